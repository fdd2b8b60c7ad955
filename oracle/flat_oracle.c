/*
 * Plain-C restatement of the flat kNN path (TEST INFRASTRUCTURE ONLY - never linked
 * into libknn_b200.so; used by tests/ to cross-check oracle/flat_oracle.py and by
 * bench.py's cpu_baseline leg as a scalar "port").
 *
 * Follows faiss-cpu 1.7.2 (third-party, pinned at /root/reference/poetry.lock:100-101,
 * source not vendored; restated from its published algorithm):
 *   - faiss.normalize_L2            -> oracle_normalize_l2   (call sites cath/search.py:19,
 *                                      seqvec_search/main.py:31,34, pfam/proteins_search.py:22)
 *   - IndexFlat::search, IP and L2  -> oracle_knn_flat       (cath/search.py:24,
 *                                      pfam/proteins_search.py:49, seqvec_search/main.py:45)
 * Per query: one pass over the database, a binary heap of the k best kept in
 * (score, id) order, then a heap sort so the output is best first.  A candidate
 * replaces the heap top only when it is strictly better, or equal with a lower id,
 * which makes the tie rule "lower id first" explicit.  Squared L2 uses the BLAS-path
 * formula |x|^2 + |y|^2 - 2<x,y> clamped at 0 (the reference's L2 workload,
 * cath/search.py:32, has nq >= 20 and therefore takes that path in faiss).
 * Parity status: see the header of flat_oracle.py ("PINNED"/"UNPINNED").
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

#define METRIC_INNER_PRODUCT 0
#define METRIC_L2 1

static float dot_f32(const float* a, const float* b, long d) {
    /* eight partial sums, the shape of an AVX fvec_inner_product */
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long i = 0;
    for (; i + 8 <= d; i += 8)
        for (int l = 0; l < 8; ++l) acc[l] += a[i + l] * b[i + l];
    float s = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
    for (; i < d; ++i) s += a[i] * b[i];
    return s;
}

void oracle_normalize_l2(float* x, long n, long d) {
    for (long i = 0; i < n; ++i) {
        float* r = x + i * d;
        float nr = dot_f32(r, r, d);
        if (nr > 0) {
            const float inv = 1.0f / sqrtf(nr);
            for (long j = 0; j < d; ++j) r[j] *= inv;
        }
    }
}

/* "a is worse than b" in best-first order; key = score for IP (larger is better),
 * -dist for L2, ties resolved by the lower id being better. */
static inline int worse(float ka, int64_t ia, float kb, int64_t ib) {
    return ka < kb || (ka == kb && ia > ib);
}

static void sift_down(float* hk, int64_t* hi, long n, long p) {
    for (;;) {
        long c = 2 * p + 1;
        if (c >= n) return;
        if (c + 1 < n && worse(hk[c + 1], hi[c + 1], hk[c], hi[c])) c++;
        if (!worse(hk[c], hi[c], hk[p], hi[p])) return;
        float tk = hk[p]; hk[p] = hk[c]; hk[c] = tk;
        int64_t ti = hi[p]; hi[p] = hi[c]; hi[c] = ti;
        p = c;
    }
}

typedef struct {
    const float *xq, *xb, *ynorm;
    long q0, q1, nb, d, k;
    int metric;
    float* D;
    int64_t* I;
} job_t;

static void* knn_range(void* arg) {
    const job_t* jb = (const job_t*)arg;
    const long nb = jb->nb, d = jb->d, k = jb->k;
    const int metric = jb->metric;
    float* hk = (float*)malloc(sizeof(float) * (size_t)k);
    int64_t* hi = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    for (long q = jb->q0; q < jb->q1; ++q) {
        const float* x = jb->xq + q * d;
        long hn = 0; /* root (index 0) = worst kept element */
        const float xnorm = metric == METRIC_L2 ? dot_f32(x, x, d) : 0.f;
        for (long j = 0; j < nb; ++j) {
            float ip = dot_f32(x, jb->xb + j * d, d);
            float key;
            if (metric == METRIC_L2) {
                float dis = xnorm + jb->ynorm[j] - 2.0f * ip;
                if (dis < 0) dis = 0;
                key = -dis;
            } else {
                key = ip;
            }
            if (key != key) continue; /* NaN never enters the heap */
            if (hn < k) {
                long p = hn++;
                hk[p] = key; hi[p] = j;
                while (p > 0) {
                    long par = (p - 1) / 2;
                    if (!worse(hk[p], hi[p], hk[par], hi[par])) break;
                    float tk = hk[p]; hk[p] = hk[par]; hk[par] = tk;
                    int64_t ti = hi[p]; hi[p] = hi[par]; hi[par] = ti;
                    p = par;
                }
            } else if (worse(hk[0], hi[0], key, j)) {
                hk[0] = key; hi[0] = j;
                sift_down(hk, hi, hn, 0);
            }
        }
        /* heap sort: repeatedly move the worst to the end -> best first */
        for (long n = hn; n > 1; --n) {
            float tk = hk[0]; hk[0] = hk[n - 1]; hk[n - 1] = tk;
            int64_t ti = hi[0]; hi[0] = hi[n - 1]; hi[n - 1] = ti;
            sift_down(hk, hi, n - 1, 0);
        }
        for (long r = 0; r < k; ++r) {
            if (r < hn) {
                jb->D[q * k + r] = metric == METRIC_L2 ? -hk[r] : hk[r];
                jb->I[q * k + r] = hi[r];
            } else { /* heap neutral element, label -1 */
                jb->D[q * k + r] = metric == METRIC_L2 ? FLT_MAX : -FLT_MAX;
                jb->I[q * k + r] = -1;
            }
        }
    }
    free(hk);
    free(hi);
    return NULL;
}

/* nthreads <= 1: scalar, in the calling thread; otherwise queries are split statically
 * over pthreads (faiss parallelises the same loop over queries with OpenMP). */
void oracle_knn_flat(const float* xq, const float* xb, long nq, long nb, long d, long k,
                     int metric, float* D, int64_t* I, int nthreads) {
    float* ynorm = NULL;
    if (metric == METRIC_L2) {
        ynorm = (float*)malloc(sizeof(float) * (size_t)(nb > 0 ? nb : 1));
        for (long j = 0; j < nb; ++j) ynorm[j] = dot_f32(xb + j * d, xb + j * d, d);
    }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > nq) nthreads = nq > 0 ? (int)nq : 1;
    job_t jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) {
        job_t jb = {xq, xb, ynorm, nq * t / nthreads, nq * (t + 1) / nthreads, nb, d, k, metric, D, I};
        jobs[t] = jb;
    }
    if (nthreads == 1) {
        knn_range(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, knn_range, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
    free(ynorm);
}
