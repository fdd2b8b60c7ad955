// C ABI of libknn_b200.so and the host-side orchestration of a search.
//
// Replaces, for the flat path only, the faiss calls of the reference drivers (paths relative
// to /root/reference): faiss.normalize_L2 (cath/search.py:19), faiss.IndexFlat
// (cath/search.py:20), index.add (cath/search.py:22), index.search (cath/search.py:24,
// pfam/proteins_search.py:49, seqvec_search/main.py:45).
//
// Two device paths, both CUDA only:
//   exact  : fp32 scan kernel -> dense scores -> two-level radix select          (small nq / small N)
//   tensor : tcgen05 GEMM over 16-bit shadow copies (fp16 or bf16) with threshold-filter epilogue over growing database panels,
//            per-panel threshold tightening, exact fp32 rerank of the surviving candidates,
//            final radix select                                                   (large batches)
// The tensor path returns exactly what the exact path returns: the filter keeps a provable
// superset of the true top-k (DESIGN.md "error bound") and the rerank recomputes the scores
// with the scan kernel's arithmetic.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace knn {

static thread_local std::string g_error;
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t want) {
        if (want <= bytes) return KNN_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            return KNN_ERR_MEMORY;
        }
        bytes = want;
        return KNN_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; prev = -1; }
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace knn

using namespace knn;

struct knn_index {
    int d = 0, dp = 0, metric = 0, device = 0;
    unsigned flags = 0;
    int64_t ntotal = 0, capacity = 0;
    float* xb_f32 = nullptr;          // [capacity x dp] zero-padded rows (absent with BF16_STORAGE)
    h16_t* xb_h16 = nullptr;          // [capacity x dp] 16-bit shadow rows, format shadow_fmt (bf16 with BF16_STORAGE: the rows themselves)
    int shadow_fmt = kFmtBF16;        // format of ALL shadow rows (see desired_shadow / prepare_shadow)
    int shadow_mbits = 7;             // explicit mantissa bits kept in bf16 shadow rows
    bool stats_dirty = false;         // rows were added since the shadow's rounding error was last looked at
    bool fp16_unfit = false;          // the data leaves fp16's exponent range (measured): automatic choice stays bf16
    float* ynorm2 = nullptr;          // [capacity]
    DbStats* stats = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t add_event = nullptr;  // last ingest; searches on another stream wait on it
    GemmPlan* plan = nullptr;
    // workspaces: exact path / staging
    DevBuf stage, xq_f32, xnorm2, eps, scores, lists_s, lists_i, overflow;
    DevBuf ovf_q, ovf_idx, ovf_x, ovf_D, ovf_I;  // per-query overflow flags of a call and the repair staging
    DevBuf h_xq, h_D, h_I;
    // tensor path: per-query state of the filter (queries, thresholds, candidate lists)
    struct TensorWs {
        DevBuf xq_f32, xq_h16, xnorm2, eps, thr, counts, cand_s, cand_i;
    };
    TensorWs ws1;  // one query batch (single-call search)
    TensorWs ws2;  // all queries of a two-phase search (filter ... exchange ... finish)
    struct Pending {
        bool active = false, tensor = false;
        int64_t nq = 0, qb = 0, nbatches = 0;
        int k = 0, cap = 0;
    } pend;
    // parameters
    int path_param = 0;
    int64_t query_batch = 16384;
    int profile = 0;
    int64_t tensor_min_nq = 1, tensor_min_n = 8192;  // measured: the bf16 stream beats the fp32 scan from nq = 1 (profiles/)
    int cta_group = 2;
    int l2_hints = 0;
    int debug_skip_epilogue = 0;
    int gemm_stages = 0;
    int stream_kernel = 1;
    int panel_ratio = 0;  // 0: automatic
    int64_t small_batch_nq = 256;  // batches up to this size use growth ratio 8
    int shadow_param = 0;          // 16-bit format of the tensor-core operands: 0 automatic (desired_shadow), 1 bf16, 2 fp16
    int mbits_param = 0;           // mantissa bits kept in bf16 operands: 0 automatic, 2..7
    // statistics of the last search
    int last_path = 0;
    long long st_launches = 0, st_gemm_launches = 0, st_candidates = 0, st_overflow_batches = 0, st_overflow_queries = 0, st_rerank_pairs = 0, st_shadow_conversions = 0;
    double st_gemm_ms = 0, st_rerank_ms = 0;
    std::vector<cudaEvent_t> ev_pool;   // profile = 1: pairs of events around the GEMM launches ...
    size_t ev_used = 0;
    std::vector<cudaEvent_t> ev_pool_r; // ... and around the rerank launches
    size_t ev_used_r = 0;
};

namespace {

// profile = 1: after the stream has been synchronised, fold the event pairs into the statistics
void collect_profile(knn_index* ix) {
    for (size_t i = 0; i + 1 < ix->ev_used; i += 2) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ix->ev_pool[i], ix->ev_pool[i + 1]);
        ix->st_gemm_ms += ms;
    }
    for (size_t i = 0; i + 1 < ix->ev_used_r; i += 2) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ix->ev_pool_r[i], ix->ev_pool_r[i + 1]);
        ix->st_rerank_ms += ms;
    }
}

bool bf16_only(const knn_index* ix) { return (ix->flags & KNN_FLAG_BF16_STORAGE) != 0; }

int grow(knn_index* ix, int64_t want_rows, bool exact = false) {
    if (want_rows <= ix->capacity) return KNN_OK;
    if (want_rows >= int64_t(0xFFFFFFFFll)) {
        set_error("an index holds at most 2^32-2 rows");
        return KNN_ERR_LIMIT;
    }
    int64_t cap = exact ? want_rows : ix->capacity + ix->capacity / 2;
    if (cap < want_rows) cap = want_rows;
    if (cap < 1024) cap = 1024;
    // rows may have been ingested on a caller's stream: settle everything before moving them
    KNN_CHECK_CUDA(cudaDeviceSynchronize());
    float* nf = nullptr;
    h16_t* nb = nullptr;
    float* nn = nullptr;
    cudaError_t e = cudaSuccess;
    if (!bf16_only(ix)) e = cudaMalloc(&nf, size_t(cap) * ix->dp * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&nb, size_t(cap) * ix->dp * sizeof(h16_t));
    if (e == cudaSuccess) e = cudaMalloc(&nn, size_t(cap) * sizeof(float));
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (nf) cudaFree(nf);
        if (nb) cudaFree(nb);
        if (nn) cudaFree(nn);
        set_error("cannot allocate storage for %lld rows of %d floats: %s", (long long)cap, ix->dp,
                  cudaGetErrorString(e));
        return KNN_ERR_MEMORY;
    }
    if (ix->ntotal > 0) {
        if (nf) KNN_CHECK_CUDA(cudaMemcpyAsync(nf, ix->xb_f32, size_t(ix->ntotal) * ix->dp * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
        KNN_CHECK_CUDA(cudaMemcpyAsync(nb, ix->xb_h16, size_t(ix->ntotal) * ix->dp * sizeof(h16_t), cudaMemcpyDeviceToDevice, ix->stream));
        KNN_CHECK_CUDA(cudaMemcpyAsync(nn, ix->ynorm2, size_t(ix->ntotal) * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
        KNN_CHECK_CUDA(cudaStreamSynchronize(ix->stream));
    }
    if (ix->xb_f32) cudaFree(ix->xb_f32);
    if (ix->xb_h16) cudaFree(ix->xb_h16);
    if (ix->ynorm2) cudaFree(ix->ynorm2);
    ix->xb_f32 = nf;
    ix->xb_h16 = nb;
    ix->ynorm2 = nn;
    ix->capacity = cap;
    return KNN_OK;
}

cudaEvent_t next_event_r(knn_index* ix) {
    if (ix->ev_used_r == ix->ev_pool_r.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ix->ev_pool_r.push_back(e);
    }
    return ix->ev_pool_r[ix->ev_used_r++];
}

cudaEvent_t next_event(knn_index* ix) {
    if (ix->ev_used == ix->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ix->ev_pool.push_back(e);
    }
    return ix->ev_pool[ix->ev_used++];
}

// ---- 16-bit shadow format -------------------------------------------------------------------
const __nv_bfloat16* bf16_rows(const knn_index* ix) { return reinterpret_cast<const __nv_bfloat16*>(ix->xb_h16); }

constexpr int kAutoMantissaBits = 7;

struct ShadowChoice {
    int fmt, mbits;
};

// Which 16-bit format serves a search of the k best among n rows (k = 0: not known yet, at add time).
// Both formats run tcgen05.mma kind::f16 at the same rate per clock; they differ in two measured ways
// (profiles/r01_shadow_formats.md): fp16's 3 extra mantissa bits shrink the rounding term of the error bound 8 x,
// i.e. fewer candidates to rescore exactly - and they cost multiplier power, which under the board's power cap is
// ~6 % of the clock in a long GEMM.  So fp16 where the rescoring dominates (k large against n), bf16 where the
// GEMM does.  The shadow is rewritten from the fp32 master rows when the regime changes (one pass over the rows).
ShadowChoice desired_shadow(const knn_index* ix, int64_t k, int64_t n) {
    const int mb = ix->mbits_param ? ix->mbits_param : kAutoMantissaBits;
    if (bf16_only(ix)) return {kFmtBF16, 7};  // the bf16 values are the database
    if (ix->shadow_param == 1) return {kFmtBF16, mb};
    if (ix->shadow_param == 2) return {kFmtFP16, 10};
    const bool rescoring_heavy = k > 0 ? n < 10000 * k : n < (int64_t(1) << 20);
    if (rescoring_heavy && !ix->fp16_unfit) return {kFmtFP16, 10};
    return {kFmtBF16, mb};
}

int query_mbits(const knn_index* ix) {
    if (ix->shadow_fmt != kFmtBF16) return 10;
    return bf16_only(ix) ? (ix->mbits_param ? ix->mbits_param : kAutoMantissaBits) : ix->shadow_mbits;
}

// Rewrites every shadow row from the fp32 master rows (and re-measures the rounding error).
int convert_shadow(knn_index* ix, ShadowChoice c, cudaStream_t s) {
    if (ix->ntotal > 0) {
        KNN_CHECK_CUDA(cudaMemsetAsync(ix->stats, 0, sizeof(DbStats), s));
        KNN_CHECK(launch_ingest(ix->xb_f32, ix->dp, ix->ntotal, ix->dp, ix->dp, nullptr, ix->xb_h16, c.fmt, c.mbits, false,
                                nullptr, ix->stats, s));
        ix->st_shadow_conversions++;
    }
    ix->shadow_fmt = c.fmt;
    ix->shadow_mbits = c.mbits;
    return KNN_OK;
}

// Before a tensor-path search for the k best: bring the shadow rows into the format that serves it.  fp16 is only
// usable while the data sits inside its exponent range: after rows were added (or rewritten) the MEASURED rounding
// error is compared with what bf16 guarantees (|y - bf16(y)| <= 2^-9 |y|); when fp16 does worse than a quarter of
// that on the worst row (saturated or flushed values), the automatic choice falls back to bf16 for this database.
int prepare_shadow(knn_index* ix, int64_t k, cudaStream_t s) {
    if (bf16_only(ix) || ix->ntotal == 0) return KNN_OK;
    ShadowChoice want = desired_shadow(ix, k, ix->ntotal);
    if (want.fmt != ix->shadow_fmt || want.mbits != ix->shadow_mbits) {
        KNN_CHECK(convert_shadow(ix, want, s));
        ix->stats_dirty = true;
    }
    if (ix->stats_dirty && ix->shadow_fmt == kFmtFP16 && ix->shadow_param == 0) {
        DbStats h;
        KNN_CHECK_CUDA(cudaMemcpyAsync(&h, ix->stats, sizeof(h), cudaMemcpyDeviceToHost, s));
        KNN_CHECK_CUDA(cudaStreamSynchronize(s));
        float n2, d2;
        memcpy(&n2, &h.max_norm2, 4);
        memcpy(&d2, &h.max_dnorm2, 4);
        if (d2 > n2 * 9.5367431640625e-07f /* 2^-20 */) {
            ix->fp16_unfit = true;
            KNN_CHECK(convert_shadow(ix, desired_shadow(ix, k, ix->ntotal), s));
        }
    }
    ix->stats_dirty = false;
    return KNN_OK;
}

// ---- exact path ---------------------------------------------------------------------------
int search_exact(knn_index* ix, int64_t nq, const float* xq_dev, int k, float* D, int64_t* I, int64_t id_base,
                 cudaStream_t s) {
    const int largest = ix->metric == KNN_METRIC_INNER_PRODUCT;
    const int64_t N = ix->ntotal;
    const int seg_len = 32768;
    const int64_t nseg = (N + seg_len - 1) / seg_len;
    const int64_t list_ld = nseg * k;
    // batch sizes from fixed budgets: lists <= 1 GiB, dense scores <= 512 MiB
    int64_t qb = nq < 1024 ? nq : 1024;
    while (qb > 8 && qb * list_ld * 8 > (int64_t(1) << 30)) qb /= 2;
    int64_t chunk = ((int64_t(512) << 20) / (qb * 4)) / seg_len * seg_len;
    if (chunk < seg_len) chunk = seg_len;
    if (chunk > nseg * seg_len) chunk = nseg * seg_len;
    KNN_CHECK(ix->xq_f32.ensure(size_t(qb) * ix->dp * sizeof(float)));
    KNN_CHECK(ix->xnorm2.ensure(size_t(qb) * sizeof(float)));
    KNN_CHECK(ix->eps.ensure(size_t(qb) * sizeof(float)));
    KNN_CHECK(ix->scores.ensure(size_t(qb) * chunk * sizeof(float)));
    KNN_CHECK(ix->lists_s.ensure(size_t(qb) * list_ld * sizeof(float)));
    KNN_CHECK(ix->lists_i.ensure(size_t(qb) * list_ld * sizeof(uint32_t)));
    for (int64_t q0 = 0; q0 < nq; q0 += qb) {
        const int64_t nb = nq - q0 < qb ? nq - q0 : qb;
        KNN_CHECK(launch_prep_queries(xq_dev + q0 * ix->d, nb, nb, ix->d, ix->dp, ix->xq_f32.as<float>(), nullptr, kFmtBF16, 7,
                                      ix->xnorm2.as<float>(), ix->eps.as<float>(), ix->stats, ix->metric, s));
        for (int64_t j0 = 0; j0 < N; j0 += chunk) {
            const int64_t j1 = j0 + chunk < N ? j0 + chunk : N;
            KNN_CHECK(launch_scan_f32(ix->xq_f32.as<float>(), ix->xnorm2.as<float>(), nb, ix->dp, ix->xb_f32,
                                      bf16_rows(ix), ix->ynorm2, j0, j1, ix->metric, ix->scores.as<float>(), chunk, s));
            KNN_CHECK(launch_select_dense(ix->scores.as<float>(), chunk, j1 - j0, nb, seg_len, uint32_t(j0), k, largest,
                                          ix->lists_s.as<float>(), ix->lists_i.as<uint32_t>(), list_ld,
                                          (j0 / seg_len) * k, s));
        }
        KNN_CHECK(launch_select_final(ix->lists_s.as<float>(), ix->lists_i.as<uint32_t>(), nullptr, list_ld, list_ld,
                                      nb, k, largest, D + q0 * k, I + q0 * k, id_base, s));
    }
    return KNN_OK;
}

// ---- tensor path --------------------------------------------------------------------------
int candidate_capacity(int k) {
    int cap = 8192;
    while (cap < 8 * k) cap *= 2;
    return cap;
}

int tensor_ws_ensure(knn_index::TensorWs& W, int64_t rows, int dp, int cap) {
    KNN_CHECK(W.xq_f32.ensure(size_t(rows) * dp * sizeof(float)));
    KNN_CHECK(W.xq_h16.ensure(size_t(rows) * dp * sizeof(h16_t)));
    KNN_CHECK(W.xnorm2.ensure(size_t(rows) * sizeof(float)));
    KNN_CHECK(W.eps.ensure(size_t(rows) * sizeof(float)));
    KNN_CHECK(W.thr.ensure(size_t(rows) * sizeof(float)));
    KNN_CHECK(W.counts.ensure(size_t(rows) * sizeof(int)));
    KNN_CHECK(W.cand_s.ensure(size_t(rows) * cap * sizeof(float)));
    KNN_CHECK(W.cand_i.ensure(size_t(rows) * cap * sizeof(uint32_t)));
    return KNN_OK;
}

FilterState filter_state(knn_index::TensorWs& W, int64_t off, int cap, int* ovf = nullptr) {
    FilterState st;
    st.ovf = ovf;
    st.thr = W.thr.as<float>() + off;
    st.counts = W.counts.as<int>() + off;
    st.cand_scores = W.cand_s.as<float>() + off * cap;
    st.cand_ids = W.cand_i.as<uint32_t>() + off * cap;
    st.cap = cap;
    return st;
}

int tensor_prepare(knn_index* ix) {
    if (!ix->plan) KNN_CHECK(gemm_plan_create(&ix->plan, ix->device));
    gemm_plan_set_cta_group(ix->plan, ix->cta_group);
    gemm_plan_set_l2_hints(ix->plan, ix->l2_hints);
    gemm_plan_set_debug(ix->plan, ix->debug_skip_epilogue);
    gemm_plan_set_stages(ix->plan, ix->gemm_stages);
    gemm_plan_set_stream_kernel(ix->plan, ix->stream_kernel);
    return KNN_OK;
}

// Filter phase of one query batch (rows [off, off+nb) of W): tensor-core scores over database panels of
// growing size, survivors appended to the candidate lists, thresholds tightened after every panel.
// On return thr[q] = (k-th best approximate score) - 2 eps[q].
int tensor_filter_batch(knn_index* ix, knn_index::TensorWs& W, int64_t off, int64_t nb, const float* xq_batch, int k,
                        int cap, int* d_overflow, int* d_ovf_q, cudaStream_t s) {
    const int64_t N = ix->ntotal;
    const int64_t nb_pad = round_up(nb, 256);
    // first panel: stored densely (every score), sized so that the list it leaves fits the register-resident
    // tighten (<= 1024 entries) for small k, and holds well over k rows for large k
    int64_t first_panel = std::max<int64_t>(1024, round_up(4 * int64_t(k), 256));
    if (first_panel > cap / 2) first_panel = cap / 2;
    if (first_panel > N) first_panel = N;
    // small batches are launch-bound, not list-bound: fewer, faster growing panels (7 k' survivors per panel and query)
    // (only while the ~7 k' survivors of the second panel fit next to the dense first panel: k' ~ 1.7 k)
    const bool fast_growth = nb <= ix->small_batch_nq && 12 * int64_t(k) + first_panel <= cap;
    const int ratio = ix->panel_ratio >= 2 ? ix->panel_ratio : (fast_growth ? 8 : 2);
    FilterState st = filter_state(W, off, cap, d_ovf_q);
    float* xq_f32 = W.xq_f32.as<float>() + off * ix->dp;
    h16_t* xq_h16 = W.xq_h16.as<h16_t>() + off * ix->dp;
    const int fmt_q = ix->shadow_fmt;  // tcgen05.mma kind::f16 takes one 16-bit format per launch (mixing them is an illegal instruction)
    KNN_CHECK(launch_prep_queries(xq_batch, nb, nb_pad, ix->d, ix->dp, xq_f32, xq_h16, fmt_q, query_mbits(ix), W.xnorm2.as<float>() + off,
                                  W.eps.as<float>() + off, ix->stats, ix->metric, s));
    KNN_CHECK(launch_init_filter(st, nb, nb_pad, int(first_panel), s));
    int64_t j0 = 0;
    int panel = 0;
    while (j0 < N) {
        // rows seen after each panel: P, 4P, 8P, 16P, ... (P = first, densely stored panel).  A panel runs with the
        // threshold of its start, so it appends about (r - 1) k' survivors per query (r = growth ratio, k' = rows
        // within 2 eps of the k-th): doubling keeps that at k' per panel (measured: r = 4, 8, 16 are slower).
        const int64_t len = panel == 0 ? first_panel : (panel == 1 && ratio == 2 && ix->panel_ratio < 2 ? 3 * j0 : (ratio - 1) * j0);
        const int64_t j1 = j0 + len < N ? j0 + len : N;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (ix->profile) {
            e0 = next_event(ix);
            e1 = next_event(ix);
            cudaEventRecord(e0, s);
        }
        KNN_CHECK(gemm_filter_launch(ix->plan, xq_h16, fmt_q, nb, nb_pad, ix->dp, ix->xb_h16, ix->shadow_fmt, N, ix->ynorm2, j0, j1,
                                     ix->metric, panel == 0, st, s));
        if (ix->profile) cudaEventRecord(e1, s);
        ix->st_gemm_launches++;
        KNN_CHECK(launch_tighten(st, W.eps.as<float>() + off, nb, k, 1, nullptr, d_overflow, s));
        j0 = j1;
        ++panel;
    }
    return KNN_OK;
}

// Finish phase of one batch: exact fp32 rescoring of every candidate at or above thr, then the sorted top-k.
// `lower` (optional) is a lower bound of the TRUE k-th best score per query obtained elsewhere (other shards):
// a row can only be in the global top-k if its approximate score is >= lower - eps.
int tensor_finish_batch(knn_index* ix, knn_index::TensorWs& W, int64_t off, int64_t nb, int k, int cap, const float* lower,
                        float* D, int64_t* I, int64_t id_base, cudaStream_t s) {
    const int largest = ix->metric == KNN_METRIC_INNER_PRODUCT;
    FilterState st = filter_state(W, off, cap);
    if (lower) KNN_CHECK(launch_apply_lower(st.thr, W.eps.as<float>() + off, lower, nb, s));
    if (ix->profile) cudaEventRecord(next_event_r(ix), s);
    KNN_CHECK(launch_rerank(W.xq_f32.as<float>() + off * ix->dp, W.xnorm2.as<float>() + off, nb, ix->dp, ix->xb_f32,
                            bf16_rows(ix), ix->ynorm2, ix->metric, st.cand_scores, st.cand_ids, st.counts, st.thr, cap, s));
    if (ix->profile) cudaEventRecord(next_event_r(ix), s);
    KNN_CHECK(launch_select_final(st.cand_scores, st.cand_ids, st.counts, cap, 0, nb, k, largest, D, I, id_base, s));
    return KNN_OK;
}

int redo_overflowed(knn_index* ix, int64_t nq, int64_t qb, int64_t nbatches, const float* xq_dev, int k, float* D,
                    int64_t* I, int64_t id_base, cudaStream_t s) {
    // A candidate list that ran past its capacity (heavily duplicated / clustered scores) is never truncated
    // silently: exactly the queries it happened to are redone with the exact scan and their rows replaced.
    std::vector<int> h_overflow(size_t(nbatches), 0);
    KNN_CHECK_CUDA(cudaMemcpyAsync(h_overflow.data(), ix->overflow.p, sizeof(int) * size_t(nbatches), cudaMemcpyDeviceToHost, s));
    KNN_CHECK_CUDA(cudaStreamSynchronize(s));
    std::vector<int> redo;
    std::vector<int> flags;
    for (int64_t b = 0; b < nbatches; ++b) {
        if (!h_overflow[size_t(b)]) continue;
        const int64_t q0 = b * qb;
        const int64_t nb = nq - q0 < qb ? nq - q0 : qb;
        ix->st_overflow_batches++;
        flags.resize(size_t(nb));
        KNN_CHECK_CUDA(cudaMemcpyAsync(flags.data(), ix->ovf_q.as<int>() + q0, sizeof(int) * size_t(nb), cudaMemcpyDeviceToHost, s));
        KNN_CHECK_CUDA(cudaStreamSynchronize(s));
        for (int64_t i = 0; i < nb; ++i)
            if (flags[size_t(i)]) redo.push_back(int(q0 + i));
    }
    ix->st_overflow_queries += (long long)redo.size();
    const int64_t chunk = 1024;
    for (size_t c0 = 0; c0 < redo.size(); c0 += size_t(chunk)) {
        const int64_t n = int64_t(std::min(redo.size() - c0, size_t(chunk)));
        KNN_CHECK(ix->ovf_idx.ensure(sizeof(int) * size_t(chunk)));
        KNN_CHECK(ix->ovf_x.ensure(sizeof(float) * size_t(chunk) * ix->d));
        KNN_CHECK(ix->ovf_D.ensure(sizeof(float) * size_t(chunk) * k));
        KNN_CHECK(ix->ovf_I.ensure(sizeof(int64_t) * size_t(chunk) * k));
        KNN_CHECK_CUDA(cudaMemcpyAsync(ix->ovf_idx.p, redo.data() + c0, sizeof(int) * size_t(n), cudaMemcpyHostToDevice, s));
        KNN_CHECK(launch_gather_rows(xq_dev, ix->d, ix->ovf_idx.as<int>(), n, ix->ovf_x.as<float>(), s));
        KNN_CHECK(search_exact(ix, n, ix->ovf_x.as<float>(), k, ix->ovf_D.as<float>(), ix->ovf_I.as<int64_t>(), id_base, s));
        KNN_CHECK(launch_scatter_results(ix->ovf_D.as<float>(), ix->ovf_I.as<int64_t>(), ix->ovf_idx.as<int>(), n, k, D, I, s));
        KNN_CHECK_CUDA(cudaStreamSynchronize(s));  // `redo` chunk consumed
    }
    return KNN_OK;
}

int search_tensor(knn_index* ix, int64_t nq, const float* xq_dev, int k, float* D, int64_t* I, int64_t id_base,
                  cudaStream_t s) {
    const int cap = candidate_capacity(k);
    int64_t qb = ix->query_batch;
    if (qb > nq) qb = nq;
    qb = round_up(qb, 256);
    const int64_t nbatches = (nq + qb - 1) / qb;
    KNN_CHECK(tensor_ws_ensure(ix->ws1, qb, ix->dp, cap));
    KNN_CHECK(ix->overflow.ensure(sizeof(int) * size_t(nbatches)));
    KNN_CHECK(ix->ovf_q.ensure(sizeof(int) * size_t(nq)));
    KNN_CHECK(tensor_prepare(ix));
    KNN_CHECK_CUDA(cudaMemsetAsync(ix->overflow.p, 0, sizeof(int) * size_t(nbatches), s));
    for (int64_t b = 0; b < nbatches; ++b) {
        const int64_t q0 = b * qb;
        const int64_t nb = nq - q0 < qb ? nq - q0 : qb;
        KNN_CHECK(tensor_filter_batch(ix, ix->ws1, 0, nb, xq_dev + q0 * ix->d, k, cap, ix->overflow.as<int>() + b,
                                      ix->ovf_q.as<int>() + q0, s));
        KNN_CHECK(tensor_finish_batch(ix, ix->ws1, 0, nb, k, cap, nullptr, D + q0 * k, I + q0 * k, id_base, s));
    }
    return redo_overflowed(ix, nq, qb, nbatches, xq_dev, k, D, I, id_base, s);
}

bool use_tensor_path(const knn_index* ix, int64_t nq, int k) {
    bool tensor = ix->path_param == 2 || (ix->path_param == 0 && nq >= ix->tensor_min_nq && ix->ntotal >= ix->tensor_min_n);
    if (ix->ntotal < k) tensor = false;
    return tensor;
}

int search_dev_impl(knn_index* ix, int64_t nq, const float* xq_dev, int64_t k64, float* D, int64_t* I, int64_t id_base,
                    cudaStream_t s) {
    if (nq < 0 || k64 <= 0 || (nq > 0 && (!xq_dev || !D || !I))) {
        set_error("search: invalid arguments (nq=%lld, k=%lld)", (long long)nq, (long long)k64);
        return KNN_ERR_INVALID;
    }
    if (k64 > KNN_MAX_K) {
        set_error("search: k=%lld exceeds KNN_MAX_K=%d", (long long)k64, KNN_MAX_K);
        return KNN_ERR_LIMIT;
    }
    if (nq == 0) return KNN_OK;
    KNN_CHECK_CUDA(cudaStreamWaitEvent(s, ix->add_event, 0));
    const int k = int(k64);
    const long long launches0 = g_launches.load();
    ix->st_gemm_launches = 0;
    ix->st_gemm_ms = 0;
    ix->st_rerank_ms = 0;
    ix->ev_used_r = 0;
    ix->st_overflow_batches = 0;
    ix->st_overflow_queries = 0;
    ix->ev_used = 0;
    int rc;
    if (ix->ntotal == 0) {
        // nothing to search: all padding (select over an empty list)
        KNN_CHECK(ix->lists_s.ensure(16));
        KNN_CHECK(ix->lists_i.ensure(16));
        rc = launch_select_final(ix->lists_s.as<float>(), ix->lists_i.as<uint32_t>(), nullptr, 0, 0, nq, k,
                                 ix->metric == KNN_METRIC_INNER_PRODUCT, D, I, id_base, s);
        ix->last_path = 1;
    } else {
        const bool tensor = use_tensor_path(ix, nq, k);
        ix->last_path = tensor ? 2 : 1;
        if (tensor) KNN_CHECK(prepare_shadow(ix, k, s));
        rc = tensor ? search_tensor(ix, nq, xq_dev, k, D, I, id_base, s) : search_exact(ix, nq, xq_dev, k, D, I, id_base, s);
    }
    if (rc != KNN_OK) return rc;
    if (ix->profile && ix->ev_used) {
        KNN_CHECK_CUDA(cudaStreamSynchronize(s));
        collect_profile(ix);
    }
    ix->st_launches = g_launches.load() - launches0;
    return KNN_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

const char* knn_last_error(void) { return g_error.c_str(); }

int knn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int64_t knn_kernel_launches(void) { return g_launches.load(); }

int knn_normalize_l2_dev(float* x_dev, int64_t n, int64_t d, void* stream) {
    if (n < 0 || d <= 0 || (n > 0 && !x_dev)) {
        set_error("normalize_l2: invalid arguments");
        return KNN_ERR_INVALID;
    }
    return launch_normalize_l2(x_dev, n, d, static_cast<cudaStream_t>(stream));
}

int knn_normalize_l2(float* x, int64_t n, int64_t d, int device) {
    if (n < 0 || d <= 0 || (n > 0 && !x)) {
        set_error("normalize_l2: invalid arguments");
        return KNN_ERR_INVALID;
    }
    if (n == 0) return KNN_OK;
    DeviceGuard g(device);
    if (!g.ok) {
        set_error("normalize_l2: cannot select CUDA device %d (no CPU fallback)", device);
        return KNN_ERR_CUDA;
    }
    // stream the matrix through a bounded device buffer
    const int64_t rows_per = std::max<int64_t>(1, (int64_t(256) << 20) / (d * 4));
    float* buf = nullptr;
    const int64_t rows_alloc = n < rows_per ? n : rows_per;
    KNN_CHECK_CUDA(cudaMalloc(&buf, size_t(rows_alloc) * d * sizeof(float)));
    int rc = KNN_OK;
    for (int64_t r0 = 0; r0 < n && rc == KNN_OK; r0 += rows_per) {
        const int64_t nr = n - r0 < rows_per ? n - r0 : rows_per;
        cudaError_t e = cudaMemcpy(buf, x + r0 * d, size_t(nr) * d * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            rc = launch_normalize_l2(buf, nr, d, nullptr);
            if (rc == KNN_OK) e = cudaMemcpy(x + r0 * d, buf, size_t(nr) * d * sizeof(float), cudaMemcpyDeviceToHost);
        }
        if (e != cudaSuccess) {
            set_error("normalize_l2: copy failed: %s", cudaGetErrorString(e));
            rc = KNN_ERR_CUDA;
        }
    }
    cudaFree(buf);
    return rc;
}

int knn_index_create(knn_index** out, int d, int metric, int device, unsigned flags) {
    if (!out || d <= 0 || (metric != KNN_METRIC_INNER_PRODUCT && metric != KNN_METRIC_L2)) {
        set_error("index_create: invalid arguments (d=%d, metric=%d)", d, metric);
        return KNN_ERR_INVALID;
    }
    int ndev = knn_device_count();
    if (device < 0 || device >= ndev) {
        set_error("index_create: CUDA device %d not available (%d visible); this library has no CPU fallback", device, ndev);
        return KNN_ERR_CUDA;
    }
    DeviceGuard g(device);
    if (!g.ok) {
        set_error("index_create: cannot select device %d", device);
        return KNN_ERR_CUDA;
    }
    knn_index* ix = new knn_index();
    ix->d = d;
    ix->dp = int(round_up(d, kDimAlign));
    ix->metric = metric;
    ix->device = device;
    ix->flags = flags;
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ix->add_event, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&ix->stats, sizeof(DbStats));
    if (e == cudaSuccess) e = cudaMemset(ix->stats, 0, sizeof(DbStats));
    if (e != cudaSuccess) {
        set_error("index_create: %s", cudaGetErrorString(e));
        delete ix;
        return KNN_ERR_CUDA;
    }
    *out = ix;
    return KNN_OK;
}

int knn_index_free(knn_index* ix) {
    if (!ix) return KNN_OK;
    DeviceGuard g(ix->device);
    cudaStreamSynchronize(ix->stream);
    for (DevBuf* b : {&ix->stage, &ix->xq_f32, &ix->xnorm2, &ix->eps, &ix->scores, &ix->lists_s, &ix->lists_i,
                      &ix->overflow, &ix->ovf_q, &ix->ovf_idx, &ix->ovf_x, &ix->ovf_D, &ix->ovf_I, &ix->h_xq, &ix->h_D, &ix->h_I})
        b->release();
    for (knn_index::TensorWs* W : {&ix->ws1, &ix->ws2})
        for (DevBuf* b : {&W->xq_f32, &W->xq_h16, &W->xnorm2, &W->eps, &W->thr, &W->counts, &W->cand_s, &W->cand_i}) b->release();
    if (ix->xb_f32) cudaFree(ix->xb_f32);
    if (ix->xb_h16) cudaFree(ix->xb_h16);
    if (ix->ynorm2) cudaFree(ix->ynorm2);
    if (ix->stats) cudaFree(ix->stats);
    for (cudaEvent_t e : ix->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : ix->ev_pool_r) cudaEventDestroy(e);
    if (ix->add_event) cudaEventDestroy(ix->add_event);
    if (ix->plan) gemm_plan_destroy(ix->plan);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
    return KNN_OK;
}

int knn_index_reset(knn_index* ix) {
    if (!ix) return KNN_ERR_INVALID;
    DeviceGuard g(ix->device);
    ix->ntotal = 0;
    ix->stats_dirty = false;
    ix->fp16_unfit = false;
    KNN_CHECK_CUDA(cudaMemsetAsync(ix->stats, 0, sizeof(DbStats), ix->stream));
    KNN_CHECK_CUDA(cudaStreamSynchronize(ix->stream));
    return KNN_OK;
}

int knn_index_reserve(knn_index* ix, int64_t n) {
    if (!ix || n < 0) return KNN_ERR_INVALID;
    DeviceGuard g(ix->device);
    // exact size: reserve is how large databases avoid the 1.5x growth slack
    return grow(ix, n, /*exact=*/true);
}

int knn_index_add_dev(knn_index* ix, int64_t n, const float* x_dev, void* stream) {
    if (!ix || n < 0 || (n > 0 && !x_dev)) {
        set_error("add: invalid arguments");
        return KNN_ERR_INVALID;
    }
    if (n == 0) return KNN_OK;
    DeviceGuard g(ix->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    KNN_CHECK(grow(ix, ix->ntotal + n));
    const int64_t r0 = ix->ntotal;
    if (r0 == 0) {  // an empty index takes the format its expected size asks for; later rows follow the rows already there
        const ShadowChoice c = desired_shadow(ix, 0, std::max<int64_t>(ix->capacity, n));
        ix->shadow_fmt = c.fmt;
        ix->shadow_mbits = c.mbits;
    }
    KNN_CHECK(launch_ingest(x_dev, ix->d, n, ix->d, ix->dp, ix->xb_f32 ? ix->xb_f32 + r0 * ix->dp : nullptr,
                            ix->xb_h16 + r0 * ix->dp, ix->shadow_fmt, ix->shadow_mbits, bf16_only(ix), ix->ynorm2 + r0,
                            ix->stats, s));
    KNN_CHECK_CUDA(cudaEventRecord(ix->add_event, s));
    ix->ntotal += n;
    ix->stats_dirty = true;
    return KNN_OK;
}

int knn_index_add(knn_index* ix, int64_t n, const float* x) {
    if (!ix || n < 0 || (n > 0 && !x)) {
        set_error("add: invalid arguments");
        return KNN_ERR_INVALID;
    }
    if (n == 0) return KNN_OK;
    DeviceGuard g(ix->device);
    KNN_CHECK(grow(ix, ix->ntotal + n));
    const int64_t rows_per = std::max<int64_t>(1, (int64_t(256) << 20) / (int64_t(ix->d) * 4));
    KNN_CHECK(ix->stage.ensure(size_t(n < rows_per ? n : rows_per) * ix->d * sizeof(float)));
    for (int64_t r0 = 0; r0 < n; r0 += rows_per) {
        const int64_t nr = n - r0 < rows_per ? n - r0 : rows_per;
        KNN_CHECK_CUDA(cudaMemcpyAsync(ix->stage.p, x + r0 * ix->d, size_t(nr) * ix->d * sizeof(float), cudaMemcpyHostToDevice, ix->stream));
        KNN_CHECK(knn_index_add_dev(ix, nr, ix->stage.as<float>(), ix->stream));
        KNN_CHECK_CUDA(cudaStreamSynchronize(ix->stream));  // the staging buffer is reused; `add` copies (faiss semantics)
    }
    return KNN_OK;
}

int64_t knn_index_ntotal(const knn_index* ix) { return ix ? ix->ntotal : -1; }
int knn_index_d(const knn_index* ix) { return ix ? ix->d : -1; }
int knn_index_metric(const knn_index* ix) { return ix ? ix->metric : -1; }

int knn_index_search_dev(knn_index* ix, int64_t nq, const float* xq_dev, int64_t k, float* D_dev, int64_t* I_dev,
                         int64_t id_base, void* stream) {
    if (!ix) {
        set_error("search: null index");
        return KNN_ERR_INVALID;
    }
    DeviceGuard g(ix->device);
    return search_dev_impl(ix, nq, xq_dev, k, D_dev, I_dev, id_base, static_cast<cudaStream_t>(stream));
}

int knn_index_search(knn_index* ix, int64_t nq, const float* xq, int64_t k, float* D, int64_t* I) {
    if (!ix || nq < 0 || k <= 0 || (nq > 0 && (!xq || !D || !I))) {
        set_error("search: invalid arguments (nq=%lld, k=%lld)", (long long)nq, (long long)k);
        return KNN_ERR_INVALID;
    }
    if (k > KNN_MAX_K) {
        set_error("search: k=%lld exceeds KNN_MAX_K=%d", (long long)k, KNN_MAX_K);
        return KNN_ERR_LIMIT;
    }
    if (nq == 0) return KNN_OK;
    DeviceGuard g(ix->device);
    // host batches bound the device staging buffers; each is H2D -> search -> D2H on one stream
    const int64_t hb = 65536;
    const int64_t nb_max = nq < hb ? nq : hb;
    KNN_CHECK(ix->h_xq.ensure(size_t(nb_max) * ix->d * sizeof(float)));
    KNN_CHECK(ix->h_D.ensure(size_t(nb_max) * k * sizeof(float)));
    KNN_CHECK(ix->h_I.ensure(size_t(nb_max) * k * sizeof(int64_t)));
    long long launches = 0, gemm_launches = 0, overflow = 0, overflow_q = 0;
    double gemm_ms = 0;
    for (int64_t q0 = 0; q0 < nq; q0 += hb) {
        const int64_t nb = nq - q0 < hb ? nq - q0 : hb;
        KNN_CHECK_CUDA(cudaMemcpyAsync(ix->h_xq.p, xq + q0 * ix->d, size_t(nb) * ix->d * sizeof(float), cudaMemcpyHostToDevice, ix->stream));
        KNN_CHECK(search_dev_impl(ix, nb, ix->h_xq.as<float>(), k, ix->h_D.as<float>(), ix->h_I.as<int64_t>(), 0, ix->stream));
        KNN_CHECK_CUDA(cudaMemcpyAsync(D + q0 * k, ix->h_D.p, size_t(nb) * k * sizeof(float), cudaMemcpyDeviceToHost, ix->stream));
        KNN_CHECK_CUDA(cudaMemcpyAsync(I + q0 * k, ix->h_I.p, size_t(nb) * k * sizeof(int64_t), cudaMemcpyDeviceToHost, ix->stream));
        KNN_CHECK_CUDA(cudaStreamSynchronize(ix->stream));
        launches += ix->st_launches;
        gemm_launches += ix->st_gemm_launches;
        gemm_ms += ix->st_gemm_ms;
        overflow += ix->st_overflow_batches;
        overflow_q += ix->st_overflow_queries;
    }
    ix->st_launches = launches;
    ix->st_gemm_launches = gemm_launches;
    ix->st_gemm_ms = gemm_ms;
    ix->st_overflow_batches = overflow;
    ix->st_overflow_queries = overflow_q;
    return KNN_OK;
}

int knn_index_search_filter_dev(knn_index* ix, int64_t nq, const float* xq_dev, int64_t k64, float* lower_dev, int64_t j,
                                float* lower_j_dev, void* stream) {
    if (!ix || nq <= 0 || k64 <= 0 || !xq_dev || !lower_dev) {
        set_error("search_filter: invalid arguments");
        return KNN_ERR_INVALID;
    }
    if (k64 > KNN_MAX_K) {
        set_error("search_filter: k=%lld exceeds KNN_MAX_K=%d", (long long)k64, KNN_MAX_K);
        return KNN_ERR_LIMIT;
    }
    if (nq > (int64_t(1) << 17)) {
        set_error("search_filter: at most 131072 queries per two-phase search (candidate lists stay resident)");
        return KNN_ERR_LIMIT;
    }
    DeviceGuard g(ix->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int k = int(k64);
    KNN_CHECK_CUDA(cudaStreamWaitEvent(s, ix->add_event, 0));
    ix->st_gemm_launches = 0;
    ix->st_gemm_ms = 0;
    ix->st_rerank_ms = 0;
    ix->ev_used_r = 0;
    ix->st_overflow_batches = 0;
    ix->st_overflow_queries = 0;
    ix->ev_used = 0;
    auto& P = ix->pend;
    P = knn_index::Pending();
    P.nq = nq;
    P.k = k;
    P.tensor = ix->ntotal > 0 && use_tensor_path(ix, nq, k);
    ix->last_path = P.tensor ? 2 : 1;
    if (!P.tensor) {  // exact path: nothing to filter, no bound to offer
        KNN_CHECK(launch_fill_f32(lower_dev, nq, -FLT_MAX, s));
        if (lower_j_dev) KNN_CHECK(launch_fill_f32(lower_j_dev, nq, -FLT_MAX, s));
        P.active = true;
        return KNN_OK;
    }
    KNN_CHECK(prepare_shadow(ix, k, s));
    P.cap = candidate_capacity(k);
    P.qb = round_up(std::min<int64_t>(ix->query_batch, nq), 256);
    P.nbatches = (nq + P.qb - 1) / P.qb;
    KNN_CHECK(tensor_ws_ensure(ix->ws2, P.nbatches * P.qb, ix->dp, P.cap));
    KNN_CHECK(ix->overflow.ensure(sizeof(int) * size_t(P.nbatches)));
    KNN_CHECK(ix->ovf_q.ensure(sizeof(int) * size_t(nq)));
    KNN_CHECK(tensor_prepare(ix));
    KNN_CHECK_CUDA(cudaMemsetAsync(ix->overflow.p, 0, sizeof(int) * size_t(P.nbatches), s));
    for (int64_t b = 0; b < P.nbatches; ++b) {
        const int64_t q0 = b * P.qb;
        const int64_t nb = nq - q0 < P.qb ? nq - q0 : P.qb;
        KNN_CHECK(tensor_filter_batch(ix, ix->ws2, q0, nb, xq_dev + q0 * ix->d, k, P.cap, ix->overflow.as<int>() + b,
                                      ix->ovf_q.as<int>() + q0, s));
    }
    // lower[q] = thr + eps = (k-th best approximate score) - eps: a lower bound of the true k-th best score
    KNN_CHECK(launch_export_lower(ix->ws2.thr.as<float>(), ix->ws2.eps.as<float>(), nq, lower_dev, s));
    if (lower_j_dev) {
        // second bound: (j-th best approximate score) - eps, j <= k.  With G shards and j = ceil(k / G) the MIN of
        // it over the shards is a lower bound of the global k-th best (every shard holds j rows at or above it).
        if (j < 1 || j > k) {
            set_error("search_filter: j must be in [1, k]");
            return KNN_ERR_INVALID;
        }
        for (int64_t b = 0; b < P.nbatches; ++b) {
            const int64_t q0 = b * P.qb;
            const int64_t nb = nq - q0 < P.qb ? nq - q0 : P.qb;
            KNN_CHECK(launch_kth_lower(filter_state(ix->ws2, q0, P.cap), ix->ws2.eps.as<float>() + q0, nb, int(j), lower_j_dev + q0, s));
        }
    }
    P.active = true;
    return KNN_OK;
}

int knn_index_search_finish_dev(knn_index* ix, int64_t nq, const float* xq_dev, int64_t k64, const float* lower_dev,
                                float* D_dev, int64_t* I_dev, int64_t id_base, void* stream) {
    if (!ix || !xq_dev || !D_dev || !I_dev) {
        set_error("search_finish: invalid arguments");
        return KNN_ERR_INVALID;
    }
    auto& P = ix->pend;
    if (!P.active || P.nq != nq || P.k != int(k64)) {
        set_error("search_finish: no matching search_filter call is pending");
        return KNN_ERR_INVALID;
    }
    DeviceGuard g(ix->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    P.active = false;
    if (!P.tensor) return search_dev_impl(ix, nq, xq_dev, k64, D_dev, I_dev, id_base, s);
    const long long launches0 = g_launches.load();
    for (int64_t b = 0; b < P.nbatches; ++b) {
        const int64_t q0 = b * P.qb;
        const int64_t nb = nq - q0 < P.qb ? nq - q0 : P.qb;
        KNN_CHECK(tensor_finish_batch(ix, ix->ws2, q0, nb, P.k, P.cap, lower_dev ? lower_dev + q0 : nullptr, D_dev + q0 * P.k,
                                      I_dev + q0 * P.k, id_base, s));
    }
    KNN_CHECK(redo_overflowed(ix, nq, P.qb, P.nbatches, xq_dev, P.k, D_dev, I_dev, id_base, s));
    if (ix->profile && ix->ev_used) {
        KNN_CHECK_CUDA(cudaStreamSynchronize(s));
        collect_profile(ix);
    }
    ix->st_launches = g_launches.load() - launches0;
    return KNN_OK;
}

int knn_index_reconstruct(knn_index* ix, int64_t i0, int64_t n, float* out) {
    if (!ix || i0 < 0 || n < 0 || i0 + n > ix->ntotal || (n > 0 && !out)) {
        set_error("reconstruct: invalid range");
        return KNN_ERR_INVALID;
    }
    if (n == 0) return KNN_OK;
    DeviceGuard g(ix->device);
    KNN_CHECK_CUDA(cudaStreamSynchronize(ix->stream));
    if (ix->xb_f32) {
        KNN_CHECK_CUDA(cudaMemcpy2D(out, size_t(ix->d) * sizeof(float), ix->xb_f32 + i0 * ix->dp, size_t(ix->dp) * sizeof(float),
                                    size_t(ix->d) * sizeof(float), size_t(n), cudaMemcpyDeviceToHost));
    } else {
        std::vector<uint16_t> tmp(size_t(n) * ix->d);
        KNN_CHECK_CUDA(cudaMemcpy2D(tmp.data(), size_t(ix->d) * 2, ix->xb_h16 + i0 * ix->dp, size_t(ix->dp) * 2,
                                    size_t(ix->d) * 2, size_t(n), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tmp.size(); ++i) {
            uint32_t u = uint32_t(tmp[i]) << 16;
            memcpy(out + i, &u, 4);
        }
    }
    return KNN_OK;
}

int knn_merge_topk_dev(int metric, int64_t nq, int64_t k, int nlists, const float* D_lists_dev, const int64_t* I_lists_dev,
                       float* D_out_dev, int64_t* I_out_dev, void* stream) {
    if (nq < 0 || k <= 0 || nlists <= 0 || (nq > 0 && (!D_lists_dev || !I_lists_dev || !D_out_dev || !I_out_dev))) {
        set_error("merge: invalid arguments");
        return KNN_ERR_INVALID;
    }
    if (k > KNN_MAX_K) {
        set_error("merge: k=%lld exceeds KNN_MAX_K=%d", (long long)k, KNN_MAX_K);
        return KNN_ERR_LIMIT;
    }
    return launch_merge_lists(D_lists_dev, I_lists_dev, nlists, nq, int(k), metric == KNN_METRIC_INNER_PRODUCT, D_out_dev,
                              I_out_dev, static_cast<cudaStream_t>(stream));
}

int knn_index_set_param(knn_index* ix, const char* name, int64_t value) {
    if (!ix || !name) return KNN_ERR_INVALID;
    std::string n(name);
    if (n == "path" && value >= 0 && value <= 2) ix->path_param = int(value);
    else if (n == "query_batch" && value >= 128) ix->query_batch = round_up(value, 256);
    else if (n == "profile") ix->profile = value != 0;
    else if (n == "cta_group" && (value == 1 || value == 2)) ix->cta_group = int(value);
    else if (n == "l2_hints") ix->l2_hints = value != 0;
    else if (n == "gemm_stages" && value >= 0 && value <= 6) ix->gemm_stages = int(value);
    else if (n == "stream_kernel") ix->stream_kernel = value != 0;
    else if (n == "panel_ratio" && value >= 0 && value <= 64) ix->panel_ratio = int(value);
    else if (n == "small_batch_nq" && value >= 0) ix->small_batch_nq = value;
    else if (n == "debug_skip_epilogue") ix->debug_skip_epilogue = int(value);
    else if (n == "shadow_fmt" && value >= 0 && value <= 2) {  // takes effect at the next search (prepare_shadow)
        if (bf16_only(ix) && value == 2) {
            set_error("set_param: an index with bf16 storage keeps bf16 rows");
            return KNN_ERR_INVALID;
        }
        ix->shadow_param = int(value);
    }
    else if (n == "mantissa_bits" && (value == 0 || (value >= 2 && value <= 7))) ix->mbits_param = int(value);
    else if (n == "tensor_min_nq" && value >= 1) ix->tensor_min_nq = value;
    else if (n == "tensor_min_n" && value >= 1) ix->tensor_min_n = value;
    else {
        set_error("set_param: unknown parameter or bad value: %s=%lld", name, (long long)value);
        return KNN_ERR_INVALID;
    }
    return KNN_OK;
}

int knn_index_get_stat(const knn_index* ix, const char* name, double* out) {
    if (!ix || !name || !out) return KNN_ERR_INVALID;
    std::string n(name);
    if (n == "path") *out = ix->last_path;
    else if (n == "launches") *out = double(ix->st_launches);
    else if (n == "gemm_launches") *out = double(ix->st_gemm_launches);
    else if (n == "gemm_ms") *out = ix->st_gemm_ms;
    else if (n == "rerank_ms") *out = ix->st_rerank_ms;
    else if (n == "overflow_batches") *out = double(ix->st_overflow_batches);
    else if (n == "overflow_queries") *out = double(ix->st_overflow_queries);
    else if (n == "capacity") *out = double(ix->capacity);
    else if (n == "shadow_fmt") *out = ix->shadow_fmt == kFmtFP16 ? 2 : 1;
    else if (n == "mantissa_bits") *out = ix->shadow_fmt == kFmtFP16 ? 10 : ix->shadow_mbits;
    else if (n == "shadow_conversions") *out = double(ix->st_shadow_conversions);
    else {
        set_error("get_stat: unknown statistic %s", name);
        return KNN_ERR_INVALID;
    }
    return KNN_OK;
}

}  // extern "C"
