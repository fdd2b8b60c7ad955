// What the reference does with (D, I) right after index.search, on the device (SURVEY.md section 8,
// rows f3 and f4): label gathers + AUC1 / TP counting, self-hit removal, and the MMseqs2 prefilter-database
// text writer.  All of it is integer / byte work bounded by HBM traffic: one pass over (D, I), coalesced
// along the hit dimension, label tables gathered through L2.
//
// Reference code replaced (paths relative to /root/reference):
//   seqvec_search/main.py:53-82                      evaluate_faiss + evaluate      -> knn_eval_family_dev
//   cath/cath.py:76-84                               compute_is_correct             -> knn_eval_levels_dev
//   pfam/proteins.py:201-207                         compute_correctness_array      -> knn_eval_sets_dev (correct)
//   pfam/proteins_shared.py:139-157                  compute_auc1                   -> knn_eval_sets_dev (lead)
//   pfam/proteins.py:85-122                          remove_self_hit                -> knn_remove_self_hit_dev
//   seqvec_search/mmseqs/_write_prefilter_db.py:52-97 write_prefilter_db            -> knn_prefilter_measure_dev / _emit_dev
#include "common.cuh"

namespace knn {
namespace {

// Python / numpy index semantics of the reference's gathers: a negative id counts from the end.
__device__ __forceinline__ int64_t wrap_id(int64_t id, int64_t n) { return id < 0 ? id + n : id; }

// ---- AUC1 / TP with one family label per row (seqvec_search/main.py:64-82) --------------------------------
// One warp per query.  lead = length of the leading run of hits of the query's family, tp = their number.
__global__ void __launch_bounds__(256)
eval_family_kernel(const int64_t* __restrict__ I, int64_t nq, int k, const int32_t* __restrict__ query_family,
                   const int32_t* __restrict__ db_family, int64_t n_db, int32_t* __restrict__ lead,
                   int32_t* __restrict__ tp, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t q = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int32_t want = query_family[q];
    const int64_t* row = I + q * k;
    int run = 0, hits = 0;
    bool open = true;
    for (int h0 = 0; h0 < k; h0 += 32) {
        const int h = h0 + lane;
        bool in = h < k, match = false;
        if (in) {
            const int64_t id = wrap_id(row[h], n_db);
            if (id < 0 || id >= n_db) atomicOr(err, 2);
            else match = db_family[id] == want;
        }
        const unsigned act = __ballot_sync(0xffffffffu, in);
        const unsigned m = __ballot_sync(0xffffffffu, match);
        hits += __popc(m);
        if (open) {
            const unsigned miss = act & ~m;
            if (miss) {
                run += __ffs(int(miss)) - 1;
                open = false;
            } else {
                run += __popc(act);
            }
        }
    }
    if (lane == 0) {
        lead[q] = run;
        tp[q] = hits;
    }
}

// ---- per-level label equality (cath/cath.py:76-84) -----------------------------------------------------------
// out[q][l][h] = mapping[q][l] == mapping[I[q][h]][l]; one thread per (q, h), writes coalesced along h.
__global__ void __launch_bounds__(256)
eval_levels_kernel(const int64_t* __restrict__ I, int64_t q0, int k, const int32_t* __restrict__ mapping, int levels,
                   int64_t n_db, uint8_t* __restrict__ out, int* __restrict__ err) {
    const int64_t q = q0 + blockIdx.y;
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= k) return;
    const int64_t id = wrap_id(I[q * k + h], n_db);
    const bool ok = id >= 0 && id < n_db;
    if (!ok) atomicOr(err, 2);
    const int32_t* mq = mapping + q * levels;
    const int32_t* mh = mapping + (ok ? id : 0) * levels;
    for (int l = 0; l < levels; ++l) out[(q * levels + l) * k + h] = ok && (__ldg(mq + l) == __ldg(mh + l));
}

// ---- membership in a per-query set of rows (pfam/proteins.py:201-207, proteins_shared.py:139-157) -----------
// Sets in CSR form with sorted members.  One warp per query, one binary search per hit.
__global__ void __launch_bounds__(256)
eval_sets_kernel(const int64_t* __restrict__ I, int64_t nq, int k, const int64_t* __restrict__ offsets,
                 const int64_t* __restrict__ members, int64_t n_db_wrap, uint8_t* __restrict__ correct,
                 int32_t* __restrict__ lead) {
    const int lane = threadIdx.x & 31;
    const int64_t q = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int64_t s0 = offsets[q], s1 = offsets[q + 1];
    const int64_t* row = I + q * k;
    int run = 0;
    bool open = true;
    for (int h0 = 0; h0 < k; h0 += 32) {
        const int h = h0 + lane;
        const bool in = h < k;
        bool plain = false, wrapped = false;
        if (in) {
            const int64_t id = row[h];
            auto member = [&](int64_t v) {
                int64_t lo = s0, hi = s1;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (__ldg(members + mid) < v) lo = mid + 1;
                    else hi = mid;
                }
                return lo < s1 && __ldg(members + lo) == v;
            };
            plain = member(id);                        // `hit in all_correct`: plain value membership
            wrapped = (id < 0 && n_db_wrap > 0) ? member(id + n_db_wrap) : plain;  // target_ids[hit] wraps
            if (correct) correct[q * k + h] = plain;
        }
        if (lead && open) {
            const unsigned act = __ballot_sync(0xffffffffu, in);
            const unsigned m = __ballot_sync(0xffffffffu, wrapped);
            const unsigned miss = act & ~m;
            if (miss) {
                run += __ffs(int(miss)) - 1;
                open = false;
            } else {
                run += __popc(act);
            }
        }
    }
    if (lead && lane == 0) lead[q] = run;
}

// ---- self-hit removal (pfam/proteins.py:85-122) ----------------------------------------------------------------
// In place, one warp per row: p = first column holding the query's own id (k - 1 and "missing" when absent);
// columns [0, p) move one to the right and the self hit goes to column 0.  The caller then drops column 0.
template <typename T>
__device__ __forceinline__ void rotate_right_to(T* row, int p, int lane) {
    const T self = row[p];
    __syncwarp();
    for (int hi = p; hi >= 1; hi -= 32) {  // columns (hi - 32, hi] take the value of their left neighbour
        const int j = hi - lane;
        T v{};
        if (j >= 1) v = row[j - 1];
        __syncwarp();
        if (j >= 1) row[j] = v;
        __syncwarp();
    }
    if (lane == 0) row[0] = self;
}

__global__ void __launch_bounds__(256)
remove_self_hit_kernel(int64_t* I, float* D, int64_t nq, int k,
                       const int64_t* __restrict__ self_ids, unsigned long long* __restrict__ n_missing) {
    const int lane = threadIdx.x & 31;
    const int64_t q = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int64_t self = self_ids ? self_ids[q] : q;
    int64_t* row = I + q * k;
    int p = -1;
    for (int h0 = 0; h0 < k && p < 0; h0 += 32) {
        const int h = h0 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, h < k && row[h] == self);
        if (m) p = h0 + __ffs(int(m)) - 1;
    }
    if (p == 0) return;
    if (p < 0) {
        p = k - 1;
        if (lane == 0) atomicAdd(n_missing, 1ull);
    }
    rotate_right_to(row, p, lane);
    if (D) rotate_right_to(D + q * k, p, lane);
}

// ---- MMseqs2 prefilter database text (seqvec_search/mmseqs/_write_prefilter_db.py:52-97) ---------------------
// Per hit with id != -1 one line "<train_map[hit]>\t<int(clip(score) * 100)>\t0\n", per query a NUL terminator;
// per query one index line "<test_map[query]>\t<offset>\t<length>\n".
typedef unsigned __int128 u128;

__device__ __forceinline__ int dec_digits_u32(uint32_t v) {
    return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6
         : v < 10000000u ? 7 : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}
__device__ __forceinline__ int dec_digits_u64(unsigned long long v) {
    if (v < 4294967296ull) return dec_digits_u32(uint32_t(v));  // ids and cosine scores live here: 32-bit arithmetic only
    int n = 9;
    v /= 1000000000ull;
    while (v >= 10ull) { v /= 10ull; ++n; }
    return n + 1;
}
// writes the decimal digits of v so that the last one lands at end[-1]; returns the number written
__device__ __forceinline__ int put_u32(uint8_t* end, uint32_t v) {
    int n = 0;
    do {
        *--end = uint8_t('0' + v % 10u);
        v /= 10u;
        ++n;
    } while (v);
    return n;
}
__device__ __forceinline__ int put_u64(uint8_t* end, unsigned long long v) {
    if (v < 4294967296ull) return put_u32(end, uint32_t(v));
    int n = 0;
    do {
        *--end = uint8_t('0' + v % 10ull);
        v /= 10ull;
        ++n;
    } while (v);
    return n;
}
constexpr unsigned long long kTen19 = 10000000000000000000ull;

struct BigInt {   // sign and magnitude of a truncated float32 (|v| < 2^128)
    bool neg;
    u128 mag;
};
// int(float32) of Python: truncation toward zero, exact for every finite value
__device__ __forceinline__ BigInt trunc_f32(float v) {
    const uint32_t u = __float_as_uint(v);
    BigInt b;
    b.neg = (u >> 31) != 0;
    const int e = int((u >> 23) & 0xFF);
    const uint32_t m = (u & 0x7FFFFFu) | (e ? 0x800000u : 0u);
    const int sh = (e ? e : 1) - 150;  // value = m * 2^sh
    if (sh >= 0) b.mag = u128(m) << sh;
    else b.mag = sh <= -32 ? u128(0) : u128(m >> (-sh));
    if (b.mag == 0) b.neg = false;  // int(-0.5) == 0 prints "0"
    return b;
}
__device__ __forceinline__ int dec_len(const BigInt& b) {
    if ((b.mag >> 64) == 0) return (b.neg ? 1 : 0) + dec_digits_u64((unsigned long long)b.mag);  // no 128-bit division
    const unsigned long long hi = (unsigned long long)(b.mag / kTen19);
    const unsigned long long lo = (unsigned long long)(b.mag % kTen19);
    return (b.neg ? 1 : 0) + (hi ? dec_digits_u64(hi) + 19 : dec_digits_u64(lo));
}
__device__ __forceinline__ void put_big(uint8_t* end, const BigInt& b, int len) {
    if ((b.mag >> 64) == 0) {
        end -= put_u64(end, (unsigned long long)b.mag);
        if (b.neg) *--end = '-';
        return;
    }
    const unsigned long long hi = (unsigned long long)(b.mag / kTen19);
    unsigned long long lo = (unsigned long long)(b.mag % kTen19);
    if (hi) {
        for (int i = 0; i < 19; ++i) {
            *--end = uint8_t('0' + lo % 10ull);
            lo /= 10ull;
        }
        end -= put_u64(end, hi);
    } else {
        end -= put_u64(end, lo);
    }
    if (b.neg) *--end = '-';
    (void)len;
}
__device__ __forceinline__ int dec_len_i64(int64_t v) {
    return v < 0 ? 1 + dec_digits_u64(0ull - (unsigned long long)v) : dec_digits_u64((unsigned long long)v);
}
__device__ __forceinline__ void put_i64(uint8_t* end, int64_t v) {
    if (v < 0) {
        end -= put_u64(end, 0ull - (unsigned long long)v);
        *--end = '-';
    } else {
        put_u64(end, (unsigned long long)v);
    }
}
// numpy.clip(scores, -(10**30), 10**30) * 100 in float32 (numpy >= 2 keeps the array's dtype)
__device__ __forceinline__ float clip_times_100(float s, int clip) {
    if (clip && s == s) s = fminf(fmaxf(s, -1.0e30f), 1.0e30f);  // numpy.clip propagates NaN
    return __fmul_rn(s, 100.0f);
}

struct HitLine {
    int64_t id;   // translated id
    BigInt score;
    int len;      // 0: skipped hit
};
__device__ __forceinline__ HitLine make_line(int64_t hit, float score, const int64_t* __restrict__ train_map,
                                             int64_t n_train, int clip, int* err) {
    HitLine L;
    L.len = 0;
    L.id = 0;
    L.score.neg = false;
    L.score.mag = 0;
    if (hit == -1) return L;
    const int64_t w = wrap_id(hit, n_train);
    if (w < 0 || w >= n_train) {
        atomicOr(err, 2);  // IndexError in the reference
        return L;
    }
    const float v = clip_times_100(score, clip);
    if (v != v) {
        atomicOr(err, 1);  // ValueError: cannot convert float NaN to integer
        return L;
    }
    if (isinf(v)) {
        atomicOr(err, 4);  // OverflowError: cannot convert float infinity to integer (clip=False only)
        return L;
    }
    L.id = __ldg(train_map + w);
    L.score = trunc_f32(v);
    L.len = dec_len_i64(L.id) + 1 + dec_len(L.score) + 3;
    return L;
}

// section length of every query: sum of its line lengths + 1 (NUL).  One warp per query.
__global__ void __launch_bounds__(256)
prefilter_measure_kernel(const int64_t* __restrict__ I, const float* __restrict__ D, int64_t nq, int k,
                         const int64_t* __restrict__ train_map, int64_t n_train, int clip,
                         int64_t* __restrict__ sec_len, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t q = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    int total = 0;
    for (int h = lane; h < k; h += 32) total += make_line(I[q * k + h], D[q * k + h], train_map, n_train, clip, err).len;
    total = __reduce_add_sync(0xffffffffu, total);
    if (lane == 0) sec_len[q] = int64_t(total) + 1;
}

__global__ void prefilter_index_len_kernel(const int64_t* __restrict__ queries, const int64_t* __restrict__ test_map,
                                           int64_t n_test, int64_t nq, const int64_t* __restrict__ sec_off,
                                           int64_t* __restrict__ idx_len, int* __restrict__ err) {
    const int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int64_t w = wrap_id(queries ? queries[q] : q, n_test);
    int64_t name = 0;
    if (w < 0 || w >= n_test) atomicOr(err, 2);
    else name = test_map[w];
    idx_len[q] = dec_len_i64(name) + 1 + dec_len_i64(sec_off[q]) + 1 + dec_len_i64(sec_off[q + 1] - sec_off[q]) + 1;
}

__global__ void prefilter_index_emit_kernel(const int64_t* __restrict__ queries, const int64_t* __restrict__ test_map,
                                            int64_t n_test, int64_t nq, const int64_t* __restrict__ sec_off,
                                            const int64_t* __restrict__ idx_off, uint8_t* __restrict__ out) {
    const int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int64_t w = wrap_id(queries ? queries[q] : q, n_test);
    const int64_t name = (w < 0 || w >= n_test) ? 0 : test_map[w];
    uint8_t* end = out + idx_off[q + 1];
    const int64_t len = sec_off[q + 1] - sec_off[q];
    *--end = '\n';
    put_i64(end, len);
    end -= dec_len_i64(len);
    *--end = '\t';
    put_i64(end, sec_off[q]);
    end -= dec_len_i64(sec_off[q]);
    *--end = '\t';
    put_i64(end, name);
}

// One CTA per query.  Thread t owns the contiguous hits [t * C, (t + 1) * C): a block scan of the per-thread byte
// counts gives every thread its position inside the section.  Sections up to kStageBytes are assembled in shared
// memory (at the same 16-byte phase as their place in the file) and copied out with 16-byte stores.
constexpr int kEmitThreads = 256;
constexpr int kStageBytes = 40 * 1024;

__global__ void __launch_bounds__(kEmitThreads)
prefilter_emit_kernel(const int64_t* __restrict__ I, const float* __restrict__ D, int k,
                      const int64_t* __restrict__ train_map, int64_t n_train, int clip,
                      const int64_t* __restrict__ sec_off, uint8_t* __restrict__ out, int* __restrict__ err) {
    __shared__ __align__(16) uint8_t stage[kStageBytes + 16];
    __shared__ int warp_tot[kEmitThreads / 32];
    const int64_t q = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int C = (k + kEmitThreads - 1) / kEmitThreads;
    const int h0 = t * C, h1 = min(k, h0 + C);
    int mine = 0;
    for (int h = h0; h < h1; ++h) mine += make_line(I[q * k + h], D[q * k + h], train_map, n_train, clip, err).len;
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int pos = incl - mine;
    for (int w = 0; w < warp; ++w) pos += warp_tot[w];
    const int64_t g0 = sec_off[q];
    const int sec = int(sec_off[q + 1] - g0);  // includes the NUL
    const int phase = int(g0 & 15);
    const bool staged = sec <= kStageBytes;
    uint8_t* base = staged ? stage + phase : out + g0;
    for (int h = h0; h < h1; ++h) {
        const HitLine L = make_line(I[q * k + h], D[q * k + h], train_map, n_train, clip, err);
        if (!L.len) continue;
        uint8_t* end = base + pos + L.len;
        *--end = '\n';
        *--end = '0';
        *--end = '\t';
        put_big(end, L.score, 0);
        end -= dec_len(L.score);
        *--end = '\t';
        put_i64(end, L.id);
        pos += L.len;
    }
    if (t == kEmitThreads - 1) base[sec - 1] = 0;  // "every section is null delimited"
    if (!staged) return;
    __syncthreads();
    // copy stage[phase, phase + sec) -> out[g0, g0 + sec): bytes up to the first 16-byte boundary, vectors, tail bytes
    uint8_t* dst = out + (g0 - phase);  // 16-byte aligned (cudaMalloc base is)
    const int first = phase, last = phase + sec;
    const int v0 = (first + 15) & ~15, v1 = last & ~15;
    if (v0 <= v1) {
        for (int i = first + t; i < v0; i += kEmitThreads) dst[i] = stage[i];
        for (int i = v0 / 16 + t; i < v1 / 16; i += kEmitThreads)
            reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(stage)[i];
        for (int i = v1 + t; i < last; i += kEmitThreads) dst[i] = stage[i];
    } else {
        for (int i = first + t; i < last; i += kEmitThreads) dst[i] = stage[i];
    }
}

// ---- exclusive scan of int64 (section lengths -> file offsets) -------------------------------------------------
constexpr int kScanBlock = 1024;

__global__ void __launch_bounds__(kScanBlock)
scan_block_kernel(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ out, int64_t* __restrict__ block_sum) {
    __shared__ int64_t warp_tot[32];
    const int64_t i = int64_t(blockIdx.x) * kScanBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t v = i < n ? in[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int64_t o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int64_t w = warp_tot[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t o = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += o;
        }
        warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int64_t before = warp ? warp_tot[warp - 1] : 0;
    if (i < n) out[i] = before + incl - v;
    if (block_sum && threadIdx.x == kScanBlock - 1) block_sum[blockIdx.x] = before + incl;
}

__global__ void scan_add_kernel(int64_t* __restrict__ out, int64_t n, const int64_t* __restrict__ block_off) {
    const int64_t i = int64_t(blockIdx.x) * kScanBlock + threadIdx.x;
    if (i < n) out[i] += block_off[blockIdx.x];
}

__global__ void scan_total_kernel(const int64_t* __restrict__ in, const int64_t* __restrict__ excl, int64_t n,
                                  int64_t* __restrict__ total) {
    *total = n ? excl[n - 1] + in[n - 1] : 0;
}

// out[0..n) = exclusive scan of in[0..n), out[n] = total.  `in` may not alias `out`.
int exclusive_scan_i64(const int64_t* in, int64_t n, int64_t* out, cudaStream_t s) {
    if (n <= 0) {
        KNN_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), s));
        return KNN_OK;
    }
    const int64_t nblk = (n + kScanBlock - 1) / kScanBlock;
    int64_t* sums = nullptr;
    if (nblk > 1) KNN_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&sums), sizeof(int64_t) * size_t(2 * nblk + 1), s));
    scan_block_kernel<<<unsigned(nblk), kScanBlock, 0, s>>>(in, n, out, sums);
    KNN_CHECK_LAUNCH();
    if (nblk > 1) {
        int rc = exclusive_scan_i64(sums, nblk, sums + nblk, s);
        if (rc != KNN_OK) return rc;
        scan_add_kernel<<<unsigned(nblk), kScanBlock, 0, s>>>(out, n, sums + nblk);
        KNN_CHECK_LAUNCH();
        KNN_CHECK_CUDA(cudaFreeAsync(sums, s));
    }
    scan_total_kernel<<<1, 1, 0, s>>>(in, out, n, out + n);
    KNN_CHECK_LAUNCH();
    return KNN_OK;
}

int check_hits(const char* what, int64_t nq, int64_t k, const void* I) {
    if (nq < 0 || k <= 0 || k > (int64_t(1) << 20) || (nq > 0 && !I)) {
        set_error("%s: invalid arguments (nq=%lld, k=%lld)", what, (long long)nq, (long long)k);
        return KNN_ERR_INVALID;
    }
    return KNN_OK;
}

}  // namespace
}  // namespace knn

using namespace knn;

extern "C" {

int knn_eval_family_dev(int64_t nq, int64_t k, const int64_t* I_dev, const int32_t* query_family_dev,
                        const int32_t* db_family_dev, int64_t n_db, int32_t* lead_dev, int32_t* tp_dev, int* err_dev,
                        void* stream) {
    KNN_CHECK(check_hits("eval_family", nq, k, I_dev));
    PtrDeviceGuard dev_guard(I_dev ? static_cast<const void*>(I_dev) : static_cast<const void*>(nullptr));
    if (nq == 0) return KNN_OK;
    if (!query_family_dev || !db_family_dev || !lead_dev || !tp_dev || !err_dev || n_db <= 0) {
        set_error("eval_family: null argument or empty database");
        return KNN_ERR_INVALID;
    }
    eval_family_kernel<<<unsigned((nq + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        I_dev, nq, int(k), query_family_dev, db_family_dev, n_db, lead_dev, tp_dev, err_dev);
    KNN_CHECK_LAUNCH();
    return KNN_OK;
}

int knn_eval_levels_dev(int64_t nq, int64_t k, const int64_t* I_dev, const int32_t* mapping_dev, int levels, int64_t n_db,
                        uint8_t* out_dev, int* err_dev, void* stream) {
    KNN_CHECK(check_hits("eval_levels", nq, k, I_dev));
    PtrDeviceGuard dev_guard(I_dev ? static_cast<const void*>(I_dev) : static_cast<const void*>(nullptr));
    if (nq == 0) return KNN_OK;
    if (!mapping_dev || !out_dev || !err_dev || levels <= 0 || n_db < nq) {
        set_error("eval_levels: null argument, levels <= 0 or fewer label rows than queries");
        return KNN_ERR_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int64_t q0 = 0; q0 < nq; q0 += 65535) {
        const int64_t qn = nq - q0 < 65535 ? nq - q0 : 65535;
        dim3 grid(unsigned((k + 255) / 256), unsigned(qn));
        eval_levels_kernel<<<grid, 256, 0, s>>>(I_dev, q0, int(k), mapping_dev, levels, n_db, out_dev, err_dev);
        KNN_CHECK_LAUNCH();
    }
    return KNN_OK;
}

int knn_eval_sets_dev(int64_t nq, int64_t k, const int64_t* I_dev, const int64_t* set_offsets_dev,
                      const int64_t* set_members_dev, int64_t n_db_wrap, uint8_t* correct_dev, int32_t* lead_dev,
                      void* stream) {
    KNN_CHECK(check_hits("eval_sets", nq, k, I_dev));
    PtrDeviceGuard dev_guard(I_dev ? static_cast<const void*>(I_dev) : static_cast<const void*>(nullptr));
    if (nq == 0) return KNN_OK;
    if (!set_offsets_dev || (!correct_dev && !lead_dev)) {
        set_error("eval_sets: null argument");
        return KNN_ERR_INVALID;
    }
    eval_sets_kernel<<<unsigned((nq + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        I_dev, nq, int(k), set_offsets_dev, set_members_dev, n_db_wrap, correct_dev, lead_dev);
    KNN_CHECK_LAUNCH();
    return KNN_OK;
}

int knn_remove_self_hit_dev(int64_t nq, int64_t k, int64_t* I_dev, float* D_dev, const int64_t* self_ids_dev,
                            uint64_t* n_missing_dev, void* stream) {
    KNN_CHECK(check_hits("remove_self_hit", nq, k, I_dev));
    PtrDeviceGuard dev_guard(I_dev ? static_cast<const void*>(I_dev) : static_cast<const void*>(nullptr));
    if (nq == 0) return KNN_OK;
    if (!n_missing_dev) {
        set_error("remove_self_hit: null counter");
        return KNN_ERR_INVALID;
    }
    remove_self_hit_kernel<<<unsigned((nq + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        I_dev, D_dev, nq, int(k), self_ids_dev, reinterpret_cast<unsigned long long*>(n_missing_dev));
    KNN_CHECK_LAUNCH();
    return KNN_OK;
}

int knn_prefilter_measure_dev(int64_t nq, int64_t k, const int64_t* I_dev, const float* D_dev, const int64_t* queries_dev,
                              const int64_t* test_map_dev, int64_t n_test, const int64_t* train_map_dev, int64_t n_train,
                              int clip, int64_t* sec_off_dev, int64_t* idx_off_dev, int* err_dev, void* stream) {
    KNN_CHECK(check_hits("prefilter_measure", nq, k, I_dev));
    PtrDeviceGuard dev_guard(I_dev ? static_cast<const void*>(I_dev) : static_cast<const void*>(nullptr));
    if (!sec_off_dev || !idx_off_dev || !err_dev || (nq > 0 && (!D_dev || !test_map_dev || !train_map_dev)) || n_test <= 0 ||
        n_train <= 0) {
        set_error("prefilter_measure: null argument or empty id map");
        return KNN_ERR_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int64_t* len = nullptr;
    KNN_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&len), sizeof(int64_t) * size_t(nq + 1), s));
    if (nq > 0) {
        prefilter_measure_kernel<<<unsigned((nq + 7) / 8), 256, 0, s>>>(I_dev, D_dev, nq, int(k), train_map_dev, n_train,
                                                                        clip, len, err_dev);
        KNN_CHECK_LAUNCH();
    }
    KNN_CHECK(exclusive_scan_i64(len, nq, sec_off_dev, s));
    if (nq > 0) {
        prefilter_index_len_kernel<<<unsigned((nq + 255) / 256), 256, 0, s>>>(queries_dev, test_map_dev, n_test, nq,
                                                                              sec_off_dev, len, err_dev);
        KNN_CHECK_LAUNCH();
    }
    KNN_CHECK(exclusive_scan_i64(len, nq, idx_off_dev, s));
    KNN_CHECK_CUDA(cudaFreeAsync(len, s));
    return KNN_OK;
}

int knn_prefilter_emit_dev(int64_t nq, int64_t k, const int64_t* I_dev, const float* D_dev, const int64_t* queries_dev,
                           const int64_t* test_map_dev, int64_t n_test, const int64_t* train_map_dev, int64_t n_train,
                           int clip, const int64_t* sec_off_dev, const int64_t* idx_off_dev, uint8_t* data_dev,
                           uint8_t* index_dev, int* err_dev, void* stream) {
    KNN_CHECK(check_hits("prefilter_emit", nq, k, I_dev));
    PtrDeviceGuard dev_guard(I_dev ? static_cast<const void*>(I_dev) : static_cast<const void*>(nullptr));
    if (nq == 0) return KNN_OK;
    if (!D_dev || !test_map_dev || !train_map_dev || !sec_off_dev || !idx_off_dev || !data_dev || !index_dev || !err_dev) {
        set_error("prefilter_emit: null argument");
        return KNN_ERR_INVALID;
    }
    if (reinterpret_cast<uintptr_t>(data_dev) % 16 != 0) {
        set_error("prefilter_emit: data_dev must be 16-byte aligned");
        return KNN_ERR_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    prefilter_emit_kernel<<<unsigned(nq), kEmitThreads, 0, s>>>(I_dev, D_dev, int(k), train_map_dev, n_train, clip,
                                                               sec_off_dev, data_dev, err_dev);
    KNN_CHECK_LAUNCH();
    prefilter_index_emit_kernel<<<unsigned((nq + 255) / 256), 256, 0, s>>>(queries_dev, test_map_dev, n_test, nq,
                                                                           sec_off_dev, idx_off_dev, index_dev);
    KNN_CHECK_LAUNCH();
    return KNN_OK;
}

}  // extern "C"
