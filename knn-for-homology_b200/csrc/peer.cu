// Cross-shard exchange fused with the k-way merge, over NVLink peer memory.
//
// New relative to the reference (single process, no sharding).  Every rank owns one GPU and a cudaMalloc'ed
// exchange buffer that the other ranks of the box map with CUDA IPC.  After the per-shard search each rank merges
// ITS slice of the queries: the kernel reads the G per-shard (D, I) rows of a query straight out of the peers'
// memory (coalesced NVLink loads into shared memory), ranks every candidate against the other lists (the lists are
// sorted, so the final position of an element is its own rank plus one binary search per other list - no serial
// tournament), and stores the merged row into EVERY rank's output buffer.  One kernel replaces
// all-gather (G x the payload per rank) + a full merge on every rank: each result row crosses NVLink once in and
// once out, and every rank merges nq / G queries.
#include <algorithm>

#include "common.cuh"

namespace knn {
namespace {

constexpr int kMaxPeers = 16;

struct PeerTable {
    const float* D[kMaxPeers];      // per-shard scores  [nq][k], rank l's buffer
    const int64_t* I[kMaxPeers];    // per-shard global ids
    float* D_out[kMaxPeers];        // merged result, one copy per rank
    int64_t* I_out[kMaxPeers];
};

__global__ void __launch_bounds__(256)
merge_peer_kernel(PeerTable t, int nranks, int64_t q0, int k, int largest) {
    extern __shared__ uint64_t sk[];  // [nranks * k] candidate keys, [k] merged keys
    const int64_t q = q0 + blockIdx.x;
    const int tid = threadIdx.x, T = blockDim.x;
    const int total = nranks * k;
    uint64_t* merged = sk + total;
    for (int i = tid; i < total; i += T) {
        const int l = i / k, r = i - l * k;
        const int64_t off = q * k + r;
        const int64_t id = t.I[l][off];
        sk[i] = make_key(t.D[l][off], id < 0 ? kInvalidId : uint32_t(id), largest);
    }
    __syncthreads();
    for (int i = tid; i < total; i += T) {
        const int l = i / k;
        const uint64_t key = sk[i];
        int pos = i - l * k;  // elements of the own (sorted) list in front of it
        for (int l2 = 0; l2 < nranks && pos < k; ++l2) {
            if (l2 == l) continue;
            // number of elements of list l2 that come first: key2 > key, or key2 == key (padding) and l2 < l
            const uint64_t* a = sk + l2 * k;
            int lo = 0, hi = k;
            if (l2 < l) {
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (a[mid] >= key) lo = mid + 1;
                    else hi = mid;
                }
            } else {
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (a[mid] > key) lo = mid + 1;
                    else hi = mid;
                }
            }
            pos += lo;
        }
        if (pos < k) merged[pos] = key;
    }
    __syncthreads();
    for (int dst = 0; dst < nranks; ++dst) {
        float* D = t.D_out[dst] + q * k;
        int64_t* I = t.I_out[dst] + q * k;
        for (int r = tid; r < k; r += T) {
            const uint64_t key = merged[r];
            D[r] = key ? key_score(key, largest) : (largest ? -FLT_MAX : FLT_MAX);
            I[r] = key ? int64_t(key_id(key)) : int64_t(-1);
        }
    }
}

// ---- bound exchange of the two-phase search, fused into two tiny kernels over peer memory ----
// Per query batch every shard contributes 2 x batch_rows floats (lower bound on its k-th best score, minus the bound
// on its j-th best); the combined bound is the element-wise MAX over the shards.  An NCCL all-reduce per batch would
// do, but its kernel wants ~100 registers x 640 threads: it does not fit next to the resident GEMM CTA and would take
// an SM from the next panel launch.  Instead: `push` stores this rank's slice into slot [rank] of EVERY rank's buffer
// (NVLink stores), fences, and raises flag [rank] there to the call's epoch; `wait_max` (32 threads per CTA, a few
// registers: co-resident with anything) spins on the local flags of all ranks and then reduces the local slots.
// No rank ever waits inside a kernel for something that depends on its own later work, so there is no cycle; a
// bounded spin (~4 s) turns a lost peer into an error flag instead of a hang.
struct BoundsPeers {
    float* slots[kMaxPeers];       // rank l's slot array [nranks][n] (for this batch)
    unsigned* flags[kMaxPeers];    // rank l's flags [nranks]          (for this batch)
};

__global__ void __launch_bounds__(256)
bounds_push_kernel(BoundsPeers t, int nranks, int rank, int64_t n, const float* __restrict__ src, unsigned epoch,
                   unsigned* __restrict__ done_counter) {
    const int64_t total = int64_t(nranks) * n;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int dst = int(i / n);
        const int64_t e = i - int64_t(dst) * n;
        t.slots[dst][int64_t(rank) * n + e] = src[e];
    }
    __threadfence_system();  // this thread's stores are visible system-wide before the counter / flags move
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned arrived = atomicAdd(done_counter, 1u) + 1u;
        if (arrived == gridDim.x) {  // last CTA: every store of the grid is fenced
            *done_counter = 0;
            __threadfence_system();
            for (int dst = 0; dst < nranks; ++dst)
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(t.flags[dst] + rank), "r"(epoch) : "memory");
        }
    }
}

__global__ void __launch_bounds__(256)
bounds_wait_max_kernel(const float* __restrict__ slots, const unsigned* __restrict__ flags, int nranks, int64_t n, unsigned epoch,
                       float* __restrict__ out, int* __restrict__ timeout_flag) {
    if (threadIdx.x < nranks) {
        const unsigned* f = flags + threadIdx.x;
        const long long t0 = clock64();
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (int(v - epoch) >= 0) break;
            if (clock64() - t0 > 8000000000ll) {  // ~4 s at 2 GHz: a peer never arrived
                atomicExch(timeout_flag, 1);
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        float m = -FLT_MAX;
        for (int r = 0; r < nranks; ++r) {
            const float v = __ldcg(slots + int64_t(r) * n + i);  // written by a peer over NVLink: read at L2
            m = v > m ? v : m;
        }
        out[i] = m;
    }
}

}  // namespace
}  // namespace knn

using namespace knn;

extern "C" {

int knn_peer_buffer_alloc(void** out, int64_t bytes, int device) {
    if (!out || bytes <= 0) {
        set_error("peer_buffer_alloc: invalid arguments");
        return KNN_ERR_INVALID;
    }
    int prev = -1;
    KNN_CHECK_CUDA(cudaGetDevice(&prev));
    KNN_CHECK_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaMalloc(out, size_t(bytes));  // plain cudaMalloc: exportable with cudaIpcGetMemHandle
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("peer_buffer_alloc: cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
        return KNN_ERR_MEMORY;
    }
    return KNN_OK;
}

int knn_peer_buffer_free(void* p) {
    if (p) KNN_CHECK_CUDA(cudaFree(p));
    return KNN_OK;
}

int knn_peer_handle_get(const void* dev_ptr, unsigned char* handle64) {
    if (!dev_ptr || !handle64) {
        set_error("peer_handle_get: null argument");
        return KNN_ERR_INVALID;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    KNN_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle64, &h, 64);
    return KNN_OK;
}

int knn_peer_handle_open(const unsigned char* handle64, int device, void** out) {
    if (!handle64 || !out) {
        set_error("peer_handle_open: null argument");
        return KNN_ERR_INVALID;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    int prev = -1;
    KNN_CHECK_CUDA(cudaGetDevice(&prev));
    KNN_CHECK_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("peer_handle_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
        return KNN_ERR_CUDA;
    }
    return KNN_OK;
}

int knn_peer_handle_close(void* p) {
    if (p) KNN_CHECK_CUDA(cudaIpcCloseMemHandle(p));
    return KNN_OK;
}

int knn_bounds_push_peer_dev(int nranks, int rank, int64_t n, const float* src_dev, void* const* slots_peer,
                             void* const* flags_peer, uint32_t epoch, uint32_t* counter_dev, void* stream) {
    if (nranks <= 0 || nranks > kMaxPeers || rank < 0 || rank >= nranks || n <= 0 || !src_dev || !slots_peer || !flags_peer || !counter_dev) {
        set_error("bounds_push: invalid arguments (at most %d ranks)", kMaxPeers);
        return KNN_ERR_INVALID;
    }
    BoundsPeers t;
    memset(&t, 0, sizeof(t));
    for (int l = 0; l < nranks; ++l) {
        if (!slots_peer[l] || !flags_peer[l]) {
            set_error("bounds_push: null peer pointer for rank %d", l);
            return KNN_ERR_INVALID;
        }
        t.slots[l] = static_cast<float*>(slots_peer[l]);
        t.flags[l] = static_cast<unsigned*>(flags_peer[l]);
    }
    PtrDeviceGuard g(src_dev);
    const int64_t total = int64_t(nranks) * n;
    const int grid = int(std::min<int64_t>(32, (total + 255) / 256));
    bounds_push_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(t, nranks, rank, n, src_dev, epoch, counter_dev);
    KNN_CHECK_LAUNCH();
    return KNN_OK;
}

int knn_bounds_wait_max_dev(int nranks, int64_t n, const float* slots_dev, const uint32_t* flags_dev, uint32_t epoch,
                            float* out_dev, int* timeout_flag_dev, void* stream) {
    if (nranks <= 0 || nranks > kMaxPeers || n <= 0 || !slots_dev || !flags_dev || !out_dev || !timeout_flag_dev) {
        set_error("bounds_wait_max: invalid arguments");
        return KNN_ERR_INVALID;
    }
    PtrDeviceGuard g(out_dev);
    const int grid = int(std::min<int64_t>(16, (n + 255) / 256));
    bounds_wait_max_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(slots_dev, flags_dev, nranks, n, epoch, out_dev,
                                                                                timeout_flag_dev);
    KNN_CHECK_LAUNCH();
    return KNN_OK;
}

int knn_merge_topk_peer_dev(int metric, int64_t nq, int64_t k, int nranks, int64_t q0, int64_t q1,
                            const void* const* D_peer, const void* const* I_peer, void* const* D_out_peer,
                            void* const* I_out_peer, void* stream) {
    if (nq < 0 || k <= 0 || nranks <= 0 || q0 < 0 || q1 < q0 || q1 > nq || !D_peer || !I_peer || !D_out_peer || !I_out_peer) {
        set_error("merge_peer: invalid arguments");
        return KNN_ERR_INVALID;
    }
    const size_t smem = size_t(nranks + 1) * size_t(k) * sizeof(uint64_t);
    if (nranks > kMaxPeers || k > KNN_MAX_K || smem > 200 * 1024) {
        set_error("merge_peer: at most %d ranks and (ranks + 1) * k * 8 bytes <= 200 KiB of shared memory (got %d ranks, k=%lld)",
                  kMaxPeers, nranks, (long long)k);
        return KNN_ERR_LIMIT;
    }
    if (q1 == q0) return KNN_OK;
    PeerTable t;
    memset(&t, 0, sizeof(t));
    for (int l = 0; l < nranks; ++l) {
        if (!D_peer[l] || !I_peer[l] || !D_out_peer[l] || !I_out_peer[l]) {
            set_error("merge_peer: null peer pointer for rank %d", l);
            return KNN_ERR_INVALID;
        }
        t.D[l] = static_cast<const float*>(D_peer[l]);
        t.I[l] = static_cast<const int64_t*>(I_peer[l]);
        t.D_out[l] = static_cast<float*>(D_out_peer[l]);
        t.I_out[l] = static_cast<int64_t*>(I_out_peer[l]);
    }
    // the calling rank's device owns its own output buffer; the pointer tables do not say which entry that is, so the
    // stream decides (NULL stream: the caller's current device)
    DeviceGuard dev_guard;
    if (stream) {
        int sdev = -1;
        if (cudaStreamGetDevice(static_cast<cudaStream_t>(stream), &sdev) == cudaSuccess) dev_guard.select(sdev);
        else cudaGetLastError();
    }
    KNN_CHECK_CUDA(cudaFuncSetAttribute(merge_peer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int64_t a = q0; a < q1; a += 65535 * 16) {  // grid.x limit is 2^31-1; chunking keeps launches bounded anyway
        const int64_t n = q1 - a < 65535 * 16 ? q1 - a : 65535 * 16;
        merge_peer_kernel<<<unsigned(n), 256, smem, s>>>(t, nranks, a, int(k), metric == KNN_METRIC_INNER_PRODUCT);
        KNN_CHECK_LAUNCH();
    }
    return KNN_OK;
}

}  // extern "C"
