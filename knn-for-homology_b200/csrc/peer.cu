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
