// Does a small kernel on stream B run while a persistent, shared-memory-heavy (cluster) kernel occupies every SM
// on stream A?  Prints when the small kernel finished relative to the big one for a few configurations.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void big(long long cycles, int* sink) {
    extern __shared__ unsigned char sm[];
    long long t0 = clock64();
    while (clock64() - t0 < cycles) {
        if (sm[threadIdx.x] == 77) atomicAdd(sink, 1);
    }
}
__global__ void small_k(int* out, int smem_touch) {
    extern __shared__ unsigned char sm2[];
    if (smem_touch) sm2[threadIdx.x] = 1;
    long long t0 = clock64();
    while (clock64() - t0 < 20000) {}
    if (threadIdx.x == 0) atomicAdd(out, 1);
}

int main() {
    int* d;
    cudaMalloc(&d, 8);
    cudaMemset(d, 0, 8);
    cudaStream_t a, b;
    cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    cudaEvent_t ea0, ea1, eb0, eb1;
    cudaEventCreate(&ea0); cudaEventCreate(&ea1); cudaEventCreate(&eb0); cudaEventCreate(&eb1);
    for (int cluster = 1; cluster <= 2; ++cluster) {
        for (int big_smem : {165136, 100000, 220000}) {
            for (int small_smem : {0, 4096, 34320})
            for (int carve : {-1, 100}) {  // preferred shared-memory carveout of BOTH kernels: driver default / max shared
                cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, 230000);
                cudaFuncSetAttribute(small_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
                cudaFuncSetAttribute(small_k, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
                cudaFuncSetAttribute(big, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(148); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = big_smem; cfg.stream = a;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaDeviceSynchronize();
                cudaEventRecord(ea0, a);
                cudaLaunchKernelEx(&cfg, big, (long long)4000000, d);  // ~2+ ms
                cudaEventRecord(ea1, a);
                cudaEventRecord(eb0, b);
                small_k<<<2048, 256, small_smem, b>>>(d + 1, small_smem > 0);
                cudaEventRecord(eb1, b);
                cudaDeviceSynchronize();
                float big_ms, small_end_ms, small_ms;
                cudaEventElapsedTime(&big_ms, ea0, ea1);
                cudaEventElapsedTime(&small_end_ms, ea0, eb1);
                cudaEventElapsedTime(&small_ms, eb0, eb1);
                printf("cluster=%d big_smem=%d small_smem=%d carveout=%d: big %.3f ms, small finished at %.3f ms (took %.3f) -> %s  [%s]\n", cluster,
                       big_smem, small_smem, carve, big_ms, small_end_ms, small_ms, small_end_ms < big_ms * 0.9 ? "OVERLAPPED" : "serialized",
                       cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
