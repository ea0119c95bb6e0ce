// Microbenchmark (diagnostics, not product): tcgen05.ld throughput per SM for different warp counts / shapes.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../modal-examples_b200/csrc/ptx.cuh"
using namespace b200;

__device__ __forceinline__ void ld_x64(uint32_t taddr, uint32_t (&r)[64]) {
    tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
    tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
}

template <int MODE>
__global__ void k(unsigned long long* out, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // x32, wait each
            uint32_t r[32];
            tmem_ld_32x32b_x32(base + ((i * 32) & 255) + (warp >> 2) * 256, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j];
        } else {  // 2 x x32 in flight
            uint32_t r[64];
            ld_x64(base + ((i * 64) & 255) + (warp >> 2) * 256, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 64; ++j) acc ^= r[j];
        }
    }
    const unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678) out[1000] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 8192);
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {1, 4, 8}) {
            if (mode == 0) k<0><<<1, warps * 32>>>(d, iters); else k<1><<<1, warps * 32>>>(d, iters);
            cudaDeviceSynchronize();
            unsigned long long c = 0;
            cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * warps * 32 * (mode ? 64 : 32) * 4;
            printf("mode %d warps %d: %llu cycles, %.1f B/clk per SM, %.1f cycles per warp-x32\n", mode, warps, c, bytes / c,
                   (double)c / iters / (mode ? 2 : 1));
        }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
