// Microbenchmark (diagnostics): the attention kernel's exp phase in isolation -- 64 scores per thread:
// FFMA -> MUFU.EX2 -> row sum + fp16 pack -> swizzled STS.128 -- for 1 or 2 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../modal-examples_b200/csrc/ptx.cuh"
using namespace b200;

__global__ void k(const float* in, float* out, unsigned long long* cyc, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int r = threadIdx.x & 127;
    const uint32_t swz = r & 7;
    const uint32_t row_ptr = smem_u32(smem) + (threadIdx.x >> 7) * 16384 + r * 128;
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = in[threadIdx.x * 64 + i];
    float tot = 0.f;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float neg_ms = -0.5f - 1e-3f * it;
        float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p0 = ex2_approx(fmaf(v[q * 8 + 2 * e], 0.18033688f, neg_ms));
                const float p1 = ex2_approx(fmaf(v[q * 8 + 2 * e + 1], 0.18033688f, neg_ms));
                ls0 += p0; ls1 += p1;
                pk[e] = pack_half2(p0, p1);
            }
            sts128(row_ptr + ((static_cast<uint32_t>(q) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        tot += ls0 + ls1;
        fence_proxy_async_smem();
    }
    const unsigned long long t1 = clock64();
    out[threadIdx.x] = tot;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float *in, *out; unsigned long long* c;
    cudaMalloc(&in, 256 * 64 * 4); cudaMemset(in, 0, 256 * 64 * 4); cudaMalloc(&out, 4096); cudaMalloc(&c, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    const int iters = 2000;
    for (int threads : {128, 256}) {
        k<<<1, threads, 40000>>>(in, out, c, iters); cudaDeviceSynchronize();
        unsigned long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
        printf("%d warps/SMSP: %.0f cycles per 64-score exp phase (per warp), err %s\n", threads / 128, (double)cy / iters, cudaGetErrorString(cudaGetLastError()));
    }
}
