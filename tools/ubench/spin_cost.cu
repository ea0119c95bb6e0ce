// Microbenchmark (diagnostics): what do warps spinning in mbar_wait() cost a productive warp on the same SM sub-partition?
// One warpgroup runs the attention exp phase (as tools/ubench/exp_phase.cu); 0..4 further warpgroups wait on an mbarrier
// that only completes when the first is done -- (a) with the try_wait spin loop of ptx.cuh, (b) parked in bar.sync.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../modal-examples_b200/csrc/ptx.cuh"
using namespace b200;

__global__ void k(const float* in, float* out, unsigned long long* cyc, int iters, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* done = reinterpret_cast<uint64_t*>(smem + 32768);
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile uint32_t*>(smem + 32768 + 64) = 0;
        mbar_init(done, 128);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x >= 128) {
        if (mode == 0) mbar_wait(done, 0);
        else if (mode == 1) named_bar_sync(1, blockDim.x);
        else {
            // mode 2: FMA-pipe background at full rate; mode 3: bursts of 128 FMAs separated by ~500-cycle sleeps
            float a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x % 128 + i];
            volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(smem + 32768 + 64);
            while (*flag == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 1.0001f, 0.5f);
                if (mode == 3) __nanosleep(300);
            }
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += a[i];
            out[threadIdx.x] = t;
        }
        return;
    }
    const int r = threadIdx.x & 127;
    const uint32_t swz = r & 7;
    const uint32_t row_ptr = smem_u32(smem) + r * 128;
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = in[threadIdx.x * 64 + i];
    float tot = 0.f;
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
    if (mode == 0) mbar_arrive(done);
    else if (mode == 1) named_bar_sync(1, blockDim.x);
    else *reinterpret_cast<volatile uint32_t*>(smem + 32768 + 64) = 1;
}
int main() {
    float *in, *out; unsigned long long* c;
    cudaMalloc(&in, 128 * 64 * 4); cudaMemset(in, 0, 128 * 64 * 4); cudaMalloc(&out, 4096); cudaMalloc(&c, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    const int iters = 2000;
    for (int mode = 0; mode < 4; ++mode)
        for (int extra = 0; extra <= 4; ++extra) {
            k<<<1, 128 * (1 + extra), 40000>>>(in, out, c, iters, mode); cudaDeviceSynchronize();
            unsigned long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
            printf("%s, %d waiting warps/SMSP: %.0f cycles per 64-score exp phase, err %s\n", mode == 0 ? "mbar_wait spin" : mode == 1 ? "bar.sync" : mode == 2 ? "FMA loop" : "FMA bursts", extra,
                   (double)cy / iters, cudaGetErrorString(cudaGetLastError()));
        }
}
