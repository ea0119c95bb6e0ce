// Microbenchmark (diagnostics): MUFU.EX2 / F2FP / FFMA throughput per SM vs resident warps.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(float* out, unsigned long long* cyc, int iters) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
    unsigned acc = 0;
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) { asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); x[i] -= 1.0f; }            // MUFU + FADD
            if (MODE == 1) { x[i] = fmaf(x[i], 0.999f, -0.5f); }                                          // FFMA only
            if (MODE == 2) { __half2 h = __floats2half2_rn(x[i], x[(i + 1) & 15]); acc ^= *reinterpret_cast<unsigned*>(&h); x[i] += 1.0f; }  // F2FP + FADD
            if (MODE == 3) { float t = fmaf(x[i], 0.18f, -3.0f); asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(t)); x[i] = x[i] * 0.5f + t; }  // FFMA+MUFU+FFMA
        }
    }
    unsigned long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* o; unsigned long long* c; cudaMalloc(&o, 1 << 20); cudaMalloc(&c, 4096);
    const int iters = 2000;
    const char* names[4] = {"ex2+fadd", "ffma", "f2fp.pack+fadd", "ffma+ex2+ffma"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {4, 8, 16, 32}) {
            if (mode == 0) k<0><<<1, warps * 32>>>(o, c, iters); if (mode == 1) k<1><<<1, warps * 32>>>(o, c, iters);
            if (mode == 2) k<2><<<1, warps * 32>>>(o, c, iters); if (mode == 3) k<3><<<1, warps * 32>>>(o, c, iters);
            cudaDeviceSynchronize();
            unsigned long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
            double ops = (double)iters * 16 * warps * 32;
            printf("%-16s warps/SM %2d: %.2f elem/clk/SM  (%.1f cycles per warp-iteration-element)\n", names[mode], warps, ops / cy, (double)cy / iters / 16);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
