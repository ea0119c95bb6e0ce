// Microbenchmark (diagnostics): how many SMs can a grid of thread-block clusters occupy at once on this part?
// Every CTA takes a whole SM (200 KB of dynamic shared memory), records its SM id and start time, then spins 100 us.
// CTAs of the first wave start together; the count of distinct SMs among them is the usable SM count for that cluster size.
// (B300 notes: cluster size 4 strands 16 of 148 SMs because of GPCs with 18 SMs; this checks the B200 in the pool.)
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
#include <set>
__global__ void k(unsigned* smid, unsigned long long* t_start) {
    extern __shared__ char smem[];
    unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (threadIdx.x == 0) { smid[blockIdx.x] = s; t_start[blockIdx.x] = t; smem[0] = 1; }
    while (true) { unsigned long long n; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n)); if (n - t > 100000) break; }
}
int main() {
    unsigned* d_s; unsigned long long* d_t; const int G = 296;
    cudaMalloc(&d_s, G * 4); cudaMalloc(&d_t, G * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(G / cs * cs); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {unsigned(cs), 1, 1};
        cfg.attrs = at; cfg.numAttrs = 1;
        int maxc = -1; cudaOccupancyMaxActiveClusters(&maxc, k, &cfg);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k, d_s, d_t); cudaDeviceSynchronize();
        std::vector<unsigned> s(G); std::vector<unsigned long long> t(G);
        cudaMemcpy(s.data(), d_s, G * 4, cudaMemcpyDeviceToHost); cudaMemcpy(t.data(), d_t, G * 8, cudaMemcpyDeviceToHost);
        const int n = G / cs * cs; unsigned long long t0 = *std::min_element(t.begin(), t.begin() + n);
        std::set<unsigned> first; int nfirst = 0;
        for (int i = 0; i < n; ++i) if (t[i] - t0 < 50000) { first.insert(s[i]); ++nfirst; }
        printf("cluster %2d: grid %d, first wave %d CTAs on %zu distinct SMs; cudaOccupancyMaxActiveClusters %d (x%d = %d CTAs); %s\n", cs, n, nfirst,
               first.size(), maxc, cs, maxc * cs, cudaGetErrorString(e == cudaSuccess ? cudaGetLastError() : e));
    }
}
