// Persistent warp-specialised GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias)
//
//   A: fp16 row-major [M,K] (activations), W: fp16 row-major [N,K] (HF Linear weight, y = x W^T + b)
//   -> both operands are K-major, the canonical UMMA "TN" case.
//   TMA (128B swizzle) -> 4-stage smem ring -> tcgen05.mma 128x256x16 (one issuing thread) ->
//   fp32 accumulators in TMEM, double buffered (2 x 256 columns) -> 8 epilogue warps read TMEM
//   (tcgen05.ld 32x32b; 8 warps, two per SMSP), fuse bias / erf-GELU / fp32 residual add, store to global.
//
// Replaces (inside TEI, un-vendored; restated from HF modeling_bert.py): the Linear layers of
// BertSelfAttention :143-207 (fused QKV), BertSelfOutput.dense :287-298, BertIntermediate :330-342
// (dense + GELU) and BertOutput.dense :345-356.
#include "kernels.h"
#include "ptx.cuh"

namespace b200 {
namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES = BN * BK * 2;          // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 48 KB
constexpr int NUM_EPI_WARPS = 8;  // two per SMSP: warps 4..7 take columns [0,128), warps 8..11 take [128,256)
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;

// GELU(x) = x * Phi(x), Phi via erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7 on erf):
//   a = |x|/sqrt2, t = 1/(1 + p a), q = 0.5 * poly(t) * exp(-a^2);  Phi = x >= 0 ? 1 - q : q.
// 2 MUFU (rcp, ex2) + ~12 FMA-pipe ops per element, so the FFN1 epilogue (128x256 elements per tile)
// stays inside the issue budget of the tile's MMA time.
__device__ __forceinline__ float gelu_erf(float x) {
    const float ax = fabsf(x);
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752f, ax, 1.0f)));
    float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    p = fmaf(p, t, 0.5f * 1.421413741f);
    p = fmaf(p, t, 0.5f * -0.284496736f);
    p = fmaf(p, t, 0.5f * 0.254829592f);
    const float e = ex2_approx((ax * (-0.5f * 1.4426950408889634f)) * ax);
    const float q = (p * t) * e;
    const float phi = x >= 0.f ? 1.0f - q : q;
    return x * phi;
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const float* __restrict__ bias, const float* __restrict__ resid, void* __restrict__ out, int M, int N,
            int K) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;            // [STAGES] TMA -> MMA
    uint64_t* empty = bars + STAGES;  // [STAGES] MMA -> TMA
    uint64_t* tfull = bars + 2 * STAGES;      // [2] MMA -> epilogue
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
    const int lane = lane_id();
    const int num_m = (M + BM - 1) / BM;
    const int num_n = N / BN;
    const int num_tiles = num_m * num_n;
    const int kblocks = K / BK;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tma_a);
        prefetch_tmap(&tma_b);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], NUM_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // ---------------------------------------------------------------- TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / num_n, n_blk = tile % num_n;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(sa, &tma_a, &full[stage], kb * BK, m_blk * BM);
                    tma_load_2d(sa + A_BYTES, &tma_b, &full[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc = make_idesc_f16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        umma_f16_ss(d_tmem, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc,
                                    (kb | k) != 0);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[as]);
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // -------------------------------------------------------------------- epilogue warps
        const int ew = warp & 3;          // this warp may touch TMEM lanes [32*ew, 32*ew+32)
        const int half = (warp - 4) >> 2;  // which 128 columns of the tile
        constexpr int CHUNKS = BN / 2 / 32;  // 4 chunks of 32 columns per warp
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            const int row = m_blk * BM + ew * 32 + lane;
            const bool row_ok = row < M;
            const int col0 = n_blk * BN + half * (BN / 2);
            const size_t row_off = static_cast<size_t>(row) * N + col0;
            float4 rq[8];  // residual of the chunk about to be processed (prefetched: it does not depend on the MMA)
            if constexpr (EPI == EPI_BIAS_RES_F32) {
                if (row_ok) {
                    const float4* r4 = reinterpret_cast<const float4*>(resid + row_off);
#pragma unroll
                    for (int j = 0; j < 8; ++j) rq[j] = r4[j];
                }
            }
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < CHUNKS; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN + half * (BN / 2) + c * 32, r);
                float4 rn[8];
                if constexpr (EPI == EPI_BIAS_RES_F32) {
                    if (row_ok && c + 1 < CHUNKS) {
                        const float4* r4 = reinterpret_cast<const float4*>(resid + row_off + (c + 1) * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) rn[j] = r4[j];
                    }
                }
                const float4* b4 = reinterpret_cast<const float4*>(bias + col0 + c * 32);
                float4 bq[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) bq[j] = __ldg(b4 + j);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + bq[j].x;
                    v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bq[j].y;
                    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bq[j].z;
                    v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bq[j].w;
                }
                if constexpr (EPI == EPI_BIAS_RES_F32) {
                    if (row_ok) {
                        float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(out) + row_off + c * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            o4[j] = make_float4(v[4 * j] + rq[j].x, v[4 * j + 1] + rq[j].y, v[4 * j + 2] + rq[j].z,
                                                v[4 * j + 3] + rq[j].w);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) rq[j] = rn[j];
                } else {
                    if constexpr (EPI == EPI_BIAS_GELU_F16) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                    }
                    if (row_ok) {
                        uint4* o4 = reinterpret_cast<uint4*>(static_cast<__half*>(out) + row_off + c * 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            o4[j] = make_uint4(pack_half2(v[8 * j], v[8 * j + 1]), pack_half2(v[8 * j + 2], v[8 * j + 3]),
                                               pack_half2(v[8 * j + 4], v[8 * j + 5]),
                                               pack_half2(v[8 * j + 6], v[8 * j + 7]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace gemm

template <int EPI>
static cudaError_t launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const float* bias, const float* resid,
                              void* out, int M, int N, int K, int sm_count, cudaStream_t stream) {
    const int tiles = ((M + gemm::BM - 1) / gemm::BM) * (N / gemm::BN);
    const int grid = tiles < sm_count ? tiles : sm_count;
    gemm::gemm_kernel<EPI><<<grid, gemm::NUM_THREADS, gemm::SMEM_BYTES, stream>>>(ta, tb, bias, resid, out, M, N, K);
    return cudaGetLastError();
}

cudaError_t gemm_init_device() {
    cudaError_t e;
    e = cudaFuncSetAttribute(gemm::gemm_kernel<EPI_BIAS_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             gemm::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm::gemm_kernel<EPI_BIAS_GELU_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             gemm::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm::gemm_kernel<EPI_BIAS_RES_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                gemm::SMEM_BYTES);
}

cudaError_t launch_gemm(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const float* bias, const float* resid,
                        void* out, int M, int N, int K, int sm_count, cudaStream_t stream) {
    if (N % gemm::BN != 0 || K % gemm::BK != 0 || M <= 0) return cudaErrorInvalidValue;
    switch (epi) {
        case EPI_BIAS_F16: return launch_one<EPI_BIAS_F16>(ta, tb, bias, resid, out, M, N, K, sm_count, stream);
        case EPI_BIAS_GELU_F16: return launch_one<EPI_BIAS_GELU_F16>(ta, tb, bias, resid, out, M, N, K, sm_count, stream);
        case EPI_BIAS_RES_F32: return launch_one<EPI_BIAS_RES_F32>(ta, tb, bias, resid, out, M, N, K, sm_count, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace b200
