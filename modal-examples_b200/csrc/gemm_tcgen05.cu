// CTA-pair GEMM for sm_100a (tcgen05 cta_group::2):  C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias)
//
// A cluster of two CTAs (one TPC) owns a 256 x 256 output tile.  CTA r of the pair TMA-loads its own
// 128 rows of A and its own 128 rows (= output columns) of W per 64-wide k-block, so each SM pulls
// 32 KB per k-block from L2 instead of the 48 KB a lone 128x256 CTA needs (the first, single-CTA version of
// this kernel ran into the ~10 TB/s L2->smem fill limit at ~900 TFLOP/s; see profiles/README.md).
// The leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16); each
// CTA's TMEM receives its 128 rows x 256 fp32 columns, double buffered (2 x 256 columns), and each
// CTA's 8 epilogue warps drain them (bias / erf-GELU / fp32 residual add) while the next tile's MMAs run.
//
//   full[s]   (leader's)  : leader producer arrive.expect_tx(64 KB); both CTAs' TMA loads complete_tx on it
//   empty[s]  (per CTA)   : tcgen05.commit multicast to both CTAs once the MMAs that read stage s retire
//   tfull[a]  (per CTA)   : commit multicast after a tile's last MMA
//   tempty[a] (leader's)  : 2 x 8 epilogue warps arrive (the peer's through mapa / shared::cluster)
//
// Replaces (inside TEI, un-vendored; restated from HF modeling_bert.py): the Linear layers of
// BertSelfAttention :143-207 (fused QKV), BertSelfOutput.dense :287-298, BertIntermediate :330-342
// (dense + GELU) and BertOutput.dense :345-356.
#include "kernels.h"
#include "ptx.cuh"

namespace b200 {
namespace gemm {

constexpr int BM = 128;        // rows per CTA (256 per pair)
constexpr int BN = 256;        // columns per pair tile; each CTA stages 128 of them
constexpr int BK = 64, STAGES = 6;
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;      // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 32 KB per CTA
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;

__device__ __forceinline__ float gelu_erf(float x) {
    // x * Phi(x), erf by Abramowitz-Stegun 7.1.26 (abs err <= 1.5e-7): 2 MUFU + ~12 FMA-pipe ops
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
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const float* __restrict__ bias, const float* __restrict__ resid, void* __restrict__ out, int M, int N,
                 int K, int dbg_mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                     // [STAGES]
    uint64_t* empty = bars + STAGES;           // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;       // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
    const int lane = lane_id();
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int num_m = (M + 2 * BM - 1) / (2 * BM);
    const int num_n = N / BN;
    const int num_tiles = num_m * num_n;
    const int kblocks = K / BK;
    const int nstages = (dbg_mode >> 4) ? (dbg_mode >> 4) : STAGES;  // diagnostics may use a shorter ring
    const int dmode = dbg_mode & 0xF;

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
            mbar_init(&tempty[s], 2 * NUM_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair<512>(tmem_slot);
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();  // both CTAs' barriers are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // ---------------------------------------------------------------- TMA producer (both CTAs)
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int m_blk = tile / num_n, n_blk = tile % num_n;
                const int a_row = m_blk * (2 * BM) + cta_rank * BM;
                const int b_row = n_blk * BN + cta_rank * (BN / 2);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    if (dmode == 1) {  // diagnostics: no loads, the MMAs run on whatever is in smem
                        if (leader) mbar_arrive(&full[stage]);
                    } else {
                        if (leader) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
                        tma_load_2d_pair(sa, &tma_a, &full[stage], kb * BK, a_row);
                        tma_load_2d_pair(sa + A_BYTES, &tma_b, &full[stage], kb * BK, b_row);
                    }
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
            // ---------------------------------------------------------------- MMA issuer (leader only)
            constexpr uint32_t idesc = make_idesc_f16(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES;
                    if (dmode != 2) {  // diagnostics: mode 2 = loads only, no MMA
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            umma_f16_ss_pair(d_tmem, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc,
                                             (kb | k) != 0);
                        }
                    }
                    umma_commit_pair(&empty[stage], 0b11);
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair(&tfull[as], 0b11);
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // -------------------------------------------------------------------- epilogue warps (both CTAs)
        const int ew = warp & 3;           // TMEM lanes [32*ew, 32*ew+32)
        const int half = (warp - 4) >> 2;  // which 128 columns of the tile
        constexpr int CHUNKS = BN / 2 / 32;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            const int row = m_blk * (2 * BM) + cta_rank * BM + ew * 32 + lane;
            const bool row_ok = row < M;
            const int col0 = n_blk * BN + half * (BN / 2);
            const size_t row_off = static_cast<size_t>(row) * N + col0;
            float4 rq[8];  // residual chunk, prefetched: it does not depend on the MMA
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
            if (lane == 0) mbar_arrive_cluster(&tempty[as], 0);  // the leader's barrier gates the next MMA chain
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    // neither CTA may leave (or free TMEM) while its partner can still read its smem / signal its barriers
    __syncwarp();  // re-converge the single-lane role loops before the .aligned cluster barrier
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair<512>(tmem_base);
    }
}

}  // namespace gemm

template <int EPI>
static cudaError_t launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const float* bias, const float* resid,
                               void* out, int M, int N, int K, int sm_count, cudaStream_t stream, int dbg_mode) {
    const int tiles = ((M + 2 * gemm::BM - 1) / (2 * gemm::BM)) * (N / gemm::BN);
    int pairs = sm_count / 2;
    if (tiles < pairs) pairs = tiles;
    gemm::gemm_pair_kernel<EPI><<<2 * pairs, gemm::NUM_THREADS, gemm::SMEM_BYTES, stream>>>(ta, tb, bias, resid, out, M, N, K, dbg_mode);
    return cudaGetLastError();
}

cudaError_t gemm_init_device() {
    cudaError_t e;
    e = cudaFuncSetAttribute(gemm::gemm_pair_kernel<EPI_BIAS_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm::gemm_pair_kernel<EPI_BIAS_GELU_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm::gemm_pair_kernel<EPI_BIAS_RES_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES);
}

// tb: 2D map over W {K, N} with box {64, 128} (each CTA stages half of the tile's columns).  Bits 8+ of `epi`
// select a diagnostic mode (1 = no TMA loads, 2 = no MMA) used only by tools/gemm_diag.py.
cudaError_t launch_gemm(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const float* bias, const float* resid,
                             void* out, int M, int N, int K, int sm_count, cudaStream_t stream) {
    const int dbg_mode = epi >> 8;
    epi &= 0xFF;
    if (N % gemm::BN != 0 || K % gemm::BK != 0 || M <= 0) return cudaErrorInvalidValue;
    switch (epi) {
        case EPI_BIAS_F16: return launch_pair<EPI_BIAS_F16>(ta, tb, bias, resid, out, M, N, K, sm_count, stream, dbg_mode);
        case EPI_BIAS_GELU_F16: return launch_pair<EPI_BIAS_GELU_F16>(ta, tb, bias, resid, out, M, N, K, sm_count, stream, dbg_mode);
        case EPI_BIAS_RES_F32: return launch_pair<EPI_BIAS_RES_F32>(ta, tb, bias, resid, out, M, N, K, sm_count, stream, dbg_mode);
    }
    return cudaErrorInvalidValue;
}

}  // namespace b200
