// CTA-pair GEMM for sm_100a (tcgen05 cta_group::2):  C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias)
//
// A cluster of two CTAs (one TPC) owns a 256 x 256 output tile.  CTA r of the pair TMA-loads its own
// 128 rows of A and its own 128 rows (= output columns) of W per 64-wide k-block, so each SM pulls
// 32 KB per k-block from L2 instead of the 48 KB a lone 128x256 CTA needs (the first, single-CTA version of
// this kernel ran into the ~9.5 TB/s L2->SM fill limit at ~900 TFLOP/s; see profiles/README.md).
// The leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16); each
// CTA's TMEM receives its 128 rows x 256 fp32 columns, double buffered (2 x 256 columns), and each
// CTA's 8 epilogue warps drain them while the next tile's MMAs run.
//
// Epilogue data movement is TMA on both sides: a row-per-thread LDG/STG touches 32 different 128-byte
// lines per warp instruction (32 L1 wavefronts), which made the fp32 residual epilogue LSU-bound.  Instead
// each column half of the tile (4 warps, thread = row) works on 128-row x 128-byte staging buffers in
// shared memory (128B-swizzled, so a thread's 16-byte accesses are bank-conflict free per quarter warp):
//   EPI_BIAS_RES_SPLIT : a DMA thread TMA-loads the PRE-LayerNorm residual chunk [128 x 32] of y = hi + lo (two fp16
//                      arrays, 64-byte rows, 64B swizzle) into the buffer, the compute threads re-apply that LayerNorm
//                      from the row's statistics, add accumulator + bias, split the new value into hi' + lo' in place
//                      and accumulate the new row's (sum, M2) partial; the DMA thread TMA-stores both chunks back.
//   EPI_BIAS(_GELU)_F16 : compute threads apply the folded LayerNorm  rstd * acc + c  (see kernels.h),
//                      write fp16 [128 x 64] chunks, the DMA thread TMA-stores them.
// No LayerNorm kernel exists in the product path: the statistics travel as per-128-column partials next to the residual.
//
//   full[s]   (leader's)  : leader producer arrive.expect_tx(64 KB); both CTAs' TMA loads complete_tx on it
//   empty[s]  (per CTA)   : tcgen05.commit multicast to both CTAs once the MMAs that read stage s retire
//   tfull[a]  (per CTA)   : commit multicast after a tile's last MMA
//   tempty[a] (leader's)  : 2 x 8 epilogue warps arrive (the peer's through mapa / shared::cluster)
//   rfull[h][b] / cdone[h][b] (per CTA): staging buffer b of column half h is loaded/free  /  computed
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
constexpr int BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;      // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 32 KB per CTA
constexpr int EBUF_BYTES = 128 * 128;           // one staging buffer: 128 rows x 128 bytes
constexpr int OFF_EBUF = STAGES * STAGE_BYTES;  // [2 halves][2 buffers]
constexpr int OFF_BIAS = OFF_EBUF + 4 * EBUF_BYTES;  // float [2 parities][3: bias, gamma, beta][256]
constexpr int OFF_BAR = OFF_BIAS + 2 * 3 * 256 * 4;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;

// erf-GELU x * Phi(x) on two values at a time, erf by Abramowitz-Stegun 7.1.26 (abs err <= 1.5e-7): per value 2 MUFU (rcp, ex2)
// + the FMA-pipe part as packed fp32 (FFMA2 / FMUL2: same rounding per element, half the issue slots -- the GELU epilogue shares
// its SM sub-partitions with the TMA and MMA-issuing threads)
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
    const float a0 = fabsf(x0), a1 = fabsf(x1);
    const uint64_t ax = pack_f32x2(a0, a1);
    float u0, u1;
    unpack_f32x2(fma2_f32(ax, pack_f32x2(0.3275911f * 0.70710678118654752f, 0.3275911f * 0.70710678118654752f), pack_f32x2(1.0f, 1.0f)), u0, u1);
    float t0, t1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(u0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(u1));
    const uint64_t t = pack_f32x2(t0, t1);
    uint64_t p = fma2_f32(pack_f32x2(0.5f * 1.061405429f, 0.5f * 1.061405429f), t, pack_f32x2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
    p = fma2_f32(p, t, pack_f32x2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
    p = fma2_f32(p, t, pack_f32x2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
    p = fma2_f32(p, t, pack_f32x2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
    float z0, z1;
    unpack_f32x2(mul2_f32(mul2_f32(ax, pack_f32x2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f)), ax), z0, z1);
    const uint64_t q2 = mul2_f32(mul2_f32(p, t), pack_f32x2(ex2_approx(z0), ex2_approx(z1)));
    float q0, q1;
    unpack_f32x2(q2, q0, q1);
    const float phi0 = x0 >= 0.f ? 1.0f - q0 : q0, phi1 = x1 >= 0.f ? 1.0f - q1 : q1;
    unpack_f32x2(mul2_f32(pack_f32x2(x0, x1), pack_f32x2(phi0, phi1)), x0, x1);
}

__device__ __forceinline__ void gelu_quick2(float& x0, float& x1) {
    // CLIP's quick_gelu: x * sigmoid(1.702 x) = x / (1 + 2^(-1.702 log2(e) x)), two values at a time: 2 MUFU per value + 3 packed ops
    const uint64_t x = pack_f32x2(x0, x1);
    float z0, z1;
    unpack_f32x2(mul2_f32(x, pack_f32x2(-1.702f * 1.4426950408889634f, -1.702f * 1.4426950408889634f)), z0, z1);
    float u0, u1;
    unpack_f32x2(add2_f32(pack_f32x2(ex2_approx(z0), ex2_approx(z1)), pack_f32x2(1.0f, 1.0f)), u0, u1);
    float r0, r1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(u0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(u1));
    unpack_f32x2(mul2_f32(x, pack_f32x2(r0, r1)), x0, x1);
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_lo, const GemmEpi ep,
                 int M, int N, int K, int dbg_mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* full = bars;                     // [STAGES]
    uint64_t* empty = bars + STAGES;           // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;       // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
    uint64_t* rfull = bars + 2 * STAGES + 4;   // [2 halves][2 buffers]
    uint64_t* cdone = bars + 2 * STAGES + 8;   // [2 halves][2 buffers]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 12);
    float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS);

    griddep_launch_dependents();
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
#ifdef B200RT_DIAG  // tools/gemm_diag.py only: 1 = no TMA loads, 2 = no MMA; high nibble = ring length
    const int nstages = (dbg_mode >> 4) ? (dbg_mode >> 4) : STAGES;
    const int dmode = dbg_mode & 0xF;
#else
    constexpr int nstages = STAGES;
    constexpr int dmode = 0;
#endif
    // staging-buffer fills per tile per column half: fp32 [128x32] x 4, or fp16 [128x64] x 2
    constexpr int NBUF_PER_TILE = (EPI == EPI_BIAS_RES_SPLIT) ? 4 : 2;
    const float* __restrict__ bias = ep.bias;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tma_a);
        prefetch_tmap(&tma_b);
        prefetch_tmap(&tma_out);
        if constexpr (EPI == EPI_BIAS_RES_SPLIT) prefetch_tmap(&tma_lo);
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
        for (int s = 0; s < 4; ++s) {
            mbar_init(&rfull[s], 1);
            mbar_init(&cdone[s], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair<512>(tmem_slot);
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();  // both CTAs' barriers are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();  // everything above overlapped the previous kernel's tail; nothing it wrote has been read yet

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
    } else if (warp == 3) {
        if (lane < 2) {
            // ---------------------------------------------------------------- epilogue DMA threads (one per column half)
            const int h = lane;
            uint8_t* ebuf = smem + OFF_EBUF + h * (2 * EBUF_BYTES);
            int my_tiles = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) ++my_tiles;
            const int total = my_tiles * NBUF_PER_TILE;
            auto coords = [&](int g, int& c0, int& c1) {
                const int tile = pair + (g / NBUF_PER_TILE) * num_pairs;
                const int m_blk = tile / num_n, n_blk = tile % num_n;
                const int sub = g % NBUF_PER_TILE;
                c0 = n_blk * BN + h * (BN / 2) + sub * (EPI == EPI_BIAS_RES_SPLIT ? 32 : 64);
                c1 = m_blk * (2 * BM) + cta_rank * BM;
            };
            auto fill = [&](int g) {  // make buffer g&1 ready for the compute threads
                const int b = g & 1;
                if constexpr (EPI == EPI_BIAS_RES_SPLIT) {
                    int c0, c1;
                    coords(g, c0, c1);
                    mbar_arrive_expect_tx(&rfull[h * 2 + b], EBUF_BYTES);
                    tma_load_2d(ebuf + b * EBUF_BYTES, &tma_out, &rfull[h * 2 + b], c0, c1);                  // hi chunk
                    tma_load_2d(ebuf + b * EBUF_BYTES + EBUF_BYTES / 2, &tma_lo, &rfull[h * 2 + b], c0, c1);  // lo chunk
                } else {
                    mbar_arrive(&rfull[h * 2 + b]);  // nothing to load: just "buffer is free"
                }
            };
            if (total > 0) fill(0);
            if (total > 1) fill(1);
            for (int g = 0; g < total; ++g) {
                const int b = g & 1;
                mbar_wait(&cdone[h * 2 + b], (g >> 1) & 1);
                int c0, c1;
                coords(g, c0, c1);
                tma_store_2d(&tma_out, ebuf + b * EBUF_BYTES, c0, c1);
                if constexpr (EPI == EPI_BIAS_RES_SPLIT) tma_store_2d(&tma_lo, ebuf + b * EBUF_BYTES + EBUF_BYTES / 2, c0, c1);
                tma_store_commit();
                if (g + 2 < total) {
                    tma_store_wait_read<0>();  // the store has finished reading buffer b
                    fill(g + 2);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before the CTA may exit
        }
    } else if (warp >= 4) {
        // -------------------------------------------------------------------- epilogue compute warps (both CTAs)
        const int ew = warp & 3;         // TMEM lanes [32*ew, 32*ew+32)
        const int h = (warp - 4) >> 2;   // which 128 columns of the tile
        const int r = ew * 32 + lane;    // row within the CTA's 128 rows == index within the half
        const uint32_t ebuf = smem_u32(smem + OFF_EBUF + h * (2 * EBUF_BYTES));
        const uint32_t swz = static_cast<uint32_t>(r & 7);          // 128B swizzle: 16-byte chunk ^= row & 7
        const uint32_t swz64 = static_cast<uint32_t>((r >> 1) & 3);  // 64B swizzle (64-byte rows): chunk ^= (row >> 1) & 3
        int as = 0;
        uint32_t aphase = 0;
        uint32_t g = 0;  // staging-buffer use counter of this half
        int tpar = 0;
        // Per-tile inputs of this thread -- its row's statistic partials and its column's epilogue vector entries -- are
        // fetched one tile AHEAD: their global-load latency (two dependent round trips when done at the top of the tile) sat
        // on the epilogue's critical path and cost FFN1, whose GELU epilogue has no slack against the MMAs, ~15 %.
        const bool has_ln = ep.stats_in != nullptr;
        float4 pf_st[3] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        float pf_bias = 0.f, pf_g = 1.f, pf_b = 0.f;
        auto prefetch = [&](int tile_) {
            if (tile_ >= num_tiles) return;
            const int n_blk_ = tile_ % num_n;
            const int col_ = n_blk_ * BN + h * (BN / 2) + r;
            pf_bias = __ldg(bias + col_);
            if constexpr (EPI == EPI_BIAS_RES_SPLIT) {
                if (has_ln) {
                    pf_g = __ldg(ep.ln_gamma + col_);
                    pf_b = __ldg(ep.ln_beta + col_);
                }
            }
            if (has_ln) {
                const int grow_ = (tile_ / num_n) * (2 * BM) + cta_rank * BM + r;
                const float4* __restrict__ pp = reinterpret_cast<const float4*>(ep.stats_in + static_cast<size_t>(grow_) * STAT_PARTS);
                pf_st[0] = __ldg(pp);
                pf_st[1] = __ldg(pp + 1);
                pf_st[2] = __ldg(pp + 2);
            }
        };
        prefetch(pair);
        for (int tile = pair; tile < num_tiles; tile += num_pairs, tpar ^= 1) {
            const int n_blk = tile % num_n;
            const int grow = (tile / num_n) * (2 * BM) + cta_rank * BM + r;  // this thread's global row
            // stage this tile's per-column vectors (128 floats per half) once; double buffered across tiles
            const uint32_t sb = smem_u32(sbias + tpar * 768 + h * 128);  // bias; gamma at +256 floats, beta at +512
            sts32f(sb + r * 4, pf_bias);
            if constexpr (EPI == EPI_BIAS_RES_SPLIT) {
                sts32f(sb + 1024 + r * 4, pf_g);
                sts32f(sb + 2048 + r * 4, pf_b);
            }
            // statistics of the LayerNorm this epilogue folds (EPI 0/1) or re-applies to the residual (EPI 2): combine the
            // row's (sum, M2) partials
            float ln_mean = 0.f, ln_rstd = 1.f;
            if (has_ln) {
                const float2 p0 = make_float2(pf_st[0].x, pf_st[0].y), p1 = make_float2(pf_st[0].z, pf_st[0].w);
                const float2 p2 = make_float2(pf_st[1].x, pf_st[1].y), p3 = make_float2(pf_st[1].z, pf_st[1].w);
                const float2 p4 = make_float2(pf_st[2].x, pf_st[2].y), p5 = make_float2(pf_st[2].z, pf_st[2].w);
                constexpr float inv_n = 1.0f / (128.0f * STAT_PARTS);
                ln_mean = (((p0.x + p1.x) + (p2.x + p3.x)) + (p4.x + p5.x)) * inv_n;
                auto m2_of = [&](const float2& pj) {
                    const float d = pj.x * (1.0f / 128.0f) - ln_mean;
                    return fmaf(128.0f * d, d, pj.y);
                };
                const float m2 = ((m2_of(p0) + m2_of(p1)) + (m2_of(p2) + m2_of(p3))) + (m2_of(p4) + m2_of(p5));
                ln_rstd = rsqrtf(m2 * inv_n + ep.eps);
            }
            prefetch(tile + num_pairs);  // in flight while this tile is processed
            named_bar_sync(1 + h, 128);
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            // running statistics of this thread's 128 new values (EPI 2): count 32 c, mean, M2
            float run_mean = 0.f, run_m2 = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t acc[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN + h * (BN / 2) + c * 32, acc);
                const uint32_t b = g & 1;
                if constexpr (EPI == EPI_BIAS_RES_SPLIT) {
                    const uint32_t row_hi = ebuf + b * EBUF_BYTES + r * 64;
                    const uint32_t row_lo = row_hi + EBUF_BYTES / 2;
                    mbar_wait(&rfull[h * 2 + b], (g >> 1) & 1);  // residual chunk has landed
                    tmem_ld_wait();
                    if (c == 3) {  // last TMEM read of this tile: release the accumulator to the MMA warp early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(&tempty[as], 0);
                    }
                    float v[32];
                    float csum = 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {  // 8 columns per 16-byte chunk of hi / lo
                        const uint32_t off = (static_cast<uint32_t>(q) ^ swz64) << 4;
                        uint32_t hh[4], ll[4];
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hh[0]), "=r"(hh[1]), "=r"(hh[2]), "=r"(hh[3]) : "r"(row_hi + off) : "memory");
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ll[0]), "=r"(ll[1]), "=r"(ll[2]), "=r"(ll[3]) : "r"(row_lo + off) : "memory");
#pragma unroll
                        for (int e2 = 0; e2 < 2; ++e2) {  // 4 columns at a time (one float4 of each per-column vector)
                            const int cc = c * 32 + q * 8 + e2 * 4;
                            const float4 bv = lds128f(sb + cc * 4);
                            const float4 gm = lds128f(sb + 1024 + cc * 4);
                            const float4 bt = lds128f(sb + 2048 + cc * 4);
                            const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&hh[e2 * 2]));
                            const float2 h1 = __half22float2(*reinterpret_cast<const __half2*>(&hh[e2 * 2 + 1]));
                            const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&ll[e2 * 2]));
                            const float2 l1 = __half22float2(*reinterpret_cast<const __half2*>(&ll[e2 * 2 + 1]));
                            const float y0 = h0.x + l0.x, y1 = h0.y + l0.y, y2 = h1.x + l1.x, y3 = h1.y + l1.y;
                            // LN(y) = fmaf((y - mean) * rstd, gamma, beta); new value = acc + bias + LN(y)
                            const int i = q * 8 + e2 * 4;
                            v[i] = __uint_as_float(acc[i]) + bv.x + fmaf((y0 - ln_mean) * ln_rstd, gm.x, bt.x);
                            v[i + 1] = __uint_as_float(acc[i + 1]) + bv.y + fmaf((y1 - ln_mean) * ln_rstd, gm.y, bt.y);
                            v[i + 2] = __uint_as_float(acc[i + 2]) + bv.z + fmaf((y2 - ln_mean) * ln_rstd, gm.z, bt.z);
                            v[i + 3] = __uint_as_float(acc[i + 3]) + bv.w + fmaf((y3 - ln_mean) * ln_rstd, gm.w, bt.w);
                            csum += (v[i] + v[i + 1]) + (v[i + 2] + v[i + 3]);
                        }
                        // split v = hi' + lo' and write both back in place
                        uint32_t nh[4], nl[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float a0 = v[q * 8 + 2 * e], a1 = v[q * 8 + 2 * e + 1];
                            const __half2 hp = __floats2half2_rn(a0, a1);
                            const float2 hb = __half22float2(hp);
                            nh[e] = *reinterpret_cast<const uint32_t*>(&hp);
                            nl[e] = pack_half2(a0 - hb.x, a1 - hb.y);
                        }
                        sts128(row_hi + off, nh[0], nh[1], nh[2], nh[3]);
                        sts128(row_lo + off, nl[0], nl[1], nl[2], nl[3]);
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&cdone[h * 2 + b]);
                    ++g;
                    // merge this chunk's (mean, M2) into the running statistics (Chan et al.)
                    const float cmean = csum * (1.0f / 32.0f);
                    float cm2 = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float d = v[i] - cmean;
                        cm2 = fmaf(d, d, cm2);
                    }
                    const float n_run = 32.0f * c, n_new = n_run + 32.0f;
                    const float delta = cmean - run_mean;
                    run_mean += delta * (32.0f / n_new);
                    run_m2 += cm2 + delta * delta * (n_run * 32.0f / n_new);
                } else {
                    const uint32_t row_ptr = ebuf + b * EBUF_BYTES + r * 128;
                    if ((c & 1) == 0) mbar_wait(&rfull[h * 2 + b], (g >> 1) & 1);  // buffer is free
                    tmem_ld_wait();
                    if (c == 3) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(&tempty[as], 0);
                    }
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b4 = lds128f(sb + (c * 32 + 4 * j) * 4);  // bias (or W beta + b): broadcast read
                        // folded LayerNorm: rstd * acc + c -- the weight rows are centred at load time (sum_k W''[n,k] = 0), so
                        // the row mean cancels inside the tensor core: sum_k (y_k - mean) W''_nk = sum_k y_k W''_nk
                        v[4 * j] = fmaf(ln_rstd, __uint_as_float(acc[4 * j]), b4.x);
                        v[4 * j + 1] = fmaf(ln_rstd, __uint_as_float(acc[4 * j + 1]), b4.y);
                        v[4 * j + 2] = fmaf(ln_rstd, __uint_as_float(acc[4 * j + 2]), b4.z);
                        v[4 * j + 3] = fmaf(ln_rstd, __uint_as_float(acc[4 * j + 3]), b4.w);
                    }
                    if constexpr (EPI == EPI_BIAS_GELU_F16) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) gelu_erf2(v[j], v[j + 1]);
                    }
                    if constexpr (EPI == EPI_BIAS_QGELU_F16) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) gelu_quick2(v[j], v[j + 1]);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q) ^ swz;
                        sts128(row_ptr + (chunk << 4), pack_half2(v[8 * q], v[8 * q + 1]), pack_half2(v[8 * q + 2], v[8 * q + 3]),
                               pack_half2(v[8 * q + 4], v[8 * q + 5]), pack_half2(v[8 * q + 6], v[8 * q + 7]));
                    }
                    if (c & 1) {
                        fence_proxy_async_smem();
                        mbar_arrive(&cdone[h * 2 + b]);
                        ++g;
                    }
                }
            }
            if constexpr (EPI == EPI_BIAS_RES_SPLIT) {
                // this thread's 128 columns of the new row: one (sum, M2) partial
                if (ep.stats_out != nullptr)
                    ep.stats_out[static_cast<size_t>(grow) * (N / 128) + n_blk * 2 + h] = make_float2(run_mean * 128.0f, run_m2);
            }
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
static cudaError_t launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& tlo,
                               const GemmEpi& e, int M, int N, int K, int sm_count, cudaStream_t stream, int dbg_mode) {
    const int tiles = ((M + 2 * gemm::BM - 1) / (2 * gemm::BM)) * (N / gemm::BN);
    int pairs = sm_count / 2;
    if (tiles < pairs) pairs = tiles;
    return launch_pdl(gemm::gemm_pair_kernel<EPI>, dim3(2 * pairs), dim3(gemm::NUM_THREADS), gemm::SMEM_BYTES, stream, ta, tb, tout, tlo, e, M, N, K, dbg_mode);
}

cudaError_t gemm_init_device() {
    cudaError_t e;
    e = cudaFuncSetAttribute(gemm::gemm_pair_kernel<EPI_BIAS_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm::gemm_pair_kernel<EPI_BIAS_GELU_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm::gemm_pair_kernel<EPI_BIAS_QGELU_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm::gemm_pair_kernel<EPI_BIAS_RES_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES);
}

// Bits 8+ of `epi` select a diagnostic mode in B200RT_DIAG builds (tools/gemm_diag.py); they are ignored otherwise.
cudaError_t launch_gemm(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap* tlo,
                        const GemmEpi& e, int M, int N, int K, int sm_count, cudaStream_t stream) {
    const int dbg_mode = epi >> 8;
    epi &= 0xFF;
    if (N % gemm::BN != 0 || K % gemm::BK != 0 || M <= 0 || e.bias == nullptr) return cudaErrorInvalidValue;
    if (e.stats_in != nullptr && e.parts_in != STAT_PARTS) return cudaErrorInvalidValue;  // the LayerNorm'd width is HIDDEN
    switch (epi) {
        case EPI_BIAS_F16: return launch_pair<EPI_BIAS_F16>(ta, tb, tout, tout, e, M, N, K, sm_count, stream, dbg_mode);
        case EPI_BIAS_GELU_F16: return launch_pair<EPI_BIAS_GELU_F16>(ta, tb, tout, tout, e, M, N, K, sm_count, stream, dbg_mode);
        case EPI_BIAS_QGELU_F16: return launch_pair<EPI_BIAS_QGELU_F16>(ta, tb, tout, tout, e, M, N, K, sm_count, stream, dbg_mode);
        case EPI_BIAS_RES_SPLIT:
            if (tlo == nullptr || (e.stats_in != nullptr && (e.ln_gamma == nullptr || e.ln_beta == nullptr))) return cudaErrorInvalidValue;
            if (e.stats_out != nullptr && e.stats_out == e.stats_in) return cudaErrorInvalidValue;
            return launch_pair<EPI_BIAS_RES_SPLIT>(ta, tb, tout, *tlo, e, M, N, K, sm_count, stream, dbg_mode);
    }
    return cudaErrorInvalidValue;
}

}  // namespace b200
