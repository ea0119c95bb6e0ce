// Multi-head self-attention for S <= 512, d = 64, on tcgen05 (sm_100a).
//
// One CTA per (item, head); it keeps that head's K and V (<= 512 x 64 fp16 each) resident in shared
// memory and walks the item's 128-row query tiles (Q double-buffered by TMA).
//
//   S_j = Q . K_j^T   for every 128-key block j: 128x128 fp32 in TMEM columns [128j, 128j+128)
//   softmax warps (thread = query row) read S_j once, take the block max m_j, write
//   P_j = exp2((S_j - m_j) * scale*log2e) as fp16 into a 128B-swizzled smem tile, keep l_j = sum(P_j)
//   O_j = P_j . V_j   accumulates into TMEM columns [128j, 128j+64) (S_j is dead by then)
//   epilogue: O = sum_j w_j O_j / sum_j w_j l_j with w_j = exp2((m_j - max_j m_j) * scale*log2e)
//
// Because every block keeps its own (m_j, l_j, O_j) there is no running-max rescale of an
// accumulator in TMEM and no second pass over the scores.  Keys >= len are masked to -inf before
// the max (exactly P = 0, matching HF's additive -inf mask); key blocks wholly past len are skipped.
//
// Restates BertSelfAttention.forward (HF modeling_bert.py:143-207) for the TEI /embed path the
// reference calls at 06_gpu_and_ml/embeddings/text_embeddings_inference.py:100.
#include "kernels.h"
#include "ptx.cuh"

namespace b200 {
namespace attn {

constexpr int QT = 128;     // query rows per tile
constexpr int KB = 128;     // keys per block
constexpr int D = HEAD_DIM;  // 64
constexpr int MAX_KB = 4;   // S <= 512
constexpr int TILE_BYTES = 128 * D * 2;  // 16 KB: 128 rows x 128 B
constexpr int OFF_Q = 0;                          // 2 x 16 KB
constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;     // 4 x 16 KB
constexpr int OFF_V = OFF_K + MAX_KB * TILE_BYTES;  // 4 x 16 KB
constexpr int OFF_P = OFF_V + MAX_KB * TILE_BYTES;  // 2 x 32 KB
constexpr int OFF_BAR = OFF_P + 2 * 2 * TILE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int NUM_THREADS = 256;

// softmax_scale * log2(e) with softmax_scale = 1/sqrt(64)
constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tq, const int32_t* __restrict__ lens, __half* __restrict__ ctx,
                 int S) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* k_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* q_full = bars + 2;    // [2]
    uint64_t* q_empty = bars + 4;   // [2]
    uint64_t* s_full = bars + 6;    // [4]
    uint64_t* p_full = bars + 10;   // [2]
    uint64_t* p_empty = bars + 12;  // [2]
    uint64_t* o_full = bars + 14;
    uint64_t* tmem_free = bars + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
    const int lane = lane_id();
    const int b = blockIdx.x / HEADS;
    const int h = blockIdx.x % HEADS;
    int len = lens[b];
    len = len < 1 ? 1 : (len > S ? S : len);
    const int nq = (S + QT - 1) / QT;
    const int nkb = (len + KB - 1) / KB;

    if (warp == 0 && elect_one()) prefetch_tmap(&tq);
    if (warp == 1 && elect_one()) {
        mbar_init(k_full, 1);
        mbar_init(v_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
            mbar_init(&p_full[i], 128);
            mbar_init(&p_empty[i], 1);
        }
        for (int i = 0; i < MAX_KB; ++i) mbar_init(&s_full[i], 1);
        mbar_init(o_full, 1);
        mbar_init(tmem_free, 128);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------------------------------------ TMA producer
            mbar_arrive_expect_tx(&q_full[0], TILE_BYTES);
            tma_load_3d(smem + OFF_Q, &tq, &q_full[0], h * D, 0, b);
            mbar_arrive_expect_tx(k_full, nkb * TILE_BYTES);
            for (int j = 0; j < nkb; ++j)
                tma_load_3d(smem + OFF_K + j * TILE_BYTES, &tq, k_full, HIDDEN + h * D, j * KB, b);
            mbar_arrive_expect_tx(v_full, nkb * TILE_BYTES);
            for (int j = 0; j < nkb; ++j)
                tma_load_3d(smem + OFF_V + j * TILE_BYTES, &tq, v_full, 2 * HIDDEN + h * D, j * KB, b);
            for (int qt = 1; qt < nq; ++qt) {
                const int slot = qt & 1;
                mbar_wait(&q_empty[slot], ((qt >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&q_full[slot], TILE_BYTES);
                tma_load_3d(smem + OFF_Q + slot * TILE_BYTES, &tq, &q_full[slot], h * D, qt * QT, b);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc_s = make_idesc_f16(QT, KB);         // 128 x 128, both K-major
            constexpr uint32_t idesc_o = make_idesc_f16(QT, D, 0, 1);    // 128 x 64, B (= V) MN-major
            const uint32_t k_addr = smem_u32(smem + OFF_K);
            const uint32_t v_addr = smem_u32(smem + OFF_V);
            const uint32_t p_addr = smem_u32(smem + OFF_P);
            uint32_t pcount = 0;
            mbar_wait(k_full, 0);
            for (int qt = 0; qt < nq; ++qt) {
                const int slot = qt & 1;
                mbar_wait(&q_full[slot], (qt >> 1) & 1);
                if (qt > 0) mbar_wait(tmem_free, (qt - 1) & 1);
                tc_fence_after();
                const uint32_t q_addr = smem_u32(smem + OFF_Q + slot * TILE_BYTES);
                for (int j = 0; j < nkb; ++j) {
#pragma unroll
                    for (int k = 0; k < D / 16; ++k) {
                        umma_f16_ss(tmem_base + j * KB, make_sw128_desc(q_addr + k * 32),
                                    make_sw128_desc(k_addr + j * TILE_BYTES + k * 32), idesc_s, k != 0);
                    }
                    umma_commit(&s_full[j]);
                }
                umma_commit(&q_empty[slot]);
                if (qt == 0) mbar_wait(v_full, 0);
                for (int j = 0; j < nkb; ++j) {
                    const uint32_t pb = pcount & 1;
                    mbar_wait(&p_full[pb], (pcount >> 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < KB / 16; ++kk) {
                        const uint32_t a = p_addr + pb * (2 * TILE_BYTES) + (kk >> 2) * TILE_BYTES + (kk & 3) * 32;
                        const uint32_t bv = v_addr + j * TILE_BYTES + kk * (16 * 128);
                        umma_f16_ss(tmem_base + j * KB, make_sw128_desc(a), make_sw128_desc(bv), idesc_o, kk != 0);
                    }
                    umma_commit(&p_empty[pb]);
                    ++pcount;
                }
                umma_commit(o_full);
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- softmax + epilogue warps
        const int ew = warp - 4;
        const int r = ew * 32 + lane;  // query row within the tile == TMEM lane
        const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
        uint8_t* p_base = smem + OFF_P;
        uint32_t pcount = 0;
        for (int qt = 0; qt < nq; ++qt) {
            float m_blk[MAX_KB], l_blk[MAX_KB];
#pragma unroll
            for (int j = 0; j < MAX_KB; ++j) {
                m_blk[j] = -INFINITY;
                l_blk[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < MAX_KB; ++j) {
                if (j < nkb) {
                    mbar_wait(&s_full[j], qt & 1);
                    tc_fence_after();
                    const uint32_t pb = pcount & 1;
                    const int valid = len - j * KB;  // keys [0, valid) of this block are real (>= 1)
                    // pass 1: block max over the 128 scores of this row
                    float mx = -INFINITY;
#pragma unroll 1
                    for (int c = 0; c < KB / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tmem_base + lane_base + j * KB + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float s = (c * 32 + i < valid) ? __uint_as_float(v[i]) : -INFINITY;
                            mx = fmaxf(mx, s);
                        }
                    }
                    const float neg_ms = -mx * kScaleLog2e;
                    mbar_wait(&p_empty[pb], ((pcount >> 1) & 1) ^ 1);
                    // pass 2: P = exp2(s*c - m*c), row sum, fp16 pack, swizzled smem store
                    float lsum = 0.f;
                    uint8_t* p_tile = p_base + pb * (2 * TILE_BYTES);
#pragma unroll 1
                    for (int c = 0; c < KB / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tmem_base + lane_base + j * KB + c * 32, v);
                        tmem_ld_wait();
                        uint32_t packed[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float s0 = (c * 32 + i < valid) ? __uint_as_float(v[i]) : -INFINITY;
                            const float s1 = (c * 32 + i + 1 < valid) ? __uint_as_float(v[i + 1]) : -INFINITY;
                            const float p0 = ex2_approx(fmaf(s0, kScaleLog2e, neg_ms));
                            const float p1 = ex2_approx(fmaf(s1, kScaleLog2e, neg_ms));
                            lsum += p0 + p1;
                            packed[i >> 1] = pack_half2(p0, p1);
                        }
                        // keys c*32 .. c*32+31 of row r -> sub-tile (c>>1), 16B chunks ((c&1)*4 + q) ^ (r&7)
                        uint8_t* row_ptr = p_tile + (c >> 1) * TILE_BYTES + r * 128;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
                            *reinterpret_cast<uint4*>(row_ptr + chunk * 16) =
                                make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                        }
                    }
                    m_blk[j] = mx;
                    l_blk[j] = lsum;
                    tc_fence_before();          // our TMEM reads of S_j precede the MMA that overwrites it with O_j
                    fence_proxy_async_smem();   // P_j visible to the tensor core's async-proxy reads
                    mbar_arrive(&p_full[pb]);
                    ++pcount;
                }
            }
            // combine the per-block partial results
            float m_all = m_blk[0];
#pragma unroll
            for (int j = 1; j < MAX_KB; ++j) m_all = fmaxf(m_all, m_blk[j]);
            float w[MAX_KB];
            float L = 0.f;
#pragma unroll
            for (int j = 0; j < MAX_KB; ++j) {
                w[j] = (j < nkb) ? ex2_approx((m_blk[j] - m_all) * kScaleLog2e) : 0.f;
                L += w[j] * l_blk[j];
            }
            const float inv_l = 1.0f / L;
            mbar_wait(o_full, qt & 1);
            tc_fence_after();
            float acc[D];
#pragma unroll
            for (int i = 0; i < D; ++i) acc[i] = 0.f;
#pragma unroll
            for (int j = 0; j < MAX_KB; ++j) {
                if (j < nkb) {
#pragma unroll
                    for (int c = 0; c < D / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tmem_base + lane_base + j * KB + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[c * 32 + i] = fmaf(w[j], __uint_as_float(v[i]), acc[c * 32 + i]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tmem_free);
            const int q_row = qt * QT + r;
            if (q_row < S) {
                uint4* dst = reinterpret_cast<uint4*>(ctx + (static_cast<size_t>(b) * S + q_row) * HIDDEN + h * D);
#pragma unroll
                for (int i = 0; i < D / 8; ++i) {
                    dst[i] = make_uint4(pack_half2(acc[8 * i] * inv_l, acc[8 * i + 1] * inv_l),
                                        pack_half2(acc[8 * i + 2] * inv_l, acc[8 * i + 3] * inv_l),
                                        pack_half2(acc[8 * i + 4] * inv_l, acc[8 * i + 5] * inv_l),
                                        pack_half2(acc[8 * i + 6] * inv_l, acc[8 * i + 7] * inv_l));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace attn

cudaError_t attention_init_device() {
    return cudaFuncSetAttribute(attn::attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES);
}

cudaError_t launch_attention(const CUtensorMap& tq, const int32_t* lens, __half* ctx, int B, int S,
                             cudaStream_t stream) {
    if (S < 1 || S > attn::MAX_KB * attn::KB || B < 1) return cudaErrorInvalidValue;
    attn::attention_kernel<<<B * HEADS, attn::NUM_THREADS, attn::SMEM_BYTES, stream>>>(tq, lens, ctx, S);
    return cudaGetLastError();
}

}  // namespace b200
