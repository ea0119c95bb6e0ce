// Multi-head self-attention for S <= 512, d = 64, on tcgen05 (sm_100a).
//
// One CTA per (item, head); it keeps that head's K and V (<= 512 x 64 fp16 each) resident in shared
// memory and walks the item's 128-row query tiles (Q double-buffered by TMA).  Per query tile:
//
//   S_j  = Q . K_j^T        one tcgen05.mma chain per 128-key block j -> TMEM columns [128j, 128j+128)
//   the scores are consumed in 64-key SUB-blocks sb = 2j+g by two softmax warpgroups (g = 0/1,
//   thread = query row): one pass over 64 TMEM columns held in registers -> sub-block max m_sb,
//   P_sb = exp2((S - m_sb) * scale*log2e) as fp16 into a 128B-swizzled smem tile (3-deep ring),
//   l_sb = sum(P_sb)
//   O_sb = P_sb . V_sb      accumulates into TMEM columns [64sb, 64sb+64)  (S_sb is dead by then)
//   epilogue: O = sum_sb w_sb O_sb / sum_sb w_sb l_sb,  w_sb = exp2((m_sb - max m) * scale*log2e);
//   thread (row, g) combines d-columns [32g, 32g+32) from all sub-blocks; (m, l) pairs are exchanged
//   through 8 KB of shared memory.
//
// Every sub-block keeps its own (m, l, O): no running-max rescale of a TMEM accumulator and no
// second pass over the scores.  Keys >= len are masked to -inf before the max (exactly P = 0, matching
// HF's additive -inf mask); sub-blocks wholly past len are skipped.
//
// Restates BertSelfAttention.forward (HF modeling_bert.py:143-207) for the TEI /embed path the
// reference calls at 06_gpu_and_ml/embeddings/text_embeddings_inference.py:100.
#include "kernels.h"
#include "ptx.cuh"

namespace b200 {
namespace attn {

constexpr int QT = 128;      // query rows per tile
constexpr int KB = 128;      // keys per S block (one MMA chain)
constexpr int SB = 64;       // keys per softmax / PV sub-block
constexpr int D = HEAD_DIM;  // 64
constexpr int MAX_KB = 4;    // S <= 512
constexpr int MAX_SB = 8;
constexpr int PRING = 3;
constexpr int TILE_BYTES = 128 * D * 2;  // 16 KB: 128 rows x 128 B
constexpr int OFF_Q = 0;                            // 2 x 16 KB
constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;       // 4 x 16 KB
constexpr int OFF_V = OFF_K + MAX_KB * TILE_BYTES;  // 4 x 16 KB
constexpr int OFF_P = OFF_V + MAX_KB * TILE_BYTES;  // 3 x 16 KB
constexpr int OFF_ML = OFF_P + PRING * TILE_BYTES;  // float2 [MAX_SB][128] = 8 KB
constexpr int OFF_BAR = OFF_ML + MAX_SB * QT * 8;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int NUM_SOFTMAX_THREADS = 256;
constexpr int NUM_THREADS = 128 + NUM_SOFTMAX_THREADS;

// softmax_scale * log2(e) with softmax_scale = 1/sqrt(64)
constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tq, const int32_t* __restrict__ lens, __half* __restrict__ ctx,
                 int S, unsigned long long* __restrict__ dbg) {
    // dbg (diagnostics only, normally NULL): CTA 0 records clock64() stamps, 16 slots per query tile for each of
    // three observers (softmax warpgroup 0 / 1 lane 0, MMA thread)
#define ATT_STAMP(obs, qt, slot)                                                            \
    do {                                                                                    \
        if (dbg != nullptr && blockIdx.x == 0) dbg[((obs) * 8 + (qt)) * 16 + (slot)] = clock64(); \
    } while (0)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* k_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* q_full = bars + 2;    // [2]
    uint64_t* q_empty = bars + 4;   // [2]
    uint64_t* s_full = bars + 6;    // [4]
    uint64_t* p_full = bars + 10;   // [3]
    uint64_t* p_empty = bars + 13;  // [3]
    uint64_t* o_full = bars + 16;
    uint64_t* tmem_free = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    float2* ml = reinterpret_cast<float2*>(smem + OFF_ML);  // [MAX_SB][128]

    const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
    const int lane = lane_id();
    const int b = blockIdx.x / HEADS;
    const int h = blockIdx.x % HEADS;
    int len = lens[b];
    len = len < 1 ? 1 : (len > S ? S : len);
    const int nq = (S + QT - 1) / QT;
    const int nsb = (len + SB - 1) / SB;   // valid 64-key sub-blocks
    const int nkb = (nsb + 1) / 2;         // 128-key blocks holding them

    if (warp == 0 && elect_one()) prefetch_tmap(&tq);
    if (warp == 1 && elect_one()) {
        mbar_init(k_full, 1);
        mbar_init(v_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
        }
        for (int i = 0; i < PRING; ++i) {
            mbar_init(&p_full[i], 128);
            mbar_init(&p_empty[i], 1);
        }
        for (int i = 0; i < MAX_KB; ++i) mbar_init(&s_full[i], 1);
        mbar_init(o_full, 1);
        mbar_init(tmem_free, NUM_SOFTMAX_THREADS);
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
            constexpr uint32_t idesc_s = make_idesc_f16(QT, KB);       // 128 x 128, both K-major
            constexpr uint32_t idesc_o = make_idesc_f16(QT, D, 0, 1);  // 128 x 64, B (= V) MN-major
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
                ATT_STAMP(2, qt, 0);
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
                ATT_STAMP(2, qt, 1);
                if (qt == 0) mbar_wait(v_full, 0);
                for (int sb = 0; sb < nsb; ++sb) {
                    const uint32_t pb = pcount % PRING;
                    mbar_wait(&p_full[pb], (pcount / PRING) & 1);
                    tc_fence_after();
                    ATT_STAMP(2, qt, 2 + sb);
#pragma unroll
                    for (int kk = 0; kk < SB / 16; ++kk) {
                        const uint32_t a = p_addr + pb * TILE_BYTES + kk * 32;
                        const uint32_t bv = v_addr + (sb * SB + kk * 16) * 128;  // key row -> 128 B
                        umma_f16_ss(tmem_base + sb * SB, make_sw128_desc(a), make_sw128_desc(bv), idesc_o, kk != 0);
                    }
                    umma_commit(&p_empty[pb]);
                    ++pcount;
                }
                umma_commit(o_full);
                ATT_STAMP(2, qt, 10);
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- softmax + epilogue warps
        const int e = warp - 4;
        const int g = e >> 2;                 // which 64-key half of every 128-key block
        const int r = (e & 3) * 32 + lane;    // query row within the tile == TMEM lane
        const uint32_t lane_base = static_cast<uint32_t>((e & 3) * 32) << 16;
        uint8_t* p_base = smem + OFF_P;
        const bool obs = (e & 3) == 0 && lane == 0;
        for (int qt = 0; qt < nq; ++qt) {
            if (obs) ATT_STAMP(g, qt, 0);
#pragma unroll 1
            for (int j = 0; j < nkb; ++j) {
                const int sb = 2 * j + g;
                if (sb >= nsb) break;
                mbar_wait(&s_full[j], qt & 1);
                tc_fence_after();
                if (obs) ATT_STAMP(g, qt, 1 + 2 * j);
                const int valid = len - sb * SB;  // keys [0, valid) of this sub-block are real (>= 1)
                uint32_t v[64];
                tmem_ld_32x32b_x32(tmem_base + lane_base + sb * SB, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                tmem_ld_32x32b_x32(tmem_base + lane_base + sb * SB + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                tmem_ld_wait();
                float mx = -INFINITY;
                if (valid >= SB) {
#pragma unroll
                    for (int i = 0; i < SB; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < SB; ++i) {
                        if (i >= valid) v[i] = __float_as_uint(-INFINITY);
                        mx = fmaxf(mx, __uint_as_float(v[i]));
                    }
                }
                const float neg_ms = -mx * kScaleLog2e;
                // the P ring is shared by both warpgroups and is consumed in sub-block order
                const uint32_t pidx = static_cast<uint32_t>(qt) * nsb + sb;
                const uint32_t pb = pidx % PRING;
                mbar_wait(&p_empty[pb], ((pidx / PRING) & 1) ^ 1);
                float lsum = 0.f;
                uint8_t* row_ptr = p_base + pb * TILE_BYTES + r * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t pk[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(v[q * 8 + 2 * i]), kScaleLog2e, neg_ms));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(v[q * 8 + 2 * i + 1]), kScaleLog2e, neg_ms));
                        lsum += p0 + p1;
                        pk[i] = pack_half2(p0, p1);
                    }
                    // keys 8q .. 8q+7 of row r -> 16-byte chunk q ^ (r & 7) of the row's 128 bytes
                    *reinterpret_cast<uint4*>(row_ptr + ((q ^ (r & 7)) * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                ml[sb * QT + r] = make_float2(mx, lsum);
                tc_fence_before();         // our TMEM reads of S_sb precede the MMA that overwrites it with O_sb
                fence_proxy_async_smem();  // P_sb visible to the tensor core's async-proxy reads
                mbar_arrive(&p_full[pb]);
                if (obs) ATT_STAMP(g, qt, 2 + 2 * j);
            }
            if (obs) ATT_STAMP(g, qt, 9);
            // every (m, l) of this tile is in smem once all softmax threads are here
            named_bar_sync(1, NUM_SOFTMAX_THREADS);
            if (obs) ATT_STAMP(g, qt, 10);
            float m_all = -INFINITY;
            for (int sb = 0; sb < nsb; ++sb) m_all = fmaxf(m_all, ml[sb * QT + r].x);
            float w[MAX_SB];
            float L = 0.f;
#pragma unroll
            for (int sb = 0; sb < MAX_SB; ++sb) {
                if (sb < nsb) {
                    const float2 t = ml[sb * QT + r];
                    w[sb] = ex2_approx((t.x - m_all) * kScaleLog2e);
                    L = fmaf(w[sb], t.y, L);
                } else {
                    w[sb] = 0.f;
                }
            }
            const float inv_l = 1.0f / L;
            mbar_wait(o_full, qt & 1);
            tc_fence_after();
            if (obs) ATT_STAMP(g, qt, 11);
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
            for (int sb = 0; sb < MAX_SB; ++sb) {
                if (sb < nsb) {
                    uint32_t o[32];
                    tmem_ld_32x32b_x32(tmem_base + lane_base + sb * SB + g * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = fmaf(w[sb], __uint_as_float(o[i]), acc[i]);
                }
            }
            tc_fence_before();
            mbar_arrive(tmem_free);  // also orders our reads of `ml` before the next tile's writes
            if (obs) ATT_STAMP(g, qt, 12);
            const int q_row = qt * QT + r;
            if (q_row < S) {
                uint4* dst = reinterpret_cast<uint4*>(ctx + (static_cast<size_t>(b) * S + q_row) * HIDDEN + h * D + g * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
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
                             cudaStream_t stream, unsigned long long* dbg) {
    if (S < 1 || S > attn::MAX_KB * attn::KB || B < 1) return cudaErrorInvalidValue;
    attn::attention_kernel<<<B * HEADS, attn::NUM_THREADS, attn::SMEM_BYTES, stream>>>(tq, lens, ctx, S, dbg);
    return cudaGetLastError();
}

}  // namespace b200
