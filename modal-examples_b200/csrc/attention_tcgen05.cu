// Multi-head self-attention for S <= 512, d = 64, on tcgen05 (sm_100a).
//
// One CTA per (item, head) keeps that head's K and V (<= 512 x 64 fp16 each) resident in shared memory and
// runs TWO independent "machines" over the item's 128-row query tiles (machine m takes tiles m, m+2, ...).
// A machine owns 256 TMEM columns organised as a ring of four 64-column slots, one Q buffer, a 2-slot P
// ring, its own MMA-issuing thread and one softmax warpgroup (thread = query row).  Its work is a stream of
// 64-key sub-blocks c = 0, 1, 2, ... (continuing across its tiles), software-pipelined per slot:
//
//   MMA thread : S_c = Q . K_c^T (128x64) into slot c&3, issued two sub-blocks ahead of the softmax
//                O_c = P_c . V_c overwrites S_c's slot as soon as P_c is in shared memory
//   softmax    : one pass over the 64 scores of its row in registers -> max m_c, P_c = exp2((S - m_c) k) as fp16
//                into a 128B-swizzled smem tile, l_c = sum(P_c); then, lagging two sub-blocks behind, folds
//                O_{c-2} into the running (max, sum, accumulator) in registers and frees the slot
//   tile end   : ctx = acc / l
//
// Every sub-block keeps its own (m, l, O): the accumulator in TMEM is never rescaled and the scores are
// read exactly once.  Keys >= len are masked to -inf before the max (exactly P = 0, matching HF's additive
// -inf mask); sub-blocks wholly past len are skipped.  k = log2(e) / sqrt(64).
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
constexpr int TILE_BYTES = 128 * D * 2;  // 16 KB: 128 rows x 128 B
constexpr int OFF_Q = 0;                            // [2 machines] x 16 KB
constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;       // 4 x 16 KB
constexpr int OFF_V = OFF_K + MAX_KB * TILE_BYTES;  // 4 x 16 KB
constexpr int OFF_P = OFF_V + MAX_KB * TILE_BYTES;  // [2 machines][2 slots] x 16 KB
constexpr int OFF_BAR = OFF_P + 4 * TILE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 384 + 1024;
constexpr int NUM_THREADS = 128 + 256;

// softmax_scale * log2(e) with softmax_scale = 1/sqrt(64)
constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;

__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tq, const int32_t* __restrict__ lens, __half* __restrict__ ctx,
                 int S, unsigned long long* __restrict__ dbg) {
    // dbg (diagnostics, normally NULL): CTA 0 records clock64() stamps; observer o in {softmax m0, softmax m1, mma m0,
    // mma m1}, 32 sub-blocks x 8 slots each (tools/attn_timeline.py prints them)
#define ATT_STAMP(o, c, slot)                                                                          \
    do {                                                                                               \
        if (dbg != nullptr && blockIdx.x == 0 && (c) < 32) dbg[((o) * 32 + (c)) * 8 + (slot)] = clock64(); \
    } while (0)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* k_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* q_full = bars + 2;      // [m]
    uint64_t* q_empty = bars + 4;     // [m]
    uint64_t* s_full = bars + 6;      // [m][4]  S_c landed in slot c&3
    uint64_t* o_full = bars + 14;     // [m][4]  O_c landed in slot c&3
    uint64_t* slot_free = bars + 22;  // [m][4]  O_c has been folded into the registers
    uint64_t* p_full = bars + 30;     // [m][2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);

    const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
    const int lane = lane_id();
    const int b = blockIdx.x / HEADS;
    const int h = blockIdx.x % HEADS;
    int len = lens[b];
    len = len < 1 ? 1 : (len > S ? S : len);
    const int nq = (S + QT - 1) / QT;
    const int nsb = (len + SB - 1) / SB;       // valid 64-key sub-blocks
    const int nkb = (nsb + 1) / 2;             // 128-key blocks holding them

    if (warp == 0 && elect_one()) prefetch_tmap(&tq);
    if (warp == 1 && elect_one()) {
        mbar_init(k_full, 1);
        mbar_init(v_full, 1);
        for (int m = 0; m < 2; ++m) {
            mbar_init(&q_full[m], 1);
            mbar_init(&q_empty[m], 1);
            for (int i = 0; i < 4; ++i) {
                mbar_init(&s_full[m * 4 + i], 1);
                mbar_init(&o_full[m * 4 + i], 1);
                mbar_init(&slot_free[m * 4 + i], 128);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&p_full[m * 2 + i], 128);
            }
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
            // ------------------------------------------------------------ TMA producer
            for (int qt = 0; qt < nq && qt < 2; ++qt) {
                mbar_arrive_expect_tx(&q_full[qt], TILE_BYTES);
                tma_load_3d(smem + OFF_Q + qt * TILE_BYTES, &tq, &q_full[qt], h * D, qt * QT, b);
            }
            mbar_arrive_expect_tx(k_full, nkb * TILE_BYTES);
            for (int j = 0; j < nkb; ++j)
                tma_load_3d(smem + OFF_K + j * TILE_BYTES, &tq, k_full, HIDDEN + h * D, j * KB, b);
            mbar_arrive_expect_tx(v_full, nkb * TILE_BYTES);
            for (int j = 0; j < nkb; ++j)
                tma_load_3d(smem + OFF_V + j * TILE_BYTES, &tq, v_full, 2 * HIDDEN + h * D, j * KB, b);
            for (int qt = 2; qt < nq; ++qt) {
                const int m = qt & 1, tl = qt >> 1;
                mbar_wait(&q_empty[m], (tl - 1) & 1);
                mbar_arrive_expect_tx(&q_full[m], TILE_BYTES);
                tma_load_3d(smem + OFF_Q + m * TILE_BYTES, &tq, &q_full[m], h * D, qt * QT, b);
            }
        }
    } else if (warp == 1 || warp == 3) {
        if (elect_one()) {
            // ------------------------------------------------------------ MMA issuer of machine m
            const int m = warp >> 1;                                    // warp 1 -> 0, warp 3 -> 1
            constexpr uint32_t idesc_s = make_idesc_f16(QT, SB);       // 128 x 64, both K-major
            constexpr uint32_t idesc_o = make_idesc_f16(QT, D, 0, 1);  // 128 x 64, B (= V) MN-major
            const uint32_t q_addr = smem_u32(smem + OFF_Q + m * TILE_BYTES);
            const uint32_t k_addr = smem_u32(smem + OFF_K);
            const uint32_t v_addr = smem_u32(smem + OFF_V);
            const uint32_t p_addr = smem_u32(smem + OFF_P + m * 2 * TILE_BYTES);
            const uint32_t tm = tmem_base + m * 256;
            const int ntiles = (nq - m + 1) / 2;          // query tiles of this machine
            const int total = ntiles * nsb;               // its stream of sub-blocks
            if (total > 0) mbar_wait(k_full, 0);
            auto issue_s = [&](int c) {
                const int t = c / nsb, sb = c - t * nsb;
                const uint32_t slot = c & 3;
                if (sb == 0) mbar_wait(&q_full[m], t & 1);
                if (c >= 4) mbar_wait(&slot_free[m * 4 + slot], ((c >> 2) - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    umma_f16_ss(tm + slot * SB, make_sw128_desc(q_addr + k * 32),
                                make_sw128_desc(k_addr + sb * (SB * 128) + k * 32), idesc_s, k != 0);
                }
                umma_commit(&s_full[m * 4 + slot]);
                if (sb == nsb - 1) umma_commit(&q_empty[m]);  // this tile's Q is no longer needed once these retire
            };
            if (total > 0) issue_s(0);
            if (total > 1) issue_s(1);
            if (total > 0) mbar_wait(v_full, 0);
            for (int c = 0; c < total; ++c) {
                const int t = c / nsb, sb = c - t * nsb;
                const uint32_t ps = c & 1, slot = c & 3;
                ATT_STAMP(2 + m, c, 0);
                mbar_wait(&p_full[m * 2 + ps], (c >> 1) & 1);
                tc_fence_after();
                ATT_STAMP(2 + m, c, 1);
#pragma unroll
                for (int kk = 0; kk < SB / 16; ++kk) {
                    const uint32_t a = p_addr + ps * TILE_BYTES + kk * 32;
                    const uint32_t bv = v_addr + (sb * SB + kk * 16) * 128;  // key row -> 128 B
                    umma_f16_ss(tm + slot * SB, make_sw128_desc(a), make_sw128_desc(bv), idesc_o, kk != 0);
                }
                umma_commit(&o_full[m * 4 + slot]);
                ATT_STAMP(2 + m, c, 2);
                if (c + 2 < total) issue_s(c + 2);
                ATT_STAMP(2 + m, c, 3);
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- softmax + epilogue warpgroup of machine m
        const int m = (warp - 4) >> 2;
        const int r = (warp & 3) * 32 + lane;  // query row within the tile == TMEM lane
        const uint32_t tm = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + m * 256;
        const uint32_t p_base = smem_u32(smem + OFF_P + m * 2 * TILE_BYTES);
        const uint32_t swz = static_cast<uint32_t>(r & 7);
        uint32_t c = 0;  // sub-block counter of this machine (continues across its tiles)
        const bool obs = (warp & 3) == 0 && lane == 0;
        for (int qt = m; qt < nq; qt += 2) {
            float acc[D];
#pragma unroll
            for (int i = 0; i < D; ++i) acc[i] = 0.f;
            float m_run = -INFINITY, l_run = 0.f;
            float mq[2] = {0.f, 0.f}, lq[2] = {0.f, 0.f};  // (m, l) of the two sub-blocks not yet folded
            // fold sub-block cj (its O is in slot cj&3) into the running state and free the slot
            // `need_wait`: in the steady state O_cj is already known to be complete (see the s_full note below)
            auto fold = [&](uint32_t cj, float m_j, float l_j, bool need_wait) {
                const float m_new = fmaxf(m_run, m_j);
                const float sc = ex2_approx((m_run - m_new) * kScaleLog2e);  // 0 for the tile's first sub-block
                const float w = ex2_approx((m_j - m_new) * kScaleLog2e);
                l_run = fmaf(w, l_j, l_run * sc);
                m_run = m_new;
                const uint32_t slot = cj & 3;
                if (need_wait) {
                    mbar_wait(&o_full[m * 4 + slot], (cj >> 2) & 1);
                    tc_fence_after();
                }
#pragma unroll
                for (int half = 0; half < D / 32; ++half) {
                    uint32_t o[32];
                    tmem_ld_32x32b_x32(tm + slot * SB + half * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc[half * 32 + e] = fmaf(w, __uint_as_float(o[e]), acc[half * 32 + e] * sc);
                }
                tc_fence_before();
                mbar_arrive(&slot_free[m * 4 + slot]);
            };
#pragma unroll 1
            for (int sb = 0; sb < nsb; ++sb, ++c) {
                const uint32_t slot = c & 3, ps = c & 1;
                if (obs) ATT_STAMP(m, c, 0);
                // The MMA thread issues PV_{c-2} before S_c and tcgen05.commit arrives only when ALL its earlier MMAs have
                // retired, so s_full(c) also tells us that O_{c-2} is complete and that P slot c&1 has been consumed:
                // one barrier wait per sub-block instead of three (each costs ~150-190 cycles even when already complete).
                mbar_wait(&s_full[m * 4 + slot], (c >> 2) & 1);
                tc_fence_after();
                if (obs) ATT_STAMP(m, c, 1);
                const int valid = len - sb * SB;  // keys [0, valid) of this sub-block are real (>= 1)
                uint32_t v[64];
                tmem_ld_32x32b_x32(tm + slot * SB, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                tmem_ld_32x32b_x32(tm + slot * SB + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                tmem_ld_wait();
                float mx = -INFINITY;
                if (valid >= SB) {
#pragma unroll
                    for (int e = 0; e < SB; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
                } else {
#pragma unroll
                    for (int e = 0; e < SB; ++e) {
                        if (e >= valid) v[e] = __float_as_uint(-INFINITY);
                        mx = fmaxf(mx, __uint_as_float(v[e]));
                    }
                }
                const float neg_ms = -mx * kScaleLog2e;
                if (obs) ATT_STAMP(m, c, 3);
                float ls0 = 0.f, ls1 = 0.f;
                const uint32_t row_ptr = p_base + ps * TILE_BYTES + r * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(v[q * 8 + 2 * e]), kScaleLog2e, neg_ms));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(v[q * 8 + 2 * e + 1]), kScaleLog2e, neg_ms));
                        ls0 += p0;
                        ls1 += p1;
                        pk[e] = pack_half2(p0, p1);
                    }
                    // keys 8q .. 8q+7 of row r -> 16-byte chunk q ^ (r & 7) of the row's 128 bytes
                    sts128(row_ptr + ((static_cast<uint32_t>(q) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
                }
                tc_fence_before();         // our TMEM reads of S_c precede the MMA that overwrites it with O_c
                fence_proxy_async_smem();  // P_c visible to the tensor core's async-proxy reads
                mbar_arrive(&p_full[m * 2 + ps]);
                if (obs) ATT_STAMP(m, c, 4);
                if (sb >= 2) fold(c - 2, mq[sb & 1], lq[sb & 1], false);
                if (obs) ATT_STAMP(m, c, 5);
                mq[sb & 1] = mx;
                lq[sb & 1] = ls0 + ls1;
            }
            // drain the (up to) two sub-blocks still in flight
            if (nsb >= 2) fold(c - 2, mq[nsb & 1], lq[nsb & 1], true);
            fold(c - 1, mq[(nsb - 1) & 1], lq[(nsb - 1) & 1], true);
            if (obs) ATT_STAMP(m, c - 1, 6);
            const float inv_l = 1.0f / l_run;
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
                             cudaStream_t stream, unsigned long long* dbg) {
    if (S < 1 || S > attn::MAX_KB * attn::KB || B < 1) return cudaErrorInvalidValue;
    attn::attention_kernel<<<B * HEADS, attn::NUM_THREADS, attn::SMEM_BYTES, stream>>>(tq, lens, ctx, S, dbg);
    return cudaGetLastError();
}

}  // namespace b200
