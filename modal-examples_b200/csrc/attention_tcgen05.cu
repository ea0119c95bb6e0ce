// Multi-head self-attention for S <= 512, d = 64, on tcgen05 (sm_100a).
//
// One CTA per (item, head) keeps that head's K and V (<= 512 x 64 fp16 each) resident in shared memory and
// streams the item's 128-row query tiles through them as a sequence of 64-key sub-blocks c = 0, 1, 2, ...
// (8 per 512-key tile, continuing across tiles).  What bounds the kernel is the exp: one warp's MUFU stream sustains an
// EX2 per ~11 cycles and its 64-score exp phase takes ~660 cycles alone, 1100+ when the same SM sub-partition also
// issues an accumulator fold (tools/ubench/spin_cost.cu).  So everything that is not the exp is kept off the exp warps:
//
//   TMEM         : two 64-column O accumulators (tile t -> t&1) + a ring of 2 NEXP 64-column S slots
//   S thread     : S_c = Q . K_c^T (128x64) into S slot c % NSLOT, as far ahead as free slots allow
//   P.V thread   : O_t += P_c . V_c accumulates IN TMEM across the tile's sub-blocks (no per-sub-block fold; a second
//                  issuing thread because one thread doing both chains was itself the bottleneck at ~1270 cycles/sub-block)
//   NEXP exp WGs : warpgroup w takes sub-blocks c = w (mod NEXP), thread = query row: one TMEM read of the 64 scores,
//                  row max, P_c = exp2((S - m) k) as fp16 into its own 128B-swizzled smem tile, partial row sum.
//                  m is the row's RUNNING reference maximum, handed from sub-block to sub-block through shared memory;
//                  it only moves when the new maximum exceeds it by more than 2^8 in the exp2 domain (P <= 256 is exact
//                  enough in fp16 and the sums are fp32), so the accumulator in TMEM is rescaled (tcgen05.ld/st by
//                  the exp warp that saw the jump) a handful of times per tile instead of once per sub-block.
//                  Product configuration: NEXP = 3, free-running (640 threads; the warpgroups re-balance registers).
//   epilogue WG  : once per tile: l = sum of the warpgroups' partial sums brought to the final m, ctx = O / l as fp16
//                  through a swizzled staging tile (the tile's own, now idle, Q buffer) and one TMA store --
//                  row-per-thread global stores cost 32 LSU wavefronts each and stalled the exp warps' shared-memory
//                  stores behind them.
//
// Keys >= len are masked to -inf before the max (exactly P = 0, matching HF's additive -inf mask); sub-blocks wholly
// past len are skipped.  k = log2(e) / sqrt(64).
//
// Batches with fewer (item, head) units than SMs launch one CTA per (item, head, query tile) instead.  Either way sub-block
// c of tile t goes to warpgroup (t nsb + c) mod NEXP and sees the same reference maxima, so the output does not depend
// on the launch shape: an item's embedding is bit-identical in any batch (tests: scheduler / size-independent properties).
//
// Restates BertSelfAttention.forward (HF modeling_bert.py:143-207) for the TEI /embed path the
// reference calls at 06_gpu_and_ml/embeddings/text_embeddings_inference.py:100.
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace b200 {
namespace attn {

constexpr int QT = 128;      // query rows per tile
constexpr int KB = 128;      // keys per K/V smem tile (TMA box)
constexpr int SB = 64;       // keys per sub-block
constexpr int D = HEAD_DIM;  // 64
constexpr int MAX_KB = 4;    // S <= 512
constexpr int MAX_NQ = 4;
constexpr int MAX_NEXP = 3;  // exp warpgroups == P buffers == accumulators per tile (template parameter NEXP: 2 or 3)
constexpr uint32_t TM_O = 0;        // accumulators: tile parity -> 2 x 64 columns
constexpr uint32_t TM_S = 2 * D;    // S ring: NSLOT x 64 columns
constexpr int TILE_BYTES = 128 * D * 2;  // 16 KB: 128 rows x 128 B
constexpr int OFF_Q = 0;                            // 2 x 16 KB (double buffered across tiles); doubles as the epilogue's
                                                    // staging tile once the tile's MMAs have retired
constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;       // 4 x 16 KB
constexpr int OFF_V = OFF_K + MAX_KB * TILE_BYTES;  // 4 x 16 KB
constexpr int OFF_MR = OFF_V + MAX_KB * TILE_BYTES; // float [MAX_NEXP][128]: reference max after warpgroup w's latest sub-block
constexpr int OFF_LS = OFF_MR + MAX_NEXP * QT * 4;  // float2 [MAX_NQ][MAX_NEXP][128]: (reference max, partial row sum) per tile
constexpr int OFF_BAR = OFF_LS + MAX_NQ * MAX_NEXP * QT * 8;
constexpr int OFF_P = (OFF_BAR + 512 + 1023) / 1024 * 1024;  // NEXP x 16 KB, 1024-aligned for the 128B swizzle
static_assert(OFF_P % 1024 == 0 && OFF_V % 1024 == 0, "swizzled tiles must be 1024-byte aligned");
constexpr int smem_bytes(int nexp) { return OFF_P + nexp * TILE_BYTES + 1024; }
constexpr int num_threads(int nexp) { return 128 + nexp * 128 + 128; }
static_assert(smem_bytes(MAX_NEXP) <= 232448, "shared memory");
constexpr int BAR_TOKEN = 1;               // named barriers 1 .. NEXP: MUFU token
constexpr int BAR_MAX = 1 + MAX_NEXP;      // named barriers 4 .. 6: running-max hand-off
constexpr int BAR_EPI = 1 + 2 * MAX_NEXP;  // the epilogue warpgroup's own barrier

// softmax_scale * log2(e) with softmax_scale = 1/sqrt(64)
constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;
// the reference max follows the true max only when it is exceeded by more than this (raw score units): P <= 2^8
constexpr float kRescaleThreshold = 8.0f / kScaleLog2e;

// NEXP exp warpgroups; USE_TOKEN: their exp phases take turns on the MUFU.  With NEXP = 3 the CTA has 640 threads (96
// registers each at launch) and the warpgroups re-balance WITHIN that pool of 640 x 96: 120 for the exp warps, 80 for the
// epilogue, 40 for the rest (asking for more than the launch allocation holds makes setmaxnreg.inc spin forever).
template <int NEXP, bool USE_TOKEN>
__global__ void __launch_bounds__(num_threads(NEXP), 1)
attention_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tctx,
                 const int32_t* __restrict__ lens, int S, int split, unsigned long long* __restrict__ dbg) {
    // dbg (diagnostics, normally NULL): CTA 0 records clock64() stamps; observer o in {exp WG 0..2, epilogue WG, P.V thread},
    // 32 sub-blocks x 8 slots each (tools/attn_timeline.py prints them)
#define ATT_STAMP(o, c, slot)                                                                          \
    do {                                                                                               \
        if (dbg != nullptr && blockIdx.x == 0 && (c) < 32) dbg[((o) * 32 + (c)) * 8 + (slot)] = clock64(); \
    } while (0)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int NSLOT = 2 * NEXP;  // TMEM S slots of 64 columns (<= 6) == barrier ring: every barrier has one waiter
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* k_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* q_full = bars + 2;       // [2]
    uint64_t* q_empty = bars + 4;      // [2]
    uint64_t* o_done = bars + 6;       // [2]  the tile's last P.V has retired
    uint64_t* o_free = bars + 8;       // [2]  the epilogue has read the accumulator
    constexpr int NRING = NSLOT;     // (== lcm(NSLOT, NEXP): a warpgroup meets the phases of its barriers in order)
    uint64_t* s_full = bars + 10;      // [NRING]  S_c landed
    uint64_t* s_free = bars + 16;      // [NRING]  S_c is in the exp warpgroup's registers
    uint64_t* p_full = bars + 22;      // [NEXP]  P buffer w written
    uint64_t* pv_done = bars + 25;     // [NEXP]  the P.V reading P buffer w has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

    const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
    const int lane = lane_id();
    // split != 0 (small batches): one CTA per (item, head, query tile) instead of per (item, head) -- K and V are then
    // loaded once per tile, but a single item spreads over 48 SMs instead of 12
    const int nq_all = (S + QT - 1) / QT;
    int unit = blockIdx.x, t0 = 0;
    if (split) {
        t0 = unit % nq_all;
        unit /= nq_all;
    }
    const int b = unit / HEADS;
    const int h = unit % HEADS;
    int len = lens[b];
    len = len < 1 ? 1 : (len > S ? S : len);
    const int nq = split ? 1 : nq_all;    // query tiles of this CTA: t0 .. t0 + nq - 1
    const int nsb = (len + SB - 1) / SB;  // valid 64-key sub-blocks per tile
    const int nkb = (nsb + 1) / 2;        // 128-key K/V tiles holding them
    const int total = nq * nsb;           // the CTA's stream of sub-blocks
    // Sub-block c of this CTA is the item's sub-block t0 nsb + c and goes to warpgroup (c_off + c) % NEXP: the split launch
    // gives every sub-block to the warpgroup the unsplit one would (the row sums are grouped by warpgroup, so that an
    // item's embedding is bit-identical in any batch)
    const int c_off = (t0 * nsb) % NEXP;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tq);
        prefetch_tmap(&tctx);
    }
    if (warp == 1 && elect_one()) {
        mbar_init(k_full, 1);
        mbar_init(v_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
            mbar_init(&o_done[i], 1);
            mbar_init(&o_free[i], 128);
        }
        for (int i = 0; i < NRING; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 128);
        }
        for (int i = 0; i < NEXP; ++i) {
            mbar_init(&p_full[i], 128);
            mbar_init(&pv_done[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // (each role's code must be dominated by its own setmaxnreg for ptxas to allocate against the new budget)
    if (warp < 4) {
      if constexpr (NEXP == 3) setmaxnreg_dec<40>();
      if (warp == 0) {
        if (elect_one()) {
            // ------------------------------------------------------------ TMA producer
            for (int t = 0; t < nq && t < 2; ++t) {
                mbar_arrive_expect_tx(&q_full[t], TILE_BYTES);
                tma_load_3d(smem + OFF_Q + t * TILE_BYTES, &tq, &q_full[t], h * D, (t0 + t) * QT, b);
            }
            mbar_arrive_expect_tx(k_full, nkb * TILE_BYTES);
            for (int j = 0; j < nkb; ++j)
                tma_load_3d(smem + OFF_K + j * TILE_BYTES, &tq, k_full, HIDDEN + h * D, j * KB, b);
            mbar_arrive_expect_tx(v_full, nkb * TILE_BYTES);
            for (int j = 0; j < nkb; ++j)
                tma_load_3d(smem + OFF_V + j * TILE_BYTES, &tq, v_full, 2 * HIDDEN + h * D, j * KB, b);
            for (int t = 2; t < nq; ++t) {
                const int qb = t & 1;
                mbar_wait(&q_empty[qb], ((t >> 1) - 1) & 1);  // tile t-2 has left the buffer (epilogue)
                mbar_arrive_expect_tx(&q_full[qb], TILE_BYTES);
                tma_load_3d(smem + OFF_Q + qb * TILE_BYTES, &tq, &q_full[qb], h * D, (t0 + t) * QT, b);
            }
        }
      } else if (warp == 1) {
        if (elect_one()) {
            // ------------------------------------------------------------ S issuer: S_c = Q . K_c^T into S slot c&3
            constexpr uint32_t idesc_s = make_idesc_f16(QT, SB);  // 128 x 64, both K-major
            const uint32_t q_addr = smem_u32(smem + OFF_Q);
            const uint32_t k_addr = smem_u32(smem + OFF_K);
            mbar_wait(k_full, 0);
            int t = 0, sb = 0;
            for (int c = 0; c < total; ++c) {  // runs as far ahead as free slots allow
                const uint32_t slot = c % NSLOT;
                if (sb == 0) mbar_wait(&q_full[t & 1], (t >> 1) & 1);
                if (c >= NSLOT) mbar_wait(&s_free[(c - NSLOT) % NRING], ((c - NSLOT) / NRING) & 1);  // the slot's previous S
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    umma_f16_ss(tmem_base + TM_S + slot * SB, make_sw128_desc(q_addr + (t & 1) * TILE_BYTES + k * 32),
                                make_sw128_desc(k_addr + sb * (SB * 128) + k * 32), idesc_s, k != 0);
                }
                umma_commit(&s_full[c % NRING]);
                if (++sb == nsb) {
                    sb = 0;
                    ++t;
                }
            }
        }
      } else if (warp == 3) {
        if (elect_one()) {
            // ------------------------------------------------------------ P.V issuer: O_t (+)= P_c . V_c
            constexpr uint32_t idesc_o = make_idesc_f16(QT, D, 0, 1);  // 128 x 64, B (= V) MN-major
            const uint32_t v_addr = smem_u32(smem + OFF_V);
            const uint32_t p_addr = smem_u32(smem + OFF_P);
            if (total > 0) mbar_wait(v_full, 0);
            int t = 0, sb = 0;
            uint32_t pb = c_off, phases = 0;  // bit w of phases: parity of the phase of p_full[w] awaited next
            for (int c = 0; c < total; ++c) {
                ATT_STAMP(4, c, 0);
                mbar_wait(&p_full[pb], (phases >> pb) & 1);
                phases ^= 1u << pb;
                if (sb == 0 && t >= 2) mbar_wait(&o_free[t & 1], ((t >> 1) - 1) & 1);  // tile t-2 has been written out
                tc_fence_after();
                ATT_STAMP(4, c, 1);
#pragma unroll
                for (int kk = 0; kk < SB / 16; ++kk) {
                    const uint32_t a = p_addr + pb * TILE_BYTES + kk * 32;
                    const uint32_t bv = v_addr + (sb * SB + kk * 16) * 128;  // key row -> 128 B
                    umma_f16_ss(tmem_base + TM_O + (t & 1) * D, make_sw128_desc(a), make_sw128_desc(bv), idesc_o,
                                (sb | kk) != 0);
                }
                umma_commit(&pv_done[pb]);  // P buffer pb is free again; O_t is complete up to sub-block c
                ATT_STAMP(4, c, 2);
                if (++sb == nsb) {
                    umma_commit(&o_done[t & 1]);
                    sb = 0;
                    ++t;
                }
                if (++pb == NEXP) pb = 0;
            }
        }
      }
    } else if (warp < 4 + 4 * NEXP) {
        // ---------------------------------------------------------------- exp warpgroup w: sub-blocks c = w (mod NEXP)
        if constexpr (NEXP == 3) setmaxnreg_inc<120>();
        const int w = (warp - 4) >> 2;
        const int wp = (w + NEXP - 1) % NEXP;  // the warpgroup that handles c - 1
        const int r = (warp & 3) * 32 + lane;  // query row within the tile == TMEM lane
        const uint32_t tm = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t p_row = smem_u32(smem + OFF_P + w * TILE_BYTES) + r * 128;
        const uint32_t mr_self = smem_u32(smem + OFF_MR) + (w * QT + r) * 4;
        const uint32_t mr_prev = smem_u32(smem + OFF_MR) + (wp * QT + r) * 4;
        const uint32_t ls_self = smem_u32(smem + OFF_LS) + (w * QT + r) * 8;
        const uint32_t swz = static_cast<uint32_t>(r & 7);
        const bool obs = (warp & 3) == 0 && lane == 0;
        int t = 0, sb = (w + NEXP - c_off) % NEXP;
        while (sb >= nsb) {
            sb -= nsb;
            ++t;
        }
        float l_w = 0.f, m_ref = 0.f;  // this warpgroup's partial row sum of the tile, relative to m_ref
        uint32_t use = 0;              // how often this warpgroup's P buffer has been filled
        const int c_first_w = (w + NEXP - c_off) % NEXP;    // this warpgroup's first sub-block
        const int c_first_wp = (wp + NEXP - c_off) % NEXP;  // ... and its predecessor's
        if (USE_TOKEN && w == (c_off + NEXP - 1) % NEXP) named_bar_arrive(BAR_TOKEN + c_off, 256);  // warpgroup c_off goes first
#pragma unroll 1
        for (int c = c_first_w; c < total; c += NEXP, ++use) {
            const uint32_t slot = c % NSLOT;
            if (obs) ATT_STAMP(w, c, 0);
            mbar_wait(&s_full[c % NRING], (c / NRING) & 1);
            tc_fence_after();
            if (obs) ATT_STAMP(w, c, 1);
            const int valid = len - sb * SB;  // keys [0, valid) of this sub-block are real (>= 1)
            uint32_t v[64];
            auto load_scores = [&]() {
                tmem_ld_32x32b_x32(tm + TM_S + slot * SB, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                tmem_ld_32x32b_x32(tm + TM_S + slot * SB + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                tmem_ld_wait();
                if (valid < SB) {
#pragma unroll
                    for (int e = 0; e < SB; ++e)
                        if (e >= valid) v[e] = __float_as_uint(-INFINITY);
                }
            };
            load_scores();
            float mx = -INFINITY;
#pragma unroll
            for (int e = 0; e < SB; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
            // running reference max of the tile's accumulator: take over the previous sub-block's unless this one
            // exceeds it by more than 2^8 in the exp2 domain
            float m_used = mx, m_prev = mx;
            if (c > 0) named_bar_sync(BAR_MAX + wp, 256);  // (c-1)'s reference max is in shared memory
            if (sb != 0) {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m_prev) : "r"(mr_prev) : "memory");
                m_used = (mx - m_prev > kRescaleThreshold) ? mx : m_prev;
            }
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(mr_self), "f"(m_used) : "memory");
            named_bar_arrive(BAR_MAX + w, 256);
            if (sb < NEXP) {  // this warpgroup's first sub-block of the tile
                l_w = 0.f;
            } else {
                l_w *= ex2_approx((m_ref - m_used) * kScaleLog2e);
            }
            m_ref = m_used;
            if (obs) ATT_STAMP(w, c, 2);
            if (sb != 0 && __any_sync(0xffffffffu, m_used != m_prev)) {
                // rare: rescale this warp's 32 rows of the accumulator by 2^((m_prev - m_used) k) (1 where unchanged).
                // The scores are dropped and read again afterwards so that this path costs the common one no registers.
                // Our own previous P.V first (its barrier's phases are met in order), then sub-block c-1's: its
                // predecessor on that barrier was issued before ours and has therefore retired too.
                if (use > 0) mbar_wait(&pv_done[w], (use - 1) & 1);
                mbar_wait(&pv_done[wp], ((c - 1 - c_first_wp) / NEXP) & 1);  // O_t is complete up to sub-block c-1
                tc_fence_after();
                const float f = ex2_approx((m_prev - m_used) * kScaleLog2e);
                const uint32_t o_addr = tm + TM_O + (t & 1) * D;
#pragma unroll
                for (int part = 0; part < D / 32; ++part) {
                    tmem_ld_32x32b_x32(o_addr + part * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * f);
                    tmem_st_32x32b_x32(o_addr + part * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                }
                tmem_st_wait();
                load_scores();
            }
            tc_fence_before();
            mbar_arrive(&s_free[c % NRING]);  // the scores live in registers from here on
            const float neg_ms = -m_used * kScaleLog2e;
            if (use > 0) mbar_wait(&pv_done[w], (use - 1) & 1);  // the previous P of this buffer has been consumed
            if (USE_TOKEN) named_bar_sync(BAR_TOKEN + w, 256);
            if (obs) ATT_STAMP(w, c, 3);
            float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x0 = fmaf(__uint_as_float(v[q * 8 + 2 * e]), kScaleLog2e, neg_ms);
                    const float x1 = fmaf(__uint_as_float(v[q * 8 + 2 * e + 1]), kScaleLog2e, neg_ms);
                    const float p0 = ex2_approx(x0);
                    const float p1 = ex2_approx(x1);
                    ls0 += p0;
                    ls1 += p1;
                    pk[e] = pack_half2(p0, p1);
                }
                // keys 8q .. 8q+7 of row r -> 16-byte chunk q ^ (r & 7) of the row's 128 bytes
                sts128(p_row + ((static_cast<uint32_t>(q) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
            if (USE_TOKEN) named_bar_arrive(BAR_TOKEN + (w + 1 == NEXP ? 0 : w + 1), 256);  // MUFU to the next warpgroup
            l_w += ls0 + ls1;
            if (sb + NEXP >= nsb)  // this warpgroup's last sub-block of the tile
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ls_self + t * (MAX_NEXP * QT * 8)), "f"(m_ref), "f"(l_w) : "memory");
            tc_fence_before();         // our TMEM accesses precede the MMA that accumulates into O_t
            fence_proxy_async_smem();  // P_c visible to the tensor core's async-proxy reads
            mbar_arrive(&p_full[w]);   // (release: also publishes (m, l) to the epilogue via the o_done chain)
            if (obs) ATT_STAMP(w, c, 4);
            sb += NEXP;
            while (sb >= nsb) {
                sb -= nsb;
                ++t;
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue warpgroup: ctx = O_t / l
        if constexpr (NEXP == 3) setmaxnreg_dec<80>();
        const int r = (warp & 3) * 32 + lane;
        const uint32_t tm = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t ls_base = smem_u32(smem + OFF_LS) + r * 8;
        const uint32_t swz = static_cast<uint32_t>(r & 7);
        const bool obs = (warp & 3) == 0 && lane == 0;
#pragma unroll 1
        for (int t = 0; t < nq; ++t) {
            if (obs) ATT_STAMP(3, t, 0);
            mbar_wait(&o_done[t & 1], (t >> 1) & 1);
            tc_fence_after();
            if (obs) ATT_STAMP(3, t, 1);
            // l = sum of the warpgroups' partial row sums brought to the tile's final reference max (the last sub-block's)
            const int c_first = (t0 + t) * nsb;  // (item-wide index: decides which warpgroup had which sub-block)
            const int w_last = (c_first + nsb - 1) % NEXP;
            float m_fin, l_fin;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m_fin), "=f"(l_fin) : "r"(ls_base + (t * MAX_NEXP + w_last) * (QT * 8)) : "memory");
#pragma unroll
            for (int w = 0; w < NEXP; ++w) {
                const int first_sb = (w - c_first % NEXP + NEXP) % NEXP;  // warpgroup w's first sub-block in this tile
                if (w != w_last && first_sb < nsb) {
                    float m_w, l_w;
                    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m_w), "=f"(l_w) : "r"(ls_base + (t * MAX_NEXP + w) * (QT * 8)) : "memory");
                    l_fin = fmaf(l_w, ex2_approx((m_w - m_fin) * kScaleLog2e), l_fin);
                }
            }
            const float inv_l = 1.0f / l_fin;
            // staging tile = this tile's Q buffer: all of the tile's MMAs have retired (o_done), and tile t+2's Q is only
            // loaded into it once our TMA store has read it back out (q_empty below)
            const uint32_t o_row = smem_u32(smem + OFF_Q + (t & 1) * TILE_BYTES) + r * 128;
#pragma unroll
            for (int part = 0; part < D / 32; ++part) {
                uint32_t o[32];
                tmem_ld_32x32b_x32(tm + TM_O + (t & 1) * D + part * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // output dims 8j .. 8j+7 of row r -> 16-byte chunk j ^ (r & 7) of the row's 128 bytes
                    const uint32_t j = part * 4 + i;
                    sts128(o_row + ((j ^ swz) << 4),
                           pack_half2(__uint_as_float(o[8 * i]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l),
                           pack_half2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l),
                           pack_half2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l),
                           pack_half2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l));
                }
            }
            tc_fence_before();
            mbar_arrive(&o_free[t & 1]);
            fence_proxy_async_smem();
            named_bar_sync(BAR_EPI, 128);
            if (warp == 4 + 4 * NEXP && lane == 0) {  // rows past S are clipped by the tensor map
                tma_store_3d(&tctx, smem + OFF_Q + (t & 1) * TILE_BYTES, h * D, (t0 + t) * QT, b);
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(&q_empty[t & 1]);
            }
            if (obs) ATT_STAMP(3, t, 2);
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

namespace {
// diagnostics: B200RT_ATTN_VARIANT = "<2|3><t|n>": number of exp warpgroups, and whether their exp phases take turns on the
// MUFU (t) or overlap freely (n).  The product default is the fastest measured combination.
constexpr int kDefaultVariant = 3;  // 3n
int attention_variant() {
    static const int v = [] {
        const char* e = getenv("B200RT_ATTN_VARIANT");
        if (!e || (e[0] != '2' && e[0] != '3') || (e[1] != 't' && e[1] != 'n')) return kDefaultVariant;
        return (e[0] == '3' ? 2 : 0) + (e[1] == 'n' ? 1 : 0);
    }();
    return v;
}
template <int NEXP, bool TOKEN>
cudaError_t set_smem() {
    return cudaFuncSetAttribute(attn::attention_kernel<NEXP, TOKEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::smem_bytes(NEXP));
}
template <int NEXP, bool TOKEN>
void launch(const CUtensorMap& tq, const CUtensorMap& tctx, const int32_t* lens, int B, int S, cudaStream_t stream,
            unsigned long long* dbg) {
    // fewer (item, head) units than SMs: split them by query tile (diagnostic stamps keep the unsplit layout)
    const int nq = (S + attn::QT - 1) / attn::QT;
    static const bool no_split = getenv("B200RT_ATTN_NOSPLIT") != nullptr;  // diagnostics
    const int split = (B * HEADS < 148 && nq > 1 && dbg == nullptr && !no_split) ? 1 : 0;
    attn::attention_kernel<NEXP, TOKEN><<<B * HEADS * (split ? nq : 1), attn::num_threads(NEXP), attn::smem_bytes(NEXP), stream>>>(
        tq, tctx, lens, S, split, dbg);
}
}  // namespace

cudaError_t attention_init_device() {
    cudaError_t e = set_smem<2, true>();
    if (e == cudaSuccess) e = set_smem<2, false>();
    if (e == cudaSuccess) e = set_smem<3, true>();
    if (e == cudaSuccess) e = set_smem<3, false>();
    return e;
}

cudaError_t launch_attention(const CUtensorMap& tq, const CUtensorMap& tctx, const int32_t* lens, int B, int S,
                             cudaStream_t stream, unsigned long long* dbg) {
    if (S < 1 || S > attn::MAX_KB * attn::KB || B < 1) return cudaErrorInvalidValue;
    switch (attention_variant()) {
        case 1: launch<2, false>(tq, tctx, lens, B, S, stream, dbg); break;
        case 2: launch<3, true>(tq, tctx, lens, B, S, stream, dbg); break;
        case 3: launch<3, false>(tq, tctx, lens, B, S, stream, dbg); break;
        default: launch<2, true>(tq, tctx, lens, B, S, stream, dbg); break;
    }
    return cudaGetLastError();
}

}  // namespace b200
