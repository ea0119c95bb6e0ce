// Multi-head self-attention for S <= 512, d = 64, on tcgen05 (sm_100a).
//
// One CTA per (item, head) keeps that head's K and V (<= 512 x 64 fp16 each) resident in shared memory and
// streams the item's 128-row query tiles through them as a sequence of 64-key sub-blocks c = 0, 1, 2, ...
// (8 per 512-key tile, continuing across tiles).  What bounds the kernel is the exp: one warp's MUFU stream sustains an
// EX2 per ~11 cycles and its 64-score exp phase takes ~660 cycles alone, 1100+ when the same SM sub-partition also
// issues an accumulator fold (tools/ubench/spin_cost.cu).  So everything that is not the exp is kept off the exp warps:
//
//   TMEM         : one 64-column O accumulator per (tile parity, exp warpgroup) + a ring of four 64-column S slots
//   S thread     : S_c = Q . K_c^T (128x64) into S slot c & 3, as far ahead as free slots allow
//   P.V thread   : O_{t,w} += P_c . V_c accumulates IN TMEM across warpgroup w's sub-blocks of tile t (no per-sub-block
//                  fold; a second issuing thread because one thread doing both chains was itself the bottleneck at
//                  ~1270 cycles per sub-block)
//   2 exp WGs    : warpgroup w takes sub-blocks c = w (mod 2), thread = query row: one TMEM read of the 64 scores,
//                  row max, P_c = exp2((S - m) k) as fp16 into P buffer c & 1 (128B-swizzled), row sum.  m is the
//                  REFERENCE maximum of the warpgroup's own accumulator: it only moves when a sub-block's maximum exceeds
//                  it by more than 2^8 in the exp2 domain (P <= 256 is exact enough in fp16 and the sums are fp32), so the
//                  accumulator in TMEM is rescaled (tcgen05.ld/st by the exp warp that saw the jump) a handful of
//                  times per tile instead of once per sub-block, and the warpgroups never exchange anything.
//                  Their exp phases take turns on the MUFU (a named-barrier token).
//   epilogue WG  : once per tile: brings the warpgroups' (m, l, O) to a common reference, ctx = O / l as fp16 through a
//                  swizzled staging tile and one TMA store -- row-per-thread global stores cost 32 LSU wavefronts each
//                  and stalled the exp warps' shared-memory stores behind them.
//
// Keys >= len are masked to -inf before the max (exactly P = 0, matching HF's additive -inf mask); sub-blocks wholly
// past len are skipped.  k = log2(e) / sqrt(64).  Batches with fewer (item, head) units than SMs are split by query tile.
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
constexpr int NEXP = 2;      // exp warpgroups == P buffers == accumulators per tile
constexpr int NSLOT = 4;     // TMEM S slots of 64 columns == the ring of their barriers
constexpr uint32_t TM_O = 0;             // accumulators: (tile parity, warpgroup) -> 2 x NEXP x 64 columns
constexpr uint32_t TM_S = 2 * NEXP * D;  // S ring: NSLOT x 64 columns
constexpr int TILE_BYTES = 128 * D * 2;  // 16 KB: 128 rows x 128 B
constexpr int OFF_Q = 0;                            // 2 x 16 KB (double buffered across tiles)
constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;       // 4 x 16 KB
constexpr int OFF_V = OFF_K + MAX_KB * TILE_BYTES;  // 4 x 16 KB
constexpr int OFF_P = OFF_V + MAX_KB * TILE_BYTES;  // 2 x 16 KB: P_c goes to buffer c & 1 whichever warpgroup produced it
constexpr int OFF_O = OFF_P + 2 * TILE_BYTES;       // 16 KB: the epilogue's staging tile for the TMA store
constexpr int OFF_LS = OFF_O + TILE_BYTES;          // float2 [MAX_NQ][NEXP][128]: (reference max, row sum) per tile and warpgroup
constexpr int OFF_BAR = OFF_LS + MAX_NQ * NEXP * QT * 8;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
static_assert(OFF_P % 1024 == 0 && OFF_O % 1024 == 0 && OFF_V % 1024 == 0, "swizzled tiles must be 1024-byte aligned");
static_assert(SMEM_BYTES <= 232448, "shared memory");
constexpr int NUM_THREADS = 128 + NEXP * 128 + 128;
constexpr int BAR_TOKEN = 1;               // named barriers 1 .. NEXP: MUFU token
constexpr int BAR_EPI = 1 + NEXP;      // the epilogue warpgroup's own barrier

// softmax_scale * log2(e) with softmax_scale = 1/sqrt(64)
constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;
// the reference max follows the true max only when it is exceeded by more than this (raw score units): P <= 2^8
constexpr float kRescaleThreshold = 8.0f / kScaleLog2e;

// USE_TOKEN: the exp phases of the two warpgroups take turns on the MUFU (product) or overlap freely (diagnostics).
template <bool USE_TOKEN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
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
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* k_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* q_full = bars + 2;       // [2]
    uint64_t* q_empty = bars + 4;      // [2]
    uint64_t* o_done = bars + 6;       // [2]  the tile's last P.V has retired
    uint64_t* o_free = bars + 8;       // [2]  the epilogue has read the accumulators
    // Rings of NRING barriers indexed by c % NRING.  Every barrier has ONE waiter, which meets its phases in order: a
    // waiter that skipped a phase would find "the phase of parity p complete" true at once (seen with 2 slots and 3
    // warpgroups, where a warpgroup's first wait was for phase 1 of a barrier whose phase 0 belonged to another one).
    constexpr int NRING = NSLOT;
    uint64_t* s_full = bars + 10;      // [NRING]  S_c landed
    uint64_t* s_free = bars + 16;      // [NRING]  S_c is in the exp warpgroup's registers
    uint64_t* p_full = bars + 22;      // [NRING]  P_c written (buffer c & 1)
    uint64_t* pv_done = bars + 28;     // [NRING]  P.V of sub-block c has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 34);

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
        for (int i = 0; i < NRING; ++i) {
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

    if (warp < 4) {
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
                mbar_wait(&q_empty[qb], ((t >> 1) - 1) & 1);  // tile t-2's S MMAs have retired
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
                const uint32_t slot = c & (NSLOT - 1);
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
                    umma_commit(&q_empty[t & 1]);  // this tile's Q is no longer needed once these retire
                    sb = 0;
                    ++t;
                }
            }
        }
      } else if (warp == 3) {
        if (elect_one()) {
            // ------------------------------------------------------------ P.V issuer: O_{t,w} (+)= P_c . V_c, w = c mod NEXP
            constexpr uint32_t idesc_o = make_idesc_f16(QT, D, 0, 1);  // 128 x 64, B (= V) MN-major
            const uint32_t v_addr = smem_u32(smem + OFF_V);
            const uint32_t p_addr = smem_u32(smem + OFF_P);
            if (total > 0) mbar_wait(v_full, 0);
            int t = 0, sb = 0, w = 0;
            for (int c = 0; c < total; ++c) {
                ATT_STAMP(4, c, 0);
                mbar_wait(&p_full[c % NRING], (c / NRING) & 1);
                if (sb == 0 && t >= 2) mbar_wait(&o_free[t & 1], ((t >> 1) - 1) & 1);  // tile t-2 has been written out
                tc_fence_after();
                ATT_STAMP(4, c, 1);
#pragma unroll
                for (int kk = 0; kk < SB / 16; ++kk) {
                    const uint32_t a = p_addr + (c & 1) * TILE_BYTES + kk * 32;
                    const uint32_t bv = v_addr + (sb * SB + kk * 16) * 128;  // key row -> 128 B
                    // warpgroup w's accumulator of tile t; its first sub-block of the tile (sb < NEXP) overwrites
                    umma_f16_ss(tmem_base + TM_O + ((t & 1) * NEXP + w) * D, make_sw128_desc(a), make_sw128_desc(bv),
                                idesc_o, (sb >= NEXP) || kk != 0);
                }
                umma_commit(&pv_done[c % NRING]);  // P buffer c & 1 is free again; O_{t,w} is stable until w's next P.V
                ATT_STAMP(4, c, 2);
                if (++sb == nsb) {
                    umma_commit(&o_done[t & 1]);
                    sb = 0;
                    ++t;
                }
                if (++w == NEXP) w = 0;
            }
        }
      }
    } else if (warp < 4 + 4 * NEXP) {
        // ---------------------------------------------------------------- exp warpgroup w: sub-blocks c = w (mod NEXP)
        const int w = (warp - 4) >> 2;
        const int r = (warp & 3) * 32 + lane;  // query row within the tile == TMEM lane
        const uint32_t tm = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t p_base = smem_u32(smem + OFF_P) + r * 128;
        const uint32_t ls_self = smem_u32(smem + OFF_LS) + (w * QT + r) * 8;
        const uint32_t swz = static_cast<uint32_t>(r & 7);
        const bool obs = (warp & 3) == 0 && lane == 0;
        int t = 0, sb = w;
        while (sb >= nsb) {
            sb -= nsb;
            ++t;
        }
        float l_w = 0.f, m_ref = 0.f;  // this warpgroup's row sum of the tile, and the reference max it is relative to
        if (USE_TOKEN && w == NEXP - 1) named_bar_arrive(BAR_TOKEN, 256);  // warpgroup 0 goes first
#pragma unroll 1
        for (int c = w; c < total; c += NEXP) {
            const uint32_t slot = c & (NSLOT - 1);
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
            // reference max of this warpgroup's accumulator: moves only when exceeded by more than 2^8 in the exp2 domain
            const bool first = sb < NEXP;  // this warpgroup's first sub-block of the tile
            const float m_prev = m_ref;
            if (first || mx - m_ref > kRescaleThreshold) m_ref = mx;
            if (first) l_w = 0.f;
            if (obs) ATT_STAMP(w, c, 2);
            if (!first && __any_sync(0xffffffffu, m_ref != m_prev)) {
                // rare: rescale this warp's 32 rows of the accumulator by 2^((m_prev - m_ref) k) (1 where unchanged).
                // The scores are dropped and read again afterwards so that this path costs the common one no registers.
                // sub-block c-2's P.V has retired, hence (commit order) our own previous one, c-NEXP: the accumulator is stable
                if (c >= 2) mbar_wait(&pv_done[(c - 2) % NRING], ((c - 2) / NRING) & 1);
                tc_fence_after();
                const float f = ex2_approx((m_prev - m_ref) * kScaleLog2e);
                l_w *= f;
                const uint32_t o_addr = tm + TM_O + ((t & 1) * NEXP + w) * D;
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
            const float neg_ms = -m_ref * kScaleLog2e;
            // P buffer c & 1 last held P_{c-2}: wait for that P.V (each barrier of the ring has this one waiter)
            if (c >= 2) mbar_wait(&pv_done[(c - 2) % NRING], ((c - 2) / NRING) & 1);
            const uint32_t p_row = p_base + (c & 1) * TILE_BYTES;
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
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ls_self + t * (NEXP * QT * 8)), "f"(m_ref), "f"(l_w) : "memory");
            tc_fence_before();         // our TMEM accesses precede the MMA that accumulates into O_t
            fence_proxy_async_smem();  // P_c visible to the tensor core's async-proxy reads
            mbar_arrive(&p_full[c % NRING]);  // (release: also publishes (m, l) to the epilogue via the o_done chain)
            if (obs) ATT_STAMP(w, c, 4);
            sb += NEXP;
            while (sb >= nsb) {
                sb -= nsb;
                ++t;
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue warpgroup: ctx = O_t / l
        const int r = (warp & 3) * 32 + lane;
        const uint32_t tm = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t ls_base = smem_u32(smem + OFF_LS) + r * 8;
        const uint32_t o_row = smem_u32(smem + OFF_O) + r * 128;
        const uint32_t swz = static_cast<uint32_t>(r & 7);
        const bool obs = (warp & 3) == 0 && lane == 0;
#pragma unroll 1
        for (int t = 0; t < nq; ++t) {
            if (obs) ATT_STAMP(3, t, 0);
            mbar_wait(&o_done[t & 1], (t >> 1) & 1);
            tc_fence_after();
            if (obs) ATT_STAMP(3, t, 1);
            // combine the warpgroups' accumulators: O = sum_w 2^((m_w - m) k) O_w, l likewise, m = max_w m_w
            const int c_first = t * nsb;
            float m_w[NEXP], l_w[NEXP], f_w[NEXP];
            bool part_w[NEXP];
            float m_fin = -INFINITY;
#pragma unroll
            for (int w = 0; w < NEXP; ++w) {
                part_w[w] = (w - c_first % NEXP + NEXP) % NEXP < nsb;  // warpgroup w had a sub-block in this tile
                m_w[w] = -INFINITY;
                l_w[w] = 0.f;
                if (part_w[w]) {
                    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m_w[w]), "=f"(l_w[w]) : "r"(ls_base + (t * NEXP + w) * (QT * 8)) : "memory");
                    m_fin = fmaxf(m_fin, m_w[w]);
                }
            }
            float l_fin = 0.f;
#pragma unroll
            for (int w = 0; w < NEXP; ++w) {
                f_w[w] = part_w[w] ? ex2_approx((m_w[w] - m_fin) * kScaleLog2e) : 0.f;
                l_fin = fmaf(l_w[w], f_w[w], l_fin);
            }
            const float inv_l = 1.0f / l_fin;
#pragma unroll
            for (int w = 0; w < NEXP; ++w) f_w[w] *= inv_l;
            // the previous tile's TMA store has finished reading the staging tile
            if (warp == 4 + 4 * NEXP && lane == 0) tma_store_wait_read<0>();
            named_bar_sync(BAR_EPI, 128);
#pragma unroll
            for (int part = 0; part < D / 32; ++part) {
                float acc[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) acc[e] = 0.f;
#pragma unroll
                for (int w = 0; w < NEXP; ++w) {
                    if (part_w[w]) {  // (uniform)
                        uint32_t o[32];
                        tmem_ld_32x32b_x32(tm + TM_O + ((t & 1) * NEXP + w) * D + part * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 32; ++e) acc[e] = fmaf(f_w[w], __uint_as_float(o[e]), acc[e]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // output dims 8j .. 8j+7 of row r -> 16-byte chunk j ^ (r & 7) of the row's 128 bytes
                    const uint32_t j = part * 4 + i;
                    sts128(o_row + ((j ^ swz) << 4), pack_half2(acc[8 * i], acc[8 * i + 1]),
                           pack_half2(acc[8 * i + 2], acc[8 * i + 3]), pack_half2(acc[8 * i + 4], acc[8 * i + 5]),
                           pack_half2(acc[8 * i + 6], acc[8 * i + 7]));
                }
            }
            tc_fence_before();
            mbar_arrive(&o_free[t & 1]);
            fence_proxy_async_smem();
            named_bar_sync(BAR_EPI, 128);
            if (warp == 4 + 4 * NEXP && lane == 0) {  // rows past S are clipped by the tensor map
                tma_store_3d(&tctx, smem + OFF_O, h * D, (t0 + t) * QT, b);
                tma_store_commit();
            }
            if (obs) ATT_STAMP(3, t, 2);
        }
        if (warp == 4 + 4 * NEXP && lane == 0) tma_store_wait_read<0>();  // shared memory must outlive the last store's reads
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
// diagnostics: B200RT_ATTN_VARIANT=n lets the exp phases of the two warpgroups overlap instead of taking turns on the MUFU
bool attention_free_running() {
    static const bool v = [] {
        const char* e = getenv("B200RT_ATTN_VARIANT");
        return e && e[0] == 'n';
    }();
    return v;
}
}  // namespace

cudaError_t attention_init_device() {
    cudaError_t e = cudaFuncSetAttribute(attn::attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn::attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES);
    return e;
}

cudaError_t launch_attention(const CUtensorMap& tq, const CUtensorMap& tctx, const int32_t* lens, int B, int S,
                             cudaStream_t stream, unsigned long long* dbg) {
    if (S < 1 || S > attn::MAX_KB * attn::KB || B < 1) return cudaErrorInvalidValue;
    // fewer (item, head) units than SMs: split them by query tile (diagnostic stamps keep the unsplit layout)
    const int nq = (S + attn::QT - 1) / attn::QT;
    const int split = (B * HEADS < 148 && nq > 1 && dbg == nullptr) ? 1 : 0;
    const int grid = B * HEADS * (split ? nq : 1);
    if (attention_free_running())
        attn::attention_kernel<false><<<grid, attn::NUM_THREADS, attn::SMEM_BYTES, stream>>>(tq, tctx, lens, S, split, dbg);
    else
        attn::attention_kernel<true><<<grid, attn::NUM_THREADS, attn::SMEM_BYTES, stream>>>(tq, tctx, lens, S, split, dbg);
    return cudaGetLastError();
}

}  // namespace b200
