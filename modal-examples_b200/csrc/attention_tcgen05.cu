// Multi-head self-attention for S <= 512, d = 64, on tcgen05 (sm_100a): persistent CTAs.
//
// A work unit is one (item, head) -- or, for batches with fewer units than SMs, one (item, head, query tile).  One CTA per
// SM walks the units u = blockIdx.x, blockIdx.x + gridDim.x, ...  For the unit in hand it keeps that head's K and V
// (<= 512 x 64 fp16 each, four 128-key tiles) resident in shared memory and streams the item's 128-row query tiles through
// them as a sequence of 64-key sub-blocks c = 0, 1, 2, ... (8 per 512-key tile, continuing across tiles).
//
// Round 1 launched one CTA per unit and paid ~7 k cycles of prologue (tensor-map fetch, barrier init, TMEM alloc, 80 KB of
// Q + K from L2/HBM) on every one of them, ~20 % of a unit (profiles/attn_timeline_r01.txt).  Here the prologue is paid once
// per SM, and the NEXT unit's operands are prefetched under the current unit's tail: K tile j of the next unit is loaded as
// soon as the current unit's LAST query tile has issued its S for the sub-blocks of tile j (k_free[j], a tcgen05.commit),
// V tile j once the last tile's P.V has consumed it (v_free[j]), and query tiles flow through their two buffers as one
// sequence across units.  All barrier phases are kept in per-role running counters, since they no longer restart per CTA.
//
// What bounds the kernel is the exp: one warp's MUFU stream sustains an EX2 per ~11 cycles and its 64-score exp phase takes
// ~660 cycles alone, 1100+ when the same SM sub-partition also issues an accumulator fold (tools/ubench/spin_cost.cu).  So
// everything that is not the exp is kept off the exp warps:
//
//   TMEM         : two 64-column O accumulators (tile parity) + NSLOT = 6 64-column S slots, two per exp warpgroup.  P_c never
//                  touches shared memory: the exp warps write it (fp16, 32 columns) over the S_c they have just read and
//                  the P.V MMA takes its A operand from tensor memory.  (Through a swizzled smem tile -- 8 STS.128 per row, a
//                  proxy fence, 16 KB written and read back per sub-block -- the launch was 2.2 % slower at full length and
//                  4.2 % on ragged batches.)
//   TMA thread   : Q tiles, K / V tiles of this and the next unit (see above)
//   S thread     : S_c = Q . K_c^T (128x64) into one of the two S slots of the warpgroup that will take sub-block c, as far
//                  ahead as free slots allow -- across unit boundaries too.  A slot is free again once the P.V that read the
//                  P in it has RETIRED (the tracker's counter), not when its S has been read.
//   P.V thread   : O_t += P_c . V_c (A = P_c in TMEM, B = V_c in smem) accumulates IN TMEM across the tile's sub-blocks
//   3 exp WGs    : warpgroup w takes the sub-blocks of a unit with (c_off + c) % 3 == w, thread = query row: one TMEM read of
//                  the 64 scores, row max, P_c = exp2((S - m) k) as fp16 back into the slot (tcgen05.st), partial row
//                  sum.  m is the row's RUNNING reference maximum, handed from sub-block to sub-block through shared memory;
//                  it only moves when the new maximum exceeds it by more than 2^8 in the exp2 domain (P <= 256 is exact
//                  enough in fp16 and the sums are fp32), so the accumulator in TMEM is rescaled (tcgen05.ld/st by the exp
//                  warp that saw the jump) a handful of times per tile instead of once per sub-block.
//   epilogue WG  : once per tile: l = sum of the warpgroups' partial sums brought to the final m, ctx = O / l as fp16
//                  through a swizzled staging tile (the tile's own, now idle, Q buffer) and one TMA store.
//
// Keys >= len are masked to -inf before the max (exactly P = 0, matching HF's additive -inf mask); sub-blocks wholly past
// len are skipped.  k = log2(e) / sqrt(64).
//
// Which warpgroup gets which sub-block depends only on the item: sub-block c of tile t goes to warpgroup (t nsb + c) mod 3
// whether the item travels as one unit or split by query tile, alone or in any batch, on any CTA -- the row sums are grouped
// by warpgroup, so this keeps an item's embedding bit-identical in any batch (tests: scheduler / size-independent properties).
//
// Restates BertSelfAttention.forward (HF modeling_bert.py:143-207) for the TEI /embed path the
// reference calls at 06_gpu_and_ml/embeddings/text_embeddings_inference.py:100.
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
constexpr int NEXP = 3;      // exp warpgroups
constexpr int NSLOT = 6;     // TMEM S slots of 64 columns, two per exp warpgroup; P_c (fp16) overwrites the first 32 columns of S_c's
constexpr uint32_t TM_O = 0;        // accumulators: tile parity -> 2 x 64 columns
constexpr uint32_t TM_S = 2 * D;    // S ring: NSLOT x 64 columns
constexpr int TILE_BYTES = 128 * D * 2;  // 16 KB: 128 rows x 128 B
constexpr int OFF_Q = 0;                            // 2 x 16 KB (double buffered across tiles); doubles as the epilogue's
                                                    // staging tile once the tile's MMAs have retired
constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;       // 4 x 16 KB
constexpr int OFF_V = OFF_K + MAX_KB * TILE_BYTES;  // 4 x 16 KB
constexpr int OFF_MR = OFF_V + MAX_KB * TILE_BYTES; // float [NEXP][128]: reference max after warpgroup w's latest sub-block
constexpr int OFF_LS = OFF_MR + NEXP * QT * 4;      // float2 [4 tiles in flight][NEXP][128]: (reference max, partial row sum)
constexpr int OFF_BAR = OFF_LS + MAX_NQ * NEXP * QT * 8;
static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0, "swizzled tiles must be 1024-byte aligned");
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
constexpr int NUM_THREADS = 128 + NEXP * 128 + 128;  // 640
static_assert(SMEM_BYTES <= 232448, "shared memory");
// named barriers 1 .. 9: running-max hand-off, one per (producing warpgroup, consuming warpgroup) pair.  Inside a unit the
// consumer of warpgroup w's sub-block is always warpgroup w+1, but across a unit boundary it is whoever owns the next unit's
// first sub-block: with one barrier per producer, a warpgroup running ahead into the next unit could pair its sync with an
// arrival meant for another consumer.  Per pair, arrivals and syncs both follow the CTA's sub-block order.
constexpr int BAR_MAX = 1;
constexpr int BAR_EPI = 1 + NEXP * NEXP;  // the epilogue warpgroup's own barrier
static_assert(BAR_EPI < 16, "named barrier ids");

// softmax_scale * log2(e) with softmax_scale = 1/sqrt(64)
constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;
// the reference max follows the true max only when it is exceeded by more than this (raw score units): P <= 2^8
constexpr float kRescaleThreshold = 8.0f / kScaleLog2e;

#ifdef B200RT_DIAG
// progress dump: every role notes (unit, sub-block, step) in shared memory; a timed-out mbarrier wait prints them all
#define ATT_PROG(role, u_, c_, step_)                                                                    \
    do {                                                                                                  \
        volatile uint32_t* pw = reinterpret_cast<volatile uint32_t*>(prog) + (role) * 4;                  \
        pw[0] = (u_); pw[1] = (c_); pw[2] = (step_);                                                      \
    } while (0)
// dbg (tools/attn_timeline.py): CTA 0 records clock64() stamps; observer o in {exp WG 0..2, epilogue WG, P.V thread},
// 32 sub-blocks x 8 slots each
#define ATT_STAMP(o, c, slot)                                                                          \
    do {                                                                                               \
        if (dbg != nullptr && blockIdx.x == 0 && (c) < 32) dbg[((o) * 32 + (c)) * 8 + (slot)] = clock64(); \
    } while (0)
#else
#define ATT_STAMP(o, c, slot) do { } while (0)
#define ATT_PROG(role, u_, c_, step_) do { } while (0)
#endif

struct Unit {
    int b, h, t0, nq;  // item, head, first query tile, query tiles
    int len, nsb, nkb; // valid keys, 64-key sub-blocks, 128-key K/V tiles
    int total;         // sub-blocks of the unit = nq * nsb
    int c_off;         // sub-block c of the unit goes to warpgroup (c_off + c) % NEXP
};

__device__ __forceinline__ Unit decode_unit(int u, int nq_all, int split, const int32_t* __restrict__ lens, int S) {
    Unit U;
    int unit = u;
    U.t0 = 0;
    if (split) {
        U.t0 = unit % nq_all;
        unit /= nq_all;
    }
    U.b = unit / HEADS;
    U.h = unit % HEADS;
    int len = __ldg(lens + U.b);
    len = len < 1 ? 1 : (len > S ? S : len);
    U.len = len;
    U.nq = split ? 1 : nq_all;
    U.nsb = (len + SB - 1) / SB;
    U.nkb = (U.nsb + 1) / 2;
    U.total = U.nq * U.nsb;
    U.c_off = (U.t0 * U.nsb) % NEXP;
    return U;
}

// 640 threads (96 registers each at launch); the warpgroups re-balance WITHIN that pool of 640 x 96: 120 for the exp warps,
// 64 for the epilogue, 56 for the single-thread roles (their loops pace the kernel: no spills there);
// (asking for more than the launch allocation holds makes setmaxnreg.inc spin forever).
__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tctx,
                 const int32_t* __restrict__ lens, int S, int split, int n_units, unsigned long long* __restrict__ dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* k_full = bars + 0;       // [4]  K tile j of the unit has landed
    uint64_t* k_free = bars + 4;       // [4]  the unit's last S that reads K tile j has retired
    uint64_t* v_full = bars + 8;       // [4]
    uint64_t* v_free = bars + 12;      // [4]  the unit's last P.V that reads V tile j has retired
    uint64_t* q_full = bars + 16;      // [2]
    uint64_t* q_empty = bars + 18;     // [2]  the epilogue's TMA store has read the buffer back out
    uint64_t* o_done = bars + 20;      // [2]  the tile's last P.V has retired
    uint64_t* o_free = bars + 22;      // [2]  the epilogue has read the accumulator
    uint64_t* s_full = bars + 24;      // [NSLOT]  S_g landed
    uint64_t* p_ready = bars + 30;     // [NSLOT]  P_g has replaced S_g in the slot (all 128 rows)
    uint64_t* pv_done = bars + 36;     // [NSLOT]  the P.V reading the P in slot s has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 42);
    uint32_t* retired = tmem_slot + 1;  // sub-blocks of this CTA whose P.V has retired (written by the tracker thread only)
    uint32_t* slot_need = tmem_slot + 36;  // [NSLOT] 1 + index of the slot's last sub-block (S issuer's own bookkeeping)
#ifdef B200RT_DIAG
    uint32_t* prog = tmem_slot + 4;     // [8 roles][4]
    g_attn_prog = smem_u32(prog);
#endif

    griddep_launch_dependents();
    const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
    const int lane = lane_id();
    const int nq_all = (S + QT - 1) / QT;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tq);
        prefetch_tmap(&tctx);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < MAX_KB; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_free[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_free[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
            mbar_init(&o_done[i], 1);
            mbar_init(&o_free[i], 128);
        }
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_ready[i], 128);
        }
        for (int i = 0; i < NSLOT; ++i) mbar_init(&pv_done[i], 1);
        *reinterpret_cast<volatile uint32_t*>(retired) = 0;
        for (int i = 0; i < NSLOT; ++i) reinterpret_cast<volatile uint32_t*>(slot_need)[i] = 0;
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();  // the prologue above overlapped the previous kernel's tail; nothing it wrote (qkv, lens) has been read yet

    // (each role's code must be dominated by its own setmaxnreg for ptxas to allocate against the new budget)
    if (warp < 4) {
      setmaxnreg_dec<56>();
      if (warp == 0) {
        if (elect_one()) {
            // ------------------------------------------------------------ TMA producer
            // bit j of *_any: tile j has been loaded before; bit j of *_par: parity of the number of loads so far (bit masks, not
            // arrays: a dynamically indexed local array would live in local memory)
            uint32_t k_any = 0, k_par = 0, v_any = 0, v_par = 0;
            uint32_t gt = 0;  // the CTA's running query-tile index
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const Unit U = decode_unit(u, nq_all, split, lens, S);
                ATT_PROG(0, u, 0, 0);
                auto load_q = [&](int t) {
                    const uint32_t g = gt + t, qb = g & 1;
                    if (g >= 2) mbar_wait(&q_empty[qb], ((g >> 1) - 1) & 1);  // tile g-2 has left the buffer (epilogue)
                    mbar_arrive_expect_tx(&q_full[qb], TILE_BYTES);
                    tma_load_3d(smem + OFF_Q + qb * TILE_BYTES, &tq, &q_full[qb], U.h * D, (U.t0 + t) * QT, U.b);
                };
                // first what the unit's first S needs, then V, then the remaining query tiles: the wait for the second
                // query buffer only ends with the previous unit, and must not hold the K prefetch back
                load_q(0);
                for (int j = 0; j < U.nkb; ++j) {
                    if ((k_any >> j) & 1) mbar_wait(&k_free[j], ((k_par >> j) & 1) ^ 1);  // the previous user is done with K tile j
                    mbar_arrive_expect_tx(&k_full[j], TILE_BYTES);
                    tma_load_3d(smem + OFF_K + j * TILE_BYTES, &tq, &k_full[j], HIDDEN + U.h * D, j * KB, U.b);
                    k_any |= 1u << j;
                    k_par ^= 1u << j;
                }
                for (int j = 0; j < U.nkb; ++j) {
                    if ((v_any >> j) & 1) mbar_wait(&v_free[j], ((v_par >> j) & 1) ^ 1);
                    mbar_arrive_expect_tx(&v_full[j], TILE_BYTES);
                    tma_load_3d(smem + OFF_V + j * TILE_BYTES, &tq, &v_full[j], 2 * HIDDEN + U.h * D, j * KB, U.b);
                    v_any |= 1u << j;
                    v_par ^= 1u << j;
                }
                for (int t = 1; t < U.nq; ++t) load_q(t);
                gt += U.nq;
            }
        }
      } else if (warp == 1) {
        if (elect_one()) {
            // ------------------------------------------------------------ S issuer: S_c = Q . K_c^T
            // This thread's instruction stream paces the kernel once everything else overlaps (ncu: 130 instructions per
            // sub-block at ~7 cycles each made it the bottleneck), so the loop is nested by (unit, tile, sub-block) with
            // everything that can be hoisted hoisted: descriptor low words are running sums (a 128B-swizzled tile address
            // enters the descriptor as addr >> 4, so +32 bytes along K is +2 and one 64-key sub-block is +512), the slot
            // bookkeeping is one bit per warpgroup and one shared-memory word per slot.
            //
            // S slots belong to warpgroups: warpgroup w consumes S slots 2w and 2w+1 alternately.  A barrier must have ONE
            // waiter that meets its phases in order (a first wait for phase 1 of a barrier whose phase 0 the waiter never saw
            // passes at once): with slots handed out by the running sub-block index, a warpgroup that sat out the CTA's first
            // short units would start on the second phase of somebody else's slot.
            // A slot is free again once the P.V that read the P in it has RETIRED: slot_need[s] is 1 + the CTA-wide index of
            // the sub-block that used slot s last (0: never used), compared with the tracker's counter.
            constexpr uint32_t idesc_s = make_idesc_f16(QT, SB);  // 128 x 64, both K-major
            constexpr uint32_t kDescHi = static_cast<uint32_t>(make_sw128_desc(0) >> 32);
            const uint32_t q_lo0 = static_cast<uint32_t>(make_sw128_desc(smem_u32(smem + OFF_Q)));
            const uint32_t k_lo0 = static_cast<uint32_t>(make_sw128_desc(smem_u32(smem + OFF_K)));
            const uint32_t need_addr = smem_u32(slot_need);
            const uint32_t retired_addr = smem_u32(retired);
            uint32_t k_par = 0;  // bit j: parity of the number of units that used K tile j so far
            uint32_t par = 0;    // bit w: which of its two slots warpgroup w takes next
            uint32_t gt = 0;     // running query-tile index
            uint32_t gidx = 0;   // running sub-block index
#ifdef B200RT_DIAG
            int sdbg = 0;
#endif
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const Unit U = decode_unit(u, nq_all, split, lens, S);
                uint32_t wg = U.c_off;
                for (int t = 0; t < U.nq; ++t) {
                    const uint32_t g = gt + t;
                    mbar_wait(&q_full[g & 1], (g >> 1) & 1);
                    const uint32_t q_lo = q_lo0 + (g & 1) * (TILE_BYTES >> 4);
                    uint32_t k_lo = k_lo0;
                    const bool first = t == 0, last = t == U.nq - 1;
                    for (int sb = 0; sb < U.nsb; ++sb) {  // runs as far ahead as free slots allow
                        const uint32_t slot = 2 * wg + ((par >> wg) & 1);
                        par ^= 1u << wg;
                        ATT_PROG(1, u, t * U.nsb + sb, slot);
                        ATT_STAMP(5, sdbg, 0);
                        if (first && (sb & 1) == 0) mbar_wait(&k_full[sb >> 1], (k_par >> (sb >> 1)) & 1);  // first touch of the K tile
                        ATT_STAMP(5, sdbg, 1);
                        {
                            uint32_t need, v;
                            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(need) : "r"(need_addr + slot * 4) : "memory");
                            do {
                                asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(retired_addr) : "memory");
                            } while (v < need);
                        }
                        tc_fence_after();
                        ATT_STAMP(5, sdbg, 2);
                        const uint32_t d_tmem = tmem_base + TM_S + slot * SB;
#pragma unroll
                        for (int k = 0; k < D / 16; ++k) {
                            umma_f16_ss(d_tmem, (static_cast<uint64_t>(kDescHi) << 32) | (q_lo + 2 * k),
                                        (static_cast<uint64_t>(kDescHi) << 32) | (k_lo + 2 * k), idesc_s, k != 0);
                        }
                        umma_commit(&s_full[slot]);
                        ++gidx;
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(need_addr + slot * 4), "r"(gidx) : "memory");
                        ATT_STAMP(5, sdbg, 3);
#ifdef B200RT_DIAG
                        ++sdbg;
#endif
                        // the unit's last tile is through with K tile j after its odd sub-block (or the last one): free it
                        if (last && ((sb & 1) == 1 || sb == U.nsb - 1)) umma_commit(&k_free[sb >> 1]);
                        k_lo += (SB * 128) >> 4;
                        wg = wg == NEXP - 1 ? 0 : wg + 1;
                    }
                }
                k_par ^= (1u << U.nkb) - 1;
                gt += U.nq;
            }
        }
      } else if (warp == 2) {
        if (elect_one()) {
            // ------------------------------------------------------------ P.V retire tracker
            // The ONLY waiter on pv_done[]: it meets every phase of every one of them in the CTA's sub-block order and
            // publishes the count of retired sub-blocks.  mbarrier parity waits are only sound for a waiter that can be at
            // most one phase behind; the exp warpgroups are not (after a unit boundary the sub-block a rescale depends on may
            // belong to a warpgroup whose previous P.V they never waited for), so they read this counter instead before
            // touching the accumulator, and the S issuer reads it before handing a slot out again -- which also keeps every
            // pv_done[] (one per slot) at most one phase ahead of this thread.
            const uint32_t retired_addr = smem_u32(retired);
            uint32_t phases = 0, g = 0;
            uint32_t par = 0;  // bit w: parity of warpgroup w's sub-block count (== which of its two slots)
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const Unit U = decode_unit(u, nq_all, split, lens, S);
                uint32_t pb = U.c_off;
                for (int c = 0; c < U.total; ++c) {
                    ATT_PROG(2, u, c, g);
                    ATT_STAMP(6, g, 0);
                    const uint32_t tslot = 2 * pb + ((par >> pb) & 1);
                    par ^= 1u << pb;
                    mbar_wait(&pv_done[tslot], (phases >> tslot) & 1);
                    phases ^= 1u << tslot;
                    ATT_STAMP(6, g, 1);
                    ++g;
                    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(retired_addr), "r"(g) : "memory");
                    pb = pb == NEXP - 1 ? 0 : pb + 1;
                }
            }
        }
      } else if (warp == 3) {
        if (elect_one()) {
            // ------------------------------------------------------------ P.V issuer: O_t (+)= P_c . V_c
            // (same lean loop structure as the S issuer; A = P_c in tensor memory, 8 columns per 16-key step; B = V_c: one
            // key row is 128 B, so 16 keys are +128 in the descriptor's address field and a 64-key sub-block +512)
            constexpr uint32_t idesc_o = make_idesc_f16(QT, D, 0, 1);  // 128 x 64, B (= V) MN-major
            constexpr uint32_t kDescHi = static_cast<uint32_t>(make_sw128_desc(0) >> 32);
            const uint32_t v_lo0 = static_cast<uint32_t>(make_sw128_desc(smem_u32(smem + OFF_V)));
            uint32_t v_par = 0;
            uint32_t gt = 0;
            uint32_t phases = 0;  // bit s: parity of the phase of p_ready[s] awaited next
            uint32_t par = 0;     // bit w: parity of the number of sub-blocks warpgroup w has had (== which of its two S slots)
#ifdef B200RT_DIAG
            int gs_dbg = 0;
#endif
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const Unit U = decode_unit(u, nq_all, split, lens, S);
                uint32_t pb = U.c_off;
                for (int t = 0; t < U.nq; ++t) {
                    const uint32_t g = gt + t;
                    const uint32_t d_tmem = tmem_base + TM_O + (g & 1) * D;
                    uint32_t v_lo = v_lo0;
                    const bool first = t == 0, last = t == U.nq - 1;
                    for (int sb = 0; sb < U.nsb; ++sb) {
                        ATT_STAMP(4, gs_dbg, 0);
                        ATT_PROG(3, u, t * U.nsb + sb, pb);
                        // P_c is announced on its SLOT's barrier: a warpgroup may be two sub-blocks ahead of this thread (its
                        // two slots), so a per-warpgroup barrier could complete two phases before its first wait here
                        const uint32_t pslot = 2 * pb + ((par >> pb) & 1);
                        par ^= 1u << pb;
                        mbar_wait(&p_ready[pslot], (phases >> pslot) & 1);
                        phases ^= 1u << pslot;
                        ATT_STAMP(4, gs_dbg, 3);
                        if (sb == 0 && g >= 2) mbar_wait(&o_free[g & 1], ((g >> 1) - 1) & 1);  // tile g-2 has been written out
                        if (first && (sb & 1) == 0) mbar_wait(&v_full[sb >> 1], (v_par >> (sb >> 1)) & 1);
                        tc_fence_after();
                        ATT_STAMP(4, gs_dbg, 1);
                        const uint32_t p_tmem = tmem_base + TM_S + pslot * SB;  // P_c sits in its S slot
#pragma unroll
                        for (int kk = 0; kk < SB / 16; ++kk) {
                            umma_f16_ts(d_tmem, p_tmem + kk * 8, (static_cast<uint64_t>(kDescHi) << 32) | (v_lo + kk * ((16 * 128) >> 4)),
                                        idesc_o, (sb | kk) != 0);
                        }
                        umma_commit(&pv_done[pslot]);  // the P in this slot has been read; O_t is complete up to sub-block c
                        ATT_STAMP(4, gs_dbg, 2);
#ifdef B200RT_DIAG
                        ++gs_dbg;
#endif
                        if (last && ((sb & 1) == 1 || sb == U.nsb - 1)) umma_commit(&v_free[sb >> 1]);
                        if (sb == U.nsb - 1) umma_commit(&o_done[g & 1]);
                        v_lo += (SB * 128) >> 4;
                        pb = pb == NEXP - 1 ? 0 : pb + 1;
                        ATT_STAMP(4, gs_dbg - 1, 4);
                    }
                }
                v_par ^= (1u << U.nkb) - 1;
                gt += U.nq;
            }
        }
      }
    } else if (warp < 4 + 4 * NEXP) {
        // ---------------------------------------------------------------- exp warpgroup w
        setmaxnreg_inc<120>();
        const int w = (warp - 4) >> 2;
        const int wp = (w + NEXP - 1) % NEXP;  // the warpgroup that handles c - 1 inside a unit
        const int r = (warp & 3) * 32 + lane;  // query row within the tile == TMEM lane
        const uint32_t tm = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t mr_base = smem_u32(smem + OFF_MR) + r * 4;
        const uint32_t mr_self = mr_base + w * QT * 4;
        const uint32_t ls_self = smem_u32(smem + OFF_LS) + (w * QT + r) * 8;
        const bool obs = lane == 0 && ((warp & 3) == 0 || w == 0);
        const int ow = (warp & 3) == 0 ? w : 6 + (warp & 3);  // observer row: warp 0 of every warpgroup, all warps of warpgroup 0
        (void)obs; (void)ow;
        uint32_t scnt = 0;                    // sub-blocks this warpgroup has taken so far
        const uint32_t retired_addr = smem_u32(retired);
        auto wait_retired = [&](uint32_t n) {  // until the P.V of the CTA's first n sub-blocks have retired
            uint32_t v;
            do {
                asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(retired_addr) : "memory");
            } while (v < n);
        };
        uint32_t gs_base = 0, gt = 0;         // running sub-block / query-tile index of the unit's first
        int prev_last_wg = -1;                // warpgroup of the previous unit's last sub-block (-1: no previous unit)
        float l_w = 0.f, m_ref = 0.f;         // this warpgroup's partial row sum of the tile, relative to m_ref
#pragma unroll 1
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const Unit U = decode_unit(u, nq_all, split, lens, S);
            const int c_first_w = (w + NEXP - U.c_off) % NEXP;    // this warpgroup's first sub-block of the unit
            // the warpgroup of the NEXT unit's first sub-block (-1: this is the CTA's last unit)
            int next_first_wg = -1;
            if (u + static_cast<int>(gridDim.x) < n_units) next_first_wg = decode_unit(u + gridDim.x, nq_all, split, lens, S).c_off;
            int t = 0, sb = c_first_w;
            while (sb >= U.nsb && t < U.nq) {
                sb -= U.nsb;
                ++t;
            }
#pragma unroll 1
            for (int c = c_first_w; c < U.total; c += NEXP) {
                const uint32_t g = gs_base + c;
                const uint32_t slot = 2 * w + (scnt & 1);  // this warpgroup's own two S slots, alternately (see the S issuer)
                const uint32_t gtile = gt + t;
                if (obs) ATT_STAMP(ow, g, 0);
                if (obs) ATT_PROG(4 + w, u, c, 1);
                mbar_wait(&s_full[slot], (scnt >> 1) & 1);
                ++scnt;
                tc_fence_after();
                if (obs) ATT_STAMP(ow, g, 1);
                const int valid = U.len - sb * SB;  // keys [0, valid) of this sub-block are real (>= 1)
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
                // exceeds it by more than 2^8 in the exp2 domain.  The hand-off chain runs through EVERY consecutive pair of
                // sub-blocks of the CTA (also across units, where it only orders the reuse of the hand-off words); a pair
                // that falls to the same warpgroup is ordered by program order and skips the barrier.
                float m_used = mx, m_prev = mx;
                const int pred_wg = c > 0 ? wp : prev_last_wg;
                if (obs) ATT_PROG(4 + w, u, c, 2 + 16 * (pred_wg + 1));
                if (obs) ATT_STAMP(ow, g, 5);
                if (pred_wg >= 0 && pred_wg != w) named_bar_sync(BAR_MAX + pred_wg * NEXP + w, 256);  // the predecessor's max is in smem
                if (obs) ATT_STAMP(ow, g, 6);
                if (sb != 0) {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m_prev) : "r"(mr_base + wp * QT * 4) : "memory");
                    m_used = (mx - m_prev > kRescaleThreshold) ? mx : m_prev;
                }
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(mr_self), "f"(m_used) : "memory");
                const int succ_wg = c + 1 < U.total ? (w + 1) % NEXP : next_first_wg;
                if (obs) ATT_PROG(4 + w, u, c, 3 + 16 * (succ_wg + 1));
                if (succ_wg >= 0 && succ_wg != w) named_bar_arrive(BAR_MAX + w * NEXP + succ_wg, 256);
                if (sb < NEXP) {  // this warpgroup's first sub-block of the tile
                    l_w = 0.f;
                } else {
                    l_w *= ex2_approx((m_ref - m_used) * kScaleLog2e);
                }
                m_ref = m_used;
                if (obs) ATT_STAMP(ow, g, 2);
                if (sb != 0 && __any_sync(0xffffffffu, m_used != m_prev)) {
                    // rare: rescale this warp's 32 rows of the accumulator by 2^((m_prev - m_used) k) (1 where unchanged).
                    // The scores are dropped and read again afterwards so that this path costs the common one no registers.
                    // Every P.V before ours must have retired: O_t is then complete up to sub-block c-1 and nothing is in flight
                    // on it (the P.V of sub-block c waits for our P).
                    wait_retired(g);
                    tc_fence_after();
                    const float f = ex2_approx((m_prev - m_used) * kScaleLog2e);
                    const uint32_t o_addr = tm + TM_O + (gtile & 1) * D;
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
                // P_c overwrites the first 32 columns of its own S slot (this thread's row: loaded above); the S issuer hands
                // the slot out again only after the P.V that reads P_c has retired
                const float neg_ms = -m_used * kScaleLog2e;
                const uint32_t p_tm = tm + TM_S + slot * SB;
                if (obs) ATT_STAMP(ow, g, 3);
                float ls0 = 0.f, ls1 = 0.f;
                uint32_t pk_prev[4];
                // two scores per FFMA2 / FADD2 (same rounding, same summation order as the scalar loop: bit-identical)
                const uint64_t k2 = pack_f32x2(kScaleLog2e, kScaleLog2e), nm2 = pack_f32x2(neg_ms, neg_ms);
                uint64_t ls2 = pack_f32x2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x0, x1;
                        unpack_f32x2(fma2_f32(pack_f32x2(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1])), k2, nm2), x0, x1);
                        const float p0 = ex2_approx(x0);
                        const float p1 = ex2_approx(x1);
                        ls2 = add2_f32(ls2, pack_f32x2(p0, p1));
                        pk[e] = pack_half2(p0, p1);
                    }
                    if (q & 1) tmem_st_32x32b_x8(p_tm + (q >> 1) * 8, pk_prev[0], pk_prev[1], pk_prev[2], pk_prev[3], pk[0], pk[1], pk[2], pk[3]);
                    else { pk_prev[0] = pk[0]; pk_prev[1] = pk[1]; pk_prev[2] = pk[2]; pk_prev[3] = pk[3]; }
                }
                unpack_f32x2(ls2, ls0, ls1);
                l_w += ls0 + ls1;
                if (sb + NEXP >= U.nsb)  // this warpgroup's last sub-block of the tile
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ls_self + (gtile % MAX_NQ) * (NEXP * QT * 8)), "f"(m_ref), "f"(l_w) : "memory");
                tmem_st_wait();            // P_c is in tensor memory
                tc_fence_before();         // ... ahead of the MMA that reads it and accumulates into O_t
                mbar_arrive(&p_ready[slot]);  // (release: also publishes (m, l) to the epilogue via the o_done chain)
                if (obs) ATT_PROG(4 + w, u, c, 6);
                if (obs) ATT_STAMP(ow, g, 4);
                sb += NEXP;
                while (sb >= U.nsb && t < U.nq) {
                    sb -= U.nsb;
                    ++t;
                }
            }
            // bookkeeping for the next unit
            prev_last_wg = (U.c_off + U.total - 1) % NEXP;
            gs_base += U.total;
            gt += U.nq;
        }
    } else {
        // ---------------------------------------------------------------- epilogue warpgroup: ctx = O_t / l
        setmaxnreg_dec<64>();
        const int r = (warp & 3) * 32 + lane;
        const uint32_t tm = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t ls_base = smem_u32(smem + OFF_LS) + r * 8;
        const uint32_t swz = static_cast<uint32_t>(r & 7);
        const bool obs = (warp & 3) == 0 && lane == 0;
        (void)obs;
        uint32_t gt = 0;
#pragma unroll 1
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const Unit U = decode_unit(u, nq_all, split, lens, S);
#pragma unroll 1
            for (int t = 0; t < U.nq; ++t) {
                const uint32_t g = gt + t;
                if (obs) ATT_STAMP(3, g, 0);
                if (obs) ATT_PROG(7, u, t, g);
                mbar_wait(&o_done[g & 1], (g >> 1) & 1);
                tc_fence_after();
                if (obs) ATT_STAMP(3, g, 1);
                // l = sum of the warpgroups' partial row sums brought to the tile's final reference max (the last sub-block's)
                const int c_first = (U.t0 + t) * U.nsb;  // (item-wide index: decides which warpgroup had which sub-block)
                const int w_last = (c_first + U.nsb - 1) % NEXP;
                const uint32_t ls_tile = ls_base + (g % MAX_NQ) * (NEXP * QT * 8);
                float m_fin, l_fin;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m_fin), "=f"(l_fin) : "r"(ls_tile + w_last * (QT * 8)) : "memory");
#pragma unroll
                for (int w = 0; w < NEXP; ++w) {
                    const int first_sb = (w - c_first % NEXP + NEXP) % NEXP;  // warpgroup w's first sub-block in this tile
                    if (w != w_last && first_sb < U.nsb) {
                        float m_w, l_w;
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m_w), "=f"(l_w) : "r"(ls_tile + w * (QT * 8)) : "memory");
                        l_fin = fmaf(l_w, ex2_approx((m_w - m_fin) * kScaleLog2e), l_fin);
                    }
                }
                const float inv_l = 1.0f / l_fin;
                // staging tile = this tile's Q buffer: all of the tile's MMAs have retired (o_done), and tile g+2's Q is only
                // loaded into it once our TMA store has read it back out (q_empty below)
                const uint32_t o_row = smem_u32(smem + OFF_Q + (g & 1) * TILE_BYTES) + r * 128;
#pragma unroll
                for (int part = 0; part < D / 32; ++part) {
                    uint32_t o[32];
                    tmem_ld_32x32b_x32(tm + TM_O + (g & 1) * D + part * 32, o);
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
                mbar_arrive(&o_free[g & 1]);
                fence_proxy_async_smem();
                named_bar_sync(BAR_EPI, 128);
                if (warp == 4 + 4 * NEXP && lane == 0) {  // rows past S are clipped by the tensor map
                    tma_store_3d(&tctx, smem + OFF_Q + (g & 1) * TILE_BYTES, U.h * D, (U.t0 + t) * QT, U.b);
                    tma_store_commit();
                    tma_store_wait_read<0>();
                    mbar_arrive(&q_empty[g & 1]);
                }
                if (obs) ATT_STAMP(3, g, 2);
            }
            gt += U.nq;
        }
        if (warp == 4 + 4 * NEXP && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
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

cudaError_t launch_attention(const CUtensorMap& tq, const CUtensorMap& tctx, const int32_t* lens, int B, int S, int sm_count,
                             cudaStream_t stream, unsigned long long* dbg, bool pairs) {
    if (S < 1 || S > attn::MAX_KB * attn::KB || B < 1 || sm_count < 1) return cudaErrorInvalidValue;
    // fewer (item, head) units than SMs: split them by query tile (K and V are then loaded once per tile, but a single item
    // spreads over 48 SMs instead of 12)
    const int nq = (S + attn::QT - 1) / attn::QT;
    const int split = (B * HEADS < sm_count && nq > 1) ? 1 : 0;
    const int n_units = B * HEADS * (split ? nq : 1);
    int grid = n_units < sm_count ? n_units : sm_count;
    if (!pairs || grid < 2) {
        return launch_pdl(attn::attention_kernel, dim3(grid), dim3(attn::NUM_THREADS), attn::SMEM_BYTES, stream, tq, tctx, lens, S, split, n_units, dbg);
    }
    grid &= ~1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(attn::NUM_THREADS);
    cfg.dynamicSmemBytes = attn::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, attn::attention_kernel, tq, tctx, lens, S, split, n_units, dbg);
}

}  // namespace b200
