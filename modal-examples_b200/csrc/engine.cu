// b200rt engine: replica pool, cold-start weight broadcast, the per-shard BERT forward, the C++ stream
// scheduler behind Function.map, and the C ABI declared in include/b200rt.h.
//
// Threads: callers (any number, any thread) -> submit queue -> dispatcher (forms waves, stages ids into
// pinned memory, H2D to the root GPU, scatter kernel into peer HBM, enqueues each shard's forward, whose
// last kernel stores straight into the root's gather buffer, then one D2H) -> completer (waits the wave's
// event, hands rows to the tickets' caller-owned buffers, wakes waiters).  No NCCL call and no host
// round-trip between H2D and D2H.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/b200rt.h"
#include "../../include/b200rt_debug.h"
#include "kernels.h"

namespace b200 {
cudaError_t gemm_init_device();
cudaError_t attention_init_device();
cudaError_t kernels_init_device() {
    cudaError_t e = gemm_init_device();
    if (e != cudaSuccess) return e;
    return attention_init_device();
}
}  // namespace b200

using namespace b200;

namespace {

// ------------------------------------------------------------------------------------------ errors

thread_local std::string t_last_error;
std::atomic<bool> g_poisoned{false};

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_last_error = buf;
    if (code == B200RT_E_CUDA) g_poisoned.store(true);
    return code;
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return fail(B200RT_E_CUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

// ------------------------------------------------------------------------------------------ tensor maps

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int load_driver_entry() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return fail(B200RT_E_CUDA, "cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

// fp16 2D row-major [rows, cols] -> box {64, box_rows}, 128B swizzle
int make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200RT_E_CUDA, "cuTensorMapEncodeTiled(2d %llu x %llu) -> %d",
                                       (unsigned long long)rows, (unsigned long long)cols, (int)r);
    return 0;
}

// fp16 2D row-major [rows, cols] -> box {32, 128} (64-byte rows), 64B swizzle: the residual epilogue's hi / lo staging chunks
int make_map_2d_chunk(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200RT_E_CUDA, "cuTensorMapEncodeTiled(chunk %llu x %llu) -> %d",
                                       (unsigned long long)rows, (unsigned long long)cols, (int)r);
    return 0;
}

// fp16 [B, S, width] (qkv: width 2304, ctx: width 768) -> box {64, 128, 1}, 128B swizzle (rows past S are zero-filled
// on load and clipped on store)
int make_map_qkv(CUtensorMap* m, const void* base, uint64_t B, uint64_t S, uint64_t width = QKV_DIM) {
    cuuint64_t dims[3] = {width, S, B};
    cuuint64_t strides[2] = {width * 2, S * width * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200RT_E_CUDA, "cuTensorMapEncodeTiled(qkv B=%llu S=%llu) -> %d",
                                       (unsigned long long)B, (unsigned long long)S, (int)r);
    return 0;
}

// ------------------------------------------------------------------------------------------ model / device state

constexpr int NSLOT = 4;
constexpr int MIN_ITEMS_PER_REPLICA = 8;  // a wave uses fewer replicas rather than giving one fewer items than this
constexpr int MAX_SEQ = 512;
constexpr int VIT_WAVE_ITEMS = 64;  // images one replica takes per wave (64 x 197 tokens; 38.5 MB of pixels per wave slot)

struct LayerW {
    __half *qkv_w, *ao_w, *ff1_w, *ff2_w;
    float *qkv_b, *ao_b, *ln1_g, *ln1_b, *ff1_b, *ff2_b, *ln2_g, *ln2_b;
    // LayerNorm folded into the consuming projections at load time (kernels.h):
    // qkv_w / ff1_w hold fp16(gamma o W, rows centred); *_c = W beta + b
    float *qkv_c, *ff1_c;
    CUtensorMap m_qkv, m_ao, m_ff1, m_ff2;  // box {64,128}: each CTA of a pair stages half of the tile's columns
};

struct DevWeights {
    float* f32_arena = nullptr;  // embeddings, biases, LayerNorm params
    __half* f16_arena = nullptr; // GEMM weights
    float *word, *pos, *type, *emb_g, *emb_b;
    std::vector<LayerW> layers;
    // ViT (kind "vit"): patch projection (fp16, the stride-p convolution as a [H, 3 p p] matrix), class / position embeddings,
    // pre / post LayerNorm, visual projection; `pos` above holds the position table.  For a ViT layer, LayerW::ln1_* is the
    // LayerNorm folded into QKV and ln2_* the one folded into FFN1 (pre-LN: neither is re-applied to the residual).
    __half* patch_w = nullptr;
    CUtensorMap m_patch;
    float *cls = nullptr, *pre_g = nullptr, *pre_b = nullptr, *post_g = nullptr, *post_b = nullptr, *proj_w = nullptr, *zero_bias = nullptr;
};

enum ModelKind { KIND_BERT = 0, KIND_VIT = 1 };

struct Model {
    ModelKind kind = KIND_BERT;
    b200rt_bert_config cfg{};
    b200rt_vit_config vcfg{};
    int tokens = 0;          // vit: (image / patch)^2 + 1
    int out_dim = HIDDEN;    // floats per item in the result
    size_t item_bytes = 0;   // vit: bytes of one item's pixels
    size_t f32_elems = 0, f16_elems = 0;
    std::vector<DevWeights> per_dev;
};

struct Dev {
    int id = -1;
    int sm_count = 148;
    int sm_gemm = 148, sm_attn = 148;  // SMs a GEMM / an attention launch of this replica asks for (see B200RT_INIT_SPLIT_SMS)
    bool attn_pairs = false;           // launch the attention CTAs as clusters of 2 so that they take whole TPCs
    cudaStream_t compute = nullptr;
    // workspace (capacity cap_rows rows)
    __half *yhi = nullptr, *ylo = nullptr;  // residual stream y = hi + lo, PRE-LayerNorm; hi doubles as the GEMM A operand
    float2* pstats[2] = {nullptr, nullptr};  // [rows][STAT_PARTS] (sum, M2) partials: [0] embedding LN / LN2 inputs, [1] LN1 inputs
    float* x32_dbg = nullptr;    // tests only (b200rt_debug_hidden): normalised fp32 rows, allocated on first use
    __half *qkv = nullptr, *ctx = nullptr, *ffn = nullptr;
    CUtensorMap m_yhi, m_ctx, m_ffn;      // fp16 [rows,*] box {64,128}: GEMM A operands (m_ffn is also FFN1's output map)
    CUtensorMap m_qkv2d;                  // fp16 [rows,2304] box {64,128}: QKV GEMM output
    CUtensorMap m_yhi_c, m_ylo_c;         // fp16 [rows,768] box {32,128} SW64: the residual epilogues update hi / lo in place
    std::unordered_map<int, std::pair<CUtensorMap, CUtensorMap>> m_qkv_by_S;  // (qkv, ctx) 3D maps per padded length
    // wave input slots (written by the root's scatter kernel, possibly over NVLink)
    int32_t* ids_in[NSLOT] = {};
    int32_t* lens_in[NSLOT] = {};
    float* pix_in[NSLOT] = {};     // vit: this replica's share of a wave's pixels (allocated when the first vit model loads)
    cudaStream_t copy = nullptr;   // vit: H2D of wave k+1's pixels runs here, under the forward of wave k on `compute`
    cudaEvent_t ev_pix[NSLOT] = {};
    int32_t* lens_const = nullptr; // vit: every item has `tokens` keys
    CUtensorMap m_im2col;          // vit: the ffn buffer viewed as the im2col matrix [rows, 768]
    cudaEvent_t ev_done[NSLOT];
    cudaEvent_t ev_begin[NSLOT], ev_end[NSLOT];  // timing: the forward itself on this replica's compute stream
    int last_slot = -1;                          // completer only: slot of this replica's previous wave (gap statistics)
    std::mutex mu;  // serialises host-side enqueue on this replica (scheduler vs. embed_device/debug)
    // Every forward on this replica uses the same workspace (yhi, ylo, qkv, ctx, ffn, pstats), whichever stream it is
    // enqueued on: ev_ws is recorded behind each forward and waited on ahead of the next one, so forwards from the
    // scheduler's stream and from callers' streams (b200rt_embed_device) execute one after the other on the device.
    cudaEvent_t ev_ws = nullptr;
    bool ev_ws_armed = false;
    // launcher thread: enqueues this replica's share of a wave (the ~85 launches per replica would otherwise be issued
    // by the one dispatcher thread for all replicas in turn)
    std::thread launcher;
    std::mutex lmu;
    std::condition_variable lcv;
    std::function<void()> ltask;
    bool lstop = false;
};

struct Model;
struct Ticket {
    uint64_t id;
    int model;
    const Model* mp = nullptr;  // resolved at submit time (the models vector may grow while the ticket is queued)
    std::vector<int32_t> ids;   // private copy [n, S] (empty when the caller lent us its buffer: `borrowed`)
    const int32_t* ids_ext = nullptr;
    const float* pixels = nullptr;  // vit tickets: [n_items, 3, image, image] fp32, caller-owned until completion
    bool borrowed = false, ids_pinned = false, out_pinned = false;
    const int32_t* ids_ptr() const { return borrowed ? ids_ext : ids.data(); }
    std::vector<int32_t> lens;
    int n_items, S;
    float* out;
    // Items travel in length buckets of 64 tokens (each bucket is padded to its own length, not to the input's max_len):
    // `order` lists the item indices sorted by bucket (stable), empty when that is the identity (all items in one bucket).
    std::vector<int32_t> order;
    int item_at(int pos) const { return order.empty() ? pos : order[pos]; }
    std::atomic<int> remaining; // items not yet delivered
    bool done = false;
    int waiters = 0;            // threads blocked in b200rt_wait on this ticket: poll_any must not hand it out
    int status = 0;
    std::string error;
};

// A run of a ticket's items that share a length bucket: positions [begin, begin + count) of the ticket's `order`.
struct Run {
    std::shared_ptr<Ticket> t;
    int begin, count, bucket;
    int next = 0;  // dispatcher cursor within the run
};

struct Segment {
    std::shared_ptr<Ticket> t;
    int ticket_off, count, wave_off;  // ticket_off: position in the ticket's `order`
};

struct Wave {
    int slot;
    int n_items, S;
    int first_dev = 0;        // lowest replica that took part (its events time the forward)
    uint32_t dev_mask = 0;    // replicas that took part
    bool direct_h2d = false;  // some segments were DMA'd from caller-pinned memory
    int out_dim = HIDDEN;     // floats per result row
    std::vector<Segment> segs;
};

struct Runtime {
    std::vector<std::unique_ptr<Dev>> devs;
    int cap_items = 0;   // per replica per wave, at 512 tokens; 0 = one item per SM of the first device (see rt_init)
    int cap_rows = 0;    // cap_items*512 rounded up to 128
    std::vector<std::unique_ptr<Model>> models;
    // root staging per slot
    int32_t *h_ids[NSLOT], *h_lens[NSLOT];
    float* h_out[NSLOT];
    int32_t *d_ids_stage[NSLOT], *d_lens_stage[NSLOT];
    float* d_out_gather[NSLOT];
    cudaStream_t s_in = nullptr, s_out = nullptr;  // on root
    cudaEvent_t ev_scatter[NSLOT], ev_wave[NSLOT], ev_t0[NSLOT], ev_fwd_end[NSLOT];
    // queues
    std::mutex mu;
    std::condition_variable cv_submit, cv_done, cv_slot, cv_wave, cv_idle;
    std::deque<Run> pending;  // runs in arrival order; a wave takes the front run's bucket and every later run of that bucket
    std::unordered_map<uint64_t, std::shared_ptr<Ticket>> tickets;
    std::deque<uint64_t> finished_unclaimed;
    std::deque<Wave> inflight;
    bool slot_busy[NSLOT] = {};
    int fill_window_us = 200;
    size_t vit_item_bytes = 0;  // bytes of one image of the loaded vit model(s)
    std::chrono::steady_clock::time_point last_submit = std::chrono::steady_clock::now();
    int active_calls = 0;  // threads inside b200rt_wait / b200rt_poll_any (shutdown waits for them to leave)
    uint64_t next_ticket = 1, next_wave = 0;
    bool stopping = false;
    std::thread dispatcher, completer;
    b200rt_stats_t stats{};
    int last_end_slot = -1, last_end_dev = -1;  // completer only
    std::mutex stats_mu;
    std::string async_error;
};

std::mutex g_rt_mu;
Runtime* g_rt = nullptr;

// ------------------------------------------------------------------------------------------ forward

struct Prof {
    std::vector<std::string> names;
    std::vector<cudaEvent_t> evs;
};

int get_qkv_map(Dev& d, int S, const CUtensorMap** out, const CUtensorMap** out_ctx) {
    auto it = d.m_qkv_by_S.find(S);
    if (it == d.m_qkv_by_S.end()) {
        CUtensorMap m, mc;
        const uint64_t B = static_cast<uint64_t>(g_rt->cap_rows) / S;
        int rc = make_map_qkv(&m, d.qkv, B, S);
        if (rc) return rc;
        rc = make_map_qkv(&mc, d.ctx, B, S, HIDDEN);
        if (rc) return rc;
        it = d.m_qkv_by_S.emplace(S, std::make_pair(m, mc)).first;
    }
    *out = &it->second.first;
    *out_ctx = &it->second.second;
    return 0;
}

// Enqueue the forward of one padded batch [B, S] on `stream`.  out: fp32 [B, 768], may be peer memory.
// n_layers >= 0 (debug): stop early and materialise the post-LN hidden state in d.x32_dbg.
int forward_enqueue(Dev& d, const Model& m, int dev_index, const int32_t* ids, const int32_t* lens, int B, int S, float* out,
                    cudaStream_t stream, int n_layers = -1, Prof* prof = nullptr, uint64_t* launches = nullptr) {
    const DevWeights& w = m.per_dev[dev_index];
    const b200rt_bert_config& c = m.cfg;
    const int L = n_layers < 0 ? c.layers : n_layers;
    const bool full = n_layers < 0;
    const int M = B * S;
    if (M > g_rt->cap_rows) return fail(B200RT_E_INVALID, "batch of %d x %d tokens exceeds wave capacity %d rows", B, S, g_rt->cap_rows);
    const CUtensorMap *mq = nullptr, *mc = nullptr;
    if (int rc = get_qkv_map(d, S, &mq, &mc)) return rc;
    uint64_t nl = 0;
    auto mark = [&](const char* name) {
        if (prof) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, stream);
            prof->names.push_back(name);
            prof->evs.push_back(e);
        }
    };
    mark("begin");
    float* const dbg = full ? nullptr : d.x32_dbg;
    CUDA_TRY(launch_embed(ids, w.word, w.pos, w.type, d.yhi, d.ylo, d.pstats[0], M, S, c.vocab, stream));
    ++nl; mark("embed");
    const float *prev_g = w.emb_g, *prev_b = w.emb_b;  // the LayerNorm that applies to the current residual rows
    for (int l = 0; l < L; ++l) {
        const LayerW& lw = w.layers[l];
        // QKV = LN_prev(y) Wqkv^T + b, the LayerNorm folded into the epilogue
        GemmEpi e_qkv{lw.qkv_c, d.pstats[0], STAT_PARTS, nullptr, nullptr, nullptr, c.eps};
        CUDA_TRY(launch_gemm(EPI_BIAS_F16, d.m_yhi, lw.m_qkv, d.m_qkv2d, nullptr, e_qkv, M, QKV_DIM, HIDDEN, d.sm_gemm, stream));
        ++nl; mark("gemm_qkv");
        CUDA_TRY(launch_attention(*mq, *mc, lens, B, S, d.sm_attn, stream, nullptr, d.attn_pairs));
        ++nl; mark("attention");
        // y <- ctx Wao^T + b + LN_prev(y); new row statistics (LN1's) into pstats[1]
        GemmEpi e_ao{lw.ao_b, d.pstats[0], STAT_PARTS, prev_g, prev_b, d.pstats[1], c.eps};
        CUDA_TRY(launch_gemm(EPI_BIAS_RES_SPLIT, d.m_ctx, lw.m_ao, d.m_yhi_c, &d.m_ylo_c, e_ao, M, HIDDEN, HIDDEN, d.sm_gemm, stream));
        ++nl; mark("gemm_attn_out");
        GemmEpi e_ff1{lw.ff1_c, d.pstats[1], STAT_PARTS, nullptr, nullptr, nullptr, c.eps};
        CUDA_TRY(launch_gemm(EPI_BIAS_GELU_F16, d.m_yhi, lw.m_ff1, d.m_ffn, nullptr, e_ff1, M, c.inter, HIDDEN, d.sm_gemm, stream));
        ++nl; mark("gemm_ffn1_gelu");
        // y <- ffn W2^T + b + LN1(y); LN2's statistics into pstats[0]
        GemmEpi e_ff2{lw.ff2_b, d.pstats[1], STAT_PARTS, lw.ln1_g, lw.ln1_b, d.pstats[0], c.eps};
        CUDA_TRY(launch_gemm(EPI_BIAS_RES_SPLIT, d.m_ffn, lw.m_ff2, d.m_yhi_c, &d.m_ylo_c, e_ff2, M, HIDDEN, c.inter, d.sm_gemm, stream));
        ++nl; mark("gemm_ffn2");
        prev_g = lw.ln2_g;
        prev_b = lw.ln2_b;
    }
    if (full) {
        CUDA_TRY(launch_pool_normalize(d.yhi, d.ylo, prev_g, prev_b, out, B, S, c.eps, stream));
        ++nl; mark("pool_normalize");
    } else {  // debug: materialise the post-LayerNorm hidden state
        CUDA_TRY(launch_ln_materialize(d.yhi, d.ylo, prev_g, prev_b, dbg, M, c.eps, stream));
        ++nl;
    }
    if (launches) *launches += nl;
    return 0;
}

// Forward of one batch of images on `stream`: pixels fp32 [B, 3, image, image] in this replica's HBM; out fp32 [B, proj],
// may be peer memory.  n_layers >= 0 (debug): stop after that many layers (the residual stream stays in yhi / ylo).
// CLIPVisionTransformer is pre-LN: LayerNorm1 / LayerNorm2 are folded into QKV / FFN1, the residual GEMMs add the raw
// stream (no LayerNorm re-applied) and emit the statistics the next fold needs.
int forward_enqueue_vit(Dev& d, const Model& m, int dev_index, const float* pixels, int B, float* out, cudaStream_t stream,
                        int n_layers = -1, uint64_t* launches = nullptr) {
    const DevWeights& w = m.per_dev[dev_index];
    const b200rt_vit_config& c = m.vcfg;
    const int T = m.tokens, M = B * T;
    const int L = n_layers < 0 ? c.layers : n_layers;
    if (M > g_rt->cap_rows || B > VIT_WAVE_ITEMS) return fail(B200RT_E_INVALID, "batch of %d images exceeds the wave capacity (%d)", B, VIT_WAVE_ITEMS);
    const CUtensorMap *mq = nullptr, *mc = nullptr;
    if (int rc = get_qkv_map(d, T, &mq, &mc)) return rc;
    uint64_t nl = 0;
    // patches -> im2col rows in the (idle) ffn buffer -> patch projection into the (idle) ctx buffer -> + class / position
    // embeddings, pre_layrnorm -> residual stream + statistics
    CUDA_TRY(launch_im2col(pixels, d.ffn, B, c.image, c.patch, stream));
    GemmEpi e_patch{w.zero_bias, nullptr, STAT_PARTS, nullptr, nullptr, nullptr, c.eps};
    CUDA_TRY(launch_gemm(EPI_BIAS_F16, d.m_im2col, w.m_patch, d.m_ctx, nullptr, e_patch, M, HIDDEN, 3 * c.patch * c.patch, d.sm_gemm, stream));
    CUDA_TRY(launch_vit_embed(d.ctx, w.cls, w.pos, w.pre_g, w.pre_b, d.yhi, d.ylo, d.pstats[0], M, T, c.eps, stream));
    nl += 3;
    for (int l = 0; l < L; ++l) {
        const LayerW& lw = w.layers[l];
        GemmEpi e_qkv{lw.qkv_c, d.pstats[0], STAT_PARTS, nullptr, nullptr, nullptr, c.eps};
        CUDA_TRY(launch_gemm(EPI_BIAS_F16, d.m_yhi, lw.m_qkv, d.m_qkv2d, nullptr, e_qkv, M, QKV_DIM, HIDDEN, d.sm_gemm, stream));
        CUDA_TRY(launch_attention(*mq, *mc, d.lens_const, B, T, d.sm_attn, stream, nullptr, d.attn_pairs));
        GemmEpi e_ao{lw.ao_b, nullptr, STAT_PARTS, nullptr, nullptr, d.pstats[1], c.eps};  // h += ctx Wo^T + b (raw residual)
        CUDA_TRY(launch_gemm(EPI_BIAS_RES_SPLIT, d.m_ctx, lw.m_ao, d.m_yhi_c, &d.m_ylo_c, e_ao, M, HIDDEN, HIDDEN, d.sm_gemm, stream));
        GemmEpi e_ff1{lw.ff1_c, d.pstats[1], STAT_PARTS, nullptr, nullptr, nullptr, c.eps};
        CUDA_TRY(launch_gemm(EPI_BIAS_QGELU_F16, d.m_yhi, lw.m_ff1, d.m_ffn, nullptr, e_ff1, M, c.inter, HIDDEN, d.sm_gemm, stream));
        GemmEpi e_ff2{lw.ff2_b, nullptr, STAT_PARTS, nullptr, nullptr, d.pstats[0], c.eps};
        CUDA_TRY(launch_gemm(EPI_BIAS_RES_SPLIT, d.m_ffn, lw.m_ff2, d.m_yhi_c, &d.m_ylo_c, e_ff2, M, HIDDEN, c.inter, d.sm_gemm, stream));
        nl += 5;
    }
    if (n_layers < 0) {
        CUDA_TRY(launch_vit_pool(d.yhi, d.ylo, w.post_g, w.post_b, w.proj_w, out, B, T, c.proj, c.eps, stream));
        ++nl;
    }
    if (launches) *launches += nl;
    return 0;
}

// Orders a ViT forward behind the previous forward on the replica's workspace, like forward() does (caller holds d.mu).
int forward_vit(Dev& d, const Model& m, int dev_index, const float* pixels, int B, float* out, cudaStream_t stream, int n_layers = -1,
                uint64_t* launches = nullptr) {
    if (d.ev_ws_armed) CUDA_TRY(cudaStreamWaitEvent(stream, d.ev_ws, 0));
    if (int rc = forward_enqueue_vit(d, m, dev_index, pixels, B, out, stream, n_layers, launches)) return rc;
    CUDA_TRY(cudaEventRecord(d.ev_ws, stream));
    d.ev_ws_armed = true;
    return 0;
}

// Small batches are launch-bound (a single 512-token item is ~85 launches of 5-40 us kernels): below
// GRAPH_MAX_ROWS rows the launch sequence is captured once per (buffers, shape) into a CUDA graph and replayed.
// Large waves keep plain launches (their kernels are long enough to hide the launch cost, and graphs per shape would
// only add memory).  The caller holds d.mu.
constexpr int GRAPH_MAX_ROWS = 8192;
constexpr size_t GRAPH_CACHE_MAX = 96;

struct GraphKey {
    const void *model, *ids, *lens, *out;
    cudaStream_t stream;
    int B, S;
    bool operator<(const GraphKey& o) const {
        return std::tie(model, ids, lens, out, stream, B, S) < std::tie(o.model, o.ids, o.lens, o.out, o.stream, o.B, o.S);
    }
};
struct GraphEntry {
    cudaGraphExec_t exec;
    uint64_t launches;
    uint64_t last_use;
};
struct GraphCache {
    std::map<GraphKey, GraphEntry> entries;
    uint64_t tick = 0;
};
std::mutex g_graph_mu;
std::map<Dev*, GraphCache> g_graphs;
GraphCache& graph_cache(Dev& d) {
    std::lock_guard<std::mutex> lk(g_graph_mu);
    return g_graphs[&d];
}
void drop_graphs(Dev& d) {  // the replica is going away: its cached launches point into freed buffers
    std::lock_guard<std::mutex> lk(g_graph_mu);
    auto it = g_graphs.find(&d);
    if (it == g_graphs.end()) return;
    for (auto& kv : it->second.entries) cudaGraphExecDestroy(kv.second.exec);
    g_graphs.erase(it);
}

// The caller holds d.mu.  Orders this forward behind the previous one on the replica's workspace (see Dev::ev_ws).
int forward(Dev& d, const Model& m, int dev_index, const int32_t* ids, const int32_t* lens, int B, int S, float* out,
            cudaStream_t stream, int n_layers = -1, Prof* prof = nullptr, uint64_t* launches = nullptr) {
    if (d.ev_ws_armed) CUDA_TRY(cudaStreamWaitEvent(stream, d.ev_ws, 0));
    int rc = 0;
    bool use_graph = !(n_layers >= 0 || prof || B * S > GRAPH_MAX_ROWS);
#ifdef B200RT_DIAG
    static const bool no_graphs = getenv("B200RT_NO_GRAPHS") != nullptr;
    if (no_graphs) use_graph = false;
#endif
    if (!use_graph) {
        rc = forward_enqueue(d, m, dev_index, ids, lens, B, S, out, stream, n_layers, prof, launches);
    } else {
        GraphCache& cache = graph_cache(d);
        const GraphKey key{&m, ids, lens, out, stream, B, S};
        auto it = cache.entries.find(key);
        if (it == cache.entries.end()) {
            if (cache.entries.size() >= GRAPH_CACHE_MAX) {  // evict the least recently replayed graph
                auto victim = cache.entries.begin();
                for (auto j = cache.entries.begin(); j != cache.entries.end(); ++j)
                    if (j->second.last_use < victim->second.last_use) victim = j;
                cudaGraphExecDestroy(victim->second.exec);
                cache.entries.erase(victim);
            }
            const CUtensorMap *mq = nullptr, *mc = nullptr;
            if (int r2 = get_qkv_map(d, S, &mq, &mc)) return r2;  // host-side map creation happens outside the capture
            uint64_t nl = 0;
            CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            rc = forward_enqueue(d, m, dev_index, ids, lens, B, S, out, stream, -1, nullptr, &nl);
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(stream, &graph);
            if (rc) {
                if (graph) cudaGraphDestroy(graph);
                return rc;
            }
            CUDA_TRY(ce);
            GraphEntry ent{nullptr, nl, 0};
            CUDA_TRY(cudaGraphInstantiate(&ent.exec, graph, 0));
            cudaGraphDestroy(graph);
            it = cache.entries.emplace(key, ent).first;
        }
        it->second.last_use = ++cache.tick;
        CUDA_TRY(cudaGraphLaunch(it->second.exec, stream));
        if (launches) *launches += it->second.launches;
    }
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(d.ev_ws, stream));
    d.ev_ws_armed = true;
    return 0;
}

// ------------------------------------------------------------------------------------------ scheduler

void finish_ticket_locked(Runtime& rt, const std::shared_ptr<Ticket>& t, int status, const std::string& err) {
    if (t->done) return;
    t->done = true;
    t->status = status;
    t->error = err;
    if (t->waiters == 0) rt.finished_unclaimed.push_back(t->id);  // a blocked b200rt_wait owns the ticket otherwise
    rt.stats.tickets++;
}

void fail_all(Runtime& rt, const std::string& err) {
    std::lock_guard<std::mutex> lk(rt.mu);
    rt.async_error = err;
    for (auto& kv : rt.tickets) finish_ticket_locked(rt, kv.second, B200RT_E_CUDA, err);
    rt.pending.clear();
    rt.cv_done.notify_all();
}

#define SCHED_TRY(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            g_poisoned.store(true);                                                                  \
            fail_all(rt, std::string(#expr) + ": " + cudaGetErrorString(_e));                        \
            return;                                                                                  \
        }                                                                                            \
    } while (0)

// Pinned allocations handed out by b200rt_alloc_pinned: buffers inside one are DMA targets/sources as they are
// (no staging copy through the scheduler's own pinned slots).
std::mutex g_pin_mu;
std::map<uintptr_t, size_t> g_pinned;  // base -> bytes
bool is_pinned(const void* p, size_t nbytes) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    auto it = g_pinned.upper_bound(a);
    if (it == g_pinned.begin()) return false;
    --it;
    return a >= it->first && a + nbytes <= it->first + it->second;
}

// Padded lengths are bucketed to multiples of 64 tokens (the attention kernel's key sub-block): tickets whose max_len
// falls into the same bucket share a wave, padded to the bucket size (at most 63 wasted tokens per item).
inline int bucket_of(int S) { return std::min(MAX_SEQ, (S + 63) / 64 * 64); }

void launcher_main(Dev* dp) {
    Dev& d = *dp;
    cudaSetDevice(d.id);
    for (;;) {
        std::function<void()> task;
        {
            std::unique_lock<std::mutex> lk(d.lmu);
            d.lcv.wait(lk, [&] { return d.lstop || d.ltask; });
            if (d.lstop && !d.ltask) return;
            task = std::move(d.ltask);
            d.ltask = nullptr;
        }
        task();
    }
}

void dispatcher_main(Runtime* rtp) {
    Runtime& rt = *rtp;
    const int G = static_cast<int>(rt.devs.size());
    Dev& root = *rt.devs[0];
    int rr = 0;  // first replica of the next wave that does not need the whole pool (whole inputs round-robin)
    for (;;) {
        Wave wv;
        {
            std::unique_lock<std::mutex> lk(rt.mu);
            rt.cv_submit.wait(lk, [&] { return rt.stopping || !rt.pending.empty(); });
            if (rt.stopping) return;
            // a free slot
            int slot = static_cast<int>(rt.next_wave % NSLOT);
            rt.cv_slot.wait(lk, [&] { return rt.stopping || !rt.slot_busy[slot]; });
            if (rt.stopping) return;
            if (rt.pending.empty()) continue;
            // Fill policy: a wave that would go out partial waits while submissions keep arriving -- until it is full, or no
            // ticket has arrived for a quiet period (pool busy: 200 us; idle: 40 us, the latency path), or a hard limit
            // (busy 2 ms, idle 1 ms).  A burst of small inputs hitting an idle pool (the start of every .map()) would otherwise
            // go out as a train of fragments, one per free wave slot.
            {
                const int S0 = rt.pending.front().bucket;
                const Model* m0 = rt.pending.front().t->mp;
                const int cap0 = (m0->kind == KIND_VIT ? VIT_WAVE_ITEMS : rt.cap_rows / S0) * G;
                auto queued = [&] {
                    int n = 0;
                    for (auto& r : rt.pending)
                        if (r.bucket == S0 && r.t->mp == m0) n += r.count - r.next;
                    return n;
                };
                bool busy = false;
                for (int s2 = 0; s2 < NSLOT; ++s2) busy |= rt.slot_busy[s2];
                const auto quiet = std::chrono::microseconds(busy ? rt.fill_window_us : rt.fill_window_us / 5);
                const auto hard = std::chrono::steady_clock::now() + std::chrono::microseconds(busy ? 10 * rt.fill_window_us : 5 * rt.fill_window_us);
                while (!rt.pending.empty() && queued() < cap0) {
                    const auto now = std::chrono::steady_clock::now();
                    const auto deadline = std::min(hard, rt.last_submit + quiet);
                    if (now >= deadline) break;
                    rt.cv_submit.wait_until(lk, deadline);
                    if (rt.stopping) return;
                }
                if (rt.pending.empty()) continue;
            }
            rt.slot_busy[slot] = true;
            rt.next_wave++;
            wv.slot = slot;
            wv.S = rt.pending.front().bucket;
            wv.n_items = 0;
            const Model* model = rt.pending.front().t->mp;
            // token capacity scales with 512/S: a wave holds cap_rows tokens per replica
            const int cap_items_S = (model->kind == KIND_VIT ? VIT_WAVE_ITEMS : rt.cap_rows / wv.S) * G;
            for (auto it = rt.pending.begin(); it != rt.pending.end() && wv.n_items < cap_items_S;) {
                if (it->bucket != wv.S || it->t->mp != model) {
                    ++it;
                    continue;
                }
                const int take = std::min(it->count - it->next, cap_items_S - wv.n_items);
                wv.segs.push_back(Segment{it->t, it->begin + it->next, take, wv.n_items});
                it->next += take;
                wv.n_items += take;
                it = it->next == it->count ? rt.pending.erase(it) : it + 1;
            }
        }
        const int slot = wv.slot, S = wv.S, n = wv.n_items;
        const Model& model = *wv.segs[0].t->mp;
        auto t_host0 = std::chrono::steady_clock::now();
        SCHED_TRY(cudaSetDevice(root.id));
        SCHED_TRY(cudaEventRecord(rt.ev_t0[slot], rt.s_in));
        // Stage the wave's ids on the root GPU.  A segment whose ids the caller lent us in pinned memory, at the
        // wave's own padded length, is DMA'd from where it lies; anything else goes through the slot's pinned buffer
        // (re-strided to the bucket length when the ticket's max_len is shorter).
        const bool vit = model.kind == KIND_VIT;
        const int out_dim = model.out_dim;
        bool staged = false;
        auto direct_ids = [&](const Segment& sg) {  // lent, pinned, already at the wave's padded length, items in place
            const Ticket& t = *sg.t;
            return t.borrowed && t.ids_pinned && t.S == S && t.order.empty();
        };
        for (const Segment& sg : wv.segs) {
            if (vit) break;  // image payloads are pulled by each replica itself (below): nothing is staged on the root
            const Ticket& t = *sg.t;
            int32_t* hdst = rt.h_ids[slot] + static_cast<size_t>(sg.wave_off) * S;
            if (direct_ids(sg)) {
                SCHED_TRY(cudaMemcpyAsync(rt.d_ids_stage[slot] + static_cast<size_t>(sg.wave_off) * S,
                                          t.ids_ptr() + static_cast<size_t>(sg.ticket_off) * t.S,
                                          static_cast<size_t>(sg.count) * S * 4, cudaMemcpyHostToDevice, rt.s_in));
                wv.direct_h2d = true;
                continue;
            }
            staged = true;
            if (t.S == S && t.order.empty()) {
                memcpy(hdst, t.ids_ptr() + static_cast<size_t>(sg.ticket_off) * t.S, static_cast<size_t>(sg.count) * S * 4);
            } else {  // re-stride to the bucket length (every item of the run has len <= S), gathering through `order`
                const int ncopy = std::min(t.S, S);
                for (int i = 0; i < sg.count; ++i) {
                    const int32_t* src = t.ids_ptr() + static_cast<size_t>(t.item_at(sg.ticket_off + i)) * t.S;
                    memcpy(hdst + static_cast<size_t>(i) * S, src, static_cast<size_t>(ncopy) * 4);
                    if (S > ncopy) memset(hdst + static_cast<size_t>(i) * S + ncopy, 0, static_cast<size_t>(S - ncopy) * 4);
                }
            }
        }
        if (!vit)
            for (const Segment& sg : wv.segs)
                for (int i = 0; i < sg.count; ++i) rt.h_lens[slot][sg.wave_off + i] = sg.t->lens[sg.t->item_at(sg.ticket_off + i)];
        auto t_host1 = std::chrono::steady_clock::now();
        if (staged) {
            if (!wv.direct_h2d) {
                SCHED_TRY(cudaMemcpyAsync(rt.d_ids_stage[slot], rt.h_ids[slot], static_cast<size_t>(n) * S * 4,
                                          cudaMemcpyHostToDevice, rt.s_in));
            } else {  // mixed wave: only the staged segments come from the slot buffer
                for (const Segment& sg : wv.segs) {
                    if (direct_ids(sg)) continue;
                    SCHED_TRY(cudaMemcpyAsync(rt.d_ids_stage[slot] + static_cast<size_t>(sg.wave_off) * S,
                                              rt.h_ids[slot] + static_cast<size_t>(sg.wave_off) * S,
                                              static_cast<size_t>(sg.count) * S * 4, cudaMemcpyHostToDevice, rt.s_in));
                }
            }
        }
        if (!vit)
            SCHED_TRY(cudaMemcpyAsync(rt.d_lens_stage[slot], rt.h_lens[slot], static_cast<size_t>(n) * 4,
                                      cudaMemcpyHostToDevice, rt.s_in));
        // Replicas used by this wave: everyone when there is enough work, otherwise as many as get at least
        // MIN_ITEMS_PER_REPLICA items each, starting from a rotating replica so that consecutive small waves land on
        // different GPUs.  Contiguous item ranges, as even as possible.
        ScatterPlan plan{};
        plan.n_shards = G;
        plan.S = S;
        int used = std::min(G, std::max(1, n / MIN_ITEMS_PER_REPLICA));
        const int first = used == G ? 0 : rr;
        if (used != G) rr = (rr + used) % G;
        const int per = (n + used - 1) / used;
        uint64_t peer_bytes = 0;
        for (int g = 0; g < G; ++g) {
            plan.dst_ids[g] = rt.devs[g]->ids_in[slot];
            plan.dst_lens[g] = rt.devs[g]->lens_in[slot];
            plan.item_begin[g] = 0;
            plan.item_count[g] = 0;
        }
        for (int k = 0; k < used; ++k) {
            const int g = (first + k) % G;
            const int b0 = std::min(n, k * per), b1 = std::min(n, (k + 1) * per);
            plan.item_begin[g] = b0;
            plan.item_count[g] = b1 - b0;
            if (g != 0) peer_bytes += static_cast<uint64_t>(b1 - b0) * (vit ? out_dim * 4 : S * 4 + 4 + out_dim * 4);
        }
        if (!vit) SCHED_TRY(launch_scatter(rt.d_ids_stage[slot], rt.d_lens_stage[slot], plan, rt.s_in));
        SCHED_TRY(cudaEventRecord(rt.ev_scatter[slot], rt.s_in));
        // every participating replica's launcher thread enqueues its own forward (~85 launches each, in parallel)
        std::atomic<uint64_t> launches{1};
        std::atomic<int> left{0};
        std::mutex jm;
        std::condition_variable jcv;
        std::string err;
        int err_code = 0;
        for (int g = 0; g < G; ++g)
            if (plan.item_count[g] > 0) left.fetch_add(1);
        wv.first_dev = -1;
        for (int g = 0; g < G; ++g) {
            if (plan.item_count[g] == 0) continue;
            if (wv.first_dev < 0) wv.first_dev = g;
            wv.dev_mask |= 1u << g;
            Dev& d = *rt.devs[g];
            auto task = [&, g]() {
                Dev& dd = *rt.devs[g];
                std::string e;
                int code = 0;
                auto ck = [&](cudaError_t ce, const char* what) {
                    if (ce != cudaSuccess && e.empty()) { e = std::string(what) + ": " + cudaGetErrorString(ce); code = B200RT_E_CUDA; }
                    return ce == cudaSuccess;
                };
                {
                    std::lock_guard<std::mutex> dl(dd.mu);
                    uint64_t nl = 0;
                    float* dst = rt.d_out_gather[slot] + static_cast<size_t>(plan.item_begin[g]) * out_dim;
                    if (ck(cudaStreamWaitEvent(dd.compute, rt.ev_scatter[slot], 0), "cudaStreamWaitEvent") &&
                        ck(cudaEventRecord(dd.ev_begin[slot], dd.compute), "cudaEventRecord")) {
                        int rc = 0;
                        if (vit) {
                            // this replica's items, segment by segment, straight from the callers' buffers over its own PCIe link
                            const int b0 = plan.item_begin[g], b1 = b0 + plan.item_count[g];
                            for (const Segment& sg : wv.segs) {
                                const int lo = std::max(b0, sg.wave_off), hi = std::min(b1, sg.wave_off + sg.count);
                                if (lo >= hi) continue;
                                const float* src = sg.t->pixels + static_cast<size_t>(sg.ticket_off + (lo - sg.wave_off)) * (model.item_bytes / 4);
                                if (!ck(cudaMemcpyAsync(dd.pix_in[slot] + static_cast<size_t>(lo - b0) * (model.item_bytes / 4), src,
                                                        static_cast<size_t>(hi - lo) * model.item_bytes, cudaMemcpyHostToDevice, dd.copy), "cudaMemcpyAsync(pixels)"))
                                    break;
                            }
                            if (e.empty() && ck(cudaEventRecord(dd.ev_pix[slot], dd.copy), "cudaEventRecord") &&
                                ck(cudaStreamWaitEvent(dd.compute, dd.ev_pix[slot], 0), "cudaStreamWaitEvent"))
                                rc = forward_vit(dd, model, g, dd.pix_in[slot], plan.item_count[g], dst, dd.compute, -1, &nl);
                        } else {
                            rc = forward(dd, model, g, dd.ids_in[slot], dd.lens_in[slot], plan.item_count[g], S, dst, dd.compute, -1, nullptr, &nl);
                        }
                        if (rc) { e = t_last_error; code = rc; }
                        else {
                            ck(cudaEventRecord(dd.ev_end[slot], dd.compute), "cudaEventRecord");
                            ck(cudaEventRecord(dd.ev_done[slot], dd.compute), "cudaEventRecord");
                        }
                    }
                    launches.fetch_add(nl);
                }
                std::lock_guard<std::mutex> jl(jm);
                if (!e.empty() && err.empty()) { err = e; err_code = code; }
                left.fetch_sub(1);
                jcv.notify_one();
            };
            if (used == 1 && g == 0) {
                task();  // the root's own share of a one-replica wave: no hand-off (latency path)
            } else {
                std::lock_guard<std::mutex> ll(d.lmu);
                d.ltask = task;
                d.lcv.notify_one();
            }
        }
        {
            std::unique_lock<std::mutex> jl(jm);
            jcv.wait(jl, [&] { return left.load() == 0; });
        }
        if (!err.empty()) {
            if (err_code == B200RT_E_CUDA) g_poisoned.store(true);
            fail_all(rt, err);
            return;
        }
        SCHED_TRY(cudaSetDevice(root.id));
        for (int g = 0; g < G; ++g)
            if (plan.item_count[g] > 0) SCHED_TRY(cudaStreamWaitEvent(rt.s_out, rt.devs[g]->ev_done[slot], 0));
        SCHED_TRY(cudaEventRecord(rt.ev_fwd_end[slot], rt.s_out));
        // D2H: rows of a ticket whose `out` lies in pinned memory go straight there; the rest through the slot buffer
        bool any_staged_out = false;
        for (const Segment& sg : wv.segs) {
            if (!(sg.t->out_pinned && sg.t->order.empty())) { any_staged_out = true; continue; }
            SCHED_TRY(cudaMemcpyAsync(sg.t->out + static_cast<size_t>(sg.ticket_off) * out_dim,
                                      rt.d_out_gather[slot] + static_cast<size_t>(sg.wave_off) * out_dim,
                                      static_cast<size_t>(sg.count) * out_dim * 4, cudaMemcpyDeviceToHost, rt.s_out));
        }
        if (any_staged_out)
            SCHED_TRY(cudaMemcpyAsync(rt.h_out[slot], rt.d_out_gather[slot], static_cast<size_t>(n) * out_dim * 4,
                                      cudaMemcpyDeviceToHost, rt.s_out));
        SCHED_TRY(cudaEventRecord(rt.ev_wave[slot], rt.s_out));
        {
            std::lock_guard<std::mutex> sl(rt.stats_mu);
            rt.stats.kernel_launches += launches.load();
            rt.stats.h2d_bytes += vit ? static_cast<uint64_t>(n) * model.item_bytes : static_cast<uint64_t>(n) * (S * 4 + 4);
            rt.stats.d2h_bytes += static_cast<uint64_t>(n) * out_dim * 4;
            rt.stats.peer_bytes += peer_bytes;
            rt.stats.stage_us += std::chrono::duration<double, std::micro>(t_host1 - t_host0).count();
            rt.stats.dispatch_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_host0).count();
        }
        wv.out_dim = out_dim;
        {
            std::lock_guard<std::mutex> lk(rt.mu);
            rt.inflight.push_back(std::move(wv));
            rt.cv_wave.notify_one();
        }
    }
}

void completer_main(Runtime* rtp) {
    Runtime& rt = *rtp;
    cudaSetDevice(rt.devs[0]->id);
    for (;;) {
        Wave wv;
        {
            std::unique_lock<std::mutex> lk(rt.mu);
            rt.cv_wave.wait(lk, [&] { return rt.stopping || !rt.inflight.empty(); });
            if (rt.inflight.empty()) {
                if (rt.stopping) return;
                continue;
            }
            wv = std::move(rt.inflight.front());
            rt.inflight.pop_front();
        }
        const int slot = wv.slot;
        cudaError_t e = cudaEventSynchronize(rt.ev_wave[slot]);
        if (e != cudaSuccess) {
            g_poisoned.store(true);
            fail_all(rt, std::string("wave failed: ") + cudaGetErrorString(e));
            return;
        }
        Dev& fd = *rt.devs[wv.first_dev < 0 ? 0 : wv.first_dev];
        float ms_in = 0, ms_fwd = 0, ms_out = 0;
        cudaEventElapsedTime(&ms_in, rt.ev_t0[slot], rt.ev_scatter[slot]);
        if (cudaEventElapsedTime(&ms_fwd, fd.ev_begin[slot], fd.ev_end[slot]) != cudaSuccess) { ms_fwd = 0; cudaGetLastError(); }
        float ms_gap = 0;
        if (rt.last_end_slot >= 0 && rt.last_end_dev == wv.first_dev &&
            cudaEventElapsedTime(&ms_gap, fd.ev_end[rt.last_end_slot], fd.ev_begin[slot]) != cudaSuccess) {
            ms_gap = 0;
            cudaGetLastError();
        }
        rt.last_end_slot = slot;
        rt.last_end_dev = wv.first_dev;
        cudaEventElapsedTime(&ms_out, rt.ev_fwd_end[slot], rt.ev_wave[slot]);
        float ms_fwd_max = 0, ms_gap_max = 0;  // over the participating replicas (each replica's events on its own device)
        for (size_t g = 0; g < rt.devs.size(); ++g) {
            if (!((wv.dev_mask >> g) & 1)) { rt.devs[g]->last_slot = -1; continue; }
            Dev& dg = *rt.devs[g];
            float f = 0, gp = 0;
            if (cudaEventElapsedTime(&f, dg.ev_begin[slot], dg.ev_end[slot]) != cudaSuccess) { f = 0; cudaGetLastError(); }
            if (dg.last_slot >= 0 && cudaEventElapsedTime(&gp, dg.ev_end[dg.last_slot], dg.ev_begin[slot]) != cudaSuccess) { gp = 0; cudaGetLastError(); }
            dg.last_slot = slot;
            ms_fwd_max = std::max(ms_fwd_max, f);
            ms_gap_max = std::max(ms_gap_max, gp);
        }
        {
            // rows of tickets that already failed or were shut down are not delivered: their owners may have freed `out`
            std::vector<const Segment*> live;
            {
                std::lock_guard<std::mutex> lk(rt.mu);
                for (const Segment& sg : wv.segs)
                    if (!sg.t->done && !(sg.t->out_pinned && sg.t->order.empty())) live.push_back(&sg);
            }
            const size_t od = static_cast<size_t>(wv.out_dim);
            for (const Segment* sg : live) {
                const float* src = rt.h_out[slot] + static_cast<size_t>(sg->wave_off) * od;
                if (sg->t->order.empty()) {
                    memcpy(sg->t->out + static_cast<size_t>(sg->ticket_off) * od, src, static_cast<size_t>(sg->count) * od * 4);
                } else {  // rows go back to the items' own positions
                    for (int i = 0; i < sg->count; ++i)
                        memcpy(sg->t->out + static_cast<size_t>(sg->t->order[sg->ticket_off + i]) * od, src + static_cast<size_t>(i) * od, od * 4);
                }
            }
        }
        {
            std::lock_guard<std::mutex> sl(rt.stats_mu);
            rt.stats.items += wv.n_items;
            rt.stats.waves++;
            rt.stats.h2d_scatter_us += ms_in * 1e3;
            rt.stats.forward_us += ms_fwd * 1e3;
            rt.stats.gap_us += (ms_gap > 0 ? ms_gap : 0) * 1e3;
            rt.stats.d2h_us += ms_out * 1e3;
            rt.stats.forward_max_us += ms_fwd_max * 1e3;
            rt.stats.gap_max_us += ms_gap_max * 1e3;
        }
        {
            std::lock_guard<std::mutex> lk(rt.mu);
            for (const Segment& sg : wv.segs) {
                if (sg.t->remaining.fetch_sub(sg.count) == sg.count) finish_ticket_locked(rt, sg.t, 0, "");
            }
            rt.slot_busy[slot] = false;
            rt.cv_slot.notify_all();
            rt.cv_done.notify_all();
        }
    }
}

// ------------------------------------------------------------------------------------------ init / load

int alloc_dev(Runtime& rt, Dev& d) {
    CUDA_TRY(cudaSetDevice(d.id));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, d.id));
    if (prop.major != 10)
        return fail(B200RT_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", d.id, prop.major, prop.minor);
    d.sm_count = prop.multiProcessorCount;
    if (d.sm_attn >= d.sm_count || d.sm_attn < 2) {  // whole-GPU replica
        d.sm_gemm = d.sm_attn = d.sm_count;
        d.attn_pairs = false;
    } else {  // two replicas share this GPU: GEMM launches leave sm_attn SMs to the other replica's attention launches
        d.sm_attn &= ~1;
        d.sm_gemm = (d.sm_count - d.sm_attn) & ~1;
        d.attn_pairs = true;
    }
    CUDA_TRY(kernels_init_device());
    CUDA_TRY(cudaStreamCreateWithFlags(&d.compute, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&d.ev_ws, cudaEventDisableTiming));
    const size_t R = rt.cap_rows;
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaMalloc(&d.pstats[i], R * STAT_PARTS * sizeof(float2)));
        CUDA_TRY(cudaMemset(d.pstats[i], 0, R * STAT_PARTS * sizeof(float2)));
    }
    CUDA_TRY(cudaMalloc(&d.yhi, R * HIDDEN * 2));
    CUDA_TRY(cudaMalloc(&d.ylo, R * HIDDEN * 2));
    CUDA_TRY(cudaMalloc(&d.qkv, R * QKV_DIM * 2));
    CUDA_TRY(cudaMalloc(&d.ctx, R * HIDDEN * 2));
    CUDA_TRY(cudaMalloc(&d.ffn, R * 3072 * 2));
    // padded / stale rows must stay finite (0 * NaN would leak through masked attention probabilities)
    CUDA_TRY(cudaMemset(d.yhi, 0, R * HIDDEN * 2));
    CUDA_TRY(cudaMemset(d.ylo, 0, R * HIDDEN * 2));
    CUDA_TRY(cudaMemset(d.qkv, 0, R * QKV_DIM * 2));
    CUDA_TRY(cudaMemset(d.ctx, 0, R * HIDDEN * 2));
    CUDA_TRY(cudaMemset(d.ffn, 0, R * 3072 * 2));
    if (int rc = make_map_2d(&d.m_yhi, d.yhi, R, HIDDEN, 128)) return rc;
    if (int rc = make_map_2d(&d.m_ctx, d.ctx, R, HIDDEN, 128)) return rc;
    if (int rc = make_map_2d(&d.m_ffn, d.ffn, R, 3072, 128)) return rc;
    if (int rc = make_map_2d(&d.m_qkv2d, d.qkv, R, QKV_DIM, 128)) return rc;
    if (int rc = make_map_2d_chunk(&d.m_yhi_c, d.yhi, R, HIDDEN)) return rc;
    if (int rc = make_map_2d_chunk(&d.m_ylo_c, d.ylo, R, HIDDEN)) return rc;
    for (int s = 0; s < NSLOT; ++s) {
        CUDA_TRY(cudaMalloc(&d.ids_in[s], R * 4));
        CUDA_TRY(cudaMalloc(&d.lens_in[s], R * 4));
        CUDA_TRY(cudaMemset(d.ids_in[s], 0, R * 4));
        CUDA_TRY(cudaEventCreateWithFlags(&d.ev_done[s], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreate(&d.ev_begin[s]));
        CUDA_TRY(cudaEventCreate(&d.ev_end[s]));
    }
    return 0;
}

int rt_init(const int* devices, int n, uint32_t flags) {
    std::lock_guard<std::mutex> lk(g_rt_mu);
    if (g_rt) return fail(B200RT_E_STATE, "b200rt already initialised");
    if (g_poisoned.load()) return fail(B200RT_E_CUDA, "context poisoned by an earlier CUDA error");
    if (n < 1 || n > ScatterPlan::MAX_SHARDS) return fail(B200RT_E_INVALID, "n_gpus must be in [1, %d], got %d", ScatterPlan::MAX_SHARDS, n);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(B200RT_E_CUDA, "no CUDA device available (%s); b200rt has no CPU path", cudaGetErrorString(e));
    auto rt = std::make_unique<Runtime>();
    if (flags & 0xFFFFu) rt->cap_items = static_cast<int>(flags & 0xFFFFu);  // B200RT_INIT_WAVE_ITEMS(n)
    if (rt->cap_items == 0) {
        // Default: as many 512-token items as the GPU has SMs.  A wave of B items is 6B / 18B / 24B pair-GEMM tiles and 12B
        // attention units, so with B = #SMs (= 2 x CTA pairs) every kernel of the forward runs a whole number of rounds
        // (148 SMs: 12 / 36 / 48 tiles per pair, 12 units per SM); 128 items left attn-out, FFN2 and attention at 10.4 -> 11
        // rounds (tools/wave_size_probe.py: 108.3 vs 110.2 us per item).
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devices ? devices[0] : 0) != cudaSuccess || sms < 2) sms = 128;
        rt->cap_items = sms & ~1;
    }
    if (rt->cap_items > 1024) return fail(B200RT_E_INVALID, "wave capacity of %d items per replica is out of range (1..1024)", rt->cap_items);
    rt->cap_rows = ((rt->cap_items * MAX_SEQ + 255) / 256) * 256;
    g_rt = rt.get();  // forward() and friends read capacity through g_rt
    auto bail = [&](int rc) { g_rt = nullptr; return rc; };
    if (int rc = load_driver_entry()) return bail(rc);
    const int split_sms = static_cast<int>((flags >> 16) & 0xFFu);  // B200RT_INIT_SPLIT_SMS(k): two replicas per GPU
    if (split_sms && 2 * n > ScatterPlan::MAX_SHARDS) return bail(fail(B200RT_E_INVALID, "split replicas: at most %d GPUs", ScatterPlan::MAX_SHARDS / 2));
    for (int i = 0; i < n; ++i) {
        for (int half = 0; half < (split_sms ? 2 : 1); ++half) {
            auto d = std::make_unique<Dev>();
            d->id = devices ? devices[i] : i;
            if (d->id < 0 || d->id >= count) return bail(fail(B200RT_E_INVALID, "device %d not present (%d visible)", d->id, count));
            d->sm_attn = split_sms ? split_sms : 1 << 20;
            if (int rc = alloc_dev(*rt, *d)) return bail(rc);
            rt->devs.push_back(std::move(d));
        }
    }
    n = static_cast<int>(rt->devs.size());
    // peer access both ways between every pair (scatter writes root->peer, gather writes peer->root)
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            if (i == j || rt->devs[i]->id == rt->devs[j]->id) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, rt->devs[i]->id, rt->devs[j]->id);
            if (!can) return bail(fail(B200RT_E_UNSUPPORTED, "no peer access between GPU %d and %d", rt->devs[i]->id, rt->devs[j]->id));
            cudaSetDevice(rt->devs[i]->id);
            cudaError_t pe = cudaDeviceEnablePeerAccess(rt->devs[j]->id, 0);
            if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (pe != cudaSuccess) return bail(fail(B200RT_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(pe)));
        }
    Dev& root = *rt->devs[0];
    if (cudaSetDevice(root.id) != cudaSuccess) return bail(fail(B200RT_E_CUDA, "cudaSetDevice(root)"));
    const size_t wave_rows = static_cast<size_t>(rt->cap_rows) * n;
    auto ck = [&](cudaError_t ce, const char* what) { return ce == cudaSuccess ? 0 : fail(B200RT_E_CUDA, "%s: %s", what, cudaGetErrorString(ce)); };
    if (int rc = ck(cudaStreamCreateWithFlags(&rt->s_in, cudaStreamNonBlocking), "stream")) return bail(rc);
    if (int rc = ck(cudaStreamCreateWithFlags(&rt->s_out, cudaStreamNonBlocking), "stream")) return bail(rc);
    for (int s = 0; s < NSLOT; ++s) {
        if (int rc = ck(cudaMallocHost(&rt->h_ids[s], wave_rows * 4), "cudaMallocHost ids")) return bail(rc);
        if (int rc = ck(cudaMallocHost(&rt->h_lens[s], wave_rows * 4), "cudaMallocHost lens")) return bail(rc);
        if (int rc = ck(cudaMallocHost(&rt->h_out[s], wave_rows * HIDDEN * 4), "cudaMallocHost out")) return bail(rc);
        if (int rc = ck(cudaMalloc(&rt->d_ids_stage[s], wave_rows * 4), "cudaMalloc")) return bail(rc);
        if (int rc = ck(cudaMalloc(&rt->d_lens_stage[s], wave_rows * 4), "cudaMalloc")) return bail(rc);
        if (int rc = ck(cudaMalloc(&rt->d_out_gather[s], wave_rows * HIDDEN * 4), "cudaMalloc")) return bail(rc);
        cudaEventCreate(&rt->ev_scatter[s]);
        cudaEventCreate(&rt->ev_wave[s]);
        cudaEventCreate(&rt->ev_t0[s]);
        cudaEventCreate(&rt->ev_fwd_end[s]);
    }
    for (auto& d : rt->devs) {
        cudaSetDevice(d->id);
        if (int rc = ck(cudaDeviceSynchronize(), "device sync after init")) return bail(rc);
    }
    for (auto& d : rt->devs) d->launcher = std::thread(launcher_main, d.get());
    rt->dispatcher = std::thread(dispatcher_main, rt.get());
    rt->completer = std::thread(completer_main, rt.get());
    g_rt = rt.release();
    return 0;
}

struct BlobCursor {
    size_t f32 = 0, f16 = 0;
};

int model_load(const b200rt_bert_config& c, const float* blob, size_t nbytes, int* model_out) {
    Runtime& rt = *g_rt;
    if (c.hidden != HIDDEN || c.heads != HEADS || c.inter != 3072 || c.max_pos > MAX_SEQ || c.max_pos < 1 ||
        c.layers < 1 || c.vocab < 1 || c.type_vocab < 1)
        return fail(B200RT_E_UNSUPPORTED,
                    "kernels are specialised for hidden=768 heads=12 inter=3072 max_pos<=512 (got %d/%d/%d/%d)", c.hidden,
                    c.heads, c.inter, c.max_pos);
    const size_t H = HIDDEN, I = 3072;
    const size_t emb = (static_cast<size_t>(c.vocab) + c.max_pos + c.type_vocab) * H + 2 * H;
    const size_t per_layer_w = 3 * H * H + H * H + I * H + H * I;
    const size_t per_layer_p = 3 * H + H + 2 * H + I + H + 2 * H;
    const size_t total = emb + static_cast<size_t>(c.layers) * (per_layer_w + per_layer_p);
    if (nbytes != total * 4) return fail(B200RT_E_INVALID, "weight blob is %zu bytes, geometry needs %zu", nbytes, total * 4);

    auto m = std::make_unique<Model>();
    m->cfg = c;
    const size_t per_layer_x = 3 * H + I;  // derived vectors: qkv_c, ff1_c
    const size_t blob_f32 = emb + static_cast<size_t>(c.layers) * per_layer_p;
    m->f32_elems = blob_f32 + static_cast<size_t>(c.layers) * per_layer_x;
    m->f16_elems = static_cast<size_t>(c.layers) * per_layer_w;
    m->per_dev.resize(rt.devs.size());

    // root: upload the fp32 blob once, carve fp32 params / convert GEMM weights to fp16 on the GPU.  The LayerNorm in front of
    // a projection is folded into it here (kernels.h): QKV of layer l takes the previous layer's LN2 (the embedding LN for
    // l = 0), FFN1 takes the layer's own LN1.
    Dev& root = *rt.devs[0];
    std::lock_guard<std::mutex> dl(root.mu);
    CUDA_TRY(cudaSetDevice(root.id));
    float* d_blob = nullptr;
    CUDA_TRY(cudaMalloc(&d_blob, nbytes));
    CUDA_TRY(cudaMemcpy(d_blob, blob, nbytes, cudaMemcpyHostToDevice));
    for (size_t g = 0; g < rt.devs.size(); ++g) {
        CUDA_TRY(cudaSetDevice(rt.devs[g]->id));
        CUDA_TRY(cudaMalloc(&m->per_dev[g].f32_arena, m->f32_elems * 4));
        CUDA_TRY(cudaMalloc(&m->per_dev[g].f16_arena, m->f16_elems * 2));
    }
    CUDA_TRY(cudaSetDevice(root.id));
    DevWeights& rw = m->per_dev[0];
    {
        size_t src = 0, o32 = 0, o16 = 0, ox = blob_f32;
        auto take32 = [&](size_t n) -> cudaError_t {
            cudaError_t e = cudaMemcpyAsync(rw.f32_arena + o32, d_blob + src, n * 4, cudaMemcpyDeviceToDevice, root.compute);
            src += n; o32 += n;
            return e;
        };
        // W [N,K] at the cursor -> fp16 (gamma folded in and rows centred when given; W beta + b into the derived region)
        auto take16 = [&](size_t N_, size_t K_, const float* gamma, const float* beta, const float* bias) -> cudaError_t {
            float* cv = gamma ? rw.f32_arena + ox : nullptr;
            cudaError_t e = launch_fold_ln(d_blob + src, gamma, beta, bias, rw.f16_arena + o16, cv, static_cast<int>(N_),
                                           static_cast<int>(K_), root.compute);
            src += N_ * K_; o16 += N_ * K_;
            if (gamma) ox += N_;
            return e;
        };
        const float* ln_g = d_blob + emb - 2 * H;  // the LayerNorm in front of the next QKV: embedding LN first
        const float* ln_b = d_blob + emb - H;
        CUDA_TRY(take32(emb));
        for (int l = 0; l < c.layers; ++l) {
            const float* qkv_b = d_blob + src + 3 * H * H;
            CUDA_TRY(take16(3 * H, H, ln_g, ln_b, qkv_b)); CUDA_TRY(take32(3 * H));      // qkv.w (folded), qkv.b
            CUDA_TRY(take16(H, H, nullptr, nullptr, nullptr));                            // ao.w
            const float* ln1_g = d_blob + src + H;
            const float* ln1_b = d_blob + src + 2 * H;
            CUDA_TRY(take32(H + 2 * H));                                                  // ao.b, ln1.g, ln1.b
            const float* ff1_b = d_blob + src + I * H;
            CUDA_TRY(take16(I, H, ln1_g, ln1_b, ff1_b)); CUDA_TRY(take32(I));            // ff1.w (folded), ff1.b
            CUDA_TRY(take16(H, I, nullptr, nullptr, nullptr));                            // ff2.w
            ln_g = d_blob + src + H;                                                      // ln2 of this layer feeds the next QKV
            ln_b = d_blob + src + 2 * H;
            CUDA_TRY(take32(H + 2 * H));                                                  // ff2.b, ln2.g, ln2.b
        }
    }
    CUDA_TRY(cudaStreamSynchronize(root.compute));
    CUDA_TRY(cudaFree(d_blob));
    // one-time broadcast of the device-resident weights to the other replicas over NVLink
    for (size_t g = 1; g < rt.devs.size(); ++g) {
        CUDA_TRY(cudaMemcpyPeerAsync(m->per_dev[g].f32_arena, rt.devs[g]->id, rw.f32_arena, root.id, m->f32_elems * 4, root.compute));
        CUDA_TRY(cudaMemcpyPeerAsync(m->per_dev[g].f16_arena, rt.devs[g]->id, rw.f16_arena, root.id, m->f16_elems * 2, root.compute));
    }
    CUDA_TRY(cudaStreamSynchronize(root.compute));
    // carve pointers + weight tensor maps per replica
    for (size_t g = 0; g < rt.devs.size(); ++g) {
        DevWeights& w = m->per_dev[g];
        CUDA_TRY(cudaSetDevice(rt.devs[g]->id));
        float* p = w.f32_arena;
        __half* q = w.f16_arena;
        w.word = p; p += static_cast<size_t>(c.vocab) * H;
        w.pos = p;  p += static_cast<size_t>(c.max_pos) * H;
        w.type = p; p += static_cast<size_t>(c.type_vocab) * H;
        w.emb_g = p; p += H;
        w.emb_b = p; p += H;
        w.layers.resize(c.layers);
        for (int l = 0; l < c.layers; ++l) {
            LayerW& lw = w.layers[l];
            lw.qkv_w = q; q += 3 * H * H;  lw.qkv_b = p; p += 3 * H;
            lw.ao_w = q;  q += H * H;      lw.ao_b = p;  p += H;  lw.ln1_g = p; p += H;  lw.ln1_b = p; p += H;
            lw.ff1_w = q; q += I * H;      lw.ff1_b = p; p += I;
            lw.ff2_w = q; q += H * I;      lw.ff2_b = p; p += H;  lw.ln2_g = p; p += H;  lw.ln2_b = p; p += H;
            float* x = w.f32_arena + blob_f32 + static_cast<size_t>(l) * per_layer_x;
            lw.qkv_c = x; lw.ff1_c = x + 3 * H;
            if (int rc = make_map_2d(&lw.m_qkv, lw.qkv_w, 3 * H, H, 128)) return rc;
            if (int rc = make_map_2d(&lw.m_ao, lw.ao_w, H, H, 128)) return rc;
            if (int rc = make_map_2d(&lw.m_ff1, lw.ff1_w, I, H, 128)) return rc;
            if (int rc = make_map_2d(&lw.m_ff2, lw.ff2_w, H, I, 128)) return rc;
        }
    }
    CUDA_TRY(cudaSetDevice(root.id));
    {
        std::lock_guard<std::mutex> lk(rt.mu);
        rt.models.push_back(std::move(m));
        *model_out = static_cast<int>(rt.models.size()) - 1;
    }
    return 0;
}

int model_load_vit(const b200rt_vit_config& c, const float* blob, size_t nbytes, int* model_out) {
    Runtime& rt = *g_rt;
    const int grid = c.patch > 0 ? c.image / c.patch : 0;
    const int T = grid * grid + 1;
    const size_t H = HIDDEN, I = 3072, PD = static_cast<size_t>(3) * c.patch * c.patch, P = static_cast<size_t>(c.proj);
    if (c.hidden != HIDDEN || c.heads != HEADS || c.inter != 3072 || c.layers < 1 || c.patch < 2 || c.image % c.patch != 0 || T > MAX_SEQ ||
        PD != HIDDEN || c.patch % 2 != 0 || c.proj < 1 || c.proj > 1024)
        return fail(B200RT_E_UNSUPPORTED, "kernels are specialised for ViT-B: hidden=768 heads=12 inter=3072, <= 512 tokens, 3*patch^2 %% 64 == 0, proj <= 1024 "
                                          "(got hidden %d heads %d inter %d image %d patch %d proj %d)", c.hidden, c.heads, c.inter, c.image, c.patch, c.proj);
    const size_t per_layer_w = 3 * H * H + H * H + I * H + H * I;
    const size_t per_layer_p = 2 * H + 3 * H + H + 2 * H + I + H;  // ln1, qkv.b, ao.b, ln2, ff1.b, ff2.b
    const size_t head = H * PD + H + static_cast<size_t>(T) * H + 2 * H, tail = 2 * H + P * H;
    const size_t total = head + static_cast<size_t>(c.layers) * (per_layer_w + per_layer_p) + tail;
    if (nbytes != total * 4) return fail(B200RT_E_INVALID, "weight blob is %zu bytes, geometry needs %zu", nbytes, total * 4);

    auto m = std::make_unique<Model>();
    m->kind = KIND_VIT;
    m->vcfg = c;
    m->tokens = T;
    m->out_dim = c.proj;
    m->item_bytes = static_cast<size_t>(3) * c.image * c.image * 4;
    const size_t per_layer_x = 3 * H + I;  // qkv_c, ff1_c
    m->f32_elems = (head - H * PD) + static_cast<size_t>(c.layers) * (per_layer_p + per_layer_x) + tail + H /* zero bias */;
    m->f16_elems = H * PD + static_cast<size_t>(c.layers) * per_layer_w;
    m->per_dev.resize(rt.devs.size());

    Dev& root = *rt.devs[0];
    std::lock_guard<std::mutex> dl(root.mu);
    CUDA_TRY(cudaSetDevice(root.id));
    float* d_blob = nullptr;
    CUDA_TRY(cudaMalloc(&d_blob, nbytes));
    CUDA_TRY(cudaMemcpy(d_blob, blob, nbytes, cudaMemcpyHostToDevice));
    for (size_t g = 0; g < rt.devs.size(); ++g) {
        Dev& d = *rt.devs[g];
        CUDA_TRY(cudaSetDevice(d.id));
        CUDA_TRY(cudaMalloc(&m->per_dev[g].f32_arena, m->f32_elems * 4));
        CUDA_TRY(cudaMalloc(&m->per_dev[g].f16_arena, m->f16_elems * 2));
        // per-replica image input slots and constants, once
        if (!d.lens_const) {
            std::vector<int32_t> lens(VIT_WAVE_ITEMS, T);
            CUDA_TRY(cudaMalloc(&d.lens_const, VIT_WAVE_ITEMS * 4));
            CUDA_TRY(cudaMemcpy(d.lens_const, lens.data(), VIT_WAVE_ITEMS * 4, cudaMemcpyHostToDevice));
            if (int rc = make_map_2d(&d.m_im2col, d.ffn, static_cast<uint64_t>(rt.cap_rows) * 4, HIDDEN, 128)) return rc;
        }
        if (!d.copy) CUDA_TRY(cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking));
        for (int sl = 0; sl < NSLOT; ++sl) {
            if (!d.pix_in[sl]) CUDA_TRY(cudaMalloc(&d.pix_in[sl], VIT_WAVE_ITEMS * m->item_bytes));
            if (!d.ev_pix[sl]) CUDA_TRY(cudaEventCreateWithFlags(&d.ev_pix[sl], cudaEventDisableTiming));
        }
    }
    if (rt.vit_item_bytes != 0 && rt.vit_item_bytes != m->item_bytes)
        return fail(B200RT_E_UNSUPPORTED, "a second vit model with a different image size is not supported in one runtime");
    rt.vit_item_bytes = m->item_bytes;
    CUDA_TRY(cudaSetDevice(root.id));
    DevWeights& rw = m->per_dev[0];
    {
        size_t src = 0, o32 = 0, o16 = 0;
        auto take32 = [&](size_t n) -> cudaError_t {
            cudaError_t e = cudaMemcpyAsync(rw.f32_arena + o32, d_blob + src, n * 4, cudaMemcpyDeviceToDevice, root.compute);
            src += n; o32 += n;
            return e;
        };
        auto skip32 = [&](size_t n) { o32 += n; };  // derived vectors are written in place by the fold kernel
        auto take16 = [&](size_t N_, size_t K_, const float* gamma, const float* beta, const float* bias, float* cv) -> cudaError_t {
            cudaError_t e = launch_fold_ln(d_blob + src, gamma, beta, bias, rw.f16_arena + o16, cv, static_cast<int>(N_), static_cast<int>(K_), root.compute);
            src += N_ * K_; o16 += N_ * K_;
            return e;
        };
        CUDA_TRY(take16(H, PD, nullptr, nullptr, nullptr, nullptr));   // patch.w
        CUDA_TRY(take32(H + static_cast<size_t>(T) * H + 2 * H));       // cls, pos, pre.g, pre.b
        for (int l = 0; l < c.layers; ++l) {
            const float* ln1_g = d_blob + src;
            const float* ln1_b = d_blob + src + H;
            CUDA_TRY(take32(2 * H));                                    // ln1.g, ln1.b
            const float* qkv_b = d_blob + src + 3 * H * H;
            float* qkv_c = rw.f32_arena + o32 + (3 * H) + H + 2 * H + I + H;  // after qkv.b ao.b ln2.g ln2.b ff1.b ff2.b
            CUDA_TRY(take16(3 * H, H, ln1_g, ln1_b, qkv_b, qkv_c));     // qkv.w folded with ln1
            CUDA_TRY(take32(3 * H));                                    // qkv.b
            CUDA_TRY(take16(H, H, nullptr, nullptr, nullptr, nullptr)); // ao.w
            CUDA_TRY(take32(H));                                        // ao.b
            const float* ln2_g = d_blob + src;
            const float* ln2_b = d_blob + src + H;
            CUDA_TRY(take32(2 * H));                                    // ln2.g, ln2.b
            const float* ff1_b = d_blob + src + I * H;
            float* ff1_c = rw.f32_arena + o32 + I + H + 3 * H;          // after ff1.b ff2.b qkv_c
            CUDA_TRY(take16(I, H, ln2_g, ln2_b, ff1_b, ff1_c));         // ff1.w folded with ln2
            CUDA_TRY(take32(I));                                        // ff1.b
            CUDA_TRY(take16(H, I, nullptr, nullptr, nullptr, nullptr)); // ff2.w
            CUDA_TRY(take32(H));                                        // ff2.b
            skip32(3 * H + I);                                          // qkv_c, ff1_c
        }
        CUDA_TRY(take32(2 * H + P * H));                                // post.g, post.b, proj.w
        CUDA_TRY(cudaMemsetAsync(rw.f32_arena + o32, 0, H * 4, root.compute));  // zero bias for the patch projection
        o32 += H;
        if (o32 != m->f32_elems || o16 != m->f16_elems || src != total) return fail(B200RT_E_STATE, "internal: vit blob carve mismatch");
    }
    CUDA_TRY(cudaStreamSynchronize(root.compute));
    CUDA_TRY(cudaFree(d_blob));
    for (size_t g = 1; g < rt.devs.size(); ++g) {
        CUDA_TRY(cudaMemcpyPeerAsync(m->per_dev[g].f32_arena, rt.devs[g]->id, rw.f32_arena, root.id, m->f32_elems * 4, root.compute));
        CUDA_TRY(cudaMemcpyPeerAsync(m->per_dev[g].f16_arena, rt.devs[g]->id, rw.f16_arena, root.id, m->f16_elems * 2, root.compute));
    }
    CUDA_TRY(cudaStreamSynchronize(root.compute));
    for (size_t g = 0; g < rt.devs.size(); ++g) {
        DevWeights& w = m->per_dev[g];
        CUDA_TRY(cudaSetDevice(rt.devs[g]->id));
        float* p = w.f32_arena;
        __half* q = w.f16_arena;
        w.patch_w = q; q += H * PD;
        if (int rc = make_map_2d(&w.m_patch, w.patch_w, H, PD, 128)) return rc;
        w.cls = p; p += H;
        w.pos = p; p += static_cast<size_t>(T) * H;
        w.pre_g = p; p += H;
        w.pre_b = p; p += H;
        w.layers.resize(c.layers);
        for (int l = 0; l < c.layers; ++l) {
            LayerW& lw = w.layers[l];
            lw.ln1_g = p; p += H;  lw.ln1_b = p; p += H;
            lw.qkv_w = q; q += 3 * H * H;  lw.qkv_b = p; p += 3 * H;
            lw.ao_w = q;  q += H * H;      lw.ao_b = p;  p += H;
            lw.ln2_g = p; p += H;  lw.ln2_b = p; p += H;
            lw.ff1_w = q; q += I * H;      lw.ff1_b = p; p += I;
            lw.ff2_w = q; q += H * I;      lw.ff2_b = p; p += H;
            lw.qkv_c = p; p += 3 * H;
            lw.ff1_c = p; p += I;
            if (int rc = make_map_2d(&lw.m_qkv, lw.qkv_w, 3 * H, H, 128)) return rc;
            if (int rc = make_map_2d(&lw.m_ao, lw.ao_w, H, H, 128)) return rc;
            if (int rc = make_map_2d(&lw.m_ff1, lw.ff1_w, I, H, 128)) return rc;
            if (int rc = make_map_2d(&lw.m_ff2, lw.ff2_w, H, I, 128)) return rc;
        }
        w.post_g = p; p += H;
        w.post_b = p; p += H;
        w.proj_w = p; p += P * H;
        w.zero_bias = p; p += H;
    }
    CUDA_TRY(cudaSetDevice(root.id));
    {
        std::lock_guard<std::mutex> lk(rt.mu);
        rt.models.push_back(std::move(m));
        *model_out = static_cast<int>(rt.models.size()) - 1;
    }
    return 0;
}

Runtime* live_rt() {
    if (g_poisoned.load()) {
        fail(B200RT_E_CUDA, "context poisoned by an earlier CUDA error%s%s", g_rt && !g_rt->async_error.empty() ? ": " : "",
             g_rt ? g_rt->async_error.c_str() : "");
        return nullptr;
    }
    if (!g_rt) {
        fail(B200RT_E_STATE, "b200rt_init has not been called");
        return nullptr;
    }
    return g_rt;
}

const Model* get_model(Runtime& rt, int model) {
    std::lock_guard<std::mutex> lk(rt.mu);
    if (model < 0 || model >= static_cast<int>(rt.models.size())) {
        fail(B200RT_E_INVALID, "unknown model handle %d", model);
        return nullptr;
    }
    return rt.models[model].get();
}

int check_ids(const b200rt_bert_config& c, const int32_t* ids, const int32_t* lens, int n_items, int max_len) {
    if (!ids || n_items < 1) return fail(B200RT_E_INVALID, "empty input (n_items=%d)", n_items);
    if (max_len < 1 || max_len > c.max_pos) return fail(B200RT_E_INVALID, "max_len %d outside [1, %d]", max_len, c.max_pos);
    for (int i = 0; i < n_items; ++i) {
        const int L = lens ? lens[i] : max_len;
        if (L < 1 || L > max_len) return fail(B200RT_E_INVALID, "lens[%d] = %d outside [1, %d]", i, L, max_len);
        const int32_t* row = ids + static_cast<size_t>(i) * max_len;
        for (int j = 0; j < L; ++j)
            if (row[j] < 0 || row[j] >= c.vocab)
                return fail(B200RT_E_INVALID, "ids[%d][%d] = %d outside the vocabulary [0, %d)", i, j, row[j], c.vocab);
    }
    return 0;
}

}  // namespace

// ============================================================================================ C ABI

extern "C" {

int b200rt_init(int n_gpus, uint32_t flags) { return rt_init(nullptr, n_gpus, flags); }
int b200rt_init_devices(const int* devices, int n_gpus, uint32_t flags) { return rt_init(devices, n_gpus, flags); }

int b200rt_num_gpus(void) {
    Runtime* rt = live_rt();
    return rt ? static_cast<int>(rt->devs.size()) : B200RT_E_STATE;
}

int b200rt_wave_capacity_items(void) {
    Runtime* rt = live_rt();
    return rt ? rt->cap_rows / MAX_SEQ : B200RT_E_STATE;
}

int b200rt_model_load(const char* kind, const void* cfg, const void* weights, size_t nbytes, int* model_out) {
    Runtime* rt = live_rt();
    if (!rt) return g_poisoned.load() ? B200RT_E_CUDA : B200RT_E_STATE;
    if (!kind || (strcmp(kind, "bert") != 0 && strcmp(kind, "vit") != 0))
        return fail(B200RT_E_UNSUPPORTED, "unknown model kind '%s' (\"bert\" or \"vit\")", kind ? kind : "(null)");
    if (!cfg || !weights || !model_out) return fail(B200RT_E_INVALID, "null argument");
    if (strcmp(kind, "vit") == 0)
        return model_load_vit(*static_cast<const b200rt_vit_config*>(cfg), static_cast<const float*>(weights), nbytes, model_out);
    return model_load(*static_cast<const b200rt_bert_config*>(cfg), static_cast<const float*>(weights), nbytes, model_out);
}

static int submit_impl(int model, const int32_t* ids, const int32_t* lens, int n_items, int max_len, float* out,
                       uint32_t flags, uint64_t* ticket_out) {
    Runtime* rt = live_rt();
    if (!rt) return g_poisoned.load() ? B200RT_E_CUDA : B200RT_E_STATE;
    const Model* m = get_model(*rt, model);
    if (!m) return B200RT_E_INVALID;
    if (!out || !ticket_out) return fail(B200RT_E_INVALID, "null out / ticket_out");
    if (m->kind != KIND_BERT) return fail(B200RT_E_INVALID, "model %d takes pixels (b200rt_submit_pixels), not token ids", model);
    if (flags & ~static_cast<uint32_t>(B200RT_SUBMIT_BORROW_IDS)) return fail(B200RT_E_INVALID, "unknown submit flags 0x%x", flags);
    if (int rc = check_ids(m->cfg, ids, lens, n_items, max_len)) return rc;
    auto t = std::make_shared<Ticket>();
    t->model = model;
    t->mp = m;
    t->n_items = n_items;
    t->S = max_len;
    t->out = out;
    t->remaining.store(n_items);
    const size_t id_bytes = static_cast<size_t>(n_items) * max_len * 4;
    if (flags & B200RT_SUBMIT_BORROW_IDS) {  // the caller keeps ids valid and unchanged until the ticket completes
        t->borrowed = true;
        t->ids_ext = ids;
        t->ids_pinned = is_pinned(ids, id_bytes);
    } else {
        t->ids.assign(ids, ids + static_cast<size_t>(n_items) * max_len);
    }
    t->out_pinned = is_pinned(out, static_cast<size_t>(n_items) * HIDDEN * 4);
    if (lens) t->lens.assign(lens, lens + n_items);
    else t->lens.assign(n_items, max_len);
    // runs of items per 64-token length bucket (TEI batches by token budget, un-padded; here every bucket is padded to its
    // own length, so a ragged input wastes at most 63 tokens per item instead of max_len - len)
    std::vector<Run> runs;
    {
        bool one_bucket = true;
        const int b0 = bucket_of(t->lens[0]);
        for (int i = 1; i < n_items && one_bucket; ++i) one_bucket = bucket_of(t->lens[i]) == b0;
        if (one_bucket) {
            runs.push_back(Run{t, 0, n_items, b0});
        } else {
            t->order.resize(n_items);
            for (int i = 0; i < n_items; ++i) t->order[i] = i;
            std::stable_sort(t->order.begin(), t->order.end(), [&](int a, int b) { return bucket_of(t->lens[a]) < bucket_of(t->lens[b]); });
            for (int i = 0; i < n_items;) {
                const int bk = bucket_of(t->lens[t->order[i]]);
                int j = i;
                while (j < n_items && bucket_of(t->lens[t->order[j]]) == bk) ++j;
                runs.push_back(Run{t, i, j - i, bk});
                i = j;
            }
        }
    }
    {
        std::lock_guard<std::mutex> lk(rt->mu);
        if (rt->stopping) return fail(B200RT_E_STATE, "runtime is shutting down");
        t->id = rt->next_ticket++;
        rt->tickets.emplace(t->id, t);
        for (auto& r : runs) rt->pending.push_back(r);
        rt->last_submit = std::chrono::steady_clock::now();
        *ticket_out = t->id;
    }
    rt->cv_submit.notify_one();
    return 0;
}

int b200rt_submit(int model, const int32_t* ids, const int32_t* lens, int n_items, int max_len, float* out,
                  uint64_t* ticket_out) {
    return submit_impl(model, ids, lens, n_items, max_len, out, 0, ticket_out);
}

int b200rt_submit_ex(int model, const int32_t* ids, const int32_t* lens, int n_items, int max_len, float* out,
                     uint32_t flags, uint64_t* ticket_out) {
    return submit_impl(model, ids, lens, n_items, max_len, out, flags, ticket_out);
}

int b200rt_submit_pixels(int model, const float* pixels, int n_items, float* out, uint64_t* ticket_out) {
    Runtime* rt = live_rt();
    if (!rt) return g_poisoned.load() ? B200RT_E_CUDA : B200RT_E_STATE;
    const Model* m = get_model(*rt, model);
    if (!m) return B200RT_E_INVALID;
    if (m->kind != KIND_VIT) return fail(B200RT_E_INVALID, "model %d takes token ids (b200rt_submit), not pixels", model);
    if (!pixels || !out || !ticket_out || n_items < 1) return fail(B200RT_E_INVALID, "null / empty argument");
    auto t = std::make_shared<Ticket>();
    t->model = model;
    t->mp = m;
    t->n_items = n_items;
    t->S = m->tokens;
    t->out = out;
    t->pixels = pixels;
    t->borrowed = true;
    t->remaining.store(n_items);
    t->out_pinned = is_pinned(out, static_cast<size_t>(n_items) * m->out_dim * 4);
    {
        std::lock_guard<std::mutex> lk(rt->mu);
        if (rt->stopping) return fail(B200RT_E_STATE, "runtime is shutting down");
        t->id = rt->next_ticket++;
        rt->tickets.emplace(t->id, t);
        rt->pending.push_back(Run{t, 0, n_items, -m->tokens});  // negative bucket keys: image waves never mix with text waves
        rt->last_submit = std::chrono::steady_clock::now();
        *ticket_out = t->id;
    }
    rt->cv_submit.notify_one();
    return 0;
}

static int reap_locked(Runtime& rt, uint64_t id, const std::shared_ptr<Ticket>& t) {
    for (auto it = rt.finished_unclaimed.begin(); it != rt.finished_unclaimed.end(); ++it)
        if (*it == id) { rt.finished_unclaimed.erase(it); break; }
    rt.tickets.erase(id);
    if (t->status != 0) return fail(t->status, "ticket %llu failed: %s", (unsigned long long)id, t->error.c_str());
    return 0;
}

// Entry/exit bookkeeping of the blocking calls: b200rt_shutdown waits until nobody is inside before it frees the
// runtime (callers are woken with "runtime shut down" first).
struct CallGuard {
    Runtime* rt = nullptr;
    std::unique_lock<std::mutex> lk;
    CallGuard() {
        std::lock_guard<std::mutex> g(g_rt_mu);
        rt = g_rt;
        if (rt) {
            lk = std::unique_lock<std::mutex>(rt->mu);
            rt->active_calls++;
        }
    }
    ~CallGuard() {
        if (rt) {
            if (!lk.owns_lock()) lk.lock();
            rt->active_calls--;
            rt->cv_idle.notify_all();
            lk.unlock();  // the last access to *rt: b200rt_shutdown may free it as soon as the count reaches zero
            lk.release();
        }
    }
};

int b200rt_wait(uint64_t ticket, int timeout_ms) {
    CallGuard g;
    Runtime* rt = g.rt;
    if (!rt) return fail(B200RT_E_STATE, "b200rt_init has not been called");
    auto& lk = g.lk;
    auto it = rt->tickets.find(ticket);
    if (it == rt->tickets.end()) return fail(B200RT_E_INVALID, "unknown or already reaped ticket %llu", (unsigned long long)ticket);
    std::shared_ptr<Ticket> t = it->second;
    // from here on the ticket belongs to this waiter: poll_any will not hand it out
    t->waiters++;
    for (auto f = rt->finished_unclaimed.begin(); f != rt->finished_unclaimed.end(); ++f)
        if (*f == ticket) { rt->finished_unclaimed.erase(f); break; }
    auto pred = [&] { return t->done; };
    bool ok = true;
    if (timeout_ms < 0) rt->cv_done.wait(lk, pred);
    else ok = rt->cv_done.wait_for(lk, std::chrono::milliseconds(timeout_ms), pred);
    t->waiters--;
    if (rt->tickets.find(ticket) == rt->tickets.end())  // a second waiter on the same ticket lost the race
        return fail(B200RT_E_INVALID, "ticket %llu was reaped by another waiter", (unsigned long long)ticket);
    if (!ok) {
        if (t->done && t->waiters == 0) rt->finished_unclaimed.push_back(ticket);  // (cannot happen: done => ok)
        return B200RT_TIMEOUT;
    }
    return reap_locked(*rt, ticket, t);
}

int b200rt_poll_any(uint64_t* ticket_out, int timeout_ms) {
    if (!ticket_out) return fail(B200RT_E_INVALID, "null ticket_out");
    CallGuard g;
    Runtime* rt = g.rt;
    if (!rt) return fail(B200RT_E_STATE, "b200rt_init has not been called");
    auto& lk = g.lk;
    auto pred = [&] { return !rt->finished_unclaimed.empty() || rt->stopping; };
    if (timeout_ms < 0) rt->cv_done.wait(lk, pred);
    else if (!rt->cv_done.wait_for(lk, std::chrono::milliseconds(timeout_ms), pred)) return B200RT_TIMEOUT;
    if (rt->finished_unclaimed.empty()) return fail(B200RT_E_STATE, "runtime shut down");
    const uint64_t id = rt->finished_unclaimed.front();
    auto t = rt->tickets[id];
    *ticket_out = id;
    return reap_locked(*rt, id, t);
}

int b200rt_embed_device(int model, int gpu, const int32_t* d_ids, const int32_t* d_lens, int n_items, int max_len,
                        float* d_out, void* stream) {
    Runtime* rt = live_rt();
    if (!rt) return g_poisoned.load() ? B200RT_E_CUDA : B200RT_E_STATE;
    const Model* m = get_model(*rt, model);
    if (!m) return B200RT_E_INVALID;
    if (gpu < 0 || gpu >= static_cast<int>(rt->devs.size())) return fail(B200RT_E_INVALID, "gpu index %d outside the pool", gpu);
    if (m->kind != KIND_BERT) return fail(B200RT_E_UNSUPPORTED, "b200rt_embed_device takes token ids: model %d is not a text encoder", model);
    if (!d_ids || !d_lens || !d_out || n_items < 1 || max_len < 1 || max_len > m->cfg.max_pos) return fail(B200RT_E_INVALID, "bad argument");
    Dev& d = *rt->devs[gpu];
    std::lock_guard<std::mutex> dl(d.mu);
    CUDA_TRY(cudaSetDevice(d.id));
    uint64_t launches = 0;
    int rc = forward(d, *m, gpu, d_ids, d_lens, n_items, max_len, d_out, stream ? static_cast<cudaStream_t>(stream) : d.compute,
                     -1, nullptr, &launches);
    std::lock_guard<std::mutex> sl(rt->stats_mu);
    rt->stats.kernel_launches += launches;
    return rc;
}

int b200rt_device_sync(int gpu) {
    Runtime* rt = live_rt();
    if (!rt) return g_poisoned.load() ? B200RT_E_CUDA : B200RT_E_STATE;
    if (gpu < 0 || gpu >= static_cast<int>(rt->devs.size())) return fail(B200RT_E_INVALID, "gpu index %d outside the pool", gpu);
    CUDA_TRY(cudaSetDevice(rt->devs[gpu]->id));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

void* b200rt_alloc_pinned(size_t nbytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, nbytes) != cudaSuccess) {
        cudaGetLastError();
        fail(B200RT_E_NOMEM, "cudaMallocHost(%zu) failed", nbytes);
        return nullptr;
    }
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        g_pinned[reinterpret_cast<uintptr_t>(p)] = nbytes;
    }
    return p;
}
void b200rt_free_pinned(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        g_pinned.erase(reinterpret_cast<uintptr_t>(p));
    }
    cudaFreeHost(p);
}

int b200rt_stats(b200rt_stats_t* out) {
    Runtime* rt = g_rt;
    if (!rt) return fail(B200RT_E_STATE, "b200rt_init has not been called");
    if (!out) return fail(B200RT_E_INVALID, "null out");
    std::lock_guard<std::mutex> sl(rt->stats_mu);
    *out = rt->stats;
    return 0;
}

const char* b200rt_last_error(void) { return t_last_error.c_str(); }

void b200rt_shutdown(void) {
    std::lock_guard<std::mutex> glk(g_rt_mu);
    Runtime* rt = g_rt;
    if (!rt) return;
    {
        std::lock_guard<std::mutex> lk(rt->mu);
        rt->stopping = true;
    }
    // 1. no new waves: the dispatcher leaves at its next wait; 2. the completer drains the waves already in flight
    // (their rows still go to live tickets' buffers) and leaves; 3. only then are the remaining tickets failed and
    // their waiters woken, so nothing writes a caller's `out` after that caller was told the runtime is gone.
    rt->cv_submit.notify_all();
    rt->cv_slot.notify_all();
    if (rt->dispatcher.joinable()) rt->dispatcher.join();
    rt->cv_wave.notify_all();
    if (rt->completer.joinable()) rt->completer.join();
    {
        std::unique_lock<std::mutex> lk(rt->mu);
        for (auto& kv : rt->tickets) finish_ticket_locked(*rt, kv.second, B200RT_E_STATE, "runtime shut down");
        rt->pending.clear();
        rt->cv_done.notify_all();
        rt->cv_idle.wait(lk, [&] { return rt->active_calls == 0; });  // blocked b200rt_wait / poll_any callers have left
    }
    for (auto& d : rt->devs) {
        {
            std::lock_guard<std::mutex> ll(d->lmu);
            d->lstop = true;
        }
        d->lcv.notify_all();
        if (d->launcher.joinable()) d->launcher.join();
    }
    for (auto& d : rt->devs) {
        cudaSetDevice(d->id);
        cudaDeviceSynchronize();
    }
    for (auto& m : rt->models)
        for (size_t g = 0; g < m->per_dev.size(); ++g) {
            cudaSetDevice(rt->devs[g]->id);
            cudaFree(m->per_dev[g].f32_arena);
            cudaFree(m->per_dev[g].f16_arena);
        }
    for (auto& d : rt->devs) {
        cudaSetDevice(d->id);
        drop_graphs(*d);
        cudaFree(d->x32_dbg); cudaFree(d->pstats[0]); cudaFree(d->pstats[1]); cudaFree(d->yhi); cudaFree(d->ylo); cudaFree(d->qkv); cudaFree(d->ctx); cudaFree(d->ffn);
        cudaFree(d->lens_const);
        if (d->copy) cudaStreamDestroy(d->copy);
        for (int s = 0; s < NSLOT; ++s) if (d->ev_pix[s]) cudaEventDestroy(d->ev_pix[s]);
        for (int s = 0; s < NSLOT; ++s) { cudaFree(d->pix_in[s]); cudaFree(d->ids_in[s]); cudaFree(d->lens_in[s]); cudaEventDestroy(d->ev_done[s]); cudaEventDestroy(d->ev_begin[s]); cudaEventDestroy(d->ev_end[s]); }
        cudaEventDestroy(d->ev_ws);
        cudaStreamDestroy(d->compute);
    }
    cudaSetDevice(rt->devs[0]->id);
    for (int s = 0; s < NSLOT; ++s) {
        cudaFreeHost(rt->h_ids[s]); cudaFreeHost(rt->h_lens[s]); cudaFreeHost(rt->h_out[s]);
        cudaFree(rt->d_ids_stage[s]); cudaFree(rt->d_lens_stage[s]); cudaFree(rt->d_out_gather[s]);
        cudaEventDestroy(rt->ev_scatter[s]); cudaEventDestroy(rt->ev_wave[s]); cudaEventDestroy(rt->ev_t0[s]); cudaEventDestroy(rt->ev_fwd_end[s]);
    }
    cudaStreamDestroy(rt->s_in);
    cudaStreamDestroy(rt->s_out);
    delete rt;
    g_rt = nullptr;
}

// ------------------------------------------------------------------------------------------ debug entry points

int b200rt_debug_gemm(int epi, const uint16_t* a, const uint16_t* w, const float* bias, const float* resid, void* out,
                      int M, int N, int K, int iters, float* ms_out, const float* ln_stats,
                      const float* ln_gamma, const float* ln_beta, float eps, float* stats_out) {
    Runtime* rt = live_rt();
    if (!rt) return g_poisoned.load() ? B200RT_E_CUDA : B200RT_E_STATE;
    if (!a || !w || !bias || !out || M < 1 || N % 256 || K % 64 || epi < 0 || (epi & 0xFF) > 2) return fail(B200RT_E_INVALID, "bad gemm arguments");
    const int epi_full = epi;  // bits 8+ = diagnostic mode of the pair kernel (B200RT_DIAG builds)
    epi &= 0xFF;
    if (ln_stats && epi != 2 && K % 128) return fail(B200RT_E_INVALID, "LayerNorm fold needs K % 128 == 0");
    if (ln_stats && epi == 2 && (!ln_gamma || !ln_beta)) return fail(B200RT_E_INVALID, "LayerNorm re-apply needs gamma and beta");
    Dev& d = *rt->devs[0];
    std::lock_guard<std::mutex> dl(d.mu);
    CUDA_TRY(cudaSetDevice(d.id));
    const size_t Mp = (static_cast<size_t>(M) + 255) / 256 * 256;
    const int parts_in = epi == 2 ? N / 128 : K / 128;
    const int parts_out = N / 128;
    __half *da = nullptr, *dw = nullptr, *dhi = nullptr, *dlo = nullptr, *dhi0 = nullptr, *dlo0 = nullptr;
    float *db = nullptr, *dg = nullptr, *dbt = nullptr;
    float2 *dsi = nullptr, *dso = nullptr;
    void* dout = nullptr;
    CUDA_TRY(cudaMalloc(&da, Mp * K * 2));
    CUDA_TRY(cudaMemset(da, 0, Mp * K * 2));
    CUDA_TRY(cudaMalloc(&dw, static_cast<size_t>(N) * K * 2));
    CUDA_TRY(cudaMalloc(&db, static_cast<size_t>(N) * 4));
    CUDA_TRY(cudaMemcpy(da, a, static_cast<size_t>(M) * K * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dw, w, static_cast<size_t>(N) * K * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(db, bias, static_cast<size_t>(N) * 4, cudaMemcpyHostToDevice));
    if (ln_stats) {
        CUDA_TRY(cudaMalloc(&dsi, Mp * parts_in * sizeof(float2)));
        CUDA_TRY(cudaMemset(dsi, 0, Mp * parts_in * sizeof(float2)));
        CUDA_TRY(cudaMemcpy(dsi, ln_stats, static_cast<size_t>(M) * parts_in * sizeof(float2), cudaMemcpyHostToDevice));
    }
    auto up = [&](float** dp, const float* h) -> cudaError_t {
        if (!h) return cudaSuccess;
        cudaError_t e = cudaMalloc(dp, static_cast<size_t>(N) * 4);
        return e != cudaSuccess ? e : cudaMemcpy(*dp, h, static_cast<size_t>(N) * 4, cudaMemcpyHostToDevice);
    };
    CUDA_TRY(up(&dg, ln_gamma));
    CUDA_TRY(up(&dbt, ln_beta));
    CUtensorMap ta, tb, tout, tlo;
    if (int rc = make_map_2d(&ta, da, Mp, K, 128)) return rc;
    if (int rc = make_map_2d(&tb, dw, N, K, 128)) return rc;
    std::vector<__half> hhi, hlo;
    if (epi == 2) {
        if (!resid) return fail(B200RT_E_INVALID, "epi 2 needs resid");
        // the residual travels as hi + lo (fp16 + fp16): split on the host, keep a pristine copy for the final launch
        hhi.assign(Mp * N, __float2half_rn(0.f));
        hlo.assign(Mp * N, __float2half_rn(0.f));
        for (size_t i = 0; i < static_cast<size_t>(M) * N; ++i) {
            hhi[i] = __float2half_rn(resid[i]);
            hlo[i] = __float2half_rn(resid[i] - __half2float(hhi[i]));
        }
        for (__half** pp : {&dhi, &dlo, &dhi0, &dlo0}) CUDA_TRY(cudaMalloc(pp, Mp * N * 2));
        CUDA_TRY(cudaMemcpy(dhi0, hhi.data(), Mp * N * 2, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dlo0, hlo.data(), Mp * N * 2, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dhi, dhi0, Mp * N * 2, cudaMemcpyDeviceToDevice));
        CUDA_TRY(cudaMemcpy(dlo, dlo0, Mp * N * 2, cudaMemcpyDeviceToDevice));
        CUDA_TRY(cudaMalloc(&dso, Mp * parts_out * sizeof(float2)));
        if (int rc = make_map_2d_chunk(&tout, dhi, Mp, N)) return rc;
        if (int rc = make_map_2d_chunk(&tlo, dlo, Mp, N)) return rc;
    } else {
        CUDA_TRY(cudaMalloc(&dout, Mp * N * 2));
        CUDA_TRY(cudaMemset(dout, 0xFF, Mp * N * 2));
        if (int rc = make_map_2d(&tout, dout, Mp, N, 128)) return rc;
    }
    GemmEpi e{db, dsi, parts_in, dg, dbt, dso, eps};
    // epi 2 updates hi / lo in place (y = acc + bias + LN(y)): reload the residual before the launch whose result is returned
    auto reload = [&]() -> cudaError_t {
        if (epi != 2) return cudaSuccess;
        cudaError_t ce = cudaMemcpyAsync(dhi, dhi0, Mp * N * 2, cudaMemcpyDeviceToDevice, d.compute);
        return ce != cudaSuccess ? ce : cudaMemcpyAsync(dlo, dlo0, Mp * N * 2, cudaMemcpyDeviceToDevice, d.compute);
    };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    if (iters < 1) iters = 1;
    CUDA_TRY(cudaEventRecord(e0, d.compute));
    for (int i = 0; i < iters; ++i) CUDA_TRY(launch_gemm(epi_full, ta, tb, tout, epi == 2 ? &tlo : nullptr, e, M, N, K, d.sm_count, d.compute));  // timing (values drift in place)
    CUDA_TRY(cudaEventRecord(e1, d.compute));
    CUDA_TRY(reload());
    CUDA_TRY(launch_gemm(epi_full, ta, tb, tout, epi == 2 ? &tlo : nullptr, e, M, N, K, d.sm_count, d.compute));  // the result that is returned
    CUDA_TRY(cudaStreamSynchronize(d.compute));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms / iters;
    if (epi == 2) {
        CUDA_TRY(cudaMemcpy(hhi.data(), dhi, Mp * N * 2, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(hlo.data(), dlo, Mp * N * 2, cudaMemcpyDeviceToHost));
        float* o = static_cast<float*>(out);
        for (size_t i = 0; i < static_cast<size_t>(M) * N; ++i) o[i] = __half2float(hhi[i]) + __half2float(hlo[i]);
        if (stats_out) CUDA_TRY(cudaMemcpy(stats_out, dso, static_cast<size_t>(M) * parts_out * sizeof(float2), cudaMemcpyDeviceToHost));
    } else {
        CUDA_TRY(cudaMemcpy(out, dout, static_cast<size_t>(M) * N * 2, cudaMemcpyDeviceToHost));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(da); cudaFree(dw); cudaFree(db); cudaFree(dg); cudaFree(dbt); cudaFree(dsi); cudaFree(dso);
    cudaFree(dhi); cudaFree(dlo); cudaFree(dhi0); cudaFree(dlo0); cudaFree(dout);
    return 0;
}

int b200rt_debug_attention(const uint16_t* qkv, const int32_t* lens, uint16_t* ctx, int B, int S, int iters, float* ms_out) {
    Runtime* rt = live_rt();
    if (!rt) return g_poisoned.load() ? B200RT_E_CUDA : B200RT_E_STATE;
    if (!qkv || !lens || !ctx || B < 1 || S < 1 || S > MAX_SEQ) return fail(B200RT_E_INVALID, "bad attention arguments");
    Dev& d = *rt->devs[0];
    std::lock_guard<std::mutex> dl(d.mu);
    CUDA_TRY(cudaSetDevice(d.id));
    const size_t M = static_cast<size_t>(B) * S;
    __half *dq = nullptr, *dc = nullptr;
    int32_t* dl_ = nullptr;
    CUDA_TRY(cudaMalloc(&dq, M * QKV_DIM * 2));
    CUDA_TRY(cudaMalloc(&dc, M * HIDDEN * 2));
    CUDA_TRY(cudaMalloc(&dl_, static_cast<size_t>(B) * 4));
    CUDA_TRY(cudaMemset(dc, 0xFF, M * HIDDEN * 2));
    CUDA_TRY(cudaMemcpy(dq, qkv, M * QKV_DIM * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dl_, lens, static_cast<size_t>(B) * 4, cudaMemcpyHostToDevice));
    CUtensorMap tq, tc;
    if (int rc = make_map_qkv(&tq, dq, B, S)) return rc;
    if (int rc = make_map_qkv(&tc, dc, B, S, HIDDEN)) return rc;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    if (iters < 1) iters = 1;
    CUDA_TRY(launch_attention(tq, tc, dl_, B, S, d.sm_count, d.compute));
    CUDA_TRY(cudaEventRecord(e0, d.compute));
    for (int i = 0; i < iters; ++i) CUDA_TRY(launch_attention(tq, tc, dl_, B, S, d.sm_count, d.compute));
    CUDA_TRY(cudaEventRecord(e1, d.compute));
    CUDA_TRY(cudaStreamSynchronize(d.compute));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms / iters;
    CUDA_TRY(cudaMemcpy(ctx, dc, M * HIDDEN * 2, cudaMemcpyDeviceToHost));
#ifdef B200RT_DIAG
    if (const char* path = getenv("B200RT_ATTN_STAMPS")) {  // diagnostics: per-phase clock stamps of CTA 0 -> text file
        unsigned long long* dstamp = nullptr;
        std::vector<unsigned long long> hs(10 * 32 * 8, 0);
        CUDA_TRY(cudaMalloc(&dstamp, hs.size() * 8));
        CUDA_TRY(cudaMemset(dstamp, 0, hs.size() * 8));
        CUDA_TRY(launch_attention(tq, tc, dl_, B, S, d.sm_count, d.compute, dstamp));
        CUDA_TRY(cudaStreamSynchronize(d.compute));
        CUDA_TRY(cudaMemcpy(hs.data(), dstamp, hs.size() * 8, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (auto v : hs) if (v && v < t0) t0 = v;
        if (FILE* f = fopen(path, "w")) {
            const char* names[10] = {"exp_wg0", "exp_wg1", "exp_wg2", "epilogue", "pv", "s_issue", "tracker", "wg0_warp1", "wg0_warp2", "wg0_warp3"};
            for (int o = 0; o < 10; ++o)
                for (int c = 0; c < 32; ++c) {
                    bool any = false;
                    for (int sl = 0; sl < 8; ++sl) any |= hs[(o * 32 + c) * 8 + sl] != 0;
                    if (!any) continue;
                    fprintf(f, "%s c%02d:", names[o], c);
                    for (int sl = 0; sl < 8; ++sl) {
                        unsigned long long v = hs[(o * 32 + c) * 8 + sl];
                        if (v) fprintf(f, " %llu", v - t0); else fprintf(f, " -");
                    }
                    fprintf(f, "\n");
                }
            fclose(f);
        }
        cudaFree(dstamp);
    }
#endif
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dq); cudaFree(dc); cudaFree(dl_);
    return 0;
}

int b200rt_debug_hidden(int model, const int32_t* ids, const int32_t* lens, int n_items, int max_len, int n_layers,
                        float* hidden_out) {
    Runtime* rt = live_rt();
    if (!rt) return g_poisoned.load() ? B200RT_E_CUDA : B200RT_E_STATE;
    const Model* m = get_model(*rt, model);
    if (!m) return B200RT_E_INVALID;
    if (m->kind != KIND_BERT) return fail(B200RT_E_INVALID, "model %d is not a text encoder", model);
    if (int rc = check_ids(m->cfg, ids, lens, n_items, max_len)) return rc;
    if (n_layers < 0 || n_layers > m->cfg.layers || !hidden_out) return fail(B200RT_E_INVALID, "bad n_layers / hidden_out");
    Dev& d = *rt->devs[0];
    std::lock_guard<std::mutex> dl(d.mu);
    CUDA_TRY(cudaSetDevice(d.id));
    const size_t M = static_cast<size_t>(n_items) * max_len;
    if (M > static_cast<size_t>(rt->cap_rows)) return fail(B200RT_E_INVALID, "batch exceeds wave capacity");
    std::vector<int32_t> l(n_items, max_len);
    if (lens) l.assign(lens, lens + n_items);
    int32_t *dids = nullptr, *dlens = nullptr;
    CUDA_TRY(cudaMalloc(&dids, M * 4));
    CUDA_TRY(cudaMalloc(&dlens, static_cast<size_t>(n_items) * 4));
    CUDA_TRY(cudaMemcpy(dids, ids, M * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dlens, l.data(), static_cast<size_t>(n_items) * 4, cudaMemcpyHostToDevice));
    if (!d.x32_dbg) CUDA_TRY(cudaMalloc(&d.x32_dbg, static_cast<size_t>(rt->cap_rows) * HIDDEN * 4));
    int rc = forward(d, *m, 0, dids, dlens, n_items, max_len, nullptr, d.compute, n_layers);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(d.compute));
    CUDA_TRY(cudaMemcpy(hidden_out, d.x32_dbg, M * HIDDEN * 4, cudaMemcpyDeviceToHost));
    cudaFree(dids); cudaFree(dlens);
    return 0;
}

int b200rt_debug_vit_hidden(int model, const float* pixels, int n_items, int n_layers, float* hidden_out) {
    Runtime* rt = live_rt();
    if (!rt) return g_poisoned.load() ? B200RT_E_CUDA : B200RT_E_STATE;
    const Model* m = get_model(*rt, model);
    if (!m) return B200RT_E_INVALID;
    if (m->kind != KIND_VIT || !pixels || !hidden_out || n_items < 1 || n_items > VIT_WAVE_ITEMS || n_layers < 0 || n_layers > m->vcfg.layers)
        return fail(B200RT_E_INVALID, "bad arguments");
    Dev& d = *rt->devs[0];
    std::lock_guard<std::mutex> dl(d.mu);
    CUDA_TRY(cudaSetDevice(d.id));
    CUDA_TRY(cudaMemcpy(d.pix_in[0], pixels, static_cast<size_t>(n_items) * m->item_bytes, cudaMemcpyHostToDevice));
    if (int rc = forward_vit(d, *m, 0, d.pix_in[0], n_items, nullptr, d.compute, n_layers)) return rc;
    CUDA_TRY(cudaStreamSynchronize(d.compute));
    const size_t n = static_cast<size_t>(n_items) * m->tokens * HIDDEN;
    std::vector<__half> hi(n), lo(n);
    CUDA_TRY(cudaMemcpy(hi.data(), d.yhi, n * 2, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(lo.data(), d.ylo, n * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) hidden_out[i] = __half2float(hi[i]) + __half2float(lo[i]);
    return 0;
}

int b200rt_debug_profile_forward(int model, int n_items, int max_len, int iters, char* names_out, size_t names_cap,
                                 float* ms_out, int* n_out, int cap) {
    Runtime* rt = live_rt();
    if (!rt) return g_poisoned.load() ? B200RT_E_CUDA : B200RT_E_STATE;
    const Model* m = get_model(*rt, model);
    if (!m) return B200RT_E_INVALID;
    if (m->kind != KIND_BERT) return fail(B200RT_E_INVALID, "model %d is not a text encoder", model);
    if (!names_out || !ms_out || !n_out || n_items < 1 || max_len < 1 || max_len > m->cfg.max_pos) return fail(B200RT_E_INVALID, "bad argument");
    Dev& d = *rt->devs[0];
    std::lock_guard<std::mutex> dl(d.mu);
    CUDA_TRY(cudaSetDevice(d.id));
    const size_t M = static_cast<size_t>(n_items) * max_len;
    if (M > static_cast<size_t>(rt->cap_rows)) return fail(B200RT_E_INVALID, "batch exceeds wave capacity");
    std::vector<int32_t> ids(M), l(n_items, max_len);
    uint32_t x = 12345;
    for (auto& v : ids) { x = x * 1664525u + 1013904223u; v = 1000 + static_cast<int32_t>((x >> 8) % (m->cfg.vocab - 1000)); }
    int32_t *dids = nullptr, *dlens = nullptr;
    float* dout = nullptr;
    CUDA_TRY(cudaMalloc(&dids, M * 4));
    CUDA_TRY(cudaMalloc(&dlens, static_cast<size_t>(n_items) * 4));
    CUDA_TRY(cudaMalloc(&dout, static_cast<size_t>(n_items) * HIDDEN * 4));
    CUDA_TRY(cudaMemcpy(dids, ids.data(), M * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dlens, l.data(), static_cast<size_t>(n_items) * 4, cudaMemcpyHostToDevice));
    if (iters < 1) iters = 1;
    std::map<std::string, double> acc;
    std::vector<std::string> order;
    for (int it = 0; it < iters + 1; ++it) {
        Prof prof;
        int rc = forward(d, *m, 0, dids, dlens, n_items, max_len, dout, d.compute, -1, &prof);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(d.compute));
        for (size_t i = 1; i < prof.evs.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, prof.evs[i - 1], prof.evs[i]);
            if (it > 0) {
                if (!acc.count(prof.names[i])) order.push_back(prof.names[i]);
                acc[prof.names[i]] += ms;
            }
        }
        for (auto e : prof.evs) cudaEventDestroy(e);
    }
    size_t off = 0;
    int n = 0;
    for (const auto& name : order) {
        if (n >= cap || off + name.size() + 1 > names_cap) break;
        memcpy(names_out + off, name.c_str(), name.size() + 1);
        off += name.size() + 1;
        ms_out[n++] = static_cast<float>(acc[name] / iters);
    }
    *n_out = n;
    cudaFree(dids); cudaFree(dlens); cudaFree(dout);
    return 0;
}

}  // extern "C"
