// Internal launch interface between the engine (host C++) and the sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum GemmEpilogue : int {
    EPI_BIAS_F16 = 0,       // out fp16 = LNfold(acc) + bias                     (fused QKV projection)
    EPI_BIAS_GELU_F16 = 1,  // out fp16 = gelu_erf(LNfold(acc) + bias)           (FFN up-projection)
    EPI_BIAS_RES_SPLIT = 2, // y (fp16 hi + fp16 lo, in place) = acc + bias + LN(y)  (attention-out / FFN down) + row statistics
    EPI_BIAS_QGELU_F16 = 3, // out fp16 = quick_gelu(LNfold(acc) + bias), x * sigmoid(1.702 x)   (CLIP MLP up-projection)
};

constexpr int HIDDEN = 768;
constexpr int HEADS = 12;
constexpr int HEAD_DIM = 64;
constexpr int QKV_DIM = 3 * HIDDEN;
constexpr int STAT_PARTS = HIDDEN / 128;  // row statistics travel as one (sum, M2) partial per 128 columns

// One-time per-device setup (dynamic shared memory opt-in for every kernel). Call with the device current.
cudaError_t kernels_init_device();

// The residual stream lives in HBM as y = hi + lo (two fp16 arrays, 22 mantissa bits together), stored PRE-LayerNorm, with
// the row statistics of the LayerNorm that applies to it as STAT_PARTS partials per row: (sum, M2 about the partial's own
// mean) over each 128-column slice, written by whoever produced the row (embedding kernel or a residual GEMM epilogue).
// No kernel ever materialises LN(y):
//   * a GEMM that consumes LN(y) reads `hi` as its fp16 A operand against W'' = fp16(gamma o W - rowmean(gamma o W)): gamma folded
//     into the weight's columns and every row centred at load time, so that sum_k (y_k - mean) W''_nk = sum_k y_k W''_nk and the
//     row mean never has to be subtracted; its epilogue applies  rstd * acc + (W beta + b)   -- "LNfold";
//   * a GEMM that adds the residual LN(y) re-applies (y - mean) * rstd * gamma + beta in its epilogue.
//
// C[M,N] = epi(A[M,K] . W[N,K]^T) on CTA pairs.  All operands move by TMA:
//   ta  : fp16 A   {K, rows>=M}  box {64,128} SW128     tb  : fp16 W {K, N} box {64,128} SW128
//   EPI 0/1: tout = fp16 out {N, rows} box {64,128} SW128; e.stats_in (or NULL: plain acc + bias), e.bias
//   EPI 2  : tout = hi, tlo = lo, both fp16 {N, rows} box {32,128} SW64 (read and written in place); e.stats_in + e.ln_gamma/
//            e.ln_beta describe the LayerNorm of the residual (stats_in NULL: the residual is added as is); e.stats_out
//            receives the partials of the new rows (N / 128 per row; must not alias stats_in: other tiles still read those).
//   N % 256 == 0, K % 64 == 0; rows past M are clipped by the maps' own bounds.
struct GemmEpi {
    const float* bias;        // [N]
    const float2* stats_in;   // [rows][parts_in]
    int parts_in;             // partials per row of stats_in (LayerNorm width / 128)
    const float* ln_gamma;    // [N] EPI 2
    const float* ln_beta;     // [N] EPI 2
    float2* stats_out;        // [rows][N / 128] EPI 2 (may be NULL)
    float eps;
};
cudaError_t launch_gemm(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap* tlo,
                        const GemmEpi& e, int M, int N, int K, int sm_count, cudaStream_t stream);

// gamma != NULL: w_out[n,k] = fp16(w[n,k] gamma[k] - mean_k(w[n,:] gamma)), cvec[n] = sum_k w[n,k] beta[k] + bias[n]
// gamma == NULL: w_out = fp16(w) (cvec untouched)
cudaError_t launch_fold_ln(const float* w, const float* gamma, const float* beta, const float* bias, __half* w_out, float* cvec,
                           int N, int K, cudaStream_t stream);

// Multi-head self-attention over a padded batch: qkv fp16 [B, S, 2304] (Q | K | V, head-major within each),
// lens[B] valid keys per item, ctx fp16 [B*S, 768].  tq: 3D map over qkv {2304, S, B}, tctx: 3D map over ctx {768, S, B};
// both box {64,128,1}, 128B swizzle.  Persistent: min(units, sm_count) CTAs walk the (item, head[, query tile]) units.
// dbg: B200RT_DIAG builds only -- device buffer of 5*32*8 clock stamps written by CTA 0 (NULL in the product path)
// pairs: launch the CTAs as clusters of 2 (no cooperation: only so that they occupy whole TPCs next to a CTA-pair GEMM)
cudaError_t launch_attention(const CUtensorMap& tq, const CUtensorMap& tctx, const int32_t* lens, int B, int S, int sm_count,
                             cudaStream_t stream, unsigned long long* dbg = nullptr, bool pairs = false);

// word + position + token_type(0) embedding gather -> y = hi + lo (pre-LN residual stream) + the row's statistic partials.
cudaError_t launch_embed(const int32_t* ids, const float* word, const float* pos, const float* type0, __half* yhi, __half* ylo,
                         float2* stats, int n_tokens, int S, int vocab, cudaStream_t stream);

// debug only: x32[row] = LN(hi + lo) with the row's own statistics (the product path never materialises this)
cudaError_t launch_ln_materialize(const __half* yhi, const __half* ylo, const float* gamma, const float* beta, float* x32,
                                  int n_rows, float eps, cudaStream_t stream);

// final LayerNorm of the CLS row of every item + L2 normalise; row i is stored at out + i * 768,
// where `out` may be a peer-mapped pointer into the root GPU's gather buffer (the fused gather).
cudaError_t launch_pool_normalize(const __half* yhi, const __half* ylo, const float* gamma, const float* beta, float* out,
                                  int n_items, int S, float eps, cudaStream_t stream);

// ViT image tower (CLIP): pixels fp32 [n, 3, img, img] -> im2col A fp16 [n*T, 3 p p] (class-token rows zero); patch GEMM output +
// class / position embeddings + pre_layrnorm -> residual stream + statistics; pooled head: post_layernorm(class token) ->
// projection [P, 768] -> L2 normalise, stored at out + i * P (possibly peer memory).
cudaError_t launch_im2col(const float* pixels, __half* a, int n_items, int img, int p, cudaStream_t stream);
cudaError_t launch_vit_embed(const __half* patch_out, const float* cls, const float* pos, const float* gamma, const float* beta,
                             __half* yhi, __half* ylo, float2* stats, int n_rows, int T, float eps, cudaStream_t stream);
cudaError_t launch_vit_pool(const __half* yhi, const __half* ylo, const float* gamma, const float* beta, const float* proj, float* out,
                            int n_items, int T, int P, float eps, cudaStream_t stream);

// fp32 -> fp16 (weight conversion at model load)
cudaError_t launch_f32_to_f16(const float* src, __half* dst, size_t n, cudaStream_t stream);

// Root-side scatter of a wave's token ids / lengths into each shard's (possibly peer) input slot.
struct ScatterPlan {
    static constexpr int MAX_SHARDS = 16;
    int n_shards;
    int S;  // padded length of every item in the wave
    int32_t* dst_ids[MAX_SHARDS];
    int32_t* dst_lens[MAX_SHARDS];
    int item_begin[MAX_SHARDS];
    int item_count[MAX_SHARDS];
};
cudaError_t launch_scatter(const int32_t* src_ids, const int32_t* src_lens, const ScatterPlan& plan,
                           cudaStream_t stream);

}  // namespace b200

#ifdef __CUDACC__
#include <utility>
// Launch with programmatic stream serialisation: the kernel's CTAs may be scheduled while the previous kernel of the
// stream drains (see ptx.cuh: griddep_launch_dependents / griddep_wait); the kernel itself waits before it reads anything
// the previous kernel wrote.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}
#endif
