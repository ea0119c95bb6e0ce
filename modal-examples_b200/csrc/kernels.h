// Internal launch interface between the engine (host C++) and the sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum GemmEpilogue : int {
    EPI_BIAS_F16 = 0,       // out fp16 = acc + bias                      (fused QKV projection)
    EPI_BIAS_GELU_F16 = 1,  // out fp16 = gelu_erf(acc + bias)            (FFN up-projection)
    EPI_BIAS_RES_F32 = 2,   // y fp32 (in place) = acc + bias + LN(y)     (attention-out / FFN down); LN(y) =
                            //   (y - mean) * rstd * gamma + beta from the row statistics of the previous LayerNorm
};

constexpr int HIDDEN = 768;
constexpr int HEADS = 12;
constexpr int HEAD_DIM = 64;
constexpr int QKV_DIM = 3 * HIDDEN;

// One-time per-device setup (dynamic shared memory opt-in for every kernel). Call with the device current.
cudaError_t kernels_init_device();

// C[M,N] = epi(A[M,K] . W[N,K]^T + bias) on CTA pairs.  All operands move by TMA (128B swizzle):
//   ta  : fp16 A   {K, rows>=M}  box {64,128}        tb  : fp16 W {K, N} box {64,128}
//   tout: fp16 out {N, rows} box {64,128} (EPI 0/1)  or  fp32 out {N, rows} box {32,128} (EPI 2)
//   EPI 2 reads the pre-LN residual through the same map `tout` (in place) and needs `ln`: the statistics and
//         affine of the LayerNorm that produced this GEMM's residual input (NULL for EPI 0/1).  With ln.stats ==
//         NULL the residual is added as is (plain `out += acc + bias`, used by the kernel-level tests).
//   N % 256 == 0, K % 64 == 0; rows past M are clipped by the maps' own bounds.
struct LnRef {
    const float2* stats;  // [rows] (mean, rstd)
    const float* gamma;   // [N]
    const float* beta;    // [N]
};
cudaError_t launch_gemm(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const LnRef* ln,
                        const float* bias, int M, int N, int K, int sm_count, cudaStream_t stream);

// Multi-head self-attention over a padded batch: qkv fp16 [B, S, 2304] (Q | K | V, head-major within each),
// lens[B] valid keys per item, ctx fp16 [B*S, 768].  tq: 3D map over qkv {2304, S, B}, tctx: 3D map over ctx {768, S, B};
// both box {64,128,1}, 128B swizzle.
// dbg: optional device buffer of 5*32*8 clock stamps written by CTA 0 (diagnostics; NULL in the product path)
cudaError_t launch_attention(const CUtensorMap& tq, const CUtensorMap& tctx, const int32_t* lens, int B, int S,
                             cudaStream_t stream, unsigned long long* dbg = nullptr);

// word + position + token_type(0) embedding gather -> y32 (fp32 pre-LN sum = residual stream), LayerNorm ->
// x16 (GEMM operand) + stats (mean, rstd per row).  x32_dbg (tests only, else NULL): normalised fp32 rows.
cudaError_t launch_embed_ln(const int32_t* ids, const float* word, const float* pos, const float* type0,
                            const float* gamma, const float* beta, float* y32, __half* x16, float2* stats, float* x32_dbg,
                            int n_tokens, int S, int vocab, float eps, cudaStream_t stream);

// LayerNorm over rows of y (fp32 pre-LN sum) -> x16 + stats; x32_dbg as above
cudaError_t launch_ln(const float* y, const float* gamma, const float* beta, __half* x16, float2* stats, float* x32_dbg,
                      int n_rows, float eps, cudaStream_t stream);

// final LayerNorm of the CLS row of every item + L2 normalise; row i is stored at out + (out_row0 + i) * 768,
// where `out` may be a peer-mapped pointer into the root GPU's gather buffer (the fused gather).
cudaError_t launch_pool_normalize(const float* y, const float* gamma, const float* beta, float* out, int n_items,
                                  int S, float eps, cudaStream_t stream);

// fp32 -> fp16 (weight conversion at model load)
cudaError_t launch_f32_to_f16(const float* src, __half* dst, size_t n, cudaStream_t stream);

// Root-side scatter of a wave's token ids / lengths into each shard's (possibly peer) input slot.
struct ScatterPlan {
    static constexpr int MAX_SHARDS = 8;
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
