/* b200rt debug / test entry points: single kernels behind host buffers so that tests/ can compare each
 * one against the oracle.  Not part of the drop-in boundary; same conventions as b200rt.h.           */
#ifndef B200RT_DEBUG_H
#define B200RT_DEBUG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* out[M,N] = epi(A[M,K] . W[N,K]^T + bias (+ resid)) on replica 0.  a, w: fp16 bit patterns (uint16),
 * bias fp32 [N], resid fp32 [M,N] or NULL.  epi: 0 bias->fp16 out, 1 bias+gelu->fp16 out, 2 bias+resid->fp32
 * out (the residual travels through the kernel as fp16 hi + fp16 lo; out = hi + lo).  out is uint16 [M,N] for epi 0/1
 * and float [M,N] for epi 2.  ms_out (optional): device time of `iters` back-to-back launches divided by iters.
 * LayerNorm in the epilogue (optional; all NULL = none): ln_stats = [M][parts] (sum, M2) float pairs, one per 128 columns
 * of the LayerNorm'd row (parts = K/128 for epi 0/1, N/128 for epi 2).
 *   epi 0/1: `a` is the PRE-LayerNorm row in fp16, `w` = fp16(gamma o W with every row centred: sum_k w[n,k] = 0),
 *            bias = W beta + b:  out = epi(rstd * acc + bias)
 *   epi 2  : out = acc + bias + ((resid - mean) * rstd * ln_gamma + ln_beta); stats_out (optional) receives [M][N/128]
 *            (sum, M2) partials of the new rows.                                                             */
int b200rt_debug_gemm(int epi, const uint16_t* a, const uint16_t* w, const float* bias, const float* resid,
                      void* out, int M, int N, int K, int iters, float* ms_out, const float* ln_stats,
                      const float* ln_gamma, const float* ln_beta, float eps, float* stats_out);

/* ctx[B*S,768] (fp16 bits) = multi-head attention over qkv[B*S,2304] (fp16 bits), lens[B].           */
int b200rt_debug_attention(const uint16_t* qkv, const int32_t* lens, uint16_t* ctx, int B, int S, int iters,
                           float* ms_out);

/* Hidden state after `n_layers` encoder layers (0 = embedding LayerNorm output) of a loaded model,
 * hidden_out fp32 [n_items*max_len, hidden] on the host.                                              */
int b200rt_debug_hidden(int model, const int32_t* ids, const int32_t* lens, int n_items, int max_len, int n_layers,
                        float* hidden_out);

/* ViT model: the residual stream after `n_layers` encoder layers (0 = pre_layrnorm output) for pixels [n_items, 3, image,
 * image], hidden_out fp32 [n_items * tokens, hidden] on the host (pre-LN architecture: the stream itself, not a LayerNorm of it). */
int b200rt_debug_vit_hidden(int model, const float* pixels, int n_items, int n_layers, float* hidden_out);

/* Per-kernel device times (ms, CUDA events on the compute stream) of one forward of a resident batch:
 * names_out receives up to cap NUL-terminated names packed in a char buffer, ms_out the times.        */
int b200rt_debug_profile_forward(int model, int n_items, int max_len, int iters, char* names_out, size_t names_cap,
                                 float* ms_out, int* n_out, int cap);

#ifdef __cplusplus
}
#endif
#endif
