// HBM-bound row-wise kernels (sm_100a): embedding gather + LayerNorm, LayerNorm, CLS pooling +
// final LayerNorm + L2 normalise with the store aimed at the root GPU's gather buffer, the root-side
// scatter of token ids into (peer) shard inputs, and the fp32 -> fp16 weight conversion.
//
// All of them: one warp per 768-wide row, 128-bit loads/stores (6 float4 per lane), statistics by
// warp shuffle in fp32, two-pass variance on register-resident data (no E[x^2]-E[x]^2 cancellation).
//
// The fp32 residual stream is stored PRE-LayerNorm (y32) together with per-row (mean, rstd): the LayerNorm
// kernels write only the fp16 GEMM operand (x16) and the statistics (4.5 KB/row of traffic instead of 7.5),
// and the next residual-adding GEMM epilogue re-applies (y - mean) * rstd * gamma + beta with the very same
// fp32 operations, so the normalised fp32 row is bit-identical to the one a materialising kernel would write.
//
// Restates BertEmbeddings.forward (HF modeling_bert.py:72-111), the LayerNorm halves of
// BertSelfOutput :287-298 / BertOutput :345-356, and sentence-transformers' CLS pooling + Normalize
// (reference 06_gpu_and_ml/gpu_snapshot.py:58, `normalize_embeddings=True`).
#include "kernels.h"
#include "ptx.cuh"

namespace b200 {
namespace rw {

constexpr int H = HIDDEN;          // 768
constexpr int V4 = H / 4 / 32;     // float4 per lane = 6
constexpr int WARPS_PER_BLOCK = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
    return v;
}

// normalise the 24 register-resident values of this lane in place; returns (mean, rstd) of the row
__device__ __forceinline__ float2 ln_inplace(float4 (&x)[V4], const float* __restrict__ gamma,
                                             const float* __restrict__ beta, float eps, int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    const float mean = warp_sum(s) * (1.0f / H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
        q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const float4 g = __ldg(g4 + i * 32 + lane);
        const float4 b = __ldg(b4 + i * 32 + lane);
        x[i].x = fmaf(x[i].x * rstd, g.x, b.x);
        x[i].y = fmaf(x[i].y * rstd, g.y, b.y);
        x[i].z = fmaf(x[i].z * rstd, g.z, b.z);
        x[i].w = fmaf(x[i].w * rstd, g.w, b.w);
    }
    return make_float2(mean, rstd);
}

// fp16 copy of the normalised row (the GEMM operand); x32 only for the debug materialisation
__device__ __forceinline__ void store_row(const float4 (&x)[V4], float* __restrict__ x32, __half* __restrict__ x16,
                                          size_t row, int lane) {
    uint2* o16 = reinterpret_cast<uint2*>(x16 + row * H);
#pragma unroll
    for (int i = 0; i < V4; ++i) o16[i * 32 + lane] = make_uint2(pack_half2(x[i].x, x[i].y), pack_half2(x[i].z, x[i].w));
    if (x32 != nullptr) {
        float4* o32 = reinterpret_cast<float4*>(x32 + row * H);
#pragma unroll
        for (int i = 0; i < V4; ++i) o32[i * 32 + lane] = x[i];
    }
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
embed_ln_kernel(const int32_t* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos,
                const float* __restrict__ type0, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ y32, __half* __restrict__ x16, float2* __restrict__ stats, float* __restrict__ x32_dbg,
                int n_tokens, int S, int vocab, float eps) {
    const int lane = threadIdx.x & 31;
    const int tok = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (tok >= n_tokens) return;
    int id = ids[tok];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // submit() rejects out-of-range ids; never read OOB
    const int p = tok % S;
    const float4* w4 = reinterpret_cast<const float4*>(word + static_cast<size_t>(id) * H);
    const float4* p4 = reinterpret_cast<const float4*>(pos + static_cast<size_t>(p) * H);
    const float4* t4 = reinterpret_cast<const float4*>(type0);
    float4 x[V4];
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const float4 a = __ldg(w4 + i * 32 + lane);
        const float4 b = __ldg(p4 + i * 32 + lane);
        const float4 c = __ldg(t4 + i * 32 + lane);
        // HF order: (word + token_type) + position
        x[i] = make_float4((a.x + c.x) + b.x, (a.y + c.y) + b.y, (a.z + c.z) + b.z, (a.w + c.w) + b.w);
    }
    // the residual stream is kept PRE-LayerNorm (y32) plus per-row (mean, rstd); consumers re-apply the affine
    float4* y4 = reinterpret_cast<float4*>(y32 + static_cast<size_t>(tok) * H);
#pragma unroll
    for (int i = 0; i < V4; ++i) y4[i * 32 + lane] = x[i];
    const float2 st = ln_inplace(x, gamma, beta, eps, lane);
    if (lane == 0) stats[tok] = st;
    store_row(x, x32_dbg, x16, tok, lane);
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
ln_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
          __half* __restrict__ x16, float2* __restrict__ stats, float* __restrict__ x32_dbg, int n_rows, float eps) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float4* y4 = reinterpret_cast<const float4*>(y + static_cast<size_t>(row) * H);
    float4 x[V4];
#pragma unroll
    for (int i = 0; i < V4; ++i) x[i] = y4[i * 32 + lane];
    const float2 st = ln_inplace(x, gamma, beta, eps, lane);
    if (lane == 0) stats[row] = st;
    store_row(x, x32_dbg, x16, row, lane);
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
pool_normalize_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                      float* __restrict__ out, int n_items, int S, float eps) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const float4* y4 = reinterpret_cast<const float4*>(y + static_cast<size_t>(item) * S * H);  // CLS row
    float4 x[V4];
#pragma unroll
    for (int i = 0; i < V4; ++i) x[i] = y4[i * 32 + lane];
    ln_inplace(x, gamma, beta, eps, lane);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
    const float inv = 1.0f / fmaxf(sqrtf(warp_sum(q)), 1e-12f);  // torch.nn.functional.normalize eps
    float4* o4 = reinterpret_cast<float4*>(out + static_cast<size_t>(item) * H);  // may be peer memory
#pragma unroll
    for (int i = 0; i < V4; ++i)
        o4[i * 32 + lane] = make_float4(x[i].x * inv, x[i].y * inv, x[i].z * inv, x[i].w * inv);
}

__global__ void __launch_bounds__(256)
f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n4) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
    }
}

// blockIdx.y = shard.  Copies ids[item_begin*S .. (item_begin+item_count)*S) and the matching lens to the
// shard's input slot; dst pointers of remote shards are peer-mapped, so the stores cross NVLink directly.
__global__ void __launch_bounds__(256)
scatter_kernel(const int32_t* __restrict__ src_ids, const int32_t* __restrict__ src_lens, const ScatterPlan plan) {
    const int sh = blockIdx.y;
    const int cnt = plan.item_count[sh];
    if (cnt <= 0) return;
    const int S = plan.S;
    const size_t n = static_cast<size_t>(cnt) * S;
    const int32_t* s = src_ids + static_cast<size_t>(plan.item_begin[sh]) * S;
    int32_t* d = plan.dst_ids[sh];
    const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    if ((((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0)) {
        const size_t n4 = n / 4;
        for (size_t i = tid; i < n4; i += stride) reinterpret_cast<int4*>(d)[i] = reinterpret_cast<const int4*>(s)[i];
        for (size_t i = n4 * 4 + tid; i < n; i += stride) d[i] = s[i];
    } else {
        for (size_t i = tid; i < n; i += stride) d[i] = s[i];
    }
    for (size_t i = tid; i < static_cast<size_t>(cnt); i += stride) plan.dst_lens[sh][i] = src_lens[plan.item_begin[sh] + i];
}

}  // namespace rw

cudaError_t launch_embed_ln(const int32_t* ids, const float* word, const float* pos, const float* type0,
                            const float* gamma, const float* beta, float* y32, __half* x16, float2* stats, float* x32_dbg,
                            int n_tokens, int S, int vocab, float eps, cudaStream_t stream) {
    const int grid = (n_tokens + rw::WARPS_PER_BLOCK - 1) / rw::WARPS_PER_BLOCK;
    rw::embed_ln_kernel<<<grid, rw::WARPS_PER_BLOCK * 32, 0, stream>>>(ids, word, pos, type0, gamma, beta, y32, x16, stats,
                                                                       x32_dbg, n_tokens, S, vocab, eps);
    return cudaGetLastError();
}

cudaError_t launch_ln(const float* y, const float* gamma, const float* beta, __half* x16, float2* stats, float* x32_dbg,
                      int n_rows, float eps, cudaStream_t stream) {
    const int grid = (n_rows + rw::WARPS_PER_BLOCK - 1) / rw::WARPS_PER_BLOCK;
    rw::ln_kernel<<<grid, rw::WARPS_PER_BLOCK * 32, 0, stream>>>(y, gamma, beta, x16, stats, x32_dbg, n_rows, eps);
    return cudaGetLastError();
}

cudaError_t launch_pool_normalize(const float* y, const float* gamma, const float* beta, float* out, int n_items,
                                  int S, float eps, cudaStream_t stream) {
    const int grid = (n_items + rw::WARPS_PER_BLOCK - 1) / rw::WARPS_PER_BLOCK;
    rw::pool_normalize_kernel<<<grid, rw::WARPS_PER_BLOCK * 32, 0, stream>>>(y, gamma, beta, out, n_items, S, eps);
    return cudaGetLastError();
}

cudaError_t launch_f32_to_f16(const float* src, __half* dst, size_t n, cudaStream_t stream) {
    if (n % 4 != 0) return cudaErrorInvalidValue;
    rw::f32_to_f16_kernel<<<148 * 8, 256, 0, stream>>>(src, dst, n / 4);
    return cudaGetLastError();
}

cudaError_t launch_scatter(const int32_t* src_ids, const int32_t* src_lens, const ScatterPlan& plan,
                           cudaStream_t stream) {
    if (plan.n_shards < 1 || plan.n_shards > ScatterPlan::MAX_SHARDS) return cudaErrorInvalidValue;
    int max_cnt = 0;
    for (int i = 0; i < plan.n_shards; ++i) max_cnt = plan.item_count[i] > max_cnt ? plan.item_count[i] : max_cnt;
    if (max_cnt == 0) return cudaSuccess;
    const size_t n4 = (static_cast<size_t>(max_cnt) * plan.S + 3) / 4;
    int gx = static_cast<int>((n4 + 255) / 256);
    gx = gx < 1 ? 1 : (gx > 148 * 4 ? 148 * 4 : gx);
    rw::scatter_kernel<<<dim3(gx, plan.n_shards), 256, 0, stream>>>(src_ids, src_lens, plan);
    return cudaGetLastError();
}

}  // namespace b200
