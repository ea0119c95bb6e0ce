// HBM-bound row-wise kernels (sm_100a): embedding gather into the split residual stream, CLS pooling + final LayerNorm +
// L2 normalise with the store aimed at the root GPU's gather buffer, the root-side scatter of token ids into (peer) shard
// inputs, the load-time weight preparation (fp32 -> fp16, LayerNorm gamma folded into weight columns), and a debug-only
// LayerNorm materialisation.
//
// Row kernels: one warp per 768-wide row, 128-bit accesses, statistics by warp shuffle in fp32, two-pass variance on
// register-resident data (no E[x^2]-E[x]^2 cancellation).
//
// The residual stream is stored PRE-LayerNorm as y = hi + lo (fp16 + fp16) together with STAT_PARTS (sum, M2) partials per
// row; there is no LayerNorm kernel in the product path -- the consuming GEMM folds it into its epilogue (kernels.h).
//
// Restates BertEmbeddings.forward (HF modeling_bert.py:72-111), the LayerNorm halves of
// BertSelfOutput :287-298 / BertOutput :345-356, and sentence-transformers' CLS pooling + Normalize
// (reference 06_gpu_and_ml/gpu_snapshot.py:58, `normalize_embeddings=True`).
#include "kernels.h"
#include "ptx.cuh"

namespace b200 {
namespace rw {

constexpr int H = HIDDEN;          // 768
constexpr int V4 = H / 4 / 32;     // float4 per lane = 6: float4 i of lane l holds columns [128 i + 4 l, +4)
constexpr int WARPS_PER_BLOCK = 8;
static_assert(V4 == STAT_PARTS, "float4 i of every lane covers exactly the i-th 128-column slice");

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

// y = hi + lo of one row -> 24 fp32 values per lane (same column mapping as a float4 row)
__device__ __forceinline__ void load_split_row(const __half* __restrict__ yhi, const __half* __restrict__ ylo, size_t row, int lane,
                                               float4 (&x)[V4]) {
    const uint2* h2 = reinterpret_cast<const uint2*>(yhi + row * H);
    const uint2* l2 = reinterpret_cast<const uint2*>(ylo + row * H);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const uint2 hv = h2[i * 32 + lane], lv = l2[i * 32 + lane];
        const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&hv.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&hv.y));
        const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&lv.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&lv.y));
        x[i] = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
    }
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
embed_kernel(const int32_t* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos,
             const float* __restrict__ type0, __half* __restrict__ yhi, __half* __restrict__ ylo, float2* __restrict__ stats,
             int n_tokens, int S, int vocab) {
    const int lane = threadIdx.x & 31;
    const int tok = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (tok >= n_tokens) return;
    int id = ids[tok];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // submit() rejects out-of-range ids; never read OOB
    const int p = tok % S;
    const float4* w4 = reinterpret_cast<const float4*>(word + static_cast<size_t>(id) * H);
    const float4* p4 = reinterpret_cast<const float4*>(pos + static_cast<size_t>(p) * H);
    const float4* t4 = reinterpret_cast<const float4*>(type0);
    uint2* oh = reinterpret_cast<uint2*>(yhi + static_cast<size_t>(tok) * H);
    uint2* ol = reinterpret_cast<uint2*>(ylo + static_cast<size_t>(tok) * H);
    float ps = 0.f, pq = 0.f;  // lane i < STAT_PARTS keeps partial i
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const float4 a = __ldg(w4 + i * 32 + lane);
        const float4 b = __ldg(p4 + i * 32 + lane);
        const float4 c = __ldg(t4 + i * 32 + lane);
        // HF order: (word + token_type) + position
        const float4 x = make_float4((a.x + c.x) + b.x, (a.y + c.y) + b.y, (a.z + c.z) + b.z, (a.w + c.w) + b.w);
        // the residual stream is kept PRE-LayerNorm as hi + lo; consumers fold / re-apply the LayerNorm
        const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        oh[i * 32 + lane] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        ol[i * 32 + lane] = make_uint2(pack_half2(x.x - f0.x, x.y - f0.y), pack_half2(x.z - f1.x, x.w - f1.y));
        // (sum, M2 about its own mean) of this 128-column slice: float4 i of the 32 lanes is exactly slice i
        const float s = warp_sum((x.x + x.y) + (x.z + x.w));
        const float m = s * (1.0f / 128.0f);
        const float dx = x.x - m, dy = x.y - m, dz = x.z - m, dw = x.w - m;
        const float q = warp_sum((dx * dx + dy * dy) + (dz * dz + dw * dw));
        if (lane == i) { ps = s; pq = q; }
    }
    if (lane < STAT_PARTS) stats[static_cast<size_t>(tok) * STAT_PARTS + lane] = make_float2(ps, pq);
}

// debug only (b200rt_debug_hidden): materialise LN(hi + lo) in fp32
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
ln_materialize_kernel(const __half* __restrict__ yhi, const __half* __restrict__ ylo, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float* __restrict__ x32, int n_rows, float eps) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    float4 x[V4];
    load_split_row(yhi, ylo, row, lane, x);
    ln_inplace(x, gamma, beta, eps, lane);
    float4* o32 = reinterpret_cast<float4*>(x32 + static_cast<size_t>(row) * H);
#pragma unroll
    for (int i = 0; i < V4; ++i) o32[i * 32 + lane] = x[i];
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
pool_normalize_kernel(const __half* __restrict__ yhi, const __half* __restrict__ ylo, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float* __restrict__ out, int n_items, int S, float eps) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (item >= n_items) return;
    float4 x[V4];
    load_split_row(yhi, ylo, static_cast<size_t>(item) * S, lane, x);  // CLS row
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

// ----------------------------------------------------------------------------------------------- ViT image tower (CLIP)
// Restates CLIPVisionEmbeddings.forward / the pooled head of CLIPVisionTransformer + visual_projection (HF modeling_clip.py)
// for the reference's image-embedding example (06_gpu_and_ml/embeddings/image_embeddings_infinity.py:76-77, 298-306).

// pixels fp32 [B, 3, img, img] -> A fp16 [B * T, 3 p p], T = (img/p)^2 + 1: row b*T is zero (the class token takes no patch),
// row b*T + 1 + (py*grid + px) holds patch (py, px) flattened (channel, row, column) -- the stride-p convolution as a GEMM.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
im2col_kernel(const float* __restrict__ pixels, __half* __restrict__ a, int n_rows, int img, int p, int grid) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int T = grid * grid + 1, pd = 3 * p * p;
    const int b = row / T, tk = row % T;
    __half2* o = reinterpret_cast<__half2*>(a + static_cast<size_t>(row) * pd);
    if (tk == 0) {
        for (int e = lane; e < pd / 2; e += 32) o[e] = __floats2half2_rn(0.f, 0.f);
        return;
    }
    const int py = (tk - 1) / grid, px = (tk - 1) % grid;
    const float* base = pixels + static_cast<size_t>(b) * 3 * img * img + static_cast<size_t>(py * p) * img + px * p;
    for (int e = lane * 2; e < pd; e += 64) {
        const int c = e / (p * p), ky = (e / p) % p, kx = e % p;  // kx even: the pair stays inside one patch row
        const float2 v = *reinterpret_cast<const float2*>(base + (static_cast<size_t>(c) * img + ky) * img + kx);
        o[e / 2] = __floats2half2_rn(v.x, v.y);
    }
}

// row b*T + tk: e = (tk == 0 ? class_embedding : patch GEMM output) + position[tk]; h = pre_layrnorm(e) becomes the residual
// stream (hi + lo) together with its (sum, M2) partials -- the statistics LayerNorm1 of the first layer folds.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
vit_embed_kernel(const __half* __restrict__ patch_out, const float* __restrict__ cls, const float* __restrict__ pos,
                 const float* __restrict__ gamma, const float* __restrict__ beta, __half* __restrict__ yhi, __half* __restrict__ ylo,
                 float2* __restrict__ stats, int n_rows, int T, float eps) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int tk = row % T;
    const uint2* p2 = reinterpret_cast<const uint2*>(patch_out + static_cast<size_t>(row) * H);
    const float4* c4 = reinterpret_cast<const float4*>(cls);
    const float4* q4 = reinterpret_cast<const float4*>(pos + static_cast<size_t>(tk) * H);
    float4 x[V4];
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        float4 v;
        if (tk == 0) {
            v = __ldg(c4 + i * 32 + lane);
        } else {
            const uint2 hv = p2[i * 32 + lane];
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hv.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&hv.y));
            v = make_float4(a.x, a.y, b.x, b.y);
        }
        const float4 q = __ldg(q4 + i * 32 + lane);
        x[i] = make_float4(v.x + q.x, v.y + q.y, v.z + q.z, v.w + q.w);
    }
    ln_inplace(x, gamma, beta, eps, lane);
    uint2* oh = reinterpret_cast<uint2*>(yhi + static_cast<size_t>(row) * H);
    uint2* ol = reinterpret_cast<uint2*>(ylo + static_cast<size_t>(row) * H);
    float ps = 0.f, pq = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const float4 v = x[i];
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        oh[i * 32 + lane] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        ol[i * 32 + lane] = make_uint2(pack_half2(v.x - f0.x, v.y - f0.y), pack_half2(v.z - f1.x, v.w - f1.y));
        const float sm = warp_sum((v.x + v.y) + (v.z + v.w));
        const float m = sm * (1.0f / 128.0f);
        const float dx = v.x - m, dy = v.y - m, dz = v.z - m, dw = v.w - m;
        const float q = warp_sum((dx * dx + dy * dy) + (dz * dz + dw * dw));
        if (lane == i) { ps = sm; pq = q; }
    }
    if (lane < STAT_PARTS) stats[static_cast<size_t>(row) * STAT_PARTS + lane] = make_float2(ps, pq);
}

// One block per image: post_layernorm of the class-token row, visual projection [P, H], L2 normalise; row i is stored at
// out + i * P, possibly peer memory (the fused gather).
constexpr int VIT_POOL_MAX_P = 1024;
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
vit_pool_kernel(const __half* __restrict__ yhi, const __half* __restrict__ ylo, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ proj, float* __restrict__ out, int T, int P, float eps) {
    __shared__ float s_out[VIT_POOL_MAX_P];
    __shared__ float s_red[WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x;
    float4 x[V4];
    load_split_row(yhi, ylo, static_cast<size_t>(item) * T, lane, x);  // class-token row (every warp keeps its own copy)
    ln_inplace(x, gamma, beta, eps, lane);
    float sq = 0.f;
    for (int n = warp; n < P; n += WARPS_PER_BLOCK) {
        const float4* w4 = reinterpret_cast<const float4*>(proj + static_cast<size_t>(n) * H);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < V4; ++i) {
            const float4 w = __ldg(w4 + i * 32 + lane);
            acc = fmaf(x[i].x, w.x, fmaf(x[i].y, w.y, fmaf(x[i].z, w.z, fmaf(x[i].w, w.w, acc))));
        }
        acc = warp_sum(acc);
        if (lane == 0) s_out[n] = acc;
        sq = fmaf(acc, acc, sq);  // identical on every lane after warp_sum
    }
    if (lane == 0) s_red[warp] = sq;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS_PER_BLOCK; ++w) tot += s_red[w];
    const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
    float* o = out + static_cast<size_t>(item) * P;
    for (int n = threadIdx.x; n < P; n += WARPS_PER_BLOCK * 32) o[n] = s_out[n] * inv;
}

// Load-time weight preparation, one warp per output row n of W[N,K] (see kernels.h, "LNfold"):
//   gamma given: w_out[n,k] = fp16(w[n,k] gamma[k] - m_n),  m_n = mean_k(w[n,k] gamma[k])   (rows centred: the LayerNorm's mean
//                subtraction then happens inside the GEMM),  cvec[n] = sum_k w[n,k] beta[k] + bias[n]
//   gamma NULL : w_out[n,k] = fp16(w[n,k])
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
fold_ln_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ bias, __half* __restrict__ w_out, float* __restrict__ cvec, int N, int K) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (n >= N) return;
    const float* wr = w + static_cast<size_t>(n) * K;
    __half* orow = w_out + static_cast<size_t>(n) * K;
    float m = 0.f;
    if (gamma != nullptr) {
        float s = 0.f, c = 0.f;
        for (int k = lane * 2; k < K; k += 64) {
            const float2 wv = *reinterpret_cast<const float2*>(wr + k);
            s += fmaf(wv.x, __ldg(gamma + k), wv.y * __ldg(gamma + k + 1));
            c = fmaf(wv.x, __ldg(beta + k), fmaf(wv.y, __ldg(beta + k + 1), c));
        }
        m = warp_sum(s) / static_cast<float>(K);
        c = warp_sum(c);
        if (lane == 0 && cvec != nullptr) cvec[n] = c + (bias != nullptr ? bias[n] : 0.f);
    }
    for (int k = lane * 2; k < K; k += 64) {
        const float2 wv = *reinterpret_cast<const float2*>(wr + k);
        const float g0 = gamma != nullptr ? __ldg(gamma + k) : 1.f, g1 = gamma != nullptr ? __ldg(gamma + k + 1) : 1.f;
        *reinterpret_cast<__half2*>(orow + k) = __floats2half2_rn(fmaf(wv.x, g0, -m), fmaf(wv.y, g1, -m));
    }
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

cudaError_t launch_embed(const int32_t* ids, const float* word, const float* pos, const float* type0, __half* yhi, __half* ylo,
                         float2* stats, int n_tokens, int S, int vocab, cudaStream_t stream) {
    const int grid = (n_tokens + rw::WARPS_PER_BLOCK - 1) / rw::WARPS_PER_BLOCK;
    rw::embed_kernel<<<grid, rw::WARPS_PER_BLOCK * 32, 0, stream>>>(ids, word, pos, type0, yhi, ylo, stats, n_tokens, S, vocab);
    return cudaGetLastError();
}

cudaError_t launch_ln_materialize(const __half* yhi, const __half* ylo, const float* gamma, const float* beta, float* x32,
                                  int n_rows, float eps, cudaStream_t stream) {
    const int grid = (n_rows + rw::WARPS_PER_BLOCK - 1) / rw::WARPS_PER_BLOCK;
    rw::ln_materialize_kernel<<<grid, rw::WARPS_PER_BLOCK * 32, 0, stream>>>(yhi, ylo, gamma, beta, x32, n_rows, eps);
    return cudaGetLastError();
}

cudaError_t launch_pool_normalize(const __half* yhi, const __half* ylo, const float* gamma, const float* beta, float* out,
                                  int n_items, int S, float eps, cudaStream_t stream) {
    const int grid = (n_items + rw::WARPS_PER_BLOCK - 1) / rw::WARPS_PER_BLOCK;
    rw::pool_normalize_kernel<<<grid, rw::WARPS_PER_BLOCK * 32, 0, stream>>>(yhi, ylo, gamma, beta, out, n_items, S, eps);
    return cudaGetLastError();
}

cudaError_t launch_im2col(const float* pixels, __half* a, int n_items, int img, int p, cudaStream_t stream) {
    const int grid_sz = img / p, T = grid_sz * grid_sz + 1;
    if (img % p != 0 || (3 * p * p) % 64 != 0 || p % 2 != 0) return cudaErrorInvalidValue;
    const int rows = n_items * T;
    rw::im2col_kernel<<<(rows + rw::WARPS_PER_BLOCK - 1) / rw::WARPS_PER_BLOCK, rw::WARPS_PER_BLOCK * 32, 0, stream>>>(pixels, a, rows, img, p, grid_sz);
    return cudaGetLastError();
}

cudaError_t launch_vit_embed(const __half* patch_out, const float* cls, const float* pos, const float* gamma, const float* beta,
                             __half* yhi, __half* ylo, float2* stats, int n_rows, int T, float eps, cudaStream_t stream) {
    rw::vit_embed_kernel<<<(n_rows + rw::WARPS_PER_BLOCK - 1) / rw::WARPS_PER_BLOCK, rw::WARPS_PER_BLOCK * 32, 0, stream>>>(
        patch_out, cls, pos, gamma, beta, yhi, ylo, stats, n_rows, T, eps);
    return cudaGetLastError();
}

cudaError_t launch_vit_pool(const __half* yhi, const __half* ylo, const float* gamma, const float* beta, const float* proj, float* out,
                            int n_items, int T, int P, float eps, cudaStream_t stream) {
    if (P < 1 || P > rw::VIT_POOL_MAX_P) return cudaErrorInvalidValue;
    rw::vit_pool_kernel<<<n_items, rw::WARPS_PER_BLOCK * 32, 0, stream>>>(yhi, ylo, gamma, beta, proj, out, T, P, eps);
    return cudaGetLastError();
}

cudaError_t launch_fold_ln(const float* w, const float* gamma, const float* beta, const float* bias, __half* w_out, float* cvec,
                           int N, int K, cudaStream_t stream) {
    if (K % 64 != 0 || N < 1) return cudaErrorInvalidValue;
    const int grid = (N + rw::WARPS_PER_BLOCK - 1) / rw::WARPS_PER_BLOCK;
    rw::fold_ln_kernel<<<grid, rw::WARPS_PER_BLOCK * 32, 0, stream>>>(w, gamma, beta, bias, w_out, cvec, N, K);
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
