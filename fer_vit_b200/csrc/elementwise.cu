// Small HBM-bound helpers around the GEMMs: casts / weight caches, im2col, cls rows, token gathers,
// deterministic column sums, split-K reduction, adapter backward glue.
#include "common.cuh"
#include "kernels.h"

namespace fervit {

namespace ew {

// ---- fp32 -> activation dtype cast (w+ latents into the A operand of the token projection) ----
template <typename AT>
__global__ void cast_kernel(const float* __restrict__ src, AT* __restrict__ dst, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(src + i * 4);
    store4<AT>(dst + i * 4, v);
  }
}

constexpr int WC_BATCH = 32;

// ---- fp32 [R,C] -> bf16 [R,C] and/or bf16 [C,R] (weight caches: W for forward, W^T for dgrad) ----
__global__ void weight_cache_kernel(const float* __restrict__ src, int R, int C, bf16* __restrict__ dst,
                                    bf16* __restrict__ dst_t) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      v = src[(size_t)r * C + c];
      if (dst) dst[(size_t)r * C + c] = __float2bfloat16_rn(v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (dst_t) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (r < R && c < C) dst_t[(size_t)c * R + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
  }
}

// ---- the same for up to WC_BATCH matrices in one launch (the trainable weights are re-cast every step) ----
struct WcBatch {
  const float* src[WC_BATCH];
  bf16* dst[WC_BATCH];
  bf16* dst_t[WC_BATCH];
  int R[WC_BATCH], C[WC_BATCH];
  int tile0[WC_BATCH + 1];  // first 64x64 tile of each matrix in the flat grid
  int n;
};
// 64 x 64 tiles, 256 threads: float4 loads, 8-byte row-major stores, and the transposed copy written as bf16 pairs
// along r (one warp = 128 contiguous bytes of one transposed row). Matrices whose shape or alignment does not allow the
// vector forms take the element-wise path of the same tile.
__global__ void __launch_bounds__(256) weight_cache_batch_kernel(const __grid_constant__ WcBatch b) {
  __shared__ float tile[64][65];
  int m = 0;
  while (m + 1 < b.n && (int)blockIdx.x >= b.tile0[m + 1]) ++m;
  const int R = b.R[m], C = b.C[m];
  const int t = blockIdx.x - b.tile0[m];
  const int tiles_c = (C + 63) / 64;
  const int c0 = (t % tiles_c) * 64, r0 = (t / tiles_c) * 64;
  const float* __restrict__ src = b.src[m];
  bf16* __restrict__ dst = b.dst[m];
  bf16* __restrict__ dst_t = b.dst_t[m];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                   (dst == nullptr || (reinterpret_cast<uintptr_t>(dst) & 7) == 0);
  if (vec) {
    const int cq = (tid & 15) * 4;           // 16 threads cover the 64 columns of a row
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = i * 16 + (tid >> 4);
      const int r = r0 + rr, c = c0 + cq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < R && c < C) {
        v = *reinterpret_cast<const float4*>(src + (size_t)r * C + c);
        if (dst) {
          uint2 o;
          o.x = pack_bf16x2(v.x, v.y);
          o.y = pack_bf16x2(v.z, v.w);
          *reinterpret_cast<uint2*>(dst + (size_t)r * C + c) = o;
        }
      }
      tile[rr][cq] = v.x; tile[rr][cq + 1] = v.y; tile[rr][cq + 2] = v.z; tile[rr][cq + 3] = v.w;
    }
  } else {
    for (int i = tid; i < 64 * 64; i += 256) {
      const int rr = i >> 6, cc = i & 63;
      const int r = r0 + rr, c = c0 + cc;
      float v = 0.f;
      if (r < R && c < C) {
        v = src[(size_t)r * C + c];
        if (dst) dst[(size_t)r * C + c] = __float2bfloat16_rn(v);
      }
      tile[rr][cc] = v;
    }
  }
  __syncthreads();
  if (!dst_t) return;
  if ((R % 2 == 0) && ((reinterpret_cast<uintptr_t>(dst_t) & 3) == 0)) {
    const int rr = threadIdx.x * 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int cc = i * 8 + threadIdx.y;
      const int c = c0 + cc, r = r0 + rr;
      if (c < C && r < R)   // R even, r even: r + 1 < R
        *reinterpret_cast<uint32_t*>(dst_t + (size_t)c * R + r) = pack_bf16x2(tile[rr][cc], tile[rr + 1][cc]);
    }
  } else {
    for (int i = tid; i < 64 * 64; i += 256) {
      const int cc = i >> 6, rr = i & 63;
      const int c = c0 + cc, r = r0 + rr;
      if (c < C && r < R) dst_t[(size_t)c * R + r] = __float2bfloat16_rn(tile[rr][cc]);
    }
  }
}

// ---- LayerNorm folded into the weight that follows it (frozen norm + frozen nn.Linear, bf16 mode) ----
//   W'[j,k] = W[j,k] * gamma[k]  (bf16, also transposed),  b'[j] = b[j] + sum_k W[j,k] beta[k]  (fp32, unrounded W),
//   cs[j] = sum_k float(bf16(W'[j,k]))  — the column sum the consumer's epilogue multiplies the row's mean shift by,
//   taken over exactly the values the tensor core sees.
// One warp per output row j; fixed summation order (deterministic). Runs when a frozen weight (re)enters the cache.
__global__ void fold_ln_weight_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, int R, int C,
                                      bf16* __restrict__ dst, bf16* __restrict__ dst_t, float* __restrict__ bfold,
                                      float* __restrict__ cs) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= R) return;
  float sb = 0.f, sc = 0.f;
  for (int k = lane; k < C; k += 32) {
    const float w = W[(size_t)j * C + k];
    const bf16 wf = __float2bfloat16_rn(w * gamma[k]);
    dst[(size_t)j * C + k] = wf;
    dst_t[(size_t)k * R + j] = wf;
    sb = fmaf(w, beta[k], sb);
    sc += __bfloat162float(wf);
  }
  sb = warp_sum(sb);
  sc = warp_sum(sc);
  if (lane == 0) {
    bfold[j] = (b ? b[j] : 0.f) + sb;
    cs[j] = sc;
  }
}

// ---- im2col for stride == kernel patchify (image_vit.py:27-43): x [B,C,H,W] -> A [B*L, C*P*P] ----
template <typename AT>
__global__ void im2col_kernel(const float* __restrict__ x, AT* __restrict__ out, int B, int C, int H, int W,
                              int P) {
  const int gw = W / P, gh = H / P;
  const int Kd = C * P * P;
  const size_t total4 = (size_t)B * gh * gw * Kd / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 4;
    const int k = (int)(e % Kd);
    const size_t tok = e / Kd;
    const int px = (int)(tok % gw);
    const int py = (int)((tok / gw) % gh);
    const int b = (int)(tok / ((size_t)gw * gh));
    const int j = k % P, ii = (k / P) % P, c = k / (P * P);
    const float4 v = *reinterpret_cast<const float4*>(x + (((size_t)b * C + c) * H + (py * P + ii)) * W + px * P + j);
    store4<AT>(out + e, v);
  }
}

// ---- x0[b,0,:] = cls + pos[0] (latent_vit.py:41-44, hybrid_latent_vit.py:218-222, image_vit.py:151-155) ----
template <typename AT>
__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x0,
                                AT* __restrict__ x0_at, int B, int S, int E, Dropout drop) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    float v = cls[c] + pos[c];
    if (drop.threshold) {
      const uint64_t idx = (uint64_t)b * S * E + c;
      v = drop_keep(drop.eff(), drop.site, idx, drop.threshold) ? v * drop.scale : 0.f;
    }
    x0[(size_t)b * S * E + c] = v;
    if (x0_at) x0_at[(size_t)b * S * E + c] = from_f32<AT>(v);
  }
}

// ---- input dropout over the token rows of x0 (ImageViT, image_vit.py:156); cls rows handled above ----
template <typename AT>
__global__ void token_dropout_kernel(float* __restrict__ x0, AT* __restrict__ x0_at, int B, int S, int E, Dropout drop) {
  const size_t total = (size_t)B * S * E;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)((i / E) % S);
    if (s == 0) continue;
    const float v = drop_keep(drop.eff(), drop.site, i, drop.threshold) ? x0[i] * drop.scale : 0.f;
    x0[i] = v;
    if (x0_at) x0_at[i] = from_f32<AT>(v);
  }
}

// ---- dx0 [B,S,E] fp32 rows 1.. -> compact [B*L, E] activation dtype (A operand of the input wgrad) ----
template <typename AT>
__global__ void gather_tokens_kernel(const float* __restrict__ dx0, AT* __restrict__ out, int B, int L, int E,
                                     Dropout drop) {
  const size_t total4 = (size_t)B * L * E / 4;
  const int S = L + 1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 4;
    const int c = (int)(e % E);
    const size_t tok = e / E;
    const int l = (int)(tok % L);
    const size_t b = tok / L;
    const size_t src = (b * S + 1 + l) * E + c;
    float4 v = *reinterpret_cast<const float4*>(dx0 + src);
    if (drop.threshold) {
      v.x = drop_keep(drop.eff(), drop.site, src + 0, drop.threshold) ? v.x * drop.scale : 0.f;
      v.y = drop_keep(drop.eff(), drop.site, src + 1, drop.threshold) ? v.y * drop.scale : 0.f;
      v.z = drop_keep(drop.eff(), drop.site, src + 2, drop.threshold) ? v.z * drop.scale : 0.f;
      v.w = drop_keep(drop.eff(), drop.site, src + 3, drop.threshold) ? v.w * drop.scale : 0.f;
    }
    store4<AT>(out + e, v);
  }
}

// ---- deterministic column sum: in [R, C] (row stride ld) -> partial [chunks][C] -> out [C] ----
// block (32, 8): x = group of 4 columns, y = row lane; every thread keeps 8 independent 16-byte loads in flight,
// the 8 row lanes are combined through shared memory in a fixed order.
__device__ __forceinline__ float4 colsum_block_reduce(float4 acc, float4 (*red)[32]) {
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.y == 0) {
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      const float4 v = red[y][threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
  }
  return t;
}
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ in, int R, int C, long long ld, int rows_per_chunk,
                      float* __restrict__ partial, Dropout drop) {
  __shared__ float4 red[8][32];
  pdl_trigger();
  pdl_grid_sync();
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(R, r0 + rows_per_chunk);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
#pragma unroll 8
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      float4 v = load4<T>(in + (size_t)r * ld + c);
      if (drop.threshold) {
        const uint64_t base = (uint64_t)r * ld + c, seed = drop.eff();
        v.x = drop_keep(seed, drop.site, base + 0, drop.threshold) ? v.x * drop.scale : 0.f;
        v.y = drop_keep(seed, drop.site, base + 1, drop.threshold) ? v.y * drop.scale : 0.f;
        v.z = drop_keep(seed, drop.site, base + 2, drop.threshold) ? v.z * drop.scale : 0.f;
        v.w = drop_keep(seed, drop.site, base + 3, drop.threshold) ? v.w * drop.scale : 0.f;
      }
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  const float4 t = colsum_block_reduce(acc, red);
  if (threadIdx.y == 0 && c < C) *reinterpret_cast<float4*>(partial + (size_t)blockIdx.y * C + c) = t;
}
// finishing pass, block (8, 32): one CTA per 32 columns (x = group of 4 columns), 32 row lanes walk the partial rows, so
// even 256 partial rows are 8 independent loads per thread; the row lanes are combined in a fixed order
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int chunks, int C, const float* alpha_ptr, float alpha,
                    float* __restrict__ out, float* __restrict__ out_hi, int split) {
  __shared__ float4 red[32][8];
  pdl_trigger();
  pdl_grid_sync();
  const int c = (blockIdx.x * 8 + threadIdx.x) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
#pragma unroll 8
    for (int k = threadIdx.y; k < chunks; k += 32) {
      const float4 v = *reinterpret_cast<const float4*>(partial + (size_t)k * C + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int y = 0; y < 32; ++y) {
      const float4 v = red[y][threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    if (alpha_ptr) alpha *= __ldg(alpha_ptr);
    // columns [split, C) may go to a second tensor (LayerNorm: [gamma | beta] partial rows -> two gradients)
    float* dst = (out_hi && c >= split) ? out_hi + (c - split) : out + c;
    *reinterpret_cast<float4*>(dst) = make_float4(t.x * alpha, t.y * alpha, t.z * alpha, t.w * alpha);
  }
}

// ---- split-K reduction: out[i] = alpha * sum_s partial[s][i] ----
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, size_t n4, const float* alpha_ptr,
                                     float alpha, float* __restrict__ out) {
  pdl_trigger();
  pdl_grid_sync();
  if (alpha_ptr) alpha *= __ldg(alpha_ptr);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int s = 0; s < splits; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)s * n4 + i) * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x *= alpha; acc.y *= alpha; acc.z *= alpha; acc.w *= alpha;
    *reinterpret_cast<float4*>(out + i * 4) = acc;
  }
}

// ---- deferred gradient finalisation of the AdapterModules (hybrid_latent_vit.py:249-265) ----
// During a block's backward only PARTIAL sums are produced (split-K slabs of the two weight-gradient GEMMs, per-chunk
// column sums of dy and du). Once per backward stage group, two launches finish every block of the group:
//   dW2 = alpha * sum_s w2_part[s]          dW1 = sum_s w1_part[s]
//   db2 = alpha * colsum(dy)                db1 = colsum(du)
//   dalpha = <W2, sum_s w2_part[s]> + b2 . colsum(dy)
// (dalpha = sum dy * (W2 g + b2); sum_t dy_t^T W2 g_t = <W2, dy^T g>, so no [T, A] intermediate is needed.)
// All sums run in a fixed order: deterministic, no atomics.
struct AdFinBatch {
  AdapterGradJob job[AD_FIN_MAX];
  int n;
};
constexpr int AD_FIN_CTAS = 24;  // CTAs per job in phase 1; the last one also does the bias / column-sum part

__global__ void __launch_bounds__(256)
adapter_grad_phase1_kernel(const __grid_constant__ AdFinBatch b, float* __restrict__ dalpha_part) {
  __shared__ float red[8];
  const AdapterGradJob& j = b.job[blockIdx.y];
  const float alpha = __ldg(j.alpha_ptr);
  const int n4 = j.E * j.A / 4;
  float dot = 0.f;
  // dW2 [E, A] and dW1 [A, E]: same element count
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 a2 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a2;
#pragma unroll 4
    for (int s = 0; s < j.s2; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(j.w2_part + ((size_t)s * n4 + i) * 4);
      a2.x += v.x; a2.y += v.y; a2.z += v.z; a2.w += v.w;
    }
#pragma unroll 4
    for (int s = 0; s < j.s1; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(j.w1_part + ((size_t)s * n4 + i) * 4);
      a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
    }
    const float4 w = __ldg(reinterpret_cast<const float4*>(j.W2) + i);
    dot += (w.x * a2.x + w.y * a2.y) + (w.z * a2.z + w.w * a2.w);
    *reinterpret_cast<float4*>(j.dW2 + (size_t)i * 4) = make_float4(alpha * a2.x, alpha * a2.y, alpha * a2.z, alpha * a2.w);
    *reinterpret_cast<float4*>(j.dW1 + (size_t)i * 4) = a1;
  }
  {
    // column sums, spread over the job's CTAs: this CTA owns 32 columns; thread (c, k0) adds chunks k0, k0+8, ...
    // and the eight partial sums of a column are combined in a fixed order.
    //   db2 = alpha * colsum(dy) [E], db1 = colsum(du) [A]; dalpha += b2 . colsum(dy)
    __shared__ float cred[8][33];
    const int c = threadIdx.x & 31, k0 = threadIdx.x >> 5;
    for (int pass = 0; pass < 2; ++pass) {
      const int C = pass == 0 ? j.E : j.A;
      const float* __restrict__ src = pass == 0 ? j.cs_dy : j.cs_du;
      for (int c0 = blockIdx.x * 32; c0 < C; c0 += gridDim.x * 32) {
        const int col = c0 + c;
        float cs = 0.f;
        if (col < C)
          for (int k = k0; k < j.chunks; k += 8) cs += src[(size_t)k * C + col];
        cred[k0][c] = cs;
        __syncthreads();
        if (k0 == 0 && col < C) {
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) t += cred[w][c];
          if (pass == 0) {
            dot += __ldg(j.b2 + col) * t;
            j.db2[col] = alpha * t;
          } else {
            j.db1[col] = t;
          }
        }
        __syncthreads();
      }
    }
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    dalpha_part[blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void adapter_grad_phase2_kernel(const __grid_constant__ AdFinBatch b, const float* __restrict__ dalpha_part,
                                           int parts) {
  const int jb = blockIdx.x * blockDim.x + threadIdx.x;
  if (jb >= b.n) return;
  float t = 0.f;
  for (int i = 0; i < parts; ++i) t += dalpha_part[jb * parts + i];
  b.job[jb].dalpha[0] = t;
}

__global__ void dropout_mask_kernel(float* __restrict__ out, size_t n, Dropout drop) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = drop_keep(drop.eff(), drop.site, i, drop.threshold) ? drop.scale : 0.f;
}

__global__ void fill_kernel(float* __restrict__ out, size_t n, float v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = v;
}

static inline int grid_for(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = (size_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace ew

template <typename AT>
int cast_to_act(const float* src, AT* dst, size_t n, cudaStream_t stream) {
  FV_CHECK(n % 4 == 0, "cast_to_act: element count must be a multiple of 4");
  if (n == 0) return 0;
  ew::cast_kernel<AT><<<ew::grid_for(n / 4, 256), 256, 0, stream>>>(src, dst, n / 4);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template int cast_to_act<float>(const float*, float*, size_t, cudaStream_t);
template int cast_to_act<bf16>(const float*, bf16*, size_t, cudaStream_t);

int weight_cache(const float* src, int R, int C, bf16* dst, bf16* dst_t, cudaStream_t stream) {
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32)), block(32, 8);
  ew::weight_cache_kernel<<<grid, block, 0, stream>>>(src, R, C, dst, dst_t);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

int weight_cache_batch(const float* const* src, const int* R, const int* C, bf16* const* dst, bf16* const* dst_t, int n,
                       cudaStream_t stream) {
  for (int base = 0; base < n; base += ew::WC_BATCH) {
    ew::WcBatch b;
    memset(&b, 0, sizeof(b));
    b.n = (n - base < ew::WC_BATCH) ? n - base : ew::WC_BATCH;
    int tiles = 0;
    for (int i = 0; i < b.n; ++i) {
      b.src[i] = src[base + i]; b.dst[i] = dst[base + i]; b.dst_t[i] = dst_t[base + i];
      b.R[i] = R[base + i]; b.C[i] = C[base + i];
      b.tile0[i] = tiles;
      tiles += ceil_div(R[base + i], 64) * ceil_div(C[base + i], 64);
    }
    b.tile0[b.n] = tiles;
    if (tiles == 0) continue;
    ew::weight_cache_batch_kernel<<<tiles, dim3(32, 8), 0, stream>>>(b);
    FV_COUNT_LAUNCH();
    FV_LAUNCH_CHECK();
  }
  return 0;
}

int fold_ln_weight(const float* W, const float* b, const float* gamma, const float* beta, int R, int C, bf16* dst,
                   bf16* dst_t, float* bfold, float* cs, cudaStream_t stream) {
  FV_CHECK(W && gamma && beta && dst && dst_t && bfold && cs, "fold_ln_weight: null argument");
  ew::fold_ln_weight_kernel<<<ceil_div(R, 8), 256, 0, stream>>>(W, b, gamma, beta, R, C, dst, dst_t, bfold, cs);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

template <typename AT>
int im2col(const float* x, AT* out, int B, int C, int H, int W, int P, cudaStream_t stream) {
  FV_CHECK(P % 4 == 0 && H % P == 0 && W % P == 0, "im2col: patch size must divide the image and be a multiple of 4");
  const size_t n4 = (size_t)B * C * H * W / 4;
  ew::im2col_kernel<AT><<<ew::grid_for(n4, 256), 256, 0, stream>>>(x, out, B, C, H, W, P);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template int im2col<float>(const float*, float*, int, int, int, int, int, cudaStream_t);
template int im2col<bf16>(const float*, bf16*, int, int, int, int, int, cudaStream_t);

template <typename AT>
int cls_rows(const float* cls, const float* pos, float* x0, AT* x0_at, int B, int S, int E, Dropout drop,
             cudaStream_t stream) {
  ew::cls_rows_kernel<AT><<<B, 256, 0, stream>>>(cls, pos, x0, x0_at, B, S, E, drop);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template int cls_rows<float>(const float*, const float*, float*, float*, int, int, int, Dropout, cudaStream_t);
template int cls_rows<bf16>(const float*, const float*, float*, bf16*, int, int, int, Dropout, cudaStream_t);

template <typename AT>
int token_dropout(float* x0, AT* x0_at, int B, int S, int E, Dropout drop, cudaStream_t stream) {
  if (!drop.threshold) return 0;
  const size_t n = (size_t)B * S * E;
  ew::token_dropout_kernel<AT><<<ew::grid_for(n, 256), 256, 0, stream>>>(x0, x0_at, B, S, E, drop);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template int token_dropout<float>(float*, float*, int, int, int, Dropout, cudaStream_t);
template int token_dropout<bf16>(float*, bf16*, int, int, int, Dropout, cudaStream_t);

template <typename AT>
int gather_tokens(const float* dx0, AT* out, int B, int L, int E, Dropout drop, cudaStream_t stream) {
  FV_CHECK(E % 4 == 0, "gather_tokens: E must be a multiple of 4");
  const size_t n4 = (size_t)B * L * E / 4;
  ew::gather_tokens_kernel<AT><<<ew::grid_for(n4, 256), 256, 0, stream>>>(dx0, out, B, L, E, drop);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template int gather_tokens<float>(const float*, float*, int, int, int, Dropout, cudaStream_t);
template int gather_tokens<bf16>(const float*, bf16*, int, int, int, Dropout, cudaStream_t);

int colsum_chunks(int R) {
  int chunks = ceil_div(R, 64);
  if (chunks > 256) chunks = 256;
  if (chunks < 1) chunks = 1;
  return chunks;
}

// scratch: [colsum_chunks(R)][C] fp32
template <typename T>
int colsum(const T* in, int R, int C, long long ld, float* scratch, const float* alpha_ptr, float alpha, float* out,
           Dropout drop, cudaStream_t stream) {
  FV_CHECK(C % 4 == 0 && ld % 4 == 0, "colsum: column count and leading dimension must be multiples of 4");
  const int chunks = colsum_chunks(R);
  const int rpc = ceil_div(R, chunks);
  dim3 grid(ceil_div(C, 128), chunks), block(32, 8);
  FV_CUDA(launch_pdl(ew::colsum_partial_kernel<T>, grid, block, 0, stream, in, R, C, ld, rpc, scratch, drop));
  FV_COUNT_LAUNCH();
  FV_CUDA(launch_pdl(ew::colsum_final_kernel, dim3(ceil_div(C, 32)), dim3(8, 32), 0, stream, (const float*)scratch, chunks,
                     C, alpha_ptr, alpha, out, (float*)nullptr, 0));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template int colsum<float>(const float*, int, int, long long, float*, const float*, float, float*, Dropout,
                           cudaStream_t);
template int colsum<bf16>(const bf16*, int, int, long long, float*, const float*, float, float*, Dropout,
                          cudaStream_t);

// columns [0, split) -> out, [split, C) -> out_hi (when given)
int colsum_reduce_partials(const float* partial, int chunks, int C, float* out, cudaStream_t stream, float* out_hi,
                           int split) {
  FV_CHECK(C % 4 == 0 && split % 4 == 0, "colsum_reduce_partials: column counts must be multiples of 4");
  FV_CUDA(launch_pdl(ew::colsum_final_kernel, dim3(ceil_div(C, 32)), dim3(8, 32), 0, stream, partial, chunks, C,
                     (const float*)nullptr, 1.0f, out, out_hi, split));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

int splitk_reduce(const float* partial, int splits, size_t n, const float* alpha_ptr, float alpha, float* out,
                  cudaStream_t stream) {
  FV_CHECK(n % 4 == 0, "splitk_reduce: element count must be a multiple of 4");
  FV_CUDA(launch_pdl(ew::splitk_reduce_kernel, dim3(ew::grid_for(n / 4, 64)), dim3(64), 0, stream, partial, splits,
                     n / 4, alpha_ptr, alpha, out));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

// per-chunk column sums only: scratch [colsum_chunks(R)][C]; reduced later (adapter_grad_finalize)
template <typename T>
int colsum_partial(const T* in, int R, int C, long long ld, float* scratch, cudaStream_t stream) {
  FV_CHECK(C % 4 == 0 && ld % 4 == 0, "colsum_partial: column count and leading dimension must be multiples of 4");
  const int chunks = colsum_chunks(R);
  const int rpc = ceil_div(R, chunks);
  dim3 grid(ceil_div(C, 128), chunks), block(32, 8);
  FV_CUDA(launch_pdl(ew::colsum_partial_kernel<T>, grid, block, 0, stream, in, R, C, ld, rpc, scratch,
                     make_dropout(0.f, 0, 0)));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
// the same with the chunk height given: scratch [ceil(R / rows_per_chunk)][C] (grouped column sums: R = groups x rows,
// rows_per_chunk dividing the group height keeps every chunk inside one group)
template <typename T>
int colsum_partial_rows(const T* in, int R, int C, long long ld, int rows_per_chunk, float* scratch, cudaStream_t stream) {
  FV_CHECK(C % 4 == 0 && ld % 4 == 0 && rows_per_chunk > 0, "colsum_partial_rows: bad shape");
  dim3 grid(ceil_div(C, 128), ceil_div(R, rows_per_chunk)), block(32, 8);
  FV_CHECK(grid.y <= 65535, "colsum_partial_rows: too many chunks");
  FV_CUDA(launch_pdl(ew::colsum_partial_kernel<T>, grid, block, 0, stream, in, R, C, ld, rows_per_chunk, scratch,
                     make_dropout(0.f, 0, 0)));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template int colsum_partial_rows<float>(const float*, int, int, long long, int, float*, cudaStream_t);
template int colsum_partial_rows<bf16>(const bf16*, int, int, long long, int, float*, cudaStream_t);
template int colsum_partial<float>(const float*, int, int, long long, float*, cudaStream_t);
template int colsum_partial<bf16>(const bf16*, int, int, long long, float*, cudaStream_t);

int adapter_grad_finalize_scratch_floats() { return AD_FIN_MAX * ew::AD_FIN_CTAS; }

// jobs: one per AdapterModule whose gradients are due; scratch: adapter_grad_finalize_scratch_floats() floats
int adapter_grad_finalize(const AdapterGradJob* jobs, int n, float* scratch, cudaStream_t stream) {
  for (int base = 0; base < n; base += AD_FIN_MAX) {
    ew::AdFinBatch b;
    memset(&b, 0, sizeof(b));
    b.n = (n - base < AD_FIN_MAX) ? n - base : AD_FIN_MAX;
    for (int i = 0; i < b.n; ++i) {
      b.job[i] = jobs[base + i];
      FV_CHECK((b.job[i].E * b.job[i].A) % 4 == 0, "adapter_grad_finalize: E * A must be a multiple of 4");
    }
    ew::adapter_grad_phase1_kernel<<<dim3(ew::AD_FIN_CTAS, b.n), 256, 0, stream>>>(b, scratch);
    FV_COUNT_LAUNCH();
    ew::adapter_grad_phase2_kernel<<<1, 32, 0, stream>>>(b, scratch, ew::AD_FIN_CTAS);
    FV_COUNT_LAUNCH();
    FV_LAUNCH_CHECK();
  }
  return 0;
}

int dropout_mask(float* out, size_t n, Dropout drop, cudaStream_t stream) {
  ew::dropout_mask_kernel<<<ew::grid_for(n, 256), 256, 0, stream>>>(out, n, drop);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

int fill_f32(float* out, size_t n, float v, cudaStream_t stream) {
  if (n == 0) return 0;
  ew::fill_kernel<<<ew::grid_for(n, 256), 256, 0, stream>>>(out, n, v);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

}  // namespace fervit
