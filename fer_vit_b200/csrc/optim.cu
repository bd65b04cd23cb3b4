// Fused AdamW (+ optional global-norm gradient clipping) over a list of fp32 parameter tensors: the step right after
// the hot path. Replaces `optim.AdamW(...)` + `clip_grad_norm_` of the reference trainers
// (train_hybrid_latent_vit.py:63-117, 244-248: five layer-wise LR groups; train_latent_vit_v2.py:132-133: clipping).
//   p <- p * (1 - lr * wd);  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2
//   p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)          (torch.optim.AdamW, amsgrad off)
// Tensors are passed BY VALUE in batches (pointer tables in kernel parameters), so nothing is staged through device
// memory and the launches are CUDA-graph capturable; the step counter lives on the device (one float, incremented by
// the first launch), per-group hyper-parameters in a small device table the host refreshes when they change.
#include "common.cuh"
#include "kernels.h"

namespace fervit {
namespace opt {

constexpr int BATCH = 24;        // tensors per launch
constexpr int CHUNK = 4096;      // elements per CTA (256 threads x 4 x 4)

struct Batch {
  float* p[BATCH];
  float* g[BATCH];
  float* m[BATCH];
  float* v[BATCH];
  long long numel[BATCH];
  int group[BATCH];
  int chunk0[BATCH + 1];         // first CTA of each tensor in the flat grid
  int n;
};

__device__ __forceinline__ int find_tensor(const Batch& b, int cta) {
  int t = 0;
  while (t + 1 < b.n && cta >= b.chunk0[t + 1]) ++t;
  return t;
}

// sum of squares per CTA -> partial[cta]
__global__ void __launch_bounds__(256) sumsq_kernel(const __grid_constant__ Batch b, float* __restrict__ partial) {
  __shared__ float red[8];
  pdl_trigger();
  pdl_grid_sync();
  const int t = find_tensor(b, blockIdx.x);
  const long long base = (long long)(blockIdx.x - b.chunk0[t]) * CHUNK;
  const float* __restrict__ g = b.g[t];
  const long long n = b.numel[t];
  float s = 0.f;
  for (int i = threadIdx.x; i < CHUNK; i += 256) {
    const long long k = base + i;
    if (k < n) { const float x = g[k]; s += x * x; }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += red[w];
    partial[blockIdx.x] = a;
  }
}
// clip_coef = min(1, max_norm / (sqrt(sum partial) + 1e-6))   (torch.nn.utils.clip_grad_norm_); fixed summation order
__global__ void __launch_bounds__(256) clip_coef_kernel(const float* __restrict__ partial, int n, float max_norm,
                                                        float* __restrict__ coef, float* __restrict__ total_norm) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = sqrtf(red[0]);
    const float c = max_norm / (norm + 1e-6f);
    coef[0] = c < 1.0f ? c : 1.0f;
    if (total_norm) total_norm[0] = norm;
  }
}
__global__ void step_increment_kernel(float* step) { step[0] += 1.0f; }

// hyper: [groups][5] = lr, beta1, beta2, eps, weight_decay. step[0] already holds t (1-based) of this update.
__global__ void __launch_bounds__(256)
adamw_kernel(const __grid_constant__ Batch b, const float* __restrict__ hyper, const float* __restrict__ step,
             const float* __restrict__ clip_coef, int write_back_grad) {
  pdl_trigger();
  pdl_grid_sync();
  const int t = find_tensor(b, blockIdx.x);
  const long long base = (long long)(blockIdx.x - b.chunk0[t]) * CHUNK;
  const long long n = b.numel[t];
  const float* h = hyper + b.group[t] * 5;
  const float lr = h[0], b1 = h[1], b2 = h[2], eps = h[3], wd = h[4];
  const float tt = step[0];
  const float bc1 = 1.0f - powf(b1, tt);
  const float bc2_rsqrt = rsqrtf(1.0f - powf(b2, tt));
  const float step_size = lr / bc1;
  const float decay = 1.0f - lr * wd;
  const float gs = clip_coef ? clip_coef[0] : 1.0f;
  float* __restrict__ p = b.p[t];
  float* __restrict__ g = b.g[t];
  float* __restrict__ m = b.m[t];
  float* __restrict__ v = b.v[t];
  const bool wb = write_back_grad && clip_coef;
  auto upd = [&](float& pk, float& gk, float& mk, float& vk) {
    gk *= gs;
    mk = b1 * mk + (1.0f - b1) * gk;
    vk = b2 * vk + (1.0f - b2) * gk * gk;
    const float denom = sqrtf(vk) * bc2_rsqrt + eps;
    pk = pk * decay - step_size * (mk / denom);
  };
  // 16-byte accesses, four independent vectors per thread in flight (all tensors are at least 16-byte aligned: torch
  // allocations and 64-element-aligned slices of the flat gradient buffer)
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
#pragma unroll
    for (int it = 0; it < CHUNK / (256 * 4); ++it) {
      const long long k = base + (long long)(it * 256 + threadIdx.x) * 4;
      if (k + 3 < n) {
        float4 pp = *reinterpret_cast<float4*>(p + k), gg = *reinterpret_cast<float4*>(g + k);
        float4 mm = *reinterpret_cast<float4*>(m + k), vv = *reinterpret_cast<float4*>(v + k);
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        *reinterpret_cast<float4*>(p + k) = pp;
        *reinterpret_cast<float4*>(m + k) = mm;
        *reinterpret_cast<float4*>(v + k) = vv;
        if (wb) *reinterpret_cast<float4*>(g + k) = gg;
      } else {
        for (long long q = k; q < n && q < k + 4; ++q) {
          float pk = p[q], gk = g[q], mk = m[q], vk = v[q];
          upd(pk, gk, mk, vk);
          p[q] = pk; m[q] = mk; v[q] = vk;
          if (wb) g[q] = gk;
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < CHUNK; i += 256) {
      const long long k = base + i;
      if (k >= n) break;
      float pk = p[k], gk = g[k], mk = m[k], vk = v[k];
      upd(pk, gk, mk, vk);
      p[k] = pk; m[k] = mk; v[k] = vk;
      if (wb) g[k] = gk;
    }
  }
}

}  // namespace opt

long long adamw_scratch_floats(int n, const long long* numel) {
  long long chunks = 0;
  for (int i = 0; i < n; ++i) chunks += (numel[i] + opt::CHUNK - 1) / opt::CHUNK;
  return chunks + 8;
}

// scratch: adamw_scratch_floats() floats when max_norm > 0 (partials + [coef, total_norm]); may be null otherwise.
int adamw_step(int n, float* const* p, float* const* g, float* const* m, float* const* v, const long long* numel,
               const int* group, const float* hyper, float* step, float max_norm, float* scratch, cudaStream_t stream) {
  FV_CHECK(n >= 0 && hyper && step, "adamw_step: null argument");
  if (n == 0) return 0;
  opt::step_increment_kernel<<<1, 1, 0, stream>>>(step);
  FV_COUNT_LAUNCH();
  auto fill = [&](opt::Batch& b, int base) {
    memset(&b, 0, sizeof(b));
    b.n = (n - base < opt::BATCH) ? n - base : opt::BATCH;
    int chunks = 0;
    for (int i = 0; i < b.n; ++i) {
      b.p[i] = p[base + i]; b.g[i] = g[base + i]; b.m[i] = m[base + i]; b.v[i] = v[base + i];
      b.numel[i] = numel[base + i]; b.group[i] = group[base + i];
      b.chunk0[i] = chunks;
      chunks += (int)((numel[base + i] + opt::CHUNK - 1) / opt::CHUNK);
    }
    b.chunk0[b.n] = chunks;
    return chunks;
  };
  float* coef = nullptr;
  if (max_norm > 0.f) {
    FV_CHECK(scratch != nullptr, "adamw_step: gradient clipping needs the scratch buffer");
    int total = 0;
    for (int base = 0; base < n; base += opt::BATCH) {
      opt::Batch b;
      const int chunks = fill(b, base);
      if (chunks == 0) continue;
      FV_CUDA(launch_pdl(opt::sumsq_kernel, dim3(chunks), dim3(256), 0, stream, b, scratch + total));
      FV_COUNT_LAUNCH();
      total += chunks;
    }
    coef = scratch + total;
    opt::clip_coef_kernel<<<1, 256, 0, stream>>>(scratch, total, max_norm, coef, coef + 1);
    FV_COUNT_LAUNCH();
  }
  for (int base = 0; base < n; base += opt::BATCH) {
    opt::Batch b;
    const int chunks = fill(b, base);
    if (chunks == 0) continue;
    FV_CUDA(launch_pdl(opt::adamw_kernel, dim3(chunks), dim3(256), 0, stream, b, hyper, (const float*)step,
                       (const float*)coef, 1));
    FV_COUNT_LAUNCH();
  }
  FV_LAUNCH_CHECK();
  return 0;
}

}  // namespace fervit
