// LayerNorm forward / backward over the fp32 residual stream, one warp per token row, fp32 statistics.
// Replaces ATen native_layer_norm behind timm Block.norm1/norm2 (eps 1e-6, hybrid_latent_vit.py:227-233),
// nn.TransformerEncoderLayer.norm1/norm2 (eps 1e-5, latent_vit.py:24-31, image_vit.py:101-113).
// HBM-bound: a row is read once into registers (CH chunks of 128 columns per warp, CH a template parameter so
// E = 768 keeps 6 float4 per array, not 8), statistics by warp shuffles, every output written once.
#include "common.cuh"
#include "kernels.h"

namespace fervit {

namespace ln {

constexpr int WARPS = 8;

template <typename AT, int CH>
__global__ void __launch_bounds__(WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              float eps, int rows, int E, float* __restrict__ out_f32, AT* __restrict__ out_at,
              float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_trigger();
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + (size_t)row * E;
  float4 v[CH];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < E) {
      v[i] = *reinterpret_cast<const float4*>(xr + c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < E) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)E + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < E) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 y;
      y.x = (v[i].x - mean) * rstd * g.x + b.x;
      y.y = (v[i].y - mean) * rstd * g.y + b.y;
      y.z = (v[i].z - mean) * rstd * g.z + b.z;
      y.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)row * E + c) = y;
      if (out_at) store4<AT>(out_at + (size_t)row * E + c, y);
    }
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) (+ dres). Optional per-CTA partial sums of
// dgamma = sum dy*xhat and dbeta = sum dy, reduced afterwards in a fixed order (deterministic).
template <typename DT, typename AT, bool WGRAD, int CH>
__global__ void __launch_bounds__(WARPS * 32)
ln_bwd_kernel(const DT* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dres,
              int rows, int E, float* __restrict__ dx_f32, AT* __restrict__ dx_at, float* __restrict__ partial,
              Dropout at_drop) {
  __shared__ float red[WGRAD ? WARPS * 32 * 4 : 1];
  pdl_trigger();
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float4 dg[WGRAD ? CH : 1], db[WGRAD ? CH : 1];
  if (WGRAD) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int row = blockIdx.x * WARPS + warp; row < rows; row += gridDim.x * WARPS) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[CH], gd[CH], rsd[CH];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < E) {
        // the residual gradient is fetched up front with the other operands: one memory round trip per row, not two
        if (dres) rsd[i] = *reinterpret_cast<const float4*>(dres + (size_t)row * E + c);
        const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)row * E + c);
        const float4 d = load4<DT>(dy + (size_t)row * E + c);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        gd[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        s1 += (gd[i].x + gd[i].y) + (gd[i].z + gd[i].w);
        s2 += (gd[i].x * xh[i].x + gd[i].y * xh[i].y) + (gd[i].z * xh[i].z + gd[i].w * xh[i].w);
        if (WGRAD) {
          dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
          db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
        }
      }
    }
    s1 = warp_sum(s1) / (float)E;
    s2 = warp_sum(s2) / (float)E;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < E) {
        float4 o;
        o.x = rs * (gd[i].x - s1 - xh[i].x * s2);
        o.y = rs * (gd[i].y - s1 - xh[i].y * s2);
        o.z = rs * (gd[i].z - s1 - xh[i].z * s2);
        o.w = rs * (gd[i].w - s1 - xh[i].w * s2);
        if (dres) { o.x += rsd[i].x; o.y += rsd[i].y; o.z += rsd[i].z; o.w += rsd[i].w; }
        if (dx_f32) *reinterpret_cast<float4*>(dx_f32 + (size_t)row * E + c) = o;
        if (dx_at) {
          if (at_drop.threshold) {
            // the activation-dtype copy feeds the dgrad/wgrad of a sub-layer whose output went through dropout
            const uint64_t base = (uint64_t)row * E + c, seed = at_drop.eff();
            o.x = drop_keep(seed, at_drop.site, base + 0, at_drop.threshold) ? o.x * at_drop.scale : 0.f;
            o.y = drop_keep(seed, at_drop.site, base + 1, at_drop.threshold) ? o.y * at_drop.scale : 0.f;
            o.z = drop_keep(seed, at_drop.site, base + 2, at_drop.threshold) ? o.z * at_drop.scale : 0.f;
            o.w = drop_keep(seed, at_drop.site, base + 3, at_drop.threshold) ? o.w * at_drop.scale : 0.f;
          }
          store4<AT>(dx_at + (size_t)row * E + c, o);
        }
      }
    }
  }
  if (WGRAD) {
    // cross-warp reduction in shared memory, chunk by chunk; partial layout [grid][2][E]
    float* pg = partial + (size_t)blockIdx.x * 2 * E;
    float* pb = pg + E;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c0 = i * 128;
      if (c0 >= E) break;  // block-uniform
      for (int which = 0; which < 2; ++which) {
        const float4 val = which == 0 ? dg[i] : db[i];
        *reinterpret_cast<float4*>(&red[(warp * 32 + lane) * 4]) = val;
        __syncthreads();
        if (warp == 0) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int w = 0; w < WARPS; ++w) {
            const float4 t = *reinterpret_cast<const float4*>(&red[(w * 32 + lane) * 4]);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          }
          const int c = c0 + lane * 4;
          if (c < E) *reinterpret_cast<float4*>((which == 0 ? pg : pb) + c) = acc;
        }
        __syncthreads();
      }
    }
  }
}

static inline int chunks_for(int E) { return (E + 127) / 128; }

// Streaming backward (no parameter gradients, no dropout), TWO warps per row: each warp owns half of the row's 128-column
// chunks, the two partial sums meet in shared memory (named barrier per warp pair). ncu on the one-warp-per-row form at
// 4864 x 768: 120 registers -> 2 CTAs/SM, 2.05 waves, long-scoreboard stalls 9 per issue, L2 throughput 16 % of peak: a
// latency problem, not a bandwidth one. Half the per-thread state (<= 64 registers) puts 32 warps on an SM.
// 6 warps = 3 rows per CTA, 6 CTAs per SM: 18 rows in flight per SM, so the 4864 rows of batch 256 (32.9 per SM) take
// two rounds; 8-warp CTAs (16 rows in flight) needed a third, nearly empty one.
constexpr int PAIR_WARPS = 6;
template <typename DT, typename AT, int CHH>
__global__ void __launch_bounds__(PAIR_WARPS * 32, 6)
ln_bwd_pair_kernel(const DT* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dres,
                   int rows, int E, float* __restrict__ dx_f32, AT* __restrict__ dx_at) {
  __shared__ float2 red[PAIR_WARPS];
  pdl_trigger();
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int pair = warp >> 1, hw = warp & 1;
  const int row = blockIdx.x * (PAIR_WARPS / 2) + pair;
  const bool live = row < rows;
  float4 xv[CHH], dv[CHH], rsd[CHH];
  float s1 = 0.f, s2 = 0.f;
  float mu = 0.f, rs = 0.f;
  if (live) {
    mu = mean[row];
    rs = rstd[row];
#pragma unroll
    for (int i = 0; i < CHH; ++i) {
      const int c = lane * 4 + (hw * CHH + i) * 128;
      if (c < E) {
        if (dres) rsd[i] = *reinterpret_cast<const float4*>(dres + (size_t)row * E + c);
        xv[i] = *reinterpret_cast<const float4*>(x + (size_t)row * E + c);
        dv[i] = load4<DT>(dy + (size_t)row * E + c);
      }
    }
#pragma unroll
    for (int i = 0; i < CHH; ++i) {
      const int c = lane * 4 + (hw * CHH + i) * 128;
      if (c < E) {
        // gamma == nullptr: the norm's scale was folded into the weight that follows it, dy already is gamma * dh
        const float4 g = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
        xv[i] = make_float4((xv[i].x - mu) * rs, (xv[i].y - mu) * rs, (xv[i].z - mu) * rs, (xv[i].w - mu) * rs);  // xhat
        dv[i] = make_float4(dv[i].x * g.x, dv[i].y * g.y, dv[i].z * g.z, dv[i].w * g.w);                          // g * dy
        s1 += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
        s2 += (dv[i].x * xv[i].x + dv[i].y * xv[i].y) + (dv[i].z * xv[i].z + dv[i].w * xv[i].w);
      }
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) red[warp] = make_float2(s1, s2);
  asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
  const float2 a = red[pair * 2], b = red[pair * 2 + 1];   // fixed order: low half first
  s1 = (a.x + b.x) / (float)E;
  s2 = (a.y + b.y) / (float)E;
  if (!live) return;
#pragma unroll
  for (int i = 0; i < CHH; ++i) {
    const int c = lane * 4 + (hw * CHH + i) * 128;
    if (c < E) {
      float4 o;
      o.x = rs * (dv[i].x - s1 - xv[i].x * s2);
      o.y = rs * (dv[i].y - s1 - xv[i].y * s2);
      o.z = rs * (dv[i].z - s1 - xv[i].z * s2);
      o.w = rs * (dv[i].w - s1 - xv[i].w * s2);
      if (dres) { o.x += rsd[i].x; o.y += rsd[i].y; o.z += rsd[i].z; o.w += rsd[i].w; }
      if (dx_f32) *reinterpret_cast<float4*>(dx_f32 + (size_t)row * E + c) = o;
      if (dx_at) store4<AT>(dx_at + (size_t)row * E + c, o);
    }
  }
}

}  // namespace ln

#define FV_LN_DISPATCH(CALL)                      \
  do {                                            \
    const int ch_ = ln::chunks_for(E);            \
    if (ch_ <= 2) { CALL(2); }                    \
    else if (ch_ <= 4) { CALL(4); }               \
    else if (ch_ <= 6) { CALL(6); }               \
    else { CALL(8); }                             \
  } while (0)

template <typename AT>
int layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int rows, int E, float* out_f32,
                  AT* out_at, float* mean, float* rstd, cudaStream_t stream) {
  FV_CHECK(E % 4 == 0 && E <= 1024, "layernorm: E must be a multiple of 4 and <= 1024 (got %d)", E);
  if (rows <= 0) return 0;
  // algorithmic bytes: read x fp32, write each requested output, 8 B of statistics per row
  ProfScope prof(2, (double)rows * E * (4.0 + (out_f32 ? 4.0 : 0.0) + (out_at ? (double)sizeof(AT) : 0.0)) + rows * 8.0,
                 stream);
#define FV_LN_FWD(CH_)                                                                                          \
  FV_CUDA(launch_pdl(ln::ln_fwd_kernel<AT, CH_>, dim3(ceil_div(rows, ln::WARPS)), dim3(ln::WARPS * 32), 0, stream, x, \
                     gamma, beta, eps, rows, E, out_f32, out_at, mean, rstd))
  FV_LN_DISPATCH(FV_LN_FWD);
#undef FV_LN_FWD
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

int layernorm_bwd_grid(int rows) {
  int g = ceil_div(rows, ln::WARPS);
  const int cap = num_sms() * 2;
  return g < cap ? g : cap;
}

// partial: [layernorm_bwd_grid(rows)][2][E] fp32 when wgrad is requested (else null)
template <typename DT, typename AT>
int layernorm_bwd(const DT* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                  const float* dres, int rows, int E, float* dx_f32, AT* dx_at, float* partial, Dropout at_drop,
                  cudaStream_t stream) {
  FV_CHECK(E % 4 == 0 && E <= 1024, "layernorm: E must be a multiple of 4 and <= 1024 (got %d)", E);
  if (rows <= 0) return 0;
  FV_CHECK(gamma != nullptr || (partial == nullptr && !at_drop.threshold),
           "layernorm_bwd: a folded (null) gamma is only defined for the streaming form (no parameter gradients)");
  // algorithmic bytes: read dy and x (and dres), write each requested output, 8 B of statistics per row
  ProfScope prof(2, (double)rows * E * ((double)sizeof(DT) + 4.0 + (dres ? 4.0 : 0.0) + (dx_f32 ? 4.0 : 0.0) +
                                        (dx_at ? (double)sizeof(AT) : 0.0)) + rows * 8.0, stream);
  if (partial) {
    const int grid = layernorm_bwd_grid(rows);
#define FV_LN_BWD_W(CH_)                                                                                         \
  FV_CUDA(launch_pdl(ln::ln_bwd_kernel<DT, AT, true, CH_>, dim3(grid), dim3(ln::WARPS * 32), 0, stream, dy, x, mean,  \
                     rstd, gamma, dres, rows, E, dx_f32, dx_at, partial, at_drop))
    FV_LN_DISPATCH(FV_LN_BWD_W);
#undef FV_LN_BWD_W
  } else if (!at_drop.threshold) {
    // streaming form: two warps per row (see ln_bwd_pair_kernel)
#define FV_LN_BWD_P(CH_)                                                                                          \
  FV_CUDA(launch_pdl(ln::ln_bwd_pair_kernel<DT, AT, CH_ / 2>, dim3(ceil_div(rows, ln::PAIR_WARPS / 2)),           \
                     dim3(ln::PAIR_WARPS * 32), 0, stream, dy, x, mean, rstd, gamma, dres, rows, E, dx_f32, dx_at))
    FV_LN_DISPATCH(FV_LN_BWD_P);
#undef FV_LN_BWD_P
  } else {
#define FV_LN_BWD(CH_)                                                                                    \
  FV_CUDA(launch_pdl(ln::ln_bwd_kernel<DT, AT, false, CH_>, dim3(ceil_div(rows, ln::WARPS)), dim3(ln::WARPS * 32), 0,   \
                     stream, dy, x, mean, rstd, gamma, dres, rows, E, dx_f32, dx_at, (float*)nullptr, at_drop))
    FV_LN_DISPATCH(FV_LN_BWD);
#undef FV_LN_BWD
  }
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

template int layernorm_fwd<float>(const float*, const float*, const float*, float, int, int, float*, float*, float*,
                                  float*, cudaStream_t);
template int layernorm_fwd<bf16>(const float*, const float*, const float*, float, int, int, float*, bf16*, float*,
                                 float*, cudaStream_t);
template int layernorm_bwd<float, float>(const float*, const float*, const float*, const float*, const float*,
                                         const float*, int, int, float*, float*, float*, Dropout, cudaStream_t);
template int layernorm_bwd<bf16, bf16>(const bf16*, const float*, const float*, const float*, const float*,
                                       const float*, int, int, float*, bf16*, float*, Dropout, cudaStream_t);
template int layernorm_bwd<float, bf16>(const float*, const float*, const float*, const float*, const float*,
                                        const float*, int, int, float*, bf16*, float*, Dropout, cudaStream_t);

}  // namespace fervit
