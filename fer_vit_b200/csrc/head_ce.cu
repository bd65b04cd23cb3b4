// Classifier head (CLS gather -> LayerNorm -> [Dropout] -> Linear) and cross-entropy, forward and backward.
// Head: latent_vit.py:33-36,46-47 (LN, Linear); hybrid_latent_vit.py:110-114,236-237 (LN, Dropout(0.1), Linear);
//       image_vit.py:116-117,161-164 (norm, head).
// Loss: nn.CrossEntropyLoss(weight=?, label_smoothing=?) as built at train_hybrid_latent_vit.py:236-241 and
//       train_latent_vit.py:248-253; mean reduction divides by sum_i w[y_i].
#include "common.cuh"
#include "kernels.h"

namespace fervit {

namespace head {

constexpr int MAXC = 16;
constexpr int MAXCH = 8;

// one warp per sample
__global__ void __launch_bounds__(128)
head_fwd_kernel(const float* __restrict__ x, int B, int S, int E, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, const float* __restrict__ W, const float* __restrict__ bias,
                int C, Dropout drop, float* __restrict__ logits, float* __restrict__ mean_out,
                float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* xr = x + (size_t)b * S * E;
  float4 v[MAXCH];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < E) {
      v[i] = *reinterpret_cast<const float4*>(xr + c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < E) {
      const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + bb * bb) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)E + eps);
  if (lane == 0) { mean_out[b] = mean; rstd_out[b] = rstd; }
  float acc[MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; ++k) acc[k] = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < E) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
      float hv[4] = {(v[i].x - mean) * rstd * g.x + be.x, (v[i].y - mean) * rstd * g.y + be.y,
                     (v[i].z - mean) * rstd * g.z + be.z, (v[i].w - mean) * rstd * g.w + be.w};
      if (drop.threshold) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
          hv[t] = drop_keep(drop.eff(), drop.site, (uint64_t)b * E + c + t, drop.threshold) ? hv[t] * drop.scale : 0.f;
      }
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        if (k < C) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)k * E + c));
          acc[k] += (hv[0] * w.x + hv[1] * w.y) + (hv[2] * w.z + hv[3] * w.w);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < MAXC; ++k) {
    if (k < C) {
      const float t = warp_sum(acc[k]);
      if (lane == 0) logits[(size_t)b * C + k] = t + bias[k];
    }
  }
}

// dx_final[b, 0, :] = LN backward of dn = mask * (dlogits W); rows 1.. of the sample are zero.
template <typename AT>
__global__ void __launch_bounds__(128)
head_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dlogits, int B, int S, int E,
                   const float* __restrict__ gamma, const float* __restrict__ W, int C,
                   const float* __restrict__ mean, const float* __restrict__ rstd, Dropout drop,
                   float* __restrict__ dx_f32, AT* __restrict__ dx_at) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* xr = x + (size_t)b * S * E;
  const float mu = mean[b], rs = rstd[b];
  float dl[MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; ++k) dl[k] = (k < C) ? dlogits[(size_t)b * C + k] : 0.f;
  float4 xh[MAXCH], gd[MAXCH];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < E) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + c);
      float dn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        if (k < C) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)k * E + c));
          dn[0] += dl[k] * w.x; dn[1] += dl[k] * w.y; dn[2] += dl[k] * w.z; dn[3] += dl[k] * w.w;
        }
      }
      if (drop.threshold) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
          dn[t] = drop_keep(drop.eff(), drop.site, (uint64_t)b * E + c + t, drop.threshold) ? dn[t] * drop.scale : 0.f;
      }
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      gd[i] = make_float4(dn[0] * g.x, dn[1] * g.y, dn[2] * g.z, dn[3] * g.w);
      s1 += (gd[i].x + gd[i].y) + (gd[i].z + gd[i].w);
      s2 += (gd[i].x * xh[i].x + gd[i].y * xh[i].y) + (gd[i].z * xh[i].z + gd[i].w * xh[i].w);
    }
  }
  s1 = warp_sum(s1) / (float)E;
  s2 = warp_sum(s2) / (float)E;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < E) {
      float4 o;
      o.x = rs * (gd[i].x - s1 - xh[i].x * s2);
      o.y = rs * (gd[i].y - s1 - xh[i].y * s2);
      o.z = rs * (gd[i].z - s1 - xh[i].z * s2);
      o.w = rs * (gd[i].w - s1 - xh[i].w * s2);
      if (dx_f32) *reinterpret_cast<float4*>(dx_f32 + (size_t)b * S * E + c) = o;
      if (dx_at) store4<AT>(dx_at + (size_t)b * S * E + c, o);
    }
  }
  // zero the non-cls rows of this sample
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int idx = E + lane * 4; idx < S * E; idx += 128) {
    if (dx_f32) *reinterpret_cast<float4*>(dx_f32 + (size_t)b * S * E + idx) = z;
    if (dx_at) store4<AT>(dx_at + (size_t)b * S * E + idx, z);
  }
}

// partial[chunk][C+2][E]: rows 0..C-1 = dW, row C = dgamma, row C+1 = dbeta ; thread per column e
__global__ void __launch_bounds__(128)
head_wgrad_partial_kernel(const float* __restrict__ x, const float* __restrict__ dlogits, int B, int S, int E,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ W, int C, const float* __restrict__ mean,
                          const float* __restrict__ rstd, Dropout drop, int b_per_chunk,
                          float* __restrict__ partial) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int b0 = blockIdx.y * b_per_chunk;
  const int b1 = min(B, b0 + b_per_chunk);
  float dw[MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; ++k) dw[k] = 0.f;
  float dg = 0.f, db = 0.f;
  float wcol[MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; ++k) wcol[k] = (k < C) ? W[(size_t)k * E + e] : 0.f;
  const float g = gamma[e], be = beta[e];
  for (int b = b0; b < b1; ++b) {
    const float xh = (x[(size_t)b * S * E + e] - mean[b]) * rstd[b];
    float m = 1.f;
    if (drop.threshold) m = drop_keep(drop.eff(), drop.site, (uint64_t)b * E + e, drop.threshold) ? drop.scale : 0.f;
    const float h = (xh * g + be) * m;
    float dn = 0.f;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      if (k < C) {
        const float d = dlogits[(size_t)b * C + k];
        dw[k] += d * h;
        dn += d * wcol[k];
      }
    }
    dn *= m;
    dg += dn * xh;
    db += dn;
  }
  float* p = partial + (size_t)blockIdx.y * (C + 2) * E;
#pragma unroll
  for (int k = 0; k < MAXC; ++k)
    if (k < C) p[(size_t)k * E + e] = dw[k];
  p[(size_t)C * E + e] = dg;
  p[(size_t)(C + 1) * E + e] = db;
}

__global__ void head_wgrad_final_kernel(const float* __restrict__ partial, int chunks, int C, int E,
                                        const float* __restrict__ dlogits, int B, float* __restrict__ dW,
                                        float* __restrict__ dgamma, float* __restrict__ dbeta,
                                        float* __restrict__ dbias) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = (C + 2) * E;
  if (idx < total) {
    float s = 0.f;
    for (int k = 0; k < chunks; ++k) s += partial[(size_t)k * total + idx];
    const int r = idx / E, e = idx % E;
    if (r < C) dW[(size_t)r * E + e] = s;
    else if (r == C) dgamma[e] = s;
    else dbeta[e] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x < C) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dlogits[(size_t)b * C + threadIdx.x];
    dbias[threadIdx.x] = s;
  }
}

// ------------------------------- cross entropy -------------------------------
// single CTA; loss = sum_i [ (1-eps) w[y_i] (-lp[i,y_i]) + (eps/C) sum_c w[c] (-lp[i,c]) ] / den
// dlogits[i,c] = (softmax[i,c] * sum_c t[i,c] - t[i,c]) / den,  t[i,c] = (1-eps) w[y_i] [c==y_i] + (eps/C) w[c]
__global__ void __launch_bounds__(256)
ce_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, const float* __restrict__ weight,
          float smoothing, int B, int C, const float* __restrict__ den_in, float grad_scale,
          float* __restrict__ loss_out, float* __restrict__ dlogits, float* __restrict__ den_out) {
  __shared__ float red[8];
  __shared__ float s_den;
  const int tid = threadIdx.x;
  float part = 0.f;
  if (den_in == nullptr) {
    for (int i = tid; i < B; i += blockDim.x) part += weight ? weight[labels[i]] : 1.0f;
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];
      s_den = t;
    }
    __syncthreads();
  } else {
    if (tid == 0) s_den = den_in[0];
    __syncthreads();
  }
  const float den = s_den;
  float lsum = 0.f;
  for (int i = tid; i < B; i += blockDim.x) {
    const float* z = logits + (size_t)i * C;
    float zl[MAXC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      zl[c] = (c < C) ? z[c] : -INFINITY;
      mx = fmaxf(mx, zl[c]);
    }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) se += expf(zl[c] - mx);
    const float lse = mx + logf(se);
    const int y = (int)labels[i];
    const float wy = weight ? weight[y] : 1.0f;
    float li = 0.f, tsum = 0.f;
    float t[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        const float wc = weight ? weight[c] : 1.0f;
        t[c] = (c == y ? (1.0f - smoothing) * wy : 0.f) + (smoothing / (float)C) * wc;
        li += t[c] * (lse - zl[c]);
        tsum += t[c];
      }
    }
    lsum += li;
    if (dlogits) {
      // sum_c dlogits[i,c] == 0 exactly (softmax sums to 1), so the target class takes minus the sum of the
      // others: avoids the cancellation in softmax[y] - 1 when the sample is already classified with p ~ 1
      float others = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C && c != y) {
          const float d = grad_scale * (expf(zl[c] - lse) * tsum - t[c]) / den;
          dlogits[(size_t)i * C + c] = d;
          others += d;
        }
      }
      dlogits[(size_t)i * C + y] = -others;
    }
  }
  lsum = warp_sum(lsum);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = lsum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    loss_out[0] = t / den;
    if (den_out) den_out[0] = den;
  }
}


// Mixup loss of the LatentViT trainers (train_latent_vit.py:131):
//   loss = lam * CE(z, y) + (1 - lam) * CE(z, y[index]), each term with its own denominator sum_i w[label_i];
// one launch, one pass over the logits: per row the two target distributions are merged as
//   t[i,c] = lam * ta[i,c] / den_a + (1 - lam) * tb[i,c] / den_b,  loss = sum_i sum_c t[i,c] (lse_i - z[i,c]),
//   dlogits[i,c] = softmax[i,c] * sum_c t[i,c] - t[i,c].
__global__ void __launch_bounds__(256)
ce_mixup_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                const long long* __restrict__ index, const float* __restrict__ weight, float smoothing, int B, int C,
                float lam_host, const float* __restrict__ lam_dev, float grad_scale, float* __restrict__ loss_out,
                float* __restrict__ dlogits) {
  __shared__ float red[2][8];
  __shared__ float s_den[2];
  const int tid = threadIdx.x;
  const float lam = lam_dev ? lam_dev[0] : lam_host;
  float pa = 0.f, pb = 0.f;
  for (int i = tid; i < B; i += blockDim.x) {
    pa += weight ? weight[labels[i]] : 1.0f;
    pb += weight ? weight[labels[index[i]]] : 1.0f;
  }
  pa = warp_sum(pa);
  pb = warp_sum(pb);
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = pa;
    red[1][tid >> 5] = pb;
  }
  __syncthreads();
  if (tid < 2) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[tid][w];
    s_den[tid] = t;
  }
  __syncthreads();
  const float ka = lam / s_den[0], kb = (1.0f - lam) / s_den[1];
  float lsum = 0.f;
  for (int i = tid; i < B; i += blockDim.x) {
    const float* z = logits + (size_t)i * C;
    float zl[MAXC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      zl[c] = (c < C) ? z[c] : -INFINITY;
      mx = fmaxf(mx, zl[c]);
    }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) se += expf(zl[c] - mx);
    const float lse = mx + logf(se);
    const int ya = (int)labels[i], yb = (int)labels[index[i]];
    const float wa = weight ? weight[ya] : 1.0f, wb = weight ? weight[yb] : 1.0f;
    float li = 0.f, tsum = 0.f;
    float t[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        const float wc = weight ? weight[c] : 1.0f;
        const float sm = (smoothing / (float)C) * wc;
        t[c] = ka * ((c == ya ? (1.0f - smoothing) * wa : 0.f) + sm) +
               kb * ((c == yb ? (1.0f - smoothing) * wb : 0.f) + sm);
        li += t[c] * (lse - zl[c]);
        tsum += t[c];
      }
    }
    lsum += li;
    if (dlogits) {
      // the row of dlogits sums to zero exactly; the heavier target class takes minus the sum of the others
      const int ypiv = (lam >= 0.5f) ? ya : yb;
      float others = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C && c != ypiv) {
          const float d = grad_scale * (expf(zl[c] - lse) * tsum - t[c]);
          dlogits[(size_t)i * C + c] = d;
          others += d;
        }
      }
      dlogits[(size_t)i * C + ypiv] = -others;
    }
  }
  lsum = warp_sum(lsum);
  __syncthreads();
  if ((tid & 31) == 0) red[0][tid >> 5] = lsum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[0][w];
    loss_out[0] = t;
  }
}

}  // namespace head

int head_fwd(const float* x, int B, int S, int E, const float* gamma, const float* beta, float eps, const float* W,
             const float* bias, int C, Dropout drop, float* logits, float* mean, float* rstd, cudaStream_t stream) {
  FV_CHECK(E % 4 == 0 && E <= head::MAXCH * 128, "head: E must be a multiple of 4 and <= 1024 (got %d)", E);
  FV_CHECK(C >= 1 && C <= head::MAXC, "head: num_classes must be in [1, %d] (got %d)", head::MAXC, C);
  head::head_fwd_kernel<<<ceil_div(B, 4), 128, 0, stream>>>(x, B, S, E, gamma, beta, eps, W, bias, C, drop, logits,
                                                            mean, rstd);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

int head_wgrad_chunks(int B) {
  int c = ceil_div(B, 32);
  if (c > 128) c = 128;
  return c < 1 ? 1 : c;
}

// scratch: [head_wgrad_chunks(B)][C+2][E] fp32. Any of dW/dgamma/dbeta/dbias may not be null when wgrad != 0.
template <typename AT>
int head_bwd(const float* x, const float* dlogits, int B, int S, int E, const float* gamma, const float* beta,
             const float* W, int C, const float* mean, const float* rstd, Dropout drop, float* dx_f32, AT* dx_at,
             int wgrad, float* scratch, float* dW, float* dgamma, float* dbeta, float* dbias, cudaStream_t stream,
             cudaStream_t wgrad_stream) {
  FV_CHECK(E % 4 == 0 && E <= head::MAXCH * 128, "head: E must be a multiple of 4 and <= 1024 (got %d)", E);
  FV_CHECK(C >= 1 && C <= head::MAXC, "head: num_classes must be in [1, %d] (got %d)", head::MAXC, C);
  head::head_bwd_dx_kernel<AT><<<ceil_div(B, 4), 128, 0, stream>>>(x, dlogits, B, S, E, gamma, W, C, mean, rstd, drop,
                                                                  dx_f32, dx_at);
  FV_COUNT_LAUNCH();
  if (wgrad) {
    const int chunks = head_wgrad_chunks(B);
    const int bpc = ceil_div(B, chunks);
    dim3 grid(ceil_div(E, 128), chunks);
    // the parameter gradients do not feed the dgrad chain: the caller may put them on a side stream
    head::head_wgrad_partial_kernel<<<grid, 128, 0, wgrad_stream>>>(x, dlogits, B, S, E, gamma, beta, W, C, mean, rstd,
                                                                   drop, bpc, scratch);
    FV_COUNT_LAUNCH();
    head::head_wgrad_final_kernel<<<ceil_div((C + 2) * E, 256), 256, 0, wgrad_stream>>>(scratch, chunks, C, E, dlogits, B,
                                                                                       dW, dgamma, dbeta, dbias);
    FV_COUNT_LAUNCH();
  }
  FV_LAUNCH_CHECK();
  return 0;
}
template int head_bwd<float>(const float*, const float*, int, int, int, const float*, const float*, const float*, int,
                             const float*, const float*, Dropout, float*, float*, int, float*, float*, float*, float*,
                             float*, cudaStream_t, cudaStream_t);
template int head_bwd<bf16>(const float*, const float*, int, int, int, const float*, const float*, const float*, int,
                            const float*, const float*, Dropout, float*, bf16*, int, float*, float*, float*, float*,
                            float*, cudaStream_t, cudaStream_t);

int cross_entropy(const float* logits, const long long* labels, const float* weight, float smoothing, int B, int C,
                  const float* den_in, float grad_scale, float* loss, float* dlogits, float* den_out,
                  cudaStream_t stream) {
  FV_CHECK(C >= 1 && C <= head::MAXC, "cross_entropy: num_classes must be in [1, %d] (got %d)", head::MAXC, C);
  FV_CHECK(B >= 1, "cross_entropy: empty batch");
  head::ce_kernel<<<1, 256, 0, stream>>>(logits, labels, weight, smoothing, B, C, den_in, grad_scale, loss, dlogits,
                                         den_out);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

int cross_entropy_mixup(const float* logits, const long long* labels, const long long* index, const float* weight,
                        float smoothing, int B, int C, float lam, const float* lam_dev, float grad_scale, float* loss,
                        float* dlogits, cudaStream_t stream) {
  FV_CHECK(C >= 1 && C <= head::MAXC, "cross_entropy_mixup: num_classes must be in [1, %d] (got %d)", head::MAXC, C);
  FV_CHECK(B >= 1, "cross_entropy_mixup: empty batch");
  FV_CHECK(lam_dev || (lam >= 0.f && lam <= 1.f), "cross_entropy_mixup: lam must be in [0, 1] (got %g)", (double)lam);
  head::ce_mixup_kernel<<<1, 256, 0, stream>>>(logits, labels, index, weight, smoothing, B, C, lam, lam_dev,
                                               grad_scale, loss, dlogits);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

}  // namespace fervit
