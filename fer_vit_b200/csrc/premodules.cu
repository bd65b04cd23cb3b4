// w+ pre-modules of LatentViTv2 fused into one pass: SemanticPE -> LayerWiseNorm -> LEAM
// (application order of latent_vit_v2.py:82-84).
//   SPE : u = x + group_embed[groups[l]] + layer_embed[l]                  modules/semantic_pe.py:44-48
//   LWN : n = LN_l(u) (per-position gamma/beta, eps 1e-5);                 modules/layer_wise_norm.py:42-50
//         z = u + sigmoid(gate[l]) * (n - u) when use_residual else n
//   LEAM: y = z * sigmoid(w[l])                                            modules/leam.py:39-40
// One warp per (sample, layer) row of D <= 1024 values; backward re-derives u, n, z from x.
#include "common.cuh"
#include "kernels.h"

namespace fervit {

namespace pre {

constexpr int MAXCH = 8;


__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

template <typename AT, int CH>
__global__ void __launch_bounds__(128)
pre_fwd_kernel(const float* __restrict__ x, int B, int L, int D, Params p, float* __restrict__ out_f32,
               AT* __restrict__ out_at) {
  const int lane = threadIdx.x & 31;
  const size_t row = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= (size_t)B * L) return;
  const int l = (int)(row % L);
  const float* xr = x + row * D;
  float4 u[CH];
  float s = 0.f;
  const int grp = p.use_spe ? (int)p.groups[l] : 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < D) {
      u[i] = *reinterpret_cast<const float4*>(xr + c);
      if (p.use_spe) {
        const float4 ge = __ldg(reinterpret_cast<const float4*>(p.group_embed + (size_t)grp * D + c));
        const float4 le = __ldg(reinterpret_cast<const float4*>(p.layer_embed + (size_t)l * D + c));
        // pe = group + layer first, then x + pe (same association as the reference)
        u[i].x += ge.x + le.x; u[i].y += ge.y + le.y; u[i].z += ge.z + le.z; u[i].w += ge.w + le.w;
      }
      s += (u[i].x + u[i].y) + (u[i].z + u[i].w);
    }
  }
  float mean = 0.f, rstd = 0.f, sg = 1.f;
  if (p.use_lwn) {
    mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < D) {
        const float a = u[i].x - mean, b = u[i].y - mean, cc = u[i].z - mean, d = u[i].w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    rstd = rsqrtf(warp_sum(q) / (float)D + p.eps);
    if (p.use_res) sg = sigmoidf_(p.gate[l]);
  }
  const float sw = p.use_leam ? sigmoidf_(p.leam_w[l]) : 1.0f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < D) {
      float4 z = u[i];
      if (p.use_lwn) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + (size_t)l * D + c));
        const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + (size_t)l * D + c));
        float4 n;
        n.x = (u[i].x - mean) * rstd * g.x + be.x;
        n.y = (u[i].y - mean) * rstd * g.y + be.y;
        n.z = (u[i].z - mean) * rstd * g.z + be.z;
        n.w = (u[i].w - mean) * rstd * g.w + be.w;
        if (p.use_res) {
          z.x = u[i].x + sg * (n.x - u[i].x); z.y = u[i].y + sg * (n.y - u[i].y);
          z.z = u[i].z + sg * (n.z - u[i].z); z.w = u[i].w + sg * (n.w - u[i].w);
        } else {
          z = n;
        }
      }
      z.x *= sw; z.y *= sw; z.z *= sw; z.w *= sw;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * D + c) = z;
      if (out_at) store4<AT>(out_at + row * D + c, z);
    }
  }
}

// Backward. grid = (L, chunks); CTA (l, chunk) walks samples b = chunk*bpc .. with its 4 warps.
// partial_vec[chunk][l][3][D] : dgamma, dbeta, dlayer_embed ; partial_sc[chunk][l][2] : dleam_raw, dgate_raw
template <typename DT, int CH>
__global__ void __launch_bounds__(128)
pre_bwd_kernel(const float* __restrict__ x, const DT* __restrict__ dy, int B, int L, int D, Params p, int bpc,
               float* __restrict__ dx, float* __restrict__ partial_vec, float* __restrict__ partial_sc) {
  __shared__ float red[4 * 32 * 4];
  __shared__ float red_sc[4][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = blockIdx.x;
  const int b0 = blockIdx.y * bpc, b1 = min(B, b0 + bpc);
  const int grp = p.use_spe ? (int)p.groups[l] : 0;
  float4 dgam[CH], dbet[CH], dle[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    dgam[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    dbet[i] = dgam[i];
    dle[i] = dgam[i];
  }
  float dleam = 0.f, dgate = 0.f;
  const float sw = p.use_leam ? sigmoidf_(p.leam_w[l]) : 1.0f;
  const float sg = (p.use_lwn && p.use_res) ? sigmoidf_(p.gate[l]) : 1.0f;
  for (int b = b0 + warp; b < b1; b += 4) {
    const size_t row = (size_t)b * L + l;
    float4 u[CH];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < D) {
        u[i] = *reinterpret_cast<const float4*>(x + row * D + c);
        if (p.use_spe) {
          const float4 ge = __ldg(reinterpret_cast<const float4*>(p.group_embed + (size_t)grp * D + c));
          const float4 le = __ldg(reinterpret_cast<const float4*>(p.layer_embed + (size_t)l * D + c));
          u[i].x += ge.x + le.x; u[i].y += ge.y + le.y; u[i].z += ge.z + le.z; u[i].w += ge.w + le.w;
        }
        s += (u[i].x + u[i].y) + (u[i].z + u[i].w);
      }
    }
    float mean = 0.f, rstd = 0.f;
    if (p.use_lwn) {
      mean = warp_sum(s) / (float)D;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < D) {
          const float a = u[i].x - mean, bb = u[i].y - mean, cc = u[i].z - mean, d = u[i].w - mean;
          q += (a * a + bb * bb) + (cc * cc + d * d);
        }
      }
      rstd = rsqrtf(warp_sum(q) / (float)D + p.eps);
    }
    // pass 1: dz, dn, sums for the LN backward
    float4 dnv[CH], xh[CH], dud[CH];
    float s1 = 0.f, s2 = 0.f, acc_leam = 0.f, acc_gate = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < D) {
        const float4 d = load4<DT>(dy + row * D + c);
        const float dyv[4] = {d.x, d.y, d.z, d.w};
        const float uv[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
        float gv[4] = {1.f, 1.f, 1.f, 1.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.use_lwn) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + (size_t)l * D + c));
          const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + (size_t)l * D + c));
          gv[0] = g.x; gv[1] = g.y; gv[2] = g.z; gv[3] = g.w;
          bv[0] = be.x; bv[1] = be.y; bv[2] = be.z; bv[3] = be.w;
        }
        float xhv[4], dn[4], du_direct[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float z = uv[t];
          float n = 0.f;
          xhv[t] = 0.f;
          if (p.use_lwn) {
            xhv[t] = (uv[t] - mean) * rstd;
            n = xhv[t] * gv[t] + bv[t];
            z = p.use_res ? uv[t] + sg * (n - uv[t]) : n;
          }
          acc_leam += dyv[t] * z;
          const float dz = dyv[t] * sw;
          if (p.use_lwn) {
            if (p.use_res) {
              dn[t] = dz * sg;
              du_direct[t] = dz * (1.0f - sg);
              acc_gate += dz * (n - uv[t]);
            } else {
              dn[t] = dz;
              du_direct[t] = 0.f;
            }
            s1 += dn[t] * gv[t];
            s2 += dn[t] * gv[t] * xhv[t];
          } else {
            dn[t] = 0.f;
            du_direct[t] = dz;
          }
        }
        dnv[i] = make_float4(dn[0], dn[1], dn[2], dn[3]);
        xh[i] = make_float4(xhv[0], xhv[1], xhv[2], xhv[3]);
        dud[i] = make_float4(du_direct[0], du_direct[1], du_direct[2], du_direct[3]);
        if (p.use_lwn) {
          dgam[i].x += dn[0] * xhv[0]; dgam[i].y += dn[1] * xhv[1]; dgam[i].z += dn[2] * xhv[2]; dgam[i].w += dn[3] * xhv[3];
          dbet[i].x += dn[0]; dbet[i].y += dn[1]; dbet[i].z += dn[2]; dbet[i].w += dn[3];
        }
      }
    }
    dleam += acc_leam;
    dgate += acc_gate;
    if (p.use_lwn) {
      s1 = warp_sum(s1) / (float)D;
      s2 = warp_sum(s2) / (float)D;
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < D) {
        float4 du = dud[i];
        if (p.use_lwn) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + (size_t)l * D + c));
          du.x += rstd * (dnv[i].x * g.x - s1 - xh[i].x * s2);
          du.y += rstd * (dnv[i].y * g.y - s1 - xh[i].y * s2);
          du.z += rstd * (dnv[i].z * g.z - s1 - xh[i].z * s2);
          du.w += rstd * (dnv[i].w * g.w - s1 - xh[i].w * s2);
        }
        dle[i].x += du.x; dle[i].y += du.y; dle[i].z += du.z; dle[i].w += du.w;
        if (dx) *reinterpret_cast<float4*>(dx + row * D + c) = du;
      }
    }
  }
  // cross-warp reduction
  dleam = warp_sum(dleam);
  dgate = warp_sum(dgate);
  if (lane == 0) { red_sc[warp][0] = dleam; red_sc[warp][1] = dgate; }
  float* pv = partial_vec + ((size_t)blockIdx.y * L + l) * 3 * D;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    if (i * 128 >= D) break;
    for (int which = 0; which < 3; ++which) {
      const float4 val = which == 0 ? dgam[i] : (which == 1 ? dbet[i] : dle[i]);
      *reinterpret_cast<float4*>(&red[(warp * 32 + lane) * 4]) = val;
      __syncthreads();
      if (warp == 0) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < 4; ++w) {
          const float4 t = *reinterpret_cast<const float4*>(&red[(w * 32 + lane) * 4]);
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        const int c = i * 128 + lane * 4;
        if (c < D) *reinterpret_cast<float4*>(pv + (size_t)which * D + c) = acc;
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    float a = 0.f, g = 0.f;
    for (int w = 0; w < 4; ++w) { a += red_sc[w][0]; g += red_sc[w][1]; }
    partial_sc[((size_t)blockIdx.y * L + l) * 2 + 0] = a;
    partial_sc[((size_t)blockIdx.y * L + l) * 2 + 1] = g;
  }
}

// grid = L CTAs; reduces chunks in order; group embed grads summed over member layers by CTA 0..2 afterwards
__global__ void pre_bwd_final_kernel(const float* __restrict__ partial_vec, const float* __restrict__ partial_sc,
                                     int chunks, int L, int D, Params p, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, float* __restrict__ dlayer, float* __restrict__ dgate,
                                     float* __restrict__ dleam) {
  const int l = blockIdx.x;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f, b = 0.f, e = 0.f;
    for (int k = 0; k < chunks; ++k) {
      const float* pv = partial_vec + ((size_t)k * L + l) * 3 * D;
      a += pv[c]; b += pv[D + c]; e += pv[2 * D + c];
    }
    if (p.use_lwn) { dgamma[(size_t)l * D + c] = a; dbeta[(size_t)l * D + c] = b; }
    if (p.use_spe) dlayer[(size_t)l * D + c] = e;
  }
  if (threadIdx.x == 0) {
    float a = 0.f, g = 0.f;
    for (int k = 0; k < chunks; ++k) {
      a += partial_sc[((size_t)k * L + l) * 2 + 0];
      g += partial_sc[((size_t)k * L + l) * 2 + 1];
    }
    if (p.use_leam) {
      const float s = sigmoidf_(p.leam_w[l]);
      dleam[l] = a * s * (1.0f - s);
    }
    if (p.use_lwn && p.use_res) {
      const float s = sigmoidf_(p.gate[l]);
      dgate[l] = g * s * (1.0f - s);
    }
  }
}
__global__ void pre_group_grad_kernel(const float* __restrict__ dlayer, const long long* __restrict__ groups, int L,
                                      int D, float* __restrict__ dgroup) {
  const int g = blockIdx.x;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < L; ++l)
      if ((int)groups[l] == g) s += dlayer[(size_t)l * D + c];
    dgroup[(size_t)g * D + c] = s;
  }
}

}  // namespace pre


template <typename AT>
int premodules_fwd(const float* x, int B, int L, int D, const PreParams& p, float* out_f32, AT* out_at,
                   cudaStream_t stream) {
  FV_CHECK(D % 4 == 0 && D <= pre::MAXCH * 128, "premodules: D must be a multiple of 4 and <= 1024 (got %d)", D);
  const size_t rows = (size_t)B * L;
  if (D <= 512) pre::pre_fwd_kernel<AT, 4><<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>(x, B, L, D, p, out_f32, out_at);
  else pre::pre_fwd_kernel<AT, 8><<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>(x, B, L, D, p, out_f32, out_at);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template int premodules_fwd<float>(const float*, int, int, int, const PreParams&, float*, float*, cudaStream_t);
template int premodules_fwd<bf16>(const float*, int, int, int, const PreParams&, float*, bf16*, cudaStream_t);

int premodules_chunks(int B) {
  int c = ceil_div(B, 16);
  if (c > 64) c = 64;
  return c < 1 ? 1 : c;
}

// scratch: chunks * L * (3*D + 2) floats
template <typename DT>
int premodules_bwd(const float* x, const DT* dy, int B, int L, int D, const PreParams& p, float* dx, float* scratch,
                   float* dgamma, float* dbeta, float* dlayer, float* dgroup, float* dgate, float* dleam,
                   cudaStream_t stream) {
  FV_CHECK(D % 4 == 0 && D <= pre::MAXCH * 128, "premodules: D must be a multiple of 4 and <= 1024 (got %d)", D);
  const int chunks = premodules_chunks(B);
  const int bpc = ceil_div(B, chunks);
  float* pvec = scratch;
  float* psc = scratch + (size_t)chunks * L * 3 * D;
  dim3 grid(L, chunks);
  if (D <= 512) pre::pre_bwd_kernel<DT, 4><<<grid, 128, 0, stream>>>(x, dy, B, L, D, p, bpc, dx, pvec, psc);
  else pre::pre_bwd_kernel<DT, 8><<<grid, 128, 0, stream>>>(x, dy, B, L, D, p, bpc, dx, pvec, psc);
  FV_COUNT_LAUNCH();
  pre::pre_bwd_final_kernel<<<L, 256, 0, stream>>>(pvec, psc, chunks, L, D, p, dgamma, dbeta, dlayer, dgate, dleam);
  FV_COUNT_LAUNCH();
  if (p.use_spe) {
    pre::pre_group_grad_kernel<<<3, 256, 0, stream>>>(dlayer, p.groups, L, D, dgroup);
    FV_COUNT_LAUNCH();
  }
  FV_LAUNCH_CHECK();
  return 0;
}
template int premodules_bwd<float>(const float*, const float*, int, int, int, const PreParams&, float*, float*,
                                   float*, float*, float*, float*, float*, float*, cudaStream_t);
template int premodules_bwd<bf16>(const float*, const bf16*, int, int, int, const PreParams&, float*, float*, float*,
                                  float*, float*, float*, float*, float*, cudaStream_t);

}  // namespace fervit
