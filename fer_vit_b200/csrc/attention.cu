// Short-sequence multi-head attention (S = 19 w+ tokens + cls, 37 in concat mode, 197 for 224^2 patches).
// The whole K/V (and in backward Q/dO) of one (sample, head) lives in shared memory as fp32; one thread owns one
// query row (forward, dQ) or one key row (dK, dV); softmax is the online form, so no S x S buffer exists.
// Replaces F.scaled_dot_product_attention behind timm Attention (hybrid_latent_vit.py:227-233) and
// nn.MultiheadAttention (latent_vit.py:24-31, image_vit.py:101-113), including its attention-weight dropout.
//
// qkv layout: [B*S, 3E], columns [Q | K | V], head h at columns h*HD .. (timm qkv / torch in_proj packing).
#include "common.cuh"
#include "kernels.h"
#include <type_traits>

namespace fervit {

namespace attn {

template <int HD> struct Pad { static constexpr int LD = HD + 4; };

template <typename AT, int HD>
__device__ __forceinline__ void load_rows(const AT* __restrict__ base, size_t row_stride, int S, float* __restrict__ dst,
                                          int tid, int nthreads) {
  constexpr int LD = Pad<HD>::LD;
  constexpr int V = HD / 4;
  for (int i = tid; i < S * V; i += nthreads) {
    const int r = i / V, c = (i % V) * 4;
    const float4 v = load4<AT>(base + (size_t)r * row_stride + c);
    *reinterpret_cast<float4*>(dst + r * LD + c) = v;
  }
}

// grid: ceil(B*H / G) CTAs; block: G groups of SP = ceil32(S) threads; group g handles (b,h) = blockIdx*G + g
template <typename AT, int HD>
__global__ void attn_fwd_kernel(const AT* __restrict__ qkv, AT* __restrict__ out, float* __restrict__ lse, int B,
                                int S, int H, int SP, int G, float scale, Dropout drop) {
  extern __shared__ float smem[];
  constexpr int LD = Pad<HD>::LD;
  const int E = H * HD;
  const int g = threadIdx.x / SP;
  const int i = threadIdx.x % SP;
  const int bh = blockIdx.x * G + g;
  const bool valid = bh < B * H;
  float* Ks = smem + (size_t)g * 2 * S * LD;
  float* Vs = Ks + S * LD;
  const int b = valid ? bh / H : 0, h = valid ? bh % H : 0;
  const AT* base = qkv + (size_t)b * S * 3 * E + h * HD;
  if (valid) {
    load_rows<AT, HD>(base + E, 3 * E, S, Ks, i, SP);
    load_rows<AT, HD>(base + 2 * E, 3 * E, S, Vs, i, SP);
  }
  __syncthreads();
  if (!valid || i >= S) return;
  float q[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 v = load4<AT>(base + (size_t)i * 3 * E + d);
    q[d] = v.x * scale; q[d + 1] = v.y * scale; q[d + 2] = v.z * scale; q[d + 3] = v.w * scale;
    acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const uint64_t drop_base = ((uint64_t)bh * S + i) * S;
  for (int j = 0; j < S; ++j) {
    const float* kj = Ks + j * LD;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 kv = *reinterpret_cast<const float4*>(kj + d);
      s0 = fmaf(q[d], kv.x, s0); s1 = fmaf(q[d + 1], kv.y, s1);
      s2 = fmaf(q[d + 2], kv.z, s2); s3 = fmaf(q[d + 3], kv.w, s3);
    }
    const float s = (s0 + s1) + (s2 + s3);
    if (s > m) {
      const float corr = __expf(m - s);
      l *= corr;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] *= corr;
      m = s;
    }
    const float pj = __expf(s - m);
    l += pj;
    float pw = pj;
    if (drop.threshold) pw = drop_keep(drop.eff(), drop.site, drop_base + j, drop.threshold) ? pj * drop.scale : 0.f;
    const float* vj = Vs + j * LD;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 vv = *reinterpret_cast<const float4*>(vj + d);
      acc[d] = fmaf(pw, vv.x, acc[d]); acc[d + 1] = fmaf(pw, vv.y, acc[d + 1]);
      acc[d + 2] = fmaf(pw, vv.z, acc[d + 2]); acc[d + 3] = fmaf(pw, vv.w, acc[d + 3]);
    }
  }
  const float inv = 1.0f / l;
  AT* o = out + ((size_t)b * S + i) * E + h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 4)
    store4<AT>(o + d, make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv));
  if (lse) lse[(size_t)bh * S + i] = m + __logf(l);
}

// Backward. smem per group: Q, K, V, dO as fp32 [S][LD], then LSE[S], D[S].
//   phase A (thread = query i): dQ_i = scale * sum_j dS_ij K_j
//   phase B (thread = key j)  : dK_j = scale * sum_i dS_ij Q_i ; dV_j = sum_i Ptilde_ij dO_i
template <typename AT, int HD>
__global__ void attn_bwd_kernel(const AT* __restrict__ qkv, const AT* __restrict__ out, const AT* __restrict__ dout,
                                const float* __restrict__ lse, AT* __restrict__ dqkv, int B, int S, int H, int SP,
                                int G, float scale, Dropout drop) {
  extern __shared__ float smem[];
  constexpr int LD = Pad<HD>::LD;
  const int E = H * HD;
  const int g = threadIdx.x / SP;
  const int i = threadIdx.x % SP;
  const int bh = blockIdx.x * G + g;
  const bool valid = bh < B * H;
  const int SA = (S + 3) & ~3;  // keep every group 16-byte aligned
  const size_t per_group = (size_t)4 * S * LD + 2 * SA;
  float* Qs = smem + (size_t)g * per_group;
  float* Ks = Qs + S * LD;
  float* Vs = Ks + S * LD;
  float* dOs = Vs + S * LD;
  float* Ls = dOs + S * LD;
  float* Ds = Ls + SA;
  const int b = valid ? bh / H : 0, h = valid ? bh % H : 0;
  const AT* base = qkv + (size_t)b * S * 3 * E + h * HD;
  const AT* obase = out + (size_t)b * S * E + h * HD;
  const AT* dobase = dout + (size_t)b * S * E + h * HD;
  if (valid) {
    load_rows<AT, HD>(base, 3 * E, S, Qs, i, SP);
    load_rows<AT, HD>(base + E, 3 * E, S, Ks, i, SP);
    load_rows<AT, HD>(base + 2 * E, 3 * E, S, Vs, i, SP);
    load_rows<AT, HD>(dobase, E, S, dOs, i, SP);
    if (i < S) {
      // D_i = dO_i . O_i
      float dsum = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 o = load4<AT>(obase + (size_t)i * E + d);
        const float4 dd = load4<AT>(dobase + (size_t)i * E + d);
        dsum += (o.x * dd.x + o.y * dd.y) + (o.z * dd.z + o.w * dd.w);
      }
      Ds[i] = dsum;
      Ls[i] = lse[(size_t)bh * S + i];
    }
  }
  __syncthreads();
  if (!valid || i >= S) return;
  AT* dq_out = dqkv + ((size_t)b * S + i) * 3 * E + h * HD;
  // ---------------- phase A: dQ_i ----------------
  {
    float dq[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = 0.f;
    const float* qi = Qs + i * LD;
    const float* doi = dOs + i * LD;
    const float li = Ls[i], di = Ds[i];
    const uint64_t drop_base = ((uint64_t)bh * S + i) * S;
    for (int j = 0; j < S; ++j) {
      const float* kj = Ks + j * LD;
      const float* vj = Vs + j * LD;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(qi + d);
        const float4 kv = *reinterpret_cast<const float4*>(kj + d);
        const float4 ov = *reinterpret_cast<const float4*>(doi + d);
        const float4 vv = *reinterpret_cast<const float4*>(vj + d);
        s += (qv.x * kv.x + qv.y * kv.y) + (qv.z * kv.z + qv.w * kv.w);
        dp += (ov.x * vv.x + ov.y * vv.y) + (ov.z * vv.z + ov.w * vv.w);
      }
      const float pij = __expf(s * scale - li);
      if (drop.threshold) dp = drop_keep(drop.eff(), drop.site, drop_base + j, drop.threshold) ? dp * drop.scale : 0.f;
      const float ds = pij * (dp - di) * scale;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(kj + d);
        dq[d] = fmaf(ds, kv.x, dq[d]); dq[d + 1] = fmaf(ds, kv.y, dq[d + 1]);
        dq[d + 2] = fmaf(ds, kv.z, dq[d + 2]); dq[d + 3] = fmaf(ds, kv.w, dq[d + 3]);
      }
    }
#pragma unroll
    for (int d = 0; d < HD; d += 4) store4<AT>(dq_out + d, make_float4(dq[d], dq[d + 1], dq[d + 2], dq[d + 3]));
  }
  // ---------------- phase B: dK_j, dV_j (this thread's row index is the key j = i) ----------------
  {
    const int j = i;
    float dk[HD], dv[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    const float* kj = Ks + j * LD;
    const float* vj = Vs + j * LD;
    for (int r = 0; r < S; ++r) {
      const float* qr = Qs + r * LD;
      const float* dor = dOs + r * LD;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(qr + d);
        const float4 kv = *reinterpret_cast<const float4*>(kj + d);
        const float4 ov = *reinterpret_cast<const float4*>(dor + d);
        const float4 vv = *reinterpret_cast<const float4*>(vj + d);
        s += (qv.x * kv.x + qv.y * kv.y) + (qv.z * kv.z + qv.w * kv.w);
        dp += (ov.x * vv.x + ov.y * vv.y) + (ov.z * vv.z + ov.w * vv.w);
      }
      const float prj = __expf(s * scale - Ls[r]);
      float pt = prj;
      if (drop.threshold) {
        const bool keep = drop_keep(drop.eff(), drop.site, ((uint64_t)bh * S + r) * S + j, drop.threshold);
        pt = keep ? prj * drop.scale : 0.f;
        dp = keep ? dp * drop.scale : 0.f;
      }
      const float ds = prj * (dp - Ds[r]) * scale;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(qr + d);
        const float4 ov = *reinterpret_cast<const float4*>(dor + d);
        dk[d] = fmaf(ds, qv.x, dk[d]); dk[d + 1] = fmaf(ds, qv.y, dk[d + 1]);
        dk[d + 2] = fmaf(ds, qv.z, dk[d + 2]); dk[d + 3] = fmaf(ds, qv.w, dk[d + 3]);
        dv[d] = fmaf(pt, ov.x, dv[d]); dv[d + 1] = fmaf(pt, ov.y, dv[d + 1]);
        dv[d + 2] = fmaf(pt, ov.z, dv[d + 2]); dv[d + 3] = fmaf(pt, ov.w, dv[d + 3]);
      }
    }
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      store4<AT>(dq_out + E + d, make_float4(dk[d], dk[d + 1], dk[d + 2], dk[d + 3]));
      store4<AT>(dq_out + 2 * E + d, make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]));
    }
  }
}

static inline void geometry(int S, int& SP, int& G) {
  SP = ((S + 31) / 32) * 32;
  G = SP >= 128 ? 1 : 128 / SP;
}

}  // namespace attn

template <typename AT, int HD>
static int attention_fwd_t(const AT* qkv, AT* out, float* lse, int B, int S, int H, Dropout drop, cudaStream_t stream) {
  int SP, G;
  attn::geometry(S, SP, G);
  const size_t smem = (size_t)G * 2 * S * attn::Pad<HD>::LD * sizeof(float);
  FV_CHECK(smem <= 227 * 1024, "attention_fwd: sequence length %d does not fit in shared memory", S);
  FV_CHECK(SP * G <= 1024, "attention_fwd: sequence length %d too long for one CTA", S);
  auto kern = attn::attn_fwd_kernel<AT, HD>;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    FV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const float scale = 1.0f / sqrtf((float)HD);
  // algorithmic bytes: read Q,K,V once, write O once (+ LSE)
  ProfScope prof(1, (double)B * S * H * HD * 4.0 * sizeof(AT) + (double)B * H * S * 4.0, stream);
  kern<<<ceil_div(B * H, G), SP * G, smem, stream>>>(qkv, out, lse, B, S, H, SP, G, scale, drop);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

template <typename AT, int HD>
static int attention_bwd_t(const AT* qkv, const AT* out, const AT* dout, const float* lse, AT* dqkv, int B, int S,
                           int H, Dropout drop, cudaStream_t stream) {
  int SP, G;
  attn::geometry(S, SP, G);
  const size_t smem = (size_t)G * ((size_t)4 * S * attn::Pad<HD>::LD + 2 * ((S + 3) & ~3)) * sizeof(float);
  FV_CHECK(smem <= 227 * 1024, "attention_bwd: sequence length %d does not fit in shared memory", S);
  FV_CHECK(SP * G <= 1024, "attention_bwd: sequence length %d too long for one CTA", S);
  auto kern = attn::attn_bwd_kernel<AT, HD>;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    FV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const float scale = 1.0f / sqrtf((float)HD);
  // algorithmic bytes: read Q,K,V,O,dO once, write dQ,dK,dV once (+ LSE)
  ProfScope prof(1, (double)B * S * H * HD * 8.0 * sizeof(AT) + (double)B * H * S * 4.0, stream);
  kern<<<ceil_div(B * H, G), SP * G, smem, stream>>>(qkv, out, dout, lse, dqkv, B, S, H, SP, G, scale, drop);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

template <typename AT>
int attention_fwd(const AT* qkv, AT* out, float* lse, int B, int S, int H, int HD, Dropout drop, cudaStream_t stream) {
  if constexpr (std::is_same<AT, bf16>::value) {
    // S <= 32: warp-per-head tensor-core kernel (attention_tc.cu); longer sequences use the kernel below
    if (attention_tc_supported(S, HD)) return attention_tc_fwd(qkv, out, lse, B, S, H, HD, drop, stream);
    // 32 < S <= 256: CTA-per-head tensor-core kernel (attention_tc_long.cu)
    if (attention_tc_long_supported(S, HD)) return attention_tc_long_fwd(qkv, out, lse, B, S, H, HD, drop, stream);
  }
  if (HD == 64) return attention_fwd_t<AT, 64>(qkv, out, lse, B, S, H, drop, stream);
  if (HD == 48) return attention_fwd_t<AT, 48>(qkv, out, lse, B, S, H, drop, stream);
  if (HD == 32) return attention_fwd_t<AT, 32>(qkv, out, lse, B, S, H, drop, stream);
  FV_CHECK(false, "attention: head dim %d not supported (32, 48, 64)", HD);
}
template <typename AT>
int attention_bwd(const AT* qkv, const AT* out, const AT* dout, const float* lse, AT* dqkv, int B, int S, int H,
                  int HD, Dropout drop, cudaStream_t stream) {
  if constexpr (std::is_same<AT, bf16>::value) {
    if (attention_tc_supported(S, HD)) return attention_tc_bwd(qkv, out, dout, lse, dqkv, B, S, H, HD, drop, stream);
    if (attention_tc_long_supported(S, HD))
      return attention_tc_long_bwd(qkv, out, dout, lse, dqkv, B, S, H, HD, drop, stream);
  }
  if (HD == 64) return attention_bwd_t<AT, 64>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
  if (HD == 48) return attention_bwd_t<AT, 48>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
  if (HD == 32) return attention_bwd_t<AT, 32>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
  FV_CHECK(false, "attention: head dim %d not supported (32, 48, 64)", HD);
}

template int attention_fwd<float>(const float*, float*, float*, int, int, int, int, Dropout, cudaStream_t);
template int attention_fwd<bf16>(const bf16*, bf16*, float*, int, int, int, int, Dropout, cudaStream_t);
template int attention_bwd<float>(const float*, const float*, const float*, const float*, float*, int, int, int, int,
                                  Dropout, cudaStream_t);
template int attention_bwd<bf16>(const bf16*, const bf16*, const bf16*, const float*, bf16*, int, int, int, int,
                                 Dropout, cudaStream_t);

}  // namespace fervit
