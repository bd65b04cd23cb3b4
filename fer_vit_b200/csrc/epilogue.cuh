// The one GEMM epilogue used by every dense contraction on the path (see common.cuh: Epilogue).
//
// It is specialised at COMPILE time on its "kind": a kernel instantiation carries only the code its epilogue needs
// (exact-erf GELU and its derivative, the dropout hash and the token-row remap are each ~40-100 instructions; with
// all of them inlined behind runtime flags the persistent tcgen05 kernel outgrew the instruction cache and its
// epilogue warps ran ~2x slower than the MMAs they were supposed to hide behind — measured).
#pragma once
#include "common.cuh"

namespace fervit {

enum EpiKind {
  EPK_PLAIN = 0,     // bias / alpha / residual / outputs only
  EPK_GELU,          // + forward GELU (pre-activation optionally stored); fast erf (common.cuh), bf16 kernels only
  EPK_RELU,          // + forward ReLU
  EPK_GELU_BWD,      // accumulator times GELU'(aux)
  EPK_RELU_BWD,      // accumulator times ReLU'(aux)
  EPK_MUL_BWD,       // accumulator times aux (aux = saved activation derivative, ACT_DERIV)
  EPK_REMAP,         // token rows skip the cls slot, + position rows (input projection)
  EPK_GENERIC,       // everything behind runtime flags (dropout, unusual combinations)
  EPK_COUNT
};

// host: the cheapest kind that implements `e`
static inline int epilogue_kind(const Epilogue& e) {
  if (e.drop.threshold) return EPK_GENERIC;
  if (e.remap_L > 0) return (e.act == ACT_NONE && e.act_bwd == ACT_NONE) ? EPK_REMAP : EPK_GENERIC;
  if (e.act != ACT_NONE && e.act_bwd != ACT_NONE) return EPK_GENERIC;
  if (e.act == ACT_GELU) return EPK_GELU;
  if (e.act == ACT_RELU) return EPK_RELU;
  if (e.act_bwd == ACT_GELU) return EPK_GELU_BWD;
  if (e.act_bwd == ACT_RELU) return EPK_RELU_BWD;
  if (e.act_bwd == ACT_DERIV) return EPK_MUL_BWD;
  return EPK_PLAIN;
}

// the kind `e` would have without its dropout (the CTA-pair kernel carries dropout as a separate template flag)
static inline int epilogue_kind_nodrop(const Epilogue& e) {
  Epilogue q = e;
  q.drop.threshold = 0;
  return epilogue_kind(q);
}

// Apply the epilogue to NV consecutive columns [col, col+NV) of logical row `row`.
// Caller guarantees row < M and col + NV <= N, NV % 4 == 0 and col % 4 == 0.
// pre_aux / pre_res: side inputs the caller already fetched (software-pipelined epilogues, NV == 4); nullptr = load here.
template <typename AT, int NV, int KIND>
__device__ __forceinline__ void epilogue_apply(const Epilogue& e, float alpha, int row, int col, int N,
                                               float (&v)[NV], const float4* pre_aux = nullptr,
                                               const float4* pre_res = nullptr) {
  constexpr bool GEN = (KIND == EPK_GENERIC);
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col + i));
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (GEN ? (e.act_bwd != ACT_NONE) : (KIND == EPK_GELU_BWD || KIND == EPK_RELU_BWD)) {
    const AT* aux = reinterpret_cast<const AT*>(e.aux) + (size_t)row * N + col;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      const float4 a = pre_aux ? *pre_aux : load4<AT>(aux + i);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float d;
        if (GEN) d = act_bwd(e.act_bwd, av[t]);
        else if (KIND == EPK_GELU_BWD) d = gelu_bwd_fast(av[t]);
        else d = av[t] > 0.0f ? 1.0f : 0.0f;
        v[i + t] *= d;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] *= alpha;
  if (GEN ? (e.act != ACT_NONE) : (KIND == EPK_GELU || KIND == EPK_RELU)) {
    if (e.out_pre) {
      AT* p = reinterpret_cast<AT*>(e.out_pre) + (size_t)row * N + col;
      const int a = GEN ? e.act : (KIND == EPK_GELU ? ACT_GELU : ACT_RELU);
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        if (e.pre_is_deriv)
          store4<AT>(p + i, make_float4(act_bwd(a, v[i]), act_bwd(a, v[i + 1]), act_bwd(a, v[i + 2]),
                                        act_bwd(a, v[i + 3])));
        else
          store4<AT>(p + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (GEN) v[i] = act_fwd(e.act, v[i]);
      else if (KIND == EPK_GELU) v[i] = gelu_fwd_fast(v[i]);
      else v[i] = fmaxf(v[i], 0.0f);
    }
  }
  if (GEN) {
    if (e.drop.threshold) {
      const uint64_t base = (uint64_t)row * (uint64_t)N + (uint64_t)col;
      const uint64_t seed = e.drop.eff();
#pragma unroll
      for (int i = 0; i < NV; ++i)
        v[i] = drop_keep(seed, e.drop.site, base + i, e.drop.threshold) ? v[i] * e.drop.scale : 0.0f;
    }
  }
  int orow = row;
  if (GEN ? (e.remap_L > 0) : (KIND == EPK_REMAP)) {
    const int b = row / e.remap_L;
    const int l = row - b * e.remap_L;
    orow = b * (e.remap_L + 1) + 1 + l;
    if (e.pos) {
      const float* pp = e.pos + (size_t)(1 + l) * N + col;
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(pp + i));
        v[i] += r.x; v[i + 1] += r.y; v[i + 2] += r.z; v[i + 3] += r.w;
      }
    }
  }
  const size_t ooff = (size_t)orow * e.ldo + col;
  if (e.residual) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      const float4 r = pre_res ? *pre_res : *reinterpret_cast<const float4*>(e.residual + ooff + i);
      v[i] += r.x; v[i + 1] += r.y; v[i + 2] += r.z; v[i + 3] += r.w;
    }
  }
  if (e.out) {
    AT* p = reinterpret_cast<AT*>(e.out) + ooff;
#pragma unroll
    for (int i = 0; i < NV; i += 4) store4<AT>(p + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
  }
  if (e.out_f32) {
    float* p = e.out_f32 + ooff;
#pragma unroll
    for (int i = 0; i < NV; i += 4)
      *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
}

// Fetch the side inputs of one 4-column group ahead of time (plain row mapping only: KIND != REMAP/GENERIC-remap).
template <typename AT, int KIND>
__device__ __forceinline__ void epilogue_prefetch(const Epilogue& e, int row, int col, int N, float4& aux, float4& res) {
  constexpr bool GEN = (KIND == EPK_GENERIC);
  if (GEN ? (e.act_bwd != ACT_NONE) : (KIND == EPK_GELU_BWD || KIND == EPK_RELU_BWD))
    aux = load4<AT>(reinterpret_cast<const AT*>(e.aux) + (size_t)row * N + col);
  if (e.residual) res = *reinterpret_cast<const float4*>(e.residual + (size_t)row * e.ldo + col);
}

}  // namespace fervit
