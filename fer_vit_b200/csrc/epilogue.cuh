// The one GEMM epilogue used by every dense contraction on the path (see common.cuh: Epilogue).
#pragma once
#include "common.cuh"

namespace fervit {

// Apply the epilogue to NV consecutive columns [col, col+NV) of logical row `row`.
// Caller guarantees row < M and col + NV <= N, NV % 4 == 0 and col % 4 == 0.
template <typename AT, int NV>
__device__ __forceinline__ void epilogue_apply(const Epilogue& e, float alpha, int row, int col,
                                               int N, float (&v)[NV]) {
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col + i));
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (e.act_bwd != ACT_NONE) {
    const AT* aux = reinterpret_cast<const AT*>(e.aux) + (size_t)row * N + col;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      const float4 a = load4<AT>(aux + i);
      v[i] *= act_bwd(e.act_bwd, a.x);
      v[i + 1] *= act_bwd(e.act_bwd, a.y);
      v[i + 2] *= act_bwd(e.act_bwd, a.z);
      v[i + 3] *= act_bwd(e.act_bwd, a.w);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] *= alpha;
  if (e.out_pre) {
    AT* p = reinterpret_cast<AT*>(e.out_pre) + (size_t)row * N + col;
#pragma unroll
    for (int i = 0; i < NV; i += 4) store4<AT>(p + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
  }
  if (e.act != ACT_NONE) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = act_fwd(e.act, v[i]);
  }
  if (e.drop.threshold) {
    const uint64_t base = (uint64_t)row * (uint64_t)N + (uint64_t)col;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      v[i] = drop_keep(e.drop.eff(), e.drop.site, base + i, e.drop.threshold) ? v[i] * e.drop.scale : 0.0f;
  }
  int orow = row;
  int prow = 0;
  if (e.remap_L > 0) {
    const int b = row / e.remap_L;
    const int l = row - b * e.remap_L;
    orow = b * (e.remap_L + 1) + 1 + l;
    prow = 1 + l;
  }
  const size_t ooff = (size_t)orow * e.ldo + col;
  if (e.residual) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      const float4 r = *reinterpret_cast<const float4*>(e.residual + ooff + i);
      v[i] += r.x; v[i + 1] += r.y; v[i + 2] += r.z; v[i + 3] += r.w;
    }
  }
  if (e.pos) {
    const float* pp = e.pos + (size_t)prow * N + col;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(pp + i));
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

}  // namespace fervit
