// Warp-level tensor-core fragment helpers shared by the attention kernels (attention_tc.cu: S <= 32;
// attention_tc_long.cu: 32 < S <= 256): mma.sync m16n8k16 bf16 with ldmatrix(.trans) operand loads from shared memory.
#pragma once
#include "common.cuh"

namespace fervit {
namespace attn_frag {

template <int HD> struct Lay { static constexpr int LD = HD + 8; };  // row stride (elements): conflict-free ldmatrix

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
// A fragment of rows [mt*16, +16), k in [ks*16, +16) of a row-major smem matrix M[row][k]
template <int LD>
__device__ __forceinline__ void lda(uint32_t (&a)[4], const bf16* M, int mt, int ks, int lane) {
  ldsm_x4(a, M + (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + ks * 16 + (lane >> 4) * 8);
}
// B fragments of TWO n-tiles (n in [np*16, +16)), k in [ks*16, +16), from M[n][k] (k contiguous): r[0..1] tile 2np, r[2..3] tile 2np+1
template <int LD>
__device__ __forceinline__ void ldb(uint32_t (&r)[4], const bf16* M, int np, int ks, int lane) {
  ldsm_x4(r, M + (np * 16 + (lane & 7) + (lane >> 4) * 8) * LD + ks * 16 + ((lane >> 3) & 1) * 8);
}
// B fragments of TWO n-tiles from M[k][n] (n contiguous): k in [kk*16, +16), n in [np*16, +16)
template <int LD>
__device__ __forceinline__ void ldbt(uint32_t (&r)[4], const bf16* M, int np, int kk, int lane) {
  ldsm_x4_t(r, M + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + np * 16 + (lane >> 4) * 8);
}
// A fragment of the TRANSPOSE: rows m in [mt*16, +16), k in [kk*16, +16), from a matrix stored as M[k][m] (m contiguous)
template <int LD>
__device__ __forceinline__ void lda_t(uint32_t (&a)[4], const bf16* M, int mt, int kk, int lane) {
  ldsm_x4_t(a, M + (kk * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * LD + mt * 16 + ((lane >> 3) & 1) * 8);
}

}  // namespace attn_frag
}  // namespace fervit
