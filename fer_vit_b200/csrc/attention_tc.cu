// Short-sequence attention on tensor cores for S <= 32 (19 w+ tokens + cls): one warp owns one (sample, head).
// Q.K^T, P.V and the five backward products are 16x8x16 bf16 MMAs (mma.sync) on fragments loaded straight from
// global memory; softmax runs on the accumulator fragments with quad shuffles; only the operands needed in
// transposed form (V for P.V; K, Q, dO for the backward products) are staged through shared memory.
// The whole problem is 19x64 per operand, far below a tcgen05 tile (M = 128), so the legacy warp-level MMA is the
// right-sized instruction here; the kernel is bound by its HBM/L2 traffic (~9.5 flop/B).
//
// qkv layout: [B*S, 3E], columns [Q | K | V], head h at columns h*HD (timm qkv / torch in_proj packing).
#include "common.cuh"
#include "kernels.h"

namespace fervit {

namespace attn_tc {

constexpr int WARPS = 4;
constexpr int SP = 32;        // padded sequence
constexpr int LDT = SP + 8;   // row stride (elements) of a transposed [HD][SP] operand in smem: conflict-free LDS.32

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// 32-bit load of two consecutive bf16 of row `r` (zero beyond S rows)
__device__ __forceinline__ uint32_t ld2(const bf16* __restrict__ base, size_t row_stride, int r, int c, int S) {
  return r < S ? __ldg(reinterpret_cast<const unsigned int*>(base + (size_t)r * row_stride + c)) : 0u;
}
// A fragment (16x16, row-major source): m-tile mt, k-step ks
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16* __restrict__ base, size_t rs, int mt, int ks, int g,
                                       int q, int S) {
  const int r0 = mt * 16 + g, c0 = ks * 16 + 2 * q;
  a[0] = ld2(base, rs, r0, c0, S);
  a[1] = ld2(base, rs, r0 + 8, c0, S);
  a[2] = ld2(base, rs, r0, c0 + 8, S);
  a[3] = ld2(base, rs, r0 + 8, c0 + 8, S);
}
// B fragment (16x8, "col"): B[k][n] = M[n][k] with M row-major: n-tile nt, k-step ks
__device__ __forceinline__ void load_b(uint32_t (&b)[2], const bf16* __restrict__ base, size_t rs, int nt, int ks, int g,
                                       int q, int S) {
  const int n = nt * 8 + g, c0 = ks * 16 + 2 * q;
  b[0] = ld2(base, rs, n, c0, S);
  b[1] = ld2(base, rs, n, c0 + 8, S);
}
// B fragment from a transposed smem operand T[HD][LDT] (T[d][j] = M[j][d]): B[k=j][n=d], n-tile nt (over d), k-step ks (over j)
__device__ __forceinline__ void load_bt(uint32_t (&b)[2], const bf16* T, int nt, int ks, int g, int q) {
  const bf16* p = T + (nt * 8 + g) * LDT + ks * 16 + 2 * q;
  b[0] = *reinterpret_cast<const uint32_t*>(p);
  b[1] = *reinterpret_cast<const uint32_t*>(p + 8);
}
// stage M[S][HD] (global, row stride rs) transposed into T[HD][LDT], zero-filling columns S..31
template <int HD>
__device__ __forceinline__ void stage_t(bf16* T, const bf16* __restrict__ base, size_t rs, int S, int lane) {
  constexpr int CH = HD / 8;
  for (int i = lane; i < SP * CH; i += 32) {
    const int j = i / CH, d0 = (i % CH) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (j < S) v = __ldg(reinterpret_cast<const uint4*>(base + (size_t)j * rs + d0));
    const bf16* e = reinterpret_cast<const bf16*>(&v);
#pragma unroll
    for (int t = 0; t < 8; ++t) T[(d0 + t) * LDT + j] = e[t];
  }
}

template <int HD>
__global__ void __launch_bounds__(WARPS * 32)
attn_tc_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int B, int S, int H,
                   float scale, Dropout drop) {
  __shared__ __align__(16) bf16 smem_vt[WARPS][HD * LDT];
  constexpr int KS = HD / 16, ND = HD / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int bh = blockIdx.x * WARPS + warp;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh % H;
  const int E = H * HD;
  const size_t rs = (size_t)3 * E;
  const bf16* Qg = qkv + (size_t)b * S * rs + h * HD;
  const bf16* Kg = Qg + E;
  const bf16* Vg = Qg + 2 * E;
  bf16* Vt = smem_vt[warp];
  stage_t<HD>(Vt, Vg, rs, S, lane);
  __syncwarp();
  const uint64_t dseed = drop.threshold ? drop.eff() : 0;
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    if (mt * 16 >= S) break;
    float c[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a[4];
      load_a(a, Qg, rs, mt, ks, g, q, S);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t bb[2];
        load_b(bb, Kg, rs, nt, ks, g, q, S);
        mma16816(c[nt], a, bb);
      }
    }
    // softmax over keys for rows r0 = mt*16+g (elements 0,1) and r1 = r0+8 (elements 2,3)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * q + (e & 1);
        c[nt][e] = col < S ? c[nt][e] * scale : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], c[nt][e]);
      }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        c[nt][e] = __expf(c[nt][e] - mx[e >> 1]);
        sum[e >> 1] += c[nt][e];
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
    }
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const float inv[2] = {1.0f / sum[0], 1.0f / sum[1]};
    if (lse && q == 0) {
      if (r0 < S) lse[(size_t)bh * S + r0] = mx[0] + __logf(sum[0]);
      if (r1 < S) lse[(size_t)bh * S + r1] = mx[1] + __logf(sum[1]);
    }
    if (drop.threshold) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = (e >> 1) ? r1 : r0, col = nt * 8 + 2 * q + (e & 1);
          const uint64_t idx = ((uint64_t)bh * S + row) * S + col;
          c[nt][e] = drop_keep(dseed, drop.site, idx, drop.threshold) ? c[nt][e] * drop.scale : 0.f;
        }
    }
    // O = P V: the accumulator fragments of P are exactly the A fragments of the next MMA
    float o[ND][4];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[nd][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16x2(c[2 * kk][0] * inv[0], c[2 * kk][1] * inv[0]);
      a[1] = pack_bf16x2(c[2 * kk][2] * inv[1], c[2 * kk][3] * inv[1]);
      a[2] = pack_bf16x2(c[2 * kk + 1][0] * inv[0], c[2 * kk + 1][1] * inv[0]);
      a[3] = pack_bf16x2(c[2 * kk + 1][2] * inv[1], c[2 * kk + 1][3] * inv[1]);
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        uint32_t bb[2];
        load_bt(bb, Vt, nd, kk, g, q);
        mma16816(o[nd], a, bb);
      }
    }
    bf16* Og = out + (size_t)b * S * E + h * HD;
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      const int col = nd * 8 + 2 * q;
      if (r0 < S) *reinterpret_cast<uint32_t*>(Og + (size_t)r0 * E + col) = pack_bf16x2(o[nd][0], o[nd][1]);
      if (r1 < S) *reinterpret_cast<uint32_t*>(Og + (size_t)r1 * E + col) = pack_bf16x2(o[nd][2], o[nd][3]);
    }
  }
}

// Backward. Phase 1 (query-major): S = Q K^T, dP = dO V^T -> dS -> dQ = dS K.
//           Phase 2 (key-major):   S^T = K Q^T, dP^T = V dO^T -> P^T, dS^T -> dV = P^T dO, dK = dS^T Q.
template <int HD>
__global__ void __launch_bounds__(WARPS * 32)
attn_tc_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ dout,
                   const float* __restrict__ lse, bf16* __restrict__ dqkv, int B, int S, int H, float scale,
                   Dropout drop) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int KS = HD / 16, ND = HD / 8;
  constexpr int PER_WARP = 3 * HD * LDT * 2 + 2 * SP * 4;   // Kt, Qt, dOt (bf16) + LSE, D (fp32)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int bh = blockIdx.x * WARPS + warp;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh % H;
  const int E = H * HD;
  const size_t rs = (size_t)3 * E;
  const bf16* Qg = qkv + (size_t)b * S * rs + h * HD;
  const bf16* Kg = Qg + E;
  const bf16* Vg = Qg + 2 * E;
  const bf16* Og = out + (size_t)b * S * E + h * HD;
  const bf16* dOg = dout + (size_t)b * S * E + h * HD;
  bf16* Kt = reinterpret_cast<bf16*>(smem_raw + (size_t)warp * PER_WARP);
  bf16* Qt = Kt + HD * LDT;
  bf16* dOt = Qt + HD * LDT;
  float* Ls = reinterpret_cast<float*>(dOt + HD * LDT);
  float* Ds = Ls + SP;
  stage_t<HD>(Kt, Kg, rs, S, lane);
  stage_t<HD>(Qt, Qg, rs, S, lane);
  stage_t<HD>(dOt, dOg, (size_t)E, S, lane);
  {
    // D_i = dO_i . O_i ; rows beyond S get LSE = +inf so their probabilities vanish
    float dsum = 0.f;
    if (lane < S) {
#pragma unroll
      for (int d0 = 0; d0 < HD; d0 += 8) {
        const uint4 ov = __ldg(reinterpret_cast<const uint4*>(Og + (size_t)lane * E + d0));
        const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dOg + (size_t)lane * E + d0));
        const uint32_t* op = reinterpret_cast<const uint32_t*>(&ov);
        const uint32_t* dp = reinterpret_cast<const uint32_t*>(&dv);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 a = unpack_bf16x2(op[t]), bb = unpack_bf16x2(dp[t]);
          dsum += a.x * bb.x + a.y * bb.y;
        }
      }
    }
    Ds[lane] = dsum;
    Ls[lane] = lane < S ? lse[(size_t)bh * S + lane] : INFINITY;
  }
  __syncwarp();
  const uint64_t dseed = drop.threshold ? drop.eff() : 0;
  bf16* dQg = dqkv + (size_t)b * S * rs + h * HD;
  bf16* dKg = dQg + E;
  bf16* dVg = dQg + 2 * E;

  // ---------------- phase 1: dQ ----------------
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    if (mt * 16 >= S) break;
    float c[4][4], dp[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { c[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t aq[4], ad[4];
      load_a(aq, Qg, rs, mt, ks, g, q, S);
      load_a(ad, dOg, (size_t)E, mt, ks, g, q, S);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t bk[2], bv[2];
        load_b(bk, Kg, rs, nt, ks, g, q, S);
        load_b(bv, Vg, rs, nt, ks, g, q, S);
        mma16816(c[nt], aq, bk);
        mma16816(dp[nt], ad, bv);
      }
    }
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const float l0 = Ls[r0], l1 = Ls[r1], d0v = Ds[r0], d1v = Ds[r1];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = (e >> 1) ? r1 : r0, col = nt * 8 + 2 * q + (e & 1);
        const float p = col < S ? __expf(c[nt][e] * scale - ((e >> 1) ? l1 : l0)) : 0.f;
        float dpv = dp[nt][e];
        if (drop.threshold) {
          const uint64_t idx = ((uint64_t)bh * S + row) * S + col;
          dpv = drop_keep(dseed, drop.site, idx, drop.threshold) ? dpv * drop.scale : 0.f;
        }
        c[nt][e] = p * (dpv - ((e >> 1) ? d1v : d0v)) * scale;   // dS
      }
    float dq[ND][4];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[nd][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16x2(c[2 * kk][0], c[2 * kk][1]);
      a[1] = pack_bf16x2(c[2 * kk][2], c[2 * kk][3]);
      a[2] = pack_bf16x2(c[2 * kk + 1][0], c[2 * kk + 1][1]);
      a[3] = pack_bf16x2(c[2 * kk + 1][2], c[2 * kk + 1][3]);
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        uint32_t bb[2];
        load_bt(bb, Kt, nd, kk, g, q);
        mma16816(dq[nd], a, bb);
      }
    }
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      const int col = nd * 8 + 2 * q;
      if (r0 < S) *reinterpret_cast<uint32_t*>(dQg + (size_t)r0 * rs + col) = pack_bf16x2(dq[nd][0], dq[nd][1]);
      if (r1 < S) *reinterpret_cast<uint32_t*>(dQg + (size_t)r1 * rs + col) = pack_bf16x2(dq[nd][2], dq[nd][3]);
    }
  }

  // ---------------- phase 2: dK, dV (rows are keys j, columns are queries i) ----------------
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    if (mt * 16 >= S) break;
    float c[4][4], dp[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { c[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t ak[4], av[4];
      load_a(ak, Kg, rs, mt, ks, g, q, S);
      load_a(av, Vg, rs, mt, ks, g, q, S);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t bq[2], bd[2];
        load_b(bq, Qg, rs, nt, ks, g, q, S);
        load_b(bd, dOg, (size_t)E, nt, ks, g, q, S);
        mma16816(c[nt], ak, bq);
        mma16816(dp[nt], av, bd);
      }
    }
    const int j0 = mt * 16 + g, j1 = j0 + 8;
    float pt[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = (e >> 1) ? j1 : j0, i = nt * 8 + 2 * q + (e & 1);
        const float p = (j < S) ? __expf(c[nt][e] * scale - Ls[i]) : 0.f;   // Ls[i >= S] = +inf -> 0
        float ptv = p, dpv = dp[nt][e];
        if (drop.threshold) {
          const uint64_t idx = ((uint64_t)bh * S + i) * S + j;
          const bool keep = drop_keep(dseed, drop.site, idx, drop.threshold);
          ptv = keep ? p * drop.scale : 0.f;
          dpv = keep ? dpv * drop.scale : 0.f;
        }
        pt[nt][e] = ptv;                                   // P~^T
        c[nt][e] = p * (dpv - Ds[i]) * scale;              // dS^T
      }
    float dv[ND][4], dk[ND][4];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) { dv[nd][e] = 0.f; dk[nd][e] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t ap[4], as[4];
      ap[0] = pack_bf16x2(pt[2 * kk][0], pt[2 * kk][1]);
      ap[1] = pack_bf16x2(pt[2 * kk][2], pt[2 * kk][3]);
      ap[2] = pack_bf16x2(pt[2 * kk + 1][0], pt[2 * kk + 1][1]);
      ap[3] = pack_bf16x2(pt[2 * kk + 1][2], pt[2 * kk + 1][3]);
      as[0] = pack_bf16x2(c[2 * kk][0], c[2 * kk][1]);
      as[1] = pack_bf16x2(c[2 * kk][2], c[2 * kk][3]);
      as[2] = pack_bf16x2(c[2 * kk + 1][0], c[2 * kk + 1][1]);
      as[3] = pack_bf16x2(c[2 * kk + 1][2], c[2 * kk + 1][3]);
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        uint32_t bo[2], bq[2];
        load_bt(bo, dOt, nd, kk, g, q);
        load_bt(bq, Qt, nd, kk, g, q);
        mma16816(dv[nd], ap, bo);
        mma16816(dk[nd], as, bq);
      }
    }
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      const int col = nd * 8 + 2 * q;
      if (j0 < S) {
        *reinterpret_cast<uint32_t*>(dKg + (size_t)j0 * rs + col) = pack_bf16x2(dk[nd][0], dk[nd][1]);
        *reinterpret_cast<uint32_t*>(dVg + (size_t)j0 * rs + col) = pack_bf16x2(dv[nd][0], dv[nd][1]);
      }
      if (j1 < S) {
        *reinterpret_cast<uint32_t*>(dKg + (size_t)j1 * rs + col) = pack_bf16x2(dk[nd][2], dk[nd][3]);
        *reinterpret_cast<uint32_t*>(dVg + (size_t)j1 * rs + col) = pack_bf16x2(dv[nd][2], dv[nd][3]);
      }
    }
  }
}

template <int HD>
int launch_fwd(const bf16* qkv, bf16* out, float* lse, int B, int S, int H, Dropout drop, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)HD);
  ProfScope prof(1, (double)B * S * H * HD * 4.0 * sizeof(bf16) + (double)B * H * S * 4.0, stream);
  attn_tc_fwd_kernel<HD><<<ceil_div(B * H, WARPS), WARPS * 32, 0, stream>>>(qkv, out, lse, B, S, H, scale, drop);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template <int HD>
int launch_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int B, int S, int H,
               Dropout drop, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)HD);
  constexpr int smem = WARPS * (3 * HD * LDT * 2 + 2 * SP * 4);
  static bool attr = false;
  if (!attr && smem > 48 * 1024) {
    FV_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  ProfScope prof(1, (double)B * S * H * HD * 8.0 * sizeof(bf16) + (double)B * H * S * 4.0, stream);
  attn_tc_bwd_kernel<HD><<<ceil_div(B * H, WARPS), WARPS * 32, smem, stream>>>(qkv, out, dout, lse, dqkv, B, S, H,
                                                                              scale, drop);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

}  // namespace attn_tc

bool attention_tc_supported(int S, int HD) { return S <= attn_tc::SP && (HD == 64 || HD == 48 || HD == 32); }

int attention_tc_fwd(const bf16* qkv, bf16* out, float* lse, int B, int S, int H, int HD, Dropout drop,
                     cudaStream_t stream) {
  if (HD == 64) return attn_tc::launch_fwd<64>(qkv, out, lse, B, S, H, drop, stream);
  if (HD == 48) return attn_tc::launch_fwd<48>(qkv, out, lse, B, S, H, drop, stream);
  if (HD == 32) return attn_tc::launch_fwd<32>(qkv, out, lse, B, S, H, drop, stream);
  FV_CHECK(false, "attention_tc: head dim %d not supported", HD);
}
int attention_tc_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int B, int S,
                     int H, int HD, Dropout drop, cudaStream_t stream) {
  if (HD == 64) return attn_tc::launch_bwd<64>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
  if (HD == 48) return attn_tc::launch_bwd<48>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
  if (HD == 32) return attn_tc::launch_bwd<32>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
  FV_CHECK(false, "attention_tc: head dim %d not supported", HD);
}

}  // namespace fervit
