// Short-sequence attention on tensor cores for S <= 32 (19 w+ tokens + cls): two warps own one (sample, head), one per
// 16-row tile, sharing the staged operands.
// Q, K, V (and dO) of the head are staged ONCE into shared memory with coalesced 16-byte loads (rows beyond S
// zero-filled); every MMA operand — including the transposed ones (V for P.V; K, Q, dO for the backward products) —
// is then fetched with ldmatrix / ldmatrix.trans, so nothing is stored transposed. Q.K^T, P.V and the five backward
// products are 16x8x16 bf16 MMAs (mma.sync); softmax runs on the accumulator fragments with quad shuffles.
// A 19x64 problem is far below a tcgen05 tile (M = 128), so the warp-level MMA is the right-sized instruction; the
// kernel is bound by its HBM/L2 traffic (~9.5 flop/B).
//
// qkv layout: [B*S, 3E], columns [Q | K | V], head h at columns h*HD (timm qkv / torch in_proj packing).
#include "common.cuh"
#include "kernels.h"
#include "attn_frag.cuh"
#include <stdlib.h>

namespace fervit {

namespace attn_tc {

constexpr int HEADS = 4;            // (sample, head) problems per CTA
constexpr int WARPS = 2 * HEADS;    // two warps per problem: one per 16-row tile of queries (phase 1) / keys (phase 2)
constexpr int SP = 32;  // padded sequence

using namespace attn_frag;

// stage M[S][HD] (global, row stride rs) into smem [32][LD], zero-filling rows S..31. Asynchronous 16-byte copies
// (cp.async, L2 -> smem without a register round trip): all ~5 copies per lane and matrix are in flight at once, where
// a load-then-store loop exposed one global-memory latency per unrolled pair. Caller: stage_wait() before reading.
template <int HD>
__device__ __forceinline__ void stage(bf16* dst, const bf16* __restrict__ src, size_t rs, int S, int t64) {
  constexpr int CH = HD / 8, LD = Lay<HD>::LD;
#pragma unroll
  for (int i = t64; i < SP * CH; i += 64) {
    const int r = i / CH, c = (i % CH) * 8;
    if (r < S) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst + r * LD + c)),
                   "l"(src + (size_t)r * rs + c)
                   : "memory");
    } else {
      *reinterpret_cast<uint4*>(dst + r * LD + c) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}
// both warps of a problem have issued their copies: wait for one's own, then meet the partner (named barrier 1 + pair)
__device__ __forceinline__ void stage_wait(int pair) {
  asm volatile("cp.async.wait_all;" ::: "memory");
  asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
}

template <int HD>
__global__ void __launch_bounds__(WARPS * 32)
attn_tc_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int B, int S, int H,
                   float scale, Dropout drop) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int KS = HD / 16, ND = HD / 8, LD = Lay<HD>::LD;
  constexpr int MAT = SP * LD;  // elements per staged matrix
  pdl_trigger();
  pdl_grid_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int pair = warp >> 1, wm = warp & 1, t64 = threadIdx.x & 63;
  const int bh = blockIdx.x * HEADS + pair;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh % H;
  const int E = H * HD;
  const size_t rs = (size_t)3 * E;
  const bf16* Qg = qkv + (size_t)b * S * rs + h * HD;
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw) + (size_t)pair * 3 * MAT;
  bf16* Ks = Qs + MAT;
  bf16* Vs = Ks + MAT;
  stage<HD>(Qs, Qg, rs, S, t64);
  stage<HD>(Ks, Qg + E, rs, S, t64);
  stage<HD>(Vs, Qg + 2 * E, rs, S, t64);
  stage_wait(pair);
  const uint64_t dseed = drop.threshold ? drop.eff() : 0;
  // this warp's 16 query rows (the second warp of a 19-token problem carries 3 live rows: cheap, and in parallel)
#pragma unroll 1
  for (int mt = wm, once = 0; once < 1 && mt * 16 < S; ++once) {
    float c[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a[4];
      lda<LD>(a, Qs, mt, ks, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bb[4];
        ldb<LD>(bb, Ks, np, ks, lane);
        mma16816(c[2 * np], a, bb[0], bb[1]);
        if ((2 * np + 1) * 8 < S) mma16816(c[2 * np + 1], a, bb[2], bb[3]);   // skip all-padding key tiles
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
      for (int np = 0; np < ND / 2; ++np) {
        uint32_t bb[4];
        ldbt<LD>(bb, Vs, np, kk, lane);
        mma16816(o[2 * np], a, bb[0], bb[1]);
        mma16816(o[2 * np + 1], a, bb[2], bb[3]);
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

// Backward. Phase 1 (query-major, one 16-query tile per warp): S = Q K^T, dP = dO V^T -> P~, dS -> dQ = dS K; P~ and dS
//           are also parked in shared memory (bf16, [query][key]).
//           Phase 2 (key-major, one 16-key tile per warp): dV = P~^T dO, dK = dS^T Q with the A operands fetched
//           TRANSPOSED from the parked tiles (ldmatrix.trans) instead of recomputing S^T and dP^T.
template <int HD>
__global__ void __launch_bounds__(WARPS * 32)
attn_tc_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ dout,
                   const float* __restrict__ lse, bf16* __restrict__ dqkv, int B, int S, int H, float scale,
                   Dropout drop) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int KS = HD / 16, ND = HD / 8, LD = Lay<HD>::LD;
  constexpr int MAT = SP * LD;
  constexpr int LDP = SP + 8;                          // row stride of the parked P~ / dS tiles (elements)
  // per problem: Q, K, V, dO (bf16) + LSE, D (fp32) + P~, dS (bf16), bytes
  constexpr int PER_WARP = 4 * MAT * 2 + 2 * SP * 4 + 2 * SP * LDP * 2;
  pdl_trigger();
  pdl_grid_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int pair = warp >> 1, wm = warp & 1, t64 = threadIdx.x & 63;
  const int bh = blockIdx.x * HEADS + pair;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh % H;
  const int E = H * HD;
  const size_t rs = (size_t)3 * E;
  const bf16* Qg = qkv + (size_t)b * S * rs + h * HD;
  const bf16* Og = out + (size_t)b * S * E + h * HD;
  const bf16* dOg = dout + (size_t)b * S * E + h * HD;
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw + (size_t)pair * PER_WARP);
  bf16* Ks = Qs + MAT;
  bf16* Vs = Ks + MAT;
  bf16* dOs = Vs + MAT;
  float* Ls = reinterpret_cast<float*>(dOs + MAT);
  float* Ds = Ls + SP;
  bf16* Ps = reinterpret_cast<bf16*>(Ds + SP);
  bf16* dSs = Ps + SP * LDP;
  stage<HD>(Qs, Qg, rs, S, t64);
  stage<HD>(Ks, Qg + E, rs, S, t64);
  stage<HD>(Vs, Qg + 2 * E, rs, S, t64);
  stage<HD>(dOs, dOg, (size_t)E, S, t64);
  if (wm == 0) {
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
  stage_wait(pair);
  const uint64_t dseed = drop.threshold ? drop.eff() : 0;
  bf16* dQg = dqkv + (size_t)b * S * rs + h * HD;
  bf16* dKg = dQg + E;
  bf16* dVg = dQg + 2 * E;

  // ---------------- phase 1: dQ ----------------
#pragma unroll 1
  for (int mt = wm, once = 0; once < 1 && mt * 16 < S; ++once) {
    float c[4][4], dp[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { c[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t aq[4], ad[4];
      lda<LD>(aq, Qs, mt, ks, lane);
      lda<LD>(ad, dOs, mt, ks, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bk[4], bv[4];
        ldb<LD>(bk, Ks, np, ks, lane);
        ldb<LD>(bv, Vs, np, ks, lane);
        mma16816(c[2 * np], aq, bk[0], bk[1]);
        mma16816(dp[2 * np], ad, bv[0], bv[1]);
        if ((2 * np + 1) * 8 < S) {   // skip all-padding key tiles
          mma16816(c[2 * np + 1], aq, bk[2], bk[3]);
          mma16816(dp[2 * np + 1], ad, bv[2], bv[3]);
        }
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
        float dpv = dp[nt][e], ptv = p;
        if (drop.threshold) {
          const uint64_t idx = ((uint64_t)bh * S + row) * S + col;
          const bool keep = drop_keep(dseed, drop.site, idx, drop.threshold);
          dpv = keep ? dpv * drop.scale : 0.f;
          ptv = keep ? p * drop.scale : 0.f;
        }
        c[nt][e] = p * (dpv - ((e >> 1) ? d1v : d0v)) * scale;   // dS
        dp[nt][e] = ptv;                                         // P~ (dropout applied), reused below
      }
    // park P~ and dS for phase 2 (rows beyond S are zero: their LSE is +inf)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = nt * 8 + 2 * q;
      *reinterpret_cast<uint32_t*>(Ps + r0 * LDP + col) = pack_bf16x2(dp[nt][0], dp[nt][1]);
      *reinterpret_cast<uint32_t*>(Ps + r1 * LDP + col) = pack_bf16x2(dp[nt][2], dp[nt][3]);
      *reinterpret_cast<uint32_t*>(dSs + r0 * LDP + col) = pack_bf16x2(c[nt][0], c[nt][1]);
      *reinterpret_cast<uint32_t*>(dSs + r1 * LDP + col) = pack_bf16x2(c[nt][2], c[nt][3]);
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
      for (int np = 0; np < ND / 2; ++np) {
        uint32_t bb[4];
        ldbt<LD>(bb, Ks, np, kk, lane);
        mma16816(dq[2 * np], a, bb[0], bb[1]);
        mma16816(dq[2 * np + 1], a, bb[2], bb[3]);
      }
    }
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      const int col = nd * 8 + 2 * q;
      if (r0 < S) *reinterpret_cast<uint32_t*>(dQg + (size_t)r0 * rs + col) = pack_bf16x2(dq[nd][0], dq[nd][1]);
      if (r1 < S) *reinterpret_cast<uint32_t*>(dQg + (size_t)r1 * rs + col) = pack_bf16x2(dq[nd][2], dq[nd][3]);
    }
  }

  // ---------------- phase 2: dK, dV (rows are keys j; the reduction runs over queries i) ----------------
  if (wm * 16 >= S) {
    // this warp had no query tile: its half of the parked tiles is still unwritten
#pragma unroll
    for (int i = lane; i < 16 * (LDP / 2); i += 32) {
      reinterpret_cast<uint32_t*>(Ps + 16 * LDP)[i] = 0u;
      reinterpret_cast<uint32_t*>(dSs + 16 * LDP)[i] = 0u;
    }
  }
  asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");   // both query tiles are parked
#pragma unroll 1
  for (int mt = wm, once = 0; once < 1 && mt * 16 < S; ++once) {
    const int j0 = mt * 16 + g, j1 = j0 + 8;
    float dv[ND][4], dk[ND][4];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) { dv[nd][e] = 0.f; dk[nd][e] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      if (kk * 16 >= S) break;   // all-padding query block
      uint32_t ap[4], as[4];
      lda_t<LDP>(ap, Ps, mt, kk, lane);    // P~^T  [keys x queries]
      lda_t<LDP>(as, dSs, mt, kk, lane);   // dS^T
#pragma unroll
      for (int np = 0; np < ND / 2; ++np) {
        uint32_t bo[4], bq[4];
        ldbt<LD>(bo, dOs, np, kk, lane);
        ldbt<LD>(bq, Qs, np, kk, lane);
        mma16816(dv[2 * np], ap, bo[0], bo[1]);
        mma16816(dv[2 * np + 1], ap, bo[2], bo[3]);
        mma16816(dk[2 * np], as, bq[0], bq[1]);
        mma16816(dk[2 * np + 1], as, bq[2], bq[3]);
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

// =================================================================================================
// Version 2 (default): persistent, double-buffered, compact tiles, coalesced stores.
//
// What ncu said about version 1 at batch 256 (profiles/r01_ncu_full_attn_bwd_s19.txt): 2 CTAs of 8 warps per SM, 21 % of
// peak warps active, DRAM 17 % busy — every (sample, head) problem paid its global-load latency, its barrier and its
// scattered 4-byte stores in series with nothing to overlap them. Version 2 keeps two warps per problem but
//   * makes the CTA persistent: a warp pair walks problems p, p + stride, ... and the cp.async loads of the NEXT
//     problem are issued before the current one is computed (two operand buffers per pair);
//   * stores only the S live token rows of each operand tile (20 instead of 32 rows at S = 19): ldmatrix row addresses
//     of the padding tokens point at one shared all-zero row, so the tiles shrink by 37 % and the second buffer fits at
//     the same occupancy;
//   * (backward) derives D_i = sum_j P_ij dP_ij from the accumulator fragments instead of re-reading O: one operand
//     less to move (dO . O and sum_j P dP are the same number) and no per-lane row loads in front of the barrier;
//   * stages results in the dead operand tiles and writes them with 16-byte row-contiguous stores.
// =================================================================================================
__device__ __forceinline__ void pair_bar(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory"); }

template <int LD>
__device__ __forceinline__ const bf16* tok_row(const bf16* M, const bf16* Z, int tok, int S) {
  return tok < S ? M + tok * LD : Z;
}
// the attn_frag loaders with token rows >= S redirected to the zero row Z
template <int LD>
__device__ __forceinline__ void lda_c(uint32_t (&a)[4], const bf16* M, const bf16* Z, int S, int mt, int ks, int lane) {
  ldsm_x4(a, tok_row<LD>(M, Z, mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, S) + ks * 16 + (lane >> 4) * 8);
}
template <int LD>
__device__ __forceinline__ void ldb_c(uint32_t (&r)[4], const bf16* M, const bf16* Z, int S, int np, int ks, int lane) {
  ldsm_x4(r, tok_row<LD>(M, Z, np * 16 + (lane & 7) + (lane >> 4) * 8, S) + ks * 16 + ((lane >> 3) & 1) * 8);
}
template <int LD>
__device__ __forceinline__ void ldbt_c(uint32_t (&r)[4], const bf16* M, const bf16* Z, int S, int np, int kk, int lane) {
  ldsm_x4_t(r, tok_row<LD>(M, Z, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, S) + np * 16 + (lane >> 4) * 8);
}
template <int LD>
__device__ __forceinline__ void lda_t_c(uint32_t (&a)[4], const bf16* M, const bf16* Z, int S, int mt, int kk, int lane) {
  ldsm_x4_t(a, tok_row<LD>(M, Z, kk * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, S) + mt * 16 + ((lane >> 3) & 1) * 8);
}
// live rows only
template <int HD>
__device__ __forceinline__ void stage_rows(bf16* dst, const bf16* __restrict__ src, size_t rs, int S, int t64) {
  constexpr int CH = HD / 8, LD = Lay<HD>::LD;
  for (int i = t64; i < S * CH; i += 64) {
    const int r = i / CH, c = (i % CH) * 8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst + r * LD + c)),
                 "l"(src + (size_t)r * rs + c)
                 : "memory");
  }
}
// staged tile (live rows) -> global rows of stride rs, 16 bytes per lane, 8 lanes per 128-byte row
template <int HD>
__device__ __forceinline__ void copy_out(bf16* __restrict__ dst, size_t rs, const bf16* src, int S, int t64) {
  constexpr int CH = HD / 8, LD = Lay<HD>::LD;
  for (int i = t64; i < S * CH; i += 64) {
    const int r = i / CH, c = (i % CH) * 8;
    *reinterpret_cast<uint4*>(dst + (size_t)r * rs + c) = *reinterpret_cast<const uint4*>(src + r * LD + c);
  }
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int HD>
__global__ void __launch_bounds__(WARPS * 32, 3)
attn_tc_fwd2_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int B, int S, int H,
                    float scale, Dropout drop, int SA) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int KS = HD / 16, ND = HD / 8, LD = Lay<HD>::LD;
  const int MAT = SA * LD;
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int pair = warp >> 1, wm = warp & 1, t64 = threadIdx.x & 63;
  bf16* Z = reinterpret_cast<bf16*>(smem_raw);
  bf16* base = Z + LD + (size_t)pair * 6 * MAT;             // [2 buffers][Q, K, V]
  if (threadIdx.x < LD / 2) reinterpret_cast<uint32_t*>(Z)[threadIdx.x] = 0u;
  __syncthreads();
  pdl_grid_sync();
  const int BH = B * H, E = H * HD, stride = gridDim.x * HEADS;
  const size_t rs = (size_t)3 * E;
  int p = blockIdx.x * HEADS + pair;
  auto issue = [&](int prob, int bi) {
    const bf16* Qg = qkv + (size_t)(prob / H) * S * rs + (prob % H) * HD;
    bf16* d = base + (size_t)bi * 3 * MAT;
    stage_rows<HD>(d, Qg, rs, S, t64);
    stage_rows<HD>(d + MAT, Qg + E, rs, S, t64);
    stage_rows<HD>(d + 2 * MAT, Qg + 2 * E, rs, S, t64);
  };
  if (p < BH) issue(p, 0);
  cp_commit();
  const uint64_t dseed = drop.threshold ? drop.eff() : 0;
  const int mt = wm;
#pragma unroll 1
  for (int cur = 0; p < BH; p += stride, cur ^= 1) {
    if (p + stride < BH) issue(p + stride, cur ^ 1);
    cp_commit();
    cp_wait_1();
    pair_bar(pair);
    const bf16* Qs = base + (size_t)cur * 3 * MAT;
    const bf16* Ks = Qs + MAT;
    const bf16* Vs = Ks + MAT;
    const int bh = p, b = p / H, h = p % H;
    uint32_t opk[ND][2];
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    if (mt * 16 < S) {
      float c[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t a[4];
        lda_c<LD>(a, Qs, Z, S, mt, ks, lane);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t bb[4];
          ldb_c<LD>(bb, Ks, Z, S, np, ks, lane);
          mma16816(c[2 * np], a, bb[0], bb[1]);
          if ((2 * np + 1) * 8 < S) mma16816(c[2 * np + 1], a, bb[2], bb[3]);   // skip all-padding key tiles
        }
      }
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
        for (int np = 0; np < ND / 2; ++np) {
          uint32_t bb[4];
          ldbt_c<LD>(bb, Vs, Z, S, np, kk, lane);
          mma16816(o[2 * np], a, bb[0], bb[1]);
          mma16816(o[2 * np + 1], a, bb[2], bb[3]);
        }
      }
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        opk[nd][0] = pack_bf16x2(o[nd][0], o[nd][1]);
        opk[nd][1] = pack_bf16x2(o[nd][2], o[nd][3]);
      }
    }
    pair_bar(pair);                                   // both warps are done reading Q, K, V of this buffer
    bf16* Os = const_cast<bf16*>(Qs);                 // O rows are staged where Q was
    if (mt * 16 < S) {
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        const int col = nd * 8 + 2 * q;
        if (r0 < S) *reinterpret_cast<uint32_t*>(Os + r0 * LD + col) = opk[nd][0];
        if (r1 < S) *reinterpret_cast<uint32_t*>(Os + r1 * LD + col) = opk[nd][1];
      }
    }
    pair_bar(pair);
    copy_out<HD>(out + (size_t)b * S * E + h * HD, (size_t)E, Os, S, t64);
    pair_bar(pair);                                   // the staged rows have been read: the buffer may be refilled
  }
}

template <int HD>
__global__ void __launch_bounds__(WARPS * 32, 2)
attn_tc_bwd2_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
                    bf16* __restrict__ dqkv, int B, int S, int H, float scale, Dropout drop, int SA) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int KS = HD / 16, ND = HD / 8, LD = Lay<HD>::LD;
  constexpr int LDP = SP + 8;                          // row stride of the parked P~ / dS tiles (elements)
  const int MAT = SA * LD;
  const int BUF = 4 * MAT + 2 * SP;                    // elements: Q, K, V, dO + 32 fp32 LSE values
  const int PER_PAIR = 2 * BUF + 2 * SA * LDP;
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int pair = warp >> 1, wm = warp & 1, t64 = threadIdx.x & 63;
  bf16* Z = reinterpret_cast<bf16*>(smem_raw);
  bf16* base = Z + LD + (size_t)pair * PER_PAIR;
  bf16* Ps = base + 2 * BUF;
  bf16* dSs = Ps + SA * LDP;
  if (threadIdx.x < LD / 2) reinterpret_cast<uint32_t*>(Z)[threadIdx.x] = 0u;
  if (t64 < SP) {                                      // LSE of padding rows = +inf: their probabilities vanish
    reinterpret_cast<float*>(base + 4 * MAT)[t64] = INFINITY;
    reinterpret_cast<float*>(base + BUF + 4 * MAT)[t64] = INFINITY;
  }
  __syncthreads();
  pdl_grid_sync();
  const int BH = B * H, E = H * HD, stride = gridDim.x * HEADS;
  const size_t rs = (size_t)3 * E;
  int p = blockIdx.x * HEADS + pair;
  auto issue = [&](int prob, int bi) {
    const int b = prob / H, h = prob % H;
    const bf16* Qg = qkv + (size_t)b * S * rs + h * HD;
    bf16* d = base + (size_t)bi * BUF;
    stage_rows<HD>(d, Qg, rs, S, t64);
    stage_rows<HD>(d + MAT, Qg + E, rs, S, t64);
    stage_rows<HD>(d + 2 * MAT, Qg + 2 * E, rs, S, t64);
    stage_rows<HD>(d + 3 * MAT, dout + (size_t)b * S * E + h * HD, (size_t)E, S, t64);
    if (t64 < S)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(reinterpret_cast<float*>(d + 4 * MAT) + t64)),
                   "l"(lse + (size_t)prob * S + t64)
                   : "memory");
  };
  if (p < BH) issue(p, 0);
  cp_commit();
  const uint64_t dseed = drop.threshold ? drop.eff() : 0;
  const int mt = wm;
  const bool live = mt * 16 < S;
#pragma unroll 1
  for (int cur = 0; p < BH; p += stride, cur ^= 1) {
    if (p + stride < BH) issue(p + stride, cur ^ 1);
    cp_commit();
    cp_wait_1();
    pair_bar(pair);
    bf16* Qs = base + (size_t)cur * BUF;
    bf16* Ks = Qs + MAT;
    bf16* Vs = Ks + MAT;
    bf16* dOs = Vs + MAT;
    const float* Ls = reinterpret_cast<const float*>(dOs + MAT);
    const int bh = p, b = p / H, h = p % H;
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    // ---------------- phase 1: S = Q K^T, dP = dO V^T -> P~, dS (parked, bf16, [query][key]) ----------------
    if (live) {
      float c[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { c[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t aq[4], ad[4];
        lda_c<LD>(aq, Qs, Z, S, mt, ks, lane);
        lda_c<LD>(ad, dOs, Z, S, mt, ks, lane);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t bk[4], bv[4];
          ldb_c<LD>(bk, Ks, Z, S, np, ks, lane);
          ldb_c<LD>(bv, Vs, Z, S, np, ks, lane);
          mma16816(c[2 * np], aq, bk[0], bk[1]);
          mma16816(dp[2 * np], ad, bv[0], bv[1]);
          if ((2 * np + 1) * 8 < S) {   // skip all-padding key tiles
            mma16816(c[2 * np + 1], aq, bk[2], bk[3]);
            mma16816(dp[2 * np + 1], ad, bv[2], bv[3]);
          }
        }
      }
      const float l0 = Ls[r0], l1 = Ls[r1];
      float dsum[2] = {0.f, 0.f};
      uint32_t kmask = 0u;   // dropout keep bits of this thread's 16 scores: the counter hash runs once per score
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = (e >> 1) ? r1 : r0, col = nt * 8 + 2 * q + (e & 1);
          const float pr = col < S ? __expf(c[nt][e] * scale - ((e >> 1) ? l1 : l0)) : 0.f;
          float dpv = dp[nt][e];
          if (drop.threshold) {
            const uint64_t idx = ((uint64_t)bh * S + row) * S + col;
            const bool keep = drop_keep(dseed, drop.site, idx, drop.threshold);
            kmask |= (uint32_t)keep << (nt * 4 + e);
            dpv = keep ? dpv * drop.scale : 0.f;   // dP = dP~ (.) mask
          }
          dsum[e >> 1] += pr * dpv;                                // D_i = sum_j P_ij dP_ij  (= dO_i . O_i)
          c[nt][e] = pr;
          dp[nt][e] = dpv;
        }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 1);
        dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 2);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float ds[4], pt[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ds[e] = c[nt][e] * (dp[nt][e] - dsum[e >> 1]) * scale;   // dS
          pt[e] = c[nt][e];
          if (drop.threshold) pt[e] = ((kmask >> (nt * 4 + e)) & 1u) ? pt[e] * drop.scale : 0.f;   // P~
        }
        const int col = nt * 8 + 2 * q;
        if (r0 < S) {
          *reinterpret_cast<uint32_t*>(Ps + r0 * LDP + col) = pack_bf16x2(pt[0], pt[1]);
          *reinterpret_cast<uint32_t*>(dSs + r0 * LDP + col) = pack_bf16x2(ds[0], ds[1]);
        }
        if (r1 < S) {
          *reinterpret_cast<uint32_t*>(Ps + r1 * LDP + col) = pack_bf16x2(pt[2], pt[3]);
          *reinterpret_cast<uint32_t*>(dSs + r1 * LDP + col) = pack_bf16x2(ds[2], ds[3]);
        }
      }
    }
    pair_bar(pair);                                    // both query tiles are parked
    // ---------------- phase 2: dV = P~^T dO, dK = dS^T Q (rows = keys), dQ = dS K (rows = queries) ----------------
    uint32_t kpk[ND][2], vpk[ND][2], qpk[ND][2];
    if (live) {
      {
        float dv[ND][4], dk[ND][4];
#pragma unroll
        for (int nd = 0; nd < ND; ++nd)
#pragma unroll
          for (int e = 0; e < 4; ++e) { dv[nd][e] = 0.f; dk[nd][e] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          if (kk * 16 >= S) break;   // all-padding query block
          uint32_t ap[4], as[4];
          lda_t_c<LDP>(ap, Ps, Z, S, mt, kk, lane);    // P~^T  [keys x queries]
          lda_t_c<LDP>(as, dSs, Z, S, mt, kk, lane);   // dS^T
#pragma unroll
          for (int np = 0; np < ND / 2; ++np) {
            uint32_t bo[4], bq[4];
            ldbt_c<LD>(bo, dOs, Z, S, np, kk, lane);
            ldbt_c<LD>(bq, Qs, Z, S, np, kk, lane);
            mma16816(dv[2 * np], ap, bo[0], bo[1]);
            mma16816(dv[2 * np + 1], ap, bo[2], bo[3]);
            mma16816(dk[2 * np], as, bq[0], bq[1]);
            mma16816(dk[2 * np + 1], as, bq[2], bq[3]);
          }
        }
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
          kpk[nd][0] = pack_bf16x2(dk[nd][0], dk[nd][1]); kpk[nd][1] = pack_bf16x2(dk[nd][2], dk[nd][3]);
          vpk[nd][0] = pack_bf16x2(dv[nd][0], dv[nd][1]); vpk[nd][1] = pack_bf16x2(dv[nd][2], dv[nd][3]);
        }
      }
      float dq[ND][4];
#pragma unroll
      for (int nd = 0; nd < ND; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[nd][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        if (kk * 16 >= S) break;     // all-padding key block
        uint32_t a[4];
        lda_c<LDP>(a, dSs, Z, S, mt, kk, lane);        // dS rows = this warp's queries, k = keys
#pragma unroll
        for (int np = 0; np < ND / 2; ++np) {
          uint32_t bb[4];
          ldbt_c<LD>(bb, Ks, Z, S, np, kk, lane);
          mma16816(dq[2 * np], a, bb[0], bb[1]);
          mma16816(dq[2 * np + 1], a, bb[2], bb[3]);
        }
      }
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        qpk[nd][0] = pack_bf16x2(dq[nd][0], dq[nd][1]);
        qpk[nd][1] = pack_bf16x2(dq[nd][2], dq[nd][3]);
      }
    }
    pair_bar(pair);                                    // every operand read of this problem is done
    if (live) {                                        // results are staged where Q, K, V were
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        const int col = nd * 8 + 2 * q;
        if (r0 < S) {
          *reinterpret_cast<uint32_t*>(Qs + r0 * LD + col) = qpk[nd][0];
          *reinterpret_cast<uint32_t*>(Ks + r0 * LD + col) = kpk[nd][0];
          *reinterpret_cast<uint32_t*>(Vs + r0 * LD + col) = vpk[nd][0];
        }
        if (r1 < S) {
          *reinterpret_cast<uint32_t*>(Qs + r1 * LD + col) = qpk[nd][1];
          *reinterpret_cast<uint32_t*>(Ks + r1 * LD + col) = kpk[nd][1];
          *reinterpret_cast<uint32_t*>(Vs + r1 * LD + col) = vpk[nd][1];
        }
      }
    }
    pair_bar(pair);
    bf16* dQg = dqkv + (size_t)b * S * rs + h * HD;
    copy_out<HD>(dQg, rs, Qs, S, t64);
    copy_out<HD>(dQg + E, rs, Ks, S, t64);
    copy_out<HD>(dQg + 2 * E, rs, Vs, S, t64);
    pair_bar(pair);                                    // staged rows read: the buffer may be refilled
  }
}

static bool use_v1() {
  static int v1 = -1;
  if (v1 < 0) { const char* e = getenv("FERVIT_ATTN_V1"); v1 = (e && atoi(e)) ? 1 : 0; }
  return v1 == 1;
}
// persistent grid: as many CTAs as fit at once, trimmed so that every warp pair gets the same number of problems
static int persistent_grid(int BH, size_t smem, int max_ctas_per_sm) {
  int per_sm = (int)(232448 / (smem + 1024));
  if (per_sm > max_ctas_per_sm) per_sm = max_ctas_per_sm;
  if (per_sm < 1) per_sm = 1;
  const int resident = num_sms() * per_sm;
  const int want = ceil_div(BH, HEADS);
  if (want <= resident) return want;
  const int rounds = ceil_div(want, resident);
  return ceil_div(want, rounds);
}

template <int HD>
int launch_fwd(const bf16* qkv, bf16* out, float* lse, int B, int S, int H, Dropout drop, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)HD);
  ProfScope prof(1, (double)B * S * H * HD * 4.0 * sizeof(bf16) + (double)B * H * S * 4.0, stream);
  if (!use_v1()) {
    constexpr int LD = Lay<HD>::LD;
    const int SA = (S + 3) & ~3;
    const size_t smem = (size_t)(LD + HEADS * 6 * SA * LD) * 2;
    static size_t attr = 0;
    if (smem > attr) {
      FV_CUDA(cudaFuncSetAttribute(attn_tc_fwd2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    FV_CUDA(launch_pdl(attn_tc_fwd2_kernel<HD>, dim3(persistent_grid(B * H, smem, 3)), dim3(WARPS * 32), smem, stream,
                       qkv, out, lse, B, S, H, scale, drop, SA));
  } else {
    constexpr int smem = HEADS * 3 * SP * Lay<HD>::LD * 2;
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
      FV_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr = true;
    }
    FV_CUDA(launch_pdl(attn_tc_fwd_kernel<HD>, dim3(ceil_div(B * H, HEADS)), dim3(WARPS * 32), (size_t)smem, stream, qkv,
                       out, lse, B, S, H, scale, drop));
  }
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template <int HD>
int launch_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int B, int S, int H,
               Dropout drop, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)HD);
  // version 2 does not read O (D = sum_j P dP), so its algorithmic bytes are 7 rows of HD per token, not 8
  ProfScope prof(1, (double)B * S * H * HD * (use_v1() ? 8.0 : 7.0) * sizeof(bf16) + (double)B * H * S * 4.0, stream);
  if (!use_v1()) {
    constexpr int LD = Lay<HD>::LD;
    const int SA = (S + 3) & ~3;
    const size_t smem = (size_t)(LD + HEADS * (2 * (4 * SA * LD + 2 * SP) + 2 * SA * (SP + 8))) * 2;
    static size_t attr = 0;
    if (smem > attr) {
      FV_CUDA(cudaFuncSetAttribute(attn_tc_bwd2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    FV_CUDA(launch_pdl(attn_tc_bwd2_kernel<HD>, dim3(persistent_grid(B * H, smem, 2)), dim3(WARPS * 32), smem, stream,
                       qkv, dout, lse, dqkv, B, S, H, scale, drop, SA));
  } else {
    constexpr int smem = HEADS * (4 * SP * Lay<HD>::LD * 2 + 2 * SP * 4 + 2 * SP * (SP + 8) * 2);
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
      FV_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr = true;
    }
    FV_CUDA(launch_pdl(attn_tc_bwd_kernel<HD>, dim3(ceil_div(B * H, HEADS)), dim3(WARPS * 32), (size_t)smem, stream, qkv,
                       out, dout, lse, dqkv, B, S, H, scale, drop));
  }
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
