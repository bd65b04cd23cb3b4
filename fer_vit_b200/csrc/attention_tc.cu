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

template <int HD>
int launch_fwd(const bf16* qkv, bf16* out, float* lse, int B, int S, int H, Dropout drop, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)HD);
  constexpr int smem = HEADS * 3 * SP * Lay<HD>::LD * 2;
  static bool attr = false;
  if (!attr && smem > 48 * 1024) {
    FV_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  ProfScope prof(1, (double)B * S * H * HD * 4.0 * sizeof(bf16) + (double)B * H * S * 4.0, stream);
  FV_CUDA(launch_pdl(attn_tc_fwd_kernel<HD>, dim3(ceil_div(B * H, HEADS)), dim3(WARPS * 32), (size_t)smem, stream, qkv, out,
                     lse, B, S, H, scale, drop));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template <int HD>
int launch_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int B, int S, int H,
               Dropout drop, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)HD);
  constexpr int smem = HEADS * (4 * SP * Lay<HD>::LD * 2 + 2 * SP * 4 + 2 * SP * (SP + 8) * 2);
  static bool attr = false;
  if (!attr && smem > 48 * 1024) {
    FV_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  ProfScope prof(1, (double)B * S * H * HD * 8.0 * sizeof(bf16) + (double)B * H * S * 4.0, stream);
  FV_CUDA(launch_pdl(attn_tc_bwd_kernel<HD>, dim3(ceil_div(B * H, HEADS)), dim3(WARPS * 32), (size_t)smem, stream, qkv, out,
                     dout, lse, dqkv, B, S, H, scale, drop));
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
