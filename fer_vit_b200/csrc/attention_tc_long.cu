// Tensor-core attention for 32 < S <= 256 (ImageViT: 196 patches + cls = 197 tokens; ExpressionAwareViT concat mode:
// 37 tokens), bf16, head dim 32 / 48 / 64. One CTA per (sample, head): Q, K, V (and dO) of the head are staged ONCE in
// shared memory with cp.async, each warp owns 16-row tiles, every product is a chain of mma.sync m16n8k16 with
// ldmatrix(.trans) operands (nothing is stored transposed), softmax runs on the accumulator fragments.
//   forward : per 16-query tile two passes over the keys in blocks of 16 (row maximum, then exp / row sum / O += p V with
//             the accumulator fragments of p as A operands; O is normalised once at the end): no online rescaling.
//   backward: phase 1, per 16-query tile, walks the keys in blocks of 16: S, dP -> P~, dS (the row statistics come from
//             the forward pass' LSE and D = rowsum(dO * O)), dQ += dS K. Phase 2, per 16-key tile, walks the queries
//             in blocks of 16 and recomputes S^T, dP^T so that dV = P~^T dO and dK = dS^T Q accumulate in the
//             registers of ONE warp: no cross-warp reduction, no atomics (deterministic).
// Replaces F.scaled_dot_product_attention inside nn.TransformerEncoderLayer (image_vit.py:101-113, dropout on the
// attention weights in training) for the sequence lengths the warp-per-head kernel (attention_tc.cu) does not cover;
// before it, these ran on the CUDA-core kernel of attention.cu (ImageViT config 2: 30.5 ms per step at batch 64).
#include "common.cuh"
#include "kernels.h"
#include "attn_frag.cuh"

namespace fervit {
namespace attn_long {

using namespace attn_frag;


__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
// stage M[S][HD] (global, row stride rs) into smem [SPAD][LD], zero-filling rows S..SPAD-1
template <int HD>
__device__ __forceinline__ void stage(bf16* dst, const bf16* __restrict__ src, size_t rs, int S, int SPAD, int tid,
                                      int nthreads) {
  constexpr int CH = HD / 8, LD = Lay<HD>::LD;
  for (int i = tid; i < SPAD * CH; i += nthreads) {
    const int r = i / CH, c = (i % CH) * 8;
    if (r < S) cp_async16(dst + r * LD + c, src + (size_t)r * rs + c);
    else *reinterpret_cast<uint4*>(dst + r * LD + c) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ------------------------------------------------------------------------------------------------ forward
// One warp per 16-query tile (MAXW = the most warps a CTA may have: ceil(S / 16) when that fits, else the tiles are
// dealt round-robin). Two passes over the keys in blocks of 16 keep the register footprint small enough for two CTAs
// per SM: pass 1 only finds the row maximum of the scaled scores, pass 2 recomputes the score block, forms
// p = exp(s - max), accumulates the row sum and O += p V unnormalised; 1 / sum (and the dropout scale) is applied to
// the 16 x HD output once. The extra Q K^T pass is cheaper than holding the whole score row (104 fp32 registers at
// S = 197), which capped the kernel at 8 warps per SM.
template <int HD, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, MAXW <= 13 ? 2 : 1)
attn_long_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int S, int H,
                     int SPAD, float scale, Dropout drop) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int KS = HD / 16, ND = HD / 8, LD = Lay<HD>::LD;
  pdl_trigger();
  pdl_grid_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int bh = blockIdx.x;
  const int b = bh / H, h = bh % H;
  const int E = H * HD;
  const size_t rs = (size_t)3 * E;
  const bf16* Qg = qkv + (size_t)b * S * rs + h * HD;
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + SPAD * LD;
  bf16* Vs = Ks + SPAD * LD;
  stage<HD>(Qs, Qg, rs, S, SPAD, threadIdx.x, blockDim.x);
  stage<HD>(Ks, Qg + E, rs, S, SPAD, threadIdx.x, blockDim.x);
  stage<HD>(Vs, Qg + 2 * E, rs, S, SPAD, threadIdx.x, blockDim.x);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const uint64_t dseed = drop.threshold ? drop.eff() : 0;
  bf16* Og = out + (size_t)b * S * E + h * HD;
  const int nblk = (S + 15) >> 4;
#pragma unroll 1
  for (int mt = warp; mt < nblk; mt += nwarps) {
    uint32_t aq[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) lda<LD>(aq[ks], Qs, mt, ks, lane);
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    // ---- pass 1: row maxima (rows r0: elements 0,1; r1: elements 2,3) ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll 2
    for (int kb = 0; kb < nblk; ++kb) {
      float c[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bk[4];
        ldb<LD>(bk, Ks, kb, ks, lane);
        mma16816(c[0], aq[ks], bk[0], bk[1]);
        mma16816(c[1], aq[ks], bk[2], bk[3]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = kb * 16 + nt * 8 + 2 * q + (e & 1);
          if (col < S) mx[e >> 1] = fmaxf(mx[e >> 1], c[nt][e] * scale);
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    // ---- pass 2: p = exp(s - max), row sums, O += p V ----
    float sum[2] = {0.f, 0.f};
    float o[ND][4];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[nd][e] = 0.f;
#pragma unroll 1
    for (int kb = 0; kb < nblk; ++kb) {
      float c[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bk[4];
        ldb<LD>(bk, Ks, kb, ks, lane);
        mma16816(c[0], aq[ks], bk[0], bk[1]);
        mma16816(c[1], aq[ks], bk[2], bk[3]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = kb * 16 + nt * 8 + 2 * q + (e & 1);
          float p = col < S ? __expf(c[nt][e] * scale - mx[e >> 1]) : 0.f;
          sum[e >> 1] += p;
          if (drop.threshold) {
            const uint64_t idx = ((uint64_t)bh * S + ((e >> 1) ? r1 : r0)) * S + col;
            if (!drop_keep(dseed, drop.site, idx, drop.threshold)) p = 0.f;
          }
          c[nt][e] = p;
        }
      // the accumulator fragments of P are exactly the A fragments of the next MMA
      uint32_t a[4];
      a[0] = pack_bf16x2(c[0][0], c[0][1]);
      a[1] = pack_bf16x2(c[0][2], c[0][3]);
      a[2] = pack_bf16x2(c[1][0], c[1][1]);
      a[3] = pack_bf16x2(c[1][2], c[1][3]);
#pragma unroll
      for (int np = 0; np < ND / 2; ++np) {
        uint32_t bb[4];
        ldbt<LD>(bb, Vs, np, kb, lane);
        mma16816(o[2 * np], a, bb[0], bb[1]);
        mma16816(o[2 * np + 1], a, bb[2], bb[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
    }
    if (lse && q == 0) {
      if (r0 < S) lse[(size_t)bh * S + r0] = mx[0] + __logf(sum[0]);
      if (r1 < S) lse[(size_t)bh * S + r1] = mx[1] + __logf(sum[1]);
    }
    const float ds = drop.threshold ? drop.scale : 1.0f;
    const float inv[2] = {ds / sum[0], ds / sum[1]};
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      const int col = nd * 8 + 2 * q;
      if (r0 < S) *reinterpret_cast<uint32_t*>(Og + (size_t)r0 * E + col) = pack_bf16x2(o[nd][0] * inv[0], o[nd][1] * inv[0]);
      if (r1 < S) *reinterpret_cast<uint32_t*>(Og + (size_t)r1 * E + col) = pack_bf16x2(o[nd][2] * inv[1], o[nd][3] * inv[1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// MAXW = most warps of a CTA: one warp per 16-row tile when ceil(S / 16) <= 13 (both phases in ONE round instead of two
// uneven ones with 8 warps), else 8 warps taking the tiles round-robin
// (13 warps put 4 on one SM sub-partition: 16384 / 4 / 32 = 128 registers per thread, a few spilled words at HD = 64)
template <int HD, int MAXW>
__global__ void __launch_bounds__(MAXW * 32)
attn_long_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ dout,
                     const float* __restrict__ lse, bf16* __restrict__ dqkv, int S, int H, int SPAD, float scale,
                     Dropout drop) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int KS = HD / 16, ND = HD / 8, LD = Lay<HD>::LD;
  pdl_trigger();
  pdl_grid_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int bh = blockIdx.x;
  const int b = bh / H, h = bh % H;
  const int E = H * HD;
  const size_t rs = (size_t)3 * E;
  const int MAT = SPAD * LD;
  const bf16* Qg = qkv + (size_t)b * S * rs + h * HD;
  const bf16* Og = out + (size_t)b * S * E + h * HD;
  const bf16* dOg = dout + (size_t)b * S * E + h * HD;
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + MAT;
  bf16* Vs = Ks + MAT;
  bf16* dOs = Vs + MAT;
  float* Ls = reinterpret_cast<float*>(dOs + MAT);
  float* Ds = Ls + SPAD;
  // dropout keep bits of the head, written by phase 1 and read (transposed) by phase 2 so that the counter hash runs
  // once per score: byte [(query * nblk + key block) * 4 + q] holds the 4 keys lane q of a quad owns in that block
  uint8_t* keepb = reinterpret_cast<uint8_t*>(Ds + SPAD);
  const int NTH = blockDim.x, nwarps = blockDim.x >> 5;
  stage<HD>(Qs, Qg, rs, S, SPAD, threadIdx.x, NTH);
  stage<HD>(Ks, Qg + E, rs, S, SPAD, threadIdx.x, NTH);
  stage<HD>(Vs, Qg + 2 * E, rs, S, SPAD, threadIdx.x, NTH);
  stage<HD>(dOs, dOg, (size_t)E, S, SPAD, threadIdx.x, NTH);
  // D_i = dO_i . O_i ; rows beyond S get LSE = +inf so their probabilities vanish
  for (int i = threadIdx.x; i < SPAD; i += NTH) {
    float dsum = 0.f;
    if (i < S) {
#pragma unroll
      for (int d0 = 0; d0 < HD; d0 += 8) {
        const uint4 ov = __ldg(reinterpret_cast<const uint4*>(Og + (size_t)i * E + d0));
        const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dOg + (size_t)i * E + d0));
        const uint32_t* op = reinterpret_cast<const uint32_t*>(&ov);
        const uint32_t* dp = reinterpret_cast<const uint32_t*>(&dv);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 a = unpack_bf16x2(op[t]), bb = unpack_bf16x2(dp[t]);
          dsum += a.x * bb.x + a.y * bb.y;
        }
      }
    }
    Ds[i] = dsum;
    Ls[i] = i < S ? lse[(size_t)bh * S + i] : INFINITY;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const uint64_t dseed = drop.threshold ? drop.eff() : 0;
  bf16* dQg = dqkv + (size_t)b * S * rs + h * HD;
  bf16* dKg = dQg + E;
  bf16* dVg = dQg + 2 * E;
  const int nblk = SPAD / 16;

  // ---------------- phase 1: dQ (rows are queries; the keys are walked in blocks of 16) ----------------
#pragma unroll 1
  for (int mt = warp; mt * 16 < S; mt += nwarps) {
    uint32_t aq[KS][4], ad[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      lda<LD>(aq[ks], Qs, mt, ks, lane);
      lda<LD>(ad[ks], dOs, mt, ks, lane);
    }
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const float l0 = Ls[r0], l1 = Ls[r1], d0v = Ds[r0], d1v = Ds[r1];
    float dq[ND][4];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[nd][e] = 0.f;
#pragma unroll 1
    for (int kb = 0; kb < nblk; ++kb) {
      if (kb * 16 >= S) break;
      float c[2][4], dp[2][4];
      uint32_t nib[2] = {0u, 0u};
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { c[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bk[4], bv[4];
        ldb<LD>(bk, Ks, kb, ks, lane);
        ldb<LD>(bv, Vs, kb, ks, lane);
        mma16816(c[0], aq[ks], bk[0], bk[1]);
        mma16816(c[1], aq[ks], bk[2], bk[3]);
        mma16816(dp[0], ad[ks], bv[0], bv[1]);
        mma16816(dp[1], ad[ks], bv[2], bv[3]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = (e >> 1) ? r1 : r0, col = kb * 16 + nt * 8 + 2 * q + (e & 1);
          const float p = col < S ? __expf(c[nt][e] * scale - ((e >> 1) ? l1 : l0)) : 0.f;
          float dpv = dp[nt][e];
          if (drop.threshold) {
            const uint64_t idx = ((uint64_t)bh * S + row) * S + col;
            const bool keep = drop_keep(dseed, drop.site, idx, drop.threshold);
            nib[e >> 1] |= (uint32_t)keep << (nt * 2 + (e & 1));
            dpv = keep ? dpv * drop.scale : 0.f;
          }
          c[nt][e] = p * (dpv - ((e >> 1) ? d1v : d0v)) * scale;   // dS
        }
      if (drop.threshold) {
        keepb[(r0 * nblk + kb) * 4 + q] = (uint8_t)nib[0];
        keepb[(r1 * nblk + kb) * 4 + q] = (uint8_t)nib[1];
      }
      uint32_t a[4];
      a[0] = pack_bf16x2(c[0][0], c[0][1]);
      a[1] = pack_bf16x2(c[0][2], c[0][3]);
      a[2] = pack_bf16x2(c[1][0], c[1][1]);
      a[3] = pack_bf16x2(c[1][2], c[1][3]);
#pragma unroll
      for (int np = 0; np < ND / 2; ++np) {
        uint32_t bb[4];
        ldbt<LD>(bb, Ks, np, kb, lane);
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

  if (drop.threshold) __syncthreads();   // phase 2 reads the keep bits every warp wrote in phase 1

  // ---------------- phase 2: dK, dV (rows are keys j; the queries i are walked in blocks of 16) ----------------
#pragma unroll 1
  for (int mt = warp; mt * 16 < S; mt += nwarps) {
    uint32_t ak[KS][4], av[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      lda<LD>(ak[ks], Ks, mt, ks, lane);
      lda<LD>(av[ks], Vs, mt, ks, lane);
    }
    const int j0 = mt * 16 + g, j1 = j0 + 8;
    float dv[ND][4], dk[ND][4];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) { dv[nd][e] = 0.f; dk[nd][e] = 0.f; }
#pragma unroll 1
    for (int qb = 0; qb < nblk; ++qb) {
      if (qb * 16 >= S) break;
      float c[2][4], dp[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { c[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bq[4], bd[4];
        ldb<LD>(bq, Qs, qb, ks, lane);
        ldb<LD>(bd, dOs, qb, ks, lane);
        mma16816(c[0], ak[ks], bq[0], bq[1]);     // S^T block: rows keys, columns queries
        mma16816(c[1], ak[ks], bq[2], bq[3]);
        mma16816(dp[0], av[ks], bd[0], bd[1]);    // dP^T block
        mma16816(dp[1], av[ks], bd[2], bd[3]);
      }
      float pt[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = (e >> 1) ? j1 : j0, i = qb * 16 + nt * 8 + 2 * q + (e & 1);
          const float p = (j < S) ? __expf(c[nt][e] * scale - Ls[i]) : 0.f;   // Ls[i >= S] = +inf -> 0
          float ptv = p, dpv = dp[nt][e];
          if (drop.threshold) {
            // key j sits in block mt at column g (+8): lane g/2 of the quad, bit (g & 1) (+2) of its byte
            const uint32_t byte = keepb[(i * nblk + mt) * 4 + (g >> 1)];
            const bool keep = (byte >> ((g & 1) + ((e >> 1) << 1))) & 1u;
            ptv = keep ? p * drop.scale : 0.f;
            dpv = keep ? dpv * drop.scale : 0.f;
          }
          pt[nt][e] = ptv;                             // P~^T
          c[nt][e] = p * (dpv - Ds[i]) * scale;        // dS^T
        }
      uint32_t ap[4], as[4];
      ap[0] = pack_bf16x2(pt[0][0], pt[0][1]);
      ap[1] = pack_bf16x2(pt[0][2], pt[0][3]);
      ap[2] = pack_bf16x2(pt[1][0], pt[1][1]);
      ap[3] = pack_bf16x2(pt[1][2], pt[1][3]);
      as[0] = pack_bf16x2(c[0][0], c[0][1]);
      as[1] = pack_bf16x2(c[0][2], c[0][3]);
      as[2] = pack_bf16x2(c[1][0], c[1][1]);
      as[3] = pack_bf16x2(c[1][2], c[1][3]);
#pragma unroll
      for (int np = 0; np < ND / 2; ++np) {
        uint32_t bo[4], bq2[4];
        ldbt<LD>(bo, dOs, np, qb, lane);
        ldbt<LD>(bq2, Qs, np, qb, lane);
        mma16816(dv[2 * np], ap, bo[0], bo[1]);
        mma16816(dv[2 * np + 1], ap, bo[2], bo[3]);
        mma16816(dk[2 * np], as, bq2[0], bq2[1]);
        mma16816(dk[2 * np + 1], as, bq2[2], bq2[3]);
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

// ------------------------------------------------------------------------------------------------ host
template <int HD, int MAXW>
int launch_fwd(const bf16* qkv, bf16* out, float* lse, int B, int S, int H, Dropout drop, cudaStream_t stream) {
  const int SPAD = (S + 15) / 16 * 16;
  const int smem = 3 * SPAD * Lay<HD>::LD * 2;
  static int smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    FV_CUDA(cudaFuncSetAttribute(attn_long_fwd_kernel<HD, MAXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  const int tiles = SPAD / 16;
  const int warps = tiles < MAXW ? tiles : MAXW;
  const float scale = 1.0f / sqrtf((float)HD);
  ProfScope prof(1, (double)B * S * H * HD * 4.0 * sizeof(bf16) + (double)B * H * S * 4.0, stream);
  FV_CUDA(launch_pdl(attn_long_fwd_kernel<HD, MAXW>, dim3(B * H), dim3(warps * 32), (size_t)smem, stream, qkv, out, lse,
                     S, H, SPAD, scale, drop));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template <int HD>
int dispatch_fwd(const bf16* qkv, bf16* out, float* lse, int B, int S, int H, Dropout drop, cudaStream_t stream) {
  if (S <= 208) return launch_fwd<HD, 13>(qkv, out, lse, B, S, H, drop, stream);
  return launch_fwd<HD, 16>(qkv, out, lse, B, S, H, drop, stream);
}
template <int HD, int MAXW>
int launch_bwd_w(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int B, int S, int H,
                 Dropout drop, cudaStream_t stream) {
  const int SPAD = (S + 15) / 16 * 16;
  const int smem = 4 * SPAD * Lay<HD>::LD * 2 + 2 * SPAD * 4 + (drop.threshold ? SPAD * (SPAD / 16) * 4 : 0);
  static int smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    FV_CUDA(cudaFuncSetAttribute(attn_long_bwd_kernel<HD, MAXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  const int tiles = SPAD / 16;
  const int warps = tiles < MAXW ? tiles : MAXW;
  const float scale = 1.0f / sqrtf((float)HD);
  ProfScope prof(1, (double)B * S * H * HD * 8.0 * sizeof(bf16) + (double)B * H * S * 4.0, stream);
  FV_CUDA(launch_pdl(attn_long_bwd_kernel<HD, MAXW>, dim3(B * H), dim3(warps * 32), (size_t)smem, stream, qkv, out, dout,
                     lse, dqkv, S, H, SPAD, scale, drop));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}
template <int HD>
int launch_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int B, int S, int H,
               Dropout drop, cudaStream_t stream) {
  if (S <= 208) return launch_bwd_w<HD, 13>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
  return launch_bwd_w<HD, 8>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
}

}  // namespace attn_long

bool attention_tc_long_supported(int S, int HD) {
  static int off = -1;
  if (off < 0) { const char* s = getenv("FERVIT_ATTN_LONG"); off = (s && atoi(s) == 0) ? 1 : 0; }
  return off == 0 && S > 32 && S <= 256 && (HD == 64 || HD == 48 || HD == 32);
}

int attention_tc_long_fwd(const bf16* qkv, bf16* out, float* lse, int B, int S, int H, int HD, Dropout drop,
                          cudaStream_t stream) {
  if (HD == 64) return attn_long::dispatch_fwd<64>(qkv, out, lse, B, S, H, drop, stream);
  if (HD == 48) return attn_long::dispatch_fwd<48>(qkv, out, lse, B, S, H, drop, stream);
  if (HD == 32) return attn_long::dispatch_fwd<32>(qkv, out, lse, B, S, H, drop, stream);
  FV_CHECK(false, "attention_tc_long: head dim %d not supported", HD);
}
int attention_tc_long_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int B,
                          int S, int H, int HD, Dropout drop, cudaStream_t stream) {
  if (HD == 64) return attn_long::launch_bwd<64>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
  if (HD == 48) return attn_long::launch_bwd<48>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
  if (HD == 32) return attn_long::launch_bwd<32>(qkv, out, dout, lse, dqkv, B, S, H, drop, stream);
  FV_CHECK(false, "attention_tc_long: head dim %d not supported", HD);
}

}  // namespace fervit
