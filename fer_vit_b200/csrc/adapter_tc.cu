// Fused AdapterModule on the tensor cores (hybrid_latent_vit.py:249-265), forward and input-gradient in ONE kernel each:
//
//   forward : y  = x  + alpha * (GELU(x W1^T + b1) W2^T + b2)         saves g = GELU(u) and d = GELU'(u)   [T, 64] bf16
//   backward: dx = dy + (alpha * (dy W2) * d) W1                       saves du = alpha * (dy W2) * d       [T, 64] bf16
//
// Both are "out = res + f(in · A^T) · B^T" with a 64-wide bottleneck, so one kernel serves both:
//   phase 1  H[128, 64]  = IN[128, E] · A[64, E]^T        TMA ring -> tcgen05.mma (M = 128, N = 64) -> TMEM
//   phase 2  G = f(H) in registers (thread = row), written as the bf16 A OPERAND of the next MMA straight into shared
//            memory in the canonical SWIZZLE_128B K-major layout (what TMA would have produced); g / d / du to global
//   phase 3  Y[128, 256] = G[128, 64] · B[256, 64]^T      one k-block, N = 256, accumulator in TMEM
//   phase 4  out = res + alpha * (Y + b2): fp32 residual tiles TMA-loaded, updated in place, TMA-stored (+ bf16 copy)
// The two skinny GEMMs it replaces each paid a full launch, prologue and epilogue round trip for ~0.5 GFLOP
// (12.7 + 11.8 us forward, 10.1 + 11.8 us backward per block at batch 256) and moved the [T, 64] intermediate
// through HBM twice. Grid: one CTA per (128-row block, 256-column group); phase 1 is recomputed per column group (the
// input tile comes from L2).
#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace fervit {
namespace adp {

using namespace ptx;

constexpr int BM = 128, BK = 64, AD = 64, NC = 256, UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = EPI_WARPS * 32 + 128;
constexpr int W_TMA = EPI_WARPS, W_MMA = EPI_WARPS + 1, W_ALLOC = EPI_WARPS + 2;
constexpr int STAGES = 7;   // phase 1 is latency-bound: bytes in flight are what counts (3 stages: 6.4 us; 7: ~3 us)
constexpr int IN_BYTES = BM * BK * 2;    // 16 KB
constexpr int AW_BYTES = AD * BK * 2;    // 8 KB
constexpr int STAGE_BYTES = IN_BYTES + AW_BYTES;
constexpr int G_BYTES = BM * AD * 2;     // 16 KB: the bf16 A operand of phase 3
constexpr int BW_BYTES = NC * AD * 2;    // 32 KB
constexpr int XT = 32 * 32 * 4, YT = 32 * 32 * 2;
constexpr int WARP_STAGING = 2 * XT + 2 * YT;
constexpr int BAR_BYTES = 512;
// The epilogue staging tiles ALIAS the input half of the phase-1 ring: every ring stage is free once the phase-1
// accumulator is complete (h_full), which is when the residual tiles start to travel.
static_assert(EPI_WARPS * WARP_STAGING <= STAGES * IN_BYTES, "staging must fit in the aliased ring region");
// b1 [64] and this column group's b2 [256] staged once by the (otherwise idle) epilogue warps during phase 1: phase 2
// and every phase-4 chunk paid one global-load latency for them (2.8 us of phase 2, profiles/r02_adapter_timeline.jsonl)
constexpr int BIAS_BYTES = (AD + NC) * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + G_BYTES + BW_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
constexpr int TMEM_COLS = 512;           // H at columns [0, 64), Y at [256, 512)
static_assert(SMEM_BYTES <= 232448, "adapter kernel: shared memory");

// FERVIT_GEMM_DEBUG bit 64: clock64 stamps of CTA 0 at its phase boundaries (tools/gemm_timeline.py --adapter)
__device__ unsigned long long g_adapter_timeline[32];

struct Params {
  int debug;
  int T, E, groups;        // rows, width, column groups (E / 256)
  int backward;            // 0: forward (f = GELU + b1), 1: backward (f = alpha * h * d)
  const float* b1;         // [64]   forward
  const float* b2;         // [E]    forward
  const float* alpha_ptr;  // device scalar
  const bf16* d_in;        // [T, 64] backward: GELU'(u) saved by the forward pass
  bf16* s0;                // [T, 64] forward: g;  backward: du
  bf16* s1;                // [T, 64] forward: d;  backward: unused
  int has_out_bf16;
  // folded LayerNorm, producer side (common.cuh: Epilogue): the rows of `out` are the input of the next block's norm1;
  // the bf16 copy becomes bf16(y - mref[row]) and each warp writes the {sum, sum of squares} of its 128-column part
  float* lnp_part;
  const float* lnp_mref;
};

__global__ void __launch_bounds__(THREADS, 1)
adapter_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_aw,
               const __grid_constant__ CUtensorMap tm_bw, const __grid_constant__ CUtensorMap tm_res,
               const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_outb,
               const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_in = smem;
  uint8_t* smem_aw = smem + STAGES * IN_BYTES;
  uint8_t* smem_g = smem + STAGES * STAGE_BYTES;
  uint8_t* smem_bw = smem_g + G_BYTES;
  uint8_t* staging = smem_in;   // aliased: valid after h_full
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bw + BW_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* h_full = bars + 2 * STAGES;      // phase 1 accumulator complete
  uint64_t* g_ready = h_full + 1;            // G operand written by all epilogue warps
  uint64_t* y_full = h_full + 2;             // phase 3 accumulator complete
  uint64_t* b_full = h_full + 3;             // B tile landed
  uint64_t* ld_bar = h_full + 4;             // [EPI_WARPS][2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(ld_bar + 2 * EPI_WARPS);
  float* sb1 = reinterpret_cast<float*>(smem_bw + BW_BYTES + BAR_BYTES);   // [64]
  float* sb2 = sb1 + AD;                                                   // [256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_blk = blockIdx.x / p.groups, cg = blockIdx.x % p.groups;
  const int total_kb = (p.E + BK - 1) / BK;

  const bool tl_on = (p.debug & 64) && blockIdx.x == 0;
  auto TL = [&](int slot) {
    if (tl_on) g_adapter_timeline[slot] = (unsigned long long)clock64();
  };
  pdl_trigger();
  if (tl_on && threadIdx.x == 0) {
    TL(0);
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_adapter_timeline[30] = t;
  }
  if (warp == W_TMA && lane == 0) {
    prefetch_tmap(&tm_in); prefetch_tmap(&tm_aw); prefetch_tmap(&tm_bw);
    prefetch_tmap(&tm_res); prefetch_tmap(&tm_out);
    if (p.has_out_bf16) prefetch_tmap(&tm_outb);
  }
  if (warp == W_MMA && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(h_full, 1);
    mbar_init(g_ready, EPI_WARPS);
    mbar_init(y_full, 1);
    mbar_init(b_full, 1);
    for (int i = 0; i < 2 * EPI_WARPS; ++i) mbar_init(&ld_bar[i], 1);
    mbar_fence_init();
  }
  if (warp == W_ALLOC) {
    tmem_alloc<1>(tmem_base_slot, (uint32_t)TMEM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (threadIdx.x == 0) TL(1);
  pdl_grid_sync();
  if (threadIdx.x == 0) TL(2);

  if (warp == W_TMA) {
    if (lane == 0) {
      mbar_expect_tx(b_full, BW_BYTES);
      tma_load_2d(smem_bw, &tm_bw, b_full, 0, cg * NC);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 11);
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        tma_load_2d(smem_in + stage * IN_BYTES, &tm_in, &full_bar[stage], kb * BK, row_blk * BM);
        tma_load_2d(smem_aw + stage * AW_BYTES, &tm_aw, &full_bar[stage], kb * BK, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == W_MMA) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(BM, AD, false, false);
      constexpr uint32_t idesc2 = make_idesc(BM, NC, false, false);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase, 12);
        tc_fence_after();
        if (kb == 0) TL(3);
        const uint32_t a_addr = smem_u32(smem_in + stage * IN_BYTES);
        const uint32_t b_addr = smem_u32(smem_aw + stage * AW_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16<1>(tmem_base, make_smem_desc(a_addr + k * (UMMA_K * 2), 16, 1024),
                       make_smem_desc(b_addr + k * (UMMA_K * 2), 16, 1024), idesc1, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(h_full);
      TL(4);
      // phase 3: the epilogue warps have written G (generic proxy -> fenced), the B tile has landed
      mbar_wait(g_ready, 0, 13);
      mbar_wait(b_full, 0, 14);
      tc_fence_after();
      const uint32_t g_addr = smem_u32(smem_g), w_addr = smem_u32(smem_bw);
#pragma unroll
      for (int k = 0; k < AD / UMMA_K; ++k)
        umma_bf16<1>(tmem_base + 256u, make_smem_desc(g_addr + k * (UMMA_K * 2), 16, 1024),
                     make_smem_desc(w_addr + k * (UMMA_K * 2), 16, 1024), idesc2, k > 0 ? 1u : 0u);
      umma_commit(y_full);
      TL(7);
    }
  } else if (warp < EPI_WARPS) {
    const int quarter = warp & 3, half = warp >> 2;
    const int r = lane;
    const int rloc = quarter * 32 + r;              // row inside the 128-row tile
    const int row = row_blk * BM + rloc;
    const int row0 = row_blk * BM + quarter * 32;   // first row of this warp (TMA boxes)
    uint8_t* wst = staging + warp * WARP_STAGING;
    uint8_t* Xs = wst;
    uint8_t* Ys = wst + 2 * XT;
    uint64_t* my_ld = ld_bar + warp * 2;
    const float alpha = __ldg(p.alpha_ptr);
    const int col_base = cg * NC + half * (NC / 2);
    const bool rows_live = row0 < p.T;
    auto issue_res = [&](int j) {
      mbar_expect_tx(&my_ld[j & 1], XT);
      tma_load_2d(Xs + (j & 1) * XT, &tm_res, &my_ld[j & 1], col_base + j * 32, row0);
    };
    uint4 dv[4];
    if (!p.backward) {
      const int t = threadIdx.x;   // 0 .. 255 (the epilogue warps), while the ring feeds phase 1
      if (t < AD) sb1[t] = __ldg(p.b1 + t);
      sb2[t] = __ldg(p.b2 + cg * NC + t);
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    } else {
      // the saved GELU'(u) of this thread's row travels while phase 1 runs
      const uint4* dp = reinterpret_cast<const uint4*>(p.d_in + (size_t)row * AD + half * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) dv[c] = row < p.T ? __ldg(dp + c) : make_uint4(0u, 0u, 0u, 0u);
    }
    // ---------------- phase 2: G = f(H) ----------------
    mbar_wait(h_full, 0, 15);
    tc_fence_after();
    if (warp == 0 && lane == 0) TL(5);
    // the ring is free now (all phase-1 MMAs have completed): the residual tiles travel during phases 2-3
    if (lane == 0 && rows_live) { issue_res(0); issue_res(1); }
    {
      uint32_t hr[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 32), hr);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        asm volatile("" : "+r"(hr[i]));
        v[i] = __uint_as_float(hr[i]);
      }
      uint32_t gq[16], dq[16];
      if (!p.backward) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float g0 = v[i] + sb1[half * 32 + i], g1 = v[i + 1] + sb1[half * 32 + i + 1], d0, d1;
          gelu_fwd_deriv_poly2(g0, g1, d0, d1);      // packed fp32 (FFMA2), as in the GEMM epilogues
          gq[i >> 1] = pack_bf16x2(g0, g1);
          dq[i >> 1] = pack_bf16x2(d0, d1);
        }
      } else {
        const uint32_t* dw = reinterpret_cast<const uint32_t*>(dv);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 d = unpack_bf16x2(dw[i >> 1]);
          gq[i >> 1] = pack_bf16x2(alpha * v[i] * d.x, alpha * v[i + 1] * d.y);
        }
      }
      // the A operand of phase 3: row rloc, 16-byte chunks half*4 .. half*4+3 of its 128-byte K-major row, 128B swizzle
      uint8_t* grow = smem_g + rloc * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(grow + (((half * 4 + c) ^ (rloc & 7)) << 4)) =
            make_uint4(gq[4 * c], gq[4 * c + 1], gq[4 * c + 2], gq[4 * c + 3]);
      fence_proxy_async_smem();   // generic-proxy writes of G -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_ready);
      // The [T, 64] tensors the backward pass / the weight gradients need are written AFTER phase 3 has been released
      // (they sat in front of the arrive: 0.6 us of every row block's critical path) and by two different column
      // groups, so that no single CTA of a row block carries both.
      const int cg_d = p.groups > 1 ? 1 : 0;
      if (cg == 0 && row < p.T) {
        uint4* s0 = reinterpret_cast<uint4*>(p.s0 + (size_t)row * AD + half * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) s0[c] = make_uint4(gq[4 * c], gq[4 * c + 1], gq[4 * c + 2], gq[4 * c + 3]);
      }
      if (!p.backward && cg == cg_d && row < p.T) {
        uint4* s1 = reinterpret_cast<uint4*>(p.s1 + (size_t)row * AD + half * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) s1[c] = make_uint4(dq[4 * c], dq[4 * c + 1], dq[4 * c + 2], dq[4 * c + 3]);
      }
    }
    if (warp == 0 && lane == 0) TL(6);

    // ---------------- phase 4: out = res + alpha * (Y + b2) ----------------
    mbar_wait(y_full, 0, 16);
    tc_fence_after();
    if (warp == 0 && lane == 0) TL(8);
    if (rows_live) {
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + 256u + (uint32_t)(half * (NC / 2));
      const uint32_t xsw = (uint32_t)(r & 7), ysw = (uint32_t)((r >> 1) & 3);
      uint32_t rr[32];
      tmem_ld32(taddr0, rr);
      float ln_s1 = 0.f, ln_s2 = 0.f;
      const float ln_mr = (p.lnp_part && p.lnp_mref && row < p.T) ? __ldg(p.lnp_mref + row) : 0.f;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const int b = j & 1;
        const int col = col_base + j * 32;
        mbar_wait(&my_ld[b], (j >> 1) & 1, 17);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          asm volatile("" : "+r"(rr[i]));
          v[i] = __uint_as_float(rr[i]);
        }
        if (j + 1 < 4) tmem_ld32(taddr0 + (uint32_t)((j + 1) * 32), rr);
        if (!p.backward) {
          const float4* b4 = reinterpret_cast<const float4*>(sb2 + half * (NC / 2) + j * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i];
            v[4 * i] = alpha * (v[4 * i] + bb.x); v[4 * i + 1] = alpha * (v[4 * i + 1] + bb.y);
            v[4 * i + 2] = alpha * (v[4 * i + 2] + bb.z); v[4 * i + 3] = alpha * (v[4 * i + 3] + bb.w);
          }
        }
        if (lane == 0) tma_store_wait_read<1>();   // the bf16 tile [b] was the source of chunk j-2's store
        __syncwarp();
        uint8_t* xrow = Xs + b * XT + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint8_t* q = xrow + ((c ^ xsw) << 4);
          const float4 x = *reinterpret_cast<const float4*>(q);
          v[4 * c] += x.x; v[4 * c + 1] += x.y; v[4 * c + 2] += x.z; v[4 * c + 3] += x.w;
          *reinterpret_cast<float4*>(q) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
        if (p.lnp_part) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] -= ln_mr;
            ln_s1 += v[i];
            ln_s2 = fmaf(v[i], v[i], ln_s2);
          }
        }
        if (p.has_out_bf16) {
          uint8_t* yrow = Ys + b * YT + r * 64;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(yrow + ((c ^ ysw) << 4)) =
                make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                           pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tm_out, Xs + b * XT, col, row0);
          if (p.has_out_bf16) tma_store_2d(&tm_outb, Ys + b * YT, col, row0);
          tma_store_commit();
          if (j + 2 < 4) {
            tma_store_wait_read<0>();   // in-place tile: fully read before the next residual lands in it
            issue_res(j + 2);
          }
        }
        if (warp == 0 && lane == 0) TL(10 + j);
      }
      if (lane == 0) tma_store_wait_read<0>();
      if (warp == 0 && lane == 0) TL(14);
      if (p.lnp_part && row < p.T)
        reinterpret_cast<float2*>(p.lnp_part)[(size_t)row * (p.E >> 7) + (col_base >> 7)] = make_float2(ln_s1, ln_s2);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == W_ALLOC) tmem_dealloc<1>(tmem_base, (uint32_t)TMEM_COLS);
  if (tl_on && threadIdx.x == 0) {
    TL(15);
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_adapter_timeline[31] = t;
  }
}

}  // namespace adp

// diagnostics: the phase stamps of CTA 0 of the last adapter launch (32 values; slots in tools/gemm_timeline.py)
int adapter_timeline(unsigned long long* out, int n) {
  FV_CHECK(n >= 32, "adapter timeline: need room for 32 values");
  FV_CUDA(cudaMemcpyFromSymbol(out, adp::g_adapter_timeline, sizeof(unsigned long long) * 32));
  return 0;
}

bool adapter_fused_supported(int T, int E, int A) {
  static int off = -1;
  if (off < 0) { const char* s = getenv("FERVIT_ADAPTER_FUSED"); off = (s && atoi(s) == 0) ? 1 : 0; }
  return off == 0 && T > 0 && A == adp::AD && E % adp::NC == 0 && E >= adp::NC;
}

// forward (backward == 0): in = x bf16 [T,E], Aw = W1 bf16 [64,E], Bw = W2 bf16 [E,64], res = x fp32, out = y fp32,
//                          s0 = g, s1 = d (both [T,64] bf16), b1 [64], b2 [E]
// backward (backward == 1): in = dy bf16 [T,E], Aw = W2^T bf16 [64,E], Bw = W1^T bf16 [E,64], res = dy fp32,
//                          out = dx fp32 (+ out_bf16), d_in = d, s0 = du
int adapter_fused(int backward, const bf16* in, const bf16* Aw, const bf16* Bw, const float* res, const float* b1,
                  const float* b2, const float* alpha_ptr, const bf16* d_in, bf16* s0, bf16* s1, float* out,
                  bf16* out_bf16, int T, int E, cudaStream_t stream, float* lnp_part, const float* lnp_mref) {
  FV_CHECK(!lnp_part || (out_bf16 && !backward), "adapter_fused: a LayerNorm producer is a forward pass with a bf16 copy");
  FV_CHECK(adapter_fused_supported(T, E, adp::AD), "adapter_fused: unsupported shape T=%d E=%d", T, E);
  FV_CHECK(in && Aw && Bw && res && alpha_ptr && s0 && out, "adapter_fused: null argument");
  FV_CHECK(backward ? (d_in != nullptr) : (b1 && b2 && s1), "adapter_fused: missing operand for this direction");
  CUtensorMap t_in, t_aw, t_bw, t_res, t_out, t_outb;
  FV_TRY(make_tmap_2d(&t_in, in, 2, (uint64_t)E, (uint64_t)T, (uint64_t)E, adp::BK, adp::BM, 128));
  FV_TRY(make_tmap_2d(&t_aw, Aw, 2, (uint64_t)E, (uint64_t)adp::AD, (uint64_t)E, adp::BK, adp::AD, 128));
  FV_TRY(make_tmap_2d(&t_bw, Bw, 2, (uint64_t)adp::AD, (uint64_t)E, (uint64_t)adp::AD, adp::AD, adp::NC, 128));
  FV_TRY(make_tmap_2d(&t_res, res, 4, (uint64_t)E, (uint64_t)T, (uint64_t)E, 32, 32, 128));
  FV_TRY(make_tmap_2d(&t_out, out, 4, (uint64_t)E, (uint64_t)T, (uint64_t)E, 32, 32, 128));
  t_outb = t_in;
  if (out_bf16) FV_TRY(make_tmap_2d(&t_outb, out_bf16, 2, (uint64_t)E, (uint64_t)T, (uint64_t)E, 32, 32, 64));
  adp::Params p;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("FERVIT_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  p.T = T; p.E = E; p.groups = E / adp::NC; p.backward = backward;
  p.b1 = b1; p.b2 = b2; p.alpha_ptr = alpha_ptr; p.d_in = d_in; p.s0 = s0; p.s1 = s1;
  p.has_out_bf16 = out_bf16 != nullptr;
  p.lnp_part = lnp_part; p.lnp_mref = lnp_mref;
  static bool attr_set = false;
  if (!attr_set) {
    FV_CUDA(cudaFuncSetAttribute(adp::adapter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, adp::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = ceil_div(T, adp::BM) * p.groups;
  // FLOPs of the two contractions (phase 1 counted once, as the GEMM pair it replaces)
  ProfScope prof(5, 4.0 * T * (double)E * adp::AD, stream);
  FV_CUDA(launch_pdl(adp::adapter_kernel, dim3(grid), dim3(adp::THREADS), (size_t)adp::SMEM_BYTES, stream, t_in, t_aw, t_bw,
                     t_res, t_out, t_outb, p));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

}  // namespace fervit
