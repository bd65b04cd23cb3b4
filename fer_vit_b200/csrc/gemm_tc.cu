// bf16 GEMM on the 5th-generation tensor cores: TMA -> shared-memory ring -> tcgen05.mma into TMEM ->
// tcgen05.ld epilogue. One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer (one lane)
//   warp 1      MMA issuer   (one lane)
//   warp 2      TMEM allocator / deallocator
//   warps 4..11 epilogue: lane quarter = warp % 4, column half = (warp - 4) / 4
// Two TMEM accumulator buffers let the epilogue of tile i overlap the MMAs of tile i+1.
//
//   C[M,N] = A · B^T   with A "K-major"  [M,K] row-major   (activations)
//                            or "MN-major" [K,M] row-major   (wgrad: tokens are the reduction dim)
//                           B "K-major"  [N,K] row-major   (nn.Linear weight layout)
//                            or "MN-major" [K,N] row-major
// The drop-in replaces the cuBLAS calls behind nn.Linear / F.linear in the reference's blocks
// (timm Block via hybrid_latent_vit.py:227-233; nn.TransformerEncoderLayer via latent_vit.py:24-31).
#include "common.cuh"
#include "kernels.h"
#include "epilogue.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace fervit {

namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;       // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 128 + EPI_WARPS * 32;

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : ((BN == 128) ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // 128 / 256 / 512: powers of two
  static constexpr int BAR_BYTES = 256;
  // per-epilogue-warp transposition buffer: 32 rows x 16 columns fp32, row stride 20 words (conflict-free float4)
  static constexpr int EPI_LD = 20;
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_LD * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_BYTES + 1024;  // +1024: manual alignment
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch then fails loudly) instead of hanging the GPU. The slow path is
// kept out of line so the hot loops stay small.
__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("fervit gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor (sm_100 format, version 1), SWIZZLE_128B.
//   K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major: 64-element (128 B) MN groups `lbo` bytes apart, 8-k-row groups 1024 B apart (SBO).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, M = 128, N = BN, major-ness per operand.
__host__ __device__ constexpr uint32_t make_idesc(int bn, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Params {
  int M, N, K;        // logical problem
  int splits;         // split-K factor (1 = none); >1 writes fp32 partials to epi.out_f32 + split*M*N
  int kb_per_split;   // k-blocks per split
  int debug;          // bit mask for timing experiments (results are garbage): 1 skip TMA loads, 2 skip MMAs, 4 skip the epilogue, 8 epilogue = TMEM loads only
  Epilogue epi;
};

template <int BN, bool A_MN, bool B_MN, int KIND>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const Params p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  float* epi_stage = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (p.M + BM - 1) / BM;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int total_kb = (p.K + BK - 1) / BK;
  const int units = m_blocks * n_blocks * p.splits;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_grid_sync();  // the prologue above does not depend on the previous kernel's output

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = u % p.splits;
        const int t = u / p.splits;
        const int m_blk = t % m_blocks;
        const int n_blk = t / m_blocks;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.debug & 1) {
            mbar_arrive(&full_bar[stage]);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * C::A_BYTES;
          uint8_t* sb = smem_b + stage * C::B_BYTES;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sa + j * (64 * BK * 2), &tmap_a, &full_bar[stage], m_blk * BM + j * 64, kb * BK);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (64 * BK * 2), &tmap_b, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
          } else {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int split = u % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (p.debug & 2) {
            mbar_arrive(&empty_bar[stage]);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          const uint32_t a_addr = smem_u32(smem_a + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: advance 16 elements = 32 B inside the 128 B swizzle row.
            // MN-major: advance 16 k-rows of 128 B = 2048 B.
            const uint64_t adesc = A_MN ? make_smem_desc(a_addr + k * (UMMA_K * 128), 64 * BK * 2, 1024)
                                        : make_smem_desc(a_addr + k * (UMMA_K * 2), 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc(b_addr + k * (UMMA_K * 128), 64 * BK * 2, 1024)
                                        : make_smem_desc(b_addr + k * (UMMA_K * 2), 16, 1024);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (p.debug & 2) mbar_arrive(&tmem_full[buf]);
        else umma_commit(&tmem_full[buf]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;          // TMEM lanes 32*quarter .. +31 are the only ones this warp may read
    const int half = (warp - 4) >> 2;      // column half of the tile
    constexpr int COLS_PER_WARP = BN / 2;
    float alpha = p.epi.alpha;
    if (p.epi.alpha_ptr) alpha *= __ldg(p.epi.alpha_ptr);
    int it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const int split = u % p.splits;
      const int t = u / p.splits;
      const int m_blk = t % m_blocks;
      const int n_blk = t / m_blocks;
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      Epilogue e = p.epi;
      if (p.splits > 1) e.out_f32 = p.epi.out_f32 + (size_t)split * (size_t)p.M * (size_t)p.N;
      if (p.debug & 16) { e.out = nullptr; e.out_f32 = nullptr; e.out_pre = nullptr; }
      mbar_wait(&tmem_full[buf], acc_phase);
      tcgen05_fence_after();
      const int row0 = m_blk * BM + quarter * 32;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + half * COLS_PER_WARP);
      float* stg = epi_stage + (warp - 4) * (32 * C::EPI_LD);
      // Side inputs (residual, activation-derivative operand) are software-pipelined one 16-column chunk ahead, so
      // their global-load latency overlaps the TMEM load / transposition / math of the current chunk.
      constexpr bool PIPE = (KIND != EPK_REMAP && KIND != EPK_GENERIC);
      const int rl = lane >> 2, c4 = (lane & 3) * 4;
      const int col_base = n_blk * BN + half * COLS_PER_WARP;
      float4 aux_nx[4], res_nx[4];
      if (PIPE && col_base < p.N) {
#pragma unroll
        for (int itr = 0; itr < 4; ++itr)
          if (row0 + itr * 8 + rl < p.M)
            epilogue_prefetch<bf16, KIND>(e, row0 + itr * 8 + rl, col_base + c4, p.N, aux_nx[itr], res_nx[itr]);
      }
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 16) {
        const int col = col_base + c;
        if (col >= p.N) break;  // warp-uniform
        if (p.debug & 4) break;
        float v[16];
        tmem_ld16(taddr0 + (uint32_t)c, v);
        if (p.debug & 8) continue;
        float4 aux_cur[4], res_cur[4];
        if (PIPE) {
#pragma unroll
          for (int itr = 0; itr < 4; ++itr) { aux_cur[itr] = aux_nx[itr]; res_cur[itr] = res_nx[itr]; }
          if (c + 16 < COLS_PER_WARP && col + 16 < p.N) {
#pragma unroll
            for (int itr = 0; itr < 4; ++itr)
              if (row0 + itr * 8 + rl < p.M)
                epilogue_prefetch<bf16, KIND>(e, row0 + itr * 8 + rl, col + 16 + c4, p.N, aux_nx[itr], res_nx[itr]);
          }
        }
        // TMEM hands each lane one ROW of the tile; global memory wants lanes side by side along a row.
        // Transpose the 32x16 chunk through shared memory so every epilogue load/store below is coalesced.
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(stg + lane * C::EPI_LD + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        __syncwarp();
#pragma unroll
        for (int itr = 0; itr < 4; ++itr) {
          const int r = itr * 8 + rl;
          const float4 w = *reinterpret_cast<const float4*>(stg + r * C::EPI_LD + c4);
          float w4[4] = {w.x, w.y, w.z, w.w};
          if (p.debug & 32) { if (w4[0] == 1.2345e30f) e.out_f32[0] = w4[1]; continue; }
          if (row0 + r < p.M) {
            if (PIPE) epilogue_apply<bf16, 4, KIND>(e, alpha, row0 + r, col + c4, p.N, w4, &aux_cur[itr], &res_cur[itr]);
            else epilogue_apply<bf16, 4, KIND>(e, alpha, row0 + r, col + c4, p.N, w4);
          }
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ----------------------------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [outer, inner] matrix with leading dimension ld (elements);
// box = 64 inner elements (128 B, SWIZZLE_128B) x box_outer rows; out-of-range elements read as zero.
static int make_tmap(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  FV_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  FV_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "tcgen05 GEMM operand must be 16-byte aligned");
  FV_CHECK((ld * 2) % 16 == 0, "tcgen05 GEMM operand leading dimension must be a multiple of 8 elements (got %llu)",
           (unsigned long long)ld);
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {64u, box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FV_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

template <int BN, bool A_MN, bool B_MN, int KIND>
static int launch(const bf16* A, int lda, const bf16* B, int ldb, const Params& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  CUtensorMap ta, tb;
  if (A_MN) FV_TRY(make_tmap(&ta, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)lda, 64));
  else      FV_TRY(make_tmap(&ta, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda, BM));
  if (B_MN) FV_TRY(make_tmap(&tb, B, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldb, 64));
  else      FV_TRY(make_tmap(&tb, B, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldb, BN));
  static bool attr_set = false;
  if (!attr_set) {
    FV_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 C::SMEM_BYTES));
    attr_set = true;
  }
  const int m_blocks = ceil_div(p.M, BM), n_blocks = ceil_div(p.N, BN);
  const int units = m_blocks * n_blocks * p.splits;
  const int grid = units < num_sms() ? units : num_sms();
  ProfScope prof(4, 2.0 * p.M * (double)p.N * p.K, stream);
  FV_CUDA(launch_pdl(gemm_tc_kernel<BN, A_MN, B_MN, KIND>, dim3(grid), dim3(THREADS), (size_t)C::SMEM_BYTES, stream, ta,
                     tb, p));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

// Pick the N tile. Measured on B200 (tools/gemm_bench.py, FERVIT_GEMM_DEBUG=5): a cta_group::1 M=128 MMA takes the
// same ~128 cycles for N = 128 and N = 256 (1660 vs 870 TFLOP/s with loads and epilogue off), so the per-tile cost is
// ~ (k-blocks * 512 cycles + an epilogue term proportional to BN) whatever BN is: use the widest tile that is not
// mostly padding, and fall back to a narrower one only when it needs fewer waves.
static int choose_bn(int M, int N, int kblocks, int splits) {
  const int cands[3] = {256, 128, 64};
  long long best_cost = -1;
  int best = 256;
  const int sms = num_sms();
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn > 64 && N <= bn / 2) continue;  // mostly padding
    const long long tiles = (long long)ceil_div(M, BM) * ceil_div(N, bn) * splits;
    const long long waves = ceil_div_ll(tiles, sms);
    const long long cost = waves * ((long long)kblocks * 512 + 6 * bn + 600);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace tc

// C[M,N] (+epilogue) = op(A) op(B)^T on tcgen05. a_mn / b_mn select MN-major operands (see file header).
// splits > 1: fp32 partial sums land in epi.out_f32[split][M][N]; caller reduces them.
int gemm_bf16_tc(const bf16* A, int lda, bool a_mn, const bf16* B, int ldb, bool b_mn, int M, int N, int K,
                 int splits, int force_bn, const Epilogue& epi, cudaStream_t stream) {
  FV_CHECK(M > 0 && N > 0 && K > 0, "gemm_bf16_tc: empty problem M=%d N=%d K=%d", M, N, K);
  FV_CHECK(N % 16 == 0, "gemm_bf16_tc: N must be a multiple of 16 (got %d)", N);
  FV_CHECK(epi.ldo % 8 == 0 || (epi.out == nullptr && epi.ldo % 4 == 0), "gemm_bf16_tc: ldo must be a multiple of 8");
  tc::Params p;
  p.M = M; p.N = N; p.K = K;
  const int total_kb = ceil_div(K, tc::BK);
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  p.kb_per_split = ceil_div(total_kb, splits);
  p.splits = ceil_div(total_kb, p.kb_per_split);
  p.epi = epi;
  {
    // timing experiments only (tools/gemm_bench.py): re-read per call when the variable existed at the first call
    static int dbg = -1;
    if (dbg != 0) { const char* e = getenv("FERVIT_GEMM_DEBUG"); dbg = e ? (atoi(e) | (1 << 30)) : 0; }
    p.debug = dbg & ~(1 << 30);
  }
  if (p.splits > 1)
    FV_CHECK(epi.out_f32 != nullptr && epi.out == nullptr && epi.bias == nullptr && epi.residual == nullptr &&
                 epi.act == 0 && epi.act_bwd == 0 && epi.out_pre == nullptr && epi.remap_L == 0 && epi.ldo == N,
             "gemm_bf16_tc: split-K supports only a plain fp32 partial output");
  const int kind = epilogue_kind(epi);
  // forward / dgrad shapes go to the CTA-pair kernel (gemm_tc2.cu); this single-CTA kernel keeps the MN-major
  // (wgrad), split-K, row-remapping and dropout epilogues
  const int kind2 = epilogue_kind_nodrop(epi);   // the CTA-pair kernel takes dropout as a flag beside the kind
  if (!a_mn && !b_mn && p.splits == 1 && force_bn >= 0 && gemm_bf16_tc2_supported(M, N, K, lda, ldb, epi, kind2))
    return gemm_bf16_tc2(A, lda, B, ldb, M, N, K, force_bn, epi, kind2, stream);
  // weight gradients (both operands MN-major, plain fp32 slabs) of 256-wide shapes: the CTA-pair kernel of gemm_wgrad2.cu
  if (a_mn && b_mn && force_bn >= 0 && kind == EPK_PLAIN && epi.out_f32 != nullptr && epi.out == nullptr &&
      epi.bias == nullptr && epi.residual == nullptr && epi.remap_L == 0 && epi.ldo == N && !epi.drop.threshold &&
      gemm_wgrad2_supported(M, N, K, lda, ldb))
    return gemm_wgrad2(A, lda, B, ldb, M, N, K, p.splits, p.kb_per_split, epi.out_f32, epi.alpha_ptr, epi.alpha, stream);
  FV_CHECK(!epi.ln_part && !epi.lnp_part, "gemm_bf16_tc: a folded LayerNorm needs the CTA-pair kernel, which does not "
           "support this problem (M=%d N=%d K=%d)", M, N, K);
  if (force_bn < 0) force_bn = -force_bn;  // negative: force this kernel with that tile width (benchmarks)
  int bn = force_bn > 0 ? force_bn : tc::choose_bn(M, N, p.kb_per_split, p.splits);
  if (a_mn != b_mn) { set_error("gemm_bf16_tc: mixed operand major-ness is not instantiated"); return 1; }
#define FV_TC_KIND(BN_, K_) case K_: return tc::launch<BN_, false, false, K_>(A, lda, B, ldb, p, stream);
#define FV_TC_DISPATCH(BN_)                                                                  \
  do {                                                                                       \
    if (a_mn) {                                                                              \
      if (kind != EPK_PLAIN) { set_error("gemm_bf16_tc: wgrad supports a plain epilogue only"); return 1; } \
      return tc::launch<BN_, true, true, EPK_PLAIN>(A, lda, B, ldb, p, stream);              \
    }                                                                                        \
    switch (kind) {                                                                          \
      FV_TC_KIND(BN_, EPK_PLAIN) FV_TC_KIND(BN_, EPK_GELU) FV_TC_KIND(BN_, EPK_RELU)         \
      FV_TC_KIND(BN_, EPK_GELU_BWD) FV_TC_KIND(BN_, EPK_RELU_BWD) FV_TC_KIND(BN_, EPK_REMAP)  \
      default: return tc::launch<BN_, false, false, EPK_GENERIC>(A, lda, B, ldb, p, stream);  \
    }                                                                                        \
  } while (0)
  if (bn == 256) FV_TC_DISPATCH(256);
  if (bn == 128) FV_TC_DISPATCH(128);
  if (bn == 64) FV_TC_DISPATCH(64);
#undef FV_TC_KIND
#undef FV_TC_DISPATCH
  FV_CHECK(false, "gemm_bf16_tc: unsupported BN %d", bn);
}

// effective split count after clamping (callers size the partial buffer with it)
int gemm_bf16_tc_effective_splits(int K, int splits) {
  const int total_kb = ceil_div(K, tc::BK);
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  const int per = ceil_div(total_kb, splits);
  return ceil_div(total_kb, per);
}

}  // namespace fervit
