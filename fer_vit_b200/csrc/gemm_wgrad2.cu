// Weight-gradient GEMM on a CTA pair: dW[M, N] (+)= dY^T X with the TOKENS as the reduction dimension,
//   A = dY [K rows, M columns] and B = X [K rows, N columns], both row-major in global memory, i.e. both operands
//   MN-major for the tensor core (no transposes are ever materialised), split-K over the token rows, fp32 partial
//   slabs out[split][M][N] reduced by the caller in a fixed order (deterministic).
// The same CTA-pair machinery as gemm_tc2.cu (cluster of 2, tcgen05.mma.cta_group::2 with M = 256, N = 256, one MMA lane
// in the leader, TMA loads of both CTAs counted on the leader's mbarrier, TMEM accumulators double-buffered) with
//   * MN-major operand tiles: a k-block of an operand is two TMA boxes of [64 k-rows x 64 columns] (SWIZZLE_128B: 64
//     rows of 128 bytes), the shared-memory descriptor walks them with LBO = one box, SBO = 8 k-rows, and a UMMA_K step
//     is 16 k-rows = 2048 bytes;
//   * work units (256 x 256 tile, k-split), persistent round-robin over the pairs;
//   * an epilogue that only moves the accumulator: tcgen05.ld (thread = row, 32 columns) -> 128-byte row stores.
// (A split-K reduction INSIDE the launch - the last CTA of a tile's splits re-reading all slabs in split order - was built
// and measured: 24 -> 97 us per launch. Every thread's __threadfence per unit and a reduction with the epilogue's
// thread-per-row access pattern on two SMs cost far more than the separate, coalesced 6 us kernel on all of them.)
// It takes the all-parameters-trainable models' weight gradients (configs 1 / 2 / 4: 25 launches per step) from the
// single-CTA kernel of gemm_tc.cu, which keeps the shapes this one does not cover (M or N below 256: the adapters).
#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace fervit {

int make_tmap_2d(CUtensorMap* map, const void* ptr, int esize, uint64_t inner, uint64_t outer, uint64_t ld,
                 uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);   // gemm_tc2.cu

namespace wg2 {

using namespace ptx;

constexpr int BM = 128;          // rows of dW per CTA; the pair covers 256
constexpr int BN = 256;          // columns of dW per pair tile; each CTA stages 128 of them
constexpr int BK = 64;           // token rows per k-block
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int COL_WARPS = 4;      // optional bias-gradient warps: column sums of the dY tile while it sits in shared memory
constexpr int THREADS = EPI_WARPS * 32 + 128 + COL_WARPS * 32;
// the single-lane roles keep the highest warp ids (the sub-partition arbiter prefers them); the column-sum warps sit below
constexpr int W_COL = EPI_WARPS, W_TMA = EPI_WARPS + COL_WARPS, W_MMA = W_TMA + 1, W_ALLOC = W_TMA + 2;
constexpr int BOX_BYTES = 64 * BK * 2;            // one [64 k-rows x 64 columns] box: 8 KB
constexpr int A_BYTES = (BM / 64) * BOX_BYTES;    // 16 KB
constexpr int B_BYTES = (BN / 2 / 64) * BOX_BYTES;  // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGES = 6;
constexpr int BAR_BYTES = 256;   // 3 * STAGES + 4 barriers + the TMEM base slot
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 2 * BN;
static_assert(SMEM_BYTES <= 232448, "wgrad2: shared memory");

struct Params {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_per_split;
  float* out;              // [splits][M][N] fp32
  const float* alpha_ptr;  // optional device scalar (splits == 1 only)
  float alpha;
  // optional: column sums of A = dY over the token rows of each split, colsum[2 * split + h][M] (the bias gradient's
  // slabs; h = which 32 rows of every 64-row k-block). The
  // dY tile is in shared memory anyway: four extra warps add it up per k-block (units of n-tile 0 only) instead of a
  // separate kernel streaming dY from HBM a second time.
  float* colsum;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
wgrad2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]  the leader's is the live one
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]  per CTA
  uint64_t* tmem_full = bars + 2 * STAGES;       // [2]       per CTA
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]       the leader's is the live one
  uint64_t* landed_bar = bars + 2 * STAGES + 4;  // [STAGES]  per CTA: "this stage's operands are in shared memory"
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int total_kb = (p.K + BK - 1) / BK;
  const int units = p.m_tiles * p.n_tiles * p.splits;

  pdl_trigger();
  if (warp == W_TMA && lane == 0) {
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
  }
  if (warp == W_MMA && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], p.colsum ? 1 + COL_WARPS : 1);   // the MMAs' commit (+ this CTA's column-sum warps)
      mbar_init(&landed_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 2 * EPI_WARPS);   // the epilogue warps of BOTH CTAs
    }
    mbar_fence_init();
  }
  if (warp == W_ALLOC) {
    tmem_alloc<2>(tmem_base_slot, (uint32_t)TMEM_COLS);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_grid_sync();   // the prologue above does not depend on the previous kernel's output

  // unit u -> (tile, split); the splits of a tile are adjacent, so the pairs running at one time share operand columns
  auto decode = [&](int u, int& mt, int& nt, int& ka, int& ke) {
    const int split = u % p.splits, t = u / p.splits;
    nt = t % p.n_tiles;
    mt = t / p.n_tiles;
    ka = split * p.kb_per_split;
    ke = min(total_kb, ka + p.kb_per_split);
    return split;
  };

  if (warp == W_TMA) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = pair_id; u < units; u += num_pairs) {
        int mt, nt, ka, ke;
        decode(u, mt, nt, ka, ke);
        const int col_a = mt * (2 * BM) + (int)rank * BM;       // this CTA's 128 rows of dW = columns of dY
        const int col_b = nt * BN + (int)rank * (BN / 2);       // this CTA's half of the tile's columns = columns of X
        for (int kb = ka; kb < ke; ++kb) {
          mbar_wait_parked(&empty_bar[stage], phase ^ 1, 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const uint32_t full0 = mapa(smem_u32(&full_bar[stage]), 0);
          uint8_t* sa = smem_a + stage * A_BYTES;
          uint8_t* sb = smem_b + stage * B_BYTES;
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sa + j * BOX_BYTES, &tm_a, full0, col_a + j * 64, kb * BK);
#pragma unroll
          for (int j = 0; j < BN / 2 / 64; ++j) tma_load_2d_pair(sb + j * BOX_BYTES, &tm_b, full0, col_b + j * 64, kb * BK);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(2 * BM, BN, true, true);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = pair_id; u < units; u += num_pairs, ++it) {
        int mt, nt, ka, ke;
        decode(u, mt, nt, ka, ke);
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait_parked(&tmem_empty[buf], acc_phase ^ 1, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = ka; kb < ke; ++kb) {
          mbar_wait_parked(&full_bar[stage], phase, 3);
          tc_fence_after();
          if (p.colsum) {   // only the leader's barrier sees the TMA bytes: pass "landed" on to both CTAs' warps
            mbar_arrive(&landed_bar[stage]);
            mbar_arrive_cluster(mapa(smem_u32(&landed_bar[stage]), 1));
          }
          const uint32_t a_addr = smem_u32(smem_a + stage * A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // MN-major: 64-element groups one box apart (LBO), 8 k-rows = 1024 B (SBO); 16 k-rows = 2048 B per step
            const uint64_t adesc = make_smem_desc(a_addr + k * (UMMA_K * 128), BOX_BYTES, 1024);
            const uint64_t bdesc = make_smem_desc(b_addr + k * (UMMA_K * 128), BOX_BYTES, 1024);
            umma_bf16<2>(tmem_d, adesc, bdesc, idesc, (kb > ka || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage], 3);   // frees the slot in both CTAs once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tmem_full[buf], 3);       // accumulator complete, both CTAs
      }
    }
  } else if (warp < EPI_WARPS) {
    // ===================== epilogue (both CTAs): accumulator -> fp32 slab =====================
    const int quarter = warp & 3;   // TMEM lanes 32*quarter .. +31 are the only ones this warp may read
    const int half = warp >> 2;     // column half of the tile
    const uint32_t tmem_empty0[2] = {mapa(smem_u32(&tmem_empty[0]), 0), mapa(smem_u32(&tmem_empty[1]), 0)};
    float alpha = p.alpha;
    if (p.alpha_ptr) alpha *= __ldg(p.alpha_ptr);
    int it = 0;
    for (int u = pair_id; u < units; u += num_pairs, ++it) {
      int mt, nt, ka, ke;
      const int split = decode(u, mt, nt, ka, ke);
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = mt * (2 * BM) + (int)rank * BM + quarter * 32 + lane;
      const int col0 = nt * BN + half * (BN / 2);
      float* orow = p.out + ((size_t)split * p.M + (size_t)row) * p.N;
      mbar_wait_cluster(&tmem_full[buf], acc_phase, 4);
      tc_fence_after();
      const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + half * (BN / 2));
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 32) {
        uint32_t rr[32];
        tmem_ld32(ta + (uint32_t)c, rr);
        tmem_ld_wait();
        if (row < p.M) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int col = col0 + c + 4 * i;
            if (col < p.N)
              *reinterpret_cast<float4*>(orow + col) =
                  make_float4(__uint_as_float(rr[4 * i]) * alpha, __uint_as_float(rr[4 * i + 1]) * alpha,
                              __uint_as_float(rr[4 * i + 2]) * alpha, __uint_as_float(rr[4 * i + 3]) * alpha);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty0[buf]);
    }
  } else if (warp >= W_COL && warp < W_COL + COL_WARPS && p.colsum) {
    // ===================== column sums of the dY tile (both CTAs) =====================
    // warp cw: box cw / 2 (64 columns = 32 lanes x 2), rows (cw & 1) * 32 .. +32 of every k-block; a k-block is 64 rows
    // of 128 bytes whose 16-byte chunks are XOR-swizzled by row & 7 (SWIZZLE_128B). The two row halves go to two slabs.
    const int cw = warp - W_COL;
    const int c = 2 * lane;
    const uint32_t chunk = (uint32_t)(c >> 3), within = (uint32_t)(c & 7) * 2;
    const int r0 = (cw & 1) * 32;
    int stage = 0;
    uint32_t phase = 0;
    for (int u = pair_id; u < units; u += num_pairs) {
      int mt, nt, ka, ke;
      const int split = decode(u, mt, nt, ka, ke);
      float2 acc = make_float2(0.f, 0.f);
      for (int kb = ka; kb < ke; ++kb) {
        mbar_wait_parked(&landed_bar[stage], phase, 5);
        if (nt == 0) {
          const uint8_t* box = smem_a + stage * A_BYTES + (cw >> 1) * BOX_BYTES;
#pragma unroll 16
          for (int r = r0; r < r0 + 32; ++r) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(box + r * 128 + ((chunk ^ (uint32_t)(r & 7)) << 4) + within);
            const float2 f = unpack_bf16x2(v);
            acc.x += f.x;
            acc.y += f.y;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (nt == 0) {
        const int col = mt * (2 * BM) + (int)rank * BM + (cw >> 1) * 64 + c;
        if (col < p.M)   // M is a multiple of 64: col + 1 < M too
          *reinterpret_cast<float2*>(p.colsum + ((size_t)split * 2 + (cw & 1)) * p.M + col) = acc;
      }
    }
  }

  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  if (warp == W_ALLOC) tmem_dealloc<2>(tmem_base, (uint32_t)TMEM_COLS);
}

static bool enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FERVIT_WGRAD2"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on == 1;
}

}  // namespace wg2

// dW [M, N] from A = dY [K, M] (ld lda) and B = X [K, N] (ld ldb): shapes the CTA-pair kernel takes
bool gemm_wgrad2_supported(int M, int N, int K, int lda, int ldb) {
  // K (token rows) >= 2048: below that a unit is a handful of k-blocks and the pair kernel's longer prologue and the
  // extra slabs cost more than its MMA rate returns (config 1, 608 rows: 1.52 -> 1.61 ms per step when it took them)
  return wg2::enabled() && M >= 256 && N >= 256 && M % 64 == 0 && N % 64 == 0 && K >= 2048 && lda % 8 == 0 &&
         ldb % 8 == 0;
}

// preferred split-K factor for such a shape (0: not supported): about one unit per CTA pair, at least 4 k-blocks each
int gemm_wgrad2_splits(int M, int N, int K) {
  if (!gemm_wgrad2_supported(M, N, K, M, N)) return 0;
  const int tiles = ceil_div(M, 2 * wg2::BM) * ceil_div(N, wg2::BN);
  const int pairs = num_sms() / 2;
  int s = pairs / tiles;
  const int max_s = ceil_div(K, 4 * wg2::BK);
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  const int total_kb = ceil_div(K, wg2::BK);
  const int per = ceil_div(total_kb, s);
  return ceil_div(total_kb, per);
}

int gemm_wgrad2(const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, int splits, int kb_per_split,
                float* out, const float* alpha_ptr, float alpha, cudaStream_t stream, float* colsum) {
  FV_CHECK(gemm_wgrad2_supported(M, N, K, lda, ldb), "gemm_wgrad2: unsupported problem M=%d N=%d K=%d", M, N, K);
  FV_CHECK(splits >= 1 && kb_per_split >= 1 && (long long)splits * kb_per_split >= ceil_div(K, wg2::BK),
           "gemm_wgrad2: the splits do not cover K");
  FV_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0 && N % 4 == 0, "gemm_wgrad2: output must be 16-byte aligned");
  CUtensorMap ta, tb;
  FV_TRY(make_tmap_2d(&ta, A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, wg2::BK, 128));
  FV_TRY(make_tmap_2d(&tb, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, wg2::BK, 128));
  wg2::Params p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = ceil_div(M, 2 * wg2::BM);
  p.n_tiles = ceil_div(N, wg2::BN);
  p.splits = splits;
  p.kb_per_split = kb_per_split;
  p.out = out;
  p.alpha_ptr = alpha_ptr;
  p.alpha = alpha;
  p.colsum = colsum;
  static bool attr_set = false;
  if (!attr_set) {
    FV_CUDA(cudaFuncSetAttribute(wg2::wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg2::SMEM_BYTES));
    attr_set = true;
  }
  const int units = p.m_tiles * p.n_tiles * splits;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (units < pairs ? units : pairs);
  ProfScope prof(4, 2.0 * M * (double)N * K, stream);
  FV_CUDA(launch_pdl(wg2::wgrad2_kernel, dim3(grid), dim3(wg2::THREADS), (size_t)wg2::SMEM_BYTES, stream, ta, tb, p));
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

}  // namespace fervit
