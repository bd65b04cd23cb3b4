// L2 -> shared-memory bandwidth probe for the tcgen05 GEMM's operand stream (no MMAs): is the operand ring bound by
// what an SM can ingest or by what L2 can put out, and does TMA multicast of the shared A tile lift that bound?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probes/tma_probe tools/probes/tma_probe.cu -lcuda
//   tools/probes/tma_probe            # prints one JSON line per mode
//
// Every CTA streams the operand boxes of a 4864 x 2304 x 768 GEMM exactly as gemm_tc2 does (per k-block: a 128-row x
// 64-column bf16 box of A and one of B, 16 KB each, SWIZZLE_128B, ring of STAGES slots) and drops them.
//   mode 0  cluster 1: every CTA loads its own A box and B box (what the GEMM does today)
//   mode 1  cluster 2: the two CTAs need the SAME A box; each loads half of it (64 rows) and multicasts to both
//   mode 2  cluster 4: four CTAs share the A box (32 rows each, multicast to all four)
//   mode 3  cluster 2: the two CTAs need the same A box and both load all of it (unicast duplicates)
//   mode 6  cluster 1, two producer lanes (one issues the A boxes, one the B boxes)
//   mode 7  cluster 1, stages of TWO k-blocks (4 boxes = 64 KB per iteration, 3 stages): per-iteration or per-byte bound?
//   mode 4  cluster 1: A only (16 KB per k-block)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int STAGES = 6;
constexpr int BOX_ROWS = 128, BOX_COLS = 64;
constexpr int BOX_BYTES = BOX_ROWS * BOX_COLS * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void arrive_remote(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait(uint64_t* b, uint32_t parity, int tag) {
  const long long t0 = clock64();
  while (!try_wait(b, parity)) {
    if (clock64() - t0 > 2000000000ll) { printf("probe: wait timed out tag %d block %d\n", tag, blockIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

struct Args { int mode, cluster, tiles, kblocks, m_blocks, n_blocks; unsigned long long* out; };

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_as,
             const __grid_constant__ CUtensorMap tm_b, const Args a) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * 2 * BOX_BYTES);
  uint64_t* empty = full + STAGES;
  const uint32_t rank = a.cluster > 1 ? ctarank() : 0;
  const int cs = a.cluster;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], a.mode == 6 ? 2 : 1); mbar_init(&empty[s], cs); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (cs > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const int total = a.tiles * a.kblocks;
  if (threadIdx.x == 0 || (a.mode == 6 && threadIdx.x == 64)) {
    // ---- producer (one lane, like the GEMM's TMA warp; mode 6: a second lane issues the B boxes) ----
    const int group = blockIdx.x / cs;              // CTAs of one cluster share the A rows of `group`
    const int ngroups = gridDim.x / cs;
    const bool two = a.mode == 6;
    const bool do_a = !two || threadIdx.x == 0, do_b = a.mode != 4 && (!two || threadIdx.x == 64);
    const uint32_t bytes = (do_a ? BOX_BYTES : 0) + (do_b ? BOX_BYTES : 0);
    const int sub = BOX_ROWS / cs;                  // rows of the shared A box this CTA fetches (multicast modes)
    int s = 0;
    uint32_t ph = 0;
    if (a.mode == 7) {
      const int NST = STAGES / 2;
      for (int tile = 0; tile < a.tiles; ++tile) {
        const int unit = group + tile * ngroups;
        const int mb = (unit / a.n_blocks) % a.m_blocks, nb = unit % a.n_blocks;
        const int row_a = mb * BOX_ROWS, row_b = (nb % (a.n_blocks * 2)) * BOX_ROWS;
        for (int kb = 0; kb < a.kblocks; kb += 2) {
          if (tile > 0 || kb >= 2 * NST) wait(&empty[s], ph ^ 1, 1);
          expect_tx(&full[s], 4 * BOX_BYTES);
          uint8_t* sa = smem + s * 4 * BOX_BYTES;
          tma_load(sa, &tm_a, &full[s], kb * BOX_COLS, row_a);
          tma_load(sa + BOX_BYTES, &tm_a, &full[s], (kb + 1) * BOX_COLS, row_a);
          tma_load(sa + 2 * BOX_BYTES, &tm_b, &full[s], kb * BOX_COLS, row_b);
          tma_load(sa + 3 * BOX_BYTES, &tm_b, &full[s], (kb + 1) * BOX_COLS, row_b);
          if (++s == NST) { s = 0; ph ^= 1; }
        }
      }
    } else
    for (int tile = 0; tile < a.tiles; ++tile) {
      const int unit = group + tile * ngroups;                 // like the GEMM: n fastest
      const int mb = (unit / a.n_blocks) % a.m_blocks, nb = unit % a.n_blocks;
      const int row_a = mb * BOX_ROWS + ((a.mode == 1 || a.mode == 2) ? (int)rank * sub : 0);
      const int row_b = ((nb * cs + (int)rank) % (a.n_blocks * 2)) * BOX_ROWS;
      for (int kb = 0; kb < a.kblocks; ++kb) {
        if (tile > 0 || kb >= STAGES) wait(&empty[s], ph ^ 1, 1);
        expect_tx(&full[s], bytes);
        uint8_t* sa = smem + s * 2 * BOX_BYTES;
        if (do_a) {
          if (a.mode == 1 || a.mode == 2)
            tma_load_mc(sa + rank * sub * 128, &tm_as, &full[s], kb * BOX_COLS, row_a, (uint16_t)((1u << cs) - 1));
          else
            tma_load(sa, &tm_a, &full[s], kb * BOX_COLS, row_a);
        }
        if (do_b) tma_load(sa + BOX_BYTES, &tm_b, &full[s], kb * BOX_COLS, row_b);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (threadIdx.x == 32) {
    // ---- consumer (one lane, like the GEMM's MMA warp): drop the stage as soon as it has landed ----
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    uint32_t eaddr[4];
    for (int r = 0; r < cs; ++r) eaddr[r] = 0;
    const int nst = a.mode == 7 ? STAGES / 2 : STAGES;
    const int iters = a.mode == 7 ? total / 2 : total;
    for (int j = 0; j < iters; ++j) {
      const int sj = j % nst;
      wait(&full[sj], (j / nst) & 1, 2);
      for (int r = 0; r < cs; ++r) arrive_remote(mapa(smem_u32(&empty[sj]), r));
    }
    const long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    a.out[blockIdx.x * 2] = t1 - t0;
    a.out[blockIdx.x * 2 + 1] = (unsigned long long)(c1 - c0);
  }
  __syncthreads();
  if (cs > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn fn, void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * 2};
  cuuint32_t box[2] = {BOX_COLS, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn fn = (EncodeFn)fp;
  const int M = 4864, N = 2304, K = 768;
  __nv_bfloat16 *A, *B;
  CK(cudaMalloc(&A, (size_t)M * K * 2));
  CK(cudaMalloc(&B, (size_t)N * K * 2));
  CK(cudaMemset(A, 0, (size_t)M * K * 2));
  CK(cudaMemset(B, 0, (size_t)N * K * 2));
  unsigned long long* out;
  CK(cudaMalloc(&out, 148 * 2 * 8));
  const int smem = STAGES * 2 * BOX_BYTES + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int modes[7][2] = {{0, 1}, {1, 2}, {2, 4}, {3, 2}, {4, 1}, {6, 1}, {7, 1}};
  for (int rep = 0; rep < 2; ++rep)
    for (int mi = 0; mi < 7; ++mi) {
      const int mode = modes[mi][0], cs = modes[mi][1];
      int grid = 148 / cs * cs;
      if (cs == 4) grid = 144;
      CUtensorMap ta = make_map(fn, A, M, K, BOX_ROWS), tas = make_map(fn, A, M, K, BOX_ROWS / cs),
                  tb = make_map(fn, B, N, K, BOX_ROWS);
      Args a{mode, cs, 16, K / BOX_COLS, M / BOX_ROWS, N / 256, out};
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      CK(cudaLaunchKernelEx(&cfg, probe_kernel, ta, tas, tb, a));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      unsigned long long h[148 * 2];
      CK(cudaMemcpy(h, out, grid * 16, cudaMemcpyDeviceToHost));
      double ns = 0, cyc = 0;
      for (int i = 0; i < grid; ++i) { if (h[2 * i] > ns) ns = (double)h[2 * i]; if (h[2 * i + 1] > cyc) cyc = (double)h[2 * i + 1]; }
      const double per_cta = 16.0 * (K / BOX_COLS) * (mode == 4 ? 1 : 2) * BOX_BYTES;
      if (rep == 1)
        printf("{\"mode\": %d, \"cluster\": %d, \"grid\": %d, \"kernel_us\": %.2f, \"event_us\": %.2f, \"ingest_bytes_per_cta\": %.0f, "
               "\"ingest_B_per_clk_per_sm\": %.1f, \"ingest_TBps_chip\": %.2f, \"sm_ghz\": %.3f}\n",
               mode, cs, grid, ns / 1e3, ms * 1e3, per_cta, per_cta / cyc, per_cta * grid / ns / 1e3, cyc / ns);
    }
  return 0;
}
