// Shared device/host helpers for the fervit_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace fervit {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// Error handling: every C-ABI entry returns 0 on success; the message of the last failure is kept
// per host thread and read back with fervit_last_error().
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define FV_CHECK(cond, ...)                \
  do {                                     \
    if (!(cond)) {                         \
      ::fervit::set_error(__VA_ARGS__);    \
      return 1;                            \
    }                                      \
  } while (0)

#define FV_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::fervit::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                          __FILE__, __LINE__);                                          \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

#define FV_LAUNCH_CHECK() FV_CUDA(cudaGetLastError())

#define FV_TRY(expr)          \
  do {                        \
    int _s = (expr);          \
    if (_s != 0) return _s;   \
  } while (0)

// number of kernels launched by this library since load (bench.py reports it as gpu_launches)
extern unsigned long long g_launch_count;
#define FV_COUNT_LAUNCH() (++::fervit::g_launch_count)

int num_sms();

// Programmatic dependent launch (PDL): kernels launched through launch_pdl() may be scheduled while the previous
// kernel of the stream is still draining; they call pdl_grid_sync() before touching global memory, so only their
// launch latency and their input-independent prologue (barrier init, TMEM allocation, descriptor prefetch) overlap the
// predecessor's tail. FERVIT_PDL=0 turns the launch attribute off (the device-side calls then return immediately).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#if defined(__CUDACC__)
// let the next kernel of the stream start launching; then wait until the previous one has completed and flushed
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_grid_sync() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// per-kernel-class event timing (runtime.cu); classes: 0 tcgen05 GEMM (flops), 1 attention (bytes),
// 2 LayerNorm (bytes), 3 fp32 GEMM (flops)
bool prof_enabled();
bool prof_serial();
int prof_open(int cls, double work, cudaStream_t st);
void prof_close(int id, cudaStream_t st);
void prof_enable(int mode);
int prof_read(int cls, double* ms, double* work, long long* count);
struct ProfScope {
  int id; cudaStream_t st;
  ProfScope(int cls, double work, cudaStream_t s) : id(-1), st(s) { if (prof_enabled()) id = prof_open(cls, work, s); }
  ~ProfScope() { if (id >= 0) prof_close(id, st); }
};

// ---------------------------------------------------------------------------------------------
// Activation ids used by GEMM epilogues and the oracle alike.
// ---------------------------------------------------------------------------------------------
// ACT_DERIV (backward only): aux already holds act'(pre), saved by a forward epilogue with pre_is_deriv set
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_DERIV = 3 };

__device__ __forceinline__ float gelu_fwd(float x) {
  // exact-erf GELU: torch nn.GELU() default, timm Mlp, AdapterModule (hybrid_latent_vit.py:259)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_bwd(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// bf16-mode epilogues: erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16's 2^-9 output rounding),
// sharing exp(-u^2/2) between the cdf and the pdf: ~15 instructions instead of ~45 for erff + expf.
//   erf(x) = 1 - (a1 t + ... + a5 t^5) exp(-x^2), t = 1 / (1 + p x), x >= 0
__device__ __forceinline__ void gelu_fast_parts(float u, float& cdf, float& pdf_times_sqrt2pi) {
  const float au = fabsf(u);
  const float t = __fdividef(1.0f, fmaf(0.3275911f * 0.70710678118654752440f, au, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;                                                              // exp(-u^2 / 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u * u * (-0.5f * 1.44269504088896340736f)));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  cdf = 0.5f + copysignf(0.5f * erf_abs, u);
  pdf_times_sqrt2pi = e;
}
__device__ __forceinline__ float gelu_fwd_fast(float u) {
  float cdf, e;
  gelu_fast_parts(u, cdf, e);
  return u * cdf;
}
__device__ __forceinline__ float gelu_bwd_fast(float u) {
  float cdf, e;
  gelu_fast_parts(u, cdf, e);
  return fmaf(u * 0.39894228040143267794f, e, cdf);
}
// Tensor-core GEMM epilogues (gemm_tc2.cu): MUFU-free polynomials, because at ~23 instructions + 2 MUFU per element
// the erf form above made the fc1 / fc2-dgrad epilogues slower than the MMAs they hide behind (measured).
//   Phi(u)   = 0.5 + uc * P8(uc^2),  uc = clamp(u, -4, 4): |error| <= 7e-6 inside, <= 3.2e-5 in the clamped tails
//   gelu'(u) = 0.5 + uc * Q9(uc^2):                        |error| <= 6e-5 inside, <= 5.4e-4 in the clamped tails
// (minimax fits of the exact erf forms; bf16 outputs round at 2^-9 = 2e-3 relative).
__host__ __device__ __forceinline__ float gelu_fwd_poly(float u) {
  const float uc = fminf(fmaxf(u, -4.0f), 4.0f);
  const float s = uc * uc;
  float p = 3.463219783e-01f / 4294967296.0f;                    // coefficients of (s/16)^k, k = 8 .. 0
  p = fmaf(p, s, -1.879982349e+00f / 268435456.0f);
  p = fmaf(p, s, 4.556960448e+00f / 16777216.0f);
  p = fmaf(p, s, -6.600791621e+00f / 1048576.0f);
  p = fmaf(p, s, 6.482043223e+00f / 65536.0f);
  p = fmaf(p, s, -4.644546053e+00f / 4096.0f);
  p = fmaf(p, s, 2.528634271e+00f / 256.0f);
  p = fmaf(p, s, -1.062569537e+00f / 16.0f);
  p = fmaf(p, s, 3.989227100e-01f);
  return u * fmaf(uc, p, 0.5f);
}
// GELU and its derivative in one pass (forward epilogues that save act'(u) for the backward GEMM instead of u):
// Phi by the polynomial above, the density by one ex2 on the otherwise idle MUFU pipe.
__device__ __forceinline__ void gelu_fwd_deriv_poly(float u, float& g, float& d) {
  const float uc = fminf(fmaxf(u, -4.0f), 4.0f);
  const float s = uc * uc;
  float p = 3.463219783e-01f / 4294967296.0f;
  p = fmaf(p, s, -1.879982349e+00f / 268435456.0f);
  p = fmaf(p, s, 4.556960448e+00f / 16777216.0f);
  p = fmaf(p, s, -6.600791621e+00f / 1048576.0f);
  p = fmaf(p, s, 6.482043223e+00f / 65536.0f);
  p = fmaf(p, s, -4.644546053e+00f / 4096.0f);
  p = fmaf(p, s, 2.528634271e+00f / 256.0f);
  p = fmaf(p, s, -1.062569537e+00f / 16.0f);
  p = fmaf(p, s, 3.989227100e-01f);
  const float cdf = fmaf(uc, p, 0.5f);
  float e;  // exp(-u^2 / 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u * u * (-0.5f * 1.44269504088896340736f)));
  g = u * cdf;
  d = fmaf(u * 0.39894228040143267794f, e, cdf);
}
// Scalar mirror of the packed (f32x2) epilogue math below, op for op (same fmaf order, same saturating final FMA): the
// host-side test (tests/host/mix_hash_host.cu) checks it against erf, the kernels use the packed form.
__host__ __device__ __forceinline__ void gelu_fwd_deriv_packed_mirror(float u, float& g, float& d) {
  const float s = u * u;
  const float sc = fminf(s, 16.0f);
  float p = 3.463219783e-01f / 4294967296.0f;
  p = fmaf(p, sc, -1.879982349e+00f / 268435456.0f);
  p = fmaf(p, sc, 4.556960448e+00f / 16777216.0f);
  p = fmaf(p, sc, -6.600791621e+00f / 1048576.0f);
  p = fmaf(p, sc, 6.482043223e+00f / 65536.0f);
  p = fmaf(p, sc, -4.644546053e+00f / 4096.0f);
  p = fmaf(p, sc, 2.528634271e+00f / 256.0f);
  p = fmaf(p, sc, -1.062569537e+00f / 16.0f);
  p = fmaf(p, sc, 3.989227100e-01f);
  const float cdf = fminf(fmaxf(fmaf(u, p, 0.5f), 0.0f), 1.0f);   // fma.rn.sat: |u| > 4 overshoots and clamps
  const float e = exp2f(s * (-0.5f * 1.44269504088896340736f));   // device: ex2.approx
  g = u * cdf;
  d = fmaf(u * 0.39894228040143267794f, e, cdf);
}
// Two elements per instruction: Blackwell's packed fp32 pipe (fma/mul.rn.f32x2 -> FFMA2 / FMUL2, same IEEE fp32 results
// as the scalar forms above). The GELU GEMM epilogues are bound by warp-instruction ISSUE, not by fp32 throughput
// (profiles/r02_gemm_timeline.jsonl: 6.2 us of epilogue per 128 x 256 tile against 4.2-4.8 us of MMAs); packing halves
// the issue slots of the polynomial: 9.5 instead of 19 per element. The clamp of u is replaced by ONE min on u^2 and a
// saturating final FMA: for |u| > 4, u * P8(16) overshoots +-0.5, and .sat clamps the cdf to [0, 1] (exact there to
// 3.2e-5, as before).
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
#define FV_F2C(v) f2_pack((v), (v))
// cdf Phi(u) of two elements; s = u^2 (both) is returned for the density
__device__ __forceinline__ void gelu_cdf2(float u0, float u1, unsigned long long u, unsigned long long& s, float& c0,
                                          float& c1) {
  s = f2_mul(u, u);
  float s0, s1;
  f2_unpack(s, s0, s1);
  const unsigned long long sc = f2_pack(fminf(s0, 16.0f), fminf(s1, 16.0f));
  // Horner: an Estrin split (dependency depth 4 instead of 8) was measured SLOWER here — 2.0 instead of 1.5 us of
  // math per 64-column chunk: the epilogue is bound by the fp32 pipe and by registers, not by the FMA latency chain
  unsigned long long p = FV_F2C(3.463219783e-01f / 4294967296.0f);
  p = f2_fma(p, sc, FV_F2C(-1.879982349e+00f / 268435456.0f));
  p = f2_fma(p, sc, FV_F2C(4.556960448e+00f / 16777216.0f));
  p = f2_fma(p, sc, FV_F2C(-6.600791621e+00f / 1048576.0f));
  p = f2_fma(p, sc, FV_F2C(6.482043223e+00f / 65536.0f));
  p = f2_fma(p, sc, FV_F2C(-4.644546053e+00f / 4096.0f));
  p = f2_fma(p, sc, FV_F2C(2.528634271e+00f / 256.0f));
  p = f2_fma(p, sc, FV_F2C(-1.062569537e+00f / 16.0f));
  p = f2_fma(p, sc, FV_F2C(3.989227100e-01f));
  float p0, p1;
  f2_unpack(p, p0, p1);
  asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(c0) : "f"(u0), "f"(p0));
  asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(c1) : "f"(u1), "f"(p1));
}
// (x0, x1) <- GELU(x0, x1)
__device__ __forceinline__ void gelu_fwd_poly2(float& x0, float& x1) {
  const unsigned long long u = f2_pack(x0, x1);
  unsigned long long s;
  float c0, c1;
  gelu_cdf2(x0, x1, u, s, c0, c1);
  f2_unpack(f2_mul(u, f2_pack(c0, c1)), x0, x1);
}
// (x0, x1) <- GELU(x0, x1), (d0, d1) <- GELU'(x0, x1)
__device__ __forceinline__ void gelu_fwd_deriv_poly2(float& x0, float& x1, float& d0, float& d1) {
  const unsigned long long u = f2_pack(x0, x1);
  unsigned long long s;
  float c0, c1;
  gelu_cdf2(x0, x1, u, s, c0, c1);
  float t0, t1, e0, e1;
  f2_unpack(f2_mul(s, FV_F2C(-0.5f * 1.44269504088896340736f)), t0, t1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));   // exp(-u^2 / 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  const unsigned long long cdf = f2_pack(c0, c1);
  const unsigned long long g = f2_mul(u, cdf);
  const unsigned long long d = f2_fma(f2_mul(u, FV_F2C(0.39894228040143267794f)), f2_pack(e0, e1), cdf);
  f2_unpack(g, x0, x1);
  f2_unpack(d, d0, d1);
}
#endif
__host__ __device__ __forceinline__ float gelu_bwd_poly(float u) {
  const float uc = fminf(fmaxf(u, -4.0f), 4.0f);
  const float s = uc * uc;
  float p = -3.606366035e+00f / 68719476736.0f;                  // (s/16)^k, k = 9 .. 0
  p = fmaf(p, s, 2.116166659e+01f / 4294967296.0f);
  p = fmaf(p, s, -5.557563405e+01f / 268435456.0f);
  p = fmaf(p, s, 8.697387332e+01f / 16777216.0f);
  p = fmaf(p, s, -9.125917374e+01f / 1048576.0f);
  p = fmaf(p, s, 6.839210087e+01f / 65536.0f);
  p = fmaf(p, s, -3.771883499e+01f / 4096.0f);
  p = fmaf(p, s, 1.521041010e+01f / 256.0f);
  p = fmaf(p, s, -4.250744890e+00f / 16.0f);
  p = fmaf(p, s, 7.978260665e-01f);
  return fmaf(uc, p, 0.5f);
}
__device__ __forceinline__ float act_fwd(int act, float x) {
  if (act == ACT_RELU) return fmaxf(x, 0.0f);
  if (act == ACT_GELU) return gelu_fwd(x);
  return x;
}
__device__ __forceinline__ float act_bwd(int act, float pre) {
  if (act == ACT_DERIV) return pre;
  if (act == ACT_RELU) return pre > 0.0f ? 1.0f : 0.0f;
  if (act == ACT_GELU) return gelu_bwd(pre);
  return 1.0f;
}

// ---------------------------------------------------------------------------------------------
// Counter-based dropout: keep(seed, site, idx) is a pure function, so backward recomputes the
// mask instead of storing it, and tests can materialise the very same mask for the oracle.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix_hash64(uint64_t seed, uint32_t site, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(site + 1) + idx * 0xD1B54A32D192ED03ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t mix_hash(uint64_t seed, uint32_t site, uint64_t idx) {
  return (uint32_t)(mix_hash64(seed, site, idx) >> 32);
}
// The dropout generator proper. mix_hash64 costs ~25 integer instructions per element (three 64-bit multiplies), which
// made the mask the largest instruction consumer of the attention kernels with dropout on (85 % of the S = 197 forward
// kernel's instructions, profiles/r02_ncu_attn_long.txt). drop_hash keeps the 64-bit mix for the KEY only - one value
// per (seed, site), loop-invariant in every kernel - and spends two 32-bit multiply-xorshift rounds per element
// (lowbias32 constants), the second key word entering between the rounds so that different keys give different
// functions, not translations of one. ~10 instructions per element.
__host__ __device__ __forceinline__ uint32_t drop_hash(uint64_t seed, uint32_t site, uint64_t idx) {
  const uint64_t key = mix_hash64(seed, site, 0x5DEECE66Dull);
  uint32_t x = (uint32_t)idx * 0x9E3779B1u + (uint32_t)key;
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= (uint32_t)(key >> 32) + (uint32_t)(idx >> 32) * 0x85EBCA6Bu;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
// threshold = round(p * 2^32); keep iff hash >= threshold
__host__ __device__ __forceinline__ bool drop_keep(uint64_t seed, uint32_t site, uint64_t idx,
                                                   uint32_t threshold) {
  return drop_hash(seed, site, idx) >= threshold;
}
static inline uint32_t drop_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t <= 0) return 0u;
  if (t >= 4294967295.0) return 4294967295u;
  return (uint32_t)t;
}

struct Dropout {       // passed by value to kernels; p == 0 disables
  uint64_t seed;
  // optional device-resident counter added to `seed`: lets a captured CUDA graph draw a fresh mask per replay
  const unsigned long long* seed_ptr;
  uint32_t site;
  uint32_t threshold;  // 0 = off
  float scale;         // 1/(1-p)
  __device__ __forceinline__ uint64_t eff() const { return seed_ptr ? seed + (uint64_t)(*seed_ptr) : seed; }
};
static inline Dropout make_dropout(float p, uint64_t seed, uint32_t site,
                                   const unsigned long long* seed_ptr = nullptr) {
  Dropout d;
  d.seed = seed;
  d.seed_ptr = seed_ptr;
  d.site = site;
  d.threshold = (p > 0.f) ? drop_threshold(p) : 0u;
  d.scale = (p > 0.f) ? 1.0f / (1.0f - p) : 1.0f;
  return d;
}

// ---------------------------------------------------------------------------------------------
// GEMM epilogue contract shared by the SIMT fp32 GEMM and the tcgen05 bf16 GEMM.
//   v = acc (+ bias[col]); v *= act'(aux[row,col]) if act_bwd; v *= alpha; pre = v; v = act(v);
//   v = dropout(v); v += residual[orow,col]; v += pos[1 + row % remap_L, col]
//   orow = row (or (row / L) * (L+1) + 1 + row % L when remap_L > 0: token rows skip the cls slot)
// ---------------------------------------------------------------------------------------------
struct Epilogue {
  const float* bias;       // [N] or null
  const float* residual;   // fp32 [Mout, ldo] or null
  const void* aux;         // activation-dtype [M, N] pre-activation for act_bwd, or null
  const float* alpha_ptr;  // device scalar or null
  float alpha;             // host scalar (1.0 = none)
  int act;                 // Act applied forward
  int act_bwd;             // Act whose derivative at aux multiplies the accumulator
  void* out;               // activation-dtype output [Mout, ldo] or null
  float* out_f32;          // fp32 output [Mout, ldo] or null
  void* out_pre;           // activation-dtype pre-activation output [M, N] or null
  int pre_is_deriv;        // 1: out_pre receives act'(pre) instead of pre (the backward GEMM then uses ACT_DERIV)
  int remap_L;             // 0 = rows map 1:1
  const float* pos;        // [(L+1), N] position rows, used with remap_L
  int ldo;                 // leading dimension of out / out_f32 / residual (elements)
  Dropout drop;
  void* sk_ws;             // stream-K scratch of the CTA-pair GEMM (gemm_tc2.cu), zero-initialised flags first; or null
  size_t sk_bytes;
  // ---- LayerNorm folded into the GEMMs either side of it (bf16 mode, frozen norm + frozen weight) ----
  // y = LN(x) W^T + b with LN(x) = (x - mu) rstd gamma + beta is computed as
  //     y_j = rstd * (acc_j - delta * cs_j) + b'_j,   acc = A W'^T,  A = bf16(x - mref),  W' = W diag(gamma),
  //     delta = mu - mref,  cs_j = sum_k W'_jk (of the bf16-rounded W'),  b' = b + W beta  (passed as `bias`).
  // CONSUMER (this GEMM applies the norm): the row statistics come from per-128-column partial sums the producer
  // left; the exact mean / rstd are written out for the LayerNorm backward.
  const float* ln_part;    // [M, ln_parts, 2] {sum, sum of squares} of (x - mref) per 128-column part; null = no fold
  const float* ln_mref;    // [M] what the producer subtracted (null = 0)
  const float* ln_cs;      // [N]
  float* ln_mean;          // [M] out
  float* ln_rstd;          // [M] out
  float ln_eps;
  int ln_parts;            // E / 128
  // PRODUCER (this GEMM's fp32 output rows are the input of a folded norm): `out` receives bf16(v - mref[row]) and
  // the partial sums of this GEMM's 128-column parts are written (each exactly once: deterministic)
  float* lnp_part;         // [M, N / 128, 2] out; null = plain output
  const float* lnp_mref;   // [M] or null
};

static inline Epilogue make_epilogue() {
  Epilogue e;
  memset(&e, 0, sizeof(e));
  e.alpha = 1.0f;
  return e;
}

// ---------------------------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  float2 r;
  r.x = __uint_as_float(v << 16);
  r.y = __uint_as_float(v & 0xffff0000u);
  return r;
}

// Load / store 4 consecutive activation values (16-byte aligned for float, 8-byte for bf16).
template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 load4<bf16>(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace fervit
