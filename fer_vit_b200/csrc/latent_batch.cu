// Latent batch producer: the data-side prologue of the LatentViT train step, in one HBM-bound kernel.
//
// Reference path (per step): `LatentFERDataset.__getitem__` loads one w+ latent [18,512] per sample and applies
// `LatentAugment` (data/latent_dataset.py:28-49: additive Gaussian noise, one scale factor per sample, element
// mask), the DataLoader stacks B of them, and the trainer blends the batch with a permuted copy of itself
// (train_latent_vit.py:119-127, mixup). Here the latents live packed in HBM ([N, row] fp32, 36 KB per sample) and one
// launch does gather -> augment -> mixup:
//
//   a[b]   = keep(b) * scale(b) * (latents[idx[b]] + noise_std * normal(b))     (each stage optional)
//   out[b] = lam * a[b] + (1 - lam) * a[mix_index[b]]
//
// Every random draw is a pure function of (seed, batch position, element), so the partner row a[mix_index[b]] is
// recomputed instead of staged (its latent is a second, mostly L2-resident read) and tests can replay the draws on the
// host. One 64-bit mix feeds a Box-Muller pair (two normals) or four 16-bit mask lanes, so a group of four elements
// costs three mixes. Multiplications and additions are kept un-contracted (__fmul_rn/__fadd_rn) so that, given the
// draws, the result is the reference's fp32 expression bit for bit.
// Algorithmic bytes per sample: read row*4 (+ row*4 partner with mixup) + write row*4.
#include "common.cuh"
#include "kernels.h"

namespace fervit {
namespace lbatch {

constexpr uint32_t SITE_NOISE = 0x4C410000u;   // idx = element pair: high word -> radius, low word -> angle
constexpr uint32_t SITE_SCALE = 0x4C410002u;   // idx = batch position
constexpr uint32_t SITE_MASK = 0x4C410003u;    // idx = element quad: four 16-bit lanes

struct Params {
  const float* latents;
  const long long* labels;
  const long long* sample_idx;
  const long long* mix_index;
  float* out;
  long long* labels_out;
  long long row;       // elements per sample, multiple of 4
  long long n_rows;    // rows in `latents` (bounds check of sample_idx)
  int B;
  float noise_std;
  float scale_min, scale_span;   // span < 0: no scaling
  uint32_t mask_threshold16;     // keep iff lane >= threshold; 0: no mask
  uint64_t seed;
  const unsigned long long* seed_dev;
  float lam, one_minus_lam;
  const float* lam_dev;
  int* status;                   // device flag: set to 1 on an out-of-range index
};

__device__ __forceinline__ float u01(uint32_t h) { return ((float)h + 0.5f) * (1.0f / 4294967296.0f); }   // (0, 1)

// augmentation of the four elements [e, e+4) of batch position b, applied to the loaded values v
__device__ __forceinline__ float4 augment(const Params& p, uint64_t seed, float scale, float4 v, int b, long long e) {
  const uint64_t base = (uint64_t)b * (uint64_t)p.row + (uint64_t)e;
  if (p.noise_std > 0.f) {
    float n[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint64_t z = mix_hash64(seed, SITE_NOISE, (base >> 1) + h);
      const float u1 = u01((uint32_t)(z >> 32));
      const float th = 6.283185307179586f * (u01((uint32_t)z) - 0.5f);      // (-pi, pi): MUFU sin/cos range
      const float rad = sqrtf(-1.3862943611198906f * __log2f(u1));         // sqrt(-2 ln u1)
      n[2 * h] = rad * __cosf(th);
      n[2 * h + 1] = rad * __sinf(th);
    }
    v.x = __fadd_rn(v.x, __fmul_rn(n[0], p.noise_std));
    v.y = __fadd_rn(v.y, __fmul_rn(n[1], p.noise_std));
    v.z = __fadd_rn(v.z, __fmul_rn(n[2], p.noise_std));
    v.w = __fadd_rn(v.w, __fmul_rn(n[3], p.noise_std));
  }
  if (p.scale_span >= 0.f) {
    v.x *= scale;
    v.y *= scale;
    v.z *= scale;
    v.w *= scale;
  }
  if (p.mask_threshold16) {
    const uint64_t z = mix_hash64(seed, SITE_MASK, base >> 2);
    v.x = ((uint32_t)(z) & 0xffffu) >= p.mask_threshold16 ? v.x : 0.f;
    v.y = ((uint32_t)(z >> 16) & 0xffffu) >= p.mask_threshold16 ? v.y : 0.f;
    v.z = ((uint32_t)(z >> 32) & 0xffffu) >= p.mask_threshold16 ? v.z : 0.f;
    v.w = ((uint32_t)(z >> 48)) >= p.mask_threshold16 ? v.w : 0.f;
  }
  return v;
}

__device__ __forceinline__ float sample_scale(const Params& p, uint64_t seed, int b) {
  return p.scale_span >= 0.f ? p.scale_min + p.scale_span * u01(mix_hash(seed, SITE_SCALE, (uint64_t)b)) : 1.0f;
}

__device__ __forceinline__ long long source_row(const Params& p, int b) {
  long long r = p.sample_idx ? p.sample_idx[b] : (long long)b;
  if (r < 0 || r >= p.n_rows) {
    if (p.status) atomicExch(p.status, 1);
    r = 0;
  }
  return r;
}

// One CTA walks (sample, 2048-element chunk) units: 256 threads x two float4 groups, all loads of a unit (own row and
// mixup partner) issued before any arithmetic.
template <bool MIX>
__global__ void __launch_bounds__(256)
latent_batch_kernel(const Params p) {
  const uint64_t seed = p.seed_dev ? p.seed + (uint64_t)(*p.seed_dev) : p.seed;
  const float lam = p.lam_dev ? p.lam_dev[0] : p.lam;
  const float om = p.lam_dev ? 1.0f - lam : p.one_minus_lam;
  const long long chunks = (p.row + 2047) / 2048;
  const long long units = (long long)p.B * chunks;
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const int b = (int)(u / chunks);
    const long long e0 = (u - (long long)b * chunks) * 2048 + (long long)threadIdx.x * 4;
    const long long r = source_row(p, b);
    int b2 = b;
    long long r2 = r;
    if (MIX) {
      long long m = p.mix_index[b];
      if (m < 0 || m >= p.B) {
        if (p.status) atomicExch(p.status, 1);
        m = b;
      }
      b2 = (int)m;
      r2 = source_row(p, b2);
    }
    float4 v[2], w[2];
    bool on[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const long long e = e0 + g * 1024;
      on[g] = e < p.row;
      if (on[g]) {
        v[g] = __ldg(reinterpret_cast<const float4*>(p.latents + r * p.row + e));
        if (MIX) w[g] = __ldg(reinterpret_cast<const float4*>(p.latents + r2 * p.row + e));
      }
    }
    const float s1 = sample_scale(p, seed, b);
    const float s2 = MIX ? sample_scale(p, seed, b2) : 1.0f;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (!on[g]) continue;
      const long long e = e0 + g * 1024;
      float4 a = augment(p, seed, s1, v[g], b, e);
      if (MIX) {
        const float4 c = augment(p, seed, s2, w[g], b2, e);
        // the reference expression lam * x + (1 - lam) * x[index], un-contracted
        a.x = __fadd_rn(__fmul_rn(lam, a.x), __fmul_rn(om, c.x));
        a.y = __fadd_rn(__fmul_rn(lam, a.y), __fmul_rn(om, c.y));
        a.z = __fadd_rn(__fmul_rn(lam, a.z), __fmul_rn(om, c.z));
        a.w = __fadd_rn(__fmul_rn(lam, a.w), __fmul_rn(om, c.w));
      }
      __stcs(reinterpret_cast<float4*>(p.out + (long long)b * p.row + e), a);
    }
    if (threadIdx.x == 0 && e0 == 0 && p.labels_out && p.labels) p.labels_out[b] = p.labels[r];
  }
}

}  // namespace lbatch

int latent_batch(const float* latents, const long long* labels, long long n_rows, const long long* sample_idx, int B,
                 long long row, float noise_std, int use_scale, float scale_min, float scale_max, float mask_prob,
                 uint64_t seed, const unsigned long long* seed_dev, const long long* mix_index, double lam,
                 const float* lam_dev, float* out, long long* labels_out, int* status, cudaStream_t stream) {
  FV_CHECK(B >= 1, "latent_batch: empty batch");
  FV_CHECK(row >= 4 && row % 4 == 0, "latent_batch: row length must be a positive multiple of 4 (got %lld)", row);
  FV_CHECK(((uintptr_t)latents & 15) == 0 && ((uintptr_t)out & 15) == 0, "latent_batch: 16-byte alignment required");
  FV_CHECK(noise_std >= 0.f && mask_prob >= 0.f && mask_prob < 1.f, "latent_batch: invalid augmentation parameter");
  FV_CHECK(!use_scale || scale_max >= scale_min, "latent_batch: scale_range must be (min, max) with max >= min");
  FV_CHECK((const void*)latents != (const void*)out, "latent_batch: in-place is not supported (rows are re-read)");
  lbatch::Params p;
  p.latents = latents;
  p.labels = labels;
  p.sample_idx = sample_idx;
  p.mix_index = mix_index;
  p.out = out;
  p.labels_out = labels_out;
  p.row = row;
  p.n_rows = n_rows;
  p.B = B;
  p.noise_std = noise_std;
  p.scale_min = scale_min;
  p.scale_span = use_scale ? (scale_max - scale_min) : -1.0f;
  p.mask_threshold16 = mask_prob > 0.f ? (uint32_t)((double)mask_prob * 65536.0) : 0u;
  if (mask_prob > 0.f && p.mask_threshold16 == 0) p.mask_threshold16 = 1;
  p.seed = seed;
  p.seed_dev = seed_dev;
  // python evaluates (1 - lam) in double before the tensor multiply rounds it to fp32
  p.lam = (float)lam;
  p.one_minus_lam = (float)(1.0 - lam);
  p.lam_dev = lam_dev;
  p.status = status;
  const long long units = (long long)B * ((row + 2047) / 2048);
  long long blocks = units;
  const long long cap = (long long)num_sms() * 8;      // 8 resident CTAs of 256 threads per SM, grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (mix_index)
    lbatch::latent_batch_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(p);
  else
    lbatch::latent_batch_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(p);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

}  // namespace fervit
