// LatentDecomposer (models_fer_vit/latent_decomposer.py:82-173): split a w+ latent into its projection on C fixed
// expression directions and the orthogonal rest, in one HBM-bound kernel.
//
//   coef[b,c] = <w[b], n_c>                                   (flattened latent, unit-norm directions)
//   w_expr[b] = sum_c coef[b,c] n_c   (all_classes)   |   coef[b,c*] n_c*,  c* = argmax_c |coef[b,c]|   (max_class)
//   w_id[b]   = w[b] - w_expr[b]
//   out       = w_expr | w_id | w_id + alpha * w_expr | [w_expr ; w_id]        (output_mode)
//
// The reference does two skinny matmuls ([B,9216]x[9216,C], [B,C]x[C,9216]) plus 1-3 elementwise passes; here a CTA
// keeps R = 4 latents in shared memory, streams the C directions from L2 once per pass for all four (the directions
// are 258 KB in total and stay L2-resident), reduces the 4 x C coefficients in the block and writes the output in the
// same launch. Algorithmic bytes per sample: read row*4, write row*4 (2*row*4 for concat). The input has no gradient
// and the directions are buffers, so there is no backward.
#include "common.cuh"
#include "kernels.h"

namespace fervit {
namespace ldec {

constexpr int R = 4;          // latents per CTA pass
constexpr int THREADS = 512;
constexpr int MAXC = 8;

struct Params {
  const float* w;        // [B, row]
  const float* dirs;     // [C, row]
  float* out;            // [B, row] or [B, 2, row] (concat); may be null (scores only)
  float* scores;         // [B, C] or null
  long long row;
  int B, C;
  int max_class;         // decompose_mode
  int output_mode;       // 0 expr_only, 1 id_only, 2 enhanced, 3 concat
  float alpha;
};

template <int C_>
__global__ void __launch_bounds__(THREADS)
latent_decompose_kernel(const Params p) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                                   // [R][row]
  __shared__ float s_part[THREADS / 32][R * C_];
  __shared__ float s_coef[R][C_];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long quads = p.row >> 2;
  const int groups = (p.B + R - 1) / R;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int b0 = grp * R;
    const int nb = min(R, p.B - b0);
    // pass 1: stage the latents, accumulate the R x C dot products
    float acc[R][C_];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < C_; ++c) acc[r][c] = 0.f;
    for (long long q = tid; q < quads; q += THREADS) {
      float4 x[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        x[r] = r < nb ? __ldcs(reinterpret_cast<const float4*>(p.w + (long long)(b0 + r) * p.row) + q)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(xs + (long long)r * p.row)[q] = x[r];
      }
#pragma unroll
      for (int c = 0; c < C_; ++c) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(p.dirs + (long long)c * p.row) + q);
#pragma unroll
        for (int r = 0; r < R; ++r)
          acc[r][c] += x[r].x * d.x + x[r].y * d.y + x[r].z * d.z + x[r].w * d.w;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < C_; ++c) {
        const float v = warp_sum(acc[r][c]);
        if (lane == 0) s_part[warp][r * C_ + c] = v;
      }
    __syncthreads();
    if (tid < R * C_) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) t += s_part[w][tid];      // fixed order: deterministic
      s_coef[tid / C_][tid % C_] = t;
      const int b = b0 + tid / C_;
      if (p.scores && b < p.B) p.scores[(long long)b * C_ + tid % C_] = t;
    }
    __syncthreads();
    if (p.out) {
      float coef[R][C_];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        int best = 0;
        float bestv = -1.f;
#pragma unroll
        for (int c = 0; c < C_; ++c) {
          coef[r][c] = s_coef[r][c];
          if (fabsf(coef[r][c]) > bestv) {       // first maximum, as torch.argmax
            bestv = fabsf(coef[r][c]);
            best = c;
          }
        }
        if (p.max_class) {
#pragma unroll
          for (int c = 0; c < C_; ++c) coef[r][c] = (c == best) ? coef[r][c] : 0.f;
        }
      }
      // pass 2: w_expr from the directions (L2) and the staged latents, output in the requested form
      const long long ostride = p.output_mode == 3 ? 2 * p.row : p.row;
      for (long long q = tid; q < quads; q += THREADS) {
        float4 e[R];
#pragma unroll
        for (int r = 0; r < R; ++r) e[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < C_; ++c) {
          const float4 d = __ldg(reinterpret_cast<const float4*>(p.dirs + (long long)c * p.row) + q);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            e[r].x += coef[r][c] * d.x;
            e[r].y += coef[r][c] * d.y;
            e[r].z += coef[r][c] * d.z;
            e[r].w += coef[r][c] * d.w;
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r >= nb) continue;
          float4* o = reinterpret_cast<float4*>(p.out + (long long)(b0 + r) * ostride) + q;
          if (p.output_mode == 0) {
            __stcs(o, e[r]);
            continue;
          }
          const float4 x = reinterpret_cast<const float4*>(xs + (long long)r * p.row)[q];
          const float4 id = make_float4(x.x - e[r].x, x.y - e[r].y, x.z - e[r].z, x.w - e[r].w);
          if (p.output_mode == 1) {
            __stcs(o, id);
          } else if (p.output_mode == 2) {
            __stcs(o, make_float4(id.x + p.alpha * e[r].x, id.y + p.alpha * e[r].y, id.z + p.alpha * e[r].z,
                                  id.w + p.alpha * e[r].w));
          } else {
            __stcs(o, e[r]);
            __stcs(o + quads, id);
          }
        }
      }
    }
    __syncthreads();      // xs and s_coef are reused by the next group
  }
}

template <int C_>
static int launch(const Params& p, cudaStream_t stream) {
  const size_t smem = (size_t)R * p.row * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    FV_CUDA(cudaFuncSetAttribute(latent_decompose_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int groups = (p.B + R - 1) / R;
  int blocks = groups < num_sms() ? groups : num_sms();
  latent_decompose_kernel<C_><<<blocks, THREADS, smem, stream>>>(p);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

}  // namespace ldec

int latent_decompose(const float* w, const float* dirs, int B, int C, long long row, int max_class, int output_mode,
                     float alpha, float* out, float* scores, cudaStream_t stream) {
  FV_CHECK(B >= 1, "latent_decompose: empty batch");
  FV_CHECK(C >= 1 && C <= ldec::MAXC, "latent_decompose: number of directions must be in [1, %d] (got %d)",
           ldec::MAXC, C);
  FV_CHECK(row >= 4 && row % 4 == 0, "latent_decompose: row length must be a positive multiple of 4 (got %lld)", row);
  FV_CHECK((size_t)ldec::R * row * sizeof(float) <= 200 * 1024,
           "latent_decompose: row of %lld elements does not fit the shared-memory staging (max 12800)", row);
  FV_CHECK(output_mode >= 0 && output_mode <= 3, "latent_decompose: unknown output_mode %d", output_mode);
  FV_CHECK(out || scores, "latent_decompose: nothing to compute");
  FV_CHECK(((uintptr_t)w & 15) == 0 && ((uintptr_t)dirs & 15) == 0 && ((uintptr_t)out & 15) == 0,
           "latent_decompose: 16-byte alignment required");
  ldec::Params p;
  p.w = w;
  p.dirs = dirs;
  p.out = out;
  p.scores = scores;
  p.row = row;
  p.B = B;
  p.C = C;
  p.max_class = max_class;
  p.output_mode = output_mode;
  p.alpha = alpha;
  switch (C) {
    case 1: return ldec::launch<1>(p, stream);
    case 2: return ldec::launch<2>(p, stream);
    case 3: return ldec::launch<3>(p, stream);
    case 4: return ldec::launch<4>(p, stream);
    case 5: return ldec::launch<5>(p, stream);
    case 6: return ldec::launch<6>(p, stream);
    case 7: return ldec::launch<7>(p, stream);
    default: return ldec::launch<8>(p, stream);
  }
}

}  // namespace fervit
