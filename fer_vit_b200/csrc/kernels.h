// Internal launch wrappers implemented across the .cu files of this directory.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace fervit {

// gemm_tc.cu
int gemm_bf16_tc(const bf16* A, int lda, bool a_mn, const bf16* B, int ldb, bool b_mn, int M, int N, int K, int splits,
                 int force_bn, const Epilogue& epi, cudaStream_t stream);
int gemm_bf16_tc_effective_splits(int K, int splits);
// CTA-pair weight-gradient GEMM (gemm_wgrad2.cu): dW[M,N] = A^T B over K token rows, A [K,M], B [K,N], fp32 slabs
bool gemm_wgrad2_supported(int M, int N, int K, int lda, int ldb);
int gemm_wgrad2_splits(int M, int N, int K);
// colsum (optional): [2 * splits][M] slabs of the column sums of A over each split's token rows (the bias gradient)
int gemm_wgrad2(const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, int splits, int kb_per_split,
                float* out, const float* alpha_ptr, float alpha, cudaStream_t stream, float* colsum = nullptr);
// gemm_tc2.cu (CTA-pair kernel; K-major operands, TMA epilogue)
bool gemm_bf16_tc2_supported(int M, int N, int K, int lda, int ldb, const Epilogue& e, int kind);
int gemm_bf16_tc2(const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, int force_bn, const Epilogue& e,
                  int kind, cudaStream_t stream);
int gemm_tc2_clock_probe(double* ns, double* cycles);
int gemm_tc2_timeline(unsigned long long* out, int n);
int adapter_timeline(unsigned long long* out, int n);
int gemm_tc2_prof(int op, cudaStream_t stream);
int gemm_tc2_prof_read(double* us, double* flops, long long* launches, double* per_launch, int cap);
size_t gemm_tc2_scratch_bytes();
int gemm_tc2_set_scratch(void* ptr, size_t bytes);
int make_tmap_2d(CUtensorMap* map, const void* ptr, int esize, uint64_t inner, uint64_t outer, uint64_t ld,
                 uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);
// adapter_tc.cu: AdapterModule forward / input-gradient fused into one tensor-core kernel each (bf16, bottleneck 64)
bool adapter_fused_supported(int T, int E, int A);
int adapter_fused(int backward, const bf16* in, const bf16* Aw, const bf16* Bw, const float* res, const float* b1,
                  const float* b2, const float* alpha_ptr, const bf16* d_in, bf16* s0, bf16* s1, float* out,
                  bf16* out_bf16, int T, int E, cudaStream_t stream, float* lnp_part = nullptr,
                  const float* lnp_mref = nullptr);
// gemm_simt.cu
int gemm_f32_simt(const float* A, long long sam, long long sak, const float* B, long long sbn, long long sbk, int M,
                  int N, int K, int splits, const Epilogue& epi, cudaStream_t stream);
int gemm_f32_simt_effective_splits(int K, int splits);

// layernorm.cu
template <typename AT>
int layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int rows, int E, float* out_f32,
                  AT* out_at, float* mean, float* rstd, cudaStream_t stream);
int layernorm_bwd_grid(int rows);
template <typename DT, typename AT>
int layernorm_bwd(const DT* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                  const float* dres, int rows, int E, float* dx_f32, AT* dx_at, float* partial, Dropout at_drop,
                  cudaStream_t stream);

// elementwise.cu
template <typename AT> int cast_to_act(const float* src, AT* dst, size_t n, cudaStream_t stream);
int weight_cache(const float* src, int R, int C, bf16* dst, bf16* dst_t, cudaStream_t stream);
int weight_cache_batch(const float* const* src, const int* R, const int* C, bf16* const* dst, bf16* const* dst_t, int n,
                       cudaStream_t stream);
// W' = W diag(gamma) (bf16 + transposed), b' = b + W beta, cs = column sums of bf16 W' (LayerNorm folded into nn.Linear)
int fold_ln_weight(const float* W, const float* b, const float* gamma, const float* beta, int R, int C, bf16* dst,
                   bf16* dst_t, float* bfold, float* cs, cudaStream_t stream);
template <typename AT> int im2col(const float* x, AT* out, int B, int C, int H, int W, int P, cudaStream_t stream);
template <typename AT>
int cls_rows(const float* cls, const float* pos, float* x0, AT* x0_at, int B, int S, int E, Dropout drop,
             cudaStream_t stream);
template <typename AT>
int token_dropout(float* x0, AT* x0_at, int B, int S, int E, Dropout drop, cudaStream_t stream);
template <typename AT>
int gather_tokens(const float* dx0, AT* out, int B, int L, int E, Dropout drop, cudaStream_t stream);
int colsum_chunks(int R);
template <typename T>
int colsum(const T* in, int R, int C, long long ld, float* scratch, const float* alpha_ptr, float alpha, float* out,
           Dropout drop, cudaStream_t stream);
int colsum_reduce_partials(const float* partial, int chunks, int C, float* out, cudaStream_t stream,
                           float* out_hi = nullptr, int split = 0);
int splitk_reduce(const float* partial, int splits, size_t n, const float* alpha_ptr, float alpha, float* out,
                  cudaStream_t stream);
template <typename T>
int colsum_partial(const T* in, int R, int C, long long ld, float* scratch, cudaStream_t stream);
template <typename T>
int colsum_partial_rows(const T* in, int R, int C, long long ld, int rows_per_chunk, float* scratch, cudaStream_t stream);
// deferred AdapterModule gradient finalisation (see elementwise.cu)
constexpr int AD_FIN_MAX = 16;
struct AdapterGradJob {
  const float* w2_part;  // [s2][E*A] split-K slabs of dy^T g
  const float* w1_part;  // [s1][A*E] split-K slabs of du^T x
  const float* cs_dy;    // [chunks][E] per-chunk column sums of dy
  const float* cs_du;    // [chunks][A] per-chunk column sums of du
  const float* W2;       // [E, A] fp32 parameter (adapter.2.weight)
  const float* b2;       // [E]
  const float* alpha_ptr;
  float *dW2, *dW1, *db2, *db1, *dalpha;
  int s2, s1, chunks, E, A;
};
int adapter_grad_finalize_scratch_floats();
int adapter_grad_finalize(const AdapterGradJob* jobs, int n, float* scratch, cudaStream_t stream);
int dropout_mask(float* out, size_t n, Dropout drop, cudaStream_t stream);
int fill_f32(float* out, size_t n, float v, cudaStream_t stream);

// optim.cu
long long adamw_scratch_floats(int n, const long long* numel);
int adamw_step(int n, float* const* p, float* const* g, float* const* m, float* const* v, const long long* numel,
               const int* group, const float* hyper, float* step, float max_norm, float* scratch, cudaStream_t stream);

// attention.cu
template <typename AT>
int attention_fwd(const AT* qkv, AT* out, float* lse, int B, int S, int H, int HD, Dropout drop, cudaStream_t stream);
template <typename AT>
int attention_bwd(const AT* qkv, const AT* out, const AT* dout, const float* lse, AT* dqkv, int B, int S, int H,
                  int HD, Dropout drop, cudaStream_t stream);

// attention_tc.cu (bf16, S <= 32)
bool attention_tc_supported(int S, int HD);
int attention_tc_fwd(const bf16* qkv, bf16* out, float* lse, int B, int S, int H, int HD, Dropout drop,
                     cudaStream_t stream);
int attention_tc_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int B, int S,
                     int H, int HD, Dropout drop, cudaStream_t stream);

// attention_tc_long.cu (bf16, 32 < S <= 256)
bool attention_tc_long_supported(int S, int HD);
int attention_tc_long_fwd(const bf16* qkv, bf16* out, float* lse, int B, int S, int H, int HD, Dropout drop,
                          cudaStream_t stream);
int attention_tc_long_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int B,
                          int S, int H, int HD, Dropout drop, cudaStream_t stream);

// head_ce.cu
int head_fwd(const float* x, int B, int S, int E, const float* gamma, const float* beta, float eps, const float* W,
             const float* bias, int C, Dropout drop, float* logits, float* mean, float* rstd, cudaStream_t stream);
int head_wgrad_chunks(int B);
template <typename AT>
int head_bwd(const float* x, const float* dlogits, int B, int S, int E, const float* gamma, const float* beta,
             const float* W, int C, const float* mean, const float* rstd, Dropout drop, float* dx_f32, AT* dx_at,
             int wgrad, float* scratch, float* dW, float* dgamma, float* dbeta, float* dbias, cudaStream_t stream,
             cudaStream_t wgrad_stream);
int cross_entropy(const float* logits, const long long* labels, const float* weight, float smoothing, int B, int C,
                  const float* den_in, float grad_scale, float* loss, float* dlogits, float* den_out,
                  cudaStream_t stream);
int cross_entropy_mixup(const float* logits, const long long* labels, const long long* index, const float* weight,
                        float smoothing, int B, int C, float lam, const float* lam_dev, float grad_scale, float* loss,
                        float* dlogits, cudaStream_t stream);

// latent_batch.cu
int latent_batch(const float* latents, const long long* labels, long long n_rows, const long long* sample_idx, int B,
                 long long row, float noise_std, int use_scale, float scale_min, float scale_max, float mask_prob,
                 uint64_t seed, const unsigned long long* seed_dev, const long long* mix_index, double lam,
                 const float* lam_dev, float* out, long long* labels_out, int* status, cudaStream_t stream);

// latent_decompose.cu
int latent_decompose(const float* w, const float* dirs, int B, int C, long long row, int max_class, int output_mode,
                     float alpha, float* out, float* scores, cudaStream_t stream);

// premodules.cu
namespace pre {
struct Params {
  int use_spe, use_lwn, use_res, use_leam;
  const float* group_embed;  // [3, D]
  const float* layer_embed;  // [L, D]
  const long long* groups;   // [L]
  const float* gamma;        // [L, D]
  const float* beta;         // [L, D]
  const float* gate;         // [L]
  const float* leam_w;       // [L]
  float eps;
};
}  // namespace pre
typedef pre::Params PreParams;
template <typename AT>
int premodules_fwd(const float* x, int B, int L, int D, const PreParams& p, float* out_f32, AT* out_at,
                   cudaStream_t stream);
int premodules_chunks(int B);
template <typename DT>
int premodules_bwd(const float* x, const DT* dy, int B, int L, int D, const PreParams& p, float* dx, float* scratch,
                   float* dgamma, float* dbeta, float* dlayer, float* dgroup, float* dgate, float* dleam,
                   cudaStream_t stream);

}  // namespace fervit
