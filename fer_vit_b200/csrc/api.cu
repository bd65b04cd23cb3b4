// C-ABI entry points other than the plan (plan.cu): library info, loss, stand-alone pre-modules and the
// operator-level calls the parity tests use. See include/fervit_b200.h for the contracts.
#include "common.cuh"
#include "kernels.h"
#include "fervit_b200.h"

using namespace fervit;

#define FV_API extern "C" __attribute__((visibility("default")))

static inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }

FV_API int fervit_abi_version(void) { return FERVIT_ABI_VERSION; }
FV_API const char* fervit_last_error(void) { return get_error(); }
FV_API unsigned long long fervit_launch_count(void) { return g_launch_count; }

FV_API int fervit_profile_enable(int on) {
  prof_enable(on);
  return 0;
}
FV_API int fervit_profile_read(int kernel_class, double* ms, double* work, long long* launches) {
  FV_CHECK(ms && work && launches, "profile_read: null argument");
  FV_CHECK(prof_read(kernel_class, ms, work, launches) == 0, "profile_read: CUDA event query failed");
  return 0;
}

FV_API int fervit_cross_entropy(const float* logits, const long long* labels, const float* weight,
                                float label_smoothing, int B, int C, const float* den_in, float grad_scale,
                                float* loss, float* dlogits, float* den_out, void* stream) {
  FV_CHECK(logits && labels && loss, "cross_entropy: null argument");
  return cross_entropy(logits, labels, weight, label_smoothing, B, C, den_in, grad_scale, loss, dlogits, den_out,
                       S_(stream));
}

FV_API int fervit_cross_entropy_mixup(const float* logits, const long long* labels, const long long* mix_index,
                                      const float* weight, float label_smoothing, int B, int C, float lam,
                                      const float* lam_dev, float grad_scale, float* loss, float* dlogits,
                                      void* stream) {
  FV_CHECK(logits && labels && mix_index && loss, "cross_entropy_mixup: null argument");
  return cross_entropy_mixup(logits, labels, mix_index, weight, label_smoothing, B, C, lam, lam_dev, grad_scale, loss,
                             dlogits, S_(stream));
}

FV_API int fervit_latent_batch(const float* latents, const long long* labels, long long n_rows,
                               const long long* sample_idx, int B, long long row_elems,
                               const fervit_latent_augment* aug, unsigned long long seed,
                               const unsigned long long* seed_dev, const long long* mix_index, double lam,
                               const float* lam_dev, float* out, long long* labels_out, int* status, void* stream) {
  FV_CHECK(latents && out, "latent_batch: null argument");
  FV_CHECK(n_rows >= 1, "latent_batch: empty latent table");
  FV_CHECK(sample_idx || B <= n_rows, "latent_batch: batch larger than the latent table and no sample_idx");
  FV_CHECK(!labels_out || labels, "latent_batch: labels_out needs labels");
  FV_CHECK(!mix_index || lam_dev || (lam >= 0.0 && lam <= 1.0), "latent_batch: lam must be in [0, 1] (got %g)", lam);
  return latent_batch(latents, labels, n_rows, sample_idx, B, row_elems, aug ? aug->noise_std : 0.f,
                      aug ? aug->use_scale : 0, aug ? aug->scale_min : 1.f, aug ? aug->scale_max : 1.f,
                      aug ? aug->mask_prob : 0.f, seed, seed_dev, mix_index, lam, lam_dev, out, labels_out, status,
                      S_(stream));
}

FV_API int fervit_latent_decompose(const float* w_plus, const float* directions, int B, int C, long long row_elems,
                                   int decompose_mode, int output_mode, float enhance_alpha, float* out,
                                   float* scores, void* stream) {
  FV_CHECK(w_plus && directions, "latent_decompose: null argument");
  FV_CHECK(decompose_mode == 0 || decompose_mode == 1, "latent_decompose: unknown decompose_mode %d", decompose_mode);
  return latent_decompose(w_plus, directions, B, C, row_elems, decompose_mode, output_mode, enhance_alpha, out, scores,
                          S_(stream));
}

static PreParams to_pre(const fervit_premodules* p) {
  PreParams q;
  q.use_spe = p->use_spe; q.use_lwn = p->use_lwn; q.use_res = p->use_lwn_res; q.use_leam = p->use_leam;
  q.group_embed = p->group_embed; q.layer_embed = p->layer_embed; q.groups = p->groups;
  q.gamma = p->gamma; q.beta = p->beta; q.gate = p->gate; q.leam_w = p->leam_w; q.eps = p->eps;
  return q;
}

FV_API int fervit_premodules_forward(const fervit_premodules* p, const float* x, int B, int L, int D, float* y,
                                     void* stream) {
  FV_CHECK(p && x && y, "premodules_forward: null argument");
  return premodules_fwd<float>(x, B, L, D, to_pre(p), y, nullptr, S_(stream));
}

FV_API long long fervit_premodules_scratch_floats(int B, int L, int D) {
  return (long long)premodules_chunks(B) * L * (3LL * D + 2);
}

FV_API int fervit_premodules_backward(const fervit_premodules* p, const float* x, const float* dy, int B, int L,
                                      int D, float* dx, float* scratch, float* dgamma, float* dbeta,
                                      float* dlayer_embed, float* dgroup_embed, float* dgate, float* dleam,
                                      void* stream) {
  FV_CHECK(p && x && dy && scratch, "premodules_backward: null argument");
  if (p->use_lwn) FV_CHECK(dgamma && dbeta, "premodules_backward: LWN gradients required");
  if (p->use_lwn && p->use_lwn_res) FV_CHECK(dgate != nullptr, "premodules_backward: gate gradient required");
  if (p->use_spe) FV_CHECK(dlayer_embed && dgroup_embed, "premodules_backward: SPE gradients required");
  if (p->use_leam) FV_CHECK(dleam != nullptr, "premodules_backward: LEAM gradient required");
  return premodules_bwd<float>(x, dy, B, L, D, to_pre(p), dx, scratch, dgamma, dbeta, dlayer_embed, dgroup_embed,
                               dgate, dleam, S_(stream));
}

FV_API int fervit_linear_forward(int act_dtype, const void* x, const void* W, const float* bias,
                                 const float* residual, int M, int N, int K, int act, void* out, float* out_f32,
                                 void* pre, int force_bn, void* stream) {
  FV_CHECK(x && W, "linear_forward: null argument");
  Epilogue e = make_epilogue();
  e.bias = bias; e.residual = residual; e.act = act & 0xff; e.out = out; e.out_f32 = out_f32; e.out_pre = pre; e.ldo = N;
  e.pre_is_deriv = (act & 0x100) ? 1 : 0;
  if (act_dtype == FERVIT_F32)
    return gemm_f32_simt((const float*)x, K, 1, (const float*)W, K, 1, M, N, K, 1, e, S_(stream));
  FV_CHECK(act_dtype == FERVIT_BF16, "linear_forward: unknown dtype %d", act_dtype);
  return gemm_bf16_tc((const bf16*)x, K, false, (const bf16*)W, K, false, M, N, K, 1, force_bn, e, S_(stream));
}

FV_API int fervit_adapter_forward(const void* x_bf16, const float* x_f32, const void* W1, const float* b1, const void* W2,
                                  const float* b2, const float* alpha, int T, int E, void* g, void* d, float* y,
                                  void* stream) {
  FV_CHECK(adapter_fused_supported(T, E, 64), "adapter_forward: needs E a multiple of 256 and bottleneck 64 (T=%d E=%d)",
           T, E);
  return adapter_fused(0, (const bf16*)x_bf16, (const bf16*)W1, (const bf16*)W2, x_f32, b1, b2, alpha, nullptr, (bf16*)g,
                       (bf16*)d, y, nullptr, T, E, S_(stream));
}

FV_API int fervit_adapter_backward_input(const void* dy_bf16, const float* dy_f32, const void* W2t, const void* W1t,
                                         const float* alpha, const void* d, int T, int E, void* du, float* dx,
                                         void* dx_bf16, void* stream) {
  FV_CHECK(adapter_fused_supported(T, E, 64), "adapter_backward_input: needs E a multiple of 256 and bottleneck 64");
  return adapter_fused(1, (const bf16*)dy_bf16, (const bf16*)W2t, (const bf16*)W1t, dy_f32, nullptr, nullptr, alpha,
                       (const bf16*)d, (bf16*)du, nullptr, dx, (bf16*)dx_bf16, T, E, S_(stream));
}

FV_API long long fervit_adamw_scratch_floats(int n, const long long* numel) {
  return (n > 0 && numel) ? adamw_scratch_floats(n, numel) : 8;
}

FV_API int fervit_adamw_step(int n, void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                             const long long* numel, const int* group, const float* hyper, float* step, float max_norm,
                             float* scratch, void* stream) {
  FV_CHECK(n == 0 || (params && grads && exp_avg && exp_avg_sq && numel && group), "adamw_step: null argument");
  return adamw_step(n, reinterpret_cast<float* const*>(params), reinterpret_cast<float* const*>(grads),
                    reinterpret_cast<float* const*>(exp_avg), reinterpret_cast<float* const*>(exp_avg_sq), numel, group,
                    hyper, step, max_norm, scratch, S_(stream));
}

FV_API long long fervit_gemm_scratch_bytes(void) { return (long long)gemm_tc2_scratch_bytes(); }
FV_API int fervit_set_gemm_scratch(void* ptr, long long bytes) { return gemm_tc2_set_scratch(ptr, (size_t)bytes); }

FV_API int fervit_debug_gemm_clock(double* ns, double* cycles) {
  FV_CHECK(ns && cycles, "debug_gemm_clock: null argument");
  return gemm_tc2_clock_probe(ns, cycles);
}

FV_API int fervit_gemm_prof(int op, void* stream) { return gemm_tc2_prof(op, S_(stream)); }
FV_API int fervit_gemm_prof_read(double* us, double* flops, long long* launches, double* per_launch, int cap) {
  FV_CHECK(us && flops && launches, "gemm_prof_read: null argument");
  return gemm_tc2_prof_read(us, flops, launches, per_launch, cap);
}

FV_API int fervit_debug_gemm_timeline(unsigned long long* out, int n) {
  FV_CHECK(out, "debug_gemm_timeline: null argument");
  return gemm_tc2_timeline(out, n);
}
FV_API int fervit_debug_adapter_timeline(unsigned long long* out, int n) {
  FV_CHECK(out, "debug_adapter_timeline: null argument");
  return adapter_timeline(out, n);
}

FV_API int fervit_linear_dgrad(int act_dtype, const void* dy, const void* Wt, const void* aux, const float* residual,
                               int M, int N, int K, int act, void* out, float* out_f32, int force_bn, void* stream) {
  FV_CHECK(dy && Wt, "linear_dgrad: null argument");
  FV_CHECK(act == ACT_NONE || aux != nullptr, "linear_dgrad: aux (pre-activation) is required with an activation");
  Epilogue e = make_epilogue();
  e.residual = residual; e.act_bwd = act; e.aux = aux; e.out = out; e.out_f32 = out_f32; e.ldo = K;
  // GEMM view: rows M, output columns K, reduction over N; B operand = Wt [K, N]
  if (act_dtype == FERVIT_F32)
    return gemm_f32_simt((const float*)dy, N, 1, (const float*)Wt, N, 1, M, K, N, 1, e, S_(stream));
  FV_CHECK(act_dtype == FERVIT_BF16, "linear_dgrad: unknown dtype %d", act_dtype);
  return gemm_bf16_tc((const bf16*)dy, N, false, (const bf16*)Wt, N, false, M, K, N, 1, force_bn, e, S_(stream));
}

static int op_wgrad_splits(int act_dtype, int M, int N, int K) {
  // same policy as the plan (plan.cu: wgrad_splits), restated on the public shapes: dW[N,K] reduced over M rows
  const int sms = num_sms();
  if (act_dtype == FERVIT_BF16) {
    const int s2 = gemm_wgrad2_splits(N, K, M);   // shapes the CTA-pair kernel takes: one 256 x 256 unit per pair
    if (s2 > 0) return s2;
    const int tiles = ceil_div(N, 128) * ceil_div(K, 128);
    int s = sms / (tiles > 0 ? tiles : 1);
    const int max_s = ceil_div(M, 256);
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    return gemm_bf16_tc_effective_splits(M, s);
  }
  const int tiles = ceil_div(N, 64) * ceil_div(K, 64);
  int s = (2 * sms) / (tiles > 0 ? tiles : 1);
  const int max_s = ceil_div(M, 128);
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return gemm_f32_simt_effective_splits(M, s);
}

FV_API long long fervit_linear_wgrad_scratch_floats(int M, int N, int K) {
  const int a = op_wgrad_splits(FERVIT_BF16, M, N, K), b = op_wgrad_splits(FERVIT_F32, M, N, K);
  return (long long)(a > b ? a : b) * N * K;
}

FV_API int fervit_linear_wgrad(int act_dtype, const void* dY, const void* X, int M, int N, int K, float alpha,
                               float* dW, float* scratch, void* stream) {
  FV_CHECK(dY && X && dW && scratch, "linear_wgrad: null argument");
  const int splits = op_wgrad_splits(act_dtype, M, N, K);
  Epilogue e = make_epilogue();
  e.ldo = K;
  if (splits > 1) e.out_f32 = scratch;
  else { e.out_f32 = dW; e.alpha = alpha; }
  if (act_dtype == FERVIT_F32) {
    FV_TRY(gemm_f32_simt((const float*)dY, 1, N, (const float*)X, 1, K, N, K, M, splits, e, S_(stream)));
  } else {
    FV_CHECK(act_dtype == FERVIT_BF16, "linear_wgrad: unknown dtype %d", act_dtype);
    FV_TRY(gemm_bf16_tc((const bf16*)dY, N, true, (const bf16*)X, K, true, N, K, M, splits, 0, e, S_(stream)));
  }
  if (splits > 1) FV_TRY(splitk_reduce(scratch, splits, (size_t)N * K, nullptr, alpha, dW, S_(stream)));
  return 0;
}

FV_API long long fervit_linear_wgrad_bias_scratch_floats(int M, int N, int K) {
  return fervit_linear_wgrad_scratch_floats(M, N, K) + 256ll * N;
}

FV_API int fervit_linear_wgrad_bias(int act_dtype, const void* dY, const void* X, int M, int N, int K, float alpha,
                                    float* dW, float* db, float* scratch, void* stream) {
  FV_CHECK(dY && X && dW && db && scratch, "linear_wgrad_bias: null argument");
  if (act_dtype == FERVIT_BF16 && gemm_wgrad2_supported(N, K, M, N, K)) {
    // one launch: the CTA-pair kernel adds up the dY tiles it stages for the weight gradient
    const int splits = op_wgrad_splits(act_dtype, M, N, K);
    const int per = ceil_div(ceil_div(M, 64), splits);
    float* cs = scratch + (size_t)splits * N * K;
    FV_TRY(gemm_wgrad2((const bf16*)dY, N, (const bf16*)X, K, N, K, M, splits, per, splits > 1 ? scratch : dW, nullptr,
                       splits > 1 ? 1.0f : alpha, S_(stream), cs));
    if (splits > 1) FV_TRY(splitk_reduce(scratch, splits, (size_t)N * K, nullptr, alpha, dW, S_(stream)));
    return colsum_reduce_partials(cs, 2 * splits, N, db, S_(stream));
  }
  FV_TRY(fervit_linear_wgrad(act_dtype, dY, X, M, N, K, alpha, dW, scratch, stream));
  const Dropout nd = make_dropout(0.f, 0, 0);
  if (act_dtype == FERVIT_F32) return colsum<float>((const float*)dY, M, N, N, scratch, nullptr, 1.0f, db, nd, S_(stream));
  return colsum<bf16>((const bf16*)dY, M, N, N, scratch, nullptr, 1.0f, db, nd, S_(stream));
}

FV_API int fervit_layernorm_forward(int act_dtype, const float* x, const float* gamma, const float* beta, float eps,
                                    int rows, int E, float* y_f32, void* y_act, float* mean, float* rstd,
                                    void* stream) {
  FV_CHECK(x && gamma && beta, "layernorm_forward: null argument");
  if (act_dtype == FERVIT_F32)
    return layernorm_fwd<float>(x, gamma, beta, eps, rows, E, y_f32, (float*)y_act, mean, rstd, S_(stream));
  return layernorm_fwd<bf16>(x, gamma, beta, eps, rows, E, y_f32, (bf16*)y_act, mean, rstd, S_(stream));
}

FV_API long long fervit_layernorm_scratch_floats(int rows, int E) {
  return ((long long)layernorm_bwd_grid(rows) + 1) * 2 * E;
}

FV_API int fervit_layernorm_backward(int act_dtype, const void* dy, const float* x, const float* mean,
                                     const float* rstd, const float* gamma, const float* dres, int rows, int E,
                                     float* dx_f32, void* dx_act, float* scratch, float* dgamma, float* dbeta,
                                     void* stream) {
  FV_CHECK(dy && x && mean && rstd && gamma, "layernorm_backward: null argument");
  const bool wg = dgamma != nullptr;
  if (wg) FV_CHECK(dbeta && scratch, "layernorm_backward: dbeta and scratch are required with dgamma");
  const Dropout nd = make_dropout(0.f, 0, 0);
  if (act_dtype == FERVIT_F32) {
    FV_TRY((layernorm_bwd<float, float>((const float*)dy, x, mean, rstd, gamma, dres, rows, E, dx_f32, (float*)dx_act,
                                        wg ? scratch : nullptr, nd, S_(stream))));
  } else {
    FV_TRY((layernorm_bwd<bf16, bf16>((const bf16*)dy, x, mean, rstd, gamma, dres, rows, E, dx_f32, (bf16*)dx_act,
                                      wg ? scratch : nullptr, nd, S_(stream))));
  }
  if (wg) {
    const int g = layernorm_bwd_grid(rows);
    FV_TRY(colsum_reduce_partials(scratch, g, 2 * E, dgamma, S_(stream), dbeta, E));
  }
  return 0;
}

FV_API int fervit_attention_forward(int act_dtype, const void* qkv, int B, int S, int H, int hd, float dropout_p,
                                    unsigned long long seed, unsigned int site, void* out, float* lse,
                                    void* stream) {
  FV_CHECK(qkv && out, "attention_forward: null argument");
  const Dropout d = make_dropout(dropout_p, seed, site);
  if (act_dtype == FERVIT_F32)
    return attention_fwd<float>((const float*)qkv, (float*)out, lse, B, S, H, hd, d, S_(stream));
  return attention_fwd<bf16>((const bf16*)qkv, (bf16*)out, lse, B, S, H, hd, d, S_(stream));
}

FV_API int fervit_attention_backward(int act_dtype, const void* qkv, const void* out, const void* dout,
                                     const float* lse, int B, int S, int H, int hd, float dropout_p,
                                     unsigned long long seed, unsigned int site, void* dqkv, void* stream) {
  FV_CHECK(qkv && out && dout && lse && dqkv, "attention_backward: null argument");
  const Dropout d = make_dropout(dropout_p, seed, site);
  if (act_dtype == FERVIT_F32)
    return attention_bwd<float>((const float*)qkv, (const float*)out, (const float*)dout, lse, (float*)dqkv, B, S, H,
                                hd, d, S_(stream));
  return attention_bwd<bf16>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, B, S, H, hd, d,
                             S_(stream));
}

FV_API int fervit_dropout_mask(float* out, long long n, float p, unsigned long long seed, unsigned int site,
                               void* stream) {
  FV_CHECK(out && n >= 0, "dropout_mask: bad argument");
  return dropout_mask(out, (size_t)n, make_dropout(p, seed, site), S_(stream));
}

FV_API int fervit_cast_bf16(const float* src, void* dst, long long n, void* stream) {
  FV_CHECK(src && dst && n >= 0, "cast_bf16: bad argument");
  return cast_to_act<bf16>(src, (bf16*)dst, (size_t)n, S_(stream));
}
