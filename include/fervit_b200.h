/* fervit_b200 — C ABI of the B200-native (sm_100a) LatentViT / HybridLatentViT / ImageViT train-step path.
 *
 * The reference (yuki-ominato/FER-ViT) has no FFI of its own: its boundary for this path is the Python
 * nn.Module surface. This header is what a binding for that surface talks to; the Python
 * host layer in fer_vit_b200/ (ctypes) is such a binding. Every entry point
 *   - takes raw device pointers, explicit sizes and a cudaStream_t (as void*), never a torch type;
 *   - returns 0 on success, non-zero on failure (message via fervit_last_error()), never throws;
 *   - launches asynchronously on the given stream and is CUDA-graph capturable (no host sync, no allocation).
 * There is no CPU fallback: without a CUDA device the compute entry points fail.
 *
 * Each declaration cites the reference interface it replaces (file:line under the reference repo).
 */
#ifndef FERVIT_B200_H_
#define FERVIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FERVIT_ABI_VERSION 1

/* ----------------------------------------------------------------------------------------------
 * Library
 * -------------------------------------------------------------------------------------------- */
int fervit_abi_version(void);
const char* fervit_last_error(void);
/* number of kernels this library has launched since load (bench.py: gpu_launches) */
unsigned long long fervit_launch_count(void);

/* Per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline leg; never on inside a
 * timed region or a graph capture). Classes: 0 CTA-pair tcgen05 GEMM (work = FLOPs), 1 attention (algorithmic
 * bytes), 2 LayerNorm (algorithmic bytes), 3 fp32 CUDA-core GEMM (FLOPs), 4 single-CTA tcgen05 GEMM (weight gradients,
 * token projection; FLOPs), 5 fused AdapterModule kernel (FLOPs of its two contractions). enable(1) clears earlier
 * records and starts recording; enable(2) records nothing but keeps the backward pass on one stream (no side branch),
 * the launch order the per-launch records belong to; enable(0) switches both off. read() sums the records of one
 * class: total device milliseconds, total work, launch count. */
int fervit_profile_enable(int on);
int fervit_profile_read(int kernel_class, double* ms, double* work, long long* launches);

/* dtypes of activation buffers */
#define FERVIT_F32 0  /* fp32 mode: CUDA-core fp32 GEMMs, the 1e-4 parity mode                      */
#define FERVIT_BF16 1 /* bf16 mode: tcgen05/TMEM GEMMs fed by TMA, fp32 accumulate, fp32 residual   */

#define FERVIT_ACT_NONE 0
#define FERVIT_ACT_RELU 1 /* nn.TransformerEncoderLayer default (latent_vit.py:24-30)                */
#define FERVIT_ACT_GELU 2 /* exact-erf GELU: timm Mlp, AdapterModule, image_vit.py:106              */

/* ----------------------------------------------------------------------------------------------
 * Whole-model plan: the forward / backward of one model instance.
 * Replaces, for the three model classes,
 *   LatentViT.forward            models_fer_vit/latent_vit.py:38-48
 *   LatentViTv2.forward          models_fer_vit/latent_vit_v2.py:75-85   (SPE -> LWN -> LEAM pre-modules)
 *   HybridLatentViT.forward      models_fer_vit/hybrid_latent_vit.py:205-239 (+ AdapterModule :249-265,
 *                                timm Block semantics SURVEY.md §8a row a9)
 *   ImageViT.forward             models_fer_vit/image_vit.py:138-166 (+ PatchEmbedding :34-44)
 * and the autograd backward of each (loss.backward(), train_hybrid_latent_vit.py:134).
 * -------------------------------------------------------------------------------------------- */
typedef struct fervit_plan fervit_plan;

typedef struct fervit_config {
  int mode;          /* FERVIT_F32 | FERVIT_BF16 */
  int input_kind;    /* 0: latent tokens x[B,L,Din] (Linear); 1: image x[B,img_c,img_h,img_w] (Conv2d k=s=patch) */
  int L;             /* tokens per sample without the cls token (18 w+ layers, 196 patches) */
  int Din;           /* latent_dim, or img_c*patch*patch */
  int E;             /* embed dim */
  int depth;         /* transformer blocks */
  int H;             /* heads; head dim E/H must be 32, 48 or 64 */
  int F;             /* MLP hidden dim */
  int C;             /* classes (<= 16) */
  int norm_first;    /* 1: pre-norm timm Block; 0: post-norm nn.TransformerEncoderLayer */
  int act;           /* FERVIT_ACT_RELU | FERVIT_ACT_GELU */
  float eps_block;   /* 1e-6 timm, 1e-5 torch */
  float eps_head;    /* 1e-5 */
  int adapter_dim;   /* 0: none; else bottleneck width of AdapterModule after every block */
  float dropout;     /* p of the torch layers' four dropout sites (and ImageViT input dropout); 0 for timm */
  float head_dropout;/* Hybrid head nn.Dropout(0.1), hybrid_latent_vit.py:112 */
  int input_dropout; /* 1: apply `dropout` to cls+tokens+pos (image_vit.py:156) */
  int img_c, img_h, img_w, patch;
  int use_spe, use_lwn, use_lwn_res, use_leam; /* LatentViTv2 pre-modules */
  float eps_lwn;     /* 1e-5 */
} fervit_config;

/* Parameter slots. params[] / grads[] arrays are indexed FERVIT_G_* for globals and
 * FERVIT_NUM_GLOBAL + block * FERVIT_NUM_BLOCK + FERVIT_B_* for block parameters.
 * All parameters are fp32, contiguous, in the reference's own layouts (nn.Linear weight = [out, in]). */
enum {
  FERVIT_G_IN_W = 0,     /* input_proj.weight [E,Din] | patch_embed.proj.weight [E, img_c*patch*patch] */
  FERVIT_G_IN_B,         /* [E] */
  FERVIT_G_CLS,          /* cls_token [E] */
  FERVIT_G_POS,          /* pos_emb / pos_embed [(L+1),E] */
  FERVIT_G_HEAD_LN_W,    /* mlp_head.0 | head.0 | norm  weight [E] */
  FERVIT_G_HEAD_LN_B,
  FERVIT_G_HEAD_W,       /* mlp_head.1 | head.2 | head  weight [C,E] */
  FERVIT_G_HEAD_B,       /* [C] */
  FERVIT_G_SPE_GROUP,    /* spe.group_embed.weight [3,Din] */
  FERVIT_G_SPE_LAYER,    /* spe.layer_embed.weight [L,Din] */
  FERVIT_G_LWN_GAMMA,    /* stack of lwn.norms.{l}.weight [L,Din] */
  FERVIT_G_LWN_BETA,     /* stack of lwn.norms.{l}.bias   [L,Din] */
  FERVIT_G_LWN_GATE,     /* lwn.gate [L] */
  FERVIT_G_LEAM_W,       /* leam.layer_weights [L] */
  FERVIT_G_SPE_GROUPS,   /* spe.groups int64 [L] (buffer, never a gradient) */
  FERVIT_NUM_GLOBAL = 16
};
enum {
  FERVIT_B_LN1_W = 0, FERVIT_B_LN1_B,
  FERVIT_B_QKV_W,  /* attn.qkv.weight | self_attn.in_proj_weight [3E,E], rows [Q;K;V] */
  FERVIT_B_QKV_B,
  FERVIT_B_PROJ_W, /* attn.proj.weight | self_attn.out_proj.weight [E,E] */
  FERVIT_B_PROJ_B,
  FERVIT_B_LN2_W, FERVIT_B_LN2_B,
  FERVIT_B_FC1_W,  /* mlp.fc1.weight | linear1.weight [F,E] */
  FERVIT_B_FC1_B,
  FERVIT_B_FC2_W,  /* mlp.fc2.weight | linear2.weight [E,F] */
  FERVIT_B_FC2_B,
  FERVIT_B_AD1_W,  /* adapters.{i}.adapter.0.weight [A,E] */
  FERVIT_B_AD1_B,
  FERVIT_B_AD2_W,  /* adapters.{i}.adapter.2.weight [E,A] */
  FERVIT_B_AD2_B,
  FERVIT_B_ALPHA,  /* adapters.{i}.alpha [1] */
  FERVIT_NUM_BLOCK = 17
};

int fervit_plan_create(const fervit_config* cfg, fervit_plan** out);
void fervit_plan_destroy(fervit_plan* plan);
int fervit_plan_num_slots(const fervit_plan* plan);
/* element count of a slot's parameter (0 when the configuration does not use it) */
long long fervit_plan_slot_numel(const fervit_plan* plan, int slot);

/* params[n]: device pointers of the fp32 parameters (NULL for unused slots). Pointers are borrowed: PyTorch
 * owns the nn.Parameters (SURVEY.md §8b "Ownership"). */
int fervit_plan_set_params(fervit_plan* plan, const void* const* params, int n);

/* bf16 weight cache (W and W^T of every GEMM weight), a derived non-persistent copy that the caller owns and
 * refreshes after the optimizer changed the fp32 masters. slots == NULL refreshes every cached weight. */
/* LayerNorm folding (bf16 mode, pre-norm blocks, E a multiple of 128): for every block whose norm1 (norm2) AND the
 * qkv (fc1) weight and bias behind it are frozen, the norm is not run as a kernel. The weight cache then holds
 * W diag(gamma) for that slot (+ b' = b + W beta and the column sums cs, fp32); the GEMM that produces the norm's input
 * writes a bf16 copy centred on the row's previous mean plus per-128-column {sum, sum of squares}; the qkv / fc1 GEMM
 * applies y = rstd (acc - (mu - mref) cs) + b' in its epilogue and saves mu / rstd for the backward pass, whose
 * LayerNorm kernel then runs with gamma = 1 (the dgrad GEMM already multiplied by gamma). fold_norm1[0] must be 0
 * (block 0's norm1 reads the token projection). Call before fervit_plan_refresh_wcache; changing a flag makes that
 * slot's cache entry stale. Reference: timm Block.norm1 / norm2 called at hybrid_latent_vit.py:228. */
int fervit_plan_set_ln_fold(fervit_plan* plan, const int* fold_norm1, const int* fold_norm2, int depth);
long long fervit_plan_wcache_bytes(const fervit_plan* plan);
int fervit_plan_set_wcache(fervit_plan* plan, void* ptr, long long bytes);
int fervit_plan_refresh_wcache(fervit_plan* plan, const int* slots, int n, void* stream);

/* Activation workspace for batch B. save_for_backward = 0 gives the (smaller) inference workspace. */
long long fervit_plan_workspace_bytes(const fervit_plan* plan, int B, int save_for_backward);

/* logits[B,C] = model(x). training != 0 enables dropout; save_for_backward != 0 keeps the activations backward
 * needs in `ws`. Dropout is counter-based: keep(seed + *seed_dev, site, element) is a pure function, so backward
 * recomputes the masks from the same (seed, seed_dev). seed_dev (device scalar, may be NULL) lets a captured CUDA
 * graph draw fresh masks on every replay by bumping one device counter. */
int fervit_plan_forward(fervit_plan* plan, const float* x, int B, void* ws, long long ws_bytes, int training,
                        int save_for_backward, unsigned long long seed, const unsigned long long* seed_dev,
                        float* logits, void* stream);

/* Backward from dlogits[B,C]. grads[n]: where each parameter's gradient is WRITTEN (not accumulated); NULL = not
 * needed (frozen). The pass is cut into stages so a data-parallel host can all-reduce finished gradient buckets
 * while earlier blocks are still running: stage 0 = head, stage 1+k = block depth-1-k, stage depth+1 = input
 * projection / cls / pos / pre-modules. Run stages [stage_begin, stage_end) in increasing order. */
int fervit_plan_num_stages(const fervit_plan* plan);
int fervit_plan_backward(fervit_plan* plan, const float* x, int B, void* ws, long long ws_bytes, int training,
                         unsigned long long seed, const unsigned long long* seed_dev, const float* dlogits,
                         float* const* grads, int n, int stage_begin, int stage_end, void* stream);

/* Where a forward pass that saved for backward left one of its per-block activations inside the caller's workspace
 * `ws` (the same B as that forward). Parity tests read the activation SELECTOR of the MLP from here: with ReLU the
 * saved derivative is the 0/1 active set the kernel actually used, which the oracle is then given (like a dropout
 * mask), so that the continuous arithmetic is compared on the same selector. Buffers are in the plan's activation
 * dtype (bf16 mode: bf16; fp32 mode: float). */
enum { FERVIT_SAVED_ACT_DERIV = 0,  /* act'(fc1 pre-activation)  [B*S, F] */
       FERVIT_SAVED_ACT_OUT = 1,    /* act(fc1 pre-activation) after dropout [B*S, F] */
       FERVIT_SAVED_QKV = 2 };      /* [B*S, 3E] */
int fervit_plan_saved_buffer(const fervit_plan* plan, void* ws, int B, int block, int which, void** ptr,
                             long long* numel);

/* ----------------------------------------------------------------------------------------------
 * Loss: nn.CrossEntropyLoss(weight, label_smoothing), mean reduction
 * (train_hybrid_latent_vit.py:236-241, train_latent_vit.py:248-253). loss[1]; dlogits[B,C] = grad_scale * dloss/dlogits
 * (may be NULL). den_in (device scalar, may be NULL) overrides sum_i w[y_i], e.g. the global-batch value under
 * data parallelism. den_out may be NULL.
 * -------------------------------------------------------------------------------------------- */
int fervit_cross_entropy(const float* logits, const long long* labels, const float* weight, float label_smoothing,
                         int B, int C, const float* den_in, float grad_scale, float* loss, float* dlogits,
                         float* den_out, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Mixup loss of the LatentViT trainers (train_latent_vit.py:131, train_latent_vit_v2.py the same expression):
 *   loss = lam * CE(z, labels) + (1 - lam) * CE(z, labels[mix_index]),  CE as fervit_cross_entropy (each term divides
 * by its own sum_i w[label_i]). One launch; dlogits may be NULL. lam_dev (device scalar, may be NULL) overrides lam,
 * so a captured CUDA graph can draw a new lam per replay.
 * -------------------------------------------------------------------------------------------- */
int fervit_cross_entropy_mixup(const float* logits, const long long* labels, const long long* mix_index,
                               const float* weight, float label_smoothing, int B, int C, float lam,
                               const float* lam_dev, float grad_scale, float* loss, float* dlogits, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Latent batch producer: gather -> LatentAugment -> mixup in one launch over a packed latent table in HBM.
 * Replaces, for latents already on the device, LatentFERDataset.__getitem__ + LatentAugment.__call__
 * (data/latent_dataset.py:93-116, 28-49), the DataLoader's stacking and the mixup blend
 * `lam * latents + (1 - lam) * latents[index]` (train_latent_vit.py:119-127).
 *   a[b]   = keep(b) * scale(b) * (latents[sample_idx[b]] + noise_std * normal(b))   (order as the reference)
 *   out[b] = lam * a[b] + (1 - lam) * a[mix_index[b]]                               (mix_index NULL: out = a)
 * latents [n_rows, row_elems] fp32; sample_idx [B] (NULL: rows 0..B-1); labels [n_rows] -> labels_out [B] (both may
 * be NULL); aug NULL = no augmentation. status (device int, may be NULL) is set to 1 if an index is out of range (the
 * row is then read from row 0 / not mixed). out must not alias latents.
 * Draws are counter-based, a pure function of (s = seed + *seed_dev, batch position b, element e), so a test can
 * replay them on the host. With h64(site, i) = the 64-bit mix of csrc/common.cuh:mix_hash64(s, site, i),
 * u(x) = (x + 0.5) / 2^32 and g = b * row_elems + e:
 *   normal pair (g even, g + 1): z = h64(0x4C410000, g / 2), r = sqrt(-2 ln u(z >> 32)),
 *                                t = 2 pi (u(z & 0xffffffff) - 1/2), normal(g) = r cos t, normal(g + 1) = r sin t
 *   scale(b) = scale_min + (scale_max - scale_min) * u(h64(0x4C410002, b) >> 32)
 *   keep(g)  = ((h64(0x4C410003, g / 4) >> 16 (g mod 4)) & 0xffff) >= max(1, floor(mask_prob * 2^16))
 * lam is a double because the reference's `(1 - lam)` is evaluated in double before the tensor multiply rounds it.
 * -------------------------------------------------------------------------------------------- */
typedef struct fervit_latent_augment {
  float noise_std;            /* 0 disables */
  int use_scale;              /* scale_range is not None */
  float scale_min, scale_max; /* one uniform factor per sample */
  float mask_prob;            /* 0 disables; element kept with probability 1 - mask_prob */
} fervit_latent_augment;
int fervit_latent_batch(const float* latents, const long long* labels, long long n_rows, const long long* sample_idx,
                        int B, long long row_elems, const fervit_latent_augment* aug, unsigned long long seed,
                        const unsigned long long* seed_dev, const long long* mix_index, double lam,
                        const float* lam_dev, float* out, long long* labels_out, int* status, void* stream);

/* ----------------------------------------------------------------------------------------------
 * LatentDecomposer (models_fer_vit/latent_decomposer.py:82-173), the front-end of ExpressionAwareViT
 * (expression_aware_vit.py:109-122): coefficients on C unit-norm expression directions, expression part, identity
 * part and the requested combination in one launch.
 *   coef[b,c] = <w_plus[b], directions[c]>  over the flattened row;  scores [B,C] (may be NULL) receives them
 *   w_expr = sum_c coef n_c (decompose_mode 0 'all_classes') or the single |coef|-largest term (1 'max_class')
 *   out = w_expr (output_mode 0 'expr_only') | w - w_expr (1 'id_only') | w_id + enhance_alpha * w_expr (2 'enhanced')
 *       | [w_expr ; w_id] as [B, 2, row_elems] (3 'concat').  out may be NULL when only scores are wanted.
 * w_plus [B,row_elems], directions [C,row_elems] fp32, C <= 8, row_elems % 4 == 0 and <= 12800. No backward: the
 * input carries no gradient and the directions are buffers.
 * -------------------------------------------------------------------------------------------- */
int fervit_latent_decompose(const float* w_plus, const float* directions, int B, int C, long long row_elems,
                            int decompose_mode, int output_mode, float enhance_alpha, float* out, float* scores,
                            void* stream);

/* ----------------------------------------------------------------------------------------------
 * Stand-alone pre-modules (modules/leam.py:31-40, modules/layer_wise_norm.py:35-50,
 * modules/semantic_pe.py:36-48), fused; any subset via the use_* flags. fp32 in / fp32 out.
 * scratch for backward: fervit_premodules_scratch_floats(B, L, D) floats.
 * -------------------------------------------------------------------------------------------- */
typedef struct fervit_premodules {
  int use_spe, use_lwn, use_lwn_res, use_leam;
  const float* group_embed;   /* [3,D]  */
  const float* layer_embed;   /* [L,D]  */
  const long long* groups;    /* [L]    */
  const float* gamma;         /* [L,D]  */
  const float* beta;          /* [L,D]  */
  const float* gate;          /* [L]    */
  const float* leam_w;        /* [L]    */
  float eps;
} fervit_premodules;
int fervit_premodules_forward(const fervit_premodules* p, const float* x, int B, int L, int D, float* y, void* stream);
long long fervit_premodules_scratch_floats(int B, int L, int D);
int fervit_premodules_backward(const fervit_premodules* p, const float* x, const float* dy, int B, int L, int D,
                               float* dx, float* scratch, float* dgamma, float* dbeta, float* dlayer_embed,
                               float* dgroup_embed, float* dgate, float* dleam, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Operator-level entry points (parity tests and micro-benchmarks call these; the plan uses the same kernels).
 * act_dtype: FERVIT_F32 or FERVIT_BF16 for the buffers typed void*.
 * -------------------------------------------------------------------------------------------- */

/* y = act(x W^T + b) (+ residual): nn.Linear / F.linear. fp32 mode: x, W fp32; bf16 mode: x, W bf16 ([N,K]).
 * out (act_dtype) and out_f32 may each be NULL; residual fp32 [M,N] may be NULL; pre (act_dtype, pre-activation) may
 * be NULL; act | 0x100 makes `pre` receive act'(pre-activation) instead (what the plan saves for the backward GEMM).
 * force_bn (ignored in fp32 mode): 0 = auto; 128 / 256 = N tile of the CTA-pair tcgen05 kernel; -64 / -128 /
 * -256 = force the single-CTA tcgen05 kernel with that N tile (benchmarks and tests). */
int fervit_linear_forward(int act_dtype, const void* x, const void* W, const float* bias, const float* residual,
                          int M, int N, int K, int act, void* out, float* out_f32, void* pre, int force_bn,
                          void* stream);
/* dx[M,K] = (dy[M,N] Wt^T) * act'(aux) (+ residual): the input gradient of nn.Linear fused with the derivative of the
 * activation that preceded it (autograd of F.linear + nn.GELU / nn.ReLU). Wt is the TRANSPOSED weight [K,N] (act_dtype);
 * aux (act_dtype, [M,K], pre-activation) is required when act != 0; residual fp32 [M,K] may be NULL; out / out_f32 may
 * each be NULL. force_bn as above. */
int fervit_linear_dgrad(int act_dtype, const void* dy, const void* Wt, const void* aux, const float* residual, int M,
                        int N, int K, int act, void* out, float* out_f32, int force_bn, void* stream);
/* dW[N,K] = alpha * dY^T X  (dY [M,N], X [M,K], act_dtype): the weight gradient of nn.Linear.
 * scratch: fervit_linear_wgrad_scratch_floats(M,N,K) floats. */
long long fervit_linear_wgrad_scratch_floats(int M, int N, int K);
int fervit_linear_wgrad(int act_dtype, const void* dY, const void* X, int M, int N, int K, float alpha, float* dW,
                        float* scratch, void* stream);
/* The same plus the bias gradient db[N] = colsum(dY) (nn.Linear's bias.grad). bf16 shapes with N, K >= 256 and
 * M >= 2048 rows take ONE launch of the CTA-pair kernel (csrc/gemm_wgrad2.cu), which adds up the dY tiles it stages for
 * the weight gradient; everything else is fervit_linear_wgrad followed by a column-sum pass.
 * scratch: fervit_linear_wgrad_bias_scratch_floats(M,N,K) floats. */
long long fervit_linear_wgrad_bias_scratch_floats(int M, int N, int K);
int fervit_linear_wgrad_bias(int act_dtype, const void* dY, const void* X, int M, int N, int K, float alpha, float* dW,
                             float* db, float* scratch, void* stream);

/* nn.LayerNorm over rows of fp32 x[rows,E]; y in act_dtype and/or fp32; mean/rstd [rows] saved for backward. */
int fervit_layernorm_forward(int act_dtype, const float* x, const float* gamma, const float* beta, float eps, int rows,
                             int E, float* y_f32, void* y_act, float* mean, float* rstd, void* stream);
/* dx = LN'(dy) (+ dres). dy is act_dtype. dgamma/dbeta NULL = not needed.
 * scratch: fervit_layernorm_scratch_floats(rows,E) floats when dgamma is requested. */
long long fervit_layernorm_scratch_floats(int rows, int E);
int fervit_layernorm_backward(int act_dtype, const void* dy, const float* x, const float* mean, const float* rstd,
                              const float* gamma, const float* dres, int rows, int E, float* dx_f32, void* dx_act,
                              float* scratch, float* dgamma, float* dbeta, void* stream);

/* softmax(q k^T / sqrt(hd)) v per (sample, head): F.scaled_dot_product_attention / nn.MultiheadAttention core.
 * qkv [B*S, 3*H*hd] columns [Q|K|V]; out [B*S, H*hd]; lse [B*H*S] (may be NULL in forward-only use). */
int fervit_attention_forward(int act_dtype, const void* qkv, int B, int S, int H, int hd, float dropout_p,
                             unsigned long long seed, unsigned int site, void* out, float* lse, void* stream);
int fervit_attention_backward(int act_dtype, const void* qkv, const void* out, const void* dout, const float* lse,
                              int B, int S, int H, int hd, float dropout_p, unsigned long long seed,
                              unsigned int site, void* dqkv, void* stream);

/* keep-mask scaled by 1/(1-p) of dropout site `site` for n elements (tests feed it to the oracle) */
int fervit_dropout_mask(float* out, long long n, float p, unsigned long long seed, unsigned int site, void* stream);
/* site ids used by the plan: block b has sites 8*b + {0: attention weights, 1: after out-proj, 2: after the MLP
 * activation, 3: after fc2}; input dropout and head dropout use the two ids below. */
#define FERVIT_SITE_INPUT 0xFFFF0u
#define FERVIT_SITE_HEAD 0xFFFF1u

/* AdapterModule (hybrid_latent_vit.py:249-265) fused on the tensor cores, bf16 mode, bottleneck 64, E % 256 == 0:
 *   forward : y = x + alpha * (GELU(x W1^T + b1) W2^T + b2); saves g = GELU(u) and d = GELU'(u), both [T,64] bf16.
 *             x_bf16 / x_f32: the same input as bf16 (MMA operand) and fp32 (residual); W1 [64,E], W2 [E,64] bf16.
 *   backward: du = alpha * (dy W2) * d, dx = dy + du W1. W2t [64,E] and W1t [E,64] are the TRANSPOSED weights (bf16);
 *             dx_bf16 may be NULL. Weight gradients: fervit_linear_wgrad on (dy, g) and (du, x). */
int fervit_adapter_forward(const void* x_bf16, const float* x_f32, const void* W1, const float* b1, const void* W2,
                           const float* b2, const float* alpha, int T, int E, void* g, void* d, float* y, void* stream);
int fervit_adapter_backward_input(const void* dy_bf16, const float* dy_f32, const void* W2t, const void* W1t,
                                  const float* alpha, const void* d, int T, int E, void* du, float* dx, void* dx_bf16,
                                  void* stream);

/* Fused AdamW over n fp32 tensors (device pointer tables live in HOST memory; they are passed to the kernels by
 * value): torch.optim.AdamW semantics (decoupled weight decay, bias correction, amsgrad off) with per-tensor
 * hyper-parameter groups, the step the reference trainers run right after backward
 * (train_hybrid_latent_vit.py:63-117, 244-248; clip_grad_norm_ at train_latent_vit_v2.py:132-133).
 *   hyper: device [groups][5] = lr, beta1, beta2, eps, weight_decay; group[i] selects the row of tensor i
 *   step : device float, the number of updates done so far (incremented here: capturable in a CUDA graph)
 *   max_norm > 0: gradients are scaled by min(1, max_norm / (||g||_2 + 1e-6)) over ALL n tensors first (and written
 *   back scaled, as clip_grad_norm_ leaves them); scratch: fervit_adamw_scratch_floats(n, numel) floats, the clip
 *   coefficient and the total norm end up in its last two used floats. */
long long fervit_adamw_scratch_floats(int n, const long long* numel);
int fervit_adamw_step(int n, void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                      const long long* numel, const int* group, const float* hyper, float* step, float max_norm,
                      float* scratch, void* stream);

/* Stream-K scratch for the stand-alone GEMM entry points (fervit_linear_forward / fervit_linear_dgrad): when a GEMM's
 * last round of 256 x 256 tiles under-fills the GPU (e.g. 57 tiles on 74 SM pairs) and K >= 1024, the tiles of that
 * round are split along K across all pairs and the fp32 partial accumulators travel through this buffer. ptr: device
 * buffer of fervit_gemm_scratch_bytes() bytes, 256-byte aligned, whose first 4096 bytes are zero, alive until replaced;
 * NULL switches stream-K off for these entry points. Plans carry their own region inside the weight cache. */
long long fervit_gemm_scratch_bytes(void);
int fervit_set_gemm_scratch(void* ptr, long long bytes);

/* Diagnostics: with FERVIT_GEMM_DEBUG bit 8 set, the CTA-pair GEMM records the wall time (ns, %globaltimer) and the SM
 * cycle count (clock64) of its CTA 0; cycles / ns = the SM clock in GHz while the kernel ran. */
int fervit_debug_gemm_clock(double* ns, double* cycles);

/* In-kernel launch timer of the CTA-pair tcgen05 GEMM (bench.py's roofline leg). While on, every launch takes the next
 * slot and its CTAs record {first start once the grid dependency has resolved, last exit} in %globaltimer ns; a CUDA
 * graph captured while it is on keeps its slots, so the launches of one replay are timed where they run (PDL edges and
 * parallel branches intact — host-side events cannot sit there). op 1: on, restart slot numbering; op 2: clear the
 * stamps recorded so far (async on `stream`, enqueue before the replay to be timed); op 0: off.
 * read: sum over the slots stamped since the last clear — device microseconds, FLOPs (2MNK), launches; per_launch
 * (optional, cap records of 8 doubles, in launch order): {us, flops, M, N, K, epilogue kind + 16 * fp32 variant, start, end
 * in microseconds after the first launch's start}. */
int fervit_gemm_prof(int op, void* stream);
int fervit_gemm_prof_read(double* us, double* flops, long long* launches, double* per_launch, int cap);

/* Diagnostics: with FERVIT_GEMM_DEBUG bit 64 set, the CTA-pair GEMM stamps clock64 at every phase boundary of two CTAs
 * (row 0: CTA 0, row 1: leader of the last pair); out receives 2 x 64 values (slot meanings: csrc/gemm_tc2.cu TL_*). */
int fervit_debug_gemm_timeline(unsigned long long* out, int n);
/* same for the fused AdapterModule kernel (CTA 0, 32 values) */
int fervit_debug_adapter_timeline(unsigned long long* out, int n);

/* fp32 -> bf16 cast (n multiple of 4) */
int fervit_cast_bf16(const float* src, void* dst, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FERVIT_B200_H_ */
