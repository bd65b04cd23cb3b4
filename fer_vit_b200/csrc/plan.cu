// Whole-model plan: sequences the kernels of one LatentViT / LatentViTv2 / HybridLatentViT / ImageViT forward and
// backward on one stream. This is the native runtime behind the Python nn.Module boundary: it owns no memory
// (parameters, weight cache and workspace are caller-provided device buffers) and never synchronises.
//
// Reference semantics implemented here:
//   input stage   latent_vit.py:40-44, hybrid_latent_vit.py:215-222, image_vit.py:148-156 (+ latent_vit_v2.py:82-84)
//   pre-norm blk  timm Block (x += attn(norm1(x)); x += mlp(norm2(x))), hybrid_latent_vit.py:227-233
//   adapter       AdapterModule.forward, hybrid_latent_vit.py:264-265
//   post-norm blk nn.TransformerEncoderLayer training path (x = norm1(x + sa(x)); x = norm2(x + ff(x)))
//   head          latent_vit.py:46-47, hybrid_latent_vit.py:236-237, image_vit.py:161-164
#include "common.cuh"
#include "kernels.h"
#include "fervit_b200.h"
#include <type_traits>
#include <vector>
#include <stdlib.h>

namespace fervit {

namespace {

struct Arena {
  char* base;
  size_t off;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct BlockBufs {
  float *m1, *r1, *m2, *r2;  // LayerNorm statistics [T]
  void* xn1;    // pre-norm: norm1(x) [T,E] act
  void* qkv;    // [T,3E] act
  float* lse;   // [B*H*S]
  void* ao;     // attention output [T,E] act
  float* x_mid; // pre-norm: x after the attention residual; post-norm: x + sa(x) before norm1   [T,E] fp32
  void* xn2;    // pre-norm: norm2(x_mid); post-norm: norm1 output (MLP input)                   [T,E] act
  void* u1;     // act'(fc1 pre-activation) [T,F] act: saved by the forward epilogue, multiplied in by fc2's dgrad
  void* g1;     // fc1 activation (after dropout) [T,F] act
  float* rr2;   // post-norm: x1 + ff(x1) before norm2 [T,E] fp32
  void* x2_at;  // adapter input [T,E] act
  void* ua;     // act'(adapter pre-activation) [T,A] act
  void* ga;     // adapter activation [T,A] act
};

struct Bufs {
  void* a_in;                  // [B*L, Din] act: A operand of the token projection
  std::vector<float*> x;       // depth+1 residual-stream tensors [T,E] fp32
  std::vector<void*> x_at;     // post-norm: act copies of x[i]
  std::vector<BlockBufs> blk;
  float *head_mean, *head_rstd;
  float* tmp_f32;              // [T,E]
  float* ln_part[2];           // folded LayerNorm: per-row, per-128-column {sum, sum of squares}; [0] norm1, [1] norm2
  // backward transients
  float* dx[2];
  void* dx_at[2];
  void* d_big;                 // [T, max(3E,F)] act
  void* d_e1;                  // [T,E] act
  void* d_e2;                  // [T,E] act
  void* du_ad;                 // [T,A] act
  void* dtok;                  // [B*L,E] act
  void* dain;                  // [B*L,Din] act
  // per-block partial sums of the AdapterModule gradients, finished once per backward stage group
  struct AdParts { float *w2_part, *w1_part, *cs_dy, *cs_du; int s2, s1, chunks; };
  std::vector<AdParts> ad;
  // deferred form (ad_deferred()): the adapters' dy / du / g / x of ALL blocks sit in contiguous pools [depth][T][.], so
  // that one split-K launch per weight and one column-sum launch per bias serve every block of a backward call
  bool ad_defer = false;
  void *ad_dy_pool = nullptr, *ad_du_pool = nullptr;     // [depth][T][E], [depth][T][A] act
  float *ad_w2_pool = nullptr, *ad_w1_pool = nullptr;    // [ad_slabs][E*A] split-K slabs
  float *ad_csdy_pool = nullptr, *ad_csdu_pool = nullptr;  // [depth][AD_CPB][E], [depth][AD_CPB][A]
  int ad_slabs = 0;
  float* ad_fin;               // scratch of adapter_grad_finalize
  float* head_scratch;         // partial sums of the head's parameter gradients (own buffer: may run on the side stream)
  float* scratch;
  size_t scratch_floats;
};

}  // namespace

}  // namespace fervit

using namespace fervit;

struct fervit_plan {
  fervit_config cfg;
  int S, T_per_sample, HD, A;
  bool has_pre;
  std::vector<const void*> params;
  std::vector<size_t> wb_off, wbt_off;  // byte offsets into the weight cache per slot (SIZE_MAX = not cached)
  size_t wcache_bytes;
  char* wcache;
  size_t sk_off = SIZE_MAX, sk_bytes = 0;  // stream-K scratch of the CTA-pair GEMM, inside the weight cache buffer
  // LayerNorm folded into the GEMM that follows it (frozen norm + frozen weight, bf16 mode): per block,
  // norm1 -> qkv and norm2 -> fc1. The weight cache then holds W diag(gamma) for those slots, plus b' and cs.
  std::vector<int> fold1, fold2;
  std::vector<size_t> fb_off, fcs_off;     // per slot: byte offsets of b' [out] and cs [out] (fp32) in the weight cache
  const float* FB(int slot) const { return reinterpret_cast<const float*>(wcache + fb_off[slot]); }
  const float* FCS(int slot) const { return reinterpret_cast<const float*>(wcache + fcs_off[slot]); }
  int bwd_cur;  // which of dx[0]/dx[1] holds the running gradient between backward stages
  // Side stream of the backward pass: the adapter weight-gradient GEMMs and column sums do not feed the dgrad chain,
  // so they run beside it and fill the SMs the chain leaves idle (57-tile GEMMs, partial last waves). Fork/join by
  // events; inside a CUDA-graph capture this becomes a parallel branch of the graph.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_tail = nullptr;
  ~fervit_plan() {
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_tail) cudaEventDestroy(ev_tail);
    if (side) cudaStreamDestroy(side);
  }

  int nslots() const { return FERVIT_NUM_GLOBAL + cfg.depth * FERVIT_NUM_BLOCK; }
  static int bslot(int blk, int s) { return FERVIT_NUM_GLOBAL + blk * FERVIT_NUM_BLOCK + s; }
  const float* P(int slot) const { return reinterpret_cast<const float*>(params[slot]); }
  const float* PB(int blk, int s) const { return P(bslot(blk, s)); }
  const bf16* WB(int slot) const { return reinterpret_cast<const bf16*>(wcache + wb_off[slot]); }
  const bf16* WBT(int slot) const { return reinterpret_cast<const bf16*>(wcache + wbt_off[slot]); }
};

namespace fervit {
namespace {

// rows/cols of a GEMM weight slot ([out, in]); false if the slot is not a GEMM weight
bool weight_shape(const fervit_plan* p, int slot, int* R, int* C) {
  const fervit_config& c = p->cfg;
  if (slot == FERVIT_G_IN_W) { *R = c.E; *C = c.Din; return true; }
  if (slot < FERVIT_NUM_GLOBAL) return false;
  const int s = (slot - FERVIT_NUM_GLOBAL) % FERVIT_NUM_BLOCK;
  switch (s) {
    case FERVIT_B_QKV_W: *R = 3 * c.E; *C = c.E; return true;
    case FERVIT_B_PROJ_W: *R = c.E; *C = c.E; return true;
    case FERVIT_B_FC1_W: *R = c.F; *C = c.E; return true;
    case FERVIT_B_FC2_W: *R = c.E; *C = c.F; return true;
    case FERVIT_B_AD1_W: if (c.adapter_dim) { *R = c.adapter_dim; *C = c.E; return true; } return false;
    case FERVIT_B_AD2_W: if (c.adapter_dim) { *R = c.E; *C = c.adapter_dim; return true; } return false;
    default: return false;
  }
}

long long slot_numel(const fervit_plan* p, int slot) {
  const fervit_config& c = p->cfg;
  const long long E = c.E, L = c.L, D = c.Din, A = c.adapter_dim;
  if (slot < FERVIT_NUM_GLOBAL) {
    switch (slot) {
      case FERVIT_G_IN_W: return E * D;
      case FERVIT_G_IN_B: return E;
      case FERVIT_G_CLS: return E;
      case FERVIT_G_POS: return (L + 1) * E;
      case FERVIT_G_HEAD_LN_W: case FERVIT_G_HEAD_LN_B: return E;
      case FERVIT_G_HEAD_W: return (long long)c.C * E;
      case FERVIT_G_HEAD_B: return c.C;
      case FERVIT_G_SPE_GROUP: return c.use_spe ? 3 * D : 0;
      case FERVIT_G_SPE_LAYER: return c.use_spe ? L * D : 0;
      case FERVIT_G_LWN_GAMMA: case FERVIT_G_LWN_BETA: return c.use_lwn ? L * D : 0;
      case FERVIT_G_LWN_GATE: return (c.use_lwn && c.use_lwn_res) ? L : 0;
      case FERVIT_G_LEAM_W: return c.use_leam ? L : 0;
      case FERVIT_G_SPE_GROUPS: return c.use_spe ? L : 0;
      default: return 0;
    }
  }
  const int s = (slot - FERVIT_NUM_GLOBAL) % FERVIT_NUM_BLOCK;
  switch (s) {
    case FERVIT_B_LN1_W: case FERVIT_B_LN1_B: case FERVIT_B_LN2_W: case FERVIT_B_LN2_B: return E;
    case FERVIT_B_QKV_W: return 3 * E * E;
    case FERVIT_B_QKV_B: return 3 * E;
    case FERVIT_B_PROJ_W: return E * E;
    case FERVIT_B_PROJ_B: return E;
    case FERVIT_B_FC1_W: return (long long)c.F * E;
    case FERVIT_B_FC1_B: return c.F;
    case FERVIT_B_FC2_W: return (long long)c.F * E;
    case FERVIT_B_FC2_B: return E;
    case FERVIT_B_AD1_W: case FERVIT_B_AD2_W: return A * E;
    case FERVIT_B_AD1_B: return A;
    case FERVIT_B_AD2_B: return A ? E : 0;
    case FERVIT_B_ALPHA: return A ? 1 : 0;
    default: return 0;
  }
}

// split-K factor for a weight gradient dW[Nout,Kin] reduced over T rows
int wgrad_splits(bool bf16_mode, int Nout, int Kin, int T) {
  const int sms = num_sms();
  if (bf16_mode) {
    const int s2 = gemm_wgrad2_splits(Nout, Kin, T);   // shapes the CTA-pair kernel takes (gemm_wgrad2.cu)
    if (s2 > 0) return s2;
    const int tiles = ceil_div(Nout, 128) * ceil_div(Kin, 128);
    int s = sms / (tiles > 0 ? tiles : 1);
    const int max_s = ceil_div(T, 256);  // at least 4 k-blocks of 64 per split
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    return gemm_bf16_tc_effective_splits(T, s);
  }
  const int tiles = ceil_div(Nout, 64) * ceil_div(Kin, 64);
  int s = (2 * sms) / (tiles > 0 ? tiles : 1);
  const int max_s = ceil_div(T, 128);
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return gemm_f32_simt_effective_splits(T, s);
}

// Deferred AdapterModule gradients (bf16, pre-norm, fused adapter kernel, T a multiple of the GEMM's k-block): instead
// of four small launches per block beside the main chain, every block's dy and du are kept and ONE split-K GEMM per
// weight (K = blocks x T, each split inside one block) plus ONE column-sum launch per bias finish all blocks of the
// backward call. FERVIT_ADAPTER_DEFER=0 restores the per-block form.
constexpr int AD_CPB = 16;   // column-sum chunks per block
bool ad_deferred(const fervit_plan* p, int B, bool save) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FERVIT_ADAPTER_DEFER"); on = (e && atoi(e) == 0) ? 0 : 1; }
  const fervit_config& c = p->cfg;
  const long long T = (long long)B * p->S;
  return on == 1 && save && c.mode == FERVIT_BF16 && c.norm_first && c.adapter_dim > 0 && T % 64 == 0 &&
         T % AD_CPB == 0 && adapter_fused_supported((int)T, c.E, c.adapter_dim);
}
// split-K slabs per block for n blocks: the divisor d of T / 64 (at least 4 k-blocks per split) that minimises
// waves x (k-blocks per split + ~8 k-blocks' worth of ramp and epilogue) for tiles x n x d work units
int ad_defer_spb(int n, int T, int tiles) {
  const int kb = T / 64, sms = num_sms();
  int best = 1;
  long long best_cost = -1;
  for (int d = 1; d <= kb; ++d) {
    if (kb % d || (kb / d < 4 && d > 1)) continue;
    const long long units = (long long)tiles * n * d;
    const long long cost = ((units + sms - 1) / sms) * (kb / d + 8);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = d; }
  }
  return best;
}
int ad_defer_slabs(const fervit_plan* p) {
  const int a = 4 * num_sms() / 6 + 1;   // more than ad_defer_spb() ever picks for 6 tiles; the caller clamps anyway
  return a > p->cfg.depth ? a : p->cfg.depth;
}

size_t scratch_floats_for(const fervit_plan* p, int B) {
  const fervit_config& c = p->cfg;
  const bool bf = c.mode == FERVIT_BF16;
  const int T = B * p->S, Tl = B * c.L;
  size_t m = 1024;
  auto upd = [&](size_t v) { if (v > m) m = v; };
  auto wg = [&](int Nout, int Kin, int rows) { upd((size_t)wgrad_splits(bf, Nout, Kin, rows) * ((size_t)Nout * Kin + 2 * Nout)); };
  wg(c.E, c.Din, Tl);
  wg(3 * c.E, c.E, T); wg(c.E, c.E, T); wg(c.F, c.E, T); wg(c.E, c.F, T);
  if (c.adapter_dim) { wg(c.adapter_dim, c.E, T); wg(c.E, c.adapter_dim, T); }
  const int widest = 3 * c.E > c.F ? 3 * c.E : c.F;
  upd((size_t)colsum_chunks(T) * widest);
  upd((size_t)colsum_chunks(B) * p->S * c.E);
  upd((size_t)colsum_chunks(Tl) * c.E);
  upd(((size_t)layernorm_bwd_grid(T) + 1) * 2 * c.E);
  upd((size_t)head_wgrad_chunks(B) * (c.C + 2) * c.E);
  if (p->has_pre) upd((size_t)premodules_chunks(B) * c.L * (3 * c.Din + 2));
  return m;
}

template <typename AT>
void carve(const fervit_plan* p, int B, bool save, Arena& ar, Bufs& b) {
  const fervit_config& c = p->cfg;
  const bool bf = c.mode == FERVIT_BF16;
  const size_t T = (size_t)B * p->S, Tl = (size_t)B * c.L, E = c.E, F = c.F, A = c.adapter_dim;
  const bool post = !c.norm_first;
  const bool need_ain_copy = bf || c.input_kind == 1 || p->has_pre;
  b.a_in = need_ain_copy ? ar.take<AT>(Tl * c.Din) : nullptr;
  const int nx = save ? c.depth + 1 : 2;
  b.x.assign(c.depth + 1, nullptr);
  b.x_at.assign(c.depth + 1, nullptr);
  std::vector<float*> xs(nx);
  std::vector<void*> xas(nx, nullptr);
  for (int i = 0; i < nx; ++i) {
    xs[i] = ar.take<float>(T * E);
    if (post && bf) xas[i] = ar.take<AT>(T * E);
  }
  for (int i = 0; i <= c.depth; ++i) {
    b.x[i] = xs[i % nx];
    b.x_at[i] = (post && bf) ? xas[i % nx] : (void*)b.x[i];
  }
  b.blk.resize(c.depth);
  b.ad_defer = ad_deferred(p, B, save);
  AT *x2_pool = nullptr, *ga_pool = nullptr;
  if (b.ad_defer) {
    x2_pool = ar.take<AT>((size_t)c.depth * T * E);
    ga_pool = ar.take<AT>((size_t)c.depth * T * A);
  }
  BlockBufs shared{};
  for (int i = 0; i < c.depth; ++i) {
    BlockBufs k{};
    if (save || i == 0) {
      k.m1 = ar.take<float>(T); k.r1 = ar.take<float>(T); k.m2 = ar.take<float>(T); k.r2 = ar.take<float>(T);
      k.xn1 = post ? nullptr : (void*)ar.take<AT>(T * E);
      k.qkv = ar.take<AT>(T * 3 * E);
      k.lse = ar.take<float>((size_t)B * c.H * p->S);
      k.ao = ar.take<AT>(T * E);
      k.x_mid = ar.take<float>(T * E);
      k.xn2 = ar.take<AT>(T * E);
      k.u1 = ar.take<AT>(T * F);
      k.g1 = ar.take<AT>(T * F);
      k.rr2 = post ? ar.take<float>(T * E) : nullptr;
      if (A) {
        k.x2_at = b.ad_defer ? (void*)(x2_pool + (size_t)i * T * E) : (void*)ar.take<AT>(T * E);
        k.ua = ar.take<AT>(T * A);
        k.ga = b.ad_defer ? (void*)(ga_pool + (size_t)i * T * A) : (void*)ar.take<AT>(T * A);
      }
      shared = k;
    } else {
      k = shared;
    }
    b.blk[i] = k;
  }
  b.head_mean = ar.take<float>(B);
  b.head_rstd = ar.take<float>(B);
  b.tmp_f32 = ar.take<float>(T * E);
  b.ln_part[0] = b.ln_part[1] = nullptr;
  if (bf && !post && E % 128 == 0) {
    b.ln_part[0] = ar.take<float>(T * (E / 128) * 2);
    b.ln_part[1] = ar.take<float>(T * (E / 128) * 2);
  }
  if (save) {
    for (int i = 0; i < 2; ++i) {
      b.dx[i] = ar.take<float>(T * E);
      b.dx_at[i] = bf ? (void*)ar.take<AT>(T * E) : (void*)b.dx[i];
    }
    b.d_big = ar.take<AT>(T * (3 * E > F ? 3 * E : F));
    b.d_e1 = ar.take<AT>(T * E);
    b.d_e2 = ar.take<AT>(T * E);
    if (A && b.ad_defer) {
      b.du_ad = nullptr;
      b.ad_dy_pool = ar.take<AT>((size_t)c.depth * T * E);
      b.ad_du_pool = ar.take<AT>((size_t)c.depth * T * A);
      b.ad_slabs = ad_defer_slabs(p);
      b.ad_w2_pool = ar.take<float>((size_t)b.ad_slabs * E * A);
      b.ad_w1_pool = ar.take<float>((size_t)b.ad_slabs * E * A);
      b.ad_csdy_pool = ar.take<float>((size_t)c.depth * AD_CPB * E);
      b.ad_csdu_pool = ar.take<float>((size_t)c.depth * AD_CPB * A);
      b.ad_fin = ar.take<float>((size_t)adapter_grad_finalize_scratch_floats());
    } else if (A) {
      b.du_ad = ar.take<AT>(T * A);
      b.ad.resize(c.depth);
      for (int i = 0; i < c.depth; ++i) {
        Bufs::AdParts& q = b.ad[i];
        q.s2 = wgrad_splits(bf, (int)E, (int)A, (int)T);
        q.s1 = wgrad_splits(bf, (int)A, (int)E, (int)T);
        q.chunks = colsum_chunks((int)T);
        q.w2_part = ar.take<float>((size_t)q.s2 * E * A);
        q.w1_part = ar.take<float>((size_t)q.s1 * A * E);
        q.cs_dy = ar.take<float>((size_t)q.chunks * E);
        q.cs_du = ar.take<float>((size_t)q.chunks * A);
      }
      b.ad_fin = ar.take<float>((size_t)adapter_grad_finalize_scratch_floats());
    }
    b.dtok = ar.take<AT>(Tl * E);
    b.dain = p->has_pre ? (void*)ar.take<AT>(Tl * c.Din) : nullptr;
    b.head_scratch = ar.take<float>((size_t)head_wgrad_chunks(B) * (c.C + 2) * E);
    b.scratch_floats = scratch_floats_for(p, B);
    b.scratch = ar.take<float>(b.scratch_floats);
  }
}

struct Ctx {
  const fervit_plan* p;
  cudaStream_t st;
  bool training;
  const unsigned long long* seed_dev;
  uint64_t seed;
  Dropout site(int blk, int k) const {
    const float pr = training ? p->cfg.dropout : 0.f;
    return make_dropout(pr, seed, (uint32_t)(blk * 8 + k), seed_dev);
  }
  Dropout none() const { return make_dropout(0.f, 0, 0); }
};

// Y = A W^T (forward) or dX = dY W (transposed) with the shared epilogue; W is parameter slot `slot` ([Nout, Kin]).
template <typename AT>
int linear(const Ctx& c, const AT* A, int M, int slot, bool transposed, const Epilogue& epi) {
  int Nout = 0, Kin = 0;
  if (!weight_shape(c.p, slot, &Nout, &Kin)) { set_error("linear: slot %d is not a GEMM weight", slot); return 1; }
  if constexpr (std::is_same<AT, float>::value) {
    const float* W = c.p->P(slot);
    if (!transposed) return gemm_f32_simt(A, Kin, 1, W, Kin, 1, M, Nout, Kin, 1, epi, c.st);
    return gemm_f32_simt(A, Nout, 1, W, 1, Kin, M, Kin, Nout, 1, epi, c.st);
  } else {
    Epilogue e2 = epi;
    static int sk_on = -1;   // opt-in: measured slower at batch 256 (gemm_tc2.cu: Params), and it gives up the
    if (sk_on < 0) {         // batch-invariance of the results (the summation order then depends on the tile count)
      const char* s = getenv("FERVIT_GEMM_STREAMK");
      sk_on = (s && atoi(s) == 1) ? 1 : 0;
    }
    if (sk_on && c.p->sk_bytes && c.st != c.p->side) {   // main-stream GEMMs run one after another: one region serves all
      e2.sk_ws = c.p->wcache + c.p->sk_off;
      e2.sk_bytes = c.p->sk_bytes;
    }
    if (!transposed) return gemm_bf16_tc(A, Kin, false, c.p->WB(slot), Kin, false, M, Nout, Kin, 1, 0, e2, c.st);
    return gemm_bf16_tc(A, Nout, false, c.p->WBT(slot), Nout, false, M, Kin, Nout, 1, 0, e2, c.st);
  }
}

// dW[Nout,Kin] = alpha * dY^T X, reduced over T rows (tokens); deterministic split-K.
// db (optional, bf16 mode): the bias gradient colsum(dY) from the same launch when the CTA-pair kernel takes the shape
// (gemm_wgrad2.cu adds the dY tiles up while they sit in shared memory); *did_bias tells the caller whether it did.
template <typename AT>
int wgrad(const Ctx& c, const AT* dY, int Nout, const AT* X, int Kin, int T, const float* alpha_ptr, float* dW,
          float* scratch, float* db = nullptr, bool* did_bias = nullptr) {
  const bool bf = !std::is_same<AT, float>::value;
  const int splits = wgrad_splits(bf, Nout, Kin, T);
  if (did_bias) *did_bias = false;
  if constexpr (!std::is_same<AT, float>::value) {
    if (db && did_bias && gemm_wgrad2_supported(Nout, Kin, T, Nout, Kin)) {
      const int total_kb = ceil_div(T, 64), per = ceil_div(total_kb, splits);
      float* slabs = scratch;
      float* cs = scratch + (size_t)splits * Nout * Kin;
      FV_TRY(gemm_wgrad2(dY, Nout, X, Kin, Nout, Kin, T, splits, per, splits > 1 ? slabs : dW,
                         splits > 1 ? nullptr : alpha_ptr, 1.0f, c.st, cs));
      if (splits > 1) FV_TRY(splitk_reduce(slabs, splits, (size_t)Nout * Kin, alpha_ptr, 1.0f, dW, c.st));
      FV_TRY(colsum_reduce_partials(cs, 2 * splits, Nout, db, c.st));
      *did_bias = true;
      return 0;
    }
  }
  Epilogue e = make_epilogue();
  e.ldo = Kin;
  if (splits > 1) {
    e.out_f32 = scratch;
  } else {
    e.out_f32 = dW;
    e.alpha_ptr = alpha_ptr;
  }
  if constexpr (std::is_same<AT, float>::value) {
    FV_TRY(gemm_f32_simt(dY, 1, Nout, X, 1, Kin, Nout, Kin, T, splits, e, c.st));
  } else {
    FV_TRY(gemm_bf16_tc(dY, Nout, true, X, Kin, true, Nout, Kin, T, splits, 0, e, c.st));
  }
  if (splits > 1) FV_TRY(splitk_reduce(scratch, splits, (size_t)Nout * Kin, alpha_ptr, 1.0f, dW, c.st));
  return 0;
}

// split-K slabs of dY^T X only: partial[splits][Nout*Kin] (splits = wgrad_splits(...)); summed later in a fixed order
template <typename AT>
int wgrad_partial(const Ctx& c, const AT* dY, int Nout, const AT* X, int Kin, int T, int splits, float* partial) {
  Epilogue e = make_epilogue();
  e.ldo = Kin;
  e.out_f32 = partial;
  if constexpr (std::is_same<AT, float>::value) {
    return gemm_f32_simt(dY, 1, Nout, X, 1, Kin, Nout, Kin, T, splits, e, c.st);
  } else {
    return gemm_bf16_tc(dY, Nout, true, X, Kin, true, Nout, Kin, T, splits, 0, e, c.st);
  }
}

int ensure_side_stream(fervit_plan* p) {
  if (!p->side) {
    FV_CUDA(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
    FV_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    FV_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    FV_CUDA(cudaEventCreateWithFlags(&p->ev_tail, cudaEventDisableTiming));
  }
  return 0;
}

bool use_side_stream() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FERVIT_SIDE_STREAM"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on == 1 && !prof_serial();   // the per-kernel event profile wants one serial stream
}

PreParams pre_params(const fervit_plan* p) {
  PreParams q;
  const fervit_config& c = p->cfg;
  q.use_spe = c.use_spe; q.use_lwn = c.use_lwn; q.use_res = c.use_lwn_res; q.use_leam = c.use_leam;
  q.group_embed = p->P(FERVIT_G_SPE_GROUP);
  q.layer_embed = p->P(FERVIT_G_SPE_LAYER);
  q.groups = reinterpret_cast<const long long*>(p->params[FERVIT_G_SPE_GROUPS]);
  q.gamma = p->P(FERVIT_G_LWN_GAMMA);
  q.beta = p->P(FERVIT_G_LWN_BETA);
  q.gate = p->P(FERVIT_G_LWN_GATE);
  q.leam_w = p->P(FERVIT_G_LEAM_W);
  q.eps = c.eps_lwn;
  return q;
}

template <typename AT>
int forward_impl(fervit_plan* p, const float* x, int B, void* ws, long long ws_bytes, bool training, bool save,
                 uint64_t seed, const unsigned long long* seed_dev, float* logits, cudaStream_t st) {
  const fervit_config& c = p->cfg;
  constexpr bool F32 = std::is_same<AT, float>::value;
  Arena ar{reinterpret_cast<char*>(ws), 0};
  Bufs b;
  carve<AT>(p, B, save, ar, b);
  FV_CHECK((long long)ar.off <= ws_bytes, "forward: workspace too small (%lld < %zu bytes)", ws_bytes, ar.off);
  Ctx cx{p, st, training, seed_dev, seed};
  const int S = p->S, T = B * S, Tl = B * c.L, E = c.E, F = c.F, A = c.adapter_dim;
  const bool post = !c.norm_first;

  // ---------------- input stage ----------------
  const AT* a_in;
  if (c.input_kind == 1) {
    FV_TRY(im2col<AT>(x, (AT*)b.a_in, B, c.img_c, c.img_h, c.img_w, c.patch, st));
    a_in = (const AT*)b.a_in;
  } else if (p->has_pre) {
    FV_TRY(premodules_fwd<AT>(x, B, c.L, c.Din, pre_params(p), F32 ? (float*)b.a_in : nullptr,
                              F32 ? nullptr : (AT*)b.a_in, st));
    a_in = (const AT*)b.a_in;
  } else if constexpr (F32) {
    a_in = x;
  } else {
    FV_TRY(cast_to_act<AT>(x, (AT*)b.a_in, (size_t)Tl * c.Din, st));
    a_in = (const AT*)b.a_in;
  }
  {
    Epilogue e = make_epilogue();
    e.bias = p->P(FERVIT_G_IN_B);
    e.remap_L = c.L;
    e.pos = p->P(FERVIT_G_POS);
    e.ldo = E;
    e.out_f32 = b.x[0];
    if (post && !F32) e.out = b.x_at[0];
    FV_TRY(linear<AT>(cx, a_in, Tl, FERVIT_G_IN_W, false, e));
    const Dropout din = make_dropout((training && c.input_dropout) ? c.dropout : 0.f, seed, FERVIT_SITE_INPUT, seed_dev);
    FV_TRY(cls_rows<AT>(p->P(FERVIT_G_CLS), p->P(FERVIT_G_POS), b.x[0], (post && !F32) ? (AT*)b.x_at[0] : nullptr, B, S,
                        E, din, st));
    FV_TRY(token_dropout<AT>(b.x[0], (post && !F32) ? (AT*)b.x_at[0] : nullptr, B, S, E, din, st));
  }

  // ---------------- blocks ----------------
  for (int i = 0; i < c.depth; ++i) {
    BlockBufs& k = b.blk[i];
    if (!post) {
      // Folded LayerNorms (frozen norm + frozen weight, bf16 mode): no norm kernel. The GEMM that PRODUCES the norm's
      // input also writes its bf16 copy centred on the row's previous mean and the per-part sums; the GEMM that
      // CONSUMES it applies rstd, the mean shift and the folded bias in its epilogue (common.cuh: Epilogue).
      const bool f1 = !F32 && p->fold1[i], f2 = !F32 && p->fold2[i];
      const bool f1_next = !F32 && i + 1 < c.depth && p->fold1[i + 1];
      const int qkv_slot = p->bslot(i, FERVIT_B_QKV_W), fc1_slot = p->bslot(i, FERVIT_B_FC1_W);
      if (!f1)
        FV_TRY(layernorm_fwd<AT>(b.x[i], p->PB(i, FERVIT_B_LN1_W), p->PB(i, FERVIT_B_LN1_B), c.eps_block, T, E, nullptr,
                                 (AT*)k.xn1, k.m1, k.r1, st));
      Epilogue e = make_epilogue();
      e.bias = p->PB(i, FERVIT_B_QKV_B); e.out = k.qkv; e.ldo = 3 * E;
      if (f1) {
        e.bias = p->FB(qkv_slot); e.ln_cs = p->FCS(qkv_slot); e.ln_part = b.ln_part[0]; e.ln_mref = b.blk[i - 1].m2;
        e.ln_mean = k.m1; e.ln_rstd = k.r1; e.ln_eps = c.eps_block; e.ln_parts = E / 128;
      }
      FV_TRY(linear<AT>(cx, (const AT*)k.xn1, T, qkv_slot, false, e));
      FV_TRY(attention_fwd<AT>((const AT*)k.qkv, (AT*)k.ao, k.lse, B, S, c.H, p->HD, cx.site(i, 0), st));
      e = make_epilogue();
      e.bias = p->PB(i, FERVIT_B_PROJ_B); e.residual = b.x[i]; e.out_f32 = k.x_mid; e.ldo = E; e.drop = cx.site(i, 1);
      if (f2) { e.out = k.xn2; e.lnp_part = b.ln_part[1]; e.lnp_mref = k.m1; }
      FV_TRY(linear<AT>(cx, (const AT*)k.ao, T, p->bslot(i, FERVIT_B_PROJ_W), false, e));
      if (!f2)
        FV_TRY(layernorm_fwd<AT>(k.x_mid, p->PB(i, FERVIT_B_LN2_W), p->PB(i, FERVIT_B_LN2_B), c.eps_block, T, E, nullptr,
                                 (AT*)k.xn2, k.m2, k.r2, st));
      e = make_epilogue();
      e.bias = p->PB(i, FERVIT_B_FC1_B); e.act = c.act; e.out_pre = k.u1; e.pre_is_deriv = 1; e.out = k.g1; e.ldo = F; e.drop = cx.site(i, 2);
      if (f2) {
        e.bias = p->FB(fc1_slot); e.ln_cs = p->FCS(fc1_slot); e.ln_part = b.ln_part[1]; e.ln_mref = k.m1;
        e.ln_mean = k.m2; e.ln_rstd = k.r2; e.ln_eps = c.eps_block; e.ln_parts = E / 128;
      }
      FV_TRY(linear<AT>(cx, (const AT*)k.xn2, T, fc1_slot, false, e));
      e = make_epilogue();
      e.bias = p->PB(i, FERVIT_B_FC2_B); e.residual = k.x_mid; e.ldo = E; e.drop = cx.site(i, 3);
      void* next_xn1 = f1_next ? b.blk[i + 1].xn1 : nullptr;   // where the next block's (folded) norm1 input goes
      if (A) {
        float* x2 = F32 ? (float*)k.x2_at : b.tmp_f32;
        e.out_f32 = x2;
        if (!F32) e.out = k.x2_at;
        FV_TRY(linear<AT>(cx, (const AT*)k.g1, T, p->bslot(i, FERVIT_B_FC2_W), false, e));
        bool fused = false;
        if constexpr (!F32) {
          if (adapter_fused_supported(T, E, A)) {
            // down-projection, GELU, up-projection, alpha and residual in one tensor-core kernel (adapter_tc.cu)
            FV_TRY(adapter_fused(0, (const bf16*)k.x2_at, p->WB(p->bslot(i, FERVIT_B_AD1_W)),
                                 p->WB(p->bslot(i, FERVIT_B_AD2_W)), x2, p->PB(i, FERVIT_B_AD1_B),
                                 p->PB(i, FERVIT_B_AD2_B), p->PB(i, FERVIT_B_ALPHA), nullptr, (bf16*)k.ga, (bf16*)k.ua,
                                 b.x[i + 1], (bf16*)next_xn1, T, E, st, f1_next ? b.ln_part[0] : nullptr,
                                 f1_next ? k.m2 : nullptr));
            fused = true;
          }
        }
        if (!fused) {
          Epilogue d = make_epilogue();
          d.bias = p->PB(i, FERVIT_B_AD1_B); d.act = ACT_GELU; d.out_pre = k.ua; d.pre_is_deriv = 1; d.out = k.ga; d.ldo = A;
          FV_TRY(linear<AT>(cx, (const AT*)k.x2_at, T, p->bslot(i, FERVIT_B_AD1_W), false, d));
          Epilogue u = make_epilogue();
          u.bias = p->PB(i, FERVIT_B_AD2_B); u.alpha_ptr = p->PB(i, FERVIT_B_ALPHA); u.residual = x2;
          u.out_f32 = b.x[i + 1]; u.ldo = E;
          if (f1_next) { u.out = next_xn1; u.lnp_part = b.ln_part[0]; u.lnp_mref = k.m2; }
          FV_TRY(linear<AT>(cx, (const AT*)k.ga, T, p->bslot(i, FERVIT_B_AD2_W), false, u));
        }
      } else {
        e.out_f32 = b.x[i + 1];
        if (f1_next) { e.out = next_xn1; e.lnp_part = b.ln_part[0]; e.lnp_mref = k.m2; }
        FV_TRY(linear<AT>(cx, (const AT*)k.g1, T, p->bslot(i, FERVIT_B_FC2_W), false, e));
      }
    } else {
      Epilogue e = make_epilogue();
      e.bias = p->PB(i, FERVIT_B_QKV_B); e.out = k.qkv; e.ldo = 3 * E;
      FV_TRY(linear<AT>(cx, (const AT*)b.x_at[i], T, p->bslot(i, FERVIT_B_QKV_W), false, e));
      FV_TRY(attention_fwd<AT>((const AT*)k.qkv, (AT*)k.ao, k.lse, B, S, c.H, p->HD, cx.site(i, 0), st));
      e = make_epilogue();
      e.bias = p->PB(i, FERVIT_B_PROJ_B); e.residual = b.x[i]; e.out_f32 = k.x_mid; e.ldo = E; e.drop = cx.site(i, 1);
      FV_TRY(linear<AT>(cx, (const AT*)k.ao, T, p->bslot(i, FERVIT_B_PROJ_W), false, e));
      float* x1 = F32 ? (float*)k.xn2 : b.tmp_f32;
      FV_TRY(layernorm_fwd<AT>(k.x_mid, p->PB(i, FERVIT_B_LN1_W), p->PB(i, FERVIT_B_LN1_B), c.eps_block, T, E, x1,
                               F32 ? nullptr : (AT*)k.xn2, k.m1, k.r1, st));
      e = make_epilogue();
      e.bias = p->PB(i, FERVIT_B_FC1_B); e.act = c.act; e.out_pre = k.u1; e.pre_is_deriv = 1; e.out = k.g1; e.ldo = F; e.drop = cx.site(i, 2);
      FV_TRY(linear<AT>(cx, (const AT*)k.xn2, T, p->bslot(i, FERVIT_B_FC1_W), false, e));
      e = make_epilogue();
      e.bias = p->PB(i, FERVIT_B_FC2_B); e.residual = x1; e.out_f32 = k.rr2; e.ldo = E; e.drop = cx.site(i, 3);
      FV_TRY(linear<AT>(cx, (const AT*)k.g1, T, p->bslot(i, FERVIT_B_FC2_W), false, e));
      FV_TRY(layernorm_fwd<AT>(k.rr2, p->PB(i, FERVIT_B_LN2_W), p->PB(i, FERVIT_B_LN2_B), c.eps_block, T, E, b.x[i + 1],
                               F32 ? nullptr : (AT*)b.x_at[i + 1], k.m2, k.r2, st));
    }
  }

  // ---------------- head ----------------
  const Dropout dh = make_dropout(training ? c.head_dropout : 0.f, seed, FERVIT_SITE_HEAD, seed_dev);
  FV_TRY(head_fwd(b.x[c.depth], B, S, E, p->P(FERVIT_G_HEAD_LN_W), p->P(FERVIT_G_HEAD_LN_B), c.eps_head,
                  p->P(FERVIT_G_HEAD_W), p->P(FERVIT_G_HEAD_B), c.C, dh, logits, b.head_mean, b.head_rstd, st));
  return 0;
}

template <typename AT>
int backward_impl(fervit_plan* p, const float* x, int B, void* ws, long long ws_bytes, bool training, uint64_t seed,
                  const unsigned long long* seed_dev,
                  const float* dlogits, float* const* G, int stage_begin, int stage_end, cudaStream_t st) {
  const fervit_config& c = p->cfg;
  constexpr bool F32 = std::is_same<AT, float>::value;
  Arena ar{reinterpret_cast<char*>(ws), 0};
  Bufs b;
  carve<AT>(p, B, true, ar, b);
  FV_CHECK((long long)ar.off <= ws_bytes, "backward: workspace too small (%lld < %zu bytes)", ws_bytes, ar.off);
  Ctx cx{p, st, training, seed_dev, seed};
  const int S = p->S, T = B * S, Tl = B * c.L, E = c.E, F = c.F, A = c.adapter_dim;
  const bool post = !c.norm_first;
  auto GB = [&](int blk, int s) -> float* { return G[p->bslot(blk, s)]; };
  const Dropout nodrop = cx.none();
  std::vector<AdapterGradJob> ad_jobs;  // adapters whose partial gradients were produced in this call
  // deferred form: the bf16 gradient ENTERING block i's stage (written by the head or by block i+1's norm1 backward) and
  // the adapter's du live in per-block pool slots; ad_lo..ad_hi = blocks of this call whose adapter gradients are due
  auto AD_DY = [&](int blk) { return (AT*)b.ad_dy_pool + (size_t)blk * T * E; };
  auto AD_DU = [&](int blk) { return (AT*)b.ad_du_pool + (size_t)blk * T * A; };
  int ad_lo = c.depth, ad_hi = -1;
  bool side_pending = false;            // side-stream work of the current block not yet joined
  bool side_used = false;               // anything at all went to the side stream in this call

  for (int stage = stage_begin; stage < stage_end; ++stage) {
    if (stage == 0) {
      // ---------------- head ----------------
      p->bwd_cur = 0;
      const Dropout dh = make_dropout(training ? c.head_dropout : 0.f, seed, FERVIT_SITE_HEAD, seed_dev);
      const int wg = G[FERVIT_G_HEAD_W] != nullptr;
      if (wg)
        FV_CHECK(G[FERVIT_G_HEAD_B] && G[FERVIT_G_HEAD_LN_W] && G[FERVIT_G_HEAD_LN_B],
                 "backward: head gradients must be requested together");
      cudaStream_t hs = st;
      if (wg && use_side_stream()) {   // head parameter gradients beside the dgrad chain; joined at the end of this call
        FV_TRY(ensure_side_stream(p));
        FV_CUDA(cudaEventRecord(p->ev_fork, st));
        FV_CUDA(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
        hs = p->side;
        side_used = true;
      }
      FV_TRY(head_bwd<AT>(b.x[c.depth], dlogits, B, S, E, p->P(FERVIT_G_HEAD_LN_W), p->P(FERVIT_G_HEAD_LN_B),
                          p->P(FERVIT_G_HEAD_W), c.C, b.head_mean, b.head_rstd, dh, b.dx[0],
                          F32 ? nullptr : (b.ad_defer ? AD_DY(c.depth - 1) : (AT*)b.dx_at[0]), wg, b.head_scratch,
                          G[FERVIT_G_HEAD_W],
                          G[FERVIT_G_HEAD_LN_W], G[FERVIT_G_HEAD_LN_B], G[FERVIT_G_HEAD_B], st, hs));
    } else if (stage <= c.depth) {
      const int i = c.depth - stage;
      BlockBufs& k = b.blk[i];
      int cur = p->bwd_cur;
      auto DX = [&](int w) { return b.dx[w]; };
      auto DXA = [&](int w) { return (AT*)b.dx_at[w]; };
      auto ATOUT = [&](int w) -> AT* { return F32 ? nullptr : (AT*)b.dx_at[w]; };
      if (!post) {
        // ---- adapter ----  (AdapterModule backward, hybrid_latent_vit.py:264-265)
        if (A) {
          bool fused = false;
          if constexpr (!F32) fused = adapter_fused_supported(T, E, A);
          Epilogue e = make_epilogue();
          if (fused) {
            // du = alpha * (dy W2) * gelu'(u) and dx = dy + du W1 in one tensor-core kernel (adapter_tc.cu)
            if constexpr (!F32)
              FV_TRY(adapter_fused(1, (const bf16*)(b.ad_defer ? AD_DY(i) : DXA(cur)),
                                   p->WBT(p->bslot(i, FERVIT_B_AD2_W)), p->WBT(p->bslot(i, FERVIT_B_AD1_W)), DX(cur),
                                   nullptr, nullptr, p->PB(i, FERVIT_B_ALPHA), (const bf16*)k.ua,
                                   (bf16*)(b.ad_defer ? AD_DU(i) : (AT*)b.du_ad), nullptr, DX(cur ^ 1),
                                   (bf16*)ATOUT(cur ^ 1), T, E, st));
          } else {
            // du = alpha * (dy W2) * gelu'(u): one GEMM, derivative and alpha in its epilogue
            e.act_bwd = ACT_DERIV; e.aux = k.ua; e.alpha_ptr = p->PB(i, FERVIT_B_ALPHA); e.out = b.du_ad; e.ldo = A;
            FV_TRY(linear<AT>(cx, DXA(cur), T, p->bslot(i, FERVIT_B_AD2_W), true, e));
          }
          if (GB(i, FERVIT_B_AD2_W)) {
            FV_CHECK(GB(i, FERVIT_B_AD2_B) && GB(i, FERVIT_B_AD1_W) && GB(i, FERVIT_B_AD1_B) && GB(i, FERVIT_B_ALPHA),
                     "backward: adapter gradients must be requested together");
            if (b.ad_defer) {   // nothing now: one grouped launch per weight / bias at the end of this call
              FV_CHECK(fused, "backward: deferred adapter gradients need the fused adapter kernel");
              if (i < ad_lo) ad_lo = i;
              if (i > ad_hi) ad_hi = i;
            } else {
            // partial sums only; adapter_grad_finalize() below finishes all blocks of this stage group at once.
            // They read dy (DX/DXA(cur)) and du, which the main chain overwrites only at norm2's backward: fork
            // here, join there.
            const Bufs::AdParts& q = b.ad[i];
            cudaStream_t ws = st;
            if (use_side_stream()) {
              FV_TRY(ensure_side_stream(p));
              FV_CUDA(cudaEventRecord(p->ev_fork, st));
              FV_CUDA(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
              ws = p->side;
              side_pending = true;
              side_used = true;
            }
            Ctx cw = cx;
            cw.st = ws;
            FV_TRY(wgrad_partial<AT>(cw, DXA(cur), E, (const AT*)k.ga, A, T, q.s2, q.w2_part));
            FV_TRY(colsum_partial<float>(DX(cur), T, E, E, q.cs_dy, ws));
            FV_TRY(wgrad_partial<AT>(cw, (const AT*)b.du_ad, A, (const AT*)k.x2_at, E, T, q.s1, q.w1_part));
            FV_TRY(colsum_partial<AT>((const AT*)b.du_ad, T, A, A, q.cs_du, ws));
            if (side_pending) FV_CUDA(cudaEventRecord(p->ev_join, p->side));
            AdapterGradJob job;
            job.w2_part = q.w2_part; job.w1_part = q.w1_part; job.cs_dy = q.cs_dy; job.cs_du = q.cs_du;
            job.W2 = p->PB(i, FERVIT_B_AD2_W); job.b2 = p->PB(i, FERVIT_B_AD2_B);
            job.alpha_ptr = p->PB(i, FERVIT_B_ALPHA);
            job.dW2 = GB(i, FERVIT_B_AD2_W); job.dW1 = GB(i, FERVIT_B_AD1_W); job.db2 = GB(i, FERVIT_B_AD2_B);
            job.db1 = GB(i, FERVIT_B_AD1_B); job.dalpha = GB(i, FERVIT_B_ALPHA);
            job.s2 = q.s2; job.s1 = q.s1; job.chunks = q.chunks; job.E = E; job.A = A;
            ad_jobs.push_back(job);
            }
          }
          if (!fused) {
            e = make_epilogue();
            e.residual = DX(cur); e.out_f32 = DX(cur ^ 1); e.out = ATOUT(cur ^ 1); e.ldo = E;
            FV_TRY(linear<AT>(cx, (const AT*)b.du_ad, T, p->bslot(i, FERVIT_B_AD1_W), true, e));
          }
          cur ^= 1;
        }
        // ---- MLP ----
        if (GB(i, FERVIT_B_FC2_W)) {
          // (the bias gradient comes out of the weight-gradient launch when the bf16 copy IS the gradient: no dropout)
          bool did = false;
          FV_TRY(wgrad<AT>(cx, DXA(cur), E, (const AT*)k.g1, F, T, nullptr, GB(i, FERVIT_B_FC2_W), b.scratch,
                           cx.site(i, 3).threshold ? nullptr : GB(i, FERVIT_B_FC2_B), &did));
          if (!did)
            FV_TRY(colsum<float>(DX(cur), T, E, E, b.scratch, nullptr, 1.0f, GB(i, FERVIT_B_FC2_B), cx.site(i, 3), st));
        }
        Epilogue e = make_epilogue();
        e.act_bwd = ACT_DERIV; e.aux = k.u1; e.out = b.d_big; e.ldo = F; e.drop = cx.site(i, 2);
        FV_TRY(linear<AT>(cx, DXA(cur), T, p->bslot(i, FERVIT_B_FC2_W), true, e));
        if (GB(i, FERVIT_B_FC1_W)) {
          bool did = false;
          FV_TRY(wgrad<AT>(cx, (const AT*)b.d_big, F, (const AT*)k.xn2, E, T, nullptr, GB(i, FERVIT_B_FC1_W), b.scratch,
                           GB(i, FERVIT_B_FC1_B), &did));
          if (!did)
            FV_TRY(colsum<AT>((const AT*)b.d_big, T, F, F, b.scratch, nullptr, 1.0f, GB(i, FERVIT_B_FC1_B), nodrop, st));
        }
        e = make_epilogue();
        e.out = b.d_e1; e.ldo = E;
        FV_TRY(linear<AT>(cx, (const AT*)b.d_big, T, p->bslot(i, FERVIT_B_FC1_W), true, e));
        if (side_pending) {  // norm2's backward overwrites the gradient buffers the side stream reads
          FV_CUDA(cudaStreamWaitEvent(st, p->ev_join, 0));
          side_pending = false;
        }
        {
          float* part = GB(i, FERVIT_B_LN2_W) ? b.scratch : nullptr;
          // a folded norm2: fc1's cached transpose is (W diag(gamma))^T, so d_e1 already is gamma * dh
          const float* g2 = (!F32 && p->fold2[i]) ? nullptr : p->PB(i, FERVIT_B_LN2_W);
          FV_TRY((layernorm_bwd<AT, AT>((const AT*)b.d_e1, k.x_mid, k.m2, k.r2, g2, DX(cur), T, E,
                                        DX(cur ^ 1), ATOUT(cur ^ 1), part, nodrop, st)));
          if (part) {
            const int g = layernorm_bwd_grid(T);
            // partial layout [g][2][E]: gamma rows then beta rows -> reduce as a [g, 2E] matrix into a temp pair
            FV_CHECK(GB(i, FERVIT_B_LN2_B), "backward: LayerNorm gradients must be requested together");
            FV_TRY(colsum_reduce_partials(part, g, 2 * E, GB(i, FERVIT_B_LN2_W), st, GB(i, FERVIT_B_LN2_B), E));
          }
          cur ^= 1;
        }
        // ---- attention ----
        if (GB(i, FERVIT_B_PROJ_W)) {
          bool did = false;
          FV_TRY(wgrad<AT>(cx, DXA(cur), E, (const AT*)k.ao, E, T, nullptr, GB(i, FERVIT_B_PROJ_W), b.scratch,
                           GB(i, FERVIT_B_PROJ_B), &did));
          if (!did)
            FV_TRY(colsum<float>(DX(cur), T, E, E, b.scratch, nullptr, 1.0f, GB(i, FERVIT_B_PROJ_B), nodrop, st));
        }
        e = make_epilogue();
        e.out = b.d_e2; e.ldo = E;
        FV_TRY(linear<AT>(cx, DXA(cur), T, p->bslot(i, FERVIT_B_PROJ_W), true, e));
        FV_TRY(attention_bwd<AT>((const AT*)k.qkv, (const AT*)k.ao, (const AT*)b.d_e2, k.lse, (AT*)b.d_big, B, S, c.H,
                                 p->HD, cx.site(i, 0), st));
        if (GB(i, FERVIT_B_QKV_W)) {
          bool did = false;
          FV_TRY(wgrad<AT>(cx, (const AT*)b.d_big, 3 * E, (const AT*)k.xn1, E, T, nullptr, GB(i, FERVIT_B_QKV_W),
                           b.scratch, GB(i, FERVIT_B_QKV_B), &did));
          if (!did)
            FV_TRY(colsum<AT>((const AT*)b.d_big, T, 3 * E, 3 * E, b.scratch, nullptr, 1.0f, GB(i, FERVIT_B_QKV_B), nodrop,
                              st));
        }
        e = make_epilogue();
        e.out = b.d_e1; e.ldo = E;
        FV_TRY(linear<AT>(cx, (const AT*)b.d_big, T, p->bslot(i, FERVIT_B_QKV_W), true, e));
        {
          float* part = GB(i, FERVIT_B_LN1_W) ? b.scratch : nullptr;
          const float* g1 = (!F32 && p->fold1[i]) ? nullptr : p->PB(i, FERVIT_B_LN1_W);
          // deferred adapter gradients: the act copy is the dy of block i-1's adapter and goes to that block's slot
          AT* at_out = (b.ad_defer && i > 0) ? AD_DY(i - 1) : ATOUT(cur ^ 1);
          FV_TRY((layernorm_bwd<AT, AT>((const AT*)b.d_e1, b.x[i], k.m1, k.r1, g1, DX(cur), T, E,
                                        DX(cur ^ 1), at_out, part, nodrop, st)));
          if (part) {
            const int g = layernorm_bwd_grid(T);
            FV_CHECK(GB(i, FERVIT_B_LN1_B), "backward: LayerNorm gradients must be requested together");
            FV_TRY(colsum_reduce_partials(part, g, 2 * E, GB(i, FERVIT_B_LN1_W), st, GB(i, FERVIT_B_LN1_B), E));
          }
          cur ^= 1;
        }
      } else {
        // ---------------- post-norm block ----------------
        // norm2 backward: dr2 (fp32, residual path) and its dropout-masked act copy (fc2 path)
        {
          float* part = GB(i, FERVIT_B_LN2_W) ? b.scratch : nullptr;
          FV_TRY((layernorm_bwd<float, AT>(DX(cur), k.rr2, k.m2, k.r2, p->PB(i, FERVIT_B_LN2_W), nullptr, T, E,
                                           DX(cur ^ 1), (AT*)b.d_e1, part, cx.site(i, 3), st)));
          if (part) {
            const int g = layernorm_bwd_grid(T);
            FV_CHECK(GB(i, FERVIT_B_LN2_B), "backward: LayerNorm gradients must be requested together");
            FV_TRY(colsum_reduce_partials(part, g, 2 * E, GB(i, FERVIT_B_LN2_W), st, GB(i, FERVIT_B_LN2_B), E));
          }
          cur ^= 1;
        }
        const AT* dy2 = (const AT*)b.d_e1;  // masked dr2
        if (GB(i, FERVIT_B_FC2_W)) {
          bool did = false;   // dy2 IS the masked gradient, so its column sums are the bias gradient
          FV_TRY(wgrad<AT>(cx, dy2, E, (const AT*)k.g1, F, T, nullptr, GB(i, FERVIT_B_FC2_W), b.scratch,
                           GB(i, FERVIT_B_FC2_B), &did));
          if (!did)
            FV_TRY(colsum<float>(DX(cur), T, E, E, b.scratch, nullptr, 1.0f, GB(i, FERVIT_B_FC2_B), cx.site(i, 3), st));
        }
        Epilogue e = make_epilogue();
        e.act_bwd = ACT_DERIV; e.aux = k.u1; e.out = b.d_big; e.ldo = F; e.drop = cx.site(i, 2);
        FV_TRY(linear<AT>(cx, dy2, T, p->bslot(i, FERVIT_B_FC2_W), true, e));
        if (GB(i, FERVIT_B_FC1_W)) {
          bool did = false;
          FV_TRY(wgrad<AT>(cx, (const AT*)b.d_big, F, (const AT*)k.xn2, E, T, nullptr, GB(i, FERVIT_B_FC1_W), b.scratch,
                           GB(i, FERVIT_B_FC1_B), &did));
          if (!did)
            FV_TRY(colsum<AT>((const AT*)b.d_big, T, F, F, b.scratch, nullptr, 1.0f, GB(i, FERVIT_B_FC1_B), nodrop, st));
        }
        e = make_epilogue();
        e.residual = DX(cur); e.out_f32 = DX(cur ^ 1); e.ldo = E;
        FV_TRY(linear<AT>(cx, (const AT*)b.d_big, T, p->bslot(i, FERVIT_B_FC1_W), true, e));
        cur ^= 1;
        // norm1 backward
        {
          float* part = GB(i, FERVIT_B_LN1_W) ? b.scratch : nullptr;
          FV_TRY((layernorm_bwd<float, AT>(DX(cur), k.x_mid, k.m1, k.r1, p->PB(i, FERVIT_B_LN1_W), nullptr, T, E,
                                           DX(cur ^ 1), (AT*)b.d_e1, part, cx.site(i, 1), st)));
          if (part) {
            const int g = layernorm_bwd_grid(T);
            FV_CHECK(GB(i, FERVIT_B_LN1_B), "backward: LayerNorm gradients must be requested together");
            FV_TRY(colsum_reduce_partials(part, g, 2 * E, GB(i, FERVIT_B_LN1_W), st, GB(i, FERVIT_B_LN1_B), E));
          }
          cur ^= 1;
        }
        const AT* dy1 = (const AT*)b.d_e1;  // masked dr1
        if (GB(i, FERVIT_B_PROJ_W)) {
          bool did = false;   // dy1 IS the masked gradient
          FV_TRY(wgrad<AT>(cx, dy1, E, (const AT*)k.ao, E, T, nullptr, GB(i, FERVIT_B_PROJ_W), b.scratch,
                           GB(i, FERVIT_B_PROJ_B), &did));
          if (!did)
            FV_TRY(colsum<float>(DX(cur), T, E, E, b.scratch, nullptr, 1.0f, GB(i, FERVIT_B_PROJ_B), cx.site(i, 1), st));
        }
        e = make_epilogue();
        e.out = b.d_e2; e.ldo = E;
        FV_TRY(linear<AT>(cx, dy1, T, p->bslot(i, FERVIT_B_PROJ_W), true, e));
        FV_TRY(attention_bwd<AT>((const AT*)k.qkv, (const AT*)k.ao, (const AT*)b.d_e2, k.lse, (AT*)b.d_big, B, S, c.H,
                                 p->HD, cx.site(i, 0), st));
        if (GB(i, FERVIT_B_QKV_W)) {
          bool did = false;
          FV_TRY(wgrad<AT>(cx, (const AT*)b.d_big, 3 * E, (const AT*)b.x_at[i], E, T, nullptr, GB(i, FERVIT_B_QKV_W),
                           b.scratch, GB(i, FERVIT_B_QKV_B), &did));
          if (!did)
            FV_TRY(colsum<AT>((const AT*)b.d_big, T, 3 * E, 3 * E, b.scratch, nullptr, 1.0f, GB(i, FERVIT_B_QKV_B), nodrop,
                              st));
        }
        e = make_epilogue();
        e.residual = DX(cur); e.out_f32 = DX(cur ^ 1); e.ldo = E;
        FV_TRY(linear<AT>(cx, (const AT*)b.d_big, T, p->bslot(i, FERVIT_B_QKV_W), true, e));
        cur ^= 1;
      }
      p->bwd_cur = cur;
    } else {
      // ---------------- input stage ----------------
      const int cur = p->bwd_cur;
      const Dropout din = make_dropout((training && c.input_dropout) ? c.dropout : 0.f, seed, FERVIT_SITE_INPUT, seed_dev);
      if (G[FERVIT_G_POS]) {
        FV_TRY(colsum<float>(b.dx[cur], B, S * E, (long long)S * E, b.scratch, nullptr, 1.0f, G[FERVIT_G_POS], din, st));
        if (G[FERVIT_G_CLS])
          FV_CUDA(cudaMemcpyAsync(G[FERVIT_G_CLS], G[FERVIT_G_POS], E * sizeof(float), cudaMemcpyDeviceToDevice, st));
      } else if (G[FERVIT_G_CLS]) {
        FV_TRY(colsum<float>(b.dx[cur], B, E, (long long)S * E, b.scratch, nullptr, 1.0f, G[FERVIT_G_CLS], din, st));
      }
      const bool pre_grads = p->has_pre && (G[FERVIT_G_SPE_LAYER] || G[FERVIT_G_LWN_GAMMA] || G[FERVIT_G_LEAM_W] ||
                                            G[FERVIT_G_LWN_GATE] || G[FERVIT_G_SPE_GROUP]);
      if (G[FERVIT_G_IN_W] || pre_grads) {
        FV_TRY(gather_tokens<AT>(b.dx[cur], (AT*)b.dtok, B, c.L, E, din, st));
        const AT* a_in = (F32 && c.input_kind == 0 && !p->has_pre) ? (const AT*)x : (const AT*)b.a_in;
        if (G[FERVIT_G_IN_W]) {
          FV_CHECK(G[FERVIT_G_IN_B] != nullptr, "backward: input projection weight and bias gradients go together");
          FV_TRY(wgrad<AT>(cx, (const AT*)b.dtok, E, a_in, c.Din, Tl, nullptr, G[FERVIT_G_IN_W], b.scratch));
          FV_TRY(colsum<AT>((const AT*)b.dtok, Tl, E, E, b.scratch, nullptr, 1.0f, G[FERVIT_G_IN_B], nodrop, st));
        }
        if (pre_grads) {
          Epilogue e = make_epilogue();
          e.out = b.dain; e.ldo = c.Din;
          FV_TRY(linear<AT>(cx, (const AT*)b.dtok, Tl, FERVIT_G_IN_W, true, e));
          FV_TRY(premodules_bwd<AT>(x, (const AT*)b.dain, B, c.L, c.Din, pre_params(p), nullptr, b.scratch,
                                    G[FERVIT_G_LWN_GAMMA], G[FERVIT_G_LWN_BETA], G[FERVIT_G_SPE_LAYER],
                                    G[FERVIT_G_SPE_GROUP], G[FERVIT_G_LWN_GATE], G[FERVIT_G_LEAM_W], st));
        }
      }
    }
  }
  if (side_used) {  // everything forked in this call is complete before the gradients are finalised / handed out
    FV_CUDA(cudaEventRecord(p->ev_tail, p->side));
    FV_CUDA(cudaStreamWaitEvent(st, p->ev_tail, 0));
  }
  if (ad_hi >= ad_lo) {
    if constexpr (!F32) {
      // blocks ad_lo..ad_hi are contiguous in the pools: K = n*T rows, every split / chunk inside one block
      const int n = ad_hi - ad_lo + 1;
      int spb = ad_defer_spb(n, T, 6);
      while (n * spb > b.ad_slabs && spb > 1) --spb;
      while ((T / 64) % spb) --spb;
      const int rpc = T / AD_CPB;
      FV_TRY(wgrad_partial<AT>(cx, AD_DY(ad_lo), E, (const AT*)b.blk[ad_lo].ga, A, n * T, n * spb, b.ad_w2_pool));
      FV_TRY(wgrad_partial<AT>(cx, AD_DU(ad_lo), A, (const AT*)b.blk[ad_lo].x2_at, E, n * T, n * spb, b.ad_w1_pool));
      FV_TRY(colsum_partial_rows<AT>(AD_DY(ad_lo), n * T, E, E, rpc, b.ad_csdy_pool, st));
      FV_TRY(colsum_partial_rows<AT>(AD_DU(ad_lo), n * T, A, A, rpc, b.ad_csdu_pool, st));
      for (int i = ad_lo; i <= ad_hi; ++i) {
        if (!GB(i, FERVIT_B_AD2_W)) continue;   // a frozen adapter inside the range: its slabs are simply not used
        const int g = i - ad_lo;
        AdapterGradJob job;
        job.w2_part = b.ad_w2_pool + (size_t)g * spb * E * A;
        job.w1_part = b.ad_w1_pool + (size_t)g * spb * E * A;
        job.cs_dy = b.ad_csdy_pool + (size_t)g * AD_CPB * E;
        job.cs_du = b.ad_csdu_pool + (size_t)g * AD_CPB * A;
        job.W2 = p->PB(i, FERVIT_B_AD2_W); job.b2 = p->PB(i, FERVIT_B_AD2_B);
        job.alpha_ptr = p->PB(i, FERVIT_B_ALPHA);
        job.dW2 = GB(i, FERVIT_B_AD2_W); job.dW1 = GB(i, FERVIT_B_AD1_W); job.db2 = GB(i, FERVIT_B_AD2_B);
        job.db1 = GB(i, FERVIT_B_AD1_B); job.dalpha = GB(i, FERVIT_B_ALPHA);
        job.s2 = spb; job.s1 = spb; job.chunks = AD_CPB; job.E = E; job.A = A;
        ad_jobs.push_back(job);
      }
    }
  }
  if (!ad_jobs.empty()) FV_TRY(adapter_grad_finalize(ad_jobs.data(), (int)ad_jobs.size(), b.ad_fin, st));
  return 0;
}

}  // namespace
}  // namespace fervit

// =================================================================================================
// C ABI
// =================================================================================================
#define FV_API extern "C" __attribute__((visibility("default")))

FV_API int fervit_plan_create(const fervit_config* cfg, fervit_plan** out) {
  FV_CHECK(cfg && out, "plan_create: null argument");
  const fervit_config& c = *cfg;
  FV_CHECK(c.mode == FERVIT_F32 || c.mode == FERVIT_BF16, "plan_create: unknown mode %d", c.mode);
  FV_CHECK(c.E > 0 && c.H > 0 && c.E % c.H == 0, "plan_create: embed dim %d not divisible by heads %d", c.E, c.H);
  const int hd = c.E / c.H;
  FV_CHECK(hd == 32 || hd == 48 || hd == 64, "plan_create: head dim %d not supported (32, 48, 64)", hd);
  FV_CHECK(c.E % 16 == 0 && c.F % 16 == 0 && c.Din % 16 == 0, "plan_create: E, F and Din must be multiples of 16");
  FV_CHECK(c.E <= 1024, "plan_create: E must be <= 1024");
  FV_CHECK(c.adapter_dim % 16 == 0, "plan_create: adapter_dim must be a multiple of 16");
  FV_CHECK(c.C >= 1 && c.C <= 16, "plan_create: num_classes must be in [1,16]");
  FV_CHECK(c.depth >= 1 && c.L >= 1, "plan_create: depth and L must be positive");
  FV_CHECK(c.act == FERVIT_ACT_RELU || c.act == FERVIT_ACT_GELU, "plan_create: activation must be relu or gelu");
  FV_CHECK(!(c.adapter_dim && !c.norm_first), "plan_create: adapters are only defined for pre-norm (timm) blocks");
  FV_CHECK(!(c.norm_first && c.dropout > 0.f), "plan_create: block dropout is only defined for post-norm (torch) blocks");
  if (c.input_kind == 1) {
    FV_CHECK(c.patch > 0 && c.img_h % c.patch == 0 && c.img_w % c.patch == 0, "plan_create: patch must divide the image");
    FV_CHECK((c.img_h / c.patch) * (c.img_w / c.patch) == c.L, "plan_create: L must equal the number of patches");
    FV_CHECK(c.img_c * c.patch * c.patch == c.Din, "plan_create: Din must equal img_c*patch*patch");
    FV_CHECK(!(c.use_spe || c.use_lwn || c.use_leam), "plan_create: pre-modules apply to latent inputs only");
  }
  if (c.use_spe || c.use_lwn || c.use_leam) FV_CHECK(c.Din <= 1024, "plan_create: latent_dim must be <= 1024 with pre-modules");
  fervit_plan* p = new fervit_plan();
  p->cfg = c;
  p->S = c.L + 1;
  p->HD = hd;
  p->A = c.adapter_dim;
  p->has_pre = c.use_spe || c.use_lwn || c.use_leam;
  p->params.assign(p->nslots(), nullptr);
  p->wb_off.assign(p->nslots(), SIZE_MAX);
  p->wbt_off.assign(p->nslots(), SIZE_MAX);
  p->wcache = nullptr;
  p->bwd_cur = 0;
  size_t off = 0;
  if (c.mode == FERVIT_BF16) {
    for (int s = 0; s < p->nslots(); ++s) {
      int R, C;
      if (weight_shape(p, s, &R, &C)) {
        off = (off + 255) & ~size_t(255);
        p->wb_off[s] = off;
        off += (size_t)R * C * 2;
        off = (off + 255) & ~size_t(255);
        p->wbt_off[s] = off;
        off += (size_t)R * C * 2;
      }
    }
  }
  p->fold1.assign(c.depth, 0);
  p->fold2.assign(c.depth, 0);
  p->fb_off.assign(p->nslots(), SIZE_MAX);
  p->fcs_off.assign(p->nslots(), SIZE_MAX);
  if (c.mode == FERVIT_BF16 && c.norm_first) {
    for (int i = 0; i < c.depth; ++i) {
      const int slots[2] = {fervit_plan::bslot(i, FERVIT_B_QKV_W), fervit_plan::bslot(i, FERVIT_B_FC1_W)};
      for (int s : slots) {
        int R, C;
        if (!weight_shape(p, s, &R, &C)) continue;
        off = (off + 255) & ~size_t(255);
        p->fb_off[s] = off;
        off += (size_t)R * 4;
        off = (off + 255) & ~size_t(255);
        p->fcs_off[s] = off;
        off += (size_t)R * 4;
      }
    }
  }
  if (c.mode == FERVIT_BF16) {
    off = (off + 255) & ~size_t(255);
    p->sk_off = off;
    p->sk_bytes = gemm_tc2_scratch_bytes();
    off += p->sk_bytes;
  }
  p->wcache_bytes = off;
  *out = p;
  return 0;
}

FV_API void fervit_plan_destroy(fervit_plan* plan) { delete plan; }

FV_API int fervit_plan_num_slots(const fervit_plan* plan) { return plan ? plan->nslots() : 0; }

FV_API long long fervit_plan_slot_numel(const fervit_plan* plan, int slot) {
  if (!plan || slot < 0 || slot >= plan->nslots()) return 0;
  return slot_numel(plan, slot);
}

FV_API int fervit_plan_set_params(fervit_plan* plan, const void* const* params, int n) {
  FV_CHECK(plan && params, "set_params: null argument");
  FV_CHECK(n == plan->nslots(), "set_params: expected %d slots, got %d", plan->nslots(), n);
  for (int s = 0; s < n; ++s) {
    if (slot_numel(plan, s) > 0) FV_CHECK(params[s] != nullptr, "set_params: parameter slot %d is required", s);
    plan->params[s] = params[s];
  }
  return 0;
}

FV_API long long fervit_plan_wcache_bytes(const fervit_plan* plan) { return plan ? (long long)plan->wcache_bytes : 0; }

FV_API int fervit_plan_set_wcache(fervit_plan* plan, void* ptr, long long bytes) {
  FV_CHECK(plan, "set_wcache: null plan");
  FV_CHECK(bytes >= (long long)plan->wcache_bytes, "set_wcache: buffer too small");
  FV_CHECK(plan->wcache_bytes == 0 || ((uintptr_t)ptr & 255) == 0, "set_wcache: buffer must be 256-byte aligned");
  plan->wcache = reinterpret_cast<char*>(ptr);
  if (plan->sk_bytes) {
    // arrival flags of the stream-K GEMMs start at zero and every launch leaves them at zero (set-up path: blocking)
    FV_CUDA(cudaMemset(plan->wcache + plan->sk_off, 0, 4096));
    FV_CUDA(cudaDeviceSynchronize());
  }
  return 0;
}

FV_API int fervit_plan_refresh_wcache(fervit_plan* plan, const int* slots, int n, void* stream) {
  FV_CHECK(plan, "refresh_wcache: null plan");
  if (plan->cfg.mode != FERVIT_BF16) return 0;
  FV_CHECK(plan->wcache != nullptr, "refresh_wcache: no cache buffer set");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // every requested matrix goes into batched launches (32 matrices each): the trainable weights are re-cast on every
  // step, and one launch per matrix would cost more than the casts themselves
  std::vector<const float*> src;
  std::vector<bf16*> dst, dst_t;
  std::vector<int> Rs, Cs;
  auto add = [&](int s) -> int {
    int R, C;
    if (!weight_shape(plan, s, &R, &C)) return 0;
    FV_CHECK(plan->params[s] != nullptr, "refresh_wcache: parameter slot %d not set", s);
    if (s >= FERVIT_NUM_GLOBAL) {
      const int blk = (s - FERVIT_NUM_GLOBAL) / FERVIT_NUM_BLOCK, w = (s - FERVIT_NUM_GLOBAL) % FERVIT_NUM_BLOCK;
      const bool f1 = w == FERVIT_B_QKV_W && plan->fold1[blk], f2 = w == FERVIT_B_FC1_W && plan->fold2[blk];
      if (f1 || f2) {   // W diag(gamma), b + W beta, column sums: the norm in front of this weight is folded into it
        const float* g = plan->PB(blk, f1 ? FERVIT_B_LN1_W : FERVIT_B_LN2_W);
        const float* be = plan->PB(blk, f1 ? FERVIT_B_LN1_B : FERVIT_B_LN2_B);
        const float* bias = plan->PB(blk, f1 ? FERVIT_B_QKV_B : FERVIT_B_FC1_B);
        FV_CHECK(g && be && bias, "refresh_wcache: a folded LayerNorm needs its parameters and the bias bound");
        return fold_ln_weight(plan->P(s), bias, g, be, R, C, const_cast<bf16*>(plan->WB(s)),
                              const_cast<bf16*>(plan->WBT(s)), const_cast<float*>(plan->FB(s)),
                              const_cast<float*>(plan->FCS(s)), st);
      }
    }
    src.push_back(plan->P(s));
    dst.push_back(const_cast<bf16*>(plan->WB(s)));
    dst_t.push_back(const_cast<bf16*>(plan->WBT(s)));
    Rs.push_back(R);
    Cs.push_back(C);
    return 0;
  };
  if (slots == nullptr) {
    for (int s = 0; s < plan->nslots(); ++s) FV_TRY(add(s));
  } else {
    for (int i = 0; i < n; ++i) {
      FV_CHECK(slots[i] >= 0 && slots[i] < plan->nslots(), "refresh_wcache: bad slot %d", slots[i]);
      FV_TRY(add(slots[i]));
    }
  }
  if (src.empty()) return 0;
  return weight_cache_batch(src.data(), Rs.data(), Cs.data(), dst.data(), dst_t.data(), (int)src.size(), st);
}

FV_API int fervit_plan_set_ln_fold(fervit_plan* plan, const int* fold_norm1, const int* fold_norm2, int depth) {
  FV_CHECK(plan && fold_norm1 && fold_norm2, "set_ln_fold: null argument");
  const fervit_config& c = plan->cfg;
  FV_CHECK(depth == c.depth, "set_ln_fold: expected %d blocks, got %d", c.depth, depth);
  bool any = false;
  for (int i = 0; i < depth; ++i) any = any || fold_norm1[i] || fold_norm2[i];
  if (any) {
    FV_CHECK(c.mode == FERVIT_BF16 && c.norm_first, "set_ln_fold: folding is defined for bf16-mode pre-norm blocks only");
    FV_CHECK(c.E % 128 == 0 && c.E <= 1024, "set_ln_fold: embed dim %d must be a multiple of 128 (<= 1024)", c.E);
    FV_CHECK(!fold_norm1[0], "set_ln_fold: block 0's norm1 reads the token projection's output and stays a kernel");
  }
  for (int i = 0; i < depth; ++i) {
    plan->fold1[i] = fold_norm1[i] ? 1 : 0;
    plan->fold2[i] = fold_norm2[i] ? 1 : 0;
  }
  return 0;
}

FV_API long long fervit_plan_workspace_bytes(const fervit_plan* plan, int B, int save_for_backward) {
  if (!plan || B <= 0) return 0;
  Arena ar{nullptr, 0};
  Bufs b;
  if (plan->cfg.mode == FERVIT_BF16) carve<bf16>(plan, B, save_for_backward != 0, ar, b);
  else carve<float>(plan, B, save_for_backward != 0, ar, b);
  return (long long)ar.off + 256;
}

FV_API int fervit_plan_forward(fervit_plan* plan, const float* x, int B, void* ws, long long ws_bytes, int training,
                               int save_for_backward, unsigned long long seed, const unsigned long long* seed_dev,
                               float* logits, void* stream) {
  FV_CHECK(plan && x && ws && logits, "forward: null argument");
  FV_CHECK(B > 0, "forward: empty batch");
  FV_CHECK(((uintptr_t)ws & 255) == 0, "forward: workspace must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (plan->cfg.mode == FERVIT_BF16) {
    FV_CHECK(plan->wcache != nullptr, "forward: bf16 mode needs the weight cache (set_wcache + refresh_wcache)");
    return forward_impl<bf16>(plan, x, B, ws, ws_bytes, training != 0, save_for_backward != 0, seed, seed_dev, logits, st);
  }
  return forward_impl<float>(plan, x, B, ws, ws_bytes, training != 0, save_for_backward != 0, seed, seed_dev, logits, st);
}

FV_API int fervit_plan_num_stages(const fervit_plan* plan) { return plan ? plan->cfg.depth + 2 : 0; }

FV_API int fervit_plan_saved_buffer(const fervit_plan* plan, void* ws, int B, int block, int which, void** ptr,
                                    long long* numel) {
  FV_CHECK(plan && ws && ptr && numel, "saved_buffer: null argument");
  FV_CHECK(B > 0 && block >= 0 && block < plan->cfg.depth, "saved_buffer: bad batch or block");
  Arena ar{reinterpret_cast<char*>(ws), 0};
  Bufs b;
  if (plan->cfg.mode == FERVIT_BF16) carve<bf16>(plan, B, true, ar, b);
  else carve<float>(plan, B, true, ar, b);
  const long long T = (long long)B * plan->S;
  switch (which) {
    case FERVIT_SAVED_ACT_DERIV: *ptr = b.blk[block].u1; *numel = T * plan->cfg.F; break;
    case FERVIT_SAVED_ACT_OUT: *ptr = b.blk[block].g1; *numel = T * plan->cfg.F; break;
    case FERVIT_SAVED_QKV: *ptr = b.blk[block].qkv; *numel = T * 3 * plan->cfg.E; break;
    default: FV_CHECK(false, "saved_buffer: unknown buffer id %d", which);
  }
  return 0;
}

FV_API int fervit_plan_backward(fervit_plan* plan, const float* x, int B, void* ws, long long ws_bytes, int training,
                                unsigned long long seed, const unsigned long long* seed_dev, const float* dlogits,
                                float* const* grads, int n,
                                int stage_begin, int stage_end, void* stream) {
  FV_CHECK(plan && x && ws && dlogits && grads, "backward: null argument");
  FV_CHECK(n == plan->nslots(), "backward: expected %d gradient slots, got %d", plan->nslots(), n);
  FV_CHECK(stage_begin >= 0 && stage_end <= plan->cfg.depth + 2 && stage_begin <= stage_end, "backward: bad stage range");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (plan->cfg.mode == FERVIT_BF16)
    return backward_impl<bf16>(plan, x, B, ws, ws_bytes, training != 0, seed, seed_dev, dlogits, grads, stage_begin, stage_end, st);
  return backward_impl<float>(plan, x, B, ws, ws_bytes, training != 0, seed, seed_dev, dlogits, grads, stage_begin, stage_end, st);
}
