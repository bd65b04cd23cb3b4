"""ORACLE — test infrastructure, not product code.

A CPU restatement, in elementary torch tensor arithmetic (matmul / exp / erf / mean — no nn.Module, no fused
attention, no nn.functional layers), of the reference's algorithm for the train-step path. Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline leg may import this package; fer_vit_b200/ never does.

Every function cites the reference file:line it follows (paths relative to the reference repo). Two third-party
dependencies hold arithmetic that is not in the reference tree:
  * torch.nn.TransformerEncoderLayer / MultiheadAttention (reference pins pytorch=2.5.1, environment.yml:113) —
    restated from the installed torch (2.11.0) sources, torch/nn/modules/transformer.py:944-982 (training path) and
    torch/nn/functional.py multi_head_attention_forward; checked against the real modules in tests/golden/make_golden.py.
  * timm.models.vision_transformer.Block (reference pins timm=1.0.17, environment.yml:124; NOT installed here) —
    restated from the published algorithm (pre-norm; LayerNorm eps 1e-6; fused [Q;K;V] Linear with bias; scale
    hd^-0.5; softmax over keys; exact-erf GELU MLP; no dropout / layer-scale / drop-path at create_model defaults).
    Parity for this block is pinned against two independent pre-norm implementations available here
    (nn.TransformerEncoderLayer(norm_first=True, activation='gelu', layer_norm_eps=1e-6) and HF transformers ViTLayer),
    not against timm itself: see DESIGN.md "Oracle pinning".

Inputs are a `state_dict`-style mapping with the reference's own keys, so the same dictionary drives the reference
classes, this oracle and the CUDA implementation. Gradients come from autograd over these elementary ops.
Dropout is expressed as explicit multiplicative masks (already scaled by 1/(1-p)) supplied by the caller, because
torch's CPU and CUDA generators cannot reproduce each other's masks.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional

import torch

Tensor = torch.Tensor
Masks = Optional[Mapping[object, Tensor]]


def _mask(x: Tensor, masks: Masks, key) -> Tensor:
    if masks is not None and key in masks:
        return x * masks[key].to(x.dtype)
    return x


# ------------------------------------------------------------------------------------------------
# Elementary layers
# ------------------------------------------------------------------------------------------------
def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """nn.Linear: x W^T + b (latent_vit.py:40, hybrid_latent_vit.py:215)."""
    y = x @ w.t()
    return y + b if b is not None else y


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """nn.LayerNorm over the last dim, biased variance (latent_vit.py:34, hybrid_latent_vit.py:111)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu(x: Tensor) -> Tensor:
    """Exact-erf GELU: nn.GELU() default (hybrid_latent_vit.py:259), timm Mlp act, activation='gelu' (image_vit.py:106)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def relu(x: Tensor) -> Tensor:
    """nn.TransformerEncoderLayer default activation (latent_vit.py:24-30 passes none)."""
    return torch.clamp_min(x, 0.0)


def attention(x: Tensor, w_qkv: Tensor, b_qkv: Tensor, w_o: Tensor, b_o: Tensor, heads: int, masks: Masks = None,
              blk: int = 0) -> Tensor:
    """Multi-head self-attention with packed [Q;K;V] projection (rows of w_qkv), each head-major.

    Same packing for timm Attention.qkv and torch in_proj_weight. Scale hd^-0.5, softmax
    over keys, optional dropout on the attention weights (torch layers only), then out-projection.
    """
    B, S, E = x.shape
    hd = E // heads
    qkv = linear(x, w_qkv, b_qkv).reshape(B, S, 3, heads, hd).permute(2, 0, 3, 1, 4)  # 3,B,H,S,hd
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
    s = s - s.max(dim=-1, keepdim=True).values
    p = torch.exp(s)
    p = p / p.sum(dim=-1, keepdim=True)
    p = _mask(p, masks, ("attn", blk))
    o = (p @ v).permute(0, 2, 1, 3).reshape(B, S, E)
    return linear(o, w_o, b_o)


# ------------------------------------------------------------------------------------------------
# Blocks
# ------------------------------------------------------------------------------------------------
def timm_block(x: Tensor, sd: Mapping[str, Tensor], prefix: str, heads: int, eps: float = 1e-6) -> Tensor:
    """timm Block.forward: x = x + attn(norm1(x)); x = x + mlp(norm2(x)) (called at hybrid_latent_vit.py:228/233)."""
    h = layer_norm(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], eps)
    x = x + attention(h, sd[prefix + "attn.qkv.weight"], sd[prefix + "attn.qkv.bias"], sd[prefix + "attn.proj.weight"],
                      sd[prefix + "attn.proj.bias"], heads)
    h = layer_norm(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], eps)
    h = gelu(linear(h, sd[prefix + "mlp.fc1.weight"], sd[prefix + "mlp.fc1.bias"]))
    return x + linear(h, sd[prefix + "mlp.fc2.weight"], sd[prefix + "mlp.fc2.bias"])


def adapter(x: Tensor, sd: Mapping[str, Tensor], prefix: str) -> Tensor:
    """AdapterModule.forward: x + alpha * (W2 GELU(W1 x + b1) + b2) (hybrid_latent_vit.py:264-265)."""
    h = gelu(linear(x, sd[prefix + "adapter.0.weight"], sd[prefix + "adapter.0.bias"]))
    return x + sd[prefix + "alpha"] * linear(h, sd[prefix + "adapter.2.weight"], sd[prefix + "adapter.2.bias"])


def torch_encoder_layer(x: Tensor, sd: Mapping[str, Tensor], prefix: str, heads: int, act, eps: float = 1e-5,
                        masks: Masks = None, blk: int = 0) -> Tensor:
    """nn.TransformerEncoderLayer post-norm training path (torch/nn/modules/transformer.py:944-958):
    x = norm1(x + dropout1(self_attn(x))); x = norm2(x + dropout2(linear2(dropout(act(linear1(x))))))."""
    a = attention(x, sd[prefix + "self_attn.in_proj_weight"], sd[prefix + "self_attn.in_proj_bias"],
                  sd[prefix + "self_attn.out_proj.weight"], sd[prefix + "self_attn.out_proj.bias"], heads, masks, blk)
    x = layer_norm(x + _mask(a, masks, ("drop1", blk)), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], eps)
    u = linear(x, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"])
    if masks is not None and "trace" in masks:
        masks["trace"][("ffn_pre", blk)] = u.detach()
    if masks is not None and ("relu", blk) in masks:
        # ReLU with an INJECTED active set (0/1, the one the implementation under test used): relu(u) = u * [u > 0] is
        # a selector times a linear map, and the selector is discontinuous exactly like a dropout mask — a unit whose
        # pre-activation is within round-off of zero may legitimately fall on either side. Given the selector, value
        # and gradient are the reference's (d relu / du = the selector).
        h = u * masks[("relu", blk)].to(u.dtype)
    else:
        h = act(u)
    h = _mask(h, masks, ("ffn", blk))
    f = _mask(linear(h, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"]), masks, ("drop2", blk))
    return layer_norm(x + f, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], eps)


# ------------------------------------------------------------------------------------------------
# w+ pre-modules
# ------------------------------------------------------------------------------------------------
LAYER_GROUPS = [0] * 4 + [1] * 8 + [2] * 6  # modules/semantic_pe.py:6-8


def semantic_pe(x: Tensor, group_embed: Tensor, layer_embed: Tensor, groups: Tensor) -> Tensor:
    """SemanticPE.forward: x + (group_embed[groups] + layer_embed[arange(L)]) (modules/semantic_pe.py:44-48)."""
    Ln = x.shape[1]
    pe = group_embed[groups[:Ln]] + layer_embed[:Ln]
    return x + pe.unsqueeze(0)


def layer_wise_norm(x: Tensor, gammas: Tensor, betas: Tensor, gate: Optional[Tensor], eps: float = 1e-5) -> Tensor:
    """LayerWiseNorm.forward: per-position LayerNorm; optional x + sigmoid(gate)(normed - x)
    (modules/layer_wise_norm.py:42-50). gammas/betas: [L, D]."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    normed = (x - mu) / torch.sqrt(var + eps) * gammas.unsqueeze(0) + betas.unsqueeze(0)
    if gate is None:
        return normed
    g = torch.sigmoid(gate).reshape(1, -1, 1)
    return x + g * (normed - x)


def leam(x: Tensor, layer_weights: Tensor) -> Tensor:
    """LEAM.forward: x * sigmoid(w)[None, :, None] (modules/leam.py:39-40)."""
    return x * torch.sigmoid(layer_weights).reshape(1, -1, 1)


def pre_modules(x: Tensor, sd: Mapping[str, Tensor], use_spe: bool, use_lwn: bool, use_lwn_residual: bool,
                use_leam: bool, prefix: str = "") -> Tensor:
    """LatentViTv2 pre-processing in the shipped order SPE -> LWN -> LEAM (latent_vit_v2.py:82-84)."""
    if use_spe:
        x = semantic_pe(x, sd[prefix + "spe.group_embed.weight"], sd[prefix + "spe.layer_embed.weight"],
                        sd[prefix + "spe.groups"])
    if use_lwn:
        Ln = x.shape[1]
        gam = torch.stack([sd[f"{prefix}lwn.norms.{i}.weight"] for i in range(Ln)])
        bet = torch.stack([sd[f"{prefix}lwn.norms.{i}.bias"] for i in range(Ln)])
        x = layer_wise_norm(x, gam, bet, sd[prefix + "lwn.gate"] if use_lwn_residual else None)
    if use_leam:
        x = leam(x, sd[prefix + "leam.layer_weights"])
    return x


# ------------------------------------------------------------------------------------------------
# Models
# ------------------------------------------------------------------------------------------------
def latent_vit_forward(sd: Mapping[str, Tensor], x: Tensor, depth: int, heads: int, masks: Masks = None,
                       prefix: str = "") -> Tensor:
    """LatentViT.forward (latent_vit.py:38-48): proj -> [cls; tokens] + pos_emb -> post-norm ReLU encoder ->
    cls row -> LayerNorm -> Linear."""
    h = linear(x, sd[prefix + "input_proj.weight"], sd[prefix + "input_proj.bias"])
    B = h.shape[0]
    h = torch.cat([sd[prefix + "cls_token"].expand(B, -1, -1), h], dim=1) + sd[prefix + "pos_emb"]
    for i in range(depth):
        h = torch_encoder_layer(h, sd, f"{prefix}transformer.layers.{i}.", heads, relu, 1e-5, masks, i)
    c = layer_norm(h[:, 0], sd[prefix + "mlp_head.0.weight"], sd[prefix + "mlp_head.0.bias"], 1e-5)
    return linear(c, sd[prefix + "mlp_head.1.weight"], sd[prefix + "mlp_head.1.bias"])


def latent_vit_v2_forward(sd: Mapping[str, Tensor], x: Tensor, depth: int, heads: int, use_spe: bool, use_lwn: bool,
                          use_lwn_residual: bool, use_leam: bool, masks: Masks = None) -> Tensor:
    """LatentViTv2.forward (latent_vit_v2.py:75-85)."""
    x = pre_modules(x, sd, use_spe, use_lwn, use_lwn_residual, use_leam)
    return latent_vit_forward(sd, x, depth, heads, masks, prefix="backbone.")


def hybrid_forward(sd: Mapping[str, Tensor], x: Tensor, depth: int, heads: int, use_adapter: bool,
                   masks: Masks = None) -> Tensor:
    """HybridLatentViT.forward (hybrid_latent_vit.py:205-239): proj -> [cls; tokens] + pos_embed ->
    (block, adapter)* -> cls row -> LayerNorm -> Dropout -> Linear."""
    h = linear(x, sd["input_proj.weight"], sd["input_proj.bias"])
    B = h.shape[0]
    h = torch.cat([sd["cls_token"].expand(B, -1, -1), h], dim=1) + sd["pos_embed"]
    for i in range(depth):
        h = timm_block(h, sd, f"transformer.{i}.", heads)
        if use_adapter:
            h = adapter(h, sd, f"adapters.{i}.")
    c = layer_norm(h[:, 0], sd["head.0.weight"], sd["head.0.bias"], 1e-5)
    c = _mask(c, masks, "head")
    return linear(c, sd["head.2.weight"], sd["head.2.bias"])


def patchify(img: Tensor, patch: int) -> Tensor:
    """Conv2d(k = s = patch) as a matmul operand: [B, C, H, W] -> [B, N, C*patch*patch], patch rows in the
    flatten(2).transpose(1, 2) order of PatchEmbedding.forward (image_vit.py:41-43)."""
    B, Cc, Hh, Ww = img.shape
    gh, gw = Hh // patch, Ww // patch
    p = img.reshape(B, Cc, gh, patch, gw, patch).permute(0, 2, 4, 1, 3, 5)
    return p.reshape(B, gh * gw, Cc * patch * patch)


def image_vit_forward(sd: Mapping[str, Tensor], img: Tensor, depth: int, heads: int, patch: int,
                      masks: Masks = None) -> Tensor:
    """ImageViT.forward (image_vit.py:138-166): patch embed -> cls -> + pos_embed -> dropout -> post-norm GELU
    encoder -> cls row -> norm -> head."""
    w = sd["patch_embed.proj.weight"]
    h = linear(patchify(img, patch), w.reshape(w.shape[0], -1), sd["patch_embed.proj.bias"])
    B = h.shape[0]
    h = torch.cat([sd["cls_token"].expand(B, -1, -1), h], dim=1) + sd["pos_embed"]
    h = _mask(h, masks, "input")
    for i in range(depth):
        h = torch_encoder_layer(h, sd, f"transformer.layers.{i}.", heads, gelu, 1e-5, masks, i)
    c = layer_norm(h[:, 0], sd["norm.weight"], sd["norm.bias"], 1e-5)
    return linear(c, sd["head.weight"], sd["head.bias"])


# ------------------------------------------------------------------------------------------------
# Loss
# ------------------------------------------------------------------------------------------------
def cross_entropy(logits: Tensor, labels: Tensor, weight: Optional[Tensor] = None, label_smoothing: float = 0.0) -> Tensor:
    """nn.CrossEntropyLoss(weight, label_smoothing), mean reduction (train_hybrid_latent_vit.py:236-241,
    train_latent_vit.py:248-253): sum_i [(1-e) w[y_i] (-lp[i,y_i]) + (e/C) sum_c w[c] (-lp[i,c])] / sum_i w[y_i]."""
    z = logits - logits.max(dim=-1, keepdim=True).values
    lp = z - torch.log(torch.exp(z).sum(dim=-1, keepdim=True))
    Cn = logits.shape[-1]
    w = weight.to(logits.dtype) if weight is not None else torch.ones(Cn, dtype=logits.dtype)
    wy = w[labels]
    nll = -(lp[torch.arange(logits.shape[0]), labels]) * wy
    smooth = -(lp * w.unsqueeze(0)).sum(dim=-1)
    den = wy.sum()
    return ((1.0 - label_smoothing) * nll.sum() + (label_smoothing / Cn) * smooth.sum()) / den


def grads_of(loss: Tensor, sd: Mapping[str, Tensor]) -> Dict[str, Tensor]:
    """d loss / d sd[k] for every entry that requires grad."""
    keys = [k for k, v in sd.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(keys, gs)}


# ------------------------------------------------------------------------------------------------
# Data-side prologue of the LatentViT train step: LatentAugment and mixup
# ------------------------------------------------------------------------------------------------
def latent_augment(latent: Tensor, noise_std: float, scale_range, mask_prob: float, normal: Optional[Tensor] = None,
                   scale: Optional[Tensor] = None, keep: Optional[Tensor] = None) -> Tensor:
    """``LatentAugment.__call__`` (data/latent_dataset.py:28-49) with the random draws injected: additive noise
    ``normal * noise_std``, then one ``scale`` factor per sample, then the element mask ``keep``; a disabled stage
    (``noise_std == 0`` / ``scale_range is None`` / ``mask_prob == 0``) is skipped as in the reference. ``latent`` is
    one sample ``[18, 512]`` or a batch ``[B, 18, 512]`` (then ``scale`` is ``[B]``: the dataset augments per sample)."""
    out = latent.clone()
    if noise_std > 0:
        out = out + normal.to(out.dtype) * noise_std
    if scale_range is not None:
        s = scale.to(out.dtype)
        out = out * (s.reshape(-1, *([1] * (out.dim() - 1))) if out.dim() == 3 else s)
    if mask_prob > 0:
        out = out * keep.to(out.dtype)
    return out


def mixup(latents: Tensor, index: Tensor, lam: float) -> Tensor:
    """``mixed_latents = lam * latents + (1 - lam) * latents[index]`` (train_latent_vit.py:126-127)."""
    return lam * latents + (1 - lam) * latents[index]


def mixup_loss(logits: Tensor, labels: Tensor, index: Tensor, lam: float, weight: Optional[Tensor] = None,
               label_smoothing: float = 0.0) -> Tensor:
    """``lam * criterion(logits, labels) + (1 - lam) * criterion(logits, labels[index])`` (train_latent_vit.py:131)."""
    return (lam * cross_entropy(logits, labels, weight, label_smoothing)
            + (1 - lam) * cross_entropy(logits, labels[index], weight, label_smoothing))


# Counter-based draws of the device-side augmentation (include/fervit_b200.h: fervit_latent_batch). This is the
# definition of the product's generator restated in numpy so that tests can replay the very draws the kernel used; the
# reference draws from torch's global CPU generator, which no device kernel can reproduce.
_SITE_NOISE, _SITE_SCALE, _SITE_MASK = 0x4C410000, 0x4C410002, 0x4C410003


def mix_hash64(seed: int, site: int, idx):
    """64-bit mix of (seed, site, idx) - fer_vit_b200/csrc/common.cuh: mix_hash64 (splitmix64 finaliser)."""
    import numpy as np
    with np.errstate(over="ignore"):
        idx = np.asarray(idx, dtype=np.uint64)
        z = (np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(site + 1)
             + idx * np.uint64(0xD1B54A32D192ED03))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def mix_hash(seed: int, site: int, idx):
    """Its high word (common.cuh: mix_hash): the generator of the data path's draws."""
    import numpy as np
    return (mix_hash64(seed, site, idx) >> np.uint64(32)).astype(np.uint32)


def drop_hash(seed: int, site: int, idx):
    """The kernels' dropout generator (common.cuh: drop_hash): a 64-bit key per (seed, site) from mix_hash64, then two
    32-bit multiply-xorshift rounds per element, the key's high word (and the index's) entering between them."""
    import numpy as np
    with np.errstate(over="ignore"):
        idx = np.asarray(idx, dtype=np.uint64)
        key = int(mix_hash64(seed, site, [0x5DEECE66D])[0])
        k1, k2 = np.uint32(key & 0xFFFFFFFF), np.uint32(key >> 32)
        lo = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hi = (idx >> np.uint64(32)).astype(np.uint32)
        x = lo * np.uint32(0x9E3779B1) + k1
        x = x ^ (x >> np.uint32(16))
        x = x * np.uint32(0x7FEB352D)
        x = x ^ (k2 + hi * np.uint32(0x85EBCA6B))
        x = x ^ (x >> np.uint32(15))
        x = x * np.uint32(0x846CA68B)
        return x ^ (x >> np.uint32(16))


def dropout_keep(seed: int, site: int, idx, p: float):
    """keep-mask of the kernels' counter-based dropout (common.cuh: drop_keep / drop_threshold): element `idx` of
    dropout site `site` survives iff drop_hash is >= floor(p * 2^32); survivors are scaled by 1 / (1 - p)."""
    import numpy as np
    t = float(np.float32(p)) * 4294967296.0
    thr = 0 if t <= 0 else (4294967295 if t >= 4294967295.0 else int(t))
    return torch.from_numpy(drop_hash(seed, site, idx) >= np.uint32(thr))


def latent_augment_draws(seed: int, B: int, row: int, scale_range=None, mask_prob: float = 0.0):
    """(normal [B, row] f64, scale [B] f64, keep [B, row] bool) for batch positions 0..B-1 under `seed`, as
    include/fervit_b200.h (fervit_latent_batch) defines them."""
    import numpy as np
    u = lambda h: ((h & np.uint64(0xFFFFFFFF)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 4294967296.0)
    g = np.arange(B * row, dtype=np.uint64)
    z = mix_hash64(seed, _SITE_NOISE, g >> np.uint64(1))
    r = np.sqrt(-2.0 * np.log(u(z >> np.uint64(32)).astype(np.float64)))
    t = 2.0 * np.pi * (u(z).astype(np.float64) - 0.5)
    normal = np.where((g & np.uint64(1)) == 0, r * np.cos(t), r * np.sin(t)).reshape(B, row)
    lo, hi = scale_range if scale_range is not None else (1.0, 1.0)
    # the kernel forms the uniforms and the scale factor in fp32
    uf = u(mix_hash64(seed, _SITE_SCALE, np.arange(B, dtype=np.uint64)) >> np.uint64(32))
    scale = (np.float32(lo) + (np.float32(hi) - np.float32(lo)) * uf).astype(np.float64)
    thr = max(1, int(float(np.float32(mask_prob)) * 65536.0)) if mask_prob > 0 else 0
    lanes = (mix_hash64(seed, _SITE_MASK, g >> np.uint64(2)) >> (np.uint64(16) * (g & np.uint64(3)))) & np.uint64(0xFFFF)
    keep = (lanes >= np.uint64(thr)).reshape(B, row)
    return torch.from_numpy(normal), torch.from_numpy(scale), torch.from_numpy(keep)


# ------------------------------------------------------------------------------------------------
# LatentDecomposer / ExpressionAwareViT front-end
# ------------------------------------------------------------------------------------------------
def normalize_directions(directions: Tensor) -> Tensor:
    """Stacked directions [C, L, D] with every flattened direction scaled to unit norm
    (latent_decomposer.py:58-65)."""
    flat = directions.reshape(directions.shape[0], -1)
    return (flat / (flat.norm(dim=1, keepdim=True) + 1e-12)).reshape(directions.shape)


def latent_decompose(w_plus: Tensor, directions: Tensor, mode: str = "all_classes"):
    """(w_expr, w_id, coefficients) of latent_decomposer.py:82-118: coefficients = flattened w+ times the unit
    directions; the expression part is their recombination (all classes) or the single largest-|coef| term."""
    B = w_plus.shape[0]
    d = directions.reshape(directions.shape[0], -1).to(w_plus.dtype)
    flat = w_plus.reshape(B, -1)
    coef = flat @ d.T
    if mode == "all_classes":
        expr = coef @ d
    elif mode == "max_class":
        best = coef.abs().argmax(dim=1)
        expr = coef[torch.arange(B), best].reshape(B, 1) * d[best]
    else:
        raise ValueError(f"Unknown mode: {mode!r}")
    expr = expr.reshape(w_plus.shape)
    return expr, w_plus - expr, coef


def decomposer_forward(w_plus: Tensor, directions: Tensor, output_mode: str = "expr_only", enhance_alpha: float = 2.0,
                       decompose_mode: str = "all_classes") -> Tensor:
    """LatentDecomposer.forward (latent_decomposer.py:146-173)."""
    expr, ident, _ = latent_decompose(w_plus, directions, decompose_mode)
    if output_mode == "expr_only":
        return expr
    if output_mode == "id_only":
        return ident
    if output_mode == "enhanced":
        return ident + enhance_alpha * expr
    if output_mode == "concat":
        return torch.cat([expr, ident], dim=1)
    raise ValueError(f"Unknown output_mode: {output_mode!r}")


# ------------------------------------------------------------------------------------------------
# Optimizer step of the trainers: clip_grad_norm_ + AdamW (train_latent_vit_v2.py:132-134, :263;
# train_hybrid_latent_vit.py:244-248)
# ------------------------------------------------------------------------------------------------
def clip_grad_norm(grads: Mapping[str, Tensor], max_norm: float):
    """torch.nn.utils.clip_grad_norm_ (L2): every gradient times min(1, max_norm / (total_norm + 1e-6)).
    Returns (clipped gradients, total norm)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return {k: g * coef.to(g.dtype) for k, g in grads.items()}, total


def adamw_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, step: int, lr: float,
               betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
    """One torch.optim.AdamW update (decoupled weight decay, bias correction, amsgrad off) for step number `step`
    (1-based). Returns (param, exp_avg, exp_avg_sq)."""
    b1, b2 = betas
    param = param * (1 - lr * weight_decay)
    exp_avg = b1 * exp_avg + (1 - b1) * grad
    exp_avg_sq = b2 * exp_avg_sq + (1 - b2) * grad * grad
    denom = exp_avg_sq.sqrt() / math.sqrt(1 - b2 ** step) + eps
    return param - (lr / (1 - b1 ** step)) * exp_avg / denom, exp_avg, exp_avg_sq
