"""Parameter containers with timm's ViT attribute names, used when timm itself is not installed.

``timm.models.vision_transformer.Block`` (timm 1.0.17, pinned by the reference's environment.yml:124) is a
third-party dependency that is absent from this image. HybridLatentViT only needs the blocks' parameters under the
keys ``norm1.*, attn.qkv.*, attn.proj.*, norm2.*, mlp.fc1.*, mlp.fc2.*`` plus ``embed_dim``, ``cls_token`` and
``pos_embed`` of the ViT they came from (hybrid_latent_vit.py:68-93), so this file provides exactly that shell.
The torch-op ``forward`` methods exist for API compatibility (evaluate_model.py:255 calls ``model.transformer(x)``
directly); HybridLatentViT.forward never calls them.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

# name -> (embed_dim, depth, heads); all patch16 / 224 => 196 patches + cls
VIT_CONFIGS = {
    "vit_tiny_patch16_224": (192, 12, 3),
    "vit_small_patch16_224": (384, 12, 6),
    "vit_base_patch16_224": (768, 12, 12),
}


def register_vit_config(name: str, embed_dim: int, depth: int, heads: int) -> None:
    """Extra (e.g. test-sized) ViT shapes for use_pretrained=False."""
    VIT_CONFIGS[name] = (embed_dim, depth, heads)


_WARNED_TORCH_PATH = False


class Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, Cd = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        return self.proj(x.transpose(1, 2).reshape(B, N, Cd))


class Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    """Pre-norm block: x += attn(norm1(x)); x += mlp(norm2(x)); LayerNorm eps 1e-6 (timm ViT default)."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        """PyTorch-op evaluation of the block: compatibility surface for callers that invoke `model.transformer(x)`
        directly (evaluate_model.py:255, an analysis pass). The model's own forward never comes here (it runs the
        native plan); recording an autograd graph through this path would be a fallback off the CUDA kernels, so it
        warns (once) instead of staying silent."""
        global _WARNED_TORCH_PATH
        if not _WARNED_TORCH_PATH and torch.is_grad_enabled() and (
                x.requires_grad or any(p.requires_grad for p in self.parameters())):
            import warnings
            warnings.warn("fer_vit_b200: Block.forward is a PyTorch-op compatibility path for analysis code; training "
                          "should go through the model's forward (the native sm_100a plan)", stacklevel=2)
            _WARNED_TORCH_PATH = True
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class VisionTransformerShell(nn.Module):
    """The three attributes of a timm ViT that HybridLatentViT reads, randomly initialised with timm's scheme
    (trunc_normal std .02 weights, zero biases, cls ~ N(0, 1e-6), pos_embed trunc_normal std .02)."""

    def __init__(self, name: str):
        super().__init__()
        if name not in VIT_CONFIGS:
            raise ValueError(f"unknown ViT '{name}' (known without timm: {sorted(VIT_CONFIGS)})")
        dim, depth, heads = VIT_CONFIGS[name]
        self.embed_dim = dim
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, 197, dim))
        self.blocks = nn.Sequential(*[Block(dim, heads) for _ in range(depth)])
        nn.init.normal_(self.cls_token, std=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)


def create_vit(name: str, pretrained: bool):
    """timm.create_model(name, pretrained, num_classes=0) when timm is importable, else the shell above
    (random init only: pretrained weights need timm and its download)."""
    try:
        import timm  # type: ignore
    except ImportError:
        timm = None
    if timm is not None and (name not in VIT_CONFIGS or pretrained):
        return timm.create_model(name, pretrained=pretrained, num_classes=0)
    if pretrained:
        raise ImportError("timm is required to load pretrained ViT weights. Install with: pip install timm "
                          "(use_pretrained=False builds the same architecture with random weights)")
    return VisionTransformerShell(name)


def check_block_structure(blocks) -> None:
    """The native plan implements the timm ViT block at create_model defaults: pre-norm LayerNorm,
    fused qkv Linear WITH bias, no q/k norm, no LayerScale, no DropPath / projection dropout, Mlp = fc1 -> GELU -> fc2.
    Any other timm variant (deit3 / LayerScale models, qkv_bias=False, qk_norm=True, SwiGLU MLPs ...) would run through
    the kernels with silently wrong arithmetic, so it is refused here by name."""
    def is_identity(m):
        return m is None or isinstance(m, nn.Identity)
    for i, b in enumerate(blocks):
        problems = []
        for name in ("ls1", "ls2", "drop_path1", "drop_path2", "drop_path"):
            if not is_identity(getattr(b, name, None)):
                problems.append(f"{name} = {type(getattr(b, name)).__name__}")
        attn = getattr(b, "attn", None)
        mlp = getattr(b, "mlp", None)
        if attn is None or mlp is None or not hasattr(b, "norm1") or not hasattr(b, "norm2"):
            problems.append("missing norm1 / attn / norm2 / mlp")
        else:
            for name in ("q_norm", "k_norm"):
                if not is_identity(getattr(attn, name, None)):
                    problems.append(f"attn.{name} = {type(getattr(attn, name)).__name__}")
            if getattr(attn.qkv, "bias", None) is None or getattr(attn.proj, "bias", None) is None:
                problems.append("attn.qkv / attn.proj without bias")
            for name in ("attn_drop", "proj_drop"):
                d = getattr(attn, name, None)
                if d is not None and float(getattr(d, "p", 0.0)) > 0.0:
                    problems.append(f"attn.{name}.p = {d.p}")
            if not (hasattr(mlp, "fc1") and hasattr(mlp, "fc2")) or getattr(mlp.fc1, "bias", None) is None \
                    or getattr(mlp.fc2, "bias", None) is None:
                problems.append("mlp is not fc1 -> act -> fc2 with biases")
            elif not isinstance(getattr(mlp, "act", nn.GELU()), nn.GELU) or getattr(mlp.act, "approximate", "none") != "none":
                problems.append(f"mlp.act = {type(mlp.act).__name__} (exact-erf GELU expected)")
            if not is_identity(getattr(mlp, "norm", None)):
                problems.append("mlp.norm is not Identity")
            for name in ("drop1", "drop2", "drop"):
                d = getattr(mlp, name, None)
                if d is not None and float(getattr(d, "p", 0.0)) > 0.0:
                    problems.append(f"mlp.{name}.p = {d.p}")
            if not isinstance(b.norm1, nn.LayerNorm) or not isinstance(b.norm2, nn.LayerNorm):
                problems.append("norm1 / norm2 are not nn.LayerNorm")
        if problems:
            raise NotImplementedError(f"fer_vit_b200: transformer block {i} is not a default timm ViT block and has no "
                                      f"native kernel path: {'; '.join(problems)}")
