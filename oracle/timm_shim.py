"""ORACLE — test infrastructure, not product code.

Stand-in for `timm` (timm 1.0.17 is pinned by the reference, environment.yml:124, but is not installed in this image
and cannot be: no network). It exposes only what models_fer_vit/hybrid_latent_vit.py touches:
`timm.create_model(name, pretrained, num_classes)` -> object with `.embed_dim`, `.cls_token`, `.pos_embed`, `.blocks`
(hybrid_latent_vit.py:68-72, 75, 82-83, 120-152, 160-164), so that the reference file runs UNMODIFIED when
`install()` has put this module into sys.modules['timm'] before it is imported.

The block restates timm.models.vision_transformer.{Block, Attention, Mlp} at create_model defaults from the published
algorithm (pre-norm, LayerNorm eps 1e-6, fused qkv Linear with bias, scale hd^-0.5, exact-erf GELU, no dropout /
layer-scale / drop-path). "parity unpinned against timm itself" — see DESIGN.md.
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn

_CONFIGS = {
    "vit_tiny_patch16_224": (192, 12, 3),
    "vit_small_patch16_224": (384, 12, 6),
    "vit_base_patch16_224": (768, 12, 12),
    # test-sized shapes for golden fixtures
    "vit_test_patch16_224": (64, 2, 2),
    "vit_test4_patch16_224": (128, 4, 2),
}


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, 4 * dim)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class VisionTransformer(nn.Module):
    def __init__(self, dim, depth, heads):
        super().__init__()
        self.embed_dim = dim
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.randn(1, 197, dim) * 0.02)
        self.blocks = nn.Sequential(*[Block(dim, heads) for _ in range(depth)])
        nn.init.normal_(self.cls_token, std=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)


def create_model(name, pretrained=False, num_classes=0, **kwargs):
    if pretrained:
        raise RuntimeError("timm shim: pretrained weights are not available offline")
    dim, depth, heads = _CONFIGS[name]
    return VisionTransformer(dim, depth, heads)


def install() -> None:
    """Register this module as `timm` (only if the real one is absent)."""
    try:
        import timm  # noqa: F401
        return
    except ImportError:
        pass
    mod = types.ModuleType("timm")
    mod.create_model = create_model
    mod.__version__ = "0.0-shim"
    sys.modules["timm"] = mod
