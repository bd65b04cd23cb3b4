"""ImageViT baseline trained from scratch on 224x224 images — drop-in for models_fer_vit/image_vit.py."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from .. import _lib as L
from ..native_module import NativeModule, base_config, encoder_layer_tensors


class PatchEmbedding(nn.Module):
    """Conv2d(in, E, k=s=patch) patchify (image_vit.py:11-44); key ``proj.*``.

    Inside ImageViT the convolution is an im2col pass + tensor-core GEMM whose epilogue adds bias and positions;
    this torch-op forward is only the stand-alone compatibility surface."""

    def __init__(self, img_size: int = 224, patch_size: int = 16, in_channels: int = 3, embed_dim: int = 768):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.n_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.proj(x).flatten(2).transpose(1, 2)


class ImageViT(NativeModule):
    """Patch embed + cls + positions + dropout + post-norm GELU encoder + LayerNorm + Linear (image_vit.py:53-166).

    Keys: ``cls_token, pos_embed, patch_embed.proj.*, transformer.layers.{i}.*, norm.*, head.*``.
    """

    def __init__(self, img_size: int = 224, patch_size: int = 16, in_channels: int = 3, embed_dim: int = 768,
                 depth: int = 12, heads: int = 12, mlp_dim: int = 3072, num_classes: int = 7, dropout: float = 0.1):
        super().__init__()
        self.patch_size = patch_size
        self.n_patches = (img_size // patch_size) ** 2
        self.patch_embed = PatchEmbedding(img_size=img_size, patch_size=patch_size, in_channels=in_channels,
                                          embed_dim=embed_dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, self.n_patches + 1, embed_dim))
        self.dropout = nn.Dropout(dropout)
        layer = nn.TransformerEncoderLayer(d_model=embed_dim, nhead=heads, dim_feedforward=mlp_dim, dropout=dropout,
                                           activation="gelu", batch_first=True, norm_first=False)
        self.transformer = nn.TransformerEncoder(layer, num_layers=depth)
        self.norm = nn.LayerNorm(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self._dims = dict(img_size=img_size, in_channels=in_channels, embed_dim=embed_dim, depth=depth, heads=heads,
                          mlp_dim=mlp_dim, num_classes=num_classes, dropout=float(dropout))
        self._init_weights()

    def _init_weights(self) -> None:
        """trunc_normal(0.02) for positions, cls and nn.Linear weights; zero biases; unit LayerNorm
        (image_vit.py:122-136). in_proj_weight is a bare Parameter, so it keeps torch's xavier init."""
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def _plan_config(self) -> L.Config:
        d = self._dims
        c = base_config()
        c.input_kind = 1
        c.L = self.n_patches
        c.Din = d["in_channels"] * self.patch_size * self.patch_size
        c.E, c.depth, c.H, c.F, c.C = d["embed_dim"], d["depth"], d["heads"], d["mlp_dim"], d["num_classes"]
        c.norm_first, c.act = 0, L.ACT_GELU
        c.eps_block = self.transformer.layers[0].norm1.eps
        c.eps_head = self.norm.eps
        c.dropout = d["dropout"]
        c.input_dropout = 1
        c.img_c, c.img_h, c.img_w, c.patch = d["in_channels"], d["img_size"], d["img_size"], self.patch_size
        return c

    def _plan_tensors(self) -> Dict[int, torch.Tensor]:
        t = {
            L.G_IN_W: self.patch_embed.proj.weight, L.G_IN_B: self.patch_embed.proj.bias,
            L.G_CLS: self.cls_token, L.G_POS: self.pos_embed,
            L.G_HEAD_LN_W: self.norm.weight, L.G_HEAD_LN_B: self.norm.bias,
            L.G_HEAD_W: self.head.weight, L.G_HEAD_B: self.head.bias,
        }
        for i, layer in enumerate(self.transformer.layers):
            t.update(encoder_layer_tensors(layer, i))
        return t

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (B, C, H, W) fp32 on CUDA -> logits (B, num_classes)."""
        return self._native_forward(x)


def create_vit_small(num_classes: int = 7, img_size: int = 224) -> ImageViT:
    """ViT-Small/16 shape (image_vit.py:169-179)."""
    return ImageViT(img_size=img_size, patch_size=16, embed_dim=384, depth=12, heads=6, mlp_dim=1536,
                    num_classes=num_classes)


def create_vit_base(num_classes: int = 7, img_size: int = 224) -> ImageViT:
    """ViT-Base/16 shape (image_vit.py:182-192)."""
    return ImageViT(img_size=img_size, patch_size=16, embed_dim=768, depth=12, heads=12, mlp_dim=3072,
                    num_classes=num_classes)


def create_vit_tiny(num_classes: int = 7, img_size: int = 224) -> ImageViT:
    """ViT-Tiny/16 shape (image_vit.py:195-205)."""
    return ImageViT(img_size=img_size, patch_size=16, embed_dim=192, depth=12, heads=3, mlp_dim=768,
                    num_classes=num_classes)
