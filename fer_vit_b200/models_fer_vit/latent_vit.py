"""LatentViT on StyleGAN2 w+ tokens — drop-in for models_fer_vit/latent_vit.py of the reference."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from .. import _lib as L
from ..native_module import NativeModule, base_config, encoder_layer_tensors


class LatentViT(NativeModule):
    """Linear(latent_dim -> embed_dim) + cls + learned positions + post-norm ReLU encoder + LayerNorm/Linear head.

    Constructor, attributes (``input_proj``, ``cls_token``, ``pos_emb``, ``transformer``, ``mlp_head``, ``seq_len``)
    and state_dict keys follow models_fer_vit/latent_vit.py:6-36. The torch sub-modules are parameter containers
    (so initialisation and checkpoints are interchangeable with the reference); ``forward`` (latent_vit.py:38-48)
    runs the native plan instead of calling them.
    """

    def __init__(self, latent_dim: int = 512, seq_len: int = 18, embed_dim: int = 512, depth: int = 6,
                 heads: int = 8, mlp_dim: int = 2048, num_classes: int = 7, dropout: float = 0.1) -> None:
        super().__init__()
        self.seq_len = seq_len
        self.input_proj = nn.Linear(latent_dim, embed_dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.pos_emb = nn.Parameter(torch.randn(1, seq_len + 1, embed_dim))
        layer = nn.TransformerEncoderLayer(d_model=embed_dim, nhead=heads, dim_feedforward=mlp_dim,
                                           dropout=dropout, batch_first=True)
        self.transformer = nn.TransformerEncoder(layer, num_layers=depth)
        self.mlp_head = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, num_classes))
        self._dims = dict(latent_dim=latent_dim, embed_dim=embed_dim, depth=depth, heads=heads, mlp_dim=mlp_dim,
                          num_classes=num_classes, dropout=float(dropout))

    # ---- native plan description -------------------------------------------------------------
    def _plan_config(self, **pre) -> L.Config:
        d = self._dims
        c = base_config()
        c.input_kind, c.L, c.Din, c.E, c.depth, c.H, c.F, c.C = (0, self.seq_len, d["latent_dim"], d["embed_dim"],
                                                                 d["depth"], d["heads"], d["mlp_dim"],
                                                                 d["num_classes"])
        c.norm_first, c.act = 0, L.ACT_RELU           # nn.TransformerEncoderLayer defaults
        c.eps_block = self.transformer.layers[0].norm1.eps
        c.eps_head = self.mlp_head[0].eps
        c.dropout = d["dropout"]
        for k, v in pre.items():
            setattr(c, k, int(v))
        return c

    def _plan_tensors(self) -> Dict[int, torch.Tensor]:
        t = {
            L.G_IN_W: self.input_proj.weight, L.G_IN_B: self.input_proj.bias,
            L.G_CLS: self.cls_token, L.G_POS: self.pos_emb,
            L.G_HEAD_LN_W: self.mlp_head[0].weight, L.G_HEAD_LN_B: self.mlp_head[0].bias,
            L.G_HEAD_W: self.mlp_head[1].weight, L.G_HEAD_B: self.mlp_head[1].bias,
        }
        for i, layer in enumerate(self.transformer.layers):
            t.update(encoder_layer_tensors(layer, i))
        return t

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (B, seq_len, latent_dim) fp32 on CUDA -> logits (B, num_classes)."""
        return self._native_forward(x)
