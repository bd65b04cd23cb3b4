"""LatentViT v2: SemanticPE -> LayerWiseNorm -> LEAM -> LatentViT (drop-in for models_fer_vit/latent_vit_v2.py)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import _lib as L
from ..modules import LEAM, LayerWiseNorm, SemanticPE
from ..native_module import NativeModule
from .latent_vit import LatentViT


class LatentViTv2(NativeModule):
    """LatentViT with optional w+ pre-modules, applied in the order SPE -> LWN -> LEAM (latent_vit_v2.py:82-84).

    Keys: ``spe.*``, ``lwn.*``, ``leam.*``, ``backbone.*`` (latent_vit_v2.py:53-67). The whole model — pre-modules
    fused into one kernel that also writes the bf16 A operand of the token projection — runs as one native plan.
    """

    def __init__(self, latent_dim: int = 512, seq_len: int = 18, embed_dim: int = 512, depth: int = 6,
                 heads: int = 8, mlp_dim: int = 2048, num_classes: int = 7, dropout: float = 0.1,
                 use_lwn: bool = False, use_lwn_residual: bool = False, use_spe: bool = False,
                 use_leam: bool = False):
        super().__init__()
        self.lwn = LayerWiseNorm(seq_len, latent_dim, use_residual=use_lwn_residual) if use_lwn else nn.Identity()
        self.spe = SemanticPE(latent_dim, seq_len) if use_spe else nn.Identity()
        self.leam = LEAM(seq_len) if use_leam else nn.Identity()
        self.backbone = LatentViT(latent_dim=latent_dim, seq_len=seq_len, embed_dim=embed_dim, depth=depth,
                                  heads=heads, mlp_dim=mlp_dim, num_classes=num_classes, dropout=dropout)
        self.use_lwn = use_lwn
        self.use_lwn_residual = use_lwn_residual
        self.use_spe = use_spe
        self.use_leam = use_leam

    def _plan_config(self) -> L.Config:
        c = self.backbone._plan_config(use_spe=self.use_spe, use_lwn=self.use_lwn,
                                       use_lwn_res=self.use_lwn and self.use_lwn_residual, use_leam=self.use_leam)
        if self.use_lwn:
            c.eps_lwn = self.lwn.norms[0].eps
        return c

    def _plan_tensors(self) -> Dict[int, torch.Tensor]:
        t = self.backbone._plan_tensors()
        if self.use_spe:
            t[L.G_SPE_GROUP] = self.spe.group_embed.weight
            t[L.G_SPE_LAYER] = self.spe.layer_embed.weight
            t[L.G_SPE_GROUPS] = self.spe.groups
        if self.use_lwn:
            t[L.G_LWN_GAMMA], t[L.G_LWN_BETA] = self.lwn.stacked()
            if self.use_lwn_residual:
                t[L.G_LWN_GATE] = self.lwn.gate
        if self.use_leam:
            t[L.G_LEAM_W] = self.leam.layer_weights
        return t

    def forward(self, w_plus: torch.Tensor) -> torch.Tensor:
        """w_plus: (B, seq_len, latent_dim) -> logits (B, num_classes)."""
        return self._native_forward(w_plus)

    def get_leam_weights(self) -> Optional[torch.Tensor]:
        """sigmoid LEAM weights for plotting, or None when use_leam is off (latent_vit_v2.py:87-91)."""
        return self.leam.get_weights() if self.use_leam else None

    def get_config(self) -> dict:
        """Model flags for experiment logs (latent_vit_v2.py:93-101)."""
        return {"model": "LatentViTv2", "use_lwn": self.use_lwn, "use_lwn_residual": self.use_lwn_residual,
                "use_spe": self.use_spe, "use_leam": self.use_leam}
