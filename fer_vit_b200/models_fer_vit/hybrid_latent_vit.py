"""HybridLatentViT: (pretrained) ViT encoder blocks on w+ tokens, optional bottleneck adapters — drop-in for
models_fer_vit/hybrid_latent_vit.py of the reference."""
from __future__ import annotations

from typing import Dict, Literal, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib as L
from ..native_module import NativeModule, base_config
from .vit_blocks import check_block_structure, create_vit


class AdapterModule(nn.Module):
    """x + alpha * W2 GELU(W1 x + b1) + b2 after every block (hybrid_latent_vit.py:249-265).

    Keys ``adapter.0.*``, ``adapter.2.*``, ``alpha``. Inside HybridLatentViT the adapter runs as two tensor-core GEMMs
    with fused epilogues; this torch-op forward is only the stand-alone compatibility surface.
    """

    def __init__(self, embed_dim: int, adapter_dim: int):
        super().__init__()
        self.adapter = nn.Sequential(nn.Linear(embed_dim, adapter_dim), nn.GELU(), nn.Linear(adapter_dim, embed_dim))
        self.alpha = nn.Parameter(torch.ones(1) * 0.1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x + self.alpha * self.adapter(x)


class HybridLatentViT(NativeModule):
    """latent (B, 18, 512) -> Linear -> [cls; tokens] + pos -> ViT blocks (+ adapters) -> LN/Dropout/Linear head.

    Same constructor, attributes (``input_proj``, ``cls_token``, ``pos_embed``, ``transformer``, ``adapters``,
    ``head``, ``use_adapter``, ``embed_dim`` ...) and state_dict keys as hybrid_latent_vit.py:30-116.
    """

    def __init__(self, latent_dim: int = 512, seq_len: int = 18,
                 pretrained_model_name: str = "vit_small_patch16_224", num_classes: int = 7,
                 use_pretrained: bool = True, freeze_transformer: bool = False,
                 freeze_stages: Optional[int] = None, adapter_dim: Optional[int] = None, verbose: bool = True):
        super().__init__()
        self.latent_dim = latent_dim
        self.seq_len = seq_len
        self.num_classes = num_classes
        self.pretrained_model_name = pretrained_model_name
        self.use_adapter = adapter_dim is not None
        self._verbose = verbose

        vit = create_vit(pretrained_model_name, use_pretrained)
        self.embed_dim = vit.embed_dim
        self.input_proj = nn.Linear(latent_dim, self.embed_dim)
        if hasattr(vit, "cls_token"):
            self.cls_token = nn.Parameter(vit.cls_token.data.clone())
        else:
            self.cls_token = nn.Parameter(torch.randn(1, 1, self.embed_dim))
        self.pos_embed = self._init_position_embedding(vit, seq_len)
        self.transformer = self._extract_transformer(vit)
        if self.use_adapter:
            self.adapters = nn.ModuleList([AdapterModule(self.embed_dim, adapter_dim)
                                           for _ in range(len(self.transformer))])
        if freeze_transformer:
            self._freeze_transformer()
        elif freeze_stages is not None:
            self._freeze_stages(freeze_stages)
        self.head = nn.Sequential(nn.LayerNorm(self.embed_dim), nn.Dropout(0.1),
                                  nn.Linear(self.embed_dim, num_classes))
        self._adapter_dim = adapter_dim or 0
        self._print_model_info()

    # ---- construction helpers (hybrid_latent_vit.py:118-203) ----------------------------------
    def _init_position_embedding(self, vit, seq_len: int) -> nn.Parameter:
        if not hasattr(vit, "pos_embed"):
            return nn.Parameter(torch.randn(1, seq_len + 1, self.embed_dim))
        pos = vit.pos_embed                                    # (1, N+1, E)
        if seq_len == pos.size(1) - 1:
            return nn.Parameter(pos.data.clone())
        # cls row kept; the N patch rows are linearly resampled along the sequence axis to seq_len rows
        patch = F.interpolate(pos[:, 1:, :].permute(0, 2, 1), size=seq_len, mode="linear", align_corners=False)
        return nn.Parameter(torch.cat([pos[:, 0:1, :], patch.permute(0, 2, 1)], dim=1).detach().clone())

    def _extract_transformer(self, vit):
        if not hasattr(vit, "blocks"):
            raise AttributeError(f"Cannot extract transformer blocks from {self.pretrained_model_name}. "
                                 f"Model structure may be different.")
        return vit.blocks

    def _freeze_transformer(self) -> None:
        for p in self.transformer.parameters():
            p.requires_grad = False

    def _freeze_stages(self, n_stages: int) -> None:
        for i in range(min(n_stages, len(self.transformer))):
            for p in self.transformer[i].parameters():
                p.requires_grad = False

    def _print_model_info(self) -> None:
        if not self._verbose:
            return
        total = sum(p.numel() for p in self.parameters())
        train = sum(p.numel() for p in self.parameters() if p.requires_grad)
        print(f"HybridLatentViT[{self.pretrained_model_name}] tokens=({self.seq_len},{self.latent_dim}) "
              f"E={self.embed_dim} blocks={len(self.transformer)} classes={self.num_classes} "
              f"adapter={'yes' if self.use_adapter else 'no'} params total={total:,} trainable={train:,} "
              f"({100.0 * train / total:.1f}%) frozen={total - train:,}")

    def unfreeze_all(self) -> None:
        """Make every parameter trainable again (hybrid_latent_vit.py:241-246)."""
        for p in self.parameters():
            p.requires_grad = True
        self._print_model_info()

    # ---- native plan description ---------------------------------------------------------------
    def _plan_config(self) -> L.Config:
        check_block_structure(self.transformer)
        blk0 = self.transformer[0]
        c = base_config()
        c.input_kind, c.L, c.Din, c.E, c.depth = 0, self.seq_len, self.latent_dim, self.embed_dim, len(self.transformer)
        c.H = blk0.attn.num_heads
        c.F = blk0.mlp.fc1.out_features
        c.C = self.num_classes
        c.norm_first, c.act = 1, L.ACT_GELU
        c.eps_block = blk0.norm1.eps
        c.eps_head = self.head[0].eps
        c.adapter_dim = self._adapter_dim
        c.dropout = 0.0
        c.head_dropout = float(self.head[1].p)
        return c

    def _plan_tensors(self) -> Dict[int, torch.Tensor]:
        t = {
            L.G_IN_W: self.input_proj.weight, L.G_IN_B: self.input_proj.bias,
            L.G_CLS: self.cls_token, L.G_POS: self.pos_embed,
            L.G_HEAD_LN_W: self.head[0].weight, L.G_HEAD_LN_B: self.head[0].bias,
            L.G_HEAD_W: self.head[2].weight, L.G_HEAD_B: self.head[2].bias,
        }
        for i, b in enumerate(self.transformer):
            t[L.bslot(i, L.B_LN1_W)], t[L.bslot(i, L.B_LN1_B)] = b.norm1.weight, b.norm1.bias
            t[L.bslot(i, L.B_QKV_W)], t[L.bslot(i, L.B_QKV_B)] = b.attn.qkv.weight, b.attn.qkv.bias
            t[L.bslot(i, L.B_PROJ_W)], t[L.bslot(i, L.B_PROJ_B)] = b.attn.proj.weight, b.attn.proj.bias
            t[L.bslot(i, L.B_LN2_W)], t[L.bslot(i, L.B_LN2_B)] = b.norm2.weight, b.norm2.bias
            t[L.bslot(i, L.B_FC1_W)], t[L.bslot(i, L.B_FC1_B)] = b.mlp.fc1.weight, b.mlp.fc1.bias
            t[L.bslot(i, L.B_FC2_W)], t[L.bslot(i, L.B_FC2_B)] = b.mlp.fc2.weight, b.mlp.fc2.bias
            if self.use_adapter:
                a = self.adapters[i]
                t[L.bslot(i, L.B_AD1_W)], t[L.bslot(i, L.B_AD1_B)] = a.adapter[0].weight, a.adapter[0].bias
                t[L.bslot(i, L.B_AD2_W)], t[L.bslot(i, L.B_AD2_B)] = a.adapter[2].weight, a.adapter[2].bias
                t[L.bslot(i, L.B_ALPHA)] = a.alpha
        return t

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (B, seq_len, latent_dim) fp32 on CUDA -> logits (B, num_classes) (hybrid_latent_vit.py:205-239)."""
        return self._native_forward(x)


def create_hybrid_latent_vit(latent_dim: int = 512, seq_len: int = 18,
                             model_size: Literal["tiny", "small", "base"] = "small", num_classes: int = 7,
                             use_pretrained: bool = True, freeze_transformer: bool = False,
                             freeze_stages: Optional[int] = None, use_adapter: bool = False,
                             adapter_dim: int = 64) -> HybridLatentViT:
    """Factory with the reference's keyword set (hybrid_latent_vit.py:268-310)."""
    names = {"tiny": "vit_tiny_patch16_224", "small": "vit_small_patch16_224", "base": "vit_base_patch16_224"}
    return HybridLatentViT(latent_dim=latent_dim, seq_len=seq_len,
                           pretrained_model_name=names.get(model_size, "vit_small_patch16_224"),
                           num_classes=num_classes, use_pretrained=use_pretrained,
                           freeze_transformer=freeze_transformer, freeze_stages=freeze_stages,
                           adapter_dim=adapter_dim if use_adapter else None)


# Training strategies the reference recommends (hybrid_latent_vit.py:314-343): same keys and flag values.
RECOMMENDED_STRATEGIES = {
    "full_finetune": {"freeze_transformer": False, "freeze_stages": None, "use_adapter": False, "lr": 1e-4,
                      "description": "train every parameter (best accuracy, slowest)"},
    "partial_freeze": {"freeze_transformer": False, "freeze_stages": 6, "use_adapter": False, "lr": 3e-4,
                       "description": "freeze the lower six blocks (balanced)"},
    "adapter": {"freeze_transformer": True, "freeze_stages": None, "use_adapter": True, "lr": 1e-3,
                "description": "train adapters only (fastest, memory-efficient)"},
    "linear_probe": {"freeze_transformer": True, "freeze_stages": None, "use_adapter": False, "lr": 1e-3,
                     "description": "train the classification head only (baseline)"},
}
