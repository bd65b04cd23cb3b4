"""Drop-in ``ExpressionAwareViT`` (reference models_fer_vit/expression_aware_vit.py:24-134): a fixed
``LatentDecomposer`` in front of a ``HybridLatentViT``; only the ViT side trains. Both halves run on the native
kernels (one decomposer launch, then the model plan)."""
from __future__ import annotations

from typing import Literal, Optional

import torch
import torch.nn as nn

from .latent_decomposer import LatentDecomposer
from .hybrid_latent_vit import HybridLatentViT, create_hybrid_latent_vit


class ExpressionAwareViT(nn.Module):
    def __init__(self, decomposer: LatentDecomposer, vit_model: HybridLatentViT,
                 output_mode: Literal['expr_only', 'id_only', 'enhanced', 'concat'] = 'expr_only',
                 enhance_alpha: float = 2.0,
                 decompose_mode: Literal['all_classes', 'max_class'] = 'all_classes'):
        super().__init__()
        self.decomposer = decomposer
        self.vit = vit_model
        self.output_mode = output_mode
        self.enhance_alpha = enhance_alpha
        self.decompose_mode = decompose_mode
        print(f"\n[ExpressionAwareViT]")
        print(f"  decompose_mode : {decompose_mode}")
        print(f"  output_mode    : {output_mode}")
        if output_mode == 'enhanced':
            print(f"  enhance_alpha  : {enhance_alpha}")

    @classmethod
    def from_config(cls, directions_path: str, model_size: str = 'small', num_classes: int = 7,
                    use_pretrained: bool = True, freeze_transformer: bool = False,
                    freeze_stages: Optional[int] = None, use_adapter: bool = False, adapter_dim: int = 64,
                    output_mode: Literal['expr_only', 'id_only', 'enhanced', 'concat'] = 'expr_only',
                    enhance_alpha: float = 2.0,
                    decompose_mode: Literal['all_classes', 'max_class'] = 'all_classes') -> 'ExpressionAwareViT':
        decomposer = LatentDecomposer.from_file(directions_path)
        # 'concat' feeds expression and identity parts side by side: twice the sequence (expression_aware_vit.py:89)
        seq_len = decomposer.seq_len * (2 if output_mode == 'concat' else 1)
        vit = create_hybrid_latent_vit(latent_dim=decomposer.latent_dim, seq_len=seq_len, model_size=model_size,
                                       num_classes=num_classes, use_pretrained=use_pretrained,
                                       freeze_transformer=freeze_transformer, freeze_stages=freeze_stages,
                                       use_adapter=use_adapter, adapter_dim=adapter_dim)
        return cls(decomposer=decomposer, vit_model=vit, output_mode=output_mode, enhance_alpha=enhance_alpha,
                   decompose_mode=decompose_mode)

    def forward(self, w_plus: torch.Tensor) -> torch.Tensor:
        x = self.decomposer(w_plus, output_mode=self.output_mode, enhance_alpha=self.enhance_alpha,
                            decompose_mode=self.decompose_mode)
        return self.vit(x)

    def get_trainable_params(self):
        return [p for p in self.vit.parameters() if p.requires_grad]

    def print_info(self):
        total = sum(p.numel() for p in self.parameters())
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        print(f"\n[ExpressionAwareViT] Parameters:")
        print(f"  Total      : {total:,}")
        print(f"  Trainable  : {trainable:,} ({trainable/total*100:.1f}%)")
        print(f"  Decomposer : fixed (SVM directions, not trained)")
