"""Drop-in ``ExpressionAwareViT`` (reference models_fer_vit/expression_aware_vit.py:24-134): a fixed
``LatentDecomposer`` in front of a ``HybridLatentViT``; only the ViT side trains. Both halves run on the native
kernels: one ``fervit_latent_decompose`` launch, then the model plan."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .hybrid_latent_vit import HybridLatentViT, create_hybrid_latent_vit
from .latent_decomposer import LatentDecomposer

_OUTPUT_MODES = ("expr_only", "id_only", "enhanced", "concat")
_DECOMPOSE_MODES = ("all_classes", "max_class")


def _numel(params) -> int:
    return sum(p.numel() for p in params)


class ExpressionAwareViT(nn.Module):
    """Attributes as the reference: ``decomposer``, ``vit``, ``output_mode``, ``enhance_alpha``, ``decompose_mode``."""

    def __init__(self, decomposer: LatentDecomposer, vit_model: HybridLatentViT, output_mode: str = "expr_only",
                 enhance_alpha: float = 2.0, decompose_mode: str = "all_classes"):
        super().__init__()
        self.decomposer, self.vit = decomposer, vit_model
        self.output_mode, self.enhance_alpha, self.decompose_mode = output_mode, enhance_alpha, decompose_mode
        lines = ["", "[ExpressionAwareViT]", f"  decompose_mode : {decompose_mode}", f"  output_mode    : {output_mode}"]
        if output_mode == "enhanced":
            lines.append(f"  enhance_alpha  : {enhance_alpha}")
        print("\n".join(lines))

    @classmethod
    def from_config(cls, directions_path: str, model_size: str = "small", num_classes: int = 7,
                    use_pretrained: bool = True, freeze_transformer: bool = False,
                    freeze_stages: Optional[int] = None, use_adapter: bool = False, adapter_dim: int = 64,
                    output_mode: str = "expr_only", enhance_alpha: float = 2.0,
                    decompose_mode: str = "all_classes") -> "ExpressionAwareViT":
        """Factory of expression_aware_vit.py:56-107: directions file + HybridLatentViT options -> model."""
        front = LatentDecomposer.from_file(directions_path)
        # 'concat' lays the expression and identity parts side by side: the ViT sees twice the tokens (:89)
        tokens = front.seq_len * (2 if output_mode == "concat" else 1)
        vit_options = dict(model_size=model_size, num_classes=num_classes, use_pretrained=use_pretrained,
                           freeze_transformer=freeze_transformer, freeze_stages=freeze_stages,
                           use_adapter=use_adapter, adapter_dim=adapter_dim)
        backbone = create_hybrid_latent_vit(latent_dim=front.latent_dim, seq_len=tokens, **vit_options)
        return cls(front, backbone, output_mode, enhance_alpha, decompose_mode)

    def forward(self, w_plus: torch.Tensor) -> torch.Tensor:
        """w+ ``(B, 18, 512)`` -> logits ``(B, num_classes)`` (expression_aware_vit.py:109-122)."""
        tokens = self.decomposer(w_plus, self.output_mode, self.enhance_alpha, self.decompose_mode)
        return self.vit(tokens)

    def forward_with_scores(self, w_plus: torch.Tensor):
        """(logits, expression scores ``(B, C)``): the coefficients come out of the same decomposer launch that
        produces the ViT input, so monitoring them (the SVM decision values of the directions) costs nothing extra."""
        tokens, scores = self.decomposer._run(w_plus, self.decompose_mode, self.output_mode, self.enhance_alpha,
                                              scores=True)
        return self.vit(tokens), scores

    def extra_repr(self) -> str:
        return (f"output_mode={self.output_mode!r}, decompose_mode={self.decompose_mode!r}, "
                f"enhance_alpha={self.enhance_alpha}, tokens={self.vit.seq_len}")

    def get_trainable_params(self) -> List[nn.Parameter]:
        """The ViT side's trainable parameters; the decomposer holds buffers only (:124-126)."""
        return [p for p in self.vit.parameters() if p.requires_grad]

    def print_info(self) -> None:
        total, trainable = _numel(self.parameters()), _numel(self.get_trainable_params())
        print("\n[ExpressionAwareViT] Parameters:\n"
              f"  Total      : {total:,}\n"
              f"  Trainable  : {trainable:,} ({trainable / total * 100:.1f}%)\n"
              "  Decomposer : fixed (SVM directions, not trained)")
