"""w+ pre-modules (drop-in for the reference's top-level ``modules`` package)."""
from .leam import LEAM
from .semantic_pe import SemanticPE
from .layer_wise_norm import LayerWiseNorm

__all__ = ["LEAM", "SemanticPE", "LayerWiseNorm"]
