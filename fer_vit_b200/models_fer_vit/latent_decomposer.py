"""Drop-in ``LatentDecomposer`` (reference models_fer_vit/latent_decomposer.py:31-173): same constructor, buffer
(``directions`` [C, 18, 512], re-normalised), methods and modes; the arithmetic is one launch of
``fervit_latent_decompose`` (include/fervit_b200.h) instead of two skinny matmuls plus elementwise passes."""
from __future__ import annotations

from typing import Dict, Literal

import torch
import torch.nn as nn

from .. import _lib as L

EMOTION_NAMES = {0: 'angry', 1: 'disgust', 2: 'fear', 3: 'happy', 4: 'neutral', 5: 'sad', 6: 'surprise'}

_DECOMPOSE = {"all_classes": 0, "max_class": 1}
_OUTPUT = {"expr_only": 0, "id_only": 1, "enhanced": 2, "concat": 3}


class LatentDecomposer(nn.Module):
    def __init__(self, directions: Dict[int, torch.Tensor], seq_len: int = 18, latent_dim: int = 512):
        super().__init__()
        self.seq_len = seq_len
        self.latent_dim = latent_dim
        self.num_classes = len(directions)
        # stack to (C, 18, 512), re-normalise each flattened direction (latent_decomposer.py:58-65)
        dirs = torch.stack([directions[i] for i in range(self.num_classes)], dim=0)
        flat = dirs.reshape(self.num_classes, -1).float()
        flat = flat / (flat.norm(dim=1, keepdim=True) + 1e-12)
        self.register_buffer('directions', flat.view(self.num_classes, seq_len, latent_dim).contiguous())

    @classmethod
    def from_file(cls, path: str) -> 'LatentDecomposer':
        data = torch.load(path, map_location='cpu', weights_only=False)
        directions = data['directions']
        seq_len = data.get('seq_len', 18)
        latent_dim = data.get('latent_dim', 512)
        method = data.get('method', 'unknown')
        print(f"Loaded '{method}' expression directions: {path}")
        print(f"  Classes  : {list(directions.keys())}")
        print(f"  Direction shape: ({seq_len}, {latent_dim}) x {len(directions)} classes")
        return cls(directions, seq_len, latent_dim)

    # ------------------------------------------------------------------ native call
    def _run(self, w_plus: torch.Tensor, decompose_mode: str, output_mode, alpha: float = 1.0, scores: bool = False):
        if decompose_mode not in _DECOMPOSE:
            raise ValueError(f"Unknown mode: {decompose_mode!r}")
        if output_mode is not None and output_mode not in _OUTPUT:
            raise ValueError(f"Unknown output_mode: {output_mode!r}")
        if not w_plus.is_cuda or not self.directions.is_cuda:
            raise RuntimeError("fer_vit_b200: LatentDecomposer runs on CUDA tensors only (no CPU fallback) - move the "
                               "module and its input to a CUDA device")
        if w_plus.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("fer_vit_b200: LatentDecomposer has no backward (the reference feeds it data latents, "
                               "which carry no gradient)")
        B = w_plus.size(0)
        row = self.seq_len * self.latent_dim
        if w_plus.numel() != B * row:
            raise RuntimeError(f"fer_vit_b200: expected w+ of shape (B, {self.seq_len}, {self.latent_dim}), got "
                               f"{tuple(w_plus.shape)}")
        w = w_plus.detach().contiguous().float()
        out = None
        if output_mode is not None:
            rows = 2 * self.seq_len if output_mode == "concat" else self.seq_len
            out = torch.empty(B, rows, self.latent_dim, dtype=torch.float32, device=w.device)
        sc = torch.empty(B, self.num_classes, dtype=torch.float32, device=w.device) if scores else None
        L.check(L.lib().fervit_latent_decompose(
            w.data_ptr(), self.directions.data_ptr(), B, self.num_classes, row, _DECOMPOSE[decompose_mode],
            _OUTPUT[output_mode] if output_mode is not None else 0, float(alpha),
            out.data_ptr() if out is not None else None, sc.data_ptr() if sc is not None else None,
            torch.cuda.current_stream().cuda_stream))
        return out, sc

    # ------------------------------------------------------------------ reference API
    def decompose(self, w_plus: torch.Tensor, mode: Literal['all_classes', 'max_class'] = 'all_classes'):
        """(w_expr, w_id), each (B, 18, 512) (latent_decomposer.py:82-118)."""
        both, _ = self._run(w_plus, mode, "concat")
        return both[:, :self.seq_len], both[:, self.seq_len:]

    def get_expression_scores(self, w_plus: torch.Tensor) -> torch.Tensor:
        """(B, num_classes) projection coefficients (latent_decomposer.py:120-130)."""
        return self._run(w_plus, "all_classes", None, scores=True)[1]

    def enhance_expression(self, w_plus: torch.Tensor, alpha: float = 2.0,
                           mode: Literal['all_classes', 'max_class'] = 'all_classes') -> torch.Tensor:
        """w_id + alpha * w_expr (latent_decomposer.py:132-144)."""
        return self._run(w_plus, mode, "enhanced", alpha)[0]

    def forward(self, w_plus: torch.Tensor,
                output_mode: Literal['expr_only', 'id_only', 'enhanced', 'concat'] = 'expr_only',
                enhance_alpha: float = 2.0,
                decompose_mode: Literal['all_classes', 'max_class'] = 'all_classes') -> torch.Tensor:
        """The ViT input (latent_decomposer.py:146-173): (B, 18, 512), or (B, 36, 512) for 'concat'."""
        return self._run(w_plus, decompose_mode, output_mode, enhance_alpha)[0]
