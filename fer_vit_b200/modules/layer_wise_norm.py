"""LayerWiseNorm — one LayerNorm per w+ layer, optional sigmoid-gated residual (modules/layer_wise_norm.py)."""
import torch
import torch.nn as nn

from ._fused import premodules


class LayerWiseNorm(nn.Module):
    """n = LN_l(x[:, l]); y = n, or x + sigmoid(gate[l]) * (n - x) with use_residual (gate initialised to -5).

    Keys ``norms.{l}.weight|bias`` and ``gate`` as in the reference (modules/layer_wise_norm.py:25-33).
    """

    def __init__(self, num_layers: int = 18, d_model: int = 512, use_residual: bool = False):
        super().__init__()
        self.norms = nn.ModuleList([nn.LayerNorm(d_model) for _ in range(num_layers)])
        self.use_residual = use_residual
        if use_residual:
            self.gate = nn.Parameter(torch.full((num_layers,), -5.0))

    def stacked(self):
        """([L, D] gamma, [L, D] beta) views for the fused kernel; autograd splits their gradients back per layer."""
        gamma = torch.stack([n.weight for n in self.norms])
        beta = torch.stack([n.bias for n in self.norms])
        return gamma, beta

    def forward(self, w_plus: torch.Tensor) -> torch.Tensor:
        gamma, beta = self.stacked()
        return premodules(w_plus, gamma=gamma, beta=beta, gate=self.gate if self.use_residual else None,
                          eps=self.norms[0].eps)
