"""LEAM — layer-wise expression attention mask over the 18 w+ layers (reference: modules/leam.py:6-44)."""
import torch
import torch.nn as nn

from ._fused import premodules


class LEAM(nn.Module):
    """y[b, l, :] = x[b, l, :] * sigmoid(layer_weights[l]).

    Same constructor and state_dict key (``layer_weights``) as the reference (modules/leam.py:17-29); the product is
    computed by the fused native pre-module kernel.
    """

    def __init__(self, num_layers: int = 18, init_coarse: float = 0.5, init_fine: float = 0.5):
        super().__init__()
        w = torch.ones(num_layers)
        w[:4] = init_coarse    # coarse layers 1-4
        w[12:] = init_fine     # fine layers 13-18
        self.layer_weights = nn.Parameter(w)

    def forward(self, w_plus: torch.Tensor) -> torch.Tensor:
        return premodules(w_plus, leam_w=self.layer_weights)

    def get_weights(self) -> torch.Tensor:
        """sigmoid(layer_weights), detached on the CPU, for plotting (reference: modules/leam.py:42-44)."""
        return torch.sigmoid(self.layer_weights).detach().cpu()
