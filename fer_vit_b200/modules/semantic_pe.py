"""SemanticPE — coarse/medium/fine group embedding + per-layer embedding added to w+ (modules/semantic_pe.py)."""
import torch
import torch.nn as nn

from ._fused import premodules

# group id of each of the 18 StyleGAN2 w+ layers: 4 coarse, 8 medium, 6 fine (modules/semantic_pe.py:6-8)
_LAYER_GROUPS = [0] * 4 + [1] * 8 + [2] * 6


class SemanticPE(nn.Module):
    """y[b, l, :] = x[b, l, :] + group_embed[groups[l]] + layer_embed[l].

    Keys ``group_embed.weight``, ``layer_embed.weight`` and the persistent buffer ``groups`` as in the reference
    (modules/semantic_pe.py:25-34).
    """

    def __init__(self, d_model: int = 512, num_layers: int = 18):
        super().__init__()
        self.group_embed = nn.Embedding(3, d_model)
        self.layer_embed = nn.Embedding(num_layers, d_model)
        self.register_buffer("groups", torch.tensor(_LAYER_GROUPS, dtype=torch.long))

    def forward(self, w_plus: torch.Tensor) -> torch.Tensor:
        n = w_plus.size(1)
        return premodules(w_plus, groups=self.groups, group_embed=self.group_embed.weight,
                          layer_embed=self.layer_embed.weight[:n])
