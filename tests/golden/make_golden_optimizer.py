"""Golden fixture for the optimizer end of the train step, from the UNMODIFIED reference trainer:

  v2_train_epoch.npz  one batch through `train_epoch` of train/train_latent_vit_v2.py:106-148 (mixup alpha 0.4, weighted
                      CE with label smoothing, `clip_grad_norm_` at 0.05 so that clipping is active, AdamW lr 1e-3 /
                      weight decay 0.01) on the `latent_vit_v2` fixture's weights: loss, train accuracy, and every
                      parameter AFTER the optimizer step.

Pins the oracle's `clip_grad_norm` and `adamw_step` restatements (fp64) on it.

    python tests/golden/make_golden_optimizer.py
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import reference_math as R  # noqa: E402
from tests.util import load_golden, relerr  # noqa: E402


class OneBatch(list):
    def __init__(self, x, y):
        super().__init__([(x, y)])
        self.dataset = range(x.shape[0])


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "train"))
    tr = importlib.import_module("train.train_latent_vit_v2")
    assert tr.__file__.startswith(REF)
    g = load_golden("latent_vit_v2")
    B, alpha, smoothing, clip, lr, wd = 8, 0.4, 0.1, 0.05, 1e-3, 0.01
    x = torch.randn(B, 18, 64, generator=torch.Generator().manual_seed(15)) * 0.7 + 0.3
    y = torch.randint(0, 7, (B,), generator=torch.Generator().manual_seed(16))
    w = torch.rand(7, generator=torch.Generator().manual_seed(17)) + 0.5
    model = tr.LatentViTv2(latent_dim=64, seq_len=18, embed_dim=64, depth=2, heads=2, mlp_dim=128, num_classes=7,
                           dropout=0.0, use_lwn=True, use_lwn_residual=True, use_spe=True, use_leam=True)
    model.load_state_dict(g["sd"], strict=True)
    model = model.double()
    crit = nn.CrossEntropyLoss(weight=w.double(), label_smoothing=smoothing)
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wd)
    args = argparse.Namespace(mixup=alpha, grad_clip=clip)
    np.random.seed(3)
    torch.manual_seed(3)
    lam = float(np.random.beta(alpha, alpha))
    index = torch.randperm(B)
    np.random.seed(3)
    torch.manual_seed(3)
    loss, acc, f1 = tr.train_epoch(model, OneBatch(x.double(), y), opt, crit, torch.device("cpu"), args)
    after = {k: v.detach().clone() for k, v in model.state_dict().items()}

    # oracle: forward + mixup loss + gradients + clipping + AdamW, fp64
    sd = {k: (v.double().requires_grad_(v.is_floating_point() and k in dict(model.named_parameters()))
              if v.is_floating_point() else v) for k, v in g["sd"].items()}
    logits = R.latent_vit_v2_forward(sd, R.mixup(x.double(), index, lam), 2, 2, True, True, True, True)
    oloss = R.mixup_loss(logits, y, index, lam, w.double(), smoothing)
    grads = R.grads_of(oloss, sd)
    clipped, total = R.clip_grad_norm(grads, clip)
    assert float(total) > clip, "clipping must be active in this fixture"
    worst = 0.0
    for k, gr in clipped.items():
        p0 = sd[k].detach()
        p1, _, _ = R.adamw_step(p0, gr, torch.zeros_like(p0), torch.zeros_like(p0), 1, lr, (0.9, 0.999), 1e-8, wd)
        # compare the UPDATE (the parameters themselves agree trivially to 1e-3)
        worst = max(worst, relerr(p1 - p0, after[k] - p0))
    print(f"  v2 train_epoch: lam {lam:.6f}, loss {loss:.8f} (oracle diff {abs(oloss.item() - loss):.2e}), grad norm "
          f"{float(total):.4f} -> clipped to {clip}; worst parameter-update error oracle vs reference {worst:.2e}; "
          f"train acc {acc:.4f}")
    assert abs(oloss.item() - loss) < 1e-10 and worst < 1e-7
    out = {"x": x.numpy(), "y": y.numpy(), "index": index.numpy(), "lam": np.float64(lam), "loss": np.float64(loss),
           "accuracy": np.float64(acc), "class_weight": w.numpy(), "label_smoothing": np.float64(smoothing),
           "grad_clip": np.float64(clip), "lr": np.float64(lr), "weight_decay": np.float64(wd),
           "total_norm": np.float64(float(total))}
    for k, v in after.items():
        if v.is_floating_point():
            out["after/" + k] = v.numpy()          # fp64: the update is ~1e-3 of the parameter
    np.savez_compressed(os.path.join(HERE, "v2_train_epoch.npz"), **out)
    print("wrote v2_train_epoch.npz")


if __name__ == "__main__":
    main()
