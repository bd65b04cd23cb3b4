"""Golden fixtures for the data-side prologue of the LatentViT train step, from the UNMODIFIED reference code:

  latent_augment.npz  `LatentAugment.__call__` (data/latent_dataset.py:28-49) on seeded latents, with the draws torch's
                      generator produced for it captured by replaying the same generator calls;
  mixup_step.npz      one batch through the reference's own `train_epoch` (train/train_latent_vit.py:108-142, mixup
                      alpha 0.4, class weights, label smoothing 0.1) on the `latent_vit` fixture's weights: loss,
                      every gradient, and the train accuracy of its extra no-grad forward.

It also pins the oracle's restatements (`latent_augment`, `mixup`, `mixup_loss`) on those outputs in fp64.
Runs only where /root/reference exists; the .npz files are committed.

    python tests/golden/make_golden_data_path.py
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


def augment_cases():
    sys.path.insert(0, REF)
    ds = importlib.import_module("data.latent_dataset")
    assert ds.__file__.startswith(REF)
    out = {}
    cases = {"all": (0.1, (0.9, 1.1), 0.1), "noise": (0.05, None, 0.0), "scale_mask": (0.0, (0.5, 1.5), 0.3),
             "off": (0.0, None, 0.0)}
    for name, (std, rng, p) in cases.items():
        x = torch.randn(18, 512, generator=torch.Generator().manual_seed(11)) + 0.3
        # replay the generator calls of LatentAugment.__call__ to capture its draws ...
        torch.manual_seed(123)
        normal = torch.randn_like(x) if std > 0 else torch.zeros_like(x)
        scale = torch.empty(1).uniform_(*rng).item() if rng is not None else 1.0
        keep = (torch.rand_like(x) > p) if p > 0 else torch.ones_like(x, dtype=torch.bool)
        # ... then run the reference on the same generator state
        torch.manual_seed(123)
        y = ds.LatentAugment(noise_std=std, scale_range=rng, mask_prob=p)(x)
        ours = R.latent_augment(x, std, rng, p, normal, torch.tensor(scale), keep)
        assert torch.equal(ours, y), f"oracle latent_augment disagrees with the reference [{name}]"
        o64 = R.latent_augment(x.double(), std, rng, p, normal.double(), torch.tensor(scale, dtype=torch.float64), keep)
        print(f"  latent_augment [{name}]: fp32 bit-exact, fp64 vs fp32 reference {relerr(o64, y):.2e}")
        out.update({f"{name}/x": x.numpy(), f"{name}/normal": normal.numpy(), f"{name}/scale": np.float32(scale),
                    f"{name}/keep": keep.numpy(), f"{name}/out": y.numpy(), f"{name}/noise_std": np.float32(std),
                    f"{name}/scale_range": np.asarray(rng if rng is not None else (0.0, 0.0), dtype=np.float32),
                    f"{name}/use_scale": np.int32(rng is not None), f"{name}/mask_prob": np.float32(p)})
    np.savez_compressed(os.path.join(HERE, "latent_augment.npz"), **out)
    print("wrote latent_augment.npz")


class OneBatch(list):
    """What train_epoch needs of a DataLoader: iteration and len(loader.dataset)."""

    def __init__(self, x, y):
        super().__init__([(x, y)])
        self.dataset = range(x.shape[0])


def mixup_step():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "train"))
    tr = importlib.import_module("train.train_latent_vit")
    assert tr.__file__.startswith(REF)
    g = load_golden("latent_vit")
    B, alpha, smoothing = 8, 0.4, 0.1
    x = torch.randn(B, 18, 64, generator=torch.Generator().manual_seed(5))
    y = torch.randint(0, 7, (B,), generator=torch.Generator().manual_seed(6))
    w = g["class_weight"]
    res = {}
    for dtype in (torch.float64, torch.float32):
        model = tr.LatentViT(latent_dim=64, seq_len=18, embed_dim=64, depth=2, heads=2, mlp_dim=128, num_classes=7,
                             dropout=0.0)
        model.load_state_dict(g["sd"], strict=True)
        model = model.to(dtype)
        crit = nn.CrossEntropyLoss(weight=w.to(dtype), label_smoothing=smoothing)
        opt = torch.optim.SGD(model.parameters(), lr=0.0)      # keeps the weights, leaves p.grad in place
        tr.args = argparse.Namespace(mixup=alpha)               # train_epoch reads the script's global `args`
        np.random.seed(3)   # Beta(0.4, 0.4) draw 0.348: both terms of the blend carry weight
        torch.manual_seed(3)
        lam = float(np.random.beta(alpha, alpha))
        index = torch.randperm(B)
        np.random.seed(3)   # Beta(0.4, 0.4) draw 0.348: both terms of the blend carry weight
        torch.manual_seed(3)
        loss, acc, f1 = tr.train_epoch(model, OneBatch(x.to(dtype), y), opt, crit, torch.device("cpu"))
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        res[dtype] = (loss, acc, grads, lam, index)
    (l64, acc, g64, lam, index), (l32, acc32, g32, _, _) = res[torch.float64], res[torch.float32]
    assert acc == acc32
    # pin the oracle restatement (fp64) on the reference's step
    sd = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v) for k, v in g["sd"].items()}
    logits = R.latent_vit_forward(sd, R.mixup(x.double(), index, lam), 2, 2)
    loss = R.mixup_loss(logits, y, index, lam, w.double(), smoothing)
    grads = R.grads_of(loss, sd)
    e_g = max(relerr(grads[k], g64[k]) for k in g64)
    with torch.no_grad():
        pred = R.latent_vit_forward(sd, x.double(), 2, 2).argmax(-1)
    print(f"  mixup step: lam {lam:.6f}, loss {l64:.8f}; oracle fp64: loss {abs(loss.item() - l64):.2e}, worst grad "
          f"{e_g:.2e}; train acc {acc:.4f} (oracle {(pred == y).float().mean().item():.4f}); fp32 reference loss "
          f"diff {abs(l32 - l64):.2e}")
    assert abs(loss.item() - l64) < 1e-10 and e_g < 1e-8 and abs((pred == y).float().mean().item() - acc) < 1e-9
    out = {"x": x.numpy(), "y": y.numpy(), "index": index.numpy(), "lam": np.float64(lam), "loss": np.float64(l64),
           "accuracy": np.float64(acc), "pred": pred.numpy(), "class_weight": w.numpy(),
           "label_smoothing": np.float32(smoothing)}
    for k, v in g64.items():
        out["grad/" + k] = v.float().numpy()
    np.savez_compressed(os.path.join(HERE, "mixup_step.npz"), **out)
    print("wrote mixup_step.npz")


if __name__ == "__main__":
    augment_cases()
    mixup_step()
