"""Golden fixture for the headline configuration's WHOLE optimizer step, from the unmodified reference trainer:

  hybrid_train_epoch.npz  one batch through `train_epoch` of train/train_hybrid_latent_vit.py:120-147 in train mode
                          (head Dropout(0.1) active) on the `hybrid_adapter` fixture's weights (frozen blocks +
                          adapters), optimizer = AdamW over the trainer's own `get_optimizer_groups` (:63-117: five
                          groups, lr x10 / x1 / x10 / x10 / x5, no decay on pos_embed + cls_token): loss, accuracy, the
                          head-dropout mask torch drew (captured by replaying the generator), every parameter after
                          the step.

Pins the oracle (forward with the injected mask, CE, gradients, per-group AdamW) on it in fp64.

    python tests/golden/make_golden_hybrid_step.py
"""
from __future__ import annotations

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
from oracle import timm_shim  # noqa: E402
from tests.util import load_golden, relerr  # noqa: E402


class OneBatch(list):
    def __init__(self, x, y):
        super().__init__([(x, y)])
        self.dataset = range(x.shape[0])


def group_of(model, groups):
    """parameter name -> (lr, weight_decay) under the trainer's grouping"""
    by_id = {id(p): (g["lr"], g["weight_decay"]) for g in groups for p in g["params"]}
    return {k: by_id[id(p)] for k, p in model.named_parameters() if id(p) in by_id}


def main():
    timm_shim.install()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "train"))
    tr = importlib.import_module("train.train_hybrid_latent_vit")
    assert tr.__file__.startswith(REF)
    g = load_golden("hybrid_adapter")
    B, lr, wd = 8, 1e-4, 0.01
    x = torch.randn(B, 18, 64, generator=torch.Generator().manual_seed(25))
    y = torch.randint(0, 7, (B,), generator=torch.Generator().manual_seed(26))
    model = tr.HybridLatentViT(latent_dim=64, seq_len=18, pretrained_model_name="vit_test_patch16_224", num_classes=7,
                               use_pretrained=False, freeze_transformer=True, adapter_dim=16)
    model.load_state_dict(g["sd"], strict=True)
    model = model.double()
    groups = tr.get_optimizer_groups(model, lr, wd)
    hyper = group_of(model, groups)
    opt = torch.optim.AdamW(groups, lr=lr, weight_decay=wd)
    crit = nn.CrossEntropyLoss()
    # the only generator draw of the step is the head Dropout(0.1) on the [B, E] cls features: replay it
    torch.manual_seed(9)
    mask = torch.nn.functional.dropout(torch.ones(B, 64, dtype=torch.float64), 0.1, True)
    torch.manual_seed(9)
    loss, acc, f1 = tr.train_epoch(model, OneBatch(x.double(), y), opt, crit, torch.device("cpu"))
    after = {k: v.detach().clone() for k, v in model.state_dict().items()}

    sd = {k: (v.double().requires_grad_(k in hyper) if v.is_floating_point() else v) for k, v in g["sd"].items()}
    logits = R.hybrid_forward(sd, x.double(), 2, 2, True, {"head": mask})
    oloss = R.cross_entropy(logits, y)
    grads = R.grads_of(oloss, sd)
    worst = 0.0
    for k, (glr, gwd) in hyper.items():
        p0 = sd[k].detach()
        p1, _, _ = R.adamw_step(p0, grads[k], torch.zeros_like(p0), torch.zeros_like(p0), 1, glr, (0.9, 0.999), 1e-8, gwd)
        worst = max(worst, relerr(p1 - p0, after[k] - p0))
    frozen_same = all(torch.equal(after[k], g["sd"][k].double()) for k in after if k.startswith("transformer."))
    oacc = (logits.argmax(-1) == y).double().mean().item()
    print(f"  hybrid train_epoch: loss {loss:.8f} (oracle diff {abs(oloss.item() - loss):.2e}); {len(hyper)} trainable "
          f"tensors in {len(groups)} groups; worst update error oracle vs reference {worst:.2e}; frozen blocks untouched: "
          f"{frozen_same}; accuracy {acc:.4f} (oracle {oacc:.4f})")
    assert abs(oloss.item() - loss) < 1e-10 and worst < 1e-7 and frozen_same and abs(oacc - acc) < 1e-12
    out = {"x": x.numpy(), "y": y.numpy(), "head_mask": mask.numpy(), "loss": np.float64(loss),
           "accuracy": np.float64(acc), "lr": np.float64(lr), "weight_decay": np.float64(wd)}
    for k, v in after.items():
        if k in hyper:
            out["after/" + k] = v.numpy()
            out["hyper/" + k] = np.asarray(hyper[k], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "hybrid_train_epoch.npz"), **out)
    print("wrote hybrid_train_epoch.npz")


if __name__ == "__main__":
    main()
