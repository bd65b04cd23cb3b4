"""Golden fixture for the ExpressionAwareViT front-end, from the UNMODIFIED reference classes
(models_fer_vit/latent_decomposer.py, expression_aware_vit.py; HybridLatentViT over the timm shim):

  expression_aware.npz  LatentDecomposer outputs in every (decompose_mode, output_mode) on seeded latents, its scores,
                        and one ExpressionAwareViT train step in 'concat' mode (36 + 1 tokens, frozen blocks +
                        adapters): logits, loss, every trainable gradient.

Also pins the oracle's restatement (reference_math.latent_decompose / decomposer_forward) in fp64.

    python tests/golden/make_golden_expression.py
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
from tests.util import relerr  # noqa: E402

MODES = [(d, o) for d in ("all_classes", "max_class") for o in ("expr_only", "id_only", "enhanced", "concat")]


def main():
    timm_shim.install()
    sys.path.insert(0, REF)
    dec = importlib.import_module("models_fer_vit.latent_decomposer")
    eav = importlib.import_module("models_fer_vit.expression_aware_vit")
    hyb = importlib.import_module("models_fer_vit.hybrid_latent_vit")
    for m in (dec, eav, hyb):
        assert m.__file__.startswith(REF)
    g = torch.Generator().manual_seed(21)
    L, D, C, B = 18, 64, 7, 5
    raw = {i: torch.randn(L, D, generator=g) * (0.5 + i) for i in range(C)}      # un-normalised, as an SVM gives them
    x = torch.randn(B, L, D, generator=g) + 0.2
    out = {"x": x.numpy(), "alpha": np.float32(1.7)}
    for i in range(C):
        out[f"raw_direction/{i}"] = raw[i].numpy()
    d32 = dec.LatentDecomposer(raw, L, D)
    d64 = dec.LatentDecomposer({i: v.double() for i, v in raw.items()}, L, D).double()
    out["directions"] = d32.directions.numpy()
    assert relerr(R.normalize_directions(torch.stack([raw[i] for i in range(C)]).double()), d64.directions) < 1e-15
    out["scores"] = d64.get_expression_scores(x.double()).float().numpy()
    for dm, om in MODES:
        ref = d64(x.double(), output_mode=om, enhance_alpha=1.7, decompose_mode=dm)
        ours = R.decomposer_forward(x.double(), d64.directions, om, 1.7, dm)
        assert relerr(ours, ref) < 1e-14, (dm, om)
        e32 = relerr(d32(x, output_mode=om, enhance_alpha=1.7, decompose_mode=dm), ref)
        out[f"out/{dm}/{om}"] = ref.float().numpy()
        print(f"  decomposer [{dm:11s} {om:9s}] oracle fp64 exact; reference fp32 vs fp64 {e32:.2e}")
    e, i_, _ = R.latent_decompose(x.double(), d64.directions)
    re, ri = d64.decompose(x.double())
    assert relerr(e, re) < 1e-14 and relerr(i_, ri) < 1e-14

    # ---- one ExpressionAwareViT step, concat mode (S = 37), frozen blocks + adapters, eval-mode head dropout
    torch.manual_seed(46)
    vit = hyb.HybridLatentViT(latent_dim=D, seq_len=2 * L, pretrained_model_name="vit_test_patch16_224", num_classes=7,
                              use_pretrained=False, freeze_transformer=True, adapter_dim=16)
    gp = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in vit.named_parameters():
            if p.dim() == 1 and "alpha" not in n:
                p.add_(0.1 * torch.randn(p.shape, generator=gp))
        for i, a in enumerate(vit.adapters):
            a.alpha.fill_(0.1 + 0.05 * i)
    sd = {k: v.detach().clone() for k, v in vit.state_dict().items()}
    model = eav.ExpressionAwareViT(d64, vit.double(), output_mode="concat", decompose_mode="all_classes").eval()
    y = torch.randint(0, 7, (B,), generator=g)
    logits = model(x.double())
    loss = nn.CrossEntropyLoss()(logits, y)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.vit.named_parameters() if p.grad is not None}
    assert len(grads) == len(model.get_trainable_params())
    # oracle: decomposer restatement feeding the hybrid restatement
    sd64 = {k: (v.double().requires_grad_(k in grads) if v.is_floating_point() else v) for k, v in sd.items()}
    ol = R.hybrid_forward(sd64, R.decomposer_forward(x.double(), d64.directions, "concat"), 2, 2, True)
    og = R.grads_of(R.cross_entropy(ol, y), sd64)
    e_g = max(relerr(og[k], grads[k]) for k in grads)
    print(f"  ExpressionAwareViT concat step: oracle fp64 logits {relerr(ol, logits):.2e}, worst grad {e_g:.2e}")
    assert relerr(ol, logits) < 1e-10 and e_g < 1e-8
    out.update({"y": y.numpy(), "logits": logits.detach().float().numpy(), "loss": np.float64(loss.item())})
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    for k, v in grads.items():
        out["grad/" + k] = v.float().numpy()
    np.savez_compressed(os.path.join(HERE, "expression_aware.npz"), **out)
    print("wrote expression_aware.npz")


if __name__ == "__main__":
    main()
