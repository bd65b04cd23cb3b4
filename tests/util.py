"""Shared helpers of the parity tests: golden fixtures, model construction from a fixture, error metrics."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["latent_vit", "latent_vit_v2", "hybrid_adapter", "hybrid_full", "image_vit"]


def load_golden(name: str) -> Dict[str, object]:
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {"sd": {}, "grad": {}, "meta": {}}
    for k in z.files:
        if k.startswith("sd/"):
            out["sd"][k[3:]] = torch.from_numpy(z[k])
        elif k.startswith("grad/"):
            out["grad"][k[5:]] = torch.from_numpy(z[k])
        elif k.startswith("meta/"):
            out["meta"][k[5:]] = int(z[k])
        else:
            out[k] = torch.from_numpy(np.asarray(z[k]))
    out["class_weight"] = out.get("class_weight", None)
    out["label_smoothing"] = float(out["label_smoothing"])
    return out


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    """Norm-wise relative error ||a - b|| / ||b|| (the metric of every parity gate in this repo)."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def oracle_forward(name: str, sd, x, masks=None):
    from oracle import reference_math as R
    if name == "latent_vit":
        return R.latent_vit_forward(sd, x, 2, 2, masks)
    if name == "latent_vit_v2":
        return R.latent_vit_v2_forward(sd, x, 2, 2, True, True, True, True, masks)
    if name == "hybrid_adapter":
        return R.hybrid_forward(sd, x, 2, 2, True, masks)
    if name == "hybrid_full":
        return R.hybrid_forward(sd, x, 2, 2, False, masks)
    if name == "image_vit":
        return R.image_vit_forward(sd, x, 2, 2, 16, masks)
    raise KeyError(name)


def build_model(name: str, precision: str = "fp32"):
    """The drop-in model class shaped like fixture `name` (weights still random)."""
    import fer_vit_b200 as fv
    from fer_vit_b200.models_fer_vit.vit_blocks import register_vit_config
    fv.set_default_precision(precision)
    register_vit_config("vit_test_patch16_224", 64, 2, 2)
    if name == "latent_vit":
        return fv.LatentViT(latent_dim=64, seq_len=18, embed_dim=64, depth=2, heads=2, mlp_dim=128, num_classes=7,
                            dropout=0.0)
    if name == "latent_vit_v2":
        return fv.LatentViTv2(latent_dim=64, seq_len=18, embed_dim=64, depth=2, heads=2, mlp_dim=128, num_classes=7,
                              dropout=0.0, use_lwn=True, use_lwn_residual=True, use_spe=True, use_leam=True)
    if name == "hybrid_adapter":
        return fv.HybridLatentViT(latent_dim=64, seq_len=18, pretrained_model_name="vit_test_patch16_224",
                                  num_classes=7, use_pretrained=False, freeze_transformer=True, adapter_dim=16,
                                  verbose=False)
    if name == "hybrid_full":
        return fv.HybridLatentViT(latent_dim=64, seq_len=18, pretrained_model_name="vit_test_patch16_224",
                                  num_classes=7, use_pretrained=False, freeze_transformer=False, adapter_dim=None,
                                  verbose=False)
    if name == "image_vit":
        return fv.ImageViT(img_size=32, patch_size=16, in_channels=3, embed_dim=64, depth=2, heads=2, mlp_dim=128,
                           num_classes=7, dropout=0.0)
    raise KeyError(name)
