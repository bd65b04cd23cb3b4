"""ORACLE — test infrastructure, not product code.

Runs the UNMODIFIED reference classes (and, for the headline configuration, the reference trainer's own
`train_epoch`) from `baseline/_ref/` (oracle/make_ref.py) on synthetic inputs, on the host CPU or — as the same-box
comparison point BASELINE.md §5 asks for — with torch eager on the GPU. Only bench.py's `--impl reference` arm and its
`cpu_baseline` / `gpu_eager_reference` legs call this; nothing here is on the product path.

The reference step timed is train/train_hybrid_latent_vit.py:127-142 (zero_grad, forward, CrossEntropyLoss, backward,
`optim.AdamW.step()`, `loss.item()`, argmax, `.cpu()`), with `optim.AdamW(model.parameters(), lr, weight_decay)` as at
:248. HybridLatentViT needs `timm`, which this image does not have: oracle/timm_shim.py stands in for it (DESIGN.md
"Oracle and pinning").
"""
from __future__ import annotations

import importlib
import os
import sys
import time
from typing import Dict, Optional

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

CONFIGS = {
    # name: (BASELINE.json config index, default batch)
    "hybrid": (2, 256),
    "latent_vit": (0, 32),
    "image_vit": (1, 64),
    "latent_vit_v2": (3, 512),
}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "models_fer_vit"))


def _import_reference():
    """Put baseline/_ref first on sys.path (the reference's scripts do the same with their repo root,
    train_latent_vit.py:20-23) and register the timm stand-in."""
    if not available():
        raise RuntimeError(f"{REF_DIR} is missing: run `python -m oracle.make_ref` where /root/reference exists")
    from . import timm_shim
    timm_shim.install()
    for p in (REF_DIR, os.path.join(REF_DIR, "train")):
        if p not in sys.path:
            sys.path.insert(0, p)


class _Batches(list):
    """A DataLoader stand-in: a list of (x, y) batches with the `.dataset` the trainers take `len()` of."""

    def __init__(self, batches):
        super().__init__(batches)
        self.dataset = range(sum(int(b[0].shape[0]) for b in batches))


def build_model(config: str) -> nn.Module:
    _import_reference()
    if config == "hybrid":
        m = importlib.import_module("models_fer_vit.hybrid_latent_vit")
        assert m.__file__.startswith(REF_DIR), m.__file__
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):     # the constructor prints a parameter summary
            return m.create_hybrid_latent_vit(latent_dim=512, seq_len=18, model_size="base", num_classes=7,
                                              use_pretrained=False, freeze_transformer=True, freeze_stages=None,
                                              use_adapter=True, adapter_dim=64)
    if config == "latent_vit":
        m = importlib.import_module("models_fer_vit.latent_vit")
        assert m.__file__.startswith(REF_DIR), m.__file__
        return m.LatentViT(latent_dim=512, seq_len=18, embed_dim=512, depth=6, heads=8, mlp_dim=2048, num_classes=7,
                           dropout=0.1)
    if config == "latent_vit_v2":
        m = importlib.import_module("models_fer_vit.latent_vit_v2")
        assert m.__file__.startswith(REF_DIR), m.__file__
        return m.LatentViTv2(latent_dim=512, seq_len=18, embed_dim=512, depth=6, heads=8, mlp_dim=2048, num_classes=7,
                             dropout=0.1, use_lwn=True, use_lwn_residual=True, use_spe=True, use_leam=True)
    if config == "image_vit":
        m = importlib.import_module("models_fer_vit.image_vit")
        assert m.__file__.startswith(REF_DIR), m.__file__
        return m.ImageViT(img_size=224, patch_size=16, in_channels=3, embed_dim=512, depth=6, heads=8, mlp_dim=2048,
                          num_classes=7, dropout=0.1)
    raise ValueError(config)


def synthetic_batch(config: str, B: int, seed: int, device) -> tuple:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, 224, 224, generator=g) if config == "image_vit" else torch.randn(B, 18, 512, generator=g)
    y = torch.randint(0, 7, (B,), generator=g)
    return x.to(device), y.to(device)


def _hybrid_train_epoch():
    """The reference trainer's own loop (train/train_hybrid_latent_vit.py:120-147), or None if its module does not
    import here (it pulls in sklearn and tensorboard at import time)."""
    try:
        tr = importlib.import_module("train.train_hybrid_latent_vit")
        assert tr.__file__.startswith(REF_DIR), tr.__file__
        return tr.train_epoch
    except Exception:
        return None


def _plain_epoch(model, loader, optimizer, criterion, device):
    """The a14 skeleton of every reference trainer (zero_grad, forward, loss, backward, step, .item(), argmax, .cpu()):
    train_hybrid_latent_vit.py:127-142 / train_image_vit.py:117-137, without the sklearn metrics at the end."""
    model.train()
    total, preds = 0.0, []
    for x, y in loader:
        x, y = x.to(device), y.to(device)
        optimizer.zero_grad()
        logits = model(x)
        loss = criterion(logits, y)
        loss.backward()
        optimizer.step()
        total += loss.item() * x.size(0)
        preds.extend(logits.argmax(dim=1).cpu().numpy())
    return total / len(loader.dataset), 0.0, 0.0


def time_train_steps(config: str, B: int, steps: int, warmup: int, device: str = "cpu", autocast_bf16: bool = False,
                     threads: Optional[int] = None, seed: int = 42) -> Dict[str, object]:
    """Seconds per reference train step: `warmup` untimed then `steps` timed steps through the trainer's loop."""
    if threads and device == "cpu":
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    dev = torch.device(device)
    model = build_model(config).to(dev)
    opt = torch.optim.AdamW([p for p in model.parameters()], lr=1e-3, weight_decay=0.01)
    crit = nn.CrossEntropyLoss()
    epoch = (_hybrid_train_epoch() if config == "hybrid" else None) or _plain_epoch
    batches = [synthetic_batch(config, B, seed + 1 + i, dev) for i in range(min(4, max(1, steps)))]

    def run(n):
        loader = _Batches([batches[i % len(batches)] for i in range(n)])
        if autocast_bf16:
            with torch.autocast(device_type=dev.type, dtype=torch.bfloat16):
                return epoch(model, loader, opt, crit, dev)
        return epoch(model, loader, opt, crit, dev)

    if warmup > 0:
        run(warmup)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss, _, _ = run(steps)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"s_per_step": dt / steps, "samples_per_s": B * steps / dt, "B": B, "steps": steps, "loss": float(loss),
            "loop": "train_hybrid_latent_vit.train_epoch" if epoch is not _plain_epoch else "train-step skeleton",
            "threads": torch.get_num_threads() if dev.type == "cpu" else None}
