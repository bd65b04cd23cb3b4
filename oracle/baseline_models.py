"""ORACLE — test infrastructure, not product code.

Seeded random `state_dict`s with the reference's keys and shapes for each model family, and a timed fwd+loss+bwd step
of the oracle on host cores (bench.py's cpu_baseline / `--impl reference` arm). No reference code is needed, so this
travels to the GPU box.
"""
from __future__ import annotations

import time
from typing import Dict

import torch

from . import reference_math as R


def _lin(sd, name, out_f, in_f, g, std=0.02):
    sd[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * std
    sd[name + ".bias"] = torch.randn(out_f, generator=g) * 0.01


def _ln(sd, name, dim, g):
    sd[name + ".weight"] = 1.0 + 0.05 * torch.randn(dim, generator=g)
    sd[name + ".bias"] = 0.05 * torch.randn(dim, generator=g)


def hybrid_state_dict(E=768, depth=12, heads=12, latent_dim=512, seq_len=18, num_classes=7, adapter_dim=64,
                      seed=0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    _lin(sd, "input_proj", E, latent_dim, g)
    sd["cls_token"] = torch.randn(1, 1, E, generator=g) * 0.02
    sd["pos_embed"] = torch.randn(1, seq_len + 1, E, generator=g) * 0.02
    for i in range(depth):
        p = f"transformer.{i}."
        _ln(sd, p + "norm1", E, g)
        _lin(sd, p + "attn.qkv", 3 * E, E, g)
        _lin(sd, p + "attn.proj", E, E, g)
        _ln(sd, p + "norm2", E, g)
        _lin(sd, p + "mlp.fc1", 4 * E, E, g)
        _lin(sd, p + "mlp.fc2", E, 4 * E, g)
        if adapter_dim:
            a = f"adapters.{i}."
            _lin(sd, a + "adapter.0", adapter_dim, E, g)
            _lin(sd, a + "adapter.2", E, adapter_dim, g)
            sd[a + "alpha"] = torch.ones(1) * 0.1
    _ln(sd, "head.0", E, g)
    _lin(sd, "head.2", num_classes, E, g)
    return sd


def hybrid_trainable(sd: Dict[str, torch.Tensor], freeze_transformer=True) -> None:
    for k, v in sd.items():
        v.requires_grad_(not (freeze_transformer and k.startswith("transformer.")))


def time_hybrid_step(B: int, steps: int, warmup: int, threads: int, E=768, depth=12, heads=12, adapter_dim=64,
                     seed=0) -> Dict[str, float]:
    """Median seconds of one fp32 fwd + CE + bwd oracle step (frozen backbone + adapters) on `threads` host cores."""
    torch.set_num_threads(threads)
    sd = hybrid_state_dict(E, depth, heads, adapter_dim=adapter_dim, seed=seed)
    hybrid_trainable(sd)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, 18, 512, generator=g)
    y = torch.randint(0, 7, (B,), generator=g)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        logits = R.hybrid_forward(sd, x, depth, heads, adapter_dim > 0)
        loss = R.cross_entropy(logits, y)
        R.grads_of(loss, sd)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    times.sort()
    return {"median_s": times[len(times) // 2], "min_s": times[0], "B": B, "steps": steps}
