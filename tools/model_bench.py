"""Train-step throughput of the other BASELINE configs (parity-test cases, not bench.py lines): one JSON line each.

    python tools/model_bench.py [--steps 20] [--precision bf16]

config 1: LatentViT 512/6/8/2048, batch 32     config 2: ImageViT 512/6/8/2048 on 224x224 images, batch 64
config 4: LatentViTv2 (LEAM + SemanticPE + LayerWiseNorm with residual gate), batch 512
Step = zero_grad + forward + cross-entropy (label smoothing 0.1 for the LatentViT trainers) + backward + FusedAdamW,
replayed from one CUDA graph; dropout 0.1 as in the reference defaults.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fer_vit_b200 as fv  # noqa: E402


def run(name, model, x, y, steps, smoothing):
    model = model.cuda().train()
    opt = fv.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
    loss_fn = lambda lg, yy: fv.cross_entropy(lg, yy, None, smoothing)
    step = fv.GraphedTrainStep(model, opt, x, y, loss_fn=loss_fn)
    for _ in range(3):
        step(x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": name, "batch": x.shape[0], "ms_per_step": round(ms, 3),
                      "samples_per_s": round(x.shape[0] / ms * 1e3, 1), "launches_per_step": step.launches_per_replay,
                      "loss": round(float(step(x, y)), 4)}), flush=True)


def run_trainer_step(name, model, B, steps):
    """The LatentViT trainers' whole step (train_latent_vit.py:115-142) from a packed latent cache: gather + the
    default train augmentation + mixup, mixup loss (label smoothing 0.1), backward, AdamW, and the extra no-grad
    forward for train accuracy - one graph replay per step."""
    model = model.cuda().train()
    N = 16384                                              # 16,384 latents = 604 MB (> L2)
    cache = fv.PackedLatentCache(torch.randn(N, 18, 512), torch.randint(0, 7, (N,)), fv.get_latent_train_transforms())
    opt = fv.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
    step = fv.GraphedMixupTrainStep(model, opt, cache, B, fv.CrossEntropyLoss(None, 0.1))
    idx = [torch.randint(0, N, (B,), device="cuda") for _ in range(8)]
    mix = [torch.randperm(B, device="cuda") for _ in range(8)]
    for i in range(3):
        step(idx[i], mix[i], 0.4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        step(idx[i % 8], mix[i % 8], 0.4)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    loss, correct = step(idx[0], mix[0], 0.4)
    cache.check()
    print(json.dumps({"config": name, "batch": B, "ms_per_step": round(ms, 3),
                      "samples_per_s": round(B / ms * 1e3, 1), "launches_per_step": step.launches_per_replay,
                      "loss": round(float(loss), 4), "train_correct": int(correct)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    fv.set_default_precision(a.precision)
    torch.manual_seed(42)
    g = torch.Generator(device="cuda").manual_seed(1)
    lat = lambda B: torch.randn(B, 18, 512, device="cuda", generator=g)
    lab = lambda B: torch.randint(0, 7, (B,), device="cuda", generator=g)
    run("1: LatentViT 512/6/8/2048", fv.LatentViT(), lat(32), lab(32), a.steps, 0.1)
    run("4: LatentViTv2 all pre-modules", fv.LatentViTv2(use_lwn=True, use_lwn_residual=True, use_spe=True,
                                                         use_leam=True), lat(512), lab(512), a.steps, 0.1)
    run_trainer_step("4 + data path: LatentViTv2, cache batch + augment + mixup + accuracy pass",
                     fv.LatentViTv2(use_lwn=True, use_lwn_residual=True, use_spe=True, use_leam=True), 512, a.steps)
    run_trainer_step("1 + data path: LatentViT, cache batch + augment + mixup + accuracy pass", fv.LatentViT(), 32,
                     a.steps)
    run("2: ImageViT 512/6/8/2048 @224", fv.ImageViT(embed_dim=512, depth=6, heads=8, mlp_dim=2048),
        torch.randn(64, 3, 224, 224, device="cuda", generator=g), lab(64), a.steps, 0.0)


if __name__ == "__main__":
    main()
