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
    run("2: ImageViT 512/6/8/2048 @224", fv.ImageViT(embed_dim=512, depth=6, heads=8, mlp_dim=2048),
        torch.randn(64, 3, 224, 224, device="cuda", generator=g), lab(64), a.steps, 0.0)


if __name__ == "__main__":
    main()
