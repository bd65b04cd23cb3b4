"""HBM roofline of the latent batch producer (fervit_latent_batch): achieved GB/s on algorithmic bytes
(read row*4 per sample, + row*4 partner with mixup, + write row*4) against MEASURED_PEAKS.json.

    python tools/latent_batch_bench.py [--batch 512 4096] [--table 65536] [--iters 50]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fer_vit_b200 as fv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[512, 4096])
    ap.add_argument("--table", type=int, default=65536)          # 65,536 latents = 2.4 GB, > L2
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    peak = 6551.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    table = torch.randn(a.table, 18, 512, device="cuda")
    labels = torch.randint(0, 7, (a.table,), device="cuda")
    aug = fv.LatentAugment(0.1, (0.9, 1.1), 0.1)
    for B in a.batch:
        out = torch.empty(B, 18, 512, device="cuda")
        perm = torch.randperm(B, device="cuda")
        for mode in ("gather", "gather+augment", "gather+augment+mixup"):
            t = aug if "augment" in mode else None
            mix = perm if "mixup" in mode else None
            idxs = [torch.randint(0, a.table, (B,), device="cuda") for _ in range(a.iters)]   # fresh rows: no L2 reuse
            for i in range(3):
                fv.latent_batch(table, labels, idxs[i], t, mix, 0.3, 1 + i, None, out)
            # the launches go into one CUDA graph: the Python call costs more than the kernel at batch 512
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for i in range(a.iters):
                    fv.latent_batch(table, labels, idxs[i], t, mix, 0.3, 10 + i, None, out)
            graph.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / a.iters
            nbytes = B * 18 * 512 * 4 * (3 if mix is not None else 2)
            gbs = nbytes / us / 1e3
            print(json.dumps({"kernel": "latent_batch", "mode": mode, "batch": B, "us": round(us, 2),
                              "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3),
                              "samples_per_s": round(B / us * 1e6)}), flush=True)

    # LatentDecomposer (fervit_latent_decompose): read row*4 + write row*4 (2x for concat) per sample
    dec = fv.LatentDecomposer({i: torch.randn(18, 512) for i in range(7)}, 18, 512).cuda()
    for B in a.batch:
        xs = [torch.randn(B, 18, 512, device="cuda") for _ in range(8)]          # 8 x B latents in rotation
        for om in ("expr_only", "concat"):
            for i in range(3):
                dec(xs[i], output_mode=om)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for i in range(a.iters):
                    dec(xs[i % 8], output_mode=om)
            graph.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / a.iters
            nbytes = B * 18 * 512 * 4 * (3 if om == "concat" else 2)
            gbs = nbytes / us / 1e3
            print(json.dumps({"kernel": "latent_decompose", "mode": om, "batch": B, "us": round(us, 2),
                              "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3),
                              "samples_per_s": round(B / us * 1e6)}), flush=True)


if __name__ == "__main__":
    main()
