"""Microseconds per launch of the latency-bound kernels of the step (attention fwd/bwd, LayerNorm fwd/bwd) at the
config-3 shapes, graph-replayed over rotating buffers (> L2) so that neither the Python call nor a warm L2 hides the
kernel. One JSON line per kernel with the achieved GB/s on algorithmic bytes against MEASURED_PEAKS.json.

    python tools/small_kernel_bench.py [--batch 256] [--iters 24]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fer_vit_b200 import _lib as L  # noqa: E402


def replay_us(fn, iters):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(iters):
            fn(i)
    graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=24)
    ap.add_argument("--seq", type=int, default=19, help="tokens per sample (197 = ImageViT, config 2)")
    ap.add_argument("--heads", type=int, default=12)
    ap.add_argument("--drop", type=float, default=0.0, help="dropout on the attention weights")
    a = ap.parse_args()
    peak = 6551.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    lib = L.lib()
    B, S, H, hd = a.batch, a.seq, a.heads, 64
    E, T = H * hd, a.batch * a.seq
    n = 8                                                   # buffer sets in rotation: 8 x ~60 MB > L2
    st = lambda: torch.cuda.current_stream().cuda_stream
    bf = torch.bfloat16
    qkv = [torch.randn(T, 3 * E, device="cuda").to(bf) for _ in range(n)]
    dout = [torch.randn(T, E, device="cuda").to(bf) for _ in range(n)]
    out = [torch.empty(T, E, device="cuda", dtype=bf) for _ in range(n)]
    lse = [torch.empty(B * H * S, device="cuda") for _ in range(n)]
    dqkv = [torch.empty(T, 3 * E, device="cuda", dtype=bf) for _ in range(n)]
    x = [torch.randn(T, E, device="cuda") for _ in range(n)]
    dres = [torch.randn(T, E, device="cuda") for _ in range(n)]
    dx = [torch.empty(T, E, device="cuda") for _ in range(n)]
    y = [torch.empty(T, E, device="cuda", dtype=bf) for _ in range(n)]
    mean = [torch.empty(T, device="cuda") for _ in range(n)]
    rstd = [torch.empty(T, device="cuda") for _ in range(n)]
    gamma, beta = torch.ones(E, device="cuda"), torch.zeros(E, device="cuda")

    def attn_fwd(i):
        k = i % n
        L.check(lib.fervit_attention_forward(L.BF16, qkv[k].data_ptr(), B, S, H, hd, a.drop, 1234, 7, out[k].data_ptr(),
                                             lse[k].data_ptr(), st()))

    def attn_bwd(i):
        k = i % n
        L.check(lib.fervit_attention_backward(L.BF16, qkv[k].data_ptr(), out[k].data_ptr(), dout[k].data_ptr(),
                                              lse[k].data_ptr(), B, S, H, hd, a.drop, 1234, 7, dqkv[k].data_ptr(), st()))

    def ln_fwd(i):
        k = i % n
        L.check(lib.fervit_layernorm_forward(L.BF16, x[k].data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-6, T, E,
                                             None, y[k].data_ptr(), mean[k].data_ptr(), rstd[k].data_ptr(), st()))

    def ln_bwd(i):
        k = i % n
        L.check(lib.fervit_layernorm_backward(L.BF16, dout[k].data_ptr(), x[k].data_ptr(), mean[k].data_ptr(),
                                              rstd[k].data_ptr(), gamma.data_ptr(), dres[k].data_ptr(), T, E,
                                              dx[k].data_ptr(), y[k].data_ptr(), None, None, None, st()))

    for i in range(n):                                      # valid forward state for the backward kernels
        attn_fwd(i)
        ln_fwd(i)
    cases = [("attention_fwd", attn_fwd, B * H * S * hd * 2 * 4 + B * H * S * 4),
             ("attention_bwd", attn_bwd, B * H * S * hd * 2 * 8 + B * H * S * 4),
             ("layernorm_fwd", ln_fwd, T * E * (4 + 2) + T * 8),
             ("layernorm_bwd", ln_bwd, T * E * (2 + 4 + 4 + 4 + 2) + T * 8)]
    for name, fn, nbytes in cases:
        us = replay_us(fn, a.iters)
        gbs = nbytes / us / 1e3
        print(json.dumps({"kernel": name, "batch": B, "seq": S, "heads": H, "attn_dropout": a.drop, "us": round(us, 2), "algorithmic_mb": round(nbytes / 1e6, 1),
                          "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3)}), flush=True)


if __name__ == "__main__":
    main()
