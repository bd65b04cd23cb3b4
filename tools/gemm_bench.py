"""Micro-benchmark of the tcgen05 GEMM through the C ABI (CUDA events, L2 flushed between launches).

    python tools/gemm_bench.py [--shapes vitb] [--bn 0,128,256] [--iters 20] [--one M,N,K,bn]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fer_vit_b200 import _lib as L  # noqa: E402

VITB = [("qkv", 4864, 2304, 768), ("proj", 4864, 768, 768), ("fc1", 4864, 3072, 768), ("fc2", 4864, 768, 3072),
        ("qkv_dgrad", 4864, 768, 2304), ("in_proj", 4608, 768, 512), ("ad_down", 4864, 64, 768), ("ad_up", 4864, 768, 64),
        ("qkv_b512", 9728, 2304, 768), ("fc1_b2048", 38912, 3072, 768)]


def time_gemm(M, N, K, bn, iters, act=0, bias=False, residual=False, out_f32=False, flush=None):
    x = torch.randn(M, K, device="cuda").bfloat16()
    W = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    of = torch.empty(M, N, device="cuda") if out_f32 else None
    b = torch.randn(N, device="cuda") if bias else None
    r = torch.randn(M, N, device="cuda") if residual else None
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if act else None
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr() if t is not None else None

    def run():
        L.check(L.lib().fervit_linear_forward(L.BF16, x.data_ptr(), W.data_ptr(), p(b), p(r), M, N, K, act,
                                              None if out_f32 else out.data_ptr(), p(of), p(pre), bn, st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bn", default="0,128,-256")
    ap.add_argument("--iters", type=int, default=15)
    ap.add_argument("--one", default="")
    ap.add_argument("--noflush", action="store_true")
    ap.add_argument("--kinds", default="plain", help="epilogue families for --sweep: plain,bias,gelu,res_f32")
    ap.add_argument("--sweep", default="", help="M,N,K,bn[;M,N,K,bn...]: time each shape under every isolation flag")
    a = ap.parse_args()
    flush = None if a.noflush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    if a.sweep:
        # isolation runs (results are garbage, timings only): see Params::debug in csrc/gemm_tc2.cu (bn >= 0, CTA-pair
        # kernel) and csrc/gemm_tc.cu (bn < 0, single-CTA kernel). kw selects the epilogue family.
        import ctypes
        names = {0: "full", 4: "no epilogue", 5: "mma only", 6: "tma only", 3: "epilogue only"}
        kws = {"plain": {}, "bias": dict(bias=True), "gelu": dict(act=2, bias=True),
               "res_f32": dict(bias=True, residual=True, out_f32=True)}
        for shape in a.sweep.split(";"):
            M, N, K, bn = [int(v) for v in shape.split(",")]
            for kname in a.kinds.split(","):
                for dbg, nm in names.items():
                    os.environ["FERVIT_GEMM_DEBUG"] = str(dbg | 8)
                    t = time_gemm(M, N, K, bn, a.iters, flush=flush, **kws[kname])
                    ns, cyc = ctypes.c_double(), ctypes.c_double()
                    ghz = None
                    if bn >= 0:
                        L.check(L.lib().fervit_debug_gemm_clock(ctypes.byref(ns), ctypes.byref(cyc)))
                        ghz = round(cyc.value / max(ns.value, 1.0), 3)
                    print(json.dumps({"M": M, "N": N, "K": K, "bn": bn, "epi": kname, "debug": dbg, "what": nm,
                                      "us": round(t * 1e6, 1), "tflops": round(2 * M * N * K / t / 1e12, 1),
                                      "kernel_us_cta0": round(ns.value / 1e3, 1), "sm_ghz": ghz}), flush=True)
        os.environ["FERVIT_GEMM_DEBUG"] = "0"
        return
    if a.one:
        M, N, K, bn = [int(v) for v in a.one.split(",")]
        t = time_gemm(M, N, K, bn, a.iters, flush=flush)
        print(json.dumps({"M": M, "N": N, "K": K, "bn": bn, "us": t * 1e6, "tflops": 2 * M * N * K / t / 1e12}))
        return
    for name, M, N, K in VITB:
        for bn in [int(v) for v in a.bn.split(",")]:
            if abs(bn) > 64 and N <= abs(bn) // 2:
                continue
            t = time_gemm(M, N, K, bn, a.iters, flush=flush)
            print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "bn": bn, "us": round(t * 1e6, 1),
                              "tflops": round(2 * M * N * K / t / 1e12, 1)}), flush=True)
    # epilogue cost: fc1 with GELU + two outputs, proj with residual fp32
    for name, M, N, K, kw in [("fc1+gelu", 4864, 3072, 768, dict(act=2, bias=True)),
                              ("proj+res_f32", 4864, 768, 768, dict(bias=True, residual=True, out_f32=True)),
                              ("fc2+res_f32", 4864, 768, 3072, dict(bias=True, residual=True, out_f32=True))]:
        for bn in sorted({0, *[int(v) for v in a.bn.split(",") if int(v) >= 0]}):
            t = time_gemm(M, N, K, bn, a.iters, flush=flush, **kw)
            print(json.dumps({"gemm": name, "bn": bn, "us": round(t * 1e6, 1),
                              "tflops": round(2 * M * N * K / t / 1e12, 1)}), flush=True)


if __name__ == "__main__":
    main()
