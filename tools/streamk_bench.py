"""Stream-K A/B on the stand-alone GEMM entry point: microseconds per launch with and without the scratch buffer.

    python tools/streamk_bench.py [M N K]...
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fer_vit_b200 import _lib as L  # noqa: E402


def bench(fn, iters=40):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    shapes = [(4864, 768, 3072), (4864, 768, 2304), (9728, 768, 3072)]
    if len(sys.argv) > 3:
        a = list(map(int, sys.argv[1:]))
        shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)]
    lib = L.lib()
    st = torch.cuda.current_stream().cuda_stream
    nbytes = int(lib.fervit_gemm_scratch_bytes())
    scratch = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    for M, N, K in shapes:
        x = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        r = torch.randn(M, N, device="cuda")
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        of = torch.empty(M, N, device="cuda")
        f_bf = lambda: lib.fervit_linear_forward(L.BF16, x.data_ptr(), W.data_ptr(), None, None, M, N, K, L.ACT_NONE,
                                                 out.data_ptr(), None, None, 0, st)
        f_res = lambda: lib.fervit_linear_forward(L.BF16, x.data_ptr(), W.data_ptr(), None, r.data_ptr(), M, N, K,
                                                  L.ACT_NONE, None, of.data_ptr(), None, 0, st)
        row = {}
        for name, f in (("bf16_out", f_bf), ("f32_residual", f_res)):
            L.check(lib.fervit_set_gemm_scratch(None, 0))
            t0 = bench(f)
            L.check(lib.fervit_set_gemm_scratch(scratch.data_ptr(), nbytes))
            t1 = bench(f)
            L.check(lib.fervit_set_gemm_scratch(None, 0))
            row[name] = (round(t0, 2), round(t1, 2))
        print(M, N, K, {k: f"plain {v[0]} us, stream-K {v[1]} us" for k, v in row.items()}, flush=True)


if __name__ == "__main__":
    main()
