"""Phase timeline of the CTA-pair tcgen05 GEMM (FERVIT_GEMM_DEBUG bit 64): where a launch spends its time.

    python tools/gemm_timeline.py [--shapes qkv,proj,fc1,fc2,fc2_dgrad] [--m 4864]

For each batch-256 shape of the ViT-B block the kernel stamps clock64 at its phase boundaries in two CTAs (CTA 0 and
the leader of the last pair); the table below is in microseconds from kernel entry (SM clock from %globaltimer).
The per-chunk epilogue stamps ("item 0 chunk j: ...") need a library built with FERVIT_TL_CHUNKS=1
(`FERVIT_TL_CHUNKS=1 python -m fer_vit_b200.build --force`): they cost 2-3 % of every GEMM and are compiled out by default.
Results of the GEMM are unaffected by the stamps; timings are of one warm launch with operands resident in L2, the
situation inside the train step.
"""
import argparse
import ctypes
import json
import os
import sys

os.environ["FERVIT_GEMM_DEBUG"] = "64"

import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fer_vit_b200 import _lib as L  # noqa: E402

SLOTS = {0: "entry", 1: "prologue done", 2: "grid dependency", 3: "first operands landed", 8: "first TMA issued",
         9: "last TMA issued", 18: "stores drained", 19: "final cluster sync", 20: "end"}
CHUNK = ["side input landed", "accumulator in registers", "math done", "staging tile free", "staged + fenced",
         "stores issued"]
for j in range(4):
    for k, nm in enumerate(CHUNK):
        SLOTS[32 + 6 * j + k] = f"item 0 chunk {j}: {nm}"
for it in range(4):
    SLOTS[4 + it] = f"MMAs of item {it} issued"
    SLOTS[10 + 2 * it] = f"accumulator {it} ready (warp 0)"
    SLOTS[11 + 2 * it] = f"epilogue {it} done (warp 0)"
    SLOTS[21 + it] = f"epilogue {it} done (warp 7)"


def shapes(M):
    E, F = 768, 3072
    return {
        "qkv": dict(N=3 * E, K=E, bias=True),
        "proj": dict(N=E, K=E, bias=True, residual=True, out_f32=True),
        "fc1": dict(N=F, K=E, bias=True, act=2 | 0x100),
        "fc2": dict(N=E, K=F, bias=True, residual=True, out_f32=True),
        "fc2_dgrad": dict(N=F, K=E, mul_bwd=True),
        "fc1_dgrad": dict(N=E, K=F),
        "qkv_dgrad": dict(N=E, K=3 * E),
    }


def run(name, M, N, K, bias=False, residual=False, out_f32=False, act=0, mul_bwd=False, iters=5, bn=0):
    x = torch.randn(M, K, device="cuda").bfloat16()
    W = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    of = torch.empty(M, N, device="cuda") if out_f32 else None
    b = torch.randn(N, device="cuda") if bias else None
    r = torch.randn(M, N, device="cuda") if residual else None
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if act else None
    aux = torch.randn(M, N, device="cuda").bfloat16() if mul_bwd else None
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr() if t is not None else None
    lib = L.lib()

    def launch():
        if mul_bwd:
            # dX = dY W (transposed weight [K_in = N here]) with the saved derivative multiplied in
            L.check(lib.fervit_linear_dgrad(L.BF16, x.data_ptr(), W.data_ptr(), p(aux), None, M, K, N, 3,
                                            out.data_ptr(), None, bn, st))   # reduction over K (API name: N)
        else:
            L.check(lib.fervit_linear_forward(L.BF16, x.data_ptr(), W.data_ptr(), p(b), p(r), M, N, K, act,
                                              None if out_f32 else out.data_ptr(), p(of), p(pre), bn, st))
    ev = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        ev.append(e0.elapsed_time(e1) * 1e3)
    buf = (ctypes.c_ulonglong * 128)()
    L.check(lib.fervit_debug_gemm_timeline(buf, 128))
    rows = []
    for row in range(2):
        t = list(buf[row * 64:(row + 1) * 64])
        cyc = t[20] - t[0]
        ns = t[31] - t[30]
        ghz = cyc / max(ns, 1)
        rec = {"cta": "first" if row == 0 else "last pair", "kernel_us": round(ns / 1e3, 2), "sm_ghz": round(ghz, 3)}
        for s, nm in sorted(SLOTS.items()):
            if t[s] >= t[0] and t[s] != 0 and t[s] <= t[20] + 10:
                rec[nm] = round((t[s] - t[0]) / ghz / 1e3, 2)
        rows.append(rec)
    return {"gemm": name, "M": M, "N": N, "K": K, "bn": bn, "event_us_median": round(sorted(ev)[len(ev) // 2], 1), "ctas": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="qkv,proj,fc1,fc2,fc2_dgrad,fc1_dgrad")
    ap.add_argument("--m", type=int, default=4864)
    ap.add_argument("--bn", type=int, default=0, help="force the N tile of the CTA-pair kernel (128 / 256)")
    a = ap.parse_args()
    sh = shapes(a.m)
    for name in a.shapes.split(","):
        kw = dict(sh[name])
        N, K = kw.pop("N"), kw.pop("K")
        print(json.dumps(run(name, a.m, N, K, bn=a.bn, **kw)), flush=True)


if __name__ == "__main__" and "--adapter" not in sys.argv:
    main()


def adapter_timeline(T=4864, E=768):
    """Phase stamps of CTA 0 of the fused AdapterModule kernel, forward and backward-input."""
    names = {0: "entry", 1: "prologue done", 2: "grid dependency", 3: "first operands landed",
             4: "phase 1 MMAs issued", 5: "H ready (warp 0)", 6: "G written (warp 0)", 7: "phase 3 MMAs issued",
             8: "Y ready (warp 0)", 10: "chunk 0 done", 11: "chunk 1 done", 12: "chunk 2 done", 13: "chunk 3 done",
             14: "stores drained", 15: "end"}
    lib = L.lib()
    x = torch.randn(T, E, device="cuda")
    xb = x.bfloat16()
    W1 = (torch.randn(64, E, device="cuda") / E ** 0.5).bfloat16()
    W2 = (torch.randn(E, 64, device="cuda") / 8).bfloat16()
    b1, b2 = torch.randn(64, device="cuda"), torch.randn(E, device="cuda")
    alpha = torch.tensor([0.3], device="cuda")
    g = torch.empty(T, 64, device="cuda", dtype=torch.bfloat16)
    d = torch.empty_like(g)
    y = torch.empty(T, E, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for which in ("forward", "backward"):
        ev = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if which == "forward":
                L.check(lib.fervit_adapter_forward(xb.data_ptr(), x.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(),
                                                   b2.data_ptr(), alpha.data_ptr(), T, E, g.data_ptr(), d.data_ptr(),
                                                   y.data_ptr(), st))
            else:
                L.check(lib.fervit_adapter_backward_input(xb.data_ptr(), x.data_ptr(), W2.t().contiguous().data_ptr(),
                                                          W1.t().contiguous().data_ptr(), alpha.data_ptr(), d.data_ptr(),
                                                          T, E, g.data_ptr(), y.data_ptr(), xb.data_ptr(), st))
            e1.record()
            torch.cuda.synchronize()
            ev.append(e0.elapsed_time(e1) * 1e3)
        buf = (ctypes.c_ulonglong * 32)()
        L.check(lib.fervit_debug_adapter_timeline(buf, 32))
        t = list(buf)
        ghz = (t[15] - t[0]) / max(t[31] - t[30], 1)
        rec = {"kernel": "adapter " + which, "event_us_median": round(sorted(ev)[2], 1),
               "kernel_us": round((t[31] - t[30]) / 1e3, 2), "sm_ghz": round(ghz, 3)}
        for s_, nm in sorted(names.items()):
            if t[s_] >= t[0] and t[s_] <= t[15] + 10:
                rec[nm] = round((t[s_] - t[0]) / ghz / 1e3, 2)
        out.append(rec)
    return out


if "--adapter" in sys.argv:
    for r in adapter_timeline():
        print(json.dumps(r), flush=True)
