"""GPU: operator-level parity of the C-ABI kernels against elementary torch arithmetic on the same inputs.

Tolerances: fp32 kernels 1e-5 norm-wise relative (BASELINE.json fp32 gate is 1e-4 end to end); bf16 kernels are
compared with a reference computed in fp32 from the SAME bf16-rounded inputs, so only accumulation order and the
bf16 rounding of the output differ: 4e-3 norm-wise (bf16 has 8 mantissa bits: 2^-9 = 2e-3 per rounding).
"""
import ctypes as C
import json
import math
import os

import pytest
import torch

from tests.util import relerr

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, **vals):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_metrics.jsonl"), "a") as fh:
        fh.write(json.dumps({"test": name, **vals}) + "\n")


@pytest.fixture(scope="module")
def lib():
    from fer_vit_b200 import _lib as L
    assert torch.cuda.is_available()
    return L


def st():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return t.data_ptr() if t is not None else None


def gelu(x):
    return 0.5 * x * (1 + torch.erf(x / math.sqrt(2)))


# ------------------------------------------------------------------------------------------------ GEMM
def run_linear(L, dtype, x, W, bias, residual, act, want_out, want_f32, want_pre, force_bn=0):
    M, K = x.shape
    N = W.shape[0]
    tdt = torch.float32 if dtype == L.F32 else torch.bfloat16
    out = torch.full((M, N), float("nan"), dtype=tdt, device="cuda") if want_out else None
    of = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda") if want_f32 else None
    pre = torch.full((M, N), float("nan"), dtype=tdt, device="cuda") if want_pre else None
    L.check(L.lib().fervit_linear_forward(dtype, x.data_ptr(), W.data_ptr(), ptr(bias), ptr(residual), M, N, K, act,
                                          ptr(out), ptr(of), ptr(pre), force_bn, st()))
    torch.cuda.synchronize()
    return out, of, pre


@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (200, 192, 100), (608, 1536, 512), (1000, 64, 2048)])
def test_gemm_fp32_simt(lib, M, N, K):
    L = lib
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g)
    out, of, pre = run_linear(L, L.F32, x, W, b, r, L.ACT_GELU, True, True, True)
    u = (x.double() @ W.double().t() + b.double())
    ref = gelu(u) + r.double()
    e1, e2 = relerr(of, ref), relerr(pre, u)
    record("gemm_fp32_simt", M=M, N=N, K=K, err_out=e1, err_pre=e2)
    assert e1 < 1e-5 and e2 < 1e-5 and torch.equal(out, of)


TC_SHAPES = [
    # M, N, K, force_bn
    (128, 64, 64, 64), (128, 128, 64, 128), (128, 256, 64, 256),      # one tile, one k-block
    (128, 128, 256, 128),                                              # k loop / descriptor K-advance
    (384, 256, 512, 0),                                                # several tiles
    (300, 192, 136, 0),                                                # ragged M (TMA zero fill), N % BN != 0, K tail
    (4864, 2304, 768, 0), (4864, 768, 768, 0), (4864, 3072, 768, 0), (4864, 768, 3072, 0),   # ViT-B @ B=256
    (4864, 64, 768, 0), (4864, 768, 64, 0),                            # adapter
    (4608, 768, 512, 0),                                               # token projection
    (19 * 7, 768, 768, 0),                                             # tiny batch
]


@pytest.mark.parametrize("M,N,K,bn", TC_SHAPES)
def test_gemm_bf16_tcgen05(lib, M, N, K, bn):
    L = lib
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g)
    # plain
    out, of, _ = run_linear(L, L.BF16, x, W, None, None, L.ACT_NONE, True, True, False, bn)
    ref = x.float() @ W.float().t()
    e_plain = relerr(of, ref)
    # full epilogue
    out2, of2, pre2 = run_linear(L, L.BF16, x, W, b, r, L.ACT_GELU, True, True, True, bn)
    u = ref + b
    ref2 = gelu(u) + r
    e_full, e_pre, e_bf = relerr(of2, ref2), relerr(pre2.float(), u), relerr(out2.float(), ref2)
    record("gemm_bf16_tcgen05", M=M, N=N, K=K, bn=bn, err_plain=e_plain, err_full=e_full, err_pre=e_pre, err_bf16=e_bf)
    assert not torch.isnan(of).any() and not torch.isnan(of2).any()
    assert e_plain < 2e-5, "fp32-accumulated output must match an fp32 matmul of the same bf16 inputs"
    assert e_full < 2e-5 and e_pre < 4e-3 and e_bf < 4e-3
    assert relerr(out.float(), ref) < 4e-3


TC2_SHAPES = [
    # M, N, K, force_bn  (CTA-pair kernel: N % 32 == 0, K % 8 == 0)
    (128, 64, 64, 0), (256, 256, 64, 0), (256, 128, 64, 256),         # one pair tile; padded second CTA / N half
    (300, 192, 136, 0),                                                # odd number of 128-row blocks, K tail
    (384, 256, 512, 128), (1000, 320, 264, 0),                         # N % BN != 0
    (4864, 2304, 768, 0), (4864, 768, 768, 0), (4864, 3072, 768, 0), (4864, 768, 3072, 0),   # ViT-B @ B=256
    (4864, 3072, 768, 128), (4864, 64, 768, 0), (4864, 768, 64, 0),
    (19 * 7, 768, 768, 0), (9728, 2304, 768, 0),
]


def drelu(x):
    return (x > 0).double()


def dgelu(x):
    return 0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)


@pytest.mark.parametrize("M,N,K,bn", TC2_SHAPES)
def test_gemm_bf16_tcgen05_pair(lib, M, N, K, bn):
    """CTA-pair tcgen05 kernel (gemm_tc2.cu): every epilogue family against fp64 math on the same bf16 inputs, and
    bit-for-bit against the single-CTA kernel where both implement the epilogue."""
    L = lib
    g = torch.Generator(device="cuda").manual_seed(M * 5 + N * 11 + K)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g)
    acc = x.double() @ W.double().t()
    errs = {}
    # (a) bias, bf16 output only
    out, _, _ = run_linear(L, L.BF16, x, W, b, None, L.ACT_NONE, True, False, False, bn)
    errs["bias_bf16"] = relerr(out.float(), acc + b)
    assert errs["bias_bf16"] < 4e-3
    # (b) bias + fp32 residual -> fp32 and bf16 outputs; fp32 output without residual
    out, of, _ = run_linear(L, L.BF16, x, W, b, r, L.ACT_NONE, True, True, False, bn)
    errs["res_f32"] = relerr(of, acc + b + r)
    assert errs["res_f32"] < 2e-5 and relerr(out.float(), acc + b + r) < 4e-3
    _, of1, _ = run_linear(L, L.BF16, x, W, b, r, L.ACT_NONE, False, True, False, -256)
    assert torch.equal(of, of1), "CTA-pair and single-CTA kernels must agree bit for bit (same fp32 accumulation order)"
    _, of, _ = run_linear(L, L.BF16, x, W, None, None, L.ACT_NONE, False, True, False, bn)
    errs["plain_f32"] = relerr(of, acc)
    assert errs["plain_f32"] < 2e-5
    # (c) forward activation with the pre-activation saved
    for act, fn in ((L.ACT_GELU, gelu), (L.ACT_RELU, torch.relu)):
        out, _, pre = run_linear(L, L.BF16, x, W, b, None, act, True, False, True, bn)
        u = acc + b
        e_o, e_p = relerr(out.float(), fn(u)), relerr(pre.float(), u)
        errs[f"act{act}"] = e_o
        assert e_o < 4e-3 and e_p < 4e-3
    # (d) input gradient times the activation derivative: GEMM view rows M, columns N, reduction K
    aux = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    for act, dfn in ((L.ACT_GELU, dgelu), (L.ACT_RELU, drelu)):
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        L.check(L.lib().fervit_linear_dgrad(L.BF16, x.data_ptr(), W.data_ptr(), aux.data_ptr(), None, M, K, N, act,
                                            out.data_ptr(), None, bn, st()))
        torch.cuda.synchronize()
        e = relerr(out.float(), acc * dfn(aux.double()))
        errs[f"dact{act}"] = e
        assert e < 4e-3
    record("gemm_bf16_tcgen05_pair", M=M, N=N, K=K, bn=bn, **errs)


@pytest.mark.parametrize("M,N,K", [(4864, 768, 3072), (4864, 768, 2304), (4700, 768, 3072), (9728, 768, 3072),
                                   (4864, 2304, 1024), (4864, 768, 1024), (19 * 64, 768, 3072)])
def test_gemm_bf16_tcgen05_stream_k(lib, M, N, K):
    """Stream-K over the under-filled last round of the CTA-pair GEMM (57 tiles on 74 pairs at batch 256; 74 + 40 at
    batch 512; 148 + 23 for the QKV shape; a 15-tile problem spread over all pairs): bf16 output, fp32 residual output
    (single-tile aliased staging) and dgrad, against fp64 math on the same inputs; deterministic and re-armed (two
    launches bit-identical); within fp32 rounding of the same GEMM without stream-K."""
    L = lib
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 7 * K)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g)
    acc = x.double() @ W.double().t()
    base_bf, base_f32, _ = run_linear(L, L.BF16, x, W, b, r, L.ACT_NONE, True, True, False, 0)     # scratch unset
    nbytes = int(L.lib().fervit_gemm_scratch_bytes())
    scratch = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    L.check(L.lib().fervit_set_gemm_scratch(scratch.data_ptr(), nbytes))
    try:
        out1, of1, _ = run_linear(L, L.BF16, x, W, b, r, L.ACT_NONE, True, True, False, 0)
        out2, of2, _ = run_linear(L, L.BF16, x, W, b, r, L.ACT_NONE, True, True, False, 0)
        ob, _, _ = run_linear(L, L.BF16, x, W, b, None, L.ACT_NONE, True, False, False, 0)
        dg = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        L.check(L.lib().fervit_linear_dgrad(L.BF16, x.data_ptr(), W.data_ptr(), None, None, M, K, N, L.ACT_NONE,
                                            dg.data_ptr(), None, 0, st()))
        torch.cuda.synchronize()
        assert int(scratch[:4096].view(torch.int32).abs().sum()) == 0, "arrival flags must be back at zero"
    finally:
        L.check(L.lib().fervit_set_gemm_scratch(None, 0))
    e = {"f32": relerr(of1, acc + b + r), "bf16": relerr(ob.float(), acc + b), "dgrad": relerr(dg.float(), acc),
         "vs_plain": relerr(of1, base_f32)}
    record("gemm_bf16_tcgen05_stream_k", M=M, N=N, K=K, **e)
    assert torch.equal(of1, of2) and torch.equal(out1, out2)
    assert e["f32"] < 2e-5 and e["bf16"] < 4e-3 and e["dgrad"] < 4e-3 and e["vs_plain"] < 2e-6
    assert relerr(out1.float(), acc + b + r) < 4e-3 and relerr(base_bf.float(), acc + b + r) < 4e-3


@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
@pytest.mark.parametrize("M,N,K", [(256, 64, 64), (4864, 64, 768), (4864, 768, 64), (608, 1536, 512),
                                   (1000, 128, 208), (4608, 768, 512), (4864, 768, 3072),
                                   # CTA-pair kernel (gemm_wgrad2.cu): N, K >= 256; ragged token tail, ragged 256-tiles,
                                   # one unit per pair and several, the configs' real shapes
                                   (9728, 2048, 512), (12608, 512, 2048), (2100, 320, 256), (2048, 256, 448),
                                   (9728, 1536, 512), (12608, 512, 768)])
def test_linear_wgrad(lib, dtype_name, M, N, K):
    """dW[N,K] = alpha * dY^T X over M token rows: MN-major tcgen05 operands (bf16) / strided fp32 GEMM, split-K."""
    L = lib
    dtype = L.F32 if dtype_name == "fp32" else L.BF16
    tdt = torch.float32 if dtype == L.F32 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(M + 13 * N + K)
    dY = torch.randn(M, N, device="cuda", generator=g).to(tdt)
    X = torch.randn(M, K, device="cuda", generator=g).to(tdt)
    dW = torch.full((N, K), float("nan"), device="cuda")
    scratch = torch.empty(int(L.lib().fervit_linear_wgrad_scratch_floats(M, N, K)), device="cuda")
    L.check(L.lib().fervit_linear_wgrad(dtype, dY.data_ptr(), X.data_ptr(), M, N, K, 0.5, dW.data_ptr(),
                                        scratch.data_ptr(), st()))
    torch.cuda.synchronize()
    ref = 0.5 * (dY.double().t() @ X.double())
    e = relerr(dW, ref)
    record("linear_wgrad", dtype=dtype_name, M=M, N=N, K=K, err=e)
    assert e < 2e-5


@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
@pytest.mark.parametrize("M,N,K", [(9728, 2048, 512), (12608, 512, 2048), (2100, 320, 256), (4864, 768, 512),
                                   (608, 1536, 512), (4864, 64, 768)])
def test_linear_wgrad_with_bias_gradient(lib, dtype_name, M, N, K):
    """dW and db = colsum(dY) from one call: the CTA-pair kernel's in-kernel column sums (bf16, N, K >= 256, M >= 2048)
    and the two-pass fallback give the same answers as fp64 math; the fused form is deterministic (two calls agree bit
    for bit)."""
    L = lib
    dtype = L.F32 if dtype_name == "fp32" else L.BF16
    tdt = torch.float32 if dtype == L.F32 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(M + 7 * N + K)
    dY = (torch.randn(M, N, device="cuda", generator=g) + 0.1).to(tdt)
    X = torch.randn(M, K, device="cuda", generator=g).to(tdt)
    scratch = torch.empty(int(L.lib().fervit_linear_wgrad_bias_scratch_floats(M, N, K)), device="cuda")
    outs = []
    for _ in range(2):
        dW = torch.full((N, K), float("nan"), device="cuda")
        db = torch.full((N,), float("nan"), device="cuda")
        L.check(L.lib().fervit_linear_wgrad_bias(dtype, dY.data_ptr(), X.data_ptr(), M, N, K, 0.5, dW.data_ptr(),
                                                 db.data_ptr(), scratch.data_ptr(), st()))
        torch.cuda.synchronize()
        outs.append((dW, db))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    e_w = relerr(outs[0][0], 0.5 * (dY.double().t() @ X.double()))
    e_b = relerr(outs[0][1], dY.double().sum(0))
    record("linear_wgrad_bias", dtype=dtype_name, M=M, N=N, K=K, err_w=e_w, err_b=e_b)
    assert e_w < 2e-5 and e_b < 2e-5


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,E", [(37, 64), (608, 512), (4864, 768), (130, 192)])
def test_layernorm(lib, dtype_name, rows, E):
    L = lib
    dtype = L.F32 if dtype_name == "fp32" else L.BF16
    tdt = torch.float32 if dtype == L.F32 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(rows + E)
    x = torch.randn(rows, E, device="cuda", generator=g) * 2 + 0.5
    gam = 1 + 0.2 * torch.randn(E, device="cuda", generator=g)
    bet = 0.2 * torch.randn(E, device="cuda", generator=g)
    dy = torch.randn(rows, E, device="cuda", generator=g).to(tdt)
    dres = torch.randn(rows, E, device="cuda", generator=g)
    yf = torch.empty(rows, E, device="cuda")
    ya = torch.empty(rows, E, device="cuda", dtype=tdt)
    mean = torch.empty(rows, device="cuda"); rstd = torch.empty(rows, device="cuda")
    L.check(L.lib().fervit_layernorm_forward(dtype, x.data_ptr(), gam.data_ptr(), bet.data_ptr(), 1e-6, rows, E,
                                             yf.data_ptr(), ya.data_ptr(), mean.data_ptr(), rstd.data_ptr(), st()))
    xd = x.double().requires_grad_(True); gd = gam.double().requires_grad_(True); bd = bet.double().requires_grad_(True)
    mu = xd.mean(-1, keepdim=True); var = ((xd - mu) ** 2).mean(-1, keepdim=True)
    ref = (xd - mu) / torch.sqrt(var + 1e-6) * gd + bd
    dxr, dgr, dbr = torch.autograd.grad(ref, [xd, gd, bd], dy.double())
    dxf = torch.empty(rows, E, device="cuda"); dxa = torch.empty(rows, E, device="cuda", dtype=tdt)
    dgam = torch.empty(E, device="cuda"); dbet = torch.empty(E, device="cuda")
    scratch = torch.empty(int(L.lib().fervit_layernorm_scratch_floats(rows, E)), device="cuda")
    L.check(L.lib().fervit_layernorm_backward(dtype, dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                              gam.data_ptr(), dres.data_ptr(), rows, E, dxf.data_ptr(), dxa.data_ptr(),
                                              scratch.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), st()))
    torch.cuda.synchronize()
    errs = dict(y=relerr(yf, ref), y_act=relerr(ya.float(), ref), dx=relerr(dxf, dxr + dres.double()),
                dgamma=relerr(dgam, dgr), dbeta=relerr(dbet, dbr))
    record("layernorm", dtype=dtype_name, rows=rows, E=E, **errs)
    tol_act = 1e-5 if dtype == L.F32 else 4e-3
    assert errs["y"] < 1e-5 and errs["y_act"] < tol_act and errs["dx"] < 1e-5
    assert errs["dgamma"] < 1e-5 and errs["dbeta"] < 1e-5


# ------------------------------------------------------------------------------------------------ attention
def attn_ref(qkv, B, S, H, hd, mask=None):
    E = H * hd
    q, k, v = qkv.reshape(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    p = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    if mask is not None:
        p = p * mask
    return (p @ v).permute(0, 2, 1, 3).reshape(B * S, E)


@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
@pytest.mark.parametrize("B,S,H,hd,pdrop", [(5, 19, 2, 32, 0.0), (33, 19, 12, 64, 0.0), (3, 37, 6, 64, 0.0),
                                            (2, 197, 8, 64, 0.0), (2, 197, 8, 48, 0.0), (4, 19, 8, 64, 0.1),
                                            (2, 5, 2, 32, 0.1), (2, 197, 4, 64, 0.1), (3, 100, 2, 32, 0.1),
                                            (2, 256, 2, 64, 0.0), (1, 33, 3, 48, 0.0), (2, 130, 2, 64, 0.1),
                                            (2, 256, 2, 64, 0.1), (2, 209, 2, 48, 0.1),   # > 208 keys: 8 warps, two rounds
                                            # persistent short-sequence kernel: several problems per warp pair (the
                                            # double-buffered loop), a ragged last CTA, one 16-row tile only, S = 32
                                            (256, 19, 12, 64, 0.0), (131, 19, 3, 64, 0.1), (300, 12, 6, 32, 0.1),
                                            (97, 32, 5, 48, 0.0), (64, 16, 4, 64, 0.1), (1, 1, 1, 64, 0.0)])
def test_attention(lib, dtype_name, B, S, H, hd, pdrop):
    L = lib
    dtype = L.F32 if dtype_name == "fp32" else L.BF16
    if dtype == L.F32 and S > 197:
        pytest.skip("the fp32 (CUDA-core) kernel keeps a whole head in shared memory: S <= 197 at head dim 64")
    tdt = torch.float32 if dtype == L.F32 else torch.bfloat16
    E = H * hd
    g = torch.Generator(device="cuda").manual_seed(B + S + H + hd)
    qkv = torch.randn(B * S, 3 * E, device="cuda", generator=g).to(tdt)
    dout = torch.randn(B * S, E, device="cuda", generator=g).to(tdt)
    out = torch.empty(B * S, E, device="cuda", dtype=tdt)
    lse = torch.empty(B * H * S, device="cuda")
    dqkv = torch.empty(B * S, 3 * E, device="cuda", dtype=tdt)
    seed, site = 1234, 8
    L.check(L.lib().fervit_attention_forward(dtype, qkv.data_ptr(), B, S, H, hd, pdrop, seed, site, out.data_ptr(),
                                             lse.data_ptr(), st()))
    L.check(L.lib().fervit_attention_backward(dtype, qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                              B, S, H, hd, pdrop, seed, site, dqkv.data_ptr(), st()))
    mask = None
    if pdrop > 0:
        mask = torch.empty(B * H * S * S, device="cuda")
        L.check(L.lib().fervit_dropout_mask(mask.data_ptr(), mask.numel(), pdrop, seed, site, st()))
        mask = mask.reshape(B, H, S, S).double()
        keep = (mask > 0).double().mean().item()
        assert abs(keep - (1 - pdrop)) < 0.05 and torch.all((mask == 0) | ((mask - 1 / (1 - pdrop)).abs() < 1e-6))
    torch.cuda.synchronize()
    qd = qkv.double().requires_grad_(True)
    ref = attn_ref(qd, B, S, H, hd, mask)
    dref, = torch.autograd.grad(ref, qd, dout.double())
    e_o, e_d = relerr(out.float(), ref), relerr(dqkv.float(), dref)
    record("attention", dtype=dtype_name, B=B, S=S, H=H, hd=hd, p=pdrop, err_out=e_o, err_dqkv=e_d)
    tol = 2e-5 if dtype == L.F32 else 6e-3
    assert e_o < tol and e_d < tol


# ------------------------------------------------------------------------------------------------ loss
@pytest.mark.parametrize("B", [1, 6, 256, 4096])
@pytest.mark.parametrize("weighted,eps", [(False, 0.0), (True, 0.0), (False, 0.1), (True, 0.1)])
def test_cross_entropy(lib, B, weighted, eps):
    import fer_vit_b200 as fv
    g = torch.Generator(device="cuda").manual_seed(B)
    z = (3 * torch.randn(B, 7, device="cuda", generator=g)).requires_grad_(True)
    y = torch.randint(0, 7, (B,), device="cuda", generator=g)
    w = (torch.rand(7, device="cuda", generator=g) + 0.5) if weighted else None
    loss = fv.cross_entropy(z, y, w, eps)
    gz, = torch.autograd.grad(loss, z)
    zr = z.detach().double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(zr, y, weight=w.double() if weighted else None, label_smoothing=eps)
    gr, = torch.autograd.grad(ref, zr)
    # the loss of a tiny batch can be ~1e-4: gate its error relative to max(1, |loss|)
    e_l, e_g = abs(loss.item() - ref.item()) / max(1.0, abs(ref.item())), relerr(gz, gr)
    record("cross_entropy", B=B, weighted=weighted, eps=eps, err_loss=e_l, err_grad=e_g)
    assert e_l < 1e-5 and e_g < 1e-5


# ------------------------------------------------------------------------------------------------ pre-modules
@pytest.mark.parametrize("flags", [(1, 0, 0, 0), (0, 1, 0, 0), (0, 1, 1, 0), (0, 0, 0, 1), (1, 1, 1, 1), (1, 1, 0, 1)])
@pytest.mark.parametrize("B,D", [(3, 64), (37, 512)])
def test_premodules(lib, flags, B, D):
    import fer_vit_b200 as fv
    from oracle import reference_math as R
    use_spe, use_lwn, use_res, use_leam = flags
    torch.manual_seed(sum(flags) + B)
    Ln = 18
    x = (torch.randn(B, Ln, D) * 0.8 + 0.2).cuda().requires_grad_(True)
    mods = {}
    if use_spe:
        mods["spe"] = fv.SemanticPE(D, Ln).cuda()
    if use_lwn:
        mods["lwn"] = fv.LayerWiseNorm(Ln, D, use_residual=bool(use_res)).cuda()
        with torch.no_grad():
            for n in mods["lwn"].norms:
                n.weight.add_(0.2 * torch.randn_like(n.weight)); n.bias.add_(0.2 * torch.randn_like(n.bias))
            if use_res:
                mods["lwn"].gate.add_(4.5 + torch.randn_like(mods["lwn"].gate))
    if use_leam:
        mods["leam"] = fv.LEAM(Ln).cuda()
    h = x
    for k in ("spe", "lwn", "leam"):
        if k in mods:
            h = mods[k](h)
    dy = torch.randn_like(h)
    params = [p for m in mods.values() for p in m.parameters()]
    names = [f"{k}.{n}" for k, m in mods.items() for n, _ in m.named_parameters()]
    grads = torch.autograd.grad(h, [x] + params, dy)
    sd = {}
    for k, m in mods.items():
        for n, v in m.state_dict().items():
            t = v.detach().cpu()
            sd[f"{k}.{n}"] = t.double().requires_grad_(True) if t.is_floating_point() else t
    xr = x.detach().cpu().double().requires_grad_(True)
    ref = R.pre_modules(xr, sd, bool(use_spe), bool(use_lwn), bool(use_res), bool(use_leam))
    rg = torch.autograd.grad(ref, [xr] + [sd[n] for n in names], dy.cpu().double())
    errs = {"y": relerr(h, ref), "dx": relerr(grads[0], rg[0])}
    for n, a, b in zip(names, grads[1:], rg[1:]):
        errs["d" + n] = relerr(a, b)
    record("premodules", flags=list(flags), B=B, D=D, worst=max(errs.values()))
    assert max(errs.values()) < 2e-5, errs


# ------------------------------------------------------------------------------------------------ optimizer
@pytest.mark.parametrize("max_norm", [None, 0.5])
def test_fused_adamw_matches_torch(lib, max_norm):
    """fer_vit_b200.FusedAdamW against torch.optim.AdamW (+ clip_grad_norm_) on two hyper-parameter groups, ragged
    tensor sizes, four steps with fresh gradients; also the state_dict round trip."""
    import fer_vit_b200 as fv
    g = torch.Generator(device="cuda").manual_seed(7)
    shapes = [(768, 64), (64,), (1,), (5, 4097), (19, 768), (7, 768)]
    ref = [torch.nn.Parameter(torch.randn(*s, device="cuda", generator=g)) for s in shapes]
    our = [torch.nn.Parameter(p.detach().clone()) for p in ref]

    def groups(ps):
        return [dict(params=ps[:3], lr=3e-3, weight_decay=0.05), dict(params=ps[3:], lr=1e-2, betas=(0.8, 0.95))]
    o_ref = torch.optim.AdamW(groups(ref), lr=1e-3, weight_decay=0.01)
    o_our = fv.FusedAdamW(groups(our), lr=1e-3, weight_decay=0.01, max_grad_norm=max_norm)
    for step in range(4):
        if step == 2:   # a scheduler-style change
            for o in (o_ref, o_our):
                o.param_groups[0]["lr"] = 1e-3
            sd = o_our.state_dict()
            o_our = fv.FusedAdamW(groups(our), lr=1e-3, weight_decay=0.01, max_grad_norm=max_norm)
            o_our.load_state_dict(sd)
        for a, b in zip(ref, our):
            gr = torch.randn(a.shape, device="cuda", generator=g) * (1 + step)
            a.grad = gr.clone()
            b.grad = gr.clone()
        if max_norm:
            tn = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        o_ref.step()
        o_our.step()
        if max_norm:
            assert abs(float(o_our.last_total_norm) - float(tn)) <= 1e-5 * float(tn)
            for a, b in zip(ref, our):
                assert relerr(b.grad, a.grad) < 1e-6, "gradients are left clipped, as clip_grad_norm_ leaves them"
    torch.cuda.synchronize()
    errs = [relerr(b, a) for a, b in zip(ref, our)]
    record("fused_adamw", max_norm=max_norm or 0.0, err=max(errs))
    assert max(errs) < 2e-6, errs


def test_fused_adamw_step_counter_rules(lib):
    """One shared device step counter: a parameter that joins after the optimizer has stepped is refused
    (it would be bias-corrected with t instead of 1), and load_state_dict() into an optimizer that has ALREADY stepped
    re-seeds the counter from the loaded state, so the next update is step t+1 of the loaded run, as in torch."""
    import fer_vit_b200 as fv
    g = torch.Generator(device="cuda").manual_seed(11)
    mk = lambda: [torch.nn.Parameter(torch.randn(33, 17, device="cuda", generator=g)) for _ in range(2)]
    ref, our = mk(), None
    our = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.01)
    o_our = fv.FusedAdamW(our, lr=1e-2, weight_decay=0.01)
    grads = [[torch.randn(33, 17, device="cuda", generator=g) for _ in range(2)] for _ in range(5)]

    def run(o, ps, k):
        for p_, gr in zip(ps, grads[k]):
            p_.grad = gr.clone()
        o.step()
    for k in range(3):
        run(o_ref, ref, k); run(o_our, our, k)
    import copy
    sd = copy.deepcopy(o_our.state_dict())     # a snapshot (state_dict() itself returns references, as in torch)
    run(o_our, our, 3)                         # the optimizer moves on (counter = 4) ...
    with torch.no_grad():                      # ... then is rolled back to the saved run
        for a, b in zip(ref, our):
            b.copy_(a)
    o_our.load_state_dict(sd)                  # moments and step = 3 restored into the SAME optimizer object
    run(o_ref, ref, 3); run(o_our, our, 3)
    assert max(relerr(b, a) for a, b in zip(ref, our)) < 2e-6
    # late joiner
    late = torch.nn.Parameter(torch.randn(5, device="cuda", generator=g))
    o_our.add_param_group({"params": [late]})
    for p_ in our:
        p_.grad = torch.zeros_like(p_)
    late.grad = torch.ones_like(late)
    with pytest.raises(RuntimeError, match="without optimizer state"):
        o_our.step()


# ------------------------------------------------------------------------------------------------ fused adapter
@pytest.mark.parametrize("T,E", [(128, 256), (300, 768), (4864, 768), (19 * 7, 512)])
def test_adapter_fused_tcgen05(lib, T, E):
    """AdapterModule forward and input-gradient as one tensor-core kernel each (adapter_tc.cu) against fp64 math on the
    same bf16 operands (hybrid_latent_vit.py:249-265)."""
    L = lib
    g_ = torch.Generator(device="cuda").manual_seed(T + E)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g_)
    x = rn(T, E)
    xb = x.bfloat16()
    W1 = (rn(64, E) / math.sqrt(E)).bfloat16()
    W2 = (rn(E, 64) / 8).bfloat16()
    b1, b2 = 0.1 * rn(64), 0.1 * rn(E)
    alpha = torch.tensor([0.37], device="cuda")
    gq = torch.full((T, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    dq = torch.full_like(gq, float("nan"))
    y = torch.full((T, E), float("nan"), device="cuda")
    L.check(L.lib().fervit_adapter_forward(xb.data_ptr(), x.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(),
                                           b2.data_ptr(), alpha.data_ptr(), T, E, gq.data_ptr(), dq.data_ptr(),
                                           y.data_ptr(), st()))
    torch.cuda.synchronize()
    u = xb.double() @ W1.double().t() + b1.double()
    g_ref, d_ref = gelu(u), dgelu(u)
    y_ref = x.double() + alpha.double() * (gq.double() @ W2.double().t() + b2.double())   # from the kernel's own bf16 g
    e_g, e_d, e_y = relerr(gq.float(), g_ref), relerr(dq.float(), d_ref), relerr(y, y_ref)
    assert e_g < 4e-3 and e_d < 4e-3, (e_g, e_d)
    assert e_y < 2e-6, e_y
    # backward: du = alpha * (dy W2) * d ; dx = dy + du W1
    dy = rn(T, E)
    dyb = dy.bfloat16()
    W2t, W1t = W2.t().contiguous(), W1.t().contiguous()
    du = torch.full((T, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    dx = torch.full((T, E), float("nan"), device="cuda")
    dxb = torch.full((T, E), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.lib().fervit_adapter_backward_input(dyb.data_ptr(), dy.data_ptr(), W2t.data_ptr(), W1t.data_ptr(),
                                                  alpha.data_ptr(), dq.data_ptr(), T, E, du.data_ptr(), dx.data_ptr(),
                                                  dxb.data_ptr(), st()))
    torch.cuda.synchronize()
    du_ref = alpha.double() * (dyb.double() @ W2.double()) * dq.double()
    dx_ref = dy.double() + du.double() @ W1.double()
    e_du, e_dx, e_dxb = relerr(du.float(), du_ref), relerr(dx, dx_ref), relerr(dxb.float(), dx_ref)
    record("adapter_fused", T=T, E=E, err_g=e_g, err_d=e_d, err_y=e_y, err_du=e_du, err_dx=e_dx)
    assert e_du < 4e-3 and e_dx < 2e-6 and e_dxb < 4e-3, (e_du, e_dx, e_dxb)
