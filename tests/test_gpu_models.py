"""GPU: whole-model parity of the drop-in classes (through the C-ABI plan) against
  (a) the golden vectors produced by the UNMODIFIED reference classes (tests/golden/*.npz), and
  (b) the oracle run live on the same seeded inputs at the reference's real widths,
plus size-independent properties at BASELINE.json's full batch sizes.

Gates (BASELINE.json north_star): logits and every gradient within 1e-4 (fp32 mode) / 2e-2 (bf16 mode) norm-wise
relative error, identical top-1 predictions.
"""
import os

import pytest
import torch

from tests.test_gpu_ops import record
from tests.util import FIXTURES, build_model, load_golden, oracle_forward, relerr

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}


def step(model, x, y, weight=None, smoothing=0.0):
    import fer_vit_b200 as fv
    model.zero_grad(set_to_none=True)
    logits = model(x)
    loss = fv.cross_entropy(logits, y, weight, smoothing)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return logits.detach(), loss.detach(), grads


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", FIXTURES)
def test_model_matches_reference_golden(name, precision):
    g = load_golden(name)
    model = build_model(name, precision)
    model.load_state_dict(g["sd"], strict=True)          # same keys / shapes as the reference checkpoint
    model = model.cuda()
    # the Hybrid fixtures were generated in eval() (its head Dropout(0.1) is hard-coded); the others in train()
    model.train(not name.startswith("hybrid"))
    w = g["class_weight"].cuda() if g["class_weight"] is not None else None
    logits, loss, grads = step(model, g["x"].cuda(), g["y"].cuda(), w, g["label_smoothing"])
    tol = TOL[precision]
    e_l = relerr(logits, g["logits"])
    errs = {k: relerr(grads[k], g["grad"][k]) for k in g["grad"]}
    worst = max(errs, key=errs.get)
    record("model_vs_golden", fixture=name, precision=precision, err_logits=e_l,
           err_loss=abs(loss.item() - float(g["loss"])), worst_grad=errs[worst], worst_key=worst)
    assert set(grads) == set(g["grad"]), set(grads) ^ set(g["grad"])
    assert e_l < tol, e_l
    assert errs[worst] < tol, (worst, errs[worst])
    assert abs(loss.item() - float(g["loss"])) < tol * max(1.0, abs(float(g["loss"])))
    if precision == "fp32":
        assert torch.equal(logits.argmax(-1).cpu(), g["logits"].argmax(-1))


def _oracle_step(fwd, sd, x, y, masks=None, weight=None, smoothing=0.0):
    """fp64 oracle on the CPU: its own round-off is negligible against both gates."""
    from oracle import reference_math as R
    sd = {k: (v.detach().double().requires_grad_(v.requires_grad) if v.is_floating_point() else v) for k, v in sd.items()}
    logits = fwd(sd, x.double(), masks)
    loss = R.cross_entropy(logits, y, weight.double() if weight is not None else None, smoothing)
    return logits.detach(), loss.detach(), R.grads_of(loss, sd)


def _group(grads):
    """The per-adapter scalars `adapters.{i}.alpha` are compared as ONE 12-vector: a single alpha's gradient is a sum
    of B*S*E signed terms that can cancel to ~0 (adapters.10 does at random init), where a per-scalar relative
    error is ill-conditioned in any finite precision."""
    out, alphas = {}, []
    for k in sorted(grads):
        if k.endswith(".alpha"):
            alphas.append(grads[k].reshape(-1).double().cpu())
        else:
            out[k] = grads[k]
    if alphas:
        out["adapters.*.alpha"] = torch.cat(alphas)
    return out


def _global_err(grads, ref):
    a = torch.cat([grads[k].reshape(-1).double().cpu() for k in sorted(ref)])
    b = torch.cat([ref[k].reshape(-1).double().cpu() for k in sorted(ref)])
    return float((a - b).norm() / b.norm())


def _compare(tag, precision, got, ref, extra=None, baseline=None):
    """Gates (BASELINE.json north_star, un-widened): logits < tol; EVERY gradient tensor < tol; the whole gradient
    vector < tol; tol = 1e-4 (fp32 mode) / 2e-2 (bf16 mode). `baseline` (torch's own bf16 autocast run of the same
    parameters) is recorded for information only and gates nothing."""
    logits, loss, grads = got
    rl, rloss, rg = ref
    assert set(grads) == set(rg)
    grads, rg = _group(grads), _group(rg)
    e_l = relerr(logits, rl)
    errs = {k: relerr(grads[k], rg[k]) for k in rg}
    e_glob = _global_err(grads, rg)
    tol = TOL[precision]
    rec = {}
    if baseline is not None:
        bg = _group(baseline[2])
        berr = {k: relerr(bg[k], rg[k]) for k in rg}
        rec = {"torch_autocast_err_logits": relerr(baseline[0], rl), "torch_autocast_worst_grad": max(berr.values()),
               "torch_autocast_global_grad": _global_err(bg, rg)}
    worst = max(errs, key=errs.get)
    record(tag, precision=precision, err_logits=e_l, err_loss=abs(loss.item() - rloss.item()), worst_grad=errs[worst],
           worst_key=worst, global_grad=e_glob, **rec, **(extra or {}))
    assert e_l < tol, e_l
    assert errs[worst] < tol, (worst, errs[worst], tol)
    assert e_glob < tol, (e_glob, tol)
    # top-1 must agree wherever the reference's top-2 margin exceeds the logit tolerance
    top2 = rl.float().topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 4 * tol * rl.abs().max()
    assert torch.equal(logits.argmax(-1).cpu()[clear], rl.argmax(-1)[clear])


# ReLU models (LatentViT: nn.TransformerEncoderLayer's default activation, latent_vit.py:24-31). relu(u) = u * [u > 0]:
# a 0/1 SELECTOR times a linear map. The selector is discontinuous in u exactly like a dropout mask is in its random
# draw: a unit whose pre-activation lies within round-off of zero legitimately falls on either side, and one flipped
# unit moves a linear1 gradient by O(1/sqrt(T*F)) whatever the precision of everything else (bf16 operands flip
# ~0.3 % of the units; torch's own bf16 autocast shows the same 3-6e-2 gradient error against an fp64 reference).
# So, as for dropout, the oracle is handed the selector the kernels actually used (read back from the saved forward
# state through the C ABI, fervit_plan_saved_buffer) and the PLAIN gates apply to everything else; separately the
# selector itself is checked against the oracle's own: it may differ only where |u_ref| is within the precision mode's
# round-off of zero (SEL_BAND x rms(u_ref)), and only on a small fraction of the units.
SEL_BAND = {"fp32": 2e-4, "bf16": 1e-1}
SEL_FRAC = {"fp32": 1e-4, "bf16": 2e-2}


def _relu_selectors(model, depth, B, S):
    r = model.plan_runner()
    return {("relu", i): r.saved_activation(i, 0).reshape(B, S, -1).cpu().double() for i in range(depth)}


def _check_selectors(tag, precision, sel, trace):
    frac, band = 0.0, 0.0
    for (name, i), m in sel.items():
        u = trace[("ffn_pre", i)].double()
        ref = (u > 0).double()
        flips = m != ref
        frac = max(frac, float(flips.double().mean()))
        if flips.any():
            band = max(band, float(u[flips].abs().max() / u.pow(2).mean().sqrt()))
    record(tag + "_relu_selector", precision=precision, worst_layer_flip_fraction=frac, worst_flip_abs_u_over_rms=band)
    assert frac < SEL_FRAC[precision], frac
    assert band < SEL_BAND[precision], band


def _relu_reference(tag, precision, fwd, sd, x, y, sel, masks=None, weight=None, smoothing=0.0):
    """fp64 oracle step with the kernels' ReLU selectors injected (+ any dropout masks), after checking those
    selectors against the oracle's own pre-activations."""
    free = dict(masks or {})
    free["trace"] = {}
    with torch.no_grad():
        fwd({k: (v.detach().double() if v.is_floating_point() else v) for k, v in sd.items()}, x.double(), free)
    _check_selectors(tag, precision, sel, free["trace"])
    given = dict(masks or {})
    given.update(sel)
    return _oracle_step(fwd, sd, x, y, given, weight, smoothing)


def _torch_autocast_step(model, x, y, smoothing=0.0, pre=None):
    """The same parameters run through torch's own modules under bf16 autocast (cuBLAS + SDPA): the error floor of a
    bf16 implementation, NOT part of the product path."""
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        h = pre(x) if pre is not None else x
        bb = model.backbone if hasattr(model, "backbone") else model
        h = bb.input_proj(h)
        h = torch.cat([bb.cls_token.expand(h.shape[0], -1, -1), h], dim=1) + bb.pos_emb
        h = bb.transformer(h)
        logits = bb.mlp_head(h[:, 0]).float()
    loss = torch.nn.functional.cross_entropy(logits, y, label_smoothing=smoothing)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    return logits.detach(), loss.detach(), grads


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_hybrid_vit_base_adapter_vs_oracle(precision):
    """BASELINE config 3 shape: frozen ViT-B/16 blocks + Adapter(64) on 18x512 w+ tokens (reduced batch for the CPU oracle)."""
    import fer_vit_b200 as fv
    from oracle import baseline_models as BM
    from oracle import reference_math as R
    fv.set_default_precision(precision)
    sd = BM.hybrid_state_dict(seed=3)
    for i in range(12):
        sd[f"adapters.{i}.alpha"] = torch.ones(1) * (0.1 + 0.02 * i)
    model = fv.create_hybrid_latent_vit(model_size="base", use_pretrained=False, freeze_transformer=True,
                                        use_adapter=True, adapter_dim=64)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(11)
    B = 8
    # second input distribution of SURVEY.md 8d: non-zero mean, like real pSp latents (stresses LayerNorm)
    x = 0.5 * torch.randn(B, 18, 512, generator=g) + 0.3 * torch.randn(1, 18, 512, generator=g)
    y = torch.randint(0, 7, (B,), generator=g)
    BM.hybrid_trainable(sd)
    ref = _oracle_step(lambda s, xx, m: R.hybrid_forward(s, xx, 12, 12, True, m), sd, x, y)
    got = step(model, x.cuda(), y.cuda())
    assert sum(v.numel() for v in got[2].values()) == 1_605_907      # trainable set of SURVEY.md 8a row a7
    _compare("hybrid_vitb_adapter_vs_oracle", precision, got, ref, {"B": B})


@pytest.mark.gpu
def test_hybrid_grouped_adapter_gradients_vs_oracle():
    """T = 64 x 19 rows is a multiple of the GEMM k-block, so the bf16 plan keeps every block's dy / du and finishes
    the AdapterModule gradients of all 12 blocks with one split-K launch per weight and one column-sum launch per bias
    (plan.cu: ad_deferred); smaller batches (the test above) take the per-block form. Same oracle, same gates; the
    launch count tells the two forms apart."""
    import fer_vit_b200 as fv
    from fer_vit_b200 import _lib
    from oracle import baseline_models as BM
    from oracle import reference_math as R
    fv.set_default_precision("bf16")
    sd = BM.hybrid_state_dict(seed=3)
    for i in range(12):
        sd[f"adapters.{i}.alpha"] = torch.ones(1) * (0.1 + 0.02 * i)
    model = fv.create_hybrid_latent_vit(model_size="base", use_pretrained=False, freeze_transformer=True,
                                        use_adapter=True, adapter_dim=64)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(12)
    B = 64
    x = 0.5 * torch.randn(B, 18, 512, generator=g) + 0.3 * torch.randn(1, 18, 512, generator=g)
    y = torch.randint(0, 7, (B,), generator=g)
    BM.hybrid_trainable(sd)
    ref = _oracle_step(lambda s, xx, m: R.hybrid_forward(s, xx, 12, 12, True, m), sd, x, y)
    n0 = _lib.launch_count()
    got = step(model, x.cuda(), y.cuda())
    grouped = _lib.launch_count() - n0
    n0 = _lib.launch_count()
    step(model, x[:63].cuda(), y[:63].cuda())     # T = 63 x 19 rows: not a multiple of 64, per-block form
    per_block = _lib.launch_count() - n0
    if os.environ.get("FERVIT_ADAPTER_DEFER", "1") != "0":
        assert grouped < per_block, (grouped, per_block)   # 12 x 4 side launches became 4
    _compare("hybrid_vitb_grouped_adapter_grads_vs_oracle", "bf16", got, ref, {"B": B, "launches": grouped,
                                                                              "launches_per_block_form": per_block})
    # one adapter frozen in the middle of the grouped range: its gradients are absent, every other one is unchanged
    for k, p in model.named_parameters():
        if k.startswith("adapters.5."):
            p.requires_grad_(False)
    part = step(model, x.cuda(), y.cuda())[2]
    assert not any(k.startswith("adapters.5.") for k in part)
    assert set(part) == {k for k in got[2] if not k.startswith("adapters.5.")}
    for k, v in part.items():
        assert torch.equal(v, got[2][k]), k


@pytest.mark.gpu
def test_hybrid_full_finetune_2128_rows_vs_oracle():
    """`freeze_transformer=False` (train_hybrid_latent_vit.py: --no_freeze) at 112 samples = 2128 token rows, bf16: every
    block's weight AND bias gradients come from the CTA-pair kernel (gemm_wgrad2.cu: M, N >= 256, >= 2048 rows), norms
    un-folded, adapters on the per-block path (2128 is not a multiple of 64). Same oracle, same gates."""
    import fer_vit_b200 as fv
    from oracle import baseline_models as BM
    from oracle import reference_math as R
    fv.set_default_precision("bf16")
    sd = BM.hybrid_state_dict(seed=4)
    for i in range(12):
        sd[f"adapters.{i}.alpha"] = torch.ones(1) * (0.1 + 0.02 * i)
    model = fv.create_hybrid_latent_vit(model_size="base", use_pretrained=False, freeze_transformer=False,
                                        use_adapter=True, adapter_dim=64)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(13)
    B = 112
    x = 0.5 * torch.randn(B, 18, 512, generator=g) + 0.3 * torch.randn(1, 18, 512, generator=g)
    y = torch.randint(0, 7, (B,), generator=g)
    BM.hybrid_trainable(sd, freeze_transformer=False)
    ref = _oracle_step(lambda s, xx, m: R.hybrid_forward(s, xx, 12, 12, True, m), sd, x, y)
    got = step(model, x.cuda(), y.cuda())
    assert any(k.startswith("transformer.") for k in got[2])
    _compare("hybrid_vitb_full_finetune_vs_oracle", "bf16", got, ref, {"B": B})


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_hybrid_head_dropout_train_mode(precision):
    """train(): the head's Dropout(0.1) is active; the oracle receives the very mask the kernel drew."""
    from fer_vit_b200 import _lib as L
    from oracle import reference_math as R
    g = load_golden("hybrid_adapter")
    model = build_model("hybrid_adapter", precision)
    model.load_state_dict(g["sd"], strict=True)
    model = model.cuda().train()
    runner = model.plan_runner()
    runner._next_seed = lambda training: 4242
    x, y = g["x"], g["y"]
    got = step(model, x.cuda(), y.cuda())
    B, E = x.shape[0], 64
    mask = torch.empty(B * E, device="cuda")
    L.check(L.lib().fervit_dropout_mask(mask.data_ptr(), B * E, 0.1, 4242, L.SITE_HEAD, torch.cuda.current_stream().cuda_stream))
    masks = {"head": mask.reshape(B, E).cpu().double()}
    sd = {k: (v.double().requires_grad_(k in g["grad"]) if v.is_floating_point() else v) for k, v in g["sd"].items()}
    ref = _oracle_step(lambda s, xx, m: R.hybrid_forward(s, xx, 2, 2, True, m), sd, x.double(), y, masks)
    assert (masks["head"] == 0).any()
    _compare("hybrid_head_dropout", precision, got, ref)


@pytest.mark.parametrize("precision,B", [("fp32", 32), ("bf16", 32), ("bf16", 112)])
def test_latent_vit_default_config_vs_oracle(precision, B):
    """BASELINE config 1: LatentViT 512/d6/h8/2048 on 18x512 tokens, batch 32. Batch 112 (2128 token rows) is past the
    2048 rows from which the bf16 plan takes weight AND bias gradients from the CTA-pair kernel (gemm_wgrad2.cu), the
    path configs 2 and 4 train on."""
    import fer_vit_b200 as fv
    from oracle import reference_math as R
    fv.set_default_precision(precision)
    torch.manual_seed(5)
    model = fv.LatentViT(dropout=0.0)
    with torch.no_grad():                         # de-correlate the deep-copied layers, shrink randn cls/pos
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
        for i, layer in enumerate(model.transformer.layers):
            for p in layer.parameters():
                if p.dim() > 1:
                    p.add_(0.02 * torch.randn_like(p))
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(42)
    x = torch.randn(B, 18, 512, generator=g); y = torch.randint(0, 7, (B,), generator=g)
    model = model.cuda().train()
    model.plan_runner().keep_workspace = True
    import fer_vit_b200 as fv2
    model.zero_grad(set_to_none=True)
    logits = model(x.cuda())
    loss = fv2.cross_entropy(logits, y.cuda(), None, 0.1)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    sel = _relu_selectors(model, 6, B, 19)
    ref = _relu_reference("latent_vit_cfg1", precision, lambda s, xx, m: R.latent_vit_forward(s, xx, 6, 8, m), sd, x, y,
                          sel, smoothing=0.1)
    base = _torch_autocast_step(model, x.cuda(), y.cuda(), 0.1) if precision == "bf16" else None
    _compare("latent_vit_cfg1_vs_oracle", precision, (logits.detach(), loss.detach(), grads), ref, {"B": B},
             baseline=base)


@pytest.mark.parametrize("dims", [(64, 2, 128, 2, 5), (256, 4, 768, 2, 37)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_latent_vit_dropout_masks_vs_oracle(precision, dims):
    """dropout = 0.1 in train(): all four per-layer sites (attention weights, after out-proj, after ReLU, after
    linear2) use counter-based masks that the test materialises and hands to the oracle. The second shape (E = 256,
    F = 768, 703 token rows) spans several 256 x 256 tiles with ragged M and N tails, so the dropout epilogues of the
    CTA-pair GEMM index the mask across tile boundaries."""
    import fer_vit_b200 as fv
    from fer_vit_b200 import _lib as L
    from oracle import reference_math as R
    fv.set_default_precision(precision)
    torch.manual_seed(9)
    E, H, F, depth, B = dims
    S = 19
    model = fv.LatentViT(latent_dim=64, embed_dim=E, depth=depth, heads=H, mlp_dim=F, dropout=0.1)
    sd = {k: v.detach().clone().double().requires_grad_(True) for k, v in model.state_dict().items()}
    model = model.cuda().train()
    runner = model.plan_runner()
    runner._next_seed = lambda training: 99
    runner.keep_workspace = True
    x = torch.randn(B, 18, 64); y = torch.randint(0, 7, (B,))
    got = step(model, x.cuda(), y.cuda())
    sel = _relu_selectors(model, depth, B, S)

    def mask(n, site, shape):
        m = torch.empty(n, device="cuda")
        L.check(L.lib().fervit_dropout_mask(m.data_ptr(), n, 0.1, 99, site, torch.cuda.current_stream().cuda_stream))
        return m.reshape(shape).cpu().double()
    masks = {}
    for i in range(depth):
        masks[("attn", i)] = mask(B * H * S * S, 8 * i + 0, (B, H, S, S))
        masks[("drop1", i)] = mask(B * S * E, 8 * i + 1, (B, S, E))
        masks[("ffn", i)] = mask(B * S * F, 8 * i + 2, (B, S, F))
        masks[("drop2", i)] = mask(B * S * E, 8 * i + 3, (B, S, E))
    ref = _relu_reference("latent_vit_dropout", precision, lambda s, xx, m: R.latent_vit_forward(s, xx, depth, H, m),
                          sd, x.double(), y, sel, masks)
    _compare("latent_vit_dropout", precision, got, ref)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_latent_vit_v2_config4_vs_oracle(precision):
    """BASELINE config 4 shape: LatentViTv2 with LEAM + SemanticPE + LayerWiseNorm(residual) (reduced batch)."""
    import fer_vit_b200 as fv
    from oracle import reference_math as R
    fv.set_default_precision(precision)
    torch.manual_seed(6)
    model = fv.LatentViTv2(dropout=0.0, use_lwn=True, use_lwn_residual=True, use_spe=True, use_leam=True)
    with torch.no_grad():
        model.lwn.gate.add_(4.0 + torch.randn(18))
        for p in model.parameters():
            if p.dim() == 1 and p.numel() > 18:
                p.add_(0.05 * torch.randn_like(p))
    sd = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v.clone())
          for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(43)
    B = 16
    x = 0.6 * torch.randn(B, 18, 512, generator=g) + 0.2; y = torch.randint(0, 7, (B,), generator=g)
    model = model.cuda().train()
    model.plan_runner().keep_workspace = True
    got = step(model, x.cuda(), y.cuda())
    sel = _relu_selectors(model, 6, B, 19)
    ref = _relu_reference("latent_vit_v2_cfg4", precision,
                          lambda s, xx, m: R.latent_vit_v2_forward(s, xx, 6, 8, True, True, True, True, m), sd, x, y, sel)
    base = None
    if precision == "bf16":
        base = _torch_autocast_step(model, x.cuda(), y.cuda(), 0.0,
                                    pre=lambda t: model.leam(model.lwn(model.spe(t))))
    _compare("latent_vit_v2_cfg4_vs_oracle", precision, got, ref, baseline=base)


@pytest.mark.parametrize("precision,B", [("fp32", 3), ("bf16", 3), ("bf16", 11)])
def test_image_vit_config2_vs_oracle(precision, B):
    """BASELINE config 2 shape: ImageViT 512/d6/h8/2048 on 224x224 (S = 197), reduced batch. 11 images are 2167 token
    rows: the CTA-pair weight + bias gradient kernel's path (gemm_wgrad2.cu) under the same oracle and gates."""
    import fer_vit_b200 as fv
    from oracle import reference_math as R
    fv.set_default_precision(precision)
    torch.manual_seed(7)
    model = fv.ImageViT(embed_dim=512, depth=6, heads=8, mlp_dim=2048, dropout=0.0)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(44)
    x = torch.randn(B, 3, 224, 224, generator=g); y = torch.randint(0, 7, (B,), generator=g)
    ref = _oracle_step(lambda s, xx, m: R.image_vit_forward(s, xx, 6, 8, 16, m), sd, x, y)
    got = step(model.cuda().train(), x.cuda(), y.cuda())
    _compare("image_vit_cfg2_vs_oracle", precision, got, ref, {"B": B})


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_batch_properties_config3(precision):
    """Size-independent properties at BASELINE config 3's full size (ViT-B + adapters, batch 256):
    samples are independent, so the logits of a batch equal bit-for-bit the logits of its halves, and the
    mean-loss gradient of the batch is the average of the halves' gradients."""
    import fer_vit_b200 as fv
    fv.set_default_precision(precision)
    torch.manual_seed(8)
    model = fv.create_hybrid_latent_vit(model_size="base", use_pretrained=False, freeze_transformer=True,
                                        use_adapter=True, adapter_dim=64).cuda().eval()
    B = 256
    x = torch.randn(B, 18, 512, device="cuda"); y = torch.randint(0, 7, (B,), device="cuda")
    full = step(model, x, y)
    a = step(model, x[:128].contiguous(), y[:128].contiguous())
    b = step(model, x[128:].contiguous(), y[128:].contiguous())
    assert torch.equal(full[0], torch.cat([a[0], b[0]])), "logits must not depend on batch composition"
    worst = 0.0
    for k in full[2]:
        worst = max(worst, relerr(full[2][k], 0.5 * (a[2][k] + b[2][k])))
    record("full_batch_properties_cfg3", precision=precision, worst_grad_additivity=worst)
    assert worst < (1e-4 if precision == "fp32" else 5e-3)
    assert torch.isfinite(full[1]) and all(torch.isfinite(v).all() for v in full[2].values())


def test_no_cpu_fallback():
    import fer_vit_b200 as fv
    model = build_model("latent_vit", "fp32")
    with pytest.raises(RuntimeError, match="no CPU"):
        model(torch.randn(2, 18, 64))


def test_inference_and_checkpoint_roundtrip(tmp_path):
    """evaluate_model.py-style use: save a checkpoint, rebuild from config, strict load, no_grad eval forward."""
    import fer_vit_b200 as fv
    model = build_model("hybrid_adapter", "fp32").cuda().eval()
    x = torch.randn(9, 18, 64, device="cuda")
    with torch.no_grad():
        a = model(x)
    path = tmp_path / "ckpt.pt"
    torch.save({"model_state_dict": model.state_dict()}, path)
    again = build_model("hybrid_adapter", "fp32")
    again.load_state_dict(torch.load(path)["model_state_dict"], strict=True)
    again = again.cuda().eval()
    with torch.no_grad():
        b = again(x)
    assert torch.equal(a, b)
    assert not any(k.startswith("_") for k in model.state_dict())
    # in-place optimizer updates invalidate the bf16 weight cache
    again.precision = "bf16"
    with torch.no_grad():
        c = again(x)
        again.input_proj.weight.mul_(2.0)
        d = again(x)
    assert relerr(c, a) < 2e-2 and relerr(d, c) > 1e-3


@pytest.mark.gpu
def test_graphed_train_step_matches_eager():
    """fer_vit_b200.GraphedTrainStep replays exactly the kernels of the eager step: after the same three AdamW steps on
    the same batches, two models that started identical hold bit-identical parameters and losses. Dropout off (eval-mode
    head) so both draw no masks; the optimizer updates make the trainable bf16 weight re-casts part of the graph."""
    import copy
    import fer_vit_b200 as fv
    g = load_golden("hybrid_adapter")
    torch.manual_seed(0)
    m1 = build_model("hybrid_adapter", "bf16")
    m1.load_state_dict(g["sd"], strict=True)
    m1 = m1.cuda().eval()
    m2 = copy.deepcopy(m1)
    xs = [torch.randn_like(g["x"]).cuda() for _ in range(3)]
    ys = [torch.randint(0, 7, g["y"].shape).cuda() for _ in range(3)]

    def make_opt(m):
        return torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-2, fused=True, capturable=True)
    o1, o2 = make_opt(m1), make_opt(m2)
    # the graphed step warms up on its example batch (3 eager steps): give the eager twin the same history
    stepper = fv.GraphedTrainStep(m2, o2, xs[0], ys[0], warmup=3)
    losses1, losses2 = [], []
    for _ in range(3):  # the 3 warm-up steps on (xs[0], ys[0]); the capture pass records kernels without running them
        o1.zero_grad(set_to_none=True)
        fv.cross_entropy(m1(xs[0]), ys[0]).backward()
        o1.step()
    for x, y in zip(xs, ys):
        o1.zero_grad(set_to_none=True)
        l1 = fv.cross_entropy(m1(x), y)
        l1.backward()
        o1.step()
        losses1.append(float(l1.detach()))
        losses2.append(float(stepper(x, y)))
    torch.cuda.synchronize()
    assert stepper.launches_per_replay > 50
    assert losses1 == losses2, (losses1, losses2)
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), k


@pytest.mark.gpu
def test_fused_optimizer_updates_reach_the_bf16_weight_cache():
    """torch.optim.AdamW(fused=True) mutates parameters without bumping Tensor._version, so the derived bf16 weight
    cache cannot rely on version counters for trainable weights: after a fused step the model must compute with the
    NEW weights — its logits equal those of a fresh copy (fresh cache) built from its state_dict."""
    import copy
    g = load_golden("hybrid_adapter")
    m = build_model("hybrid_adapter", "bf16")
    m.load_state_dict(g["sd"], strict=True)
    m = m.cuda().eval()
    x, y = g["x"].cuda(), g["y"].cuda()
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=5e-2, fused=True)
    before = m(x).detach().clone()
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        fv_loss = __import__("fer_vit_b200").cross_entropy(m(x), y)
        fv_loss.backward()
        opt.step()
    after = m(x).detach()
    fresh = copy.deepcopy(m)
    assert torch.equal(after, fresh(x).detach()), "stale bf16 weight cache after a fused optimizer step"
    assert relerr(after, before) > 1e-3, "the optimizer steps should have moved the logits"


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("size,E,heads,B", [("small", 384, 6, 5), ("tiny", 192, 3, 1)])
def test_hybrid_small_widths_and_odd_batches_vs_oracle(precision, size, E, heads, B):
    """The reference's default `--model_size small` (train_hybrid_latent_vit.py:393-394) and `tiny`: widths that are
    not multiples of 256 (the adapters take the unfused two-GEMM path, N % BN != 0 tiles in every GEMM), with odd
    batch sizes down to a single sample (T = 19 rows: every tile is mostly padding)."""
    import fer_vit_b200 as fv
    from oracle import baseline_models as BM
    from oracle import reference_math as R
    fv.set_default_precision(precision)
    sd = BM.hybrid_state_dict(E=E, depth=12, heads=heads, seed=5)
    for i in range(12):
        sd[f"adapters.{i}.alpha"] = torch.ones(1) * (0.1 + 0.03 * i)
    model = fv.create_hybrid_latent_vit(model_size=size, use_pretrained=False, freeze_transformer=True,
                                        use_adapter=True, adapter_dim=64)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(17 + B)
    x = torch.randn(B, 18, 512, generator=g)
    y = torch.randint(0, 7, (B,), generator=g)
    BM.hybrid_trainable(sd)
    ref = _oracle_step(lambda s, xx, m: R.hybrid_forward(s, xx, 12, heads, True, m), sd, x, y)
    got = step(model, x.cuda(), y.cuda())
    _compare(f"hybrid_{size}_adapter_vs_oracle", precision, got, ref, {"B": B})


@pytest.mark.gpu
def test_graphed_step_with_dropout_bf16():
    """LatentViT in train mode (dropout 0.1 at every site) under GraphedTrainStep: the dropout masks come from the
    counter-based hash with the device-side seed counter incremented INSIDE the graph, so every replay draws fresh
    masks (two replays on the same batch with a zero learning rate give different losses) and training on a fixed
    batch still converges."""
    import fer_vit_b200 as fv
    fv.set_default_precision("bf16")
    torch.manual_seed(3)
    model = fv.LatentViT(latent_dim=512, seq_len=18, embed_dim=256, depth=2, heads=4, mlp_dim=512, num_classes=7,
                         dropout=0.1).cuda().train()
    x = torch.randn(64, 18, 512, device="cuda")
    y = torch.randint(0, 7, (64,), device="cuda")
    opt = fv.FusedAdamW(model.parameters(), lr=0.0, weight_decay=0.0)
    stepper = fv.GraphedTrainStep(model, opt, x, y)
    l1 = float(stepper(x, y))
    l2 = float(stepper(x, y))
    assert l1 != l2 and abs(l1 - l2) < 0.5, (l1, l2)     # same weights (lr = 0), fresh masks
    for g in opt.param_groups:
        g["lr"] = 2e-3
    opt.refresh_hyper()
    losses = [float(stepper(x, y)) for _ in range(30)]
    assert all(map(lambda v: v == v and v < 1e3, losses))
    assert sum(losses[-5:]) / 5 < sum(losses[:5]) / 5 - 0.05, (losses[:5], losses[-5:])


@pytest.mark.parametrize("size,E,heads", [("base", 768, 12), ("small", 384, 6)])
def test_layernorm_fold_vs_unfolded_and_oracle(size, E, heads):
    """Frozen norm1 / norm2 folded into the qkv / fc1 GEMMs (fervit_plan_set_ln_fold; bf16 mode): same logits and
    gradients as the un-folded path within bf16 noise, both inside the 2e-2 gate against the fp64 oracle on the
    non-zero-mean input distribution (the case that stresses the mean-shift term), and 23 kernel launches fewer per
    forward (every norm except block 0's norm1). One block is then given a trainable norm1 and a trainable fc1 bias:
    those two norms fall back to the kernel, their gradients appear and still match the oracle."""
    import os
    import fer_vit_b200 as fv
    from fer_vit_b200 import _lib as L
    from oracle import baseline_models as BM
    from oracle import reference_math as R
    fv.set_default_precision("bf16")
    sd = BM.hybrid_state_dict(E=E, depth=12, heads=heads, seed=7)
    model = fv.create_hybrid_latent_vit(model_size=size, use_pretrained=False, freeze_transformer=True, use_adapter=True,
                                        adapter_dim=64)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(21)
    B = 8
    x = 0.5 * torch.randn(B, 18, 512, generator=g) + 0.3 * torch.randn(1, 18, 512, generator=g)
    y = torch.randint(0, 7, (B,), generator=g)
    BM.hybrid_trainable(sd)
    ref = _oracle_step(lambda s, xx, m: R.hybrid_forward(s, xx, 12, heads, True, m), sd, x, y)
    old = os.environ.get("FERVIT_LN_FOLD")
    try:
        out = {}
        for flag in ("0", "3"):
            os.environ["FERVIT_LN_FOLD"] = flag
            step(model, x.cuda(), y.cuda())                      # cache refresh for this setting
            n0 = L.launch_count()
            with torch.no_grad():
                inferred = model(x.cuda())
            out[flag] = (step(model, x.cuda(), y.cuda()), L.launch_count() - n0)
            # the no-grad forward (shared activation buffers, nothing saved) runs the same kernels: identical logits
            assert torch.equal(inferred, out[flag][0][0]), flag
        (plain, n_plain), (folded, n_folded) = out["0"], out["3"]
        runner = model.plan_runner()
        assert sum(runner.fold1) == 11 and sum(runner.fold2) == 12
        e_l = relerr(folded[0], plain[0])
        e_g = max(relerr(folded[2][k], plain[2][k]) for k in plain[2] if not k.endswith(".alpha"))
        record("layernorm_fold_vs_unfolded", size=size, err_logits=e_l, worst_grad=e_g)
        assert e_l < 1e-2 and e_g < 2e-2, (e_l, e_g)
        _compare(f"hybrid_{size}_ln_unfolded_vs_oracle", "bf16", plain, ref)
        _compare(f"hybrid_{size}_ln_folded_vs_oracle", "bf16", folded, ref)
        # forward + (forward + backward): 2 x 23 norm kernels fewer
        assert n_plain - n_folded == 46, (n_plain, n_folded)
        # per-parameter: a trainable norm1 weight in block 3 and a trainable fc1 bias in block 5 un-fold those two norms
        named = dict(model.named_parameters())
        named["transformer.3.norm1.weight"].requires_grad_(True)
        named["transformer.5.mlp.fc1.bias"].requires_grad_(True)
        sd["transformer.3.norm1.weight"].requires_grad_(True)
        sd["transformer.5.mlp.fc1.bias"].requires_grad_(True)
        ref2 = _oracle_step(lambda s, xx, m: R.hybrid_forward(s, xx, 12, heads, True, m), sd, x, y)
        got2 = step(model, x.cuda(), y.cuda())
        assert runner.fold1[3] == 0 and runner.fold2[5] == 0 and sum(runner.fold1) == 10 and sum(runner.fold2) == 11
        assert "transformer.3.norm1.weight" in got2[2] and "transformer.5.mlp.fc1.bias" in got2[2]
        _compare(f"hybrid_{size}_ln_partly_folded_vs_oracle", "bf16", got2, ref2)
    finally:
        if old is None:
            os.environ.pop("FERVIT_LN_FOLD", None)
        else:
            os.environ["FERVIT_LN_FOLD"] = old
