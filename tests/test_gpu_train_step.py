"""GPU: the COMPOSED train step (GraphedTrainStep: zero_grad + plan forward + loss + plan backward + FusedAdamW, one
CUDA-graph replay) against the reference TRAINERS' own goldens:

  tests/golden/v2_train_epoch.npz      one batch of train/train_latent_vit_v2.py:107-143 (mixup, class-weighted smoothed
                                       CE, clip_grad_norm_ active, AdamW) — no dropout in that fixture, so every
                                       parameter after the step is compared DIRECTLY with what the unmodified reference
                                       trainer produced;
  tests/golden/hybrid_train_epoch.npz  one batch of train/train_hybrid_latent_vit.py:120-147 in train mode (five
                                       layer-wise LR groups of :63-117). Its head Dropout(0.1) mask came from torch's
                                       CPU generator, which a counter-based kernel cannot reproduce, so the GPU step is
                                       compared with the oracle fed the mask the KERNEL drew; the oracle itself is
                                       pinned on that golden (tests/test_data_path_cpu.py).

What is compared is the UPDATE p_after - p_before of every trainable tensor. The first AdamW step is
lr * g / (|g| + eps): a sign function of the gradient, so an element whose gradient is ~0 can flip; with fp32-mode
gradients (relative error ~1e-6) that moves a tensor's update by < 5e-3, the gate used here; bf16-mode gradients
(~1e-2) flip a few per cent of the signs by construction, so bf16 is gated on the loss, the clipped gradient norm and
the agreement of the update's sign (> 95 % of all elements).

Also here: per-parameter requires_grad inside transformer blocks (weight/bias pairs are produced together
by the native backward; the host layer must still honour each flag).
"""
import numpy as np
import os
import pytest
import torch

from tests.test_gpu_ops import record
from tests.util import GOLDEN, build_model, load_golden, relerr

pytestmark = pytest.mark.gpu


def _reset_optimizer_state(opt):
    """GraphedTrainStep's warm-up takes real optimizer steps; put moments and the step counter back to zero IN PLACE
    (the captured graph holds their addresses) so the next replay is AdamW step 1."""
    for st in opt.state.values():
        st["exp_avg"].zero_()
        st["exp_avg_sq"].zero_()
    opt._step_dev.zero_()


def _sign_agreement(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((torch.sign(a) == torch.sign(b)).double().mean())


def _bf16_update_gate(upds, refs):
    """bf16 mode: the sign of the first AdamW update agrees on > 95 % of ALL elements and on > 75 % of every tensor
    (tiny tensors of these test-sized models hold a few dozen elements, a handful of them with gradients ~0)."""
    a = torch.cat([u.reshape(-1) for u in upds.values()])
    b = torch.cat([refs[k].reshape(-1) for k in upds])
    per = {k: _sign_agreement(upds[k], refs[k]) for k in upds}
    return _sign_agreement(a, b), min(per.values())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_v2_trainer_step_matches_reference_golden(precision):
    import fer_vit_b200 as fv
    z = np.load(os.path.join(GOLDEN, "v2_train_epoch.npz"))
    g = load_golden("latent_vit_v2")
    after = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("after/")}
    model = build_model("latent_vit_v2", precision)
    model.load_state_dict(g["sd"], strict=True)
    model = model.cuda().train()
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    x = torch.from_numpy(z["x"]).float().cuda()
    y = torch.from_numpy(z["y"]).cuda()
    index = torch.from_numpy(z["index"]).cuda()
    lam, w = float(z["lam"]), torch.from_numpy(z["class_weight"]).float().cuda()
    smoothing, clip = float(z["label_smoothing"]), float(z["grad_clip"])
    opt = fv.FusedAdamW(model.parameters(), lr=float(z["lr"]), weight_decay=float(z["weight_decay"]), max_grad_norm=clip)
    mixed = fv.mixup(x, index, lam)                                        # train_latent_vit_v2.py:125-126
    loss_fn = lambda lg, yy: fv.mixup_cross_entropy(lg, yy, index, lam, w, smoothing)   # :130
    stepper = fv.GraphedTrainStep(model, opt, mixed, y, loss_fn=loss_fn, warmup=1)
    with torch.no_grad():
        for k, p in model.named_parameters():
            p.copy_(before[k])
    _reset_optimizer_state(opt)
    loss = float(stepper(mixed, y))
    torch.cuda.synchronize()
    total_norm = float(opt.last_total_norm)
    errs, upds, refs = {}, {}, {}
    for k, p in model.named_parameters():
        refs[k] = after[k] - before[k].double().cpu()
        upds[k] = p.detach().double().cpu() - before[k].double().cpu()
        errs[k] = relerr(upds[k], refs[k])
    worst = max(errs, key=errs.get)
    sign_all, sign_min = _bf16_update_gate(upds, refs)
    record("graphed_v2_trainer_step_vs_golden", precision=precision, err_loss=abs(loss - float(z["loss"])),
           err_total_norm=abs(total_norm - float(z["total_norm"])) / float(z["total_norm"]),
           worst_update=errs[worst], worst_key=worst, sign_agreement_all=sign_all, sign_agreement_worst_tensor=sign_min)
    assert set(errs) == set(after) - {"spe.groups"}
    tol = 1e-4 if precision == "fp32" else 2e-2
    assert abs(loss - float(z["loss"])) < tol * max(1.0, abs(float(z["loss"])))
    assert abs(total_norm - float(z["total_norm"])) < tol * float(z["total_norm"])
    assert total_norm > clip                     # clipping is active in this fixture
    if precision == "fp32":
        assert errs[worst] < 5e-3, (worst, errs[worst])
    else:
        assert sign_all > 0.95 and sign_min > 0.75, (sign_all, sign_min)
    stepper.close()
    with pytest.raises(RuntimeError, match="closed"):
        stepper(mixed, y)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_hybrid_trainer_step_matches_pinned_oracle(precision):
    import fer_vit_b200 as fv
    from fer_vit_b200 import _lib as L
    from oracle import reference_math as R
    z = np.load(os.path.join(GOLDEN, "hybrid_train_epoch.npz"))
    g = load_golden("hybrid_adapter")
    hyper = {k[6:]: tuple(float(v) for v in z[k]) for k in z.files if k.startswith("hyper/")}
    model = build_model("hybrid_adapter", precision)
    model.load_state_dict(g["sd"], strict=True)
    model = model.cuda().train()                                 # head Dropout(0.1) active, as in the trainer
    named = dict(model.named_parameters())
    assert {k for k, p in named.items() if p.requires_grad} == set(hyper)
    # the trainer's five layer-wise groups (train_hybrid_latent_vit.py:63-117) as (lr, weight_decay) classes
    groups = {}
    for k, hw in hyper.items():
        groups.setdefault(hw, []).append(named[k])
    opt = fv.FusedAdamW([{"params": ps, "lr": lr, "weight_decay": wd} for (lr, wd), ps in groups.items()])
    before = {k: named[k].detach().clone() for k in hyper}
    x = torch.from_numpy(z["x"]).float().cuda()
    y = torch.from_numpy(z["y"]).cuda()
    runner = model.plan_runner()
    runner._next_seed = lambda training: 4242
    stepper = fv.GraphedTrainStep(model, opt, x, y, warmup=1)
    with torch.no_grad():
        for k in hyper:
            named[k].copy_(before[k])
    _reset_optimizer_state(opt)
    loss = float(stepper(x, y))
    torch.cuda.synchronize()
    # the mask of THIS replay: host seed baked into the graph + the device counter incremented inside it
    B, E = x.shape[0], 64
    mask = torch.empty(B * E, device="cuda")
    L.check(L.lib().fervit_dropout_mask(mask.data_ptr(), B * E, 0.1, 4242 + int(stepper._seed_dev.item()), L.SITE_HEAD,
                                        torch.cuda.current_stream().cuda_stream))
    mask = mask.reshape(B, E).cpu().double()
    assert (mask == 0).any() and (mask > 1).any()
    sd = {k: (v.double().requires_grad_(k in hyper) if v.is_floating_point() else v) for k, v in g["sd"].items()}
    logits = R.hybrid_forward(sd, x.double().cpu(), 2, 2, True, {"head": mask})
    oloss = R.cross_entropy(logits, y.cpu())
    grads = R.grads_of(oloss, sd)
    errs, upds, refs = {}, {}, {}
    for k, (lr, wd) in hyper.items():
        p0 = sd[k].detach()
        p1, _, _ = R.adamw_step(p0, grads[k], torch.zeros_like(p0), torch.zeros_like(p0), 1, lr, (0.9, 0.999), 1e-8, wd)
        upds[k] = named[k].detach().double().cpu() - before[k].double().cpu()
        refs[k] = p1 - p0
        errs[k] = relerr(upds[k], refs[k])
    worst = max(errs, key=errs.get)
    sign_all, sign_min = _bf16_update_gate(upds, refs)
    record("graphed_hybrid_trainer_step_vs_pinned_oracle", precision=precision, err_loss=abs(loss - oloss.item()),
           worst_update=errs[worst], worst_key=worst, sign_agreement_all=sign_all, sign_agreement_worst_tensor=sign_min)
    tol = 1e-4 if precision == "fp32" else 2e-2
    assert abs(loss - oloss.item()) < tol * max(1.0, abs(oloss.item()))
    # frozen blocks untouched
    for k, p in named.items():
        if k.startswith("transformer."):
            assert torch.equal(p.detach().cpu(), g["sd"][k]), k
    if precision == "fp32":
        assert errs[worst] < 5e-3, (worst, errs[worst])
    else:
        assert sign_all > 0.95 and sign_min > 0.75, (sign_all, sign_min)
    stepper.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_per_parameter_requires_grad_inside_blocks(precision):
    """A weight frozen beside a trainable bias (and the reverse) inside transformer blocks: the trainable member gets
    the reference gradient, the frozen one gets none, nothing faults (plan.cu block backward)."""
    import fer_vit_b200 as fv
    g = load_golden("hybrid_full")
    model = build_model("hybrid_full", precision)
    model.load_state_dict(g["sd"], strict=True)
    model = model.cuda().eval()
    frozen = ["transformer.0.attn.qkv.weight", "transformer.0.mlp.fc1.bias", "transformer.1.norm1.weight",
              "transformer.1.attn.proj.bias", "transformer.1.mlp.fc2.weight", "transformer.0.norm2.bias",
              "input_proj.bias", "head.0.weight"]
    named = dict(model.named_parameters())
    for k in frozen:
        named[k].requires_grad_(False)
    model.zero_grad(set_to_none=True)
    w = g["class_weight"].cuda() if g["class_weight"] is not None else None
    loss = fv.cross_entropy(model(g["x"].cuda()), g["y"].cuda(), w, g["label_smoothing"])
    loss.backward()
    torch.cuda.synchronize()
    tol = 1e-4 if precision == "fp32" else 2e-2
    errs = {}
    for k, p in named.items():
        if k in frozen:
            assert p.grad is None, k
        else:
            assert p.grad is not None, k
            errs[k] = relerr(p.grad, g["grad"][k])
    worst = max(errs, key=errs.get)
    record("per_parameter_requires_grad", precision=precision, worst_grad=errs[worst], worst_key=worst)
    assert errs[worst] < tol, (worst, errs[worst])


def test_graphed_step_takes_a_partial_last_batch():
    """A batch of another size than the captured one (the last batch of an epoch with drop_last=False) runs the same
    step host-launched instead of being refused; graph replays before and after it still work, and the whole sequence
    equals the eager loop bit for bit."""
    import copy
    import fer_vit_b200 as fv
    g = load_golden("hybrid_adapter")
    gen = torch.Generator().manual_seed(3)
    xs = [torch.randn(b, 18, 64, generator=gen).cuda() for b in (8, 8, 5, 8)]
    ys = [torch.randint(0, 7, (x.shape[0],), generator=gen).cuda() for x in xs]

    def make():
        m = build_model("hybrid_adapter", "bf16")
        m.load_state_dict(g["sd"], strict=True)
        m = m.cuda().eval()                      # eval: no head-dropout draws, so both loops are deterministic
        o = fv.FusedAdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.01)
        return m, o
    m1, o1 = make()
    stepper = fv.GraphedTrainStep(m1, o1, xs[0], ys[0], warmup=1)
    sd0 = copy.deepcopy(m1.state_dict())
    m2, o2 = make()
    m2.load_state_dict(sd0)
    o2.load_state_dict(copy.deepcopy(o1.state_dict()))
    l1 = [float(stepper(x, y)) for x, y in zip(xs, ys)]
    l2 = []
    for x, y in zip(xs, ys):
        o2.zero_grad(set_to_none=True)
        loss = fv.cross_entropy(m2(x), y)
        loss.backward()
        o2.step()
        l2.append(float(loss))
    assert l1 == l2, (l1, l2)
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    with pytest.raises(RuntimeError, match="captured for inputs"):
        stepper(torch.randn(8, 17, 64, device="cuda"), ys[0])
    stepper.close()
