"""GPU: the data-side prologue of the LatentViT train step through the C ABI -
`fervit_latent_batch` (gather + LatentAugment + mixup) and `fervit_cross_entropy_mixup` against the oracle, and the
reference's own mixup train step (golden fixture of train/train_latent_vit.py:108-142) through the drop-in classes.

Tolerances: the kernels compute in fp32, the oracle in fp64 from the same draws: 2e-6 norm-wise for the batch (the
Gaussian uses device logf/sincospif), masks exact, gather/mixup-only paths bit-exact against the fp32 expression."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import reference_math as R
from tests.util import GOLDEN, build_model, load_golden, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, **vals):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_metrics.jsonl"), "a") as fh:
        fh.write(json.dumps({"test": name, **vals}) + "\n")


def oracle_batch(table, idx, B, std, rng, p, seed, mix, lam):
    row = table[0].numel()
    src = table[idx] if idx is not None else table[:B]
    normal, scale, keep = R.latent_augment_draws(seed, B, row, rng, p)
    shp = src.shape
    a = R.latent_augment(src.double(), std, rng, p, normal.reshape(shp), scale, keep.reshape(shp))
    return (R.mixup(a, mix, lam) if mix is not None else a), keep.reshape(shp)


@pytest.mark.parametrize("B", [1, 7, 64])
@pytest.mark.parametrize("aug", ["off", "all", "noise", "scale_mask"])
@pytest.mark.parametrize("mix", [False, True])
def test_latent_batch_matches_oracle(B, aug, mix):
    import fer_vit_b200 as fv
    std, rng, p = {"off": (0.0, None, 0.0), "all": (0.1, (0.9, 1.1), 0.1), "noise": (0.05, None, 0.0),
                   "scale_mask": (0.0, (0.5, 1.5), 0.3)}[aug]
    g = torch.Generator().manual_seed(100 + B)
    N = 3 * B + 2
    table = torch.randn(N, 18, 512, generator=g) + 0.3
    labels = torch.randint(0, 7, (N,), generator=g)
    idx = torch.randint(0, N, (B,), generator=g)
    perm = torch.randperm(B, generator=g) if mix else None
    lam = 0.348 if mix else 1.0
    seed = 0xC0FFEE + B
    t = fv.LatentAugment(std, rng, p)
    out, lab = fv.latent_batch(table.cuda(), labels.cuda(), idx.cuda(), t, perm.cuda() if mix else None, lam, seed)
    ref, keep = oracle_batch(table, idx, B, std, rng, p, seed, perm, lam)
    assert torch.equal(lab.cpu(), labels[idx])
    e = relerr(out, ref)
    record("latent_batch", B=B, aug=aug, mixup=mix, err=e)
    assert e < 2e-6
    if aug == "off":                                              # pure gather / blend: the fp32 expression, bit-exact
        want = table[idx]
        want = lam * want + (1 - lam) * want[perm] if mix else want
        assert torch.equal(out.cpu(), want)
    if p > 0 and not mix:                                         # masked elements are exactly zero, and only those
        assert torch.equal(out.cpu() == 0, ~keep)


def test_latent_batch_device_scalars_and_graph_replay():
    """lam and the seed counter are read on the device, so one captured launch yields a new batch per replay."""
    import fer_vit_b200 as fv
    g = torch.Generator().manual_seed(3)
    table = torch.randn(16, 18, 512, generator=g)
    dev_table = table.cuda()
    idx = torch.arange(16).cuda()
    perm = torch.randperm(16, generator=g)
    perm_dev = perm.cuda()
    lam_dev = torch.tensor([0.25], device="cuda")
    seed_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
    t = fv.LatentAugment(0.1, (0.9, 1.1), 0.1)
    out = torch.empty(16, 18, 512, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fv.latent_batch(dev_table, None, idx, t, perm_dev, lam_dev, 77, seed_dev, out)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fv.latent_batch(dev_table, None, idx, t, perm_dev, lam_dev, 77, seed_dev, out)
    for step, lam in ((0, 0.25), (5, 0.8)):
        lam_dev.fill_(lam)
        seed_dev.fill_(step)
        graph.replay()
        ref, _ = oracle_batch(table, None, 16, 0.1, (0.9, 1.1), 0.1, 77 + step, perm, float(np.float32(lam)))
        assert relerr(out, ref) < 2e-6


def test_latent_augment_golden_inputs():
    """On the reference fixture's latent: identity when every stage is off (bit-exact with the reference output), and
    with the stages on, the reference's arithmetic on the kernel's own draws."""
    import fer_vit_b200 as fv
    z = np.load(os.path.join(GOLDEN, "latent_augment.npz"))
    x = torch.from_numpy(z["off/x"])
    assert torch.equal(fv.LatentAugment()(x.cuda()).cpu(), torch.from_numpy(z["off/out"]))
    t = fv.LatentAugment(float(z["all/noise_std"]), tuple(float(v) for v in z["all/scale_range"]),
                         float(z["all/mask_prob"]))
    out = t(x.cuda(), seed=9)
    normal, scale, keep = R.latent_augment_draws(9, 1, x.numel(), t.scale_range, t.mask_prob)
    ref = R.latent_augment(x.double(), t.noise_std, t.scale_range, t.mask_prob, normal.reshape(x.shape), scale[0],
                           keep.reshape(x.shape))
    assert relerr(out, ref) < 2e-6
    # successive calls draw fresh noise, as the reference transform does
    assert not torch.equal(t(x.cuda()), t(x.cuda()))


def test_packed_latent_cache(tmp_path):
    import fer_vit_b200 as fv
    g = torch.Generator().manual_seed(8)
    lat = torch.randn(12, 18, 512, generator=g)
    lab = torch.randint(0, 7, (12,), generator=g)
    for i in range(12):                                           # the reference's per-sample files
        torch.save({"latent": lat[i].clone(), "label": int(lab[i]), "img_path": f"{i}.png"}, tmp_path / f"s{i:03d}.pt")
    cache = fv.PackedLatentCache.from_dir(str(tmp_path))
    assert len(cache) == 12 and cache.latents.is_cuda
    assert cache.get_class_counts() == {int(c): int((lab == c).sum()) for c in lab.unique()}
    assert cache.get_class_names()[3] == "happy"
    idx = torch.tensor([5, 0, 11, 5], device="cuda")
    x, y = cache.batch(idx)
    assert torch.equal(x.cpu(), lat[idx.cpu()]) and torch.equal(y.cpu(), lab[idx.cpu()])
    cache.check()
    cache.batch(torch.tensor([3, 12], device="cuda"))            # 12 is out of range
    with pytest.raises(IndexError):
        cache.check()


@pytest.mark.parametrize("B", [1, 8, 300])
@pytest.mark.parametrize("weighted,eps", [(False, 0.0), (True, 0.1)])
@pytest.mark.parametrize("lam", [0.0, 0.348, 1.0])
def test_mixup_cross_entropy(B, weighted, eps, lam):
    import fer_vit_b200 as fv
    g = torch.Generator().manual_seed(B)
    z = (torch.randn(B, 7, generator=g) * 3).cuda().requires_grad_(True)
    y = torch.randint(0, 7, (B,), generator=g)
    idx = torch.randperm(B, generator=g)
    w = (torch.rand(7, generator=g) + 0.5) if weighted else None
    loss = fv.mixup_cross_entropy(z, y.cuda(), idx.cuda(), lam, w.cuda() if weighted else None, eps)
    loss.backward()
    zr = z.detach().double().cpu().requires_grad_(True)
    ref = R.mixup_loss(zr, y, idx, lam, w.double() if weighted else None, eps)
    ref.backward()
    e_l, e_g = abs(loss.item() - ref.item()) / abs(ref.item()), relerr(z.grad, zr.grad)
    record("mixup_cross_entropy", B=B, weighted=weighted, eps=eps, lam=lam, err_loss=e_l, err_grad=e_g)
    assert e_l < 2e-6 and e_g < 1e-5
    # lam on the device gives the same numbers
    z2 = z.detach().clone().requires_grad_(True)
    l2 = fv.mixup_cross_entropy(z2, y.cuda(), idx.cuda(), torch.tensor([lam], device="cuda"),
                                w.cuda() if weighted else None, eps)
    l2.backward()
    assert torch.equal(l2, loss) and torch.equal(z2.grad, z.grad)
    # and it is the composition the reference writes (train_latent_vit.py:131) on the single-target kernel
    comp = lam * fv.cross_entropy(z.detach(), y.cuda(), w.cuda() if weighted else None, eps) + \
        (1 - lam) * fv.cross_entropy(z.detach(), y[idx].cuda(), w.cuda() if weighted else None, eps)
    assert abs(comp.item() - loss.item()) < 2e-6 * max(1.0, abs(loss.item()))


@pytest.mark.parametrize("precision,tol_l,tol_g", [("fp32", 1e-4, 1e-4), ("bf16", 2e-2, 2e-2)])
def test_mixup_train_step_matches_reference_golden(precision, tol_l, tol_g):
    """One batch of the reference's train_epoch (mixup 0.4, weighted CE, smoothing 0.1): loss, every gradient, and the
    predictions of its extra no-grad forward on the un-mixed batch."""
    import fer_vit_b200 as fv
    z = np.load(os.path.join(GOLDEN, "mixup_step.npz"))
    g = load_golden("latent_vit")
    model = build_model("latent_vit", precision)
    model.load_state_dict(g["sd"], strict=True)
    model = model.cuda().train()
    x, y, index = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["y"]).cuda(), torch.from_numpy(z["index"]).cuda()
    lam = float(z["lam"])
    crit = fv.CrossEntropyLoss(torch.from_numpy(z["class_weight"]).cuda(), float(z["label_smoothing"]))
    logits = model(fv.mixup(x, index, lam))
    loss = crit.mixup(logits, y, index, lam)
    loss.backward()
    e_l = abs(loss.item() - float(z["loss"])) / abs(float(z["loss"]))
    errs = {k: relerr(p.grad, torch.from_numpy(z["grad/" + k])) for k, p in model.named_parameters()}
    worst = max(errs, key=errs.get)
    record("mixup_train_step", precision=precision, err_loss=e_l, err_grad=errs[worst], worst=worst)
    assert e_l < tol_l
    assert errs[worst] < tol_g, (worst, errs[worst])
    with torch.no_grad():                                         # train-mode forward without saved activations
        pred = model(x).argmax(-1)
    assert torch.equal(pred.cpu(), torch.from_numpy(z["pred"]))
    assert abs((pred == y).double().mean().item() - float(z["accuracy"])) < 1e-12


# ------------------------------------------------------------------------------------------------
# LatentDecomposer / ExpressionAwareViT front-end
# ------------------------------------------------------------------------------------------------
DEC_MODES = [(d, o) for d in ("all_classes", "max_class") for o in ("expr_only", "id_only", "enhanced", "concat")]


@pytest.mark.parametrize("dm,om", DEC_MODES)
def test_latent_decomposer_matches_reference_golden(dm, om):
    import fer_vit_b200 as fv
    z = np.load(os.path.join(GOLDEN, "expression_aware.npz"))
    d = fv.LatentDecomposer({i: torch.from_numpy(z[f"raw_direction/{i}"]) for i in range(7)}, 18, 64).cuda()
    x = torch.from_numpy(z["x"]).cuda()
    out = d(x, output_mode=om, enhance_alpha=float(z["alpha"]), decompose_mode=dm)
    e = relerr(out, torch.from_numpy(z[f"out/{dm}/{om}"]))
    record("latent_decomposer_golden", decompose_mode=dm, output_mode=om, err=e)
    assert e < 1e-5
    assert relerr(d.get_expression_scores(x), torch.from_numpy(z["scores"])) < 1e-5
    we, wi = d.decompose(x, mode=dm)
    assert relerr(we + wi, x) < 1e-6
    assert relerr(d.enhance_expression(x, float(z["alpha"]), dm), torch.from_numpy(z[f"out/{dm}/enhanced"])) < 1e-5


@pytest.mark.parametrize("B", [1, 3, 64, 597])
@pytest.mark.parametrize("C", [1, 7, 8])
def test_latent_decomposer_full_size_against_oracle(B, C):
    """w+ of the real size (18 x 512), ragged batch sizes (R = 4 latents per CTA pass), every direction count."""
    import fer_vit_b200 as fv
    g = torch.Generator().manual_seed(B * 10 + C)
    raw = {i: torch.randn(18, 512, generator=g) for i in range(C)}
    d = fv.LatentDecomposer(raw, 18, 512).cuda()
    x = torch.randn(B, 18, 512, generator=g) + 0.1
    dirs = R.normalize_directions(torch.stack([raw[i] for i in range(C)]).double())
    for dm, om in (("all_classes", "concat"), ("max_class", "enhanced"), ("all_classes", "expr_only")):
        out = d(x.cuda(), output_mode=om, enhance_alpha=2.0, decompose_mode=dm)
        ref = R.decomposer_forward(x.double(), dirs, om, 2.0, dm)
        e = relerr(out, ref)
        record("latent_decomposer", B=B, C=C, decompose_mode=dm, output_mode=om, err=e)
        assert e < 1e-5, (dm, om, e)
    assert relerr(d.get_expression_scores(x.cuda()), R.latent_decompose(x.double(), dirs)[2]) < 1e-5


@pytest.mark.parametrize("precision,tol_l,tol_g", [("fp32", 1e-4, 1e-4), ("bf16", 2e-2, 2e-2)])
def test_expression_aware_vit_step_matches_reference_golden(precision, tol_l, tol_g):
    """ExpressionAwareViT in 'concat' mode (36 + 1 tokens), frozen blocks + adapters: logits, loss and every trainable
    gradient of the reference step; identical top-1."""
    import fer_vit_b200 as fv
    from fer_vit_b200.models_fer_vit.vit_blocks import register_vit_config
    z = np.load(os.path.join(GOLDEN, "expression_aware.npz"))
    fv.set_default_precision(precision)
    register_vit_config("vit_test_patch16_224", 64, 2, 2)
    dec = fv.LatentDecomposer({i: torch.from_numpy(z[f"raw_direction/{i}"]) for i in range(7)}, 18, 64)
    vit = fv.HybridLatentViT(latent_dim=64, seq_len=36, pretrained_model_name="vit_test_patch16_224", num_classes=7,
                             use_pretrained=False, freeze_transformer=True, adapter_dim=16, verbose=False)
    vit.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}, strict=True)
    model = fv.ExpressionAwareViT(dec, vit, output_mode="concat").cuda().eval()
    logits = model(torch.from_numpy(z["x"]).cuda())
    loss = fv.cross_entropy(logits, torch.from_numpy(z["y"]).cuda())
    loss.backward()
    e_l = relerr(logits, torch.from_numpy(z["logits"]))
    grads = {k: p.grad for k, p in model.vit.named_parameters() if p.grad is not None}
    want = {k[5:] for k in z.files if k.startswith("grad/")}
    assert set(grads) == want and len(model.get_trainable_params()) == len(want)
    errs = {k: relerr(v, torch.from_numpy(z["grad/" + k])) for k, v in grads.items()}
    worst = max(errs, key=errs.get)
    record("expression_aware_vit_step", precision=precision, err_logits=e_l, err_grad=errs[worst], worst=worst)
    assert e_l < tol_l and abs(loss.item() - float(z["loss"])) < tol_l * 10
    assert errs[worst] < tol_g, (worst, errs[worst])
    assert torch.equal(logits.argmax(-1).cpu(), torch.from_numpy(z["logits"]).argmax(-1))
    with torch.no_grad():                                         # scores ride along with the same decomposer launch
        lg2, sc = model.forward_with_scores(torch.from_numpy(z["x"]).cuda())
    assert torch.equal(lg2, logits.detach()) and relerr(sc, torch.from_numpy(z["scores"])) < 1e-5


def test_graphed_mixup_train_step_matches_eager():
    """GraphedMixupTrainStep = the LatentViT trainer's whole step (batch from the packed cache with augmentation and
    mixup, mixup loss, backward, AdamW, accuracy pass) as one graph replay: bit-identical parameters, losses and
    correct-counts with the same step composed eagerly from the public pieces, including the per-replay seed."""
    import copy
    import fer_vit_b200 as fv
    g = load_golden("latent_vit")
    m1 = build_model("latent_vit", "bf16")
    m1.load_state_dict(g["sd"], strict=True)
    m1 = m1.cuda().train()                      # dropout 0.0 in this fixture: no masks drawn
    m2 = copy.deepcopy(m1)
    gen = torch.Generator().manual_seed(12)
    N, B = 40, 16
    lat = torch.randn(N, 18, 64, generator=gen)
    lab = torch.randint(0, 7, (N,), generator=gen)
    aug = fv.LatentAugment(0.1, (0.9, 1.1), 0.1)
    cache = fv.PackedLatentCache(lat, lab, aug)
    crit = fv.CrossEntropyLoss(torch.tensor([1.0, 2.0, 0.5, 1.0, 1.5, 0.7, 1.2]).cuda(), 0.1)
    o1 = fv.FusedAdamW(m1.parameters(), lr=1e-3)
    o2 = fv.FusedAdamW(m2.parameters(), lr=1e-3)
    stepper = fv.GraphedMixupTrainStep(m2, o2, cache, B, crit, seed=1000, warmup=3)
    assert stepper.launches_per_replay > 30

    def eager(idx, mix, lam, k):                # k = value of the device seed counter during that step
        mixed, labels = cache.batch(idx, mix, lam, seed=1000 + k)
        o1.zero_grad(set_to_none=True)
        loss = crit.mixup(m1(mixed), labels, mix, lam)
        loss.backward()
        o1.step()
        with torch.no_grad():
            clean, _ = cache.batch(idx, None, 1.0, seed=1000 + k)
            correct = (m1(clean).argmax(1) == labels).sum()
        return float(loss.detach()), int(correct)

    ar = torch.arange(B, device="cuda")
    for k in (1, 2, 3):                         # the three warm-up steps of the constructor (identity batch, lam = 1)
        eager(ar % N, ar, 1.0, k)
    for s, lam in enumerate((0.348, 0.9, 0.05)):
        idx = torch.randint(0, N, (B,), generator=gen).cuda()
        mix = torch.randperm(B, generator=gen).cuda()
        l1, c1 = eager(idx, mix, float(np.float32(lam)), 4 + s)
        l2, c2 = stepper(idx, mix, lam)
        assert (l1, c1) == (float(l2), int(c2)), (s, l1, float(l2), c1, int(c2))
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), k
    cache.check()
