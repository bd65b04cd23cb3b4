"""CPU: the data-side prologue of the LatentViT train step - the oracle's restatements against the
golden vectors the UNMODIFIED reference produced (tests/golden/make_golden_data_path.py), the definition of the
counter-based draws, and the host layer's error behaviour (no CPU path)."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_math as R
from tests.util import GOLDEN, load_golden, relerr

CASES = ["all", "noise", "scale_mask", "off"]


def load_augment(case):
    z = np.load(os.path.join(GOLDEN, "latent_augment.npz"))
    g = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(case + "/")}
    rng = tuple(float(v) for v in g["scale_range"]) if int(g["use_scale"]) else None
    return g, float(g["noise_std"]), rng, float(g["mask_prob"])


@pytest.mark.parametrize("case", CASES)
def test_oracle_latent_augment_matches_reference_golden(case):
    g, std, rng, p = load_augment(case)
    x = torch.from_numpy(g["x"])
    out = R.latent_augment(x, std, rng, p, torch.from_numpy(g["normal"]), torch.tensor(float(g["scale"])),
                           torch.from_numpy(g["keep"]))
    assert torch.equal(out, torch.from_numpy(g["out"]))          # fp32: bit-exact with the reference
    if case == "off":
        assert torch.equal(out, x)
    # batched form (one scale factor per sample) reduces to the per-sample form
    xb = torch.stack([x, 2 * x])
    nb = torch.stack([torch.from_numpy(g["normal"])] * 2)
    kb = torch.stack([torch.from_numpy(g["keep"])] * 2)
    ob = R.latent_augment(xb, std, rng, p, nb, torch.tensor([float(g["scale"]), 1.0]), kb)
    assert torch.equal(ob[0], out)


def test_oracle_mixup_step_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN, "mixup_step.npz"))
    g = load_golden("latent_vit")
    x, y, index = torch.from_numpy(z["x"]), torch.from_numpy(z["y"]), torch.from_numpy(z["index"])
    lam, w, eps = float(z["lam"]), torch.from_numpy(z["class_weight"]), float(z["label_smoothing"])
    for dtype, tol_loss, tol_g in ((torch.float64, 1e-9, 1e-5), (torch.float32, 1e-5, 2e-4)):
        sd = {k: (v.to(dtype).requires_grad_(True) if v.is_floating_point() else v) for k, v in g["sd"].items()}
        logits = R.latent_vit_forward(sd, R.mixup(x.to(dtype), index, lam), 2, 2)
        loss = R.mixup_loss(logits, y, index, lam, w.to(dtype), eps)
        grads = R.grads_of(loss, sd)
        assert abs(loss.item() - float(z["loss"])) < tol_loss
        for k in grads:                                           # golden grads are stored in fp32
            assert relerr(grads[k], torch.from_numpy(z["grad/" + k])) < tol_g, k
        with torch.no_grad():                                     # the trainer's extra forward on the un-mixed batch
            pred = R.latent_vit_forward(sd, x.to(dtype), 2, 2).argmax(-1)
        assert torch.equal(pred, torch.from_numpy(z["pred"]))
        assert abs((pred == y).double().mean().item() - float(z["accuracy"])) < 1e-12


def test_mixup_loss_limits():
    torch.manual_seed(1)
    zl = torch.randn(9, 7, dtype=torch.float64)
    y = torch.randint(0, 7, (9,))
    idx = torch.randperm(9)
    w = torch.rand(7, dtype=torch.float64) + 0.5
    assert torch.allclose(R.mixup_loss(zl, y, idx, 1.0, w, 0.1), R.cross_entropy(zl, y, w, 0.1))
    assert torch.allclose(R.mixup_loss(zl, y, idx, 0.0, w, 0.1), R.cross_entropy(zl, y[idx], w, 0.1))
    assert torch.equal(R.mixup(zl, idx, 1.0), zl)


def test_counter_based_draws_definition():
    """Deterministic, position-keyed, and distributed as LatentAugment's draws (N(0,1), U(lo,hi), Bernoulli(1-p))."""
    n, s, k = R.latent_augment_draws(42, 32, 9216, (0.9, 1.1), 0.1)
    n2, s2, k2 = R.latent_augment_draws(42, 32, 9216, (0.9, 1.1), 0.1)
    assert torch.equal(n, n2) and torch.equal(s, s2) and torch.equal(k, k2)
    n3, _, k3 = R.latent_augment_draws(43, 32, 9216, (0.9, 1.1), 0.1)
    assert not torch.equal(n, n3) and not torch.equal(k, k3)
    # a prefix of the batch has the same draws: they depend on the batch position only
    n4, s4, k4 = R.latent_augment_draws(42, 5, 9216, (0.9, 1.1), 0.1)
    assert torch.equal(n4, n[:5]) and torch.equal(s4, s[:5]) and torch.equal(k4, k[:5])
    assert abs(n.mean().item()) < 5e-3 and abs(n.std().item() - 1) < 5e-3
    assert abs((n ** 4).mean().item() - 3) < 0.05                # Gaussian kurtosis
    assert abs((n[:, 0::2] * n[:, 1::2]).mean().item()) < 5e-3   # the two normals of a Box-Muller pair are uncorrelated
    assert 0.9 <= s.min().item() and s.max().item() <= 1.1 and s.std().item() > 0.03
    assert abs(k.double().mean().item() - 0.9) < 2e-3
    assert abs((k[:, :-1] & k[:, 1:]).double().mean().item() - 0.81) < 3e-3
    # the 32-bit mix is the kernels' generator: known answers, so a change on either side is caught on the CPU
    assert [int(v) for v in R.mix_hash(0, 0, np.arange(3))] == [int(v) for v in R.mix_hash(0, 0, [0, 1, 2])]
    assert int(R.mix_hash(1234, 0x4C410003, 77)) == int(R.mix_hash(1234, 0x4C410003, [77])[0])


def test_data_path_has_no_cpu_fallback(tmp_path):
    import fer_vit_b200 as fv
    x = torch.randn(4, 18, 512)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fv.latent_batch(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fv.LatentAugment(noise_std=0.1)(x[0])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fv.mixup(x, torch.randperm(4), 0.3)
    with pytest.raises(RuntimeError, match="CUDA"):
        fv.mixup_cross_entropy(torch.randn(4, 7, requires_grad=True), torch.zeros(4, dtype=torch.long),
                               torch.randperm(4), 0.3)
    # directory errors of the reference dataset (data/latent_dataset.py:75-86)
    with pytest.raises(FileNotFoundError):
        fv.PackedLatentCache.from_dir(str(tmp_path / "missing"))
    with pytest.raises(ValueError, match="No .pt files"):
        fv.PackedLatentCache.from_dir(str(tmp_path))
    torch.save({"latent": torch.randn(18, 512), "label": 3, "img_path": "a.png"}, tmp_path / "a.pt")
    if not torch.cuda.is_available():
        with pytest.raises((RuntimeError, AssertionError)):
            fv.PackedLatentCache.from_dir(str(tmp_path))         # no GPU: refuses instead of keeping a host copy
    # same constructor surface as the reference transform factories (data/latent_dataset.py:138-162)
    t = fv.get_latent_train_transforms()
    assert (t.noise_std, t.scale_range, t.mask_prob) == (0.1, (0.9, 1.1), 0.1)
    assert fv.get_latent_val_transforms() is None
    assert fv.LatentAugment().params() is None


# ------------------------------------------------------------------------------------------------
# LatentDecomposer / ExpressionAwareViT front-end
# ------------------------------------------------------------------------------------------------
DEC_MODES = [(d, o) for d in ("all_classes", "max_class") for o in ("expr_only", "id_only", "enhanced", "concat")]


@pytest.mark.parametrize("dm,om", DEC_MODES)
def test_oracle_decomposer_matches_reference_golden(dm, om):
    z = np.load(os.path.join(GOLDEN, "expression_aware.npz"))
    x, dirs = torch.from_numpy(z["x"]), torch.from_numpy(z["directions"])
    raw = torch.stack([torch.from_numpy(z[f"raw_direction/{i}"]) for i in range(7)])
    assert relerr(R.normalize_directions(raw), dirs) < 1e-7
    want = torch.from_numpy(z[f"out/{dm}/{om}"])
    assert relerr(R.decomposer_forward(x.double(), dirs.double(), om, float(z["alpha"]), dm), want) < 2e-7
    assert relerr(R.decomposer_forward(x, dirs, om, float(z["alpha"]), dm), want) < 2e-6
    e, i, coef = R.latent_decompose(x.double(), dirs.double(), dm)
    assert relerr(coef, torch.from_numpy(z["scores"])) < 2e-7
    assert relerr(e + i, x) < 1e-15                                 # the two parts always add back to w+
    if dm == "all_classes":                                         # identity part carries no expression score (up
        # to the directions' mutual overlap: exact only for orthogonal directions) - projecting twice is stable
        e2, _, _ = R.latent_decompose(e, dirs.double(), dm)
        assert e2.shape == e.shape


def test_oracle_expression_aware_step_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN, "expression_aware.npz"))
    gk = [k[5:] for k in z.files if k.startswith("grad/")]
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    sd = {k: (v.double().requires_grad_(k in gk) if v.is_floating_point() else v) for k, v in sd.items()}
    x = R.decomposer_forward(torch.from_numpy(z["x"]).double(), torch.from_numpy(z["directions"]).double(), "concat")
    assert x.shape[1] == 36
    logits = R.hybrid_forward(sd, x, 2, 2, True)
    loss = R.cross_entropy(logits, torch.from_numpy(z["y"]))
    grads = R.grads_of(loss, sd)
    assert relerr(logits, torch.from_numpy(z["logits"])) < 1e-6
    assert abs(loss.item() - float(z["loss"])) < 1e-6
    for k in gk:
        assert relerr(grads[k], torch.from_numpy(z["grad/" + k])) < 1e-5, k


def test_decomposer_host_surface():
    import fer_vit_b200 as fv
    z = np.load(os.path.join(GOLDEN, "expression_aware.npz"))
    raw = {i: torch.from_numpy(z[f"raw_direction/{i}"]) for i in range(7)}
    d = fv.LatentDecomposer(raw, 18, 64)
    assert (d.seq_len, d.latent_dim, d.num_classes) == (18, 64, 7)
    assert list(d.state_dict()) == ["directions"] and not list(d.parameters())
    assert relerr(d.directions, torch.from_numpy(z["directions"])) < 1e-7
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d(torch.randn(2, 18, 64))
    with pytest.raises(ValueError, match="Unknown output_mode"):
        d(torch.randn(2, 18, 64), output_mode="nope")
    with pytest.raises(ValueError, match="Unknown mode"):
        d.decompose(torch.randn(2, 18, 64), mode="nope")


def test_dropout_generator_statistics():
    """drop_hash (two 32-bit rounds under a 64-bit key): keep rate, independence of neighbouring elements, of
    consecutive seeds (a graph replay = seed + 1) and of sites, on 2^20 elements; bit balance of the raw hash."""
    n = 1 << 20
    idx = np.arange(n)
    for p in (0.1, 0.5):
        a = R.dropout_keep(1234, 8, idx, p).numpy()
        b = R.dropout_keep(1235, 8, idx, p).numpy()
        c = R.dropout_keep(1234, 9, idx, p).numpy()
        tol = 5 * np.sqrt(p * (1 - p) / n)
        for m in (a, b, c):
            assert abs(m.mean() - (1 - p)) < tol
        for u, v in ((a[:-1], a[1:]), (a[:-197], a[197:]), (a, b), (a, c), (a[:-1], b[1:]), (a[1:], b[:-1])):
            corr = np.corrcoef(u.astype(np.float64), v.astype(np.float64))[0, 1]
            assert abs(corr) < 5 / np.sqrt(n), corr
    h = R.drop_hash(99, 3, idx)
    for bit in range(32):
        frac = float(((h >> np.uint32(bit)) & np.uint32(1)).mean())
        assert abs(frac - 0.5) < 5 * 0.5 / np.sqrt(n), (bit, frac)
    # rows of an attention matrix (index = row * S + col) must not repeat each other's masks
    rows = R.dropout_keep(7, 0, idx[:197 * 197], 0.1).numpy().reshape(197, 197).astype(np.float64)
    cc = np.corrcoef(rows)
    off = cc[~np.eye(197, dtype=bool)]
    assert np.abs(off).max() < 6 / np.sqrt(197) and abs(off.mean()) < 0.01


def test_generator_restatement_matches_the_c_source(tmp_path):
    """The kernels' counter-based generator (csrc/common.cuh, __host__ __device__) compiled for the HOST and run here:
    the oracle's numpy mix_hash64 / mix_hash give the same 64- and 32-bit values, and the dropout threshold rounding is
    the one the oracle's mask replay assumes. Pins the draws of dropout, LatentAugment and mixup without a GPU."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "mix_hash_host")
    subprocess.run([nvcc, "-O1", "-std=c++17", "--expt-relaxed-constexpr", "-I", os.path.join(root, "fer_vit_b200", "csrc"),
                    "-I", os.path.join(root, "include"), os.path.join(root, "tests", "host", "mix_hash_host.cu"),
                    "-o", exe], check=True, capture_output=True, timeout=300)
    cases = [(0, 0, 0), (42, 3, 17), (0xC0FFEE, 0x4C410000, 123456789), (2 ** 63 + 5, 0x4C410003, 2 ** 40 + 7),
             (0xFFFFFFFFFFFFFFFF, 0xFFFF1, 0xFFFFFFFF), (1000, 0x4C410002, 31)]
    args = [str(v) for c in cases for v in c]
    lines = subprocess.run([exe, *args], check=True, capture_output=True, text=True, timeout=60).stdout.split("\n")
    for (seed, site, idx), line in zip(cases, lines):
        h64, h32 = (int(v) for v in line.split())
        assert int(R.mix_hash64(seed, site, [idx])[0]) == h64, (seed, site, idx)
        assert int(R.mix_hash(seed, site, [idx])[0]) == h32
        assert h32 == h64 >> 32
    hline = lines[len(cases)].split()
    assert hline[0] == "H"
    for k, idx in enumerate((0, 77, 0xFFFFFFFF, 2 ** 40 + 7)):
        assert int(R.drop_hash(0xC0FFEE, 0x4C41 + k, [idx])[0]) == int(hline[1 + k]), idx
    lines = lines[:len(cases)] + lines[len(cases) + 1:]
    for line, p in zip(lines[len(cases):len(cases) + 2], (0.1, 0.5)):
        keep = R.dropout_keep(12345, 9, np.arange(1000, 1064), p)
        assert line.split()[1] == "".join("1" if k else "0" for k in keep.tolist()), p
    # GELU polynomials of the GEMM epilogues vs the exact erf forms on [-8, 8], step 1e-4: the bounds common.cuh states
    # (absolute; |gelu| error inside the clamp is u * 7e-6 <= 2.8e-5)
    g_in, g_tail, d_in, d_tail = (float(v) for v in lines[len(cases) + 2].split()[1:])
    assert g_in < 3e-5 and g_tail < 3e-4 and d_in < 7e-5 and d_tail < 6e-4, (g_in, g_tail, d_in, d_tail)
    # the packed forward epilogue (min on u^2 + saturating cdf; value and derivative): |u| > 4 is exact to the cdf's tail
    e_in, e_tail, ed_in, ed_tail = (float(v) for v in lines[len(cases) + 3].split()[1:])
    assert lines[len(cases) + 3].startswith("E ")
    assert e_in < 3e-5 and e_tail < 3e-4 and ed_in < 3e-5 and ed_tail < 3e-4, (e_in, e_tail, ed_in, ed_tail)
    thr = [int(v) for v in lines[len(cases) + 4].split()[1:]]
    assert thr == [int(float(np.float32(p)) * 4294967296.0) for p in (0.1, 0.5, 0.999)]


def test_stream_k_schedule_invariants(tmp_path):
    """The stream-K work decomposition the GEMM kernel runs (csrc/gemm_tc2_sched.cuh), replayed on the host for every
    tile count up to 260 x six K depths x three pair counts: every k-block of every tile exactly once, at most two
    segments per pair with the head first, one partial (one scratch slot) per pair, and the owner's expected partial
    count / writer range equal to what the other pairs actually produce - the conditions under which the device
    protocol is deadlock-free and deterministic."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "streamk_sched_host")
    subprocess.run([nvcc, "-O1", "-std=c++17", "-I", os.path.join(root, "fer_vit_b200", "csrc"),
                    os.path.join(root, "tests", "host", "streamk_sched_host.cu"), "-o", exe], check=True,
                   capture_output=True, timeout=300)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120).stdout.strip()
    assert out.startswith("OK "), out
    assert int(out.split()[1]) > 500
    assert int(out.split()[2]) == 1200      # tail splitting (no stream-K): every column of every tile exactly once


def test_oracle_optimizer_step_matches_reference_trainer_golden():
    """One batch of the reference's train_latent_vit_v2.train_epoch (mixup, weighted smoothed CE, clip_grad_norm_ active,
    AdamW): the oracle's forward + mixup loss + gradients + `clip_grad_norm` + `adamw_step` reproduce the loss, the
    gradient norm and EVERY parameter update of the unmodified trainer."""
    z = np.load(os.path.join(GOLDEN, "v2_train_epoch.npz"))
    g = load_golden("latent_vit_v2")
    after = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("after/")}
    sd = {k: (v.double().requires_grad_(k in after and k != "spe.groups") if v.is_floating_point() else v)
          for k, v in g["sd"].items()}
    x, y, index = torch.from_numpy(z["x"]).double(), torch.from_numpy(z["y"]), torch.from_numpy(z["index"])
    lam, w = float(z["lam"]), torch.from_numpy(z["class_weight"]).double()
    logits = R.latent_vit_v2_forward(sd, R.mixup(x, index, lam), 2, 2, True, True, True, True)
    loss = R.mixup_loss(logits, y, index, lam, w, float(z["label_smoothing"]))
    assert abs(loss.item() - float(z["loss"])) < 1e-10
    grads = R.grads_of(loss, sd)
    clipped, total = R.clip_grad_norm(grads, float(z["grad_clip"]))
    assert abs(float(total) - float(z["total_norm"])) < 1e-9 and float(total) > float(z["grad_clip"])
    assert set(clipped) == set(after)
    for k, gr in clipped.items():
        p0 = sd[k].detach()
        p1, m, v = R.adamw_step(p0, gr, torch.zeros_like(p0), torch.zeros_like(p0), 1, float(z["lr"]), (0.9, 0.999), 1e-8,
                                float(z["weight_decay"]))
        assert relerr(p1 - p0, after[k] - p0) < 1e-7, k
    # second step of the restatement against torch's own AdamW (state carried over, no clipping)
    p = torch.randn(5, 7, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    tp = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([tp], lr=3e-3, betas=(0.8, 0.95), eps=1e-7, weight_decay=0.1)
    m, v, q = torch.zeros_like(p), torch.zeros_like(p), p.clone()
    for step in (1, 2, 3):
        gr = torch.randn(5, 7, dtype=torch.float64, generator=torch.Generator().manual_seed(10 + step))
        tp.grad = gr.clone()
        opt.step()
        q, m, v = R.adamw_step(q, gr, m, v, step, 3e-3, (0.8, 0.95), 1e-7, 0.1)
        assert relerr(q, tp.detach()) < 1e-14


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="needs the reference tree")
def test_packed_cache_reads_a_directory_like_the_reference_dataset(tmp_path):
    """PackedLatentCache.read_dir (the host half of from_dir) beside the unmodified LatentFERDataset on the same
    directory of per-sample .pt files: same order (sorted names, non-.pt files ignored), same latents and labels, same
    class counts; a corrupt file raises the reference's RuntimeError."""
    import importlib
    import sys
    import fer_vit_b200 as fv
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    ds_mod = importlib.import_module("data.latent_dataset")
    assert ds_mod.__file__.startswith("/root/reference")
    g = torch.Generator().manual_seed(6)
    names = ["img_0010.pt", "img_0002.pt", "zz.pt", "a_first.pt", "img_0001.pt"]
    for i, n in enumerate(names):
        torch.save({"latent": torch.randn(18, 512, generator=g), "label": int(i * 3 % 7), "img_path": n + ".png"},
                   tmp_path / n)
    (tmp_path / "notes.txt").write_text("not a latent")
    ref = ds_mod.LatentFERDataset(str(tmp_path))
    lat, lab = fv.PackedLatentCache.read_dir(str(tmp_path))
    assert len(ref) == lat.shape[0] == len(names)
    for i in range(len(ref)):
        x, y = ref[i]
        assert torch.equal(lat[i], x) and int(lab[i]) == y
    counts = {int(v): int(c) for v, c in zip(*torch.unique(lab, return_counts=True))}
    assert counts == ref.get_class_counts()
    assert fv.PackedLatentCache.CLASS_NAMES == ref.get_class_names()
    (tmp_path / "broken.pt").write_bytes(b"not a checkpoint")
    with pytest.raises(RuntimeError, match="Error loading"):
        fv.PackedLatentCache.read_dir(str(tmp_path))
    with pytest.raises(RuntimeError, match="Error loading"):
        ds_mod.LatentFERDataset(str(tmp_path))[1]          # 'broken.pt' sorts second


def test_oracle_hybrid_trainer_step_matches_reference_golden():
    """The headline configuration's whole optimizer step as the reference trainer runs it
    (train_hybrid_latent_vit.train_epoch in train mode: head Dropout(0.1) with the mask torch drew; AdamW over the
    trainer's own get_optimizer_groups - lr x10 / x10 / x10 / x5, no decay on pos_embed + cls_token; frozen blocks):
    the oracle reproduces the loss, the train accuracy and every trainable tensor's update."""
    z = np.load(os.path.join(GOLDEN, "hybrid_train_epoch.npz"))
    g = load_golden("hybrid_adapter")
    hyper = {k[6:]: tuple(float(v) for v in z[k]) for k in z.files if k.startswith("hyper/")}
    assert len(hyper) == 18 and not any(k.startswith("transformer.") for k in hyper)
    assert hyper["pos_embed"] == (5e-4, 0.0) and hyper["head.2.weight"] == (1e-3, 0.01)
    assert hyper["adapters.0.alpha"] == (1e-3, 0.01) and hyper["input_proj.weight"] == (1e-3, 0.01)
    sd = {k: (v.double().requires_grad_(k in hyper) if v.is_floating_point() else v) for k, v in g["sd"].items()}
    y = torch.from_numpy(z["y"])
    logits = R.hybrid_forward(sd, torch.from_numpy(z["x"]).double(), 2, 2, True,
                              {"head": torch.from_numpy(z["head_mask"])})
    loss = R.cross_entropy(logits, y)
    assert abs(loss.item() - float(z["loss"])) < 1e-10
    assert abs((logits.argmax(-1) == y).double().mean().item() - float(z["accuracy"])) < 1e-12
    grads = R.grads_of(loss, sd)
    for k, (lr, wd) in hyper.items():
        p0 = sd[k].detach()
        p1, _, _ = R.adamw_step(p0, grads[k], torch.zeros_like(p0), torch.zeros_like(p0), 1, lr, (0.9, 0.999), 1e-8, wd)
        assert relerr(p1 - p0, torch.from_numpy(z["after/" + k]) - p0) < 1e-7, k
