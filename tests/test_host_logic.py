"""CPU: host-side logic of the drop-in layer — checkpoint/key compatibility with the reference, plan descriptions,
failure behaviour without CUDA, and the data-parallel gradient bucketing on gloo with world_size 2."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import FIXTURES, build_model, load_golden


@pytest.mark.parametrize("name", FIXTURES)
def test_state_dict_is_interchangeable_with_the_reference(name):
    """Same keys, shapes and dtypes as the checkpoint the UNMODIFIED reference class wrote (golden fixture)."""
    g = load_golden(name)
    model = build_model(name, "fp32")
    sd = model.state_dict()
    assert set(sd) == set(g["sd"])
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(g["sd"][k].shape), k
        assert v.dtype == g["sd"][k].dtype, k
    model.load_state_dict(g["sd"], strict=True)
    # requires_grad pattern = set of tensors the reference produced gradients for
    trainable = {k for k, p in model.named_parameters() if p.requires_grad}
    assert trainable == set(g["grad"])


def test_parameter_counts_match_the_survey():
    import fer_vit_b200 as fv
    assert sum(p.numel() for p in fv.LatentViT().parameters()) == 19_191_815
    v2 = fv.LatentViTv2(use_lwn=True, use_lwn_residual=True, use_spe=True, use_leam=True)
    assert sum(p.numel() for p in v2.parameters()) == 19_221_035
    hy = fv.create_hybrid_latent_vit(model_size="base", use_pretrained=False, freeze_transformer=True,
                                     use_adapter=True, adapter_dim=64)
    assert sum(p.numel() for p in hy.parameters() if p.requires_grad) == 1_605_907
    assert sum(p.numel() for p in hy.transformer.parameters()) == 85_054_464
    assert hy.pos_embed.shape == (1, 19, 768) and hy.embed_dim == 768 and hy.use_adapter


def test_hybrid_api_surface_used_by_the_reference_callers():
    """train_hybrid_latent_vit.py:63-117 and evaluate_model.py:231-296 touch these attributes."""
    import fer_vit_b200 as fv
    m = fv.create_hybrid_latent_vit(model_size="tiny", use_pretrained=False, freeze_stages=6, use_adapter=False)
    assert len(m.transformer) == 12 and hasattr(m.transformer[0], "attn")
    assert not any(p.requires_grad for p in m.transformer[5].parameters())
    assert all(p.requires_grad for p in m.transformer[6].parameters())
    m.unfreeze_all()
    assert all(p.requires_grad for p in m.parameters())
    for attr in ("input_proj", "cls_token", "pos_embed", "head", "latent_dim", "seq_len", "num_classes",
                 "pretrained_model_name"):
        assert hasattr(m, attr)
    # the blocks container is callable on CPU tensors (compatibility surface of evaluate_model.py:255)
    out = m.transformer(torch.randn(2, 19, 192))
    assert out.shape == (2, 19, 192)
    assert set(fv.RECOMMENDED_STRATEGIES) == {"full_finetune", "partial_freeze", "adapter", "linear_probe"}
    with pytest.raises(ImportError):
        fv.create_hybrid_latent_vit(model_size="tiny", use_pretrained=True)   # needs timm + its download


def test_position_embedding_interpolation_matches_reference_formula():
    """cls row kept, 196 patch rows linearly resampled to seq_len rows (hybrid_latent_vit.py:130-152)."""
    import torch.nn.functional as F
    import fer_vit_b200 as fv
    from fer_vit_b200.models_fer_vit.vit_blocks import VisionTransformerShell
    torch.manual_seed(0)
    vit = VisionTransformerShell("vit_tiny_patch16_224")
    m = fv.HybridLatentViT.__new__(fv.HybridLatentViT)
    torch.nn.Module.__init__(m)
    m.embed_dim = 192
    pos = m._init_position_embedding(vit, 18)
    ref = torch.cat([vit.pos_embed[:, :1], F.interpolate(vit.pos_embed[:, 1:].permute(0, 2, 1), size=18, mode="linear",
                                                         align_corners=False).permute(0, 2, 1)], dim=1)
    assert torch.equal(pos.data, ref.data)


def test_plan_description_and_precision_switch():
    from fer_vit_b200 import _lib as L
    model = build_model("hybrid_adapter", "bf16")
    cfg = model._plan_config()
    assert (cfg.norm_first, cfg.act, cfg.adapter_dim, cfg.E, cfg.depth, cfg.H, cfg.F) == (1, L.ACT_GELU, 16, 64, 2, 2, 256)
    assert abs(cfg.eps_block - 1e-6) < 1e-12 and abs(cfg.head_dropout - 0.1) < 1e-7
    runner = model.plan_runner()
    assert runner.bf16 and runner.nstages == 4
    slots = [s for stage in runner.stage_slots for s in stage]
    assert len(slots) == len(set(slots))                      # every slot lives in exactly one backward stage
    model.precision = "fp32"
    assert not model.plan_runner().bf16
    lat = build_model("latent_vit", "fp32")._plan_config()
    assert (lat.norm_first, lat.act) == (0, L.ACT_RELU) and abs(lat.eps_block - 1e-5) < 1e-9
    img = build_model("image_vit", "fp32")._plan_config()
    assert (img.input_kind, img.L, img.Din, img.input_dropout) == (1, 4, 768, 1)
    with pytest.raises(ValueError):
        model.precision = "fp16"


def test_runner_is_not_part_of_model_state():
    model = build_model("latent_vit_v2", "fp32")
    model.plan_runner()
    clone = copy.deepcopy(model)
    assert clone.__dict__["_runner"] is None
    assert not any(k.startswith("_") for k in model.state_dict())
    assert clone.get_config()["model"] == "LatentViTv2" and clone.get_leam_weights().shape == (18,)


def test_cpu_tensors_are_rejected_loudly():
    model = build_model("latent_vit", "fp32")
    with pytest.raises(RuntimeError, match="no CPU"):
        model(torch.randn(2, 18, 64))
    import fer_vit_b200 as fv
    with pytest.raises(RuntimeError, match="no CPU"):
        fv.cross_entropy(torch.randn(4, 7, requires_grad=True), torch.zeros(4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU"):
        fv.LEAM()(torch.randn(2, 18, 512))


# ----------------------------------------------------------------------------------------------------------
# data parallel host logic on gloo, world_size 2
# ----------------------------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fer_vit_b200.parallel import GradBucketer, global_ce_denominator, sync_parameters
    b = GradBucketer(num_buckets=3)
    groups = b.stage_groups(14)                      # ViT-B: head + 12 blocks + input stage
    assert groups[0][0] == 0 and groups[-1][1] == 14 and all(a[1] == c[0] for a, c in zip(groups, groups[1:]))
    # a fake flat gradient buffer in backward order, reduced bucket by bucket like PlanRunner.backward does
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    bounds = [0, 300, 650, 1000]
    for lo, hi in zip(bounds, bounds[1:]):
        b.reduce_async(flat[lo:hi])
    b.finish()
    expect = torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(flat, expect)
    # parameters broadcast from rank 0
    lin = torch.nn.Linear(4, 3)
    with torch.no_grad():
        lin.weight.fill_(float(rank + 1))
    sync_parameters(lin)
    ok = ok and bool((lin.weight == 1.0).all())
    # class-weighted CE: local denominator = global sum / world
    w = torch.tensor([1.0, 2.0, 3.0])
    labels = torch.tensor([0, 1]) if rank == 0 else torch.tensor([2, 2])
    den = global_ce_denominator(labels, w)
    ok = ok and abs(den.item() - (1 + 2 + 3 + 3) / world) < 1e-6
    ok = ok and b.calls == 3 and b.bytes_reduced == 4000
    out[rank] = ok
    dist.destroy_process_group()


def test_grad_bucketer_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out[0] and out[1]


def test_weighted_ce_is_rank_invariant_under_dp():
    """Oracle-level check of the DP rule: averaging the per-rank gradients computed with den = global/world equals
    the single-process gradient of the class-weighted mean loss."""
    from oracle import reference_math as R
    torch.manual_seed(0)
    z = torch.randn(8, 7, dtype=torch.float64, requires_grad=True)
    y = torch.randint(0, 7, (8,))
    w = torch.rand(7, dtype=torch.float64) + 0.5
    full, = torch.autograd.grad(R.cross_entropy(z, y, w), z)
    den_local = w[y].sum() / 2
    parts = []
    for sl in (slice(0, 4), slice(4, 8)):
        zz = z[sl].detach().requires_grad_(True)
        lp = torch.log_softmax(zz, -1)
        loss = -(lp[torch.arange(4), y[sl]] * w[y[sl]]).sum() / den_local
        parts.append(torch.autograd.grad(loss, zz)[0] / 2)    # all-reduce AVG
    assert torch.allclose(torch.cat(parts), full, atol=1e-12)


def test_fused_adamw_has_no_cpu_fallback():
    import fer_vit_b200 as fv
    p = torch.nn.Parameter(torch.randn(4, 4))
    opt = fv.FusedAdamW([p], lr=1e-3)
    assert opt.defaults["capturable"] is True
    p.grad = torch.randn(4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm) runs without a GPU and prints ONE
    JSON line with the contract's keys: same metric / unit / config as the product arm, impl = reference, a cpu_baseline
    describing the run and an e2e object repeating the value with zero host<->device bytes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--batch", "32"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_launch_summary_matches_its_launch_list():
    """profiles/: the per-kernel summary of the final step is what tools/ncu_summary.py derives from the committed ncu
    launch list (the GEMM share of the step quoted in bench.py / DESIGN.md comes from it)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = os.path.join(root, "profiles", "r01_launches_step_final.csv")
    want = open(os.path.join(root, "profiles", "r01_launches_step_final_summary.txt")).read().strip().splitlines()
    got = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), csv_path], capture_output=True,
                         text=True, timeout=120).stdout.strip().splitlines()
    assert got[0] == want[0]
    assert got[1:6] == want[1:6]


@pytest.mark.skipif(not os.path.isdir("/root/reference/train"), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("frozen,adapter", [(True, True), (False, True), (False, False)])
def test_reference_optimizer_grouping_runs_on_the_drop_in_model(frozen, adapter):
    """The reference trainer's OWN `get_optimizer_groups` (train/train_hybrid_latent_vit.py:63-117), imported unmodified,
    applied to the drop-in HybridLatentViT: it finds input_proj / transformer / adapters / head / pos_embed + cls_token
    exactly as on the reference model (same groups, same parameter counts, same learning rates), and the groups feed
    FusedAdamW's per-group hyper-parameter table."""
    import importlib
    import sys
    import fer_vit_b200 as fv
    from fer_vit_b200.models_fer_vit.vit_blocks import register_vit_config
    from oracle import timm_shim
    timm_shim.install()
    for pth in ("/root/reference", "/root/reference/train"):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    tr = importlib.import_module("train.train_hybrid_latent_vit")
    assert tr.__file__.startswith("/root/reference")
    register_vit_config("vit_test_patch16_224", 64, 2, 2)
    kw = dict(latent_dim=64, seq_len=18, pretrained_model_name="vit_test_patch16_224", num_classes=7,
              use_pretrained=False, freeze_transformer=frozen, adapter_dim=16 if adapter else None)
    ours = fv.HybridLatentViT(verbose=False, **kw)
    ref = tr.HybridLatentViT(**kw)
    g_ours = tr.get_optimizer_groups(ours, 1e-4, 0.01)
    g_ref = tr.get_optimizer_groups(ref, 1e-4, 0.01)
    assert len(g_ours) == len(g_ref) == (3 + (0 if frozen else 1) + (1 if adapter else 0))
    for a, b in zip(g_ours, g_ref):
        assert a["lr"] == b["lr"] and a["weight_decay"] == b["weight_decay"]
        assert [tuple(p.shape) for p in a["params"]] == [tuple(p.shape) for p in b["params"]]
        # the reference builds the input-projection and head groups from .parameters() without a requires_grad filter
        assert [p.requires_grad for p in a["params"]] == [p.requires_grad for p in b["params"]]
    opt = fv.FusedAdamW(g_ours, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0)
    assert [r[0] for r in opt._hyper_rows()] == [g["lr"] for g in g_ref]
    assert opt._hyper_rows()[-1][4] == 0.0          # pos_embed / cls_token: no weight decay


@pytest.mark.skipif(not os.path.isdir("/root/reference/models_fer_vit"), reason="needs the reference tree")
@pytest.mark.parametrize("output_mode", ["expr_only", "concat"])
def test_expression_aware_factory_matches_the_reference(tmp_path, output_mode):
    """ExpressionAwareViT.from_config on a directions file (the format compute_expression_directions.py writes), beside
    the unmodified reference factory: same token count ('concat' doubles it), same state_dict keys and shapes, same
    trainable-parameter list, same normalised directions."""
    import importlib
    import sys
    import fer_vit_b200 as fv
    from oracle import timm_shim
    timm_shim.install()
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    ref_mod = importlib.import_module("models_fer_vit.expression_aware_vit")
    assert ref_mod.__file__.startswith("/root/reference")
    g = torch.Generator().manual_seed(4)
    path = str(tmp_path / "binary_directions.pt")
    torch.save({"directions": {i: torch.randn(18, 512, generator=g) for i in range(7)}, "seq_len": 18, "latent_dim": 512,
                "method": "binary"}, path)
    kw = dict(model_size="tiny", num_classes=7, use_pretrained=False, freeze_transformer=True, use_adapter=True,
              adapter_dim=8, output_mode=output_mode, enhance_alpha=1.5, decompose_mode="max_class")
    ours = fv.ExpressionAwareViT.from_config(path, **kw)
    ref = ref_mod.ExpressionAwareViT.from_config(path, **kw)
    assert ours.vit.seq_len == ref.vit.seq_len == (36 if output_mode == "concat" else 18)
    assert (ours.output_mode, ours.enhance_alpha, ours.decompose_mode) == (ref.output_mode, ref.enhance_alpha,
                                                                           ref.decompose_mode)
    so, sr = ours.state_dict(), ref.state_dict()
    assert list(so) == list(sr)
    assert all(so[k].shape == sr[k].shape and so[k].dtype == sr[k].dtype for k in so)
    assert torch.allclose(so["decomposer.directions"], sr["decomposer.directions"], atol=1e-7)
    assert [tuple(p.shape) for p in ours.get_trainable_params()] == [tuple(p.shape) for p in ref.get_trainable_params()]
    ref.load_state_dict(so, strict=True)            # a checkpoint of the drop-in loads into the reference, and back
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.print_info()


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may
    import it. The product package (Python and CUDA sources) must not mention it in an import, an include or a path."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#\s*include\s+[\"<].*oracle)|sys\.path.*oracle|baseline/_ref")
    bad = []
    for base, _, files in os.walk(os.path.join(root, "fer_vit_b200")):
        if "_build" in base or "__pycache__" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                for n, line in enumerate(open(os.path.join(base, f), errors="ignore"), 1):
                    if pat.search(line):
                        bad.append((os.path.join(base, f), n, line.strip()))
    assert not bad, bad
    # and the two allowed importers outside tests/ confine it to the legs named above
    bench = open(os.path.join(root, "bench.py")).read()
    for m in re.finditer(r"^(\s*)(from oracle|import oracle)", bench, re.M):
        assert len(m.group(1)) >= 4, "bench.py imports oracle at module level"
