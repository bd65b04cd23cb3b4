#!/usr/bin/env python
"""bench.py — train samples/sec of the FER-ViT train step on B200 (headline: HybridLatentViT, frozen ViT-B/16 + Adapter64).

    python bench.py --gpus 1 --steps 20 --warmup 5                      # BASELINE config 3: batch 256 on 1 B200
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                          # data parallel, 256 samples per GPU (weak scaling)
    python bench.py --config latent_vit|image_vit|latent_vit_v2         # BASELINE configs 1 / 2 / 4, same JSON contract
    python bench.py --impl reference --steps 3 --warmup 1               # the UNMODIFIED reference classes on host cores

One step = zero_grad + forward + cross-entropy + backward (+ NCCL all-reduce of the trainable gradients when N > 1)
+ AdamW on the trainable parameters, on synthetic inputs (randn) and random-init weights of the named architecture.
Prints ONE JSON line (DESIGN.md "Measurement" explains every field).

Scaling: every N runs the SAME per-GPU workload (config 3's batch 256 per GPU, `"scaling": "weak"`), so value_N /
(N * value_1) is the scaling efficiency. BASELINE config 5 (global batch 4096 split over the GPUs) is measured in the
same run and reported in the `config5` object of every line, with its own 1-GPU batch-4096 denominator measured on
rank 0 of the same box.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# per config: BASELINE.json index, default per-GPU batch, algorithmic FLOPs per sample of one train step
CONFIGS = {
    "hybrid": dict(index=3, batch=256, flop=6.658e9,
                   metric="train samples/sec (HybridLatentViT: frozen ViT-B/16 + Adapter64, S=19 tokens, bf16)",
                   workload="BASELINE config 3: HybridLatentViT frozen timm-style ViT-B/16 + Adapter(64) on w+ tokens "
                            "18x512, batch 256 per B200"),
    "latent_vit": dict(index=1, batch=32, flop=2.184e9,
                       metric="train samples/sec (LatentViT 512/d6/h8/2048, S=19 tokens)",
                       workload="BASELINE config 1: LatentViT (post-norm, ReLU, dropout 0.1) on w+ latents 18x512, "
                                "batch 32 per B200"),
    "image_vit": dict(index=2, batch=64, flop=24.05e9,
                      metric="train samples/sec (ImageViT 512/d6/h8/2048, 224x224 images, S=197 tokens)",
                      workload="BASELINE config 2: ImageViT d6 h8 (post-norm, GELU, dropout 0.1) from scratch on "
                               "224x224 images, batch 64 per B200"),
    "latent_vit_v2": dict(index=4, batch=512, flop=2.184e9,
                          metric="train samples/sec (LatentViTv2: LEAM + SemanticPE + LayerWiseNorm, S=19 tokens)",
                          workload="BASELINE config 4: LatentViTv2 with LEAM + semantic_pe + layer_wise_norm (residual "
                                   "gate), batch 512 per B200"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"],
                "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(power),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------------------
# Reference arm: the UNMODIFIED reference classes from baseline/_ref (oracle/make_ref.py) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def _quiet_stdout():
    """The reference's packages print at import / construction time; bench.py's stdout carries ONE JSON line."""
    import contextlib
    return contextlib.redirect_stdout(sys.stderr)


def cpu_reference(config: str, B: int, steps: int, warmup: int):
    """(record, kind): the reference train step on all host cores. kind "reference" = the real classes from
    baseline/_ref; "port" = the oracle restatement (only when baseline/_ref did not travel; hybrid only)."""
    import torch
    cores = os.cpu_count() or 1
    from oracle import ref_runner as RR
    if RR.available():
        with _quiet_stdout():
            r = RR.time_train_steps(config, B, steps, warmup, device="cpu", threads=cores)
        return r, "reference", cores
    if config != "hybrid":
        raise RuntimeError("baseline/_ref is missing and the oracle port times the hybrid configuration only")
    from oracle import baseline_models as BM
    r = BM.time_hybrid_step(B, max(1, steps), max(0, warmup), cores)
    return {"s_per_step": r["median_s"], "samples_per_s": B / r["median_s"], "B": B, "steps": steps,
            "loop": "oracle port: fwd+CE+bwd, no optimizer", "threads": torch.get_num_threads()}, "port", cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cfg = CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r, kind, cores = cpu_reference(args.config, B, steps, warmup)
    value = r["samples_per_s"]
    what = ("the reference's own classes and train loop (" + r["loop"] + ": zero_grad, forward, CrossEntropyLoss, "
            "backward, optim.AdamW.step, loss.item) imported unmodified from baseline/_ref" if kind == "reference" else
            "oracle port of the reference (baseline/_ref missing)")
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"] + " — reference arm: " + what + ", fp32 on the host CPU",
                   "per_gpu_batch": B, "global_batch": B,
                   "sample": f"every step is one full batch of {B} samples (the arm's own config); {steps} timed "
                             f"steps after {warmup} warm-up"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{steps} steps of batch {B}, torch {torch.__version__} CPU, "
                                   f"{r.get('threads')} threads, loop: {r['loop']}"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# Native arm
# ------------------------------------------------------------------------------------------------------------------
def build_model(fv, config: str):
    if config == "hybrid":
        return fv.create_hybrid_latent_vit(latent_dim=512, seq_len=18, model_size="base", num_classes=7,
                                           use_pretrained=False, freeze_transformer=True, use_adapter=True,
                                           adapter_dim=64)
    if config == "latent_vit":
        return fv.LatentViT()
    if config == "latent_vit_v2":
        return fv.LatentViTv2(use_lwn=True, use_lwn_residual=True, use_spe=True, use_leam=True)
    if config == "image_vit":
        return fv.ImageViT(embed_dim=512, depth=6, heads=8, mlp_dim=2048)
    raise ValueError(config)


def input_shape(config: str):
    return (3, 224, 224) if config == "image_vit" else (18, 512)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="hybrid", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the BASELINE config's batch)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-steps", type=int, default=2, help="timed reference steps of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the torch-eager run of the reference classes on "
                                                                "the GPU (N = 1 only)")
    ap.add_argument("--no-config5", action="store_true", help="skip the global-batch-4096 leg (hybrid only)")
    ap.add_argument("--no-dp-check", action="store_true")
    ap.add_argument("--buckets", type=int, default=2)
    ap.add_argument("--profile-only", action="store_true", help="run warm-up + the timed steps and exit (for ncu)")
    ap.add_argument("--torch-optim", action="store_true", help="torch.optim.AdamW(fused=True) instead of FusedAdamW")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying "
                                                            "the captured CUDA graph of the step")
    ap.add_argument("--watchdog", type=int, default=0, help="dump every thread's Python stack to stderr after this "
                                                             "many seconds and exit (diagnosing multi-GPU hangs)")
    args = ap.parse_args()
    if args.impl == "native":
        args.warmup = max(args.warmup, 3)
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import fer_vit_b200 as fv
    from fer_vit_b200 import _lib as L
    from fer_vit_b200 import parallel

    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; fer_vit_b200 has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"bench.py: WORLD_SIZE={world} differs from --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    n_gpus = world

    B = args.batch or cfg["batch"]
    global_batch = B * n_gpus
    shape = input_shape(args.config)
    torch.manual_seed(42)                    # same weights on every rank (then broadcast anyway)
    fv.set_default_precision(args.precision)
    model = build_model(fv, args.config).to(dev).train()
    bucketer = parallel.enable_data_parallel(model, num_buckets=args.buckets) if n_gpus > 1 else None
    params = [p for p in model.parameters() if p.requires_grad]
    n_trainable = sum(p.numel() for p in params)
    if args.torch_optim:
        opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.01, fused=True, capturable=not args.no_graph)
    else:
        opt = fv.FusedAdamW(params, lr=1e-3, weight_decay=0.01)   # the repo's own fused optimizer

    # synthetic inputs: a rotating pool larger than the 126 MB L2; a different seed per rank = a different shard
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    bytes_per_batch = B * math.prod(shape) * 4
    n_pool = max(2, math.ceil(160e6 / bytes_per_batch))
    pool_x = torch.randn(n_pool, B, *shape, device=dev, generator=g)
    pool_y = torch.randint(0, 7, (n_pool, B), device=dev, generator=g)

    def make_eager(m, o):
        def eager_step(x, y):
            o.zero_grad(set_to_none=True)
            loss = fv.cross_entropy(m(x), y)
            loss.backward()
            o.step()
            return loss
        return eager_step

    eager_step = make_eager(model, opt)
    # the public train-step API: the whole step captured once as a CUDA graph, replayed per batch
    graphed = None if args.no_graph else fv.GraphedTrainStep(model, opt, pool_x[0], pool_y[0])
    train_step = eager_step if graphed is None else graphed
    live_graphs = [graphed] if graphed is not None else []

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(steps):
            fn(i)
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if n_gpus > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---------------- device-resident timing (value) ----------------
    for i in range(args.warmup):
        train_step(pool_x[i % n_pool], pool_y[i % n_pool])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.launch_count()
    ms = timed(lambda i: train_step(pool_x[i % n_pool], pool_y[i % n_pool]), args.steps)
    launches = L.launch_count() - launches0
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps
    clocks = sampler.stop() if rank == 0 else {}
    if args.profile_only:
        return
    value = global_batch * args.steps / (ms / 1e3)

    # ---------------- end-to-end timing (e2e): host batches, H2D inside the timed region, loss read back ----------------
    hg = torch.Generator().manual_seed(4321 + rank)           # every rank feeds its own shard from host memory
    host_x = [torch.randn(B, *shape, generator=hg).pin_memory() for _ in range(4)]
    host_y = [torch.randint(0, 7, (B,), generator=hg).pin_memory() for _ in range(4)]
    host_loss = torch.zeros(args.steps + 8).pin_memory()
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty(B, *shape, device=dev), torch.empty(B, dtype=torch.int64, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])
            bufs[k][0].copy_(host_x[i % 4], non_blocking=True)
            bufs[k][1].copy_(host_y[i % 4], non_blocking=True)
            ready[k].record(copy_stream)

    def e2e_step(i):
        k = i % 2
        if i == 0:
            prefetch(0)
        prefetch(i + 1)                       # next batch's H2D overlaps this step's compute
        torch.cuda.current_stream().wait_event(ready[k])
        loss = train_step(bufs[k][0], bufs[k][1])
        consumed[k].record()
        host_loss[i % host_loss.numel()].copy_(loss.detach(), non_blocking=True)   # D2H of the step's result

    for k in range(2):
        consumed[k].record()
    for i in range(3):
        e2e_step(i)
    torch.cuda.synchronize()
    for k in range(2):
        consumed[k].record()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = global_batch * args.steps / (ms_e2e / 1e3)
    h2d = B * math.prod(shape) * 4 + B * 8
    loss_last = float(host_loss[(args.steps - 1) % host_loss.numel()])

    # ---------------- roofline leg ----------------
    # Dominant kernel (the CTA-pair tcgen05 GEMM): timed WHERE IT RUNS, inside a replay of the step's CUDA graph. Events
    # cannot sit inside the graph without cutting its PDL edges and parallel branch, so the kernel stamps itself:
    # every launch records {first CTA start after its grid dependency resolved, last CTA exit} in %globaltimer ns
    # (fervit_gemm_prof; the graph is captured once more with the timer on). Sum of those = the GEMMs' share of the very
    # step `value` measures. The other classes (attention, LayerNorm, single-CTA GEMM, adapter) are timed with CUDA
    # events around every launch of a host-launched pass queued behind a spin kernel (back to back, one stream).
    pk = peaks()
    lib = L.lib()
    roof, hbm, classes = None, {}, {}
    st_ptr = torch.cuda.current_stream().cuda_stream
    if graphed is not None and args.precision == "bf16":
        L.check(lib.fervit_gemm_prof(1, None))
        prof_graph = fv.GraphedTrainStep(model, opt, pool_x[0], pool_y[0])
        L.check(lib.fervit_gemm_prof(0, None))
        live_graphs.append(prof_graph)
        reps, cap = 5, 512
        tot_us = tot_fl = tot_ms = 0.0
        nl = 0
        by_shape = {}
        kinds = {0: "plain", 1: "gelu", 2: "relu", 3: "gelu'", 4: "relu'", 5: "x saved act'"}
        for i in range(reps + 1):
            L.check(lib.fervit_gemm_prof(2, st_ptr))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            prof_graph(pool_x[i % n_pool], pool_y[i % n_pool])
            e1.record()
            torch.cuda.synchronize()
            us, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
            rec = (ctypes.c_double * (8 * cap))()
            L.check(lib.fervit_gemm_prof_read(ctypes.byref(us), ctypes.byref(fl), ctypes.byref(n), rec, cap))
            if i == 0:
                continue
            tot_us += us.value; tot_fl += fl.value; tot_ms += e0.elapsed_time(e1); nl = n.value
            last_rec = [tuple(rec[8 * j:8 * j + 8]) for j in range(min(n.value, cap))]
            for j in range(min(n.value, cap)):
                d, f, M_, N_, K_, kd = rec[8 * j:8 * j + 6]
                key = (int(M_), int(N_), int(K_), int(kd))
                a = by_shape.setdefault(key, [0.0, 0.0, 0])
                a[0] += d; a[1] += f; a[2] += 1
        if nl:
            achieved = tot_fl / (tot_us * 1e-6) / 1e12
            traffic, traffic_note = None, None
            tpath = os.path.join(ROOT, "profiles", "r02_ncu_gemm_traffic.json")
            if not os.path.exists(tpath):
                tpath = os.path.join(ROOT, "profiles", "r01_ncu_gemm_traffic.json")
            if os.path.exists(tpath) and args.config == "hybrid":
                tj = json.load(open(tpath))
                traffic = tj["dram_bytes_per_launch"]
                traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum of one launch ({tj['launch']}) from the "
                                f"committed ncu --set full capture {os.path.basename(tpath)}; algorithmic bytes "
                                f"{tj['algorithmic_bytes_per_launch']}; {tj['note']}")
            shapes = []
            for (M_, N_, K_, kd), (d, f, c) in sorted(by_shape.items(), key=lambda kv: -kv[1][0]):
                shapes.append({"M": M_, "N": N_, "K": K_, "epilogue": kinds.get(kd % 16, str(kd % 16)) +
                               (" + fp32 residual stream" if kd >= 16 else ""), "launches_per_step": c // reps,
                               "avg_us": d / c, "tflops": f / (d * 1e-6) / 1e12,
                               "frac_of_peak": f / (d * 1e-6) / 1e12 / pk["tflops_sustained"]})
            # fc1 -> fc2 (forward) and fc2-dgrad -> fc1-dgrad (backward) are launched back to back with nothing between
            # them: the time from the first one's last CTA exit to the second one's first CTA past its grid dependency
            # is the pure kernel-boundary cost inside the graph (PDL edge)
            seq = sorted(last_rec, key=lambda r: r[6])
            gaps = [b[6] - a[7] for a, b in zip(seq, seq[1:])
                    if int(a[3]) == int(b[4]) and int(a[4]) == int(b[3]) and int(a[3]) > int(a[4])
                    and int(a[5]) % 16 in (1, 2, 5)]
            roof = {"bound": "tensor", "kernel": "tc2::gemm_tc2_kernel (CTA-pair tcgen05.mma cta_group::2 kind::f16, "
                                                 "TMA-fed, TMEM accumulators, TMA-store epilogue: every forward / dgrad GEMM)",
                    "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / pk["tflops_sustained"], "traffic": traffic, "traffic_note": traffic_note,
                    "peak_source": pk["source"] + " (sustained cuBLAS bf16: kernel timed inside a long step)",
                    "launches_per_step": nl, "flops_per_step": tot_fl / reps,
                    "avg_launch_us": tot_us / reps / nl, "ms_per_step_in_kernel": tot_us / reps / 1e3,
                    "share_of_step": (tot_us / reps / 1e3) / (tot_ms / reps), "step_ms_of_timed_replays": tot_ms / reps,
                    "by_shape": shapes,
                    "back_to_back_gap_us": ({"median": statistics.median(gaps), "min": min(gaps), "max": max(gaps),
                                             "pairs": len(gaps), "what": "last CTA exit of fc1 (fc2-dgrad) -> first CTA of "
                                             "fc2 (fc1-dgrad) past its grid dependency, same replay: the cost of one "
                                             "kernel boundary inside the graph"} if gaps else None),
                    "how": "in-kernel %globaltimer stamps (min start after the grid dependency, max exit over the CTAs "
                           "of each launch) collected from " + str(reps) + " replays of the step's CUDA graph captured "
                           "with the timer on; share_of_step = sum of the launch durations / the replay's own duration "
                           "(CUDA events around the replay). Compare with the kernel's share in the ncu launch list "
                           "under profiles/."}
    psteps = 3
    blocker = int(0.045 * 1.9e9)   # ~45 ms of spin: one host-launched step takes the host ~5-8 ms to enqueue

    def serial_pass(mode):
        lib.fervit_profile_enable(2)
        eager_step(pool_x[0], pool_y[0])      # untimed: allocator, caches
        torch.cuda.synchronize()
        lib.fervit_profile_enable(mode)       # mode 1 clears the records and starts recording
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(blocker)
        e0.record()
        for i in range(psteps):
            eager_step(pool_x[i % n_pool], pool_y[i % n_pool])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / psteps

    serial_ms = serial_pass(2)
    serial_pass(1)

    def read(cls):
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        L.check(lib.fervit_profile_read(cls, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value / psteps, b.value / psteps, c.value // psteps
    names = {0: "tc2::gemm_tc2_kernel", 4: "tc::gemm_tc_kernel (single-CTA tcgen05 GEMM: weight gradients, token "
             "projection)", 5: "adp::adapter_kernel (fused AdapterModule)", 3: "gemm_simt_kernel (fp32 mode)",
             1: "attention", 2: "layernorm"}
    for cls, nm in names.items():
        t_ms, work, n = read(cls)
        if n:
            classes[cls] = {"kernel": nm, "ms_per_step": t_ms, "launches_per_step": n, "work_per_step": work,
                            "share_of_serial_step": t_ms / serial_ms}
    lib.fervit_profile_enable(0)
    if roof is None and (0 in classes or 3 in classes):   # fp32 mode / --no-graph: event-timed fall-back
        c0 = classes[0 if 0 in classes else 3]
        achieved = c0["work_per_step"] / (c0["ms_per_step"] * 1e-3) / 1e12
        roof = {"bound": "tensor" if 0 in classes else "fp32 CUDA cores", "kernel": c0["kernel"], "achieved": achieved,
                "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"],
                "traffic": None, "launches_per_step": c0["launches_per_step"],
                "avg_launch_us": c0["ms_per_step"] * 1e3 / c0["launches_per_step"],
                "ms_per_step_in_kernel": c0["ms_per_step"], "share_of_step": c0["share_of_serial_step"],
                "how": "CUDA events around every launch of a host-launched serial pass"}
    if roof is not None:
        roof["serial_pass"] = {"ms_per_step": serial_ms, "how": "the same kernels launched from the host on one stream "
                               "behind a spin kernel, CUDA events around every launch (serialised, warm caches, "
                               "~2-4 us of event overhead inside every interval): the other kernel classes"}
        for cls in (0, 4, 5):
            if cls in classes:
                c = classes[cls]
                roof.setdefault("serial_pass_tensor_kernels", []).append(
                    {"kernel": c["kernel"], "achieved_tflops": c["work_per_step"] / (c["ms_per_step"] * 1e-3) / 1e12,
                     "ms_per_step": c["ms_per_step"], "launches_per_step": c["launches_per_step"],
                     "share_of_serial_step": c["share_of_serial_step"]})
    for cls, key in ((1, "attention"), (2, "layernorm")):
        if cls in classes:
            c = classes[cls]
            gbs = c["work_per_step"] / (c["ms_per_step"] * 1e-3) / 1e9
            hbm[key] = {"achieved_gbs": gbs, "frac": gbs / pk["hbm_gbs"], "ms_per_step": c["ms_per_step"],
                        "launches_per_step": c["launches_per_step"], "share_of_serial_step": c["share_of_serial_step"]}

    # ---------------- data-parallel correctness on this hardware (N > 1, outside every timed region) ----------------
    dp_check = None
    if n_gpus > 1 and not args.no_dp_check:
        dp_check = run_dp_check(torch, dist, fv, model, opt, dev, shape, rank, n_gpus)

    # ---------------- BASELINE config 5: global batch 4096 split over the GPUs (hybrid only) ----------------
    config5 = None
    if args.config == "hybrid" and not args.no_config5 and not args.batch:
        config5 = run_config5(torch, dist, fv, model, opt, dev, rank, n_gpus, args, make_eager, live_graphs)

    # ---------------- CPU baseline + GPU eager baseline of the reference classes (rank 0, N = 1 only) ----------------
    cpu, gpu_eager = None, None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        try:
            r, kind, cores = cpu_reference(args.config, B, args.cpu_steps, 1)
            cpu = {"value": r["samples_per_s"], "unit": "samples/s", "cores": cores, "kind": kind,
                   "sample": f"{args.cpu_steps} timed steps (after 1 warm-up) of batch {B}: {r['loop']} of the "
                             f"unmodified reference classes, fp32, {r.get('threads')} host threads"}
        except Exception as e:  # the baseline must never cost the run its line
            cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"failed: {type(e).__name__}: {e}"}
    if rank == 0 and n_gpus == 1 and not args.no_gpu_eager:
        gpu_eager = {"what": "the unmodified reference classes (baseline/_ref) run by torch eager on this same B200 "
                             "(cuBLASLt GEMMs, SDPA, native LayerNorm; BASELINE.md section 5), same batch, same loop "
                             "(zero_grad, fwd, CE, bwd, AdamW.step, loss.item)", "unit": "samples/s"}
        try:
            from oracle import ref_runner as RR
            with _quiet_stdout():
                for key, ac in (("fp32", False), ("autocast_bf16", True)):
                    r = RR.time_train_steps(args.config, B, 10, 3, device=f"cuda:{local_rank}", autocast_bf16=ac)
                    gpu_eager[key] = {"value": r["samples_per_s"], "ms_per_step": r["s_per_step"] * 1e3}
            gpu_eager["native_over_autocast_bf16"] = value / gpu_eager["autocast_bf16"]["value"]
        except Exception as e:
            gpu_eager["error"] = f"{type(e).__name__}: {e}"

    if rank == 0:
        step_flops = cfg["flop"] * global_batch
        line = {
            "metric": cfg["metric"], "value": value, "unit": "samples/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {
                "workload": cfg["workload"] + (f"; data parallel over {n_gpus} GPUs, NCCL all-reduce of the "
                                               f"{n_trainable:,} trainable gradients in {args.buckets} buckets overlapped "
                                               "with backward" if n_gpus > 1 else ""),
                "baseline_config": cfg["index"],
                "global_batch": global_batch, "per_gpu_batch": B,
                "seq_len": 197 if args.config == "image_vit" else 19, "parallelism": f"dp{n_gpus}",
                "step": "zero_grad + fwd + CE + bwd + fused AdamW over the trainable set (" +
                        ("torch.optim.AdamW fused" if args.torch_optim else "fer_vit_b200.FusedAdamW") + ")" +
                        ("" if graphed is None else ", replayed from one captured CUDA graph (fer_vit_b200.GraphedTrainStep)"),
                "l2": f"inputs rotate over a {n_pool * bytes_per_batch / 1e6:.0f} MB pool (> 126 MB L2)",
            },
            "model_tflops": step_flops * args.steps / (ms / 1e3) / 1e12 / n_gpus,
            "model_frac_of_peak": step_flops * args.steps / (ms / 1e3) / 1e12 / n_gpus / pk["tflops_sustained"],
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "last_loss": loss_last},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "hbm_kernels": hbm,
            "cpu_baseline": cpu,
            "gpu_eager_reference": gpu_eager,
            "config5": config5,
            "dp_check": dp_check,
        }
        if bucketer is not None:
            line["allreduce"] = {"bytes_per_step": 4 * n_trainable, "calls_per_step": len(bucketer.stage_groups(
                model.plan_runner().nstages))}
        print(json.dumps(line), flush=True)

    # ---------------- teardown ----------------
    for gr in live_graphs:
        gr.close()          # a live graph holding captured collectives blocks ncclCommDestroy
    if n_gpus > 1:
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        done = threading.Event()

        def teardown():
            dist.destroy_process_group()
            done.set()
        t = threading.Thread(target=teardown, daemon=True)
        t.start()
        if not done.wait(timeout=60):
            print("bench.py: destroy_process_group() did not return within 60 s; leaving without it", file=sys.stderr)
            sys.stderr.flush()
            os._exit(0)


def run_dp_check(torch, dist, fv, model, opt, dev, shape, rank, world):
    """Data parallelism on THIS hardware: (1) the averaged gradient every rank holds after one DP backward equals the
    gradient of the concatenated global batch computed on one GPU; (2) after the optimizer steps of the timed regions
    (different data on every rank) all replicas still hold bit-identical parameters."""
    runner = model.plan_runner()
    # Per-rank batch of the check: both sides must take the same kernels, or the comparison measures two roundings
    # instead of the reduction. The bf16 plan switches the weight / bias gradients to the CTA-pair kernel (bias gradient
    # from the bf16 copy of dY) at 2048 token rows, so the 19-token models are checked at 128 samples per rank (2432 rows
    # on a rank, more on the single-GPU side); the hybrid at 64 (grouped adapter gradients on both sides) and ImageViT at
    # 64 (12608 rows) already are on one side of it.
    Bc = 128 if (len(shape) == 2 and not hasattr(model, "adapters")) else 64
    g = torch.Generator(device=dev).manual_seed(777 + rank)
    x = torch.randn(Bc, *shape, device=dev, generator=g)
    y = torch.randint(0, 7, (Bc,), device=dev, generator=g)
    xs = [torch.empty_like(x) for _ in range(world)]
    ys = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(xs, x)
    dist.all_gather(ys, y)
    was_training = model.training
    model.eval()       # no dropout draws: the two gradients must be the same function of the same data
    named = [(k, p) for k, p in model.named_parameters() if p.requires_grad]

    def grads():
        return torch.cat([p.grad.detach().double().reshape(-1) for _, p in named])
    opt.zero_grad(set_to_none=True)
    fv.cross_entropy(model(x), y).backward()                      # bucketed all-reduce (AVG) inside
    g_dp = grads()
    sync = runner.grad_sync
    runner.grad_sync = None
    opt.zero_grad(set_to_none=True)
    fv.cross_entropy(model(torch.cat(xs)), torch.cat(ys)).backward()   # the whole global batch on this one GPU
    g_one = grads()
    runner.grad_sync = sync
    opt.zero_grad(set_to_none=True)
    err = ((g_dp - g_one).norm() / g_one.norm()).reshape(1).float()
    worst = torch.zeros(1, device=dev)
    off = 0
    for _, p in named:
        n = p.numel()
        d = (g_dp[off:off + n] - g_one[off:off + n]).norm() / g_one[off:off + n].norm().clamp_min(1e-30)
        worst = torch.maximum(worst, d.reshape(1).float())
        off += n
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    # replicas: a 64-bit checksum of the raw bits of every trainable parameter, compared across ranks
    bits = torch.stack([p.detach().view(torch.int32).long().sum() for _, p in named])
    all_bits = [torch.empty_like(bits) for _ in range(world)]
    dist.all_gather(all_bits, bits)
    identical = all(bool(torch.equal(all_bits[0], b)) for b in all_bits[1:])
    if was_training:
        model.train()
    tol, tol_tensor = 1e-4, 1e-3   # whole gradient / any single tensor (small bias tensors carry the rounding noise)
    return {"dp_grad_vs_single_gpu_global_batch_relerr": float(err.item()),
            "worst_tensor_relerr": float(worst.item()), "tolerance": tol, "worst_tensor_tolerance": tol_tensor,
            "grad_ok": bool(err.item() < tol and worst.item() < tol_tensor),
            "replicas_bit_identical_after_training_steps": identical,
            "global_batch": Bc * world, "tensors": len(named),
            "note": "whole-gradient and worst-tensor relative error (max over ranks) between the NCCL-averaged "
                    "gradient of a sharded batch and the gradient of the same global batch computed on one GPU, "
                    "in the model's precision mode (the bf16 forward/dgrad are row-independent; only the fp32 "
                    "reduction order of the weight gradients differs), at a per-rank batch that keeps both sides on the "
                    "same kernels; parameter bit-checksums all-gathered after the timed AdamW steps"}


def run_config5(torch, dist, fv, model, opt, dev, rank, world, args, make_eager, live_graphs):
    """BASELINE config 5: global batch 4096 over the N GPUs (strong scaling) + its 1-GPU denominator on rank 0."""
    GB = 4096
    Bs = GB // world
    runner = model.plan_runner()

    def time_steps(B, use_dp, ks=10):
        sync = runner.grad_sync
        if not use_dp:
            runner.grad_sync = None
        g = torch.Generator(device=dev).manual_seed(99 + rank)
        gx = torch.randn(2, B, 18, 512, device=dev, generator=g)
        gy = torch.randint(0, 7, (2, B), device=dev, generator=g)
        step = make_eager(model, opt) if args.no_graph else fv.GraphedTrainStep(model, opt, gx[0], gy[0])
        for i in range(3):
            step(gx[i % 2], gy[i % 2])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if use_dp and world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for i in range(ks):
            step(gx[i % 2], gy[i % 2])
        e1.record()
        if use_dp and world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / ks], device=dev)
        if use_dp and world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if not args.no_graph:
            step.close()
        runner.grad_sync = sync
        del gx, gy, step
        torch.cuda.empty_cache()
        return float(ms.item())

    out = {"workload": f"BASELINE config 5: same model, global batch {GB} ({Bs}/GPU), data parallel over {world} GPUs",
           "scaling": "strong", "global_batch": GB, "per_gpu_batch": Bs, "unit": "samples/s"}
    ms_n = time_steps(Bs, True)
    out.update(value=GB / (ms_n * 1e-3), ms_per_step=ms_n)
    if world > 1:
        # the like-for-like denominator: the whole global batch on ONE GPU of this box (rank 0; the others wait)
        ref = torch.zeros(1, device=dev)
        if rank == 0:
            ref[0] = time_steps(GB, False)
        dist.broadcast(ref, 0)
        ms_1 = float(ref.item())
        out.update(one_gpu_b4096_value=GB / (ms_1 * 1e-3), one_gpu_b4096_ms_per_step=ms_1,
                   strong_efficiency_vs_1gpu_b4096=(GB / (ms_n * 1e-3)) / (world * GB / (ms_1 * 1e-3)))
    return out


if __name__ == "__main__":
    main()
