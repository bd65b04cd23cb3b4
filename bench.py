#!/usr/bin/env python
"""bench.py — train samples/sec of the HybridLatentViT (frozen ViT-B/16 + Adapter64) bf16 train step on B200.

    python bench.py --gpus 1 --steps 20 --warmup 5                      # BASELINE config 3: batch 256 on 1 B200
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                          # BASELINE config 5: global batch 4096
    python bench.py --impl reference --steps 3 --warmup 1               # the reference algorithm on host cores

One step = zero_grad + forward + cross-entropy + backward (+ NCCL all-reduce of the trainable gradients when N > 1)
+ AdamW on the trainable parameters, on synthetic w+ latents (randn, 18x512) and random-init weights of the named
architecture. Prints ONE JSON line (see DESIGN.md "Measurement" for every field).
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
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train samples/sec (HybridLatentViT: frozen ViT-B/16 + Adapter64, S=19 tokens, bf16)"
FLOP_PER_SAMPLE = 6.658e9        # fwd 3.300 + bwd 3.358 GFLOP (frozen backbone: dgrad only)
GEMM_FLOP_PER_SAMPLE = 6.48e9    # big GEMMs only (K1/K4/K6/K7 and their dgrads)


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


def reference_arm(args):
    """The reference algorithm (oracle port: the reference itself is not installable) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import baseline_models as BM
    cores = os.cpu_count() or 1
    B = args.cpu_batch
    r = BM.time_hybrid_step(B, max(1, args.steps), max(0, args.warmup), cores)
    # mean over the timed steps is what "K steps" means; median reported beside it
    value = B / r["median_s"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["median_s"] * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "HybridLatentViT vit_base_patch16_224 frozen + Adapter(64), w+ tokens 18x512, "
                               "fwd+CE+bwd, fp32 on host CPU",
                   "sample": f"batch {B} per step (bounded sample of the batch-256 workload)"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of batch {B}, torch {torch.__version__} CPU, "
                                   f"{torch.get_num_threads()} threads"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 256 at N=1, 4096/N otherwise)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--buckets", type=int, default=2)
    ap.add_argument("--profile-only", action="store_true", help="run warm-up + the timed steps and exit (for ncu)")
    ap.add_argument("--torch-optim", action="store_true", help="torch.optim.AdamW(fused=True) instead of FusedAdamW")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying "
                                                            "the captured CUDA graph of the step")
    ap.add_argument("--no-scaling-ref", action="store_true", help="skip the batch-4096 single-GPU reference run at N = 1")
    ap.add_argument("--watchdog", type=int, default=0, help="dump every thread's Python stack to stderr after this "
                                                             "many seconds and exit (diagnosing multi-GPU hangs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
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

    B = args.batch or (256 if n_gpus == 1 else 4096 // n_gpus)
    global_batch = B * n_gpus
    torch.manual_seed(42)                    # same weights on every rank (then broadcast anyway)
    fv.set_default_precision(args.precision)
    model = fv.create_hybrid_latent_vit(latent_dim=512, seq_len=18, model_size="base", num_classes=7,
                                        use_pretrained=False, freeze_transformer=True, use_adapter=True,
                                        adapter_dim=64)
    model = model.to(dev).train()
    bucketer = parallel.enable_data_parallel(model, num_buckets=args.buckets) if n_gpus > 1 else None
    params = [p for p in model.parameters() if p.requires_grad]
    if args.torch_optim:
        opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.01, fused=True, capturable=not args.no_graph)
    else:
        opt = fv.FusedAdamW(params, lr=1e-3, weight_decay=0.01)   # the repo's own fused optimizer

    # synthetic inputs: a rotating pool larger than the 126 MB L2 (different seed per rank = different shard)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    bytes_per_batch = B * 18 * 512 * 4
    n_pool = max(2, math.ceil(160e6 / bytes_per_batch))
    pool_x = torch.randn(n_pool, B, 18, 512, device=dev, generator=g)
    pool_y = torch.randint(0, 7, (n_pool, B), device=dev, generator=g)

    def eager_step(x, y):
        opt.zero_grad(set_to_none=True)
        logits = model(x)
        loss = fv.cross_entropy(logits, y)
        loss.backward()
        opt.step()
        return loss

    # the public train-step API: the whole step captured once as a CUDA graph, replayed per batch
    graphed = None if args.no_graph else fv.GraphedTrainStep(model, opt, pool_x[0], pool_y[0])
    train_step = eager_step if graphed is None else graphed

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
    host_x = [torch.randn(B, 18, 512).pin_memory() for _ in range(4)]
    host_y = [torch.randint(0, 7, (B,)).pin_memory() for _ in range(4)]
    host_loss = torch.zeros(args.steps + 8).pin_memory()
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty(B, 18, 512, device=dev), torch.empty(B, dtype=torch.int64, device=dev)) for _ in range(2)]
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
    h2d = B * 18 * 512 * 4 + B * 8
    loss_last = float(host_loss[(args.steps - 1) % host_loss.numel()])

    # ---------------- roofline leg: dominant kernel (tcgen05 GEMM) timed with CUDA events per launch ----------------
    pk = peaks()
    lib = L.lib()
    roof = None
    hbm = {}
    # every rank runs these steps (they contain the gradient all-reduce); only rank 0 records and reports
    torch.cuda.synchronize()
    psteps = 3
    # the serial host-launched step the per-launch events belong to (the timed region above replays a graph, where
    # PDL and the side stream overlap kernels): timed first without the per-kernel events, which slow the host down
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eager_step(pool_x[0], pool_y[0])
    pe0.record()
    for i in range(psteps):
        eager_step(pool_x[i % n_pool], pool_y[i % n_pool])
    pe1.record()
    torch.cuda.synchronize()
    profile_step_ms = pe0.elapsed_time(pe1) / psteps
    if rank == 0:
        lib.fervit_profile_enable(1)
    for i in range(psteps):
        eager_step(pool_x[i % n_pool], pool_y[i % n_pool])   # per-kernel events need host launches
    torch.cuda.synchronize()
    if rank == 0:

        def read(cls):
            a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
            L.check(lib.fervit_profile_read(cls, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
            return a.value, b.value, c.value
        g_ms, g_flops, g_n = read(0)
        a_ms, a_bytes, a_n = read(1)
        l_ms, l_bytes, l_n = read(2)
        lib.fervit_profile_enable(0)
        if g_n:
            achieved = g_flops / (g_ms * 1e-3) / 1e12
            traffic, traffic_note, ncu_share = None, None, None
            tpath = os.path.join(ROOT, "profiles", "r01_ncu_gemm_traffic.json")
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                traffic = tj["dram_bytes_per_launch"]
                ncu_share = tj.get("ncu_share_of_step")
                traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum of one launch ({tj['launch']}), "
                                f"algorithmic bytes {tj['algorithmic_bytes_per_launch']}; {tj['note']}")
            roof = {"bound": "tensor", "kernel": "tc2::gemm_tc2_kernel (CTA-pair tcgen05.mma cta_group::2 kind::f16, "
                                                 "TMA-fed, TMEM accumulators, TMA-store epilogue) + tc::gemm_tc_kernel "
                                                 "(single-CTA, MN-major wgrad)",
                    "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / pk["tflops_sustained"], "traffic": traffic, "traffic_note": traffic_note,
                    "peak_source": pk["source"] +
                    " (sustained cuBLAS bf16: kernel timed inside a long step)",
                    "launches_per_step": g_n // psteps, "flops_per_step": g_flops / psteps,
                    "ms_per_step_in_kernel": g_ms / psteps,
                    # share of the serial host-launched step (compare with the ncu launch list in profiles/)
                    # (the event intervals of a host-launched pass include the host's launch latency whenever the GPU
                    # runs dry, ~5 us per launch here, so this live share reads high; the committed ncu list gives
                    # share_of_step_ncu)
                    "share_of_step": (g_ms / psteps) / profile_step_ms,
                    "share_of_step_ncu": ncu_share,
                    "eager_ms_per_step": profile_step_ms}
        if a_n:
            hbm["attention"] = {"achieved_gbs": a_bytes / (a_ms * 1e-3) / 1e9, "frac": a_bytes / (a_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                                "ms_per_step": a_ms / psteps, "launches_per_step": a_n // psteps}
        if l_n:
            hbm["layernorm"] = {"achieved_gbs": l_bytes / (l_ms * 1e-3) / 1e9, "frac": l_bytes / (l_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                                "ms_per_step": l_ms / psteps, "launches_per_step": l_n // psteps}

    # ---------------- strong-scaling denominator (N = 1 only): config 5's global batch 4096 on this one GPU ----------------
    # The N > 1 runs of this script are BASELINE config 5 (global batch 4096 split over the GPUs: strong scaling); the
    # N = 1 default is config 3 (batch 256). The like-for-like denominator of the 2/4/8-GPU values is therefore this
    # number, not `value`.
    scaling_ref = None
    if rank == 0 and n_gpus == 1 and not args.batch and not args.no_scaling_ref:
        Bs = 4096
        gx = torch.randn(2, Bs, 18, 512, device=dev)
        gy = torch.randint(0, 7, (2, Bs), device=dev)
        big = eager_step if args.no_graph else fv.GraphedTrainStep(model, opt, gx[0], gy[0])
        for i in range(3):
            big(gx[i % 2], gy[i % 2])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        ks = 10
        for i in range(ks):
            big(gx[i % 2], gy[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms_s = e0.elapsed_time(e1) / ks
        scaling_ref = {"global_batch": Bs, "n_gpus": 1, "value": Bs / (ms_s * 1e-3), "unit": "samples/s",
                       "ms_per_step": ms_s, "steps": ks,
                       "note": "BASELINE config 5 on one GPU: the denominator for the strong-scaling efficiency of the "
                               "2/4/8-GPU runs of this script (their global batch is 4096 too)"}
        del big, gx, gy

    # ---------------- CPU baseline (oracle port on this box's host cores; rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        from oracle import baseline_models as BM
        cores = os.cpu_count() or 1
        r = BM.time_hybrid_step(args.cpu_batch, 3, 1, cores)
        cpu = {"value": args.cpu_batch / r["median_s"], "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"3 steps of batch {args.cpu_batch} (median), fp32 fwd+CE+bwd of the same model on the host CPU"}

    if rank == 0:
        step_flops = FLOP_PER_SAMPLE * global_batch
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if n_gpus == 1 else "strong", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {
                "workload": ("BASELINE config 3: HybridLatentViT frozen timm-style ViT-B/16 + Adapter(64) on w+ tokens "
                             "18x512, batch 256 on 1 B200" if n_gpus == 1 else
                             f"BASELINE config 5: same model, data parallel, global batch {global_batch} "
                             f"({B}/GPU), NCCL all-reduce of 1,605,907 trainable grads in {args.buckets} buckets "
                             "overlapped with backward"),
                "global_batch": global_batch, "per_gpu_batch": B, "seq_len": 19, "parallelism": f"dp{n_gpus}",
                "step": "zero_grad + fwd + CE + bwd + fused AdamW over the trainable set (" +
                        ("torch.optim.AdamW fused" if args.torch_optim else "fer_vit_b200.FusedAdamW") + ")" +
                        ("" if graphed is None else ", replayed from one captured CUDA graph (fer_vit_b200.GraphedTrainStep)"),
                "l2": f"inputs rotate over a {n_pool * bytes_per_batch / 1e6:.0f} MB pool (> 126 MB L2); the step's own "
                      "activation working set is > 1.5 GB",
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
            "strong_scaling_ref_1gpu": scaling_ref,
        }
        if bucketer is not None:
            line["allreduce"] = {"bytes_per_step": bucketer.bytes_reduced // max(1, bucketer.calls) * args.buckets,
                                 "calls_per_step": args.buckets}
        print(json.dumps(line), flush=True)
    if n_gpus > 1:
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        if graphed is not None:
            # ncclCommDestroy blocks while a live CUDA graph still holds captured collectives (observed on 2 x B200:
            # both ranks parked in destroy_process_group); the benchmark has printed its line, so leave without the
            # communicator teardown
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
