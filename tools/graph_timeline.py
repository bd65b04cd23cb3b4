"""Per-kernel timeline of a graphed train step (any BASELINE config), taken from CUPTI activity records.

The step's CUDA graph is replayed under torch.profiler (kernel activities only); for every kernel name the tool
reports launches per step, the mean CUPTI span and the mean time the kernel adds to the step's dependency chain. A diagnostic: it says where the non-GEMM half of the step goes
and what each kernel boundary costs inside the graph. Numbers under a profiler are not bench values.

    python tools/graph_timeline.py [--batch 256] [--replays 5] --out profiles/rNN_graph_timeline.json
"""
import argparse
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fer_vit_b200 as fv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="hybrid", choices=["hybrid", "latent_vit", "image_vit", "latent_vit_v2"])
    ap.add_argument("--batch", type=int, default=0, help="default: the BASELINE config's batch")
    ap.add_argument("--freeze-adapters", action="store_true", help="hybrid: no adapter gradients (what do they cost?)")
    ap.add_argument("--replays", type=int, default=5)
    ap.add_argument("--out", default=None, help="write the JSON here instead of stdout")
    ap.add_argument("--dump", default=None, help="also write the raw (name, start_us, end_us) kernel records here")
    args = ap.parse_args()
    import bench
    torch.manual_seed(0)
    B = args.batch or bench.CONFIGS[args.config]["batch"]
    m = bench.build_model(fv, args.config).cuda().train()
    if args.freeze_adapters:
        for k, p in m.named_parameters():
            if "adapter" in k:
                p.requires_grad_(False)
    o = fv.FusedAdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.01)
    x = torch.randn(B, *bench.input_shape(args.config), device="cuda")
    y = torch.randint(0, 7, (B,), device="cuda")
    step = fv.GraphedTrainStep(m, o, x, y)
    for _ in range(5):
        step.graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_plain = e0.elapsed_time(e1) / 20

    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.replays):
            step.graph.replay()
        torch.cuda.synchronize()
    path = "/tmp/fervit_graph_timeline_trace.json"
    prof.export_chrome_trace(path)
    Ev = collections.namedtuple("Ev", "name start end")
    evs = [Ev(e["name"], float(e["ts"]), float(e["ts"]) + float(e["dur"]))
           for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    evs.sort(key=lambda e: e.start)
    if args.dump:
        t0 = evs[0].start
        json.dump([[e.name.split("(")[0][:90], round(e.start - t0, 3), round(e.end - t0, 3)] for e in evs],
                  open(args.dump, "w"))
    n = len(evs) // args.replays
    # one replay = n kernels; drop the first replay (profiler warm-up), keep whole replays
    evs = evs[n:n * args.replays]
    reps = args.replays - 1
    # With programmatic dependent launch a kernel's CTAs start (and wait on the grid dependency) while its
    # predecessor is still draining, so CUPTI durations overlap. The step is one dependency chain, so the time a
    # kernel ADDS to the step is end_i - max(end_{i-1}, start_i) with the kernels ordered by end time ("exclusive").
    evs.sort(key=lambda e: e.end)
    stat = collections.OrderedDict()
    t_first, t_last = min(e.start for e in evs), evs[-1].end
    idle = 0.0
    for i, e in enumerate(evs):
        name = e.name.split("(")[0][:90]
        s = stat.setdefault(name, [0, 0.0, 0.0])
        s[0] += 1
        s[1] += e.end - e.start
        if i % n:
            prev = evs[i - 1].end
            s[2] += e.end - max(prev, e.start)
            idle += max(e.start - prev, 0.0)
        else:
            s[2] += e.end - e.start
    rows = [{"kernel": k, "launches_per_step": v[0] / reps, "avg_span_us": v[1] / v[0],
             "avg_exclusive_us": v[2] / v[0], "exclusive_us_per_step": v[2] / reps} for k, v in stat.items()]
    rows.sort(key=lambda r: -r["exclusive_us_per_step"])
    text = (json.dumps({"config": args.config, "batch": B, "ms_per_step_no_profiler": ms_plain,
                      "ms_per_step_under_profiler": (t_last - t_first) / reps / 1e3, "kernels_per_step": n,
                      "exclusive_us_per_step": sum(r["exclusive_us_per_step"] for r in rows), "idle_us_per_step": idle / reps,
                      "note": "span = CUPTI start..end (includes the wait on the grid dependency); exclusive = time the "
                              "kernel adds to the chain; idle = no kernel resident",
                      "kernels": rows}, indent=1))
    if args.out:
        open(args.out, "w").write(text + "\n")
    else:
        print(text)


if __name__ == "__main__":
    main()
