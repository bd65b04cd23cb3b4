"""Would two half-batch chains run side by side fill the SMs the batch-256 step leaves idle?

Two independent HybridLatentViT replicas (config 3 model) at batch 128 each, their graphed train steps replayed
concurrently on two streams, against one replica at batch 256: aggregate samples/s. (A probe for DESIGN.md; nothing
in the product path uses it.)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fer_vit_b200 as fv  # noqa: E402


def make(B):
    m = fv.create_hybrid_latent_vit(model_size="base", use_pretrained=False, freeze_transformer=True, use_adapter=True,
                                    adapter_dim=64).cuda().train()
    o = fv.FusedAdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.01)
    x = torch.randn(B, 18, 512, device="cuda")
    y = torch.randint(0, 7, (B,), device="cuda")
    return fv.GraphedTrainStep(m, o, x, y), x, y


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    torch.manual_seed(0)
    one, x, y = make(256)
    ms_one = timed(lambda: one.graph.replay())
    halves = [make(128) for _ in range(2)]
    ms_half = timed(lambda: halves[0][0].graph.replay())
    streams = [torch.cuda.Stream() for _ in range(2)]
    main_s = torch.cuda.current_stream()

    def both():
        for (st, _, _), s in zip(halves, streams):
            s.wait_stream(main_s)
            with torch.cuda.stream(s):
                st.graph.replay()
        for s in streams:
            main_s.wait_stream(s)
    ms_both = timed(both)
    print(json.dumps({"batch256_ms": round(ms_one, 4), "batch128_ms": round(ms_half, 4),
                      "two_x_batch128_concurrent_ms": round(ms_both, 4),
                      "samples_per_s_one": round(256 / ms_one * 1e3), "samples_per_s_two_halves": round(256 / ms_both * 1e3)}))


if __name__ == "__main__":
    main()
