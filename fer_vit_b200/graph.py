"""CUDA-graph replay of a whole train step.

The reference's step (train_hybrid_latent_vit.py:127-142: zero_grad -> forward -> loss -> backward -> optimizer.step)
is ~360 kernel launches of 5-40 us each at batch 256. Replaying them from one captured CUDA graph removes the host
launch path and the inter-kernel gaps it leaves at small per-GPU batches (the 8-GPU end of the data-parallel
configuration), with no tracing compiler involved: the graph holds exactly the kernels the eager step launches.

    step = GraphedTrainStep(model, optimizer, example_x, example_y)      # optimizer: capturable=True
    loss = step(x, y)        # x, y device tensors of the example's shape; returns the (static) loss tensor

Works with parallel.enable_data_parallel(): the bucketed NCCL all-reduces are captured in place, still overlapped
with the backward stages that follow them.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import runtime


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, example_x: torch.Tensor,
                 example_y: torch.Tensor, loss_fn: Optional[Callable] = None, warmup: int = 3):
        if not example_x.is_cuda:
            raise RuntimeError("fer_vit_b200: GraphedTrainStep needs CUDA tensors (there is no CPU fallback)")
        for grp in optimizer.param_groups:
            if not grp.get("capturable", False):
                raise RuntimeError("fer_vit_b200: GraphedTrainStep needs an optimizer built with capturable=True "
                                   "(e.g. torch.optim.AdamW(params, fused=True, capturable=True))")
        self.model = model
        self.optimizer = optimizer
        self.loss_fn = loss_fn or runtime.cross_entropy
        self.static_x = example_x.detach().clone()
        self.static_y = example_y.detach().clone()
        runner = model.plan_runner()
        if runner.seed_dev is None:
            # dropout masks are a pure function of (seed, site, element); the captured host seed is fixed, this device
            # counter (incremented inside the graph) makes every replay draw fresh masks
            runner.seed_dev = torch.zeros(1, dtype=torch.int64, device=example_x.device)
        self._seed_dev = runner.seed_dev

        side = torch.cuda.Stream(device=example_x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):      # allocator, cuBLAS-free: warms the caching allocator and NCCL
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        launches0 = _lib.launch_count()
        mode = "thread_local" if torch.distributed.is_available() and torch.distributed.is_initialized() else "global"
        with torch.cuda.graph(self.graph, capture_error_mode=mode):
            self.static_loss = self._eager()
        self.launches_per_replay = _lib.launch_count() - launches0   # native kernels inside one replay
        self.replays = 0

    def _eager(self) -> torch.Tensor:
        self._seed_dev.add_(1)
        self.optimizer.zero_grad(set_to_none=True)
        logits = self.model(self.static_x)
        loss = self.loss_fn(logits, self.static_y)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if x.shape != self.static_x.shape or y.shape != self.static_y.shape:
            raise RuntimeError(f"fer_vit_b200: GraphedTrainStep was captured for {tuple(self.static_x.shape)} / "
                               f"{tuple(self.static_y.shape)}, got {tuple(x.shape)} / {tuple(y.shape)}")
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_loss
