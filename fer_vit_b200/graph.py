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
        self._capture(example_x.device, warmup)

    def _capture(self, device, warmup: int) -> None:
        model, optimizer = self.model, self.optimizer
        runner = model.plan_runner()
        if runner.seed_dev is None:
            # dropout masks are a pure function of (seed, site, element); the captured host seed is fixed, this device
            # counter (incremented inside the graph) makes every replay draw fresh masks
            runner.seed_dev = torch.zeros(1, dtype=torch.int64, device=device)
        self._seed_dev = runner.seed_dev

        side = torch.cuda.Stream(device=device)
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

    def close(self) -> None:
        """Release the captured graph. Required before ``torch.distributed.destroy_process_group()`` when the step
        was captured under ``enable_data_parallel``: a live graph keeps the captured NCCL collectives (and with them
        the communicator) alive, and ``ncclCommDestroy`` then blocks."""
        if getattr(self, "graph", None) is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def _check_open(self) -> None:
        if self.graph is None:
            raise RuntimeError("fer_vit_b200: this graphed train step has been closed")

    def _host_launched(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """The same step with every kernel launched from the host (same native kernels, no graph): the path of a
        batch whose shape differs from the captured one, e.g. the last, partial batch of an epoch
        (DataLoader(drop_last=False), train_hybrid_latent_vit.py:212-217)."""
        self._seed_dev.add_(1)
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model(x), y)
        loss.backward()
        self.optimizer.step()
        # the captured graph addresses the gradient tensors of ITS replays: drop the ones just produced so the next
        # replay does not see stale .grad objects of another shape's backward
        self.optimizer.zero_grad(set_to_none=True)
        return loss.detach()

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self._check_open()
        if x.shape[1:] != self.static_x.shape[1:] or x.dtype != self.static_x.dtype or y.dim() != self.static_y.dim():
            raise RuntimeError(f"fer_vit_b200: GraphedTrainStep was captured for inputs {tuple(self.static_x.shape)} / "
                               f"{tuple(self.static_y.shape)}, got {tuple(x.shape)} / {tuple(y.shape)}")
        if x.shape[0] != self.static_x.shape[0]:
            if x.shape[0] != y.shape[0] or x.shape[0] == 0:
                raise RuntimeError("fer_vit_b200: GraphedTrainStep: inputs and labels must hold the same, non-zero "
                                   "number of samples")
            return self._host_launched(x, y)      # another batch size: same kernels, launched from the host
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_loss


class GraphedMixupTrainStep(GraphedTrainStep):
    """The LatentViT trainers' whole step (train_latent_vit.py:115-142) as one graph replay, fed from a
    ``PackedLatentCache``: gather + ``LatentAugment`` + mixup -> zero_grad -> forward -> mixup loss -> backward ->
    optimizer step -> (optionally) the trainer's no-grad forward on the un-mixed batch for train accuracy.

        step = GraphedMixupTrainStep(model, optimizer, cache, batch_size, criterion)     # fer_vit_b200.CrossEntropyLoss
        loss, correct = step(sample_idx, mix_index, lam)     # device index tensors [batch_size], python float lam

    ``loss`` and ``correct`` (number of correct predictions of the accuracy pass, or None) are static device tensors;
    read them off the hot path. The augmentation draws a fresh stream per replay (device-side seed counter); the
    accuracy pass sees the same augmented latents as the training pass, as in the reference (augmentation happens in
    the dataset, before mixup)."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, cache, batch_size: int,
                 criterion=None, train_accuracy: bool = True, seed: int = 0, warmup: int = 3):
        dev = cache.latents.device
        for grp in optimizer.param_groups:
            if not grp.get("capturable", False):
                raise RuntimeError("fer_vit_b200: GraphedMixupTrainStep needs an optimizer built with capturable=True")
        self.model = model
        self.optimizer = optimizer
        self.cache = cache
        self.criterion = criterion or runtime.CrossEntropyLoss()
        self.train_accuracy = train_accuracy
        self.seed = int(seed)
        self.static_idx = torch.arange(batch_size, device=dev) % len(cache)
        self.static_mix = torch.arange(batch_size, device=dev)
        self.static_lam = torch.ones(1, dtype=torch.float32, device=dev)
        self.static_correct = None
        self._capture(dev, warmup)

    def _eager(self) -> torch.Tensor:
        self._seed_dev.add_(1)
        c = self.cache
        mixed, labels = c.batch(self.static_idx, self.static_mix, self.static_lam, self.seed, self._seed_dev)
        self.optimizer.zero_grad(set_to_none=True)
        logits = self.model(mixed)
        loss = self.criterion.mixup(logits, labels, self.static_mix, self.static_lam)
        loss.backward()
        self.optimizer.step()
        if self.train_accuracy:
            with torch.no_grad():
                clean, _ = c.batch(self.static_idx, None, 1.0, self.seed, self._seed_dev)
                self.static_correct = (self.model(clean).argmax(dim=1) == labels).sum()
        return loss.detach()

    def __call__(self, sample_idx: torch.Tensor, mix_index: torch.Tensor, lam: float):
        self._check_open()
        if sample_idx.shape != self.static_idx.shape or mix_index.shape != self.static_mix.shape:
            raise RuntimeError(f"fer_vit_b200: GraphedMixupTrainStep was captured for batches of "
                               f"{self.static_idx.numel()}, got {sample_idx.numel()} / {mix_index.numel()}")
        self.static_idx.copy_(sample_idx, non_blocking=True)
        self.static_mix.copy_(mix_index, non_blocking=True)
        self.static_lam.fill_(float(lam))
        self.graph.replay()
        self.replays += 1
        return self.static_loss, self.static_correct
