"""Data-parallel training of the drop-in models: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

The reference is single-process; data parallelism is the one place this path shards (samples are
independent in forward and backward). Only TRAINABLE-parameter gradients are communicated — 1,605,907 fp32 values
(6.4 MB) for ViT-B + Adapter(64); the frozen backbone (85 M parameters) is replicated and never sent.

The native backward writes all gradients of a step into ONE flat fp32 buffer laid out in backward order
(head | block 11 ... block 0 | input stage), so a bucket is a contiguous slice. The backward is cut into
`num_buckets` stage groups; the all-reduce of a finished bucket is launched asynchronously (NCCL's own stream) while
the next group of blocks is still computing, and the compute stream waits for the buckets only after the last
group. At this payload the collective is latency-bound (tens of microseconds); the cuts are made from the end so
that only the input stage's 1.6 MB is exposed after the backward pass (`GradBucketer.stage_groups`). Measured at 8
GPUs (profiles/r02_buckets_8gpu.jsonl): 1 / 2 / 3 / 4 buckets = 3.443 / 3.434 / 3.460 / 3.499 ms per step at 256 samples
per GPU — every extra collective costs more (its launch latency and the SMs NCCL takes from the backward GEMMs) than
its overlap returns, so 2 is the default; models whose whole parameter set trains (77 MB of gradients) may want more.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


class GradBucketer:
    """Overlapped, bucketed gradient averaging for runtime.PlanRunner.backward()."""

    def __init__(self, process_group: Optional[dist.ProcessGroup] = None, num_buckets: int = 2):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.num_buckets = max(1, int(num_buckets))
        self._pending: List[Tuple[object, torch.Tensor]] = []
        backend = dist.get_backend(process_group)
        self._native_avg = backend == "nccl"
        self.bytes_reduced = 0
        self.calls = 0

    def stage_groups(self, nstages: int) -> List[Tuple[int, int]]:
        """Contiguous stage ranges [b, e), one per bucket; the head and the latest blocks come first.

        The all-reduce of a bucket overlaps only the stages that run AFTER it, so the cuts are made from the end: the
        last bucket is the input stage alone (its collective is fully exposed whatever it holds: keep it to the token
        projection's 1.6 MB), the bucket before it the two earliest blocks (its collective hides behind the input
        stage), and the remaining buckets split the head and the later blocks evenly."""
        n = min(self.num_buckets, nstages)
        if n <= 1:
            return [(0, nstages)]
        if n == 2:
            cuts = [0, nstages - 1, nstages]
        else:
            tail = max(1, nstages - 3)
            head = [round(i * tail / (n - 2)) for i in range(n - 1)]
            cuts = head + [nstages - 1, nstages]
        return [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]

    def reduce_async(self, flat: torch.Tensor) -> None:
        op = dist.ReduceOp.AVG if self._native_avg else dist.ReduceOp.SUM
        work = dist.all_reduce(flat, op=op, group=self.group, async_op=True)
        self._pending.append((work, flat))
        self.bytes_reduced += flat.numel() * flat.element_size()
        self.calls += 1

    def finish(self) -> None:
        for work, flat in self._pending:
            work.wait()  # NCCL: the current stream waits on the collective; the host does not block
            if not self._native_avg:
                flat.div_(self.world)
        self._pending.clear()


def sync_parameters(model: torch.nn.Module, src: int = 0, process_group=None) -> None:
    """Broadcast every parameter and buffer from `src` once, so replicas start identical (frozen weights included;
    afterwards they never travel again)."""
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t, src=src, group=process_group)


def enable_data_parallel(model, process_group=None, num_buckets: int = 2, broadcast: bool = True) -> GradBucketer:
    """Attach overlapped gradient averaging to a fer_vit_b200 model (any NativeModule)."""
    if broadcast:
        sync_parameters(model, 0, process_group)
    bucketer = GradBucketer(process_group, num_buckets)
    model.plan_runner().grad_sync = bucketer
    return bucketer


def global_ce_denominator(labels: torch.Tensor, weight: Optional[torch.Tensor], process_group=None) -> torch.Tensor:
    """sum_i w[y_i] over the GLOBAL batch: pass it as `den` to fer_vit_b200.cross_entropy so a
    class-weighted mean loss has the single-GPU value; the all-reduce then AVERAGES gradients, so the local
    denominator is the global one divided by the world size."""
    if weight is None:
        local = torch.tensor(float(labels.numel()), device=labels.device)
    else:
        local = weight.to(labels.device)[labels].sum().float()
    total = local.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=process_group)
    return (total / dist.get_world_size(process_group)).reshape(1)
