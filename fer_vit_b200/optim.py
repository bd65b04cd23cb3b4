"""Fused AdamW (+ optional global-norm clipping) over the trainable parameters, on the native kernels.

Drop-in for ``torch.optim.AdamW`` as the reference trainers build it — one group
(``train_hybrid_latent_vit.py:244-248``) or the five layer-wise learning-rate groups of
``train_hybrid_latent_vit.py:63-117`` — with ``clip_grad_norm_`` (``train_latent_vit_v2.py:132-133``) folded in:

    opt = fer_vit_b200.FusedAdamW(param_groups, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0)
    loss.backward(); opt.step()

Same update rule as torch (decoupled weight decay, bias correction, amsgrad off); state keys ``exp_avg`` /
``exp_avg_sq`` / ``step`` as torch's, so ``state_dict()`` round-trips. The step counter is a device tensor and the
tensors travel to the kernels by value, so ``step()`` is CUDA-graph capturable (``GraphedTrainStep`` accepts it).
Learning-rate schedulers work: group hyper-parameters are re-uploaded when the host values change (outside a graph
capture; inside a captured graph they are whatever the device table holds at replay time — update the table with
``refresh_hyper()`` between replays).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("FusedAdamW: invalid hyper-parameter")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=True)
        super().__init__(params, defaults)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self._hyper_host = None
        self._hyper_dev: Optional[torch.Tensor] = None
        self._step_dev: Optional[torch.Tensor] = None
        self._scratch: Optional[torch.Tensor] = None
        self.last_total_norm: Optional[torch.Tensor] = None   # view of the scratch buffer (valid with clipping)
        self._stepped = False

    # ------------------------------------------------------------------ device tables
    def _hyper_rows(self):
        return [(float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                 float(g["weight_decay"])) for g in self.param_groups]

    def refresh_hyper(self) -> None:
        """Upload the groups' current lr / betas / eps / weight_decay to the device table (call after a scheduler
        step when the optimizer runs inside a captured CUDA graph; eager steps do it automatically)."""
        rows = self._hyper_rows()
        if self._hyper_dev is None:
            return
        self._hyper_dev.copy_(torch.tensor(rows, dtype=torch.float32), non_blocking=False)
        self._hyper_host = rows

    def load_state_dict(self, state_dict) -> None:
        """torch semantics, plus: the shared device step counter is re-seeded from the loaded state (also when this
        optimizer has already stepped, e.g. after a GraphedTrainStep warm-up or a resume in the same process)."""
        super().load_state_dict(state_dict)
        steps = [float(torch.as_tensor(st["step"]).reshape(-1)[0]) for st in self.state.values() if "step" in st]
        if steps:
            if max(steps) != min(steps):
                raise RuntimeError("fer_vit_b200: FusedAdamW keeps one step counter for all tensors; the loaded state "
                                   f"holds different steps ({min(steps)} .. {max(steps)})")
            if self._step_dev is not None:
                self._step_dev.fill_(steps[0])      # in place: a captured graph keeps reading this tensor
                for st in self.state.values():
                    st["step"] = self._step_dev
        self._stepped = bool(steps) and steps[0] > 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ps, gs, ms, vs, numel, group = [], [], [], [], [], []
        dev = None
        for gi, grp in enumerate(self.param_groups):
            for p in grp["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("fer_vit_b200: FusedAdamW runs on CUDA tensors only (no CPU fallback)")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("fer_vit_b200: FusedAdamW needs dense float32 parameters and gradients")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("fer_vit_b200: FusedAdamW needs contiguous parameters and gradients")
                dev = p.device
                st = self.state[p]
                if len(st) != 0 and self._step_dev is None:
                    # state restored by load_state_dict(): adopt its counter as the shared device counter
                    self._step_dev = torch.as_tensor(st["step"], dtype=torch.float32).reshape(-1)[:1].to(p.device).clone()
                if len(st) != 0 and st.get("step") is not self._step_dev:
                    st["step"] = self._step_dev
                if len(st) == 0:
                    if self._step_dev is not None and self._stepped:
                        # every tensor shares ONE device step counter (that is what makes the step graph-capturable); a
                        # tensor that joins later would be bias-corrected with the others' t instead of t = 1
                        raise RuntimeError(
                            "fer_vit_b200: FusedAdamW saw a parameter without optimizer state after it has already "
                            "stepped (a parameter was unfrozen or added mid-training). Build a new FusedAdamW over the "
                            "new trainable set (state_dict() / load_state_dict() carry the moments over)")
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if self._step_dev is None:
                        self._step_dev = torch.zeros(1, dtype=torch.float32, device=p.device)
                    st["step"] = self._step_dev       # one shared device counter (every tensor steps together)
                ps.append(p.data_ptr()); gs.append(p.grad.data_ptr())
                ms.append(st["exp_avg"].data_ptr()); vs.append(st["exp_avg_sq"].data_ptr())
                numel.append(p.numel()); group.append(gi)
        n = len(ps)
        if n == 0:
            return loss
        capturing = torch.cuda.is_current_stream_capturing()
        rows = self._hyper_rows()
        if self._hyper_dev is None or self._hyper_dev.shape[0] != len(rows) or self._hyper_dev.device != dev:
            if capturing:
                raise RuntimeError("fer_vit_b200: run FusedAdamW.step() once eagerly before capturing it in a graph")
            self._hyper_dev = torch.tensor(rows, dtype=torch.float32, device=dev)
            self._hyper_host = rows
        elif rows != self._hyper_host and not capturing:
            self.refresh_hyper()
        numel_arr = (C.c_longlong * n)(*numel)
        if self.max_grad_norm > 0:
            need = int(L.lib().fervit_adamw_scratch_floats(n, numel_arr))
            if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
                if capturing:
                    raise RuntimeError("fer_vit_b200: run FusedAdamW.step() once eagerly before capturing it")
                self._scratch = torch.zeros(need, dtype=torch.float32, device=dev)
            chunks = sum((k + 4095) // 4096 for k in numel)
            self.last_total_norm = self._scratch[chunks + 1]
        arr = lambda xs: (C.c_void_p * n)(*xs)
        L.check(L.lib().fervit_adamw_step(n, arr(ps), arr(gs), arr(ms), arr(vs), numel_arr, (C.c_int * n)(*group),
                                          self._hyper_dev.data_ptr(), self._step_dev.data_ptr(),
                                          C.c_float(self.max_grad_norm),
                                          self._scratch.data_ptr() if self._scratch is not None else None,
                                          torch.cuda.current_stream().cuda_stream))
        self._stepped = True
        return loss
