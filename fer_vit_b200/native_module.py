"""Base class of the drop-in model classes: lazily builds the native plan and routes forward() through it."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib as L
from .runtime import PlanRunner, run_plan

_DEFAULT_PRECISION = os.environ.get("FERVIT_PRECISION", "bf16")


def set_default_precision(precision: str) -> None:
    """'bf16' (tcgen05 tensor-core GEMMs, fp32 accumulate and residual stream; 2e-2 parity) or
    'fp32' (CUDA-core fp32 GEMMs; 1e-4 parity). Applies to models constructed afterwards."""
    global _DEFAULT_PRECISION
    if precision not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _DEFAULT_PRECISION = precision


def get_default_precision() -> str:
    return _DEFAULT_PRECISION


class NativeModule(nn.Module):
    """nn.Module whose forward runs as one native plan. Subclasses provide _plan_config() and _plan_tensors()."""

    def __init__(self) -> None:
        super().__init__()
        self.__dict__["_precision"] = _DEFAULT_PRECISION
        self.__dict__["_runner"] = None
        self.__dict__["_runner_key"] = None

    # precision is host-side state, not a parameter: it never appears in state_dict()
    @property
    def precision(self) -> str:
        return self.__dict__["_precision"]

    @precision.setter
    def precision(self, value: str) -> None:
        if value not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.__dict__["_precision"] = value
        self.__dict__["_runner"] = None

    def _plan_config(self) -> L.Config:  # pragma: no cover - abstract
        raise NotImplementedError

    def _plan_tensors(self) -> Dict[int, torch.Tensor]:  # pragma: no cover - abstract
        raise NotImplementedError

    def _plan_key(self):
        return None

    def plan_runner(self) -> PlanRunner:
        key = (self.precision, self._plan_key())
        if self.__dict__["_runner"] is None or self.__dict__["_runner_key"] != key:
            cfg = self._plan_config()
            cfg.mode = L.BF16 if self.precision == "bf16" else L.F32
            grad_sync = getattr(self.__dict__["_runner"], "grad_sync", None)
            self.__dict__["_runner"] = PlanRunner(cfg)
            self.__dict__["_runner"].grad_sync = grad_sync
            self.__dict__["_runner_key"] = key
        return self.__dict__["_runner"]

    def _native_forward(self, x: torch.Tensor) -> torch.Tensor:
        # the native launches go to the CURRENT device's stream and the library's per-process caches (SM count, kernel
        # attributes) belong to it: run under the input's device so that a model on cuda:1 works without the caller
        # having called torch.cuda.set_device(1), as it would with PyTorch ops
        if x.is_cuda and x.device.index != torch.cuda.current_device():
            with torch.cuda.device(x.device):
                return run_plan(self.plan_runner(), self._plan_tensors(), x, self.training)
        return run_plan(self.plan_runner(), self._plan_tensors(), x, self.training)

    def __getstate__(self):
        state = super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__.copy()
        state = dict(state)
        state["_runner"] = None
        state["_runner_key"] = None
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ("_runner", "_runner_key") else copy.deepcopy(v, memo)
        return new


def encoder_layer_tensors(layer: nn.TransformerEncoderLayer, blk: int) -> Dict[int, torch.Tensor]:
    """Slots of one nn.TransformerEncoderLayer (keys transformer.layers.{i}.*)."""
    a = layer.self_attn
    return {
        L.bslot(blk, L.B_LN1_W): layer.norm1.weight, L.bslot(blk, L.B_LN1_B): layer.norm1.bias,
        L.bslot(blk, L.B_QKV_W): a.in_proj_weight, L.bslot(blk, L.B_QKV_B): a.in_proj_bias,
        L.bslot(blk, L.B_PROJ_W): a.out_proj.weight, L.bslot(blk, L.B_PROJ_B): a.out_proj.bias,
        L.bslot(blk, L.B_LN2_W): layer.norm2.weight, L.bslot(blk, L.B_LN2_B): layer.norm2.bias,
        L.bslot(blk, L.B_FC1_W): layer.linear1.weight, L.bslot(blk, L.B_FC1_B): layer.linear1.bias,
        L.bslot(blk, L.B_FC2_W): layer.linear2.weight, L.bslot(blk, L.B_FC2_B): layer.linear2.bias,
    }


def base_config() -> L.Config:
    c = L.Config()
    c.eps_head = 1e-5
    c.eps_lwn = 1e-5
    return c
