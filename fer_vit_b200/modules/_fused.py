"""Autograd wrapper of the fused pre-module kernel (fervit_premodules_forward / _backward)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib as L
from ..runtime import _require_cuda, _stream_ptr


def _ptr(t: Optional[torch.Tensor]):
    return t.data_ptr() if t is not None else None


class _PreModulesFunction(torch.autograd.Function):
    """y = LEAM(LWN(SPE(x))) with any subset of the three stages (flags)."""

    @staticmethod
    def forward(ctx, x, flags, eps, groups, group_embed, layer_embed, gamma, beta, gate, leam_w):
        _require_cuda(x, "the input")
        use_spe, use_lwn, use_res, use_leam = flags
        xc = x.contiguous().float()
        B, Ln, D = xc.shape
        tensors = [group_embed, layer_embed, gamma, beta, gate, leam_w]
        tensors = [t.contiguous() if t is not None else None for t in tensors]
        p = L.PreModules(int(use_spe), int(use_lwn), int(use_res), int(use_leam), _ptr(tensors[0]), _ptr(tensors[1]),
                         _ptr(groups), _ptr(tensors[2]), _ptr(tensors[3]), _ptr(tensors[4]), _ptr(tensors[5]), eps)
        y = torch.empty_like(xc)
        L.check(L.lib().fervit_premodules_forward(C.byref(p), xc.data_ptr(), B, Ln, D, y.data_ptr(), _stream_ptr()))
        ctx.flags, ctx.eps = flags, eps
        ctx.save_for_backward(xc, groups, *[t if t is not None else torch.empty(0, device=xc.device) for t in tensors])
        ctx.present = [t is not None for t in tensors]
        return y

    @staticmethod
    def backward(ctx, dy):
        use_spe, use_lwn, use_res, use_leam = ctx.flags
        xc, groups, *saved = ctx.saved_tensors
        tensors = [t if ok else None for t, ok in zip(saved, ctx.present)]
        B, Ln, D = xc.shape
        p = L.PreModules(int(use_spe), int(use_lwn), int(use_res), int(use_leam), _ptr(tensors[0]), _ptr(tensors[1]),
                         _ptr(groups) if groups.numel() else None, _ptr(tensors[2]), _ptr(tensors[3]),
                         _ptr(tensors[4]), _ptr(tensors[5]), ctx.eps)
        dev = xc.device
        dyc = dy.contiguous().float()
        dx = torch.empty_like(xc)
        scratch = torch.empty(int(L.lib().fervit_premodules_scratch_floats(B, Ln, D)), dtype=torch.float32, device=dev)
        dgamma = torch.empty(Ln, D, device=dev) if use_lwn else None
        dbeta = torch.empty(Ln, D, device=dev) if use_lwn else None
        dgate = torch.empty(Ln, device=dev) if (use_lwn and use_res) else None
        dlayer = torch.empty(Ln, D, device=dev) if use_spe else None
        dgroup = torch.empty(3, D, device=dev) if use_spe else None
        dleam = torch.empty(Ln, device=dev) if use_leam else None
        L.check(L.lib().fervit_premodules_backward(C.byref(p), xc.data_ptr(), dyc.data_ptr(), B, Ln, D, dx.data_ptr(),
                                                   scratch.data_ptr(), _ptr(dgamma), _ptr(dbeta), _ptr(dlayer),
                                                   _ptr(dgroup), _ptr(dgate), _ptr(dleam), _stream_ptr()))
        return dx, None, None, None, dgroup, dlayer, dgamma, dbeta, dgate, dleam


def premodules(x, *, groups=None, group_embed=None, layer_embed=None, gamma=None, beta=None, gate=None, leam_w=None,
               eps: float = 1e-5):
    flags = (group_embed is not None, gamma is not None, gate is not None, leam_w is not None)
    if groups is None:
        groups = torch.empty(0, dtype=torch.int64, device=x.device)
    return _PreModulesFunction.apply(x, flags, eps, groups, group_embed, layer_embed, gamma, beta, gate, leam_w)
