"""Host-side runtime: binds nn.Parameters to a native plan and exposes it to autograd.

PyTorch owns every parameter, gradient, activation workspace and output tensor (SURVEY.md §8b "Ownership"); the
native plan only borrows device pointers for the duration of a call. All compute happens in
libfervit_b200.so on ``torch.cuda.current_stream()``; there is no PyTorch-op or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L

_GEMM_BLOCK_SLOTS = (L.B_QKV_W, L.B_PROJ_W, L.B_FC1_W, L.B_FC2_W, L.B_AD1_W, L.B_AD2_W)

# Parameter gradients the native backward produces TOGETHER (one kernel sequence writes all members: plan.cu keys each
# weight/bias pair on the weight's pointer, the adapter and head groups on one member, the pre-modules on their flags).
# A drop-in nn.Module must honour per-parameter requires_grad, so a request for any member is widened to the whole
# group on the native side and the extra gradients are dropped before they reach autograd.
_BLOCK_GRAD_GROUPS = ((L.B_LN1_W, L.B_LN1_B), (L.B_QKV_W, L.B_QKV_B), (L.B_PROJ_W, L.B_PROJ_B), (L.B_LN2_W, L.B_LN2_B),
                      (L.B_FC1_W, L.B_FC1_B), (L.B_FC2_W, L.B_FC2_B),
                      (L.B_AD1_W, L.B_AD1_B, L.B_AD2_W, L.B_AD2_B, L.B_ALPHA))
_GLOBAL_GRAD_GROUPS = ((L.G_HEAD_LN_W, L.G_HEAD_LN_B, L.G_HEAD_W, L.G_HEAD_B), (L.G_IN_W, L.G_IN_B),
                       (L.G_SPE_GROUP, L.G_SPE_LAYER, L.G_LWN_GAMMA, L.G_LWN_BETA, L.G_LWN_GATE, L.G_LEAM_W))


def grad_groups(depth: int):
    """Every group of parameter slots whose gradients the native backward computes together."""
    groups = [tuple(g) for g in _GLOBAL_GRAD_GROUPS]
    for b in range(depth):
        groups += [tuple(L.bslot(b, s) for s in g) for g in _BLOCK_GRAD_GROUPS]
    return groups


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"fer_vit_b200: {what} is on {t.device}; this implementation runs on CUDA (sm_100a) only and has no CPU "
            "fallback — move the model and its inputs to a CUDA device")


class PlanRunner:
    """One native plan (include/fervit_b200.h: fervit_plan) plus the host bookkeeping around it."""

    def __init__(self, cfg: L.Config):
        self.cfg = cfg
        self._lib = L.lib()
        h = C.c_void_p()
        L.check(self._lib.fervit_plan_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.nslots = self._lib.fervit_plan_num_slots(h)
        self.nstages = self._lib.fervit_plan_num_stages(h)
        self.depth = cfg.depth
        self.bf16 = cfg.mode == L.BF16
        self._ptrs: Optional[tuple] = None
        self._wcache: Optional[torch.Tensor] = None
        self._wstate: Dict[int, tuple] = {}
        self._ws_bytes: Dict[tuple, int] = {}
        self._infer_ws: Dict[int, torch.Tensor] = {}
        self.seed_dev: Optional[torch.Tensor] = None  # device uint64 counter for CUDA-graph replays
        self._seed_calls = 0
        self.grad_sync = None  # optional parallel.GradBucketer
        self.keep_workspace = False   # tests: keep the last saved-for-backward workspace alive (saved_activation)
        self.last_ws: Optional[torch.Tensor] = None
        self.last_batch = 0
        # gradient layout in backward order: head, blocks depth-1..0, input stage
        head = [L.G_HEAD_LN_W, L.G_HEAD_LN_B, L.G_HEAD_W, L.G_HEAD_B]
        inp = [L.G_IN_W, L.G_IN_B, L.G_CLS, L.G_POS, L.G_SPE_GROUP, L.G_SPE_LAYER, L.G_LWN_GAMMA, L.G_LWN_BETA,
               L.G_LWN_GATE, L.G_LEAM_W]
        self.stage_slots: List[List[int]] = [head]
        for k in range(cfg.depth):
            blk = cfg.depth - 1 - k
            self.stage_slots.append([L.bslot(blk, s) for s in range(L.NUM_BLOCK)])
        self.stage_slots.append(inp)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.fervit_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def slot_numel(self, slot: int) -> int:
        return int(self._lib.fervit_plan_slot_numel(self._h, slot))

    # ------------------------------------------------------------------ binding
    def bind(self, tensors: Dict[int, torch.Tensor]) -> None:
        """(Re)bind parameter pointers and refresh stale bf16 weight-cache entries."""
        devs = {t.device for t in tensors.values()}
        if len(devs) > 1:
            raise RuntimeError(f"fer_vit_b200: the model's parameters live on several devices ({sorted(map(str, devs))})")
        ptrs = [0] * self.nslots
        for s, t in tensors.items():
            ptrs[s] = t.data_ptr()
        key = tuple(ptrs)
        if key != self._ptrs:
            for s, t in tensors.items():
                _require_cuda(t, "a parameter")
                want = torch.int64 if s == L.G_SPE_GROUPS else torch.float32
                if t.dtype != want or not t.is_contiguous():
                    raise RuntimeError(f"fer_vit_b200: parameter slot {s} must be contiguous {want} (got {t.dtype})")
                n = self.slot_numel(s)
                if n != t.numel():
                    raise RuntimeError(f"fer_vit_b200: parameter slot {s} has {t.numel()} elements, expected {n}")
            arr = (C.c_void_p * self.nslots)(*[p if p else None for p in ptrs])
            L.check(self._lib.fervit_plan_set_params(self._h, arr, self.nslots))
            self._ptrs = key  # stale cache entries are detected per slot below (data_ptr, version)
        if self.bf16:
            dev = next(iter(tensors.values())).device
            self._update_ln_fold(tensors)
            if self._wcache is None or self._wcache.device != dev:
                nbytes = int(self._lib.fervit_plan_wcache_bytes(self._h))
                self._wcache = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
                L.check(self._lib.fervit_plan_set_wcache(self._h, self._wcache.data_ptr(), self._wcache.numel()))
                self._wstate.clear()
            stale = []
            gemm_slots = [L.G_IN_W] + [L.bslot(b, s) for b in range(self.depth) for s in _GEMM_BLOCK_SLOTS]
            for s in gemm_slots:
                t = tensors.get(s)
                if t is None:
                    continue
                st = (t.data_ptr(), t._version) + self._fold_state(s, tensors)
                # frozen weights: re-cast when (data_ptr, version) moved (load_state_dict, manual edits). Trainable
                # weights: re-cast on EVERY call — fused optimizers (torch.optim.AdamW(fused=True)) update parameters
                # without bumping the version counter, and under CUDA-graph capture the host sees no update at all,
                # so the cast has to be part of the step itself (one batched kernel launch, see plan.cu).
                if self._wstate.get(s) != st or t.requires_grad:
                    stale.append(s)
                    self._wstate[s] = st
            if stale:
                arr = (C.c_int * len(stale))(*stale)
                L.check(self._lib.fervit_plan_refresh_wcache(self._h, arr, len(stale), _stream_ptr()))

    # ------------------------------------------------------------------ LayerNorm folding
    def _update_ln_fold(self, tensors: Dict[int, torch.Tensor]) -> None:
        """Frozen norm + frozen weight behind it => the norm is folded into that GEMM (include/fervit_b200.h:
        fervit_plan_set_ln_fold). Decided per block from requires_grad; FERVIT_LN_FOLD=0 switches it off (1 / 2: norm1 /
        norm2 only)."""
        d = self.depth
        f1, f2 = [0] * d, [0] * d
        cfg = self.cfg
        try:   # bit 0: norm1 -> qkv, bit 1: norm2 -> fc1 (A/B runs); default both
            mask = int(os.environ.get("FERVIT_LN_FOLD", "3"))
        except ValueError:
            mask = 3
        if mask and cfg.norm_first and cfg.E % 128 == 0 and cfg.E <= 1024:
            def frozen(blk, names):
                return all((L.bslot(blk, n) in tensors) and not tensors[L.bslot(blk, n)].requires_grad for n in names)
            for i in range(d):
                f1[i] = int(bool(mask & 1) and i > 0 and frozen(i, (L.B_LN1_W, L.B_LN1_B, L.B_QKV_W, L.B_QKV_B)))
                f2[i] = int(bool(mask & 2) and frozen(i, (L.B_LN2_W, L.B_LN2_B, L.B_FC1_W, L.B_FC1_B)))
        key = (tuple(f1), tuple(f2))
        if key != getattr(self, "_fold_key", None):
            L.check(self._lib.fervit_plan_set_ln_fold(self._h, (C.c_int * d)(*f1), (C.c_int * d)(*f2), d))
            self._fold_key = key
            self.fold1, self.fold2 = f1, f2

    def _fold_state(self, slot: int, tensors: Dict[int, torch.Tensor]) -> tuple:
        """Extra cache-staleness key of a GEMM weight slot: a folded slot's cache entry also depends on the norm's
        gamma / beta and on the bias."""
        if slot < L.NUM_GLOBAL or not getattr(self, "_fold_key", None):
            return ()
        blk, w = divmod(slot - L.NUM_GLOBAL, L.NUM_BLOCK)
        if w == L.B_QKV_W and self.fold1[blk]:
            deps = (L.B_LN1_W, L.B_LN1_B, L.B_QKV_B)
        elif w == L.B_FC1_W and self.fold2[blk]:
            deps = (L.B_LN2_W, L.B_LN2_B, L.B_FC1_B)
        else:
            return ("plain",)
        out = ["folded"]
        for n in deps:
            t = tensors[L.bslot(blk, n)]
            out += [t.data_ptr(), t._version]
        return tuple(out)

    # ------------------------------------------------------------------ workspaces
    def workspace_bytes(self, B: int, save: bool) -> int:
        k = (B, save)
        if k not in self._ws_bytes:
            self._ws_bytes[k] = int(self._lib.fervit_plan_workspace_bytes(self._h, B, 1 if save else 0))
        return self._ws_bytes[k]

    def _next_seed(self, training: bool) -> int:
        if not training or (self.cfg.dropout <= 0.0 and self.cfg.head_dropout <= 0.0):
            return 0
        if torch.cuda.is_current_stream_capturing():
            # the graph bakes this host value in; per-replay variation comes from seed_dev
            self._seed_calls += 1
            return (torch.initial_seed() + 0x9E3779B97F4A7C15 * self._seed_calls) & 0x7FFFFFFFFFFFFFFF
        return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())

    # ------------------------------------------------------------------ forward / backward
    def forward(self, x: torch.Tensor, training: bool, save: bool):
        _require_cuda(x, "the input")
        if self._wcache is not None and self._wcache.device != x.device:
            raise RuntimeError(f"fer_vit_b200: the input is on {x.device} but the model is on {self._wcache.device}")
        if x.dtype != torch.float32:
            raise RuntimeError(f"fer_vit_b200: input must be float32 (got {x.dtype})")
        x = x.contiguous()
        B = x.shape[0]
        nbytes = self.workspace_bytes(B, save)
        if save:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        else:
            ws = self._infer_ws.get(B)
            if ws is None or ws.device != x.device:
                ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
                self._infer_ws = {B: ws}
        logits = torch.empty((B, self.cfg.C), dtype=torch.float32, device=x.device)
        seed = self._next_seed(training)
        sd = self.seed_dev.data_ptr() if self.seed_dev is not None else None
        L.check(self._lib.fervit_plan_forward(self._h, x.data_ptr(), B, ws.data_ptr(), ws.numel(),
                                              1 if training else 0, 1 if save else 0, seed, sd, logits.data_ptr(),
                                              _stream_ptr()))
        if save and self.keep_workspace:
            self.last_ws, self.last_batch = ws, B
        return logits, ws, seed, x

    def saved_activation(self, block: int, which: int = 0) -> torch.Tensor:
        """A per-block activation of the last forward that saved for backward (needs ``keep_workspace = True`` before
        that forward): ``which`` 0 = act'(fc1 pre-activation) — for ReLU the 0/1 active set —, 1 = the activation,
        2 = qkv. Returned as a float32 copy [B*S, width]."""
        if self.last_ws is None:
            raise RuntimeError("fer_vit_b200: set plan_runner().keep_workspace = True before the forward pass")
        ptr, n = C.c_void_p(), C.c_longlong()
        L.check(self._lib.fervit_plan_saved_buffer(self._h, self.last_ws.data_ptr(), self.last_batch, block, which,
                                                   C.byref(ptr), C.byref(n)))
        esize = 2 if self.bf16 else 4
        off = ptr.value - self.last_ws.data_ptr()
        raw = self.last_ws[off:off + n.value * esize]
        t = raw.view(torch.bfloat16 if self.bf16 else torch.float32).float()
        return t.reshape(self.last_batch * (self.cfg.L + 1), -1)

    def backward(self, x: torch.Tensor, ws: torch.Tensor, training: bool, seed: int, dlogits: torch.Tensor,
                 want: Dict[int, torch.Size]) -> Dict[int, torch.Tensor]:
        """Run the native backward; returns {slot: gradient view} for the requested slots (`want`: slot -> shape)."""
        B = x.shape[0]
        # flat fp32 gradient buffer laid out in backward order, 256-byte aligned segments
        offsets: Dict[int, int] = {}
        stage_ranges = []
        off = 0
        for slots in self.stage_slots:
            begin = off
            for s in slots:
                if s in want:
                    offsets[s] = off
                    off += (want[s].numel() + 63) // 64 * 64
            stage_ranges.append((begin, off))
        # zero-filled: a slot the native code did not write must read as "no gradient", never as stale memory
        flat = torch.zeros(max(off, 1), dtype=torch.float32, device=x.device)
        base = flat.data_ptr()
        ptrs = [None] * self.nslots
        for s, o in offsets.items():
            ptrs[s] = base + 4 * o
        arr = (C.c_void_p * self.nslots)(*ptrs)
        dl = dlogits.contiguous()
        sd = self.seed_dev.data_ptr() if self.seed_dev is not None else None
        st = _stream_ptr()

        def run(b: int, e: int) -> None:
            L.check(self._lib.fervit_plan_backward(self._h, x.data_ptr(), B, ws.data_ptr(), ws.numel(),
                                                   1 if training else 0, seed, sd, dl.data_ptr(), arr, self.nslots,
                                                   b, e, st))

        if self.grad_sync is None:
            run(0, self.nstages)
        else:
            # data parallel: all-reduce each finished bucket while earlier blocks are still computing
            for (b, e) in self.grad_sync.stage_groups(self.nstages):
                run(b, e)
                lo, hi = stage_ranges[b][0], stage_ranges[e - 1][1]
                if hi > lo:
                    self.grad_sync.reduce_async(flat[lo:hi])
            self.grad_sync.finish()
        return {s: flat[o:o + want[s].numel()].view(want[s]) for s, o in offsets.items()}


class _PlanFunction(torch.autograd.Function):
    """logits = plan(x; parameters). One autograd node for the whole model."""

    @staticmethod
    def forward(ctx, runner: PlanRunner, slots: Sequence[int], training: bool, save: bool, x: torch.Tensor, *tensors):
        runner.bind(dict(zip(slots, tensors)))
        logits, ws, seed, xc = runner.forward(x, training, save)
        if save:
            ctx.runner, ctx.slots, ctx.training, ctx.seed, ctx.ws = runner, tuple(slots), training, seed, ws
            ctx.shapes = [t.shape for t in tensors]
            # saving the parameter tensors keeps derived inputs (e.g. the stacked LayerWiseNorm gammas) alive until
            # backward, whose kernels read them again, and lets autograd detect in-place edits in between
            ctx.save_for_backward(xc, *tensors)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, *tensors = ctx.saved_tensors
        if x.device.index != torch.cuda.current_device():   # autograd's thread: same device as the forward
            torch.cuda.set_device(x.device)
        ctx.runner.bind(dict(zip(ctx.slots, tensors)))   # another forward may have re-bound the plan since
        if ctx.needs_input_grad[4]:
            raise NotImplementedError("fer_vit_b200: gradient with respect to the model input is not implemented "
                                      "(the reference train step never requests it)")
        need = ctx.needs_input_grad[5:]
        shapes = dict(zip(ctx.slots, ctx.shapes))
        asked = {s for s, n in zip(ctx.slots, need) if n and s != L.G_SPE_GROUPS}
        native = set(asked)
        for grp in grad_groups(ctx.runner.depth):     # per-parameter requires_grad: widen to what is computed together
            if native.intersection(grp):
                native.update(s for s in grp if s in shapes)
        want = {s: shapes[s] for s in native}
        grads = ctx.runner.backward(x, ctx.ws, ctx.training, ctx.seed, dlogits, want)
        ctx.ws = None
        out = [grads.get(s) if n else None for s, n in zip(ctx.slots, need)]
        return (None, None, None, None, None, *out)


def run_plan(runner: PlanRunner, tensors: Dict[int, torch.Tensor], x: torch.Tensor, training: bool) -> torch.Tensor:
    slots = sorted(tensors)
    # grad mode is off inside Function.forward, so decide here whether backward will need the activations
    save = torch.is_grad_enabled() and any(t.requires_grad for t in tensors.values())
    return _PlanFunction.apply(runner, slots, training, save, x, *[tensors[s] for s in slots])


# ----------------------------------------------------------------------------------------------
# Loss
# ----------------------------------------------------------------------------------------------
class _CrossEntropyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, weight, label_smoothing, den):
        _require_cuda(logits, "logits")
        lg = logits.contiguous().float()
        lb = labels.contiguous()
        if lb.dtype != torch.int64:
            lb = lb.long()
        B, Cn = lg.shape
        loss = torch.empty((), dtype=torch.float32, device=lg.device)
        need = logits.requires_grad
        dl = torch.empty_like(lg) if need else None
        w = weight.contiguous().float() if weight is not None else None
        L.check(L.lib().fervit_cross_entropy(lg.data_ptr(), lb.data_ptr(), w.data_ptr() if w is not None else None,
                                             float(label_smoothing), B, Cn,
                                             den.data_ptr() if den is not None else None, 1.0, loss.data_ptr(),
                                             dl.data_ptr() if need else None, None, _stream_ptr()))
        if need:
            ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None, None, None


def cross_entropy(logits, labels, weight=None, label_smoothing: float = 0.0, den=None):
    """Mean cross-entropy with class weights / label smoothing; `den` (device scalar) overrides sum_i w[y_i]."""
    return _CrossEntropyFunction.apply(logits, labels, weight, label_smoothing, den)


class _MixupCrossEntropyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, index, lam, weight, label_smoothing):
        _require_cuda(logits, "logits")
        lg = logits.contiguous().float()
        lb = labels.contiguous().long()
        ix = index.contiguous().long()
        if not lb.is_cuda or not ix.is_cuda:
            raise RuntimeError("fer_vit_b200: mixup_cross_entropy needs labels and index on the CUDA device")
        B, Cn = lg.shape
        if lb.numel() != B or ix.numel() != B:
            raise RuntimeError("fer_vit_b200: mixup_cross_entropy: labels and index must have one entry per row")
        lam_dev = lam if isinstance(lam, torch.Tensor) else None
        if lam_dev is not None and (not lam_dev.is_cuda or lam_dev.dtype != torch.float32 or lam_dev.numel() != 1):
            raise RuntimeError("fer_vit_b200: a tensor lam must be one float32 element on the CUDA device")
        loss = torch.empty((), dtype=torch.float32, device=lg.device)
        need = logits.requires_grad
        dl = torch.empty_like(lg) if need else None
        w = weight.contiguous().float() if weight is not None else None
        L.check(L.lib().fervit_cross_entropy_mixup(
            lg.data_ptr(), lb.data_ptr(), ix.data_ptr(), w.data_ptr() if w is not None else None,
            float(label_smoothing), B, Cn, 0.0 if lam_dev is not None else float(lam),
            lam_dev.data_ptr() if lam_dev is not None else None, 1.0, loss.data_ptr(),
            dl.data_ptr() if need else None, _stream_ptr()))
        if need:
            ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None, None, None, None


def mixup_cross_entropy(logits, labels, index, lam, weight=None, label_smoothing: float = 0.0):
    """``lam * criterion(logits, labels) + (1 - lam) * criterion(logits, labels[index])`` of the LatentViT trainers
    (train_latent_vit.py:131) in one kernel. ``lam``: python float, or a one-element float32 CUDA tensor (read on the
    device, so a captured CUDA graph can change it between replays)."""
    return _MixupCrossEntropyFunction.apply(logits, labels, index, lam, weight, label_smoothing)


class CrossEntropyLoss(torch.nn.Module):
    """Drop-in for nn.CrossEntropyLoss(weight=?, label_smoothing=?) as the reference trainers build it
    (train_hybrid_latent_vit.py:236-241, train_latent_vit.py:248-253), computed by the native kernel."""

    def __init__(self, weight: Optional[torch.Tensor] = None, label_smoothing: float = 0.0):
        super().__init__()
        self.register_buffer("weight", weight)
        self.label_smoothing = label_smoothing

    def forward(self, logits, labels, den=None):
        return cross_entropy(logits, labels, self.weight, self.label_smoothing, den)

    def mixup(self, logits, labels, index, lam):
        """lam * self(logits, labels) + (1 - lam) * self(logits, labels[index]), fused (train_latent_vit.py:131)."""
        return mixup_cross_entropy(logits, labels, index, lam, self.weight, self.label_smoothing)
