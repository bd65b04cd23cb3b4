"""Device-side latent data path: packed w+ latents in HBM, ``LatentAugment`` and mixup in one kernel.

The reference feeds the LatentViT trainers one ``.pt`` file per sample through ``LatentFERDataset`` /
``DataLoader(num_workers=4)`` (data/latent_dataset.py:52-135) with ``LatentAugment`` applied per sample on the host
(:6-49), then blends the batch with a permuted copy of itself (train_latent_vit.py:119-127). A latent is 36 KB, so a
whole FER-scale training set (tens of thousands of samples, ~1-2 GB) fits in HBM many times over:
``PackedLatentCache`` reads the same ``.pt`` files once and every batch afterwards is one launch of
``fervit_latent_batch`` (include/fervit_b200.h) - gather, augment and mixup at HBM speed, no host work per sample.

The augmentation keeps the reference's definition (noise -> one scale factor per sample -> element mask) but draws
from a counter-based generator instead of torch's global CPU generator, so a given (seed, batch position) is
reproducible and tests can replay the draws (oracle/reference_math.py: ``latent_augment_draws``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class LatentAugment:
    """Same constructor as the reference ``LatentAugment`` (data/latent_dataset.py:12-26). Calling it on a CUDA latent
    ``[18, 512]`` or batch ``[B, 18, 512]`` runs the native kernel (each row of a batch gets its own scale factor, as
    each sample does in the reference dataset); CPU tensors raise - there is no host path."""

    def __init__(self, noise_std: float = 0.0, scale_range: Optional[Tuple[float, float]] = None,
                 mask_prob: float = 0.0):
        self.noise_std = noise_std
        self.scale_range = scale_range
        self.mask_prob = mask_prob
        self.seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
        self._calls = 0

    def params(self) -> Optional[L.LatentAugmentParams]:
        if not (self.noise_std > 0 or self.scale_range is not None or self.mask_prob > 0):
            return None
        lo, hi = self.scale_range if self.scale_range is not None else (1.0, 1.0)
        return L.LatentAugmentParams(float(max(self.noise_std, 0.0)), int(self.scale_range is not None), float(lo),
                                     float(hi), float(max(self.mask_prob, 0.0)))

    def next_seed(self) -> int:
        """A fresh stream of draws per call (the reference advances torch's global generator)."""
        self._calls += 1
        return (self.seed + 0x9E3779B97F4A7C15 * self._calls) & 0xFFFFFFFFFFFFFFFF

    def __call__(self, latent: torch.Tensor, seed: Optional[int] = None) -> torch.Tensor:
        batched = latent.dim() == 3
        x = latent if batched else latent.unsqueeze(0)
        out, _ = latent_batch(x, transform=self, seed=seed)
        return out if batched else out[0]


def latent_batch(latents: torch.Tensor, labels: Optional[torch.Tensor] = None,
                 sample_idx: Optional[torch.Tensor] = None, transform: Optional[LatentAugment] = None,
                 mix_index: Optional[torch.Tensor] = None, lam=1.0, seed: Optional[int] = None,
                 seed_dev: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                 status: Optional[torch.Tensor] = None):
    """One launch of ``fervit_latent_batch``: ``out[b] = lam * a[b] + (1 - lam) * a[mix_index[b]]`` with
    ``a[b] = transform(latents[sample_idx[b]])``. Returns ``(out [B, ...], labels_out [B] or None)``.

    ``lam`` may be a one-element float32 CUDA tensor and ``seed_dev`` a one-element int64/uint64 CUDA counter; both are
    read on the device, so a captured CUDA graph sees new values on every replay."""
    if not latents.is_cuda:
        raise RuntimeError("fer_vit_b200: latent_batch runs on CUDA tensors only (no CPU fallback) - keep the packed "
                           "latents on the GPU")
    if latents.dtype != torch.float32 or not latents.is_contiguous() or latents.dim() < 2:
        raise RuntimeError("fer_vit_b200: latents must be a contiguous float32 tensor [N, ...]")
    n_rows = latents.shape[0]
    row = latents[0].numel()
    dev = latents.device

    def _idx(t, name):
        if t is None:
            return None
        if not t.is_cuda or t.dtype != torch.int64 or not t.is_contiguous() or t.dim() != 1:
            raise RuntimeError(f"fer_vit_b200: {name} must be a contiguous 1-D int64 CUDA tensor")
        return t

    sample_idx = _idx(sample_idx, "sample_idx")
    mix_index = _idx(mix_index, "mix_index")
    B = sample_idx.numel() if sample_idx is not None else n_rows
    if mix_index is not None and mix_index.numel() != B:
        raise RuntimeError("fer_vit_b200: mix_index must have one entry per batch row")
    if labels is not None:
        labels = _idx(labels, "labels")
        if labels.numel() != n_rows:
            raise RuntimeError("fer_vit_b200: labels must have one entry per latent row")
    if out is None:
        out = torch.empty((B,) + tuple(latents.shape[1:]), dtype=torch.float32, device=dev)
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != B * row or out.device != dev:
        raise RuntimeError("fer_vit_b200: out must be a contiguous float32 CUDA tensor of B rows")
    labels_out = torch.empty(B, dtype=torch.int64, device=dev) if labels is not None else None
    lam_dev = lam if isinstance(lam, torch.Tensor) else None
    if lam_dev is not None and (not lam_dev.is_cuda or lam_dev.dtype != torch.float32 or lam_dev.numel() != 1):
        raise RuntimeError("fer_vit_b200: a tensor lam must be one float32 element on the CUDA device")
    if seed_dev is not None and (not seed_dev.is_cuda or seed_dev.element_size() != 8 or seed_dev.numel() != 1):
        raise RuntimeError("fer_vit_b200: seed_dev must be one 64-bit integer on the CUDA device")
    params = transform.params() if transform is not None else None
    if seed is None:
        seed = transform.next_seed() if transform is not None else 0
    L.check(L.lib().fervit_latent_batch(
        latents.data_ptr(), labels.data_ptr() if labels is not None else None, n_rows,
        sample_idx.data_ptr() if sample_idx is not None else None, B, row,
        C.byref(params) if params is not None else None, C.c_ulonglong(seed & 0xFFFFFFFFFFFFFFFF),
        seed_dev.data_ptr() if seed_dev is not None else None,
        mix_index.data_ptr() if mix_index is not None else None, 0.0 if lam_dev is not None else float(lam),
        lam_dev.data_ptr() if lam_dev is not None else None, out.data_ptr(),
        labels_out.data_ptr() if labels_out is not None else None,
        status.data_ptr() if status is not None else None, _stream_ptr()))
    return out, labels_out


def mixup(latents: torch.Tensor, index: torch.Tensor, lam) -> torch.Tensor:
    """``lam * latents + (1 - lam) * latents[index]`` (train_latent_vit.py:126-127) as one kernel."""
    out, _ = latent_batch(latents, mix_index=index, lam=lam)
    return out


class PackedLatentCache:
    """All latents of a ``LatentFERDataset`` directory packed as one ``[N, 18, 512]`` float32 tensor (+ ``[N]`` int64
    labels) in device memory. Same directory format, ordering and errors as the reference dataset
    (data/latent_dataset.py:71-116): every ``*.pt`` file, sorted by name, holding ``{'latent', 'label', ...}``."""

    CLASS_NAMES = {0: 'angry', 1: 'disgust', 2: 'fear', 3: 'happy', 4: 'neutral', 5: 'sad', 6: 'surprise'}

    def __init__(self, latents: torch.Tensor, labels: torch.Tensor, transform: Optional[LatentAugment] = None,
                 device="cuda"):
        if latents.dim() < 2 or latents.shape[0] != labels.numel():
            raise ValueError("PackedLatentCache: latents [N, ...] and labels [N] must agree")
        self.latents = latents.to(device=device, dtype=torch.float32).contiguous()
        self.labels = labels.to(device=device, dtype=torch.int64).contiguous()
        if not self.latents.is_cuda:
            raise RuntimeError("fer_vit_b200: PackedLatentCache lives in GPU memory (no CPU fallback)")
        self.transform = transform
        self.status = torch.zeros(1, dtype=torch.int32, device=self.latents.device)

    @staticmethod
    def read_dir(latent_dir: str):
        """(latents [N, 18, 512] float32, labels [N] int64) on the host: every ``*.pt`` of the directory in the
        reference dataset's order (sorted file names), read once."""
        if not os.path.exists(latent_dir):
            raise FileNotFoundError(f"Latent directory not found: {latent_dir}")
        files = [os.path.join(latent_dir, f) for f in sorted(os.listdir(latent_dir)) if f.endswith(".pt")]
        if not files:
            raise ValueError(f"No .pt files found in {latent_dir}")
        lat, lab = [], []
        for fp in files:
            try:
                d = torch.load(fp, map_location="cpu", weights_only=True)
                lat.append(d["latent"].float())
                lab.append(int(d["label"]))
            except Exception as e:  # same contract as the reference __getitem__
                raise RuntimeError(f"Error loading {fp}: {e}")
        return torch.stack(lat), torch.tensor(lab, dtype=torch.int64)

    @classmethod
    def from_dir(cls, latent_dir: str, transform: Optional[LatentAugment] = None, device="cuda"):
        host, labels = cls.read_dir(latent_dir)
        if torch.cuda.is_available():
            host = host.pin_memory()
        return cls(host, labels, transform, device)

    def __len__(self) -> int:
        return self.latents.shape[0]

    def get_class_counts(self) -> Dict[int, int]:
        vals, cnt = torch.unique(self.labels, return_counts=True)
        return {int(v): int(c) for v, c in zip(vals.tolist(), cnt.tolist())}

    def get_class_names(self) -> Dict[int, str]:
        return dict(self.CLASS_NAMES)

    def batch(self, sample_idx: torch.Tensor, mix_index: Optional[torch.Tensor] = None, lam=1.0,
              seed: Optional[int] = None, seed_dev: Optional[torch.Tensor] = None, augment: bool = True,
              out: Optional[torch.Tensor] = None):
        """(latents [B, 18, 512], labels [B]) for the given sample indices: gather + transform (+ mixup with
        ``mix_index``/``lam``) in one launch. The labels are the un-mixed ones; pass ``mix_index`` to
        ``mixup_cross_entropy`` as the reference passes ``labels[index]`` to its criterion."""
        return latent_batch(self.latents, self.labels, sample_idx, self.transform if augment else None, mix_index,
                            lam, seed, seed_dev, out, self.status)

    def check(self) -> None:
        """Raise if any batch so far carried an out-of-range index (one device->host read; call it off the hot path)."""
        if int(self.status.item()) != 0:
            raise IndexError("fer_vit_b200: PackedLatentCache.batch saw an out-of-range sample or mixup index")


def get_latent_train_transforms(noise_std: float = 0.1, scale_range: Tuple[float, float] = (0.9, 1.1),
                                mask_prob: float = 0.1) -> LatentAugment:
    """data/latent_dataset.py:138-152."""
    return LatentAugment(noise_std=noise_std, scale_range=scale_range, mask_prob=mask_prob)


def get_latent_val_transforms() -> None:
    """data/latent_dataset.py:155-162."""
    return None
