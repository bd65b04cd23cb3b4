"""ORACLE — test infrastructure, not product code.

Recipe for `baseline/_ref/`: a verbatim copy of the reference's pure-Python packages that hold the train-step path
(models_fer_vit, modules, data, train, utils), taken from `/root/reference` where it lies. The reference has no
setup.py / pyproject.toml (pip cannot install it) and nothing to compile, so "installing" it is this copy.
`baseline/_ref/` is git-ignored (reference sources never enter this repository's history) but travels to the GPU box
with the gpurun snapshot, where `/root/reference` does not exist; `bench.py --impl reference` and the `cpu_baseline` /
`gpu_eager_reference` legs import the UNMODIFIED classes from it (oracle/ref_runner.py).

    python -m oracle.make_ref          # also run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("models_fer_vit", "modules", "data", "train", "utils")


def make_ref(src: str = SRC, dst: str = DST) -> str | None:
    """Copy the packages; returns dst, or None when the reference tree is not present (GPU box: uses the shipped copy)."""
    if not os.path.isdir(src):
        return dst if os.path.isdir(dst) else None
    os.makedirs(dst, exist_ok=True)
    for pkg in PACKAGES:
        s, d = os.path.join(src, pkg), os.path.join(dst, pkg)
        if not os.path.isdir(s):
            continue
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pt", "*.pth", "*.npz"))
    with open(os.path.join(dst, "PROVENANCE.txt"), "w") as fh:
        fh.write(f"verbatim copy of {', '.join(PACKAGES)} from {src} (yuki-ominato/FER-ViT), made by oracle/make_ref.py; "
                 "not tracked by git\n")
    return dst


if __name__ == "__main__":
    print(make_ref())
