"""Build libfervit_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m fer_vit_b200.build [--force] [--verbose]

The shared library is the C-ABI boundary declared in include/fervit_b200.h. It is git-ignored but ships to the
GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libfervit_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE, "-I", CSRC,
]


if os.environ.get("FERVIT_TL_CHUNKS", "0") not in ("", "0"):
    NVCC_FLAGS.append("-DFERVIT_TL_CHUNKS")   # per-chunk epilogue stamps for tools/gemm_timeline.py (slows every GEMM 2-3 %)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest() -> str:
    h = hashlib.sha256()
    for d in (CSRC, INCLUDE):
        for f in sorted(os.listdir(d)):
            if f.endswith((".h", ".cuh")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: str, obj: str, verbose: bool) -> None:
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}")


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    hdig = _headers_digest()
    jobs = []
    objs = []
    for src in _sources():
        with open(src, "rb") as fh:
            dig = hashlib.sha256(fh.read() + hdig.encode()).hexdigest()[:16]
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + "." + dig + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            futs = [ex.submit(_compile_one, s, o, verbose) for s, o in jobs]
            for f in futs:
                f.result()
    stamp = os.path.join(BUILD, "link.stamp")
    want = "\n".join(objs)
    have = open(stamp).read() if os.path.exists(stamp) else ""
    if jobs or force or want != have or not os.path.exists(LIB):
        tmp = LIB + ".tmp"   # link beside the target, then rename: a reader never sees a half-written library
        cmd = [_nvcc(), "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static", "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        os.replace(tmp, LIB)
        with open(stamp, "w") as fh:
            fh.write(want)
        # stale objects of older source revisions
        keep = set(objs)
        for f in os.listdir(BUILD):
            p = os.path.join(BUILD, f)
            if f.endswith(".o") and p not in keep:
                os.remove(p)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
