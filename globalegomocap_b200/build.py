"""Builds globalegomocap_b200/libgem_b200.so in-tree with nvcc for sm_100a.

    python -m globalegomocap_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libgem_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
# translation units that emulate ATen's separately-rounded fp32 ops must not contract mul+add
SOURCES = {
    "api.cu": [],
    "energy.cu": ["--fmad=false"],
    "lbfgs.cu": ["--fmad=false"],
    "pose_ops.cu": ["--fmad=false"],
    "lift.cu": ["--fmad=false"],
    "metrics.cu": [],
    "gemm_simt.cu": [],
    "gemm_tc.cu": [],
    "gemm_tap_tc.cu": ["--fmad=false"],       # includes energy_device.cuh (the chain kernel's energy prologue)
}


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "gem_b200.h"))
    objs = []
    logs = []
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        stamp = obj + ".sha"
        dig = _digest([path] + headers) + " ".join(extra)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [_nvcc(), *ARCH, *COMMON, *extra, "-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(res.stderr)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        if verbose:
            sys.stderr.write(res.stderr)
        with open(stamp, "w") as f:
            f.write(dig)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc(), *ARCH, "-shared", "-o", LIB, *objs]  # static cudart (nvcc default)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
