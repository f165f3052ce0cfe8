"""Builds libbhray.so (the C-ABI library: CUDA kernels + host code) in-tree with nvcc for sm_100a.

    python -m bhusie_b200.build [--force] [--verbose] [--hash]

Flags that matter for parity: --fmad=false (no implicit FMA contraction),
IEEE division and square root (nvcc defaults -prec-div=true -prec-sqrt=true; no -use_fast_math).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libbhray.so")
SOURCES = ["bh_abi.cu", "bh_multi.cu", "ray_kernels.cu", "model_host.cpp"]
HEADERS = ["bh_device.h", "bh_objects.h", "detmath.cuh", "ray_impl.cuh", "post_impl.cuh", os.path.join("..", "..", "include", "bh_abi.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off",
    "-shared", "-cudart", "static", "-lz",
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libbhray.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


KERNEL_SOURCES = ["ray_impl.cuh", "ray_kernels.cu", "detmath.cuh", "bh_device.h"]


def kernel_source_hash() -> str:
    """Identifies the device code of the ray pass (sources + the flags that shape it): profiles/trace_kernel_dram.json
    carries the hash of the source its ncu capture was taken from, and bench.py refuses the constants when it differs."""
    import hashlib
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def build_library(force: bool = False, verbose: bool = False, extra: list[str] | None = None, out: str | None = None) -> str:
    """Builds libbhray.so.  `extra`/`out` build an experimental variant (tuning runs) beside the product library."""
    out = out or LIB_PATH
    if not force and out == LIB_PATH and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path(), *NVCC_FLAGS, *(extra or []), "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
           "-o", out, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libbhray.so")
    return out


if __name__ == "__main__":
    if "--hash" in sys.argv:
        print(kernel_source_hash())
    else:
        print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
