"""Build libmcpower_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot).

-fmad=false: the reference's f32 data plane is plain mul/add by contract (rng.rs:199-202); every fused
operation in the kernels is an explicit fma().  -lineinfo keeps ncu's source page usable.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
SO = HERE / "libmcpower_b200.so"
SOURCES = ["api.cu"]
DEPS = ["api.cu", "kernels.cuh", "device_math.cuh", "plan.h", "spec.hpp", "critvals.hpp", "json.hpp",
        "../../include/mcpower_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not SO.exists():
        return True
    t = SO.stat().st_mtime
    return any((CSRC / d).resolve().stat().st_mtime > t for d in DEPS) or Path(__file__).stat().st_mtime > t


def build(force: bool = False, verbose: bool = False, extra=()) -> Path:
    if not force and not needs_build():
        return SO
    cmd = [nvcc(), *NVCC_FLAGS, *extra, "-o", str(SO), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        print(" ".join(cmd), flush=True)
    env = dict(os.environ)
    # the image exports CC=/opt/gcc/bin/gcc (no libgomp spec); nvcc must use the system host compiler
    res = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], capture_output=True, text=True, env=env)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else ())
    print(SO)
