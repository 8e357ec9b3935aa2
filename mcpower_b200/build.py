"""Build libmcpower_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot).

The template-heavy kernels are split over several translation units (kernel_tu.cu compiled with different
-DMCP_TU_ID / -DMCP_P_LO / -DMCP_P_HI [/-DMCP_TU_GLM]) that build in parallel; api.cu holds the host side.

-fmad=false: the reference's f32 data plane is plain mul/add by contract (rng.rs:199-202); every fused
operation in the kernels is an explicit fma().  -lineinfo keeps ncu's source page usable.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = Path(os.environ.get("MCPOWER_B200_OBJ_DIR", "/tmp/mcpower_b200_build"))  # objects stay out of the repo snapshot
SO = Path(os.environ.get("MCPOWER_B200_SO_OUT", HERE / "libmcpower_b200.so"))  # variants (tools/build_variant.sh) link elsewhere
DEPS = ["api.cu", "kernel_tu.cu", "kernels.cuh", "kernels_q4.cuh", "kernels_lean.cuh", "kernels_glm.cuh", "kernels_lmm.cuh", "device_math.cuh", "pipe_bench.cuh", "launch.h",
        "plan.h", "spec.hpp", "contract.hpp", "orchestrator.hpp", "critvals.hpp", "json.hpp", "../../include/mcpower_b200.h"]
# (tu id, P_lo, P_hi, glm) — must match MCP_TU_LIST in csrc/launch.h
TUS = [(0, 1, 6, False), (1, 7, 9, False), (2, 10, 11, False), (3, 12, 13, False), (4, 14, 15, False), (5, 16, 16, False),
       (6, 1, 8, "GLM"), (7, 9, 12, "GLM"), (8, 13, 16, "GLM"),
       # the mixed-model units are the slowest to compile (random-slope and extra-grouping fits per width): four of them
       (9, 1, 6, "LMM"), (10, 7, 10, "LMM"), (17, 11, 13, "LMM"), (18, 14, 16, "LMM"),
       # padded wide widths (plan.h: MCP_WIDE_P_A / MCP_WIDE_P_B): the split path's kernels only
       (11, 24, 24, False), (12, 32, 32, False), (13, 24, 24, "GLM"), (14, 32, 32, "GLM"), (15, 24, 24, "LMM"), (16, 32, 32, "LMM")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-ccbin", "/usr/bin/g++"]
# the image exports CC=/opt/gcc/bin/gcc (no libgomp spec); nvcc must use the system host compiler (-ccbin)


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


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False, extra=(), only_tus=None) -> Path:
    if not force and not needs_build():
        return SO
    OBJ.mkdir(parents=True, exist_ok=True)
    cc = nvcc()
    jobs = [([cc, *NVCC_FLAGS, *extra, "-c", "-o", str(OBJ / "api.o"), str(CSRC / "api.cu")], OBJ / "api.o")]
    for tid, lo, hi, kind in TUS:
        defs = [f"-DMCP_TU_ID={tid}", f"-DMCP_P_LO={lo}", f"-DMCP_P_HI={hi}"] + ([f"-DMCP_TU_{kind}"] if kind else [])
        obj = OBJ / f"tu{tid}.o"
        jobs.append(([cc, *NVCC_FLAGS, *extra, *defs, "-c", "-o", str(obj), str(CSRC / "kernel_tu.cu")], obj))
    # only_tus (developer iteration): recompile just these objects ("api" or a TU id) and relink with the cached rest
    todo = [j for j in jobs if only_tus is None or j[1].stem.replace("tu", "") in only_tus or not j[1].exists()]
    logs = []
    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
        for out in ex.map(lambda j: _run(j[0], verbose), todo):
            logs.append(out)
    _run([cc, "-shared", "-o", str(SO), *[str(j[1]) for j in jobs], "-ccbin", "/usr/bin/g++"], verbose)
    if verbose:
        sys.stderr.write("".join(logs))
    return SO


if __name__ == "__main__":
    only = None
    if "--tus" in sys.argv:
        only = sys.argv[sys.argv.index("--tus") + 1].split(",")
    extra = ["-Xptxas", "-v"] if "--ptxas" in sys.argv else []
    extra += [a for a in sys.argv[1:] if a.startswith("-D")]
    build(force="--force" in sys.argv or only is not None, verbose=True, extra=extra, only_tus=only)
    print(SO)
