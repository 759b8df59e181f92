"""In-tree build of libgmmloc_b200.so (hand-written sm_100a CUDA + the C ABI) with nvcc.

    python -m gmmloc_b200.build [--force] [--verbose]

The library travels to the GPU box inside the repo snapshot (git-ignored, not gpurun-ignored).  nvcc
cross-compiles sm_100a without a GPU.  point_gn.cu and view.cu are compiled with -fmad=false (they replay the CPU
oracle's fp64 expression trees bit for bit); the other kernels use explicit intrinsics where exactness
matters and free FMA contraction elsewhere.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB = LIB_DIR / "libgmmloc_b200.so"
OBJ_DIR = PKG / "build"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-Xcompiler", "-fno-fast-math", "--expt-relaxed-constexpr"]
SOURCES = {
    "api.cu": [],
    "map_build.cu": [],
    "grid_build.cu": [],
    "assoc_scan.cu": [],
    "frame_gn.cu": [],
    "frame_gn_mixed.cu": [],
    "frame_gn_world.cu": [],
    "frame_gn_stream.cu": [],
    "pose_lm.cu": [],
    "local_ba.cu": [],
    "point_gn.cu": ["-fmad=false"],
    "view.cu": ["-fmad=false"],
}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return exe


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "gmmloc_b200.h", Path(__file__)]
    return max(p.stat().st_mtime for p in hdrs)


def build(force: bool = False, verbose: bool = False) -> Path:
    LIB_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(exist_ok=True)
    hdr_m = _deps_mtime()
    jobs = []
    objs = []
    env_extra = os.environ.get("GMMLOC_NVCC_EXTRA", "").split()
    if "--debug-env" in sys.argv:  # A/B builds: the library honours the GMMLOC_* environment overrides (common.cuh)
        env_extra.append("-DGMMLOC_DEBUG_ENV")
    for src, extra in SOURCES.items():
        extra = list(extra) + env_extra
        s = CSRC / src
        o = OBJ_DIR / (src + ".o")
        objs.append(o)
        if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, hdr_m):
            cmd = [nvcc(), *ARCH, *COMMON, *extra, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {cmd[-3]}")

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 2)) as ex:
            list(ex.map(run, jobs))
    if jobs or not LIB.exists():
        tmp = LIB.with_suffix(".so.tmp")
        cmd = [nvcc(), *ARCH, "-shared", "-o", str(tmp), *map(str, objs)]
        run(cmd)
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
