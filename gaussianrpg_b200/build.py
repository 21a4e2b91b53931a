"""Build libgrpg_b200.so (the C-ABI library, include/grpg_b200.h) in-tree with nvcc for sm_100a.

    python -m gaussianrpg_b200.build [--force] [--verbose]

The .so lands next to this file so that it travels with the repository snapshot to the GPU
box (no JIT cache).  nvcc cross-compiles sm_100a without a GPU.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libgrpg_b200.so"
OBJ_DIR = HERE / "_build"
SOURCES = ["api.cu", "preprocess_fwd.cu", "binning.cu", "blend_fwd.cu", "blend_bwd.cu", "preprocess_bwd.cu",
           "loss_ssim.cu", "image_epilogue.cu", "scene_compose.cu", "optim.cu", "sky_cubemap.cu", "simple_knn.cu", "record_exchange.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; cannot build libgrpg_b200.so")


def _newest_dep_mtime() -> float:
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "grpg_b200.h", HERE.parent / "include" / "grpg_loss.h", HERE.parent / "include" / "grpg_image.h", HERE.parent / "include" / "grpg_compose.h", HERE.parent / "include" / "grpg_optim.h", HERE.parent / "include" / "grpg_sky.h", HERE.parent / "include" / "grpg_knn.h", Path(__file__)]
    return max(p.stat().st_mtime for p in deps)


def needs_build() -> bool:
    return (not LIB.exists()) or LIB.stat().st_mtime < _newest_dep_mtime()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    hdr_mtime = max(p.stat().st_mtime for p in list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "grpg_b200.h", HERE.parent / "include" / "grpg_loss.h", HERE.parent / "include" / "grpg_image.h", HERE.parent / "include" / "grpg_compose.h", HERE.parent / "include" / "grpg_optim.h", HERE.parent / "include" / "grpg_sky.h", HERE.parent / "include" / "grpg_knn.h"])

    def compile_one(src: str) -> Path:
        s = CSRC / src
        o = OBJ_DIR / (src + ".o")
        if not force and o.exists() and o.stat().st_mtime >= max(s.stat().st_mtime, hdr_mtime):
            return o
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, flush=True)
        return o

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    print(build(force=args.force, verbose=args.verbose))
