"""Build the UNMODIFIED reference extension for sm_100a into oracle/_ref/ (git-ignored).

Sources are compiled where they lie under /root/reference (copied to a temp dir only because
the reference tree is read-only and setuptools writes build products next to setup.py); nothing
but the installed package (its __init__.py + _C*.so) lands in oracle/_ref/.  Used as the GPU
oracle for bit-exact index parity and as the `--impl reference` arm of bench.py.

    python oracle/build_ref.py [--force]

Flags: TORCH_CUDA_ARCH_LIST=10.0a (the reference's setup.py:30 passes no -gencode and no
fast-math) and NVCC_APPEND_FLAGS='-include cstdint' (gcc-13 needs <cstdint> for
cuda_rasterizer/rasterizer_impl.h:24,40-60; a flag, not a source edit).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/submodules/diff-gaussian-rasterization")
DST = HERE / "_ref"


def available() -> bool:
    return any((DST / "diff_gaussian_rasterization").glob("_C*.so"))


def build(force: bool = False) -> Path | None:
    if available() and not force:
        return DST
    if not REF_SRC.exists():
        return None  # on the GPU box: only the prebuilt files are used
    with tempfile.TemporaryDirectory(prefix="grpg_refbuild_") as tmp:
        work = Path(tmp) / "dgr"
        shutil.copytree(REF_SRC, work)
        env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0a", NVCC_APPEND_FLAGS="-include cstdint", MAX_JOBS="6")
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=work, env=env,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference build failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
        pkg = DST / "diff_gaussian_rasterization"
        if pkg.exists():
            shutil.rmtree(pkg)
        pkg.mkdir(parents=True)
        for f in (work / "diff_gaussian_rasterization").iterdir():
            if f.suffix in (".py", ".so"):
                shutil.copy2(f, pkg / f.name)
    return DST


KNN_SRC = Path("/root/reference/submodules/simple-knn")


def knn_available() -> bool:
    return any((DST / "simple_knn").glob("_C*.so"))


def build_knn(force: bool = False) -> Path | None:
    """The UNMODIFIED simple-knn extension (distCUDA2), same recipe; `-include cfloat` supplies FLT_MAX, which
    simple_knn.cu uses without including <cfloat> (a flag, not a source edit)."""
    if knn_available() and not force:
        return DST
    if not KNN_SRC.exists():
        return None
    with tempfile.TemporaryDirectory(prefix="grpg_refbuild_") as tmp:
        work = Path(tmp) / "knn"
        shutil.copytree(KNN_SRC, work, ignore=shutil.ignore_patterns("dist", "*.egg-info", "build"))
        env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0a", NVCC_APPEND_FLAGS="-include cfloat", MAX_JOBS="6")
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=work, env=env,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference simple-knn build failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
        pkg = DST / "simple_knn"
        if pkg.exists():
            shutil.rmtree(pkg)
        pkg.mkdir(parents=True)
        (pkg / "__init__.py").write_text("")
        for f in (work / "simple_knn").iterdir():
            if f.suffix == ".so":
                shutil.copy2(f, pkg / f.name)
    return DST


def load_knn():
    """The reference's `simple_knn._C` under a private name (ours owns `simple_knn`)."""
    import importlib.util
    if not knn_available():
        raise RuntimeError("oracle/_ref/simple_knn is not built")
    name = "_grpg_reference_simple_knn_C"
    if name in sys.modules:
        return sys.modules[name]
    import torch  # noqa: F401
    so = next((DST / "simple_knn").glob("_C*.so"))
    spec = importlib.util.spec_from_file_location("_C", so)  # the module's PyInit symbol is PyInit__C
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def load():
    """Import the reference package under a private name (ours owns `diff_gaussian_rasterization`)."""
    import importlib.util
    if not available():
        raise RuntimeError("oracle/_ref is not built (run python oracle/build_ref.py where /root/reference exists)")
    name = "_grpg_reference_dgr"
    if name in sys.modules:
        return sys.modules[name]
    import torch  # noqa: F401  (the extension links against libtorch)
    pkg = DST / "diff_gaussian_rasterization"
    spec = importlib.util.spec_from_file_location(name, pkg / "__init__.py", submodule_search_locations=[str(pkg)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_knn(force="--force" in sys.argv))
