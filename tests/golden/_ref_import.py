"""Import pieces of the reference's Python product tree (/root/reference/lib) on the CPU of the build container.

Used only by the golden-vector generators in this directory.  Third-party packages that are absent from the image
(plyfile, bidict, roma, simple_knn, nvdiffrast, ...) and `lib.config` (which parses sys.argv and reads a YAML at
import time) are replaced by inert stubs; the reference functions under test are imported untouched.  The reference
hard-codes `.cuda()` / `device='cuda'` in a few of them; `cpu_device()` redirects those to the CPU for the duration
of a call so that the very same lines run here without a GPU.
"""
import contextlib
import importlib
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
REFERENCE = Path("/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return None

    def get(self, k, default=None):
        return default


def install_stubs():
    for name in ("roma", "plyfile", "bidict", "simple_knn", "simple_knn._C", "nvdiffrast", "nvdiffrast.torch",
                 "imageio", "matplotlib", "matplotlib.pyplot", "open3d", "trimesh", "skimage", "lpips"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    cfgmod = types.ModuleType("lib.config")
    cfgmod.cfg = _Stub("cfg")
    cfgmod.args = _Stub("args")
    sys.modules["lib.config"] = cfgmod
    for p in (str(REFERENCE), str(ROOT)):  # ROOT provides the drop-in `diff_gaussian_rasterization` package
        if p not in sys.path:
            sys.path.insert(0, p)


def load(module: str):
    install_stubs()
    return importlib.import_module(module)


@contextlib.contextmanager
def cpu_device():
    """Run reference code that says `.cuda()` / `device='cuda'` on the CPU."""
    orig_cuda, orig_zeros, orig_ones = torch.Tensor.cuda, torch.zeros, torch.ones

    def _strip(fn):
        def wrapped(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k = dict(k)
                k.pop("device")
            return fn(*a, **k)
        return wrapped

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.zeros, torch.ones = _strip(orig_zeros), _strip(orig_ones)
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.zeros, torch.ones = orig_cuda, orig_zeros, orig_ones
