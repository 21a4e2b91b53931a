"""No-GPU checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, workspace layout queries run on the host, and the Python surface validates arguments
exactly like the reference before touching the device."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from gaussianrpg_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def _declared_functions():
    names = []
    for header in sorted(p.name for p in (ROOT / "include").glob("*.h")):
        text = (ROOT / "include" / header).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(grpg_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_functions()
    assert declared, "header parse failed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    assert set(declared) == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert lib.grpg_version() >= 100


def test_layout_queries_without_gpu():
    lib = _lib.load()
    g, b, i = _lib.GeomLayout(), _lib.BinningLayout(), _lib.ImageLayout()
    assert lib.grpg_get_geometry_layout(2_000_000, C.byref(g)) == 0
    assert lib.grpg_get_binning_layout(20_000_000, C.byref(b)) == 0
    assert lib.grpg_get_image_layout(1920, 1280, C.byref(i)) == 0
    assert g.rec == 0 and g.total_bytes > 2_000_000 * 48 and g.total_bytes < 2_000_000 * 160
    assert b.point_list % 256 == 0 and b.tile_keys % 256 == 0 and b.total_bytes < 20_000_000 * 24
    assert i.ranges >= 1920 * 1280 * 4 and i.total_bytes < 1920 * 1280 * 4 + 9600 * 8 + 1024
    assert lib.grpg_get_geometry_layout(-1, C.byref(g)) != 0 and b"bad arguments" in lib.grpg_last_error()
    assert lib.grpg_backward_workspace_bytes(1000, 0) >= 48000


def test_python_surface_matches_reference_contract():
    import diff_gaussian_rasterization as dgr
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")
    for fn in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible", "rasterize_gaussians_filter"):
        assert callable(getattr(dgr._C, fn))
    import inspect
    sig = inspect.signature(GaussianRasterizer.forward)
    assert list(sig.parameters)[1:] == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                        "rotations", "cov3D_precomp", "semantics"]
    assert all(sig.parameters[p].default is None for p in list(sig.parameters)[4:])


def test_validation_happens_before_any_device_work():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    st = GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                       torch.zeros(3), False, False)
    r = GaussianRasterizer(st)
    m, o = torch.rand(4, 3), torch.rand(4, 1)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, None, o, scales=torch.rand(4, 3), rotations=torch.rand(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(m, None, o, shs=torch.rand(4, 1, 3))
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):
        r(torch.rand(4, 2), None, o, shs=torch.rand(4, 1, 3), scales=torch.rand(4, 3), rotations=torch.rand(4, 4))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    st = GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                       torch.zeros(3), False, False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        GaussianRasterizer(st)(torch.rand(4, 3), None, torch.rand(4, 1), shs=torch.rand(4, 1, 3),
                               scales=torch.rand(4, 3), rotations=torch.rand(4, 4))


def test_ctypes_structs_match_the_c_header(tmp_path):
    """Compile a tiny C program against include/grpg_b200.h and compare sizeof/offsetof of every struct field
    with the ctypes mirrors in gaussianrpg_b200/_lib.py."""
    import subprocess
    structs = {"grpg_forward_args": _lib.ForwardArgs, "grpg_backward_args": _lib.BackwardArgs,
               "grpg_geom_layout": _lib.GeomLayout, "grpg_binning_layout": _lib.BinningLayout,
               "grpg_image_layout": _lib.ImageLayout, "grpg_l1_ssim_args": _lib.L1SsimArgs, "grpg_rgb8_args": _lib.Rgb8Args,
               "grpg_compose_submodel": _lib.ComposeSubmodel,
               "grpg_adam_tensor": _lib.AdamTensor, "grpg_stats_submodel": _lib.StatsSubmodel,
               "grpg_sky_args": _lib.SkyArgs, "grpg_exchange_args": _lib.ExchangeArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "grpg_b200.h"', '#include "grpg_loss.h"', '#include "grpg_image.h"', '#include "grpg_compose.h"', '#include "grpg_optim.h"', '#include "grpg_sky.h"', '#include "grpg_knn.h"',
             'int main(void){']
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('return 0;}')
    src = tmp_path / "offsets.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "offsets"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        cname, fname, val = line.split()
        ct = structs[cname]
        want = C.sizeof(ct) if fname == "sizeof" else getattr(ct, fname).offset
        assert int(val) == want, f"{cname}.{fname}: C {val} vs ctypes {want}"
