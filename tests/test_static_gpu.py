"""Static-capacity forward (grpg_forward_static): no host synchronisation, device-side instance counts, CUDA-graph
capturable.  It must give the default path's images and gradients, report its counters, flag an overflow instead of
writing out of bounds, and replay correctly from a captured graph with new inputs in the same buffers."""
import pytest
import torch

import cases
import gaussianrpg_b200 as grpg
from gaussianrpg_b200 import _C, debug, synthetic
from gaussianrpg_b200.rasterizer import GaussianRasterizer

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset_static():
    yield
    grpg.set_static_binning(None)
    _C._static_counts.clear()


def _scene(dev, **kw):
    return synthetic.street_scene(P=60000, W=400, H=270, n_actors=2, actor_points=3000, seed=9, **kw).to(dev)


def test_static_forward_equals_default_and_reports_counts(cuda_device):
    sc = _scene(cuda_device)
    P, W, H = sc.means3D.shape[0], sc.width, sc.height
    base = cases.raw_forward(_C, sc)
    parsed = debug.parse_buffers(P, base[0], W, H, base[6], base[7], base[8])
    dL = [t.to(cuda_device) for t in cases.loss_grads(sc)]
    gbase = cases.raw_backward(_C, sc, base, dL)
    cap = parsed["num_binned"] + 1000
    st = cases.raw_forward(_C, sc, _static=cap)
    gst = cases.raw_backward(_C, sc, st, dL)
    torch.cuda.synchronize()
    assert st[0] == cap
    for i in (1, 2, 3, 5):
        assert torch.equal(st[i], base[i])
    for n, a, b in zip(cases.GRAD_NAMES, gst, gbase):
        if a.numel():
            assert cases.rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4, n
    (counts,) = grpg.static_binning_counts().values()
    assert counts == (parsed["num_binned"], base[0], 0, int((parsed["tiles_touched"] != 0).sum()))
    grpg.check_static_binning()  # no overflow: does not raise
    # exactly enough room is enough; one instance short is flagged and nothing is written out of bounds
    st2 = cases.raw_forward(_C, sc, _static=parsed["num_binned"])
    torch.cuda.synchronize()
    assert torch.equal(st2[1], base[1])
    grpg.check_static_binning()
    guard = cases.raw_forward(_C, sc, _static=parsed["num_binned"] // 2)
    torch.cuda.synchronize()
    assert tuple(grpg.static_binning_counts().values())[0][2] == 1
    with pytest.raises(RuntimeError, match="capacity"):
        grpg.check_static_binning()
    assert bool(torch.isfinite(guard[1]).all())


def test_static_mode_through_the_module_and_bands(cuda_device):
    sc = _scene(cuda_device)
    kw = sc.raster_kwargs()
    base = GaussianRasterizer(sc.settings())(means2D=None, **kw)
    grpg.set_static_binning(4_000_000)
    out = GaussianRasterizer(sc.settings())(means2D=None, **kw)
    torch.cuda.synchronize()
    grpg.check_static_binning()
    for a, b in zip(base, out):
        assert torch.equal(a, b)
    # tile-row bands (multi-GPU sharding) in static mode reproduce the default path's bands
    E = torch.Tensor([])
    sem = torch.zeros(sc.means3D.shape[0], 0, device=cuda_device)
    for r in range(3):
        args = (sc.bg, sc.means3D, E, sem, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix, sc.projmatrix,
                sc.tanfovx, sc.tanfovy, sc.height, sc.width, sc.shs, sc.sh_degree, sc.campos, False, False)
        a = _C.rasterize_gaussians(*args, _band=(3, r), _static=0)
        b = _C.rasterize_gaussians(*args, _band=(3, r))
        assert torch.equal(a[1], b[1]) and torch.equal(a[3], b[3])


def test_training_step_replays_from_a_cuda_graph(cuda_device):
    """forward + loss + backward captured once (torch.cuda.graph) and replayed with a different camera written into the
    same input buffers: images and gradients must equal the eager default path for that camera."""
    dev = cuda_device
    scenes = [_scene(dev, cam_z=z) for z in (0.0, 1.5)]
    sc0 = scenes[0]
    grpg.set_static_binning(6_000_000)
    leaves = {k: getattr(sc0, k).clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    view, proj, campos = sc0.viewmatrix.clone(), sc0.projmatrix.clone(), sc0.campos.clone()
    weight = torch.rand(3, sc0.height, sc0.width, device=dev)
    static_out = {}

    def step():
        for v in leaves.values():
            v.grad = None
        st = sc0.settings()._replace(viewmatrix=view, projmatrix=proj, campos=campos)
        color, radii, depth, alpha, _ = GaussianRasterizer(st)(means3D=leaves["means3D"], means2D=None,
                                                               opacities=leaves["opacities"], shs=leaves["shs"],
                                                               scales=leaves["scales"], rotations=leaves["rotations"])
        loss = (color * weight).sum() + 0.1 * depth.sum() + alpha.sum()
        loss.backward()
        static_out.update(color=color, loss=loss)

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up on a side stream, as torch.cuda.graph requires
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    grads_static = {k: v.grad for k, v in leaves.items()}
    for sc in reversed(scenes):
        view.copy_(sc.viewmatrix); proj.copy_(sc.projmatrix); campos.copy_(sc.campos)
        graph.replay()
        torch.cuda.synchronize()
        grpg.check_static_binning()
        got_color = static_out["color"].clone()
        got = {k: g.clone() for k, g in grads_static.items()}
        # eager reference for the same camera through the default (synchronising) path
        grpg.set_static_binning(None)
        ref_leaves = {k: v.detach().clone().requires_grad_(True) for k, v in leaves.items()}
        color, radii, depth, alpha, _ = GaussianRasterizer(sc.settings())(
            means3D=ref_leaves["means3D"], means2D=None, opacities=ref_leaves["opacities"], shs=ref_leaves["shs"],
            scales=ref_leaves["scales"], rotations=ref_leaves["rotations"])
        ((color * weight).sum() + 0.1 * depth.sum() + alpha.sum()).backward()
        grpg.set_static_binning(6_000_000)
        assert torch.equal(got_color, color)
        for k in got:
            assert cases.rel_err(got[k].cpu().numpy(), ref_leaves[k].grad.cpu().numpy()) <= 1e-4, k
