"""Ad-hoc GPU check: ours vs the reference extension (oracle/_ref) on seeded scenes."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import torch
from gaussianrpg_b200 import synthetic, _C, debug
import build_ref
from refparse import parse_reference

ref = build_ref.load()
dev = torch.device("cuda:0")

def run_both(sc, S_grad=True, label=""):
    sc = sc.to(dev)
    P, W, H = sc.means3D.shape[0], sc.width, sc.height
    E = torch.Tensor([])
    sem = sc.semantics if sc.semantics is not None else torch.zeros(P, 0, device=dev)
    def args():
        return (sc.bg, sc.means3D, sc.colors_precomp if sc.colors_precomp is not None else E, sem, sc.opacities,
                sc.scales if sc.scales is not None else E, sc.rotations if sc.rotations is not None else E, sc.scale_modifier,
                sc.cov3D_precomp if sc.cov3D_precomp is not None else E, sc.viewmatrix, sc.projmatrix, sc.tanfovx, sc.tanfovy,
                H, W, sc.shs if sc.shs is not None else E, sc.sh_degree, sc.campos, False, False)
    torch.cuda.synchronize()
    o = _C.rasterize_gaussians(*args(), _reference_binning=True); torch.cuda.synchronize()
    op = _C.rasterize_gaussians(*args()); torch.cuda.synchronize()  # product default: clipped rectangles
    r = ref._C.rasterize_gaussians(*args()); torch.cuda.synchronize()
    Ro, Rr = o[0], r[0]
    res = dict(label=label or sc.name, P=P, R_ours=Ro, R_ref=Rr)
    for name, i in (("color", 1), ("depth", 2), ("alpha", 3), ("semantic", 4)):
        if o[i].numel():
            res[f"maxabs_{name}"] = float((o[i] - r[i]).abs().max())
            res[f"neq_{name}"] = int((o[i] != r[i]).sum())
    res["neq_radii"] = int((o[5] != r[5]).sum())
    res["product_path"] = dict(num_rendered=op[0], neq_color=int((op[1] != r[1]).sum()), neq_depth=int((op[2] != r[2]).sum()),
                               neq_alpha=int((op[3] != r[3]).sum()),
                               binned=debug.parse_buffers(P, op[0], W, H, op[6], op[7], op[8])["num_binned"])
    mine = debug.parse_buffers(P, Ro, W, H, o[6], o[7], o[8])
    theirs = parse_reference(P, Rr, W, H, r[6], r[7], r[8])
    vis = r[5] > 0
    res["V"] = int(vis.sum())
    for k in ("depths", "means2D", "cov3D", "conic_opacity", "rgb"):
        a, b = mine[k][vis], theirs[k][vis]
        res[f"neq_{k}"] = int((a.view(torch.int32) != b.contiguous().view(torch.int32)).sum()) if a.numel() else 0
    res["neq_tiles_touched"] = int((mine["tiles_touched"] != theirs["tiles_touched"]).sum())
    if Ro == Rr and Ro > 0:
        res["neq_point_list"] = int((mine["point_list"] != theirs["point_list"]).sum())
        res["neq_keys"] = int((mine["point_list_keys"] != theirs["point_list_keys"]).sum())
    res["neq_ranges"] = int((mine["ranges"] != theirs["ranges"]).sum())
    res["neq_n_contrib"] = int((mine["n_contrib"] != theirs["n_contrib"]).sum())
    res["mean_n_contrib"] = float(theirs["n_contrib"].float().mean())
    # backward
    g = torch.Generator(device="cpu").manual_seed(1)
    dc = torch.randn(3, H, W, generator=g).to(dev); dd = (torch.randn(1, H, W, generator=g) * 0.1).to(dev)
    da = torch.randn(1, H, W, generator=g).to(dev); ds = torch.randn(sem.shape[1], H, W, generator=g).to(dev)
    def bargs(t):
        return (sc.bg, sc.means3D, t[5], sc.colors_precomp if sc.colors_precomp is not None else E,
                sc.scales if sc.scales is not None else E, sc.rotations if sc.rotations is not None else E, sc.scale_modifier,
                sc.cov3D_precomp if sc.cov3D_precomp is not None else E, sc.viewmatrix, sc.projmatrix, sc.tanfovx, sc.tanfovy,
                dc, dd, da, ds, sc.shs if sc.shs is not None else E, sc.sh_degree, sc.campos, t[6], t[0], t[7], t[8], t[3], sem, False)
    go = _C.rasterize_gaussians_backward(*bargs(op)); torch.cuda.synchronize()
    gr = ref._C.rasterize_gaussians_backward(*bargs(r)); torch.cuda.synchronize()
    names = ["dmeans2D", "dcolors", "dopacity", "dmeans3D", "dcov3D", "dsh", "dscales", "drot", "dsem"]
    for n, a, b in zip(names, go, gr):
        if a.numel():
            den = float(b.abs().max()) + 1e-30
            res[f"rel_{n}"] = float((a - b).abs().max()) / den
    # timing
    def timeit(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
    res["ms_fwd_ours"] = timeit(lambda: _C.rasterize_gaussians(*args()))
    res["ms_fwd_ref"] = timeit(lambda: ref._C.rasterize_gaussians(*args()))
    res["ms_bwd_ours"] = timeit(lambda: _C.rasterize_gaussians_backward(*bargs(op)))
    res["ms_bwd_ref"] = timeit(lambda: ref._C.rasterize_gaussians_backward(*bargs(r)))
    print(json.dumps(res), flush=True)
    return res

if __name__ == "__main__":
    out = []
    out.append(run_both(synthetic.plumbing_scene(P=128, S=3)))
    out.append(run_both(synthetic.plumbing_scene(P=4096, W=200, H=120, S=0, sh_degree=3, white_bg=True)))
    out.append(run_both(synthetic.test_script_scene(P=10000, S=15)))
    out.append(run_both(synthetic.street_scene(P=200000, W=640, H=400, n_actors=2, actor_points=5000)))
    full = len(sys.argv) > 1 and sys.argv[1] == "full"
    if full:
        out.append(run_both(synthetic.street_scene()))
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "gpu_check.json").write_text(json.dumps(out, indent=1))
