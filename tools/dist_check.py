"""torchrun check: the sharded operator (NCCL, every output / gradient mode, eager and from a CUDA graph) against the
single-GPU operator on every rank.  Prints one line per mode on rank 0 and exits non-zero on a mismatch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py [full]
"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch, torch.distributed as dist
import cases
import gaussianrpg_b200 as grpg
import gaussianrpg_b200.dist as gd
from gaussianrpg_b200 import synthetic
from gaussianrpg_b200.dist import ShardedGaussianRasterizer
from diff_gaussian_rasterization import GaussianRasterizer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if len(sys.argv) > 1 and sys.argv[1] == "full":
    sc_cpu = synthetic.street_scene(P=2000000, W=1920, H=1280, seed=0)
else:
    sc_cpu = synthetic.street_scene(P=300000, W=1920, H=1066, n_actors=4, actor_points=5000, seed=3)  # ragged last row
sc = sc_cpu.to(dev)
P, H, W = sc.means3D.shape[0], sc.height, sc.width
dL = [t.to(dev) for t in cases.loss_grads(sc_cpu)]
NAMES = ("means3D", "opacities", "shs", "scales", "rotations")


def run(rast, band=False):
    leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in NAMES}
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
    out = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
               scales=leaves["scales"], rotations=leaves["rotations"])
    w = [gd.frame_to_band(t, world, rank) for t in dL[:3]] if band else dL[:3]
    ((out[0] * w[0]).sum() + (out[2] * w[1]).sum() + (out[3] * w[2]).sum()).backward()
    return out, {**{k: v.grad for k, v in leaves.items()}, "means2D": m2d.grad}


ref_out, ref_g = run(GaussianRasterizer(sc.settings()))
ref_frame = torch.cat([ref_out[0], ref_out[2], ref_out[3]], 0)
fail = 0
for output in ("frame", "band"):
    for gradients in ("full", "shard"):
        for fused in ((True, False) if output == "frame" else (True,)):
            gd.FUSED_FORWARD_GATHER = fused
            rast = ShardedGaussianRasterizer(sc.settings(), output=output, gradients=gradients)
            for rep in range(3):  # repeated: the peer frames alternate, a missing barrier would show up as stale rows
                out, g = run(rast, band=(output == "band"))
            torch.cuda.synchronize()
            got = torch.cat([out[0], out[2], out[3]], 0)
            want = ref_frame if output == "frame" else gd.frame_to_band(ref_frame, world, rank)
            if output == "band" and H % 16:  # rows of the padded last tile row hold no pixels
                rows = min(got.shape[1], want.shape[1])
                got, want = got[:, :rows], want[:, :rows]
            img_ok = torch.equal(got, want) and torch.equal(out[1], ref_out[1])
            if gradients == "full":
                errs = {k: cases.rel_err(g[k].cpu().numpy(), ref_g[k].cpu().numpy()) for k in ref_g}
                outside = 0
            else:
                b, c = gd.gaussian_slice(P, world, rank)
                errs = {k: float((g[k][b:b + c] - ref_g[k][b:b + c]).abs().max() / ref_g[k].abs().max()) for k in ref_g}
                outside = sum(int((torch.cat([g[k][:b].reshape(-1), g[k][b + c:].reshape(-1)]) != 0).sum()) for k in ref_g)
            res = torch.tensor([0.0 if img_ok else 1.0, max(errs.values()), float(outside)], device=dev)
            dist.all_reduce(res, op=dist.ReduceOp.MAX)
            bad = float(res[0]) != 0.0 or float(res[1]) > 1e-3 or float(res[2]) != 0
            fail += int(bad)
            if rank == 0:
                print(f"output={output} gradients={gradients} fused_peer_store={fused if output == 'frame' else '-'}: images bit-identical "
                      f"on all ranks: {float(res[0]) == 0.0}, max grad rel err {float(res[1]):.3g}, non-zeros outside the shard "
                      f"{int(res[2])} -> {'FAIL' if bad else 'ok'}", flush=True)

# the training step (band / shard) in static-capacity mode, replayed from a CUDA graph
grpg.set_static_binning(int(1.3 * 16_000_000 / world) + 100000)
rast = ShardedGaussianRasterizer(sc.settings(), output="band", gradients="shard")
leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in NAMES}
w = [gd.frame_to_band(t, world, rank) for t in dL[:3]]
hold = {}


def step():
    for v in leaves.values():
        v.grad = None
    out = rast(means3D=leaves["means3D"], means2D=None, opacities=leaves["opacities"], shs=leaves["shs"],
               scales=leaves["scales"], rotations=leaves["rotations"])
    ((out[0] * w[0]).sum() + (out[2] * w[1]).sum() + (out[3] * w[2]).sum()).backward()
    hold["color"] = out[0].detach()


try:
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize(); dist.barrier()
    hold.clear()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    grpg.check_static_binning()
    b, c = gd.gaussian_slice(P, world, rank)
    errs = {k: float((leaves[k].grad[b:b + c] - ref_g[k][b:b + c]).abs().max() / ref_g[k].abs().max()) for k in NAMES}
    want = gd.frame_to_band(ref_out[0], world, rank)
    rows = min(want.shape[1], hold["color"].shape[1])
    ok = torch.equal(hold["color"][:, :rows], want[:, :rows])
    res = torch.tensor([0.0 if ok else 1.0, max(errs.values())], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    bad = float(res[0]) != 0.0 or float(res[1]) > 1e-3
    fail += int(bad)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"CUDA graph of the training step (band / shard, static capacity): images bit-identical {float(res[0]) == 0.0}, "
              f"max grad rel err {float(res[1]):.3g}, {e0.elapsed_time(e1) / 20:.3f} ms per replay -> {'FAIL' if bad else 'ok'}", flush=True)
except Exception as exc:
    if rank == 0:
        print("CUDA graph capture of the training step failed:", repr(exc)[:500], flush=True)
grpg.set_static_binning(None)
torch.cuda.synchronize()
dist.barrier()
# a NCCL communicator with captured graphs can block in the finalisers: the check is complete, leave without them
sys.stdout.flush(); sys.stderr.flush()
os._exit(1 if fail else 0)
