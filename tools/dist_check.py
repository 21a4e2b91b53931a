"""torchrun check: sharded op (NCCL) vs the single-GPU op on every rank."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch, torch.distributed as dist
import cases
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
    sc_cpu = synthetic.street_scene(P=300000, W=1920, H=1066, n_actors=4, actor_points=5000, seed=3)
sc = sc_cpu.to(dev)
dL = [t.to(dev) for t in cases.loss_grads(sc_cpu)]

def run(rast_cls, **kw):
    leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    m2d = torch.zeros(sc.means3D.shape[0], 3, device=dev, requires_grad=True)
    out = rast_cls(sc.settings(), **kw)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                       shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    loss = (out[0] * dL[0]).sum() + (out[2] * dL[1]).sum() + (out[3] * dL[2]).sum()
    loss.backward()
    return out, {**{k: v.grad for k, v in leaves.items()}, "means2D": m2d.grad}

import gaussianrpg_b200.dist as gdist

def fwd_ms(n=30):
    rast = ShardedGaussianRasterizer(sc.settings())
    m2d = torch.zeros(sc.means3D.shape[0], 3, device=dev)
    with torch.no_grad():
        for _ in range(5):
            rast(means3D=sc.means3D, means2D=m2d, opacities=sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            rast(means3D=sc.means3D, means2D=m2d, opacities=sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ref_out, ref_g = run(GaussianRasterizer)
def step_ms(n=20):
    for _ in range(3):
        run(ShardedGaussianRasterizer)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        run(ShardedGaussianRasterizer)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for fused, fused_rec in ((True, True), (True, False), (False, False)):
    gdist.FUSED_FORWARD_GATHER = fused
    gdist.FUSED_RECORD_REDUCE_MAX_RANKS = 8 if fused_rec else 0
    for rep in range(3):  # repeated: the peer frame is reused, so a missing barrier would show up as stale rows
        sh_out, sh_g = run(ShardedGaussianRasterizer)
    torch.cuda.synchronize()
    ok = all(torch.equal(a, b) for a, b in zip((ref_out[0], ref_out[1], ref_out[2], ref_out[3]), (sh_out[0], sh_out[1], sh_out[2], sh_out[3])))
    errs = {k: cases.rel_err(sh_g[k].cpu().numpy(), ref_g[k].cpu().numpy()) for k in ref_g}
    res = torch.tensor([0.0 if ok else 1.0, max(errs.values())], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    ms = fwd_ms()
    sms = step_ms()
    if rank == 0:
        print(f"fused_peer_store={fused} fused_record_reduce={fused_rec} peer_frame={'yes' if any(v is not None for v in gdist._PeerFrame._cache.values()) else 'no'}:",
              "images bit-identical on all ranks:", float(res[0]) == 0.0, "max grad rel err: %.3g" % float(res[1]),
              "sharded forward %.3f ms" % ms, "fwd+bwd (incl. leaf clones) %.3f ms" % sms)
dist.barrier()
dist.destroy_process_group()
