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

ref_out, ref_g = run(GaussianRasterizer)
sh_out, sh_g = run(ShardedGaussianRasterizer)
torch.cuda.synchronize()
ok = all(torch.equal(a, b) for a, b in zip((ref_out[0], ref_out[1], ref_out[2], ref_out[3]), (sh_out[0], sh_out[1], sh_out[2], sh_out[3])))
errs = {k: cases.rel_err(sh_g[k].cpu().numpy(), ref_g[k].cpu().numpy()) for k in ref_g}
res = torch.tensor([1.0 if ok else 0.0, max(errs.values())], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN if False else dist.ReduceOp.MAX)
if rank == 0:
    print("images bit-identical on rank0:", ok, "max grad rel err:", errs)
dist.barrier()
dist.destroy_process_group()
