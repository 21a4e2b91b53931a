"""Where does the end-to-end step lose time relative to the device-resident step?  Variants of bench.py's e2e loop."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import bench
from gaussianrpg_b200 import synthetic
import diff_gaussian_rasterization as dgr

dev = torch.device("cuda:0")
sc_cpu = synthetic.street_scene()
sc = sc_cpu.to(dev)
H, W = sc.height, sc.width
g = torch.Generator().manual_seed(123)
gt_host = torch.rand(3, H, W, generator=g).pin_memory()
w_depth = (torch.rand(1, H, W, generator=g) * 0.01).to(dev)
w_alpha = (torch.rand(1, H, W, generator=g) * 0.1).to(dev)
gt_dev = gt_host.to(dev)
cam_host = torch.cat([sc_cpu.viewmatrix.flatten(), sc_cpu.projmatrix.flatten(), sc_cpu.campos.flatten()]).pin_memory()
fwd_only, fwd_bwd, leaves = bench.make_step_fns(dgr, sc, gt_dev, w_depth, w_alpha)
view, proj, campos = sc.viewmatrix, sc.projmatrix, sc.campos
copy_stream = torch.cuda.Stream(device=dev)

def resident():
    fwd_bwd(view, proj, campos, gt_dev)
def resident_sync():
    return float(fwd_bwd(view, proj, campos, gt_dev).item())
def cam_only_sync():
    cam = cam_host.to(dev, non_blocking=True)
    return float(fwd_bwd(cam[:16].view(4, 4), cam[16:32].view(4, 4), cam[32:35], gt_dev).item())
def h2d_main_sync():
    cam = cam_host.to(dev, non_blocking=True)
    gt = gt_host.to(dev, non_blocking=True)
    return float(fwd_bwd(cam[:16].view(4, 4), cam[16:32].view(4, 4), cam[32:35], gt).item())
def h2d_side_sync():
    cam = cam_host.to(dev, non_blocking=True)
    copy_stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(copy_stream):
        gt = gt_host.to(dev, non_blocking=True)
        ev = copy_stream.record_event()
    return float(fwd_bwd(cam[:16].view(4, 4), cam[16:32].view(4, 4), cam[32:35], gt, ev).item())
def h2d_only():
    gt = gt_host.to(dev, non_blocking=True)

for name, fn in (("resident", resident), ("resident+item", resident_sync), ("cam+item", cam_only_sync),
                 ("h2d main+item", h2d_main_sync), ("h2d side+item", h2d_side_sync), ("h2d alone", h2d_only)):
    for _ in range(5):
        fn()
    ms = bench.time_region(fn, 20, False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    print(f"{name:16s} events {ms / 20:.3f} ms/step   wall {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/step")
