"""One fused-loss step at 3x1280x1920 (for ncu captures)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from gaussianrpg_b200 import loss_utils
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
gt = torch.rand(3, 1280, 1920, generator=g).to(dev)
img = (gt + 0.1 * torch.randn(3, 1280, 1920, generator=g).to(dev)).clamp(0, 1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    x = img.clone().requires_grad_(True)
    loss_utils.l1_ssim_loss(x, gt, 0.2).backward()
torch.cuda.synchronize()
