"""ncu target: LayerNorm forward / backward launches at B16 x 64x64 x 192 (SwinIR block shapes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200.ops.sr_b200 import swin_ops as so
dev = torch.device('cuda:0')
x = torch.randn((16, 64, 64, 192), device=dev).to(torch.bfloat16)
x[..., 180:] = 0
gy = torch.randn_like(x)
g = torch.ones(180, device=dev); b = torch.zeros(180, device=dev)
for _ in range(4):
    y, mean, rstd = so.layernorm_fwd(x, g, b, 180)
    so.layernorm_bwd(gy, x, mean, rstd, g, 180, gres=gy)
torch.cuda.synchronize()
