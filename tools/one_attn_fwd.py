"""ncu target: a few launches of the window-attention FORWARD at B16 x 64x64 (shift 4)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200.ops.sr_b200 import swin_ops as so
dev = torch.device('cuda:0')
B, H, W = 16, 64, 64
x576 = torch.randn((B, H, W, 576), device=dev).to(torch.bfloat16)
table = torch.randn(225, 6, device=dev) * 0.1
for _ in range(6):
    so.window_attention_fwd(x576, table, 6, 8, 4, 30**-0.5)
torch.cuda.synchronize()
