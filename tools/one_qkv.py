import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200.ops.sr_b200 import raw
dev = torch.device('cuda:0')
x = torch.randn((16, 64, 64, 192), device=dev).to(torch.bfloat16)
w = torch.randn(576, 192, device=dev) * 0.05
wp = raw.pack_weight(w, 576, 192); bias = torch.zeros(576, device=dev)
for _ in range(4):
    raw.tapgemm(x, wp, ksize=1, cout=576, bias=bias)
torch.cuda.synchronize()
