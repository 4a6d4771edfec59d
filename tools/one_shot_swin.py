import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200 import _lib as L
from basicsr4rs_b200.ops.sr_b200 import raw, swin_ops as so
dev = torch.device('cuda:0'); B, H, W = 16, 64, 64
def bf(*s): return torch.randn(s, device=dev).to(torch.bfloat16)
x192, x384, x576 = bf(B, H, W, 192), bf(B, H, W, 384), bf(B, H, W, 576)
w = torch.randn(576, 192, device=dev) * 0.05; wp = raw.pack_weight(w, 576, 192); bias = torch.zeros(576, device=dev)
w1 = torch.randn(384, 192, device=dev) * 0.05; wp1 = raw.pack_weight(w1, 384, 192); b1 = torch.zeros(384, device=dev)
table = torch.randn(225, 6, device=dev) * 0.1
for _ in range(3):
    raw.tapgemm(x192, wp, ksize=1, cout=576, bias=bias)
    raw.tapgemm(x192, wp1, ksize=1, cout=384, bias=b1, act=L.ACT_GELU, want_aux=True)
    so.window_attention_fwd(x576, table, 6, 8, 4, 30**-0.5)
    so.window_attention_bwd(x576, x192, table, 6, 8, 4, 30**-0.5)
    g = torch.ones(180, device=dev); y, m, r = so.layernorm_fwd(x192, g, g, 180)
    so.layernorm_bwd(x192, x192, m, r, g, 180, gres=x192)
torch.cuda.synchronize(); print('ok')
