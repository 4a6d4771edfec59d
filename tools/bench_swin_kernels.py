"""Stand-alone timings of the SwinIR-block kernels at B16 x 64x64 (65536 tokens), back-to-back launches
(inputs of one launch are L2-warm from the previous one, as inside a real step)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200 import _lib as L
from basicsr4rs_b200.ops.sr_b200 import raw, swin_ops as so

dev = torch.device('cuda:0')
B, H, W = 16, 64, 64
T = B * H * W
def bf(*shape): return torch.randn(shape, device=dev).to(torch.bfloat16)
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(0.02 * 1.9e9))  # queue the launches behind a device-side delay: device time, not host pace
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
rows = []
def rec(name, us, gflop=None, mbytes=None):
    r = dict(kernel=name, us=round(us, 1))
    if gflop: r['tflops'] = round(gflop / us * 1e-3 * 1e3, 1)
    if mbytes: r['gbs'] = round(mbytes / us * 1e3 / 1e3 * 1e3 / 1e3, 1) if False else round(mbytes * 1e6 / (us * 1e-6) / 1e9, 1)
    rows.append(r); print(json.dumps(r), flush=True)

x192 = bf(B, H, W, 192); x384 = bf(B, H, W, 384); x576 = bf(B, H, W, 576)
def lin(n, k): return (torch.randn(n, k, device=dev) * 0.05)
def gemm_case(name, xin, n, k, **kw):
    w = lin(n, k); wp = raw.pack_weight(w, n, k); bias = torch.zeros(n, device=dev)
    us = timeit(lambda: raw.tapgemm(xin, wp, ksize=1, cout=n, bias=bias, **kw))
    rec(name, us, gflop=2.0 * T * n * k / 1e9, mbytes=(T * k * 2 + T * n * 2) / 1e6)
gemm_case('qkv 192->576', x192, 576, 192)
gemm_case('proj 192->192 +res', x192, 192, 192, residual=x192)
gemm_case('fc1 192->384 gelu+aux', x192, 384, 192, act=L.ACT_GELU, want_aux=True)
gemm_case('fc2 384->192 +res', x384, 192, 384, residual=x192)
gemm_case('dgrad fc2 (N=384,K=192) dgelu mask', x192, 384, 192, mask_src=x384, mask_mode=L.MASK_DGELU, flip=True)
gemm_case('dgrad qkv (N=192,K=576)', x576, 192, 576, flip=True)
for n, k, dy, xx in ((192, 384, x192, x384), (384, 192, x384, x192), (576, 192, x576, x192), (192, 192, x192, x192)):
    us = timeit(lambda: raw.wgrad(dy, xx, ksize=1))
    rec(f'wgrad N={n} K={k}', us, gflop=2.0 * T * n * k / 1e9)
    us = timeit(lambda: raw.colsum(dy)); rec(f'colsum C={n}', us, mbytes=T * n * 2 / 1e6)
g = torch.ones(180, device=dev); b_ = torch.zeros(180, device=dev)
us = timeit(lambda: so.layernorm_fwd(x192, g, b_, 180)); rec('layernorm_fwd', us, mbytes=2 * T * 192 * 2 / 1e6)
y, mean, rstd = so.layernorm_fwd(x192, g, b_, 180)
us = timeit(lambda: so.layernorm_bwd(x192, x192, mean, rstd, g, 180, gres=x192)); rec('layernorm_bwd(+gres)', us, mbytes=4 * T * 192 * 2 / 1e6)
table = torch.randn(225, 6, device=dev) * 0.1
for shift in (0, 4):
    us = timeit(lambda: so.window_attention_fwd(x576, table, 6, 8, shift, 30**-0.5)); rec(f'attn_fwd shift{shift}', us, mbytes=4 * T * 180 * 2 / 1e6)
    _, stats = so.window_attention_fwd(x576, table, 6, 8, shift, 30**-0.5, want_stats=True)
    us = timeit(lambda: so.window_attention_bwd(x576, x192, table, 6, 8, shift, 30**-0.5, stats=stats, use_tc=True)); rec(f'attn_bwd shift{shift}', us, mbytes=7 * T * 180 * 2 / 1e6)
    us = timeit(lambda: so.window_attention_bwd(x576, x192, table, 6, 8, shift, 30**-0.5)); rec(f'attn_bwd(mma.sync) shift{shift}', us, mbytes=7 * T * 180 * 2 / 1e6)
w = lin(576, 192)
us = timeit(lambda: raw.pack_weight(w, 576, 192)); rec('pack 576x192', us)
acc = torch.zeros(1, 576, 192, device=dev)
us = timeit(lambda: raw.unpack_wgrad(acc, (576, 192))); rec('unpack 576x192', us)
us = timeit(lambda: torch.zeros(200000, device=dev)); rec('torch.zeros 0.8MB', us)
