"""Tap-folded vs per-tap weight gradient of the 64-channel 3x3 layers (RCAN / EDSR-M), split-K sweep."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200.ops.sr_b200 import raw
dev = torch.device('cuda:0')
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.02 * 1.9e9))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (b, h, w, n, k) in ((16, 48, 48, 64, 64), (16, 48, 48, 256, 64), (16, 96, 96, 64, 64)):
    dy = torch.randn((b, h, w, n), device=dev).to(torch.bfloat16)
    x = torch.randn((b, h, w, k), device=dev).to(torch.bfloat16)
    row = []
    for nofold in ('1', ''):
        for kp in ('', '64'):
            for sp in ('', '12', '24', '36'):
                for key in ('SRB_WG_FOLD', 'SRB_WG_KPIX', 'SRB_WG_SPLITS'): os.environ.pop(key, None)
                os.environ['SRB_WG_FOLD'] = '0' if nofold else '1'
                if kp: os.environ['SRB_WG_KPIX'] = kp
                if sp: os.environ['SRB_WG_SPLITS'] = sp
                us = timeit(lambda: raw.wgrad(dy, x, ksize=3))
                row.append(f'{"per-tap" if nofold else "fold"}/kp{kp or "-"}/s{sp or "-"}:{us:.1f}')
    print(f'wgrad B{b} {h}x{w} N={n} K={k}: ' + '  '.join(row), flush=True)
