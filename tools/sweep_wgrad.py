"""Split-K / stage-size sweep of the SwinIR linear weight-gradient shapes (env knobs SRB_WG_SPLITS, SRB_WG_KPIX are
read per call).  Launches are queued behind a device-side delay so the timings are not host-bound."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200.ops.sr_b200 import raw
dev = torch.device('cuda:0')
B, H, W = 16, 64, 64
def bf(c): return torch.randn((B, H, W, c), device=dev).to(torch.bfloat16)
xs = {c: bf(c) for c in (64, 192, 384, 576)}
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.02 * 1.9e9))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
shapes = [(192, 192, 1), (576, 192, 1), (384, 192, 1), (192, 384, 1), (64, 64, 3)]
for n, k, ks in shapes:
    row = []
    for kp in ('', '64', '128', '256'):
        if kp == '256' and k != 64: continue
        for sp in ('', '8', '16', '24', '37', '74'):
            os.environ.pop('SRB_WG_SPLITS', None); os.environ.pop('SRB_WG_KPIX', None)
            if sp: os.environ['SRB_WG_SPLITS'] = sp
            if kp: os.environ['SRB_WG_KPIX'] = kp
            try:  # (every call also zero-fills its fp32 accumulator: ~2 us, the same in every column)
                us = timeit(lambda: raw.wgrad(xs[n], xs[k], ksize=ks))
            except Exception:
                us = float('nan')
            row.append(f'kp{kp or "-"}/s{sp or "-"}:{us:.1f}')
    print(f'wgrad N={n} K={k} ks={ks}: ' + '  '.join(row), flush=True)
