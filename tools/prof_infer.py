"""Per-kernel GPU time of one 1024x1024 inference tile (torch.profiler / CUPTI).   python tools/prof_infer.py swinir_infer_1024"""
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_all import INFER  # noqa: E402
from basicsr4rs_b200.archs import build_network  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'swinir_infer_1024'
opt, tile, _ = INFER[name]
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = build_network(opt).to(dev).eval()
x = torch.rand((1, 3, tile, tile), device=dev)
with torch.no_grad():
    for _ in range(2):
        net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    net(x)
    e1.record()
    torch.cuda.synchronize()
    print(f'## {name}: {e0.elapsed_time(e1):.3f} ms/tile unprofiled = {(4 * tile)**2 / e0.elapsed_time(e1) / 1e3:.1f} output MPix/s')
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        net(x)
        torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    k = re.sub(r'\(.*', '', ev.name)
    k = re.sub(r'^void ', '', k)[:70]
    agg[k][0] += 1
    agg[k][1] += ev.device_time
    tot += ev.device_time
print(f'kernel time {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f'{t:10.1f} us {n:5d} x {t / n:8.1f} us {100 * t / tot:5.1f}%  {k}')
