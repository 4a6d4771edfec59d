"""Per-kernel GPU time of a CUDA-graph-replayed training step (torch.profiler / CUPTI; low overhead, kernels
are timed inside the real replay, unlike an ncu launch list).   python tools/prof_step.py edsr_l [steps]"""
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_all import TRAIN  # noqa: E402
from basicsr4rs_b200.archs import build_network  # noqa: E402

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
graph = os.environ.get('NO_GRAPH') is None
opt, batch, lr, _ = TRAIN[name]
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = build_network(dict(opt, cuda_graph=graph)).to(dev).train()
if os.environ.get('TORCH_ADAM'):
    optim = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.99), fused=True)
else:
    from basicsr4rs_b200.utils.fused_adam import FusedAdam
    optim = FusedAdam(net.parameters(), lr=1e-4, betas=(0.9, 0.99))
lq = torch.rand((batch, 3, lr, lr), device=dev)
gt = torch.rand((batch, 3, 4 * lr, 4 * lr), device=dev)


def step():
    optim.zero_grad(set_to_none=True)
    (net(lq) - gt).abs().mean().backward()
    optim.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
print(f'## {name}: {e0.elapsed_time(e1) / steps:.3f} ms/step unprofiled')
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
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
print(f'kernel time {tot / steps / 1e3:.3f} ms/step over {sum(a[0] for a in agg.values()) // steps} launches/step')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f'{t / steps:10.1f} us {n // steps:5d} x {t / n:8.1f} us {100 * t / tot:5.1f}%  {k}')
