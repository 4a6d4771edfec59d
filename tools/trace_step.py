"""Warm-cache, in-sequence kernel times of one eager train step: a CUDA event after every C-ABI launch; the delta
between consecutive events is attributed to the later launch (torch's own kernels in between included)."""
import collections, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_all import TRAIN
from basicsr4rs_b200 import _lib as L
from basicsr4rs_b200.archs import build_network
name = sys.argv[1]
opt, batch, lr, _ = TRAIN[name]
dev = torch.device('cuda:0'); torch.manual_seed(0)
net = build_network(opt).to(dev).train()
optim = torch.optim.Adam(net.parameters(), lr=1e-4)
lq = torch.rand((batch, 3, lr, lr), device=dev); gt = torch.rand((batch, 3, 4 * lr, 4 * lr), device=dev)
def step():
    optim.zero_grad(set_to_none=True); (net(lq) - gt).abs().mean().backward(); optim.step()
for _ in range(3): step()
torch.cuda.synchronize()
L.TRACE = []
e0 = torch.cuda.Event(enable_timing=True); e0.record()
step()
e1 = torch.cuda.Event(enable_timing=True); e1.record()
torch.cuda.synchronize()
tr, L.TRACE = L.TRACE, None
agg = collections.defaultdict(lambda: [0, 0.0]); prev = e0
for what, ev in tr:
    dt = prev.elapsed_time(ev) * 1e3; prev = ev
    agg[what][0] += 1; agg[what][1] += dt
total = e0.elapsed_time(e1) * 1e3
print(json.dumps({'config': name, 'step_us': total, 'launches': len(tr)}))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{t:10.1f} us {n:5d} x {t/n:8.1f} us {100*t/total:5.1f}%  {k}')
