"""One eager training step of a BASELINE config (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_all import TRAIN
from basicsr4rs_b200.archs import build_network
name = sys.argv[1]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
opt, batch, lr, _ = TRAIN[name]
dev = torch.device('cuda:0'); torch.manual_seed(0)
net = build_network(opt).to(dev).train()
optim = torch.optim.Adam(net.parameters(), lr=1e-4)
lq = torch.rand((batch, 3, lr, lr), device=dev); gt = torch.rand((batch, 3, 4 * lr, 4 * lr), device=dev)
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize(); torch.cuda.nvtx.range_push('last_step')
    optim.zero_grad(set_to_none=True); (net(lq) - gt).abs().mean().backward(); optim.step()
torch.cuda.synchronize(); print('done')
