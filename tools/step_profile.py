"""CPU issue time vs GPU time of one EDSR-L train step (is the step launch-bound?)."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import EDSR_L, BATCH, LR
from basicsr4rs_b200.archs import build_network
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = build_network(EDSR_L).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-4)
lq = torch.rand((BATCH, 3, LR, LR), device=dev); gt = torch.rand((BATCH, 3, 4*LR, 4*LR), device=dev)
def fwd_bwd():
    opt.zero_grad(set_to_none=True)
    loss = (net(lq) - gt).abs().mean(); loss.backward()
for _ in range(3): fwd_bwd(); opt.step()
torch.cuda.synchronize()
res = {}
for name, fn in (('fwd_bwd', fwd_bwd), ('adam', opt.step)):
    cpu, gpu = [], []
    for _ in range(5):
        if name == 'adam': fwd_bwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        cpu.append((t1 - t0) * 1e3); gpu.append(e0.elapsed_time(e1))
    res[name] = dict(cpu_issue_ms=min(cpu), gpu_ms=min(gpu))
with torch.no_grad():
    for _ in range(2): net(lq)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); net(lq); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    res['fwd_nograd'] = dict(cpu_issue_ms=(t1 - t0) * 1e3, gpu_ms=e0.elapsed_time(e1))
print(json.dumps(res))
