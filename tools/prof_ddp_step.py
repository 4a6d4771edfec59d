"""Timeline evidence for the DDP scaling loss: per-kernel GPU time of the EDSR-L training
step under DistributedDataParallel, with the NCCL kernels named, how much of their run time overlaps compute kernels,
and how much of the step they are exposed (nothing else running).  torch.profiler / CUPTI on rank 0.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/prof_ddp_step.py [--grad-comm bf16]
"""
import argparse
import collections
import os
import re
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import BATCH, EDSR_L, LR, synthetic_batch  # noqa: E402
from basicsr4rs_b200.archs import build_network  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--grad-comm', default='fp32', choices=['fp32', 'bf16'])
ap.add_argument('--steps', type=int, default=4)
ap.add_argument('--ddp', default='torch', choices=['torch', 'flat'])
ap.add_argument('--bucket-mb', type=int, default=25)
args = ap.parse_args()
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
torch.manual_seed(0)
net = build_network(dict(EDSR_L, cuda_graph=True, graph_segments=4, graph_input_shape=[BATCH, 3, LR, LR],
                         flat_grads=args.ddp == 'flat')).to(dev)
model = net
if world > 1 and args.ddp == 'flat':
    from basicsr4rs_b200.utils.flat_ddp import FlatDDP
    model = FlatDDP(net, bucket_mb=args.bucket_mb)
elif world > 1:
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], gradient_as_bucket_view=True,
                                                      bucket_cap_mb=args.bucket_mb)
    if args.grad_comm == 'bf16':
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        model.register_comm_hook(None, default_hooks.bf16_compress_hook)
optim = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.99), fused=True)
lq, gt = (t.to(dev) for t in synthetic_batch(rank))


def step():
    optim.zero_grad(set_to_none=True)
    (model(lq) - gt).abs().mean().backward()
    optim.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [(ev.time_range.start, ev.time_range.end, re.sub(r'^void ', '', re.sub(r'\(.*', '', ev.name))[:60])
           for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort()
    nccl = [e for e in evs if 'nccl' in e[2].lower()]
    comp = [e for e in evs if 'nccl' not in e[2].lower()]

    def union(iv):
        out, cur = 0.0, None
        for a, b, _ in sorted(iv):
            if cur is None or a > cur[1]:
                if cur is not None:
                    out += cur[1] - cur[0]
                cur = [a, b]
            else:
                cur[1] = max(cur[1], b)
        return out + (cur[1] - cur[0] if cur else 0.0)

    def overlap(a_iv, b_iv):  # time where both sets are active = |A| + |B| - |A u B|
        return union(a_iv) + union(b_iv) - union(a_iv + b_iv)

    n = args.steps
    t_nccl, t_comp = union(nccl) / n, union(comp) / n
    t_both = overlap(nccl, comp) / n
    print(f'## EDSR-L B16 DDP world={world} ddp={args.ddp} grad-comm={args.grad_comm} bucket={args.bucket_mb} MB: {ms:.3f} ms/step unprofiled')
    print(f'compute kernels busy {t_comp / 1e3:.3f} ms/step; NCCL kernels busy {t_nccl / 1e3:.3f} ms/step, of which '
          f'{t_both / 1e3:.3f} ms overlap compute and {(t_nccl - t_both) / 1e3:.3f} ms are exposed')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for a, b, k in evs:
        agg[k][0] += 1
        agg[k][1] += b - a
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f'{t / n:10.1f} us {c // n:5d} x {t / c:8.1f} us  {k}')
    # the 256->256 tap-GEMM with and without an NCCL kernel in flight
    tap = [e for e in comp if 'tapgemm_kernel<256' in e[2]]
    if tap and nccl:
        import bisect
        starts = [e[0] for e in nccl]
        with_n, without = [], []
        for a, b, _ in tap:
            i = bisect.bisect_right(starts, b)
            hit = any(nccl[j][0] < b and nccl[j][1] > a for j in range(max(0, i - 8), min(len(nccl), i + 1)))
            (with_n if hit else without).append(b - a)
        if with_n and without:
            print(f'tapgemm<256>: {sum(without) / len(without):.1f} us alone ({len(without)}), '
                  f'{sum(with_n) / len(with_n):.1f} us with an NCCL kernel in flight ({len(with_n)})')
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
