#!/usr/bin/env python
"""Headline benchmark: SR train patches/s of EDSR-L x4 (BASELINE.json configs[1]) on N B200s.

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                          # reference algorithm on the host CPU

One "step" = zero_grad + forward + L1 loss + backward + Adam step of EDSR-L x4 (32 ResBlocks, 256 ch,
res_scale 0.1) on 16 synthetic 3x48x48 LR patches per GPU (weak scaling, DDP over NCCL).  Prints ONE
JSON line (see the task contract): `value` = device-resident throughput, `e2e` = the same through the
public nn.Module API with pinned-host inputs copied H2D and the loss read back D2H every step,
`roofline` = the dominant kernel (conv3x3 256->256 tap-GEMM) timed live with CUDA events,
`cpu_baseline` = the CPU oracle port on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EDSR_L = dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=256, num_block=32, upscale=4, res_scale=0.1,
              img_range=255., rgb_mean=[0.4488, 0.4371, 0.4040])
BATCH, LR = 16, 48
CONFIG = {'launch': 'CUDA-graph replay (4 segments) of fwd+bwd; L1 loss + torch.optim.Adam(fused=True) eager', 'workload': 'EDSR-L x4 train step (fwd + L1 + bwd + Adam), 16x3x48x48 LR patches per GPU',
          'arch': 'EDSR num_feat=256 num_block=32 res_scale=0.1 upscale=4', 'batch_per_gpu': BATCH,
          'lr_patch': LR, 'parallelism': 'ddp', 'l2': 'per-step working set (>2 GB of activations) exceeds the 126 MB L2'}
FLOP_PER_PATCH_FWD_BWD = 694.66e9
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, mean of the four 256->256 launches in
# profiles/r01_ncu_full_tapgemm_v2.txt (one `ncu --set full` capture of this script): 39.09 / 20.14 / 40.10 / 20.14 MB.  The
# algorithmic bytes are 18.9 MB in + 18.9 MB out + 1.2 MB weights (+18.9 MB residual for conv2 / dgrad-conv1); the
# output is still L2-resident when the kernel ends, so DRAM sees only the compulsory reads -- no wasted re-reads.
ROOFLINE_TRAFFIC_BYTES = 29.87e6


def synthetic_batch(rank, batch=BATCH, lr=LR, scale=4):
    """Per-rank synthetic shard: LR patches in [0,1) and GT, seeded 1234 + rank."""
    g = torch.Generator().manual_seed(1234 + rank)
    lq = torch.rand((batch, 3, lr, lr), generator=g)
    gt = torch.rand((batch, 3, scale * lr, scale * lr), generator=g)
    return lq, gt


def max_over_ranks(ms, device, world):
    """Device-timed milliseconds -> max over ranks (the job is as slow as its slowest rank)."""
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def cpu_oracle_step_time(n_patches, threads, iters=1):
    """Reference algorithm (oracle port, fp32) on the host CPU: fwd + L1 + bwd of EDSR-L on `n_patches`."""
    from oracle import sr_oracle
    from basicsr4rs_b200.archs import build_network
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = build_network(EDSR_L)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    lq = torch.rand((n_patches, 3, LR, LR), generator=g)
    gt = torch.rand((n_patches, 3, 4 * LR, 4 * LR), generator=g)
    best = None
    for _ in range(iters + 1):  # first pass = warm-up
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        out = sr_oracle.edsr_forward(sd, lq, num_block=32, upscale=4, res_scale=0.1, img_range=255.)
        (out - gt).abs().mean().backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 2
    times = []
    for _ in range(max(1, args.warmup > 0) + args.steps):
        times.append(cpu_oracle_step_time(n, threads, iters=0))
        if sum(times) > 150:
            break
    t = sorted(times[1:] or times)[len(times[1:] or times) // 2]
    value = n / t
    line = {'impl': 'reference', 'metric': 'SR train patches/s (fwd+bwd)', 'value': value, 'unit': 'patches/s',
            'n_gpus': args.gpus, 'steps': len(times[1:] or times), 'warmup': 1, 'ms_per_step': t * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(CONFIG, parallelism='host threads'),
            'cpu_baseline': {'value': value, 'unit': 'patches/s', 'cores': threads, 'kind': 'port',
                             'sample': f'{n} of 16 patches per step, EDSR-L x4 fwd+L1+bwd, oracle/sr_oracle.py fp32'},
            'e2e': {'value': value, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), file=JSON_OUT, flush=True)


JSON_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='srb200', choices=['srb200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of CUDA-graph replay')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    # stdout carries exactly ONE line, the JSON record: everything else a library may print there (NCCL prints its
    # version banner to stdout when NCCL_DEBUG=VERSION|WARN) is routed to stderr at the file-descriptor level
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)

    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from torch import nn
    from basicsr4rs_b200 import _lib as L
    from basicsr4rs_b200.archs import build_network
    from basicsr4rs_b200.ops.sr_b200 import raw

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    L.check(L.load().srb200_check_device(local_rank), 'srb200_check_device')

    torch.manual_seed(0)
    net = build_network(dict(EDSR_L, cuda_graph=not args.no_graph, graph_segments=4,
                             graph_input_shape=[BATCH, 3, LR, LR])).to(dev)  # graphs captured here, before DDP
    model = nn.parallel.DistributedDataParallel(net, device_ids=[local_rank], gradient_as_bucket_view=True) \
        if world > 1 else net
    optim = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.99), fused=True)
    crit = nn.L1Loss()

    lq_h, gt_h = (t.pin_memory() for t in synthetic_batch(rank))
    lq_d, gt_d = lq_h.to(dev), gt_h.to(dev)

    def train_step(lq, gt):
        optim.zero_grad(set_to_none=True)
        loss = crit(model(lq), gt)
        loss.backward()
        optim.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev, world)

    for _ in range(args.warmup):
        train_step(lq_d, gt_d)

    # ---- device-resident leg (`value`)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    n0 = L.launch_count
    ms = timed(lambda: train_step(lq_d, gt_d), args.steps)
    launches = L.launch_count - n0  # eager C-ABI calls (0 when the step is replayed from CUDA graphs)
    if net.cuda_graph:
        from basicsr4rs_b200.archs.graphed import GRAPHS
        launches += args.steps * sum(GRAPHS[net].kernels_per_step.values())
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end leg: pinned host inputs -> H2D, public nn.Module API, loss read back
    def e2e_step():
        lq = lq_h.to(dev, non_blocking=True)
        gt = gt_h.to(dev, non_blocking=True)
        return train_step(lq, gt).item()

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    # ---- roofline leg: the dominant kernel timed with CUDA events on its launching stream, inside real
    # train steps.  Events cannot be recorded inside a graph replay, so these steps launch eagerly.
    probe = raw.EventProbe(lambda kind, info: kind == 'tapgemm' and info == (BATCH, LR, LR, 256, 256, 3))
    graph_flag, net.cuda_graph = net.cuda_graph, False
    train_step(lq_d, gt_d)
    raw.PROBE = probe
    for _ in range(min(args.steps, 5)):
        # an eager step is CPU-bound (one Python call per launch): queue it behind a device-side delay so that the
        # kernels run back to back as they do in the graph replay -- otherwise every event pair also times the idle gap
        # between the start event and a kernel the host has not launched yet
        torch.cuda._sleep(int(0.08 * 1.9e9))
        train_step(lq_d, gt_d)
    raw.PROBE = None
    kernel_ms = probe.ms()
    net.cuda_graph = graph_flag

    if rank == 0:
        pk, pk_kind = peaks()
        patches = world * BATCH * args.steps
        value = patches / (ms / 1e3)
        kflop = 2.0 * BATCH * LR * LR * 256 * 256 * 9
        k_avg = sum(kernel_ms) / max(1, len(kernel_ms))
        achieved = kflop / (k_avg / 1e3) / 1e12 if k_avg > 0 else None
        peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
        line = {
            'metric': 'SR train patches/s (fwd+bwd)', 'value': value, 'unit': 'patches/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic', 'config': CONFIG,
            'e2e': {'value': patches / (ms_e2e / 1e3), 'unit': 'patches/s',
                    'h2d_bytes_per_step': (lq_h.numel() + gt_h.numel()) * 4, 'd2h_bytes_per_step': 4},
            'gpu_launches': launches, 'clocks': clocks,
            'model_tflops': value * FLOP_PER_PATCH_FWD_BWD / world / 1e12,
            'roofline': {'bound': 'tensor', 'kernel': 'tapgemm_kernel<256,true> (cta_group::2) conv3x3 256->256, fprop + dgrad launches',
                         'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': achieved / peak if achieved else None, 'traffic': ROOFLINE_TRAFFIC_BYTES,
                         'traffic_unit': 'bytes/launch (ncu --set full, profiles/r01_ncu_full_tapgemm_v2.txt)',
                         'peak_kind': f'{pk_kind} bf16 sustained (timed inside a long step)',
                         'launches_timed': len(kernel_ms), 'avg_ms': k_avg},
        }
        if not args.no_cpu_baseline and world == 1:  # reported at N=1 only (torchrun pins OMP threads to 1)
            threads = os.cpu_count() or 1
            t = cpu_oracle_step_time(4, threads, iters=2)
            line['cpu_baseline'] = {'value': 4.0 / t, 'unit': 'patches/s', 'cores': threads, 'kind': 'port',
                                    'sample': '4 of 16 patches, EDSR-L x4 fwd+L1+bwd, oracle/sr_oracle.py fp32, best of 3'}
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
