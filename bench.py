#!/usr/bin/env python
"""Headline benchmark: SR train patches/s of EDSR-L x4 (BASELINE.json configs[1]) on N B200s, plus -- at N=1 -- every
other BASELINE config under ``configs`` in the same JSON line.

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                          # the reference's own archs on the host CPU

One "step" = zero_grad + forward + L1 loss + backward + Adam step of EDSR-L x4 (32 ResBlocks, 256 ch,
res_scale 0.1) on 16 synthetic 3x48x48 LR patches per GPU (weak scaling, DDP over NCCL).  Prints ONE
JSON line (see the task contract): `value` = device-resident throughput, `e2e` = the same through the
public nn.Module API with pinned-host inputs copied H2D and the loss read back D2H every step (prefetched on a copy
stream / read back asynchronously, like the reference's CUDAPrefetcher -- class PrefetchedSteps),
`roofline` = the dominant kernel (conv3x3 256->256 tap-GEMM) timed live with CUDA events,
`cpu_baseline` = the unmodified reference arch (oracle/_ref, vendored by oracle/make_ref.py; the oracle port when
absent) on a bounded sample of the same workload, `configs` = EDSR-M / RCAN / SwinIR B4+B16 training and EDSR-L /
SwinIR 1024x1024-tile inference, each with value, ms and the rooflines of its own dominant kernels.
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
EDSR_M = dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=64, num_block=16, upscale=4, res_scale=1, img_range=255.)
RCAN = dict(type='RCAN', num_in_ch=3, num_out_ch=3, num_feat=64, num_group=10, num_block=20, squeeze_factor=16,
            upscale=4, res_scale=1, img_range=255.)
SWINIR = dict(type='SwinIR', upscale=4, in_chans=3, img_size=64, window_size=8, img_range=1., depths=[6] * 6,
              embed_dim=180, num_heads=[6] * 6, mlp_ratio=2, upsampler='pixelshuffle', resi_connection='1conv')
BATCH, LR = 16, 48
WORKLOAD = 'EDSR-L x4 train step (fwd + L1 + bwd + Adam), 16x3x48x48 LR patches per GPU'
# `config` names the WORKLOAD only and is identical in both arms (the driver compares them); how each arm runs it is
# under `launch`
CONFIG = {'workload': WORKLOAD, 'arch': 'EDSR num_feat=256 num_block=32 res_scale=0.1 upscale=4', 'batch_per_gpu': BATCH,
          'lr_patch': LR, 'l2': 'per-step working set (>2 GB of activations) exceeds the 126 MB L2'}
FLOP_PER_PATCH_FWD_BWD = 694.66e9
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, mean of the four 256->256 launches in
# profiles/r01_ncu_full_tapgemm_v2.txt (one `ncu --set full` capture of this script): 39.09 / 20.14 / 40.10 / 20.14 MB.  The
# algorithmic bytes are 18.9 MB in + 18.9 MB out + 1.2 MB weights (+18.9 MB residual for conv2 / dgrad-conv1); the
# output is still L2-resident when the kernel ends, so DRAM sees only the compulsory reads -- no wasted re-reads.
ROOFLINE_TRAFFIC_BYTES = 29.87e6

# name: (network opt, batch, LR patch, GFLOP fwd+bwd per patch [BASELINE.md section 2], CUDA-graph segments)
TRAIN_CONFIGS = {
    'edsr_m': (EDSR_M, 16, 48, 27.41, 4),
    'rcan': (RCAN, 16, 48, 220.04, 5),
    'swinir_b4': (SWINIR, 4, 64, 321.30, 6),
    'swinir_b16': (SWINIR, 16, 64, 321.30, 6),
}
# name: (network opt, LR tile, GFLOP fwd per LR pixel)
INFER_CONFIGS = {
    'edsr_l_infer_1024': (EDSR_L, 1024, 100.5e-3),
    'swinir_infer_1024': (SWINIR, 1024, 26.15e-3),
}
LOGICAL = {192: 180, 576: 540, 384: 360}  # SwinIR's padded -> algorithmic channel counts (padding is not credited)


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


# ------------------------------------------------------------------ the reference on the host CPU
def cpu_reference_step(n_patches, threads):
    """Returns (step_fn, kind): ``step_fn()`` runs fwd + L1 + bwd of EDSR-L x4 on ``n_patches`` synthetic 48x48 patches
    on the host CPU in fp32 and returns the seconds it took.  kind = 'reference': the UNMODIFIED reference
    ``basicsr.archs.edsr_arch.EDSR`` (from /root/reference here, from the verbatim copies in oracle/_ref on the GPU
    box); 'port': oracle/sr_oracle.py when neither exists."""
    from oracle import ref_shim
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1234)
    lq = torch.rand((n_patches, 3, LR, LR), generator=g)
    gt = torch.rand((n_patches, 3, 4 * LR, 4 * LR), generator=g)
    kw = {k: v for k, v in EDSR_L.items() if k != 'type'}
    if ref_shim.available():
        net = ref_shim.load_reference_archs().EDSR(**kw).train()
        crit = torch.nn.L1Loss()

        def step():
            net.zero_grad(set_to_none=True)
            t0 = time.perf_counter()
            crit(net(lq), gt).backward()
            return time.perf_counter() - t0
        return step, 'reference'
    from oracle import sr_oracle
    from basicsr4rs_b200.archs import build_network
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in build_network(EDSR_L).state_dict().items()}

    def step():
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        out = sr_oracle.edsr_forward(sd, lq, num_block=32, upscale=4, res_scale=0.1, img_range=255.)
        (out - gt).abs().mean().backward()
        return time.perf_counter() - t0
    return step, 'port'


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, same metric / config;
    each step = a bounded sample (2 of the 16 patches) so that K steps end within a few minutes."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 2
    step, kind = cpu_reference_step(n, threads)
    for _ in range(args.warmup):
        step()
    times = []
    for _ in range(args.steps):
        times.append(step())
        if sum(times) > 240:
            break
    t = sum(times) / len(times)
    value = n / t
    what = 'basicsr.archs.edsr_arch.EDSR of the reference, verbatim (oracle/_ref), torch fp32 CPU' if kind == 'reference' \
        else 'oracle/sr_oracle.py fp32 (reference tree absent)'
    line = {'impl': 'reference', 'metric': 'SR train patches/s (fwd+bwd)', 'value': value, 'unit': 'patches/s',
            'n_gpus': args.gpus, 'steps': len(times), 'warmup': args.warmup, 'ms_per_step': t * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(CONFIG),
            'launch': {'how': f'eager torch CPU ops, {threads} host threads; fwd + L1 + bwd of a {n}-patch sample per step',
                       'parallelism': 'host threads (rank 0 only)'},
            'cpu_baseline': {'value': value, 'unit': 'patches/s', 'cores': threads, 'kind': kind,
                             'sample': f'{n} of 16 patches per step, EDSR-L x4 fwd+L1+bwd, {what}'},
            'e2e': {'value': value, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), file=JSON_OUT, flush=True)


JSON_OUT = sys.stdout


# ------------------------------------------------------------------ rooflines from probed launches
def summarize_rooflines(records, pk, n_steps=1):
    """records: (kind, info, ms) of probed launches -> {kernel family: roofline dict}.  GEMM-shaped kernels by their
    algorithmic FLOP against the measured bf16 burst peak (sustained beside it), the rest by algorithmic bytes against
    the measured HBM copy bandwidth (DESIGN.md section 3 states the per-unit figures)."""
    fam = {}

    def add(name, bound, work, ms):
        f = fam.setdefault(name, {'bound': bound, 'work': 0.0, 'ms': 0.0, 'n': 0})
        f['work'] += work
        f['ms'] += ms
        f['n'] += 1

    for kind, info, ms in records:
        if kind == 'tapgemm':
            b, h, w, cin, cout, ks = info
            flop = 2.0 * b * h * w * LOGICAL.get(cin, cin) * LOGICAL.get(cout, cout) * ks * ks
            add(f'tapgemm {"conv3x3" if ks == 3 else "linear"} {cin}->{cout}', 'tensor', flop, ms)
        elif kind == 'wgrad':
            b, h, w, n, k, ks = info
            add(f'wgrad {"conv3x3" if ks == 3 else "linear"} {k}->{n}', 'tensor',
                2.0 * b * h * w * LOGICAL.get(n, n) * LOGICAL.get(k, k) * ks * ks, ms)
        elif kind in ('window_attn_fwd', 'window_attn_bwd'):
            b, h, w, heads, ws = info
            c = heads * 30 if heads == 6 else heads * 32  # un-padded embed dim of the classical config (head_dim 30)
            t = b * h * w
            # fwd: read q,k,v + write o = 4 T C e; bwd: read q,k,v,dO + write dq,dk,dv = 7 T C e  (e = 2 bytes)
            add(kind, 'hbm', (4 if kind.endswith('fwd') else 7) * t * c * 2.0, ms)
        elif kind == 'layernorm_fwd':
            t, c, cp = info
            add(kind, 'hbm', 2.0 * t * c * 2, ms)
        elif kind == 'layernorm_bwd':
            t, c, cp, gres = info
            add(kind, 'hbm', (4.0 if gres else 3.0) * t * c * 2, ms)
        elif kind == 'ca_forward':
            b, hw, c, f32 = info
            add(kind, 'hbm', b * hw * c * (2 + 4 + 2 + 4 if f32 else 6.0), ms)  # t, skip in, y out (+ fp32 twins)
        elif kind == 'ca_backward':
            b, hw, c = info
            add(kind, 'hbm', 3.0 * b * hw * c * 2, ms)  # read g, t; write gt (g's second read comes from L2)
    out = {}
    for name, f in fam.items():
        if f['ms'] <= 0:
            continue
        if f['bound'] == 'tensor':
            ach = f['work'] / (f['ms'] / 1e3) / 1e12
            out[name] = {'bound': 'tensor', 'achieved': ach, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s',
                         'frac': ach / pk['bf16_tflops'],
                         'frac_of_sustained': ach / pk.get('bf16_tflops_sustained', pk['bf16_tflops']),
                         'launches_timed': f['n'], 'avg_ms': f['ms'] / f['n'], 'step_ms': f['ms'] / n_steps}
        else:
            ach = f['work'] / (f['ms'] / 1e3) / 1e9
            out[name] = {'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                         'frac': ach / pk['hbm_gbs'], 'launches_timed': f['n'], 'avg_ms': f['ms'] / f['n'],
                         'step_ms': f['ms'] / n_steps}
    return out


def probe_eager_steps(net, step_fn, n_steps, raw):
    """Per-kernel CUDA-event times inside real EAGER train steps queued behind a device-side delay (so the kernels run
    back to back as in the graph replay and no event pair times a host launch gap).  Returns the probe's records."""
    probe = raw.EventProbe(lambda kind, info: True, limit=200000)
    graph_flag, net.cuda_graph = getattr(net, 'cuda_graph', False), False
    step_fn()
    raw.PROBE = probe
    for _ in range(n_steps):
        torch.cuda._sleep(int(0.08 * 1.9e9))
        step_fn()
    raw.PROBE = None
    recs = probe.records()
    net.cuda_graph = graph_flag
    return recs


def dominant(rooflines):
    """The family with the largest share of the step (what `roofline` of a config names); others go under `kernels`."""
    if not rooflines:
        return None, {}
    name = max(rooflines, key=lambda k: rooflines[k]['step_ms'])
    return dict(rooflines[name], kernel=name), {k: v for k, v in rooflines.items() if k != name}


class PrefetchedSteps:
    """End-to-end loop the way the reference feeds its models (CUDAPrefetcher, basicsr/data/prefetch_dataloader.py:
    82-122: the next batch is uploaded on a side stream while the current step computes): every step's LR / GT batch
    is copied from PINNED HOST memory to one of two device buffers on a copy stream, and every step's loss is read back
    to pinned host memory -- asynchronously, the host looks at the values after the last step (what
    utils/train_hooks.install_lazy_loss_log does to BaseModel.reduce_loss_dict).  All copies are inside the timed region."""

    def __init__(self, lq_h, gt_h, dev, step_fn):
        self.lq_h, self.gt_h, self.dev, self.step_fn = lq_h, gt_h, dev, step_fn
        self.copy = torch.cuda.Stream(device=dev)
        self.bufs = [(torch.empty(lq_h.shape, device=dev), torch.empty(gt_h.shape, device=dev)) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]

    def _prefetch(self, k):
        i = k & 1
        with torch.cuda.stream(self.copy):
            if k >= 2:
                self.copy.wait_event(self.consumed[i])  # step k - 2 has read this buffer
            self.bufs[i][0].copy_(self.lq_h, non_blocking=True)
            self.bufs[i][1].copy_(self.gt_h, non_blocking=True)
            self.ready[i].record(self.copy)

    def run(self, steps):
        main = torch.cuda.current_stream(self.dev)
        losses = torch.empty((steps,), dtype=torch.float32).pin_memory()
        self._prefetch(0)
        for k in range(steps):
            i = k & 1
            main.wait_event(self.ready[i])
            if k + 1 < steps:
                self._prefetch(k + 1)
            loss = self.step_fn(*self.bufs[i])
            self.consumed[i].record(main)
            losses[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        return losses


def make_adam(params, kind='srb200'):
    if kind == 'torch':
        return torch.optim.Adam(params, lr=1e-4, betas=(0.9, 0.99), fused=True)
    from basicsr4rs_b200.utils.fused_adam import FusedAdam
    return FusedAdam(params, lr=1e-4, betas=(0.9, 0.99))


def bench_train_config(name, dev, steps, pk, raw, build_network, L, optim_kind='srb200'):
    opt, batch, lr, gflop, segs = TRAIN_CONFIGS[name]
    torch.manual_seed(0)
    net = build_network(dict(opt, cuda_graph=True, graph_segments=segs)).to(dev).train()
    optim = make_adam(net.parameters(), optim_kind)
    crit = torch.nn.L1Loss()
    lq_h, gt_h = (t.pin_memory() for t in synthetic_batch(0, batch, lr))
    lq, gt = lq_h.to(dev), gt_h.to(dev)

    def step(a=lq, b=gt):
        optim.zero_grad(set_to_none=True)
        loss = crit(net(a), b)
        loss.backward()
        optim.step()
        return loss

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps

    pipe = PrefetchedSteps(lq_h, gt_h, dev, step)
    pipe.run(2)
    torch.cuda.synchronize()
    e0.record()
    losses = pipe.run(steps)
    e1.record()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(losses).all())
    ms_e2e = e0.elapsed_time(e1) / steps
    roof, others = dominant(summarize_rooflines(probe_eager_steps(net, step, 2, raw), pk, n_steps=2))
    row = {'workload': f'{opt["type"]} x4 train step (fwd + L1 + bwd + Adam), {batch}x3x{lr}x{lr} LR patches, '
                       f'CUDA-graph replay, training mode' + (' (drop_path_rate 0.1)' if opt['type'] == 'SwinIR' else ''),
           'value': batch / ms * 1e3, 'unit': 'patches/s', 'ms_per_step': ms,
           'e2e': {'value': batch / ms_e2e * 1e3, 'unit': 'patches/s',
                   'h2d_bytes_per_step': (lq_h.numel() + gt_h.numel()) * 4, 'd2h_bytes_per_step': 4},
           'model_tflops': batch * gflop / ms, 'roofline': roof, 'kernels': others,
           'mem_gb': torch.cuda.max_memory_allocated() / 2**30}
    del net, optim
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    return row


def bench_infer_config(name, dev, steps, pk, raw, build_network):
    """BASELINE config 5 on one GPU: x4 inference of 1024x1024 LR tiles.  `value` = fp32 tile resident in HBM, network
    only; `e2e` = through the scene driver (basicsr4rs_b200.utils.tiling.TiledUpscaler): the uint8 LR tile starts in
    pinned host memory, the uint8 4096x4096 result ends in pinned host memory (entry / exit kernels, H2D, D2H timed)."""
    from basicsr4rs_b200.utils.tiling import TiledUpscaler
    opt, tile, gflop_px = INFER_CONFIGS[name]
    torch.manual_seed(0)
    net = build_network(dict(opt)).to(dev).eval()
    g = torch.Generator().manual_seed(1234)
    scene = torch.randint(0, 256, (tile, tile, 3), generator=g, dtype=torch.uint8).pin_memory()
    y_h = torch.empty((4 * tile, 4 * tile, 3), dtype=torch.uint8).pin_memory()
    n = max(2, steps // 4)
    with torch.no_grad():
        x = raw.patch_from_u8([scene.to(dev)], [0], [0], [0], tile, tile)
        for _ in range(2):
            net(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            net(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        up = TiledUpscaler(net, scale=4, tile=tile, multiple=8 if opt['type'] == 'SwinIR' else 1)
        _, _, _, h2d, d2h = up.upscale(scene, out=y_h)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            up.upscale(scene, out=y_h)
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1) / n
        probe = raw.EventProbe(lambda kind, info: True, limit=200000)
        raw.PROBE = probe
        torch.cuda._sleep(int(0.05 * 1.9e9))
        net(x)
        raw.PROBE = None
        roof, others = dominant(summarize_rooflines(probe.records(), pk))
    mpix = (4 * tile)**2 / 1e6
    row = {'workload': f'{opt["type"]} x4 inference, one {tile}x{tile} LR tile -> {4 * tile}x{4 * tile}, eval / no_grad',
           'value': mpix / ms * 1e3, 'unit': 'output MPix/s', 'ms_per_tile': ms,
           'e2e': {'value': mpix / ms_e2e * 1e3, 'unit': 'output MPix/s', 'h2d_bytes_per_step': h2d,
                   'd2h_bytes_per_step': d2h, 'path': 'TiledUpscaler: uint8 pinned host -> uint8 pinned host'},
           'model_tflops': gflop_px * tile * tile / ms, 'roofline': roof, 'kernels': others,
           'mem_gb': torch.cuda.max_memory_allocated() / 2**30}
    del net
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='srb200', choices=['srb200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of CUDA-graph replay')
    ap.add_argument('--grad-comm', default='fp32', choices=['fp32', 'bf16'],
                    help="DDP gradient all-reduce payload: fp32 (the reference's) or torch's bf16_compress_hook")
    ap.add_argument('--ddp', default='flat', choices=['flat', 'torch'],
                    help="N > 1: 'flat' = basicsr4rs_b200.utils.flat_ddp.FlatDDP (gradients written straight into one flat "
                         "buffer that NCCL all-reduces, no per-parameter bucket copies), 'torch' = DistributedDataParallel "
                         "exactly as BaseModel.model_to_device wraps it")
    ap.add_argument('--optim', default='srb200', choices=['srb200', 'torch'],
                    help="Adam step: 'srb200' = one srb200_multi_adam launch (utils/fused_adam.FusedAdam), 'torch' = "
                         "torch.optim.Adam(fused=True)")
    ap.add_argument('--segments', type=int, default=4, help='CUDA-graph segments of the headline net (the gradients of the '
                    'LAST backward segment are all-reduced with nothing left to overlap; 8 measured no better at N=4)')
    ap.add_argument('--configs', default='all', help="'all', 'none' or a comma list of the other BASELINE configs "
                    f"({', '.join(list(TRAIN_CONFIGS) + list(INFER_CONFIGS))}); measured at N=1 only")
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
    flat = world > 1 and args.ddp == 'flat'
    segments = max(1, args.segments)
    net = build_network(dict(EDSR_L, cuda_graph=not args.no_graph, graph_segments=segments, flat_grads=flat,
                             graph_input_shape=[BATCH, 3, LR, LR])).to(dev)  # graphs captured here, before DDP
    if flat:
        from basicsr4rs_b200.utils.flat_ddp import FlatDDP
        model = FlatDDP(net)
    else:
        model = nn.parallel.DistributedDataParallel(net, device_ids=[local_rank], gradient_as_bucket_view=True) \
            if world > 1 else net
    if world > 1 and not flat and args.grad_comm == 'bf16':
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        model.register_comm_hook(None, default_hooks.bf16_compress_hook)
    optim = make_adam(model.parameters(), args.optim)
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

    # ---- end-to-end leg: pinned host inputs -> H2D every step, public nn.Module API, every loss read back to the host
    pipe = PrefetchedSteps(lq_h, gt_h, dev, train_step)
    pipe.run(2)
    e2e_losses = []
    ms_e2e = timed(lambda: e2e_losses.append(pipe.run(args.steps)), 1)
    assert bool(torch.isfinite(e2e_losses[0]).all()), 'e2e: a loss read back from the device is not finite'

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
        peak = pk['bf16_tflops']  # burst: the probed steps are five ~10 ms bursts, each behind an 80 ms idle delay
        sustained = pk.get('bf16_tflops_sustained', peak)
        line = {
            'metric': 'SR train patches/s (fwd+bwd)', 'value': value, 'unit': 'patches/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': dict(CONFIG),
            'launch': {'how': ('eager launches' if args.no_graph else f'CUDA-graph replay ({segments} segments)') + ' of fwd+bwd; L1 loss + '
                              + ('one-launch Adam (utils/fused_adam.FusedAdam)' if args.optim == 'srb200'
                                 else 'torch.optim.Adam(fused=True)') + ' eager',
                       'parallelism': 'one process per GPU, weak scaling' if world > 1 else 'single GPU',
                       'grad_comm': (('fp32 NCCL all-reduce (SUM of pre-divided gradients) of one flat gradient buffer, '
                                      'utils/flat_ddp.FlatDDP' if flat
                                      else args.grad_comm + ' buckets, torch DistributedDataParallel')
                                     if world > 1 else 'none (1 GPU)')},
            'e2e': {'value': patches / (ms_e2e / 1e3), 'unit': 'patches/s',
                    'h2d_bytes_per_step': (lq_h.numel() + gt_h.numel()) * 4, 'd2h_bytes_per_step': 4,
                    'path': 'nn.Module API; every step: LR+GT batch pinned host -> device on a copy stream (double-buffered, '
                            'as the reference CUDAPrefetcher does), loss -> pinned host; all copies inside the timed region'},
            'gpu_launches': launches, 'clocks': clocks,
            'model_tflops': value * FLOP_PER_PATCH_FWD_BWD / world / 1e12,
            'roofline': {'bound': 'tensor', 'kernel': 'tapgemm_kernel<256,true> (cta_group::2) conv3x3 256->256, fprop + dgrad launches',
                         'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': achieved / peak if achieved else None,
                         'frac_of_sustained': achieved / sustained if achieved else None,
                         'traffic': ROOFLINE_TRAFFIC_BYTES,
                         'traffic_unit': 'bytes/launch (ncu --set full, profiles/r01_ncu_full_tapgemm_v2.txt)',
                         'peak_kind': f'{pk_kind} bf16 burst (cuBLAS 8192^3 best of 10); sustained {sustained} beside it',
                         'launches_timed': len(kernel_ms), 'avg_ms': k_avg},
        }
        if not args.no_cpu_baseline and world == 1:  # reported at N=1 only (torchrun pins OMP threads to 1)
            threads = os.cpu_count() or 1
            n = 4
            step, kind = cpu_reference_step(n, threads)
            step()  # warm-up
            t = min(step() for _ in range(2))
            line['cpu_baseline'] = {'value': n / t, 'unit': 'patches/s', 'cores': threads, 'kind': kind,
                                    'sample': f'{n} of 16 patches, EDSR-L x4 fwd+L1+bwd, fp32, best of 2 ('
                                              + ('unmodified reference arch, oracle/_ref' if kind == 'reference'
                                                 else 'oracle/sr_oracle.py port') + ')'}
        # ---- every other BASELINE config (N=1 only; the scaling runs keep to the headline workload)
        want = [] if args.configs == 'none' else (list(TRAIN_CONFIGS) + list(INFER_CONFIGS) if args.configs == 'all'
                                                  else [c for c in args.configs.split(',') if c])
        if world == 1 and want:
            del model, optim, net
            torch.cuda.empty_cache()
            cfgs = {}
            for name in want:
                try:
                    if name in TRAIN_CONFIGS:
                        cfgs[name] = bench_train_config(name, dev, max(5, args.steps // 2), pk, raw, build_network, L, args.optim)
                    elif name in INFER_CONFIGS:
                        cfgs[name] = bench_infer_config(name, dev, args.steps, pk, raw, build_network)
                except Exception as e:  # a failing side config must not take the headline line with it
                    cfgs[name] = {'error': f'{type(e).__name__}: {e}'[:300]}
            line['configs'] = cfgs
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
