"""Throughput of every BASELINE.json config on one GPU (training patches/s with CUDA-graph replay, inference
output MPix/s), next to the algorithmic FLOP rate.  Not the driver's bench (that is bench.py = config 2).

    python tools/bench_all.py [--steps 10] [--only edsr_l,swinir_b16,...] > gpurun_out/bench_all.jsonl
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200 import _lib as L  # noqa: E402
from basicsr4rs_b200.archs import build_network  # noqa: E402

SWIN = dict(type='SwinIR', upscale=4, in_chans=3, img_size=64, window_size=8, img_range=1., depths=[6] * 6,
            embed_dim=180, num_heads=[6] * 6, mlp_ratio=2, upsampler='pixelshuffle', resi_connection='1conv')
TRAIN = {
    # name: (opt, batch, lr_size, GFLOP fwd+bwd per patch [BASELINE.md section 2])
    'edsr_m': (dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=64, num_block=16, upscale=4, res_scale=1), 16, 48,
               27.41),
    'edsr_l': (dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=256, num_block=32, upscale=4, res_scale=0.1), 16,
               48, 694.66),
    'rcan': (dict(type='RCAN', num_in_ch=3, num_out_ch=3, num_feat=64, num_group=10, num_block=20, squeeze_factor=16,
                  upscale=4, res_scale=1), 16, 48, 220.04),
    'swinir_b4': (SWIN, 4, 64, 321.30),
    'swinir_b16': (SWIN, 16, 64, 321.30),
}
INFER = {
    # name: (opt, tile, GFLOP fwd per LR pixel)
    'edsr_l_infer_1024': (TRAIN['edsr_l'][0], 1024, 100.5e-3),
    'swinir_infer_1024': (SWIN, 1024, 26.15e-3),
}


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--only', default='')
    ap.add_argument('--no-graph', action='store_true')
    args = ap.parse_args()
    only = set(filter(None, args.only.split(',')))
    dev = torch.device('cuda:0')
    for name, (opt, batch, lr, gflop) in TRAIN.items():
        if only and name not in only:
            continue
        torch.manual_seed(0)
        net = build_network(dict(opt, cuda_graph=not args.no_graph)).to(dev).train()
        optim = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.99), fused=True)
        lq = torch.rand((batch, 3, lr, lr), device=dev)
        gt = torch.rand((batch, 3, 4 * lr, 4 * lr), device=dev)

        def step():
            optim.zero_grad(set_to_none=True)
            loss = (net(lq) - gt).abs().mean()
            loss.backward()
            optim.step()

        t0 = time.time()
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        setup = time.time() - t0
        n0 = L.launch_count
        ms = timed(step, args.steps)
        row = dict(config=name, mode='train', batch=batch, ms_per_step=ms, patches_per_s=batch / ms * 1e3,
                   model_tflops=batch * gflop / ms, setup_s=setup, eager_launches_per_step=(L.launch_count - n0) / args.steps,
                   mem_gb=torch.cuda.max_memory_allocated() / 2**30)
        print(json.dumps(row), flush=True)
        del net, optim
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    for name, (opt, tile, gflop_px) in INFER.items():
        if only and name not in only:
            continue
        torch.manual_seed(0)
        net = build_network(opt).to(dev).eval()
        x = torch.rand((1, 3, tile, tile), device=dev)
        with torch.no_grad():
            for _ in range(2):
                net(x)
            ms = timed(lambda: net(x), max(2, args.steps // 2))
        row = dict(config=name, mode='infer', tile=tile, ms_per_tile=ms, out_mpix_per_s=(4 * tile)**2 / ms / 1e3,
                   model_tflops=gflop_px * tile * tile / ms, mem_gb=torch.cuda.max_memory_allocated() / 2**30)
        print(json.dumps(row), flush=True)
        del net
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == '__main__':
    main()
