"""Per-kernel timings at the BASELINE.json layer shapes (CUDA events, L2 flushed between iterations).

Usage (GPU box):  python tools/bench_kernels.py [--iters 20] > gpurun_out/kernels.json
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200 import _lib as L  # noqa: E402
from basicsr4rs_b200.ops.sr_b200 import raw  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []

    def conv_case(name, b, h, w, cin, cout, ks=3):
        x = torch.randn((b, h, w, cin), device=dev).to(torch.bfloat16)
        wt = torch.randn((cout, cin, ks, ks), device=dev) * 0.02
        wp = raw.pack_weight(wt, cout, cin)
        wpt = raw.pack_weight(wt, cout, cin, transpose=True)
        bias = torch.zeros(cout, device=dev)
        dy = torch.randn((b, h, w, cout), device=dev).to(torch.bfloat16)
        flops = 2.0 * b * h * w * cin * cout * ks * ks
        for kind, fn in (
            ('fprop', lambda: raw.tapgemm(x, wp, ksize=ks, cout=cout, bias=bias, act=L.ACT_RELU)),
            ('dgrad', lambda: raw.tapgemm(dy, wpt, ksize=ks, cout=cin, flip=True)),
            ('wgrad', lambda: raw.wgrad(dy, x, ksize=ks)),
        ):
            med, best = timeit(fn, args.iters, flush)
            rows.append(dict(kernel=f'{name}.{kind}', ms_median=med, ms_best=best, tflops_median=flops / med / 1e9,
                             tflops_best=flops / best / 1e9))
            print(json.dumps(rows[-1]), flush=True)

    conv_case('edsrL_body_256x256_48', 16, 48, 48, 256, 256)
    conv_case('edsrL_up1_256x1024_48', 16, 48, 48, 256, 1024)
    conv_case('edsrL_up2_256x1024_96', 16, 96, 96, 256, 1024)
    conv_case('rcan_body_64x64_48', 16, 48, 48, 64, 64)
    conv_case('swinir_conv_192x192_64', 16, 64, 64, 192, 192)
    conv_case('swinir_qkv_192x576_64', 16, 64, 64, 192, 576, ks=1)
    conv_case('swinir_fc1_192x384_64', 16, 64, 64, 192, 384, ks=1)

    # HBM-bound remaps
    x = torch.randn((16, 192, 192, 256), device=dev).to(torch.bfloat16)
    med, best = timeit(lambda: raw.pixel_shuffle_nhwc(x, 2, inverse=True), args.iters, flush)
    nbytes = 2 * x.numel() * 2
    rows.append(dict(kernel='pixel_unshuffle_nhwc_16x192x192x256', ms_median=med, gbs_median=nbytes / med / 1e6))
    print(json.dumps(rows[-1]), flush=True)
    t = torch.randn((16, 64, 64, 192), device=dev).to(torch.bfloat16)
    med, best = timeit(lambda: raw.window_partition(t, 8, 4), args.iters, flush)
    rows.append(dict(kernel='window_partition_shift_16x64x64x192', ms_median=med,
                     gbs_median=2 * t.numel() * 2 / med / 1e6))
    print(json.dumps(rows[-1]), flush=True)


if __name__ == '__main__':
    main()
