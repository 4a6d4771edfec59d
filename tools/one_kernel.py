"""ncu / microbenchmark target: a few launches of ONE hot-path kernel at its BASELINE shape.

    python tools/one_kernel.py NAME [--time]

NAME: attn_fwd attn_bwd ln_fwd ln_bwd qkv proj fc1_gelu fc1_gelu_aux fc2 fc2_dgrad wgrad_fc1 (SwinIR, B16 x 64x64 tokens)
      qkv_fold_1m fc2_stats_1m (SwinIR evaluation with folded LayerNorm, one 1024 x 1024 tile)
      conv64 conv64_dgrad wgrad64 ca_fwd ca_bwd (RCAN, 16 x 48x48 x 64)   conv256 wgrad256 (EDSR-L, 16 x 48x48 x 256)
--time: print the device time of back-to-back launches queued behind a device-side delay instead of just launching."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200 import _lib as L  # noqa: E402
from basicsr4rs_b200.ops.sr_b200 import raw, swin_ops as so  # noqa: E402

dev = torch.device('cuda:0')


def bf(*shape):
    return torch.randn(shape, device=dev).to(torch.bfloat16)


def swin(name):
    B, H, W = 16, 64, 64
    x192, x384, x576 = bf(B, H, W, 192), bf(B, H, W, 384), bf(B, H, W, 576)
    x192[..., 180:] = 0

    def lin(n, k):
        w = torch.randn(n, k, device=dev) * 0.05
        return raw.pack_weight(w, n, k), torch.zeros(n, device=dev)

    table = torch.randn(225, 6, device=dev) * 0.1
    g, b_ = torch.ones(180, device=dev), torch.zeros(180, device=dev)
    if name == 'attn_fwd':
        return lambda: so.window_attention_fwd(x576, table, 6, 8, 4, 30**-0.5, want_stats=True)
    if name == 'attn_bwd':
        _, stats = so.window_attention_fwd(x576, table, 6, 8, 4, 30**-0.5, want_stats=True)
        return lambda: so.window_attention_bwd(x576, x192, table, 6, 8, 4, 30**-0.5, stats=stats)
    if name == 'ln_fwd':
        return lambda: so.layernorm_fwd(x192, g, b_, 180)
    if name == 'ln_fwd_1m':  # one 1024 x 1024 inference tile
        big = bf(1, 1024, 1024, 192)
        return lambda: so.layernorm_fwd(big, g, b_, 180)
    if name == 'ln_bwd':
        _, mean, rstd = so.layernorm_fwd(x192, g, b_, 180)
        return lambda: so.layernorm_bwd(x192, x192, mean, rstd, g, 180, gres=x192)
    if name == 'qkv':
        wp, bias = lin(576, 192)
        return lambda: raw.tapgemm(x192, wp, ksize=1, cout=576, bias=bias)
    if name == 'proj':
        wp, bias = lin(192, 192)
        return lambda: raw.tapgemm(x192, wp, ksize=1, cout=192, bias=bias, residual=x192)
    if name == 'fc1_gelu':
        wp, bias = lin(384, 192)
        return lambda: raw.tapgemm(x192, wp, ksize=1, cout=384, bias=bias, act=L.ACT_GELU)
    if name == 'fc1_gelu_aux':
        wp, bias = lin(384, 192)
        return lambda: raw.tapgemm(x192, wp, ksize=1, cout=384, bias=bias, act=L.ACT_GELU, want_aux=True, aux_grad=True)
    if name == 'fc2':
        wp, bias = lin(192, 384)
        return lambda: raw.tapgemm(x384, wp, ksize=1, cout=192, bias=bias, residual=x192)
    if name == 'fc2_dgrad':
        wp, _ = lin(384, 192)
        return lambda: raw.tapgemm(x192, wp, ksize=1, cout=384, flip=True, mask_src=x384, mask_mode=L.MASK_MUL)
    if name == 'wgrad_fc1':
        return lambda: raw.wgrad(x384, x192, ksize=1)
    if name in ('qkv_fold_1m', 'fc2_stats_1m'):  # evaluation kernels of the folded LayerNorm on one 1024 x 1024 tile
        big192 = bf(1, 1024, 1024, 192)
        big192[..., 180:] = 0
        if name == 'fc2_stats_1m':
            big384 = bf(1, 1024, 1024, 384)
            wp, bias = lin(192, 384)
            return lambda: raw.tapgemm(big384, wp, ksize=1, cout=192, bias=bias, residual=big192, ln_out=(180, 1e-5))
        wp, bias = lin(576, 192)
        xf = big192[..., :180].float()
        stats = torch.stack([xf.mean(-1), torch.rsqrt(xf.var(-1, unbiased=False) + 1e-5)], dim=-1).contiguous()
        wsum = wp.float().sum(dim=2).reshape(-1).contiguous()
        return lambda: raw.tapgemm(big192, wp, ksize=1, cout=576, bias=bias, ln_in=(stats, wsum))
    raise SystemExit(f'unknown kernel {name}')


def conv(name):
    c = 256 if name.endswith('256') else 64
    B, H, W = 16, 48, 48
    x, dy = bf(B, H, W, c), bf(B, H, W, c)
    w = torch.randn(c, c, 3, 3, device=dev) * 0.02
    bias = torch.zeros(c, device=dev)
    if name.startswith('conv'):
        flip = name.endswith('dgrad')
        wp = raw.pack_weight(w, c, c, transpose=flip)
        return lambda: raw.tapgemm(x, wp, ksize=3, cout=c, bias=None if flip else bias, act=L.ACT_NONE if flip else L.ACT_RELU,
                                   flip=flip)
    return lambda: raw.wgrad(dy, x, ksize=3)


def ca(name):
    B, H, W, C, CR = 16, 48, 48, 64, 4
    t, x, gy = bf(B, H, W, C), bf(B, H, W, C), bf(B, H, W, C)
    x32 = x.float()
    w1 = torch.randn(CR, C, 1, 1, device=dev) * 0.2
    b1 = torch.zeros(CR, device=dev)
    w2 = torch.randn(C, CR, 1, 1, device=dev) * 0.5
    b2 = torch.zeros(C, device=dev)
    p = raw.channel_pool(t)
    if name == 'ca_fwd':
        return lambda: raw.ca_forward(t, None, p, w1, b1, w2, b2, 1.0, x32=x32, want_f32=True)
    _, z, s = raw.ca_forward(t, x, p, w1, b1, w2, b2, 1.0)

    def run():
        with raw.zero_arena(dev, B * C + C + 64):
            raw.ca_backward(gy, t, s, z, p, w1, w2, 1.0)
    return run


def main():
    name = sys.argv[1]
    fn = ca(name) if name.startswith('ca_') else conv(name) if name.startswith(('conv', 'wgrad6', 'wgrad2')) else swin(name)
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    if '--time' in sys.argv:
        iters = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(0.02 * 1.9e9))
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({'kernel': name, 'us': round(e0.elapsed_time(e1) / iters * 1e3, 2)}))


if __name__ == '__main__':
    main()
