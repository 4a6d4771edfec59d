"""SwinIR parity on the B200 (``-m gpu``): LayerNorm and fused window-attention kernels against plain
PyTorch fp32 compositions of the reference's own steps (roll -> window_partition -> softmax(qk^T*s + bias +
mask) v -> window_reverse -> roll), the fused block against the oracle block, and the SwinIR arch against
the reference golden vectors / the CPU oracle (bars: see tests/test_gpu_nets.py)."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import sr_oracle
from tests.test_gpu_nets import GOLDEN, MAX_ABS, PSNR_TOL, _build, _check_grad

pytestmark = pytest.mark.gpu


def _ops():
    from basicsr4rs_b200.ops.sr_b200 import swin_ops
    return swin_ops


@pytest.mark.parametrize('c,cp,tokens', [(180, 192, 1000), (60, 64, 333), (360, 384, 77)])
def test_layernorm_fwd_bwd(cuda, c, cp, tokens):
    so = _ops()
    g = torch.Generator().manual_seed(0)
    x = torch.zeros((1, tokens, 1, cp))
    x[..., :c] = torch.randn((1, tokens, 1, c), generator=g) * 3 + 1
    x = x.to(cuda).to(torch.bfloat16)
    gamma = (1 + 0.1 * torch.randn(c, generator=g)).to(cuda)
    beta = (0.1 * torch.randn(c, generator=g)).to(cuda)
    y, mean, rstd = so.layernorm_fwd(x, gamma, beta, c)
    xr = x[..., :c].float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (c,), gr, br, 1e-5)
    assert torch.allclose(y[..., :c].float(), ref, atol=3e-2, rtol=1e-2)
    assert torch.count_nonzero(y[..., c:]) == 0
    assert torch.allclose(mean, xr.detach().mean(-1).flatten(), atol=1e-4)
    gy = torch.zeros_like(x)
    gy[..., :c] = torch.randn((1, tokens, 1, c), generator=g).to(cuda).to(torch.bfloat16)
    gres = torch.zeros_like(x)
    gres[..., :c] = torch.randn((1, tokens, 1, c), generator=g).to(cuda).to(torch.bfloat16)
    ref.backward(gy[..., :c].float())
    gx, gg, gb = so.layernorm_bwd(gy, x, mean, rstd, gamma, c, gres=gres)
    want = xr.grad + gres[..., :c].float()
    assert torch.allclose(gx[..., :c].float(), want, atol=2e-2 * want.abs().max().item())
    assert torch.allclose(gg, gr.grad, rtol=1e-3, atol=1e-3 * gr.grad.abs().max().item())
    assert torch.allclose(gb, br.grad, rtol=1e-3, atol=1e-3 * br.grad.abs().max().item())


def _ref_window_attention(qkv_pad, table, nh, hd, ws, shift, h, w):
    """fp32 PyTorch restatement of swinir_arch.py:293-316 + :151-172 given the (padded) qkv tensor."""
    b = qkv_pad.shape[0]
    ca = nh * 32
    q, k, v = [qkv_pad[..., i * ca:(i + 1) * ca].reshape(b, h, w, nh, 32)[..., :hd] for i in range(3)]

    def part(t):  # [b,h,w,nh,hd] -> [b*nW, nh, ws*ws, hd]
        if shift:
            t = torch.roll(t, shifts=(-shift, -shift), dims=(1, 2))
        t = sr_oracle.window_partition(t.reshape(b, h, w, nh * hd), ws).reshape(-1, ws * ws, nh, hd)
        return t.permute(0, 2, 1, 3)

    q, k, v = part(q), part(k), part(v)
    attn = (q * hd**-0.5) @ k.transpose(-2, -1)
    idx = sr_oracle.relative_position_index(ws).reshape(-1).to(table.device)
    attn = attn + table[idx].reshape(ws * ws, ws * ws, nh).permute(2, 0, 1).unsqueeze(0)
    if shift:
        mask = sr_oracle.calculate_mask(h, w, ws, shift).to(table.device)
        nw = mask.shape[0]
        attn = (attn.reshape(b, nw, nh, ws * ws, ws * ws) + mask[None, :, None]).reshape(-1, nh, ws * ws, ws * ws)
    out = (torch.softmax(attn, -1) @ v).transpose(1, 2).reshape(-1, ws, ws, nh * hd)
    out = sr_oracle.window_reverse(out, ws, h, w)
    if shift:
        out = torch.roll(out, shifts=(shift, shift), dims=(1, 2))
    return out.reshape(b, h, w, nh, hd)


@pytest.mark.parametrize('b,h,w,nh,hd,shift,ws', [(2, 16, 24, 6, 30, 0, 8), (2, 16, 24, 6, 30, 4, 8), (1, 8, 8, 6, 10, 0, 8),
                                                  (3, 24, 40, 4, 32, 4, 8), (16, 64, 64, 6, 30, 4, 8),
                                                  (1, 64, 64, 6, 30, 4, 8), (1, 24, 16, 2, 32, 3, 8),
                                                  # the fork's remote-sensing recipes: 6-wide windows; and 7 (Swin default)
                                                  (2, 12, 18, 6, 30, 0, 6), (2, 12, 18, 6, 30, 3, 6), (1, 48, 48, 6, 30, 3, 6),
                                                  (1, 14, 21, 3, 20, 3, 7), (1, 8, 12, 2, 16, 2, 4)])
def test_window_attention_fwd_bwd(cuda, b, h, w, nh, hd, shift, ws):
    so = _ops()
    g = torch.Generator().manual_seed(1)
    ca = nh * 32
    qkv = torch.zeros((b, h, w, 3, nh, 32))
    qkv[..., :hd] = torch.randn((b, h, w, 3, nh, hd), generator=g)
    qkv = qkv.reshape(b, h, w, 3 * ca).to(cuda).to(torch.bfloat16)
    table = (torch.randn(((2 * ws - 1)**2, nh), generator=g) * 0.5).to(cuda)
    out, stats = so.window_attention_fwd(qkv, table, nh, ws, shift, hd**-0.5, want_stats=True)
    assert torch.equal(out, so.window_attention_fwd(qkv, table, nh, ws, shift, hd**-0.5))
    qr = qkv.float().requires_grad_(True)
    tr = table.clone().requires_grad_(True)
    ref = _ref_window_attention(qr, tr, nh, hd, ws, shift, h, w)
    got = out.reshape(b, h, w, nh, 32)
    assert torch.count_nonzero(got[..., hd:]) == 0
    err = (got[..., :hd].float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item(), f'fwd err {err:.3e}'
    go = torch.zeros((b, h, w, nh, 32))
    go[..., :hd] = torch.randn((b, h, w, nh, hd), generator=g)
    go = go.reshape(b, h, w, ca).to(cuda).to(torch.bfloat16)
    ref.backward(go.reshape(b, h, w, nh, 32)[..., :hd].float())
    # with the forward's statistics buffer (window 8: the tcgen05 backward) and without it (mma.sync, recomputes them)
    for st, tc in ((stats, True), (stats, False), (None, True)):
        gqkv, gtable = so.window_attention_bwd(qkv, go, table, nh, ws, shift, hd**-0.5, stats=st, use_tc=tc)
        want = qr.grad
        assert torch.count_nonzero(gqkv.reshape(b, h, w, 3, nh, 32)[..., hd:]) == 0
        rel = ((gqkv.float() - want).norm() / want.norm()).item()
        assert rel <= 2e-2, f'gqkv rel-L2 {rel:.3e} (stats: {st is not None})'
        relt = ((gtable - tr.grad).norm() / tr.grad.norm()).item()
        assert relt <= 2e-2, f'gtable rel-L2 {relt:.3e} (stats: {st is not None})'


@pytest.mark.parametrize('shift', [0, 4])
def test_window_attention_out_alpha_and_alpha_on_bias(cuda, shift):
    """The attention branch's per-sample DropPath factor carried by the attention output (``out_alpha``: every row of
    sample b times alpha[b], the SRB200_ATTN_ONES lane holds alpha[b]) and the matching Linear epilogue
    (SRB200_EXT_ALPHA_ON_BIAS: v = acc + alpha[b] * bias): together alpha (o W^T + bias), exactly as the plain path."""
    so, raw = _ops(), _ops().raw
    b, h, w, nh, hd = 3, 16, 16, 6, 30
    g = torch.Generator().manual_seed(5)
    ca = nh * 32
    qkv = torch.zeros((b, h, w, 3, nh, 32))
    qkv[..., :hd] = torch.randn((b, h, w, 3, nh, hd), generator=g)
    qkv = qkv.reshape(b, h, w, 3 * ca).to(cuda).to(torch.bfloat16)
    table = (torch.randn((225, nh), generator=g) * 0.5).to(cuda)
    alpha = torch.tensor([0.0, 1.0 / 0.9, 0.5], device=cuda)
    assert so.attn_on_tc(qkv, nh, 8, shift)
    plain = so.window_attention_fwd(qkv, table, nh, 8, shift, hd**-0.5, ones=True)
    scaled = so.window_attention_fwd(qkv, table, nh, 8, shift, hd**-0.5, ones=True, alpha=alpha)
    want = plain.float() * alpha.view(b, 1, 1, 1)
    assert (scaled.float() - want).abs().max().item() <= 2.0**-7 * want.abs().max().item()
    assert torch.equal(scaled[0], torch.zeros_like(scaled[0]))            # a dropped sample is exactly zero
    assert torch.allclose(scaled[..., 31].float(), alpha.view(b, 1, 1).expand(b, h, w), rtol=2.0**-8)  # the ones lane
    # proj with the factor already in its input == proj with the factor in its epilogue
    wt = (torch.randn((192, ca), generator=g) * 0.05).to(cuda)
    bias = torch.randn((192,), generator=g).to(cuda)
    res = torch.randn((b, h, w, 192), generator=g).to(cuda).to(torch.bfloat16)
    wp = raw.pack_weight(wt, 192, ca)
    y_epi = raw.tapgemm(plain, wp, ksize=1, cout=192, bias=bias, residual=res, alpha_per_sample=alpha)
    y_in = raw.tapgemm(scaled, wp, ksize=1, cout=192, bias=bias, residual=res, alpha_per_sample=alpha, alpha_on_bias=True)
    assert (y_epi.float() - y_in.float()).abs().max().item() <= 3e-2 * y_epi.float().abs().max().item()
    assert torch.equal(y_in[0], (res[0].float() + 0.0).to(torch.bfloat16))  # alpha = 0: the skip passes through untouched


def test_swin_block_vs_oracle(cuda):
    """One shifted SwinTransformerBlock (C=180, 6 heads, ws 8, shift 4) forward + all gradients."""
    so = _ops()
    c, nh, hidden, b, h, w = 180, 6, 360, 2, 16, 24
    g = torch.Generator().manual_seed(2)
    shapes = {'norm1.weight': (c,), 'norm1.bias': (c,), 'attn.qkv.weight': (3 * c, c), 'attn.qkv.bias': (3 * c,),
              'attn.relative_position_bias_table': (225, nh), 'attn.proj.weight': (c, c), 'attn.proj.bias': (c,),
              'norm2.weight': (c,), 'norm2.bias': (c,), 'mlp.fc1.weight': (hidden, c), 'mlp.fc1.bias': (hidden,),
              'mlp.fc2.weight': (c, hidden), 'mlp.fc2.bias': (c,)}
    sd = sr_oracle.fill_state_dict_({'blk.' + k: torch.zeros(s) for k, s in shapes.items()})
    x = torch.randn((b, h * w, c), generator=g)
    for shift in (0, 4):
        ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xr = x.clone().requires_grad_(True)
        ref = sr_oracle.swin_block(ref_sd, 'blk', xr, h, w, nh, 8, shift)
        gy = torch.randn((b, h * w, c), generator=g)
        ref.backward(gy)
        p = {k[4:]: v.to(cuda).requires_grad_(True) for k, v in sd.items()}
        xp = torch.zeros((b, h, w, 192))
        xp[..., :c] = x.reshape(b, h, w, c)
        xp = xp.to(cuda).to(torch.bfloat16).requires_grad_(True)
        out = so.swin_block(xp, p['norm1.weight'], p['norm1.bias'], p['attn.qkv.weight'], p['attn.qkv.bias'],
                            p['attn.relative_position_bias_table'], p['attn.proj.weight'], p['attn.proj.bias'],
                            p['norm2.weight'], p['norm2.bias'], p['mlp.fc1.weight'], p['mlp.fc1.bias'],
                            p['mlp.fc2.weight'], p['mlp.fc2.bias'], nh, 8, shift)
        err = (out[..., :c].float().cpu().reshape(b, h * w, c) - ref.detach()).abs().max().item()
        assert err <= 2e-2 * ref.abs().max().item(), f'shift {shift}: fwd err {err:.3e}'
        assert torch.count_nonzero(out[..., c:]) == 0
        gyp = torch.zeros((b, h, w, 192))
        gyp[..., :c] = gy.reshape(b, h, w, c)
        out.backward(gyp.to(cuda).to(torch.bfloat16))
        rel = ((xp.grad[..., :c].float().cpu().reshape(b, h * w, c) - xr.grad).norm() / xr.grad.norm()).item()
        assert rel <= 3e-2, f'shift {shift}: gx rel-L2 {rel:.3e}'
        for k in shapes:
            want = ref_sd['blk.' + k].grad
            rel = ((p[k].grad.cpu() - want).norm() / want.norm()).item()
            assert rel <= 3e-2, f'shift {shift}: {k} rel-L2 {rel:.3e}'


@pytest.mark.parametrize('case', ['swinir_c180_d2x2_x4', 'swinir_c60_d2_x2', 'swinir_c60_d2_direct_x2',
                                  'swinir_c60_d2_nearest_x4', 'swinir_c60_ws6_in4_x2', 'swinir_c60_d2_denoise_x1'])
def test_swinir_matches_reference_golden(cuda, case):
    fx = torch.load(os.path.join(GOLDEN, case + '.pt'), weights_only=False)
    net = _build(fx, cuda)
    out = net(fx['x'].to(cuda))
    err = (out.detach().cpu() - fx['out']).abs().max().item()
    print(f'{case}: max-abs {err:.3e}')
    assert err <= MAX_ABS, f'max-abs {err:.3e}'
    assert abs(sr_oracle.psnr(out.detach().cpu(), fx['gt']) - sr_oracle.psnr(fx['out'], fx['gt'])) <= PSNR_TOL
    ((out - fx['gt'].to(cuda))**2).mean().backward()
    params = dict(net.named_parameters())
    for k, want in fx['grads'].items():
        _check_grad(params[k].grad, want, fx.get('autocast_rel', {}).get(k))


def test_swinir_full_size_vs_oracle(cuda):
    """BASELINE config 4: embed 180, 6 RSTB x 6 layers, window 8, 6 heads, x4, 64x64 LR, default seeded init."""
    from basicsr4rs_b200.archs import build_network
    kw = dict(upscale=4, in_chans=3, img_size=64, window_size=8, img_range=1., depths=[6] * 6, embed_dim=180,
              num_heads=[6] * 6, mlp_ratio=2, upsampler='pixelshuffle', resi_connection='1conv')
    torch.manual_seed(0)
    net = build_network(dict(type='SwinIR', **kw))
    assert sum(p.numel() for p in net.parameters()) == 11900199 and len(net.state_dict()) == 550
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(cuda).eval()
    g = torch.Generator().manual_seed(1234)
    x = torch.rand((1, 3, 64, 64), generator=g)
    with torch.no_grad():
        out = net(x.to(cuda)).cpu()
        ref = sr_oracle.swinir_forward(sd, x, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6, window_size=8,
                                       upscale=4, img_range=1.)
    err = (out - ref).abs().max().item()
    print(f'SwinIR full: max-abs {err:.3e}, PSNR vs oracle {sr_oracle.psnr(out, ref, crop=1):.1f} dB')
    assert err <= MAX_ABS, f'max-abs {err:.3e}'
    # training mode with stochastic depth runs and produces finite gradients
    net.train()
    o = net(torch.rand((2, 3, 64, 64), device=cuda))
    o.mean().backward()
    assert all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)


def test_nearest_up2_bit_exact(cuda):
    """F.interpolate(scale_factor=2, mode='nearest') (swinir_arch.py:911-912) as an NHWC remap, and its gradient."""
    import torch.nn.functional as F
    from basicsr4rs_b200.ops.sr_b200 import raw
    x = torch.randn((2, 7, 5, 64), device=cuda).to(torch.bfloat16)
    up = raw.nearest_up2(x)
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2, mode='nearest').permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    g = torch.randn((2, 14, 10, 64), device=cuda).to(torch.bfloat16)
    back = raw.nearest_up2(g, inverse=True)
    want = g.float().view(2, 7, 2, 5, 2, 64).sum((2, 4)).to(torch.bfloat16)
    assert torch.equal(back.view(torch.int16), want.view(torch.int16))


# ------------------------------------------------------------------ fp32 mode of the token path (north_star: <= 1e-4)
@pytest.mark.parametrize('c,cp,tokens', [(180, 192, 1000), (60, 64, 333), (360, 384, 77), (30, 64, 5)])
def test_layernorm_f32_kernel(cuda, c, cp, tokens):
    """srb200_layernorm_f32: fp32 result within fp32 rounding of F.layer_norm; the split output is exactly the
    [hi | lo | hi] bf16 split of that result; pad channels are written as zero whatever the input holds there."""
    from basicsr4rs_b200.ops.sr_b200 import fp32_mode as f32
    g = torch.Generator().manual_seed(0)
    x = torch.randn((1, tokens, 1, cp), generator=g) * 3 + 1  # (pads deliberately non-zero)
    norm = torch.nn.LayerNorm(c, eps=1e-5)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(c, generator=g))
        norm.bias.copy_(0.1 * torch.randn(c, generator=g))
    ref = F.layer_norm(x[..., :c].double(), (c,), norm.weight.double(), norm.bias.double(), 1e-5)
    norm = norm.to(cuda)
    y = f32.layer_norm(x.to(cuda), norm)
    assert torch.count_nonzero(y[..., c:]) == 0
    err = (y[..., :c].double().cpu() - ref).abs().max().item()
    assert err <= 5e-6, f'layernorm_f32 max-abs {err:.3e}'
    ys = f32.layer_norm(x.to(cuda), norm, split=True)
    assert torch.equal(ys, f32.split3(y))


@pytest.mark.parametrize('b,h,w,nh,hd,shift,ws', [(2, 16, 24, 6, 30, 0, 8), (2, 16, 24, 6, 30, 4, 8), (1, 8, 8, 6, 10, 0, 8),
                                                  (3, 24, 40, 4, 32, 4, 8), (1, 24, 16, 2, 32, 3, 8),
                                                  (2, 12, 18, 6, 30, 3, 6), (1, 14, 21, 3, 20, 3, 7), (1, 8, 12, 2, 16, 2, 4)])
def test_window_attention_f32_kernel(cuda, b, h, w, nh, hd, shift, ws):
    """srb200_window_attention_f32 against the fp64 restatement of swinir_arch.py:151-172 + :293-316 (roll, partition,
    bias, analytic mask, softmax, reverse): fp32-rounding agreement, zero pad lanes, split output == split3(fp32)."""
    from basicsr4rs_b200.ops.sr_b200 import fp32_mode as f32
    g = torch.Generator().manual_seed(1)
    ca = (nh * 32 + 63) // 64 * 64
    qkv = torch.zeros((b, h, w, 3, ca))
    packed = torch.zeros((b, h, w, 3, nh, 32))
    packed[..., :hd] = torch.randn((b, h, w, 3, nh, hd), generator=g) * 1.5
    qkv[..., :nh * 32] = packed.reshape(b, h, w, 3, nh * 32)
    table = torch.randn(((2 * ws - 1)**2, nh), generator=g) * 0.5
    ref = _ref_window_attention(packed.reshape(b, h, w, 3 * nh * 32).double(), table.double(), nh, hd, ws, shift, h, w)
    qd = qkv.reshape(b, h, w, 3 * ca).to(cuda)
    out = f32.window_attention(qd, table.to(cuda), nh, ws, shift, hd**-0.5, split=False)
    got = out[..., :nh * 32].reshape(b, h, w, nh, 32)
    assert torch.count_nonzero(got[..., hd:]) == 0
    err = (got[..., :hd].double().cpu() - ref).abs().max().item()
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()), f'window_attention_f32 max-abs {err:.3e}'
    outs = f32.window_attention(qd, table.to(cuda), nh, ws, shift, hd**-0.5, split=True)
    want = f32.split3(out)
    for part in range(3):  # (channels beyond the heads are not written by the kernel)
        assert torch.equal(outs[..., part * ca:part * ca + nh * 32], want[..., part * ca:part * ca + nh * 32]), part


@pytest.mark.parametrize('case', ['swinir_c180_d2x2_x4', 'swinir_c60_d2_x2', 'swinir_c60_d2_direct_x2',
                                  'swinir_c60_d2_nearest_x4', 'swinir_c60_ws6_in4_x2', 'swinir_c60_d2_denoise_x1'])
def test_swinir_fp32_mode_matches_reference_golden_within_1e4(cuda, case):
    """compute_dtype='fp32' on SwinIR (split-operand tap-GEMMs + fp32 LayerNorm / window attention kernels): every
    reconstruction branch within 1e-4 max-abs of the unmodified reference's fp32 output."""
    fx = torch.load(os.path.join(GOLDEN, case + '.pt'), weights_only=False)
    net = _build(fx, cuda)
    with torch.no_grad():
        out16 = net(fx['x'].to(cuda)).cpu()
        net.compute_dtype = 'fp32'
        out = net(fx['x'].to(cuda)).cpu()
    err, err16 = (out - fx['out']).abs().max().item(), (out16 - fx['out']).abs().max().item()
    print(f'{case}: fp32 mode max-abs {err:.2e} (bf16 path {err16:.2e})')
    assert out.shape == fx['out'].shape and err <= 1e-4, f'max-abs {err:.3e}'
    assert err < 0.2 * err16


def test_swinir_fp32_mode_full_size_within_1e4(cuda):
    """BASELINE config 4 (embed 180, 6 x 6 blocks, window 8, x4, 64x64 LR) in fp32 mode against the fp32 oracle."""
    from basicsr4rs_b200.archs import build_network
    kw = dict(upscale=4, in_chans=3, img_size=64, window_size=8, img_range=1., depths=[6] * 6, embed_dim=180,
              num_heads=[6] * 6, mlp_ratio=2, upsampler='pixelshuffle', resi_connection='1conv')
    torch.manual_seed(0)
    net = build_network(dict(type='SwinIR', compute_dtype='fp32', **kw))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(cuda).eval()
    x = torch.rand((1, 3, 64, 64), generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        out = net(x.to(cuda)).cpu()
        ref = sr_oracle.swinir_forward(sd, x, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6, window_size=8,
                                       upscale=4, img_range=1.)
    err = (out - ref).abs().max().item()
    print(f'SwinIR full size, fp32 mode: max-abs {err:.2e}')
    assert err <= 1e-4, f'max-abs {err:.3e}'


# ------------------------------------------------------------------ LayerNorm folded into the surrounding GEMMs (evaluation)
@pytest.mark.parametrize('c,cs,n_out,act', [(180, 192, 576, 'none'), (180, 192, 384, 'gelu'), (120, 128, 384, 'gelu')])
def test_tapgemm_layernorm_fold(cuda, c, cs, n_out, act):
    """srb200_tapgemm_ext.ln_stats_out / ln_stats_in: the producer's (mean, rstd) of the rows it stores match a
    LayerNorm's statistics of that bf16 tensor and leave the stored tensor bit-identical; the consumer on the RAW rows
    with gamma-folded weights reproduces Linear(LayerNorm(x)) (fp32 PyTorch) at least as closely as the
    LayerNorm-kernel + plain-GEMM path does."""
    so, raw = _ops(), _ops().raw
    from basicsr4rs_b200 import _lib as L
    g = torch.Generator().manual_seed(3)
    b, h, w = 2, 24, 40
    o = torch.zeros((b, h, w, cs))
    o[..., :c] = torch.randn((b, h, w, c), generator=g)
    res = torch.zeros((b, h, w, cs))
    res[..., :c] = torch.randn((b, h, w, c), generator=g) * 2 + 0.7
    o, res = o.to(cuda).to(torch.bfloat16), res.to(cuda).to(torch.bfloat16)
    wp_ = (torch.randn((c, c), generator=g) * 0.05).to(cuda)
    bp_ = (torch.randn((c,), generator=g) * 0.1).to(cuda)
    bias_p = torch.zeros(cs, device=cuda)
    bias_p[:c] = bp_
    packed_p = raw.pack_weight(wp_, cs, cs)
    x1 = raw.tapgemm(o, packed_p, ksize=1, cout=cs, bias=bias_p, residual=res)
    x1s, st = raw.tapgemm(o, packed_p, ksize=1, cout=cs, bias=bias_p, residual=res, ln_out=(c, 1e-5))
    assert torch.equal(x1, x1s)
    xf = x1[..., :c].float()
    mean, var = xf.mean(-1), xf.var(-1, unbiased=False)
    assert torch.allclose(st[..., 0], mean, atol=1e-5, rtol=1e-5)
    assert torch.allclose(st[..., 1], torch.rsqrt(var + 1e-5), rtol=2e-4)
    # consumer
    gamma = (1 + 0.2 * torch.randn(c, generator=g)).to(cuda)
    beta = (0.2 * torch.randn(c, generator=g)).to(cuda)
    n_real = n_out - 8
    wl = (torch.randn((n_real, c), generator=g) * 0.05).to(cuda)
    bl = (torch.randn((n_real,), generator=g) * 0.1).to(cuda)
    a = L.ACT_GELU if act == 'gelu' else L.ACT_NONE
    ref = F.linear(F.layer_norm(xf, (c,), gamma, beta, 1e-5), wl, bl)
    ref = F.gelu(ref) if act == 'gelu' else ref
    # (1) the unfolded path: LayerNorm kernel -> plain GEMM
    xn = so.layernorm_fwd(x1, gamma, beta, c)[0]
    bias_l = torch.zeros(n_out, device=cuda)
    bias_l[:n_real] = bl
    y_plain = raw.tapgemm(xn, raw.pack_weight(wl, n_out, cs), ksize=1, cout=n_out, bias=bias_l, act=a)
    # (2) folded
    packed_f = raw.pack_weight((wl * gamma[None, :]).contiguous(), n_out, cs)
    wsum = packed_f.float().sum(dim=2).reshape(-1).contiguous()
    bias_f = torch.zeros(n_out, device=cuda)
    bias_f[:n_real] = wl @ beta + bl
    y_fold = raw.tapgemm(x1, packed_f, ksize=1, cout=n_out, bias=bias_f, act=a, ln_in=(st, wsum))
    assert torch.count_nonzero(y_fold[..., n_real:]) == 0
    e_plain = (y_plain[..., :n_real].float() - ref).abs().max().item()
    e_fold = (y_fold[..., :n_real].float() - ref).abs().max().item()
    print(f'fold: max-abs vs fp32 {e_fold:.3e} (LayerNorm kernel + GEMM: {e_plain:.3e})')
    assert e_fold <= 2e-2 * ref.abs().max().item() and e_fold <= 1.5 * e_plain + 1e-3


def test_swinir_eval_with_folded_layernorm(cuda, monkeypatch):
    """BasicLayer's evaluation path on big tiles (LayerNorm folded into qkv / fc1, statistics from proj / fc2), forced on
    for a small input: same output as the LayerNorm-kernel path within bf16 noise, and within the 1e-2 bar of the
    reference golden fixture."""
    from basicsr4rs_b200.archs import swinir_arch
    fx = torch.load(os.path.join(GOLDEN, 'swinir_c180_d2x2_x4.pt'), weights_only=False)
    net = _build(fx, cuda)
    with torch.no_grad():
        plain = net(fx['x'].to(cuda)).cpu()
        monkeypatch.setattr(swinir_arch.BasicLayer, 'FOLD_MIN_TOKENS', 0)
        n0 = _launches()
        folded = net(fx['x'].to(cuda)).cpu()
        n_fold = _launches() - n0
        n0 = _launches()
        monkeypatch.setattr(swinir_arch.BasicLayer, 'FOLD_MIN_TOKENS', 1 << 60)
        net(fx['x'].to(cuda))
        n_plain = _launches() - n0
    assert n_fold != n_plain, 'the folded path did not run'
    err_p, err_f = (plain - fx['out']).abs().max().item(), (folded - fx['out']).abs().max().item()
    print(f'folded LayerNorm: max-abs {err_f:.3e} (LayerNorm kernels {err_p:.3e}); {n_fold} vs {n_plain} launches')
    assert err_f <= MAX_ABS, err_f
    assert (folded - plain).abs().max().item() <= 5e-3


def _launches():
    from basicsr4rs_b200 import _lib
    return _lib.launch_count
