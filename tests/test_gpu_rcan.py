"""RCAN parity on the B200 (``-m gpu``): channel-attention kernels vs PyTorch fp32, the RCAN arch vs the
reference golden vectors and vs the CPU oracle (see tests/test_gpu_nets.py for the bars)."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import sr_oracle
from tests.test_gpu_nets import GOLDEN, MAX_ABS, PSNR_TOL, _build, _check_grad

pytestmark = pytest.mark.gpu


def test_channel_attention_kernels(cuda):
    from basicsr4rs_b200.ops.sr_b200 import raw
    b, h, w, c, cr = 3, 20, 12, 64, 4
    g = torch.Generator().manual_seed(0)
    t = torch.randn((b, h, w, c), generator=g).to(cuda).to(torch.bfloat16)
    x = torch.randn((b, h, w, c), generator=g).to(cuda).to(torch.bfloat16)
    w1 = (torch.randn((cr, c, 1, 1), generator=g) * 0.2).to(cuda)
    b1 = (torch.randn((cr,), generator=g) * 0.1).to(cuda)
    w2 = (torch.randn((c, cr, 1, 1), generator=g) * 0.5).to(cuda)
    b2 = (torch.randn((c,), generator=g) * 0.1).to(cuda)
    p = raw.channel_pool(t)
    assert torch.allclose(p, t.float().mean((1, 2)), atol=1e-5)
    z, s = raw.ca_fc(p, w1, b1, w2, b2)
    z_ref = F.relu(F.linear(p, w1.view(cr, c), b1))
    s_ref = torch.sigmoid(F.linear(z_ref, w2.view(c, cr), b2))
    assert torch.allclose(z, z_ref, atol=1e-5) and torch.allclose(s, s_ref, atol=1e-5)
    y = raw.ca_apply(t, x, s, 0.5)
    y_ref = x.float() + 0.5 * t.float() * s_ref.view(b, 1, 1, c)
    assert torch.allclose(y.float(), y_ref, atol=2e-2, rtol=1e-2)
    # backward pieces against autograd of the same fp32 expression
    tf = t.float().requires_grad_(True)
    params = [w1.clone().requires_grad_(True), b1.clone().requires_grad_(True), w2.clone().requires_grad_(True),
              b2.clone().requires_grad_(True)]
    pr = tf.mean((1, 2))
    sr = torch.sigmoid(F.linear(F.relu(F.linear(pr, params[0].view(cr, c), params[1])), params[2].view(c, cr),
                                params[3]))
    yr = x.float() + 0.5 * tf * sr.view(b, 1, 1, c)
    gy = torch.randn((b, h, w, c), generator=g).to(cuda).to(torch.bfloat16)
    yr.backward(gy.float())
    gs = raw.channel_dot(gy, t, scale=0.5)
    gw1, gb1, gw2, gb2, gp = raw.ca_fc_bwd(gs, s, z, p, w1, w2)
    for got, want in ((gw1, params[0].grad), (gb1, params[1].grad), (gw2, params[2].grad), (gb2, params[3].grad)):
        assert torch.allclose(got, want, rtol=1e-3, atol=1e-4 * want.abs().max().item())
    gt = raw.ca_apply_bwd(gy, s, gp, 0.5)
    assert torch.allclose(gt.float(), tf.grad, atol=2e-2 * tf.grad.abs().max().item())
    # the fused one-launch forms the RCAB Function uses must agree with the pieces
    for x32 in (None, x.float()):
        yf, zf, sf = raw.ca_forward(t, x if x32 is None else None, p, w1, b1, w2, b2, 0.5, x32=x32,
                                    want_f32=x32 is not None)
        if x32 is not None:
            yf, y32 = yf
            assert torch.allclose(y32, y_ref, atol=1e-5, rtol=1e-5)
        assert torch.equal(yf, y) and torch.allclose(zf, z, atol=1e-6) and torch.allclose(sf, s, atol=1e-6)
    with raw.zero_arena(cuda, b * c + c + 64):
        fw1, fb1, fw2, fb2, fgt, fcs = raw.ca_backward(gy, t, s, z, p, w1, w2, 0.5)
    for got, want in ((fw1, gw1), (fb1, gb1), (fw2, gw2), (fb2, gb2)):
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-5 * want.abs().max().item())
    assert (fgt.float() - gt.float()).abs().max().item() <= 1e-2 * gt.float().abs().max().item()
    assert torch.allclose(fcs, fgt.float().sum((0, 1, 2)), rtol=2e-3, atol=5e-2)  # (sums unrounded fp32)


@pytest.mark.parametrize('b,hw,c,cr', [(16, 48, 64, 4), (1, 7, 128, 8), (5, 33, 64, 16)])
def test_fused_channel_attention_backward_shapes(cuda, b, hw, c, cr):
    """The persistent arrive / release kernel at the bench shape, a single tiny image and a ragged one."""
    from basicsr4rs_b200.ops.sr_b200 import raw
    g = torch.Generator().manual_seed(b * 131 + hw)
    t = torch.randn((b, hw, hw, c), generator=g).to(cuda).to(torch.bfloat16)
    gy = torch.randn((b, hw, hw, c), generator=g).to(cuda).to(torch.bfloat16)
    w1 = (torch.randn((cr, c, 1, 1), generator=g) * 0.2).to(cuda)
    b1 = (torch.randn((cr,), generator=g) * 0.1).to(cuda)
    w2 = (torch.randn((c, cr, 1, 1), generator=g) * 0.5).to(cuda)
    b2 = (torch.randn((c,), generator=g) * 0.1).to(cuda)
    p = raw.channel_pool(t)
    z, s = raw.ca_fc(p, w1, b1, w2, b2)
    gs = raw.channel_dot(gy, t, scale=1.0)
    gw1, gb1, gw2, gb2, gp = raw.ca_fc_bwd(gs, s, z, p, w1, w2)
    gt = raw.ca_apply_bwd(gy, s, gp, 1.0)
    for _ in range(3):  # repeated launches: the sync words come zeroed from the arena every time
        with raw.zero_arena(cuda, b * c + c + 64):
            fw1, fb1, fw2, fb2, fgt, fcs = raw.ca_backward(gy, t, s, z, p, w1, w2, 1.0)
        for got, want in ((fw1, gw1), (fb1, gb1), (fw2, gw2), (fb2, gb2)):
            assert torch.allclose(got, want, rtol=2e-3, atol=2e-4 * want.abs().max().item())
        assert (fgt.float() - gt.float()).abs().max().item() <= 1e-2 * gt.float().abs().max().item()
        want_cs = (gy.float() * s.view(b, 1, 1, c) + gp.view(b, 1, 1, c) / (hw * hw)).sum((0, 1, 2))  # unrounded
        assert torch.allclose(fcs, want_cs, rtol=2e-3, atol=2e-3 * want_cs.abs().max().item())


@pytest.mark.parametrize('case', ['rcan_f64_g2_b2_x4', 'rcan_f64_g1_b2_sq8_x3'])
def test_rcan_matches_reference_golden(cuda, case):
    fx = torch.load(os.path.join(GOLDEN, case + '.pt'), weights_only=False)
    net = _build(fx, cuda)
    out = net(fx['x'].to(cuda))
    err = (out.detach().cpu() - fx['out']).abs().max().item()
    assert err <= MAX_ABS, f'max-abs {err:.3e}'
    assert abs(sr_oracle.psnr(out.detach().cpu(), fx['gt']) - sr_oracle.psnr(fx['out'], fx['gt'])) <= PSNR_TOL
    ((out - fx['gt'].to(cuda))**2).mean().backward()
    params = dict(net.named_parameters())
    for k, want in fx['grads'].items():
        _check_grad(params[k].grad, want, fx.get('autocast_rel', {}).get(k))


def test_rcan_full_depth_vs_oracle(cuda):
    """BASELINE config 3: 10 groups x 20 RCAB, default seeded init, 48x48 -- the deep-residual drift check."""
    from basicsr4rs_b200.archs import build_network
    kw = dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_group=10, num_block=20, squeeze_factor=16, upscale=4,
              res_scale=1, img_range=255.)
    torch.manual_seed(0)
    net = build_network(dict(type='RCAN', **kw))
    assert sum(p.numel() for p in net.parameters()) == 15592355
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(cuda).eval()
    g = torch.Generator().manual_seed(1234)
    x = torch.rand((1, 3, 48, 48), generator=g)
    with torch.no_grad():
        out = net(x.to(cuda)).cpu()
        ref = sr_oracle.rcan_forward(sd, x, num_group=10, num_block=20, upscale=4, res_scale=1, img_range=255.)
    err = (out - ref).abs().max().item()
    print(f'RCAN full depth: max-abs {err:.3e}, PSNR vs oracle {sr_oracle.psnr(out, ref, crop=1):.1f} dB')
    assert err <= MAX_ABS, f'max-abs {err:.3e}'
