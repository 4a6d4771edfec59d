"""Network-level parity on the B200 (``-m gpu``): the registered archs of this package, computing through
the C-ABI, against (a) the golden vectors generated from the unmodified reference
(tests/golden/*.pt) and (b) the CPU oracle on freshly seeded inputs.

Bars (BASELINE.md section 4, bf16 tensor-core path): output max-abs <= 1e-2 on [0,1] images and
|dPSNR| <= 0.05 dB.

Gradients (north_star states no bar; recorded choice): gradients of a SMOOTH loss are compared per tensor with
cosine similarity >= 0.995 and relative L2 <= 1e-1.  The relative-L2 bar cannot be the 2e-2 SURVEY.md
guessed: a bf16 forward perturbs pre-activations by ~2^-9 relative, which flips the ReLU mask at the
~0.3 % of positions closest to zero, and every flipped position changes its gradient wholesale -- a
relative L2 of ~sqrt(0.003) = 5e-2 that no kernel can avoid.  ``test_edsr_l_full_size_vs_oracle`` therefore
also calibrates against PyTorch's own bf16 autocast of the oracle on the CPU: our error must not exceed
2x the error bf16 autocast itself makes against fp32.
"""
import glob
import os

import pytest
import torch

from oracle import sr_oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MAX_ABS = 1e-2
PSNR_TOL = 0.05
GRAD_REL_L2 = 1e-1
GRAD_COS = 0.995


def _build(fx, dev):
    from basicsr4rs_b200.archs import build_network
    torch.manual_seed(0)
    net = build_network(dict(type=fx['arch'], **fx['kwargs']))
    sd = net.state_dict()
    assert list(sd.keys()) == fx['state_keys']
    sr_oracle.fill_state_dict_(sd)
    net.load_state_dict(sd, strict=True)
    return net.to(dev).eval()


def _check_grad(got, want, calib=None):
    """rel-L2 / cosine bars; ``calib`` = the error PyTorch's CPU bf16 autocast of the reference makes on this
    gradient (fixture['autocast_rel']), which loosens the bar to 2x that when it is the larger one."""
    got = got.detach().float().cpu()
    if isinstance(want, dict):
        g, w = got.flatten()[::want['stride']], want['sample']
        assert abs(got.double().norm().item() - want['l2'].item()) <= GRAD_REL_L2 * want['l2'].item()
    else:
        g, w = got, want
    rel = ((g - w).norm() / (w.norm() + 1e-12)).item()
    cos = (torch.dot(g.flatten().double(), w.flatten().double()) / (g.double().norm() * w.double().norm() + 1e-30)).item()
    bar = max(GRAD_REL_L2, 2.0 * calib) if calib is not None else GRAD_REL_L2
    cos_bar = GRAD_COS if bar == GRAD_REL_L2 else 1.0 - 0.5 * bar * bar * 1.5
    assert rel <= bar and cos >= cos_bar, f'relative L2 {rel:.3e} (bar {bar:.3e}), cosine {cos:.5f}'


def _golden_cases(prefix):
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, prefix + '*.pt')))


@pytest.mark.parametrize('case', _golden_cases('edsr_'))
def test_edsr_matches_reference_golden(cuda, case):
    fx = torch.load(os.path.join(GOLDEN, case + '.pt'), weights_only=False)
    net = _build(fx, cuda)
    out = net(fx['x'].to(cuda))
    assert out.shape == fx['out'].shape and out.dtype == torch.float32
    err = (out.detach().cpu() - fx['out']).abs().max().item()
    assert err <= MAX_ABS, f'max-abs {err:.3e}'
    dpsnr = abs(sr_oracle.psnr(out.detach().cpu(), fx['gt']) - sr_oracle.psnr(fx['out'], fx['gt']))
    assert dpsnr <= PSNR_TOL
    # PSNR of our output against the reference output itself must be very high
    assert sr_oracle.psnr(out.detach().cpu().clamp(0, 1), fx['out'].clamp(0, 1), crop=1) > 50.0
    loss = (out - fx['gt'].to(cuda)).abs().mean()
    assert abs(loss.item() - fx['loss'].item()) <= 1e-3
    assert fx['grad_loss'] == 'mse'  # gradients of the smooth loss (see tests/golden/make_golden.py)
    ((out - fx['gt'].to(cuda))**2).mean().backward()
    params = dict(net.named_parameters())
    for k, want in fx['grads'].items():
        _check_grad(params[k].grad, want, fx.get('autocast_rel', {}).get(k))


def test_edsr_l_full_size_vs_oracle(cuda):
    """BASELINE config 2 arch (32 blocks, 256 ch, res_scale 0.1) at 48x48, default seeded init."""
    from basicsr4rs_b200.archs import build_network
    kw = dict(num_in_ch=3, num_out_ch=3, num_feat=256, num_block=32, upscale=4, res_scale=0.1, img_range=255.)
    torch.manual_seed(0)
    net = build_network(dict(type='EDSR', **kw))
    assert sum(p.numel() for p in net.parameters()) == 43089923  # SURVEY.md section 8c known answer
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(cuda)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand((2, 3, 48, 48), generator=g)
    gt = torch.rand((2, 3, 192, 192), generator=g)
    out = net(x.to(cuda))
    ((out - gt.to(cuda))**2).mean().backward()
    for v in sd.values():
        v.requires_grad_(True)
    ref = sr_oracle.edsr_forward(sd, x, num_block=32, upscale=4, res_scale=0.1, img_range=255.)
    ((ref - gt)**2).mean().backward()
    err = (out.detach().cpu() - ref.detach()).abs().max().item()
    assert err <= MAX_ABS, f'max-abs {err:.3e}'
    assert abs(sr_oracle.psnr(out.detach().cpu(), gt) - sr_oracle.psnr(ref.detach(), gt)) <= PSNR_TOL
    # calibration: the same oracle under PyTorch's CPU bf16 autocast
    sd16 = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    with torch.autocast('cpu', dtype=torch.bfloat16):
        ref16 = sr_oracle.edsr_forward(sd16, x, num_block=32, upscale=4, res_scale=0.1, img_range=255.)
    ((ref16.float() - gt)**2).mean().backward()
    params = dict(net.named_parameters())
    report = {}
    for k in ['conv_first.weight', 'body.0.conv1.weight', 'body.31.conv2.weight', 'body.15.conv1.bias',
              'conv_after_body.weight', 'upsample.0.weight', 'upsample.2.weight', 'upsample.2.bias',
              'conv_last.weight', 'conv_last.bias']:
        want = sd[k].grad
        rel = ((params[k].grad.cpu() - want).norm() / want.norm()).item()
        rel16 = ((sd16[k].grad - want).norm() / want.norm()).item()
        report[k] = (rel, rel16)
        assert rel <= max(2.0 * rel16, 3e-2), f'{k}: relative L2 {rel:.3e} vs CPU bf16 autocast {rel16:.3e}'
    print('grad rel-L2 (ours, cpu-bf16-autocast):', {k: (round(a, 4), round(b, 4)) for k, (a, b) in report.items()})


def test_edsr_module_contract(cuda):
    """state_dict round trip, deepcopy (EMA copy, sr_model.py:51), no_grad, train/eval, odd sizes."""
    import copy
    from basicsr4rs_b200.archs import build_network
    torch.manual_seed(0)
    net = build_network(dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=64, num_block=2, upscale=2,
                             res_scale=1.0)).to(cuda)
    x = torch.rand((1, 3, 13, 21), device=cuda)
    with torch.no_grad():
        a = net(x)
    ema = copy.deepcopy(net).eval()
    with torch.no_grad():
        b = ema(x)
    assert a.shape == (1, 3, 26, 42) and torch.equal(a, b)
    # in-place weight update must invalidate the packed-weight cache
    with torch.no_grad():
        net.conv_last.weight.mul_(0.5)
        net.conv_last.bias.zero_()
        c = net(x)
    assert not torch.equal(a, c)
    net2 = build_network(dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=64, num_block=2, upscale=2,
                              res_scale=1.0)).to(cuda)
    net2.load_state_dict(net.state_dict(), strict=True)
    with torch.no_grad():
        assert torch.equal(net2(x), c)
    with pytest.raises(RuntimeError):
        net(x.cpu())


def test_residual_block_standalone_nchw(cuda):
    """ResidualBlockNoBN keeps the reference's NCHW nn.Module contract for other archs (arch_util.py:64-88)."""
    from basicsr4rs_b200.archs.arch_util import ResidualBlockNoBN
    torch.manual_seed(0)
    blk = ResidualBlockNoBN(num_feat=64, res_scale=0.5).to(cuda)
    x = torch.randn((2, 64, 10, 12), device=cuda, requires_grad=True)
    y = blk(x)
    sd = {k: v.detach().cpu() for k, v in blk.state_dict().items()}
    ref = sr_oracle.residual_block_nobn({f'b.{k}': v for k, v in sd.items()}, 'b', x.detach().cpu(), 0.5)
    assert (y.detach().cpu() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    y.sum().backward()
    assert x.grad is not None and x.grad.shape == x.shape


def test_edsr_cuda_graph_replay_matches_eager(cuda):
    """cuda_graph=True (archs/graphed.py): replayed forward/backward == eager, weights stay live."""
    from basicsr4rs_b200.archs import build_network
    kw = dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=64, num_block=4, upscale=4, res_scale=0.1)
    torch.manual_seed(0)
    eager = build_network(kw).to(cuda).train()
    graphed = build_network(dict(kw, cuda_graph=True, graph_segments=2)).to(cuda).train()
    graphed.load_state_dict(eager.state_dict())
    opt_e = torch.optim.SGD(eager.parameters(), lr=1e-3)
    opt_g = torch.optim.SGD(graphed.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(3)
    for step in range(3):  # step 0 captures, steps 1-2 replay with UPDATED weights
        x = torch.rand((2, 3, 16, 16), generator=g).to(cuda)
        gt = torch.rand((2, 3, 64, 64), generator=g).to(cuda)
        outs = []
        for net, opt in ((eager, opt_e), (graphed, opt_g)):
            opt.zero_grad(set_to_none=True)
            out = net(x)
            ((out - gt)**2).mean().backward()
            outs.append(out.detach().clone())
            opt.step()
        assert torch.allclose(outs[0], outs[1], atol=1e-5), f'step {step}'
        ge, gg = eager.body[1].conv1.weight.grad, graphed.body[1].conv1.weight.grad
        assert torch.allclose(ge, gg, rtol=1e-3, atol=1e-6 * ge.abs().max().item() + 1e-9), f'step {step}'
    # eval / no_grad falls back to eager launches and other shapes still work
    graphed.eval()
    with torch.no_grad():
        y = graphed(torch.rand((1, 3, 9, 11), device=cuda))
    assert y.shape == (1, 3, 36, 44)
    import copy
    copy.deepcopy(graphed)  # EMA copy must not trip over CUDA graph objects


def test_fused_ema_matches_reference_loop(cuda):
    """utils/ema.FusedEMA == BaseModel.model_ema's per-parameter loop (base_model.py:75-82), one launch."""
    import copy
    from basicsr4rs_b200 import _lib as L
    from basicsr4rs_b200.archs import build_network
    from basicsr4rs_b200.utils.ema import FusedEMA
    torch.manual_seed(0)
    net = build_network(dict(type='RCAN', num_in_ch=3, num_out_ch=3, num_feat=64, num_group=2, num_block=3,
                             squeeze_factor=16, upscale=4)).to(cuda)
    ema = copy.deepcopy(net)
    ref = copy.deepcopy(net)
    with torch.no_grad():
        for p_ in net.parameters():
            p_.add_(torch.randn_like(p_) * 0.1)
    fused = FusedEMA(net, ema)
    n0 = L.launch_count
    for decay in (0.999, 0.9):
        fused.step(decay)
        src = dict(net.named_parameters())
        with torch.no_grad():
            for k, p_ in ref.named_parameters():
                p_.data.mul_(decay).add_(src[k].data, alpha=1 - decay)
    assert L.launch_count - n0 == 2
    for (k, a), (_, b) in zip(ema.named_parameters(), ref.named_parameters()):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), k


def test_amp_gradscaler_step_like_srrs_model(cuda):
    """SRRSModel.optimize_parameters (srrs_model.py:24-88) wraps the forward in torch.cuda.amp.autocast and scales the
    loss with a GradScaler.  The srb200 Functions compute in bf16 regardless of autocast and pass loss-scaled
    gradients straight through (bf16 has fp32's exponent range), so after ``scaler.unscale_`` the parameter
    gradients equal those of the plain step and ``scaler.step`` performs the update (no inf/nan skip)."""
    from basicsr4rs_b200.archs import build_network
    opt = dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=64, num_block=2, upscale=2, res_scale=1.0)
    torch.manual_seed(0)
    net_a = build_network(dict(opt)).to(cuda)
    net_b = build_network(dict(opt)).to(cuda)
    net_b.load_state_dict(net_a.state_dict())
    x = torch.rand((2, 3, 16, 16), device=cuda)
    gt = torch.rand((2, 3, 32, 32), device=cuda)
    (net_a(x) - gt).abs().mean().backward()
    optim = torch.optim.Adam(net_b.parameters(), lr=1e-3)
    scaler = torch.amp.GradScaler('cuda', init_scale=65536.0)
    before = [p.detach().clone() for p in net_b.parameters()]
    with torch.autocast('cuda', dtype=torch.float16):
        out = net_b(x)
        loss = (out.float() - gt).abs().mean()
    assert out.dtype == torch.float32  # the archs return the caller's dtype (fp32 image in -> fp32 image out)
    scaler.scale(loss).backward()
    scaler.unscale_(optim)
    for (k, pa), pb in zip(net_a.named_parameters(), net_b.parameters()):
        assert torch.isfinite(pb.grad).all(), k
        rel = ((pa.grad - pb.grad).norm() / (pa.grad.norm() + 1e-12)).item()
        assert rel <= 2e-2, f'{k}: {rel:.3e}'  # identical kernels; only the 2^16 pre-scaling of bf16 intermediates differs
    scaler.step(optim)
    scaler.update()
    assert scaler.get_scale() == 65536.0  # no overflow was detected
    assert any(not torch.equal(b, p.detach()) for b, p in zip(before, net_b.parameters()))
