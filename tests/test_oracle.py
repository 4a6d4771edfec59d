"""Pin ``oracle/sr_oracle.py`` (CPU): against the committed golden fixtures everywhere, and against
the live unmodified reference where /root/reference is mounted (build container only)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_shim, sr_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, '*_x[1-8].pt')))


def oracle_forward(fx, sd, x):
    kw = fx['kwargs']
    if fx['arch'] == 'EDSR':
        return sr_oracle.edsr_forward(sd, x, num_block=kw['num_block'], upscale=kw['upscale'],
                                      res_scale=kw['res_scale'], img_range=kw['img_range'])
    if fx['arch'] == 'RCAN':
        return sr_oracle.rcan_forward(sd, x, num_group=kw['num_group'], num_block=kw['num_block'],
                                      upscale=kw['upscale'], res_scale=kw['res_scale'], img_range=kw['img_range'])
    return sr_oracle.swinir_forward(sd, x, embed_dim=kw['embed_dim'], depths=kw['depths'], num_heads=kw['num_heads'],
                                    window_size=kw['window_size'], upscale=kw['upscale'], img_range=kw['img_range'],
                                    in_chans=kw.get('in_chans', 3), upsampler=kw.get('upsampler', 'pixelshuffle'),
                                    resi_connection=kw.get('resi_connection', '1conv'))


def state_dict_from_fixture(fx):
    sd = {k: torch.zeros(s) for k, s in fx['state_shapes'].items() if not k.endswith(('relative_position_index',
                                                                                      'attn_mask'))}
    return sr_oracle.fill_state_dict_(sd)


def check_grad(got, want, tol):
    if isinstance(want, dict):
        flat = got.flatten()
        assert torch.allclose(flat[::want['stride']], want['sample'], rtol=tol, atol=tol * want['sample'].abs().max())
        assert abs(flat.double().norm().item() - want['l2'].item()) <= tol * want['l2'].item()
    else:
        assert torch.allclose(got, want, rtol=tol, atol=tol * want.abs().max().item())


def test_fixtures_present():
    assert len(CASES) >= 6


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_golden(case):
    fx = torch.load(os.path.join(GOLDEN, case + '.pt'), weights_only=False)
    sd = state_dict_from_fixture(fx)
    for v in sd.values():
        v.requires_grad_(True)
    out = oracle_forward(fx, sd, fx['x'])
    assert out.shape == fx['out'].shape
    assert (out - fx['out']).abs().max().item() <= 1e-4  # fp32 bar of north_star
    loss = (out - fx['gt']).abs().mean()
    assert abs(loss.item() - fx['loss'].item()) <= 1e-6
    ((out - fx['gt'])**2).mean().backward()
    for k, want in fx['grads'].items():
        check_grad(sd[k].grad, want, 2e-4)
    assert sum(v.numel() for k, v in sd.items()) == fx['n_params']


def test_known_answers():
    ka = torch.load(os.path.join(GOLDEN, 'swinir_known_answers.pt'), weights_only=False)
    assert torch.equal(sr_oracle.relative_position_index(8), ka['relative_position_index_ws8'])
    assert int(ka['relative_position_index_ws8'].max()) == 224
    m = sr_oracle.calculate_mask(64, 64, 8, 4)
    assert m.shape == (64, 64, 64)
    assert torch.equal(torch.unique(m), ka['attn_mask_values']) and ka['attn_mask_values'].tolist() == [-100.0, 0.0]
    assert np.array_equal(np.packbits((m != 0).numpy()), ka['attn_mask_64x64_ws8_s4_packed'].numpy())
    assert torch.equal(sr_oracle.calculate_mask(24, 40, 8, 4).to(torch.int8), ka['attn_mask_24x40_ws8_s4'])
    assert abs(ka['scale_d30'].item() - 30**-0.5) < 1e-15


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not mounted (GPU box)')
@pytest.mark.parametrize('arch,kwargs,shape', [
    ('EDSR', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=3, upscale=2, res_scale=0.1, img_range=255.),
     (1, 3, 9, 11)),
    ('RCAN', dict(num_in_ch=3, num_out_ch=3, num_feat=32, num_group=2, num_block=3, squeeze_factor=8, upscale=3,
                  res_scale=1.0, img_range=255.), (2, 3, 8, 8)),
    ('SwinIR', dict(upscale=2, in_chans=3, img_size=24, window_size=8, img_range=1., depths=[3, 2], embed_dim=60,
                    num_heads=[6, 6], mlp_ratio=2, upsampler='pixelshuffle', resi_connection='1conv'), (1, 3, 24, 16)),
])
def test_oracle_matches_live_reference(arch, kwargs, shape):
    """Default (seeded) init of the reference itself -- not the by-name filler -- and a non-square input
    (SwinIR: x_size != input_resolution, so the reference rebuilds the mask, swinir_arch.py:305-306)."""
    ref = ref_shim.load_reference_archs()
    torch.manual_seed(7)
    net = getattr(ref, arch)(**kwargs).eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.rand(shape)
    with torch.no_grad():
        want = net(x)
        got = oracle_forward({'arch': arch, 'kwargs': kwargs}, sd, x)
    assert (got - want).abs().max().item() <= 1e-5


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not available')
def test_oracle_droppath_and_ape_match_live_reference_training_mode():
    """TRAINING mode with stochastic depth (the configuration bench.py times for SwinIR) and ``ape=True``: the oracle,
    fed the per-block keep masks drawn in the reference's own RNG order (``drop_path_masks``), reproduces the live
    reference's output and gradients.  This pins the oracle leg the GPU DropPath parity test relies on."""
    ref = ref_shim.load_reference_archs()
    kw = dict(upscale=2, in_chans=3, img_size=16, window_size=8, img_range=1., depths=[2, 2], embed_dim=60,
              num_heads=[6, 6], mlp_ratio=2, upsampler='pixelshuffle', resi_connection='1conv', drop_path_rate=0.5,
              ape=True)
    torch.manual_seed(11)
    net = ref.SwinIR(**kw).train()
    with torch.no_grad():
        net.absolute_pos_embed.normal_(std=0.1)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in net.state_dict().items()}
    x = torch.rand((4, 3, 16, 16))
    torch.manual_seed(123)
    want = net(x)
    torch.manual_seed(123)
    masks = sr_oracle.drop_path_masks(4, kw['depths'], kw['drop_path_rate'])
    assert any((m[0] == 0).any() or (m[1] == 0).any() for m in masks.values())  # some branch really was dropped
    got = sr_oracle.swinir_forward(sd, x, embed_dim=60, depths=[2, 2], num_heads=[6, 6], window_size=8, upscale=2,
                                   img_range=1., drop_masks=masks, ape=True)
    assert (got - want).abs().max().item() <= 1e-5
    (want**2).mean().backward()
    (got**2).mean().backward()
    for k, p in net.named_parameters():
        assert torch.allclose(sd[k].grad, p.grad, rtol=1e-4, atol=1e-6 * p.grad.abs().max().item() + 1e-9), k
