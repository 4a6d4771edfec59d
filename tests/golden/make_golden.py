"""Generate golden vectors from the UNMODIFIED reference archs (run in the build container).

    python tests/golden/make_golden.py        # needs /root/reference (or BASICSR_REF_ROOT)

For each case the reference nn.Module is built, its state dict is overwritten by the by-name seeded
filler (oracle.sr_oracle.fill_state_dict_, so no weights need to be shipped), and
input / output / L1 loss value / selected gradients of the smooth loss ``mean((out-gt)^2)`` are saved to ``tests/golden/<case>.pt`` (fp32, CPU).
The GPU box has no /root/reference: tests read only these fixtures.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.sr_oracle import fill_state_dict_  # noqa: E402

CASES = {
    # name: (arch, ctor kwargs, input shape)
    'edsr_f64_b2_x4': ('EDSR', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=2, upscale=4, res_scale=0.1,
                                    img_range=255.), (2, 3, 12, 16)),
    'edsr_f64_b1_x3': ('EDSR', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=1, upscale=3, res_scale=1.0,
                                    img_range=255.), (1, 3, 10, 9)),
    'edsr_f256_b2_x2': ('EDSR', dict(num_in_ch=3, num_out_ch=3, num_feat=256, num_block=2, upscale=2, res_scale=0.1,
                                     img_range=255.), (1, 3, 16, 16)),
    'rcan_f64_g2_b2_x4': ('RCAN', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_group=2, num_block=2,
                                       squeeze_factor=16, upscale=4, res_scale=1.0, img_range=255.), (2, 3, 12, 12)),
    'swinir_c180_d2x2_x4': ('SwinIR', dict(upscale=4, in_chans=3, img_size=16, window_size=8, img_range=1.,
                                           depths=[2, 2], embed_dim=180, num_heads=[6, 6], mlp_ratio=2,
                                           upsampler='pixelshuffle', resi_connection='1conv'), (2, 3, 16, 24)),
    'swinir_c60_d2_x2': ('SwinIR', dict(upscale=2, in_chans=3, img_size=16, window_size=8, img_range=1., depths=[2],
                                        embed_dim=60, num_heads=[6], mlp_ratio=2, upsampler='pixelshuffle',
                                        resi_connection='1conv'), (1, 3, 16, 16)),
    # the other reconstruction branches of SwinIR (swinir_arch.py:901-918)
    'swinir_c60_d2_direct_x2': ('SwinIR', dict(upscale=2, in_chans=3, img_size=16, window_size=8, img_range=1.,
                                               depths=[2], embed_dim=60, num_heads=[6], mlp_ratio=2,
                                               upsampler='pixelshuffledirect', resi_connection='1conv'),
                                (1, 3, 16, 24)),
    'swinir_c60_d2_nearest_x4': ('SwinIR', dict(upscale=4, in_chans=3, img_size=16, window_size=8, img_range=1.,
                                                depths=[2], embed_dim=60, num_heads=[6], mlp_ratio=2,
                                                upsampler='nearest+conv', resi_connection='3conv'), (1, 3, 16, 16)),
    # denoising / JPEG-artifact branch: no upsampler, residual image (swinir_arch.py:915-918)
    'swinir_c60_d2_denoise_x1': ('SwinIR', dict(upscale=1, in_chans=3, img_size=16, window_size=8, img_range=1.,
                                                depths=[2], embed_dim=60, num_heads=[6], mlp_ratio=2, upsampler='',
                                                resi_connection='1conv'), (1, 3, 16, 24)),
    # the x3 (single PixelShuffle(3)) and x8 (three PixelShuffle(2) stages) Upsample heads (arch_util.py:119-138)
    'rcan_f64_g1_b2_sq8_x3': ('RCAN', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_group=1, num_block=2,
                                           squeeze_factor=8, upscale=3, res_scale=1.0, img_range=255.), (1, 3, 10, 14)),
    'edsr_f64_b1_x8': ('EDSR', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=1, upscale=8, res_scale=1.0,
                                    img_range=255.), (1, 3, 8, 8)),
    # the fork's remote-sensing shape: 6-wide windows, 4 input bands (train_SwinIR_L2S288_scratch.yml:43-54)
    'swinir_c60_ws6_in4_x2': ('SwinIR', dict(upscale=2, in_chans=4, img_size=12, window_size=6, img_range=1.,
                                             depths=[2], embed_dim=60, num_heads=[6], mlp_ratio=2,
                                             upsampler='pixelshuffle', resi_connection='1conv'), (1, 4, 12, 18)),
}

GRAD_KEYS = {
    'EDSR': ['conv_first.weight', 'conv_first.bias', 'body.0.conv1.weight', 'body.0.conv2.bias', 'upsample.0.weight',
             'upsample.0.bias', 'conv_last.weight', 'conv_last.bias'],
    'RCAN': ['conv_first.weight', 'body.0.residual_group.0.rcab.0.weight', 'body.0.residual_group.0.rcab.2.bias',
             'body.0.residual_group.1.rcab.3.attention.1.weight', 'body.0.residual_group.1.rcab.3.attention.3.bias',
             'body.1.conv.weight', 'conv_last.bias'],
    'SwinIR': ['conv_first.weight', 'patch_embed.norm.weight', 'layers.0.residual_group.blocks.0.norm1.bias',
               'layers.0.residual_group.blocks.0.attn.qkv.weight', 'layers.0.residual_group.blocks.0.attn.qkv.bias',
               'layers.0.residual_group.blocks.1.attn.relative_position_bias_table',
               'layers.0.residual_group.blocks.1.attn.proj.weight', 'layers.0.residual_group.blocks.1.mlp.fc1.weight',
               'layers.0.residual_group.blocks.1.mlp.fc2.bias', 'layers.0.conv.weight', 'norm.weight',
               'conv_after_body.bias', 'conv_before_upsample.0.weight', 'upsample.0.weight', 'conv_last.weight',
               'upsample.0.bias', 'conv_up1.weight', 'conv_up2.bias', 'conv_hr.weight', 'conv_after_body.0.weight',
               'conv_after_body.2.bias', 'conv_last.bias'],
}


def compress(g, limit=8192, nsample=2048):
    """Small gradients are stored whole; large ones as a strided sample + two checksums."""
    if g.numel() <= limit:
        return g.clone()
    stride = g.numel() // nsample
    flat = g.flatten()
    return {'stride': stride, 'sample': flat[::stride].clone(), 'sum': flat.double().sum(),
            'l2': flat.double().norm()}


def run_case(ref, name):
    arch, kwargs, in_shape = CASES[name]
    torch.manual_seed(0)
    net = getattr(ref, arch)(**kwargs)
    sd = net.state_dict()
    fill_state_dict_(sd)
    net.load_state_dict(sd, strict=True)
    net.eval()  # SwinIR: drop_path off
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(in_shape, generator=g)
    out = net(x)
    gt = torch.rand(out.shape, generator=g)
    loss = (out - gt).abs().mean()  # L1Loss value (losses/basic_loss.py:28)
    # gradients are taken of a SMOOTH loss: d L1 / d out = sign(out - gt) / N flips wherever a bf16-path output
    # crosses gt, which would make gradient parity measure the loss' discontinuity instead of the kernels
    ((out - gt)**2).mean().backward()
    params = dict(net.named_parameters())
    keys = [k for k in GRAD_KEYS[arch] if k in params]
    # calibration: the gradient error PyTorch's own CPU bf16 autocast of the SAME reference net makes on each key.
    # A bf16 tensor-core path cannot be expected to beat it; tests allow max(fixed bar, 2 x this)
    fp32_grads = {k: params[k].grad.clone() for k in keys}
    net.zero_grad()
    with torch.autocast('cpu', dtype=torch.bfloat16):
        out_ac = net(x)
    ((out_ac.float() - gt)**2).mean().backward()
    autocast_rel = {k: ((params[k].grad - fp32_grads[k]).norm() / (fp32_grads[k].norm() + 1e-12)).item() for k in keys}
    for k in keys:
        params[k].grad = fp32_grads[k]
    fixture = {
        'arch': arch, 'kwargs': kwargs, 'x': x, 'gt': gt, 'out': out.detach(), 'loss': loss.detach(),
        'grads': {k: compress(params[k].grad) for k in keys}, 'grad_loss': 'mse', 'autocast_rel': autocast_rel,
        'state_keys': list(sd.keys()),
        'state_shapes': {k: tuple(v.shape) for k, v in sd.items()},
        'n_params': sum(p.numel() for p in net.parameters()),
    }
    torch.save(fixture, os.path.join(HERE, name + '.pt'))
    print(f'{name}: out {tuple(out.shape)} loss {loss.item():.6f} params {fixture["n_params"]}')


def known_answers(ref):
    """Known-answer facts about the SwinIR index buffers."""
    torch.manual_seed(0)
    net = ref.SwinIR(upscale=4, in_chans=3, img_size=64, window_size=8, img_range=1., depths=[2], embed_dim=180,
                     num_heads=[6], mlp_ratio=2, upsampler='pixelshuffle', resi_connection='1conv')
    blk = net.layers[0].residual_group.blocks[1]
    fx = {
        'relative_position_index_ws8': blk.attn.relative_position_index.clone(),
        'attn_mask_64x64_ws8_s4_packed': torch.from_numpy(__import__('numpy').packbits((blk.attn_mask != 0).numpy())),
        'attn_mask_values': torch.unique(blk.attn_mask),
        'attn_mask_24x40_ws8_s4': blk.calculate_mask((24, 40)).to(torch.int8),
        'scale_d30': torch.tensor(blk.attn.scale, dtype=torch.float64),
    }
    torch.save(fx, os.path.join(HERE, 'swinir_known_answers.pt'))
    print('known answers saved')


if __name__ == '__main__':
    ref = ref_shim.load_reference_archs()
    only = sys.argv[1:]  # optional: regenerate only the named cases (existing fixtures stay byte-identical)
    for case in CASES:
        if not only or case in only:
            run_case(ref, case)
    if not only:
        known_answers(ref)
