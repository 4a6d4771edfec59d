"""B200-native twin of the hot-path part of ``basicsr/archs/arch_util.py``.

Same public names, constructor arguments, parameter names / shapes / init order as the reference
(ResidualBlockNoBN :64-88, Upsample :123-142, make_layer :48-61, default_init_weights :17-45,
trunc_normal_ :266-327, to_2tuple :331-345), so state dicts, YAML options and seeded random inits are
interchangeable.  Compute goes through ``basicsr4rs_b200.ops.sr_b200`` (hand-written sm_100a kernels);
there is no PyTorch / CPU fallback: a non-CUDA input raises.

Every block has two entry points:
  ``forward(x)``       -- the reference's nn.Module contract, NCHW float tensor in / out;
  ``forward_nhwc(t)``  -- NHWC bf16 channel-padded tensor in / out, used when the surrounding network is
                          also one of this package's archs (no layout round trips between blocks).
"""
import collections.abc
import math
import warnings
from itertools import repeat

import torch
from torch import nn as nn
from torch.nn import init as init
from torch.nn.modules.batchnorm import _BatchNorm

from ..ops import sr_b200 as ops


def require_cuda(x, who):
    if not x.is_cuda:
        raise RuntimeError(f'{who}: the srb200 hot path is CUDA (sm_100a) only; got a {x.device} tensor. '
                           'There is no CPU fallback by design.')


@torch.no_grad()
def default_init_weights(module_list, scale=1, bias_fill=0, **kwargs):
    """Kaiming-normal init scaled by ``scale`` (reference arch_util.py:17-45; same RNG consumption order)."""
    if not isinstance(module_list, list):
        module_list = [module_list]
    for module in module_list:
        for m in module.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                init.kaiming_normal_(m.weight, **kwargs)
                m.weight.data *= scale
                if m.bias is not None:
                    m.bias.data.fill_(bias_fill)
            elif isinstance(m, _BatchNorm):
                init.constant_(m.weight, 1)
                if m.bias is not None:
                    m.bias.data.fill_(bias_fill)


def make_layer(basic_block, num_basic_block, **kwarg):
    """nn.Sequential of ``num_basic_block`` blocks (reference arch_util.py:48-61)."""
    return nn.Sequential(*[basic_block(**kwarg) for _ in range(num_basic_block)])


def nchw_roundtrip(fn, x, who):
    """Run an NHWC-bf16 block on an NCHW float tensor (stand-alone use of a block)."""
    require_cuda(x, who)
    c = x.shape[1]
    with torch.cuda.device(x.device):
        t = ops.image_to_nhwc(x, None, 1.0, ops.pad64(c))
        y = fn(t)
        out = _NHWCToImage.apply(y, c)
    return out.to(x.dtype)


class _NHWCToImage(torch.autograd.Function):

    @staticmethod
    def forward(ctx, t, c):
        ctx.c_pad = t.shape[-1]
        return ops.raw.nhwc_to_nchw(t.contiguous(), c)

    @staticmethod
    def backward(ctx, g):
        return ops.raw.nchw_to_nhwc(g.contiguous().float(), ctx.c_pad), None


class ResidualBlockNoBN(nn.Module):
    """Residual block without BN: ``x + conv2(relu(conv1(x))) * res_scale`` (reference arch_util.py:64-88)."""

    def __init__(self, num_feat=64, res_scale=1, pytorch_init=False):
        super(ResidualBlockNoBN, self).__init__()
        self.res_scale = res_scale
        self.conv1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        self.relu = nn.ReLU(inplace=True)

        if not pytorch_init:
            default_init_weights([self.conv1, self.conv2], 0.1)

    def forward_nhwc(self, t):
        return ops.res_block_nobn(t, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                                  self.res_scale)

    def forward(self, x):
        return nchw_roundtrip(self.forward_nhwc, x, 'ResidualBlockNoBN')


class Upsample(nn.Sequential):
    """[conv F->4F, PixelShuffle(2)] x log2(scale) or conv F->9F + PixelShuffle(3) (reference
    arch_util.py:123-142).  The PixelShuffle is fused into the conv's epilogue store."""

    def __init__(self, scale, num_feat):
        m = []
        if (scale & (scale - 1)) == 0:  # scale = 2^n
            for _ in range(int(math.log(scale, 2))):
                m.append(nn.Conv2d(num_feat, 4 * num_feat, 3, 1, 1))
                m.append(nn.PixelShuffle(2))
        elif scale == 3:
            m.append(nn.Conv2d(num_feat, 9 * num_feat, 3, 1, 1))
            m.append(nn.PixelShuffle(3))
        else:
            raise ValueError(f'scale {scale} is not supported. Supported scales: 2^n and 3.')
        super(Upsample, self).__init__(*m)

    def forward_nhwc(self, t):
        mods = list(self)
        for conv, shuffle in zip(mods[0::2], mods[1::2]):
            t = ops.conv_nhwc(t, conv.weight, conv.bias, shuffle_r=shuffle.upscale_factor)
        return t

    def forward(self, x):
        return nchw_roundtrip(self.forward_nhwc, x, 'Upsample')


def _no_grad_trunc_normal_(tensor, mean, std, a, b):
    # inverse-CDF sampling of a truncated normal (same recipe and RNG use as reference arch_util.py:266-300)
    def norm_cdf(x):
        return (1. + math.erf(x / math.sqrt(2.))) / 2.

    if (mean < a - 2 * std) or (mean > b + 2 * std):
        warnings.warn(
            'mean is more than 2 std from [a, b] in nn.init.trunc_normal_. '
            'The distribution of values may be incorrect.', stacklevel=2)

    with torch.no_grad():
        low = norm_cdf((a - mean) / std)
        up = norm_cdf((b - mean) / std)
        tensor.uniform_(2 * low - 1, 2 * up - 1)
        tensor.erfinv_()
        tensor.mul_(std * math.sqrt(2.))
        tensor.add_(mean)
        tensor.clamp_(min=a, max=b)
        return tensor


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """Truncated-normal init (reference arch_util.py:303-327)."""
    return _no_grad_trunc_normal_(tensor, mean, std, a, b)


def _ntuple(n):

    def parse(x):
        if isinstance(x, collections.abc.Iterable):
            return x
        return tuple(repeat(x, n))

    return parse


to_1tuple = _ntuple(1)
to_2tuple = _ntuple(2)
to_3tuple = _ntuple(3)
to_4tuple = _ntuple(4)
to_ntuple = _ntuple
