"""RCAN on the srb200 kernels -- drop-in for the reference's ``basicsr/archs/rcan_arch.py`` (:8-135):
same class names, constructor / YAML keys, parameter names, shapes and init order."""
import torch
from torch import nn as nn

from ..ops import sr_b200 as ops
from ..utils.registry import ARCH_REGISTRY
from .arch_util import Upsample, make_layer, nchw_roundtrip, require_cuda
from .graphed import ArchMixin, Segment, chain_wire, graphed_forward, split_even


class ChannelAttention(nn.Module):
    """x * sigmoid(conv1x1(relu(conv1x1(avgpool(x)))))  (reference rcan_arch.py:8-24)."""

    def __init__(self, num_feat, squeeze_factor=16):
        super(ChannelAttention, self).__init__()
        self.attention = nn.Sequential(
            nn.AdaptiveAvgPool2d(1), nn.Conv2d(num_feat, num_feat // squeeze_factor, 1, padding=0),
            nn.ReLU(inplace=True), nn.Conv2d(num_feat // squeeze_factor, num_feat, 1, padding=0), nn.Sigmoid())

    def fc_params(self):
        a, b = self.attention[1], self.attention[3]
        return a.weight, a.bias, b.weight, b.bias

    def forward_nhwc(self, t):
        return ops.channel_attention(t, *self.fc_params())

    def forward(self, x):
        """Reference contract (rcan_arch.py:22-24): ``x * self.attention(x)`` on an NCHW float tensor."""
        return nchw_roundtrip(self.forward_nhwc, x, 'ChannelAttention')


class RCAB(nn.Module):
    """Residual channel attention block (reference rcan_arch.py:27-46), one fused autograd function."""

    def __init__(self, num_feat, squeeze_factor=16, res_scale=1):
        super(RCAB, self).__init__()
        self.res_scale = res_scale

        self.rcab = nn.Sequential(
            nn.Conv2d(num_feat, num_feat, 3, 1, 1), nn.ReLU(True), nn.Conv2d(num_feat, num_feat, 3, 1, 1),
            ChannelAttention(num_feat, squeeze_factor))

    def forward_nhwc(self, t, t32=None):
        """``t32``: optional fp32 twin of the bf16 stream ``t``; returns (y, y32) when given."""
        c1, c2, ca = self.rcab[0], self.rcab[2], self.rcab[3]
        return ops.rcab(t, c1.weight, c1.bias, c2.weight, c2.bias, *ca.fc_params(), self.res_scale, x32=t32)

    def forward(self, x):
        return nchw_roundtrip(self.forward_nhwc, x, 'RCAB')


class ResidualGroup(nn.Module):
    """num_block x RCAB, conv, + skip (reference rcan_arch.py:49-68); the skip add rides the conv epilogue."""

    def __init__(self, num_feat, num_block, squeeze_factor=16, res_scale=1):
        super(ResidualGroup, self).__init__()

        self.residual_group = make_layer(
            RCAB, num_block, num_feat=num_feat, squeeze_factor=squeeze_factor, res_scale=res_scale)
        self.conv = nn.Conv2d(num_feat, num_feat, 3, 1, 1)

    def forward_nhwc(self, t, t32=None):
        if t32 is None:
            r = t
            for block in self.residual_group:
                r = block.forward_nhwc(r)
            return ops.conv_nhwc(r, self.conv.weight, self.conv.bias, residual=t)
        r, r32 = t, t32
        for block in self.residual_group:
            r, r32 = block.forward_nhwc(r, r32)
        return ops.conv_nhwc(r, self.conv.weight, self.conv.bias, residual=t, residual32=t32, want_f32=True)

    def forward(self, x):
        return nchw_roundtrip(self.forward_nhwc, x, 'ResidualGroup')


@ARCH_REGISTRY.register()
class RCAN(ArchMixin, nn.Module):
    """Residual Channel Attention Network (reference rcan_arch.py:71-135).

    Args (identical to the reference): num_in_ch, num_out_ch, num_feat=64, num_group=10, num_block=16,
    squeeze_factor=16, upscale=4, res_scale=1, img_range=255., rgb_mean=(0.4488, 0.4371, 0.4040).
    """
    graph_pdl = True  # back-to-back GEMM chain: captured with programmatic dependent launch (archs/graphed.py)

    def __init__(self,
                 num_in_ch,
                 num_out_ch,
                 num_feat=64,
                 num_group=10,
                 num_block=16,
                 squeeze_factor=16,
                 upscale=4,
                 res_scale=1,
                 img_range=255.,
                 rgb_mean=(0.4488, 0.4371, 0.4040),
                 cuda_graph=False,
                 graph_segments=5,
                 graph_input_shape=None,
                 compute_dtype='bf16',
                 flat_grads=False):
        super(RCAN, self).__init__()
        if compute_dtype not in ('bf16', 'fp32'):
            raise ValueError(f"compute_dtype must be 'bf16' or 'fp32', got {compute_dtype!r}")
        self.compute_dtype = compute_dtype  # 'fp32': evaluation in fp32-class arithmetic (ops/sr_b200/fp32_mode.py)
        self.cuda_graph = cuda_graph          # optional: replay training fwd/bwd from CUDA graphs (archs/graphed.py)
        self.graph_segments = graph_segments
        self.graph_input_shape = graph_input_shape
        self.flat_grads = bool(flat_grads)  # gradients land in ONE flat buffer (utils/flat_ddp.py) -- for FlatDDP

        self.img_range = img_range
        self.mean = torch.Tensor(rgb_mean).view(1, 3, 1, 1)

        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = make_layer(
            ResidualGroup,
            num_group,
            num_feat=num_feat,
            num_block=num_block,
            squeeze_factor=squeeze_factor,
            res_scale=res_scale)
        self.conv_after_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.upsample = Upsample(upscale, num_feat)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)

    # The skip stream is carried twice: bf16 (differentiable, feeds the tensor cores) and fp32 (what the skip
    # adds read and write).  200 stacked res_scale=1 additions in bf16 alone drift past the 1e-2 output bar
    # (1.16e-2 for bf16 autocast of the reference itself).
    def _head(self, x):
        mean = self._device_mean(x)
        t = ops.image_to_nhwc(x, mean, self.img_range, ops.pad64(x.shape[1]))
        return ops.conv_nhwc(t, self.conv_first.weight, self.conv_first.bias, want_f32=True)

    def _tail(self, res, res32, first, first32):
        res = ops.conv_nhwc(res, self.conv_after_body.weight, self.conv_after_body.bias, residual=first,
                            residual32=first32)
        up = self.upsample.forward_nhwc(res)
        return ops.conv_to_image(up, self.conv_last.weight, self.conv_last.bias, 1.0 / self.img_range,
                                 self._mean_dev)

    def _build_segments(self):
        groups = split_even(list(self.body), self.graph_segments)

        def run(gs):
            def fn(res, res32):
                for grp in gs:
                    res, res32 = grp.forward_nhwc(res, res32)
                return res, res32
            return fn

        def head_fn(x):
            first, first32 = self._head(x)
            res, res32 = run(groups[0])(first, first32)
            return res, res32, first, first32

        segs = [Segment(head_fn, [self.conv_first] + groups[0])]
        segs += [Segment(run(g), g) for g in groups[1:]]
        segs.append(Segment(self._tail, [self.conv_after_body, self.upsample, self.conv_last]))
        for seg, g in zip(segs, groups):  # the per-sample pool sums of every RCAB of the segment: one fill
            seg.arena_floats = self._pool_floats(g)
        return segs

    def _pool_floats(self, groups, batch=64):
        """fp32 scratch the channel-attention pooling of ``groups`` needs (B x C per RCAB, B <= ``batch``)."""
        n_rcab = sum(len(grp.residual_group) for grp in groups)
        return n_rcab * (batch * self.conv_first.weight.shape[0] + 8)

    def _forward_fp32(self, x):
        """rcan_arch.py:124-135 in fp32 mode (no autograd): split-bf16 tap-GEMMs with fp32 accumulation for the 411 convs;
        the channel attention (a [B, C] pool, two tiny FCs, one multiply-add pass) in plain fp32."""
        from ..ops.sr_b200 import fp32_mode as f32
        from .. import _lib as L
        mean = self._device_mean(x)
        t = f32.image_to_nhwc32(x, mean, self.img_range, ops.pad64(x.shape[1]))
        first = f32.conv(t, self.conv_first.weight, self.conv_first.bias)
        res = first
        for group in self.body:
            r = res
            for blk in group.residual_group:
                c1, c2, ca = blk.rcab[0], blk.rcab[2], blk.rcab[3]
                u = f32.conv(f32.conv(r, c1.weight, c1.bias, act=L.ACT_RELU), c2.weight, c2.bias)
                wa1, ba1, wa2, ba2 = ca.fc_params()
                _, s = ops.raw.ca_fc(u.mean(dim=(1, 2)).contiguous(), wa1.detach().contiguous(), ba1.detach(),
                                     wa2.detach().contiguous(), ba2.detach())
                r = r + blk.res_scale * u * s.view(s.shape[0], 1, 1, -1)
            res = f32.conv(r, group.conv.weight, group.conv.bias, residual32=res)
        res = f32.conv(res, self.conv_after_body.weight, self.conv_after_body.bias, residual32=first)
        mods = list(self.upsample)
        for conv, shuffle in zip(mods[0::2], mods[1::2]):
            res = f32.conv(res, conv.weight, conv.bias, shuffle_r=shuffle.upscale_factor)
        return f32.conv_to_image(res, self.conv_last.weight, self.conv_last.bias, 1.0 / self.img_range, mean)

    def _forward(self, x):
        require_cuda(x, 'RCAN')
        if self.compute_dtype == 'fp32' and not torch.is_grad_enabled():
            out = self._forward_fp32(x)
            return out if out.dtype == x.dtype else out.to(x.dtype)
        if self.cuda_graph and self.training and torch.is_grad_enabled():
            self._device_mean(x)
            nseg = len(split_even(list(self.body), self.graph_segments)) + 1
            out = graphed_forward(self, x, self._build_segments, chain_wire(nseg, carry=2))
            return out if out.dtype == x.dtype else out.to(x.dtype)
        with ops.raw.zero_arena(x.device, self._pool_floats(list(self.body), batch=x.shape[0])):
            first, first32 = self._head(x)
            res, res32 = first, first32
            for group in self.body:
                res, res32 = group.forward_nhwc(res, res32)
            out = self._tail(res, res32, first, first32)
        return out if out.dtype == x.dtype else out.to(x.dtype)
