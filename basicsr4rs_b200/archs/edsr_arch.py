"""EDSR on the srb200 kernels -- drop-in for the reference's ``basicsr/archs/edsr_arch.py`` (:8-61):
same registry name, constructor / YAML keys, parameter names, shapes and init order."""
import torch
from torch import nn as nn

from ..ops import sr_b200 as ops
from ..utils.registry import ARCH_REGISTRY
from .arch_util import ResidualBlockNoBN, Upsample, make_layer, require_cuda
from .graphed import GRAPHS, ArchMixin, GraphedSegments, Segment


_MeanShiftMixin = ArchMixin  # ``self.mean`` stays a plain tensor attribute (edsr_arch.py:42,51)


@ARCH_REGISTRY.register()
class EDSR(ArchMixin, nn.Module):
    """EDSR: mean-shift -> conv_first -> num_block x ResidualBlockNoBN -> conv_after_body + skip -> Upsample
    -> conv_last -> inverse shift.

    Args (identical to the reference): num_in_ch, num_out_ch, num_feat=64, num_block=16, upscale=4,
    res_scale=1, img_range=255., rgb_mean=(0.4488, 0.4371, 0.4040).
    Optional, new (defaults keep every existing YAML valid): cuda_graph=False replays the training
    forward/backward as ``graph_segments`` CUDA graphs per input shape (see archs/graphed.py).
    """

    def __init__(self,
                 num_in_ch,
                 num_out_ch,
                 num_feat=64,
                 num_block=16,
                 upscale=4,
                 res_scale=1,
                 img_range=255.,
                 rgb_mean=(0.4488, 0.4371, 0.4040),
                 cuda_graph=False,
                 graph_segments=4,
                 graph_input_shape=None,
                 compute_dtype='bf16',
                 flat_grads=False):
        super(EDSR, self).__init__()
        if compute_dtype not in ('bf16', 'fp32'):
            raise ValueError(f"compute_dtype must be 'bf16' or 'fp32', got {compute_dtype!r}")
        # 'fp32': evaluation in fp32-class arithmetic (error-compensated bf16 operands, ops/sr_b200/fp32_mode.py) --
        # north_star's "TF32/fp32 mode", max-abs <= 1e-4 against the reference; training always runs the bf16 path
        self.compute_dtype = compute_dtype
        self.cuda_graph = cuda_graph
        self.graph_segments = max(1, int(graph_segments))
        self.graph_input_shape = graph_input_shape  # e.g. [16, 3, 48, 48]: capture on .to(device), before DDP
        self.flat_grads = bool(flat_grads)  # gradients land in ONE flat buffer (utils/flat_ddp.py) -- for FlatDDP

        self.img_range = img_range
        self.mean = torch.Tensor(rgb_mean).view(1, 3, 1, 1)

        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = make_layer(ResidualBlockNoBN, num_block, num_feat=num_feat, res_scale=res_scale, pytorch_init=True)
        self.conv_after_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.upsample = Upsample(upscale, num_feat)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)

    # ---- pieces (shared by the eager path and the captured segments)
    def _head(self, x):
        mean = self._device_mean(x)
        # (x - mean) * img_range fused into the NCHW fp32 -> NHWC bf16 entry (edsr_arch.py:53)
        t = ops.image_to_nhwc(x, mean, self.img_range, ops.pad64(x.shape[1]))
        return ops.conv_nhwc(t, self.conv_first.weight, self.conv_first.bias)

    def _tail(self, res, first):
        # conv_after_body + `res += x` in one epilogue (edsr_arch.py:55-56)
        res = ops.conv_nhwc(res, self.conv_after_body.weight, self.conv_after_body.bias, residual=first)
        up = self.upsample.forward_nhwc(res)
        # conv_last + `x / img_range + mean` + NCHW fp32 exit (edsr_arch.py:58-59)
        return ops.conv_to_image(up, self.conv_last.weight, self.conv_last.bias, 1.0 / self.img_range,
                                 self._mean_dev)

    def _build_segments(self):
        blocks = list(self.body)
        n = min(self.graph_segments, max(1, len(blocks)))
        cuts = [round(i * len(blocks) / n) for i in range(n + 1)]
        groups = [blocks[cuts[i]:cuts[i + 1]] for i in range(n)]

        def run(group):
            def fn(t):
                for blk in group:
                    t = blk.forward_nhwc(t)
                return t
            return fn

        def head_fn(x):
            first = self._head(x)
            return run(groups[0])(first), first

        segs = [Segment(head_fn, [self.conv_first] + groups[0])]
        segs += [Segment(run(g), g) for g in groups[1:]]
        segs.append(Segment(self._tail, [self.conv_after_body, self.upsample, self.conv_last]))
        return segs

    def _forward_fp32(self, x):
        """edsr_arch.py:50-61 in fp32 mode (no autograd): fp32 NHWC activations, split-bf16 tap-GEMMs."""
        from ..ops.sr_b200 import fp32_mode as f32
        from .. import _lib as L
        mean = self._device_mean(x)
        t = f32.image_to_nhwc32(x, mean, self.img_range, ops.pad64(x.shape[1]))
        first = f32.conv(t, self.conv_first.weight, self.conv_first.bias)
        res = first
        for blk in self.body:
            h = f32.conv(res, blk.conv1.weight, blk.conv1.bias, act=L.ACT_RELU)
            res = f32.conv(h, blk.conv2.weight, blk.conv2.bias, alpha=blk.res_scale, residual32=res)
        res = f32.conv(res, self.conv_after_body.weight, self.conv_after_body.bias, residual32=first)
        mods = list(self.upsample)
        for conv, shuffle in zip(mods[0::2], mods[1::2]):
            res = f32.conv(res, conv.weight, conv.bias, shuffle_r=shuffle.upscale_factor)
        return f32.conv_to_image(res, self.conv_last.weight, self.conv_last.bias, 1.0 / self.img_range, mean)

    def _forward(self, x):
        require_cuda(x, 'EDSR')
        if self.compute_dtype == 'fp32' and not torch.is_grad_enabled():
            out = self._forward_fp32(x)
            return out if out.dtype == x.dtype else out.to(x.dtype)
        if self.cuda_graph and self.training and torch.is_grad_enabled():
            graphs = GRAPHS.get(self)
            if graphs is None:
                nseg = len(self._build_segments())

                def wire(call, inp):
                    res, first = call(0, inp)
                    for i in range(1, nseg - 1):
                        res = call(i, res)
                    return call(nseg - 1, res, first)

                self._device_mean(x)
                graphs = GRAPHS[self] = GraphedSegments(self._build_segments, wire, book=self._pack_book(), pdl=True)
            out = graphs(x.contiguous().float(), True)
            return out if out.dtype == x.dtype else out.to(x.dtype)
        first = self._head(x)
        res = first
        for block in self.body:
            res = block.forward_nhwc(res)
        out = self._tail(res, first)
        return out if out.dtype == x.dtype else out.to(x.dtype)
