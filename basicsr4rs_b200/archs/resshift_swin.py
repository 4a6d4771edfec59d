"""NCHW twin of the window attention, for ResShift's Swin blocks.

``basicsr/archs/resshift/swin_transformer.py`` carries a second copy of the SwinIR attention whose blocks keep the
feature map in NCHW: ``window_partition`` (:34-46) / ``window_reverse`` (:48-62) take and return [B, C, H, W], and
``WindowAttention`` (:64-146) is the same arithmetic as ``swinir_arch.py:95-175`` (same parameter names, so state dicts
interchange).  ``UNetModelSwin`` (unet_arch.py:735) itself -- GroupNorm, 1x1-conv Mlp, timestep embeddings -- is outside
the hot path of SURVEY.md section 8; what is provided here is the attention it calls, on the srb200 kernels:

* :func:`window_partition` / :func:`window_reverse` with the reference's NCHW signatures (bit-exact remaps);
* :class:`WindowAttention` -- drop-in for the reference class (``forward(x, mask)`` on partitioned windows), plus
  :meth:`WindowAttention.forward_nchw`, which replaces the whole ``roll -> window_partition -> attn -> window_reverse ->
  roll`` sequence of a ResShift ``SwinTransformerBlock.forward`` (:222-262) by one NCHW -> NHWC entry kernel, the qkv
  tap-GEMM, the fused window-attention kernel (shift, partition, mask and reverse are its addressing), the proj
  tap-GEMM and one exit kernel.
"""
import torch

from ..ops import sr_b200 as ops
from .arch_util import _NHWCToImage, require_cuda
from .swinir_arch import WindowAttention as _WindowAttention


def window_partition(x, window_size):
    """[B, C, H, W] -> [num_windows*B, ws, ws, C] (reference resshift/swin_transformer.py:34-46).  The remap kernel serves
    tensors outside autograd; a tensor that needs a gradient takes the (equally bit-exact) view / permute expression."""
    t = x.permute(0, 2, 3, 1).contiguous()
    if t.is_cuda and t.element_size() in (2, 4) and not (torch.is_grad_enabled() and x.requires_grad):
        return ops.raw.window_partition(t, window_size, 0)
    b, h, w, c = t.shape
    return t.view(b, h // window_size, window_size, w // window_size, window_size, c).permute(0, 1, 3, 2, 4, 5) \
        .contiguous().view(-1, window_size, window_size, c)


def window_reverse(windows, window_size, H, W):
    """[num_windows*B, ws, ws, C] -> [B, C, H, W] (reference resshift/swin_transformer.py:48-62)."""
    if windows.is_cuda and windows.element_size() in (2, 4) and not (torch.is_grad_enabled() and windows.requires_grad):
        t = ops.raw.window_reverse(windows.contiguous(), window_size, H, W, 0)
    else:
        b = windows.shape[0] // ((H // window_size) * (W // window_size))
        t = windows.view(b, H // window_size, W // window_size, window_size, window_size, -1) \
            .permute(0, 1, 3, 2, 4, 5).contiguous().view(b, H, W, -1)
    return t.permute(0, 3, 1, 2).contiguous()


def _shift_mask(h, w, ws, shift):
    """0 / -100 SW-MSA mask (reference resshift/swin_transformer.py:207-220, same as swinir_arch.py:262-281)."""
    img = torch.zeros((1, h, w, 1))
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = img.view(1, h // ws, ws, w // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, float(-100.0)).masked_fill(m == 0, float(0.0))


class WindowAttention(_WindowAttention):
    """Same constructor, parameters and ``forward(x, mask)`` as the reference (:64-146)."""

    def forward_nchw(self, x, shift_size=0):
        """x [B, C, H, W] float -> [B, C, H, W]: shifted-window attention of the whole map (H, W multiples of the
        window; the SW-MSA mask of :207-220 is derived analytically by the kernel)."""
        require_cuda(x, 'WindowAttention.forward_nchw')
        ws = self.window_size[0]
        if not self.on_kernels():
            b, c, h, w = x.shape
            xs = torch.roll(x, shifts=(-shift_size, -shift_size), dims=(2, 3)) if shift_size > 0 else x
            win = window_partition(xs, ws).view(-1, ws * ws, c)
            mask = _shift_mask(h, w, ws, shift_size).to(x.device) if shift_size > 0 else None
            y = window_reverse(self.torch_forward(win, mask).view(-1, ws, ws, c), ws, h, w)
            return torch.roll(y, shifts=(shift_size, shift_size), dims=(2, 3)) if shift_size > 0 else y
        with torch.cuda.device(x.device):
            t = ops.image_to_nhwc(x, None, 1.0, ops.pad64(x.shape[1]))
            y = self.forward_nhwc(t, shift_size)
            return _NHWCToImage.apply(y, x.shape[1]).to(x.dtype)
