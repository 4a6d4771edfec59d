"""Fused multi-tensor EMA of the generator weights.

``BaseModel.model_ema`` (basicsr/models/base_model.py:75-82) does, every iteration::

    for k in net_g_ema_params: net_g_ema_params[k].data.mul_(decay).add_(net_g_params[k].data, alpha=1 - decay)

i.e. two tiny kernels per parameter (RCAN: 3260 launches, more than a whole training step of this repo takes).
:class:`FusedEMA` does the same update for all parameters in ONE ``srb200_multi_axpby`` launch.  Drop-in use from a
model class::

    self._ema = FusedEMA(self.get_bare_model(self.net_g), self.net_g_ema)      # once, after both nets are on the GPU
    def model_ema(self, decay=0.999): self._ema.step(decay)
"""
import ctypes

import numpy as np
import torch

from .. import _lib as L

_ITEM = np.dtype([('src', '<i8'), ('dst', '<i8'), ('n', '<i8'), ('chunk_begin', '<i8')])


class FusedEMA:

    def __init__(self, net, net_ema):
        self._net_ema = net_ema
        src = dict(net.named_parameters())
        self.pairs = [(src[k], p) for k, p in net_ema.named_parameters()]  # same keys, like the reference loop
        for s, d in self.pairs:
            if not (s.is_cuda and d.is_cuda and s.dtype == torch.float32 and d.dtype == torch.float32 and
                    s.is_contiguous() and d.is_contiguous() and s.shape == d.shape):
                raise RuntimeError('FusedEMA needs contiguous fp32 CUDA parameters of equal shape')
        self._ptrs = None
        self._table = None

    def _build(self):
        tab = np.zeros(len(self.pairs), dtype=_ITEM)
        chunk = 0
        for i, (s, d) in enumerate(self.pairs):
            tab[i] = (s.data_ptr(), d.data_ptr(), s.numel(), chunk)
            chunk += (s.numel() + 1023) // 1024
        dev = self.pairs[0][0].device
        self._table = torch.from_numpy(tab.view(np.uint8).reshape(len(tab), -1).copy()).to(dev)
        self._chunks = chunk

    @torch.no_grad()
    def step(self, decay=0.999):
        ptrs = [(s.data_ptr(), d.data_ptr()) for s, d in self.pairs]
        if ptrs != self._ptrs:  # first call, or a parameter was re-allocated (.to(), load_state_dict keeps storage)
            self._ptrs = ptrs
            self._build()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.check(L.load().srb200_multi_axpby(ctypes.c_void_p(self._table.data_ptr()), len(self.pairs), self._chunks,
                                            float(decay), float(1.0 - decay), stream), 'multi_axpby')
        # the kernel wrote through raw pointers: ``_version`` of the EMA parameters did not move, so the bf16 operands
        # packed from them (PackBook / per-parameter caches) must be told
        invalidate_packs(self._net_ema)


def invalidate_packs(net):
    """Drop every cached bf16 operand derived from ``net``'s parameters (call after writing them through ``.data`` or
    raw pointers while grad mode is on; no_grad forwards re-derive them anyway)."""
    from ..archs.graphed import BOOKS
    for m in net.modules():
        book = BOOKS.get(m)
        if book is not None:
            book.invalidate()
    for p in net.parameters():
        p.__dict__.pop('_srb200_pack', None)
