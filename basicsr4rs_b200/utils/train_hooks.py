"""Step-overhead hooks on the caller's side -- opt-in, the reference models stay as
they are.

``BaseModel.reduce_loss_dict`` (basicsr/models/base_model.py:376-401) runs EVERY iteration: it stacks the loss
tensors, ``dist.reduce``s them and calls ``.item()`` on each -- a host<->device synchronisation per step that drains
the launch queue (the next step's kernels cannot be queued behind the current one) -- although the values are only
looked at every ``print_freq`` iterations (train.py:177-182, utils/logger.py:107-108 formats them with ``:.4e``).
:func:`install_lazy_loss_log` replaces it by a version that keeps the values on the device: no collective and no
synchronisation until something formats, converts or compares a value, i.e. when the logger prints.

``optimizer_g.zero_grad()`` (sr_model.py:92): torch >= 2.0 defaults to ``set_to_none=True`` -- no kernel is launched,
nothing to remove (measured: 0 launches per step in tools/prof_step.py).
"""
import types
from collections import OrderedDict

import torch


class LazyScalar:
    """A loss value that stays on the device until it is read.  Reading (``float()``, ``format``, comparison, ``item``)
    performs the cross-rank average of the reference's ``reduce_loss_dict`` and the single device->host copy."""

    __slots__ = ('_t', '_dist', '_world', '_v')

    def __init__(self, tensor, dist_on=False, world=1):
        self._t, self._dist, self._world, self._v = tensor.detach(), dist_on, world, None

    def item(self):
        if self._v is None:
            t = self._t.float().mean()
            if self._dist:
                # every rank reads at the same iteration (the logger's print_freq), so the collective matches up; all
                # ranks get the mean (the reference divides on rank 0 only and prints there)
                t = t.clone()
                torch.distributed.all_reduce(t)
                t /= self._world
            self._v = t.item()
            self._t = None
        return self._v

    def __float__(self):
        return self.item()

    def __format__(self, spec):
        return format(self.item(), spec)

    def __repr__(self):
        return f'LazyScalar({self.item()!r})' if self._v is not None else 'LazyScalar(<on device>)'

    def __lt__(self, other):
        return self.item() < float(other)

    def __gt__(self, other):
        return self.item() > float(other)

    def __eq__(self, other):
        return self.item() == float(other)

    def __hash__(self):
        return id(self)


def lazy_reduce_loss_dict(self, loss_dict):
    """Drop-in for ``BaseModel.reduce_loss_dict``: same keys, values are :class:`LazyScalar` instead of floats."""
    dist_on = bool(self.opt.get('dist', False))
    world = int(self.opt.get('world_size', 1))
    return OrderedDict((k, LazyScalar(v, dist_on, world)) for k, v in loss_dict.items())


def install_lazy_loss_log(model):
    """``model``: a reference ``SRModel`` (or any ``BaseModel``).  Returns it for chaining."""
    model.reduce_loss_dict = types.MethodType(lazy_reduce_loss_dict, model)
    return model


# ------------------------------------------------------------------ tiled validation / test
def tiled_test(self, tile=1024, overlap=32, min_pixels=None):
    """Drop-in for ``SRModel.test`` (sr_model.py:120-129) / ``SwinIRModel.test`` (swinir_model.py:14-36) on scenes too
    large for one forward: ``self.lq`` [B,C,H,W] is cut into ``tile`` x ``tile`` LR tiles overlapping by ``overlap``
    pixels (utils/tiling.tiled_forward; every tile padded to a multiple of the network's ``window_size`` the way
    ``SwinIRModel.test`` pads the whole image, seams cropped at the middle of the overlap), each forwarded under
    ``no_grad`` by ``net_g_ema`` when the model keeps one, else by ``net_g`` in eval mode (restored to train mode
    afterwards, as the reference does).  Images of up to ``min_pixels`` LR pixels (default: one tile) take ONE forward,
    exactly as the reference's ``test`` -- padding and cropping included."""
    from .tiling import pad_to_multiple, tiled_forward
    opt = self.opt
    multiple = int(opt.get('network_g', {}).get('window_size', 1) or 1)
    scale = int(opt.get('scale', 1))
    ema = hasattr(self, 'net_g_ema')
    net = self.net_g_ema if ema else self.net_g
    was_training = net.training
    net.eval()
    lq = self.lq
    b, c, h, w = lq.shape
    limit = tile * tile if min_pixels is None else min_pixels
    try:
        with torch.no_grad():
            if h * w <= limit:
                img, _ = pad_to_multiple(lq, multiple)
                self.output = net(img)[:, :, :h * scale, :w * scale]
            else:
                outs = [tiled_forward(net, lq[i:i + 1], tile, scale, overlap=overlap, multiple=multiple)[0]
                        for i in range(b)]
                self.output = outs[0] if b == 1 else torch.cat(outs, 0)
    finally:
        if not ema and was_training:
            net.train()


def install_tiled_test(model, tile=1024, overlap=32, min_pixels=None):
    """``model``: a reference ``SRModel`` / ``SwinIRModel`` (anything with ``lq``, ``opt`` and ``net_g``): its ``test()``
    becomes :func:`tiled_test`; ``validation`` / ``nondist_validation`` (sr_model.py:131-201) call it unchanged."""
    model.test = types.MethodType(lambda self: tiled_test(self, tile, overlap, min_pixels), model)
    return model
