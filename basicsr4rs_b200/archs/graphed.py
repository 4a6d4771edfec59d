"""CUDA-graph replay of an arch's forward and backward (opt-in, ``cuda_graph: true`` under ``network_g``).

The hot path is hundreds of short kernels per step (EDSR-L: ~480 launches, 20-60 us each); issued eagerly
from Python they are launch-bound.  Instead of a tracing compiler, the arch is cut into a few segments,
each captured once per input shape with ``torch.cuda.make_graphed_callables`` (forward graph + backward
graph, weight repacking included in the capture) and replayed thereafter.  The module stays a drop-in
``nn.Module``: the reference's ``SRModel.optimize_parameters`` (sr_model.py:91-118) and DDP
(base_model.py:98-99) are unchanged; parameter gradients leave each segment's backward through normal
autograd, so DDP's bucket all-reduce overlaps the backward of the remaining segments.
"""
import weakref

import torch
from torch import nn as nn

from .. import _lib as L

# arch instance -> GraphedSegments; kept outside the module so that copy.deepcopy (EMA copy, sr_model.py:51),
# state_dict and pickling never see CUDA graph objects
GRAPHS = weakref.WeakKeyDictionary()


def capturing():
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


class Segment(nn.Module):
    """A slice of an arch sharing the parent's sub-modules (and therefore its parameters)."""

    def __init__(self, fn, modules):
        super().__init__()
        self.mods = nn.ModuleList(modules)
        self._fn = fn

    def forward(self, *tensors):
        return self._fn(*tensors)


class GraphedSegments:
    """Per-input-signature cache of captured segments.

    ``build_segments()`` returns the list of :class:`Segment`; ``wire(call, x)`` runs the chain through
    ``call(i, *tensors)`` (the arch decides which outputs feed which segment)."""

    WARMUP = 3

    def __init__(self, build_segments, wire):
        self._build = build_segments
        self._wire = wire
        self._cache = {}
        self.kernels_per_step = {}

    def __call__(self, x, training):
        key = (tuple(x.shape), x.dtype, str(x.device), bool(training))
        if key not in self._cache:
            self._cache[key] = self._capture(x, training, key)
        graphed = self._cache[key]
        return self._wire(lambda i, *a: graphed[i](*a), x)

    def _capture(self, x, training, key):
        segs = self._build()
        for s in segs:
            s.train(training)
        record = {}

        def eager_call(i, *args):
            record[i] = args
            return segs[i](*args)

        with torch.no_grad():  # one eager pass only to learn every segment's input signature
            self._wire(eager_call, x)
        sample_args = tuple(
            tuple(a.detach().clone().requires_grad_(a.is_floating_point() and i > 0) for a in record[i])
            for i in range(len(segs)))
        n0 = L.launch_count
        graphed = torch.cuda.make_graphed_callables(tuple(segs), sample_args, num_warmup_iters=self.WARMUP)
        self.kernels_per_step[key] = (L.launch_count - n0) // (self.WARMUP + 1)
        return list(graphed)
