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
from contextlib import nullcontext as _nullcontext

import torch
from torch import nn as nn

from .. import _lib as L
from ..ops.sr_b200 import raw
from ..ops.sr_b200.sr_b200 import PackBook, pack_book

# arch instance -> GraphedSegments; kept outside the module so that copy.deepcopy (EMA copy, sr_model.py:51),
# state_dict and pickling never see CUDA graph objects
GRAPHS = weakref.WeakKeyDictionary()
BOOKS = weakref.WeakKeyDictionary()  # arch instance -> PackBook (bf16 GEMM operands of all its weights)


def capturing():
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


class Segment(nn.Module):
    """A slice of an arch sharing the parent's sub-modules (and therefore its parameters)."""

    def __init__(self, fn, modules):
        super().__init__()
        self.mods = nn.ModuleList(modules)
        self._fn = fn
        self.book = None  # set on the FIRST segment: its forward graph starts with the batched weight repack
        self.arena_floats = 0  # > 0: all zero-initialised fp32 scratch of this segment's forward from ONE fill
        # the zero-initialised fp32 scratch of the BACKWARD of this segment's Functions, from one fill recorded in the
        # forward (raw.backward_scratch); measured by the first grad-enabled run per input shape
        self._bwd_floats = {}

    def _run(self, *tensors):
        if not torch.is_grad_enabled():
            return self._run_fwd(*tensors)
        key = tuple(tuple(t.shape) for t in tensors)
        need = self._bwd_floats.get(key)
        with raw.backward_scratch(tensors[0].device, need) as scratch:
            out = self._run_fwd(*tensors)
        if need is None:
            self._bwd_floats[key] = scratch.measured
        return out

    def _run_fwd(self, *tensors):
        if self.arena_floats > 0:
            with raw.zero_arena(tensors[0].device, self.arena_floats):
                return self._fn(*tensors)
        return self._fn(*tensors)

    def forward(self, *tensors):
        if self.book is not None:
            with pack_book(self.book):
                self.book.refresh()
                return self._run(*tensors)
        return self._run(*tensors)


class GraphedSegments:
    """Per-input-signature cache of captured segments.

    ``build_segments()`` returns the list of :class:`Segment`; ``wire(call, x)`` runs the chain through
    ``call(i, *tensors)`` (the arch decides which outputs feed which segment)."""

    WARMUP = 3

    def __init__(self, build_segments, wire, book=None, pdl=False):
        self._build = build_segments
        self._wire = wire
        self._book = book
        self._pdl = pdl  # capture the GEMM launches with programmatic dependent launch (srb200_set_pdl)
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
        segs[0].book = self._book

        def eager_call(i, *args):
            record[i] = args
            return segs[i](*args)

        with torch.no_grad():  # one eager pass only to learn every segment's input signature
            self._wire(eager_call, x)
        # the differentiable stream between segments is bf16; fp32 tensors (image input, fp32 skip twins) are
        # carried along without gradients
        sample_args = tuple(
            tuple(a.detach().clone().requires_grad_(a.dtype == torch.bfloat16 and i > 0) for a in record[i])
            for i in range(len(segs)))
        if self._book is not None:
            # one warm-up forward+backward of every segment (side stream, like make_graphed_callables' own
            # warm-up) so that every packed operand -- fprop and dgrad layouts -- is registered in the PackBook
            # and the device table of pack items exists before anything is captured
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for seg, args in zip(segs, sample_args):
                    outs = seg(*args)
                    outs = tuple(o for o in (outs if isinstance(outs, tuple) else (outs,)) if o.requires_grad)
                    ins = [a for a in args if a.requires_grad] + [p for p in seg.parameters() if p.requires_grad]
                    if training and outs and ins:
                        torch.autograd.grad(outs, ins, [torch.ones_like(o) for o in outs], allow_unused=True)
                    del outs
                self._book.refresh(force=True)
            torch.cuda.current_stream().wait_stream(side)
            self._book.capture_ok = True
        n0 = L.launch_count
        prev = L.load().srb200_set_pdl(1 if self._pdl else 0)
        try:
            graphed = torch.cuda.make_graphed_callables(tuple(segs), sample_args, num_warmup_iters=self.WARMUP)
        finally:
            L.load().srb200_set_pdl(prev)
        self.kernels_per_step[key] = (L.launch_count - n0) // (self.WARMUP + 1)
        return list(graphed)


def graphed_forward(module, x, build_segments, wire):
    """Replay ``module``'s training forward/backward for input ``x`` from CUDA graphs (captured on first use)."""
    graphs = GRAPHS.get(module)
    if graphs is None:
        graphs = GRAPHS[module] = GraphedSegments(build_segments, wire, book=module._pack_book(),
                                                  pdl=getattr(module, 'graph_pdl', False))
    return graphs(x.contiguous().float(), True)


def chain_wire(nseg, carry=1):
    """wire() for the common shape: segment 0 returns (stream..., *carry); middle segments map stream -> stream;
    the last segment consumes (stream..., *carry).  ``carry`` = how many trailing outputs of segment 0 skip
    straight to the last segment (the long skip connection)."""

    def wire(call, x):
        out = call(0, x)
        stream, kept = out[:-carry], out[-carry:]
        for i in range(1, nseg - 1):
            stream = call(i, *stream)
            if not isinstance(stream, tuple):
                stream = (stream,)
        return call(nseg - 1, *stream, *kept)

    return wire


def split_even(items, n):
    n = min(max(1, n), max(1, len(items)))
    cuts = [round(i * len(items) / n) for i in range(n + 1)]
    return [items[cuts[i]:cuts[i + 1]] for i in range(n)]


def precapture(module):
    """Capture ``module``'s training graphs for ``module.graph_input_shape`` right after it lands on the GPU.

    DDP stashes an AccumulateGrad node per parameter on the stream it was constructed on; capturing a backward
    afterwards would make that (legacy) stream wait on the capturing stream, which CUDA rejects.  The reference
    wraps with DDP right after ``net.to(device)`` (base_model.py:94-99), so the archs call this from
    ``nn.Module._apply`` -- i.e. between ``.to(device)`` and the DDP wrap, with no change to the caller."""
    shape = getattr(module, 'graph_input_shape', None)
    if not getattr(module, 'cuda_graph', False) or not shape:
        return
    p = next(module.parameters(), None)
    if p is None or not p.is_cuda:
        return
    key_done = ('precaptured', tuple(shape), str(p.device))
    if module.__dict__.get('_srb200_precaptured') == key_done:
        return
    module.__dict__['_srb200_precaptured'] = key_done
    was_training = module.training
    module.train()
    with torch.enable_grad(), torch.cuda.device(p.device):
        module(torch.zeros(tuple(shape), dtype=torch.float32, device=p.device))
    module.train(was_training)


class ArchMixin:
    """Shared by the three archs: device copy of the (non-buffer) mean, and graph pre-capture on ``.to(device)``."""

    def _device_mean(self, x):
        m = getattr(self, '_mean_dev', None)
        if m is None or m.device != x.device:
            m = self.mean.detach().to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
            self._mean_dev = m
        return m

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if getattr(self, 'flat_grads', False) and next(self.parameters()).is_cuda:
            # the flat gradient buffer must exist BEFORE the graphs are captured: the captures record its addresses
            from ..utils.flat_ddp import flat_grads_of
            flat_grads_of(self)
        precapture(self)
        return out

    def _pack_book(self):
        book = BOOKS.get(self)
        if book is None:
            book = BOOKS[self] = PackBook()
        return book

    def forward(self, x):
        """nn.Module contract of the reference archs (NCHW float in / out); every packed weight requested below
        belongs to this network's :class:`PackBook`."""
        book = self._pack_book()
        # (device guard: the C ABI launches on the current device's stream -- a net on cuda:1 must not launch on cuda:0)
        with torch.cuda.device(x.device) if x.is_cuda else _nullcontext(), pack_book(book):
            if not torch.is_grad_enabled():
                # evaluation (SRModel.test runs net_g_ema this way, sr_model.py:120-129): the EMA loop updates
                # parameters through ``.data`` (base_model.py:81-82), which does not bump ``_version`` -- the version
                # tags prove nothing here, so every packed operand is re-derived by ONE launch per forward
                if book.entries and not capturing():
                    book.invalidate()
                    book.refresh()
                return self._forward(x)
            # eager training: the backward's zeroed scratch of every Function from one fill (raw.backward_scratch);
            # sized by a measuring first pass per input shape (0 when CUDA-graph segments bring their own)
            needs = self.__dict__.setdefault('_bwd_need', {})
            need = needs.get(tuple(x.shape))
            with raw.backward_scratch(x.device, need) as scratch:
                out = self._forward(x)
            if need is None:
                needs[tuple(x.shape)] = scratch.measured
            return out
