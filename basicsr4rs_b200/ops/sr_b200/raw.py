"""Thin tensor-level wrappers over the srb200 C-ABI (no autograd).

Each function validates tensors (CUDA, contiguous, dtype), allocates the outputs with PyTorch's
caching allocator, forwards raw ``data_ptr()``s on the *current* stream and turns a non-zero
status into ``RuntimeError`` -- the same contract as the reference's pybind shims
(basicsr/ops/fused_act/src/fused_bias_act.cpp:10-26, CHECK_CUDA / CHECK_CONTIGUOUS).
"""
import ctypes
import threading

import torch

from ... import _lib as L


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class EventProbe:
    """Times selected kernel launches with CUDA events on the launching stream (bench.py's roofline
    leg): ``probe = EventProbe(lambda kind, info: ...)``; ``raw.PROBE = probe``; read ``probe.ms()`` (durations)
    or ``probe.records()`` ((kind, info, ms) triples)."""

    def __init__(self, predicate, limit=8192):
        self.predicate, self.limit, self.pairs = predicate, limit, []

    def begin(self, kind, info):
        if len(self.pairs) >= self.limit or not self.predicate(kind, info):
            return None
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), kind, info)
        ev[0].record()
        return ev

    def end(self, ev):
        if ev is not None:
            ev[1].record()
            self.pairs.append(ev)

    def ms(self):
        torch.cuda.synchronize()
        return [p[0].elapsed_time(p[1]) for p in self.pairs]

    def records(self):
        torch.cuda.synchronize()
        return [(p[2], p[3], p[0].elapsed_time(p[1])) for p in self.pairs]


def probed(kind, info, launch):
    """Run ``launch()`` (one C-ABI call) between two events when a probe is active and wants this kernel."""
    ev = PROBE.begin(kind, info) if PROBE is not None else None
    out = launch()
    if ev is not None:
        PROBE.end(ev)
    return out


PROBE = None

# ------------------------------------------------------------------ side branch for weight gradients
# A layer's weight-gradient GEMM and its data-gradient GEMM both consume dY and neither needs the other; only the
# data gradient is on the critical path of the backward chain.  ``with side_branch(device): acc = wgrad(...)`` puts the
# launch on a per-device side stream that waits for what the current stream has been given so far; ``side_join``
# makes the current stream wait for the side stream again (before the gradients are finalised, and always before the
# Function returns: the operands are main-stream tensors that must outlive the side kernels).  Inside a CUDA-graph
# capture the event pair becomes a fork / join of the graph.  SRB_SIDE_WGRAD=0 keeps everything on one stream.
import os as _os

SIDE_WGRAD = _os.environ.get('SRB_SIDE_WGRAD', '1') == '1'
_side_streams = {}
_side_main = threading.local()  # the stream a side branch forked from, while inside the branch


def _side_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    s = _side_streams.get(key)
    if s is None:
        s = _side_streams[key] = torch.cuda.Stream(device=key)
    return s


class side_branch:
    def __init__(self, device):
        self.device = device
        self.on = SIDE_WGRAD and PROBE is None  # (the event probe times launches on ONE stream)

    def __enter__(self):
        if self.on:
            side = _side_stream(self.device)
            main = torch.cuda.current_stream(self.device)
            side.wait_stream(main)
            self.ctx = torch.cuda.stream(side)
            self.ctx.__enter__()
            _side_main.cur = main
        return self

    def __exit__(self, *exc):
        if self.on:
            _side_main.cur = None
            self.ctx.__exit__(*exc)
        return False


def side_join(device):
    if SIDE_WGRAD and PROBE is None:
        torch.cuda.current_stream(device).wait_stream(_side_stream(device))


_arena = threading.local()


class zero_arena:
    """All fp32 zero-initialised scratch (split-K accumulators, column sums, ...) requested by the wrappers
    inside the ``with`` block is carved out of ONE ``torch.zeros`` -- one fill kernel instead of one per buffer."""

    def __init__(self, device, n_floats, buf=None):
        # buf: an already zeroed fp32 buffer to carve from instead of a fresh fill (see backward_scratch)
        self.device, self.n = device, int(n_floats) + 64
        self.buf = buf if (buf is not None and buf.numel() >= self.n and buf.device == device) else None

    def __enter__(self):
        self.prev = getattr(_arena, 'cur', None)
        buf = self.buf if self.buf is not None else torch.zeros((self.n,), dtype=torch.float32, device=self.device)
        _arena.cur = [buf, 0]
        return self

    def __exit__(self, *exc):
        _arena.cur = self.prev
        return False


_bwd_scratch = threading.local()


class backward_scratch:
    """Zeroed fp32 scratch for the BACKWARD of the Functions whose forward runs inside the block: one fill for all
    of them (in the forward, where a CUDA-graph segment records it once) instead of one fill at the head of every
    layer's backward, where it also breaks the programmatic-dependent-launch chain between two GEMMs.

    ``n_floats=None`` only MEASURES: nothing is handed out, ``self.measured`` is what the block asked for."""

    def __init__(self, device, n_floats):
        self.device, self.n, self.measured = device, (None if n_floats is None else int(n_floats)), 0

    def __enter__(self):
        self.prev = getattr(_bwd_scratch, 'cur', None)
        if self.n is None:
            _bwd_scratch.cur = [None, 0]
        elif self.n > 0:
            _bwd_scratch.cur = [torch.zeros((self.n,), dtype=torch.float32, device=self.device), 0]
        else:
            _bwd_scratch.cur = None
        return self

    def __exit__(self, *exc):
        cur = _bwd_scratch.cur
        if cur is not None:
            self.measured = cur[1]
        _bwd_scratch.cur = self.prev
        return False


def take_backward_scratch(n_floats, device):
    """A zeroed buffer big enough for ``zero_arena(device, n_floats, buf=...)``, or None (no provider / no room)."""
    cur = getattr(_bwd_scratch, 'cur', None)
    if cur is None:
        return None
    n = int(n_floats) + 64
    start = (cur[1] + 3) // 4 * 4
    if cur[0] is None:  # measuring pass
        cur[1] = start + n
        return None
    if cur[0].device != device or start + n > cur[0].numel():
        return None
    cur[1] = start + n
    return cur[0][start:start + n]


def stash_backward_scratch(ctx, n_floats, device):
    """Function.forward side: reserve the backward's zeroed scratch on ``ctx`` (None without a provider)."""
    ctx.scratch = take_backward_scratch(n_floats, device) if any(ctx.needs_input_grad) else None


def backward_arena(ctx, device, n_floats):
    """Function.backward side: the zero arena of this backward, carved from the stashed scratch when there is one
    (single use: a second backward through the same node falls back to its own fill)."""
    scratch, ctx.scratch = getattr(ctx, 'scratch', None), None
    return zero_arena(device, n_floats, buf=scratch)


def zeros_f32(shape, device):
    """fp32 zeros, 16-byte aligned, from the active :class:`zero_arena` when it has room."""
    n = 1
    for d in shape:
        n *= int(d)
    cur = getattr(_arena, 'cur', None)
    if cur is not None and cur[0].device == device:
        start = (cur[1] + 3) // 4 * 4
        if start + n <= cur[0].numel():
            cur[1] = start + n
            return cur[0][start:start + n].view(shape)
    out = torch.zeros(shape, dtype=torch.float32, device=device)
    main = getattr(_side_main, 'cur', None)
    if main is not None:  # allocated on the side stream, consumed (and dropped) on the main one
        out.record_stream(main)
    return out


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _chk(t, name, dtype=None):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor')
    if t.device.index != torch.cuda.current_device():
        # the C ABI launches on the CURRENT device's stream (it never changes the device itself)
        raise RuntimeError(f'{name} lives on {t.device} but the current device is cuda:{torch.cuda.current_device()}: '
                           'wrap the call in torch.cuda.device(tensor.device) (the archs\' forward does)')
    if not t.is_contiguous():
        raise RuntimeError(f'{name} must be contiguous')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f'{name} must be {dtype}, got {t.dtype}')


def _esize(t):
    es = t.element_size()
    if es not in (2, 4):
        raise RuntimeError('index remaps support 2- and 4-byte element types')
    return es


# ------------------------------------------------------------------ index remaps (bit-exact)
def pixel_shuffle_nchw(x, r, inverse=False):
    """nn.PixelShuffle(r) (arch_util.py:135,138) / its inverse on NCHW tensors."""
    _chk(x, 'x')
    lib = L.load()
    if not inverse:
        b, c, h, w = x.shape
        assert c % (r * r) == 0
        co = c // (r * r)
        out = torch.empty((b, co, h * r, w * r), dtype=x.dtype, device=x.device)
    else:
        b, co, hh, ww = x.shape
        assert hh % r == 0 and ww % r == 0
        h, w = hh // r, ww // r
        out = torch.empty((b, co * r * r, h, w), dtype=x.dtype, device=x.device)
    L.check(lib.srb200_pixel_shuffle_nchw(_ptr(x), _ptr(out), b, co, h, w, r, _esize(x), int(inverse), _stream()),
            'pixel_shuffle_nchw')
    return out


def pixel_shuffle_nhwc(x, r, inverse=False):
    _chk(x, 'x')
    lib = L.load()
    if not inverse:
        b, h, w, c = x.shape
        co = c // (r * r)
        out = torch.empty((b, h * r, w * r, co), dtype=x.dtype, device=x.device)
    else:
        b, hh, ww, co = x.shape
        h, w = hh // r, ww // r
        out = torch.empty((b, h, w, co * r * r), dtype=x.dtype, device=x.device)
    L.check(lib.srb200_pixel_shuffle_nhwc(_ptr(x), _ptr(out), b, co, h, w, r, _esize(x), int(inverse), _stream()),
            'pixel_shuffle_nhwc')
    return out


def window_partition(x, ws, shift=0):
    """roll(-shift) + window_partition (swinir_arch.py:293-300): [B,H,W,C] -> [B*nW, ws, ws, C]."""
    _chk(x, 'x')
    b, h, w, c = x.shape
    out = torch.empty((b * (h // ws) * (w // ws), ws, ws, c), dtype=x.dtype, device=x.device)
    L.check(L.load().srb200_window_remap(_ptr(x), _ptr(out), b, h, w, c, ws, shift, _esize(x), 0, _stream()),
            'window_partition')
    return out


def window_reverse(win, ws, h, w, shift=0):
    """window_reverse + roll(+shift) (swinir_arch.py:309-316): [B*nW, ws, ws, C] -> [B,H,W,C]."""
    _chk(win, 'win')
    c = win.shape[-1]
    b = win.shape[0] // ((h // ws) * (w // ws))
    out = torch.empty((b, h, w, c), dtype=win.dtype, device=win.device)
    L.check(L.load().srb200_window_remap(_ptr(win), _ptr(out), b, h, w, c, ws, shift, _esize(win), 1, _stream()),
            'window_reverse')
    return out


def roll_nhwc(x, sy, sx):
    _chk(x, 'x')
    b, h, w, c = x.shape
    out = torch.empty_like(x)
    L.check(L.load().srb200_roll_nhwc(_ptr(x), _ptr(out), b, h, w, c, sy, sx, _esize(x), _stream()), 'roll_nhwc')
    return out


# ------------------------------------------------------------------ layout entry / exit
def nchw_to_nhwc(x, c_pad, shift=None, scale=1.0):
    _chk(x, 'x', torch.float32)
    b, c, h, w = x.shape
    out = torch.empty((b, h, w, c_pad), dtype=torch.bfloat16, device=x.device)
    L.check(L.load().srb200_nchw_to_nhwc(_ptr(x), _ptr(out), b, c, h, w, c_pad, _ptr(shift), float(scale), _stream()),
            'nchw_to_nhwc')
    return out


def nhwc_to_nchw(x, c, shift=None, scale=1.0):
    _chk(x, 'x', torch.bfloat16)
    b, h, w, c_pad = x.shape
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=x.device)
    L.check(L.load().srb200_nhwc_to_nchw(_ptr(x), _ptr(out), b, c, h, w, c_pad, _ptr(shift), float(scale), _stream()),
            'nhwc_to_nchw')
    return out


def nearest_up2(x, inverse=False):
    """NHWC bf16 nearest-neighbour x2 (inverse=False) or its gradient, the 2x2 block sum (inverse=True)."""
    _chk(x, 'x', torch.bfloat16)
    b, h, w, c = x.shape
    if not inverse:
        out = torch.empty((b, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=x.device)
        L.check(L.load().srb200_nearest_up2(_ptr(x), _ptr(out), b, h, w, c, 0, _stream()), 'nearest_up2')
    else:
        assert h % 2 == 0 and w % 2 == 0
        out = torch.empty((b, h // 2, w // 2, c), dtype=torch.bfloat16, device=x.device)
        L.check(L.load().srb200_nearest_up2(_ptr(x), _ptr(out), b, h // 2, w // 2, c, 1, _stream()), 'nearest_up2_bwd')
    return out


def tap_stencil(t32, c, bias=None, shift=None, scale=1.0):
    """fp32 [B,H,W,Tp] per-tap partial products -> NCHW fp32 image: sum of the 9 shifted taps + bias, affine."""
    _chk(t32, 't32', torch.float32)
    b, h, w, tp = t32.shape
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=t32.device)
    L.check(L.load().srb200_tap_stencil(_ptr(t32), _ptr(out), b, c, h, w, tp, _ptr(bias), _ptr(shift), float(scale),
                                        _stream()), 'tap_stencil')
    return out


def tap_im2col(g, gp, scale=1.0):
    """NCHW fp32 image gradient [B,C,H,W] -> NHWC bf16 [B,H,W,gp] with channel tap*C+c = scale * g[c] shifted by -tap."""
    _chk(g, 'g', torch.float32)
    b, c, h, w = g.shape
    out = torch.empty((b, h, w, gp), dtype=torch.bfloat16, device=g.device)
    L.check(L.load().srb200_tap_im2col(_ptr(g), _ptr(out), b, c, h, w, gp, float(scale), _stream()), 'tap_im2col')
    return out


# ------------------------------------------------------------------ weights
def pack_weight(w, n_pad, k_pad, perm_out=None, perm_in=None, transpose=False, out=None):
    """fp32 [Co, Ci, kh, kw] (or [Co, Ci]) -> bf16 [taps, n_pad, k_pad] (or [taps, k_pad, n_pad])."""
    _chk(w, 'w', torch.float32)
    co, ci = w.shape[0], w.shape[1]
    taps = w.numel() // (co * ci)
    shape = (taps, k_pad, n_pad) if transpose else (taps, n_pad, k_pad)
    if out is None:
        out = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    L.check(L.load().srb200_pack_weight(_ptr(w), co, ci, taps, _ptr(perm_out), n_pad, _ptr(perm_in), k_pad,
                                        int(transpose), _ptr(out), _stream()), 'pack_weight')
    return out


def _item_table(rows):
    import numpy as np
    tab = np.zeros(len(rows), dtype=np.dtype(L.PACK_ITEM_FIELDS))
    chunk = 0
    for i, r in enumerate(rows):
        tab[i]['src'] = r['src'].data_ptr()
        tab[i]['dst'] = r['dst'].data_ptr()
        tab[i]['perm_out'] = r['perm_out'].data_ptr() if r.get('perm_out') is not None else 0
        tab[i]['perm_in'] = r['perm_in'].data_ptr() if r.get('perm_in') is not None else 0
        for k in ('Co', 'Ci', 'taps', 'Np', 'Kp'):
            tab[i][k] = int(r[k])
        tab[i]['transpose'] = int(r.get('transpose', 0))  # 0 / 1: bf16 operand layouts, 2: fp32 copy (bias)
        tab[i]['alpha'] = float(r.get('alpha', 1.0))
        tab[i]['reserved'] = int(r.get('fill_index', -1)) + 1  # fp32-copy items: one pad element := alpha
        tab[i]['chunk_begin'] = chunk
        chunk += (int(r['Np']) * int(r['Kp']) + 255) // 256
    return tab, chunk


MAX_INLINE_ITEMS = 8


# Gradient sink (basicsr4rs_b200.utils.flat_ddp.FlatGrads): parameter storage address -> a preallocated fp32 view of that
# parameter's slice of ONE flat gradient buffer.  finalize_grads writes a sunk parameter's gradient straight into its
# slice (the address is fixed, so CUDA-graph captures record it) -- already multiplied by the sink's scale, 1 / world_size
# under data parallelism, so that the all-reduce can be a plain SUM (NCCL's in-switch NVLS reduction does not do
# pre-multiplied averages) -- and returns a fresh alias of it, which autograd adopts as ``param.grad`` without a copy:
# the data-parallel all-reduce then runs on the flat buffer itself.
GRAD_SINK = {}  # parameter data_ptr -> (fp32 view of its slice, scale applied to the gradient written there)


def _grad_target(param, shape, device):
    """(destination tensor, extra gradient scale): the parameter's slice of the flat buffer when it has one."""
    if param is not None and GRAD_SINK:
        hit = GRAD_SINK.get(param.data_ptr())
        if hit is not None and hit[0].device == device and tuple(hit[0].shape) == tuple(shape):
            if param.grad is None:
                return hit
            # gradient accumulation (a second backward before zero_grad): param.grad already aliases the slice, writing
            # it again would lose the first pass -- hand autograd a fresh, equally scaled tensor to add instead
            return torch.empty(tuple(shape), dtype=torch.float32, device=device), hit[1]
    return torch.empty(tuple(shape), dtype=torch.float32, device=device), 1.0


def finalize_grads(items, params=None):
    """All weight and bias gradients of one layer's backward in ONE launch (srb200_unpack_wgrads_inline).

    ``items``: list of ('w', acc [taps,Np,Kp] fp32, w_shape, perm_out, perm_in, alpha) or
    ('b', colsum [Np] fp32, n_bias, perm_out, alpha).  Returns the gradients in the parameter layouts.
    ``params`` (optional, aligned with ``items``): the parameters the gradients belong to -- those registered in
    :data:`GRAD_SINK` receive their gradient in place."""
    rows, outs = [], []
    params = params if params is not None else [None] * len(items)
    for it, prm in zip(items, params):
        if it[0] == 'w':
            _, acc, w_shape, perm_out, perm_in, alpha = it
            taps, n_pad, k_pad = acc.shape
            # every parameter element has exactly one packed slot (the index maps are onto: padding only ADDS
            # packed slots), so the scatter writes the whole gradient -- no zero fill needed
            g, sc = _grad_target(prm, w_shape, acc.device)
            rows.append(dict(src=acc, dst=g, Co=w_shape[0], Ci=w_shape[1], taps=taps, Np=n_pad, Kp=k_pad,
                             perm_out=perm_out, perm_in=perm_in, alpha=alpha * sc))
        elif it[0] == 'bcol':
            # a bias gradient that sits in COLUMN `col` of a weight-gradient accumulator [1, Np, Kp] (the layer's input
            # carried a constant-one pad channel there, see srb200_layernorm_fwd): a strided [Np x 1] item
            _, acc, col, n_bias, perm_out, alpha = it
            _, n_pad, k_pad = acc.shape
            g, sc = _grad_target(prm, (n_bias,), acc.device)
            rows.append(dict(src=acc[0, :, col:], dst=g, Co=n_bias, Ci=1, taps=1, Np=n_pad, Kp=k_pad, perm_out=perm_out,
                             alpha=alpha * sc))
        else:
            _, cs, n_bias, perm_out, alpha = it
            g, sc = _grad_target(prm, (n_bias,), cs.device)
            rows.append(dict(src=cs, dst=g, Co=n_bias, Ci=1, taps=1, Np=cs.numel(), Kp=1, perm_out=perm_out,
                             alpha=alpha * sc))
        outs.append(g)
    for i in range(0, len(rows), MAX_INLINE_ITEMS):
        tab, _ = _item_table(rows[i:i + MAX_INLINE_ITEMS])
        L.check(L.load().srb200_unpack_wgrads_inline(tab.ctypes.data_as(ctypes.c_void_p), len(tab), _stream()),
                'unpack_wgrads_inline')
    # (fresh tensor objects: autograd adopts an incoming gradient as ``param.grad`` without a copy only when nobody else
    # holds that tensor object -- the sink keeps its own view)
    return [g.detach() for g in outs] if GRAD_SINK else outs


def pack_items(rows, device):
    """Device table of ``srb200_pack_item`` rows for :func:`pack_weights` / :func:`unpack_wgrads`.

    ``rows``: dicts with src, dst (tensors), Co, Ci, taps, Np, Kp and optional perm_out, perm_in (int32 tensors),
    transpose, alpha.  Returns (table tensor, n_items, total_chunks); keep the tensors of ``rows`` alive."""
    import numpy as np
    tab, chunk = _item_table(rows)
    t = torch.from_numpy(tab.view(np.uint8).reshape(len(rows), -1).copy()).to(device)
    return t, len(rows), chunk


def pack_weights(table, n_items, total_chunks):
    """One launch: every fp32 weight of the table -> its bf16 GEMM operand (srb200_pack_weights)."""
    L.check(L.load().srb200_pack_weights(_ptr(table), n_items, total_chunks, _stream()), 'pack_weights')


def unpack_wgrads(table, n_items, total_chunks):
    """One launch: every fp32 [taps, Np, Kp] accumulator of the table -> gradient in the parameter layout."""
    L.check(L.load().srb200_unpack_wgrads(_ptr(table), n_items, total_chunks, _stream()), 'unpack_wgrads')


def unpack_wgrad(acc, w_shape, perm_out=None, perm_in=None, alpha=1.0):
    """fp32 [taps, n_pad, k_pad] accumulator -> fp32 gradient shaped like the nn.Parameter."""
    _chk(acc, 'acc', torch.float32)
    taps, n_pad, k_pad = acc.shape
    co, ci = w_shape[0], w_shape[1]
    alloc = torch.zeros if (perm_out is not None or perm_in is not None) else torch.empty
    gw = alloc(tuple(w_shape), dtype=torch.float32, device=acc.device)
    L.check(L.load().srb200_unpack_wgrad(_ptr(acc), _ptr(gw), co, ci, taps, _ptr(perm_out), n_pad, _ptr(perm_in),
                                         k_pad, float(alpha), _stream()), 'unpack_wgrad')
    return gw


# ------------------------------------------------------------------ tap-GEMM
def tapgemm(x, wp, *, ksize, cout, bias=None, act=L.ACT_NONE, act_slope=0.0, alpha=1.0, mask_src=None,
            mask_mode=L.MASK_NONE, mask_slope=0.0, residual=None, flip=False, src_r=1, out_mode=L.OUT_NHWC,
            out_r=1, out_c=0, out_scale=1.0, out_shift=None, want_aux=False, residual_f32=None, want_f32=False,
            alpha_per_sample=None, want_colsum=False, aux_grad=False, alpha_on_bias=False, ln_in=None, ln_out=None):
    """conv3x3 / conv1x1 / Linear on NHWC bf16 (see include/srb200.h: srb200_tapgemm).

    ``ln_in=(stats, wsum)``: LayerNorm of the input rows folded into this GEMM (``x`` is the RAW tensor, ``wp`` carries
    gamma, ``bias`` carries W beta + b); ``ln_out=(channels, eps)``: also return fp32 [B,H,W,2] (mean, rstd) of the
    stored rows for the next folded GEMM (appended last to the result tuple)."""
    _chk(x, 'x', torch.bfloat16)
    _chk(wp, 'wp', torch.bfloat16)
    b, hf, wf, cin = x.shape
    h, w = hf // src_r, wf // src_r
    d = L.TapGemmDesc(B=b, H=h, W=w, Cin=cin, src_r=src_r, Cout=cout, ksize=ksize, flip=int(flip), act=act,
                      act_slope=act_slope, alpha=alpha, mask_mode=mask_mode, mask_slope=mask_slope,
                      out_mode=out_mode, out_r=out_r, out_c=out_c, out_scale=out_scale)
    if out_mode == L.OUT_NHWC:
        out = torch.empty((b, h, w, cout), dtype=torch.bfloat16, device=x.device)
    elif out_mode == L.OUT_SHUFFLE:
        out = torch.empty((b, h * out_r, w * out_r, cout // (out_r * out_r)), dtype=torch.bfloat16, device=x.device)
    else:
        out = torch.empty((b, out_c, h, w), dtype=torch.float32, device=x.device)
    aux = torch.empty_like(out) if want_aux else None
    for t, name in ((mask_src, 'mask_src'), (residual, 'residual')):
        if t is not None:
            _chk(t, name, torch.bfloat16)
            if t.shape != out.shape:
                raise RuntimeError(f'{name} shape {tuple(t.shape)} != out shape {tuple(out.shape)}')
    if bias is not None:
        _chk(bias, 'bias', torch.float32)
        assert bias.numel() == cout
    out32 = torch.empty(out.shape, dtype=torch.float32, device=x.device) if want_f32 else None
    ext = None
    csum = None
    per_image = want_colsum == 'image_mean'  # [B, cout] per-sample means (RCAN's AdaptiveAvgPool2d(1))
    if want_colsum:  # fp32 [cout] column sums of the stored result, accumulated by the epilogue (bias gradient)
        if out_mode != L.OUT_NHWC or cout % 64 != 0:
            raise RuntimeError('want_colsum needs a plain NHWC output with cout % 64 == 0')
        csum = zeros_f32((b, cout) if per_image else (cout,), x.device)
    stats_out = torch.empty((b, h, w, 2), dtype=torch.float32, device=x.device) if ln_out is not None else None
    if ln_in is not None:
        _chk(ln_in[0], 'ln_in stats', torch.float32)
        _chk(ln_in[1], 'ln_in wsum', torch.float32)
        assert ln_in[0].numel() == 2 * b * h * w and ln_in[1].numel() == cout
    if (residual_f32 is not None or want_f32 or alpha_per_sample is not None or csum is not None or aux_grad or
            ln_in is not None or ln_out is not None):
        for t, name in ((residual_f32, 'residual_f32'), (alpha_per_sample, 'alpha_per_sample')):
            if t is not None:
                _chk(t, name, torch.float32)
        ext = ctypes.byref(L.TapGemmExt(residual_f32.data_ptr() if residual_f32 is not None else None,
                                        out32.data_ptr() if out32 is not None else None,
                                        alpha_per_sample.data_ptr() if alpha_per_sample is not None else None,
                                        1 if aux_grad else 0, 1 if per_image else 0,
                                        csum.data_ptr() if csum is not None else None,
                                        1.0 / (h * w) if per_image else 1.0,
                                        L.EXT_ALPHA_ON_BIAS if (alpha_on_bias and alpha_per_sample is not None) else 0,
                                        ln_in[0].data_ptr() if ln_in is not None else None,
                                        ln_in[1].data_ptr() if ln_in is not None else None,
                                        stats_out.data_ptr() if stats_out is not None else None,
                                        int(ln_out[0]) if ln_out is not None else 0,
                                        float(ln_out[1]) if ln_out is not None else 0.0))
    ev = PROBE.begin('tapgemm', (b, h, w, cin * src_r * src_r, cout, ksize)) if PROBE is not None else None
    L.check(L.load().srb200_tapgemm(ctypes.byref(d), _ptr(x), _ptr(wp), _ptr(bias), _ptr(mask_src), _ptr(residual),
                                    _ptr(out_shift), _ptr(out), _ptr(aux), ext, _stream()), 'tapgemm')
    if ev is not None:
        PROBE.end(ev)
    res = (out,) + ((aux,) if want_aux else ()) + ((out32,) if want_f32 else ()) + ((csum,) if want_colsum else ()) + \
        ((stats_out,) if ln_out is not None else ())
    return res if len(res) > 1 else out


def wgrad(dy, x, *, ksize, dy_r=1):
    """fp32 [taps, N, K] = sum_pixels dY^T X_shifted (see srb200_wgrad)."""
    _chk(dy, 'dy', torch.bfloat16)
    _chk(x, 'x', torch.bfloat16)
    b, h, w, k = x.shape
    n = dy.shape[-1] * dy_r * dy_r
    assert dy.shape[0] == b and dy.shape[1] == h * dy_r and dy.shape[2] == w * dy_r
    acc = zeros_f32((ksize * ksize, n, k), x.device)
    probed('wgrad', (b, h, w, n, k, ksize), lambda: L.check(
        L.load().srb200_wgrad(_ptr(dy), _ptr(x), _ptr(acc), b, h, w, n, k, ksize, dy_r, _stream()), 'wgrad'))
    return acc


def colsum(dy, r=1):
    """bias gradient: fp32 [r*r*C] column sums of an NHWC bf16 tensor (phase-separated when r > 1)."""
    _chk(dy, 'dy', torch.bfloat16)
    c = dy.shape[-1]
    rows = dy.numel() // c
    out = zeros_f32((r * r * c,), dy.device)
    L.check(L.load().srb200_colsum(_ptr(dy), _ptr(out), rows, c, r, dy.shape[2] if dy.dim() == 4 else 0, _stream()),
            'colsum')
    return out


def act_bwd(g, y, slope=0.0):
    """g * act'(y) for ReLU / LeakyReLU given the forward output y."""
    _chk(g, 'g', torch.bfloat16)
    _chk(y, 'y', torch.bfloat16)
    out = torch.empty_like(g)
    L.check(L.load().srb200_act_bwd(_ptr(g), _ptr(y), _ptr(out), g.numel(), float(slope), _stream()), 'act_bwd')
    return out


# ------------------------------------------------------------------ RCAN channel attention
def channel_pool(t):
    """AdaptiveAvgPool2d(1) (rcan_arch.py:19) of NHWC bf16 -> fp32 [B, C]."""
    _chk(t, 't', torch.bfloat16)
    b, h, w, c = t.shape
    out = zeros_f32((b, c), t.device)
    L.check(L.load().srb200_channel_pool(_ptr(t), _ptr(out), b, h * w, c, _stream()), 'channel_pool')
    return out


def channel_dot(a, m, scale=1.0):
    """fp32 [B, C] = scale * sum_hw a * m."""
    _chk(a, 'a', torch.bfloat16)
    _chk(m, 'm', torch.bfloat16)
    b, h, w, c = a.shape
    out = zeros_f32((b, c), a.device)
    L.check(L.load().srb200_channel_dot(_ptr(a), _ptr(m), _ptr(out), b, h * w, c, float(scale), _stream()),
            'channel_dot')
    return out


def ca_fc(p, w1, b1, w2, b2):
    """z = relu(W1 p + b1), s = sigmoid(W2 z + b2) (rcan_arch.py:19-20); weights as stored ([Cr,C,1,1], [C,Cr,1,1])."""
    for t, n in ((p, 'p'), (w1, 'w1'), (b1, 'b1'), (w2, 'w2'), (b2, 'b2')):
        _chk(t, n, torch.float32)
    b, c = p.shape
    cr = w1.shape[0]
    z = torch.empty((b, cr), dtype=torch.float32, device=p.device)
    s = torch.empty((b, c), dtype=torch.float32, device=p.device)
    L.check(L.load().srb200_ca_fc(_ptr(p), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(z), _ptr(s), b, c, cr,
                                  _stream()), 'ca_fc')
    return z, s


def ca_apply(t, x, s, res_scale, x32=None, want_f32=False):
    """y = x + res_scale * t * s[b, c] (rcan_arch.py:24,45-46); optional fp32 skip stream in / out."""
    _chk(t, 't', torch.bfloat16)
    if x32 is not None:
        _chk(x32, 'x32', torch.float32)
    else:
        _chk(x, 'x', torch.bfloat16)
    b, h, w, c = t.shape
    y = torch.empty_like(t)
    y32 = torch.empty(t.shape, dtype=torch.float32, device=t.device) if want_f32 else None
    L.check(L.load().srb200_ca_apply(_ptr(t), _ptr(x) if x32 is None else None, _ptr(x32), _ptr(s), _ptr(y),
                                     _ptr(y32), b, h * w, c, float(res_scale), _stream()), 'ca_apply')
    return (y, y32) if want_f32 else y


def ca_forward(t, x, p, w1, b1, w2, b2, res_scale, x32=None, want_f32=False):
    """Fused ``ca_fc`` + ``ca_apply`` (one launch): returns (y | (y, y32), z, s)."""
    _chk(t, 't', torch.bfloat16)
    if x32 is not None:
        _chk(x32, 'x32', torch.float32)
    else:
        _chk(x, 'x', torch.bfloat16)
    for tt, n in ((p, 'p'), (w1, 'w1'), (b1, 'b1'), (w2, 'w2'), (b2, 'b2')):
        _chk(tt, n, torch.float32)
    b, h, w, c = t.shape
    cr = w1.shape[0]
    zs = torch.empty((b * (cr + c),), dtype=torch.float32, device=t.device)
    z, s = zs[:b * cr].view(b, cr), zs[b * cr:].view(b, c)
    y = torch.empty_like(t)
    y32 = torch.empty(t.shape, dtype=torch.float32, device=t.device) if want_f32 else None
    probed('ca_forward', (b, h * w, c, x32 is not None), lambda: L.check(L.load().srb200_ca_forward(
        _ptr(t), _ptr(x) if x32 is None else None, _ptr(x32), _ptr(p), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(z),
        _ptr(s), _ptr(y), _ptr(y32), b, h * w, c, cr, float(res_scale), _stream()), 'ca_forward'))
    return ((y, y32) if want_f32 else y), z, s


def ca_backward(g, t, s, z, p, w1, w2, res_scale):
    """Fused ``channel_dot`` -> ``ca_fc_bwd`` -> ``ca_apply_bwd`` (one persistent launch).

    Returns (gw1, gb1, gw2, gb2, gt, colsum(gt)).  Needs zeroed scratch: taken from the active zero arena."""
    _chk(g, 'g', torch.bfloat16)
    _chk(t, 't', torch.bfloat16)
    b, h, w, c = g.shape
    cr = z.shape[1]
    dev = g.device
    gs = zeros_f32((b, c), dev)
    cs = zeros_f32((c,), dev)
    sync = zeros_f32((4,), dev)
    out = torch.empty((2 * cr * c + cr + c,), dtype=torch.float32, device=dev)
    o = 0
    gw1 = out[o:o + cr * c].view(cr, c, 1, 1); o += cr * c
    gw2 = out[o:o + cr * c].view(c, cr, 1, 1); o += cr * c
    gb1 = out[o:o + cr]; o += cr
    gb2 = out[o:o + c]
    gt = torch.empty_like(g)
    probed('ca_backward', (b, h * w, c), lambda: L.check(L.load().srb200_ca_backward(
        _ptr(g), _ptr(t), _ptr(s), _ptr(z), _ptr(p), _ptr(w1), _ptr(w2), _ptr(gs), _ptr(gw1), _ptr(gb1), _ptr(gw2),
        _ptr(gb2), _ptr(gt), _ptr(cs), _ptr(sync), b, h * w, c, cr, float(res_scale), _stream()), 'ca_backward'))
    return gw1, gb1, gw2, gb2, gt, cs


def ca_fc_bwd(gs, s, z, p, w1, w2):
    b, c = s.shape
    cr = z.shape[1]
    dev = s.device
    gw1 = torch.empty((cr, c, 1, 1), dtype=torch.float32, device=dev)
    gb1 = torch.empty((cr,), dtype=torch.float32, device=dev)
    gw2 = torch.empty((c, cr, 1, 1), dtype=torch.float32, device=dev)
    gb2 = torch.empty((c,), dtype=torch.float32, device=dev)
    gp = torch.empty((b, c), dtype=torch.float32, device=dev)
    L.check(L.load().srb200_ca_fc_bwd(_ptr(gs), _ptr(s), _ptr(z), _ptr(p), _ptr(w1), _ptr(w2), _ptr(gw1), _ptr(gb1),
                                      _ptr(gw2), _ptr(gb2), _ptr(gp), b, c, cr, _stream()), 'ca_fc_bwd')
    return gw1, gb1, gw2, gb2, gp


def ca_apply_bwd(g, s, gp, res_scale, want_colsum=False):
    """gt = res_scale * g * s[b, c] + gp[b, c] / HW  (+ its fp32 column sums: the bias gradient of conv2)."""
    _chk(g, 'g', torch.bfloat16)
    b, h, w, c = g.shape
    gt = torch.empty_like(g)
    cs = zeros_f32((c,), g.device) if want_colsum else None
    L.check(L.load().srb200_ca_apply_bwd(_ptr(g), _ptr(s), _ptr(gp), _ptr(gt), b, h * w, c, float(res_scale),
                                         _ptr(cs), _stream()), 'ca_apply_bwd')
    return (gt, cs) if want_colsum else gt


# ------------------------------------------------------------------ image entry / exit (callers' side)
def patch_items(rows, device):
    """Device table of ``srb200_patch_item``: ``rows`` = (image uint8 [H,W,C] CUDA tensor, top, left, flags)."""
    import numpy as np
    tab = np.zeros(len(rows), dtype=np.dtype(L.PATCH_ITEM_FIELDS))
    for i, (img, top, left, flags) in enumerate(rows):
        tab[i] = (img.data_ptr(), img.stride(0) * img.element_size(), top, left, flags, 0)
    return torch.from_numpy(tab.view(np.uint8).reshape(len(rows), -1).copy()).to(device, non_blocking=True)


def patch_from_u8(images, tops, lefts, flags, ph, pw, bgr2rgb=True, div=255.0):
    """Batched crop + augment + img2tensor (transforms.py:28-96,166-225; img_util.py:9-37) on the GPU.

    ``images``: list of uint8 [H, W, C] CUDA tensors (rows contiguous); item i is the ``ph x pw`` crop of images[i] at
    (tops[i], lefts[i]) with flags[i] (bit 0 hflip, bit 1 vflip, bit 2 rot90).  Returns fp32 [n, C, oh, ow]."""
    n = len(images)
    c = images[0].shape[2]
    rot = [bool(f & 4) for f in flags]
    if any(rot) and not all(rot) and ph != pw:
        raise RuntimeError('mixed rot90 flags need square patches')
    oh, ow = (pw, ph) if rot[0] else (ph, pw)
    for img, t, l in zip(images, tops, lefts):
        if img.dtype != torch.uint8 or not img.is_cuda or img.dim() != 3 or img.stride(2) != 1 or img.stride(1) != c:
            raise RuntimeError('patch_from_u8 wants uint8 [H, W, C] CUDA images with contiguous rows')
        if t < 0 or l < 0 or t + ph > img.shape[0] or l + pw > img.shape[1]:
            raise RuntimeError('crop outside the image')
    dev = images[0].device
    tab = patch_items(list(zip(images, tops, lefts, flags)), dev)
    out = torch.empty((n, c, oh, ow), dtype=torch.float32, device=dev)
    L.check(L.load().srb200_patch_from_u8(_ptr(tab), n, c, ph, pw, int(bgr2rgb), float(div), _ptr(out), _stream()),
            'patch_from_u8')
    return out


def tile_blend_add(sr_tile, acc, y0, x0, ov, guard):
    """acc [C, H, W] += ramp-weighted sr_tile [C, th, tw]; ``ov`` = (top, bottom, left, right) overlaps in pixels."""
    _chk(sr_tile, 'sr_tile', torch.float32)
    _chk(acc, 'acc', torch.float32)
    c, th, tw = sr_tile.shape[-3:]
    L.check(L.load().srb200_tile_blend_add(_ptr(sr_tile), _ptr(acc), c, th, tw, acc.shape[-2], acc.shape[-1], int(y0),
                                           int(x0), int(ov[0]), int(ov[1]), int(ov[2]), int(ov[3]), int(guard),
                                           _stream()), 'tile_blend_add')


def tensor2img_u8(src, lo=0.0, hi=1.0, rgb2bgr=True, out=None):
    """tensor2img (img_util.py:40-96) on the GPU: fp32 [C, H, W] -> uint8 [H, W, C]."""
    _chk(src, 'src', torch.float32)
    c, h, w = src.shape[-3:]
    if out is None:
        out = torch.empty((h, w, c), dtype=torch.uint8, device=src.device)
    L.check(L.load().srb200_tensor2img_u8(_ptr(src), _ptr(out), c, h, w, float(lo), float(hi), int(rgb2bgr), _stream()),
            'tensor2img_u8')
    return out
