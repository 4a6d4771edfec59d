"""autograd.Function wrappers around the srb200 C-ABI -- the counterpart of the reference's
``basicsr/ops/<op>/<op>.py`` modules (e.g. ops/fused_act/fused_act.py:30-78): ``ctx.save_for_backward``
in forward, explicit kernels in backward, gradients returned through normal autograd so DDP's
bucket hooks fire during backward (basicsr/models/base_model.py:98-99).

Inside a network every activation is an NHWC bf16 tensor whose channel count is padded to a multiple
of 64 ("NHWC64"); nn.Parameters stay fp32 in the reference's OIHW / [out,in] layout and are repacked
to bf16 GEMM operands on the fly (cached per parameter version, never stored in the state dict).
"""
import threading
import weakref

import torch
from torch.autograd import Function

from ... import _lib as L
from . import raw

ACT_CODES = {None: L.ACT_NONE, 'relu': L.ACT_RELU, 'lrelu': L.ACT_LRELU, 'gelu': L.ACT_GELU}


def pad64(c):
    return (c + 63) // 64 * 64


# ------------------------------------------------------------------ derived caches
_perm_cache = {}


def shuffle_perm(c_feat, r, device):
    """packed out-channel p = ij*c_feat + c  ->  original nn.Conv2d out-channel c*r*r + ij, so that one
    GEMM N-tile holds a single PixelShuffle phase (arch_util.py:134-138)."""
    key = ('shuffle', c_feat, r, str(device))
    if key not in _perm_cache:
        p = torch.arange(c_feat * r * r, device=device)
        _perm_cache[key] = ((p % c_feat) * (r * r) + p // c_feat).to(torch.int32)
    return _perm_cache[key]


class PackBook:
    """All bf16 GEMM operands of ONE network, repacked by ONE kernel launch per optimizer step.

    A training step of EDSR-L touches 137 packed operands (fprop + dgrad layouts of 69 weights), RCAN 829;
    packing them one launch each costs more than the packing itself (8-13 us per launch against ~65 us of HBM
    traffic in total).  The book owns a persistent bf16 buffer per (weight, layout) and a device table of
    ``srb200_pack_item`` rows; :meth:`refresh` re-derives every buffer from the live fp32 parameters with
    ``srb200_pack_weights``.  Entries register themselves the first time an arch asks for them
    (:func:`_packed`), so the book needs no knowledge of the architecture.

    Eager mode: the first :func:`_packed` call that finds a parameter newer than its pack refreshes the whole
    book.  CUDA-graph mode: the refresh launch is captured at the head of the first segment's forward graph
    (archs/graphed.py) and the captured kernels read the persistent buffers.
    """

    def __init__(self):
        self.entries = {}     # key -> [weakref(weight), buf, spec dict, tag]
        self.table = None     # (device tensor, n_items, total_chunks)
        self.table_ptrs = None
        self.capture_ok = False   # set by the graph capture driver once the refresh launch is part of the graph

    def _rows(self):
        rows, ptrs = [], []
        for key in list(self.entries):
            wref, buf, spec, _ = self.entries[key]
            w = wref()
            if w is None:
                del self.entries[key]
                continue
            rows.append(dict(spec, src=w, dst=buf))
            ptrs.append(w.data_ptr())
        return rows, ptrs

    def get(self, weight, kind, n_pad, k_pad, perm_out, perm_in):
        key = (id(weight), kind, n_pad, k_pad, id(perm_out) if perm_out is not None else 0,
               id(perm_in) if perm_in is not None else 0)
        e = self.entries.get(key)
        capturing = torch.cuda.is_current_stream_capturing()
        if e is None or e[0]() is not weight:
            if capturing:
                return None
            w = weight.detach()
            if not w.is_contiguous():
                return None
            co, ci = w.shape[0], w.shape[1]
            taps = w.numel() // (co * ci)
            buf = raw.pack_weight(w, n_pad, k_pad, perm_out=perm_out, perm_in=perm_in, transpose=(kind == 'dgrad'))
            spec = dict(Co=co, Ci=ci, taps=taps, Np=n_pad, Kp=k_pad, perm_out=perm_out, perm_in=perm_in,
                        transpose=(kind == 'dgrad'))
            self.entries[key] = [weakref.ref(weight), buf, spec, (weight.data_ptr(), weight._version)]
            self.table = None
            return buf
        if capturing:
            return e[1] if self.capture_ok else None
        if e[3] != (weight.data_ptr(), weight._version):
            self.refresh()
        return e[1]

    def get_bias(self, bias, n_pad, perm_out, fill=None):
        """fp32 [n_pad] bias in the packed channel order (zero padded): an fp32 item of the same batched launch.
        ``fill = (index, value)``: that one pad element holds ``value`` instead of zero."""
        key = (id(bias), 'bias', n_pad, 1, id(perm_out) if perm_out is not None else 0, fill or 0)
        e = self.entries.get(key)
        capturing = torch.cuda.is_current_stream_capturing()
        if e is None or e[0]() is not bias:
            if capturing or bias.dtype != torch.float32 or not bias.is_contiguous():
                return None
            buf = torch.empty((n_pad,), dtype=torch.float32, device=bias.device)
            spec = dict(Co=bias.numel(), Ci=1, taps=1, Np=n_pad, Kp=1, perm_out=perm_out, perm_in=None, transpose=2)
            if fill is not None:
                spec.update(fill_index=int(fill[0]), alpha=float(fill[1]))
            self.entries[key] = [weakref.ref(bias), buf, spec, None]  # tag None: filled by the refresh below
            self.table = None
            self.refresh(force=True)
            return buf
        if capturing:
            return e[1] if self.capture_ok else None
        if e[3] != (bias.data_ptr(), bias._version):
            self.refresh()
        return e[1]

    def invalidate(self):
        """Forget every version tag: the next :meth:`get` / :meth:`refresh` repacks the whole book.  For writers that
        change parameters without bumping ``_version`` (``p.data.mul_()`` of the reference EMA loop,
        base_model.py:81-82; raw-pointer kernels such as ``srb200_multi_axpby``)."""
        for e in self.entries.values():
            e[3] = None

    def refresh(self, force=False):
        """Repack every entry whose parameter changed (all of them after an optimizer step) in one launch."""
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            if self.table is None:
                raise RuntimeError('PackBook: the item table must be built before CUDA-graph capture')
            raw.pack_weights(*self.table)
            return
        stale = force
        ptrs = []
        for key in list(self.entries):
            e = self.entries[key]
            w = e[0]()
            if w is None:
                del self.entries[key]
                self.table = None
                continue
            tag = (w.data_ptr(), w._version)
            if tag != e[3]:
                stale = True
                e[3] = tag
            ptrs.append(tag[0])
        if not self.entries or not (stale or self.table is None):
            return
        if self.table is None or ptrs != self.table_ptrs:
            rows, self.table_ptrs = self._rows()
            self._keep = rows  # keeps perm tensors / buffers referenced by raw pointers alive
            self.table = raw.pack_items(rows, rows[0]['dst'].device)
        if stale:
            raw.pack_weights(*self.table)


_book_tls = threading.local()


class pack_book:
    """``with pack_book(book):`` -- operands requested inside (an arch's forward) are owned by ``book``; the
    backward of those layers finds the same book through the parameter."""

    def __init__(self, book):
        self.book = book

    def __enter__(self):
        self.prev = getattr(_book_tls, 'book', None)
        _book_tls.book = self.book
        return self.book

    def __exit__(self, *exc):
        _book_tls.book = self.prev
        return False


def _packed(weight, kind, n_pad, k_pad, perm_out=None, perm_in=None):
    """bf16 GEMM operand of an fp32 parameter.  Inside an arch (``pack_book``) it lives in the arch's
    :class:`PackBook`; otherwise it is cached on the parameter until the parameter is modified in place
    (optimizer step bumps ``_version``) or re-allocated (``.to(device)`` / deepcopy change ``data_ptr``)."""
    book = getattr(_book_tls, 'book', None)
    if book is None:
        ref = weight.__dict__.get('_srb200_book')
        book = ref() if ref is not None else None
    else:
        weight.__dict__['_srb200_book'] = weakref.ref(book)
    if book is not None:
        buf = book.get(weight, kind, n_pad, k_pad, perm_out, perm_in)
        if buf is not None:
            return buf
    if torch.cuda.is_current_stream_capturing():
        # inside a CUDA-graph capture the repack kernel must be part of the graph (it re-reads the live
        # parameter storage on every replay), so the host-side cache is bypassed
        return raw.pack_weight(weight.detach(), n_pad, k_pad, perm_out=perm_out, perm_in=perm_in,
                               transpose=(kind == 'dgrad'))
    cache = weight.__dict__.setdefault('_srb200_pack', {})
    tag = (weight.data_ptr(), weight._version)
    if cache.get('tag') != tag or not torch.is_grad_enabled():
        # (no_grad = evaluation of a possibly EMA-updated copy: ``p.data.mul_()`` leaves ``_version`` alone, so the
        # tag proves nothing there -- repack; one small launch per weight)
        cache.clear()
        cache['tag'] = tag
    if kind not in cache:
        w = weight.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        cache[kind] = raw.pack_weight(w, n_pad, k_pad, perm_out=perm_out, perm_in=perm_in,
                                      transpose=(kind == 'dgrad'))
    return cache[kind]


def _padded_bias(bias, n_pad, perm_out=None, fill=None):
    """fp32 [n_pad] bias in the packed channel order.  ``fill = (index, value)``: one PAD element holds ``value``."""
    if bias is None:
        return None
    if perm_out is None and bias.numel() == n_pad and bias.dtype == torch.float32:
        return bias.detach()  # already in the packed layout
    book = getattr(_book_tls, 'book', None)
    if book is None:
        ref = bias.__dict__.get('_srb200_book')
        book = ref() if ref is not None else None
    else:
        bias.__dict__['_srb200_book'] = weakref.ref(book)
    if book is not None:
        buf = book.get_bias(bias, n_pad, perm_out, fill)
        if buf is not None:
            return buf
    capturing = torch.cuda.is_current_stream_capturing()  # see _packed: no host-side cache inside a capture
    cache = bias.__dict__.setdefault('_srb200_pack', {})
    tag = (bias.data_ptr(), bias._version)
    if cache.get('tag') != tag or not torch.is_grad_enabled():
        cache.clear()
        cache['tag'] = tag
    ck = ('b', fill)
    if capturing or ck not in cache:
        b = bias.detach().float()
        if perm_out is not None:
            key = ('safe', id(perm_out))
            if key not in _perm_cache:
                _perm_cache[key] = (perm_out.clamp(min=0).long(), perm_out >= 0)
            safe, valid = _perm_cache[key]
            bp = torch.where(valid, b[safe], torch.zeros((), device=b.device))
        elif b.numel() == n_pad:
            bp = b
        else:
            bp = torch.zeros(n_pad, device=b.device)
            bp[:b.numel()] = b
        bp = bp.contiguous()
        if fill is not None:
            bp = bp.clone() if bp is b else bp
            bp[fill[0]] = fill[1]
        if capturing:
            return bp
        cache[ck] = bp
    return cache[ck]


# The bias gradient of a layer is the column sum of its dY, and dY is almost always the OUTPUT of the previous
# backward tap-GEMM.  That kernel's epilogue can accumulate the column sums for free (``want_colsum``); the
# producer parks them here and the consumer -- the very next backward function of a sequential trunk -- picks
# them up by the identity of the gradient tensor's storage.  One slot per thread: anything else falls back to
# the standalone colsum kernel.
_colsum_tls = threading.local()


def _stash_colsum(t, csum):
    # the slot keeps `t` alive, so no other tensor can be handed its storage while the entry is valid
    _colsum_tls.slot = (t, t._version, csum)


def _colsum_of(g):
    """fp32 column sums of the NHWC bf16 gradient ``g``: taken from the producer's epilogue when available."""
    slot = getattr(_colsum_tls, 'slot', None)
    _colsum_tls.slot = None
    if slot is not None:
        t, ver, csum = slot
        if t.data_ptr() == g.data_ptr() and t.shape == g.shape and t.stride() == g.stride() and t._version == ver:
            return csum
    return raw.colsum(g)


# ------------------------------------------------------------------ NCHW fp32 image -> NHWC64 bf16
class _ImageToNHWC(Function):
    """(x - mean) * img_range (edsr_arch.py:53) fused with the NCHW fp32 -> NHWC bf16 conversion."""

    @staticmethod
    def forward(ctx, x, shift, scale, c_pad):
        ctx.c = x.shape[1]
        ctx.scale = scale
        return raw.nchw_to_nhwc(x.contiguous().float(), c_pad, shift=shift, scale=scale)

    @staticmethod
    def backward(ctx, g):
        gx = raw.nhwc_to_nchw(g.contiguous(), ctx.c, shift=None, scale=ctx.scale)
        return gx, None, None, None


def image_to_nhwc(x, shift, scale, c_pad=64):
    return _ImageToNHWC.apply(x, shift, scale, c_pad)


# ------------------------------------------------------------------ generic conv / linear
class _ConvNHWC(Function):
    """y = act(conv(x) + b) * alpha (+ residual), optional fused PixelShuffle store.

    Replaces nn.Conv2d(cin, cout, k, 1, k//2) [+ ReLU / LeakyReLU] [+ skip add] [+ nn.PixelShuffle(r)]
    (arch_util.py:79-88,134-138; edsr_arch.py:44-56; swinir_arch.py:532,818,834-837).
    """

    @staticmethod
    def forward(ctx, x, weight, bias, residual, ksize, act, slope, alpha, shuffle_r, residual32=None,
                want_f32=False):
        cout, cin = weight.shape[0], weight.shape[1]
        k_pad = x.shape[-1]
        assert k_pad == pad64(cin), f'input has {k_pad} channels, expected {pad64(cin)}'
        if shuffle_r > 1:
            c_feat = cout // (shuffle_r * shuffle_r)
            assert c_feat % 64 == 0, 'fused PixelShuffle needs num_feat % 64 == 0'
            perm = shuffle_perm(c_feat, shuffle_r, x.device)
            n_pad = cout
        else:
            perm = None
            n_pad = pad64(cout)
        wp = _packed(weight, 'fprop', n_pad, k_pad, perm_out=perm)
        bp = _padded_bias(bias, n_pad, perm)
        y = raw.tapgemm(x, wp, ksize=ksize, cout=n_pad, bias=bp, act=ACT_CODES[act], act_slope=slope, alpha=alpha,
                        residual=residual if residual32 is None else None, residual_f32=residual32,
                        want_f32=want_f32, out_mode=L.OUT_SHUFFLE if shuffle_r > 1 else L.OUT_NHWC,
                        out_r=shuffle_r)
        y32 = None
        if want_f32:
            y, y32 = y
        ctx.save_for_backward(x, weight, bias, y if act is not None else None)
        ctx.cfg = (ksize, act, slope, alpha, shuffle_r, n_pad, k_pad)
        raw.stash_backward_scratch(ctx, ksize * ksize * n_pad * k_pad + n_pad + k_pad + 32, x.device)
        ctx.perm = perm
        ctx.has_res = residual is not None
        if want_f32:
            ctx.mark_non_differentiable(y32)
            ctx.set_materialize_grads(False)  # (no zero-filled gradient for the fp32 twin, see _RCAB)
            return y, y32
        return y

    @staticmethod
    def backward(ctx, g, *unused):
        if g is None:
            return (None,) * 11
        x, weight, bias, y = ctx.saved_tensors
        ksize, act, slope, alpha, shuffle_r, n_pad, k_pad = ctx.cfg
        perm = ctx.perm
        g = g.contiguous()
        g_res = g if ctx.has_res else None
        gbp = None
        if act is not None:
            assert act in ('relu', 'lrelu'), 'gelu backward goes through the fused MLP function'
            g = raw.act_bwd(g, y, slope if act == 'lrelu' else 0.0)
        elif bias is not None and ctx.needs_input_grad[2] and shuffle_r == 1:
            gbp = _colsum_of(g)
        _colsum_tls.slot = None
        gx = gw = gb = None
        with raw.backward_arena(ctx, x.device, ksize * ksize * n_pad * k_pad + n_pad + k_pad + 32):
            if ctx.needs_input_grad[1]:
                acc = raw.wgrad(g, x, ksize=ksize, dy_r=shuffle_r)
            if bias is not None and ctx.needs_input_grad[2] and gbp is None:
                gbp = raw.colsum(g, r=shuffle_r)
            items, prms = [], []
            if ctx.needs_input_grad[1]:
                items.append(('w', acc, weight.shape, perm, None, alpha))
                prms.append(weight)
            if bias is not None and ctx.needs_input_grad[2]:
                items.append(('b', gbp, bias.numel(), perm, alpha))
                prms.append(bias)
            if items:
                grads = raw.finalize_grads(items, prms)
                if ctx.needs_input_grad[1]:
                    gw = grads[0]
                if bias is not None and ctx.needs_input_grad[2]:
                    gb = grads[-1]
            if ctx.needs_input_grad[0]:
                wpt = _packed(weight, 'dgrad', n_pad, k_pad, perm_out=perm)
                if ksize == 3 and k_pad % 64 == 0:
                    # the layer below is a conv with a bias: hand it the column sums of its dY
                    gx, cs = raw.tapgemm(g, wpt, ksize=ksize, cout=k_pad, alpha=alpha, flip=True, src_r=shuffle_r,
                                         want_colsum=True)
                    _stash_colsum(gx, cs)
                else:
                    gx = raw.tapgemm(g, wpt, ksize=ksize, cout=k_pad, alpha=alpha, flip=True, src_r=shuffle_r)
        return gx, gw, gb, g_res, None, None, None, None, None, None, None


def conv_nhwc(x, weight, bias=None, residual=None, act=None, slope=0.0, alpha=1.0, shuffle_r=1, residual32=None,
              want_f32=False):
    """conv3x3 (weight [Co,Ci,3,3]), conv1x1 or Linear (weight [Co,Ci,1,1] / [Co,Ci]) on NHWC64 bf16.

    ``residual`` is the differentiable bf16 skip input; ``residual32`` (optional, same values in fp32, never
    differentiated) is what the epilogue actually adds when given.  ``want_f32`` also returns the fp32 result."""
    ksize = weight.shape[-1] if weight.dim() == 4 else 1
    return _ConvNHWC.apply(x, weight, bias, residual, ksize, act, slope, float(alpha), shuffle_r, residual32,
                           want_f32)


# ------------------------------------------------------------------ fused ResidualBlockNoBN
class _ResBlockNoBN(Function):
    """x + conv2(relu(conv1(x))) * res_scale  (arch_util.py:85-88) as two tap-GEMMs whose epilogues
    carry bias+ReLU and bias*res_scale+identity; backward fuses the ReLU mask and the skip-gradient
    add into the two dgrad epilogues (no standalone elementwise pass in either direction)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, res_scale):
        c = w1.shape[0]
        cp = pad64(c)
        assert x.shape[-1] == cp
        h = raw.tapgemm(x, _packed(w1, 'fprop', cp, cp), ksize=3, cout=cp, bias=_padded_bias(b1, cp), act=L.ACT_RELU)
        y = raw.tapgemm(h, _packed(w2, 'fprop', cp, cp), ksize=3, cout=cp, bias=_padded_bias(b2, cp), alpha=res_scale,
                        residual=x)
        ctx.save_for_backward(x, h, w1, b1, w2, b2)
        ctx.res_scale = res_scale
        raw.stash_backward_scratch(ctx, 2 * 9 * cp * cp + 3 * cp + 32, x.device)
        return y

    @staticmethod
    def backward(ctx, g):
        x, h, w1, b1, w2, b2 = ctx.saved_tensors
        s = ctx.res_scale
        cp = x.shape[-1]
        g = g.contiguous()
        with raw.backward_arena(ctx, x.device, 2 * 9 * cp * cp + 3 * cp + 32):
            return _ResBlockNoBN._backward(ctx, g, x, h, w1, b1, w2, b2, s, cp)

    @staticmethod
    def _backward(ctx, g, x, h, w1, b1, w2, b2, s, cp):
        gw1 = gb1 = gw2 = gb2 = gx = None
        items, prms = [], []
        if b2 is not None and ctx.needs_input_grad[4]:
            items.append(('b', _colsum_of(g), b2.numel(), None, s))  # from the epilogue that produced g, if any
            prms.append(b2)
        _colsum_tls.slot = None
        dev = x.device
        if ctx.needs_input_grad[3]:
            with raw.side_branch(dev):  # weight gradients next to the data-gradient chain
                items.append(('w', raw.wgrad(g, h, ksize=3), w2.shape, None, None, s))
                prms.append(w2)
        # d(pre-activation of conv1) = dgrad_conv2(s*g) masked by relu'(h); its column sums = conv1's bias gradient
        want_gb1 = b1 is not None and ctx.needs_input_grad[2]
        gh = raw.tapgemm(g, _packed(w2, 'dgrad', cp, cp), ksize=3, cout=cp, alpha=s, flip=True, mask_src=h,
                         mask_mode=L.MASK_SIGN, mask_slope=0.0, want_colsum=want_gb1)
        if want_gb1:
            gh, cs = gh
            items.append(('b', cs, b1.numel(), None, 1.0))
            prms.append(b1)
        if ctx.needs_input_grad[1]:
            with raw.side_branch(dev):
                items.append(('w', raw.wgrad(gh, x, ksize=3), w1.shape, None, None, 1.0))
                prms.append(w1)
        if ctx.needs_input_grad[0]:
            gx, cs = raw.tapgemm(gh, _packed(w1, 'dgrad', cp, cp), ksize=3, cout=cp, flip=True, residual=g,
                                 want_colsum=True)
            _stash_colsum(gx, cs)
        raw.side_join(dev)
        grads = iter(raw.finalize_grads(items, prms)) if items else iter(())  # one launch for all four gradients
        if b2 is not None and ctx.needs_input_grad[4]:
            gb2 = next(grads)
        if ctx.needs_input_grad[3]:
            gw2 = next(grads)
        if want_gb1:
            gb1 = next(grads)
        if ctx.needs_input_grad[1]:
            gw1 = next(grads)
        return gx, gw1, gb1, gw2, gb2, None


def res_block_nobn(x, w1, b1, w2, b2, res_scale):
    return _ResBlockNoBN.apply(x, w1, b1, w2, b2, float(res_scale))


# ------------------------------------------------------------------ conv_last + exit
def _tap_folded_weight(weight, k_pad, tp):
    """fp32 [tp, k_pad]: row tap*C + c = weight[c, :, ky, kx] (tap = 3*ky + kx), zero padded."""
    cout, cin = weight.shape[0], weight.shape[1]
    wf = torch.zeros((tp, k_pad), dtype=torch.float32, device=weight.device)
    wf[:9 * cout, :cin] = weight.detach().permute(2, 3, 0, 1).reshape(9 * cout, cin)
    return wf


class _ConvToImage(Function):
    """conv_last (F -> num_out_ch <= 7, 3x3) fused with ``x / img_range + mean`` and the NHWC bf16 -> NCHW fp32 exit
    (edsr_arch.py:58-59; rcan_arch.py:132-133; swinir_arch.py:900,920).

    The conv is HBM-bound on its input, so it runs as ONE 1x1 tap-GEMM producing the nine per-tap partial products
    of every pixel (T[p][tap*C+c], fp32) followed by a 9-tap stencil sum over the tiny T tensor
    (srb200_tap_stencil) -- the input is read once, not nine times.  Backward: the image gradient is spread to the
    same tap-major channel layout (srb200_tap_im2col), after which the data and weight gradients are single-tap
    GEMMs too and the bias gradient is three of its column sums."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_scale, out_shift):
        cout, cin = weight.shape[0], weight.shape[1]
        assert cout <= 7 and weight.shape[-1] == 3, 'fused image exit: 3x3 conv with up to 7 output channels'
        k_pad = x.shape[-1]
        assert k_pad == pad64(cin)
        tp = pad64(9 * cout)
        wf = _tap_folded_weight(weight, k_pad, tp)
        _, t32 = raw.tapgemm(x, wf.to(torch.bfloat16).view(1, tp, k_pad), ksize=1, cout=tp, want_f32=True)
        y = raw.tap_stencil(t32, cout, bias=bias.detach() if bias is not None else None, shift=out_shift,
                            scale=out_scale)
        ctx.save_for_backward(x, weight, bias, wf)  # (wf: the backward's data-gradient operand is its transpose)
        ctx.out_scale = out_scale
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, bias, wf = ctx.saved_tensors
        cout, cin = weight.shape[0], weight.shape[1]
        k_pad = x.shape[-1]
        tp = pad64(9 * cout)
        # G[p][tap*C + c] = out_scale * g[c](p - tap): d(conv out) in the tap-major layout of the forward's T
        gn = raw.tap_im2col(g.contiguous().float(), tp, scale=ctx.out_scale)
        gx = gw = gb = None
        with raw.zero_arena(x.device, tp * k_pad + tp + k_pad + 64):
            if ctx.needs_input_grad[1]:
                acc = raw.wgrad(gn, x, ksize=1)  # [1, tp, k_pad] = G^T X
                gw = acc[0, :9 * cout, :cin].reshape(3, 3, cout, cin).permute(2, 3, 0, 1).contiguous()
            if bias is not None and ctx.needs_input_grad[2]:
                gb = raw.colsum(gn)[4 * cout:5 * cout].clone()  # centre tap: every pixel of g exactly once
            if ctx.needs_input_grad[0]:
                wd = wf.t().contiguous().to(torch.bfloat16).view(1, k_pad, tp)
                gx, cs = raw.tapgemm(gn, wd, ksize=1, cout=k_pad, want_colsum=True)
                _stash_colsum(gx, cs)
        return gx, gw, gb, None, None


class _NearestUp2(Function):
    """F.interpolate(scale_factor=2, mode='nearest') on NHWC bf16 (swinir_arch.py:911-912)."""

    @staticmethod
    def forward(ctx, x):
        return raw.nearest_up2(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return raw.nearest_up2(g.contiguous(), inverse=True)


def nearest_up2(x):
    return _NearestUp2.apply(x)


class _ShuffleToImage(Function):
    """nn.PixelShuffle(r) of the first r*r*C channels of an NHWC tensor, fused exit to the NCHW fp32 image with
    ``x * scale + shift`` -- the tail of SwinIR's 'pixelshuffledirect' branch (UpsampleOneStep, swinir_arch.py:669-690,
    followed by :920).  ``x32`` carries the conv result in fp32 (never differentiated): the image is not rounded to
    bf16 on its way out."""

    @staticmethod
    def forward(ctx, x, x32, c_img, r, scale, shift):
        cc = c_img * r * r
        ctx.cfg = (c_img, r, scale, x.shape[-1])
        hr = raw.pixel_shuffle_nhwc(x32[..., :cc].contiguous(), r)        # fp32 [B, rH, rW, c_img], bit-exact remap
        out = hr.permute(0, 3, 1, 2) * scale
        if shift is not None:
            out = out + shift.view(1, -1, 1, 1)
        return out.contiguous()

    @staticmethod
    def backward(ctx, g):
        c_img, r, scale, cp = ctx.cfg
        gn = raw.nchw_to_nhwc(g.contiguous().float(), (c_img + 7) // 8 * 8, shift=None, scale=scale)[..., :c_img].contiguous()
        lr = raw.pixel_shuffle_nhwc(gn, r, inverse=True)                  # bf16 [B, H, W, c_img*r*r]
        gx = torch.zeros(lr.shape[:3] + (cp,), dtype=torch.bfloat16, device=g.device)
        gx[..., :lr.shape[-1]] = lr
        return gx, None, None, None, None, None


def shuffle_to_image(x, x32, c_img, r, scale, shift):
    return _ShuffleToImage.apply(x, x32, c_img, r, float(scale), shift)


def conv_to_image(x, weight, bias, out_scale, out_shift):
    """conv_last + ``x * out_scale + out_shift`` + NCHW fp32 exit.  Up to 7 output bands take the fused tap-stencil
    exit; more (hyperspectral stacks) run the generic tap-GEMM and a layout kernel, the affine on the small image."""
    if weight.shape[0] <= 7 and weight.shape[-1] == 3:
        return _ConvToImage.apply(x, weight, bias, float(out_scale), out_shift)
    y = _NHWCToNCHW.apply(conv_nhwc(x, weight, bias), weight.shape[0]) * float(out_scale)
    return y + out_shift.view(1, -1, 1, 1) if out_shift is not None else y


class _NHWCToNCHW(Function):
    """[B,H,W,Cp] bf16 -> [B,C,H,W] fp32 (first C channels) and back."""

    @staticmethod
    def forward(ctx, t, c):
        ctx.c_pad = t.shape[-1]
        return raw.nhwc_to_nchw(t.contiguous(), c)

    @staticmethod
    def backward(ctx, g):
        return raw.nchw_to_nhwc(g.contiguous().float(), ctx.c_pad), None


# ------------------------------------------------------------------ fused RCAB (RCAN)
def rcab_backward_floats(batch, cp):
    """Zeroed fp32 scratch of one RCAB backward: two split-K weight-gradient accumulators, column sums, d s, sync."""
    return 2 * 9 * cp * cp + 3 * cp + batch * cp + 64


class _RCAB(Function):
    """x + res_scale * CA(conv2(relu(conv1(x))))  (rcan_arch.py:36-46, ChannelAttention :16-24).

    forward : tap-GEMM(+bias+ReLU) ; tap-GEMM(+bias, +pool in the epilogue) ; fused FC+sigmoid+scale+residual pass
    backward: one persistent launch (d s ; FC backward ; d t + colsum) ; wgrad/colsum/dgrad(+ReLU mask) for conv2 ;
              wgrad/colsum/dgrad(+skip gradient) for conv1
    """

    @staticmethod
    def forward(ctx, x, x32, w1, b1, w2, b2, wa1, ba1, wa2, ba2, res_scale):
        # x is the differentiable bf16 stream; x32 (optional, never differentiated) is the same stream in
        # fp32 -- when given, the skip add reads / writes it and (y, y32) is returned
        c = w1.shape[0]
        cp = pad64(c)
        assert x.shape[-1] == cp == c, 'RCAB kernels need num_feat % 64 == 0'
        h = raw.tapgemm(x, _packed(w1, 'fprop', cp, cp), ksize=3, cout=cp, bias=_padded_bias(b1, cp), act=L.ACT_RELU)
        # conv2; its epilogue also accumulates the per-sample channel means = AdaptiveAvgPool2d(1) (rcan_arch.py:19)
        t, p = raw.tapgemm(h, _packed(w2, 'fprop', cp, cp), ksize=3, cout=cp, bias=_padded_bias(b2, cp),
                           want_colsum='image_mean')
        # FC + sigmoid + scale + residual in one launch (every CTA recomputes the tiny FC in shared memory)
        y, z, s = raw.ca_forward(t, x if x32 is None else None, p, wa1.detach().contiguous(), ba1.detach(),
                                 wa2.detach().contiguous(), ba2.detach(), res_scale, x32=x32,
                                 want_f32=x32 is not None)
        ctx.save_for_backward(x, h, t, p, z, s, w1, b1, w2, b2, wa1, wa2)
        ctx.res_scale = res_scale
        # the backward's zero-initialised scratch, pre-zeroed by the segment-wide fill when the arch provides one
        raw.stash_backward_scratch(ctx, rcab_backward_floats(x.shape[0], cp), x.device)
        if x32 is None:
            return y
        y, y32 = y
        ctx.mark_non_differentiable(y32)
        # (without this autograd MATERIALISES a zero gradient for the non-differentiable fp32 twin before every backward:
        # a 9.4 MB fill per RCAB, 200 launches and 1.9 GB of stores per step at B16 x 48 x 48)
        ctx.set_materialize_grads(False)
        return y, y32

    @staticmethod
    def backward(ctx, g, *unused):
        if g is None:  # (possible with set_materialize_grads(False): nothing flowed into y)
            return (None,) * 11
        x, h, t, p, z, s, w1, b1, w2, b2, wa1, wa2 = ctx.saved_tensors
        rs = ctx.res_scale
        cp = x.shape[-1]
        g = g.contiguous()
        with raw.backward_arena(ctx, x.device, rcab_backward_floats(x.shape[0], cp)):
            return _RCAB._backward(ctx, g, x, h, t, p, z, s, w1, b1, w2, b2, wa1, wa2, rs, cp)

    @staticmethod
    def _backward(ctx, g, x, h, t, p, z, s, w1, b1, w2, b2, wa1, wa2, rs, cp):
        # d s = res_scale * sum_hw g*t  ->  FC backward  ->  d t (+ its column sums = conv2's bias grad): one launch
        gwa1, gba1, gwa2, gba2, gt, cs2 = raw.ca_backward(g, t, s, z, p, wa1.detach().contiguous(),
                                                         wa2.detach().contiguous(), rs)
        dev = x.device
        # the weight gradients ride a side stream next to the data-gradient chain (raw.side_branch)
        with raw.side_branch(dev):
            acc2 = raw.wgrad(gt, h, ksize=3)
        gh, cs1 = raw.tapgemm(gt, _packed(w2, 'dgrad', cp, cp), ksize=3, cout=cp, flip=True, mask_src=h,
                              mask_mode=L.MASK_SIGN, mask_slope=0.0, want_colsum=True)
        with raw.side_branch(dev):
            acc1 = raw.wgrad(gh, x, ksize=3)
        gx = None
        if ctx.needs_input_grad[0]:
            gx = raw.tapgemm(gh, _packed(w1, 'dgrad', cp, cp), ksize=3, cout=cp, flip=True, residual=g)
        raw.side_join(dev)
        gw2, gb2, gw1, gb1 = raw.finalize_grads([('w', acc2, w2.shape, None, None, 1.0), ('b', cs2, b2.numel(), None, 1.0),
                                                 ('w', acc1, w1.shape, None, None, 1.0), ('b', cs1, b1.numel(), None, 1.0)],
                                                [w2, b2, w1, b1])
        return gx, None, gw1, gb1, gw2, gb2, gwa1, gba1, gwa2, gba2, None


def rcab(x, w1, b1, w2, b2, wa1, ba1, wa2, ba2, res_scale, x32=None):
    """Returns y, or (y, y32) when the fp32 skip stream ``x32`` is carried along."""
    return _RCAB.apply(x, x32, w1, b1, w2, b2, wa1, ba1, wa2, ba2, float(res_scale))


# ------------------------------------------------------------------ stand-alone ChannelAttention (RCAN)
class _ChannelAttention(Function):
    """x * sigmoid(W2 relu(W1 avgpool(x) + b1) + b2)  (rcan_arch.py:16-24) on an NHWC bf16 tensor, from the unfused
    kernels (pool -> FC -> scale); inside an RCAB the fused :class:`_RCAB` is used instead."""

    @staticmethod
    def forward(ctx, t, wa1, ba1, wa2, ba2):
        p = raw.channel_pool(t)
        z, s = raw.ca_fc(p, wa1.detach().contiguous(), ba1.detach(), wa2.detach().contiguous(), ba2.detach())
        y = raw.ca_apply(t, torch.zeros_like(t), s, 1.0)
        ctx.save_for_backward(t, p, z, s, wa1, wa2)
        return y

    @staticmethod
    def backward(ctx, g):
        t, p, z, s, wa1, wa2 = ctx.saved_tensors
        g = g.contiguous()
        gs = raw.channel_dot(g, t)
        gw1, gb1, gw2, gb2, gp = raw.ca_fc_bwd(gs, s, z, p, wa1.detach().contiguous(), wa2.detach().contiguous())
        gt = raw.ca_apply_bwd(g, s, gp, 1.0)
        return gt, gw1, gb1, gw2, gb2


def channel_attention(t, wa1, ba1, wa2, ba2):
    return _ChannelAttention.apply(t, wa1, ba1, wa2, ba2)
