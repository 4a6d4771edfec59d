"""SwinIR ops on the srb200 C-ABI: LayerNorm, fused window attention and the fused
SwinTransformerBlock autograd function (reference swinir_arch.py:283-323).

Layouts (all NHWC bf16, tokens == pixels):
  stream  [B,H,W,Cs]      Cs = pad64(C)              channels >= C are zero
  qkv     [B,H,W,3*Ca]    Ca = num_heads*32          channel = which*Ca + head*32 + d, d >= head_dim zero
  attn o  [B,H,W,Ca]
  mlp     [B,H,W,Ch]      Ch = pad64(hidden)
The padding lives only in the packed bf16 weight copies (index maps below); parameters keep the
reference's shapes, so state dicts are interchangeable.
"""
import os

import torch
from torch.autograd import Function

from ... import _lib as L
from . import raw
from .raw import _chk, _ptr, _stream
from .sr_b200 import _packed, _padded_bias, _perm_cache, pad64

HD_PAD = 32
# v with GELU(v) = v Phi(v) = 1 (any v whose GELU rounds to 1.0 in bf16 would do)
def _gelu_one():
    import math
    lo, hi = 1.0, 1.3
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        lo, hi = (mid, hi) if 0.5 * mid * (1.0 + math.erf(mid / math.sqrt(2.0))) < 1.0 else (lo, mid)
    return 0.5 * (lo + hi)


GELU_ONE = _gelu_one()
LN_EPS = 1e-5


# ------------------------------------------------------------------ raw wrappers
def layernorm_fwd(x, gamma, beta, c, eps=LN_EPS, ones_channel=-1):
    """``ones_channel``: a pad channel of the output set to 1.0 (the consuming Linear's bias gradient then falls out of
    its weight-gradient GEMM, see include/srb200.h)."""
    _chk(x, 'x', torch.bfloat16)
    cp = x.shape[-1]
    t = x.numel() // cp
    y = torch.empty_like(x)
    mean = torch.empty((t,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((t,), dtype=torch.float32, device=x.device)
    raw.probed('layernorm_fwd', (t, c, cp), lambda: L.check(L.load().srb200_layernorm_fwd(
        _ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean), _ptr(rstd), t, c, cp, float(eps), int(ones_channel),
        _stream()),
        'layernorm_fwd'))
    return y, mean, rstd


def layernorm_bwd(gy, x, mean, rstd, gamma, c, gres=None):
    _chk(gy, 'gy', torch.bfloat16)
    cp = x.shape[-1]
    t = x.numel() // cp
    gx = torch.empty_like(x)
    gg = raw.zeros_f32((c,), x.device)
    gb = raw.zeros_f32((c,), x.device)
    raw.probed('layernorm_bwd', (t, c, cp, gres is not None), lambda: L.check(L.load().srb200_layernorm_bwd(
        _ptr(gy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(gres), _ptr(gx), _ptr(gg), _ptr(gb), t, c, cp,
        _stream()), 'layernorm_bwd'))
    return gx, gg, gb


def scale_rows(g, alpha):
    _chk(g, 'g', torch.bfloat16)
    out = torch.empty_like(g)
    b = g.shape[0]
    L.check(L.load().srb200_scale_rows(_ptr(g), _ptr(alpha), _ptr(out), b, g.numel() // b, _stream()), 'scale_rows')
    return out


ATTN_TC_MAX_HEADS = 16  # kMaxHeads of csrc/attention_tc.cuh


def attn_on_tc(qkv, num_heads, ws, shift):
    """Does this shape run on the tcgen05 window-attention kernels (attention_tc.cu)?"""
    return ws == 8 and num_heads % 2 == 0 and num_heads <= ATTN_TC_MAX_HEADS and shift in (0, 4) and \
        qkv.shape[1] % 8 == 0 and qkv.shape[2] % 8 == 0 and qkv.shape[3] == 3 * num_heads * HD_PAD and \
        os.environ.get('SRB_ATTN_MMA_SYNC') is None


def window_attention_fwd(qkv, table, num_heads, ws, shift, scale, want_stats=False, ones=False, alpha=None):
    """Fused window attention.  ``want_stats``: also return the softmax statistics buffer (fp32
    [windows, heads, ws*ws], opaque) that lets :func:`window_attention_bwd` skip the row maxima / sums.
    ``ones``: output channel 31 (a pad lane of head 0, head_dim < 32) carries 1.0 (SRB200_ATTN_ONES).
    ``alpha`` (fp32 [B], tcgen05 shapes only -- see :func:`attn_on_tc`): per-sample scale of the output rows."""
    _chk(qkv, 'qkv', torch.bfloat16)
    _chk(table, 'rpb_table', torch.float32)
    b, h, w, c3 = qkv.shape
    ca = c3 // 3
    out = torch.empty((b, h, w, ca), dtype=torch.bfloat16, device=qkv.device)
    stats = torch.empty((b * (h // ws) * (w // ws), num_heads, ws * ws), dtype=torch.float32, device=qkv.device) \
        if want_stats else None
    raw.probed('window_attn_fwd', (b, h, w, num_heads, ws), lambda: L.check(L.load().srb200_window_attention_fwd(
        _ptr(qkv), _ptr(table), _ptr(out), _ptr(stats), b, h, w, num_heads, ca, ws, shift, float(scale), int(bool(ones)),
        _ptr(alpha), _stream()), 'window_attention_fwd'))
    return (out, stats) if want_stats else out


# Window 8 with the forward's statistics buffer: the tcgen05 backward (attention_tc_bwd.cu, 95 us at B16 x 64x64 against
# 118 us for the mma.sync kernel); SRB_ATTN_BWD_TC=0 forces the mma.sync kernel (tests run both).
ATTN_BWD_TC = os.environ.get('SRB_ATTN_BWD_TC', '1') == '1'


def window_attention_bwd(qkv, gout, table, num_heads, ws, shift, scale, stats=None, use_tc=None):
    _chk(gout, 'gout', torch.bfloat16)
    b, h, w, c3 = qkv.shape
    ca = c3 // 3
    gqkv = torch.empty_like(qkv)
    gtable = raw.zeros_f32(tuple(table.shape), table.device)
    use_tc = ATTN_BWD_TC if use_tc is None else use_tc
    if not use_tc:
        stats = None  # without the statistics buffer the library takes the mma.sync kernel
    raw.probed('window_attn_bwd', (b, h, w, num_heads, ws), lambda: L.check(L.load().srb200_window_attention_bwd(
        _ptr(qkv), _ptr(gout), _ptr(table), _ptr(stats), _ptr(gqkv), _ptr(gtable), None, b, h, w, num_heads, ca,
        ws, shift, float(scale), _stream()), 'window_attention_bwd'))
    return gqkv, gtable


# ------------------------------------------------------------------ index maps (packed -> original, -1 = zero)
def head_perm(num_heads, head_dim, parts, device):
    """packed p = part*Ca + head*32 + d  ->  original part*C + head*head_dim + d (or -1 when d >= head_dim)."""
    key = ('head', num_heads, head_dim, parts, str(device))
    if key not in _perm_cache:
        ca = num_heads * HD_PAD
        p = torch.arange(parts * ca, device=device)
        part, rem = p // ca, p % ca
        head, d = rem // HD_PAD, rem % HD_PAD
        orig = part * (num_heads * head_dim) + head * head_dim + d
        _perm_cache[key] = torch.where(d < head_dim, orig, torch.full_like(orig, -1)).to(torch.int32)
    return _perm_cache[key]


# ------------------------------------------------------------------ LayerNorm
class _LayerNorm(Function):

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        c = gamma.numel()
        y, mean, rstd = layernorm_fwd(x, gamma.detach(), beta.detach(), c, eps)
        ctx.save_for_backward(x, mean, rstd, gamma)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, mean, rstd, gamma = ctx.saved_tensors
        gx, gg, gb = layernorm_bwd(gy.contiguous(), x, mean, rstd, gamma.detach(), gamma.numel())
        return gx, gg, gb, None


def layer_norm(x, gamma, beta, eps=LN_EPS):
    return _LayerNorm.apply(x, gamma, beta, eps)


# ------------------------------------------------------------------ fused SwinTransformerBlock
class _SwinBlock(Function):
    """x + DropPath(proj(WindowAttn(LN1(x)))) then + DropPath(fc2(GELU(fc1(LN2(.)))))  (swinir_arch.py:283-323).

    forward : LN ; tap-GEMM qkv(+bias) ; fused window attention ; tap-GEMM proj(+bias, x alpha[b], + x) ; LN ;
              tap-GEMM fc1(+bias, GELU, pre-activation kept) ; tap-GEMM fc2(+bias, x alpha[b], + x1)
    backward: the mirror image, with d GELU fused into fc2's dgrad epilogue and both skip-gradient adds fused
              into the LayerNorm backward kernels.
    """

    @staticmethod
    def forward(ctx, x, n1w, n1b, qkv_w, qkv_b, table, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b,
                num_heads, ws, shift, alpha1, alpha2, eps=LN_EPS):
        c = n1w.numel()
        cs = x.shape[-1]
        assert cs == pad64(c)
        hd = c // num_heads
        assert hd <= HD_PAD and (num_heads * HD_PAD) % 64 == 0, 'window attention kernel: head_dim <= 32, even heads'
        ca = num_heads * HD_PAD
        hidden = fc1_w.shape[0]
        ch = pad64(hidden)
        dev = x.device
        p_qkv = head_perm(num_heads, hd, 3, dev)
        p_o = head_perm(num_heads, hd, 1, dev)
        scale = hd**-0.5

        # a constant-one pad channel in both LayerNorm outputs: the qkv / fc1 bias gradients then come out of the weight-
        # gradient GEMMs (column `ones`), no column-sum passes in the backward
        ones = cs - 1 if (c < cs and any(ctx.needs_input_grad)) else -1
        xn, mean1, rstd1 = layernorm_fwd(x, n1w.detach(), n1b.detach(), c, eps, ones)
        qkv = raw.tapgemm(xn, _packed(qkv_w, 'fprop', 3 * ca, cs, perm_out=p_qkv), ksize=1, cout=3 * ca,
                          bias=_padded_bias(qkv_b, 3 * ca, p_qkv))
        want_stats = any(ctx.needs_input_grad)
        o_ones = ones >= 0 and hd < HD_PAD  # same trick for proj: its bias gradient = column 31 of its weight gradient
        # on the tcgen05 kernels the attention output carries the branch's per-sample DropPath factor (o' = alpha1[b] * o),
        # proj then scales only its bias, and the backward needs no row-scaled copy of the incoming gradient
        o_alpha = alpha1 is not None and attn_on_tc(qkv, num_heads, ws, shift)
        o = window_attention_fwd(qkv, table.detach(), num_heads, ws, shift, scale, want_stats=want_stats, ones=o_ones,
                                 alpha=alpha1 if o_alpha else None)
        stats = None
        if want_stats:
            o, stats = o
        x1 = raw.tapgemm(o, _packed(proj_w, 'fprop', cs, ca, perm_in=p_o), ksize=1, cout=cs,
                         bias=_padded_bias(proj_b, cs), residual=x, alpha_per_sample=alpha1, alpha_on_bias=o_alpha)
        xn2, mean2, rstd2 = layernorm_fwd(x1, n2w.detach(), n2b.detach(), c, eps, ones)
        # fc1's last PAD output channel gets the bias v with GELU(v) = 1: h carries a constant one there (fc2's packed
        # weights are zero on pad inputs), so fc2's weight-gradient GEMM also yields fc2's bias gradient -- no column-sum pass
        h_ones = hidden < ch and fc1_b is not None
        b1p = _padded_bias(fc1_b, ch, fill=(ch - 1, GELU_ONE) if h_ones else None)
        if want_stats:
            # h carries the MLP branch's per-sample DropPath factor (h' = alpha2[b] * GELU(a); the stored derivative does
            # not): fc2 then only scales its bias, and in the backward dW2 = g2^T h' (and fc2's bias gradient, its pad
            # column) need no row-scaled copy of the incoming gradient
            h, a = raw.tapgemm(xn2, _packed(fc1_w, 'fprop', ch, cs), ksize=1, cout=ch, bias=b1p, act=L.ACT_GELU,
                               want_aux=True, aux_grad=True, alpha_per_sample=alpha2)  # a = gelu'(fc1 out)
        else:  # evaluation: no derivative tensor (50 MB of stores and half of the epilogue math at B16 x 64 x 64)
            h, a = raw.tapgemm(xn2, _packed(fc1_w, 'fprop', ch, cs), ksize=1, cout=ch, bias=b1p, act=L.ACT_GELU,
                               alpha_per_sample=alpha2), None
        x2 = raw.tapgemm(h, _packed(fc2_w, 'fprop', cs, ch), ksize=1, cout=cs, bias=_padded_bias(fc2_b, cs),
                         residual=x1, alpha_per_sample=alpha2, alpha_on_bias=True)
        ctx.save_for_backward(x, mean1, rstd1, xn, qkv, o, x1, mean2, rstd2, xn2, a, h, n1w, qkv_w, qkv_b, table,
                              proj_w, proj_b, n2w, fc1_w, fc1_b, fc2_w, fc2_b, alpha1, alpha2, stats)
        ctx.cfg = (c, cs, ca, ch, hd, num_heads, ws, shift, scale)
        ctx.ones = ones
        ctx.o_ones = o_ones
        ctx.h_ones = h_ones
        ctx.o_alpha = o_alpha
        raw.stash_backward_scratch(ctx, _SwinBlock._scratch_floats(c, cs, ca, ch, table.numel()), dev)
        return x2

    @staticmethod
    def _scratch_floats(c, cs, ca, ch, table_numel):
        return 2 * cs * ch + cs * ca + 3 * ca * cs + 2 * cs + ch + 3 * ca + 4 * c + table_numel + 64 + \
            (ca // HD_PAD) * 4096 + 96  # (+ the attention backward's dS-sum workspace)

    @staticmethod
    def backward(ctx, g2):
        # (read ONCE: under torch.utils.checkpoint(use_reentrant=False) -- SwinIR's use_checkpoint -- a second
        # ctx.saved_tensors raises "already unpacked once")
        saved = ctx.saved_tensors
        (x, mean1, rstd1, xn, qkv, o, x1, mean2, rstd2, xn2, a, h, n1w, qkv_w, qkv_b, table, proj_w, proj_b, n2w,
         fc1_w, fc1_b, fc2_w, fc2_b, alpha1, alpha2, stats) = saved
        c, cs, ca, ch, hd, num_heads, ws, shift, scale = ctx.cfg
        dev = x.device
        p_qkv = head_perm(num_heads, hd, 3, dev)
        p_o = head_perm(num_heads, hd, 1, dev)
        g2 = g2.contiguous()
        arena = raw.backward_arena(ctx, dev, _SwinBlock._scratch_floats(c, cs, ca, ch, table.numel()))
        with arena:
            return _SwinBlock._backward(ctx, g2, arena, saved)

    @staticmethod
    def _backward(ctx, g2, arena, saved):
        (x, mean1, rstd1, xn, qkv, o, x1, mean2, rstd2, xn2, a, h, n1w, qkv_w, qkv_b, table, proj_w, proj_b, n2w,
         fc1_w, fc1_b, fc2_w, fc2_b, alpha1, alpha2, stats) = saved
        c, cs, ca, ch, hd, num_heads, ws, shift, scale = ctx.cfg
        dev = x.device
        p_qkv = head_perm(num_heads, hd, 3, dev)
        p_o = head_perm(num_heads, hd, 1, dev)
        # ---- MLP branch
        # (h already carries alpha2[b], see forward: the weight gradient takes g2 as it is, the data gradient gets the
        # factor in its epilogue -- no scale_rows pass)
        with raw.side_branch(dev):  # the four weight gradients ride a side stream next to the data-gradient chain
            acc_fc2 = raw.wgrad(g2, h, ksize=1)
        if ctx.h_ones:
            fc2_b_item = ('bcol', acc_fc2, ch - 1, fc2_b.numel(), None, 1.0)
        else:
            g2s = scale_rows(g2, alpha2) if alpha2 is not None else g2
            fc2_b_item = ('b', raw.colsum(g2s), fc2_b.numel(), None, 1.0)
        ones = ctx.ones
        if ones >= 0:  # fc1's bias gradient = column `ones` of acc_fc1 (xn2 carries a constant-one channel there)
            ga = raw.tapgemm(g2, _packed(fc2_w, 'dgrad', cs, ch), ksize=1, cout=ch, flip=True, mask_src=a,
                             mask_mode=L.MASK_MUL, alpha_per_sample=alpha2)
            cs_fc1 = None
        else:
            ga, cs_fc1 = raw.tapgemm(g2, _packed(fc2_w, 'dgrad', cs, ch), ksize=1, cout=ch, flip=True, mask_src=a,
                                     mask_mode=L.MASK_MUL, want_colsum=True,
                                     alpha_per_sample=alpha2)  # column sums = fc1's bias gradient
        with raw.side_branch(dev):
            acc_fc1 = raw.wgrad(ga, xn2, ksize=1)
        gxn2 = raw.tapgemm(ga, _packed(fc1_w, 'dgrad', ch, cs), ksize=1, cout=cs, flip=True)
        gx1, g_n2w, g_n2b = layernorm_bwd(gxn2, x1, mean2, rstd2, n2w.detach(), c, gres=g2)
        # ---- attention branch
        if ctx.o_alpha:  # o already carries alpha1[b] (forward): dW = gx1^T o', the data gradient scales in its epilogue
            g1s, a1_epi = gx1, alpha1
        else:
            g1s, a1_epi = (scale_rows(gx1, alpha1) if alpha1 is not None else gx1), None
        with raw.side_branch(dev):
            acc_proj = raw.wgrad(g1s, o, ksize=1)
        if ctx.o_ones:
            proj_b_item = ('bcol', acc_proj, HD_PAD - 1, proj_b.numel(), None, 1.0)
        else:
            gb = scale_rows(gx1, alpha1) if (ctx.o_alpha and alpha1 is not None) else g1s
            proj_b_item = ('b', raw.colsum(gb), proj_b.numel(), None, 1.0)
        go = raw.tapgemm(g1s, _packed(proj_w, 'dgrad', cs, ca, perm_in=p_o), ksize=1, cout=ca, flip=True,
                         alpha_per_sample=a1_epi)
        gqkv, g_table = window_attention_bwd(qkv, go, table.detach(), num_heads, ws, shift, scale, stats=stats)
        with raw.side_branch(dev):
            acc_qkv = raw.wgrad(gqkv, xn, ksize=1)
        gxn = raw.tapgemm(gqkv, _packed(qkv_w, 'dgrad', 3 * ca, cs, perm_out=p_qkv), ksize=1, cout=cs, flip=True)
        gx, g_n1w, g_n1b = layernorm_bwd(gxn, x, mean1, rstd1, n1w.detach(), c, gres=gx1)
        raw.side_join(dev)
        # ---- all eight parameter gradients of the four Linear layers leave through ONE launch
        fc1_b_item = ('bcol', acc_fc1, ones, fc1_b.numel(), None, 1.0) if ones >= 0 else \
            ('b', cs_fc1, fc1_b.numel(), None, 1.0)
        items = [('w', acc_fc2, fc2_w.shape, None, None, 1.0), fc2_b_item,
                 ('w', acc_fc1, fc1_w.shape, None, None, 1.0), fc1_b_item,
                 ('w', acc_proj, proj_w.shape, None, p_o, 1.0), proj_b_item,
                 ('w', acc_qkv, qkv_w.shape, p_qkv, None, 1.0)]
        prms = [fc2_w, fc2_b, fc1_w, fc1_b, proj_w, proj_b, qkv_w]
        if qkv_b is not None:
            items.append(('bcol', acc_qkv, ones, qkv_b.numel(), p_qkv, 1.0) if ones >= 0 else
                         ('b', raw.colsum(gqkv), qkv_b.numel(), p_qkv, 1.0))
            prms.append(qkv_b)
        grads = raw.finalize_grads(items, prms)
        g_fc2_w, g_fc2_b, g_fc1_w, g_fc1_b, g_proj_w, g_proj_b, g_qkv_w = grads[:7]
        g_qkv_b = grads[7] if qkv_b is not None else None
        return (gx, g_n1w, g_n1b, g_qkv_w, g_qkv_b, g_table, g_proj_w, g_proj_b, g_n2w, g_n2b, g_fc1_w, g_fc1_b,
                g_fc2_w, g_fc2_b, None, None, None, None, None, None)


def swin_block(x, n1w, n1b, qkv_w, qkv_b, table, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b, num_heads, ws,
               shift, alpha1=None, alpha2=None, eps=LN_EPS):
    return _SwinBlock.apply(x, n1w, n1b, qkv_w, qkv_b, table, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b,
                            num_heads, ws, shift, alpha1, alpha2, float(eps))


# ------------------------------------------------------------------ stand-alone Mlp / WindowAttention
class _Mlp(Function):
    """fc2(GELU(fc1(x)))  (swinir_arch.py:54-60) on an NHWC64 bf16 tensor: two tap-GEMMs, GELU and its derivative in
    fc1's epilogue -- the MLP branch of :class:`_SwinBlock` without the LayerNorm / skip around it."""

    @staticmethod
    def forward(ctx, x, fc1_w, fc1_b, fc2_w, fc2_b):
        cs = x.shape[-1]
        hidden, cout = fc1_w.shape[0], fc2_w.shape[0]
        ch, co = pad64(hidden), pad64(cout)
        assert cs == pad64(fc1_w.shape[1])
        h, a = raw.tapgemm(x, _packed(fc1_w, 'fprop', ch, cs), ksize=1, cout=ch, bias=_padded_bias(fc1_b, ch),
                           act=L.ACT_GELU, want_aux=True, aux_grad=True)
        y = raw.tapgemm(h, _packed(fc2_w, 'fprop', co, ch), ksize=1, cout=co, bias=_padded_bias(fc2_b, co))
        ctx.save_for_backward(x, a, h, fc1_w, fc1_b, fc2_w, fc2_b)
        raw.stash_backward_scratch(ctx, 2 * co * ch + 2 * cs * ch + 2 * (co + ch + cs) + 64, x.device)
        return y

    @staticmethod
    def backward(ctx, g):
        x, a, h, fc1_w, fc1_b, fc2_w, fc2_b = ctx.saved_tensors
        cs, ch, co = x.shape[-1], h.shape[-1], g.shape[-1]
        g = g.contiguous()
        with raw.backward_arena(ctx, x.device, 2 * co * ch + 2 * cs * ch + 2 * (co + ch + cs) + 64):
            acc2 = raw.wgrad(g, h, ksize=1)
            cs2 = raw.colsum(g)
            ga, cs1 = raw.tapgemm(g, _packed(fc2_w, 'dgrad', co, ch), ksize=1, cout=ch, flip=True, mask_src=a,
                                  mask_mode=L.MASK_MUL, want_colsum=True)
            acc1 = raw.wgrad(ga, x, ksize=1)
            gx = raw.tapgemm(ga, _packed(fc1_w, 'dgrad', ch, cs), ksize=1, cout=cs, flip=True) \
                if ctx.needs_input_grad[0] else None
            items = [('w', acc1, fc1_w.shape, None, None, 1.0), ('w', acc2, fc2_w.shape, None, None, 1.0)]
            if fc1_b is not None:
                items.append(('b', cs1, fc1_b.numel(), None, 1.0))
            if fc2_b is not None:
                items.append(('b', cs2, fc2_b.numel(), None, 1.0))
            grads = raw.finalize_grads(items)
        g1w, g2w = grads[0], grads[1]
        rest = iter(grads[2:])
        g1b = next(rest) if fc1_b is not None else None
        g2b = next(rest) if fc2_b is not None else None
        return gx, g1w, g1b, g2w, g2b


def mlp(x, fc1_w, fc1_b, fc2_w, fc2_b):
    return _Mlp.apply(x, fc1_w, fc1_b, fc2_w, fc2_b)


class _WindowAttn(Function):
    """proj(softmax(scale q k^T + rpb [+ analytic SW-MSA mask]) v) on an NHWC64 bf16 tensor (swinir_arch.py:144-175):
    qkv tap-GEMM, the fused window-attention kernel (partition / shift / reverse are its addressing), proj tap-GEMM --
    the attention branch of :class:`_SwinBlock` without the LayerNorm / skip around it."""

    @staticmethod
    def forward(ctx, x, qkv_w, qkv_b, table, proj_w, proj_b, num_heads, ws, shift):
        c = qkv_w.shape[1]
        cs = x.shape[-1]
        assert cs == pad64(c)
        hd = c // num_heads
        assert hd <= HD_PAD and (num_heads * HD_PAD) % 64 == 0, 'window attention kernel: head_dim <= 32, even heads'
        ca = num_heads * HD_PAD
        p_qkv = head_perm(num_heads, hd, 3, x.device)
        p_o = head_perm(num_heads, hd, 1, x.device)
        scale = hd**-0.5
        qkv = raw.tapgemm(x, _packed(qkv_w, 'fprop', 3 * ca, cs, perm_out=p_qkv), ksize=1, cout=3 * ca,
                          bias=_padded_bias(qkv_b, 3 * ca, p_qkv))
        o, stats = window_attention_fwd(qkv, table.detach(), num_heads, ws, shift, scale, want_stats=True)
        y = raw.tapgemm(o, _packed(proj_w, 'fprop', cs, ca, perm_in=p_o), ksize=1, cout=cs,
                        bias=_padded_bias(proj_b, cs))
        ctx.save_for_backward(x, qkv, o, qkv_w, qkv_b, table, proj_w, proj_b, stats)
        ctx.cfg = (c, cs, ca, hd, num_heads, ws, shift, scale)
        raw.stash_backward_scratch(ctx, cs * ca + 3 * ca * cs + 2 * cs + 3 * ca + table.numel() + 64, x.device)
        return y

    @staticmethod
    def backward(ctx, g):
        x, qkv, o, qkv_w, qkv_b, table, proj_w, proj_b, stats = ctx.saved_tensors
        c, cs, ca, hd, num_heads, ws, shift, scale = ctx.cfg
        p_qkv = head_perm(num_heads, hd, 3, x.device)
        p_o = head_perm(num_heads, hd, 1, x.device)
        g = g.contiguous()
        with raw.backward_arena(ctx, x.device, cs * ca + 3 * ca * cs + 2 * cs + 3 * ca + table.numel() + 64):
            acc_proj = raw.wgrad(g, o, ksize=1)
            cs_proj = raw.colsum(g)
            go = raw.tapgemm(g, _packed(proj_w, 'dgrad', cs, ca, perm_in=p_o), ksize=1, cout=ca, flip=True)
            gqkv, g_table = window_attention_bwd(qkv, go, table.detach(), num_heads, ws, shift, scale, stats=stats)
            acc_qkv = raw.wgrad(gqkv, x, ksize=1)
            gx = raw.tapgemm(gqkv, _packed(qkv_w, 'dgrad', 3 * ca, cs, perm_out=p_qkv), ksize=1, cout=cs, flip=True) \
                if ctx.needs_input_grad[0] else None
            items = [('w', acc_proj, proj_w.shape, None, p_o, 1.0), ('w', acc_qkv, qkv_w.shape, p_qkv, None, 1.0)]
            if proj_b is not None:
                items.append(('b', cs_proj, proj_b.numel(), None, 1.0))
            if qkv_b is not None:
                items.append(('b', raw.colsum(gqkv), qkv_b.numel(), p_qkv, 1.0))
            grads = raw.finalize_grads(items)
        g_proj_w, g_qkv_w = grads[0], grads[1]
        rest = iter(grads[2:])
        g_proj_b = next(rest) if proj_b is not None else None
        g_qkv_b = next(rest) if qkv_b is not None else None
        return gx, g_qkv_w, g_qkv_b, g_table.clone(), g_proj_w, g_proj_b, None, None, None


def window_attn(x, qkv_w, qkv_b, table, proj_w, proj_b, num_heads, ws, shift):
    return _WindowAttn.apply(x, qkv_w, qkv_b, table, proj_w, proj_b, num_heads, ws, shift)


# ------------------------------------------------------------------ token <-> NHWC64 plumbing (module boundaries only)
def tokens_to_nhwc(x, h, w):
    """[B, h*w, C] (any float dtype) -> [B, h, w, pad64(C)] bf16, differentiable.  Only the stand-alone ``forward`` of a
    sub-module pays this; inside a SwinIR network the stream never leaves the NHWC64 layout."""
    b, _, c = x.shape
    cp = pad64(c)
    t = x.reshape(b, h, w, c)
    if cp != c:
        t = torch.nn.functional.pad(t, (0, cp - c))
    return t.to(torch.bfloat16).contiguous()


def nhwc_to_tokens(t, c, dtype):
    b, h, w, _ = t.shape
    return t[..., :c].reshape(b, h * w, c).to(dtype)
