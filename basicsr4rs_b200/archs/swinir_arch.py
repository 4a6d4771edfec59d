"""SwinIR on the srb200 kernels -- drop-in for the reference's ``basicsr/archs/swinir_arch.py``.

Module tree, constructor / YAML keys, parameter and buffer names, shapes and init order follow the
reference (WindowAttention :95-191, SwinTransformerBlock :194-341, BasicLayer :393-477, RSTB :480-568,
PatchEmbed / PatchUnEmbed :571-644, Upsample(OneStep) :647-690, SwinIR :693-933) so that state dicts
(550 entries for the classical x4 model, incl. the ``relative_position_index`` / ``attn_mask`` buffers) and
seeded random inits are interchangeable.  What differs is how ``forward`` computes:

* tokens [B, H*W, C] *are* NHWC pixels, so PatchEmbed / PatchUnEmbed / window_partition / window_reverse /
  torch.roll never materialise -- they are address arithmetic inside the fused window-attention kernel;
* one autograd function per SwinTransformerBlock (ops/sr_b200/swin_ops.py): LayerNorm kernels, tcgen05
  tap-GEMMs for qkv / proj / fc1 / fc2 with bias, GELU, DropPath scale and skip add in their epilogues;
* the shifted-window mask is analytic (any H, W multiple of the window), the stored ``attn_mask`` buffer
  is kept only for state-dict compatibility -- nothing is rebuilt on the CPU per call (cf. :305-306);
* 3x3 convs, pixel-shuffle upsampler and image entry / exit as in EDSR.
"""

import os

import torch
import torch.nn as nn

from ..ops import sr_b200 as ops
from ..ops.sr_b200 import swin_ops
from ..utils.registry import ARCH_REGISTRY
from .arch_util import Upsample, require_cuda, to_2tuple, trunc_normal_
from .graphed import ArchMixin, Segment, chain_wire, graphed_forward, split_even

KERNEL_WINDOW = 8  # the fused attention kernel holds a window in 64-row tiles: window sizes 2..8


def drop_path_scale(batch, drop_prob, training, device):
    """Per-sample stochastic-depth factor floor(keep + U[0,1)) / keep as fp32 [B] (reference :14-26), or None."""
    if drop_prob == 0. or not training:
        return None
    keep = 1 - drop_prob
    return ((keep + torch.rand((batch,), dtype=torch.float32, device=device)).floor_() / keep).contiguous()


_drop_path_scale_impl = drop_path_scale
_KEEP_CACHE = {}


def drop_path_bank(batch, probs, training, device):
    """The stochastic-depth factors of a whole group of DropPath uses from ONE draw: a list with one fp32 [B] tensor (or
    None where ``p == 0``) per entry of ``probs``.  Drawn one by one (rand, add, floor, divide per use) they are four
    1.4 us launches per use -- 288 per SwinIR step; a bank of rows costs the same four launches per RSTB layer."""
    if drop_path_scale is not _drop_path_scale_impl:  # replaced (the parity tests inject masks): one call per use, in order
        return [drop_path_scale(batch, p, training, device) for p in probs]
    if not training or not any(p > 0. for p in probs):
        return [None] * len(probs)
    key = (tuple(probs), str(device))
    keep = _KEEP_CACHE.get(key)
    if keep is None:
        keep = _KEEP_CACHE[key] = torch.tensor([1. - p for p in probs], dtype=torch.float32, device=device).view(-1, 1)
    bank = (keep + torch.rand((len(probs), batch), dtype=torch.float32, device=device)).floor_() / keep
    return [bank[i] if p > 0. else None for i, p in enumerate(probs)]


class DropPath(nn.Module):
    """Stochastic depth per sample; in the fused block it becomes a per-sample epilogue scale."""

    def __init__(self, drop_prob=None):
        super(DropPath, self).__init__()
        self.drop_prob = drop_prob

    def scale(self, batch, device):
        return drop_path_scale(batch, self.drop_prob or 0., self.training, device)

    def forward(self, x):
        """Stand-alone use (reference :14-26, :39-40): x * floor(keep + U[0,1)) / keep per sample."""
        a = self.scale(x.shape[0], x.device)
        return x if a is None else x * a.to(x.dtype).view((-1,) + (1,) * (x.dim() - 1))


class Mlp(nn.Module):

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def _on_kernels(self):
        return isinstance(self.act, nn.GELU) and getattr(self.act, 'approximate', 'none') == 'none' and \
            not (self.training and self.drop.p > 0.)

    def forward_nhwc(self, t):
        """[..., pad64(C)] bf16 -> [..., pad64(out)] bf16: two tap-GEMMs with GELU in fc1's epilogue."""
        if self._on_kernels():
            return swin_ops.mlp(t, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)
        # dropout > 0 / another activation: the reference expression (:54-60) on the un-padded channels
        c, co = self.fc1.in_features, self.fc2.out_features
        y = self.drop(self.fc2(self.drop(self.act(self.fc1(t[..., :c].float())))))
        return nn.functional.pad(y, (0, ops.pad64(co) - co)).to(torch.bfloat16).contiguous()

    def forward(self, x):
        """Reference contract (:54-60): x [..., in_features] float -> [..., out_features]."""
        require_cuda(x, 'Mlp')
        lead = x.shape[:-1]
        t = swin_ops.tokens_to_nhwc(x.reshape(1, -1, x.shape[-1]), 1, x.numel() // x.shape[-1])
        y = self.forward_nhwc(t)
        return swin_ops.nhwc_to_tokens(y, self.fc2.out_features, x.dtype).reshape(lead + (self.fc2.out_features,))


def window_partition(x, window_size):
    """(b, h, w, c) -> (num_windows*b, ws, ws, c); bit-exact remap kernel (reference :63-75)."""
    return ops.raw.window_partition(x.contiguous(), window_size, 0)


def window_reverse(windows, window_size, h, w):
    """(num_windows*b, ws, ws, c) -> (b, h, w, c) (reference :78-92)."""
    return ops.raw.window_reverse(windows.contiguous(), window_size, h, w, 0)


def _relative_position_index(ws):
    ch, cw = torch.arange(ws[0]), torch.arange(ws[1])
    coords = torch.stack(torch.meshgrid([ch, cw], indexing='ij')).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws[0] - 1
    rel[:, :, 1] += ws[1] - 1
    rel[:, :, 0] *= 2 * ws[1] - 1
    return rel.sum(-1)


class WindowAttention(nn.Module):
    """Parameter holder of W-MSA / SW-MSA (relative_position_bias_table, qkv, proj)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim = dim
        self.window_size = window_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim**-0.5

        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1), num_heads))
        self.register_buffer('relative_position_index', _relative_position_index(window_size))

        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

        trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)
        self._qk_scale_override = qk_scale

    def on_kernels(self):
        """True when the fused kernel covers this configuration; otherwise callers run the reference expression."""
        hd = self.dim // self.num_heads
        return (self.window_size[0] == self.window_size[1] and 2 <= self.window_size[0] <= KERNEL_WINDOW and
                hd <= swin_ops.HD_PAD and self.num_heads % 2 == 0 and self.dim % self.num_heads == 0 and
                self._qk_scale_override is None and
                not (self.training and (self.attn_drop.p > 0. or self.proj_drop.p > 0.)))

    def forward_nhwc(self, t, shift):
        """t [B,H,W,pad64(C)] bf16 un-partitioned; partition / shift / mask / reverse happen inside the kernel."""
        return swin_ops.window_attn(t, self.qkv.weight, self.qkv.bias, self.relative_position_bias_table,
                                    self.proj.weight, self.proj.bias, self.num_heads, self.window_size[0], shift)

    def torch_forward(self, x, mask=None):
        """The reference expression (:144-175) on stock torch ops -- the fall-through for configurations the fused
        kernel does not cover (window > 8, odd head counts, head_dim > 32, qk_scale, dropout > 0, explicit masks)."""
        b_, n, c = x.shape
        qkv = self.qkv(x).reshape(b_, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * self.scale, qkv[1], qkv[2]
        attn = q @ k.transpose(-2, -1)
        nt = self.window_size[0] * self.window_size[1]
        bias = self.relative_position_bias_table[self.relative_position_index.view(-1)].view(nt, nt, -1)
        attn = attn + bias.permute(2, 0, 1).contiguous().unsqueeze(0)
        if mask is not None:
            nw = mask.shape[0]
            attn = attn.view(b_ // nw, nw, self.num_heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
            attn = attn.view(-1, self.num_heads, n, n)
        attn = self.attn_drop(self.softmax(attn))
        x = (attn @ v).transpose(1, 2).reshape(b_, n, c)
        return self.proj_drop(self.proj(x))

    def forward(self, x, mask=None):
        """Reference contract (:144-175): x (num_windows*b, n, c) already partitioned; mask (nW, n, n) or None.
        Without an explicit mask the windows run through the fused kernel as ws x ws images (one window each)."""
        require_cuda(x, 'WindowAttention')
        ws = self.window_size[0]
        if mask is None and self.on_kernels() and x.shape[1] == ws * ws:
            y = self.forward_nhwc(swin_ops.tokens_to_nhwc(x, ws, ws), 0)
            return swin_ops.nhwc_to_tokens(y, self.dim, x.dtype)
        return self.torch_forward(x, mask)

    def extra_repr(self) -> str:
        return f'dim={self.dim}, window_size={self.window_size}, num_heads={self.num_heads}'

    def flops(self, n):
        """MACs of one window with n tokens (reference :180-191)."""
        hd = self.dim // self.num_heads
        return n * self.dim * 3 * self.dim + 2 * self.num_heads * n * hd * n + n * self.dim * self.dim


class SwinTransformerBlock(nn.Module):

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        if min(self.input_resolution) <= self.window_size:
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, 'shift_size must in 0-window_size'

        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=to_2tuple(self.window_size), num_heads=num_heads,
                                    qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

        self.register_buffer('attn_mask', self.calculate_mask(self.input_resolution) if self.shift_size > 0 else None)

    def calculate_mask(self, x_size):
        """0 / -100 SW-MSA mask; state-dict compatibility only (the kernel derives it analytically)."""
        h, w = x_size
        ws, s = self.window_size, self.shift_size
        img = torch.zeros((1, h, w, 1))
        cnt = 0
        for hs in (slice(0, -ws), slice(-ws, -s), slice(-s, None)):
            for wsl in (slice(0, -ws), slice(-ws, -s), slice(-s, None)):
                img[:, hs, wsl, :] = cnt
                cnt += 1
        mw = img.view(1, h // ws, ws, w // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
        m = mw.unsqueeze(1) - mw.unsqueeze(2)
        return m.masked_fill(m != 0, float(-100.0)).masked_fill(m == 0, float(0.0))

    def on_kernels(self):
        """The fused block function needs the fused attention kernel, nn.LayerNorm, an exact-erf GELU Mlp and no
        dropout; anything else takes :meth:`_forward_nhwc_general` (reference expression for the uncovered piece)."""
        return (self.attn.on_kernels() and self.mlp._on_kernels() and type(self.norm1) is nn.LayerNorm and
                type(self.norm2) is nn.LayerNorm and self.norm1.eps == self.norm2.eps and
                self.norm1.elementwise_affine and self.norm2.elementwise_affine)

    def drop_probs(self):
        """The two stochastic-depth probabilities of this block (attention branch, MLP branch)."""
        p = float(self.drop_path.drop_prob or 0.) if isinstance(self.drop_path, DropPath) else 0.
        return [p, p]

    def forward_nhwc(self, t, alphas=None):
        """t: [B, H, W, pad64(C)] bf16.  ``alphas``: this block's two per-sample DropPath factors when the caller drew
        them for a whole layer at once (:func:`drop_path_bank`)."""
        if not self.on_kernels():
            return self._forward_nhwc_general(t)
        b = t.shape[0]
        dp = self.drop_path
        if alphas is not None:
            a1, a2 = alphas
        else:
            a1 = dp.scale(b, t.device) if isinstance(dp, DropPath) else None
            a2 = dp.scale(b, t.device) if isinstance(dp, DropPath) else None
        at, m = self.attn, self.mlp
        return swin_ops.swin_block(t, self.norm1.weight, self.norm1.bias, at.qkv.weight, at.qkv.bias,
                                   at.relative_position_bias_table, at.proj.weight, at.proj.bias, self.norm2.weight,
                                   self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias,
                                   self.num_heads, self.window_size, self.shift_size, a1, a2, eps=self.norm1.eps)

    def forward_nhwc_folded(self, t, stats, fold_qkv, fold_fc1):
        """Evaluation forward with both LayerNorms folded into the Linears that consume them (no LayerNorm launch, no
        normalised copy of the tokens): proj / fc2 also emit (mean, rstd) of the rows they store
        (``srb200_tapgemm_ext.ln_stats_out``), qkv / fc1 run on the raw rows with gamma-scaled weights and undo the
        mean in their epilogue (``ln_stats_in``).  ``stats`` = the statistics of ``t`` from the previous block's fc2,
        or None (first block of a layer: LayerNorm kernel + plain qkv).  ``fold_*`` = (packed W diag(gamma), row sums
        of it, W beta + b) from :meth:`BasicLayer._fold_bank`.  Returns (x2, stats of x2)."""
        from .. import _lib as L
        raw = ops.raw
        at, m = self.attn, self.mlp
        c, nh = self.dim, self.num_heads
        hd = c // nh
        cs = t.shape[-1]
        ca = nh * swin_ops.HD_PAD
        ch = ops.pad64(m.fc1.weight.shape[0])
        eps = self.norm1.eps
        p_o = swin_ops.head_perm(nh, hd, 1, t.device)
        if stats is None:
            p_qkv = swin_ops.head_perm(nh, hd, 3, t.device)
            xn = swin_ops.layernorm_fwd(t, self.norm1.weight.detach(), self.norm1.bias.detach(), c, eps)[0]
            qkv = raw.tapgemm(xn, swin_ops._packed(at.qkv.weight, 'fprop', 3 * ca, cs, perm_out=p_qkv), ksize=1, cout=3 * ca,
                              bias=swin_ops._padded_bias(at.qkv.bias, 3 * ca, p_qkv))
        else:
            qkv = raw.tapgemm(t, fold_qkv[0], ksize=1, cout=3 * ca, bias=fold_qkv[2], ln_in=(stats, fold_qkv[1]))
        o = swin_ops.window_attention_fwd(qkv, at.relative_position_bias_table.detach(), nh, self.window_size,
                                          self.shift_size, at.scale)
        x1, st1 = raw.tapgemm(o, swin_ops._packed(at.proj.weight, 'fprop', cs, ca, perm_in=p_o), ksize=1, cout=cs,
                              bias=swin_ops._padded_bias(at.proj.bias, cs), residual=t, ln_out=(c, eps))
        h = raw.tapgemm(x1, fold_fc1[0], ksize=1, cout=ch, bias=fold_fc1[2], act=L.ACT_GELU, ln_in=(st1, fold_fc1[1]))
        return raw.tapgemm(h, swin_ops._packed(m.fc2.weight, 'fprop', cs, ch), ksize=1, cout=cs,
                           bias=swin_ops._padded_bias(m.fc2.bias, cs), residual=x1, ln_out=(c, eps))

    def _forward_nhwc_general(self, t):
        """Fall-through for configurations without a fused kernel (they must keep running, not
        raise): window_size > 8, odd head counts, head_dim > 32, qk_scale, dropout > 0, a custom norm_layer / act_layer.
        The uncovered piece is the reference expression (:283-323) on stock torch ops over the GPU tensors; everything
        the kernels do cover (LayerNorm, Mlp, the attention of a supported shape) still goes through them."""
        b, h, w, cp = t.shape
        c, ws, sh = self.dim, self.window_size, self.shift_size

        def norm(mod, z):
            if type(mod) is nn.LayerNorm and mod.elementwise_affine:
                return swin_ops.layer_norm(z, mod.weight, mod.bias, mod.eps)
            y = mod(z[..., :c].float())
            return nn.functional.pad(y, (0, cp - c)).to(torch.bfloat16).contiguous()

        def residual(z, y):  # z + DropPath(y), both [B,H,W,cp] bf16 (padding channels stay zero)
            y = self.drop_path(y) if isinstance(self.drop_path, DropPath) else y
            return (z.float() + y.float()).to(torch.bfloat16)

        xn = norm(self.norm1, t)
        if self.attn.on_kernels():
            y = self.attn.forward_nhwc(xn, sh)
        else:
            xs = xn[..., :c].float()
            if sh > 0:
                xs = torch.roll(xs, shifts=(-sh, -sh), dims=(1, 2))
            win = xs.view(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, c)
            mask = None
            if sh > 0:
                mask = self.attn_mask if tuple(self.input_resolution) == (h, w) else self.calculate_mask((h, w))
                mask = mask.to(t.device)
            aw = self.attn.torch_forward(win, mask)
            ys = aw.view(b, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c)
            if sh > 0:
                ys = torch.roll(ys, shifts=(sh, sh), dims=(1, 2))
            y = nn.functional.pad(ys, (0, cp - c)).to(torch.bfloat16)
        x1 = residual(t, y)
        return residual(x1, self.mlp.forward_nhwc(norm(self.norm2, x1)))

    def forward(self, x, x_size):
        """Reference contract (:283-323): x [B, h*w, C] tokens, x_size = (h, w)."""
        require_cuda(x, 'SwinTransformerBlock')
        h, w = x_size
        y = self.forward_nhwc(swin_ops.tokens_to_nhwc(x, h, w))
        return swin_ops.nhwc_to_tokens(y, self.dim, x.dtype)

    def flops(self):
        """Reference :329-341."""
        h, w = self.input_resolution
        nw = h * w / self.window_size / self.window_size
        return (2 * self.dim * h * w + nw * self.attn.flops(self.window_size * self.window_size) +
                2 * h * w * self.dim * self.dim * self.mlp_ratio)

    def extra_repr(self) -> str:
        return (f'dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads}, '
                f'window_size={self.window_size}, shift_size={self.shift_size}, mlp_ratio={self.mlp_ratio}')


class BasicLayer(nn.Module):

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, downsample=None,
                 use_checkpoint=False):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads,
                                 window_size=window_size, shift_size=0 if (i % 2 == 0) else window_size // 2,
                                 mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop,
                                 attn_drop=attn_drop,
                                 drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                 norm_layer=norm_layer) for i in range(depth)
        ])
        self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer) if downsample is not None else None

    # evaluation on big tiles (tiled inference): LayerNorm folded into the Linears; below this many tokens the per-forward
    # folding of the weights (a dozen small launches per layer) is not worth it
    FOLD_MIN_TOKENS = int(os.environ.get('SRB_SWIN_FOLD_MIN_TOKENS', '200000'))

    def _fold_ok(self, t):
        if torch.is_grad_enabled() or self.training or t.shape[0] * t.shape[1] * t.shape[2] < self.FOLD_MIN_TOKENS:
            return False
        b0 = self.blocks[0]

        def n_tile(cout):  # srb200_tapgemm's N tile: the folding epilogues exist for the 128- / 192-wide two-team kernels
            return 256 if cout % 256 == 0 else 192 if cout % 192 == 0 else 128 if cout % 128 == 0 else 64

        widths = (3 * b0.num_heads * swin_ops.HD_PAD, ops.pad64(b0.mlp.fc1.weight.shape[0]))
        return (t.shape[-1] in (128, 192) and all(n_tile(n) in (128, 192) for n in widths) and all(blk.on_kernels() for blk in self.blocks) and
                all(blk.norm1.eps == b0.norm1.eps and blk.num_heads == b0.num_heads for blk in self.blocks) and
                all(blk.mlp.fc1.bias is not None and blk.attn.proj.bias is not None and blk.mlp.fc2.bias is not None
                    for blk in self.blocks))

    def _fold_bank(self, cs, dev):
        """Per block: (bf16 packed W diag(gamma), fp32 row sums of that packed operand, fp32 W beta + b) for qkv (norm1
        folded) and fc1 (norm2 folded), derived afresh per forward (evaluation nets are EMA copies updated through
        ``.data``), batched over the layer's blocks: two stacks, two products, one pack launch per weight."""
        raw = ops.raw
        blocks = list(self.blocks)
        nh = blocks[0].num_heads
        hd = self.dim // nh

        def bank(lins, norms, n_pad, perm):
            w = torch.stack([lin.weight.detach().float() for lin in lins])            # [n, Co, C]
            g = torch.stack([nm.weight.detach().float() for nm in norms])             # [n, C]
            bt = torch.stack([nm.bias.detach().float() for nm in norms])
            wg = (w * g[:, None, :]).contiguous()
            packed = torch.empty((len(lins), 1, n_pad, cs), dtype=torch.bfloat16, device=dev)
            for i in range(len(lins)):
                raw.pack_weight(wg[i], n_pad, cs, perm_out=perm, out=packed[i])
            wsum = packed.float().sum(dim=3).reshape(len(lins), n_pad).contiguous()
            b = torch.bmm(w, bt[:, :, None])[:, :, 0]                                 # W beta
            if lins[0].bias is not None:
                b = b + torch.stack([lin.bias.detach().float() for lin in lins])
            if perm is None:
                bp = nn.functional.pad(b, (0, n_pad - b.shape[1]))
            else:
                bp = torch.where(perm >= 0, b[:, perm.clamp(min=0).long()], torch.zeros((), device=dev))
            bp = bp.contiguous()
            return [(packed[i], wsum[i], bp[i]) for i in range(len(lins))]

        ca = nh * swin_ops.HD_PAD
        ch = ops.pad64(blocks[0].mlp.fc1.weight.shape[0])
        qkv = bank([blk.attn.qkv for blk in blocks], [blk.norm1 for blk in blocks], 3 * ca,
                   swin_ops.head_perm(nh, hd, 3, dev))
        fc1 = bank([blk.mlp.fc1 for blk in blocks], [blk.norm2 for blk in blocks], ch, None)
        return qkv, fc1

    def forward_nhwc(self, t, stats=None):
        """``stats`` (folded evaluation only): (mean, rstd) of the rows of ``t`` when its producer -- the previous RSTB's
        conv -- emitted them; the first block then needs no LayerNorm launch either."""
        if self._fold_ok(t):
            fq, f1 = self._fold_bank(t.shape[-1], t.device)
            for i, blk in enumerate(self.blocks):
                t, stats = blk.forward_nhwc_folded(t, stats, fq[i], f1[i])
            return t
        # all DropPath factors of the layer from one draw (blocks without a fused kernel draw their own)
        bank = None
        if self.training and all(blk.on_kernels() for blk in self.blocks):
            probs = [p for blk in self.blocks for p in blk.drop_probs()]
            bank = drop_path_bank(t.shape[0], probs, True, t.device)
        for i, blk in enumerate(self.blocks):
            al = (bank[2 * i], bank[2 * i + 1]) if bank is not None else None
            if self.use_checkpoint and torch.is_grad_enabled():
                t = torch.utils.checkpoint.checkpoint(blk.forward_nhwc, t, al, use_reentrant=False)
            else:
                t = blk.forward_nhwc(t, al)
        return t

    def forward(self, x, x_size):
        """Reference contract (:458-466): x [B, h*w, C] tokens."""
        require_cuda(x, 'BasicLayer')
        y = self.forward_nhwc(swin_ops.tokens_to_nhwc(x, *x_size))
        if self.downsample is not None:
            return self.downsample(swin_ops.nhwc_to_tokens(y, self.dim, x.dtype))
        return swin_ops.nhwc_to_tokens(y, self.dim, x.dtype)

    def extra_repr(self) -> str:
        return f'dim={self.dim}, input_resolution={self.input_resolution}, depth={self.depth}'

    def flops(self):
        flops = sum(blk.flops() for blk in self.blocks)
        if self.downsample is not None:
            flops += self.downsample.flops()
        return flops


class PatchEmbed(nn.Module):
    """[B,C,H,W] -> [B,H*W,C] (+ LayerNorm).  In NHWC the reshape is the identity; only the norm computes."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.patches_resolution = [img_size[0] // patch_size[0], img_size[1] // patch_size[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward_nhwc(self, t):
        return swin_ops.layer_norm(t, self.norm.weight, self.norm.bias, self.norm.eps) if self.norm is not None else t

    def forward(self, x):
        """Reference contract (:600-604): [B,C,H,W] -> [B,H*W,C] (+ LayerNorm)."""
        b, c, h, w = x.shape
        x = x.flatten(2).transpose(1, 2)
        if self.norm is None:
            return x
        require_cuda(x, 'PatchEmbed')
        return swin_ops.nhwc_to_tokens(self.forward_nhwc(swin_ops.tokens_to_nhwc(x, h, w)), c, x.dtype)

    def flops(self):
        h, w = self.img_size
        return h * w * self.embed_dim if self.norm is not None else 0


class PatchUnEmbed(nn.Module):
    """[B,H*W,C] -> [B,C,H,W]; the identity in NHWC."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.patches_resolution = [img_size[0] // patch_size[0], img_size[1] // patch_size[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans = in_chans
        self.embed_dim = embed_dim

    def forward(self, x, x_size):
        """Reference contract (:638-640): [B,H*W,C] -> [B,C,H,W]."""
        return x.transpose(1, 2).contiguous().view(x.shape[0], self.embed_dim, x_size[0], x_size[1])

    def flops(self):
        return 0


class RSTB(nn.Module):
    """Residual Swin Transformer Block: BasicLayer -> conv3x3 -> + x; the skip add rides the conv epilogue."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, downsample=None,
                 use_checkpoint=False, img_size=224, patch_size=4, resi_connection='1conv'):
        super(RSTB, self).__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.residual_group = BasicLayer(dim=dim, input_resolution=input_resolution, depth=depth,
                                         num_heads=num_heads, window_size=window_size, mlp_ratio=mlp_ratio,
                                         qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                                         drop_path=drop_path, norm_layer=norm_layer, downsample=downsample,
                                         use_checkpoint=use_checkpoint)
        if resi_connection == '1conv':
            self.conv = nn.Conv2d(dim, dim, 3, 1, 1)
        elif resi_connection == '3conv':
            self.conv = nn.Sequential(
                nn.Conv2d(dim, dim // 4, 3, 1, 1), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                nn.Conv2d(dim // 4, dim // 4, 1, 1, 0), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                nn.Conv2d(dim // 4, dim, 3, 1, 1))
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=0, embed_dim=dim,
                                      norm_layer=None)
        self.patch_unembed = PatchUnEmbed(img_size=img_size, patch_size=patch_size, in_chans=0, embed_dim=dim,
                                          norm_layer=None)

    def forward_nhwc(self, t):
        return _resi_conv(self.conv, self.residual_group.forward_nhwc(t), t)

    def forward_nhwc_folded(self, t, stats):
        """Evaluation on big tiles: (RSTB(t), row statistics of it or None).  The '1conv' residual connection emits the
        (mean, rstd) of the rows it stores, so the NEXT layer's first block folds its LayerNorm as well."""
        grp = self.residual_group
        if not grp._fold_ok(t):
            return self.forward_nhwc(t), None
        r = grp.forward_nhwc(t, stats)
        if not isinstance(self.conv, nn.Conv2d) or self.conv.bias is None:
            return _resi_conv(self.conv, r, t), None
        cs = t.shape[-1]
        return ops.raw.tapgemm(r, swin_ops._packed(self.conv.weight, 'fprop', cs, cs), ksize=3, cout=cs,
                               bias=swin_ops._padded_bias(self.conv.bias, cs), residual=t,
                               ln_out=(self.dim, grp.blocks[0].norm1.eps))

    def forward(self, x, x_size):
        """Reference contract (:557-558): x [B, h*w, C] tokens."""
        require_cuda(x, 'RSTB')
        y = self.forward_nhwc(swin_ops.tokens_to_nhwc(x, *x_size))
        return swin_ops.nhwc_to_tokens(y, self.dim, x.dtype)

    def flops(self):
        h, w = self.input_resolution
        return self.residual_group.flops() + h * w * self.dim * self.dim * 9 + self.patch_embed.flops() + \
            self.patch_unembed.flops()


def _resi_conv(conv, r, skip):
    """'1conv' / '3conv' residual connection (reference :532-539, :818-824) + skip, on NHWC bf16."""
    if isinstance(conv, nn.Conv2d):
        return ops.conv_nhwc(r, conv.weight, conv.bias, residual=skip)
    c0, c2, c4 = conv[0], conv[2], conv[4]
    r = ops.conv_nhwc(r, c0.weight, c0.bias, act='lrelu', slope=0.2)
    r = ops.conv_nhwc(r, c2.weight, c2.bias, act='lrelu', slope=0.2)
    return ops.conv_nhwc(r, c4.weight, c4.bias, residual=skip)


class UpsampleOneStep(nn.Sequential):
    """conv(num_feat -> scale^2 * num_out_ch) + PixelShuffle for lightweight SR (reference :669-690)."""

    def __init__(self, scale, num_feat, num_out_ch, input_resolution=None):
        self.num_feat = num_feat
        self.input_resolution = input_resolution
        m = [nn.Conv2d(num_feat, (scale**2) * num_out_ch, 3, 1, 1), nn.PixelShuffle(scale)]
        super(UpsampleOneStep, self).__init__(*m)

    def flops(self):
        h, w = self.input_resolution
        return h * w * self.num_feat * 3 * 9


@ARCH_REGISTRY.register()
class SwinIR(ArchMixin, nn.Module):
    # captured with programmatic dependent launch (archs/graphed.py): +0.7-1 %; SRB_SWIN_PDL=0 turns it off
    graph_pdl = os.environ.get('SRB_SWIN_PDL', '1') != '0'
    """SwinIR (classical / lightweight / real-world SR, denoising) -- same constructor as the reference."""

    def __init__(self, img_size=64, patch_size=1, in_chans=3, embed_dim=96, depths=(6, 6, 6, 6),
                 num_heads=(6, 6, 6, 6), window_size=7, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, upscale=2, img_range=1., upsampler='', resi_connection='1conv', **kwargs):
        super(SwinIR, self).__init__()
        # optional extras tolerated through **kwargs like any unknown key of the reference (:744):
        # cuda_graph / graph_segments replay the training fwd/bwd from CUDA graphs (archs/graphed.py)
        self.cuda_graph = bool(kwargs.get('cuda_graph', False))
        self.graph_segments = int(kwargs.get('graph_segments', 6))
        self.graph_input_shape = kwargs.get('graph_input_shape', None)
        self.flat_grads = bool(kwargs.get('flat_grads', False))  # ONE flat gradient buffer (utils/flat_ddp.py)
        # 'fp32': evaluation in fp32-class arithmetic (north_star "1e-4 in the TF32/fp32 mode"): split-bf16 tap-GEMMs,
        # fp32 LayerNorm / window attention kernels (ops/sr_b200/fp32_mode.py, csrc/swin_f32.cu)
        self.compute_dtype = kwargs.get('compute_dtype', 'bf16')
        if self.compute_dtype not in ('bf16', 'fp32'):
            raise ValueError(f"compute_dtype must be 'bf16' or 'fp32', got {self.compute_dtype!r}")
        num_in_ch = in_chans
        num_out_ch = in_chans
        num_feat = 64
        self.img_range = img_range
        if in_chans == 3:
            self.mean = torch.Tensor((0.4488, 0.4371, 0.4040)).view(1, 3, 1, 1)
        else:
            self.mean = torch.zeros(1, 1, 1, 1)
        self.upscale = upscale
        self.upsampler = upsampler

        self.conv_first = nn.Conv2d(num_in_ch, embed_dim, 3, 1, 1)

        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.ape = ape
        self.patch_norm = patch_norm
        self.num_features = embed_dim
        self.mlp_ratio = mlp_ratio

        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=embed_dim,
                                      embed_dim=embed_dim, norm_layer=norm_layer if self.patch_norm else None)
        num_patches = self.patch_embed.num_patches
        patches_resolution = self.patch_embed.patches_resolution
        self.patches_resolution = patches_resolution
        self.patch_unembed = PatchUnEmbed(img_size=img_size, patch_size=patch_size, in_chans=embed_dim,
                                          embed_dim=embed_dim, norm_layer=norm_layer if self.patch_norm else None)

        if self.ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
            trunc_normal_(self.absolute_pos_embed, std=.02)
        self.pos_drop = nn.Dropout(p=drop_rate)

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i_layer in range(self.num_layers):
            self.layers.append(
                RSTB(dim=embed_dim, input_resolution=(patches_resolution[0], patches_resolution[1]),
                     depth=depths[i_layer], num_heads=num_heads[i_layer], window_size=window_size,
                     mlp_ratio=self.mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                     attn_drop=attn_drop_rate, drop_path=dpr[sum(depths[:i_layer]):sum(depths[:i_layer + 1])],
                     norm_layer=norm_layer, downsample=None, use_checkpoint=use_checkpoint, img_size=img_size,
                     patch_size=patch_size, resi_connection=resi_connection))
        self.norm = norm_layer(self.num_features)

        if resi_connection == '1conv':
            self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        elif resi_connection == '3conv':
            self.conv_after_body = nn.Sequential(
                nn.Conv2d(embed_dim, embed_dim // 4, 3, 1, 1), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                nn.Conv2d(embed_dim // 4, embed_dim // 4, 1, 1, 0), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                nn.Conv2d(embed_dim // 4, embed_dim, 3, 1, 1))

        if self.upsampler == 'pixelshuffle':
            self.conv_before_upsample = nn.Sequential(
                nn.Conv2d(embed_dim, num_feat, 3, 1, 1), nn.LeakyReLU(inplace=True))
            self.upsample = Upsample(upscale, num_feat)
            self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
        elif self.upsampler == 'pixelshuffledirect':
            self.upsample = UpsampleOneStep(upscale, embed_dim, num_out_ch,
                                            (patches_resolution[0], patches_resolution[1]))
        elif self.upsampler == 'nearest+conv':
            assert self.upscale == 4, 'only support x4 now.'
            self.conv_before_upsample = nn.Sequential(
                nn.Conv2d(embed_dim, num_feat, 3, 1, 1), nn.LeakyReLU(inplace=True))
            self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
            self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
            self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
            self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
            self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        else:
            self.conv_last = nn.Conv2d(embed_dim, num_out_ch, 3, 1, 1)

        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'absolute_pos_embed'}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table'}

    def _embed(self, t):
        """patch_embed(+norm) [+ absolute position embedding] [+ pos_drop] (reference :878-881)."""
        t = self.patch_embed.forward_nhwc(t)
        if self.ape:
            b, h, w, cp = t.shape
            pe = self.absolute_pos_embed
            if pe.shape[1] != h * w:
                raise RuntimeError(f'ape=True: absolute_pos_embed holds {pe.shape[1]} positions, the input has {h * w} '
                                   '(same restriction as the reference, :880)')
            pe = nn.functional.pad(pe.view(1, h, w, -1), (0, cp - pe.shape[-1]))
            t = (t.float() + pe).to(torch.bfloat16)
        if self.training and self.pos_drop.p > 0.:
            t = self.pos_drop(t)
        return t

    def forward_features_nhwc(self, t):
        """patch_embed(+norm) -> RSTBs -> norm -> patch_unembed, all on [B,H,W,pad64(C)] bf16 (reference :876-889)."""
        t = self._embed(t)
        for layer in self.layers:
            t = layer.forward_nhwc(t)
        return swin_ops.layer_norm(t, self.norm.weight, self.norm.bias, self.norm.eps)

    def flops(self):
        """Reference :922-933."""
        h, w = self.patches_resolution
        flops = h * w * 3 * self.embed_dim * 9 + self.patch_embed.flops()
        flops += sum(layer.flops() for layer in self.layers)
        flops += h * w * 3 * self.embed_dim * self.embed_dim
        if hasattr(self.upsample, 'flops') if hasattr(self, 'upsample') else False:
            flops += self.upsample.flops()
        return flops

    def _head(self, x):
        mean = self._device_mean(x) if self.mean.numel() == x.shape[1] else None
        t = ops.image_to_nhwc(x, mean, self.img_range, ops.pad64(x.shape[1]))
        return ops.conv_nhwc(t, self.conv_first.weight, self.conv_first.bias)

    def _tail(self, t, first, x=None):
        mean = getattr(self, '_mean_dev', None) if self.mean.numel() == self.conv_first.weight.shape[1] else None
        t = swin_ops.layer_norm(t, self.norm.weight, self.norm.bias, self.norm.eps)
        feat = _resi_conv(self.conv_after_body, t, first)
        inv = 1.0 / self.img_range
        if self.upsampler == 'pixelshuffle':
            c = self.conv_before_upsample[0]
            u = ops.conv_nhwc(feat, c.weight, c.bias, act='lrelu', slope=self.conv_before_upsample[1].negative_slope)
            u = self.upsample.forward_nhwc(u)
            return ops.conv_to_image(u, self.conv_last.weight, self.conv_last.bias, inv, mean)
        if self.upsampler == '':
            # denoising / JPEG-CAR: x + conv_last(res) in normalised space == x + conv_last(res) / img_range
            return x.float() + ops.conv_to_image(feat, self.conv_last.weight, self.conv_last.bias, inv, None)
        if self.upsampler == 'pixelshuffledirect':
            # lightweight SR (reference :901-905): one conv to scale^2 * num_out_ch channels + PixelShuffle(scale)
            conv, shuffle = self.upsample[0], self.upsample[1]
            u, u32 = ops.conv_nhwc(feat, conv.weight, conv.bias, want_f32=True)
            return ops.shuffle_to_image(u, u32, conv.weight.shape[0] // shuffle.upscale_factor**2,
                                        shuffle.upscale_factor, inv, mean)
        if self.upsampler == 'nearest+conv':
            # real-world SR (reference :906-913): conv+lrelu, 2 x (nearest x2 -> conv -> lrelu 0.2), conv_hr, conv_last
            c = self.conv_before_upsample[0]
            u = ops.conv_nhwc(feat, c.weight, c.bias, act='lrelu', slope=self.conv_before_upsample[1].negative_slope)
            s02 = self.lrelu.negative_slope
            u = ops.conv_nhwc(ops.nearest_up2(u), self.conv_up1.weight, self.conv_up1.bias, act='lrelu', slope=s02)
            u = ops.conv_nhwc(ops.nearest_up2(u), self.conv_up2.weight, self.conv_up2.bias, act='lrelu', slope=s02)
            u = ops.conv_nhwc(u, self.conv_hr.weight, self.conv_hr.bias, act='lrelu', slope=s02)
            return ops.conv_to_image(u, self.conv_last.weight, self.conv_last.bias, inv, mean)
        raise ValueError(f"unknown upsampler '{self.upsampler}'")

    def _build_segments(self):
        groups = split_even(list(self.layers), self.graph_segments)

        def run(gs):
            def fn(t):
                for layer in gs:
                    t = layer.forward_nhwc(t)
                return t
            return fn

        def head_fn(x):
            first = self._head(x)
            return run(groups[0])(self._embed(first)), first

        tail_mods = [self.norm, self.conv_after_body] + [getattr(self, n) for n in
                                                         ('conv_before_upsample', 'upsample', 'conv_up1', 'conv_up2',
                                                          'conv_hr', 'conv_last') if hasattr(self, n)]
        segs = [Segment(head_fn, [self.conv_first, self.patch_embed] + groups[0])]
        segs += [Segment(run(g), g) for g in groups[1:]]
        segs.append(Segment(self._tail, tail_mods))
        return segs

    # ------------------------------------------------------------------ fp32 mode (evaluation)
    def _fp32_supported(self):
        return all(blk.on_kernels() for layer in self.layers for blk in layer.residual_group.blocks) and \
            (self.patch_embed.norm is None or type(self.patch_embed.norm) is nn.LayerNorm) and \
            type(self.norm) is nn.LayerNorm

    @staticmethod
    def _block_fp32(blk, t32):
        """SwinTransformerBlock.forward (reference :283-323) on fp32 NHWC [B,H,W,pad64(C)], evaluation mode (DropPath and
        dropout are the identity): 4 split-operand tap-GEMMs, 2 fp32 LayerNorms and the fp32 window attention, the two
        skip adds in the proj / fc2 epilogues."""
        from ..ops.sr_b200 import fp32_mode as f32
        from .. import _lib as L
        at, m = blk.attn, blk.mlp
        nh = blk.num_heads
        hd = blk.dim // nh
        cs = t32.shape[-1]
        ca = nh * swin_ops.HD_PAD
        p_qkv = swin_ops.head_perm(nh, hd, 3, t32.device)
        p_o = swin_ops.head_perm(nh, hd, 1, t32.device)
        qkv = f32.conv_split(f32.layer_norm(t32, blk.norm1, split=True), at.qkv.weight, at.qkv.bias, n_pad=3 * ca,
                             perm_out=p_qkv)
        o = f32.window_attention(qkv, at.relative_position_bias_table, nh, blk.window_size, blk.shift_size, at.scale)
        x1 = f32.conv_split(o, at.proj.weight, at.proj.bias, n_pad=cs, perm_in=p_o, residual32=t32)
        h = f32.conv_split(f32.layer_norm(x1, blk.norm2, split=True), m.fc1.weight, m.fc1.bias, act=L.ACT_GELU)
        return f32.conv(h, m.fc2.weight, m.fc2.bias, residual32=x1)

    @staticmethod
    def _resi_conv_fp32(conv, rs, skip32):
        """'1conv' / '3conv' residual connection + skip (reference :532-539, :818-824) in fp32 mode; ``rs`` is the
        split operand [hi | lo | hi] of the first conv."""
        from ..ops.sr_b200 import fp32_mode as f32
        from .. import _lib as L
        if isinstance(conv, nn.Conv2d):
            return f32.conv_split(rs, conv.weight, conv.bias, residual32=skip32)
        c0, c2, c4 = conv[0], conv[2], conv[4]
        r32 = f32.conv_split(rs, c0.weight, c0.bias, act=L.ACT_LRELU, slope=conv[1].negative_slope)
        r32 = f32.conv(r32, c2.weight, c2.bias, act=L.ACT_LRELU, slope=conv[3].negative_slope)
        return f32.conv(r32, c4.weight, c4.bias, residual32=skip32)

    def _forward_fp32(self, x):
        """swinir_arch.py:876-921 in fp32 mode (no autograd, evaluation): fp32 NHWC activations end to end."""
        from ..ops.sr_b200 import fp32_mode as f32
        from .. import _lib as L
        cin = x.shape[1]
        mean = self._device_mean(x) if self.mean.numel() == cin else None
        first = f32.conv(f32.image_to_nhwc32(x, mean, self.img_range, ops.pad64(cin)), self.conv_first.weight,
                         self.conv_first.bias)
        t = first
        if self.patch_embed.norm is not None:
            t = f32.layer_norm(t, self.patch_embed.norm)
        if self.ape:
            b, h, w, cp = t.shape
            pe = self.absolute_pos_embed
            if pe.shape[1] != h * w:
                raise RuntimeError(f'ape=True: absolute_pos_embed holds {pe.shape[1]} positions, the input has {h * w} '
                                   '(same restriction as the reference, :880)')
            t = t + nn.functional.pad(pe.detach().float().view(1, h, w, -1), (0, cp - pe.shape[-1]))
        for layer in self.layers:
            r = t
            for blk in layer.residual_group.blocks:
                r = self._block_fp32(blk, r)
            t = self._resi_conv_fp32(layer.conv, f32.split3(r), t)
        feat = self._resi_conv_fp32(self.conv_after_body, f32.layer_norm(t, self.norm, split=True), first)
        inv = 1.0 / self.img_range
        if self.upsampler == 'pixelshuffle':
            c = self.conv_before_upsample[0]
            u = f32.conv(feat, c.weight, c.bias, act=L.ACT_LRELU, slope=self.conv_before_upsample[1].negative_slope)
            mods = list(self.upsample)
            for conv, shuffle in zip(mods[0::2], mods[1::2]):
                u = f32.conv(u, conv.weight, conv.bias, shuffle_r=shuffle.upscale_factor)
            return f32.conv_to_image(u, self.conv_last.weight, self.conv_last.bias, inv, mean)
        if self.upsampler == '':
            return x.float() + f32.conv_to_image(feat, self.conv_last.weight, self.conv_last.bias, inv, None)
        if self.upsampler == 'pixelshuffledirect':
            conv, shuffle = self.upsample[0], self.upsample[1]
            r = shuffle.upscale_factor
            u = f32.conv(feat, conv.weight, conv.bias)[..., :conv.weight.shape[0]]  # [B,H,W,c_img*r*r], channel c*r*r+i*r+j
            b, h, w, _ = u.shape
            img = u.reshape(b, h, w, -1, r, r).permute(0, 3, 1, 4, 2, 5).reshape(b, -1, h * r, w * r) * inv
            return img + mean.view(1, -1, 1, 1) if mean is not None else img
        if self.upsampler == 'nearest+conv':
            def up2(z):  # F.interpolate(scale_factor=2, mode='nearest') on NHWC: a pure index remap
                b, h, w, c = z.shape
                return z[:, :, None, :, None, :].expand(b, h, 2, w, 2, c).reshape(b, 2 * h, 2 * w, c)
            c = self.conv_before_upsample[0]
            s02 = self.lrelu.negative_slope
            u = f32.conv(feat, c.weight, c.bias, act=L.ACT_LRELU, slope=self.conv_before_upsample[1].negative_slope)
            u = f32.conv(up2(u), self.conv_up1.weight, self.conv_up1.bias, act=L.ACT_LRELU, slope=s02)
            u = f32.conv(up2(u), self.conv_up2.weight, self.conv_up2.bias, act=L.ACT_LRELU, slope=s02)
            u = f32.conv(u, self.conv_hr.weight, self.conv_hr.bias, act=L.ACT_LRELU, slope=s02)
            return f32.conv_to_image(u, self.conv_last.weight, self.conv_last.bias, inv, mean)
        raise ValueError(f"unknown upsampler '{self.upsampler}'")

    def _forward(self, x):
        require_cuda(x, 'SwinIR')
        if self.compute_dtype == 'fp32' and not torch.is_grad_enabled() and not self.training:
            if not self._fp32_supported():
                raise NotImplementedError("compute_dtype='fp32' covers the configurations of the fused Swin block "
                                          '(window_size <= 8, even head count, head_dim <= 32, nn.LayerNorm, GELU Mlp)')
            out = self._forward_fp32(x)
            return out if out.dtype == x.dtype else out.to(x.dtype)
        if self.cuda_graph and self.training and torch.is_grad_enabled() and self.upsampler != '':
            nseg = len(split_even(list(self.layers), self.graph_segments)) + 1
            out = graphed_forward(self, x, self._build_segments, chain_wire(nseg, carry=1))
            return out if out.dtype == x.dtype else out.to(x.dtype)
        first = self._head(x)
        t = self._embed(first)
        if not torch.is_grad_enabled() and not self.training:
            stats = None  # (evaluation: row statistics handed from one RSTB's conv to the next layer's first block)
            for layer in self.layers:
                t, stats = layer.forward_nhwc_folded(t, stats)
        else:
            for layer in self.layers:
                t = layer.forward_nhwc(t)
        out = self._tail(t, first, x)
        return out if out.dtype == x.dtype else out.to(x.dtype)
