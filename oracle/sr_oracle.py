"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference SR hot path.

A *functional* restatement (state-dict in, tensor out; no nn.Module, no registry) of what
watercore2001/BasicSR4RS computes in basicsr/archs/{arch_util,edsr_arch,rcan_arch,swinir_arch}.py.
All arithmetic is stock ``torch`` fp32 (the reference's only dependency on this path is PyTorch,
requirements.txt:1), run on the CPU.  It is the checker for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it;
the product package ``basicsr4rs_b200`` never does.

Pinning: the reference's own tests hold no golden vectors for EDSR / RCAN / SwinIR (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference itself: ``tests/golden/make_golden.py``
imports the unmodified reference archs (oracle/ref_shim.py) in the build container and commits
input / output / gradient fixtures; ``tests/test_oracle.py`` checks this file against those
fixtures everywhere and against the live reference where /root/reference is mounted.
"""
import math
import zlib

import torch
import torch.nn.functional as F

RGB_MEAN = (0.4488, 0.4371, 0.4040)  # edsr_arch.py:38, rcan_arch.py:106, swinir_arch.py:751


# ------------------------------------------------------------------ deterministic weights
def fill_state_dict_(sd, salt=0):
    """Fill a state dict *by key name* (stable crc32 seed) so every process / machine agrees without
    shipping weights.  Integer buffers and masks are left untouched."""
    for key, t in sd.items():
        if not t.is_floating_point() or key.endswith('attn_mask'):
            continue
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + salt) & 0x7FFFFFFF)
        if key.endswith('relative_position_bias_table'):
            v = torch.randn(t.shape, generator=g) * 0.2
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            v = torch.randn(t.shape, generator=g) * (0.7 / math.sqrt(fan_in))
        elif 'norm' in key and key.endswith('weight'):
            v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        else:
            v = 0.05 * torch.randn(t.shape, generator=g)
        t.copy_(v.to(t.dtype))
    return sd


# ------------------------------------------------------------------ shared blocks
def conv(sd, prefix, x, padding=1):
    return F.conv2d(x, sd[prefix + '.weight'], sd.get(prefix + '.bias'), stride=1, padding=padding)


def residual_block_nobn(sd, prefix, x, res_scale):
    """arch_util.py:85-88: identity + conv2(relu(conv1(x))) * res_scale."""
    out = conv(sd, prefix + '.conv2', F.relu(conv(sd, prefix + '.conv1', x)))
    return x + out * res_scale


def upsample(sd, prefix, x, scale):
    """arch_util.py:131-142: [conv F->4F, PixelShuffle(2)] x log2(scale), or conv F->9F + PixelShuffle(3)."""
    if scale & (scale - 1) == 0:
        for i in range(int(math.log2(scale))):
            x = F.pixel_shuffle(conv(sd, f'{prefix}.{2 * i}', x), 2)
    elif scale == 3:
        x = F.pixel_shuffle(conv(sd, f'{prefix}.0', x), 3)
    else:
        raise ValueError(f'scale {scale} is not supported. Supported scales: 2^n and 3.')
    return x


def _mean(x, rgb_mean):
    return torch.tensor(rgb_mean, dtype=x.dtype, device=x.device).view(1, -1, 1, 1)


# ------------------------------------------------------------------ EDSR (edsr_arch.py:50-61)
def edsr_forward(sd, x, num_block=16, upscale=4, res_scale=1.0, img_range=255., rgb_mean=RGB_MEAN):
    mean = _mean(x, rgb_mean)
    x = (x - mean) * img_range
    x = conv(sd, 'conv_first', x)
    res = x
    for i in range(num_block):
        res = residual_block_nobn(sd, f'body.{i}', res, res_scale)
    res = conv(sd, 'conv_after_body', res) + x
    x = conv(sd, 'conv_last', upsample(sd, 'upsample', res, upscale))
    return x / img_range + mean


# ------------------------------------------------------------------ RCAN (rcan_arch.py)
def channel_attention(sd, prefix, x):
    """rcan_arch.py:16-24: x * sigmoid(W2 relu(W1 avgpool(x) + b1) + b2)."""
    y = x.mean(dim=(2, 3), keepdim=True)
    y = F.relu(conv(sd, prefix + '.attention.1', y, padding=0))
    y = torch.sigmoid(conv(sd, prefix + '.attention.3', y, padding=0))
    return x * y


def rcab(sd, prefix, x, res_scale):
    """rcan_arch.py:36-46."""
    r = conv(sd, prefix + '.rcab.2', F.relu(conv(sd, prefix + '.rcab.0', x)))
    r = channel_attention(sd, prefix + '.rcab.3', r)
    return r * res_scale + x


def rcan_forward(sd, x, num_group=10, num_block=16, upscale=4, res_scale=1.0, img_range=255., rgb_mean=RGB_MEAN):
    """rcan_arch.py:124-135; ResidualGroup :66-68."""
    mean = _mean(x, rgb_mean)
    x = (x - mean) * img_range
    x = conv(sd, 'conv_first', x)
    res = x
    for g in range(num_group):
        t = res
        for b in range(num_block):
            t = rcab(sd, f'body.{g}.residual_group.{b}', t, res_scale)
        res = conv(sd, f'body.{g}.conv', t) + res
    res = conv(sd, 'conv_after_body', res) + x
    x = conv(sd, 'conv_last', upsample(sd, 'upsample', res, upscale))
    return x / img_range + mean


# ------------------------------------------------------------------ SwinIR (swinir_arch.py)
def window_partition(x, ws):
    """swinir_arch.py:63-75."""
    b, h, w, c = x.shape
    x = x.reshape(b, h // ws, ws, w // ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, c)


def window_reverse(win, ws, h, w):
    """swinir_arch.py:78-92."""
    b = win.shape[0] // ((h // ws) * (w // ws))
    x = win.reshape(b, h // ws, w // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, -1)


def relative_position_index(ws):
    """swinir_arch.py:122-133."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing='ij')).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def calculate_mask(h, w, ws, shift):
    """swinir_arch.py:262-281: 0 / -100 mask of the shifted-window regions."""
    img = torch.zeros((1, h, w, 1))
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).reshape(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return torch.where(m != 0, torch.full_like(m, -100.0), torch.zeros_like(m))


def window_attention(sd, prefix, xw, mask, num_heads, ws):
    """swinir_arch.py:144-175."""
    b_, n, c = xw.shape
    hd = c // num_heads
    qkv = F.linear(xw, sd[prefix + '.qkv.weight'], sd.get(prefix + '.qkv.bias'))
    qkv = qkv.reshape(b_, n, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd**-0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    idx = relative_position_index(ws).reshape(-1).to(xw.device)
    bias = sd[prefix + '.relative_position_bias_table'][idx].reshape(n, n, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        attn = attn.reshape(b_ // nw, nw, num_heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.reshape(-1, num_heads, n, n)
    attn = torch.softmax(attn, dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(b_, n, c)
    return F.linear(out, sd[prefix + '.proj.weight'], sd[prefix + '.proj.bias'])


def swin_block(sd, prefix, x, h, w, num_heads, ws, shift, drop_keep=None):
    """swinir_arch.py:283-323 (eval mode / drop_path 0 unless ``drop_keep`` = (mask1, mask2, keep_prob))."""
    b, _, c = x.shape
    if min(h, w) <= ws:  # swinir_arch.py:233-236
        shift, ws = 0, min(h, w)
    shortcut = x
    t = F.layer_norm(x, (c,), sd[prefix + '.norm1.weight'], sd[prefix + '.norm1.bias'], 1e-5).reshape(b, h, w, c)
    if shift > 0:
        t = torch.roll(t, shifts=(-shift, -shift), dims=(1, 2))
    xw = window_partition(t, ws).reshape(-1, ws * ws, c)
    mask = calculate_mask(h, w, ws, shift).to(x.device) if shift > 0 else None
    aw = window_attention(sd, prefix + '.attn', xw, mask, num_heads, ws).reshape(-1, ws, ws, c)
    t = window_reverse(aw, ws, h, w)
    if shift > 0:
        t = torch.roll(t, shifts=(shift, shift), dims=(1, 2))
    t = t.reshape(b, h * w, c)
    if drop_keep is not None:
        t = t / drop_keep[2] * drop_keep[0]
    x = shortcut + t
    m = F.layer_norm(x, (c,), sd[prefix + '.norm2.weight'], sd[prefix + '.norm2.bias'], 1e-5)
    m = F.linear(F.gelu(F.linear(m, sd[prefix + '.mlp.fc1.weight'], sd[prefix + '.mlp.fc1.bias'])),
                 sd[prefix + '.mlp.fc2.weight'], sd[prefix + '.mlp.fc2.bias'])
    if drop_keep is not None:
        m = m / drop_keep[2] * drop_keep[1]
    return x + m


def drop_path_masks(batch, depths, drop_path_rate, generator=None):
    """Per-block stochastic-depth draws of a TRAINING forward, in the order the reference consumes its RNG
    (swinir_arch.py:14-26 called twice per block from :320-321; rates from the linspace of :796): a dict
    block prefix -> (mask1 [B,1,1], mask2 [B,1,1], keep_prob), masks = floor(keep_prob + U[0,1)).  With
    ``generator=None`` the global CPU RNG is used, exactly like ``torch.rand`` inside the reference's ``drop_path``."""
    rates = [r.item() for r in torch.linspace(0, drop_path_rate, sum(depths))]
    out, i = {}, 0
    for li, depth in enumerate(depths):
        for bi in range(depth):
            rate = rates[i]
            i += 1
            if rate == 0.:
                continue
            keep = 1 - rate
            m1 = (keep + torch.rand((batch, 1, 1), generator=generator)).floor_()
            m2 = (keep + torch.rand((batch, 1, 1), generator=generator)).floor_()
            out[f'layers.{li}.residual_group.blocks.{bi}'] = (m1, m2, keep)
    return out


def swinir_forward(sd, x, embed_dim=180, depths=(6, 6, 6, 6, 6, 6), num_heads=(6, 6, 6, 6, 6, 6), window_size=8,
                   upscale=4, img_range=1., in_chans=3, upsampler='pixelshuffle', resi_connection='1conv',
                   drop_masks=None, ape=False):
    """swinir_arch.py:891-920: the classical-SR ('pixelshuffle') branch and the three other
    reconstruction branches ('pixelshuffledirect' :901-905, 'nearest+conv' :906-913, '' :914-918); conv_after_body as
    '1conv' or '3conv' (:818-826).  Eval mode unless ``drop_masks`` (see :func:`drop_path_masks`) supplies the
    stochastic-depth draws of a training forward; ``ape`` adds ``absolute_pos_embed`` (:879-880)."""
    def resi(prefix, v):
        if resi_connection == '1conv':
            return conv(sd, prefix, v)
        # '3conv' (:533-538, :821-826): conv3x3 C->C/4, lrelu 0.2, conv1x1, lrelu 0.2, conv3x3 C/4->C
        v = F.leaky_relu(conv(sd, prefix + '.0', v), 0.2)
        v = F.leaky_relu(conv(sd, prefix + '.2', v, padding=0), 0.2)
        return conv(sd, prefix + '.4', v)

    mean = _mean(x, RGB_MEAN) if in_chans == 3 else torch.zeros(1, 1, 1, 1, device=x.device)
    x_in = x
    x = (x - mean) * img_range
    x = conv(sd, 'conv_first', x)
    b, c, h, w = x.shape
    t = x.flatten(2).transpose(1, 2)  # PatchEmbed :600-604
    t = F.layer_norm(t, (c,), sd['patch_embed.norm.weight'], sd['patch_embed.norm.bias'], 1e-5)
    if ape:
        t = t + sd['absolute_pos_embed']
    for li, depth in enumerate(depths):
        r = t
        for bi in range(depth):
            shift = 0 if bi % 2 == 0 else window_size // 2
            prefix = f'layers.{li}.residual_group.blocks.{bi}'
            r = swin_block(sd, prefix, r, h, w, num_heads[li], window_size, shift,
                           drop_keep=(drop_masks or {}).get(prefix))
        r = r.transpose(1, 2).reshape(b, c, h, w)  # PatchUnEmbed :638-640
        r = resi(f'layers.{li}.conv', r)  # RSTB.conv :530-538
        t = r.flatten(2).transpose(1, 2) + t  # RSTB :557-558
    t = F.layer_norm(t, (c,), sd['norm.weight'], sd['norm.bias'], 1e-5)
    feat = t.transpose(1, 2).reshape(b, c, h, w)
    body = resi('conv_after_body', feat)
    if upsampler == '':  # denoising / JPEG CAR: image-space residual
        return ((x_in - mean) * img_range + conv(sd, 'conv_last', body + x)) / img_range + mean
    x = body + x
    if upsampler == 'pixelshuffle':
        x = F.leaky_relu(conv(sd, 'conv_before_upsample.0', x), 0.01)
        x = conv(sd, 'conv_last', upsample(sd, 'upsample', x, upscale))
    elif upsampler == 'pixelshuffledirect':
        x = F.pixel_shuffle(conv(sd, 'upsample.0', x), upscale)
    elif upsampler == 'nearest+conv':
        x = F.leaky_relu(conv(sd, 'conv_before_upsample.0', x), 0.01)
        x = F.leaky_relu(conv(sd, 'conv_up1', F.interpolate(x, scale_factor=2, mode='nearest')), 0.2)
        x = F.leaky_relu(conv(sd, 'conv_up2', F.interpolate(x, scale_factor=2, mode='nearest')), 0.2)
        x = conv(sd, 'conv_last', F.leaky_relu(conv(sd, 'conv_hr', x), 0.2))
    else:
        raise ValueError(upsampler)
    return x / img_range + mean


# ------------------------------------------------------------------ metric used by the parity bar
def psnr(a, b, crop=4):
    """basicsr/metrics/psnr_ssim.py:11-47 on [0,1] NCHW tensors: crop border, x255, float64 MSE."""
    a = a[..., crop:-crop, crop:-crop].double() * 255.0
    b = b[..., crop:-crop, crop:-crop].double() * 255.0
    mse = ((a - b)**2).mean().item()
    return float('inf') if mse == 0 else 10.0 * math.log10(255.0 * 255.0 / mse)
