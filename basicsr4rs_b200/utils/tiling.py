"""Tiled inference over large remote-sensing scenes, sharded across the GPUs of one box.

BASELINE.json config 5 treats a scene as a grid of independent LR tiles (no halo exchange): rank ``r`` of ``W``
takes tiles ``r, r + W, ...`` and runs its replica of the network on them under ``no_grad`` -- replicas only, no
collective on the data path.  The reference itself has no tiling
(inference/inference_swinir.py:57-69 pads and forwards the whole image); these helpers are what a caller adds
around ``SRModel.test`` (sr_model.py:120-129) / ``SwinIRModel.test`` (swinir_model.py:14-36).
"""
import torch


def tile_grid(height, width, tile, overlap=0):
    """Top-left corners and sizes of the LR tiles covering an ``height x width`` image.

    Tiles are ``tile x tile`` (clipped at the borders, never empty); consecutive tiles overlap by ``overlap``
    pixels so a caller can blend or crop seams.  Returns a list of ``(y, x, h, w)`` in row-major order."""
    if tile <= 0 or overlap < 0 or overlap >= tile:
        raise ValueError(f'need tile > overlap >= 0, got tile={tile} overlap={overlap}')
    if height <= 0 or width <= 0:
        return []
    step = tile - overlap

    def starts(size):
        if size <= tile:
            return [0]
        s = list(range(0, size - tile, step)) + [size - tile]
        return sorted(set(s))

    return [(y, x, min(tile, height - y), min(tile, width - x)) for y in starts(height) for x in starts(width)]


def shard(items, rank, world):
    """The work items of ``rank`` out of ``world`` (round-robin: ``items[rank::world]``)."""
    if not 0 <= rank < world:
        raise ValueError(f'rank {rank} outside world {world}')
    return list(items)[rank::world]


def pad_to_multiple(x, multiple, mode='reflect'):
    """Pad H, W of an NCHW tensor up to a multiple (SwinIRModel.test pads to the window size the same way,
    swinir_model.py:16-24).  Returns the padded tensor and the original (h, w)."""
    h, w = x.shape[-2:]
    ph, pw = (-h) % multiple, (-w) % multiple
    if ph == 0 and pw == 0:
        return x, (h, w)
    if mode == 'reflect':  # flip-concatenate like the reference (works for pads larger than the image too)
        x = torch.cat([x, torch.flip(x, [2])], 2)[:, :, :h + ph, :]
        x = torch.cat([x, torch.flip(x, [3])], 3)[:, :, :, :w + pw]
        return x, (h, w)
    return torch.nn.functional.pad(x, (0, pw, 0, ph), mode=mode), (h, w)


@torch.no_grad()
def tiled_forward(net, image, tile, scale, overlap=0, multiple=1, rank=0, world=1, out=None):
    """Super-resolve the tiles of ``image`` ([1,C,H,W], on the net's device) that belong to ``rank``.

    Each tile is forwarded independently; with ``overlap`` > 0 the inner half of every seam is kept (crop
    blending).  Writes into / returns an ``[1,C,scale*H,scale*W]`` tensor in which only this rank's tiles are
    filled; the caller gathers or writes tiles out (no collective here).  Returns (out, tiles_done)."""
    _, c, h, w = image.shape
    if out is None:
        out = torch.zeros((1, c, h * scale, w * scale), dtype=image.dtype, device=image.device)
    mine = shard(tile_grid(h, w, tile, overlap), rank, world)
    half = overlap // 2
    for (y, x, th, tw) in mine:
        lr, (oh, ow) = pad_to_multiple(image[:, :, y:y + th, x:x + tw], multiple)
        sr = net(lr)[:, :, :oh * scale, :ow * scale]
        # keep the inner part of overlapping seams (image borders keep everything)
        y0 = half if y > 0 else 0
        x0 = half if x > 0 else 0
        y1 = th - (half if y + th < h else 0)
        x1 = tw - (half if x + tw < w else 0)
        out[:, :, (y + y0) * scale:(y + y1) * scale, (x + x0) * scale:(x + x1) * scale] = \
            sr[:, :, y0 * scale:y1 * scale, x0 * scale:x1 * scale]
    return out, len(mine)
