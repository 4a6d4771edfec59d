"""Tiled inference over large remote-sensing scenes, sharded across the GPUs of one box.

BASELINE.json config 5 treats a scene as a grid of independent LR tiles (no halo exchange): rank ``r`` of ``W``
takes tiles ``r, r + W, ...`` and runs its replica of the network on them under ``no_grad`` -- replicas only, no
collective on the data path.  The reference itself has no tiling
(inference/inference_swinir.py:57-69 pads and forwards the whole image); these helpers are what a caller adds
around ``SRModel.test`` (sr_model.py:120-129) / ``SwinIRModel.test`` (swinir_model.py:14-36).
"""
import torch


def tile_grid(height, width, tile, overlap=0, uniform=False):
    """Top-left corners and sizes of the LR tiles covering an ``height x width`` image.

    Tiles are ``tile x tile`` (clipped at the borders, never empty); consecutive tiles overlap by ``overlap``
    pixels so a caller can blend or crop seams.  Returns a list of ``(y, x, h, w)`` in row-major order.
    ``uniform``: every pair of neighbours overlaps by exactly ``overlap`` and the last tile of a row / column is
    clipped instead of shifted back (needs ``overlap <= tile / 2``) -- then no pixel lies in more than two tiles per
    axis, which is what linear-ramp blending assumes."""
    if tile <= 0 or overlap < 0 or overlap >= tile:
        raise ValueError(f'need tile > overlap >= 0, got tile={tile} overlap={overlap}')
    if uniform and 2 * overlap > tile:
        raise ValueError(f'uniform grids need overlap <= tile / 2, got tile={tile} overlap={overlap}')
    if height <= 0 or width <= 0:
        return []
    step = tile - overlap

    def starts(size):
        if size <= tile:
            return [0]
        if uniform:
            return list(range(0, size - overlap, step))
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


# ------------------------------------------------------------------ end-to-end scene driver
def band_rows(height, rank, world, multiple=1):
    """Rows [y0, y1) of the LR scene that ``rank`` of ``world`` owns: contiguous bands of near-equal height, cut at
    multiples of ``multiple``.  Contiguous (not round-robin) ownership is what lets overlapping tiles be BLENDED without
    any exchange: a rank computes every tile that touches its band, tiles on a band border are computed by both
    neighbours (replicas only)."""
    if not 0 <= rank < world:
        raise ValueError(f'rank {rank} outside world {world}')
    units = (height + multiple - 1) // multiple
    y0 = (units * rank // world) * multiple
    y1 = min(height, (units * (rank + 1) // world) * multiple)
    return y0, y1


def neighbour_overlaps(tiles, index):
    """(top, bottom, left, right) overlap in LR pixels of tile ``index`` with its grid neighbours (0 at the borders).
    ``tiles`` is the row-major list of :func:`tile_grid`."""
    y, x, th, tw = tiles[index]
    ys = sorted({t[0] for t in tiles})
    xs = sorted({t[1] for t in tiles})
    iy, ix = ys.index(y), xs.index(x)
    height = {t[0]: t[2] for t in tiles}
    width = {t[1]: t[3] for t in tiles}
    top = (ys[iy - 1] + height[ys[iy - 1]] - y) if iy > 0 else 0
    bottom = (y + th - ys[iy + 1]) if iy + 1 < len(ys) else 0
    left = (xs[ix - 1] + width[xs[ix - 1]] - x) if ix > 0 else 0
    right = (x + tw - xs[ix + 1]) if ix + 1 < len(xs) else 0
    return max(top, 0), max(bottom, 0), max(left, 0), max(right, 0)


class TiledUpscaler:
    """uint8 scene in pinned host memory -> super-resolved uint8 scene in pinned host memory, tile by tile.

    Per tile: H2D of the uint8 LR tile on a copy stream, ``srb200_patch_from_u8`` (BGR->RGB, HWC->CHW, /255),
    window padding, the network, ``srb200_tile_blend_add`` into the rank's fp32 band accumulator (linear-ramp seams;
    ``guard`` LR pixels at every interior tile edge carry no weight -- choose it >= the network's receptive-field
    radius and the blended scene equals the whole-image result); after the last tile ``srb200_tensor2img_u8``
    (clamp, x255, round, RGB->BGR -- ``tensor2img``, img_util.py:40-96) and one D2H of the band.  The reference has no
    tiling (inference/inference_swinir.py:57-69 forwards the whole image); this is what a caller puts around
    ``SRModel.test`` / ``SwinIRModel.test`` for remote-sensing scenes.  Replicas only: no collective."""

    def __init__(self, net, scale, tile, overlap=0, guard=0, multiple=1, bgr=True):
        if overlap and overlap <= 2 * guard:
            raise ValueError('overlap must exceed 2 * guard (the ramp lives between the two guards)')
        self.net, self.scale, self.tile, self.overlap, self.guard = net, scale, tile, overlap, guard
        self.multiple, self.bgr = multiple, bgr
        self.copy_stream = None

    @torch.no_grad()
    def upscale(self, scene_u8, rank=0, world=1, out=None):
        """``scene_u8``: uint8 [H, W, C] host tensor (pinned for async copies).  Returns (band uint8 [s*(y1-y0), s*W, C]
        host tensor, (y0, y1), tiles_done, bytes_h2d, bytes_d2h)."""
        from ..ops.sr_b200 import raw
        dev = next(self.net.parameters()).device
        h, w, c = scene_u8.shape
        s = self.scale
        y0, y1 = band_rows(h, rank, world, self.multiple)
        tiles = tile_grid(h, w, self.tile, self.overlap, uniform=True)
        mine = [i for i, (ty, tx, th, tw) in enumerate(tiles) if ty < y1 and ty + th > y0]
        with torch.cuda.device(dev):
            if self.copy_stream is None:
                self.copy_stream = torch.cuda.Stream()
            main = torch.cuda.current_stream()
            acc = torch.zeros((c, (y1 - y0) * s, w * s), dtype=torch.float32, device=dev)
            staged = {}
            h2d = 0

            def stage(i):  # async H2D of tile i (its rows are contiguous slabs of the pinned scene)
                ty, tx, th, tw = tiles[i]
                with torch.cuda.stream(self.copy_stream):
                    t = scene_u8[ty:ty + th, tx:tx + tw].to(dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
                staged[i] = (t, ev)
                return t.numel()

            if mine:
                h2d += stage(mine[0])
            for k, i in enumerate(mine):
                if k + 1 < len(mine):
                    h2d += stage(mine[k + 1])  # the next tile's copy overlaps this tile's compute
                t, ev = staged.pop(i)
                main.wait_event(ev)
                ty, tx, th, tw = tiles[i]
                t = t.contiguous()
                lr = raw.patch_from_u8([t], [0], [0], [0], th, tw, bgr2rgb=self.bgr)
                lr, (oh, ow) = pad_to_multiple(lr, self.multiple)
                sr = self.net(lr)[0, :, :oh * s, :ow * s].contiguous().float()
                ov = [v * s for v in neighbour_overlaps(tiles, i)]
                raw.tile_blend_add(sr, acc, (ty - y0) * s, tx * s, ov, self.guard * s)
                t.record_stream(main)
            band = raw.tensor2img_u8(acc, 0.0, 1.0, rgb2bgr=self.bgr)
            if out is None:
                out = torch.empty(band.shape, dtype=torch.uint8).pin_memory()
            out.copy_(band, non_blocking=True)
            main.synchronize()
        return out, (y0, y1), len(mine), h2d, band.numel()
