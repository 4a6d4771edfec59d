"""Callers' side of the hot path on the GPU (``-m gpu``): the input pipeline (crop / flip / rot / uint8 -> fp32 CHW RGB)
against the reference's own host functions, tensor2img, and the tiled scene driver on the REAL archs -- a blended tiled
EDSR scene must equal the whole-image forward."""
import numpy as np
import pytest
import torch

from oracle import sr_oracle

pytestmark = pytest.mark.gpu


def _ref_patch(img_u8, top, left, p, flags):
    """paired_random_crop's slice + augment (cv2.flip / transpose) + img2tensor, restated with numpy
    (basicsr/data/transforms.py:74-90,190-200; utils/img_util.py:23-37)."""
    img = img_u8.astype(np.float32) / 255.0
    patch = img[top:top + p, left:left + p, :].copy()
    if flags & 1:
        patch = patch[:, ::-1, :]
    if flags & 2:
        patch = patch[::-1, :, :]
    if flags & 4:
        patch = patch.transpose(1, 0, 2)
    patch = patch[:, :, ::-1]  # BGR -> RGB
    return torch.from_numpy(np.ascontiguousarray(patch.transpose(2, 0, 1)))


def test_patch_from_u8_matches_reference_pipeline(cuda):
    from basicsr4rs_b200.ops.sr_b200 import raw
    rng = np.random.RandomState(0)
    imgs = [rng.randint(0, 256, size=(37 + 5 * i, 53 + 3 * i, 3), dtype=np.uint8) for i in range(4)]
    dev = [torch.from_numpy(a).to(cuda) for a in imgs]
    p = 24
    rows = [(i % 4, (7 * i) % (imgs[i % 4].shape[0] - p + 1), (11 * i) % (imgs[i % 4].shape[1] - p + 1), i % 8)
            for i in range(16)]
    out = raw.patch_from_u8([dev[i] for i, _, _, _ in rows], [r[1] for r in rows], [r[2] for r in rows],
                            [r[3] for r in rows], p, p)
    assert out.shape == (16, 3, p, p)
    for k, (i, top, left, flags) in enumerate(rows):
        assert torch.equal(out[k].cpu(), _ref_patch(imgs[i], top, left, p, flags)), (k, flags)


def test_gpu_paired_patches_follow_the_reference_distribution(cuda):
    """GT crop = scale x LR crop with identical flags; the plan uses the reference's value ranges."""
    from basicsr4rs_b200.utils.data_gpu import GpuPairedPatches
    rng = np.random.RandomState(1)
    lq = [rng.randint(0, 256, size=(40, 56, 3), dtype=np.uint8) for _ in range(3)]
    gt = [np.repeat(np.repeat(a, 4, 0), 4, 1) for a in lq]  # GT = nearest x4 of LR: every GT patch is then checkable
    ds = GpuPairedPatches(lq, gt, scale=4, gt_size=64, device=cuda, seed=3)
    plan = ds.plan(32)
    assert {r[3] for r in plan} > {0} and all(0 <= r[1] <= 40 - 16 and 0 <= r[2] <= 56 - 16 for r in plan)
    batch = ds.gather(plan)
    assert batch['lq'].shape == (32, 3, 16, 16) and batch['gt'].shape == (32, 3, 64, 64)
    up = batch['lq'].repeat_interleave(4, 2).repeat_interleave(4, 3)
    assert torch.equal(up, batch['gt'])
    for k, (i, top, left, flags) in enumerate(plan[:8]):
        assert torch.equal(batch['lq'][k].cpu(), _ref_patch(lq[i], top, left, 16, flags))
    with pytest.raises(ValueError):
        GpuPairedPatches(lq, [g[:-1] for g in gt], scale=4, gt_size=64, device=cuda)


def test_tensor2img_u8_matches_reference(cuda):
    """clamp -> normalise -> x255 -> numpy round (half to even) -> uint8 -> BGR HWC (img_util.py:40-96)."""
    from basicsr4rs_b200.ops.sr_b200 import raw
    g = torch.Generator().manual_seed(0)
    x = torch.rand((3, 33, 47), generator=g) * 1.4 - 0.2
    x[0, 0, :8] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, 0.0, 1.0, -3.0, 7.0])  # ties and clamps
    got = raw.tensor2img_u8(x.to(cuda)).cpu().numpy()
    t = x.clone().clamp_(0, 1)
    ref = (t.numpy().transpose(1, 2, 0)[:, :, ::-1] * 255.0).round().astype(np.uint8)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    got = raw.tensor2img_u8(x.to(cuda), -0.2, 1.2, rgb2bgr=False).cpu().numpy()
    ref = (((x.clamp(-0.2, 1.2) + 0.2) / 1.4).numpy().transpose(1, 2, 0) * 255.0).round().astype(np.uint8)
    assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1  # (the division is a reciprocal multiply here)


@pytest.mark.parametrize('world', [1, 3])
def test_tiled_scene_equals_whole_image_on_the_real_edsr(cuda, world):
    """TiledUpscaler (uint8 in, blended tiles, uint8 out; bands per rank) vs ONE forward of the whole scene through
    the same EDSR: with guard >= the receptive-field radius the two agree to rounding (<= 1 uint8 level)."""
    from basicsr4rs_b200.archs import build_network
    from basicsr4rs_b200.ops.sr_b200 import raw
    from basicsr4rs_b200.utils.tiling import TiledUpscaler
    torch.manual_seed(0)
    net = build_network(dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=64, num_block=2, upscale=4,
                             res_scale=1.0)).to(cuda).eval()
    rng = np.random.RandomState(5)
    base = rng.randint(0, 256, size=(12, 17, 3)).astype(np.float32)
    scene = torch.from_numpy(np.clip(np.kron(base, np.ones((8, 8, 1), dtype=np.float32)) +
                                     rng.randn(96, 136, 3) * 6, 0, 255).astype(np.uint8))
    with torch.no_grad():
        lr = raw.patch_from_u8([scene.to(cuda)], [0], [0], [0], 96, 136)
        whole = raw.tensor2img_u8(net(lr)[0].contiguous()).cpu()
    up = TiledUpscaler(net, scale=4, tile=48, overlap=24, guard=8)
    parts = []
    for rank in range(world):
        band, (y0, y1), n, h2d, d2h = up.upscale(scene.pin_memory(), rank=rank, world=world)
        assert band.shape == ((y1 - y0) * 4, 136 * 4, 3) and n > 0 and h2d > 0 and d2h == band.numel()
        parts.append(band.clone())
    tiled = torch.cat(parts, 0)
    assert tiled.shape == whole.shape
    diff = (tiled.int() - whole.int()).abs()
    assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 0.02, (diff.max().item(),
                                                                               (diff > 0).float().mean().item())
    # crop seams (no blending) also reproduce the scene away from rounding
    crop = TiledUpscaler(net, scale=4, tile=48, overlap=0)
    band, _, _, _, _ = crop.upscale(scene.pin_memory())
    assert band.shape == whole.shape  # (seams differ there: no halo)


def test_tiled_swinir_scene_runs_with_window_padding(cuda):
    from basicsr4rs_b200.archs import build_network
    from basicsr4rs_b200.utils.tiling import TiledUpscaler
    torch.manual_seed(0)
    net = build_network(dict(type='SwinIR', upscale=2, in_chans=3, img_size=32, window_size=8, img_range=1.,
                             depths=[2], embed_dim=60, num_heads=[6], mlp_ratio=2, upsampler='pixelshuffle',
                             resi_connection='1conv')).to(cuda).eval()
    scene = torch.randint(0, 256, (75, 61, 3), dtype=torch.uint8).pin_memory()
    up = TiledUpscaler(net, scale=2, tile=40, overlap=16, guard=4, multiple=8)
    band, (y0, y1), n, _, _ = up.upscale(scene)
    assert band.shape == (150, 122, 3) and (y0, y1) == (0, 75) and n == 6
    assert band.float().std().item() > 1.0
    assert sr_oracle is not None
