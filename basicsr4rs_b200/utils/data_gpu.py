"""GPU-side input pipeline for paired SR training patches.

The reference builds every LR / GT patch on the host: ``PairedImageDataset.__getitem__``
(basicsr/data/paired_image_dataset.py:63-103) decodes to float32 HWC BGR, ``paired_random_crop``
(data/transforms.py:28-96) crops, ``augment`` (:166-225) flips / transposes with cv2 and ``img2tensor``
(utils/img_util.py:9-37) makes CHW RGB tensors, then ``CUDAPrefetcher`` (data/prefetch_dataloader.py:82-122) copies the
float batch (16 x (3x48x48 + 3x192x192) x 4 B = 7.5 MB per step).  At >= 10k patches/s on 8 GPUs that host pipeline is
the limiter.  Here the decoded images stay uint8 and are uploaded ONCE (or per epoch shard); each step draws crop
origins and flip / rotation flags with the reference's own distribution and one kernel per tensor
(``srb200_patch_from_u8``) produces the fp32 NCHW RGB batch the arch's ``forward`` takes -- 16 patches = 2 launches,
no host pixels touched.
"""
import random

import torch

from ..ops.sr_b200 import raw


class GpuPairedPatches:
    """Holds uint8 HWC (BGR, as cv2 decodes) LR / GT image pairs on the device and samples training batches.

    ``draw(batch)`` mirrors ``PairedImageDataset`` in training phase: a uniformly random image, a uniformly random LR
    crop origin with the GT origin at ``scale`` times it (transforms.py:74-90), then ``hflip`` / ``vflip`` / ``rot90``
    each with probability 1/2 when enabled (transforms.py:190-192), BGR->RGB, HWC->CHW, /255."""

    def __init__(self, lq_images, gt_images, scale, gt_size, use_hflip=True, use_rot=True, device='cuda', seed=None):
        if len(lq_images) != len(gt_images) or not lq_images:
            raise ValueError('need as many LR as GT images')
        self.scale, self.gt_size, self.lq_size = scale, gt_size, gt_size // scale
        self.use_hflip, self.use_rot = use_hflip, use_rot
        self.rng = random.Random(seed)
        self.lq, self.gt = [], []
        for lq, gt in zip(lq_images, gt_images):
            lq, gt = torch.as_tensor(lq), torch.as_tensor(gt)
            if lq.dtype != torch.uint8 or gt.dtype != torch.uint8 or lq.dim() != 3:
                raise ValueError('images must be uint8 HWC')
            if gt.shape[0] != lq.shape[0] * scale or gt.shape[1] != lq.shape[1] * scale:
                raise ValueError(f'Scale mismatches. GT {tuple(gt.shape[:2])} is not {scale}x LQ {tuple(lq.shape[:2])}')
            if lq.shape[0] < self.lq_size or lq.shape[1] < self.lq_size:
                raise ValueError(f'LQ {tuple(lq.shape[:2])} is smaller than patch size {self.lq_size}')
            self.lq.append(lq.to(device).contiguous())
            self.gt.append(gt.to(device).contiguous())

    def plan(self, batch):
        """Host side of a draw: [(image index, top, left, flags)] with the reference's distributions."""
        rows = []
        for _ in range(batch):
            i = self.rng.randrange(len(self.lq))
            h, w = self.lq[i].shape[:2]
            top = self.rng.randint(0, h - self.lq_size)
            left = self.rng.randint(0, w - self.lq_size)
            hflip = self.use_hflip and self.rng.random() < 0.5
            vflip = self.use_rot and self.rng.random() < 0.5
            rot90 = self.use_rot and self.rng.random() < 0.5
            rows.append((i, top, left, int(hflip) | (int(vflip) << 1) | (int(rot90) << 2)))
        return rows

    def gather(self, rows):
        """Device side: the fp32 [B,3,p,p] LR batch and [B,3,s*p,s*p] GT batch of a plan (two kernel launches)."""
        idx, tops, lefts, flags = zip(*rows)
        s = self.scale
        lq = raw.patch_from_u8([self.lq[i] for i in idx], tops, lefts, flags, self.lq_size, self.lq_size)
        gt = raw.patch_from_u8([self.gt[i] for i in idx], [t * s for t in tops], [l * s for l in lefts], flags,
                               self.gt_size, self.gt_size)
        return {'lq': lq, 'gt': gt}

    def draw(self, batch):
        return self.gather(self.plan(batch))
