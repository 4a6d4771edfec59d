#!/usr/bin/env python
"""BASELINE.json config 5 end to end: tiled x4 inference of a synthetic remote-sensing-scale scene, sharded across the
GPUs of one box in row bands (replicas only, no collective on the data path).

    python tools/bench_infer.py --arch edsr_l --scene 2048 --tile 1024
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_infer.py --arch swinir ...

The scene is a uint8 HWC image in PINNED HOST memory; every rank uploads the tiles of its band, super-resolves them
through basicsr4rs_b200.utils.tiling.TiledUpscaler (uint8 -> fp32 entry kernel, the network, blended seams, tensor2img
exit kernel) and reads its uint8 band back to pinned host memory -- H2D and D2H are inside the timed region.  Prints one
JSON line: output MPix/s of the whole job, wall-clock of the slowest rank between two barriers (device-synchronised).
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200.archs import build_network  # noqa: E402
from basicsr4rs_b200.utils import tiling  # noqa: E402
from bench import EDSR_L, SWINIR  # noqa: E402

GFLOP_PER_LR_PIXEL = {'edsr_l': 100.5e-3, 'swinir': 26.15e-3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--arch', default='edsr_l', choices=['edsr_l', 'swinir'])
    ap.add_argument('--scene', type=int, default=2048, help='LR scene is scene x scene pixels')
    ap.add_argument('--tile', type=int, default=1024)
    ap.add_argument('--overlap', type=int, default=0, help='LR pixels shared by neighbouring tiles (blended)')
    ap.add_argument('--guard', type=int, default=0)
    ap.add_argument('--reps', type=int, default=2)
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    net = build_network(dict(EDSR_L if args.arch == 'edsr_l' else SWINIR)).to(dev).eval()
    g = torch.Generator().manual_seed(1234)
    scene = torch.randint(0, 256, (args.scene, args.scene, 3), generator=g, dtype=torch.uint8).pin_memory()
    up = tiling.TiledUpscaler(net, scale=4, tile=args.tile, overlap=args.overlap, guard=args.guard,
                              multiple=8 if args.arch == 'swinir' else 1)
    y0, y1 = tiling.band_rows(args.scene, rank, world, 8)
    out = torch.empty(((y1 - y0) * 4, args.scene * 4, 3), dtype=torch.uint8).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    _, _, n_tiles, h2d, d2h = up.upscale(scene, rank, world, out=out)  # warm-up (allocator, tensor maps, clocks)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        up.upscale(scene, rank, world, out=out)
    barrier()
    ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.reps], dtype=torch.float64, device=dev)
    tiles = torch.tensor([n_tiles], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tiles)
    if rank == 0:
        out_pix = (4 * args.scene)**2
        print(json.dumps({'config': f'{args.arch} x4 tiled inference, uint8 scene in pinned host memory -> uint8 scene in '
                                    'pinned host memory', 'n_gpus': world, 'scene_lr': args.scene, 'tile_lr': args.tile,
                          'overlap_lr': args.overlap, 'tiles_computed': int(tiles.item()), 'ms': ms.item(),
                          'out_mpix_per_s': out_pix / ms.item() / 1e3,
                          'model_tflops': GFLOP_PER_LR_PIXEL[args.arch] * args.scene**2 / ms.item() / world,
                          'h2d_bytes_rank0': h2d, 'd2h_bytes_rank0': d2h,
                          'scaling': 'strong (fixed scene), row bands per rank, replicas only'}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
