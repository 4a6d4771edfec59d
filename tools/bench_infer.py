#!/usr/bin/env python
"""BASELINE.json config 5: tiled x4 inference of a synthetic remote-sensing-scale scene, LR tiles sharded across the
GPUs of one box (replicas only, no collective on the data path).

    python tools/bench_infer.py --arch edsr_l --scene 4096 --tile 1024
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_infer.py --arch swinir ...

Prints one JSON line: output MPix/s of the whole job (all ranks), device-timed, max over ranks.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200.archs import build_network  # noqa: E402
from basicsr4rs_b200.utils import tiling  # noqa: E402
from tools.bench_all import INFER  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--arch', default='edsr_l', choices=['edsr_l', 'swinir'])
    ap.add_argument('--scene', type=int, default=4096, help='LR scene is scene x scene pixels')
    ap.add_argument('--tile', type=int, default=1024)
    ap.add_argument('--reps', type=int, default=2)
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    opt, _, gflop_px = INFER['edsr_l_infer_1024' if args.arch == 'edsr_l' else 'swinir_infer_1024']
    torch.manual_seed(0)
    net = build_network(dict(opt)).to(dev).eval()
    tiles = tiling.shard(tiling.tile_grid(args.scene, args.scene, args.tile), rank, world)
    g = torch.Generator().manual_seed(1234)
    lr_tile = torch.rand((1, 3, args.tile, args.tile), generator=g).to(dev)  # every tile: same synthetic content

    def run():
        n = 0
        with torch.no_grad():
            for _ in tiles:
                net(lr_tile)
                n += 1
        return n

    run() if len(tiles) <= 2 else (net(lr_tile), net(lr_tile))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    if rank == 0:
        n_tiles = len(tiling.tile_grid(args.scene, args.scene, args.tile))
        out_pix = n_tiles * (4 * args.tile)**2
        print(json.dumps({'config': f'{args.arch} x4 tiled inference', 'n_gpus': world, 'scene_lr': args.scene,
                          'tile_lr': args.tile, 'tiles': n_tiles, 'tiles_per_gpu': len(tiles), 'ms': ms.item(),
                          'out_mpix_per_s': out_pix / ms.item() / 1e3,
                          'model_tflops': gflop_px * args.tile * args.tile * n_tiles / ms.item() / world,
                          'scaling': 'strong (fixed scene), replicas only'}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
