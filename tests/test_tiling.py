"""Host logic of the tiled multi-GPU inference path (BASELINE.json config 5): tile grids, rank sharding,
window padding and seam handling -- CPU, with a plain bicubic-free stand-in network (nearest x``s`` upscale),
plus a world_size-2 gloo run in which two ranks fill disjoint tiles of the same scene."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from basicsr4rs_b200.utils import tiling  # noqa: E402


class _Nearest(torch.nn.Module):
    def __init__(self, s):
        super().__init__()
        self.s = s

    def forward(self, x):
        return x.repeat_interleave(self.s, 2).repeat_interleave(self.s, 3)


@pytest.mark.parametrize('h,w,tile,overlap', [(1024, 1024, 256, 0), (100, 70, 32, 8), (48, 48, 64, 0), (65, 33, 32, 31),
                                              (1, 1, 8, 0)])
def test_tile_grid_covers_image(h, w, tile, overlap):
    tiles = tiling.tile_grid(h, w, tile, overlap)
    cover = torch.zeros((h, w), dtype=torch.int32)
    for (y, x, th, tw) in tiles:
        assert 0 < th <= tile and 0 < tw <= tile and y + th <= h and x + tw <= w
        cover[y:y + th, x:x + tw] += 1
    assert (cover > 0).all()
    if overlap == 0 and h % tile == 0 and w % tile == 0:
        assert (cover == 1).all() and len(tiles) == (h // tile) * (w // tile)


def test_tile_grid_edge_cases():
    assert tiling.tile_grid(0, 10, 8) == []
    with pytest.raises(ValueError):
        tiling.tile_grid(10, 10, 8, overlap=8)
    with pytest.raises(ValueError):
        tiling.shard([1, 2, 3], 2, 2)


def test_shards_partition_the_work():
    items = list(range(37))
    for world in (1, 2, 4, 8):
        parts = [tiling.shard(items, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == items
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_pad_to_multiple_matches_reference_reflect():
    """Same flip-concatenate padding as SwinIRModel.test (swinir_model.py:16-24), window_size 8."""
    x = torch.arange(2 * 3 * 13 * 21, dtype=torch.float32).view(2, 3, 13, 21)
    y, (h, w) = tiling.pad_to_multiple(x, 8)
    assert (h, w) == (13, 21) and y.shape[-2:] == (16, 24)
    ref = torch.cat([x, torch.flip(x, [2])], 2)[:, :, :16, :]
    ref = torch.cat([ref, torch.flip(ref, [3])], 3)[:, :, :, :24]
    assert torch.equal(y, ref)
    z, _ = tiling.pad_to_multiple(y, 8)
    assert z is y


@pytest.mark.parametrize('overlap', [0, 8])
def test_tiled_forward_equals_whole_image(overlap):
    net = _Nearest(4)
    img = torch.rand((1, 3, 90, 61))
    whole = net(img)
    out, n = tiling.tiled_forward(net, img, tile=32, scale=4, overlap=overlap, multiple=8)
    assert n == len(tiling.tile_grid(90, 61, 32, overlap))
    assert torch.equal(out, whole)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    net = _Nearest(2)
    img = torch.arange(3 * 48 * 64, dtype=torch.float32).view(1, 3, 48, 64) + 1.0
    out, n = tiling.tiled_forward(net, img, tile=16, scale=2, rank=rank, world=world)
    # replicas only: the test (not the product path) sums the disjoint partial scenes to check the partition
    filled = (out != 0).to(torch.int32)
    dist.all_reduce(out)
    dist.all_reduce(filled)
    counts = torch.tensor([n])
    dist.all_reduce(counts)
    ok = torch.equal(out, net(img)) and int(filled.max()) == 1 and int(filled.min()) == 1 and \
        int(counts) == len(tiling.tile_grid(48, 64, 16)) == 12
    if rank == 0:
        q.put(bool(ok))
    dist.destroy_process_group()


def test_two_ranks_fill_disjoint_tiles():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_tiled_test_hook_matches_the_reference_test_contract():
    """utils/train_hooks.install_tiled_test on a stand-in model object (lq / opt / net_g [/ net_g_ema], the attributes
    SRModel.test and SwinIRModel.test use): a pointwise stand-in network makes tiled == whole-image exact, so the hook's
    plumbing is what is tested -- window padding and cropping, EMA preference, train-mode restore, batch handling."""
    import types
    import torch
    from basicsr4rs_b200.utils.train_hooks import install_tiled_test

    class Up(torch.nn.Module):  # x2 nearest upsample of 2*x + 1: pointwise, so tiling cannot change any pixel
        def __init__(self, multiple):
            super().__init__()
            self.multiple, self.calls, self.w = multiple, 0, torch.nn.Parameter(torch.ones(1))

        def forward(self, x):
            assert x.shape[-2] % self.multiple == 0 and x.shape[-1] % self.multiple == 0, 'caller must pad'
            self.calls += 1
            return torch.repeat_interleave(torch.repeat_interleave(2 * x + self.w, 2, dim=2), 2, dim=3)

    lq = torch.rand((2, 3, 37, 50))
    want = torch.repeat_interleave(torch.repeat_interleave(2 * lq + 1, 2, dim=2), 2, dim=3)
    # SwinIRModel-like: window_size 8, no EMA copy, net_g in train mode
    m = types.SimpleNamespace(opt={'scale': 2, 'network_g': {'window_size': 8}}, lq=lq, net_g=Up(8).train())
    install_tiled_test(m, tile=16, overlap=4)
    m.test()
    assert m.output.shape == want.shape and torch.equal(m.output, want)
    assert m.net_g.training and m.net_g.calls > 2          # tiled (several forwards), train mode restored
    # one forward when the image fits a tile -- the reference's own behaviour, padding and cropping included
    m2 = types.SimpleNamespace(opt={'scale': 2, 'network_g': {'window_size': 8}}, lq=lq, net_g=Up(8).train())
    install_tiled_test(m2, tile=64, overlap=4)
    m2.test()
    assert torch.equal(m2.output, want) and m2.net_g.calls == 1
    # SRModel-like with an EMA copy: net_g_ema is the one evaluated, net_g is left alone
    m3 = types.SimpleNamespace(opt={'scale': 2, 'network_g': {}}, lq=lq[:1], net_g=Up(1).train(), net_g_ema=Up(1))
    install_tiled_test(m3, tile=20, overlap=6)
    m3.test()
    assert torch.equal(m3.output, want[:1]) and m3.net_g.calls == 0 and m3.net_g_ema.calls > 1 and m3.net_g.training
