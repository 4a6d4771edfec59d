"""world_size-2 ``gloo`` test (CPU) of the multi-GPU host logic bench.py uses: per-rank synthetic shards differ,
the max-over-ranks step time reduction, DDP gradient averaging through custom autograd Functions that
return parameter gradients the way the srb200 Functions do, and rank-0-only reference arm."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _ScaledLinear(torch.autograd.Function):
    """Stand-in with the contract of ops/sr_b200 Functions: explicit backward returning (gx, gw, gb)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return x @ w.t() + b

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        return g @ w, g.t() @ x, g.sum(0)


class _Net(torch.nn.Module):

    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(8, 4)

    def forward(self, x):
        return _ScaledLinear.apply(x, self.lin.weight, self.lin.bias)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    lq, gt = bench.synthetic_batch(rank, batch=2, lr=8)
    # shards must differ between ranks and be reproducible
    lq2, _ = bench.synthetic_batch(rank, batch=2, lr=8)
    assert torch.equal(lq, lq2)
    gathered = [torch.zeros_like(lq) for _ in range(world)]
    dist.all_gather(gathered, lq)
    assert not torch.equal(gathered[0], gathered[1])
    # max-over-ranks timing
    ms = bench.max_over_ranks(10.0 + rank, torch.device('cpu'), world)
    assert ms == 10.0 + world - 1
    # DDP averages the gradients returned by a custom Function
    torch.manual_seed(0)
    net = torch.nn.parallel.DistributedDataParallel(_Net())
    x = torch.full((3, 8), float(rank + 1))
    net(x).sum().backward()
    g = net.module.lin.weight.grad.clone()
    want = torch.full((4, 8), 3.0 * sum(r + 1 for r in range(world)) / world)
    assert torch.allclose(g, want), (g, want)
    # FlatDDP (utils/flat_ddp.py): same averaging on ONE flat gradient buffer, .grad aliases its slice, two steps
    from basicsr4rs_b200.utils.flat_ddp import FlatDDP
    torch.manual_seed(1 + rank)  # different initial weights per rank: the wrap must broadcast rank 0's
    fnet = FlatDDP(_Net(), bucket_mb=1)
    w0 = [torch.zeros_like(fnet.module.lin.weight) for _ in range(world)]
    dist.all_gather(w0, fnet.module.lin.weight.detach())
    assert torch.equal(w0[0], w0[1])
    for step in range(2):
        for p in fnet.parameters():
            p.grad = None
        fnet(x * (step + 1)).sum().backward()
        gw, gb = fnet.module.lin.weight.grad, fnet.module.lin.bias.grad
        assert torch.allclose(gw, want * (step + 1)), (gw, want)
        assert torch.allclose(gb, torch.full((4,), 3.0))
        assert gw.data_ptr() == fnet.flat.views[0].data_ptr() and gb.data_ptr() == fnet.flat.views[1].data_ptr()
    if rank == 0:
        torch.save({'ok': True}, out)
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    out = str(tmp_path / 'ok.pt')
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    assert torch.load(out)['ok']


def test_reference_arm_only_rank0_prints():
    """``bench.py --impl reference`` under torchrun: ranks != 0 exit 0 without work or output."""
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''


def _lazy_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from collections import OrderedDict
    from basicsr4rs_b200.utils.train_hooks import install_lazy_loss_log

    class M:  # the two attributes BaseModel.reduce_loss_dict reads (base_model.py:385,393-394)
        opt = {'dist': True, 'rank': rank, 'world_size': world}

    m = install_lazy_loss_log(M())
    logs = [m.reduce_loss_dict(OrderedDict(l_pix=torch.tensor(float(rank + 1 + it)), l_percep=torch.ones(3) * it))
            for it in range(5)]  # five iterations, nothing read: no collective was issued
    last = logs[-1]
    q.put((rank, f"{last['l_pix']:.4e}", float(last['l_percep']), list(last.keys())))
    dist.destroy_process_group()


def test_lazy_loss_log_matches_reduce_loss_dict_world2():
    """install_lazy_loss_log: same keys, the value read at print time is the cross-rank mean of base_model.py:376-401."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_lazy_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    for rank, s, percep, keys in got:
        assert keys == ['l_pix', 'l_percep']
        assert s == f'{(5 + 6) / 2:.4e}' and percep == 4.0  # iteration 4: ranks hold 5 and 6 -> mean 5.5
