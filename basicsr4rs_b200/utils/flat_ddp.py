"""Data parallelism on ONE flat gradient buffer (opt-in replacement for the DDP wrap of ``BaseModel.model_to_device``,
basicsr/models/base_model.py:94-99).

Why: under ``torch.nn.parallel.DistributedDataParallel`` every parameter's gradient is COPIED into its bucket as it
arrives (``mark_variable_ready_dense``: one ``mul_out`` per parameter and step -- the CUDA-graph replay hands autograd
its static output tensors, which never alias DDP's buckets).  For EDSR-L that is 126 launches of ~5 us per step: the
CUPTI timeline of a 2-GPU step (tools/prof_ddp_step.py, profiles/r02_ddp2_fp32.txt) shows 0.63 ms of them against
0.62 ms of NCCL kernels -- most of the 6 % that weak scaling loses between 1 and 2 GPUs and never loses again.

Here the layers' one-launch gradient finalize (``raw.finalize_grads``) writes every weight / bias gradient straight into
its slice of a flat fp32 buffer (``FlatGrads``, registered in ``raw.GRAD_SINK`` BEFORE the CUDA graphs are captured, so
the captures record the final addresses); autograd adopts the returned aliases as ``param.grad`` without a copy, and
``FlatDDP`` all-reduces the buffer itself -- bucket by bucket as soon as all gradients of a bucket have arrived, waited
for at the end of ``backward()`` exactly like DDP.  The finalize launches pre-divide by the world size (their ``alpha``),
so the collective is a plain SUM with no division kernel: ``ReduceOp.AVG`` is a pre-multiplied sum, which keeps NCCL off
its in-switch NVLS reduction (measured on 8 GPUs: five ``AllReduce_Sum_f32_RING_LL`` kernels of 760 us per step).  Gradients
that reach autograd by other routes (LayerNorm, bias tables, the image-exit conv) are folded in with one
``torch._foreach_copy_`` per bucket.  One exchange step, no custom collective: NCCL over NVLink does the reduction.
"""
import weakref

import torch
import torch.distributed as dist
from torch import nn

from ..ops.sr_b200 import raw


class FlatGrads:
    """One fp32 buffer holding the gradient of every trainable parameter of ``module``, in registration order."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError('FlatGrads: the module has no trainable parameters')
        dev = self.params[0].device
        self.offsets, total = [], 0
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32:
                raise ValueError('FlatGrads: fp32 parameters on one device')
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4  # 16-byte aligned slices
        self.buf = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.views = [self.buf[o:o + p.numel()].view(p.shape) for o, p in zip(self.offsets, self.params)]
        self.total = total
        # gradients written through the sink are pre-divided by the world size known NOW (the CUDA graphs captured after
        # this point bake the factor into their finalize launches): the all-reduce is then a plain SUM
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.prescale = 1.0 / self.world
        self.sunk = dev.type == 'cuda'
        if self.sunk:
            for p, v in zip(self.params, self.views):
                raw.GRAD_SINK[p.data_ptr()] = (v, self.prescale)

    def release(self):
        for p, v in zip(self.params, self.views):
            hit = raw.GRAD_SINK.get(p.data_ptr())
            if hit is not None and hit[0] is v:
                del raw.GRAD_SINK[p.data_ptr()]


def flat_grads_of(module):
    """The module's FlatGrads (created on first use; the archs create it on ``.to(device)`` when built with
    ``flat_grads=True``, i.e. before their CUDA graphs are captured)."""
    fg = module.__dict__.get('_srb200_flat_grads')
    if fg is None or fg.params[0].data_ptr() != next(p for p in module.parameters() if p.requires_grad).data_ptr():
        if fg is not None:
            fg.release()
        fg = module.__dict__['_srb200_flat_grads'] = FlatGrads(module)
        # the sink is keyed by storage address: drop the entries when the module dies, or a later, unrelated parameter
        # that the allocator places at the same address would have its gradient redirected (and pre-divided)
        weakref.finalize(module, _unsink, [p.data_ptr() for p in fg.params], fg.buf.data_ptr())
    return fg


def _unsink(ptrs, buf_ptr):
    for ptr in ptrs:
        hit = raw.GRAD_SINK.get(ptr)
        if hit is not None and hit[0].untyped_storage().data_ptr() == buf_ptr:
            del raw.GRAD_SINK[ptr]


class FlatDDP(nn.Module):
    """``FlatDDP(net)`` where the reference writes ``DistributedDataParallel(net, device_ids=[...])``.

    ``bucket_mb``: all-reduce granularity (default 40 MB, 5 buckets for EDSR-L's 172 MB of gradients); ``last_mb``: size of
    the bucket that is reduced last (the first parameters), whose all-reduce nothing overlaps."""

    def __init__(self, module, bucket_mb=40, process_group=None, last_mb=8):
        super().__init__()
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.flat = flat_grads_of(module)
        fg = self.flat
        # parameters start out identical on every rank (DDP broadcasts rank 0's state as well)
        if self.world > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, 0, group=process_group)
        # Buckets = contiguous index ranges, cut from the FRONT of the registration order: gradients arrive roughly back to
        # front, so the bucket holding the first parameters is reduced last, with nothing left to overlap it -- it is kept
        # small (``last_mb``); the others hold ``bucket_mb``.
        self.bucket_of, self.buckets = {}, []  # param index -> bucket, bucket = [lo, hi, param indices]
        cur, cap = None, int(last_mb * 2**20 / 4)
        for i in range(len(fg.params)):
            lo = fg.offsets[i]
            hi = fg.offsets[i] + (fg.params[i].numel() + 3) // 4 * 4
            if cur is None or (hi - cur[0]) > cap:
                if cur is not None:
                    cap = int(bucket_mb * 2**20 / 4)
                cur = [lo, hi, []]
                self.buckets.append(cur)
            cur[1] = hi
            cur[2].append(i)
            self.bucket_of[i] = len(self.buckets) - 1
        self._pending = None
        self._works = []
        if fg.sunk and fg.world != self.world:
            raise RuntimeError('FlatDDP: the flat gradient buffer was created for another world size -- build the network '
                               'after init_process_group')
        for i, p in enumerate(fg.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # ------------------------------------------------------------------ backward side
    def _make_hook(self, i):
        def hook(param):
            if self._pending is None:
                self._pending = [len(b[2]) for b in self.buckets]
                self._launched = [False] * len(self.buckets)
                torch.autograd.Variable._execution_engine.queue_callback(self._finish)
            b = self.bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._reduce(b)
        return hook

    def _reduce(self, b):
        fg = self.flat
        lo, hi, idx = self.buckets[b]
        self._launched[b] = True
        # gradients that did not come through the sink (LayerNorm, bias tables, the image-exit conv; everything on a CPU
        # module): fold them into the flat buffer pre-divided like the sunk ones, re-point .grad at the slice
        src, dst = [], []
        for i in idx:
            p = fg.params[i]
            if p.grad is not None and p.grad.data_ptr() != fg.views[i].data_ptr():
                src.append(p.grad.detach().reshape(fg.views[i].shape))
                dst.append(fg.views[i])
        if src:
            torch._foreach_copy_(dst, src)
            if self.world > 1:
                torch._foreach_mul_(dst, 1.0 / self.world)
            for i in idx:
                p = fg.params[i]
                if p.grad is not None and p.grad.data_ptr() != fg.views[i].data_ptr():
                    p.grad = fg.views[i].detach()
        if self.world > 1:
            self._works.append(dist.all_reduce(fg.buf[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _finish(self):
        """End of the backward pass (still inside ``backward()``): reduce what is left, make the current stream wait."""
        for b in range(len(self.buckets)):
            if not self._launched[b]:  # a bucket with parameters that received no gradient this pass
                self._reduce(b)
        for w in self._works:
            w.wait()
        self._works = []
        self._pending = None
