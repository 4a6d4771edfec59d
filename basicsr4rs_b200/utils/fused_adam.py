"""One-launch Adam (step overhead on the caller's side) -- opt-in, a drop-in for the
``torch.optim.Adam`` that ``BaseModel.get_optimizer`` builds from ``optim_g: {type: Adam, lr, betas, weight_decay}``
(basicsr/models/base_model.py:107-124; stepped by ``SRModel.optimize_parameters``, sr_model.py:113).

``torch.optim.Adam(fused=True)`` packs ~36 tensors per launch into kernel arguments: 61 launches / 1.1 ms per step for
RCAN's 1660 parameters, 89 / 0.7 ms for SwinIR, 6 / 0.24 ms for EDSR-L.  :class:`FusedAdam` walks a device table of
(param, grad, exp_avg, exp_avg_sq) pointers in ONE ``srb200_multi_adam`` launch with the same arithmetic; the table is
rebuilt only when a pointer changes (under CUDA-graph replay and with ``flat_grads`` the gradients keep their
addresses).  Same constructor keywords, same ``param_groups`` (LR schedulers work unchanged) and the same
``state_dict`` layout (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter), so checkpoints interchange with
``torch.optim.Adam``.  Not supported: ``amsgrad``, ``maximize``, sparse gradients, closures that re-evaluate.
"""
import ctypes

import numpy as np
import torch

from .. import _lib as L

_ITEM = np.dtype([('p', '<i8'), ('g', '<i8'), ('m', '<i8'), ('v', '<i8'), ('n', '<i8'), ('chunk_begin', '<i8')])


class FusedAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, **unused):
        if amsgrad:
            raise NotImplementedError('FusedAdam: amsgrad is not implemented (the reference recipes do not use it)')
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._tables = {}  # group index -> (pointer key, device table, chunks, n)
        self._steps = {}   # group index -> update count

    def _init_group(self, gi, group):
        """exp_avg / exp_avg_sq of a group live in two flat buffers (one allocation, one fill each)."""
        missing = [p for p in group['params'] if p.requires_grad and 'exp_avg' not in self.state[p]]
        if not missing:
            return
        total = sum((p.numel() + 3) // 4 * 4 for p in missing)
        dev = missing[0].device
        flat_m = torch.zeros((total,), dtype=torch.float32, device=dev)
        flat_v = torch.zeros((total,), dtype=torch.float32, device=dev)
        o = 0
        for p in missing:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError('FusedAdam needs contiguous fp32 CUDA parameters')
            st = self.state[p]
            st['step'] = torch.tensor(0.0)
            st['exp_avg'] = flat_m[o:o + p.numel()].view(p.shape)
            st['exp_avg_sq'] = flat_v[o:o + p.numel()].view(p.shape)
            o += (p.numel() + 3) // 4 * 4

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        for gi, group in enumerate(self.param_groups):
            self._init_group(gi, group)
            live = [p for p in group['params'] if p.grad is not None]
            if not live:
                continue
            step = int(self._steps.get(gi, 0))
            if step == 0:  # resumed from a checkpoint written by torch.optim.Adam / by state_dict() below
                step = int(max(float(self.state[p]['step']) for p in live))
            step += 1
            self._steps[gi] = step
            key = [p.data_ptr() for p in live] + [p.grad.data_ptr() for p in live] + \
                  [self.state[p]['exp_avg'].data_ptr() for p in live[:1]]
            cached = self._tables.get(gi)
            if cached is None or cached[0] != key:
                tab = np.zeros(len(live), dtype=_ITEM)
                chunk = 0
                for i, p in enumerate(live):
                    g = p.grad
                    if g.is_sparse or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device:
                        raise RuntimeError('FusedAdam needs dense contiguous fp32 gradients on the parameter device')
                    st = self.state[p]
                    tab[i] = (p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(),
                              p.numel(), chunk)
                    chunk += (p.numel() + 1023) // 1024
                table = torch.from_numpy(tab.view(np.uint8).reshape(len(tab), -1).copy()).to(live[0].device)
                cached = self._tables[gi] = (key, table, chunk, len(live))
            _, table, chunks, n = cached
            dev = live[0].device
            with torch.cuda.device(dev):
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                b1, b2 = group['betas']
                L.check(lib.srb200_multi_adam(ctypes.c_void_p(table.data_ptr()), n, chunks, float(group['lr']), float(b1),
                                              float(b2), float(group['eps']), float(group['weight_decay']), step, stream),
                        'multi_adam')
            # the kernel wrote the parameters through raw pointers (``_version`` did not move): the bf16 GEMM operands
            # packed from them must be re-derived before their next use
            books = set()
            for p in live:
                p.__dict__.pop('_srb200_pack', None)
                ref = p.__dict__.get('_srb200_book')
                book = ref() if ref is not None else None
                if book is not None and id(book) not in books:
                    books.add(id(book))
                    book.invalidate()
        return loss

    def state_dict(self):
        for gi, group in enumerate(self.param_groups):
            step = float(self._steps.get(gi, 0))
            for p in group['params']:
                if p in self.state and 'step' in self.state[p]:
                    self.state[p]['step'] = torch.tensor(step)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}
        for gi, group in enumerate(self.param_groups):
            steps = [float(self.state[p]['step']) for p in group['params'] if p in self.state and 'step' in self.state[p]]
            self._steps[gi] = int(max(steps)) if steps else 0
