"""Read-only, write-only and copy HBM bandwidth on one GPU (torch kernels, CUDA events, best of 10) -- the
denominators for write-dominated kernels such as SwinIR's qkv GEMM (25 MB in, 75 MB out)."""
import torch
dev = torch.device('cuda:0')
n = 1 << 29  # 512 Mi bf16 = 1 GiB
a = torch.ones(n, dtype=torch.bfloat16, device=dev)
b = torch.empty_like(a)
def best(fn, iters=10):
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts) * 1e-3
gib = n * 2
print(f'write-only  (fill_)      : {gib / best(lambda: b.fill_(1.0)) / 1e9:8.1f} GB/s')
print(f'read-only   (sum)        : {gib / best(lambda: a.sum()) / 1e9:8.1f} GB/s')
print(f'copy        (read+write) : {2 * gib / best(lambda: b.copy_(a)) / 1e9:8.1f} GB/s')
c = torch.empty(3 * n, dtype=torch.bfloat16, device=dev)
print(f'1 read : 3 write (repeat): {4 * gib / best(lambda: torch.cat([a, a, a], out=c)) / 1e9:8.1f} GB/s')
