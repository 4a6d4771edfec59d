"""SwinIR use_checkpoint=True vs False on one GPU: same outputs, same gradients?  (Round 1 ended with the fix for the
double ``ctx.saved_tensors`` read unverified on a GPU -- run this first, then turn it into a parity test.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200.archs import build_network  # noqa: E402

dev = torch.device('cuda:0')
kw = dict(type='SwinIR', upscale=2, in_chans=3, img_size=16, window_size=8, img_range=1., depths=[2, 2], embed_dim=60,
          num_heads=[6, 6], mlp_ratio=2, upsampler='pixelshuffle', resi_connection='1conv', drop_path_rate=0.0)
torch.manual_seed(0)
a = build_network(dict(kw)).to(dev).train()
b = build_network(dict(kw, use_checkpoint=True)).to(dev).train()
b.load_state_dict(a.state_dict())
x = torch.rand((2, 3, 16, 24), generator=torch.Generator().manual_seed(3)).to(dev)
for net in (a, b):
    net(x).square().mean().backward()
worst = 0.0
for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
    if p.grad is not None:
        worst = max(worst, ((p.grad - q.grad).norm() / (p.grad.norm() + 1e-12)).item())
print('use_checkpoint: worst gradient rel-L2 difference', worst, '; outputs equal:', torch.equal(a(x), b(x)))
