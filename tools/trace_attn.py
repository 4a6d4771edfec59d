"""clock64 timeline of CTA 0 of the tcgen05 window-attention kernel (srb200_debug_set_attn_trace): per stage, when each
warp role waited and worked.  python tools/trace_attn.py [fwd|bwd] [shift]"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200 import _lib as L
from basicsr4rs_b200.ops.sr_b200 import swin_ops as so

which = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
shift = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device('cuda:0')
B, H, W = 16, 64, 64
qkv = torch.randn((B, H, W, 576), device=dev).to(torch.bfloat16)
go = torch.randn((B, H, W, 192), device=dev).to(torch.bfloat16)
table = torch.randn(225, 6, device=dev) * 0.1
_, stats = so.window_attention_fwd(qkv, table, 6, 8, shift, 30**-0.5, want_stats=True)
run = (lambda: so.window_attention_fwd(qkv, table, 6, 8, shift, 30**-0.5)) if which == 'fwd' else \
    (lambda: so.window_attention_bwd(qkv, go, table, 6, 8, shift, 30**-0.5, stats=stats, use_tc=True))
for _ in range(3):
    run()
buf = torch.zeros(4 * 64 * 8 + 4 * 160, dtype=torch.int64, device=dev)
L.load().srb200_debug_set_attn_trace(ctypes.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
L.load().srb200_debug_set_attn_trace(None)
wall = buf.cpu()[4 * 64 * 8:].view(160, 4)
t = buf.cpu()[:4 * 64 * 8].view(4, 64, 8)
t0 = int(t[t > 0].min())
names = ['producer', 'mma', 'softmax', 'epilogue']
for it in range(16):
    if int(t[:, it].max()) == 0:
        break
    print(f'stage {it:2d}: ' + ' | '.join(f'{names[r]} ' + ' '.join(f'{int(v) - t0:6d}' if v > 0 else '     -' for v in t[r, it, :6])
                                     for r in range(4)))
if which == 'bwd' and int(wall.max()) > 0:
    w0 = int(wall[:, 0][wall[:, 0] > 0].min())
    live = wall[wall[:, 0] > 0]
    rel = lambda col: [(int(v) - w0) / 1e3 for v in live[:, col] if v > 0]
    st, lp, en, last = rel(0), rel(1), rel(2), rel(3)
    print(f'wall clock (us after the first CTA start), {len(st)} CTAs: start max {max(st):.1f} | main loop done min {min(lp):.1f} '
          f'median {sorted(lp)[len(lp) // 2]:.1f} max {max(lp):.1f} | end (ordinary CTAs) max {max(en) if en else 0:.1f} | '
          f'last CTA (bins the sheet) end {max(last) if last else 0:.1f}')
