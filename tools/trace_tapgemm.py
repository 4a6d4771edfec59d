"""Per-role clock64 timeline of CTA 0 for one tap-GEMM launch (debug aid)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr4rs_b200 import _lib as L
from basicsr4rs_b200.ops.sr_b200 import raw
dev = torch.device('cuda:0')
def run(name, b, h, w, cin, cout, ks, **kw):
    x = torch.randn((b, h, w, cin), device=dev).to(torch.bfloat16)
    wt = torch.randn((cout, cin, ks, ks), device=dev) * 0.02
    wp = raw.pack_weight(wt, cout, cin); bias = torch.zeros(cout, device=dev)
    for _ in range(3): raw.tapgemm(x, wp, ksize=ks, cout=cout, bias=bias, **kw)
    buf = torch.zeros(4 * 64 + 4 * 148, dtype=torch.int64, device=dev)
    L.load().srb200_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    raw.tapgemm(x, wp, ksize=ks, cout=cout, bias=bias, **kw)
    torch.cuda.synchronize(); L.load().srb200_debug_set_trace(None)
    full = buf.cpu(); t = full[:256].view(4, 64); t0 = int(t[0, 63])
    c = full[256:].view(148, 4)
    c = c[c[:, 0] > 0]
    g0 = int(c[:, 0].min())
    dur_ns = (c[:, 1] - c[:, 0]).float(); dur_clk = (c[:, 3] - c[:, 2]).float()
    print(f'== {name}: {c.shape[0]} CTAs; start spread {int(c[:,0].max())-g0} ns; kernel wall (first start -> last end) {int(c[:,1].max())-g0} ns;'
          f' CTA duration ns min/mean/max {dur_ns.min():.0f}/{dur_ns.mean():.0f}/{dur_ns.max():.0f}; clk min/mean/max {dur_clk.min():.0f}/{dur_clk.mean():.0f}/{dur_clk.max():.0f};'
          f' eff GHz {(dur_clk/dur_ns).mean():.3f}')
    rel = lambda v: (int(v) - t0) if int(v) else None
    print(f'== {name}: cycles relative to kernel start (CTA 0)')
    print(' producer tile starts :', [rel(v) for v in t[0, :12]])
    print(' mma tile start/end   :', [(rel(t[1, 2*i]), rel(t[1, 2*i+1])) for i in range(12) if int(t[1, 2*i])])
    print(' producer k-block (before wait, after A issue) tile 0:', [(rel(t[3, 2*i]), rel(t[3, 2*i+1])) for i in range(16) if int(t[3, 2*i])])
    print(' epi tile 2 chunk phases (store slot free, team in, staged operand landed, math + staging done, team out):', [[rel(v) for v in t[3, 32+7*c:37+7*c]] for c in range(4) if int(t[3, 32+7*c])])
    print(' epi  acc ready/drained:', [(rel(t[2, 2*i]), rel(t[2, 2*i+1])) for i in range(12) if int(t[2, 2*i])])
def run_wgrad(name, b, h, w, cin, cout, ks):
    x = torch.randn((b, h, w, cin), device=dev).to(torch.bfloat16)
    dy = torch.randn((b, h, w, cout), device=dev).to(torch.bfloat16)
    for _ in range(3): raw.wgrad(dy, x, ksize=ks)
    buf = torch.zeros(128, dtype=torch.int64, device=dev)
    L.load().srb200_debug_set_wgrad_trace(ctypes.c_void_p(buf.data_ptr()))
    raw.wgrad(dy, x, ksize=ks)
    torch.cuda.synchronize(); L.load().srb200_debug_set_wgrad_trace(None)
    t = buf.cpu(); t0 = int(t[0]); rel = lambda v: int(v) - t0 if int(v) else None
    print(f'== wgrad {name}: mma issued-all {rel(t[1])}, acc ready {rel(t[2])}, cta end {rel(t[3])}')
    print('   k-block ready times:', [rel(v) for v in t[8:48]])
    print('   producer (before wait, after wait, after issue):', [(rel(t[64+3*i]), rel(t[65+3*i]), rel(t[66+3*i])) for i in range(20)])
if len(sys.argv) > 1 and sys.argv[1] == 'wgrad':
    run_wgrad('proj 192x192', 16, 64, 64, 192, 192, 1)
    run_wgrad('qkv 192->576', 16, 64, 64, 192, 576, 1)
    run_wgrad('fc1 192->384', 16, 64, 64, 192, 384, 1)
    run_wgrad('fc2 384->192', 16, 64, 64, 384, 192, 1)
    run_wgrad('rcan 64x64 3x3', 16, 48, 48, 64, 64, 3)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == 'rcan':
    run('rcan 64->64 48x48 relu', 16, 48, 48, 64, 64, 3, act=L.ACT_RELU)
    run('rcan 64->64 48x48', 16, 48, 48, 64, 64, 3)
    run('rcan 64->64 48x48 B8', 8, 48, 48, 64, 64, 3)
    run('edsr body 256->256 48x48 relu', 16, 48, 48, 256, 256, 3, act=L.ACT_RELU)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == 'mlp':
    x384 = torch.randn((16, 64, 64, 384), device=dev).to(torch.bfloat16)
    x192 = torch.randn((16, 64, 64, 192), device=dev).to(torch.bfloat16)
    run('fc2 dgrad 192->384 x mask', 16, 64, 64, 192, 384, 1, mask_src=x384, mask_mode=L.MASK_MUL)
    run('proj 192->192 + residual', 16, 64, 64, 192, 192, 1, residual=x192)
    run('fc2 384->192 + residual', 16, 64, 64, 384, 192, 1, residual=x192)
    run('plain 192->384', 16, 64, 64, 192, 384, 1)
    sys.exit(0)
run('qkv 65536x192 -> 576', 16, 64, 64, 192, 576, 1)
run('fc1 192->384 gelu+aux', 16, 64, 64, 192, 384, 1, act=L.ACT_GELU, want_aux=True)
run('fc1 192->384 plain', 16, 64, 64, 192, 384, 1)
run('fc1 192->384 gelu no aux', 16, 64, 64, 192, 384, 1, act=L.ACT_GELU)
run('edsr body 256->256 48x48 relu', 16, 48, 48, 256, 256, 3, act=L.ACT_RELU)
run('rcan 64->64 48x48', 16, 48, 48, 64, 64, 3)

run_wgrad('edsr body 256x256 48x48', 16, 48, 48, 256, 256, 3)
