"""CPU-side checks (``-m "not gpu"``): the C-ABI library loads and exports every symbol the header
declares, the registry / plugin surface behaves like the reference's, constructors reproduce the
reference's parameter names, shapes and seeded init, and the product path refuses CPU tensors."""
import os
import re

import pytest
import torch

from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from basicsr4rs_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, 'include', 'srb200.h')).read()
    declared = set(re.findall(r'\b(srb200_[a-z0-9_]+)\s*\(', header))
    declared -= {'srb200_stream_t'}
    assert len(declared) >= 15
    handle = lib.load()
    for name in declared:
        assert hasattr(handle, name), f'{name} declared in include/srb200.h but not exported'
    assert declared == set(lib.SIGNATURES), (declared ^ set(lib.SIGNATURES))
    assert handle.srb200_version().startswith(b'srb200')
    assert b'invalid' in handle.srb200_strerror(-1)


def test_struct_layout_matches_header(lib):
    import ctypes
    # 17 x 4-byte fields, no padding (include/srb200.h: srb200_tapgemm_desc)
    assert ctypes.sizeof(lib.TapGemmDesc) == 17 * 4


def test_ctypes_mirrors_have_the_c_layout(tmp_path):
    """sizeof / offsetof of srb200_tapgemm_desc and srb200_tapgemm_ext as gcc lays them out from include/srb200.h
    against the ctypes mirrors in _lib.py (a field added on one side only would shift every later pointer)."""
    import ctypes
    import shutil
    import subprocess
    from basicsr4rs_b200 import _lib as L
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ext_fields = [f[0] for f in L.TapGemmExt._fields_]
    desc_fields = [f[0] for f in L.TapGemmDesc._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "srb200.h"', 'int main(void) {',
             '  printf("%zu %zu\\n", sizeof(srb200_tapgemm_desc), sizeof(srb200_tapgemm_ext));']
    lines += [f'  printf("%zu\\n", offsetof(srb200_tapgemm_desc, {f}));' for f in desc_fields]
    lines += [f'  printf("%zu\\n", offsetof(srb200_tapgemm_ext, {f}));' for f in ext_fields]
    lines += ['  return 0;', '}']
    src = tmp_path / 'abi.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'abi'
    r = subprocess.run(['gcc', '-std=c99', '-I', os.path.join(root, 'include'), str(src), '-o', str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr  # (also: every mirrored field name exists in the header's struct)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    nums = [int(v) for v in out]
    assert nums[0] == ctypes.sizeof(L.TapGemmDesc) and nums[1] == ctypes.sizeof(L.TapGemmExt), nums[:2]
    want = [getattr(L.TapGemmDesc, f).offset for f in desc_fields] + [getattr(L.TapGemmExt, f).offset for f in ext_fields]
    assert nums[2:] == want


def test_registry_semantics():
    from basicsr4rs_b200.utils.registry import Registry
    reg = Registry('arch')

    @reg.register()
    class Foo:
        pass

    assert reg.get('Foo') is Foo and 'Foo' in reg
    with pytest.raises(AssertionError):
        reg.register(Foo)
    with pytest.raises(KeyError):
        reg.get('Bar')


def test_build_network_and_cpu_refusal():
    from basicsr4rs_b200.archs import ARCH_REGISTRY, build_network
    assert 'EDSR' in ARCH_REGISTRY
    opt = dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=64, num_block=1, upscale=2)
    net = build_network(opt)
    assert opt['type'] == 'EDSR'  # opt is deep-copied, not mutated (archs/__init__.py:18)
    assert 'mean' not in net.state_dict()  # plain attribute, edsr_arch.py:42
    str(net)
    with pytest.raises(RuntimeError, match='CUDA'):
        net(torch.rand(1, 3, 8, 8))
    with pytest.raises(ValueError):
        build_network(dict(type='EDSR', num_in_ch=3, num_out_ch=3, upscale=5))


EDSR_M = dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=16, upscale=4, res_scale=1, img_range=255.)


def test_edsr_param_count_known_answer():
    from basicsr4rs_b200.archs import build_network
    net = build_network(dict(type='EDSR', **EDSR_M))
    assert sum(p.numel() for p in net.parameters()) == 1517571


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not mounted (GPU box)')
@pytest.mark.parametrize('arch,kwargs', [
    ('EDSR', EDSR_M),
    ('EDSR', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=2, upscale=3, res_scale=0.1)),
    ('RCAN', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_group=2, num_block=3, squeeze_factor=16, upscale=4)),
    ('SwinIR', dict(upscale=4, in_chans=3, img_size=64, window_size=8, img_range=1., depths=[6] * 6, embed_dim=180,
                    num_heads=[6] * 6, mlp_ratio=2, upsampler='pixelshuffle', resi_connection='1conv')),
    ('SwinIR', dict(upscale=2, in_chans=4, img_size=48, window_size=6, img_range=1., depths=[2, 2], embed_dim=60,
                    num_heads=[6, 6], mlp_ratio=2, upsampler='pixelshuffledirect', resi_connection='3conv', ape=True)),
])
def test_seeded_init_identical_to_reference(arch, kwargs):
    """Same seed -> bit-identical state dict (names, order, shapes, values): 'identical random-init weights'."""
    from basicsr4rs_b200.archs import build_network
    ref = ref_shim.load_reference_archs()
    torch.manual_seed(123)
    ours = build_network(dict(type=arch, **kwargs)).state_dict()
    torch.manual_seed(123)
    theirs = getattr(ref, arch)(**kwargs).state_dict()
    assert list(ours.keys()) == list(theirs.keys())
    for k in ours:
        assert ours[k].dtype == theirs[k].dtype and torch.equal(ours[k], theirs[k]), k


def test_pack_item_table_matches_the_c_struct():
    """The numpy mirror of ``srb200_pack_item`` (include/srb200.h) has the C layout: 96-byte rows, 8-byte fields at
    the offsets the header implies, chunk_begin = running sum of ceil(Np*Kp/256)."""
    import numpy as np
    from basicsr4rs_b200 import _lib as L
    from basicsr4rs_b200.ops.sr_b200 import raw
    dt = np.dtype(L.PACK_ITEM_FIELDS)
    assert dt.itemsize == 96
    assert [dt.fields[n][1] for n in ('src', 'dst', 'perm_out', 'perm_in', 'Co', 'Ci', 'taps', 'Np', 'Kp', 'transpose',
                                      'chunk_begin', 'alpha', 'reserved')] == [0, 8, 16, 24, 32, 40, 48, 56, 64, 72, 80,
                                                                               88, 92]
    a, b = torch.zeros(4), torch.zeros(4)
    rows = [dict(src=a, dst=b, Co=64, Ci=64, taps=9, Np=64, Kp=64),
            dict(src=a, dst=b, Co=3, Ci=64, taps=9, Np=16, Kp=64, transpose=True, alpha=0.5),
            dict(src=a, dst=b, Co=180, Ci=1, taps=1, Np=192, Kp=1)]
    tab, total = raw._item_table(rows)
    assert list(tab['chunk_begin']) == [0, 16, 20] and total == 21
    assert tab[1]['transpose'] == 1 and tab[1]['alpha'] == 0.5 and tab[0]['src'] == a.data_ptr()


def test_backward_scratch_provider_measures_then_hands_out_zeroed_slices():
    """raw.backward_scratch: a measuring pass sizes the one fill; slices are disjoint, 16-byte aligned, big enough
    for the zero_arena they back, and a provider without room degrades to None (the Function then fills itself)."""
    from basicsr4rs_b200.ops.sr_b200 import raw
    cpu = torch.device('cpu')
    assert raw.take_backward_scratch(100, cpu) is None  # no provider
    with raw.backward_scratch(cpu, None) as m:
        assert raw.take_backward_scratch(100, cpu) is None and raw.take_backward_scratch(7, cpu) is None
    assert m.measured >= 100 + 64 + 7 + 64
    with raw.backward_scratch(cpu, m.measured):
        a = raw.take_backward_scratch(100, cpu)
        b = raw.take_backward_scratch(7, cpu)
        assert a.numel() == 164 and b.numel() == 71 and not a.any() and not b.any()
        assert a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
        assert a.data_ptr() + 4 * a.numel() <= b.data_ptr()
        assert raw.take_backward_scratch(1, cpu) is None  # exhausted
        with raw.zero_arena(cpu, 100, buf=a):   # carves from the stashed slice, no new fill
            z = raw.zeros_f32((10, 10), cpu)
            assert z.data_ptr() == a.data_ptr()
        with raw.zero_arena(cpu, 100, buf=b):   # too small: falls back to its own zeros
            assert raw.zeros_f32((10, 10), cpu).data_ptr() != b.data_ptr()
    assert raw.take_backward_scratch(100, cpu) is None

    class Ctx:
        needs_input_grad = (True, False)
    ctx = Ctx()
    with raw.backward_scratch(cpu, 1000):
        raw.stash_backward_scratch(ctx, 100, cpu)
    first = ctx.scratch
    assert first is not None
    with raw.backward_arena(ctx, cpu, 100):
        assert raw.zeros_f32((4,), cpu).data_ptr() == first.data_ptr()
    assert ctx.scratch is None  # single use
    with raw.backward_arena(ctx, cpu, 100):
        assert raw.zeros_f32((4,), cpu).data_ptr() != first.data_ptr()


def test_public_header_is_plain_c(tmp_path):
    """include/srb200.h is the drop-in boundary: it must compile as C99 (and as C++) on its own -- no torch, no CUDA
    headers, plain pointers and sizes only."""
    import shutil
    import subprocess
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'srb200.h')
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    for lang, std in (('c', '-std=c99'), ('c++', '-std=c++17')):
        r = subprocess.run(['gcc', std, '-Wall', '-Wextra', '-pedantic', '-Werror', '-fsyntax-only', '-x', lang, hdr],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    includes = re.findall(r'^\s*#\s*include\s*[<"]([^>"]+)[>"]', open(hdr).read(), flags=re.M)
    assert set(includes) <= {'stdint.h', 'stddef.h'}, includes


def test_segment_sizes_its_backward_scratch_on_the_first_grad_enabled_run():
    """archs/graphed.Segment: the first grad-enabled run per input shape only measures what its Functions ask for,
    later runs hand the slices out of one fill; no_grad runs never allocate."""
    from basicsr4rs_b200.archs.graphed import Segment
    from basicsr4rs_b200.ops.sr_b200 import raw
    got = []

    def fn(t):
        got.append((raw.take_backward_scratch(100, t.device), raw.take_backward_scratch(28, t.device)))
        return t * 2

    seg = Segment(fn, [torch.nn.Identity()])
    x = torch.ones(2, 3)
    seg(x)                                    # measuring pass
    assert got[-1] == (None, None)
    seg(x)                                    # sized pass
    a, b = got[-1]
    assert a is not None and b is not None and a.numel() == 164 and b.numel() == 92 and not a.any()
    with torch.no_grad():
        seg(x)
    assert got[-1] == (None, None)
    seg(torch.ones(4, 3))                     # new shape: measured again
    assert got[-1] == (None, None)
    seg(torch.ones(4, 3))
    assert got[-1][0] is not None


def test_drop_path_bank_draws_a_layer_at_once_with_the_reference_distribution(monkeypatch):
    """All DropPath factors of an RSTB layer from one draw (swinir_arch.drop_path_bank): every row is
    floor(keep + U) / keep per sample as in the reference (swinir_arch.py:14-26), p = 0 entries are None, nothing is
    drawn in eval mode, and a replaced ``drop_path_scale`` (the parity tests inject masks) is honoured call by call."""
    from basicsr4rs_b200.archs import swinir_arch as sa
    torch.manual_seed(0)
    probs = [0.0, 0.0, 0.05, 0.05, 0.3, 0.3]
    bank = sa.drop_path_bank(4096, probs, True, torch.device('cpu'))
    assert bank[0] is None and bank[1] is None
    for p, row in zip(probs[2:], bank[2:]):
        keep = 1.0 - p
        assert row.shape == (4096,) and row.dtype == torch.float32 and row.is_contiguous()
        vals = set(round(v, 5) for v in row.unique().tolist())
        assert vals <= {0.0, round(1.0 / keep, 5)}
        assert abs((row > 0).float().mean().item() - keep) < 0.03     # P(keep) = 1 - p
        assert abs(row.mean().item() - 1.0) < 0.05                     # unbiased
    assert not torch.equal(bank[2], bank[3])                            # the two branches of a block draw separately
    assert sa.drop_path_bank(8, probs, False, torch.device('cpu')) == [None] * 6
    calls = []

    def injected(batch, p, training, device):
        calls.append(p)
        return None if p == 0 else torch.full((batch,), 2.0)

    monkeypatch.setattr(sa, 'drop_path_scale', injected)
    out = sa.drop_path_bank(3, probs, True, torch.device('cpu'))
    assert calls == probs and out[0] is None and torch.equal(out[5], torch.full((3,), 2.0))


def test_flat_grads_layout_and_buckets():
    """FlatGrads / FlatDDP host logic on CPU: 16-byte aligned slices in registration order, buckets in reverse
    registration order that cover the buffer exactly once."""
    from basicsr4rs_b200.utils.flat_ddp import FlatDDP, FlatGrads
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 300))
    fg = FlatGrads(net)
    assert [o % 4 for o in fg.offsets] == [0] * 6 and fg.offsets == sorted(fg.offsets)
    for p, v in zip(fg.params, fg.views):
        assert v.shape == p.shape and v.untyped_storage().data_ptr() == fg.buf.untyped_storage().data_ptr()
    ddp = FlatDDP(net, bucket_mb=300 * 4 / 2**20, last_mb=40 * 4 / 2**20)   # ~300 floats per bucket, ~40 in the last one
    fg = ddp.flat                                     # (the wrap owns the network's buffer: flat_grads_of)
    seen = sorted(i for b in ddp.buckets for i in b[2])
    assert seen == list(range(6))
    assert ddp.buckets[0][2] == [0] and len(ddp.buckets) >= 3   # the bucket reduced last (first parameters) is small
    spans = sorted((b[0], b[1]) for b in ddp.buckets)
    assert spans[0][0] == 0 and spans[-1][1] == fg.total and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    # without a process group a backward still leaves .grad aliasing the flat slices
    net(torch.ones(2, 7)).sum().backward()
    for p, v in zip(fg.params, fg.views):
        assert p.grad.data_ptr() == v.data_ptr()


def test_layernorm_fold_bank_host_logic(monkeypatch):
    """BasicLayer._fold_bank / _fold_ok on the CPU with the pack launch replaced by its definition (bf16 [1, Np, Kp],
    rows permuted by perm_out, zero padded): the folded operands reproduce Linear(LayerNorm(x)) through
    rstd * (W'x - mean * wsum) + bias', in the packed channel order, for qkv (head-padded layout) and fc1."""
    from basicsr4rs_b200.archs import swinir_arch
    from basicsr4rs_b200.ops.sr_b200 import raw, swin_ops

    def fake_pack(w, n_pad, k_pad, perm_out=None, perm_in=None, transpose=False, out=None):
        assert perm_in is None and not transpose
        full = torch.zeros((n_pad, k_pad))
        rows = torch.arange(n_pad) if perm_out is None else perm_out.long()
        ok = (rows >= 0) & (rows < w.shape[0])
        full[ok, :w.shape[1]] = w[rows[ok]]
        res = full.to(torch.bfloat16).reshape(1, n_pad, k_pad)
        if out is not None:
            out.copy_(res)
            return out
        return res

    monkeypatch.setattr(raw, 'pack_weight', fake_pack)
    torch.manual_seed(0)
    layer = swinir_arch.BasicLayer(dim=180, input_resolution=(16, 16), depth=2, num_heads=6, window_size=8, mlp_ratio=2.)
    layer.eval()
    for blk in layer.blocks:
        for nm in (blk.norm1, blk.norm2):
            nm.weight.data.normal_(1.0, 0.2)
            nm.bias.data.normal_(0.0, 0.2)
        for lin in (blk.attn.qkv, blk.mlp.fc1):
            lin.bias.data.normal_(0.0, 0.1)
    with torch.no_grad():
        assert layer._fold_ok(torch.zeros((1, 512, 512, 192)))          # a big evaluation tile
        assert not layer._fold_ok(torch.zeros((1, 64, 64, 192)))        # below FOLD_MIN_TOKENS
        assert not layer._fold_ok(torch.zeros((1, 512, 512, 64)))       # no two-team kernel for 64-wide rows
        layer.train()
        assert not layer._fold_ok(torch.zeros((1, 512, 512, 192)))
        layer.eval()
        fq, f1 = layer._fold_bank(192, torch.device('cpu'))
        x = torch.randn(50, 180) * 2 + 0.3
        xp = torch.nn.functional.pad(x, (0, 12))
        mean = x.mean(-1, keepdim=True)
        rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5)
        for i, blk in enumerate(layer.blocks):
            perm = swin_ops.head_perm(6, 30, 3, torch.device('cpu')).long()
            for (wp, wsum, bp), lin, nm, pm in ((fq[i], blk.attn.qkv, blk.norm1, perm), (f1[i], blk.mlp.fc1, blk.norm2, None)):
                ref = lin(nm(x))
                got = rstd * (xp @ wp[0].float().T - mean * wsum[None]) + bp[None]
                if pm is None:
                    want = torch.nn.functional.pad(ref, (0, got.shape[1] - ref.shape[1]))
                else:
                    want = torch.where(pm >= 0, ref[:, pm.clamp(min=0)], torch.zeros(()))
                # (bf16 rounding of W' is the only difference: ~2^-9 relative per weight)
                assert (got - want).abs().max().item() <= 2e-2 * ref.abs().max().item()
                assert torch.count_nonzero(got[:, (want == 0).all(0)]) == 0   # pad channels stay exactly zero


def test_fp32_mode_operand_split_algebra():
    """The arithmetic behind ``compute_dtype='fp32'`` (ops/sr_b200/fp32_mode.py, srb200_split3_bf16), on the CPU: with
    x = hi + lo (+ O(2^-17 |x|)) in bf16 pairs, the concatenated reduction [x_hi | x_lo | x_hi] . [w_hi ; w_hi ; w_lo],
    accumulated in fp32 as the tensor core does, matches the fp64 product to ~2^-16 relative -- 30x tighter than a
    plain bf16 GEMM and inside the 1e-4 bar of the north star."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn((64, 256), generator=g)
    w = torch.randn((32, 256), generator=g) * 0.05

    def split(t):
        hi = t.to(torch.bfloat16).float()
        return hi, (t - hi).to(torch.bfloat16).float()

    xh, xl = split(x)
    wh, wl = split(w)
    assert (x - (xh + xl)).abs().max() <= 2.0**-16 * x.abs().max()
    ref = x.double() @ w.double().T
    got = (torch.cat([xh, xl, xh], 1) @ torch.cat([wh, wh, wl], 1).T)          # fp32 accumulation
    plain = xh @ wh.T
    scale = (x.abs().double() @ w.abs().double().T).max().item()               # sum |x||w|: the rounding-error scale
    err, err_plain = (got.double() - ref).abs().max().item(), (plain.double() - ref).abs().max().item()
    assert err <= 2.0**-14 * scale and err_plain >= 30 * err, (err, err_plain, scale)
