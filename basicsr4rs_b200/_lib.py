"""ctypes binding of the srb200 C-ABI (include/srb200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (plain ``nvcc -shared``) and
loaded from ``basicsr4rs_b200/csrc/libsrb200.so``.  There is **no fallback**: if the library is
missing, or a call returns a non-zero status, a ``RuntimeError`` is raised -- the product path
never routes through PyTorch ops or the CPU oracle.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# this repo: basicsr4rs_b200/csrc/libsrb200.so; grafted into the reference (tools/graft_into_reference.py): the op
# package holds the library next to this file and the sources under src/
_SRC_DIR = os.path.join(_HERE, 'csrc') if os.path.isdir(os.path.join(_HERE, 'csrc')) else os.path.join(_HERE, 'src')
LIB_PATH = os.path.join(_HERE, 'csrc', 'libsrb200.so') if os.path.isdir(os.path.join(_HERE, 'csrc')) \
    else os.path.join(_HERE, 'libsrb200.so')
NVCC_FLAGS = ['-std=c++17', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-shared',
              '-cudart', 'shared', '-Xcompiler', '-fPIC', '-Xlinker', '-rpath=/usr/local/cuda/lib64']

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_GELU = 0, 1, 2, 3
OUT_NHWC, OUT_SHUFFLE, OUT_NCHW_F32 = 0, 1, 2
MASK_NONE, MASK_SIGN, MASK_DGELU, MASK_MUL = 0, 1, 2, 3


class TapGemmDesc(ctypes.Structure):
    """Mirror of ``srb200_tapgemm_desc`` (include/srb200.h)."""
    _fields_ = [
        ('B', c_int32), ('H', c_int32), ('W', c_int32), ('Cin', c_int32), ('src_r', c_int32),
        ('Cout', c_int32), ('ksize', c_int32), ('flip', c_int32), ('act', c_int32),
        ('act_slope', c_float), ('alpha', c_float), ('mask_mode', c_int32), ('mask_slope', c_float),
        ('out_mode', c_int32), ('out_r', c_int32), ('out_c', c_int32), ('out_scale', c_float),
    ]


EXT_ALPHA_ON_BIAS = 1  # srb200_tapgemm_ext.flags


class TapGemmExt(ctypes.Structure):
    """Mirror of ``srb200_tapgemm_ext``."""
    _fields_ = [('residual_f32', c_void_p), ('out_f32', c_void_p), ('alpha_per_sample', c_void_p),
                ('aux_mode', ctypes.c_int32), ('colsum_per_image', ctypes.c_int32), ('colsum', c_void_p),
                ('colsum_scale', c_float), ('flags', ctypes.c_int32),
                ('ln_stats_in', c_void_p), ('ln_wsum', c_void_p), ('ln_stats_out', c_void_p),
                ('ln_channels', ctypes.c_int32), ('ln_eps', c_float)]


# numpy mirror of ``srb200_patch_item`` (one crop of a batched patch extraction)
PATCH_ITEM_FIELDS = [('src', '<i8'), ('pitch', '<i8'), ('top', '<i4'), ('left', '<i4'), ('flags', '<i4'),
                     ('reserved', '<i4')]

# numpy mirror of ``srb200_pack_item`` (one row per weight of a batched pack / unpack launch)
PACK_ITEM_FIELDS = [('src', '<i8'), ('dst', '<i8'), ('perm_out', '<i8'), ('perm_in', '<i8'), ('Co', '<i8'),
                    ('Ci', '<i8'), ('taps', '<i8'), ('Np', '<i8'), ('Kp', '<i8'), ('transpose', '<i8'),
                    ('chunk_begin', '<i8'), ('alpha', '<f4'), ('reserved', '<i4')]


# every symbol include/srb200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    'srb200_version': (c_char_p, []),
    'srb200_strerror': (c_char_p, [c_int]),
    'srb200_check_device': (c_int, [c_int]),
    'srb200_pixel_shuffle_nchw': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p]),
    'srb200_pixel_shuffle_nhwc': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p]),
    'srb200_window_remap': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p]),
    'srb200_roll_nhwc': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'srb200_nchw_to_nhwc': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_float,
                                    c_void_p]),
    'srb200_nhwc_to_nchw': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_float,
                                    c_void_p]),
    'srb200_pack_weight': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                   c_void_p]),
    'srb200_unpack_wgrad': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                    c_float, c_void_p]),
    'srb200_nearest_up2': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'srb200_tap_stencil': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float,
                                   c_void_p]),
    'srb200_tap_im2col': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    'srb200_multi_axpby': (c_int, [c_void_p, c_int, c_int64, c_float, c_float, c_void_p]),
    'srb200_multi_adam': (c_int, [c_void_p, c_int, c_int64, c_double, c_double, c_double, c_double, c_double, c_int64,
                                  c_void_p]),
    'srb200_pack_weights': (c_int, [c_void_p, c_int, c_int64, c_void_p]),
    'srb200_unpack_wgrads': (c_int, [c_void_p, c_int, c_int64, c_void_p]),
    'srb200_unpack_wgrads_inline': (c_int, [c_void_p, c_int, c_void_p]),
    'srb200_tapgemm': (c_int, [POINTER(TapGemmDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, POINTER(TapGemmExt), c_void_p]),
    'srb200_wgrad': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_void_p]),
    'srb200_colsum': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    'srb200_act_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    'srb200_layernorm_fwd': (c_int, [c_void_p] * 6 + [c_int64, c_int, c_int, c_float, c_int, c_void_p]),
    'srb200_layernorm_bwd': (c_int, [c_void_p] * 9 + [c_int64, c_int, c_int, c_void_p]),
    'srb200_scale_rows': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    'srb200_window_attention_fwd': (c_int, [c_void_p] * 4 + [c_int] * 7 + [c_float, c_int, c_void_p, c_void_p]),
    'srb200_window_attention_bwd': (c_int, [c_void_p] * 7 + [c_int] * 7 + [c_float, c_void_p]),
    'srb200_channel_pool': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'srb200_channel_dot': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    'srb200_ca_fc': (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_void_p]),
    'srb200_ca_apply': (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_float, c_void_p]),
    'srb200_ca_forward': (c_int, [c_void_p] * 12 + [c_int, c_int, c_int, c_int, c_float, c_void_p]),
    'srb200_ca_backward': (c_int, [c_void_p] * 15 + [c_int, c_int, c_int, c_int, c_float, c_void_p]),
    'srb200_ca_fc_bwd': (c_int, [c_void_p] * 11 + [c_int, c_int, c_int, c_void_p]),
    'srb200_set_pdl': (c_int, [c_int]),
    'srb200_debug_set_trace': (c_int, [c_void_p]),
    'srb200_debug_set_wgrad_trace': (c_int, [c_void_p]),
    'srb200_debug_set_attn_trace': (c_int, [c_void_p]),
    'srb200_patch_from_u8': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    'srb200_tile_blend_add': (c_int, [c_void_p, c_void_p] + [c_int] * 12 + [c_void_p]),
    'srb200_split3_bf16': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    'srb200_layernorm_f32': (c_int, [c_void_p] * 5 + [c_int64, c_int, c_int, c_float, c_void_p]),
    'srb200_window_attention_f32': (c_int, [c_void_p] * 4 + [c_int] * 7 + [c_float, c_void_p]),
    'srb200_tensor2img_u8': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p]),
    'srb200_ca_apply_bwd': (c_int, [c_void_p] * 4 + [c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
}

_lib = None


def _include_dir():
    for d in (os.path.join(os.path.dirname(_HERE), 'include'), os.path.join(_HERE, 'include')):
        if os.path.exists(os.path.join(d, 'srb200.h')):
            return d
    raise RuntimeError('include/srb200.h not found next to the srb200 sources')


def jit_build(force=False):
    """``BASICSR_JIT=True`` (the reference's convention, ops/fused_act/fused_act.py:8-27): compile the library on first
    import.  Rank-safe: an exclusive ``flock`` on a lock file next to the library serialises the ranks of a torchrun
    job; the first one builds (into a temporary name, then an atomic rename), the others find it fresh."""
    import fcntl
    import subprocess
    srcs = sorted(os.path.join(_SRC_DIR, f) for f in os.listdir(_SRC_DIR) if f.endswith('.cu'))
    deps = srcs + [os.path.join(_SRC_DIR, f) for f in os.listdir(_SRC_DIR) if f.endswith(('.cuh', '.h'))]
    with open(LIB_PATH + '.lock', 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            fresh = os.path.exists(LIB_PATH) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_PATH) for d in deps)
            if fresh and not force:
                return LIB_PATH
            nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
            tmp = f'{LIB_PATH}.{os.getpid()}.tmp'
            subprocess.run([nvcc] + NVCC_FLAGS + ['-I', _include_dir(), '-o', tmp] + srcs, check=True, cwd=_SRC_DIR)
            os.replace(tmp, LIB_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def load():
    """Load libsrb200.so once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get('BASICSR_JIT') == 'True':
        jit_build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: build it with `python __graft_entry__.py build` '
                           '(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU/PyTorch fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


launch_count = 0  # every successful C-ABI compute call enqueues exactly one CUDA kernel
TRACE = None      # optional list: (what, cuda event recorded right after the launch) -- tools/trace_step.py


def check(status, what):
    global launch_count
    launch_count += 1
    if TRACE is not None:
        import torch
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        TRACE.append((what, ev))
    if status != 0:
        msg = load().srb200_strerror(status).decode()
        raise RuntimeError(f'{what} failed: {msg} (status {status})')
