"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference archs for oracle pinning.

``import basicsr`` fails in this image (basicsr/__init__.py:3-12 star-imports modules that need
matplotlib / lmdb / a generated version.py), so this installs a 4-entry ``sys.modules`` shim
 and loads only
  basicsr/utils/registry.py, basicsr/ops/dcn (pure-Python part), basicsr/archs/{arch_util,edsr_arch,
  rcan_arch,swinir_arch}.py
from the read-only reference tree.  Used by ``tests/golden/make_golden.py`` (run in the build
container, where /root/reference is mounted) and by the ``not gpu`` tests that pin ``oracle/sr_oracle.py``
against the live reference.  /root/reference does not exist on the GPU box; there the verbatim copies vendored
by ``oracle/make_ref.py`` into ``oracle/_ref`` (git-ignored, shipped by gpurun) are loaded instead, by the
``-m gpu`` parity tests and by ``bench.py --impl reference`` -- never by the product package.
"""
import importlib
import logging
import os
import sys
import types

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def _pick_root():
    """$BASICSR_REF_ROOT, else the mounted reference tree, else the verbatim copies oracle/make_ref.py vendored
    into oracle/_ref (the only form in which the reference reaches the GPU box)."""
    env = os.environ.get('BASICSR_REF_ROOT')
    if env:
        return env
    for root in ('/root/reference', _VENDORED):
        if os.path.isdir(os.path.join(root, 'basicsr', 'archs')):
            return root
    return '/root/reference'


REF_ROOT = _pick_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'basicsr', 'archs'))


def load_reference_archs():
    """Return a namespace with EDSR, RCAN, SwinIR and arch_util of the reference tree."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    if 'basicsr.archs.swinir_arch' not in sys.modules or not getattr(sys.modules.get('basicsr'), '_srb200_shim', False):
        for name in [m for m in sys.modules if m == 'basicsr' or m.startswith('basicsr.')]:
            del sys.modules[name]
        pkg = types.ModuleType('basicsr')
        pkg.__path__ = [os.path.join(REF_ROOT, 'basicsr')]
        pkg._srb200_shim = True
        utils = types.ModuleType('basicsr.utils')
        utils.__path__ = [os.path.join(REF_ROOT, 'basicsr', 'utils')]
        utils.get_root_logger = lambda *a, **k: logging.getLogger('basicsr')
        ops = types.ModuleType('basicsr.ops')
        ops.__path__ = [os.path.join(REF_ROOT, 'basicsr', 'ops')]
        archs = types.ModuleType('basicsr.archs')
        archs.__path__ = [os.path.join(REF_ROOT, 'basicsr', 'archs')]
        sys.modules.update({'basicsr': pkg, 'basicsr.utils': utils, 'basicsr.ops': ops, 'basicsr.archs': archs})
        importlib.import_module('basicsr.utils.registry')
    ns = types.SimpleNamespace()
    ns.arch_util = importlib.import_module('basicsr.archs.arch_util')
    ns.EDSR = importlib.import_module('basicsr.archs.edsr_arch').EDSR
    ns.RCAN = importlib.import_module('basicsr.archs.rcan_arch').RCAN
    ns.swinir_arch = importlib.import_module('basicsr.archs.swinir_arch')
    ns.SwinIR = ns.swinir_arch.SwinIR
    return ns
