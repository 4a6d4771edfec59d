"""TEST INFRASTRUCTURE ONLY -- recipe that vendors the UNMODIFIED reference arch files into ``oracle/_ref/``.

The reference's implementation of the hot path is pure Python (basicsr/archs/{arch_util,edsr_arch,rcan_arch,
swinir_arch}.py on stock ``torch``); nothing has to be compiled.  ``/root/reference`` does not exist on the GPU box,
so this script -- run by ``__graft_entry__.build()`` in the build container -- copies the handful of files
``oracle/ref_shim.py`` imports, byte for byte, to ``oracle/_ref/basicsr/...``.  ``oracle/_ref/`` is git-ignored
(no reference source enters the history) but NOT gpurun-ignored, so it travels to the box like a built ``.so``.
There it serves as (a) the ``--impl reference`` arm of ``bench.py`` (``cpu_baseline.kind = "reference"``) and
(b) the live checker of the ``-m gpu`` parity tests.  Never imported by the product package.

    python -m oracle.make_ref            # refresh oracle/_ref from $BASICSR_REF_ROOT (default /root/reference)
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
# the files oracle/ref_shim.py imports (the hot path itself) ...
FILES = [
    'basicsr/utils/registry.py',
    'basicsr/ops/dcn/__init__.py',
    'basicsr/ops/dcn/deform_conv.py',
    'basicsr/archs/arch_util.py',
    'basicsr/archs/edsr_arch.py',
    'basicsr/archs/rcan_arch.py',
    'basicsr/archs/swinir_arch.py',
    'LICENSE.txt',
]
# ... and the CALLERS of the path, which tools/graft_into_reference.py needs to prove the drop-in boundary on the
# reference itself (build_network, SRModel.optimize_parameters, SwinIRModel.test, the shipped YAML options): the
# Python sources of the package and the option files, nothing else (no assets, docs, data, notebooks).
TREES = [('basicsr', ('.py',)), ('options', ('.yml', '.yaml'))]


def main(src_root=None, quiet=False):
    src_root = src_root or os.environ.get('BASICSR_REF_ROOT', '/root/reference')
    if not os.path.isdir(os.path.join(src_root, 'basicsr', 'archs')):
        raise RuntimeError(f'reference tree not found at {src_root}')
    copied = 0
    for rel in FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            copied += 1
    n_tree = 0
    for top, exts in TREES:
        for dirpath, _, names in os.walk(os.path.join(src_root, top)):
            for name in names:
                if not name.endswith(exts):
                    continue
                src = os.path.join(dirpath, name)
                dst = os.path.join(DST, os.path.relpath(src, src_root))
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
                    shutil.copyfile(src, dst)
                    copied += 1
                n_tree += 1
    with open(os.path.join(DST, 'PROVENANCE.txt'), 'w') as f:
        f.write(f'verbatim copies of {len(FILES)} + {n_tree} files from {src_root} (watercore2001/BasicSR4RS), made by '
                'oracle/make_ref.py; git-ignored, test infrastructure only (never imported by basicsr4rs_b200)\n')
    if not quiet:
        print(f'[make_ref] {copied} file(s) refreshed under {DST}')
    return DST


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else None)
