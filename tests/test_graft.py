"""The drop-in boundary proven ON THE REFERENCE ITSELF (tools/graft_into_reference.py): a copy of the reference tree
gets the four arch files + the op package, and then the reference's OWN code -- ``basicsr/archs/__init__.py`` registry
scan and ``build_network`` (:11-23), ``SRModel.__init__ / feed_data / optimize_parameters / test`` (sr_model.py:23-129),
``BaseModel.model_to_device`` with DistributedDataParallel (base_model.py:87-105), ``model_ema`` (:75-82),
``SwinIRModel.test`` (swinir_model.py:14-36) -- drives the B200 kernels.  The reference tree is /root/reference in the
build container and the verbatim copy oracle/_ref (oracle/make_ref.py) on the GPU box."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import graft_into_reference as gr  # noqa: E402


def _have_reference():
    try:
        gr.find_reference()
        return True
    except RuntimeError:
        return False


needs_ref = pytest.mark.skipif(not _have_reference(), reason='no reference tree with basicsr/models')


@pytest.fixture(scope='module')
def grafted(tmp_path_factory):
    root = gr.graft(out_dir=str(tmp_path_factory.mktemp('graft')))
    ns = gr.import_grafted(root)
    yield ns
    for name in [m for m in sys.modules if m == 'basicsr' or m.startswith('basicsr.')]:
        del sys.modules[name]  # the oracle's ref_shim installs its own 'basicsr' shim: do not leak this one into it


@needs_ref
def test_reference_registry_serves_the_srb200_archs(grafted):
    assert gr.check(grafted)
    reg = grafted.registry.ARCH_REGISTRY
    assert len(reg._obj_map) >= 30  # every other arch of the reference still registers
    # seeded state dicts of the grafted archs equal those of the unmodified reference (checkpoints interchange)
    from oracle import ref_shim
    kw = dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_group=2, num_block=2, squeeze_factor=16, upscale=2)
    torch.manual_seed(5)
    ours = reg.get('RCAN')(**kw).state_dict()
    mods = {m: sys.modules[m] for m in list(sys.modules) if m == 'basicsr' or m.startswith('basicsr.')}
    try:
        ref = ref_shim.load_reference_archs()
        torch.manual_seed(5)
        theirs = ref.RCAN(**kw).state_dict()
    finally:
        for m in [m for m in sys.modules if m == 'basicsr' or m.startswith('basicsr.')]:
            del sys.modules[m]
        sys.modules.update(mods)
    assert list(ours) == list(theirs) and all(torch.equal(ours[k], theirs[k]) for k in ours)


@needs_ref
def test_shipped_yaml_options_build_through_the_reference_registry(grafted):
    """network_g blocks of the reference's own option files construct unchanged (YAML keys are the ctor kwargs)."""
    import yaml
    opt_root = os.path.join(grafted.root, 'options')
    if not os.path.isdir(opt_root):
        pytest.skip('options/ not vendored')
    built = 0
    for rel in ('train/EDSR/train_EDSR_Lx4.yml', 'train/EDSR/train_EDSR_Mx4.yml', 'test/RCAN/test_RCAN.yml',
                'train/SwinIR/train_SwinIR_SRx4_scratch.yml', 'train/SwinIR/train_SwinIR_L2S288_scratch.yml'):
        path = os.path.join(opt_root, rel)
        if not os.path.exists(path):
            continue
        with open(path) as f:
            opt = yaml.safe_load(f)
        net = grafted.archs.build_network(opt['network_g'])
        assert hasattr(net, '_pack_book'), rel
        built += 1
    assert built >= 3


def _sr_opt(network_g, dist=False, ema=0.999):
    return {'num_gpu': 1, 'is_train': True, 'dist': dist, 'rank': 0, 'world_size': 1, 'scale': network_g.get('upscale', 4), 'network_g': network_g,
            'path': {'pretrain_network_g': None, 'strict_load_g': True},
            'train': {'ema_decay': ema, 'optim_g': {'type': 'Adam', 'lr': 2e-4, 'weight_decay': 0, 'betas': [0.9, 0.99]},
                      'scheduler': {'type': 'MultiStepLR', 'milestones': [1000], 'gamma': 0.5},
                      'pixel_opt': {'type': 'L1Loss', 'loss_weight': 1.0, 'reduction': 'mean'}}}


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize('arch', ['EDSR', 'RCAN', 'SwinIR'])
def test_reference_srmodel_trains_on_the_b200_path(cuda, grafted, arch):
    """The reference's SRModel (unmodified) with DDP wrap + EMA: three optimize_parameters steps and a test() on the
    grafted archs.  The loss must fall on a fixed batch, the EMA copy must follow, test() must use the fresh EMA
    weights (the packed-weight cache must not go stale through ``p.data`` updates)."""
    import torch.distributed as dist
    from basicsr4rs_b200 import _lib as L
    nets = {
        'EDSR': dict(type='EDSR', num_in_ch=3, num_out_ch=3, num_feat=64, num_block=4, upscale=4, res_scale=1,
                     img_range=255., rgb_mean=[0.4488, 0.4371, 0.4040]),
        'RCAN': dict(type='RCAN', num_in_ch=3, num_out_ch=3, num_feat=64, num_group=2, num_block=2, squeeze_factor=16,
                     upscale=4, res_scale=1, img_range=255., rgb_mean=[0.4488, 0.4371, 0.4040]),
        'SwinIR': dict(type='SwinIR', upscale=4, in_chans=3, img_size=16, window_size=8, img_range=1., depths=[2, 2],
                       embed_dim=60, num_heads=[6, 6], mlp_ratio=2, upsampler='pixelshuffle', resi_connection='1conv'),
    }
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29533')
    os.environ['LOCAL_RANK'] = '0'
    if not dist.is_initialized():
        dist.init_process_group('nccl', rank=0, world_size=1, device_id=cuda)
    torch.manual_seed(0)
    cls = grafted.swinir_model.SwinIRModel if arch == 'SwinIR' else grafted.sr_model.SRModel
    model = cls(_sr_opt(nets[arch], dist=True))
    assert isinstance(model.net_g, torch.nn.parallel.DistributedDataParallel)
    if arch == 'RCAN':  # the opt-in hook of SURVEY.md 8f rank 1: loss values stay on the device until they are printed
        from basicsr4rs_b200.utils.train_hooks import install_lazy_loss_log
        install_lazy_loss_log(model)
    # the library loaded is the one inside the grafted op package, called through the grafted binding
    lib_mod = sys.modules['basicsr.ops.sr_b200._lib']
    assert lib_mod.LIB_PATH.startswith(grafted.root)
    g = torch.Generator().manual_seed(1)
    lr = 16 if arch == 'SwinIR' else 24
    data = {'lq': torch.rand((4, 3, lr, lr), generator=g), 'gt': torch.rand((4, 3, 4 * lr, 4 * lr), generator=g)}
    model.feed_data(data)
    n0 = lib_mod.launch_count
    losses = []
    for it in range(1, 7):
        model.optimize_parameters(it)
        losses.append(model.get_current_log()['l_pix'])
    assert lib_mod.launch_count > n0, 'no srb200 kernel was launched'
    assert losses[-1] < losses[0], losses
    model.test()
    out1 = model.output.clone()
    assert out1.shape == (4, 3, 4 * lr, 4 * lr) and torch.isfinite(out1).all()
    # EMA keeps moving -> the next test() must differ (stale packed weights would repeat out1 bit for bit)
    for it in range(7, 10):
        model.optimize_parameters(it)
    model.test()
    assert not torch.equal(out1, model.output)
    # SwinIRModel.test pads odd sizes to the window and crops (swinir_model.py:14-36)
    if arch == 'SwinIR':
        model.feed_data({'lq': torch.rand((1, 3, 19, 13), generator=g)})
        model.test()
        assert model.output.shape == (1, 3, 76, 52)
    del L
