"""Entry points: train.py / test.py keep the reference's command line and run end to end (CPU part: argument surface; GPU part: a short
training run with validation, checkpoint, resume, then test.py on the result) and the DROP-IN proof: the reference's OWN training and
validation functions (train.py:49-129, :132-189) drive this repository's model class after the import swap of INTEGRATION.md.

The drop-in proof needs the reference tree: /root/reference (build container) or baseline/_ref (a git-ignored copy that travels to the
GPU box); it is skipped where neither exists."""
import importlib.util
import json
import os
import shlex
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the README recipes of the reference (README.md:111-127) + the common options (README.md:105)
README_RECIPES = {
    'smmnist': '--ny 20 --nz 20 --beta_z 2 --nt_cond 5 --nt_inf 5 --dataset smmnist --nc 1 --seq_len 15',
    'smmnist_det': '--ny 20 --nz 20 --beta_z 2 --nt_cond 5 --nt_inf 5 --dataset smmnist --deterministic --nc 1 --seq_len 15 '
                   '--lr_scheduling_burnin 800000 --lr_scheduling_n_iter 100000',
    'kth': '--ny 50 --nz 50 --n_euler_steps 2 --res_gain 1.2 --archi vgg --skipco --nt_cond 10 --nt_inf 3 --obs_scale 0.2 --batch_size 100 '
           '--dataset kth --nc 1 --seq_len 20 --lr_scheduling_burnin 150000 --lr_scheduling_n_iter 50000 --val_interval 5000 --seq_len_test 30',
    'human': '--ny 50 --nz 50 --n_euler_steps 2 --res_gain 1.2 --archi vgg --skipco --nt_cond 8 --nt_inf 3 --obs_scale 0.2 --batch_size 100 '
             '--dataset human --nc 3 --seq_len 16 --lr_scheduling_burnin 325000 --lr_scheduling_n_iter 25000 --val_interval 20000 '
             '--batch_size_test 8 --seq_len_test 53',
    'bair': '--ny 50 --nz 50 --n_euler_steps 2 --archi vgg --skipco --nt_cond 2 --nt_inf 2 --obs_scale 0.71 --batch_size 192 --dataset bair '
            '--nc 3 --seq_len 12 --lr_scheduling_burnin 1000000 --lr_scheduling_n_iter 500000',
}
# every option of the reference's args.py:28-165 with its default (None = required there)
REFERENCE_FLAGS = dict(seed=None, save_path='REQ', torch_amp=False, apex_amp=False, amp_opt_lvl='O1', keep_batchnorm_fp32=None, apex_verbose=False,
                       local_rank=0, device=None, n_workers=4, nhx=128, ny='REQ', nz='REQ', n_euler_steps=1, nt_inf='REQ', obs_scale=1,
                       archi='dcgan', skipco=False, nf=64, nh_res=512, nlayers_res=4, nh_inf=256, nlayers_inf=3, res_gain=1.41, beta_y=1,
                       beta_z=1, l2_res=1, batch_size=128, lr=0.0003, lr_scheduling_burnin=1000000, lr_scheduling_n_iter=100000,
                       dataset='REQ', data_dir='REQ', seq_len='REQ', ndigits=2, max_speed=4, deterministic=False, subsampling=8, nx=64,
                       nc='REQ', val_interval=20000, chkpt_interval=None, batch_size_test=16, n_iter_test=25, nt_cond='REQ', n_samples_test=100,
                       seq_len_test=None)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_train_accepts_every_reference_flag_and_readme_recipe():
    tr = _load('srvp_b200_train', os.path.join(ROOT, 'train.py'))
    p = tr.create_args()
    for name, recipe in README_RECIPES.items():
        opt = p.parse_args(shlex.split(recipe) + ['--data_dir', '/data', '--save_path', '/tmp/x', '--seed', '3', '--device', '0', '1',
                                                  '--n_workers', '8', '--torch_amp'])
        assert opt.ny in (20, 50) and opt.device == [0, 1] and opt.torch_amp
    opt = p.parse_args(shlex.split(README_RECIPES['bair']) + ['--save_path', '/tmp/x'])
    for flag, default in REFERENCE_FLAGS.items():
        assert hasattr(opt, flag), f'missing reference option --{flag}'
        if default not in ('REQ',) and flag not in ('local_rank',) and f'--{flag}' not in README_RECIPES['bair']:
            assert getattr(opt, flag) == default, (flag, getattr(opt, flag), default)
    ref_args = '/root/reference/args.py'
    if os.path.exists(ref_args):          # build container only: the list above is complete
        import re
        flags = set(re.findall(r"add\('--(\w+)'", open(ref_args).read()))
        assert flags == set(REFERENCE_FLAGS), flags ^ set(REFERENCE_FLAGS)


def test_test_py_accepts_reference_flags():
    te = _load('srvp_b200_test', os.path.join(ROOT, 'test.py'))
    opt = te.create_args().parse_args(shlex.split('--data_dir /d --xp_dir /x --lpips_dir /l --nt_gen 53 --n_samples 100 --batch_size 16 --fvd '
                                                  '--n_euler_steps 2 --nt_cond 8 --model_name model_best.pt --device 0 --test_seed 5'))
    assert opt.nt_gen == 53 and opt.fvd and opt.test_seed == 5 and opt.model_name == 'model_best.pt'


@pytest.mark.gpu
def test_train_validate_checkpoint_resume_and_test(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    tr = _load('srvp_b200_train', os.path.join(ROOT, 'train.py'))
    te = _load('srvp_b200_test', os.path.join(ROOT, 'test.py'))
    save = str(tmp_path / 'run')
    base = ('--ny 20 --nz 20 --n_euler_steps 2 --archi vgg --skipco --nt_cond 3 --nt_inf 2 --obs_scale 0.71 --batch_size 6 --dataset synthetic '
            '--nc 3 --seq_len 5 --seq_len_test 7 --seed 4 --val_interval 3 --chkpt_interval 3 --batch_size_test 2 --n_iter_test 2 --n_samples_test 4 '
            f'--sample_batch 3 --log_interval 1 --save_path {save}')
    opt = tr.create_args().parse_args(shlex.split(base + ' --lr_scheduling_burnin 4 --lr_scheduling_n_iter 2'))
    assert tr.main(opt) == 0
    for f in ('model.pt', 'model_3.pt', 'model_best.pt', 'state.pt', 'config.json'):
        assert os.path.exists(os.path.join(save, f)), f
    st = torch.load(os.path.join(save, 'state.pt'), map_location='cpu', weights_only=False)
    assert st['itr'] == 6 and st['best_val_metric'] is not None and len(st['optimizer']['state']) > 60
    # resume: two more iterations on top of the saved optimizer / scheduler / iteration state
    opt2 = tr.create_args().parse_args(shlex.split(base + f' --lr_scheduling_burnin 6 --lr_scheduling_n_iter 2 --resume {save}/state.pt'))
    assert tr.main(opt2) == 0
    st2 = torch.load(os.path.join(save, 'state.pt'), map_location='cpu', weights_only=False)
    assert st2['itr'] == 8
    k = next(iter(st2['optimizer']['state']))
    assert float(st2['optimizer']['state'][k]['step']) == 8.0
    # test.py on the trained run: batched rollouts and the reference's per-sample loop give metrics of the same distribution
    out = te.main(te.create_args().parse_args(shlex.split(f'--xp_dir {save} --nt_gen 9 --n_samples 4 --batch_size 2 --n_videos 4 --sample_batch 2')))
    assert out['psnr'].shape == (4,) and out['ssim'].shape == (4,) and (out['ssim'] <= 1).all() and (out['psnr'] > 0).all()
    out2 = te.main(te.create_args().parse_args(shlex.split(f'--xp_dir {save} --nt_gen 9 --n_samples 4 --batch_size 2 --n_videos 4 --sample_batch 0')))
    assert abs(float(out['psnr'].mean()) - float(out2['psnr'].mean())) < 2.0
    for f in ('results.npz', 'psnr_best.npz', 'ssim_worst.npz', 'random_1.npz'):
        assert os.path.exists(os.path.join(save, f)), f


def _reference_dir():
    for d in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.exists(os.path.join(d, 'train.py')) and os.path.exists(os.path.join(d, 'module', 'srvp.py')):
            return d
    return None


@pytest.mark.gpu
def test_reference_training_loop_drives_the_dropin_class():
    """`import module.srvp` -> `srvp_b200.module.srvp`. The reference's own train() and evaluate() run unmodified
    on top of this repository's class: finite, decreasing loss over a few steps, a finite validation metric."""
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    ref = _reference_dir()
    if ref is None:
        pytest.skip('reference tree not available (neither /root/reference nor baseline/_ref)')
    import srvp_b200.module as our_pkg
    import srvp_b200.module.srvp as our_srvp
    import srvp_b200.module.utils as our_utils
    saved = {k: sys.modules.get(k) for k in ('module', 'module.srvp', 'module.utils', 'configargparse', 'args', 'helper', 'data', 'data.base')}
    stub = types.ModuleType('configargparse')       # args.py only needs ArgumentParser with .add
    import argparse

    class _P(argparse.ArgumentParser):
        def add(self, *a, **k):
            return self.add_argument(*a, **k)
    stub.ArgumentParser = stub.ArgParser = _P
    stub.ArgumentDefaultsHelpFormatter = argparse.ArgumentDefaultsHelpFormatter
    sys.modules.update({'module': our_pkg, 'module.srvp': our_srvp, 'module.utils': our_utils, 'configargparse': stub})
    sys.path.insert(0, ref)
    try:
        for k in ('args', 'helper', 'data', 'data.base'):
            sys.modules.pop(k, None)
        ref_train = _load('reference_train', os.path.join(ref, 'train.py'))
        assert ref_train.srvp is our_srvp and ref_train.utils is our_utils
        helper = sys.modules['helper']
        opt = helper.DotDict(n_euler_steps=2, obs_scale=0.71, beta_y=1, beta_z=1, l2_res=1, torch_amp=False, apex_amp=False, nt_cond=3,
                             n_iter_test=1, n_samples_test=3)
        dev = torch.device('cuda')
        torch.manual_seed(2)
        model = our_srvp.StochasticLatentResidualVideoPredictor(64, 3, 64, 128, 20, 20, True, 2, 256, 3, 512, 4, 'vgg')
        model.init(res_gain=1.41)
        model.to(dev)
        optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)          # the reference's optimizer (train.py:289)
        batch = torch.rand(6, 8, 3, 64, 64, generator=torch.Generator().manual_seed(3))
        losses = []
        for _ in range(6):
            model.train()
            loss, nll, kl_y_0, kl_z = ref_train.train(model, optimizer, None, batch, dev, opt)
            losses.append(loss)
        assert all(l == l and abs(l) < 1e9 for l in losses), losses
        assert losses[-1] < losses[0], losses
        model.eval()
        # The reference's evaluate() indexes a CPU tensor with a CUDA index (train.py:181), which PyTorch >= 2 rejects on any CUDA run
        # whatever the model class; it is therefore run with that one index moved to the CPU (torch.arange(n).to(device) -> cpu).
        real_arange = torch.arange

        class _CpuIndex(torch.Tensor):
            pass

        def arange_cpu_to(*a, **k):
            t = real_arange(*a, **k)
            if not k and len(a) == 1 and isinstance(a[0], int):
                t = t.as_subclass(_CpuIndex)
            return t
        _CpuIndex.to = lambda self, *a, **k: (self.as_subclass(torch.Tensor) if (a and isinstance(a[0], torch.device) and not k)
                                              else torch.Tensor.to(self.as_subclass(torch.Tensor), *a, **k))
        ref_train.torch.arange = arange_cpu_to
        try:
            val = ref_train.evaluate(model, [torch.rand(7, 2, 3, 64, 64)], dev, opt)
        finally:
            ref_train.torch.arange = real_arange
        assert val == val and -60 < val < 0, val
        print('reference train(): losses', [round(l, 1) for l in losses], '; reference evaluate():', val)
    finally:
        sys.path.remove(ref)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
