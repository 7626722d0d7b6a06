"""Shared helpers for the tests: case construction identical to oracle/make_golden.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ARG_ORDER = ['nx', 'nc', 'nf', 'nhx', 'ny', 'nz', 'skipco', 'nt_inf', 'nh_inf', 'nlayers_inf', 'nh_res', 'nlayers_res', 'archi']
GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + '.pt'), weights_only=False)


def build_model(cfg, res_gain, seed):
    from srvp_b200.module.srvp import StochasticLatentResidualVideoPredictor
    torch.manual_seed(seed)
    m = StochasticLatentResidualVideoPredictor(*[cfg[k] for k in ARG_ORDER])
    m.init(res_gain=res_gain)
    return m


def make_input(T, B, nc, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(T, B, nc, 64, 64, generator=g)


def model_loss(out, x, loss_cfg):
    """train.py:90-106 on our model's outputs with our module.utils helpers."""
    import torch.distributions as distrib
    from srvp_b200.module import utils
    x_, y, z, _, q_y_0_params, q_z_params, p_z_params, res = out
    n = x.shape[1]
    nll = utils.neg_logprob(x_, x, scale=loss_cfg['obs_scale']).sum()
    q_y_0 = utils.make_normal_from_raw_params(q_y_0_params)
    kl_y_0 = distrib.kl_divergence(q_y_0, distrib.Normal(0, 1)).sum()
    q_z, p_z = utils.make_normal_from_raw_params(q_z_params), utils.make_normal_from_raw_params(p_z_params)
    kl_z = distrib.kl_divergence(q_z, p_z).sum()
    loss = nll + loss_cfg['beta_y'] * kl_y_0 + loss_cfg['beta_z'] * kl_z
    if loss_cfg['l2_res'] > 0:
        loss = loss + loss_cfg['l2_res'] * torch.norm(res, p=2, dim=2).sum()
    return loss / n, nll, kl_y_0, kl_z


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
