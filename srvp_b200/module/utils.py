"""Host-side helpers with the reference's names (reference: module/utils.py).

`make_normal_from_raw_params`, `rsample_normal` and `neg_logprob` are what train.py:92-98 calls on the model outputs;
they are thin torch.distributions wrappers (plumbing around the outputs, not part of the device hot path).
"""
import torch
import torch.distributions as distrib
import torch.nn as nn
import torch.nn.functional as F


def activation_factory(name):
    """Activation placeholder modules (reference: module/utils.py:24-48). They only keep Sequential indices aligned."""
    table = {'relu': lambda: nn.ReLU(inplace=True), 'leaky_relu': lambda: nn.LeakyReLU(0.2, inplace=True),
             'elu': lambda: nn.ELU(inplace=True), 'sigmoid': nn.Sigmoid, 'tanh': nn.Tanh}
    if name not in table:
        raise ValueError(f'Activation function \'{name}\' not yet implemented')
    return table[name]()


def init_weight(m, init_type='normal', init_gain=0.02):
    """Weight initialisation with the reference's rules (module/utils.py:51-85)."""
    kind = type(m).__name__
    if kind in ('Conv2d', 'ConvTranspose2d', 'Linear'):
        if init_type == 'normal':
            nn.init.normal_(m.weight.data, 0.0, init_gain)
        elif init_type == 'xavier':
            nn.init.xavier_normal_(m.weight.data, gain=init_gain)
        elif init_type == 'kaiming':
            nn.init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
        elif init_type == 'orthogonal':
            nn.init.orthogonal_(m.weight.data, gain=init_gain)
        else:
            raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
        if getattr(m, 'bias', None) is not None:
            nn.init.constant_(m.bias.data, 0.0)
    elif kind in ('BatchNorm2d', 'SyncBatchNorm'):
        if m.weight is not None:
            nn.init.normal_(m.weight.data, 1.0, init_gain)
        if m.bias is not None:
            nn.init.constant_(m.bias.data, 0.0)


def make_normal_from_raw_params(raw_params, scale_stddev=1, dim=-1, eps=1e-8):
    """Normal(loc, softplus(raw_scale) + eps) from concatenated raw parameters (module/utils.py:88-112)."""
    loc, raw_scale = torch.chunk(raw_params, 2, dim)
    assert loc.shape[dim] == raw_scale.shape[dim]
    return distrib.Normal(loc, (F.softplus(raw_scale) + eps) * scale_stddev)


def rsample_normal(raw_params, scale_stddev=1):
    """Reparameterised sample (module/utils.py:115-134); eps drawn in the params' dtype on the params' device."""
    return make_normal_from_raw_params(raw_params, scale_stddev=scale_stddev).rsample()


def neg_logprob(loc, data, scale=1):
    """Gaussian negative log-likelihood with fixed scale (module/utils.py:137-159)."""
    return -distrib.Normal(loc, scale).log_prob(data)
