"""MLP parameter container with the reference's state-dict layout (reference: module/mlp.py:21-90)."""
import torch.nn as nn

from . import utils


def make_lin_block(n_inp, n_out, activation):
    """[activation ->] Linear, as an nn.Sequential (module/mlp.py:21-44)."""
    mods = [] if activation == 'none' else [utils.activation_factory(activation)]
    mods.append(nn.Linear(n_inp, n_out))
    return nn.Sequential(*mods)


class MLP(nn.Module):
    """Linear -> (ReLU -> Linear) x (n_layers - 1). Keys: module.0.0, module.1.1, ... (module/mlp.py:47-90)."""

    def __init__(self, n_inp, n_hid, n_out, n_layers, activation='relu'):
        super().__init__()
        assert n_hid == 0 or n_layers > 1
        self.module = nn.Sequential(*[
            make_lin_block(n_inp if i == 0 else n_hid, n_out if i == n_layers - 1 else n_hid, activation if i > 0 else 'none')
            for i in range(n_layers)])

    def linears(self):
        return [blk[-1] for blk in self.module]

    def forward(self, x):
        # Host/torch path used only by CPU-side tooling; the training hot path runs the fused latent kernels.
        return self.module(x)
