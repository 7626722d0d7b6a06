"""Encoder / decoder parameter containers with the reference's module tree (reference: module/conv.py).

These classes only HOLD parameters and buffers with the reference's names, shapes, construction order and init, so that
`state_dict()` / `load_state_dict()` / same-seed initialisation / SyncBatchNorm conversion / optimizers are drop-in
compatible. The arithmetic is in srvp_b200/engine.py (hand-written CUDA through the C ABI);
calling `forward` on a container dispatches there.
"""
import torch.nn as nn

from . import utils


def encoder_factory(name, nx, nc, nh, nf):
    """module/conv.py:23-49"""
    if name == 'dcgan':
        return DCGAN64Encoder(nc, nh, nf)
    if name == 'vgg':
        return VGG64Encoder(nc, nh, nf)
    raise ValueError(f'No encoder named \'{name}\'')


def decoder_factory(name, nx, nc, ny, nf, skip=False):
    """module/conv.py:52-78"""
    if name == 'dcgan':
        return DCGAN64Decoder(nc, ny, nf, skip)
    if name == 'vgg':
        return VGG64Decoder(nc, ny, nf, skip)
    raise ValueError(f'No decoder named \'{name}\'')


def make_conv_block(conv, activation, bn=True):
    """conv -> [BatchNorm2d] -> [activation] (module/conv.py:81-107; note the order: BN before the activation)."""
    mods = [conv]
    if bn:
        mods.append(nn.BatchNorm2d(conv.out_channels))
    if activation != 'none':
        mods.append(utils.activation_factory(activation))
    return nn.Sequential(*mods)


def _c3(cin, cout):
    return make_conv_block(nn.Conv2d(cin, cout, 3, 1, 1, bias=False), activation='leaky_relu')


class BaseEncoder(nn.Module):
    archi = None

    def __init__(self, nh):
        super().__init__()
        self.nh = nh

    def forward(self, x, return_skip=False):
        from .. import engine
        return engine.encoder_forward_nchw(self, x, return_skip)


class DCGAN64Encoder(BaseEncoder):
    """module/conv.py:157-179"""
    archi = 'dcgan'

    def __init__(self, nc, nh, nf):
        super().__init__(nh)
        self.nc, self.nf = nc, nf
        chans = [nc, nf, nf * 2, nf * 4, nf * 8]
        self.conv = nn.ModuleList([
            make_conv_block(nn.Conv2d(chans[i], chans[i + 1], 4, 2, 1, bias=False), activation='leaky_relu', bn=i > 0)
            for i in range(4)])
        self.last_conv = make_conv_block(nn.Conv2d(nf * 8, nh, 4, 1, 0, bias=False), activation='tanh')


class VGG64Encoder(BaseEncoder):
    """module/conv.py:182-224"""
    archi = 'vgg'

    def __init__(self, nc, nh, nf):
        super().__init__(nh)
        self.nc, self.nf = nc, nf
        widths = [(nc, nf, 2), (nf, nf * 2, 2), (nf * 2, nf * 4, 3), (nf * 4, nf * 8, 3)]
        stages = []
        for i, (cin, cout, depth) in enumerate(widths):
            mods = [] if i == 0 else [nn.MaxPool2d(kernel_size=2, stride=2, padding=0)]
            mods += [_c3(cin if d == 0 else cout, cout) for d in range(depth)]
            stages.append(nn.Sequential(*mods))
        self.conv = nn.ModuleList(stages)
        self.last_conv = nn.Sequential(
            nn.MaxPool2d(kernel_size=2, stride=2, padding=0),
            make_conv_block(nn.Conv2d(nf * 8, nh, 4, 1, 0, bias=False), activation='tanh'))


class BaseDecoder(nn.Module):
    archi = None

    def __init__(self, ny, skip):
        super().__init__()
        self.ny = ny
        self.skip = skip

    def forward(self, z, skip=None, sigmoid=True):
        assert skip is None and not self.skip or self.skip and skip is not None
        from .. import engine
        return engine.decoder_forward_nchw(self, z, skip, sigmoid)


class DCGAN64Decoder(BaseDecoder):
    """module/conv.py:278-305"""
    archi = 'dcgan'

    def __init__(self, nc, ny, nf, skip):
        super().__init__(ny, skip)
        self.nc, self.nf = nc, nf
        coef = 2 if skip else 1
        self.first_upconv = make_conv_block(nn.ConvTranspose2d(ny, nf * 8, 4, 1, 0, bias=False), activation='leaky_relu')
        self.conv = nn.ModuleList([
            make_conv_block(nn.ConvTranspose2d(nf * 8 * coef, nf * 4, 4, 2, 1, bias=False), activation='leaky_relu'),
            make_conv_block(nn.ConvTranspose2d(nf * 4 * coef, nf * 2, 4, 2, 1, bias=False), activation='leaky_relu'),
            make_conv_block(nn.ConvTranspose2d(nf * 2 * coef, nf, 4, 2, 1, bias=False), activation='leaky_relu'),
            nn.ConvTranspose2d(nf * coef, nc, 4, 2, 1, bias=False),
        ])


class VGG64Decoder(BaseDecoder):
    """module/conv.py:308-355"""
    archi = 'vgg'

    def __init__(self, nc, ny, nf, skip):
        super().__init__(ny, skip)
        self.nc, self.nf = nc, nf
        coef = 2 if skip else 1
        up = lambda: nn.Upsample(scale_factor=2, mode='nearest')
        self.first_upconv = nn.Sequential(
            make_conv_block(nn.ConvTranspose2d(ny, nf * 8, 4, 1, 0, bias=False), activation='leaky_relu'), up())
        self.conv = nn.ModuleList([
            nn.Sequential(_c3(nf * 8 * coef, nf * 8), _c3(nf * 8, nf * 8), _c3(nf * 8, nf * 4), up()),
            nn.Sequential(_c3(nf * 4 * coef, nf * 4), _c3(nf * 4, nf * 4), _c3(nf * 4, nf * 2), up()),
            nn.Sequential(_c3(nf * 2 * coef, nf * 2), _c3(nf * 2, nf), up()),
            nn.Sequential(_c3(nf * coef, nf), nn.ConvTranspose2d(nf, nc, 3, 1, 1, bias=False)),
        ])
