"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the SRVP hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.

It restates, as plain functions over a state-dict (no nn.Module, explicit random draws), the algorithm of the reference:
  encoder / decoder stacks       module/conv.py:129-154, :157-224, :249-275, :278-355
  encode / skip selection        module/srvp.py:156-193
  infer_w / infer_y / infer_z    module/srvp.py:229-298
  generate (Euler residual loop) module/srvp.py:300-413
  decode / forward               module/srvp.py:195-227, :415-470
  Gaussian helpers               module/utils.py:88-159
  ELBO assembly                  train.py:88-106
The arithmetic itself lives in PyTorch (third-party; reference pins torch==1.4.0, this image has torch 2.11 CPU): conv2d,
conv_transpose2d, batch_norm, max_pool2d, interpolate(nearest), linear, LSTM cell equations (restated below), softplus.

PINNING: the reference has no tests or golden vectors. The oracle is pinned by running the reference
itself: oracle/make_golden.py imports /root/reference in the build container, checks this file against it on identical
weights / inputs / random draws, and writes tests/golden/*.pt, which tests/test_oracle.py replays without the reference.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU = 0.2
# True: run the LSTM through torch's fused CPU kernel (the arithmetic nn.LSTM itself uses -> bit-exact pinning against the
# reference); False: the explicit cell equations below (differs by ~1e-7, which fp32 chaos through LeakyReLU / max-pool kinks
# amplifies to ~6e-3 of max|g| on single weight-gradient entries -- measured by oracle/make_golden.py).
USE_ATEN_LSTM = True
BN_EPS = 1e-5


# ---------------------------------------------------------------------------------------------- random draws
def draw_randoms(cfg, nt_x, nt, bsz, training, n_obs=None):
    """Draws, from torch's default CPU generator and in the reference's consumption order:
    skip frame per video (srvp.py:185), nt_inf random frames per video (srvp.py:246), y_0 noise (srvp.py:277),
    one z noise per generated frame (srvp.py:387/392 via utils.py:133)."""
    r = {}
    if training:
        if cfg['skipco']:
            r['t_skip'] = torch.randint(nt_x, size=(bsz,))
        r['t_w'] = torch.stack([torch.randperm(nt_x)[:cfg['nt_inf']] for _ in range(bsz)], 1)
    r['eps_y'] = torch.empty(bsz, cfg['ny']).normal_()
    r['eps_z'] = [torch.empty(bsz, cfg['nz']).normal_() for _ in range(nt - 1)]
    return r


# ---------------------------------------------------------------------------------------------- building blocks
# EMULATE_BF16: when True, convolution / GEMM operands and raw conv outputs of the encoder and decoder are rounded to bf16
# (straight-through gradient), i.e. the storage precision of the CUDA path. Tests use it to separate rounding-induced
# deviations (which fp32 chaos through LeakyReLU / max-pool kinks amplifies) from logic errors. Default: exact fp32.
EMULATE_BF16 = False


def _rq(t):
    if not EMULATE_BF16:
        return t
    return t + (t.detach().to(torch.bfloat16).float() - t.detach())


def _bn(h, sd, prefix, training, stats_out=None):
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    if training:
        # running buffers are passed (on clones) so that the same CPU batch-norm kernel as nn.BatchNorm2d is selected
        rm, rv = sd[prefix + '.running_mean'].detach().clone(), sd[prefix + '.running_var'].detach().clone()
        out = F.batch_norm(h, rm, rv, w, b, True, 0.1, BN_EPS)
        if stats_out is not None:
            dims = [0, 2, 3]
            stats_out[prefix] = (h.mean(dims).detach(), h.var(dims, unbiased=True).detach())
        return out
    return F.batch_norm(h, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], w, b, False, 0.0, BN_EPS)


def _block(h, sd, prefix, training, stats_out, stride=1, pad=1, act='lrelu', bn=True, transposed=False):
    w = _rq(sd[prefix + '.0.weight'])
    h = _rq(h)
    h = F.conv_transpose2d(h, w, None, stride, pad) if transposed else F.conv2d(h, w, None, stride, pad)
    if bn and not (act == 'tanh'):
        h = _rq(h)  # raw conv outputs are stored as bf16 (the 1x1 encoder head stays fp32)
    if bn:
        h = _bn(h, sd, prefix + '.1', training, stats_out)
    if act == 'lrelu':
        h = F.leaky_relu(h, LRELU)
    elif act == 'tanh':
        h = torch.tanh(h)
    return h


VGG_ENC = [[(0, 0), (0, 1)], [(1, 1), (1, 2)], [(2, 1), (2, 2), (2, 3)], [(3, 1), (3, 2), (3, 3)]]
VGG_DEC = [[(0, 0), (0, 1), (0, 2)], [(1, 0), (1, 1), (1, 2)], [(2, 0), (2, 1)], [(3, 0)]]


def encoder(sd, cfg, x_flat, training, stats_out=None):
    """x_flat (N, nc, 64, 64) -> h (N, nhx), skips deepest first (conv.py:129-154)."""
    h, skips = x_flat, []
    if cfg['archi'] == 'vgg':
        for i, stage in enumerate(VGG_ENC):
            if i > 0:
                h = F.max_pool2d(h, 2, 2)
            for (a, b) in stage:
                h = _block(h, sd, f'encoder.conv.{a}.{b}', training, stats_out)
            skips.append(h)
        h = F.max_pool2d(h, 2, 2)
        h = _block(h, sd, 'encoder.last_conv.1', training, stats_out, stride=1, pad=0, act='tanh')
    else:
        for i in range(4):
            h = _block(h, sd, f'encoder.conv.{i}', training, stats_out, stride=2, pad=1, bn=i > 0)
            skips.append(h)
        h = _block(h, sd, 'encoder.last_conv', training, stats_out, stride=1, pad=0, act='tanh')
    return h.reshape(-1, cfg['nhx']), skips[::-1]


def decoder(sd, cfg, z, skip, training, stats_out=None):
    """z (N, nh_inf + ny), skip list deepest first or None -> x_hat (N, nc, 64, 64) (conv.py:249-275)."""
    h = z.reshape(*z.shape, 1, 1)
    if cfg['archi'] == 'vgg':
        h = _block(h, sd, 'decoder.first_upconv.0', training, stats_out, stride=1, pad=0, transposed=True)
        h = F.interpolate(h, scale_factor=2, mode='nearest')
        for i, stage in enumerate(VGG_DEC):
            if skip is not None:
                h = torch.cat([h, skip[i]], 1)
            for (a, b) in stage:
                h = _block(h, sd, f'decoder.conv.{a}.{b}', training, stats_out)
            if i < 3:
                h = F.interpolate(h, scale_factor=2, mode='nearest')
        h = F.conv_transpose2d(_rq(h), _rq(sd['decoder.conv.3.1.weight']), None, 1, 1)
    else:
        h = _block(h, sd, 'decoder.first_upconv', training, stats_out, stride=1, pad=0, transposed=True)
        for i in range(3):
            if skip is not None:
                h = torch.cat([h, skip[i]], 1)
            h = _block(h, sd, f'decoder.conv.{i}', training, stats_out, stride=2, pad=1, transposed=True)
        if skip is not None:
            h = torch.cat([h, skip[3]], 1)
        h = F.conv_transpose2d(_rq(h), _rq(sd['decoder.conv.3.weight']), None, 2, 1)
    return torch.sigmoid(h)


def mlp(sd, prefix, x, n_layers):
    """Linear -> (ReLU -> Linear)* (mlp.py:47-90)."""
    for i in range(n_layers):
        if i > 0:
            x = F.relu(x)
        key = f'{prefix}.module.{i}.{0 if i == 0 else 1}'
        x = F.linear(x, sd[key + '.weight'], sd[key + '.bias'])
    return x


def lstm(sd, hx):
    """Single-layer LSTM, zero initial state, gate order (i, f, g, o) (srvp.py:132, :366; torch.nn.LSTM semantics)."""
    wi, wh, bi, bh = sd['inf_z.weight_ih_l0'], sd['inf_z.weight_hh_l0'], sd['inf_z.bias_ih_l0'], sd['inf_z.bias_hh_l0']
    nh = wh.shape[1]
    if USE_ATEN_LSTM:  # same fused CPU kernel as nn.LSTM: makes the pinning against the reference bit-exact
        z0 = hx.new_zeros(1, hx.shape[1], nh)
        # train flag: only dropout (0.0 here) depends on it on CPU; cuDNN needs it set to run the backward pass (bench.py's eager-GPU leg)
        return torch._VF.lstm(hx, (z0, z0), [wi, wh, bi, bh], True, 1, 0.0, hx.is_cuda and torch.is_grad_enabled(), False, False)[0]
    h = hx.new_zeros(hx.shape[1], nh)
    c = hx.new_zeros(hx.shape[1], nh)
    out = []
    for t in range(hx.shape[0]):
        g = F.linear(hx[t], wi, bi) + F.linear(h, wh, bh)
        i_, f_, g_, o_ = g.chunk(4, 1)
        c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
        h = torch.sigmoid(o_) * torch.tanh(c)
        out.append(h)
    return torch.stack(out)


def sample(raw, eps):
    """loc + eps * (softplus(raw_scale) + 1e-8) (utils.py:88-134)."""
    loc, raw_scale = raw.chunk(2, -1)
    return loc + eps * (F.softplus(raw_scale) + 1e-8)


# ---------------------------------------------------------------------------------------------- model
def forward(sd, cfg, x, nt, dt, rnd, training, stats_out=None):
    """Restates StochasticLatentResidualVideoPredictor.forward (srvp.py:415-470) with explicit random draws `rnd`."""
    nt_x, bsz = x.shape[0], x.shape[1]
    hx_flat, skips = encoder(sd, cfg, x.reshape(nt_x * bsz, *x.shape[2:]), training, stats_out)
    hx = hx_flat.view(nt_x, bsz, cfg['nhx'])
    ar = torch.arange(bsz)
    if cfg['skipco']:
        t_skip = rnd['t_skip'] if training else torch.full((bsz,), nt_x - 1, dtype=torch.long)
        skips = [s.view(nt_x, bsz, *s.shape[1:])[t_skip, ar] for s in skips]
    else:
        skips = None
    # infer_w (srvp.py:229-256)
    if training:
        h = hx[rnd['t_w'].reshape(-1), ar.repeat(cfg['nt_inf'], 1).reshape(-1)].view(cfg['nt_inf'], bsz, cfg['nhx'])
    else:
        h = hx[-cfg['nt_inf']:]
    h = F.relu(F.linear(h, sd['w_proj.0.weight'], sd['w_proj.0.bias'])).sum(0)
    w = torch.tanh(F.linear(h, sd['w_inf.0.weight'], sd['w_inf.0.bias']))
    # infer_y (srvp.py:258-278)
    q_y0 = mlp(sd, 'q_y', hx[:cfg['nt_inf']].permute(1, 0, 2).reshape(bsz, cfg['nt_inf'] * cfg['nhx']), cfg['nlayers_inf'])
    y0 = sample(q_y0, rnd['eps_y'])
    # generate (srvp.py:325-413)
    hx_z = lstm(sd, hx)
    osamp = int(1 / dt)
    ys, zs, qz, pz, res = [y0], [], [], [], []
    y_tm1, t_data = y0, 0
    for t in np.linspace(dt, nt - 1, osamp * (nt - 1)):
        prev, t_data = t_data, int(math.ceil(t))
        if t_data != prev:
            p = mlp(sd, 'p_z', y_tm1, cfg['nlayers_res'])
            pz.append(p)
            if t_data < nt_x:
                q = F.linear(hx_z[t_data], sd['q_z.weight'], sd['q_z.bias'])
                qz.append(q)
                z_t = sample(q, rnd['eps_z'][t_data - 1])
            else:
                assert not training
                z_t = sample(p, rnd['eps_z'][t_data - 1])
            zs.append(z_t)
        r = dt * mlp(sd, 'dynamics', torch.cat([y_tm1, zs[-1]], 1), cfg['nlayers_res'])
        y_tm1 = y_tm1 + r
        if float(t).is_integer():
            ys.append(y_tm1)
        res.append(r)
    y = torch.stack(ys)
    # decode (srvp.py:195-227)
    n = y.shape[0]
    dec_inp = torch.cat([w.repeat(n, 1, 1).view(n * bsz, -1), y.reshape(n * bsz, cfg['ny'])], 1)
    sk = None
    if skips is not None:
        sk = [s.expand(n, *s.shape).reshape(n * bsz, *s.shape[1:]) for s in skips]
    x_ = decoder(sd, cfg, dec_inp, sk, training, stats_out).view(n, bsz, *x.shape[2:])
    return dict(x_=x_, y=y, z=torch.stack(zs), w=w, q_y_0_params=q_y0, q_z_params=torch.stack(qz) if qz else None,
                p_z_params=torch.stack(pz), res=torch.stack(res), hx=hx)


def kl_normal(loc_q, scale_q, loc_p, scale_p):
    """KL(N(q) || N(p)), closed form as torch.distributions.kl._kl_normal_normal."""
    var_ratio = (scale_q / scale_p) ** 2
    t1 = ((loc_q - loc_p) / scale_p) ** 2
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


def elbo(out, x, cfg_loss):
    """train.py:90-106: (loss, nll, kl_y_0, kl_z), loss batch-averaged, the others summed."""
    s = cfg_loss['obs_scale']
    n = x.shape[1]
    nll = (((x - out['x_']) ** 2) / (2 * s * s) + math.log(s) + 0.5 * math.log(2 * math.pi)).sum()
    ly, ry = out['q_y_0_params'].chunk(2, -1)
    kl_y = kl_normal(ly, F.softplus(ry) + 1e-8, torch.zeros_like(ly), torch.ones_like(ly)).sum()
    lq, rq = out['q_z_params'].chunk(2, -1)
    lp, rp = out['p_z_params'].chunk(2, -1)
    kl_z = kl_normal(lq, F.softplus(rq) + 1e-8, lp, F.softplus(rp) + 1e-8).sum()
    loss = nll + cfg_loss['beta_y'] * kl_y + cfg_loss['beta_z'] * kl_z
    if cfg_loss['l2_res'] > 0:
        loss = loss + cfg_loss['l2_res'] * torch.norm(out['res'], p=2, dim=2).sum()
    return loss / n, nll, kl_y, kl_z
