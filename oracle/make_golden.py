"""Pins the oracle against the reference itself and writes tests/golden/*.pt (run in the build container only).

  python oracle/make_golden.py            # needs /root/reference (read-only); never runs on the GPU box

For every case: same seed -> reference model (module/srvp.py) and our parameter containers must produce identical
state-dicts; the reference forward/backward (train.py:88-119 loss) on CPU fp32 is compared with oracle/srvp_oracle.py
on identical weights, inputs and random draws; the fingerprints are stored as small fixtures.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
sys.path.insert(0, ROOT)

import torch.distributions as distrib  # noqa: E402
from oracle import srvp_oracle as O  # noqa: E402

CASES = {
    # BAIR hyper-parameters (reference README.md:127) at a small T / B
    'vgg_skip_nc3': dict(cfg=dict(nx=64, nc=3, nf=64, nhx=128, ny=50, nz=50, skipco=True, nt_inf=2, nh_inf=256, nlayers_inf=3,
                                  nh_res=512, nlayers_res=4, archi='vgg'),
                         T=4, B=3, dt=0.5, loss=dict(obs_scale=0.71, beta_y=1.0, beta_z=1.0, l2_res=1.0), res_gain=1.41),
    # KTH-like: single channel VGG with skips
    'vgg_skip_nc1': dict(cfg=dict(nx=64, nc=1, nf=64, nhx=128, ny=50, nz=50, skipco=True, nt_inf=3, nh_inf=256, nlayers_inf=3,
                                  nh_res=512, nlayers_res=4, archi='vgg'),
                         T=5, B=2, dt=0.5, loss=dict(obs_scale=0.2, beta_y=1.0, beta_z=1.0, l2_res=1.0), res_gain=1.2),
    # SM-MNIST hyper-parameters (README.md:111) at a small T / B
    'dcgan_nc1': dict(cfg=dict(nx=64, nc=1, nf=64, nhx=128, ny=20, nz=20, skipco=False, nt_inf=5, nh_inf=256, nlayers_inf=3,
                               nh_res=512, nlayers_res=4, archi='dcgan'),
                      T=6, B=2, dt=1.0, loss=dict(obs_scale=1.0, beta_y=1.0, beta_z=2.0, l2_res=1.0), res_gain=1.41),
    # DCGAN with skip connections and colour frames (the `--skipco` flag of the reference with its default architecture)
    'dcgan_skip_nc3': dict(cfg=dict(nx=64, nc=3, nf=64, nhx=128, ny=20, nz=20, skipco=True, nt_inf=2, nh_inf=256, nlayers_inf=3,
                                    nh_res=512, nlayers_res=4, archi='dcgan'),
                           T=5, B=3, dt=0.5, loss=dict(obs_scale=1.0, beta_y=1.0, beta_z=2.0, l2_res=1.0), res_gain=1.41),
}
ARG_ORDER = ['nx', 'nc', 'nf', 'nhx', 'ny', 'nz', 'skipco', 'nt_inf', 'nh_inf', 'nlayers_inf', 'nh_res', 'nlayers_res', 'archi']
SEED_MODEL, SEED_INPUT, SEED_FWD = 1, 123, 7


def build_reference(case):
    sys.path.insert(0, REF)
    import module.srvp as ref_srvp
    torch.manual_seed(SEED_MODEL)
    m = ref_srvp.StochasticLatentResidualVideoPredictor(*[case['cfg'][k] for k in ARG_ORDER])
    m.init(res_gain=case['res_gain'])
    return m


def build_ours(case):
    from srvp_b200.module.srvp import StochasticLatentResidualVideoPredictor
    torch.manual_seed(SEED_MODEL)
    m = StochasticLatentResidualVideoPredictor(*[case['cfg'][k] for k in ARG_ORDER])
    m.init(res_gain=case['res_gain'])
    return m


def make_input(case):
    g = torch.Generator().manual_seed(SEED_INPUT)
    return torch.rand(case['T'], case['B'], case['cfg']['nc'], 64, 64, generator=g)


def sd_checksum(sd):
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items() if v.dtype.is_floating_point}


def reference_loss(ref_utils, out, x, loss_cfg):
    """train.py:90-106 with the reference's own helpers."""
    x_, y, z, _, q_y_0_params, q_z_params, p_z_params, res = out
    n = x.shape[1]
    nll = ref_utils.neg_logprob(x_, x, scale=loss_cfg['obs_scale']).sum()
    q_y_0 = ref_utils.make_normal_from_raw_params(q_y_0_params)
    kl_y_0 = distrib.kl_divergence(q_y_0, distrib.Normal(0, 1)).sum()
    q_z, p_z = ref_utils.make_normal_from_raw_params(q_z_params), ref_utils.make_normal_from_raw_params(p_z_params)
    kl_z = distrib.kl_divergence(q_z, p_z).sum()
    loss = nll + loss_cfg['beta_y'] * kl_y_0 + loss_cfg['beta_z'] * kl_z
    if loss_cfg['l2_res'] > 0:
        loss = loss + loss_cfg['l2_res'] * torch.norm(res, p=2, dim=2).sum()
    return loss / n, nll, kl_y_0, kl_z


def maxdiff(a, b):
    return float((a - b).abs().max())


def run_case(name, case):
    import module.utils as ref_utils
    cfg = case['cfg']
    ref = build_reference(case)
    ours = build_ours(case)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys()), 'state-dict keys differ'
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_ours[k]), f'same-seed init differs at {k}'
    x = make_input(case)
    T, B = case['T'], case['B']
    sd0 = {k: v.clone() for k, v in sd_ref.items()}

    # ---- training-mode forward / backward of the reference
    ref.train()
    torch.manual_seed(SEED_FWD)
    out = ref(x, T, dt=case['dt'])
    loss, nll, kl_y, kl_z = reference_loss(ref_utils, out, x, case['loss'])
    loss.backward()
    grads = {k: p.grad.clone() for k, p in ref.named_parameters()}
    sd_after = {k: v.clone() for k, v in ref.state_dict().items()}

    # ---- oracle on the same weights / draws
    torch.manual_seed(SEED_FWD)
    rnd = O.draw_randoms(cfg, T, T, B, training=True)
    sdo = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in sd0.items()}
    stats = {}
    o = O.forward(sdo, cfg, x, T, case['dt'], rnd, training=True, stats_out=stats)
    o_loss, o_nll, o_kly, o_klz = O.elbo(o, x, case['loss'])
    o_loss.backward()
    names = ['x_', 'y', 'z', 'w', 'q_y_0_params', 'q_z_params', 'p_z_params', 'res']
    report = {}
    for i, n in enumerate(names):
        report[n] = maxdiff(out[i], o[n])
    report['loss_rel'] = abs(float(loss) - float(o_loss)) / abs(float(loss))
    gerr = 0.0
    for k, g in grads.items():
        e = maxdiff(g, sdo[k].grad) / (float(g.abs().max()) + 1e-12)
        gerr = max(gerr, e)
    report['grad_rel_max'] = gerr
    # running statistics update (momentum 0.1, unbiased variance)
    rs_err = 0.0
    for pfx, (mean, var) in stats.items():
        rm = 0.9 * sd0[pfx + '.running_mean'] + 0.1 * mean
        rv = 0.9 * sd0[pfx + '.running_var'] + 0.1 * var
        rs_err = max(rs_err, maxdiff(rm, sd_after[pfx + '.running_mean']), maxdiff(rv, sd_after[pfx + '.running_var']))
    report['running_stats'] = rs_err
    print(name, 'train', {k: f'{v:.2e}' for k, v in report.items()})
    assert all(v < 1e-6 for k, v in report.items() if k not in ('grad_rel_max',)), report
    assert report['grad_rel_max'] < 1e-4, report

    # ---- eval-mode forward with prediction beyond the conditioning frames (test.py:239-246 usage pattern)
    ref.eval()
    nt_cond, nt_pred = max(cfg['nt_inf'], T - 2), T + 2
    with torch.no_grad():
        torch.manual_seed(SEED_FWD)
        out_e = ref(x[:nt_cond], nt_pred, dt=case['dt'])
        torch.manual_seed(SEED_FWD)
        rnd_e = O.draw_randoms(cfg, nt_cond, nt_pred, B, training=False)
        o_e = O.forward(sd_after, cfg, x[:nt_cond], nt_pred, case['dt'], rnd_e, training=False)
    rep_e = {n: maxdiff(out_e[i], o_e[n]) for i, n in enumerate(names) if out_e[i] is not None}
    print(name, 'eval ', {k: f'{v:.2e}' for k, v in rep_e.items()})
    assert all(v < 1e-6 for v in rep_e.values()), rep_e

    def small(t):
        return t.detach().clone()

    gold = dict(
        name=name, cfg=cfg, T=T, B=B, dt=case['dt'], loss_cfg=case['loss'], res_gain=case['res_gain'],
        seeds=dict(model=SEED_MODEL, input=SEED_INPUT, fwd=SEED_FWD),
        weights_checksum=sd_checksum(sd0),
        train=dict(loss=float(loss), nll=float(nll), kl_y_0=float(kl_y), kl_z=float(kl_z),
                   y=small(out[1]), z=small(out[2]), w=small(out[3]), q_y_0_params=small(out[4]), q_z_params=small(out[5]),
                   p_z_params=small(out[6]), res=small(out[7]), hx=small(o['hx']),
                   x_mean=float(out[0].double().mean()), x_frame0=small(out[0][0, 0]), x_last=small(out[0][-1, -1]),
                   x_sub=small(out[0][:, :, :, ::8, ::8]),
                   grad_norm={k: float(g.double().norm()) for k, g in grads.items()},
                   grad_small={k: small(g) for k, g in grads.items() if g.numel() <= 4096},
                   running_after={k: small(v) for k, v in sd_after.items() if 'running' in k and v.numel() <= 128}),
        eval=dict(nt_cond=nt_cond, nt_pred=nt_pred, y=small(out_e[1]), z=small(out_e[2]), w=small(out_e[3]),
                  p_z_params=small(out_e[6]), x_mean=float(out_e[0].double().mean()), x_sub=small(out_e[0][:, :, :, ::8, ::8])),
        oracle_vs_reference=dict(train=report, eval=rep_e),
    )
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    path = os.path.join(ROOT, 'tests', 'golden', name + '.pt')
    torch.save(gold, path)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    torch.set_num_threads(8)
    sys.path.insert(0, REF)
    only = sys.argv[1:]            # optional: names of the cases to (re)generate
    for name, case in CASES.items():
        if not only or name in only:
            run_case(name, case)
