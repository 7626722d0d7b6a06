"""GPU suite, part 4: edge configurations of the drop-in model against the live oracle (same weights, inputs and random draws):
no skip connections with VGG64, a single video, four Euler sub-steps per frame, one inference frame, odd batch sizes,
`remove_intermediate=False`. Tiny batches: the ELBO tolerance is 5e-4 (rounding noise ~ 1/sqrt(#pixels), see test_gpu_model.py) and the
per-pixel MSE tolerance 2e-4 (batch statistics over 3 frames amplify the bf16 rounding of the activations: 5.0e-5 measured for the
single-video case, < 1e-5 at the fixture sizes)."""
import pytest
import torch

from common import build_model, make_input, model_loss

pytestmark = pytest.mark.gpu
BASE = dict(nx=64, nf=64, nhx=128, nh_inf=256, nlayers_inf=3, nh_res=512, nlayers_res=4)
CASES = {
    'vgg_noskip_b1_os4': (dict(BASE, nc=1, ny=20, nz=20, skipco=False, nt_inf=1, archi='vgg'), 3, 1, 0.25),
    'dcgan_skip_nc1_b5': (dict(BASE, nc=1, ny=20, nz=20, skipco=True, nt_inf=2, archi='dcgan'), 4, 5, 1.0),
    'vgg_skip_b7_os1': (dict(BASE, nc=3, ny=50, nz=50, skipco=True, nt_inf=3, archi='vgg'), 3, 7, 1.0),
}
LOSS = dict(obs_scale=1.0, beta_y=1.0, beta_z=1.0, l2_res=1.0)


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    return 'cuda'


@pytest.mark.parametrize('name', list(CASES))
def test_edge_configuration_matches_oracle(dev, name):
    from oracle import srvp_oracle as O
    cfg, T, B, dt = CASES[name]
    m = build_model(cfg, 1.41, 5)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(dev).train()
    x = make_input(T, B, cfg['nc'], 77)
    torch.manual_seed(13)
    out = m(x.to(dev), T, dt=dt)
    loss = model_loss(out, x.to(dev), LOSS)[0]
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    torch.manual_seed(13)
    rnd = O.draw_randoms(cfg, T, T, B, training=True)
    with torch.no_grad():
        o = O.forward(sd, cfg, x, T, dt, rnd, training=True)
        ref = float(O.elbo(o, x, LOSS)[0])
    assert float(loss) == pytest.approx(ref, rel=5e-4)
    assert float(((out[0].detach().cpu() - o['x_']) ** 2).mean()) < 2e-4
    assert tuple(out[7].shape) == (int(round(1 / dt)) * (T - 1), B, cfg['ny'])


def test_generate_keeps_intermediate_states_on_request(dev):
    cfg, T, B, dt = CASES['vgg_noskip_b1_os4']
    m = build_model(cfg, 1.41, 5).to(dev).eval()
    x = make_input(T, 2, cfg['nc'], 3).to(dev)
    with torch.no_grad():
        torch.manual_seed(1)
        a = m(x, T, dt=dt)
        torch.manual_seed(1)
        b = m(x, T, dt=dt, remove_intermediate=False)
    assert a[1].shape[0] == T and b[1].shape[0] == 4 * (T - 1) + 1
    assert torch.equal(a[1], b[1][::4])
    assert b[0].shape[0] == b[1].shape[0]      # one decoded frame per kept state (srvp.py:460-466)
