"""GPU suite, part 3: the 4x4 / stride-2 / pad-1 convolution family of DCGAN64 (reference module/conv.py:157-179, :278-305) through
the C ABI against torch on identical bf16-rounded operands, and the DCGAN64 model against the golden fixture / the live CPU oracle.

Tolerances as in test_gpu_kernels.py: bf16 outputs 1.5e-2 of the tensor maximum, fp32 contractions 5e-3, reductions 1e-3.
"""
import pytest
import torch
import torch.nn.functional as F

from common import build_model, load_golden, make_input, model_loss, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    torch.manual_seed(0)
    return 'cuda'


def bf(x):
    return x.to(torch.bfloat16).float()


def maxerr(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-30))


def test_s2d_layout_kernels(dev):
    from srvp_b200 import ops
    for nc in (1, 3):
        x = torch.rand(5, nc, 64, 64, device=dev)
        s = ops.nchw_to_s2d_bf16(x, 16).float()
        ref = torch.zeros(5, 32, 32, 16, device=dev)
        for py in range(2):
            for px in range(2):
                ref[..., (py * 2 + px) * nc:(py * 2 + px + 1) * nc] = bf(x[:, :, py::2, px::2]).permute(0, 2, 3, 1)
        assert torch.equal(s, ref)
        dx = torch.randn(5, nc, 64, 64, device=dev)
        xh = torch.rand(5, nc, 64, 64, device=dev)
        d = ops.sigmoid_bwd_s2d(dx, xh).float()
        full = dx * xh * (1 - xh)
        ref = torch.zeros(5, 32, 32, 16, device=dev)
        for py in range(2):
            for px in range(2):
                ref[..., (py * 2 + px) * nc:(py * 2 + px + 1) * nc] = bf(full[:, :, py::2, px::2]).permute(0, 2, 3, 1)
        assert maxerr(d, ref) < 1e-2


@pytest.mark.parametrize('cin,cout,res', [(64, 128, 16), (128, 256, 8), (256, 512, 4)])
def test_down_conv_family(dev, cin, cout, res):
    """Conv2d(cin, cout, 4, 2, 1) forward with fused BN + LeakyReLU input, BN statistics, weight gradient and data gradient."""
    from srvp_b200 import ops, engine, _lib
    F_ = 7
    z = torch.randn(F_, 2 * res, 2 * res, cin, device=dev).to(torch.bfloat16)
    sc, sh = torch.rand(cin, device=dev) + 0.5, torch.randn(cin, device=dev) * 0.3
    w = torch.randn(cout, cin, 4, 4, device=dev) * 0.05
    a_ref = bf(F.leaky_relu(z.float() * sc + sh, 0.2)).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    w_ref = bf(w).requires_grad_(True)
    out_ref = F.conv2d(a_ref, w_ref, None, 2, 1)
    srcs = engine._s2d_sources(z, cin, sc, sh)
    wp = ops.pack_conv4x4s2(w, _lib.W4_DOWN, cout, cin, cin * 16, 16)
    out, partial, a_out = ops.conv3x3(srcs, wp, F_, res, res, cout, stats=True, save_input=True, tap_masks=engine._down_masks(cin, 4 * cin), taps=4)
    assert maxerr(out.permute(0, 3, 1, 2), out_ref) < 1.5e-2
    st = partial.sum(0)
    o = out.float()
    assert torch.allclose(st[:, 0], o.sum((0, 1, 2)), rtol=1e-3, atol=1e-2 * float(o.abs().sum((0, 1, 2)).max()))
    assert torch.allclose(st[:, 1], (o * o).sum((0, 1, 2)), rtol=1e-3)
    # the loader's copy of the operand is the space-to-depth image of the activated input
    a_s2d = torch.cat([a_ref.detach()[:, :, py::2, px::2] for py in range(2) for px in range(2)], 1).permute(0, 2, 3, 1)
    assert maxerr(a_out, a_s2d) < 1e-2
    # gradients
    dz = torch.randn(F_, res, res, cout, device=dev).to(torch.bfloat16)
    out_ref.backward(dz.float().permute(0, 3, 1, 2))
    dw = torch.zeros_like(w)
    ops.wgrad3x3(a_out, 4 * cin, dz, cout, F_, res, res, cout, 4 * cin, dw, 'conv', map4=_lib.W4_DOWN, phase_channels=cin, strides=(cin * 16, 16))
    assert maxerr(dw, w_ref.grad) < 5e-3
    da = torch.empty(F_, 2 * res, 2 * res, cin, dtype=torch.bfloat16, device=dev)
    engine._up_conv([ops.Src(dz, cout)], w, 16, cin * 16, F_, res, res, cin, cout, stats=False, a_out=None, out=da)
    assert maxerr(da.permute(0, 3, 1, 2), a_ref.grad) < 1.5e-2


@pytest.mark.parametrize('cins,cout,res', [([512], 256, 4), ([256, 256], 128, 8), ([128], 64, 16)])
def test_up_conv_family(dev, cins, cout, res):
    """ConvTranspose2d(cin, cout, 4, 2, 1) forward (optionally on cat[h, skip]), BN statistics, weight gradient, data gradient."""
    from srvp_b200 import ops, engine, _lib
    F_ = 6
    cin = sum(cins)
    srcs, refs = [], []
    for c in cins:
        zz = torch.randn(F_, res, res, c, device=dev).to(torch.bfloat16)
        sc, sh = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.3
        srcs.append(ops.Src(zz, c, sc, sh, None, 0, 0, True))
        refs.append(bf(F.leaky_relu(zz.float() * sc + sh, 0.2)).permute(0, 3, 1, 2))
    a_ref = torch.cat(refs, 1).contiguous().requires_grad_(True)
    w = torch.randn(cin, cout, 4, 4, device=dev) * 0.05
    w_ref = bf(w).requires_grad_(True)
    out_ref = F.conv_transpose2d(a_ref, w_ref, None, 2, 1)
    out = torch.empty(F_, 2 * res, 2 * res, cout, dtype=torch.bfloat16, device=dev)
    a_out = torch.empty(F_, res, res, cin, dtype=torch.bfloat16, device=dev)
    partial = engine._up_conv(srcs, w, 16, cout * 16, F_, res, res, cout, cin, stats=True, a_out=a_out, out=out)
    assert maxerr(out.permute(0, 3, 1, 2), out_ref) < 1.5e-2
    assert maxerr(a_out.permute(0, 3, 1, 2), a_ref.detach()) < 1e-2
    st, o = partial.sum(0), out.float()
    assert torch.allclose(st[:, 0], o.sum((0, 1, 2)), rtol=1e-3, atol=1e-2 * float(o.abs().sum((0, 1, 2)).max()))
    assert torch.allclose(st[:, 1], (o * o).sum((0, 1, 2)), rtol=1e-3)
    # backward through BN(train) + LeakyReLU of THIS layer, written as the space-to-depth image, then both gradients
    dzf = torch.randn(F_, 2 * res, 2 * res, cout, device=dev).to(torch.bfloat16)
    dzs = torch.cat([dzf[:, py::2, px::2, :] for py in range(2) for px in range(2)], 3).contiguous()
    out_ref.backward(dzf.float().permute(0, 3, 1, 2))
    dw = torch.zeros_like(w)
    ops.wgrad3x3(dzs, 4 * cout, a_out, cin, F_, res, res, cin, 4 * cout, dw, 'conv', map4=_lib.W4_DOWN, phase_channels=cout, strides=(cout * 16, 16))
    assert maxerr(dw, w_ref.grad) < 5e-3
    wp = ops.pack_conv4x4s2(w, _lib.W4_DOWN, cin, cout, cout * 16, 16)
    da, _ = ops.conv3x3([ops.Src(dzs, 4 * cout)], wp, F_, res, res, cin, tap_masks=engine._down_masks(cout, 4 * cout), taps=4)
    assert maxerr(da.permute(0, 3, 1, 2), a_ref.grad) < 1.5e-2


def test_bn_bwd_space_to_depth_output(dev):
    from srvp_b200 import ops
    from srvp_b200.ops import BNState
    F_, H, C = 5, 8, 64
    z = torch.randn(F_, H, H, C, device=dev).to(torch.bfloat16)
    da = torch.randn(F_, H, H, C, device=dev).to(torch.bfloat16)
    st = BNState(C, dev)
    zf = z.float()
    mean, var = zf.mean((0, 1, 2)), zf.var((0, 1, 2), unbiased=False)
    gamma = torch.rand(C, device=dev) + 0.5
    st.mean.copy_(mean); st.invstd.copy_((var + 1e-5).rsqrt()); st.scale.copy_(gamma * st.invstd); st.shift.copy_(-mean * st.scale)
    g1, b1, g2, b2 = (torch.zeros(C, device=dev) for _ in range(4))
    dense = ops.bn_bwd(z, st, gamma, g1, b1, da, 0, F_, H, H, C)
    s2d = ops.bn_bwd(z, st, gamma, g2, b2, da, 0, F_, H, H, C, g_s2d=True)
    ref = torch.cat([dense[:, py::2, px::2, :] for py in range(2) for px in range(2)], 3)
    assert torch.equal(s2d, ref) and torch.equal(g1, g2)


@pytest.mark.parametrize('nc,cins', [(1, [64]), (3, [64, 64])])
def test_last_layer_family(dev, nc, cins):
    """ConvTranspose2d(cin, nc, 4, 2, 1) + sigmoid -> NCHW fp32, its weight gradient and data gradient."""
    from srvp_b200 import ops, _lib
    F_, res = 5, 32
    cin = sum(cins)
    srcs, refs = [], []
    for c in cins:
        zz = torch.randn(F_, res, res, c, device=dev).to(torch.bfloat16)
        sc, sh = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.3
        srcs.append(ops.Src(zz, c, sc, sh, None, 0, 0, True))
        refs.append(bf(F.leaky_relu(zz.float() * sc + sh, 0.2)).permute(0, 3, 1, 2))
    a_ref = torch.cat(refs, 1).contiguous().requires_grad_(True)
    w = torch.randn(cin, nc, 4, 4, device=dev) * 0.05
    w_ref = bf(w).requires_grad_(True)
    pre = F.conv_transpose2d(a_ref, w_ref, None, 2, 1)
    x_ref = torch.sigmoid(pre)
    wp = ops.pack_conv4x4s2(w, _lib.W4_UP_ALL, nc, cin, 16, nc * 16)
    x_hat, _, a_out = ops.conv3x3(srcs, wp, F_, res, res, 4 * nc, sigmoid_nchw=True, sigmoid_d2s=True, save_input=True, taps=4)
    assert x_hat.shape == (F_, nc, 64, 64)
    assert float((x_hat - x_ref).abs().max()) < 2e-3
    d_x = torch.randn(F_, nc, 64, 64, device=dev)
    dz16 = ops.sigmoid_bwd_s2d(d_x, x_hat)
    dpre = bf(d_x * x_hat * (1 - x_hat))
    pre.backward(dpre)
    dw = torch.zeros_like(w)
    ops.wgrad3x3(a_out, cin, dz16, 16, F_, res, res, 4 * nc, cin, dw, 'convT', map4=_lib.W4_UP_ALL, phase_channels=nc, strides=(16, nc * 16))
    assert maxerr(dw, w_ref.grad) < 5e-3
    da = torch.empty(F_, res, res, cin, dtype=torch.bfloat16, device=dev)
    for n0 in range(0, cin, 64):
        wpd = ops.pack_conv4x4s2(w, _lib.W4_DOWN, 64, nc, nc * 16, 16, n_offset=n0)
        ops.conv3x3([ops.Src(dz16, 16)], wpd, F_, res, res, 64, out=da, out_cpitch=cin, out_coff=n0, taps=4)
    assert maxerr(da.permute(0, 3, 1, 2), a_ref.grad) < 1.5e-2


def test_first_layer_family(dev):
    """Conv2d(nc, 64, 4, 2, 1) on the space-to-depth input image + its weight gradient (no data gradient: the input is the video)."""
    from srvp_b200 import ops, _lib
    for nc in (1, 3):
        F_ = 6
        x = torch.rand(F_, nc, 64, 64, device=dev)
        w = torch.randn(64, nc, 4, 4, device=dev) * 0.2
        w_ref = bf(w).requires_grad_(True)
        out_ref = F.conv2d(bf(x), w_ref, None, 2, 1)
        xs = ops.nchw_to_s2d_bf16(x, 16)
        wp = ops.pack_conv4x4s2(w, _lib.W4_DOWN, 64, nc, nc * 16, 16)
        out, _ = ops.conv3x3([ops.Src(xs, 16)], wp, F_, 32, 32, 64, taps=4)
        assert maxerr(out.permute(0, 3, 1, 2), out_ref) < 1.5e-2
        dz = torch.randn(F_, 32, 32, 64, device=dev).to(torch.bfloat16)
        out_ref.backward(dz.float().permute(0, 3, 1, 2))
        dw = torch.zeros_like(w)
        ops.wgrad3x3(xs, 16, dz, 64, F_, 32, 32, 64, 4 * nc, dw, 'conv', map4=_lib.W4_DOWN, phase_channels=nc, strides=(nc * 16, 16))
        assert maxerr(dw, w_ref.grad) < 5e-3


# ------------------------------------------------------------------------------------------------------------------ model level
DCGAN_SKIP = dict(nx=64, nc=3, nf=64, nhx=128, ny=20, nz=20, skipco=True, nt_inf=2, nh_inf=256, nlayers_inf=3, nh_res=512, nlayers_res=4,
                  archi='dcgan')
LOSS = dict(obs_scale=1.0, beta_y=1.0, beta_z=2.0, l2_res=1.0)


def _oracle_run(sd0, cfg, x, T, dt, seed, dev, loss_cfg, bf16=False):
    """fp32 (or torch-autocast bf16) oracle forward + backward on the GPU with the same random draws; returns (elbo terms, grads, out)."""
    from oracle import srvp_oracle as O
    sdo = {k: v.to(dev).clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in sd0.items()}
    torch.manual_seed(seed)
    rnd = O.draw_randoms(cfg, T, T, x.shape[1], training=True)
    rnd = {k: ([e.to(dev) for e in v] if isinstance(v, list) else v.to(dev)) for k, v in rnd.items()}
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    O.USE_ATEN_LSTM = False
    try:
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=bf16):
            o = O.forward(sdo, cfg, x.to(dev), T, dt, rnd, training=True)
        terms = O.elbo({k: (v.float() if torch.is_tensor(v) else v) for k, v in o.items()}, x.to(dev), loss_cfg)
        terms[0].backward()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
        O.USE_ATEN_LSTM = True
    return [float(t) for t in terms], {k: v.grad for k, v in sdo.items() if v.requires_grad}, o


@pytest.mark.parametrize('case', ['golden_noskip_nc1', 'golden_skip_nc3', 'skip_nc3'])
def test_dcgan_model_training_step(dev, case):
    """Full DCGAN64 training forward + backward: ELBO terms, outputs, running statistics and parameter gradients. The two golden cases
    are anchored on fixtures produced by the reference itself (oracle/make_golden.py), all three on the live oracle."""
    if case.startswith('golden'):
        g = load_golden('dcgan_nc1' if case == 'golden_noskip_nc1' else 'dcgan_skip_nc3')
        cfg, T, B, dt, loss_cfg, res_gain, seeds = g['cfg'], g['T'], g['B'], g['dt'], g['loss_cfg'], g['res_gain'], g['seeds']
    else:
        g = None
        cfg, T, B, dt, loss_cfg, res_gain, seeds = DCGAN_SKIP, 5, 6, 0.5, LOSS, 1.41, dict(model=3, input=31, fwd=17)
    m = build_model(cfg, res_gain, seeds['model'])
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(dev).train()
    x = make_input(T, B, cfg['nc'], seeds['input'])
    torch.manual_seed(seeds['fwd'])
    out = m(x.to(dev), T, dt=dt)
    loss, nll, kl_y, kl_z = model_loss(out, x.to(dev), loss_cfg)
    loss.backward()
    torch.cuda.synchronize()
    ref_terms, ref_grads, o = _oracle_run(sd0, cfg, x, T, dt, seeds['fwd'], dev, loss_cfg)
    if g is not None:   # the committed fixture (produced by the reference itself) is the first anchor
        t = g['train']
        assert float(loss) == pytest.approx(t['loss'], rel=1e-4)
        assert float(nll) == pytest.approx(t['nll'], rel=1e-4)
        assert float(kl_y) == pytest.approx(t['kl_y_0'], rel=1e-2)
        assert float(kl_z) == pytest.approx(t['kl_z'], rel=1e-2)
        sub = out[0][:, :, :, ::8, ::8].detach().cpu()
        assert float(((sub - t['x_sub']) ** 2).mean()) < 5e-5
        sd = m.state_dict()
        for k, v in t['running_after'].items():
            assert torch.allclose(sd[k].cpu(), v, rtol=5e-2, atol=6e-3), k
    assert float(loss) == pytest.approx(ref_terms[0], rel=1e-4)
    assert float(nll) == pytest.approx(ref_terms[1], rel=1e-4)
    assert float(kl_y) == pytest.approx(ref_terms[2], rel=1e-2)
    assert float(kl_z) == pytest.approx(ref_terms[3], rel=1e-2)
    assert float(((out[0] - o['x_']) ** 2).mean()) < 5e-5
    # gradients: judged against the deviation torch's own bf16 autocast of the same semantics shows (DESIGN.md "gradient parity")
    _, auto_grads, _ = _oracle_run(sd0, cfg, x, T, dt, seeds['fwd'], dev, loss_cfg, bf16=True)
    ours = {k: p.grad for k, p in m.named_parameters()}
    assert all(v is not None and torch.isfinite(v).all() for v in ours.values())
    e_ours = sorted(rel_l2(ours[k], ref_grads[k]) for k in ours)
    e_auto = sorted(rel_l2(auto_grads[k], ref_grads[k]) for k in ours)
    med = lambda v: v[len(v) // 2]
    assert med(e_ours) < 1.5 * med(e_auto) + 1e-3, (med(e_ours), med(e_auto))
    assert e_ours[-1] < 1.5 * e_auto[-1] + 1e-2, (e_ours[-1], e_auto[-1])
    last = 'decoder.conv.3.weight'
    assert rel_l2(ours[last], ref_grads[last]) < 2e-2


def test_dcgan_eval_rollout(dev):
    """Eval mode (running statistics, prior sampling beyond the conditioning frames) against the golden fixture, and the public
    encode / decode API with DCGAN skip shapes."""
    g = load_golden('dcgan_nc1')
    e = g['eval']
    m = build_model(g['cfg'], g['res_gain'], g['seeds']['model']).to(dev)
    x = make_input(g['T'], g['B'], g['cfg']['nc'], g['seeds']['input'])
    m.train()
    torch.manual_seed(g['seeds']['fwd'])
    with torch.no_grad():
        m(x.to(dev), g['T'], dt=g['dt'])
    m.eval()
    with torch.no_grad():
        torch.manual_seed(g['seeds']['fwd'])
        out = m(x[:e['nt_cond']].to(dev), e['nt_pred'], dt=g['dt'])
    assert out[0].shape[0] == e['nt_pred']
    for i, n in [(1, 'y'), (2, 'z'), (3, 'w'), (6, 'p_z_params')]:
        assert rel_l2(out[i], e[n]) < 1e-1, n
    sub = out[0][:, :, :, ::8, ::8].cpu()
    assert float(((sub - e['x_sub']) ** 2).mean()) < 1e-4
    ms = build_model(DCGAN_SKIP, 1.41, 3).to(dev).eval()
    xs = make_input(4, 3, 3, 5).to(dev)
    with torch.no_grad():
        hx, skips = ms.encode(xs)
        assert [tuple(s.shape[1:]) for s in skips] == [(512, 4, 4), (256, 8, 8), (128, 16, 16), (64, 32, 32)]
        o = ms(xs, 4, dt=1.0)
        x2 = ms.decode(o[3], o[1], skips)
    assert float(((x2 - o[0]) ** 2).mean()) < 1e-4
