"""GPU suite, part 1: every kernel family through the C ABI against torch on identical (bf16-rounded) operands.

Tolerances: outputs stored in bf16 carry a 2^-9 relative rounding (checked as <= 1.5e-2 of the tensor maximum); fp32 outputs
of bf16 x bf16 -> fp32 contractions are checked at 5e-3 of the tensor maximum; statistics / reductions at 1e-3 relative.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    torch.manual_seed(0)
    return 'cuda'


def bf(x):
    return x.to(torch.bfloat16).float()


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def ref_src(z_nhwc, scale, shift, lrelu, mode, frame_map):
    a = z_nhwc.float()
    if scale is not None:
        a = a * scale + shift
    if lrelu:
        a = F.leaky_relu(a, 0.2)
    a = bf(a).permute(0, 3, 1, 2)
    if mode == 1:
        a = F.max_pool2d(a, 2)
    elif mode == 2:
        a = F.interpolate(a, scale_factor=2, mode='nearest')
    if frame_map is not None:
        a = a[frame_map.long()]
    return a


def make_srcs(dev, frames, H, W, cins, modes, use_bn=True, fmap=False):
    from srvp_b200 import ops
    srcs, refs = [], []
    for cin, mode in zip(cins, modes):
        Hs, Ws = (H * 2, W * 2) if mode == 1 else (H // 2, W // 2) if mode == 2 else (H, W)
        fm, nf = None, frames
        if fmap and len(srcs) == 1:
            nf = max(1, frames // 2)
            fm = torch.randint(nf, (frames,), device=dev, dtype=torch.int32)
        z = torch.randn(nf, Hs, Ws, cin, device=dev).to(torch.bfloat16)
        sc = (torch.rand(cin, device=dev) + 0.5) if use_bn else None
        sh = (torch.randn(cin, device=dev) * 0.3) if use_bn else None
        srcs.append(ops.Src(z, cin, sc, sh, fm, 0, mode, use_bn))
        refs.append(ref_src(z, sc, sh, use_bn, mode, fm))
    return srcs, torch.cat(refs, 1)


CONV_CASES = [
    # name, frames, H, W, cins, cout, modes, use_bn, kind, fmap, sigmoid
    ('64->64@64', 3, 64, 64, [64], 64, [0], True, 'conv', False, False),
    ('64->128@32 pool', 5, 32, 32, [64], 128, [1], True, 'conv', False, False),
    ('128->256@16 pool', 7, 16, 16, [128], 256, [1], True, 'conv', False, False),
    ('512->512@8', 9, 8, 8, [512], 512, [0], True, 'conv', False, False),
    ('512+512->512@8 up+skip', 6, 8, 8, [512, 512], 512, [2, 0], True, 'conv', True, False),
    ('64+64->64@64 up+skip', 2, 64, 64, [64, 64], 64, [2, 0], True, 'conv', True, False),
    ('3(16)->64@64 thin', 3, 64, 64, [16], 64, [0], False, 'conv', False, False),
    ('64->3 convT sigmoid', 3, 64, 64, [64], 3, [0], True, 'convT', False, True),
    ('1 frame ragged', 1, 8, 8, [256], 256, [0], True, 'conv', False, False),
]


@pytest.mark.parametrize('case', CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv3x3_forward_stats_saved_input_and_wgrad(dev, case):
    """conv3x3 (fused loader + BN-stat epilogue + saved input) and wgrad3x3 against F.conv2d / autograd."""
    from srvp_b200 import ops
    name, frames, H, W, cins, cout, modes, use_bn, kind, fmap, sigmoid = case
    srcs, a = make_srcs(dev, frames, H, W, cins, modes, use_bn, fmap)
    cin_tot = sum(cins)
    cin_real = 3 if cin_tot == 16 else cin_tot
    if kind == 'conv':
        w = (torch.randn(cout, cin_real, 3, 3, device=dev) * 0.05).requires_grad_(True)
        ref = F.conv2d(a[:, :cin_real], bf(w.detach()), padding=1)
        refg = F.conv2d(a[:, :cin_real], w, padding=1)
    else:
        w = (torch.randn(cin_real, cout, 3, 3, device=dev) * 0.05).requires_grad_(True)
        ref = F.conv_transpose2d(a[:, :cin_real], bf(w.detach()), padding=1)
        refg = F.conv_transpose2d(a[:, :cin_real], w, padding=1)
    wp = ops.pack_conv3x3(w.detach(), kind)
    out, st, a_out = ops.conv3x3(srcs, wp, frames, H, W, cout, stats=not sigmoid, sigmoid_nchw=sigmoid, save_input=True)
    if sigmoid:
        got, ref = out, torch.sigmoid(ref)
    else:
        got = out.float().permute(0, 3, 1, 2)
    assert float((got - ref).abs().max() / ref.abs().max()) < 1.5e-2
    assert float((a_out.float() - a.permute(0, 2, 3, 1)).abs().max()) < 4e-2          # one bf16 ulp at |a| <= 8
    if st is not None:
        s = st.double().sum(0)
        o = out.double()
        s_ref = torch.stack([o.sum((0, 1, 2)), (o * o).sum((0, 1, 2))], 1)
        assert float(((s - s_ref).abs() / (s_ref.abs() + 1)).max()) < 1e-3
    # weight gradient from the saved input
    cpad = 16 if cout <= 16 else cout
    dz = torch.zeros(frames, H, W, cpad, device=dev, dtype=torch.bfloat16)
    dz[..., :cout] = (torch.randn(frames, H, W, cout, device=dev) * 0.1).to(torch.bfloat16)
    refg.backward(dz[..., :cout].float().permute(0, 3, 1, 2))
    dw = torch.zeros_like(w)
    ops.wgrad3x3(a_out, cin_tot, dz, cpad, frames, H, W, cout, cin_real, dw, kind)
    assert float((dw - w.grad).abs().max() / w.grad.abs().max()) < 5e-3


@pytest.mark.parametrize('kind,cin,cout', [('conv', 128, 64), ('conv', 64, 256), ('convT', 64, 3)])
def test_conv3x3_data_gradient(dev, kind, cin, cout):
    from srvp_b200 import ops
    Fr, H, W = 4, 16, 16
    a = torch.randn(Fr, cin, H, W, device=dev).to(torch.bfloat16).float().requires_grad_(True)
    if kind == 'conv':
        w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
        out = F.conv2d(a, bf(w), padding=1)
    else:
        w = torch.randn(cin, cout, 3, 3, device=dev) * 0.05
        out = F.conv_transpose2d(a, bf(w), padding=1)
    dz = (torch.randn_like(out) * 0.1).to(torch.bfloat16)
    out.backward(dz.float())
    cpad = ops.padded_k(cout)
    dzn = torch.zeros(Fr, H, W, cpad, device=dev, dtype=torch.bfloat16)
    dzn[..., :cout] = dz.permute(0, 2, 3, 1)
    da, _ = ops.conv3x3([ops.Src(dzn, cpad)], ops.pack_conv3x3(w, kind + '_dgrad'), Fr, H, W, cin)
    assert rel(da.float().permute(0, 3, 1, 2), a.grad) < 5e-3


class _BN:
    pass


@pytest.mark.parametrize('mode,with_skip,C,H', [(0, False, 64, 16), (1, False, 64, 16), (2, False, 64, 16), (0, True, 64, 16),
                                                (1, True, 128, 8), (0, False, 512, 8)])
def test_bn_lrelu_pool_backward(dev, mode, with_skip, C, H):
    """srvp_bn_bwd_{reduce,finalize,apply} against autograd of BatchNorm2d(train) -> LeakyReLU -> MaxPool2d / Upsample."""
    from srvp_b200 import ops
    Fr, W = 6, H
    z = torch.randn(Fr, H, W, C, device=dev).to(torch.bfloat16)
    gamma = (torch.rand(C, device=dev) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device=dev) * 0.2).requires_grad_(True)
    zf = z.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    a = F.leaky_relu(F.batch_norm(zf, None, None, gamma, beta, True, 0.0, 1e-5), 0.2)
    out = F.max_pool2d(a, 2) if mode == 1 else F.interpolate(a, scale_factor=2, mode='nearest') if mode == 2 else a
    da = (torch.randn_like(out) * 0.1).to(torch.bfloat16)
    bn = _BN()
    bn.weight, bn.bias = gamma.detach(), beta.detach()
    st = ops.BNState(C, dev)
    ops.bn_finalize(ops.channel_stats(z.view(-1, C)), float(Fr * H * W), bn, st, training_update=False)
    loss = (out * da.float()).sum()
    nt, B, skip, inv = 3, 2, None, None
    if with_skip:
        skip = (torch.randn(nt * B, H, W, 2 * C, device=dev) * 0.1).to(torch.bfloat16)
        inv = torch.full((Fr,), -1, dtype=torch.int32, device=dev)
        inv[1], inv[4] = 0, 1
        sk = skip[..., C:].float().view(nt, B, H, W, C).sum(0).permute(0, 3, 1, 2)
        loss = loss + (a[1] * sk[0]).sum() + (a[4] * sk[1]).sum()
    loss.backward()
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dz = ops.bn_bwd(z, st, gamma.detach(), dg, db, da.permute(0, 2, 3, 1).contiguous(), mode, Fr, H, W, C, skip=skip, skip_coff=C, nt=nt, B=B,
                    inv_map=inv)
    assert rel(dz.float(), zf.grad.permute(0, 2, 3, 1)) < 5e-3
    assert rel(dg, gamma.grad) < 1e-3 and rel(db, beta.grad) < 1e-3


@pytest.mark.parametrize('C,H,with_skip', [(64, 16, True), (512, 8, True), (256, 8, False), (128, 32, True)])
def test_bn_pooled_backward_lean_kernel(dev, C, H, with_skip):
    """bn_bwd_pool_kernel (max-pooled consumer, skip gradient already summed over time: nt = 1 as the per-video split of the decoder
    produces it) against autograd of BatchNorm2d(train) -> LeakyReLU -> MaxPool2d, plus the skip consumer on selected frames."""
    from srvp_b200 import ops
    Fr, W, B = 7, H, 3
    torch.manual_seed(C + H)
    z = torch.randn(Fr, H, W, C, device=dev).to(torch.bfloat16)
    gamma = (torch.rand(C, device=dev) + 0.5)
    gamma[::3] *= -1                                   # negative scales: arg-MIN routing
    gamma.requires_grad_(True)
    beta = (torch.randn(C, device=dev) * 0.2).requires_grad_(True)
    zf = z.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    a = F.leaky_relu(F.batch_norm(zf, None, None, gamma, beta, True, 0.0, 1e-5), 0.2)
    out = F.max_pool2d(a, 2)
    da = (torch.randn_like(out) * 0.1).to(torch.bfloat16)
    bn = _BN()
    bn.weight, bn.bias = gamma.detach(), beta.detach()
    st = ops.BNState(C, dev)
    ops.bn_finalize(ops.channel_stats(z.view(-1, C)), float(Fr * H * W), bn, st, training_update=False)
    loss = (out * da.float()).sum()
    skip, inv = None, None
    if with_skip:
        skip = (torch.randn(B, H, W, C + 64, device=dev) * 0.1).to(torch.bfloat16)
        inv = torch.full((Fr,), -1, dtype=torch.int32, device=dev)
        sel = [1, 4, 6]
        for v, f in enumerate(sel):
            inv[f] = v
            loss = loss + (a[f] * skip[v, ..., 64:].float().permute(2, 0, 1)).sum()
    loss.backward()
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dz = ops.bn_bwd(z, st, gamma.detach(), dg, db, da.permute(0, 2, 3, 1).contiguous(), 1, Fr, H, W, C, skip=skip, skip_coff=64, nt=1, B=B,
                    inv_map=inv)
    assert rel(dz.float(), zf.grad.permute(0, 2, 3, 1)) < 5e-3
    assert rel(dg, gamma.grad) < 1e-3 and rel(db, beta.grad) < 1e-3


def test_bn_finalize_statistics_and_running_update(dev):
    from srvp_b200 import ops
    C, rows = 96, 5000
    z = (torch.randn(rows, C, device=dev) * 2 + 0.5).to(torch.bfloat16)
    bn = torch.nn.BatchNorm2d(C).to(dev)
    ref = torch.nn.BatchNorm2d(C).to(dev)
    st = ops.BNState(C, dev)
    ops.bn_finalize(ops.channel_stats(z), float(rows), bn, st)
    ref.train()
    ref(z.float().t().reshape(1, C, rows, 1))
    assert torch.allclose(bn.running_mean, ref.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(bn.running_var, ref.running_var, rtol=1e-4, atol=1e-5)
    zf = z.float()
    assert torch.allclose(st.mean, zf.mean(0), rtol=1e-4, atol=1e-5)
    assert torch.allclose(st.invstd, 1 / torch.sqrt(zf.var(0, unbiased=False) + 1e-5), rtol=1e-4)


@pytest.mark.parametrize('M,N,K,ta,tb', [(200, 128, 8192, False, False), (37, 8192, 306, False, True), (306, 700, 450, True, True),
                                         (128, 8192, 2304, True, True), (1, 1, 1, False, False), (513, 129, 65, True, False)])
def test_gemm_strided_operands(dev, M, N, K, ta, tb):
    """srvp_gemm with K-major / MN-major fp32 and bf16 operands, ragged sizes, bias and activation epilogues."""
    from srvp_b200 import ops, _lib
    A = torch.randn(K, M, device=dev).t() if ta else torch.randn(M, K, device=dev)
    Bm = (torch.randn(K, N, device=dev).to(torch.bfloat16).t() if tb else torch.randn(N, K, device=dev).to(torch.bfloat16))
    bias = torch.randn(N, device=dev)
    ref = bf(A) @ Bm.float().t() + bias
    c = torch.empty(M, N, device=dev)
    ops.gemm(A, Bm, c, bias=bias)
    assert float((c - ref).abs().max() / ref.abs().max()) < 5e-3
    c2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(A, Bm, c2, bias=bias, act=_lib.ACT_RELU)
    assert float((c2.float() - F.relu(ref)).abs().max() / ref.abs().max()) < 1.5e-2
    c3 = torch.ones(M, N, device=dev)
    ops.gemm(A, Bm, c3, accumulate=True)              # may use split-K with red.add
    assert float((c3 - (ref - bias + 1)).abs().max() / ref.abs().max()) < 5e-3


def _mlp_q(mlp, x):
    rq = lambda t: t + (t.detach().to(torch.bfloat16).float() - t.detach())
    for i, lin in enumerate(mlp.linears()):
        if i > 0:
            x = F.relu(x)
        x = F.linear(rq(x), rq(lin.weight), lin.bias)
    return x


def _ref_loop(p_z, dyn, y0, z_post, eps, nt, os_, dt, n_post):
    y, ys, pzs, zs, ress = y0, [y0], [], [], []
    for s in range(os_ * (nt - 1)):
        fr = s // os_
        if s % os_ == 0:
            pp = _mlp_q(p_z, y)
            pzs.append(pp)
            if fr < n_post:
                z = z_post[fr]
            else:
                mu, rho = pp.chunk(2, -1)
                z = mu + (F.softplus(rho) + 1e-8) * eps[fr]
            zs.append(z)
        r = dt * _mlp_q(dyn, torch.cat([y, zs[-1]], 1))
        y = y + r
        ys.append(y)
        ress.append(r)
    return torch.stack(ys), torch.stack(pzs), torch.stack(zs), torch.stack(ress)


@pytest.mark.parametrize('B,ny,nz,nh,nl,nt,os_,n_post', [(16, 50, 50, 512, 4, 4, 2, 3), (37, 20, 20, 512, 4, 6, 1, 5), (24, 50, 50, 512, 4, 10, 2, 4),
                                                        (8, 50, 50, 256, 3, 5, 2, 4)])
def test_latent_loop_forward(dev, B, ny, nz, nh, nl, nt, os_, n_post):
    """The persistent Euler-loop kernel (posterior and prior-sampled frames) against the step-by-step torch loop."""
    from srvp_b200 import latent
    from srvp_b200.module.mlp import MLP
    p_z, dyn = MLP(ny, nh, 2 * nz, nl).to(dev), MLP(ny + nz, nh, ny, nl).to(dev)
    y0, z_post, eps = torch.randn(B, ny, device=dev), torch.randn(max(n_post, 1), B, nz, device=dev), torch.randn(nt - 1, B, nz, device=dev)
    dt = 1.0 / os_
    with torch.no_grad():
        ry, rp, rz, rr = _ref_loop(p_z, dyn, y0, z_post, eps, nt, os_, dt, n_post)
        out = latent.latent_fwd(p_z.linears(), dyn.linears(), y0, z_post, eps, nt, os_, dt, n_post, nh)
    assert rel(out['y_all'], ry) < 1e-2 and rel(out['pz'], rp) < 1e-2 and rel(out['z'], rz) < 1e-2 and rel(out['res'], rr) < 1.5e-2


@pytest.mark.parametrize('B,ny,nz,nh,nl,nt,os_', [(16, 50, 50, 512, 4, 4, 2), (37, 20, 20, 512, 4, 6, 1), (8, 50, 50, 256, 3, 5, 2)])
def test_latent_loop_backward(dev, B, ny, nz, nh, nl, nt, os_):
    """Reverse-time kernel + GEMM weight gradients against autograd of the torch loop with the same bf16 operand rounding.
    (ReLU masks of nearly-zero units may differ between the two summation orders: 3e-2 relative L2 per tensor.)"""
    from srvp_b200 import latent
    from srvp_b200.module.mlp import MLP
    p_z, dyn = MLP(ny, nh, 2 * nz, nl).to(dev), MLP(ny + nz, nh, ny, nl).to(dev)
    n_post, dt, S = nt - 1, 1.0 / os_, os_ * (nt - 1)
    y0 = torch.randn(B, ny, device=dev, requires_grad=True)
    z_post = torch.randn(n_post, B, nz, device=dev, requires_grad=True)
    ry, rp, rz, rr = _ref_loop(p_z, dyn, y0, z_post, None, nt, os_, dt, n_post)
    g_y = torch.zeros(S + 1, B, ny, device=dev)
    g_y[::os_] = torch.randn(nt, B, ny, device=dev)
    g_res, g_pz = torch.randn(S, B, ny, device=dev) * 0.1, torch.randn(nt - 1, B, 2 * nz, device=dev) * 0.1
    ((ry * g_y).sum() + (rr * g_res).sum() + (rp * g_pz).sum()).backward()
    with torch.no_grad():
        fwd = latent.latent_fwd(p_z.linears(), dyn.linears(), y0.detach(), z_post.detach(), None, nt, os_, dt, n_post, nh)
        d_y0, d_z, gp, gd = latent.latent_bwd(p_z.linears(), dyn.linears(), fwd, g_y, g_res, g_pz, nt, os_, dt, nh)
    assert rel(d_y0, y0.grad) < 3e-2 and rel(d_z, z_post.grad) < 3e-2
    for lin, (dW, db) in list(zip(p_z.linears(), gp)) + list(zip(dyn.linears(), gd)):
        assert rel(dW, lin.weight.grad) < 3e-2 and rel(db, lin.bias.grad) < 3e-2


def test_layout_roundtrip_and_sigmoid_backward(dev):
    from srvp_b200 import ops
    x = torch.rand(5, 3, 64, 64, device=dev)
    nh = ops.nchw_to_nhwc_bf16(x, 16)
    assert torch.equal(nh[..., 3:], torch.zeros_like(nh[..., 3:]))
    assert torch.equal(ops.nhwc_to_nchw_f32(nh, 3), bf(x))
    xh = torch.sigmoid(torch.randn(5, 3, 64, 64, device=dev))
    dx = torch.randn_like(xh)
    dz = ops.sigmoid_bwd(dx, xh)
    assert rel(dz[..., :3].float().permute(0, 3, 1, 2), dx * xh * (1 - xh)) < 5e-3
    t = torch.randn(7, 33, 65, device=dev)
    assert torch.equal(ops.transpose_last2(t), t.transpose(1, 2).contiguous())


# ------------------------------------------------------------------------------------------------ fp32 inference networks
@pytest.mark.parametrize('act', [None, 'relu', 'tanh'])
def test_linear_f32_forward_backward(dev, act):
    """srvp_linear_f32 + srvp_act_bwd_f32 + srvp_colsum against nn.Linear autograd (fp32, matmul TF32 off): 1e-5 relative."""
    from srvp_b200 import infer
    lin = torch.nn.Linear(130, 77).to(dev)
    x = torch.randn(3, 67, 130, device=dev, requires_grad=True)
    y = infer.linear(x, lin, act)
    x2 = x.detach().clone().requires_grad_(True)
    lin2 = torch.nn.Linear(130, 77).to(dev)
    lin2.load_state_dict(lin.state_dict())
    y2 = lin2(x2)
    y2 = torch.relu(y2) if act == 'relu' else torch.tanh(y2) if act == 'tanh' else y2
    assert rel(y, y2) < 1e-5
    g = torch.randn_like(y)
    y.backward(g)
    y2.backward(g)
    assert rel(x.grad, x2.grad) < 1e-5 and rel(lin.weight.grad, lin2.weight.grad) < 1e-5 and rel(lin.bias.grad, lin2.bias.grad) < 1e-5


@pytest.mark.parametrize('T,B', [(12, 192), (5, 3), (1, 7)])
def test_lstm_forward_backward(dev, T, B):
    """One-launch LSTM recurrence and its reverse-time pass against nn.LSTM (cuDNN, TF32 disabled): 2e-5 relative."""
    from srvp_b200 import infer
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = torch.nn.LSTM(128, 256, 1).to(dev)
        ours = torch.nn.LSTM(128, 256, 1).to(dev)
        ours.load_state_dict(ref.state_dict())
        x = torch.randn(T, B, 128, device=dev, requires_grad=True)
        x2 = x.detach().clone().requires_grad_(True)
        h = infer.lstm(x, ours)
        h2 = ref(x2)[0]
        assert rel(h, h2) < 2e-5
        g = torch.randn_like(h)
        h.backward(g)
        h2.backward(g)
        assert rel(x.grad, x2.grad) < 2e-5
        for n in ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0']:
            if T == 1 and n == 'weight_hh_l0':
                assert float(getattr(ours, n).grad.abs().max()) == 0.0
                continue
            assert rel(getattr(ours, n).grad, getattr(ref, n).grad) < 2e-5, n
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


# ------------------------------------------------------------------------------------------------ fused ELBO reductions
def test_fused_elbo_matches_torch_distributions(dev):
    """srvp_b200.elbo.elbo against the reference's loss assembly (train.py:90-106 via module/utils.py): values 1e-6, gradients 1e-5."""
    from common import model_loss
    from srvp_b200 import elbo
    T, B, ny, nz, S = 5, 7, 20, 30, 8
    g = torch.Generator(device='cpu').manual_seed(4)
    mk = lambda *s: torch.randn(*s, generator=g).to(dev).requires_grad_(True)
    x = torch.rand(T, B, 3, 64, 64, generator=g).to(dev)
    x_ = torch.rand(T, B, 3, 64, 64, generator=g).to(dev).requires_grad_(True)
    qy, qz, pz, res = mk(B, 2 * ny), mk(T - 1, B, 2 * nz), mk(T - 1, B, 2 * nz), mk(S, B, ny)
    with torch.no_grad():
        qz[0, 0, nz:] = 25.0      # softplus threshold branch
        res[0, 0] = 0.0           # zero residual row: norm gradient 0
    loss_cfg = dict(obs_scale=0.71, beta_y=1.5, beta_z=2.0, l2_res=0.7)
    out = (x_, None, None, None, qy, qz, pz, res)
    l1, n1, ky1, kz1 = elbo.elbo(out, x, **loss_cfg)
    l1.backward()
    g1 = [t.grad.clone() for t in (x_, qy, qz, pz, res)]
    for t in (x_, qy, qz, pz, res):
        t.grad = None
    l2, n2, ky2, kz2 = model_loss(out, x, loss_cfg)
    l2.backward()
    for a, b in [(l1, l2), (n1, n2), (ky1, ky2), (kz1, kz2)]:
        assert float(a) == pytest.approx(float(b), rel=2e-6)
    for a, t in zip(g1, (x_, qy, qz, pz, res)):
        assert rel(a, t.grad) < 1e-5


# ------------------------------------------------------------------------------------------------ training-step edges
def test_adam_multi_matches_torch_adam(dev):
    """srvp_b200.optim.Adam (one launch for all tensors) against torch.optim.Adam over 5 steps, with an LR scheduler in the loop."""
    from srvp_b200.optim import Adam
    shapes = [(64, 3, 3, 3), (64,), (512, 1024, 3, 3), (1, ), (100, 512), (33333,)]
    g = torch.Generator().manual_seed(2)
    p1 = [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in shapes]
    p2 = [torch.nn.Parameter(p.detach().clone()) for p in p1]
    o1, o2 = Adam(p1, lr=3e-4), torch.optim.Adam(p2, lr=3e-4)
    s1 = torch.optim.lr_scheduler.LambdaLR(o1, lr_lambda=lambda i: max(0, (10 - i) / 10))
    s2 = torch.optim.lr_scheduler.LambdaLR(o2, lr_lambda=lambda i: max(0, (10 - i) / 10))
    for step in range(5):
        for a, b in zip(p1, p2):
            gr = torch.randn(a.shape, generator=g).to(dev) * (10.0 ** (step - 2))
            a.grad, b.grad = gr.clone(), gr.clone()
        o1.step(); o2.step(); s1.step(); s2.step()
    for a, b in zip(p1, p2):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
    sd1, sd2 = o1.state_dict(), o2.state_dict()
    assert set(sd1['state'][0].keys()) == set(sd2['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}
    assert torch.allclose(sd1['state'][2]['exp_avg_sq'], sd2['state'][2]['exp_avg_sq'], rtol=1e-5, atol=1e-12)


def test_rsample_and_u8_input(dev):
    from srvp_b200 import infer, ops
    params = torch.randn(7, 5, 40, device=dev, requires_grad=True)
    with torch.no_grad():
        params[0, 0, 20:] = 30.0
    eps = torch.randn(7, 5, 20, device=dev)
    z = infer.rsample(params, eps)
    p2 = params.detach().clone().requires_grad_(True)
    loc, rho = p2.chunk(2, -1)
    z2 = loc + eps * (F.softplus(rho) + 1e-8)
    assert rel(z, z2) < 1e-6
    gz = torch.randn_like(z)
    z.backward(gz); z2.backward(gz)
    assert rel(params.grad, p2.grad) < 1e-6
    # uint8 (B, T, H, W, C) -> (T*B, H, W, 16) bf16 in [0, 1]
    v = torch.randint(0, 256, (3, 4, 64, 64, 3), dtype=torch.uint8, device=dev)
    out = ops.u8_to_nhwc_bf16(v, 16)
    ref = torch.zeros(4 * 3, 64, 64, 16, device=dev)
    ref[..., :3] = (v.float() / 255).permute(1, 0, 2, 3, 4).reshape(12, 64, 64, 3)
    assert torch.equal(out.float(), bf(ref))
    x = ops.u8_to_tbchw_f32(v)
    assert torch.allclose(x, (v.float() / 255).permute(1, 0, 4, 2, 3), rtol=1e-6, atol=0)


# ------------------------------------------------------------------------------------------------ thin operands at 64x64 (csrc/thin.cu)
@pytest.mark.parametrize('nc,frames', [(3, 5), (1, 2), (3, 40)])
def test_thin_conv_first_encoder_block(dev, nc, frames):
    """Conv2d(nc, 64, 3, 1, 1) on the padded 16-channel image tensor through the im2col kernel (the shape engine._encoder_fwd launches: no
    fused transform, no saved input): output, BN statistics rows (one per CTA, srvp_conv3x3_num_mtiles) and run-to-run determinism."""
    from srvp_b200 import ops
    torch.manual_seed(nc * 100 + frames)
    x = torch.zeros(frames, 64, 64, 16, device=dev, dtype=torch.bfloat16)
    x[..., :nc] = torch.rand(frames, 64, 64, nc, device=dev).to(torch.bfloat16)
    w = torch.randn(64, nc, 3, 3, device=dev) * 0.2
    ref = F.conv2d(x[..., :nc].float().permute(0, 3, 1, 2), bf(w), padding=1).permute(0, 2, 3, 1)
    wp = ops.pack_conv3x3(w, 'conv')
    out, st = ops.conv3x3([ops.Src(x, 16)], wp, frames, 64, 64, 64, stats=True, cin_real=nc)
    assert st.shape[0] == ops.conv3x3_stats_rows(frames, 64, 64, 64, 16)
    assert float((out.float() - ref).abs().max() / ref.abs().max()) < 1.5e-2
    s = st.double().sum(0)
    o = out.double()
    s_ref = torch.stack([o.sum((0, 1, 2)), (o * o).sum((0, 1, 2))], 1)
    assert float(((s - s_ref).abs() / (s_ref.abs() + 1)).max()) < 1e-3
    out2, st2 = ops.conv3x3([ops.Src(x, 16)], wp, frames, 64, 64, 64, stats=True, cin_real=nc)
    assert torch.equal(out, out2) and torch.equal(st, st2)


def test_thin_conv_head_data_gradient(dev):
    """Data gradient of ConvTranspose2d(64, nc, 3, 1, 1): the thin dz (16 padded channels) convolved back to 64 channels."""
    from srvp_b200 import ops
    frames, nc = 3, 3
    dz = torch.zeros(frames, 64, 64, 16, device=dev, dtype=torch.bfloat16)
    dz[..., :nc] = (torch.randn(frames, 64, 64, nc, device=dev) * 0.1).to(torch.bfloat16)
    wt = torch.randn(64, nc, 3, 3, device=dev) * 0.2            # ConvTranspose2d weight (cin, cout, 3, 3)
    a = torch.randn(frames, 64, 64, 64, device=dev, requires_grad=True)
    y = F.conv_transpose2d(a.permute(0, 3, 1, 2), bf(wt), padding=1)
    y.backward(dz[..., :nc].float().permute(0, 3, 1, 2))
    da, _ = ops.conv3x3([ops.Src(dz, 16)], ops.pack_conv3x3(wt, 'convT_dgrad'), frames, 64, 64, 64, cin_real=nc)
    assert float((da.float() - a.grad).abs().max() / a.grad.abs().max()) < 1.5e-2


@pytest.mark.parametrize('nc,frames', [(3, 4), (1, 3), (3, 37)])
def test_thin_wgrad_both_ends(dev, nc, frames):
    """Weight gradients with a thin operand: the first encoder block (images thin, dz wide) and the decoder head (dz thin, activations
    wide and computed on the way in from the raw z with its batch-norm affine + LeakyReLU) against torch autograd."""
    from srvp_b200 import ops
    torch.manual_seed(nc + frames)
    # first encoder block: Conv2d(nc, 64)
    x = torch.zeros(frames, 64, 64, 16, device=dev, dtype=torch.bfloat16)
    x[..., :nc] = torch.rand(frames, 64, 64, nc, device=dev).to(torch.bfloat16)
    dz = (torch.randn(frames, 64, 64, 64, device=dev) * 0.1).to(torch.bfloat16)
    w = torch.zeros(64, nc, 3, 3, device=dev, requires_grad=True)
    F.conv2d(x[..., :nc].float().permute(0, 3, 1, 2), w, padding=1).backward(dz.float().permute(0, 3, 1, 2))
    dw = torch.zeros(64, nc, 3, 3, device=dev)
    ops.wgrad3x3(x, 16, dz, 64, frames, 64, 64, 64, nc, dw, 'conv')
    assert rel(dw, w.grad) < 2e-3
    # decoder head: ConvTranspose2d(64, nc) over a = lrelu(z * scale + shift)
    z = torch.randn(frames, 64, 64, 64, device=dev).to(torch.bfloat16)
    sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.3
    a = bf(F.leaky_relu(z.float() * sc + sh, 0.2))
    dzt = torch.zeros(frames, 64, 64, 16, device=dev, dtype=torch.bfloat16)
    dzt[..., :nc] = (torch.randn(frames, 64, 64, nc, device=dev) * 0.1).to(torch.bfloat16)
    wt = torch.zeros(64, nc, 3, 3, device=dev, requires_grad=True)
    F.conv_transpose2d(a.permute(0, 3, 1, 2), wt, padding=1).backward(dzt[..., :nc].float().permute(0, 3, 1, 2))
    dwt = torch.zeros(64, nc, 3, 3, device=dev)
    ops.wgrad3x3(z, 64, dzt, 16, frames, 64, 64, nc, 64, dwt, 'convT', act_affine=(sc, sh, True))
    assert rel(dwt, wt.grad) < 2e-3
    # the same head gradient from a materialised activation (no transform) must agree as well
    dwt2 = torch.zeros(64, nc, 3, 3, device=dev)
    ops.wgrad3x3(a.to(torch.bfloat16), 64, dzt, 16, frames, 64, 64, nc, 64, dwt2, 'convT')
    assert rel(dwt2, wt.grad) < 2e-3


def test_packed_weight_cache_follows_the_weights(dev):
    """ops.pack_conv3x3 returns a cached operand; it must follow in-place updates of the weight (version counter), updates through raw
    pointers by srvp_b200.optim.Adam (PACK_EPOCH) and training-mode forwards, and re-pack every registered operand in one launch."""
    from srvp_b200 import ops, _lib
    from srvp_b200.optim import Adam
    w1 = torch.nn.Parameter(torch.randn(64, 64, 3, 3, device=dev) * 0.1)
    w2 = torch.nn.Parameter(torch.randn(128, 64, 3, 3, device=dev) * 0.1)
    fresh = lambda w, kind: ops.pack_conv3x3(w.detach().clone(), kind, out=torch.empty(ops.padded_n(w.shape[0] if kind == 'conv' else w.shape[1]) *
                                                                                   ops.padded_k(w.shape[1] if kind == 'conv' else w.shape[0]) * 9,
                                                                                   dtype=torch.bfloat16, device=dev))
    a1, a2, a3 = ops.pack_conv3x3(w1, 'conv'), ops.pack_conv3x3(w2, 'conv'), ops.pack_conv3x3(w2, 'conv_dgrad')
    assert ops.pack_conv3x3(w1, 'conv') is a1                      # cache hit: same buffer, no launch
    n0 = _lib.lib().srvp_launch_count()
    ops.pack_conv3x3(w1, 'conv'); ops.pack_conv3x3(w2, 'conv_dgrad')
    assert _lib.lib().srvp_launch_count() == n0
    with torch.no_grad():
        w1.mul_(2.0)                                               # in-place update: version counter
        w2.add_(0.5)
    n0 = _lib.lib().srvp_launch_count()
    b1 = ops.pack_conv3x3(w1, 'conv')
    assert _lib.lib().srvp_launch_count() == n0 + 1                # ONE launch refreshed all three operands
    assert b1 is a1 and torch.equal(b1, fresh(w1, 'conv'))
    assert torch.equal(ops.pack_conv3x3(w2, 'conv'), fresh(w2, 'conv')) and torch.equal(ops.pack_conv3x3(w2, 'conv_dgrad'), fresh(w2, 'conv_dgrad'))
    assert _lib.lib().srvp_launch_count() == n0 + 1 + 3            # (the three `fresh` packs themselves)
    opt = Adam([w1, w2], lr=1e-2)
    w1.grad, w2.grad = torch.ones_like(w1), torch.ones_like(w2)
    opt.step()                                                     # raw-pointer update: PACK_EPOCH
    assert torch.equal(ops.pack_conv3x3(w1, 'conv'), fresh(w1, 'conv')) and torch.equal(ops.pack_conv3x3(w2, 'conv_dgrad'), fresh(w2, 'conv_dgrad'))
    del w2                                                         # a dead weight drops out of the job table at the next refresh
    with torch.no_grad():
        w1.mul_(0.5)
    assert torch.equal(ops.pack_conv3x3(w1, 'conv'), fresh(w1, 'conv'))
