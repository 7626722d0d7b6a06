"""GPU suite, part 2: the drop-in model (forward, ELBO, backward, eval rollout) against the golden fixtures produced by the
reference and against the live CPU oracle, plus size-independent properties at the full BASELINE size.

Stated tolerances (bf16 operands, fp32 accumulation / statistics / latents; north star: ELBO within 1e-4 relative):
  total ELBO, NLL            rel <= 1e-4          per-pixel MSE of x_hat   <= 5e-5
  KL(y_0), KL(z) terms       rel <= 1e-2          latent tensors (y, z, ...) rel. L2 <= 8e-2
  parameter gradients        no worse than 1.5x what torch's own bf16 autocast of the reference semantics produces on the same
                             GPU (median over tensors of the relative L2 deviation from fp32), see DESIGN.md "gradient parity".
"""
import pytest
import torch

from common import build_model, load_golden, make_input, model_loss, rel_l2

pytestmark = pytest.mark.gpu
CASES = ['vgg_skip_nc3', 'vgg_skip_nc1']


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    return 'cuda'


@pytest.fixture(scope='module', params=CASES)
def run(request, dev):
    g = load_golden(request.param)
    m = build_model(g['cfg'], g['res_gain'], g['seeds']['model'])
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(dev).train()
    x = make_input(g['T'], g['B'], g['cfg']['nc'], g['seeds']['input'])
    torch.manual_seed(g['seeds']['fwd'])
    out = m(x.to(dev), g['T'], dt=g['dt'])
    loss, nll, kl_y, kl_z = model_loss(out, x.to(dev), g['loss_cfg'])
    loss.backward()
    torch.cuda.synchronize()
    return dict(g=g, m=m, sd0=sd0, x=x, out=out, loss=float(loss), nll=float(nll), kl_y=float(kl_y), kl_z=float(kl_z))


def test_elbo_matches_reference(run):
    """Tiny fixtures: the NLL deviation is a sum of N independent bf16 rounding effects, relative size ~ 1/sqrt(N) * 1/obs_scale^2.
    At obs_scale 0.71 (BAIR) the 1e-4 target holds even at T*B = 12 frames; the obs_scale 0.2 (KTH) fixture has only 41k pixels
    and is allowed 5e-4 here; test_elbo_at_moderate_size_kth checks the 1e-4 target on a less tiny batch."""
    t = run['g']['train']
    tol = 1e-4 if run['g']['loss_cfg']['obs_scale'] > 0.5 else 5e-4
    assert run['loss'] == pytest.approx(t['loss'], rel=tol)
    assert run['nll'] == pytest.approx(t['nll'], rel=tol)
    assert run['kl_y'] == pytest.approx(t['kl_y_0'], rel=1e-2)
    assert run['kl_z'] == pytest.approx(t['kl_z'], rel=1e-2)


def test_outputs_match_reference(run):
    t, out = run['g']['train'], run['out']
    names = ['x_', 'y', 'z', 'w', 'q_y_0_params', 'q_z_params', 'p_z_params', 'res']
    for i, n in enumerate(names):
        if n in t:
            assert rel_l2(out[i], t[n]) < 8e-2, n
    sub = out[0][:, :, :, ::8, ::8].detach().cpu()
    assert float(((sub - t['x_sub']) ** 2).mean()) < 5e-5
    assert float(out[0].mean()) == pytest.approx(t['x_mean'], abs=1e-3)


def test_running_statistics_updated_like_reference(run):
    sd = run['m'].state_dict()
    for k, v in run['g']['train']['running_after'].items():
        # running = 0.9*init + 0.1*batch statistic: bf16 activations move the deep layers' batch means by a few 1e-2
        assert torch.allclose(sd[k].cpu(), v, rtol=5e-2, atol=6e-3), k
    assert int(sd['encoder.conv.0.0.1.num_batches_tracked']) == 1


def test_gradients_vs_oracle_and_autocast_noise(run, dev):
    """fp32 gradients of the oracle vs ours, judged against the deviation torch's bf16 autocast of the same semantics shows."""
    from oracle import srvp_oracle as O
    g, m = run['g'], run['m']
    cfg = g['cfg']
    res = {}
    for mode in ('fp32', 'bf16'):
        sdo = {k: v.to(dev).clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in run['sd0'].items()}
        torch.manual_seed(g['seeds']['fwd'])
        rnd = O.draw_randoms(cfg, g['T'], g['T'], g['B'], training=True)
        rnd = {k: ([e.to(dev) for e in v] if isinstance(v, list) else v.to(dev)) for k, v in rnd.items()}
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        O.USE_ATEN_LSTM = False
        try:
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(mode == 'bf16')):
                o = O.forward(sdo, cfg, run['x'].to(dev), g['T'], g['dt'], rnd, training=True)
            l = O.elbo({k: (v.float() if torch.is_tensor(v) else v) for k, v in o.items()}, run['x'].to(dev), g['loss_cfg'])[0]
            l.backward()
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
            O.USE_ATEN_LSTM = True
        res[mode] = {k: v.grad for k, v in sdo.items() if v.requires_grad}
    ours = {k: p.grad for k, p in m.named_parameters()}
    e_ours = sorted(rel_l2(ours[k], res['fp32'][k]) for k in ours)
    e_auto = sorted(rel_l2(res['bf16'][k], res['fp32'][k]) for k in ours)
    med = lambda v: v[len(v) // 2]
    assert med(e_ours) < 1.5 * med(e_auto) + 1e-3, (med(e_ours), med(e_auto))
    assert e_ours[-1] < 1.5 * e_auto[-1] + 1e-2
    # the layers next to the loss are not yet affected by amplified activation noise: tight check there
    for k in ['decoder.conv.3.1.weight', 'decoder.conv.3.0.1.weight', 'decoder.conv.3.0.1.bias']:
        assert rel_l2(ours[k], res['fp32'][k]) < 2e-2, k


def test_eval_rollout_matches_reference(run, dev):
    """Eval mode: running statistics, last-frame skips, prior sampling beyond the conditioning frames (test.py:235-246 pattern)."""
    g = run['g']
    e = g['eval']
    m = build_model(g['cfg'], g['res_gain'], g['seeds']['model']).to(dev)
    # the fixture's eval pass ran after ONE training forward (running statistics updated once): replay that
    m.train()
    torch.manual_seed(g['seeds']['fwd'])
    with torch.no_grad():
        m(run['x'].to(dev), g['T'], dt=g['dt'])
    m.eval()
    with torch.no_grad():
        torch.manual_seed(g['seeds']['fwd'])
        out = m(run['x'][:e['nt_cond']].to(dev), e['nt_pred'], dt=g['dt'])
    assert out[0].shape[0] == e['nt_pred'] and out[5] is not None and out[5].shape[0] == e['nt_cond'] - 1
    for i, n in [(1, 'y'), (2, 'z'), (3, 'w'), (6, 'p_z_params')]:
        assert rel_l2(out[i], e[n]) < 1e-1, n
    sub = out[0][:, :, :, ::8, ::8].cpu()
    assert float(((sub - e['x_sub']) ** 2).mean()) < 1e-4
    # public encode/decode API used by test.py: skips are (B, C, H, W) tensors, deepest first
    with torch.no_grad():
        hx, skips = m.encode(run['x'][:e['nt_cond']].to(dev))
        assert [tuple(s.shape[1:]) for s in skips] == [(512, 8, 8), (256, 16, 16), (128, 32, 32), (64, 64, 64)]
        x2 = m.decode(out[3], out[1][1:], skips)
    assert float(((x2 - out[0][1:]) ** 2).mean()) < 1e-4


def test_elbo_at_moderate_size_kth(dev):
    """KTH hyper-parameters (nc=1, obs_scale 0.2, nt_inf 3) at T=6, B=24 against the live CPU oracle: total ELBO within 1e-4."""
    from oracle import srvp_oracle as O
    g = load_golden('vgg_skip_nc1')
    cfg, T, B = g['cfg'], 6, 24
    m = build_model(cfg, g['res_gain'], 1)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(dev).train()
    x = make_input(T, B, cfg['nc'], 321)
    torch.manual_seed(9)
    with torch.no_grad():
        out = m(x.to(dev), T, dt=g['dt'])
        loss = float(model_loss(out, x.to(dev), g['loss_cfg'])[0])
        torch.manual_seed(9)
        rnd = O.draw_randoms(cfg, T, T, B, training=True)
        ref = float(O.elbo(O.forward(sd, cfg, x, T, g['dt'], rnd, training=True), x, g['loss_cfg'])[0])
    assert loss == pytest.approx(ref, rel=1e-4)


def test_full_size_properties(dev):
    """BASELINE size (BAIR: T=12, B=192): determinism, finite ELBO, video-permutation equivariance of the training forward."""
    import bench
    from srvp_b200.module.srvp import StochasticLatentResidualVideoPredictor
    torch.manual_seed(1)
    m = StochasticLatentResidualVideoPredictor(*[bench.CFG[k] for k in bench.ARG_ORDER])
    m.init()
    m = m.to(dev).train()
    x = torch.rand(bench.SEQ_LEN, bench.BATCH, 3, 64, 64, generator=torch.Generator().manual_seed(5)).to(dev)

    def fwd(xx, perm=None):
        torch.manual_seed(11)
        # identical random draws for both calls; with a permutation the per-video draws are permuted alike
        import srvp_b200.module.srvp as S
        out = m(xx, bench.SEQ_LEN, dt=bench.DT)
        return out, float(bench.elbo_loss(out, xx))

    with torch.no_grad():
        o1, l1 = fwd(x)
        o2, l2 = fwd(x)
    assert l1 == l2 and torch.equal(o1[0], o2[0]), 'the training forward is not deterministic'
    assert torch.isfinite(o1[0]).all() and 1e4 < l1 < 1e6
    assert float(o1[0].min()) >= 0.0 and float(o1[0].max()) <= 1.0
    # eval mode has no per-video random frame choice: permuting the videos permutes the outputs (batch-norm uses running stats)
    m.eval()
    perm = torch.randperm(bench.BATCH, generator=torch.Generator().manual_seed(3)).to(dev)
    with torch.no_grad():
        m.noise_device = 'cuda'
        torch.manual_seed(11); torch.cuda.manual_seed(11)
        hx1, _ = m.encode(x[:4])
        hx2, _ = m.encode(x[:4, perm])
    assert torch.allclose(hx1[:, perm], hx2, rtol=0, atol=2e-2)


def test_batched_sample_rollout_decode_matches_per_sample(dev):
    """test.py's batched sampling (samples folded into the batch, skips shared through the frame map) decodes exactly what the
    reference-style per-sample loop decodes from the same latents."""
    g = load_golden('vgg_skip_nc3')
    m = build_model(g['cfg'], g['res_gain'], 2).to(dev).eval()
    B, S, nt = 2, 3, 4
    x = make_input(3, B, g['cfg']['nc'], 11).to(dev)
    with torch.no_grad():
        hx, handle = m._encode_fused(x)
        w = m.infer_w(hx)
        y = torch.randn(nt, S * B, g['cfg']['ny'], generator=torch.Generator().manual_seed(3)).to(dev)
        xb = m._decode_fused(w.repeat(S, 1), y, handle.levels, handle.frame_map.repeat(S), None)
        skips = m.encode(x)[1]
        for s in range(S):
            xs = m.decode(w, y[:, s * B:(s + 1) * B].contiguous(), skips)
            assert float(((xs - xb[:, s * B:(s + 1) * B]) ** 2).mean()) < 1e-5


def test_weight_gradient_stream_matches_inline(dev):
    """The step bench.py / train.py run (GradBucket: gradients accumulated in place, weight gradients recorded and issued on their own
    stream, the decoder's low-resolution ones held back, joined by allreduce_mean() / Adam.step()) must produce the gradients of the
    plain autograd path with everything on one stream. Differences: summation order of the split-K atomics only."""
    from srvp_b200 import ops, parallel
    g = load_golden('vgg_skip_nc3')
    x = make_input(6, 8, g['cfg']['nc'], 5).to(dev)

    def grads(bucketed, stream_on):
        m = build_model(g['cfg'], g['res_gain'], g['seeds']['model']).to(dev).train()
        old = (ops.WGRAD_STREAM, ops.DEFER_JOIN, parallel.ACTIVE_BUCKET)
        ops.WGRAD_STREAM, ops.DEFER_JOIN = int(stream_on), False
        try:
            bucket = None
            if bucketed:
                bucket = parallel.GradBucket(list(m.parameters()), early=list(m.decoder.parameters()))
                parallel.ACTIVE_BUCKET = bucket
                bucket.zero()
            torch.manual_seed(3)
            loss = model_loss(m(x, 6, dt=0.5), x, g['loss_cfg'])[0]
            n0 = ops.SIDE_LAUNCHES[0]
            loss.backward()
            assert (ops.SIDE_LAUNCHES[0] > n0) == bool(stream_on)
            if bucketed:
                bucket.allreduce_mean()        # joins the weight-gradient stream
            torch.cuda.synchronize()
            return {n: p.grad.detach().clone() for n, p in m.named_parameters()}
        finally:
            ops.WGRAD_STREAM, ops.DEFER_JOIN, parallel.ACTIVE_BUCKET = old

    ref = grads(False, False)
    for bucketed, stream_on in ((True, True), (False, True), (True, False)):
        got = grads(bucketed, stream_on)
        for n, r in ref.items():
            assert rel_l2(got[n], r) < 2e-3, (bucketed, stream_on, n, rel_l2(got[n], r))
