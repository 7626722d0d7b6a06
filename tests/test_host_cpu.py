"""CPU suite, part 2: host logic, the C-ABI exports, and the multi-rank plumbing on gloo (no kernel is launched here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from common import ROOT, ARG_ORDER, build_model


def test_library_exports_every_declared_symbol():
    from srvp_b200 import _lib
    lib = _lib.lib()
    hdr = open(os.path.join(ROOT, 'include', 'srvp_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(srvp_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in include/srvp_b200.h but not exported'
    assert set(_lib.EXPORTS) == declared
    assert lib.srvp_version() == 100
    assert isinstance(lib.srvp_last_error(), bytes)


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors must have the C layout (checked against a tiny C program compiled with gcc)."""
    from srvp_b200 import _lib
    src = '#include <stdio.h>\n#include "srvp_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(srvp_conv_src), ' \
          'sizeof(srvp_conv3x3_args), sizeof(srvp_wgrad3x3_args), sizeof(srvp_bn_bwd_args), sizeof(srvp_gemm_args), ' \
          'sizeof(srvp_latent_fwd_args), sizeof(srvp_latent_bwd_args), sizeof(srvp_linear_args), sizeof(srvp_pack_job));return 0;}\n'
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, 't.c'), 'w').write(src)
        subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), os.path.join(d, 't.c'), '-o', os.path.join(d, 't')], check=True)
        sizes = [int(v) for v in subprocess.run([os.path.join(d, 't')], capture_output=True, text=True, check=True).stdout.split()]
    mirrors = [_lib.ConvSrc, _lib.Conv3x3Args, _lib.Wgrad3x3Args, _lib.BnBwdArgs, _lib.GemmArgs, _lib.LatentFwdArgs, _lib.LatentBwdArgs, _lib.LinearArgs,
               _lib.PackJob]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_no_cpu_fallback():
    """The product path refuses to run without CUDA instead of silently falling back."""
    cfg = dict(nx=64, nc=1, nf=64, nhx=128, ny=20, nz=20, skipco=True, nt_inf=2, nh_inf=256, nlayers_inf=3, nh_res=512, nlayers_res=4,
               archi='vgg')
    m = build_model(cfg, 1.41, 0)
    with pytest.raises(RuntimeError):
        m.encoder(torch.rand(2, 1, 64, 64))


def test_class_surface_matches_reference_contract():
    """Attributes, sub-module names and method names used by train.py / test.py."""
    cfg = dict(nx=64, nc=3, nf=64, nhx=128, ny=50, nz=50, skipco=True, nt_inf=2, nh_inf=256, nlayers_inf=3, nh_res=512, nlayers_res=4,
               archi='vgg')
    m = build_model(cfg, 1.41, 0)
    for a in ['nx', 'nc', 'ny', 'nz', 'skipco', 'nt_inf', 'nh_inf', 'nlayers_inf', 'nh_res', 'nlayers_res', 'nhx']:
        assert getattr(m, a) == cfg[a]
    assert [n for n, _ in m.named_children()] == ['encoder', 'decoder', 'w_proj', 'w_inf', 'q_y', 'inf_z', 'q_z', 'p_z', 'dynamics']
    for meth in ['init', 'encode', 'decode', 'infer_w', 'infer_y', 'infer_z', '_residual_step', 'generate', 'forward']:
        assert callable(getattr(m, meth))
    assert sum(p.numel() for p in m.parameters()) == 23847390          # SURVEY.md App. E (VGG nc=3)
    assert len(m.state_dict()) == 159
    with pytest.raises(ValueError):
        from srvp_b200.module import conv
        conv.encoder_factory('resnet', 64, 3, 128, 64)
    sbn = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m)               # train.py:283 walks the BatchNorm2d children
    assert any(isinstance(x, torch.nn.SyncBatchNorm) for x in sbn.modules())


def test_rng_consumption_order_matches_reference():
    """Host draws (skip frame, infer_w frames, noise) come from torch's CPU generator in the reference's order (App. D)."""
    from oracle import srvp_oracle as O
    cfg = dict(skipco=True, nt_inf=2, ny=5, nz=7)
    torch.manual_seed(3)
    r = O.draw_randoms(cfg, 6, 6, 4, training=True)
    torch.manual_seed(3)
    t_skip = torch.randint(6, size=(4,))
    t_w = torch.stack([torch.randperm(6)[:2] for _ in range(4)], 1)
    e_y = torch.empty(4, 5).normal_()
    e_z = [torch.empty(4, 7).normal_() for _ in range(5)]
    assert torch.equal(r['t_skip'], t_skip) and torch.equal(r['t_w'], t_w) and torch.equal(r['eps_y'], e_y)
    assert all(torch.equal(a, b) for a, b in zip(r['eps_z'], e_z))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from srvp_b200 import parallel
    # batch sharding: every rank gets a contiguous slice of the videos; together they cover the global batch exactly once
    lo, hi = parallel.shard_bounds(10, rank, world)
    # gradient averaging over one flat bucket
    g = [torch.full((3,), float(rank + 1)), torch.full((2, 2), float(10 * (rank + 1)))]
    parallel.allreduce_mean_(g)
    # batch-norm statistic exchange: (sum, sumsq, count) partials are summed over ranks
    part = torch.tensor([[[1.0 + rank, 2.0 + rank]]])
    tot, cnt = parallel.allreduce_bn_partial(part, 5.5)
    q.put((rank, lo, hi, g[0].tolist(), g[1].flatten().tolist(), tot.flatten().tolist(), cnt))
    dist.destroy_process_group()


def test_two_rank_plumbing_on_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 5, 5, 10)
    for r in res:
        assert r[3] == [1.5] * 3 and r[4] == [15.0] * 4
        assert r[5] == [3.0, 5.0] and r[6] == 11.0


def test_dcgan_class_surface_and_no_cpu_fallback():
    """DCGAN64 containers: state-dict contract of SURVEY.md App. E and the same refusal to run without CUDA."""
    cfg = dict(nx=64, nc=1, nf=64, nhx=128, ny=20, nz=20, skipco=False, nt_inf=5, nh_inf=256, nlayers_inf=3, nh_res=512, nlayers_res=4,
               archi='dcgan')
    m = build_model(cfg, 1.41, 0)
    assert sum(p.numel() for p in m.parameters()) == 10678284 and len(m.state_dict()) == 82
    keys = set(m.state_dict().keys())
    for k in ['encoder.conv.0.0.weight', 'encoder.conv.3.1.running_var', 'encoder.last_conv.0.weight', 'decoder.first_upconv.0.weight',
              'decoder.conv.2.1.bias', 'decoder.conv.3.weight', 'inf_z.weight_hh_l0', 'q_y.module.2.1.weight']:
        assert k in keys, k
    assert tuple(m.state_dict()['decoder.conv.3.weight'].shape) == (64, 1, 4, 4)
    with pytest.raises(RuntimeError):
        m.encoder(torch.rand(2, 1, 64, 64))
    with pytest.raises((RuntimeError, AssertionError)):
        m(torch.rand(3, 2, 1, 64, 64), 3, dt=1.0)


def test_4x4_stride2_tap_tables_match_the_convolution_arithmetic():
    """Host functions of the 4x4 stride-2 family: the tap masks the library hands to the kernels are exactly the (phase, tap) pairs for
    which a 4x4 / stride 2 / pad 1 (transposed) convolution has a weight, and using ONLY those pairs reproduces F.conv2d /
    F.conv_transpose2d (the space-to-depth restatement of include/srvp_b200.h, checked here in fp32 on the CPU)."""
    import torch.nn.functional as F
    from srvp_b200 import _lib, ops
    down = lambda py, ty: 2 * ty + py - 1
    up = lambda py, ty: py + 3 - 2 * ty
    for kind, f in ((_lib.W4_DOWN, down), (_lib.W4_UP_PHASE, up)):
        for py in range(2):
            for px in range(2):
                want = sum(1 << (ty * 3 + tx) for ty in range(3) for tx in range(3) if 0 <= f(py, ty) < 4 and 0 <= f(px, tx) < 4)
                assert ops.tap_mask4(kind, py, px) == want and bin(want).count('1') == 4
    g = torch.Generator().manual_seed(0)
    C, Co, H = 3, 5, 8
    x, w, wt = torch.randn(2, C, H, H, generator=g), torch.randn(Co, C, 4, 4, generator=g), torch.randn(C, Co, 4, 4, generator=g)
    xs = torch.cat([x[:, :, py::2, px::2] for py in range(2) for px in range(2)], 1)
    w3 = torch.zeros(Co, 4 * C, 3, 3)
    for ph in range(4):
        m = ops.tap_mask4(_lib.W4_DOWN, ph >> 1, ph & 1)
        for tap in range(9):
            if m >> tap & 1:
                w3[:, ph * C:(ph + 1) * C, tap // 3, tap % 3] = w[:, :, down(ph >> 1, tap // 3), down(ph & 1, tap % 3)]
    assert torch.allclose(F.conv2d(xs, w3, None, 1, 1), F.conv2d(x, w, None, 2, 1), atol=1e-5)
    out = torch.zeros(2, Co, 2 * H, 2 * H)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        m = ops.tap_mask4(_lib.W4_UP_PHASE, py, px)
        w3 = torch.zeros(Co, C, 3, 3)
        for tap in range(9):
            if m >> tap & 1:
                w3[:, :, tap // 3, tap % 3] = wt[:, :, up(py, tap // 3), up(px, tap % 3)].t()
        out[:, :, py::2, px::2] = F.conv2d(x, w3, None, 1, 1)
    assert torch.allclose(out, F.conv_transpose2d(x, wt, None, 2, 1), atol=1e-5)


def test_host_side_sizing_functions():
    """Grid / buffer sizing entry points are pure host code: consistent with what the wrappers allocate."""
    from srvp_b200 import _lib
    lib = _lib.lib()
    c_int = _lib.c_int
    # statistics rows = persistent grid size, never more than the tile count
    assert lib.srvp_conv3x3_num_mtiles(c_int(1), c_int(8), c_int(8), c_int(64), c_int(64)) == 1
    big = lib.srvp_conv3x3_num_mtiles(c_int(2304), c_int(64), c_int(64), c_int(64), c_int(64))
    assert 1 < big <= 1024 and big == lib.srvp_num_sms()
    assert lib.srvp_conv3x3_nblock(c_int(16)) == 16 and lib.srvp_conv3x3_nblock(c_int(512)) == 256
    rows = lib.srvp_bn_bwd_reduce_rows(c_int(2304), c_int(64), c_int(64), c_int(64), c_int(0))
    assert 1 <= rows <= 2048 and rows * 64 <= (1 << 20)
    assert lib.srvp_bn_bwd_reduce_rows(c_int(1), c_int(4), c_int(4), c_int(512), c_int(0)) == 1
    assert lib.srvp_adam_chunk() > 0 and lib.srvp_peer_bn_buffer_bytes() > 2 * 1024 * 16
    assert lib.srvp_pack_linear_size(c_int(512), c_int(100)) >= 512 * 100


def test_peer_bn_exchange_is_gated():
    """The peer-memory SyncBatchNorm exchange is opt-in and never used without an NCCL process group."""
    from srvp_b200 import parallel
    assert parallel.PeerBN.get() is None
    os.environ['SRVP_BN_P2P'] = '1'
    try:
        assert parallel.PeerBN.get() is None      # no process group here
    finally:
        del os.environ['SRVP_BN_P2P']


def test_pack_job_descriptions_and_pack_concatenation():
    """Host logic of the packed-weight cache (ops._pack_job: what one srvp_pack_job of srvp_pack_conv3x3_multi describes) and of
    ops.concat_packs (K stages of several operands back to back inside every N block); no kernel is launched."""
    from srvp_b200 import ops
    w = torch.zeros(128, 192, 3, 3)
    # whole weight, forward and data-gradient operand
    assert ops._pack_job(w, 'conv', None)[1:] == (128, 128, 192, 192, 192 * 9, 9, 0)
    assert ops._pack_job(w, 'conv_dgrad', None)[1:] == (192, 192, 128, 128, 9, 192 * 9, 1)
    # input-channel ranges (the two halves of a convolution over cat[h, skip]): K of the forward operand, N of the data-gradient one
    p0, n_real, n_pad, k_real, k_pad, sn, sk, flip = ops._pack_job(w, 'conv', (64, 128))
    assert p0 == w.data_ptr() + 4 * 64 * 9 and (n_real, n_pad, k_real, k_pad, sn, sk, flip) == (128, 128, 128, 128, 192 * 9, 9, 0)
    assert ops._pack_job(w, 'conv_dgrad', (0, 64))[1:5] == (64, 64, 128, 128)
    # thin operands are padded to 16, everything else to multiples of 64
    assert ops._pack_job(torch.zeros(64, 3, 3, 3), 'conv', None)[2:5] == (64, 3, 16)
    assert ops._pack_job(torch.zeros(64, 3, 3, 3), 'convT', None)[1:5] == (3, 16, 64, 64)
    # concatenation: cout = 512 has two N blocks of 256; every block gets [stages of a | stages of b]
    a = torch.arange(2 * 6, dtype=torch.float32).view(2, 6).reshape(-1)         # 2 blocks x 6 elements
    b = 100 + torch.arange(2 * 4, dtype=torch.float32).view(2, 4).reshape(-1)   # 2 blocks x 4 elements
    cat = ops.concat_packs([a, b], 512)
    assert cat.tolist() == [0, 1, 2, 3, 4, 5, 100, 101, 102, 103, 6, 7, 8, 9, 10, 11, 104, 105, 106, 107]
    assert ops.concat_packs([a, b], 128).tolist() == a.tolist() + b.tolist()    # one N block: plain concatenation


def test_gradient_sinks_and_deferred_join_flag():
    """parallel.GradBucket attaches flat-buffer views as p.grad / p._srvp_sink and switches the engine to the deferred join;
    infer._direct_sinks / ops.grad_target only treat a parameter as an in-place sink while p.grad still IS that view."""
    from srvp_b200 import ops, parallel, infer
    old = ops.DEFER_JOIN
    try:
        ops.DEFER_JOIN = False
        ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7))]
        assert infer._direct_sinks(*ps) is None and ops.grad_target(ps[0])[1] is False
        b = parallel.GradBucket(ps)
        assert ops.DEFER_JOIN is True
        assert b.flat.numel() == 16 + 8 and all(p.grad is v for p, v in zip(ps, b.views))     # 16-byte aligned views
        sinks = infer._direct_sinks(ps[0], None, ps[1])
        assert sinks[0] is b.views[0] and sinks[1] is None and sinks[2] is b.views[1]
        assert ops.grad_target(ps[1]) == (b.views[1], True)
        ps[1].grad = torch.zeros(7)                       # somebody replaced the gradient: no longer an in-place sink
        assert infer._direct_sinks(*ps) is None and ops.grad_target(ps[1])[1] is False
        b.zero()                                          # re-attaches the views
        assert infer._direct_sinks(*ps) is not None
        frozen = torch.nn.Parameter(torch.zeros(3), requires_grad=False)
        frozen._srvp_sink = frozen.grad = None
        assert infer._direct_sinks(frozen) is None
    finally:
        ops.DEFER_JOIN = old
