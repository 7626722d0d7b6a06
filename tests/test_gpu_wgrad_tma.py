"""GPU suite, part 7: the TMA-fed weight-gradient kernel (srvp_b200/csrc/wgrad3x3_tma.cu) against torch autograd on identical bf16
operands, on the layer shapes of the VGG64 encoder / decoder and on the edge cases of its tiling: channel pitches /
offsets that differ from the channel count (split skip convolutions), odd frame counts (stripes that end mid-stage), transposed
convolutions (flipped taps), accumulation into a non-zero gradient, and agreement with the cp.async kernel it replaces."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    return 'cuda'


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-30))


def _reference(a_nchw, dz_nchw, kind, cout, cin):
    a = a_nchw.float().requires_grad_(False)
    if kind == 'conv':
        w = torch.zeros(cout, cin, 3, 3, device=a.device, requires_grad=True)
        out = F.conv2d(a, w, padding=1)
    else:
        w = torch.zeros(cin, cout, 3, 3, device=a.device, requires_grad=True)
        out = F.conv_transpose2d(a, w, padding=1)
    out.backward(dz_nchw.float())
    return w.grad


CASES = [
    # kind, cin, cout, H, frames
    ('conv', 64, 64, 64, 3),       # enc.conv.0.1 / dec.conv.3.0 halves
    ('conv', 64, 128, 32, 5),      # enc.conv.1.1
    ('conv', 128, 64, 32, 4),      # dec.conv.2.1
    ('conv', 128, 128, 32, 3),
    ('conv', 256, 256, 16, 7),     # enc.conv.2.x / dec.conv.1.x
    ('conv', 256, 128, 16, 6),
    ('conv', 512, 512, 8, 9),      # enc.conv.3.x / dec.conv.0.x (two stripes per stage, odd count: the last stage is half empty)
    ('conv', 512, 256, 8, 4),
    ('convT', 64, 64, 16, 5),      # transposed convolution: flipped taps
    ('conv', 64, 64, 4, 3),        # tiny image
]


@pytest.mark.parametrize('kind,cin,cout,H,frames', CASES)
def test_wgrad_tma_matches_autograd(dev, kind, cin, cout, H, frames):
    from srvp_b200 import ops
    torch.manual_seed(cin * 7 + cout + H)
    W = H
    a = (torch.randn(frames, cin, H, W, device=dev) * 0.5).to(torch.bfloat16)
    dz = (torch.randn(frames, cout, H, W, device=dev) * 0.1).to(torch.bfloat16)
    ref = _reference(a, dz, kind, cout, cin)
    a_n = a.permute(0, 2, 3, 1).contiguous()
    dz_n = dz.permute(0, 2, 3, 1).contiguous()
    dw = torch.zeros_like(ref)
    ops.wgrad3x3(a_n, cin, dz_n, cout, frames, H, W, cout, cin, dw, kind)
    torch.cuda.synchronize()
    assert rel(dw, ref) < 2e-3, rel(dw, ref)
    # accumulation into an existing gradient
    dw2 = ref.clone()
    ops.wgrad3x3(a_n, cin, dz_n, cout, frames, H, W, cout, cin, dw2, kind)
    assert rel(dw2, 2 * ref) < 2e-3


def test_wgrad_tma_channel_pitch_and_offsets(dev):
    """Operands that are channel slices of wider tensors, written into a slice of a wider weight (the per-video split of the
    convolutions over cat[h, skip]: engine._decoder_bwd)."""
    from srvp_b200 import ops
    frames, H, W, ch, cs, cout = 5, 16, 16, 128, 64, 128
    torch.manual_seed(3)
    a_full = (torch.randn(frames, H, W, ch + cs + 64, device=dev) * 0.5).to(torch.bfloat16)
    dz_full = (torch.randn(frames, H, W, cout + 64, device=dev) * 0.1).to(torch.bfloat16)
    a_s = a_full[..., ch:ch + cs]                       # channels [128, 192) of a 256-pitch tensor
    dz_s = dz_full[..., 64:64 + cout]
    ref = _reference(a_s.permute(0, 3, 1, 2), dz_s.permute(0, 3, 1, 2), 'conv', cout, cs)
    cin_tot = ch + cs
    dw = torch.zeros(cout, cin_tot, 3, 3, device=dev)
    ops.wgrad3x3(a_full, cs, dz_full, cout, frames, H, W, cout, cs, dw, 'conv', act_coff=ch, dz_coff=64, strides=(cin_tot * 9, 9), dw_offset=ch * 9)
    assert rel(dw[:, ch:], ref) < 2e-3
    assert float(dw[:, :ch].abs().max()) == 0.0


def test_wgrad_tma_equals_cp_async_kernel(dev):
    """Same launch through both kernels (SRVP_WGRAD_TMA=0 selects the cp.async one in a fresh process is not possible here: the
    switch is read once), so compare against autograd at two frame counts that exercise different split-K partitions instead."""
    from srvp_b200 import ops
    for frames in (1, 37):
        a = (torch.randn(frames, 32, 32, 128, device=dev) * 0.5).to(torch.bfloat16)
        dz = (torch.randn(frames, 32, 32, 128, device=dev) * 0.1).to(torch.bfloat16)
        ref = _reference(a.permute(0, 3, 1, 2), dz.permute(0, 3, 1, 2), 'conv', 128, 128)
        dw = torch.zeros_like(ref)
        ops.wgrad3x3(a, 128, dz, 128, frames, 32, 32, 128, 128, dw, 'conv')
        assert rel(dw, ref) < 2e-3


def test_psnr_ssim_kernel_matches_reference_formula(dev):
    """srvp_psnr_ssim vs the reference's formulas restated in torch (metrics/ssim.py:81-111 through test.py:36-58; test.py:249-250),
    including samples broadcast over one ground truth."""
    from srvp_b200 import metrics
    torch.manual_seed(0)
    S, T, B, C, H, W = 3, 4, 2, 3, 64, 64
    gt = torch.rand(T, B, C, H, W, device=dev)
    pred = (gt.unsqueeze(0) + 0.15 * torch.randn(S, T, B, C, H, W, device=dev))   # not clamped: the kernel clamps
    psnr, ssim = metrics.psnr_ssim(pred, gt, clamp=True)

    def ref_ssim(x, y):
        size, sigma, k1, k2 = 11, 1.5, 0.01, 0.03
        coords = torch.tensor([(i - (size - 1.) / 2.) for i in range(size)], device=x.device)
        g = (-coords ** 2 / (2. * sigma ** 2))
        grid = (g.view(1, -1) + g.view(-1, 1)).view(1, -1).softmax(-1)
        kern = grid.view(1, 1, size, size).expand(C, 1, size, size).contiguous()
        c1, c2 = k1 ** 2, k2 ** 2
        mu1, mu2 = F.conv2d(x, kern, groups=C), F.conv2d(y, kern, groups=C)
        s1 = F.conv2d(x * x, kern, groups=C) - mu1 ** 2
        s2 = F.conv2d(y * y, kern, groups=C) - mu2 ** 2
        s12 = F.conv2d(x * y, kern, groups=C) - mu1 * mu2
        v1, v2 = 2 * s12 + c2, s1 + s2 + c2
        return (((2 * mu1 * mu2 + c1) * v1) / ((mu1 ** 2 + mu2 ** 2 + c1) * v2)).mean(dim=[2, 3])

    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for s in range(S):
            xp = pred[s].clamp(0, 1)
            r_ssim = ref_ssim(xp.view(T * B, C, H, W), gt.view(T * B, C, H, W)).view(T, B, C)
            r_psnr = 10 * torch.log10(1 / ((xp - gt) ** 2).mean(dim=[3, 4]))
            assert torch.allclose(ssim[s], r_ssim, atol=2e-5, rtol=1e-4), float((ssim[s] - r_ssim).abs().max())
            assert torch.allclose(psnr[s], r_psnr, atol=1e-4, rtol=1e-5)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


@pytest.mark.parametrize('nc,frames,bn', [(3, 5, True), (1, 3, True), (3, 2, False)])
def test_decoder_head_kernel_matches_torch(dev, nc, frames, bn):
    """srvp_decoder_head_fwd (activation + tap-expanded transposed convolution + sigmoid) against torch on identical bf16 operands, and
    against the generic conv3x3 kernel's sigmoid epilogue; the activated copy it writes must equal the loader's."""
    from srvp_b200 import ops
    torch.manual_seed(nc * 10 + frames)
    z = torch.randn(frames, 64, 64, 64, device=dev).to(torch.bfloat16)
    sc = (torch.rand(64, device=dev) + 0.5) * torch.where(torch.rand(64, device=dev) < 0.2, -1.0, 1.0) if bn else None
    sh = torch.randn(64, device=dev) * 0.3 if bn else None
    w = torch.randn(64, nc, 3, 3, device=dev) * 0.1
    src = ops.Src(z, 64, sc, sh, None, 0, 0, True)
    xh, a_out = ops.decoder_head_fwd(src, w, frames, nc, save_input=True)
    zf = z.float()
    a = zf * sc + sh if bn else zf
    a = F.leaky_relu(a, 0.2).to(torch.bfloat16)
    assert torch.equal(a_out, a) or rel(a_out, a) < 1e-2
    ref = torch.sigmoid(F.conv_transpose2d(a_out.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1))
    assert float((xh - ref).abs().max()) < 2e-3, float((xh - ref).abs().max())
    r = ops.conv3x3([src], ops.pack_conv3x3(w, 'convT'), frames, 64, 64, nc, sigmoid_nchw=True, save_input=True)
    assert float((xh - r[0]).abs().max()) < 2e-3
    assert torch.equal(a_out, r[2])


def test_wgrad_tma_thin_operands(dev):
    """The HBM-bound ends: first encoder layer (16-channel padded input, 3 real channels) and last decoder layer (transposed convolution,
    dz padded to 16 channels, 3 real): the TMA boxes ask for 64 channels, the missing ones arrive as out-of-bound zeros."""
    from srvp_b200 import ops
    torch.manual_seed(11)
    frames, H = 5, 64
    # first layer: nn.Conv2d(3, 64): act = x padded to 16 channels
    x = torch.rand(frames, 3, H, H, device=dev).to(torch.bfloat16)
    dz = (torch.randn(frames, 64, H, H, device=dev) * 0.1).to(torch.bfloat16)
    ref = _reference(x, dz, 'conv', 64, 3)
    x16 = torch.zeros(frames, H, H, 16, device=dev, dtype=torch.bfloat16)
    x16[..., :3] = x.permute(0, 2, 3, 1)
    dw = torch.zeros_like(ref)
    ops.wgrad3x3(x16, 16, dz.permute(0, 2, 3, 1).contiguous(), 64, frames, H, H, 64, 3, dw, 'conv')
    assert rel(dw, ref) < 2e-3, rel(dw, ref)
    # last layer: nn.ConvTranspose2d(64, 3): dz padded to 16 channels
    a = (torch.randn(frames, 64, H, H, device=dev) * 0.5).to(torch.bfloat16)
    dz3 = (torch.randn(frames, 3, H, H, device=dev) * 0.1).to(torch.bfloat16)
    ref = _reference(a, dz3, 'convT', 3, 64)
    dz16 = torch.zeros(frames, H, H, 16, device=dev, dtype=torch.bfloat16)
    dz16[..., :3] = dz3.permute(0, 2, 3, 1)
    dw = torch.zeros_like(ref)
    ops.wgrad3x3(a.permute(0, 2, 3, 1).contiguous(), 64, dz16, 16, frames, H, H, 3, 64, dw, 'convT')
    assert rel(dw, ref) < 2e-3, rel(dw, ref)


@pytest.mark.parametrize('C,H', [(64, 16), (256, 8)])
def test_conv3x3_per_video_addend_split(dev, C, H):
    """conv(cat[h, s]) = conv_h(h) + conv_s(s) with s constant over time (engine._decoder_fwd): the per-video term goes through the fp32
    side output of one launch (channel-group planes) into the `add` input of the other; result vs F.conv2d over the concatenation."""
    from srvp_b200 import ops
    torch.manual_seed(C)
    nt, B = 3, 4
    Fr = nt * B
    h = (torch.randn(Fr, H, H, C, device=dev) * 0.5).to(torch.bfloat16)
    sk = (torch.randn(B, H, H, C, device=dev) * 0.5).to(torch.bfloat16)
    w = torch.randn(C, 2 * C, 3, 3, device=dev) * 0.03
    wb = w.to(torch.bfloat16).float()
    cat = torch.cat([h.float(), sk.float().repeat(nt, 1, 1, 1)], 3).permute(0, 3, 1, 2)
    ref = F.conv2d(cat, wb, padding=1).permute(0, 2, 3, 1)
    rs = ops.conv3x3([ops.Src(sk, C)], ops.pack_conv3x3(w, 'conv', cin_range=(C, C)), B, H, H, C, out_f32=True)
    assert tuple(rs[0].shape) == (C // 4, B * H * H, 4)
    r = ops.conv3x3([ops.Src(h, C)], ops.pack_conv3x3(w, 'conv', cin_range=(0, C)), Fr, H, H, C, stats=True, add=rs[0])
    assert rel(r[0].float(), ref) < 1.5e-2
    tot = r[1].sum(0)                                    # (C, 2): sum, sumsq of the stored bf16 values
    zs = r[0].float().view(-1, C)
    assert rel(tot[:, 0], zs.sum(0)) < 1e-3 and rel(tot[:, 1], (zs * zs).sum(0)) < 1e-3


@pytest.mark.parametrize('C,H', [(64, 64), (128, 32), (512, 8)])
def test_conv3x3_per_video_term_as_hilo_stages(dev, C, H):
    """The same split with the per-video term fed through the tensor core (engine._decoder_fwd at resolution >= 32): one launch stores
    conv_s(s) as [hi | lo] bf16, the per-frame launch reads it as a second source (frame -> video map) whose K stages use the centre
    tap against identity weights. Checks the hi/lo store (hi + lo = fp32 result to 2^-16), the final result vs F.conv2d over the
    concatenation (tighter than one bf16 rounding of the addend would allow), the statistics, and that a_out holds only h."""
    from srvp_b200 import ops
    torch.manual_seed(C + 1)
    nt, B = 3, 4
    Fr = nt * B
    h = (torch.randn(Fr, H, H, C, device=dev) * 0.5).to(torch.bfloat16)
    sk = (torch.randn(B, H, H, C, device=dev) * 0.5).to(torch.bfloat16)
    w = torch.randn(C, 2 * C, 3, 3, device=dev) * 0.03
    wb = w.to(torch.bfloat16).float()
    ref_s = F.conv2d(sk.float().permute(0, 3, 1, 2), wb[:, C:], padding=1).permute(0, 2, 3, 1)
    cat = torch.cat([h.float(), sk.float().repeat(nt, 1, 1, 1)], 3).permute(0, 3, 1, 2)
    ref = F.conv2d(cat, wb, padding=1).permute(0, 2, 3, 1)
    rs = ops.conv3x3([ops.Src(sk, C)], ops.pack_conv3x3(w, 'conv', cin_range=(C, C)), B, H, H, C, out_hilo=True)
    assert tuple(rs[0].shape) == (B, H, H, 2 * C)
    hl = rs[0].float()
    assert float(((hl[..., :C] + hl[..., C:]) - ref_s).abs().max() / ref_s.abs().max()) < 1e-4
    vid = torch.arange(B, dtype=torch.int32, device=dev).repeat(nt)
    wp = ops.concat_packs([ops.pack_conv3x3(w, 'conv', cin_range=(0, C)), ops.hilo_identity_pack(C, dev)], C)
    masks = [0x1ff] * (C // 64) + [0x010] * (2 * C // 64)
    r = ops.conv3x3([ops.Src(h, C), ops.Src(rs[0], 2 * C, None, None, vid, 0, 0, False)], wp, Fr, H, H, C, stats=True, save_input=True,
                    tap_masks=masks, a_out_channels=C, cin_real=C)
    assert rel(r[0].float(), ref) < 6e-3                 # output rounding only
    tot = r[1].sum(0)
    zs = r[0].float().view(-1, C)
    assert rel(tot[:, 0], zs.sum(0)) < 1e-3 and rel(tot[:, 1], (zs * zs).sum(0)) < 1e-3
    assert tuple(r[2].shape) == (Fr, H, H, C) and torch.equal(r[2], h)
