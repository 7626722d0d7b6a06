"""Ad-hoc GPU unit checks of the backward building blocks against torch autograd (fp32 on bf16-rounded inputs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from srvp_b200 import ops, _lib

dev = 'cuda'
torch.manual_seed(0)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


class FakeBN:
    pass


def test_bn_bwd(mode, Fr=6, H=16, W=16, C=64, with_skip=False):
    z = torch.randn(Fr, H, W, C, device=dev).to(torch.bfloat16)
    gamma = (torch.rand(C, device=dev) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device=dev) * 0.2).requires_grad_(True)
    zf = z.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    y = F.batch_norm(zf, None, None, gamma, beta, True, 0.0, 1e-5)
    a = F.leaky_relu(y, 0.2)
    a_r = a.detach().to(torch.bfloat16).float() + (a - a.detach())  # forward pool compares bf16-rounded values
    if mode == 1:
        out = F.max_pool2d(a_r, 2)
    elif mode == 2:
        out = F.interpolate(a, scale_factor=2, mode='nearest')
    else:
        out = a
    da = (torch.randn_like(out) * 0.1).to(torch.bfloat16)
    # our forward stats
    st = ops.BNState(C, dev)
    bn = FakeBN()
    bn.weight, bn.bias, bn.running_mean, bn.running_var = gamma.detach(), beta.detach(), None, None
    partial = ops.channel_stats(z.view(-1, C))
    ops.bn_finalize(partial, float(Fr * H * W), bn, st, training_update=False)
    loss = (out * da.float()).sum()
    nt, B = 3, 2
    skip = inv = None
    if with_skip:
        skip = (torch.randn(nt * B, H, W, 2 * C, device=dev) * 0.1).to(torch.bfloat16)
        inv = torch.full((Fr,), -1, dtype=torch.int32, device=dev)
        inv[1] = 0
        inv[4] = 1
        sk = skip[..., C:].float().view(nt, B, H, W, C).sum(0).permute(0, 3, 1, 2)  # (B,C,H,W)
        loss = loss + (a[1] * sk[0]).sum() + (a[4] * sk[1]).sum()
    loss.backward()
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dz = ops.bn_bwd(z, st, gamma.detach(), dg, db, da.permute(0, 2, 3, 1).contiguous(), mode, Fr, H, W, C, skip=skip, skip_coff=C, nt=nt, B=B, inv_map=inv)
    torch.cuda.synchronize()
    ref = zf.grad.permute(0, 2, 3, 1)
    print(f'bn_bwd mode={mode} skip={with_skip}: dz {rel(dz.float(), ref):.3e} dgamma {rel(dg, gamma.grad):.3e} dbeta {rel(db, beta.grad):.3e}', flush=True)


def test_dgrad(kind, Fr=4, H=16, W=16, cin=128, cout=64):
    a = torch.randn(Fr, cin, H, W, device=dev).to(torch.bfloat16).float().requires_grad_(True)
    if kind == 'conv':
        w = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05)
        out = F.conv2d(a, w.to(torch.bfloat16).float(), padding=1)
    else:
        w = (torch.randn(cin, cout, 3, 3, device=dev) * 0.05)
        out = F.conv_transpose2d(a, w.to(torch.bfloat16).float(), padding=1)
    dz = (torch.randn_like(out) * 0.1).to(torch.bfloat16)
    out.backward(dz.float())
    wp = ops.pack_conv3x3(w, kind + '_dgrad')
    cpad = ops.padded_k(cout)
    dzn = torch.zeros(Fr, H, W, cpad, device=dev, dtype=torch.bfloat16)
    dzn[..., :cout] = dz.permute(0, 2, 3, 1)
    da, _ = ops.conv3x3([ops.Src(dzn, cpad)], wp, Fr, H, W, cin)
    torch.cuda.synchronize()
    print(f'dgrad {kind} {cin}->{cout}: {rel(da.float().permute(0, 3, 1, 2), a.grad):.3e}', flush=True)


test_dgrad('conv')
test_dgrad('conv', cin=64, cout=256)
test_dgrad('convT', cin=64, cout=3)
for mode in (0, 1, 2):
    test_bn_bwd(mode)
test_bn_bwd(0, with_skip=True)
test_bn_bwd(1, with_skip=True)
test_bn_bwd(0, C=512, H=8, W=8)
