"""Ad-hoc GPU check of the fused latent forward kernel against the torch loop (fp32)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from srvp_b200 import latent
from srvp_b200.module.mlp import MLP

torch.manual_seed(0)
dev = 'cuda'


def ref_loop(p_z, dyn, y0, z_post, eps, nt, os_, dt, n_post):
    y, ys, pzs, zs, ress = y0, [y0], [], [], []
    for s in range(os_ * (nt - 1)):
        fr = s // os_
        if s % os_ == 0:
            pp = p_z(y)
            pzs.append(pp)
            if fr < n_post:
                z = z_post[fr]
            else:
                mu, rho = pp.chunk(2, -1)
                z = mu + (F.softplus(rho) + 1e-8) * eps[fr]
            zs.append(z)
        r = dt * dyn(torch.cat([y, zs[-1]], 1))
        y = y + r
        ys.append(y)
        ress.append(r)
    return torch.stack(ys), torch.stack(pzs), torch.stack(zs), torch.stack(ress)


def run(B, ny, nz, nh, nl, nt, os_, n_post):
    p_z = MLP(ny, nh, 2 * nz, nl).to(dev)
    dyn = MLP(ny + nz, nh, ny, nl).to(dev)
    for l in dyn.linears():
        torch.nn.init.orthogonal_(l.weight, gain=1.41)
    y0 = torch.randn(B, ny, device=dev)
    z_post = torch.randn(max(n_post, 1), B, nz, device=dev)
    eps = torch.randn(nt - 1, B, nz, device=dev)
    dt = 1.0 / os_
    with torch.no_grad():
        ry, rp, rz, rr = ref_loop(p_z, dyn, y0, z_post, eps, nt, os_, dt, n_post)
    out = latent.latent_fwd(p_z.linears(), dyn.linears(), y0, z_post, eps, nt, os_, dt, n_post, nh)
    torch.cuda.synchronize()
    def rel(a, b):
        return float((a - b).norm() / b.norm())
    e = dict(y=rel(out['y_all'], ry), pz=rel(out['pz'], rp), z=rel(out['z'], rz), res=rel(out['res'], rr))
    ok = all(v < 2e-2 for v in e.values())
    print(f'B={B} ny={ny} nz={nz} nh={nh} nl={nl} nt={nt} os={os_} n_post={n_post}:', {k: f'{v:.2e}' for k, v in e.items()}, 'PASS' if ok else 'FAIL', flush=True)
    # timing
    for _ in range(2):
        latent.latent_fwd(p_z.linears(), dyn.linears(), y0, z_post, eps, nt, os_, dt, n_post, nh)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        latent.latent_fwd(p_z.linears(), dyn.linears(), y0, z_post, eps, nt, os_, dt, n_post, nh)
    e1.record(); torch.cuda.synchronize()
    print(f'   {e0.elapsed_time(e1) / 5:.3f} ms per call (incl. weight packing)')
    return ok


ok = run(16, 50, 50, 512, 4, 4, 2, 3)
ok &= run(37, 20, 20, 512, 4, 6, 1, 5)
ok &= run(192, 50, 50, 512, 4, 12, 2, 11)
ok &= run(24, 50, 50, 512, 4, 10, 2, 4)   # eval-style: prior sampling beyond frame 4
ok &= run(8, 50, 50, 256, 3, 5, 2, 4)
print('ALL PASS' if ok else 'SOME FAILED')


class Q(torch.nn.Module):
    """MLP forward with bf16-rounded operands (straight-through), the arithmetic of the kernels."""
    def __init__(self, mlp):
        super().__init__()
        self.mlp = mlp
    def forward(self, x):
        rq = lambda t: t + (t.detach().to(torch.bfloat16).float() - t.detach())
        for i, lin in enumerate(self.mlp.linears()):
            if i > 0:
                x = F.relu(x)
            x = F.linear(rq(x), rq(lin.weight), lin.bias)
        return x


def run_bwd(B, ny, nz, nh, nl, nt, os_):
    p_z = MLP(ny, nh, 2 * nz, nl).to(dev)
    dyn = MLP(ny + nz, nh, ny, nl).to(dev)
    for l in dyn.linears():
        torch.nn.init.orthogonal_(l.weight, gain=1.41)
    n_post = nt - 1
    y0 = torch.randn(B, ny, device=dev, requires_grad=True)
    z_post = torch.randn(n_post, B, nz, device=dev, requires_grad=True)
    dt = 1.0 / os_
    S = os_ * (nt - 1)
    ry, rp, rz, rr = ref_loop(Q(p_z), Q(dyn), y0, z_post, None, nt, os_, dt, n_post)
    g_y = torch.zeros(S + 1, B, ny, device=dev)
    g_y[::os_] = torch.randn(nt, B, ny, device=dev)       # only integer-time states are consumed downstream
    g_res = torch.randn(S, B, ny, device=dev) * 0.1
    g_pz = torch.randn(nt - 1, B, 2 * nz, device=dev) * 0.1
    ((ry * g_y).sum() + (rr * g_res).sum() + (rp * g_pz).sum()).backward()
    with torch.no_grad():
        fwd = latent.latent_fwd(p_z.linears(), dyn.linears(), y0.detach(), z_post.detach(), None, nt, os_, dt, n_post, nh)
        d_y0, d_z, gp, gd = latent.latent_bwd(p_z.linears(), dyn.linears(), fwd, g_y, g_res, g_pz, nt, os_, dt, nh)
    torch.cuda.synchronize()
    def rel(a, b):
        return float((a - b).norm() / b.norm())
    e = dict(d_y0=rel(d_y0, y0.grad), d_z=rel(d_z, z_post.grad))
    for i, (lin, (dW, db)) in enumerate(zip(p_z.linears(), gp)):
        e[f'pz.W{i}'] = rel(dW, lin.weight.grad); e[f'pz.b{i}'] = rel(db, lin.bias.grad)
    for i, (lin, (dW, db)) in enumerate(zip(dyn.linears(), gd)):
        e[f'dyn.W{i}'] = rel(dW, lin.weight.grad); e[f'dyn.b{i}'] = rel(db, lin.bias.grad)
    ok = all(v < 3e-2 for v in e.values())
    print(f'BWD B={B} ny={ny} nz={nz} nh={nh} nl={nl} nt={nt} os={os_}:', {k: f'{v:.1e}' for k, v in e.items()}, 'PASS' if ok else 'FAIL', flush=True)
    return ok


okb = run_bwd(16, 50, 50, 512, 4, 4, 2)
okb &= run_bwd(37, 20, 20, 512, 4, 6, 1)
okb &= run_bwd(192, 50, 50, 512, 4, 12, 2)
okb &= run_bwd(8, 50, 50, 256, 3, 5, 2)
print('BWD ALL PASS' if okb else 'BWD SOME FAILED')
