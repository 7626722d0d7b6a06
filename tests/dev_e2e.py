"""Ad-hoc end-to-end parity run on the GPU: our model vs golden fixtures and the live CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from common import *
from oracle import srvp_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else 'vgg_skip_nc3'
g = load_golden(name)
cfg = g['cfg']
m = build_model(cfg, g['res_gain'], g['seeds']['model'])
sd0 = {k: v.clone() for k, v in m.state_dict().items()}
m = m.cuda().train()
x = make_input(g['T'], g['B'], cfg['nc'], g['seeds']['input'])
xc = x.cuda()
torch.manual_seed(g['seeds']['fwd'])
out = m(xc, g['T'], dt=g['dt'])
loss, nll, kl_y, kl_z = model_loss(out, xc, g['loss_cfg'])
loss.backward()
torch.cuda.synchronize()
t = g['train']
print(f"loss {float(loss):.4f} vs {t['loss']:.4f} rel {abs(float(loss)-t['loss'])/abs(t['loss']):.2e}")
print(f"nll  {float(nll):.4f} vs {t['nll']:.4f} rel {abs(float(nll)-t['nll'])/abs(t['nll']):.2e}")
print(f"kl_y {float(kl_y):.5f} vs {t['kl_y_0']:.5f} rel {abs(float(kl_y)-t['kl_y_0'])/abs(t['kl_y_0']):.2e}")
print(f"kl_z {float(kl_z):.4f} vs {t['kl_z']:.4f} rel {abs(float(kl_z)-t['kl_z'])/abs(t['kl_z']):.2e}")
names = ['x_', 'y', 'z', 'w', 'q_y_0_params', 'q_z_params', 'p_z_params', 'res']
for i, n in enumerate(names):
    if n in t:
        print(f'  {n:14s} rel_l2 {rel_l2(out[i], t[n]):.3e}  maxabs {float((out[i].cpu()-t[n]).abs().max()):.3e}')
print(f"  x_sub rel_l2 {rel_l2(out[0][:, :, :, ::8, ::8], t['x_sub']):.3e}; per-pixel MSE {float(((out[0][:, :, :, ::8, ::8].cpu()-t['x_sub'])**2).mean()):.3e}")
print(f"  hx rel_l2 {rel_l2(m._last_hx, t['hx']):.3e}") if hasattr(m, '_last_hx') else None
# full gradients from the live oracle
O.EMULATE_BF16 = len(sys.argv) > 2
torch.manual_seed(g['seeds']['fwd'])
rnd = O.draw_randoms(cfg, g['T'], g['T'], g['B'], training=True)
sdo = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in sd0.items()}
o = O.forward(sdo, cfg, x, g['T'], g['dt'], rnd, training=True)
ol = O.elbo(o, x, g['loss_cfg'])[0]
ol.backward()
worst = []
for k, p in m.named_parameters():
    e = rel_l2(p.grad, sdo[k].grad)
    worst.append((e, k))
worst.sort(reverse=True)
print('all grad rel_l2:')
for e, k in sorted(worst, key=lambda t: t[1]):
    print(f'   {k:40s} {e:.3e}')
print('median grad rel_l2', sorted(w[0] for w in worst)[len(worst)//2])
