"""Ad-hoc: per-frame KL(z) deviation from the fp32 oracle at the KTH shape (T = 20, untrained weights): how the chaotic growth of the
latent state amplifies bf16 rounding along the sequence."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch.nn.functional as F
from common import build_model, make_input
from oracle import srvp_oracle as O
from test_gpu_parity_full import SHAPES
cfg, loss_cfg, res_gain, T, B, dt = SHAPES['kth_shape']
m = build_model(cfg, res_gain, seed=1)
sd = {k: v.clone() for k, v in m.state_dict().items()}
m = m.cuda().train()
x = make_input(T, B, cfg['nc'], seed=123)
def klz(q, p):
    lq, rq = q.chunk(2, -1); lp, rp = p.chunk(2, -1)
    return O.kl_normal(lq, F.softplus(rq) + 1e-8, lp, F.softplus(rp) + 1e-8).sum((1, 2))
with torch.no_grad():
    torch.manual_seed(7)
    out = m(x.cuda(), T, dt=dt)
    torch.manual_seed(7)
    rnd = O.draw_randoms(cfg, T, T, B, training=True)
    o = O.forward(sd, cfg, x, T, dt, rnd, training=True)
ours = klz(out[5].cpu(), out[6].cpu()); ref = klz(o['q_z_params'], o['p_z_params'])
print('THIN', os.environ.get('SRVP_THIN', '1'))
for t in range(ours.shape[0]):
    print(f't={t + 1:2d} KL_z ours {float(ours[t]):14.2f} ref {float(ref[t]):14.2f} rel {abs(float(ours[t] - ref[t])) / abs(float(ref[t])):.2e}  |y| ref {float(o["y"][t + 1].abs().mean()):.2f}')
print('total rel', abs(float(ours.sum() - ref.sum())) / float(ref.sum()))
