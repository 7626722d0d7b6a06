"""Ad-hoc: CUDA-event time of each segment of one training step (BAIR config) + host-side time to enqueue it."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from srvp_b200.module.srvp import StochasticLatentResidualVideoPredictor
torch.manual_seed(1)
m = StochasticLatentResidualVideoPredictor(*[bench.CFG[k] for k in bench.ARG_ORDER]); m.init(); m = m.cuda().train(); m.noise_device = 'cuda'
from srvp_b200.optim import Adam
opt = Adam(m.parameters(), lr=3e-4)
x = torch.rand(12, 192, 3, 64, 64, device='cuda')
nt, dt = 12, 0.5
marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e, time.perf_counter()))
def step(rec=False):
    global marks
    marks = []
    mark('start')
    opt.zero_grad(set_to_none=True); mark('zero_grad')
    hx, handle = m._encode_fused(x); mark('encode')
    w = m.infer_w(hx); y_0, q_y_0_params = m.infer_y(hx[:m.nt_inf]); mark('infer_w_y')
    y, z, q_z_params, p_z_params, res = m.generate(y_0, hx, nt, dt); mark('generate')
    x_ = m._decode_fused(w, y, handle.levels, handle.frame_map, handle); mark('decode')
    loss = bench.elbo_loss((x_, y, z, w, q_y_0_params, q_z_params, p_z_params, res), x); mark('elbo')
    loss.backward(); mark('backward')
    opt.step(); mark('adam')
    return marks
for _ in range(4): step()
torch.cuda.synchronize()
acc = {}
N = 5
for _ in range(N):
    mk = step(); torch.cuda.synchronize()
    for (n0, e0, t0), (n1, e1, t1) in zip(mk[:-1], mk[1:]):
        a = acc.setdefault(n1, [0.0, 0.0]); a[0] += e0.elapsed_time(e1) / N; a[1] += (t1 - t0) * 1e3 / N
    a = acc.setdefault('TOTAL', [0.0, 0.0]); a[0] += mk[0][1].elapsed_time(mk[-1][1]) / N; a[1] += (mk[-1][2] - mk[0][2]) * 1e3 / N
for k, (g, c) in acc.items():
    print(f'{k:12s} gpu {g:8.3f} ms   host-enqueue {c:8.3f} ms')
