"""Ad-hoc: where does the step time go outside the library kernels? (torch profiler summary on the GPU)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from srvp_b200.module.srvp import StochasticLatentResidualVideoPredictor
torch.manual_seed(1)
m = StochasticLatentResidualVideoPredictor(*[bench.CFG[k] for k in bench.ARG_ORDER]); m.init(); m = m.cuda().train(); m.noise_device = 'cuda'
opt = torch.optim.Adam(m.parameters(), lr=3e-4, fused=True)
x = torch.rand(12, 192, 3, 64, 64, device='cuda')
def step():
    opt.zero_grad(set_to_none=True)
    out = m(x, 12, dt=0.5)
    loss = bench.elbo_loss(out, x)
    loss.backward()
    opt.step()
for _ in range(4): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(ev, key=lambda e: -e.device_time_total)[:45]
tot = sum(e.device_time_total for e in ev if e.device_type == torch.autograd.DeviceType.CUDA) if False else None
for e in rows:
    print(f'{e.key[:70]:70s} n={e.count:5d} cuda={e.device_time_total/1e3:9.3f} ms cpu={e.cpu_time_total/1e3:9.3f} ms')
