"""cProfile of the host side of BAIR training steps (run on the GPU box): where the ~10 ms of Python per step go.
   python tools/host_profile.py [--batch 192]"""
import argparse, cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from srvp_b200 import ops, parallel
from srvp_b200.optim import Adam

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=192)
ap.add_argument('--steps', type=int, default=10)
a = ap.parse_args()
dev = torch.device('cuda', 0)
model = bench.make_model(dev, 1)
params = list(model.parameters())
opt = Adam(params, lr=3e-4)
bucket = parallel.GradBucket(params, early=list(model.decoder.parameters()))
parallel.ACTIVE_BUCKET = bucket
x = torch.rand(bench.SEQ_LEN, a.batch, 3, 64, 64, device=dev)


def step():
    bucket.zero()
    out = model(x, bench.SEQ_LEN, dt=bench.DT)
    loss = bench.elbo_loss(out, x)
    loss.backward()
    opt.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(a.steps):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
st.sort_stats('cumtime').print_stats(45)
