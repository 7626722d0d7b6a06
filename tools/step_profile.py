"""Per-launch table of one BAIR training step (CUDA events around every wrapper call; run on the GPU box):
   python tools/step_profile.py [--top 60]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from srvp_b200 import ops, parallel
from srvp_b200.optim import Adam

ap = argparse.ArgumentParser()
ap.add_argument('--top', type=int, default=80)
a = ap.parse_args()
dev = torch.device('cuda', 0)
model = bench.make_model(dev, 1)
params = list(model.parameters())
opt = Adam(params, lr=3e-4)
bucket = parallel.GradBucket(params, early=list(model.decoder.parameters()))
parallel.ACTIVE_BUCKET = bucket
x = torch.rand(bench.SEQ_LEN, bench.BATCH, 3, 64, 64, device=dev)


def step():
    bucket.zero()
    out = model(x, bench.SEQ_LEN, dt=bench.DT)
    loss = bench.elbo_loss(out, x)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
ops.PROFILE = {}
step()
torch.cuda.synchronize()
rows = []
for name, recs in ops.PROFILE.items():
    if name.startswith('hbm:'):
        continue
    for r in recs:
        ms = r[0].elapsed_time(r[1])
        rows.append((ms, name, r[5], r[6], r[2], r[3]))
ops.PROFILE = None
tot = sum(r[0] for r in rows)
print(f'total {tot:.2f} ms over {len(rows)} wrapper calls')
for ms, name, tag, desc, fl, by in sorted(rows, key=lambda r: -r[0])[:a.top]:
    extra = (f'{fl / ms / 1e9:7.0f} TF/s' if fl else ' ' * 12) + (f' {by / ms / 1e6:6.0f} GB/s' if by else '')
    print(f'{ms:7.3f} ms  {name:14s} {tag:12s} {extra}  {desc}')
