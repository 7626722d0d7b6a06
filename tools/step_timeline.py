"""Two-stream timeline of one BAIR training step (run on the GPU box): every wrapper call on the main stream and every weight-gradient
launch on the side stream with its start / end time relative to the start of the step, plus how much of the side stream's busy time
lies under main-stream kernels.   python tools/step_timeline.py [--from 25 --to 60]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from srvp_b200 import ops, parallel
from srvp_b200.optim import Adam

ap = argparse.ArgumentParser()
ap.add_argument('--from', dest='t0', type=float, default=0.0)
ap.add_argument('--to', dest='t1', type=float, default=1e9)
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
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True)
end = torch.cuda.Event(enable_timing=True)
ops.TIMELINE = []
base.record()
step()
end.record()
torch.cuda.synchronize()
tl = ops.TIMELINE
ops.TIMELINE = None
rows = sorted(((base.elapsed_time(e0), base.elapsed_time(e1), name, tag, strm) for name, tag, strm, e0, e1 in tl), key=lambda r: r[0])
print(f'step {base.elapsed_time(end):.3f} ms, {len(rows)} records')
main = [(s, e) for s, e, n, t, st in rows if st == 'main']
for s, e, name, tag, strm in rows:
    if e < a.t0 or s > a.t1:
        continue
    extra = ''
    if strm == 'side':
        ov = sum(max(0.0, min(e, me) - max(s, ms)) for ms, me in main)
        extra = f'  under main-stream kernels: {ov:.3f} ms'
    print(f'{"        " if strm == "side" else "                " if strm == "aux" else ""}{s:8.3f} -> {e:8.3f}  ({e - s:6.3f})  {strm:4s} {name:14s} {tag}{extra}')
side_busy = sum(e - s for s, e, n, t, st in rows if st == 'side')
print(f'side stream busy {side_busy:.3f} ms')
