"""Ad-hoc timing of the weight-gradient launches of one BAIR training step, layer by layer (CUDA events, L2 flushed between launches):
   python tests/dev_wgrad_layers.py            # current dispatch (TMA-fed kernel where eligible)
   SRVP_WGRAD_TMA=0 python tests/dev_wgrad_layers.py   # the cp.async kernel everywhere"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srvp_b200 import ops

F_ = int(os.environ.get('FRAMES', 2304))
LAYERS = [('64->64@64', 64, 64, 64, F_), ('64->128@32', 64, 128, 32, F_), ('128->64@32', 128, 64, 32, F_), ('128->128@32', 128, 128, 32, F_),
          ('128->256@16', 128, 256, 16, F_), ('256->128@16', 256, 128, 16, F_), ('256->256@16', 256, 256, 16, F_), ('256->512@8', 256, 512, 8, F_),
          ('512->256@8', 512, 256, 8, F_), ('512->512@8', 512, 512, 8, F_), ('skip 512->512@8 (B frames)', 512, 512, 8, F_ // 12),
          ('skip 64->64@64 (B frames)', 64, 64, 64, F_ // 12)]
sel = os.environ.get('LAYERS')
if sel:
    LAYERS = [l for l in LAYERS if l[0] in sel.split(',')]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
tot = 0.0
for name, cin, cout, H, fr in LAYERS:
    a = (torch.randn(fr, H, H, cin, device='cuda') * 0.5).to(torch.bfloat16)
    dz = (torch.randn(fr, H, H, cout, device='cuda') * 0.1).to(torch.bfloat16)
    dw = torch.zeros(cout, cin, 3, 3, device='cuda')
    for _ in range(2):
        ops.wgrad3x3(a, cin, dz, cout, fr, H, H, cout, cin, dw, 'conv')
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.wgrad3x3(a, cin, dz, cout, fr, H, H, cout, cin, dw, 'conv')
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * fr * H * H * cin * cout * 9
    tot += ms
    print(f'{name:28s} {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s', flush=True)
print(f'sum {tot:.3f} ms')
