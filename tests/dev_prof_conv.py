import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srvp_b200 import ops
dev = 'cuda'
cfgs = {'e01': (2304, 64, 64, [64], 64, [0]), 'e12': (2304, 32, 32, [128], 128, [0]), 'd30': (2304, 64, 64, [64, 64], 64, [2, 0]),
        'e22': (2304, 16, 16, [256], 256, [0]), 'd00': (2304, 8, 8, [512, 512], 512, [2, 0])}
frames, H, W, cins, cout, modes = cfgs[sys.argv[1]]
srcs = []
for cin, mode in zip(cins, modes):
    Hs, Ws = (H * 2, W * 2) if mode == 1 else (H // 2, W // 2) if mode == 2 else (H, W)
    z = torch.randn(frames, Hs, Ws, cin, device=dev).to(torch.bfloat16)
    srcs.append(ops.Src(z, cin, torch.ones(cin, device=dev), torch.zeros(cin, device=dev), None, 0, mode, True))
w = torch.randn(cout, sum(cins), 3, 3, device=dev) * 0.05
wp = ops.pack_conv3x3(w, 'conv')
for _ in range(3):
    ops.conv3x3(srcs, wp, frames, H, W, cout, stats=True, save_input=True)
torch.cuda.synchronize()
