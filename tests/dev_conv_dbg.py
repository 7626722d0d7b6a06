"""Ad-hoc: conv3x3 'plain' launches (data-gradient style) with the development switches of SRVP_CONV_DBG (set in the environment)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srvp_b200 import ops
dev = 'cuda'
F_ = 2304
for (nm, H, cin, cout) in [('64->64 @64', 64, 64, 64), ('128->64 @32', 32, 128, 64), ('128->128 @32', 32, 128, 128), ('256->256 @16', 16, 256, 256), ('16->64 @64 thin', 64, 16, 64)]:
    z = torch.randn(F_, H, H, cin, device=dev).to(torch.bfloat16)
    src = ops.Src(z, cin)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    wp = ops.pack_conv3x3(w, 'conv')
    for _ in range(2):
        ops.conv3x3([src], wp, F_, H, H, cout)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        ops.conv3x3([src], wp, F_, H, H, cout)
    e1.record(); torch.cuda.synchronize()
    print(f'dbg={os.environ.get("SRVP_CONV_DBG", "0")} {nm:18s} {e0.elapsed_time(e1) / 5:7.3f} ms', flush=True)
