import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srvp_b200 import ops
dev = 'cuda'
frames, H, W, cin, cout = 2304, 32, 32, 128, 128
z = torch.randn(frames, H, W, cin, device=dev).to(torch.bfloat16)
srcs = [ops.Src(z, cin, torch.ones(cin, device=dev), torch.zeros(cin, device=dev), None, 0, 0, True)]
dz = torch.randn(frames, H, W, cout, device=dev).to(torch.bfloat16)
dw = torch.zeros(cout, cin, 3, 3, device=dev)
wp = ops.pack_conv3x3(dw, 'conv')
_, _, a_out = ops.conv3x3(srcs, wp, frames, H, W, cout, save_input=True)
for _ in range(3):
    ops.wgrad3x3(a_out, cin, dz, cout, frames, H, W, cout, cin, dw, 'conv')
torch.cuda.synchronize()

import ctypes
from srvp_b200 import _lib
def run(dbg, n=3):
    a = _lib.Wgrad3x3Args()
    a.act = _lib.ptr(a_out); a.act_channels, a.act_cpitch, a.act_coff = cin, cin, 0
    a.dz = _lib.ptr(dz); a.dz_channels, a.dz_cpitch, a.dz_coff = cout, cout, 0
    a.frames, a.H, a.W, a.cout, a.cin = frames, H, W, cout, cin
    a.dw = _lib.ptr(dw); a.stride_cout, a.stride_cin, a.flip = cin * 9, 9, dbg << 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.check(_lib.lib().srvp_wgrad3x3(ctypes.byref(a), _lib.stream_ptr()))
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        _lib.check(_lib.lib().srvp_wgrad3x3(ctypes.byref(a), _lib.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    print('dbg', dbg, e0.elapsed_time(e1) / n, 'ms', flush=True)
if len(sys.argv) > 1:
    for dbg in (0, 1, 2, 4, 5):
        run(dbg)
