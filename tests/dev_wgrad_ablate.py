"""Ad-hoc: wgrad3x3 with its development switches (1 = skip the global->shared copies, 2 = skip the MMAs) on the main layer shapes."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srvp_b200 import ops, _lib
dev = 'cuda'
F_ = 2304
for (nm, H, cin, cout) in [('e01 64->64 @64', 64, 64, 64), ('e12 128->128 @32', 32, 128, 128), ('d20h 128->128 @32', 32, 128, 128), ('e22 256->256 @16', 16, 256, 256),
                           ('e32 512->512 @8', 8, 512, 512), ('d30h 64->64 @64', 64, 64, 64)]:
    act = torch.randn(F_, H, H, cin, device=dev).to(torch.bfloat16)
    dz = torch.randn(F_, H, H, cout, device=dev).to(torch.bfloat16)
    dw = torch.zeros(cout, cin, 3, 3, device=dev)
    fl = 2.0 * F_ * H * H * cout * cin * 9
    res = []
    for dbg in (0, 1, 2):
        a = _lib.Wgrad3x3Args()
        a.act = _lib.ptr(act); a.act_channels, a.act_cpitch, a.act_coff = cin, cin, 0
        a.dz = _lib.ptr(dz); a.dz_channels, a.dz_cpitch, a.dz_coff = cout, cout, 0
        a.frames, a.H, a.W, a.cout, a.cin = F_, H, H, cout, cin
        a.dw = _lib.ptr(dw); a.stride_cout, a.stride_cin, a.flip = cin * 9, 9, dbg << 8
        for _ in range(2):
            _lib.check(_lib.lib().srvp_wgrad3x3(ctypes.byref(a), _lib.stream_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            _lib.check(_lib.lib().srvp_wgrad3x3(ctypes.byref(a), _lib.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 5)
    print(f'{nm:22s} full {res[0]:6.3f} ms ({fl / res[0] / 1e9:6.1f} TF/s) | no copies {res[1]:6.3f} ms ({fl / res[1] / 1e9:6.1f}) | no MMAs {res[2]:6.3f} ms', flush=True)
