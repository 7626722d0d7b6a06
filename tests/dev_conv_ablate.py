"""Ad-hoc: which part of the conv3x3 kernel bounds the 64-channel / pooled layers? Times ablated launches (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srvp_b200 import ops
dev = 'cuda'
F_ = 2304


def run(name, H, cin, cout, mode=0, bn=True, stats=True, save=True, n=5):
    Hs = H * 2 if mode == 1 else H // 2 if mode == 2 else H
    z = torch.randn(F_, Hs, Hs, cin, device=dev).to(torch.bfloat16)
    sc = torch.ones(cin, device=dev) if bn else None
    sh = torch.zeros(cin, device=dev) if bn else None
    src = ops.Src(z, cin, sc, sh, None, 0, mode, bn)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    wp = ops.pack_conv3x3(w, 'conv')
    for _ in range(2):
        ops.conv3x3([src], wp, F_, H, H, cout, stats=stats, save_input=save)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        ops.conv3x3([src], wp, F_, H, H, cout, stats=stats, save_input=save)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * F_ * H * H * cout * cin * 9
    print(f'{name:34s} {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s', flush=True)


for (nm, H, cin, cout, mode) in [('e01 64->64 @64', 64, 64, 64, 0), ('d21 128->64 @32', 32, 128, 64, 0), ('e11 64->128 @32 POOL2', 32, 64, 128, 1),
                                 ('e12 128->128 @32', 32, 128, 128, 0), ('e21 128->256 @16 POOL2', 16, 128, 256, 1), ('d20h 128->128 @32 UP2', 32, 128, 128, 2)]:
    run(nm + ' full', H, cin, cout, mode)
    run(nm + ' -a_out', H, cin, cout, mode, save=False)
    run(nm + ' -a_out -stats', H, cin, cout, mode, save=False, stats=False)
    run(nm + ' plain (dgrad-like)', H, cin, cout, mode, bn=False, stats=False, save=False)
