"""Ad-hoc: forward-convolution extras on the two 64-channel 64x64 layers and a 256-channel one (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srvp_b200 import ops
dev = 'cuda'
F_, B = 2304, 192
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[n // 2]


tag = f'dbg={os.environ.get("SRVP_CONV_DBG", "0")}'
for (name, C, H, mode) in [('64@64 DIRECT', 64, 64, 0), ('64@64 UP2', 64, 64, 2), ('128@32 UP2', 128, 32, 2), ('256@16 DIRECT', 256, 16, 0), ('128@32 POOL2 (64->128)', 64, 32, 1)]:
    Hs = H // 2 if mode == 2 else (H * 2 if mode == 1 else H)
    cout = 128 if mode == 1 else C
    z = torch.randn(F_, Hs, Hs, C, device=dev).to(torch.bfloat16)
    sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    w = torch.randn(cout, C, 3, 3, device=dev) * 0.05
    wp = ops.pack_conv3x3(w, 'conv')
    add = torch.randn(cout // 4, B * H * H, 4, device=dev)
    plain = ops.Src(z, C, None, None, None, 0, mode, False)
    fused = ops.Src(z, C, sc, sh, None, 0, mode, True)
    res = {}
    res['plain'] = timeit(lambda: ops.conv3x3([plain], wp, F_, H, H, cout))
    res['+bn/lrelu'] = timeit(lambda: ops.conv3x3([fused], wp, F_, H, H, cout))
    res['+stats'] = timeit(lambda: ops.conv3x3([fused], wp, F_, H, H, cout, stats=True))
    res['+a_out'] = timeit(lambda: ops.conv3x3([fused], wp, F_, H, H, cout, stats=True, save_input=True))
    res['+add'] = timeit(lambda: ops.conv3x3([fused], wp, F_, H, H, cout, stats=True, save_input=True, add=add))
    fl = 2.0 * F_ * H * H * C * cout * 9
    print(tag, f'{name:24s}', '  '.join(f'{k} {v:.3f}' for k, v in res.items()), f' | plain {fl / res["plain"] / 1e9:.0f} TF/s', flush=True)
