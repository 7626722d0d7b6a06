"""Ad-hoc: bn_bwd on the big layer shapes (for ncu / event timing)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srvp_b200 import ops
from srvp_b200.ops import BNState
dev = 'cuda'
F_ = 2304
for (nm, H, C, mode) in [('d30 64ch@64 DIRECT', 64, 64, 0), ('e12 128ch@32 POOL2', 32, 128, 1), ('d21 64ch@32 UP2', 32, 64, 2), ('e22 256ch@16 DIRECT', 16, 256, 0)]:
    z = torch.randn(F_, H, H, C, device=dev).to(torch.bfloat16)
    Hd = H // 2 if mode == 1 else H * 2 if mode == 2 else H
    da = torch.randn(F_, Hd, Hd, C, device=dev).to(torch.bfloat16)
    st = BNState(C, dev)
    st.scale.fill_(1.0); st.shift.zero_(); st.mean.zero_(); st.invstd.fill_(1.0)
    gamma = torch.ones(C, device=dev); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    for _ in range(2):
        ops.bn_bwd(z, st, gamma, dg, db, da, mode, F_, H, H, C)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        ops.bn_bwd(z, st, gamma, dg, db, da, mode, F_, H, H, C)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    n = F_ * H * H * C * 2.0
    by = 3 * n + 2 * n * (0.25 if mode == 1 else 4 if mode == 2 else 1)
    print(f'{nm:24s} {ms:7.3f} ms  {by / ms / 1e6:7.1f} GB/s (both passes)', flush=True)
