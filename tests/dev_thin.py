"""Ad-hoc: the HBM-bound ends of the network (first encoder conv, decoder head, their gradients) and the pooled batch-norm backward,
under the development switches SRVP_CONV_DBG (1 = no activation copies, 2 = no MMAs, 4 = no output stores)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srvp_b200 import ops, _lib
dev = 'cuda'
F_ = 2304
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
x16 = torch.rand(F_, 64, 64, 16, device=dev).to(torch.bfloat16)
w0 = torch.randn(64, 3, 3, 3, device=dev) * 0.05
wp0 = ops.pack_conv3x3(w0, 'conv')
print(tag, 'first conv 16->64 stats      ', round(timeit(lambda: ops.conv3x3([ops.Src(x16, 16)], wp0, F_, 64, 64, 64, stats=True, cin_real=3)), 3), 'ms', flush=True)
z = torch.randn(F_, 64, 64, 64, device=dev).to(torch.bfloat16)
sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
wf = torch.randn(64, 3, 3, 3, device=dev) * 0.05
wpf = ops.pack_conv3x3(wf, 'convT')
src = ops.Src(z, 64, sc, sh, None, 0, 0, True)
print(tag, 'head 64->3 sigmoid + a_out   ', round(timeit(lambda: ops.conv3x3([src], wpf, F_, 64, 64, 3, sigmoid_nchw=True, save_input=True)), 3), 'ms', flush=True)
print(tag, 'head 64->3 sigmoid, no a_out ', round(timeit(lambda: ops.conv3x3([src], wpf, F_, 64, 64, 3, sigmoid_nchw=True, save_input=False)), 3), 'ms', flush=True)
print(tag, 'decoder head kernel, no a_out', round(timeit(lambda: ops.decoder_head_fwd(src, wf, F_, 3, save_input=False)), 3), 'ms', flush=True)
dz16 = torch.randn(F_, 64, 64, 16, device=dev).to(torch.bfloat16)
wpd = ops.pack_conv3x3(wf, 'convT_dgrad')
print(tag, 'head dgrad 16->64            ', round(timeit(lambda: ops.conv3x3([ops.Src(dz16, 16)], wpd, F_, 64, 64, 64, cin_real=3)), 3), 'ms', flush=True)
dwf = torch.zeros_like(wf)
print(tag, 'head wgrad act64 x dz16      ', round(timeit(lambda: ops.wgrad3x3(z, 64, dz16, 16, F_, 64, 64, 3, 64, dwf, 'convT')), 3), 'ms', flush=True)
print(tag, 'head wgrad raw z + affine    ', round(timeit(lambda: ops.wgrad3x3(z, 64, dz16, 16, F_, 64, 64, 3, 64, dwf, 'convT', act_affine=(sc, sh, True))), 3), 'ms', flush=True)
dz64 = torch.randn(F_, 64, 64, 64, device=dev).to(torch.bfloat16)
dw0 = torch.zeros_like(w0)
print(tag, 'first wgrad act16 x dz64     ', round(timeit(lambda: ops.wgrad3x3(x16, 16, dz64, 64, F_, 64, 64, 64, 3, dw0, 'conv')), 3), 'ms', flush=True)
if os.environ.get('SRVP_CONV_DBG', '0') == '0':
    class BN: pass
    for (H, C) in [(64, 64), (32, 128), (16, 256), (8, 512)]:
        zz = torch.randn(F_, H, H, C, device=dev).to(torch.bfloat16)
        bn = BN(); bn.weight = torch.rand(C, device=dev) + 0.5; bn.bias = torch.zeros(C, device=dev)
        st = ops.BNState(C, dev)
        ops.bn_finalize(ops.channel_stats(zz.view(-1, C)), float(F_ * H * H), bn, st, training_update=False)
        da = torch.randn(F_, H // 2, H // 2, C, device=dev).to(torch.bfloat16)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        B = 192
        skip = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16)
        inv = torch.full((F_,), -1, dtype=torch.int32, device=dev)
        inv[torch.randperm(F_)[:B].to(dev)] = torch.arange(B, dtype=torch.int32, device=dev)
        t1 = timeit(lambda: ops.bn_bwd(zz, st, bn.weight, dg, db, da, 1, F_, H, H, C))
        t2 = timeit(lambda: ops.bn_bwd(zz, st, bn.weight, dg, db, da, 1, F_, H, H, C, skip=skip, skip_coff=0, nt=1, B=B, inv_map=inv))
        n = F_ * H * H * C * 2
        print(f'bn_bwd pooled {H}x{H} C={C}: {t1:.3f} ms, with skip {t2:.3f} ms   (z {n / 1e6:.0f} MB; ideal traffic 3.5x = {3.5 * n / 6.5e9:.3f} ms)', flush=True)
