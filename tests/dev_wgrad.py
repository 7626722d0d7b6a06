"""Ad-hoc GPU check of the wgrad3x3 kernel against torch autograd (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from srvp_b200 import ops, _lib
from dev_conv import ref_src, bf

dev = 'cuda'
torch.manual_seed(1)


def run_case(name, frames, H, W, cins, cout, modes, kind='conv', fmap=False, use_bn=True, dz_pad=None):
    srcs, refs = [], []
    for cin, mode in zip(cins, modes):
        Hs, Ws = (H * 2, W * 2) if mode == 1 else (H // 2, W // 2) if mode == 2 else (H, W)
        fm, nf = None, frames
        if fmap and len(srcs) == 1:
            nf = max(1, frames // 2)
            fm = torch.randint(nf, (frames,), device=dev, dtype=torch.int32)
        z = torch.randn(nf, Hs, Ws, cin, device=dev).to(torch.bfloat16)
        sc = (torch.rand(cin, device=dev) + 0.5) if use_bn else None
        sh = (torch.randn(cin, device=dev) * 0.3) if use_bn else None
        srcs.append(ops.Src(z, cin, sc, sh, fm, 0, mode, use_bn))
        refs.append(ref_src(z, sc, sh, use_bn, mode, fm, H, W))
    a = torch.cat(refs, 1)
    cin_tot = sum(cins)
    cin_real = 3 if cin_tot == 16 else cin_tot
    cpad = dz_pad or cout
    dz = torch.zeros(frames, H, W, cpad, device=dev, dtype=torch.bfloat16)
    dz[..., :cout] = (torch.randn(frames, H, W, cout, device=dev) * 0.1).to(torch.bfloat16)
    dzr = dz[..., :cout].float().permute(0, 3, 1, 2)
    if kind == 'conv':
        w = torch.zeros(cout, cin_real, 3, 3, device=dev, requires_grad=True)
        F.conv2d(a[:, :cin_real], w, padding=1).backward(dzr)
    else:
        w = torch.zeros(cin_real, cout, 3, 3, device=dev, requires_grad=True)
        F.conv_transpose2d(a[:, :cin_real], w, padding=1).backward(dzr)
    ref = w.grad
    dw = torch.zeros_like(ref)
    # the forward conv writes the materialised input a_out as a side effect
    wp = ops.pack_conv3x3(w.detach(), kind)
    _, _, a_out = ops.conv3x3(srcs, wp, frames, H, W, cout, save_input=True, sigmoid_nchw=(kind == 'convT'))
    a_ref = a.permute(0, 2, 3, 1)
    aerr = (a_out.float() - a_ref).abs().max().item()
    ops.wgrad3x3(a_out, cin_tot, dz, cpad, frames, H, W, cout, cin_real, dw, kind)
    torch.cuda.synchronize()
    err = (dw - ref).abs().max().item() / ref.abs().max().item()
    ok = err < 5e-3
    ok = ok and aerr < 2e-2  # a_out is bf16: 1-ulp differences vs the fp32-then-round reference
    print(f'[{name}] rel-to-max err {err:.3e} a_out err {aerr:.1e}', 'PASS' if ok else 'FAIL', flush=True)
    return ok


ok = True
ok &= run_case('64->64 @64', 3, 64, 64, [64], 64, [0])
ok &= run_case('64->128 @32 pool', 5, 32, 32, [64], 128, [1])
ok &= run_case('256->256 @16', 3, 16, 16, [256], 256, [0])
ok &= run_case('512+512->512 @8 up2+skip', 6, 8, 8, [512, 512], 512, [2, 0], fmap=True)
ok &= run_case('64+64->64 @64 up2+skip (halo on M)', 2, 64, 64, [64, 64], 64, [2, 0], fmap=True)
ok &= run_case('128->64 @32 (halo on M)', 3, 32, 32, [128], 64, [0])
ok &= run_case('16(3)->64 @64 thin', 3, 64, 64, [16], 64, [0], use_bn=False)
ok &= run_case('64->3 convT', 3, 64, 64, [64], 3, [0], kind='convT', dz_pad=16)


def bench(name, frames, H, W, cins, cout, modes, iters=3):
    srcs = []
    for cin, mode in zip(cins, modes):
        Hs, Ws = (H * 2, W * 2) if mode == 1 else (H // 2, W // 2) if mode == 2 else (H, W)
        z = torch.randn(frames, Hs, Ws, cin, device=dev).to(torch.bfloat16)
        srcs.append(ops.Src(z, cin, torch.ones(cin, device=dev), torch.zeros(cin, device=dev), None, 0, mode, True))
    dz = torch.randn(frames, H, W, cout, device=dev).to(torch.bfloat16)
    dw = torch.zeros(cout, sum(cins), 3, 3, device=dev)
    wp = ops.pack_conv3x3(dw, 'conv')
    _, _, a_out = ops.conv3x3(srcs, wp, frames, H, W, cout, save_input=True)
    srcs = a_out
    for _ in range(1):
        ops.wgrad3x3(srcs, sum(cins), dz, cout, frames, H, W, cout, sum(cins), dw, 'conv')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.wgrad3x3(srcs, sum(cins), dz, cout, frames, H, W, cout, sum(cins), dw, 'conv')
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * frames * H * W * cout * sum(cins) * 9
    print(f'[bench wgrad {name}] {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s dense-equivalent', flush=True)


if ok and len(sys.argv) > 1:
    N = 2304
    bench('enc.conv.0.1 64->64@64', N, 64, 64, [64], 64, [0])
    bench('enc.conv.1.2 128->128@32', N, 32, 32, [128], 128, [0])
    bench('enc.conv.2.2 256->256@16', N, 16, 16, [256], 256, [0])
    bench('enc.conv.3.2 512->512@8', N, 8, 8, [512], 512, [0])
    bench('dec.conv.0.0 1024->512@8', N, 8, 8, [512, 512], 512, [2, 0])
    bench('dec.conv.3.0 128->64@64', N, 64, 64, [64, 64], 64, [2, 0])
print('ALL PASS' if ok else 'SOME FAILED')
