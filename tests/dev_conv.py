"""Ad-hoc GPU check of the fused conv3x3 kernel against torch (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from srvp_b200 import ops, _lib

torch.manual_seed(0)
dev = 'cuda'


def bf(x):
    return x.to(torch.bfloat16).float()


def ref_src(z_nhwc, scale, shift, lrelu, mode, frame_map, H, W):
    a = z_nhwc.float()
    if scale is not None:
        a = a * scale + shift
    if lrelu:
        a = F.leaky_relu(a, 0.2)
    a = bf(a).permute(0, 3, 1, 2)
    if mode == _lib.SRC_POOL2:
        a = F.max_pool2d(a, 2)
    elif mode == _lib.SRC_UP2:
        a = F.interpolate(a, scale_factor=2, mode='nearest')
    if frame_map is not None:
        a = a[frame_map.long()]
    return a


def run_case(name, frames, H, W, cins, cout, modes, use_bn=True, kind='conv', fmap=False, sigmoid=False, stats=True):
    srcs, refs = [], []
    for cin, mode in zip(cins, modes):
        Hs, Ws = (H * 2, W * 2) if mode == _lib.SRC_POOL2 else (H // 2, W // 2) if mode == _lib.SRC_UP2 else (H, W)
        fm = None
        nf = frames
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
    if kind == 'conv':
        w = torch.randn(cout, cin_real, 3, 3, device=dev) * 0.05
        ref = F.conv2d(a[:, :cin_real], bf(w), padding=1)
    else:
        w = torch.randn(cin_real, cout, 3, 3, device=dev) * 0.05
        ref = F.conv_transpose2d(a[:, :cin_real], bf(w), padding=1)
    wp = ops.pack_conv3x3(w, kind)
    torch.cuda.synchronize()
    out, st = ops.conv3x3(srcs, wp, frames, H, W, cout, stats=stats and not sigmoid, sigmoid_nchw=sigmoid)
    torch.cuda.synchronize()
    if sigmoid:
        got = out
        ref = torch.sigmoid(ref)
    else:
        got = out.float().permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    msg = f'[{name}] max abs err {err:.4e} (rel to max {rel:.3e})'
    ok = rel < 1.5e-2
    if st is not None:
        s = st.double().sum(0)
        o = out.double()
        s_ref = torch.stack([o.sum((0, 1, 2)), (o * o).sum((0, 1, 2))], 1)
        serr = ((s - s_ref).abs() / (s_ref.abs() + 1)).max().item()
        msg += f' stats rel err {serr:.3e}'
        ok = ok and serr < 1e-3
    print(msg, 'PASS' if ok else 'FAIL', flush=True)
    return ok


def main():
    ok = True
    ok &= run_case('64->64 @64 direct', 3, 64, 64, [64], 64, [0])
    ok &= run_case('64->128 @32 pool', 5, 32, 32, [64], 128, [1])
    ok &= run_case('128->256 @16 pool', 7, 16, 16, [128], 256, [1])
    ok &= run_case('512->512 @8 direct', 9, 8, 8, [512], 512, [0])
    ok &= run_case('512+512->512 @8 up2+skip(fmap)', 6, 8, 8, [512, 512], 512, [2, 0], fmap=True)
    ok &= run_case('64+64->64 @64 up2+skip', 2, 64, 64, [64, 64], 64, [2, 0], fmap=True)
    ok &= run_case('16(3)->64 @64 thin no-bn', 3, 64, 64, [16], 64, [0], use_bn=False)
    ok &= run_case('64->3 convT sigmoid', 3, 64, 64, [64], 3, [0], kind='convT', sigmoid=True)
    ok &= run_case('dgrad-like 128->64 @32 no-bn', 4, 32, 32, [128], 64, [0], use_bn=False, stats=False)

    # timing of the big layers (BAIR shapes)
    global bench
    def bench(name, frames, H, W, cins, cout, modes, iters=5):
        srcs = []
        for cin, mode in zip(cins, modes):
            Hs, Ws = (H * 2, W * 2) if mode == 1 else (H // 2, W // 2) if mode == 2 else (H, W)
            z = torch.randn(frames, Hs, Ws, cin, device=dev).to(torch.bfloat16)
            srcs.append(ops.Src(z, cin, torch.ones(cin, device=dev), torch.zeros(cin, device=dev), None, 0, mode, True))
        w = torch.randn(cout, sum(cins), 3, 3, device=dev) * 0.05
        wp = ops.pack_conv3x3(w, 'conv')
        out = torch.empty(frames, H, W, cout, dtype=torch.bfloat16, device=dev)
        for _ in range(2):
            ops.conv3x3(srcs, wp, frames, H, W, cout, out=out, stats=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.conv3x3(srcs, wp, frames, H, W, cout, out=out, stats=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * frames * H * W * cout * sum(cins) * 9
        print(f'[bench {name}] {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s dense-equivalent', flush=True)


    if ok and len(sys.argv) > 1:
        N = 2304
        bench('enc.conv.0.1 64->64@64', N, 64, 64, [64], 64, [0])
        bench('enc.conv.1.2 128->128@32', N, 32, 32, [128], 128, [0])
        bench('enc.conv.2.2 256->256@16', N, 16, 16, [256], 256, [0])
        bench('enc.conv.3.2 512->512@8', N, 8, 8, [512], 512, [0])
        bench('dec.conv.0.0 1024->512@8', N, 8, 8, [512, 512], 512, [2, 0])
        bench('dec.conv.3.0 128->64@64', N, 64, 64, [64, 64], 64, [2, 0])
    print('ALL PASS' if ok else 'SOME FAILED')
    sys.exit(0 if ok else 1)

if __name__ == "__main__":
    main()
