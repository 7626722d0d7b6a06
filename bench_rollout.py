"""`bench.py --workload human_rollout`: the evaluation rollout of BASELINE.json configs[4] (reference test.py:235-254 with the README's
Human3.6M recipe): VGG64 skipco nc=3, ny=nz=50, nt_inf=3, 2 Euler steps, 8 conditioning frames -> 53 frames, 100 samples x 16 videos
per batch, PSNR + SSIM of every sample, best-of-100 selection.

One "step" = one batch of 16 videos: encoder on the conditioning frames, 100 x 16 latent rollouts (104 Euler steps each, 90 of them from
the prior), decoding of the 45 predicted frames of every sample (72 000 frames), metrics. value = predicted frames / s with the batch
resident in HBM; e2e = the same from pinned host uint8 videos with the best predictions copied back to the host. For context the same
batch is also run through the reference's own per-sample loop (encoder re-run per sample, one sample per launch) over the public API.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
CFG = dict(nx=64, nc=3, nf=64, nhx=128, ny=50, nz=50, skipco=True, nt_inf=3, nh_inf=256, nlayers_inf=3, nh_res=512, nlayers_res=4, archi='vgg')
ARG_ORDER = ['nx', 'nc', 'nf', 'nhx', 'ny', 'nz', 'skipco', 'nt_inf', 'nh_inf', 'nlayers_inf', 'nh_res', 'nlayers_res', 'archi']
NT_COND, NT_GEN, N_SAMPLES, BATCH, DT = 8, 53, 100, 16, 0.5


def run(args):
    import bench
    from srvp_b200 import _lib, ops, rollout
    from srvp_b200.module.srvp import StochasticLatentResidualVideoPredictor
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        sys.exit('human_rollout is a single-GPU workload (videos are independent: run one process per GPU for more)')
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(1)
    model = StochasticLatentResidualVideoPredictor(*[CFG[k] for k in ARG_ORDER])
    model.init(res_gain=1.2)
    model = model.to(dev).eval()
    model.noise_device = 'cuda'
    gen = torch.Generator().manual_seed(5)
    host = [torch.randint(0, 256, (BATCH, NT_GEN, 64, 64, 3), dtype=torch.uint8, generator=gen).pin_memory() for _ in range(2)]
    xdev = ops.u8_to_tbchw_f32(host[0].to(dev))
    sb = args.sample_batch if getattr(args, 'sample_batch', None) else 25

    def step(x):
        return rollout.best_of_n(model, x, NT_COND, N_SAMPLES, DT, sample_batch=sb)

    sampler = bench.ClockSampler(0)
    sampler.start()
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step(xdev)
        torch.cuda.synchronize()
        sampler.recording = True
        l0 = _lib.lib().srvp_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            r = step(xdev)
        e1.record()
        torch.cuda.synchronize()
        launches = _lib.lib().srvp_launch_count() - l0
        ms = e0.elapsed_time(e1)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        d2h = 0
        for i in range(args.steps):
            xb = ops.u8_to_tbchw_f32(host[i % 2].to(dev, non_blocking=True))
            r = step(xb)
            best = (r['x_best'] * 255).byte().cpu()
            scores = r['psnr'].max(0)[0].cpu()
            d2h = best.numel() + scores.numel() * 4
        t1.record()
        torch.cuda.synchronize()
        ms_e2e = t0.elapsed_time(t1)
        sampler.stop_flag = True
        sampler.join(timeout=2)
        # the reference's loop over the public API (encode once, then forward / generate / decode per sample): 10 samples, scaled
        import test as test_py
        n_ref = 10
        test_py.reference_loop(model, xdev, NT_COND, 2, DT, DT)
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        test_py.reference_loop(model, xdev, NT_COND, n_ref, DT, DT)
        q1.record()
        torch.cuda.synchronize()
        ms_loop = q0.elapsed_time(q1) * (N_SAMPLES / n_ref)
        # latent loop alone: Euler steps per second of the persistent kernel at this shape (100 x 16 = 1600 trajectories per launch)
        ops.PROFILE = {}
        step(xdev)
        torch.cuda.synchronize()
        prof = ops.summarize_profile(ops.PROFILE)
        ops.PROFILE = None
    frames = N_SAMPLES * BATCH * (NT_GEN - NT_COND)
    pk = bench.peaks()
    brk = {k: dict(ms_per_step=round(v['ms'], 3), launches=v['launches'],
                   tflops=round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 1) if v['flops'] else None) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])
           if not k.startswith('hbm:')}
    conv = prof.get('conv3x3', dict(ms=1e-9, flops=0, launches=1))
    lat = prof.get('latent_fwd', dict(ms=1e-9, launches=1))
    line = dict(metric='predicted frames/sec, evaluation rollout (Human3.6M recipe: 8 -> 53 frames, 2 Euler steps, 100 samples x 16 videos)',
                value=round(frames * args.steps / (ms * 1e-3), 1), unit='frames/s', n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=round(ms / args.steps, 2), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16', data='synthetic',
                config=dict(workload='Human3.6M VGG64 skipco nc=3 ny=nz=50 nt_inf=3 n_euler_steps=2: nt_cond 8 -> nt_gen 53, 100 samples x 16 videos, '
                                     'PSNR+SSIM of every sample, best-of-100 (reference test.py:235-254)', sample_batch=sb,
                            l2='72 000 decoded frames per step: > 100 GB of activations streamed'),
                e2e=dict(value=round(frames * args.steps / (ms_e2e * 1e-3), 1), unit='frames/s', h2d_bytes_per_step=host[0].numel(), d2h_bytes_per_step=d2h),
                gpu_launches=int(launches), clocks=sampler.summary(),
                roofline=dict(bound='tensor', kernel='conv3x3', achieved=round(conv['flops'] / (conv['ms'] * 1e-3) / 1e12, 1), peak=pk['tf'], unit='TFLOP/s',
                              frac=round(conv['flops'] / (conv['ms'] * 1e-3) / 1e12 / pk['tf'], 4), traffic=None,
                              share_of_kernel_time=round(conv['ms'] / sum(v['ms'] for k, v in prof.items() if not k.startswith('hbm:')), 3)),
                latent_loop=dict(ms_per_launch=round(lat['ms'] / lat['launches'], 3), launches_per_step=lat['launches'],
                                 trajectories_per_launch=sb * BATCH, euler_steps=2 * (NT_GEN - 1),
                                 note='one persistent launch per chunk of samples replaces ~2 200 ATen launches and ~104 host synchronisations of the '
                                      'reference loop per sample (SURVEY.md 3.3)'),
                reference_style_loop=dict(ms_per_step=round(ms_loop, 1), value=round(frames / (ms_loop * 1e-3), 1), unit='frames/s',
                                          note=f'same model, the reference per-sample loop through forward / generate / decode ({n_ref} samples timed, '
                                               f'scaled to {N_SAMPLES})'),
                kernel_breakdown=brk)
    print(json.dumps(line), flush=True)
