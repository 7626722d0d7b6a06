"""Benchmark of the SRVP hot path: frames/s of one full training step (forward + ELBO + backward + Adam).

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
  workload  BAIR config of BASELINE.json: VGG64 skipco nc=3 64x64, seq_len 12, batch 192 per GPU (weak scaling),
            ny=nz=50, nt_inf=2, 2 Euler steps, obs_scale 0.71 (reference README.md:127); synthetic uniform frames.
  value     frames/s with the batch already resident in HBM (device timed, CUDA events, max over ranks)
  e2e       same step through the public API from pinned HOST frames: H2D of the batch and D2H of the loss inside the
            timed region
  roofline  dominant kernel family: algorithmic dense FLOPs / CUDA-event time vs the measured bf16 peak
  conv_stages / hbm_kernels: encoder / decoder convolutions vs the tensor roofline, the HBM-bound ends of the network (first encoder
            convolution, decoder head conv+sigmoid, their gradients) vs the measured HBM peak
  gpu_eager_baseline (N=1): the reference algorithm (oracle port = the reference's own torch ops) in eager PyTorch on the SAME GPU,
            fp32 with cuDNN's default TF32 convolutions and under bf16 autocast -- the incumbent cuDNN/cuBLAS sm_100 kernels
  strong_scaling (N>1): the same step with the metric's own global batch (192) split over the ranks (reference train.py:218-219)
  cpu_baseline / --impl reference: the reference algorithm (oracle port, torch CPU fp32) on ALL host cores, bounded sample
  --workload human_rollout: the evaluation rollout of configs[4] (test.py: 16 videos x 100 samples, 8 -> 53 frames, 2 Euler steps)
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(nx=64, nc=3, nf=64, nhx=128, ny=50, nz=50, skipco=True, nt_inf=2, nh_inf=256, nlayers_inf=3, nh_res=512, nlayers_res=4,
           archi='vgg')
ARG_ORDER = ['nx', 'nc', 'nf', 'nhx', 'ny', 'nz', 'skipco', 'nt_inf', 'nh_inf', 'nlayers_inf', 'nh_res', 'nlayers_res', 'archi']
LOSS = dict(obs_scale=0.71, beta_y=1.0, beta_z=1.0, l2_res=1.0)
SEQ_LEN, BATCH, DT = 12, 192, 0.5
METRIC = 'frames/sec training step (BAIR 64x64 seq12 bs192)'
WORKLOAD = 'BAIR VGG64 skipco nc=3 64x64 seq_len=12 batch=192/GPU ny=nz=50 nt_inf=2 n_euler_steps=2; fwd+ELBO+bwd+Adam'


def select_workload(name):
    """'bair' (default, the configuration BASELINE.json's metric is quoted on) or 'smmnist' (configs[1]: DCGAN64, README.md:111), for the record."""
    global CFG, LOSS, SEQ_LEN, BATCH, DT, METRIC, WORKLOAD
    if name == 'smmnist':
        CFG = dict(nx=64, nc=1, nf=64, nhx=128, ny=20, nz=20, skipco=False, nt_inf=5, nh_inf=256, nlayers_inf=3, nh_res=512, nlayers_res=4,
                   archi='dcgan')
        LOSS = dict(obs_scale=1.0, beta_y=1.0, beta_z=2.0, l2_res=1.0)
        SEQ_LEN, BATCH, DT = 15, 128, 1.0
        METRIC = 'frames/sec training step (smmnist 64x64 seq15 bs128)'
        WORKLOAD = 'smmnist DCGAN64 nc=1 64x64 seq_len=15 batch=128/GPU ny=nz=20 nt_inf=5 n_euler_steps=1; fwd+ELBO+bwd+Adam'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tf=d['bf16_tflops_sustained'], tf_burst=d['bf16_tflops'], hbm=d['hbm_gbs'], src='measured')
    return dict(tf=1400.0, tf_burst=1590.0, hbm=6650.0, src='fallback')


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU through NVML (in-process, initialised before the timed region)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.stop_flag, self.max_mhz, self.reasons = [], False, None, set()
        self.recording = False
        self.nv, self.h = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            phys = index
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            if vis:
                phys = int(vis.split(',')[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        masks = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown, 'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown, 'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                clk = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self.recording:
                    self.samples.append(clk)
                    for name, m in masks.items():
                        if r & m:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return dict(sm_mhz=s[len(s) // 2] if s else None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(s))


def elbo_loss(out, x):
    """train.py:90-106 through the fused ELBO reductions (srvp_b200/elbo.py)."""
    from srvp_b200 import elbo
    return elbo.elbo(out, x, LOSS['obs_scale'], LOSS['beta_y'], LOSS['beta_z'], LOSS['l2_res'])[0]


def make_model(dev, world):
    from srvp_b200.module.srvp import StochasticLatentResidualVideoPredictor
    torch.manual_seed(1)
    model = StochasticLatentResidualVideoPredictor(*[CFG[k] for k in ARG_ORDER])
    model.init(res_gain=1.41)
    if world > 1:
        # reference multi-GPU semantics (train.py:283): batch-norm statistics over the GLOBAL batch
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model = model.to(dev).train()
    model.noise_device = 'cuda'
    return model


def run_ours(args):
    import torch.distributed as dist
    from srvp_b200 import ops, _lib, parallel
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    model = make_model(dev, world)
    params = [p for p in model.parameters()]
    from srvp_b200.optim import Adam
    opt = Adam(params, lr=3e-4)      # torch.optim.Adam semantics (train.py:289), all tensors in one launch
    # gradients live in ONE flat buffer: one memset + (N > 1) one in-place all-reduce per step, the decoder's part launched early
    bucket = parallel.GradBucket(params, early=list(model.decoder.parameters()))
    parallel.ACTIVE_BUCKET = bucket
    gen = torch.Generator().manual_seed(123 + rank)
    strong = args.scaling == 'strong'
    batch = BATCH // world if strong else BATCH
    assert batch * world == BATCH or not strong, 'strong scaling needs the global batch to be divisible by the number of GPUs (train.py:218)'

    def host_batches(b):
        # several distinct host batches (pinned) so that the e2e loop really moves data: uint8 (B, T, H, W, C) frames as the datasets
        # store them; the device-resident batch of the `value` loop is their fp32 (T, B, C, H, W) conversion
        return [torch.randint(0, 256, (b, SEQ_LEN, 64, 64, CFG['nc']), dtype=torch.uint8, generator=gen).pin_memory() for _ in range(2)]

    host = host_batches(batch)
    xdev = ops.u8_to_tbchw_f32(host[0].to(dev))

    def step(x):
        bucket.zero()
        out = model(x, SEQ_LEN, dt=DT)
        loss = elbo_loss(out, x)
        loss.backward()
        bucket.allreduce_mean()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(x, steps, hosts=None):
        """(ms, last loss value): `steps` training steps, device-timed; hosts: e2e mode (H2D of the uint8 batch + D2H of the loss per step)."""
        barrier()
        # the host enqueues ~300 launches per step ahead of the device: a generation-2 pass of Python's cyclic garbage collector in the
        # middle of the timed loop (50 - 300 ms with the whole torch object graph alive) starves it. Collect now, not while timing.
        gc.collect()
        gc.disable()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lv = None
        e0.record()
        for i in range(steps):
            if hosts is None:
                loss = step(x)
            else:
                xb = ops.u8_to_tbchw_f32(hosts[i % len(hosts)].to(dev, non_blocking=True))
                lv = step(xb).item()
        e1.record()
        barrier()
        gc.enable()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), (lv if lv is not None else float(loss))

    sampler = ClockSampler(local)   # NVML is initialised and polled from before the warm-up so that it cannot disturb the timed region
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(xdev)
    # the caching allocator must have reached its steady state before anything is timed: a step that still grows the pool calls
    # cudaMalloc (a device synchronisation) in the middle of the pipeline. Extra untimed steps until two consecutive ones leave the
    # reserved size and the cudaMalloc count unchanged (at most 8).
    def pool_state():
        st = torch.cuda.memory_stats(dev)
        return (st.get('reserved_bytes.all.current', 0), st.get('num_device_alloc', st.get('segment.all.allocated', 0)), st.get('num_alloc_retries', 0))
    extra_warmup, stable, prev = 0, 0, pool_state()
    while stable < 2 and extra_warmup < 8:
        step(xdev)
        torch.cuda.synchronize()
        cur = pool_state()
        stable = stable + 1 if cur == prev else 0
        prev = cur
        extra_warmup += 1
    barrier()
    sampler.recording = True
    launches0 = _lib.lib().srvp_launch_count()
    side0 = ops.SIDE_LAUNCHES[0]
    ms, _ = timed(xdev, args.steps)
    launches = _lib.lib().srvp_launch_count() - launches0
    ms_e2e, lv = timed(xdev, args.steps, hosts=host)
    side_per_step = (ops.SIDE_LAUNCHES[0] - side0) // max(1, args.steps * 2)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # the other scaling mode, for the record (N > 1 only): weak run -> also the metric's own global batch split over the ranks
    other = None
    if world > 1 and BATCH % world == 0:
        ob = BATCH if strong else BATCH // world
        ohost = host_batches(ob)
        ox = ops.u8_to_tbchw_f32(ohost[0].to(dev))
        for _ in range(3):
            step(ox)
        osteps = max(3, min(args.steps, 10))
        oms, _ = timed(ox, osteps)
        oms_e2e, _ = timed(ox, osteps, hosts=ohost)
        other = dict(scaling='weak' if strong else 'strong', global_batch=ob * world, per_gpu_batch=ob, steps=osteps,
                     ms_per_step=round(oms / osteps, 3), value=round(SEQ_LEN * ob * world * osteps / (oms * 1e-3), 1),
                     e2e_value=round(SEQ_LEN * ob * world * osteps / (oms_e2e * 1e-3), 1), unit='frames/s')
        del ox, ohost

    # per-kernel profile (2 extra steps with CUDA events around every launch)
    ops.PROFILE = {}
    for _ in range(2):
        step(xdev)
    torch.cuda.synchronize()
    prof = ops.summarize_profile(ops.PROFILE)
    ops.PROFILE = None
    pk = peaks()
    fam = {k: v for k, v in prof.items() if not k.startswith('hbm:')}
    tot_ms = sum(v['ms'] for v in fam.values())
    dom = max((k for k in fam if fam[k]['flops'] > 0), key=lambda k: fam[k]['ms'])
    d = fam[dom]
    achieved = d['flops'] / (d['ms'] * 1e-3) / 1e12
    roofline = dict(bound='tensor', kernel=dom, achieved=round(achieved, 1), peak=pk['tf'], unit='TFLOP/s', frac=round(achieved / pk['tf'], 4),
                    traffic=None, peak_source=pk['src'] + ' (sustained bf16 cuBLAS)', share_of_kernel_time=round(d['ms'] / tot_ms, 3),
                    launches_per_step=d['launches'] // 2, avg_launch_ms=round(d['ms'] / d['launches'], 4))
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')   # DRAM bytes per launch of each kernel family from the committed ncu capture
    if os.path.exists(tpath) and WORKLOAD.startswith('BAIR'):
        tj = json.load(open(tpath))
        if dom in tj:
            roofline['traffic'] = tj[dom]['dram_bytes_per_launch']
            roofline['traffic_source'] = tj[dom]['source']
    roofline['executed'] = round(d['executed_flops'] / (d['ms'] * 1e-3) / 1e12, 1)
    roofline['note'] = ('achieved = dense FLOPs of the reference ops these launches stand for (SURVEY.md 8d) / time; executed = multiply-adds '
                        'actually issued (the convolutions over cat[h, skip] are split per video, DESIGN.md section 4)')
    # north star: the encoder convolutions against the tensor-core roofline, the decoder against both (its convolutions are tensor
    # bound); forward / data-gradient launches (conv3x3) and weight-gradient launches (wgrad3x3) of each stage
    stages = {}
    for famname in ('conv3x3', 'wgrad3x3'):
        for t, v in prof.get(famname, {}).get('by_stage', {}).items():
            tf, gb = v['flops'] / (v['ms'] * 1e-3) / 1e12, v['bytes'] / (v['ms'] * 1e-3) / 1e9
            stages[f'{famname}:{t}'] = dict(ms_per_step=round(v['ms'] / 2, 3), launches=v['launches'] // 2, tflops=round(tf, 1),
                                            frac_tensor=round(tf / pk['tf'], 4), gbs=round(gb, 1), frac_hbm=round(gb / pk['hbm'], 4))
    # the HBM-bound ends of the network, each against the measured HBM peak: algorithmic bytes (real input channels read once + output
    # written once) / CUDA-event time
    hbm = {}
    for k, v in prof.items():
        if k.startswith('hbm:') and v['ms'] > 0:
            gb = v['bytes'] / (v['ms'] * 1e-3) / 1e9
            hbm[k[4:]] = dict(ms_per_launch=round(v['ms'] / v['launches'], 4), launches=v['launches'] // 2, algorithmic_mb=round(v['bytes'] / v['launches'] / 1e6, 1),
                              gbs=round(gb, 1), frac_hbm=round(gb / pk['hbm'], 4))
    breakdown = {k: dict(ms_per_step=round(v['ms'] / 2, 3), launches=v['launches'] // 2,
                         tflops=round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 1) if v['flops'] else None,
                         executed_tflops=round(v['executed_flops'] / (v['ms'] * 1e-3) / 1e12, 1) if v['executed_flops'] else None,
                         gbs=round(v['bytes'] / (v['ms'] * 1e-3) / 1e9, 1) if v['bytes'] else None) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]['ms'])}
    eager = None
    if world == 1 and not args.no_eager_baseline:
        del xdev
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(dev, steps=max(2, min(args.steps, 5)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    frames = SEQ_LEN * batch * world
    line = dict(metric=METRIC, value=round(frames * args.steps / (ms * 1e-3), 1), unit='frames/s',
                n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=round(ms / args.steps, 3), higher_is_better=True,
                scaling='strong' if strong else 'weak', vs_baseline=None, dtype='bf16', data='synthetic',
                config=dict(workload=WORKLOAD if not strong else WORKLOAD.replace(f'batch={BATCH}/GPU', f'global batch={BATCH}'),
                            global_batch=batch * world, per_gpu_batch=batch, seq_len=SEQ_LEN,
                            parallelism=f'dp{world}' + (' + SyncBN statistics, one flat gradient all-reduce' if world > 1 else ''),
                            l2='inputs larger than L2: 113 MB batch, >10 GB of activations per step',
                            extra_warmup_steps=extra_warmup,   # untimed steps after the W warm-up steps until the caching allocator stopped growing
                            e2e_input='uint8 frames (B,T,H,W,C) from pinned host memory, converted on the device',
                            parity='ELBO within 1e-4 rel of the fp32 reference at this size (tests/test_gpu_parity_full.py); KL(z) term within 1e-2'),
                e2e=dict(value=round(frames * args.steps / (ms_e2e * 1e-3), 1), unit='frames/s', h2d_bytes_per_step=host[0].numel(),
                         d2h_bytes_per_step=4),
                gpu_launches=int(launches), clocks=sampler.summary(), roofline=roofline, conv_stages=stages, hbm_kernels=hbm,
                kernel_breakdown=breakdown,
                streams=dict(weight_gradient_stream=bool(ops.WGRAD_STREAM), side_launches_per_step=side_per_step,
                             sum_kernel_ms_serialised=round(tot_ms / 2, 3),
                             note='timed steps: weight gradients (and, under the GradBucket, the head / latent weight-gradient GEMMs) run on a second '
                                  'stream next to the HBM-bound batch-norm backward and the few-CTA latent / inference backward kernels '
                                  '(srvp_b200/ops.py); roofline / conv_stages / hbm_kernels / kernel_breakdown come from 2 extra steps with '
                                  'everything on one stream and CUDA events around every launch, so their sum exceeds ms_per_step'),
                loss=lv)
    if other is not None:
        line[other['scaling'] + '_scaling'] = other
    if eager is not None:
        line['gpu_eager_baseline'] = eager
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(steps=1, warmup=1)
    print(json.dumps(line), flush=True)


def _oracle_trainer(device, batch):
    """The reference's training step (train.py:49-129) restated over the oracle port: returns (step_fn, x)."""
    from oracle import srvp_oracle as O
    from srvp_b200.module.srvp import StochasticLatentResidualVideoPredictor
    torch.manual_seed(1)
    m = StochasticLatentResidualVideoPredictor(*[CFG[k] for k in ARG_ORDER])
    m.init(res_gain=1.41)
    sd = {k: v.to(device).clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in m.state_dict().items()}
    ps = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(ps, lr=3e-4)
    x = torch.rand(SEQ_LEN, batch, CFG['nc'], 64, 64, generator=torch.Generator().manual_seed(123)).to(device)

    def step(autocast=None):
        opt.zero_grad()
        rnd = O.draw_randoms(CFG, SEQ_LEN, SEQ_LEN, batch, training=True)
        if device != 'cpu':
            rnd = {k: ([e.to(device, non_blocking=True) for e in v] if isinstance(v, list) else v.to(device, non_blocking=True)) for k, v in rnd.items()}
        if autocast is not None:
            with torch.autocast('cuda', dtype=autocast):
                o = O.forward(sd, CFG, x, SEQ_LEN, DT, rnd, training=True)
            o = {k: (v.float() if torch.is_tensor(v) else v) for k, v in o.items()}
        else:
            o = O.forward(sd, CFG, x, SEQ_LEN, DT, rnd, training=True)
        loss = O.elbo(o, x, LOSS)[0]
        loss.backward()
        opt.step()
        return loss.detach()

    return step, x


def gpu_eager_baseline(dev, steps, warmup=2):
    """The incumbent: the reference algorithm in eager PyTorch on the SAME GPU -- cuDNN / cuBLAS sm_100
    kernels, cudnn.benchmark on as reference train.py:323 -- at the full configuration, device-timed with CUDA events.
    fp32 = PyTorch defaults (TF32 cuDNN convolutions, fp32 matmuls); bf16 = torch.autocast(bfloat16)."""
    out = dict(kind='port', what='oracle port of reference train.py:49-129 (same torch ops as the reference modules), eager PyTorch '
                                 f'{torch.__version__}, cudnn.benchmark=True', batch=BATCH, steps=steps, warmup=warmup, unit='frames/s')
    bench_flag = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        for name, ac in (('fp32_tf32', None), ('bf16_autocast', torch.bfloat16)):
            try:
                step, x = _oracle_trainer(dev, BATCH)
                for _ in range(warmup):
                    step(ac)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    loss = step(ac)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                out[name] = dict(ms_per_step=round(ms, 2), value=round(SEQ_LEN * BATCH / (ms * 1e-3), 1), loss=float(loss),
                                 peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 1))
            except torch.cuda.OutOfMemoryError as e:
                out[name] = dict(error='out of memory at the full batch: ' + str(e)[:120])
            step = x = None
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
    finally:
        torch.backends.cudnn.benchmark = bench_flag
    return out


def cpu_baseline(steps, warmup, batch=8):
    """The reference algorithm (oracle port of module/srvp.py + train.py:88-119, torch CPU fp32) on ALL host cores."""
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(ncpu)      # torchrun exports OMP_NUM_THREADS=1: the CPU arm must not inherit that
    step, _ = _oracle_trainer('cpu', batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=round(SEQ_LEN * batch / dt, 2), unit='frames/s', cores=torch.get_num_threads(), host_cpus=ncpu, kind='port',
                batch=batch, s_per_step=round(dt, 3),
                sample=f'{steps} training step(s) of the same workload at batch {batch} (of {BATCH}), torch {torch.__version__} CPU fp32, '
                       f'{torch.get_num_threads()} threads, {dt:.2f} s/step')


def run_reference(args):
    """The reference's CPU implementation of the path on the host cores (rank 0 only): bounded samples at two batch sizes, so that
    the extrapolation to the full batch (frames/s independent of the batch size) is shown, not assumed."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 3)), 1
    cb = cpu_baseline(steps=steps, warmup=warm, batch=8)
    cb2 = cpu_baseline(steps=1, warmup=1, batch=24)
    cb['second_sample'] = dict(batch=24, value=cb2['value'], s_per_step=cb2['s_per_step'])
    cb['sample'] += f'; second sample at batch 24: {cb2["value"]} frames/s ({cb2["s_per_step"]} s/step)'
    best = max(cb['value'], cb2['value'])
    cb['value'] = best
    world = int(os.environ.get('WORLD_SIZE', '1'))
    line = dict(metric=METRIC, value=best, unit='frames/s', impl='reference',
                n_gpus=world, steps=steps, warmup=warm, ms_per_step=round(SEQ_LEN * BATCH / best * 1e3, 1),
                ms_per_step_note=f'extrapolated to the full batch {BATCH} from the faster of the two samples (batch 8: {cb["s_per_step"]} s, batch 24: '
                                 f'{cb2["s_per_step"]} s per step)',
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload=WORKLOAD, global_batch=BATCH, seq_len=SEQ_LEN, parallelism=f'cpu, {cb["cores"]} threads (rank 0 only)'),
                cpu_baseline=cb, e2e=dict(value=best, unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-eager-baseline', action='store_true')
    ap.add_argument('--workload', default='bair', choices=['bair', 'smmnist', 'human_rollout'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='N > 1: weak = 192 videos per GPU (default); strong = the global batch 192 split over the GPUs (train.py:218-219). '
                         'The other mode is measured for a few steps as well and reported in the same line.')
    a = ap.parse_args()
    if a.gpus > 1 and 'WORLD_SIZE' not in os.environ and a.impl == 'ours':
        # `python bench.py --gpus N` without a launcher: start one process per GPU ourselves (same command the driver uses)
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={a.gpus}', '--master-addr', '127.0.0.1',
               '--master-port', str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if a.impl == 'ours' and world != a.gpus:
        sys.exit(f'bench.py: --gpus {a.gpus} but the launcher started {world} rank(s)')
    if a.workload == 'human_rollout':
        import bench_rollout
        bench_rollout.run(a)
        sys.exit(0)
    select_workload(a.workload)
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
