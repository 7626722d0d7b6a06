"""Benchmark of the SRVP hot path: frames/s of one full training step (forward + ELBO + backward + Adam).

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
  workload  BAIR config of BASELINE.json: VGG64 skipco nc=3 64x64, seq_len 12, batch 192 per GPU (weak scaling),
            ny=nz=50, nt_inf=2, 2 Euler steps, obs_scale 0.71 (reference README.md:127); synthetic uniform frames.
  value     frames/s with the batch already resident in HBM (device timed, CUDA events, max over ranks)
  e2e       same step through the public API from pinned HOST frames: H2D of the batch and D2H of the loss inside the
            timed region
  roofline  dominant kernel family: algorithmic dense FLOPs / CUDA-event time vs the measured bf16 peak
  cpu_baseline / --impl reference: the reference algorithm (oracle port, torch CPU fp32) on the host cores, bounded sample
"""
import argparse
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


def run_ours(args):
    import torch.distributed as dist
    from srvp_b200 import ops, _lib
    from srvp_b200.module.srvp import StochasticLatentResidualVideoPredictor
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(1)
    model = StochasticLatentResidualVideoPredictor(*[CFG[k] for k in ARG_ORDER])
    model.init(res_gain=1.41)
    if world > 1:
        # reference multi-GPU semantics (train.py:283): batch-norm statistics over the GLOBAL batch
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model = model.to(dev).train()
    model.noise_device = 'cuda'
    params = [p for p in model.parameters()]
    from srvp_b200.optim import Adam
    opt = Adam(params, lr=3e-4)      # torch.optim.Adam semantics (train.py:289), all tensors in one launch
    gen = torch.Generator().manual_seed(123 + rank)
    # several distinct host batches (pinned) so that the e2e loop really moves data: uint8 (B, T, H, W, C) frames as the datasets
    # store them (28 MB per batch); the device-resident batch of the `value` loop is their fp32 (T, B, C, H, W) conversion (113 MB)
    host = [torch.randint(0, 256, (BATCH, SEQ_LEN, 64, 64, CFG['nc']), dtype=torch.uint8, generator=gen).pin_memory() for _ in range(2)]
    xdev = ops.u8_to_tbchw_f32(host[0].to(dev))

    def sync_grads():
        if world > 1:
            flat = torch._utils._flatten_dense_tensors([p.grad for p in params])
            dist.all_reduce(flat)
            flat.div_(world)
            for p, g in zip(params, torch._utils._unflatten_dense_tensors(flat, [p.grad for p in params])):
                p.grad.copy_(g)

    def step(x):
        opt.zero_grad(set_to_none=True)
        out = model(x, SEQ_LEN, dt=DT)
        loss = elbo_loss(out, x)
        loss.backward()
        sync_grads()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)   # NVML is initialised and polled from before the warm-up so that it cannot disturb the timed region
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(xdev)
    barrier()
    sampler.recording = True
    launches0 = _lib.lib().srvp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(xdev)
    e1.record()
    barrier()
    launches = _lib.lib().srvp_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    # end-to-end: host frames -> device, step, loss -> host, every step
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        xb = ops.u8_to_tbchw_f32(host[i % len(host)].to(dev, non_blocking=True))
        lv = step(xb).item()
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # per-kernel profile (2 extra steps with CUDA events around every launch)
    ops.PROFILE = {}
    for _ in range(2):
        step(xdev)
    torch.cuda.synchronize()
    prof = ops.summarize_profile(ops.PROFILE)
    ops.PROFILE = None
    pk = peaks()
    tot_ms = sum(v['ms'] for v in prof.values())
    dom = max((k for k in prof if prof[k]['flops'] > 0), key=lambda k: prof[k]['ms'])
    d = prof[dom]
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
    # north star: the encoder convolutions against the tensor-core roofline, the decoder against both (its convolutions are tensor
    # bound); figures of the conv3x3 launches of each stage
    stages = {}
    for t, v in prof.get('conv3x3', {}).get('by_stage', {}).items():
        tf, gb = v['flops'] / (v['ms'] * 1e-3) / 1e12, v['bytes'] / (v['ms'] * 1e-3) / 1e9
        stages[t] = dict(ms_per_step=round(v['ms'] / 2, 3), tflops=round(tf, 1), frac_tensor=round(tf / pk['tf'], 4), gbs=round(gb, 1),
                         frac_hbm=round(gb / pk['hbm'], 4))
    roofline['conv_stages'] = stages
    roofline['executed'] = round(d['executed_flops'] / (d['ms'] * 1e-3) / 1e12, 1)
    roofline['note'] = ('achieved = dense FLOPs of the reference ops these launches stand for (SURVEY.md 8d) / time; executed = multiply-adds '
                        'actually issued (the convolutions over cat[h, skip] are split per video, DESIGN.md section 4)')
    breakdown = {k: dict(ms_per_step=round(v['ms'] / 2, 3), launches=v['launches'] // 2,
                         tflops=round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 1) if v['flops'] else None,
                         executed_tflops=round(v['executed_flops'] / (v['ms'] * 1e-3) / 1e12, 1) if v['executed_flops'] else None,
                         gbs=round(v['bytes'] / (v['ms'] * 1e-3) / 1e9, 1) if v['bytes'] else None) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    frames = SEQ_LEN * BATCH * world
    line = dict(metric=METRIC, value=round(frames * args.steps / (ms * 1e-3), 1), unit='frames/s',
                n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=round(ms / args.steps, 3), higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='bf16', data='synthetic',
                config=dict(workload=WORKLOAD,
                            global_batch=BATCH * world, seq_len=SEQ_LEN, parallelism=f'dp{world}' + (' + SyncBN statistics' if world > 1 else ''),
                            l2='inputs larger than L2: 113 MB batch, >10 GB of activations per step',
                            e2e_input='uint8 frames (B,T,H,W,C) from pinned host memory, converted on the device'),
                e2e=dict(value=round(frames * args.steps / (ms_e2e * 1e-3), 1), unit='frames/s', h2d_bytes_per_step=host[0].numel(),
                         d2h_bytes_per_step=4),
                gpu_launches=int(launches), clocks=sampler.summary(), roofline=roofline, kernel_breakdown=breakdown, loss=lv)
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(steps=1, warmup=1)
    print(json.dumps(line), flush=True)


def cpu_baseline(steps, warmup, batch=8):
    """The reference algorithm (oracle port of module/srvp.py + train.py:88-119, torch CPU fp32) on the host cores."""
    from oracle import srvp_oracle as O
    from srvp_b200.module.srvp import StochasticLatentResidualVideoPredictor
    torch.manual_seed(1)
    m = StochasticLatentResidualVideoPredictor(*[CFG[k] for k in ARG_ORDER])
    m.init(res_gain=1.41)
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in m.state_dict().items()}
    ps = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(ps, lr=3e-4)
    x = torch.rand(SEQ_LEN, batch, CFG['nc'], 64, 64, generator=torch.Generator().manual_seed(123))

    def step():
        opt.zero_grad()
        rnd = O.draw_randoms(CFG, SEQ_LEN, SEQ_LEN, batch, training=True)
        o = O.forward(sd, CFG, x, SEQ_LEN, DT, rnd, training=True)
        loss = O.elbo(o, x, LOSS)[0]
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=round(SEQ_LEN * batch / dt, 2), unit='frames/s', cores=torch.get_num_threads(), host_cpus=os.cpu_count(), kind='port',
                sample=f'{steps} training step(s) of the same workload at batch {batch} (of {BATCH}), torch {torch.__version__} CPU fp32, '
                       f'{dt:.2f} s/step; frames/s is batch-size independent on CPU')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 3)), 1
    cb = cpu_baseline(steps=steps, warmup=warm)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    line = dict(metric=METRIC, value=cb['value'], unit='frames/s', impl='reference',
                n_gpus=world, steps=steps, warmup=warm, ms_per_step=round(SEQ_LEN * 8 / cb['value'] * 1e3, 1), higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload=WORKLOAD,
                            global_batch=BATCH * world, seq_len=SEQ_LEN, parallelism='cpu'),
                cpu_baseline=cb, e2e=dict(value=cb['value'], unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='bair', choices=['bair', 'smmnist'])
    a = ap.parse_args()
    select_workload(a.workload)
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
