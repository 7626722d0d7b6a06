"""Multi-GPU parity worker (run under torchrun with 2+ GPUs; tests/test_gpu_multigpu.py launches it and checks its verdict):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

The FULL model (encoder, inference nets, latent loop, decoder, ELBO) runs one training step on a batch sharded over the ranks with the
reference's multi-GPU semantics -- SyncBatchNorm statistics over the global batch (train.py:283), gradients averaged over the ranks
(DistributedDataParallel, train.py:314; here parallel.GradBucket's single flat all-reduce) -- and is compared with
  (a) the CPU fp32 ORACLE on the GLOBAL batch with the same per-video random draws: ELBO terms and running statistics;
  (b) the oracle run at the storage precision of the CUDA path (EMULATE_BF16) on the GPU: every parameter gradient;
  (c) this repository's single-GPU run on the global batch: every parameter gradient (tight: same arithmetic, different sharding).
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import build_model, make_input, rel_l2  # noqa: E402
from srvp_b200 import elbo, parallel  # noqa: E402
from oracle import srvp_oracle as O  # noqa: E402  (the checker; never the thing measured)


def slice_randoms(rnd, lo, hi):
    out = {}
    for k, v in rnd.items():
        if k == 'eps_z':
            out[k] = [e[lo:hi].contiguous() for e in v]
        elif k == 't_w':
            out[k] = v[:, lo:hi].contiguous()
        else:
            out[k] = v[lo:hi].contiguous()
    return out


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    cfg = dict(nx=64, nc=3, nf=64, nhx=128, ny=50, nz=50, skipco=True, nt_inf=2, nh_inf=256, nlayers_inf=3, nh_res=512, nlayers_res=4,
               archi='vgg')
    loss_cfg = dict(obs_scale=0.71, beta_y=1.0, beta_z=1.0, l2_res=1.0)
    T, B, dt = 6, 8 * world, 0.5
    x = make_input(T, B, 3, 5)
    torch.manual_seed(7)
    rnd = O.draw_randoms(cfg, T, T, B, training=True)          # identical on every rank (same seed)
    lo, hi = parallel.shard_bounds(B, rank, world)

    # ours, sharded: SyncBatchNorm containers + one flat gradient all-reduce
    m = build_model(cfg, 1.41, 1)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m).to(dev).train()
    bucket = parallel.GradBucket(list(m.parameters()), early=list(m.decoder.parameters()))
    parallel.ACTIVE_BUCKET = bucket
    bucket.zero()
    m.injected_randoms = slice_randoms(rnd, lo, hi)
    xs = x[:, lo:hi].contiguous().to(dev)
    out = m(xs, T, dt=dt)
    loss, nll, kl_y, kl_z = elbo.elbo(out, xs, **loss_cfg)
    loss.backward()
    bucket.allreduce_mean()
    terms = torch.stack([loss.detach(), nll.detach(), kl_y.detach(), kl_z.detach()]).double()
    dist.all_reduce(terms)
    terms[0] /= world                                            # loss is batch-averaged per rank; the others are sums
    g_sharded = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    parallel.ACTIVE_BUCKET = None

    ok = True
    if rank == 0:
        # (a) CPU fp32 oracle on the global batch
        with torch.no_grad():
            o = O.forward(sd0, cfg, x, T, dt, rnd, training=True)
            ref = [float(v) for v in O.elbo(o, x, loss_cfg)]
        rel = [abs(float(terms[i]) - ref[i]) / abs(ref[i]) for i in range(4)]
        print(f'world={world} ELBO {float(terms[0]):.4f} vs oracle {ref[0]:.4f}: rel loss {rel[0]:.2e} nll {rel[1]:.2e} kl_y {rel[2]:.2e} kl_z {rel[3]:.2e}')
        ok &= rel[0] < 1e-4 and rel[1] < 1e-4 and rel[2] < 1e-2 and rel[3] < 1e-2
        # running statistics = those of the global batch (oracle: F.batch_norm on the global batch updates clones; recompute here)
        st = {}
        with torch.no_grad():
            O.forward(sd0, cfg, x, T, dt, rnd, training=True, stats_out=st)
        worst = 0.0
        sd = m.state_dict()
        for prefix, (mean, var) in st.items():
            worst = max(worst, float((sd[prefix + '.running_mean'].cpu() - 0.1 * mean).abs().max()),
                        float((sd[prefix + '.running_var'].cpu() - (0.9 + 0.1 * var)).abs().max() / (1 + float(var.abs().max()))))
        print(f'  running statistics vs global-batch oracle: worst abs deviation {worst:.2e}')
        ok &= worst < 2e-2
        # (b) every parameter gradient vs the bf16-emulating oracle on the GPU (fp32 arithmetic, TF32 off)
        tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        O.EMULATE_BF16, O.USE_ATEN_LSTM = True, False
        try:
            sdo = {k: v.to(dev).clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in sd0.items()}
            rg = {k: ([e.to(dev) for e in v] if isinstance(v, list) else v.to(dev)) for k, v in rnd.items()}
            O.elbo(O.forward(sdo, cfg, x.to(dev), T, dt, rg, training=True), x.to(dev), loss_cfg)[0].backward()
        finally:
            O.EMULATE_BF16, O.USE_ATEN_LSTM = False, True
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        e_or = sorted((rel_l2(g_sharded[k], sdo[k].grad), k) for k in g_sharded)
        print(f'  gradients vs bf16-emulating oracle (global batch): median rel-L2 {e_or[len(e_or) // 2][0]:.3e}, worst {e_or[-1][0]:.3e} ({e_or[-1][1]})')
        # rounding noise amplified through kinks at this tiny size (tests/test_gpu_parity_full.py measures it per tensor against the
        # bf16-emulating oracle's own deviation from fp32); the tight logic check of the sharding is (c)
        ok &= e_or[len(e_or) // 2][0] < 0.25 and e_or[-1][0] < 0.6
        # (c) single-GPU run of this repository on the global batch
        m1 = build_model(cfg, 1.41, 1).to(dev).train()
        m1.injected_randoms = rnd
        out1 = m1(x.to(dev), T, dt=dt)
        elbo.elbo(out1, x.to(dev), **loss_cfg)[0].backward()
        e_1 = sorted((rel_l2(g_sharded[k], p.grad), k) for k, p in m1.named_parameters())
        d_1 = dict((k, e) for e, k in e_1)
        print(f'  gradients vs single-GPU run (global batch): median rel-L2 {e_1[len(e_1) // 2][0]:.3e}, worst {e_1[-1][0]:.3e} ({e_1[-1][1]})')
        # Two bf16 runs that differ by one fp32 rounding (here: the order in which the batch-norm partial sums are added) have
        # INDEPENDENT rounding noise after a few layers (a flipped bf16 rounding is a full-ulp change, so a 1e-7 perturbation grows to
        # ulp level within ~5 layers), so deep layers can only agree at the noise level measured in (b). The sharding logic itself --
        # gradient averaging, local dgamma / dbeta, GLOBAL sums in the SyncBatchNorm backward -- is pinned on the last decoder block,
        # which sits next to the loss, after ALL of those steps and before the noise has grown:
        near = ['decoder.conv.3.1.weight', 'decoder.conv.3.0.1.weight', 'decoder.conv.3.0.1.bias', 'decoder.conv.3.0.0.weight']
        print('  next to the loss: ' + ', '.join(f'{k} {d_1[k]:.2e}' for k in near))
        ok &= all(d_1[k] < 1e-2 for k in near)
        ok &= e_1[len(e_1) // 2][0] < 1.6 * e_or[len(e_or) // 2][0] + 1e-3 and e_1[-1][0] < 1.6 * e_or[-1][0] + 1e-2
        print('MULTIGPU CHECK', 'PASS' if ok else 'FAIL', flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
