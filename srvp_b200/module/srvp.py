"""B200-native StochasticLatentResidualVideoPredictor with the reference's class surface (module/srvp.py:29-470).

Same constructor signature, attributes, sub-module names (state-dict keys), method names, argument
meanings, return tuples and train/eval behaviour as the reference class, so `train.py` / `test.py` written against the
reference run unchanged. The compute is hand-written sm_100a CUDA reached through the C ABI (srvp_b200/_lib.py);
there is no PyTorch/CPU fallback for the conv-VAE path.

Random draws follow the reference's consumption order: skip-frame `randint`, per-video `randperm`,
y_0 noise, then one z noise per observed frame. `noise_device` selects which torch generator the Gaussian noise comes
from: 'cpu' (default; bit-identical stream to the reference run on CPU, used for parity) or 'cuda'.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import conv, utils
from .mlp import MLP
from .. import engine, infer


class StochasticLatentResidualVideoPredictor(nn.Module):
    def __init__(self, nx, nc, nf, nhx, ny, nz, skipco, nt_inf, nh_inf, nlayers_inf, nh_res, nlayers_res, archi):
        super().__init__()
        self.nx, self.nc, self.ny, self.nz = nx, nc, ny, nz
        self.skipco = skipco
        self.nt_inf, self.nh_inf, self.nlayers_inf = nt_inf, nh_inf, nlayers_inf
        self.nh_res, self.nlayers_res = nh_res, nlayers_res
        self.nhx = nhx
        self.noise_device = 'cpu'
        # Test hook: a dict of explicit random draws {'t_skip', 't_w', 'eps_y', 'eps_z': [...]} (the layout of oracle.draw_randoms,
        # SURVEY.md App. D) consumed instead of drawing -- lets a sharded run use its slice of the global batch's draws.
        self.injected_randoms = None
        # construction order = reference order (module/srvp.py:124-137): it fixes the same-seed default initialisation
        self.encoder = conv.encoder_factory(archi, nx, nc, nhx, nf)
        self.decoder = conv.decoder_factory(archi, nx, nc, nh_inf + ny, nf, skipco)
        self.w_proj = nn.Sequential(nn.Linear(nhx, nh_inf), nn.ReLU(inplace=True))
        self.w_inf = nn.Sequential(nn.Linear(nh_inf, nh_inf), nn.Tanh())
        self.q_y = MLP(nhx * nt_inf, nh_inf, ny * 2, nlayers_inf)
        self.inf_z = nn.LSTM(nhx, nh_inf, 1)
        self.q_z = nn.Linear(nh_inf, nz * 2)
        self.p_z = MLP(ny, nh_res, nz * 2, nlayers_res)
        self.dynamics = MLP(ny + nz, nh_res, ny, nlayers_res)

    def init(self, res_gain=1.41):
        """module/srvp.py:139-154"""
        for net in (self.encoder, self.decoder):
            net.apply(lambda m: utils.init_weight(m, init_type='normal', init_gain=0.02))
        self.dynamics.apply(lambda m: utils.init_weight(m, init_type='orthogonal', init_gain=res_gain))

    # ------------------------------------------------------------------------------------------------ noise
    @staticmethod
    def _to_device(t, device):
        """Host tensor -> device without stalling the host: a pageable-memory copy waits for the stream to drain (a hidden host
        synchronisation per call); staging through the caching pinned allocator keeps it asynchronous and stream-safe."""
        return t.pin_memory().to(device, non_blocking=True)

    def _normal(self, shape, like):
        if self.noise_device == 'cpu':
            return self._to_device(torch.empty(shape, dtype=torch.float32).normal_(), like.device)
        return torch.empty(shape, dtype=torch.float32, device=like.device).normal_()

    def _rsample(self, raw_params, key=None):
        shape = (*raw_params.shape[:-1], raw_params.shape[-1] // 2)
        inj = self.injected_randoms
        if inj is not None and key is not None and key in inj:
            return infer.rsample(raw_params, self._to_device(inj[key].float().reshape(shape), raw_params.device))
        return infer.rsample(raw_params, self._normal(shape, raw_params))

    # ------------------------------------------------------------------------------------------------ encode / decode
    def _encode_fused(self, x):
        """Returns hx (T, B, nhx) and a SkipHandle (or None). Skip frames: random per video in training, last otherwise."""
        nt, bsz = x.shape[0], x.shape[1]
        handle = engine.SkipHandle() if self.skipco else None
        hx = engine.encoder_apply(self.encoder, x.reshape(nt * bsz, *x.shape[2:]), handle).view(nt, bsz, self.nhx)
        if self.skipco:
            inj = self.injected_randoms
            if self.training:
                t = inj['t_skip'].long() if inj is not None and 't_skip' in inj else torch.randint(nt, size=(bsz,))
            else:
                t = torch.full((bsz,), nt - 1, dtype=torch.long)
            sel = (t * bsz + torch.arange(bsz)).to(torch.int32)          # encoder frame feeding each video's skip
            inv = torch.full((nt * bsz,), -1, dtype=torch.int32)
            inv[sel.long()] = torch.arange(bsz, dtype=torch.int32)
            handle.T, handle.B = nt, bsz
            handle.frame_map = self._to_device(sel, x.device)             # (B,), expanded per decode call
            handle.inv_map = self._to_device(inv, x.device)
        return hx, handle

    def encode(self, x):
        """module/srvp.py:156-193. Skips are returned as (B, C, H, W) fp32 tensors, deepest first."""
        hx, handle = self._encode_fused(x)
        if handle is None:
            return hx, None
        if self.training and torch.is_grad_enabled() and not getattr(self, '_warned_skip_grad', False):
            import warnings
            warnings.warn('srvp_b200: encode() returns the skip connections as materialised tensors outside autograd: gradients reach the '
                          'encoder through hx only. Training code should call forward() (as the reference train.py does), which routes the '
                          'skip gradients back to the encoder.')
            self._warned_skip_grad = True
        from .. import ops
        skips = []
        for (z, st, C, res) in handle.levels:
            src = ops.Src(z, C, st.scale if st is not None else None, st.shift if st is not None else None, handle.frame_map, 0, 0, True)
            skips.append(ops.nhwc_to_nchw_f32(ops.materialize(src, x.shape[1], res, res), C))
        return hx, skips

    def _decode_fused(self, w, y, levels, sel, handle):
        nt, bsz = y.shape[0], y.shape[1]
        dec_inp = torch.cat([w.repeat(nt, 1, 1).view(nt * bsz, self.nh_inf), y.reshape(nt * bsz, self.ny)], 1)
        frame_map = None
        if levels is not None:
            frame_map = sel.repeat(nt)                                    # decoder frame (t, b) -> source frame of video b
        x_flat = engine.decoder_apply(self.decoder, dec_inp, levels, frame_map, handle, sel=sel)
        return x_flat.view(nt, bsz, *x_flat.shape[1:])

    def decode(self, w, y, skip):
        """module/srvp.py:195-227. skip: list of (B, C, H, W) tensors (deepest first) or None."""
        assert skip is None and not self.skipco or self.skipco and skip is not None
        levels, sel = None, None
        if skip is not None:
            from .. import ops
            levels = [ops.nchw_to_nhwc_bf16(s.contiguous().float(), s.shape[1]) for s in skip]
            sel = torch.arange(y.shape[1], dtype=torch.int32, device=y.device)
        return self._decode_fused(w, y, levels, sel, None)

    # ------------------------------------------------------------------------------------------------ inference nets
    def infer_w(self, hx):
        """module/srvp.py:229-256"""
        nt, bsz = hx.shape[0], hx.shape[1]
        if self.training:
            inj = self.injected_randoms
            t = inj['t_w'].long() if inj is not None and 't_w' in inj else torch.stack([torch.randperm(nt)[:self.nt_inf] for _ in range(bsz)], 1)
            t = self._to_device(t, hx.device)
            index = torch.arange(bsz, device=hx.device).repeat(self.nt_inf, 1)
            h = hx[t.view(-1), index.view(-1)].view(self.nt_inf, bsz, self.nhx)
        else:
            h = hx[-self.nt_inf:]
        return infer.linear(infer.linear(h, self.w_proj[0], 'relu').sum(0), self.w_inf[0], 'tanh')

    def infer_y(self, hx):
        """module/srvp.py:258-278"""
        q_y_0_params = infer.mlp(hx.permute(1, 0, 2).reshape(hx.shape[1], self.nt_inf * self.nhx), self.q_y)
        return self._rsample(q_y_0_params, 'eps_y'), q_y_0_params

    def infer_z(self, hx):
        """module/srvp.py:280-298"""
        q_z_params = infer.linear(hx, self.q_z)
        return self._rsample(q_z_params), q_z_params

    def _residual_step(self, y_t, z_tp1, dt):
        """module/srvp.py:300-323"""
        res_tp1 = dt * infer.mlp(torch.cat([y_t, z_tp1], 1), self.dynamics)
        return y_t + res_tp1, res_tp1

    def generate(self, y_0, hx, nt, dt, remove_intermediate=True):
        """module/srvp.py:325-413. The whole Euler loop is one persistent kernel launch (srvp_b200/csrc/latent.cu)."""
        from .. import latent
        assert (1 / dt).is_integer()
        oversampling = int(1 / dt)
        bsz = y_0.shape[0]
        n_obs = len(hx)
        n_post = max(0, min(nt, n_obs) - 1)          # frames 1..n_post have an observation: z ~ q(z | x)  (srvp.py:385-388)
        if n_post < nt - 1:
            assert not self.training                 # srvp.py:391
        q_z_params, z_post = None, None
        if n_obs > 0:
            hx_z = infer.lstm(hx, self.inf_z)                # one launch for the whole recurrence (srvp.py:365-368)
        # noise in the reference's order: one (B, nz) draw per generated frame (posterior or prior alike)
        inj = self.injected_randoms
        if inj is not None and 'eps_z' in inj and nt > 1:
            eps = self._to_device(torch.stack([e.float() for e in inj['eps_z'][:nt - 1]]), y_0.device)
        else:
            eps = torch.stack([self._normal((bsz, self.nz), y_0) for _ in range(nt - 1)]) if nt > 1 else None
        if n_post > 0:
            q_z_params = infer.linear(hx_z[1:n_post + 1], self.q_z)
            z_post = infer.rsample(q_z_params, eps[:n_post])
        y_all, p_z_params, z, res = latent.latent_loop(self.p_z, self.dynamics, y_0, z_post, eps, nt, oversampling, float(dt), n_post,
                                                        self.nh_res)
        y = y_all[::oversampling] if remove_intermediate else y_all
        if n_post == nt - 1 and z_post is not None:
            z = z_post                              # keep the differentiable posterior samples in the returned tuple
        return y, (z if nt > 1 else None), q_z_params, (p_z_params if nt > 1 else None), res

    def forward(self, x, nt, dt, remove_intermediate=True):
        """module/srvp.py:415-470: returns (x_, y, z, w, q_y_0_params, q_z_params, p_z_params, res)."""
        hx, handle = self._encode_fused(x)
        w = self.infer_w(hx)
        y_0, q_y_0_params = self.infer_y(hx[:self.nt_inf])
        y, z, q_z_params, p_z_params, res = self.generate(y_0, hx, nt, dt, remove_intermediate=remove_intermediate)
        levels = handle.levels if handle is not None else None
        sel = handle.frame_map if handle is not None else None
        x_ = self._decode_fused(w, y, levels, sel, handle)
        return x_, y, z, w, q_y_0_params, q_z_params, p_z_params, res
