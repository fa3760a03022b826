"""Drop-in replacement for the reference's `autovc_replicate/proposed_autovc.py` (BASELINE config 5) on dvae_b200 kernels.

Same class names, constructor arguments, `state_dict` keys and `Generator.forward(x) -> (mel [B,1,64,80],
mel_postnet [B,1,64,80])` as the reference (autovc_replicate/proposed_autovc.py:41-220).  The network is built from the
same layer kinds as the Disentangled VAE, so it runs on the same kernels and the same engine building blocks; the
torch.nn sub-modules are parameter containers only.  Unlike the reference, importing this module has no side effect
(the reference runs a 10-sample forward at import time, :223-227).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from dvae_b200 import lib, ops
from dvae_b200.engine import Engine, GradSink, PreparedWeights
from model.disentangled_vae import ConvNorm, LinearNorm, _precision_tag

T_FRAMES, N_MELS = 64, 80
ENC_CONVS = [("encoder.convolutions.%d.0.conv" % i, "encoder.convolutions.%d.1" % i, lib.ACT_RELU) for i in range(3)]
DEC_CONVS = [("decoder.convolutions.%d.0.conv" % i, "decoder.convolutions.%d.1" % i, lib.ACT_RELU) for i in range(3)]
POST_CONVS = [("postnet.convolutions.%d.0.conv" % i, "postnet.convolutions.%d.1" % i,
               lib.ACT_TANH if i < 4 else lib.ACT_NONE) for i in range(5)]
LSTMS = {"encoder.lstm": (2, 2, 64), "decoder.lstm1": (1, 1, 512), "decoder.lstm2": (2, 1, 1024)}
LINEARS = ["encoder.latent_code.linear_layer", "decoder.dec_linear.linear_layer", "decoder.linear_projection.linear_layer"]


def _conv_bn_stack(chans, gain):
    return nn.ModuleList(nn.Sequential(ConvNorm(ci, co, kernel_size=5, stride=1, padding=2, dilation=1, w_init_gain=g),
                                       nn.BatchNorm1d(co)) for (ci, co), g in zip(chans, gain))


class Encoder(nn.Module):
    """3 x (Conv1d k5 + BN + ReLU) -> 2-layer BiLSTM(64) -> Linear 8192 -> latent_dim (:41-85)."""

    def __init__(self, dim_neck=64, latent_dim=256):
        super().__init__()
        self.dim_neck = dim_neck
        self.convolutions = _conv_bn_stack([(80, 512), (512, 512), (512, 512)], ["relu"] * 3)
        self.lstm = nn.LSTM(512, dim_neck, 2, batch_first=True, bidirectional=True)
        self.latent_code = LinearNorm(8192, latent_dim)


class Decoder(nn.Module):
    """Linear latent -> 8192 -> LSTM(512) -> 3 x (Conv1d + BN + ReLU) -> 2-layer LSTM(1024) -> Linear -> 80 (:93-136)."""

    def __init__(self, dim_neck, dim_emb, dim_pre, latent_dim=256):
        super().__init__()
        self.dec_linear = LinearNorm(latent_dim, 8192)
        self.lstm1 = nn.LSTM(dim_neck * 2, dim_pre, 1, batch_first=True)
        self.dim_neck, self.dim_emb, self.dim_pre, self.latent_dim = dim_neck, dim_emb, dim_pre, latent_dim
        self.convolutions = _conv_bn_stack([(dim_pre, dim_pre)] * 3, ["relu"] * 3)
        self.lstm2 = nn.LSTM(dim_pre, 1024, 2, batch_first=True)
        self.linear_projection = LinearNorm(1024, 80)


class Postnet(nn.Module):
    """Five Conv1d(k=5) + BatchNorm1d, tanh after the first four (:139-183)."""

    def __init__(self):
        super().__init__()
        self.convolutions = _conv_bn_stack([(80, 512), (512, 512), (512, 512), (512, 512), (512, 80)],
                                           ["tanh"] * 4 + ["linear"])


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, keep, x, *params):
        ctx.set_materialize_grads(False)
        outs, saved = module._run_forward(x, keep)
        ctx.module, ctx.saved = module, saved
        return outs

    @staticmethod
    def backward(ctx, g_mel, g_post):
        if ctx.saved is None:
            raise RuntimeError("backward through a forward that ran without gradient tracking")
        grads = ctx.module._run_backward(ctx.saved, g_mel, g_post)
        ctx.saved = None
        return (None, None, None) + tuple(grads[n] for n in ctx.module._param_names)


class Generator(nn.Module):
    """Generator network (:187-220)."""

    def __init__(self, dim_neck=64, dim_emb=256, dim_pre=512, precision: Optional[str] = None):
        super().__init__()
        if dim_neck != 64 or dim_pre != 512:
            raise ValueError("dvae_b200 implements the shipped geometry only (dim_neck=64, dim_pre=512)")
        self.encoder = Encoder(dim_neck, dim_emb)
        self.decoder = Decoder(dim_neck, dim_emb, dim_pre)
        self.postnet = Postnet()
        self._dt = _precision_tag(precision)
        # fp16 storage carries the activation-gradient stream scaled by a power of two (Engine.grad_scale).  The reference
        # defines no loss for this network, so the scale follows the incoming gradient: `_pick_grad_scale` lifts its largest
        # magnitude to ~4 (DVAE_B200_GRAD_SCALE or `.grad_scale = ...` pins it instead).
        env_scale = float(os.environ.get("DVAE_B200_GRAD_SCALE", 0))
        self._engine = Engine(self._dt, dim_emb, 1, grad_scale=env_scale or 1.0)
        self._auto_scale = self._dt == lib.F16 and not env_scale
        self._amax_reader = None
        self._param_names = [n for n, _ in self.named_parameters()]
        self._prep_cache = None
        self._debug_keep_saved = False
        self._last_saved = None

    @property
    def grad_scale(self) -> float:
        return self._engine.grad_scale

    @grad_scale.setter
    def grad_scale(self, v: float) -> None:
        self._engine.grad_scale = float(v)
        self._auto_scale = False

    def _pick_grad_scale(self, grads) -> None:
        """fp16 mode: power-of-two scale that puts the largest incoming output gradient near 4.  The magnitude is read back
        asynchronously, one step late (the first backward waits for it once), so the launch queue never drains."""
        import math
        from dvae_b200.data import AsyncScalars
        amax = torch.stack([g.detach().abs().max() for g in grads if g is not None]).max().reshape(1).float()
        if self._amax_reader is None:
            self._amax_reader = AsyncScalars(1, amax.device)
            self._amax_reader.push(amax)
            prev = [amax.item()]
        else:
            prev = self._amax_reader.push(amax)
        a = prev[0] if prev else 0.0
        if a > 0.0 and math.isfinite(a):
            self._engine.grad_scale = float(2.0 ** math.floor(math.log2(4.0 / a)))

    def _prepared(self) -> PreparedWeights:
        """See DisentangledVAE._prepared: rebuilt on re-allocation, refreshed in place when a parameter was written."""
        params = list(self.parameters())
        ptrs = tuple(p.data_ptr() for p in params)
        vers = tuple(p._version for p in params)
        c = self._prep_cache
        P = {n: p.data for n, p in self.named_parameters()}
        if c is None or c[0] != ptrs:
            if params[0].device.type != "cuda":
                raise RuntimeError("dvae_b200 runs on CUDA (sm_100a) only: move the module with .to('cuda'); there is no CPU path")
            W = PreparedWeights(self._dt, P, convs=[c_ for c_, _, _ in ENC_CONVS + DEC_CONVS + POST_CONVS], linears=LINEARS,
                                lstms=LSTMS, fused_heads=False)
            c = self._prep_cache = [ptrs, vers, W]
        elif c[1] != vers:
            c[2].refresh(P)
            c[1] = vers
        return c[2]

    # ------------------------------------------------------------------ schedule
    def _run_forward(self, x, keep):
        E, dt = self._engine, self._dt
        W = self._prepared()
        P = {n: p.data for n, p in self.named_parameters()}
        Bf = dict(self.named_buffers())
        R = x.shape[0]
        assert tuple(x.shape[1:]) == (N_MELS, T_FRAMES), f"expected [B, 80, 64], got {tuple(x.shape)}"
        saved: Optional[dict] = {} if keep else None
        sv = (lambda: [] if keep else None)
        s_enc_c, s_enc_l, s_d1, s_dc, s_d2, s_post = sv(), sv(), sv(), sv(), sv(), sv()
        x_cl = torch.empty((R, T_FRAMES, N_MELS), device=x.device, dtype=ops.act_dtype(dt))
        ops.pack_ncl_to_cl(dt, x, x_cl)
        h = E._conv_stack(W, P, Bf, x_cl, ENC_CONVS, 1, self.training, s_enc_c)
        h = E._lstm(W, "encoder.lstm", h, s_enc_l)
        flat = h.view(R, T_FRAMES * 128)
        codes, _ = ops.linear_fwd(dt, flat, W.lin[LINEARS[0]], P[LINEARS[0] + ".bias"])
        d, _ = ops.linear_fwd(dt, codes, W.lin[LINEARS[1]], P[LINEARS[1] + ".bias"])
        h = E._lstm(W, "decoder.lstm1", d.view(R, T_FRAMES, 128), s_d1)
        h = E._conv_stack(W, P, Bf, h, DEC_CONVS, 1, self.training, s_dc)
        h = E._lstm(W, "decoder.lstm2", h, s_d2)
        rec, rec32 = ops.linear_fwd(dt, h.view(R * T_FRAMES, 1024), W.lin[LINEARS[2]], P[LINEARS[2] + ".bias"], want_f32=True)
        rec = rec.view(R, T_FRAMES, N_MELS)
        post = E._conv_stack(W, P, Bf, rec, POST_CONVS, 1, self.training, s_post)
        rec32 = rec32.view(R, 1, T_FRAMES, N_MELS)
        mel_post = ops.add_f32_act(dt, rec32, post.view(R, 1, T_FRAMES, N_MELS))
        if keep:
            saved.update(R=R, W=W, enc_convs=s_enc_c, enc_lstm=s_enc_l, flat=flat, codes=codes, d=d, dec_lstm1=s_d1,
                         dec_convs=s_dc, dec_lstm2=s_d2, h_top=h, post_convs=s_post)
            if self._debug_keep_saved:
                self._last_saved = saved
        return (rec32, mel_post), saved

    def _run_backward(self, saved, g_mel, g_post) -> Dict[str, torch.Tensor]:
        E, dt = self._engine, self._dt
        W, R = saved["W"], saved["R"]
        ad = ops.act_dtype(dt)
        dev = saved["flat"].device
        sink = GradSink(dev, E.buckets, E._sink_layout)
        if self._auto_scale:
            self._pick_grad_scale((g_mel, g_post))
        shape = (R, T_FRAMES, N_MELS)
        d_post = torch.zeros(shape, device=dev, dtype=ad)
        d_rec = torch.zeros(shape, device=dev, dtype=ad)
        if g_post is not None:
            ops.prep_cast(dt, g_post.contiguous().view(shape), d_post, scale=E.grad_scale)
            ops.prep_cast(dt, g_post.contiguous().view(shape), d_rec, scale=E.grad_scale)
        if g_mel is not None:
            tmp = torch.empty(shape, device=dev, dtype=ad)
            ops.prep_cast(dt, g_mel.contiguous().view(shape), tmp, scale=E.grad_scale)
            ops.add_inplace(dt, d_rec, tmp)
        d_in = E._conv_stack_bwd(W, d_post, saved["post_convs"], sink, 1, need_dx=True)
        ops.add_inplace(dt, d_rec, d_in)
        dh = E._linear_bwd(LINEARS[2], W.lin[LINEARS[2]], d_rec.view(R * T_FRAMES, N_MELS),
                           saved["h_top"].view(R * T_FRAMES, 1024), sink)
        dh = E._lstm_bwd(W, "decoder.lstm2", dh, saved["dec_lstm2"], sink, need_dx=True)
        dh = E._conv_stack_bwd(W, dh, saved["dec_convs"], sink, 1, need_dx=True)
        dh = E._lstm_bwd(W, "decoder.lstm1", dh, saved["dec_lstm1"], sink, need_dx=True)
        d_codes = E._linear_bwd(LINEARS[1], W.lin[LINEARS[1]], dh.reshape(R, T_FRAMES * 128), saved["codes"], sink)
        d_flat = E._linear_bwd(LINEARS[0], W.lin[LINEARS[0]], d_codes, saved["flat"], sink)
        dh = E._lstm_bwd(W, "encoder.lstm", d_flat.view(R, T_FRAMES, 128), saved["enc_lstm"], sink, need_dx=True)
        E._conv_stack_bwd(W, dh, saved["enc_convs"], sink, 1, need_dx=False)
        E._join_side_stream()
        if E._sink_layout is None:
            E._sink_layout = sink.next_layout()
        return sink.finish()

    def forward(self, x):
        """x [B, 80, 64] -> (mel_outputs [B,1,64,80], mel_outputs_postnet [B,1,64,80])  (:196-220)."""
        x = x.detach().to(torch.float32).contiguous()
        params = list(self.parameters())
        keep = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if keep and not self.training:
            raise NotImplementedError("gradients in eval() mode (running-stat BatchNorm) are not implemented")
        return _GeneratorFn.apply(self, keep, x, *params)
