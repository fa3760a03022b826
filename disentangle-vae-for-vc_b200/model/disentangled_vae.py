"""Drop-in replacements for the reference's `model/disentangled_vae.py` on dvae_b200 kernels (sm_100a only).

Same import path, class names, constructor arguments, method names, returned tuples and `state_dict` keys as the
reference (v-manhlt3/Disentangle-VAE-for-VC, model/disentangled_vae.py:43-354), so `train.py` / `conversion.sh` run
unchanged.  The torch.nn sub-modules below are PARAMETER CONTAINERS ONLY (they give the 84 + 33 state_dict entries
their reference names, shapes and initialisation); their `forward` is never called.  All arithmetic runs in
libdvae_b200.so through `dvae_b200.engine`; there is no PyTorch / CPU fallback.

Extras that do not change the reference API:
  * `DisentangledVAE(..., precision="fp16"|"tf32"|"bf16"|"fp32")` (last, optional) or env DVAE_B200_PRECISION
  * `DisentangledVAE.noise_hook`: callable(shape) -> fp32 CPU/CUDA tensor, to supply the reparameterisation noise
    externally (the reference draws it on the CPU default generator, model/disentangled_vae.py:224)
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn
from torch import optim

from dvae_b200 import lib, ops
from dvae_b200 import optim as fused_optim
from dvae_b200.engine import Engine, PreparedWeights
from model.variational_base_vae import VariationalBaseModelVAE


PRECISIONS = {"bf16": lib.BF16, "tf32": lib.TF32, "fp16": lib.F16, "fp32": lib.F32}   # fp32: strict mode, CUDA cores (checking)
DEFAULT_PRECISION = "fp16"   # the fastest storage type that meets the reference-parity tolerance (DESIGN.md "Numerics")


def _precision_tag(precision: Optional[str]) -> int:
    p = (precision or os.environ.get("DVAE_B200_PRECISION", DEFAULT_PRECISION)).lower()
    if p not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {p!r}")
    return PRECISIONS[p]


def default_grad_scale(dt: int, batch_size) -> float:
    """Power-of-two factor the activation-gradient stream is carried at (dvae_b200.engine.Engine.grad_scale).  Only the
    fp16 storage mode needs one: the loss hands back gradients of magnitude mse_cof / batch_size
    (model/disentangled_vae.py:314-318), which this lifts to O(1).  Measured with random-init weights
    (scripts/diag_gradscale.py, profiles/r02_gradscale.txt): the stream then peaks at ~1.6e3 (first decoder convolution)
    -- 40x below fp16's 65504 -- and its smallest per-tensor maximum is ~3, 5e4 above the smallest normal number.
    Override with DVAE_B200_GRAD_SCALE or `model.grad_scale = ...`."""
    env = os.environ.get("DVAE_B200_GRAD_SCALE")
    if env:
        return float(env)
    if dt != lib.F16:
        return 1.0
    import math
    return float(2.0 ** (int(math.floor(math.log2(max(int(batch_size), 1)))) - 2))


def init_weights(m):
    """model/disentangled_vae.py:26-32: xavier-uniform weights; Linear bias 0.01, Conv1d bias 0."""
    if type(m) == nn.Linear:
        torch.nn.init.xavier_uniform_(m.weight)
        m.bias.data.fill_(0.01)
    if type(m) == nn.Conv1d:
        torch.nn.init.xavier_uniform_(m.weight)
        m.bias.data.fill_(0)


def tile(a, dim, n_tile):
    """model/disentangled_vae.py:35-41: every slice along `dim` repeated n_tile times in place (a repeat-interleave; the
    reference builds the index with a hard-coded torch.cuda.LongTensor).  Unused by the reference; kept for API parity."""
    return torch.repeat_interleave(a, n_tile, dim=dim)


class LinearNorm(nn.Module):
    """Container for `linear_layer` (model/disentangled_vae.py:90-100)."""

    def __init__(self, in_dim, out_dim, bias=True, w_init_gain="linear"):
        super().__init__()
        self.linear_layer = nn.Linear(in_dim, out_dim, bias=bias)
        torch.nn.init.xavier_uniform_(self.linear_layer.weight, gain=torch.nn.init.calculate_gain(w_init_gain))


class ConvNorm(nn.Module):
    """Container for `conv` (model/disentangled_vae.py:103-121)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear"):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)
        torch.nn.init.xavier_uniform_(self.conv.weight, gain=torch.nn.init.calculate_gain(w_init_gain))


class Postnet(nn.Module):
    """Five Conv1d(k=5) + BatchNorm1d, tanh after the first four (model/disentangled_vae.py:43-87)."""

    def __init__(self):
        super().__init__()
        chans = [(80, 512, "tanh"), (512, 512, "tanh"), (512, 512, "tanh"), (512, 512, "tanh"), (512, 80, "linear")]
        self.convolutions = nn.ModuleList(
            nn.Sequential(ConvNorm(ci, co, kernel_size=5, stride=1, padding=2, dilation=1, w_init_gain=g),
                          nn.BatchNorm1d(co)) for ci, co, g in chans)

    def forward(self, x):
        owner = self.__dict__.get("_owner")
        if owner is None:
            raise RuntimeError("Postnet must be used through its owning DisentangledVAE (dvae_b200 engine)")
        return owner._postnet_forward(x)


class _NetworkFn(torch.autograd.Function):
    """One autograd node for the whole network: forward = engine.forward, backward = engine.backward."""

    @staticmethod
    def forward(ctx, module, keep, sample_content, x1, x2, e1, e2, e3, *params):
        ctx.set_materialize_grads(False)
        W = module._prepared()
        P, B = module._param_dict(), module._buffer_dict()
        outs, saved = module._engine.forward(W, P, B, x1, x2, (e1, e2, e3), module.training, sample_content, keep)
        ctx.engine, ctx.W, ctx.saved, ctx.names, ctx.params = module._engine, W, saved, module._param_names, params
        if module._debug_keep_saved:
            module._last_saved = saved      # diagnostics / parity tests only
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        if ctx.saved is None:
            raise RuntimeError("backward through a forward that ran without gradient tracking")
        if ctx.engine.buckets is not None and any(p.grad is not None for p in ctx.params):
            # autograd adopts the returned bucket views as p.grad without copying; a second backward would overwrite them in
            # place instead of accumulating
            raise RuntimeError("bucketed (data-parallel) backward needs p.grad to be None: call optimizer.zero_grad() "
                               "(set_to_none=True) before every backward; gradient accumulation is not supported with buckets")
        grads = ctx.engine.backward(ctx.W, ctx.saved, gouts)
        ctx.saved = None
        return (None,) * 8 + tuple(grads[n] for n in ctx.names)


class DisentangledVAE(nn.Module):
    """Reference: model/disentangled_vae.py:124-286."""

    def __init__(self, speaker_size, input_sz=(1, 64, 80), kernel_szs=[512, 512, 512], hidden_sz: int = 256,
                 latent_sz: int = 32, c: float = 512, c_delta: float = 0.001, beta: float = 0.1, beta_delta: float = 0,
                 dim_neck=64, latent_dim=64, dim_pre=512, batch_size=10, precision: Optional[str] = None):
        super().__init__()
        if dim_neck != 64 or dim_pre != 512:
            raise ValueError("dvae_b200 implements the shipped geometry only (dim_neck=64, dim_pre=512)")
        self.batch_size = batch_size
        self._input_sz = input_sz
        self._channel_szs = [input_sz[0]] + kernel_szs
        self._hidden_sz = hidden_sz
        self._c, self._c_delta = c, c_delta
        self._beta, self._beta_delta = beta, beta_delta
        self.latent_dim = latent_dim
        self.dim_neck = dim_neck
        self.speaker_size = speaker_size

        self.postnet = Postnet()
        self.enc_modules = nn.ModuleList(
            nn.Sequential(ConvNorm(80 if i == 0 else 512, 512, kernel_size=5, stride=1, padding=2, dilation=1,
                                   w_init_gain="relu"), nn.BatchNorm1d(512)) for i in range(3))
        self.enc_lstm = nn.LSTM(dim_pre, dim_neck, 2, batch_first=True, bidirectional=True)
        self.enc_linear = LinearNorm(8192, 2048)
        self.style = LinearNorm(2048, self.speaker_size * 2)
        self.content = LinearNorm(2048, (latent_dim - self.speaker_size) * 2)
        self.dec_pre_linear1 = nn.Linear(latent_dim, 2048)
        self.dec_pre_linear2 = nn.Linear(2048, 8192)
        self.dec_lstm1 = nn.LSTM(dim_neck * 2, 512, 1, batch_first=True)
        self.dec_modules = nn.ModuleList(
            nn.Sequential(nn.Conv1d(dim_pre, dim_pre, kernel_size=5, stride=1, padding=2, dilation=1),
                          nn.BatchNorm1d(dim_pre)) for _ in range(3))
        self.dec_lstm2 = nn.LSTM(dim_pre, 1024, 2, batch_first=True)
        self.dec_linear2 = LinearNorm(1024, 80)
        self.apply(init_weights)

        self.postnet.__dict__["_owner"] = self  # plain attribute (not a registered sub-module): no state_dict recursion
        self._dt = _precision_tag(precision)
        self._engine = Engine(self._dt, latent_dim, speaker_size, grad_scale=default_grad_scale(self._dt, batch_size))
        self._param_names = [n for n, _ in self.named_parameters()]
        self._prep_cache = None
        self.noise_hook = None
        self._debug_keep_saved = False
        self._last_saved = None

    # ------------------------------------------------------------------ plumbing
    @property
    def grad_scale(self) -> float:
        return self._engine.grad_scale

    @grad_scale.setter
    def grad_scale(self, v: float) -> None:
        self._engine.grad_scale = float(v)

    def _param_dict(self):
        return {n: p.data for n, p in self.named_parameters()}

    def _buffer_dict(self):
        return dict(self.named_buffers())

    def _prepared(self) -> PreparedWeights:
        """Tensor-core copies of the parameters, re-derived whenever a parameter has been written (`_version` moves with
        every in-place update: torch optimizers bump it themselves, dvae_b200.optim.Adam bumps it explicitly because its
        kernel writes through raw pointers) or re-allocated."""
        params = list(self.parameters())
        ptrs = tuple(p.data_ptr() for p in params) + (self._dt,)
        vers = tuple(p._version for p in params)
        c = self._prep_cache
        if c is None or c[0] != ptrs:
            if params[0].device.type != "cuda":
                raise RuntimeError("dvae_b200 runs on CUDA (sm_100a) only: move the module with .to('cuda'); "
                                   "there is no CPU path")
            c = self._prep_cache = [ptrs, vers, PreparedWeights(self._dt, self._param_dict())]
        elif c[1] != vers:
            c[2].refresh(self._param_dict())
            c[1] = vers
        return c[2]

    def _noise(self, shape) -> torch.Tensor:
        """ε ~ N(0,1), drawn like the reference on the CPU default generator (:224) unless noise_hook is set."""
        dev = next(self.parameters()).device
        if self.noise_hook is not None:
            e = self.noise_hook(tuple(shape))
        else:
            # pinned staging (torch's caching host allocator) + non_blocking copy: a copy from pageable memory would make
            # the host wait for the previous step to drain before it can start enqueueing this one
            if dev.type == "cuda":
                e = torch.empty(tuple(shape), pin_memory=True).normal_()
                return e.to(device=dev, non_blocking=True)
            e = torch.empty(tuple(shape)).normal_()
        return e.to(device=dev, dtype=torch.float32).contiguous()

    @staticmethod
    def _mel(x: torch.Tensor) -> torch.Tensor:
        return x.detach().to(torch.float32).contiguous()

    # ------------------------------------------------------------------ reference API
    def encode(self, x):
        """:198-220 -> (style_mu, style_logvar, content_mu, content_logvar); gradient-free (used by conversion)."""
        W, P, B = self._prepared(), self._param_dict(), self._buffer_dict()
        x = self._mel(x)
        x_cl = torch.empty((x.shape[0], x.shape[2], x.shape[1]), device=x.device, dtype=ops.act_dtype(self._dt))
        ops.pack_ncl_to_cl(self._dt, x, x_cl)
        heads, _ = self._engine.encode_rows(W, P, B, x_cl, 1, self.training, None)
        S, L = self.speaker_size, self.latent_dim
        style, content = heads[:, :2 * S], heads[:, 2 * S:]
        return style[:, :S], style[:, S:], content[:, :L - S], content[:, L - S:]

    def _reparameterize(self, mu, logvar, train=True):
        """:222-228."""
        if not train:
            return mu
        eps = self._noise(logvar.shape)
        mu32, lv32 = mu.detach().float().contiguous(), logvar.detach().float().contiguous()
        gid = torch.arange(mu32.shape[0], device=mu32.device, dtype=torch.int32)
        return ops.group_reparam(mu32, lv32, gid, eps)

    def decode(self, z):
        """:230-248, z [N, latent_dim] -> [N, 80, 64]; gradient-free (used by conversion)."""
        W, P, B = self._prepared(), self._param_dict(), self._buffer_dict()
        z32 = z.detach().float().contiguous()
        z_act = torch.empty(z32.shape, device=z32.device, dtype=ops.act_dtype(self._dt))
        ops.prep_cast(self._dt, z32, z_act)
        _, rec32 = self._engine.decode_rows(W, P, B, z_act, 1, self.training, None)
        out, _ = ops.unpack_cl_to_ncl(self._dt, rec32, None)
        return out

    def decode_with_postnet(self, z):
        """decode(z) and decode(z) + postnet(decode(z)) in one pass (conversion tail, variational_base_vae.py:287-293)."""
        W, P, B = self._prepared(), self._param_dict(), self._buffer_dict()
        z32 = z.detach().float().contiguous()
        z_act = torch.empty(z32.shape, device=z32.device, dtype=ops.act_dtype(self._dt))
        ops.prep_cast(self._dt, z32, z_act)
        rec, rec32 = self._engine.decode_rows(W, P, B, z_act, 1, self.training, None)
        post = self._engine.postnet_rows(W, P, B, rec, 1, self.training, None)
        return ops.unpack_cl_to_ncl(self._dt, rec32, post)

    def _postnet_forward(self, x):
        W, P, B = self._prepared(), self._param_dict(), self._buffer_dict()
        x = self._mel(x)
        x_cl = torch.empty((x.shape[0], x.shape[2], x.shape[1]), device=x.device, dtype=ops.act_dtype(self._dt))
        ops.pack_ncl_to_cl(self._dt, x, x_cl)
        post = self._engine.postnet_rows(W, P, B, x_cl, 1, self.training, None)
        out, _ = ops.unpack_cl_to_ncl(self._dt, post, None)
        return out

    def forward(self, x1, x2, train=True):
        """:250-279 -> (recons_x1, recons_x2, recons_x1_hat, recons_x2_hat, q_z1_mu, q_z1_logvar, q_z2_mu,
        q_z2_logvar, z_style_mu, z_style_logvar)."""
        x1, x2 = self._mel(x1), self._mel(x2)
        R, S, L = x1.shape[0], self.speaker_size, self.latent_dim
        e1 = self._noise((R, L - S)) if train else None        # draw order of the reference: content1, content2, style
        e2 = self._noise((R, L - S)) if train else None
        e3 = self._noise((R, S))                               # style noise is always drawn (:261)
        params = list(self.parameters())
        keep = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if keep and not self.training:
            raise NotImplementedError("gradients in eval() mode (running-stat BatchNorm) are not implemented")
        return _NetworkFn.apply(self, keep, bool(train), x1, x2, e1, e2, e3, *params)

    def update_c(self):
        self._c += self._c_delta

    def update_beta(self):
        self._beta += self._beta_delta


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, batch_size, mse_cof, kl_cof, *t):
        ctx.set_materialize_grads(False)
        t = tuple(x.detach().float().contiguous() for x in t)
        out = ops.loss_fwd(*t, batch_size, mse_cof, kl_cof)
        ctx.t, ctx.cfg = t, (batch_size, mse_cof, kl_cof)
        return tuple(out[i] for i in range(8))

    @staticmethod
    def backward(ctx, *g):
        dev = ctx.t[0].device
        gout = torch.zeros(8, device=dev, dtype=torch.float32)
        for i, gi in enumerate(g):
            if gi is not None:
                gout[i:i + 1].copy_(gi.reshape(1))
        d = ops.loss_bwd(*ctx.t, *ctx.cfg, gout)
        return (None, None, None, None, None) + tuple(d)   # no gradient for batch_size/cofs/x1/x2


class ConvolutionalMulVAE(VariationalBaseModelVAE):
    """Loss + optimizer owner (model/disentangled_vae.py:288-350)."""

    def __init__(self, dataset, width, height, latent_sz, learning_rate, alpha, log_interval, normalize, batch_size,
                 speaker_size, channels=1, device=torch.device("cuda"), latent_dim=256, beta=0.1, mse_cof=10, kl_cof=10,
                 style_cof=0.1):
        super().__init__(dataset, width, height, channels, latent_sz, learning_rate, device, log_interval, batch_size)
        self.batch_size = batch_size
        self.alpha = alpha
        self.lr = learning_rate
        self.latent_dim = latent_dim
        self.mse_cof = mse_cof
        self.kl_cof = kl_cof
        self.style_cof = style_cof
        self.model = DisentangledVAE(latent_dim=self.latent_dim, beta=0.1, batch_size=batch_size,
                                     speaker_size=speaker_size).to(device)
        self.optimizer = fused_optim.Adam(self.model.parameters(), lr=self.lr)   # torch.optim.Adam semantics, one launch
        self.optimizer.overflow_hook = self._on_gradient_overflow
        self.train_losses = []
        self.test_losses = []

    def loss_functionGVAE2(self, x1, x2, x_recon1, x_recon2, recons_x1_hat, recons_x2_hat, q_z1_mu, q_z1_logvar, q_z2_mu,
                           q_z2_logvar, style_mu1, style_logvar1, train=False):
        """:310-327 -> (LOSS, MSE_x1, MSE_x2, MSE_x1_hat, MSE_x2_hat, z1_kl_loss, z2_kl_loss, z_kl_style); one fused
        kernel forward, one backward.  L1 sums / constructor batch_size; style KL reported only."""
        return _LossFn.apply(float(self.batch_size), float(self.mse_cof), float(self.kl_cof), x1, x2, x_recon1, x_recon2,
                             recons_x1_hat, recons_x2_hat, q_z1_mu, q_z1_logvar, q_z2_mu, q_z2_logvar, style_mu1,
                             style_logvar1)

    def _on_gradient_overflow(self):
        """fp16 mode: a gradient left fp16's range (the optimizer skipped those elements): halve the gradient scale."""
        if self.model._dt == lib.F16 and self.model.grad_scale > 2.0 ** -24:
            self.model.grad_scale = self.model.grad_scale / 2
            print(f"[dvae_b200] non-finite gradient: gradient scale lowered to {self.model.grad_scale:g}")

    def update_(self):
        self.model.update_c()
        self.model.update_beta()

    def compute_KL_delta_VAE(self, mu, logvar, alpha=0.95):
        """delta-VAE AR(1)-prior KL (:334-345).  Dead code in the reference (never called); kept for API parity as a
        plain torch expression -- it is not on the hot path."""
        a2 = alpha * alpha
        f = lambda x: x - torch.log(x) - 1
        kl = f(logvar[:, 0].exp()) + mu[:, 0].pow(2)
        for j in range(1, mu.shape[1]):
            kl = kl + f(logvar[:, j].exp() / (1 - a2))
            kl = kl + ((mu[:, j] - alpha * mu[:, j - 1]).pow(2) + a2 * logvar[:, j - 1]) / (1 - a2)
        return (-0.5) * torch.sum(kl)

    def update_kl(self):
        self.kl_cof = min(self.kl_cof * 2, 10)

    def set_kl(self, beta):
        self.kl = beta


def f_function(x, coef=1):
    return coef * x - torch.log(x) - 1
