"""Hand-scheduled forward / backward of the Disentangled-VAE network on the dvae_b200 kernels.

The engine owns no parameters: it reads the fp32 master parameters of the drop-in nn.Module, re-lays them for the
tensor cores (bf16 or tf32 storage, conv taps / LSTM gate interleave), runs both pair members (x1-call and x2-call
of model/disentangled_vae.py:250-279) as ONE batch of 2R rows -- BatchNorm statistics stay per call ("halves",
SURVEY F5) -- and keeps exactly the intermediates the backward schedule needs.

Data layout in HBM (R2 = rows in flight, T = 64 frames):
  activations  channels-last [R2, T, C] in the storage dtype (bf16 | fp32-as-tf32): the K-major A operand of every
               forward GEMM and, read MN-major, the operand of every weight-gradient GEMM (no transposes anywhere)
  conv weights [Cout, 5, Cin]; linear weights [N, K]; LSTM weights twice: gate-interleaved (forward) + natural (backward)
  LSTM state   h [R2, T, D*H] storage dtype, c [R2, T, D*H] fp32, activated gates [R2, T, D*4H] storage dtype
  gradients    fp32, accumulated with split-K reductions, returned in the parameters' own layouts
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from . import lib, ops

Tensor = torch.Tensor
T_FRAMES = 64
N_MELS = 80

ENC_CONVS = [("enc_modules.%d.0.conv" % i, "enc_modules.%d.1" % i, lib.ACT_RELU) for i in range(3)]
DEC_CONVS = [("dec_modules.%d.0" % i, "dec_modules.%d.1" % i, lib.ACT_RELU) for i in range(3)]
POST_CONVS = [("postnet.convolutions.%d.0.conv" % i, "postnet.convolutions.%d.1" % i,
               lib.ACT_TANH if i < 4 else lib.ACT_NONE) for i in range(5)]
# (prefix, layers, directions, hidden)
LSTMS = {"enc_lstm": (2, 2, 64), "dec_lstm1": (1, 1, 512), "dec_lstm2": (2, 1, 1024)}
LINEARS = ["enc_linear.linear_layer", "dec_pre_linear1", "dec_pre_linear2", "dec_linear2.linear_layer"]


class PreparedWeights:
    """Tensor-core copies of the fp32 parameters (rebuilt whenever the parameters change).

    The defaults describe DisentangledVAE; other networks built from the same layer kinds (the AutoVC replica) pass
    their own lists of conv / linear / LSTM parameter prefixes."""

    def __init__(self, dt: int, P: Dict[str, Tensor], convs=None, linears=None, lstms=None, fused_heads: bool = True):
        ad = ops.act_dtype(dt)
        self.dt = dt
        self.conv: Dict[str, Tensor] = {}
        self.lin: Dict[str, Tensor] = {}
        self.lstm: Dict[str, dict] = {}
        self._convs = [c for c, _, _ in ENC_CONVS + DEC_CONVS + POST_CONVS] if convs is None else convs
        self._linears = LINEARS if linears is None else linears
        self._lstms = LSTMS if lstms is None else lstms
        self._fused_heads = fused_heads
        # fp16 mode, DisentangledVAE only: split-precision operands (hi + lo, two 11-bit halves) for the layers whose cost is
        # negligible -- the first encoder convolution (K = 80 per tap), enc_linear (M = rows) and the heads -- so that the
        # input mel, those weights and `e` enter at ~22 bits instead of 11 (DESIGN.md "Numerics": the error budget)
        self.split = (dt == lib.F16 and fused_heads and convs is None
                      and os.environ.get("DVAE_B200_SPLIT", "1") == "1")
        dev = next(iter(P.values())).device
        for conv in self._convs:
            Co, Ci, _ = P[conv + ".weight"].shape
            self.conv[conv] = torch.empty((Co, 5, Ci), device=dev, dtype=ad)
        for name in self._linears:
            self.lin[name] = torch.empty_like(P[name + ".weight"], dtype=ad)   # bf16 / fp16 copy, or fp32 on the tf32 grid
        if fused_heads:
            # style + content heads fused into one [2L, 2048] GEMM (rows: style_mu, style_logvar, content_mu, content_logvar)
            ws, wc = P["style.linear_layer.weight"], P["content.linear_layer.weight"]
            n_s, n_c = ws.shape[0], wc.shape[0]
            self.heads_w = torch.empty((n_s + n_c, ws.shape[1]), device=dev, dtype=ad)
            self.heads_b = torch.empty((n_s + n_c,), device=dev, dtype=torch.float32)
            self.n_style = n_s
        if self.split:
            c0 = ENC_CONVS[0][0]
            Co, Ci, _ = P[c0 + ".weight"].shape
            self.conv0_cat = torch.empty((Co, 5, 3 * Ci), device=dev, dtype=ad)            # [w_hi | w_hi | w_lo] per tap
            wl = P["enc_linear.linear_layer.weight"]
            self.enc_linear_split = torch.empty((wl.shape[0], 2 * wl.shape[1]), device=dev, dtype=ad)    # [w_hi | w_lo]
            self.heads_w3 = torch.empty((self.heads_w.shape[0], 3 * self.heads_w.shape[1]), device=dev, dtype=ad)
            self._heads_w32 = torch.empty(self.heads_w.shape, device=dev, dtype=torch.float32)
        for prefix, (layers, D, H) in self._lstms.items():
            per_layer = []
            for l in range(layers):
                In = P[f"{prefix}.weight_ih_l{l}"].shape[1]
                per_layer.append(dict(wih_p=torch.empty((D * 4 * H, In), device=dev, dtype=ad),
                                      wih_n=torch.empty((D * 4 * H, In), device=dev, dtype=ad),
                                      whh_p=torch.empty((D, 4 * H, H), device=dev, dtype=ad),
                                      whh_n=torch.empty((D, 4 * H, H), device=dev, dtype=ad),
                                      bias_p=torch.empty((D * 4 * H,), device=dev, dtype=torch.float32), In=In))
            self.lstm[prefix] = dict(layers=per_layer, D=D, H=H)
        self.refresh(P)

    def _table(self, P: Dict[str, Tensor]):
        """The refresh as a table for ops.PrepTable: one descriptor per fp32 source tensor."""
        E = []
        for conv in self._convs:
            w = P[conv + ".weight"]
            cat = self.conv0_cat if (self.split and conv == ENC_CONVS[0][0]) else None
            E.append((ops.PREP_CONV, w, None, self.conv[conv], cat, w.shape[0], w.shape[1]))
        for name in self._linears:
            w = P[name + ".weight"]
            sp = self.enc_linear_split if (self.split and name == "enc_linear.linear_layer") else None
            E.append((ops.PREP_CAST, w, None, self.lin[name], sp, 2, w.shape[1]))
        if self.split and "enc_linear.linear_layer" not in self._linears:
            return None
        if self._fused_heads:
            n_s = self.n_style
            for w, b, rows in ((P["style.linear_layer.weight"], P["style.linear_layer.bias"], slice(0, n_s)),
                               (P["content.linear_layer.weight"], P["content.linear_layer.bias"], slice(n_s, None))):
                E.append((ops.PREP_CAST, w, None, self.heads_w[rows], self.heads_w3[rows] if self.split else None, 3, w.shape[1]))
                E.append((ops.PREP_COPY, b, None, self.heads_b[rows], None, 0, 0))
        for prefix, info in self.lstm.items():
            D, H = info["D"], info["H"]
            for l, lw in enumerate(info["layers"]):
                for d in range(D):
                    suf = "_reverse" if d == 1 else ""
                    sl = slice(d * 4 * H, (d + 1) * 4 * H)
                    w_ih, w_hh = P[f"{prefix}.weight_ih_l{l}{suf}"], P[f"{prefix}.weight_hh_l{l}{suf}"]
                    E.append((ops.PREP_LSTM_W, w_ih, None, lw["wih_n"][sl], lw["wih_p"][sl], H, w_ih.shape[1]))
                    E.append((ops.PREP_LSTM_W, w_hh, None, lw["whh_n"][d], lw["whh_p"][d], H, H))
                    E.append((ops.PREP_LSTM_BIAS, P[f"{prefix}.bias_ih_l{l}{suf}"], P[f"{prefix}.bias_hh_l{l}{suf}"], lw["bias_p"][sl],
                              None, H, 0))
        if not all(e[1].is_contiguous() and e[3].is_contiguous() for e in E):
            return None
        return E

    def refresh(self, P: Dict[str, Tensor]) -> None:
        """Re-derive every tensor-core copy from the fp32 masters, in place (after an optimizer step / load_state_dict).
        One launch (dvae_prep_all, driven by a table that is rebuilt only when a tensor has moved); DVAE_B200_PREP_ALL=0 or the
        strict-fp32 mode keep the tensor-by-tensor path (~55 launches), which the table path is tested against bit for bit."""
        dt = self.dt
        if dt in (lib.BF16, lib.F16, lib.TF32) and os.environ.get("DVAE_B200_PREP_ALL", "1") == "1":
            tab = getattr(self, "_prep_table", None)
            if tab is None or not tab.valid() or self._prep_params != tuple(P[k].data_ptr() for k in sorted(P)):
                E = self._table(P)
                tab = ops.PrepTable(dt, E, next(iter(P.values())).device) if E is not None else False
                self._prep_table = tab
                self._prep_params = tuple(P[k].data_ptr() for k in sorted(P))
            if tab:
                tab.run()
                return
        for conv in self._convs:
            cat = self.conv0_cat if (self.split and conv == ENC_CONVS[0][0]) else None
            ops.prep_conv_weight(dt, P[conv + ".weight"], out=self.conv[conv], out_cat=cat)
        if self.split:
            ops.prep_cast_split(dt, P["enc_linear.linear_layer.weight"], self.enc_linear_split, 2)
            n_s = self.n_style
            ops.copy_f32(P["style.linear_layer.weight"], self._heads_w32[:n_s])
            ops.copy_f32(P["content.linear_layer.weight"], self._heads_w32[n_s:])
            ops.prep_cast_split(dt, self._heads_w32, self.heads_w3, 3)
        for name in self._linears:
            ops.prep_cast(dt, P[name + ".weight"], self.lin[name])
        if self._fused_heads:
            n_s = self.n_style
            ops.prep_cast(dt, P["style.linear_layer.weight"], self.heads_w[:n_s])
            ops.prep_cast(dt, P["content.linear_layer.weight"], self.heads_w[n_s:])
            ops.copy_f32(P["style.linear_layer.bias"], self.heads_b[:n_s])
            ops.copy_f32(P["content.linear_layer.bias"], self.heads_b[n_s:])
        for prefix, info in self.lstm.items():
            D, H = info["D"], info["H"]
            tile = lib.lstm_gate_tile(H)
            for l, lw in enumerate(info["layers"]):
                for d in range(D):
                    suf = "_reverse" if d == 1 else ""
                    sl = slice(d * 4 * H, (d + 1) * 4 * H)
                    w_ih, w_hh = P[f"{prefix}.weight_ih_l{l}{suf}"], P[f"{prefix}.weight_hh_l{l}{suf}"]
                    ops.prep_lstm_weight(dt, w_ih, lw["wih_p"][sl], H, tile)
                    ops.prep_lstm_weight(dt, w_hh, lw["whh_p"][d], H, tile)
                    ops.prep_cast(dt, w_ih, lw["wih_n"][sl])
                    ops.prep_cast(dt, w_hh, lw["whh_n"][d])
                    ops.prep_lstm_bias(P[f"{prefix}.bias_ih_l{l}{suf}"], P[f"{prefix}.bias_hh_l{l}{suf}"], lw["bias_p"][sl], H, tile)


class GradSink:
    """Where parameter gradients land: views into the flat all-reduce buckets of dvae_b200.parallel.GradBuckets (then
    `done()` may trigger that bucket's overlapped all-reduce), or -- single process -- views into ONE flat fp32 buffer that
    is zeroed with a single fill.  The layout of that buffer (and of the zero-initialised scratch the backward needs:
    tap-major conv weight gradients, per-direction LSTM gradients) is learned from the first backward, which still
    allocates tensor by tensor; every later backward does two fills instead of ~100."""

    def __init__(self, device, buckets=None, layout: Optional[dict] = None):
        self.device = device
        self.buckets = buckets
        self.grads: Dict[str, Tensor] = {}
        self.layout = layout
        self.requests: List[tuple] = []          # (key, shape) in request order: what the next layout is built from
        self._flat = {}
        if layout is not None:
            for kind in ("grad", "scratch"):
                n = layout[kind]["total"]
                if n > 0 and not (kind == "grad" and buckets is not None):
                    self._flat[kind] = torch.zeros((n,), device=device, dtype=torch.float32)
        if buckets is not None:
            buckets.begin()

    def _zeros(self, kind: str, key: str, shape) -> Tensor:
        shape = tuple(int(d) for d in shape)
        self.requests.append((kind, key, shape))
        flat = self._flat.get(kind)
        slot = self.layout[kind]["slots"].get(key) if (flat is not None) else None
        if slot is not None and slot[1] == shape:
            n = 1
            for d in shape:
                n *= d
            return flat[slot[0]:slot[0] + n].view(shape)
        return torch.zeros(shape, device=self.device, dtype=torch.float32)

    def buf(self, name: str, shape) -> Tensor:
        """Zero-initialised fp32 buffer for the gradient of `name`."""
        t = self.buckets.view(name) if self.buckets is not None else self._zeros("grad", name, shape)
        self.grads[name] = t
        return t

    def scratch(self, key: str, shape) -> Tensor:
        """Zero-initialised fp32 scratch (same life time as the gradients of this backward)."""
        return self._zeros("scratch", key, shape)

    def done(self, name: str) -> None:
        if self.buckets is not None:
            self.buckets.ready(name)

    def put(self, name: str, src: Tensor) -> None:
        """Gradient computed elsewhere (a slice of a fused GEMM output): copy into place when bucketed."""
        if self.buckets is None:
            self.grads[name] = src
        else:
            ops.copy_f32(src.contiguous(), self.buf(name, src.shape))
        self.done(name)

    def finish(self) -> Dict[str, Tensor]:
        if self.buckets is not None:
            self.buckets.finish()
        return self.grads

    def next_layout(self) -> dict:
        """Flat-buffer layout for the following backward passes: 256-byte aligned slots in request order."""
        out = {"grad": {"slots": {}, "total": 0}, "scratch": {"slots": {}, "total": 0}}
        for kind, key, shape in self.requests:
            n = 1
            for d in shape:
                n *= d
            L = out[kind]
            if key not in L["slots"]:
                L["slots"][key] = (L["total"], shape)
                L["total"] += (n + 63) // 64 * 64
        return out


class Engine:
    def __init__(self, dt: int, latent_dim: int, speaker_size: int, bn_eps: float = 1e-5, bn_momentum: float = 0.1,
                 grad_scale: float = 1.0):
        # The activation-gradient stream is carried multiplied by `grad_scale` (a power of two, so scaling is exact) and
        # every parameter gradient is multiplied by 1 / grad_scale where it is produced.  1.0 for bf16 / tf32 storage
        # (fp32's exponent range); the fp16 mode needs it to keep small gradients inside fp16's normal range.
        self.grad_scale = float(grad_scale)
        # fp16 / tf32 modes: the convolution outputs that feed a train-mode BatchNorm are kept as unrounded fp32 (the rounding
        # of that tensor is the largest single term of the forward error budget, scripts/rounding_budget.py; it feeds
        # BatchNorm, not a tensor-core operand, so nothing requires the narrow grid); DVAE_B200_Y_F32=0 switches it off
        self.y_f32 = os.environ.get("DVAE_B200_Y_F32", "1") == "1" and dt in (lib.F16, lib.TF32)
        self.grad_stats = [] if os.environ.get("DVAE_DEBUG_GRAD_STATS") else None   # diagnostics: (name, amax) of the stream
        self.buckets = None   # set to a parallel.GradBuckets for data-parallel training
        self._sink_layout: Optional[dict] = None   # flat gradient / scratch layout learned from the first backward (GradSink)
        self.side_stream = None     # created lazily; weight-gradient GEMMs that are off the critical path run here
        self.use_side_stream = os.environ.get("DVAE_SIDE_STREAM", "1") != "0"   # A/B switch for profiling
        self._keepalive: List[tuple] = []
        self.dt = dt
        self.L = latent_dim
        self.S = speaker_size
        self.eps = bn_eps
        self.momentum = bn_momentum

    # ------------------------------------------------------------------ building blocks (forward)
    def _conv_stack(self, W: PreparedWeights, P, B, h: Tensor, convs, halves: int, training: bool, saved: Optional[list],
                    first_cat: Optional[Tensor] = None):
        """first_cat: split-precision copy [rows, T, 3C] = [hi | lo | hi] of the stack's input; the first convolution then
        contracts it with W.conv0_cat = [w_hi | w_hi | w_lo] (forward only: the backward uses the plain operands)."""
        dt = self.dt
        for li, (conv, bn, act) in enumerate(convs):
            x_in = h
            if training:
                # the convolution's epilogue also produces the BatchNorm statistics of its output (no extra pass over y)
                if li == 0 and first_cat is not None:
                    y, ws = ops.conv5_fwd_bnstats(dt, first_cat, W.conv0_cat, P[conv + ".bias"], halves, y_f32=self.y_f32)
                else:
                    y, ws = ops.conv5_fwd_bnstats(dt, x_in, W.conv[conv], P[conv + ".bias"], halves, y_f32=self.y_f32)
                C = y.shape[-1]
                h, stat = ops.bn_finalize_apply(dt, y.view(-1, C), ws, P[bn + ".weight"], P[bn + ".bias"],
                                                B[bn + ".running_mean"], B[bn + ".running_var"],
                                                B[bn + ".num_batches_tracked"], halves, act, self.eps, self.momentum)
                h = h.view_as(y)
                if saved is not None:
                    saved.append(dict(conv=conv, bn=bn, act=act, x_in=x_in, y=y, stat=stat))
            else:
                y = ops.conv5_fwd(dt, x_in, W.conv[conv], P[conv + ".bias"])
                C = y.shape[-1]
                h = ops.bn_eval_fwd(dt, y.view(-1, C), P[bn + ".weight"], P[bn + ".bias"], B[bn + ".running_mean"],
                                    B[bn + ".running_var"], act, self.eps).view_as(y)
        return h

    def _lstm(self, W: PreparedWeights, prefix: str, x: Tensor, saved: Optional[list]) -> Tensor:
        dt = self.dt
        info = W.lstm[prefix]
        D, H = info["D"], info["H"]
        rows, T, _ = x.shape
        for lw in info["layers"]:
            In = lw["In"]
            xg, _ = ops.linear_fwd(dt, x.reshape(rows * T, In), lw["wih_p"], lw["bias_p"])
            xg = xg.view(rows, T, D * 4 * H)
            h_all, c_all = ops.lstm_fwd(dt, xg, lw["whh_p"], H, D)
            if saved is not None:
                saved.append(dict(prefix=prefix, x_in=x, gates=xg, h_all=h_all, c_all=c_all, lw=lw, D=D, H=H))
            x = h_all
        return x

    def encode_rows(self, W, P, B, x_cl: Tensor, halves: int, training: bool, saved: Optional[dict],
                    x_cat: Optional[Tensor] = None):
        """x_cl [R2, T, 80] act -> (heads fp32 [R2, 2L], e act [R2, 2048]).  x_cat: split-precision copy of x_cl (fp16
        training step only, see PreparedWeights.split)."""
        dt = self.dt
        R2 = x_cl.shape[0]
        sv_conv = [] if saved is not None else None
        sv_lstm = [] if saved is not None else None
        split = W.split and training and x_cat is not None
        h = self._conv_stack(W, P, B, x_cl, ENC_CONVS, halves, training, sv_conv, first_cat=x_cat if split else None)
        h = self._lstm(W, "enc_lstm", h, sv_lstm)                      # [R2, T, 128]
        flat = h.view(R2, T_FRAMES * 128)
        if split:
            e, e_cat = ops.linear_fwd_split(dt, flat, W.enc_linear_split, P["enc_linear.linear_layer.bias"], relu=True, want_cat=True)
            _, heads = ops.linear_fwd(dt, e_cat, W.heads_w3, W.heads_b, want_f32=True, want_act=False)
        else:
            e, _ = ops.linear_fwd(dt, flat, W.lin["enc_linear.linear_layer"], P["enc_linear.linear_layer.bias"], relu=True)
            _, heads = ops.linear_fwd(dt, e, W.heads_w, W.heads_b, want_f32=True, want_act=False)
        if saved is not None:
            saved.update(enc_convs=sv_conv, enc_lstm=sv_lstm, flat=flat, e=e, heads=heads)
        return heads, e

    def decode_rows(self, W, P, B, z: Tensor, halves: int, training: bool, saved: Optional[dict]):
        """z act [R2, L] -> (rec act [R2, T, 80], rec32 fp32 [R2, T, 80])."""
        dt = self.dt
        R2 = z.shape[0]
        d1, _ = ops.linear_fwd(dt, z, W.lin["dec_pre_linear1"], P["dec_pre_linear1.bias"])
        d2, _ = ops.linear_fwd(dt, d1, W.lin["dec_pre_linear2"], P["dec_pre_linear2.bias"])
        sv_l1 = [] if saved is not None else None
        sv_conv = [] if saved is not None else None
        sv_l2 = [] if saved is not None else None
        h = self._lstm(W, "dec_lstm1", d2.view(R2, T_FRAMES, 128), sv_l1)      # [R2, T, 512]
        h = self._conv_stack(W, P, B, h, DEC_CONVS, halves, training, sv_conv)
        h = self._lstm(W, "dec_lstm2", h, sv_l2)                               # [R2, T, 1024]
        rec, rec32 = ops.linear_fwd(dt, h.view(R2 * T_FRAMES, 1024), W.lin["dec_linear2.linear_layer"],
                                    P["dec_linear2.linear_layer.bias"], want_f32=True)
        rec = rec.view(R2, T_FRAMES, N_MELS)
        rec32 = rec32.view(R2, T_FRAMES, N_MELS)
        if saved is not None:
            saved.update(z=z, d1=d1, d2=d2, dec_lstm1=sv_l1, dec_convs=sv_conv, dec_lstm2=sv_l2, h_top=h, rec=rec)
        return rec, rec32

    def postnet_rows(self, W, P, B, rec: Tensor, halves: int, training: bool, saved: Optional[dict]) -> Tensor:
        sv = [] if saved is not None else None
        out = self._conv_stack(W, P, B, rec, POST_CONVS, halves, training, sv)
        if saved is not None:
            saved["post_convs"] = sv
        return out

    # ------------------------------------------------------------------ full training-step forward
    def forward(self, W: PreparedWeights, P, B, x1: Tensor, x2: Tensor, eps: Sequence[Optional[Tensor]], training: bool,
                sample_content: bool, keep: bool):
        """Returns (outputs10, saved).  x1, x2 fp32 [R, 80, 64]."""
        dt = self.dt
        R = x1.shape[0]
        assert x1.shape == x2.shape and tuple(x1.shape[1:]) == (N_MELS, T_FRAMES), \
            f"expected [R, 80, 64] mel chunks, got {tuple(x1.shape)} (the network is locked to 64 frames)"
        saved: Optional[dict] = {} if keep else None
        x_cl = torch.empty((2 * R, T_FRAMES, N_MELS), device=x1.device, dtype=ops.act_dtype(dt))
        x_cat = torch.empty((2 * R, T_FRAMES, 3 * N_MELS), device=x1.device, dtype=ops.act_dtype(dt)) \
            if (W.split and training) else None
        ops.pack_ncl_to_cl(dt, x1, x_cl[:R], None if x_cat is None else x_cat[:R])
        ops.pack_ncl_to_cl(dt, x2, x_cl[R:], None if x_cat is None else x_cat[R:])
        heads, _ = self.encode_rows(W, P, B, x_cl, 2, training, saved, x_cat=x_cat)
        z, q, zs = ops.latent_tail_fwd(dt, heads, eps[0], eps[1], eps[2], R, self.L, self.S, sample_content)
        rec, rec32 = self.decode_rows(W, P, B, z, 2, training, saved)
        post = self.postnet_rows(W, P, B, rec, 2, training, saved)
        recon, hat = ops.unpack_cl_to_ncl(dt, rec32, post)
        if saved is not None:
            saved.update(R=R, eps=list(eps), sample_content=sample_content)
        outs = (recon[:R], recon[R:], hat[:R], hat[R:], q[0], q[1], q[2], q[3], zs[0], zs[1])
        return outs, saved

    def _side_stream_ctx(self):
        """Context that enqueues on the side stream, ordered after everything already enqueued on the current stream."""
        import contextlib
        if not self.use_side_stream:
            return contextlib.nullcontext()
        cur = torch.cuda.current_stream()
        if self.side_stream is None or self.side_stream.device != cur.device:
            self.side_stream = torch.cuda.Stream(device=cur.device)
        if self.buckets is not None:
            self.buckets.extra_streams = [self.side_stream]   # bucket all-reduces must also wait for side-stream gradients
        self.side_stream.wait_stream(cur)
        return torch.cuda.stream(self.side_stream)

    def _join_side_stream(self):
        if self.use_side_stream and self.side_stream is not None:
            torch.cuda.current_stream().wait_stream(self.side_stream)
        self._keepalive.clear()

    @staticmethod
    def discrete_decisions(saved: dict, outs, x1: Tensor, x2: Tensor) -> dict:
        """The branch decisions this forward took at the network's kinks: ReLU masks (reference layout [R, C, T] per
        call) and the signs of the L1 loss.  Parity tests hand them to the checker so that gradient comparisons are
        made at matched decisions (plain tensor bookkeeping, not part of the compute path)."""
        R = saved["R"]
        d = {}
        for group, key in (("enc_convs", "enc_modules"), ("dec_convs", "dec_modules")):
            for i, s in enumerate(saved[group]):
                y, st = s["y"].float(), s["stat"]
                for call in range(2):
                    z = y[call * R:(call + 1) * R] * st[call, 2] + st[call, 3]
                    d[f"{key}.{i}:{call}"] = (z > 0).transpose(1, 2)
        e = saved["e"].float()
        d["enc_linear:0"], d["enc_linear:1"] = e[:R] > 0, e[R:] > 0
        d["l1_signs"] = [torch.sign(outs[0] - x1), torch.sign(outs[1] - x2), torch.sign(outs[2] - x1),
                         torch.sign(outs[3] - x2)]
        return d

    # ------------------------------------------------------------------ building blocks (backward)
    def _stat(self, name: str, t: Tensor) -> None:
        """Diagnostics only (DVAE_DEBUG_GRAD_STATS=1): largest magnitude of a gradient-stream tensor, for choosing grad_scale."""
        if self.grad_stats is not None and t is not None:
            self.grad_stats.append((name, t.float().abs().max().item()))

    def _conv_stack_bwd(self, W, dout: Tensor, saved_layers: list, sink: GradSink, halves: int, need_dx: bool):
        dt = self.dt
        for i in range(len(saved_layers) - 1, -1, -1):
            s = saved_layers[i]
            y = s["y"]
            C = y.shape[-1]
            gname, bname = s["bn"] + ".weight", s["bn"] + ".bias"
            dy, _, _ = ops.bn_train_bwd(dt, dout.reshape(-1, C), y.view(-1, C), s["stat"], halves, s["act"],
                                        dgamma=sink.buf(gname, (C,)), dbeta=sink.buf(bname, (C,)), alpha=1.0 / self.grad_scale)
            dy = dy.view_as(y)
            self._stat("dy:" + s["conv"], dy)
            sink.done(gname), sink.done(bname)
            wk = W.conv[s["conv"]]
            Co, _, Ci = wk.shape
            # dX first: it is the critical path (it feeds the layer below).  The weight gradient then runs on the side stream,
            # ordered after the dgrad, so the tensor-core-bound wgrad GEMM of layer i overlaps the HBM-bound BatchNorm
            # backward of layer i-1 instead of delaying it.
            if i > 0 or need_dx:
                dout = ops.conv5_dgrad(dt, dy, wk)
            else:
                dout = None
            self._keepalive.append((dy, s["x_in"]))
            with self._side_stream_ctx():
                dwk = sink.scratch("dwk:" + s["conv"], wk.shape)
                ops.conv5_wgrad(dt, dy, s["x_in"], dwk, alpha=1.0 / self.grad_scale)
                ops.conv_wgrad_unpack(dwk, out=sink.buf(s["conv"] + ".weight", (Co, Ci, 5)))
                self._keepalive.append((dwk,))
            sink.done(s["conv"] + ".weight")
            # the conv bias feeds a train-mode BatchNorm: its gradient is identically zero (BN removes the mean)
            sink.buf(s["conv"] + ".bias", (C,))
            sink.done(s["conv"] + ".bias")
        return dout

    def _lstm_bwd(self, W, prefix: str, dh: Tensor, saved_layers: list, sink: GradSink, need_dx: bool):
        """Back-propagation through time, layer by layer.  Only dX (needed by the layer below) is on the critical path:
        the weight / bias gradient GEMMs of a layer are enqueued on a side stream so that they fill the SMs left idle by
        the next layer's (small, latency-bound) recurrence kernels."""
        dt = self.dt
        for l in range(len(saved_layers) - 1, -1, -1):
            s = saved_layers[l]
            D, H, lw = s["D"], s["H"], s["lw"]
            rows, T, _ = s["h_all"].shape
            In = lw["In"]
            da = ops.lstm_bwd(dt, dh.reshape(rows, T, D * H), s["gates"], s["c_all"], lw["whh_n"], H, D)
            da2 = da.view(rows * T, D * 4 * H)
            self._stat(f"da:{prefix}.l{l}", da)
            x2 = s["x_in"].reshape(rows * T, In)
            if l > 0 or need_dx:
                dh, _ = ops.linear_dgrad(dt, da2, lw["wih_n"])
                dh = dh.view(rows, T, In)
            else:
                dh = None
            self._keepalive.append((da, x2))
            with self._side_stream_ctx():
                lib.call("dvae_set_background", 1)
                try:
                    self._lstm_wgrads(prefix, l, s, da, da2, x2, sink)
                finally:
                    lib.call("dvae_set_background", 0)
        return dh

    def _lstm_wgrads(self, prefix: str, l: int, s: dict, da: Tensor, da2: Tensor, x2: Tensor, sink: GradSink):
        dt = self.dt
        D, H, lw = s["D"], s["H"], s["lw"]
        In = lw["In"]
        inv = 1.0 / self.grad_scale
        if D == 1:   # gradients land directly in their final buffers
            n_ih, n_hh = f"{prefix}.weight_ih_l{l}", f"{prefix}.weight_hh_l{l}"
            ops.linear_wgrad(dt, da2, x2, sink.buf(n_ih, (4 * H, In)), alpha=inv)
            sink.done(n_ih)
            ops.lstm_wgrad_hh(dt, da, s["h_all"], sink.buf(n_hh, (4 * H, H)).view(1, 4 * H, H), H, 1, alpha=inv)
            sink.done(n_hh)
            db = sink.buf(f"{prefix}.bias_ih_l{l}", (4 * H,))
            ops.colsum(dt, da2, db, alpha=inv)
            sink.done(f"{prefix}.bias_ih_l{l}")
            ops.copy_f32(db, sink.buf(f"{prefix}.bias_hh_l{l}", (4 * H,)))   # b_ih and b_hh: equal gradients
            sink.done(f"{prefix}.bias_hh_l{l}")
        else:        # both directions come out of one GEMM; slice per direction
            dwih = sink.scratch(f"dwih:{prefix}.{l}", (D * 4 * H, In))
            ops.linear_wgrad(dt, da2, x2, dwih, alpha=inv)
            dwhh = sink.scratch(f"dwhh:{prefix}.{l}", (D, 4 * H, H))
            ops.lstm_wgrad_hh(dt, da, s["h_all"], dwhh, H, D, alpha=inv)
            db = sink.scratch(f"db:{prefix}.{l}", (D * 4 * H,))
            ops.colsum(dt, da2, db, alpha=inv)
            for d in range(D):
                suf = "_reverse" if d == 1 else ""
                sl = slice(d * 4 * H, (d + 1) * 4 * H)
                sink.put(f"{prefix}.weight_ih_l{l}{suf}", dwih[sl])
                sink.put(f"{prefix}.weight_hh_l{l}{suf}", dwhh[d])
                sink.put(f"{prefix}.bias_ih_l{l}{suf}", db[sl])
                sink.put(f"{prefix}.bias_hh_l{l}{suf}", db[sl].clone())

    def _linear_bwd(self, name: str, W_act: Tensor, dy: Tensor, x: Tensor, sink: GradSink, relu_mask=None,
                    want_f32=False, need_dx: bool = True):
        dt = self.dt
        N, K = W_act.shape
        ops.linear_wgrad(dt, dy, x, sink.buf(name + ".weight", (N, K)), alpha=1.0 / self.grad_scale)
        sink.done(name + ".weight")
        ops.colsum(dt, dy, sink.buf(name + ".bias", (N,)), alpha=1.0 / self.grad_scale)
        sink.done(name + ".bias")
        if not need_dx:
            return None
        dx, dx32 = ops.linear_dgrad(dt, dy, W_act, relu_mask=relu_mask, want_f32=want_f32, want_act=not want_f32)
        return dx32 if want_f32 else dx

    # ------------------------------------------------------------------ full backward
    def backward(self, W: PreparedWeights, saved: dict, gouts: Sequence[Optional[Tensor]]) -> Dict[str, Tensor]:
        """gouts: gradients of the 10 forward outputs (None = zero).  Returns {param name: fp32 gradient}."""
        dt = self.dt
        ad = ops.act_dtype(dt)
        R = saved["R"]
        R2 = 2 * R
        dev = saved["heads"].device
        sink = GradSink(dev, self.buckets, self._sink_layout)
        grads = sink   # the helpers below take the sink
        g = [t.contiguous() if t is not None else None for t in gouts]
        # ---- residual output: recon_hat = recon + postnet(recon)  (:277-278)
        d_rec = torch.empty((R2, T_FRAMES, N_MELS), device=dev, dtype=ad)
        d_post = torch.empty((R2, T_FRAMES, N_MELS), device=dev, dtype=ad)
        gs = self.grad_scale
        ops.recon_out_bwd(dt, g[0], g[2], d_rec[:R], d_post[:R], scale=gs)
        ops.recon_out_bwd(dt, g[1], g[3], d_rec[R:], d_post[R:], scale=gs)
        self._stat("d_rec", d_rec), self._stat("d_post", d_post)
        d_in = self._conv_stack_bwd(W, d_post, saved["post_convs"], grads, 2, need_dx=True)
        ops.add_inplace(dt, d_rec, d_in)
        # ---- decoder
        d_rec2 = d_rec.view(R2 * T_FRAMES, N_MELS)
        dh = self._linear_bwd("dec_linear2.linear_layer", W.lin["dec_linear2.linear_layer"], d_rec2,
                              saved["h_top"].view(R2 * T_FRAMES, 1024), grads)
        dh = self._lstm_bwd(W, "dec_lstm2", dh, saved["dec_lstm2"], grads, need_dx=True)
        dh = self._conv_stack_bwd(W, dh, saved["dec_convs"], grads, 2, need_dx=True)
        dh = self._lstm_bwd(W, "dec_lstm1", dh, saved["dec_lstm1"], grads, need_dx=True)     # [R2, T, 128]
        d_d2 = dh.reshape(R2, T_FRAMES * 128)
        d_d1 = self._linear_bwd("dec_pre_linear2", W.lin["dec_pre_linear2"], d_d2, saved["d1"], grads)
        dz = self._linear_bwd("dec_pre_linear1", W.lin["dec_pre_linear1"], d_d1, saved["z"], grads, want_f32=True)
        # ---- latent tail (:252-272)
        eps = saved["eps"]
        dheads = ops.latent_tail_bwd(dt, saved["heads"], eps[0], eps[1], eps[2], dz, g[4:8], g[8:10], R, self.L, self.S,
                                     saved["sample_content"], gscale=gs)
        self._stat("dheads", dheads)
        # ---- encoder heads + linear
        n_s = W.n_style
        dw = sink.scratch("dw:heads", W.heads_w.shape)
        ops.linear_wgrad(dt, dheads, saved["e"], dw, alpha=1.0 / gs)
        db = sink.scratch("db:heads", (W.heads_w.shape[0],))
        ops.colsum(dt, dheads, db, alpha=1.0 / gs)
        sink.put("style.linear_layer.weight", dw[:n_s]), sink.put("content.linear_layer.weight", dw[n_s:])
        sink.put("style.linear_layer.bias", db[:n_s]), sink.put("content.linear_layer.bias", db[n_s:])
        d_e, _ = ops.linear_dgrad(dt, dheads, W.heads_w, relu_mask=saved["e"])
        d_flat = self._linear_bwd("enc_linear.linear_layer", W.lin["enc_linear.linear_layer"], d_e, saved["flat"], grads)
        dh = self._lstm_bwd(W, "enc_lstm", d_flat.view(R2, T_FRAMES, 128), saved["enc_lstm"], grads, need_dx=True)
        self._conv_stack_bwd(W, dh, saved["enc_convs"], grads, 2, need_dx=False)
        self._join_side_stream()
        if self._sink_layout is None:
            self._sink_layout = sink.next_layout()
        return sink.finish()
