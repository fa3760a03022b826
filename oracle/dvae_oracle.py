"""CPU oracle for the Disentangled-VAE hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain-PyTorch fp32 *functional* restatement of the reference algorithm
(reference = v-manhlt3/Disentangle-VAE-for-VC, paths relative to /root/reference).
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it; the product (disentangle-vae-for-vc_b200/) never does.

Parity pinning: the reference ships no tests / golden vectors, so the
oracle is pinned against the reference modules themselves, imported and executed in the
authoring container by `oracle/make_golden.py`; the resulting vectors are committed under
`tests/golden/` and `tests/test_oracle_golden.py` re-checks the oracle against them.

The network is expressed over a flat ``state_dict`` (the 84 parameters + 33 buffers of the
reference ``DisentangledVAE``) instead of an nn.Module tree, so one file shows the whole
data flow.  Layout conventions follow the reference: mel tensors are [R, 80, 64] (NCL).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

T_FRAMES = 64          # model/disentangled_vae.py:165,235 (8192 = 64 * 2*dim_neck) -- SURVEY F1
N_MELS = 80
DIM_NECK = 64
DIM_PRE = 512
BN_EPS = 1e-5          # torch.nn.BatchNorm1d default, model/disentangled_vae.py:159
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# Parameter inventory (model/disentangled_vae.py:126-195).  Used by tests and by the
# deterministic weight generator so that no 245 MB checkpoint has to be committed.
# --------------------------------------------------------------------------------------
def param_shapes(latent_dim: int = 32, speaker_size: int = 4) -> Dict[str, Tuple[int, ...]]:
    """All 84 parameter tensors, in the registration order of the reference module."""
    s: Dict[str, Tuple[int, ...]] = {}

    def bn(prefix: str, c: int):
        s[prefix + ".weight"] = (c,)
        s[prefix + ".bias"] = (c,)

    # Postnet (:43-79) is registered first (:146)
    chans = [(80, 512), (512, 512), (512, 512), (512, 512), (512, 80)]
    for i, (ci, co) in enumerate(chans):
        s[f"postnet.convolutions.{i}.0.conv.weight"] = (co, ci, 5)
        s[f"postnet.convolutions.{i}.0.conv.bias"] = (co,)
        bn(f"postnet.convolutions.{i}.1", co)
    # encoder convs (:151-162)
    for i in range(3):
        ci = 80 if i == 0 else 512
        s[f"enc_modules.{i}.0.conv.weight"] = (512, ci, 5)
        s[f"enc_modules.{i}.0.conv.bias"] = (512,)
        bn(f"enc_modules.{i}.1", 512)
    # enc_lstm (:163): 2 layers, bidirectional, H = 64
    for layer in range(2):
        inp = DIM_PRE if layer == 0 else 2 * DIM_NECK
        for suf in ("", "_reverse"):
            s[f"enc_lstm.weight_ih_l{layer}{suf}"] = (4 * DIM_NECK, inp)
            s[f"enc_lstm.weight_hh_l{layer}{suf}"] = (4 * DIM_NECK, DIM_NECK)
            s[f"enc_lstm.bias_ih_l{layer}{suf}"] = (4 * DIM_NECK,)
            s[f"enc_lstm.bias_hh_l{layer}{suf}"] = (4 * DIM_NECK,)
    s["enc_linear.linear_layer.weight"] = (2048, 8192)
    s["enc_linear.linear_layer.bias"] = (2048,)
    s["style.linear_layer.weight"] = (2 * speaker_size, 2048)
    s["style.linear_layer.bias"] = (2 * speaker_size,)
    s["content.linear_layer.weight"] = (2 * (latent_dim - speaker_size), 2048)
    s["content.linear_layer.bias"] = (2 * (latent_dim - speaker_size),)
    s["dec_pre_linear1.weight"] = (2048, latent_dim)
    s["dec_pre_linear1.bias"] = (2048,)
    s["dec_pre_linear2.weight"] = (8192, 2048)
    s["dec_pre_linear2.bias"] = (8192,)
    s["dec_lstm1.weight_ih_l0"] = (2048, 128)
    s["dec_lstm1.weight_hh_l0"] = (2048, 512)
    s["dec_lstm1.bias_ih_l0"] = (2048,)
    s["dec_lstm1.bias_hh_l0"] = (2048,)
    for i in range(3):
        s[f"dec_modules.{i}.0.weight"] = (512, 512, 5)
        s[f"dec_modules.{i}.0.bias"] = (512,)
        bn(f"dec_modules.{i}.1", 512)
    for layer in range(2):
        inp = 512 if layer == 0 else 1024
        s[f"dec_lstm2.weight_ih_l{layer}"] = (4096, inp)
        s[f"dec_lstm2.weight_hh_l{layer}"] = (4096, 1024)
        s[f"dec_lstm2.bias_ih_l{layer}"] = (4096,)
        s[f"dec_lstm2.bias_hh_l{layer}"] = (4096,)
    s["dec_linear2.linear_layer.weight"] = (80, 1024)
    s["dec_linear2.linear_layer.bias"] = (80,)
    return s


def bn_prefixes() -> List[Tuple[str, int]]:
    out = [(f"postnet.convolutions.{i}.1", 512 if i < 4 else 80) for i in range(5)]
    out += [(f"enc_modules.{i}.1", 512) for i in range(3)]
    out += [(f"dec_modules.{i}.1", 512) for i in range(3)]
    return out


def buffer_shapes() -> Dict[str, Tuple[int, ...]]:
    """The 33 BatchNorm buffers (running_mean, running_var, num_batches_tracked)."""
    s: Dict[str, Tuple[int, ...]] = {}
    for p, c in bn_prefixes():
        s[p + ".running_mean"] = (c,)
        s[p + ".running_var"] = (c,)
        s[p + ".num_batches_tracked"] = ()
    return s


def _key_rng(name: str, seed: int) -> np.random.Generator:
    import zlib
    return np.random.Generator(np.random.Philox(key=[zlib.crc32(name.encode()) & 0xFFFFFFFF, seed]))


def synth_state_dict(seed: int = 0, latent_dim: int = 32, speaker_size: int = 4,
                     randomize_bn: bool = True) -> SD:
    """Deterministic, torch-RNG-independent weights with reference-like magnitudes.

    The magnitudes mirror `init_weights` (model/disentangled_vae.py:26-32: xavier-uniform
    weights, bias 0.01 / 0) and torch's LSTM default U(-1/sqrt(H), 1/sqrt(H)); BN affine and
    running statistics are randomised (unlike a fresh reference model) so that every term of
    the BatchNorm arithmetic is exercised by the parity tests.
    """
    sd: SD = {}
    bn_names = {p for p, _ in bn_prefixes()}

    def add_buffers(prefix: str, c: int):
        # state_dict order of nn.BatchNorm1d: weight, bias, running_mean, running_var, num_batches_tracked
        g = _key_rng(prefix + ".running_mean", seed)
        arr = g.uniform(-0.1, 0.1, size=(c,)) if randomize_bn else np.zeros((c,))
        sd[prefix + ".running_mean"] = torch.from_numpy(arr.astype(np.float32))
        g = _key_rng(prefix + ".running_var", seed)
        arr = g.uniform(0.8, 1.2, size=(c,)) if randomize_bn else np.ones((c,))
        sd[prefix + ".running_var"] = torch.from_numpy(arr.astype(np.float32))
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    for name, shape in param_shapes(latent_dim, speaker_size).items():
        g = _key_rng(name, seed)
        is_bn = name.rsplit(".", 1)[0] in bn_names
        if "lstm" in name:
            hid = shape[0] // 4
            bound = 1.0 / math.sqrt(hid)
            arr = g.uniform(-bound, bound, size=shape)
        elif is_bn:
            if name.endswith(".weight"):
                arr = g.uniform(0.5, 1.5, size=shape) if randomize_bn else np.ones(shape)
            else:
                arr = g.uniform(-0.2, 0.2, size=shape) if randomize_bn else np.zeros(shape)
        elif len(shape) >= 2:
            rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
            fan_in, fan_out = shape[1] * rf, shape[0] * rf
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            arr = g.uniform(-bound, bound, size=shape)
        else:  # conv / linear bias: small but non-zero so the bias path is tested
            arr = g.uniform(-0.05, 0.05, size=shape)
        sd[name] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
        if is_bn and name.endswith(".bias"):
            add_buffers(name.rsplit(".", 1)[0], shape[0])
    return sd


def synth_inputs(R: int, seed: int = 1234, latent_dim: int = 32, speaker_size: int = 4):
    """x1, x2 ~ U[0,1) [R,80,64] and the three noise tensors (SURVEY 8d / F6 draw order)."""
    g = np.random.Generator(np.random.Philox(key=[seed, R]))
    f32 = lambda a: torch.from_numpy(a.astype(np.float32))
    x1 = f32(g.uniform(0, 1, size=(R, N_MELS, T_FRAMES)))
    x2 = f32(g.uniform(0, 1, size=(R, N_MELS, T_FRAMES)))
    eps = [f32(g.standard_normal((R, latent_dim - speaker_size))),
           f32(g.standard_normal((R, latent_dim - speaker_size))),
           f32(g.standard_normal((R, speaker_size)))]
    return x1, x2, eps


# --------------------------------------------------------------------------------------
# Layers
# --------------------------------------------------------------------------------------
def conv_bn(sd: SD, x: Tensor, conv_prefix: str, bn_prefix: str, training: bool) -> Tensor:
    """Conv1d(k=5, pad=2) followed by BatchNorm1d.

    Reference: ConvNorm.forward model/disentangled_vae.py:119-121 (or nn.Conv1d :178-189) then
    nn.BatchNorm1d (:159).  Train mode normalises with the biased batch variance over (R, T)
    and updates running stats in place (momentum 0.1, unbiased variance), exactly as torch.
    """
    y = F.conv1d(x, sd[conv_prefix + ".weight"], sd[conv_prefix + ".bias"], stride=1, padding=2)
    if training:
        sd[bn_prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(y, sd[bn_prefix + ".running_mean"], sd[bn_prefix + ".running_var"],
                        sd[bn_prefix + ".weight"], sd[bn_prefix + ".bias"],
                        training, BN_MOMENTUM, BN_EPS)


def relu_at(x: Tensor, decisions: Optional[dict], key: str) -> Tensor:
    """F.relu, or -- in matched-decision mode -- multiplication by a supplied 0/1 mask.

    ReLU (and the L1 loss) are kinks: two forwards that differ by rounding take different branches on a small fraction
    of elements, and the gradients then differ by O(sqrt(fraction)) no matter how good the arithmetic is.  Gradient
    parity tests therefore evaluate the oracle AT THE CANDIDATE'S discrete decisions (`decisions[key]`), which makes
    the oracle's backward a smooth function of the forward values."""
    if decisions is None or key not in decisions:
        return F.relu(x)
    return x * decisions[key].to(x.dtype)


def lstm(sd: SD, x: Tensor, prefix: str, num_layers: int, bidirectional: bool) -> Tensor:
    """batch_first LSTM with zero initial state (torch gate order i,f,g,o; two bias vectors).

    Reference call sites: model/disentangled_vae.py:208 (enc_lstm), :238 (dec_lstm1),
    :246 (dec_lstm2).  Uses the ATen builtin so the CPU baseline runs the same MKLDNN path
    as the reference's nn.LSTM.
    """
    R = x.shape[0]
    hid = sd[f"{prefix}.weight_hh_l0"].shape[1]
    dirs = 2 if bidirectional else 1
    flat = []
    for layer in range(num_layers):
        for suf in (("", "_reverse") if bidirectional else ("",)):
            for nm in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                flat.append(sd[f"{prefix}.{nm}_l{layer}{suf}"])
    h0 = x.new_zeros(num_layers * dirs, R, hid)
    # `train` only matters to cuDNN (its backward refuses inference-mode graphs); dropout is 0 either way
    out, _, _ = torch.lstm(x, (h0, h0.clone()), flat, True, num_layers, 0.0, torch.is_grad_enabled(), bidirectional, True)
    return out


def lstm_explicit(sd: SD, x: Tensor, prefix: str, num_layers: int, bidirectional: bool) -> Tensor:
    """Same recurrence written out step by step (cross-check of `lstm`, small sizes only)."""
    R, T, _ = x.shape
    inp = x
    for layer in range(num_layers):
        outs = []
        for suf in (("", "_reverse") if bidirectional else ("",)):
            w_ih = sd[f"{prefix}.weight_ih_l{layer}{suf}"]
            w_hh = sd[f"{prefix}.weight_hh_l{layer}{suf}"]
            b = sd[f"{prefix}.bias_ih_l{layer}{suf}"] + sd[f"{prefix}.bias_hh_l{layer}{suf}"]
            H = w_hh.shape[1]
            h = inp.new_zeros(R, H)
            c = inp.new_zeros(R, H)
            ys: List[Optional[Tensor]] = [None] * T
            order = range(T - 1, -1, -1) if suf else range(T)
            for t in order:
                g = inp[:, t] @ w_ih.t() + h @ w_hh.t() + b
                i, f, gg, o = g.split(H, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                ys[t] = h
            outs.append(torch.stack(ys, dim=1))
        inp = torch.cat(outs, dim=-1)
    return inp


def linear(sd: SD, x: Tensor, prefix: str) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


# --------------------------------------------------------------------------------------
# Network pieces (model/disentangled_vae.py)
# --------------------------------------------------------------------------------------
def encode(sd: SD, x: Tensor, training: bool, speaker_size: int = 4, latent_dim: int = 32,
           decisions: Optional[dict] = None, call: int = 0):
    """DisentangledVAE.encode, model/disentangled_vae.py:198-220."""
    R = x.shape[0]
    for i in range(3):
        x = relu_at(conv_bn(sd, x, f"enc_modules.{i}.0.conv", f"enc_modules.{i}.1", training), decisions,
                    f"enc_modules.{i}:{call}")
    x = x.transpose(1, 2)                                     # :204  [R,64,512]
    out = lstm(sd, x, "enc_lstm", 2, True)                    # :208  [R,64,128]
    out = out.reshape(R, -1)                                  # :209  time-major flatten
    out = relu_at(linear(sd, out, "enc_linear.linear_layer"), decisions, f"enc_linear:{call}")  # :211
    style = linear(sd, out, "style.linear_layer")             # :212
    content = linear(sd, out, "content.linear_layer")         # :213
    S, L = speaker_size, latent_dim
    return style[:, :S], style[:, S:], content[:, :L - S], content[:, L - S:]


def decode(sd: SD, z: Tensor, training: bool, decisions: Optional[dict] = None, call: int = 0) -> Tensor:
    """DisentangledVAE.decode, model/disentangled_vae.py:230-248."""
    R = z.shape[0]
    out = linear(sd, z, "dec_pre_linear1")
    out = linear(sd, out, "dec_pre_linear2")                  # no activation in between (:232-233)
    out = out.view(R, -1, 2 * DIM_NECK)                       # :235 [R,64,128]
    out = lstm(sd, out, "dec_lstm1", 1, False)                # :238 [R,64,512]
    out = out.transpose(-1, -2)
    for i in range(3):
        out = relu_at(conv_bn(sd, out, f"dec_modules.{i}.0", f"dec_modules.{i}.1", training), decisions,
                      f"dec_modules.{i}:{call}")
    out = out.transpose(-1, -2)
    out = lstm(sd, out, "dec_lstm2", 2, False)                # :246 [R,64,1024]
    out = linear(sd, out, "dec_linear2.linear_layer")         # :247 [R,64,80]
    return out.transpose(-1, -2)                              # :248 [R,80,64]


def postnet(sd: SD, x: Tensor, training: bool) -> Tensor:
    """Postnet.forward, model/disentangled_vae.py:81-87."""
    for i in range(4):
        x = torch.tanh(conv_bn(sd, x, f"postnet.convolutions.{i}.0.conv",
                               f"postnet.convolutions.{i}.1", training))
    return conv_bn(sd, x, "postnet.convolutions.4.0.conv", "postnet.convolutions.4.1", training)


def reparameterize(mu: Tensor, logvar: Tensor, eps: Optional[Tensor]) -> Tensor:
    """_reparameterize with externally supplied noise, model/disentangled_vae.py:222-228.

    eps=None is the `train=False` branch (returns mu)."""
    if eps is None:
        return mu
    return eps * torch.exp(0.5 * logvar) + mu


def forward(sd: SD, x1: Tensor, x2: Tensor, eps: Sequence[Tensor], training: bool = True,
            sample_content: bool = True, speaker_size: int = 4, latent_dim: int = 32, decisions: Optional[dict] = None):
    """DisentangledVAE.forward, model/disentangled_vae.py:250-279 -> the 10-tuple.

    `eps` = [eps_content1, eps_content2, eps_style] in the reference's draw order.
    `training` is the nn.Module mode (BatchNorm); `sample_content` is the `train` argument of
    forward (content noise on/off).  The style noise is always applied (:261).
    Mutates the BN buffers in `sd` exactly like the reference (x1 call first, then x2: F5).
    """
    s_mu1, s_lv1, c_mu1, c_lv1 = encode(sd, x1, training, speaker_size, latent_dim, decisions, 0)
    z_c1 = reparameterize(c_mu1, c_lv1, eps[0] if sample_content else None)
    s_mu2, s_lv2, c_mu2, c_lv2 = encode(sd, x2, training, speaker_size, latent_dim, decisions, 1)
    z_c2 = reparameterize(c_mu2, c_lv2, eps[1] if sample_content else None)
    s_mu2 = s_mu2.detach()                                    # :257
    s_lv2 = s_lv2.detach()                                    # :258
    z_s_mu = (s_mu1 + s_mu2) / 2                              # :259
    z_s_lv = (s_lv1 + s_lv2) / 2                              # :260
    z_s = reparameterize(z_s_mu, z_s_lv, eps[2])              # :261
    z1 = torch.cat((z_s, z_c1), dim=-1)
    z2 = torch.cat((z_s, z_c2), dim=-1)
    q1_mu = torch.cat((z_s_mu, c_mu1), dim=-1)
    q1_lv = torch.cat((z_s_lv, c_lv1), dim=-1)
    q2_mu = torch.cat((z_s_mu, c_mu2), dim=-1)
    q2_lv = torch.cat((z_s_lv, c_lv2), dim=-1)
    r1 = decode(sd, z1, training, decisions, 0)               # :274
    r2 = decode(sd, z2, training, decisions, 1)               # :275
    r1_hat = r1 + postnet(sd, r1, training)                   # :277
    r2_hat = r2 + postnet(sd, r2, training)                   # :278
    return r1, r2, r1_hat, r2_hat, q1_mu, q1_lv, q2_mu, q2_lv, z_s_mu, z_s_lv


def loss_gvae2(x1, x2, r1, r2, r1_hat, r2_hat, q1_mu, q1_lv, q2_mu, q2_lv, s_mu, s_lv,
               batch_size: int, mse_cof: float = 10.0, kl_cof: float = 10.0, signs: Optional[Sequence[Tensor]] = None):
    """ConvolutionalMulVAE.loss_functionGVAE2, model/disentangled_vae.py:310-327 -> 8-tuple.

    L1 sums are divided by the *constructor* batch_size; the style KL uses factor
    -1 and is reported only."""
    if signs is None:
        l1 = lambda a, b, _s: (a - b).abs().sum() / batch_size
        signs = [None] * 4
    else:   # matched-decision mode (see relu_at): |b - a| written with the supplied sign(b - a)
        l1 = lambda a, b, sg: (sg * (b - a)).sum() / batch_size
    m1, m2, m1h, m2h = l1(x1, r1, signs[0]), l1(x2, r2, signs[1]), l1(x1, r1_hat, signs[2]), l1(x2, r2_hat, signs[3])
    kl = lambda mu, lv: -0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp(), dim=-1).mean()
    k1, k2 = kl(q1_mu, q1_lv), kl(q2_mu, q2_lv)
    ks = -1.0 * torch.sum(1 + s_lv - s_mu.pow(2) - s_lv.exp()) / batch_size
    total = mse_cof * (m1 + m2 + m1h + m2h) + kl_cof * (k1 + k2)
    return total, m1, m2, m1h, m2h, k1, k2, ks


def clone_sd(sd: SD, requires_grad: bool = False, dtype=None, device=None) -> SD:
    out: SD = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if t.is_floating_point():
            if dtype is not None:
                t = t.to(dtype)
        if device is not None:
            t = t.to(device)
        if requires_grad and t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


def train_step(sd: SD, x1, x2, eps, batch_size: int, mse_cof=10.0, kl_cof=10.0,
               speaker_size: int = 4, latent_dim: int = 32, decisions: Optional[dict] = None):
    """forward + loss + backward (the timed unit: model/variational_base_vae.py:62-68).

    `sd` floating tensors must be leaves with requires_grad.  Returns (fwd10, loss8, grads)."""
    out = forward(sd, x1, x2, eps, True, True, speaker_size, latent_dim, decisions)
    losses = loss_gvae2(x1, x2, *out, batch_size=batch_size, mse_cof=mse_cof, kl_cof=kl_cof,
                        signs=None if decisions is None else decisions.get("l1_signs"))
    names = [k for k, v in sd.items() if v.requires_grad]
    grads = torch.autograd.grad(losses[0], [sd[k] for k in names], allow_unused=True)
    return out, losses, dict(zip(names, grads))


# --------------------------------------------------------------------------------------
# Conversion path (model/variational_base_vae.py:264-296, :335-348)
# --------------------------------------------------------------------------------------
def chunking_mel(mel: np.ndarray) -> np.ndarray:
    """chunking_mel, model/variational_base_vae.py:335-348: [80,T] -> [T//64+1, 80, 64].

    Non-overlapping 64-frame chunks; the last one is zero padded (a whole zero chunk when
    T % 64 == 0)."""
    n = mel.shape[1] // T_FRAMES + 1
    out = np.zeros((n, mel.shape[0], T_FRAMES), dtype=mel.dtype)
    for i in range(n):
        piece = mel[:, i * T_FRAMES:(i + 1) * T_FRAMES]
        out[i, :, :piece.shape[1]] = piece
    return out


def convert(sd: SD, source: Tensor, target: Tensor, speaker_size: int = 4, latent_dim: int = 32):
    """Core of voice_conversion_mel, model/variational_base_vae.py:277-296 (eval-mode BN).

    source [N,80,64], target [M,80,64] chunks of one utterance each.  Returns
    (recons [80, N*64], converted [80, N*64] clamped to [0,1])."""
    with torch.no_grad():
        s_mu, _, c_mu, _ = encode(sd, source, False, speaker_size, latent_dim)
        t_mu, _, _, _ = encode(sd, target, False, speaker_size, latent_dim)
        n = source.shape[0]
        src_style = s_mu.mean(dim=0, keepdim=True).repeat(n, 1)      # :281
        trg_style = t_mu.mean(dim=0, keepdim=True).repeat(n, 1)      # :282
        rec = decode(sd, torch.cat([src_style, c_mu], dim=-1), False)
        conv = decode(sd, torch.cat([trg_style, c_mu], dim=-1), False)
        conv = conv + postnet(sd, conv, False)                        # :291-293
        cat_t = lambda m: torch.cat([m[i] for i in range(m.shape[0])], dim=1)
        return cat_t(rec), torch.clamp(cat_t(conv), 0.0, 1.0)


# --------------------------------------------------------------------------------------
# Speaker-group utilities (model/utils.py) -- dead code in the reference, PoG mode oracle
# --------------------------------------------------------------------------------------
def group_segments(labels: np.ndarray):
    """Integer part of the group op: first-occurrence-ordered group index per row, group sizes.

    Bit-exact contract for the CUDA segment kernels.  Mirrors the dict insertion order of
    accumulate_group_evidence (model/utils.py:26-36)."""
    labels = np.asarray(labels).reshape(-1)
    order: Dict[int, int] = {}
    gid = np.empty(labels.shape[0], dtype=np.int64)
    for i, l in enumerate(labels.tolist()):
        if l not in order:
            order[l] = len(order)
        gid[i] = order[l]
    counts = np.bincount(gid, minlength=len(order)).astype(np.int64)
    return gid, counts


def accumulate_group_evidence(mu: np.ndarray, logvar: np.ndarray, labels: np.ndarray):
    """Product of Gaussians per label, model/utils.py:13-75, in fp32 numpy.

    var_g = 1 / sum_i 1/var_i ; mu_g = var_g * sum_i mu_i / var_i ; exact-zero variances are
    replaced by 1e-6 before inversion and before the final log.  Accumulation runs in row
    order within each group like the reference's Python loop.  Returns (group_mu, group_logvar)
    broadcast back to rows.  (The reference also overwrites its logvar argument with exp(logvar)
    in place; the oracle does not mutate.)"""
    mu = np.asarray(mu, dtype=np.float32)
    var = np.exp(np.asarray(logvar, dtype=np.float32)).astype(np.float32)
    var[var == 0.0] = np.float32(1e-6)
    gid, counts = group_segments(labels)
    G, D = counts.shape[0], mu.shape[1]
    inv_sum = np.zeros((G, D), dtype=np.float32)
    mu_sum = np.zeros((G, D), dtype=np.float32)
    one = np.float32(1.0)
    for i in range(mu.shape[0]):
        inv = (one / var[i]).astype(np.float32)
        inv_sum[gid[i]] += inv
        mu_sum[gid[i]] += mu[i] * inv
    gvar = (one / inv_sum).astype(np.float32)
    gmu = (mu_sum * gvar).astype(np.float32)
    gvar_rows = gvar[gid].copy()
    gvar_rows[gvar_rows == 0.0] = np.float32(1e-6)
    return gmu[gid], np.log(gvar_rows).astype(np.float32)


def group_wise_reparameterize(mu: np.ndarray, logvar: np.ndarray, labels: np.ndarray, eps_group: np.ndarray):
    """model/utils.py:95-116 with the per-group noise supplied: z_i = exp(0.5 lv_i) * eps_{g(i)} + mu_i.

    `eps_group` is [G, D] indexed by first-occurrence group order (the reference draws it
    N(0, 0.1) per unique label)."""
    gid, _ = group_segments(labels)
    std = np.exp(np.float32(0.5) * np.asarray(logvar, dtype=np.float32)).astype(np.float32)
    return (std * eps_group[gid].astype(np.float32) + np.asarray(mu, dtype=np.float32)).astype(np.float32)


def pair_mean_style(s_mu1, s_lv1, s_mu2, s_lv2):
    """The live 'group' op: arithmetic mean of the pair's parameters (model/disentangled_vae.py:259-260)."""
    return (s_mu1 + s_mu2) / 2, (s_lv1 + s_lv2) / 2


def assign_groups_to_ranks(counts: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Whole speaker groups per rank, contiguous group ranges balanced by row count.

    Returns [(first_group, last_group_exclusive)] per rank.  Greedy prefix split: rank r takes
    groups until its cumulative row count reaches (r+1)/world of the total."""
    total = int(sum(counts))
    out, g, acc = [], 0, 0
    for r in range(world):
        start = g
        target = total * (r + 1) / world
        while g < len(counts) and (acc + counts[g] <= target + 1e-9 or g == start) and \
                (len(counts) - g) > (world - 1 - r):
            acc += counts[g]
            g += 1
        if r == world - 1:
            while g < len(counts):
                acc += counts[g]
                g += 1
        out.append((start, g))
    return out
