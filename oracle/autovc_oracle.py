"""CPU oracle for the AutoVC-style generator (BASELINE config 5).  TEST INFRASTRUCTURE ONLY (see dvae_oracle.py).

Functional fp32 restatement of `autovc_replicate/proposed_autovc.py:41-220` over a flat state_dict, pinned against the
reference module by `oracle/make_golden.py` (tests/golden/autovc_R4.pt).  The reference defines no loss for this
network; tests and the bench use the squared error of both outputs against the input mel (a smooth loss, our choice).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from oracle.dvae_oracle import _key_rng, conv_bn, linear, lstm, relu_at

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def param_and_buffer_shapes() -> Dict[str, Tuple[int, ...]]:
    """state_dict inventory in the reference's registration order (78 parameters + 33 BatchNorm buffers)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv_bn_block(prefix, ci, co):
        s[f"{prefix}.0.conv.weight"] = (co, ci, 5)
        s[f"{prefix}.0.conv.bias"] = (co,)
        s[f"{prefix}.1.weight"] = (co,)
        s[f"{prefix}.1.bias"] = (co,)
        s[f"{prefix}.1.running_mean"] = (co,)
        s[f"{prefix}.1.running_var"] = (co,)
        s[f"{prefix}.1.num_batches_tracked"] = ()

    def lstm_block(prefix, inp, hid, layers, bi):
        for l in range(layers):
            i = inp if l == 0 else hid * (2 if bi else 1)
            for suf in (("", "_reverse") if bi else ("",)):
                s[f"{prefix}.weight_ih_l{l}{suf}"] = (4 * hid, i)
                s[f"{prefix}.weight_hh_l{l}{suf}"] = (4 * hid, hid)
                s[f"{prefix}.bias_ih_l{l}{suf}"] = (4 * hid,)
                s[f"{prefix}.bias_hh_l{l}{suf}"] = (4 * hid,)

    for i in range(3):
        conv_bn_block(f"encoder.convolutions.{i}", 80 if i == 0 else 512, 512)
    lstm_block("encoder.lstm", 512, 64, 2, True)
    s["encoder.latent_code.linear_layer.weight"] = (256, 8192)
    s["encoder.latent_code.linear_layer.bias"] = (256,)
    s["decoder.dec_linear.linear_layer.weight"] = (8192, 256)
    s["decoder.dec_linear.linear_layer.bias"] = (8192,)
    lstm_block("decoder.lstm1", 128, 512, 1, False)
    for i in range(3):
        conv_bn_block(f"decoder.convolutions.{i}", 512, 512)
    lstm_block("decoder.lstm2", 512, 1024, 2, False)
    s["decoder.linear_projection.linear_layer.weight"] = (80, 1024)
    s["decoder.linear_projection.linear_layer.bias"] = (80,)
    for i, (ci, co) in enumerate([(80, 512), (512, 512), (512, 512), (512, 512), (512, 80)]):
        conv_bn_block(f"postnet.convolutions.{i}", ci, co)
    return s


def synth_state_dict(seed: int = 0) -> SD:
    """Deterministic weights (numpy Philox, torch-RNG independent) with reference-like magnitudes; BN randomised."""
    sd: SD = {}
    for name, shape in param_and_buffer_shapes().items():
        g = _key_rng("autovc." + name, seed)
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.long)
            continue
        if name.endswith("running_mean"):
            arr = g.uniform(-0.1, 0.1, size=shape)
        elif name.endswith("running_var"):
            arr = g.uniform(0.8, 1.2, size=shape)
        elif "lstm" in name:
            bound = 1.0 / math.sqrt(shape[0] // 4)
            arr = g.uniform(-bound, bound, size=shape)
        elif ".1.weight" in name:
            arr = g.uniform(0.5, 1.5, size=shape)
        elif ".1.bias" in name:
            arr = g.uniform(-0.2, 0.2, size=shape)
        elif len(shape) >= 2:
            rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
            bound = math.sqrt(6.0 / (shape[1] * rf + shape[0] * rf))
            arr = g.uniform(-bound, bound, size=shape)
        else:
            arr = g.uniform(-0.05, 0.05, size=shape)
        sd[name] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
    return sd


def generator_forward(sd: SD, x: Tensor, training: bool = True, decisions: Optional[dict] = None):
    """Generator.forward, autovc_replicate/proposed_autovc.py:196-220 -> (mel [B,1,64,80], mel_postnet [B,1,64,80])."""
    R = x.shape[0]
    h = x
    for i in range(3):      # Encoder.forward :64-85
        h = relu_at(conv_bn(sd, h, f"encoder.convolutions.{i}.0.conv", f"encoder.convolutions.{i}.1", training),
                    decisions, f"encoder.convolutions.{i}:0")
    h = lstm(sd, h.transpose(1, 2), "encoder.lstm", 2, True).reshape(R, -1)
    codes = linear(sd, h, "encoder.latent_code.linear_layer")            # no activation (:78)
    d = linear(sd, codes, "decoder.dec_linear.linear_layer").view(R, -1, 128)   # Decoder.forward :121-136
    d = lstm(sd, d, "decoder.lstm1", 1, False).transpose(1, 2)
    for i in range(3):
        d = relu_at(conv_bn(sd, d, f"decoder.convolutions.{i}.0.conv", f"decoder.convolutions.{i}.1", training),
                    decisions, f"decoder.convolutions.{i}:0")
    d = lstm(sd, d.transpose(1, 2), "decoder.lstm2", 2, False)
    mel = linear(sd, d, "decoder.linear_projection.linear_layer")        # [B,64,80]
    p = mel.transpose(2, 1)
    for i in range(4):      # Postnet.forward :176-183
        p = torch.tanh(conv_bn(sd, p, f"postnet.convolutions.{i}.0.conv", f"postnet.convolutions.{i}.1", training))
    p = conv_bn(sd, p, "postnet.convolutions.4.0.conv", "postnet.convolutions.4.1", training)
    mel_post = mel + p.transpose(2, 1)
    return mel.unsqueeze(1), mel_post.unsqueeze(1)


def sq_loss(x: Tensor, mel: Tensor, mel_post: Tensor) -> Tensor:
    """0.5 * sum of squared errors of both outputs against the input mel (x [B,80,64])."""
    t = x.transpose(1, 2).unsqueeze(1)
    return 0.5 * ((mel - t).pow(2).sum() + (mel_post - t).pow(2).sum())


def train_step(sd: SD, x: Tensor, decisions: Optional[dict] = None):
    mel, mel_post = generator_forward(sd, x, True, decisions)
    loss = sq_loss(x, mel, mel_post)
    names = [k for k, v in sd.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
    return (mel, mel_post), loss, dict(zip(names, grads))
