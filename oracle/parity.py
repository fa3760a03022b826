"""Parity report of the drop-in modules against the fp32 oracle (TEST INFRASTRUCTURE: used by tests/ and by the `parity`
block of bench.py as the checker, never by the product path).

`parity_report(precision, R)` runs ONE training step (forward 10-tuple, loss 8-tuple, backward) of
model.disentangled_vae.ConvolutionalMulVAE on cuda and the oracle (oracle/dvae_oracle.py, a restatement of
/root/reference/model/disentangled_vae.py:250-279, :310-327 pinned on the reference's own outputs by tests/golden) on
the same synthetic inputs, weights and noise, and returns the numbers north_star's tolerance is stated in:
relative L2 error per forward tensor, relative error per loss term, and the cosine of the parameter gradients -- at
FREE branch decisions and at MATCHED decisions (the oracle re-evaluated with the candidate's ReLU masks / L1 signs)."""
from __future__ import annotations

import os
import re
from typing import Dict

import torch
import torch.nn.functional as F

from . import dvae_oracle as O

FWD_NAMES = ["recons_x1", "recons_x2", "recons_x1_hat", "recons_x2_hat", "q_z1_mu", "q_z1_logvar", "q_z2_mu", "q_z2_logvar",
             "z_style_mu", "z_style_logvar"]
LOSS_NAMES = ["LOSS", "MSE_x1", "MSE_x2", "MSE_x1_hat", "MSE_x2_hat", "z1_kl", "z2_kl", "z_kl_style"]
# conv biases that feed a train-mode BatchNorm: identically-zero gradient (rounding noise in the reference)
ZERO_GRAD = lambda k: re.search(r"(\.0\.conv\.bias$)|(^dec_modules\.\d\.0\.bias$)", k) is not None


def grad_cosines(mine: Dict[str, torch.Tensor], ref: Dict[str, torch.Tensor]):
    """(min per-tensor cosine, its parameter name, cosine over all parameters, number of tensors below 0.999)."""
    worst, dot, na, nb, below = (1.0, ""), 0.0, 0.0, 0.0, 0
    for k, o in ref.items():
        if ZERO_GRAD(k):
            continue
        g, o = mine[k].flatten().double(), o.flatten().double()
        c = F.cosine_similarity(g, o, dim=0).item()
        below += c < 0.999
        worst = min(worst, (c, k))
        dot, na, nb = dot + (g * o).sum().item(), na + (g * g).sum().item(), nb + (o * o).sum().item()
    return worst[0], worst[1], dot / (na * nb) ** 0.5, below


def parity_report(precision: str, R: int, seed: int = 0) -> dict:
    from dvae_b200.engine import Engine
    from model.disentangled_vae import ConvolutionalMulVAE
    tf32_flags = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        sd = O.synth_state_dict(seed)
        x1, x2, eps = O.synth_inputs(R)
        x1, x2, eps = x1.cuda(), x2.cuda(), [e.cuda() for e in eps]
        prev = os.environ.get("DVAE_B200_PRECISION")
        os.environ["DVAE_B200_PRECISION"] = precision      # the wrapper's constructor has no precision argument (reference API)
        try:
            w = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=R, speaker_size=4, latent_dim=32,
                                    beta=0.1, mse_cof=10, kl_cof=10, style_cof=0.1)
        finally:
            if prev is None:
                del os.environ["DVAE_B200_PRECISION"]
            else:
                os.environ["DVAE_B200_PRECISION"] = prev
        w.model.load_state_dict(sd)
        w.model.train()
        queue = list(eps)
        w.model.noise_hook = lambda shape: queue.pop(0)
        w.model._debug_keep_saved = True
        out = w.model(x1, x2)
        losses = w.loss_functionGVAE2(x1, x2, *out)
        losses[0].backward()
        mine = {k: p.grad for k, p in w.model.named_parameters()}
        osd = O.clone_sd(sd, requires_grad=True, device="cuda")
        o_out, o_losses, o_grads = O.train_step(osd, x1, x2, eps, batch_size=R)
        decisions = Engine.discrete_decisions(w.model._last_saved, [t.detach() for t in out], x1, x2)
        w.model._last_saved = None
        osd2 = O.clone_sd(sd, requires_grad=True, device="cuda")
        _, _, m_grads = O.train_step(osd2, x1, x2, eps, batch_size=R, decisions=decisions)
        fwd = {n: (a - b).norm().item() / b.norm().item() for n, a, b in zip(FWD_NAMES, out, o_out)}
        loss = {n: abs(a.item() - b.item()) / abs(b.item()) for n, a, b in zip(LOSS_NAMES, losses, o_losses)}
        fm, fk, fg, fb = grad_cosines(mine, o_grads)
        mm, mk, mg, mb = grad_cosines(mine, m_grads)
        non_hat = [v for n, v in fwd.items() if not n.endswith("hat")]
        return {
            "precision": precision, "rows_per_call": R, "checker": "fp32 oracle on cuda (cuDNN / cuBLAS, TF32 off), same inputs / weights / noise",
            "forward_rel_l2": fwd, "forward_rel_l2_max_8_non_hat": max(non_hat),
            "forward_rel_l2_max_hat": max(v for n, v in fwd.items() if n.endswith("hat")),
            "loss_rel": loss,
            "grad_cosine_matched_decisions": {"min": mm, "argmin": mk, "global": mg, "tensors_below_0.999": mb},
            "grad_cosine_free_decisions": {"min": fm, "argmin": fk, "global": fg, "tensors_below_0.999": fb},
        }
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32_flags
