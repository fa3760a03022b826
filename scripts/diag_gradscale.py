"""Diagnostic (GPU): largest magnitudes along the scaled activation-gradient stream of the fp16 mode (to check the
default Engine.grad_scale leaves head-room below 65504).  usage: python scripts/diag_gradscale.py <R>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
os.environ["DVAE_DEBUG_GRAD_STATS"] = "1"
os.environ["DVAE_B200_PRECISION"] = "fp16"
import torch

from model.disentangled_vae import ConvolutionalMulVAE
from oracle import dvae_oracle as O

R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sd = O.synth_state_dict(0)
x1, x2, eps = O.synth_inputs(R)
x1, x2, eps = x1.cuda(), x2.cuda(), [e.cuda() for e in eps]
w = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=R, speaker_size=4, latent_dim=32)
w.model.load_state_dict(sd)
w.model.train()
q = list(eps)
w.model.noise_hook = lambda shape: q.pop(0)
out = w.model(x1, x2)
w.loss_functionGVAE2(x1, x2, *out)[0].backward()
print(f"fp16 R={R} grad_scale={w.model.grad_scale}")
for name, amax in w.model._engine.grad_stats:
    print(f"  {name:40s} amax {amax:.4e}")
bad = [k for k, p in w.model.named_parameters() if not torch.isfinite(p.grad).all()]
print("non-finite parameter gradients:", bad)
