"""One warm-up + one profiled training step at the bench shape (for `ncu`).  usage: profile_step.py [bf16|tf32] [pairs] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
os.environ["DVAE_B200_PRECISION"] = prec
from model.disentangled_vae import ConvolutionalMulVAE

R = pairs * 2
torch.manual_seed(0)
w = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=R, speaker_size=4, latent_dim=32)
w.model.train()
g = torch.Generator(device="cuda").manual_seed(1)
x1, x2 = torch.rand(R, 80, 64, device="cuda", generator=g), torch.rand(R, 80, 64, device="cuda", generator=g)
noise = [torch.randn(R, 28, device="cuda"), torch.randn(R, 28, device="cuda"), torch.randn(R, 4, device="cuda")]
i = [0]


def hook(shape):
    i[0] += 1
    return noise[(i[0] - 1) % 3]


w.model.noise_hook = hook
for s in range(steps):
    for p in w.model.parameters():
        p.grad = None
    out = w.model(x1, x2)
    losses = w.loss_functionGVAE2(x1, x2, *out)
    losses[0].backward()
torch.cuda.synchronize()
print("LOSS", losses[0].item())
