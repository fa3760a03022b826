"""Host time to enqueue one training step (no synchronisation inside the loop) vs the GPU time of the step: is the step
launch-bound?  usage: cpu_enqueue_time.py [pairs]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from model.disentangled_vae import ConvolutionalMulVAE

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
R = pairs * 2
w = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=R, speaker_size=4, latent_dim=32)
w.model.train()
x1, x2 = torch.rand(R, 80, 64, device="cuda"), torch.rand(R, 80, 64, device="cuda")
noise = [torch.randn(R, 28, device="cuda"), torch.randn(R, 28, device="cuda"), torch.randn(R, 4, device="cuda")]
k = [0]


def hook(shape):
    k[0] += 1
    return noise[(k[0] - 1) % 3]


w.model.noise_hook = hook
params = list(w.model.parameters())


def step():
    for p in params:
        p.grad = None
    out = w.model(x1, x2)
    losses = w.loss_functionGVAE2(x1, x2, *out)
    losses[0].backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    step()
e1.record()
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"pairs {pairs}: host enqueue {t_enq / n * 1e3:.2f} ms/step, GPU {e0.elapsed_time(e1) / n:.2f} ms/step, wall {t_all / n * 1e3:.2f} ms/step")
