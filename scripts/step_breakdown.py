"""In-stream time of the phases of one training step (CUDA events around the engine's building blocks, no profiler).
usage: step_breakdown.py [bf16|tf32] [pairs]"""
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
os.environ["DVAE_B200_PRECISION"] = prec
from dvae_b200 import engine as E
from model.disentangled_vae import ConvolutionalMulVAE

records = []


def wrap(cls, name, label_fn):
    orig = getattr(cls, name)

    def f(self, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(self, *a, **k)
        e1.record()
        records.append((label_fn(a, k), e0, e1))
        return out
    setattr(cls, name, f)


wrap(E.Engine, "_conv_stack", lambda a, k: "fwd conv+BN stack (%s)" % a[4][0][0].split(".")[0])
wrap(E.Engine, "_lstm", lambda a, k: "fwd LSTM %s" % a[1])
wrap(E.Engine, "_conv_stack_bwd", lambda a, k: "bwd conv+BN stack (%s)" % a[2][0]["conv"].split(".")[0])
wrap(E.Engine, "_lstm_bwd", lambda a, k: "bwd LSTM %s" % a[1])
wrap(E.Engine, "_linear_bwd", lambda a, k: "bwd linear")
wrap(E.Engine, "forward", lambda a, k: "FORWARD total")
wrap(E.Engine, "backward", lambda a, k: "BACKWARD total")

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
steps = 6
for s in range(steps):
    if s == 2:
        records.clear()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True)
        t0.record()
    for p in w.model.parameters():
        p.grad = None
    out = w.model(x1, x2)
    losses = w.loss_functionGVAE2(x1, x2, *out)
    losses[0].backward()
t1 = torch.cuda.Event(enable_timing=True)
t1.record()
torch.cuda.synchronize()
n = steps - 2
agg = OrderedDict()
for label, e0, e1 in records:
    agg[label] = agg.get(label, 0.0) + e0.elapsed_time(e1)
print(f"step {t0.elapsed_time(t1) / n:.3f} ms ({prec}, {pairs} pairs); per-step phase times (ms):")
for k, v in agg.items():
    print(f"  {v / n:8.3f}  {k}")
