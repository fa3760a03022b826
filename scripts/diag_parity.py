"""Diagnostic (GPU): per-tensor forward error, loss error and the lowest gradient cosines of the drop-in model vs the
fp32 oracle.  usage: python scripts/diag_parity.py <bf16|tf32> <R>"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
import torch.nn.functional as F

name, R = sys.argv[1], int(sys.argv[2])
os.environ["DVAE_B200_PRECISION"] = name
from model.disentangled_vae import ConvolutionalMulVAE
from oracle import dvae_oracle as O

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
sd = O.synth_state_dict(0)
x1, x2, eps = O.synth_inputs(R)
x1, x2, eps = x1.cuda(), x2.cuda(), [e.cuda() for e in eps]
w = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=R, speaker_size=4, latent_dim=32)
w.model.load_state_dict(sd)
w.model.train()
q = list(eps)
w.model.noise_hook = lambda shape: q.pop(0)
w.model._debug_keep_saved = True
out = w.model(x1, x2)
losses = w.loss_functionGVAE2(x1, x2, *out)
losses[0].backward()
osd = O.clone_sd(sd, requires_grad=True, device="cuda")
o_out, o_losses, o_grads = O.train_step(osd, x1, x2, eps, batch_size=R)
print(f"== {name} R={R}")
for n, a, b in zip(["r1", "r2", "r1h", "r2h", "q1mu", "q1lv", "q2mu", "q2lv", "smu", "slv"], out, o_out):
    print(f"  {n:5s} relL2 {(a - b).norm().item() / b.norm().item():.2e}")
for n, a, b in zip(["LOSS", "m1", "m2", "m1h", "m2h", "k1", "k2", "ks"], losses, o_losses):
    print(f"  {n:5s} {a.item():.6f} vs {b.item():.6f} rel {abs(a.item() - b.item()) / abs(b.item()):.2e}")
zero = lambda k: re.search(r"(\.0\.conv\.bias$)|(^dec_modules\.\d\.0\.bias$)", k) is not None


def report(title, mine, ref, top=6):
    rows, dot, na, nb = [], 0.0, 0.0, 0.0
    for k in ref:
        if zero(k):
            continue
        g, o = mine[k].flatten().double(), ref[k].flatten().double()
        rows.append((F.cosine_similarity(g, o, dim=0).item(), (g - o).norm().item() / o.norm().item(), o.norm().item(), k))
        dot += (g * o).sum().item(); na += (g * g).sum().item(); nb += (o * o).sum().item()
    rows.sort()
    print(f"  [{title}] global cosine {dot / (na * nb) ** 0.5:.6f}; min per-tensor {rows[0][0]:.5f}; "
          f"tensors below 0.999: {sum(r[0] < 0.999 for r in rows)} of {len(rows)}")
    for c, r, n_, k in rows[:top]:
        print(f"      cos {c:.5f} relL2 {r:.2e} |g| {n_:.3e} {k}")


mine = {k: p.grad for k, p in w.model.named_parameters()}
report("dvae_b200 vs fp32 oracle, free decisions", mine, o_grads)
from dvae_b200.engine import Engine
dec = Engine.discrete_decisions(w.model._last_saved, [t.detach() for t in out], x1, x2)
osd2 = O.clone_sd(sd, requires_grad=True, device="cuda")
_, _, m_grads = O.train_step(osd2, x1, x2, eps, batch_size=R, decisions=dec)
report("dvae_b200 vs fp32 oracle, MATCHED decisions", mine, m_grads)
# context: PyTorch's own reduced-precision runs of the same oracle against its fp32 run
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
osd3 = O.clone_sd(sd, requires_grad=True, device="cuda")
t_out, _, t_grads = O.train_step(osd3, x1, x2, eps, batch_size=R)
print(f"  torch TF32 (cuDNN/cuBLAS) forward r1 relL2 {(t_out[0] - o_out[0]).norm().item() / o_out[0].norm().item():.2e}, "
      f"r1h {(t_out[2] - o_out[2]).norm().item() / o_out[2].norm().item():.2e}")
report("torch TF32 oracle vs fp32 oracle, free decisions", t_grads, o_grads, top=2)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
osd4 = O.clone_sd(sd, requires_grad=True, device="cuda")
with torch.autocast("cuda", dtype=torch.bfloat16):
    b_out, b_losses, b_grads = O.train_step(osd4, x1, x2, eps, batch_size=R)
print(f"  torch bf16 autocast forward r1 relL2 {(b_out[0].float() - o_out[0]).norm().item() / o_out[0].norm().item():.2e}, "
      f"r1h {(b_out[2].float() - o_out[2]).norm().item() / o_out[2].norm().item():.2e}")
report("torch bf16-autocast oracle vs fp32 oracle, free decisions", b_grads, o_grads, top=2)
