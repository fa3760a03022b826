"""One forward (time-resident kernel) and one backward (step-per-launch kernels) of an LSTM layer at rows = 1024 (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import lib, ops

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rows, T = 1024, 64
dt, td = lib.F16, torch.float16
xg = torch.randn(rows, T, 4 * H, device="cuda").to(td)
whh = (torch.randn(1, 4 * H, H, device="cuda") / H ** 0.5).to(td)
dh = (torch.randn(rows, T, H, device="cuda") * 0.1).to(td)
h, c = ops.lstm_fwd(dt, xg, whh, H, 1)
da = ops.lstm_bwd(dt, dh, xg, c, whh, H, 1)
torch.cuda.synchronize()
print("ok", float(h.float().abs().mean()), float(da.float().abs().mean()))
