"""Minimal launch sequence for `ncu --set full`: one conv GEMM and a few LSTM step kernels at the bench shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import lib, ops

dt = lib.BF16
rows, T, H = 1024, 4, 1024
x = torch.randn(rows, 64, 512, device="cuda").to(torch.bfloat16)
wk = (torch.randn(512, 5, 512, device="cuda") * 0.02).to(torch.bfloat16)
bias = torch.zeros(512, device="cuda")
ops.conv5_fwd(dt, x, wk, bias)                       # launch 0: conv fwd (K/K, BLOCK_N 256)
dy = torch.randn(rows, 64, 512, device="cuda").to(torch.bfloat16)
dwk = torch.zeros(512, 5, 512, device="cuda")
ops.conv5_wgrad(dt, dy, x, dwk)                      # launch 1: conv wgrad (MN/MN, split-K, atomics)
xg = torch.randn(rows, T, 4 * H, device="cuda").to(torch.bfloat16)
whh = (torch.randn(1, 4 * H, H, device="cuda") / H ** 0.5).to(torch.bfloat16)
h, c = ops.lstm_fwd(dt, xg, whh, H, 1)               # launches 2..5: LSTM fwd steps (step 0 has no MMA)
dh = torch.randn(rows, T, H, device="cuda").to(torch.bfloat16)
ops.lstm_bwd(dt, dh, xg, c, whh, H, 1)               # launches 6..9: LSTM bwd steps
H2 = 64
xg2 = torch.randn(rows, 64, 2 * 4 * H2, device="cuda").to(torch.bfloat16)
whh2 = (torch.randn(2, 4 * H2, H2, device="cuda") / H2 ** 0.5).to(torch.bfloat16)
h2, c2 = ops.lstm_fwd(dt, xg2, whh2, H2, 2)          # sequence-resident BiLSTM forward (one launch, 64 steps)
dh2 = torch.randn(rows, 64, 2 * H2, device="cuda").to(torch.bfloat16)
ops.lstm_bwd(dt, dh2, xg2, c2, whh2, H2, 2)          # sequence-resident backward
torch.cuda.synchronize()
print("done")
