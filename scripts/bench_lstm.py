"""Micro-benchmark of the LSTM recurrence kernels at the bench shape (rows = 1024).  Tile sizes come from the
DVAE_LSTM_* environment variables (see ops_gemm.cu), so run once per configuration."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import lib, ops

dt = lib.BF16
rows, T = int(os.environ.get("ROWS", "1024")), 64
cfg = {k: os.environ.get(k, "-") for k in ("DVAE_LSTM_FWD_TILE", "DVAE_LSTM_FWD_TILE_SMALL", "DVAE_LSTM_BWD_TILE")}
print("config", cfg)
for H, D in ((64, 2), (512, 1), (1024, 1)):
    xg0 = (torch.randn(rows, T, D * 4 * H, device="cuda")).to(torch.bfloat16)
    whh = (torch.randn(D, 4 * H, H, device="cuda") / H ** 0.5).to(torch.bfloat16)
    dh = torch.randn(rows, T, D * H, device="cuda").to(torch.bfloat16)

    def timeit(fn, n=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    xg = xg0.clone()
    out = {}

    def fwd():
        xg.copy_(xg0)
        out["h"], out["c"] = ops.lstm_fwd(dt, xg, whh, H, D)
    t_copy = timeit(lambda: xg.copy_(xg0))
    t_f = timeit(fwd) - t_copy
    t_b = timeit(lambda: ops.lstm_bwd(dt, dh, xg, out["c"], whh, H, D))
    fl = 2.0 * rows * 4 * H * H * D * T
    print(f"H={H:5d} D={D} fwd {t_f * 1e3 / T:7.2f} us/step ({fl / t_f / 1e9:7.1f} TF/s)   bwd {t_b * 1e3 / T:7.2f} us/step ({fl / t_b / 1e9:7.1f} TF/s)")
