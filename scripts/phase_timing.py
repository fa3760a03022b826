"""Per-phase timing of single tc_gemm launches (entry -> setup -> producer done -> accumulator ready -> epilogue done)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import lib, ops

dt = lib.BF16
cap = 4096
buf = torch.zeros(cap * 5, dtype=torch.int64, device="cuda")


def report(title, launches=1):
    torch.cuda.synchronize()
    b = buf.view(cap, 5).cpu()
    used = b[:, 0] > 0
    b = b[used].double()
    if b.numel() == 0:
        print(title, "no stamps")
        return
    t0 = b[:, 0].min()
    rel = (b - t0) / 1e3
    names = ["entry", "setup", "producer_done", "acc_ready", "epi_done"]
    print(f"{title}: {int(used.sum())} CTAs; kernel span {(b[:, 4].max() - t0) / 1e3:.1f} us")
    for i, n in enumerate(names):
        col = rel[:, i][b[:, i] > 0]
        if col.numel():
            print(f"    {n:14s} min {col.min():7.2f}  median {col.median():7.2f}  max {col.max():7.2f} us")
    d = (b[:, 4] - b[:, 3]) / 1e3
    m = (b[:, 3] - b[:, 1]) / 1e3
    print(f"    per-CTA: setup->acc_ready median {m.median():.2f} us, epilogue median {d.median():.2f} us (max {d.max():.2f})")
    buf.zero_()


rows, H = 1024, int(os.environ.get("H", "1024"))
T = 3
xg = torch.randn(rows, T, 4 * H, device="cuda").to(torch.bfloat16)
whh = (torch.randn(1, 4 * H, H, device="cuda") / H ** 0.5).to(torch.bfloat16)
dh = torch.randn(rows, T, H, device="cuda").to(torch.bfloat16)
h, c = ops.lstm_fwd(dt, xg, whh, H, 1)     # warm
ops.lstm_bwd(dt, dh, xg, c, whh, H, 1)
torch.cuda.synchronize()
lib.call("dvae_debug_timing", buf.data_ptr(), cap)
# the last launch of each call overwrites the stamps of the earlier ones (same CTA ids): T=3 -> stamps of step 2
xg = torch.randn(rows, T, 4 * H, device="cuda").to(torch.bfloat16)
torch.cuda.synchronize()
h, c = ops.lstm_fwd(dt, xg, whh, H, 1)
report(f"LSTM fwd step (H={H}, last of 3)")
ops.lstm_bwd(dt, dh, xg, c, whh, H, 1)
report(f"LSTM bwd step (H={H}, last of 3)")
x = torch.randn(rows, 64, 512, device="cuda").to(torch.bfloat16)
wk = (torch.randn(512, 5, 512, device="cuda") * 0.02).to(torch.bfloat16)
ops.conv5_fwd(dt, x, wk, torch.zeros(512, device="cuda"))
report("conv5 fwd 512->512")
lib.call("dvae_debug_timing", None, 0)
