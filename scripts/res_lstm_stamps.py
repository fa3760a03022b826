"""Per-step phase stamps of CTA 0 of the time-resident LSTM forward kernel (csrc/ops_lstm_res.cu): where a step's chain goes.
`python scripts/res_lstm_stamps.py [H]`"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import lib, ops

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bwd = False
rows, T = 1024, 64
dt, td = lib.F16, torch.float16
xg0 = torch.randn(rows, T, 4 * H, device="cuda").to(td)
whh = (torch.randn(1, 4 * H, H, device="cuda") / H ** 0.5).to(td)
dh = (torch.randn(rows, T, H, device="cuda") * 0.1).to(td)
buf = torch.zeros(T * 2 * 8, dtype=torch.int64, device="cuda")
names = ["wait", "first_kblock", "loads_issued", "mma_done", "acc_seen", "tile_done", "agent", "published"]
for rep in range(2):
    xg = xg0.clone()
    if not bwd:
        lib.call("dvae_debug_res_stamps", buf.data_ptr())
    h, c = ops.lstm_fwd(dt, xg, whh, H, 1)
    if bwd:
        lib.call("dvae_debug_res_stamps", buf.data_ptr())
        ops.lstm_bwd(dt, dh, xg, c, whh, H, 1)
    torch.cuda.synchronize()
    lib.call("dvae_debug_res_stamps", None)
s = buf.cpu().view(T, 2, 8).double()
t0 = s[1, 0, 0].item()
print(f"H={H} {'bwd' if bwd else 'fwd'}: us since step 1's first wait; columns " + " ".join(names))
for st in list(range(1, 6)) + list(range(30, 34)):
    for sl in range(2):
        print(f"  st {st:2d} slot {sl}: " + " ".join(f"{(s[st, sl, k].item() - t0) / 1e3:8.2f}" for k in range(8)))
d = (s[2:, :, :] - s[1:-1, :, :]).mean(dim=(0, 1)) / 1e3
print("  mean period per point (us): " + " ".join(f"{x:.2f}" for x in d.tolist()))
seg = (s[1:, :, 1:] - s[1:, :, :-1]).mean(dim=(0, 1)) / 1e3
print("  mean segment (us): " + " ".join(f"{names[k]}->{names[k + 1]} {seg[k].item():.2f}" for k in range(7)))
nxt = (s[2:, :, 1] - s[1:-1, :, 7]).mean().item() / 1e3
print(f"  published(st) -> flag(st+1) seen by this CTA: {nxt:.2f} us")
