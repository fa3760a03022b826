"""Per-step phase timing of the sequence-resident LSTM kernels (H = 64), CTA (0,0), thread 0, in SM clocks.
Slots: 0 before the MMA wait, 1 accumulator ready, 2 cell done, 3 A tile written, 4 fences done, 5 barrier passed,
6 next MMA issued."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import lib, ops

dt = lib.BF16
rows, T, H, D = int(os.environ.get("ROWS", "1024")), 64, 64, 2
xg = torch.randn(rows, T, D * 4 * H, device="cuda").to(torch.bfloat16)
whh = (torch.randn(D, 4 * H, H, device="cuda") / H ** 0.5).to(torch.bfloat16)
dh = torch.randn(rows, T, D * H, device="cuda").to(torch.bfloat16)
buf = torch.zeros(T * 8, device="cuda", dtype=torch.int64)
names = ["mma wait", "tmem->cell", "write A", "fences"]
for which in ("fwd", "bwd"):
    h, c = ops.lstm_fwd(dt, xg.clone(), whh, H, D)
    buf.zero_()
    lib.call("dvae_debug_seq_stamps", buf.data_ptr())
    if which == "fwd":
        ops.lstm_fwd(dt, xg.clone(), whh, H, D)
    else:
        ops.lstm_bwd(dt, dh, xg, c, whh, H, D)
    torch.cuda.synchronize()
    lib.call("dvae_debug_seq_stamps", None)
    st = buf.view(T, 8).cpu()
    print("total clk", (st[T - 1, 4] - st[0, 0]).item(), "per-step (slot0 deltas):", (st[1:, 0] - st[:-1, 0]).tolist())
    print("phase max:", (st[1:T, 1:5] - st[1:T, 0:4]).max(0).values.tolist(), "arrive->ready max", (st[1:T, 0] - st[0:T - 1, 4]).max().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xx = xg.clone()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        if which == "fwd":
            ops.lstm_fwd(dt, xx, whh, H, D)
        else:
            ops.lstm_bwd(dt, dh, xg, c, whh, H, D)
    e1.record()
    torch.cuda.synchronize()
    print(which, "kernel+host us per call", e0.elapsed_time(e1) * 1e3 / 5)
    d = (st[2:T - 1, 1:5] - st[2:T - 1, 0:4]).float()
    step = (st[3:T - 1, 0] - st[2:T - 2, 0]).float()
    print(f"{which}: step {step.median().item():.0f} clk;  " +
          "  ".join(f"{n} {d[:, i].median().item():.0f}" for i, n in enumerate(names)) +
          f"  arrive->next inputs ready {(st[3:T - 1, 0] - st[2:T - 2, 4]).float().median().item():.0f}")
