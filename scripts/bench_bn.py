"""In-stream timing of the BatchNorm kernels at the bench shape ([65536, 512] activations, two statistics halves)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import lib, ops

dt = lib.BF16
rows, C = 65536, int(os.environ.get("C", "512"))
NB = 3
ys = [torch.randn(rows, C, device="cuda").to(torch.bfloat16) for _ in range(NB)]
ds = [torch.randn(rows, C, device="cuda").to(torch.bfloat16) for _ in range(NB)]
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
rm, rv, nb = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros((), device="cuda", dtype=torch.long)


def timeit(fn, n=9):
    for i in range(NB):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i % NB)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


eb = 2
for act in (lib.ACT_RELU, lib.ACT_TANH):
    out = {}

    def f(i):
        out["o"], out["s"] = ops.bn_train_fwd(dt, ys[i], g, b, rm, rv, nb, 2, act, 1e-5, 0.1)
    t_f = timeit(f)
    t_b = timeit(lambda i: ops.bn_train_bwd(dt, ds[i], ys[i], out["s"], 2, act))
    by_f = rows * C * eb * 3        # stats read + apply read/write
    by_b = rows * C * eb * 5        # reduce reads 2, apply reads 2 writes 1
    print(f"act {act}: fwd {t_f:6.1f} us ({by_f / t_f / 1e6:5.2f} TB/s of {by_f / 1e6:.0f} MB)   bwd {t_b:6.1f} us ({by_b / t_b / 1e6:5.2f} TB/s of {by_b / 1e6:.0f} MB)")
