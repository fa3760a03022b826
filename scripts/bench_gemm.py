"""Micro-benchmark of the large contractions at the bench shape (rows = 1024 sequences x 64 frames): Conv1d(k=5) forward /
dgrad / wgrad and the LSTM input projections.  CUDA events, inputs rotated over 3 buffer sets (> L2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import lib, ops

dt = lib.BF16
R, T = int(os.environ.get("ROWS", "1024")), 64
NB = 3


def timeit(fn, n=6):
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


def bf(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


for Cin, Cout in ((512, 512), (80, 512), (512, 80)):
    xs = [bf(R, T, Cin) for _ in range(NB)]
    dys = [bf(R, T, Cout) for _ in range(NB)]
    wk = bf(Cout, 5, Cin, scale=0.02)
    bias = torch.zeros(Cout, device="cuda")
    dwk = torch.zeros(Cout, 5, Cin, device="cuda")
    fl = 2.0 * R * T * Cin * Cout * 5
    t_f = timeit(lambda i: ops.conv5_fwd(dt, xs[i], wk, bias))
    t_d = timeit(lambda i: ops.conv5_dgrad(dt, dys[i], wk))
    t_w = timeit(lambda i: ops.conv5_wgrad(dt, dys[i], xs[i], dwk))
    print(f"conv5 {Cin:4d}->{Cout:4d}: fwd {t_f:7.1f} us ({fl / t_f / 1e6:6.0f} TF/s)  dgrad {t_d:7.1f} us ({fl / t_d / 1e6:6.0f})  "
          f"wgrad {t_w:7.1f} us ({fl / t_w / 1e6:6.0f})")
for K, N in ((512, 4096), (1024, 4096), (128, 2048), (512, 512)):
    M = R * T
    xs = [bf(M, K) for _ in range(NB)]
    dys = [bf(M, N) for _ in range(NB)]
    w = bf(N, K, scale=0.02)
    bias = torch.zeros(N, device="cuda")
    dw = torch.zeros(N, K, device="cuda")
    fl = 2.0 * M * N * K
    t_f = timeit(lambda i: ops.linear_fwd(dt, xs[i], w, bias))
    t_d = timeit(lambda i: ops.linear_dgrad(dt, dys[i], w))
    t_w = timeit(lambda i: ops.linear_wgrad(dt, dys[i], xs[i], dw))
    print(f"linear M={M} K={K:4d} N={N:4d}: fwd {t_f:7.1f} us ({fl / t_f / 1e6:6.0f} TF/s)  dgrad {t_d:7.1f} us ({fl / t_d / 1e6:6.0f})  "
          f"wgrad {t_w:7.1f} us ({fl / t_w / 1e6:6.0f})")
