"""HBM roofline of the speaker-group kernels at north_star's standalone size: B = 2^22 rows, D = 32, 2^17 groups.
Algorithmic bytes: read mu, logvar (2*B*D*4) + ids, write group_mu, group_logvar (2*B*D*4) + ids again."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import ops

B, D, G = 1 << 22, 32, 1 << 17
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
g = torch.Generator(device="cuda").manual_seed(0)
mu = torch.randn(B, D, device="cuda", generator=g)
lv = torch.randn(B, D, device="cuda", generator=g) * 0.5
labels = torch.arange(B, device="cuda", dtype=torch.int64) // (B // G)       # sorted speaker ids, 32 rows per group
gid, ng = ops.segment_ids_sorted(labels)
assert int(ng.item()) == G
acc = torch.zeros(G, 2, D, device="cuda")
cnt = torch.zeros(G, device="cuda")
out_a, out_b = torch.empty_like(mu), torch.empty_like(mu)
table = torch.empty_like(acc)
from dvae_b200.lib import call, ptr, stream


def accumulate():
    acc.zero_(); cnt.zero_()
    call("dvae_group_accumulate", ops.MODE_POG, ptr(mu), ptr(lv), ptr(gid), ptr(acc), ptr(cnt), B, D, stream())


def finalize():
    call("dvae_group_finalize", ops.MODE_POG, ptr(acc), ptr(cnt), ptr(gid), ptr(table), ptr(out_a), ptr(out_b), B, G, D, stream())


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


t_zero = timeit(lambda: (acc.zero_(), cnt.zero_()))
t_acc = timeit(accumulate) - t_zero
t_fin = timeit(finalize)
b_acc = 2 * B * D * 4 + B * 4
b_fin = 2 * B * D * 4 + B * 4
res = {"B": B, "D": D, "groups": G, "peak_gbs": peak,
       "accumulate": {"ms": t_acc * 1e3, "GBps": b_acc / t_acc / 1e9, "frac": b_acc / t_acc / 1e9 / peak, "bytes": b_acc},
       "finalize": {"ms": t_fin * 1e3, "GBps": b_fin / t_fin / 1e9, "frac": b_fin / t_fin / 1e9 / peak, "bytes": b_fin},
       "both": {"ms": (t_acc + t_fin) * 1e3, "GBps": (b_acc + b_fin) / (t_acc + t_fin) / 1e9,
                "frac": (b_acc + b_fin) / (t_acc + t_fin) / 1e9 / peak}}
print(json.dumps(res))
