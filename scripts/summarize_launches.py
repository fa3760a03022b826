"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total us, share.
usage: summarize_launches.py launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    rows.append((int(r["ID"]), r["Kernel Name"], us, r.get("Grid Size", ""), r.get("Block Size", "")))
rows.sort()
rows = rows[skip:]


def short(name):
    m = re.match(r"(?:void )?(?:dvae::)?(\w+)(<.*>)?", name)
    base = m.group(1) if m else name
    targs = (m.group(2) or "") if m else ""
    if base == "tc_gemm_kernel":
        t = targs.replace("dvae::", "").replace("__nv_bfloat16", "bf16")
        return "tc_gemm" + t.replace("(bool)1", "MN").replace("(bool)0", "K").replace("(int)", "")
    return base + ("<tf32>" if "tf32_t" in targs else "")


agg = defaultdict(lambda: [0, 0.0])
for _, name, us, *_ in rows:
    a = agg[short(name)]
    a[0] += 1
    a[1] += us
total = sum(a[1] for a in agg.values())
print(f"launches {len(rows)}  total {total / 1e3:.3f} ms (serialised, cold cache: compare shares)")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / total * 100:6.2f}%  {us / 1e3:9.3f} ms  {n:6d} x {us / n:9.2f} us  {k}")

# optional: list every launch whose short name contains argv[3] (grid, duration) in launch order
if len(sys.argv) > 3:
    key = sys.argv[3]
    print(f"--- individual launches matching '{key}'")
    for i, name, us, grid, block in rows:
        if key in short(name):
            print(f"  #{i:5d} {us:9.2f} us  grid {grid:>14s}  {short(name)[:110]}")
