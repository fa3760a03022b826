"""Time-resident H = 512 / 1024 recurrence kernels (csrc/ops_lstm_res.cu) against the step-per-launch kernels on the same
inputs: forward outputs must be bit-identical (same MMA order), backward within fp32 summation-order noise; then us per step
of both.  `python scripts/check_lstm_res.py [fwd|bwd|all]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import lib, ops

what = sys.argv[1] if len(sys.argv) > 1 else "all"
T = int(os.environ.get("T", "64"))


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


for name, dt, td in (("fp16", lib.F16, torch.float16), ("bf16", lib.BF16, torch.bfloat16)):
    for H, rows in ((1024, 1024), (512, 1024), (1024, 512), (512, 2048)):
        g = torch.Generator(device="cuda").manual_seed(H + rows)
        xg0 = torch.randn(rows, T, 4 * H, device="cuda", generator=g).to(td)
        whh = (torch.randn(1, 4 * H, H, device="cuda", generator=g) / H ** 0.5).to(td)
        perm = ((torch.arange(4 * H) % 4) * H + torch.arange(4 * H) // 4).cuda()
        whh_p = whh[:, perm, :].contiguous()
        dh = (torch.randn(rows, T, H, device="cuda", generator=g) * 0.1).to(td)
        res = {}
        for mode in (0, 1):
            lib.set_lstm_resident(mode)
            xg = xg0.clone()
            h, c = ops.lstm_fwd(dt, xg, whh_p, H, 1)
            torch.cuda.synchronize()
            res[mode] = [h, c, xg]
            if what in ("bwd", "all"):
                da = ops.lstm_bwd(dt, dh, xg, c, whh, H, 1)
                torch.cuda.synchronize()
                res[mode].append(da)
        msg = f"{name} H={H} rows={rows}: "
        for i, nm in enumerate(("h", "c", "gates")):
            a, b = res[0][i].float(), res[1][i].float()
            msg += f"{nm} maxdiff {(a - b).abs().max().item():.3e} (neq {(a != b).sum().item()})  "
        if what in ("bwd", "all"):
            a, b = res[0][3].float(), res[1][3].float()
            msg += f"da rel {((a - b).norm() / a.norm()).item():.3e} max {(a - b).abs().max().item():.3e}"
        print(msg, flush=True)
        for mode in (0, 1):
            lib.set_lstm_resident(mode)
            xg = xg0.clone()
            t_copy = timeit(lambda: xg.copy_(xg0))
            keep = {}

            def fwd():
                xg.copy_(xg0)
                keep["h"], keep["c"] = ops.lstm_fwd(dt, xg, whh_p, H, 1)
            t_f = timeit(fwd) - t_copy
            line = f"    resident={mode}: fwd {t_f * 1e3 / T:7.2f} us/step"
            if what in ("bwd", "all"):
                t_b = timeit(lambda: ops.lstm_bwd(dt, dh, xg, keep["c"], whh, H, 1))
                line += f"   bwd {t_b * 1e3 / T:7.2f} us/step"
            print(line, flush=True)
        del res
lib.set_lstm_resident(1)
