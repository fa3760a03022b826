"""Cost of keeping the pre-BatchNorm convolution output in fp32 (fp16 mode): conv5_fwd_bnstats + bn_finalize_apply +
bn_train_bwd at the step's shape, y stored as fp16 vs fp32."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
from dvae_b200 import lib, ops

dt = lib.F16
R2, T, C = 1024, 64, 512
xs = [torch.randn(R2, T, C, device="cuda").half() for _ in range(3)]
ds = [torch.randn(R2 * T, C, device="cuda").half() for _ in range(3)]
wk = (torch.randn(C, 5, C, device="cuda") * 0.02).half()
b, g, be = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
rm, rv, nb = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros((), device="cuda", dtype=torch.long)


def timeit(fn, n=12):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i % 3)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for y_f32 in (False, True):
    st = {}

    def conv(i):
        st["y"], st["ws"] = ops.conv5_fwd_bnstats(dt, xs[i], wk, b, 2, y_f32=y_f32)

    def apply(i):
        st["o"], st["s"] = ops.bn_finalize_apply(dt, st["y"].view(-1, C), st["ws"], g, be, rm, rv, nb, 2, lib.ACT_RELU, 1e-5, 0.1)
    t_c = timeit(conv)
    t_a = timeit(apply)
    t_b = timeit(lambda i: ops.bn_train_bwd(dt, ds[i], st["y"].view(-1, C), st["s"], 2, lib.ACT_RELU))
    print(f"y {'fp32' if y_f32 else 'fp16'}: conv+stats {t_c:6.1f} us   bn finalize+apply {t_a:6.1f} us   bn backward {t_b:6.1f} us")
