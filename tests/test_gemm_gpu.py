"""GPU parity of the tcgen05 GEMM family (linear / conv1d-k5 / LSTM; fwd, dgrad, wgrad) against plain
PyTorch fp32 on the same (bf16- or tf32-representable) inputs.  Tolerances: the operands are exactly
representable in the storage dtype, accumulation is fp32 on both sides, so only summation order and the
rounding of the stored result differ."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTS = ["bf16", "tf32", "fp16"]


def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _dt(name):
    from dvae_b200 import lib
    return {"bf16": lib.BF16, "tf32": lib.TF32, "fp16": lib.F16}[name]


def _rand(shape, name, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randn(shape, device="cuda", generator=g) * scale
    if name == "bf16":
        return t.to(torch.bfloat16)
    if name == "fp16":
        return t.to(torch.float16)
    # tf32: keep 10 mantissa bits so the tensor core sees exactly these values
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def _tol(name, ref):
    # result rounding to the storage dtype + fp32 summation-order noise
    # (bf16: 8 mantissa bits; tf32 storage is rounded onto the 10-bit tf32 grid when stored)
    return (2.0 ** -8 if name == "bf16" else 2.0 ** -11) * ref.abs().max().item() + 1e-6


@pytest.fixture(params=["1", "2", "p"], ids=["rows128", "rows256", "persistent"])
def mt(request, monkeypatch):
    """GEMM kernel variant: one tile per CTA with a 128-row or a 256-row (two accumulators sharing B) tile, or the
    persistent kernel (one CTA per SM, double-buffered TMEM accumulators)."""
    if request.param == "p":
        monkeypatch.setenv("DVAE_GEMM_MT", "1")
        monkeypatch.setenv("DVAE_GEMM_PERSISTENT", "1")
    else:
        monkeypatch.setenv("DVAE_GEMM_MT", request.param)
        monkeypatch.setenv("DVAE_GEMM_PERSISTENT", "0")
    return request.param


@pytest.mark.parametrize("name", DTS)
@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 0), (200, 80, 96, 0), (256, 256, 512, 0), (256, 256, 512, 128),
                                      (256, 256, 512, 64), (1024, 2048, 1024, 0), (130, 56, 32, 0), (64, 4096, 128, 128)])
def test_linear_fwd(name, M, N, K, bn, mt):
    _setup()
    from dvae_b200 import ops
    dt = _dt(name)
    x, w = _rand((M, K), name, seed=1), _rand((N, K), name, 0.1, seed=2)
    b = torch.randn(N, device="cuda")
    out, out32 = ops.linear_fwd(dt, x, w, b, relu=True, want_f32=True, block_n=bn)
    ref = torch.relu(x.float() @ w.float().t() + b)
    assert (out32 - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-5
    assert (out.float() - ref).abs().max().item() <= _tol(name, ref)


@pytest.mark.parametrize("name", DTS)
@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 0), (200, 80, 96, 0), (256, 512, 256, 0), (256, 512, 256, 128),
                                      (512, 2048, 8192 // 8, 0), (130, 2048, 32, 0), (300, 64, 2048, 0)])
def test_linear_dgrad(name, M, N, K, bn, mt):
    _setup()
    from dvae_b200 import ops
    dt = _dt(name)
    dy, w = _rand((M, N), name, seed=3), _rand((N, K), name, 0.1, seed=4)
    mask = _rand((M, K), name, seed=5)
    dx, dx32 = ops.linear_dgrad(dt, dy, w, relu_mask=mask, want_f32=True, block_n=bn)
    ref = (dy.float() @ w.float()) * (mask.float() > 0)
    assert (dx32 - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-5
    assert (dx.float() - ref).abs().max().item() <= _tol(name, ref)


@pytest.mark.parametrize("name", DTS)
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 80, 96), (4096, 256, 512), (1024, 2048, 32), (777, 64, 2048)])
def test_linear_wgrad(name, M, N, K, mt):
    _setup()
    from dvae_b200 import ops
    dt = _dt(name)
    dy, x = _rand((M, N), name, seed=6), _rand((M, K), name, seed=7)
    dw = torch.zeros((N, K), device="cuda")
    ops.linear_wgrad(dt, dy, x, dw)
    ref = dy.float().t() @ x.float()
    assert (dw - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-5
    ops.linear_wgrad(dt, dy, x, dw)   # accumulates
    assert (dw - 2 * ref).abs().max().item() <= 4e-5 * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("name", DTS)
@pytest.mark.parametrize("R,Cin,Cout", [(2, 64, 64), (3, 80, 512), (8, 512, 512), (5, 512, 80), (16, 512, 512)])
def test_conv5(name, R, Cin, Cout, mt):
    _setup()
    from dvae_b200 import ops
    dt = _dt(name)
    T = 64
    x = _rand((R, T, Cin), name, seed=8)
    w = _rand((Cout, Cin, 5), name, 0.05, seed=9)        # torch layout [Cout, Cin, k]
    b = torch.randn(Cout, device="cuda")
    wk = w.permute(0, 2, 1).contiguous()                  # [Cout, 5, Cin]
    y, y32 = ops.conv5_fwd(dt, x, wk, b, want_f32=True)
    xt = x.float().transpose(1, 2).requires_grad_(True)   # NCL
    wt = w.float().requires_grad_(True)
    ref = F.conv1d(xt, wt, b, padding=2)
    ref_cl = ref.transpose(1, 2)
    assert (y32 - ref_cl).abs().max().item() <= 1e-5 * ref_cl.abs().max().item() + 1e-5
    assert (y.float() - ref_cl).abs().max().item() <= _tol(name, ref_cl)
    dy = _rand((R, T, Cout), name, seed=10)
    ref.backward(dy.float().transpose(1, 2))
    dx, dx32 = ops.conv5_dgrad(dt, dy, wk, want_f32=True)
    ref_dx = xt.grad.transpose(1, 2)
    assert (dx32 - ref_dx).abs().max().item() <= 1e-5 * ref_dx.abs().max().item() + 1e-5
    dwk = torch.zeros((Cout, 5, Cin), device="cuda")
    ops.conv5_wgrad(dt, dy, x, dwk)
    ref_dw = wt.grad.permute(0, 2, 1)
    assert (dwk - ref_dw).abs().max().item() <= 2e-5 * ref_dw.abs().max().item() + 1e-5


def _gate_perm(H, tile=None):
    n = torch.arange(4 * H)
    return (n % 4) * H + n // 4            # dst row n = 4*u + g  <-  src row g*H + u


def _lstm_ref(xproj_nat, whh, D, H):
    """xproj_nat [rows,T,D,4H] natural gate order, whh [D,4H,H] -> h_all [rows,T,D*H] (fp32 torch)."""
    rows, T = xproj_nat.shape[:2]
    outs = []
    for d in range(D):
        h = xproj_nat.new_zeros(rows, H)
        c = xproj_nat.new_zeros(rows, H)
        ys = [None] * T
        for t in (range(T) if d == 0 else range(T - 1, -1, -1)):
            a = xproj_nat[:, t, d] + h @ whh[d].t()
            i, f, g, o = a.split(H, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            ys[t] = h
        outs.append(torch.stack(ys, 1))
    return torch.cat(outs, -1)


@pytest.mark.parametrize("name", DTS)
@pytest.mark.parametrize("rows,H,D,T", [(8, 64, 2, 64), (130, 64, 1, 64), (16, 512, 1, 64), (256, 1024, 1, 64)])
def test_lstm(name, rows, H, D, T):
    _setup()
    from dvae_b200 import lib, ops
    dt = _dt(name)
    tile = lib.lstm_gate_tile(H)
    perm = _gate_perm(H, tile).cuda()
    xproj = _rand((rows, T, D, 4 * H), name, seed=11)                  # natural gate order
    whh = _rand((D, 4 * H, H), name, 1.0 / H ** 0.5, seed=12)
    xg = xproj[:, :, :, perm].reshape(rows, T, D * 4 * H).contiguous()
    whh_p = whh[:, perm, :].contiguous()
    h_all, c_all = ops.lstm_fwd(dt, xg, whh_p, H, D)
    xr = xproj.float().requires_grad_(True)
    wr = whh.float().requires_grad_(True)
    ref = _lstm_ref(xr, wr, D, H)
    err = (h_all.float() - ref).abs().max().item()
    tol = 3e-2 if name == "bf16" else 2e-3       # h is re-rounded to the storage dtype every step (64 steps)
    assert err <= tol, f"lstm fwd max err {err}"
    # backward: grad wrt the natural-order x-projection is exactly da_all
    dh = _rand((rows, T, D * H), name, seed=13)
    ref.backward(dh.float())
    da = ops.lstm_bwd(dt, dh, xg, c_all, whh.contiguous(), H, D)
    ref_da = xr.grad.reshape(rows, T, D * 4 * H)
    rel = (da.float() - ref_da).norm().item() / ref_da.norm().item()
    assert rel <= (3e-2 if name == "bf16" else 3e-3), f"lstm bwd rel err {rel}"
    dwhh = torch.zeros((D, 4 * H, H), device="cuda")
    ops.lstm_wgrad_hh(dt, da, h_all, dwhh, H, D)
    relw = (dwhh - wr.grad).norm().item() / wr.grad.norm().item()
    assert relw <= (3e-2 if name == "bf16" else 3e-3), f"lstm dW_hh rel err {relw}"


@pytest.mark.parametrize("env", [
    {"DVAE_LSTM_SEQ": "0"},                                        # H = 64 through the step-per-launch path
    {"DVAE_LSTM_SEQ_IO": "direct"},                                # sequence-resident kernels without the TMA staging
    {"DVAE_LSTM_FWD_TMA": "0", "DVAE_LSTM_BWD_REDUCE": "0"},       # direct cell epilogue, store-epilogue backward
    {"DVAE_LSTM_PAIR": "1"},                                       # CTA-pair (cta_group::2) step kernels
    {"DVAE_LSTM_BWD_FUSED": "1"},                                  # cell backward fused into the step GEMM's epilogue
], ids=lambda e: ",".join(f"{k[5:]}={v}" for k, v in e.items()))
def test_lstm_optional_paths(env):
    """The kernel variants that are not the default (selected by environment variables read once per process) stay correct:
    re-run the LSTM parity test in a child process per variant."""
    import subprocess
    import sys
    child_env = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k", "test_lstm and not optional",
                        "-p", "no:cacheprovider"], env=child_env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("env", [
    {"DVAE_GEMM_PERSISTENT": "1"},                                 # persistent CTA-pair kernels on the small test shapes
    {"DVAE_GEMM_PERSISTENT": "1", "DVAE_GEMM_PAIR": "0"},          # persistent 1-SM kernels
    {"DVAE_GEMM_PERSISTENT": "1", "DVAE_GEMM_TMA_STORE": "1"},     # staged TMA-store epilogue wherever eligible
    {"DVAE_GEMM_PERSISTENT": "0", "DVAE_WGRAD_WIDE": "0"},         # one tile per CTA only
], ids=lambda e: ",".join(f"{k[5:]}={v}" for k, v in e.items()))
def test_gemm_optional_paths(env):
    """Same for the dense-contraction variants (linear / conv forward, dgrad, wgrad)."""
    import subprocess
    import sys
    child_env = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k", "not lstm and not optional",
                        "-p", "no:cacheprovider"], env=child_env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
