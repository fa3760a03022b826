"""GPU parity of the tcgen05 GEMM family (linear / conv1d-k5 / LSTM; fwd, dgrad, wgrad) against plain
PyTorch fp32 on the same (bf16- or tf32-representable) inputs.  Tolerances: the operands are exactly
representable in the storage dtype, accumulation is fp32 on both sides, so only summation order and the
rounding of the stored result differ."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTS = ["bf16", "tf32", "fp16", "fp32"]


def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _dt(name):
    from dvae_b200 import lib
    return {"bf16": lib.BF16, "tf32": lib.TF32, "fp16": lib.F16, "fp32": lib.F32}[name]


def _rand(shape, name, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randn(shape, device="cuda", generator=g) * scale
    if name == "bf16":
        return t.to(torch.bfloat16)
    if name == "fp16":
        return t.to(torch.float16)
    if name == "fp32":
        return t                      # strict mode: nothing is rounded
    # tf32: keep 10 mantissa bits so the tensor core sees exactly these values
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def _tol(name, ref):
    # result rounding to the storage dtype + fp32 summation-order noise
    # (bf16: 8 mantissa bits; tf32 storage is rounded onto the 10-bit tf32 grid when stored)
    return {"bf16": 2.0 ** -8, "fp32": 5e-6}.get(name, 2.0 ** -11) * ref.abs().max().item() + 1e-6


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


@pytest.mark.parametrize("name,y_f32", [("bf16", False), ("tf32", False), ("tf32", True), ("fp16", False), ("fp16", True), ("fp32", False)])
@pytest.mark.parametrize("R,Cin,Cout,act", [(4, 80, 512, "relu"), (16, 512, 512, "tanh"), (32, 512, 512, "relu"), (6, 512, 80, "none")])
def test_conv5_bnstats_and_bn_apply(name, y_f32, R, Cin, Cout, act, mt):
    """The training path of every Conv1d + BatchNorm1d (+ activation) pair: `conv5_fwd_bnstats` (BatchNorm statistics fused
    into the convolution's staged store epilogue when the persistent kernel runs, a separate reduction otherwise) followed by
    `bn_finalize_apply`, against F.conv1d + F.batch_norm in train mode, with TWO statistics halves (the x1 call and the x2
    call of model/disentangled_vae.py:251-254 share one tensor) and the sequential running-statistics updates."""
    _setup()
    from dvae_b200 import lib, ops
    dt = _dt(name)
    T, halves = 64, 2
    x = _rand((R, T, Cin), name, seed=21)
    w = _rand((Cout, Cin, 5), name, 0.05, seed=22)
    b = torch.randn(Cout, device="cuda")
    gamma = torch.rand(Cout, device="cuda") + 0.5
    beta = torch.randn(Cout, device="cuda") * 0.1
    wk = w.permute(0, 2, 1).contiguous()
    y, ws = ops.conv5_fwd_bnstats(dt, x, wk, b, halves, y_f32=y_f32)
    assert y.dtype == (torch.float32 if y_f32 else ops.act_dtype(dt))
    rm, rv = torch.zeros(Cout, device="cuda"), torch.ones(Cout, device="cuda")
    nbt = torch.zeros((), device="cuda", dtype=torch.long)
    code = {"relu": lib.ACT_RELU, "tanh": lib.ACT_TANH, "none": lib.ACT_NONE}[act]
    out, stat = ops.bn_finalize_apply(dt, y.view(-1, Cout), ws, gamma, beta, rm, rv, nbt, halves, code, 1e-5, 0.1)
    assert out.dtype == ops.act_dtype(dt)
    # reference: the two halves are two consecutive module calls
    rrm, rrv = torch.zeros(Cout, device="cuda"), torch.ones(Cout, device="cuda")
    refs = []
    for hsel in range(halves):
        xh = x[hsel * R // halves:(hsel + 1) * R // halves].float().transpose(1, 2)
        z = F.batch_norm(F.conv1d(xh, w.float(), b, padding=2), rrm, rrv, gamma, beta, True, 0.1, 1e-5)
        z = F.relu(z) if act == "relu" else (torch.tanh(z) if act == "tanh" else z)
        refs.append(z.transpose(1, 2))
    ref = torch.cat(refs).reshape(-1, Cout)
    # y is rounded to the storage type before normalisation unless y_f32; the output is rounded once more
    step = 2.0 ** -8 if name == "bf16" else 2.0 ** -11
    tol = (2.5 if not y_f32 else 1.2) * step * 4.0 + 1e-5            # |BN output| stays below ~4 for these inputs
    assert (out.float() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item() / 4.0)
    assert int(nbt.item()) == halves
    assert torch.allclose(rm, rrm, atol=2e-3 if name == "bf16" else 2e-4, rtol=1e-2)
    assert torch.allclose(rv, rrv, atol=2e-3 if name == "bf16" else 2e-4, rtol=1e-2)
    # backward through the pair (BatchNorm backward reads the same y, fp32 or not)
    dout = _rand((R * T, Cout), name, seed=23)
    dy, dgamma, dbeta = ops.bn_train_bwd(dt, dout, y.view(-1, Cout), stat, halves, code)
    dys, dgs, dbs = [], 0, 0
    for hsel in range(halves):
        sl = slice(hsel * R // halves, (hsel + 1) * R // halves)
        yh = (y[sl].float()).transpose(1, 2).detach().requires_grad_(True)
        g_, b_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        z = F.batch_norm(yh, None, None, g_, b_, True, 0.1, 1e-5)
        z = F.relu(z) if act == "relu" else (torch.tanh(z) if act == "tanh" else z)
        z.backward(dout.float().view(R, T, Cout)[sl].transpose(1, 2))
        dys.append(yh.grad.transpose(1, 2))
        dgs, dbs = dgs + g_.grad, dbs + b_.grad
    ref_dy = torch.cat(dys).reshape(-1, Cout)
    assert (dy.float() - ref_dy).norm().item() <= (1e-2 if name == "bf16" else 1.5e-3) * ref_dy.norm().item()
    # a ReLU input within rounding of zero may take the other branch in the reference (one element of dout moves between the
    # two sums): allow two such channels, bound everything else tightly
    for mine, ref_ in ((dgamma, dgs), (dbeta, dbs)):
        off = (mine - ref_).abs() > 2e-3 * ref_.abs() + 2e-3 * ref_.abs().max().item()
        assert int(off.sum().item()) <= 2, (mine - ref_).abs().max().item()
        assert (mine - ref_).norm().item() <= 1e-2 * ref_.norm().item()


@pytest.mark.parametrize("M,N,K", [(256, 2048, 8192), (130, 64, 2048), (64, 2048, 32)])
def test_linear_fwd_split_precision(M, N, K):
    """fp16 mode's split-precision small layers: weights as [w_hi | w_lo] (two passes over the same A tiles), the result
    optionally re-split as [hi | lo | hi] for a following GEMM against [w_hi | w_hi | w_lo].  The chain then reproduces
    the fp32 product of the fp16 activations with the FP32 weights to ~1e-6 instead of fp16's 5e-4."""
    _setup()
    from dvae_b200 import lib, ops
    dt = lib.F16
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(M, K, device="cuda", generator=g).half()
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5          # fp32 weights, NOT fp16-representable
    b = torch.randn(N, device="cuda", generator=g)
    w_split = torch.empty(N, 2 * K, device="cuda", dtype=torch.float16)
    ops.prep_cast_split(dt, w, w_split, 2)
    assert torch.equal(w_split[:, :K], w.half())
    assert (w_split[:, :K].float() + w_split[:, K:].float() - w).abs().max().item() <= 2.0 ** -21 * w.abs().max().item()
    out, out_cat = ops.linear_fwd_split(dt, x, w_split, b, relu=True, want_cat=True)
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    assert torch.equal(out, out_cat[:, :N]) and torch.equal(out, out_cat[:, 2 * N:])
    recon = out_cat[:, :N].double() + out_cat[:, N:2 * N].double()
    # hi + lo carries the fp32 result: what is left is the tensor core's fp32 accumulation over up to 16384 products (~2e-5)
    assert (recon - ref).norm().item() <= 4e-5 * ref.norm().item()
    assert (out.double() - ref).norm().item() <= 4e-4 * ref.norm().item()    # hi alone is an fp16 rounding of it
    # a following small GEMM against [v_hi | v_hi | v_lo]
    v = torch.randn(64, N, device="cuda", generator=g) / N ** 0.5
    v3 = torch.empty(64, 3 * N, device="cuda", dtype=torch.float16)
    ops.prep_cast_split(dt, v, v3, 3)
    _, y32 = ops.linear_fwd(dt, out_cat, v3, None, want_f32=True, want_act=False)
    ref2 = ref @ v.double().t()
    assert (y32.double() - ref2).norm().item() <= 6e-5 * ref2.norm().item()
    plain = out.double() @ v.half().double().t()
    assert (plain - ref2).norm().item() > 5 * (y32.double() - ref2).norm().item()      # what the split buys


def test_pack_split_and_first_conv():
    """[hi | lo | hi] input against [w_hi | w_hi | w_lo] weights: the first encoder convolution (80 -> 512) sees its fp32 mel
    input and fp32 weights to ~1e-6."""
    _setup()
    from dvae_b200 import lib, ops
    dt = lib.F16
    R, T = 6, 64
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.rand(R, 80, T, device="cuda", generator=g)
    w = (torch.rand(512, 80, 5, device="cuda", generator=g) - 0.5) * 0.2
    b = torch.randn(512, device="cuda", generator=g) * 0.1
    x_cl = torch.empty(R, T, 80, device="cuda", dtype=torch.float16)
    x_cat = torch.empty(R, T, 240, device="cuda", dtype=torch.float16)
    ops.pack_ncl_to_cl(dt, x, x_cl, x_cat)
    assert torch.equal(x_cl, x.transpose(1, 2).half()) and torch.equal(x_cat[..., :80], x_cl) and torch.equal(x_cat[..., 160:], x_cl)
    assert (x_cat[..., :80].float() + x_cat[..., 80:160].float() - x.transpose(1, 2)).abs().max().item() <= 2.0 ** -21
    wk = torch.empty(512, 5, 80, device="cuda", dtype=torch.float16)
    wk_cat = torch.empty(512, 5, 240, device="cuda", dtype=torch.float16)
    ops.prep_conv_weight(dt, w, out=wk, out_cat=wk_cat)
    assert torch.equal(wk, w.permute(0, 2, 1).half()) and torch.equal(wk_cat[..., :80], wk) and torch.equal(wk_cat[..., 80:160], wk)
    y, _ = ops.conv5_fwd_bnstats(dt, x_cat, wk_cat, b, 2, y_f32=True)
    ref = F.conv1d(x.double(), w.double(), b.double(), padding=2).transpose(1, 2)
    assert (y.double() - ref).norm().item() <= 1e-5 * ref.norm().item()
    y_plain, _ = ops.conv5_fwd_bnstats(dt, x_cl, wk, b, 2, y_f32=True)
    assert (y_plain.double() - ref).norm().item() > 10 * (y.double() - ref).norm().item()


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
@pytest.mark.parametrize("rows,H,D,T", [(8, 64, 2, 64), (130, 64, 1, 64), (16, 512, 1, 64), (256, 1024, 1, 64),
                                        # rows % 512 == 0: the time-resident kernels of ops_lstm_res.cu (16-bit storage)
                                        (512, 1024, 1, 64), (1024, 512, 1, 64)])
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
    tol = {"bf16": 3e-2, "fp32": 2e-5}.get(name, 2e-3)       # h is re-rounded to the storage dtype every step (64 steps)
    assert err <= tol, f"lstm fwd max err {err}"
    # backward: grad wrt the natural-order x-projection is exactly da_all
    dh = _rand((rows, T, D * H), name, seed=13)
    ref.backward(dh.float())
    da = ops.lstm_bwd(dt, dh, xg, c_all, whh.contiguous(), H, D)
    ref_da = xr.grad.reshape(rows, T, D * 4 * H)
    rel = (da.float() - ref_da).norm().item() / ref_da.norm().item()
    assert rel <= {"bf16": 3e-2, "fp32": 2e-5}.get(name, 3e-3), f"lstm bwd rel err {rel}"
    dwhh = torch.zeros((D, 4 * H, H), device="cuda")
    ops.lstm_wgrad_hh(dt, da, h_all, dwhh, H, D)
    relw = (dwhh - wr.grad).norm().item() / wr.grad.norm().item()
    assert relw <= {"bf16": 3e-2, "fp32": 2e-5}.get(name, 3e-3), f"lstm dW_hh rel err {relw}"


def test_lstm_resident_forward_is_bit_identical():
    """The time-resident forward runs the same MMAs in the same order as the step-per-launch kernels: h, c and the saved gates
    are bit-identical, at both hidden sizes, with one and with two launches per call (rows > 1024)."""
    _setup()
    from dvae_b200 import lib, ops
    prev = lib.set_lstm_resident(-1)
    try:
        for name, H, rows, T in (("fp16", 1024, 1024, 64), ("bf16", 512, 2048, 16), ("fp16", 512, 512, 5)):
            dt = _dt(name)
            xg0 = _rand((rows, T, 4 * H), name, seed=21)
            whh_p = _rand((1, 4 * H, H), name, 1.0 / H ** 0.5, seed=22)
            out = {}
            for mode in (0, 1):
                lib.set_lstm_resident(mode)
                xg = xg0.clone()
                h, c = ops.lstm_fwd(dt, xg, whh_p, H, 1)
                torch.cuda.synchronize()
                out[mode] = (h, c, xg)
            for a, b, nm in zip(out[0], out[1], ("h", "c", "gates")):
                assert torch.equal(a, b), f"{name} H={H} rows={rows}: {nm} differs between the resident and the per-step kernels"
    finally:
        lib.set_lstm_resident(prev)


@pytest.mark.parametrize("env", [
    {"DVAE_LSTM_SEQ": "0"},                                        # H = 64 through the step-per-launch path
    {"DVAE_LSTM_SEQ_IO": "direct"},                                # sequence-resident kernels without the TMA staging
    {"DVAE_LSTM_FWD_TMA": "0", "DVAE_LSTM_BWD_REDUCE": "0"},       # direct cell epilogue, store-epilogue backward
    {"DVAE_LSTM_PAIR": "1"},                                       # CTA-pair (cta_group::2) step kernels
    {"DVAE_LSTM_BWD_FUSED": "1"},                                  # cell backward fused into the step GEMM's epilogue
    {"DVAE_LSTM_RES": "0"},                                        # H = 512 / 1024 through the step-per-launch path at every size
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
