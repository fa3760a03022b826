"""Tensor-level wrappers over the C-ABI (one Python function per exported op).

These allocate outputs with torch (device memory / stream plumbing only) and pass raw pointers to
libdvae_b200.so.  No arithmetic happens in Python or torch here.
Activation storage dtype `dt` is lib.BF16 (torch.bfloat16, tcgen05 kind::f16), lib.F16 (torch.float16, kind::f16) or
lib.TF32 (torch.float32 storage, tcgen05 kind::tf32); lib.F32 is the strict mode (fp32 on the CUDA cores, for the 1e-5 check).  `alpha` / `scale` arguments carry the power-of-two gradient
scale of the fp16 mode (1.0 otherwise).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import lib
from .lib import call, ptr, stream

Tensor = torch.Tensor


_ACT_DTYPES = {lib.BF16: torch.bfloat16, lib.F16: torch.float16, lib.TF32: torch.float32, lib.F32: torch.float32}


def act_dtype(dt: int) -> torch.dtype:
    return _ACT_DTYPES[dt]


def _chk(t: Tensor, dtype=None):
    assert t.is_cuda and t.is_contiguous(), "dvae_b200 ops need contiguous CUDA tensors"
    if dtype is not None:
        assert t.dtype == dtype, f"expected {dtype}, got {t.dtype}"
    return t


# ------------------------------------------------------------------------------- dense ops
def linear_fwd(dt: int, x: Tensor, w: Tensor, bias: Optional[Tensor], relu: bool = False, want_f32: bool = False,
               want_act: bool = True, block_n: int = 0) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """out[M,N] = x[M,K] @ w[N,K]^T (+bias)(relu).  Returns (out_act, out_f32)."""
    ad = act_dtype(dt)
    _chk(x, ad), _chk(w, ad)
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K
    out = torch.empty((M, N), device=x.device, dtype=ad) if want_act else None
    out32 = torch.empty((M, N), device=x.device, dtype=torch.float32) if want_f32 else None
    call("dvae_linear_fwd", dt, ptr(x), K, ptr(w), ptr(bias), ptr(out), ptr(out32), N, M, N, K, int(relu), block_n, stream())
    return out, out32


def linear_fwd_split(dt: int, x: Tensor, w_split: Tensor, bias: Optional[Tensor], relu: bool = False, want_cat: bool = False):
    """Split-precision forward of a small linear layer: w_split [N, 2K] = [w_hi | w_lo] (prep_cast_split), out = x @ (w_hi +
    w_lo)^T (+bias)(relu).  Returns (out act [M,N], out_cat act [M,3N] = [hi | lo | hi] or None)."""
    ad = act_dtype(dt)
    _chk(x, ad), _chk(w_split, ad)
    M, K = x.shape
    N = w_split.shape[0]
    assert w_split.shape[1] == 2 * K
    out = torch.empty((M, N), device=x.device, dtype=ad)
    out_cat = torch.empty((M, 3 * N), device=x.device, dtype=ad) if want_cat else None
    call("dvae_linear_fwd_split", dt, ptr(x), K, ptr(w_split), ptr(bias), ptr(out), None, N, ptr(out_cat), M, N, K, int(relu), 2, stream())
    return out, out_cat


def prep_cast_split(dt: int, src: Tensor, dst: Tensor, parts: int) -> None:
    """dst act [rows, parts*K] = [hi | lo] (parts 2) or [hi | hi | lo] (parts 3) of src fp32 [rows, K]."""
    _chk(src, torch.float32), _chk(dst, act_dtype(dt))
    rows, K = src.shape
    assert tuple(dst.shape) == (rows, parts * K)
    call("dvae_prep_cast_split", dt, ptr(src), ptr(dst), rows, K, parts, stream())


def linear_dgrad(dt: int, dy: Tensor, w: Tensor, relu_mask: Optional[Tensor] = None, want_f32: bool = False,
                 want_act: bool = True, block_n: int = 0) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """dx[M,K] = dy[M,N] @ w[N,K]; optionally zeroed where relu_mask <= 0."""
    ad = act_dtype(dt)
    _chk(dy, ad), _chk(w, ad)
    M, N = dy.shape
    K = w.shape[1]
    assert w.shape[0] == N
    dx = torch.empty((M, K), device=dy.device, dtype=ad) if want_act else None
    dx32 = torch.empty((M, K), device=dy.device, dtype=torch.float32) if want_f32 else None
    call("dvae_linear_dgrad", dt, ptr(dy), N, ptr(w), ptr(dx), ptr(dx32), ptr(relu_mask), K, M, N, K, block_n, stream())
    return dx, dx32


def linear_wgrad(dt: int, dy: Tensor, x: Tensor, dw: Tensor, alpha: float = 1.0) -> None:
    """dw[N,K] (fp32, pre-zeroed or accumulating) += alpha * dy[M,N]^T @ x[M,K]."""
    ad = act_dtype(dt)
    _chk(dy, ad), _chk(x, ad), _chk(dw, torch.float32)
    M, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == M and tuple(dw.shape) == (N, K)
    call("dvae_linear_wgrad", dt, ptr(dy), N, ptr(x), K, ptr(dw), K, M, N, K, float(alpha), stream())


def conv5_fwd(dt: int, x: Tensor, wk: Tensor, bias: Optional[Tensor], want_f32: bool = False):
    """Channels-last Conv1d(k=5,pad=2): x [R,T,Cin], wk [Cout,5,Cin] -> y [R,T,Cout] (+bias)."""
    ad = act_dtype(dt)
    _chk(x, ad), _chk(wk, ad)
    R, T, Cin = x.shape
    Cout = wk.shape[0]
    assert tuple(wk.shape) == (Cout, 5, Cin)
    y = torch.empty((R, T, Cout), device=x.device, dtype=ad)
    y32 = torch.empty((R, T, Cout), device=x.device, dtype=torch.float32) if want_f32 else None
    call("dvae_conv5_fwd", dt, ptr(x), ptr(wk), ptr(bias), ptr(y), ptr(y32), R, T, Cin, Cout, stream())
    return (y, y32) if want_f32 else y


def _y_is_f32(dt: int, y: Tensor) -> int:
    """1 when y is unrounded fp32 although the activation dtype is a 16-bit one (the fp16 mode's pre-BatchNorm tensors)."""
    ad = act_dtype(dt)
    if y.dtype == ad:
        return 0
    assert y.dtype == torch.float32, f"expected {ad} or float32, got {y.dtype}"
    return 1


def conv5_fwd_bnstats(dt: int, x: Tensor, wk: Tensor, bias: Optional[Tensor], halves: int, y_f32: bool = False):
    """conv5_fwd + the statistics pass of the train-mode BatchNorm behind it.  Returns (y, ws): ws holds the per-half
    column sums / sums of squares of y (double [halves*2*Cout + 1]) for `bn_finalize_apply`.  y_f32: store y as
    unrounded fp32 (fp16 mode: fp32 instead of fp16 storage; tf32 mode: the fp32 storage is not rounded to the tf32 grid --
    y feeds BatchNorm, not a tensor-core operand)."""
    ad = act_dtype(dt)
    _chk(x, ad), _chk(wk, ad)
    R, T, Cin = x.shape
    Cout = wk.shape[0]
    assert tuple(wk.shape) == (Cout, 5, Cin) and (R * T) % halves == 0
    y = torch.empty((R, T, Cout), device=x.device, dtype=torch.float32 if y_f32 else ad)
    ws = torch.empty((lib.workspace_bytes("bn_stats", halves, Cout) // 8,), device=x.device, dtype=torch.float64)
    call("dvae_conv5_fwd_bnstats", dt, ptr(x), ptr(wk), ptr(bias), ptr(y), int(bool(y_f32)), R, T, Cin, Cout, ptr(ws),
         R * T // halves, halves, stream())
    return y, ws


def conv5_dgrad(dt: int, dy: Tensor, wk: Tensor, want_f32: bool = False):
    ad = act_dtype(dt)
    _chk(dy, ad), _chk(wk, ad)
    R, T, Cout = dy.shape
    Cin = wk.shape[2]
    assert tuple(wk.shape) == (Cout, 5, Cin)
    dx = torch.empty((R, T, Cin), device=dy.device, dtype=ad)
    dx32 = torch.empty((R, T, Cin), device=dy.device, dtype=torch.float32) if want_f32 else None
    call("dvae_conv5_dgrad", dt, ptr(dy), ptr(wk), ptr(dx), ptr(dx32), R, T, Cin, Cout, stream())
    return (dx, dx32) if want_f32 else dx


def conv5_wgrad(dt: int, dy: Tensor, x: Tensor, dwk: Tensor, alpha: float = 1.0) -> None:
    """dwk[Cout,5,Cin] (fp32, accumulating) += alpha * sum_{r,t} dy[r,t,co] x[r,t+k-2,ci]."""
    ad = act_dtype(dt)
    _chk(dy, ad), _chk(x, ad), _chk(dwk, torch.float32)
    R, T, Cout = dy.shape
    Cin = x.shape[2]
    assert tuple(dwk.shape) == (Cout, 5, Cin)
    call("dvae_conv5_wgrad", dt, ptr(dy), ptr(x), ptr(dwk), R, T, Cin, Cout, float(alpha), stream())


def lstm_fwd(dt: int, xg: Tensor, whh_p: Tensor, H: int, D: int):
    """Recurrence over T steps.  xg [rows,T,D*4H] holds the (gate-interleaved) input projection on entry and the
    activated gates on exit.  Returns (h_all [rows,T,D*H] act, c_all [rows,T,D*H] fp32)."""
    ad = act_dtype(dt)
    _chk(xg, ad), _chk(whh_p, ad)
    rows, T, _ = xg.shape
    assert xg.shape[2] == D * 4 * H and tuple(whh_p.shape) == (D, 4 * H, H)
    h_all = torch.empty((rows, T, D * H), device=xg.device, dtype=ad)
    c_all = torch.empty((rows, T, D * H), device=xg.device, dtype=torch.float32)
    call("dvae_lstm_fwd", dt, ptr(xg), ptr(whh_p), ptr(h_all), ptr(c_all), rows, T, H, D, stream())
    return h_all, c_all


def lstm_bwd(dt: int, dh_all: Tensor, gates: Tensor, c_all: Tensor, whh_n: Tensor, H: int, D: int) -> Tensor:
    """Back-propagation through time.  Returns da_all [rows,T,D*4H] (natural i,f,g,o order per direction)."""
    ad = act_dtype(dt)
    _chk(dh_all, ad), _chk(gates, ad), _chk(c_all, torch.float32), _chk(whh_n, ad)
    rows, T, _ = dh_all.shape
    da = torch.empty((rows, T, D * 4 * H), device=dh_all.device, dtype=ad)
    dc = torch.empty((lib.workspace_bytes("lstm_bwd_dc", rows, H, D) // 4,), device=dh_all.device,
                     dtype=torch.float32)   # [2][D][rows][H]: dc carry, dh_rec scratch
    ws, tickets = _splitk_workspace(dt, rows, H, D, dh_all.device)
    call("dvae_lstm_bwd", dt, ptr(dh_all), ptr(gates), ptr(c_all), ptr(whh_n), ptr(da), ptr(dc), ptr(ws), ptr(tickets), rows, T,
         H, D, stream())
    return da


_SPLITK_CACHE = {}


def _splitk_workspace(dt: int, rows: int, H: int, D: int, device):
    """Fix-up workspace + self-resetting tickets of the split-K LSTM backward step, cached per shape and device."""
    key = (dt, rows, H, D, str(device))
    if key not in _SPLITK_CACHE:
        n_ws, n_t = lib.lstm_bwd_workspace(dt, rows, H, D)
        ws = torch.empty((n_ws,), device=device, dtype=torch.float32) if n_ws > 0 else None
        tickets = torch.zeros((max(n_t, 1),), device=device, dtype=torch.int32)
        _SPLITK_CACHE[key] = (ws, tickets)
    return _SPLITK_CACHE[key]


def lstm_wgrad_hh(dt: int, da_all: Tensor, h_all: Tensor, dwhh: Tensor, H: int, D: int, alpha: float = 1.0) -> None:
    ad = act_dtype(dt)
    _chk(da_all, ad), _chk(h_all, ad), _chk(dwhh, torch.float32)
    rows, T, _ = da_all.shape
    assert tuple(dwhh.shape) == (D, 4 * H, H)
    call("dvae_lstm_wgrad_hh", dt, ptr(da_all), ptr(h_all), ptr(dwhh), rows, T, H, D, float(alpha), stream())


# ------------------------------------------------------------------------------- weight preparation
def copy_f32(src: Tensor, dst: Tensor) -> None:
    """fp32 device copy through the library (used to scatter gradient slices into bucket views)."""
    _chk(src, torch.float32), _chk(dst, torch.float32)
    assert src.numel() == dst.numel()
    call("dvae_copy_f32", ptr(src), ptr(dst), src.numel(), stream())


def prep_cast(dt: int, src: Tensor, dst: Tensor, scale: float = 1.0) -> None:
    """dst (activation dtype) = scale * src (fp32)."""
    _chk(src, torch.float32), _chk(dst, act_dtype(dt))
    assert src.numel() == dst.numel()
    call("dvae_prep_cast", dt, ptr(src), ptr(dst), src.numel(), float(scale), stream())


PREP_CAST, PREP_CONV, PREP_LSTM_W, PREP_COPY, PREP_LSTM_BIAS = 0, 1, 2, 3, 4   # PrepKind of csrc/ops_pointwise.cu


class PrepTable:
    """Device-side table for dvae_prep_all: every tensor-core copy of the parameters re-derived in ONE launch.
    entries: (kind, src, src2, dst0, dst1, d0, d1) with fp32 sources; built once, valid while the tensors stay where they are."""

    CHUNK = 8192

    def __init__(self, dt: int, entries, device):
        import struct
        self.dt = dt
        self.keep = entries                      # keeps the tensors alive
        self.ptrs = tuple(t.data_ptr() for e in entries for t in e[1:5] if t is not None)
        raw = bytearray()
        blk_desc, blk_off = [], []
        for i, (kind, src, src2, dst0, dst1, d0, d1) in enumerate(entries):
            _chk(src, torch.float32)
            n = src.numel()
            raw += struct.pack("<QQQQqiiii", ptr(src), ptr(src2) or 0, ptr(dst0), ptr(dst1) or 0, n, kind, d0, d1, 0)
            for off in range(0, n, self.CHUNK):
                blk_desc.append(i)
                blk_off.append(off)
        assert len(raw) == 56 * len(entries)
        self.descs = torch.frombuffer(raw, dtype=torch.uint8).to(device)
        self.blk_desc = torch.tensor(blk_desc, dtype=torch.int32, device=device)
        self.blk_off = torch.tensor(blk_off, dtype=torch.int64, device=device)
        self.nblocks = len(blk_desc)

    def valid(self) -> bool:
        return self.ptrs == tuple(t.data_ptr() for e in self.keep for t in e[1:5] if t is not None)

    def run(self) -> None:
        call("dvae_prep_all", self.dt, ptr(self.descs), ptr(self.blk_desc), ptr(self.blk_off), self.nblocks, self.CHUNK, stream())


def add_f32_act(dt: int, a: Tensor, b: Tensor) -> Tensor:
    """fp32 out = a (fp32) + b (activation dtype), same shape."""
    _chk(a, torch.float32), _chk(b, act_dtype(dt))
    out = torch.empty_like(a)
    call("dvae_add_f32_act", dt, ptr(a), ptr(b), ptr(out), a.numel(), stream())
    return out


def add_inplace(dt: int, a: Tensor, b: Tensor) -> None:
    _chk(a, act_dtype(dt)), _chk(b, act_dtype(dt))
    assert a.numel() == b.numel()
    call("dvae_add_inplace", dt, ptr(a), ptr(b), a.numel(), stream())


def prep_conv_weight(dt: int, w: Tensor, out: Optional[Tensor] = None, out_cat: Optional[Tensor] = None) -> Tensor:
    """torch Conv1d weight [Co,Ci,5] fp32 -> [Co,5,Ci] activation dtype (+ out_cat [Co,5,3Ci] = [hi | hi | lo])."""
    _chk(w, torch.float32)
    Co, Ci, k = w.shape
    assert k == 5
    wk = out if out is not None else torch.empty((Co, 5, Ci), device=w.device, dtype=act_dtype(dt))
    _chk(wk, act_dtype(dt))
    assert tuple(wk.shape) == (Co, 5, Ci)
    if out_cat is not None:
        _chk(out_cat, act_dtype(dt))
        assert tuple(out_cat.shape) == (Co, 5, 3 * Ci)
    call("dvae_prep_conv_weight", dt, ptr(w), ptr(wk), ptr(out_cat), Co, Ci, stream())
    return wk


def conv_wgrad_unpack(dwk: Tensor, out: Optional[Tensor] = None) -> Tensor:
    _chk(dwk, torch.float32)
    Co, k, Ci = dwk.shape
    dw = out if out is not None else torch.empty((Co, Ci, 5), device=dwk.device, dtype=torch.float32)
    _chk(dw, torch.float32)
    assert tuple(dw.shape) == (Co, Ci, 5)
    call("dvae_conv_wgrad_unpack", ptr(dwk), ptr(dw), Co, Ci, stream())
    return dw


def prep_lstm_weight(dt: int, w: Tensor, dst: Tensor, H: int, tile: int) -> None:
    """w [4H, In] fp32 (torch gate order) -> dst [4H, In] activation dtype, rows gate-interleaved per `tile` columns."""
    _chk(w, torch.float32), _chk(dst, act_dtype(dt))
    In = w.shape[1]
    assert tuple(w.shape) == (4 * H, In) and tuple(dst.shape) == (4 * H, In)
    call("dvae_prep_lstm_weight", dt, ptr(w), ptr(dst), H, In, tile, stream())


def prep_lstm_bias(b_ih: Tensor, b_hh: Tensor, dst: Tensor, H: int, tile: int) -> None:
    _chk(b_ih, torch.float32), _chk(b_hh, torch.float32), _chk(dst, torch.float32)
    call("dvae_prep_lstm_bias", ptr(b_ih), ptr(b_hh), ptr(dst), H, tile, stream())


# ------------------------------------------------------------------------------- layout
def pack_ncl_to_cl(dt: int, x: Tensor, out: Tensor, out_cat: Optional[Tensor] = None) -> None:
    """x fp32 [R,C,T] -> out act [R,T,C] (+ out_cat act [R,T,3C] = [hi | lo | hi], the split-precision operand)."""
    _chk(x, torch.float32), _chk(out, act_dtype(dt))
    R, Cc, T = x.shape
    assert tuple(out.shape) == (R, T, Cc)
    if out_cat is not None:
        _chk(out_cat, act_dtype(dt))
        assert tuple(out_cat.shape) == (R, T, 3 * Cc)
    call("dvae_pack_ncl_to_cl", dt, ptr(x), ptr(out), ptr(out_cat), R, Cc, T, stream())


def unpack_cl_to_ncl(dt: int, a: Tensor, b: Optional[Tensor], want_a: bool = True):
    """a [R,T,C] (fp32 or act), b [R,T,C] act or None -> (a^T fp32 [R,C,T] or None, (a+b)^T fp32 or None)."""
    assert a.is_cuda and a.is_contiguous()
    R, T, Cc = a.shape
    out_a = torch.empty((R, Cc, T), device=a.device, dtype=torch.float32) if want_a else None
    out_s = torch.empty((R, Cc, T), device=a.device, dtype=torch.float32) if b is not None else None
    call("dvae_unpack_cl_to_ncl", dt, ptr(a), int(a.dtype == torch.float32), ptr(b), ptr(out_a), ptr(out_s), R, Cc, T, stream())
    return out_a, out_s


def recon_out_bwd(dt: int, g_rec: Optional[Tensor], g_hat: Optional[Tensor], d_rec: Tensor, d_post: Tensor,
                  scale: float = 1.0) -> None:
    """g_* fp32 [R,C,T] (or None) -> d_rec = scale * (g_rec + g_hat)^T, d_post = scale * g_hat^T, both act [R,T,C]."""
    R, T, Cc = d_rec.shape
    for g in (g_rec, g_hat):
        if g is not None:
            _chk(g, torch.float32)
            assert tuple(g.shape) == (R, Cc, T)
    call("dvae_recon_out_bwd", dt, ptr(g_rec), ptr(g_hat), ptr(d_rec), ptr(d_post), R, Cc, T, float(scale), stream())


def chunk_mel(dt: int, mel: Tensor, mel_off: Tensor, t_len: Tensor, chunk_first: Tensor, chunk_utt: Tensor, n_chunks: int,
              C: int = 80, T: int = 64) -> Tensor:
    """chunking_mel for many utterances: flat fp32 `mel` holding [C, t_len[u]] matrices at mel_off[u] -> act [n_chunks, T, C]."""
    _chk(mel, torch.float32), _chk(mel_off, torch.int64), _chk(t_len, torch.int32), _chk(chunk_first, torch.int32)
    _chk(chunk_utt, torch.int32)
    x_cl = torch.empty((n_chunks, T, C), device=mel.device, dtype=act_dtype(dt))
    call("dvae_chunk_mel", dt, ptr(mel), ptr(mel_off), ptr(t_len), ptr(chunk_first), ptr(chunk_utt), ptr(x_cl), n_chunks, C, T, stream())
    return x_cl


def unchunk_mel(dt: int, a: Tensor, b: Optional[Tensor], out: Tensor, out_off: Tensor, chunk_first: Tensor, chunk_utt: Tensor,
                clamp: Optional[Tuple[float, float]] = None) -> None:
    """a fp32 [n_chunks, T, C] (+ b act) -> per-utterance [C, n_u * T] matrices inside flat fp32 `out`, optionally clamped."""
    _chk(a, torch.float32), _chk(out, torch.float32), _chk(out_off, torch.int64)
    n, T, C = a.shape
    if b is not None:
        _chk(b, act_dtype(dt))
    lo, hi = clamp if clamp is not None else (0.0, 0.0)
    call("dvae_unchunk_mel", dt, ptr(a), ptr(b), ptr(out_off), ptr(chunk_first), ptr(chunk_utt), ptr(out), n, C, T,
         int(clamp is not None), float(lo), float(hi), stream())


# ------------------------------------------------------------------------------- batch norm
def bn_train_fwd(dt: int, y: Tensor, gamma: Tensor, beta: Tensor, run_mean: Optional[Tensor], run_var: Optional[Tensor],
                 num_batches: Optional[Tensor], halves: int, act: int, eps: float, momentum: float):
    """y [rows, C] (rows = halves * rows_half).  Returns (out act [rows,C], stat fp32 [halves,4,C])."""
    ad = act_dtype(dt)
    _chk(y)
    C = y.shape[-1]
    rows = y.numel() // C
    assert rows % halves == 0
    out = torch.empty_like(y, dtype=ad)
    ws = torch.empty((halves * 2 * C,), device=y.device, dtype=torch.float64)
    stat = torch.empty((halves, 4, C), device=y.device, dtype=torch.float32)
    call("dvae_bn_train_fwd", dt, ptr(y), _y_is_f32(dt, y), ptr(out), ptr(gamma), ptr(beta), ptr(run_mean), ptr(run_var), ptr(num_batches),
         ptr(ws), ptr(stat), rows // halves, halves, C, act, eps, momentum, stream())
    return out, stat


def bn_finalize_apply(dt: int, y: Tensor, ws: Tensor, gamma: Tensor, beta: Tensor, run_mean: Optional[Tensor],
                      run_var: Optional[Tensor], num_batches: Optional[Tensor], halves: int, act: int, eps: float, momentum: float):
    """bn_train_fwd without its statistics pass: `ws` comes from conv5_fwd_bnstats.  Returns (out, stat)."""
    ad = act_dtype(dt)
    _chk(y), _chk(ws, torch.float64)
    C = y.shape[-1]
    rows = y.numel() // C
    assert rows % halves == 0 and ws.numel() >= halves * 2 * C
    out = torch.empty_like(y, dtype=ad)
    stat = torch.empty((halves, 4, C), device=y.device, dtype=torch.float32)
    call("dvae_bn_finalize_apply", dt, ptr(y), _y_is_f32(dt, y), ptr(out), ptr(gamma), ptr(beta), ptr(run_mean), ptr(run_var), ptr(num_batches),
         ptr(ws), ptr(stat), rows // halves, halves, C, act, eps, momentum, stream())
    return out, stat


def bn_eval_fwd(dt: int, y: Tensor, gamma: Tensor, beta: Tensor, run_mean: Tensor, run_var: Tensor, act: int, eps: float):
    ad = act_dtype(dt)
    _chk(y, ad)
    C = y.shape[-1]
    rows = y.numel() // C
    out = torch.empty_like(y)
    stat = torch.empty((1, 4, C), device=y.device, dtype=torch.float32)
    call("dvae_bn_eval_fwd", dt, ptr(y), ptr(out), ptr(gamma), ptr(beta), ptr(run_mean), ptr(run_var), ptr(stat), rows, C, act,
         eps, stream())
    return out


def bn_train_bwd(dt: int, dout: Tensor, y: Tensor, stat: Tensor, halves: int, act: int,
                 dgamma: Optional[Tensor] = None, dbeta: Optional[Tensor] = None, alpha: float = 1.0):
    """Returns (dy act [rows,C], dgamma fp32 [C], dbeta fp32 [C]); the two parameter gradients are scaled by alpha."""
    ad = act_dtype(dt)
    _chk(dout, ad), _chk(y), _chk(stat, torch.float32)
    C = y.shape[-1]
    rows = y.numel() // C
    dy = torch.empty_like(y, dtype=ad)
    ws = torch.empty((halves * 2 * C,), device=y.device, dtype=torch.float64)
    coef = torch.empty((halves * 2 * C,), device=y.device, dtype=torch.float32)
    dgamma = dgamma if dgamma is not None else torch.empty((C,), device=y.device, dtype=torch.float32)
    dbeta = dbeta if dbeta is not None else torch.empty((C,), device=y.device, dtype=torch.float32)
    call("dvae_bn_train_bwd", dt, ptr(dout), ptr(y), _y_is_f32(dt, y), ptr(stat), ptr(ws), ptr(coef), ptr(dy), ptr(dgamma), ptr(dbeta),
         rows // halves, halves, C, act, float(alpha), stream())
    return dy, dgamma, dbeta


def colsum(dt: int, x: Tensor, out: Tensor, alpha: float = 1.0) -> None:
    """out[C] (fp32, accumulating) += alpha * column sums of x viewed as [rows, C]."""
    _chk(x, act_dtype(dt)), _chk(out, torch.float32)
    C = out.numel()
    rows = x.numel() // C
    call("dvae_colsum", dt, ptr(x), ptr(out), rows, C, C, float(alpha), stream())


# ------------------------------------------------------------------------------- latent tail / loss
def latent_tail_fwd(dt: int, heads: Tensor, eps_c1, eps_c2, eps_s: Tensor, R: int, L: int, S: int, sample_content: bool):
    _chk(heads, torch.float32)
    assert tuple(heads.shape) == (2 * R, 2 * L)
    dev = heads.device
    z = torch.empty((2 * R, L), device=dev, dtype=act_dtype(dt))
    q = [torch.empty((R, L), device=dev, dtype=torch.float32) for _ in range(4)]
    zs = [torch.empty((R, S), device=dev, dtype=torch.float32) for _ in range(2)]
    call("dvae_latent_tail_fwd", dt, ptr(heads), ptr(eps_c1), ptr(eps_c2), ptr(eps_s), ptr(z), ptr(q[0]), ptr(q[1]), ptr(q[2]),
         ptr(q[3]), ptr(zs[0]), ptr(zs[1]), R, L, S, int(sample_content), stream())
    return z, q, zs


def latent_tail_bwd(dt: int, heads: Tensor, eps_c1, eps_c2, eps_s, dz: Tensor, dq, dzs, R: int, L: int, S: int,
                    sample_content: bool, gscale: float = 1.0) -> Tensor:
    _chk(dz, torch.float32)
    dheads = torch.empty((2 * R, 2 * L), device=heads.device, dtype=act_dtype(dt))
    call("dvae_latent_tail_bwd", dt, ptr(heads), ptr(eps_c1), ptr(eps_c2), ptr(eps_s), ptr(dz), ptr(dq[0]), ptr(dq[1]), ptr(dq[2]),
         ptr(dq[3]), ptr(dzs[0]), ptr(dzs[1]), ptr(dheads), R, L, S, int(sample_content), float(gscale), stream())
    return dheads


def loss_fwd(x1, x2, r1, r2, h1, h2, q1_mu, q1_lv, q2_mu, q2_lv, s_mu, s_lv, batch_size: float, mse_cof: float,
             kl_cof: float) -> Tensor:
    ts = [x1, x2, r1, r2, h1, h2, q1_mu, q1_lv, q2_mu, q2_lv, s_mu, s_lv]
    for t in ts:
        _chk(t, torch.float32)
    n = x1.numel()
    assert all(t.numel() == n for t in ts[:6])
    rows, L = q1_mu.shape
    S = s_mu.shape[1]
    ws = torch.empty((lib.workspace_bytes("loss") // 8,), device=x1.device, dtype=torch.float64)
    out = torch.empty((8,), device=x1.device, dtype=torch.float32)
    call("dvae_loss_fwd", *[ptr(t) for t in ts[:6]], n, *[ptr(t) for t in ts[6:10]], rows, L, ptr(s_mu), ptr(s_lv), S,
         float(batch_size), float(mse_cof), float(kl_cof), ptr(ws), ptr(out), stream())
    return out


def loss_bwd(x1, x2, r1, r2, h1, h2, q1_mu, q1_lv, q2_mu, q2_lv, s_mu, s_lv, batch_size, mse_cof, kl_cof, gout: Tensor):
    ts = [x1, x2, r1, r2, h1, h2, q1_mu, q1_lv, q2_mu, q2_lv, s_mu, s_lv]
    _chk(gout, torch.float32)
    n = x1.numel()
    rows, L = q1_mu.shape
    S = s_mu.shape[1]
    outs = [torch.empty_like(t) for t in ts[2:]]
    call("dvae_loss_bwd", *[ptr(t) for t in ts[:6]], n, *[ptr(t) for t in ts[6:10]], rows, L, ptr(s_mu), ptr(s_lv), S,
         float(batch_size), float(mse_cof), float(kl_cof), ptr(gout), *[ptr(t) for t in outs], stream())
    return outs


# ------------------------------------------------------------------------------- speaker groups
MODE_POG, MODE_MEAN, MODE_RAW = 0, 1, 2


def segment_ids_sorted(labels: Tensor):
    """labels int64 [B] on device with equal ids adjacent -> (gid int32 [B], num_groups int32 [1] on device)."""
    _chk(labels, torch.int64)
    B = labels.numel()
    gid = torch.empty((B,), device=labels.device, dtype=torch.int32)
    scratch = torch.empty((max(1, (B + 1023) // 1024),), device=labels.device, dtype=torch.int32)
    ng = torch.empty((1,), device=labels.device, dtype=torch.int32)
    call("dvae_segment_ids_sorted", ptr(labels), ptr(gid), ptr(scratch), ptr(ng), B, stream())
    return gid, ng


def group_accumulate(mode: int, a: Tensor, b: Tensor, gid: Tensor, G: int):
    _chk(a, torch.float32), _chk(b, torch.float32), _chk(gid, torch.int32)
    B, D = a.shape
    acc = torch.zeros((G, 2, D), device=a.device, dtype=torch.float32)
    cnt = torch.zeros((G,), device=a.device, dtype=torch.float32)
    call("dvae_group_accumulate", mode, ptr(a), ptr(b), ptr(gid), ptr(acc), ptr(cnt), B, D, stream())
    return acc, cnt


def group_finalize(mode: int, acc: Tensor, cnt: Tensor, gid: Tensor, B: int, D: int, want_b: bool = True):
    out_a = torch.empty((B, D), device=acc.device, dtype=torch.float32)
    out_b = torch.empty((B, D), device=acc.device, dtype=torch.float32) if want_b else None
    table = torch.empty_like(acc)
    call("dvae_group_finalize", mode, ptr(acc), ptr(cnt), ptr(gid), ptr(table), ptr(out_a), ptr(out_b), B, acc.shape[0], D, stream())
    return out_a, out_b


def group_pog_bwd(mu: Tensor, logvar: Tensor, gid: Tensor, acc_f: Tensor, acc_g: Tensor):
    B, D = mu.shape
    dmu, dlv = torch.empty_like(mu), torch.empty_like(mu)
    call("dvae_group_pog_bwd", ptr(mu), ptr(logvar), ptr(gid), ptr(acc_f), ptr(acc_g), ptr(dmu), ptr(dlv), B, D, stream())
    return dmu, dlv


def group_reparam(mu: Tensor, logvar: Tensor, gid: Tensor, eps_group: Tensor) -> Tensor:
    B, D = mu.shape
    z = torch.empty_like(mu)
    call("dvae_group_reparam", ptr(mu), ptr(logvar), ptr(gid), ptr(eps_group), ptr(z), B, D, stream())
    return z
