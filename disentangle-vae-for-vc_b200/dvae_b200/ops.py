"""Tensor-level wrappers over the C-ABI (one Python function per exported op).

These allocate outputs with torch (device memory / stream plumbing only) and pass raw pointers to
libdvae_b200.so.  No arithmetic happens in Python or torch here.
Activation storage dtype `dt` is lib.BF16 (torch.bfloat16, tcgen05 kind::f16) or lib.TF32
(torch.float32 storage, tcgen05 kind::tf32).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import lib
from .lib import call, ptr, stream

Tensor = torch.Tensor


def act_dtype(dt: int) -> torch.dtype:
    return torch.bfloat16 if dt == lib.BF16 else torch.float32


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


def linear_wgrad(dt: int, dy: Tensor, x: Tensor, dw: Tensor) -> None:
    """dw[N,K] (fp32, pre-zeroed or accumulating) += dy[M,N]^T @ x[M,K]."""
    ad = act_dtype(dt)
    _chk(dy, ad), _chk(x, ad), _chk(dw, torch.float32)
    M, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == M and tuple(dw.shape) == (N, K)
    call("dvae_linear_wgrad", dt, ptr(dy), N, ptr(x), K, ptr(dw), K, M, N, K, stream())


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


def conv5_wgrad(dt: int, dy: Tensor, x: Tensor, dwk: Tensor) -> None:
    """dwk[Cout,5,Cin] (fp32, accumulating) += sum_{r,t} dy[r,t,co] x[r,t+k-2,ci]."""
    ad = act_dtype(dt)
    _chk(dy, ad), _chk(x, ad), _chk(dwk, torch.float32)
    R, T, Cout = dy.shape
    Cin = x.shape[2]
    assert tuple(dwk.shape) == (Cout, 5, Cin)
    call("dvae_conv5_wgrad", dt, ptr(dy), ptr(x), ptr(dwk), R, T, Cin, Cout, stream())


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
    dc = torch.empty((D, rows, H), device=dh_all.device, dtype=torch.float32)
    call("dvae_lstm_bwd", dt, ptr(dh_all), ptr(gates), ptr(c_all), ptr(whh_n), ptr(da), ptr(dc), rows, T, H, D, stream())
    return da


def lstm_wgrad_hh(dt: int, da_all: Tensor, h_all: Tensor, dwhh: Tensor, H: int, D: int) -> None:
    ad = act_dtype(dt)
    _chk(da_all, ad), _chk(h_all, ad), _chk(dwhh, torch.float32)
    rows, T, _ = da_all.shape
    assert tuple(dwhh.shape) == (D, 4 * H, H)
    call("dvae_lstm_wgrad_hh", dt, ptr(da_all), ptr(h_all), ptr(dwhh), rows, T, H, D, stream())
