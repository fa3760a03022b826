"""ctypes binding of libdvae_b200.so (the C-ABI declared in include/dvae_b200.h).

Every entry point takes raw device pointers, sizes and a cudaStream_t and returns an int status
(0 = ok).  `call()` raises RuntimeError with `dvae_last_error()` on a non-zero status.  The library
must exist: there is deliberately no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdvae_b200.so")

BF16, TF32, F16, F32 = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
        "(dvae_b200 has no CPU / PyTorch fallback)")

_lib = C.CDLL(LIB_PATH)
_lib.dvae_last_error.restype = C.c_char_p

_p, _i, _l, _f, _d = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_double

# name -> argtypes (all return int)
_SIGNATURES = {
    "dvae_linear_fwd": [_i, _p, _l, _p, _p, _p, _p, _l, _i, _i, _i, _i, _i, _p],
    "dvae_linear_fwd_split": [_i, _p, _l, _p, _p, _p, _p, _l, _p, _i, _i, _i, _i, _i, _p],
    "dvae_linear_dgrad": [_i, _p, _l, _p, _p, _p, _p, _l, _i, _i, _i, _i, _p],
    "dvae_linear_wgrad": [_i, _p, _l, _p, _l, _p, _l, _i, _i, _i, _f, _p],
    "dvae_conv5_fwd": [_i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "dvae_conv5_fwd_bnstats": [_i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _i, _i, _p],
    "dvae_conv5_dgrad": [_i, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "dvae_conv5_wgrad": [_i, _p, _p, _p, _i, _i, _i, _i, _f, _p],
    "dvae_lstm_fwd": [_i, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "dvae_lstm_bwd": [_i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "dvae_lstm_bwd_workspace": [_i, _i, _i, _i, _p, _p],
    "dvae_debug_timing": [_p, _i],
    "dvae_debug_seq_stamps": [_p],
    "dvae_debug_res_stamps": [_p],
    "dvae_set_background": [_i],
    "dvae_lstm_wgrad_hh": [_i, _p, _p, _p, _i, _i, _i, _i, _f, _p],
    # weight preparation / layout
    "dvae_prep_cast": [_i, _p, _p, _l, _f, _p],
    "dvae_add_inplace": [_i, _p, _p, _l, _p],
    "dvae_copy_f32": [_p, _p, _l, _p],
    "dvae_add_f32_act": [_i, _p, _p, _p, _l, _p],
    "dvae_prep_conv_weight": [_i, _p, _p, _p, _i, _i, _p],
    "dvae_prep_cast_split": [_i, _p, _p, _l, _l, _i, _p],
    "dvae_prep_all": [_i, _p, _p, _p, _i, _i, _p],
    "dvae_conv_wgrad_unpack": [_p, _p, _i, _i, _p],
    "dvae_prep_lstm_weight": [_i, _p, _p, _i, _i, _i, _p],
    "dvae_prep_lstm_bias": [_p, _p, _p, _i, _i, _p],
    "dvae_pack_ncl_to_cl": [_i, _p, _p, _p, _i, _i, _i, _p],
    "dvae_unpack_cl_to_ncl": [_i, _p, _i, _p, _p, _p, _i, _i, _i, _p],
    "dvae_recon_out_bwd": [_i, _p, _p, _p, _p, _i, _i, _i, _f, _p],
    "dvae_chunk_mel": [_i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "dvae_unchunk_mel": [_i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p],
    # batch norm / reductions
    "dvae_bn_train_fwd": [_i, _p, _i] + [_p] * 8 + [_i, _i, _i, _i, _f, _f, _p],
    "dvae_bn_finalize_apply": [_i, _p, _i] + [_p] * 8 + [_i, _i, _i, _i, _f, _f, _p],
    "dvae_bn_eval_fwd": [_i] + [_p] * 7 + [_l, _i, _i, _f, _p],
    "dvae_bn_train_bwd": [_i, _p, _p, _i] + [_p] * 6 + [_i, _i, _i, _i, _f, _p],
    "dvae_colsum": [_i, _p, _p, _l, _i, _l, _f, _p],
    # latent tail / loss / speaker groups
    "dvae_latent_tail_fwd": [_i] + [_p] * 11 + [_i, _i, _i, _i, _p],
    "dvae_latent_tail_bwd": [_i] + [_p] * 12 + [_i, _i, _i, _i, _f, _p],
    "dvae_loss_fwd": [_p] * 6 + [_l] + [_p] * 4 + [_i, _i, _p, _p, _i, _f, _f, _f, _p, _p, _p],
    "dvae_loss_bwd": [_p] * 6 + [_l] + [_p] * 4 + [_i, _i, _p, _p, _i, _f, _f, _f] + [_p] * 11 + [_p],
    "dvae_segment_ids_sorted": [_p, _p, _p, _p, _l, _p],
    "dvae_group_accumulate": [_i, _p, _p, _p, _p, _p, _l, _i, _p],
    "dvae_group_finalize": [_i, _p, _p, _p, _p, _p, _p, _l, _l, _i, _p],
    "dvae_group_pog_bwd": [_p] * 7 + [_l, _i, _p],
    "dvae_group_reparam": [_p] * 5 + [_l, _i, _p],
    # optimizer
    "dvae_adam_step": [_p] * 7 + [_i, _i, _d, _d, _d, _d, _l, _p, _p],
}
_OPTIONAL = {}


def _bind(name, argtypes):
    fn = getattr(_lib, name)
    fn.argtypes = argtypes
    fn.restype = C.c_int
    return fn


def register(name, argtypes):
    _SIGNATURES[name] = argtypes
    _FUNCS[name] = _bind(name, argtypes)


_FUNCS = {n: _bind(n, a) for n, a in _SIGNATURES.items()}
for _n in ("dvae_version", "dvae_sm_arch", "dvae_lstm_gate_tile", "dvae_lstm_launches", "dvae_lstm_launches_for", "dvae_set_lstm_resident"):
    getattr(_lib, _n).restype = C.c_int
_lib.dvae_lstm_launches.argtypes = [C.c_int, C.c_int, C.c_int]
_lib.dvae_lstm_launches_for.argtypes = [C.c_int] * 6
_lib.dvae_set_lstm_resident.argtypes = [C.c_int]
_lib.dvae_workspace_bytes.restype = C.c_long
_lib.dvae_workspace_bytes.argtypes = [C.c_char_p, C.c_long, C.c_long, C.c_long]


def workspace_bytes(op: str, n0: int = 0, n1: int = 0, n2: int = 0) -> int:
    """Bytes of scratch the library wants for `op` (it never allocates itself); raises for an unknown name."""
    n = _lib.dvae_workspace_bytes(op.encode(), n0, n1, n2)
    if n < 0:
        raise KeyError(f"dvae_workspace_bytes: unknown op {op!r}")
    return n


def ptr(t):
    """Device pointer of a tensor (or None)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# kernels launched per C-ABI call (bench.py reports the total as `gpu_launches`)
LAUNCHES = 0
_LAUNCHES_PER_CALL = {"dvae_bn_finalize_apply": 2, "dvae_bn_train_fwd": 3, "dvae_bn_train_bwd": 3, "dvae_bn_eval_fwd": 2, "dvae_segment_ids_sorted": 3,
                      "dvae_group_finalize": 2}
_LSTM_SHAPE_ARGS = {"dvae_lstm_fwd": 5, "dvae_lstm_bwd": 9}   # index of `rows` (then T, H, D): the library says how many launches


def call(name, *args):
    global LAUNCHES
    if name in _LSTM_SHAPE_ARGS:
        i = _LSTM_SHAPE_ARGS[name]
        LAUNCHES += _lib.dvae_lstm_launches_for(args[0], args[i], args[i + 1], args[i + 2], args[i + 3],
                                                1 if name == "dvae_lstm_bwd" else 0)
    else:
        LAUNCHES += _LAUNCHES_PER_CALL.get(name, 1)
    rc = _FUNCS[name](*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (status {rc}): {_lib.dvae_last_error().decode()}")


def version() -> int:
    return _lib.dvae_version()


def lstm_bwd_workspace(dt: int, rows: int, H: int, D: int):
    """(floats of split-K fix-up workspace, number of int32 tickets) for dvae_lstm_bwd at this shape."""
    ws, nt = C.c_long(0), C.c_int(0)
    rc = _FUNCS["dvae_lstm_bwd_workspace"](dt, rows, H, D, C.cast(C.byref(ws), C.c_void_p), C.cast(C.byref(nt), C.c_void_p))
    if rc != 0:
        raise RuntimeError("dvae_lstm_bwd_workspace failed")
    return ws.value, nt.value


def set_lstm_resident(on: int) -> int:
    """Switch the time-resident H = 512 / 1024 recurrence kernels on (1) / off (0); negative: query.  Returns the previous setting."""
    return _lib.dvae_set_lstm_resident(int(on))


def lstm_gate_tile(hidden: int) -> int:
    return _lib.dvae_lstm_gate_tile(hidden)


def exported_symbols():
    return sorted(_SIGNATURES.keys()) + ["dvae_last_error", "dvae_version", "dvae_sm_arch", "dvae_lstm_gate_tile",
                                             "dvae_lstm_launches", "dvae_lstm_launches_for", "dvae_set_lstm_resident", "dvae_workspace_bytes"]
