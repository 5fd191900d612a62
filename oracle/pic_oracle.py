"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front-end of the C oracle (oracle/pic_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.  All arrays are
C-contiguous numpy float32 / int32; shapes are [units, n_per_unit] (any leading
shape is flattened by the caller).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpic_oracle.so")
_lib = None

SCALE_BOUND = 0.11
LIK_BOUND = 1e-9


def build(force: bool = False) -> str:
    """Compiles the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "pic_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _f(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _d(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _c32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def scale_table(smin: float = 0.11, smax: float = 256.0, levels: int = 64) -> np.ndarray:
    """models/pic.py:12-17 computed with torch f32 ops (exp(linspace(log min, log max)));
    the f32 values are data, committed in tests/golden/scale_table.npy.  This helper only
    exists for CPU-side timing where the exact table bits do not matter."""
    return np.exp(np.linspace(np.log(smin), np.log(smax), levels)).astype(np.float32)


def quantile(x: np.ndarray, q: float):
    x = _c32(x).ravel()
    thr, a, b = C.c_float(), C.c_float(), C.c_float()
    rc = lib().pic_oracle_quantile(_f(x), C.c_int64(x.size), C.c_float(np.float32(q)),
                                   C.byref(thr), C.byref(a), C.byref(b), None)
    if rc == -2:
        raise RuntimeError("quantile() input tensor is too large")
    if rc != 0:
        raise ValueError(f"pic_oracle_quantile failed rc={rc}")
    return np.float32(thr.value), np.float32(a.value), np.float32(b.value)


def _pr_array(pr, units):
    arr = np.atleast_1d(np.asarray(pr, dtype=np.float64))
    per_unit = 1 if arr.size > 1 else 0
    if per_unit and arr.size != units:
        raise ValueError("pr must be a scalar or one value per unit")
    return np.ascontiguousarray(arr), per_unit


def channel_mask(scale: np.ndarray, pr):
    """ChannelMask.forward(scale, pr) for scale [units, n]; returns (mask, thr)."""
    scale = _c32(scale)
    units, n = scale.shape
    prs, per_unit = _pr_array(pr, units)
    mask = np.empty_like(scale)
    thr = np.empty(units, dtype=np.float32)
    rc = lib().pic_oracle_channel_mask(_f(scale), C.c_int64(n), C.c_int64(units), _d(prs),
                                       C.c_int(per_unit), _f(mask), _f(thr))
    if rc != 0:
        raise RuntimeError(f"pic_oracle_channel_mask rc={rc}")
    return mask, thr


def build_indexes(scales: np.ndarray, table: np.ndarray, scale_bound: float = SCALE_BOUND):
    scales = _c32(scales)
    table = _c32(table)
    idx = np.empty(scales.shape, dtype=np.int32)
    rc = lib().pic_oracle_build_indexes(_f(scales), C.c_int64(scales.size), _f(table),
                                        C.c_int(table.size), C.c_float(scale_bound), _i(idx))
    if rc != 0:
        raise RuntimeError(f"pic_oracle_build_indexes rc={rc}")
    return idx


def gaussian_forward(inputs, scales, means=None, noise=None, scale_bound=SCALE_BOUND,
                     lik_bound=LIK_BOUND, f64: bool = False):
    """GaussianConditional.forward; noise=None => eval (round). Returns (outputs, lik)."""
    inputs, scales, means, noise = _c32(inputs), _c32(scales), _c32(means), _c32(noise)
    if f64:
        lik = np.empty(inputs.shape, dtype=np.float64)
        lib().pic_oracle_gaussian_forward_f64(_f(inputs), _f(scales), _f(means), _f(noise),
                                              C.c_int64(inputs.size), C.c_double(scale_bound),
                                              C.c_double(lik_bound), _d(lik))
        return None, lik
    out = np.empty_like(inputs)
    lik = np.empty_like(inputs)
    lib().pic_oracle_gaussian_forward(_f(inputs), _f(scales), _f(means), _f(noise),
                                      C.c_int64(inputs.size), C.c_float(scale_bound),
                                      C.c_float(lik_bound), _f(out), _f(lik))
    return out, lik


def quantize(inputs, mode: str, means=None, noise=None, mask=None):
    modes = {"noise": 0, "dequantize": 1, "symbols": 2}
    if mode not in modes:
        raise ValueError(f'Invalid quantization mode: "{mode}"')
    inputs, means, noise, mask = _c32(inputs), _c32(means), _c32(noise), _c32(mask)
    out_f = np.empty_like(inputs) if mode != "symbols" else None
    out_i = np.empty(inputs.shape, dtype=np.int32) if mode == "symbols" else None
    rc = lib().pic_oracle_quantize(_f(inputs), _f(means), _f(noise), _f(mask),
                                   C.c_int64(inputs.size), C.c_int(modes[mode]), _f(out_f), _i(out_i))
    if rc != 0:
        raise ValueError(f"pic_oracle_quantize rc={rc}")
    return out_i if mode == "symbols" else out_f


def slice_forward(y_top, y_base, mu, scale, pr, table, noise=None, scale_bound=SCALE_BOUND,
                  lik_bound=LIK_BOUND, want=("mask", "thr", "y_hat", "lik", "idx", "symbols", "rate")):
    """One progressive slice for all units ([units, n] arrays). Returns a dict."""
    y_top, y_base, mu, scale, noise = _c32(y_top), _c32(y_base), _c32(mu), _c32(scale), _c32(noise)
    table = _c32(table)
    units, n = scale.shape
    prs, per_unit = _pr_array(pr, units)
    out = {}
    for k in ("mask", "y_hat", "lik"):
        out[k] = np.empty_like(scale) if k in want else None
    out["thr"] = np.empty(units, dtype=np.float32) if "thr" in want else None
    for k in ("idx", "symbols"):
        out[k] = np.empty(scale.shape, dtype=np.int32) if k in want else None
    out["rate"] = np.empty(units, dtype=np.float64) if "rate" in want else None
    rc = lib().pic_oracle_slice_forward(
        _f(y_top), _f(y_base), _f(mu), _f(scale), _d(prs), C.c_int(per_unit), _f(noise), _f(table),
        C.c_int(table.size), C.c_float(scale_bound), C.c_float(lik_bound), C.c_int64(n),
        C.c_int64(units), _f(out["mask"]), _f(out["thr"]), _f(out["y_hat"]), _f(out["lik"]),
        _i(out["idx"]), _i(out["symbols"]), _d(out["rate"]))
    if rc == -2:
        raise RuntimeError("quantile() input tensor is too large")
    if rc != 0:
        raise RuntimeError(f"pic_oracle_slice_forward rc={rc}")
    return {k: v for k, v in out.items() if v is not None}


def slice_backward(g_lik, g_yhat, y_top, y_base, mu, scale, mask, noise=None,
                   scale_bound=SCALE_BOUND, lik_bound=LIK_BOUND):
    g_lik, g_yhat = _c32(g_lik), _c32(g_yhat)
    y_top, y_base, mu, scale, mask, noise = map(_c32, (y_top, y_base, mu, scale, mask, noise))
    g_ytop = np.empty_like(scale)
    g_ybase = np.empty_like(scale) if y_base is not None else None
    g_mu = np.empty_like(scale)
    g_scale = np.empty_like(scale)
    lib().pic_oracle_slice_backward(_f(g_lik), _f(g_yhat), _f(y_top), _f(y_base), _f(mu), _f(scale),
                                    _f(mask), _f(noise), C.c_float(scale_bound), C.c_float(lik_bound),
                                    C.c_int64(scale.size), _f(g_ytop), _f(g_ybase), _f(g_mu),
                                    _f(g_scale))
    return dict(g_ytop=g_ytop, g_ybase=g_ybase, g_mu=g_mu, g_scale=g_scale)


def gaussian_backward(g_out, g_lik, inputs, scales, means=None, noise=None,
                      scale_bound=SCALE_BOUND, lik_bound=LIK_BOUND):
    g_out, g_lik, inputs, scales, means, noise = map(_c32, (g_out, g_lik, inputs, scales, means, noise))
    g_in = np.empty_like(inputs)
    g_sc = np.empty_like(inputs)
    g_mu = np.empty_like(inputs) if means is not None else None
    lib().pic_oracle_gaussian_backward(_f(g_out), _f(g_lik), _f(inputs), _f(scales), _f(means),
                                       _f(noise), C.c_int64(inputs.size), C.c_float(scale_bound),
                                       C.c_float(lik_bound), _f(g_in), _f(g_sc), _f(g_mu))
    return dict(g_inputs=g_in, g_scales=g_sc, g_means=g_mu)


def num_threads() -> int:
    return int(lib().pic_oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().pic_oracle_set_num_threads(C.c_int(int(n)))


# ---- elementwise neighbours of the path; numpy f32, same operation order as the reference ----
def lrp_merge(y_hat, lrp, base=None):
    """models/pic.py:635-641: lrp = 0.5 * tanh(lrp); y_hat += lrp; merge(y_hat, base) = y_hat + base."""
    t = np.float32(0.5) * np.tanh(_c32(lrp))
    out = _c32(y_hat) + t.astype(np.float32)
    return out if base is None else (out + _c32(base)).astype(np.float32)


def lrp_merge_backward(g_out, lrp):
    th = np.tanh(_c32(lrp).astype(np.float64))
    return (_c32(g_out).astype(np.float64) * 0.5 * (1.0 - th * th)).astype(np.float32)


def rem_merge(identity, ret, att_mask):
    """layers/rem.py:137-140: identity + ret * att_mask."""
    return (_c32(identity) + (_c32(ret) * _c32(att_mask)).astype(np.float32)).astype(np.float32)
