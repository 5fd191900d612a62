"""Tensor-level wrappers over the C ABI (include/pic_latent.h).

torch is used here for device memory, streams and dtype checks only; every computation is a
kernel of libpic_latent.so launched on the current CUDA stream without host synchronisation.
All tensor arguments must be CUDA float32 (int32 where stated); anything else raises -- there
is no CPU fallback on the product path.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import torch

from . import _lib
from ._lib import Q_ONES, Q_ZEROS, check, lib

Number = Union[int, float]

QUANTIZE_MODES = {"noise": 0, "dequantize": 1, "symbols": 2, "ste": 3}


# ----------------------------------------------------------------------------------- helpers
import functools


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _tensors_of(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            yield a
        elif isinstance(a, dict):
            for v in a.values():
                if isinstance(v, torch.Tensor):
                    yield v


def _on_tensor_device(fn):
    """Runs the wrapped op with the tensors' device current: the C ABI launches on the current device and
    `_stream()` returns the current device's stream, so a module moved to cuda:1 while cuda:0 is current must
    switch first.  All tensor arguments have to live on one device."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for t in _tensors_of(args, kwargs):
            if not t.is_cuda:
                continue           # reported by _require with the op's own message
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise RuntimeError(f"{fn.__name__}: all tensors must be on the same device, got {dev} and {t.device}")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _expand_like(t: Optional[torch.Tensor], ref: torch.Tensor, name: str) -> Optional[torch.Tensor]:
    """Broadcasts an operand to the shape of `ref` the way the reference's eager arithmetic does
    (entropy_models.py:146-149, 161-168 accept means of shape [B, C, 1, 1])."""
    if t is None or t.shape == ref.shape:
        return t
    try:
        return t.expand_as(ref).contiguous()
    except RuntimeError:
        raise ValueError(f"{name} of shape {tuple(t.shape)} does not broadcast to {tuple(ref.shape)}") from None


def _require(t: Optional[torch.Tensor], name: str, dtype=torch.float32) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the PIC latent path has no CPU implementation "
                           "(the reference CPU path lives in oracle/, test infrastructure only)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def pr_to_q01(pr: Number) -> float:
    """layers/channel_mask.py:133-140: pr>=10 -> ones, pr==0 -> zeros, else q = 1 - min(pr,10)*0.1
    (Python double arithmetic; torch casts it to f32 when it builds the quantile tensor)."""
    pr = float(pr)
    if pr >= 10:
        return Q_ONES
    if pr == 0:
        return Q_ZEROS
    pr = 10 if pr > 10 else pr
    pr = pr * 0.1
    return 1.0 - pr


def q01_tensor(prs: Sequence[Number], device) -> torch.Tensor:
    """Per-unit q01 control vector (one small H2D copy; build it once for a quality sweep)."""
    vals = [pr_to_q01(p) for p in prs]
    return torch.tensor(vals, dtype=torch.float32).to(device, non_blocking=True)


def bind_host_to_device_numa(device) -> Optional[int]:
    """Host-buffer callers (pic_slice_forward_host): restrict this process to the CPUs of the NUMA node the GPU
    hangs off, so that pinned buffers allocated afterwards (first touch) are local to its PCIe root port.  Returns
    the node, or None when the topology is not visible (containers) -- in which case nothing is changed."""
    import os

    try:
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = torch.cuda.get_device_properties(device).pci_domain_id
        dev_id = torch.cuda.get_device_properties(device).pci_device_id
        sysfs = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0"
        node = int(open(os.path.join(sysfs, "numa_node")).read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def _units_view(t: torch.Tensor, units: int) -> Tuple[int, int]:
    total = t.numel()
    if units <= 0 or total % units != 0:
        raise ValueError("tensor size is not a multiple of the number of units")
    return units, total // units


def workspace_bytes(n_per_unit: int, units: int) -> int:
    """Scratch bytes the select needs for `units` units of n_per_unit elements (pic_workspace_bytes)."""
    return int(lib().pic_workspace_bytes(n_per_unit, units))


def _workspace(n_per_unit: int, units: int, device, workspace: Optional[torch.Tensor] = None
               ) -> Tuple[Optional[torch.Tensor], int]:
    """Caller-provided scratch (uint8 CUDA tensor, reusable across calls on one stream) or a fresh allocation."""
    nbytes = workspace_bytes(n_per_unit, units)
    if workspace is not None:
        if not (workspace.is_cuda and workspace.dtype == torch.uint8 and workspace.is_contiguous()):
            raise TypeError("workspace must be a contiguous CUDA uint8 tensor")
        if workspace.device != torch.device(device):
            raise RuntimeError("workspace must live on the tensors' device")
        if workspace.numel() < nbytes:
            raise ValueError(f"workspace holds {workspace.numel()} bytes, {nbytes} are needed (ops.workspace_bytes)")
        return workspace, workspace.numel()
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return ws, nbytes


_FUSED_MAX = None


def fused_max_elems() -> int:
    global _FUSED_MAX
    if _FUSED_MAX is None:
        _FUSED_MAX = int(lib().pic_fused_max_elems())
    return _FUSED_MAX


def _q_args(q01, units, device):
    """Returns (scalar q01, per-unit tensor or None)."""
    if isinstance(q01, torch.Tensor):
        qt = _require(q01, "q01_per_unit")
        if qt.numel() != units:
            raise ValueError("q01_per_unit must hold one value per unit")
        return 0.5, qt
    return float(q01), None


# ----------------------------------------------------------------------------------- select / mask
@_on_tensor_device
def select_threshold(std: torch.Tensor, units: int, q01, want_ab: bool = False, workspace: Optional[torch.Tensor] = None):
    """Exact torch.quantile per unit. Returns thr [units] (and a, b if want_ab)."""
    std = _require(std, "std")
    units, n = _units_view(std, units)
    q, qt = _q_args(q01, units, std.device)
    thr = torch.empty(units, dtype=torch.float32, device=std.device)
    a = torch.empty_like(thr) if want_ab else None
    b = torch.empty_like(thr) if want_ab else None
    ws, ws_bytes = _workspace(n, units, std.device, workspace)
    check(lib().pic_select_threshold(_ptr(std), n, units, q, _ptr(qt), _ptr(thr), _ptr(a), _ptr(b),
                                     _ptr(ws), ws_bytes, _stream()), "pic_select_threshold")
    return (thr, a, b) if want_ab else thr


@_on_tensor_device
def channel_mask(std: torch.Tensor, units: int, q01, want_thr: bool = False, workspace: Optional[torch.Tensor] = None):
    """mask = (std >= quantile) as f32, same shape as std."""
    std = _require(std, "std")
    units, n = _units_view(std, units)
    q, qt = _q_args(q01, units, std.device)
    mask = torch.empty_like(std)
    thr = torch.empty(units, dtype=torch.float32, device=std.device) if want_thr else None
    ws, ws_bytes = _workspace(n, units, std.device, workspace)
    check(lib().pic_channel_mask(_ptr(std), n, units, q, _ptr(qt), _ptr(mask), _ptr(thr), _ptr(ws), ws_bytes,
                                 _stream()), "pic_channel_mask")
    return (mask, thr) if want_thr else mask


@_on_tensor_device
def attention_mask(std: torch.Tensor, units: int, q01, copies: int = 2, out: Optional[torch.Tensor] = None,
                   workspace: Optional[torch.Tensor] = None):
    """REM attention mask: channel_mask written `copies` times along dim 1, shape [units, copies * C, ...]
    (torch.cat([m] * copies, 1) of the reference) from one select + one pass over std."""
    std = _require(std, "std")
    units, n = _units_view(std, units)
    q, qt = _q_args(q01, units, std.device)
    if out is None:
        shape = (units, copies * n) if std.dim() < 2 else (std.shape[0], copies * std.shape[1]) + tuple(std.shape[2:])
        out = torch.empty(shape, dtype=torch.float32, device=std.device)
    elif out.numel() != units * copies * n or not out.is_contiguous():
        raise ValueError("out must be a contiguous tensor of units * copies * n elements")
    ws, ws_bytes = _workspace(n, units, std.device, workspace)
    check(lib().pic_attention_mask(_ptr(std), n, units, q, _ptr(qt), copies, _ptr(out), None, _ptr(ws), ws_bytes,
                                   _stream()), "pic_attention_mask")
    return out


@_on_tensor_device
def select_threshold_multi(std: torch.Tensor, units: int, prs: Sequence[Number]) -> torch.Tensor:
    """Thresholds of every unit at every quality of `prs` in one launch. Returns thr [units, len(prs)]."""
    std = _require(std, "std")
    units, n = _units_view(std, units)
    levels = len(prs)
    q = q01_tensor(list(prs) * units, std.device)          # [units, levels], level-minor
    thr = torch.empty((units, levels), dtype=torch.float32, device=std.device)
    check(lib().pic_select_threshold_multi(_ptr(std), n, units, _ptr(q), levels, _ptr(thr), _stream()),
          "pic_select_threshold_multi")
    return thr


@_on_tensor_device
def level_map(std: torch.Tensor, thr: torch.Tensor, units: int) -> torch.Tensor:
    """level[e] = first l with std[e] >= thr[u, l] (or L): delta mask of level l == (level == l)."""
    std, thr = _require(std, "std"), _require(thr, "thr")
    units, n = _units_view(std, units)
    levels = thr.numel() // units
    out = torch.empty(std.shape, dtype=torch.int32, device=std.device)
    check(lib().pic_level_map(_ptr(std), _ptr(thr), n, units, levels, _ptr(out), _stream()), "pic_level_map")
    return out


@_on_tensor_device
def rank_order(std: torch.Tensor, units: int) -> torch.Tensor:
    """Explicit ranking of every unit: unit-local element indexes sorted by (std descending, index ascending),
    int32 [units, n].  The threshold mask's support is order[:, :kept] with kept = #{std >= thr}."""
    std = _require(std, "std")
    units, n = _units_view(std, units)
    order = torch.empty((units, n), dtype=torch.int32, device=std.device)
    nbytes = int(lib().pic_rank_order_workspace_bytes(n, units))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=std.device)
    check(lib().pic_rank_order(_ptr(std), n, units, _ptr(order), _ptr(ws), nbytes, _stream()), "pic_rank_order")
    return order


@_on_tensor_device
def mask_from_threshold(std: torch.Tensor, thr: torch.Tensor, units: int) -> torch.Tensor:
    std, thr = _require(std, "std"), _require(thr, "thr")
    units, n = _units_view(std, units)
    mask = torch.empty_like(std)
    check(lib().pic_mask_from_threshold(_ptr(std), _ptr(thr), n, units, _ptr(mask), _stream()),
          "pic_mask_from_threshold")
    return mask


# ----------------------------------------------------------------------------------- fused slice
@_on_tensor_device
def slice_forward(y_top, y_base, mu, std, units: int, q01, scale_table: Optional[torch.Tensor] = None,
                  noise=None, thr_in=None, scale_bound: float = 0.11, lik_bound: float = 1e-9,
                  want=("mask", "y_hat", "lik"), out: Optional[dict] = None,
                  workspace: Optional[torch.Tensor] = None) -> dict:
    """One progressive slice for `units` units (see pic_slice_forward in include/pic_latent.h).
    want: subset of {mask, y_hat, lik, idx, symbols, thr, rate}.  `out` may carry preallocated
    output tensors (keys as in `want`) to avoid allocations in a steady-state loop."""
    y_top, y_base, mu, std = (_require(t, nm) for t, nm in
                              ((y_top, "y_top"), (y_base, "y_base"), (mu, "mu"), (std, "std")))
    noise, thr_in = _require(noise, "noise"), _require(thr_in, "thr_in")
    units, n = _units_view(std, units)
    for t, nm in ((y_top, "y_top"), (y_base, "y_base"), (mu, "mu"), (noise, "noise")):
        if t is not None and t.numel() != std.numel():
            raise ValueError(f"{nm} must have the same number of elements as std")
    q, qt = _q_args(q01, units, std.device)
    table = _require(scale_table, "scale_table") if "idx" in want else None
    if "idx" in want and table is None:
        raise ValueError("scale_table is required for idx")
    res = dict(out) if out else {}
    dev = std.device
    for k in ("mask", "y_hat", "lik"):
        if k in want and k not in res:
            res[k] = torch.empty_like(std)
    for k in ("idx", "symbols"):
        if k in want and k not in res:
            res[k] = torch.empty(std.shape, dtype=torch.int32, device=dev)
    if "thr" in want and "thr" not in res:
        res["thr"] = torch.empty(units, dtype=torch.float32, device=dev)
    if "rate" in want and "rate" not in res:
        res["rate"] = torch.empty(units, dtype=torch.float64, device=dev)
    ws, ws_bytes = _workspace(n, units, dev, workspace)
    check(lib().pic_slice_forward(
        _ptr(y_top), _ptr(y_base), _ptr(mu), _ptr(std), q, _ptr(qt), _ptr(thr_in), _ptr(noise),
        _ptr(table), 0 if table is None else table.numel(), scale_bound, lik_bound, n, units,
        _ptr(res.get("mask")), _ptr(res.get("y_hat")), _ptr(res.get("lik")), _ptr(res.get("idx")),
        _ptr(res.get("symbols")), _ptr(res.get("thr")), _ptr(res.get("rate")), _ptr(ws), ws_bytes, _stream()),
        "pic_slice_forward")
    return res


@_on_tensor_device
def slice_forward_multi(y_top, y_base, mu, std, units: int, prs: Sequence[Number], scale_table=None,
                        scale_bound: float = 0.11, lik_bound: float = 1e-9, want=("mask", "y_hat", "lik"),
                        out: Optional[dict] = None, q01_levels: Optional[torch.Tensor] = None) -> dict:
    """Quality sweep of one set of latents (pic_slice_forward_multi): every output is [units, len(prs), n]; the
    `len(prs)` qualities of a unit share its inputs.  `q01_levels` may carry a prepared [units, levels] q01 tensor."""
    y_top, y_base, mu, std = (_require(t, nm) for t, nm in
                              ((y_top, "y_top"), (y_base, "y_base"), (mu, "mu"), (std, "std")))
    units, n = _units_view(std, units)
    levels = len(prs)
    q = _require(q01_levels, "q01_levels") if q01_levels is not None else q01_tensor(list(prs) * units, std.device)
    if q.numel() != units * levels:
        raise ValueError("q01_levels must hold units * levels values")
    table = _require(scale_table, "scale_table") if "idx" in want else None
    if "idx" in want and table is None:
        raise ValueError("scale_table is required for idx")
    res = dict(out) if out else {}
    dev, shape = std.device, (units, levels, n)
    for k in ("mask", "y_hat", "lik"):
        if k in want and k not in res:
            res[k] = torch.empty(shape, dtype=torch.float32, device=dev)
    for k in ("idx", "symbols"):
        if k in want and k not in res:
            res[k] = torch.empty(shape, dtype=torch.int32, device=dev)
    if "thr" not in res:
        res["thr"] = torch.empty((units, levels), dtype=torch.float32, device=dev)
    if "rate" in want and "rate" not in res:
        res["rate"] = torch.empty((units, levels), dtype=torch.float64, device=dev)
    check(lib().pic_slice_forward_multi(
        _ptr(y_top), _ptr(y_base), _ptr(mu), _ptr(std), _ptr(q), levels, _ptr(table),
        0 if table is None else table.numel(), scale_bound, lik_bound, n, units,
        _ptr(res.get("mask")), _ptr(res.get("y_hat")), _ptr(res.get("lik")), _ptr(res.get("idx")),
        _ptr(res.get("symbols")), _ptr(res["thr"]), _ptr(res.get("rate")), _stream()), "pic_slice_forward_multi")
    return res


@_on_tensor_device
def slice_backward(g_lik, g_yhat, y_top, y_base, mu, std, mask, noise=None, scale_bound: float = 0.11,
                   lik_bound: float = 1e-9, need_base: bool = True):
    """Returns (g_ytop, g_ybase or None, g_mu, g_std)."""
    ts = [_require(t, nm) for t, nm in ((g_lik, "g_lik"), (g_yhat, "g_yhat"), (y_top, "y_top"), (y_base, "y_base"),
                                        (mu, "mu"), (std, "std"), (mask, "mask"), (noise, "noise"))]
    g_lik, g_yhat, y_top, y_base, mu, std, mask, noise = ts
    g_ytop = torch.empty_like(std)
    g_ybase = torch.empty_like(std) if (y_base is not None and need_base) else None
    g_mu = torch.empty_like(std)
    g_std = torch.empty_like(std)
    check(lib().pic_slice_backward(_ptr(g_lik), _ptr(g_yhat), _ptr(y_top), _ptr(y_base), _ptr(mu), _ptr(std),
                                   _ptr(mask), _ptr(noise), scale_bound, lik_bound, std.numel(), _ptr(g_ytop),
                                   _ptr(g_ybase), _ptr(g_mu), _ptr(g_std), _stream()), "pic_slice_backward")
    return g_ytop, g_ybase, g_mu, g_std


# ----------------------------------------------------------------------------------- un-fused ops
@_on_tensor_device
def gaussian_forward(inputs, scales, means=None, noise=None, likelihood_only: bool = False,
                     scale_bound: float = 0.11, lik_bound: float = 1e-9, want_outputs: bool = True):
    inputs, scales, means, noise = (_require(t, nm) for t, nm in
                                    ((inputs, "inputs"), (scales, "scales"), (means, "means"), (noise, "noise")))
    scales, means, noise = (_expand_like(t, inputs, nm) for t, nm in ((scales, "scales"), (means, "means"), (noise, "noise")))
    outputs = torch.empty_like(inputs) if (want_outputs and not likelihood_only) else None
    lik = torch.empty_like(inputs)
    check(lib().pic_gaussian_forward(_ptr(inputs), _ptr(scales), _ptr(means), _ptr(noise), int(likelihood_only),
                                     inputs.numel(), scale_bound, lik_bound, _ptr(outputs), _ptr(lik), _stream()),
          "pic_gaussian_forward")
    return outputs, lik


@_on_tensor_device
def gaussian_backward(g_out, g_lik, inputs, scales, means=None, noise=None, likelihood_only: bool = False,
                      scale_bound: float = 0.11, lik_bound: float = 1e-9):
    ts = [_require(t, nm) for t, nm in ((g_out, "g_out"), (g_lik, "g_lik"), (inputs, "inputs"), (scales, "scales"),
                                        (means, "means"), (noise, "noise"))]
    g_out, g_lik, inputs, scales, means, noise = ts
    g_in = torch.empty_like(inputs)
    g_sc = torch.empty_like(inputs)
    g_mu = torch.empty_like(inputs) if means is not None else None
    check(lib().pic_gaussian_backward(_ptr(g_out), _ptr(g_lik), _ptr(inputs), _ptr(scales), _ptr(means), _ptr(noise),
                                      int(likelihood_only), inputs.numel(), scale_bound, lik_bound, _ptr(g_in),
                                      _ptr(g_sc), _ptr(g_mu), _stream()), "pic_gaussian_backward")
    return g_in, g_sc, g_mu


@_on_tensor_device
def build_indexes(scales, scale_table, scale_bound: float = 0.11) -> torch.Tensor:
    scales, scale_table = _require(scales, "scales"), _require(scale_table, "scale_table")
    idx = torch.empty(scales.shape, dtype=torch.int32, device=scales.device)
    check(lib().pic_build_indexes(_ptr(scales), scales.numel(), _ptr(scale_table), scale_table.numel(), scale_bound,
                                  _ptr(idx), _stream()), "pic_build_indexes")
    return idx


@_on_tensor_device
def quantize(inputs, mode: str, means=None, noise=None, mask=None) -> torch.Tensor:
    if mode not in QUANTIZE_MODES:
        raise ValueError(f'Invalid quantization mode: "{mode}"')  # entropy_models.py:130-131
    inputs, means, noise, mask = (_require(t, nm) for t, nm in
                                  ((inputs, "inputs"), (means, "means"), (noise, "noise"), (mask, "mask")))
    means, noise, mask = (_expand_like(t, inputs, nm) for t, nm in ((means, "means"), (noise, "noise"), (mask, "mask")))
    out_f = torch.empty_like(inputs) if mode != "symbols" else None
    out_i = torch.empty(inputs.shape, dtype=torch.int32, device=inputs.device) if mode == "symbols" else None
    check(lib().pic_quantize(_ptr(inputs), _ptr(means), _ptr(noise), _ptr(mask), inputs.numel(),
                             QUANTIZE_MODES[mode], _ptr(out_f), _ptr(out_i), _stream()), "pic_quantize")
    return out_i if mode == "symbols" else out_f


@_on_tensor_device
def dequantize(symbols, means=None) -> torch.Tensor:
    symbols = _require(symbols, "symbols", torch.int32)
    means = _expand_like(_require(means, "means"), symbols, "means")
    out = torch.empty(symbols.shape, dtype=torch.float32, device=symbols.device)
    check(lib().pic_dequantize(_ptr(symbols), _ptr(means), symbols.numel(), _ptr(out), _stream()), "pic_dequantize")
    return out


@_on_tensor_device
def log_sum(x, units: int) -> torch.Tensor:
    """Per-unit sum of ln(x) in f64 (numerator of training/loss.py:45-60)."""
    x = _require(x, "x")
    units, n = _units_view(x, units)
    out = torch.empty(units, dtype=torch.float64, device=x.device)
    check(lib().pic_log_sum(_ptr(x), n, units, _ptr(out), _stream()), "pic_log_sum")
    return out


# ----------------------------------------------------------------------------------- EntropyBottleneck (z)
@_on_tensor_device
def bottleneck_forward(z, medians, params, filters, noise=None, lik_bound: float = 1e-9):
    """EntropyBottleneck.forward on z [B, C, ...]: returns (outputs, likelihood) (pic_bottleneck_forward)."""
    z, medians, params, noise = (_require(t, nm) for t, nm in ((z, "z"), (medians, "medians"), (params, "params"), (noise, "noise")))
    B, Cc = z.shape[0], z.shape[1]
    S = z.numel() // (B * Cc)
    outputs, lik = torch.empty_like(z), torch.empty_like(z)
    f = tuple(int(v) for v in filters)
    check(lib().pic_bottleneck_forward(_ptr(z), _ptr(noise), _ptr(medians), _ptr(params), f[0], f[1], f[2], f[3], B, Cc, S,
                                       lik_bound, _ptr(outputs), _ptr(lik), _stream()), "pic_bottleneck_forward")
    return outputs, lik


@_on_tensor_device
def bottleneck_backward(z, medians, params, filters, g_lik, g_out=None, noise=None, lik_bound: float = 1e-9):
    """Returns (g_z, g_params [C, per_channel], g_medians [C])."""
    ts = [_require(t, nm) for t, nm in ((z, "z"), (medians, "medians"), (params, "params"), (g_lik, "g_lik"), (g_out, "g_out"),
                                        (noise, "noise"))]
    z, medians, params, g_lik, g_out, noise = ts
    B, Cc = z.shape[0], z.shape[1]
    S = z.numel() // (B * Cc)
    g_z = torch.empty_like(z)
    g_params = torch.empty_like(params)
    g_med = torch.empty_like(medians)
    f = tuple(int(v) for v in filters)
    check(lib().pic_bottleneck_backward(_ptr(z), _ptr(noise), _ptr(medians), _ptr(params), f[0], f[1], f[2], f[3], B, Cc, S,
                                        lik_bound, _ptr(g_lik), _ptr(g_out), _ptr(g_z), _ptr(g_params), _ptr(g_med), _stream()),
          "pic_bottleneck_backward")
    return g_z, g_params, g_med
