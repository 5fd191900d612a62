"""Fused progressive-slice operator: the per-slice body of the model loops as ONE launch.

Replaces, for a batch of units, the reference lines (models/pic.py)
    583-584  y_slice = y_slice - y_slices[current_index]
    621-622  block_mask = masking(scale, pr=quality); apply_noise(block_mask, False)
    625-629  y_slice_m = (y_slice - mu) * block_mask
             _, lik = gaussian_conditional(y_slice_m, scale * block_mask, training=training)
             y_hat_slice = ste_round(y_slice - mu) * block_mask + mu
    813, 819 index = build_indexes(scale * block_mask); symbols = quantize(y_slice_m, "symbols")
(and their REM twins models/rem_pic.py:319-320, 382-391, 588-597) with autograd flowing to
y_top, y_base, mu and scale exactly as through the reference ops.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch

from . import ops


class _SliceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_top, y_base, mu, scale, q01, noise, table, scale_bound, lik_bound, want_idx, want_sym,
                want_rate, units):
        want = ["mask", "y_hat", "lik", "thr"]
        if want_idx:
            want.append("idx")
        if want_sym:
            want.append("symbols")
        if want_rate:
            want.append("rate")
        res = ops.slice_forward(y_top, y_base, mu, scale, units, q01, table, noise=noise,
                                scale_bound=scale_bound, lik_bound=lik_bound, want=tuple(want))
        ctx.save_for_backward(y_top, y_base, mu, scale, res["mask"], noise)
        ctx.cfg = (scale_bound, lik_bound)
        idx = res.get("idx", y_top.new_empty(0, dtype=torch.int32))
        sym = res.get("symbols", y_top.new_empty(0, dtype=torch.int32))
        rate = res.get("rate", y_top.new_empty(0, dtype=torch.float64))
        ctx.mark_non_differentiable(res["mask"], idx, sym, res["thr"], rate)
        return res["y_hat"], res["lik"], res["mask"], idx, sym, res["thr"], rate

    @staticmethod
    def backward(ctx, g_yhat, g_lik, *_unused):
        y_top, y_base, mu, scale, mask, noise = ctx.saved_tensors
        scale_bound, lik_bound = ctx.cfg
        g_yhat = None if g_yhat is None else g_yhat.contiguous()
        g_lik = None if g_lik is None else g_lik.contiguous()
        need_base = y_base is not None and ctx.needs_input_grad[1]
        g_ytop, g_ybase, g_mu, g_std = ops.slice_backward(g_lik, g_yhat, y_top, y_base, mu, scale, mask, noise,
                                                          scale_bound, lik_bound, need_base=need_base)
        return (g_ytop, g_ybase, g_mu, g_std) + (None,) * 9


def progressive_slice_forward(y_top: torch.Tensor, y_base: Optional[torch.Tensor], mu: torch.Tensor,
                              scale: torch.Tensor, pr: Union[float, Sequence[float], torch.Tensor],
                              gaussian_conditional=None, training: bool = False, noise: Optional[torch.Tensor] = None,
                              want_indexes: bool = False, want_symbols: bool = False, want_rate: bool = False,
                              scale_bound: float = 0.11, lik_bound: float = 1e-9, scale_table=None):
    """Returns a dict with y_hat, likelihood, mask, thr (+ indexes, symbols, rate when asked).

    y_top / y_base / mu / scale: [B, C, h, w] CUDA f32 (a unit is one image's [C,h,w] block).
    pr: the reference's quality on the 0..10 scale -- a scalar, one value per image, or a
        prepared per-unit q01 tensor (ops.q01_tensor) for quality sweeps.
    training: adds U(-1/2, 1/2) noise drawn with the same torch call as the reference
        (or the supplied `noise`).
    """
    units = scale.shape[0]
    if gaussian_conditional is not None:
        scale_bound, lik_bound = gaussian_conditional._bounds()
        scale_table = gaussian_conditional.scale_table
    if isinstance(pr, torch.Tensor):
        q01 = pr
    elif isinstance(pr, (list, tuple)):
        q01 = ops.q01_tensor(pr, scale.device)
    else:
        q01 = ops.pr_to_q01(pr)
    if training and noise is None:
        noise = torch.empty_like(scale).uniform_(-0.5, 0.5)
    if not training:
        noise = None
    y_hat, lik, mask, idx, sym, thr, rate = _SliceFn.apply(
        y_top.contiguous(), None if y_base is None else y_base.contiguous(), mu.contiguous(), scale.contiguous(),
        q01, noise, scale_table, scale_bound, lik_bound, want_indexes, want_symbols, want_rate, units)
    out = {"y_hat": y_hat, "likelihood": lik, "mask": mask, "thr": thr}
    if want_indexes:
        out["indexes"] = idx
    if want_symbols:
        out["symbols"] = sym
    if want_rate:
        out["rate"] = rate
    return out


class _LogSum(torch.autograd.Function):
    """sum(ln x) over the whole tensor as one kernel (f64 accumulation); d/dx = g / x, the gradient autograd
    gives torch.log(x).sum() in the reference's loss (training/loss.py:45-60)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.log_sum(x, 1).sum()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return (g.to(x.dtype) / x)


def rate_bpp(likelihoods: torch.Tensor, num_pixels: int) -> torch.Tensor:
    """training/loss.py:45-60: sum(log(lik)) / (-ln2 * num_pixels), as an f64 scalar tensor; differentiable
    with respect to the likelihoods (the rate term of the training loss)."""
    import math

    total = _LogSum.apply(likelihoods.contiguous())
    return total / (-math.log(2) * num_pixels)


# ---------------------------------------------------------------------------------------------------------------
# elementwise neighbours of the path: one fused pass each, with autograd
# ---------------------------------------------------------------------------------------------------------------
def _same(*ts):
    shape = ts[0].shape
    for t in ts:
        if t is not None and t.shape != shape:
            raise ValueError("tensors must have the same shape")
    return [None if t is None else ops._require(t, "tensor") for t in ts]


class _LrpMerge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_hat, lrp, base):
        y_hat, lrp, base = _same(y_hat, lrp, base)
        out = torch.empty_like(y_hat)
        ops.check(ops.lib().pic_lrp_merge(ops._ptr(y_hat), ops._ptr(lrp), ops._ptr(base), ops._ptr(out), out.numel(),
                                          ops._stream()), "pic_lrp_merge")
        ctx.save_for_backward(lrp)
        ctx.has_base = base is not None
        return out

    @staticmethod
    def backward(ctx, g):
        (lrp,) = ctx.saved_tensors
        g = g.contiguous()
        g_lrp = None
        if ctx.needs_input_grad[1]:
            g_lrp = torch.empty_like(lrp)
            ops.check(ops.lib().pic_lrp_merge_backward(ops._ptr(g), ops._ptr(lrp), ops._ptr(g_lrp), g.numel(),
                                                       ops._stream()), "pic_lrp_merge_backward")
        return (g if ctx.needs_input_grad[0] else None, g_lrp, g if (ctx.has_base and ctx.needs_input_grad[2]) else None)


def lrp_merge(y_hat_slice: torch.Tensor, lrp: torch.Tensor, base: Optional[torch.Tensor] = None) -> torch.Tensor:
    """models/pic.py:635-641 in one pass:  `lrp = 0.5 * torch.tanh(lrp); y_hat_slice += lrp;
    y_hat_slice = self.merge(y_hat_slice, y_hat_slices[current_index])`  (merge = sum; base=None skips it)."""
    return _LrpMerge.apply(y_hat_slice, lrp, base)


class _RemMerge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, identity, ret, att_mask):
        identity, ret, att_mask = _same(identity, ret, att_mask)
        out = torch.empty_like(identity)
        ops.check(ops.lib().pic_rem_merge(ops._ptr(identity), ops._ptr(ret), ops._ptr(att_mask), ops._ptr(out),
                                          out.numel(), ops._stream()), "pic_rem_merge")
        ctx.save_for_backward(att_mask)
        return out

    @staticmethod
    def backward(ctx, g):
        (att_mask,) = ctx.saved_tensors
        g = g.contiguous()
        g_ret = None
        if ctx.needs_input_grad[1]:
            g_ret = torch.empty_like(g)
            ops.check(ops.lib().pic_rem_merge_backward(ops._ptr(g), ops._ptr(att_mask), ops._ptr(g_ret), g.numel(),
                                                       ops._stream()), "pic_rem_merge_backward")
        return (g if ctx.needs_input_grad[0] else None), g_ret, None


def rem_merge(identity: torch.Tensor, ret: torch.Tensor, att_mask: torch.Tensor) -> torch.Tensor:
    """layers/rem.py:137-140 in one pass:  `ret = ret * att_mask; res = identity + ret`."""
    return _RemMerge.apply(identity, ret, att_mask)
