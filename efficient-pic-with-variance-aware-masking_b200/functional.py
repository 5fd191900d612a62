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


def rate_bpp(likelihoods: torch.Tensor, num_pixels: int) -> torch.Tensor:
    """training/loss.py:45-60: sum(log(lik)) / (-ln2 * num_pixels), as an f64 scalar tensor."""
    import math

    total = ops.log_sum(likelihoods.contiguous(), 1).sum()
    return total / (-math.log(2) * num_pixels)
