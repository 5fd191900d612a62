"""Drop-in for the reference's ``layers/channel_mask.py`` (ChannelMask, ste_round).

Same class name, constructor, method names, argument meaning, return dtype (f32 masks) and
error behaviour; the per-image ``torch.quantile`` sort + compare loop is replaced by the exact
radix-select kernels of libpic_latent.so (one launch for the whole batch, no host sync).

Reference lines: ChannelMask.forward 89-156 (live branch 132-151, "two-levels" 152-153),
ProgMask 18-49, apply_noise 81-86, ste_round 5-6.  ``delta_mask`` (52-78) and the ``cust_map``
quantile branch (96-121) are dead/broken in the reference (chained tensor comparison raises;
undefined ``bs, ch, w, h``) and are not reproduced beyond their pr>=10 / pr==0 shortcuts.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _SteRound(torch.autograd.Function):
    """torch.round(x) - x.detach() + x: value of round(), identity gradient."""

    @staticmethod
    def forward(ctx, x):
        return ops.quantize(x, "ste")

    @staticmethod
    def backward(ctx, g):
        return g


def ste_round(x: torch.Tensor) -> torch.Tensor:
    return _SteRound.apply(x)


class ChannelMask(nn.Module):
    def __init__(self, mask_policy):
        super().__init__()
        self.mask_policy = mask_policy

    # ------------------------------------------------------------------ ProgMask (18-49)
    def ProgMask(self, scale, pr):
        """`scale` is a list of per-slice blocks [1, ch, w, h]; returns the stacked masks
        [len(scale)*bs, ch, w, h] (f32).  All blocks go through ONE launch."""
        if len(scale) == 0:
            raise RuntimeError("stack expects a non-empty TensorList")
        blocks = [b if b.is_contiguous() else b.contiguous() for b in scale]
        shape0 = blocks[0].shape
        same = all(b.shape == shape0 for b in blocks)
        if not same:
            # ragged blocks cannot be stacked by the reference either (torch.stack raises)
            raise RuntimeError("stack expects each tensor to be equal size")
        bs, ch, w, h = shape0
        stacked = torch.cat([b.reshape(bs, ch * w * h) for b in blocks], dim=0)
        q01 = ops.pr_to_q01(pr)
        mask = ops.channel_mask(stacked, stacked.shape[0], q01)
        return mask.reshape(len(blocks) * bs, ch, w, h)

    def ProgLevels(self, scale, q_list):
        """All progressive levels at once (test/functions_encode.py:176-190: two ProgMask calls per level
        on the same scale list).  `scale`: list of per-slice blocks [1, ch, w, h]; `q_list`: increasing
        qualities.  Returns (level, thr): level [len(scale), ch, w, h] int32 with the first level that keeps
        each element (len(q_list) if none) -- ProgMask(q_l) - ProgMask(q_{l-1}) == (level == l), with
        q_{-1} = 0 -- and the thresholds thr [len(scale), len(q_list)]."""
        if len(scale) == 0:
            raise RuntimeError("stack expects a non-empty TensorList")
        blocks = [b if b.is_contiguous() else b.contiguous() for b in scale]
        if not all(b.shape == blocks[0].shape for b in blocks):
            raise RuntimeError("stack expects each tensor to be equal size")
        bs, ch, w, h = blocks[0].shape
        stacked = torch.cat([b.reshape(bs, ch * w * h) for b in blocks], dim=0)
        q_list = list(q_list)
        if stacked.shape[1] <= ops.fused_max_elems():
            thr = ops.select_threshold_multi(stacked, stacked.shape[0], q_list)
        else:
            # larger blocks (images beyond ~1024x1024): one large-unit select per level, as ProgMask does
            thr = torch.stack([ops.select_threshold(stacked, stacked.shape[0], ops.pr_to_q01(q)) for q in q_list], dim=1)
            thr = thr.contiguous()
        level = ops.level_map(stacked, thr, stacked.shape[0])
        return level.reshape(len(blocks) * bs, ch, w, h), thr

    def delta_mask(self, scale, pr_bar, pr):
        raise NotImplementedError("delta_mask is dead code in the reference (chained tensor comparison raises)")

    # ------------------------------------------------------------------ apply_noise (81-86)
    def apply_noise(self, mask, training):
        if training:
            mask = ste_round(mask)
        else:
            mask = ops.quantize(mask, "dequantize")  # torch.round
        return mask

    # ------------------------------------------------------------------ forward (89-156)
    def forward(self, scale, pr=0, mask_pol="point-based-std", ravel=False, cust_map=None):
        if cust_map is not None:
            if pr >= 10:
                return torch.ones_like(cust_map).to(scale.device)
            elif pr == 0:
                return torch.zeros_like(cust_map).to(scale.device)
            raise NotImplementedError("cust_map quantile branch is broken in the reference (undefined bs/ch/w/h)")

        if mask_pol is None:
            mask_pol = self.mask_policy

        shapes = scale.shape
        if ravel is False:
            bs, ch, w, h = shapes
        else:
            bs, d = shapes

        if mask_pol == "point-based-std":
            if pr >= 10:
                return torch.ones_like(scale)
            elif pr == 0:
                return torch.zeros_like(scale)
            assert scale is not None
            q01 = ops.pr_to_q01(pr)
            mask = ops.channel_mask(scale, bs, q01)
            return mask.reshape(shapes)
        elif mask_pol == "two-levels":
            return torch.zeros_like(scale) if pr == 0 else torch.ones_like(scale)
        else:
            raise NotImplementedError()

    # ------------------------------------------------------------------ REM attention mask
    def attention_mask(self, scale, pr, training=False, mu_std=False, mask_pol="point-based-std"):
        """models/rem_pic.py:181-195 (apply_latent_enhancement): the `star_mask` at quality `pr` on the
        pre-REM scale, rounded, duplicated on the channel axis when the REM refines (mu, std) jointly.
        (The reference also computes a `bar_mask` at `quality_bar` and discards it; not reproduced.)"""
        if mu_std and mask_pol == "point-based-std" and scale.dim() >= 2:
            # one select + one pass writes both halves; apply_noise is the identity on a {0,1} mask without grad
            return ops.attention_mask(scale.contiguous(), scale.shape[0], ops.pr_to_q01(pr), copies=2)
        m = self.apply_noise(self.forward(scale, pr=pr, mask_pol=mask_pol), training)
        return torch.cat([m, m], dim=1) if mu_std else m

    # ------------------------------------------------------------------ extensions (not in the reference)
    def rank_order(self, scale):
        """Explicit variance-aware ranking per image: int32 [bs, ch*w*h] element indexes (NCHW, image-local) from the
        most to the least uncertain, ties broken by ascending index.  `forward(scale, pr)` keeps exactly the first
        `kept = #{scale >= quantile}` entries -- more than ceil(pr/10 * n) only when elements tie with the threshold
        (the reference's `>=` keeps all of them)."""
        return ops.rank_order(scale.contiguous(), scale.shape[0])

    def forward_multi(self, scale, prs):
        """One quality level per image/unit: `prs` is a sequence of len(scale) qualities or a
        prepared per-unit q01 tensor (ops.q01_tensor)."""
        bs = scale.shape[0]
        q = prs if isinstance(prs, torch.Tensor) else ops.q01_tensor(prs, scale.device)
        return ops.channel_mask(scale, bs, q).reshape(scale.shape)
