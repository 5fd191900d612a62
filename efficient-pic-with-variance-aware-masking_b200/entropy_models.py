"""Drop-in for the hot-path part of the reference's ``entropy_models/entropy_models.py``.

``EntropyModel`` / ``GaussianConditional`` keep the reference's constructor signature, buffer
and sub-module names (so state_dicts load: ``scale_table``, ``scale_bound``, ``_offset``,
``_quantized_cdf``, ``_cdf_length``, ``lower_bound_scale.bound``,
``likelihood_lower_bound.bound``), method names, argument meaning, return dtypes and
exceptions.  forward / _likelihood / build_indexes / quantize / dequantize run as kernels of
libpic_latent.so with a fused backward.

Reference lines: EntropyModel 70-168 (quantize 127-153, dequantize 161-168),
GaussianConditional 528-672 (_likelihood 620-635, forward 637-652, build_indexes 654-659).
Codec side: rANS compress/decompress (206-294) and the CDF-table build in
update() (591-618) use compressai's C++ extension when it is importable, else the native coder of
this package (codec.py over include/pic_codec.h: same bit-stream by construction, byte parity with
compressai unpinned because it is not installed here).  EntropyBottleneck (297-525) is not on the path.
"""
from __future__ import annotations

import warnings
from typing import Any, List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import codec, ops


class _LowerBoundFn(torch.autograd.Function):
    """compressai.ops.LowerBound (v1.2.4): max(x, bound); gradient passes iff x >= bound or g < 0."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)) * g, None


class LowerBound(nn.Module):
    """Same buffer layout as compressai.ops.LowerBound (a 1-element f32 ``bound``)."""

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))
        self._value = float(torch.tensor(float(bound), dtype=torch.float32))  # f32-rounded

    def value(self) -> float:
        return self._value

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._value = float(self.bound.detach().cpu().reshape(-1)[0])

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound)


class _CompressaiCoder:
    """The reference's coder handle (entropy_models.py:18-52) when compressai is importable: one object exposing
    `encode_with_indexes` / `decode_with_indexes` of the named backend."""

    _BACKENDS = {
        "ans": lambda: __import__("compressai.ans", fromlist=["RansEncoder"]),
        "rangecoder": lambda: __import__("range_coder"),
    }

    def __init__(self, method):
        if type(method) is not str:
            raise ValueError(f'Invalid method type "{type(method)}"')
        import compressai

        known = compressai.available_entropy_coders()
        if method not in known:
            raise ValueError(f'Unknown entropy coder "{method}" (available: {", ".join(known)})')
        mod = self._BACKENDS[method]()
        enc, dec = (mod.RansEncoder(), mod.RansDecoder()) if method == "ans" else (mod.RangeEncoder(), mod.RangeDecoder())
        self.name = method
        self.encode_with_indexes = enc.encode_with_indexes
        self.decode_with_indexes = dec.decode_with_indexes


def _make_entropy_coder(method):
    """compressai's coder when it is installed (default method = compressai.get_entropy_coder()), else the
    package's native rANS coder (codec.RansCoder), which speaks the same bit-stream."""
    try:
        import compressai
    except ImportError:
        if method not in (None, "ans"):
            raise ValueError(f'Unknown entropy coder "{method}" (available: ans)')
        return codec.RansCoder()
    return _CompressaiCoder(compressai.get_entropy_coder() if method is None else method)


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder: Optional[str] = None,
                 entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder = _make_entropy_coder(entropy_coder)
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    # ---- pickling / deepcopy (reference 103-110): the coder handle is stored by name --------------------
    def __getstate__(self):
        attributes = self.__dict__.copy()
        attributes["entropy_coder"] = self.entropy_coder.name
        attributes.pop("_tables_cache", None)     # host copy of the CDF tables: rebuilt on demand
        attributes["_tables_gen_seen"] = None
        return attributes

    def __setstate__(self, state):
        self.__dict__ = state
        self.entropy_coder = _make_entropy_coder(self.__dict__.pop("entropy_coder"))

    def _invalidate_tables(self) -> None:
        """Called whenever the CDF buffers are replaced (update(), load_state_dict, .to()): the native coder's
        host copy is keyed on a generation counter, not on the buffer's address (the allocator reuses blocks)."""
        self._tables_gen = getattr(self, "_tables_gen", 0) + 1

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._invalidate_tables()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._invalidate_tables()
        return out

    @property
    def offset(self):
        return self._offset

    @property
    def quantized_cdf(self):
        return self._quantized_cdf

    @property
    def cdf_length(self):
        return self._cdf_length

    def forward(self, *args: Any) -> Any:
        raise NotImplementedError()

    # ---- quantize (127-153) -------------------------------------------------------------
    def quantize(self, inputs, mode, means=None, mask=None) -> Tensor:
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            half = float(0.5)
            noise = torch.empty_like(inputs).uniform_(-half, half)  # same RNG call as the reference
            return _QuantizeNoise.apply(inputs, noise, mask)
        return ops.quantize(inputs, mode, means)

    def _quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        warnings.warn("_quantize is deprecated. Use quantize instead.")
        return self.quantize(inputs, mode, means)

    # ---- dequantize (161-168) -----------------------------------------------------------
    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None) -> Tensor:
        if inputs.dtype == torch.int32 and (means is None or means.dtype == torch.float32):
            return ops.dequantize(inputs, means)
        if means is not None:  # other dtypes: plain tensor arithmetic as in the reference
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.float()
        return outputs

    @classmethod
    def _dequantize(cls, inputs: Tensor, means: Optional[Tensor] = None) -> Tensor:
        warnings.warn("_dequantize. Use dequantize instead.")
        return cls.dequantize(inputs, means)

    # ---- rANS coding (206-294) ----------------------------------------------------------
    _TABLE_CHECKS = (   # (buffer, wanted rank, message when empty, message when the rank is wrong) -- reference 185-204
        ("_quantized_cdf", 2, "Uninitialized CDFs. Run update() first", "Invalid CDF size {}"),
        ("_cdf_length", 1, "Uninitialized CDF lengths. Run update() first", "Invalid offsets size {}"),
        ("_offset", 1, "Uninitialized offsets. Run update() first", "Invalid offsets size {}"),
    )

    def _check_tables(self) -> None:
        for name, rank, empty_msg, rank_msg in self._TABLE_CHECKS:
            buf = getattr(self, name)
            if buf.numel() == 0:
                raise ValueError(empty_msg)
            if buf.dim() != rank:
                raise ValueError(rank_msg.format(buf.size()))

    def _tables(self) -> "codec.CdfTables":
        """Host copy of the CDF tables for the native coder, rebuilt when update() replaces the buffers."""
        gen = (getattr(self, "_tables_gen", 0), self._quantized_cdf._version)
        if getattr(self, "_tables_gen_seen", None) != gen or getattr(self, "_tables_cache", None) is None:
            self._check_tables()
            self._tables_cache = codec.CdfTables(self._quantized_cdf, self._cdf_length, self._offset)
            self._tables_gen_seen = gen
        return self._tables_cache

    def _stream_lists(self):
        """The tables as the Python lists compressai's coder takes."""
        return (self._quantized_cdf.tolist(), self._cdf_length.reshape(-1).int().tolist(),
                self._offset.reshape(-1).int().tolist())

    def compress(self, inputs, indexes, means=None, flag=1, already_quantize=False):
        """entropy_models.py:206-241: one stream per entry of dim 0."""
        symbols = inputs if already_quantize else self.quantize(inputs, "symbols", means)
        if inputs.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if symbols.shape != indexes.shape:
            raise ValueError("`inputs` and `indexes` should have the same size.")
        if isinstance(self.entropy_coder, codec.RansCoder):
            # one D2H copy of the int32 symbols / indexes, every stream on its own host thread
            return codec.encode_streams(symbols, indexes, self._tables())
        tables = self._stream_lists()
        flat_s, flat_i = symbols.reshape(symbols.size(0), -1).int(), indexes.reshape(indexes.size(0), -1).int()
        return [self.entropy_coder.encode_with_indexes(row_s.tolist(), row_i.tolist(), *tables)
                for row_s, row_i in zip(flat_s, flat_i)]

    def decompress(self, strings, indexes, means=None, flag=1):
        """entropy_models.py:243-294."""
        if type(strings) not in (tuple, list):
            raise ValueError("Invalid `strings` parameter type.")
        if len(strings) != indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if indexes.dim() < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        self._check_tables()
        if means is not None:
            if tuple(means.shape[:2]) != tuple(indexes.shape[:2]):
                raise ValueError("Invalid means or indexes parameters")
            if means.shape != indexes.shape and any(d != 1 for d in means.shape[2:]):
                raise ValueError("Invalid means parameters")   # broadcastable means only
        if isinstance(self.entropy_coder, codec.RansCoder):
            symbols = codec.decode_streams(strings, indexes, self._tables()).to(indexes.device)
            return self.dequantize(symbols, means)
        tables = self._stream_lists()
        symbols = self._quantized_cdf.new_empty(indexes.size())
        for k, stream in enumerate(strings):
            vals = self.entropy_coder.decode_with_indexes(stream, indexes[k].reshape(-1).int().tolist(), *tables)
            symbols[k] = torch.tensor(vals, device=symbols.device, dtype=symbols.dtype).reshape(symbols[k].shape)
        return self.dequantize(symbols, means)


class _QuantizeNoise(torch.autograd.Function):
    """inputs + noise (* mask): identity gradient to inputs (entropy_models.py:132-138)."""

    @staticmethod
    def forward(ctx, inputs, noise, mask):
        return ops.quantize(inputs, "noise", None, noise, mask)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class _GaussianForward(torch.autograd.Function):
    """GaussianConditional.forward as one kernel with a fused backward."""

    @staticmethod
    def forward(ctx, inputs, scales, means, noise, scale_bound, lik_bound, likelihood_only):
        outputs, lik = ops.gaussian_forward(inputs, scales, means, noise, likelihood_only, scale_bound, lik_bound)
        ctx.save_for_backward(inputs, scales, means, noise)
        ctx.cfg = (scale_bound, lik_bound, likelihood_only)
        if likelihood_only:
            outputs = inputs.new_empty(0)
            ctx.mark_non_differentiable(outputs)
        return outputs, lik

    @staticmethod
    def backward(ctx, g_out, g_lik):
        inputs, scales, means, noise = ctx.saved_tensors
        scale_bound, lik_bound, likelihood_only = ctx.cfg
        if likelihood_only:
            g_out = None
        g_in, g_sc, g_mu = ops.gaussian_backward(
            None if g_out is None else g_out.contiguous(), None if g_lik is None else g_lik.contiguous(),
            inputs, scales, means, noise, likelihood_only, scale_bound, lik_bound)
        return g_in, g_sc, g_mu, None, None, None, None


class GaussianConditional(EntropyModel):
    """Constructor contract of the reference (entropy_models.py:531-567): `scale_table` is None or a non-empty,
    sorted, positive list / tuple; `scale_bound` > 0 (defaults to 0.11; None takes the table's first entry)."""

    def __init__(self, scale_table: Optional[Union[List, Tuple]], *args: Any, scale_bound: float = 0.11,
                 tail_mass: float = 1e-9, **kwargs: Any):
        super().__init__(*args, **kwargs)
        if scale_table is not None:
            if type(scale_table) not in (list, tuple):
                raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
            if len(scale_table) == 0:
                raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
            ascending = all(a <= b for a, b in zip(scale_table, scale_table[1:]))
            if not ascending or min(scale_table) <= 0:
                raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = float(scale_table[0])
        if scale_bound is None or scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor([float(s) for s in scale_table])

    def _standardized_cumulative(self, inputs: Tensor) -> Tensor:
        """Phi(x) = erfc(-x / sqrt(2)) / 2 with the reference's constants (f32 erfc of c * x, c = float(-(2 ** -0.5)))."""
        return float(0.5) * torch.erfc(float(-(2 ** -0.5)) * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        from scipy.stats import norm

        return norm.ppf(quantile)

    def update_scale_table(self, scale_table):
        self.update(scale_table)
        return True

    def update(self, scale_table):
        """entropy_models.py:591-618.  Row i of the table describes round(N(0, s_i)) on the support
        [-c_i, c_i], c_i = ceil(s_i * |Phi^-1(tail_mass / 2)|), plus one escape slot holding the two tails: its pmf is
        computed with the reference's own torch ops in f32 (on the host: 64 rows) and normalised to 16-bit CDF rows by
        compressai's pmf_to_quantized_cdf when importable, else pic_pmf_to_quantized_cdf (include/pic_codec.h)."""
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        try:
            from compressai._CXX import pmf_to_quantized_cdf as normalise
        except ImportError:
            normalise = codec.pmf_to_quantized_cdf
        sigma = self.scale_table.cpu().float()
        half_width = torch.ceil(sigma * -self._standardized_quantile(self.tail_mass / 2)).int()      # c_i
        support = 2 * half_width + 1
        widest = int(support.max())
        # |k - c_i| for k = 0 .. widest-1 (int32, then f32), standardised by sigma_i: same operations as the reference
        dist = (torch.arange(widest).int() - half_width[:, None]).abs().float()
        upper = self._standardized_cumulative((0.5 - dist) / sigma[:, None])
        lower = self._standardized_cumulative((-0.5 - dist) / sigma[:, None])
        pmf, tails = upper - lower, 2 * lower[:, :1]
        table = torch.zeros((sigma.numel(), widest + 2), dtype=torch.int32)
        for row, (p, width, tail) in enumerate(zip(pmf, support.tolist(), tails)):
            cdf_row = normalise(torch.cat((p[:width], tail)).tolist(), self.entropy_coder_precision)
            table[row, : len(cdf_row)] = torch.tensor(cdf_row, dtype=torch.int32)
        self._quantized_cdf = table.to(device)
        self._offset = (-half_width).to(device)
        self._cdf_length = (support + 2).to(device)
        self._invalidate_tables()

    # ---- hot path ---------------------------------------------------------------------------
    def _bounds(self) -> Tuple[float, float]:
        lik_bound = self.likelihood_lower_bound.value() if self.use_likelihood_bound else 0.0
        return self.lower_bound_scale.value(), lik_bound

    @staticmethod
    def _same_shape(inputs, scales, means):
        if scales.shape != inputs.shape:
            scales = scales.expand_as(inputs)
        if means is not None and means.shape != inputs.shape:
            means = means.expand_as(inputs)
        return scales.contiguous(), None if means is None else means.contiguous()

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        scale_bound, _ = self._bounds()
        scales, means = self._same_shape(inputs, scales, means)
        _, lik = _GaussianForward.apply(inputs.contiguous(), scales, means, None, scale_bound, 0.0, True)
        return lik

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None, mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        if training is None:
            training = self.training
        scale_bound, lik_bound = self._bounds()
        scales, means = self._same_shape(inputs, scales, means)
        noise = None
        if training:
            half = float(0.5)
            noise = torch.empty_like(inputs).uniform_(-half, half)  # same RNG call as quantize("noise")
            if mask is not None:
                noise = noise * mask
        return _GaussianForward.apply(inputs.contiguous(), scales, means, noise, scale_bound, lik_bound, False)

    def build_indexes(self, scales: Tensor) -> Tensor:
        scale_bound, _ = self._bounds()
        return ops.build_indexes(scales, self.scale_table, scale_bound)
