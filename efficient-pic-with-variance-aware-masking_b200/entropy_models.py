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


class _EntropyCoder:
    """Proxy to compressai's rANS coder (entropy_models.py:18-52); only built when compressai exists."""

    def __init__(self, method):
        if not isinstance(method, str):
            raise ValueError(f'Invalid method type "{type(method)}"')
        from compressai import available_entropy_coders  # noqa: WPS433 (optional dependency)

        if method not in available_entropy_coders():
            methods = ", ".join(available_entropy_coders())
            raise ValueError(f'Unknown entropy coder "{method}" (available: {methods})')
        if method == "ans":
            from compressai import ans

            encoder, decoder = ans.RansEncoder(), ans.RansDecoder()
        elif method == "rangecoder":
            import range_coder

            encoder, decoder = range_coder.RangeEncoder(), range_coder.RangeDecoder()
        self.name = method
        self._encoder = encoder
        self._decoder = decoder

    def encode_with_indexes(self, *args, **kwargs):
        return self._encoder.encode_with_indexes(*args, **kwargs)

    def decode_with_indexes(self, *args, **kwargs):
        return self._decoder.decode_with_indexes(*args, **kwargs)


def _make_entropy_coder(method):
    try:
        if method is None:
            from compressai import get_entropy_coder

            method = get_entropy_coder()
        return _EntropyCoder(method)
    except ImportError:
        if method not in (None, "ans"):
            raise ValueError(f'Unknown entropy coder "{method}" (available: ans)')
        return codec.RansCoder()  # compressai absent: the native rANS coder of this package


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
    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def _tables(self) -> "codec.CdfTables":
        """Host copy of the CDF tables for the native coder, rebuilt when update() replaces the buffers."""
        key = (self._quantized_cdf.data_ptr(), self._quantized_cdf._version, self._quantized_cdf.numel())
        if getattr(self, "_tables_key", None) != key:
            self._check_cdf_size(), self._check_cdf_length(), self._check_offsets_size()
            self._tables_cache = codec.CdfTables(self._quantized_cdf, self._cdf_length, self._offset)
            self._tables_key = key
        return self._tables_cache

    def compress(self, inputs, indexes, means=None, flag=1, already_quantize=False):
        symbols = self.quantize(inputs, "symbols", means) if already_quantize is False else inputs
        if len(inputs.size()) < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if symbols.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        if isinstance(self.entropy_coder, codec.RansCoder):
            # one D2H copy of the int32 symbols / indexes, every stream (dim 0) on its own host thread
            return codec.encode_streams(symbols, indexes, self._tables())
        strings = []
        for i in range(symbols.size(0)):
            rv = self.entropy_coder.encode_with_indexes(
                symbols[i].reshape(-1).int().tolist(), indexes[i].reshape(-1).int().tolist(),
                self._quantized_cdf.tolist(), self._cdf_length.reshape(-1).int().tolist(),
                self._offset.reshape(-1).int().tolist())
            strings.append(rv)
        return strings

    def decompress(self, strings, indexes, means=None, flag=1):
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if len(indexes.size()) < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        self._check_cdf_size(), self._check_cdf_length(), self._check_offsets_size()
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, len(indexes.size())):
                    if means.size(i) != 1:
                        raise ValueError("Invalid means parameters")
        cdf = self._quantized_cdf
        if isinstance(self.entropy_coder, codec.RansCoder):
            outputs = codec.decode_streams(strings, indexes, self._tables()).to(indexes.device)
            return self.dequantize(outputs, means)
        outputs = cdf.new_empty(indexes.size())
        for i, s in enumerate(strings):
            values = self.entropy_coder.decode_with_indexes(
                s, indexes[i].reshape(-1).int().tolist(), cdf.tolist(),
                self._cdf_length.reshape(-1).int().tolist(), self._offset.reshape(-1).int().tolist())
            outputs[i] = torch.tensor(values, device=outputs.device, dtype=outputs.dtype).reshape(outputs[i].size())
        return self.dequantize(outputs, means)


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
    def __init__(self, scale_table: Optional[Union[List, Tuple]], *args: Any, scale_bound: float = 0.11,
                 tail_mass: float = 1e-9, **kwargs: Any):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = self.scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def _standardized_cumulative(self, inputs: Tensor) -> Tensor:
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        import scipy.stats

        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table):
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update(self.scale_table)
        return True

    def update(self, scale_table):
        """entropy_models.py:591-618: the pmf with the reference's own torch ops (f32, on the host: 64 rows), the
        16-bit CDF rows through compressai's pmf_to_quantized_cdf when importable, else pic_pmf_to_quantized_cdf."""
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        try:
            from compressai._CXX import pmf_to_quantized_cdf as _pmf_to_quantized_cdf
        except ImportError:
            _pmf_to_quantized_cdf = codec.pmf_to_quantized_cdf
        table = self.scale_table.cpu()
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = torch.max(pmf_length).item()
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        samples_scale = table.unsqueeze(1).float()
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            _cdf = torch.IntTensor(_pmf_to_quantized_cdf(prob.tolist(), self.entropy_coder_precision))
            cdf[i, : _cdf.size(0)] = _cdf
        self._quantized_cdf = cdf.to(device)
        self._offset = (-pmf_center).to(device)
        self._cdf_length = (pmf_length + 2).to(device)

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
