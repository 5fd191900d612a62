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


class _BottleneckFn(torch.autograd.Function):
    """EntropyBottleneck.forward as one kernel, with one backward kernel for z, the medians and every parameter."""

    @staticmethod
    def forward(ctx, z, noise, medians, lik_bound, filters, *raw):
        nl = len(filters) + 1
        mats, biases, factors = raw[:nl], raw[nl:2 * nl], raw[2 * nl:]
        C = z.shape[1]
        params = torch.cat([t.reshape(C, -1) for t in (*mats, *biases, *factors)], dim=1).contiguous()
        outputs, lik = ops.bottleneck_forward(z, medians, params, filters, noise=noise, lik_bound=lik_bound)
        ctx.save_for_backward(z, noise, medians, params)
        ctx.cfg = (lik_bound, filters, [t.shape for t in raw])
        return outputs, lik

    @staticmethod
    def backward(ctx, g_out, g_lik):
        z, noise, medians, params = ctx.saved_tensors
        lik_bound, filters, shapes = ctx.cfg
        g_out = None if g_out is None else g_out.contiguous()
        g_lik = torch.zeros_like(z) if g_lik is None else g_lik.contiguous()
        g_z, g_params, g_med = ops.bottleneck_backward(z, medians, params, filters, g_lik, g_out, noise, lik_bound)
        grads, o = [], 0
        for shp in shapes:
            cnt = shp[1] * shp[2]
            grads.append(g_params[:, o:o + cnt].reshape(shp))
            o += cnt
        return (g_z, None, g_med, None, None, *grads)


class EntropyBottleneck(EntropyModel):
    """Drop-in for the reference's EntropyBottleneck (entropy_models/entropy_models.py:296-528): same constructor,
    parameter names (_matrix{i}, _bias{i}, _factor{i}, quantiles, target) and methods.  forward() runs as one kernel
    on z where the conv wrote it (no permute), its backward as one kernel that also accumulates the parameter
    gradients; update() / compress() / decompress() keep the reference's logic over the package's coder."""

    _offset: Tensor

    def __init__(self, channels: int, *args: Any, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters: Tuple[int, ...] = (3, 3, 3, 3), **kwargs: Any):
        super().__init__(*args, **kwargs)
        import numpy as np

        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        if len(self.filters) != 4 or any(f < 1 or f > 4 for f in self.filters):
            raise ValueError("the kernel supports the reference's four hidden layers of up to 4 units")
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(self.channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(self.channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(self.channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(self.channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def _raw_parameters(self):
        nl = len(self.filters) + 1
        return ([getattr(self, f"_matrix{i:d}") for i in range(nl)] + [getattr(self, f"_bias{i:d}") for i in range(nl)] +
                [getattr(self, f"_factor{i:d}") for i in range(nl - 1)])

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        """Reference 403-420 (plain torch: used by update() and loss() on a handful of values per channel)."""
        import torch.nn.functional as F

        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                matrix, bias = matrix.detach(), bias.detach()
            logits = torch.matmul(F.softplus(matrix), logits) + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def loss(self) -> Tensor:
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def _likelihood(self, inputs: Tensor) -> Tensor:
        """Reference 422-434 on values laid out [C, 1, N] (its internal layout): no quantisation, no bound."""
        C = inputs.shape[0]
        z = inputs.reshape(C, -1).t().reshape(1, -1, C).permute(0, 2, 1).contiguous()       # [1, C, N]
        zero = torch.zeros_like(z)
        _, lik = _BottleneckFn.apply(z, zero, self._get_medians().reshape(-1).contiguous(), 0.0, self.filters,
                                     *self._raw_parameters())
        return lik.reshape(C, 1, -1)

    def forward(self, x: Tensor, training: Optional[bool] = None) -> Tuple[Tensor, Tensor]:
        if training is None:
            training = self.training
        x = x.contiguous()
        noise = None
        if training:
            # the reference draws the noise on the permuted [C, 1, B*S] tensor: same call, same stream, same layout
            C = x.shape[1]
            perm = torch.empty((C, 1, x.numel() // C), dtype=x.dtype, device=x.device).uniform_(-0.5, 0.5)
            noise = perm.reshape(C, x.shape[0], -1).permute(1, 0, 2).reshape(x.shape).contiguous()
        lik_bound = self.likelihood_lower_bound.value() if self.use_likelihood_bound else 0.0
        return _BottleneckFn.apply(x, noise, self._get_medians().reshape(-1).contiguous(), lik_bound, self.filters,
                                   *self._raw_parameters())

    def update(self, force: bool = False) -> bool:
        """Reference 355-396: per-channel pmf on [median - minima, median + maxima] and its 16-bit CDF."""
        try:
            from compressai._CXX import pmf_to_quantized_cdf as normalise
        except ImportError:
            normalise = codec.pmf_to_quantized_cdf
        medians = self.quantiles[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - self.quantiles[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(self.quantiles[:, 0, 2] - medians).int(), min=0)
        self._offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max().item())
        samples = torch.arange(max_length, device=pmf_start.device)[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative(samples - 0.5, stop_gradient=True)
        upper = self._logits_cumulative(samples + 0.5, stop_gradient=True)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32, device=pmf.device)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            row = normalise(prob.tolist(), self.entropy_coder_precision)
            cdf[i, : len(row)] = torch.tensor(row, dtype=torch.int32)
        self._quantized_cdf = cdf
        self._cdf_length = (pmf_length + 2).int()
        self._invalidate_tables()
        return True

    @staticmethod
    def _build_indexes(size):
        N, C = size[0], size[1]
        view = [1] * len(size)
        view[1] = -1
        return torch.arange(C).view(*view).int().repeat(N, 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def compress(self, x):
        indexes = self._build_indexes(x.size()).to(x.device)
        spatial_dims = len(x.size()) - 2
        medians = self._extend_ndims(self._get_medians().detach(), spatial_dims)
        medians = medians.expand(x.size(0), *([-1] * (spatial_dims + 1)))
        return super().compress(x, indexes, medians, 0)

    def decompress(self, strings, size):
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        indexes = self._build_indexes(output_size).to(self._quantized_cdf.device)
        medians = self._extend_ndims(self._get_medians().detach(), len(size))
        medians = medians.expand(len(strings), *([-1] * (len(size) + 1)))
        return super().decompress(strings, indexes, medians, 0)


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
