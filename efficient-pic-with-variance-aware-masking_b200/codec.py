"""Entropy-coder side of the codec over include/pic_codec.h.

Host code: the reference reaches CompressAI's C++ extension through Python lists here
(entropy_models.py:175-183 `pmf_to_quantized_cdf`, 230-236 `encode_with_indexes`, 280-286
`decode_with_indexes`).  This module keeps those call signatures (`RansCoder` is a drop-in for
`entropy_models._EntropyCoder`) and adds the tensor form the latent path wants: int32 symbols /
indexes produced on the GPU are copied to pinned host memory ONCE and every stream is coded on a
host thread, without `.tolist()`.

Byte parity with compressai itself is unpinned (compressai is not installed here); the coder is
checked against oracle/rans_oracle.py and by round trips (tests/test_codec.py).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import check, lib


def _i32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().to("cpu", torch.int32).numpy()
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pmf_to_quantized_cdf(pmf, precision: int = 16) -> List[int]:
    """compressai._CXX.pmf_to_quantized_cdf(pmf: list[float], precision) -> list[int]."""
    p = np.ascontiguousarray(pmf.detach().cpu().numpy() if isinstance(pmf, torch.Tensor) else pmf, dtype=np.float32)
    bad = p[(p < 0) | ~np.isfinite(p)]
    if bad.size:
        raise ValueError(f"Invalid `pmf`, non-finite or negative element found: {bad[0]}")
    out = np.empty(p.size + 1, dtype=np.int32)
    rc = lib().pic_pmf_to_quantized_cdf(_p(p), int(p.size), int(precision), _p(out))
    if rc != 0:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability.")
    return out.tolist()


class CdfTables:
    """(quantized_cdf [n_cdfs, stride], cdf_length [n_cdfs], offset [n_cdfs]) as contiguous host int32 arrays."""

    def __init__(self, cdf, cdf_length, offset):
        self.cdf = _i32(cdf)
        if self.cdf.ndim != 2:
            raise ValueError(f"Invalid CDF size {tuple(self.cdf.shape)}")
        self.length = _i32(cdf_length).reshape(-1)
        self.offset = _i32(offset).reshape(-1)
        if not (self.cdf.shape[0] == self.length.size == self.offset.size):
            raise ValueError("cdf, cdf_length and offset disagree on the number of tables")

    def args(self) -> Tuple:
        return (_p(self.cdf), int(self.cdf.shape[0]), int(self.cdf.shape[1]), _p(self.length), _p(self.offset))


def encode_streams(symbols, indexes, tables: CdfTables, threads: int = 0) -> List[bytes]:
    """symbols, indexes: [streams, ...] int32 (CUDA or CPU tensors, or arrays) -> one byte string per stream."""
    s, ix = _to_host_pair(symbols, indexes)
    streams, n = s.shape
    stride = int(lib().pic_rans_stream_bound(n))
    out = np.empty((streams, stride), dtype=np.uint8)
    nbytes = np.zeros(streams, dtype=np.int64)
    check(lib().pic_rans_encode_batch(_p(s), _p(ix), streams, n, *tables.args(), _p(out), stride, _p(nbytes), threads),
          "pic_rans_encode_batch")
    return [out[i, : nbytes[i]].tobytes() for i in range(streams)]


def decode_streams(strings: Sequence[bytes], indexes, tables: CdfTables, threads: int = 0) -> torch.Tensor:
    """-> int32 CPU tensor shaped like `indexes` ([streams, ...])."""
    ix = _i32(indexes)
    streams = len(strings)
    if streams != ix.shape[0]:
        raise ValueError("Invalid strings or indexes parameters")
    ix2 = ix.reshape(streams, -1)
    sizes = np.array([len(b) for b in strings], dtype=np.int64)
    padded = (sizes + 3) // 4 * 4
    offs = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    blob = np.zeros(int(padded.sum()) + 4, dtype=np.uint8)
    for o, b in zip(offs, strings):
        blob[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    out = np.empty_like(ix2)
    check(lib().pic_rans_decode_batch(_p(blob), _p(offs), _p(sizes), _p(ix2), streams, ix2.shape[1], *tables.args(),
                                      _p(out), threads), "pic_rans_decode_batch")
    return torch.from_numpy(out.reshape(ix.shape))


def _to_host_many(*tensors) -> List[np.ndarray]:
    """[streams, ...] int32 tensors / arrays -> contiguous host [streams, n] arrays; CUDA tensors share one pinned copy."""
    first = tensors[0]
    if isinstance(first, torch.Tensor) and first.is_cuda:
        both = torch.stack([t.to(torch.int32).reshape(first.shape[0], -1) for t in tensors])
        host = torch.empty(both.shape, dtype=torch.int32, pin_memory=True)
        host.copy_(both, non_blocking=False)
        return [host[i].numpy() for i in range(len(tensors))]
    arrs = [_i32(t) for t in tensors]
    if any(a.shape != arrs[0].shape for a in arrs):
        raise ValueError("symbols, indexes and level should have the same size.")
    return [a.reshape(a.shape[0], -1) for a in arrs]


def encode_levels(symbols, indexes, level, n_levels: int, tables: CdfTables, level_begin: int = 0,
                  threads: int = 0) -> List[List[bytes]]:
    """Progressive multi-level packing (functions_encode.py:176-196): bitstream[l][s] == the reference's
    compress(symbols * delta_l, indexes * delta_l)[s] with delta_l = (level == l), for l in [level_begin, n_levels).
    symbols / indexes / level cross PCIe once for all levels; (level, stream) pairs run on host threads."""
    s, ix, lv = _to_host_many(symbols, indexes, level)
    streams, n = s.shape
    stride = int(lib().pic_rans_stream_bound(n))
    tasks = (n_levels - level_begin) * streams
    out = np.empty((tasks, stride), dtype=np.uint8)
    nbytes = np.zeros(tasks, dtype=np.int64)
    check(lib().pic_rans_encode_levels(_p(s), _p(ix), _p(lv), streams, n, level_begin, n_levels, *tables.args(), _p(out),
                                       stride, _p(nbytes), threads), "pic_rans_encode_levels")
    return [[out[l * streams + k, : nbytes[l * streams + k]].tobytes() for k in range(streams)]
            for l in range(n_levels - level_begin)]


def decode_levels(bitstream: Sequence[Sequence[bytes]], indexes, level, tables: CdfTables, level_begin: int = 0,
                  out: Optional[torch.Tensor] = None, threads: int = 0) -> torch.Tensor:
    """Inverse of encode_levels for the levels received so far: int32 symbols (CPU tensor shaped like `indexes`),
    zero (or `out`'s previous content) where the element's level has not been received."""
    ix, lv = _to_host_many(indexes, level)
    streams, n = ix.shape
    flat = [b for lvl in bitstream for b in lvl]
    if any(len(lvl) != streams for lvl in bitstream):
        raise ValueError("Invalid strings or indexes parameters")
    sizes = np.array([len(b) for b in flat], dtype=np.int64)
    padded = (sizes + 3) // 4 * 4
    offs = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    blob = np.zeros(int(padded.sum()) + 4, dtype=np.uint8)
    for o, b in zip(offs, flat):
        blob[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    shape = tuple(indexes.shape)
    res = np.zeros((streams, n), dtype=np.int32) if out is None else _i32(out).reshape(streams, n).copy()
    check(lib().pic_rans_decode_levels(_p(blob), _p(offs), _p(sizes), _p(ix), _p(lv), streams, n, level_begin,
                                       level_begin + len(bitstream), *tables.args(), _p(res), threads),
          "pic_rans_decode_levels")
    return torch.from_numpy(res.reshape(shape))


def _to_host_pair(symbols, indexes) -> Tuple[np.ndarray, np.ndarray]:
    if isinstance(symbols, torch.Tensor) and symbols.is_cuda:
        # one D2H copy of both tensors through pinned memory instead of per-element .tolist()
        both = torch.stack([symbols.to(torch.int32).reshape(symbols.shape[0], -1),
                            indexes.to(torch.int32).reshape(indexes.shape[0], -1)])
        host = torch.empty(both.shape, dtype=torch.int32, pin_memory=True)
        host.copy_(both, non_blocking=False)
        return host[0].numpy(), host[1].numpy()
    s, ix = _i32(symbols), _i32(indexes)
    if s.shape != ix.shape:
        raise ValueError("`inputs` and `indexes` should have the same size.")
    return s.reshape(s.shape[0], -1), ix.reshape(ix.shape[0], -1)


class RansCoder:
    """Same methods as the reference's `_EntropyCoder` proxy (entropy_models.py:18-52) over compressai.ans:
    list (or array / tensor) arguments, one stream per call."""

    name = "ans"

    def encode_with_indexes(self, symbols, indexes, cdf, cdf_length, offset) -> bytes:
        s, ix = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
        if s.size != ix.size:
            raise ValueError("`symbols` and `indexes` should have the same size.")
        t = cdf if isinstance(cdf, CdfTables) else CdfTables(cdf, cdf_length, offset)
        cap = int(lib().pic_rans_stream_bound(s.size))
        out = np.empty(cap, dtype=np.uint8)
        n = int(lib().pic_rans_encode_with_indexes(_p(s), _p(ix), s.size, *t.args(), _p(out), cap))
        if n < 0:
            check(n, "pic_rans_encode_with_indexes")
        return out[:n].tobytes()

    def decode_with_indexes(self, stream: bytes, indexes, cdf, cdf_length, offset) -> List[int]:
        ix = _i32(indexes).reshape(-1)
        t = cdf if isinstance(cdf, CdfTables) else CdfTables(cdf, cdf_length, offset)
        buf = np.frombuffer(bytes(stream) + b"\0" * ((-len(stream)) % 4), dtype=np.uint8).copy()
        out = np.empty(ix.size, dtype=np.int32)
        check(lib().pic_rans_decode_with_indexes(_p(buf), len(stream), _p(ix), ix.size, *t.args(), _p(out)),
              "pic_rans_decode_with_indexes")
        return out.tolist()
