"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of the entropy-coder side of the codec
(SURVEY 8(f) rows 2-3).  Only tests/ (and scripts that time the CPU baseline) may import this.

The reference calls an un-vendored third-party C++ extension here: CompressAI 1.2.4
(environment.yml:203): `compressai._CXX.pmf_to_quantized_cdf` (entropy_models.py:175-183) and
`compressai.ans.RansEncoder.encode_with_indexes / RansDecoder.decode_with_indexes`
(entropy_models.py:230-236, 280-286), which wrap ryg_rans' public-domain `rans64.h`.
compressai is NOT installed in this image and there is no network, so this restatement of the
published algorithms (compressai/cpp_exts/ops/ops.cpp, cpp_exts/rans/rans_interface.cpp, rans64.h)
cannot be checked against the real library here: **PARITY UNPINNED** against compressai's bytes.
What it does pin: the product C++ coder (csrc/pic_rans.cpp) must produce byte-identical streams
and integer-identical CDF tables to this file on seeded inputs, and decode(encode(x)) == x.

Bit-stream format (rans_interface.cpp): one rANS64 state, 32-bit renormalisation words, CDF
precision 16 bits; symbols outside a CDF's range are escaped through the last CDF slot and coded
in 4-bit "bypass" nibbles; the encoder runs over the symbols in reverse and the stream is the
little-endian dump of the 32-bit words, state first.
"""
from __future__ import annotations

import math
import struct
from typing import List, Sequence

PRECISION = 16
BYPASS_PRECISION = 4
MAX_BYPASS_VAL = (1 << BYPASS_PRECISION) - 1
RANS64_L = 1 << 31
MASK64 = (1 << 64) - 1


def _f32(x: float) -> float:
    return struct.unpack("f", struct.pack("f", x))[0]


def pmf_to_quantized_cdf(pmf: Sequence[float], precision: int = PRECISION) -> List[int]:
    """ops.cpp pmf_to_quantized_cdf: f32 `round(p * 2^precision)` (half away from zero), renormalise
    to 2^precision with integer division, prefix-sum, force the total, then give every zero-width
    slot one count stolen from the narrowest slot that still has more than one."""
    for p in pmf:
        if p < 0 or not math.isfinite(p):
            raise ValueError(f"Invalid `pmf`, non-finite or negative element found: {p}")
    scale = float(1 << precision)
    cdf = [0] * (len(pmf) + 1)
    for i, p in enumerate(pmf):
        v = _f32(_f32(p) * scale)                  # float multiply in f32
        cdf[i + 1] = int(math.floor(v + 0.5)) & 0xFFFFFFFF   # std::round on a non-negative value -> uint32
    total = sum(cdf) & 0xFFFFFFFF
    if total == 0:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability.")
    cdf = [((1 << precision) * c) // total for c in cdf]
    for i in range(1, len(cdf)):
        cdf[i] += cdf[i - 1]
    cdf[-1] = 1 << precision
    for i in range(len(cdf) - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq, best_steal = 0xFFFFFFFF, -1
            for j in range(len(cdf) - 1):
                freq = cdf[j + 1] - cdf[j]
                if 1 < freq < best_freq:
                    best_freq, best_steal = freq, j
            assert best_steal != -1
            if best_steal < i:
                for j in range(best_steal + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, best_steal + 1):
                    cdf[j] += 1
    assert cdf[0] == 0 and cdf[-1] == (1 << precision)
    assert all(cdf[i + 1] > cdf[i] for i in range(len(cdf) - 1))
    return cdf


def encode_with_indexes(symbols: Sequence[int], indexes: Sequence[int], cdfs: Sequence[Sequence[int]],
                        cdf_sizes: Sequence[int], offsets: Sequence[int]) -> bytes:
    """RansEncoder.encode_with_indexes (= BufferedRansEncoder.encode_with_indexes + flush)."""
    assert len(symbols) == len(indexes)
    syms = []  # (start, range, bypass)
    for s, ci in zip(symbols, indexes):
        assert 0 <= ci < len(cdfs)
        cdf = cdfs[ci]
        max_value = cdf_sizes[ci] - 2
        assert 0 <= max_value < len(cdf) - 1
        value = int(s) - offsets[ci]
        raw_val = 0
        if value < 0:
            raw_val = -2 * value - 1
            value = max_value
        elif value >= max_value:
            raw_val = 2 * (value - max_value)
            value = max_value
        syms.append((cdf[value], cdf[value + 1] - cdf[value], False))
        if value == max_value:
            n_bypass = 0
            while (raw_val >> (n_bypass * BYPASS_PRECISION)) != 0:
                n_bypass += 1
            val = n_bypass
            while val >= MAX_BYPASS_VAL:
                syms.append((MAX_BYPASS_VAL, MAX_BYPASS_VAL + 1, True))
                val -= MAX_BYPASS_VAL
            syms.append((val, val + 1, True))
            for j in range(n_bypass):
                nib = (raw_val >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL
                syms.append((nib, nib + 1, True))
    x = RANS64_L
    words: List[int] = []  # emitted back to front
    for start, rng, bypass in reversed(syms):
        if not bypass:
            x_max = ((RANS64_L >> PRECISION) << 32) * rng
            if x >= x_max:
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x // rng) << PRECISION) + (x % rng) + start
        else:
            freq = 1 << (16 - BYPASS_PRECISION)
            x_max = ((RANS64_L >> 16) << 32) * freq
            if x >= x_max:
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = (x << BYPASS_PRECISION) | start
        assert x <= MASK64
    words.append((x >> 32) & 0xFFFFFFFF)   # Rans64EncFlush: ptr[0] = low word, ptr[1] = high word
    words.append(x & 0xFFFFFFFF)
    words.reverse()
    return struct.pack(f"<{len(words)}I", *words)


def decode_with_indexes(stream: bytes, indexes: Sequence[int], cdfs: Sequence[Sequence[int]],
                        cdf_sizes: Sequence[int], offsets: Sequence[int]) -> List[int]:
    """RansDecoder.decode_with_indexes."""
    assert len(stream) % 4 == 0 and len(stream) >= 8
    words = struct.unpack(f"<{len(stream) // 4}I", stream)
    pos = 2
    x = words[0] | (words[1] << 32)

    def get_bits(n_bits: int) -> int:
        nonlocal x, pos
        val = x & ((1 << n_bits) - 1)
        x >>= n_bits
        if x < RANS64_L:
            x = (x << 32) | words[pos]
            pos += 1
        return val

    out = []
    for ci in indexes:
        cdf = cdfs[ci]
        max_value = cdf_sizes[ci] - 2
        cum = x & ((1 << PRECISION) - 1)
        s = 0
        while s + 1 < cdf_sizes[ci] and cdf[s + 1] <= cum:   # first slot whose upper bound exceeds cum
            s += 1
        start, freq = cdf[s], cdf[s + 1] - cdf[s]
        x = freq * (x >> PRECISION) + (x & ((1 << PRECISION) - 1)) - start
        if x < RANS64_L:
            x = (x << 32) | words[pos]
            pos += 1
        value = s
        if value == max_value:
            val = get_bits(BYPASS_PRECISION)
            n_bypass = val
            while val == MAX_BYPASS_VAL:
                val = get_bits(BYPASS_PRECISION)
                n_bypass += val
            raw_val = 0
            for j in range(n_bypass):
                raw_val |= get_bits(BYPASS_PRECISION) << (j * BYPASS_PRECISION)
            value = raw_val >> 1
            if raw_val & 1:
                value = -value - 1
            else:
                value += max_value
        out.append(value + offsets[ci])
    return out


def gaussian_cdf_tables(scale_table: Sequence[float], tail_mass: float = 1e-9, precision: int = PRECISION):
    """GaussianConditional.update (entropy_models.py:591-618) in float64 numpy-free Python for SMALL tables,
    f32-rounded where the reference computes in f32.  Returns (quantized_cdf rows, cdf_length, offset).
    The product path computes the pmf with the same torch ops as the reference; this restatement is only
    used for its invariants (lengths, offsets, totals), not for bit-exact pmf values."""
    from statistics import NormalDist

    multiplier = -NormalDist().inv_cdf(tail_mass / 2)
    centers = [int(math.ceil(_f32(s) * multiplier)) for s in scale_table]
    lengths = [2 * c + 1 for c in centers]
    max_length = max(lengths)
    rows, out_len, offsets = [], [], []
    for s, c, ln in zip(scale_table, centers, lengths):
        def cum(v):
            return 0.5 * math.erfc(-(2 ** -0.5) * v)
        pmf = [cum((0.5 - abs(k - c)) / s) - cum((-0.5 - abs(k - c)) / s) for k in range(ln)]
        tail = 2 * cum((-0.5 - abs(0 - c)) / s)
        cdf = pmf_to_quantized_cdf(pmf + [tail], precision)
        rows.append(cdf + [0] * (max_length + 2 - len(cdf)))
        out_len.append(ln + 2)
        offsets.append(-c)
    return rows, out_len, offsets
