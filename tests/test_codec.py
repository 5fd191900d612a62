"""Entropy-coder side: the C++ coder of include/pic_codec.h against oracle/rans_oracle.py,
the golden vectors produced by the reference's own update / compress / decompress (oracle/gen_golden_codec.py)
and round trips.  Host code only: runs without a GPU.  Byte parity with compressai itself is unpinned."""
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

import pic_b200
import rans_oracle as ro
from pic_b200 import codec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "codec.npz")


@pytest.fixture(scope="module", autouse=True)
def _built():
    pic_b200.build()


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="module")
def tables(gold):
    return codec.CdfTables(gold["cdf"], gold["cdf_length"], gold["offset"])


def test_pmf_to_quantized_cdf_known_answers():
    assert codec.pmf_to_quantized_cdf([0.5, 0.5], 16) == [0, 32768, 65536]
    assert codec.pmf_to_quantized_cdf([0.25, 0.25, 0.5], 16) == [0, 16384, 32768, 65536]
    assert codec.pmf_to_quantized_cdf([0.0, 1.0], 4) == [0, 1, 16]            # zero-width slot steals one count
    assert codec.pmf_to_quantized_cdf([1.0, 0.0, 0.0], 4) == [0, 14, 15, 16]  # ... from a slot below it
    with pytest.raises(ValueError):
        codec.pmf_to_quantized_cdf([0.5, -0.1], 16)
    with pytest.raises(ValueError):
        codec.pmf_to_quantized_cdf([0.5, float("nan")], 16)
    with pytest.raises(ValueError):
        codec.pmf_to_quantized_cdf([0.0, 0.0], 16)


@pytest.mark.parametrize("seed", range(6))
def test_pmf_to_quantized_cdf_vs_oracle(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 400))
    pmf = rng.random(n).astype(np.float32) ** 8        # many near-zero entries -> the stealing loop runs
    pmf[rng.random(n) < 0.3] = 0.0
    pmf[0] += 1e-3
    pmf /= pmf.sum()
    got = codec.pmf_to_quantized_cdf(pmf.tolist(), 16)
    assert got == ro.pmf_to_quantized_cdf(pmf.tolist(), 16)
    assert got[0] == 0 and got[-1] == 65536 and all(b > a for a, b in zip(got, got[1:]))


def test_update_builds_the_reference_tables(gold):
    """GaussianConditional.update(): tables identical to the reference's update() (golden), buffer names kept."""
    gc = pic_b200.GaussianConditional(None)
    gc.update(gold["scale_table"].tolist())
    assert np.array_equal(gc._quantized_cdf.numpy(), gold["cdf"])
    assert np.array_equal(gc._cdf_length.numpy(), gold["cdf_length"])
    assert np.array_equal(gc._offset.numpy(), gold["offset"])
    assert gc._quantized_cdf.dtype == torch.int32 and set(gc.state_dict()) >= {"_offset", "_quantized_cdf", "_cdf_length"}
    for row, ln in zip(gold["cdf"], gold["cdf_length"]):
        assert row[0] == 0 and row[ln - 1] == 65536 and np.all(np.diff(row[:ln]) > 0)


def test_golden_streams_byte_exact_and_decodable(gold, tables):
    """Streams of the reference's compress() (oracle coder) == the C++ coder's, single-stream and batched."""
    sym, idx = gold["symbols"], gold["indexes"]
    offs = np.concatenate([[0], np.cumsum(gold["stream_bytes"])])
    want = [gold["stream_blob"][offs[i]:offs[i + 1]].tobytes() for i in range(len(sym))]
    coder = codec.RansCoder()
    lists = (gold["cdf"].tolist(), gold["cdf_length"].tolist(), gold["offset"].tolist())
    for i in range(len(sym)):
        got = coder.encode_with_indexes(sym[i].reshape(-1).tolist(), idx[i].reshape(-1).tolist(), *lists)
        assert got == want[i], i
        assert coder.decode_with_indexes(got, idx[i].reshape(-1).tolist(), *lists) == sym[i].reshape(-1).tolist()
    assert codec.encode_streams(torch.from_numpy(sym), torch.from_numpy(idx), tables, threads=3) == want
    back = codec.decode_streams(want, torch.from_numpy(idx), tables, threads=2)
    assert back.dtype == torch.int32 and np.array_equal(back.numpy(), sym)


def test_compress_decompress_module_api(gold):
    """EntropyModel.compress(already_quantize=True) / decompress streams through the native coder."""
    gc = pic_b200.GaussianConditional(None)
    gc.update(gold["scale_table"].tolist())
    sym, idx = torch.from_numpy(gold["symbols"]), torch.from_numpy(gold["indexes"])
    strings = gc.compress(sym, idx, already_quantize=True)
    offs = np.concatenate([[0], np.cumsum(gold["stream_bytes"])])
    assert strings == [gold["stream_blob"][offs[i]:offs[i + 1]].tobytes() for i in range(len(sym))]
    fresh = pic_b200.GaussianConditional(None)
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        fresh.decompress(strings, idx)
    with pytest.raises(ValueError, match="same size"):
        gc.compress(sym, idx[:, :4], already_quantize=True)
    with pytest.raises(ValueError, match="Invalid strings or indexes"):
        gc.decompress(strings[:2], idx)


@pytest.mark.parametrize("seed", range(4))
def test_streams_vs_oracle_seeded(seed, gold, tables):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 3000))
    idx = rng.integers(0, 64, size=n).astype(np.int32)
    scale = gold["scale_table"][idx]
    sym = np.rint(rng.normal(0, 1, n) * scale * rng.choice([1, 1, 1, 8], size=n)).astype(np.int32)  # some escapes
    lists = (gold["cdf"].tolist(), gold["cdf_length"].tolist(), gold["offset"].tolist())
    want = ro.encode_with_indexes(sym.tolist(), idx.tolist(), *lists)
    got = codec.RansCoder().encode_with_indexes(sym, idx, tables, None, None)
    assert got == want
    assert ro.decode_with_indexes(got, idx.tolist(), *lists) == sym.tolist()          # oracle decodes C++ stream
    assert codec.RansCoder().decode_with_indexes(want, idx, tables, None, None) == sym.tolist()


@settings(max_examples=40, deadline=None)
@given(st.lists(st.tuples(st.integers(-(2 ** 31) + 4000, 2 ** 31 - 4000), st.integers(0, 63)), min_size=0, max_size=200))
def test_round_trip_property(pairs):
    g = dict(np.load(GOLDEN))
    t = codec.CdfTables(g["cdf"], g["cdf_length"], g["offset"])
    sym = np.array([p[0] for p in pairs], dtype=np.int32)
    idx = np.array([p[1] for p in pairs], dtype=np.int32)
    c = codec.RansCoder()
    s = c.encode_with_indexes(sym, idx, t, None, None)
    assert len(s) % 4 == 0 and len(s) >= 8
    assert c.decode_with_indexes(s, idx, t, None, None) == sym.tolist()


def test_errors(tables):
    c = codec.RansCoder()
    with pytest.raises(ValueError):
        c.encode_with_indexes([0, 1], [0, 64], tables, None, None)          # CDF index out of range
    s = c.encode_with_indexes([3, -2, 0, 1] * 50, [5, 9, 0, 63] * 50, tables, None, None)
    with pytest.raises(ValueError):
        c.decode_with_indexes(s[:4], [5, 9, 0, 63] * 50, tables, None, None)   # shorter than the rANS state
    with pytest.raises(ValueError):
        c.decode_with_indexes(s[:12], [5, 9, 0, 63] * 50, tables, None, None)  # truncated: runs out of words
    with pytest.raises(ValueError):
        codec.CdfTables(np.zeros(5, np.int32), [5], [0])


def test_level_streams_equal_per_level_masked_streams(gold, tables):
    """encode_levels(symbols, indexes, level) == compress(symbols * delta_l, indexes * delta_l) for every level
    (functions_encode.py:176-196), and decode_levels restores exactly the received levels."""
    rng = np.random.default_rng(9)
    streams, n, levels = 5, 1777, 4
    idx = rng.integers(0, 64, size=(streams, n)).astype(np.int32)
    sym = np.rint(rng.normal(0, 1, (streams, n)) * gold["scale_table"][idx] * rng.choice([1, 1, 6], size=(streams, n)))
    sym = sym.astype(np.int32)
    level = rng.integers(0, levels + 1, size=(streams, n)).astype(np.int32)   # value `levels` = never sent
    bits = codec.encode_levels(sym, idx, level, levels, tables, threads=3)
    assert len(bits) == levels and all(len(b) == streams for b in bits)
    for l in range(levels):
        d = (level == l).astype(np.int32)
        assert bits[l] == codec.encode_streams(sym * d, idx * d, tables), l
    got = codec.decode_levels(bits[:2], idx, level, tables)
    assert np.array_equal(got.numpy(), sym * (level < 2))
    more = codec.decode_levels(bits[2:], idx, level, tables, level_begin=2, out=got)   # the next levels arrive
    assert np.array_equal(more.numpy(), sym * (level < levels))
    part = codec.encode_levels(sym, idx, level, levels, tables, level_begin=2)
    assert part == bits[2:]


def test_empty_streams(gold, tables):
    """n = 0: the stream is the bare rANS state (8 bytes), as the oracle writes it."""
    lists = (gold["cdf"].tolist(), gold["cdf_length"].tolist(), gold["offset"].tolist())
    empty = np.zeros(0, np.int32)
    s = codec.RansCoder().encode_with_indexes(empty, empty, tables, None, None)
    assert s == ro.encode_with_indexes([], [], *lists) and len(s) == 8
    assert codec.RansCoder().decode_with_indexes(s, empty, tables, None, None) == []
    assert codec.encode_streams(np.zeros((2, 0), np.int32), np.zeros((2, 0), np.int32), tables) == [s, s]


def test_decoder_survives_garbage(tables):
    """Corrupt / random streams must come back as an error or as some symbols -- never crash or hang."""
    rng = np.random.default_rng(11)
    c = codec.RansCoder()
    idx = rng.integers(0, 64, size=500).astype(np.int32)
    for k in range(200):
        blob = rng.integers(0, 256, size=int(rng.integers(8, 400)) // 4 * 4, dtype=np.uint8).tobytes()
        if k % 4 == 0:
            blob = b"\xff" * len(blob)          # all-ones state: every escape counter saturates
        try:
            out = c.decode_with_indexes(blob, idx, tables, None, None)
            assert len(out) == idx.size
        except ValueError:
            pass
