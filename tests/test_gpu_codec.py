"""GPU end-to-end of the codec side: latent path (CUDA) -> per-level symbol / index streams -> native rANS coder
-> decode -> reconstruction, following test/functions_encode.py:153-196 and functions_decode.py:186-206."""
import numpy as np
import pytest
import torch

import pic_oracle as po
import rans_oracle as ro
from _common import scale_table, trained_like

pytestmark = pytest.mark.gpu


def test_progressive_encode_decode_round_trip():
    import pic_b200 as pic

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(31)
    slices, shape = 10, (1, 32, 8, 12)
    y_top, y_base, mu, std = (np.stack(a) for a in zip(*[trained_like(rng, shape) for _ in range(slices)]))
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    table = scale_table()
    gc = pic.GaussianConditional(None).to(dev)
    gc.update(table.tolist())
    assert gc._quantized_cdf.is_cuda and gc._quantized_cdf.shape[0] == 64

    # encoder: symbols / indexes of every slice in one launch (mask of ones: pr = 10), then the level map
    full = pic.ops.slice_forward(T(y_top), T(y_base), T(mu), T(std), slices, pic.ops.pr_to_q01(10), T(table),
                                 want=("y_hat", "idx", "symbols"))
    ref = po.slice_forward(y_top.reshape(slices, -1), y_base.reshape(slices, -1), mu.reshape(slices, -1),
                           std.reshape(slices, -1), 10, table)
    symbols, indexes = full["symbols"].reshape(slices, *shape[1:]), full["idx"].reshape(slices, *shape[1:])
    assert np.array_equal(symbols.cpu().numpy().reshape(slices, -1), ref["symbols"])
    q_list = [0.5, 2.5, 5, 10]
    masking = pic.ChannelMask("point-based-std")
    level, _ = masking.ProgLevels([T(s) for s in std], q_list)
    bitstream, total_bytes = [], 0
    for l in range(len(q_list)):
        delta = (level == l).to(torch.int32)
        strings = gc.compress(symbols * delta, indexes * delta, already_quantize=True)   # 10 streams per level
        assert len(strings) == slices and all(isinstance(s, bytes) and len(s) % 4 == 0 for s in strings)
        bitstream.append(strings)
        total_bytes += sum(len(s) for s in strings)
    # the streams are the ones the oracle coder writes for the same symbols (level 1, slice 3)
    delta = (level == 1).to(torch.int32)
    lists = (gc._quantized_cdf.cpu().tolist(), gc._cdf_length.cpu().tolist(), gc._offset.cpu().tolist())
    want = ro.encode_with_indexes((symbols * delta)[3].reshape(-1).cpu().tolist(),
                                  (indexes * delta)[3].reshape(-1).cpu().tolist(), *lists)
    assert bitstream[1][3] == want

    # decoder: knows std (hence indexes and the level map), receives the streams level by level
    dec_idx = gc.build_indexes(T(std).reshape(slices, *shape[1:]))
    assert torch.equal(dec_idx, indexes)
    r_hat = torch.zeros(symbols.shape, device=dev)
    kept = torch.zeros(symbols.shape, device=dev)
    for l, strings in enumerate(bitstream):
        delta = (level == l)
        values = gc.decompress(strings, dec_idx * delta.to(torch.int32))
        assert values.dtype == torch.float32 and values.is_cuda
        r_hat += values * delta
        kept += delta
        # progressive property: after level l the reconstruction equals the masked slice at quality q_list[l]
        part = pic.ops.slice_forward(T(y_top), T(y_base), T(mu), T(std), slices, pic.ops.pr_to_q01(q_list[l]), None,
                                     want=("y_hat", "mask"))
        assert torch.equal(part["mask"].reshape(kept.shape), kept)
        assert torch.equal(part["y_hat"].reshape(kept.shape), r_hat + T(mu).reshape(kept.shape))
    # level-aware coder: symbols / indexes / level cross PCIe once; streams identical to the per-level compress()
    packed = pic.codec.encode_levels(symbols, indexes, level, len(q_list), gc._tables())
    assert packed == bitstream
    got = pic.codec.decode_levels(packed[:2], dec_idx, level, gc._tables())
    assert torch.equal(got.to(dev), symbols * (level < 2))
    assert torch.equal(r_hat, symbols.float())
    assert torch.equal(r_hat + T(mu).reshape(r_hat.shape), full["y_hat"].reshape(r_hat.shape))
    assert 0 < total_bytes < symbols.numel() * 4
