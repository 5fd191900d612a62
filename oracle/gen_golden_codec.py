"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/codec.npz (run in the build container, where
/root/reference exists):  python oracle/gen_golden_codec.py

Runs the UNMODIFIED reference `GaussianConditional.update / compress / decompress`
(entropy_models/entropy_models.py:206-294, 591-618) with the two entry points of the absent CompressAI
C++ extension supplied by oracle/rans_oracle.py (`pmf_to_quantized_cdf`, `encode/decode_with_indexes`),
and stores the CDF tables, inputs and byte streams.  The reference's torch code (pmf in f32, offsets,
lengths, the per-stream loop and list conversions) is therefore the real thing; the coder arithmetic is
the published-algorithm restatement (parity with compressai's bytes unpinned, see rans_oracle.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import rans_oracle  # noqa: E402
import ref_shim  # noqa: E402


class _OracleCoder:
    def encode_with_indexes(self, symbols, indexes, cdf, cdf_length, offset):
        return rans_oracle.encode_with_indexes(symbols, indexes, cdf, cdf_length, offset)

    def decode_with_indexes(self, stream, indexes, cdf, cdf_length, offset):
        return rans_oracle.decode_with_indexes(stream, indexes, cdf, cdf_length, offset)


def main():
    ref = ref_shim.load_reference()
    em = ref.entropy_models_module
    em._pmf_to_quantized_cdf = rans_oracle.pmf_to_quantized_cdf   # the name the reference imports from compressai._CXX
    gc = ref.GaussianConditional(None)
    gc.entropy_coder = _OracleCoder()
    gc.update(ref_shim.get_scale_table())
    out = {"cdf": gc._quantized_cdf.numpy().astype(np.int32), "cdf_length": gc._cdf_length.numpy().astype(np.int32),
           "offset": gc._offset.numpy().astype(np.int32), "scale_table": gc.scale_table.numpy()}

    rng = np.random.default_rng(2024)
    streams, shape = 4, (4, 8, 6, 5)                      # [streams, 8, 6, 5] like [10, 32, h, w]
    scales = np.exp(rng.normal(-1.0, 1.6, size=shape)).clip(0.05, 300).astype(np.float32)
    indexes = gc.build_indexes(torch.from_numpy(scales)).int()
    sym = np.rint(rng.normal(0, 1, size=shape) * scales)
    sym[0, 0, 0, :5] = [4000, -4000, 70000, -70000, 2 ** 24]   # escapes of several nibble counts
    sym[1] = 0                                                  # an all-zero stream (masked-out level)
    symbols = torch.from_numpy(sym.astype(np.int32))
    strings = gc.compress(symbols, indexes, already_quantize=True)
    back = gc.decompress(strings, indexes)
    assert torch.equal(back.int(), symbols), "oracle coder does not round-trip through the reference code"
    out.update(symbols=symbols.numpy(), indexes=indexes.numpy(),
               stream_bytes=np.array([len(s) for s in strings], dtype=np.int64),
               stream_blob=np.frombuffer(b"".join(strings), dtype=np.uint8))
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "codec.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()}, "streams", [len(s) for s in strings])


if __name__ == "__main__":
    main()
