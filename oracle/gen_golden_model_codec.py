"""TEST INFRASTRUCTURE -- the reference model's own codec loops (models/pic.py:769-831 compress, 905-960 decompress):
runs the UNMODIFIED random-init VarianceMaskingPIC on one synthetic 256x256 RGB image on CPU through
`compress(x, quality)` and `decompress(strings, shape, quality)` -- the rANS coder CompressAI would supply is
oracle/rans_oracle.py, exactly as in gen_golden_codec.py -- and records, for every progressive slice and in call order,
what the encoder and the decoder hand to the latent path and get back:

  encoder (pic.py:809-820)  masking(scale, pr)            -> block mask
                            build_indexes(scale * mask)   -> index
                            quantize((y - mu) * mask, "symbols") -> symbols
  decoder (pic.py:942-948)  masking(scale, pr)            -> block mask      (scale recomputed from decoded slices)
                            build_indexes(scale * mask)   -> index
                            decompress(strings, index)    -> rv ;  dequantize(rv, mu) -> y_hat

The generator itself asserts the encoder / decoder agreement of the reference (same scale, mask and index on both sides,
decoded symbols == encoded symbols); tests/test_gpu_parity.py::test_codec_loops_of_the_reference_model replays the
recorded calls on the CUDA drop-in.

    python oracle/gen_golden_model_codec.py        # writes tests/golden/model_codec.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden_model as base  # noqa: E402  (stubs for the absent third-party packages)
import rans_oracle  # noqa: E402
from gen_golden_codec import _OracleCoder  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "model_codec.npz")


def main():
    base.install_stubs()
    from models.pic import VarianceMaskingPIC, get_scale_table

    torch.manual_seed(0)
    torch.set_num_threads(4)
    net = VarianceMaskingPIC()
    import entropy_models.entropy_models as em
    em._pmf_to_quantized_cdf = rans_oracle.pmf_to_quantized_cdf
    net.gaussian_conditional.entropy_coder = _OracleCoder()
    net.entropy_bottleneck.entropy_coder = _OracleCoder()
    net.entropy_bottleneck.update(force=True)
    net.gaussian_conditional.update_scale_table(get_scale_table())
    net.eval()
    gc = net.gaussian_conditional
    n_prog = net.ns1 - net.ns0
    x = torch.rand(1, 3, 256, 256)
    out = {}
    for q_name, pr in (("pr2.5", 2.5), ("pr7", 7.0)):
        calls = {"mask": [], "idx": [], "quant": [], "dec": [], "deq": []}
        real = {"build": gc.build_indexes, "quantize": gc.quantize, "decompress": gc.decompress, "dequantize": gc.dequantize}

        def build_rec(scales):
            o = real["build"](scales)
            calls["idx"].append((scales.detach().clone(), o.detach().clone().int()))
            return o

        def quant_rec(inputs, mode, means=None, mask=None):
            o = real["quantize"](inputs, mode, means, mask) if mask is not None else real["quantize"](inputs, mode, means)
            calls["quant"].append((inputs.detach().clone(), mode, None if means is None else means.detach().clone(), o.detach().clone()))
            return o

        def decompress_rec(strings, indexes, *a, **k):
            o = real["decompress"](strings, indexes, *a, **k)
            calls["dec"].append(o.detach().clone())
            return o

        def dequant_rec(inputs, means=None):
            o = real["dequantize"](inputs, means)
            calls["deq"].append((inputs.detach().clone(), None if means is None else means.detach().clone(), o.detach().clone()))
            return o

        gc.build_indexes, gc.quantize, gc.decompress, gc.dequantize = build_rec, quant_rec, decompress_rec, dequant_rec
        hook = net.masking.register_forward_hook(
            lambda m, a, kw, o: calls["mask"].append(((a[0] if a else kw["scale"]).detach().clone(), o.detach().clone())), with_kwargs=True)
        with torch.no_grad():
            enc = net.compress(x, quality=pr)
        enc_calls = {k: list(v) for k, v in calls.items()}
        for v in calls.values():
            v.clear()
        with torch.no_grad():
            net.decompress(enc["strings"], enc["shape"], quality=pr)
        dec_calls = {k: list(v) for k, v in calls.items()}
        hook.remove()
        for nm in real:
            delattr(gc, {"build": "build_indexes"}.get(nm, nm))      # back to the class's methods
        # the last n_prog calls of each kind belong to the progressive slices
        e_mask, e_idx, e_q = enc_calls["mask"][-n_prog:], enc_calls["idx"][-n_prog:], enc_calls["quant"][-n_prog:]
        d_mask, d_idx, d_rv = dec_calls["mask"][-n_prog:], dec_calls["idx"][-n_prog:], dec_calls["dec"][-n_prog:]
        # gaussian_conditional.decompress() dequantises internally without means; the model's own calls carry mu (pic.py:947)
        d_deq = [c for c in dec_calls["deq"] if c[1] is not None][-n_prog:]
        assert len(e_mask) == len(d_mask) == len(e_idx) == len(d_idx) == len(e_q) == len(d_rv) == len(d_deq) == n_prog
        kept = []
        for k in range(n_prog):
            scale, mask = e_mask[k]
            assert e_q[k][1] == "symbols" and e_q[k][2] is None
            # encoder / decoder agreement of the reference itself
            assert torch.equal(scale, d_mask[k][0]) and torch.equal(mask, d_mask[k][1]), f"slice {k}: decoder scale / mask differ"
            assert torch.equal(e_idx[k][0], scale * torch.round(mask)) and torch.equal(e_idx[k][1], d_idx[k][1])
            assert torch.equal(d_rv[k].reshape(e_q[k][3].shape).int(), e_q[k][3].int()), f"slice {k}: decoded symbols differ"
            tag = f"{q_name}/slice{k}"
            out[f"{tag}/scale"] = scale.numpy()
            out[f"{tag}/mask"] = np.packbits(mask.numpy().astype(np.uint8).ravel())
            out[f"{tag}/idx"] = e_idx[k][1].numpy().astype(np.uint8)              # 0..63
            out[f"{tag}/quant_in"] = e_q[k][0].numpy()                              # (y - mu) * mask
            out[f"{tag}/symbols"] = e_q[k][3].numpy().astype(np.int32)
            out[f"{tag}/mu"] = d_deq[k][1].numpy()
            out[f"{tag}/y_hat"] = d_deq[k][2].numpy()                               # dequantize(rv, mu), before the LRP
            kept.append(float(torch.round(mask).mean()))
        out[f"{q_name}/pr"] = np.asarray(pr, np.float64)
        print(q_name, "progressive slices", n_prog, "kept fraction per slice", [round(v, 3) for v in kept],
              "bytes", sum(len(s[0]) for s in enc["strings"][0][-n_prog:]))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
