"""TEST INFRASTRUCTURE -- BASELINE config[3]: runs the UNMODIFIED random-init reference REM model
(VarianceMaskingPICREM, division_dimension = [320, 640] as in the base model, mu_std = True, dimension = "middle",
check_levels = [0.75]) on one synthetic 256x256 RGB
image on CPU and records EVERY call the model makes into the hot path, in call order:

  * masking(scale, pr=..)                      models/rem_pic.py:182-189 (bar / star masks of apply_latent_enhancement),
                                                382-391 (block mask on the REM-refined scale), and the checkpoint pass
  * the attention mask handed to the REM       models/rem_pic.py:191-195 -> layers/rem.py:137-140 ([B,64,h,w] for mu_std)
  * gaussian_conditional(inputs, scales)       models/rem_pic.py:388-389 and the base slices
  * gaussian_conditional.build_indexes(scale)  when the model's compress path is reachable (it is not: no rANS here)

    python oracle/gen_golden_model_rem.py        # writes tests/golden/model_rem.npz
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

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "model_rem.npz")


def main():
    base.install_stubs()
    from models.pic import get_scale_table
    from models.rem_pic import VarianceMaskingPICREM

    torch.manual_seed(0)
    torch.set_num_threads(4)
    net = VarianceMaskingPICREM(division_dimension=[320, 640], check_levels=[0.75], mu_std=True, dimension="middle")
    # the checkpoint pass goes through the model's compress(): the absent CompressAI C++ entry points are supplied by
    # oracle/rans_oracle.py exactly as in gen_golden_codec.py (the bytes are not recorded here, only the latent path)
    import entropy_models.entropy_models as em
    em._pmf_to_quantized_cdf = rans_oracle.pmf_to_quantized_cdf
    net.gaussian_conditional.entropy_coder = _OracleCoder()
    net.entropy_bottleneck.entropy_coder = _OracleCoder()
    net.entropy_bottleneck.update(force=True)
    net.gaussian_conditional.update_scale_table(get_scale_table())
    net.eval()
    idx_calls = []
    real_build = net.gaussian_conditional.build_indexes

    def build_indexes_recorded(scales):
        o = real_build(scales)
        idx_calls.append((scales.detach().clone(), o.detach().clone().int()))
        return o
    net.gaussian_conditional.build_indexes = build_indexes_recorded
    x = torch.rand(1, 3, 256, 256)
    out = {}
    for q_name, pr in (("pr2.5", 2.5), ("pr5", 5.0)):
        masks, atts, gcs = [], [], []
        idx_calls.clear()

        def mask_hook(m, args, kwargs, o):
            scale = args[0] if args else kwargs["scale"]
            masks.append((scale.detach().clone(), float(kwargs.get("pr", args[1] if len(args) > 1 else 0)), o.detach().clone()))

        def gc_hook(m, args, kwargs, o):
            means = kwargs.get("means", args[2] if len(args) > 2 else None)
            gcs.append((args[0].detach().clone(), args[1].detach().clone(), o[0].detach().clone(), o[1].detach().clone(),
                        None if means is None else means.detach().clone()))

        def rem_hook(m, args, kwargs):
            atts.append(args[3].detach().clone())

        hooks = [net.masking.register_forward_hook(mask_hook, with_kwargs=True),
                 net.gaussian_conditional.register_forward_hook(gc_hook, with_kwargs=True)]
        for group in net.post_latent:
            for rem in group:
                hooks.append(rem.register_forward_pre_hook(rem_hook, with_kwargs=True))
        # training/step.py:69-79: the checkpoint representation at the REM's quality, then the pass at `pr` with it
        with torch.no_grad():
            ckpt = net.ExtractChekpointRepr(x, quality=0.75, rc=False)
        out[f"{q_name}/n_mask_checkpoint"] = np.asarray(len(masks), np.int32)
        net.forward(x, quality=pr, training=False, checkpoint_ref=ckpt.detach().clone())
        for h in hooks:
            h.remove()
        print(q_name, "masking calls", len(masks), "attention masks", len(atts), "gaussian_conditional calls", len(gcs))
        out[f"{q_name}/pr"] = np.asarray(pr, np.float32)
        out[f"{q_name}/n_mask"] = np.asarray(len(masks), np.int32)
        out[f"{q_name}/n_att"] = np.asarray(len(atts), np.int32)
        out[f"{q_name}/n_gc"] = np.asarray(len(gcs), np.int32)
        for i, (scale, p, m) in enumerate(masks):
            out[f"{q_name}/mask{i}/scale"] = scale.numpy()
            out[f"{q_name}/mask{i}/pr"] = np.asarray(p, np.float64)
            out[f"{q_name}/mask{i}/mask"] = np.packbits(m.numpy().astype(np.uint8).ravel())
        for i, a in enumerate(atts):
            # the attention mask is cat([star_mask] * 2, 1) of a masking call made just before: find it
            half = a[:, : a.shape[1] // 2]
            assert torch.equal(half, a[:, a.shape[1] // 2:])
            src = [j for j, (_, _, m) in enumerate(masks) if m.shape == half.shape and torch.equal(torch.round(m), half)]
            assert src, "attention mask without a source masking call"
            out[f"{q_name}/att{i}/src"] = np.asarray(src[-1], np.int32)
            out[f"{q_name}/att{i}/shape"] = np.asarray(a.shape, np.int32)
            out[f"{q_name}/att{i}/mask"] = np.packbits(a.numpy().astype(np.uint8).ravel())
        if q_name != "pr2.5":      # the second quality keeps the masks and attention masks only (file size)
            gcs, idx_calls[:] = [], []
            out[f"{q_name}/n_gc"] = np.asarray(0, np.int32)
        out[f"{q_name}/n_idx"] = np.asarray(len(idx_calls), np.int32)
        for i, (sc, ix) in enumerate(idx_calls):      # compress side of the checkpoint pass (rem_pic.py:588-590)
            out[f"{q_name}/idx{i}/scales"], out[f"{q_name}/idx{i}/idx"] = sc.numpy(), ix.numpy()
        for i, (inp, sc, o, lik, means) in enumerate(gcs):
            out[f"{q_name}/gc{i}/inputs"], out[f"{q_name}/gc{i}/scales"] = inp.numpy(), sc.numpy()
            if means is not None:
                out[f"{q_name}/gc{i}/means"] = means.numpy()
            out[f"{q_name}/gc{i}/outputs"], out[f"{q_name}/gc{i}/lik"] = o.numpy(), lik.numpy()
            # encoder / decoder agreement on the index (models/pic.py:813 vs 942-946): build_indexes of the same scale
            out[f"{q_name}/gc{i}/idx"] = real_build(sc).int().numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
