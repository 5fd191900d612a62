"""TEST INFRASTRUCTURE -- BASELINE config[0]: runs the UNMODIFIED random-init reference model
(VarianceMaskingPIC, defaults N=192 M=640, support_progressive_slices=5) on one synthetic 256x256
RGB image on CPU at quality q in {0.25, 0.5, 1} (the reference's 0..10 `pr` scale: 2.5, 5, 10; at q = 0 the reference returns
before the progressive loop, models/pic.py:556-568, so the hot path is not exercised) and captures,
per progressive slice, the tensors that enter the hot path (y_top, y_base, mu, std as the model
computed them) and what the reference's own per-slice code produced (mask, likelihood, y_hat).

    python oracle/gen_golden_model.py        # writes tests/golden/model_c1.npz

Stubs installed for absent third-party packages (compressai, timm) only touch layers that are OFF
the path (GDN parametrisation, window attention helpers); on-path compressai.ops.LowerBound is the
restatement in oracle/ref_shim.py.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "model_c1.npz")


def install_stubs():
    ref_shim._install_compressai_stub()
    ops = sys.modules["compressai.ops"]
    par = types.ModuleType("compressai.ops.parametrizers")

    class NonNegativeParametrizer(nn.Module):  # compressai 1.2.4 (off-path: GDN only)
        def __init__(self, minimum=0, reparam_offset=2 ** -18):
            super().__init__()
            self.minimum, self.reparam_offset = float(minimum), float(reparam_offset)
            self.register_buffer("pedestal", torch.Tensor([self.reparam_offset ** 2]))
            self.lower_bound = ref_shim.LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)

        def init(self, x):
            return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

        def forward(self, x):
            return self.lower_bound(x) ** 2 - self.pedestal

    par.NonNegativeParametrizer = NonNegativeParametrizer
    ops.parametrizers = par
    sys.modules["compressai.ops.parametrizers"] = par
    timm = types.ModuleType("timm")
    tm = types.ModuleType("timm.models")
    tl = types.ModuleType("timm.models.layers")

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x

    tl.DropPath = DropPath
    tl.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    tl.trunc_normal_ = lambda t, std=1.0, **k: nn.init.trunc_normal_(t, std=std)
    timm.models, tm.layers = tm, tl
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "src"))
    import entropy_models  # the reference's own package

    em = types.ModuleType("compressai.entropy_models")
    em.EntropyBottleneck = entropy_models.EntropyBottleneck
    sys.modules["compressai.entropy_models"] = em
    sys.modules["compressai"].entropy_models = em


def main():
    install_stubs()
    from models.pic import VarianceMaskingPIC, get_scale_table

    torch.manual_seed(0)
    torch.set_num_threads(4)
    net = VarianceMaskingPIC()
    net.gaussian_conditional.scale_table = get_scale_table()
    net.eval()
    x = torch.rand(1, 3, 256, 256)
    out = {}
    for q_name, pr in (("pr2.5", 2.5), ("pr5", 5), ("pr10", 10)):  # q=0 returns before the progressive loop (pic.py:556-568)
        rec = {"mu": [], "std": [], "lrp_in": [], "mask": [], "gc_in": [], "gc_scale": [], "lik": []}
        hooks = []
        for i in range(net.ns0):
            hooks.append(net.cc_mean_transforms_prog[i].register_forward_hook(lambda m, a, o: rec["mu"].append(o.detach())))
            hooks.append(net.cc_scale_transforms_prog[i].register_forward_hook(lambda m, a, o: rec["std"].append(o.detach())))
            hooks.append(net.lrp_transforms_prog[i].register_forward_hook(lambda m, a, o: rec["lrp_in"].append(a[0].detach())))
        hooks.append(net.masking.register_forward_hook(lambda m, a, o: rec["mask"].append(o.detach())))

        def gc_hook(m, a, kw, o):
            rec["gc_in"].append(a[0].detach())
            rec["gc_scale"].append(a[1].detach())
            rec["lik"].append(o[1].detach())
        hooks.append(net.gaussian_conditional.register_forward_hook(gc_hook, with_kwargs=True))
        ys = {}
        hooks.append(net.g_a[0].register_forward_hook(lambda m, a, o: ys.__setitem__(0, o.detach())))
        hooks.append(net.g_a[1].register_forward_hook(lambda m, a, o: ys.__setitem__(1, o.detach())))
        with torch.no_grad():
            net.forward_single_quality(x, quality=pr, training=False)
        for h in hooks:
            h.remove()
        y = torch.cat([ys[0], ys[1]], dim=1)
        y_slices = y.chunk(net.num_slices, 1)
        n_prog = net.ns1 - net.ns0
        # the last n_prog gaussian_conditional / masking calls belong to the progressive slices
        lik, gc_in, gc_scale = rec["lik"][-n_prog:], rec["gc_in"][-n_prog:], rec["gc_scale"][-n_prog:]
        masks = rec["mask"][-n_prog:]
        assert len(rec["mu"]) == n_prog and len(masks) == n_prog
        for k in range(n_prog):
            y_top, y_base = y_slices[net.ns0 + k], y_slices[k]
            mu = rec["mu"][k][:, :, :y.shape[2], :y.shape[3]]
            std = rec["std"][k][:, :, :y.shape[2], :y.shape[3]]
            mask = torch.round(masks[k])
            # sanity: these really are the tensors the reference fed to its entropy model
            assert torch.equal(gc_in[k], (y_top - y_base - mu) * mask)
            assert torch.equal(gc_scale[k], std * mask)
            y_hat = rec["lrp_in"][k][:, -32:]
            tag = f"{q_name}/slice{k}"
            out[f"{tag}/y_top"], out[f"{tag}/y_base"] = y_top.numpy(), y_base.numpy()
            out[f"{tag}/mu"], out[f"{tag}/std"] = mu.numpy(), std.numpy()
            out[f"{tag}/mask"] = np.packbits(mask.numpy().astype(np.uint8).ravel())
            out[f"{tag}/lik"] = lik[k].numpy()
            out[f"{tag}/y_hat"] = y_hat.numpy()
        out[f"{q_name}/pr"] = np.asarray(pr, np.float32)
        neg = float((torch.cat([s.flatten() for s in rec["std"]]) < 0).float().mean())
        print(q_name, "slices", n_prog, "std range", float(torch.cat([s.flatten() for s in rec['std']]).abs().max()), "neg frac", neg)
    # y_top / y_base are identical across q: store them once
    for q_name in ("pr5", "pr10"):
        for k in range(10):
            for nm in ("y_top", "y_base"):
                assert np.array_equal(out[f"{q_name}/slice{k}/{nm}"], out[f"pr2.5/slice{k}/{nm}"])
                del out[f"{q_name}/slice{k}/{nm}"]
    np.savez_compressed(OUT, **{k: (v.astype(np.float16).astype(np.float32) if False else v) for k, v in out.items()})
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
