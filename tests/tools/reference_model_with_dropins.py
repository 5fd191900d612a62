"""Worker of tests/test_reference_integration.py (build container only: needs the reference tree).
Builds the UNMODIFIED reference model twice -- once as shipped, once with the three classes of INTEGRATION.md section 1
swapped for the CUDA drop-ins (the import lines of models/pic.py:4,9 and models/base.py) -- and checks that the swap is
structurally invisible: same modules, same state_dict keys and shapes, the reference's weights load with strict=True,
buffers the codec needs survive, and the drop-ins refuse CPU tensors loudly instead of computing something else."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import gen_golden_model as base          # stubs for compressai / timm (absent third-party packages)
    base.install_stubs()
    import models.base as mbase
    import models.pic as mpic
    import pic_b200

    torch.manual_seed(0)
    ref_net = mpic.VarianceMaskingPIC()
    ref_sd = ref_net.state_dict()
    ref_aux = ref_net.aux_loss()
    swapped = {"ChannelMask": pic_b200.ChannelMask, "GaussianConditional": pic_b200.GaussianConditional,
               "EntropyBottleneck": pic_b200.EntropyBottleneck}
    for mod in (mpic, mbase):
        for name, cls in swapped.items():
            if hasattr(mod, name):
                setattr(mod, name, cls)
    mpic.ste_round = pic_b200.channel_mask.ste_round
    torch.manual_seed(0)
    net = mpic.VarianceMaskingPIC()
    assert isinstance(net.masking, pic_b200.ChannelMask)
    assert isinstance(net.gaussian_conditional, pic_b200.GaussianConditional)
    assert isinstance(net.entropy_bottleneck, pic_b200.EntropyBottleneck)
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys()), sorted(set(sd) ^ set(ref_sd))[:10]
    for k in sd:
        assert sd[k].shape == ref_sd[k].shape and sd[k].dtype == ref_sd[k].dtype, k
    net.load_state_dict(ref_sd, strict=True)      # the model's own override (models/pic.py:240-247): raises on any mismatch
    for k, v in net.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    # same random initialisation of the factorised prior (same constructor arithmetic, same RNG consumption)
    for k in sd:
        if k.startswith("entropy_bottleneck."):
            assert torch.equal(sd[k], ref_sd[k]), k
    # the reference's own aux-loss / update plumbing finds the drop-in (models/base.py iterates over EntropyBottleneck)
    aux = net.aux_loss()
    assert torch.isfinite(aux) and torch.allclose(aux, ref_aux), (aux, ref_aux)
    # no silent CPU path
    for call in (lambda: net.masking(torch.rand(1, 32, 4, 4), pr=5.0),
                 lambda: net.gaussian_conditional(torch.rand(1, 32, 4, 4), torch.rand(1, 32, 4, 4)),
                 lambda: net.entropy_bottleneck(torch.rand(1, 192, 4, 4))):
        try:
            call()
        except RuntimeError as e:
            assert "CUDA" in str(e), e
        else:
            raise AssertionError("a drop-in computed on CPU tensors")
    print("REFERENCE_MODEL_WITH_DROPINS_OK", len(sd), "state_dict entries")


if __name__ == "__main__":
    main()
