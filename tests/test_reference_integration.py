"""INTEGRATION.md section 1 ("change the import lines") checked against the reference model itself.  Needs the reference
tree (oracle/ref_shim.py knows where: PIC_REFERENCE_ROOT or its default): it exists in the build container only, so this
is a CPU test that skips elsewhere -- the GPU box never sees the reference."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402  (test infrastructure: locates the reference tree, imports nothing from it)

REF = ref_shim.REFERENCE_ROOT


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "src", "models", "pic.py")), reason="reference tree not present")
def test_reference_model_builds_and_loads_with_the_dropins():
    """VarianceMaskingPIC with ChannelMask / GaussianConditional / EntropyBottleneck swapped for the drop-ins: identical
    state_dict keys, shapes and initial values, strict load of the reference's weights, the reference's aux-loss
    plumbing finds the drop-in bottleneck, and CPU tensors are refused (no silent fallback)."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "reference_model_with_dropins.py")],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0 and "REFERENCE_MODEL_WITH_DROPINS_OK" in res.stdout, res.stdout[-3000:]
