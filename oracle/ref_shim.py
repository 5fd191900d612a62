"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference Python modules.

Used only in the build container (where /root/reference exists) by
``oracle/gen_golden.py`` to generate the golden vectors committed under
``tests/golden/`` and by ``oracle/gen_golden_model*.py`` / ``gen_golden_codec.py`` (model- and coder-derived vectors); ``tests/test_oracle_golden.py`` pins the C
restatement.  Nothing on the product path, in ``-m gpu`` tests or ``smoke()``
imports this file: /root/reference does not exist on the GPU box.  The one other
user is ``bench.py --impl reference``: when (and only when) ``PIC_REFERENCE_ROOT``
points at a visible reference tree -- the build container -- it also times the
reference's own torch code for the path through this loader (``torch_reference``
in its JSON line); on the GPU box that leg is skipped.

The reference modules on the hot path are

* ``src/layers/channel_mask.py``      (pure torch; loaded by file path because
  ``layers/__init__.py`` drags in timm / compressai),
* ``src/entropy_models/entropy_models.py`` (needs ``compressai`` -- un-vendored
  pip dependency ``compressai==1.2.4``, environment.yml:203 -- only for
  ``LowerBound``, the rANS coder handle and ``pmf_to_quantized_cdf``),
* ``src/models/utils.py`` (``ste_round``), ``src/models/pic.py:16-17``
  (``get_scale_table``; restated here because importing pic.py needs the whole
  conv stack).

``compressai`` is absent from this image, so a stub is installed in
``sys.modules``.  The only arithmetic it supplies is ``LowerBound`` whose
published CompressAI 1.2.4 definition is restated below (forward
``max(x, bound)``, backward pass-through iff ``x >= bound`` or ``grad < 0``).
That boundary is therefore "parity unpinned" (see DESIGN.md); everything else
is the reference's own code running on this box's torch.
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("PIC_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REFERENCE_ROOT, "src")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_SRC, "layers", "channel_mask.py"))


# --------------------------------------------------------------------------
# compressai stub (restatement of compressai.ops.LowerBound, v1.2.4)
# --------------------------------------------------------------------------
class _LowerBoundFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        pass_through_if = (x >= bound) | (grad_output < 0)
        return pass_through_if * grad_output, None


class LowerBound(nn.Module):
    """compressai.ops.LowerBound: ``bound`` is an f32 1-element buffer."""

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFunction.apply(x, self.bound)


def _install_compressai_stub() -> None:
    if "compressai" in sys.modules and not getattr(sys.modules["compressai"], "_pic_stub", False):
        return  # a real compressai is present; use it
    pkg = types.ModuleType("compressai")
    pkg._pic_stub = True
    pkg.__path__ = []
    pkg.available_entropy_coders = lambda: ["ans"]
    pkg.get_entropy_coder = lambda: "ans"

    ans = types.ModuleType("compressai.ans")

    class _NoCoder:
        def __init__(self, *a, **k):
            pass

        def _na(self, *a, **k):
            raise RuntimeError("rANS coder (compressai C++ ext) is not available in this image")

        encode_with_indexes = decode_with_indexes = _na

    ans.RansEncoder = _NoCoder
    ans.RansDecoder = _NoCoder
    ans.BufferedRansEncoder = _NoCoder

    cxx = types.ModuleType("compressai._CXX")

    def _pmf_to_quantized_cdf(*a, **k):
        raise RuntimeError("compressai._CXX.pmf_to_quantized_cdf is not available in this image")

    cxx.pmf_to_quantized_cdf = _pmf_to_quantized_cdf

    ops = types.ModuleType("compressai.ops")
    ops.LowerBound = LowerBound

    pkg.ans, pkg._CXX, pkg.ops = ans, cxx, ops
    sys.modules.update({"compressai": pkg, "compressai.ans": ans,
                        "compressai._CXX": cxx, "compressai.ops": ops})


def _load_by_path(name: str, relpath: str):
    path = os.path.join(_SRC, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cache: dict = {}


def load_reference():
    """Returns a namespace with the reference's hot-path classes/functions."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_compressai_stub()
    cm = _load_by_path("_pic_ref_channel_mask", "layers/channel_mask.py")
    em = _load_by_path("_pic_ref_entropy_models", "entropy_models/entropy_models.py")
    mu = _load_by_path("_pic_ref_models_utils", "models/utils.py")
    ns = types.SimpleNamespace(
        ChannelMask=cm.ChannelMask,
        ste_round=cm.ste_round,
        ste_round_models=mu.ste_round,
        GaussianConditional=em.GaussianConditional,
        EntropyModel=em.EntropyModel,
        LowerBound=LowerBound,
        channel_mask_module=cm,
        entropy_models_module=em,
    )
    _cache["ns"] = ns
    return ns


def get_scale_table(min: float = 0.11, max: float = 256, levels: int = 64):
    """Restates models/pic.py:12-17 (SCALES_MIN/MAX/LEVELS, get_scale_table)."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def make_gaussian_conditional(ref=None):
    """GaussianConditional(None) + scale table installed without update()
    (update() needs the C++ pmf_to_quantized_cdf; models/pic.py:97,230-236)."""
    ref = ref or load_reference()
    gc = ref.GaussianConditional(None)
    gc.scale_table = get_scale_table()
    return gc


# --------------------------------------------------------------------------
# The per-slice composition exactly as written in the model loops.
# --------------------------------------------------------------------------
def reference_slice_forward(ref, gc, masking, y_top, y_base, mu, scale, pr, training=False,
                            noise_seed=None, with_indexes=True):
    """models/pic.py:583-584, 621-629 (single-q) / 401-402, 430-443 (multi-q) and
    809-820 (compress): one progressive slice, composed from reference calls."""
    y_slice = y_top - y_base if y_base is not None else y_top          # pic.py:583-584
    block_mask = masking(scale, pr=pr, mask_pol="point-based-std")       # pic.py:621
    block_mask = masking.apply_noise(block_mask, False)                  # pic.py:622
    y_slice_m = y_slice - mu                                             # pic.py:625
    y_slice_m = y_slice_m * block_mask                                   # pic.py:626
    if noise_seed is not None:
        torch.manual_seed(noise_seed)
    outputs, lik = gc(y_slice_m, scale * block_mask, training=training)  # pic.py:628
    y_hat = ref.ste_round(y_slice - mu) * block_mask + mu                # pic.py:629
    out = dict(mask=block_mask, outputs=outputs, lik=lik, y_hat=y_hat)
    if with_indexes:
        out["idx"] = gc.build_indexes(scale * block_mask).int()          # pic.py:813
        out["symbols"] = gc.quantize(y_slice_m, "symbols")               # pic.py:819
    return out
