"""ctypes binding of libpic_latent.so (the C ABI declared in include/pic_latent.h).

The library is built in-tree (``build()`` below or ``__graft_entry__.build()``) with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  There is no CPU fallback: if the shared
object is missing, or a tensor is not a CUDA tensor, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("PIC_LIB_PATH") or os.path.join(_PKG, "libpic_latent.so")   # override: kernel experiments
_SOURCES = [os.path.join(_PKG, "csrc", f) for f in ("pic_latent.cu", "pic_tma_select.cu", "pic_rank.cu", "pic_bottleneck.cu", "pic_host.cu", "pic_rans.cpp")]
_HEADERS = [os.path.join(_PKG, "csrc", f) for f in ("pic_math.cuh", "pic_fast.cuh", "pic_select.cuh", "pic_gselect.cuh", "pic_params.h")] + [
    os.path.join(_ROOT, "include", "pic_latent.h"), os.path.join(_ROOT, "include", "pic_codec.h")]

PIC_OK = 0
PIC_ERR_INVALID_ARGUMENT = -1
PIC_ERR_TOO_LARGE = -2
PIC_ERR_WORKSPACE = -3
PIC_ERR_CUDA = -4
PIC_ERR_UNALIGNED = -5

Q_ONES = -1.0
Q_ZEROS = 2.0

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "--threads", "0", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-ldl"]

_vp, _i64, _i32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t
_tables = [_vp, _i32, _i32, _vp, _vp]   # cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets (include/pic_codec.h)

# name -> (restype, argtypes); must list every symbol of include/pic_latent.h
SIGNATURES = {
    "pic_version": (C.c_int, []),
    "pic_error_string": (C.c_char_p, [C.c_int]),
    "pic_last_cuda_error": (C.c_int, []),
    "pic_fused_max_elems": (_i64, []),
    "pic_debug_select_counters": (C.c_int, [_vp, _vp]),
    "pic_slice_forward_plan": (C.c_int, [_i64, _i64, _i32, _vp]),
    "pic_workspace_bytes": (_sz, [_i64, _i64]),
    "pic_select_threshold": (C.c_int, [_vp, _i64, _i64, _f32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pic_select_threshold_multi": (C.c_int, [_vp, _i64, _i64, _vp, _i32, _vp, _vp]),
    "pic_level_map": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "pic_rank_order_workspace_bytes": (_sz, [_i64, _i64]),
    "pic_rank_order": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "pic_select_state_bytes": (_sz, [_i64]),
    "pic_hist_words": (_i64, []),
    "pic_select_begin": (C.c_int, [_vp, _i64, _i64, _f32, _vp, _vp]),
    "pic_hist_round": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp]),
    "pic_select_advance": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "pic_select_finish": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "pic_tiled_workspace_bytes": (_sz, [_i64]),
    "pic_dist_unique_id": (C.c_int, [_vp]),
    "pic_dist_comm_init": (C.c_int, [_vp, _i32, _i32, _vp]),
    "pic_dist_comm_destroy": (C.c_int, [_vp]),
    "pic_tiled_select_threshold": (C.c_int, [_vp, _i64, _i64, _i64, _f32, _vp, _vp, _vp, _sz, _vp, _vp]),
    "pic_tiled_sampled_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "pic_tiled_select_threshold_sampled": (C.c_int, [_vp, _i64, _i64, _i64, _f32, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "pic_dist_p2p_region_bytes": (C.c_int, [_i64, _i64, _i32, _vp, _vp]),
    "pic_dist_p2p_init": (C.c_int, [_vp, _i32, _sz, _sz, _vp]),
    "pic_dist_p2p_destroy": (C.c_int, [_vp]),
    "pic_tiled_select_threshold_p2p": (C.c_int, [_vp, _i64, _i64, _i64, _f32, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "pic_channel_mask": (C.c_int, [_vp, _i64, _i64, _f32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pic_attention_mask": (C.c_int, [_vp, _i64, _i64, _f32, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "pic_lrp_merge": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "pic_lrp_merge_backward": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "pic_rem_merge": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "pic_rem_merge_backward": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "pic_mask_from_threshold": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "pic_slice_forward": (C.c_int, [_vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _i32, _f32, _f32,
                                    _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pic_slice_forward_multi": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _f32, _f32, _i64, _i64,
                                          _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pic_slice_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _i64,
                                     _vp, _vp, _vp, _vp, _vp]),
    "pic_gaussian_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _f32, _f32, _vp, _vp, _vp]),
    "pic_gaussian_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _f32, _f32,
                                        _vp, _vp, _vp, _vp]),
    "pic_build_indexes": (C.c_int, [_vp, _i64, _vp, _i32, _f32, _vp, _vp]),
    "pic_quantize": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "pic_dequantize": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "pic_log_sum": (C.c_int, [_vp, _i64, _i64, _vp, _vp]),
    "pic_bottleneck_params_per_channel": (C.c_int, [_i32, _i32, _i32, _i32]),
    "pic_bottleneck_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _f32, _vp, _vp, _vp]),
    "pic_bottleneck_backward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _f32,
                                          _vp, _vp, _vp, _vp, _vp, _vp]),
    # include/pic_codec.h (host code: CDF tables + rANS)
    "pic_pmf_to_quantized_cdf": (C.c_int, [_vp, _i32, _i32, _vp]),
    "pic_rans_stream_bound": (_i64, [_i64]),
    "pic_rans_encode_with_indexes": (_i64, [_vp, _vp, _i64] + _tables + [_vp, _i64]),
    "pic_rans_decode_with_indexes": (C.c_int, [_vp, _i64, _vp, _i64] + _tables + [_vp]),
    "pic_rans_encode_batch": (C.c_int, [_vp, _vp, _i64, _i64] + _tables + [_vp, _i64, _vp, _i32]),
    "pic_rans_decode_batch": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64] + _tables + [_vp, _i32]),
    "pic_rans_encode_levels": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _i32] + _tables + [_vp, _i64, _vp, _i32]),
    "pic_rans_decode_levels": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32] + _tables + [_vp, _i32]),
    "pic_host_pipeline_bytes": (_sz, [_i64, _i64]),
    "pic_slice_forward_host_compact": (C.c_int, [_vp, _vp, _vp, _vp, _f32, _vp, _vp, _i32, _f32, _f32, _i64, _i64, _i64,
                                                 _vp, _vp, _vp, _vp, _vp, _sz]),
    "pic_slice_forward_host": (C.c_int, [_vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _i32, _f32, _f32,
                                         _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz]),
}

_lib = None


class PicCudaError(RuntimeError):
    pass


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libpic_latent.so")


STAMP_PATH = LIB_PATH + ".srchash"   # written next to the library by build(); travels with it, not committed


def source_hash() -> str:
    """SHA-256 over the compile flags and the bytes of every source / header of the library (sorted by name):
    identifies the sources a binary was built from, independent of file times and of who built it."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(_SOURCES + _HEADERS):
        if os.path.isfile(p):
            h.update(os.path.basename(p).encode())
            with open(p, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def built_from() -> str:
    """The source hash recorded when the library on disk was built ('' if unknown)."""
    try:
        with open(STAMP_PATH) as f:
            return f.read().strip()
    except OSError:
        return ""


def needs_build() -> bool:
    """True when the library is missing or was built from other sources than the ones on disk (hash, not mtime:
    a binary that is newer than the sources but foreign to them is rebuilt too)."""
    return not os.path.isfile(LIB_PATH) or built_from() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every CUDA source of the package for sm_100a into libpic_latent.so (in-tree)."""
    if not force and not needs_build():
        return LIB_PATH
    srcs = [s for s in _SOURCES if os.path.isfile(s)]
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)  # the image's CC points at a gcc without OpenMP specs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    with open(STAMP_PATH, "w") as f:
        f.write(source_hash() + "\n")
    if verbose:
        print(res.stdout)
    return LIB_PATH


def lib():
    """Loads the C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise PicCudaError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback for the PIC latent path.")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    """Maps C-ABI error codes to the exceptions the reference API raises."""
    if rc == PIC_OK:
        return
    msg = lib().pic_error_string(rc).decode()
    if rc == PIC_ERR_TOO_LARGE:
        raise RuntimeError("quantile() input tensor is too large")  # torch.quantile's message
    if rc == PIC_ERR_INVALID_ARGUMENT:
        raise ValueError(f"{what}: {msg}")
    if rc == PIC_ERR_CUDA:
        raise PicCudaError(f"{what}: CUDA error {lib().pic_last_cuda_error()}")
    raise RuntimeError(f"{what}: {msg} ({rc})")
