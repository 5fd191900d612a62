"""Shared tolerances / helpers for the parity tests.

Tolerances (BASELINE.json north_star: "within 1e-5 relative (FP32)"):

* masks, scale indexes, symbols, reconstructions y_hat, thresholds: BIT-EXACT.
* likelihoods: |got - ref| <= LIK_RTOL*|ref| + LIK_ATOL.  The absolute floor is needed because
  lik = upper - lower subtracts two values of magnitude <= 1 whose f32 erfc evaluations differ
  between libraries by a few ulp (ulp(0.5..1) = 6e-8); the reference's own f32 result differs
  from exact math by up to 3.8e-5 relative at large scales for the same reason.
* rate sums (sum log lik): relative RATE_RTOL.
* gradients: |got - ref| <= GRAD_RTOL*|ref| + GRAD_ATOL_REL*max|ref|.
"""
import os

import numpy as np

LIK_RTOL = 1e-5
LIK_ATOL = 3e-7
RATE_RTOL = 1e-5
GRAD_RTOL = 1e-4
GRAD_ATOL_REL = 2e-6

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PR_LIST = [0, 1e-4, 0.5, 0.75, 1, 2.5, 5, 7.3, 9.9999, 10, 11]


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def scale_table():
    return np.load(os.path.join(GOLDEN, "scale_table.npy"))


def unpack_mask(packed, shape):
    n = int(np.prod(shape))
    return np.unpackbits(packed)[:n].reshape(shape).astype(np.float32)


# every likelihood comparison of the session: (test id, what, elements, worst relative error, elements that needed
# the absolute floor, worst excess over the purely relative bound) -- printed by conftest.py at the end of the run so
# that the use of LIK_ATOL is visible, not just allowed
LIK_STATS = []


def assert_lik_close(got, ref, what="lik"):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref)
    tol = LIK_RTOL * np.abs(ref) + LIK_ATOL
    if err.size:
        rel = err / np.maximum(np.abs(ref), 1e-300)
        over = err > LIK_RTOL * np.abs(ref)
        LIK_STATS.append((os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0], what, int(err.size),
                          float(rel.max()), int(over.sum()), float((err - LIK_RTOL * np.abs(ref)).max())))
    bad = ~(err <= tol)
    assert not bad.any(), (f"{what}: {bad.sum()} of {bad.size} outside tolerance; worst abs {err.max():.3e}, "
                           f"worst excess {(err - tol).max():.3e}")


def assert_grad_close(got, ref, what="grad"):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref)
    tol = GRAD_RTOL * np.abs(ref) + GRAD_ATOL_REL * max(np.abs(ref).max(), 1e-30)
    bad = ~(err <= tol)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} outside tolerance; worst abs {err.max():.3e}"


def hashed_std(n: int, seed: int) -> np.ndarray:
    """Same exact-integer hash as oracle/gen_golden.py:hashed_std (regenerable large input)."""
    i = np.arange(n, dtype=np.uint64)
    k = (i * np.uint64(2654435761) + np.uint64(seed) * np.uint64(40503)) & np.uint64(0xFFFFFFFF)
    k ^= k >> np.uint64(15)
    k = (k * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    k ^= k >> np.uint64(13)
    return ((k >> np.uint64(8)).astype(np.float32) / np.float32(1 << 24)) - np.float32(0.5)


def trained_like(rng, shape):
    """SURVEY 8(d) 'S-trained-like' synthetic latents (numpy, host)."""
    std = np.exp(rng.normal(-1.0, 1.2, size=shape)).clip(1e-3, 300.0)
    flip = rng.random(size=shape) < 0.02
    std = np.where(flip, -std, std).astype(np.float32)
    mu = rng.normal(0, 1, size=shape).astype(np.float32)
    y_base = rng.normal(0, 2, size=shape).astype(np.float32)
    y_top = (y_base + mu + np.abs(std) * rng.normal(0, 1, size=shape)).astype(np.float32)
    return y_top, y_base, mu, std
