"""TEST INFRASTRUCTURE / provenance: derives the erfc approximation used by csrc/pic_fast.cuh.

    erfc(t) = exp(-t^2) * g(t),   g(t) = exp(t^2) erfc(t) = 1 + v R(v),   v = p t / (1 + p t),  p = 1/2

R is a degree-9 polynomial fitted (Lawson-reweighted least squares ~ minimax in relative error of g) on
t in [0, 6.5]; beyond t = 10 the kernel clamps (erfc < 1e-45).  The script prints the coefficients and the
error of an emulated-f32 evaluation (with 1-ulp reciprocal and 2-ulp exp2 error models) against f64 erfc.
Needs numpy + scipy.  Run: python oracle/fit_erfc.py
"""
import numpy as np
import scipy.special as sp

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def fit(p, N, tmax=6.5):
    vmax = 1 - 1 / (1 + p * tmax)
    k = np.arange(800)
    x = np.cos(np.pi * (k + 0.5) / 800)
    v = np.maximum((x + 1) / 2 * vmax, 1e-9)
    t = v / (p * (1 - v))
    g = sp.erfcx(t)
    y = (g - 1) / v
    w = np.ones_like(v)
    V = np.vander(v, N + 1, increasing=True)
    for _ in range(100):
        sc = (v / g) * w
        c, *_ = np.linalg.lstsq(V * sc[:, None], y * sc, rcond=None)
        r = np.abs((1 + v * (V @ c)) / g - 1)
        w = w * (r / r.max() + 1e-3) ** 0.5
        w /= w.max()
    return c, r.max()


def erfc_f32(t, p, c, rcp_ulp=0.0, ex2_ulp=0.0, rng=None):
    t = t.astype(f32)
    pt = (f32(p) * t).astype(f32)
    d = (pt + f32(1)).astype(f32)
    u = 1.0 / d.astype(np.float64)
    if rcp_ulp:
        u = u * (1 + rcp_ulp * rng.uniform(-1, 1, size=u.shape) * 6e-8)
    u = u.astype(f32)
    v = (pt * u).astype(f32)
    acc = np.full_like(t, f32(c[-1]))
    for ck in c[-2::-1]:
        acc = fma(acc, v, np.full_like(t, f32(ck)))
    g = fma(acc, v, np.ones_like(t))
    s = (t * t).astype(f32)
    e = fma(t, t, -s)
    L2Eh = f32(1.4426950408889634)
    L2El = f32(1.4426950408889634 - float(L2Eh))
    zh = (s * L2Eh).astype(f32)
    zl = fma(s, np.full_like(t, L2Eh), -zh)
    zl = fma(s, np.full_like(t, L2El), zl)
    zl = fma(e, np.full_like(t, L2Eh), zl)
    r = np.exp2(-zh.astype(np.float64))
    if ex2_ulp:
        r = r * (1 + ex2_ulp * rng.uniform(-1, 1, size=r.shape) * 6e-8)
    r = r.astype(f32)
    r = fma(r, (-zl * f32(0.6931471805599453)).astype(f32), r)
    return (g * r).astype(f32)


if __name__ == "__main__":
    p, N = 0.5, 9
    c, fit_err = fit(p, N)
    print("p =", p, "degree", N, "relative fit error of g:", fit_err)
    print("R coefficients c0..c9 (pic_fast.cuh uses -c_k (-1)^k in Horner form on w = -v):")
    print([float(f32(x)) for x in c])
    rng = np.random.default_rng(0)
    t = np.linspace(0, 6.5, 2000001)
    ref = sp.erfc(t.astype(f32).astype(np.float64))
    for rc, ex in ((0, 0), (1, 2)):
        got = erfc_f32(t, p, c, rc, ex, rng).astype(np.float64)
        rel, ab = np.abs(got / ref - 1), np.abs(got - ref)
        print(f"rcp err {rc} ulp, ex2 err {ex} ulp: max abs {ab.max():.3e} (t={t[ab.argmax()]:.3f}), "
              f"max rel {rel.max():.3e}, max rel t<2 {rel[t < 2].max():.3e}")
