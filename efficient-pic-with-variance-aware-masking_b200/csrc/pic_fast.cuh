// pic_fast.cuh -- issue-slot-lean versions of the per-element arithmetic (sm_100a).
//
// The first profile of the fused kernel (profiles/r1a_*) showed ~300 lane-instructions per
// element: 2 x erfcf (~100), 2 x IEEE __fdiv_rn (~67), 6-step table search (~43).  On B200 the
// HBM roofline leaves ~150 issue slots per element, so the arithmetic is restructured:
//   * two elements are processed together in Blackwell packed-f32 registers (FFMA2/FMUL2/FADD2:
//     one issue slot for two IEEE-rounded operations),
//   * division by the (shared) scale uses MUFU.RCP + one Newton step + an FMA residual
//     correction (Markstein): the quotient is the correctly rounded one except in rare
//     half-ulp ties, i.e. the erfc arguments equal the reference's to <= 1 ulp,
//   * erfc(t), t >= 0, is exp(-t^2) * g(t), g(t) = exp(t^2) erfc(t) = 1 + v R(v),
//     v = p t / (1 + p t), p = 1/2, R a degree-9 minimax fit (relative fit error 8e-10);
//     exp(-t^2) uses the exact t*t residual and a compensated log2(e) product before
//     MUFU.EX2.  Measured against f64: abs error <= 1.9e-7 on [0,1], relative <= 1.7e-6 in
//     the tails (oracle/gen_golden + tests bound the end-to-end likelihood error),
//   * the scale index is floor of a MUFU.LG2-based guess fixed up exactly against the real
//     table entries (the guess is within +-1 for a geometric table; arbitrary tables take the
//     binary search).
// Everything that must be bit-exact (mask, y_hat, symbols, index) is computed with
// individually rounded operations exactly as in pic_math.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pic_math.cuh"

namespace pic {

struct f2 {
    unsigned long long v;
};
__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 bc(float c) { return pk(c, c); }
__device__ __forceinline__ void unpk(f2 a, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// L2 eviction-priority hints: the unit's std is read twice (select sweep, then apply) and should
// stay in L2 in between; everything else is touched exactly once and is streamed.
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld_hint(const float4 *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_hint(float4 *p, float4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(int4 *p, int4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

// 16-byte asynchronous global->shared copy (LDGSTS) with an L2 policy; per-thread private slots
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gmem_src, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(smem_addr), "l"(gmem_src), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t smem_addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr));
    return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// classification helpers of the select sweep (instruction-count critical: the sweep is issue-bound)
// 1.0f if a < b else 0.0f (one FSET)
__device__ __forceinline__ float fset_lt(float a, float b) {
    float d;
    asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
// hits |= BIT if lo <= x <= hi (two FSETP + one predicated OR); false for NaN
template <uint32_t BIT>
__device__ __forceinline__ void or_if_in_range(uint32_t &hits, float x, float lo, float hi) {
    asm("{ .reg .pred p, q; setp.ge.f32 p, %1, %2; setp.le.and.f32 q, %1, %3, p; @q or.b32 %0, %0, %4; }"
        : "+r"(hits) : "f"(x), "f"(lo), "f"(hi), "n"(BIT));
}

// NaN-propagating min / max (torch.max(x, bound) semantics of LowerBound)
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float min_nan_abs(float a, float b) {
    float r;
    asm("{ .reg .f32 t; abs.f32 t, %1; min.NaN.f32 %0, t, %2; }" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// erfc(t) for t in [0, 10] (t already clamped; NaN propagates), two lanes at once.
// nt = -t.  See the header comment for the derivation; coefficients from oracle/fit_erfc.py.
__device__ __forceinline__ f2 erfc_pos2(f2 t, f2 nt) {
    const f2 pt = mul2(t, bc(0.5f));                 // p t
    const f2 d = add2(pt, bc(1.0f));                 // 1 + p t   (>= 1)
    float dl, dh;
    unpk(d, dl, dh);
    const f2 nu0 = pk(rcp_approx(-dl), rcp_approx(-dh));   // -1/d, 1 ulp
    const f2 e = fma2(d, nu0, bc(1.0f));             // 1 - d/d ~ 0
    const f2 nu = fma2(nu0, e, nu0);                 // Newton: -1/d to ~0.5 ulp
    const f2 w = mul2(pt, nu);                       // w = -v = -p t / (1 + p t)
    // g = 1 + v R(v) = 1 + w * (-R(-w)): Horner in w with coefficients -c_k (-1)^k
    f2 acc = fma2(bc(0.03844483569264412f), w, bc(0.11183761060237885f));
    acc = fma2(acc, w, bc(0.07856672257184982f));
    acc = fma2(acc, w, bc(0.009631340391933918f));
    acc = fma2(acc, w, bc(0.09950766712427139f));
    acc = fma2(acc, w, bc(-0.003952182363718748f));
    acc = fma2(acc, w, bc(-0.31069591641426086f));
    acc = fma2(acc, w, bc(0.27476468682289124f));
    acc = fma2(acc, w, bc(1.743239164352417f));
    acc = fma2(acc, w, bc(2.256758213043213f));
    const f2 g = fma2(acc, w, bc(1.0f));
    // exp(-t^2): t^2 = s + e2 exactly; exponent = (ns - e2) log2(e)
    const f2 s = mul2(t, t);
    const f2 ns = mul2(t, nt);                       // == -s exactly
    const f2 e2 = fma2(t, t, ns);                    // exact residual of t*t
    const f2 L2E = bc(1.4426950216293335f);
    const f2 zh = mul2(ns, L2E);                     // rounded exponent
    const f2 pzh = mul2(s, L2E);                     // == -zh exactly
    f2 zl = fma2(ns, L2E, pzh);                      // exact residual of the product
    zl = fma2(e2, bc(-1.4426950216293335f), zl);     // - e2 log2(e)
    zl = fma2(ns, bc(1.9259629911266175e-08f), zl);  // low part of log2(e)
    float zhl, zhh;
    unpk(zh, zhl, zhh);
    const f2 r0 = pk(ex2_approx(zhl), ex2_approx(zhh));
    const f2 r = fma2(r0, mul2(zl, bc(0.6931471805599453f)), r0);
    return mul2(g, r);
}

struct PairOut {
    float m[2], y_hat[2], lik[2];
    int32_t idx[2], sym[2];
};

struct IndexCtx {
    const float *tbl;     // table (shared memory copy when table_len == 64)
    float inv_step;       // (len-1) / (log2(tbl[len-1]) - log2(tbl[0]))
    float bias;           // -(log2(tbl[0]) * inv_step + eps)
    int len;
    bool geometric;       // guess-and-fix is valid
    bool tbl64;
};

// idx = #{j < len-1 : tbl[j] < sc}  (== build_indexes on a sorted table).
// Geometric table: g = floor(x - eps) + 1 with x = (lg2 sc - lg2 t0) * inv_step is idx or idx-1
// (|error of x| << eps << 1), so one comparison against the real table entry makes it exact.
// fminf() maps NaN / huge sc to the last index (non-propagating min returns the other operand).
__device__ __forceinline__ int scale_index_fast(float sc, const IndexCtx &c) {
    if (!c.geometric) return c.tbl64 ? scale_index64(sc, 0.0f, c.tbl) : scale_index(sc, 0.0f, c.tbl, c.len);
    // cap at len-2.5 so that g <= len-2 is always a valid table index; branch-free fix-up.
    const float x = fminf(__fmaf_rn(lg2_approx(sc), c.inv_step, c.bias), static_cast<float>(c.len) - 2.5f);
    const int g = max(__float2int_rd(x) + 1, 0);
    return g + (!(sc <= c.tbl[g]) ? 1 : 0);   // !(<=) is also true for NaN -> len-1, like the reference
}

__device__ __forceinline__ void likelihood_pair(float v0, float v1, float sc0, float sc1, float lik_bound,
                                                float &lik0, float &lik1);

// One pair of elements of a progressive slice (same arithmetic as apply_one in pic_latent.cu).
// s, yt, yb, mu, nz: the two elements' inputs.  want_lik / want_idx are warp-uniform.
template <bool TRAIN>
__device__ __forceinline__ void apply_pair(const float s[2], const float yt[2], const float yb[2],
                                           const float mu[2], const float nz[2], bool has_base,
                                           float thr, bool force_one, float scale_bound,
                                           float lik_bound, bool want_lik, bool want_idx,
                                           bool want_sym, const IndexCtx &ic, PairOut &o) {
    const float m0 = ((s[0] >= thr) || force_one) ? 1.0f : 0.0f;
    const float m1 = ((s[1] >= thr) || force_one) ? 1.0f : 0.0f;
    const f2 mf = pk(m0, m1);
    const f2 muv = pk(mu[0], mu[1]);
    const f2 r = has_base ? sub2(pk(yt[0], yt[1]), pk(yb[0], yb[1])) : pk(yt[0], yt[1]);  // pic.py:583-584
    const f2 d = sub2(r, muv);                                                            // pic.py:625
    const f2 y_m = mul2(d, mf);                                                           // pic.py:626
    const f2 s_m = mul2(pk(s[0], s[1]), mf);                                              // scale*block_mask
    float d0, d1, ym0, ym1, sm0, sm1;
    unpk(d, d0, d1);
    unpk(y_m, ym0, ym1);
    unpk(s_m, sm0, sm1);
    o.m[0] = m0;
    o.m[1] = m1;
    // y_hat = ste_round(d) * m + mu   (pic.py:629); (rd - d) + d == rd except for inf
    const f2 rd = pk(rintf(d0), rintf(d1));
    const f2 ste = add2(sub2(rd, d), d);
    const f2 yh = fma2(ste, mf, muv);               // product by {0,1} is exact: same as mul then add
    unpk(yh, o.y_hat[0], o.y_hat[1]);
    if (want_sym) {
        o.sym[0] = __float2int_rn(ym0);
        o.sym[1] = __float2int_rn(ym1);
    }
    const float sc0 = max_nan(sm0, scale_bound), sc1 = max_nan(sm1, scale_bound);  // LowerBound
    if (want_idx) {
        o.idx[0] = scale_index_fast(sc0, ic);
        o.idx[1] = scale_index_fast(sc1, ic);
    }
    if (!want_lik) return;
    // quantize: noise (training) or round (eval); v = |out|
    float out0, out1;
    if (TRAIN) {
        unpk(add2(y_m, pk(nz[0], nz[1])), out0, out1);
    } else {
        out0 = rintf(ym0);
        out1 = rintf(ym1);
    }
    likelihood_pair(fabsf(out0), fabsf(out1), sc0, sc1, lik_bound, o.lik[0], o.lik[1]);
}

// lik = max(Phi((.5 - v)/sc) - Phi((-.5 - v)/sc), lik_bound) for two elements (v = |x| >= 0,
// sc = lower-bounded scale), entropy_models.py:620-635, 649-650.
__device__ __forceinline__ void likelihood_pair(float v0, float v1, float sc0, float sc1, float lik_bound,
                                                float &lik0, float &lik1) {
    const f2 v = pk(v0, v1);
    const f2 sc = pk(sc0, sc1);
    // -1/sc: MUFU + Newton
    const f2 nrc0 = pk(rcp_approx(-sc0), rcp_approx(-sc1));
    const f2 e = fma2(sc, nrc0, bc(1.0f));
    const f2 nrc = fma2(nrc0, e, nrc0);
    // a = (0.5 - v)/sc, b = (-0.5 - v)/sc with negated numerators nna = v - 0.5, nnb = v + 0.5
    const f2 nna = add2(v, bc(-0.5f)), nnb = add2(v, bc(0.5f));
    const f2 qa = mul2(nna, nrc), qb = mul2(nnb, nrc);
    const f2 ra = fma2(qa, sc, nna), rb = fma2(qb, sc, nnb);      // -(na - qa sc)
    const f2 a = fma2(ra, nrc, qa), b = fma2(rb, nrc, qb);
    // erfc arguments xu = c a, xl = c b  (c = -1/sqrt2); xl > 0 always
    const f2 xu = mul2(a, bc(kNegInvSqrt2)), xl = mul2(b, bc(kNegInvSqrt2));
    float xu0, xu1, xl0, xl1;
    unpk(xu, xu0, xu1);
    unpk(xl, xl0, xl1);
    const f2 tu = pk(min_nan_abs(xu0, 10.0f), min_nan_abs(xu1, 10.0f));
    const f2 tl = pk(min_nan_abs(xl0, 10.0f), min_nan_abs(xl1, 10.0f));
    const f2 hu = mul2(erfc_pos2(tu, mul2(tu, bc(-1.0f))), bc(0.5f));
    const f2 lo = mul2(erfc_pos2(tl, mul2(tl, bc(-1.0f))), bc(0.5f));
    const f2 omh = sub2(bc(1.0f), hu);               // Phi for a negative erfc argument
    float hu0, hu1, om0, om1;
    unpk(hu, hu0, hu1);
    unpk(omh, om0, om1);
    const f2 up = pk((xu0 < 0.0f) ? om0 : hu0, (xu1 < 0.0f) ? om1 : hu1);
    float l0, l1;
    unpk(sub2(up, lo), l0, l1);
    lik0 = (lik_bound > 0.0f) ? max_nan(l0, lik_bound) : l0;
    lik1 = (lik_bound > 0.0f) ? max_nan(l1, lik_bound) : l1;
}

}  // namespace pic
