// pic_math.cuh -- per-element device arithmetic of the latent hot path (sm_100a).
//
// Every helper follows the f32 operation order of the reference's eager torch ops so that
// masks / indexes / symbols / reconstructions are bit-exact and the erfc arguments are
// bit-identical to the reference's (reference paths relative to src/):
//   layers/channel_mask.py:138-149            quantile threshold + `>=` mask
//   entropy_models/entropy_models.py:127-153  quantize
//   entropy_models/entropy_models.py:620-635  _likelihood  (+ 649-650 likelihood bound)
//   entropy_models/entropy_models.py:654-659  build_indexes
// Explicit __f*_rn intrinsics are used wherever an implicit FMA contraction would change a
// rounding that the reference performs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pic {

// float(-(2 ** -0.5)) converted to f32 by torch when it multiplies an f32 tensor
// (entropy_models.py:665-668).
__device__ constexpr float kNegInvSqrt2 = -0.70710678118654752440f;
__device__ constexpr float kInvSqrt2Pi = 0.39894228040143267794f;

enum UnitMode : int { kModeZeros = 0, kModeOnes = 1, kModeThreshold = 2, kModeDone = 3 };

__host__ __device__ __forceinline__ int unit_mode(float q01) {
    // sentinels of include/pic_latent.h: q01 < 0 -> ones (pr >= 10), q01 > 1 -> zeros (pr == 0)
    return (q01 < 0.0f) ? kModeOnes : (q01 > 1.0f) ? kModeZeros : kModeThreshold;
}

// Order-preserving map f32 -> u32 (ascending).  -0.0 is canonicalised to +0.0 so that it
// ties with +0.0 exactly as in torch.sort's float comparison.
__device__ __forceinline__ uint32_t float_to_key(float x) {
    uint32_t b = __float_as_uint(x);
    b = (b == 0x80000000u) ? 0u : b;
    const uint32_t m = static_cast<uint32_t>(static_cast<int32_t>(b) >> 31) | 0x80000000u;
    return b ^ m;
}

__device__ __forceinline__ float key_to_float(uint32_t k) {
    const uint32_t m = ((k >> 31) - 1u) | 0x80000000u;
    return __uint_as_float(k ^ m);
}

// torch.quantile's rank arithmetic (ATen quantile_compute): rank = f32(q) * f32(n-1).
__device__ __forceinline__ void quantile_ranks(float q01, int64_t n, uint32_t &lo, uint32_t &hi,
                                               float &w) {
    const float rank = __fmul_rn(q01, static_cast<float>(n - 1));
    const float fl = truncf(rank);
    lo = static_cast<uint32_t>(fl);
    hi = static_cast<uint32_t>(ceilf(rank));
    w = __fsub_rn(rank, fl);
}

// ATen lerp(a, b, w) with the FMA contraction its CPU (AVX2/AVX-512) and CUDA kernels use.
__device__ __forceinline__ float quantile_lerp(float a, float b, float w) {
    const float diff = __fsub_rn(b, a);
    return (fabsf(w) < 0.5f) ? __fmaf_rn(w, diff, a) : __fmaf_rn(-diff, __fsub_rn(1.0f, w), b);
}

// compressai.ops.LowerBound forward: torch.max(x, bound) (NaN propagates).
__device__ __forceinline__ float lower_bound(float x, float bound) {
    return !(x < bound) ? x : bound;
}

// half * erfc(const * t): GaussianConditional._standardized_cumulative.
__device__ __forceinline__ float std_cumulative(float t) {
    return 0.5f * erfcf(__fmul_rn(kNegInvSqrt2, t));
}

// _likelihood on a mean-removed value, then the likelihood lower bound.
__device__ __forceinline__ float likelihood(float value, float scale, float scale_bound,
                                            float lik_bound, float *raw = nullptr) {
    const float sc = lower_bound(scale, scale_bound);
    const float v = fabsf(value);
    const float upper = std_cumulative(__fdiv_rn(__fsub_rn(0.5f, v), sc));
    const float lower = std_cumulative(__fdiv_rn(__fsub_rn(-0.5f, v), sc));
    const float lik = __fsub_rn(upper, lower);
    if (raw) *raw = lik;
    return (lik_bound > 0.0f) ? lower_bound(lik, lik_bound) : lik;
}

// build_indexes: (len-1) - sum_{t in table[:-1]} (s <= t) on a sorted table
// == first j in [0, len-1) with s <= table[j], or len-1 (also for NaN).
__device__ __forceinline__ int scale_index(float scale, float scale_bound, const float *table,
                                           int table_len) {
    const float sc = lower_bound(scale, scale_bound);
    int lo = 0, hi = table_len - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sc <= table[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// 64-entry table specialisation: 6 fixed halving steps, no divergence.
__device__ __forceinline__ int scale_index64(float scale, float scale_bound, const float *table) {
    const float sc = lower_bound(scale, scale_bound);
    // search first j in [0,63) with sc <= table[j]; positions 0..63 -> 6-bit answer
    int pos = 0;
#pragma unroll
    for (int step = 32; step >= 1; step >>= 1) {
        const int probe = pos + step - 1;  // candidate: all entries <= probe are < sc ?
        // probe ranges over [0,62]; entry 63 (table[-1]) is never compared (table[:-1])
        const bool less = (probe < 63) && !(sc <= table[probe]);
        pos = less ? pos + step : pos;
    }
    return pos;
}

struct Gauss {
    float a, b;      // (0.5 - v)/sc, (-0.5 - v)/sc
    float sc;        // lower-bounded scale
    float raw;       // upper - lower
};

// d lik / d v and d lik / d sc with lik = Phi(a) - Phi(b):
//   phi(t) = exp(-t^2/2)/sqrt(2 pi);  dlik/dv = (phi(b) - phi(a))/sc;  dlik/dsc = (b phi(b) - a phi(a))/sc
__device__ __forceinline__ void likelihood_grads(float value, float scale, float scale_bound,
                                                 float &raw, float &dlik_dv, float &dlik_dsc) {
    const float sc = lower_bound(scale, scale_bound);
    const float v = fabsf(value);
    const float a = __fdiv_rn(__fsub_rn(0.5f, v), sc);
    const float b = __fdiv_rn(__fsub_rn(-0.5f, v), sc);
    const float ca = __fmul_rn(kNegInvSqrt2, a), cb = __fmul_rn(kNegInvSqrt2, b);
    raw = __fsub_rn(0.5f * erfcf(ca), 0.5f * erfcf(cb));
    const float pa = kInvSqrt2Pi * expf(-ca * ca);
    const float pb = kInvSqrt2Pi * expf(-cb * cb);
    const float inv = 1.0f / sc;
    dlik_dv = (pb - pa) * inv;
    dlik_dsc = (b * pb - a * pa) * inv;
}

}  // namespace pic
