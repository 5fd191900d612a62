// pic_gselect.cuh -- global sampled select: the sampled-pivot exact select split into three
// kernels so that the one pass over std runs tile by tile in GLOBAL address order (like
// slice_apply_kernel) instead of unit by unit (thousands of concurrent DRAM streams):
//   gs_pivot_kernel   grid = units          sample -> pivots (or the whole select for tiny units)
//   gs_sweep_kernel   grid = units x tiles  count below / NaN, append bracket keys to the unit's
//                                           candidate buffer in global memory
//   gs_finish_kernel  grid = units          exact ranks among the candidates -> thr (fallback:
//                                           full histogram select of the unit by this CTA)
// Included by pic_latent.cu after the shared helpers (quantile arithmetic, block selects).
#pragma once

namespace pic {

struct GsUnit {            // 32 bytes per unit in the workspace
    float plo_f, phi_f;    // bracket pivots (+-inf = open end)
    uint32_t c_below;      // #{x < plo}
    uint32_t c_cand;       // bracket elements appended (may exceed kCandMax => fallback)
    uint32_t nan_flag;
    uint32_t state;        // 0: sweep + finish pending, 1: finished by the pivot kernel
    uint32_t pad[2];
};
static_assert(sizeof(GsUnit) == 32, "GsUnit must stay 32 bytes");
constexpr int kGsTile = 8192;       // elements per CTA of gs_sweep_kernel
constexpr int kGsThreads = 256;

struct GsParams {
    const float *std;
    const float *q01_per_unit;
    float q01;
    int64_t n, units;
    GsUnit *st;
    uint32_t *cand;        // [units][cand_cap]
    float *thr, *a_out, *b_out;
    int vec;
    int64_t cand_cap;      // kCandMax for fused-size units; n/8 for large units (candidates as raw floats)
    uint32_t *below_tile;  // large units: [units][tiles] per-tile counts below the bracket (no atomics), else null
};

__global__ void __launch_bounds__(kGsThreads, 5) gs_pivot_kernel(const GsParams p) {
    constexpr int THREADS = kGsThreads;
    __shared__ __align__(16) uint32_t hist[2 * kHistBins];
    __shared__ __align__(16) uint32_t cand[kCandMax];
    __shared__ uint32_t scratch[kScratchWords];
    const int tid = threadIdx.x;
    const int64_t u = blockIdx.x;
    const int n = static_cast<int>(p.n);
    const float q = p.q01_per_unit ? p.q01_per_unit[u] : p.q01;
    const int mode = unit_mode(q);
    GsUnit st{};
    st.state = 1;
    if (mode != kModeThreshold) {
        const float t = (mode == kModeOnes) ? -INFINITY : INFINITY;
        if (tid == 0) {
            p.thr[u] = t;
            if (p.a_out) p.a_out[u] = t;
            if (p.b_out) p.b_out[u] = t;
            p.st[u] = st;
        }
        return;
    }
    const float *std_u = p.std + u * p.n;
    uint32_t lo, hi;
    float w;
    quantile_ranks(q, p.n, lo, hi, w);
    if (n <= kCandMax) {   // tiny unit: select here
        if (tid == 0) scratch[39] = 0u;
        __syncthreads();
        bool has_nan = false;
        for (int j = tid; j < n; j += THREADS) {
            const float v = __ldg(std_u + j);
            has_nan |= (v != v);
            cand[j] = float_to_key(v);
        }
        if (__any_sync(0xffffffffu, has_nan) && (tid & 31) == 0) scratch[39] = 1u;
        __syncthreads();
        uint32_t a_key, b_key;
        block_select_norm<THREADS>(cand, n, 0u, 32, hist, scratch, lo, hi, a_key, b_key);
        float a = key_to_float(a_key), b = key_to_float(b_key);
        float t = quantile_lerp(a, b, w);
        if (scratch[39] != 0u) t = a = b = __int_as_float(0x7fc00000);
        if (tid == 0) {
            p.thr[u] = t;
            if (p.a_out) p.a_out[u] = a;
            if (p.b_out) p.b_out[u] = b;
            p.st[u] = st;
        }
        return;
    }
    // sample (hashed stride) -> cand
    int S = n >> 3;
    S = S < 1024 ? 1024 : (S > kCandMax ? kCandMax : S);
    S &= ~3;
    if (p.vec) {
        const int nvec = n >> 2, S4 = S >> 2;
        const int stride = nvec / S4;
        const float4 *s4 = reinterpret_cast<const float4 *>(std_u);
        for (int i = tid; i < S4; i += THREADS) {
            const uint32_t jit = ((static_cast<uint32_t>(i) * 0x9E3779B1u) >> 12) % static_cast<uint32_t>(stride);
            const float4 v = __ldg(s4 + static_cast<size_t>(i) * stride + jit);
            reinterpret_cast<uint4 *>(cand)[i] =
                make_uint4(float_to_key(v.x), float_to_key(v.y), float_to_key(v.z), float_to_key(v.w));
        }
    } else {
        const int stride = n / S;
        for (int i = tid; i < S; i += THREADS) {
            const uint32_t jit = ((static_cast<uint32_t>(i) * 0x9E3779B1u) >> 12) % static_cast<uint32_t>(stride);
            cand[i] = float_to_key(__ldg(std_u + static_cast<size_t>(i) * stride + jit));
        }
    }
    __syncthreads();
    const float frac = static_cast<float>(lo) / static_cast<float>(n > 1 ? n - 1 : 1);
    const float kt = frac * static_cast<float>(S - 1);
    const float margin = 4.0f * sqrtf(static_cast<float>(S) * frac * (1.0f - frac)) + 4.0f;
    const int klo = static_cast<int>(floorf(kt - margin));
    const int khi = static_cast<int>(ceilf(kt + margin));
    uint32_t plo_key, phi_key;
    block_bracket_pair<THREADS>(cand, S, hist, scratch, klo > 0 ? klo : 0, khi < S - 1 ? khi : S - 1, plo_key, phi_key);
    st.plo_f = (klo > 0) ? key_to_float(plo_key) : -INFINITY;
    st.phi_f = (khi < S - 1) ? key_to_float(phi_key) : INFINITY;
    st.state = 0;
    if (tid == 0) p.st[u] = st;
}

// RAW: candidates are stored as float bits (input of the histogram rounds) instead of ordered keys.
// One CTA per 8192-element tile: the tile is parked in shared memory while it is classified (one hit bit per
// element in a register), then ONE global reservation per CTA places the tile's bracket elements in the unit's
// candidate buffer (per-warp reservations serialise on the unit's counter once units are millions of elements).
template <bool VEC, bool RAW>
__global__ void __launch_bounds__(kGsThreads, 4) gs_sweep_kernel(const GsParams p, int tiles_per_unit) {
    constexpr int THREADS = kGsThreads;
    constexpr int WARPS = THREADS / 32;
    static_assert(kGsTile == 32 * THREADS, "one hit bit per element and thread");
    __shared__ __align__(16) float4 park4[kGsTile / 4];   // 32 KB: the CTA's tile
    __shared__ uint32_t warp_off[WARPS + 1], warp_below[WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t u = blockIdx.x / tiles_per_unit;
    const int tile = blockIdx.x - static_cast<int>(u) * tiles_per_unit;
    GsUnit *st = p.st + u;
    if (st->state != 0u) return;
    const float plo_f = st->plo_f, phi_f = st->phi_f;
    const int64_t begin = static_cast<int64_t>(tile) * kGsTile;
    const int len = static_cast<int>(min(static_cast<int64_t>(kGsTile), p.n - begin));
    const float *base = p.std + u * p.n + begin;
    uint32_t *cand = p.cand + u * p.cand_cap;
    const uint32_t cap = static_cast<uint32_t>(p.cand_cap);
    const float *parkf = reinterpret_cast<const float *>(park4);
    uint32_t below = 0, hits = 0;
    bool has_nan = false;
    if (VEC) {   // bit 4*i + e  <->  element e of float4 (i * THREADS + tid)
        const int nvec = len >> 2;   // VEC: n % 4 == 0, so every tile is whole float4s
        const uint64_t pol_last = policy_evict_last();
        const float4 *s4 = reinterpret_cast<const float4 *>(base);
        f2 below2 = pk(0.0f, 0.0f);
        auto classify = [&](const float4 &q, int sh) {
            const float mx = max_nan(max_nan(q.x, q.y), max_nan(q.z, q.w));
            has_nan |= (mx != mx);
            below2 = add2(below2, pk(fset_lt(q.x, plo_f), fset_lt(q.y, plo_f)));   // exact: counts << 2^24
            below2 = add2(below2, pk(fset_lt(q.z, plo_f), fset_lt(q.w, plo_f)));
            uint32_t h4 = 0;
            or_if_in_range<1u>(h4, q.x, plo_f, phi_f);
            or_if_in_range<2u>(h4, q.y, plo_f, phi_f);
            or_if_in_range<4u>(h4, q.z, plo_f, phi_f);
            or_if_in_range<8u>(h4, q.w, plo_f, phi_f);
            hits |= h4 << sh;
        };
        if (nvec == kGsTile / 4) {
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ld_hint(s4 + tid + i * THREADS, pol_last);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                park4[i * THREADS + tid] = v[i];
                classify(v[i], 4 * i);
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                const int j = tid + i * THREADS;
                if (j < nvec) {
                    const float4 v = ld_hint(s4 + j, pol_last);
                    park4[j] = v;
                    classify(v, 4 * i);
                }
            }
        }
        float b_lo, b_hi;
        unpk(below2, b_lo, b_hi);
        below = static_cast<uint32_t>(b_lo + b_hi);
    } else {     // bit i  <->  element i * THREADS + tid
        float *parkw = reinterpret_cast<float *>(park4);
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int j = tid + i * THREADS;
            if (j < len) {
                const float x = __ldg(base + j);
                parkw[j] = x;
                has_nan |= (x != x);
                below += (x < plo_f) ? 1u : 0u;
                if (x >= plo_f && x <= phi_f) hits |= 1u << i;
            }
        }
    }
    // CTA-wide reservation: warp scan -> per-warp offsets -> one atomicAdd on the unit's counter
    const uint32_t cnt = static_cast<uint32_t>(__popc(hits));
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    below = __reduce_add_sync(0xffffffffu, below);
    if (lane == 31) {
        warp_off[warp] = incl;
        warp_below[warp] = below;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0, bsum = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = warp_off[w];
            warp_off[w] = run;
            run += c;
            bsum += warp_below[w];
        }
        warp_off[WARPS] = run ? atomicAdd(&st->c_cand, run) : 0u;
        if (p.below_tile) p.below_tile[blockIdx.x] = bsum;          // summed by gs_begin_rounds_kernel
        else if (bsum) atomicAdd(&st->c_below, bsum);
    }
    __syncthreads();
    uint32_t pos = warp_off[WARPS] + warp_off[warp] + incl - cnt;
    while (hits) {
        const int e = __ffs(hits) - 1;
        hits &= hits - 1u;
        const float x = VEC ? parkf[((e >> 2) * THREADS + tid) * 4 + (e & 3)] : parkf[e * THREADS + tid];
        if (pos < cap) cand[pos] = RAW ? __float_as_uint(x) : float_to_key(x);
        ++pos;
    }
    if (__any_sync(0xffffffffu, has_nan) && lane == 0) atomicOr(&st->nan_flag, 1u);
}

__global__ void __launch_bounds__(kGsThreads, 5) gs_finish_kernel(const GsParams p) {
    constexpr int THREADS = kGsThreads;
    __shared__ __align__(16) uint32_t hist[2 * kHistBins];
    __shared__ __align__(16) uint32_t cand[kCandMax];
    __shared__ uint32_t scratch[kScratchWords];
    const int tid = threadIdx.x;
    const int64_t u = blockIdx.x;
    const GsUnit st = p.st[u];
    if (st.state != 0u) return;
    const float q = p.q01_per_unit ? p.q01_per_unit[u] : p.q01;
    uint32_t lo, hi;
    float w;
    quantile_ranks(q, p.n, lo, hi, w);
    uint32_t a_key, b_key;
    const bool valid = st.c_cand <= static_cast<uint32_t>(kCandMax) && st.c_below <= lo && hi < st.c_below + st.c_cand;
    if (valid) {
        const uint32_t *gc = p.cand + u * p.cand_cap;
        for (int i = tid; i < static_cast<int>(st.c_cand); i += THREADS) cand[i] = gc[i];
        __syncthreads();
        const uint32_t base = float_to_key(st.plo_f);
        const uint32_t width = float_to_key(st.phi_f) - base;
        block_select_norm<THREADS>(cand, static_cast<int>(st.c_cand), base, 32 - __clz(width | 1u), hist, scratch,
                                   lo - st.c_below, hi - st.c_below, a_key, b_key);
        if (tid == 0) atomicAdd(&g_sampled_units, 1ull);
    } else {
        if (tid == 0) atomicAdd(&g_fallback_units, 1ull);
        block_select<THREADS, true>(GlobalStd{p.std + u * p.n, p.vec != 0}, static_cast<int>(p.n), hist, cand, scratch,
                                    lo, hi, false, a_key, b_key);
    }
    float a = key_to_float(a_key), b = key_to_float(b_key);
    float t = quantile_lerp(a, b, w);
    if (st.nan_flag) t = a = b = __int_as_float(0x7fc00000);
    if (tid == 0) {
        p.thr[u] = t;
        if (p.a_out) p.a_out[u] = a;
        if (p.b_out) p.b_out[u] = b;
    }
}

// Large units: after the sweep the exact ranks are found by the histogram rounds (hist_round_kernel /
// select_advance_kernel) over the unit's candidate buffer -- or, when the bracket missed or overflowed, over
// the whole unit (same launches, graceful fallback).  One thread per unit sets up the round state.
__global__ void __launch_bounds__(128) gs_begin_rounds_kernel(const GsParams p, SelectState *state, int tiles_per_unit) {
    __shared__ uint32_t red[4];
    const int64_t u = blockIdx.x;
    const int tid = threadIdx.x;
    GsUnit g = p.st[u];
    if (g.state == 0u && p.below_tile) {   // per-tile counts below the bracket -> c_below
        uint32_t acc = 0;
        for (int t = tid; t < tiles_per_unit; t += 128) acc += p.below_tile[u * tiles_per_unit + t];
        acc = __reduce_add_sync(0xffffffffu, acc);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        g.c_below = red[0] + red[1] + red[2] + red[3];
    }
    if (tid != 0) return;
    SelectState st{};
    st.q = p.q01_per_unit ? p.q01_per_unit[u] : p.q01;
    if (g.state != 0u) {          // ones / zeros / tiny unit: threshold already written by gs_pivot_kernel
        st.mode = kModeDone;
        state[u] = st;
        return;
    }
    st.mode = kModeThreshold;
    quantile_ranks(st.q, p.n, st.lo, st.hi, st.w);
    const bool valid = g.c_cand <= static_cast<uint32_t>(p.cand_cap) && g.c_below <= st.lo &&
                       st.hi < g.c_below + g.c_cand;
    if (valid) {
        st.lo -= g.c_below;
        st.hi -= g.c_below;
        st.pad[0] = 1u;           // rounds read the candidate buffer ...
        st.pad[1] = g.c_cand;     // ... of this many elements
        atomicAdd(&g_sampled_units, 1ull);
    } else {
        atomicAdd(&g_fallback_units, 1ull);
    }
    st.rank = st.lo;
    st.nan_flag = g.nan_flag;
    state[u] = st;
}

}  // namespace pic
