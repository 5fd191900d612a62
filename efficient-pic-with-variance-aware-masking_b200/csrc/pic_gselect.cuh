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
#include <cooperative_groups.h>

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
    // fused-size units: bracket (~8 % at S = 2048) must fit the kCandMax finish buffer; large units: the
    // largest sample the pivot CTA holds, for the narrowest bracket (~6 % of the unit)
    const int s_max = (p.cand_cap == kCandMax) ? kSampleMax : kCandMax;
    int S = (p.cand_cap == kCandMax) ? (n >> 4) : (n >> 3);
    S = S < 1024 ? 1024 : (S > s_max ? s_max : S);
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

// Classification of one tile (shared by gs_sweep_kernel and the tiled select's fused sweep + exchange kernel): parks
// the tile in shared memory, counts the elements below the bracket, sets one hit bit per bracket element
// (VEC: bit 4*i + e <-> element e of float4 (i * THREADS + tid); scalar: bit i <-> element i * THREADS + tid).
template <bool VEC>
__device__ __forceinline__ void gs_classify_tile(const float *base, int len, float plo_f, float phi_f, float4 *park4,
                                                 uint32_t &below, uint32_t &hits, bool &has_nan) {
    constexpr int THREADS = kGsThreads;
    const int tid = threadIdx.x;
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
    gs_classify_tile<VEC>(base, len, plo_f, phi_f, park4, below, hits, has_nan);
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

// ------------------------------------------------------------------------------------------
// Cluster select: ONE launch replaces the three histogram rounds (+ begin / advance / finish) of a large
// unit.  A thread-block cluster of 8 CTAs owns one unit: every round each CTA histograms its stripe of the
// unit's candidate buffer (L2-resident, just written by gs_sweep_kernel) into its own shared memory, the
// cluster synchronises, and every CTA sums the eight histograms through distributed shared memory and
// finds the rank's bin redundantly -- no global histograms, no memsets, no kernel boundaries between rounds.
// Candidate keys are normalised to the bracket, (key - key(plo)) << clz(width), so the 11/11/10-bit digits
// spread over all bins.  A unit whose bracket missed or overflowed is selected from all of std by the same
// code (base 0, shift 0): slower, exact.
// ------------------------------------------------------------------------------------------
constexpr int kClusterCtas = 8;
constexpr int kClusterThreads = 1024;

template <bool VEC>
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
gs_cluster_select_kernel(const GsParams p, int tiles_per_unit) {
    namespace cg = cooperative_groups;
    constexpr int THREADS = kClusterThreads;
    constexpr int UNROLL = 4;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ __align__(16) uint32_t hist[2][kHistBins];   // this CTA's histogram, double-buffered across rounds
    __shared__ __align__(16) uint32_t merged[kHistBins];    // sum over the cluster
    __shared__ uint32_t scratch[kScratchWords];
    __shared__ uint32_t red[THREADS / 32];
    __shared__ uint32_t mn_local;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cluster.block_rank();
    const int64_t u = blockIdx.y;
    GsUnit g = p.st[u];
    if (g.state != 0u) return;   // uniform over the cluster: ones / zeros / tiny unit, done by gs_pivot_kernel
    if (p.below_tile) {          // per-tile counts below the bracket -> c_below
        uint32_t acc = 0;
        for (int t = tid; t < tiles_per_unit; t += THREADS) acc += p.below_tile[u * tiles_per_unit + t];
        acc = __reduce_add_sync(0xffffffffu, acc);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        acc = (lane < THREADS / 32) ? red[lane] : 0u;
        g.c_below = __reduce_add_sync(0xffffffffu, acc);
    }
    const float q = p.q01_per_unit ? p.q01_per_unit[u] : p.q01;
    uint32_t lo, hi;
    float w;
    quantile_ranks(q, p.n, lo, hi, w);
    const bool valid = g.c_cand <= static_cast<uint32_t>(p.cand_cap) && g.c_below <= lo && hi < g.c_below + g.c_cand;
    const float *src = p.std + u * p.n;
    int64_t len = p.n;
    uint32_t base = 0u;
    int lsh = 0;
    bool vec = VEC;
    if (valid) {
        src = reinterpret_cast<const float *>(p.cand) + u * p.cand_cap;   // 16-byte aligned: cand_cap % 4 == 0
        len = g.c_cand;
        lo -= g.c_below;
        hi -= g.c_below;
        base = float_to_key(g.plo_f);
        lsh = __clz((float_to_key(g.phi_f) - base) | 1u);
        vec = true;
    }
    if (tid == 0 && crank == 0) atomicAdd(valid ? &g_sampled_units : &g_fallback_units, 1ull);
    const int64_t nvec = vec ? (len >> 2) : 0;
    const int64_t v0 = nvec * crank / kClusterCtas, v1 = nvec * (crank + 1) / kClusterCtas;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    uint32_t prefix = 0, rank = lo, below_total = 0, count = 0, mn = 0xffffffffu, last_bin = 0;
    uint32_t a_key = 0, b_key = 0;
    bool resolved = false;   // uniform over the cluster (every CTA holds the same merged histograms)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        if (resolved) break;
        constexpr int kUp[3] = {32, 21, 10};   // bits above round r's digit
        const int shift = round_shift(r), nbins = round_bins(r), up = kUp[r];
        uint32_t *h = hist[r & 1];
        for (int j = tid; j < nbins; j += THREADS) h[j] = 0u;
        __syncthreads();
        const uint32_t want = (r > 0) ? (prefix >> up) : 0u;
        auto visit = [&](float x, bool inb) {
            const uint32_t k = (float_to_key(x) - base) << lsh;
            bool match = inb;
            if (r > 0) {
                const uint32_t hb = k >> up;
                if (r == 2 && inb && hb > want) mn = min(mn, k);
                match = inb && hb == want;
            }
            hist_add(h, (k >> shift) & round_mask(r), match);
        };
        for (int64_t jb = v0 + (tid - lane) * UNROLL; jb < v1; jb += THREADS * UNROLL) {
            float4 v[UNROLL];
            bool inb[UNROLL];
#pragma unroll
            for (int i = 0; i < UNROLL; ++i) {
                const int64_t j = jb + i * 32 + lane;
                inb[i] = j < v1;
                v[i] = inb[i] ? __ldg(s4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < UNROLL; ++i) {
                visit(v[i].x, inb[i]); visit(v[i].y, inb[i]); visit(v[i].z, inb[i]); visit(v[i].w, inb[i]);
            }
        }
        // scalar remainder (everything when the source is not float4-addressable), striped over the cluster
        const int64_t t0 = nvec << 2, tn = len - t0;
        const int64_t e0 = t0 + tn * crank / kClusterCtas, e1 = t0 + tn * (crank + 1) / kClusterCtas;
        for (int64_t jb = e0 + tid - lane; jb < e1; jb += THREADS) {
            const int64_t j = jb + lane;
            const bool inb = j < e1;
            visit(inb ? __ldg(src + j) : 0.0f, inb);
        }
        if (r == 2) {
            mn = __reduce_min_sync(0xffffffffu, mn);
            if (tid == 0) mn_local = 0xffffffffu;
            __syncthreads();
            if (lane == 0 && mn != 0xffffffffu) atomicMin(&mn_local, mn);
        }
        __syncthreads();
        cluster.sync();   // every CTA's histogram of this round is complete and visible
        for (int j = tid; j < nbins; j += THREADS) {
            uint32_t sum = 0;
#pragma unroll
            for (int c = 0; c < kClusterCtas; ++c) sum += cluster.map_shared_rank(h, c)[j];
            merged[j] = sum;
        }
        __syncthreads();
        const BinHit hit = block_find_bin<THREADS>(merged, nbins, rank, scratch);
        prefix |= hit.bin << shift;
        rank -= hit.below;
        below_total += hit.below;
        count = hit.count;
        last_bin = hit.bin;
        if (r == 1 && lsh >= 10) {
            // the normalised keys have no bits below the second digit (bracket narrower than 2^22 keys, the usual
            // case): the 22-bit prefix IS the key, and its successor is the next non-empty bin of this round -- the
            // third pass is only needed when that bin lies in another first-round bucket
            if (hi < below_total + count) {
                a_key = b_key = prefix;
                resolved = true;
            } else {
                const uint32_t nb = block_next_nonempty<THREADS>(merged, round_bins(1), last_bin, scratch);
                if (nb != 0xffffffffu) {
                    a_key = prefix;
                    b_key = (prefix & ~(round_mask(1) << round_shift(1))) | (nb << round_shift(1));
                    resolved = true;
                }
            }
        }
    }
    if (!resolved) {
        a_key = b_key = prefix;
    }
    if (!resolved && !(hi < below_total + count)) {
        const uint32_t nb = block_next_nonempty<THREADS>(merged, round_bins(2), last_bin, scratch);
        if (nb != 0xffffffffu) {
            b_key = (prefix & ~round_mask(2)) | nb;
        } else {          // next key lies beyond the last digit's range: smallest key above, over the cluster
            uint32_t m = 0xffffffffu;
            for (int c = 0; c < kClusterCtas; ++c) m = min(m, *cluster.map_shared_rank(&mn_local, c));
            b_key = m;
        }
    }
    cluster.sync();       // no CTA may exit while a peer still reads its shared memory
    if (crank != 0 || tid != 0) return;
    float a = key_to_float((a_key >> lsh) + base), b = key_to_float((b_key >> lsh) + base);
    float t = quantile_lerp(a, b, w);
    if (g.nan_flag) t = a = b = __int_as_float(0x7fc00000);
    p.thr[u] = t;
    if (p.a_out) p.a_out[u] = a;
    if (p.b_out) p.b_out[u] = b;
}

}  // namespace pic
