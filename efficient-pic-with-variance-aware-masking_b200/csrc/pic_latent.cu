// pic_latent.cu -- kernels + C ABI of libpic_latent.so (sm_100a only).
//
// Hot path per progressive slice (reference: models/pic.py:583-584, 621-629, 809-820):
//   select threshold (torch.quantile semantics) -> mask -> mean substitution -> round/noise
//   quantisation -> Gaussian likelihood -> scale-table index / symbols / rate.
//
// Kernels
//   slice_fused_kernel      one CTA per unit (n <= kFusedMaxElems): std tile -> order-preserving
//                           keys in shared memory + round-0 histogram in the same sweep,
//                           3-round radix select in shared memory, then the apply sweep reads
//                           std back from shared memory.  HBM traffic = compulsory bytes only.
//   hist_round_kernel /     multi-CTA select for large units and for spatially tiled units
//   select_advance_kernel / (histograms merged in global memory; all-reducible between rounds)
//   select_finish_kernel
//   slice_apply_kernel      elementwise apply with given thresholds (tile per CTA)
//   slice_backward_kernel   fused backward
//   gaussian_* / build_indexes / quantize / dequantize / log_sum   un-fused operators
// All global accesses are 128-bit when n % 4 == 0 and the pointers are 16-byte aligned.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is bound at run time (dlopen), never linked
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/pic_latent.h"
#include "pic_math.cuh"
#include "pic_fast.cuh"
#include "pic_select.cuh"
#include "pic_params.h"

namespace pic {

thread_local int g_last_cuda_error = 0;

// diagnostics: units whose sampled bracket missed (counted since library load)
__device__ unsigned long long g_fallback_units = 0ull;
__device__ unsigned long long g_sampled_units = 0ull;
#ifdef PIC_PHASE_TIMING
__device__ long long g_phase_clk[16];
#define PIC_PHASE(i) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == PIC_PHASE_BLOCK) g_phase_clk[i] = clock64(); } while (0)
#else
#define PIC_PHASE(i) do {} while (0)
#endif

struct SelectState {
    uint32_t lo, hi;
    float w, q;
    uint32_t prefix, rank, below_total;
    uint32_t a_key, b_key, need_min, nan_flag, count;
    int32_t mode;
    uint32_t pad[3];   // [0] 1 = rounds read the unit's candidate buffer (large-unit sampled select), [1] its length
};
static_assert(sizeof(SelectState) == 64, "SelectState must stay 64 bytes");

// ------------------------------------------------------------------------------------------
// per-element apply (two elements per call: pic_fast.cuh::apply_pair)
// ------------------------------------------------------------------------------------------
// Shared-memory staging of the scale table for the index computation.  Layout (floats):
// [0,64) table copy, [64,192) float2 pairs (tbl[k-1], tbl[k]), [192] lg2(t0), [193] inv_step,
// [194] geometric flag.
constexpr int kIndexSmemFloats = 64 + 128 + 8;

__device__ __forceinline__ IndexCtx index_ctx_setup(const SliceParams &p, float *sm) {
    IndexCtx ic;
    ic.len = p.table_len;
    ic.tbl64 = (p.table_len == kTableSmem) && p.idx;
    ic.tbl = p.table;
    ic.geometric = false;
    ic.inv_step = 0.0f;
    ic.bias = 0.0f;
    if (!ic.tbl64) return ic;  // uniform
    const int tid = threadIdx.x;
    if (tid < kTableSmem) sm[tid] = p.table[tid];
    if (tid == 0) {
        const float l0 = lg2_approx(p.table[0]);
        const float l1 = lg2_approx(p.table[kTableSmem - 1]);
        const float inv = static_cast<float>(kTableSmem - 1) / (l1 - l0);
        sm[192] = l0;
        sm[193] = inv;
    }
    __syncthreads();
    bool ok = true;
    if (tid < kTableSmem) {
        const float x = (lg2_approx(sm[tid]) - sm[192]) * sm[193];
        ok = fabsf(x - static_cast<float>(tid)) < 1e-3f;      // table is geometric to well within eps
    }
    const int all_ok = __syncthreads_and(ok ? 1 : 0);
    ic.tbl = sm;
    ic.inv_step = sm[193];
    ic.bias = -(sm[192] * sm[193] + 4e-3f);
    ic.geometric = all_ok != 0 && (sm[193] > 0.0f);
    return ic;
}

// Applies the slice arithmetic to elements [0, len) of a unit-local range starting at global
// element offset `off`.  keys != nullptr: std comes from the shared-memory key tile (index 0 ==
// first element of the range); otherwise from global memory.
// OUTS: compile-time set of optional tensors (bit 0 y_base, 1 mask, 2 y_hat, 3 lik, 4 idx,
// 5 symbols, 6 rate) or -1 = decided at run time from the pointers (generic flavour).
constexpr int kOutsCodec = 0x1f;   // y_base + mask, y_hat, lik, idx   (compress / eval forward)
constexpr int kOutsTrain = 0x0f;   // y_base + mask, y_hat, lik        (training forward)
constexpr int kOutsGeneric = -1;
constexpr int kOutsSelectOnly = -3; // threshold selection only (no apply sweep compiled in)

__host__ __device__ inline int outs_of(const SliceParams &p) {
    return (p.y_base ? 1 : 0) | (p.mask ? 2 : 0) | (p.y_hat ? 4 : 0) | (p.lik ? 8 : 0) | (p.idx ? 16 : 0) |
           (p.symbols ? 32 : 0) | (p.rate ? 64 : 0);
}

template <bool TRAIN, bool VEC, int THREADS, int OUTS>
__device__ __forceinline__ float apply_range(const SliceParams &p, int64_t off, int len,
                                             const uint32_t *keys, int mode, float thr_in,
                                             const IndexCtx &ic, float4 *stage = nullptr, int64_t in_off = -1) {
    // in_off: element offset of the INPUTS when it differs from the outputs' (multi-quality apply: `repeat`
    // consecutive output units share one input unit)
    if (in_off < 0) in_off = off;
    float rate_acc = 0.0f;
    const int tid = threadIdx.x;
    const bool full = p.apply_kind == 2;
    const bool force_one = mode == kModeOnes;
    const float thr = (mode == kModeZeros) ? __int_as_float(0x7fc00000) : thr_in;  // s >= NaN is false
    const int outs = (OUTS >= 0) ? OUTS : outs_of(p);
    const bool has_base = outs & 1, want_mask = outs & 2, want_yhat = outs & 4;
    const bool want_lik = outs & (8 | 64), store_lik = outs & 8;
    const bool want_idx = outs & 16, want_sym = outs & 32, want_rate = outs & 64;
    if (VEC) {
        const int nvec = len >> 2;
        const uint64_t pol_first = policy_evict_first();
        // one 32-bit vector index for every tensor (host guarantees total elements < 2^34)
        const uint32_t vbase = static_cast<uint32_t>(off >> 2);
        const uint32_t vin = static_cast<uint32_t>(in_off >> 2);
        const float4 *std4 = reinterpret_cast<const float4 *>(p.std) + vin - vbase;   // inputs are indexed with vi too
        const float4 *yt4 = reinterpret_cast<const float4 *>(p.y_top) + vin - vbase;
        const float4 *yb4 = reinterpret_cast<const float4 *>(p.y_base) + vin - vbase;
        const float4 *mu4 = reinterpret_cast<const float4 *>(p.mu) + vin - vbase;
        const float4 *nz4 = reinterpret_cast<const float4 *>(p.noise) + vin - vbase;
        const uint4 *k4 = reinterpret_cast<const uint4 *>(keys);
        // Software pipeline: with a stage buffer the inputs of iteration i+1 are copied
        // global->shared (cp.async, per-thread private 16-byte slots: no barrier needed) while
        // iteration i is computed, so ~2x the bytes are in flight per SM at no register cost.
        constexpr int NARR = TRAIN ? 5 : 4;
        const bool piped = (stage != nullptr) && full && (keys == nullptr);
        // slot(st, arr) = sbase + st * kStageBytes + arr * kArrBytes (32-bit shared addresses)
        constexpr uint32_t kArrBytes = THREADS * sizeof(float4), kStageBytes = NARR * kArrBytes;
        const uint32_t sbase = piped ? static_cast<uint32_t>(__cvta_generic_to_shared(stage + tid)) : 0u;
        auto prefetch = [&](int jj, int st) {
            const uint32_t v = vbase + static_cast<uint32_t>(jj);
            const uint32_t sa = sbase + (st ? kStageBytes : 0u);
            cp_async16(sa, std4 + v, pol_first);
            cp_async16(sa + kArrBytes, yt4 + v, pol_first);
            if (has_base) cp_async16(sa + 2 * kArrBytes, yb4 + v, pol_first);
            cp_async16(sa + 3 * kArrBytes, mu4 + v, pol_first);
            if (TRAIN) cp_async16(sa + (NARR - 1) * kArrBytes, nz4 + v, pol_first);
        };
        // The sweep walks the unit BACKWARDS: in the fused kernel the select sweep has just read
        // std front to back, so the tail of the unit is the part most likely still in L2.
        const int iters = (nvec - tid + THREADS - 1) / THREADS;     // this thread's float4 count (may be <= 0)
        const int jlast = tid + (iters - 1) * THREADS;
        if (piped && iters > 0) prefetch(jlast, 0);
        if (piped) cp_async_commit();
        int st = 0;
        for (int j = jlast; j >= 0; j -= THREADS, st ^= 1) {
            const uint32_t vi = vbase + static_cast<uint32_t>(j);
            float s[4];
            float4 ytv, muv, ybv = make_float4(0.f, 0.f, 0.f, 0.f), nzv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (piped) {
                if (j - THREADS >= 0) prefetch(j - THREADS, st ^ 1);
                cp_async_commit();
                cp_async_wait<1>();          // everything but the newest group has landed
                const uint32_t sa = sbase + (st ? kStageBytes : 0u);
                const float4 v = lds128(sa);
                s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
                ytv = lds128(sa + kArrBytes);
                if (has_base) ybv = lds128(sa + 2 * kArrBytes);
                muv = lds128(sa + 3 * kArrBytes);
                if (TRAIN) nzv = lds128(sa + (NARR - 1) * kArrBytes);
            } else {
                if (keys) {
                    const uint4 k = k4[j];
                    s[0] = key_to_float(k.x); s[1] = key_to_float(k.y);
                    s[2] = key_to_float(k.z); s[3] = key_to_float(k.w);
                } else {
                    const float4 v = ld_hint(std4 + vi, pol_first);   // second (last) use of std: may leave L2
                    s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
                }
                if (!full) {
                    float mk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) mk[e] = ((s[e] >= thr) || force_one) ? 1.0f : 0.0f;
                    reinterpret_cast<float4 *>(p.mask)[vi] = make_float4(mk[0], mk[1], mk[2], mk[3]);
                    continue;
                }
                ytv = ld_hint(yt4 + vi, pol_first);
                muv = ld_hint(mu4 + vi, pol_first);
                if (has_base) ybv = ld_hint(yb4 + vi, pol_first);
                if (TRAIN) nzv = ld_hint(nz4 + vi, pol_first);
            }
            const float yt[4] = {ytv.x, ytv.y, ytv.z, ytv.w}, yb[4] = {ybv.x, ybv.y, ybv.z, ybv.w};
            const float mu[4] = {muv.x, muv.y, muv.z, muv.w}, nz[4] = {nzv.x, nzv.y, nzv.z, nzv.w};
            PairOut o[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                apply_pair<TRAIN>(s + 2 * h, yt + 2 * h, yb + 2 * h, mu + 2 * h, nz + 2 * h, has_base, thr,
                                  force_one, p.scale_bound, p.lik_bound, want_lik, want_idx, want_sym, ic, o[h]);
                if (want_rate) rate_acc += logf(o[h].lik[0]) + logf(o[h].lik[1]);
            }
            if (want_mask) st_hint(reinterpret_cast<float4 *>(p.mask) + vi, make_float4(o[0].m[0], o[0].m[1], o[1].m[0], o[1].m[1]), pol_first);
            if (want_yhat) st_hint(reinterpret_cast<float4 *>(p.y_hat) + vi, make_float4(o[0].y_hat[0], o[0].y_hat[1], o[1].y_hat[0], o[1].y_hat[1]), pol_first);
            if (store_lik) st_hint(reinterpret_cast<float4 *>(p.lik) + vi, make_float4(o[0].lik[0], o[0].lik[1], o[1].lik[0], o[1].lik[1]), pol_first);
            if (want_idx) st_hint(reinterpret_cast<int4 *>(p.idx) + vi, make_int4(o[0].idx[0], o[0].idx[1], o[1].idx[0], o[1].idx[1]), pol_first);
            if (want_sym) st_hint(reinterpret_cast<int4 *>(p.symbols) + vi, make_int4(o[0].sym[0], o[0].sym[1], o[1].sym[0], o[1].sym[1]), pol_first);
        }
        if (piped) cp_async_wait<0>();
    } else {
        for (int j = tid; j < len; j += THREADS) {
            const float sv = keys ? key_to_float(keys[j]) : __ldg(p.std + in_off + j);
            if (!full) {
                p.mask[off + j] = ((sv >= thr) || force_one) ? 1.0f : 0.0f;
                continue;
            }
            const float ytv = __ldg(p.y_top + in_off + j);
            const float ybv = has_base ? __ldg(p.y_base + in_off + j) : 0.0f;
            const float muv = __ldg(p.mu + in_off + j);
            const float nzv = TRAIN ? __ldg(p.noise + in_off + j) : 0.0f;
            const float s[2] = {sv, sv}, yt[2] = {ytv, ytv}, yb[2] = {ybv, ybv}, mu[2] = {muv, muv}, nz[2] = {nzv, nzv};
            PairOut o;
            apply_pair<TRAIN>(s, yt, yb, mu, nz, has_base, thr, force_one, p.scale_bound, p.lik_bound, want_lik,
                              want_idx, want_sym, ic, o);
            if (want_rate) rate_acc += logf(o.lik[0]);
            if (want_mask) p.mask[off + j] = o.m[0];
            if (want_yhat) p.y_hat[off + j] = o.y_hat[0];
            if (store_lik) p.lik[off + j] = o.lik[0];
            if (want_idx) p.idx[off + j] = o.idx[0];
            if (want_sym) p.symbols[off + j] = o.sym[0];
        }
    }
    return rate_acc;
}

// block reduction of the per-thread rate partials in f64; result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ double block_sum_f64(float v, double *sh /* THREADS/32 doubles */) {
    double d = static_cast<double>(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = d;
    __syncthreads();
    double total = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < THREADS / 32; ++w) total += sh[w];
    __syncthreads();
    return total;
}

// ------------------------------------------------------------------------------------------
// fused kernel: one CTA per unit, several CTAs resident per SM
// ------------------------------------------------------------------------------------------
// Sampled-pivot exact select (Floyd-Rivest style) of the two order statistics of one unit:
//   1. S <= 4096 keys are sampled (strided with a hashed jitter) into shared memory,
//   2. two sample order statistics bracketing the wanted rank by ~4 sigma become pivots,
//   3. ONE streaming sweep over the unit's std (the only HBM read of std) counts the keys
//      below the bracket and compacts the bracket (about n * 8 sigma / S keys) into `cand`,
//   4. the exact ranks are found among the candidates with the shared-memory radix select.
// Returns false (uniformly) when the bracket missed or overflowed; the caller then runs the
// full histogram select.  NaN presence is reported through scratch[39].
// WIDE (latency variants of the select-only kernel, one or two wide CTAs per SM): a 4 x THREADS float4 park
// area outside the histogram, so every thread keeps four 128-bit loads in flight whatever the CTA width.
#ifdef PIC_SELECT_PIPE
constexpr bool kSelectPipe = true;    // experiment: cp.async double-buffered park area in the 256-thread select
#else
constexpr bool kSelectPipe = false;
#endif
template <int THREADS, bool VEC, bool WIDE = false, bool PIPE = false>
__device__ __forceinline__ bool sampled_select(const float *std_u, int n, uint32_t lo, uint32_t hi,
                                               uint32_t *hist, uint32_t *cand, uint32_t *scratch,
                                               uint32_t &a_key, uint32_t &b_key, float4 *park_wide = nullptr) {
    const int tid = threadIdx.x;
    int S = n >> 4;                    // ~6 % of the unit's cache lines are touched by the sample
    S = S < 1024 ? 1024 : (S > kSampleMax ? kSampleMax : S);
    S &= ~3;
    PIC_PHASE(0);
    // ---- 1. sample ----------------------------------------------------------------------
    if (VEC) {
        const int nvec = n >> 2, S4 = S >> 2;
        const int stride = nvec / S4;  // >= 1 because n > kCandMax >= S
        const float4 *s4 = reinterpret_cast<const float4 *>(std_u);
        const uint64_t pol_keep = policy_evict_last();   // the sweep re-reads these lines within microseconds
        for (int i = tid; i < S4; i += THREADS) {
            const uint32_t jit = ((static_cast<uint32_t>(i) * 0x9E3779B1u) >> 12) % static_cast<uint32_t>(stride);
            const float4 v = ld_hint(s4 + static_cast<size_t>(i) * stride + jit, pol_keep);
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
    PIC_PHASE(1);
    // ---- 2. pivots: sample ranks kt -+ 4 sigma, resolved together to 22-bit buckets --------
    const float frac = static_cast<float>(lo) / static_cast<float>(n > 1 ? n - 1 : 1);
    const float kt = frac * static_cast<float>(S - 1);
    const float margin = 3.5f * sqrtf(static_cast<float>(S) * frac * (1.0f - frac)) + 4.0f;
    const int klo = static_cast<int>(floorf(kt - margin));
    const int khi = static_cast<int>(ceilf(kt + margin));
    uint32_t plo_key, phi_key;
    block_bracket_pair<THREADS>(cand, S, hist, scratch, klo > 0 ? klo : 0, khi < S - 1 ? khi : S - 1,
                                plo_key, phi_key);
    // open-ended bracket at the extremes of the sample; pivots as floats for the sweep
    const float plo_f = (klo > 0) ? key_to_float(plo_key) : -INFINITY;
    const float phi_f = (khi < S - 1) ? key_to_float(phi_key) : INFINITY;
    if (tid == 0) { scratch[40] = 0u; scratch[41] = 0u; }
    __syncthreads();  // cand (the sample) may now be overwritten
    PIC_PHASE(2);
    // ---- 3. sweep: count below, append the bracket's keys ------------------------------------
    // float-domain compares (== key order for non-NaN, -0 == +0); NaN fails every compare.
    // Each iteration's values are parked in the (currently idle) histogram region of shared
    // memory -- per-thread private float4 slots -- so a hit can be fetched by its bit index
    // without dynamic register indexing and appended as a key right away.
    uint32_t below = 0;
    bool has_nan = false;
    if (VEC) {
        // float4 per thread per iteration: the 16 KB histogram region parks 4 x 256 float4; wider CTAs get
        // their own park area (kWideThreads) or fewer loads in flight (512)
        constexpr int VPI = (THREADS <= 256 || WIDE) ? 4 : 2;
        const int nvec = n >> 2;
        const uint64_t pol_last = policy_evict_last();
        const float4 *s4 = reinterpret_cast<const float4 *>(std_u);
        float4 *park4 = WIDE ? park_wide : reinterpret_cast<float4 *>(hist);
        const float *park = reinterpret_cast<const float *>(park4) + tid * 4;   // the current iteration's slots
        f2 below2 = pk(0.0f, 0.0f);
        auto classify4 = [&](const float4 &q, uint32_t &hits4) {   // hits4: bits 0..3 of this float4
            const float mx = max_nan(max_nan(q.x, q.y), max_nan(q.z, q.w));
            has_nan |= (mx != mx);
            below2 = add2(below2, pk(fset_lt(q.x, plo_f), fset_lt(q.y, plo_f)));   // exact: counts << 2^24
            below2 = add2(below2, pk(fset_lt(q.z, plo_f), fset_lt(q.w, plo_f)));
            or_if_in_range<1u>(hits4, q.x, plo_f, phi_f);
            or_if_in_range<2u>(hits4, q.y, plo_f, phi_f);
            or_if_in_range<4u>(hits4, q.z, plo_f, phi_f);
            or_if_in_range<8u>(hits4, q.w, plo_f, phi_f);
        };
        auto classify = [&](const float4 &q, uint32_t &hits, int sh) {
            uint32_t h4 = 0;
            classify4(q, h4);
            hits |= h4 << sh;
        };
        auto append = [&](uint32_t hits) {   // bit e -> park[(e >> 2) * THREADS * 4 + (e & 3)]
            // one slot reservation per WARP (inclusive shuffle scan of the hit counts): the
            // per-CTA counter would otherwise see ~2k serialised same-address atomics per unit
            const uint32_t cnt = static_cast<uint32_t>(__popc(hits));
            if (__ballot_sync(0xffffffffu, cnt != 0u) == 0u) return;
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if ((tid & 31) >= d) incl += t;
            }
            uint32_t base = 0;
            if ((tid & 31) == 31) base = atomicAdd(&scratch[40], incl);
            base = __shfl_sync(0xffffffffu, base, 31);
            uint32_t pos = base + incl - cnt;
            while (hits) {
                const int e = __ffs(hits) - 1;
                hits &= hits - 1u;
                const float x = park[(e >> 2) * (THREADS * 4) + (e & 3)];
                if (pos < static_cast<uint32_t>(kCandMax)) cand[pos] = float_to_key(x);
                ++pos;
            }
        };
        // warp-uniform trip counts (jb is the warp's first index): out-of-range lanes idle
        const int lane = tid & 31;
        int jb = tid - lane;
        if (PIPE) {
            // full iterations through two park buffers filled by cp.async: the next iteration's loads are in
            // flight while this one is classified (per-thread private slots: no barrier needed)
            float4 *const buf0 = reinterpret_cast<float4 *>(hist);
            const int iters = nvec / (VPI * THREADS);
            auto issue = [&](int it, int b) {
#pragma unroll
                for (int i = 0; i < VPI; ++i)
                    cp_async16(static_cast<uint32_t>(__cvta_generic_to_shared((b ? park_wide : buf0) + i * THREADS + tid)),
                               s4 + it * (VPI * THREADS) + i * THREADS + tid, pol_last);
                cp_async_commit();
            };
            if (iters > 0) issue(0, 0);
            for (int it = 0; it < iters; ++it) {
                const int b = it & 1;
                if (it + 1 < iters) {
                    issue(it + 1, b ^ 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                const float4 *cur = b ? park_wide : buf0;
                park = reinterpret_cast<const float *>(cur) + tid * 4;
                uint32_t hits = 0;
#pragma unroll
                for (int i = 0; i < VPI; ++i) classify(cur[i * THREADS + tid], hits, 4 * i);
                append(hits);
            }
            park = reinterpret_cast<const float *>(park4) + tid * 4;
            jb += iters * VPI * THREADS;
        }
        for (; jb + (VPI - 1) * THREADS + 31 < nvec; jb += VPI * THREADS) {   // VPI independent 128-bit loads in flight
            const int j = jb + lane;
            float4 v[VPI];
#pragma unroll
            for (int i = 0; i < VPI; ++i) v[i] = ld_hint(s4 + j + i * THREADS, pol_last);
            uint32_t hits = 0;
#pragma unroll
            for (int i = 0; i < VPI; ++i) {
                park4[i * THREADS + tid] = v[i];
                classify(v[i], hits, 4 * i);
            }
            append(hits);
        }
        for (; jb < nvec; jb += THREADS) {
            const int j = jb + lane;
            uint32_t hits = 0;
            if (j < nvec) {
                const float4 v0 = ld_hint(s4 + j, pol_last);
                park4[tid] = v0;
                classify(v0, hits, 0);
            }
            append(hits);
        }
        float b_lo, b_hi;
        unpk(below2, b_lo, b_hi);
        below = static_cast<uint32_t>(b_lo + b_hi);
    } else {
        for (int j = tid; j < n; j += THREADS) {
            const float x = __ldg(std_u + j);
            has_nan |= (x != x);
            below += (x < plo_f) ? 1u : 0u;
            if (x >= plo_f && x <= phi_f) {
                const uint32_t pos = atomicAdd(&scratch[40], 1u);
                if (pos < static_cast<uint32_t>(kCandMax)) cand[pos] = float_to_key(x);
            }
        }
    }
    below = __reduce_add_sync(0xffffffffu, below);
    if ((tid & 31) == 0 && below) atomicAdd(&scratch[41], below);
    if (__any_sync(0xffffffffu, has_nan) && (tid & 31) == 0) scratch[39] = 1u;
    __syncthreads();
    const uint32_t c_cand = scratch[40], c_below = scratch[41];
    PIC_PHASE(3);
    // ---- 4. exact ranks among the candidates ------------------------------------------------
    const bool valid = c_cand <= static_cast<uint32_t>(kCandMax) && c_below <= lo && hi < c_below + c_cand;
    if (!valid) {
        __syncthreads();
        return false;
    }
    __syncthreads();   // the park area is the histogram: everyone must be done with it
    const uint32_t base = float_to_key(plo_f);
    const uint32_t width = float_to_key(phi_f) - base;
    block_select_norm<THREADS>(cand, static_cast<int>(c_cand), base, 32 - __clz(width | 1u), hist, scratch,
                               lo - c_below, hi - c_below, a_key, b_key);
    PIC_PHASE(4);
    return true;
}

template <int THREADS>
struct FusedSmem {
    uint32_t scratch[kScratchWords];
    float index[kIndexSmemFloats];
    double red[THREADS / 32];
};

// dynamic shared memory: max(select buffers, cp.async stage buffer) -- the two are never live
// at the same time (select finishes before the unit's apply sweep starts)
constexpr size_t kSelectSmemBytes = (2 * kHistBins + kCandMax) * sizeof(uint32_t);
template <int THREADS, int OUTS>
__host__ __device__ constexpr bool wide_select() { return OUTS == -3 /* kOutsSelectOnly */ && THREADS > 256; }
template <int THREADS>
__host__ __device__ constexpr size_t wide_park_bytes() { return size_t(4) * THREADS * sizeof(float4); }   // 32 KB (512) / 64 KB (1024)
template <bool TRAIN, int THREADS>
constexpr size_t fused_dyn_smem() {
    const size_t stage = size_t(2) * (TRAIN ? 5 : 4) * THREADS * sizeof(float4);
    return stage > kSelectSmemBytes ? stage : kSelectSmemBytes;
}

template <bool TRAIN, bool VEC, int THREADS, int OUTS>
__global__ void __launch_bounds__(THREADS, (OUTS == kOutsSelectOnly && !kSelectPipe ? 1280 : 1024) / THREADS)
slice_fused_kernel(const SliceParams p) {
    __shared__ __align__(16) FusedSmem<THREADS> sm;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    float4 *stage = p.use_stage ? reinterpret_cast<float4 *>(dyn_smem) : nullptr;
    uint32_t *hist = reinterpret_cast<uint32_t *>(dyn_smem);
    uint32_t *cand = hist + 2 * kHistBins;
    constexpr bool WIDE = wide_select<THREADS, OUTS>();
    constexpr bool PIPE = kSelectPipe && OUTS == kOutsSelectOnly && THREADS == 256 && VEC;
    float4 *park_wide = (WIDE || PIPE) ? reinterpret_cast<float4 *>(dyn_smem + kSelectSmemBytes) : nullptr;
    uint32_t *scratch = sm.scratch;
    const int n = static_cast<int>(p.n);
    const int tid = threadIdx.x;
    IndexCtx ic{};
    if (OUTS != kOutsSelectOnly) ic = index_ctx_setup(p, sm.index);

    for (int64_t u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int64_t off = ((OUTS == kOutsSelectOnly && p.repeat > 1) ? u / p.repeat : u) * p.n;
        const float q = p.q01_per_unit ? p.q01_per_unit[u] : p.q01;
        const int mode = unit_mode(q);
        float thr = (mode == kModeOnes) ? -INFINITY : INFINITY;
        float a_val = thr, b_val = thr;
        if (p.thr_in && mode == kModeThreshold) {
            thr = p.thr_in[u];
            a_val = b_val = thr;
        } else if (mode == kModeThreshold) {
            if (tid == 0) scratch[39] = 0u;
            __syncthreads();
            uint32_t lo, hi;
            float w;
            quantile_ranks(q, p.n, lo, hi, w);
            uint32_t a_key = 0, b_key = 0;
            const float *std_u = p.std + off;
            if (n <= kCandMax) {
                // small unit: all keys are candidates
                bool has_nan = false;
                for (int j = tid; j < n; j += THREADS) {
                    const float v = __ldg(std_u + j);
                    has_nan |= (v != v);
                    cand[j] = float_to_key(v);
                }
                if (__any_sync(0xffffffffu, has_nan) && (tid & 31) == 0) scratch[39] = 1u;
                __syncthreads();
                block_select_norm<THREADS>(cand, n, 0u, 32, hist, scratch, lo, hi, a_key, b_key);
            } else if (sampled_select<THREADS, VEC, WIDE, PIPE>(std_u, n, lo, hi, hist, cand, scratch, a_key, b_key, park_wide)) {
                if (tid == 0) atomicAdd(&g_sampled_units, 1ull);
            } else {
                if (tid == 0) atomicAdd(&g_fallback_units, 1ull);
                // bracket missed or overflowed (heavy ties, adversarial order): full histogram select
                // over the unit's std (L2-resident after the sweep)
                block_select<THREADS, true>(GlobalStd{std_u, VEC}, n, hist, cand, scratch, lo, hi, false, a_key, b_key);
            }
            a_val = key_to_float(a_key);
            b_val = key_to_float(b_key);
            thr = quantile_lerp(a_val, b_val, w);
            if (scratch[39] != 0u) thr = a_val = b_val = __int_as_float(0x7fc00000);
        }
        if (tid == 0) {
            if (p.thr_out) p.thr_out[u] = thr;
            if (p.a_out) p.a_out[u] = a_val;
            if (p.b_out) p.b_out[u] = b_val;
        }
        if (OUTS != kOutsSelectOnly && p.apply_kind != 0) {
            const float acc = apply_range<TRAIN, VEC, THREADS, OUTS>(p, off, n, nullptr, mode, thr, ic, stage);
            if ((OUTS >= 0) ? ((OUTS & 64) != 0) : (p.rate != nullptr)) {
                const double total = block_sum_f64<THREADS>(acc, sm.red);
                if (tid == 0) p.rate[u] = total;
            }
        }
        __syncthreads();  // hist / cand / scratch are reused by the next unit
    }
}

}  // namespace pic
#include "pic_gselect.cuh"
namespace pic {

// ------------------------------------------------------------------------------------------
// multi-CTA select rounds (large units; spatially tiled units)
// ------------------------------------------------------------------------------------------
__global__ void select_begin_kernel(SelectState *state, int64_t n_total, int64_t units, float q01,
                                    const float *q01_per_unit) {
    const int64_t u = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (u >= units) return;
    SelectState st;
    const float q = q01_per_unit ? q01_per_unit[u] : q01;
    st.q = q;
    st.mode = unit_mode(q);
    st.lo = st.hi = 0;
    st.w = 0.0f;
    if (st.mode == kModeThreshold) quantile_ranks(q, n_total, st.lo, st.hi, st.w);
    st.prefix = 0;
    st.rank = st.lo;
    st.below_total = 0;
    st.a_key = st.b_key = 0;
    st.need_min = 0;
    st.nan_flag = 0;
    st.count = 0;
    st.pad[0] = st.pad[1] = st.pad[2] = 0;
    state[u] = st;
}

// grid: (chunks, units).  hist: [units][kHistWords] (pre-zeroed); min_above: [units] (0xffffffff).
// cand != nullptr: units whose state says so read their candidate buffer (raw floats, cand_cap per unit)
// instead of std -- the large-unit sampled select; the other units (bracket fallback) read all of std.
template <int ROUND, bool VEC>
__global__ void __launch_bounds__(512) hist_round_kernel(const float *std, int64_t n_local,
                                                         const SelectState *state, uint32_t *hist,
                                                         uint32_t *min_above, const float *cand, int64_t cand_cap) {
    constexpr int THREADS = 512;
    __shared__ uint32_t sh[kHistBins];
    const int64_t u = blockIdx.y;
    const SelectState st = state[u];
    if (st.mode != kModeThreshold) return;
    const bool from_cand = cand != nullptr && st.pad[0] != 0u;
    if (from_cand) n_local = st.pad[1];
    if (static_cast<int64_t>(blockIdx.x) * kRoundChunk >= n_local) return;
    constexpr int shift = round_shift(ROUND);
    constexpr int nbins = round_bins(ROUND);
    constexpr int up = shift + (ROUND == 1 ? 11 : 10);  // bits above this round's digit
    const int tid = threadIdx.x, lane = tid & 31;
    for (int j = tid; j < nbins; j += THREADS) sh[j] = 0u;
    __syncthreads();
    const int64_t begin = static_cast<int64_t>(blockIdx.x) * kRoundChunk;
    const int64_t end = min(n_local, begin + kRoundChunk);
    const float *base = from_cand ? cand + u * cand_cap : std + u * n_local;
    const uint32_t want = (ROUND > 0) ? (st.prefix >> up) : 0u;
    uint32_t mn = 0xffffffffu;
    bool has_nan = false;
    auto visit = [&](float s, bool inb) {
        if (ROUND == 0) has_nan |= (inb && s != s);
        const uint32_t k = float_to_key(s);
        bool match = inb;
        if (ROUND > 0) {
            const uint32_t hb = k >> up;
            if (ROUND == 2 && inb && hb > want) mn = min(mn, k);
            match = inb && hb == want;
        }
        hist_add(sh, (k >> shift) & round_mask(ROUND), match);
    };
    if (VEC) {
        const int64_t nvec = (end - begin) >> 2;  // chunk starts are multiples of 4
        const float4 *p4 = reinterpret_cast<const float4 *>(base + begin);
        for (int64_t jb = tid - lane; jb < nvec; jb += THREADS) {
            const int64_t j = jb + lane;
            const bool inb = j < nvec;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inb) v = __ldg(p4 + j);
            visit(v.x, inb); visit(v.y, inb); visit(v.z, inb); visit(v.w, inb);
        }
        for (int64_t jb = begin + (nvec << 2); jb < end; jb += 32) {  // < 4 leftover elements
            if (tid < 32) {
                const int64_t j = jb + lane;
                const bool inb = j < end;
                visit(inb ? __ldg(base + j) : 0.0f, inb);
            }
        }
    } else {
        for (int64_t jb = begin + tid - lane; jb < end; jb += THREADS) {
            const int64_t j = jb + lane;
            const bool inb = j < end;
            visit(inb ? __ldg(base + j) : 0.0f, inb);
        }
    }
    __syncthreads();
    uint32_t *gh = hist + u * kHistWords;
    for (int j = tid; j < nbins; j += THREADS) {
        const uint32_t c = sh[j];
        if (c) atomicAdd(&gh[j], c);
    }
    if (ROUND == 0 && __any_sync(0xffffffffu, has_nan) && lane == 0) atomicAdd(&gh[kHistBins], 1u);
    if (ROUND == 2) {
        mn = __reduce_min_sync(0xffffffffu, mn);
        if (lane == 0 && mn != 0xffffffffu) atomicMin(&min_above[u], mn);
    }
}

template <int ROUND>
__global__ void __launch_bounds__(256) select_advance_kernel(SelectState *state, const uint32_t *hist) {
    constexpr int THREADS = 256;
    __shared__ uint32_t scratch[kScratchWords];
    const int64_t u = blockIdx.x;
    SelectState st = state[u];
    if (st.mode != kModeThreshold) return;
    const uint32_t *gh = hist + u * kHistWords;
    const BinHit hit = block_find_bin<THREADS>(gh, round_bins(ROUND), st.rank, scratch);
    st.prefix |= hit.bin << round_shift(ROUND);
    st.rank -= hit.below;
    st.below_total += hit.below;
    if (ROUND == 0) st.nan_flag |= gh[kHistBins];
    if (ROUND == 2) {
        st.a_key = st.prefix;
        st.count = hit.count;
        if (st.hi < st.below_total + hit.count) {
            st.b_key = st.a_key;
            st.need_min = 0;
        } else {
            const uint32_t nb = block_next_nonempty<THREADS>(gh, round_bins(2), hit.bin, scratch);
            st.need_min = (nb == 0xffffffffu) ? 1u : 0u;
            st.b_key = (st.prefix & ~round_mask(2)) | (nb & round_mask(2));
        }
    }
    if (threadIdx.x == 0) state[u] = st;
}

__global__ void select_finish_kernel(const SelectState *state, const uint32_t *min_above,
                                     int64_t units, float *thr_out, float *a_out, float *b_out) {
    const int64_t u = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (u >= units) return;
    const SelectState st = state[u];
    float thr, a, b;
    if (st.mode == kModeDone) return;   // written by gs_pivot_kernel
    if (st.mode == kModeOnes) {
        thr = a = b = -INFINITY;
    } else if (st.mode == kModeZeros) {
        thr = a = b = INFINITY;
    } else if (st.nan_flag) {
        thr = a = b = __int_as_float(0x7fc00000);
    } else {
        a = key_to_float(st.a_key);
        b = key_to_float(st.need_min ? min_above[u] : st.b_key);
        thr = quantile_lerp(a, b, st.w);
    }
    if (thr_out) thr_out[u] = thr;
    if (a_out) a_out[u] = a;
    if (b_out) b_out[u] = b;
}

// ------------------------------------------------------------------------------------------
// apply with given thresholds: grid = units * tiles_per_unit
// ------------------------------------------------------------------------------------------
template <bool TRAIN, bool VEC, int OUTS>
__global__ void __launch_bounds__(256) slice_apply_kernel(const SliceParams p, int tiles_per_unit) {
    constexpr int THREADS = 256;
    __shared__ __align__(16) float tbl[kIndexSmemFloats];
    __shared__ double red[THREADS / 32];
    const int tid = threadIdx.x;
    const IndexCtx ic = index_ctx_setup(p, tbl);
    const int64_t u = blockIdx.x / tiles_per_unit;
    const int tile = blockIdx.x - static_cast<int>(u) * tiles_per_unit;
    const float q = p.q01_per_unit ? p.q01_per_unit[u] : p.q01;
    const int mode = unit_mode(q);
    const float thr = (mode == kModeThreshold) ? p.thr_in[u] : ((mode == kModeOnes) ? -INFINITY : INFINITY);
    if (tile == 0 && tid == 0 && p.thr_out) p.thr_out[u] = thr;
    const int64_t begin = static_cast<int64_t>(tile) * kApplyTile;
    const int len = static_cast<int>(min(static_cast<int64_t>(kApplyTile), p.n - begin));
    // multi-quality apply: `repeat` consecutive output units (one per quality level) read the same input unit
    const int64_t u_in = (p.repeat > 1) ? u / p.repeat : u;
    const float acc = apply_range<TRAIN, VEC, THREADS, OUTS>(p, u * p.n + begin, len, nullptr, mode, thr, ic, nullptr,
                                                             u_in * p.n + begin);
    if (p.rate) {
        const double total = block_sum_f64<THREADS>(acc, red);
        if (tid == 0) atomicAdd(&p.rate[u], total);
    }
}

// ------------------------------------------------------------------------------------------
// multi-quality apply (pic_slice_forward_multi).  Every output of an element depends on the quality level only through
// its mask bit m = (std >= thr_level): the slice arithmetic is evaluated ONCE per element for m = 1 ("kept") and once
// for m = 0 ("masked": mean substitution, the bound scale, symbol 0) with the very same routine the per-level kernel
// uses (apply_pair), and each level then only compares, selects and stores -- ~10 instead of 87 lane-instructions
// per output element, which turns the sweep from instruction-bound into store-bound.  One CTA = a 1024-element chunk
// of one input unit x a group of levels.  grid = (chunks_per_unit * groups, input units); eval mode, 128-bit path.
// ------------------------------------------------------------------------------------------
constexpr int kMultiChunk = 1024;
__global__ void __launch_bounds__(256) slice_apply_multi_kernel(const SliceParams p, int levels, int groups) {
    constexpr int THREADS = 256;
    static_assert(kMultiChunk == 4 * THREADS, "one float4 per thread");
    __shared__ __align__(16) float tbl[kIndexSmemFloats];
    __shared__ double red[THREADS / 32];
    const int tid = threadIdx.x;
    const IndexCtx ic = index_ctx_setup(p, tbl);
    const int64_t u_in = blockIdx.y;
    const int chunk = blockIdx.x / groups, grp = blockIdx.x - chunk * groups;
    const int l0 = static_cast<int>(static_cast<int64_t>(levels) * grp / groups);
    const int l1 = static_cast<int>(static_cast<int64_t>(levels) * (grp + 1) / groups);
    const int64_t begin = static_cast<int64_t>(chunk) * kMultiChunk;
    const int nvec = static_cast<int>(min(static_cast<int64_t>(kMultiChunk), p.n - begin)) >> 2;
    const bool live = tid < nvec;
    const int outs = outs_of(p);
    const bool has_base = outs & 1, want_mask = outs & 2, want_yhat = outs & 4;
    const bool want_lik = outs & (8 | 64), store_lik = outs & 8;
    const bool want_idx = outs & 16, want_sym = outs & 32, want_rate = outs & 64;
    const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();
    const int64_t vin = ((u_in * p.n + begin) >> 2) + tid;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // the other level groups of this chunk read the same lines: keep them in L2
    const float4 sv = live ? ld_hint(reinterpret_cast<const float4 *>(p.std) + vin, pol_last) : zero4;
    const float4 ytv = live ? ld_hint(reinterpret_cast<const float4 *>(p.y_top) + vin, pol_last) : zero4;
    const float4 muv = live ? ld_hint(reinterpret_cast<const float4 *>(p.mu) + vin, pol_last) : zero4;
    const float4 ybv = (live && has_base) ? ld_hint(reinterpret_cast<const float4 *>(p.y_base) + vin, pol_last) : zero4;
    const float s[4] = {sv.x, sv.y, sv.z, sv.w}, yt[4] = {ytv.x, ytv.y, ytv.z, ytv.w};
    const float yb[4] = {ybv.x, ybv.y, ybv.z, ybv.w}, mu[4] = {muv.x, muv.y, muv.z, muv.w};
    const float nz[2] = {0.0f, 0.0f};
    PairOut kept[2], gone[2];
    float lg_kept[4] = {0.f, 0.f, 0.f, 0.f}, lg_gone[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        apply_pair<false>(s + 2 * h, yt + 2 * h, yb + 2 * h, mu + 2 * h, nz, has_base, 0.0f, true, p.scale_bound, p.lik_bound,
                          want_lik, want_idx, want_sym, ic, kept[h]);
        apply_pair<false>(s + 2 * h, yt + 2 * h, yb + 2 * h, mu + 2 * h, nz, has_base, __int_as_float(0x7fc00000), false,
                          p.scale_bound, p.lik_bound, want_lik, want_idx, want_sym, ic, gone[h]);
        if (want_rate) {
            lg_kept[2 * h] = logf(kept[h].lik[0]); lg_kept[2 * h + 1] = logf(kept[h].lik[1]);
            lg_gone[2 * h] = logf(gone[h].lik[0]); lg_gone[2 * h + 1] = logf(gone[h].lik[1]);
        }
    }
    for (int l = l0; l < l1; ++l) {
        const int64_t u = u_in * levels + l;
        const int mode = unit_mode(p.q01_per_unit[u]);
        const bool force_one = mode == kModeOnes;
        const float thr = (mode == kModeThreshold) ? p.thr_in[u] : __int_as_float(0x7fc00000);   // s >= NaN is false
        bool m[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) m[e] = (s[e] >= thr) || force_one;
        float rate_acc = 0.0f;
        if (live) {
            const int64_t vi = ((u * p.n + begin) >> 2) + tid;
            auto pick_f = [&](const float (&a)[2], const float (&b)[2], int e) { return m[e] ? a[e & 1] : b[e & 1]; };
            auto pick_i = [&](const int32_t (&a)[2], const int32_t (&b)[2], int e) { return m[e] ? a[e & 1] : b[e & 1]; };
            if (want_mask)
                st_hint(reinterpret_cast<float4 *>(p.mask) + vi,
                        make_float4(m[0] ? 1.f : 0.f, m[1] ? 1.f : 0.f, m[2] ? 1.f : 0.f, m[3] ? 1.f : 0.f), pol_first);
            if (want_yhat)
                st_hint(reinterpret_cast<float4 *>(p.y_hat) + vi,
                        make_float4(pick_f(kept[0].y_hat, gone[0].y_hat, 0), pick_f(kept[0].y_hat, gone[0].y_hat, 1),
                                    pick_f(kept[1].y_hat, gone[1].y_hat, 2), pick_f(kept[1].y_hat, gone[1].y_hat, 3)), pol_first);
            if (store_lik)
                st_hint(reinterpret_cast<float4 *>(p.lik) + vi,
                        make_float4(pick_f(kept[0].lik, gone[0].lik, 0), pick_f(kept[0].lik, gone[0].lik, 1),
                                    pick_f(kept[1].lik, gone[1].lik, 2), pick_f(kept[1].lik, gone[1].lik, 3)), pol_first);
            if (want_idx)
                st_hint(reinterpret_cast<int4 *>(p.idx) + vi,
                        make_int4(pick_i(kept[0].idx, gone[0].idx, 0), pick_i(kept[0].idx, gone[0].idx, 1),
                                  pick_i(kept[1].idx, gone[1].idx, 2), pick_i(kept[1].idx, gone[1].idx, 3)), pol_first);
            if (want_sym)
                st_hint(reinterpret_cast<int4 *>(p.symbols) + vi,
                        make_int4(pick_i(kept[0].sym, gone[0].sym, 0), pick_i(kept[0].sym, gone[0].sym, 1),
                                  pick_i(kept[1].sym, gone[1].sym, 2), pick_i(kept[1].sym, gone[1].sym, 3)), pol_first);
            if (want_rate) {
                // same pairwise order as the per-level kernel: (e0 + e1) of each pair, pairs in order
                rate_acc += (m[0] ? lg_kept[0] : lg_gone[0]) + (m[1] ? lg_kept[1] : lg_gone[1]);
                rate_acc += (m[2] ? lg_kept[2] : lg_gone[2]) + (m[3] ? lg_kept[3] : lg_gone[3]);
            }
        }
        if (want_rate) {
            const double total = block_sum_f64<THREADS>(rate_acc, red);
            if (tid == 0) atomicAdd(&p.rate[u], total);
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward of the slice
// ------------------------------------------------------------------------------------------
struct BwdParams {
    const float *g_lik, *g_yhat, *y_top, *y_base, *mu, *std, *mask, *noise;
    float scale_bound, lik_bound;
    int64_t n;
    float *g_ytop, *g_ybase, *g_mu, *g_std;
};

template <bool TRAIN>
__device__ __forceinline__ void backward_one(const BwdParams &p, float gl, float gy, float yt, float yb,
                                             float mu, float s, float m, float nz, float &g_d,
                                             float &g_mu, float &g_s) {
    const float r = p.y_base ? __fsub_rn(yt, yb) : yt;
    const float d = __fsub_rn(r, mu);
    const float y_m = __fmul_rn(d, m);
    const float s_m = __fmul_rn(s, m);
    const float x = TRAIN ? __fadd_rn(y_m, nz) : rintf(y_m);
    float raw, dv, dsc;
    likelihood_grads(x, s_m, p.scale_bound, raw, dv, dsc);
    // likelihood LowerBound backward: pass iff raw >= bound or grad < 0
    const float glr = (p.lik_bound > 0.0f) ? (((raw >= p.lik_bound) || (gl < 0.0f)) ? gl : 0.0f) : gl;
    const float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
    const float g_x = TRAIN ? glr * dv * sgn : 0.0f;
    const float g_sc = glr * dsc;
    const float g_sm = ((s_m >= p.scale_bound) || (g_sc < 0.0f)) ? g_sc : 0.0f;
    g_d = g_x * m + gy * m;
    g_mu = gy - g_d;
    g_s = g_sm * m;
}

template <bool TRAIN, bool VEC>
__global__ void __launch_bounds__(256) slice_backward_kernel(const BwdParams p) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (VEC) {
        const int64_t nvec = p.n >> 2;
        for (int64_t j = t0; j < nvec; j += stride) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 gl = p.g_lik ? __ldg(reinterpret_cast<const float4 *>(p.g_lik) + j) : z;
            const float4 gy = p.g_yhat ? __ldg(reinterpret_cast<const float4 *>(p.g_yhat) + j) : z;
            const float4 yt = __ldg(reinterpret_cast<const float4 *>(p.y_top) + j);
            const float4 yb = p.y_base ? __ldg(reinterpret_cast<const float4 *>(p.y_base) + j) : z;
            const float4 mu = __ldg(reinterpret_cast<const float4 *>(p.mu) + j);
            const float4 s = __ldg(reinterpret_cast<const float4 *>(p.std) + j);
            const float4 m = __ldg(reinterpret_cast<const float4 *>(p.mask) + j);
            const float4 nz = TRAIN ? __ldg(reinterpret_cast<const float4 *>(p.noise) + j) : z;
            float4 gd, gm, gs;
            backward_one<TRAIN>(p, gl.x, gy.x, yt.x, yb.x, mu.x, s.x, m.x, nz.x, gd.x, gm.x, gs.x);
            backward_one<TRAIN>(p, gl.y, gy.y, yt.y, yb.y, mu.y, s.y, m.y, nz.y, gd.y, gm.y, gs.y);
            backward_one<TRAIN>(p, gl.z, gy.z, yt.z, yb.z, mu.z, s.z, m.z, nz.z, gd.z, gm.z, gs.z);
            backward_one<TRAIN>(p, gl.w, gy.w, yt.w, yb.w, mu.w, s.w, m.w, nz.w, gd.w, gm.w, gs.w);
            if (p.g_ytop) reinterpret_cast<float4 *>(p.g_ytop)[j] = gd;
            if (p.g_ybase) reinterpret_cast<float4 *>(p.g_ybase)[j] = make_float4(-gd.x, -gd.y, -gd.z, -gd.w);
            if (p.g_mu) reinterpret_cast<float4 *>(p.g_mu)[j] = gm;
            if (p.g_std) reinterpret_cast<float4 *>(p.g_std)[j] = gs;
        }
    } else {
        for (int64_t j = t0; j < p.n; j += stride) {
            float gd, gm, gs;
            backward_one<TRAIN>(p, p.g_lik ? p.g_lik[j] : 0.f, p.g_yhat ? p.g_yhat[j] : 0.f, p.y_top[j],
                                p.y_base ? p.y_base[j] : 0.f, p.mu[j], p.std[j], p.mask[j],
                                TRAIN ? p.noise[j] : 0.f, gd, gm, gs);
            if (p.g_ytop) p.g_ytop[j] = gd;
            if (p.g_ybase) p.g_ybase[j] = -gd;
            if (p.g_mu) p.g_mu[j] = gm;
            if (p.g_std) p.g_std[j] = gs;
        }
    }
}

// ------------------------------------------------------------------------------------------
// un-fused operators
// ------------------------------------------------------------------------------------------
// GaussianConditional.forward / _likelihood (entropy_models.py:620-652).
// kind: 0 eval (round), 1 training (noise), 2 likelihood only (outputs := inputs)
template <int KIND>
__global__ void __launch_bounds__(256) gaussian_forward_kernel(const float *inputs, const float *scales,
                                                               const float *means, const float *noise,
                                                               int64_t n, float scale_bound, float lik_bound,
                                                               float *outputs, float *lik, bool vec) {
    const float lb = (KIND == 2) ? 0.0f : lik_bound;
    auto quant = [&](float x, float mu, float nz, float &out, float &value) {
        if (KIND == 1) {
            out = __fadd_rn(x, nz);
        } else if (KIND == 0) {
            float t = means ? __fsub_rn(x, mu) : x;
            t = rintf(t);
            out = means ? __fadd_rn(t, mu) : t;
        } else {
            out = x;
        }
        value = means ? __fsub_rn(out, mu) : out;
    };
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (vec) {
        // 128-bit accesses, two packed-f32 likelihood pairs per float4 (pic_fast.cuh)
        const int64_t nvec = n >> 2;
        for (int64_t j = t0; j < nvec; j += stride) {
            const float4 x4 = __ldg(reinterpret_cast<const float4 *>(inputs) + j);
            const float4 s4 = __ldg(reinterpret_cast<const float4 *>(scales) + j);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 m4 = means ? __ldg(reinterpret_cast<const float4 *>(means) + j) : z;
            const float4 n4 = (KIND == 1) ? __ldg(reinterpret_cast<const float4 *>(noise) + j) : z;
            const float x[4] = {x4.x, x4.y, x4.z, x4.w}, sc[4] = {s4.x, s4.y, s4.z, s4.w};
            const float mu[4] = {m4.x, m4.y, m4.z, m4.w}, nz[4] = {n4.x, n4.y, n4.z, n4.w};
            float o[4], v[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) quant(x[e], mu[e], nz[e], o[e], v[e]);
            if (outputs) reinterpret_cast<float4 *>(outputs)[j] = make_float4(o[0], o[1], o[2], o[3]);
            if (lik) {
                likelihood_pair(fabsf(v[0]), fabsf(v[1]), max_nan(sc[0], scale_bound), max_nan(sc[1], scale_bound), lb, l[0], l[1]);
                likelihood_pair(fabsf(v[2]), fabsf(v[3]), max_nan(sc[2], scale_bound), max_nan(sc[3], scale_bound), lb, l[2], l[3]);
                reinterpret_cast<float4 *>(lik)[j] = make_float4(l[0], l[1], l[2], l[3]);
            }
        }
        return;
    }
    // scalar path: two elements per iteration through the packed-f32 likelihood
    const int64_t pairs = n >> 1;
    for (int64_t j = t0; j < pairs; j += stride) {
        const int64_t i0 = 2 * j, i1 = i0 + 1;
        float o0, o1, v0, v1;
        quant(inputs[i0], means ? means[i0] : 0.0f, (KIND == 1) ? noise[i0] : 0.0f, o0, v0);
        quant(inputs[i1], means ? means[i1] : 0.0f, (KIND == 1) ? noise[i1] : 0.0f, o1, v1);
        if (outputs) { outputs[i0] = o0; outputs[i1] = o1; }
        if (lik) {
            float l0, l1;
            likelihood_pair(fabsf(v0), fabsf(v1), max_nan(scales[i0], scale_bound), max_nan(scales[i1], scale_bound),
                            lb, l0, l1);
            lik[i0] = l0;
            lik[i1] = l1;
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i0 = n - 1;
        float o0, v0, l0, l1;
        quant(inputs[i0], means ? means[i0] : 0.0f, (KIND == 1) ? noise[i0] : 0.0f, o0, v0);
        if (outputs) outputs[i0] = o0;
        if (lik) {
            const float sc = max_nan(scales[i0], scale_bound);
            likelihood_pair(fabsf(v0), fabsf(v0), sc, sc, lb, l0, l1);
            lik[i0] = l0;
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) gaussian_backward_kernel(const float *g_out, const float *g_lik,
                                                                const float *inputs, const float *scales,
                                                                const float *means, const float *noise,
                                                                int64_t n, float scale_bound, float lik_bound,
                                                                float *g_inputs, float *g_scales, float *g_means) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += stride) {
        const float xin = inputs[j];
        const float mu = means ? means[j] : 0.0f;
        float out;
        if (KIND == 1) out = __fadd_rn(xin, noise[j]);
        else if (KIND == 0) {
            float t = means ? __fsub_rn(xin, mu) : xin;
            t = rintf(t);
            out = means ? __fadd_rn(t, mu) : t;
        } else out = xin;
        const float x = means ? __fsub_rn(out, mu) : out;
        const float sraw = scales[j];
        float raw, dv, dsc;
        likelihood_grads(x, sraw, scale_bound, raw, dv, dsc);
        const float gl = g_lik ? g_lik[j] : 0.0f;
        const float lb = (KIND == 2) ? 0.0f : lik_bound;
        const float glr = (lb > 0.0f) ? (((raw >= lb) || (gl < 0.0f)) ? gl : 0.0f) : gl;
        const float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
        const float g_x = glr * dv * sgn;       // grad wrt values = outputs - means
        const float g_sc = glr * dsc;
        const float go = g_out ? g_out[j] : 0.0f;
        const float g_outputs = go + g_x;
        // outputs: noise/identity -> d/dinputs = 1 ; eval -> round() blocks inputs, d/dmeans = 1
        if (g_inputs) g_inputs[j] = (KIND == 0) ? 0.0f : g_outputs;
        if (g_means) g_means[j] = ((KIND == 0) ? g_outputs : 0.0f) - g_x;
        if (g_scales) g_scales[j] = ((sraw >= scale_bound) || (g_sc < 0.0f)) ? g_sc : 0.0f;
    }
}

__global__ void __launch_bounds__(256) build_indexes_kernel(const float *scales, int64_t n, const float *table,
                                                            int table_len, float scale_bound, int32_t *idx, bool vec) {
    __shared__ __align__(16) float sm[kIndexSmemFloats];
    SliceParams p{};
    p.table = table;
    p.table_len = table_len;
    p.idx = idx;
    const IndexCtx ic = index_ctx_setup(p, sm);   // shared table copy + MUFU.LG2 guess parameters
    __syncthreads();
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (vec) {
        const int64_t nvec = n >> 2;
        for (int64_t j = t0; j < nvec; j += stride) {
            const float4 s4 = __ldg(reinterpret_cast<const float4 *>(scales) + j);
            reinterpret_cast<int4 *>(idx)[j] =
                make_int4(scale_index_fast(max_nan(s4.x, scale_bound), ic), scale_index_fast(max_nan(s4.y, scale_bound), ic),
                          scale_index_fast(max_nan(s4.z, scale_bound), ic), scale_index_fast(max_nan(s4.w, scale_bound), ic));
        }
        return;
    }
    for (int64_t j = t0; j < n; j += stride) idx[j] = scale_index_fast(max_nan(scales[j], scale_bound), ic);
}

__device__ __forceinline__ float quantize_one(float x, float mu, float nz, float mk, int mode, bool has_mean,
                                              bool has_mask) {
    if (mode == PIC_QUANTIZE_NOISE) return __fadd_rn(x, has_mask ? __fmul_rn(nz, mk) : nz);
    if (mode == PIC_QUANTIZE_STE) return __fadd_rn(__fsub_rn(rintf(x), x), x);
    const float t = rintf(has_mean ? __fsub_rn(x, mu) : x);
    if (mode == PIC_QUANTIZE_DEQUANTIZE) return has_mean ? __fadd_rn(t, mu) : t;
    return t;  // PIC_QUANTIZE_SYMBOLS: the caller converts to int32
}

// vec: n % 4 == 0 and every pointer 16-byte aligned -> one 128-bit access per tensor and thread
__global__ void __launch_bounds__(256) quantize_kernel(const float *inputs, const float *means, const float *noise,
                                                       const float *mask, int64_t n, int mode, float *out_f,
                                                       int32_t *out_i, bool vec) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool has_mean = means != nullptr, has_mask = mask != nullptr, is_noise = mode == PIC_QUANTIZE_NOISE;
    if (vec) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t j = t0; j < (n >> 2); j += stride) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(inputs) + j);
            const float4 m = (has_mean && !is_noise) ? __ldg(reinterpret_cast<const float4 *>(means) + j) : z;
            const float4 nz = is_noise ? __ldg(reinterpret_cast<const float4 *>(noise) + j) : z;
            const float4 mk = (is_noise && has_mask) ? __ldg(reinterpret_cast<const float4 *>(mask) + j) : z;
            const float4 r = make_float4(quantize_one(x.x, m.x, nz.x, mk.x, mode, has_mean, has_mask),
                                         quantize_one(x.y, m.y, nz.y, mk.y, mode, has_mean, has_mask),
                                         quantize_one(x.z, m.z, nz.z, mk.z, mode, has_mean, has_mask),
                                         quantize_one(x.w, m.w, nz.w, mk.w, mode, has_mean, has_mask));
            if (mode == PIC_QUANTIZE_SYMBOLS)
                reinterpret_cast<int4 *>(out_i)[j] =
                    make_int4(__float2int_rn(r.x), __float2int_rn(r.y), __float2int_rn(r.z), __float2int_rn(r.w));
            else
                reinterpret_cast<float4 *>(out_f)[j] = r;
        }
        return;
    }
    for (int64_t j = t0; j < n; j += stride) {
        const float r = quantize_one(inputs[j], (has_mean && !is_noise) ? means[j] : 0.0f, is_noise ? noise[j] : 0.0f,
                                     (is_noise && has_mask) ? mask[j] : 0.0f, mode, has_mean, has_mask);
        if (mode == PIC_QUANTIZE_SYMBOLS) out_i[j] = __float2int_rn(r);
        else out_f[j] = r;
    }
}

__global__ void __launch_bounds__(256) dequantize_kernel(const int32_t *symbols, const float *means, int64_t n,
                                                         float *out, bool vec) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (vec) {
        for (int64_t j = t0; j < (n >> 2); j += stride) {
            const int4 q = __ldg(reinterpret_cast<const int4 *>(symbols) + j);
            float4 v = make_float4(static_cast<float>(q.x), static_cast<float>(q.y), static_cast<float>(q.z),
                                   static_cast<float>(q.w));
            if (means) {
                const float4 m = __ldg(reinterpret_cast<const float4 *>(means) + j);
                v = make_float4(__fadd_rn(v.x, m.x), __fadd_rn(v.y, m.y), __fadd_rn(v.z, m.z), __fadd_rn(v.w, m.w));
            }
            reinterpret_cast<float4 *>(out)[j] = v;
        }
        return;
    }
    for (int64_t j = t0; j < n; j += stride) {
        const float v = static_cast<float>(symbols[j]);
        out[j] = means ? __fadd_rn(v, means[j]) : v;
    }
}

// vec: n_per_unit % 4 == 0 and aligned pointers, so a float4 never straddles two units.
// copies > 1: output layout [units][copies][n_per_unit] -- the REM attention mask cat([m, m], 1)
// (models/rem_pic.py:181-195) written once instead of mask + two copies.
__global__ void __launch_bounds__(256) mask_from_threshold_kernel(const float *std, const float *thr,
                                                                  int64_t n_per_unit, int64_t total,
                                                                  float *mask, bool vec, int copies, float q01,
                                                                  const float *q01_per_unit) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (vec) {
        for (int64_t j = t0; j < (total >> 2); j += stride) {
            const int64_t u = (4 * j) / n_per_unit;
            const int mode = unit_mode(q01_per_unit ? __ldg(q01_per_unit + u) : q01);   // ones / zeros ignore std
            const float t = __ldg(thr + u);
            const float4 s = __ldg(reinterpret_cast<const float4 *>(std) + j);
            float4 m = make_float4((s.x >= t) ? 1.0f : 0.0f, (s.y >= t) ? 1.0f : 0.0f,
                                   (s.z >= t) ? 1.0f : 0.0f, (s.w >= t) ? 1.0f : 0.0f);
            if (mode != kModeThreshold) m.x = m.y = m.z = m.w = (mode == kModeOnes) ? 1.0f : 0.0f;
            float *dst = mask + 4 * j + u * (copies - 1) * n_per_unit;
            for (int c = 0; c < copies; ++c) *reinterpret_cast<float4 *>(dst + c * n_per_unit) = m;
        }
        return;
    }
    for (int64_t j = t0; j < total; j += stride) {
        const int64_t u = j / n_per_unit;
        const int mode = unit_mode(q01_per_unit ? q01_per_unit[u] : q01);
        const float m = (mode != kModeThreshold) ? ((mode == kModeOnes) ? 1.0f : 0.0f) : ((std[j] >= thr[u]) ? 1.0f : 0.0f);
        float *dst = mask + j + u * (copies - 1) * n_per_unit;
        for (int c = 0; c < copies; ++c) dst[c * n_per_unit] = m;
    }
}

// ------------------------------------------------------------------------------------------
// Elementwise neighbours of the path, one pass each instead of three:
//   LRP epilogue + merge  (models/pic.py:635-641):  out = (y_hat + 0.5 * tanh(lrp)) + base
//   REM merge             (layers/rem.py:137-140):  out = identity + ret * att_mask
// KIND 0 forward LRP, 1 backward LRP (g_lrp = g * 0.5 * (1 - tanh(lrp)^2)), 2 forward REM, 3 backward REM (g * mask)
// ------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ float neighbour_op(float a, float b, float c, bool has_c) {
    if (KIND == 0) {
        const float t = __fadd_rn(a, __fmul_rn(0.5f, tanhf(b)));
        return has_c ? __fadd_rn(t, c) : t;
    }
    if (KIND == 1) {
        const float th = tanhf(b);
        return __fmul_rn(a, __fmul_rn(0.5f, __fsub_rn(1.0f, __fmul_rn(th, th))));
    }
    if (KIND == 2) return __fadd_rn(a, __fmul_rn(b, c));
    return __fmul_rn(a, b);
}

template <int KIND>
__global__ void __launch_bounds__(256) neighbour_kernel(const float *a, const float *b, const float *c, float *out,
                                                        int64_t n, bool vec) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool has_c = c != nullptr;
    if (vec) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t j = t0; j < (n >> 2); j += stride) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(a) + j);
            const float4 y = __ldg(reinterpret_cast<const float4 *>(b) + j);
            const float4 w = has_c ? __ldg(reinterpret_cast<const float4 *>(c) + j) : z;
            reinterpret_cast<float4 *>(out)[j] =
                make_float4(neighbour_op<KIND>(x.x, y.x, w.x, has_c), neighbour_op<KIND>(x.y, y.y, w.y, has_c),
                            neighbour_op<KIND>(x.z, y.z, w.z, has_c), neighbour_op<KIND>(x.w, y.w, w.w, has_c));
        }
        return;
    }
    for (int64_t j = t0; j < n; j += stride) out[j] = neighbour_op<KIND>(a[j], b[j], has_c ? c[j] : 0.0f, has_c);
}

// Progressive level map (test/functions_encode.py:176-190, functions_decode.py:186-200): for thresholds
// thr[u][0..Q) of increasing quality, level[e] = first l with std[e] >= thr[u][l], or Q when no level keeps
// the element.  The delta mask of level l (ProgMask(q_l) - ProgMask(q_{l-1})) is exactly (level == l).
__global__ void __launch_bounds__(256) level_map_kernel(const float *std, const float *thr, int64_t n_per_unit,
                                                        int64_t total, int levels, int32_t *level) {
    extern __shared__ float sthr[];   // levels floats of the CTA's unit (grid-stride loops stay inside one unit)
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    int64_t cached_u = -1;
    for (int64_t j0 = static_cast<int64_t>(blockIdx.x) * blockDim.x; j0 < total; j0 += stride) {
        const int64_t u = j0 / n_per_unit;                       // unit of the CTA's first element
        if (u != cached_u) {
            __syncthreads();
            for (int l = threadIdx.x; l < levels; l += blockDim.x) sthr[l] = thr[u * levels + l];
            __syncthreads();
            cached_u = u;
        }
        const int64_t j = j0 + threadIdx.x;
        if (j >= total) continue;
        const float s = std[j];
        const int64_t uj = j / n_per_unit;
        int lv = levels;
        if (uj == u) {
            for (int l = 0; l < levels; ++l)
                if (s >= sthr[l]) { lv = l; break; }
        } else {                                                  // CTA straddles a unit boundary
            for (int l = 0; l < levels; ++l)
                if (s >= thr[uj * levels + l]) { lv = l; break; }
        }
        level[j] = lv;
    }
}

// per-unit sum ln(x): grid (chunks, units), atomics in f64 (out pre-zeroed)
__global__ void __launch_bounds__(256) log_sum_kernel(const float *x, int64_t n_per_unit, double *out) {
    __shared__ double red[8];
    const int64_t u = blockIdx.y;
    const float *base = x + u * n_per_unit;
    float acc = 0.0f;
    int cnt = 0;
    double dacc = 0.0;
    for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < n_per_unit;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        acc += logf(base[j]);
        if (++cnt == 64) { dacc += acc; acc = 0.0f; cnt = 0; }
    }
    dacc += acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_down_sync(0xffffffffu, dacc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dacc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(&out[u], t);
    }
}

// ------------------------------------------------------------------------------------------
// launch plumbing
// ------------------------------------------------------------------------------------------
template <bool TRAIN, bool VEC, int THREADS, int OUTS>
static int launch_fused_t(const SliceParams &p_in, cudaStream_t stream) {
    auto kern = slice_fused_kernel<TRAIN, VEC, THREADS, OUTS>;
    SliceParams p = p_in;
    p.use_stage = (VEC && p.apply_kind == 2) ? 1 : 0;
    const size_t smem = (OUTS == kOutsSelectOnly)
                            ? kSelectSmemBytes + (wide_select<THREADS, OUTS>() ? wide_park_bytes<THREADS>()
                                                  : (kSelectPipe && THREADS == 256 ? size_t(16384) : 0))
                            : fused_dyn_smem<TRAIN, THREADS>();
    // per instantiation AND per device: the attribute and the occupancy belong to the device that launches
    static bool configured[64] = {false};
    static int occ_blocks[64][2] = {{0}};
    int dev = 0;
    PIC_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return PIC_ERR_INVALID_ARGUMENT;
    if (!configured[dev]) {
        PIC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured[dev] = true;
    }
    int &occ = occ_blocks[dev][p.use_stage];
    if (occ == 0) {
        int q = 1;
        PIC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, THREADS, smem));
        occ = q < 1 ? 1 : q;
    }
    const int64_t max_grid = static_cast<int64_t>(sm_count()) * occ;
    const int grid = static_cast<int>(p.units < max_grid ? p.units : max_grid);
    kern<<<grid, THREADS, smem, stream>>>(p);
    return launch_status();
}

template <bool TRAIN, bool VEC, int OUTS>
static int launch_fused_v(const SliceParams &p, cudaStream_t stream) {
    // few large units: wider CTAs finish each unit sooner; otherwise 4 x 256-thread CTAs per SM
    // give the best overlap of one unit's (latency-bound) select with other units' apply sweeps
    static const int forced = [] { const char *e = getenv("PIC_FUSED_THREADS"); return e ? atoi(e) : 0; }();
    // select-only with at most one unit per SM (per-slice launches, single images): one 1024-thread CTA per SM
    // with 64 KB of loads in flight sweeps a unit ~3x sooner than 512 threads with 16 KB
    if (OUTS == kOutsSelectOnly && (forced == kWideThreads || (forced == 0 && p.n > 16384 && p.units <= sm_count())))
        return launch_fused_t<false, VEC, kWideThreads, kOutsSelectOnly>(p, stream);
    if (forced == 512 || (forced == 0 && p.n > 16384 && p.units < 2 * static_cast<int64_t>(sm_count())))
        return launch_fused_t<TRAIN, VEC, 512, OUTS>(p, stream);
    return launch_fused_t<TRAIN, VEC, 256, OUTS>(p, stream);
}

static bool slice_vec_ok(const SliceParams &p) {
    if (p.n % 4 != 0) return false;
    const void *ptrs[] = {p.y_top, p.y_base, p.mu, p.std, p.noise, p.mask, p.y_hat, p.lik, p.idx, p.symbols};
    for (const void *q : ptrs)
        if (q && !aligned16(q)) return false;
    return true;
}

// the vectorised kernels index every tensor with one 32-bit float4 index
constexpr int64_t kMaxElemsPerLaunch = int64_t(1) << 33;

template <bool TRAIN>
static int launch_fused_f(const SliceParams &p, bool vec, cudaStream_t stream) {
    if (p.apply_kind == 0 && vec) {
        // select-only launches: measured choice between the three select kernels (scripts/select_tune.py)
        static const int prefer_tma = [] { const char *e = getenv("PIC_PREFER_TMA"); return e ? atoi(e) : 0; }();
        if (prefer_tma && select_tma_usable(p)) return launch_select_tma(p, stream);
        if (select_lean_usable(p)) return launch_select_lean(p, stream);
        if (select_tma_usable(p)) return launch_select_tma(p, stream);
    }
    if (p.apply_kind == 0) return vec ? launch_fused_v<false, true, kOutsSelectOnly>(p, stream)
                                      : launch_fused_v<false, false, kOutsSelectOnly>(p, stream);
    if (!vec) return launch_fused_v<TRAIN, false, kOutsGeneric>(p, stream);
    const int outs = (p.apply_kind == 2) ? outs_of(p) : -2;
    if (outs == kOutsCodec) return launch_fused_v<TRAIN, true, kOutsCodec>(p, stream);
    if (outs == kOutsTrain) return launch_fused_v<TRAIN, true, kOutsTrain>(p, stream);
    return launch_fused_v<TRAIN, true, kOutsGeneric>(p, stream);
}

static SliceParams slice_units(const SliceParams &p, int64_t u0, int64_t count) {
    SliceParams q = p;
    const int64_t e = u0 * p.n;
    q.units = count;
    q.y_top = p.y_top ? p.y_top + e : nullptr;  q.y_base = p.y_base ? p.y_base + e : nullptr;
    q.mu = p.mu ? p.mu + e : nullptr;            q.std = p.std ? p.std + e : nullptr;
    q.noise = p.noise ? p.noise + e : nullptr;   q.mask = p.mask ? p.mask + e : nullptr;
    q.y_hat = p.y_hat ? p.y_hat + e : nullptr;   q.lik = p.lik ? p.lik + e : nullptr;
    q.idx = p.idx ? p.idx + e : nullptr;         q.symbols = p.symbols ? p.symbols + e : nullptr;
    q.q01_per_unit = p.q01_per_unit ? p.q01_per_unit + u0 : nullptr;
    q.thr_in = p.thr_in ? p.thr_in + u0 : nullptr;
    q.thr_out = p.thr_out ? p.thr_out + u0 : nullptr;
    q.a_out = p.a_out ? p.a_out + u0 : nullptr;  q.b_out = p.b_out ? p.b_out + u0 : nullptr;
    q.rate = p.rate ? p.rate + u0 : nullptr;
    return q;
}

static int launch_fused(const SliceParams &p, cudaStream_t stream) {
    const bool vec = slice_vec_ok(p);
    const bool train = p.noise != nullptr && p.apply_kind == 2;
    const int64_t max_units = kMaxElemsPerLaunch / p.n;
    for (int64_t u0 = 0; u0 < p.units; u0 += max_units) {
        const int64_t cnt = (p.units - u0 < max_units) ? (p.units - u0) : max_units;
        const SliceParams q = (u0 == 0 && cnt == p.units) ? p : slice_units(p, u0, cnt);
        const int rc = train ? launch_fused_f<true>(q, vec, stream) : launch_fused_f<false>(q, vec, stream);
        if (rc != PIC_OK) return rc;
    }
    return PIC_OK;
}

template <bool TRAIN>
static int launch_apply_f(const SliceParams &p, bool vec, int grid, int tiles, cudaStream_t stream) {
    if (!vec) {
        slice_apply_kernel<TRAIN, false, kOutsGeneric><<<grid, 256, 0, stream>>>(p, tiles);
    } else {
        const int outs = (p.apply_kind == 2) ? outs_of(p) : -2;
        if (outs == kOutsCodec) slice_apply_kernel<TRAIN, true, kOutsCodec><<<grid, 256, 0, stream>>>(p, tiles);
        else if (outs == kOutsTrain) slice_apply_kernel<TRAIN, true, kOutsTrain><<<grid, 256, 0, stream>>>(p, tiles);
        else slice_apply_kernel<TRAIN, true, kOutsGeneric><<<grid, 256, 0, stream>>>(p, tiles);
    }
    return launch_status();
}

static int launch_apply(const SliceParams &p, cudaStream_t stream) {
    const bool vec = slice_vec_ok(p);
    const bool train = p.noise != nullptr && p.apply_kind == 2;
    const int tiles = static_cast<int>((p.n + kApplyTile - 1) / kApplyTile);
    if (p.rate) PIC_CUDA_CHECK(cudaMemsetAsync(p.rate, 0, sizeof(double) * p.units, stream));
    const int64_t max_units = kMaxElemsPerLaunch / p.n < 1 ? 1 : kMaxElemsPerLaunch / p.n;
    for (int64_t u0 = 0; u0 < p.units; u0 += max_units) {
        const int64_t cnt = (p.units - u0 < max_units) ? (p.units - u0) : max_units;
        const SliceParams q = (u0 == 0 && cnt == p.units) ? p : slice_units(p, u0, cnt);
        const int64_t grid64 = cnt * tiles;
        if (grid64 > 0x7fffffffLL) return PIC_ERR_TOO_LARGE;
        const int rc = train ? launch_apply_f<true>(q, vec, static_cast<int>(grid64), tiles, stream)
                             : launch_apply_f<false>(q, vec, static_cast<int>(grid64), tiles, stream);
        if (rc != PIC_OK) return rc;
    }
    return PIC_OK;
}

// global sampled select (3 launches).  ws layout: [GsUnit x units][cand kCandMax x units][thr x units]
static size_t gs_ws_bytes(int64_t units) {
    return static_cast<size_t>(units) * (sizeof(GsUnit) + kCandMax * sizeof(uint32_t) + sizeof(float)) + 256;
}
static float *gs_thr_buffer(void *ws, int64_t units) {
    unsigned char *b = static_cast<unsigned char *>(ws);
    return reinterpret_cast<float *>(b + static_cast<size_t>(units) * (sizeof(GsUnit) + kCandMax * sizeof(uint32_t)));
}
static int select_sampled_global(const float *std, int64_t n, int64_t units, float q01, const float *q01_per_unit,
                                 float *thr, float *a_out, float *b_out, void *ws, cudaStream_t stream) {
    GsParams g{};
    g.std = std; g.q01_per_unit = q01_per_unit; g.q01 = q01; g.n = n; g.units = units;
    g.st = static_cast<GsUnit *>(ws);
    g.cand = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(ws) + static_cast<size_t>(units) * sizeof(GsUnit));
    g.thr = thr; g.a_out = a_out; g.b_out = b_out;
    g.vec = ((n % 4 == 0) && aligned16(std)) ? 1 : 0;
    g.cand_cap = kCandMax;
    const int tiles = static_cast<int>((n + kGsTile - 1) / kGsTile);
    if (units * tiles > 0x7fffffffLL) return PIC_ERR_TOO_LARGE;
    gs_pivot_kernel<<<static_cast<unsigned>(units), kGsThreads, 0, stream>>>(g);
    if (n > kCandMax) {
        if (g.vec) gs_sweep_kernel<true, false><<<static_cast<unsigned>(units * tiles), kGsThreads, 0, stream>>>(g, tiles);
        else gs_sweep_kernel<false, false><<<static_cast<unsigned>(units * tiles), kGsThreads, 0, stream>>>(g, tiles);
        gs_finish_kernel<<<static_cast<unsigned>(units), kGsThreads, 0, stream>>>(g);
    }
    return launch_status();
}

// workspace layout of the multi-launch select: [state][hist x3 rounds][min_above]
struct RoundsWs {
    SelectState *state;
    uint32_t *hist[3];
    uint32_t *min_above;
    float *thr;
};

static size_t rounds_ws_bytes(int64_t units) {
    size_t b = 0;
    b += static_cast<size_t>(units) * sizeof(SelectState);
    b += static_cast<size_t>(units) * kHistWords * 4 * 3;
    b += ((static_cast<size_t>(units) * 4 + 63) / 64) * 64;  // min_above
    b += ((static_cast<size_t>(units) * 4 + 63) / 64) * 64;  // thr
    return b;
}

static RoundsWs carve_ws(void *ws, int64_t units) {
    RoundsWs r;
    unsigned char *p = static_cast<unsigned char *>(ws);
    r.state = reinterpret_cast<SelectState *>(p);
    p += static_cast<size_t>(units) * sizeof(SelectState);
    for (int i = 0; i < 3; ++i) {
        r.hist[i] = reinterpret_cast<uint32_t *>(p);
        p += static_cast<size_t>(units) * kHistWords * 4;
    }
    r.min_above = reinterpret_cast<uint32_t *>(p);
    p += ((static_cast<size_t>(units) * 4 + 63) / 64) * 64;
    r.thr = reinterpret_cast<float *>(p);
    return r;
}

static int launch_hist_round(const float *std, int64_t n_local, int64_t units, int round,
                             const SelectState *state, uint32_t *hist, uint32_t *min_above,
                             cudaStream_t stream, const float *cand = nullptr, int64_t cand_cap = 0) {
    if (units > 65535) return PIC_ERR_TOO_LARGE;
    PIC_CUDA_CHECK(cudaMemsetAsync(hist, 0, static_cast<size_t>(units) * kHistWords * 4, stream));
    if (round == 2) PIC_CUDA_CHECK(cudaMemsetAsync(min_above, 0xff, static_cast<size_t>(units) * 4, stream));
    if (n_local == 0) return PIC_OK;
    const bool vec = (n_local % 4 == 0) && aligned16(std);
    dim3 grid(static_cast<unsigned>((n_local + kRoundChunk - 1) / kRoundChunk), static_cast<unsigned>(units));
#define PIC_LAUNCH_ROUND(R)                                                                                  \
    if (vec) hist_round_kernel<R, true><<<grid, 512, 0, stream>>>(std, n_local, state, hist, min_above, cand, cand_cap); \
    else hist_round_kernel<R, false><<<grid, 512, 0, stream>>>(std, n_local, state, hist, min_above, cand, cand_cap)
    if (round == 0) { PIC_LAUNCH_ROUND(0); }
    else if (round == 1) { PIC_LAUNCH_ROUND(1); }
    else { PIC_LAUNCH_ROUND(2); }
#undef PIC_LAUNCH_ROUND
    return launch_status();
}

static int launch_advance(SelectState *state, const uint32_t *hist, int64_t units, int round,
                          cudaStream_t stream) {
    const unsigned grid = static_cast<unsigned>(units);
    if (round == 0) select_advance_kernel<0><<<grid, 256, 0, stream>>>(state, hist);
    else if (round == 1) select_advance_kernel<1><<<grid, 256, 0, stream>>>(state, hist);
    else select_advance_kernel<2><<<grid, 256, 0, stream>>>(state, hist);
    return launch_status();
}

// full multi-launch select on one device
static int select_rounds(const float *std, int64_t n, int64_t units, float q01, const float *q01_per_unit,
                         float *thr_out, float *a_out, float *b_out, void *ws, cudaStream_t stream) {
    RoundsWs w = carve_ws(ws, units);
    select_begin_kernel<<<static_cast<unsigned>((units + 127) / 128), 128, 0, stream>>>(w.state, n, units, q01, q01_per_unit);
    int rc = launch_status();
    for (int r = 0; r < 3 && rc == PIC_OK; ++r) {
        rc = launch_hist_round(std, n, units, r, w.state, w.hist[r], w.min_above, stream);
        if (rc == PIC_OK) rc = launch_advance(w.state, w.hist[r], units, r, stream);
    }
    if (rc != PIC_OK) return rc;
    select_finish_kernel<<<static_cast<unsigned>((units + 127) / 128), 128, 0, stream>>>(w.state, w.min_above, units,
                                                                                      thr_out, a_out, b_out);
    return launch_status();
}

// debugging switches of the large-unit select (defaults: sampled sweep + cluster select)
static int env_large_sampled() {
    static const int v = [] { const char *e = getenv("PIC_LARGE_SAMPLED"); return e ? atoi(e) : 1; }();
    return v;
}
static int env_cluster_select() {
    static const int v = [] { const char *e = getenv("PIC_CLUSTER_SELECT"); return e ? atoi(e) : 1; }();
    return v;
}

// Large units (> kFusedMaxElems) on one device: sampled pivots -> one tile-ordered sweep that counts the
// elements below the bracket and compacts the bracket (~6 % of the unit) into a candidate buffer -> the three
// histogram rounds over the candidates only.  Units whose bracket missed or overflowed run the same rounds over
// the whole unit (exact either way).  ws layout: [rounds workspace][GsUnit x units][cand_cap floats x units][below counts x tiles x units].
static int64_t large_cand_cap(int64_t n) { return ((n >> 3) + 3) & ~int64_t(3); }
static size_t large_ws_bytes(int64_t n, int64_t units) {
    const size_t rounds = (rounds_ws_bytes(units) + 255) / 256 * 256;
    const size_t tiles = static_cast<size_t>((n + kGsTile - 1) / kGsTile);
    return rounds + static_cast<size_t>(units) * (sizeof(GsUnit) + (static_cast<size_t>(large_cand_cap(n)) + tiles) * 4) + 512;
}
static int select_large(const float *std, int64_t n, int64_t units, float q01, const float *q01_per_unit,
                        float *thr_out, float *a_out, float *b_out, void *ws, cudaStream_t stream) {
    RoundsWs w = carve_ws(ws, units);
    unsigned char *extra = static_cast<unsigned char *>(ws) + (rounds_ws_bytes(units) + 255) / 256 * 256;
    GsParams g{};
    g.std = std; g.q01_per_unit = q01_per_unit; g.q01 = q01; g.n = n; g.units = units;
    g.st = reinterpret_cast<GsUnit *>(extra);
    g.cand = reinterpret_cast<uint32_t *>(extra + (static_cast<size_t>(units) * sizeof(GsUnit) + 255) / 256 * 256);
    g.thr = thr_out; g.a_out = a_out; g.b_out = b_out;
    g.vec = ((n % 4 == 0) && aligned16(std)) ? 1 : 0;
    g.cand_cap = large_cand_cap(n);
    const int tiles = static_cast<int>((n + kGsTile - 1) / kGsTile);
    g.below_tile = g.cand + static_cast<size_t>(units) * static_cast<size_t>(g.cand_cap);
    if (units * tiles > 0x7fffffffLL) return PIC_ERR_TOO_LARGE;
    gs_pivot_kernel<<<static_cast<unsigned>(units), kGsThreads, 0, stream>>>(g);
    if (g.vec) gs_sweep_kernel<true, true><<<static_cast<unsigned>(units * tiles), kGsThreads, 0, stream>>>(g, tiles);
    else gs_sweep_kernel<false, true><<<static_cast<unsigned>(units * tiles), kGsThreads, 0, stream>>>(g, tiles);
    if (env_cluster_select() && units <= 65535) {
        const dim3 grid(kClusterCtas, static_cast<unsigned>(units));
        if (g.vec) gs_cluster_select_kernel<true><<<grid, kClusterThreads, 0, stream>>>(g, tiles);
        else gs_cluster_select_kernel<false><<<grid, kClusterThreads, 0, stream>>>(g, tiles);
        return launch_status();
    }
    gs_begin_rounds_kernel<<<static_cast<unsigned>(units), 128, 0, stream>>>(g, w.state, tiles);
    int rc = launch_status();
    for (int r = 0; r < 3 && rc == PIC_OK; ++r) {
        rc = launch_hist_round(std, n, units, r, w.state, w.hist[r], w.min_above, stream,
                               reinterpret_cast<const float *>(g.cand), g.cand_cap);
        if (rc == PIC_OK) rc = launch_advance(w.state, w.hist[r], units, r, stream);
    }
    if (rc != PIC_OK) return rc;
    select_finish_kernel<<<static_cast<unsigned>((units + 127) / 128), 128, 0, stream>>>(w.state, w.min_above, units,
                                                                                      thr_out, a_out, b_out);
    return launch_status();
}
// picks the sampled select when the workspace allows it (pic_workspace_bytes), else the plain rounds
static int select_large_or_rounds(const float *std, int64_t n, int64_t units, float q01, const float *q01_per_unit,
                                  float *thr_out, float *a_out, float *b_out, void *ws, size_t ws_bytes,
                                  cudaStream_t stream) {
    if (env_large_sampled() && ws_bytes >= large_ws_bytes(n, units))
        return select_large(std, n, units, q01, q01_per_unit, thr_out, a_out, b_out, ws, stream);
    return select_rounds(std, n, units, q01, q01_per_unit, thr_out, a_out, b_out, ws, stream);
}

// ------------------------------------------------------------------------------------------
// Spatially tiled units, sampled protocol (two collectives): kernels
// ------------------------------------------------------------------------------------------
// K1: s_slot hashed-stride samples of every unit's local band -> out[units][s_slot]
struct TilePivotRanks {
    int klo, khi, S;
};
__device__ __forceinline__ TilePivotRanks tile_pivot_ranks(float q, int64_t n_total, int S) {
    uint32_t lo, hi;
    float w;
    quantile_ranks(q, n_total, lo, hi, w);
    const float frac = static_cast<float>(lo) / static_cast<float>(n_total > 1 ? n_total - 1 : 1);
    const float kt = frac * static_cast<float>(S - 1);
    const float margin = 4.0f * sqrtf(static_cast<float>(S) * frac * (1.0f - frac)) + 4.0f;
    TilePivotRanks r;
    r.klo = static_cast<int>(floorf(kt - margin));
    r.khi = static_cast<int>(ceilf(kt + margin));
    r.S = S;
    return r;
}

// sample i of unit u's band: hashed stride over the band; a band shorter than the slot is sampled with repetition
__device__ __forceinline__ float tile_sample_value(const float *std_local, int64_t n_local, int64_t u, int s_slot, int i) {
    const uint64_t stride = static_cast<uint64_t>(n_local) >= static_cast<uint64_t>(s_slot) ? static_cast<uint64_t>(n_local) / s_slot : 1u;
    const uint32_t jit = static_cast<uint32_t>((static_cast<uint64_t>(static_cast<uint32_t>(i) * 0x9E3779B1u) * stride) >> 32);
    const uint64_t at = (static_cast<uint64_t>(i) * stride + jit) % static_cast<uint64_t>(n_local);
    return __ldg(std_local + u * n_local + at);
}

// also: the two ranks of the pooled sample (world * s_slot values per unit) whose elements become the bracket pivots
// (0xffffffff: open end / no select needed)
__global__ void __launch_bounds__(256) tile_sample_kernel(const float *std_local, int64_t n_local, int64_t units, int s_slot,
                                                          float *out, uint32_t *rank_in, int world, int64_t n_total, float q01,
                                                          const float *q01_per_unit) {
    const int64_t u = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i == 0) {
        const float q = q01_per_unit ? q01_per_unit[u] : q01;
        const int S = world * s_slot;
        uint32_t rl = 0xffffffffu, rh = 0xffffffffu;
        if (unit_mode(q) == kModeThreshold) {
            const TilePivotRanks t = tile_pivot_ranks(q, n_total, S);
            if (t.klo > 0) rl = static_cast<uint32_t>(t.klo);
            if (t.khi < S - 1) rh = static_cast<uint32_t>(t.khi);
        }
        rank_in[2 * u] = rl;
        rank_in[2 * u + 1] = rh;
    }
    if (i >= s_slot) return;
    out[u * s_slot + i] = tile_sample_value(std_local, n_local, u, s_slot, i);
}

// K4: pivots -> GsUnit (sweep state); units that need no select are answered here
__global__ void tile_pivots_kernel(const float *piv /*[2 * units]*/, GsUnit *st, int64_t units, int64_t n_total, int S, float q01,
                                   const float *q01_per_unit, float *thr_out) {
    const int64_t u = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (u >= units) return;
    const float q = q01_per_unit ? q01_per_unit[u] : q01;
    const int mode = unit_mode(q);
    GsUnit g{};
    if (mode != kModeThreshold) {
        g.state = 1;
        thr_out[u] = (mode == kModeOnes) ? -INFINITY : INFINITY;
    } else {
        const TilePivotRanks t = tile_pivot_ranks(q, n_total, S);
        g.plo_f = (t.klo > 0) ? piv[2 * u] : -INFINITY;
        g.phi_f = (t.khi < S - 1) ? piv[2 * u + 1] : INFINITY;
        if (g.plo_f != g.plo_f || g.phi_f != g.phi_f) {     // a NaN in the sample: the unit's threshold is NaN
            g.nan_flag = 1;
            g.plo_f = -INFINITY;
            g.phi_f = -INFINITY;                             // empty bracket: nothing to exchange
        }
        g.state = 0;
    }
    st[u] = g;
}

// K6: header of this rank's exchange slot: [c_below, c_cand, nan, overflow], candidates follow (written by the sweep)
__global__ void __launch_bounds__(128) tile_pack_kernel(const GsUnit *st, const uint32_t *below_tile, int tiles, uint32_t *send,
                                                        int64_t stride, uint32_t cap_x) {
    __shared__ uint32_t red[4];
    const int64_t u = blockIdx.x;
    const int tid = threadIdx.x;
    uint32_t acc = 0;
    for (int t = tid; t < tiles; t += 128) acc += below_tile[u * tiles + t];
    acc = __reduce_add_sync(0xffffffffu, acc);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid != 0) return;
    const GsUnit g = st[u];
    uint32_t *h = send + u * stride;
    const bool active = g.state == 0u;
    h[0] = active ? red[0] + red[1] + red[2] + red[3] : 0u;
    h[1] = active ? (g.c_cand < cap_x ? g.c_cand : cap_x) : 0u;
    h[2] = g.nan_flag;
    h[3] = (active && g.c_cand > cap_x) ? 1u : 0u;
}

// K7: per unit, sum the ranks' headers and line their candidates up in one buffer (`split` CTAs per rank segment); a
// unit whose bracket does not hold both order statistics (or whose slot overflowed) is counted in *invalid and left out
// of the final select.  all_x: [units][world][stride].  Only the headers are read (never st[u], which CTA (0, u) updates).
__global__ void __launch_bounds__(256) tile_merge_kernel(const uint32_t *all_x, int world, int split, int64_t units, int64_t stride,
                                                         GsUnit *st, uint32_t *cand_all, int64_t cap_all, uint32_t cap_x, int64_t n_total,
                                                         float q01, const float *q01_per_unit, uint32_t *invalid, float *thr_out) {
    __shared__ uint32_t sh[2];
    const int64_t u = blockIdx.y;
    const int r = blockIdx.x / split, part = blockIdx.x - r * split, tid = threadIdx.x;
    const float q = q01_per_unit ? q01_per_unit[u] : q01;
    if (unit_mode(q) != kModeThreshold) return;     // ones / zeros: answered by tile_pivots_kernel
    if (tid == 0) {
        uint32_t below = 0, run = 0, nan = 0, ovf = 0, mine = 0;
        for (int k = 0; k < world; ++k) {
            const uint32_t *h = all_x + (u * world + k) * stride;
            if (k == r) mine = run;
            below += h[0];
            run += h[1] < cap_x ? h[1] : cap_x;      // a header that never arrived (exchange time-out) must not index past the slot
            nan |= h[2];
            ovf |= h[3] | (h[1] > cap_x ? 1u : 0u);
        }
        uint32_t lo, hi;
        float w;
        quantile_ranks(q, n_total, lo, hi, w);
        const bool valid = ovf == 0u && below <= lo && hi < below + run;
        sh[0] = (valid && nan == 0u) ? 1u : 0u;
        sh[1] = mine;
        if (blockIdx.x == 0) {
            if (nan != 0u) thr_out[u] = __int_as_float(0x7fc00000);   // any NaN in the unit: NaN threshold (torch.quantile)
            else if (!valid) atomicAdd(invalid, 1u);                  // the caller falls back to the histogram rounds
            GsUnit *g = st + u;
            g->c_below = below;
            g->c_cand = run;
            g->nan_flag = nan;
            g->state = (nan != 0u || !valid) ? 1u : 0u;               // 1: nothing (left) to select in the cluster select
        }
    }
    __syncthreads();
    if (sh[0] == 0u) return;
    const uint32_t *src = all_x + (u * world + r) * stride;
    const uint32_t cnt = src[1] < cap_x ? src[1] : cap_x;
    const uint32_t i0 = static_cast<uint32_t>(static_cast<uint64_t>(cnt) * part / split);
    const uint32_t i1 = static_cast<uint32_t>(static_cast<uint64_t>(cnt) * (part + 1) / split);
    uint32_t *dst = cand_all + u * cap_all + sh[1];
    for (uint32_t i = i0 + tid; i < i1; i += 256) dst[i] = src[4 + i];
}

// ---- peer-memory exchange (NVLink / NVSwitch stores into the peers' windows) --------------------------------------------
// Every rank owns one window (cudaMalloc + CUDA IPC, mapped by all peers).  Layout: [flags 2 x 64 u32][local words]
// [region 0 | region 1].  An exchange = every rank stores its rows into the same place of EVERY window (its own
// included), fences, then raises flags[region][rank] = epoch in every window and waits until all of its own flags reached
// the epoch.  Two regions alternate within one tiled select, which is what makes reuse safe without double buffering:
// a rank can only start writing region R of call k+1 after it saw the peers' flags of the exchange in between, and a
// peer raises those only after (in stream order) it finished reading region R of call k.
constexpr int kP2pMaxWorld = 64;
constexpr size_t kP2pFlagsOff = 0, kP2pLocalOff = 1024, kP2pDataOff = 4096;
__device__ unsigned long long g_p2p_timeout_ns = 4000000000ull;   // bounded wait for the peers' flags (PIC_P2P_TIMEOUT_MS)
struct P2pLocal {
    uint32_t done[2], epoch[2], error, pad[3];
};
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Every CTA of an exchanging grid calls p2p_ticket after its stores (all threads); it returns true in the last CTA to
// arrive, which then (after any last-minute stores of its own) calls p2p_signal_wait with all of its threads.
__device__ __forceinline__ bool p2p_ticket(unsigned char *const *windows, int rank, int region, uint32_t total_ctas, uint32_t *sh_last) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // bar.sync ordered the CTA's stores before this thread; its system-scope fence is cumulative over them
        __threadfence_system();
        P2pLocal *loc = reinterpret_cast<P2pLocal *>(windows[rank] + kP2pLocalOff);
        *sh_last = (atomicAdd(&loc->done[region], 1u) == total_ctas - 1u) ? 1u : 0u;
    }
    __syncthreads();
    return *sh_last != 0u;
}
__device__ __forceinline__ void p2p_signal_wait(unsigned char *const *windows, int rank, int world, int region, uint32_t *status) {
    const int tid = threadIdx.x;
    P2pLocal *loc = reinterpret_cast<P2pLocal *>(windows[rank] + kP2pLocalOff);
    __syncthreads();
    if (tid == 0) __threadfence_system();
    __syncthreads();
    const uint32_t e = loc->epoch[region] + 1u;
    if (tid < world) {
        st_release_sys(reinterpret_cast<uint32_t *>(windows[tid] + kP2pFlagsOff) + region * kP2pMaxWorld + rank, e);
        const uint32_t *mine = reinterpret_cast<const uint32_t *>(windows[rank] + kP2pFlagsOff) + region * kP2pMaxWorld + tid;
        const uint64_t t0 = global_ns();
        while (static_cast<int32_t>(ld_acquire_sys(mine) - e) < 0) {
            __nanosleep(100);
            if (global_ns() - t0 > g_p2p_timeout_ns) {    // a peer never arrived (default 4 s): report instead of hanging the GPU
                atomicAdd(&loc->error, 1u);
                if (status) atomicAdd(status, 0x10000u);
                break;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        loc->done[region] = 0u;
        loc->epoch[region] = e;
    }
}

// rows: `units` rows of this rank; row u = src[u * src_stride ..], `fixed_words` words long, or (fixed_words < 0) a
// 4-word header + src[u * src_stride + 1] payload words.  Lands at window[dst_off + (u * world + rank) * dst_stride].
// grid (chunks, units); the last CTA to finish signals and waits.
// Pack mode (st != nullptr): the row's 4-word header [c_below, c_cand, nan, overflow] is computed here from the sweep's
// per-tile counts instead of being read from src (tile_pack_kernel folded into the exchange).
struct P2pPack {
    const GsUnit *st;
    const uint32_t *below_tile;
    int tiles;
    uint32_t cap_x;
};
__global__ void __launch_bounds__(256) p2p_exchange_kernel(unsigned char *const *windows, int rank, int world, int region,
                                                            const uint32_t *src, int64_t src_stride, int fixed_words,
                                                            size_t dst_off, int64_t dst_stride, uint32_t *status, P2pPack pk) {
    __shared__ uint32_t sh_last;
    __shared__ uint32_t red[8];
    __shared__ uint4 sh_head;
    const int64_t u = blockIdx.y;
    const int tid = threadIdx.x;
    const uint32_t *row = src + u * src_stride;
    int64_t words;
    if (pk.st) {
        uint32_t acc = 0;
        for (int t = tid; t < pk.tiles; t += 256) acc += pk.below_tile[u * pk.tiles + t];
        acc = __reduce_add_sync(0xffffffffu, acc);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            const GsUnit g = pk.st[u];
            const bool active = g.state == 0u;
            uint32_t below = 0;
            for (int k = 0; k < 8; ++k) below += red[k];
            sh_head = make_uint4(active ? below : 0u, active ? (g.c_cand < pk.cap_x ? g.c_cand : pk.cap_x) : 0u, g.nan_flag,
                                 (active && g.c_cand > pk.cap_x) ? 1u : 0u);
        }
        __syncthreads();
        words = 4 + static_cast<int64_t>(sh_head.y);
    } else {
        words = fixed_words >= 0 ? fixed_words : 4 + static_cast<int64_t>(row[1]);
    }
    if (words > dst_stride) words = dst_stride;
    const int64_t vecs = (words + 3) >> 2;     // rows are 16-byte aligned and padded to whole vectors on both sides
    const int64_t v0 = vecs * blockIdx.x / gridDim.x, v1 = vecs * (blockIdx.x + 1) / gridDim.x;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(row);
    for (int k = 0; k < world; ++k) {
        int p = rank + k;                       // start with the own window, every rank walks the peers in a different order
        if (p >= world) p -= world;
        uint4 *d4 = reinterpret_cast<uint4 *>(windows[p] + dst_off) + ((u * world + rank) * dst_stride >> 2);
        for (int64_t i = v0 + tid; i < v1; i += 256) d4[i] = (pk.st && i == 0) ? sh_head : __ldg(s4 + i);
    }
    if (p2p_ticket(windows, rank, region, gridDim.x * gridDim.y, &sh_last)) p2p_signal_wait(windows, rank, world, region, status);
}

// ---- first exchange folded into the kernel that produces the data -------------------------------------------------------
// Sample + exchange: every sampled element is stored straight into all windows (coalesced 128-byte rows per warp).
// (The same fusion for the sweep + second exchange -- every tile CTA appending its bracket elements to all windows --
// was built and measured SLOWER: 0.371 vs 0.315 ms per 2-GPU step.  Each of the thousands of tile CTAs then needs a
// system-scope fence before its ticket and the candidates leave as ~64-byte fragments, whereas the dedicated exchange
// kernel moves them as 512-byte rows from 80 CTAs.  So the sweep keeps writing a local slot.)
__global__ void __launch_bounds__(256) tile_sample_exchange_kernel(const float *std_local, int64_t n_local, int64_t units, int s_slot,
                                                                   uint32_t *rank_in, int64_t n_total, float q01,
                                                                   const float *q01_per_unit, unsigned char *const *windows, int rank,
                                                                   int world, size_t dst_off, uint32_t *status) {
    __shared__ uint32_t sh_last;
    const int64_t u = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i == 0) {
        const float q = q01_per_unit ? q01_per_unit[u] : q01;
        const int S = world * s_slot;
        uint32_t rl = 0xffffffffu, rh = 0xffffffffu;
        if (unit_mode(q) == kModeThreshold) {
            const TilePivotRanks t = tile_pivot_ranks(q, n_total, S);
            if (t.klo > 0) rl = static_cast<uint32_t>(t.klo);
            if (t.khi < S - 1) rh = static_cast<uint32_t>(t.khi);
        }
        rank_in[2 * u] = rl;
        rank_in[2 * u + 1] = rh;
    }
    if (i < s_slot) {
        const float v = tile_sample_value(std_local, n_local, u, s_slot, i);
        const int64_t at = (u * world + rank) * s_slot + i;
        for (int k = 0; k < world; ++k) {
            int p = rank + k;
            if (p >= world) p -= world;
            reinterpret_cast<float *>(windows[p] + dst_off)[at] = v;
        }
    }
    if (p2p_ticket(windows, rank, 0, gridDim.x * gridDim.y, &sh_last)) p2p_signal_wait(windows, rank, world, 0, status);
}

static int check_common(int64_t n_per_unit, int64_t units) {
    if (n_per_unit <= 0 || units <= 0) return PIC_ERR_INVALID_ARGUMENT;
    if (n_per_unit > (int64_t(1) << 24)) return PIC_ERR_TOO_LARGE;
    return PIC_OK;
}

static int elementwise_grid(int64_t n, int per_thread = 4) {
    const int64_t blocks = (n + 256LL * per_thread - 1) / (256LL * per_thread);
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    const int64_t g = blocks < cap ? blocks : cap;
    return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace pic

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace pic;

extern "C" {

int pic_version(void) { return 100; }

const char *pic_error_string(int code) {
    switch (code) {
        case PIC_OK: return "ok";
        case PIC_ERR_INVALID_ARGUMENT: return "invalid argument";
        case PIC_ERR_TOO_LARGE: return "quantile() input tensor is too large";
        case PIC_ERR_WORKSPACE: return "workspace too small";
        case PIC_ERR_CUDA: return "CUDA runtime error";
        case PIC_ERR_UNALIGNED: return "pointer is not 4-byte aligned";
        default: return "unknown error";
    }
}

int pic_last_cuda_error(void) { return g_last_cuda_error; }

int64_t pic_fused_max_elems(void) { return kFusedMaxElems; }

int pic_slice_forward_plan(int64_t n_per_unit, int64_t units, int needs_select, int *n_kernels) {
    if (n_per_unit <= 0 || units <= 0 || !n_kernels) return PIC_ERR_INVALID_ARGUMENT;
    static const int two_kernel = [] { const char *e = getenv("PIC_TWO_KERNEL"); return e ? atoi(e) : 1; }();
    static const int gsel = [] { const char *e = getenv("PIC_GLOBAL_SELECT"); return e ? atoi(e) : 0; }();
    if (n_per_unit > kFusedMaxElems) {
        // with the full workspace: pivot + sweep + cluster select + apply; plain rounds: begin + 3 x (hist,
        // advance) + finish + apply
        *n_kernels = !needs_select ? 1 : (!env_large_sampled() ? 9 : (env_cluster_select() && units <= 65535 ? 4 : 11));
        return 2;
    }
    if (two_kernel && (!needs_select || n_per_unit >= kTwoKernelMinUnit)) {
        *n_kernels = needs_select ? (gsel ? 4 : 2) : 1;   // select kernel(s) + tile-ordered apply
        return 1;
    }
    *n_kernels = 1;                              // single fused kernel
    return 0;
}

#ifdef PIC_PHASE_TIMING
extern "C" int pic_debug_phase_clocks(long long *out) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, g_phase_clk, sizeof(long long) * 16) == cudaSuccess ? 0 : -4;
}
#endif

int pic_debug_select_counters(unsigned long long *sampled, unsigned long long *fallback) {
    PIC_CUDA_CHECK(cudaDeviceSynchronize());
    PIC_CUDA_CHECK(cudaMemcpyFromSymbol(sampled, g_sampled_units, sizeof(unsigned long long)));
    PIC_CUDA_CHECK(cudaMemcpyFromSymbol(fallback, g_fallback_units, sizeof(unsigned long long)));
    select_tma_counters(sampled, fallback);
    return PIC_OK;
}

size_t pic_workspace_bytes(int64_t n_per_unit, int64_t units) {
    if (units <= 0) return 256;
    if (n_per_unit <= kFusedMaxElems) return gs_ws_bytes(units);  // pivots, candidate buffers, thresholds
    return large_ws_bytes(n_per_unit, units);  // rounds state + pivots + candidate buffers (n/8 per unit)
}

size_t pic_select_state_bytes(int64_t units) { return static_cast<size_t>(units < 1 ? 1 : units) * sizeof(SelectState); }
int64_t pic_hist_words(void) { return kHistWords; }

int pic_select_threshold(const float *std, int64_t n_per_unit, int64_t units, float q01,
                         const float *q01_per_unit, float *thr_out, float *a_out, float *b_out,
                         void *ws, size_t ws_bytes, pic_stream_t stream_) {
    int rc = check_common(n_per_unit, units);
    if (rc != PIC_OK) return rc;
    if (!std || !thr_out) return PIC_ERR_INVALID_ARGUMENT;
    if (!aligned4(std)) return PIC_ERR_UNALIGNED;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_per_unit <= kFusedMaxElems) {
        static const int gsel = [] { const char *e = getenv("PIC_GLOBAL_SELECT"); return e ? atoi(e) : 0; }();
        if (gsel && ws && ws_bytes >= gs_ws_bytes(units) && n_per_unit * units >= kTwoKernelMinElems &&
            n_per_unit >= kTwoKernelMinUnit)
            return select_sampled_global(std, n_per_unit, units, q01, q01_per_unit, thr_out, a_out, b_out, ws, stream);
        SliceParams p{};
        p.std = std; p.q01 = q01; p.q01_per_unit = q01_per_unit;
        p.n = n_per_unit; p.units = units;
        p.thr_out = thr_out; p.a_out = a_out; p.b_out = b_out;
        p.apply_kind = 0;
        return launch_fused(p, stream);
    }
    if (ws_bytes < rounds_ws_bytes(units) || !ws) return PIC_ERR_WORKSPACE;
    return select_large_or_rounds(std, n_per_unit, units, q01, q01_per_unit, thr_out, a_out, b_out, ws, ws_bytes, stream);
}

int pic_select_threshold_multi(const float *std, int64_t n_per_unit, int64_t units, const float *q01_levels,
                               int levels, float *thr_out, pic_stream_t stream_) {
    int rc = check_common(n_per_unit, units);
    if (rc != PIC_OK) return rc;
    if (!std || !q01_levels || !thr_out || levels < 1) return PIC_ERR_INVALID_ARGUMENT;
    if (n_per_unit > kFusedMaxElems) return PIC_ERR_TOO_LARGE;   // larger units: call pic_select_threshold per level
    if (!aligned4(std)) return PIC_ERR_UNALIGNED;
    SliceParams p{};
    p.std = std;
    p.q01_per_unit = q01_levels;          // [units * levels], level-minor
    p.n = n_per_unit;
    p.units = units * levels;             // virtual units: `levels` consecutive ones share one std block
    p.repeat = levels;
    p.thr_out = thr_out;
    p.apply_kind = 0;
    return launch_fused(p, static_cast<cudaStream_t>(stream_));
}

int pic_level_map(const float *std, const float *thr, int64_t n_per_unit, int64_t units, int levels,
                  int32_t *level, pic_stream_t stream_) {
    if (n_per_unit <= 0 || units <= 0 || levels < 1 || levels > 4096 || !std || !thr || !level)
        return PIC_ERR_INVALID_ARGUMENT;
    const int64_t total = n_per_unit * units;
    level_map_kernel<<<elementwise_grid(total, 1), 256, levels * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(
        std, thr, n_per_unit, total, levels, level);
    return launch_status();
}

int pic_select_begin(void *state, int64_t n_total, int64_t units, float q01, const float *q01_per_unit,
                     pic_stream_t stream_) {
    int rc = check_common(n_total, units);
    if (rc != PIC_OK) return rc;
    if (!state) return PIC_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    select_begin_kernel<<<static_cast<unsigned>((units + 127) / 128), 128, 0, stream>>>(
        static_cast<SelectState *>(state), n_total, units, q01, q01_per_unit);
    return launch_status();
}

int pic_hist_round(const float *std_local, int64_t n_local, int64_t units, int round, const void *state,
                   uint32_t *hist, uint32_t *min_above, pic_stream_t stream_) {
    if (round < 0 || round > 2 || units <= 0 || n_local < 0 || !state || !hist) return PIC_ERR_INVALID_ARGUMENT;
    if (round == 2 && !min_above) return PIC_ERR_INVALID_ARGUMENT;
    if (n_local > 0 && !std_local) return PIC_ERR_INVALID_ARGUMENT;
    return launch_hist_round(std_local, n_local, units, round, static_cast<const SelectState *>(state), hist,
                             min_above, static_cast<cudaStream_t>(stream_));
}

int pic_select_advance(void *state, const uint32_t *hist, int64_t units, int round, pic_stream_t stream_) {
    if (round < 0 || round > 2 || units <= 0 || !state || !hist) return PIC_ERR_INVALID_ARGUMENT;
    return launch_advance(static_cast<SelectState *>(state), hist, units, round, static_cast<cudaStream_t>(stream_));
}

int pic_select_finish(const void *state, const uint32_t *min_above, int64_t units, float *thr_out, float *a_out,
                      float *b_out, pic_stream_t stream_) {
    if (units <= 0 || !state || !min_above || !thr_out) return PIC_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    select_finish_kernel<<<static_cast<unsigned>((units + 127) / 128), 128, 0, stream>>>(
        static_cast<const SelectState *>(state), min_above, units, thr_out, a_out, b_out);
    return launch_status();
}

// ------------------------------------------------------------------------------------------
// Spatially tiled select with NCCL issued from this library (one call = begin + 3 x (histogram kernel,
// all-reduce, advance) + min all-reduce + finish on the caller's stream: no host work between the steps).
// NCCL is the copy torch already loaded (dlopen of libnccl.so.2); uint32 sum / min are native NCCL types.
// ------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    decltype(&ncclGetUniqueId) get_unique_id = nullptr;
    decltype(&ncclCommInitRank) comm_init_rank = nullptr;
    decltype(&ncclCommDestroy) comm_destroy = nullptr;
    decltype(&ncclAllReduce) all_reduce = nullptr;
    decltype(&ncclAllGather) all_gather = nullptr;
    decltype(&ncclCommCount) comm_count = nullptr;
    decltype(&ncclGroupStart) group_start = nullptr;
    decltype(&ncclGroupEnd) group_end = nullptr;
    bool ok = false;
};
const NcclApi &nccl_api() {
    static const NcclApi api = [] {
        NcclApi a;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the one already in the process (torch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
        if (!h) {
            const char *path = getenv("PIC_NCCL_LIB");
            if (path) h = dlopen(path, RTLD_NOW);
        }
        if (!h) return a;
        a.get_unique_id = reinterpret_cast<decltype(a.get_unique_id)>(dlsym(h, "ncclGetUniqueId"));
        a.comm_init_rank = reinterpret_cast<decltype(a.comm_init_rank)>(dlsym(h, "ncclCommInitRank"));
        a.comm_destroy = reinterpret_cast<decltype(a.comm_destroy)>(dlsym(h, "ncclCommDestroy"));
        a.all_reduce = reinterpret_cast<decltype(a.all_reduce)>(dlsym(h, "ncclAllReduce"));
        a.all_gather = reinterpret_cast<decltype(a.all_gather)>(dlsym(h, "ncclAllGather"));
        a.comm_count = reinterpret_cast<decltype(a.comm_count)>(dlsym(h, "ncclCommCount"));
        a.group_start = reinterpret_cast<decltype(a.group_start)>(dlsym(h, "ncclGroupStart"));
        a.group_end = reinterpret_cast<decltype(a.group_end)>(dlsym(h, "ncclGroupEnd"));
        a.ok = a.get_unique_id && a.comm_init_rank && a.comm_destroy && a.all_reduce && a.all_gather && a.comm_count &&
               a.group_start && a.group_end;
        return a;
    }();
    return api;
}
inline int nccl_status(ncclResult_t r) {
    if (r == ncclSuccess) return PIC_OK;
    pic::g_last_cuda_error = 100000 + static_cast<int>(r);   // NCCL results are reported above the CUDA range
    return PIC_ERR_CUDA;
}
}  // namespace

size_t pic_tiled_workspace_bytes(int64_t units) { return rounds_ws_bytes(units < 1 ? 1 : units); }

int pic_dist_unique_id(unsigned char *id_out) {
    static_assert(sizeof(ncclUniqueId) == PIC_DIST_ID_BYTES, "ncclUniqueId is 128 bytes");
    if (!id_out) return PIC_ERR_INVALID_ARGUMENT;
    if (!nccl_api().ok) return PIC_ERR_CUDA;
    ncclUniqueId id;
    const int rc = nccl_status(nccl_api().get_unique_id(&id));
    if (rc == PIC_OK) memcpy(id_out, &id, sizeof(id));
    return rc;
}

int pic_dist_comm_init(const unsigned char *id, int rank, int world_size, void **comm_out) {
    if (!id || !comm_out || world_size < 1 || rank < 0 || rank >= world_size) return PIC_ERR_INVALID_ARGUMENT;
    if (!nccl_api().ok) return PIC_ERR_CUDA;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    const int rc = nccl_status(nccl_api().comm_init_rank(&comm, world_size, uid, rank));
    *comm_out = comm;
    return rc;
}

int pic_dist_comm_destroy(void *comm) {
    if (!comm) return PIC_OK;
    if (!nccl_api().ok) return PIC_ERR_CUDA;
    return nccl_status(nccl_api().comm_destroy(static_cast<ncclComm_t>(comm)));
}

int pic_tiled_select_threshold(const float *std_local, int64_t n_local, int64_t n_total, int64_t units, float q01,
                               const float *q01_per_unit, float *thr_out, void *ws, size_t ws_bytes, void *comm_,
                               pic_stream_t stream_) {
    int rc = check_common(n_total, units);
    if (rc != PIC_OK) return rc;
    if (n_local < 0 || n_local > n_total || (n_local > 0 && !std_local) || !thr_out || !comm_) return PIC_ERR_INVALID_ARGUMENT;
    if (!ws || ws_bytes < rounds_ws_bytes(units)) return PIC_ERR_WORKSPACE;
    if (!nccl_api().ok) return PIC_ERR_CUDA;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ncclComm_t comm = static_cast<ncclComm_t>(comm_);
    RoundsWs w = carve_ws(ws, units);
    select_begin_kernel<<<static_cast<unsigned>((units + 127) / 128), 128, 0, stream>>>(w.state, n_total, units, q01,
                                                                                      q01_per_unit);
    rc = launch_status();
    const NcclApi &nc = nccl_api();
    for (int r = 0; r < 3 && rc == PIC_OK; ++r) {
        rc = launch_hist_round(std_local, n_local, units, r, w.state, w.hist[r], w.min_above, stream);
        if (rc != PIC_OK) break;
        // the last round's kernel also produced the per-rank minimum key above the bucket: its MIN all-reduce
        // rides in the same NCCL group (one fused launch) as the histogram's SUM all-reduce
        if (r == 2) rc = nccl_status(nc.group_start());
        if (rc == PIC_OK)
            rc = nccl_status(nc.all_reduce(w.hist[r], w.hist[r], static_cast<size_t>(units) * kHistWords, ncclUint32,
                                           ncclSum, comm, stream));
        if (r == 2) {
            if (rc == PIC_OK)
                rc = nccl_status(nc.all_reduce(w.min_above, w.min_above, static_cast<size_t>(units), ncclUint32, ncclMin, comm, stream));
            const int rc_end = nccl_status(nc.group_end());
            if (rc == PIC_OK) rc = rc_end;
        }
        if (rc == PIC_OK) rc = launch_advance(w.state, w.hist[r], units, r, stream);
    }
    if (rc != PIC_OK) return rc;
    select_finish_kernel<<<static_cast<unsigned>((units + 127) / 128), 128, 0, stream>>>(w.state, w.min_above, units,
                                                                                      thr_out, nullptr, nullptr);
    return launch_status();
}

// Sampled protocol of the tiled select: TWO collectives instead of four, one pass over the band instead of three.
//   sample the band -> all-gather the samples -> every rank derives the SAME bracket pivots from the pooled sample ->
//   one tile-ordered sweep of the band (count below, collect the bracket's elements) -> all-gather counts + candidates ->
//   every rank selects the exact order statistics among the pooled candidates (cluster select) -> identical thresholds.
// A bracket that missed (or a slot that overflowed) is detected identically on every rank; then -- and only then -- the
// histogram-round protocol runs.  That decision needs one 4-byte read-back: this entry synchronises `stream` once and is
// therefore not CUDA-graph capturable (pic_tiled_select_threshold is).
namespace {
struct TiledPlan {
    int world, s_slot, S, tiles;
    int64_t cap_x, stride, cap_all;
    size_t off_send_samp, off_pooled, off_rank, off_piv, off_st, off_below, off_send_x, off_all_x, off_cand_all,
        off_invalid, off_rounds, total;
};
size_t up256(size_t v) { return (v + 255) / 256 * 256; }
// slot_margin: how many times its fair share of the bracket's elements one rank's slot can hold (bands of real images
// differ: a threshold inside the texture of the lower half puts most candidates into the lower bands).  The NCCL
// transport moves whole slots, so it stays at 2; the peer-memory transport moves only what exists and takes 4.
constexpr double kSlotMarginNccl = 2.0, kSlotMarginP2p = 4.0;
bool tiled_plan(int64_t n_local, int64_t n_total, int64_t units, int world, double slot_margin, TiledPlan &t) {
    if (world < 1 || world > 64 || units < 1 || n_local < 1) return false;
    t.world = world;
    int s_slot = (32768 / world) & ~3;
    if (s_slot < 1284) s_slot = 1284;                         // pooled sample > kCandMax for every world size
    while (static_cast<int64_t>(s_slot) * world > kFusedMaxElems) s_slot -= 4;
    t.s_slot = s_slot;
    t.S = s_slot * world;
    if (t.S <= kCandMax) return false;
    t.tiles = static_cast<int>((n_local + kGsTile - 1) / kGsTile);
    const double frac = (8.0 * sqrt(t.S * 0.25) + 10.0) / t.S;
    // exchange sizes must be the same on every rank: they derive from n_total / world, not from this rank's band (a band
    // much larger than its share overflows its slot -> every rank sees the flag -> histogram rounds)
    const double share = static_cast<double>((n_total + world - 1) / world);
    int64_t cap = static_cast<int64_t>(slot_margin * frac * share) + 1024;
    if (cap > static_cast<int64_t>(share) + 1024) cap = static_cast<int64_t>(share) + 1024;
    t.cap_x = (cap + 3) & ~int64_t(3);
    t.stride = t.cap_x + 4;
    t.cap_all = t.cap_x * world;
    if (t.cap_all > (int64_t(1) << 31) || units * t.tiles > 0x7fffffffLL) return false;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += up256(bytes); return at; };
    t.off_send_samp = take(static_cast<size_t>(units) * s_slot * 4);
    t.off_pooled = take(static_cast<size_t>(units) * t.S * 4);
    t.off_rank = take(static_cast<size_t>(units) * 8);
    t.off_piv = take(static_cast<size_t>(units) * 8);
    t.off_st = take(static_cast<size_t>(units) * sizeof(GsUnit));
    t.off_below = take(static_cast<size_t>(units) * t.tiles * 4);
    t.off_send_x = take(static_cast<size_t>(units) * t.stride * 4 + 64);
    t.off_all_x = take(static_cast<size_t>(world) * units * t.stride * 4 + 64);
    t.off_cand_all = take(static_cast<size_t>(units) * t.cap_all * 4);
    t.off_invalid = take(256);
    t.off_rounds = take(rounds_ws_bytes(units));
    t.total = o + 256;
    return true;
}
}  // namespace

size_t pic_tiled_sampled_workspace_bytes(int64_t n_local, int64_t n_total, int64_t units, int world_size) {
    TiledPlan t;
    if (!tiled_plan(n_local, n_total, units < 1 ? 1 : units, world_size, kSlotMarginP2p, t)) return rounds_ws_bytes(units < 1 ? 1 : units);
    return t.total;
}

namespace {
struct P2pCtx {
    int rank = 0, world = 1;
    size_t bytes = 0, region_bytes[2] = {0, 0};
    unsigned char *window[kP2pMaxWorld] = {};   // window[rank] is this rank's own allocation
    unsigned char **windows_dev = nullptr;      // the same table in device memory
    ncclComm_t comm = nullptr;
    int device = 0;
};

// all ranks meet here (used around window setup / teardown only)
int p2p_barrier(const P2pCtx &c, uint32_t *scratch_dev, cudaStream_t stream) {
    const int rc = nccl_status(nccl_api().all_reduce(scratch_dev, scratch_dev, 1, ncclUint32, ncclSum, c.comm, stream));
    if (rc != PIC_OK) return rc;
    PIC_CUDA_CHECK(cudaStreamSynchronize(stream));
    return PIC_OK;
}

int tiled_sampled_impl(const float *std_local, int64_t n_local, int64_t n_total, int64_t units, float q01,
                       const float *q01_per_unit, float *thr_out, void *ws, size_t ws_bytes, void *comm_, P2pCtx *p2p,
                       uint32_t *status_dev, pic_stream_t stream_, int *used_fallback) {
    int rc = check_common(n_total, units);
    if (rc != PIC_OK) return rc;
    if (n_local < 0 || n_local > n_total || (n_local > 0 && !std_local) || !thr_out || !comm_ || !ws) return PIC_ERR_INVALID_ARGUMENT;
    if (!nccl_api().ok) return PIC_ERR_CUDA;
    const NcclApi &nc = nccl_api();
    ncclComm_t comm = static_cast<ncclComm_t>(comm_);
    int world = 0;
    rc = nccl_status(nc.comm_count(comm, &world));
    if (rc != PIC_OK) return rc;
    if (used_fallback) *used_fallback = 0;
    TiledPlan t;
    // every rank must take the same branch: the plan depends on this rank's band, so ranks with unequal bands agree only
    // when all of them can plan -- bands smaller than a sample slot are the caller's job to avoid (documented)
    if (n_local < 1 || !aligned4(std_local)) return PIC_ERR_INVALID_ARGUMENT;
    const bool planned = tiled_plan(n_local, n_total, units, world, p2p ? kSlotMarginP2p : kSlotMarginNccl, t) && ws_bytes >= t.total &&
                         units <= 65535;
    const size_t x1_bytes = planned ? static_cast<size_t>(units) * t.S * 4 : 0;
    const size_t x2_bytes = planned ? static_cast<size_t>(world) * units * t.stride * 4 : 0;
    if (p2p) {
        // peer-memory transport: no host decision inside (graph capturable), so "cannot plan" is an error, not a fallback
        if (!planned || p2p->world != world || x1_bytes > p2p->region_bytes[0] || x2_bytes > p2p->region_bytes[1] || !status_dev)
            return PIC_ERR_WORKSPACE;
    } else if (!planned) {
        if (used_fallback) *used_fallback = 1;
        return pic_tiled_select_threshold(std_local, n_local, n_total, units, q01, q01_per_unit, thr_out, ws, ws_bytes, comm_, stream_);
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    unsigned char *b = static_cast<unsigned char *>(ws);
    b = reinterpret_cast<unsigned char *>(up256(reinterpret_cast<size_t>(b)));
    float *send_samp = reinterpret_cast<float *>(b + t.off_send_samp);
    float *pooled = reinterpret_cast<float *>(b + t.off_pooled);
    uint32_t *rank_in = reinterpret_cast<uint32_t *>(b + t.off_rank);
    float *piv = reinterpret_cast<float *>(b + t.off_piv);
    GsUnit *st = reinterpret_cast<GsUnit *>(b + t.off_st);
    uint32_t *below_tile = reinterpret_cast<uint32_t *>(b + t.off_below);
    uint32_t *send_x = reinterpret_cast<uint32_t *>(b + t.off_send_x);
    uint32_t *all_x = reinterpret_cast<uint32_t *>(b + t.off_all_x);
    uint32_t *cand_all = reinterpret_cast<uint32_t *>(b + t.off_cand_all);
    uint32_t *invalid = reinterpret_cast<uint32_t *>(b + t.off_invalid);
    void *rounds_ws = b + t.off_rounds;
    size_t x1_off = 0, x2_off = 0;
    if (p2p) {
        x1_off = kP2pDataOff;
        x2_off = kP2pDataOff + p2p->region_bytes[0];
        pooled = reinterpret_cast<float *>(p2p->window[p2p->rank] + x1_off);
        all_x = reinterpret_cast<uint32_t *>(p2p->window[p2p->rank] + x2_off);
        invalid = status_dev;                       // the caller owns (and clears) the status word
    } else {
        PIC_CUDA_CHECK(cudaMemsetAsync(invalid, 0, 4, stream));
    }
    // 1. sample the band (and derive the pivot ranks of the pooled sample), exchange: unit u's pooled sample is contiguous
    if (p2p) {
        // one kernel: the samples are stored straight into every rank's window
        tile_sample_exchange_kernel<<<dim3((t.s_slot + 255) / 256, static_cast<unsigned>(units)), 256, 0, stream>>>(
            std_local, n_local, units, t.s_slot, rank_in, n_total, q01, q01_per_unit, p2p->windows_dev, p2p->rank, world, x1_off,
            status_dev);
        rc = launch_status();
    } else {
        tile_sample_kernel<<<dim3((t.s_slot + 255) / 256, static_cast<unsigned>(units)), 256, 0, stream>>>(
            std_local, n_local, units, t.s_slot, send_samp, rank_in, world, n_total, q01, q01_per_unit);
        rc = launch_status();
        if (rc != PIC_OK) return rc;
        // one all-gather per unit inside a group (a single fused NCCL launch)
        rc = nccl_status(nc.group_start());
        for (int64_t u = 0; u < units && rc == PIC_OK; ++u)
            rc = nccl_status(nc.all_gather(send_samp + u * t.s_slot, pooled + u * t.S, static_cast<size_t>(t.s_slot), ncclFloat, comm, stream));
        const int rc_end = nccl_status(nc.group_end());
        if (rc == PIC_OK) rc = rc_end;
    }
    if (rc != PIC_OK) return rc;
    // 2. pivots: the two order statistics of every pooled sample (lean select, explicit ranks, two virtual units per unit)
    {
        SliceParams sp{};
        sp.std = pooled; sp.n = t.S; sp.units = units * 2; sp.repeat = 2; sp.rank_in = rank_in;
        sp.q01 = 0.5f; sp.thr_out = nullptr; sp.a_out = piv; sp.apply_kind = 0;
        rc = launch_select_lean(sp, stream);
        if (rc != PIC_OK) return rc;
    }
    tile_pivots_kernel<<<static_cast<unsigned>((units + 127) / 128), 128, 0, stream>>>(piv, st, units, n_total, t.S, q01,
                                                                                      q01_per_unit, thr_out);
    rc = launch_status();
    if (rc != PIC_OK) return rc;
    // 3. one sweep of the band: per-tile counts below the bracket, bracket elements straight into the exchange slot;
    // 4. exchange counts + candidates
    GsParams g{};
    g.std = std_local; g.q01_per_unit = q01_per_unit; g.q01 = q01; g.n = n_local; g.units = units;
    g.st = st; g.cand = send_x + 4; g.cand_cap = t.stride; g.below_tile = below_tile;
    g.vec = (n_local % 4 == 0 && aligned16(std_local)) ? 1 : 0;
    g.thr = thr_out;
    if (g.vec) gs_sweep_kernel<true, true><<<static_cast<unsigned>(units * t.tiles), kGsThreads, 0, stream>>>(g, t.tiles);
    else gs_sweep_kernel<false, true><<<static_cast<unsigned>(units * t.tiles), kGsThreads, 0, stream>>>(g, t.tiles);
    rc = launch_status();
    if (rc != PIC_OK) return rc;
    if (p2p) {
        // only the header (computed inside the exchange kernel) and the candidates that exist cross the links (NCCL has
        // to move the whole fixed-size slot)
        p2p_exchange_kernel<<<dim3(8, static_cast<unsigned>(units)), 256, 0, stream>>>(
            p2p->windows_dev, p2p->rank, world, 1, send_x, t.stride, -1, x2_off, t.stride, status_dev,
            P2pPack{st, below_tile, t.tiles, static_cast<uint32_t>(t.cap_x)});
        rc = launch_status();
    } else {
        tile_pack_kernel<<<static_cast<unsigned>(units), 128, 0, stream>>>(st, below_tile, t.tiles, send_x, t.stride,
                                                                           static_cast<uint32_t>(t.cap_x));
        rc = launch_status();
        if (rc != PIC_OK) return rc;
        rc = nccl_status(nc.group_start());
        for (int64_t u = 0; u < units && rc == PIC_OK; ++u)
            rc = nccl_status(nc.all_gather(send_x + u * t.stride, all_x + u * world * t.stride, static_cast<size_t>(t.stride), ncclUint32,
                                           comm, stream));
        const int rc_end = nccl_status(nc.group_end());
        if (rc == PIC_OK) rc = rc_end;
    }
    if (rc != PIC_OK) return rc;
    // 5. merge, exact select among the pooled candidates
    const int split = world >= 32 ? 1 : 32 / world;
    tile_merge_kernel<<<dim3(static_cast<unsigned>(world * split), static_cast<unsigned>(units)), 256, 0, stream>>>(
        all_x, world, split, units, t.stride, st, cand_all, t.cap_all, static_cast<uint32_t>(t.cap_x), n_total, q01, q01_per_unit,
        invalid, thr_out);
    rc = launch_status();
    if (rc != PIC_OK) return rc;
    GsParams f{};
    f.std = nullptr; f.q01_per_unit = q01_per_unit; f.q01 = q01; f.n = n_total; f.units = units;
    f.st = st; f.cand = cand_all; f.cand_cap = t.cap_all; f.below_tile = nullptr; f.vec = 1;
    f.thr = thr_out;
    gs_cluster_select_kernel<true><<<dim3(kClusterCtas, static_cast<unsigned>(units)), kClusterThreads, 0, stream>>>(f, 0);
    rc = launch_status();
    if (rc != PIC_OK) return rc;
    if (p2p) return PIC_OK;     // the caller reads the status word at its next synchronisation point
    // 5. the one read-back: did every unit's bracket hold?  (identical on all ranks: derived from all-gathered data)
    uint32_t n_invalid = 0;
    PIC_CUDA_CHECK(cudaMemcpyAsync(&n_invalid, invalid, 4, cudaMemcpyDeviceToHost, stream));
    PIC_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (n_invalid != 0u) {
        if (used_fallback) *used_fallback = 1;
        return pic_tiled_select_threshold(std_local, n_local, n_total, units, q01, q01_per_unit, thr_out, rounds_ws,
                                          rounds_ws_bytes(units), comm_, stream_);
    }
    return PIC_OK;
}
}  // namespace

int pic_tiled_select_threshold_sampled(const float *std_local, int64_t n_local, int64_t n_total, int64_t units, float q01,
                                       const float *q01_per_unit, float *thr_out, void *ws, size_t ws_bytes, void *comm_,
                                       pic_stream_t stream_, int *used_fallback) {
    return tiled_sampled_impl(std_local, n_local, n_total, units, q01, q01_per_unit, thr_out, ws, ws_bytes, comm_, nullptr, nullptr,
                              stream_, used_fallback);
}

// ---- peer-memory transport of the sampled protocol ---------------------------------------------------------------------
int pic_dist_p2p_region_bytes(int64_t n_total, int64_t units, int world_size, size_t *sample_bytes, size_t *cand_bytes) {
    TiledPlan t;
    if (!sample_bytes || !cand_bytes || n_total < 1 || units < 1) return PIC_ERR_INVALID_ARGUMENT;
    if (!tiled_plan(n_total, n_total, units, world_size, kSlotMarginP2p, t)) return PIC_ERR_TOO_LARGE;
    *sample_bytes = up256(static_cast<size_t>(units) * t.S * 4);
    *cand_bytes = up256(static_cast<size_t>(world_size) * units * t.stride * 4);
    return PIC_OK;
}

int pic_dist_p2p_init(void *comm_, int rank, size_t sample_bytes, size_t cand_bytes, void **p2p_out) {
    if (!comm_ || !p2p_out || rank < 0) return PIC_ERR_INVALID_ARGUMENT;
    if (!nccl_api().ok) return PIC_ERR_CUDA;
    *p2p_out = nullptr;
    P2pCtx *c = new P2pCtx();
    c->comm = static_cast<ncclComm_t>(comm_);
    int rc = nccl_status(nccl_api().comm_count(c->comm, &c->world));
    if (rc != PIC_OK || c->world > kP2pMaxWorld || rank >= c->world) {
        delete c;
        return rc != PIC_OK ? rc : PIC_ERR_INVALID_ARGUMENT;
    }
    c->rank = rank;
    c->region_bytes[0] = up256(sample_bytes);
    c->region_bytes[1] = up256(cand_bytes);
    c->bytes = kP2pDataOff + c->region_bytes[0] + c->region_bytes[1];
    cudaStream_t stream = nullptr;
    unsigned char *handles_dev = nullptr;
    auto fail = [&](int code) {
        for (int p = 0; p < c->world; ++p)
            if (p != c->rank && c->window[p]) cudaIpcCloseMemHandle(c->window[p]);
        if (c->window[c->rank]) cudaFree(c->window[c->rank]);
        if (c->windows_dev) cudaFree(c->windows_dev);
        if (handles_dev) cudaFree(handles_dev);
        if (stream) cudaStreamDestroy(stream);
        delete c;
        return code;
    };
#define PIC_P2P_TRY(call)                                                  \
    do {                                                                   \
        const cudaError_t e_ = (call);                                     \
        if (e_ != cudaSuccess) {                                           \
            pic::g_last_cuda_error = static_cast<int>(e_);                 \
            (void)cudaGetLastError();                                      \
            return fail(PIC_ERR_CUDA);                                     \
        }                                                                  \
    } while (0)
    PIC_P2P_TRY(cudaGetDevice(&c->device));
    PIC_P2P_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaIpcMemHandle_t mine;
    // a rank that cannot allocate or export still has to take part in the all-gather below (else the peers hang): it
    // sends zeros and every rank fails together
    cudaError_t exported = cudaMalloc(&c->window[c->rank], c->bytes);
    if (exported == cudaSuccess) exported = cudaMemsetAsync(c->window[c->rank], 0, c->bytes, stream);
    if (exported == cudaSuccess) exported = cudaStreamSynchronize(stream);
    if (exported == cudaSuccess) exported = cudaIpcGetMemHandle(&mine, c->window[c->rank]);
    if (exported != cudaSuccess) {
        pic::g_last_cuda_error = static_cast<int>(exported);
        (void)cudaGetLastError();
        memset(&mine, 0, sizeof(mine));
    }
    const size_t hb = sizeof(cudaIpcMemHandle_t) + 8;     // handle + "ok" marker
    PIC_P2P_TRY(cudaMalloc(&handles_dev, hb * (c->world + 1)));
    unsigned char send[sizeof(cudaIpcMemHandle_t) + 8] = {};
    memcpy(send, &mine, sizeof(mine));
    send[sizeof(mine)] = exported == cudaSuccess ? 1 : 0;
    PIC_P2P_TRY(cudaMemcpyAsync(handles_dev + hb * c->world, send, hb, cudaMemcpyHostToDevice, stream));
    rc = nccl_status(nccl_api().all_gather(handles_dev + hb * c->world, handles_dev, hb, ncclUint8, c->comm, stream));
    if (rc != PIC_OK) return fail(rc);
    std::vector<unsigned char> all(hb * c->world);
    PIC_P2P_TRY(cudaMemcpyAsync(all.data(), handles_dev, all.size(), cudaMemcpyDeviceToHost, stream));
    PIC_P2P_TRY(cudaStreamSynchronize(stream));
    bool every_ok = true;
    for (int p = 0; p < c->world; ++p) every_ok = every_ok && all[hb * p + sizeof(cudaIpcMemHandle_t)] == 1;
    int open_failed = 0;
    if (every_ok) {
        for (int p = 0; p < c->world && !open_failed; ++p) {
            if (p == c->rank) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + hb * p, sizeof(h));
            void *ptr = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                pic::g_last_cuda_error = static_cast<int>(e);
                (void)cudaGetLastError();
                open_failed = 1;
            } else {
                c->window[p] = static_cast<unsigned char *>(ptr);
            }
        }
    }
    // agree on the outcome: nobody may start storing into a window some rank failed to map (and then freed)
    uint32_t *flag_dev = reinterpret_cast<uint32_t *>(handles_dev);
    const uint32_t bad = (every_ok && !open_failed) ? 0u : 1u;
    PIC_P2P_TRY(cudaMemcpyAsync(flag_dev, &bad, 4, cudaMemcpyHostToDevice, stream));
    rc = p2p_barrier(*c, flag_dev, stream);
    if (rc != PIC_OK) return fail(rc);
    uint32_t bad_total = 0;
    PIC_P2P_TRY(cudaMemcpy(&bad_total, flag_dev, 4, cudaMemcpyDeviceToHost));
    if (bad_total != 0u) return fail(PIC_ERR_CUDA);
    if (const char *e = getenv("PIC_P2P_TIMEOUT_MS")) {     // ranks that may skew by more than the default 4 s raise it
        long long ms = atoll(e);
        if (ms < 100) ms = 100;
        const unsigned long long ns = static_cast<unsigned long long>(ms) * 1000000ull;
        PIC_P2P_TRY(cudaMemcpyToSymbol(g_p2p_timeout_ns, &ns, sizeof(ns)));
    }
    PIC_P2P_TRY(cudaMalloc(&c->windows_dev, sizeof(unsigned char *) * kP2pMaxWorld));
    PIC_P2P_TRY(cudaMemcpy(c->windows_dev, c->window, sizeof(unsigned char *) * kP2pMaxWorld, cudaMemcpyHostToDevice));
    cudaFree(handles_dev);
    cudaStreamDestroy(stream);
#undef PIC_P2P_TRY
    *p2p_out = c;
    return PIC_OK;
}

int pic_dist_p2p_destroy(void *p2p_) {
    if (!p2p_) return PIC_OK;
    P2pCtx *c = static_cast<P2pCtx *>(p2p_);
    int rc = PIC_OK;
    uint32_t *word = nullptr;
    // peers may still be storing into this window: meet first, unmap, meet again, then free
    if (cudaMalloc(&word, 4) == cudaSuccess && cudaMemset(word, 0, 4) == cudaSuccess) {
        cudaDeviceSynchronize();
        rc = p2p_barrier(*c, word, nullptr);
        for (int p = 0; p < c->world; ++p)
            if (p != c->rank && c->window[p]) cudaIpcCloseMemHandle(c->window[p]);
        if (rc == PIC_OK) rc = p2p_barrier(*c, word, nullptr);
    } else {
        rc = PIC_ERR_CUDA;
    }
    if (word) cudaFree(word);
    if (c->window[c->rank]) cudaFree(c->window[c->rank]);
    if (c->windows_dev) cudaFree(c->windows_dev);
    delete c;
    return rc;
}

int pic_tiled_select_threshold_p2p(const float *std_local, int64_t n_local, int64_t n_total, int64_t units, float q01,
                                   const float *q01_per_unit, float *thr_out, void *ws, size_t ws_bytes, void *p2p_,
                                   uint32_t *status_dev, pic_stream_t stream_) {
    if (!p2p_) return PIC_ERR_INVALID_ARGUMENT;
    P2pCtx *c = static_cast<P2pCtx *>(p2p_);
    return tiled_sampled_impl(std_local, n_local, n_total, units, q01, q01_per_unit, thr_out, ws, ws_bytes, c->comm, c, status_dev,
                              stream_, nullptr);
}

int pic_channel_mask(const float *std, int64_t n_per_unit, int64_t units, float q01, const float *q01_per_unit,
                     float *mask, float *thr_out, void *ws, size_t ws_bytes, pic_stream_t stream_) {
    int rc = check_common(n_per_unit, units);
    if (rc != PIC_OK) return rc;
    if (!std || !mask) return PIC_ERR_INVALID_ARGUMENT;
    if (!aligned4(std) || !aligned4(mask)) return PIC_ERR_UNALIGNED;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SliceParams p{};
    p.std = std; p.q01 = q01; p.q01_per_unit = q01_per_unit;
    p.n = n_per_unit; p.units = units;
    p.mask = mask; p.thr_out = thr_out;
    p.apply_kind = 1;
    const bool needs_select = q01_per_unit || unit_mode(q01) == kModeThreshold;
    if (n_per_unit <= kFusedMaxElems) return launch_fused(p, stream);
    if (needs_select) {
        if (ws_bytes < rounds_ws_bytes(units) || !ws) return PIC_ERR_WORKSPACE;
        RoundsWs w = carve_ws(ws, units);
        rc = select_large_or_rounds(std, n_per_unit, units, q01, q01_per_unit, w.thr, nullptr, nullptr, ws, ws_bytes,
                                    stream);
        if (rc != PIC_OK) return rc;
        p.thr_in = w.thr;
    }
    return launch_apply(p, stream);
}

int pic_attention_mask(const float *std, int64_t n_per_unit, int64_t units, float q01, const float *q01_per_unit,
                       int copies, float *mask, float *thr_out, void *ws, size_t ws_bytes, pic_stream_t stream_) {
    int rc = check_common(n_per_unit, units);
    if (rc != PIC_OK) return rc;
    if (!std || !mask || copies < 1 || copies > 8) return PIC_ERR_INVALID_ARGUMENT;
    if (!aligned4(std) || !aligned4(mask)) return PIC_ERR_UNALIGNED;
    if (copies == 1) return pic_channel_mask(std, n_per_unit, units, q01, q01_per_unit, mask, thr_out, ws, ws_bytes, stream_);
    if (!ws || ws_bytes < pic_workspace_bytes(n_per_unit, units)) return PIC_ERR_WORKSPACE;
    // thresholds (select only; ones / zeros sentinels give -inf / +inf), then one pass writes every copy
    float *thr = thr_out;
    if (!thr) thr = (n_per_unit <= kFusedMaxElems) ? gs_thr_buffer(ws, units) : carve_ws(ws, units).thr;
    rc = pic_select_threshold(std, n_per_unit, units, q01, q01_per_unit, thr, nullptr, nullptr, ws, ws_bytes, stream_);
    if (rc != PIC_OK) return rc;
    const int64_t total = n_per_unit * units;
    const bool vec = (n_per_unit % 4 == 0) && aligned16(std) && aligned16(mask);
    mask_from_threshold_kernel<<<elementwise_grid(total, vec ? 4 : 1), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        std, thr, n_per_unit, total, mask, vec, copies, q01, q01_per_unit);
    return launch_status();
}

int pic_mask_from_threshold(const float *std, const float *thr, int64_t n_per_unit, int64_t units, float *mask,
                            pic_stream_t stream_) {
    if (n_per_unit <= 0 || units <= 0 || !std || !thr || !mask) return PIC_ERR_INVALID_ARGUMENT;
    const int64_t total = n_per_unit * units;
    const bool vec = (n_per_unit % 4 == 0) && aligned16(std) && aligned16(mask);
    mask_from_threshold_kernel<<<elementwise_grid(total, vec ? 4 : 1), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        std, thr, n_per_unit, total, mask, vec, 1, 0.5f, nullptr);
    return launch_status();
}

int pic_slice_forward(const float *y_top, const float *y_base, const float *mu, const float *std, float q01,
                      const float *q01_per_unit, const float *thr_in, const float *noise,
                      const float *scale_table, int table_len, float scale_bound, float lik_bound,
                      int64_t n_per_unit, int64_t units, float *mask, float *y_hat, float *lik, int32_t *idx,
                      int32_t *symbols, float *thr_out, double *rate, void *ws, size_t ws_bytes,
                      pic_stream_t stream_) {
    int rc = check_common(n_per_unit, units);
    if (rc != PIC_OK) return rc;
    if (!y_top || !mu || !std) return PIC_ERR_INVALID_ARGUMENT;
    if (idx && (!scale_table || table_len < 1)) return PIC_ERR_INVALID_ARGUMENT;
    if (!(scale_bound > 0.0f)) return PIC_ERR_INVALID_ARGUMENT;
    const void *ptrs[] = {y_top, y_base, mu, std, noise, mask, y_hat, lik, idx, symbols};
    for (const void *q : ptrs)
        if (q && !aligned4(q)) return PIC_ERR_UNALIGNED;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SliceParams p{};
    p.y_top = y_top; p.y_base = y_base; p.mu = mu; p.std = std;
    p.q01 = q01; p.q01_per_unit = q01_per_unit; p.thr_in = thr_in; p.noise = noise;
    p.table = scale_table; p.table_len = table_len;
    p.scale_bound = scale_bound; p.lik_bound = lik_bound;
    p.n = n_per_unit; p.units = units;
    p.mask = mask; p.y_hat = y_hat; p.lik = lik; p.idx = idx; p.symbols = symbols;
    p.thr_out = thr_out; p.rate = rate;
    p.apply_kind = 2;
    if (n_per_unit <= kFusedMaxElems) {
        const bool sel = !thr_in && (q01_per_unit || unit_mode(q01) == kModeThreshold);
        const bool have_ws = ws && ws_bytes >= gs_ws_bytes(units);
        float *thr_buf = thr_out ? thr_out : (have_ws ? gs_thr_buffer(ws, units) : nullptr);
        static const int two_kernel = [] { const char *e = getenv("PIC_TWO_KERNEL"); return e ? atoi(e) : 1; }();
        // thresholds known: the tile-ordered apply kernel (global-order streaming, ~99 % of roofline)
        if (two_kernel && thr_in) return launch_apply(p, stream);
        // measured cross-over (scripts/small_batch.py, phase_bench.py): for units of >= 32768 elements the
        // select kernel + tile-ordered apply wins at every batch size (few units: the apply spreads over
        // tiles; many units: global-order streaming); below that per-unit fixed costs favour one fused launch
        if (two_kernel && sel && thr_buf && n_per_unit >= kTwoKernelMinUnit) {
            // large batch: lean select kernel (6 CTAs/SM) -> thresholds -> tile-ordered apply kernel.
            // The apply kernel walks memory in global order and reaches ~99 % of the HBM roofline,
            // which beats one-CTA-per-unit streaming (thousands of concurrent DRAM streams).
            static const int gsel = [] { const char *e = getenv("PIC_GLOBAL_SELECT"); return e ? atoi(e) : 0; }();
            if (gsel && have_ws) {
                // global sampled select: std streamed tile by tile in global address order
                rc = select_sampled_global(std, n_per_unit, units, q01, q01_per_unit, thr_buf, nullptr, nullptr, ws, stream);
            } else {
                SliceParams ps = p;
                ps.apply_kind = 0;
                ps.thr_out = thr_buf;
                rc = launch_fused(ps, stream);
            }
            if (rc != PIC_OK) return rc;
            p.thr_in = thr_buf;
            p.thr_out = nullptr;
            return launch_apply(p, stream);
        }
        return launch_fused(p, stream);
    }
    const bool needs_select = !thr_in && (q01_per_unit || unit_mode(q01) == kModeThreshold);
    if (needs_select) {
        if (ws_bytes < rounds_ws_bytes(units) || !ws) return PIC_ERR_WORKSPACE;
        RoundsWs w = carve_ws(ws, units);
        rc = select_large_or_rounds(std, n_per_unit, units, q01, q01_per_unit, w.thr, nullptr, nullptr, ws, ws_bytes,
                                    stream);
        if (rc != PIC_OK) return rc;
        p.thr_in = w.thr;
    }
    return launch_apply(p, stream);
}

int pic_slice_forward_multi(const float *y_top, const float *y_base, const float *mu, const float *std,
                            const float *q01_levels, int levels, const float *scale_table, int table_len,
                            float scale_bound, float lik_bound, int64_t n_per_unit, int64_t units, float *mask,
                            float *y_hat, float *lik, int32_t *idx, int32_t *symbols, float *thr_out, double *rate,
                            pic_stream_t stream_) {
    int rc = check_common(n_per_unit, units);
    if (rc != PIC_OK) return rc;
    if (!y_top || !mu || !std || !q01_levels || !thr_out || levels < 1) return PIC_ERR_INVALID_ARGUMENT;
    if (idx && (!scale_table || table_len < 1)) return PIC_ERR_INVALID_ARGUMENT;
    if (!(scale_bound > 0.0f)) return PIC_ERR_INVALID_ARGUMENT;
    if (n_per_unit > kFusedMaxElems) return PIC_ERR_TOO_LARGE;   // larger units: one pic_slice_forward per level
    if (units * static_cast<int64_t>(levels) * n_per_unit > kMaxElemsPerLaunch) return PIC_ERR_TOO_LARGE;
    const void *ptrs[] = {y_top, y_base, mu, std, mask, y_hat, lik, idx, symbols};
    for (const void *q : ptrs)
        if (q && !aligned4(q)) return PIC_ERR_UNALIGNED;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    rc = pic_select_threshold_multi(std, n_per_unit, units, q01_levels, levels, thr_out, stream_);
    if (rc != PIC_OK) return rc;
    SliceParams p{};
    p.y_top = y_top; p.y_base = y_base; p.mu = mu; p.std = std;
    p.q01_per_unit = q01_levels; p.thr_in = thr_out;
    p.table = scale_table; p.table_len = table_len;
    p.scale_bound = scale_bound; p.lik_bound = lik_bound;
    p.n = n_per_unit; p.units = units * levels; p.repeat = levels;
    p.mask = mask; p.y_hat = y_hat; p.lik = lik; p.idx = idx; p.symbols = symbols; p.rate = rate;
    p.apply_kind = 2;
    static const int shared_eval = [] { const char *e = getenv("PIC_MULTI_SHARED_EVAL"); return e ? atoi(e) : 1; }();
    if (shared_eval && levels >= 4 && slice_vec_ok(p) && units <= 65535) {
        // kept / masked evaluation shared by the levels of a group: enough groups to fill the GPU about twice over, but at
        // least 8 levels per group (the two full evaluations per element are what the group amortises)
        const int64_t chunks = (n_per_unit + kMultiChunk - 1) / kMultiChunk;
        int64_t groups = (2LL * sm_count() * 4 + units * chunks - 1) / (units * chunks);
        if (groups > levels / 8) groups = levels / 8;
        if (groups < 1) groups = 1;
        if (chunks * groups <= 0x7fffffffLL) {
            if (rate) PIC_CUDA_CHECK(cudaMemsetAsync(rate, 0, sizeof(double) * p.units, stream));
            slice_apply_multi_kernel<<<dim3(static_cast<unsigned>(chunks * groups), static_cast<unsigned>(units)), 256, 0, stream>>>(
                p, levels, static_cast<int>(groups));
            return launch_status();
        }
    }
    return launch_apply(p, stream);
}

int pic_slice_backward(const float *g_lik, const float *g_yhat, const float *y_top, const float *y_base,
                       const float *mu, const float *std, const float *mask, const float *noise,
                       float scale_bound, float lik_bound, int64_t n, float *g_ytop, float *g_ybase, float *g_mu,
                       float *g_std, pic_stream_t stream_) {
    if (n <= 0 || !y_top || !mu || !std || !mask) return PIC_ERR_INVALID_ARGUMENT;
    BwdParams p{g_lik, g_yhat, y_top, y_base, mu, std, mask, noise, scale_bound, lik_bound, n,
                g_ytop, g_ybase, g_mu, g_std};
    const void *ptrs[] = {g_lik, g_yhat, y_top, y_base, mu, std, mask, noise, g_ytop, g_ybase, g_mu, g_std};
    bool vec = (n % 4 == 0);
    for (const void *q : ptrs) {
        if (q && !aligned4(q)) return PIC_ERR_UNALIGNED;
        if (q && !aligned16(q)) vec = false;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int grid = elementwise_grid(n, 4);
    if (noise) {
        if (vec) slice_backward_kernel<true, true><<<grid, 256, 0, stream>>>(p);
        else slice_backward_kernel<true, false><<<grid, 256, 0, stream>>>(p);
    } else {
        if (vec) slice_backward_kernel<false, true><<<grid, 256, 0, stream>>>(p);
        else slice_backward_kernel<false, false><<<grid, 256, 0, stream>>>(p);
    }
    return launch_status();
}

int pic_gaussian_forward(const float *inputs, const float *scales, const float *means, const float *noise,
                         int likelihood_only, int64_t n, float scale_bound, float lik_bound, float *outputs,
                         float *lik, pic_stream_t stream_) {
    if (n <= 0 || !inputs || !scales) return PIC_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int grid = elementwise_grid(n, 4);
    bool vec = (n % 4 == 0);
    for (const void *q : {static_cast<const void *>(inputs), static_cast<const void *>(scales), static_cast<const void *>(means),
                          static_cast<const void *>(noise), static_cast<const void *>(outputs), static_cast<const void *>(lik)})
        if (q && !aligned16(q)) vec = false;
    if (likelihood_only) gaussian_forward_kernel<2><<<grid, 256, 0, stream>>>(inputs, scales, means, noise, n, scale_bound, lik_bound, outputs, lik, vec);
    else if (noise) gaussian_forward_kernel<1><<<grid, 256, 0, stream>>>(inputs, scales, means, noise, n, scale_bound, lik_bound, outputs, lik, vec);
    else gaussian_forward_kernel<0><<<grid, 256, 0, stream>>>(inputs, scales, means, noise, n, scale_bound, lik_bound, outputs, lik, vec);
    return launch_status();
}

int pic_gaussian_backward(const float *g_out, const float *g_lik, const float *inputs, const float *scales,
                          const float *means, const float *noise, int likelihood_only, int64_t n,
                          float scale_bound, float lik_bound, float *g_inputs, float *g_scales, float *g_means,
                          pic_stream_t stream_) {
    if (n <= 0 || !inputs || !scales) return PIC_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int grid = elementwise_grid(n, 2);
    if (likelihood_only) gaussian_backward_kernel<2><<<grid, 256, 0, stream>>>(g_out, g_lik, inputs, scales, means, noise, n, scale_bound, lik_bound, g_inputs, g_scales, g_means);
    else if (noise) gaussian_backward_kernel<1><<<grid, 256, 0, stream>>>(g_out, g_lik, inputs, scales, means, noise, n, scale_bound, lik_bound, g_inputs, g_scales, g_means);
    else gaussian_backward_kernel<0><<<grid, 256, 0, stream>>>(g_out, g_lik, inputs, scales, means, noise, n, scale_bound, lik_bound, g_inputs, g_scales, g_means);
    return launch_status();
}

int pic_build_indexes(const float *scales, int64_t n, const float *scale_table, int table_len, float scale_bound,
                      int32_t *idx, pic_stream_t stream_) {
    if (n <= 0 || !scales || !scale_table || table_len < 1 || !idx) return PIC_ERR_INVALID_ARGUMENT;
    const bool vec = (n % 4 == 0) && aligned16(scales) && aligned16(idx);
    build_indexes_kernel<<<elementwise_grid(n, vec ? 4 : 1), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        scales, n, scale_table, table_len, scale_bound, idx, vec);
    return launch_status();
}

int pic_quantize(const float *inputs, const float *means, const float *noise, const float *mask, int64_t n,
                 int mode, float *out_f32, int32_t *out_i32, pic_stream_t stream_) {
    if (mode < 0 || mode > 3) return PIC_ERR_INVALID_ARGUMENT;  // ValueError in the reference
    if (n <= 0 || !inputs) return PIC_ERR_INVALID_ARGUMENT;
    if (mode == PIC_QUANTIZE_NOISE && (!noise || !out_f32)) return PIC_ERR_INVALID_ARGUMENT;
    if ((mode == PIC_QUANTIZE_DEQUANTIZE || mode == PIC_QUANTIZE_STE) && !out_f32) return PIC_ERR_INVALID_ARGUMENT;
    if (mode == PIC_QUANTIZE_SYMBOLS && !out_i32) return PIC_ERR_INVALID_ARGUMENT;
    bool vec = (n % 4 == 0);
    for (const void *q : {static_cast<const void *>(inputs), static_cast<const void *>(means), static_cast<const void *>(noise),
                          static_cast<const void *>(mask), static_cast<const void *>(out_f32), static_cast<const void *>(out_i32)})
        if (q && !aligned16(q)) vec = false;
    quantize_kernel<<<elementwise_grid(n, vec ? 4 : 1), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        inputs, means, noise, mask, n, mode, out_f32, out_i32, vec);
    return launch_status();
}

int pic_dequantize(const int32_t *symbols, const float *means, int64_t n, float *out, pic_stream_t stream_) {
    if (n <= 0 || !symbols || !out) return PIC_ERR_INVALID_ARGUMENT;
    const bool vec = (n % 4 == 0) && aligned16(symbols) && aligned16(out) && (!means || aligned16(means));
    dequantize_kernel<<<elementwise_grid(n, vec ? 4 : 1), 256, 0, static_cast<cudaStream_t>(stream_)>>>(symbols, means, n,
                                                                                                       out, vec);
    return launch_status();
}

static int launch_neighbour(int kind, const float *a, const float *b, const float *c, float *out, int64_t n,
                            pic_stream_t stream_) {
    if (n <= 0 || !a || !b || !out) return PIC_ERR_INVALID_ARGUMENT;
    bool vec = (n % 4 == 0);
    for (const void *q : {static_cast<const void *>(a), static_cast<const void *>(b), static_cast<const void *>(c),
                          static_cast<const void *>(out)}) {
        if (q && !aligned4(q)) return PIC_ERR_UNALIGNED;
        if (q && !aligned16(q)) vec = false;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int grid = elementwise_grid(n, vec ? 4 : 1);
    if (kind == 0) neighbour_kernel<0><<<grid, 256, 0, stream>>>(a, b, c, out, n, vec);
    else if (kind == 1) neighbour_kernel<1><<<grid, 256, 0, stream>>>(a, b, c, out, n, vec);
    else if (kind == 2) neighbour_kernel<2><<<grid, 256, 0, stream>>>(a, b, c, out, n, vec);
    else neighbour_kernel<3><<<grid, 256, 0, stream>>>(a, b, c, out, n, vec);
    return launch_status();
}

int pic_lrp_merge(const float *y_hat, const float *lrp, const float *base, float *out, int64_t n, pic_stream_t stream) {
    return launch_neighbour(0, y_hat, lrp, base, out, n, stream);
}
int pic_lrp_merge_backward(const float *g_out, const float *lrp, float *g_lrp, int64_t n, pic_stream_t stream) {
    return launch_neighbour(1, g_out, lrp, nullptr, g_lrp, n, stream);
}
int pic_rem_merge(const float *identity, const float *ret, const float *att_mask, float *out, int64_t n,
                  pic_stream_t stream) {
    if (!att_mask) return PIC_ERR_INVALID_ARGUMENT;
    return launch_neighbour(2, identity, ret, att_mask, out, n, stream);
}
int pic_rem_merge_backward(const float *g_out, const float *att_mask, float *g_ret, int64_t n, pic_stream_t stream) {
    return launch_neighbour(3, g_out, att_mask, nullptr, g_ret, n, stream);
}

int pic_log_sum(const float *x, int64_t n_per_unit, int64_t units, double *out, pic_stream_t stream_) {
    if (n_per_unit <= 0 || units <= 0 || units > 65535 || !x || !out) return PIC_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PIC_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * units, stream));
    int chunks = static_cast<int>((n_per_unit + 256 * 16 - 1) / (256 * 16));
    const int cap = (sm_count() * 8 + static_cast<int>(units) - 1) / static_cast<int>(units);
    if (chunks > cap) chunks = cap < 1 ? 1 : cap;
    log_sum_kernel<<<dim3(chunks, static_cast<unsigned>(units)), 256, 0, stream>>>(x, n_per_unit, out);
    return launch_status();
}

}  // extern "C"
