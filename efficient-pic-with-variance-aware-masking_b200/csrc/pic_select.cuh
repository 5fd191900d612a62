// pic_select.cuh -- exact two-order-statistic radix select on order-preserving u32 keys.
//
// Replaces the full sort inside torch.quantile (layers/channel_mask.py:41,145): finds
// a = sorted[lo] and b = sorted[hi] (hi in {lo, lo+1}) with three histogram rounds over
// 11/11/10 key bits.  The histogram lives in shared memory (warp-aggregated atomics when a
// whole warp hits one bin), the bin search is a block-wide shuffle scan.
//
// Two users:
//   * block_select(): one CTA owns the whole unit, keys staged in shared memory (fused kernel)
//   * the round kernels in pic_latent.cu: many CTAs per unit, histograms merged in global
//     memory (large units, and spatially tiled units whose histograms are all-reduced).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pic {

constexpr int kRadixRounds = 3;
constexpr int kHistBins = 2048;  // bins of round 0 and 1; round 2 uses the first 1024
__host__ __device__ constexpr int round_shift(int r) { return r == 0 ? 21 : (r == 1 ? 10 : 0); }
__host__ __device__ constexpr int round_bins(int r) { return r == 2 ? 1024 : 2048; }
__host__ __device__ constexpr uint32_t round_mask(int r) { return r == 2 ? 1023u : 2047u; }

// Histogram increment.  When every active lane of the warp targets the same bin (heavy ties,
// all-equal tiles, masked-out regions) one lane adds the population count instead of 32
// serialised same-address atomics.
__device__ __forceinline__ void hist_add(uint32_t *hist, uint32_t bin, bool active) {
    const unsigned ballot = __ballot_sync(0xffffffffu, active);
    if (ballot == 0u) return;
    const int leader = __ffs(ballot) - 1;
    const uint32_t lead_bin = __shfl_sync(0xffffffffu, bin, leader);
    const bool same = __all_sync(0xffffffffu, !active || bin == lead_bin);
    if (same) {
        if ((threadIdx.x & 31) == leader) atomicAdd(&hist[lead_bin], __popc(ballot));
    } else if (active) {
        atomicAdd(&hist[bin], 1u);
    }
}

struct BinHit {
    uint32_t bin;     // bin that contains the wanted rank
    uint32_t below;   // number of keys in lower bins
    uint32_t count;   // population of `bin`
};

// Block-wide search of the bin containing `rank` (0-based) in hist[0..nbins).
// scratch: >= 40 words of shared memory.  All threads must call; result valid in all threads.
template <int THREADS>
__device__ __forceinline__ BinHit block_find_bin(const uint32_t *hist, int nbins, uint32_t rank,
                                                 uint32_t *scratch) {
    constexpr int WARPS = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (nbins + THREADS - 1) / THREADS;
    const int begin = tid * per;
    uint32_t sum = 0;
    for (int j = 0; j < per; ++j) {
        const int bidx = begin + j;
        sum += (bidx < nbins) ? hist[bidx] : 0u;
    }
    // inclusive warp scan
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < WARPS) ? scratch[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < WARPS) scratch[lane] = winc - w;  // exclusive warp offsets
    }
    __syncthreads();
    const uint32_t excl = scratch[warp] + inc - sum;
    if (sum != 0u && excl <= rank && rank < excl + sum) {
        uint32_t run = excl;
        for (int j = 0; j < per; ++j) {
            const int bidx = begin + j;
            const uint32_t c = (bidx < nbins) ? hist[bidx] : 0u;
            if (rank < run + c) {
                scratch[34] = static_cast<uint32_t>(bidx);
                scratch[35] = run;
                scratch[36] = c;
                break;
            }
            run += c;
        }
    }
    __syncthreads();
    BinHit hit{scratch[34], scratch[35], scratch[36]};
    __syncthreads();  // scratch may be reused by the caller
    return hit;
}

// Smallest non-empty bin index strictly above `bin`, or 0xffffffff.  scratch[37] is used.
template <int THREADS>
__device__ __forceinline__ uint32_t block_next_nonempty(const uint32_t *hist, int nbins,
                                                        uint32_t bin, uint32_t *scratch) {
    if (threadIdx.x == 0) scratch[37] = 0xffffffffu;
    __syncthreads();
    uint32_t best = 0xffffffffu;
    for (int j = static_cast<int>(bin) + 1 + threadIdx.x; j < nbins; j += THREADS) {
        if (hist[j] != 0u) { best = static_cast<uint32_t>(j); break; }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0 && best != 0xffffffffu) atomicMin(&scratch[37], best);
    __syncthreads();
    const uint32_t r = scratch[37];
    __syncthreads();
    return r;
}

constexpr int kCandMax = 5120;    // candidate buffer (keys of the pivot bracket / chosen bucket)
constexpr int kSampleMax = 2048;  // sample size of the sampled-pivot select (bracket ~ 8 % of the unit)

// Key sources for the block-level passes: keys already in shared memory, or a unit's std
// values in global memory (converted on the fly; L2-resident after the first sweep).
struct SmemKeys {
    const uint32_t *k;
    __device__ __forceinline__ bool vec() const { return true; }
    __device__ __forceinline__ uint4 load4(int j) const { return reinterpret_cast<const uint4 *>(k)[j]; }
    __device__ __forceinline__ uint32_t load1(int j) const { return k[j]; }
};
struct GlobalStd {
    const float *s;
    bool vec_ok;
    __device__ __forceinline__ bool vec() const { return vec_ok; }
    __device__ __forceinline__ uint4 load4(int j) const {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(s) + j);
        return make_uint4(float_to_key(v.x), float_to_key(v.y), float_to_key(v.z), float_to_key(v.w));
    }
    __device__ __forceinline__ uint32_t load1(int j) const { return float_to_key(__ldg(s + j)); }
};

// Visits every key of the source once: f(key).  Vector path when the source allows it.
template <int THREADS, typename Src, typename F>
__device__ __forceinline__ void block_for_each_key(const Src &src, int n, F f) {
    const int tid = threadIdx.x;
    if (src.vec()) {
        const int nvec = n >> 2;
        for (int j = tid; j < nvec; j += THREADS) {
            const uint4 k = src.load4(j);
            f(k.x); f(k.y); f(k.z); f(k.w);
        }
        for (int j = (nvec << 2) + tid; j < n; j += THREADS) f(src.load1(j));
    } else {
        for (int j = tid; j < n; j += THREADS) f(src.load1(j));
    }
}

// One CTA selects the order statistics lo and hi (hi in {lo, lo+1}) among the n keys of `src`.
// hist: kHistBins words; cand: kCandMax words (unused when !COMPACT); scratch: 48 words.
// round0_ready: the caller already accumulated the round-0 histogram.
// COMPACT: after any round whose chosen bin holds <= kCandMax keys, the bin is copied into
// `cand` and the remaining rounds (and the successor search) run on the candidates only.
template <int THREADS, bool COMPACT, typename Src>
__device__ __forceinline__ void block_select(const Src &src, int n, uint32_t *hist, uint32_t *cand,
                                             uint32_t *scratch, uint32_t lo, uint32_t hi,
                                             bool round0_ready, uint32_t &a_key, uint32_t &b_key) {
    const int tid = threadIdx.x;
    uint32_t prefix = 0;        // key bits decided so far (in place)
    uint32_t rank = lo;         // rank of `a` among keys matching the prefix
    uint32_t below_total = 0;   // keys strictly below the current prefix bucket
    int cur_n = n;
    bool compacted = false;
    BinHit hit{0, 0, 0};
    const SmemKeys csrc{cand};
#pragma unroll
    for (int r = 0; r < kRadixRounds; ++r) {
        const int shift = round_shift(r);
        const int nbins = round_bins(r);
        if (!(r == 0 && round0_ready)) {
            for (int j = tid; j < nbins; j += THREADS) hist[j] = 0u;
            __syncthreads();
            const int up = (r == 0) ? 0 : shift + (r == 1 ? 11 : 10);
            const uint32_t want = (r == 0) ? 0u : (prefix >> up);
            const uint32_t mask = round_mask(r);
            auto add = [&](uint32_t k) {
                if (r == 0 || (k >> up) == want) atomicAdd(&hist[(k >> shift) & mask], 1u);
            };
            if (COMPACT && compacted) block_for_each_key<THREADS>(csrc, cur_n, add);
            else block_for_each_key<THREADS>(src, n, add);
            __syncthreads();
        }
        hit = block_find_bin<THREADS>(hist, nbins, rank, scratch);
        prefix |= hit.bin << shift;
        rank -= hit.below;
        below_total += hit.below;
        if (COMPACT && r < 2 && !compacted && hit.count <= static_cast<uint32_t>(kCandMax)) {
            if (tid == 0) scratch[40] = 0u;
            __syncthreads();
            const uint32_t want = prefix >> shift;
            block_for_each_key<THREADS>(src, n, [&](uint32_t k) {
                if ((k >> shift) == want) cand[atomicAdd(&scratch[40], 1u)] = k;
            });
            __syncthreads();
            cur_n = static_cast<int>(hit.count);
            compacted = true;
        }
    }
    a_key = prefix;
    // ranks [below_total, below_total + hit.count) all hold a
    if (hi < below_total + hit.count) {
        b_key = a_key;
        return;
    }
    // successor of a: smallest key > a.  It lies in the candidate bucket unless a is its maximum.
    for (int pass = 0; pass < 2; ++pass) {
        if (tid == 0) scratch[38] = 0xffffffffu;
        __syncthreads();
        uint32_t best = 0xffffffffu;
        auto upd = [&](uint32_t k) { if (k > a_key) best = min(best, k); };
        if (COMPACT && compacted) block_for_each_key<THREADS>(csrc, cur_n, upd);
        else block_for_each_key<THREADS>(src, n, upd);
        best = __reduce_min_sync(0xffffffffu, best);
        if ((tid & 31) == 0 && best != 0xffffffffu) atomicMin(&scratch[38], best);
        __syncthreads();
        b_key = scratch[38];
        __syncthreads();
        if (b_key != 0xffffffffu || !(COMPACT && compacted)) break;
        compacted = false;   // a was the bucket maximum: look at the whole unit (rare)
    }
}

// ---------------------------------------------------------------------------------------------
// Small-set helpers used by the sampled-pivot select (keys in shared memory, m <= kCandMax)
// ---------------------------------------------------------------------------------------------
// Two-level histogram: `fine` has 2^w bins, `coarse` has min(64, 2^w) bins (coarse = fine >> cs).
// One warp finds the bin holding `rank` with two shuffle scans -- no block-wide scan, no barrier.
constexpr int kCoarseBins = 64;
__host__ __device__ constexpr int coarse_shift(int w) { return w > 6 ? w - 6 : 0; }

// Must be called by all 32 lanes of one warp.  Result valid in every lane.
__device__ __forceinline__ BinHit warp_find_bin(const uint32_t *fine, const uint32_t *coarse, int w,
                                                uint32_t rank) {
    const int lane = threadIdx.x & 31;
    const int nb = 1 << w;
    const int nc = nb < kCoarseBins ? nb : kCoarseBins;
    const int fpc = nb / nc;  // fine bins per coarse bin: 32, 16, ... or 1
    // level 1: two coarse bins per lane
    const uint32_t c0 = (2 * lane < nc) ? coarse[2 * lane] : 0u;
    const uint32_t c1 = (2 * lane + 1 < nc) ? coarse[2 * lane + 1] : 0u;
    uint32_t inc = c0 + c1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    const uint32_t excl = inc - (c0 + c1);
    const unsigned owner = __ballot_sync(0xffffffffu, (c0 + c1) != 0u && excl <= rank && rank < inc);
    const int ol = owner ? (__ffs(owner) - 1) : 0;
    const uint32_t o_excl = __shfl_sync(0xffffffffu, excl, ol);
    const uint32_t o_c0 = __shfl_sync(0xffffffffu, c0, ol);
    const bool second = rank >= o_excl + o_c0;
    const int cb = 2 * ol + (second ? 1 : 0);
    uint32_t below = o_excl + (second ? o_c0 : 0u);
    // level 2: the fpc (<= 32) fine bins of coarse bin cb
    const uint32_t f = (lane < fpc) ? fine[cb * fpc + lane] : 0u;
    uint32_t finc = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, finc, d);
        if (lane >= d) finc += t;
    }
    const uint32_t fexcl = below + finc - f;
    const unsigned fo = __ballot_sync(0xffffffffu, f != 0u && fexcl <= rank && rank < fexcl + f);
    const int fl = fo ? (__ffs(fo) - 1) : 0;
    BinHit hit;
    hit.bin = static_cast<uint32_t>(cb * fpc + fl);
    hit.below = __shfl_sync(0xffffffffu, fexcl, fl);
    hit.count = __shfl_sync(0xffffffffu, f, fl);
    return hit;
}

// scratch layout used by the helpers below
constexpr int kScrHitA = 44;      // 3 words
constexpr int kScrHitB = 48;      // 3 words
constexpr int kScrCoarseA = 64;   // 64 words
constexpr int kScrCoarseB = 128;  // 64 words

// Exact order statistics lo / hi (hi in {lo, lo+1}) of keys[0..m), all of which lie in
// [base, base + 2^bits).  Works on normalised keys k - base, most significant digit first,
// <= 11 bits per round, so a narrow bracket needs 1-2 rounds instead of 3.
// hist: kHistBins words; scratch: 192 words.
template <int THREADS>
__device__ __forceinline__ void block_select_norm(const uint32_t *keys, int m, uint32_t base, int bits,
                                                  uint32_t *hist, uint32_t *scratch, uint32_t lo,
                                                  uint32_t hi, uint32_t &a_key, uint32_t &b_key) {
    const int tid = threadIdx.x;
    uint32_t *coarse = scratch + kScrCoarseA;
    uint32_t prefix = 0, rank = lo, below_total = 0, count = static_cast<uint32_t>(m);
    int top = bits;
    while (top > 0) {
        const int w = top < 11 ? top : 11;
        const int shift = top - w;
        const int nb = 1 << w;
        const int cs = coarse_shift(w);
        for (int j = tid; j < nb; j += THREADS) hist[j] = 0u;
        if (tid < kCoarseBins) coarse[tid] = 0u;
        __syncthreads();
        const uint32_t want = (top >= 32) ? 0u : (prefix >> top);
        for (int j = tid; j < m; j += THREADS) {
            const uint32_t kn = keys[j] - base;
            if (top >= 32 || (kn >> top) == want) {
                const uint32_t d = (kn >> shift) & (nb - 1);
                atomicAdd(&hist[d], 1u);
                atomicAdd(&coarse[d >> cs], 1u);
            }
        }
        __syncthreads();
        if (tid < 32) {
            const BinHit h = warp_find_bin(hist, coarse, w, rank);
            if (tid == 0) { scratch[kScrHitA] = h.bin; scratch[kScrHitA + 1] = h.below; scratch[kScrHitA + 2] = h.count; }
        }
        __syncthreads();
        const uint32_t bin = scratch[kScrHitA], below = scratch[kScrHitA + 1];
        count = scratch[kScrHitA + 2];
        prefix |= bin << shift;
        rank -= below;
        below_total += below;
        top = shift;
    }
    a_key = base + prefix;
    if (hi < below_total + count) {
        b_key = a_key;
        return;
    }
    __syncthreads();
    if (tid == 0) scratch[38] = 0xffffffffu;
    __syncthreads();
    uint32_t best = 0xffffffffu;
    for (int j = tid; j < m; j += THREADS) {
        const uint32_t k = keys[j];
        if (k > a_key) best = min(best, k);
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((tid & 31) == 0 && best != 0xffffffffu) atomicMin(&scratch[38], best);
    __syncthreads();
    b_key = scratch[38];
    __syncthreads();
}

// Two ranks r_lo <= r_hi of keys[0..m) resolved TOGETHER to 22-bit buckets (two 11-bit rounds,
// one shared pass per round, the two searches run in two warps concurrently): returns the lower
// edge of r_lo's bucket and the upper edge of r_hi's bucket -- a bracket containing both order
// statistics.  hist: 2 * kHistBins words; scratch: 192 words.
template <int THREADS>
__device__ __forceinline__ void block_bracket_pair(const uint32_t *keys, int m, uint32_t *hist,
                                                   uint32_t *scratch, uint32_t r_lo, uint32_t r_hi,
                                                   uint32_t &lo_edge, uint32_t &hi_edge) {
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t *hA = hist, *hB = hist + kHistBins;
    uint32_t *cA = scratch + kScrCoarseA, *cB = scratch + kScrCoarseB;
    // round 0: one histogram serves both ranks
    for (int j = tid; j < kHistBins; j += THREADS) hA[j] = 0u;
    if (tid < kCoarseBins) cA[tid] = 0u;
    __syncthreads();
    for (int j = tid; j < m; j += THREADS) {
        const uint32_t d = keys[j] >> 21;
        atomicAdd(&hA[d], 1u);
        atomicAdd(&cA[d >> 5], 1u);
    }
    __syncthreads();
    if (warp < 2) {
        const BinHit h = warp_find_bin(hA, cA, 11, warp == 0 ? r_lo : r_hi);
        const int o = warp == 0 ? kScrHitA : kScrHitB;
        if ((tid & 31) == 0) { scratch[o] = h.bin; scratch[o + 1] = h.below; }
    }
    __syncthreads();
    const uint32_t a0 = scratch[kScrHitA], a0_below = scratch[kScrHitA + 1];
    const uint32_t b0 = scratch[kScrHitB], b0_below = scratch[kScrHitB + 1];
    // round 1: bits [10, 21) within each rank's round-0 bucket
    for (int j = tid; j < 2 * kHistBins; j += THREADS) hist[j] = 0u;
    if (tid < 2 * kCoarseBins) cA[tid] = 0u;   // cA and cB are contiguous
    __syncthreads();
    for (int j = tid; j < m; j += THREADS) {
        const uint32_t k = keys[j];
        const uint32_t d0 = k >> 21, d1 = (k >> 10) & 2047u;
        if (d0 == a0) { atomicAdd(&hA[d1], 1u); atomicAdd(&cA[d1 >> 5], 1u); }
        if (d0 == b0) { atomicAdd(&hB[d1], 1u); atomicAdd(&cB[d1 >> 5], 1u); }
    }
    __syncthreads();
    if (warp < 2) {
        const BinHit h = (warp == 0) ? warp_find_bin(hA, cA, 11, r_lo - a0_below)
                                     : warp_find_bin(hB, cB, 11, r_hi - b0_below);
        if ((tid & 31) == 0) scratch[(warp == 0 ? kScrHitA : kScrHitB) + 2] = h.bin;
    }
    __syncthreads();
    lo_edge = (a0 << 21) | (scratch[kScrHitA + 2] << 10);
    hi_edge = (b0 << 21) | (scratch[kScrHitB + 2] << 10) | 1023u;
    __syncthreads();
}

}  // namespace pic
