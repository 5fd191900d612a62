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

constexpr int kCandMax = 4096;  // candidate buffer (keys of the chosen round-0/1 bucket)

// Histogram of digit `r` over keys[0..n) restricted to keys whose higher bits equal those of
// `prefix`.  Plain shared-memory atomics: after round 0 the matching keys are a small subset
// (or the whole, already small, candidate buffer).
template <int THREADS>
__device__ __forceinline__ void block_hist_pass(const uint32_t *keys, int n, uint32_t *hist, int r,
                                                uint32_t prefix) {
    const int shift = round_shift(r);
    const int up = (r == 0) ? 32 : shift + (r == 1 ? 11 : 10);
    const uint32_t want = (r == 0) ? 0u : (prefix >> up);
    const uint32_t mask = round_mask(r);
    const int tid = threadIdx.x;
    const int nvec = n >> 2;
    const uint4 *k4 = reinterpret_cast<const uint4 *>(keys);
    for (int j = tid; j < nvec; j += THREADS) {
        const uint4 k = k4[j];
        const uint32_t kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (r == 0 || (kk[e] >> up) == want) atomicAdd(&hist[(kk[e] >> shift) & mask], 1u);
    }
    for (int j = (nvec << 2) + tid; j < n; j += THREADS) {
        const uint32_t k = keys[j];
        if (r == 0 || (k >> up) == want) atomicAdd(&hist[(k >> shift) & mask], 1u);
    }
}

// One CTA selects the order statistics lo and hi (hi in {lo, lo+1}) among n keys held in
// shared memory.  hist: kHistBins words; cand: kCandMax words; scratch: 48 words.
// round0_ready: the caller accumulated the round-0 histogram while staging the keys.
// After any round whose chosen bin holds <= kCandMax keys the bin is compacted into `cand`
// and the remaining rounds (and the successor search) run on the candidates only.
template <int THREADS>
__device__ __forceinline__ void block_select(const uint32_t *keys, int n, uint32_t *hist,
                                             uint32_t *cand, uint32_t *scratch, uint32_t lo,
                                             uint32_t hi, bool round0_ready, uint32_t &a_key,
                                             uint32_t &b_key) {
    const int tid = threadIdx.x;
    uint32_t prefix = 0;        // key bits decided so far (in place)
    uint32_t rank = lo;         // rank of `a` among keys matching the prefix
    uint32_t below_total = 0;   // keys strictly below the current prefix bucket
    const uint32_t *cur = keys;
    int cur_n = n;
    bool compacted = false;
    BinHit hit{0, 0, 0};
#pragma unroll
    for (int r = 0; r < kRadixRounds; ++r) {
        const int shift = round_shift(r);
        const int nbins = round_bins(r);
        if (!(r == 0 && round0_ready)) {
            for (int j = tid; j < nbins; j += THREADS) hist[j] = 0u;
            __syncthreads();
            block_hist_pass<THREADS>(cur, cur_n, hist, r, prefix);
            __syncthreads();
        }
        hit = block_find_bin<THREADS>(hist, nbins, rank, scratch);
        prefix |= hit.bin << shift;
        rank -= hit.below;
        below_total += hit.below;
        if (r < 2 && !compacted && hit.count <= static_cast<uint32_t>(kCandMax)) {
            // compact the chosen bucket: keys whose top (32 - shift) bits equal the prefix's
            if (tid == 0) scratch[40] = 0u;
            __syncthreads();
            const uint32_t want = prefix >> shift;
            const int nvec = n >> 2;
            const uint4 *k4 = reinterpret_cast<const uint4 *>(keys);
            for (int j = tid; j < nvec; j += THREADS) {
                const uint4 k = k4[j];
                const uint32_t kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if ((kk[e] >> shift) == want) cand[atomicAdd(&scratch[40], 1u)] = kk[e];
            }
            for (int j = (nvec << 2) + tid; j < n; j += THREADS) {
                const uint32_t k = keys[j];
                if ((k >> shift) == want) cand[atomicAdd(&scratch[40], 1u)] = k;
            }
            __syncthreads();
            cur = cand;
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
        for (int j = tid; j < cur_n; j += THREADS) {
            const uint32_t k = cur[j];
            if (k > a_key) best = min(best, k);
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if ((tid & 31) == 0 && best != 0xffffffffu) atomicMin(&scratch[38], best);
        __syncthreads();
        b_key = scratch[38];
        __syncthreads();
        if (b_key != 0xffffffffu || !compacted) break;
        cur = keys;   // a was the bucket maximum: look at the whole unit (rare)
        cur_n = n;
        compacted = false;
    }
}

}  // namespace pic
