// pic_tma_select.cu -- threshold select with TMA-staged tiles and warp-specialised phases (sm_100a).
//
// Replaces torch.quantile's full sort (layers/channel_mask.py:142-149, ProgMask 18-49) for units of
// 32768 <= n <= ~80 K elements (the Kodak-shape slices): the exact order statistics a = sorted[lo],
// b = sorted[hi] and thr = lerp(a, b, w) with ATen's f32 rank arithmetic (pic_math.cuh).
//
// One persistent CTA per SM, units strided over CTAs, three roles:
//   producer warp   streams the unit's std through a deep ring of shared-memory stages with cp.async.bulk
//                   (1-D TMA, mbarrier-completed); it runs ahead across unit boundaries, so HBM stays busy
//                   whatever the other warps are doing,
//   sweep warps     (20) do nothing but classify: per element (closed bracket) d = x - mid (packed add),
//                   below-count = FSET(d < -hw) + packed add, candidate test |d| <= hw (one FSETP), predicated
//                   store + IMAD pointer bump into the thread's PRIVATE candidate list -- no divergence, no
//                   scan, no atomic; NaN/inf detection rides on a packed FMA.  A thread whose list fills up
//                   continues in 8-entry blocks from a small shared pool (rare).  Lists are double-buffered,
//                   so the sweep of unit u+1 starts the moment the sweep of unit u ends,
//   helper warps    (8) run the latency-bound phases off the critical path: the pivots of unit u+1 (sample of
//                   <= 2048 keys gathered into registers, ONE 2048-bin histogram of the top key bits, the two
//                   sample ranks located by one warp each and interpolated inside their buckets) and the final
//                   phase of unit u (one 2048-bin histogram of the ~8 % candidates on a digit that is LINEAR in
//                   the value -- monotone, so every bin is an interval --, one warp finds the bin of the wanted
//                   rank, its few members are ranked exactly by key inside that warp; crowded bins (heavy ties)
//                   take radix rounds restricted to the bin).
// The pivots only have to bracket the answer: the final phase validates them (c_below <= lo, hi < c_below +
// c_cand), so any deterministic estimate is exact-safe.  A bracket that missed or a pool that ran dry falls
// back to radix rounds over the whole unit (exact either way; counted in the fallback counter).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <type_traits>

#include "pic_math.cuh"
#include "pic_fast.cuh"
#include "pic_params.h"

namespace pic {

__device__ unsigned long long g_tma_fallback_units = 0ull;
__device__ unsigned long long g_tma_sampled_units = 0ull;
#ifdef PIC_PHASE_TIMING
__device__ long long g_tma_phase_clk[16];
// role timers of block 0: TMA_T(i, who, expr) adds the cycles of expr to slot i when `who` is true
#define TMA_T(i, who, expr) do { const long long t0__ = clock64(); expr; if (who) g_tma_phase_clk[i] += clock64() - t0__; } while (0)
#else
#define TMA_T(i, who, expr) do { expr; } while (0)
#endif
#ifdef PIC_PHASE_TIMING
#define TMA_STAMP(i) do { if (threadIdx.x == blockDim.x - 256 && blockIdx.x == 0) { const long long t__ = clock64(); g_tma_phase_clk[i] += t__ - stamp__; stamp__ = t__; } } while (0)
#define TMA_STAMP0() long long stamp__ = clock64()
#else
#define TMA_STAMP(i) do {} while (0)
#define TMA_STAMP0() do {} while (0)
#endif

// ---------------------------------------------------------------------------------------------
// mbarrier / bulk-copy wrappers (shared::cta addresses as 32-bit registers)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0u;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time before it answers)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// for waits that may last microseconds: polling warps would take issue slots from the working ones
template <unsigned NS>
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(NS);
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA unit, no tensor map)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ float max3_nan(float a, float b, float c) {
    float r;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// hist[bin] += 1 if ok -- predicated, never a branch (keeps the callers' unrolled chains independent)
__device__ __forceinline__ void red_inc_if(uint32_t addr, bool ok) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; @p red.shared.add.u32 [%0], 1; }" ::"r"(addr), "r"(static_cast<uint32_t>(ok)) : "memory");
}

// Candidate pushes of the sweep: predicated store + pointer bump.  The bump is an IMAD (FMA pipe) so that the
// ALU pipe only carries the compares; `one` is an opaque 1 (kernel parameter).
// closed bracket: candidate iff |d| <= hw with d = x - mid
__device__ __forceinline__ void push_if_abs_le(uint32_t &addr, float d, float hw, float x, uint32_t stride, uint32_t one) {
    asm volatile("{ .reg .pred q; .reg .f32 t; abs.f32 t, %1; setp.le.f32 q, t, %2; @q st.shared.f32 [%0], %3; @q mad.lo.u32 %0, %4, %5, %0; }"
                 : "+r"(addr) : "f"(d), "f"(hw), "f"(x), "r"(stride), "r"(one) : "memory");
}
// open-ended bracket: candidate iff lo <= x <= hi (false for NaN)
__device__ __forceinline__ void push_if_in_range(uint32_t &addr, float x, float lo, float hi, uint32_t stride, uint32_t one) {
    asm volatile("{ .reg .pred p, q; setp.ge.f32 p, %1, %2; setp.le.and.f32 q, %1, %3, p; @q st.shared.f32 [%0], %1; @q mad.lo.u32 %0, %4, %5, %0; }"
                 : "+r"(addr) : "f"(x), "f"(lo), "f"(hi), "r"(stride), "r"(one) : "memory");
}

// Four candidate pushes at once.  The four store addresses are computed first, each in its own register, and the
// stores follow: a store keeps its address register busy until the memory pipe has read it, so bumping that same
// register right after the store (one element at a time) stalls the warp on every element.
__device__ __forceinline__ void push4_abs_le(uint32_t &addr, float d0, float d1, float d2, float d3, float hw, const float4 &v,
                                             uint32_t stride, uint32_t one) {
    asm volatile(
        "{ .reg .pred q0, q1, q2, q3; .reg .f32 t0, t1, t2, t3; .reg .u32 a1, a2, a3;\n"
        "  abs.f32 t0, %1; abs.f32 t1, %2; abs.f32 t2, %3; abs.f32 t3, %4;\n"
        "  setp.le.f32 q0, t0, %5; setp.le.f32 q1, t1, %5; setp.le.f32 q2, t2, %5; setp.le.f32 q3, t3, %5;\n"
        "  mov.u32 a1, %0; @q0 mad.lo.u32 a1, %10, %11, %0;\n"
        "  mov.u32 a2, a1; @q1 mad.lo.u32 a2, %10, %11, a1;\n"
        "  mov.u32 a3, a2; @q2 mad.lo.u32 a3, %10, %11, a2;\n"
        "  @q0 st.shared.f32 [%0], %6; @q1 st.shared.f32 [a1], %7; @q2 st.shared.f32 [a2], %8; @q3 st.shared.f32 [a3], %9;\n"
        "  mov.u32 %0, a3; @q3 mad.lo.u32 %0, %10, %11, a3; }"
        : "+r"(addr)
        : "f"(d0), "f"(d1), "f"(d2), "f"(d3), "f"(hw), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(stride), "r"(one)
        : "memory");
}
__device__ __forceinline__ void push4_in_range(uint32_t &addr, const float4 &v, float lo, float hi, uint32_t stride, uint32_t one) {
    asm volatile(
        "{ .reg .pred p, q0, q1, q2, q3; .reg .u32 a1, a2, a3;\n"
        "  setp.ge.f32 p, %1, %5; setp.le.and.f32 q0, %1, %6, p;\n"
        "  setp.ge.f32 p, %2, %5; setp.le.and.f32 q1, %2, %6, p;\n"
        "  setp.ge.f32 p, %3, %5; setp.le.and.f32 q2, %3, %6, p;\n"
        "  setp.ge.f32 p, %4, %5; setp.le.and.f32 q3, %4, %6, p;\n"
        "  mov.u32 a1, %0; @q0 mad.lo.u32 a1, %7, %8, %0;\n"
        "  mov.u32 a2, a1; @q1 mad.lo.u32 a2, %7, %8, a1;\n"
        "  mov.u32 a3, a2; @q2 mad.lo.u32 a3, %7, %8, a2;\n"
        "  @q0 st.shared.f32 [%0], %1; @q1 st.shared.f32 [a1], %2; @q2 st.shared.f32 [a2], %3; @q3 st.shared.f32 [a3], %4;\n"
        "  mov.u32 %0, a3; @q3 mad.lo.u32 %0, %7, %8, a3; }"
        : "+r"(addr)
        : "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "f"(lo), "f"(hi), "r"(stride), "r"(one)
        : "memory");
}

constexpr int kPoolBlockWords = 12;   // 8 usable entries + 4 guard (a float4 is classified between capacity checks)
constexpr int kSmallList = 32;        // a final bin with at most this many members is ranked inside one warp

struct TmaSelectConfig {
    int stages;        // ring depth
    int k0;            // private candidate entries per sweep thread (incl. 4 guard entries)
    int pool_blocks;   // shared overflow blocks of kPoolBlockWords words (per list buffer)
    uint32_t one;      // 1 (opaque to the compiler: keeps the pointer bump an IMAD)
};

// mailbox words (one mailbox per list buffer)
enum : int { kMbPlo = 0, kMbPhi, kMbMid, kMbHw, kMbClosed, kMbBelow, kMbNan, kMbOvf, kMbPoolNext, kMbMaxK, kMbWords = 16 };

// shared-memory carve-up (bytes), every part a multiple of 16
template <int NCT, int VPT>
struct TmaSmem {
    static constexpr uint32_t kStageBytes = NCT * VPT * 16;
    __host__ __device__ static size_t ring(const TmaSelectConfig &c) { return size_t(c.stages) * kStageBytes; }
    __host__ __device__ static size_t priv(const TmaSelectConfig &c) { return size_t(c.k0) * NCT * 4; }
    __host__ __device__ static size_t pool(const TmaSelectConfig &c) { return size_t(c.pool_blocks) * kPoolBlockWords * 4; }
    __host__ __device__ static size_t pool_cnt(const TmaSelectConfig &c) { return (size_t(c.pool_blocks) * 4 + 15) / 16 * 16; }
    static constexpr size_t kCnt0 = size_t(NCT) * 4;
    __host__ __device__ static size_t buf(const TmaSelectConfig &c) { return priv(c) + pool(c) + pool_cnt(c) + kCnt0; }
    static constexpr size_t kHist = (kHistBins + 8) * 4;
    static constexpr size_t kScratch = 256 * 4;
    static constexpr size_t kMail = 2 * kMbWords * 4;
    static constexpr size_t kPivot = kHistBins * 4 + 64;          // pivot group's histogram + 16 scratch words
    static constexpr size_t kLocalQ = 128 * 4;                     // q01 of this CTA's first units
    __host__ __device__ static size_t bars(const TmaSelectConfig &c) { return size_t(c.stages) * 16 + 6 * 8 + 16; }
    __host__ __device__ static size_t total(const TmaSelectConfig &c) {
        return ring(c) + 2 * buf(c) + kHist + kScratch + kMail + kPivot + kLocalQ + bars(c);
    }
};

// helper scratch words (warp_find_bin's callers use kScrHitA / kScrCoarseA of pic_select.cuh)
constexpr int kScrMinAbove = 52, kScrBinMin = 53, kScrBinMax = 54, kScrListLen = 55, kScrA = 56, kScrB = 57, kScrTmp = 58;
constexpr int kScrList = 192;   // kSmallList words
constexpr int kScrDummy = 224;  // 32 words

__device__ __forceinline__ int sample_size(int n) {
    int S = n >> 4;
    S = S < 1024 ? 1024 : (S > kSampleMax ? kSampleMax : S);
    return S & ~3;
}

// Interpolated key of sample rank position `t` (may be fractional / out of range) inside a bucket of 2^21 keys
// holding `count` sample keys.
__device__ __forceinline__ uint32_t bucket_interp(uint32_t bin, float t, uint32_t count) {
    float f = t / static_cast<float>(count);
    f = fminf(fmaxf(f, 0.0f), 1.0f);
    uint32_t off = static_cast<uint32_t>(f * 2097152.0f);
    off = off > 0x1fffffu ? 0x1fffffu : off;
    uint32_t k = (bin << 21) | off;
    // keep the pivot a non-NaN float: keys outside [key(-inf), key(+inf)] are NaN bit patterns
    k = k < 0x007fffffu ? 0x007fffffu : (k > 0xff800000u ? 0xff800000u : k);
    return k;
}

// One warp finds the bin of 0-based rank `rank` in a 2048-bin histogram without a coarse level built by others:
// lane l sums bins [64 l, 64 l + 64) with 16 conflict-free 128-bit loads (the start is rotated by the lane index),
// one scan picks the lane, its 64 bins are scanned two per lane.  `total` = population of the histogram.
__device__ __forceinline__ BinHit warp_find2048(const uint32_t *hist, uint32_t rank, uint32_t &total) {
    const int lane = threadIdx.x & 31;
    const uint4 *h4 = reinterpret_cast<const uint4 *>(hist) + 16 * lane;
    uint32_t seg = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint4 v = h4[(j + lane) & 15];
        seg += (v.x + v.y) + (v.z + v.w);
    }
    uint32_t inc = seg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    const uint32_t excl = inc - seg;
    const unsigned owner = __ballot_sync(0xffffffffu, seg != 0u && excl <= rank && rank < inc);
    const int ol = owner ? (__ffs(owner) - 1) : 0;
    const uint32_t below0 = __shfl_sync(0xffffffffu, excl, ol);
    const uint2 f = reinterpret_cast<const uint2 *>(hist)[32 * ol + lane];    // bins 64 ol + 2 lane, + 1
    uint32_t finc = f.x + f.y;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, finc, d);
        if (lane >= d) finc += t;
    }
    const uint32_t fexcl = below0 + finc - (f.x + f.y);
    const unsigned fo = __ballot_sync(0xffffffffu, (f.x + f.y) != 0u && fexcl <= rank && rank < fexcl + f.x + f.y);
    const int fl = fo ? (__ffs(fo) - 1) : 0;
    const uint32_t pe = __shfl_sync(0xffffffffu, fexcl, fl);
    const uint32_t fx = __shfl_sync(0xffffffffu, f.x, fl), fy = __shfl_sync(0xffffffffu, f.y, fl);
    const bool second = rank >= pe + fx;
    BinHit hit;
    hit.bin = static_cast<uint32_t>(64 * ol + 2 * fl + (second ? 1 : 0));
    hit.below = pe + (second ? fx : 0u);
    hit.count = second ? fy : fx;
    return hit;
}

// NCT sweep threads + one producer warp + NHT helper threads per CTA; VPT float4 per sweep thread per ring stage.
template <int NCT, int NHT, int VPT>
__global__ void __launch_bounds__(NCT + 32 + NHT, 1) select_tma_kernel(const SliceParams p, const TmaSelectConfig cfg) {
    extern __shared__ __align__(128) unsigned char dyn[];
    using L = TmaSmem<NCT, VPT>;
    constexpr uint32_t kStageBytes = L::kStageBytes;
    constexpr int kStageVec = NCT * VPT;
    constexpr uint32_t kPrivStride = NCT * 4;
    constexpr int kSweepWarps = NCT / 32, kHelperWarps = NHT / 32;
    unsigned char *ring = dyn;
    unsigned char *buf0 = dyn + L::ring(cfg);
    const size_t buf_bytes = L::buf(cfg);
    uint32_t *hist = reinterpret_cast<uint32_t *>(buf0 + 2 * buf_bytes);
    uint32_t *scratch = hist + (kHistBins + 8);
    uint32_t *coarse = scratch + kScrCoarseA;
    uint32_t *mail0 = scratch + 256;
    uint32_t *hist_p = mail0 + 2 * kMbWords;                 // pivot group
    uint32_t *scratch_p = hist_p + kHistBins;
    float *local_q = reinterpret_cast<float *>(scratch_p + 16);
    const uint32_t bars = smem_u32(local_q + 128);   // full[s] at +16 s, empty[s] at +16 s + 8, then the unit barriers
    const uint32_t ring_u32 = smem_u32(ring);
    const int stages = cfg.stages;
    const uint32_t ubar = bars + 16 * stages;               // piv_ready[b] +8 b, sweep_done[b] +16 + 8 b, final_done[b] +32 + 8 b
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = static_cast<int>(p.n);
    const int nvec = n >> 2;
    const int nchunks = (nvec + kStageVec - 1) / kStageVec;
    const int64_t stride_u = gridDim.x;
    const uint64_t pol_last = policy_evict_last();

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bars + 16 * s, 1);
            mbar_init(bars + 16 * s + 8, kSweepWarps);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(ubar + 8 * b, 1);
            mbar_init(ubar + 16 + 8 * b, kSweepWarps);
            mbar_init(ubar + 32 + 8 * b, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // q01 of this CTA's first 128 units: the roles below look them up without a global-memory round trip per unit
    if (tid < 128) {
        const int64_t uq = blockIdx.x + static_cast<int64_t>(tid) * stride_u;
        local_q[tid] = (p.q01_per_unit && uq < p.units) ? p.q01_per_unit[uq] : p.q01;
    }
    __syncthreads();

    // this CTA's i-th unit, its q01, and the first i' >= i whose unit needs a select (or the end of the sequence)
    auto unit_at = [&](int64_t i) { return blockIdx.x + i * stride_u; };
    auto q_at = [&](int64_t i) {
        if (!p.q01_per_unit) return p.q01;
        return i < 128 ? local_q[i] : p.q01_per_unit[unit_at(i)];
    };
    auto unit_std = [&](int64_t u) { return p.std + ((p.repeat > 1) ? u / p.repeat : u) * p.n; };
    auto next_stream = [&](int64_t i) {
        while (unit_at(i) < p.units && unit_mode(q_at(i)) != kModeThreshold) ++i;
        return i;
    };
    auto buf_priv = [&](int b) { return reinterpret_cast<uint32_t *>(buf0 + b * buf_bytes); };
    auto buf_pool = [&](int b) { return reinterpret_cast<uint32_t *>(buf0 + b * buf_bytes + L::priv(cfg)); };
    auto buf_pool_cnt = [&](int b) { return reinterpret_cast<uint32_t *>(buf0 + b * buf_bytes + L::priv(cfg) + L::pool(cfg)); };
    auto buf_cnt0 = [&](int b) {
        return reinterpret_cast<uint32_t *>(buf0 + b * buf_bytes + L::priv(cfg) + L::pool(cfg) + L::pool_cnt(cfg));
    };

    // ======================================= producer warp ===========================================
    if (warp == kSweepWarps) {
        if (lane == 0) {
            int ps = 0;
            uint32_t ppar = 1;   // parity of the phase that must have completed before a stage is refilled
            for (int64_t i = next_stream(0); unit_at(i) < p.units; i = next_stream(i + 1)) {
                const float4 *src = reinterpret_cast<const float4 *>(unit_std(unit_at(i)));
                for (int c = 0; c < nchunks; ++c) {
                    const int first = c * kStageVec;
                    const uint32_t bytes = static_cast<uint32_t>(min(kStageVec, nvec - first)) * 16u;
                    TMA_T(7, blockIdx.x == 0, mbar_wait_backoff<32>(bars + 16 * ps + 8, ppar));           // fresh barrier: passes at once
                    mbar_expect_tx(bars + 16 * ps, bytes);
                    bulk_g2s(ring_u32 + ps * kStageBytes, src + first, bytes, bars + 16 * ps, pol_last);
                    if (++ps == stages) { ps = 0; ppar ^= 1u; }
                }
            }
        }
        return;
    }

    // ======================================== sweep warps ============================================
    if (warp < kSweepWarps) {
        const uint32_t one = cfg.one;
        const uint32_t pool_blocks = static_cast<uint32_t>(cfg.pool_blocks);
        int cs = 0;
        uint32_t cpar = 0;
        uint32_t seq = 0;
        for (int64_t i = next_stream(0); unit_at(i) < p.units; i = next_stream(i + 1), ++seq) {
            const int b = static_cast<int>(seq & 1u);
            const uint32_t upar = (seq >> 1) & 1u;
            uint32_t *mail = mail0 + b * kMbWords;
            TMA_T(0, tid == 0 && blockIdx.x == 0, mbar_wait_backoff<64>(ubar + 32 + 8 * b, upar ^ 1u));     // the helpers are done with this buffer's previous unit
            TMA_T(1, tid == 0 && blockIdx.x == 0, mbar_wait_backoff<64>(ubar + 8 * b, upar));               // pivots of this unit
            const float plo_f = __uint_as_float(mail[kMbPlo]), phi_f = __uint_as_float(mail[kMbPhi]);
            const float mid = __uint_as_float(mail[kMbMid]), hw = __uint_as_float(mail[kMbHw]);
            const bool closed = mail[kMbClosed] != 0u;
            const uint32_t slot0 = smem_u32(buf_priv(b)) + tid * 4;
            const uint32_t pool_u32 = smem_u32(buf_pool(b));
            uint32_t *pool_cnt = buf_pool_cnt(b);
            uint32_t addr = slot0, lim = slot0 + static_cast<uint32_t>(cfg.k0 - 4) * kPrivStride, stride = kPrivStride, c0 = 0;
            int cur = -1;                 // -1: private list, else the pool block being filled
            bool ovf = false;
            f2 below2 = pk(0.0f, 0.0f);
            auto switch_block = [&]() {   // rare: the current block is (nearly) full
                if (cur < 0) c0 = (addr - slot0) / kPrivStride;
                else pool_cnt[cur] = (addr - (pool_u32 + cur * (kPoolBlockWords * 4))) >> 2;
                uint32_t nb = atomicAdd(&mail[kMbPoolNext], 1u);
                if (nb >= pool_blocks) { ovf = true; nb = pool_blocks - 1; }
                cur = static_cast<int>(nb);
                addr = pool_u32 + nb * (kPoolBlockWords * 4);
                lim = addr + (kPoolBlockWords - 4) * 4;
                stride = 4;
            };
            bool flagged = false;         // NaN (open: exact; closed: NaN or +-inf, re-checked by the helpers)
            auto run_sweep = [&](auto closed_tag) {
                constexpr bool CLOSED = decltype(closed_tag)::value;
                const f2 nmid2 = pk(-mid, -mid), zero2 = pk(0.0f, 0.0f);
                const float nhw = -hw;
                f2 nan2 = pk(0.0f, 0.0f);
                float runmax = -INFINITY;
                auto classify4 = [&](const float4 &v) {
                    if (CLOSED) {
                        const f2 v01 = pk(v.x, v.y), v23 = pk(v.z, v.w);
                        float d0, d1, d2, d3;
                        unpk(add2(v01, nmid2), d0, d1);
                        unpk(add2(v23, nmid2), d2, d3);
                        nan2 = fma2(v01, zero2, nan2);      // NaN or +-inf poisons the accumulator
                        nan2 = fma2(v23, zero2, nan2);
                        below2 = add2(below2, pk(fset_lt(d0, nhw), fset_lt(d1, nhw)));   // exact: counts << 2^24
                        below2 = add2(below2, pk(fset_lt(d2, nhw), fset_lt(d3, nhw)));
                        push4_abs_le(addr, d0, d1, d2, d3, hw, v, stride, one);
                    } else {
                        runmax = max3_nan(runmax, v.x, v.y);
                        runmax = max3_nan(runmax, v.z, v.w);
                        below2 = add2(below2, pk(fset_lt(v.x, plo_f), fset_lt(v.y, plo_f)));
                        below2 = add2(below2, pk(fset_lt(v.z, plo_f), fset_lt(v.w, plo_f)));
                        push4_in_range(addr, v, plo_f, phi_f, stride, one);
                    }
                    if (addr > lim) switch_block();
                };
                for (int c = 0; c < nchunks; ++c) {
                    TMA_T(3, tid == 0 && blockIdx.x == 0, mbar_wait_backoff<20>(bars + 16 * cs, cpar));
                    const uint32_t st = ring_u32 + cs * kStageBytes + tid * 16;
                    const int cv = min(kStageVec, nvec - c * kStageVec);
                    if (cv == kStageVec) {
                        float4 v[VPT];
#pragma unroll
                        for (int i = 0; i < VPT; ++i) v[i] = lds128(st + i * (NCT * 16));
#pragma unroll
                        for (int i = 0; i < VPT; ++i) classify4(v[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < VPT; ++i)
                            if (tid + i * NCT < cv) classify4(lds128(st + i * (NCT * 16)));
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars + 16 * cs + 8);
                    if (++cs == stages) { cs = 0; cpar ^= 1u; }
                }
                if (CLOSED) {
                    float n0, n1;
                    unpk(nan2, n0, n1);
                    flagged = (n0 != n0) || (n1 != n1);
                } else {
                    flagged = runmax != runmax;
                }
            };
            TMA_T(2, tid == 0 && blockIdx.x == 0, if (closed) run_sweep(std::true_type{}); else run_sweep(std::false_type{}));
            if (cur < 0) c0 = (addr - slot0) / kPrivStride;
            else pool_cnt[cur] = (addr - (pool_u32 + cur * (kPoolBlockWords * 4))) >> 2;
            buf_cnt0(b)[tid] = c0;
            {
                float b_lo, b_hi;
                unpk(below2, b_lo, b_hi);
                const uint32_t below = __reduce_add_sync(0xffffffffu, static_cast<uint32_t>(b_lo + b_hi));
                const uint32_t maxk = __reduce_max_sync(0xffffffffu, c0);
                const bool wflag = __any_sync(0xffffffffu, flagged);
                const bool wovf = __any_sync(0xffffffffu, ovf);
                if (lane == 0) {
                    if (below) atomicAdd(&mail[kMbBelow], below);
                    atomicMax(&mail[kMbMaxK], maxk);
                    if (wflag) mail[kMbNan] = 1u;
                    if (wovf) mail[kMbOvf] = 1u;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(ubar + 16 + 8 * b);   // lists, counts and flags of this warp are in place
        }
        return;
    }

    // ======================================== helper warps ===========================================
    // Two groups, so that neither sits on the other's critical path: the pivot group (2 warps) prepares the
    // bracket of the next units, the finish group (the other helper warps) resolves the swept ones.
    const int ht_all = tid - (NCT + 32);
    constexpr int kPivotThreads = 64, kFinishThreads = NHT - kPivotThreads;
    constexpr int kFinishWarps = kFinishThreads / 32;

    if (ht_all < kPivotThreads) {
        // ------------------------------------ pivot group -----------------------------------------------
        const int ht = ht_all, hwarp = ht >> 5;
        constexpr int KPT4 = (kSampleMax / 4 + kPivotThreads - 1) / kPivotThreads;   // sample float4 per thread
        auto psync = [] { asm volatile("bar.sync 3, %0;" ::"n"(kPivotThreads) : "memory"); };
        float4 samp[KPT4];
        auto load_sample = [&](int64_t u) {
            const int S4 = sample_size(n) >> 2;
            const uint32_t sstride = static_cast<uint32_t>(nvec / S4);
            const float4 *s4 = reinterpret_cast<const float4 *>(unit_std(u));
#pragma unroll
            for (int r = 0; r < KPT4; ++r) {
                const int i = ht + r * kPivotThreads;
                if (i < S4) {
                    const uint32_t jit = __umulhi(static_cast<uint32_t>(i) * 0x9E3779B1u, sstride);
                    samp[r] = ld_hint(s4 + static_cast<size_t>(i) * sstride + jit, pol_last);
                }
            }
        };
        // units that need no select are answered on the way
        auto next_answering = [&](int64_t i) {
            while (unit_at(i) < p.units && unit_mode(q_at(i)) != kModeThreshold) {
                if (ht == 0) {
                    const int64_t u = unit_at(i);
                    const float t = (unit_mode(q_at(i)) == kModeOnes) ? -INFINITY : INFINITY;
                    if (p.thr_out) p.thr_out[u] = t;
                    if (p.a_out) p.a_out[u] = t;
                    if (p.b_out) p.b_out[u] = t;
                }
                ++i;
            }
            return i;
        };
        int64_t i = next_answering(0);
        if (unit_at(i) < p.units) load_sample(unit_at(i));
        for (uint32_t seq = 0; unit_at(i) < p.units; ++seq) {
            const int b = static_cast<int>(seq & 1u);
            uint32_t lo, hi;
            float w;
            quantile_ranks(q_at(i), p.n, lo, hi, w);
            const int S = sample_size(n);
            // one histogram round over the sample (registers) + in-bucket interpolation
            for (int j = ht; j < kHistBins / 4; j += kPivotThreads) reinterpret_cast<uint4 *>(hist_p)[j] = make_uint4(0u, 0u, 0u, 0u);
            psync();
#pragma unroll
            for (int r = 0; r < KPT4; ++r) {
                if (ht + r * kPivotThreads < (S >> 2)) {
                    atomicAdd(&hist_p[float_to_key(samp[r].x) >> 21], 1u);
                    atomicAdd(&hist_p[float_to_key(samp[r].y) >> 21], 1u);
                    atomicAdd(&hist_p[float_to_key(samp[r].z) >> 21], 1u);
                    atomicAdd(&hist_p[float_to_key(samp[r].w) >> 21], 1u);
                }
            }
            // the next unit's sample lands while this one's bracket is computed and the sweep catches up
            const int64_t i_next = next_answering(i + 1);
            if (unit_at(i_next) < p.units) load_sample(unit_at(i_next));
            psync();
            const float frac = static_cast<float>(lo) / static_cast<float>(n > 1 ? n - 1 : 1);
            const float kt = frac * static_cast<float>(S - 1);
            const float margin = 3.5f * sqrtf(static_cast<float>(S) * frac * (1.0f - frac)) + 4.0f;
            const int klo = static_cast<int>(floorf(kt - margin));
            const int khi = static_cast<int>(ceilf(kt + margin));
            {
                const uint32_t r = hwarp == 0 ? static_cast<uint32_t>(klo > 0 ? klo : 0)
                                              : static_cast<uint32_t>(khi < S - 1 ? khi : S - 1);
                uint32_t tot;
                const BinHit h = warp_find2048(hist_p, r, tot);
                const float t = static_cast<float>(r - h.below);
                const float slack = 0.5f * sqrtf(static_cast<float>(h.count)) + 3.0f;
                const uint32_t k = (hwarp == 0) ? bucket_interp(h.bin, t - slack, h.count)
                                                : bucket_interp(h.bin, t + 1.0f + slack, h.count);
                if (lane == 0) scratch_p[hwarp] = k;
            }
            psync();
            if (ht == 0) {
                const float plo_f = (klo > 0) ? key_to_float(scratch_p[0]) : -INFINITY;
                const float phi_f = (khi < S - 1) ? key_to_float(scratch_p[1]) : INFINITY;
                const bool closed = (fabsf(plo_f) < INFINITY) && (fabsf(phi_f) < INFINITY);
                // closed bracket: classification on d = x - mid; |d| <= hw holds for every x in [plo, phi] (hw is
                // rounded up), and d is monotone in x, so {d < -hw}, {|d| <= hw}, {d > hw} partition the unit into a
                // down-set, an interval and an up-set
                const float mid = 0.5f * plo_f + 0.5f * phi_f;
                const float hw = closed ? fmaxf(__fsub_ru(phi_f, mid), __fsub_ru(mid, plo_f)) : 0.0f;
                uint32_t *mail = mail0 + b * kMbWords;
                // the mailbox and the list buffer still belong to unit seq - 2 until the finish group lets go
                mbar_wait_backoff<64>(ubar + 32 + 8 * b, ((seq >> 1) & 1u) ^ 1u);
                mail[kMbPlo] = __float_as_uint(plo_f); mail[kMbPhi] = __float_as_uint(phi_f);
                mail[kMbMid] = __float_as_uint(mid); mail[kMbHw] = __float_as_uint(hw);
                mail[kMbClosed] = closed ? 1u : 0u;
                mail[kMbBelow] = 0u; mail[kMbNan] = 0u; mail[kMbOvf] = 0u; mail[kMbPoolNext] = 0u; mail[kMbMaxK] = 0u;
                mbar_arrive(ubar + 8 * b);
            }
            psync();          // scratch_p is rewritten by the next unit
            i = i_next;
        }
        return;
    }

    // -------------------------------------- finish group ----------------------------------------------
    const int ht = ht_all - kPivotThreads, hwarp = ht >> 5;
    constexpr int kCols = (NCT + kFinishThreads - 1) / kFinishThreads;   // private lists per finish thread
    auto hsync = [] { asm volatile("bar.sync 2, %0;" ::"n"(kFinishThreads) : "memory"); };
    // coarse[c] = sum of the fine bins of coarse bin c (no second atomic per key)
    auto coarse_from_fine = [&](int w) {
        const int nb = 1 << w;
        const int nc = nb < kCoarseBins ? nb : kCoarseBins;
        const int fpc = nb / nc;
        for (int c = hwarp; c < nc; c += kFinishWarps) {
            const uint32_t v = (lane < fpc) ? hist[c * fpc + lane] : 0u;
            const uint32_t sum = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0) coarse[c] = sum;
        }
    };
    auto zero_hist = [&](int nb) {
        for (int j = ht; j < (nb + 3) / 4; j += kFinishThreads) reinterpret_cast<uint4 *>(hist)[j] = make_uint4(0u, 0u, 0u, 0u);
    };
    // hist filled -> (bin, below, count) of 0-based rank `rnk`, left in scratch[kScrHitA..]
    auto find_bin = [&](int w, uint32_t rnk) {
        hsync();
        coarse_from_fine(w);
        hsync();
        if (hwarp == 0) {
            const BinHit h = warp_find_bin(hist, coarse, w, rnk);
            if (lane == 0) { scratch[kScrHitA] = h.bin; scratch[kScrHitA + 1] = h.below; scratch[kScrHitA + 2] = h.count; }
        }
        hsync();
    };
    // Radix rounds (<= 11 bits each, most significant first) for the key of 0-based rank `rnk` among the keys
    // visited by `each(f)`, all in [kbase, kbase + 2^bits).  Returns the key; count / below_total describe its tie run.
    auto radix_rounds = [&](auto each, uint32_t kbase, int bits, uint32_t rnk, uint32_t total, uint32_t &below_total,
                            uint32_t &count) -> uint32_t {
        uint32_t prefix = 0;
        below_total = 0;
        count = total;
        int top = bits;
        while (top > 0) {
            const int wd = top < 11 ? top : 11;
            const int shift = top - wd;
            const int nb = 1 << wd;
            hsync();
            zero_hist(nb);
            hsync();
            const uint32_t want = (top >= 32) ? 0u : (prefix >> top);
            each([&](uint32_t k) {
                const uint32_t kn = k - kbase;
                if (top >= 32 || (kn >> top) == want) atomicAdd(&hist[(kn >> shift) & (nb - 1)], 1u);
            });
            find_bin(wd, rnk);
            prefix |= scratch[kScrHitA] << shift;
            rnk -= scratch[kScrHitA + 1];
            below_total += scratch[kScrHitA + 1];
            count = scratch[kScrHitA + 2];
            top = shift;
        }
        return kbase + prefix;
    };
    auto block_min = [&](uint32_t mine, int slot) -> uint32_t {
        hsync();
        if (ht == 0) scratch[slot] = 0xffffffffu;
        hsync();
        mine = __reduce_min_sync(0xffffffffu, mine);
        if (lane == 0 && mine != 0xffffffffu) atomicMin(&scratch[slot], mine);
        hsync();
        return scratch[slot];
    };

    // final phase of unit u (q01 = q) from list buffer b
    auto finish = [&](int64_t u, float q, int b) {
        uint32_t lo, hi;
        float w;
        quantile_ranks(q, p.n, lo, hi, w);
        const float *std_u = unit_std(u);
        const uint32_t *mail = mail0 + b * kMbWords;
        const float plo_f = __uint_as_float(mail[kMbPlo]), phi_f = __uint_as_float(mail[kMbPhi]);
        const float mid = __uint_as_float(mail[kMbMid]), hw = __uint_as_float(mail[kMbHw]);
        const bool closed = mail[kMbClosed] != 0u;
        const uint32_t c_below = mail[kMbBelow];
        const uint32_t npool = min(mail[kMbPoolNext], static_cast<uint32_t>(cfg.pool_blocks));
        const uint32_t maxk = mail[kMbMaxK];
        const uint32_t priv_u32 = smem_u32(buf_priv(b)), pool_u32 = smem_u32(buf_pool(b));
        const uint32_t *pool_cnt = buf_pool_cnt(b);
        const uint32_t *cnt0 = buf_cnt0(b);
        TMA_STAMP0();
        uint32_t colc[kCols];
#pragma unroll
        for (int j = 0; j < kCols; ++j) colc[j] = (ht + j * kFinishThreads < NCT) ? cnt0[ht + j * kFinishThreads] : 0u;
        // this finish thread's share of the candidates: rows of kCols private lists + pool blocks ht, ht + 192, ...
        // Several rows are loaded and processed together (independent chains: the phase is latency-bound, not
        // issue-bound); f(x, valid) must be cheap to run on padding (valid == false).
        constexpr int kRows = 3;
        auto for_my_cands = [&](auto f) {   // f(raw float, valid)
            for (uint32_t k = 0; k < maxk; k += kRows) {
                float x[kRows][kCols];
                bool v[kRows][kCols];
#pragma unroll
                for (int r = 0; r < kRows; ++r)
#pragma unroll
                    for (int j = 0; j < kCols; ++j) {
                        v[r][j] = (k + r) < colc[j];
                        x[r][j] = v[r][j] ? lds_f32(priv_u32 + ((k + r) * NCT + ht + j * kFinishThreads) * 4) : 0.0f;
                    }
#pragma unroll
                for (int r = 0; r < kRows; ++r)
#pragma unroll
                    for (int j = 0; j < kCols; ++j) f(x[r][j], v[r][j]);
            }
            for (uint32_t pb = ht; pb < npool; pb += kFinishThreads) {
                const uint32_t cnt = pool_cnt[pb];
                const uint32_t pa = pool_u32 + pb * (kPoolBlockWords * 4);
                for (uint32_t j = 0; j < cnt; ++j) f(lds_f32(pa + 4 * j), true);
            }
        };
        zero_hist(kHistBins);         // the previous unit ended with a barrier: hist / scratch are free
        if (ht == 0) {
            scratch[kScrMinAbove] = 0xffffffffu; scratch[kScrBinMin] = 0xffffffffu; scratch[kScrBinMax] = 0u;
            scratch[kScrListLen] = 0u;
        }
        bool has_nan = mail[kMbNan] != 0u;
        if (has_nan && closed) {
            // the packed-FMA flag also fires on +-inf: look for a real NaN (the unit is L2-resident)
            bool nn = false;
            const float4 *s4 = reinterpret_cast<const float4 *>(std_u);
            for (int j = ht; j < nvec; j += kFinishThreads) {
                const float4 v = __ldg(s4 + j);
                nn |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
            }
            has_nan = block_min(nn ? 0u : 0xffffffffu, kScrA) == 0u;
        }
        hsync();
        // the candidate count comes out of the histogram search below; until then only the cheap half of the check
        bool valid = mail[kMbOvf] == 0u && c_below <= lo;
        TMA_STAMP(8);
        uint32_t a_key = 0, b_key = 0;
        bool done = false;            // warp 0 / lane 0 already wrote the unit's outputs
        if (has_nan) {
            // any NaN -> NaN threshold (torch.quantile)
        } else if (valid) {
            // digit: monotone map of the candidate's value to [0, 2048).  closed: linear in d = x - mid;
            // open-ended: top bits of the bracket-normalised key.
            const float dscale = (closed && hw > 1e-30f) ? 1024.0f / hw : 0.0f;
            const uint32_t obase = float_to_key(plo_f);
            const uint32_t owidth = float_to_key(phi_f) - obase;
            const int obits = 32 - __clz(owidth | 1u);
            const int oshift = obits > 11 ? obits - 11 : 0;
            auto run_final = [&](auto closed_tag) {
            constexpr bool CLOSED = decltype(closed_tag)::value;   // compile-time: keeps the unrolled passes branch-free
            auto tval = [&](float x) { return __fmaf_rn(__fsub_rn(x, mid), dscale, 1024.0f); };
            auto digit = [&](float x) -> uint32_t {
                if (CLOSED) {
                    const int di = __float2int_rd(tval(x));
                    return static_cast<uint32_t>(di < 0 ? 0 : (di > 2047 ? 2047 : di));
                }
                return (float_to_key(x) - obase) >> oshift;
            };
            TMA_STAMP(9);
            const uint32_t hist_u32 = smem_u32(hist);
            // padding entries bump a per-lane dummy word instead: an unconditional atomic keeps the unrolled chains
            // free of branches (ptxas turns a predicated one back into a branch around its address computation)
            const uint32_t dummy_u32 = smem_u32(scratch + kScrDummy + lane);
            for_my_cands([&](float x, bool ok) {
                const uint32_t a = hist_u32 + digit(x) * 4;
                atomicAdd(reinterpret_cast<uint32_t *>(__cvta_shared_to_generic(ok ? a : dummy_u32)), 1u);
            });
            TMA_STAMP(10);
            const uint32_t rank = lo - c_below;
            hsync();
            if (hwarp == 0) {
                uint32_t tot;
                const BinHit h = warp_find2048(hist, rank, tot);
                if (lane == 0) {
                    scratch[kScrHitA] = h.bin; scratch[kScrHitA + 1] = h.below; scratch[kScrHitA + 2] = h.count;
                    scratch[kScrTmp] = tot;
                }
            }
            hsync();
            TMA_STAMP(11);
            const uint32_t bin = scratch[kScrHitA], bin_below = scratch[kScrHitA + 1], bin_count = scratch[kScrHitA + 2];
            valid = hi < c_below + scratch[kScrTmp];          // the bracket reaches far enough up
            if (!valid) return;
            const bool small = bin_count <= static_cast<uint32_t>(kSmallList);
            // members of the bin -> small list (or their key range); smallest key of the higher bins.  Branch-free per
            // entry: a thread remembers its first member and how many it saw; the rare second member re-walks the list.
            {
                uint32_t mymin = 0xffffffffu, bmin = 0xffffffffu, bmax = 0u, nhit = 0, hitk = 0;
                if (CLOSED) {
                    // digit(x) == bin  <=>  lo_t <= t < hi_t (the clamped end bins are open-ended): compares in float,
                    // keys only at the end (the key is monotone in the value)
                    const float lo_t = (bin == 0u) ? -INFINITY : static_cast<float>(bin);
                    const float hi_t = (bin == 2047u) ? INFINITY : static_cast<float>(bin + 1u);
                    float xmin = INFINITY, bxmin = INFINITY, bxmax = -INFINITY, hitx = 0.0f;
                    for_my_cands([&](float x, bool ok) {
                        const float t = tval(x);
                        const bool above = ok & (t >= hi_t);
                        const bool in = ok & (t >= lo_t) & (t < hi_t);
                        xmin = above ? fminf(xmin, x) : xmin;
                        hitx = (in & (nhit == 0u)) ? x : hitx;
                        nhit += in ? 1u : 0u;
                        bxmin = in ? fminf(bxmin, x) : bxmin;
                        bxmax = in ? fmaxf(bxmax, x) : bxmax;
                    });
                    mymin = (xmin < INFINITY) ? float_to_key(xmin) : 0xffffffffu;   // closed-bracket candidates are finite
                    hitk = float_to_key(hitx);
                    if (nhit) { bmin = float_to_key(bxmin); bmax = float_to_key(bxmax); }
                } else {
                    for_my_cands([&](float x, bool ok) {
                        const uint32_t k = float_to_key(x);
                        const uint32_t d = (k - obase) >> oshift;
                        mymin = (ok & (d > bin)) ? min(mymin, k) : mymin;
                        const bool in = ok & (d == bin);
                        hitk = (in & (nhit == 0u)) ? k : hitk;
                        nhit += in ? 1u : 0u;
                        bmin = in ? min(bmin, k) : bmin;
                        bmax = in ? max(bmax, k) : bmax;
                    });
                }
                if (small) {
                    if (nhit == 1u) {
                        scratch[kScrList + atomicAdd(&scratch[kScrListLen], 1u)] = hitk;
                    } else if (nhit > 1u) {
                        for_my_cands([&](float x, bool ok) {
                            if (ok && digit(x) == bin) scratch[kScrList + atomicAdd(&scratch[kScrListLen], 1u)] = float_to_key(x);
                        });
                    }
                }
                mymin = __reduce_min_sync(0xffffffffu, mymin);
                if (lane == 0 && mymin != 0xffffffffu) atomicMin(&scratch[kScrMinAbove], mymin);
                if (!small) {
                    bmin = __reduce_min_sync(0xffffffffu, bmin);
                    bmax = __reduce_max_sync(0xffffffffu, bmax);
                    if (lane == 0) { atomicMin(&scratch[kScrBinMin], bmin); atomicMax(&scratch[kScrBinMax], bmax); }
                }
            }
            hsync();
            TMA_STAMP(12);
            const uint32_t r_a = rank - bin_below;              // position of a inside the bin
            const uint32_t r_b = r_a + (hi - lo);               // position of b (may lie beyond the bin)
            if (small) {
                if (hwarp == 0) {
                    const bool real = static_cast<uint32_t>(lane) < bin_count;
                    const uint32_t mk = real ? scratch[kScrList + lane] : 0xffffffffu;
                    uint32_t less = 0, leq = 0;
#pragma unroll 8
                    for (int j = 0; j < kSmallList; ++j) {
                        const uint32_t other = __shfl_sync(0xffffffffu, mk, j);
                        less += (other < mk) ? 1u : 0u;
                        leq += (other <= mk) ? 1u : 0u;
                    }
                    // padding lanes hold 0xffffffff: they never count as < or <= a real key
                    const unsigned ma = __ballot_sync(0xffffffffu, real && less <= r_a && r_a < leq);
                    const unsigned mb = __ballot_sync(0xffffffffu, real && less <= r_b && r_b < leq);
                    const uint32_t ka = __shfl_sync(0xffffffffu, mk, ma ? (__ffs(ma) - 1) : 0);
                    const uint32_t kb_in = __shfl_sync(0xffffffffu, mk, mb ? (__ffs(mb) - 1) : 0);
                    if (lane == 0) {   // the answer goes out from here: nobody else needs it
                        const float a_val = key_to_float(ka), b_val = key_to_float(mb ? kb_in : scratch[kScrMinAbove]);
                        if (p.thr_out) p.thr_out[u] = quantile_lerp(a_val, b_val, w);
                        if (p.a_out) p.a_out[u] = a_val;
                        if (p.b_out) p.b_out[u] = b_val;
                    }
                }
                done = true;
            } else {
                // crowded bin (ties): radix rounds over the bin's members on keys normalised to the bin's range
                const uint32_t kmin = scratch[kScrBinMin], kmax = scratch[kScrBinMax];
                const int bits = (kmax == kmin) ? 0 : 32 - __clz(kmax - kmin);
                uint32_t below_total, count;
                a_key = radix_rounds([&](auto f) { for_my_cands([&](float x, bool ok) { if (ok && digit(x) == bin) f(float_to_key(x)); }); },
                                     kmin, bits, r_a, bin_count, below_total, count);
                if (r_b < below_total + count) {
                    b_key = a_key;
                } else {
                    // successor of a: the smallest larger key of the bin, else the smallest key of the higher bins
                    uint32_t best = 0xffffffffu;
                    for_my_cands([&](float x, bool ok) {
                        const uint32_t k = float_to_key(x);
                        if (ok && k > a_key && digit(x) == bin) best = min(best, k);
                    });
                    const uint32_t nb_in = block_min(best, kScrBinMin);
                    b_key = nb_in != 0xffffffffu ? nb_in : scratch[kScrMinAbove];
                }
            }
            };
            if (closed) run_final(std::true_type{}); else run_final(std::false_type{});
            if (valid && ht == 0) atomicAdd(&g_tma_sampled_units, 1ull);
        }
        if (!has_nan && !valid) {
            // bracket missed / pool ran dry: radix rounds over the whole unit (L2-resident after the sweep)
            if (ht == 0) atomicAdd(&g_tma_fallback_units, 1ull);
            const float4 *s4 = reinterpret_cast<const float4 *>(std_u);
            auto each_key = [&](auto f) {
                for (int j = ht; j < nvec; j += kFinishThreads) {
                    const float4 v = __ldg(s4 + j);
                    f(float_to_key(v.x)); f(float_to_key(v.y)); f(float_to_key(v.z)); f(float_to_key(v.w));
                }
            };
            uint32_t below_total, count;
            a_key = radix_rounds(each_key, 0u, 32, lo, static_cast<uint32_t>(n), below_total, count);
            if (hi < below_total + count) {
                b_key = a_key;
            } else {
                uint32_t best = 0xffffffffu;
                each_key([&](uint32_t k) { if (k > a_key) best = min(best, k); });
                b_key = block_min(best, kScrBinMin);
            }
        }
        if (ht == 0 && !done) {
            float a_val = key_to_float(a_key), b_val = key_to_float(b_key);
            float thr = quantile_lerp(a_val, b_val, w);
            if (has_nan) thr = a_val = b_val = __int_as_float(0x7fc00000);
            if (p.thr_out) p.thr_out[u] = thr;
            if (p.a_out) p.a_out[u] = a_val;
            if (p.b_out) p.b_out[u] = b_val;
        }
        TMA_STAMP(13);
        hsync();   // every finish thread is done with this buffer's lists, the mailbox, hist and scratch
        if (ht == 0) mbar_arrive(ubar + 32 + 8 * b);
    };

    uint32_t seq = 0;
    for (int64_t i = next_stream(0); unit_at(i) < p.units; i = next_stream(i + 1), ++seq) {
        const int b = static_cast<int>(seq & 1u);
        TMA_T(5, ht == 0 && blockIdx.x == 0, mbar_wait_backoff<64>(ubar + 16 + 8 * b, (seq >> 1) & 1u));   // sweep of unit seq complete
        TMA_T(6, ht == 0 && blockIdx.x == 0, finish(unit_at(i), q_at(i), b));
    }
}

// =============================================================================================
// Lean select: the same phases in ONE role, many CTAs per SM
// =============================================================================================
// For units below the TMA kernel's range (and wherever it measures faster): 256 threads per unit-at-a-time CTA,
// several CTAs resident per SM, so that one unit's latency-bound phases (pivots, final) overlap other units' sweeps.
// The sweep keeps four 128-bit loads in flight per thread and classifies straight from registers into the thread's
// private candidate list (no parking, no scan, no atomic); pivots and final phase as in the kernel above, every
// thread working on its own list.
struct LeanConfig {
    int k0;            // private candidate entries per thread (incl. 4 guard entries)
    int pool_blocks;
    uint32_t one;
    int prefetch_lines;   // 128-byte lines of the CTA's next unit pulled into L2 during the final phase (0 = off)
};
template <int CT>
struct LeanSmem {
    __host__ __device__ static size_t priv(const LeanConfig &c) { return size_t(c.k0) * CT * 4; }
    __host__ __device__ static size_t pool(const LeanConfig &c) { return size_t(c.pool_blocks) * kPoolBlockWords * 4; }
    __host__ __device__ static size_t pool_cnt(const LeanConfig &c) { return (size_t(c.pool_blocks) * 4 + 15) / 16 * 16; }
    static constexpr size_t kHist = (kHistBins + 8) * 4;
    static constexpr size_t kScratch = 256 * 4;
    __host__ __device__ static size_t total(const LeanConfig &c) { return priv(c) + pool(c) + pool_cnt(c) + kHist + kScratch; }
};
constexpr int kLsBelow = 130, kLsNan = 131, kLsOvf = 132, kLsPoolNext = 133, kLsMaxK = 134, kLsPivLo = 135, kLsPivHi = 136,
              kLsSampMin = 137, kLsSampMax = 138;   // scratch words [128, 192) are free in this kernel (no second coarse histogram)

template <int CT, int MINB, int VPI = 4, bool DB = false>
__global__ void __launch_bounds__(CT, MINB) select_lean_kernel(const SliceParams p, const LeanConfig cfg) {
    extern __shared__ __align__(128) unsigned char dyn[];
    using L = LeanSmem<CT>;
    constexpr uint32_t kPrivStride = CT * 4;
    constexpr int kWarps = CT / 32;
    constexpr int KPT4 = (kSampleMax / 4 + CT - 1) / CT;
    uint32_t *priv = reinterpret_cast<uint32_t *>(dyn);
    uint32_t *pool = reinterpret_cast<uint32_t *>(dyn + L::priv(cfg));
    uint32_t *pool_cnt = reinterpret_cast<uint32_t *>(dyn + L::priv(cfg) + L::pool(cfg));
    uint32_t *hist = reinterpret_cast<uint32_t *>(dyn + L::priv(cfg) + L::pool(cfg) + L::pool_cnt(cfg));
    uint32_t *scratch = hist + (kHistBins + 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = static_cast<int>(p.n);
    const int nvec = n >> 2;
    const uint32_t one = cfg.one;
    const uint32_t pool_blocks = static_cast<uint32_t>(cfg.pool_blocks);
    const uint32_t slot0 = smem_u32(priv) + tid * 4;
    const uint32_t pool_u32 = smem_u32(pool);
    const uint32_t hist_u32 = smem_u32(hist);
    const uint64_t pol_last = policy_evict_last();
    auto zero_hist = [&](int nb) {
        for (int j = tid; j < (nb + 3) / 4; j += CT) reinterpret_cast<uint4 *>(hist)[j] = make_uint4(0u, 0u, 0u, 0u);
    };
    auto coarse_from_fine = [&](int w) {
        uint32_t *coarse = scratch + kScrCoarseA;
        const int nb = 1 << w;
        const int nc = nb < kCoarseBins ? nb : kCoarseBins;
        const int fpc = nb / nc;
        for (int c = warp; c < nc; c += kWarps) {
            const uint32_t v = (lane < fpc) ? hist[c * fpc + lane] : 0u;
            const uint32_t sum = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0) coarse[c] = sum;
        }
    };
    auto radix_rounds = [&](auto each, uint32_t kbase, int bits, uint32_t rnk, uint32_t total, uint32_t &below_total,
                            uint32_t &count) -> uint32_t {
        uint32_t prefix = 0;
        below_total = 0;
        count = total;
        int top = bits;
        while (top > 0) {
            const int wd = top < 11 ? top : 11;
            const int shift = top - wd;
            const int nb = 1 << wd;
            __syncthreads();
            zero_hist(nb);
            __syncthreads();
            const uint32_t want = (top >= 32) ? 0u : (prefix >> top);
            each([&](uint32_t k) {
                const uint32_t kn = k - kbase;
                if (top >= 32 || (kn >> top) == want) atomicAdd(&hist[(kn >> shift) & (nb - 1)], 1u);
            });
            __syncthreads();
            coarse_from_fine(wd);
            __syncthreads();
            if (warp == 0) {
                const BinHit h = warp_find_bin(hist, scratch + kScrCoarseA, wd, rnk);
                if (lane == 0) { scratch[kScrHitA] = h.bin; scratch[kScrHitA + 1] = h.below; scratch[kScrHitA + 2] = h.count; }
            }
            __syncthreads();
            prefix |= scratch[kScrHitA] << shift;
            rnk -= scratch[kScrHitA + 1];
            below_total += scratch[kScrHitA + 1];
            count = scratch[kScrHitA + 2];
            top = shift;
        }
        return kbase + prefix;
    };
    auto block_min = [&](uint32_t mine, int slot) -> uint32_t {
        __syncthreads();
        if (tid == 0) scratch[slot] = 0xffffffffu;
        __syncthreads();
        mine = __reduce_min_sync(0xffffffffu, mine);
        if (lane == 0 && mine != 0xffffffffu) atomicMin(&scratch[slot], mine);
        __syncthreads();
        return scratch[slot];
    };

    for (int64_t u = blockIdx.x; u < p.units; u += gridDim.x) {
        const float q = p.q01_per_unit ? p.q01_per_unit[u] : p.q01;
        const int mode = p.rank_in ? (p.rank_in[u] == 0xffffffffu ? kModeZeros : kModeThreshold) : unit_mode(q);
        if (mode != kModeThreshold) {
            if (tid == 0) {
                const float t = (mode == kModeOnes) ? -INFINITY : INFINITY;
                if (p.thr_out) p.thr_out[u] = t;
                if (p.a_out) p.a_out[u] = t;
                if (p.b_out) p.b_out[u] = t;
            }
            continue;
        }
        uint32_t lo, hi;
        float w;
        if (p.rank_in) {          // explicit order statistic (pivots of the tiled select's pooled sample)
            lo = hi = min(p.rank_in[u], static_cast<uint32_t>(n - 1));
            w = 0.0f;
        } else {
            quantile_ranks(q, p.n, lo, hi, w);
        }
        const float *std_u = p.std + ((p.repeat > 1) ? u / p.repeat : u) * p.n;
        const float4 *s4 = reinterpret_cast<const float4 *>(std_u);
        const int S = sample_size(n);
#ifdef PIC_PHASE_TIMING
        long long lt0 = clock64();
#define LEAN_STAMP(i) do { if (tid == 0 && blockIdx.x == 0) { const long long t__ = clock64(); g_tma_phase_clk[i] += t__ - lt0; lt0 = t__; } } while (0)
#else
#define LEAN_STAMP(i) do {} while (0)
#endif
        // ---- sample + pivots ---------------------------------------------------------------------------
        float4 samp[KPT4];
        {
            const int S4 = S >> 2;
            const uint32_t sstride = static_cast<uint32_t>(nvec / S4);
#pragma unroll
            for (int r = 0; r < KPT4; ++r) {
                const int i = tid + r * CT;
                if (i < S4) {
                    const uint32_t jit = __umulhi(static_cast<uint32_t>(i) * 0x9E3779B1u, sstride);
                    samp[r] = ld_hint(s4 + static_cast<size_t>(i) * sstride + jit, pol_last);
                }
            }
        }
        zero_hist(kHistBins);
        if (tid == 0) {
            scratch[kLsBelow] = 0u; scratch[kLsNan] = 0u; scratch[kLsOvf] = 0u; scratch[kLsPoolNext] = 0u; scratch[kLsMaxK] = 0u;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < KPT4; ++r) {
            if (tid + r * CT < (S >> 2)) {
                atomicAdd(&hist[float_to_key(samp[r].x) >> 21], 1u);
                atomicAdd(&hist[float_to_key(samp[r].y) >> 21], 1u);
                atomicAdd(&hist[float_to_key(samp[r].z) >> 21], 1u);
                atomicAdd(&hist[float_to_key(samp[r].w) >> 21], 1u);
            }
        }
        __syncthreads();
        const float frac = static_cast<float>(lo) / static_cast<float>(n > 1 ? n - 1 : 1);
        const float kt = frac * static_cast<float>(S - 1);
        const float margin = 3.5f * sqrtf(static_cast<float>(S) * frac * (1.0f - frac)) + 4.0f;
        const int klo = static_cast<int>(floorf(kt - margin));
        const int khi = static_cast<int>(ceilf(kt + margin));
        if (warp < 2) {
            const uint32_t r = warp == 0 ? static_cast<uint32_t>(klo > 0 ? klo : 0) : static_cast<uint32_t>(khi < S - 1 ? khi : S - 1);
            uint32_t tot;
            const BinHit h = warp_find2048(hist, r, tot);
            const float t = static_cast<float>(r - h.below);
            const float slack = 0.5f * sqrtf(static_cast<float>(h.count)) + 3.0f;
            const uint32_t k = (warp == 0) ? bucket_interp(h.bin, t - slack, h.count) : bucket_interp(h.bin, t + 1.0f + slack, h.count);
            if (lane == 0) scratch[warp == 0 ? kLsPivLo : kLsPivHi] = k;
        } else if (warp < 4) {
            // lowest / highest occupied sample bucket: bounds the digit range of an open-ended bracket (final phase)
            uint32_t tot;
            const BinHit h = warp_find2048(hist, warp == 2 ? 0u : static_cast<uint32_t>(S - 1), tot);
            if (lane == 0) scratch[warp == 2 ? kLsSampMin : kLsSampMax] = h.bin;
        }
        __syncthreads();
        const float plo_f = (klo > 0) ? key_to_float(scratch[kLsPivLo]) : -INFINITY;
        const float phi_f = (khi < S - 1) ? key_to_float(scratch[kLsPivHi]) : INFINITY;
        const bool closed = (fabsf(plo_f) < INFINITY) && (fabsf(phi_f) < INFINITY);
        const float mid = 0.5f * plo_f + 0.5f * phi_f;
        const float hw = closed ? fmaxf(__fsub_ru(phi_f, mid), __fsub_ru(mid, plo_f)) : 0.0f;
        LEAN_STAMP(0);
        // ---- sweep: four 128-bit loads in flight per thread, classification from registers ----------------
        uint32_t addr = slot0, lim = slot0 + static_cast<uint32_t>(cfg.k0 - 4) * kPrivStride, stride = kPrivStride, c0 = 0;
        int cur = -1;
        bool ovf = false, flagged = false;
        f2 below2 = pk(0.0f, 0.0f);
        auto switch_block = [&]() {
            if (cur < 0) c0 = (addr - slot0) / kPrivStride;
            else pool_cnt[cur] = (addr - (pool_u32 + cur * (kPoolBlockWords * 4))) >> 2;
            uint32_t nb = atomicAdd(&scratch[kLsPoolNext], 1u);
            if (nb >= pool_blocks) { ovf = true; nb = pool_blocks - 1; }
            cur = static_cast<int>(nb);
            addr = pool_u32 + nb * (kPoolBlockWords * 4);
            lim = addr + (kPoolBlockWords - 4) * 4;
            stride = 4;
        };
        auto run_sweep = [&](auto closed_tag) {
            constexpr bool CLOSED = decltype(closed_tag)::value;
            const f2 nmid2 = pk(-mid, -mid), zero2 = pk(0.0f, 0.0f);
            const float nhw = -hw;
            f2 nan2 = pk(0.0f, 0.0f);
            float runmax = -INFINITY;
            auto classify4 = [&](const float4 &v) {
                if (CLOSED) {
                    const f2 v01 = pk(v.x, v.y), v23 = pk(v.z, v.w);
                    float d0, d1, d2, d3;
                    unpk(add2(v01, nmid2), d0, d1);
                    unpk(add2(v23, nmid2), d2, d3);
                    nan2 = fma2(v01, zero2, nan2);
                    nan2 = fma2(v23, zero2, nan2);
                    below2 = add2(below2, pk(fset_lt(d0, nhw), fset_lt(d1, nhw)));
                    below2 = add2(below2, pk(fset_lt(d2, nhw), fset_lt(d3, nhw)));
                    push4_abs_le(addr, d0, d1, d2, d3, hw, v, stride, one);
                } else {
                    runmax = max3_nan(runmax, v.x, v.y);
                    runmax = max3_nan(runmax, v.z, v.w);
                    below2 = add2(below2, pk(fset_lt(v.x, plo_f), fset_lt(v.y, plo_f)));
                    below2 = add2(below2, pk(fset_lt(v.z, plo_f), fset_lt(v.w, plo_f)));
                    push4_in_range(addr, v, plo_f, phi_f, stride, one);
                }
                if (addr > lim) switch_block();
            };
            int j = tid;
            if (DB) {
                // software pipeline: the loads of batch k+1 are issued before batch k is classified, so every thread
                // has VPI 128-bit loads in flight at all times (the sweep is bound by load latency, not by issue)
                float4 cur[VPI], nxt[VPI];
                bool have = j + (VPI - 1) * CT < nvec;
                if (have) {
#pragma unroll
                    for (int i = 0; i < VPI; ++i) cur[i] = ld_hint(s4 + j + i * CT, pol_last);
                }
                while (have) {
                    const int jn = j + VPI * CT;
                    const bool have_next = jn + (VPI - 1) * CT < nvec;
                    if (have_next) {
#pragma unroll
                        for (int i = 0; i < VPI; ++i) nxt[i] = ld_hint(s4 + jn + i * CT, pol_last);
                    }
#pragma unroll
                    for (int i = 0; i < VPI; ++i) classify4(cur[i]);
#pragma unroll
                    for (int i = 0; i < VPI; ++i) cur[i] = nxt[i];
                    j = jn;
                    have = have_next;
                }
            } else {
                for (; j + (VPI - 1) * CT < nvec; j += VPI * CT) {
                    float4 v[VPI];
#pragma unroll
                    for (int i = 0; i < VPI; ++i) v[i] = ld_hint(s4 + j + i * CT, pol_last);
#pragma unroll
                    for (int i = 0; i < VPI; ++i) classify4(v[i]);
                }
            }
            for (; j < nvec; j += CT) classify4(ld_hint(s4 + j, pol_last));
            if (CLOSED) {
                float n0, n1;
                unpk(nan2, n0, n1);
                flagged = (n0 != n0) || (n1 != n1);
            } else {
                flagged = runmax != runmax;
            }
        };
        if (closed) run_sweep(std::true_type{}); else run_sweep(std::false_type{});
        if (cur < 0) c0 = (addr - slot0) / kPrivStride;
        else pool_cnt[cur] = (addr - (pool_u32 + cur * (kPoolBlockWords * 4))) >> 2;
        {
            float b_lo, b_hi;
            unpk(below2, b_lo, b_hi);
            const uint32_t below = __reduce_add_sync(0xffffffffu, static_cast<uint32_t>(b_lo + b_hi));
            const uint32_t mk = __reduce_max_sync(0xffffffffu, c0);
            const bool wflag = __any_sync(0xffffffffu, flagged);
            const bool wovf = __any_sync(0xffffffffu, ovf);
            if (lane == 0) {
                if (below) atomicAdd(&scratch[kLsBelow], below);
                atomicMax(&scratch[kLsMaxK], mk);
                if (wflag) scratch[kLsNan] = 1u;
                if (wovf) scratch[kLsOvf] = 1u;
            }
        }
        LEAN_STAMP(1);
        if (cfg.prefetch_lines > 0) {
            // The final phase below and the next unit's pivot phase are latency-bound and leave HBM idle: pull the head
            // of this CTA's NEXT unit into L2 meanwhile, so that its sweep starts on L2 hits
            int64_t un = u + gridDim.x;
            while (un < p.units && unit_mode(p.q01_per_unit ? p.q01_per_unit[un] : p.q01) != kModeThreshold) un += gridDim.x;
            if (un < p.units) {
                const char *nxt = reinterpret_cast<const char *>(p.std + ((p.repeat > 1) ? un / p.repeat : un) * p.n);
                const int lines = min(cfg.prefetch_lines, (n * 4 + 127) / 128);
                for (int l = tid; l < lines; l += CT)
                    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(nxt + static_cast<size_t>(l) * 128));
            }
        }
        zero_hist(kHistBins);         // the sample histogram was last read before the sweep
        if (tid == 0) {
            scratch[kScrMinAbove] = 0xffffffffu; scratch[kScrBinMin] = 0xffffffffu; scratch[kScrBinMax] = 0u;
            scratch[kScrListLen] = 0u;
        }
        __syncthreads();
        LEAN_STAMP(6);
        // ---- final --------------------------------------------------------------------------------------
        const uint32_t c_below = scratch[kLsBelow];
        const uint32_t npool = min(scratch[kLsPoolNext], pool_blocks);
        const uint32_t maxk = scratch[kLsMaxK];
        constexpr int kRows = 8;
        auto for_my_cands = [&](auto f) {   // f(raw float, valid): this thread's list, kRows entries at a time
            for (uint32_t k = 0; k < maxk; k += kRows) {
                float x[kRows];
                bool v[kRows];
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    v[r] = (k + r) < c0;
                    x[r] = v[r] ? lds_f32(slot0 + (k + r) * kPrivStride) : 0.0f;
                }
#pragma unroll
                for (int r = 0; r < kRows; ++r) f(x[r], v[r]);
            }
            for (uint32_t pb = tid; pb < npool; pb += CT) {
                const uint32_t cnt = pool_cnt[pb];
                const uint32_t pa = pool_u32 + pb * (kPoolBlockWords * 4);
                for (uint32_t j = 0; j < cnt; ++j) f(lds_f32(pa + 4 * j), true);
            }
        };
        bool has_nan = scratch[kLsNan] != 0u;
        if (has_nan && closed) {
            bool nn = false;
            for (int j = tid; j < nvec; j += CT) {
                const float4 v = __ldg(s4 + j);
                nn |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
            }
            has_nan = block_min(nn ? 0u : 0xffffffffu, kScrA) == 0u;
        }
        bool valid = scratch[kLsOvf] == 0u && c_below <= lo;
        uint32_t a_key = 0, b_key = 0;
        bool done = false;
        if (has_nan) {
            // any NaN -> NaN threshold (torch.quantile)
        } else if (valid) {
            const float dscale = (closed && hw > 1e-30f) ? 1024.0f / hw : 0.0f;
            // open-ended bracket: the digit spans [okey_lo, okey_hi], the bracket's finite end and the sample's extreme
            // bucket on the open side (candidates beyond it clamp into the end bin: monotone, and few) -- digits over the
            // whole float range up to +-inf would crowd the candidates into a handful of bins
            uint32_t okey_lo = float_to_key(plo_f), okey_hi = float_to_key(phi_f);
            if (!(fabsf(plo_f) < INFINITY)) okey_lo = max(okey_lo, scratch[kLsSampMin] << 21);
            if (!(fabsf(phi_f) < INFINITY)) okey_hi = min(okey_hi, ((scratch[kLsSampMax] + 1u) << 21) - 1u);
            if (okey_hi <= okey_lo) okey_hi = okey_lo + 1u;
            const int obits = 32 - __clz((okey_hi - okey_lo) | 1u);
            const int oshift = obits > 11 ? obits - 11 : 0;
            auto run_final = [&](auto closed_tag) {
                constexpr bool CLOSED = decltype(closed_tag)::value;
                auto tval = [&](float x) { return __fmaf_rn(__fsub_rn(x, mid), dscale, 1024.0f); };
                auto kdigit = [&](uint32_t k) -> uint32_t {
                    const uint32_t d = (k - okey_lo) >> oshift;
                    return k < okey_lo ? 0u : (d > 2047u ? 2047u : d);
                };
                auto digit = [&](float x) -> uint32_t {
                    if (CLOSED) {
                        const int di = __float2int_rd(tval(x));
                        return static_cast<uint32_t>(di < 0 ? 0 : (di > 2047 ? 2047 : di));
                    }
                    return kdigit(float_to_key(x));
                };
                const uint32_t dummy_u32 = smem_u32(scratch + kScrDummy + lane);
                for_my_cands([&](float x, bool ok) {
                    const uint32_t a = hist_u32 + digit(x) * 4;
                    atomicAdd(reinterpret_cast<uint32_t *>(__cvta_shared_to_generic(ok ? a : dummy_u32)), 1u);
                });
                const uint32_t rank = lo - c_below;
                __syncthreads();
                LEAN_STAMP(3);
                if (warp == 0) {
                    uint32_t tot;
                    const BinHit h = warp_find2048(hist, rank, tot);
                    if (lane == 0) {
                        scratch[kScrHitA] = h.bin; scratch[kScrHitA + 1] = h.below; scratch[kScrHitA + 2] = h.count;
                        scratch[kScrTmp] = tot;
                    }
                }
                __syncthreads();
                const uint32_t bin = scratch[kScrHitA], bin_below = scratch[kScrHitA + 1], bin_count = scratch[kScrHitA + 2];
                valid = hi < c_below + scratch[kScrTmp];
                if (!valid) return;
                const bool small = bin_count <= static_cast<uint32_t>(kSmallList);
                LEAN_STAMP(4);
                {
                    uint32_t mymin = 0xffffffffu, bmin = 0xffffffffu, bmax = 0u, nhit = 0, hitk = 0;
                    if (CLOSED) {
                        const float lo_t = (bin == 0u) ? -INFINITY : static_cast<float>(bin);
                        const float hi_t = (bin == 2047u) ? INFINITY : static_cast<float>(bin + 1u);
                        float xmin = INFINITY, bxmin = INFINITY, bxmax = -INFINITY, hitx = 0.0f;
                        for_my_cands([&](float x, bool ok) {
                            const float t = tval(x);
                            const bool above = ok & (t >= hi_t);
                            const bool in = ok & (t >= lo_t) & (t < hi_t);
                            xmin = above ? fminf(xmin, x) : xmin;
                            hitx = (in & (nhit == 0u)) ? x : hitx;
                            nhit += in ? 1u : 0u;
                            bxmin = in ? fminf(bxmin, x) : bxmin;
                            bxmax = in ? fmaxf(bxmax, x) : bxmax;
                        });
                        mymin = (xmin < INFINITY) ? float_to_key(xmin) : 0xffffffffu;
                        hitk = float_to_key(hitx);
                        if (nhit) { bmin = float_to_key(bxmin); bmax = float_to_key(bxmax); }
                    } else {
                        for_my_cands([&](float x, bool ok) {
                            const uint32_t k = float_to_key(x);
                            const uint32_t d = kdigit(k);
                            mymin = (ok & (d > bin)) ? min(mymin, k) : mymin;
                            const bool in = ok & (d == bin);
                            hitk = (in & (nhit == 0u)) ? k : hitk;
                            nhit += in ? 1u : 0u;
                            bmin = in ? min(bmin, k) : bmin;
                            bmax = in ? max(bmax, k) : bmax;
                        });
                    }
                    if (small) {
                        if (nhit == 1u) {
                            scratch[kScrList + atomicAdd(&scratch[kScrListLen], 1u)] = hitk;
                        } else if (nhit > 1u) {
                            for_my_cands([&](float x, bool ok) {
                                if (ok && digit(x) == bin) scratch[kScrList + atomicAdd(&scratch[kScrListLen], 1u)] = float_to_key(x);
                            });
                        }
                    }
                    mymin = __reduce_min_sync(0xffffffffu, mymin);
                    if (lane == 0 && mymin != 0xffffffffu) atomicMin(&scratch[kScrMinAbove], mymin);
                    if (!small) {
                        bmin = __reduce_min_sync(0xffffffffu, bmin);
                        bmax = __reduce_max_sync(0xffffffffu, bmax);
                        if (lane == 0) { atomicMin(&scratch[kScrBinMin], bmin); atomicMax(&scratch[kScrBinMax], bmax); }
                    }
                }
                __syncthreads();
                LEAN_STAMP(5);
                const uint32_t r_a = rank - bin_below;
                const uint32_t r_b = r_a + (hi - lo);
                if (small) {
                    if (warp == 0) {
                        const bool real = static_cast<uint32_t>(lane) < bin_count;
                        const uint32_t mk = real ? scratch[kScrList + lane] : 0xffffffffu;
                        uint32_t less = 0, leq = 0;
#pragma unroll 8
                        for (int j = 0; j < kSmallList; ++j) {
                            const uint32_t other = __shfl_sync(0xffffffffu, mk, j);
                            less += (other < mk) ? 1u : 0u;
                            leq += (other <= mk) ? 1u : 0u;
                        }
                        const unsigned ma = __ballot_sync(0xffffffffu, real && less <= r_a && r_a < leq);
                        const unsigned mb = __ballot_sync(0xffffffffu, real && less <= r_b && r_b < leq);
                        const uint32_t ka = __shfl_sync(0xffffffffu, mk, ma ? (__ffs(ma) - 1) : 0);
                        const uint32_t kb_in = __shfl_sync(0xffffffffu, mk, mb ? (__ffs(mb) - 1) : 0);
                        if (lane == 0) {
                            const float a_val = key_to_float(ka), b_val = key_to_float(mb ? kb_in : scratch[kScrMinAbove]);
                            if (p.thr_out) p.thr_out[u] = quantile_lerp(a_val, b_val, w);
                            if (p.a_out) p.a_out[u] = a_val;
                            if (p.b_out) p.b_out[u] = b_val;
                        }
                    }
                    done = true;
                } else {
                    const uint32_t kmin = scratch[kScrBinMin], kmax = scratch[kScrBinMax];
                    const int bits = (kmax == kmin) ? 0 : 32 - __clz(kmax - kmin);
                    uint32_t below_total, count;
                    a_key = radix_rounds([&](auto f) { for_my_cands([&](float x, bool ok) { if (ok && digit(x) == bin) f(float_to_key(x)); }); },
                                         kmin, bits, r_a, bin_count, below_total, count);
                    if (r_b < below_total + count) {
                        b_key = a_key;
                    } else {
                        uint32_t best = 0xffffffffu;
                        for_my_cands([&](float x, bool ok) {
                            const uint32_t k = float_to_key(x);
                            if (ok && k > a_key && digit(x) == bin) best = min(best, k);
                        });
                        const uint32_t nb_in = block_min(best, kScrBinMin);
                        b_key = nb_in != 0xffffffffu ? nb_in : scratch[kScrMinAbove];
                    }
                }
            };
            if (closed) run_final(std::true_type{}); else run_final(std::false_type{});
            if (valid && tid == 0) atomicAdd(&g_tma_sampled_units, 1ull);
        }
        if (!has_nan && !valid) {
            if (tid == 0) atomicAdd(&g_tma_fallback_units, 1ull);
            auto each_key = [&](auto f) {
                for (int j = tid; j < nvec; j += CT) {
                    const float4 v = __ldg(s4 + j);
                    f(float_to_key(v.x)); f(float_to_key(v.y)); f(float_to_key(v.z)); f(float_to_key(v.w));
                }
            };
            uint32_t below_total, count;
            a_key = radix_rounds(each_key, 0u, 32, lo, static_cast<uint32_t>(n), below_total, count);
            if (hi < below_total + count) {
                b_key = a_key;
            } else {
                uint32_t best = 0xffffffffu;
                each_key([&](uint32_t k) { if (k > a_key) best = min(best, k); });
                b_key = block_min(best, kScrBinMin);
            }
        }
        if (tid == 0 && !done) {
            float a_val = key_to_float(a_key), b_val = key_to_float(b_key);
            float thr = quantile_lerp(a_val, b_val, w);
            if (has_nan) thr = a_val = b_val = __int_as_float(0x7fc00000);
            if (p.thr_out) p.thr_out[u] = thr;
            if (p.a_out) p.a_out[u] = a_val;
            if (p.b_out) p.b_out[u] = b_val;
        }
        __syncthreads();   // lists, hist and scratch are reused by the next unit
        LEAN_STAMP(2);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// private entries per sweep thread: expected bracket hits + ~2 sigma + 4 guard entries; the pool absorbs the rest
static int private_entries(int64_t n, int ct) {
    const double ept = static_cast<double>((n + ct - 1) / ct);
    int S = static_cast<int>(n >> 4);
    S = S < 1024 ? 1024 : (S > kSampleMax ? kSampleMax : S);
    const double brk = (2.0 * (3.5 * sqrt(S * 0.25) + 4.0) + 8.0) / S;
    const double mean = ept * brk;
    const int k = static_cast<int>(mean + 2.0 * sqrt(mean) + 1.0) + 4;
    return k < 8 ? 8 : k;
}

constexpr int kTmaSweepThreads = 640, kTmaHelperThreads = 256;
constexpr int kTmaVpt = 2;                // 20 KB stages: one producer thread issues a bulk copy every ~400 cycles
constexpr int64_t kTmaMinElems = 32768;   // smaller units: per-unit phases dominate, the multi-CTA-per-SM kernel wins

template <int NCT, int NHT, int VPT>
static bool tma_config(int64_t n, int smem_optin, TmaSelectConfig &cfg) {
    using L = TmaSmem<NCT, VPT>;
    static const int pool_env = env_int("PIC_TMA_POOL", 0);
    static const int stages_env = env_int("PIC_TMA_STAGES", 0);
    cfg.one = 1u;
    cfg.k0 = private_entries(n, NCT);
    cfg.pool_blocks = pool_env > 0 ? pool_env : 96;
    const int nchunks = static_cast<int>((n / 4 + NCT * VPT - 1) / (NCT * VPT));
    cfg.stages = stages_env > 0 ? stages_env : nchunks + 2;       // room for more than a whole unit
    while (cfg.stages > 4 && L::total(cfg) > static_cast<size_t>(smem_optin)) --cfg.stages;
    return L::total(cfg) <= static_cast<size_t>(smem_optin);
}

static int tma_smem_optin() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) v = 0;
        cached[dev] = v > 0 ? v : -1;
    }
    return cached[dev] > 0 ? cached[dev] : 0;
}

bool select_tma_usable(const SliceParams &p) {
    static const int enabled = env_int("PIC_TMA_SELECT", 1);
    if (!(enabled && p.apply_kind == 0 && !p.thr_in && p.n % 4 == 0 && p.n >= kTmaMinElems && p.n <= kFusedMaxElems &&
          aligned16(p.std) && p.units > 0))
        return false;
    TmaSelectConfig cfg;
    return tma_config<kTmaSweepThreads, kTmaHelperThreads, kTmaVpt>(p.n, tma_smem_optin(), cfg);
}

int launch_select_tma(const SliceParams &p, cudaStream_t stream) {
    constexpr int NCT = kTmaSweepThreads, NHT = kTmaHelperThreads, VPT = kTmaVpt;
    using L = TmaSmem<NCT, VPT>;
    auto kern = select_tma_kernel<NCT, NHT, VPT>;
    int dev = 0;
    PIC_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return PIC_ERR_INVALID_ARGUMENT;
    static bool configured[64] = {false};
    const int optin = tma_smem_optin();
    if (!configured[dev]) {
        PIC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        configured[dev] = true;
    }
    TmaSelectConfig cfg;
    if (!tma_config<NCT, NHT, VPT>(p.n, optin, cfg)) return PIC_ERR_TOO_LARGE;
    const int64_t max_grid = sm_count();
    const int grid = static_cast<int>(p.units < max_grid ? p.units : max_grid);
    kern<<<grid, NCT + 32 + NHT, L::total(cfg), stream>>>(p, cfg);
    return launch_status();
}

#ifdef PIC_PHASE_TIMING
extern "C" int pic_debug_tma_phase_clocks(long long *out, int reset) {
    cudaDeviceSynchronize();
    long long z[16] = {0};
    if (reset) return cudaMemcpyToSymbol(g_tma_phase_clk, z, sizeof(z)) == cudaSuccess ? 0 : -4;
    return cudaMemcpyFromSymbol(out, g_tma_phase_clk, sizeof(long long) * 16) == cudaSuccess ? 0 : -4;
}
#endif

bool select_lean_usable(const SliceParams &p) {
    static const int enabled = env_int("PIC_LEAN_SELECT", 1);
    return enabled && p.apply_kind == 0 && !p.thr_in && p.n % 4 == 0 && p.n > kCandMax && p.n <= kFusedMaxElems &&
           aligned16(p.std) && p.units > 0;
}

template <int CT, int MINB, int VPI = 4, bool DB = false>
static int launch_lean_t(const SliceParams &p, cudaStream_t stream) {
    using L = LeanSmem<CT>;
    auto kern = select_lean_kernel<CT, MINB, VPI, DB>;
    int dev = 0;
    PIC_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return PIC_ERR_INVALID_ARGUMENT;
    static bool configured[64] = {false};
    static int occ_cache[64] = {0};
    static size_t occ_smem[64] = {0};
    if (!configured[dev]) {
        PIC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tma_smem_optin()));
        configured[dev] = true;
    }
    static const int pool_env = env_int("PIC_LEAN_POOL", 0);
    LeanConfig cfg;
    cfg.one = 1u;
    cfg.k0 = private_entries(p.n, CT);
    cfg.pool_blocks = pool_env > 0 ? pool_env : 64;
    static const int pf_env = env_int("PIC_LEAN_PREFETCH", -1);
    cfg.prefetch_lines = pf_env >= 0 ? pf_env : 0;
    const size_t smem = L::total(cfg);
    if (smem > static_cast<size_t>(tma_smem_optin())) return PIC_ERR_TOO_LARGE;
    if (occ_cache[dev] == 0 || occ_smem[dev] != smem) {
        int occ = 1;
        PIC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CT, smem));
        occ_cache[dev] = occ < 1 ? 1 : occ;
        occ_smem[dev] = smem;
    }
    const int64_t max_grid = static_cast<int64_t>(sm_count()) * occ_cache[dev];
    const int grid = static_cast<int>(p.units < max_grid ? p.units : max_grid);
    kern<<<grid, CT, smem, stream>>>(p, cfg);
    return launch_status();
}

int launch_select_lean(const SliceParams &p, cudaStream_t stream) {
    // measured on B200 (scripts/select_tune.py; 1010 Kodak units): 256 threads x 4 CTAs/SM x 4 loads in flight 59.6 us,
    // 256 x 3 x 8 loads 56.8 us (the sweep is bound by load latency: bytes in flight beat occupancy), 5 or 6 CTAs/SM
    // (register-capped) 73-82 us, a register double-buffered load pipeline 65 us.  Smaller units prefer the 4-CTA form
    // (32768: 42.8 vs 45.7 us; 8192: 43.7 vs 49.1 us).  At most one unit per SM: 512-thread CTAs (101 units: 15.0 vs 17.0 us).
    static const int ct = env_int("PIC_LEAN_CT", 0);
    static const int vpi = env_int("PIC_LEAN_VPI", 0);
    const bool few = p.units <= sm_count();
    const int use = ct ? ct : (few ? 512 : 256);
    if (use == 512) return launch_lean_t<512, 2, 4>(p, stream);
    const int v = vpi ? vpi : (p.n >= 40000 ? 8 : 4);
    if (v == 8) return launch_lean_t<256, 3, 8>(p, stream);
    return launch_lean_t<256, 4, 4>(p, stream);
}

void select_tma_counters(unsigned long long *sampled, unsigned long long *fallback) {
    unsigned long long s = 0, f = 0;
    cudaMemcpyFromSymbol(&s, g_tma_sampled_units, sizeof(s));
    cudaMemcpyFromSymbol(&f, g_tma_fallback_units, sizeof(f));
    if (sampled) *sampled += s;
    if (fallback) *fallback += f;
}

}  // namespace pic
