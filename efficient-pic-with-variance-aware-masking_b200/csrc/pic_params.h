// pic_params.h -- declarations shared by the translation units of libpic_latent.so (host helpers, launch
// constants, the parameter block of the slice kernels).  Internal: the public interface is include/pic_latent.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pic_latent.h"
#include "pic_math.cuh"
#include "pic_select.cuh"

namespace pic {

// ------------------------------------------------------------------------------------------
// host-side helpers
// ------------------------------------------------------------------------------------------
extern thread_local int g_last_cuda_error;   // defined in pic_latent.cu

#define PIC_CUDA_CHECK(expr)                                   \
    do {                                                       \
        cudaError_t e__ = (expr);                              \
        if (e__ != cudaSuccess) {                              \
            pic::g_last_cuda_error = static_cast<int>(e__);    \
            return PIC_ERR_CUDA;                               \
        }                                                      \
    } while (0)

static inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = static_cast<int>(e);
        return PIC_ERR_CUDA;
    }
    return PIC_OK;
}

static inline int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned4(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

constexpr int kFusedMaxElems = 131072;  // one CTA per unit up to here; larger units take the multi-CTA rounds path
constexpr int kScratchWords = 192;  // block_select uses [0,48); small-set helpers [44,192)
constexpr int kTableSmem = 64;
constexpr int kHistWords = kHistBins + 8;  // [2048] = NaN count; rest padding
constexpr int kApplyTile = 8192;           // elements per CTA in slice_apply_kernel
constexpr int kWideThreads = 1024;         // latency variant of the select-only kernel (units <= SM count)
constexpr int kRoundChunk = 16384;         // elements per CTA in hist_round_kernel
constexpr int64_t kTwoKernelMinElems = int64_t(1) << 22;  // >= 4 Mi elements: select kernel + apply kernel
constexpr int64_t kTwoKernelMinUnit = 32768;               // ... and units of at least this many elements

struct SliceParams {
    const float *y_top, *y_base, *mu, *std, *q01_per_unit, *thr_in, *noise, *table;
    float q01, scale_bound, lik_bound;
    int table_len;
    int64_t n, units;
    float *mask, *y_hat, *lik;
    int32_t *idx, *symbols;
    float *thr_out, *a_out, *b_out;
    double *rate;
    int apply_kind;  // 0: select only, 1: mask only, 2: full slice
    int use_stage;   // dynamic shared memory holds the cp.async stage buffer
    int repeat;      // select-only: `repeat` consecutive (virtual) units share one std block (multi-quality select)
    const uint32_t *rank_in;   // lean select only: explicit 0-based order statistic per (virtual) unit instead of the
                               // quantile of q01 (0xffffffff: skip the unit); a_out receives the element of that rank
};


// TMA-staged select kernel (pic_tma_select.cu): thresholds of `units` units of n elements, select only.
// Requires n % 4 == 0, a 16-byte aligned std, kCandMax < n <= kFusedMaxElems and no thr_in.
bool select_tma_usable(const SliceParams &p);
int launch_select_tma(const SliceParams &p, cudaStream_t stream);
// Lean select kernel (same file): kCandMax < n <= kFusedMaxElems, n % 4 == 0, 16-byte aligned std, no thr_in.
bool select_lean_usable(const SliceParams &p);
int launch_select_lean(const SliceParams &p, cudaStream_t stream);
void select_tma_counters(unsigned long long *sampled, unsigned long long *fallback);

}  // namespace pic
