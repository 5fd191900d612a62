// pic_rank.cu -- explicit variance-aware ranking (north_star step 1) as an exported permutation.
//
// The reference never materialises a ranking: ChannelMask keeps `std >= quantile` (layers/channel_mask.py:142-149),
// so every element tied with the threshold is kept.  For callers that want the order itself (progressive bit-stream
// ordering, analysis) this file exports it under the stated tie-break: key (std DESCENDING, linear NCHW index
// ASCENDING).  Implementation: order-preserving u32 keys of std (pic_math.cuh::float_to_key, -0 == +0, NaN above
// +inf as in torch.sort) and one stable segmented LSD radix sort per call (CUB, descending) over (key, index) pairs --
// stability is what turns "equal std" into "ascending index".
// Relation to the mask: with kept = #{std >= thr} (>= ceil((1 - q) n) because of ties), the mask's support is
// exactly order[0 .. kept).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_segmented_radix_sort.cuh>

#include "pic_math.cuh"
#include "pic_params.h"

namespace pic {

__global__ void __launch_bounds__(256) rank_keys_kernel(const float *std, int64_t n_per_unit, int64_t total,
                                                        uint32_t *keys, int32_t *index) {
    for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
        const float x = std[i];
        // NaN sorts as the largest value (torch.sort): above key(+inf) = 0xff800000
        keys[i] = (x != x) ? 0xffffffffu : float_to_key(x);
        index[i] = static_cast<int32_t>(i % n_per_unit);
    }
}

struct RankWs {
    uint32_t *keys_in, *keys_out;
    int32_t *index_in;
    int64_t *offsets;
    void *cub;
    size_t cub_bytes;
};

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static size_t rank_cub_bytes(int64_t total, int64_t units) {
    size_t bytes = 0;
    cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, bytes, static_cast<const uint32_t *>(nullptr),
                                                       static_cast<uint32_t *>(nullptr), static_cast<const int32_t *>(nullptr),
                                                       static_cast<int32_t *>(nullptr), total, static_cast<int>(units),
                                                       static_cast<const int64_t *>(nullptr),
                                                       static_cast<const int64_t *>(nullptr) + 1);
    return bytes;
}

__global__ void rank_offsets_kernel(int64_t *offsets, int64_t n_per_unit, int64_t units) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i <= units) offsets[i] = i * n_per_unit;
}

}  // namespace pic

using namespace pic;

extern "C" {

size_t pic_rank_order_workspace_bytes(int64_t n_per_unit, int64_t units) {
    if (n_per_unit <= 0 || units <= 0) return 256;
    const int64_t total = n_per_unit * units;
    return 3 * align256(static_cast<size_t>(total) * 4) + align256(static_cast<size_t>(units + 1) * 8) +
           align256(rank_cub_bytes(total, units)) + 256;
}

int pic_rank_order(const float *std, int64_t n_per_unit, int64_t units, int32_t *order_out, void *ws, size_t ws_bytes,
                   pic_stream_t stream_) {
    if (n_per_unit <= 0 || units <= 0 || !std || !order_out) return PIC_ERR_INVALID_ARGUMENT;
    if (n_per_unit > (int64_t(1) << 24)) return PIC_ERR_TOO_LARGE;
    if (units > 0x7fffffffLL) return PIC_ERR_TOO_LARGE;
    if (!ws || ws_bytes < pic_rank_order_workspace_bytes(n_per_unit, units)) return PIC_ERR_WORKSPACE;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t total = n_per_unit * units;
    unsigned char *b = static_cast<unsigned char *>(ws);
    b = reinterpret_cast<unsigned char *>(align256(reinterpret_cast<size_t>(b)));
    RankWs w;
    w.keys_in = reinterpret_cast<uint32_t *>(b);  b += align256(static_cast<size_t>(total) * 4);
    w.keys_out = reinterpret_cast<uint32_t *>(b); b += align256(static_cast<size_t>(total) * 4);
    w.index_in = reinterpret_cast<int32_t *>(b);  b += align256(static_cast<size_t>(total) * 4);
    w.offsets = reinterpret_cast<int64_t *>(b);   b += align256(static_cast<size_t>(units + 1) * 8);
    w.cub = b;
    w.cub_bytes = rank_cub_bytes(total, units);
    const int64_t blocks = (total + 1023) / 1024;
    const int grid = static_cast<int>(blocks < 148 * 16 ? (blocks < 1 ? 1 : blocks) : 148 * 16);
    rank_keys_kernel<<<grid, 256, 0, stream>>>(std, n_per_unit, total, w.keys_in, w.index_in);
    rank_offsets_kernel<<<static_cast<unsigned>((units + 256) / 256), 256, 0, stream>>>(w.offsets, n_per_unit, units);
    cudaError_t e = cub::DeviceSegmentedRadixSort::SortPairsDescending(
        w.cub, w.cub_bytes, static_cast<const uint32_t *>(w.keys_in), w.keys_out, static_cast<const int32_t *>(w.index_in),
        order_out, total, static_cast<int>(units), static_cast<const int64_t *>(w.offsets),
        static_cast<const int64_t *>(w.offsets) + 1, 0, 32, stream);
    if (e != cudaSuccess) {
        g_last_cuda_error = static_cast<int>(e);
        return PIC_ERR_CUDA;
    }
    return launch_status();
}

}  // extern "C"
