// pic_host.cu -- host-buffer entry point of the C ABI (include/pic_latent.h, section 6).
//
// A caller whose latents live in host memory (the reference's CPU path, an FFI caller without
// device tensors) streams units through the device in chunks: three slots, each with its own
// stream, so chunk c's H2D copies, chunk c-1's kernel and chunk c-2's D2H copies overlap on
// the two copy engines and the SMs.  With pinned host memory the whole pipeline is
// asynchronous; the call returns after the last D2H copy has landed.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pic_latent.h"

namespace {

constexpr int kSlots = 3;
constexpr size_t kAlign = 256;

struct Pipeline {
    cudaStream_t stream[kSlots] = {nullptr, nullptr, nullptr};
    bool ready = false;
};

thread_local Pipeline g_pipe;

inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct SlotLayout {
    size_t array_bytes;  // chunk_units * n * 4
    size_t small_bytes;  // q01 / thr (f32) + rate (f64) per unit
    size_t slot_bytes;
    size_t table_bytes;
};

SlotLayout layout(int64_t n, int64_t chunk_units) {
    SlotLayout l;
    l.array_bytes = align_up(static_cast<size_t>(n) * static_cast<size_t>(chunk_units) * 4);
    l.small_bytes = align_up(static_cast<size_t>(chunk_units) * 16);
    l.slot_bytes = 10 * l.array_bytes + l.small_bytes + 256;  // +256: kernel workspace stub
    l.table_bytes = align_up(4096);
    return l;
}

// f32 {0,1} mask and int32 scale index -> one byte each (four elements per thread, 32-bit stores)
__global__ void __launch_bounds__(256) pack_u8_kernel(const float *mask, const int32_t *idx, int64_t n, uint8_t *mask8,
                                                      uint8_t *idx8) {
    const int64_t nvec = n >> 2;
    for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < nvec; i += gridDim.x * 256LL) {
        const float4 m = reinterpret_cast<const float4 *>(mask)[i];
        const int4 k = reinterpret_cast<const int4 *>(idx)[i];
        const uint32_t mb = (m.x != 0.0f ? 1u : 0u) | (m.y != 0.0f ? 1u << 8 : 0u) | (m.z != 0.0f ? 1u << 16 : 0u) |
                            (m.w != 0.0f ? 1u << 24 : 0u);
        const uint32_t kb = (static_cast<uint32_t>(k.x) & 255u) | ((static_cast<uint32_t>(k.y) & 255u) << 8) |
                            ((static_cast<uint32_t>(k.z) & 255u) << 16) | ((static_cast<uint32_t>(k.w) & 255u) << 24);
        reinterpret_cast<uint32_t *>(mask8)[i] = mb;
        reinterpret_cast<uint32_t *>(idx8)[i] = kb;
    }
    for (int64_t i = (nvec << 2) + blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        mask8[i] = mask[i] != 0.0f ? 1 : 0;
        idx8[i] = static_cast<uint8_t>(idx[i]);
    }
}

}  // namespace

extern "C" {

size_t pic_host_pipeline_bytes(int64_t n_per_unit, int64_t chunk_units) {
    if (n_per_unit <= 0 || chunk_units <= 0) return 0;
    const SlotLayout l = layout(n_per_unit, chunk_units);
    // large units additionally need the multi-launch select workspace per slot
    const size_t ws = align_up(pic_workspace_bytes(n_per_unit, chunk_units));
    return l.table_bytes + kSlots * (l.slot_bytes + ws);
}

int pic_slice_forward_host(const float *y_top, const float *y_base, const float *mu, const float *std,
                           float q01, const float *q01_per_unit_host, const float *noise,
                           const float *scale_table_host, int table_len, float scale_bound,
                           float lik_bound, int64_t n_per_unit, int64_t units, int64_t chunk_units,
                           float *mask, float *y_hat, float *lik, int32_t *idx, int32_t *symbols,
                           float *thr_out, double *rate, void *device_buf, size_t device_buf_bytes) {
    if (n_per_unit <= 0 || units <= 0 || chunk_units <= 0 || !y_top || !mu || !std || !device_buf)
        return PIC_ERR_INVALID_ARGUMENT;
    if (idx && (!scale_table_host || table_len < 1 || table_len > 1024)) return PIC_ERR_INVALID_ARGUMENT;
    if (chunk_units > units) chunk_units = units;
    if (device_buf_bytes < pic_host_pipeline_bytes(n_per_unit, chunk_units)) return PIC_ERR_WORKSPACE;
#define PIC_HOST_CHECK(expr)                        \
    do {                                            \
        cudaError_t e__ = (expr);                   \
        if (e__ != cudaSuccess) return PIC_ERR_CUDA; \
    } while (0)
    if (!g_pipe.ready) {
        for (int s = 0; s < kSlots; ++s) PIC_HOST_CHECK(cudaStreamCreateWithFlags(&g_pipe.stream[s], cudaStreamNonBlocking));
        g_pipe.ready = true;
    }
    const SlotLayout l = layout(n_per_unit, chunk_units);
    const size_t ws_bytes = align_up(pic_workspace_bytes(n_per_unit, chunk_units));
    unsigned char *base = static_cast<unsigned char *>(device_buf);
    float *d_table = reinterpret_cast<float *>(base);
    base += l.table_bytes;
    if (idx) {
        // the table is tiny; a synchronous copy keeps every slot stream independent
        PIC_HOST_CHECK(cudaMemcpy(d_table, scale_table_host, sizeof(float) * table_len, cudaMemcpyHostToDevice));
    }
    const size_t unit_bytes = static_cast<size_t>(n_per_unit) * 4;
    int rc = PIC_OK;
    int64_t chunk = 0;
    for (int64_t u0 = 0; u0 < units && rc == PIC_OK; u0 += chunk_units, ++chunk) {
        const int slot = static_cast<int>(chunk % kSlots);
        cudaStream_t st = g_pipe.stream[slot];
        const int64_t cu = (units - u0 < chunk_units) ? (units - u0) : chunk_units;
        unsigned char *sb = base + static_cast<size_t>(slot) * (l.slot_bytes + ws_bytes);
        float *d_arr[10];
        for (int a = 0; a < 10; ++a) d_arr[a] = reinterpret_cast<float *>(sb + a * l.array_bytes);
        unsigned char *small = sb + 10 * l.array_bytes;
        float *d_q = reinterpret_cast<float *>(small);
        float *d_thr = d_q + chunk_units;
        double *d_rate = reinterpret_cast<double *>(small + static_cast<size_t>(chunk_units) * 8);
        void *d_ws = sb + l.slot_bytes;
        const size_t bytes = unit_bytes * static_cast<size_t>(cu);
        const size_t off = static_cast<size_t>(u0) * static_cast<size_t>(n_per_unit);
        // inputs
        PIC_HOST_CHECK(cudaMemcpyAsync(d_arr[0], y_top + off, bytes, cudaMemcpyHostToDevice, st));
        if (y_base) PIC_HOST_CHECK(cudaMemcpyAsync(d_arr[1], y_base + off, bytes, cudaMemcpyHostToDevice, st));
        PIC_HOST_CHECK(cudaMemcpyAsync(d_arr[2], mu + off, bytes, cudaMemcpyHostToDevice, st));
        PIC_HOST_CHECK(cudaMemcpyAsync(d_arr[3], std + off, bytes, cudaMemcpyHostToDevice, st));
        if (noise) PIC_HOST_CHECK(cudaMemcpyAsync(d_arr[4], noise + off, bytes, cudaMemcpyHostToDevice, st));
        if (q01_per_unit_host)
            PIC_HOST_CHECK(cudaMemcpyAsync(d_q, q01_per_unit_host + u0, sizeof(float) * cu, cudaMemcpyHostToDevice, st));
        rc = pic_slice_forward(d_arr[0], y_base ? d_arr[1] : nullptr, d_arr[2], d_arr[3], q01,
                               q01_per_unit_host ? d_q : nullptr, nullptr, noise ? d_arr[4] : nullptr,
                               idx ? d_table : nullptr, table_len, scale_bound, lik_bound, n_per_unit, cu,
                               mask ? d_arr[5] : nullptr, y_hat ? d_arr[6] : nullptr, lik ? d_arr[7] : nullptr,
                               idx ? reinterpret_cast<int32_t *>(d_arr[8]) : nullptr,
                               symbols ? reinterpret_cast<int32_t *>(d_arr[9]) : nullptr,
                               thr_out ? d_thr : nullptr, rate ? d_rate : nullptr, d_ws, ws_bytes, st);
        if (rc != PIC_OK) break;
        // outputs
        if (mask) PIC_HOST_CHECK(cudaMemcpyAsync(mask + off, d_arr[5], bytes, cudaMemcpyDeviceToHost, st));
        if (y_hat) PIC_HOST_CHECK(cudaMemcpyAsync(y_hat + off, d_arr[6], bytes, cudaMemcpyDeviceToHost, st));
        if (lik) PIC_HOST_CHECK(cudaMemcpyAsync(lik + off, d_arr[7], bytes, cudaMemcpyDeviceToHost, st));
        if (idx) PIC_HOST_CHECK(cudaMemcpyAsync(idx + off, d_arr[8], bytes, cudaMemcpyDeviceToHost, st));
        if (symbols) PIC_HOST_CHECK(cudaMemcpyAsync(symbols + off, d_arr[9], bytes, cudaMemcpyDeviceToHost, st));
        if (thr_out) PIC_HOST_CHECK(cudaMemcpyAsync(thr_out + u0, d_thr, sizeof(float) * cu, cudaMemcpyDeviceToHost, st));
        if (rate) PIC_HOST_CHECK(cudaMemcpyAsync(rate + u0, d_rate, sizeof(double) * cu, cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < kSlots; ++s) {
        cudaError_t e = cudaStreamSynchronize(g_pipe.stream[s]);
        if (e != cudaSuccess && rc == PIC_OK) rc = PIC_ERR_CUDA;
    }
#undef PIC_HOST_CHECK
    return rc;
}

int pic_slice_forward_host_compact(const float *y_top, const float *y_base, const float *mu, const float *std, float q01,
                                   const float *q01_per_unit_host, const float *scale_table_host, int table_len,
                                   float scale_bound, float lik_bound, int64_t n_per_unit, int64_t units,
                                   int64_t chunk_units, uint8_t *mask_u8, float *y_hat, float *lik, uint8_t *idx_u8,
                                   void *device_buf, size_t device_buf_bytes) {
    if (n_per_unit <= 0 || units <= 0 || chunk_units <= 0 || !y_top || !mu || !std || !device_buf || !mask_u8 || !idx_u8)
        return PIC_ERR_INVALID_ARGUMENT;
    if (!scale_table_host || table_len < 1 || table_len > 256) return PIC_ERR_INVALID_ARGUMENT;   // indexes must fit a byte
    if (chunk_units > units) chunk_units = units;
    if (device_buf_bytes < pic_host_pipeline_bytes(n_per_unit, chunk_units)) return PIC_ERR_WORKSPACE;
#define PIC_HOST_CHECK(expr)                        \
    do {                                            \
        cudaError_t e__ = (expr);                   \
        if (e__ != cudaSuccess) return PIC_ERR_CUDA; \
    } while (0)
    if (!g_pipe.ready) {
        for (int s = 0; s < kSlots; ++s) PIC_HOST_CHECK(cudaStreamCreateWithFlags(&g_pipe.stream[s], cudaStreamNonBlocking));
        g_pipe.ready = true;
    }
    const SlotLayout l = layout(n_per_unit, chunk_units);
    const size_t ws_bytes = align_up(pic_workspace_bytes(n_per_unit, chunk_units));
    unsigned char *base = static_cast<unsigned char *>(device_buf);
    float *d_table = reinterpret_cast<float *>(base);
    base += l.table_bytes;
    PIC_HOST_CHECK(cudaMemcpy(d_table, scale_table_host, sizeof(float) * table_len, cudaMemcpyHostToDevice));
    const size_t unit_bytes = static_cast<size_t>(n_per_unit) * 4;
    int rc = PIC_OK;
    int64_t chunk = 0;
    for (int64_t u0 = 0; u0 < units && rc == PIC_OK; u0 += chunk_units, ++chunk) {
        const int slot = static_cast<int>(chunk % kSlots);
        cudaStream_t st = g_pipe.stream[slot];
        const int64_t cu = (units - u0 < chunk_units) ? (units - u0) : chunk_units;
        unsigned char *sb = base + static_cast<size_t>(slot) * (l.slot_bytes + ws_bytes);
        float *d_arr[10];
        for (int a = 0; a < 10; ++a) d_arr[a] = reinterpret_cast<float *>(sb + a * l.array_bytes);
        float *d_q = reinterpret_cast<float *>(sb + 10 * l.array_bytes);
        void *d_ws = sb + l.slot_bytes;
        const size_t bytes = unit_bytes * static_cast<size_t>(cu);
        const int64_t elems = n_per_unit * cu;
        const size_t off = static_cast<size_t>(u0) * static_cast<size_t>(n_per_unit);
        PIC_HOST_CHECK(cudaMemcpyAsync(d_arr[0], y_top + off, bytes, cudaMemcpyHostToDevice, st));
        if (y_base) PIC_HOST_CHECK(cudaMemcpyAsync(d_arr[1], y_base + off, bytes, cudaMemcpyHostToDevice, st));
        PIC_HOST_CHECK(cudaMemcpyAsync(d_arr[2], mu + off, bytes, cudaMemcpyHostToDevice, st));
        PIC_HOST_CHECK(cudaMemcpyAsync(d_arr[3], std + off, bytes, cudaMemcpyHostToDevice, st));
        if (q01_per_unit_host)
            PIC_HOST_CHECK(cudaMemcpyAsync(d_q, q01_per_unit_host + u0, sizeof(float) * cu, cudaMemcpyHostToDevice, st));
        rc = pic_slice_forward(d_arr[0], y_base ? d_arr[1] : nullptr, d_arr[2], d_arr[3], q01,
                               q01_per_unit_host ? d_q : nullptr, nullptr, nullptr, d_table, table_len, scale_bound,
                               lik_bound, n_per_unit, cu, d_arr[5], y_hat ? d_arr[6] : nullptr, lik ? d_arr[7] : nullptr,
                               reinterpret_cast<int32_t *>(d_arr[8]), nullptr, nullptr, nullptr, d_ws, ws_bytes, st);
        if (rc != PIC_OK) break;
        // one byte per mask / index element: arrays 4 (noise) and 9 (symbols) of the slot are free in this mode
        uint8_t *d_m8 = reinterpret_cast<uint8_t *>(d_arr[4]), *d_i8 = reinterpret_cast<uint8_t *>(d_arr[9]);
        const int64_t blocks = (elems / 4 + 255) / 256;
        pack_u8_kernel<<<static_cast<unsigned>(blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks)), 256, 0, st>>>(
            d_arr[5], reinterpret_cast<const int32_t *>(d_arr[8]), elems, d_m8, d_i8);
        PIC_HOST_CHECK(cudaGetLastError());
        PIC_HOST_CHECK(cudaMemcpyAsync(mask_u8 + off, d_m8, static_cast<size_t>(elems), cudaMemcpyDeviceToHost, st));
        if (y_hat) PIC_HOST_CHECK(cudaMemcpyAsync(y_hat + off, d_arr[6], bytes, cudaMemcpyDeviceToHost, st));
        if (lik) PIC_HOST_CHECK(cudaMemcpyAsync(lik + off, d_arr[7], bytes, cudaMemcpyDeviceToHost, st));
        PIC_HOST_CHECK(cudaMemcpyAsync(idx_u8 + off, d_i8, static_cast<size_t>(elems), cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < kSlots; ++s) {
        cudaError_t e = cudaStreamSynchronize(g_pipe.stream[s]);
        if (e != cudaSuccess && rc == PIC_OK) rc = PIC_ERR_CUDA;
    }
#undef PIC_HOST_CHECK
    return rc;
}

}  // extern "C"
