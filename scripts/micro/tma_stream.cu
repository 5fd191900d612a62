// Microbenchmark: how fast can ONE CTA per SM stream global memory into shared memory with cp.async.bulk?
// usage: tma_stream <chunk_bytes> <stages> <MB per CTA> <ctas_per_sm> <grid_sms>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__global__ void __launch_bounds__(160) stream_kernel(const char *src, size_t bytes_per_cta, int chunk, int stages, unsigned long long *sink, int split) {
    extern __shared__ __align__(128) unsigned char dyn[];
    const uint32_t ring = smem_u32(dyn);
    const uint32_t bars = ring + (uint32_t)stages * chunk;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(bars + 16 * s, 1); mbar_init(bars + 16 * s + 8, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const char *base = src + (size_t)blockIdx.x * bytes_per_cta;
    const int nchunks = (int)(bytes_per_cta / chunk);
    if (warp == 4) {
        if (lane == 0) {
            int ps = 0; uint32_t par = 1;
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(bars + 16 * ps + 8, par);
                mbar_expect_tx(bars + 16 * ps, chunk);
                const int piece = chunk / split;
                for (int k = 0; k < split; ++k)
                    bulk_g2s(ring + ps * chunk + k * piece, base + (size_t)c * chunk + k * piece, piece, bars + 16 * ps);
                if (++ps == stages) { ps = 0; par ^= 1; }
            }
        }
        return;
    }
    int cs = 0; uint32_t par = 0; unsigned long long acc = 0;
    for (int c = 0; c < nchunks; ++c) {
        mbar_wait(bars + 16 * cs, par);
        acc += *reinterpret_cast<const uint32_t *>(dyn + cs * chunk + tid * 4);
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 16 * cs + 8);
        if (++cs == stages) { cs = 0; par ^= 1; }
    }
    if (acc == 0x123456789ull) *sink = acc;
}
int main(int argc, char **argv) {
    const int chunk = argc > 1 ? atoi(argv[1]) : 8192, stages = argc > 2 ? atoi(argv[2]) : 8;
    const size_t mb = argc > 3 ? atoi(argv[3]) : 8;
    const int per_sm = argc > 4 ? atoi(argv[4]) : 1, sms = argc > 5 ? atoi(argv[5]) : 148, split = argc > 6 ? atoi(argv[6]) : 1;
    const int grid = sms * per_sm;
    const size_t per_cta = mb << 20, total = per_cta * grid;
    char *src; unsigned long long *sink;
    cudaMalloc(&src, total); cudaMemset(src, 1, total); cudaMalloc(&sink, 8);
    const size_t smem = (size_t)stages * chunk + stages * 16;
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        stream_kernel<<<grid, 160, smem>>>(src, per_cta, chunk, stages, sink, split);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it == 2) printf("chunk %6d stages %2d split %d ctas/sm %d sms %3d: %8.3f ms  %7.1f GB/s  (%.1f B/clk/SM @1.965GHz)  err=%d\n", chunk, stages, split, per_sm, sms, ms,
                            total / ms / 1e6, total / ms / 1e6 / sms / 1.965, (int)cudaGetLastError());
    }
    return 0;
}
