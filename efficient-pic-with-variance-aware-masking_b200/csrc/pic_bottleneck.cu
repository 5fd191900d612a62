// pic_bottleneck.cu -- EntropyBottleneck likelihood of the hyper-latent z.
//
// Reference: entropy_models/entropy_models.py:403-436 (_logits_cumulative, _likelihood), 449-492 (forward) and
// their autograd.  Per channel c a five-layer scalar network with filters (1, f1, f2, f3, f4, 1):
//     h <- softplus(M_i) h + b_i ;  h <- h + tanh(F_i) * tanh(h)      (no gate after the last layer)
// evaluated at x - 1/2 and x + 1/2:
//     s = -sign(lower + upper) ;  lik = | sigmoid(s upper) - sigmoid(s lower) | ;  lik = max(lik, bound)
// with x = round(z - median) + median (eval) or z + noise (training; the noise tensor is drawn by torch).
// The reference permutes z to [C, 1, B*S], runs ~40 small batched matmuls / elementwise kernels and permutes back;
// here one kernel reads z where the conv wrote it ([B, C, S]) and writes outputs + likelihood, and one kernel
// produces the gradients of z AND of every parameter (per-thread register accumulators, block reduction, one
// atomicAdd per parameter and CTA).  The parameters arrive packed per channel (layout below, built by the caller
// from the module's _matrix / _bias / _factor tensors): softplus and tanh of the raw parameters are applied here.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pic_latent.h"

namespace {

constexpr int kMaxF = 4;                 // filters per layer of the reference: (3, 3, 3, 3); up to 4 supported
constexpr int kLayers = 5;

struct Net {                              // activated parameters of one channel, in shared memory
    float W[kLayers][kMaxF][kMaxF];       // softplus(M_i)[out][in]
    float b[kLayers][kMaxF];
    float t[kLayers - 1][kMaxF];          // tanh(F_i)
};

struct Dims {
    int f[kLayers + 1];                   // 1, f1, f2, f3, f4, 1
    int off_m[kLayers], off_b[kLayers], off_f[kLayers - 1];
    int per_channel;
};

__host__ __device__ inline Dims make_dims(int f1, int f2, int f3, int f4) {
    Dims d;
    d.f[0] = 1; d.f[1] = f1; d.f[2] = f2; d.f[3] = f3; d.f[4] = f4; d.f[5] = 1;
    int o = 0;
    for (int i = 0; i < kLayers; ++i) { d.off_m[i] = o; o += d.f[i + 1] * d.f[i]; }
    for (int i = 0; i < kLayers; ++i) { d.off_b[i] = o; o += d.f[i + 1]; }
    for (int i = 0; i < kLayers - 1; ++i) { d.off_f[i] = o; o += d.f[i + 1]; }
    d.per_channel = o;
    return d;
}

__device__ __forceinline__ float softplus_f(float x) {      // torch: x > 20 ? x : log1p(exp(x))
    return x > 20.0f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ void load_net(const float *params, const Dims &d, int c, Net &net) {
    const float *p = params + static_cast<size_t>(c) * d.per_channel;
    for (int i = threadIdx.x; i < d.per_channel; i += blockDim.x) {
        // locate i: matrices, then biases, then factors
        if (i < d.off_b[0]) {
            int l = kLayers - 1;
            while (i < d.off_m[l]) --l;
            const int r = i - d.off_m[l];
            net.W[l][r / d.f[l]][r % d.f[l]] = softplus_f(p[i]);
        } else if (i < d.off_f[0]) {
            int l = kLayers - 1;
            while (i < d.off_b[l]) --l;
            net.b[l][i - d.off_b[l]] = p[i];
        } else {
            int l = kLayers - 2;
            while (i < d.off_f[l]) --l;
            net.t[l][i - d.off_f[l]] = tanhf(p[i]);
        }
    }
}

// forward of the scalar network; a[l][j] (pre-gate activations) kept for the backward when KEEP
template <bool KEEP>
__device__ __forceinline__ float logits(const Net &net, const Dims &d, float v, float (*a)[kMaxF], float (*h)[kMaxF]) {
    float cur[kMaxF] = {v, 0.f, 0.f, 0.f};
    for (int l = 0; l < kLayers; ++l) {
        float nxt[kMaxF];
        for (int j = 0; j < d.f[l + 1]; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < d.f[l]; ++k) acc = fmaf(net.W[l][j][k], cur[k], acc);
            acc += net.b[l][j];
            if (KEEP) { a[l][j] = acc; }
            nxt[j] = (l < kLayers - 1) ? acc + net.t[l][j] * tanhf(acc) : acc;
        }
        if (KEEP) for (int k = 0; k < d.f[l]; ++k) h[l][k] = cur[k];
        for (int j = 0; j < d.f[l + 1]; ++j) cur[j] = nxt[j];
    }
    return cur[0];
}

// one CTA per (channel, slab of the B*S positions of that channel)
__global__ void __launch_bounds__(256) eb_forward_kernel(const float *z, const float *noise, const float *medians,
                                                         const float *params, Dims d, int64_t B, int64_t C, int64_t S,
                                                         float lik_bound, float *outputs, float *lik) {
    __shared__ Net net;
    const int c = blockIdx.x;
    load_net(params, d, c, net);
    __syncthreads();
    const float med = medians[c];
    const int64_t total = B * S;
    for (int64_t e = blockIdx.y * 256LL + threadIdx.x; e < total; e += gridDim.y * 256LL) {
        const int64_t bidx = e / S, s = e - bidx * S;
        const int64_t g = (bidx * C + c) * S + s;
        const float x = noise ? z[g] + noise[g] : rintf(z[g] - med) + med;     // quantize("noise" | "dequantize", medians)
        const float lower = logits<false>(net, d, x - 0.5f, nullptr, nullptr);
        const float upper = logits<false>(net, d, x + 0.5f, nullptr, nullptr);
        const float sum = lower + upper;
        const float sgn = sum > 0.0f ? -1.0f : (sum < 0.0f ? 1.0f : 0.0f);     // -sign(lower + upper)
        float l = fabsf(sigmoid_f(sgn * upper) - sigmoid_f(sgn * lower));
        if (lik_bound > 0.0f) l = !(l < lik_bound) ? l : lik_bound;
        if (outputs) outputs[g] = x;
        lik[g] = l;
    }
}

// gradients: g_lik (and g_out, nullable) -> g_z (nullable), g_params [C][per_channel] (atomically accumulated, raw-
// parameter space: softplus' and tanh' applied), g_medians [C] (eval mode only: d outputs / d median = 1)
__global__ void __launch_bounds__(128) eb_backward_kernel(const float *z, const float *noise, const float *medians,
                                                          const float *params, Dims d, int64_t B, int64_t C, int64_t S,
                                                          float lik_bound, const float *g_lik, const float *g_out,
                                                          float *g_z, float *g_params, float *g_medians) {
    __shared__ Net net;
    __shared__ float red[64];
    const int c = blockIdx.x;
    load_net(params, d, c, net);
    __syncthreads();
    const float med = medians[c];
    float gW[kLayers][kMaxF][kMaxF] = {}, gb[kLayers][kMaxF] = {}, gt[kLayers - 1][kMaxF] = {};
    float gmed = 0.0f;
    const int64_t total = B * S;
    for (int64_t e = blockIdx.y * 128LL + threadIdx.x; e < total; e += gridDim.y * 128LL) {
        const int64_t bidx = e / S, s = e - bidx * S;
        const int64_t g = (bidx * C + c) * S + s;
        const float x = noise ? z[g] + noise[g] : rintf(z[g] - med) + med;
        float a0[kLayers][kMaxF], h0[kLayers][kMaxF], a1[kLayers][kMaxF], h1[kLayers][kMaxF];
        const float lower = logits<true>(net, d, x - 0.5f, a0, h0);
        const float upper = logits<true>(net, d, x + 0.5f, a1, h1);
        const float sum = lower + upper;
        const float sgn = sum > 0.0f ? -1.0f : (sum < 0.0f ? 1.0f : 0.0f);
        const float su = sigmoid_f(sgn * upper), sl = sigmoid_f(sgn * lower);
        const float diff = su - sl;
        const float raw = fabsf(diff);
        float gl = g_lik ? g_lik[g] : 0.0f;
        if (lik_bound > 0.0f && !(raw >= lik_bound || gl < 0.0f)) gl = 0.0f;      // LowerBound backward rule
        const float sd = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);
        const float g_up = gl * sd * sgn * su * (1.0f - su);
        const float g_lo = -gl * sd * sgn * sl * (1.0f - sl);
        float gx = 0.0f;
        for (int pass = 0; pass < 2; ++pass) {
            float (*a)[kMaxF] = pass ? a1 : a0;
            float (*h)[kMaxF] = pass ? h1 : h0;
            float gh[kMaxF] = {pass ? g_up : g_lo, 0.f, 0.f, 0.f};                // gradient of the layer's OUTPUT
            for (int l = kLayers - 1; l >= 0; --l) {
                float ga[kMaxF];
                for (int j = 0; j < d.f[l + 1]; ++j) {
                    if (l < kLayers - 1) {
                        const float th = tanhf(a[l][j]);
                        ga[j] = gh[j] * (1.0f + net.t[l][j] * (1.0f - th * th));
                        gt[l][j] += gh[j] * th;
                    } else {
                        ga[j] = gh[j];
                    }
                    gb[l][j] += ga[j];
                }
                float gin[kMaxF] = {0.f, 0.f, 0.f, 0.f};
                for (int j = 0; j < d.f[l + 1]; ++j)
                    for (int k = 0; k < d.f[l]; ++k) {
                        gW[l][j][k] += ga[j] * h[l][k];
                        gin[k] = fmaf(net.W[l][j][k], ga[j], gin[k]);
                    }
                for (int k = 0; k < d.f[l]; ++k) gh[k] = gin[k];
            }
            gx += gh[0];
        }
        const float go = g_out ? g_out[g] : 0.0f;
        if (noise) {
            if (g_z) g_z[g] = gx + go;              // outputs = z + noise: identity
        } else {
            if (g_z) g_z[g] = 0.0f;                 // round(): zero gradient to z
            gmed += gx + go;                        // outputs = round(z - med) + med: the median passes through
        }
    }
    // block reduction of every accumulator, one atomicAdd per parameter and CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto reduce_add = [&](float v, float *dst) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.0f;
            for (int w = 0; w < 4; ++w) t += red[w];
            if (t != 0.0f) atomicAdd(dst, t);
        }
        __syncthreads();
    };
    const float *p = params + static_cast<size_t>(c) * d.per_channel;
    float *gp = g_params + static_cast<size_t>(c) * d.per_channel;
    for (int l = 0; l < kLayers; ++l)
        for (int j = 0; j < d.f[l + 1]; ++j)
            for (int k = 0; k < d.f[l]; ++k) {
                const float m = p[d.off_m[l] + j * d.f[l] + k];
                reduce_add(gW[l][j][k] * sigmoid_f(m), gp + d.off_m[l] + j * d.f[l] + k);    // softplus' = sigmoid
            }
    for (int l = 0; l < kLayers; ++l)
        for (int j = 0; j < d.f[l + 1]; ++j) reduce_add(gb[l][j], gp + d.off_b[l] + j);
    for (int l = 0; l < kLayers - 1; ++l)
        for (int j = 0; j < d.f[l + 1]; ++j) {
            const float th = net.t[l][j];
            reduce_add(gt[l][j] * (1.0f - th * th), gp + d.off_f[l] + j);                    // tanh'
        }
    if (g_medians) reduce_add(gmed, g_medians + c);
}

bool dims_ok(int f1, int f2, int f3, int f4) {
    const int f[4] = {f1, f2, f3, f4};
    for (int v : f)
        if (v < 1 || v > kMaxF) return false;
    return true;
}

}  // namespace

extern "C" {

int pic_bottleneck_params_per_channel(int f1, int f2, int f3, int f4) {
    if (!dims_ok(f1, f2, f3, f4)) return PIC_ERR_INVALID_ARGUMENT;
    return make_dims(f1, f2, f3, f4).per_channel;
}

int pic_bottleneck_forward(const float *z, const float *noise, const float *medians, const float *params, int f1, int f2,
                           int f3, int f4, int64_t batch, int64_t channels, int64_t spatial, float lik_bound,
                           float *outputs, float *lik, pic_stream_t stream) {
    if (!z || !medians || !params || !lik || batch <= 0 || channels <= 0 || spatial <= 0 || !dims_ok(f1, f2, f3, f4))
        return PIC_ERR_INVALID_ARGUMENT;
    if (channels > 65535 * 32768LL) return PIC_ERR_TOO_LARGE;
    const Dims d = make_dims(f1, f2, f3, f4);
    const int64_t slabs = (batch * spatial + 256 * 8 - 1) / (256 * 8);
    dim3 grid(static_cast<unsigned>(channels), static_cast<unsigned>(slabs < 1 ? 1 : (slabs > 64 ? 64 : slabs)));
    eb_forward_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(z, noise, medians, params, d, batch, channels, spatial,
                                                                          lik_bound, outputs, lik);
    return cudaGetLastError() == cudaSuccess ? PIC_OK : PIC_ERR_CUDA;
}

int pic_bottleneck_backward(const float *z, const float *noise, const float *medians, const float *params, int f1, int f2,
                            int f3, int f4, int64_t batch, int64_t channels, int64_t spatial, float lik_bound,
                            const float *g_lik, const float *g_out, float *g_z, float *g_params, float *g_medians,
                            pic_stream_t stream_) {
    if (!z || !medians || !params || !g_params || batch <= 0 || channels <= 0 || spatial <= 0 || !dims_ok(f1, f2, f3, f4))
        return PIC_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const Dims d = make_dims(f1, f2, f3, f4);
    if (cudaMemsetAsync(g_params, 0, sizeof(float) * channels * d.per_channel, stream) != cudaSuccess) return PIC_ERR_CUDA;
    if (g_medians && cudaMemsetAsync(g_medians, 0, sizeof(float) * channels, stream) != cudaSuccess) return PIC_ERR_CUDA;
    const int64_t slabs = (batch * spatial + 128 * 8 - 1) / (128 * 8);
    dim3 grid(static_cast<unsigned>(channels), static_cast<unsigned>(slabs < 1 ? 1 : (slabs > 32 ? 32 : slabs)));
    eb_backward_kernel<<<grid, 128, 0, stream>>>(z, noise, medians, params, d, batch, channels, spatial, lik_bound, g_lik, g_out,
                                                 g_z, g_params, g_medians);
    return cudaGetLastError() == cudaSuccess ? PIC_OK : PIC_ERR_CUDA;
}

}  // extern "C"
