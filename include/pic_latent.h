/*
 * pic_latent.h -- C ABI of libpic_latent.so: the B200 (sm_100a) implementation of the
 * latent-side hot path of the PIC / REM progressive image codec
 * (reference: das-ankur/Efficient-PIC-with-Variance-Aware-Masking, paths below are
 * relative to its src/ directory).
 *
 * The reference has no FFI: its boundary for this path is the Python method surface
 *   layers/channel_mask.py:9-156      ChannelMask.{forward, ProgMask, apply_noise}, ste_round
 *   entropy_models/entropy_models.py  EntropyModel.{quantize 127-153, dequantize 161-168},
 *                                     GaussianConditional.{_likelihood 620-635, forward 637-652,
 *                                     build_indexes 654-659}
 *   models/pic.py:401-402,430-443,583-584,621-629,809-820,945-948  (per-slice glue)
 *   training/loss.py:45-60            (rate reduction)
 * Each entry point below names the reference lines it replaces.  The Python drop-in classes
 * (efficient-pic-with-variance-aware-masking_b200/) bind these symbols with ctypes; a
 * maintainer of the reference would bind them the same way.
 *
 * Conventions
 *  - All tensor pointers are caller-owned DEVICE memory on the current CUDA device, dense
 *    f32/i32, laid out as `units` consecutive blocks of `n_per_unit` elements.  A *unit* is
 *    one (image, progressive slice) block [32,h,w] of a contiguous NCHW tensor [B,32,h,w],
 *    i.e. n_per_unit = 32*h*w and units = B (times slices when slices are batched).
 *  - Every call is asynchronous on `stream` (a cudaStream_t), never synchronises the host,
 *    never allocates, keeps no state between calls and is CUDA-graph capturable.  Scratch
 *    memory is passed in (`ws`, size from the matching *_workspace_bytes()).
 *  - Inputs are never modified.  Nullable pointers are marked; a NULL output is skipped.
 *  - Quantile control: `q01` is the f32 torch.quantile argument 1 - min(pr,10)*0.1
 *    (channel_mask.py:138-140).  Values outside [0,1] select the reference's
 *    short-circuits: q01 < 0  => all-ones mask (pr >= 10, channel_mask.py:133-134),
 *                    q01 > 1  => all-zeros mask (pr == 0, channel_mask.py:135-136).
 *    `q01_per_unit` (device, nullable) overrides the scalar `q01` with one value per unit.
 *  - Return value: PIC_OK or a negative PIC_ERR_* code; nothing throws.
 */
#ifndef PIC_LATENT_H_
#define PIC_LATENT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIC_OK 0
#define PIC_ERR_INVALID_ARGUMENT (-1) /* ValueError / NotImplementedError in the reference API   */
#define PIC_ERR_TOO_LARGE (-2)        /* torch.quantile: "input tensor is too large" (n > 2^24)   */
#define PIC_ERR_WORKSPACE (-3)        /* ws_bytes smaller than *_workspace_bytes()               */
#define PIC_ERR_CUDA (-4)             /* a CUDA runtime call failed: see pic_last_cuda_error()    */
#define PIC_ERR_UNALIGNED (-5)        /* pointer not 4-byte aligned                              */

#define PIC_Q_ONES (-1.0f) /* q01 sentinel: pr >= 10 */
#define PIC_Q_ZEROS (2.0f) /* q01 sentinel: pr == 0  */

#define PIC_QUANTIZE_NOISE 0      /* entropy_models.py:132-138 */
#define PIC_QUANTIZE_DEQUANTIZE 1 /* entropy_models.py:140-149 */
#define PIC_QUANTIZE_SYMBOLS 2    /* entropy_models.py:151-153 */
#define PIC_QUANTIZE_STE 3        /* ste_round forward: round(x) - x + x (channel_mask.py:5-6) */

typedef void *pic_stream_t; /* cudaStream_t */

int pic_version(void);
const char *pic_error_string(int code);
int pic_last_cuda_error(void); /* cudaError_t of the last failing runtime call on this thread */

/* Largest n_per_unit served by the single-launch fused kernel (std tile of the unit staged in
 * shared memory between select and apply); larger units take the multi-launch path. */
int64_t pic_fused_max_elems(void);

/* Launch plan of pic_slice_forward for a problem size: returns 0 (single fused kernel), 1 (select
 * kernel + tile-ordered apply kernel) or 2 (large units: pivot + sweep + cluster select + apply kernel, given
 * the workspace of pic_workspace_bytes) and stores the number of kernel launches in *n_kernels.  needs_select = 0 when thresholds are supplied. */
int pic_slice_forward_plan(int64_t n_per_unit, int64_t units, int needs_select, int *n_kernels);

/* Diagnostics (synchronises the device): number of units served by the sampled-pivot select and
 * number that fell back to the full histogram select, since the library was loaded. */
int pic_debug_select_counters(unsigned long long *sampled, unsigned long long *fallback);

/* Scratch needed by pic_select_threshold / pic_channel_mask / pic_slice_forward (caller-owned device memory,
 * reusable across calls on the same stream).  Units up to pic_fused_max_elems(): ~20 KB per unit (pivots,
 * candidates, thresholds).  Larger units: round state + histograms + a candidate buffer of n_per_unit / 8 floats
 * and one counter per 8192 elements per unit (sampled sweep + cluster select).  A smaller workspace is accepted
 * for large units down to the plain radix rounds' need (~25 KB per unit) and selects that slower path;
 * below that the call returns PIC_ERR_WORKSPACE. */
size_t pic_workspace_bytes(int64_t n_per_unit, int64_t units);

/*
 * (1) Threshold selection: exact torch.quantile(std[u].ravel(), q01, interpolation="linear")
 * per unit (channel_mask.py:41,145; ATen quantile semantics: f32 rank = q01*f32(n-1),
 * a = sorted[floor], b = sorted[ceil], FMA lerp; any NaN => NaN).  For the q01 sentinels
 * thr_out is -inf (ones) / +inf (zeros).  a_out / b_out (nullable) receive the two order
 * statistics.
 */
int pic_select_threshold(const float *std, int64_t n_per_unit, int64_t units, float q01,
                         const float *q01_per_unit, float *thr_out, float *a_out, float *b_out,
                         void *ws, size_t ws_bytes, pic_stream_t stream);

/*
 * (1a) Multi-quality form of (1) for progressive level packing (test/functions_encode.py:176-190,
 * functions_decode.py:186-200, which call ProgMask twice per level on the same scale list):
 * q01_levels is a device array [units][levels] (level-minor); thr_out receives [units][levels].
 * All levels of a unit are selected in one launch from the same std block (n_per_unit <=
 * pic_fused_max_elems()).  pic_level_map then yields, per element, the first level l with
 * std >= thr[u][l] (or `levels`): the delta mask ProgMask(q_l) - ProgMask(q_{l-1}) of nested
 * quality levels is exactly (level == l).
 */
int pic_select_threshold_multi(const float *std, int64_t n_per_unit, int64_t units,
                               const float *q01_levels, int levels, float *thr_out,
                               pic_stream_t stream);
int pic_level_map(const float *std, const float *thr, int64_t n_per_unit, int64_t units, int levels,
                  int32_t *level, pic_stream_t stream);

/*
 * (1b) Explicit variance-aware ranking (north_star step 1).  The reference has no ranking array: its mask is
 * `std >= quantile` (layers/channel_mask.py:142-149), which keeps EVERY element tied with the threshold.  This
 * entry exports the order itself under the stated tie-break: key (std descending, linear NCHW index within the unit
 * ascending); -0 ties with +0, NaN ranks first (torch.sort's order).  order_out: [units][n_per_unit] int32, the
 * unit-local element indexes from the most to the least uncertain.  Relation to (1)/(3): with
 * kept = #{std >= thr}, the mask's support is exactly order[0 .. kept) -- kept >= ceil((1 - q01) n), with equality
 * unless elements tie with the threshold.  ws: pic_rank_order_workspace_bytes().
 */
size_t pic_rank_order_workspace_bytes(int64_t n_per_unit, int64_t units);
int pic_rank_order(const float *std, int64_t n_per_unit, int64_t units, int32_t *order_out, void *ws, size_t ws_bytes,
                   pic_stream_t stream);

/*
 * (1b) Split form of (1) for spatially tiled units (one image sharded over several GPUs,
 * SURVEY 8e).  Radix rounds r = 0,1,2 (11/11/10 key bits).  Per round every rank calls
 * pic_hist_round() on its local tile, all-reduces `hist` (uint32 sum, pic_hist_words() words
 * per unit) and calls pic_select_advance(); after round 2 ranks all-reduce `min_above`
 * (uint32 min, one word per unit) and call pic_select_finish().  `state` is
 * pic_select_state_bytes(units) bytes of device memory owned by the caller; n_total is the
 * element count of the whole (global) unit.
 */
size_t pic_select_state_bytes(int64_t units);
int64_t pic_hist_words(void);
int pic_select_begin(void *state, int64_t n_total, int64_t units, float q01,
                     const float *q01_per_unit, pic_stream_t stream);
int pic_hist_round(const float *std_local, int64_t n_local, int64_t units, int round,
                   const void *state, uint32_t *hist, uint32_t *min_above, pic_stream_t stream);
int pic_select_advance(void *state, const uint32_t *hist, int64_t units, int round,
                       pic_stream_t stream);
int pic_select_finish(const void *state, const uint32_t *min_above, int64_t units, float *thr_out,
                      float *a_out, float *b_out, pic_stream_t stream);

/*
 * (1c) The same protocol with the collectives issued by this library: one call enqueues begin, 3 x (histogram
 * kernel, ncclAllReduce(uint32, sum), advance), ncclAllReduce(uint32, min) and finish on `stream` -- no host work
 * between the steps (the Python loop over pic_hist_round + torch.distributed costs ~10 us of host time per
 * step, more than the kernels at 8 GPUs).  NCCL is bound at run time to the libnccl.so.2 already loaded in the
 * process (torch's); the communicator is created from a 128-byte unique id that the caller broadcasts over its
 * own channel (torch.distributed, MPI, a file).  ws: pic_tiled_workspace_bytes(units) bytes (~25 KB per unit).
 */
#define PIC_DIST_ID_BYTES 128
size_t pic_tiled_workspace_bytes(int64_t units);
int pic_dist_unique_id(unsigned char *id_out);                                /* rank 0 */
int pic_dist_comm_init(const unsigned char *id, int rank, int world_size, void **comm_out);
int pic_dist_comm_destroy(void *comm);
int pic_tiled_select_threshold(const float *std_local, int64_t n_local, int64_t n_total, int64_t units,
                               float q01, const float *q01_per_unit, float *thr_out, void *ws,
                               size_t ws_bytes, void *comm, pic_stream_t stream);

/*
 * (1d) Sampled protocol of (1c): TWO collectives instead of four and ONE pass over the band instead of three.  Every
 * rank samples its band, the samples are all-gathered, every rank derives the same bracket pivots from the pooled
 * sample, sweeps its band once (count below, collect the bracket's elements), the counts and candidates are
 * all-gathered and every rank selects the exact order statistics among them: thresholds are bit-identical to (1c) and to
 * the single-device select.  A bracket that missed or an exchange slot that overflowed is seen identically by every
 * rank; only then the histogram rounds of (1c) run (*used_fallback = 1).  That decision needs a 4-byte read-back: this
 * entry synchronises `stream` once per call and cannot be captured in a CUDA graph ((1c) can).  Every rank must hold
 * at least one element of each unit.  ws: pic_tiled_sampled_workspace_bytes().
 */
size_t pic_tiled_sampled_workspace_bytes(int64_t n_local, int64_t n_total, int64_t units, int world_size);
int pic_tiled_select_threshold_sampled(const float *std_local, int64_t n_local, int64_t n_total, int64_t units, float q01,
                                       const float *q01_per_unit, float *thr_out, void *ws, size_t ws_bytes, void *comm,
                                       pic_stream_t stream, int *used_fallback);

/*
 * (1e) The sampled protocol of (1d) over PEER MEMORY instead of NCCL: every rank owns one window (cudaMalloc exported
 * with CUDA IPC, mapped by all peers of the node); an exchange is a kernel that stores this rank's rows straight into
 * every peer's window over NVLink / NVSwitch, raises a flag there and waits for the peers' flags -- a few microseconds
 * instead of a collective launch, and only the candidates that exist cross the links.  Nothing on the host happens
 * between the steps and nothing is read back, so the whole select (and the apply behind it) captures into one CUDA
 * graph.  The price is that the validity check moves to the caller: *status_dev (device word, the caller zeroes it) counts
 * the units whose bracket missed or whose slot overflowed in its low 16 bits and adds 0x10000 per exchange wait that
 * timed out (a peer that never arrived: 4 s by default, PIC_P2P_TIMEOUT_MS in the environment of pic_dist_p2p_init); when it is non-zero after the caller's next synchronisation the
 * thresholds of that call are not valid and (1c) has to be run instead -- every rank sees the same low 16 bits.
 * pic_dist_p2p_init / _destroy are collective over `comm` (they exchange the IPC handles and meet through it); the
 * regions are sized by pic_dist_p2p_region_bytes for the LARGEST (n_total, units) the window will serve.  One select at
 * a time per window; all ranks issue the same sequence of selects.  Single node only.
 */
int pic_dist_p2p_region_bytes(int64_t n_total, int64_t units, int world_size, size_t *sample_bytes, size_t *cand_bytes);
int pic_dist_p2p_init(void *comm, int rank, size_t sample_bytes, size_t cand_bytes, void **p2p_out);
int pic_dist_p2p_destroy(void *p2p);
int pic_tiled_select_threshold_p2p(const float *std_local, int64_t n_local, int64_t n_total, int64_t units, float q01,
                                   const float *q01_per_unit, float *thr_out, void *ws, size_t ws_bytes, void *p2p,
                                   uint32_t *status_dev, pic_stream_t stream);

/*
 * (2) ChannelMask.forward / ProgMask (channel_mask.py:18-49, 89-151): mask = (std >= thr) as
 * f32 {0,1}; ones / zeros for the sentinels.  thr_out nullable.
 */
int pic_channel_mask(const float *std, int64_t n_per_unit, int64_t units, float q01,
                     const float *q01_per_unit, float *mask, float *thr_out, void *ws,
                     size_t ws_bytes, pic_stream_t stream);

/*
 * REM attention mask (models/rem_pic.py:181-195; consumed at layers/rem.py:137-140): the star mask of
 * pic_channel_mask written `copies` times along the channel axis, out = [units][copies][n_per_unit]
 * (copies = 2 is torch.cat([m, m], 1) for the joint (mu, std) refinement; copies = 1 is pic_channel_mask).
 * One select + one pass over std instead of mask + `copies` tensor copies.  thr_out nullable.
 */
int pic_attention_mask(const float *std, int64_t n_per_unit, int64_t units, float q01,
                       const float *q01_per_unit, int copies, float *mask, float *thr_out, void *ws,
                       size_t ws_bytes, pic_stream_t stream);

/*
 * Elementwise neighbours of the path, one pass each instead of three torch kernels.
 * LRP epilogue + merge (models/pic.py:635-641, rem_pic.py equivalents):
 *     out = (y_hat + 0.5 * tanh(lrp)) + base          base nullable (no merge); out may alias y_hat
 *     backward: g_y_hat = g_base = g_out;  g_lrp = g_out * 0.5 * (1 - tanh(lrp)^2)
 * REM merge (layers/rem.py:137-140):
 *     out = identity + ret * att_mask                 backward: g_identity = g_out; g_ret = g_out * att_mask
 */
int pic_lrp_merge(const float *y_hat, const float *lrp, const float *base, float *out, int64_t n,
                  pic_stream_t stream);
int pic_lrp_merge_backward(const float *g_out, const float *lrp, float *g_lrp, int64_t n, pic_stream_t stream);
int pic_rem_merge(const float *identity, const float *ret, const float *att_mask, float *out, int64_t n,
                  pic_stream_t stream);
int pic_rem_merge_backward(const float *g_out, const float *att_mask, float *g_ret, int64_t n,
                           pic_stream_t stream);

/* mask = (std >= thr[u]) with thresholds already known (e.g. all-reduced ones). */
int pic_mask_from_threshold(const float *std, const float *thr, int64_t n_per_unit,
                            int64_t units, float *mask, pic_stream_t stream);

/*
 * (3) One progressive slice, fused (models/pic.py:583-584, 621-629 and 809-820):
 *   r      = y_top - y_base                (y_base NULL => r = y_top, delta_encode off)
 *   mask   = std >= thr                    (thr selected here, or taken from thr_in if non-NULL)
 *   y_m    = (r - mu) * mask ;  s_m = std * mask
 *   out    = noise ? y_m + noise : round(y_m)          (GaussianConditional.quantize)
 *   lik    = max(Phi((.5-|out|)/s) - Phi((-.5-|out|)/s), lik_bound),  s = max(s_m, scale_bound)
 *   y_hat  = round(r - mu) * mask + mu                 (ste_round forward; masked => mu)
 *   idx    = #{t in scale_table[:-1] : t < s}          (build_indexes)
 *   symbols= int32(round(y_m))                         (quantize(.., "symbols"))
 *   rate[u]= sum_i ln(lik[u,i])  (f64; training/loss.py:45-60 divides by -ln2*num_pixels)
 * Outputs mask, y_hat, lik, idx, symbols, thr_out, rate are each nullable.
 * scale_table: device pointer to table_len sorted f32 (GaussianConditional.scale_table).
 */
int pic_slice_forward(const float *y_top, const float *y_base, const float *mu, const float *std,
                      float q01, const float *q01_per_unit, const float *thr_in,
                      const float *noise, const float *scale_table, int table_len,
                      float scale_bound, float lik_bound, int64_t n_per_unit, int64_t units,
                      float *mask, float *y_hat, float *lik, int32_t *idx, int32_t *symbols,
                      float *thr_out, double *rate, void *ws, size_t ws_bytes,
                      pic_stream_t stream);

/*
 * (3a) Multi-quality form of (3): the quality sweep of ONE set of latents (train.py check_levels_np sweeps,
 * test/functions_encode.py:153-196 per-level loops; BASELINE config "quality sweep q = 0..1 in 100 steps").  For each
 * of `units` input units and each of `levels` qualities q01_levels[u][l] (device, level-minor): threshold, mask,
 * y_hat, likelihood, index, symbols and rate of output unit u * levels + l.  Inputs are [units][n]; every output is
 * [units][levels][n] (thr_out / rate [units][levels]); thr_out is required (it carries the thresholds from the select
 * launch to the apply launch).  The `levels` outputs of a unit read the SAME input block: inputs cross HBM once
 * (the rest are L2 hits), so the sweep is bound by its 16 B/element of output.  Evaluation mode (no noise).
 * n_per_unit <= pic_fused_max_elems().
 */
int pic_slice_forward_multi(const float *y_top, const float *y_base, const float *mu, const float *std,
                            const float *q01_levels, int levels, const float *scale_table, int table_len,
                            float scale_bound, float lik_bound, int64_t n_per_unit, int64_t units, float *mask,
                            float *y_hat, float *lik, int32_t *idx, int32_t *symbols, float *thr_out, double *rate,
                            pic_stream_t stream);

/*
 * (4) Backward of (3) (autograd of the same lines).  `mask` is the forward
 * mask; g_lik / g_yhat nullable (treated as zero); g_ybase nullable.  noise NULL => eval
 * forward (round() blocks the gradient into y_m).
 */
int pic_slice_backward(const float *g_lik, const float *g_yhat, const float *y_top,
                       const float *y_base, const float *mu, const float *std,
                       const float *mask, const float *noise, float scale_bound, float lik_bound,
                       int64_t n, float *g_ytop, float *g_ybase, float *g_mu, float *g_std,
                       pic_stream_t stream);

/*
 * (5) Un-fused operators, for API parity with GaussianConditional / EntropyModel.
 *  pic_gaussian_forward : forward(inputs, scales, means, training) entropy_models.py:637-652
 *                         (noise NULL => eval; outputs nullable => _likelihood only with
 *                         outputs := inputs when `likelihood_only` != 0, entropy_models.py:620-635)
 *  pic_gaussian_backward: its autograd (g_out, g_lik nullable)
 *  pic_build_indexes    : build_indexes, entropy_models.py:654-659
 *  pic_quantize         : quantize, entropy_models.py:127-153 (noise*mask when mask given)
 *  pic_dequantize       : dequantize, entropy_models.py:161-168 (int32 symbols + means)
 *  pic_log_sum          : per-unit sum ln(x) (rate numerator, training/loss.py:45-60)
 */
int pic_gaussian_forward(const float *inputs, const float *scales, const float *means,
                         const float *noise, int likelihood_only, int64_t n, float scale_bound,
                         float lik_bound, float *outputs, float *lik, pic_stream_t stream);
int pic_gaussian_backward(const float *g_out, const float *g_lik, const float *inputs,
                          const float *scales, const float *means, const float *noise,
                          int likelihood_only, int64_t n, float scale_bound, float lik_bound,
                          float *g_inputs, float *g_scales, float *g_means, pic_stream_t stream);
int pic_build_indexes(const float *scales, int64_t n, const float *scale_table, int table_len,
                      float scale_bound, int32_t *idx, pic_stream_t stream);
int pic_quantize(const float *inputs, const float *means, const float *noise, const float *mask,
                 int64_t n, int mode, float *out_f32, int32_t *out_i32, pic_stream_t stream);
int pic_dequantize(const int32_t *symbols, const float *means, int64_t n, float *out,
                   pic_stream_t stream);
int pic_log_sum(const float *x, int64_t n_per_unit, int64_t units, double *out,
                pic_stream_t stream);

/*
 * (5a) EntropyBottleneck for the hyper-latent z -- entropy_models/entropy_models.py:403-436 (_logits_cumulative,
 * _likelihood) and 449-492 (forward), with their autograd.  z / noise / outputs / lik: [batch, channels, spatial]
 * (the contiguous NCHW tensor; the reference's permute to [C, 1, B*S] and back is not needed).  noise NULL: eval,
 * outputs = round(z - median) + median; else outputs = z + noise (the caller draws the noise as the reference does).
 * medians: [channels] (quantiles[:, 0, 1]).  params: [channels][pic_bottleneck_params_per_channel()] RAW parameters
 * of the per-channel scalar network with filters (1, f1, f2, f3, f4, 1), packed per channel as all _matrix{i}
 * (row-major [f_{i+1}][f_i]), then all _bias{i}, then all _factor{i} (i = 0..3); softplus / tanh are applied inside.
 * lik_bound: likelihood lower bound (0 disables).  Backward: g_params [channels][per_channel] and g_medians (nullable)
 * are overwritten; g_z (nullable) is the gradient of z (identity through the noise in training, 0 through round()).
 */
int pic_bottleneck_params_per_channel(int f1, int f2, int f3, int f4);
int pic_bottleneck_forward(const float *z, const float *noise, const float *medians, const float *params, int f1, int f2,
                           int f3, int f4, int64_t batch, int64_t channels, int64_t spatial, float lik_bound,
                           float *outputs, float *lik, pic_stream_t stream);
int pic_bottleneck_backward(const float *z, const float *noise, const float *medians, const float *params, int f1, int f2,
                            int f3, int f4, int64_t batch, int64_t channels, int64_t spatial, float lik_bound,
                            const float *g_lik, const float *g_out, float *g_z, float *g_params, float *g_medians,
                            pic_stream_t stream);

/*
 * (6) Host-buffer form of (3) for callers whose latents live in host memory (the reference's
 * CPU path, or an FFI caller without device tensors).  All tensor pointers are HOST memory
 * (pinned memory makes the copies asynchronous); units are streamed through the device in
 * chunks with the copies overlapping the kernel.  Blocks until the outputs are in host memory.
 * `device_buf`/`device_buf_bytes`: caller-owned device scratch of at least
 * pic_host_pipeline_bytes(n_per_unit, chunk_units) bytes.
 */
size_t pic_host_pipeline_bytes(int64_t n_per_unit, int64_t chunk_units);
int pic_slice_forward_host(const float *y_top, const float *y_base, const float *mu,
                           const float *std, float q01, const float *q01_per_unit_host,
                           const float *noise, const float *scale_table_host, int table_len,
                           float scale_bound, float lik_bound, int64_t n_per_unit, int64_t units,
                           int64_t chunk_units, float *mask, float *y_hat, float *lik,
                           int32_t *idx, int32_t *symbols, float *thr_out, double *rate,
                           void *device_buf, size_t device_buf_bytes);

/*
 * (6a) Compact-output form of (6) for the codec (eval) forward: the mask is {0,1} and the scale index lies in
 * [0, table_len) with table_len <= 256 (64 in the reference, models/pic.py:12-18), so both return as ONE BYTE per
 * element -- 10 instead of 16 bytes per element cross PCIe on the way back.  y_hat / lik (nullable) stay f32.
 * Same pipeline, same scratch size as (6).
 */
int pic_slice_forward_host_compact(const float *y_top, const float *y_base, const float *mu, const float *std, float q01,
                                   const float *q01_per_unit_host, const float *scale_table_host, int table_len,
                                   float scale_bound, float lik_bound, int64_t n_per_unit, int64_t units,
                                   int64_t chunk_units, uint8_t *mask_u8, float *y_hat, float *lik, uint8_t *idx_u8,
                                   void *device_buf, size_t device_buf_bytes);

#ifdef __cplusplus
}
#endif
#endif /* PIC_LATENT_H_ */
