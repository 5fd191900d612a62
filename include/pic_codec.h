/*
 * pic_codec.h -- C ABI of the entropy-coder side of the codec: the step right after
 * the latent path.  HOST code (C++), exported by the same libpic_latent.so.
 *
 * Replaces, in the reference, the un-vendored CompressAI 1.2.4 C++ extension (environment.yml:203):
 *   compressai._CXX.pmf_to_quantized_cdf                        <- entropy_models.py:175-183 (_pmf_to_cdf)
 *   compressai.ans.RansEncoder().encode_with_indexes(...)       <- entropy_models.py:230-236 (compress)
 *   compressai.ans.RansDecoder().decode_with_indexes(...)       <- entropy_models.py:280-286 (decompress)
 * and the Python-list boundary around them (`.tolist()` of every symbol / index / CDF entry per call):
 * these entry points read int32 buffers in place (pinned host memory filled by one D2H copy of the
 * symbols / indexes the latent path produced) and code many streams on host threads.
 *
 * Bit-stream: rANS64 (ryg_rans rans64.h), 32-bit words, 16-bit CDF precision, 4-bit bypass escapes,
 * identical to CompressAI's rans_interface.cpp by construction.  compressai is not installed in this image:
 * byte parity with the real library is UNPINNED; parity is against oracle/rans_oracle.py plus round trips.
 */
#ifndef PIC_CODEC_H_
#define PIC_CODEC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cdf_out[0 .. len] (len + 1 entries) from pmf[0 .. len): returns 0, or PIC_ERR_INVALID_ARGUMENT for a negative /
 * non-finite / all-zero pmf or precision outside [1, 16] (the reference raises). */
int pic_pmf_to_quantized_cdf(const float *pmf, int len, int precision, int32_t *cdf_out);

/* Upper bound of the bytes one stream of n symbols can take (every symbol escaped). */
int64_t pic_rans_stream_bound(int64_t n);

/*
 * One stream.  cdfs: [n_cdfs][cdf_stride] int32 (row i valid for cdf_sizes[i] entries), offsets[n_cdfs].
 * symbols / indexes: n int32 each.  Returns the stream length in bytes (written to out[0 .. bytes)), or a
 * negative PIC_ERR_* (INVALID_ARGUMENT: index out of range / bad table; WORKSPACE: out_cap too small).
 */
int64_t pic_rans_encode_with_indexes(const int32_t *symbols, const int32_t *indexes, int64_t n,
                                     const int32_t *cdfs, int n_cdfs, int cdf_stride,
                                     const int32_t *cdf_sizes, const int32_t *offsets, uint8_t *out,
                                     int64_t out_cap);

/* Decodes n symbols of one stream into symbols_out; returns 0 or a negative PIC_ERR_*. */
int pic_rans_decode_with_indexes(const uint8_t *stream, int64_t nbytes, const int32_t *indexes, int64_t n,
                                 const int32_t *cdfs, int n_cdfs, int cdf_stride, const int32_t *cdf_sizes,
                                 const int32_t *offsets, int32_t *symbols_out);

/*
 * Many independent streams (the reference's loop `for i in range(symbols.size(0))`, entropy_models.py:229,279)
 * on up to `threads` host threads (0 = hardware concurrency).  symbols / indexes: [streams][n]; stream s is
 * written at out + s * out_stride and its length stored in out_bytes[s].  Returns 0 or the first error.
 */
int pic_rans_encode_batch(const int32_t *symbols, const int32_t *indexes, int64_t streams, int64_t n,
                          const int32_t *cdfs, int n_cdfs, int cdf_stride, const int32_t *cdf_sizes,
                          const int32_t *offsets, uint8_t *out, int64_t out_stride, int64_t *out_bytes,
                          int threads);
/* stream s is read from in + in_offsets[s], in_bytes[s] long; symbols_out: [streams][n]. */
int pic_rans_decode_batch(const uint8_t *in, const int64_t *in_offsets, const int64_t *in_bytes,
                          const int32_t *indexes, int64_t streams, int64_t n, const int32_t *cdfs, int n_cdfs,
                          int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                          int32_t *symbols_out, int threads);

/*
 * Progressive multi-level packing (test/functions_encode.py:176-196, functions_decode.py:186-206): the
 * reference codes, per level l, the tensors symbols * delta_l and indexes * delta_l with
 * delta_l = ProgMask(q_l) - ProgMask(q_{l-1}).  With the level map of pic_level_map (level[i] = l  <=>
 * delta_l[i] = 1) the selection happens inside the coder: stream (l, s) codes, for i in [0, n),
 * (level[s][i] == l ? symbols[s][i] : 0,  level[s][i] == l ? indexes[s][i] : 0) -- byte-identical to the
 * reference's per-level streams -- so symbols, indexes and level cross PCIe ONCE for all levels.
 * Streams are ordered [level][stream]; stream (l, s) is written at out + (l * streams + s) * out_stride and
 * its length stored in out_bytes[l * streams + s].  Levels [level_begin, level_end) are coded.
 */
int pic_rans_encode_levels(const int32_t *symbols, const int32_t *indexes, const int32_t *level,
                           int64_t streams, int64_t n, int level_begin, int level_end, const int32_t *cdfs,
                           int n_cdfs, int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                           uint8_t *out, int64_t out_stride, int64_t *out_bytes, int threads);
/*
 * Decoder side: streams of levels [level_begin, level_end) (same ordering; in_offsets / in_bytes indexed
 * [(l - level_begin) * streams + s]) -> symbols_out[s][i] = decoded symbol where level[s][i] is one of those
 * levels, untouched elsewhere (initialise it to zero, or to the symbols of the levels already received).
 */
int pic_rans_decode_levels(const uint8_t *in, const int64_t *in_offsets, const int64_t *in_bytes,
                           const int32_t *indexes, const int32_t *level, int64_t streams, int64_t n,
                           int level_begin, int level_end, const int32_t *cdfs, int n_cdfs, int cdf_stride,
                           const int32_t *cdf_sizes, const int32_t *offsets, int32_t *symbols_out, int threads);

#ifdef __cplusplus
}
#endif
#endif /* PIC_CODEC_H_ */
