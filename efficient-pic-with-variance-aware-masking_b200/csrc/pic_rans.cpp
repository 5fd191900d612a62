// pic_rans.cpp -- host side of the codec right after the latent path (include/pic_codec.h):
// quantised-CDF build and rANS coding of (symbol, CDF index) streams.
//
// Restates the published algorithms the reference reaches through CompressAI 1.2.4's C++ extension
// (un-vendored; call sites entropy_models.py:175-183, 230-236, 280-286): ryg_rans' rans64.h coder
// with CompressAI's 4-bit bypass escapes, and the pmf -> 16-bit CDF normalisation.  Differences
// from that extension are structural only: symbols / indexes / tables are read from int32 buffers
// in place (no Python lists), the encoder walks the symbols backwards without materialising the
// intermediate (start, range) vector, the decoder finds the symbol by binary search, and many
// streams are coded on host threads.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/pic_codec.h"
#include "../../include/pic_latent.h"

namespace {

constexpr int kPrecision = 16;
constexpr int kBypassBits = 4;
constexpr int32_t kMaxBypass = (1 << kBypassBits) - 1;
constexpr uint64_t kRansL = 1ull << 31;

struct Tables {
    const int32_t *cdfs;
    int n_cdfs, stride;
    const int32_t *sizes, *offsets;
};

bool tables_ok(const Tables &t) {
    if (!t.cdfs || !t.sizes || !t.offsets || t.n_cdfs <= 0 || t.stride < 2) return false;
    for (int i = 0; i < t.n_cdfs; ++i)
        if (t.sizes[i] < 2 || t.sizes[i] > t.stride) return false;
    return true;
}

// Encoder writing 32-bit words from the back of a buffer.
struct Writer {
    uint32_t *begin, *ptr;
    bool overflow = false;
    inline void put(uint32_t w) {
        if (ptr == begin) { overflow = true; return; }
        *--ptr = w;
    }
};

inline void enc_put(uint64_t &x, Writer &w, uint32_t start, uint32_t freq) {
    const uint64_t x_max = ((kRansL >> kPrecision) << 32) * freq;
    if (x >= x_max) {
        w.put(static_cast<uint32_t>(x));
        x >>= 32;
    }
    x = ((x / freq) << kPrecision) + (x % freq) + start;
}

inline void enc_put_bits(uint64_t &x, Writer &w, uint32_t val) {
    constexpr uint32_t freq = 1u << (16 - kBypassBits);
    constexpr uint64_t x_max = ((kRansL >> 16) << 32) * freq;
    if (x >= x_max) {
        w.put(static_cast<uint32_t>(x));
        x >>= 32;
    }
    x = (x << kBypassBits) | val;
}

// One (start, freq) pair prepared for division-free coding (Alverson reciprocal, as in ryg_rans'
// Rans64EncSymbolInit): x' = x + bias + mulhi(x, rcp) >> shift * (2^16 - freq)  ==  ((x / freq) << 16) + x % freq + start.
struct FastSym {
    uint64_t rcp, x_max;
    uint32_t shift, bias, cmpl;
};
FastSym fast_sym(uint32_t start, uint32_t freq) {
    FastSym f;
    f.x_max = ((kRansL >> kPrecision) << 32) * freq;
    f.cmpl = (1u << kPrecision) - freq;
    if (freq < 2) {   // 1 / 1 is not representable: mulhi(x, ~0) = x - 1, compensated in the bias
        f.rcp = ~0ull;
        f.shift = 0;
        f.bias = start + (1u << kPrecision) - 1;
    } else {
        uint32_t sh = 0;
        while (freq > (1u << sh)) ++sh;
        f.rcp = static_cast<uint64_t>(((static_cast<unsigned __int128>(1) << (sh + 63)) + freq - 1) / freq);
        f.shift = sh - 1;
        f.bias = start;
    }
    return f;
}
inline void enc_put_fast(uint64_t &x, Writer &w, const FastSym &f) {
    if (x >= f.x_max) {
        w.put(static_cast<uint32_t>(x));
        x >>= 32;
    }
    const uint64_t q = static_cast<uint64_t>((static_cast<unsigned __int128>(x) * f.rcp) >> 64) >> f.shift;
    x = x + f.bias + q * f.cmpl;
}

// level != nullptr: progressive level `lv` -- elements of other levels are coded as (symbol 0, index 0), which is
// what the reference's symbols * delta / indexes * delta tensors hold there (functions_encode.py:186-190).  That one
// pair is 7/8 of all coding steps at 8 levels, so it takes the division-free form.
int64_t encode_stream(const int32_t *symbols, const int32_t *indexes, int64_t n, const Tables &t, uint8_t *out,
                      int64_t out_cap, const int32_t *level = nullptr, int32_t lv = 0) {
    if (out_cap < 8 || (reinterpret_cast<uintptr_t>(out) & 3u)) return PIC_ERR_WORKSPACE;
    Writer w;
    w.begin = reinterpret_cast<uint32_t *>(out);
    w.ptr = w.begin + out_cap / 4;
    uint32_t *const end = w.ptr;
    uint64_t x = kRansL;
    FastSym zero{};
    bool zero_ok = false;   // (symbol 0, index 0) inside table 0's regular range?
    if (level) {
        const int64_t v0 = -static_cast<int64_t>(t.offsets[0]);
        if (v0 >= 0 && v0 < t.sizes[0] - 2 && t.cdfs[v0 + 1] > t.cdfs[v0]) {
            zero = fast_sym(static_cast<uint32_t>(t.cdfs[v0]), static_cast<uint32_t>(t.cdfs[v0 + 1] - t.cdfs[v0]));
            zero_ok = true;
        }
    }
    for (int64_t i = n - 1; i >= 0; --i) {
        const bool on = !level || level[i] == lv;
        if (!on && zero_ok) {
            enc_put_fast(x, w, zero);
            continue;
        }
        const int32_t ci = on ? indexes[i] : 0;
        if (ci < 0 || ci >= t.n_cdfs) return PIC_ERR_INVALID_ARGUMENT;
        const int32_t *cdf = t.cdfs + static_cast<int64_t>(ci) * t.stride;
        const int32_t max_value = t.sizes[ci] - 2;
        int64_t value = static_cast<int64_t>(on ? symbols[i] : 0) - t.offsets[ci];
        uint64_t raw = 0;
        bool escaped = false;
        if (value < 0) {
            raw = static_cast<uint64_t>(-2 * value - 1);
            value = max_value;
            escaped = true;
        } else if (value >= max_value) {
            raw = static_cast<uint64_t>(2 * (value - max_value));
            value = max_value;
            escaped = true;
        }
        if (escaped) {   // reverse of: [count nibbles][raw nibbles, low first]
            int n_bypass = 0;
            while ((raw >> (n_bypass * kBypassBits)) != 0) ++n_bypass;
            for (int j = n_bypass - 1; j >= 0; --j) enc_put_bits(x, w, static_cast<uint32_t>((raw >> (j * kBypassBits)) & kMaxBypass));
            const int full = n_bypass / kMaxBypass;          // leading 15s, then the remainder
            enc_put_bits(x, w, static_cast<uint32_t>(n_bypass - full * kMaxBypass));
            for (int j = 0; j < full; ++j) enc_put_bits(x, w, static_cast<uint32_t>(kMaxBypass));
        }
        const uint32_t start = static_cast<uint32_t>(cdf[value]);
        const uint32_t freq = static_cast<uint32_t>(cdf[value + 1] - cdf[value]);
        if (freq == 0 || freq > (1u << kPrecision)) return PIC_ERR_INVALID_ARGUMENT;
        enc_put(x, w, start, freq);
    }
    w.put(static_cast<uint32_t>(x >> 32));
    w.put(static_cast<uint32_t>(x));
    if (w.overflow) return PIC_ERR_WORKSPACE;
    const int64_t bytes = (end - w.ptr) * 4;
    std::memmove(out, w.ptr, static_cast<size_t>(bytes));
    return bytes;
}

struct Reader {
    const uint32_t *ptr, *end;
    bool underflow = false;
    inline uint32_t get() {
        if (ptr == end) { underflow = true; return 0; }
        return *ptr++;
    }
};

inline uint32_t dec_get_bits(uint64_t &x, Reader &r) {
    const uint32_t val = static_cast<uint32_t>(x & ((1u << kBypassBits) - 1));
    x >>= kBypassBits;
    if (x < kRansL) x = (x << 32) | r.get();
    return val;
}

int decode_stream(const uint8_t *stream, int64_t nbytes, const int32_t *indexes, int64_t n, const Tables &t,
                  int32_t *out, const int32_t *level = nullptr, int32_t lv = 0) {
    if (nbytes < 8 || (nbytes & 3) || (reinterpret_cast<uintptr_t>(stream) & 3u)) return PIC_ERR_INVALID_ARGUMENT;
    Reader r;
    r.ptr = reinterpret_cast<const uint32_t *>(stream);
    r.end = r.ptr + nbytes / 4;
    uint64_t x = static_cast<uint64_t>(r.ptr[0]) | (static_cast<uint64_t>(r.ptr[1]) << 32);
    r.ptr += 2;
    constexpr uint64_t mask = (1ull << kPrecision) - 1;
    // elements of other levels were coded as (symbol 0, index 0): when the state's low bits fall in that slot (they
    // do, for a valid stream) the table search is skipped
    uint32_t z_start = 0, z_freq = 0;
    if (level) {
        const int64_t v0 = -static_cast<int64_t>(t.offsets[0]);
        if (v0 >= 0 && v0 < t.sizes[0] - 2 && t.cdfs[v0 + 1] > t.cdfs[v0]) {
            z_start = static_cast<uint32_t>(t.cdfs[v0]);
            z_freq = static_cast<uint32_t>(t.cdfs[v0 + 1] - t.cdfs[v0]);
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        const bool on = !level || level[i] == lv;
        if (!on && z_freq != 0) {
            const uint32_t rel = static_cast<uint32_t>(x & mask) - z_start;
            if (rel < z_freq) {
                x = z_freq * (x >> kPrecision) + rel;
                if (x < kRansL) x = (x << 32) | r.get();
                if (r.underflow) return PIC_ERR_INVALID_ARGUMENT;
                continue;
            }
        }
        const int32_t ci = on ? indexes[i] : 0;
        if (ci < 0 || ci >= t.n_cdfs) return PIC_ERR_INVALID_ARGUMENT;
        const int32_t *cdf = t.cdfs + static_cast<int64_t>(ci) * t.stride;
        const int32_t size = t.sizes[ci];
        const int32_t max_value = size - 2;
        const int32_t cum = static_cast<int32_t>(x & mask);
        // first entry strictly greater than cum, minus one (find_if in the reference extension)
        const int32_t s = static_cast<int32_t>(std::upper_bound(cdf, cdf + size, cum) - cdf) - 1;
        if (s < 0 || s > max_value) return PIC_ERR_INVALID_ARGUMENT;
        const uint32_t start = static_cast<uint32_t>(cdf[s]);
        const uint32_t freq = static_cast<uint32_t>(cdf[s + 1] - cdf[s]);
        x = freq * (x >> kPrecision) + (x & mask) - start;
        if (x < kRansL) x = (x << 32) | r.get();
        int64_t value = s;
        if (s == max_value) {
            int32_t val = static_cast<int32_t>(dec_get_bits(x, r));
            int32_t n_bypass = val;
            while (val == kMaxBypass) {
                val = static_cast<int32_t>(dec_get_bits(x, r));
                n_bypass += val;
                if (r.underflow) return PIC_ERR_INVALID_ARGUMENT;
            }
            if (n_bypass > 16) return PIC_ERR_INVALID_ARGUMENT;   // more than 64 raw bits: corrupt stream
            uint64_t raw = 0;
            for (int j = 0; j < n_bypass; ++j) raw |= static_cast<uint64_t>(dec_get_bits(x, r)) << (j * kBypassBits);
            value = static_cast<int64_t>(raw >> 1);
            if (raw & 1) value = -value - 1;
            else value += max_value;
        }
        if (r.underflow) return PIC_ERR_INVALID_ARGUMENT;
        if (on) out[i] = static_cast<int32_t>(value + t.offsets[ci]);
    }
    return PIC_OK;
}

template <class Fn>
int run_streams(int64_t streams, int threads, Fn fn) {
    int nt = threads > 0 ? threads : static_cast<int>(std::thread::hardware_concurrency());
    if (nt < 1) nt = 1;
    if (nt > streams) nt = static_cast<int>(streams);
    std::vector<int> rc(static_cast<size_t>(nt), PIC_OK);
    auto work = [&](int tix) {
        for (int64_t s = tix; s < streams; s += nt) {
            const int r = fn(s);
            if (r != PIC_OK && rc[tix] == PIC_OK) rc[tix] = r;
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        pool.reserve(static_cast<size_t>(nt));
        for (int i = 0; i < nt; ++i) pool.emplace_back(work, i);
        for (auto &th : pool) th.join();
    }
    for (int r : rc)
        if (r != PIC_OK) return r;
    return PIC_OK;
}

}  // namespace

extern "C" {

int pic_pmf_to_quantized_cdf(const float *pmf, int len, int precision, int32_t *cdf_out) {
    if (!pmf || !cdf_out || len < 1 || precision < 1 || precision > 16) return PIC_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < len; ++i)
        if (pmf[i] < 0 || !std::isfinite(pmf[i])) return PIC_ERR_INVALID_ARGUMENT;
    std::vector<uint32_t> cdf(static_cast<size_t>(len) + 1);
    cdf[0] = 0;
    const float scale = static_cast<float>(1 << precision);
    uint32_t total = 0;
    for (int i = 0; i < len; ++i) {
        cdf[i + 1] = static_cast<uint32_t>(std::round(pmf[i] * scale));
        total += cdf[i + 1];
    }
    if (total == 0) return PIC_ERR_INVALID_ARGUMENT;
    for (auto &c : cdf) c = static_cast<uint32_t>((static_cast<uint64_t>(1u << precision) * c) / total);
    for (int i = 1; i <= len; ++i) cdf[i] += cdf[i - 1];
    cdf[len] = 1u << precision;
    for (int i = 0; i < len; ++i) {
        if (cdf[i] != cdf[i + 1]) continue;
        // a zero-width slot: steal one count from the narrowest slot that can spare it
        uint32_t best_freq = ~0u;
        int best = -1;
        for (int j = 0; j < len; ++j) {
            const uint32_t freq = cdf[j + 1] - cdf[j];
            if (freq > 1 && freq < best_freq) {
                best_freq = freq;
                best = j;
            }
        }
        if (best < 0) return PIC_ERR_INVALID_ARGUMENT;   // more slots than 2^precision counts
        if (best < i) {
            for (int j = best + 1; j <= i; ++j) --cdf[j];
        } else {
            for (int j = i + 1; j <= best; ++j) ++cdf[j];
        }
    }
    for (int i = 0; i <= len; ++i) cdf_out[i] = static_cast<int32_t>(cdf[i]);
    return PIC_OK;
}

int64_t pic_rans_stream_bound(int64_t n) { return n < 0 ? 0 : 16 + 8 * n; }

int64_t pic_rans_encode_with_indexes(const int32_t *symbols, const int32_t *indexes, int64_t n, const int32_t *cdfs,
                                     int n_cdfs, int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                                     uint8_t *out, int64_t out_cap) {
    const Tables t{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    if (n < 0 || (n > 0 && (!symbols || !indexes)) || !out || !tables_ok(t)) return PIC_ERR_INVALID_ARGUMENT;
    return encode_stream(symbols, indexes, n, t, out, out_cap);
}

int pic_rans_decode_with_indexes(const uint8_t *stream, int64_t nbytes, const int32_t *indexes, int64_t n,
                                 const int32_t *cdfs, int n_cdfs, int cdf_stride, const int32_t *cdf_sizes,
                                 const int32_t *offsets, int32_t *symbols_out) {
    const Tables t{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    if (n < 0 || !stream || (n > 0 && (!indexes || !symbols_out)) || !tables_ok(t)) return PIC_ERR_INVALID_ARGUMENT;
    return decode_stream(stream, nbytes, indexes, n, t, symbols_out);
}

int pic_rans_encode_batch(const int32_t *symbols, const int32_t *indexes, int64_t streams, int64_t n,
                          const int32_t *cdfs, int n_cdfs, int cdf_stride, const int32_t *cdf_sizes,
                          const int32_t *offsets, uint8_t *out, int64_t out_stride, int64_t *out_bytes, int threads) {
    const Tables t{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    if (streams <= 0 || n < 0 || !symbols || !indexes || !out || !out_bytes || (out_stride & 3) || !tables_ok(t))
        return PIC_ERR_INVALID_ARGUMENT;
    return run_streams(streams, threads, [&](int64_t s) {
        const int64_t b = encode_stream(symbols + s * n, indexes + s * n, n, t, out + s * out_stride, out_stride);
        out_bytes[s] = b < 0 ? 0 : b;
        return b < 0 ? static_cast<int>(b) : PIC_OK;
    });
}

int pic_rans_decode_batch(const uint8_t *in, const int64_t *in_offsets, const int64_t *in_bytes,
                          const int32_t *indexes, int64_t streams, int64_t n, const int32_t *cdfs, int n_cdfs,
                          int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets, int32_t *symbols_out,
                          int threads) {
    const Tables t{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    if (streams <= 0 || n < 0 || !in || !in_offsets || !in_bytes || !indexes || !symbols_out || !tables_ok(t))
        return PIC_ERR_INVALID_ARGUMENT;
    return run_streams(streams, threads, [&](int64_t s) {
        return decode_stream(in + in_offsets[s], in_bytes[s], indexes + s * n, n, t, symbols_out + s * n);
    });
}

int pic_rans_encode_levels(const int32_t *symbols, const int32_t *indexes, const int32_t *level, int64_t streams,
                           int64_t n, int level_begin, int level_end, const int32_t *cdfs, int n_cdfs, int cdf_stride,
                           const int32_t *cdf_sizes, const int32_t *offsets, uint8_t *out, int64_t out_stride,
                           int64_t *out_bytes, int threads) {
    const Tables t{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    if (streams <= 0 || n < 0 || level_begin < 0 || level_end <= level_begin || !symbols || !indexes || !level || !out ||
        !out_bytes || (out_stride & 3) || !tables_ok(t))
        return PIC_ERR_INVALID_ARGUMENT;
    const int64_t tasks = static_cast<int64_t>(level_end - level_begin) * streams;
    return run_streams(tasks, threads, [&](int64_t k) {
        const int64_t s = k % streams;
        const int32_t lv = level_begin + static_cast<int32_t>(k / streams);
        const int64_t b = encode_stream(symbols + s * n, indexes + s * n, n, t, out + k * out_stride, out_stride,
                                        level + s * n, lv);
        out_bytes[k] = b < 0 ? 0 : b;
        return b < 0 ? static_cast<int>(b) : PIC_OK;
    });
}

int pic_rans_decode_levels(const uint8_t *in, const int64_t *in_offsets, const int64_t *in_bytes,
                           const int32_t *indexes, const int32_t *level, int64_t streams, int64_t n, int level_begin,
                           int level_end, const int32_t *cdfs, int n_cdfs, int cdf_stride, const int32_t *cdf_sizes,
                           const int32_t *offsets, int32_t *symbols_out, int threads) {
    const Tables t{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    if (streams <= 0 || n < 0 || level_begin < 0 || level_end <= level_begin || !in || !in_offsets || !in_bytes ||
        !indexes || !level || !symbols_out || !tables_ok(t))
        return PIC_ERR_INVALID_ARGUMENT;
    // levels of one stream write disjoint elements of symbols_out[s], so (level, stream) tasks are independent
    const int64_t tasks = static_cast<int64_t>(level_end - level_begin) * streams;
    return run_streams(tasks, threads, [&](int64_t k) {
        const int64_t s = k % streams;
        const int32_t lv = level_begin + static_cast<int32_t>(k / streams);
        return decode_stream(in + in_offsets[k], in_bytes[k], indexes + s * n, n, t, symbols_out + s * n, level + s * n, lv);
    });
}

}  // extern "C"
