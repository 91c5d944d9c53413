// jls_codec.cuh -- the JPEG-LS scan codec as per-thread device code.
//
// Everything here is `__host__ __device__` so that the exact code the kernels run can also be driven on the CPU by
// the unit tests (tests/hostemu); the product only ever calls it from the kernels in jls_kernels.cu.
//
// Two codecs are built from one set of primitives:
//   * FastLineEncoder / FastLineDecoder (jls_fast.cuh) -- restart interval = 1 line.  The previous line is all zeros
//     at every restart (reference src/scan_decoder_impl.hpp:122-127), so Rb = Rc = Rd = 0, the gradient vector
//     collapses to (0, 0, -Ra) and only contexts 0..4 exist: the whole adaptive state of a line lives in registers plus
//     80 bytes of shared memory per thread.  One thread codes one line; a warp codes 32 lines.
//   * GeneralIntervalEncoder / GeneralIntervalDecoder -- any restart interval (including none): full 2-D LOCO-I with
//     365 contexts, one thread per restart interval.  Needed to decode every conformant stream and to write
//     streams that are byte-identical to the reference's (restart interval 0).
//
// Behaviour follows the reference (team-charls/charls @ 7b9b2da); citations give the file:line restated.
#pragma once

#include "jls_common.h"

#if defined(__CUDACC__)
#define JLS_HD __host__ __device__ __forceinline__
#define JLS_HD_NOINLINE __host__ __device__ __noinline__
#else
#define JLS_HD inline
#define JLS_HD_NOINLINE
#endif

namespace jls {

// ---------------------------------------------------------------------------------------------------------------------
// bit tricks
// ---------------------------------------------------------------------------------------------------------------------
JLS_HD int clz32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __clz(static_cast<int>(v));
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

JLS_HD int clz64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __clzll(static_cast<long long>(v));
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}

JLS_HD uint32_t bswap32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0, 0x0123);
#else
    return __builtin_bswap32(v);
#endif
}

// low 32 bits of (hi:lo) >> shift, shift in [0, 31]
JLS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t shift)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift);
#else
    return shift ? (lo >> shift) | (hi << (32 - shift)) : lo;
#endif
}

// high 32 bits of (hi:lo) << shift, shift in [0, 32]
JLS_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t shift)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_lc(lo, hi, shift);
#else
    return shift == 0 ? hi : shift >= 32 ? lo : (hi << shift) | (lo >> (32 - shift));
#endif
}

// shifts whose count may reach 32 (the result is then 0): PTX defines that, C++ does not
JLS_HD uint32_t shl_sat(uint32_t v, uint32_t count)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(count));
    return r;
#else
    return count >= 32 ? 0U : v << count;
#endif
}

JLS_HD uint32_t shr_sat(uint32_t v, uint32_t count)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(count));
    return r;
#else
    return count >= 32 ? 0U : v >> count;
#endif
}

// the same for 64-bit values, count in [0, 64]
JLS_HD uint64_t shl64_sat(uint64_t v, uint32_t count)
{
#if defined(__CUDA_ARCH__)
    uint64_t r;
    asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(v), "r"(count));
    return r;
#else
    return count >= 64 ? 0U : v << count;
#endif
}

JLS_HD uint64_t shr64_sat(uint64_t v, uint32_t count)
{
#if defined(__CUDA_ARCH__)
    uint64_t r;
    asm("shr.u64 %0, %1, %2;" : "=l"(r) : "l"(v), "r"(count));
    return r;
#else
    return count >= 64 ? 0U : v >> count;
#endif
}

JLS_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

// 0x80 in every byte of w that is 0xFF (exact per byte: the addition never carries out of a byte), 0 elsewhere
JLS_HD uint32_t has_ff_byte(uint32_t w)
{
    return ((w & 0x7F7F7F7FU) + 0x01010101U) & w & 0x80808080U;
}

JLS_HD int32_t popcount32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

JLS_HD int32_t iabs(int32_t v) { return v < 0 ? -v : v; }
JLS_HD int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
JLS_HD int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
JLS_HD uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

// reference src/jpegls_algorithm.hpp:91-116
JLS_HD int32_t bit_wise_sign(int32_t v) { return v >> 31; }
JLS_HD int32_t sign_of(int32_t v) { return (v >> 31) | 1; }
JLS_HD int32_t apply_sign(int32_t v, int32_t s) { return (s ^ v) - s; }

// reference src/jpegls_algorithm.hpp:68-87 (T.87 A.5.2)
JLS_HD int32_t map_error_value(int32_t e) { return (e >> 30) ^ (2 * e); }
JLS_HD int32_t unmap_error_value(int32_t m) { return (-(m & 1)) ^ (m >> 1); }

// J[] of T.87 A.2.1 (reference src/scan_codec.hpp:18-19) in closed form:
// {0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,5,5,6,6,7,7,8,9,10,11,12,13,14,15}
JLS_HD int32_t run_order(int32_t run_index)
{
    return run_index < 16 ? (run_index >> 2) : (run_index < 24 ? (run_index >> 1) - 4 : run_index - 16);
}

// ---------------------------------------------------------------------------------------------------------------------
// sample arithmetic (reference src/default_traits.hpp, src/lossless_traits.hpp)
// ---------------------------------------------------------------------------------------------------------------------

// reference src/default_traits.hpp:111-117
JLS_HD int32_t correct_prediction(const CodecParams& p, int32_t predicted)
{
    if ((predicted & p.maxval) == predicted)
        return predicted;
    return (~(predicted >> 31)) & p.maxval;
}

// reference src/default_traits.hpp:65-68,123-139,157-163 ; lossless: src/lossless_traits.hpp:47-65
template<bool LOSSLESS>
JLS_HD int32_t compute_error_value(const CodecParams& p, int32_t e)
{
    if (LOSSLESS)
    {
        const int32_t shift = 32 - p.bits_per_sample;
        return static_cast<int32_t>(static_cast<uint32_t>(e) << shift) >> shift;
    }
    // quantize: (e + NEAR) / (2 NEAR + 1) for e > 0, -((NEAR - e) / (2 NEAR + 1)) otherwise; |e| <= 65535
    const uint32_t magnitude = static_cast<uint32_t>(iabs(e) + p.near);
    int32_t q = static_cast<int32_t>(mulhi32(magnitude, p.dq_magic));
    q = e > 0 ? q : -q;
    if (q < 0)
        q += p.range;
    if (q >= (p.range + 1) / 2)
        q -= p.range;
    return q;
}

// reference src/default_traits.hpp:77-81,166-184 ; lossless: src/lossless_traits.hpp:67-72
template<bool LOSSLESS>
JLS_HD int32_t reconstruct(const CodecParams& p, int32_t predicted, int32_t error_value)
{
    if (LOSSLESS)
        return (predicted + error_value) & p.maxval;
    int32_t v = predicted + error_value * p.dq;
    if (v < -p.near)
        v += p.range_dq;
    else if (v > p.maxval + p.near)
        v -= p.range_dq;
    return correct_prediction(p, v);
}

// reference src/jpegls_algorithm.hpp:173-194
JLS_HD int32_t quantize_gradient(const CodecParams& p, int32_t di)
{
    if (di <= -p.t3)
        return -4;
    if (di <= -p.t2)
        return -3;
    if (di <= -p.t1)
        return -2;
    if (di < -p.near)
        return -1;
    if (di <= p.near)
        return 0;
    if (di < p.t1)
        return 1;
    if (di < p.t2)
        return 2;
    if (di < p.t3)
        return 3;
    return 4;
}

// reference src/jpegls_algorithm.hpp:144-161
JLS_HD int32_t predict_med(int32_t ra, int32_t rb, int32_t rc)
{
    const int32_t sgn = bit_wise_sign(rb - ra);
    if ((sgn ^ (rc - ra)) < 0)
        return rb;
    if ((sgn ^ (rb - rc)) < 0)
        return ra;
    return ra + rb - rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// adaptive contexts (reference src/regular_mode_context.hpp, src/run_mode_context.hpp)
// ---------------------------------------------------------------------------------------------------------------------
struct alignas(16) RegularContext
{
    int32_t a, b, c, n;
};

struct RunContext
{
    int32_t a, n, nn;
};

// min k with (n << k) >= a ; reference src/regular_mode_context.hpp:99-136 (clz form :121-136)
JLS_HD int32_t golomb_parameter(int32_t a, int32_t n)
{
    int32_t k = imax(0, clz32(static_cast<uint32_t>(n)) - clz32(static_cast<uint32_t>(a)));
    if ((n << k) < a)
        ++k;
    return k;
}

// reference src/regular_mode_context.hpp:36-42
JLS_HD int32_t error_correction(const RegularContext& c, int32_t k_or_near)
{
    return k_or_near != 0 ? 0 : bit_wise_sign(2 * c.b + c.n - 1);
}

// reference src/regular_mode_context.hpp:45-94 (T.87 A.12, A.13); returns false on the invalid_data condition (:52-54)
JLS_HD bool update_regular_context(RegularContext& c, int32_t e, int32_t dq, int32_t reset)
{
    c.a += iabs(e);
    c.b += e * dq;
    const bool ok = c.a < 65536 * 256 && iabs(c.b) < 65536 * 256;
    if (c.n == reset)
    {
        c.a >>= 1;
        c.b >>= 1;
        c.n >>= 1;
    }
    ++c.n;
    if (c.b + c.n <= 0)
    {
        c.b += c.n;
        if (c.b <= -c.n)
            c.b = -c.n + 1;
        if (c.c > -128)
            --c.c;
    }
    else if (c.b > 0)
    {
        c.b -= c.n;
        if (c.b > 0)
            c.b = 0;
        if (c.c < 127)
            ++c.c;
    }
    return ok;
}

// reference src/run_mode_context.hpp:36-61
JLS_HD int32_t run_golomb_parameter(const RunContext& c, int32_t ri_type)
{
    const int32_t temp = c.a + (c.n >> 1) * ri_type;
    return golomb_parameter(imax(temp, 1), c.n);
}

// reference src/run_mode_context.hpp:102-115 (T.87 A.21)
JLS_HD int32_t run_compute_map(const RunContext& c, int32_t e, int32_t k)
{
    if (k == 0 && e > 0 && 2 * c.nn < c.n)
        return 1;
    if (e < 0 && 2 * c.nn >= c.n)
        return 1;
    if (e < 0 && k != 0)
        return 1;
    return 0;
}

// reference src/run_mode_context.hpp:64-82 (T.87 A.23)
JLS_HD void update_run_context(RunContext& c, int32_t e, int32_t e_mapped, int32_t ri_type, int32_t reset)
{
    if (e < 0)
        ++c.nn;
    c.a += (e_mapped + 1 - ri_type) >> 1;
    if (c.n == reset)
    {
        c.a >>= 1;
        c.n >>= 1;
        c.nn >>= 1;
    }
    ++c.n;
}

// reference src/run_mode_context.hpp:84-99
JLS_HD int32_t run_error_value(const RunContext& c, int32_t temp, int32_t k)
{
    const int32_t map = temp & 1;
    const int32_t e_abs = (temp + map) >> 1;
    const int32_t cond = (k != 0 || (2 * c.nn >= c.n)) ? 1 : 0;
    return cond == map ? -e_abs : e_abs;
}

// ---------------------------------------------------------------------------------------------------------------------
// Bit writer: MSB first; after a 0xFF byte the next byte carries only 7 bits (T.87 A.1; reference
// src/scan_encoder.hpp:75-180).  Bytes are staged in a register and written as aligned 32-bit words.
// ---------------------------------------------------------------------------------------------------------------------
struct BitWriter
{
    uint64_t acc;   // pending bits in the low `nbits` bits
    int32_t nbits;  // < 32 between calls
    uint32_t pend;  // pending output bytes (low `npend` bytes, oldest most significant)
    int32_t npend;  // 0..3
    bool prev_ff;   // last emitted byte was 0xFF
    bool overflow;  // ran out of destination
    uint32_t* wp;   // next aligned word
    uint32_t* wend; // end of destination
    uint32_t* base;

    JLS_HD void init(uint8_t* destination, size_t capacity) // destination 4-byte aligned
    {
        acc = 0;
        nbits = 0;
        pend = 0;
        npend = 0;
        prev_ff = false;
        overflow = false;
        base = reinterpret_cast<uint32_t*>(destination);
        wp = base;
        wend = base + capacity / 4;
    }

    JLS_HD void store_word(uint32_t big_endian_value)
    {
        if (wp < wend)
            *wp = bswap32(big_endian_value);
        else
            overflow = true;
        ++wp;
    }

    JLS_HD void emit_byte(uint32_t b)
    {
        pend = (pend << 8) | b;
        if (++npend == 4)
        {
            store_word(pend);
            npend = 0;
        }
    }

    JLS_HD void emit_word(uint32_t w)
    {
        store_word(funnel_r(w, pend, static_cast<uint32_t>(8 * npend)));
        pend = w;
    }

    JLS_HD void emit_one_stuffed_byte()
    {
        const int32_t take = prev_ff ? 7 : 8;
        const uint32_t b = static_cast<uint32_t>(acc >> (nbits - take)) & (0xFFU >> (8 - take));
        nbits -= take;
        emit_byte(b);
        prev_ff = (b == 0xFFU);
    }

    // value < 2^count, count in [0, 32]
    JLS_HD void put(uint32_t value, int32_t count)
    {
        acc = (acc << count) | value;
        nbits += count;
        if (nbits >= 32)
        {
            const uint32_t w = static_cast<uint32_t>(acc >> (nbits - 32));
            if (!prev_ff && has_ff_byte(w) == 0)
            {
                emit_word(w);
                nbits -= 32;
            }
            else
            {
                do
                {
                    emit_one_stuffed_byte();
                } while (nbits >= 32);
            }
        }
    }

    // limited-length Golomb code (T.87 A.5.3; reference src/scan_encoder_core.hpp:69-103)
    JLS_HD void put_golomb(int32_t k, int32_t mapped, int32_t limit, int32_t qbpp)
    {
        const int32_t high = mapped >> k;
        const int32_t escape = limit - qbpp - 1;
        if (high < escape && high + 1 + k <= 32)
        {
            put((1U << k) | (static_cast<uint32_t>(mapped) & ((1U << k) - 1U)), high + 1 + k);
            return;
        }
        // long code word (more than 32 bits) or escape code: unary part in at most two pieces, then the binary part
        int32_t zeros = high < escape ? high : escape;
        if (zeros > 31)
        {
            put(0, 31);
            zeros -= 31;
        }
        put(1, zeros + 1);
        if (high < escape)
            put(static_cast<uint32_t>(mapped) & ((1U << k) - 1U), k);
        else
            put(static_cast<uint32_t>(mapped - 1) & ((1U << qbpp) - 1U), qbpp);
    }

    // End of a restart interval / scan (reference src/scan_encoder.hpp:103-115): pad with zero bits to a byte; a final
    // 0xFF is followed by a zero byte.  Returns the number of bytes of the interval.
    JLS_HD uint32_t finish()
    {
        for (;;)
        {
            const int32_t take = prev_ff ? 7 : 8;
            if (nbits < take)
                break;
            emit_one_stuffed_byte();
        }
        if (nbits > 0)
        {
            const int32_t take = prev_ff ? 7 : 8;
            const uint32_t b = (static_cast<uint32_t>(acc) & ((1U << nbits) - 1U)) << (take - nbits);
            nbits = 0;
            emit_byte(b);
            prev_ff = false;
        }
        if (prev_ff)
        {
            emit_byte(0);
            prev_ff = false;
        }
        const uint32_t bytes = static_cast<uint32_t>(wp - base) * 4U + static_cast<uint32_t>(npend);
        if (npend != 0)
        {
            store_word(pend << (8 * (4 - npend)));
            npend = 0;
        }
        return bytes;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Bit reader over the byte range [begin, end) of one restart interval (the range never contains a marker).
// 64-bit left-aligned cache, refilled 4 bytes at a time from aligned 32-bit loads; bytes past `end` read as zero and
// are accounted for so that consuming them is reported (reference src/scan_decoder.hpp:250-333).
// ---------------------------------------------------------------------------------------------------------------------
struct BitReader
{
    uint64_t cache;
    int32_t valid;        // number of valid bits in cache (from the top)
    int32_t virtual_bits; // how many of the appended bits lie beyond `end`
    bool prev_ff;
    const uint32_t* wptr; // aligned word that holds the next unread byte
    uint32_t cur;         // *wptr
    uint32_t shift;       // 8 * (byte offset of the next unread byte inside *wptr)
    const uint8_t* pos;   // next unread byte
    const uint8_t* end;
    // Bits of the last code word if it was one the reference decodes through its 8-bit look-up table (a regular-mode Golomb
    // code of at most 8 bits, reference src/scan_decoder_core.hpp:46-52), else 255.  The reference peeks a whole byte for
    // those and never checks that all bits of the code it then skips were valid (src/scan_decoder.hpp:144-152, 58-64): the
    // last code word of an interval may end in up to seven zero bits that are not in the stream (interval_end_status).
    int32_t last_code_bits;
    bool marker_inside; // a 0xFF followed by a byte >= 0x80 was read as data

    JLS_HD bool leftover_is_zero_padding() // see FastReaderT
    {
        refill();
        return pos >= end && !residue();
    }

    JLS_HD uint32_t load_word(const uint32_t* w) const
    {
        return reinterpret_cast<const uint8_t*>(w) < end ? *w : 0U;
    }

    JLS_HD void init(const uint8_t* begin, const uint8_t* end_)
    {
        cache = 0;
        valid = 0;
        virtual_bits = 0;
        prev_ff = false;
        pos = begin;
        end = end_;
        last_code_bits = 255;
        marker_inside = false;
        const uintptr_t address = reinterpret_cast<uintptr_t>(begin);
        wptr = reinterpret_cast<const uint32_t*>(address & ~static_cast<uintptr_t>(3));
        shift = static_cast<uint32_t>(address & 3U) * 8U;
        cur = begin < end_ ? *wptr : 0U;
        refill();
    }

    JLS_HD void append_byte(uint32_t b, bool is_virtual)
    {
        if (prev_ff && (b & 0x80U) != 0)
            marker_inside = true; // see FastReaderT::refill_once
        const int32_t take = prev_ff ? 7 : 8;
        cache |= static_cast<uint64_t>(b & (0xFFU >> (8 - take))) << (64 - take - valid);
        valid += take;
        if (is_virtual)
            virtual_bits += take;
        prev_ff = (b == 0xFFU);
    }

    JLS_HD void refill() // brings valid above 32 (at most 64)
    {
        while (valid <= 32)
        {
            const uint32_t next = load_word(wptr + 1);
            const uint32_t w = bswap32(funnel_r(cur, next, shift)); // the next four stream bytes, first byte on top
            cur = next;
            ++wptr;
            const ptrdiff_t remaining = end - pos;
            pos += 4;
            if (remaining >= 4 && !prev_ff && has_ff_byte(w) == 0)
            {
                cache |= static_cast<uint64_t>(w) << (32 - valid);
                valid += 32;
                continue;
            }
            for (int32_t i = 0; i < 4; ++i)
            {
                const bool is_virtual = i >= remaining;
                append_byte(is_virtual ? 0U : (w >> (24 - 8 * i)) & 0xFFU, is_virtual);
            }
        }
    }

    JLS_HD bool residue() const { return cache != 0; } // bits that were not consumed are not all zero
    JLS_HD uint32_t peek(int32_t count) const { return static_cast<uint32_t>(cache >> (64 - count)); } // count 1..32

    JLS_HD void skip(int32_t count) // count 0..63
    {
        cache <<= count;
        valid -= count;
    }

    JLS_HD uint32_t read(int32_t count) // count 1..25, refills first
    {
        if (valid < 32)
            refill();
        const uint32_t v = peek(count);
        skip(count);
        return v;
    }

    // True when bits beyond the end of the interval were consumed (the reference throws invalid_data when it runs dry).
    JLS_HD bool overrun() const { return valid < virtual_bits; }
    JLS_HD int32_t overrun_bits() const { return virtual_bits - valid; }

    // Whole unread bytes left in the interval after the last decoded symbol.
    JLS_HD int64_t unread_bytes() const
    {
        const int64_t fetched_beyond = static_cast<int64_t>(pos - end); // may be negative
        const int64_t real_valid_bits = static_cast<int64_t>(valid) - virtual_bits;
        return (real_valid_bits > 0 ? real_valid_bits / 8 : 0) - (fetched_beyond > 0 ? 0 : fetched_beyond);
    }

    // limited-length Golomb code (reference src/scan_decoder.hpp:113-125,203-217); sets bad on a malformed code
    JLS_HD int32_t get_golomb(int32_t k, int32_t limit, int32_t qbpp, bool& bad)
    {
        int32_t zeros = 0;
        for (;;)
        {
            if (valid < 32)
                refill();
            const int32_t z = clz64(cache);
            if (z < valid)
            {
                zeros += z;
                cache = (cache << z) << 1;
                valid -= z + 1;
                break;
            }
            zeros += valid;
            cache = 0;
            valid = 0;
            if (overrun()) // ran off the end of the interval inside a unary code (the reference: invalid_data)
            {
                bad = true;
                return 0;
            }
        }
        if (zeros < limit - qbpp - 1)
        {
            last_code_bits = zeros + 1 + k <= 8 ? zeros + 1 + k : 255; // callers in run mode overwrite it with 255
            return k == 0 ? zeros : (zeros << k) + static_cast<int32_t>(read(k));
        }
        last_code_bits = 255;
        return static_cast<int32_t>(read(qbpp)) + 1;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Shared helpers for both codecs
// ---------------------------------------------------------------------------------------------------------------------

// reference src/scan_encoder.hpp:53-73
JLS_HD void encode_run_length(BitWriter& bw, int32_t& run_index, int32_t run_length, bool end_of_line)
{
    while (run_length >= (1 << run_order(run_index)))
    {
        bw.put(1, 1);
        run_length -= 1 << run_order(run_index);
        if (run_index < 31)
            ++run_index;
    }
    if (end_of_line)
    {
        if (run_length != 0)
            bw.put(1, 1);
    }
    else
    {
        bw.put(static_cast<uint32_t>(run_length), run_order(run_index) + 1);
    }
}

// reference src/scan_decoder_impl.hpp:305-337 (the fill loop is the caller's job); returns -1 on a run past the line end
JLS_HD int32_t decode_run_length(BitReader& br, int32_t& run_index, int32_t pixel_count)
{
    int32_t index = 0;
    while (br.read(1) != 0)
    {
        const int32_t block = 1 << run_order(run_index);
        const int32_t count = imin(block, pixel_count - index);
        index += count;
        if (count == block && run_index < 31)
            ++run_index;
        if (index == pixel_count)
            break;
    }
    if (index != pixel_count)
    {
        const int32_t j = run_order(run_index);
        if (j > 0)
            index += static_cast<int32_t>(br.read(j));
    }
    br.last_code_bits = 255; // run-length bits are read one by one, each with its own validity check (reference read_bit)
    return index > pixel_count ? -1 : index;
}

// reference src/scan_encoder_core.hpp:105-116
JLS_HD void encode_run_interruption_error(const CodecParams& p, BitWriter& bw, RunContext& c, int32_t ri_type, int32_t e,
                                          int32_t run_index)
{
    const int32_t k = run_golomb_parameter(c, ri_type);
    const int32_t map = run_compute_map(c, e, k);
    const int32_t e_mapped = 2 * iabs(e) - ri_type - map;
    bw.put_golomb(k, e_mapped, p.limit - run_order(run_index) - 1, p.qbpp);
    update_run_context(c, e, e_mapped, ri_type, p.reset);
}

// reference src/scan_decoder_core.hpp:72-80
JLS_HD int32_t decode_run_interruption_error(const CodecParams& p, BitReader& br, RunContext& c, int32_t ri_type,
                                             int32_t run_index, bool& bad)
{
    const int32_t k = run_golomb_parameter(c, ri_type);
    const int32_t e_mapped = br.get_golomb(k, p.limit - run_order(run_index) - 1, p.qbpp, bad);
    br.last_code_bits = 255; // no look-up table on this path (reference src/scan_decoder.hpp:113-125: every part checked)
    const int32_t e = run_error_value(c, e_mapped + ri_type, k);
    update_run_context(c, e, e_mapped, ri_type, p.reset);
    return e;
}

// reference src/scan_encoder_core.hpp:40-55 with an explicit sign; returns the reconstructed sample
template<bool LOSSLESS>
JLS_HD int32_t encode_regular_sample(const CodecParams& p, BitWriter& bw, RegularContext& c, int32_t sign, int32_t x,
                                     int32_t predicted, bool& bad)
{
    const int32_t k = golomb_parameter(c.a, c.n);
    if (k >= 16) // reference src/regular_mode_context.hpp:107-108,133-134
    {
        bad = true;
        return x;
    }
    const int32_t pv = correct_prediction(p, predicted + apply_sign(c.c, sign));
    const int32_t e = compute_error_value<LOSSLESS>(p, apply_sign(x - pv, sign));
    bw.put_golomb(k, map_error_value(error_correction(c, k | p.near) ^ e), p.limit, p.qbpp);
    if (!update_regular_context(c, e, p.dq, p.reset))
        bad = true;
    return LOSSLESS ? x : reconstruct<false>(p, pv, apply_sign(e, sign));
}

// reference src/scan_decoder_core.hpp:38-69
template<bool LOSSLESS>
JLS_HD int32_t decode_regular_sample(const CodecParams& p, BitReader& br, RegularContext& c, int32_t sign, int32_t predicted,
                                     bool& bad)
{
    const int32_t pv = correct_prediction(p, predicted + apply_sign(c.c, sign));
    const int32_t k = golomb_parameter(c.a, c.n);
    if (k >= 16)
    {
        bad = true;
        return 0;
    }
    int32_t e = unmap_error_value(br.get_golomb(k, p.limit, p.qbpp, bad));
    if (iabs(e) > 65535)
        bad = true;
    if (k == 0)
        e ^= error_correction(c, p.near);
    if (!update_regular_context(c, e, p.dq, p.reset))
        bad = true;
    return reconstruct<LOSSLESS>(p, pv, apply_sign(e, sign));
}

// ---------------------------------------------------------------------------------------------------------------------
// Colour transforms HP1..HP3 (reference src/color_transform.hpp:27-117).  `range` is 256 or 65536 (sample TYPE width).
// ---------------------------------------------------------------------------------------------------------------------
JLS_HD void color_forward(int32_t transform, int32_t type_mask, int32_t& c0, int32_t& c1, int32_t& c2)
{
    // transform is 1, 2 or 3 (the host validates it).  HP1 on its own path, HP2 / HP3 as selects: a chain of equality
    // tests on `transform` becomes a jump table (an indirect branch per pixel in the tile kernels).
    const int32_t range = type_mask + 1, bias = range / 2;
    const int32_t r = c0, g = c1, b = c2;
    const int32_t r_g = (r - g + bias) & type_mask;
    const int32_t b_g = (b - g + bias) & type_mask;
    if (transform == 1)
    {
        c0 = r_g;
        c1 = g & type_mask;
        c2 = b_g;
        return;
    }
    const bool hp3 = transform == 3;
    const int32_t hp2_c2 = (b - ((r + g) >> 1) + bias) & type_mask; // r, g >= 0: division == shift
    const int32_t hp3_c0 = (g + ((b_g + r_g) >> 2) - range / 4) & type_mask;
    c0 = hp3 ? hp3_c0 : r_g;
    c1 = hp3 ? b_g : (g & type_mask);
    c2 = hp3 ? r_g : hp2_c2;
}

JLS_HD void color_inverse(int32_t transform, int32_t type_mask, int32_t& c0, int32_t& c1, int32_t& c2)
{
    const int32_t range = type_mask + 1, bias = range / 2;
    const int32_t v1 = c0, v2 = c1, v3 = c2;
    const int32_t r12 = (v1 + v2 - bias) & type_mask; // R of HP1 and HP2
    if (transform == 1)
    {
        c0 = r12;
        c1 = v2 & type_mask;
        c2 = (v3 + v2 - bias) & type_mask;
        return;
    }
    const bool hp3 = transform == 3;
    const int32_t hp2_b = (v3 + ((r12 + (v2 & type_mask)) >> 1) - bias) & type_mask;
    const int32_t hp3_g = v1 - ((v3 + v2) >> 2) + range / 4;
    c0 = hp3 ? (v3 + hp3_g - bias) & type_mask : r12;
    c1 = (hp3 ? hp3_g : v2) & type_mask;
    c2 = hp3 ? (v2 + hp3_g - bias) & type_mask : hp2_b;
}

// ---------------------------------------------------------------------------------------------------------------------
// GENERAL PATH: any restart interval, full 2-D neighbourhood, 365 + 2 contexts in (thread-)local memory.
// Lines are kept as uint16_t[components][width + 2] x 2 in global scratch; index 0 and width + 1 are the edge samples
// (reference src/scan_codec.hpp:189-195).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int32_t general_context_count = 365;

struct GeneralState
{
    // The 365 regular contexts live where the caller puts them: the kernels hand every thread 365 * 16 bytes of SHARED memory.
    // As a member array they were thread-local memory, which interleaves the 32 lanes of a warp word by word -- with one
    // restart interval per scan only one lane of a warp works, every context access touched a cache line of its own and
    // the 187 KB a warp's contexts then span did not stay in L1 (31 -> ~1 MPix/s per stream, profiles/r2_notes.md).
    RegularContext* contexts;
    // Optional table of quantize_gradient(d) for d in [-MAXVAL, MAXVAL] at quant_lut[d + MAXVAL] (shared memory in the kernels
    // for samples of up to 12 bits; the reference's quantization_lut, src/scan_codec.hpp:22-99): one load instead of a chain
    // of up to eight compares, three times per sample.
    const int8_t* quant_lut;
    RunContext run_contexts[2];
    int32_t run_index;
    bool bad;

    JLS_HD void reset(const CodecParams& p) // reference src/scan_codec.hpp:163-174
    {
        const RegularContext initial = {p.a_init, 0, 0, 1};
        for (int32_t i = 0; i < general_context_count; ++i)
            contexts[i] = initial;
        for (int32_t i = 0; i < 2; ++i)
        {
            run_contexts[i].a = p.a_init;
            run_contexts[i].n = 1;
            run_contexts[i].nn = 0;
        }
        run_index = 0;
    }

    JLS_HD int32_t context_id(const CodecParams& p, int32_t ra, int32_t rb, int32_t rc, int32_t rd) const
    {
        // reference src/jpegls_algorithm.hpp:165-168
        if (quant_lut != nullptr)
        {
            const int8_t* q = quant_lut + p.maxval;
            return (q[rd - rb] * 9 + q[rb - rc]) * 9 + q[rc - ra];
        }
        return (quantize_gradient(p, rd - rb) * 9 + quantize_gradient(p, rb - rc)) * 9 + quantize_gradient(p, rc - ra);
    }
};

// Scalar line (ILV none and every component line of ILV line); reference src/scan_encoder_impl.hpp:109-144,249-275
template<bool LOSSLESS>
JLS_HD void general_encode_line(const CodecParams& p, GeneralState& s, BitWriter& bw, uint16_t* cur,
                                         const uint16_t* prev)
{
    // the neighbourhood slides along in registers, see general_decode_line
    const int32_t width = p.width;
    int32_t index = 1;
    int32_t ra = cur[0], rc = prev[0], rb = prev[1], rd = prev[2];
    while (index <= width)
    {
        const int32_t rd_next = prev[index + 2 <= width + 1 ? index + 2 : width + 1];
        const int32_t qs = s.context_id(p, ra, rb, rc, rd);
        if (qs != 0)
        {
            const int32_t sign = bit_wise_sign(qs);
            const int32_t x = encode_regular_sample<LOSSLESS>(p, bw, s.contexts[apply_sign(qs, sign)], sign, cur[index],
                                                              predict_med(ra, rb, rc), s.bad);
            if (!LOSSLESS)
                cur[index] = static_cast<uint16_t>(x); // the reconstructed value is the next line's neighbourhood
            ra = x;
            rc = rb;
            rb = rd;
            rd = rd_next;
            ++index;
            continue;
        }
        const int32_t remain = width - (index - 1);
        int32_t run_length = 0;
        while (iabs(static_cast<int32_t>(cur[index + run_length]) - ra) <= p.near)
        {
            cur[index + run_length] = static_cast<uint16_t>(ra);
            ++run_length;
            if (run_length == remain)
                break;
        }
        encode_run_length(bw, s.run_index, run_length, run_length == remain);
        if (run_length == remain)
            return;
        index += run_length;
        const int32_t x = cur[index];
        const int32_t rb2 = prev[index];
        int32_t reconstructed;
        if (iabs(ra - rb2) <= p.near) // reference src/scan_encoder_core.hpp:118-131
        {
            const int32_t e = compute_error_value<LOSSLESS>(p, x - ra);
            encode_run_interruption_error(p, bw, s.run_contexts[1], 1, e, s.run_index);
            reconstructed = reconstruct<LOSSLESS>(p, ra, e);
        }
        else
        {
            const int32_t sg = sign_of(rb2 - ra);
            const int32_t e = compute_error_value<LOSSLESS>(p, (x - rb2) * sg);
            encode_run_interruption_error(p, bw, s.run_contexts[0], 0, e, s.run_index);
            reconstructed = reconstruct<LOSSLESS>(p, rb2, e * sg);
        }
        cur[index] = static_cast<uint16_t>(reconstructed);
        if (s.run_index > 0)
            --s.run_index;
        ++index;
        ra = reconstructed;
        rc = rb2;
        rb = prev[index <= width + 1 ? index : width + 1];
        rd = prev[index + 1 <= width + 1 ? index + 1 : width + 1];
    }
}

// Sample-interleaved line; reference src/scan_encoder_impl.hpp:147-246,249-310. cur/prev: component c at + c * (width + 2).
template<bool LOSSLESS>
JLS_HD void general_encode_line_multi(const CodecParams& p, GeneralState& s, BitWriter& bw, uint16_t* cur,
                                               const uint16_t* prev)
{
    const int32_t width = p.width, nc = p.components, ps = width + 2;
    int32_t index = 1;
    while (index <= width)
    {
        int32_t qs[4];
        bool all_zero = true;
        for (int32_t c = 0; c < nc; ++c)
        {
            const uint16_t* pc = prev + c * ps;
            qs[c] = s.context_id(p, cur[c * ps + index - 1], pc[index], pc[index - 1], pc[index + 1]);
            all_zero = all_zero && qs[c] == 0;
        }
        if (!all_zero)
        {
            for (int32_t c = 0; c < nc; ++c)
            {
                const uint16_t* pc = prev + c * ps;
                uint16_t* cc = cur + c * ps;
                const int32_t sign = bit_wise_sign(qs[c]);
                cc[index] = static_cast<uint16_t>(
                    encode_regular_sample<LOSSLESS>(p, bw, s.contexts[apply_sign(qs[c], sign)], sign, cc[index],
                                                    predict_med(cc[index - 1], pc[index], pc[index - 1]), s.bad));
            }
            ++index;
            continue;
        }
        const int32_t remain = width - (index - 1);
        int32_t ra[4];
        for (int32_t c = 0; c < nc; ++c)
            ra[c] = cur[c * ps + index - 1];
        int32_t run_length = 0;
        for (;;)
        {
            bool near_all = true;
            for (int32_t c = 0; c < nc; ++c)
                near_all = near_all && iabs(static_cast<int32_t>(cur[c * ps + index + run_length]) - ra[c]) <= p.near;
            if (!near_all)
                break;
            for (int32_t c = 0; c < nc; ++c)
                cur[c * ps + index + run_length] = static_cast<uint16_t>(ra[c]);
            ++run_length;
            if (run_length == remain)
                break;
        }
        encode_run_length(bw, s.run_index, run_length, run_length == remain);
        if (run_length == remain)
            return;
        index += run_length;
        for (int32_t c = 0; c < nc; ++c)
        {
            const int32_t rb = prev[c * ps + index];
            const int32_t sg = sign_of(rb - ra[c]);
            const int32_t e = compute_error_value<LOSSLESS>(p, sg * (static_cast<int32_t>(cur[c * ps + index]) - rb));
            encode_run_interruption_error(p, bw, s.run_contexts[0], 0, e, s.run_index);
            cur[c * ps + index] = static_cast<uint16_t>(reconstruct<LOSSLESS>(p, rb, e * sg));
        }
        if (s.run_index > 0)
            --s.run_index;
        ++index;
    }
}

// reference src/scan_decoder_impl.hpp:132-159,264-281 + src/scan_decoder_core.hpp:83-92
template<bool LOSSLESS>
JLS_HD void general_decode_line(const CodecParams& p, GeneralState& s, BitReader& br, uint16_t* cur,
                                         const uint16_t* prev)
{
    // The neighbourhood slides along in registers: Ra is the sample just decoded, Rc and Rb were Rb and Rd one pixel ago,
    // and the only load per pixel -- Rd of the NEXT pixel -- is issued a whole pixel before it is needed.  (Reading Ra back
    // from the line that was stored to a moment ago put a trip to L2 on the critical path of every pixel.)
    const int32_t width = p.width;
    int32_t index = 1;
    int32_t ra = cur[0], rc = prev[0], rb = prev[1], rd = prev[2];
    while (index <= width && !s.bad)
    {
        const int32_t rd_next = prev[index + 2 <= width + 1 ? index + 2 : width + 1];
        const int32_t qs = s.context_id(p, ra, rb, rc, rd);
        if (qs != 0)
        {
            const int32_t sign = bit_wise_sign(qs);
            const int32_t x =
                decode_regular_sample<LOSSLESS>(p, br, s.contexts[apply_sign(qs, sign)], sign, predict_med(ra, rb, rc), s.bad);
            cur[index] = static_cast<uint16_t>(x);
            ra = x;
            rc = rb;
            rb = rd;
            rd = rd_next;
            ++index;
            continue;
        }
        const int32_t run_length = decode_run_length(br, s.run_index, width - (index - 1));
        if (run_length < 0)
        {
            s.bad = true;
            return;
        }
        for (int32_t i = 0; i < run_length; ++i)
            cur[index + i] = static_cast<uint16_t>(ra);
        index += run_length;
        if (index - 1 == width)
            return;
        const int32_t rb2 = prev[index];
        int32_t x;
        if (iabs(ra - rb2) <= p.near)
        {
            const int32_t e = decode_run_interruption_error(p, br, s.run_contexts[1], 1, s.run_index, s.bad);
            x = reconstruct<LOSSLESS>(p, ra, e);
        }
        else
        {
            const int32_t e = decode_run_interruption_error(p, br, s.run_contexts[0], 0, s.run_index, s.bad);
            x = reconstruct<LOSSLESS>(p, rb2, e * sign_of(rb2 - ra));
        }
        cur[index] = static_cast<uint16_t>(x);
        if (s.run_index > 0)
            --s.run_index;
        ++index;
        // the window jumped: reload it (index <= width + 1, so index + 1 may be the slot behind the right edge sample)
        ra = x;
        rc = rb2;
        rb = prev[index <= width + 1 ? index : width + 1];
        rd = prev[index + 1 <= width + 1 ? index + 1 : width + 1];
    }
}

// reference src/scan_decoder_impl.hpp:162-261,283-303 + src/scan_decoder_core.hpp:94-100
template<bool LOSSLESS>
JLS_HD void general_decode_line_multi(const CodecParams& p, GeneralState& s, BitReader& br, uint16_t* cur,
                                               const uint16_t* prev)
{
    const int32_t width = p.width, nc = p.components, ps = width + 2;
    int32_t index = 1;
    while (index <= width && !s.bad)
    {
        int32_t qs[4];
        bool all_zero = true;
        for (int32_t c = 0; c < nc; ++c)
        {
            const uint16_t* pc = prev + c * ps;
            qs[c] = s.context_id(p, cur[c * ps + index - 1], pc[index], pc[index - 1], pc[index + 1]);
            all_zero = all_zero && qs[c] == 0;
        }
        if (!all_zero)
        {
            for (int32_t c = 0; c < nc; ++c)
            {
                const uint16_t* pc = prev + c * ps;
                uint16_t* cc = cur + c * ps;
                const int32_t sign = bit_wise_sign(qs[c]);
                cc[index] = static_cast<uint16_t>(
                    decode_regular_sample<LOSSLESS>(p, br, s.contexts[apply_sign(qs[c], sign)], sign,
                                                    predict_med(cc[index - 1], pc[index], pc[index - 1]), s.bad));
            }
            ++index;
            continue;
        }
        int32_t ra[4];
        for (int32_t c = 0; c < nc; ++c)
            ra[c] = cur[c * ps + index - 1];
        const int32_t run_length = decode_run_length(br, s.run_index, width - (index - 1));
        if (run_length < 0)
        {
            s.bad = true;
            return;
        }
        for (int32_t i = 0; i < run_length; ++i)
            for (int32_t c = 0; c < nc; ++c)
                cur[c * ps + index + i] = static_cast<uint16_t>(ra[c]);
        index += run_length;
        if (index - 1 == width)
            return;
        for (int32_t c = 0; c < nc; ++c)
        {
            const int32_t rb = prev[c * ps + index];
            const int32_t e = decode_run_interruption_error(p, br, s.run_contexts[0], 0, s.run_index, s.bad);
            cur[c * ps + index] = static_cast<uint16_t>(reconstruct<LOSSLESS>(p, rb, e * sign_of(rb - ra[c])));
        }
        if (s.run_index > 0)
            --s.run_index;
        ++index;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Caller layout <-> samples (reference src/copy_to_line_buffer.hpp:26-261, src/copy_from_line_buffer.hpp:24-191)
// ---------------------------------------------------------------------------------------------------------------------
JLS_HD int32_t load_sample(const uint8_t* line, int32_t index, int32_t sample_bytes)
{
    return sample_bytes == 2 ? static_cast<int32_t>(reinterpret_cast<const uint16_t*>(line)[index])
                             : static_cast<int32_t>(line[index]);
}

JLS_HD void store_sample(uint8_t* line, int32_t index, int32_t sample_bytes, int32_t value)
{
    if (sample_bytes == 2)
        reinterpret_cast<uint16_t*>(line)[index] = static_cast<uint16_t>(value);
    else
        line[index] = static_cast<uint8_t>(value);
}

// Loads pixel `x` of a caller line into NC internal samples: masks unused high bits, or applies the forward colour
// transform (the transforming variants do not mask, copy_to_line_buffer.hpp:234-246).
template<int NC>
JLS_HD void load_pixel(const CodecParams& p, const uint8_t* line, int32_t x, int32_t (&v)[NC])
{
    const int32_t mask = (1 << p.bits_per_sample) - 1;
    for (int32_t c = 0; c < NC; ++c)
        v[c] = load_sample(line, x * NC + c, p.sample_bytes);
    if (NC == 3 && p.transform != 0)
    {
        color_forward(p.transform, p.sample_bytes == 2 ? 0xFFFF : 0xFF, v[0], v[NC > 1 ? 1 : 0], v[NC > 2 ? 2 : 0]);
    }
    else
    {
        for (int32_t c = 0; c < NC; ++c)
            v[c] &= mask;
    }
}

template<int NC>
JLS_HD void store_pixel(const CodecParams& p, uint8_t* line, int32_t x, const int32_t (&v)[NC])
{
    int32_t o[NC];
    for (int32_t c = 0; c < NC; ++c)
        o[c] = v[c];
    if (NC == 3 && p.transform != 0)
        color_inverse(p.transform, p.sample_bytes == 2 ? 0xFFFF : 0xFF, o[0], o[NC > 1 ? 1 : 0], o[NC > 2 ? 2 : 0]);
    for (int32_t c = 0; c < NC; ++c)
        store_sample(line, x * NC + c, p.sample_bytes, o[c]);
}

} // namespace jls
