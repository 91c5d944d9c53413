// jls_fast.cuh -- the restart-interval-1 line codec, written for instruction count.
//
// With one restart interval per line the previous line is all zeros at every line start (reference
// src/scan_decoder_impl.hpp:122-127), so Rb = Rc = Rd = 0: the gradient vector is (0, 0, -Ra), the MED prediction is Ra,
// and only contexts 0..4 (index |Q(-Ra)|) exist.  One thread codes one line, a warp 32 lines in
// lock step, one pixel per loop iteration.  The kernels are bound by instruction issue (ncu: profiles/), so this file
// is about executing few instructions per pixel AND about keeping the rare paths rare for the whole warp:
//   * coding parameters live in registers (HotParams), not in the constant bank;
//   * the bit writer / reader move 32 bits at a time on a straight-line path; byte-wise code runs only around 0xFF
//     bytes and at the ends of a line;
//   * the context in use is cached in registers; the five contexts of a thread sit in shared memory as
//     [context][thread] (conflict-free 16-byte accesses) and are touched only when the context changes;
//   * conditions that cannot occur for valid input are not checked in the encoder (the decoder checks everything the
//     reference checks, it sees untrusted data).
// Everything is __host__ __device__: tests/hostemu runs this very code on the CPU against the oracle.
#pragma once

#include "jls_codec.cuh"

#if defined(__GNUC__) || defined(__CUDACC__)
#define JLS_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define JLS_LIKELY(x) __builtin_expect(!!(x), 1)
#else
#define JLS_UNLIKELY(x) (x)
#define JLS_LIKELY(x) (x)
#endif

// A/B switches (tools/ab_build.py name:-DJLS_...=value builds a variant of the library; results in profiles/r1_notes.md):
// how lossless 16-bit samples are written: write_immediate (0) or write_wide (2), see FastWriter::put
#ifndef JLS_WRITER_MODE_16
#define JLS_WRITER_MODE_16 2
#endif
// stream words the three-component decoder keeps loaded ahead of their use: 1 or 2, see FastReaderT::refill_pair
#ifndef JLS_READER_DEPTH_NC3
#define JLS_READER_DEPTH_NC3 2
#endif
// round-2 switches (results in profiles/r2_notes.md):
// 1: the context in use is cached in registers and swapped with shared memory when the index changes (round 1);
// 0: every sample loads its context from shared memory and stores it back (one LDS.128 + one STS.128 instead of a compare
//    and seven predicated instructions that issue for every sample whether the context changes or not)
#ifndef JLS_CONTEXT_CACHE
#define JLS_CONTEXT_CACHE 0
#endif
// 1: the bias step of the context update works on -B with one two-sided clamp (8 instructions instead of 12)
#ifndef JLS_BIAS_NEGATED
#define JLS_BIAS_NEGATED 1
#endif
// 1: additions of the straight path that ptxas puts on the (busy) integer ALU pipe are written as multiply-adds with an
//    opaque factor 1 so that they run on the FMA pipe
#ifndef JLS_FMA_ADDS
#define JLS_FMA_ADDS 1
#endif
// 1: the one-component 8-bit decoders top their read window up every 8 samples instead of every 4 (FastLineDecoder).
// Measured (profiles/r2_notes.md): no gain -- what a top-up costs the warp is its number of refill iterations, and eight
// samples need two where four need one; cfg2 decoder 6.88 ms (4) against 6.95 ms (8).  Off.
#ifndef JLS_TOP_UP_8
#define JLS_TOP_UP_8 0
#endif

namespace jls {

// The coding parameters a line needs, held in registers.
struct HotParams
{
    int32_t t1, t2, t3, near, maxval, limit, qbpp, reset, bits, escape, dq, range, range_dq, a_init, transform;
    int32_t one; // the value 1, opaque to the compiler on the device (see FastLineState::select_context, add_fma)
    int32_t two; // the value 2, likewise (golomb_parameter_reciprocal)
    uint32_t dq_magic;
    uint32_t sign_scale; // 2^(32 - bits)
    // Optional table |Q(-Ra)| for Ra in [0, context_lut_last]; larger Ra use the last entry (they are >= T3).  The tile
    // kernels keep it in shared memory: one LDS replaces three compares, a select and the adds on the busy ALU pipe.
    const uint8_t* context_lut;
    int32_t context_lut_last;
    uint32_t context_lut_shared; // device: shared-window address of context_lut (see keep_hot_params_in_registers)
    // Optional table ceil(2^31 / N) for N in [0, RESET] (entry 0 unused): with it the Golomb parameter is two multiplies on
    // the FMA pipe and one find-leading-one instead of two find-leading-one, a clamp, a shift, a compare and a select
    // (golomb_parameter_reciprocal).
    const uint32_t* reciprocal_lut;
    uint32_t reciprocal_lut_shared;
};

constexpr int32_t context_lut_capacity = 1024; // the table covers T3 <= 1023 (defaults: 21 / 85 / 276 for 8 / 12 / 16 bit)
// The table form of the Golomb parameter is exact for N <= 64 and A < 2^24 (checked exhaustively, tests/test_hostemu.py);
// N runs up to RESET (default 64 for every bit depth), larger RESET values take the kernels without tables.
constexpr int32_t reciprocal_lut_capacity = 65;

JLS_HD uint32_t reciprocal_lut_entry(int32_t n)
{
    if (n <= 1)
        return 0x7FFFFFFFU; // n = 1: floor((2a - 1) * (2^31 - 1) / 2^32) = a - 1 exactly, and the factor stays positive
    return static_cast<uint32_t>(((1ULL << 31) + static_cast<uint64_t>(n) - 1U) / static_cast<uint64_t>(n));
}

// min k with (n << k) >= a, from reciprocal = ceil(2^31 / n):  k = bit length of floor((a - 1) / n), and
// t = floor((2a - 1) * reciprocal / 2^32) has that bit length for n <= 64, a < 2^24 (the quotient is overestimated by
// less than what it takes to reach the next power of two).  a = 0 gives t = -1, whose most significant non-sign bit
// does not exist: k = 0 like the reference.  Any a gives 0 <= k <= 31.
// `two` = 2, as a register the compiler cannot see through on the device: 2a - 1 becomes one IMAD (FMA pipe) instead of an
// IADD3 on the integer ALU pipe, which is the busy one (JLS_FMA_ADDS)
JLS_HD int32_t golomb_parameter_reciprocal(int32_t a, uint32_t reciprocal, int32_t two = 2, int32_t one = 1)
{
#if defined(__CUDA_ARCH__)
    int32_t position;
#if JLS_FMA_ADDS
    int32_t doubled;
    asm("mad.lo.s32 %0, %1, %2, -1;" : "=r"(doubled) : "r"(a), "r"(two));
#else
    const int32_t doubled = 2 * a - 1;
#endif
    asm("bfind.s32 %0, %1;" : "=r"(position) : "r"(__mulhi(doubled, static_cast<int32_t>(reciprocal))));
#if JLS_FMA_ADDS
    int32_t k;
    asm("mad.lo.s32 %0, %1, %2, 1;" : "=r"(k) : "r"(position), "r"(one)); // bfind yields -1 for 0 and for -1
    return k;
#else
    return position + 1;
#endif
#else
    (void)two;
    (void)one;
    const int64_t t = (static_cast<int64_t>(static_cast<int32_t>(2U * static_cast<uint32_t>(a) - 1U)) * static_cast<int64_t>(reciprocal)) >> 32;
    return t < 0 ? 32 - clz32(static_cast<uint32_t>(~t)) : 32 - clz32(static_cast<uint32_t>(t));
#endif
}

JLS_HD HotParams make_hot_params(const CodecParams& p)
{
    HotParams h;
    h.t1 = p.t1;
    h.t2 = p.t2;
    h.t3 = p.t3;
    h.near = p.near;
    h.maxval = p.maxval;
    h.limit = p.limit;
    h.qbpp = p.qbpp;
    h.reset = p.reset;
    h.bits = p.bits_per_sample;
    h.escape = p.limit - p.qbpp - 1;
    h.dq = p.dq;
    h.range = p.range;
    h.range_dq = p.range_dq;
    h.a_init = p.a_init;
    h.transform = p.transform;
    h.one = 1;
    h.two = 2;
    h.dq_magic = p.dq_magic;
    h.sign_scale = 1U << (32 - p.bits_per_sample);
    h.context_lut = nullptr;
    h.context_lut_last = 0;
    h.context_lut_shared = 0;
    h.reciprocal_lut = nullptr;
    h.reciprocal_lut_shared = 0;
    return h;
}

#if defined(__CUDACC__)
// Moves the parameters the pixel loop reads into registers.  Kernel parameters live in the constant bank and ptxas
// re-reads them (LDC / LDCU, one issue slot each) at every use instead of spending a register: 5 to 9 slots per pixel.
// It also sees through shuffles and register moves of such warp-uniform values, so they take a round trip through
// `scratch` (hot_scratch_words words of shared memory owned by the calling warp; all 32 lanes call this).
__device__ __forceinline__ void keep_hot_params_in_registers(HotParams& h, volatile int32_t* scratch)
{
    if ((threadIdx.x & 31U) == 0)
    {
        scratch[0] = h.t1;
        scratch[1] = h.t2;
        scratch[2] = h.t3;
        scratch[3] = h.near;
        scratch[4] = h.reset;
        scratch[5] = h.escape;
        scratch[6] = h.maxval;
        scratch[7] = h.bits;
        scratch[8] = static_cast<int32_t>(h.context_lut_shared);
        scratch[9] = static_cast<int32_t>(h.reciprocal_lut_shared);
        scratch[10] = static_cast<int32_t>(h.sign_scale);
        scratch[11] = h.context_lut_last;
        scratch[12] = h.transform;
        scratch[13] = 1;
        scratch[14] = 2;
    }
    __syncwarp();
    h.t1 = scratch[0];
    h.t2 = scratch[1];
    h.t3 = scratch[2];
    h.near = scratch[3];
    h.reset = scratch[4];
    h.escape = scratch[5];
    h.maxval = scratch[6];
    h.bits = scratch[7];
    h.context_lut_shared = static_cast<uint32_t>(scratch[8]); // same story for an address that derives from the CTA id
    h.reciprocal_lut_shared = static_cast<uint32_t>(scratch[9]);
    h.sign_scale = static_cast<uint32_t>(scratch[10]);
    h.context_lut_last = scratch[11];
    h.transform = scratch[12];
    h.one = scratch[13];
    h.two = scratch[14];
}
constexpr int hot_scratch_words = 15;
#endif

// DEPTH != 0: the sample depth is known at compile time (the tile kernels instantiate 8 and 16 for lossless data that
// fills its container); 0: it is h.bits.
template<bool LOSSLESS, int DEPTH = 0>
JLS_HD int32_t fast_error_value(const HotParams& h, int32_t e)
{
    if (LOSSLESS)
    {
        // modulo RANGE = sign extension from bit `bits` (reference src/lossless_traits.hpp:61-65)
#if defined(__CUDA_ARCH__)
        if (DEPTH != 0)
        {
            int32_t r;
            asm("bfe.s32 %0, %1, 0, %2;" : "=r"(r) : "r"(e), "n"(DEPTH)); // one SGXT
            return r;
        }
#endif
        // the left shift is a multiplication so that it goes to the FMA pipe, the integer ALU pipe is the busy one
        return static_cast<int32_t>(static_cast<uint32_t>(e) * h.sign_scale) >> (32 - h.bits);
    }
    int32_t q = static_cast<int32_t>(mulhi32(static_cast<uint32_t>(iabs(e) + h.near), h.dq_magic));
    q = e > 0 ? q : -q;
    if (q < 0)
        q += h.range;
    if (q >= (h.range + 1) / 2)
        q -= h.range;
    return q;
}

JLS_HD int32_t fast_clamp(const HotParams& h, int32_t v) // == correct_prediction: v in [0, maxval] or the nearer bound
{
    return imin(imax(v, 0), h.maxval);
}

// three-input maximum: one VIMNMX3 on sm_100a
JLS_HD uint32_t umax3(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __vimax3_u32(a, b, c);
#else
    return umax(umax(a, b), c);
#endif
}

// max(min(a + b, high), 0) -- one VIADDMNMX.RELU on sm_100a
JLS_HD int32_t add_clamp_relu(int32_t a, int32_t b, int32_t high)
{
#if defined(__CUDA_ARCH__)
    return __viaddmin_s32_relu(a, b, high);
#else
    return imax(imin(a + b, high), 0);
#endif
}

// a + b on the FMA pipe: an IMAD with the factor h.one, which the compiler cannot see through.  Inline PTX so that NVVM does
// not reassociate (x * one + 1 and x * one + 2 became one multiply and two additions on the ALU pipe); plain addition on
// the host.
JLS_HD int32_t add_fma(const HotParams& h, int32_t a, int32_t b)
{
#if defined(__CUDA_ARCH__) && JLS_FMA_ADDS
    int32_t r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(h.one), "r"(b));
    return r;
#else
    (void)h;
    return a + b;
#endif
}

// unmap_error_value with the sign mask from one SGXT (bit-field extract of bit 0, sign extended)
JLS_HD int32_t fast_unmap(int32_t m)
{
#if defined(__CUDA_ARCH__)
    int32_t sign;
    asm("bfe.s32 %0, %1, 0, 1;" : "=r"(sign) : "r"(m));
    return sign ^ (m >> 1);
#else
    return unmap_error_value(m);
#endif
}

template<bool LOSSLESS>
JLS_HD int32_t fast_reconstruct(const HotParams& h, int32_t predicted, int32_t error_value)
{
    if (LOSSLESS)
        return (predicted + error_value) & h.maxval;
    int32_t v = predicted + error_value * h.dq;
    if (v < -h.near)
        v += h.range_dq;
    else if (v > h.maxval + h.near)
        v -= h.range_dq;
    return fast_clamp(h, v);
}

// T.87 A.12 / A.13 (reference src/regular_mode_context.hpp:45-94); branch-light.
// The reference's sanity check (:52-54: A >= 2^24 or |B| >= 2^24 before the halving and clamping) is kept as a high-water
// mark that only the decoder looks at: the return value is the maximum of everything that must stay below
// sanity_limit = 2^24 (the decoder adds k << 20 and |e| << 8, whose bounds land on the same limit).  Three-input
// maxima replace a shift per quantity and the ORs.  With NEAR = 0, |B| stays below RESET + 65535 and is not looked at.
constexpr uint32_t sanity_limit = 1U << 24;

template<bool LOSSLESS>
JLS_HD uint32_t fast_update_context(const HotParams& h, RegularContext& c, int32_t e)
{
    // unsigned: on damaged input A keeps growing after the high-water mark has tripped (the line is decoded to its end) and
    // may pass 2^31; the wrapped value still reads as >= sanity_limit
    c.a = static_cast<int32_t>(static_cast<uint32_t>(c.a) + static_cast<uint32_t>(iabs(e)));
#if JLS_BIAS_NEGATED
    // The fast path keeps nb = -B in RegularContext::b.  The reference's two branches (B + N <= 0: B += N, at least 1 - N,
    // C--; B > 0: B -= N, at most 0, C++) are one step d in {-1, 0, +1} and ONE two-sided clamp, because a B that needs no
    // correction already lies in [1 - N, 0]:  nb' = clamp(nb + d N, 0, N - 1), C' = clamp(C + d, -128, 127) with
    // d = +1 for nb < 0 (B > 0) and -1 for nb >= N (B + N <= 0).  N - 1 is the count before the increment, and the clamp
    // is a single VIADDMNMX.RELU.  Halving B (arithmetic shift = floor) is ceil(nb / 2) on the negated value.
    int32_t nb = c.b - (LOSSLESS ? e : e * h.dq);
    const uint32_t water = LOSSLESS ? 0U : umax(static_cast<uint32_t>(c.a), static_cast<uint32_t>(iabs(nb)));
    int32_t n = c.n;
    if (JLS_UNLIKELY(n == h.reset))
    {
        c.a >>= 1;
        nb = (nb + 1) >> 1;
        n >>= 1;
    }
    const int32_t n1 = add_fma(h, n, 1);
    const int32_t d = nb < 0 ? 1 : (nb > n ? -1 : 0);
    c.b = add_clamp_relu(nb, d * n1, n);
    c.c = imax(imin(c.c + d, 127), -128);
    c.n = n1;
    return water;
#else
    c.b += LOSSLESS ? e : e * h.dq;
    const uint32_t water = LOSSLESS ? static_cast<uint32_t>(c.a) : umax(static_cast<uint32_t>(c.a), static_cast<uint32_t>(iabs(c.b)));
    if (JLS_UNLIKELY(c.n == h.reset))
    {
        c.a >>= 1;
        c.b >>= 1;
        c.n >>= 1;
    }
    ++c.n;
    const int32_t t = c.b + c.n;
    const bool low = t <= 0;
    const bool high = c.b > 0;
    const int32_t b_low = imax(t, 1 - c.n);
    const int32_t b_high = imin(c.b - c.n, 0);
    const int32_t c_low = imax(c.c - 1, -128);
    const int32_t c_high = imin(c.c + 1, 127);
    c.b = low ? b_low : (high ? b_high : c.b);
    c.c = low ? c_low : (high ? c_high : c.c);
    return water;
#endif
}

// the reference's error correction applies (before k and NEAR are looked at): 2 B + N < 1 (src/regular_mode_context.hpp:36-42)
JLS_HD bool fast_correction_sign(const RegularContext& c)
{
#if JLS_BIAS_NEGATED
    return 2 * c.b >= c.n;
#else
    return 2 * c.b + c.n < 1;
#endif
}

// Fills entry `index` of the context table (callers loop / stride over [0, last]).
JLS_HD uint8_t context_lut_entry(const CodecParams& p, int32_t ra_value)
{
    return static_cast<uint8_t>((ra_value >= p.t3) + (ra_value >= p.t2) + (ra_value >= p.t1) + (ra_value > p.near));
}

// ---------------------------------------------------------------------------------------------------------------------
// Bit writer for a slot that is large enough by construction (worst_case_interval_bytes): no capacity checks.
// ---------------------------------------------------------------------------------------------------------------------
enum : int
{
    write_immediate = 0,
    write_deferred = 1,
    write_wide = 2,
    // Values from 16 up: "steady" writing with the value as the longest code word of the unchecked path, see
    // FastWriter::put_golomb.  96-bit accumulator, drained to < 32 bits by all lanes together every
    // steady_pixels_per_drain pixels; in between, code words of up to MODE bits are appended WITHOUT a capacity test
    // (31 + NC * pixels * MODE <= 95), anything longer and every run-mode symbol goes through the checked put and ends
    // with a drain of its own (settle).  Round 1 tested the capacity at every code word: a compare, a branch and a
    // reconvergence pair per sample on the straight path.
    write_steady_16 = 16,
    write_steady_21 = 21,
    write_steady_32 = 32
};

template<int MODE>
constexpr bool is_steady = MODE >= 16;
template<int MODE, int NC>
constexpr int steady_pixels_per_drain = 64 / (NC * MODE) >= 4 ? 4 : 64 / (NC * MODE) >= 2 ? 2 : 1;

#ifndef JLS_STEADY_WRITER
#define JLS_STEADY_WRITER 1
#endif
// 1 (A/B, off): write_steady_21 / write_steady_32 drain one or two words in one step (FastWriter::drain_one_or_two):
// 94 -> 92 instructions per 32 samples of cfg4, encode 6.78 -> 6.87 ms
// 1 (A/B, off): the three-component decoder's top-up takes one or two words in one straight-line step
// (FastReaderT::refill_words) instead of a two-word step for the lanes that need it followed by the one-word loop
#ifndef JLS_TOP_UP_UNIFIED
#define JLS_TOP_UP_UNIFIED 0
#endif
#ifndef JLS_DRAIN_FUSED
#define JLS_DRAIN_FUSED 0
#endif
// Lossless 16-bit samples need room for ~12-bit code words, everything else (8-bit containers, near-lossless) codes a few
// bits per sample and drains less often with the short limit.
template<int NC, bool LOSSLESS, typename S>
constexpr int writer_mode = !JLS_STEADY_WRITER ? (LOSSLESS && sizeof(S) == 2 ? JLS_WRITER_MODE_16 : write_deferred)
                            : NC == 3                     ? write_steady_21
                            : NC == 4                     ? write_steady_16
                            : LOSSLESS && sizeof(S) == 2 ? write_steady_32
                                                          : write_steady_16;

// L2 eviction hints (profiles/r2g_ab_l2_hints.txt).  Bit 0 (on): the encoder's slot words are stored `evict_last`: encode
// 6.01 -> 5.84 ms (cfg2), 6.88 -> 6.78 ms (cfg4).  A 32-byte sector of a slot is written by eight separate word stores of one
// lane, and ncu shows 0.45 GB more DRAM writes AND reads than the payload for cfg2 (sectors written back half full and
// completed later by a read-modify-write); the hint does not change that traffic, so it is not a capacity effect -- it only
// shortens the stores.  Bit 1 (off): sample tiles loaded `evict_first` -- DRAM reads 2.6 -> 7.5 GB: a tile row is 32 bytes and
// the next three tiles of the row live off the 128 bytes L2 fetched for the first.
#ifndef JLS_L2_HINTS
#define JLS_L2_HINTS 1
#endif

JLS_HD void store_slot_word(uint32_t* p, uint32_t value)
{
#if defined(__CUDA_ARCH__) && (JLS_L2_HINTS & 1)
    uint64_t policy;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(value), "l"(policy) : "memory");
#else
    *p = value;
#endif
}

// bit 2 of JLS_L2_HINTS (A/B): the decoder's stream words (one 4-byte load per lane and refill, eight loads per sector)
JLS_HD uint32_t load_stream_word(const uint32_t* p)
{
#if defined(__CUDA_ARCH__) && (JLS_L2_HINTS & 4)
    uint64_t policy;
    uint32_t value;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(value) : "l"(p), "l"(policy) : "memory");
    return value;
#else
    return *p;
#endif
}

// `fallback`, or *p when `wanted`: a predicated load into the register that holds the fallback
JLS_HD uint32_t load_stream_word_if(bool wanted, const uint32_t* p, uint32_t fallback)
{
#if defined(__CUDA_ARCH__)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.u32 %0, [%1];\n\t}"
                 : "+r"(fallback)
                 : "l"(p), "r"(static_cast<uint32_t>(wanted))
                 : "memory");
    return fallback;
#else
    return wanted ? *p : fallback;
#endif
}

struct FastWriter
{
    uint64_t acc;        // pending bits in the low `nbits` bits
    uint32_t acc_hi;     // write_wide only: bits 64..95 of the accumulator
    int32_t nbits;       // < 32 between calls (write_immediate), <= 64 (write_deferred), < 96 (write_wide)
    uint32_t pend;       // pending output bytes in the low pend_shift / 8 bytes
    uint32_t pend_shift; // 8 * (number of pending bytes), 0..24
    uint32_t prev_ff;    // last emitted byte was 0xFF
    uint32_t* wp;
    uint32_t* base;

    JLS_HD void init(uint8_t* destination)
    {
        acc = 0;
        acc_hi = 0;
        nbits = 0;
        pend = 0;
        pend_shift = 0;
        prev_ff = 0;
        base = reinterpret_cast<uint32_t*>(destination);
        wp = base;
    }

    JLS_HD void emit_byte(uint32_t b)
    {
        pend = (pend << 8) | b;
        pend_shift += 8;
        if (pend_shift == 32)
        {
            store_slot_word(wp++, bswap32(pend));
            pend_shift = 0;
        }
    }

    JLS_HD void emit_one_stuffed_byte()
    {
        const int32_t take = prev_ff ? 7 : 8;
        const uint32_t b = static_cast<uint32_t>(acc >> (nbits - take)) & (0xFFU >> (8 - take));
        nbits -= take;
        emit_byte(b);
        prev_ff = (b == 0xFFU) ? 1U : 0U;
    }

    // The 32 bits on top of the accumulator, nbits >= 32 (WIDE: nbits < 96 and the accumulator is acc_hi : acc).
    template<bool WIDE>
    JLS_HD uint32_t top_word() const
    {
        if (WIDE && nbits > 64)
            return static_cast<uint32_t>(acc >> (nbits - 32)) | (acc_hi << (96 - nbits));
        return static_cast<uint32_t>(acc >> (nbits - 32));
    }

    // Moves one 32-bit word (four stuffed bytes when a 0xFF is around) from the accumulator to the slot; nbits >= 32.
    template<bool WIDE = false>
    JLS_HD void flush_word()
    {
        const uint32_t w = top_word<WIDE>();
        if (JLS_LIKELY((prev_ff | has_ff_byte(w)) == 0))
        {
            store_slot_word(wp++, bswap32(funnel_r(w, pend, pend_shift)));
            pend = w;
            nbits -= 32;
        }
        else
        {
            flush_word_stuffed<WIDE>(w);
        }
    }

    // The same word when a 0xFF byte is around: four output bytes again, the byte after a 0xFF takes seven bits, so 28 to
    // 32 bits leave the accumulator.  Straight-line (five instructions per byte): a warp takes this path whenever ONE of
    // its lanes meets a 0xFF, at a fifth of all drains for 8-bit data.  Ends with nbits < 32 like the plain path unless
    // WIDE (whose callers loop).
    template<bool WIDE>
    JLS_HD void flush_word_stuffed(uint32_t w)
    {
        for (;;)
        {
            uint32_t out = 0;
            uint32_t ff = prev_ff;
            int32_t left = 32; // bits of w not yet taken
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int32_t i = 0; i < 4; ++i)
            {
                left -= ff ? 7 : 8;
                const uint32_t b = (w >> left) & (ff ? 0x7FU : 0xFFU);
                out = (out << 8) | b;
                ff = (b == 0xFFU) ? 1U : 0U;
            }
            prev_ff = ff;
            store_slot_word(wp++, bswap32(funnel_r(out, pend, pend_shift)));
            pend = out;
            nbits -= 32 - left;
            if (WIDE || JLS_LIKELY(nbits < 32))
                return;
            w = top_word<false>(); // up to four stuffed bits stayed behind
            if ((prev_ff | has_ff_byte(w)) == 0)
            {
                store_slot_word(wp++, bswap32(funnel_r(w, pend, pend_shift)));
                pend = w;
                nbits -= 32;
                return;
            }
        }
    }

    // value < 2^count, count in [0, 32].  The accumulator holds up to 64 bits; it is drained at a fixed cadence by
    // drain() (all lanes of a warp at the same pixel), so the branch below is taken only after unusually long codes.
    // A flush that any one lane needs costs the whole warp its ~15 instructions: at ~4 bits per pixel some lane needed
    // one at almost every pixel (profiles/r1_notes.md), a drain every 4 pixels costs a quarter of that.
    // write_immediate flushes as soon as 32 bits are pending; write_deferred leaves it to drain() (the accumulator holds 64
    // bits: right for a few bits per sample); write_wide does the same with a 96-bit accumulator (lossless 16-bit data at
    // 12 bits per sample: three samples of a pixel fit between two drains, where write_immediate has a word ready in
    // 13 of 32 lanes at every sample and walks the whole warp through the flush each time).
    template<int MODE>
    JLS_HD void put(uint32_t value, int32_t count)
    {
        if (MODE == write_wide || is_steady<MODE>)
        {
            if (JLS_UNLIKELY(nbits + count > 95)) // top_word() needs nbits < 96
                drain<write_wide>();
            acc_hi = funnel_l(static_cast<uint32_t>(acc >> 32), acc_hi, static_cast<uint32_t>(count));
            acc = (acc << count) | value;
            nbits += count;
        }
        else if (MODE == write_deferred)
        {
            if (JLS_UNLIKELY(nbits + count > 64))
            {
                do
                {
                    flush_word();
                } while (nbits >= 32);
            }
            acc = (acc << count) | value;
            nbits += count;
        }
        else
        {
            acc = (acc << count) | value;
            nbits += count;
            if (nbits >= 32)
                flush_word();
        }
    }

    // steady modes: the same without the capacity test (the caller's contract guarantees nbits + count <= 95)
    JLS_HD void put_unchecked(uint32_t value, int32_t count)
    {
        acc_hi = funnel_l(static_cast<uint32_t>(acc >> 32), acc_hi, static_cast<uint32_t>(count));
        acc = (acc << count) | value;
        nbits += count;
    }

    // brings nbits below 32
    template<int MODE = write_deferred>
    JLS_HD void drain()
    {
        if (JLS_DRAIN_FUSED && (MODE == write_steady_21 || MODE == write_steady_32))
        {
            drain_one_or_two();
            return;
        }
        while (nbits >= 32)
            flush_word<MODE == write_wide || is_steady<MODE>>();
    }

    // The steady modes that see more than one word per drain and lane on average (three samples per pixel, lossless 16-bit
    // samples): ONE word (nbits < 64) or TWO in the same straight-line step, nbits < 96.  As a loop over flush_word() the
    // warp makes two trips at nearly every pixel -- 18 lanes active in the first, a few in the second (ncu, cfg4: 21 of the
    // encoder's 94 instructions per 32 samples).
    JLS_HD void drain_one_or_two()
    {
        if (nbits < 32)
            return;
        const bool two = nbits >= 64;
        const uint32_t a0 = static_cast<uint32_t>(acc), a1 = static_cast<uint32_t>(acc >> 32);
        const uint32_t t = static_cast<uint32_t>(nbits) & 31U; // nbits - 32 or nbits - 64
        const uint32_t w1 = funnel_r(two ? a1 : a0, two ? acc_hi : a1, t);
        const uint32_t w2 = two ? funnel_r(a0, a1, t) : 0U;
        if (JLS_LIKELY((prev_ff | has_ff_byte(w1) | has_ff_byte(w2)) == 0))
        {
            store_slot_word(wp, bswap32(funnel_r(w1, pend, pend_shift)));
            if (two)
                store_slot_word(wp + 1, bswap32(funnel_r(w2, w1, pend_shift)));
            pend = two ? w2 : w1;
            wp += two ? 2 : 1;
            nbits -= two ? 64 : 32;
        }
        else
        {
            while (nbits >= 32)
                flush_word<true>();
        }
    }

    // steady modes: called after symbols that went through the checked put, restores nbits < 32 for the unchecked ones
    template<int MODE>
    JLS_HD void settle()
    {
        if (is_steady<MODE>)
            drain<MODE>();
    }

    // limited-length Golomb code (T.87 A.5.3; reference src/scan_encoder_core.hpp:69-103)
    template<int MODE>
    JLS_HD void put_golomb(const HotParams& h, int32_t k, int32_t mapped, int32_t escape)
    {
        const int32_t high = mapped >> k;
        if (is_steady<MODE>)
        {
            // The short code word is built unconditionally and appended behind the rare branch, which zeroes it after
            // writing the long form: the straight path has no jump over the long form's code.
            // mapped = high << k | low and the code word is 1 << k | low: flip the bits in which high differs from 1
            uint32_t value = static_cast<uint32_t>(mapped) ^ (static_cast<uint32_t>(high ^ 1) << k);
            int32_t length = add_fma(h, high, add_fma(h, k, 1));
            // an escape (high >= escape), or longer than MODE bits: one compare -- a code word of at most min(MODE, escape) bits
            // has fewer than `escape` zeros in front (the few legal code words between the two bounds take the long form's
            // path as well; it writes them just the same), and the bound is loop invariant
            if (JLS_UNLIKELY(length > imin(MODE, escape)))
            {
                put_golomb<write_wide>(h, k, mapped, escape);
                drain<MODE>();
                value = 0;
                length = 0;
            }
            put_unchecked(value, length);
            return;
        }
        const int32_t length = high + 1 + k;
        if (JLS_LIKELY(high < imin(escape, 32 - k))) // no escape and length <= 32: one compare
        {
            // mapped = high << k | low and the code word is 1 << k | low: flip the bits in which high differs from 1
            put<MODE>(static_cast<uint32_t>(mapped) ^ (static_cast<uint32_t>(high ^ 1) << k), length);
            return;
        }
        // long code word (more than 32 bits) or escape code: unary part in at most two pieces, then the binary part
        int32_t zeros = high < escape ? high : escape;
        if (zeros > 31)
        {
            put<MODE>(0, 31);
            zeros -= 31;
        }
        put<MODE>(1, zeros + 1);
        if (high < escape)
            put<MODE>(static_cast<uint32_t>(mapped) & ((1U << k) - 1U), k);
        else
            put<MODE>(static_cast<uint32_t>(mapped - 1) & ((1U << h.qbpp) - 1U), h.qbpp);
    }

    // reference src/scan_encoder.hpp:103-115: zero-pad to a byte; a final 0xFF is followed by a zero byte
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
            prev_ff = 0;
        }
        if (prev_ff)
        {
            emit_byte(0);
            prev_ff = 0;
        }
        const uint32_t bytes = static_cast<uint32_t>(wp - base) * 4U + pend_shift / 8U;
        if (pend_shift != 0)
            store_slot_word(wp++, bswap32(pend << (32 - pend_shift)));
        return bytes;
    }
};

// reference src/scan_encoder.hpp:53-73
template<int MODE>
JLS_HD void fast_encode_run_length(FastWriter& bw, int32_t& run_index, int32_t run_length, bool end_of_line)
{
    while (run_length >= (1 << run_order(run_index)))
    {
        bw.put<MODE>(1, 1);
        run_length -= 1 << run_order(run_index);
        if (run_index < 31)
            ++run_index;
    }
    if (end_of_line)
    {
        if (run_length != 0)
            bw.put<MODE>(1, 1);
    }
    else
    {
        bw.put<MODE>(static_cast<uint32_t>(run_length), run_order(run_index) + 1);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Bit reader over [begin, end) of one interval; bytes past `end` read as zero and are accounted for.
//
// The window is 128 bits wide (four registers).  A code word needs at most 32 valid bits, so a window that is topped up
// to > 96 bits survives several pixels without a refill: the pixel loop calls top_up() every few pixels from all lanes
// of the warp at the same time, and the refill inside get_golomb()/read() is a rarely taken fall-back.  With a 64-bit
// window some lane of the warp needed a word at nearly every symbol and the whole warp paid for the refill each time.
// ---------------------------------------------------------------------------------------------------------------------
// STEADY_BITS: the longest code word of the unchecked path (get_golomb_steady); (full_mark + 1) / STEADY_BITS - 1 symbols fit
// between two top-ups: 24 bits -> 4 symbols, 12 bits -> 8 symbols.
template<int DEPTH, int STEADY_BITS = 24>
struct FastReaderT
{
    uint32_t c3, c2, c1, c0; // the window, left aligned: c3 holds the next bits
    int32_t valid;           // valid bits in the window
    int32_t virtual_bits;    // appended bits that lie beyond the end of the interval
    int32_t remaining;       // real bytes not yet moved into the window (may go negative)
    int32_t guard;           // > 0 while the word after `cur` still holds bytes of the interval
    uint32_t prev_ff;
    uint32_t bad;         // malformed code word seen
    const uint32_t* wptr; // aligned word that holds the next byte to fetch
    uint32_t cur;         // *wptr
    uint32_t ahead;       // wptr[1], loaded one refill early so that its latency is hidden behind several pixels of work
    uint32_t ahead2;      // DEPTH == 2: wptr[2].  A top-up that takes two words (three 16-bit samples per pixel) otherwise
                          // waits for the load its own first word has just issued; the one-component decoders rarely take
                          // two words and lose 2 % to the extra move, so they stay at DEPTH 1
    uint32_t shift;       // 8 * (offset of the next byte inside *wptr)
    int32_t last_code_bits; // length of the last code word if it was a regular-mode Golomb code without escape, else 255
                            // (BitReader::last_code_bits, interval_end_status: the reference's look-up table tail)

    static constexpr int32_t full_mark = 96; // a refill appends words while valid <= full_mark

    JLS_HD void init(const uint8_t* begin, const uint8_t* end)
    {
        c3 = c2 = c1 = c0 = 0;
        valid = 0;
        virtual_bits = 0;
        prev_ff = 0;
        remaining = static_cast<int32_t>(end - begin);
        const uintptr_t address = reinterpret_cast<uintptr_t>(begin);
        wptr = reinterpret_cast<const uint32_t*>(address & ~static_cast<uintptr_t>(3));
#if defined(__CUDA_ARCH__)
        __builtin_assume(__isGlobal(wptr));
#endif
        const int32_t offset = static_cast<int32_t>(address & 3U);
        shift = static_cast<uint32_t>(offset) * 8U;
        guard = remaining - (4 - offset);
        bad = 0;
        last_code_bits = 255;
        cur = remaining > 0 ? *wptr : 0U;
        ahead = guard > 0 ? wptr[1] : 0U;
        ahead2 = DEPTH == 2 && guard > 4 ? wptr[2] : 0U;
        refill();
    }

    // appends the low `count` bits of `bits` (count in [1, 32], valid + count <= 128)
    JLS_HD void append(uint32_t bits, int32_t count)
    {
        const uint64_t v = static_cast<uint64_t>(bits) << (64 - count); // left aligned in 64 bits
        if (JLS_LIKELY(valid >= 64))
        {
            const uint64_t piece = v >> (valid - 64);
            c1 |= static_cast<uint32_t>(piece >> 32);
            c0 |= static_cast<uint32_t>(piece);
        }
        else
        {
            const uint64_t piece = v >> valid;
            c3 |= static_cast<uint32_t>(piece >> 32);
            c2 |= static_cast<uint32_t>(piece);
            if (valid > 32)
                c1 |= static_cast<uint32_t>(v >> 32) << (64 - valid); // the part that crosses into the lower half
        }
        valid += count;
    }

    // drops `count` bits, count in [0, 32]
    JLS_HD void consume(int32_t count)
    {
        const uint32_t n = static_cast<uint32_t>(count);
        c3 = funnel_l(c2, c3, n);
        c2 = funnel_l(c1, c2, n);
        c1 = funnel_l(c0, c1, n);
        c0 = funnel_l(0U, c0, n);
        valid -= count;
    }

    JLS_HD void refill_once() // valid <= full_mark on entry
    {
        const uint32_t w = bswap32(funnel_r(cur, ahead, shift)); // the next four bytes, first one on top
        cur = ahead;
        ++wptr;
        guard -= 4;
#if defined(__CUDA_ARCH__)
        __builtin_assume(__isGlobal(wptr));
#endif
        if (DEPTH == 2)
        {
            ahead = ahead2;
            ahead2 = guard > 4 ? load_stream_word(wptr + 2) : 0U; // needed two refills from now
        }
        else
        {
            ahead = guard > 0 ? load_stream_word(wptr + 1) : 0U; // needed only at the next refill
        }
        if (JLS_LIKELY(remaining >= 4 && (prev_ff | has_ff_byte(w)) == 0))
        {
            append(w, 32);
        }
        else if (remaining >= 4)
        {
            // Four real bytes with a 0xFF among them or in front: the byte after a 0xFF carries seven bits, its top bit is a
            // stuffed zero.  Straight-line: the warp walks this path whenever ONE lane meets a 0xFF (three 16-bit samples
            // per pixel: at a third of all pixels, ncu), and the byte-wise loop below costs 77 instructions.
            const uint32_t ff = has_ff_byte(w);                   // 0x80 in the bytes that are 0xFF
            const uint32_t stuffed = (ff >> 8) | (prev_ff << 31); // the top bit of every byte that follows a 0xFF
            // 0xFF followed by a byte with its top bit set is a marker, and an interval never contains one: the marker
            // search would have ended it there.  (Only a wrong side table of offsets can lead a reader into one.)
            if ((w & stuffed) != 0)
                bad |= 2U;
            uint32_t bits = w & ~stuffed;
            // squeeze bit 23, 15 and 7 out where they are stuffed, top one first: what lies above moves down by one,
            // bits = (hi << (p + 1)) + lo with bit p clear  ->  (hi << p) + lo = bits - ((bits >> 1) & (~0 << p));
            // a stuffed bit 31 is clear already and only shortens the count
            bits -= (bits >> 1) & 0xFF800000U & (0U - ((stuffed >> 23) & 1U));
            bits -= (bits >> 1) & 0xFFFF8000U & (0U - ((stuffed >> 15) & 1U));
            bits -= (bits >> 1) & 0xFFFFFF80U & (0U - ((stuffed >> 7) & 1U));
            prev_ff = (ff >> 7) & 1U;
            append(bits, 32 - popcount32(stuffed));
        }
        else
        {
            // the last bytes of the interval and the zero bytes beyond its end (bit stuffing as above, byte by byte)
            uint32_t bits = 0;
            int32_t count = 0;
            for (int32_t i = 0; i < 4; ++i)
            {
                const bool is_virtual = i >= remaining;
                const uint32_t b = is_virtual ? 0U : (w >> (24 - 8 * i)) & 0xFFU;
                if (prev_ff && (b & 0x80U) != 0)
                    bad |= 2U;
                const int32_t take = prev_ff ? 7 : 8;
                bits = (bits << take) | (b & (0xFFU >> (8 - take)));
                count += take;
                if (is_virtual)
                    virtual_bits += take;
                prev_ff = (b == 0xFFU) ? 1U : 0U;
            }
            append(bits, count);
        }
        remaining -= 4;
    }

    JLS_HD void refill() // afterwards valid > full_mark
    {
        while (valid <= full_mark)
            refill_once();
    }

    // DEPTH == 2, valid <= 64: two words in one step.  Both come from registers that were loaded at an earlier top-up, and
    // the loads issued here go straight into `ahead` and `ahead2`: no instruction reads a register whose load is still
    // in flight.  (Two refill_once() in a row cannot do that: the second one has to move the word the first one has just
    // requested, and even a move waits for the load -- ncu, cfg4: a fifth of the decoder's stall samples.)
    JLS_HD void refill_pair()
    {
        const uint32_t w1 = bswap32(funnel_r(cur, ahead, shift));
        const uint32_t w2 = bswap32(funnel_r(ahead, ahead2, shift));
        if (JLS_LIKELY(remaining >= 8 && (prev_ff | has_ff_byte(w1) | has_ff_byte(w2)) == 0))
        {
            cur = ahead2;
            wptr += 2;
            guard -= 8;
#if defined(__CUDA_ARCH__)
            __builtin_assume(__isGlobal(wptr));
#endif
            ahead = guard > 0 ? load_stream_word(wptr + 1) : 0U;
            ahead2 = guard > 4 ? load_stream_word(wptr + 2) : 0U;
            // 64 bits at bit `valid` of the window; everything below the valid bits is zero
            const uint64_t v = (static_cast<uint64_t>(w1) << 32) | w2;
            const uint64_t upper = shr64_sat(v, static_cast<uint32_t>(valid));
            const uint64_t lower = shl64_sat(v, static_cast<uint32_t>(64 - valid));
            c3 |= static_cast<uint32_t>(upper >> 32);
            c2 |= static_cast<uint32_t>(upper);
            c1 |= static_cast<uint32_t>(lower >> 32);
            c0 |= static_cast<uint32_t>(lower);
            valid += 64;
            remaining -= 8;
        }
        else
        {
            refill_once();
            refill_once();
        }
    }

    // DEPTH == 2 (three 16-bit samples per pixel take more than one word per pixel and lane), valid <= full_mark: ONE word
    // (64 < valid) or TWO (valid <= 64) in the same straight-line step.  Round 1 had a two-word step for the lanes with
    // valid <= 64 followed by the one-word loop for everybody: 5 lanes walked the warp through the first (57 instructions),
    // 26 through the second (35), at every pixel (profiles/r2_notes.md).  All words come from registers that were loaded at
    // an earlier top-up, and the loads issued here go straight into `ahead` and `ahead2`: no instruction reads a register
    // whose load is still in flight.
    JLS_HD void refill_words()
    {
        const bool two = valid <= 64;
        const uint32_t w1 = bswap32(funnel_r(cur, ahead, shift));
        const uint32_t w2 = two ? bswap32(funnel_r(ahead, ahead2, shift)) : 0U;
        if (JLS_LIKELY(remaining >= 8 && (prev_ff | has_ff_byte(w1) | has_ff_byte(w2)) == 0))
        {
            const int32_t bytes = two ? 8 : 4;
            cur = two ? ahead2 : ahead;
            wptr += two ? 2 : 1;
            guard -= bytes;
            remaining -= bytes;
#if defined(__CUDA_ARCH__)
            __builtin_assume(__isGlobal(wptr));
#endif
            // A one-word lane keeps the word it holds as `ahead2` and loads one new word, a two-word lane loads two.  The load
            // is predicated and writes the register in place: as a select between the register and a loaded value it gets a
            // move behind the load, and the move waits for the load (decode 8.6 -> 9.4 ms); loading both words in every lane
            // avoids the move but adds 70 % to the L1 wavefronts of the stream loads (32 lanes, 32 lines: 9.7 ms).
            ahead = load_stream_word_if(two && guard > 0, wptr + 1, two ? 0U : ahead2);
            ahead2 = load_stream_word_if(guard > 4, wptr + 2, 0U);
            // 32 or 64 bits at bit `valid` of the window; everything below the valid bits is zero
            const uint64_t v = (static_cast<uint64_t>(w1) << 32) | w2;
            const uint64_t upper = shr64_sat(v, static_cast<uint32_t>(valid));
            const uint64_t lower = two ? shl64_sat(v, static_cast<uint32_t>(64 - valid)) : v >> (valid - 64);
            c3 |= static_cast<uint32_t>(upper >> 32);
            c2 |= static_cast<uint32_t>(upper);
            c1 |= static_cast<uint32_t>(lower >> 32);
            c0 |= static_cast<uint32_t>(lower);
            valid += 8 * bytes;
        }
        else
        {
            refill();
        }
    }

    // what top_up() guarantees: DEPTH == 2 stops at three symbols' worth (one pixel), where a lane that arrives nearly empty
    // would need a third word
    static constexpr int32_t topped_up_bits = JLS_TOP_UP_UNIFIED && DEPTH == 2 ? 3 * STEADY_BITS : full_mark + 1;

    // the pixel loop's cadence call (all lanes of a warp together)
    JLS_HD void top_up()
    {
        if (valid <= full_mark)
        {
            if (JLS_TOP_UP_UNIFIED && DEPTH == 2)
            {
                refill_words();
                if (JLS_UNLIKELY(valid < topped_up_bits))
                    refill();
            }
            else
            {
                if (DEPTH == 2 && valid <= 64)
                    refill_pair();
                refill();
            }
        }
    }

    JLS_HD uint32_t read(int32_t count) // count in [1, 31]
    {
        if (JLS_UNLIKELY(valid <= 32))
            refill();
        const uint32_t v = c3 >> (32 - count);
        consume(count);
        return v;
    }

    JLS_HD bool overrun() const { return valid < virtual_bits; }
    JLS_HD int32_t overrun_bits() const { return virtual_bits - valid; }

    // true when bits that were not consumed are not all zero
    JLS_HD bool residue() const { return (c3 | c2 | c1 | c0) != 0; }

    // Side-table mode (interval_end_status, strict): true when what is left of the interval after the last symbol is known
    // to be zero bits only -- everything real has been moved into the window (refills first) and nothing in it is set.  A
    // marker needs a 0xFF, so the marker search would have found the same interval.
    JLS_HD bool leftover_is_zero_padding()
    {
        refill();
        return remaining <= 0 && !residue();
    }

    // whole unread bytes left in the interval after the last decoded symbol
    JLS_HD int32_t unread_bytes() const
    {
        const int32_t real_valid_bits = valid - virtual_bits;
        return (real_valid_bits > 0 ? real_valid_bits / 8 : 0) + (remaining > 0 ? remaining : 0);
    }

    // The regular-mode symbols of a line are read under a contract that needs no refill test per symbol: the pixel loop
    // tops the window up to > full_mark bits at least every `steady_symbols` symbols, the straight-line path below
    // only takes code words of up to `steady_bits` bits (full_mark + 1 - 4 * 24 >= 24), and every other way of
    // consuming bits (long code words, run mode) ends with a top_up() of its own.
    static constexpr int32_t steady_bits = STEADY_BITS;
    static constexpr int32_t steady_symbols = topped_up_bits / STEADY_BITS; // 4 (97 - 4 * 24 >= 0 ...), 8, or 3 (DEPTH == 2)
    static_assert(topped_up_bits - (steady_symbols - 1) * steady_bits >= steady_bits, "the last symbol of a group finds its bits");

    JLS_HD int32_t get_golomb_steady(const HotParams& h, int32_t k, int32_t escape)
    {
#if !defined(__CUDA_ARCH__)
        if (valid < steady_bits)
            bad = 0x80000000U; // contract broken: fails every host-emulation test
#endif
        const uint32_t top = c3;
        const int32_t z = clz32(top);
        const int32_t length = z + 1 + k;
        if (JLS_LIKELY(length <= imin(steady_bits, escape)))
        {
            // not an escape (z < length <= escape), and the code word (at most steady_bits bits) is valid and sits in c3; one
            // compare against a loop-invariant bound (see FastWriter::put_golomb)
            const uint32_t remainder = shr_sat(shl_sat(top, static_cast<uint32_t>(z + 1)), static_cast<uint32_t>(32 - k));
            last_code_bits = length;
            consume(length);
            return (z << k) + static_cast<int32_t>(remainder);
        }
        const int32_t value = get_golomb(h, k, escape);
        top_up();
        return value;
    }

    // limited-length Golomb code (reference src/scan_decoder.hpp:113-125,203-217); `bad` on a malformed code
    JLS_HD int32_t get_golomb(const HotParams& h, int32_t k, int32_t escape)
    {
        if (JLS_UNLIKELY(valid <= 32))
            refill();
        const uint32_t top = c3;
        const int32_t z = clz32(top);
        if (JLS_LIKELY(z < imin(escape, 32 - k)))
        {
            // not an escape and the whole code word (z + 1 + k bits) sits in the top 32 bits (valid > 32)
            const uint32_t remainder = shr_sat(shl_sat(top, static_cast<uint32_t>(z + 1)), static_cast<uint32_t>(32 - k));
            last_code_bits = z + 1 + k;
            consume(last_code_bits);
            return (z << k) + static_cast<int32_t>(remainder);
        }
        last_code_bits = 255; // longer than 32 bits or an escape: not in the reference's look-up table
        int32_t zeros = 0;
        for (;;)
        {
            if (valid <= 32)
                refill();
            const int32_t n = clz32(c3); // valid > 32: all of c3 is valid
            if (n < 32)
            {
                zeros += n;
                consume(n + 1);
                break;
            }
            zeros += 32;
            consume(32);
            if (overrun()) // ran off the end of the interval inside a unary code (the reference: invalid_data)
            {
                bad = 1;
                return 0;
            }
        }
        if (zeros < escape)
            return k == 0 ? zeros : (zeros << k) + static_cast<int32_t>(read(k));
        return static_cast<int32_t>(read(h.qbpp)) + 1;
    }
};

// reference src/scan_decoder_impl.hpp:305-337; -1 when the run passes the end of the line
template<typename Reader>
JLS_HD int32_t fast_decode_run_length(Reader& br, int32_t& run_index, int32_t pixel_count)
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
    br.last_code_bits = 255;
    return index > pixel_count ? -1 : index;
}

// ---------------------------------------------------------------------------------------------------------------------
// Per-line state shared by encoder and decoder
// ---------------------------------------------------------------------------------------------------------------------
// LUT_MODE: lut_none = compare chain and count-leading-zeros forms; lut_clamped = tables, the context table covers
// Ra in [0, context_lut_last] and larger Ra take its last entry; lut_full = the context table covers every sample value
// (8-bit containers: 256 entries), no clamp.  `true` / `false` select lut_clamped / lut_none.
enum : int
{
    lut_none = 0,
    lut_clamped = 1,
    lut_full = 2
};

template<int NC, int LUT_MODE, int DEPTH = 0>
struct FastLineState
{
    static constexpr bool USE_LUT = LUT_MODE != lut_none;
    RegularContext* contexts; // this thread's context q lives at contexts[q * context_stride] (device: shared memory)
    int32_t context_stride;
#if defined(__CUDA_ARCH__)
    uint32_t context_base; // shared-window address of contexts[0], see begin_interval
#endif
    RegularContext cached;
    int32_t cached_index;
    uint32_t cached_reciprocal; // USE_LUT: reciprocal_lut[cached.n], reloaded whenever cached.n changes
    RunContext run_context; // scalar lines only ever use RItype 1, multi-component pixels only RItype 0
    int32_t run_index;
    int32_t ra[NC];

    JLS_HD void store_context(int32_t index, const RegularContext& c)
    {
#if defined(__CUDA_ARCH__)
        asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};"
                     :
                     : "r"(context_base + static_cast<uint32_t>(index * context_stride) * 16U), "r"(c.a), "r"(c.b), "r"(c.c), "r"(c.n)
                     : "memory");
#else
        contexts[index * context_stride] = c;
#endif
    }

    JLS_HD RegularContext load_context(int32_t index) const
    {
#if defined(__CUDA_ARCH__)
        RegularContext c;
        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(c.a), "=r"(c.b), "=r"(c.c), "=r"(c.n)
                     : "r"(context_base + static_cast<uint32_t>(index * context_stride) * 16U)
                     : "memory");
        return c;
#else
        return contexts[index * context_stride];
#endif
    }

    // Scalar lines reach regular mode only with Ra > NEAR, i.e. never in context 0: callers may pass a `ctx` that points
    // one row in front of a four-row array (the tile kernels do: 512 bytes of shared memory per block less).
    static constexpr int32_t first_context = NC == 1 ? 1 : 0;

    JLS_HD void begin_interval(const HotParams& h, RegularContext* ctx, int32_t stride)
    {
        contexts = ctx;
        context_stride = stride;
#if defined(__CUDA_ARCH__)
        // ptxas recomputes a shared address that derives from %tid at every use (S2R, S2UR, UMOV, ULEA, LEA: five issue
        // slots per pixel) instead of holding it in a register.  A volatile round trip through the thread's own context
        // slot makes the value opaque.
        static_assert(sizeof(RegularContext) == 16, "one context is one 16-byte shared-memory access");
        volatile uint32_t* own_slot = reinterpret_cast<volatile uint32_t*>(ctx + first_context * stride);
        *own_slot = static_cast<uint32_t>(__cvta_generic_to_shared(ctx));
        context_base = *own_slot;
#endif
        const RegularContext initial = {h.a_init, 0, 0, 1};
        for (int32_t q = first_context; q < 5; ++q)
            store_context(q, initial);
        cached = initial;
        cached_index = 4;
        cached_reciprocal = reciprocal_lut_entry(1);
        run_context.a = h.a_init;
        run_context.n = 1;
        run_context.nn = 0;
        begin_line();
    }

    // every line of an interval starts from an all-zero neighbourhood and run index 0
    JLS_HD void begin_line()
    {
        run_index = 0;
#pragma unroll
        for (int32_t c = 0; c < NC; ++c)
            ra[c] = 0;
    }

    // FORCE_BRANCH (device): in the three-component kernels ptxas turns the short body into predicated instructions all the
    // same -- seven issue slots per sample that do nothing on smooth data.  A loop of (opaque) one iteration cannot be
    // predicated.  Measured on cfg4: decoder 4.56 -> 4.49 ms; the encoder does not gain and keeps the plain form, and so do
    // the one-component kernels (predicated there as well; cfg2 decoder 7.49 -> 7.64 ms with the branch, encoder unchanged).
    template<bool FORCE_BRANCH = false>
    JLS_HD void select_context(const HotParams& h, int32_t index)
    {
#if !JLS_CONTEXT_CACHE
        // no caching: the context travels shared memory -> registers -> shared memory for every sample (update() stores it)
        (void)FORCE_BRANCH;
        cached_index = index;
        cached = load_context(index);
        if (USE_LUT)
            load_reciprocal(h);
#else
        // a branch, not predication: seven instructions that a warp skips for as long as its lines stay in their contexts
        if (JLS_UNLIKELY(index != cached_index))
        {
#if defined(__CUDA_ARCH__)
            for (int32_t once = FORCE_BRANCH ? h.one : 1; once > 0; --once)
#endif
            {
                store_context(cached_index, cached);
                cached = load_context(index);
                cached_index = index;
                if (USE_LUT)
                    load_reciprocal(h);
            }
        }
#endif
    }

    // after cached.n changed
    JLS_HD void load_reciprocal(const HotParams& h)
    {
#if defined(__CUDA_ARCH__)
        asm("ld.shared.u32 %0, [%1];" : "=r"(cached_reciprocal) : "r"(h.reciprocal_lut_shared + 4U * static_cast<uint32_t>(cached.n)));
#else
        cached_reciprocal = h.reciprocal_lut[cached.n];
#endif
    }

    JLS_HD int32_t golomb_k(const HotParams& h) const
    {
        return USE_LUT ? golomb_parameter_reciprocal(cached.a, cached_reciprocal, h.two, h.one) : golomb_parameter(cached.a, cached.n);
    }

    // context update of the cached context; returns the high-water mark of fast_update_context
    template<bool LOSSLESS>
    JLS_HD uint32_t update(const HotParams& h, int32_t e)
    {
        const uint32_t water = fast_update_context<LOSSLESS>(h, cached, e);
#if !JLS_CONTEXT_CACHE
        store_context(cached_index, cached);
#else
        if (USE_LUT)
            load_reciprocal(h);
#endif
        return water;
    }

    // |Q(-Ra)|: di = -Ra <= -T3 -> 4, <= -T2 -> 3, <= -T1 -> 2, < -NEAR -> 1, else 0 (jpegls_algorithm.hpp:173-194)
    static JLS_HD int32_t context_index_compare(const HotParams& h, int32_t ra_value)
    {
        // four independent compares: a select chain has fewer instructions but its latency sits on the critical path
        return (ra_value >= h.t3) + (ra_value >= h.t2) + (ra_value >= h.t1) + (ra_value > h.near);
    }

    static JLS_HD int32_t context_index(const HotParams& h, int32_t ra_value)
    {
        if (USE_LUT)
        {
#if defined(__CUDA_ARCH__)
            uint32_t q;
            const int32_t index = LUT_MODE == lut_full ? ra_value : imin(ra_value, h.context_lut_last);
            asm("ld.shared.u8 %0, [%1];" : "=r"(q) : "r"(h.context_lut_shared + static_cast<uint32_t>(index)));
            return static_cast<int32_t>(q);
#else
            return h.context_lut[LUT_MODE == lut_full ? ra_value : imin(ra_value, h.context_lut_last)];
#endif
        }
        return context_index_compare(h, ra_value);
    }

    JLS_HD bool in_run_mode(const HotParams& h) const
    {
        if (NC == 1)
            return ra[0] <= h.near;
        int32_t m = ra[0];
#pragma unroll
        for (int32_t c = 1; c < NC; ++c)
            m = imax(m, ra[c]);
        return m <= h.near;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Encoder
// ---------------------------------------------------------------------------------------------------------------------
template<int NC, bool LOSSLESS, int LUT_MODE = lut_none, int MODE = write_deferred, int DEPTH = 0>
struct FastLineEncoder : FastLineState<NC, LUT_MODE, DEPTH>
{
    FastWriter bw;
    int32_t run_count;
    // run-mode symbols take the checked put of the steady modes (see write_steady_16)
    static constexpr int RUN_WRITE = is_steady<MODE> ? write_wide : MODE;

    JLS_HD void begin(const HotParams& h, RegularContext* ctx, int32_t stride, uint8_t* slot)
    {
        this->begin_interval(h, ctx, stride);
        bw.init(slot);
        run_count = 0;
    }

    JLS_HD void begin_line()
    {
        FastLineState<NC, LUT_MODE, DEPTH>::begin_line();
        run_count = 0;
    }

    // regular mode for one sample (reference src/scan_encoder_core.hpp:40-55); prediction = Ra, sign < 0 unless q == 0
    JLS_HD int32_t regular(const HotParams& h, int32_t x, int32_t ra_value)
    {
        const int32_t q = FastLineState<NC, LUT_MODE, DEPTH>::context_index(h, ra_value);
        this->select_context(h, q);
        RegularContext& c = this->cached;
        const int32_t k = this->golomb_k(h);
        const bool negative = NC == 1 || q != 0; // a scalar line reaches regular mode only with q != 0
        const int32_t pv = add_clamp_relu(ra_value, negative ? -c.c : c.c, h.maxval); // == correct_prediction(Ra -+ C)
        const int32_t e = fast_error_value<LOSSLESS, DEPTH>(h, negative ? pv - x : x - pv);
        // map(correction ^ e) with correction in {0, -1} equals map(e) ^ (correction & 1): the reference's XOR trick
        // (src/scan_encoder_core.hpp:48-53, src/regular_mode_context.hpp:36-42) costs one conditional bit flip here
        const bool flip = (LOSSLESS ? k : (k | h.near)) == 0 && fast_correction_sign(c);
        bw.template put_golomb<MODE>(h, k, map_error_value(e) ^ (flip ? 1 : 0), h.escape);
        this->template update<LOSSLESS>(h, e);
        return LOSSLESS ? x : fast_reconstruct<false>(h, pv, negative ? -e : e);
    }

    // run interruption (reference src/scan_encoder_core.hpp:105-138)
    JLS_HD void interruption_error(const HotParams& h, int32_t ri_type, int32_t e)
    {
        RunContext& c = this->run_context;
        const int32_t k = run_golomb_parameter(c, ri_type);
        const int32_t map = run_compute_map(c, e, k);
        const int32_t e_mapped = 2 * iabs(e) - ri_type - map;
        bw.template put_golomb<RUN_WRITE>(h, k, e_mapped, h.limit - run_order(this->run_index) - 1 - h.qbpp - 1);
        update_run_context(c, e, e_mapped, ri_type, h.reset);
    }

    // Codes one pixel (NC samples, already masked / colour transformed).
    JLS_HD void pixel(const HotParams& h, const int32_t (&x)[NC])
    {
        if (JLS_UNLIKELY(this->in_run_mode(h)))
        {
            bool same = true;
#pragma unroll
            for (int32_t c = 0; c < NC; ++c)
                same = same && iabs(x[c] - this->ra[c]) <= h.near;
            if (same)
            {
                ++run_count; // reconstructed value is Ra (scan_encoder_impl.hpp:258-265)
                return;
            }
            fast_encode_run_length<RUN_WRITE>(bw, this->run_index, run_count, false);
            run_count = 0;
#pragma unroll
            for (int32_t c = 0; c < NC; ++c)
            {
                if (NC == 1)
                {
                    // Rb = 0 and Ra <= NEAR: |Ra - Rb| <= NEAR always -> RItype 1 (scan_encoder_core.hpp:118-125)
                    const int32_t e = fast_error_value<LOSSLESS, DEPTH>(h, x[c] - this->ra[c]);
                    interruption_error(h, 1, e);
                    this->ra[c] = LOSSLESS ? x[c] : fast_reconstruct<false>(h, this->ra[c], e);
                }
                else
                {
                    // per component, RItype 0, prediction Rb = 0 (scan_encoder_core.hpp:133-138)
                    const int32_t s = sign_of(-this->ra[c]);
                    const int32_t e = fast_error_value<LOSSLESS, DEPTH>(h, s * x[c]);
                    interruption_error(h, 0, e);
                    this->ra[c] = LOSSLESS ? x[c] : fast_reconstruct<false>(h, 0, e * s);
                }
            }
            if (this->run_index > 0)
                --this->run_index;
            bw.template settle<MODE>();
            return;
        }
#pragma unroll
        for (int32_t c = 0; c < NC; ++c)
            this->ra[c] = regular(h, x[c], this->ra[c]);
    }

    // called by the pixel loop every few pixels, by all lanes at the same time (see FastWriter::put)
    JLS_HD void drain() { bw.template drain<MODE>(); }

    // pixels between two drain() calls of the pixel loop: write_wide keeps whole pixels of up to ~24 bits per sample
    // between drains (anything longer drains itself inside put)
    static constexpr int32_t pixels_per_drain =
        is_steady<MODE> ? steady_pixels_per_drain<MODE, NC> : MODE == write_wide ? (NC == 1 ? 4 : NC == 2 ? 2 : 1) : 4;

    // end of a line: a run that reaches the end of the line (reference src/scan_encoder.hpp:62-68)
    JLS_HD void end_line()
    {
        if (run_count != 0)
        {
            fast_encode_run_length<RUN_WRITE>(bw, this->run_index, run_count, true);
            run_count = 0;
        }
    }

    JLS_HD uint32_t finish()
    {
        if (MODE == write_wide || is_steady<MODE>)
            bw.template drain<write_wide>(); // finish() looks at the lower 64 bits only
        return bw.finish();
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Decoder
// ---------------------------------------------------------------------------------------------------------------------
template<int NC, bool LOSSLESS, int LUT_MODE = lut_none, int DEPTH = 0>
struct FastLineDecoder : FastLineState<NC, LUT_MODE, DEPTH>
{
    // JLS_TOP_UP_8 (A/B, off): one-component lines in 8-bit containers with eight samples between two top-ups and code words
    // of up to 12 bits on the unchecked path; everything else four samples of up to 24 bits.
    static constexpr bool long_cadence = JLS_TOP_UP_8 && NC == 1 && LUT_MODE == lut_full;
    FastReaderT<NC == 3 ? JLS_READER_DEPTH_NC3 : 1, long_cadence ? 12 : 24> br;
    // 2 * (pixels of the current run still to be output) + (1 if a run-interruption pixel follows the run): one
    // register and one test on the regular-mode path
    int32_t pending;
    uint32_t high_water; // >= sanity_limit: one of the reference's sanity checks fired in regular mode (fast_update_context)

    JLS_HD void begin(const HotParams& h, RegularContext* ctx, int32_t stride, const uint8_t* begin_, const uint8_t* end_)
    {
        this->begin_interval(h, ctx, stride);
        br.init(begin_, end_);
        pending = 0;
        high_water = 0;
    }

    JLS_HD bool bad() const { return br.bad != 0 || high_water >= sanity_limit; }

    // called by the pixel loop every few pixels, by all lanes at the same time (see FastReader)
    JLS_HD void top_up() { br.top_up(); }

    JLS_HD void begin_line()
    {
        FastLineState<NC, LUT_MODE, DEPTH>::begin_line();
        pending = 0;
    }

    // reference src/scan_decoder_core.hpp:38-69
    JLS_HD int32_t regular(const HotParams& h, int32_t ra_value)
    {
        const int32_t q = FastLineState<NC, LUT_MODE, DEPTH>::context_index(h, ra_value);
        this->template select_context<NC == 3>(h, q);
        RegularContext& c = this->cached;
        const bool negative = NC == 1 || q != 0;
        const int32_t pv = add_clamp_relu(ra_value, negative ? -c.c : c.c, h.maxval); // == correct_prediction(Ra -+ C)
        const int32_t k = this->golomb_k(h);
        const bool flip = k == 0 && (LOSSLESS || h.near == 0) && fast_correction_sign(c); // see the encoder
        // unmap(m ^ 1) == ~unmap(m): the correction becomes one predicated complement behind the unmapping
        int32_t e = fast_unmap(br.get_golomb_steady(h, k, h.escape));
        if (flip)
            e = ~e;
        // The reference's sanity checks -- k >= 16 (src/regular_mode_context.hpp:107-108), |e| > 65535
        // (src/scan_decoder_core.hpp:57-58) and the context's (:52-54) -- without branches or clamps: each quantity is
        // scaled so that its bound is sanity_limit and goes into a running maximum.  k <= 31 and |e| < 2^22 on every
        // path (24 or `escape` zeros at most, shifted by k <= 15, or k >= 16 which trips the limit by itself), so the
        // products cannot wrap.
        const uint32_t mark_k = static_cast<uint32_t>(k) << 20, mark_e = static_cast<uint32_t>(iabs(e)) << 8;
        // Lossless: the context's own check (A >= 2^24 after the update) never fires first.  |e| <= 65535 puts the A before
        // the update above 2^24 - 2^16, and N <= RESET <= 255 then gives k >= 16 for this very symbol (255 << 15 < 2^24 -
        // 2^16): the k mark has tripped already.  |B| stays below RESET + 65536.  Near-lossless keeps the mark for B.
        const uint32_t water = this->template update<LOSSLESS>(h, e);
        high_water = LOSSLESS ? umax3(high_water, mark_k, mark_e) : umax(umax3(high_water, mark_k, mark_e), water);
        return fast_reconstruct<LOSSLESS>(h, pv, negative ? -e : e);
    }

    // reference src/scan_decoder_core.hpp:72-100
    JLS_HD void interruption(const HotParams& h)
    {
        RunContext& c = this->run_context;
#pragma unroll
        for (int32_t i = 0; i < NC; ++i)
        {
            const int32_t ri_type = NC == 1 ? 1 : 0;
            const int32_t k = run_golomb_parameter(c, ri_type);
            const int32_t e_mapped = br.get_golomb(h, k, h.limit - run_order(this->run_index) - 1 - h.qbpp - 1);
            br.last_code_bits = 255;
            const int32_t e = run_error_value(c, e_mapped + ri_type, k);
            update_run_context(c, e, e_mapped, ri_type, h.reset);
            if (NC == 1)
                this->ra[i] = fast_reconstruct<LOSSLESS>(h, this->ra[i], e);
            else
                this->ra[i] = fast_reconstruct<LOSSLESS>(h, 0, e * sign_of(-this->ra[i]));
        }
        if (this->run_index > 0)
            --this->run_index;
        pending = 0;
    }

    // pixels between two top_up() calls of the pixel loop (FastReader::steady_symbols regular-mode symbols at most)
    static constexpr int32_t pixels_per_top_up = long_cadence ? 8 : NC == 1 ? 4 : NC == 2 ? 2 : 1;

    // a pixel that belongs to a run, ends one or starts one
    JLS_HD void run_mode_pixel(const HotParams& h, int32_t remaining)
    {
        if (pending > 1)
        {
            pending -= 2; // a pixel of the run: Ra stays
            return;
        }
        if (pending == 0)
        {
            const int32_t length = fast_decode_run_length(br, this->run_index, remaining);
            if (length < 0)
            {
                br.bad = 1; // reference scan_decoder_impl.hpp:328-329
                return;
            }
            pending = (length != remaining ? 1 : 0);
            if (length > 0)
            {
                pending += 2 * (length - 1);
                br.top_up();
                return;
            }
        }
        if (pending != 0)
            interruption(h);
        br.top_up();
    }

    // Decodes one pixel into this->ra. remaining_a + remaining_b = pixels left in the line including this one (the sum
    // is needed on the run-mode path only and is formed there).
    JLS_HD void pixel(const HotParams& h, int32_t remaining_a, int32_t remaining_b)
    {
        // `pending != 0` implies run mode (Ra stays at the value that started the run until the interruption sample is
        // decoded), so one test covers the pixels of a run, its interruption sample and the start of a run
        if (JLS_UNLIKELY(this->in_run_mode(h)))
        {
            run_mode_pixel(h, remaining_a + remaining_b);
            return;
        }
#pragma unroll
        for (int32_t c = 0; c < NC; ++c)
            this->ra[c] = regular(h, this->ra[c]);
    }
};

} // namespace jls
