// hostemu.cpp -- TEST INFRASTRUCTURE: runs the product's per-interval device code (charls_b200/csrc/jls_interval.cuh,
// compiled for the host) serially over the intervals of a scan, so that the kernel logic can be checked against the
// oracle in the GPU-less container.  The product never links this file; on a GPU box the same comparisons run through
// the real kernels (tests/test_gpu_*.py).
#include "jls_interval.cuh"
#include "jls_params.hpp"

#include <cstring>
#include <vector>

using namespace jls;

namespace {

// The tile kernels look the context index up in a shared-memory table; odd intervals take that variant here so that
// both it and the compare chain are checked against the oracle in every multi-line test.
bool use_lut(const CodecParams& p, uint32_t interval)
{
    return p.t3 < context_lut_capacity && p.reset < reciprocal_lut_capacity && (interval & 1U) != 0;
}

const uint32_t* reciprocals()
{
    static const std::vector<uint32_t> table = [] {
        std::vector<uint32_t> t(reciprocal_lut_capacity);
        for (int32_t i = 0; i < reciprocal_lut_capacity; ++i)
            t[static_cast<size_t>(i)] = reciprocal_lut_entry(i);
        return t;
    }();
    return table.data();
}

std::vector<uint8_t> make_lut(const CodecParams& p)
{
    std::vector<uint8_t> lut(context_lut_capacity, 0);
    for (int32_t i = 0; i <= std::max(std::min(p.t3, context_lut_capacity - 1), 255); ++i) // 8-bit containers: every value
        lut[static_cast<size_t>(i)] = context_lut_entry(p, i);
    return lut;
}

} // namespace

extern "C" {

// Exhaustive check of the table form of the Golomb parameter against the definition (min k with (n << k) >= a) for
// n in [1, n_last], a in [0, a_end); returns the number of mismatches.
uint64_t hostemu_check_golomb_parameter(int32_t n_last, int32_t a_end)
{
    uint64_t mismatches = 0;
    for (int32_t n = 1; n <= n_last; ++n)
    {
        const uint32_t reciprocal = reciprocal_lut_entry(n);
        int32_t k = 0;
        for (int32_t a = 0; a < a_end; ++a)
        {
            while ((static_cast<int64_t>(n) << k) < a)
                ++k; // the definition, incrementally (k only grows with a)
            mismatches += golomb_parameter_reciprocal(a, reciprocal) != k || golomb_parameter(a, n) != k;
        }
    }
    return mismatches;
}

// returns 0 on success
int hostemu_make_params(CodecParams* out, int32_t width, int32_t height, int32_t bits, int32_t components, int32_t near_lossless,
                        int32_t interleave, int32_t transform, int32_t t1, int32_t t2, int32_t t3, int32_t reset,
                        uint32_t restart_interval)
{
    const PresetCodingParameters pc{(1 << bits) - 1, t1, t2, t3, reset};
    *out = make_codec_params(width, height, bits, components, near_lossless, interleave, transform, pc, restart_interval);
    return 0;
}

int hostemu_uses_fast_path(const CodecParams* p) { return use_fast_path(*p) ? 1 : 0; }

// Encodes one scan: per-interval coding into slots, then the serial equivalent of k_scan_offsets + k_gather.
// Returns bytes written or -(errc).
int64_t hostemu_encode_scan(const CodecParams* pp, const uint8_t* pixels, size_t stride, uint8_t* out, size_t capacity,
                            int force_general)
{
    const CodecParams& p = *pp;
    const size_t slot_bytes = worst_case_interval_bytes(p, p.lines_per_interval);
    std::vector<uint8_t> slots(slot_bytes * p.interval_count + 64);
    std::vector<uint32_t> interval_bytes(p.interval_count);
    std::vector<uint16_t> line_scratch(static_cast<size_t>(2) * p.components * (p.width + 2) * p.interval_count);
    ScanJob job{};
    job.pixels_in = pixels;
    job.stride = stride;
    job.slots = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(slots.data()) + 15) & ~uintptr_t{15});
    job.interval_bytes = interval_bytes.data();
    job.line_scratch = line_scratch.data();

    const bool fast = use_fast_path(p) && !force_general;
    const bool lossless = p.near == 0;
    const std::vector<uint8_t> lut = make_lut(p);
    RegularContext contexts[5];
    uint64_t first_error = ~0ULL;
    for (uint32_t i = 0; i < p.interval_count; ++i)
    {
        IntervalResult r;
        if (fast)
        {
#define HOSTEMU_ENCODE(NC, LL, LINE)                                                                                       \
    (use_lut(p, i) ? (p.sample_bytes == 2 ? encode_interval_fast<NC, LL, uint16_t, LINE, true>(p, job, i, contexts, 1, slot_bytes, lut.data(), reciprocals()) \
                                          : encode_interval_fast<NC, LL, uint8_t, LINE, lut_full>(p, job, i, contexts, 1, slot_bytes, lut.data(), reciprocals()))   \
                   : (p.sample_bytes == 2 ? encode_interval_fast<NC, LL, uint16_t, LINE>(p, job, i, contexts, 1, slot_bytes)                   \
                                          : encode_interval_fast<NC, LL, uint8_t, LINE>(p, job, i, contexts, 1, slot_bytes)))
            if (p.interleave == ilv_sample && p.components == 2)
                r = lossless ? HOSTEMU_ENCODE(2, true, false) : HOSTEMU_ENCODE(2, false, false);
            else if (p.interleave == ilv_sample && p.components == 4)
                r = lossless ? HOSTEMU_ENCODE(4, true, false) : HOSTEMU_ENCODE(4, false, false);
            else if (p.interleave == ilv_sample)
                r = lossless ? HOSTEMU_ENCODE(3, true, false) : HOSTEMU_ENCODE(3, false, false);
            else if (p.interleave == ilv_line)
                r = lossless ? HOSTEMU_ENCODE(1, true, true) : HOSTEMU_ENCODE(1, false, true);
            else
                r = lossless ? HOSTEMU_ENCODE(1, true, false) : HOSTEMU_ENCODE(1, false, false);
        }
        else
        {
            std::vector<RegularContext> general(general_context_count);
            r = lossless ? encode_interval_general<true>(p, job, i, slot_bytes, general.data())
                         : encode_interval_general<false>(p, job, i, slot_bytes, general.data());
        }
        interval_bytes[i] = r.bytes;
        if (r.errc != err_none)
            first_error = std::min(first_error, (static_cast<uint64_t>(i) << 8) | static_cast<uint64_t>(r.errc));
    }
    if (first_error != ~0ULL)
        return -static_cast<int64_t>(first_error & 0xFF);

    size_t position = 0;
    for (uint32_t i = 0; i < p.interval_count; ++i)
    {
        const size_t need = interval_bytes[i] + (i + 1 < p.interval_count ? 2U : 0U);
        if (position + need > capacity)
            return -err_destination_too_small;
        std::memcpy(out + position, job.slots + static_cast<size_t>(i) * slot_bytes, interval_bytes[i]);
        position += interval_bytes[i];
        if (i + 1 < p.interval_count)
        {
            out[position++] = 0xFF;
            out[position++] = static_cast<uint8_t>(0xD0 + (i & 7U));
        }
    }
    return static_cast<int64_t>(position);
}

// Decodes one scan: serial marker table (what k_marker_* + k_decode_finish compute), then per-interval decoding.
// Returns bytes consumed or -(errc).
int64_t hostemu_decode_scan(const CodecParams* pp, const uint8_t* stream, size_t size, uint8_t* pixels, size_t stride,
                            int force_general)
{
    const CodecParams& p = *pp;
    // the bit reader loads aligned 32-bit words: give it an aligned, padded copy like the engine's device buffers
    std::vector<uint32_t> aligned((size + 16) / 4 + 4, 0);
    uint8_t* data = reinterpret_cast<uint8_t*>(aligned.data());
    std::memcpy(data, stream, size);

    std::vector<uint64_t> offsets(2 * static_cast<size_t>(p.interval_count), ~0ULL);
    std::vector<uint8_t> codes(p.interval_count, 0);
    uint32_t found = 0;
    offsets[0] = 0;
    for (size_t i = 1; i < size && found < p.interval_count; ++i)
    {
        if (data[i - 1] == 0xFF && data[i] >= 0x80 && data[i] != 0xFF)
        {
            size_t begin = i - 1;
            while (begin > 0 && data[begin - 1] == 0xFF)
                --begin;
            offsets[2 * static_cast<size_t>(found) + 1] = begin;
            if (found + 1 < p.interval_count)
                offsets[2 * static_cast<size_t>(found) + 2] = i + 1;
            codes[found] = data[i];
            ++found;
        }
    }

    std::vector<uint16_t> line_scratch(static_cast<size_t>(2) * p.components * (p.width + 2) * p.interval_count);
    ScanJob job{};
    job.pixels_out = pixels;
    job.stride = stride;
    job.stream_in = data;
    job.stream_in_size = size;
    job.interval_offset = offsets.data();
    job.line_scratch = line_scratch.data();

    const bool fast = use_fast_path(p) && !force_general;
    const bool lossless = p.near == 0;
    const std::vector<uint8_t> lut = make_lut(p);
    RegularContext contexts[5];
    uint64_t first_error = ~0ULL;
    auto report = [&](uint32_t interval, int32_t errc) {
        first_error = std::min(first_error, (static_cast<uint64_t>(interval) << 8) | static_cast<uint64_t>(errc));
    };
    for (uint32_t i = 0; i < p.interval_count; ++i)
    {
        IntervalResult r;
        if (fast)
        {
#define HOSTEMU_DECODE(NC, LL, LINE)                                                                                       \
    (use_lut(p, i) ? (p.sample_bytes == 2 ? decode_interval_fast<NC, LL, uint16_t, LINE, true>(p, job, i, contexts, 1, lut.data(), reciprocals()) \
                                          : decode_interval_fast<NC, LL, uint8_t, LINE, lut_full>(p, job, i, contexts, 1, lut.data(), reciprocals()))   \
                   : (p.sample_bytes == 2 ? decode_interval_fast<NC, LL, uint16_t, LINE>(p, job, i, contexts, 1)                   \
                                          : decode_interval_fast<NC, LL, uint8_t, LINE>(p, job, i, contexts, 1)))
            if (p.interleave == ilv_sample && p.components == 2)
                r = lossless ? HOSTEMU_DECODE(2, true, false) : HOSTEMU_DECODE(2, false, false);
            else if (p.interleave == ilv_sample && p.components == 4)
                r = lossless ? HOSTEMU_DECODE(4, true, false) : HOSTEMU_DECODE(4, false, false);
            else if (p.interleave == ilv_sample)
                r = lossless ? HOSTEMU_DECODE(3, true, false) : HOSTEMU_DECODE(3, false, false);
            else if (p.interleave == ilv_line)
                r = lossless ? HOSTEMU_DECODE(1, true, true) : HOSTEMU_DECODE(1, false, true);
            else
                r = lossless ? HOSTEMU_DECODE(1, true, false) : HOSTEMU_DECODE(1, false, false);
        }
        else
        {
            std::vector<RegularContext> general(general_context_count);
            r = lossless ? decode_interval_general<true>(p, job, i, general.data()) : decode_interval_general<false>(p, job, i, general.data());
        }
        if (r.errc != err_none)
            report(i, r.errc);
    }
    // k_decode_finish
    for (uint32_t i = 0; i + 1 < p.interval_count && i < found; ++i)
    {
        if (codes[i] != 0xD0U + (i & 7U))
        {
            report(i, err_restart_marker_not_found);
            break;
        }
    }
    if (found < p.interval_count && (first_error >> 8) != found)
        report(found, err_need_more_data);
    if (first_error != ~0ULL)
        return -static_cast<int64_t>(first_error & 0xFF);
    return static_cast<int64_t>(offsets[2 * static_cast<size_t>(p.interval_count - 1) + 1]);
}

size_t hostemu_sizeof_params() { return sizeof(CodecParams); }

} // extern "C"
