// jls_interval.cuh -- what ONE thread does for ONE restart interval (encode or decode), fast and general path.
//
// Kept apart from the __global__ wrappers so that the unit tests can run exactly this code on the CPU
// (tests/hostemu) -- the kernels in jls_kernels.cu add nothing but the thread -> interval mapping, the shared-memory
// context storage and the atomic error reporting.
#pragma once

#include "jls_fast.cuh"

namespace jls {

// Outcome of coding one interval: 0 or the charls_jpegls_errc to report for it.
struct IntervalResult
{
    int32_t errc;
    uint32_t bytes; // encode: size of the interval's entropy data
};

// One component of pixel x of a line-interleaved source line (masking / colour transform like load_pixel).
JLS_HD int32_t load_line_component(const CodecParams& p, const uint8_t* line, int32_t x, int32_t c)
{
    if (p.components == 3)
    {
        int32_t v[3];
        load_pixel<3>(p, line, x, v);
        return v[c];
    }
    const int32_t mask = (1 << p.bits_per_sample) - 1;
    return load_sample(line, x * p.components + c, p.sample_bytes) & mask;
}

// ---------------------------------------------------------------------------------------------------------------------
// Fast path, restart interval = 1 line.  S = sample type of the caller's buffer (uint8_t / uint16_t).
// ---------------------------------------------------------------------------------------------------------------------
// Tells the compiler a pointer is a global-memory address (so it emits LDG/STG instead of generic LD/ST).
JLS_HD void assume_global(const void* pointer)
{
#if defined(__CUDA_ARCH__)
    __builtin_assume(__isGlobal(pointer));
#else
    (void)pointer;
#endif
}

template<typename S>
JLS_HD int32_t fast_load(const S* line, int32_t index)
{
#if defined(__CUDA_ARCH__)
    return static_cast<int32_t>(__ldg(line + index));
#else
    return static_cast<int32_t>(line[index]);
#endif
}

// pixel x of a caller line -> NC internal samples: masks unused high bits, or applies the forward colour transform
// (the transforming variants do not mask, reference src/copy_to_line_buffer.hpp:234-246)
template<int NC, typename S>
JLS_HD void fast_load_pixel(const CodecParams& p, const HotParams& h, const S* line, int32_t x, int32_t (&v)[NC])
{
    for (int32_t c = 0; c < NC; ++c)
        v[c] = fast_load(line, x * NC + c);
    if (NC == 3 && p.transform != 0)
    {
        color_forward(p.transform, sizeof(S) == 2 ? 0xFFFF : 0xFF, v[0], v[NC > 1 ? 1 : 0], v[NC > 2 ? 2 : 0]);
    }
    else if (h.bits != static_cast<int32_t>(8 * sizeof(S)))
    {
        for (int32_t c = 0; c < NC; ++c)
            v[c] &= h.maxval;
    }
}

template<int NC, typename S>
JLS_HD void fast_store_pixel(const CodecParams& p, S* line, int32_t x, const int32_t (&v)[NC])
{
    int32_t o[NC];
    for (int32_t c = 0; c < NC; ++c)
        o[c] = v[c];
    if (NC == 3 && p.transform != 0)
        color_inverse(p.transform, sizeof(S) == 2 ? 0xFFFF : 0xFF, o[0], o[NC > 1 ? 1 : 0], o[NC > 2 ? 2 : 0]);
    for (int32_t c = 0; c < NC; ++c)
        line[x * NC + c] = static_cast<S>(o[c]);
}

// LINE_ILV: the interval holds one line of each of p.components components (line interleave); NC must be 1 then.
// USE_LUT: context_lut holds context_lut_entry(p, 0 .. min(T3, capacity - 1)), reciprocal_lut holds
// reciprocal_lut_entry(0 .. RESET) (RESET < reciprocal_lut_capacity).
template<int NC, bool LOSSLESS, typename S, bool LINE_ILV, int USE_LUT = lut_none>
JLS_HD IntervalResult encode_interval_fast(const CodecParams& p, const ScanJob& job, uint32_t interval,
                                           RegularContext* contexts, int32_t context_stride, size_t slot_bytes,
                                           const uint8_t* context_lut = nullptr, const uint32_t* reciprocal_lut = nullptr)
{
    HotParams h = make_hot_params(p);
    h.context_lut = context_lut;
    h.context_lut_last = imin(p.t3, context_lut_capacity - 1);
    h.reciprocal_lut = reciprocal_lut;
    FastLineEncoder<NC, LOSSLESS, USE_LUT, writer_mode<NC, LOSSLESS, S>> enc;
    constexpr int32_t drain_mask = decltype(enc)::pixels_per_drain - 1;
    uint8_t* slot = job.slots + static_cast<size_t>(interval) * slot_bytes;
    assume_global(slot);
    enc.begin(h, contexts, context_stride, slot);
    const S* line = reinterpret_cast<const S*>(job.pixels_in + static_cast<size_t>(interval) * job.stride);
    const int32_t width = p.width;

    if constexpr (LINE_ILV)
    {
        // one line of every component forms the interval; contexts are shared, the run index restarts per component
        // (reference src/scan_decoder_impl.hpp:76-117,122-127)
        static_assert(NC == 1, "line interleave codes scalar lines");
        const uint8_t* bytes = reinterpret_cast<const uint8_t*>(line);
        for (int32_t c = 0; c < p.components; ++c)
        {
            enc.begin_line();
            for (int32_t x = 0; x < width; ++x)
            {
                if ((x & drain_mask) == drain_mask)
                    enc.drain();
                const int32_t v[1] = {load_line_component(p, bytes, x, c)};
                enc.pixel(h, v);
            }
            enc.end_line();
        }
    }
    else
    {
        for (int32_t x = 0; x < width; ++x)
        {
            if ((x & drain_mask) == drain_mask)
                enc.drain();
            int32_t v[NC];
            fast_load_pixel<NC, S>(p, h, line, x, v);
            enc.pixel(h, v);
        }
        enc.end_line();
    }

    IntervalResult result;
    result.bytes = enc.finish();
    result.errc = err_none; // the slot is large enough by construction; invalid states cannot arise from valid samples
    return result;
}

// What the reference checks when an interval / the scan ends (src/scan_decoder.hpp:71-89,335-349).
// strict: the interval came from a side table of offsets, not from the marker search.  It then counts only if the search
// would have produced the very same interval, i.e. if no marker hides in bytes the decoder did not look at: whatever is
// left after the last symbol must be zero padding.  Anything else rejects the table (the engine decodes again without it).
template<typename Reader>
JLS_HD int32_t interval_end_status(const CodecParams& p, Reader& br, bool bad, uint32_t interval,
                                   bool closing_marker_found = true, bool strict = false)
{
    if (bad)
        return err_invalid_data;
    if (strict && !br.leftover_is_zero_padding())
        return errc_offset_table_rejected;
    if (br.overrun())
    {
        // Bits beyond the end of the interval were consumed.  The reference throws invalid_data when it needs bits and finds
        // none -- except for the tail of a code word it took from its 8-bit look-up table: it peeks a byte (missing bits read
        // as zero, src/scan_decoder.hpp:144-152), skips the code's length without looking at valid_bits_ (:58-64), and if that
        // was the interval's last symbol nothing ever asks for bits again (restart marker and end_scan read bytes, :71-89,
        // 237-243).  At least one bit of the code must be real (fill_read_cache throws on an empty cache, :272-284).
        const int32_t missing = br.overrun_bits();
        if (!(br.last_code_bits <= 8 && missing < br.last_code_bits))
            return err_invalid_data;
    }
    if (!closing_marker_found)
    {
        // The data ends without a marker.  The reference's read pointer runs at most one cache (8 bytes) ahead of the
        // last symbol: at the end of the data it reports need_more_data (left to the finish step), else invalid_data.
        return br.unread_bytes() > 7 ? err_invalid_data : err_none;
    }
    if (interval + 1 == p.interval_count)
    {
        // left-over bits must be zero padding and the closing marker must follow directly
        return (br.residue() || br.unread_bytes() > 0) ? err_invalid_data : err_none;
    }
    // the reference looks for RSTm where its 64-bit read cache stopped
    return br.unread_bytes() > 7 ? err_restart_marker_not_found : err_none;
}

template<int NC, bool LOSSLESS, typename S, bool LINE_ILV, int USE_LUT = lut_none>
JLS_HD IntervalResult decode_interval_fast(const CodecParams& p, const ScanJob& job, uint32_t interval,
                                           RegularContext* contexts, int32_t context_stride,
                                           const uint8_t* context_lut = nullptr, const uint32_t* reciprocal_lut = nullptr)
{
    IntervalResult result = {err_none, 0};
    // interval_offset holds 2 entries per interval: [2i] = first byte, [2i+1] = end (first 0xFF of the closing marker)
    const uint64_t begin = job.interval_offset[2 * static_cast<size_t>(interval)];
    uint64_t end = job.interval_offset[2 * static_cast<size_t>(interval) + 1];
    if (begin == ~0ULL)
        return result; // an earlier restart marker is missing: the finish step reports it
    const bool closing_marker_found = end != ~0ULL;
    if (!closing_marker_found)
        end = job.stream_in_size; // no closing marker: decode what is there (the reference runs dry -> invalid_data)
    if (begin > end)
        return result;

    HotParams h = make_hot_params(p);
    h.context_lut = context_lut;
    h.context_lut_last = imin(p.t3, context_lut_capacity - 1);
    h.reciprocal_lut = reciprocal_lut;
    FastLineDecoder<NC, LOSSLESS, USE_LUT> dec;
    dec.begin(h, contexts, context_stride, job.stream_in + begin, job.stream_in + end);
    S* line = reinterpret_cast<S*>(job.pixels_out + static_cast<size_t>(interval) * job.stride);
    assume_global(line);
    const int32_t width = p.width;

    if constexpr (LINE_ILV)
    {
        static_assert(NC == 1, "line interleave codes scalar lines");
        {
            const int32_t nc = p.components;
            uint8_t* bytes = reinterpret_cast<uint8_t*>(line);
            for (int32_t c = 0; c < nc; ++c)
            {
                dec.begin_line();
                for (int32_t x = 0; x < width; ++x)
                {
                    if (x % dec.pixels_per_top_up == 0)
                        dec.top_up();
                    dec.pixel(h, width, -x);
                    line[x * nc + c] = static_cast<S>(dec.ra[0]);
                }
            }
            if (nc == 3 && p.transform != 0)
            {
                // inverse colour transform in place once all three component lines are there
                for (int32_t x = 0; x < width; ++x)
                {
                    int32_t v[3];
                    for (int32_t c = 0; c < 3; ++c)
                        v[c] = load_sample(bytes, x * 3 + c, p.sample_bytes);
                    store_pixel<3>(p, bytes, x, v);
                }
            }
        }
    }
    else
    {
        for (int32_t x = 0; x < width; ++x)
        {
            if (x % dec.pixels_per_top_up == 0)
                dec.top_up();
            dec.pixel(h, width, -x);
            fast_store_pixel<NC, S>(p, line, x, dec.ra);
        }
    }
    result.errc = interval_end_status(p, dec.br, dec.bad(), interval, closing_marker_found, job.offset_table.total != 0);
    return result;
}

// ---------------------------------------------------------------------------------------------------------------------
// General path: any restart interval, full 2-D LOCO-I
// ---------------------------------------------------------------------------------------------------------------------
JLS_HD void general_set_edges(uint16_t* cur, uint16_t* prev, int32_t nc, int32_t width)
{
    const int32_t ps = width + 2;
    for (int32_t c = 0; c < nc; ++c)
    {
        prev[c * ps + width + 1] = prev[c * ps + width]; // reference src/scan_codec.hpp:189-195
        cur[c * ps] = prev[c * ps + 1];
    }
}

template<bool LOSSLESS>
JLS_HD_NOINLINE IntervalResult encode_interval_general(const CodecParams& p, const ScanJob& job, uint32_t interval,
                                                       size_t slot_bytes, RegularContext* contexts, const int8_t* quant_lut = nullptr)
{
    const int32_t width = p.width, ps = width + 2, nc = p.components;
    const size_t per_interval = static_cast<size_t>(2) * nc * ps;
    uint16_t* lines = job.line_scratch + static_cast<size_t>(interval) * per_interval;
    for (size_t i = 0; i < per_interval; ++i)
        lines[i] = 0;

    GeneralState state;
    state.contexts = contexts; // general_context_count entries owned by this thread (shared memory in the kernels)
    state.quant_lut = quant_lut;
    state.reset(p);
    state.bad = false;
    int32_t run_index[4] = {0, 0, 0, 0};
    BitWriter bw;
    bw.init(job.slots + static_cast<size_t>(interval) * slot_bytes, slot_bytes);

    const uint32_t first_line = interval * p.lines_per_interval;
    const uint32_t line_end = static_cast<uint32_t>(
        imin(static_cast<int32_t>(first_line + p.lines_per_interval), p.height)); // heights are <= 100000
    for (uint32_t line = first_line; line < line_end; ++line)
    {
        uint16_t* cur = lines + ((line - first_line) & 1U) * (static_cast<size_t>(nc) * ps);
        uint16_t* prev = lines + (((line - first_line) & 1U) ^ 1U) * (static_cast<size_t>(nc) * ps);
        const uint8_t* source = job.pixels_in + static_cast<size_t>(line) * job.stride;

        // caller layout -> component lines (reference src/copy_to_line_buffer.hpp:26-261)
        for (int32_t x = 0; x < width; ++x)
        {
            if (nc == 1)
            {
                int32_t v[1];
                load_pixel<1>(p, source, x, v);
                cur[x + 1] = static_cast<uint16_t>(v[0]);
            }
            else
            {
                for (int32_t c = 0; c < nc; ++c)
                    cur[c * ps + x + 1] = static_cast<uint16_t>(load_line_component(p, source, x, c));
            }
        }
        general_set_edges(cur, prev, nc, width);
        if (p.interleave == ilv_sample)
        {
            state.run_index = run_index[0];
            general_encode_line_multi<LOSSLESS>(p, state, bw, cur, prev);
            run_index[0] = state.run_index;
        }
        else
        {
            for (int32_t c = 0; c < nc; ++c)
            {
                state.run_index = run_index[c];
                general_encode_line<LOSSLESS>(p, state, bw, cur + c * ps, prev + c * ps);
                run_index[c] = state.run_index;
            }
        }
    }

    IntervalResult result;
    result.bytes = bw.finish();
    result.errc = bw.overflow ? err_destination_too_small : (state.bad ? err_invalid_data : err_none);
    return result;
}

template<bool LOSSLESS>
JLS_HD_NOINLINE IntervalResult decode_interval_general(const CodecParams& p, const ScanJob& job, uint32_t interval,
                                                       RegularContext* contexts, const int8_t* quant_lut = nullptr)
{
    IntervalResult result = {err_none, 0};
    const uint64_t begin = job.interval_offset[2 * static_cast<size_t>(interval)];
    uint64_t end = job.interval_offset[2 * static_cast<size_t>(interval) + 1];
    if (begin == ~0ULL)
        return result;
    const bool closing_marker_found = end != ~0ULL;
    if (!closing_marker_found)
        end = job.stream_in_size;
    if (begin > end)
        return result;

    const int32_t width = p.width, ps = width + 2, nc = p.components;
    const size_t per_interval = static_cast<size_t>(2) * nc * ps;
    uint16_t* lines = job.line_scratch + static_cast<size_t>(interval) * per_interval;
    for (size_t i = 0; i < per_interval; ++i)
        lines[i] = 0;

    GeneralState state;
    state.contexts = contexts; // general_context_count entries owned by this thread (shared memory in the kernels)
    state.quant_lut = quant_lut;
    state.reset(p);
    state.bad = false;
    int32_t run_index[4] = {0, 0, 0, 0};
    BitReader br;
    br.init(job.stream_in + begin, job.stream_in + end);

    const uint32_t first_line = interval * p.lines_per_interval;
    const uint32_t line_end = static_cast<uint32_t>(imin(static_cast<int32_t>(first_line + p.lines_per_interval), p.height));
    for (uint32_t line = first_line; line < line_end && !state.bad; ++line)
    {
        uint16_t* cur = lines + ((line - first_line) & 1U) * (static_cast<size_t>(nc) * ps);
        uint16_t* prev = lines + (((line - first_line) & 1U) ^ 1U) * (static_cast<size_t>(nc) * ps);
        general_set_edges(cur, prev, nc, width);
        if (p.interleave == ilv_sample)
        {
            state.run_index = run_index[0];
            general_decode_line_multi<LOSSLESS>(p, state, br, cur, prev);
            run_index[0] = state.run_index;
        }
        else
        {
            for (int32_t c = 0; c < nc && !state.bad; ++c)
            {
                state.run_index = run_index[c];
                general_decode_line<LOSSLESS>(p, state, br, cur + c * ps, prev + c * ps);
                run_index[c] = state.run_index;
            }
        }
        if (state.bad)
            break;

        // component lines -> caller layout (reference src/copy_from_line_buffer.hpp:24-191)
        uint8_t* destination = job.pixels_out + static_cast<size_t>(line) * job.stride;
        for (int32_t x = 0; x < width; ++x)
        {
            if (nc == 3)
            {
                const int32_t v[3] = {cur[x + 1], cur[ps + x + 1], cur[2 * ps + x + 1]};
                store_pixel<3>(p, destination, x, v);
            }
            else
            {
                for (int32_t c = 0; c < nc; ++c)
                    store_sample(destination, x * nc + c, p.sample_bytes, cur[c * ps + x + 1]);
            }
        }
    }
    result.errc = interval_end_status(p, br, state.bad || br.marker_inside, interval, closing_marker_found, job.offset_table.total != 0);
    return result;
}

// Which intervals take the fast path: exactly one line per interval (scalar lines or 2..4-component pixels).
inline bool use_fast_path(const CodecParams& p)
{
    return p.lines_per_interval == 1;
}

} // namespace jls
