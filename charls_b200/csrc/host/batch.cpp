// batch.cpp -- charlsx_batch_*: frames that already live in device memory (extension, not in the reference).
//
// The per-frame container handling is the same host code the single-image ABI uses (StreamWriter / StreamReader); only
// the headers ever cross PCIe.  All frames of a batch share geometry and coding parameters, so one kernel launch per
// stage covers every line of every frame (grid.y = frame).
#include "../engine.hpp"
#include "abi_support.hpp"
#include "stream_reader.hpp"
#include "stream_writer.hpp"

#include <vector>

using namespace jls;
using namespace jls::host;

struct charlsx_batch final
{
    Engine engine;
    std::vector<BatchFrame> frames;
};

namespace {

constexpr uint32_t header_prefix_bytes = 1024;

void validate_params(const charlsx_batch_params& bp)
{
    const charls_frame_info& f = bp.frame_info;
    check_range(1U, maximum_width, f.width, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_WIDTH);
    check_range(1U, maximum_height, f.height, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_HEIGHT);
    check_range(2, 16, f.bits_per_sample, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_BITS_PER_SAMPLE);
    check_range(0, 2, bp.interleave_mode, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_INTERLEAVE_MODE);
    // single-scan frames only: one component, or 2..4 interleaved components
    if (bp.interleave_mode == 0)
        check_argument(f.component_count == 1, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_INTERLEAVE_MODE);
    else
        check_range(2, maximum_component_count_in_scan, f.component_count, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_INTERLEAVE_MODE);
    check_range(0, maximum_near_lossless(maximum_bit_sample_value(f.bits_per_sample)), bp.near_lossless,
                CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_NEAR_LOSSLESS);
    check_range(0, 3, bp.color_transformation, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_COLOR_TRANSFORMATION);
    if (bp.color_transformation != 0)
        check_argument(f.component_count == 3 && (f.bits_per_sample == 8 || f.bits_per_sample == 16) && bp.near_lossless == 0,
                       CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_COLOR_TRANSFORMATION);
}

size_t effective_stride(const charlsx_batch_params& bp)
{
    const charls_frame_info& f = bp.frame_info;
    const size_t minimum = static_cast<size_t>(f.width) * static_cast<size_t>((f.bits_per_sample + 7) / 8) *
                           (bp.interleave_mode == 0 ? 1U : static_cast<size_t>(f.component_count));
    if (bp.stride == 0)
        return minimum;
    check_argument(bp.stride >= minimum, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_STRIDE);
    return bp.stride;
}

} // namespace

extern "C" {

charlsx_batch* charlsx_batch_create(void) noexcept
{
    return new (std::nothrow) charlsx_batch;
}

void charlsx_batch_destroy(charlsx_batch* batch) noexcept
{
    delete batch;
}

} // extern "C"

namespace {

charls_jpegls_errc encode_frames(charlsx_batch* batch, const charlsx_batch_params* params, charlsx_batch_image* images, size_t count,
                                 void* cuda_stream, bool host_memory) noexcept
{
    return guarded([&] {
        check_pointer(batch);
        const charlsx_batch_params& bp = *check_pointer(params);
        check_argument(images != nullptr || count == 0);
        validate_params(bp);
        const size_t stride = effective_stride(bp);
        const charls_frame_info& f = bp.frame_info;

        // the header every frame gets: SOI [APP8 mrfx] SOF55 [DRI] [APP11 side table of interval offsets] SOS -- what the
        // single-image encoder writes
        const bool with_table = (bp.flags & CHARLSX_BATCH_OFFSET_TABLE) != 0 && bp.restart_interval != 0;
        const uint32_t lines_per_interval = bp.restart_interval != 0 && bp.restart_interval < f.height ? bp.restart_interval : f.height;
        const uint32_t intervals = (f.height + lines_per_interval - 1) / lines_per_interval;
        std::vector<uint8_t> header_bytes(128 + (with_table ? offset_table_bytes(intervals) : 0));
        uint8_t* const header = header_bytes.data();
        StreamWriter writer;
        writer.destination(header, header_bytes.size());
        writer.write_start_of_image();
        if (bp.color_transformation != 0)
            writer.write_color_transform(bp.color_transformation);
        if (writer.write_start_of_frame(f))
            writer.write_oversize_dimensions(f.height, f.width);
        if (bp.restart_interval != 0)
            writer.write_define_restart_interval(bp.restart_interval);
        StreamOffsetTable table{};
        if (with_table)
        {
            writer.write_offset_table_placeholder(intervals, table.entry_offsets);
            table.total = intervals + 1;
        }
        writer.write_start_of_scan(f.component_count, bp.near_lossless, bp.interleave_mode);

        const PresetCodingParameters preset = default_preset_parameters(maximum_bit_sample_value(f.bits_per_sample), bp.near_lossless);
        const CodecParams p = make_codec_params(static_cast<int32_t>(f.width), static_cast<int32_t>(f.height), f.bits_per_sample,
                                                f.component_count, bp.near_lossless, bp.interleave_mode,
                                                bp.interleave_mode != 0 ? bp.color_transformation : 0, preset, bp.restart_interval);
        batch->frames.resize(count);
        for (size_t i = 0; i < count; ++i)
        {
            check_argument(images[i].pixels != nullptr && images[i].stream != nullptr);
            batch->frames[i] = BatchFrame{static_cast<uint8_t*>(images[i].pixels), static_cast<uint8_t*>(images[i].stream),
                                          images[i].stream_capacity, 0, 0, 0};
        }
        const int32_t status =
            host_memory ? batch->engine.encode_batch_host(p, header, writer.bytes_written(), batch->frames.data(), count, stride,
                                                          table.total != 0 ? &table : nullptr)
                        : batch->engine.encode_batch(p, header, writer.bytes_written(), batch->frames.data(), count, stride,
                                                     static_cast<CUstream_st*>(cuda_stream), {}, table.total != 0 ? &table : nullptr);
        for (size_t i = 0; i < count; ++i)
        {
            images[i].stream_size = batch->frames[i].stream_size;
            images[i].status = batch->frames[i].status;
        }
        check_status(status);
    });
}

} // namespace

extern "C" {

charls_jpegls_errc charlsx_batch_encode(charlsx_batch* batch, const charlsx_batch_params* params, charlsx_batch_image* images,
                                        size_t count, void* cuda_stream) noexcept
{
    return encode_frames(batch, params, images, count, cuda_stream, false);
}

charls_jpegls_errc charlsx_batch_encode_host(charlsx_batch* batch, const charlsx_batch_params* params, charlsx_batch_image* images,
                                             size_t count) noexcept
{
    return encode_frames(batch, params, images, count, nullptr, true);
}

} // extern "C"

namespace {

charls_jpegls_errc decode_frames(charlsx_batch* batch, const charlsx_batch_params* params, charlsx_batch_image* images, size_t count,
                                 void* cuda_stream, bool host_memory) noexcept
{
    return guarded([&] {
        check_pointer(batch);
        const charlsx_batch_params& bp = *check_pointer(params);
        check_argument(images != nullptr || count == 0);
        validate_params(bp);
        const size_t stride = effective_stride(bp);
        const charls_frame_info& f = bp.frame_info;
        auto* stream = static_cast<CUstream_st*>(cuda_stream);

        batch->frames.resize(count);
        for (size_t i = 0; i < count; ++i)
        {
            check_argument(images[i].pixels != nullptr && images[i].stream != nullptr);
            batch->frames[i] = BatchFrame{static_cast<uint8_t*>(images[i].pixels), static_cast<uint8_t*>(images[i].stream),
                                          images[i].stream_capacity, 0, 0, 0};
        }
        if (count == 0)
            return;

        // headers are parsed on the host: one gather kernel + one copy brings the first bytes of every stream over
        // (streams in host memory are read where they are)
        // with side tables of interval offsets in the headers, SOS sits behind them: read that much more of every stream
        const uint32_t prefix_bytes =
            header_prefix_bytes + ((bp.flags & CHARLSX_BATCH_OFFSET_TABLE) != 0 ? static_cast<uint32_t>(offset_table_bytes(f.height)) : 0U);
        const uint8_t* prefixes = nullptr; // in the engine's pinned staging buffer, valid until its next call
        if (!host_memory)
            check_status(batch->engine.download_prefixes(batch->frames.data(), count, prefix_bytes, prefixes, stream));

        bool have_params = false;
        CodecParams p{};
        int32_t first_error = 0;
        std::vector<size_t> active; // frames that go to the GPU
        for (size_t i = 0; i < count; ++i)
        {
            BatchFrame& frame = batch->frames[i];
            const size_t available = frame.stream_capacity < prefix_bytes ? frame.stream_capacity : prefix_bytes;
            const charls_jpegls_errc errc = guarded([&] {
                StreamReader reader;
                reader.source(host_memory ? frame.stream : prefixes + i * prefix_bytes, available);
                reader.read_header();
                const charls_frame_info& info = reader.frame_info();
                // every frame must match the batch description
                if (info.width != f.width || info.height != f.height || info.bits_per_sample != f.bits_per_sample ||
                    info.component_count != f.component_count || reader.scan_component_count() != static_cast<uint32_t>(f.component_count) ||
                    reader.scan_near_lossless() != bp.near_lossless || reader.scan_interleave_mode() != bp.interleave_mode ||
                    reader.color_transformation() != bp.color_transformation)
                    fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT);
                const charls_jpegls_pc_parameters pc = reader.validated_preset_coding_parameters();
                const PresetCodingParameters preset{pc.maximum_sample_value, pc.threshold1, pc.threshold2, pc.threshold3, pc.reset_value};
                const CodecParams q = make_codec_params(static_cast<int32_t>(info.width), static_cast<int32_t>(info.height),
                                                        info.bits_per_sample, info.component_count, bp.near_lossless, bp.interleave_mode,
                                                        bp.interleave_mode != 0 ? bp.color_transformation : 0, preset,
                                                        reader.restart_interval());
                if (!have_params)
                {
                    p = q;
                    have_params = true;
                }
                else if (q.t1 != p.t1 || q.t2 != p.t2 || q.t3 != p.t3 || q.reset != p.reset || q.restart_interval != p.restart_interval)
                {
                    fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT); // not a uniform batch
                }
                frame.scan_offset = reader.position();
                frame.table = StreamOffsetTable{};
                if (reader.scan_offset_table().total == q.interval_count + 1U)
                {
                    frame.table.total = reader.scan_offset_table().total;
                    for (uint32_t segment = 0; segment < offset_table_max_segments; ++segment)
                        frame.table.entry_offsets[segment] = reader.scan_offset_table().entry_offsets[segment];
                }
            });
            frame.status = errc;
            if (errc == 0)
                active.push_back(i);
            else if (first_error == 0)
                first_error = errc;
        }

        if (!active.empty())
        {
            std::vector<BatchFrame> work(active.size());
            for (size_t k = 0; k < active.size(); ++k)
                work[k] = batch->frames[active[k]];
            const int32_t status = host_memory ? batch->engine.decode_batch_host(p, work.data(), work.size(), stride)
                                               : batch->engine.decode_batch(p, work.data(), work.size(), stride, stream);
            for (size_t k = 0; k < active.size(); ++k)
                batch->frames[active[k]] = work[k];
            if (status != 0 && first_error == 0)
                first_error = status;
        }
        for (size_t i = 0; i < count; ++i)
        {
            images[i].status = batch->frames[i].status;
            images[i].stream_size = batch->frames[i].stream_size;
        }
        check_status(first_error);
    });
}

} // namespace

extern "C" {

charls_jpegls_errc charlsx_batch_decode(charlsx_batch* batch, const charlsx_batch_params* params, charlsx_batch_image* images,
                                        size_t count, void* cuda_stream) noexcept
{
    return decode_frames(batch, params, images, count, cuda_stream, false);
}

charls_jpegls_errc charlsx_batch_decode_host(charlsx_batch* batch, const charlsx_batch_params* params, charlsx_batch_image* images,
                                             size_t count) noexcept
{
    return decode_frames(batch, params, images, count, nullptr, true);
}

charls_jpegls_errc charlsx_batch_get_last_coder_kernel_ms(const charlsx_batch* batch, float* milliseconds) noexcept
{
    return guarded([&] { *check_pointer(milliseconds) = check_pointer(batch)->engine.last_coder_kernel_ms(); });
}

charls_jpegls_errc charlsx_batch_get_last_kernel_launches(const charlsx_batch* batch, uint32_t* launches) noexcept
{
    return guarded([&] { *check_pointer(launches) = check_pointer(batch)->engine.last_kernel_launches(); });
}

} // extern "C"
