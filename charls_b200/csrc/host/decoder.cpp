// decoder.cpp -- the charls_jpegls_decoder_* half of the C ABI.
//
// State machine and validation follow the reference's charls_jpegls_decoder (src/charls_jpegls_decoder.cpp:21-273);
// each scan is decoded by jls::Engine (CUDA) where the reference calls make_scan_codec<scan_decoder>()->decode_scan()
// (src/charls_jpegls_decoder.cpp:186-189).
#include "../engine.hpp"
#include "abi_support.hpp"
#include "stream_reader.hpp"

#include <cstdlib>

using namespace jls;
using namespace jls::host;

struct charls_jpegls_decoder final
{
    ~charls_jpegls_decoder()
    {
        if (deferred_)
        {
            size_t ignored = 0;
            (void)engine().decode_scan_to_host_end(ignored); // the copy into the caller's buffer may still be running
        }
        Engine::release(engine_);
    }
    Engine& engine()
    {
        if (!engine_)
            engine_ = Engine::acquire();
        return *engine_;
    }

    enum class State
    {
        initial,
        source_set,
        spiff_header_read,
        spiff_header_not_found,
        header_read,
        completed
    };

    void source(const uint8_t* data, size_t size)
    {
        check_buffer(data, size);
        check_operation(state_ == State::initial);
        reader_.source(data, size);
        state_ = State::source_set;
    }

    bool read_spiff_header(charls_spiff_header* header)
    {
        check_operation(state_ == State::source_set);
        bool found = false;
        reader_.read_header(header, &found);
        state_ = found ? State::spiff_header_read : State::spiff_header_not_found;
        return found;
    }

    void read_header()
    {
        check_operation(state_ >= State::source_set && state_ < State::header_read);
        if (state_ != State::spiff_header_not_found)
            reader_.read_header();
        state_ = reader_.end_of_image() ? State::completed : State::header_read;
    }

    void check_header_read() const { check_operation(state_ >= State::header_read); }
    void check_completed() const { check_operation(state_ == State::completed); }

    charls_frame_info frame_info() const
    {
        check_header_read();
        return reader_.frame_info();
    }

    int32_t near_lossless(int32_t component) const
    {
        check_header_read();
        check_argument(static_cast<size_t>(component) < reader_.component_count());
        return reader_.near_lossless(static_cast<size_t>(component));
    }

    int32_t interleave_mode(int32_t component) const
    {
        check_header_read();
        check_argument(static_cast<size_t>(component) < reader_.component_count());
        return reader_.interleave_mode(static_cast<size_t>(component));
    }

    int32_t color_transformation() const
    {
        check_header_read();
        return reader_.color_transformation();
    }

    uint32_t restart_interval() const
    {
        check_header_read();
        return reader_.restart_interval();
    }

    charls_jpegls_pc_parameters preset_coding_parameters() const
    {
        check_header_read();
        return reader_.preset_coding_parameters();
    }

    // reference src/charls_jpegls_decoder.cpp:93-121
    size_t destination_size(size_t stride) const
    {
        const charls_frame_info info = frame_info();
        const size_t sample_bytes = static_cast<size_t>((info.bits_per_sample + 7) / 8);
        if (stride == 0)
            return checked_mul(checked_mul(checked_mul(static_cast<size_t>(info.component_count), info.height), info.width),
                               sample_bytes);
        if (interleave_mode(0) == 0)
        {
            const size_t minimum_stride = static_cast<size_t>(info.width) * sample_bytes;
            check_argument(stride >= minimum_stride, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_STRIDE);
            return checked_mul(checked_mul(stride, static_cast<size_t>(info.component_count)), info.height) - (stride - minimum_stride);
        }
        const size_t minimum_stride = static_cast<size_t>(info.width) * static_cast<size_t>(info.component_count) * sample_bytes;
        check_argument(stride >= minimum_stride, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_STRIDE);
        return checked_mul(stride, info.height) - (stride - minimum_stride);
    }

    // reference src/charls_jpegls_decoder.cpp:177-201
    // `deferred` (charlsx_jpegls_decoder_decode_to_buffer_begin): a frame that is a single scan is only issued here and
    // decode_end() completes it; frames of several scans need the host between the scans and run to the end here.
    void decode(uint8_t* destination, size_t destination_size, size_t stride, bool deferred = false)
    {
        check_buffer(destination, destination_size);
        check_operation(state_ == State::header_read && !deferred_);
        const charls_frame_info& info = reader_.frame_info();
        (void)check_stride_and_destination_size(destination_size, stride); // argument errors come before any device work
        check_status(engine().upload_stream(reader_.source_data(), reader_.source_size()));

        for (size_t component = 0;;)
        {
            const size_t scan_stride = check_stride_and_destination_size(destination_size, stride);
            const charls_jpegls_pc_parameters pc = reader_.validated_preset_coding_parameters();
            const PresetCodingParameters preset{pc.maximum_sample_value, pc.threshold1, pc.threshold2, pc.threshold3, pc.reset_value};
            const int32_t ilv = reader_.scan_interleave_mode();
            const CodecParams p = make_codec_params(static_cast<int32_t>(info.width), static_cast<int32_t>(info.height),
                                                    info.bits_per_sample, static_cast<int32_t>(reader_.scan_component_count()),
                                                    reader_.scan_near_lossless(), ilv, ilv != 0 ? reader_.color_transformation() : 0,
                                                    preset, reader_.restart_interval());
            // a side table of interval offsets in front of this scan (jls_common.h) spares the search for restart markers
            StreamOffsetTable table{};
            static const bool ignore_tables = [] {
                const char* value = std::getenv("CHARLS_B200_IGNORE_OFFSET_TABLES");
                return value && value[0] == '1';
            }();
            if (!ignore_tables && reader_.scan_offset_table().total == p.interval_count + 1U)
            {
                table.total = reader_.scan_offset_table().total;
                for (uint32_t segment = 0; segment < offset_table_max_segments; ++segment)
                    table.entry_offsets[segment] = reader_.scan_offset_table().entry_offsets[segment];
            }
            const StreamOffsetTable* use_table = table.total != 0 ? &table : nullptr;
            if (deferred && component == 0 && reader_.scan_component_count() == reader_.component_count())
            {
                check_status(engine().decode_scan_to_host_begin(p, reader_.position(), destination, scan_stride, use_table));
                deferred_ = true;
                return;
            }
            size_t consumed = 0;
            check_status(engine().decode_scan_to_host(p, reader_.position(), destination, scan_stride, consumed, use_table));
            reader_.advance(consumed);

            component += reader_.scan_component_count();
            if (component == reader_.component_count())
                break;
            destination += scan_stride * info.height;
            destination_size -= scan_stride * info.height;
            reader_.read_next_start_of_scan();
        }
        reader_.read_end_of_image();
        state_ = State::completed;
    }

    // second half of a deferred decode; a no-op when the first half ran to the end by itself
    void decode_end()
    {
        if (!deferred_)
        {
            check_operation(state_ == State::completed);
            return;
        }
        deferred_ = false;
        size_t consumed = 0;
        check_status(engine().decode_scan_to_host_end(consumed));
        reader_.advance(consumed);
        reader_.read_end_of_image();
        state_ = State::completed;
    }

    StreamReader& reader() noexcept { return reader_; }
    const StreamReader& reader() const noexcept { return reader_; }

private:
    // reference src/charls_jpegls_decoder.cpp:211-245
    size_t check_stride_and_destination_size(size_t destination_size, size_t stride) const
    {
        const charls_frame_info& info = reader_.frame_info();
        const size_t components_in_plane = reader_.scan_interleave_mode() == 0 ? 1U : reader_.scan_component_count();
        const size_t minimum_stride = components_in_plane * info.width * static_cast<size_t>((info.bits_per_sample + 7) / 8);
        if (stride == 0)
            stride = minimum_stride;
        else if (stride < minimum_stride)
            fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_STRIDE);
        const size_t unused_at_end = stride - minimum_stride;
        const size_t minimum_size = reader_.scan_interleave_mode() == 0
                                        ? stride * reader_.scan_component_count() * info.height - unused_at_end
                                        : stride * info.height - unused_at_end;
        if (destination_size < minimum_size)
            fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_SIZE);
        return stride;
    }

    State state_{State::initial};
    StreamReader reader_;
    Engine* engine_{}; // borrowed from the pool on first use
    bool deferred_{};  // decode_to_buffer_begin has issued the scan, decode_end has not been called yet
};

extern "C" {

charls_jpegls_decoder* charls_jpegls_decoder_create(void) noexcept
{
    return new (std::nothrow) charls_jpegls_decoder;
}

void charls_jpegls_decoder_destroy(const charls_jpegls_decoder* decoder) noexcept
{
    delete decoder;
}

charls_jpegls_errc charls_jpegls_decoder_set_source_buffer(charls_jpegls_decoder* decoder, const void* source_buffer,
                                                           size_t source_size_bytes) noexcept
{
    return guarded([&] { check_pointer(decoder)->source(static_cast<const uint8_t*>(source_buffer), source_size_bytes); });
}

charls_jpegls_errc charls_jpegls_decoder_read_spiff_header(charls_jpegls_decoder* decoder, charls_spiff_header* spiff_header,
                                                           int32_t* header_found) noexcept
{
    return guarded([&] {
        check_pointer(header_found);
        *header_found = check_pointer(decoder)->read_spiff_header(check_pointer(spiff_header)) ? 1 : 0;
    });
}

charls_jpegls_errc charls_jpegls_decoder_read_header(charls_jpegls_decoder* decoder) noexcept
{
    return guarded([&] { check_pointer(decoder)->read_header(); });
}

charls_jpegls_errc charls_jpegls_decoder_get_frame_info(const charls_jpegls_decoder* decoder, charls_frame_info* frame_info) noexcept
{
    return guarded([&] { *check_pointer(frame_info) = check_pointer(decoder)->frame_info(); });
}

charls_jpegls_errc charls_jpegls_decoder_get_near_lossless(const charls_jpegls_decoder* decoder, int32_t component_index,
                                                           int32_t* near_lossless) noexcept
{
    return guarded([&] { *check_pointer(near_lossless) = check_pointer(decoder)->near_lossless(component_index); });
}

charls_jpegls_errc charls_jpegls_decoder_get_interleave_mode(const charls_jpegls_decoder* decoder, int32_t component_index,
                                                             charls_interleave_mode* interleave_mode) noexcept
{
    return guarded([&] { *check_pointer(interleave_mode) = check_pointer(decoder)->interleave_mode(component_index); });
}

charls_jpegls_errc charls_jpegls_decoder_get_preset_coding_parameters(const charls_jpegls_decoder* decoder, int32_t /*reserved*/,
                                                                      charls_jpegls_pc_parameters* preset_coding_parameters) noexcept
{
    return guarded([&] { *check_pointer(preset_coding_parameters) = check_pointer(decoder)->preset_coding_parameters(); });
}

charls_jpegls_errc charls_jpegls_decoder_get_color_transformation(const charls_jpegls_decoder* decoder,
                                                                  charls_color_transformation* color_transformation) noexcept
{
    return guarded([&] { *check_pointer(color_transformation) = check_pointer(decoder)->color_transformation(); });
}

charls_jpegls_errc charls_jpegls_decoder_get_destination_size(const charls_jpegls_decoder* decoder, uint32_t stride,
                                                              size_t* destination_size_bytes) noexcept
{
    return guarded([&] { *check_pointer(destination_size_bytes) = check_pointer(decoder)->destination_size(stride); });
}

charls_jpegls_errc charls_jpegls_decoder_decode_to_buffer(charls_jpegls_decoder* decoder, void* destination_buffer,
                                                          size_t destination_size_bytes, uint32_t stride) noexcept
{
    return guarded([&] { check_pointer(decoder)->decode(static_cast<uint8_t*>(destination_buffer), destination_size_bytes, stride); });
}

charls_jpegls_errc charls_jpegls_decoder_at_comment(charls_jpegls_decoder* decoder, charls_at_comment_handler handler,
                                                    void* user_context) noexcept
{
    return guarded([&] { check_pointer(decoder)->reader().at_comment(handler, user_context); });
}

charls_jpegls_errc charls_jpegls_decoder_at_application_data(charls_jpegls_decoder* decoder,
                                                             charls_at_application_data_handler handler, void* user_context) noexcept
{
    return guarded([&] { check_pointer(decoder)->reader().at_application_data(handler, user_context); });
}

charls_jpegls_errc charls_decoder_get_compressed_data_format(const charls_jpegls_decoder* decoder,
                                                             charls_compressed_data_format* compressed_data_format) noexcept
{
    return guarded([&] { *check_pointer(compressed_data_format) = check_pointer(decoder)->reader().compressed_data_format(); });
}

charls_jpegls_errc charls_decoder_get_mapping_table_id(const charls_jpegls_decoder* decoder, int32_t component_index,
                                                       int32_t* table_id) noexcept
{
    return guarded([&] {
        check_pointer(decoder)->check_completed();
        check_argument(static_cast<size_t>(component_index) < decoder->reader().component_count());
        *check_pointer(table_id) = decoder->reader().mapping_table_id(static_cast<size_t>(component_index));
    });
}

charls_jpegls_errc charls_decoder_find_mapping_table_index(const charls_jpegls_decoder* decoder, int32_t mapping_table_id,
                                                           int32_t* index) noexcept
{
    return guarded([&] {
        check_pointer(decoder)->check_completed();
        check_range(1, 255, mapping_table_id);
        *check_pointer(index) = decoder->reader().find_mapping_table_index(static_cast<uint8_t>(mapping_table_id));
    });
}

charls_jpegls_errc charls_decoder_get_mapping_table_count(const charls_jpegls_decoder* decoder, int32_t* count) noexcept
{
    return guarded([&] {
        check_pointer(decoder)->check_completed();
        *check_pointer(count) = static_cast<int32_t>(decoder->reader().mapping_table_count());
    });
}

charls_jpegls_errc charls_decoder_get_mapping_table_info(const charls_jpegls_decoder* decoder, int32_t mapping_table_index,
                                                         charls_mapping_table_info* mapping_table_info) noexcept
{
    return guarded([&] {
        check_pointer(decoder)->check_completed();
        check_argument(static_cast<size_t>(mapping_table_index) < decoder->reader().mapping_table_count());
        *check_pointer(mapping_table_info) = decoder->reader().mapping_table_info(static_cast<size_t>(mapping_table_index));
    });
}

charls_jpegls_errc charls_decoder_get_mapping_table_data(const charls_jpegls_decoder* decoder, int32_t mapping_table_index,
                                                         void* mapping_table_data, size_t mapping_table_size_bytes) noexcept
{
    return guarded([&] {
        check_pointer(decoder)->check_completed();
        check_argument(static_cast<size_t>(mapping_table_index) < decoder->reader().mapping_table_count());
        check_buffer(mapping_table_data, mapping_table_size_bytes);
        decoder->reader().mapping_table_data(static_cast<size_t>(mapping_table_index), static_cast<uint8_t*>(mapping_table_data),
                                             mapping_table_size_bytes);
    });
}

charls_jpegls_errc charlsx_jpegls_decoder_decode_to_buffer_begin(charls_jpegls_decoder* decoder, void* destination_buffer,
                                                                 size_t destination_size_bytes, uint32_t stride) noexcept
{
    return guarded(
        [&] { check_pointer(decoder)->decode(static_cast<uint8_t*>(destination_buffer), destination_size_bytes, stride, true); });
}

charls_jpegls_errc charlsx_jpegls_decoder_decode_end(charls_jpegls_decoder* decoder) noexcept
{
    return guarded([&] { check_pointer(decoder)->decode_end(); });
}

charls_jpegls_errc charlsx_jpegls_decoder_get_restart_interval(const charls_jpegls_decoder* decoder, uint32_t* restart_interval) noexcept
{
    return guarded([&] { *check_pointer(restart_interval) = check_pointer(decoder)->restart_interval(); });
}

} // extern "C"
