// encoder.cpp -- the charls_jpegls_encoder_* half of the C ABI.
//
// State machine, validation order and error codes follow the reference's charls_jpegls_encoder
// (src/charls_jpegls_encoder.cpp:31-444); the scan itself is handed to jls::Engine (CUDA) where the reference calls
// make_scan_codec<scan_encoder>()->encode_scan() (src/charls_jpegls_encoder.cpp:285-296).
#include "../engine.hpp"
#include "abi_support.hpp"
#include "stream_writer.hpp"

#include <cstdlib>
#include <cstring>

using namespace jls;
using namespace jls::host;

struct charls_jpegls_encoder final
{
    ~charls_jpegls_encoder()
    {
        if (deferred_components_ != 0)
        {
            size_t ignored = 0;
            (void)engine().encode_scan_from_host_end(ignored); // nothing of this object may stay in flight
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
        destination_set,
        spiff_header,
        tables_and_miscellaneous,
        completed
    };

    void destination(uint8_t* data, size_t size)
    {
        check_buffer(data, size);
        check_operation(state_ <= State::destination_set);
        writer_.destination(data, size);
        state_ = State::destination_set;
    }

    void frame_info(const charls_frame_info& info)
    {
        check_range(1U, maximum_width, info.width, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_WIDTH);
        check_range(1U, maximum_height, info.height, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_HEIGHT);
        check_range(2, 16, info.bits_per_sample, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_BITS_PER_SAMPLE);
        check_range(1, maximum_component_count, info.component_count, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_COMPONENT_COUNT);
        frame_info_ = info;
    }

    void interleave_mode(int32_t mode)
    {
        check_range(0, 2, mode, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_INTERLEAVE_MODE);
        interleave_mode_ = mode;
    }

    void near_lossless(int32_t near)
    {
        check_range(0, 255, near, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_NEAR_LOSSLESS);
        near_lossless_ = near;
    }

    void encoding_options(uint32_t options)
    {
        check_argument(options <= 7U, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_ENCODING_OPTIONS);
        encoding_options_ = options;
    }

    void color_transformation(int32_t transformation)
    {
        check_range(0, 3, transformation, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_COLOR_TRANSFORMATION);
        color_transformation_ = transformation;
    }

    void mapping_table_id(int32_t component_index, int32_t table_id)
    {
        check_range(0, maximum_component_count - 1, component_index);
        check_range(0, 255, table_id);
        writer_.set_mapping_table_id(static_cast<size_t>(component_index), table_id);
    }

    void offset_table(bool enabled)
    {
        check_operation(encoded_component_count_ == 0 && state_ != State::completed);
        offset_table_ = enabled;
    }

    void restart_interval(uint32_t interval)
    {
        check_operation(encoded_component_count_ == 0 && state_ != State::completed);
        restart_interval_ = interval;
    }

    // reference src/charls_jpegls_encoder.cpp:104-114, plus what restart markers can add per interval and component: the
    // first sample after a restart is predicted from 0 whatever the image looks like and may take a whole LIMIT-bit code
    // word (a constant 10000 x 1 image costs 6 bytes per line, not a fraction of a bit), then 2 bytes RSTm, 1 byte of
    // padding and 1 stuffed byte after a trailing 0xFF; and the DRI segment.
    size_t estimated_destination_size() const
    {
        check_operation(frame_info_.width != 0);
        size_t size = checked_mul(checked_mul(checked_mul(frame_info_.width, frame_info_.height),
                                              static_cast<size_t>(frame_info_.component_count)),
                                  static_cast<size_t>((frame_info_.bits_per_sample + 7) / 8));
        size = add_saturated(size, size / 16U + 1024 + spiff_header_size_in_bytes);
        if (restart_interval_ != 0)
        {
            const size_t intervals = (frame_info_.height + restart_interval_ - 1) / restart_interval_;
            const size_t bits = static_cast<size_t>(frame_info_.bits_per_sample < 2 ? 2 : frame_info_.bits_per_sample);
            const size_t limit_bits = 2 * (bits + (bits < 8 ? 8 : bits)); // LIMIT of T.87 (reference src/default_traits.hpp:41-45)
            const size_t per_interval = (limit_bits + 7) / 8 + 4;
            size = add_saturated(size, checked_mul(checked_mul(intervals, per_interval), static_cast<size_t>(frame_info_.component_count)) + 8U);
            if (offset_table_)
                size = add_saturated(size, checked_mul(offset_table_bytes(static_cast<uint32_t>(intervals)),
                                                       static_cast<size_t>(frame_info_.component_count)));
        }
        return size;
    }

    void write_spiff_header(const charls_spiff_header& header)
    {
        check_range(1U, maximum_height, header.height, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_HEIGHT);
        check_range(1U, maximum_width, header.width, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_WIDTH);
        write_spiff_header_core(header);
    }

    void write_standard_spiff_header(int32_t color_space, int32_t resolution_units, uint32_t vertical, uint32_t horizontal)
    {
        check_operation(frame_info_.width != 0);
        write_spiff_header_core({0, frame_info_.component_count, frame_info_.height, frame_info_.width, color_space,
                                 frame_info_.bits_per_sample, 6 /* JPEG-LS */, resolution_units, vertical, horizontal});
    }

    void write_spiff_entry(uint32_t tag, const uint8_t* data, size_t size)
    {
        check_buffer(data, size);
        check_argument(tag != 1); // the end-of-directory tag has its own call
        check_argument(size <= spiff_entry_max_data_size, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_SIZE);
        check_operation(state_ == State::spiff_header);
        writer_.write_spiff_directory_entry(tag, data, size);
    }

    void write_spiff_end_of_directory()
    {
        check_operation(state_ == State::spiff_header);
        enter_tables_and_miscellaneous();
    }

    void write_comment(const uint8_t* data, size_t size)
    {
        check_buffer(data, size);
        check_argument(size <= segment_max_data_size, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_SIZE);
        check_can_write();
        enter_tables_and_miscellaneous();
        writer_.write_comment(data, size);
    }

    void write_application_data(int32_t id, const uint8_t* data, size_t size)
    {
        check_range(0, 15, id);
        check_buffer(data, size);
        check_argument(size <= segment_max_data_size, CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_SIZE);
        check_can_write();
        enter_tables_and_miscellaneous();
        writer_.write_application_data(id, data, size);
    }

    void write_mapping_table(int32_t table_id, int32_t entry_size, const uint8_t* data, size_t size)
    {
        check_range(1, 255, table_id);
        check_range(1, 255, entry_size);
        check_buffer(data, size);
        check_argument(size >= static_cast<size_t>(entry_size), CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_SIZE);
        check_can_write();
        enter_tables_and_miscellaneous();
        writer_.write_mapping_table(table_id, entry_size, data, size);
    }

    // reference src/charls_jpegls_encoder.cpp:187-236
    // `deferred` (charlsx_jpegls_encoder_encode_from_buffer_begin): a call that makes exactly one scan only issues it;
    // encode_end() completes it.  Calls that make several scans (planar components) run to the end here.
    void encode_components(const uint8_t* source, size_t source_size, int32_t source_component_count, size_t stride,
                           bool deferred = false)
    {
        check_buffer(source, source_size);
        check_can_write();
        check_operation(frame_info_.width != 0);
        if (frame_info_.component_count == 1 && interleave_mode_ != 0)
            fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_INTERLEAVE_MODE);
        if (source_component_count < 1 || source_component_count > frame_info_.component_count - encoded_component_count_ ||
            (interleave_mode_ != 0 && (source_component_count < 2 || source_component_count > maximum_component_count_in_scan)))
            fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_INTERLEAVE_MODE); // the reference leaves these unchecked (UB)
        const int32_t maximum_value = maximum_bit_sample_value(frame_info_.bits_per_sample);
        if (near_lossless_ > maximum_near_lossless(effective_maximum_sample_value(maximum_value)))
            fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_NEAR_LOSSLESS);
        const size_t scan_stride = check_stride_and_source_size(source_size, stride, source_component_count);

        const PresetCodingParameters user{user_preset_.maximum_sample_value, user_preset_.threshold1, user_preset_.threshold2,
                                          user_preset_.threshold3, user_preset_.reset_value};
        if (!validate_preset_parameters(user, maximum_value, near_lossless_, &preset_))
            fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_JPEGLS_PC_PARAMETERS);

        if (encoded_component_count_ == 0)
        {
            enter_tables_and_miscellaneous();
            if (color_transformation_ != 0)
            {
                if (!(frame_info_.component_count == 3 && (frame_info_.bits_per_sample == 8 || frame_info_.bits_per_sample == 16) &&
                      near_lossless_ == 0 && interleave_mode_ != 0))
                    fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_COLOR_TRANSFORMATION);
                writer_.write_color_transform(color_transformation_);
            }
            if (writer_.write_start_of_frame(frame_info_))
                writer_.write_oversize_dimensions(frame_info_.height, frame_info_.width);
            if (!is_default_preset(user, default_preset_parameters(maximum_value, near_lossless_)) ||
                ((encoding_options_ & 4U) != 0 && frame_info_.bits_per_sample > 12))
                writer_.write_preset_coding_parameters({preset_.maximum_sample_value, preset_.threshold1, preset_.threshold2,
                                                        preset_.threshold3, preset_.reset_value});
            // The restart interval applies to every scan that follows (ISO/IEC 14495-1 C.2.5); not in the reference.
            if (restart_interval_ != 0)
                writer_.write_define_restart_interval(restart_interval_);
        }

        if (interleave_mode_ == 0)
        {
            const size_t plane_bytes = scan_stride * frame_info_.height;
            if (deferred && source_component_count == 1)
            {
                reserve_offset_table();
                writer_.write_start_of_scan(1, near_lossless_, interleave_mode_);
                begin_scan(source, scan_stride, 1);
                deferred_components_ = 1;
                return;
            }
            for (int32_t component = 0; component < source_component_count; ++component)
            {
                reserve_offset_table();
                writer_.write_start_of_scan(1, near_lossless_, interleave_mode_);
                encode_scan(source + static_cast<size_t>(component) * plane_bytes, scan_stride, 1);
            }
        }
        else
        {
            reserve_offset_table();
            writer_.write_start_of_scan(source_component_count, near_lossless_, interleave_mode_);
            if (deferred)
            {
                begin_scan(source, scan_stride, source_component_count);
                deferred_components_ = source_component_count;
                return;
            }
            encode_scan(source, scan_stride, source_component_count);
        }

        encoded_component_count_ += source_component_count;
        if (encoded_component_count_ == frame_info_.component_count)
            write_end_of_image();
    }

    // second half of a deferred encode; a no-op when the first half ran to the end by itself
    void encode_end()
    {
        if (deferred_components_ == 0)
            return;
        const int32_t components = deferred_components_;
        deferred_components_ = 0;
        size_t written = 0;
        check_status(engine().encode_scan_from_host_end(written));
        writer_.advance(written);
        encoded_component_count_ += components;
        if (encoded_component_count_ == frame_info_.component_count)
            write_end_of_image();
    }

    void create_abbreviated_format()
    {
        check_operation(state_ == State::tables_and_miscellaneous);
        write_end_of_image();
    }

    void rewind()
    {
        check_operation(deferred_components_ == 0);
        if (state_ == State::initial)
            return;
        writer_.rewind();
        state_ = State::destination_set;
        encoded_component_count_ = 0;
    }

    size_t bytes_written() const noexcept { return writer_.bytes_written(); }
    const charls_frame_info& frame() const noexcept { return frame_info_; }
    void user_preset(const charls_jpegls_pc_parameters& pc) noexcept { user_preset_ = pc; }

private:
    void check_can_write() const
    {
        check_operation(state_ >= State::destination_set && state_ < State::completed && deferred_components_ == 0);
    }

    int32_t effective_maximum_sample_value(int32_t maximum_value) const
    {
        if (user_preset_.maximum_sample_value != 0)
        {
            if (user_preset_.maximum_sample_value < 1 || user_preset_.maximum_sample_value > maximum_value)
                fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_JPEGLS_PC_PARAMETERS);
            return user_preset_.maximum_sample_value;
        }
        return maximum_value;
    }

    // reference src/charls_jpegls_encoder.cpp:299-331
    size_t check_stride_and_source_size(size_t source_size, size_t stride, int32_t source_component_count) const
    {
        size_t minimum_stride = static_cast<size_t>(frame_info_.width) * static_cast<size_t>((frame_info_.bits_per_sample + 7) / 8);
        if (interleave_mode_ != 0)
            minimum_stride *= static_cast<size_t>(source_component_count);
        if (stride == 0)
            stride = minimum_stride;
        else if (stride < minimum_stride)
            fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_STRIDE);
        const size_t unused_at_end = stride - minimum_stride;
        const size_t minimum_source_size =
            interleave_mode_ == 0 ? checked_mul(stride * static_cast<size_t>(source_component_count), frame_info_.height) - unused_at_end
                                  : checked_mul(stride, frame_info_.height) - unused_at_end;
        if (source_size < minimum_source_size)
            fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_SIZE);
        return stride;
    }

    void write_spiff_header_core(const charls_spiff_header& header)
    {
        check_operation(state_ == State::destination_set);
        writer_.write_start_of_image();
        writer_.write_spiff_header(header);
        state_ = State::spiff_header;
    }

    // reference src/charls_jpegls_encoder.cpp:365-389
    void enter_tables_and_miscellaneous()
    {
        if (state_ == State::tables_and_miscellaneous)
            return;
        if (state_ == State::spiff_header)
            writer_.write_spiff_end_of_directory();
        else
            writer_.write_start_of_image();
        if ((encoding_options_ & 2U) != 0)
        {
            static const char version[] = "charls 3.0.0"; // same identification string as the reference build
            writer_.write_comment(reinterpret_cast<const uint8_t*>(version), sizeof(version));
        }
        state_ = State::tables_and_miscellaneous;
    }

    void encode_scan(const uint8_t* source, size_t stride, int32_t component_count)
    {
        const CodecParams p = make_codec_params(static_cast<int32_t>(frame_info_.width), static_cast<int32_t>(frame_info_.height),
                                                frame_info_.bits_per_sample, component_count, near_lossless_, interleave_mode_,
                                                interleave_mode_ != 0 ? color_transformation_ : 0, preset_, restart_interval_);
        size_t written = 0;
        check_status(engine().encode_scan_from_host(p, source, stride, writer_.remaining_data(), writer_.remaining_size(), written,
                                                    scan_table_.total != 0 ? &scan_table_ : nullptr));
        writer_.advance(written);
    }

    // Side table of interval offsets for the scan whose SOS comes next (jls_common.h): the segments are written with empty
    // entries, the engine fills them in when the scan is coded.
    void reserve_offset_table()
    {
        scan_table_ = HostOffsetTable{};
        if (!offset_table_ || restart_interval_ == 0)
            return;
        const uint32_t lines = restart_interval_ < frame_info_.height ? restart_interval_ : frame_info_.height;
        const uint32_t intervals = (frame_info_.height + lines - 1) / lines;
        size_t positions[offset_table_max_segments] = {};
        writer_.write_offset_table_placeholder(intervals, positions);
        scan_table_.total = intervals + 1;
        for (uint32_t segment = 0; segment < offset_table_segment_count(scan_table_.total); ++segment)
            scan_table_.entries[segment] = writer_.data() + positions[segment];
    }

    void begin_scan(const uint8_t* source, size_t stride, int32_t component_count)
    {
        const CodecParams p = make_codec_params(static_cast<int32_t>(frame_info_.width), static_cast<int32_t>(frame_info_.height),
                                                frame_info_.bits_per_sample, component_count, near_lossless_, interleave_mode_,
                                                interleave_mode_ != 0 ? color_transformation_ : 0, preset_, restart_interval_);
        check_status(engine().encode_scan_from_host_begin(p, source, stride, writer_.remaining_data(), writer_.remaining_size(),
                                                          scan_table_.total != 0 ? &scan_table_ : nullptr));
    }

    void write_end_of_image()
    {
        writer_.write_end_of_image((encoding_options_ & 1U) != 0);
        state_ = State::completed;
    }

    charls_frame_info frame_info_{};
    int32_t near_lossless_{};
    int32_t encoded_component_count_{};
    int32_t interleave_mode_{};
    int32_t color_transformation_{};
    uint32_t encoding_options_{};
    // One line per restart interval: every line is an independent work item.  The reference writes none (and so does this
    // library with CHARLS_B200_RESTART_INTERVAL=0 or charlsx_jpegls_encoder_set_restart_interval(encoder, 0): byte-identical
    // output, one CUDA thread per scan).
    uint32_t restart_interval_{default_restart_interval()};
    static uint32_t default_restart_interval() noexcept
    {
        static const uint32_t value = [] {
            const char* text = std::getenv("CHARLS_B200_RESTART_INTERVAL");
            return text && text[0] >= '0' && text[0] <= '9' ? static_cast<uint32_t>(std::strtoul(text, nullptr, 10)) : 1U;
        }();
        return value;
    }
    State state_{State::initial};
    StreamWriter writer_;
    charls_jpegls_pc_parameters user_preset_{};
    PresetCodingParameters preset_{};
    Engine* engine_{}; // borrowed from the pool on first use
    int32_t deferred_components_{}; // components of the scan that encode_from_buffer_begin has issued and encode_end completes
    bool offset_table_{};           // write the side table of interval offsets (charlsx_jpegls_encoder_set_offset_table)
    HostOffsetTable scan_table_{};  // where the current scan's table entries go
};

extern "C" {

charls_jpegls_encoder* charls_jpegls_encoder_create(void) noexcept
{
    return new (std::nothrow) charls_jpegls_encoder;
}

void charls_jpegls_encoder_destroy(const charls_jpegls_encoder* encoder) noexcept
{
    delete encoder;
}

charls_jpegls_errc charls_jpegls_encoder_set_destination_buffer(charls_jpegls_encoder* encoder, void* destination_buffer,
                                                                size_t destination_size_bytes) noexcept
{
    return guarded([&] { check_pointer(encoder)->destination(static_cast<uint8_t*>(destination_buffer), destination_size_bytes); });
}

charls_jpegls_errc charls_jpegls_encoder_set_frame_info(charls_jpegls_encoder* encoder, const charls_frame_info* frame_info) noexcept
{
    return guarded([&] { check_pointer(encoder)->frame_info(*check_pointer(frame_info)); });
}

charls_jpegls_errc charls_jpegls_encoder_set_near_lossless(charls_jpegls_encoder* encoder, int32_t near_lossless) noexcept
{
    return guarded([&] { check_pointer(encoder)->near_lossless(near_lossless); });
}

charls_jpegls_errc charls_jpegls_encoder_set_encoding_options(charls_jpegls_encoder* encoder,
                                                              charls_encoding_options encoding_options) noexcept
{
    return guarded([&] { check_pointer(encoder)->encoding_options(encoding_options); });
}

charls_jpegls_errc charls_jpegls_encoder_set_interleave_mode(charls_jpegls_encoder* encoder,
                                                             charls_interleave_mode interleave_mode) noexcept
{
    return guarded([&] { check_pointer(encoder)->interleave_mode(interleave_mode); });
}

charls_jpegls_errc charls_jpegls_encoder_set_preset_coding_parameters(
    charls_jpegls_encoder* encoder, const charls_jpegls_pc_parameters* preset_coding_parameters) noexcept
{
    return guarded([&] { check_pointer(encoder)->user_preset(*check_pointer(preset_coding_parameters)); });
}

charls_jpegls_errc charls_jpegls_encoder_set_color_transformation(charls_jpegls_encoder* encoder,
                                                                  charls_color_transformation color_transformation) noexcept
{
    return guarded([&] { check_pointer(encoder)->color_transformation(color_transformation); });
}

charls_jpegls_errc charls_jpegls_encoder_set_mapping_table_id(charls_jpegls_encoder* encoder, int32_t component_index,
                                                              int32_t table_id) noexcept
{
    return guarded([&] { check_pointer(encoder)->mapping_table_id(component_index, table_id); });
}

charls_jpegls_errc charls_jpegls_encoder_get_estimated_destination_size(const charls_jpegls_encoder* encoder,
                                                                        size_t* size_in_bytes) noexcept
{
    return guarded([&] { *check_pointer(size_in_bytes) = check_pointer(encoder)->estimated_destination_size(); });
}

charls_jpegls_errc charls_jpegls_encoder_write_spiff_header(charls_jpegls_encoder* encoder,
                                                            const charls_spiff_header* spiff_header) noexcept
{
    return guarded([&] { check_pointer(encoder)->write_spiff_header(*check_pointer(spiff_header)); });
}

charls_jpegls_errc charls_jpegls_encoder_write_standard_spiff_header(charls_jpegls_encoder* encoder,
                                                                     charls_spiff_color_space color_space,
                                                                     charls_spiff_resolution_units resolution_units,
                                                                     uint32_t vertical_resolution,
                                                                     uint32_t horizontal_resolution) noexcept
{
    return guarded([&] {
        check_pointer(encoder)->write_standard_spiff_header(color_space, resolution_units, vertical_resolution, horizontal_resolution);
    });
}

charls_jpegls_errc charls_jpegls_encoder_write_spiff_entry(charls_jpegls_encoder* encoder, uint32_t entry_tag,
                                                           const void* entry_data, size_t entry_data_size_bytes) noexcept
{
    return guarded(
        [&] { check_pointer(encoder)->write_spiff_entry(entry_tag, static_cast<const uint8_t*>(entry_data), entry_data_size_bytes); });
}

charls_jpegls_errc charls_jpegls_encoder_write_spiff_end_of_directory_entry(charls_jpegls_encoder* encoder) noexcept
{
    return guarded([&] { check_pointer(encoder)->write_spiff_end_of_directory(); });
}

charls_jpegls_errc charls_jpegls_encoder_write_comment(charls_jpegls_encoder* encoder, const void* comment,
                                                       size_t comment_size_bytes) noexcept
{
    return guarded([&] { check_pointer(encoder)->write_comment(static_cast<const uint8_t*>(comment), comment_size_bytes); });
}

charls_jpegls_errc charls_jpegls_encoder_write_application_data(charls_jpegls_encoder* encoder, int32_t application_data_id,
                                                                const void* application_data,
                                                                size_t application_data_size_bytes) noexcept
{
    return guarded([&] {
        check_pointer(encoder)->write_application_data(application_data_id, static_cast<const uint8_t*>(application_data),
                                                       application_data_size_bytes);
    });
}

charls_jpegls_errc charls_jpegls_encoder_write_mapping_table(charls_jpegls_encoder* encoder, int32_t table_id, int32_t entry_size,
                                                             const void* table_data, size_t table_data_size_bytes) noexcept
{
    return guarded([&] {
        check_pointer(encoder)->write_mapping_table(table_id, entry_size, static_cast<const uint8_t*>(table_data),
                                                    table_data_size_bytes);
    });
}

charls_jpegls_errc charls_jpegls_encoder_encode_from_buffer(charls_jpegls_encoder* encoder, const void* source_buffer,
                                                            size_t source_size_bytes, uint32_t stride) noexcept
{
    return guarded([&] {
        check_pointer(encoder)->encode_components(static_cast<const uint8_t*>(source_buffer), source_size_bytes,
                                                  encoder->frame().component_count, stride);
    });
}

charls_jpegls_errc charls_jpegls_encoder_encode_components_from_buffer(charls_jpegls_encoder* encoder, const void* source_buffer,
                                                                       size_t source_size_bytes, int32_t source_component_count,
                                                                       uint32_t stride) noexcept
{
    return guarded([&] {
        check_pointer(encoder)->encode_components(static_cast<const uint8_t*>(source_buffer), source_size_bytes,
                                                  source_component_count, stride);
    });
}

charls_jpegls_errc charls_jpegls_encoder_create_abbreviated_format(charls_jpegls_encoder* encoder) noexcept
{
    return guarded([&] { check_pointer(encoder)->create_abbreviated_format(); });
}

charls_jpegls_errc charls_jpegls_encoder_get_bytes_written(const charls_jpegls_encoder* encoder, size_t* bytes_written) noexcept
{
    return guarded([&] { *check_pointer(bytes_written) = check_pointer(encoder)->bytes_written(); });
}

charls_jpegls_errc charls_jpegls_encoder_rewind(charls_jpegls_encoder* encoder) noexcept
{
    return guarded([&] { check_pointer(encoder)->rewind(); });
}

charls_jpegls_errc charlsx_jpegls_encoder_set_restart_interval(charls_jpegls_encoder* encoder, uint32_t restart_interval) noexcept
{
    return guarded([&] { check_pointer(encoder)->restart_interval(restart_interval); });
}

charls_jpegls_errc charlsx_jpegls_encoder_set_offset_table(charls_jpegls_encoder* encoder, int32_t enabled) noexcept
{
    return guarded([&] { check_pointer(encoder)->offset_table(enabled != 0); });
}

charls_jpegls_errc charlsx_jpegls_encoder_encode_from_buffer_begin(charls_jpegls_encoder* encoder, const void* source_buffer,
                                                                   size_t source_size_bytes, uint32_t stride) noexcept
{
    return guarded([&] {
        check_pointer(encoder)->encode_components(static_cast<const uint8_t*>(source_buffer), source_size_bytes,
                                                  encoder->frame().component_count, stride, true);
    });
}

charls_jpegls_errc charlsx_jpegls_encoder_encode_end(charls_jpegls_encoder* encoder) noexcept
{
    return guarded([&] { check_pointer(encoder)->encode_end(); });
}

} // extern "C"
