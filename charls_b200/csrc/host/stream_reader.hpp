// stream_reader.hpp -- parses the JPEG-LS container (marker segments) on the host.
//
// Host-side counterpart of the reference's jpeg_stream_reader (src/jpeg_stream_reader.hpp:22-326,
// src/jpeg_stream_reader.cpp:87-1014): same accepted syntax, same error codes, same callback behaviour.  The entropy-coded
// data itself is never looked at here -- the CUDA engine reports how many bytes a scan consumed.
#pragma once

#include "../jls_common.h"
#include "abi_support.hpp"

#include <vector>

namespace jls::host {

class StreamReader final
{
public:
    void source(const uint8_t* data, size_t size) noexcept
    {
        begin_ = data;
        position_ = data;
        end_ = data + size;
    }

    // Reads up to (and including) the first SOS segment, or stops after a SPIFF header when one is requested and found.
    void read_header(charls_spiff_header* spiff_header = nullptr, bool* spiff_header_found = nullptr);
    void read_next_start_of_scan();
    void read_end_of_image();

    const charls_frame_info& frame_info() const noexcept { return frame_info_; }
    const charls_jpegls_pc_parameters& preset_coding_parameters() const noexcept { return preset_; }
    charls_jpegls_pc_parameters validated_preset_coding_parameters() const;
    bool end_of_image() const noexcept { return state_ == State::after_end_of_image; }
    int32_t compressed_data_format() const noexcept { return compressed_data_format_; }

    size_t component_count() const noexcept { return components_.size(); }
    int32_t near_lossless(size_t component) const noexcept { return components_[component].near_lossless; }
    int32_t interleave_mode(size_t component) const noexcept { return components_[component].interleave_mode; }
    int32_t mapping_table_id(size_t component) const noexcept { return components_[component].table_id; }

    // parameters of the scan whose SOS was read last
    uint32_t scan_component_count() const noexcept { return scan_component_count_; }
    int32_t scan_interleave_mode() const noexcept { return scan_interleave_mode_; }
    int32_t scan_near_lossless() const noexcept { return scan_near_lossless_; }
    int32_t color_transformation() const noexcept { return color_transformation_; }
    uint32_t restart_interval() const noexcept { return restart_interval_; }

    // The side table of interval offsets found in front of the current scan's SOS (jls_common.h), if it was complete and
    // well formed: entry_offsets[s] = offset in the source of segment s's first entry.  total == 0: none.
    struct OffsetTable
    {
        uint32_t total{};
        size_t entry_offsets[jls::offset_table_max_segments]{};
    };
    const OffsetTable& scan_offset_table() const noexcept { return scan_offset_table_; }

    // position of the first entropy-coded byte of the current scan, relative to the start of the source
    size_t position() const noexcept { return static_cast<size_t>(position_ - begin_); }
    size_t source_size() const noexcept { return static_cast<size_t>(end_ - begin_); }
    const uint8_t* source_data() const noexcept { return begin_; }
    void advance(size_t count) noexcept { position_ += count; }

    void at_comment(charls_at_comment_handler handler, void* context) noexcept
    {
        comment_handler_ = handler;
        comment_context_ = context;
    }
    void at_application_data(charls_at_application_data_handler handler, void* context) noexcept
    {
        application_data_handler_ = handler;
        application_data_context_ = context;
    }

    size_t mapping_table_count() const noexcept { return mapping_tables_.size(); }
    int32_t find_mapping_table_index(uint8_t table_id) const noexcept;
    charls_mapping_table_info mapping_table_info(size_t index) const;
    void mapping_table_data(size_t index, uint8_t* destination, size_t size) const;

private:
    enum class State
    {
        before_start_of_image,
        header_section,
        spiff_header_section,
        frame_section,
        scan_section,
        bit_stream_section,
        after_end_of_image
    };

    struct Component
    {
        uint8_t id;
        uint8_t near_lossless;
        uint8_t table_id;
        int32_t interleave_mode;
    };

    struct MappingTable
    {
        uint8_t table_id;
        uint8_t entry_size;
        std::vector<std::pair<const uint8_t*, size_t>> fragments;
        size_t data_size() const noexcept
        {
            size_t n = 0;
            for (const auto& f : fragments)
                n += f.second;
            return n;
        }
    };

    uint8_t byte_checked();
    uint8_t next_marker_code();
    uint8_t marker_code_after_start_byte();
    void validate_marker(uint8_t code) const;
    void begin_segment();
    void require_segment_at_least(size_t n) const;
    void require_segment_exactly(size_t n) const;
    void skip_rest_of_segment() noexcept { position_ = segment_end_; }
    size_t segment_size() const noexcept { return static_cast<size_t>(segment_end_ - segment_begin_); }

    uint8_t get8() noexcept { return *position_++; }
    uint16_t get16() noexcept
    {
        const uint16_t v = static_cast<uint16_t>((position_[0] << 8) | position_[1]);
        position_ += 2;
        return v;
    }
    uint32_t get24() noexcept
    {
        const uint32_t hi = get8();
        return (hi << 16) | get16();
    }
    uint32_t get32() noexcept
    {
        const uint32_t hi = get16();
        return (hi << 16) | get16();
    }

    void read_segment(uint8_t code, charls_spiff_header* header, bool* found);
    void read_spiff_directory_entry(uint8_t code);
    void read_start_of_frame();
    void read_start_of_scan();
    void read_preset_parameters();
    void read_restart_interval();
    uint32_t read_number_of_lines();
    void read_application_data8(charls_spiff_header* header, bool* found);
    void call_application_data_handler(uint8_t code) const;
    void read_offset_table_segment();
    void find_define_number_of_lines();
    void set_height(uint32_t height, bool final_update);
    void set_width(uint32_t width);
    uint32_t maximum_sample_value() const noexcept;
    bool abbreviated_table_specification() const;
    bool has_external_mapping_table_ids() const noexcept;
    std::vector<MappingTable>::const_iterator find_table(uint8_t id) const noexcept;

    const uint8_t* begin_{};
    const uint8_t* position_{};
    const uint8_t* end_{};
    const uint8_t* segment_begin_{};
    const uint8_t* segment_end_{};

    State state_{State::before_start_of_image};
    charls_frame_info frame_info_{};
    charls_jpegls_pc_parameters preset_{};
    std::vector<Component> components_;
    std::vector<MappingTable> mapping_tables_;
    uint32_t read_component_count_{};
    uint32_t scan_component_count_{};
    int32_t scan_interleave_mode_{};
    int32_t scan_near_lossless_{};
    int32_t color_transformation_{};
    uint32_t restart_interval_{};
    bool dnl_expected_{};
    int32_t compressed_data_format_{};
    OffsetTable pending_offset_table_{}; // segments seen since the last SOS
    uint32_t pending_offset_entries_{};  // entries they hold so far; ~0 after a malformed segment
    OffsetTable scan_offset_table_{};
    charls_at_comment_handler comment_handler_{};
    void* comment_context_{};
    charls_at_application_data_handler application_data_handler_{};
    void* application_data_context_{};
};

} // namespace jls::host
