// stream_writer.hpp -- writes the JPEG-LS container (marker segments) into the caller's destination buffer.
//
// Host-side counterpart of the reference's jpeg_stream_writer (src/jpeg_stream_writer.hpp:30-103,
// src/jpeg_stream_writer.cpp:20-245), plus the one segment the reference cannot write: DRI (ISO/IEC 14495-1 C.2.5).
#pragma once

#include "abi_support.hpp"

#include <array>
#include <vector>
#include "../jls_common.h"

#include <cstring>

namespace jls::host {

class StreamWriter final
{
public:
    void destination(uint8_t* data, size_t size) noexcept
    {
        // the write position survives, as in the reference (src/jpeg_stream_writer.hpp:129-132): a segment that did not
        // fit leaves the bytes written before it counted, whatever buffer is set next; only rewind() starts over
        data_ = data;
        size_ = size;
    }

    size_t bytes_written() const noexcept { return position_; }
    uint8_t* remaining_data() const noexcept { return data_ + position_; }
    size_t remaining_size() const noexcept { return position_ < size_ ? size_ - position_ : 0; }
    void advance(size_t count) noexcept { position_ += count; }

    void rewind() noexcept
    {
        position_ = 0;
        component_index_ = 0;
    }

    void set_mapping_table_id(size_t component_index, int32_t table_id)
    {
        if (component_index >= table_ids_.size())
            table_ids_.resize(component_index + 1);
        table_ids_[component_index] = static_cast<uint8_t>(table_id);
    }

    void write_start_of_image() { write_marker_only(marker_soi); }

    void write_end_of_image(bool even_destination_size)
    {
        if (even_destination_size && position_ % 2 != 0)
            put8(marker_start); // a fill byte in front of EOI (reference jpeg_stream_writer.cpp:26-35)
        write_marker_only(marker_eoi);
    }

    void write_spiff_header(const charls_spiff_header& h)
    {
        begin_segment(marker_app8, 30);
        static const uint8_t magic[6] = {'S', 'P', 'I', 'F', 'F', 0};
        put_bytes(magic, 6);
        put8(2); // SPIFF version 2.0 (ISO/IEC 14495-1 4.8.1)
        put8(0);
        put8(static_cast<uint8_t>(h.profile_id));
        put8(static_cast<uint8_t>(h.component_count));
        put32(h.height);
        put32(h.width);
        put8(static_cast<uint8_t>(h.color_space));
        put8(static_cast<uint8_t>(h.bits_per_sample));
        put8(static_cast<uint8_t>(h.compression_type));
        put8(static_cast<uint8_t>(h.resolution_units));
        put32(h.vertical_resolution);
        put32(h.horizontal_resolution);
    }

    void write_spiff_directory_entry(uint32_t tag, const uint8_t* data, size_t size)
    {
        begin_segment(marker_app8, 4 + size);
        put32(tag);
        put_bytes(data, size);
    }

    void write_spiff_end_of_directory()
    {
        // The EOD entry carries a dummy SOI so that a plain JPEG-LS decoder can start there (T.84 F.2.2.3).
        static const uint8_t payload[6] = {0, 0, 0, 1, 0xFF, marker_soi};
        begin_segment(marker_app8, 6);
        put_bytes(payload, 6);
    }

    // Returns true when the dimensions do not fit 16 bits and an LSE oversize segment must follow.
    bool write_start_of_frame(const charls_frame_info& f)
    {
        begin_segment(marker_sof55, 6 + static_cast<size_t>(f.component_count) * 3);
        put8(static_cast<uint8_t>(f.bits_per_sample));
        const bool oversized = f.width > 65535 || f.height > 65535;
        put16(oversized ? 0 : static_cast<uint16_t>(f.height));
        put16(oversized ? 0 : static_cast<uint16_t>(f.width));
        put8(static_cast<uint8_t>(f.component_count));
        for (int32_t id = 1; id <= f.component_count; ++id)
        {
            put8(static_cast<uint8_t>(id)); // component ids start at 1 like the T.87 H.4 sample
            put8(0x11);                     // no sub-sampling
            put8(0);                        // Tq: reserved
        }
        return oversized;
    }

    void write_color_transform(int32_t transformation)
    {
        const uint8_t payload[5] = {'m', 'r', 'f', 'x', static_cast<uint8_t>(transformation)};
        begin_segment(marker_app8, 5);
        put_bytes(payload, 5);
    }

    void write_comment(const uint8_t* data, size_t size)
    {
        begin_segment(marker_com, size);
        put_bytes(data, size);
    }

    void write_application_data(int32_t id, const uint8_t* data, size_t size)
    {
        begin_segment(static_cast<uint8_t>(marker_app0 + id), size);
        put_bytes(data, size);
    }

    void write_preset_coding_parameters(const charls_jpegls_pc_parameters& pc)
    {
        begin_segment(marker_lse, 11);
        put8(1);
        put16(static_cast<uint16_t>(pc.maximum_sample_value));
        put16(static_cast<uint16_t>(pc.threshold1));
        put16(static_cast<uint16_t>(pc.threshold2));
        put16(static_cast<uint16_t>(pc.threshold3));
        put16(static_cast<uint16_t>(pc.reset_value));
    }

    void write_oversize_dimensions(uint32_t height, uint32_t width)
    {
        begin_segment(marker_lse, 10); // ISO/IEC 14495-1 C.2.4.1.4, Wxy = 4
        put8(4);
        put8(4);
        put32(height);
        put32(width);
    }

    void write_mapping_table(int32_t table_id, int32_t entry_size, const uint8_t* data, size_t size)
    {
        constexpr size_t max_fragment = segment_max_data_size - 3;
        size_t offset = 0;
        uint8_t type = 2; // mapping table specification, then continuations (type 3)
        do
        {
            const size_t n = size - offset < max_fragment ? size - offset : max_fragment;
            begin_segment(marker_lse, 3 + n);
            put8(type);
            put8(static_cast<uint8_t>(table_id));
            put8(static_cast<uint8_t>(entry_size));
            put_bytes(data + offset, n);
            offset += n;
            type = 3;
        } while (offset < size);
    }

    // DRI, ISO/IEC 14495-1 C.2.5: 2-byte Ri when it fits, else 4 bytes.  Not present in the reference writer.
    void write_define_restart_interval(uint32_t restart_interval)
    {
        if (restart_interval < 65536)
        {
            begin_segment(marker_dri, 2);
            put16(static_cast<uint16_t>(restart_interval));
        }
        else
        {
            begin_segment(marker_dri, 4);
            put32(restart_interval);
        }
    }

    // Reserves the side table of interval offsets (jls_common.h) for the scan whose SOS follows: headers are written, the
    // entries stay zero.  entry_positions[s] = offset in the destination of segment s's first entry.
    void write_offset_table_placeholder(uint32_t intervals, size_t (&entry_positions)[jls::offset_table_max_segments])
    {
        const uint32_t total = intervals + 1U;
        uint32_t first = 0;
        for (uint32_t segment = 0; first < total; ++segment)
        {
            const uint32_t count = total - first < jls::offset_table_entries_per_segment ? total - first : jls::offset_table_entries_per_segment;
            begin_segment(jls::offset_table_marker, jls::offset_table_header_bytes + static_cast<size_t>(count) * 4U);
            static const uint8_t identifier[8] = {'J', 'L', 'S', '-', 'O', 'F', 'F', 'T'};
            put_bytes(identifier, sizeof(identifier));
            put8(1);
            put8(0);
            put32(first);
            put32(count);
            put32(total);
            entry_positions[segment] = position_;
            std::memset(data_ + position_, 0, static_cast<size_t>(count) * 4U);
            position_ += static_cast<size_t>(count) * 4U;
            first += count;
        }
    }

    uint8_t* data() const noexcept { return data_; }

    void write_start_of_scan(int32_t component_count, int32_t near_lossless, int32_t interleave_mode)
    {
        begin_segment(marker_sos, 1 + static_cast<size_t>(component_count) * 2 + 3);
        put8(static_cast<uint8_t>(component_count));
        for (int32_t i = 0; i < component_count; ++i)
        {
            put8(static_cast<uint8_t>(component_index_ + 1));
            put8(component_index_ < table_ids_.size() ? table_ids_[component_index_] : 0);
            ++component_index_;
        }
        put8(static_cast<uint8_t>(near_lossless));
        put8(static_cast<uint8_t>(interleave_mode));
        put8(0); // point transform
    }

private:
    void begin_segment(uint8_t marker, size_t data_size)
    {
        if (position_ + 4 + data_size > size_)
            fail(CHARLS_JPEGLS_ERRC_DESTINATION_TOO_SMALL);
        put8(marker_start);
        put8(marker);
        put16(static_cast<uint16_t>(data_size + 2));
    }

    void write_marker_only(uint8_t marker)
    {
        if (position_ + 2 > size_)
            fail(CHARLS_JPEGLS_ERRC_DESTINATION_TOO_SMALL);
        put8(marker_start);
        put8(marker);
    }

    void put8(uint8_t v)
    {
        if (position_ >= size_)
            fail(CHARLS_JPEGLS_ERRC_DESTINATION_TOO_SMALL);
        data_[position_++] = v;
    }
    void put16(uint16_t v)
    {
        put8(static_cast<uint8_t>(v >> 8));
        put8(static_cast<uint8_t>(v));
    }
    void put32(uint32_t v)
    {
        put16(static_cast<uint16_t>(v >> 16));
        put16(static_cast<uint16_t>(v));
    }
    void put_bytes(const uint8_t* data, size_t size)
    {
        if (size != 0)
            std::memcpy(data_ + position_, data, size);
        position_ += size;
    }

    uint8_t* data_{};
    size_t size_{};
    size_t position_{};
    size_t component_index_{};
    std::vector<uint8_t> table_ids_;
};

} // namespace jls::host
