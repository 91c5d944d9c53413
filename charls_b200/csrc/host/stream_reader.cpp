// stream_reader.cpp -- see stream_reader.hpp.  Behaviour restated from reference src/jpeg_stream_reader.cpp.
#include "stream_reader.hpp"

#include "../jls_params.hpp"

#include <algorithm>
#include <cstring>

namespace jls::host {

namespace {

// SOF markers of the other JPEG coding processes (T.81 table B.1, T.870): reported as "encoding not supported".
bool is_other_jpeg_frame_marker(uint8_t code) noexcept
{
    switch (code)
    {
    case 0xC0: case 0xC1: case 0xC2: case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xF9:
        return true;
    default:
        return false;
    }
}

bool is_restart_marker(uint8_t code) noexcept { return code >= 0xD0 && code <= 0xD7; }

bool is_valid_interleave_mode(int32_t mode) noexcept { return mode >= 0 && mode <= 2; }

} // namespace

// reference jpeg_stream_reader.cpp:87-149
void StreamReader::read_header(charls_spiff_header* spiff_header, bool* spiff_header_found)
{
    if (state_ == State::before_start_of_image)
    {
        if (next_marker_code() != marker_soi)
            fail(CHARLS_JPEGLS_ERRC_START_OF_IMAGE_MARKER_NOT_FOUND);
        components_.reserve(4);
        state_ = State::header_section;
    }

    for (;;)
    {
        const uint8_t code = next_marker_code();
        if (code == marker_eoi)
        {
            if (abbreviated_table_specification())
            {
                state_ = State::after_end_of_image;
                compressed_data_format_ = 3; // abbreviated_table_specification
                return;
            }
            fail(CHARLS_JPEGLS_ERRC_UNEXPECTED_END_OF_IMAGE_MARKER);
        }
        validate_marker(code);
        begin_segment();
        if (state_ == State::spiff_header_section)
            read_spiff_directory_entry(code);
        else
            read_segment(code, spiff_header, spiff_header_found);

        if (state_ == State::header_section && spiff_header_found && *spiff_header_found)
        {
            state_ = State::spiff_header_section;
            return;
        }
        if (state_ == State::bit_stream_section)
        {
            if (frame_info_.height == 0)
                find_define_number_of_lines();
            if (frame_info_.width < 1)
                fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_WIDTH);
            // colour transformation vs frame (reference color_transform.hpp:11-17, jpeg_stream_reader.cpp:876-881)
            if (color_transformation_ != 0 &&
                !(frame_info_.component_count == 3 && (frame_info_.bits_per_sample == 8 || frame_info_.bits_per_sample == 16) &&
                  near_lossless(0) == 0 && interleave_mode(0) != 0))
                fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_COLOR_TRANSFORMATION);
            return;
        }
    }
}

// reference jpeg_stream_reader.cpp:152-172
void StreamReader::read_end_of_image()
{
    uint8_t b = byte_checked();
    if (b == 0) // tolerated padding byte written by some legacy encoders
        b = byte_checked();
    if (b != marker_start || marker_code_after_start_byte() != marker_eoi)
        fail(CHARLS_JPEGLS_ERRC_END_OF_IMAGE_MARKER_NOT_FOUND);
    compressed_data_format_ = has_external_mapping_table_ids() ? 2 : 1; // abbreviated_image_data : interchange
    state_ = State::after_end_of_image;
}

// reference jpeg_stream_reader.cpp:175-189
void StreamReader::read_next_start_of_scan()
{
    state_ = State::scan_section;
    do
    {
        const uint8_t code = next_marker_code();
        validate_marker(code);
        begin_segment();
        read_segment(code, nullptr, nullptr);
    } while (state_ == State::scan_section);
}

uint8_t StreamReader::byte_checked()
{
    if (position_ == end_)
        fail(CHARLS_JPEGLS_ERRC_NEED_MORE_DATA);
    return *position_++;
}

uint8_t StreamReader::next_marker_code()
{
    if (byte_checked() != marker_start)
        fail(CHARLS_JPEGLS_ERRC_JPEG_MARKER_START_BYTE_NOT_FOUND);
    return marker_code_after_start_byte();
}

uint8_t StreamReader::marker_code_after_start_byte()
{
    uint8_t code = byte_checked();
    while (code == marker_start) // fill bytes, T.81 B.1.1.2
        code = byte_checked();
    return code;
}

// reference jpeg_stream_reader.cpp:216-282
void StreamReader::validate_marker(uint8_t code) const
{
    if (code == marker_sos)
    {
        if (state_ != State::scan_section)
            fail(CHARLS_JPEGLS_ERRC_UNEXPECTED_START_OF_SCAN_MARKER);
        return;
    }
    if (code == marker_sof55)
    {
        if (state_ == State::scan_section)
            fail(CHARLS_JPEGLS_ERRC_DUPLICATE_START_OF_FRAME_MARKER);
        return;
    }
    if (code == marker_dri || code == marker_lse || code == marker_com || (code >= marker_app0 && code <= marker_app15))
        return;
    if (code == marker_dnl)
    {
        if (!dnl_expected_)
            fail(CHARLS_JPEGLS_ERRC_UNEXPECTED_DEFINE_NUMBER_OF_LINES_MARKER);
        return;
    }
    if (code == marker_soi)
        fail(CHARLS_JPEGLS_ERRC_DUPLICATE_START_OF_IMAGE_MARKER);
    if (is_other_jpeg_frame_marker(code))
        fail(CHARLS_JPEGLS_ERRC_ENCODING_NOT_SUPPORTED);
    if (is_restart_marker(code))
        fail(CHARLS_JPEGLS_ERRC_UNEXPECTED_RESTART_MARKER);
    fail(CHARLS_JPEGLS_ERRC_UNKNOWN_JPEG_MARKER_FOUND);
}

// reference jpeg_stream_reader.cpp:707-716
void StreamReader::begin_segment()
{
    if (position_ + 2 > end_)
        fail(CHARLS_JPEGLS_ERRC_NEED_MORE_DATA);
    const size_t size = get16();
    if (size < 2 || position_ + (size - 2) > end_)
        fail(CHARLS_JPEGLS_ERRC_INVALID_MARKER_SEGMENT_SIZE);
    segment_begin_ = position_;
    segment_end_ = position_ + (size - 2);
}

void StreamReader::require_segment_at_least(size_t n) const
{
    if (n > segment_size())
        fail(CHARLS_JPEGLS_ERRC_INVALID_MARKER_SEGMENT_SIZE);
}

void StreamReader::require_segment_exactly(size_t n) const
{
    if (n != segment_size())
        fail(CHARLS_JPEGLS_ERRC_INVALID_MARKER_SEGMENT_SIZE);
}

// reference jpeg_stream_reader.cpp:346-407
void StreamReader::read_segment(uint8_t code, charls_spiff_header* header, bool* found)
{
    switch (code)
    {
    case marker_sof55:
        read_start_of_frame();
        break;
    case marker_sos:
        read_start_of_scan();
        break;
    case marker_lse:
        read_preset_parameters();
        break;
    case marker_dri:
        read_restart_interval();
        break;
    case marker_dnl:
        (void)read_number_of_lines();
        dnl_expected_ = false;
        break;
    case marker_app8:
        read_application_data8(header, found);
        break;
    case marker_com:
        if (comment_handler_ &&
            comment_handler_(segment_size() == 0 ? nullptr : position_, segment_size(), comment_context_) != 0)
            fail(CHARLS_JPEGLS_ERRC_CALLBACK_FAILED);
        skip_rest_of_segment();
        break;
    default: // APP0-7, APP9-15
        call_application_data_handler(code);
        if (code == jls::offset_table_marker)
            read_offset_table_segment();
        skip_rest_of_segment();
        break;
    }
}

// reference jpeg_stream_reader.cpp:409-422
void StreamReader::read_spiff_directory_entry(uint8_t code)
{
    if (code != marker_app8)
        fail(CHARLS_JPEGLS_ERRC_MISSING_END_OF_SPIFF_DIRECTORY);
    require_segment_at_least(4);
    if (get32() == 1) // end of directory
    {
        require_segment_exactly(6); // tag + the dummy SOI
        state_ = State::frame_section;
    }
    skip_rest_of_segment();
}

// reference jpeg_stream_reader.cpp:425-458
void StreamReader::read_start_of_frame()
{
    require_segment_at_least(6);
    frame_info_.bits_per_sample = get8();
    if (frame_info_.bits_per_sample < 2 || frame_info_.bits_per_sample > 16)
        fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_BITS_PER_SAMPLE);
    set_height(get16(), false);
    set_width(get16());
    frame_info_.component_count = get8();
    if (frame_info_.component_count == 0)
        fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_COMPONENT_COUNT);
    require_segment_exactly(static_cast<size_t>(frame_info_.component_count) * 3 + 6);
    for (int32_t i = 0; i < frame_info_.component_count; ++i)
    {
        const uint8_t id = get8();
        for (const Component& c : components_)
            if (c.id == id)
                fail(CHARLS_JPEGLS_ERRC_DUPLICATE_COMPONENT_ID_IN_SOF_SEGMENT);
        components_.push_back({id, 0, 0, 0});
        if (get8() != 0x11) // sub-sampling is not part of JPEG-LS as implemented
            fail(CHARLS_JPEGLS_ERRC_PARAMETER_VALUE_NOT_SUPPORTED);
        (void)get8(); // Tq
    }
    state_ = State::scan_section;
}

// reference jpeg_stream_reader.cpp:608-655
void StreamReader::read_start_of_scan()
{
    require_segment_at_least(1);
    const uint32_t count = get8();
    if (count < 1 || count > static_cast<uint32_t>(maximum_component_count_in_scan) ||
        count > static_cast<uint32_t>(frame_info_.component_count) - read_component_count_)
        fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_COMPONENT_COUNT);
    scan_component_count_ = count;
    read_component_count_ += count;
    require_segment_exactly(count * 2 + 4);

    uint8_t ids[4] = {};
    uint8_t table_ids[4] = {};
    for (uint32_t i = 0; i < count; ++i)
    {
        ids[i] = get8();
        table_ids[i] = get8();
    }
    scan_near_lossless_ = get8();
    if (scan_near_lossless_ > maximum_near_lossless(static_cast<int32_t>(maximum_sample_value())))
        fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_NEAR_LOSSLESS);
    scan_interleave_mode_ = get8();
    if (!is_valid_interleave_mode(scan_interleave_mode_) || (count == 1 && scan_interleave_mode_ != 0))
        fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_INTERLEAVE_MODE);

    for (uint32_t i = 0; i < count; ++i)
    {
        // defaults need no lookup (and tolerate unknown ids), reference jpeg_stream_reader.cpp:969-985
        if (table_ids[i] == 0 && scan_near_lossless_ == 0 && scan_interleave_mode_ == 0)
            continue;
        auto it = std::find_if(components_.begin(), components_.end(), [&](const Component& c) { return c.id == ids[i]; });
        if (it == components_.end())
            fail(CHARLS_JPEGLS_ERRC_UNKNOWN_COMPONENT_ID);
        it->near_lossless = static_cast<uint8_t>(scan_near_lossless_);
        it->table_id = table_ids[i];
        it->interleave_mode = scan_interleave_mode_;
    }
    if ((get8() & 0x0F) != 0) // point transform
        fail(CHARLS_JPEGLS_ERRC_PARAMETER_VALUE_NOT_SUPPORTED);
    state_ = State::bit_stream_section;

    // a side table of interval offsets belongs to the scan that follows it
    scan_offset_table_ = OffsetTable{};
    if (pending_offset_entries_ != ~0U && pending_offset_table_.total != 0 && pending_offset_entries_ == pending_offset_table_.total)
        scan_offset_table_ = pending_offset_table_;
    pending_offset_table_ = OffsetTable{};
    pending_offset_entries_ = 0;
}

// An APP11 segment that carries (part of) the side table of interval offsets (jls_common.h).  Anything that is not exactly
// that -- another application's APP11, an unknown version, segments out of order -- is not an error: the stream is then
// decoded like one without a table (the reference does not know the segment at all).
void StreamReader::read_offset_table_segment()
{
    static const uint8_t identifier[8] = {'J', 'L', 'S', '-', 'O', 'F', 'F', 'T'};
    if (segment_size() < jls::offset_table_header_bytes || std::memcmp(position_, identifier, sizeof(identifier)) != 0)
        return;
    const uint8_t* const saved = position_;
    position_ += sizeof(identifier);
    const uint32_t version = get8();
    (void)get8();
    const uint32_t first = get32(), count = get32(), total = get32();
    const size_t entries_at = static_cast<size_t>(position_ - begin_);
    position_ = saved;
    if (pending_offset_entries_ == ~0U)
        return;
    const uint32_t segment = first / jls::offset_table_entries_per_segment;
    const bool well_formed = version == 1 && total >= 2 && count >= 1 && first < total && first == pending_offset_entries_ &&
                             first % jls::offset_table_entries_per_segment == 0 && segment < jls::offset_table_max_segments &&
                             count == (total - first < jls::offset_table_entries_per_segment ? total - first : jls::offset_table_entries_per_segment) &&
                             segment_size() == jls::offset_table_header_bytes + static_cast<size_t>(count) * 4U &&
                             (first == 0 || total == pending_offset_table_.total);
    if (!well_formed)
    {
        pending_offset_entries_ = ~0U;
        return;
    }
    pending_offset_table_.total = total;
    pending_offset_table_.entry_offsets[segment] = entries_at;
    pending_offset_entries_ += count;
}

// reference jpeg_stream_reader.cpp:488-606
void StreamReader::read_preset_parameters()
{
    require_segment_at_least(1);
    const uint8_t type = get8();
    switch (type)
    {
    case 1: // preset coding parameters; validated later, when NEAR is known
        require_segment_exactly(11);
        preset_.maximum_sample_value = get16();
        preset_.threshold1 = get16();
        preset_.threshold2 = get16();
        preset_.threshold3 = get16();
        preset_.reset_value = get16();
        return;
    case 2: // mapping table specification
    case 3: // mapping table continuation
    {
        require_segment_at_least(3);
        const uint8_t table_id = get8();
        const uint8_t entry_size = get8();
        const auto existing = find_table(table_id);
        if (type == 2)
        {
            if (table_id == 0 || existing != mapping_tables_.cend())
                fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_MAPPING_TABLE_ID);
            MappingTable table{table_id, entry_size, {}};
            table.fragments.emplace_back(position_, static_cast<size_t>(segment_end_ - position_));
            mapping_tables_.push_back(std::move(table));
        }
        else
        {
            if (existing == mapping_tables_.cend() || existing->entry_size != entry_size)
                fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_MAPPING_TABLE_CONTINUATION);
            mapping_tables_[static_cast<size_t>(existing - mapping_tables_.cbegin())].fragments.emplace_back(
                position_, static_cast<size_t>(segment_end_ - position_));
        }
        skip_rest_of_segment();
        return;
    }
    case 4: // oversize image dimensions
    {
        require_segment_at_least(2);
        const uint8_t width_bytes = get8();
        uint32_t height, width;
        switch (width_bytes)
        {
        case 2:
            require_segment_exactly(2 + 4);
            height = get16();
            width = get16();
            break;
        case 3:
            require_segment_exactly(2 + 6);
            height = get24();
            width = get24();
            break;
        case 4:
            require_segment_exactly(2 + 8);
            height = get32();
            width = get32();
            break;
        default:
            fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_JPEGLS_PRESET_PARAMETERS);
        }
        set_height(height, false);
        set_width(width);
        return;
    }
    default:
        // types 5..13 belong to JPEG-LS part 2 (ISO/IEC 14495-2)
        fail(type <= 0x0D ? CHARLS_JPEGLS_ERRC_JPEGLS_PRESET_EXTENDED_PARAMETER_TYPE_NOT_SUPPORTED
                          : CHARLS_JPEGLS_ERRC_INVALID_JPEGLS_PRESET_PARAMETER_TYPE);
    }
}

// reference jpeg_stream_reader.cpp:586-607 -- 2, 3 or 4 byte Ri (ISO/IEC 14495-1 C.2.5)
void StreamReader::read_restart_interval()
{
    switch (segment_size())
    {
    case 2:
        restart_interval_ = get16();
        break;
    case 3:
        restart_interval_ = get24();
        break;
    case 4:
        restart_interval_ = get32();
        break;
    default:
        fail(CHARLS_JPEGLS_ERRC_INVALID_MARKER_SEGMENT_SIZE);
    }
}

// reference jpeg_stream_reader.cpp:478-495
uint32_t StreamReader::read_number_of_lines()
{
    switch (segment_size())
    {
    case 2:
        return get16();
    case 3:
        return get24();
    case 4:
        return get32();
    default:
        fail(CHARLS_JPEGLS_ERRC_INVALID_MARKER_SEGMENT_SIZE);
    }
}

// reference jpeg_stream_reader.cpp:733-822
void StreamReader::read_application_data8(charls_spiff_header* header, bool* found)
{
    call_application_data_handler(marker_app8);
    if (found)
        *found = false;

    if (segment_size() == 5)
    {
        if (std::memcmp(position_, "mrfx", 4) == 0) // HP colour transformation segment
        {
            const uint8_t transformation = position_[4];
            if (transformation <= 3)
                color_transformation_ = transformation;
            else if (transformation == 4 || transformation == 5)
                fail(CHARLS_JPEGLS_ERRC_COLOR_TRANSFORM_NOT_SUPPORTED);
            else
                fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_COLOR_TRANSFORMATION);
        }
    }
    else if (header && found && segment_size() >= 30)
    {
        static const uint8_t magic[6] = {'S', 'P', 'I', 'F', 'F', 0};
        *header = charls_spiff_header{};
        if (std::memcmp(position_, magic, 6) == 0 && position_[6] <= 2) // unknown major versions: as if absent
        {
            position_ += 8;
            header->profile_id = get8();
            header->component_count = get8();
            header->height = get32();
            header->width = get32();
            header->color_space = get8();
            header->bits_per_sample = get8();
            header->compression_type = get8();
            header->resolution_units = get8();
            header->vertical_resolution = get32();
            header->horizontal_resolution = get32();
            *found = true;
        }
    }
    skip_rest_of_segment();
}

void StreamReader::call_application_data_handler(uint8_t code) const
{
    if (application_data_handler_ &&
        application_data_handler_(code - marker_app0, segment_size() == 0 ? nullptr : position_, segment_size(),
                                  application_data_context_) != 0)
        fail(CHARLS_JPEGLS_ERRC_CALLBACK_FAILED);
}

// reference jpeg_stream_reader.cpp:921-946: with Y = 0 in SOF a DNL segment must end the first scan
void StreamReader::find_define_number_of_lines()
{
    for (const uint8_t* p = position_; p + 1 < end_; ++p)
    {
        if (p[0] != marker_start)
            continue;
        const uint8_t code = p[1];
        if (code < 0x80 || code == marker_start)
            continue;
        if (code != marker_dnl)
            break;
        const uint8_t* resume = position_;
        position_ = p + 2;
        begin_segment();
        set_height(read_number_of_lines(), true);
        dnl_expected_ = true;
        position_ = resume;
        return;
    }
    fail(CHARLS_JPEGLS_ERRC_DEFINE_NUMBER_OF_LINES_MARKER_NOT_FOUND);
}

// reference jpeg_stream_reader.cpp:884-908
void StreamReader::set_height(uint32_t height, bool final_update)
{
    if (height == 0 && !final_update)
        return;
    if (frame_info_.height != 0 || height < 1 || height > maximum_height)
        fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_HEIGHT);
    frame_info_.height = height;
}

void StreamReader::set_width(uint32_t width)
{
    if (width == 0)
        return;
    if (frame_info_.width != 0 || width > maximum_width)
        fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_WIDTH);
    frame_info_.width = width;
}

uint32_t StreamReader::maximum_sample_value() const noexcept
{
    if (preset_.maximum_sample_value != 0)
        return static_cast<uint32_t>(preset_.maximum_sample_value);
    return static_cast<uint32_t>(maximum_bit_sample_value(frame_info_.bits_per_sample));
}

// reference jpeg_stream_reader.cpp:285-295
charls_jpegls_pc_parameters StreamReader::validated_preset_coding_parameters() const
{
    const PresetCodingParameters in{preset_.maximum_sample_value, preset_.threshold1, preset_.threshold2, preset_.threshold3,
                                    preset_.reset_value};
    PresetCodingParameters out{};
    if (!validate_preset_parameters(in, maximum_bit_sample_value(frame_info_.bits_per_sample), scan_near_lossless_, &out))
        fail(CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_JPEGLS_PRESET_PARAMETERS);
    return {out.maximum_sample_value, out.threshold1, out.threshold2, out.threshold3, out.reset_value};
}

// reference jpeg_stream_reader.hpp:215-224
bool StreamReader::abbreviated_table_specification() const
{
    if (mapping_tables_.empty())
        return false;
    if (state_ == State::frame_section)
        fail(CHARLS_JPEGLS_ERRC_ABBREVIATED_FORMAT_AND_SPIFF_HEADER_MISMATCH);
    return state_ == State::header_section;
}

bool StreamReader::has_external_mapping_table_ids() const noexcept
{
    for (const Component& c : components_)
        if (c.table_id != 0 && find_table(c.table_id) == mapping_tables_.cend())
            return true;
    return false;
}

std::vector<StreamReader::MappingTable>::const_iterator StreamReader::find_table(uint8_t id) const noexcept
{
    return std::find_if(mapping_tables_.cbegin(), mapping_tables_.cend(), [id](const MappingTable& t) { return t.table_id == id; });
}

int32_t StreamReader::find_mapping_table_index(uint8_t table_id) const noexcept
{
    const auto it = find_table(table_id);
    return it == mapping_tables_.cend() ? CHARLS_MAPPING_TABLE_MISSING : static_cast<int32_t>(it - mapping_tables_.cbegin());
}

charls_mapping_table_info StreamReader::mapping_table_info(size_t index) const
{
    const MappingTable& t = mapping_tables_[index];
    return {t.table_id, t.entry_size, static_cast<uint32_t>(t.data_size())};
}

void StreamReader::mapping_table_data(size_t index, uint8_t* destination, size_t size) const
{
    const MappingTable& t = mapping_tables_[index];
    if (t.data_size() > size)
        fail(CHARLS_JPEGLS_ERRC_DESTINATION_TOO_SMALL);
    for (const auto& f : t.fragments)
    {
        std::memcpy(destination, f.first, f.second);
        destination += f.second;
    }
}

} // namespace jls::host
