// jls_common.h -- plain structs shared by the host engine and the CUDA kernels.
//
// Vocabulary (JPEG-LS / CharLS domain): a *scan* holds the entropy-coded data of 1..4 components; a *restart
// interval* is `restart_interval` lines of the scan (all components of those lines in line-interleaved mode) whose
// coding state is independent of every other interval (reference src/scan_decoder_impl.hpp:72-128).  The engine's
// unit of parallel work is the interval; with restart_interval == 1 an interval is one scan line.
#pragma once

#include <cstddef>
#include <cstdint>

namespace jls {

// charls_jpegls_errc values the device can raise (reference include/charls/public_types.h:28-88).
enum : int32_t
{
    err_none = 0,
    err_destination_too_small = 3,
    err_need_more_data = 4,
    err_invalid_data = 5,
    err_restart_marker_not_found = 23,
};

enum : int32_t
{
    ilv_none = 0,
    ilv_line = 1,
    ilv_sample = 2
};

// Everything a kernel needs to know about one scan's coding parameters.  Derived values follow the reference's
// default_traits constructor (src/default_traits.hpp:51-59) and preset parameter defaults
// (src/jpegls_preset_coding_parameters.hpp:24-57).
struct CodecParams
{
    int32_t width;           // pixels per line
    int32_t height;          // lines in the scan
    int32_t components;      // components in this scan (1..4)
    int32_t interleave;      // ilv_*
    int32_t transform;       // 0 none, 1..3 HP1..HP3 (3 components, ILV line/sample, 8 or 16 bit, lossless)
    int32_t bits_per_sample; // 2..16
    int32_t sample_bytes;    // 1 or 2
    int32_t near;            // NEAR
    int32_t maxval;          // 2^bits - 1 (the reference never builds the codec from an LSE MAXVAL, make_scan_codec.cpp:98)
    int32_t range;           // RANGE
    int32_t qbpp;            // ceil(log2 RANGE)
    int32_t bpp;             // ceil(log2 MAXVAL)
    int32_t limit;           // LIMIT
    int32_t t1, t2, t3;      // thresholds
    int32_t reset;           // RESET (already truncated to uint8_t like scan_codec.hpp:129)
    int32_t a_init;          // max(2, (RANGE + 32) / 64)
    int32_t dq;              // 2 * NEAR + 1
    uint32_t dq_magic;       // ceil(2^32 / dq): exact unsigned division by dq for n < 2^32 / dq via mulhi
    int32_t range_dq;        // RANGE * dq
    uint32_t restart_interval; // lines per interval as coded in the stream (0 = none)
    uint32_t lines_per_interval; // effective: restart_interval ? restart_interval : height
    uint32_t interval_count;   // ceil(height / lines_per_interval)
};

// ---------------------------------------------------------------------------------------------------------------------
// Optional side table of interval offsets (extension; not in the reference, which neither writes nor looks at it -- to every
// other JPEG-LS decoder it is an application data segment to skip).  One or more APP11 segments in front of a scan's SOS:
//   "JLS-OFFT" (8 bytes) | version = 1 | 0 | first entry (u32) | entries in this segment (u32) | entries in all (u32) | entries
// all big endian.  Entry j < N (N = number of restart intervals of the scan) is the offset of the first byte of interval j
// from the first entropy-coded byte of the scan, entry N the offset of the marker that closes the scan.  With the table a
// decoder finds its intervals without searching the stream for restart markers (the three marker kernels), and a rank
// that decodes only some lines knows which bytes to fetch.  It is checked against the stream before it is believed
// (k_offsets_from_table, interval_end_status); a stream that does not agree with its table is decoded as if it had none.
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint8_t offset_table_marker = 0xEB; // APP11
constexpr uint32_t offset_table_header_bytes = 22;
constexpr uint32_t offset_table_entries_per_segment = (65533U - offset_table_header_bytes) / 4U;
constexpr uint32_t offset_table_max_segments = 8; // 131 016 entries: heights go up to 100 000
constexpr int32_t errc_offset_table_rejected = 250; // internal: never leaves the engine

inline uint32_t offset_table_segment_count(uint32_t entries)
{
    return (entries + offset_table_entries_per_segment - 1U) / offset_table_entries_per_segment;
}

// bytes the table of a scan with `intervals` restart intervals takes in the stream (marker + length + header + entries)
inline size_t offset_table_bytes(uint32_t intervals)
{
    const uint32_t entries = intervals + 1U;
    return static_cast<size_t>(offset_table_segment_count(entries)) * (4U + offset_table_header_bytes) + static_cast<size_t>(entries) * 4U;
}

// Where the entries of a table sit in device memory (decode) / go (encode): segment s holds entries
// [s * offset_table_entries_per_segment, ...), big endian, at entries[s].
struct OffsetTableRef
{
    uint8_t* entries[offset_table_max_segments];
    uint32_t total; // entries in all segments = intervals + 1; 0 = no table
};

// One image (or one plane of an ILV-none image) to code: where the samples live and where the entropy bytes go.
struct ScanJob
{
    const uint8_t* pixels_in; // encode: source samples (device)
    uint8_t* pixels_out;      // decode: destination samples (device)
    size_t stride;            // bytes between lines of the sample buffer
    const uint8_t* stream_in; // decode: entropy-coded data of the scan (device), first byte after the SOS segment
    size_t stream_in_size;    // decode: bytes available from stream_in to the end of the JPEG-LS stream
    uint8_t* stream_out;      // encode: where the concatenated interval data (+RSTm markers) is written (device)
    size_t stream_out_capacity;
    uint8_t* slots;           // encode scratch: interval_count slots of slot_bytes each
    uint32_t* interval_bytes; // encode scratch: bytes produced per interval           [interval_count]
    uint64_t* interval_offset; // encode: exclusive scan of (bytes + marker)            [interval_count + 1]
                               // decode: start offset of every interval in stream_in   [interval_count + 1]
    uint16_t* line_scratch;   // general (2-D) path: 2 * components * (width + 2) samples per interval
    OffsetTableRef offset_table; // encode: where to write the side table (total = 0: none); decode: the table to check and use
    uint64_t* status;         // first-error key: (interval << 8) | charls_jpegls_errc, ~0 when no error
    uint64_t* result;         // [0] encode: total bytes written; decode: bytes consumed by the scan
                              // [1] decode: code of the marker that closes the scan
};

// Worst case size of one interval's entropy data: every sample costs at most LIMIT bits (regular, run-interruption and
// run-length codes alike), bit stuffing adds at most 1 bit per 7, plus padding and the stuffed byte after a final 0xFF.
inline size_t worst_case_interval_bytes(const CodecParams& p, uint32_t lines)
{
    const size_t bits = static_cast<size_t>(p.width) * static_cast<size_t>(p.components) * static_cast<size_t>(p.limit) *
                            static_cast<size_t>(lines) +
                        64U * lines * static_cast<size_t>(p.components);
    const size_t bytes = bits / 7 + 32;
    return (bytes + 15) & ~static_cast<size_t>(15);
}

} // namespace jls
