/*
 * charls_b200.h -- C ABI of the B200-native JPEG-LS engine.
 *
 * Part 1 declares, with identical names, argument meaning, state machines and error codes, the 48 symbols the
 * reference library (team-charls/charls @ 7b9b2da, libcharls.so.3) exports; every declaration cites the reference
 * header line it replaces (paths relative to the reference repository).  A program or binding written against the
 * reference's headers links against charls_b200/lib/libcharls.so.3 unchanged: it is this library's drop-in boundary.
 * The entropy coding behind charls_jpegls_encoder_encode_*_from_buffer and charls_jpegls_decoder_decode_to_buffer runs
 * as CUDA kernels on the current device; there is no CPU fallback (calls fail with a CUDA-less host).
 *
 * Part 2 (charlsx_*) are extensions that do not exist in the reference: choosing the restart interval the encoder
 * writes (the reference cannot write restart markers at all), device selection, and a batch interface for frames that
 * are already resident in device memory.
 *
 * Plain C99 / C++: only pointers, sizes and fixed-width integers cross the boundary.
 */
#ifndef CHARLS_B200_H
#define CHARLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#define CHARLS_B200_NOEXCEPT noexcept
#else
#define CHARLS_B200_NOEXCEPT
#endif

#if defined(CHARLS_B200_BUILD)
#define CHARLS_B200_API __attribute__((visibility("default")))
#else
#define CHARLS_B200_API extern
#endif

/* ------------------------------------------------------------------------------------------------------------------ */
/* Types (reference include/charls/public_types.h)                                                                     */
/* ------------------------------------------------------------------------------------------------------------------ */

/* charls_jpegls_errc, public_types.h:28-88.  0 = success, 1..38 run-time errors, 100..112 logic errors. */
typedef int32_t charls_jpegls_errc;
enum
{
    CHARLS_JPEGLS_ERRC_SUCCESS = 0,
    CHARLS_JPEGLS_ERRC_NOT_ENOUGH_MEMORY = 1,
    CHARLS_JPEGLS_ERRC_CALLBACK_FAILED = 2,
    CHARLS_JPEGLS_ERRC_DESTINATION_TOO_SMALL = 3,
    CHARLS_JPEGLS_ERRC_NEED_MORE_DATA = 4,
    CHARLS_JPEGLS_ERRC_INVALID_DATA = 5,
    CHARLS_JPEGLS_ERRC_ENCODING_NOT_SUPPORTED = 6,
    CHARLS_JPEGLS_ERRC_PARAMETER_VALUE_NOT_SUPPORTED = 7,
    CHARLS_JPEGLS_ERRC_COLOR_TRANSFORM_NOT_SUPPORTED = 8,
    CHARLS_JPEGLS_ERRC_JPEGLS_PRESET_EXTENDED_PARAMETER_TYPE_NOT_SUPPORTED = 9,
    CHARLS_JPEGLS_ERRC_JPEG_MARKER_START_BYTE_NOT_FOUND = 10,
    CHARLS_JPEGLS_ERRC_START_OF_IMAGE_MARKER_NOT_FOUND = 11,
    CHARLS_JPEGLS_ERRC_INVALID_SPIFF_HEADER = 12,
    CHARLS_JPEGLS_ERRC_UNKNOWN_JPEG_MARKER_FOUND = 13,
    CHARLS_JPEGLS_ERRC_UNEXPECTED_START_OF_SCAN_MARKER = 14,
    CHARLS_JPEGLS_ERRC_INVALID_MARKER_SEGMENT_SIZE = 15,
    CHARLS_JPEGLS_ERRC_DUPLICATE_START_OF_IMAGE_MARKER = 16,
    CHARLS_JPEGLS_ERRC_DUPLICATE_START_OF_FRAME_MARKER = 17,
    CHARLS_JPEGLS_ERRC_DUPLICATE_COMPONENT_ID_IN_SOF_SEGMENT = 18,
    CHARLS_JPEGLS_ERRC_UNEXPECTED_END_OF_IMAGE_MARKER = 19,
    CHARLS_JPEGLS_ERRC_INVALID_JPEGLS_PRESET_PARAMETER_TYPE = 20,
    CHARLS_JPEGLS_ERRC_MISSING_END_OF_SPIFF_DIRECTORY = 21,
    CHARLS_JPEGLS_ERRC_UNEXPECTED_RESTART_MARKER = 22,
    CHARLS_JPEGLS_ERRC_RESTART_MARKER_NOT_FOUND = 23,
    CHARLS_JPEGLS_ERRC_END_OF_IMAGE_MARKER_NOT_FOUND = 24,
    CHARLS_JPEGLS_ERRC_UNEXPECTED_DEFINE_NUMBER_OF_LINES_MARKER = 25,
    CHARLS_JPEGLS_ERRC_DEFINE_NUMBER_OF_LINES_MARKER_NOT_FOUND = 26,
    CHARLS_JPEGLS_ERRC_UNKNOWN_COMPONENT_ID = 27,
    CHARLS_JPEGLS_ERRC_ABBREVIATED_FORMAT_AND_SPIFF_HEADER_MISMATCH = 28,
    CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_WIDTH = 29,
    CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_HEIGHT = 30,
    CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_BITS_PER_SAMPLE = 31,
    CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_COMPONENT_COUNT = 32,
    CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_INTERLEAVE_MODE = 33,
    CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_NEAR_LOSSLESS = 34,
    CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_JPEGLS_PRESET_PARAMETERS = 35,
    CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_COLOR_TRANSFORMATION = 36,
    CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_MAPPING_TABLE_ID = 37,
    CHARLS_JPEGLS_ERRC_INVALID_PARAMETER_MAPPING_TABLE_CONTINUATION = 38,
    CHARLS_JPEGLS_ERRC_INVALID_OPERATION = 100,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT = 101,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_WIDTH = 102,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_HEIGHT = 103,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_BITS_PER_SAMPLE = 104,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_COMPONENT_COUNT = 105,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_INTERLEAVE_MODE = 106,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_NEAR_LOSSLESS = 107,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_JPEGLS_PC_PARAMETERS = 108,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_COLOR_TRANSFORMATION = 109,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_SIZE = 110,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_STRIDE = 111,
    CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT_ENCODING_OPTIONS = 112
};

/* public_types.h:90-187: all enumerations are 32-bit integers on the ABI */
typedef int32_t charls_interleave_mode;        /* 0 none, 1 line, 2 sample                  (public_types.h:90-95)   */
typedef int32_t charls_compressed_data_format; /* 0 unknown, 1 interchange, 2 abbreviated image data, 3 abbreviated
                                                  table specification                        (public_types.h:97-103)  */
typedef uint32_t charls_encoding_options;      /* bit flags: 1 even_destination_size, 2 include_version_number,
                                                  4 include_pc_parameters_jai                (public_types.h:105-111) */
typedef int32_t charls_color_transformation;   /* 0 none, 1 HP1, 2 HP2, 3 HP3               (public_types.h:113-119) */
typedef int32_t charls_spiff_profile_id;       /* public_types.h:121-128 */
typedef int32_t charls_spiff_color_space;      /* public_types.h:130-145 */
typedef int32_t charls_spiff_compression_type; /* public_types.h:147-156 */
typedef int32_t charls_spiff_resolution_units; /* public_types.h:158-163 */

#define CHARLS_MAPPING_TABLE_MISSING (-1) /* public_types.h:184-187 */

/* public_types.h:934-950, 40 bytes */
typedef struct charls_spiff_header
{
    charls_spiff_profile_id profile_id;
    int32_t component_count;
    uint32_t height;
    uint32_t width;
    charls_spiff_color_space color_space;
    int32_t bits_per_sample;
    charls_spiff_compression_type compression_type;
    charls_spiff_resolution_units resolution_units;
    uint32_t vertical_resolution;
    uint32_t horizontal_resolution;
} charls_spiff_header;

/* public_types.h:988-1001, 16 bytes */
typedef struct charls_frame_info
{
    uint32_t width;          /* 1..100000  (src/constants.hpp:22-25) */
    uint32_t height;         /* 1..100000 */
    int32_t bits_per_sample; /* 2..16 */
    int32_t component_count; /* 1..255 */
} charls_frame_info;

/* public_types.h:1008-1020, 20 bytes; 0 = use the ISO/IEC 14495-1 default */
typedef struct charls_jpegls_pc_parameters
{
    int32_t maximum_sample_value;
    int32_t threshold1;
    int32_t threshold2;
    int32_t threshold3;
    int32_t reset_value;
} charls_jpegls_pc_parameters;

/* public_types.h:1027-1034, 12 bytes */
typedef struct charls_mapping_table_info
{
    int32_t table_id;
    int32_t entry_size;
    uint32_t data_size;
} charls_mapping_table_info;

/* public_types.h:1037-1043: non-zero return aborts decoding with callback_failed */
typedef int32_t (*charls_at_comment_handler)(const void* data, size_t size, void* user_context);
typedef int32_t (*charls_at_application_data_handler)(int32_t application_data_id, const void* data, size_t size,
                                                      void* user_context);

typedef struct charls_jpegls_encoder charls_jpegls_encoder;
typedef struct charls_jpegls_decoder charls_jpegls_decoder;

/* ------------------------------------------------------------------------------------------------------------------ */
/* Part 1a: encoder (reference include/charls/charls_jpegls_encoder.h)                                                 */
/* ------------------------------------------------------------------------------------------------------------------ */

/* charls_jpegls_encoder.h:24-25 -- returns NULL when out of memory; destroy(NULL) is a no-op */
CHARLS_B200_API charls_jpegls_encoder* charls_jpegls_encoder_create(void) CHARLS_B200_NOEXCEPT;
CHARLS_B200_API void charls_jpegls_encoder_destroy(const charls_jpegls_encoder* encoder) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:41-43 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_set_frame_info(charls_jpegls_encoder* encoder,
                                                                        const charls_frame_info* frame_info) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:51-52 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_set_near_lossless(charls_jpegls_encoder* encoder,
                                                                           int32_t near_lossless) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:60-62 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_set_encoding_options(charls_jpegls_encoder* encoder,
                                                                              charls_encoding_options encoding_options) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:71-73 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_set_interleave_mode(charls_jpegls_encoder* encoder,
                                                                             charls_interleave_mode interleave_mode) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:84-87 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_set_preset_coding_parameters(
    charls_jpegls_encoder* encoder, const charls_jpegls_pc_parameters* preset_coding_parameters) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:98-100 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_set_color_transformation(
    charls_jpegls_encoder* encoder, charls_color_transformation color_transformation) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:109-111 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_set_mapping_table_id(charls_jpegls_encoder* encoder,
                                                                              int32_t component_index, int32_t table_id) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:122-124.  Reference: raw + raw/16 + 1024 + 34 (src/charls_jpegls_encoder.cpp:104-114);
   this library adds the restart-marker overhead of the configured restart interval. */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_get_estimated_destination_size(const charls_jpegls_encoder* encoder,
                                                                                        size_t* size_in_bytes) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:134-138 -- the buffer is borrowed until the encoder is destroyed or rewound */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_set_destination_buffer(charls_jpegls_encoder* encoder,
                                                                                void* destination_buffer,
                                                                                size_t destination_size_bytes) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:151-156 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_write_standard_spiff_header(
    charls_jpegls_encoder* encoder, charls_spiff_color_space color_space, charls_spiff_resolution_units resolution_units,
    uint32_t vertical_resolution, uint32_t horizontal_resolution) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:165-167 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_write_spiff_header(charls_jpegls_encoder* encoder,
                                                                            const charls_spiff_header* spiff_header) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:180-184 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_write_spiff_entry(charls_jpegls_encoder* encoder, uint32_t entry_tag,
                                                                           const void* entry_data,
                                                                           size_t entry_data_size_bytes) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:196-197 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_write_spiff_end_of_directory_entry(charls_jpegls_encoder* encoder) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:209-213 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_write_comment(charls_jpegls_encoder* encoder, const void* comment,
                                                                       size_t comment_size_bytes) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:226-230 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_write_application_data(charls_jpegls_encoder* encoder,
                                                                                int32_t application_data_id,
                                                                                const void* application_data,
                                                                                size_t application_data_size_bytes) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:245-249 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_write_mapping_table(charls_jpegls_encoder* encoder, int32_t table_id,
                                                                             int32_t entry_size, const void* table_data,
                                                                             size_t table_data_size_bytes) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:262-266 -- HOT ENTRY POINT.  source_buffer is HOST memory (pinned memory is copied by DMA);
   stride in bytes, 0 = tightly packed. Replaces scan_encoder::encode_scan (src/scan_encoder_impl.hpp:42-49). */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_encode_from_buffer(charls_jpegls_encoder* encoder,
                                                                            const void* source_buffer,
                                                                            size_t source_size_bytes, uint32_t stride) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:282-287 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_encode_components_from_buffer(
    charls_jpegls_encoder* encoder, const void* source_buffer, size_t source_size_bytes, int32_t source_component_count,
    uint32_t stride) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:296-297 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_create_abbreviated_format(charls_jpegls_encoder* encoder) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:305-307 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_get_bytes_written(const charls_jpegls_encoder* encoder,
                                                                           size_t* bytes_written) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_encoder.h:315-316 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_encoder_rewind(charls_jpegls_encoder* encoder) CHARLS_B200_NOEXCEPT;

/* ------------------------------------------------------------------------------------------------------------------ */
/* Part 1b: decoder (reference include/charls/charls_jpegls_decoder.h)                                                 */
/* ------------------------------------------------------------------------------------------------------------------ */

/* charls_jpegls_decoder.h:24-25 */
CHARLS_B200_API charls_jpegls_decoder* charls_jpegls_decoder_create(void) CHARLS_B200_NOEXCEPT;
CHARLS_B200_API void charls_jpegls_decoder_destroy(const charls_jpegls_decoder* decoder) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:43-47 -- the buffer is borrowed until the decoder is destroyed */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_set_source_buffer(charls_jpegls_decoder* decoder,
                                                                           const void* source_buffer,
                                                                           size_t source_size_bytes) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:58-61 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_read_spiff_header(charls_jpegls_decoder* decoder,
                                                                           charls_spiff_header* spiff_header,
                                                                           int32_t* header_found) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:68-69 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_read_header(charls_jpegls_decoder* decoder) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:80-82 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_get_frame_info(const charls_jpegls_decoder* decoder,
                                                                        charls_frame_info* frame_info) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:94-96 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_get_near_lossless(const charls_jpegls_decoder* decoder,
                                                                           int32_t component_index, int32_t* near_lossless) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:108-110 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_get_interleave_mode(const charls_jpegls_decoder* decoder,
                                                                             int32_t component_index,
                                                                             charls_interleave_mode* interleave_mode) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:122-125 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_get_preset_coding_parameters(
    const charls_jpegls_decoder* decoder, int32_t reserved, charls_jpegls_pc_parameters* preset_coding_parameters) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:136-138 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_get_color_transformation(
    const charls_jpegls_decoder* decoder, charls_color_transformation* color_transformation) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:150-152 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_get_destination_size(const charls_jpegls_decoder* decoder, uint32_t stride,
                                                                              size_t* destination_size_bytes) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:168-172 -- HOT ENTRY POINT.  destination_buffer is HOST memory.
   Replaces scan_decoder::decode_scan (src/scan_decoder_impl.hpp:40-56). */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_decode_to_buffer(charls_jpegls_decoder* decoder, void* destination_buffer,
                                                                          size_t destination_size_bytes, uint32_t stride) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:185-187 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_at_comment(charls_jpegls_decoder* decoder, charls_at_comment_handler handler,
                                                                    void* user_context) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:200-202 */
CHARLS_B200_API charls_jpegls_errc charls_jpegls_decoder_at_application_data(charls_jpegls_decoder* decoder,
                                                                             charls_at_application_data_handler handler,
                                                                             void* user_context) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:214-216 */
CHARLS_B200_API charls_jpegls_errc charls_decoder_get_compressed_data_format(
    const charls_jpegls_decoder* decoder, charls_compressed_data_format* compressed_data_format) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:228-230 */
CHARLS_B200_API charls_jpegls_errc charls_decoder_get_mapping_table_id(const charls_jpegls_decoder* decoder, int32_t component_index,
                                                                       int32_t* table_id) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:243-245 */
CHARLS_B200_API charls_jpegls_errc charls_decoder_find_mapping_table_index(const charls_jpegls_decoder* decoder,
                                                                           int32_t mapping_table_id, int32_t* index) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:256-258 */
CHARLS_B200_API charls_jpegls_errc charls_decoder_get_mapping_table_count(const charls_jpegls_decoder* decoder, int32_t* count) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:272-274 */
CHARLS_B200_API charls_jpegls_errc charls_decoder_get_mapping_table_info(const charls_jpegls_decoder* decoder,
                                                                         int32_t mapping_table_index,
                                                                         charls_mapping_table_info* mapping_table_info) CHARLS_B200_NOEXCEPT;
/* charls_jpegls_decoder.h:289-293 */
CHARLS_B200_API charls_jpegls_errc charls_decoder_get_mapping_table_data(const charls_jpegls_decoder* decoder,
                                                                         int32_t mapping_table_index, void* mapping_table_data,
                                                                         size_t mapping_table_size_bytes) CHARLS_B200_NOEXCEPT;

/* ------------------------------------------------------------------------------------------------------------------ */
/* Part 1c: miscellaneous                                                                                              */
/* ------------------------------------------------------------------------------------------------------------------ */

/* include/charls/jpegls_error.h:12 */
CHARLS_B200_API const char* charls_get_error_message(charls_jpegls_errc error_value);
/* include/charls/jpegls_error.hpp:10 -- returns a `const std::error_category*` for C++ callers */
CHARLS_B200_API const void* charls_get_jpegls_category(void);
/* include/charls/version.h:27-38 */
CHARLS_B200_API const char* charls_get_version_string(void);
CHARLS_B200_API void charls_get_version_number(int32_t* major, int32_t* minor, int32_t* patch);
/* include/charls/validate_spiff_header.h:23-24 */
CHARLS_B200_API charls_jpegls_errc charls_validate_spiff_header(const charls_spiff_header* spiff_header,
                                                                const charls_frame_info* frame_info) CHARLS_B200_NOEXCEPT;

/* ------------------------------------------------------------------------------------------------------------------ */
/* Part 2: B200 extensions (no counterpart in the reference)                                                           */
/* ------------------------------------------------------------------------------------------------------------------ */

/* Number of CUDA devices visible to this process / device used by every later call of the calling process. */
CHARLS_B200_API charls_jpegls_errc charlsx_get_device_count(int32_t* count) CHARLS_B200_NOEXCEPT;
CHARLS_B200_API charls_jpegls_errc charlsx_set_device(int32_t device_ordinal) CHARLS_B200_NOEXCEPT;

/* Restart interval (in lines) the encoder writes: a DRI segment plus RSTm markers.  Default 1 (every line an
   independent work item, the data-parallel configuration); 0 = no restart markers, which makes the output byte-identical
   to the reference encoder's (and serial).  Must be called before the first encode call. */
CHARLS_B200_API charls_jpegls_errc charlsx_jpegls_encoder_set_restart_interval(charls_jpegls_encoder* encoder,
                                                                               uint32_t restart_interval) CHARLS_B200_NOEXCEPT;
/* Restart interval found in the DRI segment (0 = none); valid after read_header. */
CHARLS_B200_API charls_jpegls_errc charlsx_jpegls_decoder_get_restart_interval(const charls_jpegls_decoder* decoder,
                                                                               uint32_t* restart_interval) CHARLS_B200_NOEXCEPT;

/* Batch interface: all frames share geometry and coding parameters, samples and streams live in DEVICE memory. */
/* Two-part forms of the hot entry points (no counterpart in the reference, whose calls are synchronous: reference
 * src/charls_jpegls_encoder.cpp:285-296, src/charls_jpegls_decoder.cpp:177-201).  `_begin` validates like the one-part call,
 * writes / has read the marker segments and ISSUES the scan on the object's CUDA stream -- input copy, kernels, and for the
 * decoder the copy of the samples into destination_buffer -- without waiting; `_end` waits for it, moves the entropy-coded
 * bytes into the destination (encoder) and finishes the object exactly as the one-part call would have (same error codes,
 * same bytes).  Between the two the caller may start work on OTHER codec objects: that is how one host thread keeps several
 * images in flight.  Source and destination buffers must stay valid until `_end` returns; no other call on the object is
 * allowed in between (invalid_operation).  Frames that are coded as several scans (planar components) run to completion
 * inside `_begin`; `_end` is then a no-op. */
CHARLS_B200_API charls_jpegls_errc charlsx_jpegls_encoder_encode_from_buffer_begin(charls_jpegls_encoder* encoder,
                                                                                   const void* source_buffer,
                                                                                   size_t source_size_bytes,
                                                                                   uint32_t stride) CHARLS_B200_NOEXCEPT;
CHARLS_B200_API charls_jpegls_errc charlsx_jpegls_encoder_encode_end(charls_jpegls_encoder* encoder) CHARLS_B200_NOEXCEPT;
CHARLS_B200_API charls_jpegls_errc charlsx_jpegls_decoder_decode_to_buffer_begin(charls_jpegls_decoder* decoder,
                                                                                 void* destination_buffer,
                                                                                 size_t destination_size_bytes,
                                                                                 uint32_t stride) CHARLS_B200_NOEXCEPT;
CHARLS_B200_API charls_jpegls_errc charlsx_jpegls_decoder_decode_end(charls_jpegls_decoder* decoder) CHARLS_B200_NOEXCEPT;

/* Side table of interval offsets (extension).  With it the encoder writes, in front of every scan's SOS, one or more
 * APP11 segments ("JLS-OFFT", charls_b200/csrc/jls_common.h) that list where every restart interval starts.  The reference --
 * like every other JPEG-LS decoder -- treats them as application data and skips them (reference
 * src/jpeg_stream_reader.cpp:442-457); this library's decoders use the table instead of searching the stream for restart
 * markers (and decode as if there were none when the stream does not agree with it), and a rank that decodes only some
 * lines of a frame knows which bytes it needs.  Off by default on the charls_* objects (their output then differs from a
 * plain DRI stream by nothing); requires a restart interval. */
CHARLS_B200_API charls_jpegls_errc charlsx_jpegls_encoder_set_offset_table(charls_jpegls_encoder* encoder,
                                                                           int32_t enabled) CHARLS_B200_NOEXCEPT;

typedef struct charlsx_batch_image
{
    void* pixels;           /* device: samples of the frame (encode: input, decode: output), 16-byte aligned */
    void* stream;           /* device: complete JPEG-LS stream (encode: output, decode: input).  The decoder loads the stream in
                             * aligned 32-bit words: up to three bytes on either side of it, inside the words that hold its first
                             * and last byte, are read and ignored. */
    size_t stream_capacity; /* encode: capacity of `stream`; decode: size of the stream in bytes */
    size_t stream_size;     /* encode: bytes written (out) */
    int32_t status;         /* charls_jpegls_errc of this frame (out) */
    int32_t reserved;
} charlsx_batch_image;

typedef struct charlsx_batch_params
{
    charls_frame_info frame_info;
    int32_t near_lossless;
    charls_interleave_mode interleave_mode; /* none requires component_count == 1 in the batch interface */
    charls_color_transformation color_transformation;
    uint32_t restart_interval; /* encode: interval to write (1 = per line); decode: ignored (read from each stream) */
    uint32_t stride;           /* bytes between lines of `pixels`, 0 = tightly packed.  Device frames: a multiple of 4 (and
                                * 4-byte aligned frames) saves a copy -- other rows are moved to an aligned pitch in device scratch
                                * memory (one more frame's worth per frame) in front of the encoder and back behind the decoder;
                                * host frames are re-pitched on their way to the device. */
    uint32_t flags;            /* CHARLSX_BATCH_* */
} charlsx_batch_params;

/* encode: write the side table of interval offsets (see charlsx_jpegls_encoder_set_offset_table); decode: the streams may
 * carry one (the headers are then read far enough to find it; streams without a table decode as always) */
#define CHARLSX_BATCH_OFFSET_TABLE 1U

typedef struct charlsx_batch charlsx_batch;

CHARLS_B200_API charlsx_batch* charlsx_batch_create(void) CHARLS_B200_NOEXCEPT;
CHARLS_B200_API void charlsx_batch_destroy(charlsx_batch* batch) CHARLS_B200_NOEXCEPT;
/* Encodes / decodes `count` frames.  `cuda_stream` is a cudaStream_t (NULL = the batch object's own stream); the call
   returns after the work has completed and per-frame status / stream_size have been written.  The return value is the
   first per-frame error (or an argument error). */
CHARLS_B200_API charls_jpegls_errc charlsx_batch_encode(charlsx_batch* batch, const charlsx_batch_params* params,
                                                        charlsx_batch_image* images, size_t count, void* cuda_stream) CHARLS_B200_NOEXCEPT;
CHARLS_B200_API charls_jpegls_errc charlsx_batch_decode(charlsx_batch* batch, const charlsx_batch_params* params,
                                                        charlsx_batch_image* images, size_t count, void* cuda_stream) CHARLS_B200_NOEXCEPT;
/* The same for frames whose samples and streams live in HOST memory (`pixels` / `stream` are host pointers; page-locked
   memory gives full PCIe speed): the library stages chunks of frames in device memory and overlaps the copies of one
   chunk with the kernels of its neighbours, so one host thread keeps the PCIe link busy.  The reference's way to code many
   images is one codec object and one call per image (include/charls/charls_jpegls_encoder.h:263-266,
   charls_jpegls_decoder.h:168-172); these two calls replace such a loop. */
CHARLS_B200_API charls_jpegls_errc charlsx_batch_encode_host(charlsx_batch* batch, const charlsx_batch_params* params,
                                                             charlsx_batch_image* images, size_t count) CHARLS_B200_NOEXCEPT;
CHARLS_B200_API charls_jpegls_errc charlsx_batch_decode_host(charlsx_batch* batch, const charlsx_batch_params* params,
                                                             charlsx_batch_image* images, size_t count) CHARLS_B200_NOEXCEPT;
/* Kernels launched by the last charlsx_batch_encode / _decode call on this object. */
CHARLS_B200_API charls_jpegls_errc charlsx_batch_get_last_kernel_launches(const charlsx_batch* batch, uint32_t* launches) CHARLS_B200_NOEXCEPT;
/* Device time (CUDA events on the launching stream, recorded directly around it) of the entropy-coding kernel of the
   last charlsx_batch_encode / _decode call on this object, in milliseconds. */
CHARLS_B200_API charls_jpegls_errc charlsx_batch_get_last_coder_kernel_ms(const charlsx_batch* batch, float* milliseconds) CHARLS_B200_NOEXCEPT;
/* Kernels launched by this library in this process so far. */
CHARLS_B200_API charls_jpegls_errc charlsx_get_kernel_launch_count(uint64_t* launches) CHARLS_B200_NOEXCEPT;

#ifdef __cplusplus
}
#endif

#endif /* CHARLS_B200_H */
