// misc.cpp -- error messages, version, SPIFF header validation, device selection (C ABI).
//
// Counterparts of reference src/jpegls_error.cpp:14-210, src/version.cpp:12-37 and src/validate_spiff_header.cpp:12-107.
#include "../engine.hpp"
#include "../jls_kernels.hpp"
#include "abi_support.hpp"

#include <string>
#include <system_error>

using namespace jls::host;

namespace {

class JpegLsCategory final : public std::error_category
{
public:
    const char* name() const noexcept override { return "charls::jpegls"; }
    std::string message(int error_value) const override { return charls_get_error_message(error_value); }
};

bool valid_color_space(int32_t color_space, int32_t component_count) noexcept
{
    switch (color_space)
    {
    case 2: // none
        return true;
    case 0:  // bi-level black
    case 15: // bi-level white: not meaningful for JPEG-LS
        return false;
    case 8: // grayscale
        return component_count == 1;
    case 1: case 3: case 4: case 9: case 10: case 11: case 14: // YCbCr variants, PhotoYCC, RGB, CMY, CIELab
        return component_count == 3;
    case 12: case 13: // CMYK, YCCK
        return component_count == 4;
    default:
        return false;
    }
}

} // namespace

extern "C" {

const char* charls_get_error_message(charls_jpegls_errc error_value)
{
    switch (error_value)
    {
    case 0: return "Success";
    case 1: return "No memory could be allocated for an internal buffer";
    case 2: return "Callback function returned a failure";
    case 3: return "The destination buffer is too small to hold all the output";
    case 4: return "The source buffer is too small, more input data was expected";
    case 5: return "Invalid JPEG-LS stream, the encoded bit stream contains a general structural problem";
    case 6: return "Invalid JPEG-LS stream: the JPEG stream is not encoded with the JPEG-LS algorithm";
    case 7: return "The JPEG-LS stream is encoded with a parameter value that is not supported by the this decoder";
    case 8: return "The HP color transform is not supported";
    case 9: return "Unsupported JPEG-LS stream: JPEG-LS preset parameters segment contains a JPEG-LS Extended (ISO/IEC 14495-2) type";
    case 10: return "Invalid JPEG-LS stream: the leading start byte (0xFF) for a JPEG marker was not found";
    case 11: return "Invalid JPEG-LS stream: first JPEG marker is not a Start Of Image (SOI) marker";
    case 12: return "Invalid JPEG-LS stream: invalid SPIFF header";
    case 13: return "Invalid JPEG-LS stream: an unknown JPEG marker code was found";
    case 14: return "Invalid JPEG-LS stream: unexpected Start Of Scan (SOS) marker found";
    case 15: return "Invalid JPEG-LS stream: segment size of a marker segment is invalid";
    case 16: return "Invalid JPEG-LS stream: more then one Start Of Image (SOI) marker";
    case 17: return "Invalid JPEG-LS stream: more then one Start Of Frame (SOF) marker";
    case 18: return "Invalid JPEG-LS stream: duplicate component identifier in the (SOF) segment";
    case 19: return "Invalid JPEG-LS stream: unexpected End Of Image (EOI) marker";
    case 20: return "Invalid JPEG-LS stream: JPEG-LS preset parameters segment contains an invalid type";
    case 21: return "Invalid JPEG-LS stream: SPIFF header without End Of Directory (EOD) entry";
    case 22: return "Invalid JPEG-LS stream: restart (RTSm) marker found outside encoded entropy data";
    case 23: return "Invalid JPEG-LS stream: missing expected restart (RTSm) marker";
    case 24: return "Invalid JPEG-LS stream: missing End Of Image (EOI) marker";
    case 25: return "Invalid JPEG-LS stream: unexpected Define Number of Lines (DNL) marker";
    case 26: return "Invalid JPEG-LS stream: missing expected Define Number of Lines (DNL) marker";
    case 27: return "Invalid JPEG-LS stream: unknown component ID in scan segment";
    case 28: return "Invalid JPEG-LS stream: mapping tables without SOF but with spiff header";
    case 29: return "Invalid JPEG-LS stream: the width (Number of samples per line) is already defined";
    case 30: return "Invalid JPEG-LS stream: the height (Number of lines) is already defined";
    case 31: return "Invalid JPEG-LS stream: the bit per sample (sample precision) parameter is not in the range [2, 16]";
    case 32: return "Invalid JPEG-LS stream: component count in the SOF segment is outside the range [1, 255]";
    case 33: return "Invalid JPEG-LS stream: interleave mode is outside the range [0, 2] or conflicts with component count";
    case 34: return "Invalid JPEG-LS stream: near-lossless is outside the range [0, min(255, MAXVAL/2)]";
    case 35: return "Invalid JPEG-LS stream: JPEG-LS preset parameters segment contains invalid values";
    case 36: return "Invalid JPEG-LS stream: Color transformation segment contains invalid values or frame info mismatch";
    case 37: return "Invalid JPEG-LS stream: mapping table ID outside valid range or duplicate";
    case 38: return "Invalid JPEG-LS stream: mapping table continuation without matching mapping table specification";
    case 100: return "Method call is invalid for the current state";
    case 101: return "Invalid argument";
    case 102: return "The width argument is outside the supported range [1, 4294967295]";
    case 103: return "The height argument is outside the supported range [1, 4294967295]";
    case 104: return "The bit per sample argument is outside the range [2, 16]";
    case 105: return "The component count argument is outside the range [1, 255]";
    case 106: return "The interleave mode is not None, Sample, Line or invalid in combination with component count";
    case 107: return "The near lossless argument is outside the range [0, min(255, MAXVAL/2)]";
    case 108: return "The argument for the JPEG-LS preset coding parameters is not valid";
    case 109: return "The argument for the color component is not (None, Hp1, Hp2, Hp3) or invalid in combination with component count";
    case 110: return "The passed size is outside the valid range";
    case 111: return "The stride argument does not match with the frame info and buffer size";
    case 112: return "The encoding options argument has an invalid value";
    case jls::errc_device_failure: return "CUDA device failure (charls_b200 extension; details were written to stderr)";
    default: return "Unknown";
    }
}

const void* charls_get_jpegls_category(void)
{
    static const JpegLsCategory instance;
    return static_cast<const std::error_category*>(&instance);
}

const char* charls_get_version_string(void)
{
    return "3.0.0"; // ABI level of the reference this library stands in for (reference include/charls/version.h:14-16)
}

void charls_get_version_number(int32_t* major, int32_t* minor, int32_t* patch)
{
    if (major)
        *major = 3;
    if (minor)
        *minor = 0;
    if (patch)
        *patch = 0;
}

charls_jpegls_errc charls_validate_spiff_header(const charls_spiff_header* spiff_header, const charls_frame_info* frame_info) noexcept
{
    return guarded([&] {
        const charls_spiff_header& h = *check_pointer(spiff_header);
        const charls_frame_info& f = *check_pointer(frame_info);
        const bool valid = h.compression_type == 6 && h.profile_id == 0 && h.resolution_units >= 0 && h.resolution_units <= 2 &&
                           h.horizontal_resolution != 0 && h.vertical_resolution != 0 && h.component_count == f.component_count &&
                           valid_color_space(h.color_space, h.component_count) && h.bits_per_sample == f.bits_per_sample &&
                           h.height == f.height && h.width == f.width;
        if (!valid)
            fail(CHARLS_JPEGLS_ERRC_INVALID_SPIFF_HEADER);
    });
}

charls_jpegls_errc charlsx_get_device_count(int32_t* count) noexcept
{
    return guarded([&] { check_status(jls::device_count(check_pointer(count))); });
}

charls_jpegls_errc charlsx_set_device(int32_t device_ordinal) noexcept
{
    return guarded([&] { check_status(jls::set_device(device_ordinal)); });
}

charls_jpegls_errc charlsx_get_kernel_launch_count(uint64_t* launches) noexcept
{
    return guarded([&] { *check_pointer(launches) = jls::kernel_launch_count(); });
}

} // extern "C"
