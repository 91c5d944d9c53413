// abi_support.hpp -- error transport and argument checks shared by the C ABI entry points.
//
// Like the reference (src/util.hpp:84-101,226-260) errors travel as C++ exceptions inside the library and are turned
// into charls_jpegls_errc values at the `extern "C"` boundary; nothing ever propagates into the caller.
#pragma once

#include "charls_b200.h"

#include <cstddef>
#include <cstdint>
#include <new>

namespace jls::host {

struct Failure final
{
    charls_jpegls_errc errc;
};

[[noreturn]] inline void fail(charls_jpegls_errc errc)
{
    throw Failure{errc};
}

inline void check_operation(bool ok)
{
    if (!ok)
        fail(CHARLS_JPEGLS_ERRC_INVALID_OPERATION);
}

inline void check_argument(bool ok, charls_jpegls_errc errc = CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT)
{
    if (!ok)
        fail(errc);
}

template<typename T>
void check_range(T minimum, T maximum, T value, charls_jpegls_errc errc = CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT)
{
    if (!(minimum <= value && value <= maximum))
        fail(errc);
}

template<typename T>
T* check_pointer(T* pointer)
{
    if (!pointer)
        fail(CHARLS_JPEGLS_ERRC_INVALID_ARGUMENT);
    return pointer;
}

inline void check_buffer(const void* data, size_t size)
{
    check_argument(data != nullptr || size == 0);
}

inline void check_status(int32_t errc)
{
    if (errc != 0)
        fail(errc);
}

inline size_t checked_mul(size_t a, size_t b)
{
    size_t result;
    if (__builtin_mul_overflow(a, b, &result))
        fail(CHARLS_JPEGLS_ERRC_PARAMETER_VALUE_NOT_SUPPORTED); // reference src/util.hpp:352-381
    return result;
}

inline size_t add_saturated(size_t a, size_t b) noexcept
{
    const size_t r = a + b;
    return r < a ? static_cast<size_t>(-1) : r;
}

// Runs `body`, mapping every failure to an error code (reference src/util.hpp:84-101).
template<typename Body>
charls_jpegls_errc guarded(Body&& body) noexcept
{
    try
    {
        body();
        return CHARLS_JPEGLS_ERRC_SUCCESS;
    }
    catch (const Failure& failure)
    {
        return failure.errc;
    }
    catch (const std::bad_alloc&)
    {
        return CHARLS_JPEGLS_ERRC_NOT_ENOUGH_MEMORY;
    }
    catch (...)
    {
        return CHARLS_JPEGLS_ERRC_NOT_ENOUGH_MEMORY;
    }
}

// limits (reference src/constants.hpp:14-52)
constexpr uint32_t maximum_width = 100000;
constexpr uint32_t maximum_height = 100000;
constexpr int32_t maximum_component_count = 255;
constexpr int32_t maximum_component_count_in_scan = 4;
constexpr size_t segment_max_data_size = 65535 - 2;
constexpr size_t spiff_entry_max_data_size = 65528;
constexpr size_t spiff_header_size_in_bytes = 34;

// marker codes (reference src/jpeg_marker_code.hpp:20-100)
constexpr uint8_t marker_start = 0xFF;
constexpr uint8_t marker_soi = 0xD8, marker_eoi = 0xD9, marker_sos = 0xDA, marker_dnl = 0xDC, marker_dri = 0xDD;
constexpr uint8_t marker_app0 = 0xE0, marker_app8 = 0xE8, marker_app15 = 0xEF, marker_com = 0xFE;
constexpr uint8_t marker_sof55 = 0xF7, marker_lse = 0xF8;

} // namespace jls::host
