// jls_params.hpp -- host-side derivation of the coding parameters (no CUDA needed).
//
// Restates the reference's parameter logic: default preset coding parameters (ISO/IEC 14495-1 C.2.4.1.1.1; reference
// src/jpegls_preset_coding_parameters.hpp:15-130) and the traits values of src/default_traits.hpp:51-59.
#pragma once

#include "jls_common.h"

#include <algorithm>

namespace jls {

struct PresetCodingParameters
{
    int32_t maximum_sample_value;
    int32_t threshold1;
    int32_t threshold2;
    int32_t threshold3;
    int32_t reset_value;
};

inline int32_t log2_ceiling(int32_t n) noexcept
{
    int32_t x = 0;
    while (n > (int32_t{1} << x))
        ++x;
    return x;
}

inline int32_t maximum_bit_sample_value(int32_t bits_per_sample) noexcept
{
    return static_cast<int32_t>((1U << bits_per_sample) - 1U);
}

inline int32_t maximum_near_lossless(int32_t maximum_sample_value) noexcept
{
    return std::min(255, maximum_sample_value / 2); // ISO/IEC 14495-1 C.2.3
}

// Figure C.3 clamp
inline int32_t clamp_threshold(int32_t i, int32_t j, int32_t maximum_sample_value) noexcept
{
    return (i > maximum_sample_value || i < j) ? j : i;
}

inline PresetCodingParameters default_preset_parameters(int32_t maximum_sample_value, int32_t near_lossless) noexcept
{
    constexpr int32_t basic_t1 = 3, basic_t2 = 7, basic_t3 = 21, default_reset = 64;
    PresetCodingParameters d{};
    d.maximum_sample_value = maximum_sample_value;
    d.reset_value = default_reset;
    if (maximum_sample_value >= 128)
    {
        const int32_t factor = (std::min(maximum_sample_value, 4095) + 128) / 256;
        d.threshold1 = clamp_threshold(factor * (basic_t1 - 2) + 2 + 3 * near_lossless, near_lossless + 1, maximum_sample_value);
        d.threshold2 = clamp_threshold(factor * (basic_t2 - 3) + 3 + 5 * near_lossless, d.threshold1, maximum_sample_value);
        d.threshold3 = clamp_threshold(factor * (basic_t3 - 4) + 4 + 7 * near_lossless, d.threshold2, maximum_sample_value);
        return d;
    }
    const int32_t factor = 256 / (maximum_sample_value + 1);
    d.threshold1 = clamp_threshold(std::max(2, basic_t1 / factor + 3 * near_lossless), near_lossless + 1, maximum_sample_value);
    d.threshold2 = clamp_threshold(std::max(3, basic_t2 / factor + 5 * near_lossless), d.threshold1, maximum_sample_value);
    d.threshold3 = clamp_threshold(std::max(4, basic_t3 / factor + 7 * near_lossless), d.threshold2, maximum_sample_value);
    return d;
}

inline bool is_default_preset(const PresetCodingParameters& pc, const PresetCodingParameters& defaults) noexcept
{
    if (pc.maximum_sample_value == 0 && pc.threshold1 == 0 && pc.threshold2 == 0 && pc.threshold3 == 0 && pc.reset_value == 0)
        return true;
    return pc.maximum_sample_value == defaults.maximum_sample_value && pc.threshold1 == defaults.threshold1 &&
           pc.threshold2 == defaults.threshold2 && pc.threshold3 == defaults.threshold3 &&
           pc.reset_value == defaults.reset_value;
}

// Table C.1 validity; zero entries mean "use the default".  On success `validated` holds the effective values.
inline bool validate_preset_parameters(const PresetCodingParameters& pc, int32_t maximum_bit_value, int32_t near_lossless,
                                       PresetCodingParameters* validated) noexcept
{
    if (pc.maximum_sample_value != 0 && (pc.maximum_sample_value < 1 || pc.maximum_sample_value > maximum_bit_value))
        return false;
    const int32_t maxval = pc.maximum_sample_value != 0 ? pc.maximum_sample_value : maximum_bit_value;
    if (pc.threshold1 != 0 && (pc.threshold1 < near_lossless + 1 || pc.threshold1 > maxval))
        return false;
    const PresetCodingParameters d = default_preset_parameters(maxval, near_lossless);
    const int32_t t1 = pc.threshold1 != 0 ? pc.threshold1 : d.threshold1;
    if (pc.threshold2 != 0 && (pc.threshold2 < t1 || pc.threshold2 > maxval))
        return false;
    const int32_t t2 = pc.threshold2 != 0 ? pc.threshold2 : d.threshold2;
    if (pc.threshold3 != 0 && (pc.threshold3 < t2 || pc.threshold3 > maxval))
        return false;
    if (pc.reset_value != 0 && (pc.reset_value < 3 || pc.reset_value > std::max(255, maxval)))
        return false;
    if (validated)
    {
        validated->maximum_sample_value = maxval;
        validated->threshold1 = t1;
        validated->threshold2 = t2;
        validated->threshold3 = pc.threshold3 != 0 ? pc.threshold3 : d.threshold3;
        validated->reset_value = pc.reset_value != 0 ? pc.reset_value : d.reset_value;
    }
    return true;
}

// All derived values the kernels need for one scan.  `pc` must be validated (non-zero thresholds / reset).
inline CodecParams make_codec_params(int32_t width, int32_t height, int32_t bits_per_sample, int32_t components_in_scan,
                                     int32_t near_lossless, int32_t interleave, int32_t transform,
                                     const PresetCodingParameters& pc, uint32_t restart_interval) noexcept
{
    CodecParams p{};
    p.width = width;
    p.height = height;
    p.components = components_in_scan;
    p.interleave = interleave;
    p.transform = transform;
    p.bits_per_sample = bits_per_sample;
    p.sample_bytes = bits_per_sample > 8 ? 2 : 1;
    p.near = near_lossless;
    p.maxval = maximum_bit_sample_value(bits_per_sample); // never the LSE MAXVAL: reference src/make_scan_codec.cpp:98
    p.dq = 2 * near_lossless + 1;
    p.range = (p.maxval + 2 * near_lossless) / p.dq + 1;
    p.qbpp = log2_ceiling(p.range);
    p.bpp = log2_ceiling(p.maxval);
    p.limit = 2 * (p.bpp + std::max(8, p.bpp));
    p.t1 = pc.threshold1;
    p.t2 = pc.threshold2;
    p.t3 = pc.threshold3;
    p.reset = pc.reset_value & 0xFF; // reference src/scan_codec.hpp:129 stores RESET in a uint8_t
    p.a_init = std::max(2, (p.range + 32) / 64);
    p.dq_magic = p.dq == 1 ? 0xFFFFFFFFU : static_cast<uint32_t>((uint64_t{1} << 32) / static_cast<uint32_t>(p.dq)) + 1U;
    p.range_dq = p.range * p.dq;
    p.restart_interval = restart_interval;
    p.lines_per_interval = restart_interval == 0 ? static_cast<uint32_t>(height)
                                                 : std::min(restart_interval, static_cast<uint32_t>(height));
    p.interval_count = (static_cast<uint32_t>(height) + p.lines_per_interval - 1) / p.lines_per_interval;
    return p;
}

} // namespace jls
