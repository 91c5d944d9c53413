// A CharLS user's program: written against the REFERENCE's own header-only C++ wrapper (include/charls/charls.hpp), not
// against anything of this repository.  tests/test_zz_abi_consumer.py compiles it with -I/root/reference/include and
// links it with charls_b200/lib/libcharls.so.3 -- the drop-in claim of DESIGN.md section 1 ("what calls it").
//   consumer header <file.jls>   host-only calls (no GPU needed): version, header parsing, sizes, error reporting
//   consumer roundtrip           encode + decode through charls::jpegls_encoder / jpegls_decoder (needs the GPU)
#include <charls/charls.hpp>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

namespace {

int fail(const char* what)
{
    std::fprintf(stderr, "consumer: %s\n", what);
    return 1;
}

int header_mode(const char* path)
{
    std::ifstream file(path, std::ios::binary);
    const std::vector<uint8_t> stream((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
    if (stream.empty())
        return fail("cannot read the stream");
    int32_t major = 0, minor = 0, patch = 0;
    charls_get_version_number(&major, &minor, &patch);
    if (major != 3 || std::strlen(charls_get_version_string()) == 0)
        return fail("version");

    charls::jpegls_decoder decoder{stream, true};
    const charls::frame_info info = decoder.frame_info();
    std::printf("%u %u %d %d %d %d %zu\n", info.width, info.height, info.bits_per_sample, info.component_count,
                decoder.get_near_lossless(), static_cast<int>(decoder.get_interleave_mode()), decoder.get_destination_size());

    charls::jpegls_encoder encoder;
    encoder.frame_info({info.width, info.height, info.bits_per_sample, info.component_count});
    if (encoder.estimated_destination_size() == 0)
        return fail("estimated size");

    // errors travel as charls::jpegls_error with the reference's codes and messages
    const std::vector<uint8_t> garbage{0x33, 0x33, 0x33, 0x33};
    try
    {
        charls::jpegls_decoder broken{garbage, true};
        return fail("garbage accepted");
    }
    catch (const charls::jpegls_error& error)
    {
        if (error.code() != charls::jpegls_errc::jpeg_marker_start_byte_not_found || std::strlen(error.what()) == 0)
            return fail("error code");
    }
    return 0;
}

int roundtrip_mode()
{
    const charls::frame_info frame{200, 120, 8, 3};
    std::vector<uint8_t> pixels(static_cast<size_t>(frame.width) * frame.height * 3);
    for (size_t i = 0; i < pixels.size(); ++i)
        pixels[i] = static_cast<uint8_t>((i * 7 + (i >> 9) * 13) & 0xFF);

    const std::vector<uint8_t> encoded =
        charls::jpegls_encoder::encode(pixels, frame, charls::interleave_mode::sample, charls::encoding_options::none);
    std::vector<uint8_t> decoded;
    const auto [info, mode] = charls::jpegls_decoder::decode(encoded, decoded);
    if (info.width != frame.width || info.height != frame.height || info.component_count != 3 ||
        mode != charls::interleave_mode::sample)
        return fail("frame info after the round trip");
    if (decoded != pixels)
        return fail("pixels after the round trip");
    std::printf("%zu\n", encoded.size());
    return 0;
}

} // namespace

int main(int argc, char** argv)
{
    try
    {
        if (argc == 3 && std::strcmp(argv[1], "header") == 0)
            return header_mode(argv[2]);
        if (argc == 2 && std::strcmp(argv[1], "roundtrip") == 0)
            return roundtrip_mode();
    }
    catch (const charls::jpegls_error& error)
    {
        std::fprintf(stderr, "consumer: jpegls_error %d: %s\n", static_cast<int>(error.code().value()), error.what());
        return 2;
    }
    return fail("usage: consumer header <file.jls> | consumer roundtrip");
}
