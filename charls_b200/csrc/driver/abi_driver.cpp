// abi_driver.cpp -- drives a CharLS-compatible C ABI from plain C++ threads, the way a C/C++ application would.
//
// bench.py's end-to-end leg and tools/e2e_probe.py call this through ctypes once per step: Python threads spend their
// time handing the interpreter lock around the ~40 short ABI calls of one image round trip, which measures the
// interpreter rather than the library.  The driver binds the library with dlopen/dlsym (the reference's own
// include/charls/charls_jpegls_encoder.h:38-264, charls_jpegls_decoder.h:41-300 entry points, nothing else), so the
// same code runs against the B200 library and against the reference's libcharls.
#include <atomic>
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <dlfcn.h>
#include <thread>
#include <vector>

namespace {

struct FrameInfo
{
    uint32_t width, height;
    int32_t bits_per_sample, component_count;
};

struct Abi
{
    void* (*encoder_create)();
    void (*encoder_destroy)(void*);
    int32_t (*encoder_set_frame_info)(void*, const FrameInfo*);
    int32_t (*encoder_set_near_lossless)(void*, int32_t);
    int32_t (*encoder_set_interleave_mode)(void*, int32_t);
    int32_t (*encoder_set_color_transformation)(void*, int32_t);
    int32_t (*encoder_set_destination_buffer)(void*, void*, size_t);
    int32_t (*encoder_encode_from_buffer)(void*, const void*, size_t, uint32_t);
    int32_t (*encoder_get_bytes_written)(void*, size_t*);
    void* (*decoder_create)();
    void (*decoder_destroy)(void*);
    int32_t (*decoder_set_source_buffer)(void*, const void*, size_t);
    int32_t (*decoder_read_header)(void*);
    int32_t (*decoder_decode_to_buffer)(void*, void*, size_t, uint32_t);
};

template<typename F>
bool bind(void* handle, const char* name, F& out)
{
    out = reinterpret_cast<F>(dlsym(handle, name));
    return out != nullptr;
}

bool load(const char* path, Abi& abi)
{
    void* handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!handle)
        return false;
    return bind(handle, "charls_jpegls_encoder_create", abi.encoder_create) &&
           bind(handle, "charls_jpegls_encoder_destroy", abi.encoder_destroy) &&
           bind(handle, "charls_jpegls_encoder_set_frame_info", abi.encoder_set_frame_info) &&
           bind(handle, "charls_jpegls_encoder_set_near_lossless", abi.encoder_set_near_lossless) &&
           bind(handle, "charls_jpegls_encoder_set_interleave_mode", abi.encoder_set_interleave_mode) &&
           bind(handle, "charls_jpegls_encoder_set_color_transformation", abi.encoder_set_color_transformation) &&
           bind(handle, "charls_jpegls_encoder_set_destination_buffer", abi.encoder_set_destination_buffer) &&
           bind(handle, "charls_jpegls_encoder_encode_from_buffer", abi.encoder_encode_from_buffer) &&
           bind(handle, "charls_jpegls_encoder_get_bytes_written", abi.encoder_get_bytes_written) &&
           bind(handle, "charls_jpegls_decoder_create", abi.decoder_create) &&
           bind(handle, "charls_jpegls_decoder_destroy", abi.decoder_destroy) &&
           bind(handle, "charls_jpegls_decoder_set_source_buffer", abi.decoder_set_source_buffer) &&
           bind(handle, "charls_jpegls_decoder_read_header", abi.decoder_read_header) &&
           bind(handle, "charls_jpegls_decoder_decode_to_buffer", abi.decoder_decode_to_buffer);
}

} // namespace

extern "C" {

struct abi_driver_job
{
    const char* library_path;
    const uint8_t* frames; // n frames, frame_bytes apart
    size_t frame_bytes;
    uint8_t* streams; // n slots, stream_capacity apart
    size_t stream_capacity;
    uint8_t* decoded; // n frames, frame_bytes apart; may be null (encode only)
    size_t* sizes;    // out: n stream sizes
    int32_t n, threads;
    uint32_t width, height;
    int32_t bits_per_sample, component_count, near_lossless, interleave_mode, color_transformation;
    int32_t per_frame; // 1: a worker encodes a frame and decodes it right away; 0: all frames encoded, then all decoded
    double seconds;    // out: wall time of the whole job
};

// Returns 0 or the first charls_jpegls_errc seen (-1: the library could not be bound).
__attribute__((visibility("default"))) int32_t abi_driver_run(abi_driver_job* job)
{
    Abi abi{};
    if (!load(job->library_path, abi))
        return -1;
    std::atomic<int32_t> first_error{0};
    const auto fail = [&](int32_t errc) {
        int32_t expected = 0;
        if (errc != 0)
            first_error.compare_exchange_strong(expected, errc);
        return errc != 0;
    };
    const auto encode = [&](int32_t i) {
        void* e = abi.encoder_create();
        const FrameInfo info{job->width, job->height, job->bits_per_sample, job->component_count};
        size_t written = 0;
        if (!fail(abi.encoder_set_frame_info(e, &info)) && !fail(abi.encoder_set_near_lossless(e, job->near_lossless)) &&
            !fail(abi.encoder_set_interleave_mode(e, job->interleave_mode)) &&
            !fail(abi.encoder_set_color_transformation(e, job->color_transformation)) &&
            !fail(abi.encoder_set_destination_buffer(e, job->streams + static_cast<size_t>(i) * job->stream_capacity,
                                                     job->stream_capacity)) &&
            !fail(abi.encoder_encode_from_buffer(e, job->frames + static_cast<size_t>(i) * job->frame_bytes, job->frame_bytes, 0)))
            fail(abi.encoder_get_bytes_written(e, &written));
        abi.encoder_destroy(e);
        job->sizes[i] = written;
    };
    const auto decode = [&](int32_t i) {
        if (!job->decoded)
            return;
        void* d = abi.decoder_create();
        if (!fail(abi.decoder_set_source_buffer(d, job->streams + static_cast<size_t>(i) * job->stream_capacity, job->sizes[i])) &&
            !fail(abi.decoder_read_header(d)))
            fail(abi.decoder_decode_to_buffer(d, job->decoded + static_cast<size_t>(i) * job->frame_bytes, job->frame_bytes, 0));
        abi.decoder_destroy(d);
    };
    const auto run_phase = [&](auto&& work) {
        std::atomic<int32_t> next{0};
        std::vector<std::thread> pool;
        const int32_t threads = job->threads < 1 ? 1 : job->threads;
        for (int32_t t = 0; t < threads; ++t)
            pool.emplace_back([&] {
                for (int32_t i = next.fetch_add(1); i < job->n; i = next.fetch_add(1))
                    work(i);
            });
        for (auto& thread : pool)
            thread.join();
    };

    const auto t0 = std::chrono::steady_clock::now();
    if (job->per_frame)
    {
        run_phase([&](int32_t i) {
            encode(i);
            decode(i);
        });
    }
    else
    {
        run_phase(encode);
        run_phase(decode);
    }
    job->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return first_error.load();
}

} // extern "C"
