// engine.hpp -- host-side owner of the device resources behind one encoder / decoder / batch object.
//
// The engine is the seam the reference has between its C ABI objects and make_scan_codec()
// (reference src/charls_jpegls_encoder.cpp:285-296, src/charls_jpegls_decoder.cpp:186-189): the ABI layer hands it one
// scan at a time (host buffers), or a whole batch of device-resident frames, and it runs the kernels.
// One engine = one CUDA stream + grow-only device buffers; engines of different objects are independent, so distinct
// encoder / decoder instances may be used from different host threads concurrently (like the reference's instances).
#pragma once

#include "jls_common.h"
#include "jls_params.hpp"

#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

struct CUstream_st;
struct CUevent_st;
struct CUgraphExec_st;

namespace jls {

// error code for CUDA failures (no counterpart in the reference's charls_jpegls_errc)
constexpr int32_t errc_device_failure = 200;
constexpr int32_t errc_not_enough_memory = 1;

int32_t set_device(int32_t ordinal) noexcept; // process-wide device used by every engine
int32_t device_count(int32_t* count) noexcept;

// Where the entries of a side table of interval offsets (jls_common.h) are, segment by segment; total = 0: no table.
struct HostOffsetTable // single-image encode: absolute host addresses inside the destination's header
{
    uint32_t total{};
    uint8_t* entries[offset_table_max_segments]{};
};
struct StreamOffsetTable // offsets from the first byte of a stream (the same for every frame of a batch encode)
{
    uint32_t total{};
    size_t entry_offsets[offset_table_max_segments]{};
};

struct BatchFrame
{
    uint8_t* pixels;        // device
    uint8_t* stream;        // device
    size_t stream_capacity; // encode: capacity; decode: stream size
    size_t stream_size;     // encode out
    int32_t status;         // out
    // decode only, filled by the caller after parsing the header on the host:
    size_t scan_offset; // offset of the first entropy-coded byte
    StreamOffsetTable table{}; // the frame's side table of interval offsets, if its header has one
};

class Engine final
{
public:
    Engine() = default;
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    // Encodes one scan whose samples are in host memory at `source` (line pitch `stride`) and writes the entropy-coded
    // data (interval data + RSTm markers) to host memory at `destination`.  Returns a charls_jpegls_errc.
    int32_t encode_scan_from_host(const CodecParams& p, const uint8_t* source, size_t stride, uint8_t* destination,
                                  size_t capacity, size_t& written, const HostOffsetTable* table = nullptr);

    // The same call in two halves (charlsx_*_begin / _end): `begin` issues the input copy, the kernels and the outcome copy
    // and returns without waiting, `end` waits and moves the entropy-coded bytes.  One pending half per engine; a host
    // thread keeps several codec objects (= engines = CUDA streams) in flight this way.
    int32_t encode_scan_from_host_begin(const CodecParams& p, const uint8_t* source, size_t stride, uint8_t* destination,
                                        size_t capacity, const HostOffsetTable* table = nullptr);
    int32_t encode_scan_from_host_end(size_t& written);
    int32_t decode_scan_to_host_begin(const CodecParams& p, size_t offset, uint8_t* destination, size_t stride,
                                      const StreamOffsetTable* table = nullptr);
    int32_t decode_scan_to_host_end(size_t& consumed);

    // Copies a complete JPEG-LS stream to the device once; decode_scan_to_host then works on offsets into it.
    int32_t upload_stream(const uint8_t* host_stream, size_t size);
    // Decodes the scan whose entropy-coded data starts `offset` bytes into the uploaded stream.
    // `consumed` = bytes of entropy-coded data (up to the marker that ends the scan).
    int32_t decode_scan_to_host(const CodecParams& p, size_t offset, uint8_t* destination, size_t stride, size_t& consumed,
                                const StreamOffsetTable* table = nullptr);

    // Device-resident batches (single-scan frames).  `header` = the bytes in front of the entropy-coded data.
    // `while_coding` (optional) runs on the host after all work of the call has been issued and before the call waits
    // for it: the host-resident path issues the next chunk's copies there, behind this chunk's kernels.
    // `table` (encode): where the header reserves the side table of interval offsets, relative to every frame's stream.
    // Decode: frames[i].table; the tables are used when every frame of the batch has one.
    int32_t encode_batch(const CodecParams& p, const uint8_t* header, size_t header_size, BatchFrame* frames, size_t count,
                         size_t stride, CUstream_st* user_stream, const std::function<int32_t()>& while_coding = {},
                         const StreamOffsetTable* table = nullptr);
    int32_t decode_batch(const CodecParams& p, BatchFrame* frames, size_t count, size_t stride, CUstream_st* user_stream,
                         const std::function<int32_t()>& while_coding = {});
    // The same for frames and streams in HOST memory (pinned for full speed): the engine stages a chunk of frames at a
    // time in device memory and overlaps the copies of one chunk with the kernels of its neighbours (three CUDA
    // streams, double-buffered staging).  BatchFrame pointers are host pointers here.
    int32_t encode_batch_host(const CodecParams& p, const uint8_t* header, size_t header_size, BatchFrame* frames, size_t count,
                              size_t stride, const StreamOffsetTable* table = nullptr);
    int32_t decode_batch_host(const CodecParams& p, BatchFrame* frames, size_t count, size_t stride);
    // Downloads the first `prefix_bytes` of every frame's stream (header parsing happens on the host).
    int32_t download_prefixes(const BatchFrame* frames, size_t count, uint32_t prefix_bytes, const uint8_t*& prefixes,
                              CUstream_st* user_stream);

    uint32_t last_kernel_launches() const noexcept { return last_launches_; }
    // Device time (CUDA events on the launching stream) of the entropy-coding kernel of the last call, in milliseconds.
    float last_coder_kernel_ms() const noexcept { return last_coder_ms_; }

    // Engines are pooled: an encoder / decoder object borrows one for its lifetime, so that programs that create a
    // codec object per image (the usual way to use the reference API) do not pay for device allocations every time.
    static Engine* acquire();
    static void release(Engine* engine) noexcept;
    static void drop_pooled_except(int device) noexcept;

private:
    struct Buffer
    {
        void* data{};
        size_t capacity{};
        bool pinned{};
    };

    int32_t prepare();
    int32_t ensure(Buffer& buffer, size_t bytes, bool pinned = false);
    void release(Buffer& buffer) noexcept;
    // Lays out the per-job scratch for `job_count` jobs and uploads the job table. Scratch pointers are filled in here.
    // `upload` false: the caller copies host_jobs_ to job_table_ itself (inside a replayed graph).
    int32_t stage_jobs(const CodecParams& p, std::vector<ScanJob>& jobs, bool encode, size_t slot_bytes, CUstream_st* stream,
                       bool upload = true);
    int32_t fetch_outcomes(size_t job_count, CUstream_st* stream);

    // The single-image calls issue the same dozen launches and copies between engine buffers for every image of a
    // series.  They are captured into a CUDA graph once and replayed: one driver call instead of twelve, which matters
    // when many host threads code images at the same time (the driver serialises their calls).
    struct GraphKey
    {
        CodecParams p;
        uint64_t slot_bytes;
        uint64_t marker_blocks;
        uint32_t encode;
        uint32_t reserved;
    };
    struct CachedGraph
    {
        GraphKey key;
        uint64_t generation; // of the engine's buffers, see ensure()
        CUgraphExec_st* exec;
        uint32_t launches; // kernels inside
    };
    template<typename Enqueue>
    int32_t replay(const GraphKey& key, Enqueue&& enqueue);
    void drop_graphs() noexcept;
    void read_coder_time() noexcept;
    int32_t wait_for(CUstream_st* stream);

    struct Pending // the half-finished single-image call (begin issued, end not yet called)
    {
        bool active{};
        bool direct{};
        uint8_t* destination{};
        size_t capacity{};
        uint64_t launches_before{};
        // encode: where the side table goes on the host; decode: what a second attempt without the table needs
        HostOffsetTable host_table{};
        bool used_table{};
        CodecParams p{};
        size_t offset{};
        size_t stride{};
    };
    Pending pending_;
    CUstream_st* stream_{};
    int device_{-1};
    Buffer pixels_, stream_buffer_, slots_, interval_bytes_, interval_offset_, line_scratch_, job_table_, outcomes_, table_buffer_,
        marker_counts_, marker_totals_, marker_codes_, header_, pointer_table_, prefixes_, repitched_, row_pointers_;
    Buffer host_outcomes_, host_jobs_, host_prefixes_, host_pointer_table_, host_row_pointers_; // pinned
    // device frames whose rows are not 4-byte aligned: copied to / from an aligned pitch around the tile kernels (engine.cu)
    int32_t repitch(const BatchFrame* frames, size_t count, size_t stride, size_t pitch, size_t span, uint32_t row_bytes,
                    uint32_t height, bool to_scratch, CUstream_st* stream);
    size_t uploaded_stream_size_{};
    uint32_t last_launches_{};
    float last_coder_ms_{};
    CUevent_st* events_[2]{};
    CUevent_st* sleep_event_{};
    std::vector<CachedGraph> graphs_;
    // host-resident batches: staging slots and the two copy streams
    int32_t prepare_staging();
    static size_t staging_chunk(size_t count, size_t bytes_per_frame) noexcept;
    static constexpr int staging_slots = 4;
    Buffer stage_pixels_[staging_slots], stage_streams_[staging_slots];
    CUstream_st* copy_in_{};
    CUstream_st* copy_out_{};
    CUevent_st* in_done_[staging_slots]{};
    CUevent_st* out_done_[staging_slots]{};
    Engine* helper_{}; // second compute engine of the host-batch pipeline (odd chunks), created on first use
    uint64_t batch_launches_before_{};
    // the two halves of encode_batch / decode_batch: issue everything, then wait and collect sizes / statuses
    int32_t encode_batch_begin(const CodecParams& p, const uint8_t* header, size_t header_size, const BatchFrame* frames, size_t count,
                               size_t stride, CUstream_st* user_stream, const StreamOffsetTable* table = nullptr);
    int32_t encode_batch_end(BatchFrame* frames, size_t count, size_t header_size, CUstream_st* user_stream);
    int32_t decode_batch_begin(const CodecParams& p, const BatchFrame* frames, size_t count, size_t stride, CUstream_st* user_stream,
                               bool use_tables);
    static bool all_frames_have_tables(const CodecParams& p, const BatchFrame* frames, size_t count) noexcept;
    int32_t decode_batch_end(BatchFrame* frames, size_t count, CUstream_st* user_stream);
    // CHARLS_B200_TRACE timeline of the single-image calls (engine.cu: Trace)
    void trace_gpu(int index) noexcept;
    uint8_t* device_view_of(void* pointer) noexcept;
    void trace_host(int index) noexcept;
    void trace_commit(int kind) noexcept;
    CUevent_st* trace_events_[4]{};
    double trace_host_ms_[4]{};
    uint64_t buffer_generation_{};
};

} // namespace jls
