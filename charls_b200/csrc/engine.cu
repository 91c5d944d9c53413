// engine.cu -- device resource management and kernel orchestration (see engine.hpp).
#include "engine.hpp"

#include "jls_interval.cuh"
#include "jls_kernels.hpp"

#include <atomic>
#include <chrono>
#include <functional>
#include <string>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <mutex>
#if defined(__linux__)
#include <sched.h>
#endif
#include <vector>

namespace jls {

namespace {

std::atomic<int> g_device{-1}; // -1: whatever device is current when the first engine is used

constexpr size_t outcome_words = 4; // per job: status key, result[0], result[1], pad

size_t align_up(size_t value, size_t alignment) noexcept
{
    return (value + alignment - 1) / alignment * alignment;
}

int32_t map_cuda_error(cudaError_t error, const char* what) noexcept
{
    if (error == cudaSuccess)
        return 0;
    std::fprintf(stderr, "charls_b200: CUDA failure in %s: %s\n", what, cudaGetErrorString(error));
    cudaGetLastError(); // clear the sticky-free error state
    return error == cudaErrorMemoryAllocation ? errc_not_enough_memory : errc_device_failure;
}

#define JLS_CUDA(expr)                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
        const int32_t jls_cuda_status = map_cuda_error((expr), #expr);                                                 \
        if (jls_cuda_status != 0)                                                                                      \
            return jls_cuda_status;                                                                                    \
    } while (0)

#define JLS_CHECK(expr)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        const int32_t jls_check_status = (expr);                                                                       \
        if (jls_check_status != 0)                                                                                     \
            return jls_check_status;                                                                                   \
    } while (0)

size_t row_bytes_of(const CodecParams& p) noexcept
{
    const size_t samples_per_pixel = p.interleave == ilv_none ? 1U : static_cast<size_t>(p.components);
    return static_cast<size_t>(p.width) * samples_per_pixel * static_cast<size_t>(p.sample_bytes);
}

} // namespace

int32_t set_device(int32_t ordinal) noexcept
{
    int count = 0;
    JLS_CUDA(cudaGetDeviceCount(&count));
    if (ordinal < 0 || ordinal >= count)
        return 101; // invalid_argument
    g_device.store(ordinal);
    Engine::drop_pooled_except(ordinal);
    return 0;
}

int32_t device_count(int32_t* count) noexcept
{
    int n = 0;
    JLS_CUDA(cudaGetDeviceCount(&n));
    *count = n;
    return 0;
}

Engine::~Engine()
{
    if (device_ >= 0)
        cudaSetDevice(device_);
    for (Buffer* b : {&pixels_, &stream_buffer_, &slots_, &interval_bytes_, &interval_offset_, &line_scratch_, &job_table_,
                      &outcomes_, &marker_counts_, &marker_totals_, &marker_codes_, &header_, &pointer_table_, &prefixes_,
                      &repitched_, &row_pointers_, &host_outcomes_, &host_jobs_, &host_prefixes_, &host_pointer_table_,
                      &host_row_pointers_})
        release(*b);
    for (auto& event : events_)
        if (event)
            cudaEventDestroy(event);
    drop_graphs();
    for (Buffer& b : stage_pixels_)
        release(b);
    for (Buffer& b : stage_streams_)
        release(b);
    delete helper_;
    for (auto& event : in_done_)
        if (event)
            cudaEventDestroy(event);
    for (auto& event : out_done_)
        if (event)
            cudaEventDestroy(event);
    if (copy_in_)
        cudaStreamDestroy(copy_in_);
    if (copy_out_)
        cudaStreamDestroy(copy_out_);
    for (auto& event : trace_events_)
        if (event)
            cudaEventDestroy(event);
    if (sleep_event_)
        cudaEventDestroy(sleep_event_);
    if (stream_)
        cudaStreamDestroy(stream_);
}

namespace {
std::mutex g_pool_mutex;
std::vector<Engine*> g_pool;
constexpr size_t pool_limit = 64;
// CHARLS_B200_TRACE=<file>: a timeline of the single-image calls (host clock and CUDA events against one base event),
// written when the library is unloaded.  tools/e2e_timeline.py reads it.  Costs nothing when the variable is not set.
struct TraceRecord
{
    uint64_t thread;
    int32_t kind; // 0 encode, 1 decode
    double host_ms[4]; // call entered, work enqueued, outcome known, call done
    float gpu_ms[4];   // stream reached the call, input copy done, kernels done, output copy done
};

struct Trace
{
    bool enabled = false;
    std::string path;
    std::mutex mutex;
    std::vector<TraceRecord> records;
    cudaEvent_t base = nullptr;
    std::chrono::steady_clock::time_point start;

    Trace()
    {
        const char* value = std::getenv("CHARLS_B200_TRACE");
        if (value && value[0])
        {
            enabled = true;
            path = value;
            start = std::chrono::steady_clock::now();
        }
    }
    ~Trace()
    {
        if (!enabled)
            return;
        if (FILE* f = std::fopen(path.c_str(), "w"))
        {
            for (const TraceRecord& r : records)
                std::fprintf(f, "%llu %d %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f\n", static_cast<unsigned long long>(r.thread),
                             r.kind, r.host_ms[0], r.host_ms[1], r.host_ms[2], r.host_ms[3], r.gpu_ms[0], r.gpu_ms[1], r.gpu_ms[2],
                             r.gpu_ms[3]);
            std::fclose(f);
        }
    }
    double now() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count(); }
};
Trace g_trace;

// Streams are multiplexed onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams that share a queue wait
// for each other's kernels.  Sixteen host threads with a codec object (= a stream) each reached 2300 single-frame round
// trips/s on device-resident data with the default and 3570 with 32 queues (profiles/r1_notes.md).  The variable is read
// when the CUDA context is created, so this only helps when the library is loaded before that; a value the user set wins.
const int g_connections_set = setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);

std::atomic<int> g_borrowed{0};           // engines that codec objects hold right now
// Codec objects in flight from which on a waiting caller sleeps instead of spinning (Engine::wait_for): at least 6, and not
// before there are more objects in flight than this process has cores to spin on -- the cores it may run on (affinity, capped by
// a cgroup CPU quota) divided by the GPUs of the box, since a multi-GPU job runs one such process per GPU.  Measured with 16
// callers on a 16-core single-GPU box: 26.8 GPix/s when they spin, 25.4 when they sleep (profiles/r2k_copy_pattern_probe.txt);
// with 8 processes x 8 callers on the 32 vCPUs of an 8-GPU box they have to sleep.
int blocking_sync_threshold()
{
    static const int value = [] {
        unsigned cpus = std::thread::hardware_concurrency();
#if defined(__linux__)
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0)
            cpus = static_cast<unsigned>(CPU_COUNT(&set));
        if (FILE* f = std::fopen("/sys/fs/cgroup/cpu.max", "r"))
        {
            char quota[32] = {};
            unsigned long period = 0;
            if (std::fscanf(f, "%31s %lu", quota, &period) == 2 && period > 0 && std::strcmp(quota, "max") != 0)
            {
                const unsigned long allowed = (std::strtoul(quota, nullptr, 10) + period - 1) / period;
                if (allowed > 0 && allowed < cpus)
                    cpus = static_cast<unsigned>(allowed);
            }
            std::fclose(f);
        }
#endif
        int devices = 1;
        if (cudaGetDeviceCount(&devices) != cudaSuccess || devices < 1)
        {
            cudaGetLastError();
            devices = 1;
        }
        const int per_device = static_cast<int>(cpus) / devices;
        return per_device + 1 > 6 ? per_device + 1 : 6;
    }();
    return value;
}
} // namespace

// The device a new piece of work belongs to: the one set with charlsx_set_device, else the calling thread's current device
// (-1 when that cannot be asked, e.g. without a driver: the engine then fails in prepare()).
static int wanted_device() noexcept
{
    int wanted = g_device.load();
    if (wanted < 0 && cudaGetDevice(&wanted) != cudaSuccess)
    {
        cudaGetLastError();
        wanted = -1;
    }
    return wanted;
}

// An engine stays on the device it was first used on (its stream and buffers live there), so the pool hands out only
// engines of the device the caller is on, or ones that have not been used yet.
Engine* Engine::acquire()
{
    g_borrowed.fetch_add(1, std::memory_order_relaxed);
    const int wanted = wanted_device();
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        for (size_t i = g_pool.size(); i-- > 0;)
        {
            Engine* engine = g_pool[i];
            if (engine->device_ < 0 || engine->device_ == wanted)
            {
                g_pool.erase(g_pool.begin() + static_cast<std::ptrdiff_t>(i));
                return engine;
            }
        }
    }
    return new Engine;
}

void Engine::release(Engine* engine) noexcept
{
    if (!engine)
        return;
    g_borrowed.fetch_sub(1, std::memory_order_relaxed);
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (g_pool.size() < pool_limit)
        {
            g_pool.push_back(engine);
            return;
        }
    }
    delete engine;
}

// charlsx_set_device: pooled engines of other devices would never be handed out again
void Engine::drop_pooled_except(int device) noexcept
{
    std::vector<Engine*> stale;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        for (size_t i = g_pool.size(); i-- > 0;)
        {
            if (g_pool[i]->device_ >= 0 && g_pool[i]->device_ != device)
            {
                stale.push_back(g_pool[i]);
                g_pool.erase(g_pool.begin() + static_cast<std::ptrdiff_t>(i));
            }
        }
    }
    for (Engine* engine : stale)
        delete engine;
}

// Waits until `stream` has drained.  A lone caller spins (lowest latency).  When many codec objects are in flight -- one
// host thread per image is how the reference API gets used in parallel -- spinning threads eat the cores (and any CPU
// quota of the container) the other callers need to feed the GPU, so from `blocking_sync_threshold` borrowed engines on
// the thread sleeps on an event instead (blocking_sync_threshold() above).  CHARLS_B200_BLOCKING_SYNC=0 / 1 forces one behaviour.
int32_t Engine::wait_for(CUstream_st* stream)
{
    static const int forced = [] {
        const char* value = std::getenv("CHARLS_B200_BLOCKING_SYNC");
        return value && (value[0] == '0' || value[0] == '1') ? value[0] - '0' : -1;
    }();
    const bool blocking = forced >= 0 ? forced == 1 : g_borrowed.load(std::memory_order_relaxed) >= blocking_sync_threshold();
    if (!blocking)
    {
        JLS_CUDA(cudaStreamSynchronize(stream));
        return 0;
    }
    if (!sleep_event_)
        JLS_CUDA(cudaEventCreateWithFlags(&sleep_event_, cudaEventBlockingSync | cudaEventDisableTiming));
    JLS_CUDA(cudaEventRecord(sleep_event_, stream));
    JLS_CUDA(cudaEventSynchronize(sleep_event_));
    return 0;
}

// The address under which the device can write `pointer` directly, or null for ordinary pageable host memory.
// Opt-in (CHARLS_B200_DIRECT_OUTPUT=1): it saves the encoder's second wait, but 128-byte stores over PCIe are slower than
// the copy engine's transfers -- with 16 callers on one B200 the round-trip rate fell from 24.9 to 22.0 GPix/s
// (profiles/r1_notes.md), so the staged copy stays the default.
uint8_t* Engine::device_view_of(void* pointer) noexcept
{
    const char* value = std::getenv("CHARLS_B200_DIRECT_OUTPUT");
    const bool enabled = value && value[0] == '1';
    if (!enabled || !pointer)
        return nullptr;
    cudaPointerAttributes attributes{};
    if (cudaPointerGetAttributes(&attributes, pointer) != cudaSuccess)
    {
        cudaGetLastError();
        return nullptr;
    }
    if (attributes.type == cudaMemoryTypeUnregistered || !attributes.devicePointer)
        return nullptr;
    if (attributes.type == cudaMemoryTypeDevice && attributes.device != device_)
        return nullptr; // another GPU's memory: take the staged copy
    return static_cast<uint8_t*>(attributes.devicePointer);
}

void Engine::trace_gpu(int index) noexcept
{
    if (!g_trace.enabled)
        return;
    if (!trace_events_[index])
        cudaEventCreate(&trace_events_[index]);
    cudaEventRecord(trace_events_[index], stream_);
}

void Engine::trace_host(int index) noexcept
{
    if (g_trace.enabled)
        trace_host_ms_[index] = g_trace.now();
}

void Engine::trace_commit(int kind) noexcept
{
    if (!g_trace.enabled)
        return;
    std::lock_guard<std::mutex> lock(g_trace.mutex);
    if (!g_trace.base)
    {
        // the base event goes first on the timeline: it is recorded before this call's events are read, on the
        // legacy-free stream of this engine, and everything already recorded is measured against it (negative is fine)
        cudaEventCreate(&g_trace.base);
        cudaEventRecord(g_trace.base, stream_);
        cudaEventSynchronize(g_trace.base);
    }
    TraceRecord r{};
    r.thread = std::hash<std::thread::id>{}(std::this_thread::get_id()) % 100000;
    r.kind = kind;
    for (int i = 0; i < 4; ++i)
    {
        r.host_ms[i] = trace_host_ms_[i];
        if (!trace_events_[i] || cudaEventElapsedTime(&r.gpu_ms[i], g_trace.base, trace_events_[i]) != cudaSuccess)
        {
            cudaGetLastError();
            r.gpu_ms[i] = -1.0F;
        }
    }
    g_trace.records.push_back(r);
}

void Engine::read_coder_time() noexcept
{
    float ms = 0.0F;
    if (cudaEventElapsedTime(&ms, events_[0], events_[1]) == cudaSuccess)
        last_coder_ms_ = ms;
    else
        cudaGetLastError();
}

void Engine::release(Buffer& buffer) noexcept
{
    if (buffer.data)
    {
        if (buffer.pinned)
            cudaFreeHost(buffer.data);
        else
            cudaFree(buffer.data);
    }
    buffer = Buffer{};
}

int32_t Engine::prepare()
{
    int wanted = g_device.load();
    if (wanted < 0)
    {
        JLS_CUDA(cudaGetDevice(&wanted));
    }
    if (device_ >= 0 && device_ != wanted)
        return 100; // invalid_operation: an object stays on the device it was first used on
    JLS_CUDA(cudaSetDevice(wanted));
    device_ = wanted;
    if (!stream_)
        JLS_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    for (auto& event : events_)
        if (!event)
            JLS_CUDA(cudaEventCreate(&event));
    return 0;
}

int32_t Engine::ensure(Buffer& buffer, size_t bytes, bool pinned)
{
    if (buffer.capacity >= bytes && buffer.data)
        return 0;
    release(buffer);
    ++buffer_generation_; // captured graphs hold the old address
    const size_t capacity = align_up(bytes + bytes / 8 + 256, 256);
    void* data = nullptr;
    // Fresh memory is zeroed once (buffers only grow, so this is rare): kernels read whole aligned words around the ends of
    // streams and the outcome copy carries a pad word, all masked or ignored -- but initcheck (tools/sanitize_gpu.sh) then has
    // nothing to report, and nothing a caller gets back can depend on what the allocator handed out.
    if (pinned)
    {
        JLS_CUDA(cudaMallocHost(&data, capacity));
        std::memset(data, 0, capacity);
    }
    else
    {
        JLS_CUDA(cudaMalloc(&data, capacity));
        buffer.data = data; // owned from here on: an early return below leaves it to release()
        buffer.capacity = capacity;
        buffer.pinned = false;
        JLS_CUDA(cudaMemset(data, 0, capacity));
        JLS_CUDA(cudaStreamSynchronize(cudaStreamLegacy)); // kernels run on non-blocking streams: the memset must be over
    }
    buffer.data = data;
    buffer.capacity = capacity;
    buffer.pinned = pinned;
    return 0;
}

int32_t Engine::stage_jobs(const CodecParams& p, std::vector<ScanJob>& jobs, bool encode, size_t slot_bytes, CUstream_st* stream,
                           bool upload)
{
    const size_t n = jobs.size();
    const size_t intervals = p.interval_count;
    const bool general = !use_fast_path(p);
    const size_t slots_per_job = encode ? intervals * slot_bytes : 0;
    const size_t offsets_per_job = 2 * intervals + 2;
    const size_t lines_per_job = general ? static_cast<size_t>(2) * p.components * (p.width + 2) * intervals : 0;

    if (encode)
    {
        JLS_CHECK(ensure(slots_, n * slots_per_job + 64));
        JLS_CHECK(ensure(interval_bytes_, n * intervals * sizeof(uint32_t)));
    }
    JLS_CHECK(ensure(interval_offset_, n * offsets_per_job * sizeof(uint64_t)));
    if (general)
        JLS_CHECK(ensure(line_scratch_, n * lines_per_job * sizeof(uint16_t) + 64));
    JLS_CHECK(ensure(outcomes_, n * outcome_words * sizeof(uint64_t)));
    JLS_CHECK(ensure(job_table_, n * sizeof(ScanJob)));
    JLS_CHECK(ensure(host_jobs_, n * sizeof(ScanJob), true));

    for (size_t i = 0; i < n; ++i)
    {
        ScanJob& job = jobs[i];
        job.slots = encode ? static_cast<uint8_t*>(slots_.data) + i * slots_per_job : nullptr;
        job.interval_bytes = encode ? static_cast<uint32_t*>(interval_bytes_.data) + i * intervals : nullptr;
        job.interval_offset = static_cast<uint64_t*>(interval_offset_.data) + i * offsets_per_job;
        job.line_scratch = general ? static_cast<uint16_t*>(line_scratch_.data) + i * lines_per_job : nullptr;
        job.status = static_cast<uint64_t*>(outcomes_.data) + i * outcome_words;
        job.result = job.status + 1;
    }
    std::memcpy(host_jobs_.data, jobs.data(), n * sizeof(ScanJob));
    if (upload)
        JLS_CUDA(cudaMemcpyAsync(job_table_.data, host_jobs_.data, n * sizeof(ScanJob), cudaMemcpyHostToDevice, stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Graph replay for the single-image calls
// ---------------------------------------------------------------------------------------------------------------------
void Engine::drop_graphs() noexcept
{
    for (CachedGraph& g : graphs_)
        cudaGraphExecDestroy(g.exec);
    graphs_.clear();
}

// `enqueue` issues work on stream_ only, reads and writes engine buffers only (their addresses are part of the captured
// graph; `key` must cover every by-value kernel argument) and makes no synchronising call.  CHARLS_B200_GRAPHS=0 turns
// the replay off (every call enqueues directly).
template<typename Enqueue>
int32_t Engine::replay(const GraphKey& key, Enqueue&& enqueue)
{
    static const bool enabled = [] {
        const char* value = std::getenv("CHARLS_B200_GRAPHS");
        return !(value && value[0] == '0');
    }();
    if (!enabled)
        return enqueue();

    if (!graphs_.empty() && graphs_.front().generation != buffer_generation_)
        drop_graphs(); // a buffer moved
    for (const CachedGraph& g : graphs_)
    {
        if (std::memcmp(&g.key, &key, sizeof(GraphKey)) == 0)
        {
            JLS_CUDA(cudaGraphLaunch(g.exec, stream_));
            count_kernel_launches(g.launches);
            return 0;
        }
    }

    const uint64_t launches_before = thread_kernel_launch_count();
    JLS_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
    const int32_t status = enqueue();
    cudaGraph_t graph = nullptr;
    const cudaError_t end = cudaStreamEndCapture(stream_, &graph);
    const uint32_t launches = static_cast<uint32_t>(thread_kernel_launch_count() - launches_before);
    if (status != 0 || end != cudaSuccess || !graph)
    {
        if (graph)
            cudaGraphDestroy(graph);
        cudaGetLastError();
        return status != 0 ? status : 200;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t made = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (made != cudaSuccess)
    {
        cudaGetLastError();
        return 200;
    }
    constexpr size_t graph_cache_limit = 8;
    if (graphs_.size() >= graph_cache_limit)
    {
        cudaGraphExecDestroy(graphs_.front().exec);
        graphs_.erase(graphs_.begin());
    }
    graphs_.push_back(CachedGraph{key, buffer_generation_, exec, launches});
    JLS_CUDA(cudaGraphLaunch(exec, stream_));
    return 0;
}

int32_t Engine::fetch_outcomes(size_t job_count, CUstream_st* stream)
{
    JLS_CHECK(ensure(host_outcomes_, job_count * outcome_words * sizeof(uint64_t), true));
    JLS_CUDA(cudaMemcpyAsync(host_outcomes_.data, outcomes_.data, job_count * outcome_words * sizeof(uint64_t),
                             cudaMemcpyDeviceToHost, stream));
    JLS_CHECK(wait_for(stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// C ABI path: host buffers
// ---------------------------------------------------------------------------------------------------------------------
int32_t Engine::encode_scan_from_host(const CodecParams& p, const uint8_t* source, size_t stride, uint8_t* destination,
                                      size_t capacity, size_t& written, const HostOffsetTable* table)
{
    written = 0;
    JLS_CHECK(encode_scan_from_host_begin(p, source, stride, destination, capacity, table));
    return encode_scan_from_host_end(written);
}

// Everything of an encode that needs no answer from the device: input copy, kernels, the 32-byte outcome.  Nothing waits.
int32_t Engine::encode_scan_from_host_begin(const CodecParams& p, const uint8_t* source, size_t stride, uint8_t* destination,
                                            size_t capacity, const HostOffsetTable* table)
{
    pending_ = Pending{};
    JLS_CHECK(prepare());
    trace_host(0);
    trace_gpu(0);
    pending_.launches_before = thread_kernel_launch_count();
    const size_t row_bytes = row_bytes_of(p);
    const size_t pitch = align_up(row_bytes, 16);
    JLS_CHECK(ensure(pixels_, pitch * static_cast<size_t>(p.height) + 64));
    if (stride == pitch)
        JLS_CUDA(cudaMemcpyAsync(pixels_.data, source, pitch * (static_cast<size_t>(p.height) - 1) + row_bytes,
                                 cudaMemcpyHostToDevice, stream_));
    else
        JLS_CUDA(cudaMemcpy2DAsync(pixels_.data, pitch, source, stride, row_bytes, static_cast<size_t>(p.height),
                                   cudaMemcpyHostToDevice, stream_));
    trace_gpu(1);

    const size_t slot_bytes = worst_case_interval_bytes(p, p.lines_per_interval);
    if (slot_bytes >= (size_t{1} << 32))
        return 7; // parameter_value_not_supported: one restart interval must stay below 4 GiB of entropy-coded data
    const size_t worst_total = static_cast<size_t>(p.interval_count) * (slot_bytes + 2);
    const size_t device_capacity = capacity < worst_total ? capacity : worst_total;
    // Opt-in: a destination the device can address (page-locked or registered host memory) is written by the gather
    // kernel itself -- no staging buffer, no second copy whose size the host has to wait for first (see device_view_of).
    uint8_t* direct = device_view_of(destination);
    if (direct && device_capacity > 1 && !device_view_of(destination + device_capacity - 1))
        direct = nullptr; // only the beginning of the buffer is registered
    if (!direct)
        JLS_CHECK(ensure(stream_buffer_, device_capacity + 64));

    std::vector<ScanJob> jobs(1);
    jobs[0] = ScanJob{};
    jobs[0].pixels_in = static_cast<const uint8_t*>(pixels_.data);
    jobs[0].stride = pitch;
    jobs[0].stream_out = direct ? direct : static_cast<uint8_t*>(stream_buffer_.data);
    jobs[0].stream_out_capacity = device_capacity;
    // side table of interval offsets: the entries are made on the device (big endian, segment after segment in
    // table_buffer_) and copied into the places the header reserves for them once the scan is done
    const bool with_table = table != nullptr && table->total == p.interval_count + 1U;
    if (with_table)
    {
        JLS_CHECK(ensure(table_buffer_, static_cast<size_t>(table->total) * 4U + 64));
        jobs[0].offset_table.total = table->total;
        for (uint32_t segment = 0; segment < offset_table_segment_count(table->total); ++segment)
            jobs[0].offset_table.entries[segment] =
                static_cast<uint8_t*>(table_buffer_.data) + static_cast<size_t>(segment) * offset_table_entries_per_segment * 4U;
    }
    JLS_CHECK(stage_jobs(p, jobs, true, slot_bytes, stream_, false));
    JLS_CHECK(ensure(host_outcomes_, outcome_words * sizeof(uint64_t), true));
    GraphKey key{};
    key.p = p;
    key.slot_bytes = slot_bytes;
    key.encode = 1;
    key.reserved = with_table ? 1U : 0U;
    JLS_CHECK(replay(key, [&]() -> int32_t {
        JLS_CUDA(cudaMemcpyAsync(job_table_.data, host_jobs_.data, sizeof(ScanJob), cudaMemcpyHostToDevice, stream_));
        JLS_CUDA(launch_encode(p, static_cast<const ScanJob*>(job_table_.data), 1, slot_bytes, stream_, nullptr, true, with_table));
        JLS_CUDA(cudaMemcpyAsync(host_outcomes_.data, outcomes_.data, outcome_words * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                 stream_));
        return 0;
    }));
    trace_gpu(2);
    trace_host(1);
    pending_.active = true;
    pending_.destination = destination;
    pending_.capacity = capacity;
    pending_.direct = direct != nullptr;
    if (with_table)
        pending_.host_table = *table;
    return 0;
}

// Waits for the outcome, then fetches the entropy-coded bytes (their count is only known now) and waits for them.
int32_t Engine::encode_scan_from_host_end(size_t& written)
{
    written = 0;
    if (!pending_.active)
        return 100; // invalid_operation
    pending_.active = false;
    JLS_CHECK(wait_for(stream_));
    trace_host(2);
    last_coder_ms_ = 0.0F; // not measured on this path (the batch interface does)
    last_launches_ = static_cast<uint32_t>(thread_kernel_launch_count() - pending_.launches_before);

    const uint64_t* outcome = static_cast<const uint64_t*>(host_outcomes_.data);
    if (outcome[0] != ~0ULL)
        return static_cast<int32_t>(outcome[0] & 0xFF);
    const uint64_t total = outcome[1];
    if (total > pending_.capacity)
        return err_destination_too_small;
    const uint32_t table_entries = pending_.host_table.total;
    for (uint32_t segment = 0; segment < offset_table_segment_count(table_entries) && table_entries != 0; ++segment)
    {
        const uint32_t first = segment * offset_table_entries_per_segment;
        const uint32_t count = table_entries - first < offset_table_entries_per_segment ? table_entries - first : offset_table_entries_per_segment;
        JLS_CUDA(cudaMemcpyAsync(pending_.host_table.entries[segment], static_cast<const uint8_t*>(table_buffer_.data) + static_cast<size_t>(first) * 4U,
                                 static_cast<size_t>(count) * 4U, cudaMemcpyDeviceToHost, stream_));
    }
    if ((total != 0 && !pending_.direct) || table_entries != 0)
    {
        if (total != 0 && !pending_.direct)
            JLS_CUDA(cudaMemcpyAsync(pending_.destination, stream_buffer_.data, total, cudaMemcpyDeviceToHost, stream_));
        trace_gpu(3);
        JLS_CHECK(wait_for(stream_));
    }
    trace_host(3);
    trace_commit(0);
    written = static_cast<size_t>(total);
    return 0;
}

int32_t Engine::upload_stream(const uint8_t* host_stream, size_t size)
{
    JLS_CHECK(prepare());
    trace_host(0);
    trace_gpu(0);
    JLS_CHECK(ensure(stream_buffer_, size + 64));
    JLS_CUDA(cudaMemcpyAsync(stream_buffer_.data, host_stream, size, cudaMemcpyHostToDevice, stream_));
    trace_gpu(1);
    uploaded_stream_size_ = size;
    return 0;
}

int32_t Engine::decode_scan_to_host(const CodecParams& p, size_t offset, uint8_t* destination, size_t stride, size_t& consumed,
                                    const StreamOffsetTable* table)
{
    consumed = 0;
    JLS_CHECK(decode_scan_to_host_begin(p, offset, destination, stride, table));
    return decode_scan_to_host_end(consumed);
}

// Kernels, outcome and the copy of the samples to the caller's buffer are all issued here; nothing waits.
int32_t Engine::decode_scan_to_host_begin(const CodecParams& p, size_t offset, uint8_t* destination, size_t stride,
                                          const StreamOffsetTable* table)
{
    pending_ = Pending{};
    JLS_CHECK(prepare());
    if (offset > uploaded_stream_size_)
        return err_need_more_data;
    pending_.launches_before = thread_kernel_launch_count();
    const size_t remaining = uploaded_stream_size_ - offset;
    const size_t row_bytes = row_bytes_of(p);
    const size_t pitch = align_up(row_bytes, 16);
    JLS_CHECK(ensure(pixels_, pitch * static_cast<size_t>(p.height) + 64));

    // the marker kernels take their bounds from the job; sizing their grid for the next MiB lets streams of similar
    // size share one captured graph
    const size_t grid_bytes = align_up(remaining + 1, size_t{1} << 20);
    const size_t blocks = marker_blocks_for(grid_bytes);
    JLS_CHECK(ensure(marker_counts_, marker_scratch_bytes(1, grid_bytes)));
    JLS_CHECK(ensure(marker_totals_, sizeof(uint32_t)));
    JLS_CHECK(ensure(marker_codes_, p.interval_count));

    std::vector<ScanJob> jobs(1);
    jobs[0] = ScanJob{};
    jobs[0].pixels_out = static_cast<uint8_t*>(pixels_.data);
    jobs[0].stride = pitch;
    jobs[0].stream_in = static_cast<const uint8_t*>(stream_buffer_.data) + offset;
    jobs[0].stream_in_size = remaining;
    // a side table of interval offsets in the stream's header replaces the marker search (it is checked on the device; if
    // the stream does not agree with it, decode_scan_to_host_end decodes again without)
    const bool with_table = table != nullptr && table->total == p.interval_count + 1U;
    if (with_table)
    {
        jobs[0].offset_table.total = table->total;
        for (uint32_t segment = 0; segment < offset_table_segment_count(table->total); ++segment)
            jobs[0].offset_table.entries[segment] = static_cast<uint8_t*>(stream_buffer_.data) + table->entry_offsets[segment];
    }
    JLS_CHECK(stage_jobs(p, jobs, false, 0, stream_, false));
    JLS_CHECK(ensure(host_outcomes_, outcome_words * sizeof(uint64_t), true));
    GraphKey key{};
    key.p = p;
    key.marker_blocks = blocks;
    key.reserved = with_table ? 1U : 0U;
    JLS_CHECK(replay(key, [&]() -> int32_t {
        JLS_CUDA(cudaMemcpyAsync(job_table_.data, host_jobs_.data, sizeof(ScanJob), cudaMemcpyHostToDevice, stream_));
        JLS_CUDA(launch_decode(p, static_cast<const ScanJob*>(job_table_.data), 1, grid_bytes,
                               static_cast<uint32_t*>(marker_counts_.data), static_cast<uint32_t*>(marker_totals_.data),
                               static_cast<uint8_t*>(marker_codes_.data), stream_, nullptr, true, with_table));
        JLS_CUDA(cudaMemcpyAsync(host_outcomes_.data, outcomes_.data, outcome_words * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                 stream_));
        return 0;
    }));
    trace_gpu(2);
    // The samples follow the kernels without a host round trip in between: the caller's buffer is specified only for a
    // successful decode (the reference leaves the lines it got to before the error), so nothing is lost when the
    // outcome turns out to be an error.
    if (stride == pitch)
        JLS_CUDA(cudaMemcpyAsync(destination, pixels_.data, pitch * (static_cast<size_t>(p.height) - 1) + row_bytes,
                                 cudaMemcpyDeviceToHost, stream_));
    else
        JLS_CUDA(cudaMemcpy2DAsync(destination, stride, pixels_.data, pitch, row_bytes, static_cast<size_t>(p.height),
                                   cudaMemcpyDeviceToHost, stream_));
    trace_gpu(3);
    trace_host(1);
    pending_.active = true;
    pending_.used_table = with_table;
    pending_.p = p;
    pending_.offset = offset;
    pending_.destination = destination;
    pending_.stride = stride;
    return 0;
}

int32_t Engine::decode_scan_to_host_end(size_t& consumed)
{
    consumed = 0;
    if (!pending_.active)
        return 100; // invalid_operation
    pending_.active = false;
    JLS_CHECK(wait_for(stream_));
    if (pending_.used_table && static_cast<const uint64_t*>(host_outcomes_.data)[0] != ~0ULL)
    {
        // Not a clean decode with the side table: the table may be wrong, or the stream damaged.  Either way the answer is
        // what the marker search gives (errors included), so decode once more without the table.
        const Pending again = pending_;
        JLS_CHECK(decode_scan_to_host_begin(again.p, again.offset, again.destination, again.stride, nullptr));
        return decode_scan_to_host_end(consumed);
    }
    trace_host(2);
    last_coder_ms_ = 0.0F; // not measured on this path (the batch interface does)
    last_launches_ = static_cast<uint32_t>(thread_kernel_launch_count() - pending_.launches_before);
    trace_host(3);
    trace_commit(1);

    const uint64_t* outcome = static_cast<const uint64_t*>(host_outcomes_.data);
    if (outcome[0] != ~0ULL)
        return static_cast<int32_t>(outcome[0] & 0xFF);
    consumed = static_cast<size_t>(outcome[1]);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Batch path: device-resident frames
// ---------------------------------------------------------------------------------------------------------------------
// Device frames whose rows do not start on 4-byte boundaries (a stride that is not a multiple of four, or such a frame
// address) would take the kernels without shared-memory tiles, at half the speed or less (profiles/r2_notes.md).  One pass
// of this kernel copies them to an aligned pitch in front of the encoder, or back behind the decoder: 2 x 2.1 GB for 128 frames
// of 4095 x 4096, about a millisecond, against 4 (encode) and 11 ms (decode) lost otherwise.
namespace {

__global__ void k_repitch(uint8_t* const* __restrict__ user_frames, size_t user_stride, uint8_t* __restrict__ scratch, size_t span,
                          size_t pitch, uint32_t row_bytes, uint32_t height, uint32_t count, bool to_scratch)
{
    const uint32_t words = (row_bytes + 3U) / 4U;
    for (uint32_t f = blockIdx.z; f < count; f += gridDim.z)
    {
        uint8_t* const frame = user_frames[f];
        for (uint32_t row = blockIdx.y; row < height; row += gridDim.y)
        {
            for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < words; t += gridDim.x * blockDim.x)
            {
                uint8_t* user = frame + static_cast<size_t>(row) * user_stride + 4U * t;
                uint32_t* aligned = reinterpret_cast<uint32_t*>(scratch + static_cast<size_t>(f) * span + static_cast<size_t>(row) * pitch) + t;
                const uint32_t n = row_bytes - 4U * t < 4U ? row_bytes - 4U * t : 4U;
                if (to_scratch)
                {
                    uint32_t word = 0;
                    for (uint32_t b = 0; b < n; ++b)
                        word |= static_cast<uint32_t>(user[b]) << (8U * b);
                    *aligned = word;
                }
                else
                {
                    const uint32_t word = *aligned;
                    for (uint32_t b = 0; b < n; ++b)
                        user[b] = static_cast<uint8_t>(word >> (8U * b));
                }
            }
        }
    }
}

} // namespace

int32_t Engine::repitch(const BatchFrame* frames, size_t count, size_t stride, size_t pitch, size_t span, uint32_t row_bytes,
                        uint32_t height, bool to_scratch, CUstream_st* stream)
{
    JLS_CHECK(ensure(row_pointers_, count * sizeof(void*)));
    JLS_CHECK(ensure(host_row_pointers_, count * sizeof(void*), true));
    auto** host_pointers = static_cast<uint8_t**>(host_row_pointers_.data);
    for (size_t i = 0; i < count; ++i)
        host_pointers[i] = frames[i].pixels;
    JLS_CUDA(cudaMemcpyAsync(row_pointers_.data, host_row_pointers_.data, count * sizeof(void*), cudaMemcpyHostToDevice, stream));
    const uint32_t words = (row_bytes + 3U) / 4U;
    const dim3 block(256);
    // a few thousand blocks that walk over rows and frames: one block per 1 KB of a row spends the time on block scheduling
    // (2 M blocks for 128 frames of 4096 lines: 2.4 ms per pass)
    const dim3 grid((words + 255U) / 256U, height < 64U ? height : 64U, count < 32U ? static_cast<uint32_t>(count) : 32U);
    k_repitch<<<grid, block, 0, stream>>>(static_cast<uint8_t* const*>(row_pointers_.data), stride, static_cast<uint8_t*>(repitched_.data),
                                          span, pitch, row_bytes, height, static_cast<uint32_t>(count), to_scratch);
    JLS_CUDA(cudaGetLastError());
    count_kernel_launches(1);
    return 0;
}

int32_t Engine::encode_batch(const CodecParams& p, const uint8_t* header, size_t header_size, BatchFrame* frames, size_t count,
                             size_t stride, CUstream_st* user_stream, const std::function<int32_t()>& while_coding,
                             const StreamOffsetTable* table)
{
    if (count == 0)
        return prepare();
    JLS_CHECK(encode_batch_begin(p, header, header_size, frames, count, stride, user_stream, table));
    if (while_coding)
        JLS_CHECK(while_coding());
    return encode_batch_end(frames, count, header_size, user_stream);
}

// Issues the whole batch (job table, kernels, headers and EOI, outcome copy) on `user_stream` or the engine's own stream
// and returns; encode_batch_end waits and fills in sizes and statuses.
int32_t Engine::encode_batch_begin(const CodecParams& p, const uint8_t* header, size_t header_size, const BatchFrame* frames,
                                   size_t count, size_t stride, CUstream_st* user_stream, const StreamOffsetTable* table)
{
    JLS_CHECK(prepare());
    cudaStream_t stream = user_stream ? user_stream : stream_;
    batch_launches_before_ = thread_kernel_launch_count();
    const size_t slot_bytes = worst_case_interval_bytes(p, p.lines_per_interval);
    if (slot_bytes >= (size_t{1} << 32))
        return 7; // parameter_value_not_supported, see encode_scan_from_host

    JLS_CHECK(ensure(header_, header_size + 16));
    JLS_CHECK(ensure(host_prefixes_, header_size + 16, true));
    std::memcpy(host_prefixes_.data, header, header_size);
    JLS_CUDA(cudaMemcpyAsync(header_.data, host_prefixes_.data, header_size, cudaMemcpyHostToDevice, stream));

    const bool with_table = table != nullptr && table->total == p.interval_count + 1U;
    // the tile kernels keep row offsets of up to three strides in 32 bits (jls_tile.cuh, TileWalk)
    bool word_aligned = stride % 4 == 0 && stride < (size_t{1} << 30);
    std::vector<ScanJob> jobs(count);
    for (size_t i = 0; i < count; ++i)
    {
        ScanJob& job = jobs[i];
        job = ScanJob{};
        word_aligned = word_aligned && reinterpret_cast<uintptr_t>(frames[i].pixels) % 4 == 0;
        job.pixels_in = frames[i].pixels;
        job.stride = stride;
        const bool room = frames[i].stream_capacity >= header_size + 2;
        job.stream_out = frames[i].stream + header_size;
        job.stream_out_capacity = room ? frames[i].stream_capacity - header_size - 2 : 0;
        if (with_table && room)
        {
            job.offset_table.total = table->total;
            for (uint32_t segment = 0; segment < offset_table_segment_count(table->total); ++segment)
                job.offset_table.entries[segment] = frames[i].stream + table->entry_offsets[segment];
        }
    }
    if (!word_aligned && use_fast_path(p) && p.interleave != ilv_line)
    {
        const size_t row_bytes = row_bytes_of(p);
        const size_t pitch = align_up(row_bytes, 16), span = align_up(pitch * static_cast<size_t>(p.height), 256);
        JLS_CHECK(ensure(repitched_, count * span + 64));
        JLS_CHECK(repitch(frames, count, stride, pitch, span, static_cast<uint32_t>(row_bytes), p.height, true, stream));
        for (size_t i = 0; i < count; ++i)
        {
            jobs[i].pixels_in = static_cast<const uint8_t*>(repitched_.data) + i * span;
            jobs[i].stride = pitch;
        }
        word_aligned = pitch < (size_t{1} << 30);
    }
    JLS_CHECK(stage_jobs(p, jobs, true, slot_bytes, stream));
    const ScanJob* device_jobs = static_cast<const ScanJob*>(job_table_.data);
    // the header (with the table's entries still zero) goes in first, the entries are written over it
    if (with_table)
        JLS_CUDA(launch_wrap_frames(device_jobs, static_cast<const uint8_t*>(header_.data), static_cast<uint32_t>(header_size),
                                    static_cast<uint32_t>(count), stream, wrap_header));
    JLS_CUDA(launch_encode(p, device_jobs, static_cast<uint32_t>(count), slot_bytes, stream, events_, word_aligned, with_table));
    JLS_CUDA(launch_wrap_frames(device_jobs, static_cast<const uint8_t*>(header_.data), static_cast<uint32_t>(header_size),
                                static_cast<uint32_t>(count), stream, with_table ? wrap_end_of_image : wrap_header | wrap_end_of_image));
    JLS_CHECK(ensure(host_outcomes_, count * outcome_words * sizeof(uint64_t), true));
    JLS_CUDA(cudaMemcpyAsync(host_outcomes_.data, outcomes_.data, count * outcome_words * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                             stream));
    return 0;
}

int32_t Engine::encode_batch_end(BatchFrame* frames, size_t count, size_t header_size, CUstream_st* user_stream)
{
    cudaStream_t stream = user_stream ? user_stream : stream_;
    JLS_CHECK(wait_for(stream));
    read_coder_time();
    last_launches_ = static_cast<uint32_t>(thread_kernel_launch_count() - batch_launches_before_);

    int32_t first_error = 0;
    const uint64_t* outcomes = static_cast<const uint64_t*>(host_outcomes_.data);
    for (size_t i = 0; i < count; ++i)
    {
        const uint64_t* outcome = outcomes + i * outcome_words;
        const bool room = frames[i].stream_capacity >= header_size + 2;
        const size_t capacity = room ? frames[i].stream_capacity - header_size - 2 : 0;
        int32_t status = outcome[0] != ~0ULL ? static_cast<int32_t>(outcome[0] & 0xFF) : 0;
        if (status == 0 && outcome[1] > capacity)
            status = err_destination_too_small;
        frames[i].status = status;
        frames[i].stream_size = status == 0 ? header_size + static_cast<size_t>(outcome[1]) + 2 : 0;
        if (status != 0 && first_error == 0)
            first_error = status;
    }
    return first_error;
}

// ---------------------------------------------------------------------------------------------------------------------
// Batch path: host-resident frames, staged through device memory chunk by chunk
// ---------------------------------------------------------------------------------------------------------------------
// The pipeline of the host-resident batch calls:
//
//   copy_in_  (stream)   H2D chunk j+2 | H2D chunk j+3 | ...
//   compute   (2 engines) kernels chunk j (this engine) / chunk j+1 (helper engine), alternating, each on its own stream
//   copy_out_ (stream)   D2H chunk j-1 | D2H chunk j   | ...
//
// over `staging_slots` device staging slots.  A chunk is small (a few frames): its kernels take the latency of one line
// however few frames it holds (~1.3 ms for 4096 samples), so two chunks are coded at the same time on two engines (own
// scratch buffers, own stream) while two more are being uploaded; the host only ever waits for the OLDEST chunk in flight
// (it needs the stream sizes before it can issue their copies).  Round 1 coded one chunk at a time on one stream with two
// slots and reached 52 % of the PCIe bound.
int32_t Engine::prepare_staging()
{
    JLS_CHECK(prepare());
    if (!copy_in_)
        JLS_CUDA(cudaStreamCreateWithFlags(&copy_in_, cudaStreamNonBlocking));
    if (!copy_out_)
        JLS_CUDA(cudaStreamCreateWithFlags(&copy_out_, cudaStreamNonBlocking));
    for (auto& event : in_done_)
        if (!event)
            JLS_CUDA(cudaEventCreateWithFlags(&event, cudaEventDisableTiming));
    for (auto& event : out_done_)
        if (!event)
            JLS_CUDA(cudaEventCreateWithFlags(&event, cudaEventDisableTiming));
    if (!helper_)
        helper_ = new (std::nothrow) Engine;
    if (!helper_)
        return errc_not_enough_memory;
    JLS_CHECK(helper_->prepare());
    return 0;
}

// Frames per staging slot: enough to amortise the per-chunk host work (a wait, a dozen driver calls), few enough that the
// pipeline fills and drains quickly and four slots plus two engines' scratch stay small.  CHARLS_B200_HOST_CHUNK overrides.
size_t Engine::staging_chunk(size_t count, size_t bytes_per_frame) noexcept
{
    static const size_t forced = [] {
        const char* value = std::getenv("CHARLS_B200_HOST_CHUNK");
        return value ? static_cast<size_t>(std::strtoul(value, nullptr, 10)) : size_t{0};
    }();
    constexpr size_t slot_budget = size_t{256} << 20;
    size_t largest = slot_budget / (bytes_per_frame ? bytes_per_frame : 1);
    largest = largest < 1 ? 1 : largest > 64 ? 64 : largest;
    size_t chunk = forced ? forced : 4;
    chunk = chunk > largest ? largest : chunk;
    return chunk > count ? count : chunk;
}

int32_t Engine::encode_batch_host(const CodecParams& p, const uint8_t* header, size_t header_size, BatchFrame* frames, size_t count,
                                  size_t stride, const StreamOffsetTable* table)
{
    if (count == 0)
        return 0;
    JLS_CHECK(prepare_staging());
    const size_t row_bytes = row_bytes_of(p);
    const size_t frame_bytes = stride * (static_cast<size_t>(p.height) - 1) + row_bytes;
    // Rows whose distance is not a multiple of four bytes would take the kernels without shared-memory tiles (half the speed):
    // the staged copy gets an aligned pitch instead, the copy engine does the re-pitching on the way in.
    const size_t staged_stride = stride % 4 == 0 ? stride : align_up(row_bytes, 16);
    const size_t frame_pitch = align_up(staged_stride * static_cast<size_t>(p.height), 256); // distance between staged frames
    const size_t slot_bytes = worst_case_interval_bytes(p, p.lines_per_interval);
    size_t stream_slot = header_size + static_cast<size_t>(p.interval_count) * (slot_bytes + 2) + 2; // nothing is longer
    size_t largest_capacity = 0;
    for (size_t i = 0; i < count; ++i)
        largest_capacity = frames[i].stream_capacity > largest_capacity ? frames[i].stream_capacity : largest_capacity;
    stream_slot = align_up(stream_slot < largest_capacity ? stream_slot : largest_capacity, 256);
    const size_t chunk = staging_chunk(count, frame_pitch + stream_slot);
    for (int s = 0; s < staging_slots; ++s)
    {
        JLS_CHECK(ensure(stage_pixels_[s], chunk * frame_pitch + 64));
        JLS_CHECK(ensure(stage_streams_[s], chunk * stream_slot + 64));
    }

    const size_t chunks = (count + chunk - 1) / chunk;
    std::vector<BatchFrame> staged[staging_slots];
    for (auto& v : staged)
        v.resize(chunk);
    const auto frames_in = [&](size_t j) { return (count - j * chunk < chunk) ? count - j * chunk : chunk; };
    const auto engine_of = [&](size_t j) -> Engine& { return (j & 1U) ? *helper_ : *this; };

    const auto upload = [&](size_t j) -> int32_t {
        const int s = static_cast<int>(j % staging_slots);
        // the slot is free once the stream copies of the chunk that used it before have left it
        JLS_CUDA(cudaStreamWaitEvent(copy_in_, out_done_[s], 0));
        const size_t first = j * chunk, n = frames_in(j);
        for (size_t k = 0; k < n; ++k)
        {
            uint8_t* staged_frame = static_cast<uint8_t*>(stage_pixels_[s].data) + k * frame_pitch;
            if (staged_stride == stride)
                JLS_CUDA(cudaMemcpyAsync(staged_frame, frames[first + k].pixels, frame_bytes, cudaMemcpyHostToDevice, copy_in_));
            else
                JLS_CUDA(cudaMemcpy2DAsync(staged_frame, staged_stride, frames[first + k].pixels, stride, row_bytes,
                                           static_cast<size_t>(p.height), cudaMemcpyHostToDevice, copy_in_));
        }
        JLS_CUDA(cudaEventRecord(in_done_[s], copy_in_));
        return 0;
    };
    const auto code = [&](size_t j) -> int32_t {
        const int s = static_cast<int>(j % staging_slots);
        const size_t first = j * chunk, n = frames_in(j);
        for (size_t k = 0; k < n; ++k)
        {
            const size_t capacity = frames[first + k].stream_capacity < stream_slot ? frames[first + k].stream_capacity : stream_slot;
            staged[s][k] = BatchFrame{static_cast<uint8_t*>(stage_pixels_[s].data) + k * frame_pitch,
                                      static_cast<uint8_t*>(stage_streams_[s].data) + k * stream_slot, capacity, 0, 0, 0};
        }
        Engine& engine = engine_of(j);
        JLS_CUDA(cudaStreamWaitEvent(engine.stream_, in_done_[s], 0));
        return engine.encode_batch_begin(p, header, header_size, staged[s].data(), n, staged_stride, engine.stream_, table);
    };
    int32_t first_error = 0;
    uint32_t launches = 0;
    float coder_ms = 0.0F;
    const auto complete = [&](size_t j) -> int32_t {
        const int s = static_cast<int>(j % staging_slots);
        const size_t first = j * chunk, n = frames_in(j);
        Engine& engine = engine_of(j);
        const int32_t status = engine.encode_batch_end(staged[s].data(), n, header_size, engine.stream_);
        launches += engine.last_launches_;
        coder_ms += engine.last_coder_ms_;
        if (status != 0 && first_error == 0)
            first_error = status;
        for (size_t k = 0; k < n; ++k)
        {
            frames[first + k].status = staged[s][k].status;
            frames[first + k].stream_size = staged[s][k].stream_size;
            if (staged[s][k].status == 0)
                JLS_CUDA(cudaMemcpyAsync(frames[first + k].stream, staged[s][k].stream, staged[s][k].stream_size, cudaMemcpyDeviceToHost,
                                         copy_out_));
        }
        JLS_CUDA(cudaEventRecord(out_done_[s], copy_out_));
        return 0;
    };

    for (int s = 0; s < staging_slots; ++s)
        JLS_CUDA(cudaEventRecord(out_done_[s], copy_out_)); // all slots start free
    JLS_CHECK(upload(0));
    if (chunks > 1)
        JLS_CHECK(upload(1));
    for (size_t j = 0; j < chunks; ++j)
    {
        JLS_CHECK(code(j));
        if (j + 2 < chunks)
            JLS_CHECK(upload(j + 2));
        if (j >= 1)
            JLS_CHECK(complete(j - 1));
    }
    JLS_CHECK(complete(chunks - 1));
    JLS_CHECK(wait_for(copy_out_));
    last_launches_ = launches;
    last_coder_ms_ = coder_ms;
    return first_error;
}

int32_t Engine::decode_batch_host(const CodecParams& p, BatchFrame* frames, size_t count, size_t stride)
{
    if (count == 0)
        return 0;
    JLS_CHECK(prepare_staging());
    bool with_tables[staging_slots] = {};
    const size_t row_bytes = row_bytes_of(p);
    const size_t frame_bytes = stride * (static_cast<size_t>(p.height) - 1) + row_bytes;
    const size_t staged_stride = stride % 4 == 0 ? stride : align_up(row_bytes, 16); // see encode_batch_host
    const size_t frame_pitch = align_up(staged_stride * static_cast<size_t>(p.height), 256);
    size_t stream_slot = 0;
    for (size_t i = 0; i < count; ++i)
        stream_slot = frames[i].stream_capacity > stream_slot ? frames[i].stream_capacity : stream_slot;
    stream_slot = align_up(stream_slot + 16, 256);
    const size_t chunk = staging_chunk(count, frame_pitch + stream_slot);
    for (int s = 0; s < staging_slots; ++s)
    {
        JLS_CHECK(ensure(stage_pixels_[s], chunk * frame_pitch + 64));
        JLS_CHECK(ensure(stage_streams_[s], chunk * stream_slot + 64));
    }

    const size_t chunks = (count + chunk - 1) / chunk;
    std::vector<BatchFrame> staged[staging_slots];
    for (auto& v : staged)
        v.resize(chunk);
    const auto frames_in = [&](size_t j) { return (count - j * chunk < chunk) ? count - j * chunk : chunk; };
    const auto engine_of = [&](size_t j) -> Engine& { return (j & 1U) ? *helper_ : *this; };

    const auto upload = [&](size_t j) -> int32_t {
        const int s = static_cast<int>(j % staging_slots);
        JLS_CUDA(cudaStreamWaitEvent(copy_in_, out_done_[s], 0));
        const size_t first = j * chunk, n = frames_in(j);
        for (size_t k = 0; k < n; ++k)
            JLS_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(stage_streams_[s].data) + k * stream_slot, frames[first + k].stream,
                                     frames[first + k].stream_capacity, cudaMemcpyHostToDevice, copy_in_));
        JLS_CUDA(cudaEventRecord(in_done_[s], copy_in_));
        return 0;
    };
    const auto code = [&](size_t j) -> int32_t {
        const int s = static_cast<int>(j % staging_slots);
        const size_t first = j * chunk, n = frames_in(j);
        for (size_t k = 0; k < n; ++k)
            staged[s][k] = BatchFrame{static_cast<uint8_t*>(stage_pixels_[s].data) + k * frame_pitch,
                                      static_cast<uint8_t*>(stage_streams_[s].data) + k * stream_slot, frames[first + k].stream_capacity, 0, 0,
                                      frames[first + k].scan_offset, frames[first + k].table};
        Engine& engine = engine_of(j);
        JLS_CUDA(cudaStreamWaitEvent(engine.stream_, in_done_[s], 0));
        with_tables[s] = all_frames_have_tables(p, staged[s].data(), n);
        return engine.decode_batch_begin(p, staged[s].data(), n, staged_stride, engine.stream_, with_tables[s]);
    };
    int32_t first_error = 0;
    uint32_t launches = 0;
    float coder_ms = 0.0F;
    const auto complete = [&](size_t j) -> int32_t {
        const int s = static_cast<int>(j % staging_slots);
        const size_t first = j * chunk, n = frames_in(j);
        Engine& engine = engine_of(j);
        int32_t status = engine.decode_batch_end(staged[s].data(), n, engine.stream_);
        if (status != 0 && with_tables[s])
        {
            // not a clean decode with the side tables: once more by marker search (decode_scan_to_host_end)
            JLS_CHECK(engine.decode_batch_begin(p, staged[s].data(), n, staged_stride, engine.stream_, false));
            status = engine.decode_batch_end(staged[s].data(), n, engine.stream_);
        }
        launches += engine.last_launches_;
        coder_ms += engine.last_coder_ms_;
        if (status != 0 && first_error == 0)
            first_error = status;
        for (size_t k = 0; k < n; ++k)
        {
            frames[first + k].status = staged[s][k].status;
            frames[first + k].stream_size = staged[s][k].stream_size;
            if (staged[s][k].status != 0)
                continue;
            if (staged_stride == stride)
                JLS_CUDA(cudaMemcpyAsync(frames[first + k].pixels, staged[s][k].pixels, frame_bytes, cudaMemcpyDeviceToHost, copy_out_));
            else
                JLS_CUDA(cudaMemcpy2DAsync(frames[first + k].pixels, stride, staged[s][k].pixels, staged_stride, row_bytes,
                                           static_cast<size_t>(p.height), cudaMemcpyDeviceToHost, copy_out_));
        }
        JLS_CUDA(cudaEventRecord(out_done_[s], copy_out_));
        return 0;
    };

    for (int s = 0; s < staging_slots; ++s)
        JLS_CUDA(cudaEventRecord(out_done_[s], copy_out_));
    JLS_CHECK(upload(0));
    if (chunks > 1)
        JLS_CHECK(upload(1));
    for (size_t j = 0; j < chunks; ++j)
    {
        JLS_CHECK(code(j));
        if (j + 2 < chunks)
            JLS_CHECK(upload(j + 2));
        if (j >= 1)
            JLS_CHECK(complete(j - 1));
    }
    JLS_CHECK(complete(chunks - 1));
    JLS_CHECK(wait_for(copy_out_));
    last_launches_ = launches;
    last_coder_ms_ = coder_ms;
    return first_error;
}

int32_t Engine::download_prefixes(const BatchFrame* frames, size_t count, uint32_t prefix_bytes, const uint8_t*& prefixes,
                                  CUstream_st* user_stream)
{
    JLS_CHECK(prepare());
    cudaStream_t stream = user_stream ? user_stream : stream_;
    const size_t table_bytes = count * (sizeof(void*) + sizeof(size_t));
    JLS_CHECK(ensure(pointer_table_, table_bytes));
    JLS_CHECK(ensure(host_pointer_table_, table_bytes, true));
    JLS_CHECK(ensure(prefixes_, count * prefix_bytes));
    JLS_CHECK(ensure(host_prefixes_, count * prefix_bytes, true));
    auto** host_pointers = static_cast<const uint8_t**>(host_pointer_table_.data);
    auto* host_sizes = reinterpret_cast<size_t*>(host_pointers + count);
    for (size_t i = 0; i < count; ++i)
    {
        host_pointers[i] = frames[i].stream;
        host_sizes[i] = frames[i].stream_capacity;
    }
    JLS_CUDA(cudaMemcpyAsync(pointer_table_.data, host_pointer_table_.data, table_bytes, cudaMemcpyHostToDevice, stream));
    auto** device_pointers = static_cast<const uint8_t**>(pointer_table_.data);
    JLS_CUDA(launch_copy_prefixes(device_pointers, reinterpret_cast<const size_t*>(device_pointers + count),
                                  static_cast<uint8_t*>(prefixes_.data), prefix_bytes, static_cast<uint32_t>(count), stream));
    JLS_CUDA(cudaMemcpyAsync(host_prefixes_.data, prefixes_.data, count * prefix_bytes, cudaMemcpyDeviceToHost, stream));
    JLS_CHECK(wait_for(stream));
    prefixes = static_cast<const uint8_t*>(host_prefixes_.data); // parsed where they are (pinned; the next call reuses the buffer)
    return 0;
}

int32_t Engine::decode_batch(const CodecParams& p, BatchFrame* frames, size_t count, size_t stride, CUstream_st* user_stream,
                             const std::function<int32_t()>& while_coding)
{
    if (count == 0)
        return prepare();
    const bool use_tables = all_frames_have_tables(p, frames, count);
    JLS_CHECK(decode_batch_begin(p, frames, count, stride, user_stream, use_tables));
    if (while_coding)
        JLS_CHECK(while_coding());
    const int32_t status = decode_batch_end(frames, count, user_stream);
    if (status == 0 || !use_tables)
        return status;
    // not a clean decode with the side tables: the answer is what the marker search gives (decode_scan_to_host_end)
    JLS_CHECK(decode_batch_begin(p, frames, count, stride, user_stream, false));
    return decode_batch_end(frames, count, user_stream);
}

bool Engine::all_frames_have_tables(const CodecParams& p, const BatchFrame* frames, size_t count) noexcept
{
    static const bool disabled = [] {
        const char* value = std::getenv("CHARLS_B200_IGNORE_OFFSET_TABLES");
        return value && value[0] == '1';
    }();
    if (disabled)
        return false;
    for (size_t i = 0; i < count; ++i)
        if (frames[i].table.total != p.interval_count + 1U)
            return false;
    return count != 0;
}

int32_t Engine::decode_batch_begin(const CodecParams& p, const BatchFrame* frames, size_t count, size_t stride,
                                   CUstream_st* user_stream, bool use_tables)
{
    JLS_CHECK(prepare());
    cudaStream_t stream = user_stream ? user_stream : stream_;
    batch_launches_before_ = thread_kernel_launch_count();

    size_t max_remaining = 0;
    // the tile kernels keep row offsets of up to three strides in 32 bits (jls_tile.cuh, TileWalk)
    bool word_aligned = stride % 4 == 0 && stride < (size_t{1} << 30);
    std::vector<ScanJob> jobs(count);
    for (size_t i = 0; i < count; ++i)
    {
        ScanJob& job = jobs[i];
        job = ScanJob{};
        word_aligned = word_aligned && reinterpret_cast<uintptr_t>(frames[i].pixels) % 4 == 0;
        job.pixels_out = frames[i].pixels;
        job.stride = stride;
        job.stream_in = frames[i].stream + frames[i].scan_offset;
        job.stream_in_size = frames[i].stream_capacity - frames[i].scan_offset;
        max_remaining = job.stream_in_size > max_remaining ? job.stream_in_size : max_remaining;
        if (use_tables)
        {
            job.offset_table.total = frames[i].table.total;
            for (uint32_t segment = 0; segment < offset_table_segment_count(frames[i].table.total); ++segment)
                job.offset_table.entries[segment] = frames[i].stream + frames[i].table.entry_offsets[segment];
        }
    }
    const bool through_scratch = !word_aligned && use_fast_path(p) && p.interleave != ilv_line;
    const size_t scratch_row_bytes = row_bytes_of(p);
    const size_t scratch_pitch = align_up(scratch_row_bytes, 16), scratch_span = align_up(scratch_pitch * static_cast<size_t>(p.height), 256);
    if (through_scratch)
    {
        // decoded at an aligned pitch, copied to the caller's rows afterwards (see k_repitch)
        JLS_CHECK(ensure(repitched_, count * scratch_span + 64));
        for (size_t i = 0; i < count; ++i)
        {
            jobs[i].pixels_out = static_cast<uint8_t*>(repitched_.data) + i * scratch_span;
            jobs[i].stride = scratch_pitch;
        }
        word_aligned = scratch_pitch < (size_t{1} << 30);
    }
    JLS_CHECK(ensure(marker_counts_, marker_scratch_bytes(count, max_remaining)));
    JLS_CHECK(ensure(marker_totals_, count * sizeof(uint32_t)));
    JLS_CHECK(ensure(marker_codes_, count * p.interval_count));
    JLS_CHECK(stage_jobs(p, jobs, false, 0, stream));
    JLS_CUDA(launch_decode(p, static_cast<const ScanJob*>(job_table_.data), static_cast<uint32_t>(count), max_remaining,
                           static_cast<uint32_t*>(marker_counts_.data), static_cast<uint32_t*>(marker_totals_.data),
                           static_cast<uint8_t*>(marker_codes_.data), stream, events_, word_aligned, use_tables));
    if (through_scratch)
        JLS_CHECK(repitch(frames, count, stride, scratch_pitch, scratch_span, static_cast<uint32_t>(scratch_row_bytes), p.height, false,
                          stream));
    JLS_CHECK(ensure(host_outcomes_, count * outcome_words * sizeof(uint64_t), true));
    JLS_CUDA(cudaMemcpyAsync(host_outcomes_.data, outcomes_.data, count * outcome_words * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                             stream));
    return 0;
}

int32_t Engine::decode_batch_end(BatchFrame* frames, size_t count, CUstream_st* user_stream)
{
    cudaStream_t stream = user_stream ? user_stream : stream_;
    JLS_CHECK(wait_for(stream));
    read_coder_time();
    last_launches_ = static_cast<uint32_t>(thread_kernel_launch_count() - batch_launches_before_);

    int32_t first_error = 0;
    const uint64_t* outcomes = static_cast<const uint64_t*>(host_outcomes_.data);
    for (size_t i = 0; i < count; ++i)
    {
        const uint64_t* outcome = outcomes + i * outcome_words;
        int32_t status = outcome[0] != ~0ULL ? static_cast<int32_t>(outcome[0] & 0xFF) : 0;
        if (status == 0 && outcome[2] != 0xD9)
            status = 24; // end_of_image_marker_not_found (reference src/jpeg_stream_reader.cpp:152-172)
        frames[i].status = status;
        frames[i].stream_size = frames[i].scan_offset + static_cast<size_t>(outcome[1]) + 2;
        if (status != 0 && first_error == 0)
            first_error = status;
    }
    return first_error;
}

} // namespace jls
