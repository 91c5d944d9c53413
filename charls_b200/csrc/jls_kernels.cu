// jls_kernels.cu -- sm_100a kernels of the JPEG-LS scan engine and their launch wrappers.
//
// Pipeline (DESIGN.md "Kernels"):
//   encode:  k_encode_fast / k_encode_general  -> one slot of entropy bytes per restart interval
//            k_scan_offsets                    -> exclusive scan of the interval sizes (+2 per RSTm marker)
//            k_gather                          -> concatenation of the intervals, RSTm markers inserted
//   decode:  k_marker_count / k_marker_scan / k_marker_write -> ordered table of the markers that delimit intervals
//            k_decode_fast / k_decode_general  -> samples
//            k_decode_finish                   -> marker-sequence validation, bytes consumed by the scan
// A thread owns one restart interval from its first to its last bit: the adaptive coder is a strict serial chain in x
// (every symbol's context state and, in the decoder, bit position depend on the previous symbol), so the parallel
// axes are intervals x images.  There is no dense contraction anywhere: the tensor cores are not used.
#include "jls_kernels.hpp"

#include "jls_interval.cuh"
#include "jls_tile.cuh"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>

namespace jls {

namespace {

#if !defined(JLS_FAST_BLOCK_THREADS)
#define JLS_FAST_BLOCK_THREADS 32
#endif
// Pixel loop bodies written out per drain / top-up group for the full-depth one-component kernels.  Measured on cfg2 (A/B on
// one box, profiles/r2_notes.md): encoder 6.27 -> 6.02 ms, decoder 7.17 -> 7.67 ms (its body is four times the encoder's
// size and no longer sits well in the instruction caches) -- so the encoder unrolls and the decoder does not.
#if !defined(JLS_UNROLL_ENCODER)
#define JLS_UNROLL_ENCODER 1
#endif
#if !defined(JLS_UNROLL_DECODER)
#define JLS_UNROLL_DECODER 0
#endif
// 1: the encoder's tiles are loaded with TMA bulk copies instead of cp.async (A/B, see jls_tile.cuh)
#if !defined(JLS_TILE_LOAD_BULK)
#define JLS_TILE_LOAD_BULK 0
#endif
constexpr int fast_block_threads = JLS_FAST_BLOCK_THREADS;
constexpr int general_block_threads = 32;

std::atomic<uint64_t> g_kernel_launches{0};
thread_local uint64_t t_kernel_launches = 0; // launches issued by this host thread

// first error in stream order wins: key = (interval << 8) | errc, smaller interval first
__device__ __forceinline__ void report_error(const ScanJob& job, uint32_t interval, int32_t errc)
{
    const unsigned long long key = (static_cast<unsigned long long>(interval) << 8) | static_cast<unsigned long long>(errc);
    atomicMin(reinterpret_cast<unsigned long long*>(job.status), key);
}

// ---------------------------------------------------------------------------------------------------------------------
// Fast path, restart interval = 1 line: one thread per line, a warp holds 32 consecutive lines.  The five regular
// contexts of a thread live in shared memory as [context][thread] so that a warp-wide 16-byte access is conflict free.
// ---------------------------------------------------------------------------------------------------------------------
template<int NC, bool LOSSLESS, typename S, bool LINE_ILV>
__global__ void __launch_bounds__(fast_block_threads)
    k_encode_fast(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs, size_t slot_bytes)
{
    __shared__ RegularContext contexts[5 * fast_block_threads];
    const ScanJob& job = jobs[blockIdx.y];
    const uint32_t interval = blockIdx.x * fast_block_threads + threadIdx.x;
    if (interval >= p.interval_count)
        return;
    const IntervalResult r =
        encode_interval_fast<NC, LOSSLESS, S, LINE_ILV>(p, job, interval, contexts + threadIdx.x, fast_block_threads, slot_bytes);
    job.interval_bytes[interval] = r.bytes;
    if (r.errc != err_none)
        report_error(job, interval, r.errc);
}

template<int NC, bool LOSSLESS, typename S, bool LINE_ILV>
__global__ void __launch_bounds__(fast_block_threads)
    k_decode_fast(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs)
{
    __shared__ RegularContext contexts[5 * fast_block_threads];
    const ScanJob& job = jobs[blockIdx.y];
    const uint32_t interval = blockIdx.x * fast_block_threads + threadIdx.x;
    if (interval >= p.interval_count)
        return;
    const IntervalResult r = decode_interval_fast<NC, LOSSLESS, S, LINE_ILV>(p, job, interval, contexts + threadIdx.x, fast_block_threads);
    if (r.errc != err_none)
        report_error(job, interval, r.errc);
}

// ---------------------------------------------------------------------------------------------------------------------
// Fast path with shared-memory tiles (jls_tile.cuh): the same per-line codec, but the samples travel between HBM and the
// lanes as coalesced [32 lines x 64/96 bytes] tiles instead of per-lane byte accesses.  Needs 4-byte aligned rows.
// ---------------------------------------------------------------------------------------------------------------------
// sample `index` (an immediate) at shared-window address `address`
template<typename S>
__device__ __forceinline__ int32_t load_shared_sample(uint32_t address, int32_t index)
{
    uint32_t v;
    if (sizeof(S) == 1)
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(address + static_cast<uint32_t>(index)) : "memory");
    else
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(address + 2U * static_cast<uint32_t>(index)) : "memory");
    return static_cast<int32_t>(v);
}

template<typename S>
__device__ __forceinline__ void store_shared_sample(uint32_t address, int32_t index, int32_t value)
{
    if (sizeof(S) == 1)
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(address + static_cast<uint32_t>(index)), "r"(value) : "memory");
    else
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(address + 2U * static_cast<uint32_t>(index)), "r"(value) : "memory");
}

#if !defined(JLS_NARROW_TILES)
#define JLS_NARROW_TILES 1
#endif
// ENCODER: the encoder double-buffers its tiles; 32-byte tiles for one-component lines (and four instead of five context
// rows) bring a block's shared memory under the 7 KB that let 32 blocks live on an SM, 48-byte tiles take the
// three-component encoders from 20 to 27 blocks (JLS_NARROW_TILES).
template<int NC, typename S = uint16_t, bool ENCODER = false>
struct TileShape
{
    // 96 (48) bytes hold whole pixels for 3 x 8 and 3 x 16 bit
    static constexpr int words = NC == 3 ? ((JLS_NARROW_TILES && ENCODER) ? 12 : 24)
                                         : (JLS_NARROW_TILES && ENCODER && NC == 1) ? 8 : 16;
    static constexpr int stride_words = words + 1;
};

// DEPTH = bits per sample when lossless data fills its container (8 in uint8_t, 16 in uint16_t): depth, MAXVAL and the
// sign extension of the error value become immediates and no sample needs masking; 0 = any depth (h.bits).
// One-component encoders are asked to fit 32 blocks on an SM (64 registers; shared memory: see TileShape), the others keep
// what ptxas picks (a cap costs the three-component encoders spills, profiles/r1_notes.md).
#if !defined(JLS_ENCODE_MIN_BLOCKS_NC1)
#define JLS_ENCODE_MIN_BLOCKS_NC1 32
#endif
template<int NC, bool LOSSLESS, typename S, int DEPTH>
__global__ void __launch_bounds__(fast_block_threads, NC == 1 ? JLS_ENCODE_MIN_BLOCKS_NC1 : NC == 3 ? 26 : 28)
    k_encode_tiled(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs, size_t slot_bytes)
{
    // JLS_TILE_LOAD_BULK (A/B, jls_tile.cuh): rows land 16-byte aligned (TW + 4 words apart) through TMA bulk copies
    constexpr bool bulk = JLS_TILE_LOAD_BULK != 0;
    constexpr int TW = TileShape<NC, S, true>::words, SW = bulk ? TW + 4 : TileShape<NC, S, true>::stride_words;
    constexpr int pixels_per_tile = TW * 4 / static_cast<int>(sizeof(S)) / NC;
    constexpr int warps = fast_block_threads / 32;
    constexpr int first_context = NC == 1 ? 1 : 0; // scalar lines never use context 0 (FastLineState::first_context)
    // Two tile buffers: the next tile is in flight while this one is coded.  (One buffer and a wait per tile leaves room
    // for more resident blocks -- 31 instead of 21 for three-component pixels -- and measured 3 % slower.)
    __shared__ RegularContext contexts[(5 - first_context) * fast_block_threads];
    __shared__ __align__(16) uint32_t tiles[warps][2][32 * SW];
    __shared__ uint64_t tile_ready[warps][2]; // bulk: one mbarrier per tile buffer
    extern __shared__ uint8_t context_lut[]; // lut_last + 1 entries, sized at launch (tiled_dynamic_shared_bytes)

    // the host only picks this kernel when T3 fits; 8-bit containers get an entry for every sample value (no clamp)
    const int32_t lut_last = sizeof(S) == 1 ? 255 : min(p.t3, context_lut_capacity - 1);
    for (int32_t i = threadIdx.x; i <= lut_last; i += fast_block_threads)
        context_lut[i] = context_lut_entry(p, i);
    __shared__ uint32_t reciprocal_lut[reciprocal_lut_capacity]; // the host only picks this kernel when RESET <= 64
    for (int32_t i = threadIdx.x; i < reciprocal_lut_capacity; i += fast_block_threads)
        reciprocal_lut[i] = reciprocal_lut_entry(i);
    __syncthreads();

    const ScanJob& job = jobs[blockIdx.y];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t first_line = (blockIdx.x * warps + warp) * 32U;
    if (first_line >= p.interval_count)
        return; // whole warp
    const uint32_t interval = first_line + lane;
    const bool active = interval < p.interval_count;

    __shared__ int32_t hot_scratch[warps][hot_scratch_words];
    HotParams h = make_hot_params(p);
    h.context_lut = context_lut;
    h.context_lut_last = lut_last;
    h.context_lut_shared = static_cast<uint32_t>(__cvta_generic_to_shared(context_lut));
    h.reciprocal_lut = reciprocal_lut;
    h.reciprocal_lut_shared = static_cast<uint32_t>(__cvta_generic_to_shared(reciprocal_lut));
    keep_hot_params_in_registers(h, hot_scratch[warp]);
    if (DEPTH != 0)
    {
        // lossless at full depth: RANGE = 2^DEPTH, qbpp = DEPTH, LIMIT = 2 (DEPTH + max(8, DEPTH)) (make_codec_params)
        h.bits = DEPTH;
        h.maxval = (1 << DEPTH) - 1;
        h.qbpp = DEPTH;
        h.limit = 2 * (DEPTH + (DEPTH > 8 ? DEPTH : 8));
        h.escape = h.limit - DEPTH - 1;
    }
    // deferred flushing unless nearly every sample fills a word anyway (lossless 16-bit data)
    FastLineEncoder<NC, LOSSLESS, sizeof(S) == 1 ? lut_full : lut_clamped, writer_mode<NC, LOSSLESS, S>, DEPTH> enc;
    constexpr int32_t drain_mask = decltype(enc)::pixels_per_drain - 1;
    // JLS_UNROLL_ENCODER: the pixel loop's body written out once per pixel of a drain group (full-depth one-component kernels)
    constexpr bool unrolled_groups = JLS_UNROLL_ENCODER && NC == 1 && DEPTH != 0 && drain_mask > 0;
    uint8_t* slot = job.slots + static_cast<size_t>(active ? interval : first_line) * slot_bytes;
    assume_global(slot);
    enc.begin(h, contexts + threadIdx.x - first_context * fast_block_threads, fast_block_threads, slot);

    const uint8_t* pixels = job.pixels_in;
    assume_global(pixels);
    const size_t stride = job.stride;
    const int32_t width = p.width;
    const int32_t row_bytes = width * NC * static_cast<int32_t>(sizeof(S));
    const int32_t tile_count = (row_bytes + TW * 4 - 1) / (TW * 4);
    const uint32_t last_line = p.interval_count - 1;
    const int32_t transform = h.transform;
    const bool mask_needed = DEPTH == 0 && h.bits != static_cast<int32_t>(8 * sizeof(S));

    // bulk copies need whole tiles, all 32 lines and 16-byte aligned rows; the last (partial) tile of a line and everything
    // else goes through cp.async
    const bool bulk_rows = bulk && first_line + 31U <= last_line && stride % 16 == 0 && reinterpret_cast<uintptr_t>(pixels) % 16 == 0;
    const auto load_tile = [&](int32_t index) {
        if (bulk_rows && (index + 1) * (TW * 4) <= row_bytes)
            tile_load_bulk<TW>(static_cast<uint32_t>(__cvta_generic_to_shared(tiles[warp][index & 1])), SW * 4U,
                               static_cast<uint32_t>(__cvta_generic_to_shared(&tile_ready[warp][index & 1])), pixels, stride, first_line,
                               index, lane);
        else
            tile_load_async<TW, SW>(tiles[warp][index & 1], pixels, stride, first_line, last_line, row_bytes, index, lane);
    };
    if (bulk)
    {
        if (lane < 2)
            mbarrier_init(static_cast<uint32_t>(__cvta_generic_to_shared(&tile_ready[warp][lane])), 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    load_tile(0);
    for (int32_t t = 0; t < tile_count; ++t)
    {
        if (t + 1 < tile_count)
        {
            if (bulk)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the buffer's last readers were ordinary loads
            load_tile(t + 1);
            cp_async_wait<1>();
        }
        else
        {
            cp_async_wait<0>();
        }
        if (bulk_rows && (t + 1) * (TW * 4) <= row_bytes)
            mbarrier_wait(static_cast<uint32_t>(__cvta_generic_to_shared(&tile_ready[warp][t & 1])), static_cast<uint32_t>(t >> 1) & 1U);
        __syncwarp();
        if (active)
        {
            // two loop registers: the sample's shared-memory address and a count-down that is loop condition and drain
            // cadence at once (pixels_per_tile is a multiple of 4: groups end where n is one); both advance on the FMA pipe
            uint32_t sample = static_cast<uint32_t>(__cvta_generic_to_shared(&tiles[warp][t & 1][lane * SW]));
            int32_t n = min(pixels_per_tile, width - t * pixels_per_tile);
            if (unrolled_groups && (n & drain_mask) == 0)
            {
                // Whole groups only (every tile of a line whose width is a multiple of the cadence): the group is written out,
                // its samples are read at immediate offsets and the two loop registers advance once per group.
                do
                {
                    enc.drain();
#pragma unroll
                    for (int32_t g = 0; g <= drain_mask; ++g)
                    {
                        int32_t v[NC];
#pragma unroll
                        for (int32_t c = 0; c < NC; ++c)
                            v[c] = load_shared_sample<S>(sample, g * NC + c);
                        enc.pixel(h, v);
                    }
                    sample = static_cast<uint32_t>(
                        add_fma(h, static_cast<int32_t>(sample), (drain_mask + 1) * NC * static_cast<int32_t>(sizeof(S))));
                    n = add_fma(h, n, -(drain_mask + 1));
                } while (n != 0);
            }
            else
            do
            {
                enc.drain(); // every fourth pixel (every pixel_per_drain-th), for all lanes of the warp together
#pragma unroll 1
                do
                {
                    int32_t v[NC];
#pragma unroll
                    for (int32_t c = 0; c < NC; ++c)
                        v[c] = load_shared_sample<S>(sample, c);
                    if (NC == 3 && transform != 0)
                    {
                        color_forward(transform, sizeof(S) == 2 ? 0xFFFF : 0xFF, v[0], v[NC > 1 ? 1 : 0], v[NC > 2 ? 2 : 0]);
                    }
                    else if (mask_needed)
                    {
#pragma unroll
                        for (int32_t c = 0; c < NC; ++c)
                            v[c] &= h.maxval;
                    }
                    enc.pixel(h, v);
                    sample = static_cast<uint32_t>(add_fma(h, static_cast<int32_t>(sample), NC * static_cast<int32_t>(sizeof(S))));
                    n = add_fma(h, n, -1);
                } while ((n & drain_mask) != 0); // the group test is the loop condition: no drain test per pixel
            } while (n != 0);
        }
        __syncwarp(); // everybody is done with this buffer before the copy two tiles ahead overwrites it
    }
    if (active)
    {
        enc.end_line();
        job.interval_bytes[interval] = enc.finish();
    }
}

template<int NC, bool LOSSLESS, typename S, int DEPTH>
__global__ void __launch_bounds__(fast_block_threads)
    k_decode_tiled(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs)
{
    constexpr int TW = TileShape<NC>::words, SW = TileShape<NC>::stride_words;
    constexpr int pixels_per_tile = TW * 4 / static_cast<int>(sizeof(S)) / NC;
    constexpr int warps = fast_block_threads / 32;
    // pixels between two top-ups of the 128-bit read window (see FastReader::get_golomb_steady)
    constexpr int refill_cadence = FastLineDecoder<NC, LOSSLESS, sizeof(S) == 1 ? lut_full : lut_clamped, DEPTH>::pixels_per_top_up;
    constexpr bool unrolled_groups = JLS_UNROLL_DECODER && NC == 1 && DEPTH != 0 && refill_cadence > 1;
    constexpr int first_context = NC == 1 ? 1 : 0; // scalar lines never use context 0 (FastLineState::first_context)
    __shared__ RegularContext contexts[(5 - first_context) * fast_block_threads];
    __shared__ uint32_t tiles[warps][32 * SW];
    extern __shared__ uint8_t context_lut[]; // lut_last + 1 entries, sized at launch (tiled_dynamic_shared_bytes)

    // the host only picks this kernel when T3 fits; 8-bit containers get an entry for every sample value (no clamp)
    const int32_t lut_last = sizeof(S) == 1 ? 255 : min(p.t3, context_lut_capacity - 1);
    for (int32_t i = threadIdx.x; i <= lut_last; i += fast_block_threads)
        context_lut[i] = context_lut_entry(p, i);
    __shared__ uint32_t reciprocal_lut[reciprocal_lut_capacity]; // the host only picks this kernel when RESET <= 64
    for (int32_t i = threadIdx.x; i < reciprocal_lut_capacity; i += fast_block_threads)
        reciprocal_lut[i] = reciprocal_lut_entry(i);
    __syncthreads();

    const ScanJob& job = jobs[blockIdx.y];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t first_line = (blockIdx.x * warps + warp) * 32U;
    if (first_line >= p.interval_count)
        return; // whole warp
    const uint32_t interval = first_line + lane;

    // interval_offset holds 2 entries per interval: [2i] = first byte, [2i+1] = end (first 0xFF of the closing marker)
    uint64_t begin = 0, end = 0;
    bool coding = interval < p.interval_count;
    bool closing_marker_found = true;
    if (coding)
    {
        begin = job.interval_offset[2 * static_cast<size_t>(interval)];
        end = job.interval_offset[2 * static_cast<size_t>(interval) + 1];
        closing_marker_found = end != ~0ULL;
        if (!closing_marker_found)
            end = job.stream_in_size; // decode what is there (the reference runs dry -> invalid_data)
        coding = begin != ~0ULL && begin <= end; // else an earlier marker is missing: k_decode_finish reports it
    }
    const uint32_t row_mask = __ballot_sync(0xFFFFFFFFU, coding);

    __shared__ int32_t hot_scratch[warps][hot_scratch_words];
    HotParams h = make_hot_params(p);
    h.context_lut = context_lut;
    h.context_lut_last = lut_last;
    h.context_lut_shared = static_cast<uint32_t>(__cvta_generic_to_shared(context_lut));
    h.reciprocal_lut = reciprocal_lut;
    h.reciprocal_lut_shared = static_cast<uint32_t>(__cvta_generic_to_shared(reciprocal_lut));
    keep_hot_params_in_registers(h, hot_scratch[warp]);
    if (DEPTH != 0)
    {
        // lossless at full depth: RANGE = 2^DEPTH, qbpp = DEPTH, LIMIT = 2 (DEPTH + max(8, DEPTH)) (make_codec_params)
        h.bits = DEPTH;
        h.maxval = (1 << DEPTH) - 1;
        h.qbpp = DEPTH;
        h.limit = 2 * (DEPTH + (DEPTH > 8 ? DEPTH : 8));
        h.escape = h.limit - DEPTH - 1;
    }
    FastLineDecoder<NC, LOSSLESS, sizeof(S) == 1 ? lut_full : lut_clamped, DEPTH> dec;
    const uint8_t* stream = job.stream_in;
    assume_global(stream);
    dec.begin(h, contexts + threadIdx.x - first_context * fast_block_threads, fast_block_threads, stream + (coding ? begin : 0),
              stream + (coding ? end : 0));

    uint8_t* pixels = job.pixels_out;
    assume_global(pixels);
    const size_t stride = job.stride;
    const int32_t width = p.width;
    const int32_t row_bytes = width * NC * static_cast<int32_t>(sizeof(S));
    const int32_t tile_count = (row_bytes + TW * 4 - 1) / (TW * 4);
    const int32_t transform = h.transform;
    uint32_t* tile = tiles[warp];

    for (int32_t t = 0; t < tile_count; ++t)
    {
        if (coding)
        {
            uint32_t sample = static_cast<uint32_t>(__cvta_generic_to_shared(&tile[lane * SW]));
            const int32_t x0 = t * pixels_per_tile;
            // two loop registers: the sample pointer and a count-down that is loop condition and refill cadence at once;
            // the pixels left in the line are needed on the rare run-mode path only
            const int32_t beyond = max(width - x0 - pixels_per_tile, 0); // pixels of the line after this tile
            int32_t n = width - x0 - beyond;
            if (unrolled_groups && (n & (refill_cadence - 1)) == 0)
            {
                constexpr int32_t unroll = refill_cadence < 4 ? refill_cadence : 4; // pixels written out per loop iteration
                do
                {
                    if (unroll == refill_cadence || (n & (refill_cadence - 1)) == 0)
                        dec.top_up();
#pragma unroll
                    for (int32_t g = 0; g < unroll; ++g)
                    {
                        dec.pixel(h, beyond, n - g);
#pragma unroll
                        for (int32_t c = 0; c < NC; ++c)
                            store_shared_sample<S>(sample, g * NC + c, dec.ra[c]);
                    }
                    sample = static_cast<uint32_t>(add_fma(h, static_cast<int32_t>(sample), unroll * NC * static_cast<int32_t>(sizeof(S))));
                    n = add_fma(h, n, -unroll);
                } while (n != 0);
            }
            else
            do
            {
                dec.top_up(); // at the start of every group of refill_cadence pixels
#pragma unroll 1
                do
                {
                    dec.pixel(h, beyond, n);
                    int32_t v[NC];
#pragma unroll
                    for (int32_t c = 0; c < NC; ++c)
                        v[c] = dec.ra[c];
                    if (NC == 3 && transform != 0)
                        color_inverse(transform, sizeof(S) == 2 ? 0xFFFF : 0xFF, v[0], v[NC > 1 ? 1 : 0], v[NC > 2 ? 2 : 0]);
#pragma unroll
                    for (int32_t c = 0; c < NC; ++c)
                        store_shared_sample<S>(sample, c, v[c]);
                    sample = static_cast<uint32_t>(add_fma(h, static_cast<int32_t>(sample), NC * static_cast<int32_t>(sizeof(S))));
                    n = add_fma(h, n, -1);
                } while ((n & (refill_cadence - 1)) != 0); // the group test is the loop condition: no top-up test per pixel
            } while (n != 0);
        }
        __syncwarp();
        tile_store<TW>(tile, pixels, stride, first_line, row_mask, row_bytes, t, lane);
        __syncwarp();
    }
    if (coding)
    {
        const int32_t errc = interval_end_status(p, dec.br, dec.bad(), interval, closing_marker_found, job.offset_table.total != 0);
        if (errc != err_none)
            report_error(job, interval, errc);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// General path: one thread per restart interval, full 2-D LOCO-I (365 contexts in local memory).
// ---------------------------------------------------------------------------------------------------------------------
// Dynamic shared memory: general_context_count contexts (5840 bytes) per thread that has an interval -- 32 in a full
// block (one block per SM then), one for scans without restart markers (one interval per scan, one working lane per warp) --
// and behind them, for samples of up to 12 bits (general_quant_lut_max_maxval), the gradient quantisation table of the block.
constexpr int32_t general_quant_lut_max_maxval = 4095;

__host__ __device__ inline size_t general_working_threads(const CodecParams& p)
{
    return p.interval_count < static_cast<uint32_t>(general_block_threads) ? p.interval_count : general_block_threads;
}

__host__ __device__ inline size_t general_quant_lut_bytes(const CodecParams& p)
{
    return p.maxval <= general_quant_lut_max_maxval ? (static_cast<size_t>(2 * p.maxval + 1) + 15U) / 16U * 16U : 0U;
}

// all threads of the block call this before any of them leaves
__device__ __forceinline__ const int8_t* fill_general_quant_lut(const CodecParams& p, uint8_t* shared)
{
    if (general_quant_lut_bytes(p) == 0)
        return nullptr;
    int8_t* lut = reinterpret_cast<int8_t*>(shared + general_working_threads(p) * general_context_count * sizeof(RegularContext));
    for (int32_t i = threadIdx.x; i <= 2 * p.maxval; i += general_block_threads)
        lut[i] = static_cast<int8_t>(quantize_gradient(p, i - p.maxval));
    __syncthreads();
    return lut;
}

template<bool LOSSLESS>
__global__ void __launch_bounds__(general_block_threads)
    k_encode_general(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs, size_t slot_bytes)
{
    extern __shared__ __align__(16) uint8_t general_shared[];
    const int8_t* quant_lut = fill_general_quant_lut(p, general_shared);
    const ScanJob& job = jobs[blockIdx.y];
    const uint32_t interval = blockIdx.x * general_block_threads + threadIdx.x;
    if (interval >= p.interval_count)
        return;
    RegularContext* contexts = reinterpret_cast<RegularContext*>(general_shared) + threadIdx.x * general_context_count;
    const IntervalResult r = encode_interval_general<LOSSLESS>(p, job, interval, slot_bytes, contexts, quant_lut);
    job.interval_bytes[interval] = r.bytes;
    if (r.errc != err_none)
        report_error(job, interval, r.errc);
}

template<bool LOSSLESS>
__global__ void __launch_bounds__(general_block_threads)
    k_decode_general(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs)
{
    extern __shared__ __align__(16) uint8_t general_shared[];
    const int8_t* quant_lut = fill_general_quant_lut(p, general_shared);
    const ScanJob& job = jobs[blockIdx.y];
    const uint32_t interval = blockIdx.x * general_block_threads + threadIdx.x;
    if (interval >= p.interval_count)
        return;
    RegularContext* contexts = reinterpret_cast<RegularContext*>(general_shared) + threadIdx.x * general_context_count;
    const IntervalResult r = decode_interval_general<LOSSLESS>(p, job, interval, contexts, quant_lut);
    if (r.errc != err_none)
        report_error(job, interval, r.errc);
}

// ---------------------------------------------------------------------------------------------------------------------
// Encode: interval sizes -> offsets -> contiguous stream with RSTm markers
// ---------------------------------------------------------------------------------------------------------------------
constexpr int scan_block_threads = 1024;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t value, uint64_t* warp_totals, uint64_t& block_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inclusive = value;
#pragma unroll
    for (int delta = 1; delta < 32; delta <<= 1)
    {
        const uint64_t other = __shfl_up_sync(0xFFFFFFFFU, inclusive, delta);
        if (lane >= delta)
            inclusive += other;
    }
    if (lane == 31)
        warp_totals[warp] = inclusive;
    __syncthreads();
    if (warp == 0)
    {
        uint64_t t = lane < (blockDim.x >> 5) ? warp_totals[lane] : 0;
        uint64_t inc = t;
#pragma unroll
        for (int delta = 1; delta < 32; delta <<= 1)
        {
            const uint64_t other = __shfl_up_sync(0xFFFFFFFFU, inc, delta);
            if (lane >= delta)
                inc += other;
        }
        warp_totals[lane] = inc - t; // exclusive prefix of the warp totals
        if (lane == 31)
            warp_totals[32] = inc;
    }
    __syncthreads();
    block_total = warp_totals[32];
    const uint64_t result = warp_totals[warp] + inclusive - value;
    __syncthreads();
    return result;
}

// one block per job
__global__ void __launch_bounds__(scan_block_threads)
    k_scan_offsets(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs)
{
    __shared__ uint64_t warp_totals[33];
    const ScanJob& job = jobs[blockIdx.x];
    uint64_t carry = 0;
    for (uint32_t base = 0; base < p.interval_count; base += scan_block_threads)
    {
        const uint32_t i = base + threadIdx.x;
        uint64_t size = 0;
        if (i < p.interval_count)
            size = static_cast<uint64_t>(job.interval_bytes[i]) + (i + 1 < p.interval_count ? 2U : 0U);
        uint64_t total;
        const uint64_t offset = carry + block_exclusive_scan(size, warp_totals, total);
        if (i < p.interval_count)
            job.interval_offset[i] = offset;
        carry += total;
    }
    if (threadIdx.x == 0)
    {
        job.interval_offset[p.interval_count] = carry;
        job.result[0] = carry;
        if (carry > job.stream_out_capacity)
            atomicMin(reinterpret_cast<unsigned long long*>(job.status),
                      (static_cast<unsigned long long>(p.interval_count) << 8) | err_destination_too_small);
    }
}

// one warp per interval: slot -> final position, then the RSTm marker (reference decoder expects FF D0+(n mod 8))
constexpr int gather_block_threads = 256;

__global__ void __launch_bounds__(gather_block_threads)
    k_gather(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs, size_t slot_bytes)
{
    const ScanJob& job = jobs[blockIdx.y];
    const uint32_t interval = blockIdx.x * (gather_block_threads / 32) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (interval >= p.interval_count)
        return;
    const uint64_t total = job.interval_offset[p.interval_count];
    if (total > job.stream_out_capacity)
        return;
    const uint32_t bytes = job.interval_bytes[interval];
    const uint8_t* source = job.slots + static_cast<size_t>(interval) * slot_bytes;
    uint8_t* destination = job.stream_out + job.interval_offset[interval];
    assume_global(source);
    assume_global(destination);

    // head: bytes until the destination is 4-byte aligned; body: aligned 32-bit stores fed by two aligned loads and a
    // funnel shift; tail: bytes
    const uint32_t head = min(bytes, static_cast<uint32_t>((4U - (reinterpret_cast<uintptr_t>(destination) & 3U)) & 3U));
    if (lane < head)
        destination[lane] = source[lane];
    const uint32_t body_words = (bytes - head) / 4U;
    const uint32_t* source_words = reinterpret_cast<const uint32_t*>(source + (head & ~3U)); // slots are 16-byte aligned
    const uint32_t shift = (head & 3U) * 8U;
    uint32_t* destination_words = reinterpret_cast<uint32_t*>(destination + head);
    // four independent word copies per lane and iteration: the loads of one iteration are in flight together
    for (uint32_t w = lane; w < body_words; w += 128)
    {
        uint32_t lo[4], hi[4];
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i)
        {
            const uint32_t index = w + 32U * i;
            lo[i] = index < body_words ? source_words[index] : 0U;
            hi[i] = (shift != 0 && index < body_words) ? source_words[index + 1] : 0U;
        }
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i)
        {
            const uint32_t index = w + 32U * i;
            if (index < body_words)
                destination_words[index] = __funnelshift_r(lo[i], hi[i], shift);
        }
    }
    const uint32_t done = head + body_words * 4U;
    if (lane < bytes - done)
        destination[done + lane] = source[done + lane];
    if (lane == 0 && interval + 1 < p.interval_count)
    {
        destination[bytes] = 0xFF;
        destination[bytes + 1] = static_cast<uint8_t>(0xD0 + (interval & 7U));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Decode front end: ordered table of markers.  A marker is a byte >= 0x80 (and != 0xFF) that follows a 0xFF; inside
// entropy-coded data 0xFF is always followed by a byte < 0x80 (T.87 A.1), so this is unambiguous.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int marker_block_threads = 256;
constexpr int marker_bytes_per_thread = 16;
constexpr int marker_bytes_per_block = marker_block_threads * marker_bytes_per_thread;

// 0x80 in every byte of x that equals 0xFF (exact, no carries between bytes)
__device__ __forceinline__ uint32_t ff_bytes(uint32_t x)
{
    return ((x & 0x7F7F7F7FU) + 0x01010101U) & x & 0x80808080U;
}

// The 16 stream bytes a thread inspects.  Chunks are laid out on absolute 16-byte boundaries of the buffer so that the
// interior ones can be fetched with one aligned 128-bit load; bit i of `mask` <=> data[base + i] is a marker code byte.
struct MarkerChunk
{
    uint32_t mask;
    int64_t base;
};

__device__ __forceinline__ int64_t marker_chunk_base(const uint8_t* data, size_t chunk_index)
{
    const int64_t lead = static_cast<int64_t>(reinterpret_cast<uintptr_t>(data) & 15U);
    return static_cast<int64_t>(chunk_index) * marker_bytes_per_thread - lead;
}

// bit i <=> byte i of the 16 bytes in q is a marker code byte; `previous` = the stream byte in front of them
__device__ __forceinline__ uint32_t marker_mask_of(const uint4& q, uint32_t previous)
{
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        // byte j of `shifted` is the stream byte in front of byte j of w[i] (little endian: earlier byte = lower bits)
        const uint32_t shifted = (w[i] << 8) | previous;
        const uint32_t m = ff_bytes(shifted) & w[i] & ~ff_bytes(w[i]); // 0x80 where FF is followed by >= 0x80, != FF
        mask |= (((m >> 7) & 1U) | ((m >> 14) & 2U) | ((m >> 21) & 4U) | ((m >> 28) & 8U)) << (4 * i);
        previous = w[i] >> 24;
    }
    return mask;
}

__device__ __forceinline__ MarkerChunk marker_chunk(const uint8_t* data, size_t size, size_t chunk_index)
{
    assume_global(data);
    MarkerChunk chunk;
    chunk.base = marker_chunk_base(data, chunk_index);
    chunk.mask = 0;
    const int64_t n = static_cast<int64_t>(size);
    if (chunk.base >= n)
        return chunk;
    if (chunk.base >= 1 && chunk.base + marker_bytes_per_thread <= n)
    {
        const uint4 q = *reinterpret_cast<const uint4*>(data + chunk.base);
        chunk.mask = marker_mask_of(q, data[chunk.base - 1]);
        return chunk;
    }
    const int64_t begin = chunk.base < 0 ? 0 : chunk.base;
    const int64_t end = chunk.base + marker_bytes_per_thread < n ? chunk.base + marker_bytes_per_thread : n;
    uint32_t previous = begin > 0 ? data[begin - 1] : 0U;
    for (int64_t i = begin; i < end; ++i)
    {
        const uint32_t b = data[i];
        if (previous == 0xFFU && b >= 0x80U && b != 0xFFU)
            chunk.mask |= 1U << static_cast<uint32_t>(i - chunk.base);
        previous = b;
    }
    return chunk;
}

// Every block covers marker_sections sections of marker_bytes_per_block stream bytes: the 16-byte loads of all sections
// are issued before any of them is used (four independent loads in flight per thread; with one section per block the
// kernel ran at 2.5 TB/s).  Counts and chunk masks stay per section, so the scan and the layout are those of one section
// per block.
#ifndef JLS_MARKER_SECTIONS
#define JLS_MARKER_SECTIONS 4
#endif
constexpr int marker_sections = JLS_MARKER_SECTIONS;

// grid (ceil(sections / marker_sections), jobs): block_counts[job][section]
__global__ void __launch_bounds__(marker_block_threads)
    k_marker_count(const ScanJob* __restrict__ jobs, uint32_t* __restrict__ block_counts, uint32_t blocks_per_job,
                   uint16_t* __restrict__ chunk_masks)
{
    const ScanJob& job = jobs[blockIdx.y];
    const uint8_t* data = job.stream_in;
    assume_global(data);
    const uint32_t first_section = blockIdx.x * marker_sections;
    uint4 q[marker_sections];
    bool fast[marker_sections];
#pragma unroll
    for (int s = 0; s < marker_sections; ++s)
    {
        const size_t chunk_index = static_cast<size_t>(first_section + s) * marker_block_threads + threadIdx.x;
        const int64_t base = marker_chunk_base(data, chunk_index);
        const bool interior = base >= 1 && base + marker_bytes_per_thread <= static_cast<int64_t>(job.stream_in_size);
        // the whole warp reads 512 consecutive stream bytes: a lane's preceding byte is its neighbour's last one
        fast[s] = __all_sync(0xFFFFFFFFU, interior);
        if (fast[s])
            q[s] = *reinterpret_cast<const uint4*>(data + base);
    }
    __shared__ uint32_t warp_sums[marker_sections][marker_block_threads / 32];
#pragma unroll
    for (int s = 0; s < marker_sections; ++s)
    {
        const uint32_t section = first_section + s;
        if (section >= blocks_per_job)
            break; // whole block
        const size_t chunk_index = static_cast<size_t>(section) * marker_block_threads + threadIdx.x;
        uint32_t mask;
        if (fast[s])
        {
            uint32_t previous = __shfl_up_sync(0xFFFFFFFFU, q[s].w >> 24, 1);
            if ((threadIdx.x & 31) == 0)
                previous = data[marker_chunk_base(data, chunk_index) - 1];
            mask = marker_mask_of(q[s], previous);
        }
        else
        {
            mask = marker_chunk(data, job.stream_in_size, chunk_index).mask;
        }
        // kept for k_marker_write: 2 bytes per 16 stream bytes instead of a second pass over the stream
        chunk_masks[static_cast<size_t>(blockIdx.y) * blocks_per_job * marker_block_threads + chunk_index] = static_cast<uint16_t>(mask);
        uint32_t sum = __popc(mask);
#pragma unroll
        for (int delta = 16; delta > 0; delta >>= 1)
            sum += __shfl_down_sync(0xFFFFFFFFU, sum, delta);
        if ((threadIdx.x & 31) == 0)
            warp_sums[s][threadIdx.x >> 5] = sum;
    }
    __syncthreads();
    if (threadIdx.x < marker_sections && first_section + threadIdx.x < blocks_per_job)
    {
        uint32_t block_sum = 0;
        for (int w = 0; w < marker_block_threads / 32; ++w)
            block_sum += warp_sums[threadIdx.x][w];
        block_counts[static_cast<size_t>(blockIdx.y) * blocks_per_job + first_section + threadIdx.x] = block_sum;
    }
}

// one block per job: exclusive scan of the per-block counts (in place)
__global__ void __launch_bounds__(scan_block_threads)
    k_marker_scan(uint32_t* __restrict__ block_counts, uint32_t blocks_per_job, uint32_t* __restrict__ marker_totals)
{
    __shared__ uint64_t warp_totals[33];
    uint32_t* counts = block_counts + static_cast<size_t>(blockIdx.x) * blocks_per_job;
    uint64_t carry = 0;
    for (uint32_t base = 0; base < blocks_per_job; base += scan_block_threads)
    {
        const uint32_t i = base + threadIdx.x;
        const uint64_t value = i < blocks_per_job ? counts[i] : 0;
        uint64_t total;
        const uint64_t offset = carry + block_exclusive_scan(value, warp_totals, total);
        if (i < blocks_per_job)
            counts[i] = static_cast<uint32_t>(offset);
        carry += total;
    }
    if (threadIdx.x == 0)
        marker_totals[blockIdx.x] = static_cast<uint32_t>(carry);
}

// Writes interval_offset[2i] / [2i+1] (begin / end of interval i) for the first interval_count markers and remembers
// each marker's code in marker_codes[job][i].  Sections per block as in k_marker_count.
__global__ void __launch_bounds__(marker_block_threads)
    k_marker_write(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs,
                   const uint32_t* __restrict__ block_counts, uint32_t blocks_per_job, uint8_t* __restrict__ marker_codes,
                   const uint16_t* __restrict__ chunk_masks)
{
    __shared__ uint32_t warp_sums[marker_sections][marker_block_threads / 32];
    const ScanJob& job = jobs[blockIdx.y];
    assume_global(job.stream_in);
    const uint32_t first_section = blockIdx.x * marker_sections;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t masks[marker_sections], inclusive[marker_sections];
#pragma unroll
    for (int s = 0; s < marker_sections; ++s)
    {
        const size_t chunk_index = static_cast<size_t>(first_section + s) * marker_block_threads + threadIdx.x;
        masks[s] = first_section + s < blocks_per_job
                       ? chunk_masks[static_cast<size_t>(blockIdx.y) * blocks_per_job * marker_block_threads + chunk_index]
                       : 0U;
    }
#pragma unroll
    for (int s = 0; s < marker_sections; ++s)
    {
        inclusive[s] = __popc(masks[s]);
#pragma unroll
        for (int delta = 1; delta < 32; delta <<= 1)
        {
            const uint32_t other = __shfl_up_sync(0xFFFFFFFFU, inclusive[s], delta);
            if (lane >= delta)
                inclusive[s] += other;
        }
        if (lane == 31)
            warp_sums[s][warp] = inclusive[s];
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0)
        job.interval_offset[0] = 0;

#pragma unroll
    for (int s = 0; s < marker_sections; ++s)
    {
        uint32_t remaining = masks[s];
        if (remaining == 0)
            continue;
        const size_t chunk_index = static_cast<size_t>(first_section + s) * marker_block_threads + threadIdx.x;
        const int64_t chunk_base = marker_chunk_base(job.stream_in, chunk_index);
        uint32_t rank = block_counts[static_cast<size_t>(blockIdx.y) * blocks_per_job + first_section + s] + inclusive[s] -
                        __popc(remaining);
        for (int w = 0; w < warp; ++w)
            rank += warp_sums[s][w];
        while (remaining != 0 && rank < p.interval_count)
        {
            const uint32_t bit = __ffs(remaining) - 1;
            remaining &= remaining - 1;
            const size_t code_position = static_cast<size_t>(chunk_base + bit);
            size_t marker_begin = code_position - 1;
            while (marker_begin > 0 && job.stream_in[marker_begin - 1] == 0xFF) // fill bytes (T.81 B.1.1.2)
                --marker_begin;
            job.interval_offset[2 * static_cast<size_t>(rank) + 1] = marker_begin;
            if (rank + 1 < p.interval_count)
                job.interval_offset[2 * static_cast<size_t>(rank) + 2] = code_position + 1;
            const uint8_t code = job.stream_in[code_position];
            marker_codes[static_cast<size_t>(blockIdx.y) * p.interval_count + rank] = code;
            // the first interval_count - 1 markers must be RSTm with m = index mod 8 (reference src/scan_decoder.hpp:335-349)
            if (rank + 1 < p.interval_count && code != 0xD0U + (rank & 7U))
                report_error(job, rank, err_restart_marker_not_found);
            ++rank;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Side table of interval offsets (jls_common.h)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t* offset_table_entry(const OffsetTableRef& table, uint32_t index)
{
    return table.entries[index / offset_table_entries_per_segment] + static_cast<size_t>(index % offset_table_entries_per_segment) * 4U;
}

// Encode: interval_offset[0 .. N] (k_scan_offsets) -> big-endian entries in the reserved segments of the stream header
__global__ void k_write_offset_table(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs)
{
    const ScanJob& job = jobs[blockIdx.y];
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (job.offset_table.total != p.interval_count + 1U || j > p.interval_count)
        return;
    const uint64_t total = job.interval_offset[p.interval_count];
    if (total > job.stream_out_capacity || total >= (1ULL << 32))
        return; // nothing was written / does not fit 32-bit entries: the entries stay zero and no decoder will believe them
    const uint32_t v = static_cast<uint32_t>(job.interval_offset[j]);
    uint8_t* entry = offset_table_entry(job.offset_table, j);
    entry[0] = static_cast<uint8_t>(v >> 24);
    entry[1] = static_cast<uint8_t>(v >> 16);
    entry[2] = static_cast<uint8_t>(v >> 8);
    entry[3] = static_cast<uint8_t>(v);
}

__device__ __forceinline__ uint32_t read_offset_entry(const OffsetTableRef& table, uint32_t index)
{
    const uint8_t* e = offset_table_entry(table, index);
    return (static_cast<uint32_t>(e[0]) << 24) | (static_cast<uint32_t>(e[1]) << 16) | (static_cast<uint32_t>(e[2]) << 8) | e[3];
}

// Decode: the table instead of the three marker kernels.  Thread i checks what the table says about interval i against the
// stream -- offsets ascending and inside the stream, FF D0+(i mod 8) directly in front of the next interval, a marker
// where the scan is said to end -- and fills interval_offset / marker_codes / marker_totals exactly as k_marker_write
// would have.  Any disagreement rejects the table for this job (errc_offset_table_rejected): the engine then decodes the
// stream again by searching for the markers.  What the table cannot know -- a marker INSIDE an interval -- the decoders
// notice themselves (a 0xFF followed by a byte >= 0x80 is an error to their bit readers; interval_end_status, strict).
__global__ void k_offsets_from_table(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs,
                                     uint32_t* __restrict__ marker_totals, uint8_t* __restrict__ marker_codes)
{
    const ScanJob& job = jobs[blockIdx.y];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = p.interval_count;
    if (i >= n)
        return;
    const uint8_t* data = job.stream_in;
    assume_global(data);
    const uint64_t begin = read_offset_entry(job.offset_table, i);
    const uint64_t next = read_offset_entry(job.offset_table, i + 1);
    const bool last = i + 1 == n;
    bool ok = (i != 0 || begin == 0) && next >= begin + (last ? 0U : 2U) && next + (last ? 2U : 0U) <= job.stream_in_size;
    uint64_t end = next;
    uint8_t code = 0;
    if (ok && last)
    {
        code = data[next + 1];
        ok = data[next] == 0xFF && code >= 0x80 && code != 0xFF;
    }
    else if (ok)
    {
        end = next - 2;
        code = data[next - 1];
        ok = data[end] == 0xFF && code == 0xD0U + (i & 7U);
        // fill bytes in front of a marker (T.81 B.1.1.2) belong to the marker: the search would have ended the interval there
        ok = ok && (end == begin || data[end - 1] != 0xFF);
    }
    if (!ok)
    {
        report_error(job, 0, errc_offset_table_rejected);
        return;
    }
    job.interval_offset[2 * static_cast<size_t>(i)] = begin;
    job.interval_offset[2 * static_cast<size_t>(i) + 1] = end;
    marker_codes[static_cast<size_t>(blockIdx.y) * n + i] = code;
    if (i == 0)
        marker_totals[blockIdx.y] = n;
}

// one thread per job: bytes consumed by the scan, its closing marker, and the data-ends-early case
__global__ void k_decode_finish(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs,
                                const uint32_t* __restrict__ marker_totals, const uint8_t* __restrict__ marker_codes,
                                uint32_t job_count)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= job_count)
        return;
    const ScanJob& job = jobs[j];
    const uint32_t found = marker_totals[j];
    const uint8_t* codes = marker_codes + static_cast<size_t>(j) * p.interval_count;
    if (found < p.interval_count)
    {
        // The data ends inside interval `found`.  The reference either ran out of bits while decoding it
        // (invalid_data, already reported by the decode kernel) or finds nothing where the marker should be.
        if ((*job.status >> 8) != found)
            report_error(job, found, err_need_more_data);
        job.result[0] = job.stream_in_size;
    }
    else
    {
        job.result[0] = job.interval_offset[2 * static_cast<size_t>(p.interval_count - 1) + 1];
        job.result[1] = codes[p.interval_count - 1]; // the marker that closes the scan (EOI, SOS, DNL, ...)
    }
}

__global__ void k_init_decode_tables(const __grid_constant__ CodecParams p, const ScanJob* __restrict__ jobs)
{
    const ScanJob& job = jobs[blockIdx.y];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * p.interval_count)
        job.interval_offset[i] = ~0ULL;
}

__global__ void k_init_status(const ScanJob* __restrict__ jobs, uint32_t job_count)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < job_count)
    {
        *reinterpret_cast<unsigned long long*>(jobs[j].status) = ~0ULL;
        jobs[j].result[0] = 0;
        jobs[j].result[1] = 0;
    }
}

// Batch encode: completes every frame's stream: header bytes in front of the entropy-coded data, EOI behind it.
// One warp per frame.  stream_out points at the first entropy byte, i.e. header_size bytes into the frame's stream.
__global__ void k_wrap_frames(const ScanJob* __restrict__ jobs, const uint8_t* __restrict__ header, uint32_t header_size,
                              uint32_t job_count, int parts)
{
    const uint32_t j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (j >= job_count)
        return;
    const ScanJob& job = jobs[j];
    if ((parts & wrap_header) != 0 && job.stream_out_capacity != 0) // capacity 0: not even header + EOI fit
    {
        uint8_t* begin = job.stream_out - header_size;
        for (uint32_t i = lane; i < header_size; i += 32)
            begin[i] = header[i];
    }
    if ((parts & wrap_end_of_image) != 0 && job.result[0] <= job.stream_out_capacity && lane == 0)
    {
        job.stream_out[job.result[0]] = 0xFF;
        job.stream_out[job.result[0] + 1] = 0xD9;
    }
}

// Batch decode: gathers the first `prefix_bytes` of every stream so that the host can parse the headers with one copy.
__global__ void k_copy_prefixes(const uint8_t* const* __restrict__ streams, const size_t* __restrict__ sizes,
                                uint8_t* __restrict__ prefixes, uint32_t prefix_bytes, uint32_t job_count)
{
    const uint32_t j = blockIdx.x;
    if (j >= job_count)
        return;
    const size_t n = sizes[j] < prefix_bytes ? sizes[j] : prefix_bytes;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x)
        prefixes[static_cast<size_t>(j) * prefix_bytes + i] = streams[j][i];
}

template<typename Kernel, typename... Args>
cudaError_t launch(Kernel kernel, dim3 grid, dim3 block, cudaStream_t stream, Args... args)
{
    kernel<<<grid, block, 0, stream>>>(args...);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    ++t_kernel_launches;
    return cudaGetLastError();
}

// CHARLS_B200_TRACE_OCCUPANCY=1 prints, once per kernel and shared-memory size, the resident blocks per SM the runtime
// computes for the tile kernels.  (A preferred shared-memory carve-out changes nothing: the driver sizes it for the
// resident blocks already.)
void trace_occupancy(const void* kernel, int block_threads, size_t dynamic_shared_bytes)
{
    static const char* const trace = std::getenv("CHARLS_B200_TRACE_OCCUPANCY");
    if (trace == nullptr)
        return;
    static std::mutex mutex;
    static std::set<std::pair<const void*, size_t>> seen;
    const std::lock_guard<std::mutex> lock(mutex);
    if (!seen.insert({kernel, dynamic_shared_bytes}).second)
        return;
    int blocks = 0;
    cudaFuncAttributes attributes{};
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, block_threads, dynamic_shared_bytes);
    cudaFuncGetAttributes(&attributes, kernel);
    std::fprintf(stderr, "charls_b200: kernel %p: %d registers, %zu + %zu bytes shared, %d resident blocks per SM\n", kernel,
                 attributes.numRegs, attributes.sharedSizeBytes, dynamic_shared_bytes, blocks);
}

template<typename Kernel, typename... Args>
cudaError_t launch_with_shared(Kernel kernel, dim3 grid, dim3 block, size_t dynamic_shared_bytes, cudaStream_t stream, Args... args)
{
    trace_occupancy(reinterpret_cast<const void*>(kernel), static_cast<int>(block.x), dynamic_shared_bytes);
    kernel<<<grid, block, dynamic_shared_bytes, stream>>>(args...);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    ++t_kernel_launches;
    return cudaGetLastError();
}

// General-path kernels: shared memory for the contexts of the threads of a block that have an interval (up to 32 x 5840
// bytes, which needs the opt-in limit).
template<typename Kernel, typename... Args>
cudaError_t launch_general(Kernel kernel, dim3 grid, const CodecParams& p, cudaStream_t stream, Args... args)
{
    const size_t bytes = general_working_threads(p) * general_context_count * sizeof(RegularContext) + general_quant_lut_bytes(p);
    if (bytes > 48U * 1024U)
    {
        const cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(bytes));
        if (e != cudaSuccess)
            return e;
    }
    return launch_with_shared(kernel, grid, dim3(general_block_threads), bytes, stream, args...);
}

// The tile kernels keep the context-index table |Q(-Ra)| in dynamic shared memory: Ra = 0 .. 255 for 8-bit containers (no
// clamp in the pixel loop), Ra = 0 .. min(T3, capacity - 1) otherwise (277 bytes for the 12..16 bit defaults).  A fixed 1 KB table per one-warp block costs three resident blocks per SM.
size_t tiled_dynamic_shared_bytes(const CodecParams& p)
{
    const size_t entries =
        p.sample_bytes == 1 ? 256U : static_cast<size_t>(p.t3 < context_lut_capacity - 1 ? p.t3 : context_lut_capacity - 1) + 1;
    return (entries + 15) / 16 * 16;
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant selection for the fast path -- the counterpart of the reference's make_scan_codec (src/make_scan_codec.cpp:40-156):
// components per pixel (1 = scalar lines, 2..4 = sample interleave), lossless or not, 8 / 16 bit containers,
// tiled (4-byte aligned rows) or per-lane access, line interleave.
// ---------------------------------------------------------------------------------------------------------------------
bool rows_tileable(const CodecParams& p, bool rows_word_aligned)
{
    // rows need not END on a word boundary: the last tile of a row is copied word by word with the row's length in hand
    // (tile_load_async, tile_store), and a stride that is a multiple of four leaves room behind such a row
    return rows_word_aligned && p.interleave != ilv_line && p.t3 < context_lut_capacity && p.reset < reciprocal_lut_capacity;
}

template<int NC, bool LL, typename S>
cudaError_t launch_encode_fast(const CodecParams& p, bool tiled, dim3 grid, dim3 block, cudaStream_t stream,
                               const ScanJob* jobs, size_t slot_bytes)
{
    if (tiled)
    {
        if constexpr (LL)
        {
            if (p.bits_per_sample == static_cast<int32_t>(8 * sizeof(S)))
                return launch_with_shared(k_encode_tiled<NC, LL, S, 8 * sizeof(S)>, grid, block, tiled_dynamic_shared_bytes(p), stream,
                                          p, jobs, slot_bytes);
        }
        return launch_with_shared(k_encode_tiled<NC, LL, S, 0>, grid, block, tiled_dynamic_shared_bytes(p), stream, p, jobs, slot_bytes);
    }
    if constexpr (NC == 1)
    {
        if (p.interleave == ilv_line)
            return launch(k_encode_fast<1, LL, S, true>, grid, block, stream, p, jobs, slot_bytes);
    }
    return launch(k_encode_fast<NC, LL, S, false>, grid, block, stream, p, jobs, slot_bytes);
}

template<int NC, bool LL, typename S>
cudaError_t launch_decode_fast(const CodecParams& p, bool tiled, dim3 grid, dim3 block, cudaStream_t stream, const ScanJob* jobs)
{
    if (tiled)
    {
        if constexpr (LL)
        {
            if (p.bits_per_sample == static_cast<int32_t>(8 * sizeof(S)))
                return launch_with_shared(k_decode_tiled<NC, LL, S, 8 * sizeof(S)>, grid, block, tiled_dynamic_shared_bytes(p), stream,
                                          p, jobs);
        }
        return launch_with_shared(k_decode_tiled<NC, LL, S, 0>, grid, block, tiled_dynamic_shared_bytes(p), stream, p, jobs);
    }
    if constexpr (NC == 1)
    {
        if (p.interleave == ilv_line)
            return launch(k_decode_fast<1, LL, S, true>, grid, block, stream, p, jobs);
    }
    return launch(k_decode_fast<NC, LL, S, false>, grid, block, stream, p, jobs);
}

template<int NC>
cudaError_t dispatch_encode_nc(const CodecParams& p, bool tiled, dim3 grid, dim3 block, cudaStream_t stream,
                               const ScanJob* jobs, size_t slot_bytes)
{
    const bool wide = p.sample_bytes == 2;
    if (p.near == 0)
        return wide ? launch_encode_fast<NC, true, uint16_t>(p, tiled, grid, block, stream, jobs, slot_bytes)
                    : launch_encode_fast<NC, true, uint8_t>(p, tiled, grid, block, stream, jobs, slot_bytes);
    return wide ? launch_encode_fast<NC, false, uint16_t>(p, tiled, grid, block, stream, jobs, slot_bytes)
                : launch_encode_fast<NC, false, uint8_t>(p, tiled, grid, block, stream, jobs, slot_bytes);
}

template<int NC>
cudaError_t dispatch_decode_nc(const CodecParams& p, bool tiled, dim3 grid, dim3 block, cudaStream_t stream, const ScanJob* jobs)
{
    const bool wide = p.sample_bytes == 2;
    if (p.near == 0)
        return wide ? launch_decode_fast<NC, true, uint16_t>(p, tiled, grid, block, stream, jobs)
                    : launch_decode_fast<NC, true, uint8_t>(p, tiled, grid, block, stream, jobs);
    return wide ? launch_decode_fast<NC, false, uint16_t>(p, tiled, grid, block, stream, jobs)
                : launch_decode_fast<NC, false, uint8_t>(p, tiled, grid, block, stream, jobs);
}

cudaError_t dispatch_encode_fast(const CodecParams& p, bool rows_word_aligned, dim3 grid, dim3 block, cudaStream_t stream,
                                 const ScanJob* jobs, size_t slot_bytes)
{
    const bool tiled = rows_tileable(p, rows_word_aligned);
    switch (p.interleave == ilv_sample ? p.components : 1)
    {
    case 3:
        return dispatch_encode_nc<3>(p, tiled, grid, block, stream, jobs, slot_bytes);
#if !defined(JLS_DEV_SUBSET) // development builds (tools/dev_sass.sh) instantiate the one- and three-component kernels only
    case 2:
        return dispatch_encode_nc<2>(p, tiled, grid, block, stream, jobs, slot_bytes);
    case 4:
        return dispatch_encode_nc<4>(p, tiled, grid, block, stream, jobs, slot_bytes);
#endif
    default:
        return dispatch_encode_nc<1>(p, tiled, grid, block, stream, jobs, slot_bytes);
    }
}

cudaError_t dispatch_decode_fast(const CodecParams& p, bool rows_word_aligned, dim3 grid, dim3 block, cudaStream_t stream,
                                 const ScanJob* jobs)
{
    const bool tiled = rows_tileable(p, rows_word_aligned);
    switch (p.interleave == ilv_sample ? p.components : 1)
    {
    case 3:
        return dispatch_decode_nc<3>(p, tiled, grid, block, stream, jobs);
#if !defined(JLS_DEV_SUBSET) // development builds (tools/dev_sass.sh) instantiate the one- and three-component kernels only
    case 2:
        return dispatch_decode_nc<2>(p, tiled, grid, block, stream, jobs);
    case 4:
        return dispatch_decode_nc<4>(p, tiled, grid, block, stream, jobs);
#endif
    default:
        return dispatch_decode_nc<1>(p, tiled, grid, block, stream, jobs);
    }
}

} // namespace

uint64_t kernel_launch_count() noexcept
{
    return g_kernel_launches.load(std::memory_order_relaxed);
}

uint64_t thread_kernel_launch_count() noexcept
{
    return t_kernel_launches;
}

void count_kernel_launches(uint32_t launches) noexcept
{
    g_kernel_launches.fetch_add(launches, std::memory_order_relaxed);
    t_kernel_launches += launches;
}

size_t marker_blocks_for(size_t stream_bytes) noexcept
{
    // + 15: the chunk grid starts at the 16-byte boundary at or below the first stream byte
    return (stream_bytes + 15 + marker_bytes_per_block - 1) / marker_bytes_per_block;
}

// Layout of the marker scratch: [jobs][blocks] uint32 block counts (rounded up to 16 bytes), then [jobs][blocks][threads]
// uint16 chunk masks.
static size_t marker_counts_bytes(size_t job_count, size_t blocks_per_job) noexcept
{
    return (job_count * blocks_per_job * sizeof(uint32_t) + 15) / 16 * 16;
}

size_t marker_scratch_bytes(size_t job_count, size_t max_stream_bytes) noexcept
{
    const size_t blocks = marker_blocks_for(max_stream_bytes);
    return marker_counts_bytes(job_count, blocks) + job_count * blocks * marker_block_threads * sizeof(uint16_t);
}

#define JLS_TRY(expr)                                                                                                  \
    do                                                                                                                 \
    {                                                                                                                  \
        const cudaError_t jls_try_error = (expr);                                                                      \
        if (jls_try_error != cudaSuccess)                                                                              \
            return jls_try_error;                                                                                      \
    } while (0)

cudaError_t launch_encode(const CodecParams& p, const ScanJob* device_jobs, uint32_t job_count, size_t slot_bytes,
                          cudaStream_t stream, cudaEvent_t* coder_events, bool rows_word_aligned, bool offset_tables)
{
    JLS_TRY(launch(k_init_status, dim3((job_count + 127) / 128), dim3(128), stream, device_jobs, job_count));
    const bool lossless = p.near == 0;
    if (coder_events)
        JLS_TRY(cudaEventRecord(coder_events[0], stream));
    if (use_fast_path(p))
    {
        const dim3 grid((p.interval_count + fast_block_threads - 1) / fast_block_threads, job_count);
        const dim3 block(fast_block_threads);
        JLS_TRY(dispatch_encode_fast(p, rows_word_aligned, grid, block, stream, device_jobs, slot_bytes));
    }
    else
    {
        const dim3 grid((p.interval_count + general_block_threads - 1) / general_block_threads, job_count);
        if (lossless)
            JLS_TRY(launch_general(k_encode_general<true>, grid, p, stream, p, device_jobs, slot_bytes));
        else
            JLS_TRY(launch_general(k_encode_general<false>, grid, p, stream, p, device_jobs, slot_bytes));
    }
    if (coder_events)
        JLS_TRY(cudaEventRecord(coder_events[1], stream));
    JLS_TRY(launch(k_scan_offsets, dim3(job_count), dim3(scan_block_threads), stream, p, device_jobs));
    if (offset_tables)
        JLS_TRY(launch(k_write_offset_table, dim3((p.interval_count + 1 + 255) / 256, job_count), dim3(256), stream, p, device_jobs));
    const dim3 gather_grid((p.interval_count + gather_block_threads / 32 - 1) / (gather_block_threads / 32), job_count);
    JLS_TRY(launch(k_gather, gather_grid, dim3(gather_block_threads), stream, p, device_jobs, slot_bytes));
    return cudaSuccess;
}

cudaError_t launch_decode(const CodecParams& p, const ScanJob* device_jobs, uint32_t job_count, size_t max_stream_bytes,
                          uint32_t* block_counts, uint32_t* marker_totals, uint8_t* marker_codes, cudaStream_t stream,
                          cudaEvent_t* coder_events, bool rows_word_aligned, bool offset_tables)
{
    const uint32_t blocks_per_job = static_cast<uint32_t>(marker_blocks_for(max_stream_bytes));
    JLS_TRY(launch(k_init_status, dim3((job_count + 127) / 128), dim3(128), stream, device_jobs, job_count));
    JLS_TRY(launch(k_init_decode_tables, dim3((2 * p.interval_count + 255) / 256, job_count), dim3(256), stream, p,
                   device_jobs));
    if (offset_tables)
    {
        // every job carries a side table of interval offsets: it is checked against the stream and used instead of the search
        JLS_TRY(launch(k_offsets_from_table, dim3((p.interval_count + 255) / 256, job_count), dim3(256), stream, p, device_jobs,
                       marker_totals, marker_codes));
    }
    else
    {
        uint16_t* chunk_masks =
            reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(block_counts) + marker_counts_bytes(job_count, blocks_per_job));
        const uint32_t marker_grid = (blocks_per_job + marker_sections - 1) / marker_sections;
        JLS_TRY(launch(k_marker_count, dim3(marker_grid, job_count), dim3(marker_block_threads), stream, device_jobs,
                       block_counts, blocks_per_job, chunk_masks));
        JLS_TRY(launch(k_marker_scan, dim3(job_count), dim3(scan_block_threads), stream, block_counts, blocks_per_job,
                       marker_totals));
        JLS_TRY(launch(k_marker_write, dim3(marker_grid, job_count), dim3(marker_block_threads), stream, p, device_jobs,
                       static_cast<const uint32_t*>(block_counts), blocks_per_job, marker_codes,
                       static_cast<const uint16_t*>(chunk_masks)));
    }

    const bool lossless = p.near == 0;
    if (coder_events)
        JLS_TRY(cudaEventRecord(coder_events[0], stream));
    if (use_fast_path(p))
    {
        const dim3 grid((p.interval_count + fast_block_threads - 1) / fast_block_threads, job_count);
        const dim3 block(fast_block_threads);
        JLS_TRY(dispatch_decode_fast(p, rows_word_aligned, grid, block, stream, device_jobs));
    }
    else
    {
        const dim3 grid((p.interval_count + general_block_threads - 1) / general_block_threads, job_count);
        if (lossless)
            JLS_TRY(launch_general(k_decode_general<true>, grid, p, stream, p, device_jobs));
        else
            JLS_TRY(launch_general(k_decode_general<false>, grid, p, stream, p, device_jobs));
    }
    if (coder_events)
        JLS_TRY(cudaEventRecord(coder_events[1], stream));
    JLS_TRY(launch(k_decode_finish, dim3((job_count + 127) / 128), dim3(128), stream, p, device_jobs,
                   static_cast<const uint32_t*>(marker_totals), static_cast<const uint8_t*>(marker_codes), job_count));
    return cudaSuccess;
}

cudaError_t launch_wrap_frames(const ScanJob* device_jobs, const uint8_t* device_header, uint32_t header_size,
                               uint32_t job_count, cudaStream_t stream, int parts)
{
    return launch(k_wrap_frames, dim3((job_count + 3) / 4), dim3(128), stream, device_jobs, device_header, header_size,
                  job_count, parts);
}

cudaError_t launch_copy_prefixes(const uint8_t* const* device_streams, const size_t* device_sizes, uint8_t* device_prefixes,
                                 uint32_t prefix_bytes, uint32_t job_count, cudaStream_t stream)
{
    return launch(k_copy_prefixes, dim3(job_count), dim3(128), stream, device_streams, device_sizes, device_prefixes,
                  prefix_bytes, job_count);
}

} // namespace jls
