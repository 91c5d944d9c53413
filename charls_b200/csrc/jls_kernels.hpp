// jls_kernels.hpp -- launch wrappers of the sm_100a kernels (implemented in jls_kernels.cu).
#pragma once

#include "jls_common.h"

#include <atomic>
#include <cuda_runtime_api.h>

namespace jls {

// Number of kernels this library has launched in this process (evidence for bench.py's "gpu_launches").
uint64_t kernel_launch_count() noexcept;        // process-wide
uint64_t thread_kernel_launch_count() noexcept; // issued by the calling host thread
// kernels that ran without going through launch(): the replay of a captured graph
void count_kernel_launches(uint32_t launches) noexcept;

// Blocks (of 4096 stream bytes) the marker kernels use for a stream of `stream_bytes`.
size_t marker_blocks_for(size_t stream_bytes) noexcept;
size_t marker_scratch_bytes(size_t job_count, size_t max_stream_bytes) noexcept;

// Encodes `job_count` scans that share the coding parameters `p`.  Afterwards, per job: result[0] = bytes written to
// stream_out (interval data + RSTm markers), status = first error key (~0 when none).
// coder_events (optional): two events recorded directly before / after the entropy-coding kernel (the dominant one).
// rows_word_aligned: every job's sample buffer and stride are multiples of 4 bytes (enables the shared-memory tile kernels).
// offset_tables: the jobs' ScanJob::offset_table entries are filled in (side table of interval offsets, jls_common.h).
cudaError_t launch_encode(const CodecParams& p, const ScanJob* device_jobs, uint32_t job_count, size_t slot_bytes,
                          cudaStream_t stream, cudaEvent_t* coder_events = nullptr, bool rows_word_aligned = false,
                          bool offset_tables = false);

// Decodes `job_count` scans.  block_counts: marker_scratch_bytes(job_count, max_stream_bytes) bytes of scratch (block
// counts, then the per-chunk marker masks); marker_totals: job_count uint32; marker_codes: job_count * interval_count
// bytes.  Afterwards result[0] = bytes consumed by the scan.
cudaError_t launch_decode(const CodecParams& p, const ScanJob* device_jobs, uint32_t job_count, size_t max_stream_bytes,
                          uint32_t* block_counts, uint32_t* marker_totals, uint8_t* marker_codes, cudaStream_t stream,
                          cudaEvent_t* coder_events = nullptr, bool rows_word_aligned = false, bool offset_tables = false);
// offset_tables (decode): EVERY job's ScanJob::offset_table is set; the intervals come from the tables (checked against the
// streams) instead of the marker search.  A job whose status is not success afterwards must be decoded again without.

// Batch encode: writes `header` in front of and EOI behind every frame's entropy-coded data (see k_wrap_frames).
// parts: the header (before the scan is coded when it reserves a side table whose entries the encoder fills in) and / or EOI
constexpr int wrap_header = 1, wrap_end_of_image = 2;
cudaError_t launch_wrap_frames(const ScanJob* device_jobs, const uint8_t* device_header, uint32_t header_size,
                               uint32_t job_count, cudaStream_t stream, int parts = wrap_header | wrap_end_of_image);

// Batch decode: prefixes[j * prefix_bytes ...] = first bytes of stream j.
cudaError_t launch_copy_prefixes(const uint8_t* const* device_streams, const size_t* device_sizes, uint8_t* device_prefixes,
                                 uint32_t prefix_bytes, uint32_t job_count, cudaStream_t stream);

} // namespace jls
