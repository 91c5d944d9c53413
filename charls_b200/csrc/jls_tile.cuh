// jls_tile.cuh -- warp-cooperative staging of sample rows through shared memory (device only).
//
// A warp codes 32 lines in lock step, every lane walking along its own line.  Touching global memory per lane would cost
// 32 L1 wavefronts per instruction (32 lanes, 32 different lines) -- ncu showed the L1 pipe saturating before the issue
// slots (profiles/r1_notes.md).  So the warp moves [32 lines x TW words] tiles instead: every copy instruction covers
// whole 64/96-byte row segments (coalesced), the tile sits in shared memory with rows padded to an odd number of
// words, and lane r then reads / writes its row with conflict-free LDS/STS (bank = (r * (TW+1) + w) mod 32 is a
// permutation of the lanes).  Loads are asynchronous (cp.async, SASS LDGSTS) and double buffered: the tile after the
// one being coded is already in flight, so HBM latency is hidden behind 64 pixels of coding work per lane.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace jls {

__device__ __forceinline__ void cp_async_4(void* smem_destination, const void* global_source, int source_bytes)
{
    const unsigned address = static_cast<unsigned>(__cvta_generic_to_shared(smem_destination));
    // source_bytes == 0: nothing is read, the destination word is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(address), "l"(global_source), "r"(source_bytes)
                 : "memory");
}

__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

template<int PENDING>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING) : "memory");
}

// Where lane `lane` starts in a [32 rows][TW words] tile when the warp walks it 32 words at a time (word k * 32 + lane of
// the tile in the k-th step), and how it moves from one step to the next: TW = 16: same word, two rows down; TW = 24:
// eight words to the right and one row down, wrapping into one more row.
template<int TW>
struct TileWalk
{
    static_assert(TW == 16 || TW == 24, "tiles are 64 or 96 bytes wide");
    uint32_t r, w;
    __device__ __forceinline__ explicit TileWalk(uint32_t lane) : r(lane / TW), w(lane % TW) {}
    // advances to the next step; returns the distance in words inside the padded shared-memory tile and reports the
    // distance in global memory as rows (`rows`) and words (`words`, may be negative)
    __device__ __forceinline__ void step(int32_t& rows, int32_t& words)
    {
        if constexpr (TW == 16)
        {
            rows = 2;
            words = 0;
        }
        else
        {
            const bool wrap = w + 8U >= TW;
            rows = wrap ? 2 : 1;
            words = wrap ? 8 - TW : 8;
            w = wrap ? w + 8U - TW : w + 8U;
        }
    }
};

// Starts the copy of tile `tile_index` (TW words of each of the warp's 32 lines) into `tile` ([32][TW + 1] words).
// Lines past `last_line` repeat the last line (never coded); words past the end of a row are zero-filled.
template<int TW>
__device__ __forceinline__ void tile_load_async(uint32_t* tile, const uint8_t* pixels, size_t stride, uint32_t first_line,
                                                uint32_t last_line, int32_t row_bytes, int32_t tile_index, uint32_t lane)
{
    if (first_line + 31U <= last_line && (tile_index + 1) * (TW * 4) <= row_bytes)
    {
        // A whole tile (all 32 lines exist, the row does not end inside it): one pointer per lane that is advanced from
        // step to step, no bounds to test.  The general form below costs ~25 instructions per step for index arithmetic
        // (profiles/r1_notes.md), this one 4.
        TileWalk<TW> walk(lane);
        const uint8_t* source = pixels + static_cast<size_t>(first_line + walk.r) * stride + tile_index * (TW * 4) + walk.w * 4U;
        unsigned destination = static_cast<unsigned>(__cvta_generic_to_shared(tile + walk.r * (TW + 1) + walk.w));
#pragma unroll
        for (int k = 0; k < TW; ++k)
        {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(destination), "l"(source) : "memory");
            int32_t rows, words;
            walk.step(rows, words);
            source += static_cast<size_t>(rows) * stride + words * 4;
            destination += static_cast<unsigned>(rows * (TW + 1) + words) * 4U;
        }
        cp_async_commit();
        return;
    }
#pragma unroll 4
    for (int k = 0; k < TW; ++k)
    {
        const uint32_t index = static_cast<uint32_t>(k) * 32U + lane;
        const uint32_t r = index / TW;
        const uint32_t w = index - r * TW;
        const int32_t byte = tile_index * (TW * 4) + static_cast<int32_t>(w) * 4;
        const bool inside = byte < row_bytes;
        const uint32_t line = min(first_line + r, last_line);
        const uint8_t* source = pixels + static_cast<size_t>(line) * stride + (inside ? byte : 0);
        cp_async_4(tile + r * (TW + 1) + w, source, inside ? 4 : 0);
    }
    cp_async_commit();
}

// Writes tile `tile_index` back: row r goes to line first_line + r if bit r of `row_mask` is set.
template<int TW>
__device__ __forceinline__ void tile_store(const uint32_t* tile, uint8_t* pixels, size_t stride, uint32_t first_line,
                                           uint32_t row_mask, int32_t row_bytes, int32_t tile_index, uint32_t lane)
{
    if (row_mask == 0xFFFFFFFFU && (tile_index + 1) * (TW * 4) <= row_bytes)
    {
        // a whole tile: see tile_load_async
        TileWalk<TW> walk(lane);
        uint8_t* destination = pixels + static_cast<size_t>(first_line + walk.r) * stride + tile_index * (TW * 4) + walk.w * 4U;
        unsigned source = static_cast<unsigned>(__cvta_generic_to_shared(tile + walk.r * (TW + 1) + walk.w));
#pragma unroll
        for (int k = 0; k < TW; ++k)
        {
            uint32_t word;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(source) : "memory");
            *reinterpret_cast<uint32_t*>(destination) = word;
            int32_t rows, words;
            walk.step(rows, words);
            destination += static_cast<size_t>(rows) * stride + words * 4;
            source += static_cast<unsigned>(rows * (TW + 1) + words) * 4U;
        }
        return;
    }
#pragma unroll 4
    for (int k = 0; k < TW; ++k)
    {
        const uint32_t index = static_cast<uint32_t>(k) * 32U + lane;
        const uint32_t r = index / TW;
        const uint32_t w = index - r * TW;
        const int32_t byte = tile_index * (TW * 4) + static_cast<int32_t>(w) * 4;
        if (byte < row_bytes && ((row_mask >> r) & 1U) != 0)
            *reinterpret_cast<uint32_t*>(pixels + static_cast<size_t>(first_line + r) * stride + byte) = tile[r * (TW + 1) + w];
    }
}

} // namespace jls
