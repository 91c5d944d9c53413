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

// How the warp walks a whole [32 rows][TW words] tile, 32 words per copy instruction (word k * 32 + lane of the tile in
// step k).  The walk repeats after `period` steps, `rows` rows further down: TW = 16: every step covers two rows;
// TW = 24: three steps cover four rows.  Per lane that is `period` offsets inside the group (kept in registers) and one
// base that advances by `rows` rows.  Written as 24 individual steps, each with its own distance to the next, ptxas hoisted
// all 24 distances out of the tile loop: 120-128 registers for three-component pixels against 64-72 for the others, and
// 16 resident warps per SM instead of 28.
template<int TW, int SW = TW + 1>
struct TileWalk
{
    static_assert(TW == 8 || TW == 12 || TW == 16 || TW == 24, "tiles are 32, 48, 64 or 96 bytes wide");
    // 32 words per step: four 32-byte rows or two 64-byte rows; 48- and 96-byte rows repeat after three steps (96 words =
    // eight / four rows)
    static constexpr int period = (TW == 24 || TW == 12) ? 3 : 1;
    static constexpr int rows = TW == 16 ? 2 : TW == 12 ? 8 : 4;
    static constexpr int groups = 32 / rows;
    uint32_t global_offset[period]; // bytes from the group's first row in global memory
    uint32_t shared_offset[period]; // bytes from the group's first row in the padded shared-memory tile
    __device__ __forceinline__ TileWalk(uint32_t lane, uint32_t stride)
    {
#pragma unroll
        for (int j = 0; j < period; ++j)
        {
            const uint32_t index = static_cast<uint32_t>(j) * 32U + lane;
            const uint32_t r = index / TW, w = index % TW;
            global_offset[j] = r * stride + w * 4U;
            shared_offset[j] = (r * SW + w) * 4U;
        }
    }
};

// Starts the copy of tile `tile_index` (TW words of each of the warp's 32 lines) into `tile` ([32][TW + 1] words).
// Lines past `last_line` repeat the last line (never coded); words past the end of a row are zero-filled.
template<int TW, int SW = TW + 1>
__device__ __forceinline__ void tile_load_async(uint32_t* tile, const uint8_t* pixels, size_t stride, uint32_t first_line,
                                                uint32_t last_line, int32_t row_bytes, int32_t tile_index, uint32_t lane)
{
    if (first_line + 31U <= last_line && (tile_index + 1) * (TW * 4) <= row_bytes)
    {
        // A whole tile (all 32 lines exist, the row does not end inside it): a base per group of rows and the lane's
        // offsets inside a group, no bounds to test.  The general form below costs ~25 instructions per step for index
        // arithmetic (profiles/r1_notes.md), this one 4.
        const TileWalk<TW, SW> walk(lane, static_cast<uint32_t>(stride));
        const uint8_t* source = pixels + static_cast<size_t>(first_line) * stride + tile_index * (TW * 4);
        unsigned destination = static_cast<unsigned>(__cvta_generic_to_shared(tile));
#if defined(JLS_L2_HINTS) && (JLS_L2_HINTS & 2)
        uint64_t load_policy;
        asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(load_policy));
#endif
#pragma unroll
        for (int g = 0; g < TileWalk<TW>::groups; ++g)
        {
#pragma unroll
            for (int j = 0; j < TileWalk<TW>::period; ++j)
#if defined(JLS_L2_HINTS) && (JLS_L2_HINTS & 2)
                asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;\n" ::"r"(destination + walk.shared_offset[j]),
                             "l"(source + walk.global_offset[j]), "l"(load_policy)
                             : "memory");
#else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(destination + walk.shared_offset[j]),
                             "l"(source + walk.global_offset[j])
                             : "memory");
#endif
            source += TileWalk<TW>::rows * stride;
            destination += TileWalk<TW>::rows * SW * 4U;
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
        // a row that does not end on a word boundary: only its own bytes are read, the rest of the word is zero-filled
        cp_async_4(tile + r * SW + w, source, inside ? min(4, row_bytes - byte) : 0);
    }
    cp_async_commit();
}

// ---------------------------------------------------------------------------------------------------------------------
// A/B variant (JLS_TILE_LOAD_BULK, off by default): the same tile through the TMA unit's bulk copies, one row per lane
// (cp.async.bulk.shared.global, SASS UBLKCP) with an mbarrier that counts the bytes.  Two things speak against it, and the
// measurement agrees (profiles/r2_notes.md): a bulk copy takes its addresses from uniform registers, so 32 lanes with 32 row
// addresses run a 32-trip ELECT / R2UR / UBLKCP loop (7 issue slots per row, against 16 LDGSTS per lane for the whole tile),
// and its destination must be 16-byte aligned, so rows cannot be padded to an odd word count: the per-lane walk along a row
// then meets 4-way bank conflicts.  A 2-D tensor map (one UTMALDG per tile) has the same alignment constraint and needs a
// descriptor per frame from the driver API.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbarrier_init(uint32_t barrier, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barrier), "r"(count) : "memory");
}

__device__ __forceinline__ void mbarrier_wait(uint32_t barrier, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}" ::"r"(barrier),
                 "r"(parity)
                 : "memory");
}

// Row r of the warp's tile -> tile + r * row_stride_bytes (a multiple of 16); every lane arrives on the barrier (count 32)
// with the bytes of its own row.  Whole tiles only (the caller falls back to tile_load_async at the edges).
template<int TW>
__device__ __forceinline__ void tile_load_bulk(uint32_t tile_shared, uint32_t row_stride_bytes, uint32_t barrier, const uint8_t* pixels,
                                               size_t stride, uint32_t first_line, int32_t tile_index, uint32_t lane)
{
    const uint8_t* source = pixels + static_cast<size_t>(first_line + lane) * stride + tile_index * (TW * 4);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barrier), "n"(TW * 4) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tile_shared + lane * row_stride_bytes),
                 "l"(source), "n"(TW * 4), "r"(barrier)
                 : "memory");
}

// Writes tile `tile_index` back: row r goes to line first_line + r if bit r of `row_mask` is set.
// one to three bytes, once per row at most: kept out of line so that the tile kernels' register allocation does not see it
__device__ __noinline__ void store_row_tail(uint8_t* destination, uint32_t word, int32_t count)
{
    for (int32_t b = 0; b < count; ++b)
        destination[b] = static_cast<uint8_t>(word >> (8 * b));
}

template<int TW>
__device__ __forceinline__ void tile_store(const uint32_t* tile, uint8_t* pixels, size_t stride, uint32_t first_line,
                                           uint32_t row_mask, int32_t row_bytes, int32_t tile_index, uint32_t lane)
{
    if (row_mask == 0xFFFFFFFFU && (tile_index + 1) * (TW * 4) <= row_bytes)
    {
        // a whole tile: see tile_load_async
        const TileWalk<TW> walk(lane, static_cast<uint32_t>(stride));
        uint8_t* destination = pixels + static_cast<size_t>(first_line) * stride + tile_index * (TW * 4);
        unsigned source = static_cast<unsigned>(__cvta_generic_to_shared(tile));
#if defined(JLS_L2_HINTS) && (JLS_L2_HINTS & 8)
        uint64_t store_policy; // A/B: decoded tiles are written once and not read again by this kernel
        asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(store_policy));
#endif
#pragma unroll
        for (int g = 0; g < TileWalk<TW>::groups; ++g)
        {
#pragma unroll
            for (int j = 0; j < TileWalk<TW>::period; ++j)
            {
                uint32_t word;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(source + walk.shared_offset[j]) : "memory");
#if defined(JLS_L2_HINTS) && (JLS_L2_HINTS & 8)
                asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(destination + walk.global_offset[j]), "r"(word),
                             "l"(store_policy)
                             : "memory");
#else
                *reinterpret_cast<uint32_t*>(destination + walk.global_offset[j]) = word;
#endif
            }
            destination += TileWalk<TW>::rows * stride;
            source += TileWalk<TW>::rows * (TW + 1) * 4U;
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
        if (byte + 4 <= row_bytes && ((row_mask >> r) & 1U) != 0)
            *reinterpret_cast<uint32_t*>(pixels + static_cast<size_t>(first_line + r) * stride + byte) = tile[r * (TW + 1) + w];
    }
    // The last one to three bytes of rows that do not end on a word boundary, lane r for row r: what follows them belongs to
    // the caller (or to the next row when the stride leaves no room).
    const int32_t tail_byte = row_bytes & ~3;
    if ((row_bytes & 3) != 0 && tail_byte >= tile_index * (TW * 4) && tail_byte < (tile_index + 1) * (TW * 4) &&
        ((row_mask >> lane) & 1U) != 0)
        store_row_tail(pixels + static_cast<size_t>(first_line + lane) * stride + tail_byte,
                       tile[lane * (TW + 1) + (tail_byte - tile_index * (TW * 4)) / 4], row_bytes & 3);
}

} // namespace jls
