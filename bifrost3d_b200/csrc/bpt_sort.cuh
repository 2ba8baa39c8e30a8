// Device-wide primitives of the build and of the queue reordering, hand-written for sm_100a (no CUB on the path):
//   * exclusive_scan: reduce / spine / apply, three launches, any length
//   * radix_sort_pairs: least-significant-digit radix sort of (key, 32-bit value) pairs, 8 bits per pass, stable.
//     One pass = per-tile digit histograms (digit-major) -> ONE exclusive scan over digits x tiles, which yields for every
//     (digit, tile) the global position of the tile's first key with that digit -> scatter. The scatter ranks the keys of
//     a tile with warp match + popc (no atomics), stages them in shared memory in sorted order and writes them out in
//     digit runs, so the global stores are coalesced run by run.
// Every kernel takes the element count either by value or from device memory (count_ptr), so that the wavefront queues,
// whose lengths only exist on the device, can be sorted without a host round trip.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bpt {
namespace sort {

constexpr int BLOCK = 256;                 // threads per CTA of every kernel in this file
constexpr int SCAN_ITEMS = 8;              // elements per thread of the scan kernels
constexpr int SCAN_TILE = BLOCK * SCAN_ITEMS;
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_ITEMS = 8;              // keys per thread of the sort kernels
constexpr int SORT_TILE = BLOCK * SORT_ITEMS; // 2048 keys: 16 KB of 64-bit keys + 8 KB of values in shared memory
constexpr int WARPS = BLOCK / 32;

__device__ __forceinline__ uint32_t element_count(uint32_t n, const uint32_t* __restrict__ count_ptr) { return count_ptr ? *count_ptr : n; }

// Exclusive prefix sum over the BLOCK values of a CTA; `total` receives the CTA's sum. `warp_sums`: 32 words of shared memory.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inclusive = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inclusive, o); if (lane >= o) inclusive += t; }
    if (lane == 31) warp_sums[warp] = inclusive;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < WARPS ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        warp_sums[lane] = s;
    }
    __syncthreads();
    const uint32_t warp_offset = warp ? warp_sums[warp - 1] : 0u;
    total = warp_sums[WARPS - 1];
    __syncthreads(); // warp_sums may be reused by the caller
    return warp_offset + inclusive - v;
}

// ---- exclusive scan of 32-bit counts ---------------------------------------------------------------------------------

static __global__ void __launch_bounds__(BLOCK) scan_reduce_kernel(const uint32_t* __restrict__ in, uint32_t n, const uint32_t* __restrict__ count_ptr,
                                                            uint32_t* __restrict__ tile_sums, uint32_t tiles) {
    __shared__ uint32_t warp_sums[32];
    n = element_count(n, count_ptr);
    tiles = min(tiles, (n + SCAN_TILE - 1) / SCAN_TILE); // a count from device memory can be far below the capacity the launch is sized for
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) sum += in[base + i];
        uint32_t total;
        block_exclusive_scan(sum, warp_sums, total);
        if (threadIdx.x == 0) tile_sums[tile] = total;
    }
}

// One CTA: exclusive scan of the tile sums in place; the grand total goes to total_out (if given).
static __global__ void __launch_bounds__(BLOCK) scan_spine_kernel(uint32_t* __restrict__ tile_sums, uint32_t tiles, uint32_t* __restrict__ total_out,
                                                                  uint32_t n, const uint32_t* __restrict__ count_ptr) {
    __shared__ uint32_t warp_sums[32];
    tiles = min(tiles, (element_count(n, count_ptr) + SCAN_TILE - 1) / SCAN_TILE);
    uint32_t carry = 0;
    for (uint32_t base = 0; base < tiles; base += BLOCK) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < tiles ? tile_sums[i] : 0u;
        uint32_t total;
        const uint32_t exclusive = block_exclusive_scan(v, warp_sums, total);
        if (i < tiles) tile_sums[i] = carry + exclusive;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

static __global__ void __launch_bounds__(BLOCK) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n,
                                                           const uint32_t* __restrict__ count_ptr, const uint32_t* __restrict__ tile_offsets, uint32_t tiles) {
    __shared__ uint32_t warp_sums[32];
    n = element_count(n, count_ptr);
    tiles = min(tiles, (n + SCAN_TILE - 1) / SCAN_TILE);
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = base + i < n ? in[base + i] : 0u; sum += v[i]; }
        uint32_t total;
        uint32_t running = tile_offsets[tile] + block_exclusive_scan(sum, warp_sums, total);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = running; running += v[i]; }
    }
}

inline uint32_t scan_tiles(uint32_t capacity) { return (capacity + SCAN_TILE - 1) / SCAN_TILE; }
// Scratch words exclusive_scan needs for up to `capacity` elements.
inline size_t scan_scratch_words(uint32_t capacity) { return (size_t)scan_tiles(capacity) + 1; }

// out[i] = sum of in[0 .. i). `in` and `out` may alias. n elements, or *count_ptr (then n is the capacity the launch is
// sized for). total_out (device, optional) receives the sum of all elements.
inline void exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, const uint32_t* count_ptr, uint32_t* scratch, uint32_t* total_out,
                           int sm_count, cudaStream_t stream) {
    const uint32_t tiles = scan_tiles(n);
    if (tiles == 0) { if (total_out) cudaMemsetAsync(total_out, 0, sizeof(uint32_t), stream); return; }
    const int grid = (int)(tiles < (uint32_t)sm_count * 8u ? tiles : (uint32_t)sm_count * 8u);
    scan_reduce_kernel<<<grid, BLOCK, 0, stream>>>(in, n, count_ptr, scratch, tiles);
    scan_spine_kernel<<<1, BLOCK, 0, stream>>>(scratch, tiles, total_out, n, count_ptr);
    scan_apply_kernel<<<grid, BLOCK, 0, stream>>>(in, out, n, count_ptr, scratch, tiles);
}
constexpr int SCAN_LAUNCHES = 3;

// ---- radix sort ------------------------------------------------------------------------------------------------------

template <typename Key>
__device__ __forceinline__ uint32_t digit_of(Key key, int shift) { return (uint32_t)(key >> shift) & (RADIX - 1); }

// Digit histogram of every tile, digit-major: tile_hist[digit * tiles + tile].
template <typename Key>
__global__ void __launch_bounds__(BLOCK) radix_histogram_kernel(const Key* __restrict__ keys, uint32_t n, const uint32_t* __restrict__ count_ptr, int shift,
                                                                uint32_t* __restrict__ tile_hist, uint32_t tiles) {
    __shared__ uint32_t hist[RADIX];
    n = element_count(n, count_ptr);
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        hist[threadIdx.x] = 0; // BLOCK == RADIX
        __syncthreads();
        const uint32_t base = tile * SORT_TILE;
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            const uint32_t index = base + i * BLOCK + threadIdx.x;
            if (index < n) atomicAdd(&hist[digit_of(keys[index], shift)], 1u);
        }
        __syncthreads();
        tile_hist[threadIdx.x * tiles + tile] = hist[threadIdx.x];
        __syncthreads();
    }
}

// Stable scatter of one tile per CTA iteration. Warp w owns the contiguous chunk [w * 32 * SORT_ITEMS, ...) of the tile and
// walks it in rounds of 32 consecutive keys, so (warp, round, lane) order is memory order.
template <typename Key>
__global__ void __launch_bounds__(BLOCK) radix_scatter_kernel(const Key* __restrict__ keys_in, const uint32_t* __restrict__ values_in,
                                                              Key* __restrict__ keys_out, uint32_t* __restrict__ values_out, uint32_t n,
                                                              const uint32_t* __restrict__ count_ptr, int shift,
                                                              const uint32_t* __restrict__ tile_offsets /* scanned tile_hist */, uint32_t tiles) {
    __shared__ Key s_keys[SORT_TILE];
    __shared__ uint32_t s_values[SORT_TILE];
    __shared__ uint32_t s_warp_hist[WARPS][RADIX]; // running per-warp digit counts, then the warp's base inside the tile's digit run
    __shared__ uint32_t s_digit_base[RADIX];       // first tile-local sorted slot of every digit
    __shared__ uint32_t s_global_base[RADIX];      // global position of the tile's first key of every digit
    __shared__ uint32_t s_warp_sums[32];
    static_assert(BLOCK == RADIX, "one thread per digit");
    n = element_count(n, count_ptr);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lanes_below = (1u << lane) - 1u;

    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t tile_base = tile * SORT_TILE;
        if (tile_base >= n) break; // uniform: n is the same for every thread
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s_warp_hist[w][threadIdx.x] = 0;
        __syncthreads();

        // Phase A: rank every key among the earlier keys of its warp's chunk with the same digit.
        Key key[SORT_ITEMS]; uint32_t value[SORT_ITEMS]; uint32_t rank[SORT_ITEMS];
        const uint32_t chunk_base = tile_base + warp * (32 * SORT_ITEMS);
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; ++r) {
            const uint32_t index = chunk_base + r * 32 + lane;
            const bool valid = index < n;
            key[r] = valid ? keys_in[index] : Key(0);
            value[r] = valid ? values_in[index] : 0u;
            // invalid lanes (past the end) take a digit of their own class: RADIX, matched separately and never stored
            const uint32_t digit = valid ? digit_of(key[r], shift) : (uint32_t)RADIX;
            const uint32_t peers = __match_any_sync(0xffffffffu, digit);
            uint32_t before = 0;
            if (valid) before = s_warp_hist[warp][digit];
            __syncwarp();
            rank[r] = before + __popc(peers & lanes_below);
            if (valid && (peers & lanes_below) == 0u) s_warp_hist[warp][digit] = before + __popc(peers); // the group's first lane
            __syncwarp();
        }
        __syncthreads();

        // Phase B: thread d turns the per-warp counts of digit d into per-warp bases and gets the tile's count of d.
        uint32_t count = 0;
        {
            const uint32_t d = threadIdx.x;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { uint32_t c = s_warp_hist[w][d]; s_warp_hist[w][d] = count; count += c; }
            s_global_base[d] = tile_offsets[d * tiles + tile];
        }
        uint32_t tile_total;
        const uint32_t digit_base = block_exclusive_scan(count, s_warp_sums, tile_total);
        s_digit_base[threadIdx.x] = digit_base;
        __syncthreads();

        // Phase C: stage in sorted order.
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; ++r) {
            const uint32_t index = chunk_base + r * 32 + lane;
            if (index < n) {
                const uint32_t d = digit_of(key[r], shift);
                const uint32_t slot = s_digit_base[d] + s_warp_hist[warp][d] + rank[r];
                s_keys[slot] = key[r]; s_values[slot] = value[r];
            }
        }
        __syncthreads();

        // Phase D: write out; consecutive threads hold consecutive keys of a digit run -> coalesced run by run.
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            const uint32_t slot = i * BLOCK + threadIdx.x;
            if (slot < tile_total) {
                const Key k = s_keys[slot];
                const uint32_t d = digit_of(k, shift);
                const uint32_t dst = s_global_base[d] + (slot - s_digit_base[d]);
                keys_out[dst] = k; values_out[dst] = s_values[slot];
            }
        }
        __syncthreads();
    }
}

inline uint32_t sort_tiles(uint32_t capacity) { return (capacity + SORT_TILE - 1) / SORT_TILE; }
// Scratch words radix_sort_pairs needs for up to `capacity` pairs: the digits x tiles histogram and the scan's own scratch.
inline size_t sort_scratch_words(uint32_t capacity) {
    const size_t hist = (size_t)RADIX * sort_tiles(capacity);
    return hist + scan_scratch_words((uint32_t)hist);
}

// Sorts (key, value) pairs by bits [begin_bit, end_bit) of the key, stable. The pairs ping-pong between (keys, values) and
// (keys_alt, values_alt); returns 0 when the sorted pairs end up in (keys, values) and 1 when they end up in the _alt
// arrays. n pairs, or *count_ptr pairs with n the capacity the launches are sized for. `launches` counts kernels launched.
template <typename Key>
inline int radix_sort_pairs(Key* keys, uint32_t* values, Key* keys_alt, uint32_t* values_alt, uint32_t n, const uint32_t* count_ptr,
                            int begin_bit, int end_bit, uint32_t* scratch, int sm_count, cudaStream_t stream, uint64_t* launches = nullptr) {
    const uint32_t tiles = sort_tiles(n);
    if (tiles == 0) return 0;
    const uint32_t hist_words = RADIX * tiles;
    uint32_t* tile_hist = scratch;
    uint32_t* scan_scratch = scratch + hist_words;
    const int grid = (int)(tiles < (uint32_t)sm_count * 8u ? tiles : (uint32_t)sm_count * 8u);
    int current = 0;
    for (int shift = begin_bit; shift < end_bit; shift += RADIX_BITS) {
        Key* k_in = current ? keys_alt : keys; uint32_t* v_in = current ? values_alt : values;
        Key* k_out = current ? keys : keys_alt; uint32_t* v_out = current ? values : values_alt;
        radix_histogram_kernel<Key><<<grid, BLOCK, 0, stream>>>(k_in, n, count_ptr, shift, tile_hist, tiles);
        exclusive_scan(tile_hist, tile_hist, hist_words, nullptr, scan_scratch, nullptr, sm_count, stream);
        radix_scatter_kernel<Key><<<grid, BLOCK, 0, stream>>>(k_in, v_in, k_out, v_out, n, count_ptr, shift, tile_hist, tiles);
        if (launches) *launches += 2 + SCAN_LAUNCHES;
        current ^= 1;
    }
    return current;
}

} // namespace sort
} // namespace bpt
