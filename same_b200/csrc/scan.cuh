// scan.cuh — device-wide exclusive prefix sums INSIDE a kernel (single pass, decoupled look-back), so that the
// "count -> scan -> fill" triplets of the hot path (window subsetting, triangle remap, frame compaction, triangle
// filter, cut emission) are one launch each: a block counts what its items will emit, learns the total of all earlier
// blocks by looking back at their published aggregates, and writes its output in place.
//
// Tile = one thread block (blockIdx.x order; blocks of a 1-D grid are dispatched in index order, so a block only ever
// waits for blocks that are already resident or finished — the same assumption cub::DeviceScan's decoupled look-back makes).
// One scan state per Section: calls on one section must come from one host thread at a time (include/same_b200.h); different
// sections — one per worker thread and stream in CandidateStream — share nothing.  Per tile and per stream one 64-bit word
//     [ epoch : 30 | status : 2 | value : 32 ]
// written and read as ONE word, so no fence is needed between a value and its status.  status 1 = "aggregate of this
// tile", 2 = "inclusive prefix up to this tile".  The epoch makes a word of an earlier launch look empty, so the state
// buffer is zeroed once when it is allocated and never again (no memset per launch).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace same {

struct ScanCtx {
    unsigned long long *state;   // [n_streams][stride]
    int stride;                  // >= number of tiles of the launch
    unsigned epoch;              // unique per launch, 1 .. 2^30-1
};

__device__ __forceinline__ unsigned long long ts_pack(unsigned epoch, unsigned status, unsigned value) {
    return ((unsigned long long)epoch << 34) | ((unsigned long long)status << 32) | (unsigned long long)value;
}
__device__ __forceinline__ void ts_store(unsigned long long *p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ts_load(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Block-wide exclusive scan of NS independent int streams (one value per thread and stream).
// `warp_tot` is shared memory, at least NS * (THREADS / 32) ints (checked at compile time: an undersized buffer would
// silently scribble over the caller's other shared arrays).  total[s] is returned to every thread.
template <int NS, int THREADS, int N>
__device__ __forceinline__ void block_exclusive_scan(const int (&x)[NS], int (&excl)[NS], int (&total)[NS], int (&warp_tot)[N]) {
    constexpr int NW = THREADS / 32;
    static_assert(N >= NS * NW, "block_exclusive_scan: shared buffer smaller than NS * (THREADS / 32) ints");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int v = x[s];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        incl[s] = v;
        if (lane == 31) warp_tot[s * NW + warp] = v;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int before = 0, all = 0;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int wt = warp_tot[s * NW + k];
            all += wt;
            if (k < warp) before += wt;
        }
        excl[s] = before + incl[s] - x[s];
        total[s] = all;
    }
    __syncthreads();   // warp_tot may be reused by the caller
}

// Exclusive prefix of this tile = sum of the aggregates of tiles 0 .. tile-1, for NS streams at once (warp s resolves
// stream s; THREADS >= 32 * NS).  Must be called by every thread of the block; `agg` is block-uniform.
// `sh` is shared memory, NS ints.
template <int NS>
__device__ __forceinline__ void tile_exclusive_prefix(const ScanCtx &c, int tile, const int (&agg)[NS], int (&prefix)[NS], int *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < NS) {
        int a = 0;
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (s == warp) a = agg[s];
        unsigned long long *st = c.state + (size_t)warp * (size_t)c.stride;
        int p = 0;
        if (tile == 0) {
            if (lane == 0) ts_store(st, ts_pack(c.epoch, 2u, (unsigned)a));
        } else {
            if (lane == 0) ts_store(st + tile, ts_pack(c.epoch, 1u, (unsigned)a));
            int base = tile - 1;
            for (;;) {
                const int idx = base - lane;
                unsigned long long v;
                bool ok;
                do {   // wait until the 32 predecessors in the window have published something in this epoch
                    v = idx >= 0 ? ts_load(st + idx) : ts_pack(c.epoch, 2u, 0u);
                    ok = (unsigned)(v >> 34) == c.epoch && ((unsigned)(v >> 32) & 3u) != 0u;
                } while (!__all_sync(0xffffffffu, ok));
                const unsigned pm = __ballot_sync(0xffffffffu, ((unsigned)(v >> 32) & 3u) == 2u);
                const int stop = pm ? __ffs(pm) - 1 : 31;   // nearest predecessor that already holds an inclusive prefix
                int contrib = lane <= stop ? (int)(unsigned)v : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                p += contrib;
                if (pm) break;
                base -= 32;
            }
            if (lane == 0) ts_store(st + tile, ts_pack(c.epoch, 2u, (unsigned)(p + a)));
        }
        if (lane == 0) sh[warp] = p;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; ++s) prefix[s] = sh[s];
    __syncthreads();
}

// Segmented variant for sums that restart at segment boundaries (violated-triangle ranks restart at every window):
// returns the carry-in of the tile = number of counted items of the segment that is open at the tile's first item, in
// all earlier tiles.  A tile that contains a boundary publishes right away the count of its LAST segment as final (the
// look-back of later tiles stops there); a tile without a boundary publishes its whole count as an aggregate first and
// carry + count once it knows its carry.  Must be called by every thread of the block; arguments are block-uniform.
__device__ __forceinline__ int tile_segmented_carry(const ScanCtx &c, int tile, int count_all, bool has_boundary, int tail_count, int *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        unsigned long long *st = c.state;
        int p = 0;
        if (tile == 0) {
            if (lane == 0) ts_store(st, ts_pack(c.epoch, 2u, (unsigned)(has_boundary ? tail_count : count_all)));
        } else {
            if (lane == 0) ts_store(st + tile, has_boundary ? ts_pack(c.epoch, 2u, (unsigned)tail_count) : ts_pack(c.epoch, 1u, (unsigned)count_all));
            int base = tile - 1;
            for (;;) {
                const int idx = base - lane;
                unsigned long long v;
                bool ok;
                do {
                    v = idx >= 0 ? ts_load(st + idx) : ts_pack(c.epoch, 2u, 0u);
                    ok = (unsigned)(v >> 34) == c.epoch && ((unsigned)(v >> 32) & 3u) != 0u;
                } while (!__all_sync(0xffffffffu, ok));
                const unsigned pm = __ballot_sync(0xffffffffu, ((unsigned)(v >> 32) & 3u) == 2u);
                const int stop = pm ? __ffs(pm) - 1 : 31;
                int contrib = lane <= stop ? (int)(unsigned)v : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                p += contrib;
                if (pm) break;
                base -= 32;
            }
            if (!has_boundary && lane == 0) ts_store(st + tile, ts_pack(c.epoch, 2u, (unsigned)(p + count_all)));
        }
        if (lane == 0) sh[0] = p;
    }
    __syncthreads();
    const int carry = sh[0];
    __syncthreads();
    return carry;
}

// Convenience: per-thread value x[s] -> its device-wide exclusive prefix; `total_before_tile + block total` of the last
// tile is the grand total.  smem: NS * (THREADS/32) + NS ints.
template <int NS, int THREADS, int N>
__device__ __forceinline__ void device_exclusive_scan(const ScanCtx &c, int tile, const int (&x)[NS], int (&excl)[NS], int (&tile_total)[NS],
                                                      int (&tile_prefix)[NS], int (&smem)[N]) {
    static_assert(N >= NS * (THREADS / 32) + NS, "device_exclusive_scan: shared buffer smaller than NS * (THREADS / 32) + NS ints");
    int local[NS];
    block_exclusive_scan<NS, THREADS>(x, local, tile_total, smem);
    tile_exclusive_prefix<NS>(c, tile, tile_total, tile_prefix, smem + NS * (THREADS / 32));
#pragma unroll
    for (int s = 0; s < NS; ++s) excl[s] = tile_prefix[s] + local[s];
}

}  // namespace same
