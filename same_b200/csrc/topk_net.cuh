// topk_net.cuh — branch-free compare-exchange networks on 32-bit keys held in registers (static indices only), used by the
// tile search of candidates.cu: a thread keeps the N smallest keys it has seen sorted in registers and folds batches of
// N pending keys into them.  Plain C++ (min/max), so tests/test_host_logic.py compiles this header with g++ and checks the
// networks exhaustively with the zero-one principle.
#pragma once

#if defined(__CUDACC__)
#define SAME_HD __host__ __device__ __forceinline__
#else
#define SAME_HD inline
#endif

namespace same {

SAME_HD unsigned net_min(unsigned a, unsigned b) { return a < b ? a : b; }
SAME_HD unsigned net_max(unsigned a, unsigned b) { return a < b ? b : a; }
SAME_HD void net_ce(unsigned &a, unsigned &b) {
    const unsigned lo = net_min(a, b), hi = net_max(a, b);
    a = lo;
    b = hi;
}

// Batcher's odd-even merge sort, ascending; N a power of two (5 / 19 / 63 compare-exchanges for N = 4 / 8 / 16).
template <int N>
SAME_HD void net_sort(unsigned (&v)[N]) {
    static_assert((N & (N - 1)) == 0, "net_sort: N must be a power of two");
#pragma unroll
    for (int p = 1; p < N; p <<= 1)
#pragma unroll
        for (int k = p; k >= 1; k >>= 1)
#pragma unroll
            for (int j = k % p; j + k < N; j += 2 * k)
#pragma unroll
                for (int i = 0; i < k; ++i)
                    if (i + j + k < N && (i + j) / (2 * p) == (i + j + k) / (2 * p)) net_ce(v[i + j], v[i + j + k]);
}

// best[] and pend[] sorted ascending.  Afterwards best[] holds the N smallest keys of the union, ascending; the return
// value is the smallest key that was dropped (the (N+1)-th smallest of the union).
// min(best[i], pend[N-1-i]) keeps exactly the N smallest and is a bitonic sequence, which log2(N) half-cleaner stages sort.
template <int N>
SAME_HD unsigned net_merge_smallest(unsigned (&best)[N], const unsigned (&pend)[N]) {
    unsigned dropped = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const unsigned a = best[i], b = pend[N - 1 - i];
        best[i] = net_min(a, b);
        dropped = net_min(dropped, net_max(a, b));
    }
#pragma unroll
    for (int j = N / 2; j >= 1; j >>= 1)
#pragma unroll
        for (int i = 0; i < N; ++i)
            if ((i & j) == 0) net_ce(best[i], best[i | j]);
    return dropped;
}

}  // namespace same
