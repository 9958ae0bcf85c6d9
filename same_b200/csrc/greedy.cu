// greedy.cu — ordered greedy selection of items whose endpoints must stay pairwise disjoint.
//
// Two loops of the reference have this shape: walk the items in ascending (key, original index) order (Python's
// stable sort) and take an item iff none of its endpoints has been taken before:
//   * the greedy MIP start (src/init_helpers.py:110-132): items = candidate pairs, key = cost, endpoints =
//     (aligned i, reference j), only rows whose best cost beats their no-match penalty take part;
//   * the batch selection of greedy_triangle_collapse (src/metacell_utils.py:423-433): items = collapsible
//     triangles, key = perimeter, endpoints = the three vertices.
// The sequential result is the unique fixed point of "an item is taken iff it is the first alive item at EVERY one
// of its endpoints; items touching a taken endpoint die", so it is computed in rounds: every alive item proposes its
// rank to its endpoints (64-bit atomicMin tagged with the round, so nothing is reset between rounds), items that win
// all their endpoints are taken, their neighbours die.  Rounds needed ~ the longest rank-monotone chain of
// conflicting items (O(log n) on geometric inputs); after MAX_ROUNDS a single thread finishes the remainder in rank
// order, which keeps pathological inputs correct.
#include "common.cuh"

namespace same {

typedef unsigned long long u64;
constexpr int GREEDY_MAX_ROUNDS = 64;

// monotone u64 image of a double (ascending; -0.0 and +0.0 coincide, as they compare equal in the reference's sort)
__device__ __forceinline__ u64 order_bits(double d) {
    if (d == 0.0) d = 0.0;
    const u64 b = (u64)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void k_greedy_keys(const double *__restrict__ key, const unsigned char *__restrict__ eligible, i64 n, u64 *__restrict__ kb,
                              i32 *__restrict__ idx, unsigned char *__restrict__ alive) {
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    kb[e] = order_bits(key[e]);
    idx[e] = (i32)e;
    alive[e] = eligible ? (eligible[e] != 0) : 1;
}
__global__ void k_greedy_rank(const i32 *__restrict__ sorted_idx, i64 n, i32 *__restrict__ rank) {
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) rank[sorted_idx[r]] = (i32)r;
}
// prune (an endpoint was taken in the previous round) + propose; counts the items still alive
template <int D>
__global__ void k_greedy_propose(const i32 *__restrict__ nodes, const i32 *__restrict__ rank, i64 n, unsigned round,
                                 const unsigned char *__restrict__ used, unsigned char *__restrict__ alive, u64 *__restrict__ best,
                                 i32 *__restrict__ n_alive) {
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = e < n && alive[e];
    if (live) {
        i32 v[D];
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = nodes[e * D + d];
        bool hit = false;
#pragma unroll
        for (int d = 0; d < D; ++d) hit |= used[v[d]] != 0;
        if (hit) { alive[e] = 0; live = false; }
        else {
            const u64 tag = ((u64)(0x7fffffffu - round) << 32) | (u64)(unsigned)rank[e];   // later rounds sort first
#pragma unroll
            for (int d = 0; d < D; ++d) atomicMin(best + v[d], tag);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, live);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_alive, __popc(m));
}
template <int D>
__global__ void k_greedy_commit(const i32 *__restrict__ nodes, const i32 *__restrict__ rank, i64 n, unsigned round,
                                const u64 *__restrict__ best, unsigned char *__restrict__ used, unsigned char *__restrict__ alive,
                                unsigned char *__restrict__ selected) {
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n || !alive[e]) return;
    const u64 tag = ((u64)(0x7fffffffu - round) << 32) | (u64)(unsigned)rank[e];
    i32 v[D];
    bool win = true;
#pragma unroll
    for (int d = 0; d < D; ++d) { v[d] = nodes[e * D + d]; win &= best[v[d]] == tag; }
    if (!win) return;
    selected[e] = 1;
    alive[e] = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) used[v[d]] = 1;
}
// remainder after GREEDY_MAX_ROUNDS, in rank order (one thread: the sequential loop itself)
template <int D>
__global__ void k_greedy_tail(const i32 *__restrict__ nodes, const i32 *__restrict__ sorted_idx, i64 n, unsigned char *__restrict__ used,
                              unsigned char *__restrict__ alive, unsigned char *__restrict__ selected) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (i64 r = 0; r < n; ++r) {
        const i32 e = sorted_idx[r];
        if (!alive[e]) continue;
        bool hit = false;
        for (int d = 0; d < D; ++d) hit |= used[nodes[(i64)e * D + d]] != 0;
        if (!hit) {
            selected[e] = 1;
            for (int d = 0; d < D; ++d) used[nodes[(i64)e * D + d]] = 1;
        }
        alive[e] = 0;
    }
}

// device arrays in, device arrays out; `selected` [n] and `used` [n_nodes] are written (zeroed here).  Returns the rounds run.
template <int D>
static int greedy_select_dev(const i32 *nodes, const double *key, const unsigned char *eligible, i64 n, i64 n_nodes, unsigned char *selected,
                             unsigned char *used, cudaStream_t s) {
    CK(cudaMemsetAsync(selected, 0, (size_t)std::max<i64>(n, 1), s));
    CK(cudaMemsetAsync(used, 0, (size_t)std::max<i64>(n_nodes, 1), s));
    if (n == 0) return 0;
    DevBuf<u64> kb, kb2, best;
    DevBuf<i32> idx, idx2, rank, d_alive;
    DevBuf<unsigned char> alive, tmp;
    kb.alloc(n, s); kb2.alloc(n, s); idx.alloc(n, s); idx2.alloc(n, s); rank.alloc(n, s); alive.alloc(n, s);
    best.alloc(n_nodes, s); d_alive.alloc(GREEDY_MAX_ROUNDS + 1, s);
    CK(cudaMemsetAsync(best.p, 0xff, sizeof(u64) * (size_t)n_nodes, s));
    d_alive.zero(s);
    const unsigned blocks = blocks_for(n, 256);
    LAUNCH(k_greedy_keys, blocks, 256, 0, s, key, eligible, n, kb.p, idx.p, alive.p);
    size_t bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kb.p, kb2.p, idx.p, idx2.p, (int)n, 0, 64, s));
    tmp.alloc((i64)bytes, s);
    CK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kb.p, kb2.p, idx.p, idx2.p, (int)n, 0, 64, s));   // stable: ties keep index order
    g_launches.fetch_add(1, std::memory_order_relaxed);
    LAUNCH(k_greedy_rank, blocks, 256, 0, s, idx2.p, n, rank.p);
    i32 *h_alive = nullptr;
    CK(cudaHostAlloc((void **)&h_alive, sizeof(i32), cudaHostAllocDefault));
    int rounds = 0;
    try {
        for (;;) {
            LAUNCH(k_greedy_propose<D>, blocks, 256, 0, s, nodes, rank.p, n, (unsigned)rounds, used, alive.p, best.p, d_alive.p + rounds);
            CK(cudaMemcpyAsync(h_alive, d_alive.p + rounds, sizeof(i32), cudaMemcpyDeviceToHost, s));
            CK(stream_wait(s));
            if (*h_alive == 0) break;
            if (rounds == GREEDY_MAX_ROUNDS) {
                LAUNCH(k_greedy_tail<D>, 1, 32, 0, s, nodes, idx2.p, n, used, alive.p, selected);
                break;
            }
            LAUNCH(k_greedy_commit<D>, blocks, 256, 0, s, nodes, rank.p, n, (unsigned)rounds, best.p, used, alive.p, selected);
            ++rounds;
        }
    } catch (...) {
        cudaFreeHost(h_alive);
        throw;
    }
    cudaFreeHost(h_alive);
    return rounds;
}

static int greedy_select_any(int degree, const i32 *nodes, const double *key, const unsigned char *eligible, i64 n, i64 n_nodes,
                             unsigned char *selected, unsigned char *used, cudaStream_t s) {
    if (degree == 2) return greedy_select_dev<2>(nodes, key, eligible, n, n_nodes, selected, used, s);
    if (degree == 3) return greedy_select_dev<3>(nodes, key, eligible, n, n_nodes, selected, used, s);
    if (degree == 1) return greedy_select_dev<1>(nodes, key, eligible, n, n_nodes, selected, used, s);
    throw Error(SAME_E_ARG, "degree must be 1, 2 or 3");
}

// stateless form: host or device arrays in, host arrays out
void greedy_select_arrays(int device, i64 n, int degree, const i32 *nodes, const double *key, const unsigned char *eligible, i64 n_nodes,
                          unsigned char *selected, unsigned char *used_out, i32 *rounds_out) {
    CK(cudaSetDevice(device));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    try {
        DevBuf<i32> d_nodes;
        DevBuf<double> d_key;
        DevBuf<unsigned char> d_el, d_sel, d_used;
        d_nodes.alloc(n * degree, s); d_key.alloc(n, s); d_sel.alloc(n, s); d_used.alloc(n_nodes, s);
        if (n) {
            CK(cudaMemcpyAsync(d_nodes.p, nodes, sizeof(i32) * (size_t)(n * degree), cudaMemcpyDefault, s));
            CK(cudaMemcpyAsync(d_key.p, key, sizeof(double) * (size_t)n, cudaMemcpyDefault, s));
            if (eligible) {
                d_el.alloc(n, s);
                CK(cudaMemcpyAsync(d_el.p, eligible, (size_t)n, cudaMemcpyDefault, s));
            }
        }
        const int rounds = greedy_select_any(degree, d_nodes.p, d_key.p, eligible ? d_el.p : nullptr, n, n_nodes, d_sel.p, d_used.p, s);
        if (n && selected) CK(cudaMemcpyAsync(selected, d_sel.p, (size_t)n, cudaMemcpyDefault, s));
        if (n_nodes && used_out) CK(cudaMemcpyAsync(used_out, d_used.p, (size_t)n_nodes, cudaMemcpyDefault, s));
        CK(stream_wait(s));
        if (rounds_out) *rounds_out = rounds;
    } catch (...) {
        stream_wait(s);
        cudaStreamDestroy(s);
        throw;
    }
    CK(stream_wait(s));
    CK(cudaStreamDestroy(s));
}

// ---- batch selection of greedy_triangle_collapse (src/metacell_utils.py:388-433) --------------------------------
// np.linalg.norm of a 2-vector is sqrt(ddot(v, v)); ddot on FMA hosts is fma(vy, vy, vx*vx), the three norms
// are added left to right.  The perimeter only orders the candidates, but on lattice-like data (ISS spots) many perimeters tie
// to the last bit, so the order is only reproduced with the reference's own rounding.
__device__ __forceinline__ double norm2_ref(double vx, double vy) { return __dsqrt_rn(__fma_rn(vy, vy, __dmul_rn(vx, vx))); }
__global__ void k_collapse_score(const double2 *__restrict__ xy, const i32 *__restrict__ type, const double *__restrict__ size, const i32 *__restrict__ tri,
                                 i64 T, double max_size, unsigned char *__restrict__ cand, double *__restrict__ perim) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const i32 a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
    const bool same = (type[a] == type[b]) & (type[b] == type[c]);                              // :391-394
    const double total = __dadd_rn(__dadd_rn(size[a], size[b]), size[c]);                       // :397-399
    const double2 A = xy[a], B = xy[b], C = xy[c];
    const double p = __dadd_rn(__dadd_rn(norm2_ref(__dsub_rn(A.x, B.x), __dsub_rn(A.y, B.y)), norm2_ref(__dsub_rn(B.x, C.x), __dsub_rn(B.y, C.y))),
                               norm2_ref(__dsub_rn(C.x, A.x), __dsub_rn(C.y, A.y)));            // :405-409
    cand[t] = same && !(total > max_size);
    perim[t] = p;
}

// score + select in one call, everything device-resident in between; host or device arrays in, host arrays out
void collapse_select_arrays(int device, i64 n, const double *xy, const i32 *type, const double *size, i64 T, const i32 *tri, double max_size,
                            unsigned char *selected, double *perim_out, i32 *rounds_out) {
    CK(cudaSetDevice(device));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    try {
        DevBuf<double2> d_xy;
        DevBuf<i32> d_type, d_tri;
        DevBuf<double> d_size, d_perim;
        DevBuf<unsigned char> d_cand, d_sel, d_used;
        d_xy.alloc(n, s); d_type.alloc(n, s); d_size.alloc(n, s); d_tri.alloc(3 * T, s); d_perim.alloc(T, s); d_cand.alloc(T, s); d_sel.alloc(T, s);
        d_used.alloc(n, s);
        int rounds = 0;
        if (T > 0) {
            CK(cudaMemcpyAsync(d_xy.p, xy, sizeof(double2) * (size_t)n, cudaMemcpyDefault, s));
            CK(cudaMemcpyAsync(d_type.p, type, sizeof(i32) * (size_t)n, cudaMemcpyDefault, s));
            CK(cudaMemcpyAsync(d_size.p, size, sizeof(double) * (size_t)n, cudaMemcpyDefault, s));
            CK(cudaMemcpyAsync(d_tri.p, tri, sizeof(i32) * (size_t)(3 * T), cudaMemcpyDefault, s));
            LAUNCH(k_collapse_score, blocks_for(T, 256), 256, 0, s, d_xy.p, d_type.p, d_size.p, d_tri.p, T, max_size, d_cand.p, d_perim.p);
            rounds = greedy_select_dev<3>(d_tri.p, d_perim.p, d_cand.p, T, n, d_sel.p, d_used.p, s);
            CK(cudaMemcpyAsync(selected, d_sel.p, (size_t)T, cudaMemcpyDefault, s));
            if (perim_out) CK(cudaMemcpyAsync(perim_out, d_perim.p, sizeof(double) * (size_t)T, cudaMemcpyDefault, s));
        }
        CK(stream_wait(s));
        if (rounds_out) *rounds_out = rounds;
    } catch (...) {
        stream_wait(s);
        cudaStreamDestroy(s);
        throw;
    }
    CK(stream_wait(s));
    CK(cudaStreamDestroy(s));
}

// ---- member means of merged metacells (src/metacell_utils.py:446-474) ---------------------------------------------
// `rows[col].mean()` of the reference is pandas' nanmean = numpy's pairwise sum of the gathered members divided by their
// count (checked against pandas in tests/test_oracle_golden.py).  The sum is reproduced operation for operation: fewer than 8
// values are added left to right starting from 0; up to 128 values go through eight running sums combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail; longer lists split in halves (first half rounded down to a
// multiple of 8).  One thread per (merged metacell, column).
__device__ double pairwise_sum_ref(const double *__restrict__ V, i64 C, i64 c, const i32 *__restrict__ pos, i64 n) {
    if (n < 8) {
        double r = 0.0;
        for (i64 i = 0; i < n; ++i) r = __dadd_rn(r, V[(i64)pos[i] * C + c]);
        return r;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = V[(i64)pos[k] * C + c];
        i64 i = 8;
        for (; i < n - (n % 8); i += 8)
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], V[(i64)pos[i + k] * C + c]);
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, V[(i64)pos[i] * C + c]);
        return res;
    }
    i64 n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(pairwise_sum_ref(V, C, c, pos, n2), pairwise_sum_ref(V, C, c, pos + n2, n - n2));
}
__global__ void k_member_mean(const double *__restrict__ V, i64 C, const i32 *__restrict__ pos, const i64 *__restrict__ ptr, i64 G,
                              double *__restrict__ out) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * C) return;
    const i64 g = t / C, c = t - g * C;
    const i64 lo = ptr[g], n = ptr[g + 1] - lo;
    out[t] = n > 0 ? __ddiv_rn(pairwise_sum_ref(V, C, c, pos + lo, n), (double)n) : __longlong_as_double(0x7ff8000000000000ll);
}

void segment_mean_arrays(int device, i64 n_rows, i64 C, const double *values, i64 G, const i64 *ptr, i64 n_members, const i32 *pos, double *out) {
    CK(cudaSetDevice(device));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    try {
        DevBuf<double> d_v, d_out;
        DevBuf<i32> d_pos;
        DevBuf<i64> d_ptr;
        d_v.alloc(n_rows * C, s); d_out.alloc(G * C, s); d_pos.alloc(n_members, s); d_ptr.alloc(G + 1, s);
        if (G > 0 && C > 0) {
            CK(cudaMemcpyAsync(d_v.p, values, sizeof(double) * (size_t)(n_rows * C), cudaMemcpyDefault, s));
            if (n_members) CK(cudaMemcpyAsync(d_pos.p, pos, sizeof(i32) * (size_t)n_members, cudaMemcpyDefault, s));
            CK(cudaMemcpyAsync(d_ptr.p, ptr, sizeof(i64) * (size_t)(G + 1), cudaMemcpyDefault, s));
            LAUNCH(k_member_mean, blocks_for(G * C, 128), 128, 0, s, d_v.p, C, d_pos.p, d_ptr.p, G, d_out.p);
            CK(cudaMemcpyAsync(out, d_out.p, sizeof(double) * (size_t)(G * C), cudaMemcpyDefault, s));
        }
        CK(stream_wait(s));
    } catch (...) {
        stream_wait(s);
        cudaStreamDestroy(s);
        throw;
    }
    CK(stream_wait(s));
    CK(cudaStreamDestroy(s));
}

// ---- greedy MIP start over a whole batch (src/init_helpers.py:110-132) ---------------------------------------
// one thread per kept aligned row: prefer_match = (best cost of the row) < no_match_penalty * size
__global__ void k_start_rows(const i32 *__restrict__ row_ptr, const double *__restrict__ cost, const double *__restrict__ ka_size, i64 nKA,
                             double no_match_penalty, unsigned char *__restrict__ prefer) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nKA) return;
    double best = INFINITY;
    for (i32 p = row_ptr[i]; p < row_ptr[i + 1]; ++p) best = fmin(best, cost[p]);   // costs are never NaN
    prefer[i] = best < __dmul_rn(no_match_penalty, ka_size[i]);
}
// endpoints of pair p in one batch-wide node space: kept aligned rows [0, nKA), kept reference rows [nKA, nKA + nKR)
__global__ void k_start_items(const int2 *__restrict__ pairs, i64 P, const i32 *__restrict__ p_off, const i32 *__restrict__ ka_off,
                              const i32 *__restrict__ kr_off, int W, i64 nKA, const unsigned char *__restrict__ prefer, i32 *__restrict__ nodes,
                              unsigned char *__restrict__ eligible) {
    const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int w = find_window(p_off, W, (i32)p);
    const int2 ij = pairs[p];
    const i32 a = ka_off[w] + ij.x;
    nodes[2 * p] = a;
    nodes[2 * p + 1] = (i32)nKA + kr_off[w] + ij.y;
    eligible[p] = prefer[a];
}
__global__ void k_start_unmatched(const unsigned char *__restrict__ used, i64 nKA, unsigned char *__restrict__ unmatched) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nKA) unmatched[i] = !used[i];
}

void batch_mip_start(Batch *b, double no_match_penalty, i32 *rounds_out) {
    cudaStream_t s = b->stream;
    REQUIRE(b->stage >= 1, SAME_E_STATE, "same_batch_mip_start before same_batch_candidates");
    batch_settle(b);
    batch_kept_columns(b);
    const i64 P = b->P, nKA = b->nKA, nKR = b->nKR;
    b->start_x.alloc(P, s);
    b->start_unmatched.alloc(nKA, s);
    b->have_start = true;
    DevBuf<unsigned char> prefer, eligible, used;
    DevBuf<i32> nodes;
    prefer.alloc(nKA, s); eligible.alloc(P, s); used.alloc(nKA + nKR, s); nodes.alloc(2 * P, s);
    if (nKA > 0) LAUNCH(k_start_rows, blocks_for(nKA, 256), 256, 0, s, b->row_ptr.p, b->cost.p, b->ka_size.p, nKA, no_match_penalty, prefer.p);
    if (P > 0)
        LAUNCH(k_start_items, blocks_for(P, 256), 256, 0, s, b->pairs.p, P, b->d_p_off.p, b->d_ka_off.p, b->d_kr_off.p, (int)b->W, nKA, prefer.p,
               nodes.p, eligible.p);
    const int rounds = greedy_select_dev<2>(nodes.p, b->cost.p, eligible.p, P, nKA + nKR, b->start_x.p, used.p, s);
    if (nKA > 0) LAUNCH(k_start_unmatched, blocks_for(nKA, 256), 256, 0, s, used.p, nKA, b->start_unmatched.p);
    if (rounds_out) *rounds_out = rounds;
}

}  // namespace same
