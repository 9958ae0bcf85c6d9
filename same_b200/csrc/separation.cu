// separation.cu — a10 lazy-constraint separation (src/same.py:621-703), a11 x/y-order diagnostics
// (src/violationhelper.py:1-134), a12 signed-area flip analysis (src/same.py:1355-1408, src/helpers.py:73-77).
//
// One thread per Delaunay triangle.  The orientation test reproduces the reference's naive fp64
// expression one IEEE operation at a time (no FMA), so the violated set is bit-identical to the
// Python callback; an exact predicate would change the set and is deliberately not used.
#include "common.cuh"

namespace same {

// matching[i] = j of the LAST pair of row i with x > 0.5 (dict overwrite, src/same.py:636-639).  A block owns
// MATCH_ROWS consecutive rows; their pairs are one contiguous range of x, which the block reads coalesced and reduces
// to one flag per pair in shared memory; every thread then scans the flags of its own row.
constexpr int MATCH_ROWS = 256, MATCH_CHUNK = 2048;
__global__ void __launch_bounds__(MATCH_ROWS) k_match_rows(const double *__restrict__ x, i32 x_base, const int2 *__restrict__ pairs,
                                                           const i32 *__restrict__ row_ptr, const i32 *__restrict__ ka_off, const i32 *__restrict__ p_off,
                                                           int W, i32 k_lo, i32 k_hi, i32 *__restrict__ match_j, i32 *__restrict__ match_p) {
    __shared__ unsigned char flag[MATCH_CHUNK];
    __shared__ i32 range[2];
    const i32 k0 = k_lo + (i32)blockIdx.x * MATCH_ROWS;
    const i32 k = k0 + threadIdx.x;
    if (threadIdx.x == 0) { range[0] = row_ptr[k0]; range[1] = row_ptr[min(k0 + MATCH_ROWS, k_hi)]; }
    i32 rs = 0, re = 0;
    if (k < k_hi) { rs = row_ptr[k]; re = row_ptr[k + 1]; }
    __syncthreads();
    const i32 pb = range[0], pe = range[1];
    i32 mp = -1;
    for (i32 c = pb; c < pe; c += MATCH_CHUNK) {
        for (i32 q = threadIdx.x; q < MATCH_CHUNK; q += MATCH_ROWS) flag[q] = (c + q < pe) && (x[c + q - x_base] > 0.5);
        __syncthreads();
        const i32 lo = max(rs, c), hi = min(re, c + MATCH_CHUNK);
        for (i32 p = lo; p < hi; ++p)
            if (flag[p - c]) mp = p;
        __syncthreads();
    }
    if (k >= k_hi) return;
    i32 mj = -1;
    if (mp >= 0) { mj = pairs[mp].y; mp -= p_off[find_window(ka_off, W, k)]; }
    match_j[k] = mj;
    match_p[k] = mp;
}

// Separation of one incumbent in ONE launch: orientation test per triangle, ranks of the violated triangles inside
// their window (exclusive scan that restarts at every window, scan.cuh), the first `cap` cuts of every window in
// ascending triangle order and the per-window counts.  A block owns SEP_TILE consecutive triangles; the geometry
// phase is striped (coalesced, independent gathers), the scan phase blocked.
constexpr int SEP_THREADS = 256, SEP_ITEMS = 4, SEP_TILE = SEP_THREADS * SEP_ITEMS;   // 4 or 8 (flags of a thread are read as one word)

__global__ void __launch_bounds__(SEP_THREADS) k_separation(const int3 *__restrict__ tri, const signed char *__restrict__ src_sign, i32 t_lo, i32 t_hi,
                                                            const i32 *__restrict__ t_off, const i32 *__restrict__ ka_off, const i32 *__restrict__ kr_off,
                                                            int W, int w_lo, const i32 *__restrict__ match_j, const i32 *__restrict__ match_p,
                                                            const double2 *__restrict__ kr_xy, i64 cap, ScanCtx sc,
                                                            i32 *__restrict__ counts /* [2*nw] zeroed: n_viol, n_checked */, i32 *__restrict__ cuts,
                                                            i32 *__restrict__ unc_list, i32 *__restrict__ unc_count) {
    __shared__ int E[SEP_TILE + 1];
    __shared__ __align__(8) unsigned char V[SEP_TILE];
    __shared__ int smem[2 * (SEP_THREADS / 32) + 2];   // two scanned streams (violated, checked) + the carry word
    const i32 t0 = t_lo + (i32)blockIdx.x * SEP_TILE;
    const i32 tend = min(t0 + SEP_TILE, t_hi);
    const int wf = find_window(t_off, W, t0), wl = find_window(t_off, W, tend - 1);
    int wk[SEP_ITEMS];
    int ck = 0;
#pragma unroll
    for (int k = 0; k < SEP_ITEMS; ++k) {
        const int pos = k * SEP_THREADS + threadIdx.x;
        const i32 t = t0 + pos;
        int w = wf;
        unsigned char vi = 0;
        if (t < tend) {
            if (wf != wl) w = find_window(t_off, W, t);
            const i32 nb = ka_off[w], rb = kr_off[w];
            const int3 v = tri[t];
            const i32 ja = match_j[nb + v.x], jb = match_j[nb + v.y], jc = match_j[nb + v.z];
            if (ja >= 0 && jb >= 0 && jc >= 0) {  // same.py:649-651
                const double2 A = kr_xy[rb + ja], B = kr_xy[rb + jb], C = kr_xy[rb + jc];
                const int rs = sign_of(orient_naive(A.x, A.y, B.x, B.y, C.x, C.y));  // same.py:658
                if (orient_uncertain(A.x, A.y, B.x, B.y, C.x, C.y)) {   // diagnostic only: listed for the host's exact check
                    const i32 at = atomicAdd(unc_count, 1);
                    if (at < UNC_CAP) unc_list[at] = t;
                }
                const int ss = src_sign[t];
                if (ss != 0 && rs != 0) {  // same.py:663-664
                    if (wf == wl) ++ck; else atomicAdd(counts + 2 * (w - w_lo) + 1, 1);
                    vi = ss != rs;  // same.py:669
                }
            }
        }
        wk[k] = w;
        V[pos] = vi;
    }
    __syncthreads();
    // blocked phase: thread owns flags [8*tid, 8*tid+8)
    static_assert(SEP_ITEMS == 4 || SEP_ITEMS == 8, "flags of a thread are read as one 32- or 64-bit word");
    const unsigned long long bits = SEP_ITEMS == 8 ? *reinterpret_cast<const unsigned long long *>(V + threadIdx.x * SEP_ITEMS)
                                                   : (unsigned long long)*reinterpret_cast<const unsigned *>(V + threadIdx.x * SEP_ITEMS);
    int nv[2] = {__popcll(bits), ck}, excl[2], tot[2];
    block_exclusive_scan<2, SEP_THREADS>(nv, excl, tot, smem);
    {
        int run = excl[0];
#pragma unroll
        for (int k = 0; k < SEP_ITEMS; ++k) {
            E[threadIdx.x * SEP_ITEMS + k] = run;
            run += (int)((bits >> (8 * k)) & 1ull);
        }
        if (threadIdx.x == SEP_THREADS - 1) E[SEP_TILE] = run;
    }
    if (wf == wl && threadIdx.x == 0 && tot[1]) atomicAdd(counts + 2 * (wf - w_lo) + 1, tot[1]);
    __syncthreads();
    // the tile's last window: does it start at or after the tile's first triangle?  Then later tiles need nothing older.
    const i32 wl_start = t_off[wl];
    const bool has_boundary = wl_start >= t0;
    const int tail = tot[0] - (has_boundary ? E[wl_start - t0] : 0);
    const int carry = tile_segmented_carry(sc, (int)blockIdx.x, tot[0], has_boundary, tail, smem);
#pragma unroll
    for (int k = 0; k < SEP_ITEMS; ++k) {
        const int pos = k * SEP_THREADS + threadIdx.x;
        const i32 t = t0 + pos;
        if (t >= tend) continue;
        const int ww = wk[k];
        const i32 ws = t_off[ww];
        const int rank = ws >= t0 ? E[pos] - E[ws - t0] : carry + E[pos];
        const bool vi = V[pos] != 0;
        if (vi && rank < cap) {
            const i32 nb = ka_off[ww];
            const int3 v = tri[t];
            reinterpret_cast<int4 *>(cuts)[(i64)(ww - w_lo) * cap + rank] = make_int4(match_p[nb + v.x], match_p[nb + v.y], match_p[nb + v.z], t - ws);
        }
        if (t == t_off[ww + 1] - 1) counts[2 * (ww - w_lo)] = rank + (vi ? 1 : 0);
    }
}

// x may be host or device memory; a device-resident solution vector is used in place
static const double *resolve_x(Batch *b, i64 w_lo, i64 w_hi, const double *x) {
    cudaStream_t s = b->stream;
    const i64 n = b->p_off[w_hi] - b->p_off[w_lo];
    if (n == 0) return x;
    REQUIRE(x != nullptr, SAME_E_ARG, "x is NULL");
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, x) == cudaSuccess && attr.type == cudaMemoryTypeDevice && attr.device == b->sec->device) return x;
    cudaGetLastError();  // clear "invalid value" for plain host pointers on older drivers
    b->x_dev.alloc(n, s);
    CK(cudaMemcpyAsync(b->x_dev.p, x, sizeof(double) * (size_t)n, cudaMemcpyDefault, s));
    return b->x_dev.p;
}

static void run_matching(Batch *b, i64 w_lo, i64 w_hi, const double *xd) {
    cudaStream_t s = b->stream;
    b->match_j.alloc(b->nKA, s);
    b->match_p.alloc(b->nKA, s);
    const i32 k_lo = (i32)b->ka_off[w_lo], k_hi = (i32)b->ka_off[w_hi];
    if (k_hi > k_lo)
        LAUNCH(k_match_rows, blocks_for(k_hi - k_lo, MATCH_ROWS), MATCH_ROWS, 0, s, xd, (i32)b->p_off[w_lo], b->pairs.p, b->row_ptr.p, b->d_ka_off.p,
               b->d_p_off.p, (int)b->W, k_lo, k_hi, b->match_j.p, b->match_p.p);
}

void batch_separation(Batch *b, i64 w_lo, i64 w_hi, const double *x, i64 cap, i64 *n_viol, i64 *n_checked, i32 *cuts) {
    cudaStream_t s = b->stream;
    REQUIRE(b->stage >= 4, SAME_E_STATE, "same_batch_separation before same_batch_tri_finalize");
    REQUIRE(w_lo >= 0 && w_hi <= b->W && w_lo < w_hi, SAME_E_ARG, "bad window range");
    REQUIRE(cap >= 0, SAME_E_ARG, "cap must be >= 0");
    batch_settle(b);
    batch_kept_columns(b);
    const i64 nw = w_hi - w_lo;
    run_matching(b, w_lo, w_hi, resolve_x(b, w_lo, w_hi, x));
    const i32 t_lo = (i32)b->t_off[w_lo], t_hi = (i32)b->t_off[w_hi];
    const i64 nt = t_hi - t_lo;
    b->sep_counts.alloc(2 * nw + 1, s);   // + the count of orientations the filter could not certify (diagnostic)
    CK(cudaMemsetAsync(b->sep_counts.p, 0, sizeof(i32) * (2 * nw + 1), s));
    // Where the cuts go: straight into the caller's block when the GPU can address it — page-locked host memory (the kernel's
    // stores cross PCIe as posted writes, no copy-engine launch behind the kernel) or device memory — else into a device buffer
    // that is copied afterwards.
    i32 *cut_dst = nullptr;
    bool direct = false;
    if (cuts && cap > 0 && nt > 0) {
        cudaPointerAttributes attr;   // asked on every call (about a microsecond): a recycled address may have changed its kind
        if (cudaPointerGetAttributes(&attr, cuts) == cudaSuccess &&
            ((attr.type == cudaMemoryTypeHost && attr.devicePointer) || (attr.type == cudaMemoryTypeDevice && attr.device == b->sec->device)))
            cut_dst = (i32 *)attr.devicePointer;
        cudaGetLastError();
        direct = cut_dst != nullptr;
    }
    if (!direct) {
        b->cuts.alloc(std::max<i64>(1, nw * cap * 4), s);
        cut_dst = b->cuts.p;
    }
    if (nt > 0) {
        const unsigned tiles = blocks_for(nt, SEP_TILE);
        LAUNCH(k_separation, tiles, SEP_THREADS, 0, s, b->tri.p, b->t_sign.p, t_lo, t_hi, b->d_t_off.p, b->d_ka_off.p, b->d_kr_off.p, (int)b->W, (int)w_lo,
               b->match_j.p, b->match_p.p, b->kr_xy.p, cap, scan_ctx(b->sec, tiles, 1, s), b->sep_counts.p, cut_dst, b->unc_list[1].p,
               b->sep_counts.p + 2 * nw);
    }
    // counts (and the cut block, unless the kernel wrote it in place) come back behind ONE synchronisation
    i32 *h_sep = b->pin_misc();   // page-locked (2 nw + 1 <= 4(W+1) + 8 ints)
    small_d2h(h_sep, b->sep_counts.p, sizeof(i32) * (2 * nw + 1), s);   // (the diagnostic count rides along: no extra round trip)
    if (!direct && cuts && cap > 0 && nt > 0) CK(cudaMemcpyAsync(cuts, b->cuts.p, sizeof(i32) * 4 * (size_t)(nw * cap), cudaMemcpyDefault, s));
    batch_sync(b);
    for (i64 w = 0; w < nw; ++w) {
        n_viol[w] = h_sep[2 * w];
        n_checked[w] = h_sep[2 * w + 1];
    }
    b->last_unc_sep = h_sep[2 * nw];
}

// triangles whose naive orientation sign the static filter could not certify: which = 0 source signs (aligned coordinates,
// src/same.py:1146), 1 = the last separation call (matched reference coordinates, src/same.py:658).  Batch-global TRI indices.
void batch_uncertain(Batch *b, int which, i64 cap, i64 *n, i32 *tri_idx) {
    cudaStream_t s = b->stream;
    REQUIRE(b->stage >= 4, SAME_E_STATE, "no triangle tables yet");
    REQUIRE(which == 0 || which == 1, SAME_E_ARG, "which must be 0 or 1");
    REQUIRE(n && cap >= 0 && (cap == 0 || tri_idx), SAME_E_ARG, "bad output");
    batch_settle(b);
    if (which == 1) {
        *n = std::max<i64>(b->last_unc_sep, 0);       // came back with the counts of the separation call (0 before the first one)
    } else {
        i32 *h = b->pin_misc();
        small_d2h(h, b->unc_count.p + which, sizeof(i32), s);
        CK(stream_wait(s));
        *n = h[0];
    }
    const i64 m = std::min<i64>(std::min<i64>(*n, UNC_CAP), cap);
    if (m > 0) {
        CK(cudaMemcpyAsync(tri_idx, b->unc_list[which].p, sizeof(i32) * (size_t)m, cudaMemcpyDefault, s));
        CK(stream_wait(s));
    }
}

// ---- a11 + a12 --------------------------------------------------------------------------------------
// 0.5 * (x1*(y2-y3) + x2*(y3-y1) + x3*(y1-y2)), left to right (src/helpers.py:77)
__device__ __forceinline__ double signed_area(double2 p1, double2 p2, double2 p3) {
    const double a = __dmul_rn(p1.x, __dsub_rn(p2.y, p3.y));
    const double b = __dmul_rn(p2.x, __dsub_rn(p3.y, p1.y));
    const double c = __dmul_rn(p3.x, __dsub_rn(p1.y, p2.y));
    return __dmul_rn(0.5, __dadd_rn(__dadd_rn(a, b), c));
}

__global__ void k_postsolve(const int3 *__restrict__ tri, i32 t_lo, i32 t_hi, const i32 *__restrict__ t_off, const i32 *__restrict__ ka_off,
                            const i32 *__restrict__ kr_off, int W, const i32 *__restrict__ match_j, const double2 *__restrict__ ka_xy,
                            const double2 *__restrict__ kr_xy, i32 *__restrict__ mask, double *__restrict__ area_before,
                            double *__restrict__ area_after, unsigned char *__restrict__ flipped) {
    const i32 t = t_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_hi) return;
    const int w = find_window(t_off, W, t);
    const i32 nb = ka_off[w], rb = kr_off[w];
    const int3 v3 = tri[t];
    const i32 v[3] = {v3.x, v3.y, v3.z};
    i32 j[3];
    double2 a[3], r[3];
    i32 m = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        j[c] = match_j[nb + v[c]];
        a[c] = ka_xy[nb + v[c]];
        if (j[c] >= 0) { m |= 1 << (8 + c); r[c] = kr_xy[rb + j[c]]; } else r[c] = make_double2(0.0, 0.0);
    }
    const int PU[3] = {0, 0, 1}, PW[3] = {1, 2, 2};
#pragma unroll
    for (int q = 0; q < 3; ++q) {  // violationhelper.py:62-75
        const int u = PU[q], ww = PW[q];
        if (j[u] < 0 || j[ww] < 0) continue;
        if ((a[u].x < a[ww].x) != (r[u].x < r[ww].x)) m |= 1 << q;
        if ((a[u].y < a[ww].y) != (r[u].y < r[ww].y)) m |= 1 << (3 + q);
    }
    mask[t] = m;
    const double ab = signed_area(a[0], a[1], a[2]);
    area_before[t] = ab;
    if (j[0] >= 0 && j[1] >= 0 && j[2] >= 0) {
        const double aa = signed_area(r[0], r[1], r[2]);
        area_after[t] = aa;
        flipped[t] = __dmul_rn(ab, aa) < 0.0;  // same.py:1398
    } else {
        area_after[t] = __longlong_as_double(0x7ff8000000000000ll);
        flipped[t] = 0;
    }
}

void batch_postsolve(Batch *b, i64 w_lo, i64 w_hi, const double *x) {
    cudaStream_t s = b->stream;
    REQUIRE(b->stage >= 4, SAME_E_STATE, "same_batch_postsolve before same_batch_tri_finalize");
    REQUIRE(w_lo >= 0 && w_hi <= b->W && w_lo < w_hi, SAME_E_ARG, "bad window range");
    batch_settle(b);
    batch_kept_columns(b);
    const double *xd = resolve_x(b, w_lo, w_hi, x);
    run_matching(b, w_lo, w_hi, xd);
    if (!b->have_post) {
        b->t_mask.alloc(b->T, s); b->area_before.alloc(b->T, s); b->area_after.alloc(b->T, s); b->flipped.alloc(b->T, s);
        b->have_post = true;
    }
    const i32 t_lo = (i32)b->t_off[w_lo], t_hi = (i32)b->t_off[w_hi];
    if (t_hi > t_lo)
        LAUNCH(k_postsolve, blocks_for(t_hi - t_lo, 256), 256, 0, s, b->tri.p, t_lo, t_hi, b->d_t_off.p, b->d_ka_off.p, b->d_kr_off.p, (int)b->W,
               b->match_j.p, b->ka_xy.p, b->kr_xy.p, b->t_mask.p, b->area_before.p, b->area_after.p, b->flipped.p);
    if (xd != x) batch_sync(b);   // the caller's host vector was read asynchronously; a device-resident x needs no wait
}

void postsolve_arrays(int device, i64 T, const i32 *tri, i64 nA, const double *a_xy, i64 nR, const double *r_xy, const i32 *match_j, i32 *mask,
                      double *area_before, double *area_after, unsigned char *flipped) {
    CK(cudaSetDevice(device));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    try {
        DevBuf<int3> d_tri; DevBuf<double2> d_a, d_r; DevBuf<i32> d_m, d_off, d_mask; DevBuf<double> d_ab, d_aa; DevBuf<unsigned char> d_fl;
        d_tri.alloc(T, s); d_a.alloc(nA, s); d_r.alloc(nR, s); d_m.alloc(nA, s); d_off.alloc(6, s);
        d_mask.alloc(T, s); d_ab.alloc(T, s); d_aa.alloc(T, s); d_fl.alloc(T, s);
        if (T) CK(cudaMemcpyAsync(d_tri.p, tri, sizeof(int3) * T, cudaMemcpyDefault, s));
        if (nA) CK(cudaMemcpyAsync(d_a.p, a_xy, sizeof(double2) * nA, cudaMemcpyDefault, s));
        if (nR) CK(cudaMemcpyAsync(d_r.p, r_xy, sizeof(double2) * nR, cudaMemcpyDefault, s));
        if (nA) CK(cudaMemcpyAsync(d_m.p, match_j, sizeof(i32) * nA, cudaMemcpyDefault, s));
        const i32 off[6] = {0, (i32)T, 0, (i32)nA, 0, (i32)nR};
        CK(cudaMemcpyAsync(d_off.p, off, sizeof(off), cudaMemcpyHostToDevice, s));
        if (T > 0)
            LAUNCH(k_postsolve, blocks_for(T, 256), 256, 0, s, d_tri.p, 0, (i32)T, d_off.p, d_off.p + 2, d_off.p + 4, 1, d_m.p, d_a.p, d_r.p, d_mask.p,
                   d_ab.p, d_aa.p, d_fl.p);
        if (T) {
            CK(cudaMemcpyAsync(mask, d_mask.p, sizeof(i32) * T, cudaMemcpyDefault, s));
            CK(cudaMemcpyAsync(area_before, d_ab.p, sizeof(double) * T, cudaMemcpyDefault, s));
            CK(cudaMemcpyAsync(area_after, d_aa.p, sizeof(double) * T, cudaMemcpyDefault, s));
            CK(cudaMemcpyAsync(flipped, d_fl.p, T, cudaMemcpyDefault, s));
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

}  // namespace same
