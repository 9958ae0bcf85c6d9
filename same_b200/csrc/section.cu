// section.cu — frames resident in HBM, rectangle counting, window subsetting (src/same.py:293-295),
// vertex-id resolution for precomputed triangulations (src/same.py:262-290).
#include "common.cuh"

namespace same {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
bool g_prof = false;
bool g_debug = getenv("SAME_B200_DEBUG") != nullptr && getenv("SAME_B200_DEBUG")[0] == '1';
std::vector<ProfRec> g_prof_recs;

void exclusive_scan_i32(const i32 *in, i32 *out, i64 n, Scratch &sc, cudaStream_t s) {
    size_t bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, s));
    void *tmp = sc.get(bytes, s);
    ProfScope prof("cub::DeviceScan::ExclusiveSum", s);
    CK(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, s));
    g_launches.fetch_add(2, std::memory_order_relaxed);  // init + scan kernels
}

void upload_offsets(const std::vector<i64> &h, DevBuf<i32> &d, cudaStream_t s) {
    std::vector<i32> t(h.size());
    for (size_t i = 0; i < h.size(); ++i) t[i] = (i32)h[i];
    d.alloc((i64)t.size(), s);
    CK(cudaMemcpyAsync(d.p, t.data(), sizeof(i32) * t.size(), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));  // t goes out of scope
}

// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long enc_f64(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
static inline double dec_f64(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

__global__ void k_fill_f64(double *p, i64 n, double v) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_iota(i32 *p, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (i32)i;
}

// out[0..3] = enc(min x), enc(max x), enc(min y), enc(max y)
__global__ void k_bbox(const double2 *__restrict__ xy, i64 n, unsigned long long *out) {
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    double mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double2 p = xy[i];
        mnx = fmin(mnx, p.x); mxx = fmax(mxx, p.x); mny = fmin(mny, p.y); mxy = fmax(mxy, p.y);
    }
    double r;
    r = BR(tmp).Reduce(mnx, cub::Min()); __syncthreads(); if (threadIdx.x == 0) atomicMin(out + 0, enc_f64(r));
    r = BR(tmp).Reduce(mxx, cub::Max()); __syncthreads(); if (threadIdx.x == 0) atomicMax(out + 1, enc_f64(r));
    r = BR(tmp).Reduce(mny, cub::Min()); __syncthreads(); if (threadIdx.x == 0) atomicMin(out + 2, enc_f64(r));
    r = BR(tmp).Reduce(mxy, cub::Max()); __syncthreads(); if (threadIdx.x == 0) atomicMax(out + 3, enc_f64(r));
}

template <typename T>
static void upload(DevBuf<T> &d, const void *src, i64 count, cudaStream_t s) {
    d.alloc(count, s);
    if (count > 0) CK(cudaMemcpyAsync(d.p, src, sizeof(T) * (size_t)count, cudaMemcpyDefault, s));
}

void section_build(Section *sec, const double *a_xy, const double *r_xy, const double *a_prob, const double *r_prob,
                   const i32 *a_type, const i32 *r_type, const double *a_size, const double *r_size) {
    cudaStream_t s = sec->stream;
    const i64 nA = sec->nA, nR = sec->nR;
    const int K = sec->K;
    upload(sec->a_xy, a_xy, nA, s);
    upload(sec->r_xy, r_xy, nR, s);
    upload(sec->a_prob, a_prob, nA * K, s);
    upload(sec->r_prob, r_prob, nR * K, s);
    if (a_type) upload(sec->a_type, a_type, nA, s); else { sec->a_type.alloc(nA, s); sec->a_type.zero(s); }
    if (r_type) upload(sec->r_type, r_type, nR, s); else { sec->r_type.alloc(nR, s); sec->r_type.zero(s); }
    if (a_size) upload(sec->a_size, a_size, nA, s);
    else { sec->a_size.alloc(nA, s); LAUNCH(k_fill_f64, blocks_for(nA, 256), 256, 0, s, sec->a_size.p, nA, 1.0); }
    if (r_size) upload(sec->r_size, r_size, nR, s);
    else { sec->r_size.alloc(nR, s); LAUNCH(k_fill_f64, blocks_for(nR, 256), 256, 0, s, sec->r_size.p, nR, 1.0); }
    // bounding box of both frames (src/same.py:481-482)
    DevBuf<unsigned long long> bb;
    bb.alloc(4, s);
    const unsigned long long init[4] = {~0ull, 0ull, ~0ull, 0ull};
    CK(cudaMemcpyAsync(bb.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
    if (nA > 0) LAUNCH(k_bbox, std::min<unsigned>(blocks_for(nA, 256), 1184), 256, 0, s, sec->a_xy.p, nA, bb.p);
    if (nR > 0) LAUNCH(k_bbox, std::min<unsigned>(blocks_for(nR, 256), 1184), 256, 0, s, sec->r_xy.p, nR, bb.p);
    unsigned long long h[4];
    CK(cudaMemcpyAsync(h, bb.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int i = 0; i < 4; ++i) sec->bbox[i] = (nA + nR > 0) ? dec_f64(h[i]) : 0.0;
}

// ---- rectangle counts ----------------------------------------------------------------
__device__ __forceinline__ bool in_rect(double2 p, const double *__restrict__ r) {
    return p.x >= r[0] && p.x < r[1] && p.y >= r[2] && p.y < r[3];
}

// grid = (chunks, M); each block counts the cells of its chunk that fall in rect m
constexpr int SUB_THREADS = 256;
constexpr int SUB_ITEMS = 8;
constexpr int SUB_CHUNK = SUB_THREADS * SUB_ITEMS;

__global__ void __launch_bounds__(SUB_THREADS) k_rect_count(const double2 *__restrict__ xy, i64 n, const double *__restrict__ rects,
                                                            i32 *__restrict__ block_counts, unsigned long long *totals) {
    const double *r = rects + 4 * (i64)blockIdx.y;
    const i64 base = (i64)blockIdx.x * SUB_CHUNK;
    int c = 0;
#pragma unroll
    for (int it = 0; it < SUB_ITEMS; ++it) {
        const i64 i = base + it * SUB_THREADS + threadIdx.x;
        c += (i < n) && in_rect(xy[i], r);
    }
    typedef cub::BlockReduce<int, SUB_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const int tot = BR(tmp).Sum(c);
    if (threadIdx.x == 0) {
        if (block_counts) block_counts[(i64)blockIdx.y * gridDim.x + blockIdx.x] = tot;
        if (totals && tot) atomicAdd(totals + blockIdx.y, (unsigned long long)tot);
    }
}

// second pass: ranks inside the block in ascending row order -> src[base + rank] = row
__global__ void __launch_bounds__(SUB_THREADS) k_rect_fill(const double2 *__restrict__ xy, i64 n, const double *__restrict__ rects,
                                                           const i32 *__restrict__ block_base, i32 *__restrict__ src) {
    const double *r = rects + 4 * (i64)blockIdx.y;
    const i64 base = (i64)blockIdx.x * SUB_CHUNK;
    typedef cub::BlockScan<int, SUB_THREADS> BS;
    __shared__ typename BS::TempStorage tmp;
    int run = block_base[(i64)blockIdx.y * gridDim.x + blockIdx.x];
#pragma unroll 1
    for (int it = 0; it < SUB_ITEMS; ++it) {
        const i64 i = base + it * SUB_THREADS + threadIdx.x;
        const int f = (i < n) && in_rect(xy[i], r);
        int rank, tot;
        BS(tmp).ExclusiveSum(f, rank, tot);
        if (f) src[run + rank] = (i32)i;
        run += tot;
        __syncthreads();
    }
}

__global__ void k_pick_offsets(const i32 *__restrict__ scanned, i64 chunks, i64 W, i32 *__restrict__ off) {
    i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w <= W) off[w] = scanned[w * chunks];
}

void section_count_rects(Section *sec, i64 m, const double *rects, i64 *cntA, i64 *cntR) {
    cudaStream_t s = sec->stream;
    DevBuf<double> d_r;
    upload(d_r, rects, 4 * m, s);
    DevBuf<unsigned long long> tot;
    tot.alloc(2 * m, s);
    tot.zero(s);
    if (sec->nA > 0) LAUNCH(k_rect_count, dim3(blocks_for(sec->nA, SUB_CHUNK), (unsigned)m), SUB_THREADS, 0, s, sec->a_xy.p, sec->nA, d_r.p, (i32 *)nullptr, tot.p);
    if (sec->nR > 0) LAUNCH(k_rect_count, dim3(blocks_for(sec->nR, SUB_CHUNK), (unsigned)m), SUB_THREADS, 0, s, sec->r_xy.p, sec->nR, d_r.p, (i32 *)nullptr, tot.p + m);
    std::vector<unsigned long long> h(2 * m);
    CK(cudaMemcpyAsync(h.data(), tot.p, sizeof(unsigned long long) * 2 * m, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (i64 i = 0; i < m; ++i) { cntA[i] = (i64)h[i]; cntR[i] = (i64)h[m + i]; }
}

// ---- window subsetting: subset_data (src/same.py:293-295) for all windows at once ---------------------------
// One thread per cell looks up the rectangles overlapping its index-grid cell, tests them exactly (half-open) and
// appends (window << rowbits | row); a key-only radix sort then yields, per window, the rows in ascending order —
// exactly `df[mask]` of the reference.  Lanes walk their candidate lists in lockstep so a warp needs one atomic per slot.
// pass 1: number of windows holding each cell; pass 2 (after an exclusive scan): keys written at scan[i].. in ascending
// window order.  The key stream is then ordered by (row, window), so ONE stable radix pass over the window bits
// yields (window, row) order — no atomics, no multi-pass sort.
template <bool FILL, typename KeyT>
__global__ void __launch_bounds__(256) k_subset_scan(const double2 *__restrict__ xy, i64 n, const double *__restrict__ rects, RectIndexDev ri,
                                                     int rowbits, i32 *__restrict__ count, const i32 *__restrict__ pos, KeyT *__restrict__ keys) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { if (!FILL) count[i] = 0; return; }
    const double2 p = xy[i];
    const int c = rect_cell(ri, p.x, p.y);
    const i32 lo = ri.cell_ptr[c], hi = ri.cell_ptr[c + 1];
    i32 out = FILL ? pos[i] : 0;
    for (i32 k = lo; k < hi; ++k) {
        const int w = ri.cell_rects[k];
        if (in_rect(p, rects + 4 * (i64)w)) {
            if (FILL) keys[out] = (KeyT)(((unsigned long long)w << rowbits) | (unsigned long long)i);
            ++out;
        }
    }
    if (!FILL) count[i] = out;
}

template <typename KeyT>
__global__ void k_subset_finish(const KeyT *__restrict__ sorted, i64 total, int rowbits, i64 W, i32 *__restrict__ src, i32 *__restrict__ off) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) src[i] = (i32)((unsigned long long)sorted[i] & ((1ull << rowbits) - 1ull));
    if (i <= W) {  // offset of window i = lower bound of key (i << rowbits)
        const unsigned long long key = (unsigned long long)i << rowbits;
        i64 lo = 0, hi = total;
        while (lo < hi) {
            const i64 mid = (lo + hi) >> 1;
            if ((unsigned long long)sorted[mid] < key) lo = mid + 1; else hi = mid;
        }
        off[i] = (i32)lo;
    }
}

static int bits_needed(i64 n) {
    int b = 1;
    while ((1ll << b) < n) ++b;
    return b;
}

template <typename KeyT>
static void subset_frame_t(Batch *b, const DevBuf<double2> &xy, i64 n, std::vector<i64> &off, DevBuf<i32> &d_off, DevBuf<i32> &src, i64 &total,
                           int rowbits, int wbits) {
    cudaStream_t s = b->stream;
    const i64 W = b->W;
    DevBuf<i32> count, pos;
    DevBuf<KeyT> keys, keys_out;
    count.alloc(n + 1, s); pos.alloc(n + 1, s);
    LAUNCH((k_subset_scan<false, KeyT>), blocks_for(n + 1, 256), 256, 0, s, xy.p, n, b->d_rects.p, b->rindex, rowbits, count.p, (const i32 *)nullptr,
           (KeyT *)nullptr);
    exclusive_scan_i32(count.p, pos.p, n + 1, b->scratch, s);
    i32 h = 0;
    CK(cudaMemcpyAsync(&h, pos.p + n, sizeof(h), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    total = (i64)h;
    REQUIRE(total >= 0, SAME_E_LIMIT, "batch exceeds 2^31 window instances");
    src.alloc(total, s);
    d_off.alloc(W + 1, s);
    keys.alloc(total, s); keys_out.alloc(total, s);
    if (total > 0) {
        LAUNCH((k_subset_scan<true, KeyT>), blocks_for(n + 1, 256), 256, 0, s, xy.p, n, b->d_rects.p, b->rindex, rowbits, (i32 *)nullptr, pos.p, keys.p);
        size_t bytes = 0;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys.p, keys_out.p, (int)total, rowbits, rowbits + wbits, s));
        void *tmp = b->scratch.get(bytes, s);
        {
            ProfScope prof("cub::DeviceRadixSort::SortKeys(subset)", s);
            CK(cub::DeviceRadixSort::SortKeys(tmp, bytes, keys.p, keys_out.p, (int)total, rowbits, rowbits + wbits, s));
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    LAUNCH((k_subset_finish<KeyT>), blocks_for(std::max<i64>(total, W + 1), 256), 256, 0, s, keys_out.p, total, rowbits, W, src.p, d_off.p);
    std::vector<i32> ho(W + 1);
    CK(cudaMemcpyAsync(ho.data(), d_off.p, sizeof(i32) * (W + 1), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    off.assign(ho.begin(), ho.end());
}

static void subset_frame(Batch *b, const DevBuf<double2> &xy, i64 n, std::vector<i64> &off, DevBuf<i32> &d_off, DevBuf<i32> &src, i64 &total) {
    const int rowbits = bits_needed(std::max<i64>(n, 2)), wbits = bits_needed(b->W + 1);
    if (rowbits + wbits <= 32) subset_frame_t<unsigned>(b, xy, n, off, d_off, src, total, rowbits, wbits);
    else subset_frame_t<unsigned long long>(b, xy, n, off, d_off, src, total, rowbits, wbits);
}

// host: uniform index grid over the section bbox, each rectangle registered in every grid cell it overlaps (+1 cell of slack)
static void build_rect_index(Batch *b) {
    Section *sec = b->sec;
    cudaStream_t s = b->stream;
    const i64 W = b->W;
    const double bx0 = sec->bbox[0], bx1 = sec->bbox[1], by0 = sec->bbox[2], by1 = sec->bbox[3];
    const double ext = std::max(std::max(bx1 - bx0, by1 - by0), 1e-300);
    int G = (int)std::min<i64>(256, std::max<i64>(1, (i64)std::ceil(std::sqrt((double)W)) * 4));
    std::vector<i32> ptr, lst;
    double cs = 1.0, inv = 1.0;
    int nx = 1, ny = 1, max_len = 0;
    for (;; G = std::max(1, G / 2)) {
        cs = ext / G;
        if (!(cs > 0.0) || !std::isfinite(cs)) cs = 1.0;
        inv = 1.0 / cs;
        nx = (int)std::floor((bx1 - bx0) * inv) + 1;
        ny = (int)std::floor((by1 - by0) * inv) + 1;
        auto cell = [&](double v, double o, int nmax) {
            double t = std::floor((v - o) * inv);
            if (!(t > -1e9)) t = -1e9;   // -inf / NaN
            if (t > 1e9) t = 1e9;
            return (int)std::min<double>(std::max<double>(t, 0.0), (double)(nmax - 1));
        };
        std::vector<i32> cnt((size_t)nx * ny + 1, 0);
        i64 total = 0;
        std::vector<int> cx0(W), cx1(W), cy0(W), cy1(W);
        for (i64 w = 0; w < W; ++w) {
            const double *r = &b->rects[4 * w];
            cx0[w] = std::max(0, cell(r[0], bx0, nx) - 1); cx1[w] = std::min(nx - 1, cell(r[1], bx0, nx) + 1);
            cy0[w] = std::max(0, cell(r[2], by0, ny) - 1); cy1[w] = std::min(ny - 1, cell(r[3], by0, ny) + 1);
            if (r[1] < bx0 || r[0] > bx1 || r[3] < by0 || r[2] > by1) { cx1[w] = cx0[w] - 1; continue; }  // misses the section
            total += (i64)(cx1[w] - cx0[w] + 1) * (cy1[w] - cy0[w] + 1);
        }
        if (total > (1ll << 24) && G > 1) continue;  // too fine for this many rectangles: coarsen
        for (i64 w = 0; w < W; ++w)
            for (int cy = cy0[w]; cy <= cy1[w] && cx1[w] >= cx0[w]; ++cy)
                for (int cx = cx0[w]; cx <= cx1[w]; ++cx) cnt[(size_t)cy * nx + cx + 1]++;
        for (size_t c = 0; c < (size_t)nx * ny; ++c) { max_len = std::max(max_len, cnt[c + 1]); cnt[c + 1] += cnt[c]; }
        ptr = cnt;
        lst.assign((size_t)std::max<i64>(total, 1), 0);
        std::vector<i32> fill(ptr.begin(), ptr.end() - 1);
        for (i64 w = 0; w < W; ++w)   // ascending w inside every cell
            for (int cy = cy0[w]; cy <= cy1[w] && cx1[w] >= cx0[w]; ++cy)
                for (int cx = cx0[w]; cx <= cx1[w]; ++cx) lst[fill[(size_t)cy * nx + cx]++] = (i32)w;
        break;
    }
    b->ri_ptr.alloc((i64)ptr.size(), s);
    b->ri_rects.alloc((i64)lst.size(), s);
    CK(cudaMemcpyAsync(b->ri_ptr.p, ptr.data(), sizeof(i32) * ptr.size(), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b->ri_rects.p, lst.data(), sizeof(i32) * lst.size(), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    b->rindex = RectIndexDev{bx0, by0, inv, nx, ny, max_len, b->ri_ptr.p, b->ri_rects.p};
}

void batch_subset(Batch *b) {
    Section *sec = b->sec;
    cudaStream_t s = b->stream;
    b->d_rects.alloc(4 * b->W, s);
    CK(cudaMemcpyAsync(b->d_rects.p, b->rects.data(), sizeof(double) * 4 * b->W, cudaMemcpyHostToDevice, s));
    build_rect_index(b);
    subset_frame(b, sec->a_xy, sec->nA, b->a_off, b->d_a_off, b->a_src, b->nAi);
    subset_frame(b, sec->r_xy, sec->nR, b->r_off, b->d_r_off, b->r_src, b->nRi);
}

// ---- vertex ids -> section rows ------------------------------------------------------
__global__ void k_iota64(i64 *p, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// id_to_row = {v: i for i, v in enumerate(vertex_ids)} -> later rows win (src/same.py:285)
__global__ void k_resolve_vids(const i64 *__restrict__ sorted_vid, const i32 *__restrict__ sorted_row, i64 n,
                               const i64 *__restrict__ tri_vid, i64 m, i32 *__restrict__ out_rows) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const i64 v = tri_vid[e];
    i64 lo = 0, hi = n;  // upper bound
    while (lo < hi) {
        const i64 mid = (lo + hi) >> 1;
        if (sorted_vid[mid] <= v) lo = mid + 1; else hi = mid;
    }
    out_rows[e] = (lo > 0 && sorted_vid[lo - 1] == v) ? sorted_row[lo - 1] : -1;
}

__global__ void k_resolve_identity(const i64 *__restrict__ tri_vid, i64 m, i64 n, i32 *__restrict__ out_rows) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const i64 v = tri_vid[e];
    out_rows[e] = (v >= 0 && v < n) ? (i32)v : -1;
}

void section_set_triangles(Section *sec, const i64 *a_vid, const i64 *tri_vid, i64 n_tri) {
    cudaStream_t s = sec->stream;
    const i64 n = sec->nA, m = 3 * n_tri;
    DevBuf<i64> d_tri;
    upload(d_tri, tri_vid, m, s);
    sec->tri_rows.alloc(m, s);
    sec->Tg = n_tri;
    if (m == 0) { CK(cudaStreamSynchronize(s)); return; }
    if (!a_vid) {
        LAUNCH(k_resolve_identity, blocks_for(m, 256), 256, 0, s, d_tri.p, m, n, sec->tri_rows.p);
    } else {
        DevBuf<i64> vid_in, vid_out;
        DevBuf<i32> row_in, row_out;
        upload(vid_in, a_vid, n, s);
        vid_out.alloc(n, s); row_in.alloc(n, s); row_out.alloc(n, s);
        LAUNCH(k_iota, blocks_for(n, 256), 256, 0, s, row_in.p, n);
        size_t bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, vid_in.p, vid_out.p, row_in.p, row_out.p, (int)n, 0, 64, s));
        void *tmp = sec->scratch.get(bytes, s);
        {
            ProfScope prof("cub::DeviceRadixSort::SortPairs(vid)", s);
            CK(cub::DeviceRadixSort::SortPairs(tmp, bytes, vid_in.p, vid_out.p, row_in.p, row_out.p, (int)n, 0, 64, s));
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        LAUNCH(k_resolve_vids, blocks_for(m, 256), 256, 0, s, vid_out.p, row_out.p, n, d_tri.p, m, sec->tri_rows.p);
    }
    CK(cudaStreamSynchronize(s));
}

}  // namespace same
