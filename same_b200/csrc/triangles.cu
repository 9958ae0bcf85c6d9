// triangles.cu — a5 _remap_triangles_by_vertex_ids (src/same.py:262-290), a6 filter_triangles_by_radius
// (src/helpers.py:233-395), a7 unconstrained-node removal (src/same.py:1056-1085), a8 triangle info
// (src/helpers.py:184-210), a9 weights + source signs (src/same.py:1128-1146), batched over windows.
#include "common.cuh"

namespace same {


// Remap = for every global triangle, every window that holds all three vertices AND whose post-KNN frame kept all
// three rows.  The subset stage left, per section row, the list of (window, instance) it belongs to (row_inst, ascending
// window, a handful of entries), and the candidate stage the kept index of every instance (newA, valid where cnt > 0):
// a thread intersects the three short lists of its triangle — no rectangle tests, no binary searches.  Pass 1 counts the
// hits per triangle and scans the counts in the same launch (scan.cuh); pass 2 writes (window << tbits | triangle) keys
// and the window-local vertex triples at pos[t]..; one stable radix pass over the window bits restores the reference's
// order (window-major, input order inside) and the window offsets are lower bounds in the sorted keys.
__device__ __forceinline__ i32 inst_in_window(const int2 *__restrict__ row_inst, i32 lo, i32 hi, i32 w) {
    for (i32 e = lo; e < hi; ++e) {
        const int2 v = row_inst[e];
        if (v.x == w) return v.y;
        if (v.x > w) break;
    }
    return -1;
}

constexpr int REMAP_CT = 1024, REMAP_CI = 2;   // counting pass: big tiles for the look-back chain

// Row table: the (window, window-local kept index or -1) entries of every aligned row, at most ROW_TAB of them, in ONE 32-byte
// sector per row — a triangle then costs three sector gathers instead of walking row_pos -> row_inst -> cnt / newA for each
// vertex (five dependent gathers per vertex).  Unused entries hold window TAB_END (above every window, the entries ascend); a
// row that lies in more than ROW_TAB windows (overlap above one half) is marked TAB_MANY and its triangles take the general walk.
constexpr int ROW_TAB = 4;
constexpr i32 TAB_END = 0x7fffffff, TAB_MANY = -2;
__global__ void k_row_table(i64 nA, const i32 *__restrict__ row_pos, const int2 *__restrict__ row_inst, const i32 *__restrict__ cnt,
                            const i32 *__restrict__ newA, const i32 *__restrict__ ka_off, int4 *__restrict__ tab) {
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nA) return;
    const i32 e0 = row_pos[r], n = row_pos[r + 1] - e0;
    i32 w[ROW_TAB], l[ROW_TAB];
#pragma unroll
    for (int k = 0; k < ROW_TAB; ++k) {
        w[k] = TAB_END; l[k] = -1;
        if (k < n) {
            const int2 v = row_inst[e0 + k];
            w[k] = v.x;
            l[k] = cnt[v.y] > 0 ? newA[v.y] - ka_off[v.x] : -1;
        }
    }
    if (n > ROW_TAB) w[0] = TAB_MANY;
    tab[2 * r] = make_int4(w[0], l[0], w[1], l[1]);
    tab[2 * r + 1] = make_int4(w[2], l[2], w[3], l[3]);
}
__device__ __forceinline__ i32 tab_find(const int4 &p, const int4 &q, i32 w) {   // local index of the row in window w, -1 if none
    return p.x == w ? p.y : (p.z == w ? p.w : (q.x == w ? q.y : (q.z == w ? q.w : -1)));
}

// visits the hits of triangle t in ascending window order: f(window, local a, local b, local c)
template <typename F>
__device__ __forceinline__ void remap_hits_walk(i32 ra, i32 rb, i32 rc, const i32 *__restrict__ row_pos, const int2 *__restrict__ row_inst,
                                                const i32 *__restrict__ cnt, const i32 *__restrict__ newA, const i32 *__restrict__ ka_off, F &&f) {
    const i32 a0 = row_pos[ra], a1 = row_pos[ra + 1];
    if (a0 == a1) return;
    const i32 b0 = row_pos[rb], b1 = row_pos[rb + 1], c0 = row_pos[rc], c1 = row_pos[rc + 1];
    for (i32 e = a0; e < a1; ++e) {
        const int2 va = row_inst[e];
        if (cnt[va.y] == 0) continue;
        const i32 ib = inst_in_window(row_inst, b0, b1, va.x);
        if (ib < 0 || cnt[ib] == 0) continue;
        const i32 ic = inst_in_window(row_inst, c0, c1, va.x);
        if (ic < 0 || cnt[ic] == 0) continue;
        const i32 kb = ka_off[va.x];
        f(va.x, newA[va.y] - kb, newA[ib] - kb, newA[ic] - kb);
    }
}
template <typename F>
__device__ __forceinline__ void remap_hits(const i32 *__restrict__ tri_rows, i64 t, const int4 *__restrict__ tab, const i32 *__restrict__ row_pos,
                                           const int2 *__restrict__ row_inst, const i32 *__restrict__ cnt, const i32 *__restrict__ newA,
                                           const i32 *__restrict__ ka_off, F &&f) {
    const i32 ra = tri_rows[3 * t], rb = tri_rows[3 * t + 1], rc = tri_rows[3 * t + 2];
    if (ra < 0 || rb < 0 || rc < 0) return;
    const int4 a0 = tab[2 * ra], a1 = tab[2 * ra + 1], b0 = tab[2 * rb], b1 = tab[2 * rb + 1], c0 = tab[2 * rc], c1 = tab[2 * rc + 1];
    if (a0.x == TAB_MANY || b0.x == TAB_MANY || c0.x == TAB_MANY) {
        remap_hits_walk(ra, rb, rc, row_pos, row_inst, cnt, newA, ka_off, f);
        return;
    }
    const i32 wa[ROW_TAB] = {a0.x, a0.z, a1.x, a1.z}, la[ROW_TAB] = {a0.y, a0.w, a1.y, a1.w};
#pragma unroll
    for (int k = 0; k < ROW_TAB; ++k) {
        if (wa[k] == TAB_END || la[k] < 0) continue;     // (an END entry has local -1 as well)
        const i32 lb = tab_find(b0, b1, wa[k]);
        if (lb < 0) continue;
        const i32 lc = tab_find(c0, c1, wa[k]);
        if (lc < 0) continue;
        f(wa[k], la[k], lb, lc);
    }
}

__global__ void __launch_bounds__(REMAP_CT) k_remap_count(const i32 *__restrict__ tri_rows, i64 Tg, const int4 *__restrict__ tab,
                                                          const i32 *__restrict__ row_pos, const int2 *__restrict__ row_inst, const i32 *__restrict__ cnt, const i32 *__restrict__ newA,
                                                          const i32 *__restrict__ ka_off, ScanCtx sc, i32 *__restrict__ pos,
                                                          int4 *__restrict__ first_hit) {
    __shared__ int smem[REMAP_CT / 32 + 1];
    const i64 base = ((i64)blockIdx.x * REMAP_CT + threadIdx.x) * REMAP_CI;
    int c[REMAP_CI], sum[1] = {0};
#pragma unroll
    for (int k = 0; k < REMAP_CI; ++k) {
        c[k] = 0;
        // most triangles lie in exactly one window: their only hit is remembered so that the fill pass does not walk the
        // row -> instance table a second time
        if (base + k < Tg) {
            int4 h = make_int4(0, 0, 0, 0);
            remap_hits(tri_rows, base + k, tab, row_pos, row_inst, cnt, newA, ka_off, [&](i32 w, i32 la, i32 lb, i32 lc) {
                if (c[k] == 0) h = make_int4(w, la, lb, lc);
                ++c[k];
            });
            if (c[k] == 1) first_hit[base + k] = h;
        }
        sum[0] += c[k];
    }
    int excl[1], tot[1], pre[1];
    device_exclusive_scan<1, REMAP_CT>(sc, (int)blockIdx.x, sum, excl, tot, pre, smem);
    int run = excl[0];
#pragma unroll
    for (int k = 0; k < REMAP_CI; ++k) {
        if (base + k <= Tg) pos[base + k] = run;
        run += c[k];
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(256) k_remap_fill(const i32 *__restrict__ tri_rows, i64 Tg, const int4 *__restrict__ tab,
                                                    const i32 *__restrict__ row_pos, const int2 *__restrict__ row_inst, const i32 *__restrict__ cnt, const i32 *__restrict__ newA,
                                                    const i32 *__restrict__ ka_off, const i32 *__restrict__ pos, const int4 *__restrict__ first_hit,
                                                    int tbits, KeyT *__restrict__ keys, i32 *__restrict__ idx, int3 *__restrict__ recs) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tg) return;
    i32 out = pos[t];
    const i32 n = pos[t + 1] - out;
    if (n == 0) return;
    if (n == 1) {   // the single hit the counting pass remembered
        const int4 h = first_hit[t];
        keys[out] = (KeyT)(((unsigned long long)h.x << tbits) | (unsigned long long)t);
        idx[out] = out;
        recs[out] = make_int3(h.y, h.z, h.w);
        return;
    }
    remap_hits(tri_rows, t, tab, row_pos, row_inst, cnt, newA, ka_off, [&](i32 w, i32 la, i32 lb, i32 lc) {
        keys[out] = (KeyT)(((unsigned long long)w << tbits) | (unsigned long long)t);
        idx[out] = out;
        recs[out] = make_int3(la, lb, lc);
        ++out;
    });
}

template <typename KeyT>
__global__ void k_remap_gather(const KeyT *__restrict__ sorted_keys, const i32 *__restrict__ sorted_idx,
                               const int3 *__restrict__ recs, i64 n, int tbits, i64 W, int3 *__restrict__ tin, i32 *__restrict__ tin_src,
                               i32 *__restrict__ tin_off) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        tin[i] = recs[sorted_idx[i]];
        tin_src[i] = (i32)((unsigned long long)sorted_keys[i] & ((1ull << tbits) - 1ull));
    }
    if (i <= W) {
        const unsigned long long key = (unsigned long long)i << tbits;
        i64 lo = 0, hi = n;
        while (lo < hi) {
            const i64 mid = (lo + hi) >> 1;
            if ((unsigned long long)sorted_keys[mid] < key) lo = mid + 1; else hi = mid;
        }
        tin_off[i] = (i32)lo;
    }
}

static void reset_triangles(Batch *b) {
    b->Tin = 0; b->T = 0; b->n_band = 0; b->tin_has_src = false; b->have_post = false;
}

template <typename KeyT>
static void remap_t(Batch *b, int tbits, int wbits) {
    Section *sec = b->sec;
    cudaStream_t s = b->stream;
    const i64 W = b->W, Tg = sec->Tg;
    DevBuf<i32> pos;
    DevBuf<int4> first_hit;
    DevBuf<int4> tab;
    pos.alloc(Tg + 1, s);
    first_hit.alloc(Tg, s);
    tab.alloc(2 * sec->nA, s);
    if (sec->nA > 0) LAUNCH(k_row_table, blocks_for(sec->nA, 256), 256, 0, s, sec->nA, b->row_pos.p, b->row_inst.p, b->cnt.p, b->newA.p, b->d_ka_off.p, tab.p);
    const unsigned tiles = blocks_for(Tg + 1, REMAP_CT * REMAP_CI);
    LAUNCH(k_remap_count, tiles, REMAP_CT, 0, s, sec->tri_rows.p, Tg, tab.p, b->row_pos.p, b->row_inst.p, b->cnt.p, b->newA.p, b->d_ka_off.p,
           scan_ctx(sec, tiles, 1, s), pos.p, first_hit.p);
    small_d2h(b->pin_misc(), pos.p + Tg, sizeof(i32), s);
    batch_sync(b);   // also brings in the candidate stage's window offsets
    b->Tin = (i64)b->pin_misc()[0];
    REQUIRE(b->Tin >= 0, SAME_E_LIMIT, "too many window triangles");
    DevBuf<KeyT> keys, keys_out;
    DevBuf<int3> recs;
    DevBuf<i32> idx, idx_out;
    b->tin.alloc(b->Tin, s); b->tin_src.alloc(b->Tin, s);
    b->d_tin_off.alloc(W + 1, s);
    keys.alloc(b->Tin, s); recs.alloc(b->Tin, s); idx.alloc(b->Tin, s);
    const KeyT *sorted_keys = keys.p;
    const i32 *sorted_idx = idx.p;
    if (b->Tin > 0) {
        LAUNCH((k_remap_fill<KeyT>), blocks_for(Tg, 256), 256, 0, s, sec->tri_rows.p, Tg, tab.p, b->row_pos.p, b->row_inst.p, b->cnt.p, b->newA.p, b->d_ka_off.p,
               pos.p, first_hit.p, tbits, keys.p, idx.p, recs.p);
        if (W > 1) {   // a single window is already in input order
            keys_out.alloc(b->Tin, s); idx_out.alloc(b->Tin, s);
            size_t bytes = 0;
            CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys_out.p, idx.p, idx_out.p, (int)b->Tin, tbits, tbits + wbits, s));
            void *tmp = b->scratch.get(bytes, s);
            {
                ProfScope prof("cub::DeviceRadixSort::SortPairs(remap)", s);
                CK(cub::DeviceRadixSort::SortPairs(tmp, bytes, keys.p, keys_out.p, idx.p, idx_out.p, (int)b->Tin, tbits, tbits + wbits, s));
            }
            g_launches.fetch_add(1, std::memory_order_relaxed);
            sorted_keys = keys_out.p;
            sorted_idx = idx_out.p;
        }
    }
    LAUNCH((k_remap_gather<KeyT>), blocks_for(std::max<i64>(b->Tin, W + 1), 256), 256, 0, s, sorted_keys, sorted_idx, recs.p, b->Tin, tbits, W, b->tin.p,
           b->tin_src.p, b->d_tin_off.p);
    small_d2h(b->pin_tin(), b->d_tin_off.p, sizeof(i32) * (W + 1), s);
    b->pend_tin = true;
}

void batch_triangles_remap(Batch *b) {
    Section *sec = b->sec;
    REQUIRE(b->stage >= 1, SAME_E_STATE, "same_batch_triangles_remap before same_batch_candidates");
    REQUIRE(sec->Tg >= 0, SAME_E_STATE, "same_section_set_triangles was not called");
    reset_triangles(b);
    int tbits = 1, wbits = 1;
    while ((1ll << tbits) < sec->Tg) ++tbits;
    while ((1ll << wbits) < b->W + 1) ++wbits;
    if (tbits + wbits <= 32) remap_t<unsigned>(b, tbits, wbits);
    else remap_t<unsigned long long>(b, tbits, wbits);
    b->tin_has_src = true;
    b->stage = 2;
}

void batch_triangles_set(Batch *b, const i32 *tri, const i64 *tri_off) {
    cudaStream_t s = b->stream;
    REQUIRE(b->stage >= 1, SAME_E_STATE, "same_batch_triangles_set before same_batch_candidates");
    batch_settle(b);
    reset_triangles(b);
    b->pend_tin = false;
    const i64 W = b->W;
    b->tin_off.assign(tri_off, tri_off + W + 1);
    REQUIRE(b->tin_off[0] == 0, SAME_E_ARG, "tri_off[0] must be 0");
    for (i64 w = 0; w < W; ++w) REQUIRE(b->tin_off[w + 1] >= b->tin_off[w], SAME_E_ARG, "tri_off must be non-decreasing");
    b->Tin = b->tin_off[W];
    b->tin.alloc(b->Tin, s);
    if (b->Tin > 0) CK(cudaMemcpyAsync(b->tin.p, tri, sizeof(int3) * (size_t)b->Tin, cudaMemcpyDefault, s));
    upload_offsets(b->tin_off, b->d_tin_off, s);
    CK(stream_wait(s));   // the caller's triangle array (possibly page-locked) was read asynchronously
    b->stage = 2;
}

// ---- a6 step 1: classification ---------------------------------------------------------------
// 1-D np.linalg.norm / np.dot go through BLAS ddot == fma(y1,y2,x1*x2)
__device__ __forceinline__ double norm2_blas(double x, double y) { return __dsqrt_rn(__fma_rn(y, y, __dmul_rn(x, x))); }
// Clipped cosine of the angle between v1 and v2 as compute_angle forms it (src/helpers.py:278-288: dot / (|v1| |v2|), np.clip),
// from the dot product and the two norms; a zero-length side gives angle 0 there, i.e. cosine 1 here.
__device__ __forceinline__ double angle_cos(double dot, double n1, double n2) {
    if (n1 == 0.0 || n2 == 0.0) return 1.0;
    return fmin(fmax(__ddiv_rn(dot, __dmul_rn(n1, n2)), -1.0), 1.0);
}

__global__ void __launch_bounds__(256, 4) k_tri_classify(const int3 *__restrict__ tin, i64 Tin, const i32 *__restrict__ tin_off, const i32 *__restrict__ ka_off, int W,
                               const double2 *__restrict__ ka_xy, const i32 *__restrict__ ka_type, double radius, int use_angle,
                               double min_angle, int ignore_same_type, unsigned char *__restrict__ cls, double *__restrict__ score,
                               i32 *__restrict__ band_idx, i32 *__restrict__ band_count) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tin) return;
    const int w = find_window(tin_off, W, (i32)t);
    const i32 nb = ka_off[w];
    const int3 v = tin[t];
    const double2 p1 = ka_xy[nb + v.x], p2 = ka_xy[nb + v.y], p3 = ka_xy[nb + v.z];
    const double s1 = norm2_blas(__dsub_rn(p2.x, p1.x), __dsub_rn(p2.y, p1.y));
    const double s2 = norm2_blas(__dsub_rn(p3.x, p2.x), __dsub_rn(p3.y, p2.y));
    const double s3 = norm2_blas(__dsub_rn(p1.x, p3.x), __dsub_rn(p1.y, p3.y));
    const double mx = fmax(s1, fmax(s2, s3));
    bool band = fabs(mx - radius) <= 1e-12 * fabs(radius);
    score[t] = __dadd_rn(__dadd_rn(s1, s2), s3);  // helpers.py:333
    unsigned char k;
    if (mx >= radius) k = SAME_TRI_DROP_RADIUS;  // helpers.py:310
    else {
        k = SAME_TRI_KEEP;
        if (use_angle) {  // helpers.py:315-321
            // The three angles share the three edge vectors e1 = p2 - p1, e2 = p3 - p2, e3 = p1 - p3 whose norms are the sides
            // above: compute_angle's vectors at a vertex are one edge and the NEGATED other one, negation is exact in every
            // operation involved (x*x, fma, division), so dot_at_vertex = -fma(ey, e'y, ex * e'x) bit for bit.  Only the smallest
            // angle is compared: min_i acos(c_i) = acos(max_i c_i) (acos decreases; a last-place wobble of the device acos
            // falls inside the guard band below, which the host re-decides with numpy).  One acos, three square roots.
            const double e1x = __dsub_rn(p2.x, p1.x), e1y = __dsub_rn(p2.y, p1.y), e2x = __dsub_rn(p3.x, p2.x), e2y = __dsub_rn(p3.y, p2.y);
            const double e3x = __dsub_rn(p1.x, p3.x), e3y = __dsub_rn(p1.y, p3.y);
            const double c1 = angle_cos(-__fma_rn(e1y, e3y, __dmul_rn(e1x, e3x)), s1, s3);   // at p1: v1 = e1, v2 = -e3
            const double c2 = angle_cos(-__fma_rn(e1y, e2y, __dmul_rn(e1x, e2x)), s1, s2);   // at p2: v1 = -e1, v2 = e2
            const double c3 = angle_cos(-__fma_rn(e3y, e2y, __dmul_rn(e3x, e2x)), s3, s2);   // at p3: v1 = e3, v2 = -e2
            const double mn = __dmul_rn(acos(fmax(c1, fmax(c2, c3))), 57.29577951308232);  // np.degrees: x * (180/pi)
            if (fabs(mn - min_angle) <= 1e-9) band = true;
            if (mn < min_angle) k = SAME_TRI_DROP_ANGLE;
        }
        if (k == SAME_TRI_KEEP && ignore_same_type) {  // helpers.py:328-330
            const i32 ta = ka_type[nb + v.x], tb = ka_type[nb + v.y], tc = ka_type[nb + v.z];
            if (ta == tb && tb == tc) k = SAME_TRI_SAME_TYPE;
        }
    }
    cls[t] = k;
    if (band) band_idx[atomicAdd(band_count, 1)] = (i32)t;
}

void batch_tri_classify(Batch *b, double radius, int use_angle, double min_angle_deg, int ignore_same_type) {
    cudaStream_t s = b->stream;
    REQUIRE(b->stage >= 2, SAME_E_STATE, "same_batch_tri_classify before triangles were provided");
    const i64 Tin = b->Tin;
    batch_kept_columns(b);   // (no synchronisation: the kernel takes the kept counts from device memory)
    b->cls.alloc(Tin, s); b->score.alloc(Tin, s); b->band_idx.alloc(Tin + 1, s);
    DevBuf<i32> bc;
    bc.alloc(1, s);
    bc.zero(s);
    if (Tin > 0)
        LAUNCH(k_tri_classify, blocks_for(Tin, 256), 256, 0, s, b->tin.p, Tin, b->d_tin_off.p, b->d_ka_off.p, (int)b->W, b->ka_xy.p, b->ka_type.p,
               radius, use_angle, min_angle_deg, ignore_same_type, b->cls.p, b->score.p, b->band_idx.p, bc.p);
    small_d2h(b->pin_misc(), bc.p, sizeof(i32), s);
    batch_sync(b);   // also brings in the remap's window offsets
    b->n_band = b->pin_misc()[0];
    b->stage = 3;
}

__global__ void k_tri_override(i64 n, const i32 *__restrict__ idx, const unsigned char *__restrict__ v, i64 Tin, unsigned char *__restrict__ cls) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && idx[i] >= 0 && idx[i] < Tin) cls[idx[i]] = v[i];
}

void batch_tri_override(Batch *b, i64 n, const i32 *idx, const unsigned char *cls) {
    cudaStream_t s = b->stream;
    REQUIRE(b->stage >= 3, SAME_E_STATE, "same_batch_tri_override before same_batch_tri_classify");
    if (n == 0) return;
    DevBuf<i32> d_i;
    DevBuf<unsigned char> d_c;
    d_i.alloc(n, s); d_c.alloc(n, s);
    CK(cudaMemcpyAsync(d_i.p, idx, sizeof(i32) * n, cudaMemcpyDefault, s));
    CK(cudaMemcpyAsync(d_c.p, cls, n, cudaMemcpyDefault, s));
    LAUNCH(k_tri_override, blocks_for(n, 256), 256, 0, s, n, d_i.p, d_c.p, b->Tin, b->cls.p);
    CK(stream_wait(s));
}

// ---- a6 step 2 -----------------------------------------------------------------------------------
__global__ void k_tri_nodes(const int3 *__restrict__ tin, i64 Tin, const i32 *__restrict__ tin_off, const i32 *__restrict__ ka_off, int W,
                            const unsigned char *__restrict__ cls, const double *__restrict__ score, int track_best,
                            i32 *__restrict__ node_valid, i32 *__restrict__ has_tri, unsigned long long *__restrict__ best_score,
                            i32 *__restrict__ kept_flag) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == Tin) kept_flag[t] = 0;
    if (t >= Tin) return;
    const unsigned char c = cls[t];
    kept_flag[t] = (c == SAME_TRI_KEEP);
    if (c < SAME_TRI_SAME_TYPE) return;
    const i32 nb = ka_off[find_window(tin_off, W, (i32)t)];
    const int3 v = tin[t];
    node_valid[nb + v.x] = 1; node_valid[nb + v.y] = 1; node_valid[nb + v.z] = 1;  // helpers.py:324-325
    if (c == SAME_TRI_KEEP) { has_tri[nb + v.x] = 1; has_tri[nb + v.y] = 1; has_tri[nb + v.z] = 1; }
    else if (track_best) {
        const unsigned long long k = (unsigned long long)__double_as_longlong(score[t]);  // score >= 0: bit order == value order
        atomicMin(best_score + nb + v.x, k); atomicMin(best_score + nb + v.y, k); atomicMin(best_score + nb + v.z, k);
    }
}
// strict '<' + first wins (helpers.py:335-339) == smallest triangle index among the minimal scores
__global__ void k_tri_best(const int3 *__restrict__ tin, i64 Tin, const i32 *__restrict__ tin_off, const i32 *__restrict__ ka_off, int W,
                           const unsigned char *__restrict__ cls, const double *__restrict__ score,
                           const unsigned long long *__restrict__ best_score, i32 *__restrict__ best_tri) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tin || cls[t] != SAME_TRI_SAME_TYPE) return;
    const i32 nb = ka_off[find_window(tin_off, W, (i32)t)];
    const int3 v = tin[t];
    const unsigned long long k = (unsigned long long)__double_as_longlong(score[t]);
    if (best_score[nb + v.x] == k) atomicMin(best_tri + nb + v.x, (i32)t);
    if (best_score[nb + v.y] == k) atomicMin(best_tri + nb + v.y, (i32)t);
    if (best_score[nb + v.z] == k) atomicMin(best_tri + nb + v.z, (i32)t);
}
// node v re-adds its best same-type triangle unless a smaller uncovered node already did (helpers.py:365-383)
// Add-back decision per node fused with the three node-indexed prefix sums of the stage (add-back triangles, unconstrained
// nodes, surviving nodes) — one launch with a three-stream in-kernel scan (scan.cuh) instead of a flag kernel and three
// device-wide scans.  A thread owns AB_ITEMS consecutive nodes; item nKA is the sentinel that receives the totals.
constexpr int AB_THREADS = 256, AB_ITEMS = 4;
__global__ void __launch_bounds__(AB_THREADS) k_addback_scan(i64 nKA, const i32 *__restrict__ ka_off, int W, const i32 *__restrict__ node_valid,
                                                             const i32 *__restrict__ has_tri, const i32 *__restrict__ best_tri,
                                                             const int3 *__restrict__ tin, int enabled, ScanCtx sc, i32 *__restrict__ ab_flag,
                                                             i32 *__restrict__ unc_flag, i32 *__restrict__ abpos, i32 *__restrict__ uncpos,
                                                             i32 *__restrict__ validpos) {
    __shared__ int smem[3 * (AB_THREADS / 32) + 3];
    const i64 base = ((i64)blockIdx.x * AB_THREADS + threadIdx.x) * AB_ITEMS;
    int fa[AB_ITEMS], fu[AB_ITEMS], fv[AB_ITEMS], sum[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < AB_ITEMS; ++k) {
        const i64 v = base + k;
        fa[k] = fu[k] = fv[k] = 0;
        if (v < nKA) {
            fv[k] = node_valid[v] != 0;
            fu[k] = !fv[k];
            if (enabled && fv[k] && !has_tri[v] && best_tri[v] != 0x7fffffff) {
                const i32 t = best_tri[v];
                const i32 nb = ka_off[find_window(ka_off, W, (i32)v)];
                const int3 tv = tin[t];
                int f = 1;
                const i32 u3[3] = {nb + tv.x, nb + tv.y, nb + tv.z};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const i32 u = u3[c];
                    if (u < (i32)v && !has_tri[u] && best_tri[u] == t) f = 0;
                }
                fa[k] = f;
            }
        }
        sum[0] += fa[k]; sum[1] += fu[k]; sum[2] += fv[k];
    }
    int excl[3], tot[3], pre[3];
    device_exclusive_scan<3, AB_THREADS>(sc, (int)blockIdx.x, sum, excl, tot, pre, smem);
    int ra = excl[0], ru = excl[1], rv = excl[2];
#pragma unroll
    for (int k = 0; k < AB_ITEMS; ++k) {
        const i64 v = base + k;
        if (v <= nKA) {
            ab_flag[v] = fa[k]; unc_flag[v] = fu[k];
            abpos[v] = ra; uncpos[v] = ru; validpos[v] = rv;
        }
        ra += fa[k]; ru += fu[k]; rv += fv[k];
    }
}
__global__ void k_tri_window_offsets(const i32 *__restrict__ kpos, const i32 *__restrict__ abpos, const i32 *__restrict__ uncpos,
                                     const i32 *__restrict__ validpos, const i32 *__restrict__ tin_off, const i32 *__restrict__ ka_off, int W,
                                     i32 *__restrict__ out /* 4*(W+1): t_off, nkept(per window), unc_off, new ka_off */) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > W) return;
    out[w] = kpos[tin_off[w]] + abpos[ka_off[w]];
    out[(W + 1) + w] = w < W ? kpos[tin_off[w + 1]] - kpos[tin_off[w]] : 0;
    out[2 * (W + 1) + w] = uncpos[ka_off[w]];
    out[3 * (W + 1) + w] = validpos[ka_off[w]];
}

__global__ void k_tri_emit_kept(const int3 *__restrict__ tin, i64 Tin, const i32 *__restrict__ tin_off, const i32 *__restrict__ ka_off, int W,
                                const i32 *__restrict__ kept_flag, const i32 *__restrict__ kpos, const i32 *__restrict__ t_off,
                                const i32 *__restrict__ validpos, int renumber, int3 *__restrict__ tri, i32 *__restrict__ tri_src) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tin || !kept_flag[t]) return;
    const int w = find_window(tin_off, W, (i32)t);
    const i32 pos = t_off[w] + (kpos[t] - kpos[tin_off[w]]);
    int3 v = tin[t];
    if (renumber) {
        const i32 nb = ka_off[w], vb = validpos[nb];
        v.x = validpos[nb + v.x] - vb; v.y = validpos[nb + v.y] - vb; v.z = validpos[nb + v.z] - vb;
    }
    tri[pos] = v;
    tri_src[pos] = (i32)t - tin_off[w];
}
__global__ void k_tri_emit_addback(i64 nKA, const i32 *__restrict__ ka_off, int W, const i32 *__restrict__ ab_flag, const i32 *__restrict__ abpos,
                                   const i32 *__restrict__ best_tri, const int3 *__restrict__ tin, const i32 *__restrict__ tin_off,
                                   const i32 *__restrict__ t_off, const i32 *__restrict__ nkept, const i32 *__restrict__ validpos, int renumber,
                                   int3 *__restrict__ tri, i32 *__restrict__ tri_src) {
    const i64 n = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nKA || !ab_flag[n]) return;
    const int w = find_window(ka_off, W, (i32)n);
    const i32 nb = ka_off[w];
    const i32 pos = t_off[w] + nkept[w] + (abpos[n] - abpos[nb]);
    const i32 t = best_tri[n];
    int3 v = tin[t];
    if (renumber) {
        const i32 vb = validpos[nb];
        v.x = validpos[nb + v.x] - vb; v.y = validpos[nb + v.y] - vb; v.z = validpos[nb + v.z] - vb;
    }
    tri[pos] = v;
    tri_src[pos] = t - tin_off[w];
}
__global__ void k_emit_unconstrained(i64 nKA, const i32 *__restrict__ ka_off, int W, const i32 *__restrict__ unc_flag,
                                     const i32 *__restrict__ uncpos, i32 *__restrict__ unc) {
    const i64 n = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nKA || !unc_flag[n]) return;
    unc[uncpos[n]] = (i32)n - ka_off[find_window(ka_off, W, (i32)n)];
}

// ---- a7: drop unconstrained nodes, renumber rows and pairs (src/same.py:1056-1085) ---------------------
__global__ void k_compact_nodes(i64 nKA, const i32 *__restrict__ node_valid, const i32 *__restrict__ validpos, const i32 *__restrict__ keepA,
                                const double2 *__restrict__ xy, const i32 *__restrict__ type, const double *__restrict__ size,
                                i32 *__restrict__ keepA2, double2 *__restrict__ xy2, i32 *__restrict__ type2, double *__restrict__ size2) {
    const i64 n = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nKA || !node_valid[n]) return;
    const i32 k = validpos[n];
    keepA2[k] = keepA[n]; xy2[k] = xy[n]; type2[k] = type[n]; size2[k] = size[n];
}
// Pairs of removed nodes go, the others are renumbered (src/same.py:1066-1070) — flag, scan and write in ONE launch
// (scan.cuh).  A block owns CP_TILE consecutive pairs (pair P is a sentinel that receives the totals): striped,
// coalesced passes compute the keep flags and later move the pairs; the scan in between is blocked over the flags in
// shared memory.  The thread that holds the first pair of a window / of a row also writes the window's new pair offset /
// the row's new row pointer.
constexpr int CP_THREADS = 256, CP_ITEMS = 8, CP_TILE = CP_THREADS * CP_ITEMS;
__global__ void __launch_bounds__(CP_THREADS) k_compact_pairs(const int2 *__restrict__ pairs, const double *__restrict__ cost, i64 P,
                                                              const i32 *__restrict__ p_off, const i32 *__restrict__ ka_off, int W,
                                                              const i32 *__restrict__ node_valid, const i32 *__restrict__ validpos,
                                                              const i32 *__restrict__ row_ptr, i64 nKA, ScanCtx sc, int2 *__restrict__ pairs2,
                                                              double *__restrict__ cost2, i32 *__restrict__ row_ptr2, i32 *__restrict__ poff2) {
    __shared__ int E[CP_TILE];
    __shared__ __align__(8) unsigned char V[CP_TILE];
    __shared__ int smem[CP_THREADS / 32 + 1];
    const i64 p0 = (i64)blockIdx.x * CP_TILE;
    const i64 pend = min(p0 + CP_TILE, P);   // real pairs of this tile: [p0, pend)
    const int wf = p0 < P ? find_window(p_off, W, (i32)p0) : 0, wl = pend > p0 ? find_window(p_off, W, (i32)(pend - 1)) : wf;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
        const int pos = k * CP_THREADS + threadIdx.x;
        const i64 p = p0 + pos;
        unsigned char f = 0;
        if (p < pend) {
            const int w = wf == wl ? wf : find_window(p_off, W, (i32)p);
            f = node_valid[ka_off[w] + pairs[p].x] != 0;
        }
        V[pos] = f;
    }
    __syncthreads();
    const unsigned long long bits = *reinterpret_cast<const unsigned long long *>(V + threadIdx.x * CP_ITEMS);
    int sum[1] = {__popcll(bits)}, excl[1], tot[1], pre[1];
    device_exclusive_scan<1, CP_THREADS>(sc, (int)blockIdx.x, sum, excl, tot, pre, smem);
    {
        int run = excl[0];
#pragma unroll
        for (int k = 0; k < CP_ITEMS; ++k) {
            E[threadIdx.x * CP_ITEMS + k] = run;
            run += (int)((bits >> (8 * k)) & 1ull);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
        const int pos = k * CP_THREADS + threadIdx.x;
        const i64 p = p0 + pos;
        if (p > P) continue;
        if (p == P) {   // totals: windows that end the pair list, and the row pointer past the last kept row
            const i32 end = pos > 0 ? E[pos - 1] + V[pos - 1] : pre[0];   // pairs kept in total
            for (int ww = W; ww >= 0 && p_off[ww] == (i32)P; --ww) poff2[ww] = end;
            row_ptr2[validpos[nKA]] = end;
            continue;
        }
        const i32 dst = E[pos];
        const int w = wf == wl ? wf : find_window(p_off, W, (i32)p);
        if (p == p_off[w])
            for (int ww = w; ww >= 0 && p_off[ww] == (i32)p; --ww) poff2[ww] = dst;   // (empty windows share the start)
        if (V[pos]) {
            const i32 nb = ka_off[w];
            const int2 q = pairs[p];
            const i32 row = nb + q.x;
            pairs2[dst] = make_int2(validpos[row] - validpos[nb], q.y);
            cost2[dst] = cost[p];
            if (p == row_ptr[row]) row_ptr2[validpos[row]] = dst;
        }
    }
}
// keep the instance -> kept-index map of the remap in step with the renumbered nodes
__global__ void k_renumber_instances(i64 nAi, const i32 *__restrict__ node_valid, const i32 *__restrict__ validpos, i32 *__restrict__ cnt,
                                     i32 *__restrict__ newA) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nAi || cnt[i] == 0) return;
    const i32 k = newA[i];
    if (node_valid[k]) newA[i] = validpos[k]; else cnt[i] = 0;
}
__global__ void k_fill_i32_tri(i32 *p, i64 n, i32 v) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- a8 + a9 ------------------------------------------------------------------------------------------
__global__ void k_tri_tables(const int3 *__restrict__ tri, i64 T, const i32 *__restrict__ t_off, const i32 *__restrict__ ka_off, int W,
                             const double2 *__restrict__ ka_xy, const double *__restrict__ ka_size, double *__restrict__ weight,
                             signed char *__restrict__ sign, double *__restrict__ bounds, i32 *__restrict__ argv, i32 *__restrict__ unc_list,
                             i32 *__restrict__ unc_count) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const i32 nb = ka_off[find_window(t_off, W, (i32)t)];
    const int3 v = tri[t];
    const i32 vv[3] = {v.x, v.y, v.z};
    const double2 p[3] = {ka_xy[nb + v.x], ka_xy[nb + v.y], ka_xy[nb + v.z]};
    weight[t] = __dadd_rn(__dadd_rn(ka_size[nb + v.x], ka_size[nb + v.y]), ka_size[nb + v.z]);  // same.py:1131-1134
    sign[t] = (signed char)sign_of(orient_naive(p[0].x, p[0].y, p[1].x, p[1].y, p[2].x, p[2].y));  // same.py:1146
    if (orient_uncertain(p[0].x, p[0].y, p[1].x, p[1].y, p[2].x, p[2].y)) {   // diagnostic only
        const i32 at = atomicAdd(unc_count, 1);
        if (at < UNC_CAP) unc_list[at] = (i32)t;
    }
    const double mnx = fmin(p[0].x, fmin(p[1].x, p[2].x)), mxx = fmax(p[0].x, fmax(p[1].x, p[2].x));
    const double mny = fmin(p[0].y, fmin(p[1].y, p[2].y)), mxy = fmax(p[0].y, fmax(p[1].y, p[2].y));
    reinterpret_cast<double4 *>(bounds)[t] = make_double4(mnx, mxx, mny, mxy);
    int amx = 0, amn = 0, amy = 0, any_ = 0;  // first vertex attaining the bound (helpers.py:203-206)
#pragma unroll
    for (int c = 2; c >= 0; --c) {
        if (p[c].x == mxx) amx = c;
        if (p[c].x == mnx) amn = c;
        if (p[c].y == mxy) amy = c;
        if (p[c].y == mny) any_ = c;
    }
    reinterpret_cast<int4 *>(argv)[t] = make_int4(vv[amx], vv[amn], vv[amy], vv[any_]);
}

// ---- a8: node -> triangle incidence as CSR (aligned_simplex_map, src/same.py:1096-1099) -----------------------------
// count / scan / fill; the (few) triangles of a node are then put in ascending order by its own thread.
__global__ void k_inc_count(const int3 *__restrict__ tri, i64 T, const i32 *__restrict__ t_off, const i32 *__restrict__ ka_off, int W,
                            i32 *__restrict__ cnt) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const i32 nb = ka_off[find_window(t_off, W, (i32)t)];
    const int3 v = tri[t];
    atomicAdd(cnt + nb + v.x, 1); atomicAdd(cnt + nb + v.y, 1); atomicAdd(cnt + nb + v.z, 1);
}
__global__ void k_inc_fill(const int3 *__restrict__ tri, i64 T, const i32 *__restrict__ t_off, const i32 *__restrict__ ka_off, int W,
                           const i32 *__restrict__ ptr, i32 *__restrict__ cursor, i32 *__restrict__ idx) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int w = find_window(t_off, W, (i32)t);
    const i32 nb = ka_off[w], tl = (i32)t - t_off[w];
    const int3 v = tri[t];
    const i32 node[3] = {nb + v.x, nb + v.y, nb + v.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if ((c == 1 && node[1] == node[0]) || (c == 2 && (node[2] == node[0] || node[2] == node[1]))) continue;   // a set: no duplicates
        idx[ptr[node[c]] + atomicAdd(cursor + node[c], 1)] = tl;
    }
}
__global__ void k_inc_sort(const i32 *__restrict__ ptr, const i32 *__restrict__ cursor, i64 nKA, i32 *__restrict__ idx, i32 *__restrict__ len) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nKA) return;
    i32 *a = idx + ptr[i];
    const int n = cursor[i];
    len[i] = n;
    for (int k = 1; k < n; ++k) {
        const i32 x = a[k];
        int j = k - 1;
        while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; --j; }
        a[j + 1] = x;
    }
}
void batch_incidence(Batch *b) {
    if (b->have_incidence) return;
    cudaStream_t s = b->stream;
    const i64 T = b->T, nKA = b->nKA, W = b->W;
    DevBuf<i32> cnt, cursor;
    cnt.alloc(nKA + 1, s); cursor.alloc(nKA + 1, s);
    cnt.zero(s); cursor.zero(s);
    b->nt_ptr.alloc(nKA + 1, s); b->nt_idx.alloc(3 * T, s); b->nt_len.alloc(nKA + 1, s);
    if (T > 0) LAUNCH(k_inc_count, blocks_for(T, 256), 256, 0, s, b->tri.p, T, b->d_t_off.p, b->d_ka_off.p, (int)W, cnt.p);
    scan_i32(b->sec, cnt.p, b->nt_ptr.p, nKA + 1, s);
    if (T > 0) LAUNCH(k_inc_fill, blocks_for(T, 256), 256, 0, s, b->tri.p, T, b->d_t_off.p, b->d_ka_off.p, (int)W, b->nt_ptr.p, cursor.p, b->nt_idx.p);
    if (nKA > 0) LAUNCH(k_inc_sort, blocks_for(nKA, 128), 128, 0, s, b->nt_ptr.p, cursor.p, nKA, b->nt_idx.p, b->nt_len.p);
    b->nt_off.assign(W + 1, 0);
    for (i64 w = 0; w <= W; ++w) b->nt_off[w] = 3 * b->t_off[w];
    b->have_incidence = true;
}

void batch_tri_finalize(Batch *b, int ignore_same_type, int ensure_min, int remove_unconstrained) {
    cudaStream_t s = b->stream;
    REQUIRE(b->stage >= 3, SAME_E_STATE, "same_batch_tri_finalize before same_batch_tri_classify");
    batch_settle(b);
    batch_kept_columns(b);
    const i64 W = b->W, Tin = b->Tin, nKA = b->nKA, P = b->P;
    const int addback = ignore_same_type && ensure_min;
    DevBuf<i32> node_valid, has_tri, best_tri, kept_flag, kpos, ab_flag, abpos, unc_flag, uncpos, validpos, woff;
    DevBuf<unsigned long long> best_score;
    node_valid.alloc(nKA + 1, s); has_tri.alloc(nKA + 1, s); best_tri.alloc(nKA + 1, s); best_score.alloc(nKA + 1, s);
    kept_flag.alloc(Tin + 1, s); kpos.alloc(Tin + 1, s);
    ab_flag.alloc(nKA + 1, s); abpos.alloc(nKA + 1, s); unc_flag.alloc(nKA + 1, s); uncpos.alloc(nKA + 1, s); validpos.alloc(nKA + 1, s);
    woff.alloc(4 * (W + 1), s);
    node_valid.zero(s); has_tri.zero(s);
    CK(cudaMemsetAsync(best_score.p, 0xff, sizeof(unsigned long long) * (nKA + 1), s));
    LAUNCH(k_fill_i32_tri, blocks_for(nKA + 1, 256), 256, 0, s, best_tri.p, nKA + 1, 0x7fffffff);
    LAUNCH(k_tri_nodes, blocks_for(Tin + 1, 256), 256, 0, s, b->tin.p, Tin, b->d_tin_off.p, b->d_ka_off.p, (int)W, b->cls.p, b->score.p, addback,
           node_valid.p, has_tri.p, best_score.p, kept_flag.p);
    if (addback && Tin > 0)
        LAUNCH(k_tri_best, blocks_for(Tin, 256), 256, 0, s, b->tin.p, Tin, b->d_tin_off.p, b->d_ka_off.p, (int)W, b->cls.p, b->score.p, best_score.p,
               best_tri.p);
    {
        const unsigned tiles = blocks_for(nKA + 1, AB_THREADS * AB_ITEMS);
        LAUNCH(k_addback_scan, tiles, AB_THREADS, 0, s, nKA, b->d_ka_off.p, (int)W, node_valid.p, has_tri.p, best_tri.p, b->tin.p, addback,
               scan_ctx(b->sec, tiles, 3, s), ab_flag.p, unc_flag.p, abpos.p, uncpos.p, validpos.p);
    }
    exclusive_scan_i32(kept_flag.p, kpos.p, Tin + 1, b->scratch, s);
    LAUNCH(k_tri_window_offsets, blocks_for(W + 1, 128), 128, 0, s, kpos.p, abpos.p, uncpos.p, validpos.p, b->d_tin_off.p, b->d_ka_off.p, (int)W, woff.p);
    const i32 *h = b->pin_misc();
    small_d2h(b->pin_misc(), woff.p, sizeof(i32) * 4 * (W + 1), s);
    batch_sync(b);   // the one host decision of this stage: sizes, and whether any node has to go
    b->t_off.assign(h, h + W + 1);
    b->unc_off.assign(h + 2 * (W + 1), h + 3 * (W + 1));
    std::vector<i64> new_ka(h + 3 * (W + 1), h + 4 * (W + 1));
    b->T = b->t_off[W];
    b->nUnc = b->unc_off[W];
    const int renumber = remove_unconstrained && b->nUnc > 0;

    b->d_t_off.alloc(W + 1, s);
    CK(cudaMemcpyAsync(b->d_t_off.p, woff.p, sizeof(i32) * (W + 1), cudaMemcpyDeviceToDevice, s));
    b->tri.alloc(b->T, s); b->tri_src.alloc(b->T, s);
    b->unc.alloc(b->nUnc, s);
    if (Tin > 0)
        LAUNCH(k_tri_emit_kept, blocks_for(Tin, 256), 256, 0, s, b->tin.p, Tin, b->d_tin_off.p, b->d_ka_off.p, (int)W, kept_flag.p, kpos.p, b->d_t_off.p,
               validpos.p, renumber, b->tri.p, b->tri_src.p);
    if (addback && nKA > 0)
        LAUNCH(k_tri_emit_addback, blocks_for(nKA, 256), 256, 0, s, nKA, b->d_ka_off.p, (int)W, ab_flag.p, abpos.p, best_tri.p, b->tin.p, b->d_tin_off.p,
               b->d_t_off.p, woff.p + (W + 1), validpos.p, renumber, b->tri.p, b->tri_src.p);
    if (b->nUnc > 0)
        LAUNCH(k_emit_unconstrained, blocks_for(nKA, 256), 256, 0, s, nKA, b->d_ka_off.p, (int)W, unc_flag.p, uncpos.p, b->unc.p);

    if (renumber) {
        const i64 nKA2 = new_ka[W];
        DevBuf<i32> keepA2, type2, row_ptr2, poff2;
        DevBuf<double2> xy2;
        DevBuf<double> size2, cost2;
        DevBuf<int2> pairs2;
        keepA2.alloc(nKA2, s); type2.alloc(nKA2, s); xy2.alloc(nKA2, s); size2.alloc(nKA2, s);
        poff2.alloc(W + 1, s); pairs2.alloc(P, s); cost2.alloc(P, s); row_ptr2.alloc(nKA2 + 1, s);   // pairs only shrink: sized by the old count
        LAUNCH(k_compact_nodes, blocks_for(nKA, 256), 256, 0, s, nKA, node_valid.p, validpos.p, b->keepA.p, b->ka_xy.p, b->ka_type.p, b->ka_size.p,
               keepA2.p, xy2.p, type2.p, size2.p);
        {
            const unsigned tiles = blocks_for(P + 1, CP_TILE);
            LAUNCH(k_compact_pairs, tiles, CP_THREADS, 0, s, b->pairs.p, b->cost.p, P, b->d_p_off.p, b->d_ka_off.p, (int)W, node_valid.p, validpos.p,
                   b->row_ptr.p, nKA, scan_ctx(b->sec, tiles, 1, s), pairs2.p, cost2.p, row_ptr2.p, poff2.p);
        }
        // the new per-window pair offsets reach the host with the next synchronisation; the device copies are made in place
        small_d2h(b->pin_renum(), poff2.p, sizeof(i32) * (W + 1), s);
        b->pend_renum = true;
        if (b->nAi > 0) LAUNCH(k_renumber_instances, blocks_for(b->nAi, 256), 256, 0, s, b->nAi, node_valid.p, validpos.p, b->cnt.p, b->newA.p);
        b->keepA.swap(keepA2); b->ka_xy.swap(xy2); b->ka_type.swap(type2); b->ka_size.swap(size2);
        b->pairs.swap(pairs2); b->cost.swap(cost2); b->row_ptr.swap(row_ptr2);
        b->nKA = nKA2;
        b->ka_off = new_ka;
        CK(cudaMemcpyAsync(b->d_ka_off.p, woff.p + 3 * (W + 1), sizeof(i32) * (W + 1), cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(b->d_p_off.p, poff2.p, sizeof(i32) * (W + 1), cudaMemcpyDeviceToDevice, s));
        b->have_groups = false;
        b->have_start = false;   // pairs and rows were renumbered
        b->have_pair_j = false;
        b->have_pair_j16 = false;
    }

    b->t_weight.alloc(b->T, s); b->t_sign.alloc(b->T, s); b->t_bounds.alloc(4 * b->T, s); b->t_argv.alloc(4 * b->T, s);
    b->unc_list[0].alloc(UNC_CAP, s); b->unc_list[1].alloc(UNC_CAP, s); b->unc_count.alloc(2, s);
    b->unc_count.zero(s);
    b->last_unc_sep = -1;
    if (b->T > 0)
        LAUNCH(k_tri_tables, blocks_for(b->T, 256), 256, 0, s, b->tri.p, b->T, b->d_t_off.p, b->d_ka_off.p, (int)W, b->ka_xy.p, b->ka_size.p, b->t_weight.p,
               b->t_sign.p, b->t_bounds.p, b->t_argv.p, b->unc_list[0].p, b->unc_count.p);
    b->stage = 4;
    b->have_post = false;
    b->have_incidence = false;
}

}  // namespace same
