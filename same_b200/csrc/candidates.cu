// candidates.cu — a1 find_knn_within_radius (src/utils.py:709-742), a2 cell-type priority
// (src/knn_utils.py:31-65), a3 pair cost (src/same.py:1182-1189), a4 constraint grouping
// (src/helpers.py:105-138) for a whole batch of windows at once.
//
// Layout: reference cells of every window are binned on a per-window uniform grid (bin key =
// window base + by*nbx + bx), sorted with CUB radix sort, and scanned ring by ring around each
// aligned cell with a per-thread sorted top-k kept in registers.  Bins and rings whose nearest
// point is already farther than the current k-th best are skipped, so the number of distance
// evaluations is ~k-dependent, not radius-dependent.  The inclusion predicate and the ranking
// are exact fp64 (d2 = dx*dx + dy*dy, no FMA; d2 <= r*r; order (d2, ref index)); only the
// pruning uses conservative bounds.
#include "common.cuh"

namespace same {

constexpr int MAX_RINGS = 4;
static double target_per_bin() {   // reference cells per bin; SAME_B200_BIN_TARGET overrides (tuning only)
    static const double v = [] {
        const char *e = getenv("SAME_B200_BIN_TARGET");
        const double x = e ? atof(e) : 0.0;
        return x >= 1.0 ? x : 12.0;
    }();
    return v;
}
constexpr int MAX_BINS_AXIS = 2048;

// ---- binning -------------------------------------------------------------------------
__device__ __forceinline__ void bin_of(const GridParams &g, double2 p, int &bx, int &by) {
    bx = (int)floor((p.x - g.x0) * g.inv_w);
    by = (int)floor((p.y - g.y0) * g.inv_w);
    bx = min(max(bx, 0), g.nbx - 1);
    by = min(max(by, 0), g.nby - 1);
}

__global__ void k_bin_keys(const i32 *__restrict__ src, const double2 *__restrict__ sec_xy, i64 n, const i32 *__restrict__ off, int W,
                           const GridParams *__restrict__ grids, unsigned *__restrict__ keys, i32 *__restrict__ vals) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int w = find_window(off, W, (i32)i);
    const GridParams g = grids[w];
    int bx, by;
    bin_of(g, sec_xy[src[i]], bx, by);
    keys[i] = (unsigned)(g.base + by * g.nbx + bx);
    vals[i] = (i32)i;
}

__global__ void k_bin_starts(const unsigned *__restrict__ sorted_keys, i64 n, i64 nbins, i32 *__restrict__ start) {
    i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nbins) return;
    i64 lo = 0, hi = n;  // lower bound of key b
    while (lo < hi) {
        const i64 mid = (lo + hi) >> 1;
        if ((i64)sorted_keys[mid] < b) lo = mid + 1; else hi = mid;
    }
    start[b] = (i32)lo;
}

__global__ void k_gather_xy(const i32 *__restrict__ inst, const i32 *__restrict__ src, const double2 *__restrict__ sec_xy, i64 n,
                            double2 *__restrict__ out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sec_xy[src[inst[i]]];
}

// ---- top-k search ----------------------------------------------------------------------
// (d, j) < (bd, bj) lexicographically; bitwise ops so that the compiler emits predicate logic, not branches (the
// short-circuit form compiled to divergent branches inside the insertion: 180 SASS instructions at ~3 active lanes)
__device__ __forceinline__ bool cand_less(double d, i32 j, double bd, i32 bj) { return (d < bd) | ((d == bd) & (j < bj)); }

// Sorted top-k of one query in registers, right-aligned: slots [KCAP-knn, KCAP) hold the list (ascending), the
// slots before it are -inf sentinels that never move, so the k-th best is always the LAST slot (a static register
// index; a runtime index would push the arrays into local memory).  Branch-free: lt[u] = candidate sorts before
// slot u (on the old values); slot u takes slot u-1 if lt[u-1], the candidate if only lt[u], else keeps its value.
// A candidate that does not beat the last slot — e.g. the (inf, INT_MAX) filler of an idle lane — is a no-op.
template <int KCAP>
__device__ __forceinline__ void topk_insert(double (&bd)[KCAP], i32 (&bj)[KCAP], double d2, i32 j) {
    bool lt[KCAP];
#pragma unroll
    for (int u = 0; u < KCAP; ++u) lt[u] = cand_less(d2, j, bd[u], bj[u]);
#pragma unroll
    for (int u = KCAP - 1; u > 0; --u) {
        bd[u] = lt[u - 1] ? bd[u - 1] : (lt[u] ? d2 : bd[u]);
        bj[u] = lt[u - 1] ? bj[u - 1] : (lt[u] ? j : bj[u]);
    }
    bd[0] = lt[0] ? d2 : bd[0];
    bj[0] = lt[0] ? j : bj[0];
}

// Search kernel.  Thread = query; every lane walks the bins around ITS OWN bin, nearest ring first, and skips bins and
// rings that cannot beat min(r^2, current k-th best) (conservative bounds; the exact fp64 predicate decides each
// candidate).  What is shared by the warp is the CONTROL FLOW: all lanes step through the same (ring, dy, dx)
// offsets and the same candidate slots of "their" bin, a lane whose bin is pruned or shorter is just predicated off.
// That keeps the warp converged, which lets a lane postpone its insertions: the sorted insertion is ~80
// instructions and only ~1 candidate in 6 needs it (ncu on the divergent version: 47 % of all warp instructions ran
// with 3 of 32 lanes active), so a lane appends a passing candidate to a small queue in shared memory and the warp
// drains all queues together when one is nearly full — the insertion network then runs with most lanes busy.
// A warp-wide shared candidate stream was tried and rejected: 30/32 lanes active, but 3-4x the distance evaluations
// and insertions in stream order (profiles/r1j_knn_warp_stream_experiment.md).
constexpr int KNN_Q = 8;      // queue slots per lane

// drain every lane's queue: as many insertion rounds as the fullest queue of the warp holds
template <int KCAP>
__device__ __forceinline__ void knn_flush(double (&bd)[KCAP], i32 (&bj)[KCAP], int &qn, const double (*q_d)[128], const i32 (*q_j)[128], int tid) {
    const int rounds = __reduce_max_sync(0xffffffffu, qn);
    for (int r = 0; r < rounds; ++r) {
        const bool have = r < qn;
        const double d2 = have ? q_d[r][tid] : INFINITY;
        const i32 j = have ? q_j[r][tid] : 0x7fffffff;
        topk_insert<KCAP>(bd, bj, d2, j);
    }
    qn = 0;
}

template <int KCAP>
__global__ void __launch_bounds__(128) k_knn(const double2 *__restrict__ sa_xy, const i32 *__restrict__ sa_inst, i64 nAi,
                                             const i32 *__restrict__ a_off, int W, const GridParams *__restrict__ grids,
                                             const i32 *__restrict__ bin_start, const double2 *__restrict__ sr_xy,
                                             const i32 *__restrict__ sr_inst, double r2, int knn, i32 *__restrict__ cand,
                                             i32 *__restrict__ cnt, i32 *__restrict__ r_used) {
    __shared__ double q_d[KNN_Q][128];
    __shared__ i32 q_j[KNN_Q][128];
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    const i64 t = (i64)blockIdx.x * blockDim.x + tid;
    const bool live = t < nAi;
    i32 inst = 0;
    double2 q = make_double2(0.0, 0.0);
    double gw = 1.0, px = 0.0, py = 0.0;
    int nbx = 0, nby = 0, base = 0, rings = -1, cbx = 0, cby = 0;
    if (live) {
        inst = sa_inst[t];
        q = sa_xy[t];
        const GridParams *gp = grids + find_window(a_off, W, inst);
        gw = gp->w; nbx = gp->nbx; nby = gp->nby; base = gp->base; rings = gp->rings;
        px = q.x - gp->x0; py = q.y - gp->y0;
        cbx = min(max((int)floor(px * gp->inv_w), 0), nbx - 1);
        cby = min(max((int)floor(py * gp->inv_w), 0), nby - 1);
    }
    const double eps = 1e-7 * gw;
    const double fx = px - cbx * gw, fy = py - cby * gw;
    const double edge = fmin(fmin(fx, gw - fx), fmin(fy, gw - fy)) - eps;   // distance to the nearest side of the own bin
    double bd[KCAP];
    i32 bj[KCAP];
    const int head = KCAP - knn;
#pragma unroll
    for (int s = 0; s < KCAP; ++s) { bd[s] = (s < head) ? -INFINITY : INFINITY; bj[s] = 0x7fffffff; }
    int qn = 0;
    bool open = live;   // this lane still has rings to visit
    const int max_rings = __reduce_max_sync(FULL, rings);
    for (int ring = 0; ring <= max_rings; ++ring) {
        if (ring > 0) {
            // nearest possible point of this ring (Chebyshev distance `ring` bins from the own bin)
            if (__any_sync(FULL, qn > 0)) knn_flush<KCAP>(bd, bj, qn, q_d, q_j, tid);   // the ring test wants the true k-th best
            const double gap = (ring - 1) * gw + edge;
            open = open && ring <= rings && !(gap > 0.0 && gap * gap > fmin(bd[KCAP - 1], r2));
            if (!__any_sync(FULL, open)) break;
        }
        for (int dy = -ring; dy <= ring; ++dy) {
            const int by = cby + dy;
            const bool row_ok = open && by >= 0 && by < nby;
            const double gy = fmax(0.0, fmax(by * gw - py, py - (by + 1) * gw) - eps);
            const double gy2 = gy * gy;
            const int step = (dy == -ring || dy == ring || ring == 0) ? 1 : 2 * ring;
            for (int dx = -ring; dx <= ring; dx += step) {
                const int bx = cbx + dx;
                const double gx = fmax(0.0, fmax(bx * gw - px, px - (bx + 1) * gw) - eps);
                const bool want = row_ok && bx >= 0 && bx < nbx && gx * gx + gy2 <= fmin(bd[KCAP - 1], r2);
                i32 s0 = 0, len = 0;
                if (want) {
                    const i32 b = base + by * nbx + bx;
                    s0 = bin_start[b];
                    len = bin_start[b + 1] - s0;
                }
                const int maxlen = __reduce_max_sync(FULL, len);
                for (int i = 0; i < maxlen; ++i) {
                    if (i < len) {
                        const double2 p = sr_xy[s0 + i];
                        const double ddx = __dsub_rn(p.x, q.x), ddy = __dsub_rn(p.y, q.y);
                        const double d2 = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
                        if (d2 <= fmin(bd[KCAP - 1], r2)) {
                            if (qn == KNN_Q) {   // rare: this lane alone filled its queue between two warp-wide drains
                                for (int r = 0; r < KNN_Q; ++r) topk_insert<KCAP>(bd, bj, q_d[r][tid], q_j[r][tid]);
                                qn = 0;
                            }
                            q_d[qn][tid] = d2;
                            q_j[qn][tid] = sr_inst[s0 + i];
                            ++qn;
                        }
                    }
                }
            }
        }
    }
    knn_flush<KCAP>(bd, bj, qn, q_d, q_j, tid);
    if (!live) return;
    int found = 0;
#pragma unroll
    for (int u = 0; u < KCAP; ++u) found += (u >= head) && (bd[u] < INFINITY);
    cnt[inst] = found;
    i32 *out = cand + (i64)inst * knn;
#pragma unroll
    for (int u = 0; u < KCAP; ++u)
        if (u >= head) {
            const bool ok = (u - head) < found;
            out[u - head] = ok ? bj[u] : -1;
            if (ok) r_used[bj[u]] = 1;
        }
}

// generic path for knn > 32: top-k lives in global scratch ([slot][query] so threads coalesce)
__global__ void __launch_bounds__(128) k_knn_big(const double2 *__restrict__ sa_xy, const i32 *__restrict__ sa_inst, i64 nAi,
                                                 const i32 *__restrict__ a_off, int W, const GridParams *__restrict__ grids,
                                                 const i32 *__restrict__ bin_start, const double2 *__restrict__ sr_xy,
                                                 const i32 *__restrict__ sr_inst, double r2, int knn, double *__restrict__ gd,
                                                 i32 *__restrict__ gj, i32 *__restrict__ cand, i32 *__restrict__ cnt,
                                                 i32 *__restrict__ r_used) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nAi) return;
    const i32 inst = sa_inst[t];
    const double2 q = sa_xy[t];
    const GridParams g = grids[find_window(a_off, W, inst)];
    int cbx, cby;
    bin_of(g, q, cbx, cby);
    const double px = q.x - g.x0, py = q.y - g.y0;
    const double eps = 1e-7 * g.w;
    int found = 0;
    double tau_d = INFINITY;
    i32 tau_j = 0x7fffffff;
#define GD(u) gd[(i64)(u) * nAi + t]
#define GJ(u) gj[(i64)(u) * nAi + t]
    for (int ring = 0; ring <= g.rings; ++ring) {
        if (ring > 0) {
            const double fx = px - cbx * g.w, fy = py - cby * g.w;
            const double gap = (ring - 1) * g.w + fmin(fmin(fx, g.w - fx), fmin(fy, g.w - fy)) - eps;
            if (gap > 0.0 && gap * gap > fmin(tau_d, r2)) break;
        }
        for (int dy = -ring; dy <= ring; ++dy) {
            const int by = cby + dy;
            if (by < 0 || by >= g.nby) continue;
            const int step = (dy == -ring || dy == ring || ring == 0) ? 1 : 2 * ring;
            const double ylo = by * g.w, yhi = ylo + g.w;
            const double gy = fmax(0.0, fmax(ylo - py, py - yhi) - eps);
            for (int dx = -ring; dx <= ring; dx += step) {
                const int bx = cbx + dx;
                if (bx < 0 || bx >= g.nbx) continue;
                const double xlo = bx * g.w, xhi = xlo + g.w;
                const double gx = fmax(0.0, fmax(xlo - px, px - xhi) - eps);
                if (gx * gx + gy * gy > fmin(tau_d, r2)) continue;
                const i32 b = g.base + by * g.nbx + bx;
                const i32 s1 = bin_start[b + 1];
                for (i32 s = bin_start[b]; s < s1; ++s) {
                    const double2 p = sr_xy[s];
                    const double ddx = __dsub_rn(p.x, q.x), ddy = __dsub_rn(p.y, q.y);
                    const double d2 = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
                    if (d2 > r2) continue;
                    const i32 j = sr_inst[s];
                    if (found == knn && !cand_less(d2, j, tau_d, tau_j)) continue;
                    int u = (found == knn) ? knn - 1 : found;
                    while (u > 0 && cand_less(d2, j, GD(u - 1), GJ(u - 1))) { GD(u) = GD(u - 1); GJ(u) = GJ(u - 1); --u; }
                    GD(u) = d2; GJ(u) = j;
                    if (found < knn) ++found;
                    if (found == knn) { tau_d = GD(knn - 1); tau_j = GJ(knn - 1); }
                }
            }
        }
    }
    cnt[inst] = found;
    for (int u = 0; u < knn; ++u) {
        cand[(i64)inst * knn + u] = (u < found) ? GJ(u) : -1;
        if (u < found) r_used[GJ(u)] = 1;
    }
#undef GD
#undef GJ
}

// ---- a2: cell-type priority ------------------------------------------------------------
// winner of ref j = smallest aligned instance whose nearest ref is j with the same type
// (equivalent parallel form of the sequential loop)
__global__ void k_claim(const i32 *__restrict__ cand, const i32 *__restrict__ cnt, int knn, i64 nAi, const i32 *__restrict__ a_src,
                        const i32 *__restrict__ r_src, const i32 *__restrict__ a_type, const i32 *__restrict__ r_type,
                        i32 *__restrict__ claim) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nAi || cnt[i] == 0) return;
    const i32 j0 = cand[i * knn];
    if (a_type[a_src[i]] == r_type[r_src[j0]]) atomicMin(claim + j0, (i32)i);
}
__global__ void k_eff(const i32 *__restrict__ cand, const i32 *__restrict__ cnt, int knn, i64 nAi, const i32 *__restrict__ claim,
                      i32 *__restrict__ eff) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nAi) return;
    const int c = cnt[i];
    eff[i] = (c > 0 && claim[cand[i * knn]] == (i32)i) ? 1 : c;
}

// ---- compaction + emission ----------------------------------------------------------------
__global__ void k_flag_pos(const i32 *__restrict__ v, i64 n, i32 *__restrict__ f) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = v[i] > 0;
    if (i == n) f[i] = 0;
}
__global__ void k_fill_i32(i32 *p, i64 n, i32 v) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// off3[0..W] / [W+1..2W+1] / [2W+2..3W+2]: kept-aligned, kept-ref and pair offsets of each window
__global__ void k_window_offsets(const i32 *__restrict__ newA, const i32 *__restrict__ newR, const i32 *__restrict__ poff,
                                 const i32 *__restrict__ a_off, const i32 *__restrict__ r_off, int W, i32 *__restrict__ off3) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > W) return;
    off3[w] = newA[a_off[w]];
    off3[W + 1 + w] = newR[r_off[w]];
    off3[2 * (W + 1) + w] = poff[a_off[w]];
}

__global__ void k_emit_aligned(const i32 *__restrict__ cnt, const i32 *__restrict__ newA, const i32 *__restrict__ poff, i64 nAi,
                               const i32 *__restrict__ a_src, const double2 *__restrict__ sec_xy, const i32 *__restrict__ sec_type,
                               const double *__restrict__ sec_size, i32 *__restrict__ keepA, double2 *__restrict__ ka_xy,
                               i32 *__restrict__ ka_type, double *__restrict__ ka_size, i32 *__restrict__ row_ptr, i64 nKA, i32 P) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) row_ptr[nKA] = P;
    if (i >= nAi || cnt[i] == 0) return;
    const i32 k = newA[i], row = a_src[i];
    keepA[k] = row;
    ka_xy[k] = sec_xy[row];
    ka_type[k] = sec_type[row];
    ka_size[k] = sec_size[row];
    row_ptr[k] = poff[i];
}
__global__ void k_emit_ref(const i32 *__restrict__ used, const i32 *__restrict__ newR, i64 nRi, const i32 *__restrict__ r_src,
                           const double2 *__restrict__ sec_xy, const double *__restrict__ sec_size, i32 *__restrict__ keepR,
                           double2 *__restrict__ kr_xy, double *__restrict__ kr_size) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRi || used[i] == 0) return;
    const i32 k = newR[i], row = r_src[i];
    keepR[k] = row;
    kr_xy[k] = sec_xy[row];
    kr_size[k] = sec_size[row];
}

// one thread per (aligned instance, slot): pair indices + cost (src/same.py:1183-1188)
__global__ void k_emit_pairs(const i32 *__restrict__ cand, const i32 *__restrict__ eff, int knn, i64 nAi, const i32 *__restrict__ newA,
                             const i32 *__restrict__ newR, const i32 *__restrict__ poff, const i32 *__restrict__ a_off, int W,
                             const i32 *__restrict__ ka_off, const i32 *__restrict__ kr_off, const i32 *__restrict__ a_src,
                             const i32 *__restrict__ r_src, const double2 *__restrict__ a_xy, const double2 *__restrict__ r_xy,
                             const double *__restrict__ a_prob, const double *__restrict__ r_prob, int K, double ct_coeff,
                             double dist_coeff, int2 *__restrict__ pairs, double *__restrict__ cost) {
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 i = tid / knn;
    const int slot = (int)(tid - i * knn);
    if (i >= nAi || slot >= eff[i]) return;
    const int w = find_window(a_off, W, (i32)i);
    const i32 jinst = cand[i * knn + slot];
    const i32 p = poff[i] + slot;
    pairs[p] = make_int2(newA[i] - ka_off[w], newR[jinst] - kr_off[w]);
    const i32 ar = a_src[i], rr = r_src[jinst];
    const double *pa = a_prob + (i64)ar * K, *pr = r_prob + (i64)rr * K;
    double s = 0.0;
    for (int c = 0; c < K; ++c) s = __dadd_rn(s, fabs(__dsub_rn(pa[c], pr[c])));  // left-to-right
    const double2 A = a_xy[ar], R = r_xy[rr];
    const double dc = __dadd_rn(fabs(__dsub_rn(A.x, R.x)), fabs(__dsub_rn(A.y, R.y)));
    cost[p] = __dadd_rn(__dmul_rn(ct_coeff, s), __dmul_rn(dist_coeff, dc));
}

static int bits_for(i64 n) {
    int b = 1;
    while ((1ll << b) < n) ++b;
    return b;
}

void batch_candidates(Batch *b, double radius, int knn, int priority, double dist_ct_coeff) {
    Section *sec = b->sec;
    cudaStream_t s = b->stream;
    const i64 W = b->W, nAi = b->nAi, nRi = b->nRi;
    REQUIRE(knn >= 1 && knn <= SAME_MAX_KNN, SAME_E_LIMIT, "knn must be in [1, SAME_MAX_KNN]");
    REQUIRE(radius >= 0 && std::isfinite(radius), SAME_E_ARG, "radius must be finite and >= 0");
    REQUIRE((double)nAi * knn < 2.0e9, SAME_E_LIMIT, "aligned cells x knn exceeds 2^31");
    b->knn = knn;
    b->radius = radius;

    // per-window grids over (window rectangle ∩ section bbox)
    std::vector<GridParams> grids(W);
    i64 nbins = 0;
    for (i64 w = 0; w < W; ++w) {
        const double *r = &b->rects[4 * w];
        double x0 = std::max(r[0], sec->bbox[0]), x1 = std::min(r[1], sec->bbox[1]);
        double y0 = std::max(r[2], sec->bbox[2]), y1 = std::min(r[3], sec->bbox[3]);
        if (!(x1 > x0)) x1 = x0;
        if (!(y1 > y0)) y1 = y0;
        const double ex = x1 - x0, ey = y1 - y0;
        const i64 nref = b->r_off[w + 1] - b->r_off[w];
        double bw = std::sqrt(target_per_bin() * std::max(ex, 1e-300) * std::max(ey, 1e-300) / (double)std::max<i64>(nref, 1));
        bw = std::max(bw, radius * (1.0 + 1e-9) / MAX_RINGS);
        bw = std::max(bw, std::max(ex, ey) / MAX_BINS_AXIS);
        if (!(bw > 0.0) || !std::isfinite(bw)) bw = 1.0;
        GridParams g;
        g.x0 = x0; g.y0 = y0; g.w = bw; g.inv_w = 1.0 / bw;
        g.nbx = (i32)std::floor(ex / bw) + 2;
        g.nby = (i32)std::floor(ey / bw) + 2;
        g.rings = (i32)std::ceil(radius / bw * (1.0 + 1e-9) + 1e-9);
        if (g.rings < 1) g.rings = 1;
        REQUIRE(nbins + (i64)g.nbx * g.nby < (1ll << 30), SAME_E_LIMIT, "bin grid too large");
        g.base = (i32)nbins;
        nbins += (i64)g.nbx * g.nby;
        grids[w] = g;
    }
    DevBuf<GridParams> d_grids;
    d_grids.alloc(W, s);
    CK(cudaMemcpyAsync(d_grids.p, grids.data(), sizeof(GridParams) * W, cudaMemcpyHostToDevice, s));

    // sort both frames' instances by bin
    DevBuf<unsigned> keys_in, keys_out;
    DevBuf<i32> vals_in, sr_inst, sa_inst, bin_start;
    DevBuf<double2> sr_xy, sa_xy;
    const i64 nmax = std::max(nAi, nRi);
    keys_in.alloc(nmax, s); keys_out.alloc(nmax, s); vals_in.alloc(nmax, s);
    sr_inst.alloc(nRi, s); sa_inst.alloc(nAi, s); sr_xy.alloc(nRi, s); sa_xy.alloc(nAi, s);
    bin_start.alloc(nbins + 1, s);
    const int kb = bits_for(nbins + 1);
    auto sort_frame = [&](const DevBuf<i32> &src, const DevBuf<double2> &xy, i64 n, const DevBuf<i32> &off, DevBuf<i32> &out_inst,
                          DevBuf<double2> &out_xy) {
        if (n == 0) return;
        LAUNCH(k_bin_keys, blocks_for(n, 256), 256, 0, s, src.p, xy.p, n, off.p, (int)W, d_grids.p, keys_in.p, vals_in.p);
        size_t bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in.p, keys_out.p, vals_in.p, out_inst.p, (int)n, 0, kb, s));
        void *tmp = b->scratch.get(bytes, s);
        {
            ProfScope prof("cub::DeviceRadixSort::SortPairs(bins)", s);
            CK(cub::DeviceRadixSort::SortPairs(tmp, bytes, keys_in.p, keys_out.p, vals_in.p, out_inst.p, (int)n, 0, kb, s));
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        LAUNCH(k_gather_xy, blocks_for(n, 256), 256, 0, s, out_inst.p, src.p, xy.p, n, out_xy.p);
    };
    sort_frame(b->r_src, sec->r_xy, nRi, b->d_r_off, sr_inst, sr_xy);
    LAUNCH(k_bin_starts, blocks_for(nbins + 1, 256), 256, 0, s, keys_out.p, nRi, nbins, bin_start.p);
    sort_frame(b->a_src, sec->a_xy, nAi, b->d_a_off, sa_inst, sa_xy);

    // top-k search
    b->cand.alloc(nAi * knn, s);
    b->cnt.alloc(nAi + 1, s);
    b->r_used.alloc(nRi + 1, s);
    b->r_used.zero(s);
    const double r2 = radius * radius;
    if (nAi > 0) {
        const unsigned grid = blocks_for(nAi, 128);
#define KNN_ARGS sa_xy.p, sa_inst.p, nAi, b->d_a_off.p, (int)W, d_grids.p, bin_start.p, sr_xy.p, sr_inst.p, r2, knn
        if (knn <= 4) LAUNCH(k_knn<4>, grid, 128, 0, s, KNN_ARGS, b->cand.p, b->cnt.p, b->r_used.p);
        else if (knn <= 8) LAUNCH(k_knn<8>, grid, 128, 0, s, KNN_ARGS, b->cand.p, b->cnt.p, b->r_used.p);
        else if (knn <= 16) LAUNCH(k_knn<16>, grid, 128, 0, s, KNN_ARGS, b->cand.p, b->cnt.p, b->r_used.p);
        else if (knn <= 32) LAUNCH(k_knn<32>, grid, 128, 0, s, KNN_ARGS, b->cand.p, b->cnt.p, b->r_used.p);
        else {
            DevBuf<double> gd;
            DevBuf<i32> gj;
            gd.alloc(nAi * knn, s); gj.alloc(nAi * knn, s);
            LAUNCH(k_knn_big, grid, 128, 0, s, KNN_ARGS, gd.p, gj.p, b->cand.p, b->cnt.p, b->r_used.p);
        }
#undef KNN_ARGS
    }

    // a2: priority filter decides how many pairs each aligned row emits; compaction of the frames is
    // still a1's (knn_utils.py:14 re-uses the frames find_knn_within_radius returned)
    const i32 *eff = b->cnt.p;
    if (priority) {
        DevBuf<i32> claim;
        claim.alloc(nRi, s);
        b->eff.alloc(nAi + 1, s);
        LAUNCH(k_fill_i32, blocks_for(nRi, 256), 256, 0, s, claim.p, nRi, 0x7fffffff);
        LAUNCH(k_claim, blocks_for(nAi, 256), 256, 0, s, b->cand.p, b->cnt.p, knn, nAi, b->a_src.p, b->r_src.p, sec->a_type.p, sec->r_type.p, claim.p);
        LAUNCH(k_eff, blocks_for(nAi, 256), 256, 0, s, b->cand.p, b->cnt.p, knn, nAi, claim.p, b->eff.p);
        eff = b->eff.p;
    }

    // compaction maps
    DevBuf<i32> flagA, newA, newR, poff, off3;
    flagA.alloc(nAi + 1, s); newA.alloc(nAi + 1, s); newR.alloc(nRi + 1, s); poff.alloc(nAi + 1, s); off3.alloc(3 * (W + 1), s);
    LAUNCH(k_flag_pos, blocks_for(nAi + 1, 256), 256, 0, s, b->cnt.p, nAi, flagA.p);
    exclusive_scan_i32(flagA.p, newA.p, nAi + 1, b->scratch, s);
    exclusive_scan_i32(b->r_used.p, newR.p, nRi + 1, b->scratch, s);   // r_used[nRi] == 0 from the memset
    LAUNCH(k_flag_pos, 1, 1, 0, s, (const i32 *)nullptr, (i64)0, (i32 *)eff + nAi);  // sentinel: eff[nAi] = 0
    exclusive_scan_i32(eff, poff.p, nAi + 1, b->scratch, s);
    LAUNCH(k_window_offsets, blocks_for(W + 1, 128), 128, 0, s, newA.p, newR.p, poff.p, b->d_a_off.p, b->d_r_off.p, (int)W, off3.p);
    std::vector<i32> h(3 * (W + 1));
    CK(cudaMemcpyAsync(h.data(), off3.p, sizeof(i32) * h.size(), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    b->ka_off.assign(h.begin(), h.begin() + W + 1);
    b->kr_off.assign(h.begin() + W + 1, h.begin() + 2 * (W + 1));
    b->p_off.assign(h.begin() + 2 * (W + 1), h.end());
    b->nKA = b->ka_off[W]; b->nKR = b->kr_off[W]; b->P = b->p_off[W];
    b->d_ka_off.alloc(W + 1, s); b->d_kr_off.alloc(W + 1, s); b->d_p_off.alloc(W + 1, s);
    CK(cudaMemcpyAsync(b->d_ka_off.p, off3.p, sizeof(i32) * (W + 1), cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(b->d_kr_off.p, off3.p + (W + 1), sizeof(i32) * (W + 1), cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(b->d_p_off.p, off3.p + 2 * (W + 1), sizeof(i32) * (W + 1), cudaMemcpyDeviceToDevice, s));

    // emission
    b->keepA.alloc(b->nKA, s); b->ka_xy.alloc(b->nKA, s); b->ka_type.alloc(b->nKA, s); b->ka_size.alloc(b->nKA, s);
    b->row_ptr.alloc(b->nKA + 1, s);
    b->keepR.alloc(b->nKR, s); b->kr_xy.alloc(b->nKR, s); b->kr_size.alloc(b->nKR, s);
    b->pairs.alloc(b->P, s); b->cost.alloc(b->P, s);
    LAUNCH(k_emit_aligned, blocks_for(std::max<i64>(nAi, 1), 256), 256, 0, s, b->cnt.p, newA.p, poff.p, nAi, b->a_src.p, sec->a_xy.p, sec->a_type.p,
           sec->a_size.p, b->keepA.p, b->ka_xy.p, b->ka_type.p, b->ka_size.p, b->row_ptr.p, b->nKA, (i32)b->P);
    if (nRi > 0)
        LAUNCH(k_emit_ref, blocks_for(nRi, 256), 256, 0, s, b->r_used.p, newR.p, nRi, b->r_src.p, sec->r_xy.p, sec->r_size.p, b->keepR.p,
               b->kr_xy.p, b->kr_size.p);
    if (nAi > 0 && b->P > 0)
        LAUNCH(k_emit_pairs, blocks_for(nAi * knn, 256), 256, 0, s, b->cand.p, eff, knn, nAi, newA.p, newR.p, poff.p, b->d_a_off.p, (int)W,
               b->d_ka_off.p, b->d_kr_off.p, b->a_src.p, b->r_src.p, sec->a_xy.p, sec->r_xy.p, sec->a_prob.p, sec->r_prob.p, sec->K,
               dist_ct_coeff, dist_ct_coeff * 0.001, b->pairs.p, b->cost.p);
    CK(cudaStreamSynchronize(s));  // temporaries (DevBuf) are released stream-ordered, but keep the stage boundary simple
    b->stage = 1;
    b->have_groups = false;
    b->Tin = b->T = 0;
}

// ---- a4: ref_to_pairs groups ----------------------------------------------------------------
// Groups must come out in FIRST-APPEARANCE order of j with pair indices ascending inside (dict insertion order,
// helpers.py:105-110).  first[j] = smallest pair index of ref j (atomicMin); a pair p is a group head iff first[j_p] == p, so
// an exclusive scan of the head flags numbers the groups in first-appearance order (window-major for free, pairs are
// window-major); group sizes come from a per-ref counter; members are scattered with a per-group cursor and every
// (small, ~knn entries) group is then sorted by its own thread.  No 8M-element radix sort.
__global__ void k_group_count(const int2 *__restrict__ pairs, i64 P, const i32 *__restrict__ p_off, const i32 *__restrict__ kr_off, int W,
                              i32 *__restrict__ first, i32 *__restrict__ cnt) {
    i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const i32 r = kr_off[find_window(p_off, W, (i32)p)] + pairs[p].y;
    atomicMin(first + r, (i32)p);
    atomicAdd(cnt + r, 1);
}
__global__ void k_group_heads(const int2 *__restrict__ pairs, i64 P, const i32 *__restrict__ p_off, const i32 *__restrict__ kr_off, int W,
                              const i32 *__restrict__ first, i32 *__restrict__ head) {
    i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) head[p] = first[kr_off[find_window(p_off, W, (i32)p)] + pairs[p].y] == (i32)p;
    if (p == P) head[p] = 0;
}
__device__ __forceinline__ unsigned long long enc_pos_f64(double d) { return (unsigned long long)__double_as_longlong(d); }  // d >= 0
__global__ void k_window_max_size(const double *__restrict__ kr_size, i64 nKR, const i32 *__restrict__ kr_off, int W,
                                  unsigned long long *__restrict__ wmax) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nKR) return;
    const double v = kr_size[i];
    if (v > 1.0) atomicMax(wmax + find_window(kr_off, W, (i32)i), enc_pos_f64(v));
}
// one thread per kept ref: its group id, node, size and limit
__global__ void k_group_setup(i64 nKR, const i32 *__restrict__ kr_off, int W, const i32 *__restrict__ first, const i32 *__restrict__ cnt,
                              const i32 *__restrict__ gid_at, const double *__restrict__ kr_size, const unsigned long long *__restrict__ wmax,
                              int max_matches, int multiplier, i32 *__restrict__ ref_gid, i32 *__restrict__ g_node, i32 *__restrict__ g_cnt,
                              i32 *__restrict__ g_limit, i64 G) {
    i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) g_cnt[G] = 0;
    if (r >= nKR) return;
    if (first[r] == 0x7fffffff) { ref_gid[r] = -1; return; }
    const int w = find_window(kr_off, W, (i32)r);
    const i32 g = gid_at[first[r]];
    ref_gid[r] = g;
    g_node[g] = (i32)r - kr_off[w];
    g_cnt[g] = cnt[r];
    const unsigned long long m = wmax[w];  // 0 = no ref with size > 1 in this window (helpers.py:121)
    int lim = max_matches;
    if (m != 0ull && kr_size[r] > 1.0) {
        const int mult = multiplier >= 0 ? multiplier : (int)__longlong_as_double((long long)m);  // int(ref_df['size'].max())
        lim = mult * max_matches;
    }
    g_limit[g] = lim;
}
__global__ void k_group_fill(const int2 *__restrict__ pairs, i64 P, const i32 *__restrict__ p_off, const i32 *__restrict__ kr_off, int W,
                             const i32 *__restrict__ ref_gid, const i32 *__restrict__ g_ptr, i32 *__restrict__ cursor, i32 *__restrict__ g_idx) {
    i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int w = find_window(p_off, W, (i32)p);
    const i32 g = ref_gid[kr_off[w] + pairs[p].y];
    g_idx[g_ptr[g] + atomicAdd(cursor + g, 1)] = (i32)p - p_off[w];
}
// ascending pair index inside every group (insertion sort; groups hold ~knn entries)
__global__ void k_group_sort(const i32 *__restrict__ g_ptr, i64 G, i32 *__restrict__ g_idx) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const i32 lo = g_ptr[g], n = g_ptr[g + 1] - lo;
    i32 *a = g_idx + lo;
    if (n <= 32) {
        i32 v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = k < n ? a[k] : 0x7fffffff;
        // odd-even transposition network would touch all 32 slots; groups are ~8 long, so a bounded insertion sort over the
        // live prefix is cheaper.  Indexing is dynamic -> local memory, but it stays in L1.
        for (int i = 1; i < n; ++i) {
            const i32 x = v[i];
            int j = i - 1;
            while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
            v[j + 1] = x;
        }
        for (int k = 0; k < n; ++k) a[k] = v[k];
    } else {
        for (int i = 1; i < n; ++i) {
            const i32 x = a[i];
            int j = i - 1;
            while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; --j; }
            a[j + 1] = x;
        }
    }
}
__global__ void k_pick(const i32 *__restrict__ scanned, const i32 *__restrict__ at, int n, i32 *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = scanned[at[i]];
}

void batch_groups(Batch *b, int max_matches, int multiplier) {
    cudaStream_t s = b->stream;
    const i64 W = b->W, P = b->P, nKR = b->nKR;
    REQUIRE(b->stage >= 1, SAME_E_STATE, "same_batch_groups before same_batch_candidates");
    b->g_off.assign(W + 1, 0);
    b->G = 0;
    b->have_groups = true;
    if (P == 0) {  // keep REF_GROUP_PTR (G + 1 = 1 element) readable
        b->g_ptr.alloc(1, s);
        b->g_ptr.zero(s);
        CK(cudaStreamSynchronize(s));
        return;
    }
    DevBuf<i32> first, cnt, head, gid_at, goff, ref_gid, g_cnt, cursor;
    DevBuf<unsigned long long> wmax;
    first.alloc(nKR, s); cnt.alloc(nKR, s); head.alloc(P + 1, s); gid_at.alloc(P + 1, s); goff.alloc(W + 1, s); ref_gid.alloc(nKR, s);
    wmax.alloc(W, s);
    wmax.zero(s);
    cnt.zero(s);
    LAUNCH(k_fill_i32, blocks_for(nKR, 256), 256, 0, s, first.p, nKR, 0x7fffffff);
    LAUNCH(k_group_count, blocks_for(P, 256), 256, 0, s, b->pairs.p, P, b->d_p_off.p, b->d_kr_off.p, (int)W, first.p, cnt.p);
    LAUNCH(k_group_heads, blocks_for(P + 1, 256), 256, 0, s, b->pairs.p, P, b->d_p_off.p, b->d_kr_off.p, (int)W, first.p, head.p);
    exclusive_scan_i32(head.p, gid_at.p, P + 1, b->scratch, s);
    LAUNCH(k_pick, blocks_for(W + 1, 128), 128, 0, s, gid_at.p, b->d_p_off.p, (int)(W + 1), goff.p);
    LAUNCH(k_window_max_size, blocks_for(nKR, 256), 256, 0, s, b->kr_size.p, nKR, b->d_kr_off.p, (int)W, wmax.p);
    std::vector<i32> h(W + 1);
    CK(cudaMemcpyAsync(h.data(), goff.p, sizeof(i32) * (W + 1), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    b->g_off.assign(h.begin(), h.end());
    const i64 G = b->G = b->g_off[W];
    b->g_node.alloc(G, s); b->g_ptr.alloc(G + 1, s); b->g_idx.alloc(P, s); b->g_limit.alloc(G, s);
    g_cnt.alloc(G + 1, s); cursor.alloc(G, s);
    cursor.zero(s);
    LAUNCH(k_group_setup, blocks_for(nKR, 256), 256, 0, s, nKR, b->d_kr_off.p, (int)W, first.p, cnt.p, gid_at.p, b->kr_size.p, wmax.p, max_matches,
           multiplier, ref_gid.p, b->g_node.p, g_cnt.p, b->g_limit.p, G);
    exclusive_scan_i32(g_cnt.p, b->g_ptr.p, G + 1, b->scratch, s);
    LAUNCH(k_group_fill, blocks_for(P, 256), 256, 0, s, b->pairs.p, P, b->d_p_off.p, b->d_kr_off.p, (int)W, ref_gid.p, b->g_ptr.p, cursor.p, b->g_idx.p);
    LAUNCH(k_group_sort, blocks_for(G, 128), 128, 0, s, b->g_ptr.p, G, b->g_idx.p);
    CK(cudaStreamSynchronize(s));
}

}  // namespace same
