/*
 * same_oracle.c — CPU restatement of the SAME hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call this file.  The product (same_b200/) never does: it
 * fails loudly when the CUDA library is missing.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py)
 * against golden vectors produced by running the unmodified Python reference in the
 * build container (tests/golden/gen_golden.py -> the .npz files under tests/golden).
 *
 * Each function restates, in plain sequential C (optionally OpenMP over independent
 * items), what the cited reference lines compute.  Paths are relative to the reference
 * repository root.  Floating point follows SURVEY.md App. A: every expression the
 * reference evaluates as scalar Python arithmetic is evaluated here one IEEE operation
 * at a time (compile with -ffp-contract=off); the two places where the reference goes
 * through BLAS ddot (1-D np.linalg.norm / np.dot, helpers.py:282-288,305-307) use an
 * explicit fma(), which is what that BLAS does for n=2 on FMA hosts.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;
typedef int32_t i32;

#define ORACLE_API __attribute__((visibility("default")))

/* Worker threads of the OpenMP loops below (bench.py: torchrun exports OMP_NUM_THREADS=1 to every rank, which would
 * silently time the CPU arm on one core).  n <= 0 leaves the setting alone; returns the thread count in effect. */
#ifdef _OPENMP
#include <omp.h>
ORACLE_API int oracle_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
#else
ORACLE_API int oracle_set_threads(int n) { (void)n; return 1; }
#endif

/* ------------------------------------------------------------------------- */
/* a1  find_knn_within_radius                      src/utils.py:709-742       */
/* ------------------------------------------------------------------------- */
/* Inclusion: dx*dx + dy*dy <= r*r (cKDTree p=2 compares squared distances,     */
/* utils.py:722).  Rank: ascending distance (utils.py:726-  */
/* 728); the reference's order among equal distances is implementation-defined  */
/* (unsorted ball query + unstable argsort), ours is (d2, original ref index).  */
/* Keep the first min(knn, count).                                              */

typedef struct { double d2; i32 j; } cand_t;

static inline int cand_less(const cand_t *a, const cand_t *b) {
    return a->d2 < b->d2 || (a->d2 == b->d2 && a->j < b->j);
}

/* insert c into sorted top[0..*n) bounded by k */
static inline void topk_insert(cand_t *top, int *n, int k, cand_t c) {
    int m = *n;
    if (m == k) {
        if (!cand_less(&c, &top[k - 1])) return;
        m = k - 1;
    }
    int p = m;
    while (p > 0 && cand_less(&c, &top[p - 1])) { top[p] = top[p - 1]; --p; }
    top[p] = c;
    *n = m + 1;
}

/* brute force: O(Na*Nr); the independent check for the gridded version */
ORACLE_API int oracle_knn_brute(const double *a_xy, i64 na, const double *r_xy, i64 nr,
                                double radius, int knn, i32 *cand_j, i32 *cnt) {
    const double r2 = radius * radius;
    cand_t *top = (cand_t *)malloc(sizeof(cand_t) * (size_t)(knn > 0 ? knn : 1));
    for (i64 i = 0; i < na; ++i) {
        int n = 0;
        const double ax = a_xy[2 * i], ay = a_xy[2 * i + 1];
        for (i64 j = 0; j < nr; ++j) {
            const double dx = r_xy[2 * j] - ax, dy = r_xy[2 * j + 1] - ay;
            const double d2 = dx * dx + dy * dy;
            if (d2 <= r2) { cand_t c = {d2, (i32)j}; topk_insert(top, &n, knn, c); }
        }
        cnt[i] = n;
        for (int t = 0; t < knn; ++t) cand_j[i * knn + t] = t < n ? top[t].j : -1;
    }
    free(top);
    return 0;
}

/* gridded version (counting-sort bins of width >= radius); same results, O(N*m) */
ORACLE_API int oracle_knn_grid(const double *a_xy, i64 na, const double *r_xy, i64 nr,
                               double radius, int knn, i32 *cand_j, i32 *cnt) {
    if (na == 0) return 0;
    if (nr == 0) { for (i64 i = 0; i < na; ++i) { cnt[i] = 0; for (int t = 0; t < knn; ++t) cand_j[i * knn + t] = -1; } return 0; }
    double x0 = r_xy[0], x1 = r_xy[0], y0 = r_xy[1], y1 = r_xy[1];
    for (i64 j = 1; j < nr; ++j) {
        if (r_xy[2 * j] < x0) x0 = r_xy[2 * j];
        if (r_xy[2 * j] > x1) x1 = r_xy[2 * j];
        if (r_xy[2 * j + 1] < y0) y0 = r_xy[2 * j + 1];
        if (r_xy[2 * j + 1] > y1) y1 = r_xy[2 * j + 1];
    }
    double w = radius > 0 ? radius * 1.000001 : 1.0;
    const double maxbins = 2048.0;
    if ((x1 - x0) / w > maxbins) w = (x1 - x0) / maxbins;
    if ((y1 - y0) / w > maxbins) w = (y1 - y0) / maxbins;
    const i64 nbx = (i64)floor((x1 - x0) / w) + 1, nby = (i64)floor((y1 - y0) / w) + 1;
    i64 *start = (i64 *)calloc((size_t)(nbx * nby + 1), sizeof(i64));
    i32 *binof = (i32 *)malloc(sizeof(i32) * (size_t)nr);
    for (i64 j = 0; j < nr; ++j) {
        i64 bx = (i64)floor((r_xy[2 * j] - x0) / w), by = (i64)floor((r_xy[2 * j + 1] - y0) / w);
        if (bx >= nbx) bx = nbx - 1;
        if (by >= nby) by = nby - 1;
        binof[j] = (i32)(by * nbx + bx);
        start[binof[j] + 1]++;
    }
    for (i64 b = 0; b < nbx * nby; ++b) start[b + 1] += start[b];
    i32 *order = (i32 *)malloc(sizeof(i32) * (size_t)nr);
    i64 *fill = (i64 *)malloc(sizeof(i64) * (size_t)(nbx * nby));
    memcpy(fill, start, sizeof(i64) * (size_t)(nbx * nby));
    for (i64 j = 0; j < nr; ++j) order[fill[binof[j]]++] = (i32)j;
    const double r2 = radius * radius;
#pragma omp parallel
    {
        cand_t *top = (cand_t *)malloc(sizeof(cand_t) * (size_t)(knn > 0 ? knn : 1));
#pragma omp for schedule(dynamic, 256)
        for (i64 i = 0; i < na; ++i) {
            int n = 0;
            const double ax = a_xy[2 * i], ay = a_xy[2 * i + 1];
            const i64 cbx = (i64)floor((ax - x0) / w), cby = (i64)floor((ay - y0) / w);
            for (i64 by = cby - 1; by <= cby + 1; ++by) {
                if (by < 0 || by >= nby) continue;
                for (i64 bx = cbx - 1; bx <= cbx + 1; ++bx) {
                    if (bx < 0 || bx >= nbx) continue;
                    const i64 b = by * nbx + bx;
                    for (i64 s = start[b]; s < start[b + 1]; ++s) {
                        const i32 j = order[s];
                        const double dx = r_xy[2 * j] - ax, dy = r_xy[2 * j + 1] - ay;
                        const double d2 = dx * dx + dy * dy;
                        if (d2 <= r2) { cand_t c = {d2, j}; topk_insert(top, &n, knn, c); }
                    }
                }
            }
            cnt[i] = n;
            for (int t = 0; t < knn; ++t) cand_j[i * knn + t] = t < n ? top[t].j : -1;
        }
        free(top);
    }
    free(start); free(binof); free(order); free(fill);
    return 0;
}

/* utils.py:733-742: unique aligned / ref indices, frames re-indexed, pairs remapped.
 * keepA/keepR receive the ORIGINAL row numbers kept (ascending); pairs are in the new
 * index space, ordered by aligned row then rank.  Returns P. */
ORACLE_API i64 oracle_knn_compact(i64 na, i64 nr, int knn, const i32 *cand_j, const i32 *cnt,
                                  i32 *keepA, i64 *n_keepA, i32 *keepR, i64 *n_keepR, i32 *pairs) {
    i32 *newR = (i32 *)malloc(sizeof(i32) * (size_t)(nr > 0 ? nr : 1));
    for (i64 j = 0; j < nr; ++j) newR[j] = 0;
    for (i64 i = 0; i < na; ++i)
        for (int t = 0; t < cnt[i]; ++t) newR[cand_j[i * knn + t]] = 1;
    i64 kr = 0;
    for (i64 j = 0; j < nr; ++j) {
        if (newR[j]) { keepR[kr] = (i32)j; newR[j] = (i32)kr++; } else newR[j] = -1;
    }
    i64 ka = 0, p = 0;
    for (i64 i = 0; i < na; ++i) {
        if (cnt[i] == 0) continue;
        keepA[ka] = (i32)i;
        for (int t = 0; t < cnt[i]; ++t) { pairs[2 * p] = (i32)ka; pairs[2 * p + 1] = newR[cand_j[i * knn + t]]; ++p; }
        ++ka;
    }
    *n_keepA = ka; *n_keepR = kr;
    free(newR);
    return p;
}

/* ------------------------------------------------------------------------- */
/* a2  find_knn_with_cell_type_priority            src/knn_utils.py:31-65     */
/* ------------------------------------------------------------------------- */
/* Input pairs are a1's output (aligned-major, distance order).  In aligned     */
/* order: if the nearest ref has the same type and is not yet claimed, keep     */
/* only that pair and claim it; otherwise keep all pairs.  Returns new P.       */
ORACLE_API i64 oracle_knn_priority(const i32 *pairs, i64 p, const i32 *typeA, const i32 *typeR, i64 nr, i32 *out_pairs) {
    unsigned char *claimed = (unsigned char *)calloc((size_t)(nr > 0 ? nr : 1), 1);
    i64 q = 0, s = 0;
    while (s < p) {
        i64 e = s;
        while (e < p && pairs[2 * e] == pairs[2 * s]) ++e;
        const i32 i = pairs[2 * s], j0 = pairs[2 * s + 1];
        if (typeR[j0] == typeA[i] && !claimed[j0]) {
            out_pairs[2 * q] = i; out_pairs[2 * q + 1] = j0; ++q; claimed[j0] = 1;
        } else {
            for (i64 t = s; t < e; ++t) { out_pairs[2 * q] = pairs[2 * t]; out_pairs[2 * q + 1] = pairs[2 * t + 1]; ++q; }
        }
        s = e;
    }
    free(claimed);
    return q;
}

/* ------------------------------------------------------------------------- */
/* a3  pair cost                                   src/same.py:1182-1189      */
/* ------------------------------------------------------------------------- */
/* dist_ct = sum_k |A[i,k]-R[j,k]| left to right (object-dtype np.sum, SURVEY   */
/* App. A.3); dist_coords = |dx| + |dy|; c = ct_coeff*dist_ct + (ct_coeff*0.001)*dist_coords */
ORACLE_API int oracle_pair_cost(const i32 *pairs, i64 p, const double *a_xy, const double *r_xy,
                                const double *a_prob, const double *r_prob, int k, double dist_ct_coeff, double *cost) {
    const double dist_coeff = dist_ct_coeff * 0.001;
#pragma omp parallel for schedule(static)
    for (i64 e = 0; e < p; ++e) {
        const i64 i = pairs[2 * e], j = pairs[2 * e + 1];
        double s = 0.0;
        for (int c = 0; c < k; ++c) s = s + fabs(a_prob[i * k + c] - r_prob[j * k + c]);
        const double dc = fabs(a_xy[2 * i] - r_xy[2 * j]) + fabs(a_xy[2 * i + 1] - r_xy[2 * j + 1]);
        const double t1 = dist_ct_coeff * s, t2 = dist_coeff * dc;
        cost[e] = t1 + t2;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a4  constraint grouping                         src/helpers.py:105-138     */
/* ------------------------------------------------------------------------- */
/* ref_to_pairs: groups in FIRST-APPEARANCE order of j, pair indices ascending  */
/* inside a group; limit_g = multiplier*max_matches if ref size>1 (and the ref  */
/* frame has any size>1) else max_matches (helpers.py:121-135).  aligned_to_    */
/* pairs likewise (pairs are aligned-major so groups are contiguous runs, but   */
/* the general first-appearance rule is implemented).  Returns #groups.         */
ORACLE_API i64 oracle_group_pairs(const i32 *pairs, i64 p, int column, i64 n_nodes,
                                  i32 *group_node, i64 *indptr, i32 *idx) {
    i32 *gid = (i32 *)malloc(sizeof(i32) * (size_t)(n_nodes > 0 ? n_nodes : 1));
    for (i64 v = 0; v < n_nodes; ++v) gid[v] = -1;
    i64 g = 0;
    for (i64 e = 0; e < p; ++e) {
        const i32 v = pairs[2 * e + column];
        if (gid[v] < 0) { gid[v] = (i32)g; group_node[g] = v; ++g; }
    }
    for (i64 t = 0; t <= g; ++t) indptr[t] = 0;
    for (i64 e = 0; e < p; ++e) indptr[gid[pairs[2 * e + column]] + 1]++;
    for (i64 t = 0; t < g; ++t) indptr[t + 1] += indptr[t];
    i64 *fill = (i64 *)malloc(sizeof(i64) * (size_t)(g > 0 ? g : 1));
    memcpy(fill, indptr, sizeof(i64) * (size_t)g);
    for (i64 e = 0; e < p; ++e) idx[fill[gid[pairs[2 * e + column]]]++] = (i32)e;
    free(gid); free(fill);
    return g;
}

/* ------------------------------------------------------------------------- */
/* a5  _remap_triangles_by_vertex_ids              src/same.py:262-290        */
/* ------------------------------------------------------------------------- */
/* id_to_row = {v: i} (later rows win on duplicate ids); triangles with any     */
/* missing vertex are dropped; input order kept.  Returns T_out.                */
typedef struct { i64 v; i32 row; } vid_t;
static int vid_cmp(const void *a, const void *b) {
    const vid_t *x = (const vid_t *)a, *y = (const vid_t *)b;
    if (x->v != y->v) return x->v < y->v ? -1 : 1;
    return x->row < y->row ? -1 : (x->row > y->row);
}
ORACLE_API i64 oracle_remap_triangles(const i64 *tri, i64 t, const i64 *vertex_ids, i64 n, i32 *out_tri, i32 *out_src) {
    vid_t *tab = (vid_t *)malloc(sizeof(vid_t) * (size_t)(n > 0 ? n : 1));
    for (i64 i = 0; i < n; ++i) { tab[i].v = vertex_ids[i]; tab[i].row = (i32)i; }
    qsort(tab, (size_t)n, sizeof(vid_t), vid_cmp);
    i64 o = 0;
    for (i64 e = 0; e < t; ++e) {
        i32 r[3]; int ok = 1;
        for (int c = 0; c < 3 && ok; ++c) {
            const i64 v = tri[3 * e + c];
            i64 lo = 0, hi = n;              /* upper bound: first entry with id > v */
            while (lo < hi) { i64 m = (lo + hi) / 2; if (tab[m].v <= v) lo = m + 1; else hi = m; }
            if (lo == 0 || tab[lo - 1].v != v) ok = 0; else r[c] = tab[lo - 1].row;
        }
        if (ok) { out_tri[3 * o] = r[0]; out_tri[3 * o + 1] = r[1]; out_tri[3 * o + 2] = r[2]; if (out_src) out_src[o] = (i32)e; ++o; }
    }
    free(tab);
    return o;
}

/* ------------------------------------------------------------------------- */
/* a6  filter_triangles_by_radius                  src/helpers.py:233-395     */
/* ------------------------------------------------------------------------- */
static inline double norm2(double x, double y) { return sqrt(fma(y, y, x * x)); }   /* helpers.py:282-283,305-307 */
/* compute_angle(p1,p2,p3): angle at p2, helpers.py:278-288 */
static inline double angle_at(double p1x, double p1y, double p2x, double p2y, double p3x, double p3y) {
    const double v1x = p1x - p2x, v1y = p1y - p2y, v2x = p3x - p2x, v2y = p3y - p2y;
    const double n1 = norm2(v1x, v1y), n2 = norm2(v2x, v2y);
    if (n1 == 0 || n2 == 0) return 0.0;
    double c = fma(v1y, v2y, v1x * v2x) / (n1 * n2);
    if (c < -1.0) c = -1.0;
    if (c > 1.0) c = 1.0;
    return acos(c) * (180.0 / M_PI);
}

/* class codes shared with the CUDA library (include/same_b200.h) */
enum { TRI_DROP_RADIUS = 0, TRI_DROP_ANGLE = 1, TRI_SAME_TYPE = 2, TRI_KEEP = 3 };

/* Per-triangle classification + perimeter; `band` (optional) marks triangles whose
 * decision sits within a few ulps of a threshold (SURVEY.md §7 hard part 1). */
ORACLE_API int oracle_tri_classify(const double *xy, const i32 *tri, i64 t, double radius, int use_angle, double min_angle_deg,
                                   const i32 *type, int ignore_same_type, unsigned char *cls, double *score, unsigned char *band) {
#pragma omp parallel for schedule(static)
    for (i64 e = 0; e < t; ++e) {
        const i32 a = tri[3 * e], b = tri[3 * e + 1], c = tri[3 * e + 2];
        const double p1x = xy[2 * a], p1y = xy[2 * a + 1], p2x = xy[2 * b], p2y = xy[2 * b + 1], p3x = xy[2 * c], p3y = xy[2 * c + 1];
        const double s1 = norm2(p2x - p1x, p2y - p1y), s2 = norm2(p3x - p2x, p3y - p2y), s3 = norm2(p1x - p3x, p1y - p3y);
        double mx = s1 > s2 ? s1 : s2; if (s3 > mx) mx = s3;
        unsigned char k, bd = fabs(mx - radius) <= 1e-12 * fabs(radius);
        score[e] = (s1 + s2) + s3;                                    /* helpers.py:333 */
        if (mx >= radius) k = TRI_DROP_RADIUS;                         /* helpers.py:310 */
        else {
            k = TRI_KEEP;
            if (use_angle) {                                           /* helpers.py:315-321 */
                const double a1 = angle_at(p2x, p2y, p1x, p1y, p3x, p3y);
                const double a2 = angle_at(p1x, p1y, p2x, p2y, p3x, p3y);
                const double a3 = angle_at(p1x, p1y, p3x, p3y, p2x, p2y);
                double mn = a1 < a2 ? a1 : a2; if (a3 < mn) mn = a3;
                if (fabs(mn - min_angle_deg) <= 1e-9) bd = 1;
                if (mn < min_angle_deg) k = TRI_DROP_ANGLE;
            }
            if (k == TRI_KEEP && ignore_same_type && type && type[a] == type[b] && type[b] == type[c]) k = TRI_SAME_TYPE; /* :328-330 */
        }
        cls[e] = k;
        if (band) band[e] = bd;
    }
    return 0;
}

/* The sequential bookkeeping of helpers.py:296-393 on top of the classification.
 * kept_src: indices into the input triangle list, in output order (input order, then
 * add-backs in ascending node order, helpers.py:365-383).  node_valid[v]=1 iff v has a
 * radius+angle-valid triangle (complement = "truly unconstrained", helpers.py:355-356). */
ORACLE_API i64 oracle_tri_select(const i32 *tri, i64 t, i64 n_points, const unsigned char *cls, const double *score,
                                 int ignore_same_type, int ensure_min, i32 *kept_src, unsigned char *node_valid) {
    unsigned char *has_tri = (unsigned char *)calloc((size_t)(n_points > 0 ? n_points : 1), 1);
    i32 *best = (i32 *)malloc(sizeof(i32) * (size_t)(n_points > 0 ? n_points : 1));
    unsigned char *added = (unsigned char *)calloc((size_t)(t > 0 ? t : 1), 1);
    for (i64 v = 0; v < n_points; ++v) { best[v] = -1; node_valid[v] = 0; }
    i64 o = 0;
    for (i64 e = 0; e < t; ++e) {
        if (cls[e] < TRI_SAME_TYPE) continue;
        for (int c = 0; c < 3; ++c) node_valid[tri[3 * e + c]] = 1;
        if (cls[e] == TRI_SAME_TYPE) {
            if (ensure_min)
                for (int c = 0; c < 3; ++c) {
                    const i32 v = tri[3 * e + c];
                    if (best[v] < 0 || score[e] < score[best[v]]) best[v] = (i32)e;   /* strict <, first wins: :335-339 */
                }
            continue;
        }
        kept_src[o++] = (i32)e; added[e] = 1;
        for (int c = 0; c < 3; ++c) has_tri[tri[3 * e + c]] = 1;
    }
    if (ignore_same_type && ensure_min)
        for (i64 v = 0; v < n_points; ++v) {
            if (has_tri[v] || !node_valid[v]) continue;
            const i32 e = best[v];
            if (e < 0 || added[e]) continue;
            /* helpers.py:379 de-duplicates by vertex tuple; identical tuples share score and
             * vertices, so the first such triangle wins every node and index-dedup is equivalent. */
            kept_src[o++] = e; added[e] = 1;
        }
    free(has_tri); free(best); free(added);
    return o;
}

/* ------------------------------------------------------------------------- */
/* a9  triangle weights + source signs             src/same.py:1128-1146      */
/* ------------------------------------------------------------------------- */
static inline double sgn(double v) { return v > 0 ? 1.0 : (v < 0 ? -1.0 : 0.0); }
static inline double orient(double ax, double ay, double bx, double by, double cx, double cy) {
    const double t1 = (bx - ax) * (cy - ay), t2 = (by - ay) * (cx - ax);
    return t1 - t2;                                                     /* same.py:658,1146 */
}
ORACLE_API int oracle_tri_tables(const double *xy, const double *size, const i32 *tri, i64 t, double *weight, signed char *sign,
                                 double *bounds, i32 *argv) {
#pragma omp parallel for schedule(static)
    for (i64 e = 0; e < t; ++e) {
        const i32 v[3] = {tri[3 * e], tri[3 * e + 1], tri[3 * e + 2]};
        const double x[3] = {xy[2 * v[0]], xy[2 * v[1]], xy[2 * v[2]]}, y[3] = {xy[2 * v[0] + 1], xy[2 * v[1] + 1], xy[2 * v[2] + 1]};
        if (weight) weight[e] = (size[v[0]] + size[v[1]]) + size[v[2]];
        sign[e] = (signed char)sgn(orient(x[0], y[0], x[1], y[1], x[2], y[2]));
        if (bounds) {                                                   /* helpers.py:184-210 */
            double mnx = x[0], mxx = x[0], mny = y[0], mxy = y[0];
            for (int c = 1; c < 3; ++c) { if (x[c] < mnx) mnx = x[c]; if (x[c] > mxx) mxx = x[c]; if (y[c] < mny) mny = y[c]; if (y[c] > mxy) mxy = y[c]; }
            bounds[4 * e] = mnx; bounds[4 * e + 1] = mxx; bounds[4 * e + 2] = mny; bounds[4 * e + 3] = mxy;
            int amx = -1, amn = -1, amy = -1, any_ = -1;                /* first vertex attaining the bound */
            for (int c = 2; c >= 0; --c) { if (x[c] == mxx) amx = c; if (x[c] == mnx) amn = c; if (y[c] == mxy) amy = c; if (y[c] == mny) any_ = c; }
            argv[4 * e] = v[amx]; argv[4 * e + 1] = v[amn]; argv[4 * e + 2] = v[amy]; argv[4 * e + 3] = v[any_];
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a10 _lazy_orientation_callback                  src/same.py:621-703        */
/* ------------------------------------------------------------------------- */
/* matching[i] = j and match_pair_idx[i] = idx of the LAST pair of row i with   */
/* x > 0.5 (dict overwrite, same.py:636-639).                                   */
ORACLE_API int oracle_matching_from_x(const double *x, const i32 *pairs, i64 p, i64 na, i32 *match_j, i32 *match_p) {
    for (i64 i = 0; i < na; ++i) { match_j[i] = -1; match_p[i] = -1; }
    for (i64 e = 0; e < p; ++e)
        if (x[e] > 0.5) { match_j[pairs[2 * e]] = pairs[2 * e + 1]; match_p[pairs[2 * e]] = (i32)e; }
    return 0;
}
/* Returns V (all violated triangles, ascending); *checked as in same.py:667. */
ORACLE_API i64 oracle_separation(const i32 *tri, i64 t, const signed char *source_sign, const i32 *match_j,
                                 const double *r_xy, i32 *viol, i64 *checked) {
    i64 v = 0, ck = 0;
    for (i64 e = 0; e < t; ++e) {
        const i32 ja = match_j[tri[3 * e]], jb = match_j[tri[3 * e + 1]], jc = match_j[tri[3 * e + 2]];
        if (ja < 0 || jb < 0 || jc < 0) continue;
        const double s = sgn(orient(r_xy[2 * ja], r_xy[2 * ja + 1], r_xy[2 * jb], r_xy[2 * jb + 1], r_xy[2 * jc], r_xy[2 * jc + 1]));
        if (source_sign[e] == 0 || s == 0) continue;
        ++ck;
        if ((double)source_sign[e] != s) viol[v++] = (i32)e;
    }
    *checked = ck;
    return v;
}

/* ------------------------------------------------------------------------- */
/* a11 verify_spatial_preservation                 src/violationhelper.py:1-134 */
/* a12 signed areas + flips       src/same.py:1355-1408, src/helpers.py:73-77   */
/* ------------------------------------------------------------------------- */
/* mask bit layout per triangle: bits0-2 x-order violation of vertex pairs      */
/* (0,1),(0,2),(1,2) in VERTEX POSITION terms; bits3-5 y-order; bits 8-10       */
/* vertex matched.  The reference enumerates pairs over the matched-vertex      */
/* sub-list, which preserves this relative order.                               */
static inline double signed_area(double x1, double y1, double x2, double y2, double x3, double y3) {
    return 0.5 * ((x1 * (y2 - y3) + x2 * (y3 - y1)) + x3 * (y1 - y2));   /* helpers.py:77 */
}
ORACLE_API int oracle_postsolve(const i32 *tri, i64 t, const double *a_xy, const double *r_xy, const i32 *match_j,
                                i32 *mask, double *area_before, double *area_after, unsigned char *flipped) {
#pragma omp parallel for schedule(static)
    for (i64 e = 0; e < t; ++e) {
        const i32 v[3] = {tri[3 * e], tri[3 * e + 1], tri[3 * e + 2]};
        const i32 j[3] = {match_j[v[0]], match_j[v[1]], match_j[v[2]]};
        i32 m = 0;
        for (int c = 0; c < 3; ++c) if (j[c] >= 0) m |= 1 << (8 + c);
        static const int P[3][2] = {{0, 1}, {0, 2}, {1, 2}};
        for (int q = 0; q < 3; ++q) {
            const int u = P[q][0], w = P[q][1];
            if (j[u] < 0 || j[w] < 0) continue;
            const int ox = a_xy[2 * v[u]] < a_xy[2 * v[w]], oy = a_xy[2 * v[u] + 1] < a_xy[2 * v[w] + 1];
            const int mx = r_xy[2 * j[u]] < r_xy[2 * j[w]], my = r_xy[2 * j[u] + 1] < r_xy[2 * j[w] + 1];
            if (ox != mx) m |= 1 << q;
            if (oy != my) m |= 1 << (3 + q);
        }
        mask[e] = m;
        const double ab = signed_area(a_xy[2 * v[0]], a_xy[2 * v[0] + 1], a_xy[2 * v[1]], a_xy[2 * v[1] + 1], a_xy[2 * v[2]], a_xy[2 * v[2] + 1]);
        area_before[e] = ab;
        if (j[0] >= 0 && j[1] >= 0 && j[2] >= 0) {
            const double aa = signed_area(r_xy[2 * j[0]], r_xy[2 * j[0] + 1], r_xy[2 * j[1]], r_xy[2 * j[1] + 1], r_xy[2 * j[2]], r_xy[2 * j[2] + 1]);
            area_after[e] = aa;
            flipped[e] = (ab * aa < 0);                                 /* same.py:1398 */
        } else { area_after[e] = NAN; flipped[e] = 0; }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a13 subset_data                                 src/same.py:293-295        */
/* ------------------------------------------------------------------------- */
ORACLE_API i64 oracle_subset(const double *xy, i64 n, double x_min, double x_max, double y_min, double y_max, i32 *rows) {
    i64 o = 0;
    for (i64 i = 0; i < n; ++i)
        if (xy[2 * i] >= x_min && xy[2 * i] < x_max && xy[2 * i + 1] >= y_min && xy[2 * i + 1] < y_max) rows[o++] = (i32)i;
    return o;
}

/* ------------------------------------------------------------------------- */
/* f3/f2  ordered greedy selection with disjoint endpoints                   */
/*        greedy MIP start                     src/init_helpers.py:110-132   */
/*        batch selection of triangle collapse src/metacell_utils.py:423-433 */
/* ------------------------------------------------------------------------- */
typedef struct { double key; i64 idx; } gitem_t;
static int gitem_cmp(const void *a, const void *b) {
    const gitem_t *x = (const gitem_t *)a, *y = (const gitem_t *)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);   /* list.sort is stable: ties stay in index order */
}
/* Items in ascending (key, index) order; an item is selected iff it is eligible (eligible == NULL: all) and none of its
 * `degree` endpoints was used by an earlier selected item.  selected[n], used[n_nodes] out.  Returns the number selected. */
ORACLE_API i64 oracle_greedy_select(i64 n, int degree, const i32 *nodes, const double *key, const unsigned char *eligible, i64 n_nodes,
                                    unsigned char *selected, unsigned char *used) {
    gitem_t *it = (gitem_t *)malloc(sizeof(gitem_t) * (size_t)(n > 0 ? n : 1));
    for (i64 e = 0; e < n; ++e) { it[e].key = key[e]; it[e].idx = e; }
    qsort(it, (size_t)n, sizeof(gitem_t), gitem_cmp);
    memset(selected, 0, (size_t)n);
    memset(used, 0, (size_t)n_nodes);
    i64 count = 0;
    for (i64 r = 0; r < n; ++r) {
        const i64 e = it[r].idx;
        if (eligible && !eligible[e]) continue;
        int hit = 0;
        for (int d = 0; d < degree; ++d) hit |= used[nodes[e * degree + d]];
        if (hit) continue;
        selected[e] = 1;
        for (int d = 0; d < degree; ++d) used[nodes[e * degree + d]] = 1;
        ++count;
    }
    free(it);
    return count;
}

/* f2  candidate test + perimeter of one collapse iteration   src/metacell_utils.py:388-409
 * np.linalg.norm(1-D) = sqrt(ddot(v, v)); ddot with n = 2 on FMA hosts = fma(vy, vy, vx*vx);
 * the three norms are added left to right. */
static double norm2_ref(double vx, double vy) { return sqrt(fma(vy, vy, vx * vx)); }
ORACLE_API void oracle_collapse_score(i64 n_tri, const i32 *tri, const double *xy, const i32 *type, const double *size, double max_size,
                                      unsigned char *cand, double *perim) {
    for (i64 t = 0; t < n_tri; ++t) {
        const i32 a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        const int same = type[a] == type[b] && type[b] == type[c];
        const double total = size[a] + size[b] + size[c];
        const double ax = xy[2 * a], ay = xy[2 * a + 1], bx = xy[2 * b], by = xy[2 * b + 1], cx = xy[2 * c], cy = xy[2 * c + 1];
        perim[t] = (norm2_ref(ax - bx, ay - by) + norm2_ref(bx - cx, by - cy)) + norm2_ref(cx - ax, cy - ay);
        cand[t] = same && !(total > max_size);
    }
}

/* f2  member means of merged metacells                 src/metacell_utils.py:446-474
 * rows[col].mean() = pandas nanmean = numpy pairwise sum / count (verified against pandas in tests/test_oracle_golden.py). */
static double pairwise_sum_ref(const double *v, i64 C, i64 c, const i32 *pos, i64 n) {
    if (n < 8) {
        double r = 0.0;
        for (i64 i = 0; i < n; ++i) r += v[(i64)pos[i] * C + c];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = v[(i64)pos[k] * C + c];
        i64 i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += v[(i64)pos[i + k] * C + c];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += v[(i64)pos[i] * C + c];
        return res;
    }
    i64 n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum_ref(v, C, c, pos, n2) + pairwise_sum_ref(v, C, c, pos + n2, n - n2);
}
ORACLE_API void oracle_segment_mean(i64 n_rows, i64 C, const double *values, i64 G, const i64 *ptr, const i32 *pos, double *out) {
    (void)n_rows;
    for (i64 g = 0; g < G; ++g)
        for (i64 c = 0; c < C; ++c) {
            const i64 n = ptr[g + 1] - ptr[g];
            out[g * C + c] = n > 0 ? pairwise_sum_ref(values, C, c, pos + ptr[g], n) / (double)n : NAN;
        }
}
