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
// reference cells per bin: the own bin plus ring 1 (nine bins) should hold the k nearest with room to spare, and no more —
// smaller bins prune better (measured at knn = 8 on the 1 M-cell section with the bucketed top-k: 4/bin 194 us, 5/bin 182 us,
// 6/bin 179 us, 8/bin 192 us, 10/bin 199 us)
static double target_per_bin(int knn) {
    static const char *env = getenv("SAME_B200_BIN_TARGET");   // tuning knob for tools/dev_knn_sweep.sh (cells per bin = value * knn / 8)
    const double scale = env ? atof(env) / 8.0 : 0.75;
    return std::max(scale > 0.0 ? 2.0 : 4.0, (scale > 0.0 ? scale : 0.75) * knn);
}
constexpr int MAX_BINS_AXIS = 2048;

// ---- binning -------------------------------------------------------------------------
__device__ __forceinline__ void bin_of(const GridParams &g, double2 p, int &bx, int &by) {
    bx = (int)floor((p.x - g.x0) * g.inv_w);
    by = (int)floor((p.y - g.y0) * g.inv_w);
    bx = min(max(bx, 0), g.nbx - 1);
    by = min(max(by, 0), g.nby - 1);
}

// Counting sort of BOTH frames' window instances by bin (item < nAi: aligned instance, else reference instance
// item - nAi): pass 1 computes the bin of every instance and counts bin populations (counter [frame][bin]); the
// counters are scanned in one launch (scan.cuh); pass 2 scatters (xy, instance) to start[bin] + a slot taken from the
// same counter.  The order of the cells INSIDE a bin is arbitrary; the search ranks candidates by the total order
// (d2, reference instance), so its result does not depend on it.
__global__ void k_bin_count(const i32 *__restrict__ a_src, const i32 *__restrict__ r_src, const double2 *__restrict__ a_xy,
                            const double2 *__restrict__ r_xy, i64 nAi, i64 nRi, const i32 *__restrict__ a_off, const i32 *__restrict__ r_off, int W,
                            const GridParams *__restrict__ grids, i64 nb1, i32 *__restrict__ bkey, i32 *__restrict__ bcnt) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nAi + nRi) return;
    const bool ref = i >= nAi;
    const i32 inst = (i32)(ref ? i - nAi : i);
    const GridParams &g = grids[find_window(ref ? r_off : a_off, W, inst)];
    int bx, by;
    bin_of(g, ref ? r_xy[r_src[inst]] : a_xy[a_src[inst]], bx, by);
    const i32 key = g.base + by * g.nbx + bx;
    bkey[i] = key;
    atomicAdd(bcnt + (ref ? nb1 : 0) + key, 1);
}

__global__ void k_bin_scatter(const i32 *__restrict__ a_src, const i32 *__restrict__ r_src, const double2 *__restrict__ a_xy,
                              const double2 *__restrict__ r_xy, i64 nAi, i64 nRi, i64 nb1, const i32 *__restrict__ bkey,
                              const i32 *__restrict__ bstart, i32 *__restrict__ bcnt, double2 *__restrict__ sorted_xy, i32 *__restrict__ sorted_inst) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nAi + nRi) return;
    const bool ref = i >= nAi;
    const i32 inst = (i32)(ref ? i - nAi : i);
    const i64 c = (ref ? nb1 : 0) + bkey[i];
    const i32 dst = bstart[c] + atomicSub(bcnt + c, 1) - 1;
    sorted_xy[dst] = ref ? r_xy[r_src[inst]] : a_xy[a_src[inst]];
    sorted_inst[dst] = inst;
}

// out[i] = sum of in[0 .. i-1] for i in [0, n): one launch.  A thread owns 16 consecutive values (four 16-byte loads);
// 256-thread blocks keep eight tiles per SM in flight, so one tile's look-back wait overlaps the others' loads.
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 16;
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_i32(const i32 *__restrict__ in, i32 *__restrict__ out, i64 n, ScanCtx sc) {
    __shared__ int smem[SCAN_THREADS / 32 + 1];
    const i64 base = ((i64)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
    int v[SCAN_ITEMS], sum[1] = {0};
    if (base + SCAN_ITEMS <= n) {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k += 4) {
            const int4 q = *reinterpret_cast<const int4 *>(in + base + k);
            v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = base + k < n ? in[base + k] : 0;
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) sum[0] += v[k];
    int excl[1], tot[1], pre[1];
    device_exclusive_scan<1, SCAN_THREADS>(sc, (int)blockIdx.x, sum, excl, tot, pre, smem);
    int run = excl[0];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { const int x = v[k]; v[k] = run; run += x; }
    if (base + SCAN_ITEMS <= n) {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k += 4) *reinterpret_cast<int4 *>(out + base + k) = make_int4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < n) out[base + k] = v[k];
    }
}
void scan_i32(Section *sec, const i32 *in, i32 *out, i64 n, cudaStream_t s) {
    if (n <= 0) return;
    const unsigned tiles = blocks_for(n, SCAN_THREADS * SCAN_ITEMS);
    LAUNCH(k_scan_i32, tiles, SCAN_THREADS, 0, s, in, out, n, scan_ctx(sec, tiles, 1, s));
}

// ---- top-k search ----------------------------------------------------------------------
// (d, j) < (bd, bj) lexicographically; bitwise ops so that the compiler emits predicate logic, not branches (the
// short-circuit form compiled to divergent branches inside the insertion: 180 SASS instructions at ~3 active lanes)
__device__ __forceinline__ bool cand_less(double d, i32 j, double bd, i32 bj) { return (d < bd) | ((d == bd) & (j < bj)); }

__device__ __forceinline__ void ring_slot(int ring, int slot, int &dx, int &dy) {   // perimeter walk: four sides of 2*ring bins
    const int side = slot / (2 * ring), k = slot - side * 2 * ring;
    dx = side == 0 ? -ring + k : (side == 1 ? ring : (side == 2 ? ring - k : -ring));
    dy = side == 0 ? -ring : (side == 1 ? -ring + k : (side == 2 ? ring : ring - k));
}

// Exact search of ONE query with the (d2, ref instance) list in local memory: the path a query takes when the bucketed
// search below cannot decide its k-th neighbour (two candidates in the same 2^-20-relative distance bucket at the
// boundary — in practice only on exact lattices).  Writes the query's outputs itself.
__device__ __noinline__ void knn_exact_one(double2 q, const GridParams &g, int cbx, int cby, const i32 *__restrict__ bin_start,
                                           const double2 *__restrict__ sr_xy, const i32 *__restrict__ sr_inst, double r2, int knn,
                                           i32 *__restrict__ out, i32 *__restrict__ cnt_out, i32 *__restrict__ r_used) {
    double ld[32];
    i32 lj[32];
    const double px = q.x - g.x0, py = q.y - g.y0;
    const double eps = 1e-7 * g.w;
    const double fx = px - cbx * g.w, fy = py - cby * g.w;
    const double edge = fmin(fmin(fx, g.w - fx), fmin(fy, g.w - fy)) - eps;
    int found = 0;
    double tau_d = INFINITY;
    i32 tau_j = 0x7fffffff;
    for (int ring = 0; ring <= g.rings; ++ring) {
        if (ring > 0) {
            const double gap = (ring - 1) * g.w + edge;
            if (gap > 0.0 && gap * gap > fmin(tau_d, r2)) break;
        }
        const int n_slots = ring == 0 ? 1 : 8 * ring;
        for (int slot = 0; slot < n_slots; ++slot) {
            int dx = 0, dy = 0;
            if (ring > 0) ring_slot(ring, slot, dx, dy);
            const int bx = cbx + dx, by = cby + dy;
            if (bx < 0 || bx >= g.nbx || by < 0 || by >= g.nby) continue;
            const double gx = fmax(0.0, fmax(bx * g.w - px, px - (bx + 1) * g.w) - eps);
            const double gy = fmax(0.0, fmax(by * g.w - py, py - (by + 1) * g.w) - eps);
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
                while (u > 0 && cand_less(d2, j, ld[u - 1], lj[u - 1])) { ld[u] = ld[u - 1]; lj[u] = lj[u - 1]; --u; }
                ld[u] = d2; lj[u] = j;
                if (found < knn) ++found;
                if (found == knn) { tau_d = ld[knn - 1]; tau_j = lj[knn - 1]; }
            }
        }
    }
    *cnt_out = found;
    for (int u = 0; u < knn; ++u) {
        out[u] = (u < found) ? lj[u] : -1;
        if (u < found) r_used[lj[u]] = 1;
    }
}

// One thread per aligned instance.  The running top-k is kept on a 32-bit KEY — the high word of the fp64 d2, a monotone
// 2^-20-relative bucket of the distance — with the candidate's position in the sorted array beside it: an insertion is
// one integer compare and four selects per slot instead of a (double, index) lexicographic compare and six selects
// (the fp64 list spent 52 % of the kernel's warp instructions inserting at 11 of 32 active lanes, profiles/r1k).
// The k candidates with the smallest keys ARE the k nearest unless the k-th and a dropped candidate share a bucket
// (`tie`): those queries take knn_exact_one.  The survivors' exact d2 are recomputed once at the end and put in
// (d2, ref instance) order by an odd-even transposition pass over an almost sorted list.
constexpr unsigned KEY_NONE = 0x7ff00000u;   // high word of +inf: above the key of every finite d2
template <int KCAP>
__global__ void __launch_bounds__(128) k_knn(const double2 *__restrict__ sa_xy, const i32 *__restrict__ sa_inst, i64 nAi,
                                             const i32 *__restrict__ a_off, int W, const GridParams *__restrict__ grids,
                                             const i32 *__restrict__ bin_start, const double2 *__restrict__ sr_xy,
                                             const i32 *__restrict__ sr_inst, double r2, int knn, i32 *__restrict__ cand,
                                             i32 *__restrict__ cnt, i32 *__restrict__ r_used, unsigned long long *__restrict__ evals) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nAi) return;
    const i32 inst = sa_inst[t];
    const double2 q = sa_xy[t];
    const GridParams g = grids[find_window(a_off, W, inst)];
    int cbx, cby;
    bin_of(g, q, cbx, cby);
    const double px = q.x - g.x0, py = q.y - g.y0;
    int nev = 0;   // distance evaluations of this query (only reported when profiling: the compute-side roofline of bench.py)

    // Sorted keys in registers, right-aligned: slots [KCAP-knn, KCAP) hold the list (ascending), the slots before it hold
    // zero keys that never move, so the k-th best is always the LAST slot (a static register index; a runtime index would push
    // the array into local memory).  A key = distance bucket (high word of d2 without its SB lowest bits, + one bucket so that
    // no real key is below the zero keys) | id of the POSITION SLOT that remembers where the candidate sits in the sorted
    // array.  Keys are all distinct, so an insertion is two integer min/max per slot; the evicted key's slot id is the free
    // position slot, written with one shared-memory store ([slot][thread], conflict-free).
    constexpr int SB = KCAP <= 4 ? 2 : (KCAP <= 8 ? 3 : (KCAP <= 16 ? 4 : 5));
    constexpr unsigned SLOT = (1u << SB) - 1u, ONE = 1u << SB, NONE_B = KEY_NONE + ONE;
    __shared__ i32 spos[KCAP * 128];
    unsigned bk[KCAP];
    const int head = KCAP - knn;
#pragma unroll
    for (int s = 0; s < KCAP; ++s) bk[s] = ((s < head) ? 0u : NONE_B) | (unsigned)s;
#define last_k bk[KCAP - 1]
    bool tie = false;
    // Candidates with d2 > thr cannot enter: thr = min(r2, upper edge of the k-th key's bucket), kept as its two words (the
    // upper edge has the bucket field of last_k as its high word — the one-bucket offset cancels — and a zero low word; it is
    // below r2 exactly when that high word is <= the high word of r2).
    const unsigned r2_hi = (unsigned)__double2hiint(r2), r2_lo = (unsigned)__double2loint(r2);
    double thr = r2;
    // Bin pruning runs in fp32 on lower bounds of the point-to-bin distance: fx, fy (position inside the own bin) and the
    // bin width are O(w), so fp32 rounding is ~1e-7 w per term; 1e-5 w of slack on every gap and thr rounded up keep the
    // test conservative.  The exact predicate is always the fp64 one on the candidate itself.
    const float wf = (float)g.w, epsf = 1e-5f * wf;
    const float fxf = (float)(px - cbx * g.w), fyf = (float)(py - cby * g.w);
    float thrf = __double2float_ru(thr);
    // The own bin and ring 1 — where nearly all the work is — are nine bins whose squared gaps are sums of four numbers: the
    // gap to the NEAR neighbour column / row (the one on the query's side of its bin) and to the FAR one.  They are walked
    // nearest first (own, near column, near row, far column, far row, then the corners), so
    // the k-th best tightens early and the far bins are pruned without an evaluation.  Rings >= 2 use the general walker.
    const int sx = (fxf + fxf < wf) ? -1 : 1, sy = (fyf + fyf < wf) ? -1 : 1;
    const float nxg = fmaxf(0.f, fminf(fxf, wf - fxf) - epsf), fxg = fmaxf(0.f, fmaxf(fxf, wf - fxf) - epsf);
    const float nyg = fmaxf(0.f, fminf(fyf, wf - fyf) - epsf), fyg = fmaxf(0.f, fmaxf(fyf, wf - fyf) - epsf);
    const float nx2 = nxg * nxg, fx2 = fxg * fxg, ny2 = nyg * nyg, fy2 = fyg * fyg;
    const float edge = fminf(nxg, nyg);                     // distance to the nearest side of the own bin (lower bound)
    // One bin: two candidates per trip — both loads are in flight before either is used (the second one re-reads the first
    // when the bin ends on an odd count and is then skipped).
    auto visit = [&](i32 b) {
        const i32 s1 = bin_start[b + 1];
        for (i32 s0 = bin_start[b]; s0 < s1; s0 += 2) {
            const bool two = s0 + 1 < s1;
            nev += two ? 2 : 1;
            const double2 pA = sr_xy[s0], pB = sr_xy[two ? s0 + 1 : s0];
            const double ax = __dsub_rn(pA.x, q.x), ay = __dsub_rn(pA.y, q.y), bx2 = __dsub_rn(pB.x, q.x), by2 = __dsub_rn(pB.y, q.y);
            const double dA = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
            const double dB = two ? __dadd_rn(__dmul_rn(bx2, bx2), __dmul_rn(by2, by2)) : INFINITY;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double d2 = h ? dB : dA;
                const i32 s = s0 + h;
                if (d2 <= thr) {
                    const unsigned hb = ((unsigned)__double2hiint(d2) & ~SLOT) + ONE, lb = last_k & ~SLOT;
                    if (hb < lb) {
                        const unsigned free_slot = last_k & SLOT;           // the evicted entry's position slot
                        const unsigned nk = hb | free_slot;
                        spos[free_slot * 128 + threadIdx.x] = s;
                        // sorted insertion of a key that differs from every key in the list (top down: each slot reads the
                        // OLD value of its left neighbour): slot u keeps its key, takes the new one, or takes slot u-1's
#pragma unroll
                        for (int u = KCAP - 1; u > 0; --u) bk[u] = min(bk[u], max(bk[u - 1], nk));
                        bk[0] = min(bk[0], nk);
                        const unsigned nb = last_k & ~SLOT;
                        tie = nb == lb;   // the dropped candidate shares the new k-th's bucket (or the list is not full yet)
                        const bool below = nb <= r2_hi;   // never for the empty-slot key: r2 is finite
                        thr = __hiloint2double((int)(below ? nb : r2_hi), (int)(below ? 0u : r2_lo));
                        thrf = __double2float_ru(thr);
                    } else if (hb == lb) {
                        tie = true;
                    }
                }
            }
        }
    };
    // own bin, then ring 1 nearest first — (N,0), (0,N), (F,0), (0,F), (N,N), (N,F), (F,N), (F,F) with N / F the near / far
    // neighbour column or row — written out so that the bin offsets, the bounds tests and the gaps are a few registers
    // computed once (a table-driven loop spent ~50 instructions per bin on decoding)
    const i32 b0 = g.base + cby * g.nbx + cbx, dyn = sy * g.nbx;
    const bool xn = (unsigned)(cbx + sx) < (unsigned)g.nbx, xf = (unsigned)(cbx - sx) < (unsigned)g.nbx;
    const bool yn = (unsigned)(cby + sy) < (unsigned)g.nby, yf = (unsigned)(cby - sy) < (unsigned)g.nby;
    visit(b0);
    if (!(edge * edge > thrf)) {                             // else nothing outside the own bin can enter any more
        if (xn && !(nx2 > thrf)) visit(b0 + sx);
        if (yn && !(ny2 > thrf)) visit(b0 + dyn);
        if (xf && !(fx2 > thrf)) visit(b0 - sx);
        if (yf && !(fy2 > thrf)) visit(b0 - dyn);
        if (xn && yn && !(nx2 + ny2 > thrf)) visit(b0 + sx + dyn);
        if (xn && yf && !(nx2 + fy2 > thrf)) visit(b0 + sx - dyn);
        if (xf && yn && !(fx2 + ny2 > thrf)) visit(b0 - sx + dyn);
        if (xf && yf && !(fx2 + fy2 > thrf)) visit(b0 - sx - dyn);
        for (int ring = 2; ring <= g.rings; ++ring) {        // general walker
            // nearest possible point of this ring (Chebyshev distance `ring` bins from the centre bin)
            const float gap = (float)(ring - 1) * wf + edge;
            if (gap * gap > thrf) break;
            for (int slot = 0; slot < 8 * ring; ++slot) {
                int dx, dy;
                ring_slot(ring, slot, dx, dy);
                // gap along one axis to a bin d bins away: d > 0: d*w - f;  d < 0: f + (-d - 1)*w;  own row/column: 0
                const float gx = dx == 0 ? 0.f : fmaxf(0.f, (dx > 0 ? (float)dx * wf - fxf : fxf - (float)(dx + 1) * wf) - epsf);
                const float gy = dy == 0 ? 0.f : fmaxf(0.f, (dy > 0 ? (float)dy * wf - fyf : fyf - (float)(dy + 1) * wf) - epsf);
                const int bx = cbx + dx, by = cby + dy;
                if (bx < 0 || bx >= g.nbx || by < 0 || by >= g.nby) continue;
                if (gx * gx + gy * gy > thrf) continue;
                visit(g.base + by * g.nbx + bx);
            }
        }
    }
    if (evals) {   // one atomic per warp (the lanes have reconverged behind the walk)
        const unsigned live = __activemask();
        const unsigned tot = __reduce_add_sync(live, (unsigned)nev);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(live) - 1)) atomicAdd(evals, (unsigned long long)tot);
    }
    i32 *out = cand + (i64)inst * knn;
    if (tie && (last_k & ~SLOT) != NONE_B) {
        knn_exact_one(q, g, cbx, cby, bin_start, sr_xy, sr_inst, r2, knn, out, cnt + inst, r_used);
        return;
    }
    // Survivors in different buckets are already in exact order; only when two live slots share a bucket (rare off a lattice)
    // are the exact d2 recomputed and the list put in (d2, instance) order by an odd-even transposition pass.
    i32 bj[KCAP], bs[KCAP];
    int found = 0;
    bool shared_bucket = false;
#pragma unroll
    for (int u = 0; u < KCAP; ++u) {
        const bool ok = (u >= head) & ((bk[u] & ~SLOT) != NONE_B);
        bs[u] = ok ? spos[(bk[u] & SLOT) * 128 + threadIdx.x] : 0;
        bj[u] = ok ? sr_inst[bs[u]] : ((u < head) ? (i32)0x80000000 : 0x7fffffff);
        found += ok;
        if (u > 0) shared_bucket |= ok & ((bk[u] & ~SLOT) == (bk[u - 1] & ~SLOT)) & (u - 1 >= head);
    }
    if (shared_bucket) {
        double bd[KCAP];
#pragma unroll
        for (int u = 0; u < KCAP; ++u) {
            const bool ok = (u >= head) & ((bk[u] & ~SLOT) != NONE_B);
            const double2 p = sr_xy[bs[u]];
            const double ddx = __dsub_rn(p.x, q.x), ddy = __dsub_rn(p.y, q.y);
            bd[u] = ok ? __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)) : ((u < head) ? -INFINITY : INFINITY);
        }
        bool swapped = true;
        while (swapped) {
            swapped = false;
#pragma unroll
            for (int par = 0; par < 2; ++par)
#pragma unroll
                for (int u = par; u + 1 < KCAP; u += 2) {
                    const bool sw = cand_less(bd[u + 1], bj[u + 1], bd[u], bj[u]);
                    const double d0 = bd[u], d1 = bd[u + 1];
                    const i32 j0 = bj[u], j1 = bj[u + 1];
                    bd[u] = sw ? d1 : d0; bd[u + 1] = sw ? d0 : d1;
                    bj[u] = sw ? j1 : j0; bj[u + 1] = sw ? j0 : j1;
                    swapped |= sw;
                }
        }
    }
    cnt[inst] = found;
#pragma unroll
    for (int u = 0; u < KCAP; ++u)
        if (u >= head) {
            const bool ok = (u - head) < found;
            out[u - head] = ok ? bj[u] : -1;
            if (ok) r_used[bj[u]] = 1;
        }
#undef last_k
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
__global__ void k_fill_i32(i32 *p, i64 n, i32 v) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// off3[0..W] / [W+1..2W+1] / [2W+2..3W+2]: kept-aligned, kept-ref and pair offsets of each window
// (also written to the three device-side offset arrays the later kernels read, instead of three device-to-device copies)
__global__ void k_window_offsets(const i32 *__restrict__ newA, const int2 *__restrict__ rmap, const i32 *__restrict__ poff,
                                 const i32 *__restrict__ a_off, const i32 *__restrict__ r_off, int W, i32 *__restrict__ off3,
                                 i32 *__restrict__ ka_off, i32 *__restrict__ kr_off, i32 *__restrict__ p_off) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > W) return;
    const i32 ka = newA[a_off[w]], kr = rmap[r_off[w]].x, po = poff[a_off[w]];
    off3[w] = ka; off3[W + 1 + w] = kr; off3[2 * (W + 1) + w] = po;
    ka_off[w] = ka; kr_off[w] = kr; p_off[w] = po;
}

// Frame compaction in one launch (src/utils.py:734-741): blocks [0, tilesA) scan the aligned instances — kept flag
// (cnt > 0) and emitted pairs (eff) together — and write the kept aligned rows; blocks [tilesA, ..) scan the used flag of
// the reference instances and write the kept reference rows.  Item nAi / nRi is a sentinel that receives the totals.
// Only the row lists are written here: the coordinate / type / size columns of the kept rows, which the triangle, grouping and
// separation stages gather through, are materialised when the first of those stages runs (batch_kept_columns) — a caller
// that only wants candidates and costs does not pay for them.
constexpr int COMPACT_THREADS = 1024, COMPACT_ITEMS = 4;   // big tiles: one L2 round trip of look-back per 32 tiles
__global__ void __launch_bounds__(COMPACT_THREADS, 2) k_compact_frames(
    const i32 *__restrict__ cnt, const i32 *__restrict__ eff, const i32 *__restrict__ r_used, i64 nAi, i64 nRi, unsigned tilesA, ScanCtx scA, ScanCtx scR,
    const i32 *__restrict__ a_src, const i32 *__restrict__ r_src, i32 *__restrict__ newA, int2 *__restrict__ rmap, i32 *__restrict__ poff,
    i32 *__restrict__ keepA, i32 *__restrict__ row_ptr, i32 *__restrict__ keepR) {
    __shared__ int smem[2 * (COMPACT_THREADS / 32) + 2];
    int f[COMPACT_ITEMS], e[COMPACT_ITEMS], sum[2] = {0, 0}, excl[2], tot[2], pre[2];
    if (blockIdx.x < tilesA) {
        const i64 base = ((i64)blockIdx.x * COMPACT_THREADS + threadIdx.x) * COMPACT_ITEMS;
#pragma unroll
        for (int k = 0; k < COMPACT_ITEMS; ++k) {
            const bool in = base + k < nAi;
            f[k] = in ? cnt[base + k] > 0 : 0;
            e[k] = in ? eff[base + k] : 0;
            sum[0] += f[k];
            sum[1] += e[k];
        }
        device_exclusive_scan<2, COMPACT_THREADS>(scA, (int)blockIdx.x, sum, excl, tot, pre, smem);
        i32 kpos = excl[0], ppos = excl[1];
#pragma unroll
        for (int k = 0; k < COMPACT_ITEMS; ++k) {
            const i64 i = base + k;
            if (i <= nAi) {
                newA[i] = kpos;
                poff[i] = ppos;
                if (i == nAi) row_ptr[kpos] = ppos;   // row_ptr[nKA] = P
                if (f[k]) {
                    keepA[kpos] = a_src[i];
                    row_ptr[kpos] = ppos;
                }
            }
            kpos += f[k];
            ppos += e[k];
        }
    } else {
        const unsigned tile = blockIdx.x - tilesA;
        const i64 base = ((i64)tile * COMPACT_THREADS + threadIdx.x) * COMPACT_ITEMS;
#pragma unroll
        for (int k = 0; k < COMPACT_ITEMS; ++k) {
            f[k] = base + k < nRi ? r_used[base + k] != 0 : 0;
            sum[0] += f[k];
        }
        device_exclusive_scan<2, COMPACT_THREADS>(scR, (int)tile, sum, excl, tot, pre, smem);
        i32 kpos = excl[0];
#pragma unroll
        for (int k = 0; k < COMPACT_ITEMS; ++k) {
            const i64 i = base + k;
            if (i <= nRi) {
                const i32 row = i < nRi ? r_src[i] : 0;
                rmap[i] = make_int2(kpos, row);   // kept index (batch-wide) and section row of every reference instance
                if (f[k]) keepR[kpos] = row;
            }
            kpos += f[k];
        }
    }
}

// Packed row records for the pair-cost kernel: [x, y, p_0 .. p_{K-1}, zero padding] per row, RS = 2 + K rounded up to even.
__global__ void k_pack_records(const double2 *__restrict__ xy, const double *__restrict__ prob, i64 n, int K, int RS, double *__restrict__ rec) {
    const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * RS) return;
    const i64 row = idx / RS;
    const int e = (int)(idx - row * RS);
    double v = 0.0;
    if (e == 0) v = xy[row].x;
    else if (e == 1) v = xy[row].y;
    else if (e - 2 < K) v = prob[row * K + (e - 2)];
    rec[idx] = v;
}
static void section_records(Section *sec, cudaStream_t s) {
    if (sec->have_rec) return;
    const int K = sec->K, RS = (2 + K + 1) & ~1;
    sec->rec_stride = RS;
    sec->a_rec.alloc(sec->nA * RS, s); sec->r_rec.alloc(sec->nR * RS, s);
    if (sec->nA > 0) LAUNCH(k_pack_records, blocks_for(sec->nA * RS, 256), 256, 0, s, sec->a_xy.p, sec->a_prob.p, sec->nA, K, RS, sec->a_rec.p);
    if (sec->nR > 0) LAUNCH(k_pack_records, blocks_for(sec->nR * RS, 256), 256, 0, s, sec->r_xy.p, sec->r_prob.p, sec->nR, K, RS, sec->r_rec.p);
    sec->have_rec = true;
}

// one thread per (aligned instance, slot): pair indices + cost (src/same.py:1183-1188).  Per pair the reference side costs one
// 8-byte gather (rmap: kept index + section row of the candidate instance, written by k_compact_frames) and one run of
// adjacent sectors (the row's packed record); the window of a warp's first row is found once and the other lanes step
// forward from it.  The L1 sum runs over the zero padding too: s + |0 - 0| == s exactly.
template <int NP>   // NP = double2 words per record (1 + ceil(K / 2)); 0 = any
__global__ void __launch_bounds__(256, 8) k_emit_pairs(const i32 *__restrict__ cand, const i32 *__restrict__ eff, int knn, int knn_shift, i64 nAi,
                                                    const i32 *__restrict__ newA, const int2 *__restrict__ rmap, const i32 *__restrict__ poff,
                                                    const i32 *__restrict__ a_off, int W, const i32 *__restrict__ ka_off,
                                                    const i32 *__restrict__ kr_off, const i32 *__restrict__ a_src, const double *__restrict__ a_rec,
                                                    const double *__restrict__ r_rec, int RS, double ct_coeff, double dist_coeff,
                                                    int2 *__restrict__ pairs, double *__restrict__ cost) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;     // nAi * knn < 2^31 (batch_candidates)
    const unsigned i = knn_shift >= 0 ? tid >> knn_shift : tid / (unsigned)knn;
    const int slot = (int)(tid - i * (unsigned)knn);
    const unsigned i0 = min(__shfl_sync(0xffffffffu, i, 0), (unsigned)(nAi > 0 ? nAi - 1 : 0));
    int w = 0;
    if ((threadIdx.x & 31) == 0) w = find_window(a_off, W, (i32)i0);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (i >= nAi || slot >= eff[i]) return;
    while (w + 1 < W && a_off[w + 1] <= (i32)i) ++w;
    const i32 jinst = cand[tid];
    const int2 m = rmap[jinst];
    const i32 p = poff[i] + slot;
    const double2 *__restrict__ ra = (const double2 *)(a_rec + (size_t)a_src[i] * RS), *__restrict__ rr = (const double2 *)(r_rec + (size_t)m.y * RS);
    const double2 A = ra[0], R = rr[0];
    double s = 0.0;
    if (NP > 0) {
#pragma unroll
        for (int c = 1; c < NP; ++c) {   // left-to-right
            const double2 u = ra[c], v = rr[c];
            s = __dadd_rn(s, fabs(__dsub_rn(u.x, v.x)));
            s = __dadd_rn(s, fabs(__dsub_rn(u.y, v.y)));
        }
    } else {
#pragma unroll 1
        for (int c = 1; 2 * c < RS; ++c) {
            const double2 u = ra[c], v = rr[c];
            s = __dadd_rn(s, fabs(__dsub_rn(u.x, v.x)));
            s = __dadd_rn(s, fabs(__dsub_rn(u.y, v.y)));
        }
    }
    pairs[p] = make_int2(newA[i] - ka_off[w], m.x - kr_off[w]);
    const double dc = __dadd_rn(fabs(__dsub_rn(A.x, R.x)), fabs(__dsub_rn(A.y, R.y)));
    cost[p] = __dadd_rn(__dmul_rn(ct_coeff, s), __dmul_rn(dist_coeff, dc));
}

// SAME_ARR_PAIR_J: the reference index of every pair as its own array (half the bytes of PAIRS over PCIe)
__global__ void k_pair_j(const int2 *__restrict__ pairs, i64 P, i32 *__restrict__ pj) {
    const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) pj[p] = pairs[p].y;
}
// coordinate / type / size columns of the kept rows of both frames (gathered once, on first use by a later stage).  The kept
// counts are read from device memory (last entry of the kept offsets), so the launch needs no host round trip.
__global__ void k_kept_columns(const i32 *__restrict__ keepA, const i32 *__restrict__ keepR, const i32 *__restrict__ ka_off,
                               const i32 *__restrict__ kr_off, int W, i64 nAi, const double2 *__restrict__ a_xy, const double2 *__restrict__ r_xy,
                               const i32 *__restrict__ a_type, const double *__restrict__ a_size, const double *__restrict__ r_size,
                               double2 *__restrict__ ka_xy, i32 *__restrict__ ka_type, double *__restrict__ ka_size, double2 *__restrict__ kr_xy,
                               double *__restrict__ kr_size) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nAi) {
        if (i >= ka_off[W]) return;
        const i32 row = keepA[i];
        ka_xy[i] = a_xy[row]; ka_type[i] = a_type[row]; ka_size[i] = a_size[row];
    } else {
        const i64 k = i - nAi;
        if (k >= kr_off[W]) return;
        const i32 row = keepR[k];
        kr_xy[k] = r_xy[row]; kr_size[k] = r_size[row];
    }
}
void batch_kept_columns(Batch *b) {
    if (b->have_kept_cols) return;
    cudaStream_t s = b->stream;
    Section *sec = b->sec;
    const i64 nAi = b->nAi, nRi = b->nRi;   // upper bounds of the kept counts
    b->ka_xy.alloc(nAi, s); b->ka_type.alloc(nAi, s); b->ka_size.alloc(nAi, s); b->kr_xy.alloc(nRi, s); b->kr_size.alloc(nRi, s);
    if (nAi + nRi > 0)
        LAUNCH(k_kept_columns, blocks_for(nAi + nRi, 256), 256, 0, s, b->keepA.p, b->keepR.p, b->d_ka_off.p, b->d_kr_off.p, (int)b->W, nAi, sec->a_xy.p,
               sec->r_xy.p, sec->a_type.p, sec->a_size.p, sec->r_size.p, b->ka_xy.p, b->ka_type.p, b->ka_size.p, b->kr_xy.p, b->kr_size.p);
    b->have_kept_cols = true;
}

__global__ void k_pair_j16(const int2 *__restrict__ pairs, i64 P, unsigned short *__restrict__ pj) {
    const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) pj[p] = (unsigned short)pairs[p].y;
}
void batch_pair_j16(Batch *b) {
    if (b->have_pair_j16) return;
    for (i64 w = 0; w < b->W; ++w)   // (the caller settled the batch: the kept offsets are on the host)
        REQUIRE(b->kr_off[w + 1] - b->kr_off[w] <= 65536, SAME_E_LIMIT, "a window keeps more than 65,536 reference rows: PAIR_J16 cannot hold its indices");
    b->pair_j16.alloc(b->P, b->stream);
    if (b->P > 0) LAUNCH(k_pair_j16, blocks_for(b->P, 256), 256, 0, b->stream, b->pairs.p, b->P, b->pair_j16.p);
    b->have_pair_j16 = true;
}
void batch_pair_j(Batch *b) {
    if (b->have_pair_j) return;
    b->pair_j.alloc(b->P, b->stream);
    if (b->P > 0) LAUNCH(k_pair_j, blocks_for(b->P, 256), 256, 0, b->stream, b->pairs.p, b->P, b->pair_j.p);
    b->have_pair_j = true;
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
        double bw = std::sqrt(target_per_bin(knn) * std::max(ex, 1e-300) * std::max(ey, 1e-300) / (double)std::max<i64>(nref, 1));
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
    small_h2d(b->arena, d_grids.p, grids.data(), sizeof(GridParams) * W, s);

    // counting sort of both frames' instances by bin
    const i64 nb1 = nbins + 1, nI = nAi + nRi;
    DevBuf<i32> bkey, bcnt, bstart, sorted_inst;
    DevBuf<double2> sorted_xy;
    bkey.alloc(nI, s); bcnt.alloc(2 * nb1, s); bstart.alloc(2 * nb1, s); sorted_inst.alloc(nI, s); sorted_xy.alloc(nI, s);
    bcnt.zero(s);
    if (nI > 0)
        LAUNCH(k_bin_count, blocks_for(nI, 256), 256, 0, s, b->a_src.p, b->r_src.p, sec->a_xy.p, sec->r_xy.p, nAi, nRi, b->d_a_off.p, b->d_r_off.p, (int)W,
               d_grids.p, nb1, bkey.p, bcnt.p);
    scan_i32(sec, bcnt.p, bstart.p, 2 * nb1, s);
    if (nI > 0)
        LAUNCH(k_bin_scatter, blocks_for(nI, 256), 256, 0, s, b->a_src.p, b->r_src.p, sec->a_xy.p, sec->r_xy.p, nAi, nRi, nb1, bkey.p, bstart.p, bcnt.p,
               sorted_xy.p, sorted_inst.p);

    // top-k search: aligned instances are sorted[0, nAi), reference instances sorted[nAi, nAi + nRi); the bin starts of
    // the reference frame (second half of bstart) are positions in the whole sorted array
    b->cand.alloc(nAi * knn, s);
    b->cnt.alloc(nAi + 1, s);
    b->r_used.alloc(nRi + 1, s);
    b->r_used.zero(s);
    const double r2 = radius * radius;
    if (nAi > 0) {
        const unsigned grid = blocks_for(nAi, 128);
#define KNN_ARGS sorted_xy.p, sorted_inst.p, nAi, b->d_a_off.p, (int)W, d_grids.p, bstart.p + nb1, sorted_xy.p, sorted_inst.p, r2, knn
        unsigned long long *evals = nullptr;
        if (g_prof) {   // profiling pass only: count the distance evaluations
            b->knn_evals.alloc(1, s);
            b->knn_evals.zero(s);
            evals = b->knn_evals.p;
        }
        if (knn <= 4) LAUNCH(k_knn<4>, grid, 128, 0, s, KNN_ARGS, b->cand.p, b->cnt.p, b->r_used.p, evals);
        else if (knn <= 8) LAUNCH(k_knn<8>, grid, 128, 0, s, KNN_ARGS, b->cand.p, b->cnt.p, b->r_used.p, evals);
        else if (knn <= 16) LAUNCH(k_knn<16>, grid, 128, 0, s, KNN_ARGS, b->cand.p, b->cnt.p, b->r_used.p, evals);
        else if (knn <= 32) LAUNCH(k_knn<32>, grid, 128, 0, s, KNN_ARGS, b->cand.p, b->cnt.p, b->r_used.p, evals);
        else {
            DevBuf<double> gd;
            DevBuf<i32> gj;
            gd.alloc(nAi * knn, s); gj.alloc(nAi * knn, s);
            LAUNCH(k_knn_big, grid, 128, 0, s, KNN_ARGS, gd.p, gj.p, b->cand.p, b->cnt.p, b->r_used.p);
        }
#undef KNN_ARGS
    }

    // first use of the columns that section_build uploads on the auxiliary stream (type codes, sizes, probabilities)
    if (sec->aux_ready) CK(cudaStreamWaitEvent(s, sec->aux_ready, 0));

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

    // compaction + emission without a host round trip: outputs are sized by their upper bounds (kept rows <= instances,
    // pairs <= knn per aligned instance); the per-window offsets come back once, at the end
    DevBuf<i32> poff, off3;
    DevBuf<int2> rmap;
    b->newA.alloc(nAi + 1, s); rmap.alloc(nRi + 1, s); poff.alloc(nAi + 1, s); off3.alloc(3 * (W + 1), s);
    b->keepA.alloc(nAi, s);
    b->row_ptr.alloc(nAi + 1, s);
    b->keepR.alloc(nRi, s);
    b->have_kept_cols = false;
    b->pairs.alloc(nAi * knn, s); b->cost.alloc(nAi * knn, s);
    const unsigned tilesA = blocks_for(nAi + 1, COMPACT_THREADS * COMPACT_ITEMS), tilesR = blocks_for(nRi + 1, COMPACT_THREADS * COMPACT_ITEMS);
    {
        scan_reserve(sec, 2 * ((i64)tilesA + tilesR), s);   // two independent scans in one launch: disjoint tile-state words
        const ScanCtx scA = scan_ctx_at(sec, 0, tilesA), scR = scan_ctx_at(sec, 2 * (i64)tilesA, tilesR);
        LAUNCH(k_compact_frames, tilesA + tilesR, COMPACT_THREADS, 0, s, b->cnt.p, eff, b->r_used.p, nAi, nRi, tilesA, scA, scR, b->a_src.p, b->r_src.p,
               b->newA.p, rmap.p, poff.p, b->keepA.p, b->row_ptr.p, b->keepR.p);
    }
    b->d_ka_off.alloc(W + 1, s); b->d_kr_off.alloc(W + 1, s); b->d_p_off.alloc(W + 1, s);
    LAUNCH(k_window_offsets, blocks_for(W + 1, 128), 128, 0, s, b->newA.p, rmap.p, poff.p, b->d_a_off.p, b->d_r_off.p, (int)W, off3.p, b->d_ka_off.p,
           b->d_kr_off.p, b->d_p_off.p);
    // (measured and not kept: one thread per row with 8 gather chains in flight, 224 us vs 129 us; rows visited in bin order so
    // that neighbouring warps gather the same reference rows, 128 us; two slots per thread, 130 us — the kernel moves 292 MB of
    // scattered 32-byte sectors through DRAM at 2.3 TB/s either way; profiles/r1m)
    if (nAi > 0) {
        section_records(sec, s);
        int knn_shift = -1;
        for (int sh = 0; sh < 7; ++sh) if ((1 << sh) == knn) knn_shift = sh;
#define EMIT(NP) LAUNCH(k_emit_pairs<NP>, blocks_for(nAi * knn, 256), 256, 0, s, b->cand.p, eff, knn, knn_shift, nAi, b->newA.p, rmap.p, poff.p,      \
                        b->d_a_off.p, (int)W, b->d_ka_off.p, b->d_kr_off.p, b->a_src.p, sec->a_rec.p, sec->r_rec.p, sec->rec_stride, dist_ct_coeff,      \
                        dist_ct_coeff * 0.001, b->pairs.p, b->cost.p)
        switch (sec->rec_stride / 2) {
        case 1: EMIT(1); break;
        case 2: EMIT(2); break;
        case 3: EMIT(3); break;
        case 4: EMIT(4); break;
        default: EMIT(0); break;
        }
#undef EMIT
    }
    // the window offsets come back asynchronously; whoever needs them on the host first waits for them (batch_settle)
    small_d2h(b->pin_cand(), off3.p, sizeof(i32) * 3 * (W + 1), s);
    b->pend_cand = true;
    b->pend_renum = false;
    b->stage = 1;
    b->have_groups = false;
    b->have_start = false;
    b->have_pair_j = false;
    b->have_pair_j16 = false;
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
                              int max_matches, int multiplier, i32 *__restrict__ ref_gid, i32 *__restrict__ g_node, i32 *__restrict__ g_cnt /* zeroed */,
                              i32 *__restrict__ g_limit) {
    i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
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
// ascending pair index inside every group.  Groups hold ~knn entries: up to 16 they are sorted in registers by a bitonic
// network of integer min/max (static indices, no divergence, no local memory); longer ones by insertion sort in place.
template <int N>
__device__ __forceinline__ void bitonic_sort_regs(i32 (&v)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const i32 lo = min(v[i], v[l]), hi = max(v[i], v[l]);
                    v[i] = up ? lo : hi;
                    v[l] = up ? hi : lo;
                }
            }
}
template <int N>
__device__ __forceinline__ void sort_group_regs(i32 *__restrict__ a, i32 n) {
    i32 v[N];
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = k < n ? a[k] : 0x7fffffff;
    bitonic_sort_regs<N>(v);
#pragma unroll
    for (int k = 0; k < N; ++k)
        if (k < n) a[k] = v[k];
}
__global__ void k_group_sort(const i32 *__restrict__ g_ptr, const i32 *__restrict__ n_groups, i32 *__restrict__ g_idx) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= *n_groups) return;
    const i32 lo = g_ptr[g], n = g_ptr[g + 1] - lo;
    i32 *a = g_idx + lo;
    if (n <= 1) return;
    if (n <= 8) sort_group_regs<8>(a, n);
    else if (n <= 16) sort_group_regs<16>(a, n);
    else {
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
    REQUIRE(b->stage >= 1, SAME_E_STATE, "same_batch_groups before same_batch_candidates");
    batch_settle(b);
    batch_kept_columns(b);
    const i64 W = b->W, P = b->P, nKR = b->nKR;
    b->g_off.assign(W + 1, 0);
    b->G = 0;
    b->have_groups = true;
    b->pend_groups = false;
    if (P == 0) {  // keep REF_GROUP_PTR (G + 1 = 1 element) readable
        b->g_ptr.alloc(1, s);
        b->g_ptr.zero(s);
        return;
    }
    // every buffer is sized by its bound (groups <= kept reference rows), so nothing has to come back before the end
    DevBuf<i32> first, cnt, head, gid_at, goff, ref_gid, g_cnt, cursor;
    DevBuf<unsigned long long> wmax;
    first.alloc(nKR, s); cnt.alloc(nKR, s); head.alloc(P + 1, s); gid_at.alloc(P + 1, s); goff.alloc(W + 1, s); ref_gid.alloc(nKR, s);
    wmax.alloc(W, s);
    wmax.zero(s);
    cnt.zero(s);
    b->g_node.alloc(nKR, s); b->g_ptr.alloc(nKR + 1, s); b->g_idx.alloc(P, s); b->g_limit.alloc(nKR, s);
    g_cnt.alloc(nKR + 1, s); cursor.alloc(nKR, s);
    g_cnt.zero(s);
    cursor.zero(s);
    LAUNCH(k_fill_i32, blocks_for(nKR, 256), 256, 0, s, first.p, nKR, 0x7fffffff);
    LAUNCH(k_group_count, blocks_for(P, 256), 256, 0, s, b->pairs.p, P, b->d_p_off.p, b->d_kr_off.p, (int)W, first.p, cnt.p);
    LAUNCH(k_group_heads, blocks_for(P + 1, 256), 256, 0, s, b->pairs.p, P, b->d_p_off.p, b->d_kr_off.p, (int)W, first.p, head.p);
    scan_i32(b->sec, head.p, gid_at.p, P + 1, s);
    LAUNCH(k_pick, blocks_for(W + 1, 128), 128, 0, s, gid_at.p, b->d_p_off.p, (int)(W + 1), goff.p);
    small_d2h(b->pin_groups(), goff.p, sizeof(i32) * (W + 1), s);
    b->pend_groups = true;
    LAUNCH(k_window_max_size, blocks_for(nKR, 256), 256, 0, s, b->kr_size.p, nKR, b->d_kr_off.p, (int)W, wmax.p);
    LAUNCH(k_group_setup, blocks_for(nKR, 256), 256, 0, s, nKR, b->d_kr_off.p, (int)W, first.p, cnt.p, gid_at.p, b->kr_size.p, wmax.p, max_matches,
           multiplier, ref_gid.p, b->g_node.p, g_cnt.p, b->g_limit.p);
    scan_i32(b->sec, g_cnt.p, b->g_ptr.p, nKR + 1, s);   // entries past the number of groups repeat the total
    LAUNCH(k_group_fill, blocks_for(P, 256), 256, 0, s, b->pairs.p, P, b->d_p_off.p, b->d_kr_off.p, (int)W, ref_gid.p, b->g_ptr.p, cursor.p, b->g_idx.p);
    LAUNCH(k_group_sort, blocks_for(nKR, 128), 128, 0, s, b->g_ptr.p, gid_at.p + P, b->g_idx.p);
}

}  // namespace same
