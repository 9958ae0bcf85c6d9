// common.cuh — shared host/device plumbing of libsame_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/same_b200.h"
#include "scan.cuh"

typedef int64_t i64;
typedef int32_t i32;

namespace same {

extern thread_local std::string g_err;
extern std::atomic<long long> g_launches;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define CK(call)                                                                                            \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            throw same::Error(SAME_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" +    \
                                               __FILE__ + ":" + std::to_string(__LINE__) + ")");          \
    } while (0)

#define REQUIRE(cond, code, msg)                                                                            \
    do {                                                                                                    \
        if (!(cond)) throw same::Error(code, std::string(msg));                                             \
    } while (0)

// Optional per-kernel timing (bench.py's roofline section): when enabled, every launch is bracketed by
// CUDA events on the launching stream; same_profile_report() sums elapsed time per kernel name.
struct ProfRec { const char *name; cudaEvent_t a, b; };
extern bool g_prof;
extern bool g_debug;  // SAME_B200_DEBUG=1: synchronise + check after every launch
extern std::vector<ProfRec> g_prof_recs;
extern std::mutex g_prof_mu;   // sections may be driven from several host threads (CandidateStream)
struct ProfScope {
    ProfRec r{nullptr, nullptr, nullptr};
    cudaStream_t s;
    ProfScope(const char *name, cudaStream_t stream) : s(stream) {
        if (!g_prof) return;
        r.name = name;
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, s);
    }
    ~ProfScope() {
        if (!r.name) return;
        cudaEventRecord(r.b, s);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof_recs.push_back(r);
    }
};

// every kernel launch goes through this macro so bench.py can report gpu_launches
#define LAUNCH(kernel, grid, block, smem, stream, ...)                                                      \
    do {                                                                                                    \
        if (same::g_debug) {                                                                                \
            cudaError_t pre__ = cudaGetLastError();                                                         \
            if (pre__ != cudaSuccess)                                                                       \
                throw same::Error(SAME_E_CUDA, std::string("stale CUDA error before " #kernel ": ") + cudaGetErrorString(pre__) + \
                                                   " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
        }                                                                                                   \
        same::ProfScope prof__(#kernel, (stream));                                                          \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                         \
        same::g_launches.fetch_add(1, std::memory_order_relaxed);                                           \
        CK(cudaGetLastError());                                                                             \
        if (same::g_debug) CK(same::stream_wait(stream));                                                   \
    } while (0)

// Host waits for a stream.  Default: cudaStreamSynchronize (the calling thread spins: lowest latency, what a solver callback
// wants).  Yielding mode (same_set_host_wait(1) / SAME_B200_HOST_WAIT=yield): the thread sleeps on a blocking-sync event
// instead — for hosts where several ranks with several section threads each share few cores and spinning waiters steal the
// cores that the other ranks' launches need.
extern std::atomic<int> g_host_wait_yield;
cudaError_t stream_wait(cudaStream_t s);   // every host wait in this library goes through it

inline unsigned blocks_for(i64 n, int per_block) { return (unsigned)std::max<i64>(1, (n + per_block - 1) / per_block); }

// Guard mode (SAME_B200_GUARD=1 / same_debug_guard): the library's own memory checker for boxes where compute-sanitizer is not
// available.  Every device buffer is allocated with a GUARD_BYTES zone of 0xA5 in front of and behind it and its body is filled
// with 0xCD on EVERY alloc() call (contents are never preserved across alloc), so a kernel that reads memory it did not write sees
// poison — NaN-like doubles, huge negative ints — and changes a parity result, and a kernel that writes outside its buffer
// breaks a zone: the zones are verified on the device when the buffer is released (guard_check counts corrupted zones,
// same_debug_guard reports them).
constexpr size_t GUARD_BYTES = 256;
extern bool g_guard;
void guard_fill(void *raw, size_t body_bytes, cudaStream_t s);     // zones + poison
void guard_check(void *raw, size_t body_bytes, cudaStream_t s);    // zones still intact?
void guard_report(int enable, i64 *corrupted, i64 *checked);

// stream-ordered device buffer
template <typename T>
struct DevBuf {
    T *p = nullptr;
    i64 n = 0;
    cudaStream_t s = nullptr;
    bool guarded = false;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) {
            if (guarded) {
                void *raw = (char *)p - GUARD_BYTES;
                guard_check(raw, sizeof(T) * (size_t)n, s);
                cudaFreeAsync(raw, s);
            } else {
                cudaFreeAsync(p, s);
            }
        }
        p = nullptr;
        n = 0;
        guarded = false;
    }
    // contents are NOT preserved
    void alloc(i64 count, cudaStream_t stream) {
        if (count <= n && p) {
            s = stream;
            if (guarded) guard_fill((char *)p - GUARD_BYTES, sizeof(T) * (size_t)n, stream);
            return;
        }
        release();
        s = stream;
        n = std::max<i64>(count, 1);
        if (g_guard) {
            void *raw = nullptr;
            CK(cudaMallocAsync(&raw, sizeof(T) * (size_t)n + 2 * GUARD_BYTES, stream));
            p = (T *)((char *)raw + GUARD_BYTES);
            guarded = true;
            guard_fill(raw, sizeof(T) * (size_t)n, stream);
        } else {
            CK(cudaMallocAsync((void **)&p, sizeof(T) * (size_t)n, stream));
        }
    }
    void zero(cudaStream_t stream) { CK(cudaMemsetAsync(p, 0, sizeof(T) * (size_t)n, stream)); }
    void swap(DevBuf &o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(s, o.s); std::swap(guarded, o.guarded); }
};

struct Scratch {
    DevBuf<unsigned char> buf;
    void *get(size_t bytes, cudaStream_t s) {
        buf.alloc((i64)bytes + 256, s);
        return buf.p;
    }
};

// exclusive prefix sum of n i32 -> out[n] (+ out[n] = total when with_total)
void exclusive_scan_i32(const i32 *in, i32 *out, i64 n, Scratch &sc, cudaStream_t s);

// Small host <-> device transfers (window offsets, counts, rectangle lists, grid parameters) do not go through the copy
// engines: a large transfer of ANOTHER section queued on the same engine (CandidateStream overlaps one section's download with
// the next one's upload and kernels) would hold a 32-byte read-back for milliseconds.  Instead a tiny kernel moves the words
// between device memory and page-locked, device-mapped host memory, ordered on the stream like any other launch.
struct PinArena {   // bump allocator over page-locked blocks of a process-wide pool (section.cu); blocks return to the pool on release
    std::vector<std::pair<char *, size_t>> blocks;
    size_t used = 0;
    void *get(size_t bytes);
    void release();
    ~PinArena() { release(); }
};
constexpr size_t SMALL_COPY_LIMIT = 256 << 10;   // larger transfers use cudaMemcpyAsync
void small_d2h(void *host_pinned, const void *dev, size_t bytes, cudaStream_t s);             // host_pinned: page-locked (cudaHostAlloc)
void small_h2d(PinArena &arena, void *dev, const void *host, size_t bytes, cudaStream_t s);   // host: any memory, staged through the arena
struct SmallCopy { void *dev; const void *host; size_t bytes; };
void small_h2d_many(PinArena &arena, const SmallCopy *items, int n, cudaStream_t s);            // up to 4 uploads, ONE launch

struct Section {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // Columns the search does not read (probabilities, type codes, sizes) are uploaded on a second stream while the window
    // subsetting, binning and search already run on the coordinates; `aux_ready` is waited for before their first use.
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t aux_ready = nullptr;
    i64 nA = 0, nR = 0;
    int K = 0;
    DevBuf<double2> a_xy, r_xy;
    DevBuf<double> a_prob, r_prob, a_size, r_size;
    DevBuf<i32> a_type, r_type;
    // one packed record per row, [x, y, p_0 .. p_{K-1}, 0-padding to an even count] (rec_stride doubles): the pair-cost kernel
    // gathers a row's coordinates and probabilities from ONE run of adjacent sectors instead of two arrays (built on the
    // device the first time a batch of this section computes costs, section_records)
    DevBuf<double> a_rec, r_rec;
    int rec_stride = 0;
    bool have_rec = false;
    double bbox[4] = {0, 0, 0, 0};  // x_min, x_max, y_min, y_max over both frames
    // precomputed triangulation (vertex ids resolved to section rows; -1 = id not present)
    i64 Tg = -1;
    DevBuf<i32> tri_rows;  // [Tg*3]
    Scratch scratch;
    PinArena arena;
    // tile states of the in-kernel prefix sums (scan.cuh): zeroed when (re)allocated, separated by epoch afterwards
    DevBuf<unsigned long long> scan_state;
    unsigned scan_epoch = 0;
};

// tile-state context for one launch of `tiles` blocks scanning `n_streams` sums; a launch that runs several independent
// scans reserves the words of all of them first and places each with scan_ctx_at
ScanCtx scan_ctx(Section *sec, i64 tiles, int n_streams, cudaStream_t s);
void scan_reserve(Section *sec, i64 words, cudaStream_t s);
ScanCtx scan_ctx_at(Section *sec, i64 word_offset, i64 tiles);
void scan_i32(Section *sec, const i32 *in, i32 *out, i64 n, cudaStream_t s);   // out[i] = sum(in[0..i-1]), one launch

struct GridParams {  // uniform bin grid of one window (reference side)
    double x0, y0, inv_w, w;
    i32 nbx, nby, base, rings;
};

// Coarse uniform grid over the section bbox listing, per grid cell, the window rectangles that overlap it: a point /
// triangle only tests the handful of rectangles of its own cell instead of all W (built on the host, batch_subset).
struct RectIndexDev {
    double x0, y0, inv_cs;
    i32 nx, ny, max_len;
    const i32 *cell_ptr;    // [nx*ny + 1]
    const i32 *cell_rects;  // rectangle ids, ascending inside a cell
};
__device__ __forceinline__ int rect_cell(const RectIndexDev &ri, double x, double y) {
    int cx = (int)floor((x - ri.x0) * ri.inv_cs), cy = (int)floor((y - ri.y0) * ri.inv_cs);
    cx = min(max(cx, 0), ri.nx - 1);
    cy = min(max(cy, 0), ri.ny - 1);
    return cy * ri.nx + cx;
}

struct Batch {
    Section *sec = nullptr;
    cudaStream_t stream = nullptr;
    i64 W = 0;
    std::vector<double> rects;  // W*4
    DevBuf<double> d_rects;
    DevBuf<i32> ri_ptr, ri_rects;
    std::vector<i32> h_ri_ptr, h_ri_rects;
    RectIndexDev rindex{};
    Scratch scratch;
    PinArena arena;
    int stage = 0;  // 0 created, 1 candidates, 2 triangles in, 3 classified, 4 finalized

    // Small per-window results (offsets, counts) are copied to this page-locked block asynchronously and parsed into the
    // host vectors at the next synchronisation (batch_sync): a stage does not have to drain the stream just to return.
    i32 *pin = nullptr;
    i64 pin_n = 0;
    bool pend_cand = false, pend_tin = false, pend_renum = false, pend_groups = false;
    i32 *pin_cand() const { return pin; }                              // 3(W+1): ka_off, kr_off, p_off
    i32 *pin_tin() const { return pin + 3 * (W + 1); }                 // W+1
    i32 *pin_renum() const { return pin + 4 * (W + 1); }               // W+1: p_off after node removal
    i32 *pin_groups() const { return pin + 5 * (W + 1); }              // W+1
    i32 *pin_misc() const { return pin + 6 * (W + 1); }                // 4(W+1) + 8 scratch for stages that synchronise themselves

    // window instances (subset_data)
    i64 nAi = 0, nRi = 0;
    std::vector<i64> a_off, r_off;      // W+1
    DevBuf<i32> d_a_off, d_r_off;       // W+1
    DevBuf<i32> a_src, r_src;           // section row of each instance
    // row -> its window instances: entries [row_pos[row], row_pos[row+1]) of row_inst hold (window, instance index
    // within the frame) in ascending window order; aligned rows first, reference row r at index nA + r
    DevBuf<i32> row_pos;                // [nA + nR + 1]
    DevBuf<int2> row_inst;              // [nAi + nRi]

    // candidates
    int knn = 0;
    double radius = 0;
    DevBuf<i32> cand, cnt, eff;         // [nAi*knn] ref instance, [nAi], [nAi] pairs emitted (priority)
    DevBuf<i32> newA;                   // [nAi+1] batch-global kept index of each aligned instance (exclusive scan of cnt > 0)
    DevBuf<i32> r_used;                 // [nRi]
    DevBuf<unsigned long long> knn_evals;   // [1] distance evaluations of the search (filled only while profiling is enabled)

    // kept nodes (post-KNN frames), batch-global "kept index" = window offset + local index
    i64 nKA = 0, nKR = 0, P = 0;
    std::vector<i64> ka_off, kr_off, p_off;   // W+1 each
    DevBuf<i32> d_ka_off, d_kr_off, d_p_off;  // W+1
    DevBuf<i32> keepA, keepR;           // section rows
    DevBuf<double2> ka_xy, kr_xy;
    DevBuf<i32> ka_type;
    DevBuf<double> ka_size, kr_size;
    bool have_kept_cols = false;        // the five columns above are gathered on first use (batch_kept_columns)
    DevBuf<int2> pairs;                 // window-local (i, j)
    DevBuf<i32> pair_j;                 // pairs[].y on its own, built on first request (SAME_ARR_PAIR_J)
    bool have_pair_j = false;
    DevBuf<unsigned short> pair_j16;    // the same in 16 bits (SAME_ARR_PAIR_J16), when every window keeps <= 65,536 reference rows
    bool have_pair_j16 = false;
    DevBuf<double> cost;
    DevBuf<i32> row_ptr;                // [nKA+1] batch-global pair offset of each aligned row

    // groups (a4)
    i64 G = 0;
    std::vector<i64> g_off;
    DevBuf<i32> g_node, g_ptr, g_idx, g_limit;
    bool have_groups = false;

    // triangles in
    i64 Tin = 0;
    std::vector<i64> tin_off;
    DevBuf<i32> d_tin_off;
    DevBuf<int3> tin;                   // window-local vertices
    DevBuf<i32> tin_src;
    bool tin_has_src = false;
    DevBuf<unsigned char> cls;
    DevBuf<double> score;
    i64 n_band = 0;
    DevBuf<i32> band_idx;
    bool use_angle = false;

    // triangles out
    i64 T = 0;
    std::vector<i64> t_off;
    DevBuf<i32> d_t_off;
    DevBuf<int3> tri;
    DevBuf<i32> tri_src;
    DevBuf<double> t_weight;
    DevBuf<signed char> t_sign;
    DevBuf<double> t_bounds;
    DevBuf<i32> t_argv;
    // node -> triangle incidence (CSR, built on first request): slots [nt_ptr[i], nt_ptr[i] + nt_len[i]) of nt_idx, ascending
    DevBuf<i32> nt_ptr, nt_idx, nt_len;
    std::vector<i64> nt_off;
    bool have_incidence = false;
    i64 nUnc = 0;
    std::vector<i64> unc_off;
    DevBuf<i32> unc;

    // separation / postsolve
    DevBuf<double> x_dev;
    DevBuf<i32> match_j, match_p;       // [nKA]
    DevBuf<i32> sep_counts, cuts;
    i64 last_unc_sep = -1;
    DevBuf<i32> unc_list[2], unc_count;   // [0] source signs (k_tri_tables), [1] last separation call: triangles the orientation filter could not decide
    DevBuf<i32> t_mask;
    DevBuf<double> area_before, area_after;
    DevBuf<unsigned char> flipped;
    bool have_post = false;

    // greedy MIP start (init_helpers.py:110-132)
    DevBuf<unsigned char> start_x, start_unmatched;   // [P], [nKA]
    bool have_start = false;
};

// cudaStreamSynchronize + parse whatever small results were pending; batch_settle only synchronises if something is
void batch_sync(Batch *b);
void batch_settle(Batch *b);
void batch_pin_acquire(Batch *b);
void batch_pin_release(Batch *b);

// implemented across the .cu files
void section_build(Section *sec, const double *a_xy, const double *r_xy, const double *a_prob, const double *r_prob,
                   const i32 *a_type, const i32 *r_type, const double *a_size, const double *r_size);
void section_count_rects(Section *sec, i64 m, const double *rects, i64 *cntA, i64 *cntR);
void section_set_triangles(Section *sec, const i64 *a_vid, const i64 *tri_vid, i64 n_tri);
void batch_subset(Batch *b);
void batch_candidates(Batch *b, double radius, int knn, int priority, double dist_ct_coeff);
void batch_groups(Batch *b, int max_matches, int multiplier);
void batch_pair_j(Batch *b);
void batch_pair_j16(Batch *b);
void batch_kept_columns(Batch *b);   // ka_xy / ka_type / ka_size / kr_xy / kr_size, gathered on first use
void batch_incidence(Batch *b);
void batch_triangles_remap(Batch *b);
void batch_triangles_set(Batch *b, const i32 *tri, const i64 *tri_off);
void batch_tri_classify(Batch *b, double radius, int use_angle, double min_angle_deg, int ignore_same_type);
void batch_tri_override(Batch *b, i64 n, const i32 *idx, const unsigned char *cls);
void batch_tri_finalize(Batch *b, int ignore_same_type, int ensure_min, int remove_unconstrained);
void batch_separation(Batch *b, i64 w_lo, i64 w_hi, const double *x, i64 cap, i64 *n_viol, i64 *n_checked, i32 *cuts);
void batch_postsolve(Batch *b, i64 w_lo, i64 w_hi, const double *x);
void batch_uncertain(Batch *b, int which, i64 cap, i64 *n, i32 *tri_idx);
void postsolve_arrays(int device, i64 T, const i32 *tri, i64 nA, const double *a_xy, i64 nR, const double *r_xy, const i32 *match_j, i32 *mask,
                      double *area_before, double *area_after, unsigned char *flipped);

void batch_mip_start(Batch *b, double no_match_penalty, i32 *rounds_out);
void greedy_select_arrays(int device, i64 n, int degree, const i32 *nodes, const double *key, const unsigned char *eligible, i64 n_nodes,
                          unsigned char *selected, unsigned char *used_out, i32 *rounds_out);

void collapse_select_arrays(int device, i64 n, const double *xy, const i32 *type, const double *size, i64 T, const i32 *tri, double max_size,
                            unsigned char *selected, double *perim_out, i32 *rounds_out);

void segment_mean_arrays(int device, i64 n_rows, i64 C, const double *values, i64 G, const i64 *ptr, i64 n_members, const i32 *pos, double *out);

double measure_fp64_peak(int device);

// upload a small host vector of i64 offsets as i32 device array
void upload_offsets(const std::vector<i64> &h, DevBuf<i32> &d, cudaStream_t s);

// ---- device helpers -----------------------------------------------------------------
// window of element idx given W+1 ascending offsets (last offset = total)
__device__ __forceinline__ int find_window(const i32 *__restrict__ off, int W, i32 idx) {
    int lo = 0, hi = W;  // invariant: off[lo] <= idx < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= idx) lo = mid; else hi = mid;
    }
    return lo;
}

// (bx-ax)*(cy-ay) - (by-ay)*(cx-ax), one IEEE operation at a time (src/same.py:658,1146)
__device__ __forceinline__ double orient_naive(double ax, double ay, double bx, double by, double cx, double cy) {
    const double t1 = __dmul_rn(__dsub_rn(bx, ax), __dsub_rn(cy, ay));
    const double t2 = __dmul_rn(__dsub_rn(by, ay), __dsub_rn(cx, ax));
    return __dsub_rn(t1, t2);
}
__device__ __forceinline__ int sign_of(double v) { return (v > 0.0) - (v < 0.0); }

// Exact-predicate DIAGNOSTIC (never changes a result): can the sign of the naive expression above differ from the sign of the
// exact determinant?  Shewchuk's static filter for this very operation order: |computed - exact| <= (3 + 16 eps) eps (|t1| + |t2|),
// eps = 2^-53.  Triangles inside the bound are listed for the host, which decides them with rational arithmetic
// (same_b200/helpers.py::exact_orientation_sign) and counts the disagreements.
__device__ __forceinline__ bool orient_uncertain(double ax, double ay, double bx, double by, double cx, double cy) {
    // two coincident vertices (an incumbent that maps two cells to the same reference cell): the determinant is exactly zero and
    // the naive expression evaluates to exactly zero as well (both products vanish, or are the same product)
    if ((ax == bx && ay == by) || (ax == cx && ay == cy) || (bx == cx && by == cy)) return false;
    const double t1 = __dmul_rn(__dsub_rn(bx, ax), __dsub_rn(cy, ay));
    const double t2 = __dmul_rn(__dsub_rn(by, ay), __dsub_rn(cx, ax));
    const double det = __dsub_rn(t1, t2);
    const double bound = 3.3306690738754716e-16 * (fabs(t1) + fabs(t2));   // (3 + 16 eps) eps, rounded up
    return !(fabs(det) > bound);   // (NaN coordinates count as uncertain)
}
constexpr int UNC_CAP = 1 << 16;   // listed triangles per call; the count is exact beyond it

}  // namespace same
