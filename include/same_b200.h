/*
 * same_b200.h — C-ABI of libsame_b200.so: the B200 (sm_100a) implementation of the
 * data-parallel hot path of SAME (rohitsinghlab/SAME).
 *
 * The reference has no FFI: its boundary is a set of Python functions.  Every entry point
 * below names the reference function (file:line, relative to the reference repository)
 * whose array-level work it replaces; the Python package `same_b200` keeps the reference's
 * own signatures on top of this header (see INTEGRATION.md for the ctypes binding a
 * maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; all functions return 0 on success, a negative
 *     SAME_E_* code otherwise; same_last_error() gives the message (thread-local).
 *   - "host or device": pointers marked HD may be host or device memory (the library
 *     copies with cudaMemcpyDefault); everything else is host memory owned by the caller.
 *   - a *section* is one pair of frames (aligned/moving + reference) resident in HBM;
 *     a *batch* is a list of windows (rectangles, src/same.py:293-295) cut from a section
 *     and processed together by every kernel.  run_same() on a single pair of frames is
 *     a batch of one unbounded window.
 *   - indices are int32 on the ABI (N < 2^31), offsets int64.  Within a batch every
 *     per-window array is the concatenation over windows; same_batch_offsets() returns the
 *     W+1 boundaries.  Indices stored inside the arrays are WINDOW-LOCAL, exactly the
 *     numbers the reference would produce for that window.
 *   - one CUDA stream per section/batch; calls on one batch must come from one thread at a
 *     time (the reference's MIPSOL callback is serialised by Gurobi, src/same.py:1241).
 */
#ifndef SAME_B200_H
#define SAME_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAME_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define SAME_API __attribute__((visibility("default")))
#else
#define SAME_API
#endif

enum {
    SAME_OK = 0,
    SAME_E_CUDA = -1,      /* a CUDA call failed; see same_last_error() */
    SAME_E_ARG = -2,       /* invalid argument */
    SAME_E_STATE = -3,     /* stage called out of order */
    SAME_E_NOPAIRS = -4,   /* no valid pairs after KNN (reference raises ValueError, src/same.py:1002-1003) */
    SAME_E_LIMIT = -5      /* size limit of this build exceeded (e.g. knn > SAME_MAX_KNN) */
};

#define SAME_MAX_KNN 256

/* triangle classes written by same_batch_tri_classify (src/helpers.py:296-344) */
enum { SAME_TRI_DROP_RADIUS = 0, SAME_TRI_DROP_ANGLE = 1, SAME_TRI_SAME_TYPE = 2, SAME_TRI_KEEP = 3 };

/* arrays retrievable with same_batch_get(); element type in brackets */
enum {
    SAME_ARR_WIN_A = 1,        /* [i32]   section rows of the aligned cells of each window   (subset_data, src/same.py:293-295) */
    SAME_ARR_WIN_R = 2,        /* [i32]   section rows of the reference cells of each window */
    SAME_ARR_KEEP_A = 3,       /* [i32]   section rows kept after KNN compaction (+ unconstrained-node removal) (src/utils.py:734-737, src/same.py:1056-1085) */
    SAME_ARR_KEEP_R = 4,       /* [i32]   same for the reference frame */
    SAME_ARR_PAIRS = 5,        /* [i32x2] valid_pairs (src/utils.py:741, src/knn_utils.py:59-65) */
    SAME_ARR_COST = 6,         /* [f64]   objective coefficient of each pair (src/same.py:1182-1189) */
    SAME_ARR_REF_GROUP_NODE = 7,  /* [i32] ref index j of each ref_to_pairs group, first-appearance order (src/helpers.py:105-110) */
    SAME_ARR_REF_GROUP_PTR = 8,   /* [i32] G_total+1 BATCH-GLOBAL offsets into REF_GROUP_IDX (subtract the window's pair offset) */
    SAME_ARR_REF_GROUP_IDX = 9,   /* [i32] pair indices, ascending inside a group */
    SAME_ARR_REF_GROUP_LIMIT = 10,/* [i32] right-hand side of max_matches_<j> (src/helpers.py:130-138) */
    SAME_ARR_ROW_PTR = 11,     /* [i32]   nKeepA_total+1 BATCH-GLOBAL pair offsets: pairs of aligned row i are [ptr[i], ptr[i+1]) (aligned_to_pairs, src/helpers.py:107-110) */
    SAME_ARR_TRI_IN = 12,      /* [i32x3] triangles before filtering, window-local vertices (src/same.py:1018-1031) */
    SAME_ARR_TRI_IN_SRC = 13,  /* [i32]   index of each TRI_IN row in the global triangle list (remap path only) */
    SAME_ARR_TRI_CLASS = 14,   /* [u8]    SAME_TRI_* per TRI_IN row */
    SAME_ARR_TRI_BAND = 15,    /* [i32]   batch-global TRI_IN indices whose radius/angle decision lies within the guard band */
    SAME_ARR_TRI = 16,         /* [i32x3] triangles after filtering + add-back + node renumbering (src/helpers.py:233-395, src/same.py:1072-1077) */
    SAME_ARR_TRI_SRC = 17,     /* [i32]   window-local TRI_IN index of each TRI row */
    SAME_ARR_TRI_WEIGHT = 18,  /* [f64]   triangle_weights (src/same.py:1128-1135) */
    SAME_ARR_TRI_SIGN = 19,    /* [i8]    source_signs (src/same.py:1139-1146) */
    SAME_ARR_TRI_BOUNDS = 20,  /* [f64x4] min_x,max_x,min_y,max_y (src/helpers.py:184-210) */
    SAME_ARR_TRI_ARGV = 21,    /* [i32x4] max_x,min_x,max_y,min_y vertex (src/helpers.py:203-206) */
    SAME_ARR_UNCONSTRAINED = 22,/* [i32]  window-local (pre-removal) indices of truly unconstrained nodes (src/helpers.py:355-358) */
    SAME_ARR_MATCH_J = 23,     /* [i32]   per kept aligned row: matched ref index or -1 (src/same.py:634-639) */
    SAME_ARR_MATCH_P = 24,     /* [i32]   per kept aligned row: pair index of the match or -1 */
    SAME_ARR_TRI_MASK = 25,    /* [i32]   post-solve bits: 0-2 x-order violation of vertex pairs (0,1),(0,2),(1,2); 3-5 y-order; 8-10 vertex matched (src/violationhelper.py:54-121) */
    SAME_ARR_AREA_BEFORE = 26, /* [f64]   calculate_signed_area on aligned XY (src/same.py:1362-1370) */
    SAME_ARR_AREA_AFTER = 27,  /* [f64]   same on matched reference XY, NaN when a vertex is unmatched (src/same.py:1379-1395) */
    SAME_ARR_FLIPPED = 28,     /* [u8]    area_before*area_after < 0 (src/same.py:1398) */
    SAME_ARR_START_X = 29,     /* [u8]    greedy MIP start: 1 on the chosen pairs (src/init_helpers.py:124-130, x_vars[..].Start) */
    SAME_ARR_START_UNMATCHED = 30,/* [u8] greedy MIP start: 1 on kept aligned rows left unmatched (src/init_helpers.py:132, no_match_vars[..].Start) */
    SAME_ARR_NODE_TRI_PTR = 32,/* [i32]   nKeepA_total+1 BATCH-GLOBAL offsets into NODE_TRI_IDX: the triangles of kept aligned node i start at ptr[i]
                                  (aligned_simplex_map as CSR, src/same.py:1096-1099; a node fills NODE_TRI_LEN[i] slots, a triangle that names
                                  a vertex twice counts once, as in the reference's sets) */
    SAME_ARR_NODE_TRI_LEN = 33,/* [i32]   number of distinct triangles of each kept aligned node */
    SAME_ARR_NODE_TRI_IDX = 34,/* [i32]   3*T_total slots, window-local TRI indices, ascending inside a node */
    SAME_ARR_PAIR_J = 31,      /* [i32]   reference index of each pair = PAIRS[:, 1] on its own.  The aligned index PAIRS[:, 0] is implied by ROW_PTR
                                  (pairs are sorted by aligned row, src/utils.py:720-731), so ROW_PTR + PAIR_J is valid_pairs in half the bytes over PCIe */
    SAME_ARR_PAIR_J16 = 35     /* [u16]   PAIR_J in two bytes per pair — a quarter of PAIRS over PCIe.  Only when every window keeps at most 65,536
                                  reference rows (the index is window-local); SAME_E_LIMIT otherwise: fall back to PAIR_J */
};

typedef struct same_section same_section_t;
typedef struct same_batch same_batch_t;

SAME_API const char *same_last_error(void);
SAME_API int same_abi_version(void);
SAME_API int same_device_count(int *count);

/* ---- section -------------------------------------------------------------------------- */
/* Upload both frames.  XY are [N,2] row-major, prob blocks [N,K] row-major in commonCT order,
 * type = integer code of `cell_type` (same code space for both frames), size = the `size`
 * column (1 when absent, src/same.py:934-939).  `stream` = a cudaStream_t to run on, or NULL
 * for a private stream.  All HD. */
SAME_API int same_section_create(int device, void *stream, int64_t n_aligned, int64_t n_ref, int n_types,
                        const double *a_xy, const double *r_xy, const double *a_prob, const double *r_prob,
                        const int32_t *a_type, const int32_t *r_type, const double *a_size, const double *r_size,
                        same_section_t **out);
/* The coordinates are on the device when same_section_create returns; the other columns (probabilities, type codes, sizes) are
 * uploaded on an auxiliary stream that overlaps the first stages of a batch.  PAGE-LOCKED host buffers are read asynchronously:
 * keep them valid and unmodified until same_section_wait_uploads() (or the first same_batch_candidates on the section) has
 * returned.  Pageable buffers are staged before same_section_create returns and may be released at once. */
SAME_API int same_section_wait_uploads(same_section_t *sec);
SAME_API int same_section_destroy(same_section_t *sec);
/* bounding box of both frames: out[4] = x_min, x_max, y_min, y_max (src/same.py:481-482) */
SAME_API int same_section_bbox(same_section_t *sec, double *out4);
/* Cells of each frame inside M half-open rectangles (x_min,x_max,y_min,y_max), for the driver's
 * small-window merge rule (src/same.py:523-542). rects HD. */
SAME_API int same_section_count_rects(same_section_t *sec, int64_t m, const double *rects, int64_t *cnt_aligned, int64_t *cnt_ref);
/* Precomputed triangulation in vertex-id space (src/same.py:262-290): a_vid[n_aligned] are the
 * `__tri_vid` values (src/same.py:962-970), tri_vid[T,3] the triangles.  HD. */
SAME_API int same_section_set_triangles(same_section_t *sec, const int64_t *a_vid, const int64_t *tri_vid, int64_t n_tri);

/* ---- batch ---------------------------------------------------------------------------- */
/* Cut W windows (subset_data, src/same.py:293-295).  rects == NULL with W == 1 means "the whole section". */
SAME_API int same_batch_create(same_section_t *sec, int64_t n_windows, const double *rects, same_batch_t **out);
SAME_API int same_batch_destroy(same_batch_t *b);
SAME_API int64_t same_batch_num_windows(same_batch_t *b);

/* a1 (+a2) + a3: find_knn_within_radius (src/utils.py:709-742), optional cell-type priority
 * (src/knn_utils.py:31-65), pair costs (src/same.py:1182-1189).  Windows that end with zero pairs
 * are reported through their sizes (the caller raises the reference's ValueError). */
SAME_API int same_batch_candidates(same_batch_t *b, double radius, int knn, int priority, double dist_ct_coeff);

/* a5: remap the section's precomputed triangles onto every window's post-KNN rows
 * (src/same.py:1028-1031). */
SAME_API int same_batch_triangles_remap(same_batch_t *b);
/* ...or hand in per-window local triangulations (scipy Delaunay of the post-KNN aligned XY,
 * src/same.py:1023): tri[T,3] window-local rows, tri_off[W+1].  HD. */
SAME_API int same_batch_triangles_set(same_batch_t *b, const int32_t *tri, const int64_t *tri_off);

/* a6 step 1: per-triangle radius / min-angle / same-type classification (src/helpers.py:296-344).
 * use_angle = 0 reproduces min_angle_deg=None.  *n_band = number of guard-band triangles. */
SAME_API int same_batch_tri_classify(same_batch_t *b, double radius, int use_angle, double min_angle_deg,
                            int ignore_same_type, int64_t *n_band);
/* optional: overwrite the class of n TRI_IN rows (batch-global indices) after a host re-decision */
SAME_API int same_batch_tri_override(same_batch_t *b, int64_t n, const int32_t *tri_index, const uint8_t *cls);
/* a6 step 2 + a7 + a8 + a9: keep / add back (src/helpers.py:346-389), optionally remove truly
 * unconstrained nodes and renumber pairs + triangles + rows (src/same.py:1056-1085), then weights,
 * source signs, bounds (src/same.py:1128-1146, src/helpers.py:184-210). */
SAME_API int same_batch_tri_finalize(same_batch_t *b, int ignore_same_type, int ensure_min_triangle_per_node,
                            int remove_unconstrained);

/* a4: ref_to_pairs groups and limits (src/helpers.py:105-138).  multiplier < 0 = None
 * (use the largest ref size of the window, src/helpers.py:123-126). */
SAME_API int same_batch_groups(same_batch_t *b, int max_matches, int ref_metacell_match_multiplier);

/* a10: lazy separation for windows [w_lo, w_hi) (src/same.py:621-703).  x HD = solution values of
 * those windows' pairs, concatenated.  Per window: n_viol, n_checked, and the first
 * min(n_viol, cap) cuts as (pair_a, pair_b, pair_c, triangle) in ascending triangle order into
 * cuts[(w - w_lo) * cap * 4 ...].  The caps/threshold logic of src/same.py:671-703 is the caller's. */
SAME_API int same_batch_separation(same_batch_t *b, int64_t w_lo, int64_t w_hi, const double *x, int64_t cap,
                          int64_t *n_viol, int64_t *n_checked, int32_t *cuts);

/* Exact-predicate diagnostic for the two orientation tests of the path, both evaluated in the reference's naive fp64 arithmetic
 * and never changed by this call: which = 0 the source signs (aligned coordinates, src/same.py:1146), which = 1 the LAST
 * same_batch_separation call (matched reference coordinates, src/same.py:658).  *n = number of triangles whose computed
 * determinant lies inside Shewchuk's static error bound for that expression, i.e. whose naive sign MAY differ from the exact
 * sign; the first min(*n, cap, 65536) batch-global TRI indices go to tri_idx (unordered).  The caller decides them exactly
 * (same_b200/helpers.py::exact_orientation_sign) and reports the count of disagreements. */
SAME_API int same_batch_uncertain(same_batch_t *b, int which, int64_t cap, int64_t *n, int32_t *tri_idx);

/* a11 + a12: post-solve analysis of windows [w_lo, w_hi) (src/violationhelper.py:1-134,
 * src/same.py:1355-1408); results through same_batch_get(TRI_MASK / AREA_* / FLIPPED / MATCH_*). */
SAME_API int same_batch_postsolve(same_batch_t *b, int64_t w_lo, int64_t w_hi, const double *x);

/* Stateless form of a11 + a12 for callers that hold plain arrays (verify_spatial_preservation called on its
 * own, src/violationhelper.py:1-134): tri[T,3] rows of a_xy, match_j[n_aligned] = matched row of r_xy or -1.
 * Outputs (host): mask[T] i32 (bit layout of SAME_ARR_TRI_MASK), area_before[T], area_after[T], flipped[T] u8. */
SAME_API int same_postsolve_arrays(int device, int64_t n_tri, const int32_t *tri, int64_t n_aligned, const double *a_xy, int64_t n_ref,
                                   const double *r_xy, const int32_t *match_j, int32_t *mask, double *area_before, double *area_after,
                                   uint8_t *flipped);

/* Greedy MIP start of every window of the batch (compute_mip_start_pairs(init_method='greedy'), src/init_helpers.py:110-132):
 * pairs in ascending (cost, pair index) order, a pair is chosen iff its aligned row prefers a match (best cost of the row
 * < no_match_penalty * size) and neither endpoint is taken yet.  Results through same_batch_get(START_X / START_UNMATCHED).
 * *rounds (may be NULL) = parallel rounds the selection took. */
SAME_API int same_batch_mip_start(same_batch_t *b, double no_match_penalty, int32_t *rounds);

/* The selection loop on its own, for callers that hold plain arrays: n items with `degree` (1..3) endpoints each
 * (nodes[n, degree], values in [0, n_nodes)), visited in ascending (key, item index) order; an item is selected iff it is
 * eligible (eligible == NULL: all are) and none of its endpoints belongs to an item selected earlier.  Used by the greedy
 * MIP start (degree 2) and by the batch selection of greedy_triangle_collapse (degree 3, key = perimeter,
 * src/metacell_utils.py:423-433).  nodes/key/eligible HD; selected[n], used[n_nodes] (may be NULL) host, u8. */
SAME_API int same_greedy_select(int device, int64_t n, int degree, const int32_t *nodes, const double *key, const uint8_t *eligible,
                                int64_t n_nodes, uint8_t *selected, uint8_t *used, int32_t *rounds);

/* One collapse iteration of greedy_triangle_collapse (src/metacell_utils.py:388-433) on the current metacells: a triangle is a
 * candidate iff its three vertices have the same type code and their sizes sum to <= max_size; candidates are visited in
 * ascending (perimeter, triangle index) order and selected iff no vertex is used yet.  The perimeter reproduces the reference's
 * arithmetic (np.linalg.norm of each side = sqrt(fma(dy, dy, dx*dx)), summed left to right).  xy[n,2], type[n], size[n],
 * tri[T,3] HD; selected[T] u8, perimeter[T] (may be NULL) host. */
SAME_API int same_collapse_select(int device, int64_t n, const double *xy, const int32_t *type, const double *size, int64_t n_tri,
                                  const int32_t *tri, double max_size, uint8_t *selected, double *perimeter, int32_t *rounds);

/* Member means of merged metacells (src/metacell_utils.py:446-474): out[g, c] = mean over the rows pos[ptr[g] .. ptr[g+1]) of
 * values[row, c], summed in the order pandas / numpy sum them (pairwise summation), divided by the member count; NaN for an
 * empty group.  values[n_rows, n_cols] row-major, ptr[n_groups + 1], pos[ptr[n_groups]] HD; out[n_groups, n_cols] host.
 * Inputs must not contain NaN (pandas would skip them; the caller keeps such columns on its own path). */
SAME_API int same_segment_mean(int device, int64_t n_rows, int64_t n_cols, const double *values, int64_t n_groups, const int64_t *ptr,
                               const int32_t *pos, double *out);

/* ---- results -------------------------------------------------------------------------- */
/* W+1 offsets (in elements) of array `what` */
SAME_API int same_batch_offsets(same_batch_t *b, int what, int64_t *off);
/* total number of elements of array `what` */
SAME_API int same_batch_length(same_batch_t *b, int what, int64_t *n);
/* copy elements [elem_lo, elem_hi) of array `what` to dst (host or device) */
SAME_API int same_batch_get(same_batch_t *b, int what, int64_t elem_lo, int64_t elem_hi, void *dst);
/* n copies issued back to back on the batch's stream with ONE synchronisation at the end:
 * array what[k], elements [lo[k], hi[k]) -> dst[k] (host, ideally pinned, or device) */
SAME_API int same_batch_get_many(same_batch_t *b, int64_t n, const int32_t *what, const int64_t *lo, const int64_t *hi, void *const *dst);
/* The same copies WITHOUT the final synchronisation: the call returns once they are queued on the batch's stream, so the caller can
 * start the next section on another stream while this one's results cross PCIe (full duplex with its uploads).  dst must be
 * page-locked (or device) memory and stay valid until same_batch_sync(b) has returned; only then may it be read. */
SAME_API int same_batch_get_many_async(same_batch_t *b, int64_t n, const int32_t *what, const int64_t *lo, const int64_t *hi, void *const *dst);
/* page-locked host memory for inputs/outputs (cudaHostAlloc / cudaFreeHost) */
SAME_API int same_pinned_alloc(int64_t bytes, void **out);
SAME_API int same_pinned_free(void *p);
/* element size in bytes of array `what` */
SAME_API int64_t same_elem_size(int what);
/* block until everything queued on the batch's stream has finished */
SAME_API int same_batch_sync(same_batch_t *b);
/* A CUDA stream owned by the caller (cudaStreamCreateWithFlags(cudaStreamNonBlocking) / cudaStreamDestroy) for hosts without their
 * own CUDA binding: pass it as `stream` to same_section_create.  Sections that rotate through a FIXED set of such streams
 * overlap one section's downloads with the next one's uploads and kernels, and the stream-ordered memory pool recycles their
 * buffers without touching the driver (a new stream per section does not: its first allocations cannot reuse memory that
 * was freed on other streams). */
SAME_API int same_stream_create(int device, void **stream);
SAME_API int same_stream_destroy(int device, void *stream);
/* How host threads wait for the device: 0 (default) = spin (cudaStreamSynchronize: lowest latency), 1 = sleep on a blocking-sync
 * event (yields the core: for hosts where many ranks / section threads share few cores).  Also SAME_B200_HOST_WAIT=yield. */
SAME_API int same_set_host_wait(int yield);
/* The library's own memory checker (for hosts where compute-sanitizer cannot run): enable = 1 / 0 switches guard mode for buffers
 * allocated from now on (-1 = leave as is; 2 = self-test: overrun a guarded probe buffer by one word on purpose, the corrupted count
 * must rise by one; SAME_B200_GUARD=1 sets the mode at load time).  In guard mode every device buffer has a
 * 256-byte canary zone on both sides, verified when the buffer is released, and its body is filled with 0xCD on every
 * (re)allocation, so that reads of memory the library did not write change the results.  Waits for the device, then returns the
 * number of buffers whose zones were found corrupted and the number of buffers checked since the library was loaded. */
SAME_API int same_debug_guard(int enable, int64_t *corrupted, int64_t *checked);
/* Counters of a batch: SAME_STAT_KNN_EVALUATIONS = distance evaluations of the last same_batch_candidates (counted only while
 * same_profile_enable(1) is in effect, -1 otherwise): bench.py's compute-side roofline = evaluations x 5 flops / kernel time. */
enum { SAME_STAT_KNN_EVALUATIONS = 1 };
SAME_API int same_batch_stat(same_batch_t *b, int what, int64_t *value);
/* Device memory the library's stream-ordered pool holds (reserved) and has handed out (used) right now, in bytes. */
SAME_API int same_mempool_stats(int device, int64_t *reserved, int64_t *used);
/* Make the pool hold at least `bytes` (one allocation that is freed again at once; the pool never returns memory to the driver):
 * a stream of overlapping sections then does not meet a cudaMalloc — tens of milliseconds when several processes share the host —
 * in the middle of its steady state.  Blocks until done. */
SAME_API int same_mempool_reserve(int device, int64_t bytes);
/* stream the batch runs on (cudaStream_t), for event timing by the caller */
SAME_API void *same_batch_stream(same_batch_t *b);
/* FP64 vector peak of the device in TFLOP/s, measured with an FMA micro-kernel (eight independent chains per thread, best of
 * three launches): the compute-side roofline denominator bench.py records next to the measured HBM copy peak. */
SAME_API int same_measure_fp64_peak(int device, double *tflops);
/* number of kernel launches issued by this library in this process (bench.py "gpu_launches") */
SAME_API int64_t same_launch_count(void);
/* Per-kernel device timing for bench.py's roofline block: while enabled every launch is bracketed by CUDA
 * events on its stream.  same_profile_report() synchronises, writes "name<TAB>launches<TAB>total_ms" lines into
 * buf (NUL-terminated, truncated to cap) and clears the records; returns the untruncated length or < 0. */
SAME_API int same_profile_enable(int on);
SAME_API int64_t same_profile_report(char *buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* SAME_B200_H */
