// separation.cu — a10 lazy-constraint separation (src/same.py:621-703), a11 x/y-order diagnostics
// (src/violationhelper.py:1-134), a12 signed-area flip analysis (src/same.py:1355-1408, src/helpers.py:73-77).
//
// One thread per Delaunay triangle.  The orientation test reproduces the reference's naive fp64
// expression one IEEE operation at a time (no FMA), so the violated set is bit-identical to the
// Python callback; an exact predicate would change the set and is deliberately not used.
#include "common.cuh"

namespace same {

// matching[i] = j of the LAST pair of row i with x > 0.5 (dict overwrite, src/same.py:636-639)
__global__ void k_match_rows(const double *__restrict__ x, i32 x_base, const int2 *__restrict__ pairs, const i32 *__restrict__ row_ptr,
                             const i32 *__restrict__ ka_off, const i32 *__restrict__ p_off, int W, i32 k_lo, i32 k_hi,
                             i32 *__restrict__ match_j, i32 *__restrict__ match_p) {
    const i32 k = k_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_hi) return;
    i32 mp = -1, mj = -1;
    const i32 e = row_ptr[k + 1];
    for (i32 p = row_ptr[k]; p < e; ++p)
        if (x[p - x_base] > 0.5) { mp = p; mj = pairs[p].y; }
    if (mp >= 0) mp -= p_off[find_window(ka_off, W, k)];
    match_j[k] = mj;
    match_p[k] = mp;
}

__global__ void __launch_bounds__(256) k_separation(const int3 *__restrict__ tri, const signed char *__restrict__ src_sign, i32 t_lo, i32 t_hi,
                                                    const i32 *__restrict__ t_off, const i32 *__restrict__ ka_off, const i32 *__restrict__ kr_off,
                                                    int W, const i32 *__restrict__ match_j, const double2 *__restrict__ kr_xy,
                                                    i32 *__restrict__ viol_flag, i32 *__restrict__ checked /* per window */) {
    const i32 t = t_lo + blockIdx.x * blockDim.x + threadIdx.x;
    int ck = 0, w = 0;
    if (t < t_hi) {
        w = find_window(t_off, W, t);
        const i32 nb = ka_off[w], rb = kr_off[w];
        const int3 v = tri[t];
        const i32 ja = match_j[nb + v.x], jb = match_j[nb + v.y], jc = match_j[nb + v.z];
        int viol = 0;
        if (ja >= 0 && jb >= 0 && jc >= 0) {  // same.py:649-651
            const double2 A = kr_xy[rb + ja], B = kr_xy[rb + jb], C = kr_xy[rb + jc];
            const int rs = sign_of(orient_naive(A.x, A.y, B.x, B.y, C.x, C.y));  // same.py:658
            const int ss = src_sign[t];
            if (ss != 0 && rs != 0) {  // same.py:663-664
                ck = 1;
                viol = ss != rs;  // same.py:669
            }
        }
        viol_flag[t - t_lo] = viol;
    } else if (t == t_hi) viol_flag[t - t_lo] = 0;
    // block-aggregated count when the whole block sits in one window
    __shared__ int w0, same_w;
    if (threadIdx.x == 0) { w0 = w; same_w = 1; }
    __syncthreads();
    if (t < t_hi && w != w0) same_w = 0;
    __syncthreads();
    if (same_w) {
        const int tot = __syncthreads_count(ck);
        if (threadIdx.x == 0 && tot) atomicAdd(checked + w0, tot);
    } else if (ck) atomicAdd(checked + w, 1);
}

__global__ void k_emit_cuts(const int3 *__restrict__ tri, i32 t_lo, i32 t_hi, const i32 *__restrict__ t_off, const i32 *__restrict__ ka_off, int W,
                            int w_lo, const i32 *__restrict__ viol_flag, const i32 *__restrict__ vpos, const i32 *__restrict__ match_p, i64 cap,
                            i32 *__restrict__ cuts) {
    const i32 t = t_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_hi || !viol_flag[t - t_lo]) return;
    const int w = find_window(t_off, W, t);
    const i64 r = vpos[t - t_lo] - vpos[t_off[w] - t_lo];
    if (r >= cap) return;
    const i32 nb = ka_off[w];
    const int3 v = tri[t];
    reinterpret_cast<int4 *>(cuts)[(i64)(w - w_lo) * cap + r] = make_int4(match_p[nb + v.x], match_p[nb + v.y], match_p[nb + v.z], t - t_off[w]);
}

__global__ void k_sep_counts(const i32 *__restrict__ vpos, const i32 *__restrict__ t_off, i32 t_lo, int w_lo, int w_hi, const i32 *__restrict__ checked,
                             i32 *__restrict__ out /* 2*(w_hi-w_lo): n_viol, n_checked */) {
    const int w = w_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= w_hi) return;
    out[2 * (w - w_lo)] = vpos[t_off[w + 1] - t_lo] - vpos[t_off[w] - t_lo];
    out[2 * (w - w_lo) + 1] = checked[w];
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
        LAUNCH(k_match_rows, blocks_for(k_hi - k_lo, 256), 256, 0, s, xd, (i32)b->p_off[w_lo], b->pairs.p, b->row_ptr.p, b->d_ka_off.p,
               b->d_p_off.p, (int)b->W, k_lo, k_hi, b->match_j.p, b->match_p.p);
}

void batch_separation(Batch *b, i64 w_lo, i64 w_hi, const double *x, i64 cap, i64 *n_viol, i64 *n_checked, i32 *cuts) {
    cudaStream_t s = b->stream;
    REQUIRE(b->stage >= 4, SAME_E_STATE, "same_batch_separation before same_batch_tri_finalize");
    REQUIRE(w_lo >= 0 && w_hi <= b->W && w_lo < w_hi, SAME_E_ARG, "bad window range");
    REQUIRE(cap >= 0, SAME_E_ARG, "cap must be >= 0");
    const i64 nw = w_hi - w_lo;
    run_matching(b, w_lo, w_hi, resolve_x(b, w_lo, w_hi, x));
    const i32 t_lo = (i32)b->t_off[w_lo], t_hi = (i32)b->t_off[w_hi];
    const i64 nt = t_hi - t_lo;
    b->viol_flag.alloc(nt + 1, s); b->viol_pos.alloc(nt + 1, s);
    b->sep_counts.alloc(b->W + 2 * nw, s);
    b->cuts.alloc(std::max<i64>(1, nw * cap * 4), s);
    CK(cudaMemsetAsync(b->sep_counts.p, 0, sizeof(i32) * (b->W + 2 * nw), s));
    LAUNCH(k_separation, blocks_for(nt + 1, 256), 256, 0, s, b->tri.p, b->t_sign.p, t_lo, t_hi, b->d_t_off.p, b->d_ka_off.p, b->d_kr_off.p, (int)b->W,
           b->match_j.p, b->kr_xy.p, b->viol_flag.p, b->sep_counts.p);
    exclusive_scan_i32(b->viol_flag.p, b->viol_pos.p, nt + 1, b->scratch, s);
    if (nt > 0 && cap > 0)
        LAUNCH(k_emit_cuts, blocks_for(nt, 256), 256, 0, s, b->tri.p, t_lo, t_hi, b->d_t_off.p, b->d_ka_off.p, (int)b->W, (int)w_lo, b->viol_flag.p,
               b->viol_pos.p, b->match_p.p, cap, b->cuts.p);
    i32 *out = b->sep_counts.p + b->W;
    LAUNCH(k_sep_counts, blocks_for(nw, 128), 128, 0, s, b->viol_pos.p, b->d_t_off.p, t_lo, (int)w_lo, (int)w_hi, b->sep_counts.p, out);
    std::vector<i32> h(2 * nw);
    CK(cudaMemcpyAsync(h.data(), out, sizeof(i32) * 2 * nw, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    i64 max_fill = 0;
    for (i64 w = 0; w < nw; ++w) {
        n_viol[w] = h[2 * w];
        n_checked[w] = h[2 * w + 1];
        max_fill = std::max<i64>(max_fill, std::min<i64>(h[2 * w], cap));
    }
    if (cuts && cap > 0 && max_fill > 0) {
        if (nw == 1) {  // the callback case: copy only what was filled
            CK(cudaMemcpyAsync(cuts, b->cuts.p, sizeof(i32) * 4 * (size_t)max_fill, cudaMemcpyDefault, s));
        } else {
            CK(cudaMemcpyAsync(cuts, b->cuts.p, sizeof(i32) * 4 * (size_t)(nw * cap), cudaMemcpyDefault, s));
        }
        CK(cudaStreamSynchronize(s));
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
    run_matching(b, w_lo, w_hi, resolve_x(b, w_lo, w_hi, x));
    if (!b->have_post) {
        b->t_mask.alloc(b->T, s); b->area_before.alloc(b->T, s); b->area_after.alloc(b->T, s); b->flipped.alloc(b->T, s);
        b->have_post = true;
    }
    const i32 t_lo = (i32)b->t_off[w_lo], t_hi = (i32)b->t_off[w_hi];
    if (t_hi > t_lo)
        LAUNCH(k_postsolve, blocks_for(t_hi - t_lo, 256), 256, 0, s, b->tri.p, t_lo, t_hi, b->d_t_off.p, b->d_ka_off.p, b->d_kr_off.p, (int)b->W,
               b->match_j.p, b->ka_xy.p, b->kr_xy.p, b->t_mask.p, b->area_before.p, b->area_after.p, b->flipped.p);
    CK(cudaStreamSynchronize(s));
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
        CK(cudaStreamSynchronize(s));
    } catch (...) {
        cudaStreamSynchronize(s);
        cudaStreamDestroy(s);
        throw;
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamDestroy(s));
}

}  // namespace same
