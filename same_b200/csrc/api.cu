// api.cu — the extern "C" surface declared in include/same_b200.h.
#include "common.cuh"

using namespace same;

namespace {

template <typename F>
int guarded(F &&f) {
    try {
        f();
        return SAME_OK;
    } catch (const same::Error &e) {
        g_err = e.what();
        cudaGetLastError();  // do not leave a non-sticky error pending for the next call
        return e.code;
    } catch (const std::exception &e) {
        g_err = e.what();
        cudaGetLastError();
        return SAME_E_CUDA;
    }
}

struct ArrayView {
    const void *p;
    i64 n;         // elements
    i64 esize;     // bytes per element
    const std::vector<i64> *off;  // per-window offsets (may be null)
};

i64 elem_size(int what) {
    switch (what) {
        case SAME_ARR_PAIRS: return 8;
        case SAME_ARR_PAIR_J: return 4;
        case SAME_ARR_PAIR_J16: return 2;
        case SAME_ARR_COST: case SAME_ARR_TRI_WEIGHT: case SAME_ARR_AREA_BEFORE: case SAME_ARR_AREA_AFTER: return 8;
        case SAME_ARR_TRI_IN: case SAME_ARR_TRI: return 12;
        case SAME_ARR_TRI_CLASS: case SAME_ARR_TRI_SIGN: case SAME_ARR_FLIPPED: case SAME_ARR_START_X: case SAME_ARR_START_UNMATCHED: return 1;
        case SAME_ARR_TRI_BOUNDS: return 32;
        case SAME_ARR_TRI_ARGV: return 16;
        case SAME_ARR_REF_GROUP_PTR: case SAME_ARR_ROW_PTR: return 4;
        default: return 4;
    }
}

ArrayView view(Batch *b, int what) {
    batch_settle(b);
    auto need = [&](int st, const char *m) { REQUIRE(b->stage >= st, SAME_E_STATE, m); };
    const i64 es = elem_size(what);
    switch (what) {
        case SAME_ARR_WIN_A: return {b->a_src.p, b->nAi, es, &b->a_off};
        case SAME_ARR_WIN_R: return {b->r_src.p, b->nRi, es, &b->r_off};
        case SAME_ARR_KEEP_A: need(1, "candidates not run"); return {b->keepA.p, b->nKA, es, &b->ka_off};
        case SAME_ARR_KEEP_R: need(1, "candidates not run"); return {b->keepR.p, b->nKR, es, &b->kr_off};
        case SAME_ARR_PAIRS: need(1, "candidates not run"); return {b->pairs.p, b->P, es, &b->p_off};
        case SAME_ARR_COST: need(1, "candidates not run"); return {b->cost.p, b->P, es, &b->p_off};
        case SAME_ARR_PAIR_J: need(1, "candidates not run"); batch_pair_j(b); return {b->pair_j.p, b->P, es, &b->p_off};
        case SAME_ARR_PAIR_J16: need(1, "candidates not run"); batch_pair_j16(b); return {b->pair_j16.p, b->P, es, &b->p_off};
        case SAME_ARR_ROW_PTR: need(1, "candidates not run"); return {b->row_ptr.p, b->nKA + 1, es, nullptr};
        case SAME_ARR_REF_GROUP_NODE: REQUIRE(b->have_groups, SAME_E_STATE, "groups not built"); return {b->g_node.p, b->G, es, &b->g_off};
        case SAME_ARR_REF_GROUP_LIMIT: REQUIRE(b->have_groups, SAME_E_STATE, "groups not built"); return {b->g_limit.p, b->G, es, &b->g_off};
        case SAME_ARR_REF_GROUP_PTR: REQUIRE(b->have_groups, SAME_E_STATE, "groups not built"); return {b->g_ptr.p, b->G + 1, es, nullptr};
        case SAME_ARR_REF_GROUP_IDX: REQUIRE(b->have_groups, SAME_E_STATE, "groups not built"); return {b->g_idx.p, b->P, es, &b->p_off};
        case SAME_ARR_TRI_IN: need(2, "no triangles"); return {b->tin.p, b->Tin, es, &b->tin_off};
        case SAME_ARR_TRI_IN_SRC: need(2, "no triangles"); REQUIRE(b->tin_has_src, SAME_E_STATE, "triangles were not remapped"); return {b->tin_src.p, b->Tin, es, &b->tin_off};
        case SAME_ARR_TRI_CLASS: need(3, "triangles not classified"); return {b->cls.p, b->Tin, es, &b->tin_off};
        case SAME_ARR_TRI_BAND: need(3, "triangles not classified"); return {b->band_idx.p, b->n_band, es, nullptr};
        case SAME_ARR_TRI: need(4, "triangles not finalized"); return {b->tri.p, b->T, es, &b->t_off};
        case SAME_ARR_TRI_SRC: need(4, "triangles not finalized"); return {b->tri_src.p, b->T, es, &b->t_off};
        case SAME_ARR_TRI_WEIGHT: need(4, "triangles not finalized"); return {b->t_weight.p, b->T, es, &b->t_off};
        case SAME_ARR_TRI_SIGN: need(4, "triangles not finalized"); return {b->t_sign.p, b->T, es, &b->t_off};
        case SAME_ARR_TRI_BOUNDS: need(4, "triangles not finalized"); return {b->t_bounds.p, b->T, es, &b->t_off};
        case SAME_ARR_TRI_ARGV: need(4, "triangles not finalized"); return {b->t_argv.p, b->T, es, &b->t_off};
        case SAME_ARR_UNCONSTRAINED: need(4, "triangles not finalized"); return {b->unc.p, b->nUnc, es, &b->unc_off};
        case SAME_ARR_MATCH_J: need(4, "no matching yet"); return {b->match_j.p, b->match_j.p ? b->nKA : 0, es, &b->ka_off};
        case SAME_ARR_MATCH_P: need(4, "no matching yet"); return {b->match_p.p, b->match_p.p ? b->nKA : 0, es, &b->ka_off};
        case SAME_ARR_TRI_MASK: REQUIRE(b->have_post, SAME_E_STATE, "postsolve not run"); return {b->t_mask.p, b->T, es, &b->t_off};
        case SAME_ARR_AREA_BEFORE: REQUIRE(b->have_post, SAME_E_STATE, "postsolve not run"); return {b->area_before.p, b->T, es, &b->t_off};
        case SAME_ARR_AREA_AFTER: REQUIRE(b->have_post, SAME_E_STATE, "postsolve not run"); return {b->area_after.p, b->T, es, &b->t_off};
        case SAME_ARR_FLIPPED: REQUIRE(b->have_post, SAME_E_STATE, "postsolve not run"); return {b->flipped.p, b->T, es, &b->t_off};
        case SAME_ARR_NODE_TRI_PTR: need(4, "triangles not finalized"); batch_incidence(b); return {b->nt_ptr.p, b->nKA + 1, es, nullptr};
        case SAME_ARR_NODE_TRI_LEN: need(4, "triangles not finalized"); batch_incidence(b); return {b->nt_len.p, b->nKA, es, &b->ka_off};
        case SAME_ARR_NODE_TRI_IDX: need(4, "triangles not finalized"); batch_incidence(b); return {b->nt_idx.p, 3 * b->T, es, &b->nt_off};
        case SAME_ARR_START_X: REQUIRE(b->have_start, SAME_E_STATE, "mip start not computed"); return {b->start_x.p, b->P, es, &b->p_off};
        case SAME_ARR_START_UNMATCHED: REQUIRE(b->have_start, SAME_E_STATE, "mip start not computed"); return {b->start_unmatched.p, b->nKA, es, &b->ka_off};
        default: throw same::Error(SAME_E_ARG, "unknown array id");
    }
}

}  // namespace

extern "C" {

const char *same_last_error(void) { return g_err.c_str(); }
int same_abi_version(void) { return SAME_B200_ABI_VERSION; }
int same_device_count(int *count) {
    return guarded([&] { CK(cudaGetDeviceCount(count)); });
}
int64_t same_launch_count(void) { return (int64_t)g_launches.load(); }
int same_measure_fp64_peak(int device, double *tflops) {
    return guarded([&] { REQUIRE(tflops, SAME_E_ARG, "NULL output"); *tflops = measure_fp64_peak(device); });
}

int same_profile_enable(int on) {
    return guarded([&] {
        for (auto &r : g_prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        g_prof_recs.clear();
        g_prof = on != 0;
    });
}

int64_t same_profile_report(char *buf, int64_t cap) {
    int64_t written = -1;
    int rc = guarded([&] {
        CK(cudaDeviceSynchronize());
        struct Acc { std::string name; long long n; double ms; };
        std::vector<Acc> acc;
        for (auto &r : g_prof_recs) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, r.a, r.b));
            cudaEventDestroy(r.a); cudaEventDestroy(r.b);
            bool hit = false;
            for (auto &a : acc) if (a.name == r.name) { a.n++; a.ms += ms; hit = true; break; }
            if (!hit) acc.push_back({r.name, 1, (double)ms});
        }
        g_prof_recs.clear();
        std::string out;
        char line[512];
        for (auto &a : acc) { snprintf(line, sizeof(line), "%s\t%lld\t%.6f\n", a.name.c_str(), a.n, a.ms); out += line; }
        written = (int64_t)out.size();
        if (buf && cap > 0) { size_t k = std::min<size_t>(out.size(), (size_t)cap - 1); memcpy(buf, out.data(), k); buf[k] = 0; }
    });
    return rc == SAME_OK ? written : rc;
}
int64_t same_elem_size(int what) { return elem_size(what); }

int same_section_create(int device, void *stream, int64_t n_aligned, int64_t n_ref, int n_types, const double *a_xy, const double *r_xy,
                        const double *a_prob, const double *r_prob, const int32_t *a_type, const int32_t *r_type, const double *a_size,
                        const double *r_size, same_section_t **out) {
    return guarded([&] {
        REQUIRE(out, SAME_E_ARG, "out is NULL");
        REQUIRE(n_aligned >= 0 && n_ref >= 0 && n_types >= 0, SAME_E_ARG, "negative size");
        REQUIRE(n_aligned < (1ll << 31) && n_ref < (1ll << 31), SAME_E_LIMIT, "frames are limited to 2^31 rows");
        REQUIRE((n_aligned == 0 || a_xy) && (n_ref == 0 || r_xy), SAME_E_ARG, "XY pointer is NULL");
        REQUIRE(n_types == 0 || ((n_aligned == 0 || a_prob) && (n_ref == 0 || r_prob)), SAME_E_ARG, "probability pointer is NULL");
        CK(cudaSetDevice(device));
        Section *sec = new Section();
        try {
            sec->device = device;
            if (stream) sec->stream = (cudaStream_t)stream;
            else { CK(cudaStreamCreateWithFlags(&sec->stream, cudaStreamNonBlocking)); sec->own_stream = true; }
            cudaMemPool_t pool;
            CK(cudaDeviceGetDefaultMemPool(&pool, device));
            unsigned long long thr = ~0ull;  // keep freed blocks cached: the pipeline reallocates the same sizes every window batch
            CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
            sec->nA = n_aligned; sec->nR = n_ref; sec->K = n_types;
            section_build(sec, a_xy, r_xy, a_prob, r_prob, a_type, r_type, a_size, r_size);
        } catch (...) {
            cudaStream_t aux = sec->aux_stream;
            cudaEvent_t ev = sec->aux_ready;
            if (aux) stream_wait(aux);   // before the buffers it writes are freed
            delete sec;
            if (aux) cudaStreamDestroy(aux);
            if (ev) cudaEventDestroy(ev);
            throw;
        }
        *out = (same_section_t *)sec;
    });
}

int same_section_wait_uploads(same_section_t *h) {
    return guarded([&] {
        REQUIRE(h, SAME_E_ARG, "section is NULL");
        Section *sec = (Section *)h;
        CK(cudaSetDevice(sec->device));
        if (sec->aux_ready) CK(cudaEventSynchronize(sec->aux_ready));
    });
}

int same_section_destroy(same_section_t *h) {
    return guarded([&] {
        Section *sec = (Section *)h;
        if (!sec) return;
        CK(cudaSetDevice(sec->device));
        cudaStream_t s = sec->stream, aux = sec->aux_stream;
        cudaEvent_t ev = sec->aux_ready;
        bool own = sec->own_stream;
        CK(stream_wait(s));
        if (aux) CK(stream_wait(aux));   // the uploads write buffers that are freed (stream-ordered, on `s`) right below
        delete sec;
        CK(stream_wait(s));
        if (aux) { CK(stream_wait(aux)); CK(cudaStreamDestroy(aux)); }
        if (ev) CK(cudaEventDestroy(ev));
        if (own) CK(cudaStreamDestroy(s));
    });
}

int same_section_bbox(same_section_t *h, double *out4) {
    return guarded([&] {
        REQUIRE(h && out4, SAME_E_ARG, "NULL argument");
        memcpy(out4, ((Section *)h)->bbox, sizeof(double) * 4);
    });
}

int same_section_count_rects(same_section_t *h, int64_t m, const double *rects, int64_t *cnt_aligned, int64_t *cnt_ref) {
    return guarded([&] {
        REQUIRE(h && (m == 0 || (rects && cnt_aligned && cnt_ref)), SAME_E_ARG, "NULL argument");
        REQUIRE(m >= 0 && m < 65536, SAME_E_LIMIT, "at most 65535 rectangles per call");
        if (m == 0) return;
        CK(cudaSetDevice(((Section *)h)->device));
        section_count_rects((Section *)h, m, rects, cnt_aligned, cnt_ref);
    });
}

int same_section_set_triangles(same_section_t *h, const int64_t *a_vid, const int64_t *tri_vid, int64_t n_tri) {
    return guarded([&] {
        REQUIRE(h && (n_tri == 0 || tri_vid), SAME_E_ARG, "NULL argument");
        REQUIRE(n_tri >= 0 && 3 * n_tri < (1ll << 31), SAME_E_LIMIT, "too many triangles");
        CK(cudaSetDevice(((Section *)h)->device));
        section_set_triangles((Section *)h, a_vid, tri_vid, n_tri);
    });
}

int same_batch_create(same_section_t *h, int64_t n_windows, const double *rects, same_batch_t **out) {
    return guarded([&] {
        REQUIRE(h && out, SAME_E_ARG, "NULL argument");
        REQUIRE(n_windows >= 1 && n_windows < 65536, SAME_E_LIMIT, "a batch holds 1..65535 windows");
        REQUIRE(rects || n_windows == 1, SAME_E_ARG, "rects is NULL");
        Section *sec = (Section *)h;
        CK(cudaSetDevice(sec->device));
        Batch *b = new Batch();
        try {
            b->sec = sec;
            b->stream = sec->stream;
            b->W = n_windows;
            if (rects) b->rects.assign(rects, rects + 4 * n_windows);
            else b->rects = {-INFINITY, INFINITY, -INFINITY, INFINITY};
            batch_subset(b);
        } catch (...) {
            batch_pin_release(b);
            delete b;
            throw;
        }
        *out = (same_batch_t *)b;
    });
}

int same_batch_destroy(same_batch_t *h) {
    return guarded([&] {
        Batch *b = (Batch *)h;
        if (!b) return;
        CK(cudaSetDevice(b->sec->device));
        cudaStream_t s = b->stream;
        CK(stream_wait(s));   // nothing queued may still read the batch's page-locked staging blocks
        batch_pin_release(b);
        delete b;
        CK(stream_wait(s));
    });
}

int64_t same_batch_num_windows(same_batch_t *h) { return h ? ((Batch *)h)->W : 0; }
void *same_batch_stream(same_batch_t *h) { return h ? (void *)((Batch *)h)->stream : nullptr; }

#define BATCH_CALL(h, body)                              \
    return guarded([&] {                                 \
        REQUIRE(h, SAME_E_ARG, "batch is NULL");         \
        Batch *b = (Batch *)h;                           \
        CK(cudaSetDevice(b->sec->device));               \
        body;                                            \
    })

int same_batch_candidates(same_batch_t *h, double radius, int knn, int priority, double dist_ct_coeff) {
    BATCH_CALL(h, batch_candidates(b, radius, knn, priority, dist_ct_coeff));
}
int same_batch_triangles_remap(same_batch_t *h) { BATCH_CALL(h, batch_triangles_remap(b)); }
int same_batch_triangles_set(same_batch_t *h, const int32_t *tri, const int64_t *tri_off) {
    BATCH_CALL(h, { REQUIRE(tri_off, SAME_E_ARG, "tri_off is NULL"); batch_triangles_set(b, tri, tri_off); });
}
int same_batch_tri_classify(same_batch_t *h, double radius, int use_angle, double min_angle_deg, int ignore_same_type, int64_t *n_band) {
    BATCH_CALL(h, { batch_tri_classify(b, radius, use_angle, min_angle_deg, ignore_same_type); if (n_band) *n_band = b->n_band; });
}
int same_batch_tri_override(same_batch_t *h, int64_t n, const int32_t *tri_index, const uint8_t *cls) {
    BATCH_CALL(h, batch_tri_override(b, n, tri_index, cls));
}
int same_batch_tri_finalize(same_batch_t *h, int ignore_same_type, int ensure_min, int remove_unconstrained) {
    BATCH_CALL(h, batch_tri_finalize(b, ignore_same_type, ensure_min, remove_unconstrained));
}
int same_batch_groups(same_batch_t *h, int max_matches, int multiplier) { BATCH_CALL(h, batch_groups(b, max_matches, multiplier)); }
int same_batch_separation(same_batch_t *h, int64_t w_lo, int64_t w_hi, const double *x, int64_t cap, int64_t *n_viol, int64_t *n_checked,
                          int32_t *cuts) {
    BATCH_CALL(h, { REQUIRE(n_viol && n_checked, SAME_E_ARG, "NULL output"); batch_separation(b, w_lo, w_hi, x, cap, n_viol, n_checked, cuts); });
}
int same_batch_postsolve(same_batch_t *h, int64_t w_lo, int64_t w_hi, const double *x) { BATCH_CALL(h, batch_postsolve(b, w_lo, w_hi, x)); }
int same_batch_uncertain(same_batch_t *h, int which, int64_t cap, int64_t *n, int32_t *tri_idx) { BATCH_CALL(h, batch_uncertain(b, which, cap, n, tri_idx)); }

int same_batch_mip_start(same_batch_t *h, double no_match_penalty, int32_t *rounds) { BATCH_CALL(h, batch_mip_start(b, no_match_penalty, rounds)); }

int same_greedy_select(int device, int64_t n, int degree, const int32_t *nodes, const double *key, const uint8_t *eligible, int64_t n_nodes,
                       uint8_t *selected, uint8_t *used, int32_t *rounds) {
    return guarded([&] {
        REQUIRE(n >= 0 && n_nodes >= 0 && n < (1ll << 31) && n_nodes < (1ll << 31), SAME_E_ARG, "bad size");
        REQUIRE(degree >= 1 && degree <= 3, SAME_E_ARG, "degree must be 1, 2 or 3");
        REQUIRE(n == 0 || (nodes && key && selected), SAME_E_ARG, "NULL argument");
        greedy_select_arrays(device, n, degree, nodes, key, eligible, n_nodes, selected, used, rounds);
    });
}

int same_collapse_select(int device, int64_t n, const double *xy, const int32_t *type, const double *size, int64_t n_tri, const int32_t *tri,
                         double max_size, uint8_t *selected, double *perimeter, int32_t *rounds) {
    return guarded([&] {
        REQUIRE(n >= 0 && n_tri >= 0 && n < (1ll << 31) && 3 * n_tri < (1ll << 31), SAME_E_ARG, "bad size");
        REQUIRE(n_tri == 0 || (xy && type && size && tri && selected), SAME_E_ARG, "NULL argument");
        collapse_select_arrays(device, n, xy, type, size, n_tri, tri, max_size, selected, perimeter, rounds);
    });
}

int same_segment_mean(int device, int64_t n_rows, int64_t n_cols, const double *values, int64_t n_groups, const int64_t *ptr, const int32_t *pos,
                      double *out) {
    return guarded([&] {
        REQUIRE(n_rows >= 0 && n_cols >= 0 && n_groups >= 0, SAME_E_ARG, "negative size");
        REQUIRE(n_rows < (1ll << 31), SAME_E_LIMIT, "at most 2^31 rows");
        if (n_groups == 0 || n_cols == 0) return;
        REQUIRE(values && ptr && out, SAME_E_ARG, "NULL argument");
        const int64_t n_members = ptr[n_groups];    // ptr is host memory in every caller of this stateless form
        REQUIRE(n_members >= 0 && (n_members == 0 || pos), SAME_E_ARG, "bad member list");
        segment_mean_arrays(device, n_rows, n_cols, values, n_groups, ptr, n_members, pos, out);
    });
}

int same_postsolve_arrays(int device, int64_t n_tri, const int32_t *tri, int64_t n_aligned, const double *a_xy, int64_t n_ref, const double *r_xy,
                          const int32_t *match_j, int32_t *mask, double *area_before, double *area_after, uint8_t *flipped) {
    return guarded([&] {
        REQUIRE(n_tri >= 0 && n_aligned >= 0 && n_ref >= 0, SAME_E_ARG, "negative size");
        REQUIRE(n_tri == 0 || (tri && mask && area_before && area_after && flipped), SAME_E_ARG, "NULL argument");
        postsolve_arrays(device, n_tri, tri, n_aligned, a_xy, n_ref, r_xy, match_j, mask, area_before, area_after, flipped);
    });
}

int same_batch_offsets(same_batch_t *h, int what, int64_t *off) {
    BATCH_CALL(h, {
        REQUIRE(off, SAME_E_ARG, "off is NULL");
        ArrayView v = view(b, what);
        REQUIRE(v.off != nullptr, SAME_E_ARG, "array has no per-window offsets");
        memcpy(off, v.off->data(), sizeof(i64) * (size_t)(b->W + 1));
    });
}

int same_batch_length(same_batch_t *h, int what, int64_t *n) {
    BATCH_CALL(h, { REQUIRE(n, SAME_E_ARG, "n is NULL"); *n = view(b, what).n; });
}

int same_batch_get(same_batch_t *h, int what, int64_t elem_lo, int64_t elem_hi, void *dst) {
    BATCH_CALL(h, {
        ArrayView v = view(b, what);
        REQUIRE(elem_lo >= 0 && elem_lo <= elem_hi && elem_hi <= v.n, SAME_E_ARG, "element range out of bounds");
        if (elem_hi > elem_lo) {
            REQUIRE(dst, SAME_E_ARG, "dst is NULL");
            CK(cudaMemcpyAsync(dst, (const char *)v.p + elem_lo * v.esize, (size_t)((elem_hi - elem_lo) * v.esize), cudaMemcpyDefault, b->stream));
            CK(stream_wait(b->stream));
        }
    });
}

// every entry is validated before the first copy is queued: an error must not leave copies into the caller's buffers in flight
static void get_many(Batch *b, int64_t n, const int32_t *what, const int64_t *lo, const int64_t *hi, void *const *dst, bool sync) {
    REQUIRE(n >= 0 && (n == 0 || (what && lo && hi && dst)), SAME_E_ARG, "NULL argument");
    std::vector<ArrayView> views((size_t)n);
    for (int64_t k = 0; k < n; ++k) {
        views[k] = view(b, what[k]);
        REQUIRE(lo[k] >= 0 && lo[k] <= hi[k] && hi[k] <= views[k].n, SAME_E_ARG, "element range out of bounds");
        REQUIRE(hi[k] == lo[k] || dst[k], SAME_E_ARG, "dst is NULL");
    }
    for (int64_t k = 0; k < n; ++k)
        if (hi[k] > lo[k])
            CK(cudaMemcpyAsync(dst[k], (const char *)views[k].p + lo[k] * views[k].esize, (size_t)((hi[k] - lo[k]) * views[k].esize), cudaMemcpyDefault,
                               b->stream));
    if (sync) CK(stream_wait(b->stream));
}

int same_batch_get_many(same_batch_t *h, int64_t n, const int32_t *what, const int64_t *lo, const int64_t *hi, void *const *dst) {
    BATCH_CALL(h, get_many(b, n, what, lo, hi, dst, true));
}
int same_batch_get_many_async(same_batch_t *h, int64_t n, const int32_t *what, const int64_t *lo, const int64_t *hi, void *const *dst) {
    BATCH_CALL(h, get_many(b, n, what, lo, hi, dst, false));
}

int same_pinned_alloc(int64_t bytes, void **out) {
    return guarded([&] {
        REQUIRE(out && bytes >= 0, SAME_E_ARG, "bad argument");
        CK(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable | cudaHostAllocMapped));   // device-addressable: kernels may write results in place
    });
}
int same_pinned_free(void *p) {
    return guarded([&] { if (p) CK(cudaFreeHost(p)); });
}

int same_batch_sync(same_batch_t *h) { BATCH_CALL(h, batch_sync(b)); }

int same_set_host_wait(int yield) {
    same::g_host_wait_yield.store(yield ? 1 : 0);
    return SAME_OK;
}
int same_debug_guard(int enable, int64_t *corrupted, int64_t *checked) {
    return guarded([&] {
        i64 c = 0, k = 0;
        same::guard_report(enable, &c, &k);
        if (corrupted) *corrupted = c;
        if (checked) *checked = k;
    });
}
int same_batch_stat(same_batch_t *h, int what, int64_t *value) {
    BATCH_CALL(h, {
        REQUIRE(value, SAME_E_ARG, "value is NULL");
        REQUIRE(what == SAME_STAT_KNN_EVALUATIONS, SAME_E_ARG, "unknown counter");
        *value = -1;
        if (b->knn_evals.p) {
            unsigned long long v = 0;
            CK(cudaMemcpyAsync(&v, b->knn_evals.p, sizeof(v), cudaMemcpyDeviceToHost, b->stream));
            CK(stream_wait(b->stream));
            *value = (int64_t)v;
        }
    });
}
int same_mempool_stats(int device, int64_t *reserved, int64_t *used) {
    return guarded([&] {
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long r = 0, u = 0;
        CK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &r));
        CK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &u));
        if (reserved) *reserved = (int64_t)r;
        if (used) *used = (int64_t)u;
    });
}
int same_mempool_reserve(int device, int64_t bytes) {
    return guarded([&] {
        REQUIRE(bytes >= 0, SAME_E_ARG, "negative size");
        if (bytes == 0) return;
        CK(cudaSetDevice(device));
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long thr = ~0ull;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        void *p = nullptr;
        CK(cudaMallocAsync(&p, (size_t)bytes, (cudaStream_t)0));
        CK(cudaFreeAsync(p, (cudaStream_t)0));
        CK(stream_wait((cudaStream_t)0));
    });
}
int same_stream_create(int device, void **stream) {
    return guarded([&] {
        REQUIRE(stream, SAME_E_ARG, "stream is NULL");
        CK(cudaSetDevice(device));
        cudaStream_t s;
        CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        *stream = (void *)s;
    });
}
int same_stream_destroy(int device, void *stream) {
    return guarded([&] {
        if (!stream) return;
        CK(cudaSetDevice(device));
        CK(stream_wait((cudaStream_t)stream));
        CK(cudaStreamDestroy((cudaStream_t)stream));
    });
}

}  // extern "C"
