// section.cu — frames resident in HBM, rectangle counting, window subsetting (src/same.py:293-295),
// vertex-id resolution for precomputed triangulations (src/same.py:262-290).
#include "common.cuh"

namespace same {

std::atomic<int> g_host_wait_yield{getenv("SAME_B200_HOST_WAIT") != nullptr && std::string(getenv("SAME_B200_HOST_WAIT")) == "yield" ? 1 : 0};
cudaError_t stream_wait(cudaStream_t s) {
    if (!g_host_wait_yield.load(std::memory_order_relaxed)) return cudaStreamSynchronize(s);
    thread_local cudaEvent_t ev = nullptr;       // one blocking-sync event per host thread (per device: threads here stay on one device)
    thread_local int ev_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!ev || ev_dev != dev) {
        if (ev) cudaEventDestroy(ev);
        e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) { ev = nullptr; return e; }
        ev_dev = dev;
    }
    e = cudaEventRecord(ev, s);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ev);
}
bool g_guard = getenv("SAME_B200_GUARD") != nullptr && getenv("SAME_B200_GUARD")[0] == '1';
static int *g_guard_bad = nullptr;   // page-locked, device-mapped: [0] corrupted zones seen, [1] buffers checked
static std::mutex g_guard_mu;
__global__ void k_guard_check(const unsigned char *__restrict__ raw, size_t body, int *__restrict__ bad) {
    const size_t i = threadIdx.x;   // one block of GUARD_BYTES threads: front zone, then back zone
    const bool broken = raw[i] != 0xA5 || raw[GUARD_BYTES + body + i] != 0xA5;
    if (__syncthreads_or(broken) && threadIdx.x == 0) atomicAdd_system(bad, 1);
    if (threadIdx.x == 0) atomicAdd_system(bad + 1, 1);
}
static int *guard_counters() {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    if (!g_guard_bad) {
        CK(cudaHostAlloc((void **)&g_guard_bad, 2 * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
        g_guard_bad[0] = g_guard_bad[1] = 0;
    }
    return g_guard_bad;
}
void guard_fill(void *raw, size_t body, cudaStream_t s) {
    CK(cudaMemsetAsync(raw, 0xA5, GUARD_BYTES, s));
    CK(cudaMemsetAsync((char *)raw + GUARD_BYTES, 0xCD, body, s));
    CK(cudaMemsetAsync((char *)raw + GUARD_BYTES + body, 0xA5, GUARD_BYTES, s));
}
void guard_check(void *raw, size_t body, cudaStream_t s) {   // (called from destructors: no throw)
    int *bad = nullptr;
    try { bad = guard_counters(); } catch (...) { return; }
    int *dbad = nullptr;
    if (cudaHostGetDevicePointer((void **)&dbad, bad, 0) != cudaSuccess) return;
    k_guard_check<<<1, (unsigned)GUARD_BYTES, 0, s>>>((const unsigned char *)raw, body, dbad);
}
__global__ void k_guard_selftest(int *p, int n) { p[n] = 1; }   // one word past the end: lands in the back zone
void guard_report(int enable, i64 *corrupted, i64 *checked) {
    if (enable == 2) {   // self-test of the checker: a deliberate overrun of a guarded buffer must be counted
        const bool was = g_guard;
        g_guard = true;
        {
            DevBuf<int> probe;
            probe.alloc(100, nullptr);
            k_guard_selftest<<<1, 1>>>(probe.p, 100);
        }
        g_guard = was;
    } else if (enable >= 0) {
        g_guard = enable != 0;
    }
    CK(cudaDeviceSynchronize());
    int *bad = guard_counters();
    if (corrupted) *corrupted = bad[0];
    if (checked) *checked = bad[1];
}
thread_local std::string g_err;
std::atomic<long long> g_launches{0};
bool g_prof = false;
bool g_debug = getenv("SAME_B200_DEBUG") != nullptr && getenv("SAME_B200_DEBUG")[0] == '1';
std::vector<ProfRec> g_prof_recs;
std::mutex g_prof_mu;

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
    CK(stream_wait(s));  // t goes out of scope
}

// ---- small transfers without the copy engines (see common.cuh) ----------------------------------------
static std::mutex g_arena_mu;
static std::vector<std::pair<char *, size_t>> g_arena_pool;
constexpr size_t ARENA_BLOCK = 1 << 20;
void *PinArena::get(size_t bytes) {
    bytes = (bytes + 15) & ~(size_t)15;
    if (blocks.empty() || used + bytes > blocks.back().second) {
        const size_t want = std::max(bytes, ARENA_BLOCK);
        std::pair<char *, size_t> blk{nullptr, 0};
        {
            std::lock_guard<std::mutex> lk(g_arena_mu);
            for (size_t k = 0; k < g_arena_pool.size(); ++k)
                if (g_arena_pool[k].second >= want) { blk = g_arena_pool[k]; g_arena_pool.erase(g_arena_pool.begin() + k); break; }
        }
        if (!blk.first) {
            CK(cudaHostAlloc((void **)&blk.first, want, cudaHostAllocPortable | cudaHostAllocMapped));
            blk.second = want;
        }
        blocks.push_back(blk);
        used = 0;
    }
    void *p = blocks.back().first + used;
    used += bytes;
    return p;
}
void PinArena::release() {
    if (blocks.empty()) return;
    std::lock_guard<std::mutex> lk(g_arena_mu);
    for (auto &b : blocks) g_arena_pool.push_back(b);   // kept for the life of the process: cudaFreeHost synchronises the device
    blocks.clear();
    used = 0;
}
__global__ void k_copy_words(unsigned *__restrict__ dst, const unsigned *__restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
static void copy_words(void *dst, const void *src, size_t bytes, cudaStream_t s) {
    const size_t n = bytes / 4;
    LAUNCH(k_copy_words, (unsigned)std::min<size_t>(64, (n + 255) / 256), 256, 0, s, (unsigned *)dst, (const unsigned *)src, n);
}
void small_d2h(void *host_pinned, const void *dev, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return;
    if (bytes > SMALL_COPY_LIMIT || (bytes & 3)) { CK(cudaMemcpyAsync(host_pinned, dev, bytes, cudaMemcpyDeviceToHost, s)); return; }
    copy_words(host_pinned, dev, bytes, s);
}
struct CopySegs { unsigned *dst[4]; unsigned long long src_word[4], n_word[4]; int n; };
__global__ void k_copy_segments(CopySegs segs, const unsigned *__restrict__ stage) {
    for (int k = 0; k < segs.n; ++k)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < segs.n_word[k]; i += (size_t)gridDim.x * blockDim.x)
            segs.dst[k][i] = stage[segs.src_word[k] + i];
}
void small_h2d_many(PinArena &arena, const SmallCopy *items, int n, cudaStream_t s) {
    size_t total = 0;
    bool ok = n <= 4;
    for (int k = 0; k < n; ++k) { ok = ok && !(items[k].bytes & 3); total += (items[k].bytes + 15) & ~(size_t)15; }
    if (!ok || total > SMALL_COPY_LIMIT) {
        for (int k = 0; k < n; ++k) small_h2d(arena, items[k].dev, items[k].host, items[k].bytes, s);
        return;
    }
    if (total == 0) return;
    char *stage = (char *)arena.get(total);
    CopySegs segs{};
    size_t off = 0, most = 0;
    for (int k = 0; k < n; ++k) {
        if (items[k].bytes == 0) continue;
        memcpy(stage + off, items[k].host, items[k].bytes);
        segs.dst[segs.n] = (unsigned *)items[k].dev; segs.src_word[segs.n] = off / 4; segs.n_word[segs.n] = items[k].bytes / 4;
        most = std::max(most, items[k].bytes / 4);
        ++segs.n;
        off += (items[k].bytes + 15) & ~(size_t)15;
    }
    LAUNCH(k_copy_segments, (unsigned)std::min<size_t>(64, (most + 255) / 256), 256, 0, s, segs, (const unsigned *)stage);
}
void small_h2d(PinArena &arena, void *dev, const void *host, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return;
    if (bytes > SMALL_COPY_LIMIT || (bytes & 3)) { CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s)); return; }
    void *stage = arena.get(bytes);
    memcpy(stage, host, bytes);
    copy_words(dev, stage, bytes, s);
}

// Page-locked staging blocks for the small per-window results live in a process-wide pool and are never handed back to the
// driver while the library is loaded: cudaHostAlloc / cudaFreeHost synchronise the whole device, which would serialise
// sections that are meant to overlap on different streams (CandidateStream).
static std::mutex g_pin_mu;
static std::vector<std::pair<i32 *, i64>> g_pin_pool;
void batch_pin_acquire(Batch *b) {
    const i64 need = 10 * (b->W + 1) + 8;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        for (size_t k = 0; k < g_pin_pool.size(); ++k)
            if (g_pin_pool[k].second >= need) {
                b->pin = g_pin_pool[k].first; b->pin_n = g_pin_pool[k].second;
                g_pin_pool.erase(g_pin_pool.begin() + k);
                return;
            }
    }
    b->pin_n = std::max<i64>(need, 4096);
    CK(cudaHostAlloc((void **)&b->pin, sizeof(i32) * (size_t)b->pin_n, cudaHostAllocPortable | cudaHostAllocMapped));
}
void batch_pin_release(Batch *b) {
    if (b->pin) {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        g_pin_pool.emplace_back(b->pin, b->pin_n);
    }
    b->pin = nullptr;
}

void batch_sync(Batch *b) {
    CK(stream_wait(b->stream));
    const i64 W = b->W;
    if (b->pend_cand) {
        const i32 *h = b->pin_cand();
        b->ka_off.assign(h, h + W + 1);
        b->kr_off.assign(h + W + 1, h + 2 * (W + 1));
        b->p_off.assign(h + 2 * (W + 1), h + 3 * (W + 1));
        b->nKA = b->ka_off[W]; b->nKR = b->kr_off[W]; b->P = b->p_off[W];
        b->pend_cand = false;
    }
    if (b->pend_tin) {
        const i32 *h = b->pin_tin();
        b->tin_off.assign(h, h + W + 1);
        b->pend_tin = false;
    }
    if (b->pend_renum) {
        const i32 *h = b->pin_renum();
        b->p_off.assign(h, h + W + 1);
        b->P = b->p_off[W];
        b->pend_renum = false;
    }
    if (b->pend_groups) {
        const i32 *h = b->pin_groups();
        b->g_off.assign(h, h + W + 1);
        b->G = b->g_off[W];
        b->pend_groups = false;
    }
}
void batch_settle(Batch *b) {
    if (b->pend_cand || b->pend_tin || b->pend_renum || b->pend_groups) batch_sync(b);
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
    // everything else goes up on the auxiliary stream: the synchronisation below (bounding box) then waits for the
    // coordinates only, and the first stages of a batch overlap the rest of the upload.  The buffers themselves are allocated
    // (and later freed) on the section's own stream, so that sections which reuse a stream recycle them through the
    // stream-ordered pool; the auxiliary stream only waits for the allocations and copies.
    sec->a_prob.alloc(nA * K, s); sec->r_prob.alloc(nR * K, s); sec->a_type.alloc(nA, s); sec->r_type.alloc(nR, s);
    sec->a_size.alloc(nA, s); sec->r_size.alloc(nR, s);
    CK(cudaStreamCreateWithFlags(&sec->aux_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&sec->aux_ready, cudaEventDisableTiming));
    cudaStream_t x = sec->aux_stream;
    CK(cudaEventRecord(sec->aux_ready, s));            // (reused below for its real purpose)
    CK(cudaStreamWaitEvent(x, sec->aux_ready, 0));
    auto put = [&](void *dst, const void *src, size_t bytes) {
        if (bytes > 0) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, x));
    };
    put(sec->a_prob.p, a_prob, sizeof(double) * (size_t)(nA * K));
    put(sec->r_prob.p, r_prob, sizeof(double) * (size_t)(nR * K));
    if (a_type) put(sec->a_type.p, a_type, sizeof(i32) * (size_t)nA); else CK(cudaMemsetAsync(sec->a_type.p, 0, sizeof(i32) * (size_t)std::max<i64>(nA, 1), x));
    if (r_type) put(sec->r_type.p, r_type, sizeof(i32) * (size_t)nR); else CK(cudaMemsetAsync(sec->r_type.p, 0, sizeof(i32) * (size_t)std::max<i64>(nR, 1), x));
    if (a_size) put(sec->a_size.p, a_size, sizeof(double) * (size_t)nA);
    else LAUNCH(k_fill_f64, blocks_for(nA, 256), 256, 0, x, sec->a_size.p, nA, 1.0);
    if (r_size) put(sec->r_size.p, r_size, sizeof(double) * (size_t)nR);
    else LAUNCH(k_fill_f64, blocks_for(nR, 256), 256, 0, x, sec->r_size.p, nR, 1.0);
    CK(cudaEventRecord(sec->aux_ready, x));
    // bounding box of both frames (src/same.py:481-482)
    DevBuf<unsigned long long> bb;
    bb.alloc(4, s);
    const unsigned long long init[4] = {~0ull, 0ull, ~0ull, 0ull};
    small_h2d(sec->arena, bb.p, init, sizeof(init), s);
    if (nA > 0) LAUNCH(k_bbox, std::min<unsigned>(blocks_for(nA, 256), 1184), 256, 0, s, sec->a_xy.p, nA, bb.p);
    if (nR > 0) LAUNCH(k_bbox, std::min<unsigned>(blocks_for(nR, 256), 1184), 256, 0, s, sec->r_xy.p, nR, bb.p);
    unsigned long long *h = (unsigned long long *)sec->arena.get(4 * sizeof(unsigned long long));
    small_d2h(h, bb.p, 4 * sizeof(unsigned long long), s);
    CK(stream_wait(s));
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
    CK(stream_wait(s));
    for (i64 i = 0; i < m; ++i) { cntA[i] = (i64)h[i]; cntR[i] = (i64)h[m + i]; }
}

void scan_reserve(Section *sec, i64 words, cudaStream_t s) {
    if (words > sec->scan_state.n || !sec->scan_state.p) {
        sec->scan_state.alloc(std::max<i64>(words, 1 << 16), s);
        sec->scan_state.zero(s);
    }
    if (sec->scan_epoch >= (1u << 30) - 16u) {   // epoch field exhausted: start over on a clean buffer
        sec->scan_state.zero(s);
        sec->scan_epoch = 0;
    }
}
ScanCtx scan_ctx_at(Section *sec, i64 word_offset, i64 tiles) {
    REQUIRE(tiles < (1ll << 31), SAME_E_LIMIT, "too many scan tiles");
    return ScanCtx{sec->scan_state.p + word_offset, (int)std::max<i64>(tiles, 1), ++sec->scan_epoch};
}
ScanCtx scan_ctx(Section *sec, i64 tiles, int n_streams, cudaStream_t s) {
    scan_reserve(sec, std::max<i64>(tiles, 1) * n_streams, s);
    return scan_ctx_at(sec, 0, tiles);
}

// ---- window subsetting: subset_data (src/same.py:293-295) for all windows and BOTH frames at once ------------
// Item = one row of the aligned frame (items [0, nA)) or of the reference frame (items [nA, nA+nR)).  A thread looks
// up the rectangles overlapping the index-grid cell of its point and tests them exactly (half-open).  Pass 1 counts
// the windows holding each row and scans the counts in the same launch (scan.cuh); pass 2 writes, at row_pos[item]..,
// one key (frame | window | row) per hit in ascending window order.  The key stream is then ordered by
// (frame, row, window), so ONE stable radix pass over the (frame | window) bits yields, per frame and window, the rows
// in ascending order — exactly `df[mask]` of the reference — with no atomics and no multi-pass sort.  The sort carries
// the original key position along, which turns into the row -> instance table the triangle remap looks vertices up in.
constexpr int SUBSET_THREADS = 256;
// the counting pass uses big tiles: the look-back chain of scan.cuh costs one L2 round trip per 32 tiles
constexpr int SUBSET_CT = 1024, SUBSET_CI = 4;
constexpr i64 SUBSET_HIST_W = 4096;   // up to this many windows the per-window sizes are counted in shared memory

__device__ __forceinline__ double2 subset_point(const double2 *__restrict__ a_xy, i64 nA, const double2 *__restrict__ r_xy, i64 item) {
    return item < nA ? a_xy[item] : r_xy[item - nA];
}

__global__ void __launch_bounds__(SUBSET_CT) k_subset_count(const double2 *__restrict__ a_xy, i64 nA, const double2 *__restrict__ r_xy, i64 nR,
                                                            const double *__restrict__ rects, RectIndexDev ri, ScanCtx sc, i64 W,
                                                            i32 *__restrict__ row_pos, i32 *__restrict__ wcount /* [2][W] zeroed, then [2] totals */) {
    __shared__ int smem[SUBSET_CT / 32 + 1];
    extern __shared__ int hist[];   // [2][W] window sizes of this block when they fit (hist_on), flushed once at the end
    const bool hist_on = W <= SUBSET_HIST_W;
    if (hist_on) {
        for (i64 k = threadIdx.x; k < 2 * W; k += SUBSET_CT) hist[k] = 0;
        __syncthreads();
    }
    const i64 n = nA + nR;
    const i64 base = ((i64)blockIdx.x * SUBSET_CT + threadIdx.x) * SUBSET_CI;
    int cnt[SUBSET_CI], sum[1] = {0};
#pragma unroll
    for (int k = 0; k < SUBSET_CI; ++k) {
        cnt[k] = 0;
        if (base + k < n) {
            const double2 p = subset_point(a_xy, nA, r_xy, base + k);
            const int c = rect_cell(ri, p.x, p.y);
            const i32 hi = ri.cell_ptr[c + 1];
            const i64 fo = base + k < nA ? 0 : W;
            for (i32 e = ri.cell_ptr[c]; e < hi; ++e) {
                const int w = ri.cell_rects[e];
                if (in_rect(p, rects + 4 * (i64)w)) {   // per-window sizes -> window offsets on the host
                    ++cnt[k];
                    atomicAdd((hist_on ? hist : wcount) + fo + w, 1);
                }
            }
        }
        sum[0] += cnt[k];
    }
    int excl[1], tot[1], pre[1];
    device_exclusive_scan<1, SUBSET_CT>(sc, (int)blockIdx.x, sum, excl, tot, pre, smem);   // (its barriers also complete hist)
    if (hist_on)
        for (i64 k = threadIdx.x; k < 2 * W; k += SUBSET_CT)
            if (hist[k]) atomicAdd(wcount + k, hist[k]);
    int run = excl[0];
#pragma unroll
    for (int k = 0; k < SUBSET_CI; ++k) {
        const i64 item = base + k;
        if (item <= n) row_pos[item] = run;
        if (item == nA) wcount[2 * W + 1] = run;
        if (item == n) wcount[2 * W] = run;
        run += cnt[k];
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(SUBSET_THREADS) k_subset_fill(const double2 *__restrict__ a_xy, i64 nA, const double2 *__restrict__ r_xy, i64 nR,
                                                                const double *__restrict__ rects, RectIndexDev ri, int rowbits, int fbit,
                                                                const i32 *__restrict__ row_pos, KeyT *__restrict__ keys, i32 *__restrict__ vals) {
    const i64 item = (i64)blockIdx.x * SUBSET_THREADS + threadIdx.x;
    if (item >= nA + nR) return;
    const double2 p = subset_point(a_xy, nA, r_xy, item);
    const unsigned long long hi_bits = (item < nA ? 0ull : (1ull << fbit)) | (unsigned long long)(item < nA ? item : item - nA);
    const int c = rect_cell(ri, p.x, p.y);
    const i32 hi = ri.cell_ptr[c + 1];
    i32 out = row_pos[item];
    for (i32 k = ri.cell_ptr[c]; k < hi; ++k) {
        const int w = ri.cell_rects[k];
        if (in_rect(p, rects + 4 * (i64)w)) {
            keys[out] = (KeyT)(hi_bits | ((unsigned long long)w << rowbits));
            vals[out] = out;
            ++out;
        }
    }
}

// sorted key i -> section row of instance i and the (window, instance) entry of its row
template <typename KeyT>
__global__ void k_subset_finish(const KeyT *__restrict__ sorted, const i32 *__restrict__ sorted_vals, i64 total, i64 totalA, int rowbits, int fbit,
                                i32 *__restrict__ a_src, i32 *__restrict__ r_src, int2 *__restrict__ row_inst) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned long long key = (unsigned long long)sorted[i];
    const i32 row = (i32)(key & ((1ull << rowbits) - 1ull));
    const i32 w = (i32)((key >> rowbits) & ((1ull << (fbit - rowbits)) - 1ull));
    const i32 inst = (i32)(i < totalA ? i : i - totalA);
    if (i < totalA) a_src[inst] = row; else r_src[inst] = row;
    row_inst[sorted_vals[i]] = make_int2(w, inst);
}

static int bits_needed(i64 n) {
    int b = 1;
    while ((1ll << b) < n) ++b;
    return b;
}

template <typename KeyT>
static void subset_frames_t(Batch *b, int rowbits, int wbits) {
    Section *sec = b->sec;
    cudaStream_t s = b->stream;
    const i64 W = b->W, nA = sec->nA, nR = sec->nR, n = nA + nR;
    const int fbit = rowbits + wbits;
    DevBuf<i32> wcount;
    wcount.alloc(2 * W + 2, s);
    wcount.zero(s);
    b->row_pos.alloc(n + 1, s);
    const unsigned tiles = blocks_for(n + 1, SUBSET_CT * SUBSET_CI);
    LAUNCH(k_subset_count, tiles, SUBSET_CT, W <= SUBSET_HIST_W ? sizeof(int) * 2 * (size_t)W : 0, s, sec->a_xy.p, nA, sec->r_xy.p, nR, b->d_rects.p, b->rindex, scan_ctx(sec, tiles, 1, s), W, b->row_pos.p,
           wcount.p);
    i32 *h = b->pin_misc();   // [2W] window sizes, [2] totals, then [2(W+1)] offsets built here for the upload
    small_d2h(h, wcount.p, sizeof(i32) * (2 * W + 2), s);
    batch_sync(b);
    const i64 total = h[2 * W], totalA = h[2 * W + 1];
    REQUIRE(total >= 0 && totalA >= 0, SAME_E_LIMIT, "batch exceeds 2^31 window instances");
    b->nAi = totalA; b->nRi = total - totalA;
    i32 *ho = h + 2 * W + 2;
    b->a_off.assign(W + 1, 0); b->r_off.assign(W + 1, 0);
    for (i64 w = 0; w < W; ++w) { b->a_off[w + 1] = b->a_off[w] + h[w]; b->r_off[w + 1] = b->r_off[w] + h[W + w]; }
    for (i64 w = 0; w <= W; ++w) { ho[w] = (i32)b->a_off[w]; ho[W + 1 + w] = (i32)b->r_off[w]; }
    b->d_a_off.alloc(W + 1, s); b->d_r_off.alloc(W + 1, s);
    {
        const SmallCopy up[2] = {{b->d_a_off.p, ho, sizeof(i32) * (size_t)(W + 1)}, {b->d_r_off.p, ho + W + 1, sizeof(i32) * (size_t)(W + 1)}};
        small_h2d_many(b->arena, up, 2, s);
    }
    b->a_src.alloc(b->nAi, s); b->r_src.alloc(b->nRi, s); b->row_inst.alloc(total, s);
    if (total == 0) return;
    DevBuf<KeyT> keys, keys_out;
    DevBuf<i32> vals, vals_out;
    keys.alloc(total, s); keys_out.alloc(total, s); vals.alloc(total, s); vals_out.alloc(total, s);
    LAUNCH((k_subset_fill<KeyT>), blocks_for(n, SUBSET_THREADS), SUBSET_THREADS, 0, s, sec->a_xy.p, nA, sec->r_xy.p, nR, b->d_rects.p, b->rindex, rowbits,
           fbit, b->row_pos.p, keys.p, vals.p);
    size_t bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys_out.p, vals.p, vals_out.p, (int)total, rowbits, fbit + 1, s));
    void *tmp = b->scratch.get(bytes, s);
    {
        ProfScope prof("cub::DeviceRadixSort::SortPairs(subset)", s);
        CK(cub::DeviceRadixSort::SortPairs(tmp, bytes, keys.p, keys_out.p, vals.p, vals_out.p, (int)total, rowbits, fbit + 1, s));
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    LAUNCH((k_subset_finish<KeyT>), blocks_for(total, 256), 256, 0, s, keys_out.p, vals_out.p, total, totalA, rowbits, fbit, b->a_src.p, b->r_src.p,
           b->row_inst.p);
}

static void subset_frames(Batch *b) {
    const int rowbits = bits_needed(std::max<i64>(std::max(b->sec->nA, b->sec->nR), 2)), wbits = bits_needed(b->W + 1);
    if (rowbits + wbits + 1 <= 32) subset_frames_t<unsigned>(b, rowbits, wbits);
    else subset_frames_t<unsigned long long>(b, rowbits, wbits);
}

// host: uniform index grid over the section bbox, each rectangle registered in every grid cell it overlaps (+1 cell of slack)
static void build_rect_index(Batch *b) {
    Section *sec = b->sec;
    cudaStream_t s = b->stream;
    const i64 W = b->W;
    const double bx0 = sec->bbox[0], bx1 = sec->bbox[1], by0 = sec->bbox[2], by1 = sec->bbox[3];
    const double ext = std::max(std::max(bx1 - bx0, by1 - by0), 1e-300);
    // ~16 index cells across a window: a point then tests ~1.6 rectangles on average instead of ~3.5 with 4 cells per window (the
    // rectangle tests were 40 % of k_subset_count's instructions, profiles/r1m)
    int G = (int)std::min<i64>(512, std::max<i64>(1, (i64)std::ceil(std::sqrt((double)W)) * 16));
    std::vector<i32> &ptr = b->h_ri_ptr, &lst = b->h_ri_rects;   // batch members: the async uploads below read them
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
    b->d_rects.alloc(4 * b->W, s);
    {
        const SmallCopy up[3] = {{b->ri_ptr.p, ptr.data(), sizeof(i32) * ptr.size()}, {b->ri_rects.p, lst.data(), sizeof(i32) * lst.size()},
                                 {b->d_rects.p, b->rects.data(), sizeof(double) * 4 * (size_t)b->W}};
        small_h2d_many(b->arena, up, 3, s);   // rectangle index and the rectangles themselves: one launch
    }
    b->rindex = RectIndexDev{bx0, by0, inv, nx, ny, max_len, b->ri_ptr.p, b->ri_rects.p};
}

void batch_subset(Batch *b) {
    Section *sec = b->sec;
    cudaStream_t s = b->stream;
    batch_pin_acquire(b);
    build_rect_index(b);   // (also uploads the rectangles)
    subset_frames(b);
}

// ---- FP64 vector peak (bench.py: the compute-side denominator SURVEY.md §8d asks to measure in the same run) ---------------
__global__ void __launch_bounds__(256) k_fp64_fma(double *__restrict__ out, int iters) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-3 * (double)(threadIdx.x + k);
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = __fma_rn(a[k], b, c);     // eight independent chains per thread
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += a[k];
    out[(i64)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
double measure_fp64_peak(int device) {
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, iters = 8192;
    double *d = nullptr;
    CK(cudaMalloc((void **)&d, sizeof(double) * (size_t)blocks * 256));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, 0));
        k_fp64_fma<<<blocks, 256>>>(d, iters);
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CK(cudaFree(d));
    return 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks / ((double)best * 1e-3) / 1e12;
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
    if (m == 0) { CK(stream_wait(s)); return; }
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
    CK(stream_wait(s));
}

}  // namespace same
