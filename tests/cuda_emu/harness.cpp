// harness.cpp -- TEST ONLY: C entry points over the kernels compiled against the CPU emulation of CUDA
// (tests/cuda_emu/cuda_runtime.h), called from tests/test_kernels_emulated.py through ctypes.  Each function
// replays what the product's driver (bonej2_b200/csrc/api.cu) does around the same launchers.
#include "common.cuh"

#include <vector>

namespace bjt {
alignas(16) uint32_t sm_raw[(128 * 1024) / 4];
namespace disc_detail { alignas(16) uint32_t zf_smem[(224 * 1024) / 4], dk_smem[(128 * 1024) / 4]; }   // paint_disc.cu      // the dynamic shared memory of edt_tile_kernel (budget 100 KB)
}

using namespace bjt;

#include <sys/mman.h>
#include <unistd.h>

// Buffers with an inaccessible page directly behind (mode 0) or in front of (mode 1) them: a kernel that reads
// or writes one element outside an array dies here instead of passing by luck.
namespace {
int g_guard_mode = 0;
struct Guarded {
    char *map = nullptr, *p = nullptr;
    size_t map_bytes = 0;
    explicit Guarded(size_t bytes) {
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        const size_t body = (bytes + page - 1) / page * page + page;          // at least one whole page of payload
        map_bytes = body + 2 * page;
        map = (char *)mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (map == MAP_FAILED) abort();
        mprotect(map, page, PROT_NONE);
        mprotect(map + page + body, page, PROT_NONE);
        if (g_guard_mode == 1) p = map + page;                                 // starts right behind the front guard
        else p = map + page + body - (bytes + 15) / 16 * 16;                   // ends (16-byte aligned) at the back guard
    }
    ~Guarded() { munmap(map, map_bytes); }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};
}

extern "C" {

void emu_set_guard_mode(int mode) { g_guard_mode = mode; }

// api.cu dev_edt: x, y, z passes; *maxval = largest squared distance
int emu_edt(const uint8_t *in, int w, int h, int d, int inverse, int thresh, uint32_t *sq, uint32_t *maxval) {
    const size_t N = (size_t)w * h * d;
    const int n = w > h ? (w > d ? w : d) : (h > d ? h : d);
    const uint32_t n0 = no_result(n);
    Guarded g_in(N), g_gx(N * 2), g_gy(N * 4), g_sq(N * 4);
    memcpy(g_in.p, in, N);
    uint32_t mv = 0;
    int launches = launch_edt_x(g_in.as<uint8_t>(), w, h, d, inverse, thresh, g_gx.as<uint16_t>(), nullptr);
    launches += launch_edt_y(g_gx.as<uint16_t>(), w, h, d, n0, g_gy.as<uint32_t>(), nullptr);
    launches += launch_edt_z(g_gy.as<uint32_t>(), w, h, d, n0, g_sq.as<uint32_t>(), nullptr, &mv, nullptr);
    memcpy(sq, g_sq.p, N * 4);
    if (maxval) *maxval = mv;
    return launches;
}

// api.cu dev_ridge, direct-index thresholds (maxval <= 65536)
int emu_ridge(const uint32_t *sq, int w, int h, int d, uint32_t maxval, uint32_t *ridge, unsigned long long *count) {
    if (maxval == 0 || maxval > 65536u) return -1;
    const size_t N = (size_t)w * h * d;
    std::vector<uint32_t> t1(maxval + 1, 0), t2(maxval + 1, 0), t3(maxval + 1, 0);
    Guarded g_sq(N * 4), g_ridge(N * 4);
    memcpy(g_sq.p, sq, N * 4);
    unsigned long long c = 0;
    int launches = launch_ridge_template(nullptr, maxval, t1.data(), t2.data(), t3.data(), nullptr);
    launches += launch_ridge(g_sq.as<uint32_t>(), w, h, d, t1.data(), t2.data(), t3.data(), g_ridge.as<uint32_t>(), &c, nullptr);
    memcpy(ridge, g_ridge.p, N * 4);
    if (count) *count = c;
    return launches;
}

// api.cu dev_paint, separable painter: variant bit 0 = u64 items, bit 1 = doubled pools.  *flags != 0: overflow.
// cap / fmax: the test hook that caps the fast kernels' capacities (0 = default); *redo = lines sent to the redo kernels
int emu_paint(const uint32_t *ridge, int w, int h, int d, int variant, int xslab_cols, int cap, int fmax, uint32_t *rsq,
              uint32_t *flags, unsigned long long *redo) {
    const size_t N = (size_t)w * h * d;
    paint_sep_set_xslab_cols(xslab_cols);
    paint_sep_set_limits(cap, fmax);
    const size_t wb = paint_sep_workspace_bytes(w, h, d, variant);
    Guarded g_ws(wb), g_ridge(N * 4), g_rsq(N * 4);
    memset(g_ws.p, 0, wb);
    memcpy(g_ridge.p, ridge, N * 4);
    memset(g_rsq.p, 0xAB, N * 4);                       // the painter must write every voxel
    uint32_t f = 0;
    const int launches = launch_paint_sep(g_ridge.as<uint32_t>(), w, h, d, variant, g_ws.p, g_rsq.as<uint32_t>(), &f, nullptr);
    memcpy(rsq, g_rsq.p, N * 4);
    if (flags) *flags = f;
    if (redo) *redo = paint_sep_redo_lines(g_ws.p);
    paint_sep_set_xslab_cols(0);
    paint_sep_set_limits(0, 0);
    return launches;
}

// api.cu dev_finish_range: 2 sqrt, cleanup, mask, NaN / calibration and statistics of planes [z_lo, z_hi).
// original may be null (mask off).  res = sum, sum2, min, max, count of the non-NaN voxels.
int emu_finish(const uint32_t *rsq, const uint8_t *original, int w, int h, int d, int z_lo, int z_hi, int inverse, int calibrate,
               float pixel_width, float *out, double *res) {
    const size_t N = (size_t)w * h * d, M = (size_t)w * h * (z_hi - z_lo);
    const int np = finish_partials(w, h, z_hi - z_lo);
    Guarded g_rsq(N * 4), g_org(N), g_out(M * 4), g_part((size_t)np * sizeof(StatsPartial)), g_res(sizeof(StatsPartial));
    memcpy(g_rsq.p, rsq, N * 4);
    if (original) memcpy(g_org.p, original, N);
    memset(g_out.p, 0xAB, M * 4);
    const int launches = launch_finish(g_rsq.as<uint32_t>(), original ? g_org.as<uint8_t>() : nullptr, w, h, d, z_lo, z_hi, inverse,
                                       calibrate, pixel_width, g_out.as<float>(), g_part.as<StatsPartial>(), np,
                                       g_res.as<StatsPartial>(), nullptr, nullptr);
    memcpy(out, g_out.p, M * 4);
    const StatsPartial r = *g_res.as<StatsPartial>();
    res[0] = r.sum; res[1] = r.sum2; res[2] = r.mn; res[3] = r.mx; res[4] = (double)r.count;
    return launches;
}

// The painter in steps on several z-slabs of one volume, staircases handed from slab to slab in memory: what
// bonej2_b200/sharded.py does over NCCL and tests/test_gpu_sharded.py does on one GPU.  cuts: nslab + 1 plane
// indices; rmax: no ball reaches further (floor(sqrt(max r^2))); variant as above.  Returns overflow flags (0 = fine),
// -1 when a slab has more than SEP_MAX_SOURCES sources on one side.
int emu_paint_slabs(const uint32_t *ridge, int w, int h, int d, const int *cuts, int nslab, int rmax, int kcap, int variant,
                    int xslab_cols, uint32_t *rsq) {
    const size_t wh = (size_t)w * h;
    paint_sep_set_xslab_cols(xslab_cols);
    struct Slab { int z0, dz; Guarded *ws, *ridge, *rsq; uint32_t flags; };
    std::vector<Slab> S;
    for (int i = 0; i < nslab; ++i) {
        Slab sl{cuts[i], cuts[i + 1] - cuts[i], nullptr, nullptr, nullptr, 0u};
        const size_t n = wh * sl.dz;
        sl.ws = new Guarded(paint_sep_workspace_bytes(w, h, sl.dz, variant));
        memset(sl.ws->p, 0, paint_sep_workspace_bytes(w, h, sl.dz, variant));
        sl.ridge = new Guarded(n * 4); sl.rsq = new Guarded(n * 4);
        memcpy(sl.ridge->p, ridge + wh * sl.z0, n * 4);
        memset(sl.rsq->p, 0xAB, n * 4);
        S.push_back(sl);
    }
    for (auto &sl : S) paint_sep_stage1(sl.ridge->as<uint32_t>(), w, h, sl.dz, variant, sl.ws->p, &sl.flags, nullptr);
    auto gap = [&](int p, int q) {          // from slab p's face plane to slab q's first plane; 0 = out of reach
        const int g = q > p ? S[q].z0 - (S[p].z0 + S[p].dz - 1) : S[p].z0 - (S[q].z0 + S[q].dz - 1);
        return g <= rmax ? g : 0;
    };
    int rc = 0;
    const int nxs = paint_sep_nxslabs(w, h, S[0].dz);
    for (int xs = 0; xs < nxs && rc == 0; ++xs) {
        const size_t ncol = (size_t)paint_sep_xslab_width(w, h, S[0].dz, xs) * h;
        for (auto &sl : S) paint_sep_stage2(w, h, sl.dz, variant, sl.ws->p, xs, &sl.flags, nullptr);
        std::vector<std::vector<Guarded *> > cnt(nslab, std::vector<Guarded *>(nslab, nullptr)), items = cnt;
        for (int p = 0; p < nslab; ++p) for (int q = 0; q < nslab; ++q) {
            if (p == q || !gap(p, q)) continue;
            cnt[p][q] = new Guarded(ncol * 4); items[p][q] = new Guarded(ncol * kcap * 8);
            paint_sep_export(w, h, S[p].dz, variant, S[p].ws->p, xs, q > p ? +1 : -1, gap(p, q), rmax + 1, cnt[p][q]->as<uint32_t>(),
                             items[p][q]->as<uint2>(), kcap, &S[p].flags, nullptr);
        }
        for (int q = 0; q < nslab; ++q) {
            SepImports imp{};
            imp.kcap = kcap;
            for (int p = 0; p < q; ++p) if (cnt[p][q]) {
                if (imp.n_lo == SEP_MAX_SOURCES) { rc = -1; break; }
                imp.lo_cnt[imp.n_lo] = cnt[p][q]->as<uint32_t>(); imp.lo_items[imp.n_lo++] = items[p][q]->as<uint2>();
            }
            for (int p = q + 1; p < nslab; ++p) if (cnt[p][q]) {
                if (imp.n_hi == SEP_MAX_SOURCES) { rc = -1; break; }
                imp.hi_cnt[imp.n_hi] = cnt[p][q]->as<uint32_t>(); imp.hi_items[imp.n_hi++] = items[p][q]->as<uint2>();
            }
            if (rc) break;
            paint_sep_stage3(w, h, S[q].dz, variant, S[q].ws->p, xs, &imp, S[q].rsq->as<uint32_t>(), &S[q].flags, nullptr);
        }
        for (auto &row : cnt) for (auto *g : row) delete g;
        for (auto &row : items) for (auto *g : row) delete g;
    }
    for (auto &sl : S) {
        if (rc == 0) { memcpy(rsq + wh * sl.z0, sl.rsq->p, wh * sl.dz * 4); rc |= (int)sl.flags; }
        delete sl.ws; delete sl.ridge; delete sl.rsq;
    }
    paint_sep_set_xslab_cols(0);
    return rc;
}

// api.cu dev_paint_disc (path 4): planes [z_lo, z_hi) of a ridge volume of d planes, in chunks of chunk_planes
// (0 = the plan's choice); items_per_voxel sizes the pool.  *flags != 0: pool overflow.
int emu_paint_disc(const uint32_t *ridge, int w, int h, int d, int z_lo, int z_hi, int chunk_planes, int items_per_voxel,
                   uint32_t *rsq, uint32_t *flags) {
    const size_t N = (size_t)w * h * d, M = (size_t)w * h * (z_hi - z_lo);
    uint32_t max_r = 0;
    for (size_t i = 0; i < N; ++i) max_r = ridge[i] > max_r ? ridge[i] : max_r;
    if (!paint_disc_supports(max_r)) return -1;
    int zc = 0;
    size_t wb = 0;
    paint_disc_plan(w, h, z_hi - z_lo, items_per_voxel, 0, &zc, &wb);
    if (chunk_planes > 0) { zc = chunk_planes; paint_disc_plan(w, h, zc, items_per_voxel, 0, &zc, &wb); zc = chunk_planes; }
    Guarded g_ws(wb), g_ridge(N * 4), g_rsq(M * 4);
    memset(g_ws.p, 0xCD, wb);
    memcpy(g_ridge.p, ridge, N * 4);
    memset(g_rsq.p, 0xAB, M * 4);                       // the painter must write every voxel
    uint32_t f = 0;
    const int launches = launch_paint_disc(g_ridge.as<uint32_t>(), w, h, d, z_lo, z_hi, max_r, zc, items_per_voxel, g_ws.p,
                                           g_rsq.as<uint32_t>(), &f, nullptr, nullptr);
    memcpy(rsq, g_rsq.p, M * 4);
    if (flags) *flags = f;
    return launches;
}

// api.cu dev_paint, scatter fallback (path 2): count, compact the ridge points, scatter-max every ball
int emu_paint_scatter(const uint32_t *ridge, int w, int h, int d, uint32_t *rsq) {
    const size_t N = (size_t)w * h * d;
    Guarded g_ridge(N * 4), g_rsq(N * 4);
    memcpy(g_ridge.p, ridge, N * 4);
    unsigned long long npts = 0;
    int launches = launch_compact_ridge(g_ridge.as<uint32_t>(), w, h, d, nullptr, &npts, nullptr);
    Guarded g_pts((size_t)(npts ? npts : 1) * sizeof(uint4));
    unsigned long long n2 = 0;
    memset(g_rsq.p, 0, N * 4);
    launches += launch_compact_ridge(g_ridge.as<uint32_t>(), w, h, d, g_pts.as<uint4>(), &n2, nullptr);
    launches += launch_paint_scatter(g_pts.as<uint4>(), npts, w, h, d, g_rsq.as<uint32_t>(), nullptr);
    memcpy(rsq, g_rsq.p, N * 4);
    return n2 == npts ? launches : -1;
}

// bjt_dev_phantom_u8: planes [z0, z0 + dz) of the synthetic phantom
int emu_phantom(const int32_t *shapes, int nshapes, int w, int h, int d, int z0, int dz, uint8_t *out) {
    const size_t M = (size_t)w * h * dz;
    Guarded g_out(M);
    memset(g_out.p, 0xAB, M);
    const int launches = launch_phantom(shapes, nshapes, w, h, d, z0, dz, g_out.as<uint8_t>(), nullptr);
    memcpy(out, g_out.p, M);
    return launches;
}

// ---- the device-level entry points bonej2_b200/sharded.py's CudaStages drives (bjt_dev_* in api.cu), on plain host
// pointers, so that class can run unchanged on CPU tensors under gloo (tests/test_sharded_emulated.py) ----
int emu_dev_edt_x(const uint8_t *in, int w, int h, int d, int inverse, int thresh, uint16_t *gx) {
    return launch_edt_x(in, w, h, d, inverse, thresh, gx, nullptr);
}
int emu_dev_edt_y(const uint16_t *gx, int w, int h, int d, int n_sentinel, uint32_t *gy) {
    return launch_edt_y(gx, w, h, d, no_result(n_sentinel), gy, nullptr);
}
int emu_dev_edt_z(const uint32_t *gy, int w, int h, int d, int n_sentinel, uint32_t *sq) {
    uint32_t mv = 0;
    return launch_edt_z(gy, w, h, d, no_result(n_sentinel), sq, nullptr, &mv, nullptr);
}
int emu_dev_ridge(const uint32_t *sq, int w, int h, int d, uint32_t *ridge) {          // api.cu dev_ridge_full
    uint32_t maxval = 0;
    const size_t N = (size_t)w * h * d;
    for (size_t i = 0; i < N; ++i) maxval = sq[i] > maxval ? sq[i] : maxval;
    if (maxval == 0) { memset(ridge, 0, N * 4); return 0; }
    return emu_ridge(sq, w, h, d, maxval, ridge, nullptr);
}
int emu_dev_paint(const uint32_t *ridge, int w, int h, int d, uint32_t *rsq) {           // whole slab + halo at once
    uint32_t flags = 0;
    const int rc = emu_paint(ridge, w, h, d, 0, 0, 0, 0, rsq, &flags, nullptr);
    return flags ? -1 : rc;
}

struct EmuPaintSession { int w, h, d, variant; void *ws; uint32_t flags; };

void *emu_paint_open(const uint32_t *ridge, int w, int h, int d, long long max_rsq) {     // bjt_dev_paint_open
    EmuPaintSession *S = new EmuPaintSession{w, h, d, max_rsq >= 65536 ? 1 : 0, nullptr, 0u};
    S->ws = calloc(1, paint_sep_workspace_bytes(w, h, d, S->variant));
    paint_sep_stage1(ridge, w, h, d, S->variant, S->ws, &S->flags, nullptr);
    return S;
}
int emu_paint_nxslabs(void *h) { EmuPaintSession *S = (EmuPaintSession *)h; return paint_sep_nxslabs(S->w, S->h, S->d); }
int emu_paint_xslab_width(void *h, int xs) { EmuPaintSession *S = (EmuPaintSession *)h; return paint_sep_xslab_width(S->w, S->h, S->d, xs); }
int emu_paint_lists(void *h, int xs) {
    EmuPaintSession *S = (EmuPaintSession *)h;
    return paint_sep_stage2(S->w, S->h, S->d, S->variant, S->ws, xs, &S->flags, nullptr);
}
int emu_paint_export(void *h, int xs, int side, int gap, int nplanes, uint32_t *cnt, void *items, int kcap) {
    EmuPaintSession *S = (EmuPaintSession *)h;
    return paint_sep_export(S->w, S->h, S->d, S->variant, S->ws, xs, side, gap, nplanes, cnt, (uint2 *)items, kcap, &S->flags, nullptr);
}
int emu_paint_sweep(void *h, int xs, int n_lo, const uint32_t *const *lo_cnt, const void *const *lo_items, int n_hi,
                    const uint32_t *const *hi_cnt, const void *const *hi_items, int kcap, uint32_t *rsq) {
    EmuPaintSession *S = (EmuPaintSession *)h;
    if (n_lo > SEP_MAX_SOURCES || n_hi > SEP_MAX_SOURCES) return -1;
    SepImports imp{};
    imp.n_lo = n_lo; imp.n_hi = n_hi; imp.kcap = kcap;
    for (int i = 0; i < n_lo; ++i) { imp.lo_cnt[i] = lo_cnt[i]; imp.lo_items[i] = (const uint2 *)lo_items[i]; }
    for (int i = 0; i < n_hi; ++i) { imp.hi_cnt[i] = hi_cnt[i]; imp.hi_items[i] = (const uint2 *)hi_items[i]; }
    return paint_sep_stage3(S->w, S->h, S->d, S->variant, S->ws, xs, &imp, rsq, &S->flags, nullptr);
}
unsigned emu_paint_close(void *h) {
    EmuPaintSession *S = (EmuPaintSession *)h;
    const unsigned f = S->flags;
    free(S->ws);
    delete S;
    return f;
}
int emu_paint_block_cols(int w, int h, int d) { (void)w; return paint_sep_natural_xslab_cols(h, d); }
void emu_paint_set_block_cols(int cols) { paint_sep_set_xslab_cols(cols); }

// bjt_dev_finish_range on host pointers; `original` may be null and may point in front of an array (the caller shifts
// it by z_lo planes, as CudaStages.finish does): no guarded copies here
int emu_dev_finish_range(const uint32_t *rsq, const uint8_t *original, int w, int h, int d, int z_lo, int z_hi, int inverse,
                         int calibrate, float pixel_width, float *out, double *res) {
    const int np = finish_partials(w, h, z_hi - z_lo);
    std::vector<StatsPartial> part((size_t)np);
    StatsPartial r;
    const int launches = launch_finish(rsq, original, w, h, d, z_lo, z_hi, inverse, calibrate, pixel_width, out, part.data(), np, &r, nullptr, nullptr);
    res[0] = r.sum; res[1] = r.sum2; res[2] = r.mn; res[3] = r.mx; res[4] = (double)r.count;
    return launches;
}

}  // extern "C"
