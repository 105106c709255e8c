// api.cu -- the C-ABI of libbonejthickness.so (include/bonej_thickness.h) and the pipeline driver.
//
// One call of bjt_thickness_u8 is one LocalThicknessWrapper.processImage(ImagePlus) plus the
// StackStatistics that ThicknessWrapper.addMapResults reads (ThicknessWrapper.java:158-196):
//   upload -> EDT x,y,z -> ridge -> paint -> 2*sqrt/cleanup/mask/NaN/calibrate/stats -> download.
// All device memory belongs to the context and is cached between calls; nothing here keeps a
// caller pointer after returning.  There is no CPU path: every failure is reported, never routed.
#include "common.cuh"
#include "../../include/bonej_thickness.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <map>
#include <vector>
#include <mutex>
#include <string>
#include <thread>

using namespace bjt;

struct bjt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // downloads that overlap the next run
    bool own_stream = false;
    char err[512] = {0};
    std::mutex mu;
    struct Buf { void *p = nullptr; size_t cap = 0; };
    std::map<std::string, Buf> bufs;
    uint32_t *h_scal = nullptr;        // pinned readback area (64 words)
    StatsPartial *h_stats = nullptr;   // pinned
    cudaEvent_t ev[16];
    int paint_path = -1;
    int64_t launches = 0;
    int sm_count = 0;
    // per-stage device time of the last separable paint (events around the x stage and every column block's y / z stage)
    std::vector<cudaEvent_t> pev;
    std::vector<cudaEvent_t> sev;        // events of the streamed paint / finish / download path
    float paint_ms[3] = {0, 0, 0};
    int paint_launches[3] = {0, 0, 0};
    // stepwise painter of a z-slab (bjt_dev_paint_open .. _close)
    struct { bool open = false; int w = 0, h = 0, d = 0, variant = 0; void *ws = nullptr; uint32_t *flags = nullptr; } ps;
    // bjt_create_multi: one single-device context per z-slab (empty for a plain context); the fields below are the
    // per-slab state of a run, written by the slab's worker and read by the others only after a phase has been joined
    std::vector<bjt_ctx *> kids;
    struct { uint32_t maxval = 0, not_binary = 0; StatsPartial part{}; uint32_t *gy = nullptr, *sqt = nullptr; } mg;
};

static char g_err[512] = "no error";

static int fail(bjt_ctx *c, int code, const char *fmt, ...) {
    char *dst = c ? c->err : g_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            cudaGetLastError();   /* reported here: it must not surface again at the next launch check */  \
            return fail(c, BJT_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                                \
    } while (0)

#define CKL()                                                                                            \
    do {                                                                                                 \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess)                                                                           \
            return fail(c, BJT_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// A few words from device memory into the page-locked readback area, written by a kernel through the mapped
// pointer instead of a device-to-host memcpy: a memcpy would queue behind whatever the copy engine is doing,
// and in a "Both" call that is the 4 N-byte download of the first map -- the second run would wait for it at its
// first readback.
__global__ void peek_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int nwords) {
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
}
static cudaError_t peek(bjt_ctx *c, void *host_dst, const void *dev_src, size_t bytes) {
    peek_kernel<<<1, 32, 0, c->stream>>>(reinterpret_cast<const uint32_t *>(dev_src), reinterpret_cast<uint32_t *>(host_dst),
                                         (int)(bytes / 4));
    return cudaGetLastError();
}

static int ws_get(bjt_ctx *c, const char *name, size_t bytes, void **out) {
    bjt_ctx::Buf &b = c->bufs[name];
    if (b.cap < bytes) {
        if (b.p) { CK(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
        if (bytes == 0) bytes = 256;
        CK(cudaMalloc(&b.p, bytes));
        b.cap = bytes;
    }
    *out = b.p;
    return BJT_OK;
}
#define WS(name, bytes, ptr)                                             \
    do {                                                                 \
        int rc_ = ws_get(c, name, (bytes), (void **)&(ptr));             \
        if (rc_) return rc_;                                             \
    } while (0)

static int check_dims(bjt_ctx *c, int w, int h, int d) {
    if (!c) return fail(nullptr, BJT_E_ARG, "null context");
    if (w <= 0 || h <= 0 || d <= 0) return fail(c, BJT_E_ARG, "non-positive volume size %d x %d x %d", w, h, d);
    if (w > MAX_EXTENT || h > MAX_EXTENT || d > MAX_EXTENT)
        return fail(c, BJT_E_ARG, "volume extent above %d is not supported", MAX_EXTENT);
    return BJT_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define ENTER(c)                                            \
    if (!(c)) return fail(nullptr, BJT_E_ARG, "null context"); \
    std::lock_guard<std::mutex> lk_((c)->mu);               \
    DeviceGuard dg_((c)->device)

extern "C" {

int bjt_create(bjt_ctx **out, int device) {
    if (!out) return fail(nullptr, BJT_E_ARG, "null output pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, BJT_E_CUDA, "no CUDA device: %s (libbonejthickness has no CPU fallback)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= ndev) return fail(nullptr, BJT_E_ARG, "device %d out of range (%d devices)", device, ndev);
    bjt_ctx *c = new bjt_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(&c->h_scal, 64 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMallocHost(&c->h_stats, sizeof(StatsPartial)) != cudaSuccess) {
        fail(nullptr, BJT_E_CUDA, "context setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return BJT_E_CUDA;
    }
    c->own_stream = true;
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) c->copy_stream = nullptr;
    for (auto &ev : c->ev) cudaEventCreate(&ev);
    *out = c;
    return BJT_OK;
}

int bjt_create_multi(bjt_ctx **out, const int *devices, int ndev) {
    if (!out) return fail(nullptr, BJT_E_ARG, "null output pointer");
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > 64) return fail(nullptr, BJT_E_ARG, "bjt_create_multi: between 1 and 64 devices");
    int rc = bjt_create(out, devices[0]);
    if (rc) return rc;
    bjt_ctx *c = *out;
    if (ndev == 1) return BJT_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    for (int i = 0; i < ndev; ++i) {
        bjt_ctx *k = nullptr;
        rc = bjt_create(&k, devices[i]);
        if (rc) { bjt_destroy(c); *out = nullptr; return rc; }
        c->kids.push_back(k);
    }
    // slabs read each other's memory (strided copies over NVLink): map every peer that can be mapped; the copies
    // still work, staged by the driver, where it cannot
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < ndev; ++j) {
            const int a = c->kids[i]->device, b = c->kids[j]->device;
            int can = 0;
            if (a == b || cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) continue;
            cudaSetDevice(a);
            cudaDeviceEnablePeerAccess(b, 0);
            cudaGetLastError();                                        // "already enabled" is fine
        }
    if (prev >= 0) cudaSetDevice(prev);
    return BJT_OK;
}

int bjt_device_count(const bjt_ctx *c) { return c ? (c->kids.empty() ? 1 : (int)c->kids.size()) : 0; }

int bjt_release_workspace(bjt_ctx *c) {
    if (!c) return BJT_E_ARG;
    for (bjt_ctx *k : c->kids) bjt_release_workspace(k);
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard dg(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &kv : c->bufs) if (kv.second.p) cudaFree(kv.second.p);
    c->bufs.clear();
    return BJT_OK;
}

void bjt_destroy(bjt_ctx *c) {
    if (!c) return;
    for (bjt_ctx *k : c->kids) bjt_destroy(k);
    c->kids.clear();
    bjt_release_workspace(c);
    DeviceGuard dg(c->device);
    for (auto &ev : c->ev) cudaEventDestroy(ev);
    for (auto &ev : c->pev) cudaEventDestroy(ev);
    for (auto &ev : c->sev) cudaEventDestroy(ev);
    if (c->h_scal) cudaFreeHost(c->h_scal);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

const char *bjt_last_error(const bjt_ctx *c) { return c ? c->err : g_err; }

int bjt_set_stream(bjt_ctx *c, void *stream) {
    if (!c) return BJT_E_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = (cudaStream_t)stream;
    c->own_stream = false;
    return BJT_OK;
}

int bjt_set_paint_path(bjt_ctx *c, int path) {
    if (!c || path < -1 || path > 4 || path == 3) return BJT_E_ARG;
    c->paint_path = path;
    for (bjt_ctx *k : c->kids) k->paint_path = path;
    return BJT_OK;
}

int bjt_device_info(bjt_ctx *c, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor) {
    if (!c) return BJT_E_ARG;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, c->device));
    if (name && name_len > 0) { strncpy(name, p.name, (size_t)name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return BJT_OK;
}

int64_t bjt_launch_count(const bjt_ctx *c) { return c ? c->launches : 0; }

int bjt_debug_set_paint_limits(bjt_ctx *c, int active_cap, int front_cap) {
    if (!c) return BJT_E_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    paint_sep_set_limits(active_cap, front_cap);
    return BJT_OK;
}

int bjt_paint_block_cols(int w, int h, int d) {
    if (w <= 0 || h <= 0 || d <= 0) return BJT_E_ARG;
    return paint_sep_natural_xslab_cols(h, d);
}

int bjt_paint_set_block_cols(int cols) {
    if (cols < 0 || (cols % 32) != 0) return BJT_E_ARG;
    paint_sep_set_xslab_cols(cols);
    return BJT_OK;
}

int bjt_debug_set_xslab_cols(int cols) { return bjt_paint_set_block_cols(cols); }

int bjt_debug_isqrt_check(bjt_ctx *c, int64_t *mismatches) {
    if (!c || !mismatches) return BJT_E_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    uint32_t *scal;
    WS("scalars", 256, scal);
    CK(cudaMemsetAsync(scal + 16, 0, 4, c->stream));
    c->launches += launch_isqrt_small_check(scal + 16, c->stream); CKL();
    CK(peek(c, c->h_scal + 16, scal + 16, 4));
    CK(cudaStreamSynchronize(c->stream));
    *mismatches = (int64_t)c->h_scal[16];
    return BJT_OK;
}

int64_t bjt_debug_paint_redo_lines(bjt_ctx *c) {
    if (!c) return -1;
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->bufs.find("paint_sep");
    if (it == c->bufs.end() || !it->second.p) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    return (int64_t)paint_sep_redo_lines(it->second.p);
}

void *bjt_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void bjt_host_free(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// stages on device buffers (no locking, no sync unless a value is needed on the host)

static void stats_out(const StatsPartial &r, bjt_stats *st) {
    if (!st) return;
    st->count = r.count; st->sum = r.sum; st->sum2 = r.sum2; st->min = r.mn; st->max = r.mx;
    const double n = (double)r.count;
    st->mean = r.sum / n;                 // 0/0 = NaN when the map is empty, as in the Java code
    double sd = 0.0;
    if (r.count > 0) {
        sd = (n * r.sum2 - r.sum * r.sum) / n;
        sd = sd > 0.0 ? sqrt(sd / (n - 1.0)) : 0.0;
    }
    st->sd = sd;
}

static int dev_edt(bjt_ctx *c, const uint8_t *d_in, int w, int h, int d, int inverse, int thresh, int n_sentinel,
                   uint16_t *gx, uint32_t *gy, uint32_t *sq, uint32_t *maxval_host, cudaEvent_t *evs) {
    const uint32_t n0 = no_result(n_sentinel);
    uint32_t *scal;
    WS("scalars", 256, scal);
    CK(cudaMemsetAsync(scal, 0, 256, c->stream));
    if (evs) CK(cudaEventRecord(evs[0], c->stream));
    c->launches += launch_edt_x(d_in, w, h, d, inverse, thresh, gx, c->stream); CKL();
    if (evs) CK(cudaEventRecord(evs[1], c->stream));
    c->launches += launch_edt_y(gx, w, h, d, n0, gy, c->stream); CKL();
    if (evs) CK(cudaEventRecord(evs[2], c->stream));
    c->launches += launch_edt_z(gy, w, h, d, n0, sq, nullptr, scal, c->stream); CKL();
    if (evs) CK(cudaEventRecord(evs[3], c->stream));
    if (maxval_host) {
        CK(peek(c, c->h_scal, scal, sizeof(uint32_t)));
        CK(cudaStreamSynchronize(c->stream));
        *maxval_host = c->h_scal[0];
    }
    return BJT_OK;
}

// thresholds -> ridge.  maxval: largest value of sq.  Up to 65536 the thresholds are simply computed for
// every r^2 in 1..maxval (values that do not occur are never looked up); beyond that only for the values
// that occur, found with a presence table.
static int dev_ridge(bjt_ctx *c, const uint32_t *sq, int w, int h, int d, uint32_t maxval, uint32_t *ridge,
                     int64_t *n_ridge) {
    const size_t N = (size_t)w * h * d;
    uint32_t *values, *t1, *t2, *t3, *scal;
    const size_t tb = ((size_t)maxval + 1) * sizeof(uint32_t);
    WS("t1", tb, t1); WS("t2", tb, t2); WS("t3", tb, t3);
    WS("scalars", 256, scal);
    CK(cudaMemsetAsync(scal + 2, 0, 16, c->stream));   // [2] = nvalues, [4..5] = ridge count (u64)
    if (maxval <= 65536u) {
        c->launches += launch_ridge_template(nullptr, maxval, t1, t2, t3, c->stream); CKL();
    } else {
        uint8_t *present;
        WS("values", tb, values);
        WS("present", (size_t)maxval + 1, present);
        CK(cudaMemsetAsync(present, 0, (size_t)maxval + 1, c->stream));
        c->launches += launch_mark_present(sq, N, present, scal + 6, c->stream); CKL();
        c->launches += launch_collect_values(present, maxval, values, scal + 2, c->stream); CKL();
        CK(peek(c, c->h_scal + 2, scal + 2, sizeof(uint32_t)));
        CK(cudaStreamSynchronize(c->stream));
        c->launches += launch_ridge_template(values, c->h_scal[2], t1, t2, t3, c->stream); CKL();
    }
    c->launches += launch_ridge(sq, w, h, d, t1, t2, t3, ridge, (unsigned long long *)(scal + 4), c->stream); CKL();
    if (n_ridge) {
        CK(peek(c, c->h_scal + 4, scal + 4, 8));
        CK(cudaStreamSynchronize(c->stream));
        *n_ridge = (int64_t)(*(unsigned long long *)(c->h_scal + 4));
    }
    return BJT_OK;
}

// ridge of an arbitrary squared map on the device: largest value first, then as above
static int dev_ridge_full(bjt_ctx *c, const uint32_t *sq, int w, int h, int d, uint32_t *ridge, int64_t *n_ridge) {
    const size_t N = (size_t)w * h * d;
    uint32_t *scal;
    WS("scalars", 256, scal);
    CK(cudaMemsetAsync(scal, 0, 4, c->stream));
    c->launches += launch_mark_present(sq, N, nullptr, scal, c->stream); CKL();
    CK(peek(c, c->h_scal, scal, 4));
    CK(cudaStreamSynchronize(c->stream));
    return dev_ridge(c, sq, w, h, d, c->h_scal[0], ridge, n_ridge);
}

// Fronts along z, then discs in the plane (paint_disc.cu), over planes [z_lo, z_hi) of a ridge volume of dext
// planes; rsq points at plane z_lo.  *done = 0 when the item pool could not be made large enough (nothing valid
// in rsq; the caller takes another path).
static int dev_paint_disc(bjt_ctx *c, const uint32_t *ridge, int w, int h, int dext, int z_lo, int z_hi, uint32_t max_r,
                          uint32_t *rsq, int *done) {
    uint32_t *scal;
    WS("scalars", 256, scal);
    *done = 0;
    for (int ipv = 3; ipv <= 48; ipv *= 4) {
        size_t have = 0;
        { auto it = c->bufs.find("paint_disc"); if (it != c->bufs.end()) have = it->second.cap; }
        int zc = 0;
        size_t wb = 0;
        paint_disc_plan(w, h, z_hi - z_lo, ipv, (size_t)6 << 30, &zc, &wb);
        if (wb > have) {
            // the workspace has to grow: only now ask the driver how much room there is (cudaMemGetInfo takes
            // milliseconds on some hosts -- measured 11 ms per call on a two-GPU box -- and a steady-state call with a
            // cached workspace must not pay it)
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            const size_t avail = have + (free_b > ((size_t)1 << 30) ? free_b - ((size_t)1 << 30) : 0);
            paint_disc_plan(w, h, z_hi - z_lo, ipv, std::min(avail, (size_t)6 << 30), &zc, &wb);
            if (wb > avail) return BJT_OK;                     // does not fit: not done
        }
        void *wsp;
        WS("paint_disc", wb, wsp);
        const int nchunks = (z_hi - z_lo + zc - 1) / zc;
        while (c->pev.size() < (size_t)(2 * nchunks + 1)) { cudaEvent_t e; CK(cudaEventCreate(&e)); c->pev.push_back(e); }
        CK(cudaMemsetAsync(scal + 8, 0, 4, c->stream));
        c->launches += launch_paint_disc(ridge, w, h, dext, z_lo, z_hi, max_r, zc, ipv, wsp, rsq, scal + 8, c->pev.data(),
                                         c->stream); CKL();
        CK(peek(c, c->h_scal + 8, scal + 8, 4));
        CK(cudaStreamSynchronize(c->stream));
        c->paint_ms[0] = c->paint_ms[1] = c->paint_ms[2] = 0.f;
        for (int k = 0; k < nchunks; ++k) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, c->pev[2 * k], c->pev[2 * k + 1]);
            cudaEventElapsedTime(&b, c->pev[2 * k + 1], c->pev[2 * k + 2]);
            c->paint_ms[0] += a; c->paint_ms[1] += b;
        }
        c->paint_launches[0] = nchunks; c->paint_launches[1] = nchunks; c->paint_launches[2] = 0;
        if (c->h_scal[8] == 0) { *done = 1; return BJT_OK; }
    }
    return BJT_OK;
}

static int dev_paint(bjt_ctx *c, const uint32_t *ridge, int w, int h, int d, int path, int64_t n_ridge_known,
                     int64_t max_r, uint32_t *rsq, int *path_used) {
    const size_t N = (size_t)w * h * d;
    if (path < 0) path = c->paint_path;
    const bool forced = path >= 0;
    if (path <= 1 || path == 4) {
        uint32_t *scal;
        WS("scalars", 256, scal);
        if (max_r < 0) {   // largest tag: decides the painter and the item width
            CK(cudaMemsetAsync(scal + 14, 0, 4, c->stream));
            c->launches += launch_mark_present(ridge, N, nullptr, scal + 14, c->stream); CKL();
            CK(peek(c, c->h_scal + 14, scal + 14, 4));
            CK(cudaStreamSynchronize(c->stream));
            max_r = c->h_scal[14];
        }
    }
    if (path < 0 || path == 4) {
        if (paint_disc_supports((uint32_t)max_r)) {
            int done = 0;
            const int rc = dev_paint_disc(c, ridge, w, h, d, 0, d, (uint32_t)max_r, rsq, &done);
            if (rc) return rc;
            if (done) { if (path_used) *path_used = 4; return BJT_OK; }
        }
        if (forced) return fail(c, BJT_E_ARG, "paint path 4 cannot take this volume (largest r^2 %lld, or no room for its items)", (long long)max_r);
    }
    if (path <= 1) {
        uint32_t *scal;
        WS("scalars", 256, scal);
        int wide = max_r >= 65536 ? 1 : 0;
        if (forced && path == 0 && wide) return fail(c, BJT_E_ARG, "paint path 0 needs r^2 < 65536 (largest is %lld)", (long long)max_r);
        if (forced && path == 1) wide = 1;
        // small capacities first, then the large variant; list overflow is detected, never silent
        for (int big = 0; big <= 1; ++big) {
            const int variant = wide | (big << 1);
            const size_t wb = paint_sep_workspace_bytes(w, h, d, variant);
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            size_t have = 0;
            { auto it = c->bufs.find("paint_sep"); if (it != c->bufs.end()) have = it->second.cap; }
            if (wb > have && wb - have + ((size_t)1 << 30) > free_b) continue;    // does not fit: next option
            void *wsp;
            WS("paint_sep", wb, wsp);
            CK(cudaMemsetAsync(scal + 8, 0, 4, c->stream));
            // x stage, then per column block the y and z stages, with an event at every boundary
            const int nxs = paint_sep_nxslabs(w, h, d);
            const size_t nev = 1 + (size_t)nxs * 2 + 1;
            while (c->pev.size() < nev) { cudaEvent_t e; CK(cudaEventCreate(&e)); c->pev.push_back(e); }
            size_t ne = 0;
            CK(cudaEventRecord(c->pev[ne++], c->stream));
            c->launches += paint_sep_stage1(ridge, w, h, d, variant, wsp, scal + 8, c->stream); CKL();
            CK(cudaEventRecord(c->pev[ne++], c->stream));
            for (int xs = 0; xs < nxs; ++xs) {
                c->launches += paint_sep_stage2(w, h, d, variant, wsp, xs, scal + 8, c->stream); CKL();
                CK(cudaEventRecord(c->pev[ne++], c->stream));
                c->launches += paint_sep_stage3(w, h, d, variant, wsp, xs, nullptr, rsq, scal + 8, c->stream); CKL();
                CK(cudaEventRecord(c->pev[ne++], c->stream));
            }
            CK(peek(c, c->h_scal + 8, scal + 8, 4));
            CK(cudaStreamSynchronize(c->stream));
            c->paint_ms[0] = c->paint_ms[1] = c->paint_ms[2] = 0.f;
            cudaEventElapsedTime(&c->paint_ms[0], c->pev[0], c->pev[1]);
            for (int xs = 0; xs < nxs; ++xs) {
                float a = 0.f, b = 0.f;
                cudaEventElapsedTime(&a, c->pev[1 + 2 * xs], c->pev[2 + 2 * xs]);
                cudaEventElapsedTime(&b, c->pev[2 + 2 * xs], c->pev[3 + 2 * xs]);
                c->paint_ms[1] += a; c->paint_ms[2] += b;
            }
            c->paint_launches[0] = 2; c->paint_launches[1] = 2 * nxs; c->paint_launches[2] = 2 * nxs;   // sweeps (each followed by its redo kernel)
            if (c->h_scal[8] == 0) { if (path_used) *path_used = wide; return BJT_OK; }
        }
        if (forced)
            return fail(c, BJT_E_INTERNAL, "separable painter (path %d) overflowed its lists (flags %u)", path, c->h_scal[8]);
    }
    // scatter
    {
        uint32_t *scal; uint4 *points;
        WS("scalars", 256, scal);
        unsigned long long npts = 0;
        if (n_ridge_known >= 0) {
            npts = (unsigned long long)n_ridge_known;
        } else {   // count first (null sink)
            CK(cudaMemsetAsync(scal + 10, 0, 8, c->stream));
            c->launches += launch_compact_ridge(ridge, w, h, d, nullptr, (unsigned long long *)(scal + 10), c->stream); CKL();
            CK(peek(c, c->h_scal + 10, scal + 10, 8));
            CK(cudaStreamSynchronize(c->stream));
            npts = *(unsigned long long *)(c->h_scal + 10);
        }
        WS("points", (size_t)(npts > 0 ? npts : 1) * sizeof(uint4), points);
        CK(cudaMemsetAsync(scal + 10, 0, 8, c->stream));
        CK(cudaMemsetAsync(rsq, 0, N * sizeof(uint32_t), c->stream));
        c->launches += launch_compact_ridge(ridge, w, h, d, points, (unsigned long long *)(scal + 10), c->stream); CKL();
        c->launches += launch_paint_scatter(points, npts, w, h, d, rsq, c->stream); CKL();
        if (path_used) *path_used = 2;
    }
    return BJT_OK;
}

// planes [z_lo, z_hi) of the painted map of a ridge volume of d planes (bjt_dev_paint_range)
static int dev_paint_range(bjt_ctx *c, const uint32_t *d_ridge, int w, int h, int d, int z_lo, int z_hi, int64_t max_rsq,
                           uint32_t *d_rsq) {
    if (z_lo < 0 || z_hi > d || z_lo >= z_hi) return fail(c, BJT_E_ARG, "paint_range: bad plane range [%d, %d) of %d", z_lo, z_hi, d);
    const size_t N = (size_t)w * h * d, wh = (size_t)w * h;
    uint32_t *scal;
    WS("scalars", 256, scal);
    if (max_rsq < 0) {
        CK(cudaMemsetAsync(scal + 14, 0, 4, c->stream));
        c->launches += launch_mark_present(d_ridge, N, nullptr, scal + 14, c->stream); CKL();
        CK(peek(c, c->h_scal + 14, scal + 14, 4));
        CK(cudaStreamSynchronize(c->stream));
        max_rsq = c->h_scal[14];
    }
    if ((c->paint_path < 0 || c->paint_path == 4) && paint_disc_supports((uint32_t)max_rsq)) {
        int done = 0;
        const int rc = dev_paint_disc(c, d_ridge, w, h, d, z_lo, z_hi, (uint32_t)max_rsq, d_rsq, &done);
        if (rc) return rc;
        if (done) return BJT_OK;
    }
    // radii beyond the disc painter (or a forced path): paint every plane of the slab, keep the requested ones
    uint32_t *tmp;
    WS("paint_range_tmp", N * sizeof(uint32_t), tmp);
    const int rc = dev_paint(c, d_ridge, w, h, d, c->paint_path == 4 ? -1 : c->paint_path, -1, max_rsq, tmp, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(d_rsq, tmp + (size_t)z_lo * wh, (size_t)(z_hi - z_lo) * wh * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return BJT_OK;
}

static int dev_finish_range(bjt_ctx *c, const uint32_t *rsq, const uint8_t *d_original, int w, int h, int d, int z_lo,
                            int z_hi, int inverse, int calibrate, double pixel_width, float *d_out, bjt_stats *stats) {
    if (z_lo < 0 || z_hi > d || z_lo >= z_hi) return fail(c, BJT_E_ARG, "bad plane range [%d, %d) of %d", z_lo, z_hi, d);
    StatsPartial *partials, *result;
    const int np = finish_partials(w, h, z_hi - z_lo);
    WS("partials", (size_t)np * sizeof(StatsPartial), partials);
    WS("stats_result", sizeof(StatsPartial), result);
    float *ttab;
    {   // the 2*sqrt table lives as long as the context's workspace and is filled once
        const bool fresh = c->bufs.find("thick_tab") == c->bufs.end();
        WS("thick_tab", thickness_table_bytes(), ttab);
        if (fresh) { c->launches += launch_thickness_table(ttab, c->stream); CKL(); }
    }
    c->launches += launch_finish(rsq, d_original, w, h, d, z_lo, z_hi, inverse, calibrate, (float)pixel_width, d_out,
                                 partials, np, result, ttab, c->stream); CKL();
    if (stats) {
        CK(peek(c, c->h_stats, result, sizeof(StatsPartial)));
        CK(cudaStreamSynchronize(c->stream));
        stats_out(*c->h_stats, stats);
    }
    return BJT_OK;
}

static int dev_finish(bjt_ctx *c, const uint32_t *rsq, const uint8_t *d_original, int w, int h, int d, int inverse,
                      int calibrate, double pixel_width, float *d_out, bjt_stats *stats) {
    return dev_finish_range(c, rsq, d_original, w, h, d, 0, d, inverse, calibrate, pixel_width, d_out, stats);
}

// Where a host entry point wants its map: with it, the pipeline paints and finishes the volume in chunks of planes and
// sends every finished chunk home on the copy stream while the next chunk is painted -- the download of a 1024^3 map
// (4 GiB over PCIe, ~75 ms) then hides behind the painter instead of following it.
struct StreamOut {
    float *host = nullptr;                  // contiguous map, or
    float *const *host_slices = nullptr;    // one array per slice
    cudaStream_t copy = nullptr;
    cudaEvent_t first = nullptr, last = nullptr;   // recorded on the copy stream around the copies (timing)
    bool used = false;                      // set when the map went out this way
};

// paint (fronts + discs) and finish in chunks of planes, each finished chunk copied out asynchronously.
// *done = 0: the disc painter could not take a chunk (nothing usable in d_out; the caller runs the plain path).
static int dev_paint_finish_streamed(bjt_ctx *c, const uint32_t *ridge, const uint8_t *d_orig, int w, int h, int d, int inverse,
                                     double pixel_width, uint32_t maxval, uint32_t *rsq, float *d_out, StreamOut *so, bjt_stats *stats,
                                     float *ms_paint, float *ms_finish, int *done) {
    const size_t wh = (size_t)w * h;
    *done = 0;
    int zc = 0;
    size_t wb = 0;
    paint_disc_plan(w, h, d, 3, (size_t)6 << 30, &zc, &wb);
    const int nchunks = (d + zc - 1) / zc + 8;           // + the halved chunks at the end
    while (c->sev.size() < (size_t)(2 * nchunks + 2)) { cudaEvent_t e; CK(cudaEventCreate(&e)); c->sev.push_back(e); }
    StatsPartial tot{0.0, 0.0, 0.0, 0.0, 0};
    bool any = false;
    float px = 0.f, py = 0.f, pf = 0.f;
    int nlaunch_disc = 0, finished = 0, nev = 0;
    bool first_copy = true;
    auto finish_to = [&](int upto) -> int {              // finish and send planes [finished, upto)
        if (upto <= finished) return BJT_OK;
        bjt_stats st;
        cudaEvent_t e0 = c->sev[nev++], e1 = c->sev[nev++];
        CK(cudaEventRecord(e0, c->stream));
        const int rc = dev_finish_range(c, rsq, d_orig, w, h, d, finished, upto, inverse, 1, pixel_width, d_out + (size_t)finished * wh, &st);
        if (rc) return rc;                               // (returns with the compute stream idle)
        CK(cudaEventRecord(e1, c->stream));
        if (st.count > 0) {
            tot.sum += st.sum; tot.sum2 += st.sum2; tot.count += st.count;
            tot.mn = any ? std::min(tot.mn, st.min) : st.min;
            tot.mx = any ? std::max(tot.mx, st.max) : st.max;
            any = true;
        }
        CK(cudaStreamWaitEvent(so->copy, e1, 0));
        if (first_copy && so->first) { CK(cudaEventRecord(so->first, so->copy)); first_copy = false; }
        if (so->host) {
            CK(cudaMemcpyAsync(so->host + (size_t)finished * wh, d_out + (size_t)finished * wh, (size_t)(upto - finished) * wh * sizeof(float),
                               cudaMemcpyDeviceToHost, so->copy));
        } else {
            for (int z = finished; z < upto; ++z)
                CK(cudaMemcpyAsync(so->host_slices[z], d_out + (size_t)z * wh, wh * sizeof(float), cudaMemcpyDeviceToHost, so->copy));
        }
        float ms = 0.f;
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        pf += ms;
        finished = upto;
        return BJT_OK;
    };
    // Chunks of zc planes, halving towards the end (.., zc, zc/2, .., 32, 32): what is left when the last chunk has been
    // painted is its finish pass and its download, and the download of a full chunk (0.5 GiB at 1024^3) is 10 ms.
    for (int z0 = 0, z1; z0 < d; z0 = z1) {
        const int left = d - z0;
        int step = zc;
        if (left < 2 * zc) step = std::max(32, (left / 2 + 31) / 32 * 32);
        z1 = std::min(d, z0 + std::min(step, zc));
        int ok = 0;
        int rc = dev_paint_disc(c, ridge, w, h, d, z0, z1, maxval, rsq + (size_t)z0 * wh, &ok);
        if (rc) return rc;
        if (!ok) {                                       // not with this painter: let the copies already queued finish, then the plain path
            CK(cudaStreamSynchronize(so->copy));
            return BJT_OK;
        }
        px += c->paint_ms[0]; py += c->paint_ms[1]; nlaunch_disc += c->paint_launches[1];
        // the cleanup of a plane reads two painted planes either side
        rc = finish_to(z1 == d ? d : z1 - 2);
        if (rc) return rc;
    }
    if (so->last) CK(cudaEventRecord(so->last, so->copy));
    c->paint_ms[0] = px; c->paint_ms[1] = py; c->paint_ms[2] = 0.f;
    c->paint_launches[0] = nlaunch_disc; c->paint_launches[1] = nlaunch_disc; c->paint_launches[2] = 0;
    if (!any) { tot.mn = INFINITY; tot.mx = -INFINITY; }
    stats_out(tot, stats);
    *ms_paint = px + py; *ms_finish = pf;
    so->used = true;
    *done = 1;
    return BJT_OK;
}

// the whole chain on device buffers
static int dev_pipeline(bjt_ctx *c, const uint8_t *d_in, int w, int h, int d, int inverse, int mask, int thresh,
                        double pixel_width, float *d_out, bjt_stats *stats, bjt_timing *tm, StreamOut *so = nullptr) {
    const size_t N = (size_t)w * h * d;
    const int n = w > h ? (w > d ? w : d) : (h > d ? h : d);
    const uint32_t n0 = no_result(n);
    uint16_t *gx; uint32_t *bufA, *bufB; uint32_t *scal;
    WS("gx", N * sizeof(uint16_t), gx);
    WS("bufA", N * sizeof(uint32_t), bufA);
    WS("bufB", N * sizeof(uint32_t), bufB);
    WS("scalars", 256, scal);
    const int64_t launches0 = c->launches;

    // input check of LocalThicknessWrapper.processImage: only 0 and 255
    CK(cudaMemsetAsync(scal + 12, 0, 4, c->stream));
    c->launches += launch_check_binary(d_in, N, scal + 12, c->stream); CKL();
    CK(peek(c, c->h_scal + 12, scal + 12, 4));

    uint32_t maxval = 0;
    int rc = dev_edt(c, d_in, w, h, d, inverse, thresh, n, gx, bufA, bufB, &maxval, c->ev);
    if (rc) return rc;
    if (c->h_scal[12]) return fail(c, BJT_E_NOT_BINARY, "input is not a binary 0/255 image");

    int64_t n_ridge = 0;
    int path_used = 3;
    uint32_t *rsq = bufB;
    bool streamed = false;
    float streamed_paint_ms = 0.f, streamed_finish_ms = 0.f;
    if (maxval == 0) {
        // no foreground for this run: sq (bufB) is all zero and doubles as the painted map
        CK(cudaEventRecord(c->ev[4], c->stream));
        CK(cudaEventRecord(c->ev[5], c->stream));
    } else if (maxval >= n0) {
        // no background anywhere: every voxel carries the sentinel radius and is covered by every ball
        c->launches += launch_fill_u32(bufB, N, n0, c->stream); CKL();
        n_ridge = (int64_t)N;
        CK(cudaEventRecord(c->ev[4], c->stream));
        CK(cudaEventRecord(c->ev[5], c->stream));
    } else {
        rc = dev_ridge(c, bufB, w, h, d, maxval, bufA, &n_ridge);
        if (rc) return rc;
        CK(cudaEventRecord(c->ev[4], c->stream));
        if (so && d_out && c->paint_path < 0 && paint_disc_supports(maxval)) {
            // paint, finish and download chunk by chunk
            int done = 0;
            rc = dev_paint_finish_streamed(c, bufA, mask ? d_in : nullptr, w, h, d, inverse, pixel_width, maxval, bufB, d_out, so, stats,
                                           &streamed_paint_ms, &streamed_finish_ms, &done);
            if (rc) return rc;
            if (done) { streamed = true; path_used = 4; }
        }
        if (!streamed) {
            rc = dev_paint(c, bufA, w, h, d, -1, n_ridge, (int64_t)maxval, bufB, &path_used);
            if (rc) return rc;
        }
        CK(cudaEventRecord(c->ev[5], c->stream));
    }
    if (!streamed) {
        rc = dev_finish(c, rsq, mask ? d_in : nullptr, w, h, d, inverse, 1, pixel_width, d_out, stats);
        if (rc) return rc;
    }
    CK(cudaEventRecord(c->ev[6], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (tm) {
        cudaEventElapsedTime(&tm->ms_edt_x, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&tm->ms_edt_y, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&tm->ms_edt_z, c->ev[2], c->ev[3]);
        cudaEventElapsedTime(&tm->ms_ridge, c->ev[3], c->ev[4]);
        cudaEventElapsedTime(&tm->ms_paint, c->ev[4], c->ev[5]);
        cudaEventElapsedTime(&tm->ms_finish, c->ev[5], c->ev[6]);
        if (streamed) { tm->ms_paint = streamed_paint_ms; tm->ms_finish = streamed_finish_ms; }   // interleaved: the sums of their pieces
        tm->ms_cleanup = tm->ms_finish;
        cudaEventElapsedTime(&tm->ms_total, c->ev[0], c->ev[6]);
        tm->n_ridge = n_ridge;
        tm->rsq_max = maxval;
        tm->paint_path = path_used;
        tm->n_launches = c->launches - launches0;
        const bool sep = path_used == 0 || path_used == 1 || path_used == 4;
        tm->ms_paint_x = sep ? c->paint_ms[0] : 0.f;
        tm->ms_paint_y = sep ? c->paint_ms[1] : 0.f;
        tm->ms_paint_z = sep ? c->paint_ms[2] : 0.f;
        tm->n_paint_y_sweeps = sep ? c->paint_launches[1] : 0;
    }
    return BJT_OK;
}



// n strided blocks (rows of `width` bytes) in one kernel launch when everything is 16-byte aligned, else one
// cudaMemcpy2DAsync each; on the context's stream
struct Block2D { void *dst; size_t dpitch; const void *src; size_t spitch, width, height; };
static int dev_copy_blocks(bjt_ctx *c, const Block2D *blk, int n) {
    if (c->sm_count == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, c->device);
        c->sm_count = v > 0 ? v : 148;
    }
    CopyBatch b;
    b.n = 0;
    for (int i = 0; i < n; ++i) {
        const Block2D &k = blk[i];
        if (k.width == 0 || k.height == 0) continue;
        const bool vec = ((uintptr_t)k.dst | (uintptr_t)k.src | k.dpitch | k.spitch | k.width) % 16 == 0;
        if (!vec) {
            CK(cudaMemcpy2DAsync(k.dst, k.dpitch, k.src, k.spitch, k.width, k.height, cudaMemcpyDefault, c->stream));
            continue;
        }
        b.d[b.n++] = CopyDesc{reinterpret_cast<const uint4 *>(k.src), reinterpret_cast<uint4 *>(k.dst), k.spitch / 16, k.dpitch / 16,
                              k.width / 16, k.height};
        if (b.n == COPY_BATCH_MAX) { c->launches += launch_copy2d_batch(b, c->sm_count, c->stream); CKL(); b.n = 0; }
    }
    if (b.n) { c->launches += launch_copy2d_batch(b, c->sm_count, c->stream); CKL(); }
    return BJT_OK;
}

// ---------------------------------------------------------------------------------------------------
// z-slabs over several GPUs behind the same calls (bjt_create_multi).  The reference seam is one blocking call
// on one thread (ThicknessWrapper.java:127-134, 190-196); here that call cuts the stack into contiguous z-slabs,
// one per device, and one host thread per device drives its slab.  A run is three phases, joined on the host:
//   A  upload the slab (first run of a call), 0/255 check, EDT x and y                       slab-local
//   B  pull my y range of every slab's y-pass result out of the peers' memory (one strided copy per peer over
//      NVLink), EDT z on whole z lines, largest squared distance
//   C  pull the squared EDT of my planes +- (H + 3), H = floor(sqrt(largest)), straight out of the peers' y-slab
//      layout (transpose back and halo in one copy per peer), ridge, paint [slab +- 2], cleanup .. statistics on
//      the slab, download
// No ball reaches further than H planes, the ridge of a plane needs one plane either side, the cleanup two: the
// maps are bitwise those of a single device (the same argument as bonej2_b200/sharded.py, which runs the same
// split with one process per GPU).  (n, sum, sum^2, min, max) are combined before mean and sd are formed.

static void split_range(int n, int parts, int i, int *lo, int *hi) {
    const int base = n / parts, rem = n % parts;
    *lo = i * base + (i < rem ? i : rem);
    *hi = *lo + base + (i < rem ? 1 : 0);
}

// f(r) for every slab: one host thread each (the phases block on their own streams), joined before returning.
// Build option BJT_SINGLE_HOST_THREAD: the slabs take turns on the calling thread (for a runtime that must not be
// entered from several threads).
template <class F> static int for_each_slab(bjt_ctx *c, int G, F f) {
    std::vector<int> rcs((size_t)G, BJT_OK);
#ifdef BJT_SINGLE_HOST_THREAD
    for (int r = 0; r < G; ++r) rcs[r] = f(r);
#else
    std::vector<std::thread> th;
    for (int r = 1; r < G; ++r) th.emplace_back([&, r]() { DeviceGuard dg(c->kids[r]->device); rcs[r] = f(r); });
    { DeviceGuard dg(c->kids[0]->device); rcs[0] = f(0); }
    for (auto &t : th) t.join();
#endif
    for (int r = 0; r < G; ++r)
        if (rcs[r]) { snprintf(c->err, sizeof c->err, "slab %d (device %d): %.400s", r, c->kids[r]->device, c->kids[r]->err); return rcs[r]; }
    return BJT_OK;
}

static int multi_run(bjt_ctx *P, const uint8_t *in, const uint8_t *const *in_slices, int w, int h, int d, int both,
                     int inverse, int mask, int threshold, double pixel_width, float *out0, float *const *out_slices,
                     float *out1, bjt_stats *st0, bjt_stats *st1, bjt_timing *tm0, bjt_timing *tm1) {
    const int G = (int)std::min<size_t>(P->kids.size(), (size_t)d);      // never an empty slab
    const size_t wh = (size_t)w * h;
    const int n = w > h ? (w > d ? w : d) : (h > d ? h : d);
    const uint32_t n0 = no_result(n);
    if (tm0) memset(tm0, 0, sizeof *tm0);
    if (tm1) memset(tm1, 0, sizeof *tm1);
    auto zr = [&](int r, int *lo, int *hi) { split_range(d, G, r, lo, hi); };
    auto yr = [&](int r, int *lo, int *hi) { split_range(h, G, r, lo, hi); };
    int64_t launches0 = 0;
    for (int r = 0; r < G; ++r) launches0 += P->kids[r]->launches;

    for (int run = 0; run < (both ? 2 : 1); ++run) {
        const int inv = both ? run : inverse;
        float *out = run == 0 ? out0 : out1;
        float *const *outs = run == 0 ? out_slices : nullptr;
        const bool want_map = out || outs;
        bjt_timing *tm = run == 0 ? tm0 : tm1;
        const auto t_start = std::chrono::steady_clock::now();

        // ---- phase A ------------------------------------------------------------------------------
        int rc = for_each_slab(P, G, [&](int r) -> int {
            bjt_ctx *c = P->kids[r], *k = c;                           // CK / WS report into the slab's context
            int z0, z1; zr(r, &z0, &z1);
            const int dz = z1 - z0;
            const size_t Ns = (size_t)dz * wh;
            uint8_t *d_in; uint16_t *gx; uint32_t *gy, *scal;
            WS("in", Ns, d_in); WS("gx", Ns * 2, gx); WS("mg_gy", Ns * 4, gy); WS("scalars", 256, scal);
            k->mg.gy = gy;
            if (run == 0) {
                if (in) CK(cudaMemcpyAsync(d_in, in + (size_t)z0 * wh, Ns, cudaMemcpyHostToDevice, k->stream));
                else for (int z = z0; z < z1; ++z)
                    CK(cudaMemcpyAsync(d_in + (size_t)(z - z0) * wh, in_slices[z], wh, cudaMemcpyHostToDevice, k->stream));
                CK(cudaMemsetAsync(scal + 12, 0, 4, k->stream));
                k->launches += launch_check_binary(d_in, Ns, scal + 12, k->stream); CKL();
                CK(peek(k, k->h_scal + 12, scal + 12, 4));
            }
            k->launches += launch_edt_x(d_in, w, h, dz, inv, threshold, gx, k->stream); CKL();
            k->launches += launch_edt_y(gx, w, h, dz, n0, gy, k->stream); CKL();
            CK(cudaStreamSynchronize(k->stream));
            if (run == 0) k->mg.not_binary = k->h_scal[12];
            return BJT_OK;
        });
        if (rc) return rc;
        for (int r = 0; r < G; ++r)
            if (P->kids[r]->mg.not_binary) return fail(P, BJT_E_NOT_BINARY, "input is not a binary 0/255 image");

        // ---- phase B ------------------------------------------------------------------------------
        rc = for_each_slab(P, G, [&](int r) -> int {
            bjt_ctx *c = P->kids[r], *k = c;
            int y0, y1; yr(r, &y0, &y1);
            const int hy = y1 - y0;
            k->mg.maxval = 0;
            uint32_t *gyt, *sqt, *scal;
            WS("mg_gyt", (size_t)d * hy * w * 4, gyt); WS("mg_sqt", (size_t)d * hy * w * 4, sqt); WS("scalars", 256, scal);
            k->mg.sqt = sqt;
            if (hy == 0) return BJT_OK;
            const size_t row = (size_t)hy * w * 4;                     // my y range of one plane: contiguous on either side
            std::vector<Block2D> blk;
            for (int i = 0; i < G; ++i) {
                const int p = (r + i) % G;                             // every device starts with a different peer
                int pz0, pz1; zr(p, &pz0, &pz1);
                blk.push_back(Block2D{gyt + (size_t)pz0 * hy * w, row, P->kids[p]->mg.gy + (size_t)y0 * w, wh * 4, row, (size_t)(pz1 - pz0)});
            }
            { const int rcb = dev_copy_blocks(k, blk.data(), (int)blk.size()); if (rcb) return rcb; }
            CK(cudaMemsetAsync(scal, 0, 4, k->stream));
            k->launches += launch_edt_z(gyt, w, hy, d, n0, sqt, nullptr, scal, k->stream); CKL();
            CK(peek(k, k->h_scal, scal, 4));
            CK(cudaStreamSynchronize(k->stream));
            k->mg.maxval = k->h_scal[0];
            return BJT_OK;
        });
        if (rc) return rc;
        uint32_t gmax = 0;
        for (int r = 0; r < G; ++r) gmax = std::max(gmax, P->kids[r]->mg.maxval);
        const int H = (int)std::floor(std::sqrt((double)gmax));

        // ---- phase C ------------------------------------------------------------------------------
        rc = for_each_slab(P, G, [&](int r) -> int {
            bjt_ctx *c = P->kids[r], *k = c;
            int z0, z1; zr(r, &z0, &z1);
            const int dz = z1 - z0;
            const int s_lo = std::max(0, z0 - 2), s_hi = std::min(d, z1 + 2);      // painted planes the cleanup reads
            const int ds = s_hi - s_lo;
            uint8_t *d_in; uint32_t *rsq; float *d_out;
            WS("in", (size_t)dz * wh, d_in);
            WS("mg_rsq", (size_t)ds * wh * 4, rsq);
            // "Both": the first map goes home on the copy stream while the second run computes into another buffer
            const bool second_buf = both && run == 1 && out0 && k->copy_stream;
            if (second_buf) WS("out2", (size_t)dz * wh * 4, d_out); else WS("out", (size_t)dz * wh * 4, d_out);
            if (gmax == 0 || gmax >= n0) {
                // no foreground for this run, or no background anywhere (every voxel carries the sentinel radius)
                if (gmax == 0) CK(cudaMemsetAsync(rsq, 0, (size_t)ds * wh * 4, k->stream));
                else { k->launches += launch_fill_u32(rsq, (size_t)ds * wh, n0, k->stream); CKL(); }
            } else {
                const int e_lo = std::max(0, z0 - 2 - H - 1), e_hi = std::min(d, z1 + 2 + H + 1);
                const int de = e_hi - e_lo;
                uint32_t *sq_ext, *ridge;
                WS("mg_sq_ext", (size_t)de * wh * 4, sq_ext); WS("mg_ridge", (size_t)de * wh * 4, ridge);
                std::vector<Block2D> blk;
                for (int i = 0; i < G; ++i) {
                    const int p = (r + i) % G;
                    int py0, py1; yr(p, &py0, &py1);
                    const size_t prow = (size_t)(py1 - py0) * w * 4;
                    blk.push_back(Block2D{sq_ext + (size_t)py0 * w, wh * 4, P->kids[p]->mg.sqt + (size_t)e_lo * (py1 - py0) * w, prow, prow, (size_t)de});
                }
                { const int rcb = dev_copy_blocks(k, blk.data(), (int)blk.size()); if (rcb) return rcb; }
                int rc2 = dev_ridge(k, sq_ext, w, h, de, gmax, ridge, nullptr);
                if (rc2) return rc2;
                // the outermost fetched planes have an incomplete neighbourhood unless they are volume faces; they lie
                // beyond reach of [s_lo, s_hi) and must not paint
                if (e_lo > 0) CK(cudaMemsetAsync(ridge, 0, wh * 4, k->stream));
                if (e_hi < d) CK(cudaMemsetAsync(ridge + (size_t)(de - 1) * wh, 0, wh * 4, k->stream));
                // paint in chunks of planes, finish the chunk before (the cleanup reads two painted planes either side) and
                // send it home on the copy stream while the next chunk is painted
                const uint8_t *orig = mask ? d_in - (size_t)(z0 - s_lo) * wh : nullptr;
                cudaStream_t cs = k->copy_stream ? k->copy_stream : k->stream;
                const int quarter = (ds + 3) / 4, chunk = std::max(32, (quarter + 31) / 32 * 32);
                StatsPartial acc{0.0, 0.0, 0.0, 0.0, 0};
                bool any = false;
                int fin = z0;
                for (int c0 = s_lo; c0 < s_hi; c0 += chunk) {
                    const int c1 = std::min(s_hi, c0 + chunk);
                    rc2 = dev_paint_range(k, ridge, w, h, de, c0 - e_lo, c1 - e_lo, (int64_t)gmax, rsq + (size_t)(c0 - s_lo) * wh);
                    if (rc2) return rc2;
                    const int upto = c1 == s_hi ? z1 : std::min(z1, c1 - 2);
                    if (upto <= fin) continue;
                    bjt_stats st;
                    rc2 = dev_finish_range(k, rsq, orig, w, h, ds, fin - s_lo, upto - s_lo, inv, 1, pixel_width,
                                           want_map ? d_out + (size_t)(fin - z0) * wh : nullptr, &st);
                    if (rc2) return rc2;                           // returns with the compute stream idle: the planes are final
                    if (st.count > 0) {
                        acc.sum += st.sum; acc.sum2 += st.sum2; acc.count += st.count;
                        acc.mn = any ? std::min(acc.mn, st.min) : st.min;
                        acc.mx = any ? std::max(acc.mx, st.max) : st.max;
                        any = true;
                    }
                    if (out) CK(cudaMemcpyAsync(out + (size_t)fin * wh, d_out + (size_t)(fin - z0) * wh, (size_t)(upto - fin) * wh * 4, cudaMemcpyDeviceToHost, cs));
                    else if (outs) for (int z = fin; z < upto; ++z)
                        CK(cudaMemcpyAsync(outs[z], d_out + (size_t)(z - z0) * wh, wh * 4, cudaMemcpyDeviceToHost, cs));
                    fin = upto;
                }
                if (!any) { acc.mn = INFINITY; acc.mx = -INFINITY; }
                k->mg.part = acc;
                return BJT_OK;                                     // the copies are joined after the last run
            }
            bjt_stats st;
            // the kernel indexes the original with the coordinates of rsq: shift the slab's pointer
            const uint8_t *orig = mask ? d_in - (size_t)(z0 - s_lo) * wh : nullptr;
            int rc2 = dev_finish_range(k, rsq, orig, w, h, ds, z0 - s_lo, z1 - s_lo, inv, 1, pixel_width, want_map ? d_out : nullptr, &st);
            if (rc2) return rc2;                                       // returns with the compute stream idle
            k->mg.part = *k->h_stats;
            cudaStream_t cs = (both && run == 0 && out0 && k->copy_stream) ? k->copy_stream : k->stream;
            if (out) CK(cudaMemcpyAsync(out + (size_t)z0 * wh, d_out, (size_t)dz * wh * 4, cudaMemcpyDeviceToHost, cs));
            else if (outs) for (int z = z0; z < z1; ++z)
                CK(cudaMemcpyAsync(outs[z], d_out + (size_t)(z - z0) * wh, wh * 4, cudaMemcpyDeviceToHost, cs));
            if (cs == k->stream) CK(cudaStreamSynchronize(cs));
            return BJT_OK;
        });
        if (rc) return rc;

        StatsPartial tot{0.0, 0.0, 0.0, 0.0, 0};
        bool first = true;
        for (int r = 0; r < G; ++r) {                                  // slab order: the same sums whatever the timing
            const StatsPartial &p = P->kids[r]->mg.part;
            if (p.count == 0) continue;
            tot.sum += p.sum; tot.sum2 += p.sum2; tot.count += p.count;
            tot.mn = first ? p.mn : std::min(tot.mn, p.mn);
            tot.mx = first ? p.mx : std::max(tot.mx, p.mx);
            first = false;
        }
        if (first) tot = P->kids[0]->mg.part;                          // empty map: min / max as the kernel leaves them
        stats_out(tot, run == 0 ? st0 : st1);
        if (tm) {
            tm->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_start).count();
            tm->rsq_max = gmax;
            tm->paint_path = (gmax == 0 || gmax >= n0) ? 3 : 4;
        }
    }
    // the first map of a "Both" call may still be on its way
    int rc = for_each_slab(P, G, [&](int r) -> int {
        bjt_ctx *c = P->kids[r];
        if (c->copy_stream) CK(cudaStreamSynchronize(c->copy_stream));
        return BJT_OK;
    });
    if (rc) return rc;
    int64_t launches1 = 0;
    for (int r = 0; r < G; ++r) launches1 += P->kids[r]->launches;
    P->launches += launches1 - launches0;
    if (tm0) tm0->n_launches = launches1 - launches0;
    return BJT_OK;
}

extern "C" {

int bjt_dev_thickness_u8(bjt_ctx *c, const uint8_t *d_in, int w, int h, int d, int inverse, int mask, int threshold,
                         double pixel_width, float *d_out_map, bjt_stats *stats, bjt_timing *timing) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    if (!d_in) return fail(c, BJT_E_ARG, "null input");
    if (timing) memset(timing, 0, sizeof *timing);
    return dev_pipeline(c, d_in, w, h, d, inverse, mask, threshold, pixel_width, d_out_map, stats, timing);
}

static int host_run(bjt_ctx *c, const uint8_t *in, const uint8_t *const *in_slices, int w, int h, int d, int both,
                    int inverse, int mask, int threshold, double pixel_width, float *out0, float *const *out_slices,
                    float *out1, bjt_stats *st0, bjt_stats *st1, bjt_timing *tm0, bjt_timing *tm1) {
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    if (!in && !in_slices) return fail(c, BJT_E_ARG, "null input");
    // every pointer is checked before anything is queued
    if (!in) for (int z = 0; z < d; ++z) if (!in_slices[z]) return fail(c, BJT_E_ARG, "null input slice %d", z);
    if (out_slices) for (int z = 0; z < d; ++z) if (!out_slices[z]) return fail(c, BJT_E_ARG, "null output slice %d", z);
    if (c->kids.size() > 1 && d > 1)
        return multi_run(c, in, in_slices, w, h, d, both, inverse, mask, threshold, pixel_width, out0, out_slices, out1, st0,
                         st1, tm0, tm1);
    // whatever way this function is left, no copy into or out of caller memory is still in flight
    struct Drain { bjt_ctx *c; ~Drain() { cudaStreamSynchronize(c->stream); if (c->copy_stream) cudaStreamSynchronize(c->copy_stream); } } drain{c};
    const size_t N = (size_t)w * h * d, wh = (size_t)w * h;
    uint8_t *d_in; float *d_out;
    WS("in", N, d_in);
    WS("out", N * sizeof(float), d_out);
    if (tm0) memset(tm0, 0, sizeof *tm0);
    if (tm1) memset(tm1, 0, sizeof *tm1);
    CK(cudaEventRecord(c->ev[8], c->stream));
    if (in) CK(cudaMemcpyAsync(d_in, in, N, cudaMemcpyHostToDevice, c->stream));
    else for (int z = 0; z < d; ++z) CK(cudaMemcpyAsync(d_in + z * wh, in_slices[z], wh, cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(c->ev[9], c->stream));
    // "Both": the download of the first map runs on the copy stream while the second run computes into
    // a second device buffer
    const bool overlap = both && out0 && c->copy_stream;
    bool streamed_any = false;
    for (int run = 0; run < (both ? 2 : 1); ++run) {
        const int inv = both ? run : inverse;
        float *out = run == 0 ? out0 : out1;
        bjt_timing *tm = run == 0 ? tm0 : tm1;
        const bool want_map = out || (run == 0 && out_slices);
        float *d_dst = d_out;
        if (run == 1 && overlap) WS("out2", N * sizeof(float), d_dst);
        // the map goes home chunk by chunk behind the painter where that is possible (its own copy stream), else in one
        // piece afterwards
        StreamOut so;
        so.host = out; so.host_slices = (run == 0 && !out) ? out_slices : nullptr; so.copy = c->copy_stream;
        so.first = c->ev[10 + 2 * run]; so.last = c->ev[11 + 2 * run];
        const bool can_stream = want_map && c->copy_stream && (so.host || so.host_slices);
        rc = dev_pipeline(c, d_in, w, h, d, inv, mask, threshold, pixel_width, want_map ? d_dst : nullptr,
                          run == 0 ? st0 : st1, tm, can_stream ? &so : nullptr);      // returns with the compute stream idle
        if (rc) return rc;
        if (so.used) { streamed_any = true; continue; }
        cudaStream_t cs = (run == 0 && overlap) ? c->copy_stream : c->stream;
        CK(cudaEventRecord(c->ev[10 + 2 * run], cs));
        if (out) CK(cudaMemcpyAsync(out, d_dst, N * sizeof(float), cudaMemcpyDeviceToHost, cs));
        else if (run == 0 && out_slices) for (int z = 0; z < d; ++z)
            CK(cudaMemcpyAsync(out_slices[z], d_dst + z * wh, wh * sizeof(float), cudaMemcpyDeviceToHost, cs));
        CK(cudaEventRecord(c->ev[11 + 2 * run], cs));
        if (!(run == 0 && overlap)) CK(cudaStreamSynchronize(cs));
    }
    if (overlap || streamed_any) CK(cudaStreamSynchronize(c->copy_stream));
    for (int run = 0; run < (both ? 2 : 1); ++run) {
        bjt_timing *tm = run == 0 ? tm0 : tm1;
        if (!tm) continue;
        if (run == 0) cudaEventElapsedTime(&tm->ms_h2d, c->ev[8], c->ev[9]);
        cudaEventElapsedTime(&tm->ms_d2h, c->ev[10 + 2 * run], c->ev[11 + 2 * run]);
        tm->ms_total += tm->ms_h2d + tm->ms_d2h;
    }
    return BJT_OK;
}

int bjt_thickness_u8(bjt_ctx *c, const uint8_t *in, int w, int h, int d, int inverse, int mask, int threshold,
                     double pixel_width, float *out_map, bjt_stats *stats, bjt_timing *timing) {
    ENTER(c);
    return host_run(c, in, nullptr, w, h, d, 0, inverse, mask, threshold, pixel_width, out_map, nullptr, nullptr, stats,
                    nullptr, timing, nullptr);
}

int bjt_thickness_u8_slices(bjt_ctx *c, const uint8_t *const *in_slices, int w, int h, int d, int inverse, int mask,
                            int threshold, double pixel_width, float *const *out_slices, bjt_stats *stats,
                            bjt_timing *timing) {
    ENTER(c);
    if (!in_slices) return fail(c, BJT_E_ARG, "null input");
    return host_run(c, nullptr, in_slices, w, h, d, 0, inverse, mask, threshold, pixel_width, nullptr, out_slices,
                    nullptr, stats, nullptr, timing, nullptr);
}

int bjt_thickness_both_u8(bjt_ctx *c, const uint8_t *in, int w, int h, int d, int mask, int threshold,
                          double pixel_width, float *out_th, float *out_sp, bjt_stats *stats_th, bjt_stats *stats_sp,
                          bjt_timing *timing_th, bjt_timing *timing_sp) {
    ENTER(c);
    return host_run(c, in, nullptr, w, h, d, 1, 0, mask, threshold, pixel_width, out_th, nullptr, out_sp, stats_th,
                    stats_sp, timing_th, timing_sp);
}

// ---- stage-level, host buffers ---------------------------------------------------------------------

int bjt_edt_sq_u8(bjt_ctx *c, const uint8_t *in, int w, int h, int d, int inverse, int threshold, uint32_t *out_sq) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    if (!in || !out_sq) return fail(c, BJT_E_ARG, "null buffer");
    const size_t N = (size_t)w * h * d;
    const int n = w > h ? (w > d ? w : d) : (h > d ? h : d);
    uint8_t *d_in; uint16_t *gx; uint32_t *bufA, *bufB;
    WS("in", N, d_in); WS("gx", N * 2, gx); WS("bufA", N * 4, bufA); WS("bufB", N * 4, bufB);
    CK(cudaMemcpyAsync(d_in, in, N, cudaMemcpyHostToDevice, c->stream));
    uint32_t maxval;
    rc = dev_edt(c, d_in, w, h, d, inverse, threshold, n, gx, bufA, bufB, &maxval, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out_sq, bufB, N * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return BJT_OK;
}

int bjt_ridge_u32(bjt_ctx *c, const uint32_t *sq, int w, int h, int d, uint32_t *out_ridge) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    if (!sq || !out_ridge) return fail(c, BJT_E_ARG, "null buffer");
    const size_t N = (size_t)w * h * d;
    uint32_t *bufA, *bufB;
    WS("bufA", N * 4, bufA); WS("bufB", N * 4, bufB);
    CK(cudaMemcpyAsync(bufB, sq, N * 4, cudaMemcpyHostToDevice, c->stream));
    rc = dev_ridge_full(c, bufB, w, h, d, bufA, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out_ridge, bufA, N * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return BJT_OK;
}

int bjt_ridge_template(bjt_ctx *c, const int32_t *rsq_values, int n, int32_t *t1, int32_t *t2, int32_t *t3) {
    ENTER(c);
    if (n < 0 || (n > 0 && (!rsq_values || !t1 || !t2 || !t3))) return fail(c, BJT_E_ARG, "bad template arguments");
    if (n == 0) return BJT_OK;
    uint32_t *v, *a, *b, *e;
    const size_t bytes = (size_t)n * 4;
    WS("values", bytes, v); WS("t1", bytes, a); WS("t2", bytes, b); WS("t3", bytes, e);
    CK(cudaMemcpyAsync(v, rsq_values, bytes, cudaMemcpyHostToDevice, c->stream));
    c->launches += launch_ridge_template_compact(v, (uint32_t)n, a, b, e, c->stream); CKL();
    CK(cudaMemcpyAsync(t1, a, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(t2, b, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(t3, e, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return BJT_OK;
}

int bjt_paint_rsq(bjt_ctx *c, const uint32_t *ridge, int w, int h, int d, int path, uint32_t *out_rsq) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    if (!ridge || !out_rsq) return fail(c, BJT_E_ARG, "null buffer");
    const size_t N = (size_t)w * h * d;
    uint32_t *bufA, *bufB;
    WS("bufA", N * 4, bufA); WS("bufB", N * 4, bufB);
    CK(cudaMemcpyAsync(bufA, ridge, N * 4, cudaMemcpyHostToDevice, c->stream));
    const int saved = c->paint_path;
    if (path >= 0) c->paint_path = path;
    rc = dev_paint(c, bufA, w, h, d, path, -1, -1, bufB, nullptr);
    c->paint_path = saved;
    if (rc) return rc;
    CK(cudaMemcpyAsync(out_rsq, bufB, N * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return BJT_OK;
}

int bjt_cleanup_f32(bjt_ctx *c, const uint32_t *rsq, const uint8_t *original, int w, int h, int d, int inverse,
                    int calibrate, double pixel_width, float *out_map, bjt_stats *stats) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    if (!rsq) return fail(c, BJT_E_ARG, "null buffer");
    const size_t N = (size_t)w * h * d;
    uint32_t *bufB; uint8_t *d_in = nullptr; float *d_out;
    WS("bufB", N * 4, bufB); WS("out", N * 4, d_out);
    CK(cudaMemcpyAsync(bufB, rsq, N * 4, cudaMemcpyHostToDevice, c->stream));
    if (original) { WS("in", N, d_in); CK(cudaMemcpyAsync(d_in, original, N, cudaMemcpyHostToDevice, c->stream)); }
    rc = dev_finish(c, bufB, d_in, w, h, d, inverse, calibrate, pixel_width, d_out, stats);
    if (rc) return rc;
    if (out_map) CK(cudaMemcpyAsync(out_map, d_out, N * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return BJT_OK;
}

int bjt_stats_f32(bjt_ctx *c, const float *map, size_t n, bjt_stats *stats) {
    ENTER(c);
    if (!map || !stats || n == 0) return fail(c, BJT_E_ARG, "bad stats arguments");
    float *d_out; StatsPartial *partials, *result;
    const int np = 148 * 8;
    WS("out", n * 4, d_out); WS("partials", (size_t)np * sizeof(StatsPartial), partials);
    WS("stats_result", sizeof(StatsPartial), result);
    CK(cudaMemcpyAsync(d_out, map, n * 4, cudaMemcpyHostToDevice, c->stream));
    c->launches += launch_stats(d_out, n, partials, np, result, c->stream); CKL();
    CK(peek(c, c->h_stats, result, sizeof(StatsPartial)));
    CK(cudaStreamSynchronize(c->stream));
    stats_out(*c->h_stats, stats);
    return BJT_OK;
}

// ---- stage-level, device buffers -------------------------------------------------------------------

int bjt_dev_edt_x(bjt_ctx *c, const uint8_t *d_in, int w, int h, int d, int inverse, int threshold, uint16_t *d_gx) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    c->launches += launch_edt_x(d_in, w, h, d, inverse, threshold, d_gx, c->stream); CKL();
    return BJT_OK;
}

int bjt_dev_edt_y(bjt_ctx *c, const uint16_t *d_gx, int w, int h, int d, int n_sentinel, uint32_t *d_gy) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    c->launches += launch_edt_y(d_gx, w, h, d, no_result(n_sentinel), d_gy, c->stream); CKL();
    return BJT_OK;
}

int bjt_dev_edt_z(bjt_ctx *c, const uint32_t *d_gy, int w, int h, int d, int n_sentinel, uint32_t *d_sq) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    const uint32_t n0 = no_result(n_sentinel);
    uint32_t *scal;
    WS("scalars", 256, scal);
    CK(cudaMemsetAsync(scal, 0, 4, c->stream));
    c->launches += launch_edt_z(d_gy, w, h, d, n0, d_sq, nullptr, scal, c->stream); CKL();
    return BJT_OK;
}

int bjt_dev_ridge(bjt_ctx *c, const uint32_t *d_sq, int w, int h, int d, uint32_t *d_ridge) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    return dev_ridge_full(c, d_sq, w, h, d, d_ridge, nullptr);
}

int bjt_dev_paint(bjt_ctx *c, const uint32_t *d_ridge, int w, int h, int d, int path, uint32_t *d_rsq) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    const int saved = c->paint_path;
    if (path >= 0) c->paint_path = path;
    rc = dev_paint(c, d_ridge, w, h, d, path, -1, -1, d_rsq, nullptr);
    c->paint_path = saved;
    return rc;
}

// ---- consumers of the map: segmented statistics ---------------------------------------------------------
static int dev_label_stats(bjt_ctx *c, const float *d_map, const int32_t *d_labels, size_t n, int n_labels,
                           const int64_t *sizes, double *out) {
    if (!d_map || !d_labels || !sizes || !out || n_labels < 1) return fail(c, BJT_E_ARG, "label_stats: null pointer or no labels");
    double *work, *d_sizes;
    uint32_t *scal;
    WS("scalars", 256, scal);
    WS("seg_work", (size_t)5 * n_labels * sizeof(double), work);
    WS("seg_sizes", (size_t)n_labels * sizeof(double), d_sizes);
    std::vector<double> hs(n_labels);
    for (int i = 0; i < n_labels; ++i) hs[i] = (double)sizes[i];
    CK(cudaMemcpyAsync(d_sizes, hs.data(), hs.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(scal + 9, 0, 4, c->stream));
    c->launches += launch_label_stats(d_map, d_labels, n, n_labels, d_sizes, work, scal + 9, c->stream); CKL();
    std::vector<double> hw((size_t)5 * n_labels);
    CK(cudaMemcpyAsync(hw.data(), work, hw.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(peek(c, c->h_scal + 9, scal + 9, 4));
    CK(cudaStreamSynchronize(c->stream));
    if (c->h_scal[9]) return fail(c, BJT_E_ARG, "label_stats: a label outside [0, %d) on a voxel with a value", n_labels);
    for (int p = 0; p < n_labels; ++p) {
        const double mean = hw[(size_t)3 * n_labels + p], ssq = hw[(size_t)4 * n_labels + p];
        out[3 * p + 0] = p ? mean : 0.0;                       // the reference never fills mean / sd of label 0
        out[3 * p + 1] = p ? std::sqrt(ssq / (double)sizes[p]) : 0.0;
        out[3 * p + 2] = hw[(size_t)n_labels + p];
    }
    return BJT_OK;
}

int bjt_dev_label_stats_f32(bjt_ctx *c, const float *d_map, const int32_t *d_labels, size_t n, int n_labels,
                            const int64_t *sizes, double *out) {
    ENTER(c);
    return dev_label_stats(c, d_map, d_labels, n, n_labels, sizes, out);
}

int bjt_label_stats_f32(bjt_ctx *c, const float *map, const int32_t *labels, size_t n, int n_labels, const int64_t *sizes,
                        double *out) {
    ENTER(c);
    if (!map || !labels || n == 0) return fail(c, BJT_E_ARG, "label_stats: null pointer or empty map");
    float *d_map; int32_t *d_lab;
    WS("out", n * sizeof(float), d_map);
    WS("bufA", n * sizeof(uint32_t), d_lab);
    CK(cudaMemcpyAsync(d_map, map, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_lab, labels, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    return dev_label_stats(c, d_map, d_lab, n, n_labels, sizes, out);
}

static int dev_slice_stats(bjt_ctx *c, const float *d_map, int w, int h, int d, int rx, int ry, int rw, int rh, double *out,
                           int64_t *counts) {
    if (!d_map || !out) return fail(c, BJT_E_ARG, "slice_stats: null pointer");
    if (rx < 0 || ry < 0 || rw < 1 || rh < 1 || rx + rw > w || ry + rh > h)
        return fail(c, BJT_E_ARG, "slice_stats: ROI %d,%d %dx%d outside the %dx%d slice", rx, ry, rw, rh, w, h);
    double *work;
    WS("seg_work", (size_t)5 * d * sizeof(double), work);
    c->launches += launch_slice_stats(d_map, w, h, d, rx, ry, rw, rh, work, c->stream); CKL();
    std::vector<double> hw((size_t)5 * d);
    CK(cudaMemcpyAsync(hw.data(), work, hw.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int z = 0; z < d; ++z) {
        const double cnt = hw[(size_t)2 * d + z];
        out[3 * z + 0] = hw[(size_t)3 * d + z];                // sum / count: NaN for an empty slice, as 0.0 / 0.0 in Java
        out[3 * z + 1] = hw[(size_t)d + z];
        out[3 * z + 2] = std::sqrt(hw[(size_t)4 * d + z] / cnt);
        if (counts) counts[z] = (int64_t)cnt;
    }
    return BJT_OK;
}

int bjt_dev_slice_stats_f32(bjt_ctx *c, const float *d_map, int w, int h, int d, int rx, int ry, int rw, int rh, double *out,
                            int64_t *counts) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    return dev_slice_stats(c, d_map, w, h, d, rx, ry, rw, rh, out, counts);
}

int bjt_slice_stats_f32(bjt_ctx *c, const float *map, int w, int h, int d, int rx, int ry, int rw, int rh, double *out,
                        int64_t *counts) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    if (!map) return fail(c, BJT_E_ARG, "slice_stats: null pointer");
    const size_t n = (size_t)w * h * d;
    float *d_map;
    WS("out", n * sizeof(float), d_map);
    CK(cudaMemcpyAsync(d_map, map, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    return dev_slice_stats(c, d_map, w, h, d, rx, ry, rw, rh, out, counts);
}

// ---- FindRidgePoints: seeds for Ellipsoid Factor ---------------------------------------------------------
int bjt_find_ridge_points_u8(bjt_ctx *c, const uint8_t *in, int w, int h, int d, double fraction, double *out_xyz,
                             int64_t cap, int64_t *n_found, float *ridge_max) {
    ENTER(c);
    int rc = check_dims(c, w + 2, h + 2, d + 2);
    if (rc) return rc;
    if (!in || !n_found || cap < 0 || (cap > 0 && !out_xyz)) return fail(c, BJT_E_ARG, "find_ridge_points: null pointer");
    const int W = w + 2, H = h + 2, D = d + 2;
    const size_t n_in = (size_t)w * h * d, n = (size_t)W * H * D;
    uint8_t *d_raw, *d_pad; uint16_t *gx; uint32_t *bufA, *bufB, *scal; float *fa, *fb; unsigned long long *found;
    WS("in", n, d_pad);
    WS("raw_in", n_in, d_raw);
    WS("gx", n * sizeof(uint16_t), gx);
    WS("bufA", n * sizeof(uint32_t), bufA);
    WS("bufB", n * sizeof(uint32_t), bufB);
    WS("out", n * sizeof(float), fa);
    WS("out2", n * sizeof(float), fb);
    WS("scalars", 256, scal);
    CK(cudaMemcpyAsync(d_raw, in, n_in, cudaMemcpyHostToDevice, c->stream));
    c->launches += launch_pad_binary(d_raw, w, h, d, d_pad, c->stream); CKL();
    uint32_t maxval = 0;
    const int nmax = W > H ? (W > D ? W : D) : (H > D ? H : D);
    rc = dev_edt(c, d_pad, W, H, D, 0, 128, nmax, gx, bufA, bufB, &maxval, c->ev);     // the border is background: no sentinel
    if (rc) return rc;
    *n_found = 0;
    if (ridge_max) *ridge_max = 0.f;
    if (maxval == 0) return BJT_OK;                                                    // no foreground
    float *fc = reinterpret_cast<float *>(bufA);
    c->launches += launch_ridge_points_map(bufB, W, H, D, fa, fb, fc, scal + 16, c->stream); CKL();
    CK(peek(c, c->h_scal + 16, scal + 16, 4));
    CK(cudaStreamSynchronize(c->stream));
    float mx;
    memcpy(&mx, c->h_scal + 16, 4);
    if (ridge_max) *ridge_max = mx;
    const float threshold = (float)fraction * mx;                                       // float x float, as getRealFloat() * getRealFloat()
    const unsigned long long ucap = (unsigned long long)(cap > 0 ? cap : 0);
    WS("points", (size_t)(ucap ? ucap : 1) * sizeof(unsigned long long), found);
    c->launches += launch_ridge_points_collect(fa, n, threshold, found, ucap, (unsigned long long *)(scal + 18), c->stream); CKL();
    CK(peek(c, c->h_scal + 18, scal + 18, 8));
    CK(cudaStreamSynchronize(c->stream));
    unsigned long long cnt;
    memcpy(&cnt, c->h_scal + 18, 8);
    *n_found = (int64_t)cnt;
    const unsigned long long take = cnt < ucap ? cnt : ucap;
    if (take) {
        std::vector<unsigned long long> idx(take);
        CK(cudaMemcpyAsync(idx.data(), found, take * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        std::sort(idx.begin(), idx.end());                                              // the reference's cursor order: x fastest
        for (unsigned long long k = 0; k < take; ++k) {
            const unsigned long long i = idx[k];
            out_xyz[3 * k + 0] = (double)(i % W) - 0.5;
            out_xyz[3 * k + 1] = (double)((i / W) % H) - 0.5;
            out_xyz[3 * k + 2] = (double)(i / ((unsigned long long)W * H)) - 0.5;
        }
    }
    return BJT_OK;
}

// Painted r^2 of planes [z_lo, z_hi) of a ridge volume of d planes (a z-slab with its halo): balls centred on any
// plane of d_ridge count, only the requested planes are painted.  max_rsq: largest tag anywhere in the whole volume
// (the ranks of a sharded run pass the same value), < 0: determined here.
int bjt_dev_paint_range(bjt_ctx *c, const uint32_t *d_ridge, int w, int h, int d, int z_lo, int z_hi, int64_t max_rsq,
                        uint32_t *d_rsq) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    if (!d_ridge || !d_rsq) return fail(c, BJT_E_ARG, "paint_range: null pointer");
    return dev_paint_range(c, d_ridge, w, h, d, z_lo, z_hi, max_rsq, d_rsq);
}

// ---- stepwise painter of one z-slab -------------------------------------------------------------------
int bjt_dev_paint_open(bjt_ctx *c, const uint32_t *d_ridge, int w, int h, int d, int64_t max_rsq) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    if (!d_ridge || max_rsq < 0) return fail(c, BJT_E_ARG, "paint_open: null ridge or negative tag bound");
    uint32_t *scal;
    WS("scalars", 256, scal);
    const int variant = max_rsq >= 65536 ? 1 : 0;
    void *wsp;
    WS("paint_sep", paint_sep_workspace_bytes(w, h, d, variant), wsp);
    CK(cudaMemsetAsync(scal + 8, 0, 4, c->stream));
    c->ps.open = true; c->ps.w = w; c->ps.h = h; c->ps.d = d; c->ps.variant = variant; c->ps.ws = wsp; c->ps.flags = scal + 8;
    c->launches += paint_sep_stage1(d_ridge, w, h, d, variant, wsp, scal + 8, c->stream); CKL();
    return BJT_OK;
}

int bjt_dev_paint_nxslabs(bjt_ctx *c) { return c && c->ps.open ? paint_sep_nxslabs(c->ps.w, c->ps.h, c->ps.d) : 0; }
int bjt_dev_paint_xslab_width(bjt_ctx *c, int xs) { return c && c->ps.open ? paint_sep_xslab_width(c->ps.w, c->ps.h, c->ps.d, xs) : 0; }

int bjt_dev_paint_lists(bjt_ctx *c, int xs) {
    ENTER(c);
    if (!c->ps.open || xs < 0 || xs >= paint_sep_nxslabs(c->ps.w, c->ps.h, c->ps.d)) return fail(c, BJT_E_ARG, "paint_lists: no open painter or bad x-slab");
    c->launches += paint_sep_stage2(c->ps.w, c->ps.h, c->ps.d, c->ps.variant, c->ps.ws, xs, c->ps.flags, c->stream); CKL();
    return BJT_OK;
}

int bjt_dev_paint_export(bjt_ctx *c, int xs, int side, int gap, int nplanes, uint32_t *d_count, void *d_items, int kcap) {
    ENTER(c);
    if (!c->ps.open || xs < 0 || xs >= paint_sep_nxslabs(c->ps.w, c->ps.h, c->ps.d)) return fail(c, BJT_E_ARG, "paint_export: no open painter or bad x-slab");
    if ((side != 1 && side != -1) || gap < 1 || nplanes < 1 || kcap < 1 || !d_count || !d_items)
        return fail(c, BJT_E_ARG, "paint_export: bad side / gap / planes / capacity");
    c->launches += paint_sep_export(c->ps.w, c->ps.h, c->ps.d, c->ps.variant, c->ps.ws, xs, side, gap, nplanes, d_count,
                                    reinterpret_cast<uint2 *>(d_items), kcap, c->ps.flags, c->stream); CKL();
    return BJT_OK;
}

int bjt_dev_paint_sweep(bjt_ctx *c, int xs, int n_lo, const uint32_t *const *lo_count, const void *const *lo_items, int n_hi,
                        const uint32_t *const *hi_count, const void *const *hi_items, int kcap, uint32_t *d_rsq) {
    ENTER(c);
    if (!c->ps.open || xs < 0 || xs >= paint_sep_nxslabs(c->ps.w, c->ps.h, c->ps.d)) return fail(c, BJT_E_ARG, "paint_sweep: no open painter or bad x-slab");
    if (n_lo < 0 || n_hi < 0 || n_lo > SEP_MAX_SOURCES || n_hi > SEP_MAX_SOURCES || !d_rsq)
        return fail(c, BJT_E_ARG, "paint_sweep: at most %d sources per side", SEP_MAX_SOURCES);
    SepImports imp{};
    imp.n_lo = n_lo; imp.n_hi = n_hi; imp.kcap = kcap;
    for (int i = 0; i < n_lo; ++i) { imp.lo_cnt[i] = lo_count[i]; imp.lo_items[i] = reinterpret_cast<const uint2 *>(lo_items[i]); }
    for (int i = 0; i < n_hi; ++i) { imp.hi_cnt[i] = hi_count[i]; imp.hi_items[i] = reinterpret_cast<const uint2 *>(hi_items[i]); }
    c->launches += paint_sep_stage3(c->ps.w, c->ps.h, c->ps.d, c->ps.variant, c->ps.ws, xs, &imp, d_rsq, c->ps.flags, c->stream); CKL();
    return BJT_OK;
}

int bjt_dev_paint_close(bjt_ctx *c, uint32_t *overflow_flags) {
    ENTER(c);
    if (!c->ps.open) return fail(c, BJT_E_ARG, "paint_close: no open painter");
    c->ps.open = false;
    CK(peek(c, c->h_scal + 8, c->ps.flags, 4));
    CK(cudaStreamSynchronize(c->stream));
    if (overflow_flags) *overflow_flags = c->h_scal[8];
    return BJT_OK;
}

int bjt_dev_finish(bjt_ctx *c, const uint32_t *d_rsq, const uint8_t *d_original, int w, int h, int d, int inverse,
                   int calibrate, double pixel_width, float *d_out, bjt_stats *stats) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    return dev_finish(c, d_rsq, d_original, w, h, d, inverse, calibrate, pixel_width, d_out, stats);
}

int bjt_dev_finish_range(bjt_ctx *c, const uint32_t *d_rsq, const uint8_t *d_original, int w, int h, int d, int z_lo,
                         int z_hi, int inverse, int calibrate, double pixel_width, float *d_out, bjt_stats *stats) {
    ENTER(c);
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    return dev_finish_range(c, d_rsq, d_original, w, h, d, z_lo, z_hi, inverse, calibrate, pixel_width, d_out, stats);
}

// ---- buffers other processes read: CUDA IPC over NVLink (one process per GPU; bonej2_b200/sharded.py) ----------
// A rank's y-pass and z-pass results stay in buffers of its own context; the other ranks map them once
// (bjt_ipc_import) and pull their share with strided copies (bjt_dev_copy2d), so the transposes of the z pass and the
// max-radius halo are device-to-device DMA over NVLink with no pack / unpack kernels and no staging buffers.
int bjt_dev_alloc(bjt_ctx *c, const char *name, size_t bytes, void **ptr) {
    ENTER(c);
    if (!name || !ptr) return fail(c, BJT_E_ARG, "dev_alloc: null argument");
    void *p;
    const std::string key = std::string("user:") + name;
    WS(key.c_str(), bytes, p);
    *ptr = p;
    return BJT_OK;
}

int bjt_ipc_export(bjt_ctx *c, const void *ptr, unsigned char *handle64) {
    ENTER(c);
    if (!ptr || !handle64) return fail(c, BJT_E_ARG, "ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void *>(ptr)));
    memcpy(handle64, &h, 64);
    return BJT_OK;
}

int bjt_ipc_import(bjt_ctx *c, const unsigned char *handle64, void **ptr) {
    ENTER(c);
    if (!ptr || !handle64) return fail(c, BJT_E_ARG, "ipc_import: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BJT_OK;
}

int bjt_ipc_close(bjt_ctx *c, void *ptr) {
    ENTER(c);
    if (ptr) CK(cudaIpcCloseMemHandle(ptr));
    return BJT_OK;
}

int bjt_dev_copy2d(bjt_ctx *c, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height) {
    ENTER(c);
    if (!dst || !src) return fail(c, BJT_E_ARG, "copy2d: null pointer");
    const Block2D b{dst, dpitch, src, spitch, width, height};
    return dev_copy_blocks(c, &b, 1);
}

int bjt_dev_copy2d_batch(bjt_ctx *c, int n, void *const *dst, const size_t *dpitch, const void *const *src, const size_t *spitch,
                         const size_t *width, const size_t *height) {
    ENTER(c);
    if (n < 0 || (n > 0 && (!dst || !dpitch || !src || !spitch || !width || !height))) return fail(c, BJT_E_ARG, "copy2d_batch: null argument");
    std::vector<Block2D> blk;
    for (int i = 0; i < n; ++i) {
        if (!dst[i] || !src[i]) return fail(c, BJT_E_ARG, "copy2d_batch: null pointer in block %d", i);
        blk.push_back(Block2D{dst[i], dpitch[i], src[i], spitch[i], width[i], height[i]});
    }
    return dev_copy_blocks(c, blk.data(), n);
}

// ---- phantom ---------------------------------------------------------------------------------------

// rasterise slices [z0, z0 + dz) into d_out; the caller holds the context's lock
static int dev_phantom(bjt_ctx *c, const int32_t *shapes_host, int nshapes, int w, int h, int d, int z0, int dz, uint8_t *d_out) {
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    if (nshapes < 0 || (nshapes > 0 && !shapes_host) || !d_out || z0 < 0 || dz <= 0 || z0 + dz > d)
        return fail(c, BJT_E_ARG, "bad phantom arguments");
    int32_t *d_shapes;
    WS("shapes", (size_t)(nshapes > 0 ? nshapes : 1) * 16 * sizeof(int32_t), d_shapes);
    if (nshapes) CK(cudaMemcpyAsync(d_shapes, shapes_host, (size_t)nshapes * 16 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    c->launches += launch_phantom(d_shapes, nshapes, w, h, d, z0, dz, d_out, c->stream); CKL();
    CK(cudaStreamSynchronize(c->stream));
    return BJT_OK;
}

int bjt_dev_phantom_u8(bjt_ctx *c, const int32_t *shapes_host, int nshapes, int w, int h, int d, int z0, int dz,
                       uint8_t *d_out) {
    ENTER(c);
    return dev_phantom(c, shapes_host, nshapes, w, h, d, z0, dz, d_out);
}

int bjt_phantom_u8(bjt_ctx *c, const int32_t *shapes, int nshapes, int w, int h, int d, uint8_t *out) {
    ENTER(c);                                            // one lock for the whole call: the context's input buffer is in use throughout
    if (!out) return fail(c, BJT_E_ARG, "null output");
    int rc = check_dims(c, w, h, d);
    if (rc) return rc;
    uint8_t *d_in;
    WS("in", (size_t)w * h * d, d_in);
    rc = dev_phantom(c, shapes, nshapes, w, h, d, 0, d, d_in);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, d_in, (size_t)w * h * d, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return BJT_OK;
}

}  // extern "C"
