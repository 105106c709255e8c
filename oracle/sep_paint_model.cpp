/*
 * sep_paint_model.cpp -- CPU model of the separable sphere painter.  TEST INFRASTRUCTURE ONLY
 * (lives beside the oracle; the product never links it).  It drives the very same per-line code
 * the CUDA kernels run (bonej2_b200/csrc/paint_sep_core.cuh is plain C++), one line after another
 * on the host, with the same list layout (forward and backward fronts kept as separate lists), so
 * the algorithm can be checked against lto_paint -- the literal Local_Thickness_Parallel
 * restatement -- without a GPU, and its list statistics can be measured.
 */
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../bonej2_b200/csrc/paint_sep_core.cuh"

using namespace bjt;

extern "C" {

typedef struct {
    long overflow_act, overflow_front;
    long max_active, max_front, max_stair;
    long items[2][2];          /* [stage 1,2][fwd,bwd] emitted items */
    double sum_active[3];
    long steps[3];
} spm_stats_t;

}

namespace {
constexpr int CAP = 1024, FMAX = 512;

struct Lists {
    std::vector<SepItem> pool;
    std::vector<uint64_t> off;
    std::vector<uint32_t> cnt;
    explicit Lists(size_t n) : off(n, 0), cnt(n, 0) {}
};

// stages 1 and 2.  src: ridge values (stage 1) or nullptr; inF/inB: input lists (stage 2)
void dyn_line(const uint32_t *src, const Lists *inF, const Lists *inB, size_t base, size_t stride, int L, int dir,
              Lists &out, spm_stats_t *st, int stage) {
    static SepActive<SepLocalStore<CAP> > A;
    static SepItem fr[FMAX];
    A.clear();
    for (int step = 0; step < L; ++step) {
        const int t = dir > 0 ? step : L - 1 - step;
        const size_t v = base + (size_t)t * stride;
        bool ok = true;
        int nf;
        if (src) {
            SepItem one; one.rho = src[v]; one.R = src[v];
            nf = dir > 0 ? A.step<SepItem, SepUnpackPlain, false, +1>(t, &one, src[v] ? 1 : 0, nullptr, 0, fr, FMAX, ok)
                         : A.step<SepItem, SepUnpackPlain, false, -1>(t, &one, src[v] ? 1 : 0, nullptr, 0, fr, FMAX, ok);
        } else {
            const SepItem *pf = inF->pool.data() + inF->off[v], *pb = inB->pool.data() + inB->off[v];
            const int cf = (int)inF->cnt[v], cb = (int)inB->cnt[v];
            nf = dir > 0 ? A.step<SepItem, SepUnpackPlain, true, +1>(t, pf, cf, pb, cb, fr, FMAX, ok)
                         : A.step<SepItem, SepUnpackPlain, true, -1>(t, pf, cf, pb, cb, fr, FMAX, ok);
        }
        if (!ok) st->overflow_act++;
        if (A.n > st->max_active) st->max_active = A.n;
        st->sum_active[stage] += A.n; st->steps[stage]++;
        if (nf > st->max_front) st->max_front = nf;
        if (nf > FMAX) { st->overflow_front++; nf = FMAX; }
        out.off[v] = out.pool.size();
        out.cnt[v] = (uint32_t)nf;
        out.pool.insert(out.pool.end(), fr, fr + nf);
        st->items[stage][dir > 0 ? 0 : 1] += nf;
    }
}

void stair_line(const Lists &inF, const Lists &inB, size_t base, int L, int dir, uint32_t *out, spm_stats_t *st) {
    static SepStair<SepStairLocal<CAP> > S;
    S.clear();
    for (int step = 0; step < L; ++step) {
        const int x = dir > 0 ? step : L - 1 - step;
        const int u = dir > 0 ? x : -x;
        const size_t v = base + (size_t)x;
        for (int which = 0; which < 2; ++which) {
            const Lists &in = which ? inB : inF;
            int hint = 0;
            for (uint32_t k = 0; k < in.cnt[v]; ++k) {
                const SepItem it = in.pool[in.off[v] + k];
                if (!S.arrive(u, it.rho, it.R, hint)) st->overflow_act++;
            }
        }
        if (S.n > st->max_stair) st->max_stair = S.n;
        st->sum_active[2] += S.n; st->steps[2]++;
        const uint32_t r = S.query(u);
        if (dir > 0) out[v] = r;
        else if (r > out[v]) out[v] = r;
    }
}
}  // namespace

extern "C" int spm_paint(const uint32_t *ridge, int w, int h, int d, uint32_t *out_rsq, spm_stats_t *st) {
    const size_t N = (size_t)w * h * d, wh = (size_t)w * h;
    memset(st, 0, sizeof *st);
    Lists l1f(N), l1b(N), l2f(N), l2b(N);
    /* stage 1: lines along z */
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
        const size_t base = (size_t)y * w + x;
        dyn_line(ridge, nullptr, nullptr, base, wh, d, +1, l1f, st, 0);
        dyn_line(ridge, nullptr, nullptr, base, wh, d, -1, l1b, st, 0);
    }
    /* stage 2: lines along y */
    for (int z = 0; z < d; ++z) for (int x = 0; x < w; ++x) {
        const size_t base = (size_t)z * wh + x;
        dyn_line(nullptr, &l1f, &l1b, base, (size_t)w, h, +1, l2f, st, 1);
        dyn_line(nullptr, &l1f, &l1b, base, (size_t)w, h, -1, l2b, st, 1);
    }
    /* stage 3: rows along x */
    for (size_t row = 0; row < (size_t)h * d; ++row) {
        stair_line(l2f, l2b, row * w, w, +1, out_rsq, st);
        stair_line(l2f, l2b, row * w, w, -1, out_rsq, st);
    }
    return (st->overflow_act || st->overflow_front) ? 1 : 0;
}

/* ---- the painter in steps, as the CUDA library offers it to z-slab sharded runs (bjt_dev_paint_open ..
 * _close in include/bonej_thickness.h), so the exchange logic of bonej2_b200/sharded.py can be tested on
 * CPU over gloo.  One x-slab (the whole width).  Boundary staircases use the kernels' layout: per column
 * y * w + x up to kcap items {reach beyond, tag}, tags descending, and a count. ---- */
namespace {
struct SpmSession {
    int w, h, d;
    Lists *l2f, *l2b;
    long overflow;
};
}

extern "C" void *spm_open(const uint32_t *ridge, int w, int h, int d) {
    const size_t N = (size_t)w * h * d, wh = (size_t)w * h;
    spm_stats_t st;
    memset(&st, 0, sizeof st);
    SpmSession *S = new SpmSession{w, h, d, nullptr, nullptr, 0};
    /* NB: spm_paint sweeps z first and x last, the kernels x first and z last; for a z-slab the plane-local
     * stages must be x and y, so here: stage 1 along x, stage 2 along y, stage 3 along z */
    Lists m1f(N), m1b(N);
    for (size_t row = 0; row < (size_t)h * d; ++row) {
        dyn_line(ridge, nullptr, nullptr, row * w, 1, w, +1, m1f, &st, 0);
        dyn_line(ridge, nullptr, nullptr, row * w, 1, w, -1, m1b, &st, 0);
    }
    S->l2f = new Lists(N); S->l2b = new Lists(N);
    for (int z = 0; z < d; ++z) for (int x = 0; x < w; ++x) {
        const size_t base = (size_t)z * wh + x;
        dyn_line(nullptr, &m1f, &m1b, base, (size_t)w, h, +1, *S->l2f, &st, 1);
        dyn_line(nullptr, &m1f, &m1b, base, (size_t)w, h, -1, *S->l2b, &st, 1);
    }
    S->overflow = st.overflow_act + st.overflow_front;
    return S;
}

static void spm_arrivals(const SpmSession *S, SepStair<SepStairLocal<CAP> > &T, size_t v, int u, long *ovf) {
    for (int which = 0; which < 2; ++which) {
        const Lists &in = which ? *S->l2b : *S->l2f;
        int hint = 0;
        for (uint32_t k = 0; k < in.cnt[v]; ++k) {
            const SepItem it = in.pool[in.off[v] + k];
            if (!T.arrive(u, it.rho, it.R, hint)) ++*ovf;
        }
    }
}

extern "C" int spm_export(void *handle, int side, int gap, int nplanes, uint32_t *cnt, uint32_t *items, int kcap) {
    SpmSession *S = (SpmSession *)handle;
    const size_t wh = (size_t)S->w * S->h;
    static SepStair<SepStairLocal<CAP> > T;
    const int np = nplanes < S->d ? nplanes : S->d;
    for (size_t col = 0; col < wh; ++col) {
        T.clear();
        for (int step = 0; step < np; ++step) {
            const int z = side > 0 ? S->d - np + step : np - 1 - step;
            const int u = side > 0 ? z : -z;
            spm_arrivals(S, T, (size_t)z * wh + col, u, &S->overflow);
            T.query(u);
        }
        const int target = (side > 0 ? S->d - 1 : 0) + gap;
        int nout = 0;
        for (int i = 0; i < T.n; ++i) {
            const int E = T.st.getE(i);
            if (E < target) continue;
            if (nout < kcap) { items[(col * kcap + nout) * 2] = (uint32_t)(E - target); items[(col * kcap + nout) * 2 + 1] = T.st.getR(i); }
            ++nout;
        }
        if (nout > kcap) { S->overflow++; nout = kcap; }
        cnt[col] = (uint32_t)nout;
    }
    return 0;
}

extern "C" int spm_sweep(void *handle, int n_lo, const uint32_t *const *lo_cnt, const uint32_t *const *lo_items, int n_hi,
                         const uint32_t *const *hi_cnt, const uint32_t *const *hi_items, int kcap, uint32_t *out_rsq) {
    SpmSession *S = (SpmSession *)handle;
    const size_t wh = (size_t)S->w * S->h;
    static SepStair<SepStairLocal<CAP> > T;
    for (size_t col = 0; col < wh; ++col) {
        for (int dir = +1; dir >= -1; dir -= 2) {
            T.clear();
            for (int step = 0; step < S->d; ++step) {
                const int z = dir > 0 ? step : S->d - 1 - step;
                const int u = dir > 0 ? z : -z;
                const size_t v = (size_t)z * wh + col;
                spm_arrivals(S, T, v, u, &S->overflow);
                for (int side = 0; side < 2; ++side) {
                    if (z != (side ? S->d - 1 : 0)) continue;
                    const int ns = side ? n_hi : n_lo;
                    for (int src = 0; src < ns; ++src) {
                        const uint32_t c = (side ? hi_cnt : lo_cnt)[src][col];
                        const uint32_t *p = (side ? hi_items : lo_items)[src] + col * kcap * 2;
                        int hint = 0;
                        for (uint32_t k = 0; k < c; ++k)
                            if (!T.arrive(u, p[2 * k], p[2 * k + 1], hint)) S->overflow++;
                    }
                }
                const uint32_t r = T.query(u);
                if (dir > 0) out_rsq[v] = r;
                else if (r > out_rsq[v]) out_rsq[v] = r;
            }
        }
    }
    return 0;
}

/* the kernels' thread -> line map (sep_patch_line), evaluated for a whole grid: lines[tid], -1 = idle thread */
extern "C" long long spm_patch_map(int na, int nb, int threads_per_block, long long *lines, long long capacity) {
    const long long need = sep_patch_threads(na, nb);
    const long long grid = (need + threads_per_block - 1) / threads_per_block * threads_per_block;
    if (grid > capacity) return -grid;
    for (long long tid = 0; tid < grid; ++tid) lines[tid] = sep_patch_line(tid, na, nb);
    return grid;
}

extern "C" long spm_close(void *handle) {
    SpmSession *S = (SpmSession *)handle;
    const long ovf = S->overflow;
    delete S->l2f; delete S->l2b; delete S;
    return ovf;
}
