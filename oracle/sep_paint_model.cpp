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
