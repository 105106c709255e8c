// park_study.cpp -- what a two-tier active set would buy in the painter's y stage (CPU measurement, research only).
//
// Today every item of a line's active set is visited at every position, and 43 % of the visits find an item that
// is dominated now but must be kept because the dominator is older (its slack falls faster) and stops dominating
// before the item dies.  How long that lasts is known in closed form (the gap between two slacks is linear in
// the position).  Variant measured here: every K positions ("scan" positions) all items are visited as today and
// a dominated item that provably stays dominated until the next scan is put in a cold tier, which the K-1
// positions in between do not touch; at the next scan the cold tier is merged back in (it is sorted by tag like
// the hot one).  The emitted fronts must be identical to today's; the tool checks that, and counts loop
// iterations per position the way a warp pays for them (max over 32 adjacent lines).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

static long g_cnt[8];
#define BJT_SEP_COUNT(k) (++g_cnt[k])
#include "../../bonej2_b200/csrc/paint_sep_core.cuh"
using namespace bjt;

namespace {
constexpr int CAP = 4096, FMAX = 2048;
struct Lists {
    std::vector<SepItem> pool;
    std::vector<uint64_t> off;
    std::vector<uint32_t> cnt;
    explicit Lists(size_t n) : off(n, 0), cnt(n, 0) {}
};
SepActive<SepLocalStore<CAP> > A;
SepItem fr[FMAX], fr2[FMAX];

inline long floordiv(long a, long b) { long q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }

// two-tier line state
struct Park {
    std::vector<SepEntry> hot[2], cold[2];
    int nh = 0, nc = 0, cur = 0, ccur = 0;
    Park() { for (int i = 0; i < 2; ++i) { hot[i].resize(CAP); cold[i].resize(CAP); } }
    void clear() { nh = nc = 0; }
    long parked = 0;

    // REACH emission only (stage 2).  dir = +-1.  scan: visit the cold tier too and re-route.
    int step(int tc, int dir, const SepItem *lf, int nlf, const SepItem *lb, int nlb, SepItem *out, bool scan, int K, long &iters) {
        const bool skip_own = dir < 0;
        const int src = cur, dst = cur ^ 1;
        const int csrc = ccur, cdst = ccur ^ 1;
        int i = 0, ic = 0, jf = 0, jb = 0, w = 0, wc = 0, nf = 0;
        int best = -1, last_e = -1;
        bool last_emitted = false;
        SepEntry bx{0, 0, tc};
        while (true) {
            // pick the source with the largest tag; ties: forward list, backward list, hot, cold
            uint32_t Rf = jf < nlf ? lf[jf].R : 0, Rb = jb < nlb ? lb[jb].R : 0, Rh = i < nh ? hot[src][i].R : 0,
                     Rc = (scan && ic < nc) ? cold[csrc][ic].R : 0;
            if (!(Rf | Rb | Rh | Rc)) break;
            SepEntry x;
            if (Rf >= Rb && Rf >= Rh && Rf >= Rc) { x = SepEntry{lf[jf].R, lf[jf].rho, tc}; ++jf; }
            else if (Rb >= Rh && Rb >= Rc) { x = SepEntry{lb[jb].R, lb[jb].rho, tc}; ++jb; }
            else if (Rh >= Rc) x = hot[src][i++];
            else x = cold[csrc][ic++];
            ++iters;
            const int dt = tc - x.t;
            const int s = (int)x.rho0 - dt * dt;
            if (s < 0) continue;
            bool to_cold = false;
            if (s > best) {
                const bool same_tag = best >= 0 && bx.R == x.R && last_emitted;
                best = s; bx = x;
                const bool may_emit = !(skip_own && x.t == tc);
                const int e = (int)sep_isqrt((uint32_t)s);
                if (same_tag) { out[nf - 1].rho = (uint32_t)e; last_e = e; }
                else if (e > last_e) {
                    last_e = e; last_emitted = may_emit;
                    if (may_emit) { out[nf].rho = (uint32_t)e; out[nf].R = x.R; ++nf; }
                } else last_emitted = false;
            } else {
                if (dir > 0 ? bx.t >= x.t : bx.t <= x.t) continue;
                const int ex = x.t + dir * sep_sqrt_upper(x.rho0);
                if ((int)bx.rho0 - (int)x.rho0 + (bx.t - x.t) * (2 * ex - bx.t - x.t) >= 0) continue;
                if (scan) {
                    // first sweep coordinate at which x's slack exceeds b's: gap(u) = D - q (2u - sb - sx) < 0
                    const long sb = (long)dir * bx.t, sx = (long)dir * x.t, q = sx - sb, D = (long)bx.rho0 - (long)x.rho0;
                    const long wake = floordiv(sb + sx + floordiv(D, q), 2) + 1;
                    const long snow = (long)dir * tc;
                    to_cold = wake >= snow + K;          // dominated at every position before the next scan
                }
            }
            if (to_cold) { cold[cdst][wc++] = x; ++parked; }
            else hot[dst][w++] = x;
            if (w >= CAP || wc >= CAP) { fprintf(stderr, "cap\n"); exit(1); }
        }
        nh = w; cur = dst;
        if (scan) { nc = wc; ccur = cdst; }
        return nf;
    }
};
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s ridge.u32 w h d\n", argv[0]); return 2; }
    const int w = atoi(argv[2]), h = atoi(argv[3]), d = atoi(argv[4]);
    const size_t N = (size_t)w * h * d, wh = (size_t)w * h;
    std::vector<uint32_t> ridge(N);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(ridge.data(), 4, N, f) != N) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(f);
    Lists l1f(N), l1b(N);
    for (size_t row = 0; row < (size_t)h * d; ++row)
        for (int dir = +1; dir >= -1; dir -= 2) {
            Lists &out = dir > 0 ? l1f : l1b;
            A.clear();
            for (int step = 0; step < w; ++step) {
                const int t = dir > 0 ? step : w - 1 - step;
                const size_t v = row * w + t;
                bool ok = true;
                SepItem one; one.rho = ridge[v]; one.R = ridge[v];
                int nf = 0;
                if (ridge[v] || A.n)
                    nf = dir > 0 ? A.step<SepItem, SepUnpackPlain, false, +1>(t, &one, ridge[v] ? 1 : 0, nullptr, 0, fr, FMAX, ok)
                                 : A.step<SepItem, SepUnpackPlain, false, -1>(t, &one, ridge[v] ? 1 : 0, nullptr, 0, fr, FMAX, ok);
                out.off[v] = out.pool.size(); out.cnt[v] = (uint32_t)nf;
                out.pool.insert(out.pool.end(), fr, fr + nf);
            }
        }
    const int Ks[] = {0, 2, 3, 4, 6};
    std::vector<uint16_t> it(2 * N);
    Park P;
    for (int K : Ks) {
        long total = 0, mismatches = 0;
        P.parked = 0;
        for (int z = 0; z < d; ++z) for (int x = 0; x < w; ++x)
            for (int dir = +1; dir >= -1; dir -= 2) {
                A.clear(); P.clear();
                for (int step = 0; step < h; ++step) {
                    const int t = dir > 0 ? step : h - 1 - step;
                    const size_t v = (size_t)z * wh + (size_t)t * w + x;
                    bool ok = true;
                    const SepItem *pf = l1f.pool.data() + l1f.off[v], *pb = l1b.pool.data() + l1b.off[v];
                    const int cf = (int)l1f.cnt[v], cb = (int)l1b.cnt[v];
                    long iters = 0;
                    if (K == 0) {
                        const long c0 = g_cnt[0];
                        if (cf | cb | A.n)
                            dir > 0 ? A.step<SepItem, SepUnpackPlain, true, +1>(t, pf, cf, pb, cb, fr, FMAX, ok)
                                    : A.step<SepItem, SepUnpackPlain, true, -1>(t, pf, cf, pb, cb, fr, FMAX, ok);
                        iters = g_cnt[0] - c0;
                    } else {
                        int nf0 = 0;
                        if (cf | cb | A.n)
                            nf0 = dir > 0 ? A.step<SepItem, SepUnpackPlain, true, +1>(t, pf, cf, pb, cb, fr, FMAX, ok)
                                          : A.step<SepItem, SepUnpackPlain, true, -1>(t, pf, cf, pb, cb, fr, FMAX, ok);
                        const int nf1 = P.step(t, dir, pf, cf, pb, cb, fr2, step % K == 0, K, iters);
                        if (nf0 != nf1 || memcmp(fr, fr2, sizeof(SepItem) * nf0)) ++mismatches;
                    }
                    it[(dir > 0 ? 0 : N) + v] = (uint16_t)iters;
                    total += iters;
                }
            }
        // a warp's cost: per position the maximum over its 32 lanes (adjacent x)
        double issued = 0;
        for (int dirsel = 0; dirsel < 2; ++dirsel) {
            const uint16_t *p = it.data() + (dirsel ? N : 0);
            for (int z = 0; z < d; ++z) for (int x0 = 0; x0 + 32 <= w; x0 += 32) for (int t = 0; t < h; ++t) {
                int mx = 0;
                const uint16_t *q = p + (size_t)z * wh + (size_t)t * w + x0;
                for (int l = 0; l < 32; ++l) mx = std::max(mx, (int)q[l]);
                issued += mx;
            }
        }
        // the same with a warp's lanes as an 8 x 4 patch of lines (x by z), the kernels' mapping
        double issued_patch = 0;
        for (int dirsel = 0; dirsel < 2; ++dirsel) {
            const uint16_t *p = it.data() + (dirsel ? N : 0);
            for (int z0 = 0; z0 + 4 <= d; z0 += 4) for (int x0 = 0; x0 + 8 <= w; x0 += 8) for (int t = 0; t < h; ++t) {
                int mx = 0;
                for (int lz = 0; lz < 4; ++lz) for (int lx = 0; lx < 8; ++lx)
                    mx = std::max(mx, (int)p[(size_t)(z0 + lz) * wh + (size_t)t * w + x0 + lx]);
                issued_patch += mx;
            }
        }
        printf("K=%2d: iterations %.3f per voxel and direction, warp iterations per position: %.3f (32 x 1 lines), %.3f (8 x 4 patch), "
               "parked %ld, front mismatches %ld\n", K, total / (2.0 * N), issued / (2.0 * N / 32), issued_patch / (2.0 * N / 32),
               P.parked, mismatches);
        fflush(stdout);
    }
    return 0;
}
