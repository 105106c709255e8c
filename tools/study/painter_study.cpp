// painter_study.cpp -- CPU measurements on the separable painter's per-line code (paint_sep_core.cuh), in the
// geometry the CUDA kernels use (stage 1 along x; stage 2 along y with the 32 lanes of a warp on adjacent x).
// Research tool, not product and not a test: it answers "what would a different lane mapping / retention rule buy"
// before GPU time is spent on it.  Build and run: tools/study/run_study.sh
//
//   input : raw u32 ridge volume (r^2 at ridge points, 0 elsewhere), w h d on the command line
//   output: visit categories, lock-step utilisation of a warp per position, and the same when lanes only
//           re-converge every K positions (a flattened per-lane loop), plus how long a retained-but-dominated
//           item stays dominated (what parking it until its wake-up position would save)
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

static long g_cnt[16];
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
SepItem fr[FMAX];
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s ridge.u32 w h d\n", argv[0]); return 2; }
    const int w = atoi(argv[2]), h = atoi(argv[3]), d = atoi(argv[4]);
    const size_t N = (size_t)w * h * d, wh = (size_t)w * h;
    std::vector<uint32_t> ridge(N);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(ridge.data(), 4, N, f) != N) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(f);
    size_t nridge = 0;
    for (size_t i = 0; i < N; ++i) nridge += ridge[i] != 0;
    printf("volume %d x %d x %d, ridge points %zu (%.1f %% of N)\n", w, h, d, nridge, 100.0 * nridge / N);

    // ---- stage 1 along x, visits per (row, position, direction) ----
    Lists l1f(N), l1b(N);
    std::vector<uint16_t> v1(2 * N);
    memset(g_cnt, 0, sizeof g_cnt);
    for (size_t row = 0; row < (size_t)h * d; ++row)
        for (int dir = +1; dir >= -1; dir -= 2) {
            Lists &out = dir > 0 ? l1f : l1b;
            A.clear();
            for (int step = 0; step < w; ++step) {
                const int t = dir > 0 ? step : w - 1 - step;
                const size_t v = row * w + t;
                bool ok = true;
                SepItem one; one.rho = ridge[v]; one.R = ridge[v];
                const long c0 = g_cnt[0];
                int nf = 0;
                if (ridge[v] || A.n)
                    nf = dir > 0 ? A.step<SepItem, SepUnpackPlain, false, +1>(t, &one, ridge[v] ? 1 : 0, nullptr, 0, fr, FMAX, ok)
                                 : A.step<SepItem, SepUnpackPlain, false, -1>(t, &one, ridge[v] ? 1 : 0, nullptr, 0, fr, FMAX, ok);
                if (!ok || nf > FMAX) { fprintf(stderr, "overflow\n"); return 1; }
                v1[(dir > 0 ? 0 : N) + v] = (uint16_t)(g_cnt[0] - c0);
                out.off[v] = out.pool.size(); out.cnt[v] = (uint32_t)nf;
                out.pool.insert(out.pool.end(), fr, fr + nf);
            }
        }
    printf("\nstage 1 (x): visits %ld = %.2f per voxel and direction; expired %.1f %%, front %.1f %%, dropped (newer) %.1f %%, "
           "dropped (for life) %.1f %%, retained dominated %.1f %%; items out %.2f per voxel\n",
           g_cnt[0], g_cnt[0] / (2.0 * N), 100.0 * g_cnt[1] / g_cnt[0], 100.0 * g_cnt[2] / g_cnt[0], 100.0 * g_cnt[3] / g_cnt[0],
           100.0 * g_cnt[4] / g_cnt[0], 100.0 * g_cnt[5] / g_cnt[0], (l1f.pool.size() + l1b.pool.size()) / (double)N);

    // ---- optional: canonical slack for stage 2 (argv[5] = 1).  What an item (rho, R) still covers in the (y, z) plane
    // is {dy^2 + dz^2 <= rho}: the same lattice points for every rho between two consecutive sums of two squares,
    // so rho may be rounded down to the largest sum of two squares, and an item whose rounded slack does not
    // exceed that of an item with a larger tag is redundant (the 2-D form of stage 2's reach pruning) ----
    if (argc > 5 && atoi(argv[5]) == 1) {
        uint32_t maxr = 0;
        for (size_t i = 0; i < N; ++i) maxr = std::max(maxr, ridge[i]);
        std::vector<uint32_t> canon(maxr + 1, 0);
        std::vector<char> is2(maxr + 1, 0);
        for (uint32_t a = 0; a * a <= maxr; ++a) for (uint32_t b = a; a * a + b * b <= maxr; ++b) is2[a * a + b * b] = 1;
        uint32_t last = 0, nsum = 0;
        for (uint32_t r = 0; r <= maxr; ++r) { if (is2[r]) { last = r; ++nsum; } canon[r] = last; }
        printf("canonical slack: %u of %u values up to max r^2 are sums of two squares\n", nsum, maxr + 1);
        size_t before = 0, after = 0;
        for (Lists *Lp : {&l1f, &l1b}) {
            Lists &Ls = *Lp;
            before += Ls.pool.size();
            std::vector<SepItem> np; np.reserve(Ls.pool.size());
            for (size_t v = 0; v < N; ++v) {
                const size_t o = Ls.off[v]; const uint32_t c = Ls.cnt[v];
                Ls.off[v] = np.size();
                long lastc = -1; uint32_t k2 = 0;
                for (uint32_t k = 0; k < c; ++k) {
                    SepItem it = Ls.pool[o + k];
                    const long cs = canon[it.rho];
                    if (cs > lastc) { lastc = cs; it.rho = (uint32_t)cs; np.push_back(it); ++k2; }
                }
                Ls.cnt[v] = k2;
            }
            Ls.pool.swap(np);
            after += Ls.pool.size();
        }
        printf("stage-1 items %zu -> %zu (%.1f %% fewer)\n", before, after, 100.0 * (before - after) / before);
    }

    // ---- stage 2 along y ----
    Lists l2f(N), l2b(N);
    std::vector<uint16_t> v2(2 * N);
    memset(g_cnt, 0, sizeof g_cnt);
    double sum_active = 0;
    for (int z = 0; z < d; ++z) for (int x = 0; x < w; ++x)
        for (int dir = +1; dir >= -1; dir -= 2) {
            Lists &out = dir > 0 ? l2f : l2b;
            A.clear();
            for (int step = 0; step < h; ++step) {
                const int t = dir > 0 ? step : h - 1 - step;
                const size_t v = (size_t)z * wh + (size_t)t * w + x;
                bool ok = true;
                const SepItem *pf = l1f.pool.data() + l1f.off[v], *pb = l1b.pool.data() + l1b.off[v];
                const int cf = (int)l1f.cnt[v], cb = (int)l1b.cnt[v];
                const long c0 = g_cnt[0];
                int nf = 0;
                if (cf | cb | A.n)
                    nf = dir > 0 ? A.step<SepItem, SepUnpackPlain, true, +1>(t, pf, cf, pb, cb, fr, FMAX, ok)
                                 : A.step<SepItem, SepUnpackPlain, true, -1>(t, pf, cf, pb, cb, fr, FMAX, ok);
                if (!ok || nf > FMAX) { fprintf(stderr, "overflow\n"); return 1; }
                v2[(dir > 0 ? 0 : N) + v] = (uint16_t)(g_cnt[0] - c0);
                sum_active += A.n;
                out.off[v] = out.pool.size(); out.cnt[v] = (uint32_t)nf;
                out.pool.insert(out.pool.end(), fr, fr + nf);
            }
        }
    printf("stage 2 (y): visits %ld = %.2f per voxel and direction; expired %.1f %%, front %.1f %%, dropped (newer) %.1f %%, "
           "dropped (for life) %.1f %%, retained dominated %.1f %%; mean active set %.2f; items out %.2f per voxel\n",
           g_cnt[0], g_cnt[0] / (2.0 * N), 100.0 * g_cnt[1] / g_cnt[0], 100.0 * g_cnt[2] / g_cnt[0], 100.0 * g_cnt[3] / g_cnt[0],
           100.0 * g_cnt[4] / g_cnt[0], 100.0 * g_cnt[5] / g_cnt[0], sum_active / (2.0 * N), (l2f.pool.size() + l2b.pool.size()) / (double)N);

    // ---- lane utilisation: a warp = 32 adjacent lines, lock-step per position (today) or per K positions ----
    // cost model per lane and position: visits + c (the per-position bookkeeping in units of one visit)
    const int Ks[] = {1, 2, 4, 8, 16, 32, 64, 1 << 20};
    for (int stage = 1; stage <= 2; ++stage) {
        const std::vector<uint16_t> &vv = stage == 1 ? v1 : v2;
        // stage 1: line = row (y, z), lanes along y (adjacent rows), position x.  stage 2: line = (x, z), lanes along x, position y.
        const int L = stage == 1 ? w : h;
        for (double c : {0.0, 0.5, 1.0}) {
            printf("stage %d, per-position overhead %.1f visit-equivalents: useful / issued lane slots when lanes re-converge every K positions\n   ", stage, c);
            for (int K : Ks) {
                double useful = 0, issued = 0;
                for (int dirsel = 0; dirsel < 2; ++dirsel) {
                    const uint16_t *p = vv.data() + (dirsel ? N : 0);
                    const size_t nl = stage == 1 ? (size_t)h * d : (size_t)w * d;      // lines
                    for (size_t l0 = 0; l0 + 32 <= nl; l0 += 32) {
                        for (int t0 = 0; t0 < L; t0 += K) {
                            const int t1 = std::min(L, t0 + (K > L ? L : K));
                            double mx = 0;
                            for (int lane = 0; lane < 32; ++lane) {
                                const size_t line = l0 + lane;
                                double s = 0;
                                for (int t = t0; t < t1; ++t) {
                                    size_t v;
                                    if (stage == 1) v = line * w + t;
                                    else { const size_t z = line / w, x = line % w; v = z * wh + (size_t)t * w + x; }
                                    s += p[v] + c;
                                }
                                useful += s; mx = std::max(mx, s);
                            }
                            issued += 32 * mx;
                        }
                    }
                }
                if (K > L) printf(" line:%.3f", useful / issued); else printf(" K=%d:%.3f", K, useful / issued);
            }
            printf("\n");
        }
    }
    // ---- lane shape: a warp's 32 lines as a patch of LX adjacent x by LZ adjacent z (stage 2), lock-step per position ----
    printf("stage 2, warp = LX x LZ patch of lines (x by z), lock-step per position: useful / issued lane slots, warp iterations per position\n   ");
    for (int LX : {32, 16, 8, 4, 2, 1}) {
        const int LZ = 32 / LX;
        double useful = 0, issued = 0, npos = 0;
        for (int dirsel = 0; dirsel < 2; ++dirsel) {
            const uint16_t *p = v2.data() + (dirsel ? N : 0);
            for (int z0 = 0; z0 + LZ <= d; z0 += LZ) for (int x0 = 0; x0 + LX <= w; x0 += LX) for (int t = 0; t < h; ++t) {
                int mx = 0;
                for (int lz = 0; lz < LZ; ++lz) for (int lx = 0; lx < LX; ++lx) {
                    const int c = p[(size_t)(z0 + lz) * wh + (size_t)t * w + x0 + lx];
                    useful += c; mx = std::max(mx, c);
                }
                issued += 32.0 * mx; npos += 1;
            }
        }
        printf(" %dx%d: %.3f (%.2f)", LX, LZ, useful / issued, issued / 32.0 / npos);
    }
    printf("\n");
    // stage 1: line = row (y, z); today's warp = 32 adjacent y.  stage 3: line = column (x, y), position z, cost taken as
    // arrivals + 1 per position (the staircase updates are proportional to the arrivals)
    printf("stage 1, warp = LY x LZ patch of rows: useful / issued (warp iterations per position)\n   ");
    for (int LY : {32, 16, 8, 4, 2, 1}) {
        const int LZ = 32 / LY;
        double useful = 0, issued = 0, npos = 0;
        for (int dirsel = 0; dirsel < 2; ++dirsel) {
            const uint16_t *p = v1.data() + (dirsel ? N : 0);
            for (int z0 = 0; z0 + LZ <= d; z0 += LZ) for (int y0 = 0; y0 + LY <= h; y0 += LY) for (int t = 0; t < w; ++t) {
                int mx = 0;
                for (int lz = 0; lz < LZ; ++lz) for (int ly = 0; ly < LY; ++ly) {
                    const int c = p[(size_t)(z0 + lz) * wh + (size_t)(y0 + ly) * w + t];
                    useful += c; mx = std::max(mx, c);
                }
                issued += 32.0 * mx; npos += 1;
            }
        }
        printf(" %dx%d: %.3f (%.2f)", LY, LZ, useful / issued, issued / 32.0 / npos);
    }
    printf("\nstage 3, warp = LX x LY patch of columns: useful / issued (warp iterations per position)\n   ");
    for (int LX : {32, 16, 8, 4, 2, 1}) {
        const int LY = 32 / LX;
        double useful = 0, issued = 0, npos = 0;
        for (int y0 = 0; y0 + LY <= h; y0 += LY) for (int x0 = 0; x0 + LX <= w; x0 += LX) for (int z = 0; z < d; ++z) {
            int mx = 0;
            for (int ly = 0; ly < LY; ++ly) for (int lx = 0; lx < LX; ++lx) {
                const size_t v = (size_t)z * wh + (size_t)(y0 + ly) * w + x0 + lx;
                const int c = (int)(l2f.cnt[v] + l2b.cnt[v]) + 1;
                useful += c; mx = std::max(mx, c);
            }
            issued += 32.0 * mx; npos += 1;
        }
        printf(" %dx%d: %.3f (%.2f)", LX, LY, useful / issued, issued / 32.0 / npos);
    }
    printf("\n");
    // ---- stage 3: what the staircase operations cost (loop trips inside arrive() / query()) ----
    {
        memset(g_cnt, 0, sizeof g_cnt);
        static SepStair<SepStairLocal<CAP> > S;
        std::vector<uint32_t> outv(N, 0);
        std::vector<uint16_t> c3(2 * N);
        double sum_n = 0; long max_n = 0;
        for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x)
            for (int dir = +1; dir >= -1; dir -= 2) {
                S.clear();
                for (int step = 0; step < d; ++step) {
                    const int z = dir > 0 ? step : d - 1 - step;
                    const int u = dir > 0 ? z : -z;
                    const size_t v = (size_t)z * wh + (size_t)y * w + x;
                    long c0 = 0; for (int k = 6; k <= 12; ++k) c0 += g_cnt[k];
                    for (int which = 0; which < 2; ++which) {
                        const Lists &in = which ? l2b : l2f;
                        int hint = 0;
                        for (uint32_t k = 0; k < in.cnt[v]; ++k) { const SepItem it = in.pool[in.off[v] + k]; S.arrive(u, it.rho, it.R, hint); }
                    }
                    const uint32_t r = S.query(u);
                    if (dir > 0) outv[v] = r; else if (r > outv[v]) outv[v] = r;
                    long c1 = 0; for (int k = 6; k <= 12; ++k) c1 += g_cnt[k];
                    c3[(dir > 0 ? 0 : N) + v] = (uint16_t)(c1 - c0 + 1);
                    sum_n += S.n; max_n = std::max<long>(max_n, S.n);
                }
            }
        const double per = 2.0 * N;
        printf("stage 3 (z), per voxel and direction: arrivals %.2f (head rejects %.2f, other rejects %.2f), search steps %.2f, "
               "removal scan %.2f, shift moves on insert %.2f, shift moves on query %.2f; staircase size mean %.2f max %ld\n",
               g_cnt[6] / per, g_cnt[7] / per, g_cnt[9] / per, g_cnt[8] / per, g_cnt[10] / per, g_cnt[11] / per, g_cnt[12] / per,
               sum_n / per, max_n);
        printf("stage 3 with cost = all loop trips + 1: warp = LX x LY patch: useful / issued (warp cost per position)\n   ");
        for (int LX : {32, 16, 8, 4}) {
            const int LY = 32 / LX;
            double useful = 0, issued = 0, npos = 0;
            for (int dirsel = 0; dirsel < 2; ++dirsel) {
                const uint16_t *p = c3.data() + (dirsel ? N : 0);
                for (int y0 = 0; y0 + LY <= h; y0 += LY) for (int x0 = 0; x0 + LX <= w; x0 += LX) for (int z = 0; z < d; ++z) {
                    int mx = 0;
                    for (int ly = 0; ly < LY; ++ly) for (int lx = 0; lx < LX; ++lx) {
                        const int c = p[(size_t)z * wh + (size_t)(y0 + ly) * w + x0 + lx];
                        useful += c; mx = std::max(mx, c);
                    }
                    issued += 32.0 * mx; npos += 1;
                }
            }
            printf(" %dx%d: %.3f (%.2f)", LX, LY, useful / issued, issued / 32.0 / npos);
        }
        printf("\n");
    }
    return 0;
}
