/*
 * lt_oracle.c -- CPU restatement of the BoneJ Thickness hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle and the timed CPU baseline.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load it.  The product path (bonej2_b200/) never links, imports or calls anything in oracle/.
 *
 * What it restates.  BoneJ's ThicknessWrapper
 * (Modern/wrapperPlugins/src/main/java/org/bonej/wrapperPlugins/ThicknessWrapper.java:123-222)
 * and ThicknessHelper (Legacy/bonej/src/main/java/org/bonej/menuWrappers/ThicknessHelper.java:106-117)
 * delegate all arithmetic to the third-party artifact sc.fiji:LocalThickness_ (package
 * sc.fiji.localThickness; declared without a version at Modern/wrapperPlugins/pom.xml:177-180,
 * version managed by the BOM org.scijava:pom-scijava:45.0.0 at pom.xml:6-8 -- the BOM is not on
 * disk, the number is believed to be 5.0.0) and to ij.process.StackStatistics (net.imagej:ij).
 * Neither artifact is vendored under /root/reference and there is no JVM in the image, so the
 * functions below restate the published algorithm of those classes:
 *   lto_edt            <- EDT_S1D (Saito-Toriwaki, Step1/2/3 threads + sqrt pass)      [A.1]
 *   lto_ridge          <- Distance_Ridge.run / createTemplate / scanCube               [A.2]
 *   lto_paint          <- Local_Thickness_Parallel.run + LTThread                      [A.3]
 *   lto_cleanup        <- Clean_Up_Local_Thickness.run / setFlag / averageInteriorNeighbors [A.4]
 *   lto_mask           <- MaskThicknessMapWithOriginal.trimOverhang                    [A.5]
 *   lto_calibrate      <- LocalThicknessWrapper calibration tail (0->NaN, x pixelWidth)[A.6]
 *   lto_stats          <- ij.process.StackStatistics (32-bit stack) as consumed at
 *                         ThicknessWrapper.java:158-180                                [A.7]
 *   lto_process_image  <- LocalThicknessWrapper.processImage chain                     [A.0]
 *
 * Pinning.  The restatement is pinned by the reference's own known-answer tests:
 *   ThicknessWrapperTest.java:225-259  (2x2x2 zeros -> NaN,NaN,NaN, 10.392304420471191, 0, 10.39...)
 *   ThicknessHelperTest.java:45-78     (rods +-1.5, spheres +-10 %, bricks +-5 %)
 * see tests/test_oracle_golden.py.  Per-voxel EDT / ridge / map values are NOT pinned by any
 * reference test ("parity unpinned" at that granularity) -- only this restatement pins them.
 *
 * The algorithmic structure (brute-force line scans in EDT steps 2/3, slice-strided threads,
 * locked scatter-max painting, separate passes) is kept on purpose: it is what the JVM
 * reference executes, so timing this file is the closest available CPU baseline.
 *
 * Layout everywhere: contiguous [z][y][x], x fastest (the Java code holds one array per z slice).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LTO_MAX_THREADS 256

typedef struct {
    int64_t count;      /* longPixelCount: voxels that are not NaN                  */
    double mean;        /* sum / n                                                   */
    double sd;          /* sqrt((n*sum2 - sum*sum)/n/(n-1)), 0 when the numerator<=0 */
    double min, max;    /* over non-NaN voxels                                       */
    double sum, sum2;
} lto_stats_t;

/* ------------------------------------------------------------------------------------------ */
/* tiny fork/join helper: the Java code spawns nThreads plain Threads per stage and joins them */
typedef struct job {
    void (*fn)(struct job *, int tid, int nthreads);
    const void *a0, *a1;
    void *o0, *o1;
    int w, h, d, inverse, thresh, noResult;
    pthread_mutex_t *locks;
} job_t;

typedef struct { job_t *job; int tid, n; } targ_t;

static void *trampoline(void *p) {
    targ_t *t = (targ_t *)p;
    t->job->fn(t->job, t->tid, t->n);
    return NULL;
}

static void run_threads(job_t *job, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > LTO_MAX_THREADS) nthreads = LTO_MAX_THREADS;
    pthread_t th[LTO_MAX_THREADS];
    targ_t ta[LTO_MAX_THREADS];
    for (int i = 0; i < nthreads; i++) {
        ta[i].job = job; ta[i].tid = i; ta[i].n = nthreads;
        pthread_create(&th[i], NULL, trampoline, &ta[i]);
    }
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
}

static inline int is_bg(uint8_t px, int thresh, int inverse) {
    return ((px & 255) < thresh) ^ (inverse != 0);
}

/* ------------------------------------------------------------------------------------------ */
/* A.1  EDT_S1D.  s holds integers carried in float, exactly as the Java float[][] stack does. */

static void edt_step1(job_t *jb, int tid, int nt) {
    const uint8_t *data = (const uint8_t *)jb->a0;
    float *s = (float *)jb->o0;
    const int w = jb->w, h = jb->h, d = jb->d, n0 = jb->noResult;
    uint8_t *background = (uint8_t *)malloc((size_t)w);
    for (int k = tid; k < d; k += nt) {            /* threads stride z */
        const uint8_t *dk = data + (size_t)k * w * h;
        float *sk = s + (size_t)k * w * h;
        for (int j = 0; j < h; j++) {
            for (int i = 0; i < w; i++) background[i] = (uint8_t)is_bg(dk[i + (size_t)w * j], jb->thresh, jb->inverse);
            for (int i = 0; i < w; i++) {
                int min = n0;
                for (int x = i; x < w; x++) {          /* scan right */
                    if (background[x]) { int t = i - x; min = t * t; break; }
                }
                for (int x = i - 1; x >= 0; x--) {     /* then left */
                    if (background[x]) { int t = i - x; t *= t; if (t < min) min = t; break; }
                }
                sk[i + (size_t)w * j] = (float)min;
            }
        }
    }
    free(background);
}

static void edt_step2(job_t *jb, int tid, int nt) {
    float *s = (float *)jb->o0;
    const int w = jb->w, h = jb->h, d = jb->d, n0 = jb->noResult;
    int *tempInt = (int *)malloc(sizeof(int) * (size_t)h);
    int *tempS = (int *)malloc(sizeof(int) * (size_t)h);
    for (int k = tid; k < d; k += nt) {            /* threads stride z */
        float *sk = s + (size_t)k * w * h;
        for (int i = 0; i < w; i++) {
            int nonempty = 0;
            for (int j = 0; j < h; j++) {
                tempS[j] = tempInt[j] = (int)sk[i + (size_t)w * j];
                if (tempS[j] > 0) nonempty = 1;
            }
            if (!nonempty) continue;                /* all-zero columns are skipped */
            for (int j = 0; j < h; j++) {
                int min = n0, delta = j;
                for (int y = 0; y < h; y++) {       /* brute-force line scan, as upstream */
                    int test = tempInt[y] + delta * delta;
                    delta--;
                    if (test < min) min = test;
                }
                tempS[j] = min;
            }
            for (int j = 0; j < h; j++) sk[i + (size_t)w * j] = (float)tempS[j];
        }
    }
    free(tempInt); free(tempS);
}

static void edt_step3(job_t *jb, int tid, int nt) {
    const uint8_t *data = (const uint8_t *)jb->a0;
    float *s = (float *)jb->o0;
    const int w = jb->w, h = jb->h, d = jb->d, n0 = jb->noResult;
    const size_t wh = (size_t)w * h;
    int *tempInt = (int *)malloc(sizeof(int) * (size_t)d);
    int *tempS = (int *)malloc(sizeof(int) * (size_t)d);
    for (int j = tid; j < h; j += nt) {            /* threads stride y */
        for (int i = 0; i < w; i++) {
            const size_t ind = i + (size_t)w * j;
            int nonempty = 0;
            for (int k = 0; k < d; k++) {
                tempS[k] = tempInt[k] = (int)s[k * wh + ind];
                if (tempS[k] > 0) nonempty = 1;
            }
            if (!nonempty) continue;
            int zStart = 0;
            while (zStart < d - 1 && tempInt[zStart] == 0) zStart++;
            if (zStart > 0) zStart--;
            int zStop = d - 1;
            while (zStop > 0 && tempInt[zStop] == 0) zStop--;
            if (zStop < d - 1) zStop++;
            for (int k = 0; k < d; k++) {
                if (!is_bg(data[k * wh + ind], jb->thresh, jb->inverse)) {   /* foreground z only */
                    int min = n0;
                    int zBegin = zStart, zEnd = zStop;
                    if (zBegin > k) zBegin = k;
                    if (zEnd < k) zEnd = k;
                    int delta = k - zBegin;
                    for (int z = zBegin; z <= zEnd; z++) {
                        int test = tempInt[z] + delta * delta;
                        delta--;
                        if (test < min) min = test;
                    }
                    tempS[k] = min;
                }
            }
            for (int k = 0; k < d; k++) s[k * wh + ind] = (float)tempS[k];
        }
    }
    free(tempInt); free(tempS);
}

static void edt_final(job_t *jb, int tid, int nt) {
    const uint8_t *data = (const uint8_t *)jb->a0;
    float *s = (float *)jb->o0;
    uint32_t *sq = (uint32_t *)jb->o1;
    const size_t wh = (size_t)jb->w * jb->h;
    for (int k = tid; k < jb->d; k += nt) {
        for (size_t ind = 0; ind < wh; ind++) {
            const size_t p = k * wh + ind;
            if (is_bg(data[p], jb->thresh, jb->inverse)) {
                s[p] = 0.0f;
                if (sq) sq[p] = 0;
            } else {
                if (sq) sq[p] = (uint32_t)(int)s[p];
                s[p] = (float)sqrt((double)s[p]);
            }
        }
    }
}

/* out_dist: float distance map (bg -> 0).  out_sq (optional): the integer squared map the
 * Java code carries in the float stack just before the sqrt pass (bg -> 0). */
int lto_edt(const uint8_t *in, int w, int h, int d, int inverse, int thresh,
            float *out_dist, uint32_t *out_sq, int nthreads) {
    int n = w; if (h > n) n = h; if (d > n) n = d;
    job_t jb; memset(&jb, 0, sizeof jb);
    jb.a0 = in; jb.o0 = out_dist; jb.o1 = out_sq;
    jb.w = w; jb.h = h; jb.d = d; jb.inverse = inverse; jb.thresh = thresh;
    jb.noResult = 3 * (n + 1) * (n + 1);
    jb.fn = edt_step1; run_threads(&jb, nthreads);
    jb.fn = edt_step2; run_threads(&jb, nthreads);
    jb.fn = edt_step3; run_threads(&jb, nthreads);
    jb.fn = edt_final; run_threads(&jb, nthreads);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A.2  Distance_Ridge */

static inline int sq_of(float v) { return (int)(v * v + 0.5f); }   /* float arithmetic, as upstream */

/* scanCube: smallest r1^2 such that the grid ball of radius r1 centred at (dx,dy,dz) contains the
 * grid ball of radius^2 rSq centred at the origin. */
static void scan_cube(int dx, int dy, int dz, const int *distSqValues, int numRadii, int *r1Sq) {
    const int dxAbs = -abs(dx), dyAbs = -abs(dy), dzAbs = -abs(dz);
    for (int ri = 0; ri < numRadii; ri++) {
        const int rSq = distSqValues[ri];
        int max = 0;
        const int r = 1 + (int)sqrt((double)rSq);
        for (int k = 0; k <= r; k++) {
            const int scank = k * k;
            const int dk = (k - dzAbs) * (k - dzAbs);
            for (int j = 0; j <= r; j++) {
                const int scankj = scank + j * j;
                if (scankj <= rSq) {
                    const int iPlus = ((int)sqrt((double)(rSq - scankj))) - dxAbs;
                    const int dkji = dk + (j - dyAbs) * (j - dyAbs) + iPlus * iPlus;
                    if (dkji > max) max = dkji;
                }
            }
        }
        r1Sq[ri] = max;
    }
}

/* exported so tests can pin the GPU threshold table against it */
int lto_ridge_template(const int *distSqValues, int numRadii, int *t1, int *t2, int *t3) {
    scan_cube(1, 0, 0, distSqValues, numRadii, t1);
    scan_cube(1, 1, 0, distSqValues, numRadii, t2);
    scan_cube(1, 1, 1, distSqValues, numRadii, t3);
    return 0;
}

/* dist: float distance map (A.1 output); ridge: float, distance at ridge points else 0. */
int lto_ridge(const float *dist, int w, int h, int d, float *ridge) {
    const size_t wh = (size_t)w * h, N = wh * d;
    float distMax = 0;
    for (size_t p = 0; p < N; p++) if (dist[p] > distMax) distMax = dist[p];
    const int rSqMax = (int)(distMax * distMax + 0.5f) + 1;
    uint8_t *occurs = (uint8_t *)calloc((size_t)rSqMax, 1);
    for (size_t p = 0; p < N; p++) occurs[sq_of(dist[p])] = 1;
    int numRadii = 0;
    for (int i = 0; i < rSqMax; i++) if (occurs[i]) numRadii++;
    int *distSqIndex = (int *)calloc((size_t)rSqMax, sizeof(int));
    int *distSqValues = (int *)malloc(sizeof(int) * (size_t)numRadii);
    int indDS = 0;
    for (int i = 0; i < rSqMax; i++) if (occurs[i]) { distSqIndex[i] = indDS; distSqValues[indDS++] = i; }
    int *tmpl[3];
    for (int c = 0; c < 3; c++) tmpl[c] = (int *)malloc(sizeof(int) * (size_t)numRadii);
    lto_ridge_template(distSqValues, numRadii, tmpl[0], tmpl[1], tmpl[2]);

    memset(ridge, 0, sizeof(float) * N);
    for (int k = 0; k < d; k++) {
        const float *sk = dist + k * wh;
        for (int j = 0; j < h; j++) for (int i = 0; i < w; i++) {
            const size_t ind = i + (size_t)w * j;
            if (!(sk[ind] > 0)) continue;
            int notRidgePoint = 0;
            const int sk0SqInd = distSqIndex[sq_of(sk[ind])];
            for (int dz = -1; dz <= 1 && !notRidgePoint; dz++) {
                const int k1 = k + dz;
                if (k1 < 0 || k1 >= d) continue;
                const float *sk1 = dist + k1 * wh;
                for (int dy = -1; dy <= 1 && !notRidgePoint; dy++) {
                    const int j1 = j + dy;
                    if (j1 < 0 || j1 >= h) continue;
                    for (int dx = -1; dx <= 1; dx++) {
                        const int i1 = i + dx;
                        if (i1 < 0 || i1 >= w) continue;
                        if (dx == 0 && dy == 0 && dz == 0) continue;
                        const int sk1Sq = sq_of(sk1[i1 + (size_t)w * j1]);
                        if (sk1Sq >= tmpl[abs(dz) + abs(dy) + abs(dx) - 1][sk0SqInd]) { notRidgePoint = 1; break; }
                    }
                }
            }
            if (!notRidgePoint) ridge[k * wh + ind] = sk[ind];
        }
    }
    free(occurs); free(distSqIndex); free(distSqValues);
    for (int c = 0; c < 3; c++) free(tmpl[c]);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A.3  Local_Thickness_Parallel: scatter-max of integer r^2 under a per-destination-slice lock */

typedef struct { int *i, *j; float *r; int n; } ridge_slice_t;

static void lt_paint_thread(job_t *jb, int tid, int nt) {
    const ridge_slice_t *rs = (const ridge_slice_t *)jb->a0;
    float *s = (float *)jb->o0;
    const int w = jb->w, h = jb->h, d = jb->d;
    const size_t wh = (size_t)w * h;
    for (int k = tid; k < d; k += nt) {            /* threads stride the SOURCE slice */
        const int nR = rs[k].n;
        for (int iR = 0; iR < nR; iR++) {
            const int i = rs[k].i[iR], j = rs[k].j[iR];
            const float r = rs[k].r[iR];
            const int rSquared = (int)(r * r + 0.5f);
            int rInt = (int)r;
            if (rInt < r) rInt++;
            int iStart = i - rInt; if (iStart < 0) iStart = 0;
            int iStop = i + rInt;  if (iStop >= w) iStop = w - 1;
            int jStart = j - rInt; if (jStart < 0) jStart = 0;
            int jStop = j + rInt;  if (jStop >= h) jStop = h - 1;
            int kStart = k - rInt; if (kStart < 0) kStart = 0;
            int kStop = k + rInt;  if (kStop >= d) kStop = d - 1;
            for (int k1 = kStart; k1 <= kStop; k1++) {
                const int r1SquaredK = (k1 - k) * (k1 - k);
                float *sk1 = s + k1 * wh;
                for (int j1 = jStart; j1 <= jStop; j1++) {
                    const int r1SquaredJK = r1SquaredK + (j1 - j) * (j1 - j);
                    if (r1SquaredJK > rSquared) continue;
                    for (int i1 = iStart; i1 <= iStop; i1++) {
                        const int r1Squared = r1SquaredJK + (i1 - i) * (i1 - i);
                        if (r1Squared <= rSquared) {
                            const size_t ind1 = i1 + (size_t)w * j1;
                            const float s1 = (float)rSquared;
                            /* unsynchronised peek then locked re-check, as upstream */
                            if (s1 > *(volatile float *)&sk1[ind1]) {
                                pthread_mutex_lock(&jb->locks[k1]);
                                if (s1 > sk1[ind1]) sk1[ind1] = s1;
                                pthread_mutex_unlock(&jb->locks[k1]);
                            }
                        }
                    }
                }
            }
        }
    }
}

/* ridge: float ridge map (A.2 output).  out: (float)(2*sqrt(max r^2)) per voxel.
 * out_rsq (optional): the painted integer r^2 before the final sqrt. */
int lto_paint(const float *ridge, int w, int h, int d, float *out, uint32_t *out_rsq, int nthreads) {
    const size_t wh = (size_t)w * h, N = wh * d;
    ridge_slice_t *rs = (ridge_slice_t *)calloc((size_t)d, sizeof(ridge_slice_t));
    for (int k = 0; k < d; k++) {
        const float *sk = ridge + k * wh;
        int nr = 0;
        for (size_t ind = 0; ind < wh; ind++) if (sk[ind] > 0) nr++;
        rs[k].n = nr;
        rs[k].i = (int *)malloc(sizeof(int) * (size_t)(nr ? nr : 1));
        rs[k].j = (int *)malloc(sizeof(int) * (size_t)(nr ? nr : 1));
        rs[k].r = (float *)malloc(sizeof(float) * (size_t)(nr ? nr : 1));
        int iR = 0;
        for (int j = 0; j < h; j++) for (int i = 0; i < w; i++) {
            const size_t ind = i + (size_t)w * j;
            if (sk[ind] > 0) { rs[k].i[iR] = i; rs[k].j[iR] = j; rs[k].r[iR] = sk[ind]; iR++; }
        }
    }
    memset(out, 0, sizeof(float) * N);             /* working copy cleared */
    pthread_mutex_t *locks = (pthread_mutex_t *)malloc(sizeof(pthread_mutex_t) * (size_t)d);
    for (int k = 0; k < d; k++) pthread_mutex_init(&locks[k], NULL);
    job_t jb; memset(&jb, 0, sizeof jb);
    jb.a0 = rs; jb.o0 = out; jb.w = w; jb.h = h; jb.d = d; jb.locks = locks;
    jb.fn = lt_paint_thread; run_threads(&jb, nthreads);
    for (int k = 0; k < d; k++) pthread_mutex_destroy(&locks[k]);
    free(locks);
    for (size_t p = 0; p < N; p++) {
        if (out_rsq) out_rsq[p] = (uint32_t)(int)out[p];
        out[p] = (float)(2 * sqrt((double)out[p]));
    }
    for (int k = 0; k < d; k++) { free(rs[k].i); free(rs[k].j); free(rs[k].r); }
    free(rs);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A.4  Clean_Up_Local_Thickness */

static inline float look(const float *s, int i, int j, int k, int w, int h, int d) {
    if (i < 0 || i >= w || j < 0 || j >= h || k < 0 || k >= d) return -1.0f;   /* out of image != 0 */
    return s[(size_t)k * w * h + i + (size_t)w * j];
}

/* visiting order: 6 faces, 12 edges, 8 corners (affects only the last ulp of the float32 sum) */
static const int8_t NB26[26][3] = {
    {0,0,-1},{0,0,1},{0,-1,0},{0,1,0},{-1,0,0},{1,0,0},
    {0,1,-1},{0,1,1},{1,-1,0},{1,1,0},{-1,0,1},{1,0,1},
    {0,-1,-1},{0,-1,1},{-1,-1,0},{-1,1,0},{-1,0,-1},{1,0,-1},
    {1,1,1},{1,-1,1},{-1,1,1},{-1,-1,1},
    {1,1,-1},{1,-1,-1},{-1,1,-1},{-1,-1,-1}
};

static float set_flag(const float *s, int i, int j, int k, int w, int h, int d) {
    if (s[(size_t)k * w * h + i + (size_t)w * j] == 0) return 0;
    for (int n = 0; n < 26; n++)
        if (look(s, i + NB26[n][0], j + NB26[n][1], k + NB26[n][2], w, h, d) == 0) return -1;
    return s[(size_t)k * w * h + i + (size_t)w * j];
}

static float average_interior_neighbors(const float *s, const float *sNew, int i, int j, int k, int w, int h, int d) {
    int n = 0;
    float sum = 0;
    for (int q = 0; q < 26; q++) {
        const float value = look(sNew, i + NB26[q][0], j + NB26[q][1], k + NB26[q][2], w, h, d);
        if (value > 0) { n++; sum += value; }
    }
    if (n > 0) return sum / n;
    return s[(size_t)k * w * h + i + (size_t)w * j];
}

int lto_cleanup(const float *s, int w, int h, int d, float *sNew) {
    const size_t wh = (size_t)w * h;
    for (int k = 0; k < d; k++) for (int j = 0; j < h; j++) for (int i = 0; i < w; i++)
        sNew[k * wh + i + (size_t)w * j] = set_flag(s, i, j, k, w, h, d);
    for (int k = 0; k < d; k++) for (int j = 0; j < h; j++) for (int i = 0; i < w; i++) {
        const size_t ind = k * wh + i + (size_t)w * j;
        if (sNew[ind] == -1) sNew[ind] = -average_interior_neighbors(s, sNew, i, j, k, w, h, d);
    }
    for (size_t p = 0; p < wh * d; p++) sNew[p] = fabsf(sNew[p]);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A.5  MaskThicknessMapWithOriginal.trimOverhang */
int lto_mask(const uint8_t *original, size_t N, int inverse, float *map) {
    const int fg = inverse ? 0 : 255;
    for (size_t p = 0; p < N; p++) if ((original[p] & 255) != fg) map[p] = 0;
    return 0;
}

/* A.6  calibration tail: 0 -> NaN, then FloatProcessor.multiply(pixelWidth) in float */
int lto_calibrate(float *map, size_t N, double pixelWidth) {
    const float pw = (float)pixelWidth;
    for (size_t p = 0; p < N; p++) {
        if (map[p] == 0.0f) map[p] = NAN;
        map[p] = map[p] * pw;
    }
    return 0;
}

/* A.7  StackStatistics on a 32-bit stack: two passes, doubles, z->y->x order, NaN excluded */
int lto_stats(const float *map, size_t N, lto_stats_t *st) {
    double mn = INFINITY, mx = -INFINITY;
    for (size_t p = 0; p < N; p++) {
        const double v = map[p];
        if (v < mn) mn = v;        /* NaN fails both comparisons */
        if (v > mx) mx = v;
    }
    int64_t n = 0; double sum = 0, sum2 = 0;
    for (size_t p = 0; p < N; p++) {
        const double v = map[p];
        if (v >= mn && v <= mx) { n++; sum += v; sum2 += v * v; }
    }
    st->count = n; st->sum = sum; st->sum2 = sum2; st->min = mn; st->max = mx;
    st->mean = sum / (double)n;
    double sd = 0;
    if (n > 0) {
        sd = ((double)n * sum2 - sum * sum) / (double)n;
        sd = sd > 0.0 ? sqrt(sd / ((double)n - 1.0)) : 0.0;
    }
    st->sd = sd;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A.0  LocalThicknessWrapper.processImage.  Returns 0, or -1 when the image is not 0/255 binary
 * (upstream: error, returns null).  out_map receives the calibrated map (NaN background).
 * Optional taps (may be NULL): sq (uint32 squared EDT), ridge (float), rsq (painted r^2).     */
int lto_process_image(const uint8_t *in, int w, int h, int d, int inverse, int mask,
                      int thresh, double pixelWidth, int calibrate, float *out_map,
                      uint32_t *tap_sq, float *tap_ridge, uint32_t *tap_rsq,
                      lto_stats_t *st, int nthreads) {
    const size_t N = (size_t)w * h * d;
    for (size_t p = 0; p < N; p++) if (in[p] != 0 && in[p] != 255) return -1;
    float *a = (float *)malloc(sizeof(float) * N);
    float *b = tap_ridge ? tap_ridge : (float *)malloc(sizeof(float) * N);
    lto_edt(in, w, h, d, inverse, thresh, a, tap_sq, nthreads);
    lto_ridge(a, w, h, d, b);
    lto_paint(b, w, h, d, a, tap_rsq, nthreads);
    lto_cleanup(a, w, h, d, out_map);
    if (mask) lto_mask(in, N, inverse, out_map);
    if (calibrate) lto_calibrate(out_map, N, pixelWidth);
    if (st) lto_stats(out_map, N, st);
    free(a);
    if (!tap_ridge) free(b);
    return 0;
}
