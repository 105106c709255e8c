/*
 * edt_exact.c -- TEST INFRASTRUCTURE: a second, independent checker for the squared Euclidean distance map at the
 * sizes the benchmark runs (512^3, 1024^3), where the brute-force restatement of EDT_S1D in lt_oracle.c (Theta(N n))
 * takes hours.  Not a restatement of the reference's loops: the lower envelope of parabolas (Felzenszwalb &
 * Huttenlocher, "Distance Transforms of Sampled Functions", 2012) computes the same min-plus line transform
 *     h(t) = min_j f(j) + (t - j)^2
 * in O(n) per line, in exact integer / double arithmetic (all values < 2^33).  What it must reproduce is
 * EDT_S1D's result as restated in SURVEY.md A.1 / lt_oracle.c:lto_edt: sq(p) = min(|p - q|^2 over background q,
 * 3 (n + 1)^2), background -> 0.  tests/test_oracle_golden.py pins it against lto_edt (and SciPy) on small volumes
 * before tests/test_gpu_configs.py uses it on the large ones.  Only tests/ loads it.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int is_bg(uint8_t px, int thresh, int inverse) { return ((px & 255) < thresh) ^ (inverse != 0); }

/* h(t) = min(min_j f[j] + (t - j)^2, cap) for t in [0, n); v, z: scratch of n and n + 1 entries */
static void envelope(const int64_t *f, int n, int64_t cap, int64_t *out, int *v, double *z) {
    int k = 0;
    v[0] = 0;
    z[0] = -1e300; z[1] = 1e300;
    for (int q = 1; q < n; q++) {
        double s;
        for (;;) {
            const int p = v[k];
            /* abscissa where the parabola rooted at q drops below the one rooted at p (exact numerator, one rounding) */
            s = (double)((f[q] + (int64_t)q * q) - (f[p] + (int64_t)p * p)) / (double)(2 * (q - p));
            if (s <= z[k] && k > 0) { k--; continue; }
            if (s <= z[k]) { k = -1; }
            break;
        }
        if (k < 0) { k = 0; v[0] = q; z[0] = -1e300; z[1] = 1e300; continue; }
        k++;
        v[k] = q; z[k] = s; z[k + 1] = 1e300;
    }
    k = 0;
    for (int t = 0; t < n; t++) {
        while (z[k + 1] < (double)t) k++;
        /* a tie or a rounding at an abscissa that is (nearly) an integer: the neighbours give the same or a larger
         * value, so taking the minimum over the adjacent envelope pieces makes the result independent of it */
        int64_t best = f[v[k]] + (int64_t)(t - v[k]) * (t - v[k]);
        if (k > 0) { const int64_t a = f[v[k - 1]] + (int64_t)(t - v[k - 1]) * (t - v[k - 1]); if (a < best) best = a; }
        if (z[k + 1] < 1e299) { const int64_t a = f[v[k + 1]] + (int64_t)(t - v[k + 1]) * (t - v[k + 1]); if (a < best) best = a; }
        out[t] = best < cap ? best : cap;
    }
}

typedef struct {
    const uint8_t *in; uint32_t *out; int w, h, d, inverse, thresh, pass; int64_t cap;
    long long next, njobs;                      /* jobs are handed out under the mutex, a few at a time */
    pthread_mutex_t mu;
} job_t;

static int take(job_t *jb, long long chunk, long long *lo, long long *hi) {
    pthread_mutex_lock(&jb->mu);
    *lo = jb->next;
    *hi = *lo + chunk < jb->njobs ? *lo + chunk : jb->njobs;
    jb->next = *hi;
    pthread_mutex_unlock(&jb->mu);
    return *lo < *hi;
}

/* x: distance to the nearest background voxel of the row, two scans */
static void *x_worker(void *arg) {
    job_t *jb = (job_t *)arg;
    const int w = jb->w;
    const int64_t cap = jb->cap;
    long long lo, hi;
    while (take(jb, 256, &lo, &hi))
        for (long long row = lo; row < hi; row++) {
            const uint8_t *src = jb->in + (size_t)row * w;
            uint32_t *dst = jb->out + (size_t)row * w;
            int64_t last = -1;
            for (int x = 0; x < w; x++) {
                if (is_bg(src[x], jb->thresh, jb->inverse)) last = x;
                const int64_t dd = last < 0 ? cap : (int64_t)(x - last) * (x - last);
                dst[x] = (uint32_t)(dd < cap ? dd : cap);
            }
            last = -1;
            for (int x = w - 1; x >= 0; x--) {
                if (is_bg(src[x], jb->thresh, jb->inverse)) last = x;
                if (last >= 0) { const int64_t dd = (int64_t)(last - x) * (last - x); if (dd < (int64_t)dst[x]) dst[x] = (uint32_t)dd; }
            }
        }
    return NULL;
}

/* y (pass 0), then z (pass 1): blocks of 16 adjacent lines are gathered so the strided axis is read in 64-byte pieces */
static void *line_worker(void *arg) {
    job_t *jb = (job_t *)arg;
    const int w = jb->w, h = jb->h, d = jb->d;
    const size_t wh = (size_t)w * h;
    const int len = jb->pass == 0 ? h : d;
    const size_t stride = jb->pass == 0 ? (size_t)w : wh;
    const size_t outer_stride = jb->pass == 0 ? wh : (size_t)w;
    const int nblk = (w + 15) / 16;
    int64_t *f = (int64_t *)malloc(sizeof(int64_t) * (size_t)len * 16);
    int64_t *o = (int64_t *)malloc(sizeof(int64_t) * (size_t)len);
    int *v = (int *)malloc(sizeof(int) * (size_t)len);
    double *z = (double *)malloc(sizeof(double) * ((size_t)len + 1));
    long long lo, hi;
    while (take(jb, 4, &lo, &hi))
        for (long long job = lo; job < hi; job++) {
            const long long outer = job / nblk;
            const int x0 = (int)(job % nblk) * 16, xn = (w - x0 < 16) ? w - x0 : 16;
            uint32_t *base = jb->out + (size_t)outer * outer_stride + x0;
            for (int t = 0; t < len; t++)
                for (int c = 0; c < xn; c++) f[(size_t)c * len + t] = base[(size_t)t * stride + c];
            for (int c = 0; c < xn; c++) {
                envelope(f + (size_t)c * len, len, jb->cap, o, v, z);
                for (int t = 0; t < len; t++) base[(size_t)t * stride + c] = (uint32_t)o[t];
            }
        }
    free(f); free(o); free(v); free(z);
    return NULL;
}

static void run(job_t *jb, void *(*fn)(void *), long long njobs, int nthreads) {
    pthread_t th[256];
    jb->next = 0; jb->njobs = njobs;
    if (nthreads > 256) nthreads = 256;
    for (int i = 1; i < nthreads; i++) pthread_create(&th[i], NULL, fn, jb);
    fn(jb);
    for (int i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
}

/* in: u8 [d][h][w]; out_sq: u32 [d][h][w] */
int edt_exact_sq(const uint8_t *in, int w, int h, int d, int inverse, int thresh, uint32_t *out_sq, int nthreads) {
    int n = w; if (h > n) n = h; if (d > n) n = d;
    job_t jb;
    memset(&jb, 0, sizeof jb);
    jb.in = in; jb.out = out_sq; jb.w = w; jb.h = h; jb.d = d; jb.inverse = inverse; jb.thresh = thresh;
    jb.cap = 3 * (int64_t)(n + 1) * (n + 1);
    pthread_mutex_init(&jb.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    run(&jb, x_worker, (long long)h * d, nthreads);
    const int nblk = (w + 15) / 16;
    jb.pass = 0; run(&jb, line_worker, (long long)d * nblk, nthreads);
    jb.pass = 1; run(&jb, line_worker, (long long)h * nblk, nthreads);
    pthread_mutex_destroy(&jb.mu);
    return 0;
}
