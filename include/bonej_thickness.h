/*
 * bonej_thickness.h -- C-ABI of libbonejthickness.so, the B200 (sm_100a) implementation of the
 * hot path behind BoneJ's `Plugins > BoneJ > Thickness` command.
 *
 * The boundary.  The reference has no FFI: ThicknessWrapper.createMap()
 * (Modern/wrapperPlugins/src/main/java/org/bonej/wrapperPlugins/ThicknessWrapper.java:190-196)
 * calls sc.fiji.localThickness.LocalThicknessWrapper.processImage(ImagePlus) and
 * ThicknessWrapper.addMapResults() (:158-180) reads ij.process.StackStatistics mean/stdDev/max.
 * The entry points below are what a JNI shim for that seam binds (see INTEGRATION.md and
 * java/): plain pointers and sizes, no CUDA / torch types.  Volumes are [z][y][x], x fastest --
 * the same byte order as the per-slice Java arrays (ImageStack.getPixels(z+1)) laid end to end
 * and as DatasetUtil raw files (Modern/utilities/.../DatasetUtil.java:293-311).
 *
 * All functions return BJT_OK (0) or a negative BJT_E_* code; bjt_last_error() gives the text.
 * There is no CPU fallback: without a usable CUDA device every compute entry point fails.
 */
#ifndef BONEJ_THICKNESS_H
#define BONEJ_THICKNESS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BJT_OK            0
#define BJT_E_CUDA       -1   /* CUDA runtime error (no device, launch failure, out of memory)   */
#define BJT_E_ARG        -2   /* bad argument (null pointer, non-positive size, volume too large) */
#define BJT_E_NOT_BINARY -3   /* input is not a 0/255 image: processImage() would return null     */
#define BJT_E_INTERNAL   -4   /* capacity fallback exhausted (should not happen)                  */

typedef struct bjt_ctx bjt_ctx;

/* ij.process.StackStatistics as consumed at ThicknessWrapper.java:158-180: statistics over the
 * non-NaN voxels of the calibrated map.  sd uses n-1 and is 0 when the numerator is <= 0.
 * count == 0 is what the wrapper turns into NaN / NaN / NaN. */
typedef struct bjt_stats {
    int64_t count;
    double  mean, sd, min, max;
    double  sum, sum2;
} bjt_stats;

/* Per-stage device times (CUDA events on the context's stream), filled when a non-NULL pointer
 * is passed.  ms_total covers upload .. download for the host entry points. */
typedef struct bjt_timing {
    float   ms_h2d, ms_edt_x, ms_edt_y, ms_edt_z, ms_ridge, ms_paint, ms_cleanup, ms_finish, ms_d2h, ms_total;
    int64_t n_ridge;        /* distance-ridge points painted                                */
    int64_t n_launches;     /* kernels of this library launched by the call                 */
    uint32_t rsq_max;       /* largest squared EDT value                                    */
    int32_t paint_path;     /* 0 = separable sweeps (16-bit items), 1 = separable sweeps (32-bit items),
                               2 = brute-force scatter, 3 = constant fast path (no background),
                               4 = fronts along z + discs in the plane (the default)                    */
    float   ms_paint_x, ms_paint_y, ms_paint_z;   /* painter by stage.  Paths 0/1: x, y, z sweeps.  Path 4:
                                                     ms_paint_x = front kernel, ms_paint_y = disc kernel   */
    int32_t n_paint_y_sweeps;                     /* launches of the dominant kernel: paths 0/1 the y-stage
                                                     sweep (2 directions x column blocks), path 4 the disc
                                                     kernel (one per chunk of planes)                       */
} bjt_timing;

/* ---- context ------------------------------------------------------------------------------- */
int  bjt_create(bjt_ctx **out, int device);            /* device < 0: current CUDA device        */
/* One context over several GPUs: the same thickness calls below (bjt_thickness_u8, _u8_slices, _both_u8) cut the
 * stack into contiguous z-slabs, one per device, driven by one host thread each; slabs exchange EDT lines and a
 * max-radius halo through peer memory (NVLink), and the maps and statistics are bitwise those of one device.  This
 * is how the single blocking call of the reference seam (ThicknessWrapper.java:127-134, 190-196) reaches 2 .. 8
 * GPUs.  devices may name a GPU more than once (slabs then share it).  Every other entry point of a multi-device
 * context runs on devices[0]. */
int  bjt_create_multi(bjt_ctx **out, const int *devices, int ndev);
int  bjt_device_count(const bjt_ctx *ctx);             /* slabs a thickness call is cut into (1 for bjt_create) */
void bjt_destroy(bjt_ctx *ctx);
const char *bjt_last_error(const bjt_ctx *ctx);        /* ctx may be NULL (creation failures)    */
int  bjt_set_stream(bjt_ctx *ctx, void *cuda_stream);  /* cudaStream_t; NULL = default stream    */
int  bjt_release_workspace(bjt_ctx *ctx);              /* free cached device buffers             */
int  bjt_device_info(bjt_ctx *ctx, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor);
/* paint_path override for tests: -1 automatic (default), 0/1/2/4 force that path */
int  bjt_set_paint_path(bjt_ctx *ctx, int path);
/* kernels of this library launched through this context so far */
int64_t bjt_launch_count(const bjt_ctx *ctx);
/* test hook: cap the fast painter kernels' per-line capacities (0 = default) so that small volumes exercise
 * the redo kernels; results must not change.  This and bjt_paint_set_block_cols below configure the separable
 * sweeps (paint paths 0 / 1, the fallback painter) for every context on the same device / in the process: set them
 * only while no other context is painting.  The default painter (path 4) reads neither. */
int  bjt_debug_set_paint_limits(bjt_ctx *ctx, int active_cap, int front_cap);
/* The painter runs its y and z stages per block of columns, sized by the volume of the block (the lists of
 * a block are held at once).  bjt_paint_block_cols: the width the library chooses for a w x h x d volume, not
 * clipped to w (a multiple of 128, at least 512; no device needed).  bjt_paint_set_block_cols: use this width
 * instead (a multiple of 32; 0 = the library's choice again); process-wide.  The ranks of a z-slab sharded run
 * exchange boundary staircases per block, so they must agree on the width although their slabs may differ in
 * thickness: bonej2_b200/sharded.py sets the smallest of the ranks' choices on every rank. */
int  bjt_paint_block_cols(int w, int h, int d);
int  bjt_paint_set_block_cols(int cols);
/* test hook: the same call under its older name */
int  bjt_debug_set_xslab_cols(int cols);
/* self-check: the disc painter's correction-free integer square root (paint path 4) against the exact one for
 * every argument it is specified for (values below 2^18); *mismatches must come back 0 */
int  bjt_debug_isqrt_check(bjt_ctx *ctx, int64_t *mismatches);
/* lines the last separable paint of this context sent to its redo kernels */
int64_t bjt_debug_paint_redo_lines(bjt_ctx *ctx);
/* page-locked host memory (what a JNI direct ByteBuffer / the bench wraps so copies are DMA) */
void *bjt_host_alloc(size_t bytes);
void  bjt_host_free(void *p);

/* ---- the seam: LocalThicknessWrapper.processImage + StackStatistics, HOST buffers ------------ *
 * in       u8 [d][h][w], must contain only 0 and 255 (else BJT_E_NOT_BINARY)
 * inverse  0 = Tb.Th (foreground), 1 = Tb.Sp (background)   (ThicknessWrapper.java:217-222)
 * mask     maskArtefacts -> MaskThicknessMapWithOriginal      (ThicknessWrapper.java:186)
 * threshold 128 in the reference; pixel_width = calibration.pixelWidth (calibratePixels = true)
 * out_map  f32 [d][h][w], NaN background, calibrated units; may be NULL (showMaps = false)     */
int bjt_thickness_u8(bjt_ctx *ctx, const uint8_t *in, int w, int h, int d, int inverse, int mask,
                     int threshold, double pixel_width, float *out_map, bjt_stats *stats,
                     bjt_timing *timing);

/* Same, for Java's one-array-per-slice stacks: d pointers to w*h bytes / floats each. */
int bjt_thickness_u8_slices(bjt_ctx *ctx, const uint8_t *const *in_slices, int w, int h, int d,
                            int inverse, int mask, int threshold, double pixel_width,
                            float *const *out_slices, bjt_stats *stats, bjt_timing *timing);

/* mapChoice = "Both": one upload, the foreground run then the inverted run
 * (ThicknessWrapper.java:127-134).  Outputs may be NULL individually. */
int bjt_thickness_both_u8(bjt_ctx *ctx, const uint8_t *in, int w, int h, int d, int mask,
                          int threshold, double pixel_width, float *out_th, float *out_sp,
                          bjt_stats *stats_th, bjt_stats *stats_sp, bjt_timing *timing_th,
                          bjt_timing *timing_sp);

/* ---- device-resident variant (inputs / outputs already in HBM; bench `value`, sharded runs) -- */
int bjt_dev_thickness_u8(bjt_ctx *ctx, const uint8_t *d_in, int w, int h, int d, int inverse,
                         int mask, int threshold, double pixel_width, float *d_out_map,
                         bjt_stats *stats, bjt_timing *timing);

/* ---- stage-level entry points, HOST buffers (parity tests against the oracle) ---------------- */
/* EDT_S1D before its sqrt pass: uint32 squared distance, background -> 0, sentinel 3(n+1)^2   */
int bjt_edt_sq_u8(bjt_ctx *ctx, const uint8_t *in, int w, int h, int d, int inverse, int threshold,
                  uint32_t *out_sq);
/* Distance_Ridge on the squared map: r^2 at ridge points, 0 elsewhere */
int bjt_ridge_u32(bjt_ctx *ctx, const uint32_t *sq, int w, int h, int d, uint32_t *out_ridge);
/* Distance_Ridge.createTemplate / scanCube for a list of r^2 values */
int bjt_ridge_template(bjt_ctx *ctx, const int32_t *rsq_values, int n, int32_t *t1, int32_t *t2, int32_t *t3);
/* Local_Thickness_Parallel before its sqrt: painted max r^2 per voxel; path as bjt_timing.paint_path (-1 auto) */
int bjt_paint_rsq(bjt_ctx *ctx, const uint32_t *ridge, int w, int h, int d, int path, uint32_t *out_rsq);
/* 2*sqrt + Clean_Up_Local_Thickness (+ mask if original != NULL) + 0->NaN, x pixel_width (if calibrate) */
int bjt_cleanup_f32(bjt_ctx *ctx, const uint32_t *rsq, const uint8_t *original, int w, int h, int d,
                    int inverse, int calibrate, double pixel_width, float *out_map, bjt_stats *stats);
int bjt_stats_f32(bjt_ctx *ctx, const float *map, size_t n, bjt_stats *stats);

/* ---- consumers of the map: the reductions Legacy plugins run over a thickness map ------------- *
 * Per particle label, ParticleAnalysis.getMeanStdDev (Legacy/bonej/.../ParticleAnalysis.java:325-364): over the
 * voxels of label p with value > 0, out[3p] = sum / sizes[p], out[3p+1] = sqrt(sum (v - mean)^2 / sizes[p]),
 * out[3p+2] = max; label 0: mean and sd stay 0.  sizes = the particles' voxel counts (the caller has them).
 * A label outside [0, n_labels) on a voxel with a value is BJT_E_ARG.                                       */
int bjt_label_stats_f32(bjt_ctx *ctx, const float *map, const int32_t *labels, size_t n, int n_labels,
                        const int64_t *sizes, double *out);
int bjt_dev_label_stats_f32(bjt_ctx *ctx, const float *d_map, const int32_t *d_labels, size_t n, int n_labels,
                            const int64_t *sizes, double *out);
/* Per slice inside an ROI rectangle, SliceGeometry (Legacy/bonej/.../SliceGeometry.java:876-900, 1172-1200): over the
 * pixels with value > 0, out[3z] = mean, out[3z+1] = max (0 if none), out[3z+2] = sqrt(sum (mean - v)^2 / count);
 * an empty slice gives NaN for mean and sd like the Java arithmetic; counts (may be NULL) = pixels counted.    */
int bjt_slice_stats_f32(bjt_ctx *ctx, const float *map, int w, int h, int d, int rx, int ry, int rw, int rh,
                        double *out, int64_t *counts);
int bjt_dev_slice_stats_f32(bjt_ctx *ctx, const float *d_map, int w, int h, int d, int rx, int ry, int rw, int rh,
                            double *out, int64_t *counts);

/* org.bonej.ops.skeletonize.FindRidgePoints (Modern/ops/.../FindRidgePoints.java:55-116), the seed finder of
 * Ellipsoid Factor, on the CUDA EDT.  in: w*h*d bytes, non-zero = foreground.  fraction: thresholdForBeingARidgePoint
 * (reference default 0.6).  Writes up to cap points (x, y, z as doubles, voxel-centre coordinates of the input, in
 * the reference's iteration order) and the number found (may exceed cap); ridge_max (may be NULL) = largest ridge value. */
int bjt_find_ridge_points_u8(bjt_ctx *ctx, const uint8_t *in, int w, int h, int d, double fraction, double *out_xyz,
                             int64_t cap, int64_t *n_found, float *ridge_max);

/* ---- stage-level entry points, DEVICE buffers (z-slab sharded orchestration, profiling) ------- *
 * n_sentinel is max(w,h,d) of the WHOLE volume (a slab passes the global value).               */
int bjt_dev_edt_x(bjt_ctx *ctx, const uint8_t *d_in, int w, int h, int d, int inverse, int threshold,
                  uint16_t *d_gx);
int bjt_dev_edt_y(bjt_ctx *ctx, const uint16_t *d_gx, int w, int h, int d, int n_sentinel, uint32_t *d_gy);
int bjt_dev_edt_z(bjt_ctx *ctx, const uint32_t *d_gy, int w, int h, int d, int n_sentinel, uint32_t *d_sq);
int bjt_dev_ridge(bjt_ctx *ctx, const uint32_t *d_sq, int w, int h, int d, uint32_t *d_ridge);
int bjt_dev_paint(bjt_ctx *ctx, const uint32_t *d_ridge, int w, int h, int d, int path, uint32_t *d_rsq);

/* The painter on a z-slab with halo planes (z-slab sharded runs; SURVEY.md 8e "max-radius halo exchange"): d_ridge
 * holds d planes, balls centred on any of them count, and planes [z_lo, z_hi) are painted into d_rsq (z_hi - z_lo
 * planes).  max_rsq = the largest r^2 of the WHOLE volume (every rank passes the same value; < 0: found here).      */
int bjt_dev_paint_range(bjt_ctx *ctx, const uint32_t *d_ridge, int w, int h, int d, int z_lo, int z_hi, int64_t max_rsq,
                        uint32_t *d_rsq);

/* The painter of one z-slab in steps (Local_Thickness_Parallel over a volume cut along z).  Its x and y
 * stages are plane-local; only the z sweeps cross slab faces, and what crosses is small: per (x, y) column
 * the staircase of (tag, reach beyond the face) of the balls centred next to the face.
 *   open(ridge of the slab's own planes, max_rsq = largest tag anywhere)      x stage
 *   for xs in 0 .. nxslabs-1  (column blocks of xslab_width(xs) columns):
 *       lists(xs)                                                             y stage of the block
 *       export(xs, side, gap, nplanes, count, items, kcap)   for every slab that lies within reach:
 *           side +1 / -1: the receiver lies above / below; gap = planes between my face plane and the
 *           receiver's first plane, counted from 1 for the adjacent slab; nplanes >= reach of the largest
 *           ball; count[ncol] and items[ncol * kcap] (8 bytes each: reach beyond, tag), ncol =
 *           xslab_width(xs) * h, column index y * xslab_width + x
 *       -- exchange the exported buffers between ranks --
 *       sweep(xs, imports from below, imports from above, kcap, rsq of the slab's own planes)   z stage
 *   close(&flags): flags != 0 means a list or staircase capacity was exceeded and rsq is NOT valid; run the
 *   slab through bjt_dev_paint on slab + halo instead (it is never silent).                              */
int bjt_dev_paint_open(bjt_ctx *ctx, const uint32_t *d_ridge, int w, int h, int d, int64_t max_rsq);
int bjt_dev_paint_nxslabs(bjt_ctx *ctx);
int bjt_dev_paint_xslab_width(bjt_ctx *ctx, int xs);
int bjt_dev_paint_lists(bjt_ctx *ctx, int xs);
int bjt_dev_paint_export(bjt_ctx *ctx, int xs, int side, int gap, int nplanes, uint32_t *d_count, void *d_items, int kcap);
int bjt_dev_paint_sweep(bjt_ctx *ctx, int xs, int n_lo, const uint32_t *const *lo_count, const void *const *lo_items,
                        int n_hi, const uint32_t *const *hi_count, const void *const *hi_items, int kcap, uint32_t *d_rsq);
int bjt_dev_paint_close(bjt_ctx *ctx, uint32_t *overflow_flags);
int bjt_dev_finish(bjt_ctx *ctx, const uint32_t *d_rsq, const uint8_t *d_original, int w, int h, int d,
                   int inverse, int calibrate, double pixel_width, float *d_out, bjt_stats *stats);
/* same on a slab with halo planes: the map and the statistics cover planes [z_lo, z_hi) of the local
 * volume only (d_out holds z_hi - z_lo planes); planes outside serve as cleanup neighbours.  stats->sum,
 * sum2, count, min, max are what ranks combine before forming mean and sd. */
int bjt_dev_finish_range(bjt_ctx *ctx, const uint32_t *d_rsq, const uint8_t *d_original, int w, int h, int d,
                         int z_lo, int z_hi, int inverse, int calibrate, double pixel_width, float *d_out,
                         bjt_stats *stats);

/* ---- device buffers that other processes read directly (one process per GPU: bonej2_b200/sharded.py) ---- *
 * bjt_dev_alloc: a named device buffer owned by the context (kept until bjt_release_workspace / bjt_destroy).
 * bjt_ipc_export / _import: a 64-byte CUDA IPC handle of such a buffer, and its mapping in another process (peer
 * access over NVLink is enabled on first use); bjt_ipc_close unmaps.  bjt_dev_copy2d: strided device-to-device copy
 * on the context's stream (height rows of width bytes), local or out of a mapped peer buffer -- the transposes
 * of the EDT z pass ("NCCL all-to-all transpose" of the north star) and the max-radius halo are such pulls. */
int bjt_dev_alloc(bjt_ctx *ctx, const char *name, size_t bytes, void **ptr);
int bjt_ipc_export(bjt_ctx *ctx, const void *ptr, unsigned char *handle64);
int bjt_ipc_import(bjt_ctx *ctx, const unsigned char *handle64, void **ptr);
int bjt_ipc_close(bjt_ctx *ctx, void *ptr);
int bjt_dev_copy2d(bjt_ctx *ctx, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height);
/* n such copies in one kernel launch (the shares of all peers at once; 16-byte aligned blocks are moved by the SMs) */
int bjt_dev_copy2d_batch(bjt_ctx *ctx, int n, void *const *dst, const size_t *dpitch, const void *const *src,
                         const size_t *spitch, const size_t *width, const size_t *height);

/* ---- synthetic trabecular phantom (workload of BASELINE.json) ------------------- */
/* shapes: n records of 16 int32 (bonej2_b200/phantom.py); out: u8 [d][h][w] in {0,255}          */
int bjt_phantom_u8(bjt_ctx *ctx, const int32_t *shapes, int nshapes, int w, int h, int d, uint8_t *out);
int bjt_dev_phantom_u8(bjt_ctx *ctx, const int32_t *shapes_host, int nshapes, int w, int h, int d,
                       int z0, int dz, uint8_t *d_out);   /* rasterise slices [z0, z0+dz) of the volume */

#ifdef __cplusplus
}
#endif
#endif /* BONEJ_THICKNESS_H */
