/*
 * oracle/gdal_restate.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, OpenMP over destination rows) of the three GDAL algorithms that the reference's hot
 * path reaches through rasterio and whose source is NOT under /root/reference:
 *
 *   - rasterio.warp.reproject(..., resampling=average)       (call sites: raster_array.py:573-577 via
 *     kernel_model.py:397, 480, 520)                          -> GDAL  GWKAverageOrModeThread   (GRA_Average)
 *   - rasterio.warp.reproject(..., resampling=cubic_spline)   (kernel_model.py:491, 520)
 *                                                             -> GDAL  GWKRealCase + GWKResample (GRA_CubicSpline)
 *   - rasterio.warp.reproject(..., resampling=nearest)        (kernel_model.py:497) -> GDAL GWKRealCase (nearest)
 *   - rasterio.fill.fillnodata(image, mask)                   (kernel_model.py:366) -> GDAL GDALFillNodata
 *
 * Dependency: rasterio>=1.1 (pyproject.toml:7, unpinned) which bundles GDAL (3.x in current wheels).  GDAL is not
 * installed in this image, so these functions restate the published algorithm (alg/gdalwarpkernel.cpp,
 * alg/rasterfill.cpp) for axis-aligned, same-CRS, north-up grids.  Nothing executable here can check them against
 * GDAL bit for bit.  PINNING: GRA_Average and GRA_CubicSpline are pinned at the known-answer level to real GDAL
 * output -- the fuse + compare table the reference publishes for its own test images (docs/cli.rst:58-72, 32
 * numbers) is reproduced to every printed digit through these functions (oracle/make_golden_docs.py,
 * tests/test_oracle_golden.py).  GDALFillNodata and nearest are anchored only on the reference's loose known-answer
 * tests (tests/test_kernel_model.py:41-117, 166-273): for those, PARITY UNPINNED.
 *
 * Grid mapping convention used by every resampler: destination pixel-EDGE coordinate u (column) maps to source
 * pixel-edge coordinate  sx*u + ox  (rows: sy*v + oy), sx, sy > 0.  Destination pixel j covers [j, j+1), its centre is
 * j + 0.5.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int gr_invalid(double v, int has_nd, double nd)
{
    if (!has_nd) return 0;
    if (isnan(nd)) return isnan(v);
    return v == nd;
}

/* ------------------------------------------------------------------------------------------------------------------
 * GRA_Average.  For each destination pixel: weighted mean of the VALID source pixels intersecting its footprint,
 * edge pixels weighted by their fractional overlap (GDAL >= 3.3), accumulated in double, row-major (y outer, x inner).
 * Destination is only written when the total weight is > 0 (caller pre-fills it with the destination nodata).
 * ------------------------------------------------------------------------------------------------------------------ */
#define DEFINE_AVERAGE(NAME, T)                                                                                       \
void NAME(const T *src, long hs, long ws, int has_nd, double nd, float *dst, long hd, long wd,                        \
          double sx, double ox, double sy, double oy)                                                                 \
{                                                                                                                     \
    _Pragma("omp parallel for schedule(static)")                                                                      \
    for (long i = 0; i < hd; i++) {                                                                                   \
        const double y_min = sy * (double)i + oy, y_max = sy * (double)(i + 1) + oy;                                  \
        long iy0 = (long)fmax(floor(y_min + 1e-10), 0.0);                                                             \
        long iy1 = (long)fmin(ceil(y_max - 1e-10), (double)hs);                                                       \
        if (y_max <= 0.0 || y_min >= (double)hs) continue;                                                            \
        if (iy0 == iy1 && iy1 < hs) iy1++;                                                                            \
        for (long j = 0; j < wd; j++) {                                                                               \
            const double x_min = sx * (double)j + ox, x_max = sx * (double)(j + 1) + ox;                              \
            long ix0 = (long)fmax(floor(x_min + 1e-10), 0.0);                                                         \
            long ix1 = (long)fmin(ceil(x_max - 1e-10), (double)ws);                                                   \
            if (x_max <= 0.0 || x_min >= (double)ws) continue;                                                        \
            if (ix0 == ix1 && ix1 < ws) ix1++;                                                                        \
            double total = 0.0, total_w = 0.0;                                                                        \
            for (long y = iy0; y < iy1; y++) {                                                                        \
                double wy = 1.0;                                                                                      \
                if (y == iy0) wy = (iy0 + 1 == iy1) ? 1.0 : 1.0 - (y_min - (double)iy0);                              \
                else if (y + 1 == iy1) wy = 1.0 - ((double)iy1 - y_max);                                              \
                const T *row = src + y * ws;                                                                          \
                for (long x = ix0; x < ix1; x++) {                                                                    \
                    const double v = (double)row[x];                                                                  \
                    if (gr_invalid(v, has_nd, nd)) continue;                                                          \
                    double w = wy;                                                                                    \
                    if (x == ix0) w = (ix0 + 1 == ix1) ? wy : wy * (1.0 - (x_min - (double)ix0));                     \
                    else if (x + 1 == ix1) w = wy * (1.0 - ((double)ix1 - x_max));                                    \
                    total_w += w;                                                                                     \
                    total += v * w;                                                                                   \
                }                                                                                                     \
            }                                                                                                         \
            if (total_w > 0.0) dst[i * wd + j] = (float)(total / total_w);                                            \
        }                                                                                                             \
    }                                                                                                                 \
}

DEFINE_AVERAGE(gr_average_f32, float)
DEFINE_AVERAGE(gr_average_f64, double)
DEFINE_AVERAGE(gr_average_u8, uint8_t)
DEFINE_AVERAGE(gr_average_u16, uint16_t)

/* cubic B-spline basis, GDAL's GWKBSpline: 1/6 [ (x+2)^3+ - 4 (x+1)^3+ + 6 x^3+ - 4 (x-1)^3+ ] */
static inline double gr_bspline(double x)
{
    const double xp2 = x + 2.0, xp1 = x + 1.0, xm1 = x - 1.0;
    const double xp2c = xp2 * xp2 * xp2;
    if (!(xp2 > 0.0)) return 0.0;
    double r = xp2c;
    if (xp1 > 0.0) {
        r += -4.0 * xp1 * xp1 * xp1;
        if (x > 0.0) {
            r += 6.0 * x * x * x;
            if (xm1 > 0.0) r += -4.0 * xm1 * xm1 * xm1;
        }
    }
    return r * 0.16666666666666666666;
}

/* ------------------------------------------------------------------------------------------------------------------
 * GRA_CubicSpline for UP-sampling (destination finer than source; no kernel scaling), nb bands that share a
 * "unified" validity = any band valid (GDAL UNIFIED_SRC_NODATA=PARTIAL default).  Per destination pixel:
 *   1. the source pixel containing the destination centre must be in range and unified-valid, else skip;
 *   2. 4x4 taps around (X-0.5, Y-0.5), B-spline weights, out-of-range / unified-invalid / band-invalid taps skipped;
 *   3. skip if sum(w) < 1e-6; divide by sum(w) only if it is outside [0.99999, 1.00001]; store as float.
 * src is [nb, hs, ws] double; dst is [nb, hd, wd] float pre-filled with the destination nodata.
 * ------------------------------------------------------------------------------------------------------------------ */
void gr_cubic_spline_up(const double *src, long nb, long hs, long ws, int has_nd, double nd, float *dst, long hd,
                        long wd, double sx, double ox, double sy, double oy)
{
    const long splane = hs * ws, dplane = hd * wd;
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < hd; i++) {
        const double src_y = sy * ((double)i + 0.5) + oy;
        long cy = (long)floor(src_y + 1e-10);
        if (src_y < 0.0) continue;
        if (cy == hs) cy--;
        if (cy < 0 || cy >= hs) continue;
        const long ky = (long)floor(src_y - 0.5);
        const double dy = src_y - 0.5 - (double)ky;
        double wy[4];
        for (int t = 0; t < 4; t++) wy[t] = gr_bspline((double)(t - 1) - dy);
        for (long j = 0; j < wd; j++) {
            const double src_x = sx * ((double)j + 0.5) + ox;
            long cx = (long)floor(src_x + 1e-10);
            if (src_x < 0.0) continue;
            if (cx == ws) cx--;
            if (cx < 0 || cx >= ws) continue;
            int any_valid = 0;
            for (long b = 0; b < nb; b++) any_valid |= !gr_invalid(src[b * splane + cy * ws + cx], has_nd, nd);
            if (!any_valid) continue;
            const long kx = (long)floor(src_x - 0.5);
            const double dx = src_x - 0.5 - (double)kx;
            double wx[4];
            for (int t = 0; t < 4; t++) wx[t] = gr_bspline((double)(t - 1) - dx);
            for (long b = 0; b < nb; b++) {
                double acc = 0.0, acc_w = 0.0;
                for (int tj = 0; tj < 4; tj++) {
                    const long y = ky + tj - 1;
                    if (y < 0 || y >= hs) continue;
                    for (int ti = 0; ti < 4; ti++) {
                        const long x = kx + ti - 1;
                        if (x < 0 || x >= ws) continue;
                        int uni = 0;
                        for (long bb = 0; bb < nb; bb++) uni |= !gr_invalid(src[bb * splane + y * ws + x], has_nd, nd);
                        if (!uni) continue;
                        const double v = src[b * splane + y * ws + x];
                        if (gr_invalid(v, has_nd, nd)) continue;
                        const double w = wx[ti] * wy[tj];
                        acc_w += w;
                        acc += v * w;
                    }
                }
                if (acc_w < 0.000001) continue;
                if (acc_w < 0.99999 || acc_w > 1.00001) acc /= acc_w;
                dst[b * dplane + i * wd + j] = (float)acc;
            }
        }
    }
}

/* GRA_NearestNeighbour: destination centre -> containing source pixel; dst pre-filled with destination nodata. */
void gr_nearest(const double *src, long hs, long ws, int has_nd, double nd, float *dst, long hd, long wd,
                double sx, double ox, double sy, double oy)
{
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < hd; i++) {
        const double src_y = sy * ((double)i + 0.5) + oy;
        long cy = (long)floor(src_y + 1e-10);
        if (src_y < 0.0) continue;
        if (cy == hs) cy--;
        if (cy < 0 || cy >= hs) continue;
        for (long j = 0; j < wd; j++) {
            const double src_x = sx * ((double)j + 0.5) + ox;
            long cx = (long)floor(src_x + 1e-10);
            if (src_x < 0.0) continue;
            if (cx == ws) cx--;
            if (cx < 0 || cx >= ws) continue;
            const double v = src[cy * ws + cx];
            if (gr_invalid(v, has_nd, nd)) continue;
            dst[i * wd + j] = (float)v;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * GDALFillNodata(target, mask, max_search_dist, smoothing_iterations = 0), inverse-distance weighting.
 *   mask != 0 : source pixel (kept);  mask == 0 : pixel to fill.
 * Pass 1 (top-down) and pass 2 (bottom-up) keep, per column, the row index and float32 value of the last valid pixel
 * seen (the current row included).  For each pixel to fill, columns x-step (left) and x+step (right), step =
 * 0..max_dist, are probed (column indices clamped to the raster); the left column feeds the top-left / bottom-left
 * quadrants, the right column (step > 0 only) the top-right / bottom-right ones; a candidate replaces the quadrant's
 * current one when dist^2 < (current dist)^2, with current dist = sqrt(previous dist^2) (so the comparison carries the
 * rounding of sqrt).  Every 4 steps the search radius shrinks to floor(max quadrant distance).  Result =
 * sum(v/d) / sum(1/d) over quadrants with d <= max_search_dist, stored as float32; untouched if no quadrant found.
 * Only ORIGINAL valid pixels are ever used as sources.
 * ------------------------------------------------------------------------------------------------------------------ */
#define GR_QUAD_CHECK(qd, qv, tx, ty, tv)                                                                             \
    if ((ty) >= 0) {                                                                                                  \
        const double ddx = (double)(tx) - (double)x, ddy = (double)(ty) - (double)y;                                  \
        const double d2 = ddx * ddx + ddy * ddy;                                                                      \
        if (d2 < (qd) * (qd)) { (qd) = sqrt(d2); (qv) = (double)(tv); }                                               \
    }

void gr_fillnodata(float *img, const uint8_t *mask, long h, long w, double max_search_dist)
{
    const long max_dist_i = (long)floor(max_search_dist);
    int32_t *top_y = (int32_t *)malloc(sizeof(int32_t) * h * w);
    float *top_v = (float *)malloc(sizeof(float) * h * w);
    int32_t *bot_y = (int32_t *)malloc(sizeof(int32_t) * h * w);
    float *bot_v = (float *)malloc(sizeof(float) * h * w);
    float *out = (float *)malloc(sizeof(float) * h * w);
    memcpy(out, img, sizeof(float) * h * w);

    for (long x = 0; x < w; x++) {
        int32_t ly = -1; float lv = 0.f;
        for (long y = 0; y < h; y++) {
            if (mask[y * w + x]) { ly = (int32_t)y; lv = img[y * w + x]; }
            top_y[y * w + x] = ly; top_v[y * w + x] = lv;
        }
        ly = -1; lv = 0.f;
        for (long y = h - 1; y >= 0; y--) {
            if (mask[y * w + x]) { ly = (int32_t)y; lv = img[y * w + x]; }
            bot_y[y * w + x] = ly; bot_v[y * w + x] = lv;
        }
    }

    #pragma omp parallel for schedule(dynamic, 4)
    for (long y = 0; y < h; y++) {
        const int32_t *ty = top_y + y * w, *by = bot_y + y * w;
        const float *tv = top_v + y * w, *bv = bot_v + y * w;
        for (long x = 0; x < w; x++) {
            if (mask[y * w + x]) continue;
            double qd[4] = {max_search_dist + 1.0, max_search_dist + 1.0, max_search_dist + 1.0,
                            max_search_dist + 1.0};
            double qv[4] = {0.0, 0.0, 0.0, 0.0};
            long this_max = max_dist_i;
            for (long step = 0; step <= this_max; step++) {
                const long lx = (x - step < 0) ? 0 : x - step;
                const long rx = (x + step > w - 1) ? w - 1 : x + step;
                GR_QUAD_CHECK(qd[0], qv[0], lx, ty[lx], tv[lx])   /* top left, includes current row */
                GR_QUAD_CHECK(qd[1], qv[1], lx, by[lx], bv[lx])   /* bottom left */
                if (step == 0) continue;                          /* right quadrants exclude the centre column */
                GR_QUAD_CHECK(qd[2], qv[2], rx, ty[rx], tv[rx])   /* top right */
                GR_QUAD_CHECK(qd[3], qv[3], rx, by[rx], bv[rx])   /* bottom right */
                if ((step & 0x3) == 0) {
                    const double m = fmax(fmax(qd[0], qd[1]), fmax(qd[2], qd[3]));
                    const long lim = (long)floor(m);
                    if (lim < this_max) this_max = lim;
                }
            }
            double wsum = 0.0, vsum = 0.0; int found = 0;
            for (int q = 0; q < 4; q++) {
                if (qd[q] <= max_search_dist) {
                    const double wq = 1.0 / qd[q];
                    found = 1;
                    wsum += wq;
                    vsum += qv[q] * wq;
                }
            }
            if (found) out[y * w + x] = (float)(vsum / wsum);
        }
    }
    memcpy(img, out, sizeof(float) * h * w);
    free(top_y); free(top_v); free(bot_y); free(bot_v); free(out);
}
