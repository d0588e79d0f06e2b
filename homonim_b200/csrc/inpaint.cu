// inpaint.cu -- low-R2 parameter in-painting and gain refit of the gain-offset model (sm_100a)
//   homonim/kernel_model.py:361-371:
//       r2_mask = (R2 > thresh) & (gain > 0) & mask
//       offset  = rasterio.fill.fillnodata(offset, r2_mask)          # GDALFillNodata, max distance 100, no smoothing
//       params[:, ~mask] = nan
//       gain[~r2_mask & mask] = (sum_ref - N * offset) / sum_src
//   The GDALFillNodata algorithm restated here is specified in oracle/gdal_restate.c (gr_fillnodata): four-quadrant
//   nearest-source search over per-column "last valid pixel above / below" tables, inverse-distance weighting.
#include "hb_common.cuh"

namespace {

struct FillTables {
    int *top_y;      // row of the last source pixel at or above (y, x); -1 if none
    float *top_v;    // its offset value
    int *bot_y;      // row of the first source pixel at or below (y, x); -1 if none
    float *bot_v;
    uint8_t *r2m;    // the r2_mask
};

// "Last source pixel at or above / below" tables.  One CTA = 32 columns (lane = column, coalesced 128-byte rows) x 32
// warps; warp w owns the row chunk [w*rc, (w+1)*rc).  Pass 1: every warp scans its chunk down and up, writes the
// r2_mask and publishes the chunk's last / first source pixel per column in shared memory.  Pass 2: every warp takes
// its carry-in from the nearest chunk above / below that has a source pixel and re-scans, writing the tables.
constexpr int kScanWarps = 32;

__global__ void __launch_bounds__(kScanWarps * 32)
inpaint_scan_kernel(const float *__restrict__ params, const float *__restrict__ sums, long h, long w, float thresh,
                    FillTables tb)
{
    __shared__ int s_dn_y[kScanWarps][32], s_up_y[kScanWarps][32];
    __shared__ float s_dn_v[kScanWarps][32], s_up_v[kScanWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long x = (long)blockIdx.x * 32 + lane;
    const bool col_ok = x < w;
    const long plane = h * w;
    const long rc = (h + kScanWarps - 1) / kScanWarps;
    const long ya = min((long)warp * rc, h), yb = min(ya + rc, h);

    // pass 1: r2_mask, and the chunk's own last (scanning down) / first (scanning up) source pixel
    int dn_y = -1, up_y = -1;
    float dn_v = 0.f, up_v = 0.f;
    if (col_ok) {
        for (long y = ya; y < yb; y++) {
            const long i = y * w + x;
            const float gain = params[i], off = params[plane + i], r2 = params[2 * plane + i];
            const bool mask = sums[2 * plane + i] >= 0.f;            // count plane holds -1 outside the mask
            const bool m = mask && (r2 > thresh) && (gain > 0.f);    // comparisons with nan are false (:363)
            tb.r2m[i] = m ? 1 : 0;
            if (m) {
                dn_y = (int)y; dn_v = off;
                if (up_y < 0) { up_y = (int)y; up_v = off; }
            }
        }
    }
    s_dn_y[warp][lane] = dn_y; s_dn_v[warp][lane] = dn_v;
    s_up_y[warp][lane] = up_y; s_up_v[warp][lane] = up_v;
    __syncthreads();
    if (!col_ok) return;
    // pass 2: carry-in from the other chunks, then the tables
    int ly = -1;
    float lv = 0.f;
    for (int ww = warp - 1; ww >= 0; ww--)
        if (s_dn_y[ww][lane] >= 0) { ly = s_dn_y[ww][lane]; lv = s_dn_v[ww][lane]; break; }
    for (long y = ya; y < yb; y++) {
        const long i = y * w + x;
        if (tb.r2m[i]) { ly = (int)y; lv = params[plane + i]; }
        tb.top_y[i] = ly;
        tb.top_v[i] = lv;
    }
    ly = -1;
    lv = 0.f;
    for (int ww = warp + 1; ww < kScanWarps; ww++)
        if (s_up_y[ww][lane] >= 0) { ly = s_up_y[ww][lane]; lv = s_up_v[ww][lane]; break; }
    for (long y = yb - 1; y >= ya; y--) {
        const long i = y * w + x;
        if (tb.r2m[i]) { ly = (int)y; lv = params[plane + i]; }
        tb.bot_y[i] = ly;
        tb.bot_v[i] = lv;
    }
}

#define HB_QUAD_CHECK(qd, qv, tx, ty, tv)                                                                   \
    if ((ty) >= 0) {                                                                                        \
        const double ddx = (double)(tx) - (double)x, ddy = (double)(ty) - (double)y;                        \
        const double d2 = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));                              \
        if (d2 < __dmul_rn((qd), (qd))) { (qd) = sqrt(d2); (qv) = (double)(tv); }                           \
    }

// thread per pixel; only pixels inside the mask that failed the R2 test do any work
__global__ void inpaint_fill_kernel(float *__restrict__ params, const float *__restrict__ sums, long h, long w,
                                    double max_dist, FillTables tb)
{
    const long n = h * w;
    const long max_dist_i = (long)floor(max_dist);
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (tb.r2m[i]) continue;
        const float fN = sums[2 * n + i];
        if (!(fN >= 0.f)) continue;                               // outside the mask: stays nan (:367)
        const long y = i / w, x = i - y * w;
        const int *ty = tb.top_y + y * w, *by = tb.bot_y + y * w;
        const float *tv = tb.top_v + y * w, *bv = tb.bot_v + y * w;
        double qd0 = max_dist + 1.0, qd1 = qd0, qd2 = qd0, qd3 = qd0;
        double qv0 = 0.0, qv1 = 0.0, qv2 = 0.0, qv3 = 0.0;
        long this_max = max_dist_i;
        for (long step = 0; step <= this_max; step++) {
            const long lx = (x - step < 0) ? 0 : x - step;
            const long rx = (x + step > w - 1) ? w - 1 : x + step;
            const int tyl = ty[lx], byl = by[lx];
            HB_QUAD_CHECK(qd0, qv0, lx, tyl, tv[lx])              // top left (includes the current row)
            HB_QUAD_CHECK(qd1, qv1, lx, byl, bv[lx])              // bottom left
            if (step == 0) continue;                              // right quadrants exclude the centre column
            const int tyr = ty[rx], byr = by[rx];
            HB_QUAD_CHECK(qd2, qv2, rx, tyr, tv[rx])              // top right
            HB_QUAD_CHECK(qd3, qv3, rx, byr, bv[rx])              // bottom right
            if ((step & 0x3) == 0) {
                const long lim = (long)floor(fmax(fmax(qd0, qd1), fmax(qd2, qd3)));
                if (lim < this_max) this_max = lim;
            }
        }
        double wsum = 0.0, vsum = 0.0;
        bool found = false;
        const double qd[4] = {qd0, qd1, qd2, qd3}, qv[4] = {qv0, qv1, qv2, qv3};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (qd[q] <= max_dist) {
                const double wq = 1.0 / qd[q];
                found = true;
                wsum = __dadd_rn(wsum, wq);
                vsum = __dadd_rn(vsum, __dmul_rn(qv[q], wq));
            }
        }
        float off = params[n + i];
        if (found) {
            off = (float)(vsum / wsum);
            params[n + i] = off;
        }
        // gain refit through (mean(src), mean(ref)) with the in-painted offset (:371), float32 arithmetic
        const float fR = sums[i], fS = sums[n + i];
        params[i] = __fdiv_rn(__fsub_rn(fR, __fmul_rn(fN, off)), fS);
    }
}

}  // namespace

extern "C" size_t hb_inpaint_workspace_bytes(long h, long w)
{
    const size_t n = (size_t)h * (size_t)w;
    return n * (4 * sizeof(float)) + ((n + 15) / 16) * 16;
}

extern "C" int hb_inpaint_refit(float *params_dev, const float *sums_dev, long h, long w, double r2_thresh,
                                double max_search_dist, void *workspace_dev, size_t workspace_bytes, void *stream)
{
    HB_REQUIRE(params_dev && sums_dev && workspace_dev && h > 0 && w > 0, "hb_inpaint_refit: bad arguments");
    HB_REQUIRE(workspace_bytes >= hb_inpaint_workspace_bytes(h, w), "hb_inpaint_refit: workspace too small");
    HB_REQUIRE(((uintptr_t)workspace_dev) % 4 == 0, "hb_inpaint_refit: workspace must be 4-byte aligned");
    HB_REQUIRE(h < 2147483647L, "hb_inpaint_refit: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h * (size_t)w;
    FillTables tb;
    tb.top_y = (int *)workspace_dev;
    tb.top_v = (float *)(tb.top_y + n);
    tb.bot_y = (int *)(tb.top_v + n);
    tb.bot_v = (float *)(tb.bot_y + n);
    tb.r2m = (uint8_t *)(tb.bot_v + n);
    // numpy compares the float32 R2 plane with a Python float: the threshold is used as float32 (NEP 50)
    const float thresh = (float)r2_thresh;
    inpaint_scan_kernel<<<(unsigned)((w + 31) / 32), kScanWarps * 32, 0, st>>>(params_dev, sums_dev, h, w, thresh, tb);
    HB_LAUNCH_OK("inpaint_scan_kernel");
    long blocks = ((long)n + 255) / 256;
    const long cap = (long)hb_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    inpaint_fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(params_dev, sums_dev, h, w, max_search_dist, tb);
    HB_LAUNCH_OK("inpaint_fill_kernel");
    return 0;
}
