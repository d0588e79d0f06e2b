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

// per pixel: the nearest source pixel (r2_mask == 1) at or above / at or below it in its column -- row (-1: none) and
// offset value, packed so that one 8-byte load fetches both
struct __align__(8) ColHit { int y; float v; };

struct FillTables {
    ColHit *top;     // last source pixel at or above (y, x)
    ColHit *bot;     // first source pixel at or below (y, x)
    uint8_t *r2m;    // the r2_mask
};

// ---- tables, small rasters (the proc grid of a RefSpace fit: h <= 8192) -----------------------------------------------
// One CTA = 8 columns x all rows, every phase parallel over rows:
//   A. r2_mask of every pixel (coalesced 32-byte row segments; a thread's rows are independent loads), collected as one
//      bit per row in per-column shared-memory bit arrays;
//   B. per column and 32-row word: the last source row before / first source row after the word (short serial pass);
//   C. every pixel finds its nearest source row above / below with two bit scans of its word (or takes B's answer),
//      gathers the offset there and writes both tables.
constexpr int kScanCols = 8;
constexpr int kScanThreads = 256;
constexpr int kScanMaxWords = 256;               // h <= 8192

__global__ void __launch_bounds__(kScanThreads)
inpaint_scan_small_kernel(const float *__restrict__ params, const float *__restrict__ sums, int h, long w, float thresh,
                          FillTables tb)
{
    __shared__ unsigned s_bits[kScanCols][kScanMaxWords];
    __shared__ int s_prev[kScanCols][kScanMaxWords], s_next[kScanCols][kScanMaxWords];
    const int t = threadIdx.x, col = t % kScanCols, rslot = t / kScanCols;
    constexpr int kRowSlots = kScanThreads / kScanCols;              // 32 rows per sweep of the CTA
    const long x = (long)blockIdx.x * kScanCols + col;
    const bool col_ok = x < w;
    const long plane = (long)h * w;
    const int words = (h + 31) / 32;
    for (int i = t; i < kScanCols * words; i += kScanThreads) s_bits[i / words][i % words] = 0u;
    __syncthreads();
    // ---- A ----------------------------------------------------------------------------------------------------------
    if (col_ok) {
        constexpr int kB = 4;
        for (int y0 = rslot; y0 < h; y0 += kB * kRowSlots) {
            float gain[kB], r2[kB], cnt[kB];
#pragma unroll
            for (int u = 0; u < kB; u++) {
                const int y = min(y0 + u * kRowSlots, h - 1);
                const long i = (long)y * w + x;
                gain[u] = __ldg(params + i); r2[u] = __ldg(params + 2 * plane + i); cnt[u] = __ldg(sums + 2 * plane + i);
            }
#pragma unroll
            for (int u = 0; u < kB; u++) {
                const int y = y0 + u * kRowSlots;
                if (y >= h) break;
                const bool mask = cnt[u] >= 0.f;                     // count plane holds -1 outside the mask
                const bool m = mask && (r2[u] > thresh) && (gain[u] > 0.f);   // comparisons with nan are false (:363)
                tb.r2m[(long)y * w + x] = m ? 1 : 0;
                if (m) atomicOr(&s_bits[col][y >> 5], 1u << (y & 31));
            }
        }
    }
    __syncthreads();
    // ---- B ----------------------------------------------------------------------------------------------------------
    if (t < kScanCols) {
        int last = -1;
        for (int wi = 0; wi < words; wi++) {
            s_prev[t][wi] = last;
            const unsigned b = s_bits[t][wi];
            if (b) last = (wi << 5) + 31 - __clz(b);
        }
    } else if (t < 2 * kScanCols) {
        const int c = t - kScanCols;
        int first = -1;
        for (int wi = words - 1; wi >= 0; wi--) {
            s_next[c][wi] = first;
            const unsigned b = s_bits[c][wi];
            if (b) first = (wi << 5) + __ffs(b) - 1;
        }
    }
    __syncthreads();
    // ---- C ----------------------------------------------------------------------------------------------------------
    if (!col_ok) return;
    const float *off = params + plane;
    constexpr int kC = 4;                                            // rows per thread in flight (independent gathers)
    for (int y0 = rslot; y0 < h; y0 += kC * kRowSlots) {
        ColHit ht[kC], hb[kC];
#pragma unroll
        for (int u = 0; u < kC; u++) {
            const int y = min(y0 + u * kRowSlots, h - 1);
            const int wi = y >> 5, b = y & 31;
            const unsigned word = s_bits[col][wi];
            const unsigned below = word & (0xFFFFFFFFu >> (31 - b)), above = word & (0xFFFFFFFFu << b);
            ht[u].y = below ? (wi << 5) + 31 - __clz(below) : s_prev[col][wi];
            hb[u].y = above ? (wi << 5) + __ffs(above) - 1 : s_next[col][wi];
        }
#pragma unroll
        for (int u = 0; u < kC; u++) {
            ht[u].v = __ldg(off + (long)max(ht[u].y, 0) * w + x);
            hb[u].v = __ldg(off + (long)max(hb[u].y, 0) * w + x);
        }
#pragma unroll
        for (int u = 0; u < kC; u++) {
            const int y = y0 + u * kRowSlots;
            if (y >= h) break;
            if (ht[u].y < 0) ht[u].v = 0.f;
            if (hb[u].y < 0) hb[u].v = 0.f;
            tb.top[(long)y * w + x] = ht[u];
            tb.bot[(long)y * w + x] = hb[u];
        }
    }
}

// ---- tables, large rasters ---------------------------------------------------------------------------------------------
// One CTA = 32 columns (lane = column, coalesced 128-byte rows) x 32 warps; warp w owns the row chunk
// [w*rc, (w+1)*rc).  Pass 1: every warp scans its chunk down and up, writes the r2_mask and publishes the chunk's
// last / first source pixel per column in shared memory.  Pass 2: every warp takes its carry-in from the nearest chunk
// above / below that has a source pixel and re-scans, writing the tables.
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
    ColHit c;
    c.y = -1; c.v = 0.f;
    for (int ww = warp - 1; ww >= 0; ww--)
        if (s_dn_y[ww][lane] >= 0) { c.y = s_dn_y[ww][lane]; c.v = s_dn_v[ww][lane]; break; }
    for (long y = ya; y < yb; y++) {
        const long i = y * w + x;
        if (tb.r2m[i]) { c.y = (int)y; c.v = params[plane + i]; }
        tb.top[i] = c;
    }
    c.y = -1; c.v = 0.f;
    for (int ww = warp + 1; ww < kScanWarps; ww++)
        if (s_up_y[ww][lane] >= 0) { c.y = s_up_y[ww][lane]; c.v = s_up_v[ww][lane]; break; }
    for (long y = yb - 1; y >= ya; y--) {
        const long i = y * w + x;
        if (tb.r2m[i]) { c.y = (int)y; c.v = params[plane + i]; }
        tb.bot[i] = c;
    }
}

#define HB_QUAD_CHECK(qd, qv, tx, hit)                                                                      \
    if ((hit).y >= 0) {                                                                                     \
        const double ddx = (double)(tx) - (double)x, ddy = (double)(hit).y - (double)y;                     \
        const double d2 = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));                              \
        if (d2 < __dmul_rn((qd), (qd))) { (qd) = sqrt(d2); (qv) = (double)(hit).v; }                        \
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
        const ColHit *top = tb.top + y * w, *bot = tb.bot + y * w;
        double qd0 = max_dist + 1.0, qd1 = qd0, qd2 = qd0, qd3 = qd0;
        double qv0 = 0.0, qv1 = 0.0, qv2 = 0.0, qv3 = 0.0;
        long this_max = max_dist_i;
        {                                                         // step 0: the centre column, left quadrants only
            const ColHit a = top[x], b = bot[x];
            HB_QUAD_CHECK(qd0, qv0, x, a)                         // top left (includes the current row)
            HB_QUAD_CHECK(qd1, qv1, x, b)                         // bottom left
        }
        // steps 1 .. this_max, eight at a time: the table entries of the eight steps are fetched together (the search
        // is a chain of dependent loads otherwise; entries beyond the current radius are read but not used).  GDAL
        // shrinks the search radius after every 4th step.
        constexpr int kSteps = 8;
        for (long s0 = 1; s0 <= this_max; s0 += kSteps) {
            ColHit tl[kSteps], bl[kSteps], tr[kSteps], br[kSteps];
#pragma unroll
            for (int u = 0; u < kSteps; u++) {
                const long step = s0 + u;
                const long lx = (x - step < 0) ? 0 : x - step;
                const long rx = (x + step > w - 1) ? w - 1 : x + step;
                tl[u] = top[lx]; bl[u] = bot[lx];
                tr[u] = top[rx]; br[u] = bot[rx];
            }
#pragma unroll
            for (int u = 0; u < kSteps; u++) {
                const long step = s0 + u;
                if (step > this_max) break;
                const long lx = (x - step < 0) ? 0 : x - step;
                const long rx = (x + step > w - 1) ? w - 1 : x + step;
                HB_QUAD_CHECK(qd0, qv0, lx, tl[u])                // top left
                HB_QUAD_CHECK(qd1, qv1, lx, bl[u])                // bottom left
                HB_QUAD_CHECK(qd2, qv2, rx, tr[u])                // top right
                HB_QUAD_CHECK(qd3, qv3, rx, br[u])                // bottom right
                if ((step & 0x3) == 0) {
                    const long lim = (long)floor(fmax(fmax(qd0, qd1), fmax(qd2, qd3)));
                    if (lim < this_max) this_max = lim;
                }
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
    HB_REQUIRE(((uintptr_t)workspace_dev) % 8 == 0, "hb_inpaint_refit: workspace must be 8-byte aligned");
    HB_REQUIRE(h < 2147483647L, "hb_inpaint_refit: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h * (size_t)w;
    FillTables tb;
    tb.top = (ColHit *)workspace_dev;
    tb.bot = tb.top + n;
    tb.r2m = (uint8_t *)(tb.bot + n);
    // numpy compares the float32 R2 plane with a Python float: the threshold is used as float32 (NEP 50)
    const float thresh = (float)r2_thresh;
    if (h <= 32L * kScanMaxWords) {
        inpaint_scan_small_kernel<<<(unsigned)((w + kScanCols - 1) / kScanCols), kScanThreads, 0, st>>>(
            params_dev, sums_dev, (int)h, w, thresh, tb);
        HB_LAUNCH_OK("inpaint_scan_small_kernel");
    } else {
        inpaint_scan_kernel<<<(unsigned)((w + 31) / 32), kScanWarps * 32, 0, st>>>(params_dev, sums_dev, h, w, thresh, tb);
        HB_LAUNCH_OK("inpaint_scan_kernel");
    }
    long blocks = ((long)n + 255) / 256;
    const long cap = (long)hb_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    inpaint_fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(params_dev, sums_dev, h, w, max_search_dist, tb);
    HB_LAUNCH_OK("inpaint_fill_kernel");
    return 0;
}
