// resample.cu -- grid-changing kernels of the kernel-model path (sm_100a):
//   * hb_downsample_average : RasterArray.reproject(resampling=average)     (reference raster_array.py:526-578,
//                             kernel_model.py:480)                            [HBM-bound: reads every hi-res pixel]
//   * hb_upsample_apply     : RefSpaceModel.apply = cubic-spline up-sampling of (gain, offset) fused with
//                             corr = gain*src + offset                        (kernel_model.py:484-503, 442-463)
//   * hb_resample_up        : plain cubic-spline / nearest up-sampling        (kernel_model.py:491, 497, 520)
//   * hb_apply_same_grid, hb_valid_mask, hb_full_coverage_mask                (kernel_model.py:375-409, 442-463)
//
// The GDAL algorithms restated here are specified in oracle/gdal_restate.c (GDAL itself is not in this image).
#include "hb_common.cuh"

namespace {

constexpr int kThreads = 256;

// =====================================================================================================================
// 1. average down-sampling
// =====================================================================================================================
constexpr int kDsVec = 8;                        // source pixels per thread per row
constexpr int kDsSpan = kThreads * kDsVec;       // source columns staged per CTA

template <typename T> struct Raw8;               // 8 consecutive source pixels, vector-loaded
template <> struct Raw8<uint16_t> {
    static __device__ __forceinline__ void load(const uint16_t *p, uint16_t (&v)[8])
    {
        const uint4 w = hb_ldg_stream16(p);
        v[0] = (uint16_t)(w.x & 0xFFFFu); v[1] = (uint16_t)(w.x >> 16);
        v[2] = (uint16_t)(w.y & 0xFFFFu); v[3] = (uint16_t)(w.y >> 16);
        v[4] = (uint16_t)(w.z & 0xFFFFu); v[5] = (uint16_t)(w.z >> 16);
        v[6] = (uint16_t)(w.w & 0xFFFFu); v[7] = (uint16_t)(w.w >> 16);
    }
};
template <> struct Raw8<uint8_t> {
    static __device__ __forceinline__ void load(const uint8_t *p, uint8_t (&v)[8])
    {
        const uint2 w = hb_ldg_stream8(p);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[k] = (uint8_t)((w.x >> (8 * k)) & 0xFFu);
            v[4 + k] = (uint8_t)((w.y >> (8 * k)) & 0xFFu);
        }
    }
};
template <> struct Raw8<float> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[8])
    {
        const uint4 a = hb_ldg_stream16(p), b = hb_ldg_stream16(p + 4);
        v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z);
        v[3] = __uint_as_float(a.w); v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y);
        v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
    }
};

template <typename T> struct IsInt { static constexpr bool value = true; };
template <> struct IsInt<float> { static constexpr bool value = false; };

template <typename T> __device__ __forceinline__ bool ds_valid(T v, const NoData &nd)
{
    if (IsInt<T>::value) return hb_valid_int((uint32_t)v, nd);
    return hb_valid((float)v, nd);
}

// fetch 8 pixels of row `row` starting at column c (may be partially / wholly outside [0, ws))
template <typename T, bool ALIGNED>
__device__ __forceinline__ void ds_fetch(const T *row, long c, long ws, T (&v)[8], uint32_t &inb)
{
    if (ALIGNED && c >= 0 && c + 8 <= ws) {
        Raw8<T>::load(row + c, v);
        inb = 0xFFu;
    } else {
        inb = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const long x = c + k;
            if (x >= 0 && x < ws) { v[k] = row[x]; inb |= (1u << k); } else v[k] = (T)0;
        }
    }
}

// One CTA: one destination row, `ndc` destination columns.  Phase 1: every thread owns 8 source columns and sums them
// down the footprint rows (interior rows have weight 1: exact integer sums for uint8/uint16, double for float32; the
// <= 2 fractionally covered rows are added in double).  Phase 2: one thread per destination pixel combines its
// footprint columns from shared memory with the fractional edge weights.
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(kThreads)
downsample_average_kernel(const T *__restrict__ src, long hs, long ws, NoData nd, float *__restrict__ dst, long hd,
                          long wd, double sx, double ox, double sy, double oy, int ndc, int chunks)
{
    __shared__ double s_sum[kDsSpan];
    __shared__ double s_w[kDsSpan];

    const long i = blockIdx.x / chunks;
    const int chunk = (int)(blockIdx.x % chunks);
    const long j0 = (long)chunk * ndc;
    const long j1 = min(j0 + (long)ndc, wd);
    const int t = threadIdx.x;
    const float qnan = __int_as_float(0x7fc00000);

    // Footprint rows.  Weights follow GDAL's COMPUTE_WEIGHT_Y on the UNCLAMPED footprint and rows outside the raster
    // are skipped -- i.e. the result equals GDAL's on a source padded with nodata, which is how the reference reads
    // its source blocks (boundless windows, raster_array.py:175-199).
    const double y_min = sy * (double)i + oy, y_max = sy * (double)(i + 1) + oy;
    const long iy0u = (long)floor(y_min + 1e-10);
    long iy1u = (long)ceil(y_max - 1e-10);
    if (iy0u == iy1u) iy1u++;                                   // footprint inside one source row
    const long iy0 = max(iy0u, 0L), iy1 = min(iy1u, hs);
    if (iy0 >= iy1) {
        for (long j = j0 + t; j < j1; j += kThreads) dst[i * wd + j] = qnan;
        return;
    }
    const double wy_first = (iy0u + 1 == iy1u) ? 1.0 : 1.0 - (y_min - (double)iy0u);
    const double wy_last = 1.0 - ((double)iy1u - y_max);
    const bool first_frac = (iy0u >= 0) && (wy_first != 1.0);
    const bool last_frac = (iy1u <= hs) && (iy0u + 1 < iy1u) && (wy_last != 1.0);
    const long ya = iy0 + (first_frac ? 1 : 0), yb = iy1 - (last_frac ? 1 : 0);   // interior rows [ya, yb)

    const double xs_min = sx * (double)j0 + ox;
    long c_lo = (long)floor(xs_min + 1e-10);
    if (c_lo < 0) c_lo = 0;
    const long c_base = (c_lo / kDsVec) * kDsVec;
    const long c = c_base + (long)t * kDsVec;

    const NoData ndk = nd;

    // ---- phase 1: column sums over interior rows ------------------------------------------------------------------
    double colsum[8], colw[8];
    if (IsInt<T>::value) {
        uint32_t isum[8], icnt[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { isum[k] = 0; icnt[k] = 0; }
#pragma unroll 4
        for (long y = ya; y < yb; y++) {
            T v[8]; uint32_t inb;
            ds_fetch<T, ALIGNED>(src + y * ws, c, ws, v, inb);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const bool ok = ((inb >> k) & 1u) && ds_valid<T>(v[k], ndk);
                isum[k] += ok ? (uint32_t)v[k] : 0u;
                icnt[k] += ok ? 1u : 0u;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) { colsum[k] = (double)isum[k]; colw[k] = (double)icnt[k]; }
    } else {
        uint32_t icnt[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { colsum[k] = 0.0; icnt[k] = 0; }
#pragma unroll 4
        for (long y = ya; y < yb; y++) {
            T v[8]; uint32_t inb;
            ds_fetch<T, ALIGNED>(src + y * ws, c, ws, v, inb);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const bool ok = ((inb >> k) & 1u) && ds_valid<T>(v[k], ndk);
                colsum[k] += ok ? (double)v[k] : 0.0;
                icnt[k] += ok ? 1u : 0u;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) colw[k] = (double)icnt[k];
    }
    // fractionally covered first / last rows
    if (first_frac || last_frac) {
#pragma unroll 1
        for (int e = 0; e < 2; e++) {
            if ((e == 0 && !first_frac) || (e == 1 && !last_frac)) continue;
            const long y = (e == 0) ? iy0u : iy1u - 1;
            const double wy = (e == 0) ? wy_first : wy_last;
            T v[8]; uint32_t inb;
            ds_fetch<T, ALIGNED>(src + y * ws, c, ws, v, inb);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const bool ok = ((inb >> k) & 1u) && ds_valid<T>(v[k], ndk);
                if (ok) { colsum[k] += wy * (double)v[k]; colw[k] += wy; }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) { s_sum[t * kDsVec + k] = colsum[k]; s_w[t * kDsVec + k] = colw[k]; }
    __syncthreads();

    // ---- phase 2: one thread per destination pixel ----------------------------------------------------------------
    for (long j = j0 + t; j < j1; j += kThreads) {
        const double x_min = sx * (double)j + ox, x_max = sx * (double)(j + 1) + ox;
        const long ix0u = (long)floor(x_min + 1e-10);
        long ix1u = (long)ceil(x_max - 1e-10);
        if (ix0u == ix1u) ix1u++;
        const long ix0 = max(ix0u, 0L), ix1 = min(ix1u, ws);
        float out = qnan;
        double total = 0.0, total_w = 0.0;
        for (long x = ix0; x < ix1; x++) {
            double wx = 1.0;
            if (x == ix0u) wx = (ix0u + 1 == ix1u) ? 1.0 : 1.0 - (x_min - (double)ix0u);
            else if (x + 1 == ix1u) wx = 1.0 - ((double)ix1u - x_max);
            const int s = (int)(x - c_base);
            total += wx * s_sum[s];
            total_w += wx * s_w[s];
        }
        if (total_w > 0.0) out = (float)(total / total_w);
        dst[i * wd + j] = out;
    }
}

template <typename T>
int launch_downsample(const void *src, long hs, long ws, NoData nd, float *dst, long hd, long wd, double sx, double ox,
                      double sy, double oy, cudaStream_t stream)
{
    // destination columns per CTA so that their footprint (plus alignment slack) fits the staged span
    long ndc = (long)floor(((double)kDsSpan - kDsVec - 2.0) / sx);
    HB_REQUIRE(ndc >= 1, "hb_downsample_average: down-sampling ratio %.1f is too large (max %d)", sx, kDsSpan - 10);
    if (ndc > wd) ndc = wd;
    const long chunks = (wd + ndc - 1) / ndc;
    const long blocks = hd * chunks;
    HB_REQUIRE(blocks > 0 && blocks < 2147483647L, "hb_downsample_average: grid too large");
    const bool aligned = ((ws * (long)sizeof(T)) % 16 == 0) && (((uintptr_t)src) % 16 == 0);
    if (aligned)
        downsample_average_kernel<T, true><<<(unsigned)blocks, kThreads, 0, stream>>>(
            (const T *)src, hs, ws, nd, dst, hd, wd, sx, ox, sy, oy, (int)ndc, (int)chunks);
    else
        downsample_average_kernel<T, false><<<(unsigned)blocks, kThreads, 0, stream>>>(
            (const T *)src, hs, ws, nd, dst, hd, wd, sx, ox, sy, oy, (int)ndc, (int)chunks);
    HB_LAUNCH_OK("downsample_average_kernel");
    return 0;
}

// =====================================================================================================================
// 2. cubic-spline up-sampling, optionally fused with the apply step
// =====================================================================================================================
// Per destination row the 4x4 B-spline interpolation is separated: (A) each coarse column is combined down its 4 tap
// rows with the row's y-weights (invalid / out-of-range taps dropped, their weight tracked), (B) each coarse cell gets
// the cubic polynomial in dx = frac(x) of its 4 tap columns -- value polynomials for both bands and, only when a tap
// is missing, weight polynomials for GDAL's renormalisation -- and (C) every destination pixel evaluates its cell's
// polynomial by Horner, fused with corr = gain*src + offset.  A/B for rows y+2 / y+1 overlap C for row y, one
// __syncthreads per row.
struct __align__(16) CellPoly {
    double g[4];   // band 0 (gain) polynomial
    double o[4];   // band 1 (offset) polynomial
};
struct __align__(16) ColComb {
    double ag, ao;   // sum_j wy_j * v_j over valid taps, per band
    double mg, mo;   // sum_j wy_j over valid taps, per band
};

constexpr int kUpRows = 32;   // destination rows per CTA

__device__ __forceinline__ double bspline_w(int tap, double d)   // tap in {-1,0,1,2}, weight B(tap - d)
{
    const double u = 1.0 - d;
    switch (tap) {
        case -1: return u * u * u * (1.0 / 6.0);
        case 0: return (4.0 + d * d * (3.0 * d - 6.0)) * (1.0 / 6.0);
        case 1: return (1.0 + d * (3.0 + d * (3.0 - 3.0 * d))) * (1.0 / 6.0);
        default: return d * d * d * (1.0 / 6.0);
    }
}

__device__ __forceinline__ void poly_from_taps(double a_m1, double a_0, double a_1, double a_2, double (&p)[4])
{
    p[0] = (a_m1 + 4.0 * a_0 + a_1) * (1.0 / 6.0);
    p[1] = (a_1 - a_m1) * 0.5;
    p[2] = (a_m1 - 2.0 * a_0 + a_1) * 0.5;
    p[3] = ((a_2 - a_m1) + 3.0 * (a_0 - a_1)) * (1.0 / 6.0);
}

__device__ __forceinline__ double horner3(const double (&p)[4], double d)
{
    return fma(fma(fma(p[3], d, p[2]), d, p[1]), d, p[0]);
}

template <typename T> struct SrcVec4;   // 4 consecutive source pixels -> float32
template <> struct SrcVec4<uint16_t> {
    static __device__ __forceinline__ void load(const uint16_t *p, float (&v)[4])
    {
        const uint2 w = hb_ldg_stream8(p);
        v[0] = hb_to_f32<uint16_t>((uint16_t)(w.x & 0xFFFFu)); v[1] = hb_to_f32<uint16_t>((uint16_t)(w.x >> 16));
        v[2] = hb_to_f32<uint16_t>((uint16_t)(w.y & 0xFFFFu)); v[3] = hb_to_f32<uint16_t>((uint16_t)(w.y >> 16));
    }
    static constexpr int kAlign = 8;
};
template <> struct SrcVec4<uint8_t> {
    static __device__ __forceinline__ void load(const uint8_t *p, float (&v)[4])
    {
        const uint32_t w = hb_ldg_stream4(p);
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = hb_to_f32<uint8_t>((uint8_t)((w >> (8 * k)) & 0xFFu));
    }
    static constexpr int kAlign = 4;
};
template <> struct SrcVec4<float> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4])
    {
        const uint4 w = hb_ldg_stream16(p);
        v[0] = __uint_as_float(w.x); v[1] = __uint_as_float(w.y); v[2] = __uint_as_float(w.z);
        v[3] = __uint_as_float(w.w);
    }
    static constexpr int kAlign = 16;
};

struct UpGeom {
    long hs, ws;          // destination (fine) grid
    long hp, wp;          // coarse grid
    double sx, ox, sy, oy;
    int ncols;            // coarse columns staged per CTA (cells + 3)
    int tile_w;           // destination columns per CTA
};

// T: storage type of the source plane (APPLY); NB: number of coarse bands (1 or 2); APPLY: fuse gain*src+offset
template <typename T, int NB, bool APPLY, int PPT, bool ALIGNED>
__global__ void __launch_bounds__(kThreads)
upsample_kernel(const T *__restrict__ src, NoData nd, const float *__restrict__ coarse, UpGeom g,
                const uint8_t *__restrict__ cover, float *__restrict__ out)
{
    constexpr int NOUT = (NB == 2 && !APPLY) ? 2 : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ncols = g.ncols;
    ColComb *s_col = reinterpret_cast<ColComb *>(smem_raw);            // [2][ncols]  stage A -> B
    CellPoly *s_poly = reinterpret_cast<CellPoly *>(s_col + 2 * ncols);   // [2][ncols]  stage B -> C  (values)
    CellPoly *s_wpoly = s_poly + 2 * ncols;                              // [2][ncols]  stage B -> C  (weights)
    uint8_t *s_miss = reinterpret_cast<uint8_t *>(s_wpoly + 2 * ncols);  // [2][ncols]  A -> B: column has a missing tap
    uint8_t *s_cokA = s_miss + 2 * ncols;                                // [2][ncols]  A -> B: centre pixel usable
    uint8_t *s_norm = s_cokA + 2 * ncols;                                // [2][ncols]  B -> C: cell needs renormalising
    uint8_t *s_cok = s_norm + 2 * ncols;                                 // [2][ncols]  B -> C: centre pixel usable

    const int t = threadIdx.x;
    const long Xt0 = (long)blockIdx.x * g.tile_w;                  // first destination column of the tile
    const long X0 = Xt0 + (long)t * PPT;                           // first destination column of this thread
    const bool px_thread = (t * PPT < g.tile_w) && (X0 < g.ws);
    const long Y0 = (long)blockIdx.y * kUpRows;
    const long Y1 = min(Y0 + (long)kUpRows, g.hs);
    const long plane = g.hp * g.wp;
    const float qnan = __int_as_float(0x7fc00000);

    // coarse column window of the tile: the taps of the tile's first pixel start at column kx - 1
    const double srcx_t0 = g.sx * ((double)Xt0 + 0.5) + g.ox;
    const long col_base = (long)floor(srcx_t0 - 0.5) - 1;

    // per-pixel, row-invariant: cell (index of its first tap column in the window), dx, centre column in the window
    int cell[PPT], ccol[PPT];
    double dx[PPT];
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        const double srcx = g.sx * ((double)(X0 + k) + 0.5) + g.ox;
        const long kx = (long)floor(srcx - 0.5);
        dx[k] = srcx - 0.5 - (double)kx;
        long cl = kx - 1 - col_base;
        long cx = (long)floor(srcx + 1e-10);
        if (cx == g.wp) cx--;
        const bool okx = (srcx >= 0.0) && (cx >= 0) && (cx < g.wp);
        long cc = okx ? (cx - col_base) : -1;
        if (cl < 0 || cl + 3 >= ncols || cc >= ncols) { cl = 0; cc = -1; }   // outside the staged window: X >= ws
        cell[k] = (int)cl;
        ccol[k] = (int)cc;
    }

    // ---- stage A: combine coarse column (col_base + t) down the 4 tap rows of destination row Y -------------------
    // the 4 tap rows are cached in registers and shifted when ky advances
    float tap_v[NB][4];
    bool tap_ok[NB][4];
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
        for (int j = 0; j < 4; j++) { tap_v[b][j] = qnan; tap_ok[b][j] = false; }
    long cached_ky = -(1L << 60);
    const long my_col = col_base + t;
    const bool col_thread = (t < ncols);
    const bool col_inside = col_thread && my_col >= 0 && my_col < g.wp;

    auto stage_a = [&](long Y, int buf) {
        if (!col_thread) return;
        const double srcy = g.sy * ((double)Y + 0.5) + g.oy;
        const long ky = (long)floor(srcy - 0.5);
        const double dy = srcy - 0.5 - (double)ky;
        if (ky != cached_ky) {
            const bool shift1 = (ky - cached_ky) == 1;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const long row = ky - 1 + j;
                const bool inside = col_inside && row >= 0 && row < g.hp;
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    if (shift1 && j < 3) {
                        tap_v[b][j] = tap_v[b][j + 1];
                        tap_ok[b][j] = tap_ok[b][j + 1];
                    } else {
                        const float v = inside ? __ldg(coarse + b * plane + row * g.wp + my_col) : qnan;
                        tap_v[b][j] = v;
                        tap_ok[b][j] = inside && !isnan(v);
                    }
                }
            }
            cached_ky = ky;
        }
        ColComb cc = {0.0, 0.0, 0.0, 0.0};
        bool all_ok = true;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double wy = bspline_w(j - 1, dy);
            // a band's tap counts iff that band is valid there (band-valid implies unified "any band" validity)
            if (tap_ok[0][j]) { cc.ag = fma(wy, (double)tap_v[0][j], cc.ag); cc.mg += wy; } else all_ok = false;
            if (NB > 1) {
                if (tap_ok[1][j]) { cc.ao = fma(wy, (double)tap_v[1][j], cc.ao); cc.mo += wy; } else all_ok = false;
            }
        }
        s_col[buf * ncols + t] = cc;
        s_miss[buf * ncols + t] = all_ok ? 0 : 1;
        // centre pixel: the coarse pixel containing the destination centre must be in range and valid in any band
        long cy = (long)floor(srcy + 1e-10);
        if (cy == g.hp) cy--;
        bool cok = (srcy >= 0.0) && cy >= 0 && cy < g.hp && col_inside;
        if (cok) {
            // cy is ky or ky + 1 (tap row 1 or 2), or ky - 1 (tap row 0) when clamped at the bottom edge
            const int j = (int)(cy - (ky - 1));
            bool any = false;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                if (jj == j) {
                    any = tap_ok[0][jj];
                    if (NB > 1) any = any || tap_ok[1][jj];
                }
            }
            cok = any;
            if (cok && cover != nullptr) cok = cover[cy * g.wp + my_col] != 0;
        }
        s_cokA[buf * ncols + t] = cok ? 1 : 0;
    };

    // ---- stage B: polynomials of the cell whose taps are staged columns t .. t+3 -----------------------------------
    auto stage_b = [&](int buf) {
        if (!col_thread) return;
        s_cok[buf * ncols + t] = s_cokA[buf * ncols + t];
        if (t + 3 >= ncols) return;
        const ColComb *c = s_col + buf * ncols + t;
        CellPoly p;
        poly_from_taps(c[0].ag, c[1].ag, c[2].ag, c[3].ag, p.g);
        if (NB > 1) poly_from_taps(c[0].ao, c[1].ao, c[2].ao, c[3].ao, p.o);
        else { p.o[0] = p.o[1] = p.o[2] = p.o[3] = 0.0; }
        s_poly[buf * ncols + t] = p;
        const uint8_t *f = s_miss + buf * ncols + t;
        const bool missing = (f[0] | f[1] | f[2] | f[3]) != 0;
        if (missing) {
            CellPoly w;
            poly_from_taps(c[0].mg, c[1].mg, c[2].mg, c[3].mg, w.g);
            if (NB > 1) poly_from_taps(c[0].mo, c[1].mo, c[2].mo, c[3].mo, w.o);
            else { w.o[0] = w.o[1] = w.o[2] = w.o[3] = 0.0; }
            s_wpoly[buf * ncols + t] = w;
        }
        s_norm[buf * ncols + t] = missing ? 1 : 0;
    };

    // software pipeline prologue
    stage_a(Y0, 0);
    __syncthreads();
    stage_b(0);
    if (Y0 + 1 < Y1) stage_a(Y0 + 1, 1);
    __syncthreads();

    for (long Y = Y0; Y < Y1; Y++) {
        const int buf = (int)((Y - Y0) & 1);
        // ---- stage C: destination pixels of row Y -------------------------------------------------------------------
        if (px_thread) {
            float s[PPT];
            bool in_row[PPT];
            const bool full_vec = (X0 + PPT <= g.ws);
            if constexpr (APPLY) {
                const T *row = src + Y * g.ws;
                bool loaded = false;
                if constexpr (ALIGNED && PPT == 4) {
                    if (full_vec) {
                        SrcVec4<T>::load(row + X0, s);
#pragma unroll
                        for (int k = 0; k < PPT; k++) in_row[k] = true;
                        loaded = true;
                    }
                }
                if (!loaded) {
#pragma unroll
                    for (int k = 0; k < PPT; k++) {
                        in_row[k] = (X0 + k) < g.ws;
                        s[k] = in_row[k] ? hb_to_f32<T>(row[X0 + k]) : 0.f;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < PPT; k++) { in_row[k] = (X0 + k) < g.ws; s[k] = 0.f; }
            }
            float res[NOUT][PPT];
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                float r0 = qnan, r1 = qnan;
                bool ok = in_row[k] && ccol[k] >= 0 && s_cok[buf * ncols + (ccol[k] < 0 ? 0 : ccol[k])];
                if (APPLY) ok = ok && hb_valid(s[k], nd);
                if (ok) {
                    const CellPoly &p = s_poly[buf * ncols + cell[k]];
                    double gv = horner3(p.g, dx[k]);
                    double ov = (NB > 1) ? horner3(p.o, dx[k]) : 0.0;
                    bool g_ok = true, o_ok = true;
                    if (s_norm[buf * ncols + cell[k]]) {
                        // GDAL GWKResample: drop if sum(w) < 1e-6, divide unless sum(w) is within 1e-5 of 1
                        const CellPoly &w = s_wpoly[buf * ncols + cell[k]];
                        const double wg = horner3(w.g, dx[k]);
                        if (wg < 0.000001) g_ok = false;
                        else if (wg < 0.99999 || wg > 1.00001) gv /= wg;
                        if (NB > 1) {
                            const double wo = horner3(w.o, dx[k]);
                            if (wo < 0.000001) o_ok = false;
                            else if (wo < 0.99999 || wo > 1.00001) ov /= wo;
                        }
                    }
                    const float gf = g_ok ? (float)gv : qnan;
                    const float of = o_ok ? (float)ov : qnan;
                    if (APPLY) {
                        r0 = __fadd_rn(__fmul_rn(gf, s[k]), of);   // two roundings, as numpy (kernel_model.py:461)
                    } else {
                        r0 = gf;
                        r1 = of;
                    }
                }
                res[0][k] = r0;
                if constexpr (NOUT == 2) res[1][k] = r1;
            }
            float *orow = out + Y * g.ws;
            bool stored = false;
            if constexpr (ALIGNED && PPT == 4) {
                if (full_vec) {
                    hb_stg_stream16(orow + X0, make_float4(res[0][0], res[0][1], res[0][2], res[0][3]));
                    if constexpr (NOUT == 2)
                        hb_stg_stream16(orow + g.hs * g.ws + X0,
                                        make_float4(res[1][0], res[1][1], res[1][2], res[1][3]));
                    stored = true;
                }
            }
            if (!stored) {
#pragma unroll
                for (int k = 0; k < PPT; k++) {
                    if (in_row[k]) {
                        orow[X0 + k] = res[0][k];
                        if constexpr (NOUT == 2) orow[g.hs * g.ws + X0 + k] = res[1][k];
                    }
                }
            }
        }
        // ---- overlap: polynomials for row Y+1, column combination for row Y+2 ----------------------------------------
        if (Y + 1 < Y1) stage_b(buf ^ 1);
        if (Y + 2 < Y1) stage_a(Y + 2, buf);
        __syncthreads();
    }
}

template <typename T, int NB, bool APPLY>
int launch_upsample(const void *src, NoData nd, const float *coarse, long hs, long ws, long hp, long wp, double sx,
                    double ox, double sy, double oy, const uint8_t *cover, float *out, cudaStream_t stream)
{
    HB_REQUIRE(sx > 0 && sy > 0 && sx <= 1.0 + 1e-9 && sy <= 1.0 + 1e-9,
               "cubic-spline up-sampling needs a destination grid at least as fine as the source (scale %.4f, %.4f)",
               sx, sy);
    // 4 pixels per thread when the coarse window of a 1024-pixel tile fits the 256 staging threads, else 1
    const bool wide = (1024.0 * sx + 6.0) <= (double)kThreads;
    UpGeom g;
    g.hs = hs; g.ws = ws; g.hp = hp; g.wp = wp; g.sx = sx; g.ox = ox; g.sy = sy; g.oy = oy;
    g.tile_w = wide ? kThreads * 4 : kThreads - 6;
    g.ncols = (int)ceil((double)g.tile_w * sx) + 5;
    HB_REQUIRE(g.ncols <= kThreads, "up-sampling tile does not fit its coarse window");
    const size_t smem = (size_t)g.ncols * 2 * (sizeof(ColComb) + 2 * sizeof(CellPoly) + 4);
    dim3 grid((unsigned)((ws + g.tile_w - 1) / g.tile_w), (unsigned)((hs + kUpRows - 1) / kUpRows));
    HB_REQUIRE(grid.y <= 65535u, "up-sampling destination has too many rows (%ld)", hs);
    const size_t align = APPLY ? SrcVec4<T>::kAlign : 16;
    const bool aligned = wide && (!APPLY || (((ws * (long)sizeof(T)) % (long)align == 0) && (((uintptr_t)src) % align == 0))) &&
                         (ws % 4 == 0) && (((uintptr_t)out) % 16 == 0);
#define HB_UP_LAUNCH(PPT_, AL_)                                                                                       \
    do {                                                                                                              \
        auto kern = upsample_kernel<T, NB, APPLY, PPT_, AL_>;                                                         \
        if (smem > 48 * 1024)                                                                                         \
            HB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        kern<<<grid, kThreads, smem, stream>>>((const T *)src, nd, coarse, g, cover, out);                            \
    } while (0)
    if (wide) {
        if (aligned) HB_UP_LAUNCH(4, true); else HB_UP_LAUNCH(4, false);
    } else {
        HB_UP_LAUNCH(1, false);
    }
#undef HB_UP_LAUNCH
    HB_LAUNCH_OK("upsample_kernel");
    return 0;
}

// nearest-neighbour up-sampling of float planes (kernel_model.py:497 semantics for float data)
__global__ void nearest_kernel(const float *__restrict__ src, long nb, long hs, long ws, NoData nd,
                               float *__restrict__ dst, long hd, long wd, double sx, double ox, double sy, double oy)
{
    const long n = hd * wd;
    const float qnan = __int_as_float(0x7fc00000);
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx / wd, j = idx - i * wd;
        const double srcy = sy * ((double)i + 0.5) + oy, srcx = sx * ((double)j + 0.5) + ox;
        long cy = (long)floor(srcy + 1e-10), cx = (long)floor(srcx + 1e-10);
        if (cy == hs) cy--;
        if (cx == ws) cx--;
        const bool ok = srcy >= 0.0 && srcx >= 0.0 && cy >= 0 && cy < hs && cx >= 0 && cx < ws;
        for (long b = 0; b < nb; b++) {
            float v = qnan;
            if (ok) {
                const float s = src[b * hs * ws + cy * ws + cx];
                if (hb_valid(s, nd)) v = s;
            }
            dst[b * n + idx] = v;
        }
    }
}

// =====================================================================================================================
// 3. element-wise kernels
// =====================================================================================================================
template <typename T>
__global__ void apply_same_grid_kernel(const T *__restrict__ src, NoData nd, int mask_src,
                                       const float *__restrict__ params, long n, float *__restrict__ corr)
{
    const float qnan = __int_as_float(0x7fc00000);
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float s = hb_to_f32<T>(src[i]);
        float r = __fadd_rn(__fmul_rn(__ldg(params + i), s), __ldg(params + n + i));   // kernel_model.py:461
        if (mask_src && !hb_valid(s, nd)) r = qnan;
        corr[i] = r;
    }
}

template <typename T>
__global__ void valid_mask_kernel(const T *__restrict__ src, long n, NoData nd, uint8_t *__restrict__ mask)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        mask[i] = hb_valid(hb_to_f32<T>(src[i]), nd) ? 1 : 0;
}

// ---- full-coverage mask (kernel_model.py:375-409) -------------------------------------------------------------------
// step 1: param pixel is "covered" iff it is valid and every in-range pixel of the other image's mask under its
//         footprint is valid (== GDAL average of the 0/1 mask is exactly 1) and the footprint is not empty.
__global__ void coverage_kernel(const uint8_t *__restrict__ in_mask, long hi, long wi, const float *__restrict__ params,
                                long hp, long wp, double sx, double ox, double sy, double oy,
                                uint8_t *__restrict__ covered)
{
    const long n = hp * wp;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx / wp, j = idx - i * wp;
        bool ok = !(isnan(params[idx]) && isnan(params[n + idx]));
        if (ok) {
            const double y_min = sy * (double)i + oy, y_max = sy * (double)(i + 1) + oy;
            const double x_min = sx * (double)j + ox, x_max = sx * (double)(j + 1) + ox;
            long iy0 = (long)floor(y_min + 1e-10), iy1 = (long)ceil(y_max - 1e-10);
            long ix0 = (long)floor(x_min + 1e-10), ix1 = (long)ceil(x_max - 1e-10);
            if (iy0 == iy1) iy1++;
            if (ix0 == ix1) ix1++;
            // a footprint that leaves the other image's raster is covered by nodata there (boundless read)
            if (iy0 < 0 || ix0 < 0 || iy1 > hi || ix1 > wi) ok = false;
            for (long y = iy0; ok && y < iy1; y++)
                for (long x = ix0; x < ix1; x++)
                    if (!in_mask[y * wi + x]) { ok = false; break; }
        }
        covered[idx] = ok ? 1 : 0;
    }
}

// step 2: erosion with a (kh+2) x (kw+2) rectangle, zero border (cv.erode BORDER_CONSTANT 0, kernel_model.py:407-408)
__global__ void erode_kernel(const uint8_t *__restrict__ in, long h, long w, int eh, int ew, uint8_t *__restrict__ out)
{
    const long n = h * w;
    const int ry = eh / 2, rx = ew / 2;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx / w, j = idx - i * w;
        bool ok = (i - ry >= 0) && (i + ry < h) && (j - rx >= 0) && (j + rx < w);
        for (long y = i - ry; ok && y <= i + ry; y++)
            for (long x = j - rx; x <= j + rx; x++)
                if (!in[y * w + x]) { ok = false; break; }
        out[idx] = ok ? 1 : 0;
    }
}

inline unsigned grid_for(long n, int threads, int per_sm = 8)
{
    long blocks = (n + threads - 1) / threads;
    const long cap = (long)hb_sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

// =====================================================================================================================
// C-ABI
// =====================================================================================================================
extern "C" int hb_downsample_average(const void *src_dev, int src_dtype, long hs, long ws, int has_nodata,
                                     double nodata, float *dst_dev, long hd, long wd, double sx, double ox, double sy,
                                     double oy, void *stream)
{
    HB_REQUIRE(src_dev && dst_dev && hs > 0 && ws > 0 && hd > 0 && wd > 0, "hb_downsample_average: bad arguments");
    HB_REQUIRE(sx > 0 && sy > 0, "hb_downsample_average: grids must have the same orientation");
    const NoData nd = hb_make_nodata(has_nodata, nodata);
    cudaStream_t st = (cudaStream_t)stream;
    switch (src_dtype) {
        case HB_U8: return launch_downsample<uint8_t>(src_dev, hs, ws, nd, dst_dev, hd, wd, sx, ox, sy, oy, st);
        case HB_U16: return launch_downsample<uint16_t>(src_dev, hs, ws, nd, dst_dev, hd, wd, sx, ox, sy, oy, st);
        case HB_F32: return launch_downsample<float>(src_dev, hs, ws, nd, dst_dev, hd, wd, sx, ox, sy, oy, st);
    }
    HB_REQUIRE(false, "hb_downsample_average: unknown dtype %d", src_dtype);
}

extern "C" int hb_upsample_apply(const void *src_dev, int src_dtype, long hs, long ws, int has_nodata, double nodata,
                                 const float *params_dev, long hp, long wp, double sx, double ox, double sy, double oy,
                                 const uint8_t *cover_dev, float *corr_dev, void *stream)
{
    HB_REQUIRE(src_dev && params_dev && corr_dev && hs > 0 && ws > 0 && hp > 0 && wp > 0,
               "hb_upsample_apply: bad arguments");
    const NoData nd = hb_make_nodata(has_nodata, nodata);
    cudaStream_t st = (cudaStream_t)stream;
    switch (src_dtype) {
        case HB_U8:
            return launch_upsample<uint8_t, 2, true>(src_dev, nd, params_dev, hs, ws, hp, wp, sx, ox, sy, oy, cover_dev,
                                                     corr_dev, st);
        case HB_U16:
            return launch_upsample<uint16_t, 2, true>(src_dev, nd, params_dev, hs, ws, hp, wp, sx, ox, sy, oy,
                                                      cover_dev, corr_dev, st);
        case HB_F32:
            return launch_upsample<float, 2, true>(src_dev, nd, params_dev, hs, ws, hp, wp, sx, ox, sy, oy, cover_dev,
                                                   corr_dev, st);
    }
    HB_REQUIRE(false, "hb_upsample_apply: unknown dtype %d", src_dtype);
}

extern "C" int hb_resample_up(const float *src_dev, long nb, long hs, long ws, int has_nodata, double nodata,
                              float *dst_dev, long hd, long wd, double sx, double ox, double sy, double oy, int method,
                              void *stream)
{
    HB_REQUIRE(src_dev && dst_dev && nb >= 1 && hs > 0 && ws > 0 && hd > 0 && wd > 0, "hb_resample_up: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const NoData nd = hb_make_nodata(has_nodata, nodata);
    if (method == HB_UP_NEAREST) {
        nearest_kernel<<<grid_for(hd * wd, 256), 256, 0, st>>>(src_dev, nb, hs, ws, nd, dst_dev, hd, wd, sx, ox, sy,
                                                               oy);
        HB_LAUNCH_OK("nearest_kernel");
        return 0;
    }
    HB_REQUIRE(method == HB_UP_CUBIC_SPLINE, "hb_resample_up: unknown method %d", method);
    HB_REQUIRE(has_nodata && isnan(nodata), "hb_resample_up: cubic-spline input must use NaN as nodata");
    HB_REQUIRE(nb <= 2, "hb_resample_up: cubic-spline supports 1 or 2 bands per call (got %ld)", nb);
    const NoData none = hb_make_nodata(0, 0.0);
    if (nb == 1)
        return launch_upsample<float, 1, false>(nullptr, none, src_dev, hd, wd, hs, ws, sx, ox, sy, oy, nullptr,
                                                dst_dev, st);
    return launch_upsample<float, 2, false>(nullptr, none, src_dev, hd, wd, hs, ws, sx, ox, sy, oy, nullptr, dst_dev,
                                            st);
}

extern "C" int hb_apply_same_grid(const void *src_dev, int src_dtype, int has_nodata, double nodata, int mask_src,
                                  const float *params_dev, long h, long w, float *corr_dev, void *stream)
{
    HB_REQUIRE(src_dev && params_dev && corr_dev && h > 0 && w > 0, "hb_apply_same_grid: bad arguments");
    const NoData nd = hb_make_nodata(has_nodata, nodata);
    cudaStream_t st = (cudaStream_t)stream;
    const long n = h * w;
    const unsigned grid = grid_for(n, 256, 16);
    switch (src_dtype) {
        case HB_U8:
            apply_same_grid_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)src_dev, nd, mask_src, params_dev, n,
                                                                  corr_dev);
            break;
        case HB_U16:
            apply_same_grid_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)src_dev, nd, mask_src, params_dev,
                                                                   n, corr_dev);
            break;
        case HB_F32:
            apply_same_grid_kernel<float><<<grid, 256, 0, st>>>((const float *)src_dev, nd, mask_src, params_dev, n,
                                                                corr_dev);
            break;
        default: HB_REQUIRE(false, "hb_apply_same_grid: unknown dtype %d", src_dtype);
    }
    HB_LAUNCH_OK("apply_same_grid_kernel");
    return 0;
}

extern "C" int hb_valid_mask(const void *src_dev, int src_dtype, long n, int has_nodata, double nodata,
                             uint8_t *mask_dev, void *stream)
{
    HB_REQUIRE(src_dev && mask_dev && n > 0, "hb_valid_mask: bad arguments");
    const NoData nd = hb_make_nodata(has_nodata, nodata);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = grid_for(n, 256, 16);
    switch (src_dtype) {
        case HB_U8: valid_mask_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)src_dev, n, nd, mask_dev); break;
        case HB_U16:
            valid_mask_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)src_dev, n, nd, mask_dev);
            break;
        case HB_F32: valid_mask_kernel<float><<<grid, 256, 0, st>>>((const float *)src_dev, n, nd, mask_dev); break;
        default: HB_REQUIRE(false, "hb_valid_mask: unknown dtype %d", src_dtype);
    }
    HB_LAUNCH_OK("valid_mask_kernel");
    return 0;
}

extern "C" int hb_full_coverage_mask(const uint8_t *in_mask_dev, long hi, long wi, const float *params_dev, long hp,
                                     long wp, double sx, double ox, double sy, double oy, int kh, int kw,
                                     uint8_t *out_dev, void *workspace_dev, void *stream)
{
    HB_REQUIRE(in_mask_dev && params_dev && out_dev && workspace_dev && hi > 0 && wi > 0 && hp > 0 && wp > 0,
               "hb_full_coverage_mask: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *covered = (uint8_t *)workspace_dev;
    coverage_kernel<<<grid_for(hp * wp, 256), 256, 0, st>>>(in_mask_dev, hi, wi, params_dev, hp, wp, sx, ox, sy, oy,
                                                            covered);
    HB_LAUNCH_OK("coverage_kernel");
    erode_kernel<<<grid_for(hp * wp, 256), 256, 0, st>>>(covered, hp, wp, kh + 2, kw + 2, out_dev);
    HB_LAUNCH_OK("erode_kernel");
    return 0;
}
