// moments.cu -- same-grid sliding-kernel model fit (sm_100a), one fused kernel:
//   masked window moment sums  ->  closed-form gain / gain-blk-offset / gain-offset solve  ->  R2
// replacing the 6 cv.boxFilter / cv.sqrBoxFilter calls and ~30 numpy passes of the reference
// (homonim/kernel_model.py:142-214 _r2_array, :231-274 _fit_gain, :276-303 _fit_gain_blk_offset,
//  :305-359 _fit_gain_offset).  The six window sums never touch HBM.
//
// Data flow per CTA (column strip x row band):
//   * every thread owns C adjacent columns and walks DOWN the band keeping the vertical running sums of its columns
//     in registers (double): add the entering row, subtract the row that left the kh-row window;
//   * per output row the kw-column window sum is a prefix difference: thread-local prefix over its C columns, warp
//     shuffle scan of the per-thread totals, per-column prefixes published once in shared memory, then
//     W(x) = Q(x+hw) - Q(x-hw-1) (+ the total of the warp the window starts in, when it ends in the next one);
//   * the epilogue evaluates the reference's formulas with the reference's precision class and rounding order for
//     every intermediate (SURVEY.md 8a numerics note): sums rounded to float32 where cv.boxFilter returns float32,
//     kept double where cv.sqrBoxFilter returns float64, float32 numerator, double denominator, no FMA contraction.
//   Zero padding beyond the raster (cv BORDER_CONSTANT) falls out of skipping out-of-range rows / columns.
#include <stdlib.h>
#include <type_traits>

#include "hb_common.cuh"

namespace {

#ifndef HB_FIT_MIN_CTAS
#define HB_FIT_MIN_CTAS 4
#endif
constexpr int kFitThreads = 128;
constexpr int kFitWarps = kFitThreads / 32;
constexpr int kMaxHalfW = 63;                    // kw <= 127: a window spans at most two 128-column warps

enum { Q_S = 0, Q_R = 1, Q_P = 2, Q_S2 = 3, Q_R2 = 4 };

// resident CTAs per SM the register allocation aims for (measured on 16384 x 16384 planes, scratch/perf_fit.py): the
// kernel is latency-bound, so the variants that carry fewer running sums trade registers for warps
#ifndef HB_FIT_CTAS_NQ4
#define HB_FIT_CTAS_NQ4 HB_FIT_MIN_CTAS
#endif
#ifndef HB_FIT_CTAS_NQ5
#define HB_FIT_CTAS_NQ5 (HB_FIT_MIN_CTAS - 1)
#endif
constexpr int fit_min_ctas(int model, int nq, int c, bool lean)
{
    if (c != 4) return HB_FIT_MIN_CTAS;
    if (nq >= 5) return lean ? HB_FIT_CTAS_NQ5 : HB_FIT_MIN_CTAS - 1;
    if (nq == 4) return lean ? HB_FIT_CTAS_NQ4 : HB_FIT_MIN_CTAS;
    return lean ? (model == HB_MODEL_GAIN ? HB_FIT_MIN_CTAS + 2 : HB_FIT_MIN_CTAS + 1) : HB_FIT_MIN_CTAS;
}

struct FitGeom {
    long h, w;
    int kh, kw;
    int hw_al;         // half kernel width rounded up to a multiple of C
    int tw_out;        // output columns per CTA
    int rows_per_band;
    long row0, nrows;  // output rows [row0, row0 + nrows) of the plane; output arrays hold just these rows
    int aligned;       // planes are 16-byte aligned and w % 4 == 0: 16-byte loads / stores
    OutSpec ospec;     // FUSED: output dtype / nodata of the corrected plane (conversion fused into the store)
};

// L2 residency hints.  Every input row is read twice by a CTA: when it enters the kh-row window and, kh rows later, when
// it leaves it.  In between the resident CTAs of the whole GPU stream ~(CTAs x kh rows x 512 columns x 8 bytes) -- tens
// of MB -- through L2, so with plain loads the second read mostly misses (measured: 2.2x the compulsory DRAM reads).
// Entering rows are therefore loaded "evict last", leaving rows (their last use) "evict first", and the output is
// written with streaming stores, which leaves L2 to the windows.
__device__ __forceinline__ unsigned long long hb_policy_evict_last()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long hb_policy_evict_first()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 hb_ldg16_hint(const float *p, unsigned long long policy)
{
    float4 r;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ void hb_stg16_stream(float *p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// a / b in double with a quotient that is correctly rounded in all but a vanishing fraction of cases (approximate
// reciprocal + 2 Newton steps + one residual correction: ~9 instructions instead of the ~35 of the IEEE routine).
// Its result is immediately rounded to float32 by the callers, so a last-bit difference in double is invisible
// except on an exact float32 tie.  Straight-line code: operands outside the form's range -- |b| not in [2^-930, 2^930),
// |a| >= 2^930, zeros / infinities / nans included -- set `bad`, and the caller repeats that pixel with IEEE divisions
// (x/0 -> +-inf, 0/0 -> nan must come out exactly as numpy produces them).  The range test reads the exponent fields
// in the integer pipe.
__device__ __forceinline__ double hb_ddiv_fast(double a, double b, bool &bad)
{
    const unsigned eb = ((unsigned)__double2hiint(b) >> 20) & 0x7ffu, ea = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
    bad = bad || (eb - 93u >= 1860u) || (ea >= 1953u);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    const double q = a * r;
    const double rem = fma(-b, q, a);
    return fma(rem, r, q);
}

// 1 / n in double for a window's pixel count n (an integer in [1, 127 * 127]), relative error < 2^-52.  Used for
// offset = x / n in float32 (kernel_model.py:351) as RN_f32(x * (1 / n)): x * (1 / n) is within 2^-51.9 of the exact
// quotient, while x / n, for a float32 x and an integer n < 2^14, is either exactly representable or at least 2^-39
// (relative) away from the nearest float32 rounding tie -- so the rounded result IS the IEEE float32 quotient; infinite
// and nan x propagate as they do in the division.  (n = 0 only occurs outside the mask, where the result is unused.)
__device__ __forceinline__ double hb_drcp_count(double n)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(n));
    double e = fma(-n, r, 1.0);
    r = fma(r, e, r);
    e = fma(-n, r, 1.0);
    return fma(r, e, r);
}

// contribution of one pixel to the running sums
template <int NQ, bool NORM>
__device__ __forceinline__ void pixel_terms(float s, float r, bool valid, double n0, double n1, double (&q)[NQ],
                                            int &cnt)
{
    double ds, dr;
    if (NORM) {
        // src * norm[0] + norm[1] in float64, two roundings (kernel_model.py:295 under numpy >= 2)
        ds = __dadd_rn(__dmul_rn((double)s, n0), n1);
        valid = valid && !isnan(ds);
        if (!valid) { ds = 0.0; r = 0.f; }       // kernel_model.py:246-247
        dr = (double)r;
    } else {
        // zero the invalid pixel once, on the float32 inputs (kernel_model.py:246-247 / 320-321): every term below is
        // then zero without further selects
        s = valid ? s : 0.f;
        r = valid ? r : 0.f;
        ds = (double)s;
        dr = (double)r;
    }
    q[Q_S] = ds;
    q[Q_R] = dr;
    if (NQ > 2) {
        // src*ref is formed in the arrays' dtype before filtering (kernel_model.py:175, 334): float32 product for
        // float32 planes, float64 when the source was normalised
        q[Q_P] = NORM ? __dmul_rn(ds, dr) : (double)__fmul_rn(s, r);
        q[Q_S2] = __dmul_rn(ds, ds);             // cv.sqrBoxFilter squares in double
    }
    if (NQ > 4) q[Q_R2] = __dmul_rn(dr, dr);
    cnt = valid ? 1 : 0;
}

// MODEL: HB_MODEL_*; WANT_R2: third band; NQ: number of double sums carried (2, 4 or 5); C: columns per thread;
// FUSED: write corr = gain * src + offset instead of the parameters.
// LEAN: both planes use NaN as nodata and the kernel is at least 3 x 3 (the float32 rasters of every large configuration).
//   The kernel is bound by instruction issue, not by memory, and the general form spends ~40 % of its instructions on
//   bookkeeping: the 4-compare nodata test with its row / column gates, validity bits pulled back out of the ring for the
//   leaving row, run-time exchange-buffer addresses, and the kh == 1 / kw == 1 special cases.  Here a pixel is valid iff
//   neither value is NaN (ONE setp.num), rows / columns outside the raster are loaded as NaN so that they need no gate,
//   the leaving row is simply tested again, and the row loop is unrolled by two so that the exchange buffer of a step is
//   an immediate in every shared-memory address.  Same additions in the same order: results are bit-identical to the
//   general form (tests/test_gpu_parity.py::test_same_grid_lean_equals_general).
//
// Shared-memory exchange of one output row (double-buffered, one barrier per row).  Per sum k a row of SLOTS doubles:
//   slots [0, TW)      warp-local inclusive prefixes of the column sums; column c lives at (c % C) * threads + c / C,
//                      conflict-free for the stores (thread t, its column i) and for the window look-ups
//   slot  TW           a permanent 0.0
//   slots TW + 1 + w   total of warp w
// The window sum of column c is  Q(c + hw) - Q(c - hw - 1) + (total of the warp the window starts in, if it ends in the
// next one), and every term's slot is a per-thread constant: where a term does not apply (no warp boundary inside the
// window, window starting at the CTA's first column) its slot is the permanent zero, so the row loop has no selects,
// no index arithmetic and no divergent branches -- the byte offsets are computed once per thread, the sum's offset is
// an immediate and the buffer's offset a warp-uniform register.  (The launcher guarantees kw <= columns per warp, so a
// window never spans more than two warps.)
template <int MODEL, bool WANT_R2, int NQ, int C, bool FUSED, bool LEAN>
__global__ void __launch_bounds__(kFitThreads, fit_min_ctas(MODEL, NQ, C, LEAN))
fit_same_grid_kernel(const float *__restrict__ src, NoData nd_s, const float *__restrict__ ref, NoData nd_r,
                     FitGeom g, const double *__restrict__ norm, float *__restrict__ params,
                     float *__restrict__ sums_out, float *__restrict__ corr_out)   // (corr_out: g.ospec.dtype elements)
{
    constexpr bool NORM = (MODEL == HB_MODEL_GAIN_BLK_OFFSET);
    constexpr bool HAS_N = (NQ > 2);
    constexpr int TW = kFitThreads * C;                       // columns per CTA including both halos
    constexpr int WCOLS = 32 * C;                             // columns per warp
    constexpr int SLOTS = TW + 8;
    constexpr int ZERO = TW;
    // (the valid count rides along as one more row of 4-byte slots, addressed with half the byte offset)
    constexpr int ROW_BYTES = SLOTS * 8, BUF_BYTES = NQ * ROW_BYTES + (HAS_N ? ROW_BYTES / 2 : 0);
    __shared__ double s_x[2 * BUF_BYTES / 8];
    char *const s_base = reinterpret_cast<char *>(s_x);
    auto q_at = [&](int bufoff, int k, unsigned off) -> double & {
        return *reinterpret_cast<double *>(s_base + bufoff + k * ROW_BYTES + off);
    };
    auto n_at = [&](int bufoff, unsigned off) -> int & {
        return *reinterpret_cast<int *>(s_base + bufoff + NQ * ROW_BYTES + (off >> 1));
    };

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int hh = g.kh / 2, hw = g.kw / 2;
    const int w = (int)g.w;
    const int cx = (int)blockIdx.x * g.tw_out - g.hw_al + t * C;     // first global column of this thread (may be < 0)
    // (row indices are 32-bit -- the launcher checks h < 2^31 -- so that the per-row tests and the row * width products
    //  are single instructions; only the final element offsets are 64-bit)
    const int h = (int)g.h;
    const int y0 = (int)g.row0 + (int)blockIdx.y * g.rows_per_band;  // output rows [y0, y1) of the plane's rows [0, h)
    const int y1 = min(y0 + g.rows_per_band, (int)(g.row0 + g.nrows));
    const long plane = g.nrows * g.w;                                // output planes hold rows [row0, row0 + nrows)
    const float qnan = __int_as_float(0x7fc00000);
    double n0 = 1.0, n1 = 0.0;
    if (NORM) { n0 = norm[0]; n1 = norm[1]; }

    if (t < 2 * NQ) q_at((t / NQ) * BUF_BYTES, t % NQ, ZERO * 8) = 0.0;
    if (HAS_N && t < 2) n_at(t * BUF_BYTES, ZERO * 8) = 0;

    // validity of a pixel without per-pixel branching on the nodata kind: "equals the nodata value" compares with nan
    // (never true) when there is no value nodata, and the nan test is OR-ed with a kernel-uniform "nan is data" flag:
    // four compare-and-combine instructions per pixel, result AND-ed with `gate` (column / row inside the raster)
    const float ndv_s = (nd_s.has && !nd_s.is_nan) ? nd_s.value : qnan, ndv_r = (nd_r.has && !nd_r.is_nan) ? nd_r.value : qnan;
    const unsigned keep_nan_s = !(nd_s.has && nd_s.is_nan), keep_nan_r = !(nd_r.has && nd_r.is_nan);
    auto valid2 = [&](float sv, float rv, unsigned gate) -> bool {
        unsigned ok;
        if constexpr (LEAN) {
            // (gate: the row exists -- warp-uniform, its setp is hoisted out of the per-pixel code)
            asm("{\n\t.reg .pred p, g;\n\tsetp.ne.u32 g, %3, 0;\n\tsetp.num.and.f32 p, %1, %2, g;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "f"(sv), "f"(rv), "r"(gate));
            return ok != 0u;
        }
        asm("{\n\t.reg .pred p, q, ks, kr;\n\t"
            "setp.ne.u32 ks, %3, 0;\n\tsetp.ne.u32 kr, %4, 0;\n\t"
            "setp.num.or.f32 p, %1, %1, ks;\n\tsetp.neu.and.f32 p, %1, %5, p;\n\t"
            "setp.num.or.f32 q, %2, %2, kr;\n\tsetp.neu.and.f32 q, %2, %6, q;\n\t"
            "and.pred p, p, q;\n\tselp.u32 %0, %7, 0, p;\n\t}"
            : "=r"(ok)
            : "f"(sv), "f"(rv), "r"(keep_nan_s), "r"(keep_nan_r), "f"(ndv_s), "f"(ndv_r), "r"(gate));
        return ok != 0u;
    };
    // bit i: this thread's column i lies inside the raster
    unsigned cmask = 0;
#pragma unroll
    for (int i = 0; i < C; i++) cmask |= ((cx + i >= 0) && (cx + i < w)) ? (1u << i) : 0u;
    asm volatile("" : "+r"(cmask));      // (kept in its register: re-deriving it from cx and w every row costs more)
    const bool vec_ok = g.aligned && (C == 4) && (cmask == 0xfu);
    // this thread's columns are output columns iff they sit between the two halos (whole-thread granularity)
    const bool out_thread = (t * C >= g.hw_al) && (t * C + C <= g.hw_al + g.tw_out);

    // per-thread constant slots of the window look-ups (see above), as byte offsets inside a row of slots
    unsigned sl_a[C], sl_b[C], sl_t0[C];
#pragma unroll
    for (int i = 0; i < C; i++) {
        const int ci = t * C + i;
        const int a = min(ci + hw, TW - 1), b = ci - hw - 1;         // (a is only clamped for halo threads: unused)
        const int wa = a / WCOLS, wb = (b >= 0) ? b / WCOLS : 0;
        sl_a[i] = 8u * ((a % C) * kFitThreads + a / C);
        sl_b[i] = 8u * ((b >= 0) ? (b % C) * kFitThreads + b / C : ZERO);
        sl_t0[i] = 8u * ((wa != wb) ? TW + 1 + wb : ZERO);
    }
    const unsigned sl_own = 8u * t, sl_tot = 8u * (TW + 1 + warp);   // this thread's store slots (column i: + i * threads)

    double V[C][NQ];
    int VN[C];
    unsigned long long vring[C];                                     // validity of the last 64 rows, per column
#pragma unroll
    for (int i = 0; i < C; i++) {
#pragma unroll
        for (int q = 0; q < NQ; q++) V[i][q] = 0.0;
        VN[i] = 0;
        vring[i] = 0ull;
    }

    const unsigned long long pol_keep = hb_policy_evict_last(), pol_drop = hb_policy_evict_first();
    // `last_use`: the row is leaving the window (its final read) -- see the L2 note above
    auto load_row = [&](int y, float (&s)[C], float (&r)[C], bool last_use) {
        const long off = (long)y * w + cx;
        const float *srow = src + off, *rrow = ref + off;
        if (vec_ok) {
            const float4 a = hb_ldg16_hint(srow, last_use ? pol_drop : pol_keep);
            const float4 b = hb_ldg16_hint(rrow, last_use ? pol_drop : pol_keep);
            s[0] = a.x; r[0] = b.x;
            if (C > 1) { s[1] = a.y; r[1] = b.y; }
            if (C > 2) { s[2] = a.z; r[2] = b.z; }
            if (C > 3) { s[3] = a.w; r[3] = b.w; }
        } else {
#pragma unroll
            for (int i = 0; i < C; i++) {
                const bool in = (cmask >> i) & 1u;
                s[i] = in ? __ldg(srow + i) : (LEAN ? qnan : 0.f);
                r[i] = in ? __ldg(rrow + i) : (LEAN ? qnan : 0.f);
            }
        }
    };

    const bool single_row = !LEAN && (g.kh == 1), single_col = !LEAN && (g.kw == 1);
    const int e_first = y0 - hh, e_last = y1 - 1 + hh;
    const int l_first = max(e_first, 0);                             // first row that ever leaves a window of this band
    float se[C], re[C], sl[C], rl[C];
#pragma unroll
    for (int i = 0; i < C; i++) { se[i] = re[i] = sl[i] = rl[i] = 0.f; }
    auto fetch = [&](int e) {
        const int l = e - g.kh;
        if constexpr (LEAN) {
            // both rows are ALWAYS loaded (row index clamped into the raster) and always consumed: a load that is only
            // consumed on some paths makes the compiler wait for it -- and for every younger load sharing its scoreboard
            // -- before it may reuse the destination registers (measured: 17 % of all stall samples).  Whether the row
            // counts is decided by the validity test's gate.
            load_row(min(max(e, 0), h - 1), se, re, false);
            load_row(min(max(l, 0), h - 1), sl, rl, true);
        } else {
            if ((unsigned)e < (unsigned)h) load_row(e, se, re, false);
            if (!single_row && (l >= l_first) && (l < h)) load_row(l, sl, rl, true);
        }
    };
    // Vertical update with one row's pixels; `rowmask` = cmask if the row exists, else 0 (rows outside the raster / before
    // the band's first window contribute as invalid pixels: all-zero terms, so the update is free of branches; their
    // registers hold stale values that the mask hides).  A LEAVING pixel's validity is the bit it shifted into the ring
    // when it entered kh rows ago -- no second test -- as long as the ring (64 rows) reaches that far.
    const bool ring_has_leaver = g.kh <= 63;
    auto accumulate = [&](const float (&s)[C], const float (&r)[C], unsigned rowmask, bool entering) {
#pragma unroll
        for (int i = 0; i < C; i++) {
            double q[NQ]; int cnt;
            bool v;
            if (LEAN || entering || !ring_has_leaver) v = valid2(s[i], r[i], (rowmask >> i) & 1u);
            else v = (((unsigned)(vring[i] >> g.kh)) & (rowmask >> i) & 1u) != 0u;
            pixel_terms<NQ, NORM>(s[i], r[i], v, n0, n1, q, cnt);
            if (entering) {
#pragma unroll
                for (int k = 0; k < NQ; k++) V[i][k] = __dadd_rn(V[i][k], q[k]);
                VN[i] += cnt;
                vring[i] = vring[i] + vring[i] + (unsigned long long)cnt;      // (ring << 1) | valid
            } else {
#pragma unroll
                for (int k = 0; k < NQ; k++) V[i][k] = __dsub_rn(V[i][k], q[k]);
                VN[i] -= cnt;
            }
        }
    };
    // a row that does not exist contributes through an all-zero row mask (LEAN: columns outside the raster were loaded
    // as NaN, so its mask is only about the row)
    auto update = [&](const float (&s)[C], const float (&r)[C], bool row_exists, bool entering) {
        accumulate(s, r, row_exists ? (LEAN ? 0xfu : cmask) : 0u, entering);
    };
    // ---- warm-up: the rows above the first output row's centre only ENTER the window (nothing leaves, no output row is
    //      completed).  Fetch them 4 at a time -- all loads in flight together -- and accumulate in row order (the same
    //      additions, in the same order, as the row-by-row steps below would perform).
    int e_begin = e_first;
    if (!single_row) {
        const int e_warm_end = y0 + hh;                              // first step that completes an output row
        for (; e_begin < e_warm_end; e_begin += 4) {
            float ws[4][C], wr[4][C];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e_begin + u;
#pragma unroll
                for (int i = 0; i < C; i++) { ws[u][i] = 0.f; wr[u][i] = 0.f; }
                if (LEAN) load_row(min(max(e, 0), h - 1), ws[u], wr[u], false);
                else if (e < e_warm_end && (unsigned)e < (unsigned)h) load_row(e, ws[u], wr[u], false);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e_begin + u;
                if (e >= e_warm_end) break;
                update(ws[u], wr[u], (unsigned)e < (unsigned)h, true);
            }
        }
        e_begin = e_warm_end;
    }
    fetch(e_begin);

    // ---- row loop: vertical update with rows e (entering) and e - kh (leaving), which were fetched one step ago; the
    //      rows of step e + 1 are requested as soon as those registers are free, so their latency hides behind the scan
    //      and the solve
    auto row_step = [&](const int e, const int bufoff) {            // bufoff: exchange buffer of this step (warp-uniform)
        const int l = e - g.kh;
        const bool has_e = (unsigned)e < (unsigned)h;
        const bool has_l = !single_row && (l >= l_first) && (l < h);
        if (single_row) {                                            // kh == 1: the window IS this row (exact)
#pragma unroll
            for (int i = 0; i < C; i++) {
#pragma unroll
                for (int k = 0; k < NQ; k++) V[i][k] = 0.0;
                VN[i] = 0;
            }
        }
        update(se, re, has_e, true);
        update(sl, rl, has_l, false);
        if (e < e_last) fetch(e + 1);
        const int y = e - hh;                                        // output row completed by this step

        // ---- horizontal window sums: prefix over this thread's columns, warp scan of thread totals ----------------
        // (kw == 1: the window sum is the column sum itself -- exact, and no exchange is needed)
        if (!single_col) {
            double pre[C][NQ], incl[NQ];
            int pren[C], incl_n = 0;
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < C; i++) { acc += V[i][k]; pre[i][k] = acc; }
                incl[k] = acc;
            }
            if (HAS_N) {
                int acc = 0;
#pragma unroll
                for (int i = 0; i < C; i++) { acc += VN[i]; pren[i] = acc; }
                incl_n = acc;
            }
            // the scans of all sums advance together: NQ + 1 independent shuffle / add chains per step
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                double up[NQ];
                int up_n = 0;
#pragma unroll
                for (int k = 0; k < NQ; k++) up[k] = __shfl_up_sync(0xffffffffu, incl[k], d);
                if (HAS_N) up_n = __shfl_up_sync(0xffffffffu, incl_n, d);
                if (lane >= d) {
#pragma unroll
                    for (int k = 0; k < NQ; k++) incl[k] += up[k];
                    if (HAS_N) incl_n += up_n;
                }
            }
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                const double excl = incl[k] - pre[C - 1][k];
                if (lane == 31) q_at(bufoff, k, sl_tot) = incl[k];
#pragma unroll
                for (int i = 0; i < C; i++) q_at(bufoff, k, sl_own + i * (kFitThreads * 8)) = excl + pre[i][k];
            }
            if (HAS_N) {
                const int excl_n = incl_n - pren[C - 1];
                if (lane == 31) n_at(bufoff, sl_tot) = incl_n;
#pragma unroll
                for (int i = 0; i < C; i++) n_at(bufoff, sl_own + i * (kFitThreads * 8)) = excl_n + pren[i];
            }
            __syncthreads();        // (double-buffered: the next row writes the other buffer, so one barrier per row)
        }

        if (!out_thread) return;
        float o_gain[C], o_off[C], o_r2[C], o_rs[C], o_ss[C], o_n[C];
        // One pixel: window sums from the exchange buffer, closed-form solve.  IEEE = false: every division takes the
        // straight-line fast form and the pixel is reported "bad" when an operand is outside that form's range; bad pixels
        // (zero / non-finite denominators: flat windows, borders of nodata areas -- rare) are then solved again with IEEE
        // divisions.  With no branch inside the common path the compiler interleaves the C independent solves instead of
        // running their ~25-deep dependent chains one after the other.
        auto solve_pixel = [&](const int i, auto ieee_tag) -> bool {
            constexpr bool IEEE = decltype(ieee_tag)::value;
            bool bad = false;
            auto ddiv = [&](double a, double b) -> double {
                if constexpr (IEEE) return a / b;
                else return hb_ddiv_fast(a, b, bad);
            };
            double W[NQ];
            int N = 0;
            if (single_col) {
#pragma unroll
                for (int k = 0; k < NQ; k++) W[k] = V[i][k];
                N = VN[i];
            } else {
#pragma unroll
                for (int k = 0; k < NQ; k++)
                    W[k] = (q_at(bufoff, k, sl_a[i]) + q_at(bufoff, k, sl_t0[i])) - q_at(bufoff, k, sl_b[i]);
                if (HAS_N) N = (n_at(bufoff, sl_a[i]) + n_at(bufoff, sl_t0[i])) - n_at(bufoff, sl_b[i]);
            }
            const bool mask = ((vring[i] >> hh) & 1ull) != 0ull;     // centre pixel valid in both images

            // ---- closed-form solve with the reference's precision classes ---------------------------------------------
            const float fR = (float)W[Q_R];                          // cv.boxFilter(f32) -> f32
            float gain = qnan, off = qnan, r2 = qnan;
            float fS = 0.f, fN = 0.f;
            if (MODEL == HB_MODEL_GAIN_OFFSET) {
                fS = (float)W[Q_S];
                const float fP = (float)W[Q_P];
                fN = (float)N;
                const double dN = (double)fN;
                const float num = __fsub_rn(__fmul_rn(fN, fP), __fmul_rn(fS, fR));                  // :338, float32
                const double den = __dsub_rn(__dmul_rn(dN, W[Q_S2]), (double)__fmul_rn(fS, fS));    // :342
                gain = (float)ddiv((double)num, den);                                                // :348
                const float onum = __fsub_rn(fR, __fmul_rn(gain, fS));
                if constexpr (IEEE) off = __fdiv_rn(onum, fN);                                       // :351
                else off = (float)__dmul_rn((double)onum, hb_drcp_count(dN));   // == onum / fN correctly rounded, see there
                if (WANT_R2) {
                    const double ss_tot = __dsub_rn(__dmul_rn(dN, W[Q_R2]), (double)__fmul_rn(fR, fR));   // :179
                    const double t1 = __dmul_rn((double)__fmul_rn(gain, gain), W[Q_S2]);
                    const float t2 = __fmul_rn(__fmul_rn(2.f, __fmul_rn(gain, off)), fS);
                    const float t3 = __fmul_rn(__fmul_rn(2.f, gain), fP);
                    const float t4 = __fmul_rn(__fmul_rn(2.f, off), fR);
                    const float t6 = __fmul_rn(fN, __fmul_rn(off, off));
                    double ss_res = __dadd_rn(t1, (double)t2);                                       // :189-195
                    ss_res = __dsub_rn(ss_res, (double)t3);
                    ss_res = __dsub_rn(ss_res, (double)t4);
                    ss_res = __dadd_rn(ss_res, W[Q_R2]);
                    ss_res = __dadd_rn(ss_res, (double)t6);
                    ss_res = __dmul_rn(ss_res, dN);                                                  // :203
                    r2 = __fsub_rn(1.f, (float)ddiv(ss_res, ss_tot));                               // :212-213
                }
            } else {
                // gain (kernel_model.py:265) -- for gain-blk-offset on the normalised, float64 source sums
                float g0;
                if (NORM) g0 = (float)ddiv((double)fR, W[Q_S]);
                else { fS = (float)W[Q_S]; g0 = __fdiv_rn(fR, fS); }   // (12 float32 instructions: cheaper than any double route)
                if (WANT_R2) {
                    fN = (float)N;
                    const double ss_tot = __dsub_rn(__dmul_rn((double)fN, W[Q_R2]), (double)__fmul_rn(fR, fR));
                    const double t1 = __dmul_rn((double)__fmul_rn(g0, g0), W[Q_S2]);
                    const float g2 = __fmul_rn(2.f, g0);
                    const double t3 = NORM ? __dmul_rn((double)g2, W[Q_P]) : (double)__fmul_rn(g2, (float)W[Q_P]);
                    double ss_res = __dadd_rn(__dsub_rn(t1, t3), W[Q_R2]);                           // :201
                    ss_res = __dmul_rn(ss_res, (double)fN);
                    r2 = __fsub_rn(1.f, (float)ddiv(ss_res, ss_tot));
                }
                if (NORM) {
                    off = (float)__dmul_rn((double)g0, n1);                                          // :301
                    gain = (float)__dmul_rn((double)g0, n0);                                         // :302
                } else {
                    gain = g0;
                    off = 0.f;                                                                       // :262
                }
            }
            o_gain[i] = mask ? gain : qnan;
            o_off[i] = mask ? off : qnan;
            o_r2[i] = mask ? r2 : qnan;
            o_rs[i] = fR; o_ss[i] = fS; o_n[i] = mask ? fN : -1.f;   // count plane: -1 marks "outside the mask"
            return bad && mask;                                      // (outside the mask the values are not used)
        };
        unsigned redo = 0u;
#pragma unroll
        for (int i = 0; i < C; i++) redo |= solve_pixel(i, std::false_type{}) ? (1u << i) : 0u;
        if (redo != 0u) {
#pragma unroll
            for (int i = 0; i < C; i++)
                if ((redo >> i) & 1u) solve_pixel(i, std::true_type{});
        }
        // ---- store -------------------------------------------------------------------------------------------------
        const long yo = (long)y - g.row0;                            // row of the output planes
        if (FUSED) {
            // fused apply (KernelModel.apply, kernel_model.py:461): corr = gain * src + offset with the centre row's
            // ORIGINAL source pixels (re-read: they entered the window kh/2 rows ago, an L2 hit), two float32
            // roundings as numpy; the parameters are not materialised
            float sc[C];
            const float *srow = src + ((long)y * w + cx);
            if (vec_ok) {
                const float4 a = hb_ldg16_hint(srow, pol_keep);
                sc[0] = a.x;
                if (C > 1) sc[1] = a.y;
                if (C > 2) sc[2] = a.z;
                if (C > 3) sc[3] = a.w;
            } else {
#pragma unroll
                for (int i = 0; i < C; i++) sc[i] = ((cmask >> i) & 1u) ? __ldg(srow + i) : 0.f;
            }
            float oc[C];
#pragma unroll
            for (int i = 0; i < C; i++) oc[i] = __fadd_rn(__fmul_rn(o_gain[i], sc[i]), o_off[i]);
            const long cpix = yo * g.w + cx;                         // pixel index in the corrected plane
            if (vec_ok) {
                hb_store4_out(corr_out, cpix, make_float4(oc[0], oc[C > 1 ? 1 : 0], oc[C > 2 ? 2 : 0], oc[C > 3 ? 3 : 0]),
                              g.ospec);
            } else {
#pragma unroll
                for (int i = 0; i < C; i++)
                    if ((cmask >> i) & 1u) hb_store1_out(corr_out, cpix + i, oc[i], g.ospec);
            }
        } else {
            float *prow = params + yo * g.w + cx;
            if (vec_ok) {
                hb_stg16_stream(prow, make_float4(o_gain[0], o_gain[C > 1 ? 1 : 0], o_gain[C > 2 ? 2 : 0], o_gain[C > 3 ? 3 : 0]));
                hb_stg16_stream(prow + plane, make_float4(o_off[0], o_off[C > 1 ? 1 : 0], o_off[C > 2 ? 2 : 0], o_off[C > 3 ? 3 : 0]));
                if (WANT_R2)
                    hb_stg16_stream(prow + 2 * plane, make_float4(o_r2[0], o_r2[C > 1 ? 1 : 0], o_r2[C > 2 ? 2 : 0], o_r2[C > 3 ? 3 : 0]));
                if (sums_out != nullptr) {
                    float *srow = sums_out + yo * g.w + cx;
                    hb_stg16_stream(srow, make_float4(o_rs[0], o_rs[C > 1 ? 1 : 0], o_rs[C > 2 ? 2 : 0], o_rs[C > 3 ? 3 : 0]));
                    hb_stg16_stream(srow + plane, make_float4(o_ss[0], o_ss[C > 1 ? 1 : 0], o_ss[C > 2 ? 2 : 0], o_ss[C > 3 ? 3 : 0]));
                    hb_stg16_stream(srow + 2 * plane, make_float4(o_n[0], o_n[C > 1 ? 1 : 0], o_n[C > 2 ? 2 : 0], o_n[C > 3 ? 3 : 0]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < C; i++) {
                    if (!((cmask >> i) & 1u)) continue;
                    prow[i] = o_gain[i];
                    prow[plane + i] = o_off[i];
                    if (WANT_R2) prow[2 * plane + i] = o_r2[i];
                    if (sums_out != nullptr) {
                        float *srow = sums_out + yo * g.w + cx;
                        srow[i] = o_rs[i];
                        srow[plane + i] = o_ss[i];
                        srow[2 * plane + i] = o_n[i];
                    }
                }
            }
        }
    };
    if constexpr (LEAN) {
        // two steps per iteration: the buffer offsets are immediates in both copies of the body
        for (int e = e_begin; e <= e_last; e += 2) {
            row_step(e, 0);
            if (e + 1 <= e_last) row_step(e + 1, BUF_BYTES);
        }
    } else {
        for (int e = e_begin; e <= e_last; e++) row_step(e, ((e - e_begin) & 1) * BUF_BYTES);
    }
}

// HOMONIM_B200_FIT_GENERAL=1 (read at every call: a test switch) forces the general form of the kernel
static bool hb_fit_force_general()
{
    const char *e = getenv("HOMONIM_B200_FIT_GENERAL");
    return e != nullptr && e[0] == '1';
}

template <int MODEL, bool WANT_R2, int NQ, int C>
int launch_fit(const float *src, NoData nd_s, const float *ref, NoData nd_r, long h, long w, long row0, long nrows,
               int kh, int kw, const double *norm, float *params, float *sums, float *corr, OutSpec ospec,
               cudaStream_t stream)
{
    FitGeom g;
    g.h = h; g.w = w; g.kh = kh; g.kw = kw; g.row0 = row0; g.nrows = nrows; g.ospec = ospec;
    const int hw = kw / 2;
    g.hw_al = ((hw + C - 1) / C) * C;
    g.tw_out = kFitThreads * C - 2 * g.hw_al;
    HB_REQUIRE(g.tw_out > 0 && kw <= 32 * C, "hb_fit_same_grid: kernel width %d is too wide for this variant", kw);
    HB_REQUIRE(w < (1L << 30) && h < (1L << 30), "hb_fit_same_grid: rasters beyond 2^30 pixels per side are not supported");
    const long xtiles = (w + g.tw_out - 1) / g.tw_out;
    // rows per band: every band re-reads (kh - 1) warm-up rows, so bands should be tall -- ~8 window heights -- unless
    // that leaves SMs idle (mid-size rasters: go down to 2 window heights to get one CTA per resident slot).  Small
    // rasters (C == 1) are latency-bound with the machine mostly empty: short bands, many CTAs.
    const long slots = (long)hb_sm_count() * HB_FIT_MIN_CTAS;
    long rpb;
    if (C == 1) {
        const long bands_t = ((long)hb_sm_count() * 8 + xtiles - 1) / xtiles;
        rpb = (nrows + bands_t - 1) / bands_t;
        if (rpb < 4) rpb = 4;
    } else {
        // score = (fraction of the last wave of resident CTAs that is filled) x (fraction of row fetches that are not
        // warm-up); candidates from 2 to 16 window heights
        double best = -1.0;
        rpb = 8L * kh;
        for (long cand = 16L * kh; cand >= 2L * kh; cand -= (kh + 1) / 2) {
            const long c = cand > nrows ? nrows : cand;
            const long ctas = xtiles * ((nrows + c - 1) / c);
            const long waves = (ctas + slots - 1) / slots;
            const double score = ((double)ctas / (double)(waves * slots)) * ((double)c / (double)(c + kh - 1));
            if (score > best) { best = score; rpb = c; }
        }
    }
    if (rpb > nrows) rpb = nrows;
    long bands;
    g.rows_per_band = (int)rpb;
    bands = (nrows + rpb - 1) / rpb;
    HB_REQUIRE(bands <= 65535, "hb_fit_same_grid: too many row bands");
    dim3 grid((unsigned)xtiles, (unsigned)bands);
    g.aligned = ((C == 4) && (w % 4 == 0) && (((uintptr_t)src) % 16 == 0) && (((uintptr_t)ref) % 16 == 0) &&
                 (params == nullptr || ((uintptr_t)params) % 16 == 0) &&
                 (sums == nullptr || ((uintptr_t)sums) % 16 == 0) &&
                 (corr == nullptr || ((uintptr_t)corr) % (4 * hb_out_size(ospec.dtype)) == 0)) ? 1 : 0;
    // the lean form (see the kernel): NaN nodata on both planes, kernel at least 3 x 3; large-raster variant only
    const bool lean = (C == 4) && nd_s.has && nd_s.is_nan && nd_r.has && nd_r.is_nan && kh > 1 && kw > 1 && !hb_fit_force_general();
    if (corr != nullptr) {
        if constexpr (!WANT_R2) {
            if (lean && C == 4)
                fit_same_grid_kernel<MODEL, false, NQ, C, true, (C == 4)><<<grid, kFitThreads, 0, stream>>>(
                    src, nd_s, ref, nd_r, g, norm, nullptr, nullptr, corr);
            else
                fit_same_grid_kernel<MODEL, false, NQ, C, true, false><<<grid, kFitThreads, 0, stream>>>(
                    src, nd_s, ref, nd_r, g, norm, nullptr, nullptr, corr);
        } else
            HB_REQUIRE(false, "hb_fit_apply_same_grid: the fused apply does not produce an R2 band");
    } else {
        if (lean && C == 4)
            fit_same_grid_kernel<MODEL, WANT_R2, NQ, C, false, (C == 4)><<<grid, kFitThreads, 0, stream>>>(
                src, nd_s, ref, nd_r, g, norm, params, sums, nullptr);
        else
            fit_same_grid_kernel<MODEL, WANT_R2, NQ, C, false, false><<<grid, kFitThreads, 0, stream>>>(
                src, nd_s, ref, nd_r, g, norm, params, sums, nullptr);
    }
    HB_LAUNCH_OK("fit_same_grid_kernel");
    return 0;
}

template <int MODEL, bool WANT_R2, int NQ>
int launch_fit_c(const float *src, NoData nd_s, const float *ref, NoData nd_r, long h, long w, long row0, long nrows,
                 int kh, int kw, const double *norm, float *params, float *sums, float *corr, OutSpec ospec,
                 cudaStream_t stream)
{
    // small rasters: 1 column per thread (128-column strips) so that the grid still spreads over the SMs.  A window
    // may span at most two warps (one warp total in the look-up): 32-column warps carry kernels up to 31 wide,
    // 128-column warps up to 127.
    if (nrows * w < ((long)4 << 20) && kw <= 31)
        return launch_fit<MODEL, WANT_R2, NQ, 1>(src, nd_s, ref, nd_r, h, w, row0, nrows, kh, kw, norm, params, sums,
                                                 corr, ospec, stream);
    return launch_fit<MODEL, WANT_R2, NQ, 4>(src, nd_s, ref, nd_r, h, w, row0, nrows, kh, kw, norm, params, sums, corr,
                                             ospec, stream);
}

}  // namespace

static int fit_dispatch(const float *src_dev, int src_has_nodata, double src_nodata, const float *ref_dev,
                        int ref_has_nodata, double ref_nodata, long h, long w, long row0, long nrows, int model, int kh,
                        int kw, int want_r2, const double *norm_dev, float *params_dev, float *sums_dev, float *corr_dev,
                        void *stream, const char *who, OutSpec ospec = hb_make_outspec(HB_F32, 0, 0.0))
{
    HB_REQUIRE(src_dev && ref_dev && (params_dev || corr_dev) && h > 0 && w > 0, "%s: bad arguments", who);
    HB_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= h, "%s: output rows [%ld, %ld) outside the %ld-row plane", who,
               row0, row0 + nrows, h);
    HB_REQUIRE(kh >= 1 && kw >= 1 && (kh & 1) && (kw & 1), "%s: kernel shape must be odd and >= 1", who);
    HB_REQUIRE(kw / 2 <= kMaxHalfW, "%s: kernel width %d > %d is not supported", who, kw, 2 * kMaxHalfW + 1);
    HB_REQUIRE(kh / 2 <= 63, "%s: kernel height %d > 127 is not supported", who, kh);
    const NoData nd_s = hb_make_nodata(src_has_nodata, src_nodata), nd_r = hb_make_nodata(ref_has_nodata, ref_nodata);
    cudaStream_t st = (cudaStream_t)stream;
#define HB_FIT_ARGS src_dev, nd_s, ref_dev, nd_r, h, w, row0, nrows, kh, kw
    switch (model) {
        case HB_MODEL_GAIN:
            HB_REQUIRE(sums_dev == nullptr, "%s: sums are only produced for the gain-offset model", who);
            if (want_r2)
                return launch_fit_c<HB_MODEL_GAIN, true, 5>(HB_FIT_ARGS, nullptr, params_dev, nullptr, corr_dev, ospec, st);
            return launch_fit_c<HB_MODEL_GAIN, false, 2>(HB_FIT_ARGS, nullptr, params_dev, nullptr, corr_dev, ospec, st);
        case HB_MODEL_GAIN_BLK_OFFSET:
            HB_REQUIRE(norm_dev != nullptr, "%s: gain-blk-offset needs the block normalisation", who);
            HB_REQUIRE(sums_dev == nullptr, "%s: sums are only produced for the gain-offset model", who);
            if (want_r2)
                return launch_fit_c<HB_MODEL_GAIN_BLK_OFFSET, true, 5>(HB_FIT_ARGS, norm_dev, params_dev, nullptr,
                                                                       corr_dev, ospec, st);
            return launch_fit_c<HB_MODEL_GAIN_BLK_OFFSET, false, 2>(HB_FIT_ARGS, norm_dev, params_dev, nullptr,
                                                                    corr_dev, ospec, st);
        case HB_MODEL_GAIN_OFFSET:
            if (want_r2)
                return launch_fit_c<HB_MODEL_GAIN_OFFSET, true, 5>(HB_FIT_ARGS, nullptr, params_dev, sums_dev, corr_dev,
                                                                   ospec, st);
            return launch_fit_c<HB_MODEL_GAIN_OFFSET, false, 4>(HB_FIT_ARGS, nullptr, params_dev, sums_dev, corr_dev,
                                                                ospec, st);
    }
#undef HB_FIT_ARGS
    HB_REQUIRE(false, "%s: unknown model %d", who, model);
}

extern "C" int hb_fit_same_grid(const float *src_dev, int src_has_nodata, double src_nodata, const float *ref_dev,
                                int ref_has_nodata, double ref_nodata, long h, long w, int model, int kh, int kw,
                                int want_r2, const double *norm_dev, float *params_dev, float *sums_dev, void *stream)
{
    HB_REQUIRE(params_dev != nullptr, "hb_fit_same_grid: bad arguments");
    return fit_dispatch(src_dev, src_has_nodata, src_nodata, ref_dev, ref_has_nodata, ref_nodata, h, w, 0, h, model, kh,
                        kw, want_r2, norm_dev, params_dev, sums_dev, nullptr, stream, "hb_fit_same_grid");
}

extern "C" int hb_fit_same_grid_rows(const float *src_dev, int src_has_nodata, double src_nodata, const float *ref_dev,
                                     int ref_has_nodata, double ref_nodata, long h, long w, long row0, long nrows,
                                     int model, int kh, int kw, int want_r2, const double *norm_dev, float *params_dev,
                                     float *sums_dev, void *stream)
{
    HB_REQUIRE(params_dev != nullptr, "hb_fit_same_grid_rows: bad arguments");
    return fit_dispatch(src_dev, src_has_nodata, src_nodata, ref_dev, ref_has_nodata, ref_nodata, h, w, row0, nrows,
                        model, kh, kw, want_r2, norm_dev, params_dev, sums_dev, nullptr, stream,
                        "hb_fit_same_grid_rows");
}

extern "C" int hb_fit_apply_same_grid(const float *src_dev, int src_has_nodata, double src_nodata, const float *ref_dev,
                                      int ref_has_nodata, double ref_nodata, long h, long w, int model, int kh, int kw,
                                      const double *norm_dev, int out_dtype, int out_has_nodata, double out_nodata,
                                      void *corr_dev, void *stream)
{
    HB_REQUIRE(corr_dev != nullptr, "hb_fit_apply_same_grid: bad arguments");
    HB_REQUIRE(hb_outspec_error(out_dtype, out_has_nodata, out_nodata) == nullptr, "hb_fit_apply_same_grid: %s",
               hb_outspec_error(out_dtype, out_has_nodata, out_nodata));
    return fit_dispatch(src_dev, src_has_nodata, src_nodata, ref_dev, ref_has_nodata, ref_nodata, h, w, 0, h, model, kh,
                        kw, 0, norm_dev, nullptr, nullptr, (float *)corr_dev, stream, "hb_fit_apply_same_grid",
                        hb_make_outspec(out_dtype, out_has_nodata, out_nodata));
}

extern "C" int hb_fit_apply_same_grid_rows(const float *src_dev, int src_has_nodata, double src_nodata,
                                           const float *ref_dev, int ref_has_nodata, double ref_nodata, long h, long w,
                                           long row0, long nrows, int model, int kh, int kw, const double *norm_dev,
                                           int out_dtype, int out_has_nodata, double out_nodata, void *corr_dev,
                                           void *stream)
{
    HB_REQUIRE(corr_dev != nullptr, "hb_fit_apply_same_grid_rows: bad arguments");
    HB_REQUIRE(hb_outspec_error(out_dtype, out_has_nodata, out_nodata) == nullptr, "hb_fit_apply_same_grid_rows: %s",
               hb_outspec_error(out_dtype, out_has_nodata, out_nodata));
    return fit_dispatch(src_dev, src_has_nodata, src_nodata, ref_dev, ref_has_nodata, ref_nodata, h, w, row0, nrows,
                        model, kh, kw, 0, norm_dev, nullptr, nullptr, (float *)corr_dev, stream,
                        "hb_fit_apply_same_grid_rows", hb_make_outspec(out_dtype, out_has_nodata, out_nodata));
}
