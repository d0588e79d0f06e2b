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
//     W(x) = Q(x+hw) - Q(x-hw-1) (+ totals of the warps in between);
//   * the epilogue evaluates the reference's formulas with the reference's precision class and rounding order for
//     every intermediate (SURVEY.md 8a numerics note): sums rounded to float32 where cv.boxFilter returns float32,
//     kept double where cv.sqrBoxFilter returns float64, float32 numerator, double denominator, no FMA contraction.
//   Zero padding beyond the raster (cv BORDER_CONSTANT) falls out of skipping out-of-range rows / columns.
#include "hb_common.cuh"

namespace {

#ifndef HB_FIT_MIN_CTAS
#define HB_FIT_MIN_CTAS 4
#endif
constexpr int kFitThreads = 128;
constexpr int kFitWarps = kFitThreads / 32;
constexpr int kMaxHalfW = 64;                    // kw <= 129

enum { Q_S = 0, Q_R = 1, Q_P = 2, Q_S2 = 3, Q_R2 = 4 };

struct FitGeom {
    long h, w;
    int kh, kw;
    int hw_al;         // half kernel width rounded up to a multiple of C
    int tw_out;        // output columns per CTA
    int rows_per_band;
};

// a / b in double with a quotient that is correctly rounded in all but a vanishing fraction of cases (approximate
// reciprocal + 2 Newton steps + one residual correction: ~9 instructions instead of the ~35 of the IEEE routine).
// Its result is immediately rounded to float32 by the callers, so a last-bit difference in double is invisible
// except on an exact float32 tie.  Zero / non-finite / extreme operands take the IEEE division (x/0 -> +-inf,
// 0/0 -> nan must come out exactly as numpy produces them).
__device__ __forceinline__ double hb_ddiv(double a, double b)
{
    const double ab = fabs(b);
    if (!(ab > 1e-280 && ab < 1e280) || !(fabs(a) < 1e280)) return a / b;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    double q = a * r;
    const double rem = fma(-b, q, a);
    return fma(rem, r, q);
}

// one step of an inclusive warp scan of doubles: v += (value of lane - d), for lanes >= d.  The shuffle's own
// "source lane in range" predicate guards the add (one predicated DADD instead of an add and two selects).
__device__ __forceinline__ double scan_step(double v, int d)
{
    int lo = __double2loint(v), hi = __double2hiint(v), ulo, uhi, p;
    asm volatile("{\n\t.reg .pred q;\n\tshfl.sync.up.b32 %0|q, %3, %5, 0, 0xffffffff;\n\t"
                 "shfl.sync.up.b32 %1, %4, %5, 0, 0xffffffff;\n\tselp.s32 %2, 1, 0, q;\n\t}"
                 : "=r"(ulo), "=r"(uhi), "=r"(p) : "r"(lo), "r"(hi), "r"(d));
    const double up = __hiloint2double(uhi, ulo);
    if (p) v += up;
    return v;
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

// MODEL: HB_MODEL_*; WANT_R2: third band; NQ: number of double sums carried (2, 4 or 5); C: columns per thread
template <int MODEL, bool WANT_R2, int NQ, int C, bool ALIGNED>
__global__ void __launch_bounds__(kFitThreads, (NQ >= 5 && C == 4) ? HB_FIT_MIN_CTAS - 1 : HB_FIT_MIN_CTAS)
fit_same_grid_kernel(const float *__restrict__ src, NoData nd_s, const float *__restrict__ ref, NoData nd_r,
                     FitGeom g, const double *__restrict__ norm, float *__restrict__ params,
                     float *__restrict__ sums_out, float *__restrict__ corr_out)
{
    constexpr bool NORM = (MODEL == HB_MODEL_GAIN_BLK_OFFSET);
    constexpr bool HAS_N = (NQ > 2);
    constexpr int TW = kFitThreads * C;                       // columns per CTA including both halos
    constexpr int WCOLS = 32 * C;                             // columns per warp
    __shared__ double s_q[2][NQ][TW];                         // per-column warp-local inclusive prefixes
    __shared__ int s_n[2][HAS_N ? TW : 1];
    __shared__ double s_tot[2][NQ][kFitWarps];                // per-warp totals
    __shared__ int s_ntot[2][kFitWarps];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int hh = g.kh / 2, hw = g.kw / 2;
    const long tile_start = (long)blockIdx.x * g.tw_out - g.hw_al;   // global column of this CTA's column 0
    const long cx = tile_start + (long)t * C;                        // first global column of this thread
    const long y0 = (long)blockIdx.y * g.rows_per_band;
    const long y1 = min(y0 + (long)g.rows_per_band, g.h);
    const long plane = g.h * g.w;
    const float qnan = __int_as_float(0x7fc00000);
    double n0 = 1.0, n1 = 0.0;
    if (NORM) { n0 = norm[0]; n1 = norm[1]; }

    // validity of a pixel without per-pixel branching on the nodata kind: "equals the nodata value" compares with nan
    // (never true) when there is no value nodata, and the nan test is switched by a kernel-uniform flag
    const float ndv_s = (nd_s.has && !nd_s.is_nan) ? nd_s.value : qnan, ndv_r = (nd_r.has && !nd_r.is_nan) ? nd_r.value : qnan;
    const bool nan_s = nd_s.has && nd_s.is_nan, nan_r = nd_r.has && nd_r.is_nan;
    auto valid2 = [&](float sv, float rv) {
        return !(sv == ndv_s) && !(rv == ndv_r) && !(nan_s && (sv != sv)) && !(nan_r && (rv != rv));
    };
    bool col_in[C];
#pragma unroll
    for (int i = 0; i < C; i++) col_in[i] = (cx + i >= 0) && (cx + i < g.w);
    const bool vec_ok = ALIGNED && (C == 4) && (cx >= 0) && (cx + C <= g.w);
    // this thread's columns are output columns iff they sit between the two halos (whole-thread granularity)
    const bool out_thread = (t * C >= g.hw_al) && (t * C + C <= g.hw_al + g.tw_out);

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

    auto load_row = [&](long y, float (&s)[C], float (&r)[C]) {
        const float *srow = src + y * g.w, *rrow = ref + y * g.w;
        if (vec_ok) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(srow + cx));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(rrow + cx));
            s[0] = a.x; r[0] = b.x;
            if (C > 1) { s[1] = a.y; r[1] = b.y; }
            if (C > 2) { s[2] = a.z; r[2] = b.z; }
            if (C > 3) { s[3] = a.w; r[3] = b.w; }
        } else {
#pragma unroll
            for (int i = 0; i < C; i++) {
                s[i] = col_in[i] ? __ldg(srow + cx + i) : 0.f;
                r[i] = col_in[i] ? __ldg(rrow + cx + i) : 0.f;
            }
        }
    };

    const bool single_row = (g.kh == 1), single_col = (g.kw == 1);
    const long e_first = y0 - hh, e_last = y1 - 1 + hh;
    // the rows of step e + 1 are fetched (into registers) before step e is computed: their latency hides behind a
    // whole row step instead of heading every step's dependency chain
    float se[C], re[C], sl[C], rl[C], nse[C], nre[C], nsl[C], nrl[C];
#pragma unroll
    for (int i = 0; i < C; i++) { se[i] = re[i] = sl[i] = rl[i] = nse[i] = nre[i] = nsl[i] = nrl[i] = 0.f; }
    auto fetch = [&](long e, float (&a)[C], float (&b)[C], float (&c)[C], float (&d)[C]) {
        const long l = e - g.kh;
        if ((e >= 0) && (e < g.h)) load_row(e, a, b);
        if (!single_row && (l >= e_first) && (l >= 0) && (l < g.h)) load_row(l, c, d);
    };
    // ---- warm-up: the rows above the first output row's centre only ENTER the window (nothing leaves, no output row is
    //      completed).  Fetch them 4 at a time -- all loads in flight together -- and accumulate in row order (the same
    //      additions, in the same order, as the row-by-row steps below would perform).
    long e_begin = e_first;
    if (!single_row) {
        const long e_warm_end = y0 + hh;                             // first step that completes an output row
        for (; e_begin < e_warm_end; e_begin += 4) {
            float ws[4][C], wr[4][C];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const long e = e_begin + u;
#pragma unroll
                for (int i = 0; i < C; i++) { ws[u][i] = 0.f; wr[u][i] = 0.f; }
                if (e < e_warm_end && e >= 0 && e < g.h) load_row(e, ws[u], wr[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const long e = e_begin + u;
                if (e >= e_warm_end) break;
                const bool has_e = (e >= 0) && (e < g.h);
#pragma unroll
                for (int i = 0; i < C; i++) {
                    bool ve = false;
                    if (has_e) {
                        double q[NQ]; int cnt;
                        ve = col_in[i] && valid2(ws[u][i], wr[u][i]);
                        pixel_terms<NQ, NORM>(ws[u][i], wr[u][i], ve, n0, n1, q, cnt);
                        ve = cnt != 0;
#pragma unroll
                        for (int k = 0; k < NQ; k++) V[i][k] = __dadd_rn(V[i][k], q[k]);
                        VN[i] += cnt;
                    }
                    vring[i] = (vring[i] << 1) | (ve ? 1ull : 0ull);
                }
            }
        }
        e_begin = e_warm_end;
    }
    fetch(e_begin, se, re, sl, rl);
    for (long e = e_begin; e <= e_last; e++) {
        // ---- vertical running sums: entering row e, leaving row e - kh ---------------------------------------------
        const long l = e - g.kh;
        const bool has_e = (e >= 0) && (e < g.h);
        const bool has_l = !single_row && (l >= e_first) && (l >= 0) && (l < g.h);
        if (e < e_last) fetch(e + 1, nse, nre, nsl, nrl);
#pragma unroll
        for (int i = 0; i < C; i++) {
            bool ve = false;
            if (has_e) {
                double q[NQ]; int cnt;
                ve = col_in[i] && valid2(se[i], re[i]);
                pixel_terms<NQ, NORM>(se[i], re[i], ve, n0, n1, q, cnt);
                ve = cnt != 0;
                if (single_row) {                                    // kh == 1: the window IS this row (exact)
#pragma unroll
                    for (int k = 0; k < NQ; k++) V[i][k] = q[k];
                    VN[i] = cnt;
                } else {
#pragma unroll
                    for (int k = 0; k < NQ; k++) V[i][k] = __dadd_rn(V[i][k], q[k]);
                    VN[i] += cnt;
                }
            } else if (single_row) {
#pragma unroll
                for (int k = 0; k < NQ; k++) V[i][k] = 0.0;
                VN[i] = 0;
            }
            vring[i] = (vring[i] << 1) | (ve ? 1ull : 0ull);
            if (has_l) {
                double q[NQ]; int cnt;
                const bool vl = col_in[i] && valid2(sl[i], rl[i]);
                pixel_terms<NQ, NORM>(sl[i], rl[i], vl, n0, n1, q, cnt);
#pragma unroll
                for (int k = 0; k < NQ; k++) V[i][k] = __dsub_rn(V[i][k], q[k]);
                VN[i] -= cnt;
            }
        }
#pragma unroll
        for (int i = 0; i < C; i++) { se[i] = nse[i]; re[i] = nre[i]; sl[i] = nsl[i]; rl[i] = nrl[i]; }   // rotate
        const long y = e - hh;                                       // output row completed by this step
        if (y < y0) continue;                                        // still filling the first window (uniform)
        const int buf = (int)(y & 1);

        // ---- horizontal window sums: prefix over this thread's columns, warp scan of thread totals ----------------
        // (kw == 1: the window sum is the column sum itself -- exact, and no exchange is needed)
        if (!single_col) {
        double pre[C][NQ];
        int pren[C];
#pragma unroll
        for (int k = 0; k < NQ; k++) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < C; i++) { acc += V[i][k]; pre[i][k] = acc; }
        }
        {
            int acc = 0;
#pragma unroll
            for (int i = 0; i < C; i++) { acc += VN[i]; pren[i] = acc; }
        }
        double excl[NQ];
        int excl_n = 0;
#pragma unroll
        for (int k = 0; k < NQ; k++) {
            double incl = pre[C - 1][k];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) incl = scan_step(incl, d);
            excl[k] = incl - pre[C - 1][k];
            if (lane == 31) s_tot[buf][k][warp] = incl;
        }
        if (HAS_N) {
            int incl = pren[C - 1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            excl_n = incl - pren[C - 1];
            if (lane == 31) s_ntot[buf][warp] = incl;
        }
#pragma unroll
        for (int i = 0; i < C; i++) {
#pragma unroll
            for (int k = 0; k < NQ; k++) s_q[buf][k][i * kFitThreads + t] = excl[k] + pre[i][k];
            if (HAS_N) s_n[buf][i * kFitThreads + t] = excl_n + pren[i];
        }
        __syncthreads();        // (double-buffered: the next row writes the other buffer, so one barrier per row)
        }

        if (!out_thread) continue;
        float o_gain[C], o_off[C], o_r2[C], o_rs[C], o_ss[C], o_n[C];
#pragma unroll
        for (int i = 0; i < C; i++) {
            const int ci = t * C + i;
            const int a = ci + hw, b = ci - hw - 1;
            // column c is stored at slot (c % C) * threads + c / C (conflict-free for both the stores and these loads);
            // the window spans at most two warps (kw <= 129 <= columns per warp + 1): one conditional total
            const int wa = a / WCOLS, wb = (b >= 0) ? b / WCOLS : 0;
            const int sa = (a % C) * kFitThreads + a / C;
            const int sb = (b >= 0) ? (b % C) * kFitThreads + b / C : 0;
            const bool cross = (wa != wb);
            const bool cross2 = (wa - wb) > 1;                        // only when kw > columns per warp
            double W[NQ];
            int N = 0;
            if (single_col) {
#pragma unroll
                for (int k = 0; k < NQ; k++) W[k] = V[i][k];
                N = VN[i];
            } else {
#pragma unroll
                // (all loads unconditional, corrections selected: no divergent branches in the row loop)
                const int wb1 = min(wb + 1, kFitWarps - 1);
                for (int k = 0; k < NQ; k++) {
                    const double qa = s_q[buf][k][sa], qb = s_q[buf][k][sb];
                    const double t0 = s_tot[buf][k][wb], t1 = s_tot[buf][k][wb1];
                    double v = qa + (cross ? t0 : -0.0);     // (x + -0.0 == x and x - 0.0 == x for every x, signed zeros too)
                    v += cross2 ? t1 : -0.0;
                    v -= (b >= 0) ? qb : 0.0;
                    W[k] = v;
                }
                if (HAS_N) {
                    int v = s_n[buf][sa] + (cross ? s_ntot[buf][wb] : 0) + (cross2 ? s_ntot[buf][wb1] : 0);
                    v -= (b >= 0) ? s_n[buf][sb] : 0;
                    N = v;
                }
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
                const float num = __fsub_rn(__fmul_rn(fN, fP), __fmul_rn(fS, fR));                  // :338, float32
                const double den = __dsub_rn(__dmul_rn((double)fN, W[Q_S2]), (double)__fmul_rn(fS, fS));   // :342
                gain = (float)hb_ddiv((double)num, den);                                             // :348
                off = __fdiv_rn(__fsub_rn(fR, __fmul_rn(gain, fS)), fN);                            // :351
                if (WANT_R2) {
                    const double ss_tot = __dsub_rn(__dmul_rn((double)fN, W[Q_R2]), (double)__fmul_rn(fR, fR));   // :179
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
                    ss_res = __dmul_rn(ss_res, (double)fN);                                          // :203
                    r2 = __fsub_rn(1.f, (float)hb_ddiv(ss_res, ss_tot));                            // :212-213
                }
            } else {
                // gain (kernel_model.py:265) -- for gain-blk-offset on the normalised, float64 source sums
                float g0;
                if (NORM) g0 = (float)hb_ddiv((double)fR, W[Q_S]);
                else { fS = (float)W[Q_S]; g0 = __fdiv_rn(fR, fS); }
                if (WANT_R2) {
                    fN = (float)N;
                    const double ss_tot = __dsub_rn(__dmul_rn((double)fN, W[Q_R2]), (double)__fmul_rn(fR, fR));
                    const double t1 = __dmul_rn((double)__fmul_rn(g0, g0), W[Q_S2]);
                    const float g2 = __fmul_rn(2.f, g0);
                    const double t3 = NORM ? __dmul_rn((double)g2, W[Q_P]) : (double)__fmul_rn(g2, (float)W[Q_P]);
                    double ss_res = __dadd_rn(__dsub_rn(t1, t3), W[Q_R2]);                           // :201
                    ss_res = __dmul_rn(ss_res, (double)fN);
                    r2 = __fsub_rn(1.f, (float)hb_ddiv(ss_res, ss_tot));
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
        }
        // ---- store -------------------------------------------------------------------------------------------------
        if (corr_out != nullptr) {
            // fused apply (KernelModel.apply, kernel_model.py:461): corr = gain * src + offset with the centre row's
            // ORIGINAL source pixels (re-read: they entered the window kh/2 rows ago, an L2 hit), two float32
            // roundings as numpy; the parameters are not materialised
            float sc[C], rdummy[C];
            load_row(y, sc, rdummy);
            float oc[C];
#pragma unroll
            for (int i = 0; i < C; i++) oc[i] = __fadd_rn(__fmul_rn(o_gain[i], sc[i]), o_off[i]);
            float *crow = corr_out + y * g.w;
            if (vec_ok && C == 4) {
                *reinterpret_cast<float4 *>(crow + cx) = make_float4(oc[0], oc[C > 1 ? 1 : 0], oc[C > 2 ? 2 : 0], oc[C > 3 ? 3 : 0]);
            } else {
#pragma unroll
                for (int i = 0; i < C; i++)
                    if (col_in[i]) crow[cx + i] = oc[i];
            }
            continue;
        }
        float *prow = params + y * g.w;
        if (vec_ok) {
            if (C == 4) {
                *reinterpret_cast<float4 *>(prow + cx) = make_float4(o_gain[0], o_gain[C > 1 ? 1 : 0],
                                                                     o_gain[C > 2 ? 2 : 0], o_gain[C > 3 ? 3 : 0]);
                *reinterpret_cast<float4 *>(prow + plane + cx) = make_float4(o_off[0], o_off[C > 1 ? 1 : 0],
                                                                             o_off[C > 2 ? 2 : 0], o_off[C > 3 ? 3 : 0]);
                if (WANT_R2)
                    *reinterpret_cast<float4 *>(prow + 2 * plane + cx) =
                        make_float4(o_r2[0], o_r2[C > 1 ? 1 : 0], o_r2[C > 2 ? 2 : 0], o_r2[C > 3 ? 3 : 0]);
                if (sums_out != nullptr) {
                    float *srow = sums_out + y * g.w;
                    *reinterpret_cast<float4 *>(srow + cx) =
                        make_float4(o_rs[0], o_rs[C > 1 ? 1 : 0], o_rs[C > 2 ? 2 : 0], o_rs[C > 3 ? 3 : 0]);
                    *reinterpret_cast<float4 *>(srow + plane + cx) =
                        make_float4(o_ss[0], o_ss[C > 1 ? 1 : 0], o_ss[C > 2 ? 2 : 0], o_ss[C > 3 ? 3 : 0]);
                    *reinterpret_cast<float4 *>(srow + 2 * plane + cx) =
                        make_float4(o_n[0], o_n[C > 1 ? 1 : 0], o_n[C > 2 ? 2 : 0], o_n[C > 3 ? 3 : 0]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < C; i++) {
                if (!col_in[i]) continue;
                prow[cx + i] = o_gain[i];
                prow[plane + cx + i] = o_off[i];
                if (WANT_R2) prow[2 * plane + cx + i] = o_r2[i];
                if (sums_out != nullptr) {
                    float *srow = sums_out + y * g.w;
                    srow[cx + i] = o_rs[i];
                    srow[plane + cx + i] = o_ss[i];
                    srow[2 * plane + cx + i] = o_n[i];
                }
            }
        }
    }
}

template <int MODEL, bool WANT_R2, int NQ, int C>
int launch_fit(const float *src, NoData nd_s, const float *ref, NoData nd_r, long h, long w, int kh, int kw,
               const double *norm, float *params, float *sums, float *corr, cudaStream_t stream)
{
    FitGeom g;
    g.h = h; g.w = w; g.kh = kh; g.kw = kw;
    const int hw = kw / 2;
    g.hw_al = ((hw + C - 1) / C) * C;
    g.tw_out = kFitThreads * C - 2 * g.hw_al;
    const long xtiles = (w + g.tw_out - 1) / g.tw_out;
    // rows per band: every band re-reads (kh - 1) warm-up rows, so bands should be tall -- ~8 window heights -- unless
    // that leaves SMs idle (mid-size rasters: go down to 2 window heights to get one CTA per resident slot).  Small
    // rasters (C == 1) are latency-bound with the machine mostly empty: short bands, many CTAs.
    const long slots = (long)hb_sm_count() * HB_FIT_MIN_CTAS;
    long rpb;
    if (C == 1) {
        const long bands_t = ((long)hb_sm_count() * 8 + xtiles - 1) / xtiles;
        rpb = (h + bands_t - 1) / bands_t;
        if (rpb < 4) rpb = 4;
    } else {
        // score = (fraction of the last wave of resident CTAs that is filled) x (fraction of row fetches that are not
        // warm-up); candidates from 2 to 16 window heights
        double best = -1.0;
        rpb = 8L * kh;
        for (long cand = 16L * kh; cand >= 2L * kh; cand -= (kh + 1) / 2) {
            const long c = cand > h ? h : cand;
            const long ctas = xtiles * ((h + c - 1) / c);
            const long waves = (ctas + slots - 1) / slots;
            const double score = ((double)ctas / (double)(waves * slots)) * ((double)c / (double)(c + kh - 1));
            if (score > best) { best = score; rpb = c; }
        }
    }
    if (rpb > h) rpb = h;
    long bands;
    g.rows_per_band = (int)rpb;
    bands = (h + rpb - 1) / rpb;
    HB_REQUIRE(bands <= 65535, "hb_fit_same_grid: too many row bands");
    dim3 grid((unsigned)xtiles, (unsigned)bands);
    const bool aligned = (C == 4) && (w % 4 == 0) && (((uintptr_t)src) % 16 == 0) && (((uintptr_t)ref) % 16 == 0) &&
                         (((uintptr_t)params) % 16 == 0) && (sums == nullptr || ((uintptr_t)sums) % 16 == 0) &&
                         (corr == nullptr || ((uintptr_t)corr) % 16 == 0);
    if (aligned)
        fit_same_grid_kernel<MODEL, WANT_R2, NQ, C, true><<<grid, kFitThreads, 0, stream>>>(src, nd_s, ref, nd_r, g,
                                                                                           norm, params, sums, corr);
    else
        fit_same_grid_kernel<MODEL, WANT_R2, NQ, C, false><<<grid, kFitThreads, 0, stream>>>(src, nd_s, ref, nd_r, g,
                                                                                            norm, params, sums, corr);
    HB_LAUNCH_OK("fit_same_grid_kernel");
    return 0;
}

template <int MODEL, bool WANT_R2, int NQ>
int launch_fit_c(const float *src, NoData nd_s, const float *ref, NoData nd_r, long h, long w, int kh, int kw,
                 const double *norm, float *params, float *sums, float *corr, cudaStream_t stream)
{
    // small rasters: 1 column per thread (128-column strips) so that the grid still spreads over the SMs
    if (h * w < (long)4 << 20)
        return launch_fit<MODEL, WANT_R2, NQ, 1>(src, nd_s, ref, nd_r, h, w, kh, kw, norm, params, sums, corr, stream);
    return launch_fit<MODEL, WANT_R2, NQ, 4>(src, nd_s, ref, nd_r, h, w, kh, kw, norm, params, sums, corr, stream);
}

}  // namespace

static int fit_dispatch(const float *src_dev, int src_has_nodata, double src_nodata, const float *ref_dev,
                        int ref_has_nodata, double ref_nodata, long h, long w, int model, int kh, int kw, int want_r2,
                        const double *norm_dev, float *params_dev, float *sums_dev, float *corr_dev, void *stream,
                        const char *who)
{
    HB_REQUIRE(src_dev && ref_dev && (params_dev || corr_dev) && h > 0 && w > 0, "%s: bad arguments", who);
    HB_REQUIRE(kh >= 1 && kw >= 1 && (kh & 1) && (kw & 1), "%s: kernel shape must be odd and >= 1", who);
    HB_REQUIRE(kw / 2 <= kMaxHalfW, "%s: kernel width %d > %d is not supported", who, kw, 2 * kMaxHalfW + 1);
    HB_REQUIRE(kh / 2 <= 63, "%s: kernel height %d > 127 is not supported", who, kh);
    const NoData nd_s = hb_make_nodata(src_has_nodata, src_nodata), nd_r = hb_make_nodata(ref_has_nodata, ref_nodata);
    cudaStream_t st = (cudaStream_t)stream;
    switch (model) {
        case HB_MODEL_GAIN:
            HB_REQUIRE(sums_dev == nullptr, "%s: sums are only produced for the gain-offset model", who);
            if (want_r2)
                return launch_fit_c<HB_MODEL_GAIN, true, 5>(src_dev, nd_s, ref_dev, nd_r, h, w, kh, kw, nullptr,
                                                            params_dev, nullptr, corr_dev, st);
            return launch_fit_c<HB_MODEL_GAIN, false, 2>(src_dev, nd_s, ref_dev, nd_r, h, w, kh, kw, nullptr,
                                                         params_dev, nullptr, corr_dev, st);
        case HB_MODEL_GAIN_BLK_OFFSET:
            HB_REQUIRE(norm_dev != nullptr, "%s: gain-blk-offset needs the block normalisation", who);
            HB_REQUIRE(sums_dev == nullptr, "%s: sums are only produced for the gain-offset model", who);
            if (want_r2)
                return launch_fit_c<HB_MODEL_GAIN_BLK_OFFSET, true, 5>(src_dev, nd_s, ref_dev, nd_r, h, w, kh, kw,
                                                                       norm_dev, params_dev, nullptr, corr_dev, st);
            return launch_fit_c<HB_MODEL_GAIN_BLK_OFFSET, false, 2>(src_dev, nd_s, ref_dev, nd_r, h, w, kh, kw,
                                                                    norm_dev, params_dev, nullptr, corr_dev, st);
        case HB_MODEL_GAIN_OFFSET:
            if (want_r2)
                return launch_fit_c<HB_MODEL_GAIN_OFFSET, true, 5>(src_dev, nd_s, ref_dev, nd_r, h, w, kh, kw, nullptr,
                                                                   params_dev, sums_dev, corr_dev, st);
            return launch_fit_c<HB_MODEL_GAIN_OFFSET, false, 4>(src_dev, nd_s, ref_dev, nd_r, h, w, kh, kw, nullptr,
                                                                params_dev, sums_dev, corr_dev, st);
    }
    HB_REQUIRE(false, "%s: unknown model %d", who, model);
}

extern "C" int hb_fit_same_grid(const float *src_dev, int src_has_nodata, double src_nodata, const float *ref_dev,
                                int ref_has_nodata, double ref_nodata, long h, long w, int model, int kh, int kw,
                                int want_r2, const double *norm_dev, float *params_dev, float *sums_dev, void *stream)
{
    HB_REQUIRE(params_dev != nullptr, "hb_fit_same_grid: bad arguments");
    return fit_dispatch(src_dev, src_has_nodata, src_nodata, ref_dev, ref_has_nodata, ref_nodata, h, w, model, kh, kw,
                        want_r2, norm_dev, params_dev, sums_dev, nullptr, stream, "hb_fit_same_grid");
}

extern "C" int hb_fit_apply_same_grid(const float *src_dev, int src_has_nodata, double src_nodata, const float *ref_dev,
                                      int ref_has_nodata, double ref_nodata, long h, long w, int model, int kh, int kw,
                                      const double *norm_dev, float *corr_dev, void *stream)
{
    HB_REQUIRE(corr_dev != nullptr, "hb_fit_apply_same_grid: bad arguments");
    return fit_dispatch(src_dev, src_has_nodata, src_nodata, ref_dev, ref_has_nodata, ref_nodata, h, w, model, kh, kw, 0,
                        norm_dev, nullptr, nullptr, corr_dev, stream, "hb_fit_apply_same_grid");
}
