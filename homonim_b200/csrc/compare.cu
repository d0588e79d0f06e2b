// compare.cu -- accuracy sums of RasterCompare: the per-block sums of
// homonim/compare.py:232-256 (`get_block_sums`) over two float32 planes on one grid, as one HBM-bound reduction.
//
//   mask = valid(src) & valid(ref); both planes are zeroed outside it (compare.py:243-247); then
//   sums = { sum(src), sum(ref), sum(src^2), sum(ref^2), sum(src*ref), sum((ref-src)^2), sum(mask) }   (:249-253)
//
// The per-pixel terms are formed in float32 exactly as numpy forms them (products and the difference are rounded to
// float32 before they are summed); the accumulation is double, in a fixed order (deterministic run to run), where
// numpy accumulates the float32 terms pairwise in float32 -- the sums here are the more accurate ones, the reference's
// differ from them by its float32 summation error.
//
// Algorithmic bytes: 8 per pixel (one read of each plane), nothing written but 7 doubles.
#include "hb_common.cuh"

namespace {

constexpr int kCmpThreads = 256;
constexpr int kCmpWarps = kCmpThreads / 32;
constexpr int kCmpSums = 7;
constexpr int kCmpMaxBlocks = 1024;
constexpr size_t kCmpHeader = 16;   // ticket counter (+ padding to keep the partials 16-byte aligned)

struct CmpAcc {
    double s = 0.0, r = 0.0, ss = 0.0, rr = 0.0, sr = 0.0, dd = 0.0;
    unsigned int n = 0u;
};

__device__ __forceinline__ void cmp_pixel(float s, float r, const NoData &nd_s, const NoData &nd_r, CmpAcc &a)
{
    const bool ok = hb_valid(s, nd_s) && hb_valid(r, nd_r);
    const float sv = ok ? s : 0.0f, rv = ok ? r : 0.0f;
    const float d = __fsub_rn(rv, sv);
    a.s += (double)sv;
    a.r += (double)rv;
    a.ss += (double)__fmul_rn(sv, sv);
    a.rr += (double)__fmul_rn(rv, rv);
    a.sr += (double)__fmul_rn(sv, rv);
    a.dd += (double)__fmul_rn(d, d);
    a.n += ok ? 1u : 0u;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Each CTA reduces a grid-strided share of the planes to 7 partial sums; the last CTA to finish (ticket) adds the
// partials of all CTAs in block order.
template <bool VEC>
__global__ void __launch_bounds__(kCmpThreads, 4)
compare_sums_kernel(const float *__restrict__ src, const NoData nd_s, const float *__restrict__ ref, const NoData nd_r,
                    long n, unsigned int *ticket, double *partials, double *sums)
{
    __shared__ double s_part[kCmpWarps][kCmpSums];
    __shared__ bool s_last;
    CmpAcc acc;
    const long tid = (long)blockIdx.x * kCmpThreads + threadIdx.x;
    const long nthreads = (long)gridDim.x * kCmpThreads;
    if (VEC) {
        const long n4 = n >> 2;
        long i = tid;
        // two independent 16-byte loads of each plane in flight per thread
        for (; i + nthreads < n4; i += 2 * nthreads) {
            const uint4 a0 = hb_ldg_stream16(src + 4 * i), b0 = hb_ldg_stream16(ref + 4 * i);
            const uint4 a1 = hb_ldg_stream16(src + 4 * (i + nthreads)), b1 = hb_ldg_stream16(ref + 4 * (i + nthreads));
            cmp_pixel(__uint_as_float(a0.x), __uint_as_float(b0.x), nd_s, nd_r, acc);
            cmp_pixel(__uint_as_float(a0.y), __uint_as_float(b0.y), nd_s, nd_r, acc);
            cmp_pixel(__uint_as_float(a0.z), __uint_as_float(b0.z), nd_s, nd_r, acc);
            cmp_pixel(__uint_as_float(a0.w), __uint_as_float(b0.w), nd_s, nd_r, acc);
            cmp_pixel(__uint_as_float(a1.x), __uint_as_float(b1.x), nd_s, nd_r, acc);
            cmp_pixel(__uint_as_float(a1.y), __uint_as_float(b1.y), nd_s, nd_r, acc);
            cmp_pixel(__uint_as_float(a1.z), __uint_as_float(b1.z), nd_s, nd_r, acc);
            cmp_pixel(__uint_as_float(a1.w), __uint_as_float(b1.w), nd_s, nd_r, acc);
        }
        for (; i < n4; i += nthreads) {
            const uint4 a0 = hb_ldg_stream16(src + 4 * i), b0 = hb_ldg_stream16(ref + 4 * i);
            cmp_pixel(__uint_as_float(a0.x), __uint_as_float(b0.x), nd_s, nd_r, acc);
            cmp_pixel(__uint_as_float(a0.y), __uint_as_float(b0.y), nd_s, nd_r, acc);
            cmp_pixel(__uint_as_float(a0.z), __uint_as_float(b0.z), nd_s, nd_r, acc);
            cmp_pixel(__uint_as_float(a0.w), __uint_as_float(b0.w), nd_s, nd_r, acc);
        }
        for (long j = (n4 << 2) + tid; j < n; j += nthreads) cmp_pixel(src[j], ref[j], nd_s, nd_r, acc);
    } else {
        for (long j = tid; j < n; j += nthreads) cmp_pixel(src[j], ref[j], nd_s, nd_r, acc);
    }

    double v[kCmpSums] = {acc.s, acc.r, acc.ss, acc.rr, acc.sr, acc.dd, (double)acc.n};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < kCmpSums; ++q) {
        const double t = warp_sum(v[q]);
        if (lane == 0) s_part[warp][q] = t;
    }
    __syncthreads();
    if (threadIdx.x < kCmpSums) {
        double t = 0.0;
#pragma unroll
        for (int wi = 0; wi < kCmpWarps; ++wi) t += s_part[wi][threadIdx.x];
        partials[(long)blockIdx.x * kCmpSums + threadIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the last CTA: warp q adds quantity q of every CTA (lane-strided in block order, then the shuffle tree)
    if (warp < kCmpSums) {
        double t = 0.0;
        for (unsigned int b = lane; b < gridDim.x; b += 32) t += __ldcg(partials + (long)b * kCmpSums + warp);
        t = warp_sum(t);
        if (lane == 0) sums[warp] = t;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

}  // namespace

extern "C" size_t hb_compare_sums_workspace_bytes(void)
{
    return kCmpHeader + (size_t)kCmpMaxBlocks * kCmpSums * sizeof(double);
}

extern "C" int hb_compare_sums(const float *src_dev, int src_has_nodata, double src_nodata, const float *ref_dev,
                               int ref_has_nodata, double ref_nodata, long n, double *sums_dev, void *workspace_dev,
                               size_t workspace_bytes, void *stream)
{
    HB_REQUIRE(src_dev && ref_dev && sums_dev && workspace_dev && n > 0, "hb_compare_sums: bad arguments");
    HB_REQUIRE(workspace_bytes >= hb_compare_sums_workspace_bytes(), "hb_compare_sums: workspace too small (%zu < %zu)",
               workspace_bytes, hb_compare_sums_workspace_bytes());
    HB_REQUIRE(((uintptr_t)workspace_dev) % 16 == 0 && ((uintptr_t)sums_dev) % 8 == 0,
               "hb_compare_sums: workspace must be 16-byte aligned and sums 8-byte aligned");
    const NoData nd_s = hb_make_nodata(src_has_nodata, src_nodata), nd_r = hb_make_nodata(ref_has_nodata, ref_nodata);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int *ticket = (unsigned int *)workspace_dev;
    double *partials = (double *)((char *)workspace_dev + kCmpHeader);
    // 8 pixels per thread per trip; whole waves of 4 CTAs per SM
    long blocks = (n + (long)kCmpThreads * 8 - 1) / ((long)kCmpThreads * 8);
    const long cap = (long)hb_sm_count() * 4 < kCmpMaxBlocks ? (long)hb_sm_count() * 4 : (long)kCmpMaxBlocks;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    HB_CUDA_OK(cudaMemsetAsync(ticket, 0, kCmpHeader, st));
    const bool vec = (((uintptr_t)src_dev) % 16 == 0) && (((uintptr_t)ref_dev) % 16 == 0);
    if (vec)
        compare_sums_kernel<true><<<(unsigned)blocks, kCmpThreads, 0, st>>>(src_dev, nd_s, ref_dev, nd_r, n, ticket,
                                                                            partials, sums_dev);
    else
        compare_sums_kernel<false><<<(unsigned)blocks, kCmpThreads, 0, st>>>(src_dev, nd_s, ref_dev, nd_r, n, ticket,
                                                                             partials, sums_dev);
    HB_LAUNCH_OK("compare_sums_kernel");
    return 0;
}
