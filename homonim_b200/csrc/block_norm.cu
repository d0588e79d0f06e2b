// block_norm.cu -- block normalisation statistics of the gain-blk-offset model (sm_100a)
//   KernelModel._fit_block_norm (homonim/kernel_model.py:216-229):
//       norm[0] = std(ref[mask]) / std(src[mask])
//       norm[1] = percentile(ref[mask], 1) - percentile(src[mask], 1) * norm[0]
//   with mask = valid(src) & valid(ref); {0, 0} when the mask is empty.
//
// Exact order statistics come from a 3-level (12 + 12 + 8 bit) radix select on the order-preserving integer image of
// the float32 values: three streaming passes, each building shared-memory histograms (warp-aggregated atomics) that a
// one-CTA "resolve" kernel turns into the next key prefix.  The two ranks numpy interpolates between (k, k + 1) of
// both planes are four simultaneous queries.  Means ride on pass 0, squared deviations (two-pass std, double
// accumulation) on pass 1.  The interpolation reproduces numpy >= 2's float32 arithmetic for float32 input
// (q = 1/float32(100), virtual index n*q + (1 - q) - 1 in float32, _lerp in float32).
#include "hb_common.cuh"

namespace {

constexpr int kBins = 4096;
constexpr int kNormThreads = 512;

struct NormState {
    unsigned long long n;          // number of valid pixels
    double sum[2];                 // sum src, sum ref
    double ssd[2];                 // sum (x - mean)^2
    double mean[2];
    unsigned long long rank[4];    // remaining rank inside the current prefix; queries: src k, src k+1, ref k, ref k+1
    unsigned int prefix[4];        // key prefix found so far
    float gamma;                   // numpy's interpolation weight
    int pad;
    unsigned long long hist[4][kBins];
};

__device__ __forceinline__ unsigned int float_key(float v)
{
    const unsigned int b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k)
{
    const unsigned int b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

// add `1` to hist[bin] for every active lane, one shared-memory atomic per distinct bin in the warp
__device__ __forceinline__ void warp_hist_add(unsigned int *hist, unsigned int bin, bool active)
{
    const unsigned int amask = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned int peers = __match_any_sync(amask, bin);
    const int leader = __ffs(peers) - 1;
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(hist + bin, (unsigned int)__popc(peers));
}

__device__ __forceinline__ double block_sum(double v, double *s_red)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = (lane < (int)(blockDim.x >> 5)) ? s_red[lane] : 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    }
    return t;   // valid in warp 0
}

// LEVEL 0: histogram of key >> 20 for both planes (queries share them: hist[0] = src, hist[2] = ref), count, sums
// LEVEL 1: histogram of (key >> 8) & 0xfff for keys matching each query's 12-bit prefix; squared deviations
// LEVEL 2: histogram of key & 0xff for keys matching each query's 24-bit prefix
template <int LEVEL>
__global__ void __launch_bounds__(kNormThreads)
norm_level_kernel(const float *__restrict__ src, NoData nd_s, const float *__restrict__ ref, NoData nd_r, long n,
                  NormState *__restrict__ st)
{
    extern __shared__ unsigned int s_hist[];               // [4][bins]
    __shared__ double s_red[kNormThreads / 32];
    constexpr int bins = (LEVEL == 2) ? 256 : kBins;
    for (int i = threadIdx.x; i < 4 * bins; i += blockDim.x) s_hist[i] = 0;
    unsigned int prefix[4] = {0, 0, 0, 0};
    double mean_s = 0.0, mean_r = 0.0;
    if (LEVEL > 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) prefix[q] = st->prefix[q];
        mean_s = st->mean[0]; mean_r = st->mean[1];
    }
    __syncthreads();

    double acc_s = 0.0, acc_r = 0.0;
    unsigned long long cnt = 0;
    const long stride = (long)gridDim.x * blockDim.x;
    const long n_round = ((n + 31) / 32) * 32;              // keep warps converged for the ballot / match
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        float s = 0.f, r = 0.f;
        bool valid = false;
        if (i < n) {
            s = __ldg(src + i); r = __ldg(ref + i);
            valid = hb_valid(s, nd_s) && hb_valid(r, nd_r);
        }
        const unsigned int ks = float_key(s), kr = float_key(r);
        if (LEVEL == 0) {
            if (valid) { acc_s += (double)s; acc_r += (double)r; cnt++; }
            warp_hist_add(s_hist + 0 * bins, ks >> 20, valid);
            warp_hist_add(s_hist + 2 * bins, kr >> 20, valid);
        } else if (LEVEL == 1) {
            if (valid) {
                const double ds = (double)s - mean_s, dr = (double)r - mean_r;
                acc_s += ds * ds; acc_r += dr * dr;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned int k = (q < 2) ? ks : kr;
                warp_hist_add(s_hist + q * bins, (k >> 8) & 0xfffu, valid && ((k >> 20) == prefix[q]));
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned int k = (q < 2) ? ks : kr;
                warp_hist_add(s_hist + q * bins, k & 0xffu, valid && ((k >> 8) == prefix[q]));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * bins; i += blockDim.x) {
        const unsigned int c = s_hist[i];
        if (c) atomicAdd(&st->hist[i / bins][i % bins], (unsigned long long)c);
    }
    if (LEVEL < 2) {
        const double ts = block_sum(acc_s, s_red);
        const double tr = block_sum(acc_r, s_red);
        if (threadIdx.x == 0) {
            if (LEVEL == 0) { atomicAdd(&st->sum[0], ts); atomicAdd(&st->sum[1], tr); }
            else { atomicAdd(&st->ssd[0], ts); atomicAdd(&st->ssd[1], tr); }
        }
        if (LEVEL == 0) {
            const double tc = block_sum((double)cnt, s_red);
            if (threadIdx.x == 0) atomicAdd(&st->n, (unsigned long long)(tc + 0.5));
        }
    }
}

// One CTA, 4 warps: warp q finds the bin of query q's rank in hist[q] (or the shared level-0 histogram), updates the
// prefix and the remaining rank, and clears the histogram for the next level.  After level 2 the four order
// statistics are known and thread 0 finishes the normalisation.
template <int LEVEL>
__global__ void __launch_bounds__(128) norm_resolve_kernel(NormState *__restrict__ st, double *__restrict__ norm)
{
    constexpr int bins = (LEVEL == 2) ? 256 : kBins;
    constexpr int per_lane = bins / 32;
    __shared__ float s_val[4];
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long n = st->n;

    if (LEVEL == 0 && threadIdx.x == 0) {
        // numpy >= 2, float32 input: q = 1 / float32(100); virtual index = n*q + (1 + q*(1 - 1 - 1)) - 1 in float32
        const float q32 = __fdiv_rn(1.f, 100.f);
        const float nf = __ull2float_rn(n);
        float vi = __fsub_rn(__fadd_rn(__fmul_rn(nf, q32), __fadd_rn(1.f, __fmul_rn(q32, -1.f))), 1.f);
        if (n == 0) vi = 0.f;
        long long k0 = (long long)floorf(vi);
        float gamma = (float)((double)vi - (double)k0);
        long long k1 = k0 + 1;
        const long long last = (long long)n - 1;
        if (vi >= (float)last) { k0 = last; k1 = last; }       // numpy: indexes above bounds -> last element
        if (k0 < 0) k0 = 0;
        if (k1 < 0) k1 = 0;
        if (k1 > last) k1 = last > 0 ? last : 0;
        st->gamma = gamma;
        st->rank[0] = (unsigned long long)k0; st->rank[1] = (unsigned long long)k1;
        st->rank[2] = (unsigned long long)k0; st->rank[3] = (unsigned long long)k1;
        const double dn = n ? (double)n : 1.0;
        st->mean[0] = st->sum[0] / dn; st->mean[1] = st->sum[1] / dn;
    }
    __syncthreads();
    if (n > 0) {
        const int hq = (LEVEL == 0) ? (q & 2) : q;              // level 0: queries share the per-plane histogram
        const unsigned long long *h = st->hist[hq];
        const unsigned long long rank = st->rank[q];
        unsigned long long local = 0;
        for (int i = 0; i < per_lane; i++) local += h[lane * per_lane + i];
        unsigned long long incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        const unsigned long long excl = incl - local;
        const bool mine = (rank >= excl) && (rank < incl);
        const unsigned int who = __ballot_sync(0xffffffffu, mine);
        if (who != 0 && lane == (__ffs(who) - 1)) {
            unsigned long long below = excl;
            int bin = lane * per_lane;
            for (int i = 0; i < per_lane; i++) {
                const unsigned long long c = h[lane * per_lane + i];
                if (rank < below + c) { bin = lane * per_lane + i; break; }
                below += c;
            }
            st->rank[q] = rank - below;
            const unsigned int p = (LEVEL == 0) ? (unsigned int)bin
                                  : (LEVEL == 1) ? ((st->prefix[q] << 12) | (unsigned int)bin)
                                                 : ((st->prefix[q] << 8) | (unsigned int)bin);
            st->prefix[q] = p;
            if (LEVEL == 2) s_val[q] = key_float(p);
        }
    }
    __syncthreads();
    // clear the histograms for the next level / next call
    for (int i = threadIdx.x; i < 4 * kBins; i += blockDim.x) st->hist[i / kBins][i % kBins] = 0ull;
    if (LEVEL == 2 && threadIdx.x == 0) {
        double n0 = 0.0, n1 = 0.0;
        if (n > 0) {
            const double dn = (double)n;
            const float std_s = (float)sqrt(st->ssd[0] / dn), std_r = (float)sqrt(st->ssd[1] / dn);   // np.std(f32) -> f32
            n0 = (double)__fdiv_rn(std_r, std_s);                                                      // :227
            const float t = st->gamma;
            float p[2];
#pragma unroll
            for (int a = 0; a < 2; a++) {                       // numpy _lerp in float32
                const float lo = s_val[2 * a], hi = s_val[2 * a + 1];
                const float d = __fsub_rn(hi, lo);
                float v = __fadd_rn(lo, __fmul_rn(d, t));
                if (t >= 0.5f) v = __fsub_rn(hi, __fmul_rn(d, __fsub_rn(1.f, t)));
                p[a] = v;
            }
            n1 = __dsub_rn((double)p[1], __dmul_rn((double)p[0], n0));                                 // :228
        }
        norm[0] = n0;
        norm[1] = n1;
    }
}

__global__ void norm_init_kernel(NormState *st)
{
    for (int i = threadIdx.x; i < 4 * kBins; i += blockDim.x) st->hist[i / kBins][i % kBins] = 0ull;
    if (threadIdx.x == 0) {
        st->n = 0; st->sum[0] = st->sum[1] = 0.0; st->ssd[0] = st->ssd[1] = 0.0; st->mean[0] = st->mean[1] = 0.0;
        for (int q = 0; q < 4; q++) { st->rank[q] = 0; st->prefix[q] = 0; }
        st->gamma = 0.f;
    }
}

}  // namespace

extern "C" size_t hb_block_norm_workspace_bytes(long n)
{
    (void)n;
    return sizeof(NormState);
}

extern "C" int hb_block_norm(const float *src_dev, int src_has_nodata, double src_nodata, const float *ref_dev,
                             int ref_has_nodata, double ref_nodata, long n, double *norm_dev, void *workspace_dev,
                             size_t workspace_bytes, void *stream)
{
    HB_REQUIRE(src_dev && ref_dev && norm_dev && workspace_dev && n > 0, "hb_block_norm: bad arguments");
    HB_REQUIRE(workspace_bytes >= sizeof(NormState), "hb_block_norm: workspace too small (%zu < %zu)", workspace_bytes,
               sizeof(NormState));
    HB_REQUIRE(((uintptr_t)workspace_dev) % 8 == 0, "hb_block_norm: workspace must be 8-byte aligned");
    const NoData nd_s = hb_make_nodata(src_has_nodata, src_nodata), nd_r = hb_make_nodata(ref_has_nodata, ref_nodata);
    cudaStream_t st = (cudaStream_t)stream;
    NormState *state = (NormState *)workspace_dev;
    long blocks = (n + kNormThreads - 1) / kNormThreads;
    const long cap = (long)hb_sm_count() * 2;
    if (blocks > cap) blocks = cap;
    const size_t smem01 = 4 * kBins * sizeof(unsigned int), smem2 = 4 * 256 * sizeof(unsigned int);
    static thread_local bool attr_set = false;
    if (!attr_set) {
        HB_CUDA_OK(cudaFuncSetAttribute(norm_level_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem01));
        HB_CUDA_OK(cudaFuncSetAttribute(norm_level_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem01));
        attr_set = true;
    }
    norm_init_kernel<<<1, 256, 0, st>>>(state);
    HB_LAUNCH_OK("norm_init_kernel");
    norm_level_kernel<0><<<(unsigned)blocks, kNormThreads, smem01, st>>>(src_dev, nd_s, ref_dev, nd_r, n, state);
    HB_LAUNCH_OK("norm_level_kernel<0>");
    norm_resolve_kernel<0><<<1, 128, 0, st>>>(state, norm_dev);
    HB_LAUNCH_OK("norm_resolve_kernel<0>");
    norm_level_kernel<1><<<(unsigned)blocks, kNormThreads, smem01, st>>>(src_dev, nd_s, ref_dev, nd_r, n, state);
    HB_LAUNCH_OK("norm_level_kernel<1>");
    norm_resolve_kernel<1><<<1, 128, 0, st>>>(state, norm_dev);
    HB_LAUNCH_OK("norm_resolve_kernel<1>");
    norm_level_kernel<2><<<(unsigned)blocks, kNormThreads, smem2, st>>>(src_dev, nd_s, ref_dev, nd_r, n, state);
    HB_LAUNCH_OK("norm_level_kernel<2>");
    norm_resolve_kernel<2><<<1, 128, 0, st>>>(state, norm_dev);
    HB_LAUNCH_OK("norm_resolve_kernel<2>");
    return 0;
}
