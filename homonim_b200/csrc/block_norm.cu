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

// What one pass over (a shard of) the planes accumulates.  It sits at the start of the workspace, so that the ranks of a
// row-band sharded raster can exchange it (hb_block_norm_accum_bytes / hb_block_norm_partial / hb_block_norm_merge).
struct NormAccum {
    unsigned long long n;          // number of valid pixels
    double sum[2];                 // sum src, sum ref
    double ssd[2];                 // sum (x - mean)^2
    unsigned long long hist[4][kBins];
};

struct NormState : NormAccum {
    double mean[2];
    unsigned long long rank[4];    // remaining rank inside the current prefix; queries: src k, src k+1, ref k, ref k+1
    unsigned int prefix[4];        // key prefix found so far
    float gamma;                   // numpy's interpolation weight
    unsigned int ticket;           // CTAs that have finished the current level (the last one resolves it)
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

// the same, with a fast path for the common case that every active lane of the warp hits the same bin (neighbouring
// pixels of an image share the upper key bits)
__device__ __forceinline__ void warp_hist_add_fast(unsigned int *hist, unsigned int bin, bool active)
{
    const unsigned int amask = __ballot_sync(0xffffffffu, active);
    if (amask == 0u || !active) return;
    int same = 0;
    __match_all_sync(amask, bin, &same);
    if (same) {
        if ((int)(threadIdx.x & 31) == __ffs(amask) - 1) atomicAdd(hist + bin, (unsigned int)__popc(amask));
        return;
    }
    const unsigned int peers = __match_any_sync(amask, bin);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + bin, (unsigned int)__popc(peers));
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
template <int LEVEL> __device__ void norm_resolve(NormState *__restrict__ st, double *__restrict__ norm);

// One pixel's contribution at LEVEL (all lanes of the warp call this together).
template <int LEVEL>
__device__ __forceinline__ void norm_pixel(float s, float r, bool valid, unsigned int *s_hist, const unsigned int (&prefix)[4],
                                           double mean_s, double mean_r, double &acc_s, double &acc_r,
                                           unsigned long long &cnt)
{
    constexpr int bins = (LEVEL == 2) ? 256 : kBins;
    const unsigned int ks = float_key(s), kr = float_key(r);
    if (LEVEL == 0) {
        if (valid) { acc_s += (double)s; acc_r += (double)r; cnt++; }
        warp_hist_add_fast(s_hist + 0 * bins, ks >> 20, valid);
        warp_hist_add_fast(s_hist + 2 * bins, kr >> 20, valid);
    } else {
        if (LEVEL == 1 && valid) {
            const double ds = (double)s - mean_s, dr = (double)r - mean_r;
            acc_s += ds * ds; acc_r += dr * dr;
        }
        // which of the four queries does this pixel's key still match?  (almost always none: skip with one vote)
        constexpr int sh = (LEVEL == 1) ? 20 : 8;
        unsigned int m = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned int k = (q < 2) ? ks : kr;
            if (valid && ((k >> sh) == prefix[q])) m |= 1u << q;
        }
        if (__ballot_sync(0xffffffffu, m != 0u) == 0u) return;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned int k = (q < 2) ? ks : kr;
            const unsigned int bin = (LEVEL == 1) ? ((k >> 8) & 0xfffu) : (k & 0xffu);
            warp_hist_add(s_hist + q * bins, bin, (m >> q) & 1u);
        }
    }
}

// LEVEL 0: histogram of key >> 20 for both planes (queries share them: hist[0] = src, hist[2] = ref), count, sums
// LEVEL 1: histogram of (key >> 8) & 0xfff for keys matching each query's 12-bit prefix; squared deviations
// LEVEL 2: histogram of key & 0xff for keys matching each query's 24-bit prefix
// The last CTA to finish a level resolves it (norm_resolve) -- no separate one-CTA launches between the passes.
// `norm` == nullptr: accumulate only (row-band shards: the level is resolved by norm_merge_kernel once every rank's
// accumulators are known).
template <int LEVEL, bool VEC>
__global__ void __launch_bounds__(kNormThreads)
norm_level_kernel(const float *__restrict__ src, NoData nd_s, const float *__restrict__ ref, NoData nd_r, long n,
                  NormState *__restrict__ st, double *__restrict__ norm)
{
    extern __shared__ unsigned int s_hist[];               // [4][bins]
    __shared__ double s_red[kNormThreads / 32];
    __shared__ int s_last;
    constexpr int bins = (LEVEL == 2) ? 256 : kBins;
    for (int i = threadIdx.x; i < 4 * bins; i += blockDim.x) s_hist[i] = 0;
    unsigned int prefix[4] = {0, 0, 0, 0};
    double mean_s = 0.0, mean_r = 0.0;
    if (LEVEL > 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) prefix[q] = __ldcg(&st->prefix[q]);
        mean_s = __ldcg(&st->mean[0]); mean_r = __ldcg(&st->mean[1]);
    }
    __syncthreads();

    double acc_s = 0.0, acc_r = 0.0;
    unsigned long long cnt = 0;
    const long stride = (long)gridDim.x * blockDim.x;
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long done = 0;                                          // pixels covered by the vector loop
    if (VEC) {
        // 4 consecutive pixels per thread and iteration (16-byte loads of both planes)
        const long n4 = n / 4, n4_round = ((n4 + 31) / 32) * 32;   // keep warps converged for the votes
        for (long g = tid; g < n4_round; g += stride) {
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = s4;
            const bool in = g < n4;
            if (in) {
                s4 = __ldg(reinterpret_cast<const float4 *>(src) + g);
                r4 = __ldg(reinterpret_cast<const float4 *>(ref) + g);
            }
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const bool valid = in && hb_valid(sv[k], nd_s) && hb_valid(rv[k], nd_r);
                norm_pixel<LEVEL>(sv[k], rv[k], valid, s_hist, prefix, mean_s, mean_r, acc_s, acc_r, cnt);
            }
        }
        done = n4 * 4;
    }
    {
        const long rem = n - done, rem_round = ((rem + 31) / 32) * 32;
        for (long i = tid; i < rem_round; i += stride) {
            float s = 0.f, r = 0.f;
            bool valid = false;
            if (i < rem) {
                s = __ldg(src + done + i); r = __ldg(ref + done + i);
                valid = hb_valid(s, nd_s) && hb_valid(r, nd_r);
            }
            norm_pixel<LEVEL>(s, r, valid, s_hist, prefix, mean_s, mean_r, acc_s, acc_r, cnt);
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
    if (norm == nullptr) return;
    // ---- the last CTA to arrive resolves the level -------------------------------------------------------------------
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __threadfence();
        norm_resolve<LEVEL>(st, norm);
    }
}

// Run by every thread of ONE CTA (the last to finish the level): warp q < 4 finds the bin of query q's rank in hist[q]
// (or the shared level-0 histogram), updates the prefix and the remaining rank; all threads then clear the histograms
// for the next level / next call.  After level 2 the four order statistics are known and thread 0 finishes the
// normalisation.  (Global state is read with ld.cg: it was written by other CTAs' atomics.)
template <int LEVEL>
__device__ void norm_resolve(NormState *__restrict__ st, double *__restrict__ norm)
{
    constexpr int bins = (LEVEL == 2) ? 256 : kBins;
    __shared__ float s_val[4];
    const int lane = threadIdx.x & 31;
    const unsigned long long n = __ldcg(&st->n);

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
        st->mean[0] = __ldcg(&st->sum[0]) / dn; st->mean[1] = __ldcg(&st->sum[1]) / dn;
    }
    __syncthreads();
    // every thread sums a few consecutive bins of each query's histogram (independent loads: one L2 round trip), a block
    // scan locates the thread -- and then the bin -- that holds each query's rank
    __shared__ unsigned long long s_wtot[4][kNormThreads / 32];
    constexpr int per_thread = (bins >= kNormThreads) ? bins / kNormThreads : 1;
    const bool t_active = (int)threadIdx.x * per_thread < bins;
    const int warp = threadIdx.x >> 5;
    unsigned long long local[4], incl[4], rk[4];
    unsigned int pf[4];
#pragma unroll
    for (int qq = 0; qq < 4; qq++) {
        // (ranks / prefixes are read by everybody BEFORE the barrier below; the winners overwrite them after it)
        rk[qq] = __ldcg(&st->rank[qq]);
        pf[qq] = __ldcg(&st->prefix[qq]);
        const int hq = (LEVEL == 0) ? (qq & 2) : qq;          // level 0: queries share the per-plane histogram
        unsigned long long acc = 0;
        if (n > 0 && t_active) {
#pragma unroll
            for (int i = 0; i < per_thread; i++) acc += __ldcg(&st->hist[hq][threadIdx.x * per_thread + i]);
        }
        local[qq] = acc;
        unsigned long long in = acc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xffffffffu, in, d);
            if (lane >= d) in += up;
        }
        incl[qq] = in;
        if (lane == 31) s_wtot[qq][warp] = in;
    }
    __syncthreads();
    if (n > 0 && t_active) {
#pragma unroll
        for (int qq = 0; qq < 4; qq++) {
            unsigned long long off = 0;
            for (int w = 0; w < warp; w++) off += s_wtot[qq][w];
            const unsigned long long rank = rk[qq];
            const unsigned long long excl = off + incl[qq] - local[qq];
            if (rank >= excl && rank < excl + local[qq]) {      // exactly one thread per query
                const int hq = (LEVEL == 0) ? (qq & 2) : qq;
                unsigned long long below = excl;
                int bin = threadIdx.x * per_thread;
                for (int i = 0; i < per_thread; i++) {
                    const unsigned long long c = __ldcg(&st->hist[hq][threadIdx.x * per_thread + i]);
                    if (rank < below + c) { bin = threadIdx.x * per_thread + i; break; }
                    below += c;
                }
                const unsigned int old_prefix = pf[qq];
                const unsigned int p = (LEVEL == 0) ? (unsigned int)bin
                                      : (LEVEL == 1) ? ((old_prefix << 12) | (unsigned int)bin)
                                                     : ((old_prefix << 8) | (unsigned int)bin);
                st->rank[qq] = rank - below;
                st->prefix[qq] = p;
                if (LEVEL == 2) s_val[qq] = key_float(p);
            }
        }
    }
    __syncthreads();
    // clear the histograms for the next level / next call
    for (int i = threadIdx.x; i < 4 * kBins; i += blockDim.x) st->hist[i / kBins][i % kBins] = 0ull;
    if (threadIdx.x == 0) st->ticket = 0u;
    if (LEVEL == 2 && threadIdx.x == 0) {
        double n0 = 0.0, n1 = 0.0;
        if (n > 0) {
            const double dn = (double)n;
            const float std_s = (float)sqrt(__ldcg(&st->ssd[0]) / dn), std_r = (float)sqrt(__ldcg(&st->ssd[1]) / dn);   // np.std(f32) -> f32
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

// Row-band shards: sum the accumulators of all ranks (`gathered`: world x NormAccum, in rank order -- a fixed order, so
// every rank gets bit-identical statistics) into this rank's state and resolve the level.  One CTA.
template <int LEVEL>
__global__ void __launch_bounds__(kNormThreads)
norm_merge_kernel(NormState *__restrict__ st, const NormAccum *__restrict__ gathered, int world, double *__restrict__ norm)
{
    constexpr int bins = (LEVEL == 2) ? 256 : kBins;
    constexpr int nq = (LEVEL == 0) ? 3 : 4;                 // level 0 fills hist[0] and hist[2] only
    for (int i = threadIdx.x; i < nq * bins; i += blockDim.x) {
        const int q = i / bins, b = i % bins;
        if (LEVEL == 0 && q == 1) continue;
        unsigned long long c = 0;
        for (int r = 0; r < world; r++) c += gathered[r].hist[q][b];
        st->hist[q][b] = c;
    }
    if (threadIdx.x == 0) {
        if (LEVEL == 0) {
            unsigned long long n = 0;
            double s0 = 0.0, s1 = 0.0;
            for (int r = 0; r < world; r++) { n += gathered[r].n; s0 += gathered[r].sum[0]; s1 += gathered[r].sum[1]; }
            st->n = n; st->sum[0] = s0; st->sum[1] = s1;
        }
        if (LEVEL == 1) {
            double d0 = 0.0, d1 = 0.0;
            for (int r = 0; r < world; r++) { d0 += gathered[r].ssd[0]; d1 += gathered[r].ssd[1]; }
            st->ssd[0] = d0; st->ssd[1] = d1;
        }
    }
    __threadfence();
    __syncthreads();
    norm_resolve<LEVEL>(st, norm);
}

__global__ void norm_init_kernel(NormState *st)
{
    for (int i = threadIdx.x; i < 4 * kBins; i += blockDim.x) st->hist[i / kBins][i % kBins] = 0ull;
    if (threadIdx.x == 0) {
        st->n = 0; st->sum[0] = st->sum[1] = 0.0; st->ssd[0] = st->ssd[1] = 0.0; st->mean[0] = st->mean[1] = 0.0;
        for (int q = 0; q < 4; q++) { st->rank[q] = 0; st->prefix[q] = 0; }
        st->gamma = 0.f;
        st->ticket = 0u;
    }
}

}  // namespace

extern "C" size_t hb_block_norm_workspace_bytes(long n)
{
    (void)n;
    return sizeof(NormState);
}

namespace {

int norm_setup_once(size_t smem01)
{
    static HbOncePerDevice attr_once;
    return hb_once_per_device(attr_once, [&]() -> int {
        HB_CUDA_OK(cudaFuncSetAttribute(norm_level_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem01));
        HB_CUDA_OK(cudaFuncSetAttribute(norm_level_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem01));
        HB_CUDA_OK(cudaFuncSetAttribute(norm_level_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem01));
        HB_CUDA_OK(cudaFuncSetAttribute(norm_level_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem01));
        return 0;
    });
}

// one streaming pass (level 0, 1 or 2) over n pixels of both planes; `norm_dev` == nullptr: accumulate only
int norm_launch_level(int level, const float *src_dev, const NoData &nd_s, const float *ref_dev, const NoData &nd_r, long n,
                      NormState *state, double *norm_dev, cudaStream_t st)
{
    long blocks = (n / 4 + kNormThreads - 1) / kNormThreads;
    if (blocks < 1) blocks = 1;
    const long cap = (long)hb_sm_count() * 2;
    if (blocks > cap) blocks = cap;
    const size_t smem01 = 4 * kBins * sizeof(unsigned int), smem2 = 4 * 256 * sizeof(unsigned int);
    const int rc = norm_setup_once(smem01);
    if (rc) return rc;
    const bool vec = (((uintptr_t)src_dev) % 16 == 0) && (((uintptr_t)ref_dev) % 16 == 0);
#define HB_NORM_LEVEL(L_, SMEM_)                                                                                      \
    do {                                                                                                              \
        if (vec) norm_level_kernel<L_, true><<<(unsigned)blocks, kNormThreads, SMEM_, st>>>(src_dev, nd_s, ref_dev, nd_r, n, state, norm_dev); \
        else norm_level_kernel<L_, false><<<(unsigned)blocks, kNormThreads, SMEM_, st>>>(src_dev, nd_s, ref_dev, nd_r, n, state, norm_dev);    \
        HB_LAUNCH_OK("norm_level_kernel");                                                                            \
    } while (0)
    if (level == 0) HB_NORM_LEVEL(0, smem01);
    else if (level == 1) HB_NORM_LEVEL(1, smem01);
    else HB_NORM_LEVEL(2, smem2);
#undef HB_NORM_LEVEL
    return 0;
}

}  // namespace

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
    norm_init_kernel<<<1, 256, 0, st>>>(state);
    HB_LAUNCH_OK("norm_init_kernel");
    for (int level = 0; level < 3; level++) {
        const int rc = norm_launch_level(level, src_dev, nd_s, ref_dev, nd_r, n, state, norm_dev, st);
        if (rc) return rc;
    }
    return 0;
}

// ---- row-band shards: the same three passes with the per-level resolve replaced by an exchange between the ranks ------
extern "C" size_t hb_block_norm_accum_bytes(void) { return sizeof(NormAccum); }

extern "C" int hb_block_norm_partial(int level, const float *src_dev, int src_has_nodata, double src_nodata,
                                     const float *ref_dev, int ref_has_nodata, double ref_nodata, long n,
                                     void *workspace_dev, size_t workspace_bytes, void *stream)
{
    HB_REQUIRE(level >= 0 && level <= 2, "hb_block_norm_partial: level must be 0, 1 or 2");
    HB_REQUIRE(workspace_dev && n >= 0 && (n == 0 || (src_dev && ref_dev)), "hb_block_norm_partial: bad arguments");
    HB_REQUIRE(workspace_bytes >= sizeof(NormState), "hb_block_norm_partial: workspace too small (%zu < %zu)",
               workspace_bytes, sizeof(NormState));
    HB_REQUIRE(((uintptr_t)workspace_dev) % 8 == 0, "hb_block_norm_partial: workspace must be 8-byte aligned");
    const NoData nd_s = hb_make_nodata(src_has_nodata, src_nodata), nd_r = hb_make_nodata(ref_has_nodata, ref_nodata);
    cudaStream_t st = (cudaStream_t)stream;
    NormState *state = (NormState *)workspace_dev;
    if (level == 0) {
        norm_init_kernel<<<1, 256, 0, st>>>(state);
        HB_LAUNCH_OK("norm_init_kernel");
    }
    if (n == 0) return 0;                                    // (a rank without rows contributes empty accumulators)
    return norm_launch_level(level, src_dev, nd_s, ref_dev, nd_r, n, state, nullptr, st);
}

extern "C" int hb_block_norm_merge(int level, const void *gathered_dev, int world, void *workspace_dev,
                                   size_t workspace_bytes, double *norm_dev, void *stream)
{
    HB_REQUIRE(level >= 0 && level <= 2, "hb_block_norm_merge: level must be 0, 1 or 2");
    HB_REQUIRE(gathered_dev && world >= 1 && workspace_dev && norm_dev, "hb_block_norm_merge: bad arguments");
    HB_REQUIRE(workspace_bytes >= sizeof(NormState), "hb_block_norm_merge: workspace too small");
    HB_REQUIRE(((uintptr_t)gathered_dev) % 8 == 0, "hb_block_norm_merge: gathered buffer must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    NormState *state = (NormState *)workspace_dev;
    const NormAccum *g = (const NormAccum *)gathered_dev;
    if (level == 0) norm_merge_kernel<0><<<1, kNormThreads, 0, st>>>(state, g, world, norm_dev);
    else if (level == 1) norm_merge_kernel<1><<<1, kNormThreads, 0, st>>>(state, g, world, norm_dev);
    else norm_merge_kernel<2><<<1, kNormThreads, 0, st>>>(state, g, world, norm_dev);
    HB_LAUNCH_OK("norm_merge_kernel");
    return 0;
}
