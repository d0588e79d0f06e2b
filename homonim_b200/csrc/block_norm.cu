// block_norm.cu -- block normalisation statistics of the gain-blk-offset model (sm_100a)
//   KernelModel._fit_block_norm (homonim/kernel_model.py:216-229):
//       norm[0] = std(ref[mask]) / std(src[mask])
//       norm[1] = percentile(ref[mask], 1) - percentile(src[mask], 1) * norm[0]
//   with mask = valid(src) & valid(ref); {0, 0} when the mask is empty.
//
// Exact order statistics come from a 3-level (12 + 12 + 8 bit) radix select on the order-preserving integer image of
// the float32 values: three streaming passes, each building shared-memory histograms (warp-aggregated atomics) that a
// one-CTA "resolve" kernel turns into the next key prefix.  The two ranks numpy interpolates between (k, k + 1) of
// both planes are four simultaneous queries.  The interpolation reproduces numpy >= 2's float32 arithmetic for float32
// input (q = 1/float32(100), virtual index (n - 1) * q in float32, _lerp in float32).
//
// The standard deviations reproduce numpy's float32 np.std TO THE BIT (blocks up to 2^28 pixels): np.std of a float32
// array is mean = float32(pairwise_sum(x) / n), then pairwise_sum((x - mean)^2) / n, square root -- all in float32, where
// pairwise_sum is numpy's recursive halving down to leaves of <= 128 elements summed with 8 interleaved accumulators
// (numpy/_core/src/umath/loops_utils.h.src).  The result depends on that exact association, so it is replayed: the valid
// pixels are compacted in C order (what `array[mask]` hands to np.std), every leaf of the recursion is summed by one
// thread in numpy's order, and the leaves are combined up the same tree.  (The R2 formula of the gain-blk-offset model
// amplifies a 1-ulp change of the block gain ~1000x, which is why "correctly rounded" was not close enough, a6.)
// Larger blocks, and the per-rank accumulators of row-band shards, use double accumulation: means on pass 0, squared
// deviations on pass 1 (within 1 float32 ulp of numpy).
#include <type_traits>

#include "hb_common.cuh"

namespace {

constexpr int kBins = 4096;
constexpr int kNormThreads = 512;

constexpr int kHead = 128, kTail = 64;      // compacted pixels a rank shares with its neighbours (a leaf is <= 128 long and
                                            // starts < 64 before the position that owns it)
constexpr int kMaxWorld = 64;
constexpr int kNodeDepth = 13, kNodes = 1 << kNodeDepth;   // see norm_nodes_kernel

// What one pass over (a shard of) the planes accumulates and the ranks of a row-band sharded raster exchange: this
// struct sits at the start of the workspace, immediately followed by the shard's pairwise-leaf sums -- together the
// "message" of hb_block_norm_message_bytes / hb_block_norm_partial / hb_block_norm_merge.
struct NormAccum {
    unsigned long long n;          // number of valid pixels (== number of compacted pixels)
    double sum[2];                 // sum src, sum ref
    double ssd[2];                 // sum (x - mean)^2
    float2 head[kHead];            // first / last compacted (src, ref) pixels of this shard
    float2 tail[kTail];
    unsigned long long hist[4][kBins];
};

struct NormState {
    float std_exact[2];            // numpy-exact float32 standard deviations (valid when has_exact)
    float meanf[2];                // numpy's float32 means
    unsigned int has_exact;
    int world, rank_id;
    unsigned int pad_;
    unsigned long long n_comp;     // number of compacted (valid) pixels of this shard
    unsigned long long off[kMaxWorld + 1];   // compacted pixels before every shard; off[world] = all of them
    double mean[2];
    unsigned long long rank[4];    // remaining rank inside the current prefix; queries: src k, src k+1, ref k, ref k+1
    unsigned int prefix[4];        // key prefix found so far
    float gamma;                   // numpy's interpolation weight
    unsigned int ticket;
    unsigned int grid_arrivals;    // grid barrier of the cooperative single-launch kernel (zeroed by the host)
    unsigned int pad2_;
};

// pointers into one workspace (see norm_layout)
struct NormWs {
    NormAccum *acc;                // message part 1
    float2 *leaf;                  // message part 2: pairwise-leaf sums of this shard (exact mode), `nleaf` entries
    NormState *st;
    float2 *comp;                  // compacted pixels with kTail entries of headroom in front and kHead behind
    unsigned int *chunk_cnt;
    unsigned long long *chunk_off;
    float2 *nodes;                 // sums of the recursion's nodes at depth kNodeDepth (large blocks: computed by many CTAs)
    unsigned char *node_has;
    long nleaf, nchunks;
    size_t msg_bytes, total;
    int exact;
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

// LEVEL 0: `comp` = (src bin << 12) | ref bin of a valid pixel, 0xffffffff for an invalid one; adds `weight` per active
// lane to hist_s[src bin] and hist_r[ref bin] with one match for the pair of histograms
__device__ __forceinline__ void warp_hist_add_pair(unsigned int *hist_s, unsigned int *hist_r, unsigned int comp,
                                                   unsigned int weight)
{
    const bool active = comp != 0xffffffffu;
    const unsigned int amask = __ballot_sync(0xffffffffu, active);
    if (amask == 0u || !active) return;
    int same = 0;
    __match_all_sync(amask, comp, &same);
    unsigned int peers = amask;
    if (!same) peers = __match_any_sync(amask, comp);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
        const unsigned int c = weight * (unsigned int)__popc(peers);
        atomicAdd(hist_s + (comp >> 12), c);
        atomicAdd(hist_r + (comp & 0xfffu), c);
    }
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
template <int LEVEL> __device__ void norm_resolve(NormAccum *__restrict__ acc, NormState *__restrict__ st, double *__restrict__ norm);

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
// The kernel only accumulates; the level is resolved by norm_merge_kernel once the accumulators of every shard are known
// (a single GPU is the one-shard case of the same sequence).
// the streaming loop of one level for the threads `tid`, `tid + stride`, ... (whole warps stay converged for the votes)
template <int LEVEL, bool VEC>
__device__ __forceinline__ void norm_level_pass(const float *__restrict__ src, const NoData &nd_s,
                                                const float *__restrict__ ref, const NoData &nd_r, long n, long tid,
                                                long stride, unsigned int *s_hist, const unsigned int (&prefix)[4],
                                                double mean_s, double mean_r, double &acc_s, double &acc_r,
                                                unsigned long long &cnt)
{
    long done = 0;                                          // pixels covered by the vector loop
    if (VEC) {
        // 4 consecutive pixels per thread and iteration (16-byte loads of both planes)
        const long n4 = n / 4, n4_round = ((n4 + 31) / 32) * 32;   // keep warps converged for the votes
        // (measured: not load-latency bound -- a 3-deep prefetch ring and 3 CTAs per SM changed nothing; level 0 waits on
        //  the shared-memory atomics / matches, level 1 is bound by integer-pipe issue, level 2 runs at 5.3 TB/s)
        for (long g = tid; g < n4_round; g += stride) {
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = s4;
            const bool in = g < n4;
            if (in) {
                s4 = __ldg(reinterpret_cast<const float4 *>(src) + g);
                r4 = __ldg(reinterpret_cast<const float4 *>(ref) + g);
            }
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
            bool valid[4];
#pragma unroll
            for (int k = 0; k < 4; k++) valid[k] = in && hb_valid(sv[k], nd_s) && hb_valid(rv[k], nd_r);
            // The warp collectives (vote / match) of the histogram updates dominate this loop, not the loads: do them
            // once per 4-pixel group where the group allows it.
            if (LEVEL == 0) {
                // one composite key per pixel -- (src bin, ref bin) -- so that ONE match serves both histograms; when every
                // lane's four pixels share their key (smooth imagery: the common case) one match serves all four
                unsigned int comp[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    comp[k] = valid[k] ? (((float_key(sv[k]) >> 20) << 12) | (float_key(rv[k]) >> 20)) : 0xffffffffu;
                    if (valid[k]) { acc_s += (double)sv[k]; acc_r += (double)rv[k]; cnt++; }
                }
                const bool uni = (comp[0] == comp[1]) && (comp[1] == comp[2]) && (comp[2] == comp[3]);
                if (__all_sync(0xffffffffu, uni)) {
                    warp_hist_add_pair(s_hist, s_hist + 2 * kBins, comp[0], 4u);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) warp_hist_add_pair(s_hist, s_hist + 2 * kBins, comp[k], 1u);
                }
            } else {
                // levels 1 / 2: almost no pixel still matches a query's prefix -- one vote per group decides
                constexpr int sh = (LEVEL == 1) ? 20 : 8;
                bool any_match = false;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned int ks = float_key(sv[k]) >> sh, kr = float_key(rv[k]) >> sh;
                    any_match = any_match || (valid[k] && (ks == prefix[0] || ks == prefix[1] || kr == prefix[2] || kr == prefix[3]));
                }
                if (__ballot_sync(0xffffffffu, any_match) == 0u) {
                    if (LEVEL == 1) {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (valid[k]) {
                                const double ds = (double)sv[k] - mean_s, dr = (double)rv[k] - mean_r;
                                acc_s += ds * ds; acc_r += dr * dr;
                            }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        norm_pixel<LEVEL>(sv[k], rv[k], valid[k], s_hist, prefix, mean_s, mean_r, acc_s, acc_r, cnt);
                }
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
}

template <int LEVEL, bool VEC>
__global__ void __launch_bounds__(kNormThreads)
norm_level_kernel(const float *__restrict__ src, NoData nd_s, const float *__restrict__ ref, NoData nd_r, long n,
                  NormAccum *__restrict__ acc, const NormState *__restrict__ st)
{
    extern __shared__ unsigned int s_hist[];               // [4][bins]
    __shared__ double s_red[kNormThreads / 32];
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
    norm_level_pass<LEVEL, VEC>(src, nd_s, ref, nd_r, n, (long)blockIdx.x * blockDim.x + threadIdx.x,
                                (long)gridDim.x * blockDim.x, s_hist, prefix, mean_s, mean_r, acc_s, acc_r, cnt);
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * bins; i += blockDim.x) {
        const unsigned int c = s_hist[i];
        if (c) atomicAdd(&acc->hist[i / bins][i % bins], (unsigned long long)c);
    }
    if (LEVEL < 2) {
        const double ts = block_sum(acc_s, s_red);
        const double tr = block_sum(acc_r, s_red);
        if (threadIdx.x == 0) {
            if (LEVEL == 0) { atomicAdd(&acc->sum[0], ts); atomicAdd(&acc->sum[1], tr); }
            else { atomicAdd(&acc->ssd[0], ts); atomicAdd(&acc->ssd[1], tr); }
        }
        if (LEVEL == 0) {
            const double tc = block_sum((double)cnt, s_red);
            if (threadIdx.x == 0) atomicAdd(&acc->n, (unsigned long long)(tc + 0.5));
        }
    }
}

// Run by every thread of ONE CTA (norm_merge_kernel): warp q < 4 finds the bin of query q's rank in hist[q]
// (or the shared level-0 histogram), updates the prefix and the remaining rank; all threads then clear the histograms
// for the next level / next call.  After level 2 the four order statistics are known and thread 0 finishes the
// normalisation.  (Global state is read with ld.cg: it was written by other CTAs' atomics.)
template <int LEVEL>
__device__ void norm_resolve(NormAccum *__restrict__ acc, NormState *__restrict__ st, double *__restrict__ norm)
{
    constexpr int bins = (LEVEL == 2) ? 256 : kBins;
    __shared__ float s_val[4];
    const int lane = threadIdx.x & 31;
    const unsigned long long n = __ldcg(&acc->n);

    if (LEVEL == 0 && threadIdx.x == 0) {
        // numpy >= 2, float32 input: q = 1 / float32(100); the "linear" method's virtual index is (n - 1) * q, formed in
        // float32 (numpy/lib/_function_base_impl.py, _QuantileMethods['linear'])
        const float q32 = __fdiv_rn(1.f, 100.f);
        float vi = (n > 0) ? __fmul_rn(__ull2float_rn(n - 1), q32) : 0.f;
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
        st->mean[0] = __ldcg(&acc->sum[0]) / dn; st->mean[1] = __ldcg(&acc->sum[1]) / dn;
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
        unsigned long long part = 0;
        if (n > 0 && t_active) {
#pragma unroll
            for (int i = 0; i < per_thread; i++) part += __ldcg(&acc->hist[hq][threadIdx.x * per_thread + i]);
        }
        local[qq] = part;
        unsigned long long in = part;
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
                // (all of the thread's bins are fetched together -- one L2 round trip; a loop that stops at the winning
                //  bin makes every iteration wait for its own load: ~10 us per level on a proc-grid block)
                unsigned long long cbin[per_thread];
#pragma unroll
                for (int i = 0; i < per_thread; i++) cbin[i] = __ldcg(&acc->hist[hq][threadIdx.x * per_thread + i]);
                unsigned long long below = excl;
                int bin = threadIdx.x * per_thread;
                bool found = false;
#pragma unroll
                for (int i = 0; i < per_thread; i++) {
                    if (!found) {
                        if (rank < below + cbin[i]) { bin = threadIdx.x * per_thread + i; found = true; }
                        else below += cbin[i];
                    }
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
    for (int i = threadIdx.x; i < 4 * kBins; i += blockDim.x) acc->hist[i / kBins][i % kBins] = 0ull;
    if (threadIdx.x == 0) st->ticket = 0u;
    if (LEVEL == 2 && threadIdx.x == 0) {
        double n0 = 0.0, n1 = 0.0;
        if (n > 0) {
            const double dn = (double)n;
            float std_s = (float)sqrt(__ldcg(&acc->ssd[0]) / dn), std_r = (float)sqrt(__ldcg(&acc->ssd[1]) / dn);   // np.std(f32) -> f32
            if (__ldcg(&st->has_exact)) { std_s = __ldcg(&st->std_exact[0]); std_r = __ldcg(&st->std_exact[1]); }
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

// Sum the accumulators of all shards (`gathered`: `world` messages of `msg_bytes` each, in rank order -- a fixed order, so
// every rank derives bit-identical statistics) into this shard's accumulators and resolve the level.  One CTA.
// Level 0 also lays out the compacted pixels of all shards (off[]) and completes this shard's compacted array with the
// neighbours' first / last pixels; levels 1 / 2 first combine the shards' pairwise-leaf sums up numpy's tree (exact mode).
template <int LEVEL> __device__ void norm_tree(const char *gathered, size_t msg_bytes, NormState *st, const float2 *nodes,
                                               const unsigned char *node_has);

template <int LEVEL>
__global__ void __launch_bounds__(kNormThreads)
norm_merge_kernel(NormWs ws, const char *__restrict__ gathered, size_t msg_bytes, int world, int rank_id,
                  double *__restrict__ norm)
{
    NormAccum *acc = ws.acc;
    NormState *st = ws.st;
    constexpr int bins = (LEVEL == 2) ? 256 : kBins;
    constexpr int nq = (LEVEL == 0) ? 3 : 4;                 // level 0 fills hist[0] and hist[2] only
    auto msg = [&](int r) { return reinterpret_cast<const NormAccum *>(gathered + (size_t)r * msg_bytes); };
    if (ws.exact && LEVEL > 0) norm_tree<LEVEL>(gathered, msg_bytes, st, ws.nodes, ws.node_has);   // (nodes: norm_nodes_kernel)
    if (gathered != reinterpret_cast<const char *>(acc)) {   // (one shard: the message IS the accumulator, nothing to add)
        for (int i = threadIdx.x; i < nq * bins; i += blockDim.x) {
            const int q = i / bins, b = i % bins;
            if (LEVEL == 0 && q == 1) continue;
            unsigned long long c = 0;
            for (int r = 0; r < world; r++) c += msg(r)->hist[q][b];
            acc->hist[q][b] = c;
        }
    }
    if (threadIdx.x == 0) {
        if (LEVEL == 0) {
            unsigned long long n = 0;
            double s0 = 0.0, s1 = 0.0;
            for (int r = 0; r < world; r++) {
                st->off[r] = n;
                n += msg(r)->n; s0 += msg(r)->sum[0]; s1 += msg(r)->sum[1];
            }
            st->off[world] = n;
            st->world = world; st->rank_id = rank_id;
            acc->n = n; acc->sum[0] = s0; acc->sum[1] = s1;
        }
        if (LEVEL == 1) {
            double d0 = 0.0, d1 = 0.0;
            for (int r = 0; r < world; r++) { d0 += msg(r)->ssd[0]; d1 += msg(r)->ssd[1]; }
            acc->ssd[0] = d0; acc->ssd[1] = d1;
        }
    }
    __threadfence();
    __syncthreads();
    if (LEVEL == 0 && ws.exact) {
        // neighbours' pixels around this shard's compacted range: kTail before (from the previous shards' tails) and
        // kHead after (from the next shards' heads)
        const unsigned long long my0 = st->off[rank_id], my1 = st->off[rank_id + 1], N = st->off[world];
        for (int q = threadIdx.x; q < kTail + kHead; q += blockDim.x) {
            const bool before = q < kTail;
            const long long gi = before ? (long long)my0 - 1 - q : (long long)my1 + (q - kTail);
            if (gi < 0 || gi >= (long long)N) continue;
            int rr = 0;
            while (rr + 1 < world && st->off[rr + 1] <= (unsigned long long)gi) rr++;
            const unsigned long long nr = st->off[rr + 1] - st->off[rr], li = (unsigned long long)gi - st->off[rr];
            float2 v;
            if (before) { const unsigned long long tl = nr < kTail ? nr : kTail; v = msg(rr)->tail[li - (nr - tl)]; }
            else v = msg(rr)->head[li];
            ws.comp[kTail + (gi - (long long)my0)] = v;
        }
    }
    norm_resolve<LEVEL>(acc, st, norm);
}

__global__ void norm_init_kernel(NormWs ws)
{
    NormAccum *acc = ws.acc;
    NormState *st = ws.st;
    for (int i = threadIdx.x; i < 4 * kBins; i += blockDim.x) acc->hist[i / kBins][i % kBins] = 0ull;
    if (threadIdx.x == 0) {
        acc->n = 0; acc->sum[0] = acc->sum[1] = 0.0; acc->ssd[0] = acc->ssd[1] = 0.0; st->mean[0] = st->mean[1] = 0.0;
        for (int q = 0; q < 4; q++) { st->rank[q] = 0; st->prefix[q] = 0; }
        st->gamma = 0.f;
        st->ticket = 0u;
        st->has_exact = 0u; st->n_comp = 0ull; st->world = 1; st->rank_id = 0;
        st->off[0] = 0ull; st->off[1] = 0ull;
        st->std_exact[0] = st->std_exact[1] = 0.f; st->meanf[0] = st->meanf[1] = 0.f;
    }
}

// ---- numpy-exact float32 np.std ------------------------------------------------------------------------------------------
constexpr int kCompChunk = 2048;              // pixels per compaction chunk (256 threads x 8 consecutive pixels)
constexpr int kCompThreads = 256;
constexpr long kExactMaxPixels = 1L << 28;    // above this the double-accumulated statistics are used

// valid pixels per chunk
__global__ void __launch_bounds__(kCompThreads)
norm_count_kernel(const float *__restrict__ src, NoData nd_s, const float *__restrict__ ref, NoData nd_r, long n,
                  unsigned int *__restrict__ chunk_cnt)
{
    const long base = (long)blockIdx.x * kCompChunk + (long)threadIdx.x * 8;
    int c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const long i = base + j;
        if (i < n) c += (hb_valid(__ldg(src + i), nd_s) && hb_valid(__ldg(ref + i), nd_r)) ? 1 : 0;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    __shared__ int s_w[kCompThreads / 32];
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < kCompThreads / 32; w++) tot += s_w[w];
        chunk_cnt[blockIdx.x] = (unsigned int)tot;
    }
}

// exclusive scan of the chunk counts (one CTA): chunk_off[c] = valid pixels before chunk c; total -> st->n_comp
__global__ void __launch_bounds__(1024)
norm_scan_kernel(NormWs ws)
{
    __shared__ unsigned long long s_part[1024];
    const long nchunks = ws.nchunks;
    const long per = (nchunks + 1023) / 1024;
    const long lo = (long)threadIdx.x * per, hi = min(lo + per, nchunks);
    unsigned long long part = 0;
    for (long c = lo; c < hi; c++) part += ws.chunk_cnt[c];
    s_part[threadIdx.x] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < 1024; t++) { const unsigned long long v = s_part[t]; s_part[t] = run; run += v; }
        ws.st->n_comp = run;
    }
    __syncthreads();
    unsigned long long run = s_part[threadIdx.x];
    for (long c = lo; c < hi; c++) { ws.chunk_off[c] = run; run += ws.chunk_cnt[c]; }
}

// stream compaction in C order: comp[kTail + k] = (src, ref) of the k-th valid pixel of this shard
__global__ void __launch_bounds__(kCompThreads)
norm_scatter_kernel(const float *__restrict__ src, NoData nd_s, const float *__restrict__ ref, NoData nd_r, long n,
                    NormWs ws)
{
    const long base = (long)blockIdx.x * kCompChunk + (long)threadIdx.x * 8;
    float sv[8], rv[8];
    unsigned int vm = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const long i = base + j;
        sv[j] = rv[j] = 0.f;
        if (i < n) {
            sv[j] = __ldg(src + i); rv[j] = __ldg(ref + i);
            if (hb_valid(sv[j], nd_s) && hb_valid(rv[j], nd_r)) vm |= 1u << j;
        }
    }
    const int mine = __popc(vm);
    int incl = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    __shared__ int s_w[kCompThreads / 32];
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += s_w[w];
    unsigned long long dst = kTail + ws.chunk_off[blockIdx.x] + (unsigned long long)(woff + incl - mine);
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (vm & (1u << j)) ws.comp[dst++] = make_float2(sv[j], rv[j]);
}

// first kHead / last kTail compacted pixels of this shard -> the message (one CTA)
__global__ void norm_edges_kernel(NormWs ws)
{
    const unsigned long long n = ws.st->n_comp;
    for (int q = threadIdx.x; q < kHead; q += blockDim.x)
        ws.acc->head[q] = ((unsigned long long)q < n) ? ws.comp[kTail + q] : make_float2(0.f, 0.f);
    const unsigned long long tl = n < kTail ? n : kTail;
    for (int q = threadIdx.x; q < kTail; q += blockDim.x)
        ws.acc->tail[q] = ((unsigned long long)q < tl) ? ws.comp[kTail + (n - tl) + q] : make_float2(0.f, 0.f);
}

// the leaf of numpy's pairwise recursion over [0, n) that contains position p
__device__ __forceinline__ void pw_leaf_of(long n, long p, long &start, long &len)
{
    start = 0; len = n;
    while (len > 128) {
        long n2 = len / 2;
        n2 -= n2 % 8;
        if (p < start + n2) len = n2;
        else { start += n2; len -= n2; }
    }
}

// Positions are GLOBAL indices into the concatenation of all shards' compacted pixels.  One thread per multiple of 64
// inside this shard's range [off[rank], off[rank + 1]); the thread at the first multiple of 64 inside a leaf (every leaf
// of a recursion over more than 128 elements is 64 .. 128 long) sums that leaf in numpy's order -- reading up to 63
// pixels before and 127 after the shard's range from the neighbours' pixels placed there by the level-0 merge.
// SQDEV: sum (x - mean)^2 with numpy's float32 subtract and multiply instead of x.
template <bool SQDEV>
__device__ __forceinline__ void norm_leaf_one(const NormWs &ws, long j)
{
    const NormState *st = ws.st;
    const long my0 = (long)st->off[st->rank_id], my1 = (long)st->off[st->rank_id + 1], N = (long)st->off[st->world];
    const long tbase = (my0 + 63) / 64;
    const long p = (tbase + j) * 64;
    if (j >= ws.nleaf || p >= my1) return;
    long start, len;
    pw_leaf_of(N, p, start, len);
    if (p - start >= 64) return;                       // an earlier multiple of 64 lies inside this leaf
    const float2 *comp = ws.comp + kTail - my0;        // comp[g] for global index g around this shard's range
    const float ms = SQDEV ? st->meanf[0] : 0.f, mr = SQDEV ? st->meanf[1] : 0.f;
    auto term = [&](long g) -> float2 {
        float2 v = comp[g];
        if (SQDEV) {
            const float ds = __fsub_rn(v.x, ms), dr = __fsub_rn(v.y, mr);
            v = make_float2(__fmul_rn(ds, ds), __fmul_rn(dr, dr));
        }
        return v;
    };
    float2 res;
    if (len < 8) {
        res = make_float2(0.f, 0.f);
        for (long i = 0; i < len; i++) { const float2 v = term(start + i); res.x = __fadd_rn(res.x, v.x); res.y = __fadd_rn(res.y, v.y); }
    } else {
        float2 r[8];
#pragma unroll
        for (int k = 0; k < 8; k++) r[k] = term(start + k);
        long i = 8;
        const long lim = len - (len % 8);
        // (32 elements fetched together, added in numpy's order: a leaf is up to 16 dependent rounds of loads otherwise)
        for (; i + 32 <= lim; i += 32) {
            float2 v[32];
#pragma unroll
            for (int k = 0; k < 32; k++) v[k] = term(start + i + k);
#pragma unroll
            for (int k = 0; k < 32; k++) { r[k & 7].x = __fadd_rn(r[k & 7].x, v[k].x); r[k & 7].y = __fadd_rn(r[k & 7].y, v[k].y); }
        }
        for (; i < lim; i += 8) {
#pragma unroll
            for (int k = 0; k < 8; k++) { const float2 v = term(start + i + k); r[k].x = __fadd_rn(r[k].x, v.x); r[k].y = __fadd_rn(r[k].y, v.y); }
        }
        res.x = __fadd_rn(__fadd_rn(__fadd_rn(r[0].x, r[1].x), __fadd_rn(r[2].x, r[3].x)),
                          __fadd_rn(__fadd_rn(r[4].x, r[5].x), __fadd_rn(r[6].x, r[7].x)));
        res.y = __fadd_rn(__fadd_rn(__fadd_rn(r[0].y, r[1].y), __fadd_rn(r[2].y, r[3].y)),
                          __fadd_rn(__fadd_rn(r[4].y, r[5].y), __fadd_rn(r[6].y, r[7].y)));
        for (; i < len; i++) { const float2 v = term(start + i); res.x = __fadd_rn(res.x, v.x); res.y = __fadd_rn(res.y, v.y); }
    }
    ws.leaf[j] = res;
}

template <bool SQDEV>
__global__ void __launch_bounds__(256)
norm_leaf_kernel(NormWs ws)
{
    norm_leaf_one<SQDEV>(ws, (long)blockIdx.x * blockDim.x + threadIdx.x);
}

// leaf sum of the leaf starting at global index `start`: it was computed by the shard holding position ceil64(start)
struct LeafTable {
    const char *gathered; size_t msg_bytes; const NormState *st;
    __device__ float2 at(long start) const
    {
        const long t = (start + 63) / 64;
        int rr = 0;
        // the shard whose range contains position 64 t (shards without pixels own no positions)
        while (rr + 1 < st->world && (long)st->off[rr + 1] <= t * 64) rr++;
        const long tbase = ((long)st->off[rr] + 63) / 64;
        const float2 *leaf = reinterpret_cast<const float2 *>(gathered + (size_t)rr * msg_bytes + sizeof(NormAccum));
        return leaf[t - tbase];
    }
};

// sum of the recursion node [start, start + len) from the leaf sums (numpy: left + right)
__device__ float2 pw_node_sum(const LeafTable &lt, long start, long len)
{
    if (len <= 128) return lt.at(start);
    long n2 = len / 2;
    n2 -= n2 % 8;
    const float2 a = pw_node_sum(lt, start, n2), b = pw_node_sum(lt, start + n2, len - n2);
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}

// the node of numpy's recursion over [0, n) reached from the root by the `depth` bits of `k` (most significant first);
// false when that path ends in a leaf before `depth` steps and k is not the path's first index (the leaf is then held by
// the index whose remaining bits are zero)
__device__ __forceinline__ bool pw_path_node(long n, int k, int depth, long &start, long &len)
{
    start = 0; len = n;
    if (n <= 0) return false;
    for (int level = 0; level < depth; level++) {
        if (len <= 128) return (k & ((1 << (depth - level)) - 1)) == 0;
        long n2 = len / 2;
        n2 -= n2 % 8;
        if ((k >> (depth - 1 - level)) & 1) { start += n2; len -= n2; }
        else len = n2;
    }
    return true;
}

// Large blocks: the sums of the recursion's nodes at depth kNodeDepth, one thread per node, so that the one-CTA merge only
// has the top of the tree left (a block of 9 M pixels has ~94 000 leaves: ~12 per node here, instead of ~180 per thread of
// the merge CTA).
__global__ void __launch_bounds__(256)
norm_nodes_kernel(NormWs ws, const char *__restrict__ gathered, size_t msg_bytes)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kNodes) return;
    const NormState *st = ws.st;
    const LeafTable lt{gathered, msg_bytes, st};
    long start, len;
    const bool mine = pw_path_node((long)st->off[st->world], k, kNodeDepth, start, len);
    ws.node_has[k] = mine ? 1 : 0;
    ws.nodes[k] = mine ? pw_node_sum(lt, start, len) : make_float2(0.f, 0.f);
}

// Combine the leaves up numpy's tree (all threads of the merge CTA): thread t takes the node reached from the root by
// the bits of t (most significant first) and sums its subtree -- from the precomputed depth-kNodeDepth nodes when given,
// else leaf by leaf -- and the top levels are reduced pairwise in shared memory (left + right).  A node that is already
// a leaf above a cut is held by the index whose remaining bits are zero; absent nodes are skipped.
// LEVEL 1: the sums of x -> numpy's float32 means; LEVEL 2: the sums of (x - mean)^2 -> the float32 standard deviations.
template <int LEVEL>
__device__ void norm_tree(const char *gathered, size_t msg_bytes, NormState *st, const float2 *nodes,
                          const unsigned char *node_has)
{
    constexpr int depth = 9, nthreads = 1 << depth;        // == kNormThreads
    static_assert(nthreads == kNormThreads, "one tree node per thread of the merge CTA");
    __shared__ float2 s_val[nthreads];
    __shared__ unsigned char s_has[nthreads];
    const long n = (long)st->off[st->world];
    const int t = threadIdx.x;
    if (nodes != nullptr) {
        constexpr int per = kNodes / nthreads;             // this thread's depth-kNodeDepth nodes: one complete subtree
        float2 v[per];
        bool has[per];
#pragma unroll
        for (int i = 0; i < per; i++) { v[i] = __ldcg(&nodes[t * per + i]); has[i] = __ldcg(&node_has[t * per + i]) != 0; }
#pragma unroll
        for (int sdist = 1; sdist < per; sdist <<= 1) {
#pragma unroll
            for (int i = 0; i < per; i += 2 * sdist) {
                if (has[i] && has[i + sdist]) v[i] = make_float2(__fadd_rn(v[i].x, v[i + sdist].x), __fadd_rn(v[i].y, v[i + sdist].y));
            }
        }
        s_has[t] = has[0] ? 1 : 0;
        s_val[t] = v[0];
    } else {
        const LeafTable lt{gathered, msg_bytes, st};
        long start, len;
        const bool mine = pw_path_node(n, t, depth, start, len);
        s_has[t] = mine ? 1 : 0;
        s_val[t] = mine ? pw_node_sum(lt, start, len) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    for (int sdist = 1; sdist < nthreads; sdist <<= 1) {
        if ((t % (2 * sdist)) == 0 && s_has[t] && s_has[t + sdist]) {
            const float2 a = s_val[t], b = s_val[t + sdist];
            s_val[t] = make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
        }
        __syncthreads();
    }
    if (t == 0 && n > 0) {
        const double dn = (double)n;
        // numpy divides the float32 sum by the integer count in double and stores float32 (_methods._var)
        const float qs = (float)((double)s_val[0].x / dn), qr = (float)((double)s_val[0].y / dn);
        if (LEVEL == 1) { st->meanf[0] = qs; st->meanf[1] = qr; }
        else { st->std_exact[0] = __fsqrt_rn(qs); st->std_exact[1] = __fsqrt_rn(qr); st->has_exact = 1u; }
    }
    __threadfence();
    __syncthreads();
}

// ---- small / mid-size blocks on one GPU: everything in ONE cooperative launch ------------------------------------------------
// On a proc grid of a few hundred pixels a side the 14 launches of the general sequence cost more than their work (ncu, 400 x
// 400: 127 us in kernels of 3 .. 20 us).  Here a few dozen co-resident CTAs (cudaLaunchCooperativeKernel) run the same steps
// back to back -- compaction, level 0, numpy's mean, level 1, numpy's standard deviation, level 2 -- separated by grid
// barriers instead of launches; CTA 0 does the one-CTA steps (resolve, tree).  The very same device functions are called
// (norm_level_pass, norm_resolve, norm_leaf_one, norm_tree), so the result is identical to the general sequence.
constexpr long kCoopMaxPixels = 1L << 22;
constexpr int kCoopMaxCtas = 64;

// all CTAs of the (co-resident) grid; `*counter` starts at 0 and counts arrivals over the whole kernel
__device__ __forceinline__ void norm_grid_sync(unsigned int *counter, unsigned int &epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned int target = (++epoch) * gridDim.x;
        unsigned int seen;
        do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}

#ifdef HB_NORM_PROF
#define NORM_TS(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_ts[i])); } } while (0)
#else
#define NORM_TS(i) do { } while (0)
#endif

template <bool VEC>
__global__ void __launch_bounds__(kNormThreads)
norm_coop_kernel(const float *__restrict__ src, NoData nd_s, const float *__restrict__ ref, NoData nd_r, long n, NormWs ws,
                 double *__restrict__ norm)
{
    extern __shared__ unsigned int s_hist[];               // [4][kBins]
    __shared__ double s_red[kNormThreads / 32];
    __shared__ int s_w[kNormThreads / 32];
    NormAccum *acc = ws.acc;
    NormState *st = ws.st;
    unsigned int *counter = &st->grid_arrivals;            // (zeroed by the host before the launch)
    unsigned int epoch = 0;
#ifdef HB_NORM_PROF
    unsigned long long prof_ts[24];
    int prof_n = 0;
#define NORM_MARK() do { NORM_TS(prof_n); prof_n++; } while (0)
#else
#define NORM_MARK() do { } while (0)
#endif
    NORM_MARK();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long gtid = (long)blockIdx.x * kNormThreads + t, gstride = (long)gridDim.x * kNormThreads;
    // ---- init; valid pixels per chunk of 8 pixels per thread (chunk = kNormThreads * 8 pixels, one CTA at a time) -------------
    constexpr int kChunk = kNormThreads * 8;
    const long nchunks = (n + kChunk - 1) / kChunk;
    if (gtid == 0) {
        acc->n = 0; acc->sum[0] = acc->sum[1] = 0.0; acc->ssd[0] = acc->ssd[1] = 0.0; st->mean[0] = st->mean[1] = 0.0;
        for (int q = 0; q < 4; q++) { st->rank[q] = 0; st->prefix[q] = 0; }
        st->gamma = 0.f; st->has_exact = 0u; st->world = 1; st->rank_id = 0;
        st->std_exact[0] = st->std_exact[1] = 0.f; st->meanf[0] = st->meanf[1] = 0.f;
    }
    for (long i = gtid; i < 4 * kBins; i += gstride) acc->hist[i / kBins][i % kBins] = 0ull;
    auto chunk_valid = [&](long c, float (&sv)[8], float (&rv)[8]) -> unsigned int {
        const long base = c * kChunk + (long)t * 8;
        unsigned int vm = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const long i = base + j;
            sv[j] = rv[j] = 0.f;
            if (i < n) {
                sv[j] = __ldg(src + i); rv[j] = __ldg(ref + i);
                if (hb_valid(sv[j], nd_s) && hb_valid(rv[j], nd_r)) vm |= 1u << j;
            }
        }
        return vm;
    };
    if (ws.exact) {
        for (long c = blockIdx.x; c < nchunks; c += gridDim.x) {
            float sv[8], rv[8];
            int cnt = __popc(chunk_valid(c, sv, rv));
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            __syncthreads();
            if (lane == 0) s_w[warp] = cnt;
            __syncthreads();
            if (t == 0) {
                int tot = 0;
                for (int w = 0; w < kNormThreads / 32; w++) tot += s_w[w];
                ws.chunk_cnt[c] = (unsigned int)tot;
            }
        }
    }
    NORM_MARK();                                           // [1] counted
    norm_grid_sync(counter, epoch);
    NORM_MARK();                                           // [2] sync
    // ---- compaction in C order (every CTA sums the counts of the chunks before its own: there are few) -------------------------
    if (ws.exact) {
        for (long c = blockIdx.x; c < nchunks; c += gridDim.x) {
            unsigned long long before = 0;
            for (long k = lane; k < c; k += 32) before += ws.chunk_cnt[k];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(0xffffffffu, before, d);
            float sv[8], rv[8];
            const unsigned int vm = chunk_valid(c, sv, rv);
            const int mine = __popc(vm);
            int incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            __syncthreads();
            if (lane == 31) s_w[warp] = incl;
            __syncthreads();
            int woff = 0;
            for (int w = 0; w < warp; w++) woff += s_w[w];
            unsigned long long dst = kTail + before + (unsigned long long)(woff + incl - mine);
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (vm & (1u << j)) ws.comp[dst++] = make_float2(sv[j], rv[j]);
            if (c == nchunks - 1 && t == kNormThreads - 1) {
                const unsigned long long total = before + (unsigned long long)(woff + incl);
                st->n_comp = total; st->off[0] = 0ull; st->off[1] = total;
            }
        }
    }
    NORM_MARK();                                           // [3] compacted
    // ---- the three levels ------------------------------------------------------------------------------------------------------
    auto level = [&](auto level_c) {
        constexpr int LEVEL = decltype(level_c)::value;
        constexpr int bins = (LEVEL == 2) ? 256 : kBins;
        for (int i = t; i < 4 * bins; i += kNormThreads) s_hist[i] = 0;
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
        norm_level_pass<LEVEL, VEC>(src, nd_s, ref, nd_r, n, gtid, gstride, s_hist, prefix, mean_s, mean_r, acc_s, acc_r, cnt);
        __syncthreads();
        for (int i = t; i < 4 * bins; i += kNormThreads) {
            const unsigned int c = s_hist[i];
            if (c) atomicAdd(&acc->hist[i / bins][i % bins], (unsigned long long)c);
        }
        if (LEVEL < 2) {
            const double ts = block_sum(acc_s, s_red);
            const double tr = block_sum(acc_r, s_red);
            if (t == 0) {
                if (LEVEL == 0) { atomicAdd(&acc->sum[0], ts); atomicAdd(&acc->sum[1], tr); }
                else { atomicAdd(&acc->ssd[0], ts); atomicAdd(&acc->ssd[1], tr); }
            }
            if (LEVEL == 0) {
                const double tc = block_sum((double)cnt, s_red);
                if (t == 0) atomicAdd(&acc->n, (unsigned long long)(tc + 0.5));
            }
        }
        NORM_MARK();                                       // pass done
        norm_grid_sync(counter, epoch);
        NORM_MARK();                                       // sync
        if (blockIdx.x == 0) {
            if (ws.exact && LEVEL > 0) norm_tree<LEVEL>(reinterpret_cast<const char *>(acc), ws.msg_bytes, st, nullptr, nullptr);
            NORM_MARK();                                   // tree
            norm_resolve<LEVEL>(acc, st, norm);           // (also clears the histograms for the next level)
        }
        NORM_MARK();                                       // resolve
        if (LEVEL < 2) norm_grid_sync(counter, epoch);     // (nothing follows the last level's resolve)
        NORM_MARK();                                       // sync
    };
    auto leaves = [&](auto sq_c) {                         // numpy's pairwise-leaf sums of x / of (x - mean)^2
        constexpr bool SQDEV = decltype(sq_c)::value;
        for (long j = gtid; j < ws.nleaf; j += gstride) norm_leaf_one<SQDEV>(ws, j);
    };
    level(std::integral_constant<int, 0>{});
    if (ws.exact) leaves(std::false_type{});               // (needs off[] and the compacted pixels; its barrier is level 1's)
    level(std::integral_constant<int, 1>{});               // CTA 0: tree<1> -> numpy's means, then the level's resolve
    if (ws.exact) leaves(std::true_type{});
    level(std::integral_constant<int, 2>{});               // CTA 0: tree<2> -> numpy's standard deviations, final resolve
#ifdef HB_NORM_PROF
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        printf("norm_coop n=%ld grid=%d:", n, (int)gridDim.x);
        for (int i = 1; i < prof_n; i++) printf(" %llu", prof_ts[i] - prof_ts[i - 1]);
        printf(" ns\n");
    }
#endif
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
// workspace of a shard of at most `n_local_max` pixels (the largest shard: all shards use the same message size) of a
// block of `n_total` pixels:  [NormAccum | leaf sums] = the exchanged message, then the private state and the compaction
// buffers
NormWs norm_layout(void *workspace, long n_local_max, long n_total)
{
    NormWs ws;
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    ws.exact = (n_total > 0 && n_total <= kExactMaxPixels) ? 1 : 0;
    ws.nleaf = ws.exact ? n_local_max / 64 + 2 : 0;
    ws.nchunks = ws.exact ? (n_local_max + kCompChunk - 1) / kCompChunk : 0;
    ws.msg_bytes = al(sizeof(NormAccum) + (size_t)ws.nleaf * sizeof(float2));
    size_t off = ws.msg_bytes;
    char *base = (char *)workspace;
    ws.acc = (NormAccum *)base;
    ws.leaf = (float2 *)(base + sizeof(NormAccum));
    ws.st = (NormState *)(base + off); off += al(sizeof(NormState));
    ws.comp = (float2 *)(base + off); off += ws.exact ? al((size_t)(n_local_max + kTail + kHead) * sizeof(float2)) : 0;
    ws.chunk_cnt = (unsigned int *)(base + off); off += al((size_t)ws.nchunks * sizeof(unsigned int));
    ws.chunk_off = (unsigned long long *)(base + off); off += al((size_t)ws.nchunks * sizeof(unsigned long long));
    ws.nodes = (float2 *)(base + off); off += ws.exact ? al((size_t)kNodes * sizeof(float2)) : 0;
    ws.node_has = (unsigned char *)(base + off); off += ws.exact ? al((size_t)kNodes) : 0;
    ws.total = off;
    return ws;
}

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

// one streaming pass (level 0, 1 or 2) over n pixels of both planes
int norm_launch_level(int level, const float *src_dev, const NoData &nd_s, const float *ref_dev, const NoData &nd_r, long n,
                      const NormWs &ws, cudaStream_t st)
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
        if (vec) norm_level_kernel<L_, true><<<(unsigned)blocks, kNormThreads, SMEM_, st>>>(src_dev, nd_s, ref_dev, nd_r, n, ws.acc, ws.st); \
        else norm_level_kernel<L_, false><<<(unsigned)blocks, kNormThreads, SMEM_, st>>>(src_dev, nd_s, ref_dev, nd_r, n, ws.acc, ws.st);    \
        HB_LAUNCH_OK("norm_level_kernel");                                                                            \
    } while (0)
    if (level == 0) HB_NORM_LEVEL(0, smem01);
    else if (level == 1) HB_NORM_LEVEL(1, smem01);
    else HB_NORM_LEVEL(2, smem2);
#undef HB_NORM_LEVEL
    return 0;
}

// this shard's part of level `level` (accumulate only)
int norm_partial(int level, const float *src_dev, const NoData &nd_s, const float *ref_dev, const NoData &nd_r, long n,
                 const NormWs &ws, cudaStream_t st)
{
    if (level == 0) {
        norm_init_kernel<<<1, 256, 0, st>>>(ws);
        HB_LAUNCH_OK("norm_init_kernel");
        if (ws.exact) {
            // compact the valid pixels in C order (what `array[mask]` hands to np.std); share the edges with the neighbours
            if (n > 0) {
                const long nchunks = (n + kCompChunk - 1) / kCompChunk;
                NormWs w = ws;
                w.nchunks = nchunks;
                norm_count_kernel<<<(unsigned)nchunks, kCompThreads, 0, st>>>(src_dev, nd_s, ref_dev, nd_r, n, ws.chunk_cnt);
                HB_LAUNCH_OK("norm_count_kernel");
                norm_scan_kernel<<<1, 1024, 0, st>>>(w);
                HB_LAUNCH_OK("norm_scan_kernel");
                norm_scatter_kernel<<<(unsigned)nchunks, kCompThreads, 0, st>>>(src_dev, nd_s, ref_dev, nd_r, n, ws);
                HB_LAUNCH_OK("norm_scatter_kernel");
            }
            norm_edges_kernel<<<1, 128, 0, st>>>(ws);
            HB_LAUNCH_OK("norm_edges_kernel");
        }
    } else if (ws.exact) {
        // pairwise-leaf sums of this shard: of x (for the mean, level 1) / of (x - mean)^2 (for the variance, level 2)
        const unsigned lgrid = (unsigned)((ws.nleaf + 255) / 256);
        if (level == 1) norm_leaf_kernel<false><<<lgrid, 256, 0, st>>>(ws);
        else norm_leaf_kernel<true><<<lgrid, 256, 0, st>>>(ws);
        HB_LAUNCH_OK("norm_leaf_kernel");
    }
    if (n == 0) return 0;                                    // (a shard without pixels contributes empty accumulators)
    return norm_launch_level(level, src_dev, nd_s, ref_dev, nd_r, n, ws, st);
}

int norm_merge(int level, const NormWs &ws, const void *gathered, int world, int rank_id, double *norm_dev, cudaStream_t st)
{
    const char *g = (const char *)gathered;
    if (ws.exact && level > 0) {
        norm_nodes_kernel<<<kNodes / 256, 256, 0, st>>>(ws, g, ws.msg_bytes);
        HB_LAUNCH_OK("norm_nodes_kernel");
    }
    if (level == 0) norm_merge_kernel<0><<<1, kNormThreads, 0, st>>>(ws, g, ws.msg_bytes, world, rank_id, norm_dev);
    else if (level == 1) norm_merge_kernel<1><<<1, kNormThreads, 0, st>>>(ws, g, ws.msg_bytes, world, rank_id, norm_dev);
    else norm_merge_kernel<2><<<1, kNormThreads, 0, st>>>(ws, g, ws.msg_bytes, world, rank_id, norm_dev);
    HB_LAUNCH_OK("norm_merge_kernel");
    return 0;
}

}  // namespace

extern "C" size_t hb_block_norm_workspace_bytes(long n)
{
    return norm_layout(nullptr, n > 0 ? n : 1, n > 0 ? n : 1).total;
}

extern "C" int hb_block_norm(const float *src_dev, int src_has_nodata, double src_nodata, const float *ref_dev,
                             int ref_has_nodata, double ref_nodata, long n, double *norm_dev, void *workspace_dev,
                             size_t workspace_bytes, void *stream)
{
    HB_REQUIRE(src_dev && ref_dev && norm_dev && workspace_dev && n > 0, "hb_block_norm: bad arguments");
    const NormWs ws = norm_layout(workspace_dev, n, n);
    HB_REQUIRE(workspace_bytes >= ws.total, "hb_block_norm: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    HB_REQUIRE(((uintptr_t)workspace_dev) % 256 == 0, "hb_block_norm: workspace must be 256-byte aligned");
    const NoData nd_s = hb_make_nodata(src_has_nodata, src_nodata), nd_r = hb_make_nodata(ref_has_nodata, ref_nodata);
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= kCoopMaxPixels) {
        // one cooperative launch (co-resident CTAs, grid barriers) instead of 14 small ones
        const size_t smem01 = 4 * kBins * sizeof(unsigned int);
        static HbOncePerDevice coop_once;
        const int rc = hb_once_per_device(coop_once, [&]() -> int {
            HB_CUDA_OK(cudaFuncSetAttribute(norm_coop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem01));
            HB_CUDA_OK(cudaFuncSetAttribute(norm_coop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem01));
            return 0;
        });
        if (rc) return rc;
        long ctas = (n + kNormThreads * 8 - 1) / (kNormThreads * 8);      // (one 2048-pixel chunk per CTA and phase up to 131k px)
        if (ctas < 1) ctas = 1;
        if (ctas > kCoopMaxCtas) ctas = kCoopMaxCtas;
        HB_CUDA_OK(cudaMemsetAsync(&ws.st->grid_arrivals, 0, sizeof(unsigned int), st));
        const bool vec = (((uintptr_t)src_dev) % 16 == 0) && (((uintptr_t)ref_dev) % 16 == 0);
        NoData a_nd_s = nd_s, a_nd_r = nd_r;
        long a_n = n;
        NormWs a_ws = ws;
        void *args[] = {(void *)&src_dev, (void *)&a_nd_s, (void *)&ref_dev, (void *)&a_nd_r, (void *)&a_n, (void *)&a_ws,
                        (void *)&norm_dev};
        const void *kern = vec ? (const void *)norm_coop_kernel<true> : (const void *)norm_coop_kernel<false>;
        HB_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3((unsigned)ctas), dim3(kNormThreads), args, smem01, st));
        hb_count_launch();
        return 0;
    }
    // the one-shard case of the sharded sequence: the shard's own message is what the merges read
    for (int level = 0; level < 3; level++) {
        int rc = norm_partial(level, src_dev, nd_s, ref_dev, nd_r, n, ws, st);
        if (rc) return rc;
        rc = norm_merge(level, ws, ws.acc, 1, 0, norm_dev, st);
        if (rc) return rc;
    }
    return 0;
}

// ---- row-band shards: the same three passes, every merge preceded by an exchange of the shards' messages ----------------------
extern "C" size_t hb_block_norm_shard_workspace_bytes(long n_local_max, long n_total)
{
    return norm_layout(nullptr, n_local_max > 0 ? n_local_max : 1, n_total).total;
}

extern "C" size_t hb_block_norm_message_bytes(long n_local_max, long n_total)
{
    return norm_layout(nullptr, n_local_max > 0 ? n_local_max : 1, n_total).msg_bytes;
}

extern "C" int hb_block_norm_partial(int level, const float *src_dev, int src_has_nodata, double src_nodata,
                                     const float *ref_dev, int ref_has_nodata, double ref_nodata, long n_local,
                                     long n_local_max, long n_total, void *workspace_dev, size_t workspace_bytes,
                                     void *stream)
{
    HB_REQUIRE(level >= 0 && level <= 2, "hb_block_norm_partial: level must be 0, 1 or 2");
    HB_REQUIRE(workspace_dev && n_local >= 0 && n_local <= n_local_max && n_local_max <= n_total &&
               (n_local == 0 || (src_dev && ref_dev)), "hb_block_norm_partial: bad arguments");
    const NormWs ws = norm_layout(workspace_dev, n_local_max, n_total);
    HB_REQUIRE(workspace_bytes >= ws.total, "hb_block_norm_partial: workspace too small (%zu < %zu)", workspace_bytes,
               ws.total);
    HB_REQUIRE(((uintptr_t)workspace_dev) % 256 == 0, "hb_block_norm_partial: workspace must be 256-byte aligned");
    const NoData nd_s = hb_make_nodata(src_has_nodata, src_nodata), nd_r = hb_make_nodata(ref_has_nodata, ref_nodata);
    return norm_partial(level, src_dev, nd_s, ref_dev, nd_r, n_local, ws, (cudaStream_t)stream);
}

extern "C" int hb_block_norm_merge(int level, const void *gathered_dev, int world, int rank, long n_local_max,
                                   long n_total, void *workspace_dev, size_t workspace_bytes, double *norm_dev,
                                   void *stream)
{
    HB_REQUIRE(level >= 0 && level <= 2, "hb_block_norm_merge: level must be 0, 1 or 2");
    HB_REQUIRE(gathered_dev && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && workspace_dev && norm_dev,
               "hb_block_norm_merge: bad arguments (at most %d shards)", kMaxWorld);
    const NormWs ws = norm_layout(workspace_dev, n_local_max, n_total);
    HB_REQUIRE(workspace_bytes >= ws.total, "hb_block_norm_merge: workspace too small");
    HB_REQUIRE(((uintptr_t)gathered_dev) % 16 == 0, "hb_block_norm_merge: gathered buffer must be 16-byte aligned");
    return norm_merge(level, ws, gathered_dev, world, rank, norm_dev, (cudaStream_t)stream);
}
