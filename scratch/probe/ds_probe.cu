// ds_probe.cu -- what bounds the average down-sampler's read pattern?  (scratch experiment, not product code)
// Each "tile" = R rows x W bytes of a pitch-P plane (the footprint of ~100 destination pixels of one destination row).
//   mode 0: one CTA per tile, every thread loads R x 16 B in batches of B (the product kernel's pattern), trivial reduce
//   mode 1: persistent CTAs, cp.async.bulk (1-D TMA copy) of whole tile rows into a 2-stage shared-memory ring + mbarrier
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ds_probe ds_probe.cu && ./ds_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int B>
__global__ void __launch_bounds__(256) probe_ldg(const uint4 *src, long pitch16, int rows, int tiles_x, unsigned *out)
{
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;
    const uint4 *p = src + (long)ty * rows * pitch16 + (long)tx * 256 + threadIdx.x;
    unsigned acc = 0;
    for (int r0 = 0; r0 < rows; r0 += B) {
        uint4 v[B];
#pragma unroll
        for (int u = 0; u < B; u++) {
            const uint4 *q = p + (long)min(r0 + u, rows - 1) * pitch16;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(q));
        }
#pragma unroll
        for (int u = 0; u < B; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ void mbar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}

// persistent: tile row = `wbytes` bytes; STAGES-deep ring
template <int STAGES>
__global__ void __launch_bounds__(256) probe_bulk(const char *src, long pitch, int rows, int wbytes, int tiles_x, int ntiles, unsigned *out)
{
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t full[STAGES];
    const int tile_bytes = rows * wbytes;
    if (threadIdx.x == 0) { for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    auto issue = [&](int tile, int s) {
        const int ty = tile / tiles_x, tx = tile % tiles_x;
        const char *p = src + (long)ty * rows * pitch + (long)tx * wbytes;
        mbar_expect(&full[s], (uint32_t)tile_bytes);
        for (int r = 0; r < rows; r++) bulk_g2s(smem + (long)s * tile_bytes + (long)r * wbytes, p + (long)r * pitch, (uint32_t)wbytes, &full[s]);
    };
    int k = 0;
    for (int tile = blockIdx.x; tile < ntiles && k < STAGES - 1; tile += gridDim.x, k++)
        if (threadIdx.x == 0) issue(tile, k);
    unsigned acc = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int s = it % STAGES;
        const int nxt = tile + (STAGES - 1) * gridDim.x;
        if (threadIdx.x == 0 && nxt < ntiles) issue(nxt, (it + STAGES - 1) % STAGES);
        mbar_wait(&full[s], (it / STAGES) & 1);
        const uint4 *t = reinterpret_cast<const uint4 *>(smem + (long)s * tile_bytes);
        for (int i = threadIdx.x; i < tile_bytes / 16; i += 256) { const uint4 v = t[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
        __syncthreads();                         // everybody is done with stage s before it is refilled
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main()
{
    const long H = 10000, Wb = 20000;            // 10000 x 10000 uint16
    char *src; unsigned *out;
    CK(cudaMalloc(&src, H * Wb + 65536)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(src, 1, H * Wb));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int rows = 20, tiles_y = (int)(H / rows);
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        {   // mode 0: 4096-byte tile rows (256 threads x 16 B), 4 full tiles per row + remainder ignored
            const int tiles_x = (int)(Wb / 4096);
            cudaEventRecord(e0);
            for (int i = 0; i < 10; i++) probe_ldg<10><<<tiles_y * tiles_x, 256>>>((const uint4 *)src, Wb / 16, rows, tiles_x, out);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)tiles_y * tiles_x * rows * 4096;
            printf("ldg batch 10 : %.1f us  %.0f GB/s\n", ms * 100, bytes / (ms * 1e-4) / 1e9);
            cudaEventRecord(e0);
            for (int i = 0; i < 10; i++) probe_ldg<20><<<tiles_y * tiles_x, 256>>>((const uint4 *)src, Wb / 16, rows, tiles_x, out);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            printf("ldg batch 20 : %.1f us  %.0f GB/s\n", ms * 100, bytes / (ms * 1e-4) / 1e9);
            cudaEventRecord(e0);
            for (int i = 0; i < 10; i++) probe_ldg<5><<<tiles_y * tiles_x, 256>>>((const uint4 *)src, Wb / 16, rows, tiles_x, out);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            printf("ldg batch 5  : %.1f us  %.0f GB/s\n", ms * 100, bytes / (ms * 1e-4) / 1e9);
        }
        for (int wbytes : {2048, 4000}) {
            const int tiles_x = (int)(Wb / wbytes), ntiles = tiles_y * tiles_x;
            const double bytes = (double)ntiles * rows * wbytes;
            {
                const size_t smem = (size_t)2 * rows * wbytes;
                CK(cudaFuncSetAttribute(probe_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                const int per_sm = (smem <= 100 * 1024) ? 2 : 1;
                cudaEventRecord(e0);
                for (int i = 0; i < 10; i++) probe_bulk<2><<<148 * per_sm, 256, smem>>>(src, Wb, rows, wbytes, tiles_x, ntiles, out);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                printf("bulk 2 stages, %d-byte rows, %d CTA/SM: %.1f us  %.0f GB/s\n", wbytes, per_sm, ms * 100, bytes / (ms * 1e-4) / 1e9);
            }
            {
                const size_t smem = (size_t)3 * rows * wbytes;
                if (smem <= 227 * 1024) {
                    CK(cudaFuncSetAttribute(probe_bulk<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    const int per_sm = (smem <= 100 * 1024) ? 2 : 1;
                    cudaEventRecord(e0);
                    for (int i = 0; i < 10; i++) probe_bulk<3><<<148 * per_sm, 256, smem>>>(src, Wb, rows, wbytes, tiles_x, ntiles, out);
                    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                    printf("bulk 3 stages, %d-byte rows, %d CTA/SM: %.1f us  %.0f GB/s\n", wbytes, per_sm, ms * 100, bytes / (ms * 1e-4) / 1e9);
                }
            }
        }
    }
    CK(cudaGetLastError());
    return 0;
}
