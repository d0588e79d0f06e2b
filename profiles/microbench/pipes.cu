// pipes.cu -- instruction-throughput microbenchmark (per SM, per clock) for the pipes the moments kernel leans on:
// FP64 add / fma, f32<->f64 conversions, warp shuffles, 64-bit shared-memory loads.  Used to size the kernel design
// (DESIGN.md section 5).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, CH = 8;

template <int OP> __global__ void __launch_bounds__(256) bench(double *out, long long *cycles, float seed)
{
    __shared__ double sm[256 * 2];
    double a[CH]; float f[CH]; int n[CH];
    for (int c = 0; c < CH; c++) { a[c] = seed + c + threadIdx.x; f[c] = seed * c + threadIdx.x; n[c] = threadIdx.x + c; }
    sm[threadIdx.x] = seed; sm[threadIdx.x + 256] = seed + 1;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (OP == 0) a[c] = __dadd_rn(a[c], 1.25);
            if (OP == 1) a[c] = __fma_rn(a[c], 1.0000001, 0.5);
            if (OP == 2) a[c] = __dmul_rn(a[c], 1.0000001);
            if (OP == 3) { a[c] = (double)f[c]; f[c] = __fadd_rn(f[c], (float)i); }               // F2F.F64.F32 + FADD
            if (OP == 4) { f[c] = (float)a[c]; a[c] = __dadd_rn(a[c], (double)1.5); }              // F2F.F32.F64 + DADD
            if (OP == 5) n[c] = __shfl_up_sync(0xffffffffu, n[c], 1) + 1;
            if (OP == 6) f[c] = __fmaf_rn(f[c], 1.0001f, 0.5f);
            if (OP == 7) a[c] += sm[(threadIdx.x + c * 32 + i) & 511];                             // LDS.64 + DADD
            if (OP == 8) { f[c] = (float)n[c]; n[c] += i; }                                        // I2F + IADD
            if (OP == 9) f[c] = __fadd_rn(f[c], 0.5f);
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; c++) s += a[c] + f[c] + n[c];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, int extra_ops)
{
    int dev, sms; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ctas = sms * 8;
    double *out; long long *cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, ctas * 8);
    bench<OP><<<ctas, 256>>>(out, cyc, 1.0f); bench<OP><<<ctas, 256>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long *h = new long long[ctas]; cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; i++) avg += h[i]; avg /= ctas;
    const double per_sm = 8.0 * 256 * ITERS * CH / avg;   // all 8 CTAs of an SM run concurrently
    printf("%-28s %8.1f lane-ops/clk/SM   (%d companion op(s) per measured op)\n", name, per_sm, extra_ops);
    cudaFree(out); cudaFree(cyc); delete[] h;
}

int main()
{
    run<0>("DADD", 0); run<1>("DFMA", 0); run<2>("DMUL", 0);
    run<3>("F2F.F64.F32 (+FADD)", 1); run<4>("F2F.F32.F64 (+DADD)", 1);
    run<5>("SHFL.UP (+IADD)", 1); run<6>("FFMA", 0); run<9>("FADD", 0);
    run<7>("LDS.64 (+DADD)", 1); run<8>("I2F (+IADD)", 1);
    return 0;
}
