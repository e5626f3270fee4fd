// Microbenchmark: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/ffma2_probe tools/dbg/ffma2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void __launch_bounds__(256) k_probe(float* out, int iters, float seed)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            if (MODE == 0) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y); }
            if (MODE == 1) a[i] = __ffma2_rn(a[i], m, c);
            if (MODE == 2) { a[i] = __ffma2_rn(a[i], m, c); a[i].x = fminf(a[i].x, 3.0e38f); a[i].y = fminf(a[i].y, 3.0e38f); }
            if (MODE == 3) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y); a[i].x = fminf(a[i].x, 3.0e38f); a[i].y = fminf(a[i].y, 3.0e38f); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, float* d)
{
    const int iters = 4096, blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe<MODE><<<blocks, 256>>>(d, 64, 1.0f);
    cudaEventRecord(e0);
    k_probe<MODE><<<blocks, 256>>>(d, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)blocks * 256 * iters * 16;
    printf("%-28s %8.3f ms  %7.2f TFMA/s  (%.1f lane-FMA/clk/SM at 1.965 GHz)\n", name, ms, fmas / ms * 1e-9,
           fmas / (ms * 1e-3) / 148 / 1.965e9);
}

int main()
{
    float* d;
    cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    run<0>("scalar FFMA", d);
    run<1>("packed FFMA2", d);
    run<3>("scalar FFMA + FMNMX", d);
    run<2>("packed FFMA2 + 2 FMNMX", d);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
