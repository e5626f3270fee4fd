// scratch: which formulation of the FAST corner score compiles to correct code on sm_100a?
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

__device__ __constant__ int c_off[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};

__device__ int score_a(const unsigned char* ring, int v, int t)
{
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - (int)ring[c_off[k]];
    int a0 = t, b0 = t;
#pragma unroll
    for (int s = 0; s < 16; s++)
    {
        int mn = d[s], mx = d[s];
#pragma unroll
        for (int j = 1; j < 9; j++)
        {
            const int e = d[(s + j) & 15];
            mn = min(mn, e);
            mx = max(mx, e);
        }
        a0 = max(a0, mn);
        b0 = max(b0, -mx);
    }
    return max(a0, b0) - 1;
}

__device__ int score_b(const unsigned char* ring, int v, int t)
{
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - (int)ring[c_off[k]];
    int lo2[16], hi2[16], lo4[16], hi4[16], lo8[16], hi8[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { lo2[k] = min(d[k], d[(k + 1) & 15]); hi2[k] = max(d[k], d[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; k++) { lo4[k] = min(lo2[k], lo2[(k + 2) & 15]); hi4[k] = max(hi2[k], hi2[(k + 2) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; k++) { lo8[k] = min(lo4[k], lo4[(k + 4) & 15]); hi8[k] = max(hi4[k], hi4[(k + 4) & 15]); }
    int a0 = t, b0 = t;
#pragma unroll
    for (int k = 0; k < 16; k++)
    {
        a0 = max(a0, min(lo8[k], d[(k + 8) & 15]));
        b0 = max(b0, -max(hi8[k], d[(k + 8) & 15]));
    }
    return max(a0, b0) - 1;
}

// OpenCV cornerScore<16> literally (d has 25 entries)
__device__ int score_c(const unsigned char* ring, int v, int t)
{
    int d[25];
#pragma unroll
    for (int k = 0; k < 25; k++) d[k] = v - (int)ring[c_off[k & 15]];
    int a0 = t;
#pragma unroll
    for (int k = 0; k < 16; k += 2)
    {
        int a = min(d[k + 1], d[k + 2]);
        a = min(a, d[k + 3]);
        if (a <= a0) continue;
        a = min(a, d[k + 4]); a = min(a, d[k + 5]); a = min(a, d[k + 6]); a = min(a, d[k + 7]); a = min(a, d[k + 8]);
        a0 = max(a0, min(a, d[k]));
        a0 = max(a0, min(a, d[k + 9]));
    }
    int b0 = -a0;
#pragma unroll
    for (int k = 0; k < 16; k += 2)
    {
        int b = max(d[k + 1], d[k + 2]);
        b = max(b, d[k + 3]); b = max(b, d[k + 4]); b = max(b, d[k + 5]);
        if (b >= b0) continue;
        b = max(b, d[k + 6]); b = max(b, d[k + 7]); b = max(b, d[k + 8]);
        b0 = min(b0, max(b, d[k]));
        b0 = min(b0, max(b, d[k + 9]));
    }
    return -b0 - 1;
}

__global__ void k(const unsigned char* rings, const unsigned char* vs, int n, int t, int* out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[3 * i] = score_a(rings + 16 * i, vs[i], t);
    out[3 * i + 1] = score_b(rings + 16 * i, vs[i], t);
    out[3 * i + 2] = score_c(rings + 16 * i, vs[i], t);
}

static int host_score(const unsigned char* ring, int v, int t)
{
    int d[16];
    for (int k = 0; k < 16; k++) d[k] = v - ring[k];
    int a0 = t, b0 = t;
    for (int s = 0; s < 16; s++)
    {
        int mn = d[s], mx = d[s];
        for (int j = 1; j < 9; j++) { int e = d[(s + j) & 15]; mn = std::min(mn, e); mx = std::max(mx, e); }
        a0 = std::max(a0, mn); b0 = std::max(b0, -mx);
    }
    return std::max(a0, b0) - 1;
}

int main()
{
    const int n = 1 << 16, t = 20;
    std::vector<unsigned char> rings(16 * n), vs(n);
    srand(1);
    for (int i = 0; i < n; i++)
    {
        vs[i] = rand() & 255;
        int base = rand() & 255, amp = rand() % 60;
        for (int k = 0; k < 16; k++) rings[16 * i + k] = (unsigned char)std::min(255, std::max(0, base + (rand() % (2 * amp + 1)) - amp));
    }
    unsigned char *dr, *dv; int* dout;
    cudaMalloc(&dr, rings.size()); cudaMalloc(&dv, n); cudaMalloc(&dout, 3 * n * sizeof(int));
    cudaMemcpy(dr, rings.data(), rings.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dv, vs.data(), n, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(dr, dv, n, t, dout);
    std::vector<int> out(3 * n);
    cudaError_t e = cudaMemcpy(out.data(), dout, out.size() * sizeof(int), cudaMemcpyDeviceToHost);
    int bad[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
    {
        const int ref = host_score(&rings[16 * i], vs[i], t);
        for (int v = 0; v < 3; v++) if (out[3 * i + v] != ref) { if (bad[v] < 3) printf("variant %d case %d: gpu %d ref %d\n", v, i, out[3 * i + v], ref); bad[v]++; }
    }
    printf("cuda %s; mismatches: loop %d, doubling %d, opencv-literal %d of %d\n", cudaGetErrorString(e), bad[0], bad[1], bad[2], n);
    return 0;
}
