// DMMA m8n8k4 throughput of ONE block as a function of its warp count (how the 4 sub-partitions share the FP64 tensor pipe).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_warps dmma_warps.cu && ./dmma_warps
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__global__ void k(double *out, long long *cyc, int iters, int sync_every)
{
    double c[8][2];
    for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i * 0.5; }
    double a0 = 1.0 + threadIdx.x * 1e-6, a1 = 0.5, b0 = 1e-3, b1 = 2e-3;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a0, b0, c[i][0], c[i][1]);
#pragma unroll
        for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a1, b1, c[i][0], c[i][1]);
        if (sync_every) __syncthreads();
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h[4];
    cudaMalloc(&out, 8 * 2048 * 4); cudaMalloc(&cyc, 64);
    const int iters = 2000;
    for (int sync = 0; sync < 2; sync++)
        for (int nw = 1; nw <= 16; nw++) {
            k<<<1, nw * 32>>>(out, cyc, iters, sync);
            cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
            double per_iter = (double)h[0] / iters;   // 16 DMMAs per warp per iteration
            printf("sync %d warps %2d: %.1f clk per 16-DMMA round  -> %.2f clk per DMMA per SM (ideal 4.0)\n", sync, nw, per_iter, per_iter / (16.0 * nw));
        }
    // two blocks of 10 warps on one SM?  (grid of 2 lands on different SMs; skip)
    return 0;
}
