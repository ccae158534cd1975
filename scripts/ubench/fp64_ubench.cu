// B200 FP64 micro-benchmarks (latency / throughput of the primitives the block solver is built from).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void lat_dfma(double *out, long long *cyc, int n)
{
    double a = out[0], b = out[1], c = out[2];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { c = fma(a, c, b); c = fma(a, c, b); c = fma(a, c, b); c = fma(a, c, b); }
    long long t1 = clock64();
    out[3] = c; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void thr_dfma(double *out, long long *cyc, int n)
{
    double c[ILP]; double a = out[0], b = out[1];
    for (int q = 0; q < ILP; q++) c[q] = out[2] + q + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int q = 0; q < ILP; q++) c[q] = fma(a, c[q], b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int q = 0; q < ILP; q++) s += c[q];
    out[4 + threadIdx.x % 4] = s; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_rcp(double *out, long long *cyc, int n)
{
    double c = out[2];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { c = 1.0 / c + 1.5; }
    long long t1 = clock64();
    out[3] = c; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_redux(unsigned *out, long long *cyc, int n)
{
    unsigned v = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { v = __reduce_max_sync(0xffffffffu, v + threadIdx.x) ; }
    long long t1 = clock64();
    out[32] = v; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_shfl(double *out, long long *cyc, int n)
{
    double v = out[threadIdx.x % 8];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { v = __shfl_sync(0xffffffffu, v, (threadIdx.x + 5) & 31) + 1.0; }
    long long t1 = clock64();
    out[8] = v; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_bar(double *out, long long *cyc, int n)
{
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { __syncthreads(); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_lds(double *out, long long *cyc, int n)
{
    __shared__ int idx[256];
    idx[threadIdx.x] = (threadIdx.x * 7 + 1) & 255;
    __syncthreads();
    int j = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { j = idx[j]; }
    long long t1 = clock64();
    out[9] = j; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_sts_bar_lds(double *out, long long *cyc, int n)
{
    __shared__ double buf[2][64];
    double v = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (threadIdx.x < 32) buf[i & 1][threadIdx.x] = v;
        __syncthreads();
        v = buf[i & 1][(threadIdx.x + 3) & 31] + 1.0;
    }
    long long t1 = clock64();
    out[10] = v; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// DMMA m8n8k4 throughput: each warp issues ILP independent accumulators
template <int ILP>
__global__ void thr_dmma(double *out, long long *cyc, int n)
{
    double a = out[0] + threadIdx.x, b = out[1];
    double c[ILP][2];
    for (int q = 0; q < ILP; q++) { c[q][0] = q; c[q][1] = 1; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int q = 0; q < ILP; q++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int q = 0; q < ILP; q++) s += c[q][0] + c[q][1];
    out[12 + threadIdx.x % 4] = s; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_dmma(double *out, long long *cyc, int n)
{
    double a = out[0] + threadIdx.x, b = out[1];
    double c0 = 0, c1 = 1;
    long long t0 = clock64();
    for (int i = 0; i < n; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    long long t1 = clock64();
    out[11] = c0 + c1; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    double *d; long long *cyc; unsigned *u;
    CK(cudaMalloc(&d, 1024)); CK(cudaMalloc(&cyc, 8 * 1024)); CK(cudaMalloc(&u, 1024));
    double h[16] = {0.999999, 1e-9, 0.5, 0}; CK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice)); CK(cudaMemset(u, 1, 1024));
    long long c; const int n = 4096;
#define RUN(name, launch, per) launch; CK(cudaDeviceSynchronize()); launch; CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost)); printf("%-34s %8.2f cycles\n", name, (double)c / (per));
    RUN("DFMA dependent latency", (lat_dfma<<<1, 32>>>(d, cyc, n)), 4.0 * n)
    RUN("rcp(1/x)+add dependent latency", (lat_rcp<<<1, 32>>>(d, cyc, n)), 1.0 * n)
    RUN("REDUX.MAX dependent latency", (lat_redux<<<1, 32>>>(u, cyc, n)), 1.0 * n)
    RUN("SHFL(64-bit)+DADD dependent", (lat_shfl<<<1, 32>>>(d, cyc, n)), 1.0 * n)
    RUN("__syncthreads 288 thr", (lat_bar<<<1, 288>>>(d, cyc, n)), 1.0 * n)
    RUN("__syncthreads 1024 thr", (lat_bar<<<1, 1024>>>(d, cyc, n)), 1.0 * n)
    RUN("LDS dependent latency", (lat_lds<<<1, 256>>>(d, cyc, n)), 1.0 * n)
    RUN("STS->bar->LDS+DADD 288 thr", (lat_sts_bar_lds<<<1, 288>>>(d, cyc, n)), 1.0 * n)
    RUN("DMMA m8n8k4 dependent latency", (lat_dmma<<<1, 32>>>(d, cyc, n)), 1.0 * n)
    // throughput: cycles per warp-instruction per SM with W warps
    for (int warps : {1, 4, 8, 16, 32}) {
        char nm[64];
        snprintf(nm, 64, "DFMA thr ILP8, %2d warps: cyc/instr", warps);
        RUN(nm, (thr_dfma<8><<<1, 32 * warps>>>(d, cyc, n)), 8.0 * n * warps)
    }
    for (int warps : {1, 4, 8, 16, 32}) {
        char nm[64];
        snprintf(nm, 64, "DMMA thr ILP8, %2d warps: cyc/instr", warps);
        RUN(nm, (thr_dmma<8><<<1, 32 * warps>>>(d, cyc, n)), 8.0 * n * warps)
    }
    // whole-chip DFMA rate
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    thr_dfma<8><<<148 * 2, 1024>>>(d, cyc, 20000); cudaDeviceSynchronize();
    cudaEventRecord(e0); thr_dfma<8><<<148 * 2, 1024>>>(d, cyc, 20000); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("whole-chip DFMA: %.2f TFLOP/s\n", 2.0 * 8 * 20000.0 * 148 * 2 * 1024 / (ms * 1e-3) / 1e12);
    thr_dmma<8><<<148 * 2, 1024>>>(d, cyc, 20000); cudaDeviceSynchronize();
    cudaEventRecord(e0); thr_dmma<8><<<148 * 2, 1024>>>(d, cyc, 20000); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    printf("whole-chip DMMA m8n8k4: %.2f TFLOP/s\n", 2.0 * 256 * 8 * 20000.0 * 148 * 2 * 32 / (ms * 1e-3) / 1e12);
    return 0;
}
