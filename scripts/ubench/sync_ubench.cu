// producer/consumer hand-off latency between two warps of one block on B200: named barriers vs shared-memory flag polling
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ID, int COUNT> __device__ __forceinline__ void bsync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT> __device__ __forceinline__ void barrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }

// warp 0 <-> warp 1 ping-pong with named barriers; extra (nw-2) warps just sync on the "pong" barrier like the matrix warps do
template <int NWTOT>
__global__ void pingpong_bar(long long *cyc, int n)
{
    const int w = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (w == 0) { barrive<1, 64>(); bsync<2, NWTOT * 32>(); }
        else if (w == 1) { bsync<1, 64>(); barrive<2, NWTOT * 32>(); }
        else { bsync<2, NWTOT * 32>(); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void pingpong_flag(long long *cyc, int n)
{
    __shared__ volatile int f0, f1;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { f0 = 0; f1 = 0; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 1; i <= n; i++) {
        if (w == 0) {
            if (lane == 0) f0 = i;
            while (f1 != i) { }
        } else if (w == 1) {
            while (f0 != i) { }
            if (lane == 0) f1 = i;
        } else {
            while (f1 != i) { }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// dependent chain of "typical" integer ops in ONE warp: how many cycles per dependent ALU instruction?
__global__ void alu_chain(unsigned *out, long long *cyc, int n)
{
    unsigned v = out[threadIdx.x & 31];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        v = (v & 0x7fffff80u) | 5u; v = max(v, 17u); v = v + 3u; v = (v >> 1) ^ 0x55u;
    }
    long long t1 = clock64();
    out[32] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void alu_indep(unsigned *out, long long *cyc, int n)
{
    unsigned v[8];
    for (int q = 0; q < 8; q++) v[q] = out[(threadIdx.x + q) & 31];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) { v[q] = (v[q] & 0x7fffff80u) | 5u; v[q] = max(v[q], 17u + q); }
    }
    long long t1 = clock64();
    unsigned s = 0; for (int q = 0; q < 8; q++) s += v[q];
    out[32] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    long long *cyc, c; unsigned *u;
    CK(cudaMalloc(&cyc, 64)); CK(cudaMalloc(&u, 1024)); CK(cudaMemset(u, 1, 1024));
    const int n = 4096;
#define RUN(name, launch, per) launch; CK(cudaDeviceSynchronize()); launch; CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost)); printf("%-52s %8.2f cycles\n", name, (double)c / (per));
    RUN("named-barrier ping-pong round trip (2 warps)", (pingpong_bar<2><<<1, 64>>>(cyc, n)), 1.0 * n)
    RUN("named-barrier ping-pong round trip (10 warps block)", (pingpong_bar<10><<<1, 320>>>(cyc, n)), 1.0 * n)
    RUN("smem flag ping-pong round trip (2 warps)", (pingpong_flag<<<1, 64>>>(cyc, n)), 1.0 * n)
    RUN("smem flag ping-pong round trip (10 warps polling)", (pingpong_flag<<<1, 320>>>(cyc, n)), 1.0 * n)
    RUN("dependent ALU op (1 warp): cycles per instr", (alu_chain<<<1, 32>>>(u, cyc, n)), 5.0 * n)
    RUN("independent ALU ops ILP8 (1 warp): cycles per instr", (alu_indep<<<1, 32>>>(u, cyc, n)), 24.0 * n)
    return 0;
}
