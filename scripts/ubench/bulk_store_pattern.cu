// Microbenchmark: is the emitted Jacobian kernel (negjac_*) bound by its store pattern?  Same grid, same addresses, no
// arithmetic: every thread (column) of a block (layer j, 128 columns) sends NROW pieces of PIECE doubles from its shared-memory row to
//   D[col][j][r * PIECE ...]   (D = [ncol][nz][72][72] doubles)
// by cp.async.bulk, waiting for the previous piece to be read before "rewriting" the buffer.  PIECE = 72: one block row per copy (what the
// kernel does); 144: two rows per copy; 5184: the whole 41 KB block in one copy (upper bound of the layout).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a bulk_store_pattern.cu -o /tmp/bsp && /tmp/bsp
#include <cstdio>
#include <cuda_runtime.h>

template <int PIECE, int TB>
__global__ void __launch_bounds__(TB) store_kernel(double *D, int nz, int ncol)
{
    extern __shared__ __align__(128) double sm[];
    const int tid = threadIdx.x;
    const int j = blockIdx.x % nz, col = (blockIdx.x / nz) * TB + tid;
    constexpr int BUF = PIECE < 1024 ? PIECE + 2 : 72 + 2;            // the whole-block variant sends the same 592 bytes 72 times over
    double *row = sm + (size_t)tid * BUF;
    for (int t = 0; t < BUF; t++) row[t] = (double)(tid + t);
    if (col >= ncol) return;
    double *blk = D + ((size_t)col * nz + j) * 5184;
    const unsigned rs = (unsigned)__cvta_generic_to_shared(row);
    constexpr int NP = 5184 / (PIECE < 1024 ? PIECE : 72);
    constexpr int BYTES = (PIECE < 1024 ? PIECE : 72) * 8;
    for (int p = 0; p < NP; p++) {
        if (p > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        row[p % 8] = (double)p;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        double *dst = blk + (size_t)p * (BYTES / 8);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(rs), "r"((unsigned)BYTES) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// plain coalesced streaming store of the same volume: the write ceiling of the device
__global__ void fill_kernel(double2 *D, size_t n2)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) D[i] = make_double2(1.0, 2.0);
}

template <class F>
static float time_ms(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 3; r++) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
}

int main()
{
    const int ncol = 4096, nz = 150;
    const size_t n = (size_t)ncol * nz * 5184;
    double *D; cudaMalloc(&D, n * 8);
    const double gb = n * 8 / 1e9;
    float ms = time_ms([&] { fill_kernel<<<148 * 16, 256>>>((double2 *)D, n / 2); });
    printf("coalesced fill              %7.3f ms  %6.0f GB/s\n", ms, gb / ms * 1e3);
    {
        auto k = store_kernel<72, 128>; size_t smem = 128 * 74 * 8;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ms = time_ms([&] { k<<<nz * (ncol / 128), 128, smem>>>(D, nz, ncol); });
        printf("bulk 576 B per thread-row   %7.3f ms  %6.0f GB/s\n", ms, gb / ms * 1e3);
    }
    {
        auto k = store_kernel<144, 128>; size_t smem = 128 * 146 * 8;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ms = time_ms([&] { k<<<nz * (ncol / 128), 128, smem>>>(D, nz, ncol); });
        printf("bulk 1152 B (two rows)      %7.3f ms  %6.0f GB/s\n", ms, gb / ms * 1e3);
    }
    {
        auto k = store_kernel<72, 256>; size_t smem = 256 * 74 * 8;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ms = time_ms([&] { k<<<nz * (ncol / 256), 256, smem>>>(D, nz, ncol); });
        printf("bulk 576 B, 256-thr blocks  %7.3f ms  %6.0f GB/s\n", ms, gb / ms * 1e3);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
