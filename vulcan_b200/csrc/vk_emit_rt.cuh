// Runtime pieces of the EMITTED chemistry kernels (vulcan_b200/emit.py writes csrc/gen/chemdf_<network>.cu, one straight-line kernel per
// network: chem_funs.chemdf, make_chem_funs.py:113-430).  One THREAD per (column, layer); a block of VK_EMIT_TB threads takes ONE layer j of
// VK_EMIT_TB consecutive columns of a batch that shares its rate coefficients (one T-P profile, vk_set_k(shared = 1)):
//   ks [nr+1]      the k row of layer j, read by every thread at the same address (shared-memory broadcast, immediate offsets)
//   ys [ni+1][LD]  y of the block's columns at layer j, transposed (row ni = the third body M): thread tid reads ys[s][tid], conflict-free
// and every thread keeps dy_s/dt of all species of the pass in registers - no barrier between the prologue and the store, so the compiler
// schedules the whole reaction list as one basic block.  Also written: the layer sums of y (ysum, numpy's association) and, in stage 2,
// y + k1/r - both are inputs of the transport stencil kernel that follows (vk_chem.cu: rhs_stencil_kernel).
// -fmad=false like every other chemistry unit: the sums and products round exactly as numpy's do.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace vk { namespace emitted {

#define VK_EMIT_TB 128
#define VK_EMIT_LD (VK_EMIT_TB + 1)

struct EmitArgs {
    int nz, ncol;
    const double *y;        // [ncol][nz][ni]
    const double *k1;       // stage 2: chemdf(y + k1/r) (op.py:2917); NULL for stage 1
    double *yk2_out;        // stage 2: y + k1/r
    const double *k;        // [nz][nr+1], shared by the batch
    const double *M;        // atm.M [ncol|1][nz]
    size_t M_cs;
    double *chem;           // out [ncol][nz][ni]
    double *ysum;           // out [ncol][nz]: np.sum(y, axis=1) / np.sum(y[:, gas_indx], axis=1) (op.py:1505-1507)
    int n_gas; const int *gas_indx;
    const int *act;         // [ncol] or NULL
};
typedef int (*EmitLaunch)(const EmitArgs &, cudaStream_t);
struct EmitEntry { unsigned long long hash; int ni, nr; const char *name; EmitLaunch fn; };
// registry of the kernels compiled into this library (vk_emit.cu); looked up by the hash of the uploaded tables in vk_network_create
void emit_register(const EmitEntry &e);
const EmitEntry *emit_find(unsigned long long hash, int ni, int nr);
struct EmitRegistrar {
    EmitRegistrar(unsigned long long hash, int ni, int nr, const char *name, EmitLaunch fn) { emit_register(EmitEntry{hash, ni, nr, name, fn}); }
};
int emit_set_smem(const void *func, size_t bytes);      // per (function, device) opt-in above 48 KB

inline size_t emit_smem_bytes(int ni, int nr)
{
    return sizeof(double) * ((size_t)(ni + 1) * VK_EMIT_LD + (size_t)nr + 2);
}

// np.sum over one layer of one column, read from the transposed tile (stride VK_EMIT_LD): numpy's pairwise association for n <= 128
// (8 accumulators r_m = a[m] + a[8+m] + .., ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail), or the plain left-to-right sum over
// gas_indx that np.sum(y[:, gas_indx], axis=1) performs (vk_device_math.cuh: row_sum)
__device__ __forceinline__ double emit_row_sum(const double *yT, int n, int n_gas, const int *gas)
{
    if (n_gas > 0) {
        double acc = yT[gas[0] * VK_EMIT_LD];
        for (int i = 1; i < n_gas; i++) acc += yT[gas[i] * VK_EMIT_LD];
        return acc;
    }
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res += yT[i * VK_EMIT_LD];
        return res;
    }
    double r[8];
#pragma unroll
    for (int m = 0; m < 8; m++) r[m] = yT[m * VK_EMIT_LD];
    const int nb = n - (n % 8);
    for (int i = 8; i < nb; i += 8) {
#pragma unroll
        for (int m = 0; m < 8; m++) r[m] += yT[(i + m) * VK_EMIT_LD];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (int i = nb; i < n; i++) res += yT[i * VK_EMIT_LD];
    return res;
}

#define VK_EMIT_PROLOGUE(NI, NR, FIRST)                                                                                   \
    extern __shared__ __align__(16) double sm[];                                                                         \
    const int tid = threadIdx.x, wrp = tid >> 5, lane = tid & 31;                                                        \
    double *ys = sm;                                                                                                     \
    double *ks = ys + ((NI) + 1) * VK_EMIT_LD;                                                                           \
    const int j = (int)(blockIdx.x % (unsigned)a.nz), col0 = (int)(blockIdx.x / (unsigned)a.nz) * VK_EMIT_TB;            \
    const int ncb = min(VK_EMIT_TB, a.ncol - col0);                                                                      \
    for (int i = tid; i <= (NR); i += VK_EMIT_TB) ks[i] = a.k[(size_t)j * ((NR) + 1) + i];                               \
    {                                                                                                                    \
        const double rr = 1. + 1. / sqrt(2.);                                                                            \
        for (int cc = wrp; cc < VK_EMIT_TB; cc += VK_EMIT_TB / 32) {                                                     \
            const size_t base = ((size_t)(col0 + cc) * a.nz + j) * (NI);                                                 \
            for (int s = lane; s < (NI); s += 32) {                                                                      \
                double v = 0.0;                                                                                          \
                if (cc < ncb) {                                                                                          \
                    v = a.y[base + s];                                                                                   \
                    if (a.k1) {                                                                                          \
                        v = v + a.k1[base + s] / rr;                                                                     \
                        if ((FIRST) && a.yk2_out) a.yk2_out[base + s] = v;                                               \
                    }                                                                                                    \
                }                                                                                                        \
                ys[s * VK_EMIT_LD + cc] = v;                                                                             \
            }                                                                                                            \
        }                                                                                                                \
        ys[(NI) * VK_EMIT_LD + tid] = (tid < ncb) ? a.M[(size_t)(col0 + tid) * a.M_cs + j] : 0.0;                        \
    }                                                                                                                    \
    __syncthreads();                                                                                                     \
    const double *const yT = ys + tid;                                                                                   \
    if ((FIRST) && a.ysum && tid < ncb) a.ysum[(size_t)(col0 + tid) * a.nz + j] = emit_row_sum(yT, (NI), a.n_gas, a.gas_indx);

#define Y(s) yT[(s) * VK_EMIT_LD]
#define K(i) ks[(i)]
#define F(s) fT[(s) * VK_EMIT_LD]

#define VK_EMIT_STORE_BEGIN(NI)                                                                                           \
    __syncthreads();                                                                                                     \
    double *const fT = ys + tid;

#define VK_EMIT_STORE_END(NI, S0, S1)                                                                                     \
    __syncthreads();                                                                                                     \
    for (int cc = wrp; cc < ncb; cc += VK_EMIT_TB / 32) {                                                                \
        if (a.act && !a.act[col0 + cc]) continue;                                                                        \
        const size_t base = ((size_t)(col0 + cc) * a.nz + j) * (NI);                                                     \
        for (int s = (S0) + lane; s < (S1); s += 32) a.chem[base + s] = ys[s * VK_EMIT_LD + cc];                         \
    }

#define VK_EMIT_LAUNCH(kern)                                                                                              \
    {                                                                                                                    \
        if (emit_set_smem((const void *)kern, smem)) return 1;                                                           \
        kern<<<(unsigned)(a.nz * ((a.ncol + VK_EMIT_TB - 1) / VK_EMIT_TB)), VK_EMIT_TB, smem, st>>>(a);                  \
        if (cudaGetLastError() != cudaSuccess) return 1;                                                                 \
    }

}}  // namespace vk::emitted
