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
    const double *k;        // [ncol|1][nz][nr+1]: shared by the batch (k_cs = 0) or per column with identical "static" rows (see KG)
    size_t k_cs;
    const double *M;        // atm.M [ncol|1][nz]
    size_t M_cs;
    double *chem;           // out [ncol][nz][ni]
    double *ysum;           // out [ncol][nz]: np.sum(y, axis=1) / np.sum(y[:, gas_indx], axis=1) (op.py:1505-1507)
    int n_gas; const int *gas_indx;
    const int *act;         // [ncol] or NULL
};
// the chemical Jacobian -J (chem_funs.neg_symjac) of one layer of 128 columns, dense rows into D; diagonal / transport terms are added by
// lhs_diag_kernel (vk_chem.cu) afterwards
struct EmitJacArgs {
    int nz, ncol;
    const double *y;        // [ncol][nz][ni]
    const double *k;        // [ncol|1][nz][nr+1]
    size_t k_cs;
    const double *M;        // atm.M [ncol|1][nz]
    size_t M_cs;
    double *D;              // [ncol][nz][NIP][NIP]
    double *ysum;           // out [ncol][nz]: the layer sums lhs_jac_tot reads (op.py:1981-1984)
    int n_gas; const int *gas_indx;
    const int *act;
};
typedef int (*EmitLaunch)(const EmitArgs &, cudaStream_t);
typedef int (*EmitJacLaunch)(const EmitJacArgs &, cudaStream_t);
// dyn / n_dyn: the k indices a run rewrites per column (photolysis, ionisation, condensation rows): the emitted code reads those through
// KG(i) from the thread's own column, everything else through K(i) from the block's shared copy of the row
struct EmitEntry { unsigned long long hash; int ni, nr; const char *name; EmitLaunch fn; EmitJacLaunch jac; const int *dyn; int n_dyn; };
// registry of the kernels compiled into this library (vk_emit.cu); looked up by the hash of the uploaded tables in vk_network_create
void emit_register(const EmitEntry &e);
const EmitEntry *emit_find(unsigned long long hash, int ni, int nr);
struct EmitRegistrar {
    EmitRegistrar(unsigned long long hash, int ni, int nr, const char *name, EmitLaunch fn, EmitJacLaunch jac, const int *dyn, int n_dyn)
    {
        emit_register(EmitEntry{hash, ni, nr, name, fn, jac, dyn, n_dyn});
    }
};
int emit_set_smem(const void *func, size_t bytes);      // per (function, device) opt-in above 48 KB

inline size_t emit_smem_bytes(int ni, int nr)
{
    return sizeof(double) * ((size_t)(ni + 1) * VK_EMIT_LD + (size_t)nr + 2);
}

inline size_t emitj_smem_bytes(int rld, int nr, int tb)
{
    return sizeof(double) * ((size_t)tb * rld + (size_t)nr + 2);
}
#ifndef VK_EMITJ_ROW_SYNC
#define VK_EMITJ_ROW_SYNC
#endif

// np.sum over one layer of one column, read from the transposed tile (stride VK_EMIT_LD): numpy's pairwise association for n <= 128
// (8 accumulators r_m = a[m] + a[8+m] + .., ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail), or the plain left-to-right sum over
// gas_indx that np.sum(y[:, gas_indx], axis=1) performs (vk_device_math.cuh: row_sum)
template <int STRIDE>
__device__ __forceinline__ double emit_row_sum(const double *yT, int n, int n_gas, const int *gas)
{
    if (n_gas > 0) {
        double acc = yT[gas[0] * STRIDE];
        for (int i = 1; i < n_gas; i++) acc += yT[gas[i] * STRIDE];
        return acc;
    }
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res += yT[i * STRIDE];
        return res;
    }
    double r[8];
#pragma unroll
    for (int m = 0; m < 8; m++) r[m] = yT[m * STRIDE];
    const int nb = n - (n % 8);
    for (int i = 8; i < nb; i += 8) {
#pragma unroll
        for (int m = 0; m < 8; m++) r[m] += yT[(i + m) * STRIDE];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (int i = nb; i < n; i++) res += yT[i * STRIDE];
    return res;
}

#define VK_EMIT_PROLOGUE(NI, NR, FIRST)                                                                                   \
    extern __shared__ __align__(16) double sm[];                                                                         \
    const int tid = threadIdx.x, wrp = tid >> 5, lane = tid & 31;                                                        \
    double *ys = sm;                                                                                                     \
    double *ks = ys + ((NI) + 1) * VK_EMIT_LD;                                                                           \
    const int j = (int)(blockIdx.x % (unsigned)a.nz), col0 = (int)(blockIdx.x / (unsigned)a.nz) * VK_EMIT_TB;            \
    const int ncb = min(VK_EMIT_TB, a.ncol - col0);                                                                      \
    if (a.act && !__syncthreads_or(tid < ncb && a.act[col0 + tid])) return;   /* every column of the block has stopped */  \
    for (int i = tid; i <= (NR); i += VK_EMIT_TB) ks[i] = a.k[(size_t)col0 * a.k_cs + (size_t)j * ((NR) + 1) + i];       \
    /* rows the run rewrites per column (KG): the thread's own column; with one shared k they are in the block's copy too */ \
    const double *const kg = a.k_cs ? a.k + (size_t)(col0 + (tid < ncb ? tid : 0)) * a.k_cs + (size_t)j * ((NR) + 1) : ks; \
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
    if ((FIRST) && a.ysum && tid < ncb) a.ysum[(size_t)(col0 + tid) * a.nz + j] = emit_row_sum<VK_EMIT_LD>(yT, (NI), a.n_gas, a.gas_indx);

#define Y(s) yT[(s) * VK_EMIT_LD]
// volatile: every use re-reads the broadcast from shared memory.  Without it the compiler keeps k values it will need again in registers and,
// out of registers, SPILLS them to local memory (2 KB per thread in the Jacobian kernel) - a shared-memory load is cheaper than either
#define K(i) (*(const volatile double *)(ks + (i)))
#define KG(i) kg[(i)]
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

// ---- Jacobian kernel: rb [TB columns][RLD] - thread tid owns row tid: first its y (staged coalesced, then loaded into registers), then
// the buffer of the Jacobian row being formed, which leaves as ONE asynchronous bulk copy (576 contiguous bytes of D for NIP = 72) issued
// by the thread itself: no block barrier in the row loop, the next row is formed while the copy drains.  Behind the row buffer (from NIP + 2)
// the thread keeps the y of the species the Jacobian uses least ("cold": ~10 % of all factor reads) - the others live in registers.
// RLD is even (16-byte aligned rows for the bulk copy): 2-way bank conflict on the thread's own accesses.
// Measured alternatives, 4096 HD189 columns: the warp storing its 32 rows with 16-byte vector stores 9.0 ms (instruction overhead);
// every entry stored straight into D with the rest of the buffer zeroed once 17.5 ms (partial-sector writes); this variant 7.2 ms.
#define VK_EMITJ_PROLOGUE(NI, NR, NIP, RLD, TB)                                                                           \
    extern __shared__ __align__(128) double sm[];                                                                        \
    const int tid = threadIdx.x, wrp = tid >> 5, lane = tid & 31;                                                        \
    double *rb = sm;                                                                                                     \
    double *ks = rb + (TB) * (RLD);                                                                                      \
    const int j = (int)(blockIdx.x % (unsigned)a.nz), col0 = (int)(blockIdx.x / (unsigned)a.nz) * (TB);                  \
    const int ncb = min((TB), a.ncol - col0);                                                                            \
    if (a.act && !__syncthreads_or(tid < ncb && a.act[col0 + tid])) return;   /* every column of the block has stopped */  \
    for (int i = tid; i <= (NR); i += (TB)) ks[i] = a.k[(size_t)col0 * a.k_cs + (size_t)j * ((NR) + 1) + i];             \
    /* rows the run rewrites per column (KG): the thread's own column; with one shared k they are in the block's copy too */ \
    const double *const kg = a.k_cs ? a.k + (size_t)(col0 + (tid < ncb ? tid : 0)) * a.k_cs + (size_t)j * ((NR) + 1) : ks; \
    for (int cc = wrp; cc < (TB); cc += (TB) / 32) {                                                                     \
        const size_t base = ((size_t)(col0 + cc) * a.nz + j) * (NI);                                                     \
        for (int s = lane; s < (NI); s += 32) rb[cc * (RLD) + s] = (cc < ncb) ? a.y[base + s] : 0.0;                     \
    }                                                                                                                    \
    __syncthreads();                                                                                                     \
    double *const rT = rb + tid * (RLD);                                                                                 \
    const bool live = tid < ncb && (!a.act || a.act[col0 + tid]);                                                        \
    const double yM = (tid < ncb) ? a.M[(size_t)(col0 + tid) * a.M_cs + j] : 0.0;                                        \
    if (a.ysum && tid < ncb) a.ysum[(size_t)(col0 + tid) * a.nz + j] = emit_row_sum<1>(rT, (NI), a.n_gas, a.gas_indx);   \
    double *const Dblk = a.D + ((size_t)(col0 + (tid < ncb ? tid : 0)) * a.nz + j) * (NIP) * (NIP);                      \
    const unsigned rT_s = (unsigned)__cvta_generic_to_shared(rT);

#define S(s) rT[(s)]
#define R(t) rT[(t)]

// (the thread has copied its y into registers / its cold slots: its row of rb becomes the row buffer)
#define VK_EMITJ_BEGIN(NI, NIP)                                                                                           \
    for (int r = 0; r < (NIP); r++) rT[r] = 0.0;

// the row buffer is about to be rewritten: the bulk copy of the previous row must have read it (placed by the generator AFTER the first
// group of entries of the row has been accumulated in registers, so that the wait overlaps that arithmetic)
#define VK_EMITJ_ROW_OPEN(s)                                                                                              \
    if ((s) > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");                                          \
    VK_EMITJ_ROW_SYNC

#define VK_EMITJ_ROW(s, NI, NIP)                                                                                          \
    if (live) {                                                                                                          \
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                                     \
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"                                       \
                     ::"l"(Dblk + (size_t)(s) * (NIP)), "r"(rT_s), "r"((unsigned)(sizeof(double) * (NIP))) : "memory");  \
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");                                                        \
    }

#define VK_EMITJ_END()                                                                                                    \
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

#define VK_EMITJ_LAUNCH(kern, TB)                                                                                         \
    {                                                                                                                    \
        if (emit_set_smem((const void *)kern, smem)) return 1;                                                           \
        kern<<<(unsigned)(a.nz * ((a.ncol + (TB) - 1) / (TB))), (TB), smem, st>>>(a);                                    \
        if (cudaGetLastError() != cudaSuccess) return 1;                                                                 \
    }

#define VK_EMIT_LAUNCH(kern)                                                                                              \
    {                                                                                                                    \
        if (emit_set_smem((const void *)kern, smem)) return 1;                                                           \
        kern<<<(unsigned)(a.nz * ((a.ncol + VK_EMIT_TB - 1) / VK_EMIT_TB)), VK_EMIT_TB, smem, st>>>(a);                  \
        if (cudaGetLastError() != cudaSuccess) return 1;                                                                 \
    }

}}  // namespace vk::emitted
