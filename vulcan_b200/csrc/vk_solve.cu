// Block-tridiagonal solver: replaces ODESolver.store_bandM + scipy.linalg.solve_banded (LAPACK dgbsv), op.py:2833-2858,
// 2914, 2929.  System per column:   dn_j x_{j-1} + D_j x_j + up_j x_{j+1} = r_j   with D_j dense ni x ni and up/dn DIAGONAL
// (the transport Jacobian freezes ysum), so the Schur complement of block Thomas is a two-sided diagonal
// scaling of the previous inverse:
//     S_0 = D_0,   S_j = D_j - diag(dn_j) W_{j-1} diag(up_{j-1}),   W_j = S_j^{-1}
//     z_j = W_j (r_j - dn_j * z_{j-1}),   x_{nz-1} = z_{nz-1},   x_j = z_j - W_j (up_j * x_{j+1})
// One thread block per column marches over the layers.  W_j is formed by in-register Gauss-Jordan elimination with
// partial (row) pivoting inside the block: the NIP x NIP matrix lives in registers in the m8n8 accumulator-fragment
// layout (lane = 4*g + t holds rows 8*tile+g, columns 8*tile+2t,2t+1), pivot row / column are broadcast through
// shared memory, the pivot search is a warp REDUX over the candidate column.  The factor is stored once (W_j) and reused
// for the second Ros2 stage and for iterative refinement - the reference factorises twice.
#include "vk_internal.cuh"

namespace vk {

__device__ __forceinline__ unsigned long long abs_bits(double v) { return (unsigned long long)__double_as_longlong(fabs(v)); }

template <int TR, int WR>
struct FactorCfg {
    static constexpr int NIP = 8 * TR * WR;
    static constexpr int NW = WR * WR;
    static constexpr int NT = NW * 32;
    static constexpr int LD = NIP + 2;
    static constexpr size_t SMEM = sizeof(double) * ((size_t)NIP * LD + 3 * NIP) + sizeof(int) * 2 * NIP;
};

struct FactorArgs {
    int nz, ni;
    const double *D;     // [ncol][nz][NIP][NIP]
    const double *up;    // [ncol][nz][NIP]
    const double *dn;
    double *W;           // [ncol][nz][NIP][NIP]
    int *status;         // [ncol]
};

template <int TR, int WR>
__global__ void __launch_bounds__(FactorCfg<TR, WR>::NT, 1) factor_kernel(FactorArgs a)
{
    using C = FactorCfg<TR, WR>;
    constexpr int NIP = C::NIP, LD = C::LD, NT = C::NT;
    extern __shared__ double smem[];
    double *Wsm = smem;                 // NIP x LD   natural-layout inverse of the previous layer
    double *colk = Wsm + NIP * LD;      // 2 x NIP    pivot column (double buffered)
    double *rowk = colk + 2 * NIP;      // NIP        pivot row
    int *piv_p = (int *)(rowk + NIP);   // p[k]: physical row chosen at step k
    int *piv_q = piv_p + NIP;           // q[r]: step at which physical row r was the pivot

    const int col = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wr = warp / WR, wc = warp % WR;
    const int ni = a.ni, nz = a.nz;
    // my rows: R(aa) = 8*(wr*TR+aa)+g ; my columns: Cc(bb,e) = 8*(wc*TR+bb)+2t+e
    const int row0 = 8 * wr * TR + g, col0 = 8 * wc * TR + 2 * t;

    double A[TR][TR][2];
    const double *Dc = a.D + (size_t)col * nz * NIP * NIP;
    const double *upc = a.up + (size_t)col * nz * NIP;
    const double *dnc = a.dn + (size_t)col * nz * NIP;
    double *Wc = a.W + (size_t)col * nz * NIP * NIP;

    for (int j = 0; j < nz; j++) {
        // ---- S_j into registers
        const double *Dj = Dc + (size_t)j * NIP * NIP;
#pragma unroll
        for (int aa = 0; aa < TR; aa++)
#pragma unroll
            for (int bb = 0; bb < TR; bb++) {
                const int r = row0 + 8 * aa, c = col0 + 8 * bb;
                double2 d = *reinterpret_cast<const double2 *>(Dj + (size_t)r * NIP + c);
                if (j > 0) {
                    const double l = dnc[(size_t)j * NIP + r];
                    const double u0 = upc[(size_t)(j - 1) * NIP + c], u1 = upc[(size_t)(j - 1) * NIP + c + 1];
                    d.x = d.x - (l * Wsm[r * LD + c]) * u0;
                    d.y = d.y - (l * Wsm[r * LD + c + 1]) * u1;
                }
                A[aa][bb][0] = d.x;
                A[aa][bb][1] = d.y;
            }
        __syncthreads();   // everyone is done reading Wsm of layer j-1
        // ---- Gauss-Jordan with implicit row pivoting
        bool used[(NIP + 31) / 32];
#pragma unroll
        for (int m = 0; m < (NIP + 31) / 32; m++) used[m] = false;
        if (tid < NIP) { piv_p[tid] = tid; piv_q[tid] = tid; }
        // publish column 0
        if (wc == 0 && t == 0) {
#pragma unroll
            for (int aa = 0; aa < TR; aa++) colk[row0 + 8 * aa] = A[aa][0][0];
        }
        bool singular = false;
        for (int k = 0; k < ni; k++) {
            const int cb = (k & 1) * NIP;
            __syncthreads();   // S1: column k published, update k-1 finished everywhere
            // pivot search (every warp redundantly): max |colk[r]| over unused rows, smallest row on ties
            unsigned long long best = 0ull;
            int brow = 0x7fffffff;
#pragma unroll
            for (int m = 0; m < (NIP + 31) / 32; m++) {
                const int r = lane + 32 * m;
                if (r < NIP && !used[m]) {
                    unsigned long long v = abs_bits(colk[cb + r]);
                    if (v > best) { best = v; brow = r; }
                }
            }
            unsigned hi = (unsigned)(best >> 32), lo = (unsigned)best;
            unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
            unsigned mlo = __reduce_max_sync(0xffffffffu, (hi == mhi) ? lo : 0u);
            unsigned prow = __reduce_min_sync(0xffffffffu, (hi == mhi && lo == mlo) ? (unsigned)brow : 0x7fffffffu);
            if ((mhi | mlo) == 0u) { singular = true; break; }
            const int p = (int)prow;
#pragma unroll
            for (int m = 0; m < (NIP + 31) / 32; m++)
                if (p == lane + 32 * m) used[m] = true;
            const double piv = colk[cb + p];
            const double inv = 1.0 / piv;
            if (tid == 0) { piv_p[k] = p; piv_q[p] = k; }
            // owners of row p publish it (with 1.0 in column k); owners of column k clear it
            const int kb = k >> 3;                  // tile column of k
            const bool own_colk = (kb / TR == wc) && (((k & 7) >> 1) == t);
            const int kbb = kb % TR, ke = k & 1;
            const int pa = (p >> 3);                // tile row of p
            const bool own_rowp = (pa / TR == wr) && ((p & 7) == g);
            const int paa = pa % TR;
            if (own_colk) {
#pragma unroll
                for (int aa = 0; aa < TR; aa++)
#pragma unroll
                    for (int bb = 0; bb < TR; bb++)
#pragma unroll
                        for (int e = 0; e < 2; e++)
                            if (bb == kbb && e == ke) A[aa][bb][e] = (own_rowp && aa == paa) ? 1.0 : 0.0;
            }
            if (own_rowp) {
#pragma unroll
                for (int aa = 0; aa < TR; aa++)
                    if (aa == paa) {
#pragma unroll
                        for (int bb = 0; bb < TR; bb++) {
                            rowk[col0 + 8 * bb] = A[aa][bb][0];
                            rowk[col0 + 8 * bb + 1] = A[aa][bb][1];
                        }
                    }
            }
            __syncthreads();   // S2: pivot row published
            double rs[TR][2], mm[TR];
#pragma unroll
            for (int bb = 0; bb < TR; bb++) {
                rs[bb][0] = rowk[col0 + 8 * bb] * inv;
                rs[bb][1] = rowk[col0 + 8 * bb + 1] * inv;
            }
#pragma unroll
            for (int aa = 0; aa < TR; aa++) mm[aa] = colk[cb + row0 + 8 * aa];
#pragma unroll
            for (int aa = 0; aa < TR; aa++) {
                const bool isp = own_rowp && (aa == paa);
#pragma unroll
                for (int bb = 0; bb < TR; bb++) {
                    A[aa][bb][0] = isp ? rs[bb][0] : fma(-mm[aa], rs[bb][0], A[aa][bb][0]);
                    A[aa][bb][1] = isp ? rs[bb][1] : fma(-mm[aa], rs[bb][1], A[aa][bb][1]);
                }
            }
            // publish column k+1 into the other buffer
            if (k + 1 < ni) {
                const int k1 = k + 1, k1b = k1 >> 3;
                if ((k1b / TR == wc) && (((k1 & 7) >> 1) == t)) {
                    const int nb = ((k1 & 1) * NIP);
                    const int bb1 = k1b % TR, e1 = k1 & 1;
#pragma unroll
                    for (int aa = 0; aa < TR; aa++)
#pragma unroll
                        for (int bb = 0; bb < TR; bb++)
#pragma unroll
                            for (int e = 0; e < 2; e++)
                                if (bb == bb1 && e == e1) colk[nb + row0 + 8 * aa] = A[aa][bb][e];
                }
            }
        }
        if (singular) {
            if (tid == 0) a.status[col] = VK_ERR_SINGULAR;
            return;   // uniform: every warp saw the same zero column
        }
        __syncthreads();   // permutation tables complete
        // ---- un-permute into shared memory:  W[q(r)][p(c)] = A[r][c]
#pragma unroll
        for (int aa = 0; aa < TR; aa++) {
            const int qr = piv_q[row0 + 8 * aa];
#pragma unroll
            for (int bb = 0; bb < TR; bb++) {
                Wsm[qr * LD + piv_p[col0 + 8 * bb]] = A[aa][bb][0];
                Wsm[qr * LD + piv_p[col0 + 8 * bb + 1]] = A[aa][bb][1];
            }
        }
        __syncthreads();
        double *Wj = Wc + (size_t)j * NIP * NIP;
        for (int q = tid; q < NIP * NIP / 2; q += NT) {
            const int r = (2 * q) / NIP, c = (2 * q) % NIP;
            double2 v = make_double2(Wsm[r * LD + c], Wsm[r * LD + c + 1]);
            *reinterpret_cast<double2 *>(Wj + (size_t)r * NIP + c) = v;
        }
        // the next iteration reads Wsm for its Schur update and syncs before overwriting it
    }
}

// ------------------------------------------------------------------------------------------------------------------
// forward + backward sweeps with the stored W_j: 4 threads per block row, 16-byte loads, shuffle reduction.
struct SolveArgs {
    int nz, ni, nip;
    const double *W, *up, *dn;   // padded layouts
    const double *rhs;           // [ncol][nz][ni]
    double *x;                   // [ncol][nz][ni]
    double *z;                   // [ncol][nz][nip] scratch
};

template <int NIP>
__global__ void __launch_bounds__(NIP * 4, 1) solve_kernel(SolveArgs a)
{
    constexpr int NCH = NIP / 8;
    __shared__ __align__(16) double tvec[NIP];
    __shared__ __align__(16) double zprev[NIP];
    const int col = blockIdx.x, tid = threadIdx.x;
    const int row = tid >> 2, part = tid & 3;
    const int nz = a.nz, ni = a.ni;
    const double *Wc = a.W + (size_t)col * nz * NIP * NIP;
    const double *upc = a.up + (size_t)col * nz * NIP;
    const double *dnc = a.dn + (size_t)col * nz * NIP;
    const double *rc = a.rhs + (size_t)col * nz * ni;
    double *xc = a.x + (size_t)col * nz * ni;
    double *zc = a.z + (size_t)col * nz * NIP;

    double2 w[NCH];
    auto load_w = [&](int j) {
        const double *Wr = Wc + (size_t)j * NIP * NIP + (size_t)row * NIP + part * 2;
#pragma unroll
        for (int i = 0; i < NCH; i++) w[i] = *reinterpret_cast<const double2 *>(Wr + i * 8);
    };
    auto matvec = [&]() -> double {
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            double2 tv = *reinterpret_cast<const double2 *>(tvec + i * 8 + part * 2);
            acc0 = fma(w[i].x, tv.x, acc0);
            acc1 = fma(w[i].y, tv.y, acc1);
        }
        double acc = acc0 + acc1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        return acc;
    };
    // ---- forward: z_j = W_j (r_j - dn_j * z_{j-1})
    if (tid < NIP) zprev[tid] = 0.0;
    load_w(0);
    for (int j = 0; j < nz; j++) {
        __syncthreads();
        if (tid < NIP) {
            double r = (tid < ni) ? rc[(size_t)j * ni + tid] : 0.0;
            tvec[tid] = (j == 0) ? r : r - dnc[(size_t)j * NIP + tid] * zprev[tid];
        }
        __syncthreads();
        double acc = matvec();
        if (j + 1 < nz) load_w(j + 1);
        if (part == 0) { zprev[row] = acc; zc[(size_t)j * NIP + row] = acc; }
    }
    // ---- backward: x_j = z_j - W_j (up_j * x_{j+1});  zprev now holds x_{j+1}
    __syncthreads();
    if (tid < ni) xc[(size_t)(nz - 1) * ni + tid] = zprev[tid];
    if (nz > 1) load_w(nz - 2);
    for (int j = nz - 2; j >= 0; j--) {
        __syncthreads();
        if (tid < NIP) tvec[tid] = upc[(size_t)j * NIP + tid] * zprev[tid];
        __syncthreads();
        double acc = matvec();
        if (j > 0) load_w(j - 1);
        if (part == 0) {
            double xv = zc[(size_t)j * NIP + row] - acc;
            zprev[row] = xv;
            if (row < ni) xc[(size_t)j * ni + row] = xv;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// residual  res = rhs - A x  (iterative refinement): one block per (column, layer), 4 threads per row
struct ResidArgs {
    int nz, ni, nip;
    const double *D, *up, *dn, *rhs, *x;
    double *res;
};
__global__ void __launch_bounds__(512) resid_kernel(ResidArgs a)
{
    extern __shared__ double xs[];   // x_j padded
    const int nz = a.nz, ni = a.ni, nip = a.nip;
    const int col = blockIdx.x / nz, j = blockIdx.x % nz;
    const int tid = threadIdx.x;
    const size_t vb = ((size_t)col * nz + j) * ni;
    for (int i = tid; i < nip; i += blockDim.x) xs[i] = (i < ni) ? a.x[vb + i] : 0.0;
    __syncthreads();
    const int row = tid >> 2, part = tid & 3;
    const bool live = row < nip;
    const double *Dr = a.D + (((size_t)col * nz + j) * nip + (live ? row : 0)) * nip;
    double acc = 0.0;
    for (int c = part * 2; live && c < nip; c += 8) {
        double2 d = *reinterpret_cast<const double2 *>(Dr + c);
        acc = fma(d.x, xs[c], acc);
        acc = fma(d.y, xs[c + 1], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0 && row < ni) {
        const size_t pb = ((size_t)col * nz + j) * nip;
        if (j + 1 < nz) acc = fma(a.up[pb + row], a.x[vb + ni + row], acc);
        if (j > 0) acc = fma(a.dn[pb + row], a.x[vb - ni + row], acc);
        a.res[vb + row] = a.rhs[vb + row] - acc;
    }
}

// ------------------------------------------------------------------------------------------------------------------
template <int TR, int WR>
static int launch_factor_t(vk_column *c, const double *D, const double *up, const double *dn, double *W, int *status)
{
    using C = FactorCfg<TR, WR>;
    static bool configured = false;
    if (!configured) {
        VK_CUDA(cudaFuncSetAttribute(factor_kernel<TR, WR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured = true;
    }
    FactorArgs a{c->nz, c->ni, D, up, dn, W, status};
    factor_kernel<TR, WR><<<c->ncol, C::NT, C::SMEM, c->stream>>>(a);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

int launch_factor(vk_column *c, const double *D, const double *up, const double *dn, double *W, int *status)
{
    switch (c->nip) {
        case 48: return launch_factor_t<2, 3>(c, D, up, dn, W, status);
        case 72: return launch_factor_t<3, 3>(c, D, up, dn, W, status);
        case 96: return launch_factor_t<4, 3>(c, D, up, dn, W, status);
        case 120: return launch_factor_t<5, 3>(c, D, up, dn, W, status);
        default: set_error("no factor kernel for this padded block size"); return VK_ERR_UNSUPPORTED;
    }
}

int launch_solve(vk_column *c, const double *W, const double *up, const double *dn, const double *rhs, double *x, double *z)
{
    SolveArgs a{c->nz, c->ni, c->nip, W, up, dn, rhs, x, z};
    switch (c->nip) {
        case 48: solve_kernel<48><<<c->ncol, 48 * 4, 0, c->stream>>>(a); break;
        case 72: solve_kernel<72><<<c->ncol, 72 * 4, 0, c->stream>>>(a); break;
        case 96: solve_kernel<96><<<c->ncol, 96 * 4, 0, c->stream>>>(a); break;
        case 120: solve_kernel<120><<<c->ncol, 120 * 4, 0, c->stream>>>(a); break;
        default: set_error("no solve kernel for this padded block size"); return VK_ERR_UNSUPPORTED;
    }
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

int launch_residual(vk_column *c, const double *D, const double *up, const double *dn, const double *rhs, const double *x, double *res)
{
    ResidArgs a{c->nz, c->ni, c->nip, D, up, dn, rhs, x, res};
    resid_kernel<<<c->ncol * c->nz, 512, sizeof(double) * c->nip, c->stream>>>(a);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

}  // namespace vk
