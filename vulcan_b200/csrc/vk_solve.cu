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

// Register layout: warp w owns the 8 columns [8w, 8w+8) of the block and ALL its rows, in the m8n8 accumulator-fragment
// pattern stacked vertically: lane = 4*g + t holds rows 8*i + g (i < NIP/8) and columns 8w + 2t, 8w + 2t + 1.
// One pivot step k:
//   owner warp (k/8): pivot search on its own column with ONE warp REDUX over packed keys {|a| high word, row} (the
//   maximum is taken on the top 13 mantissa bits: threshold partial pivoting with threshold 1 - 2^-13), reciprocal of
//   each lane's local best computed while the REDUX is in flight, multipliers l_r = a_rk / pivot published to shared memory;
//   __syncthreads (the only block barrier of the step);
//   every warp: rank-1 update of its own 8 columns; the pivot-row values it needs are its own (warp shuffle).
// dst = A[idx][e] with a block-uniform runtime idx: a uniform switch keeps A in registers (no local-memory indexing)
#define VK_SELECT_ROW(idx, e, dst)                                                     \
    switch (idx) {                                                                     \
        case 0: dst = A[0][e]; break;                                                  \
        case 1: if (NR > 1) dst = A[1 < NR ? 1 : 0][e]; break;                         \
        case 2: if (NR > 2) dst = A[2 < NR ? 2 : 0][e]; break;                         \
        case 3: if (NR > 3) dst = A[3 < NR ? 3 : 0][e]; break;                         \
        case 4: if (NR > 4) dst = A[4 < NR ? 4 : 0][e]; break;                         \
        case 5: if (NR > 5) dst = A[5 < NR ? 5 : 0][e]; break;                         \
        case 6: if (NR > 6) dst = A[6 < NR ? 6 : 0][e]; break;                         \
        case 7: if (NR > 7) dst = A[7 < NR ? 7 : 0][e]; break;                         \
        case 8: if (NR > 8) dst = A[8 < NR ? 8 : 0][e]; break;                         \
        case 9: if (NR > 9) dst = A[9 < NR ? 9 : 0][e]; break;                         \
        case 10: if (NR > 10) dst = A[10 < NR ? 10 : 0][e]; break;                     \
        case 11: if (NR > 11) dst = A[11 < NR ? 11 : 0][e]; break;                     \
        case 12: if (NR > 12) dst = A[12 < NR ? 12 : 0][e]; break;                     \
        case 13: if (NR > 13) dst = A[13 < NR ? 13 : 0][e]; break;                     \
        case 14: if (NR > 14) dst = A[14 < NR ? 14 : 0][e]; break;                     \
        default: break;                                                                \
    }

template <int NIP>
struct FactorCfg {
    static constexpr int NW = NIP / 8;
    static constexpr int NR = NIP / 8;
    static constexpr int NT = NW * 32;
    static constexpr int LD = NIP + 2;
    // Wsm + lbuf[2] + tvec + zprev (doubles), pinv[2] (double), piv_p, piv_q, pidx[2] (ints)
    static constexpr size_t SMEM = sizeof(double) * ((size_t)NIP * LD + 5 * NIP) + sizeof(int) * (2 * NIP + 4);
};

struct FactorArgs {
    int nz, ni;
    const double *D;     // [ncol][nz][NIP][NIP]
    const double *up;    // [ncol][nz][NIP]
    const double *dn;
    double *W;           // [ncol][nz][NIP][NIP]
    int *status;         // [ncol]
    const double *rhs;   // optional [ncol][nz][ni]: forward elimination fused into the factorisation
    double *z;           // [ncol][nz][NIP]
};

template <int NIP, int MINB>
__global__ void __launch_bounds__(FactorCfg<NIP>::NT, MINB) factor_kernel(FactorArgs a)
{
    using C = FactorCfg<NIP>;
    constexpr int NR = C::NR, LD = C::LD, NT = C::NT;
    extern __shared__ __align__(16) double smem[];
    double *Wsm = smem;                  // NIP x LD   natural-layout inverse of the current / previous layer
    double *lbuf = Wsm + NIP * LD;       // 2 x NIP    multipliers of step k (double buffered)
    double *tvec = lbuf + 2 * NIP;       // NIP
    double *zprev = tvec + NIP;          // NIP
    double *rscale = zprev + NIP;        // NIP  1/pivot of each physical row (deferred row scaling)
    int *piv_p = (int *)(rscale + NIP);  // p[k]: physical row chosen at step k
    int *piv_q = piv_p + NIP;            // q[r]: step at which physical row r was the pivot
    int *pidx = piv_q + NIP;             // 2

    const int col = blockIdx.x;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int ni = a.ni, nz = a.nz;
    const int c0 = 8 * w + 2 * t;        // my columns c0, c0+1 ; my rows 8*i + g

    double A[NR][2];
    const double *Dc = a.D + (size_t)col * nz * NIP * NIP;
    const double *upc = a.up + (size_t)col * nz * NIP;
    const double *dnc = a.dn + (size_t)col * nz * NIP;
    double *Wc = a.W + (size_t)col * nz * NIP * NIP;
    const bool fuse = a.rhs != nullptr;
    const double *rc = fuse ? a.rhs + (size_t)col * nz * ni : nullptr;
    double *zc = fuse ? a.z + (size_t)col * nz * NIP : nullptr;
    if (tid < NIP) zprev[tid] = 0.0;

    for (int j = 0; j < nz; j++) {
        // ---- S_j = D_j - diag(dn_j) W_{j-1} diag(up_{j-1}) into registers
        const double *Dj = Dc + (size_t)j * NIP * NIP;
        {
            double u0 = 0.0, u1 = 0.0;
            if (j > 0) { u0 = upc[(size_t)(j - 1) * NIP + c0]; u1 = upc[(size_t)(j - 1) * NIP + c0 + 1]; }
#pragma unroll
            for (int i = 0; i < NR; i++) {
                const int r = 8 * i + g;
                double2 d = *reinterpret_cast<const double2 *>(Dj + (size_t)r * NIP + c0);
                if (j > 0) {
                    const double l = dnc[(size_t)j * NIP + r];
                    const double2 wv = *reinterpret_cast<const double2 *>(Wsm + r * LD + c0);
                    d.x = d.x - (l * wv.x) * u0;
                    d.y = d.y - (l * wv.y) * u1;
                }
                A[i][0] = d.x;
                A[i][1] = d.y;
            }
        }
        if (tid < NIP) { piv_p[tid] = tid; piv_q[tid] = tid; rscale[tid] = 1.0; }
        __syncthreads();   // everyone is done reading Wsm of layer j-1
        unsigned avail = (NR >= 32) ? 0xffffffffu : ((1u << NR) - 1u);   // bit i: my row 8i+g has not been a pivot yet
        bool singular = false;
        for (int k = 0; k < ni; k++) {
            const int cb = k & 1;
            const int kc = k & 7;
            if (w == (k >> 3)) {
                // ---- owner warp: pivot search on column k (held by the 8 lanes with t == kc/2)
                const bool holder = (t == (kc >> 1));
                unsigned key[NR];
#pragma unroll
                for (int i = 0; i < NR; i++) {
                    const int hi = (kc & 1) ? __double2hiint(A[i][1]) : __double2hiint(A[i][0]);
                    key[i] = (((unsigned)hi & 0x7fffff80u) | (unsigned)(127 - (8 * i + g))) & (0u - ((avail >> i) & 1u));
                }
#pragma unroll
                for (int st = 1; st < NR; st <<= 1)
#pragma unroll
                    for (int i = 0; i + st < NR; i += 2 * st) key[i] = max(key[i], key[i + st]);
                const unsigned mk = __reduce_max_sync(0xffffffffu, holder ? key[0] : 0u);
                const int p = 127 - (int)(mk & 0x7fu);
                const int pi = p >> 3;
                double pv0 = 1.0, pv1 = 1.0;
                VK_SELECT_ROW(pi, 0, pv0);
                VK_SELECT_ROW(pi, 1, pv1);
                double pv = (kc & 1) ? pv1 : pv0;
                pv = __shfl_sync(0xffffffffu, pv, ((p & 7) << 2) | (kc >> 1));
                const double inv = 1.0 / pv;
                if (holder) {
#pragma unroll
                    for (int i = 0; i < NR; i++) {
                        const int r = 8 * i + g;
                        const double v = (kc & 1) ? A[i][1] : A[i][0];
                        const double l = (r == p) ? 0.0 : v * inv;
                        lbuf[cb * NIP + r] = l;
                        // column k of the transformed block; the pivot row is kept UNSCALED (its factor 1/pivot is applied
                        // once, when the block is written out), so its own entry in column k is pivot * (1/pivot) = 1
                        const double nv = (r == p) ? 1.0 : -l;
                        if (kc & 1) A[i][1] = nv; else A[i][0] = nv;
                    }
                }
                if (lane == 0) {
                    const bool ok = (mk >> 7) != 0u;
                    pidx[cb] = ok ? p : -1;
                    if (ok) { piv_p[k] = p; piv_q[p] = k; rscale[p] = inv; }
                }
            }
            __syncthreads();
            const int p = pidx[cb];
            if (p < 0) { singular = true; break; }
            const int pi = p >> 3, pg = p & 7;
            if (pg == g) avail &= ~(1u << pi);
            // pivot-row values of my two columns (held by lane 4*pg + t of this warp)
            double pr0 = 0.0, pr1 = 0.0;
            VK_SELECT_ROW(pi, 0, pr0);
            VK_SELECT_ROW(pi, 1, pr1);
            pr0 = __shfl_sync(0xffffffffu, pr0, (pg << 2) | t);
            pr1 = __shfl_sync(0xffffffffu, pr1, (pg << 2) | t);
            if (w == (k >> 3) && t == (kc >> 1)) {   // column k itself is already final: eliminate with 0
                if (kc & 1) pr1 = 0.0; else pr0 = 0.0;
            }
#pragma unroll
            for (int i = 0; i < NR; i++) {
                const double l = lbuf[cb * NIP + 8 * i + g];    // 0 for the pivot row
                A[i][0] = fma(-l, pr0, A[i][0]);
                A[i][1] = fma(-l, pr1, A[i][1]);
            }
        }
        if (singular) {
            if (tid == 0) a.status[col] = VK_ERR_SINGULAR;
            return;   // uniform: every thread read the same flag
        }
        // ---- un-permute into shared memory, applying the deferred pivot-row scaling:  W[q(r)][p(c)] = A[r][c] / pivot(r)
        {
            const int pc0 = piv_p[c0], pc1 = piv_p[c0 + 1];
#pragma unroll
            for (int i = 0; i < NR; i++) {
                const int qr = piv_q[8 * i + g];
                const double sc = rscale[8 * i + g];
                Wsm[qr * LD + pc0] = A[i][0] * sc;
                Wsm[qr * LD + pc1] = A[i][1] * sc;
            }
        }
        if (fuse && tid < NIP) {
            const double r = (tid < ni) ? rc[(size_t)j * ni + tid] : 0.0;
            tvec[tid] = (j == 0) ? r : r - dnc[(size_t)j * NIP + tid] * zprev[tid];
        }
        __syncthreads();
        double *Wj = Wc + (size_t)j * NIP * NIP;
        for (int q = tid; q < NIP * NIP / 2; q += NT) {
            const int r = (2 * q) / NIP, c = (2 * q) % NIP;
            *reinterpret_cast<double2 *>(Wj + (size_t)r * NIP + c) = *reinterpret_cast<const double2 *>(Wsm + r * LD + c);
        }
        if (fuse) {   // z_j = W_j (r_j - dn_j * z_{j-1}) : 4 threads per row
            const int row = tid >> 2, part = tid & 3;
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int i = 0; i < NIP / 8; i++) {
                const double2 wv = *reinterpret_cast<const double2 *>(Wsm + row * LD + i * 8 + part * 2);
                const double2 tv = *reinterpret_cast<const double2 *>(tvec + i * 8 + part * 2);
                acc0 = fma(wv.x, tv.x, acc0);
                acc1 = fma(wv.y, tv.y, acc1);
            }
            double acc = acc0 + acc1;
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (part == 0) { zprev[row] = acc; zc[(size_t)j * NIP + row] = acc; }
        }
        // the next iteration reads Wsm / zprev after its own __syncthreads-protected section
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
    int skip_fwd;                // z already holds the forward-eliminated vector (fused into the factorisation)
};

template <int NIP>
__global__ void __launch_bounds__(NIP * 4, (NIP <= 72) ? 2 : 1) solve_kernel(SolveArgs a)
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
    if (tid < NIP) zprev[tid] = a.skip_fwd ? zc[(size_t)(nz - 1) * NIP + tid] : 0.0;
    if (!a.skip_fwd) load_w(0);
    for (int j = 0; j < nz && !a.skip_fwd; j++) {
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
template <int NIP, int MINB>
static int launch_factor_t(vk_column *c, const double *D, const double *up, const double *dn, double *W, int *status,
                           const double *rhs, double *z)
{
    using C = FactorCfg<NIP>;
    static bool configured = false;
    if (!configured) {
        VK_CUDA(cudaFuncSetAttribute(factor_kernel<NIP, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured = true;
    }
    FactorArgs a{c->nz, c->ni, D, up, dn, W, status, rhs, z};
    factor_kernel<NIP, MINB><<<c->ncol, C::NT, C::SMEM, c->stream>>>(a);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

int launch_factor(vk_column *c, const double *D, const double *up, const double *dn, double *W, int *status, const double *rhs,
                  double *z)
{
    switch (c->nip) {
        case 48: return launch_factor_t<48, 2>(c, D, up, dn, W, status, rhs, z);
        case 72: return launch_factor_t<72, 2>(c, D, up, dn, W, status, rhs, z);
        case 96: return launch_factor_t<96, 1>(c, D, up, dn, W, status, rhs, z);
        case 120: return launch_factor_t<120, 1>(c, D, up, dn, W, status, rhs, z);
        default: set_error("no factor kernel for this padded block size"); return VK_ERR_UNSUPPORTED;
    }
}

int launch_solve(vk_column *c, const double *W, const double *up, const double *dn, const double *rhs, double *x, double *z, int skip_fwd)
{
    SolveArgs a{c->nz, c->ni, c->nip, W, up, dn, rhs, x, z, skip_fwd};
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
