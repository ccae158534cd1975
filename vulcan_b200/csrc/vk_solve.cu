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
#include <type_traits>

#include "vk_internal.cuh"

namespace vk {

// 1/x without the IEEE slow path: hardware approximation (2^-23) + two Newton steps (error ~1 ulp); x = 0 / inf / nan give
// inf / 0 / nan, caught by the singular-pivot flag
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// Register layout: warp w owns the 8 columns [8w, 8w+8) of the block and ALL its rows, in the m8n8 accumulator-fragment
// pattern stacked vertically: lane = 4*g + t holds rows 8*i + g (i < NIP/8) and columns 8w + 2t, 8w + 2t + 1 (the layout of
// mma.sync.m8n8k4.f64 accumulators).
//
// Pivoting: DIAGONAL pivots.  Measured on the reference's own matrices (tests + DESIGN.md §4.1): for these systems
// (1/(r h) I - J with loss terms on the diagonal; species abundances spanning 30+ decades so that rows carry wildly
// different scales) Gauss-Jordan on the diagonal is 2-5 orders of magnitude MORE accurate than LAPACK-style partial pivoting,
// which lets the largest entry of a column - a scale artefact - destroy componentwise accuracy.  A zero / non-finite pivot
// sets VK_ERR_SINGULAR for the column: the step is rejected and retried with dt/2 exactly like any other failed step.
//
// With the pivot row known in advance the per-pivot work is: owner warp (k/8) forms the multipliers a_rk / a_kk of its column
// and publishes them (double-buffered shared memory); ONE __syncthreads; every warp: 9 LDS + pivot-row broadcast by warp
// shuffle from a compile-time register (the k loop is unrolled over the row tile) + 18 FMAs.  Pivot-row scaling is deferred to
// the end of the layer, so the inverse W_j stays in registers in its natural layout: the Schur update of the next layer
// (elementwise, the couplings are diagonal), the write-out and the fused forward elimination all work from registers.
template <int NIP>
struct FactorCfg {
    static constexpr int NW = NIP / 8;           // warps
    static constexpr int NR = NIP / 8;           // rows per lane
    static constexpr int NT = NW * 32;
    // lbuf[2][NIP] + rscale[NIP] + tvec[NIP] + zpart[NW][NIP] (doubles)
    static constexpr size_t SMEM = sizeof(double) * ((size_t)(4 + NW) * NIP);
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

#ifdef VK_TRACE
__device__ long long g_trace[128 * 8];
#define TRACE(kk, ev) do { if (j == 5 && (kk) < 128 && lane == 0) g_trace[(kk) * 8 + (ev)] = clock64(); } while (0)
#else
#define TRACE(kk, ev) do { } while (0)
#endif

template <int NIP, int MINB>
__global__ void __launch_bounds__(FactorCfg<NIP>::NT, MINB) factor_kernel(FactorArgs a)
{
    using C = FactorCfg<NIP>;
    constexpr int NR = C::NR, NW = C::NW;
    extern __shared__ __align__(16) double smem[];
    double *lbuf = smem;                 // 2 x NIP   multipliers of pivot step k (double buffered)
    double *rscale = lbuf + 2 * NIP;     // NIP       1/pivot of every row (deferred pivot-row scaling)
    double *tvec = rscale + NIP;         // NIP       r_j - dn_j * z_{j-1}
    double *zpart = tvec + NIP;          // NW x NIP  per-warp partial sums of W_j tvec

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
    double zreg = 0.0;                   // z_{j-1}[tid] for tid < NIP
    int bad = 0;

    for (int j = 0; j < nz; j++) {
        // ---- S_j = D_j - diag(dn_j) W_{j-1} diag(up_{j-1}); W_{j-1} is still in A
        {
            const double *Dj = Dc + (size_t)j * NIP * NIP;
            double u0 = 0.0, u1 = 0.0;
            if (j > 0) { u0 = upc[(size_t)(j - 1) * NIP + c0]; u1 = upc[(size_t)(j - 1) * NIP + c0 + 1]; }
#pragma unroll
            for (int i = 0; i < NR; i++) {
                const int r = 8 * i + g;
                const double2 d = *reinterpret_cast<const double2 *>(Dj + (size_t)r * NIP + c0);
                if (j > 0) {
                    const double l = dnc[(size_t)j * NIP + r];
                    A[i][0] = d.x - (l * A[i][0]) * u0;
                    A[i][1] = d.y - (l * A[i][1]) * u1;
                } else {
                    A[i][0] = d.x;
                    A[i][1] = d.y;
                }
            }
        }
        if (fuse && tid < NIP) {
            const double r = (tid < ni) ? rc[(size_t)j * ni + tid] : 0.0;
            tvec[tid] = (j == 0) ? r : r - dnc[(size_t)j * NIP + tid] * zreg;
        }
        if (tid < NIP) rscale[tid] = 1.0;
        // ---- Gauss-Jordan on the diagonal.  lbuf holds the NEGATED multipliers -a_rk/a_kk (0 in the pivot row); they are
        // also the new column k of the transformed block (with 1 in the pivot row, which stays unscaled until write-out).
        // PUBLISH(KT, G, E, H, CB): the warp owning column 8*KT+G (plane E of lane pair H) forms and publishes them.
#define VK_PUBLISH(KT, G, E, H, CB)                                                                                    \
        {                                                                                                              \
            const double dkk = __shfl_sync(0xffffffffu, A[KT][E], ((G) << 2) | (H));                                   \
            const double ninv = -fast_rcp(dkk);                                                                        \
            if (!(fabs(dkk) > 0.0)) bad = 1;                                                                           \
            if (t == (H)) {                                                                                            \
                _Pragma("unroll") for (int i = 0; i < NR; i++) {                                                       \
                    A[i][E] = A[i][E] * ninv;                                                                          \
                    if (i == (KT)) { if (g == (G)) A[i][E] = 0.0; }                                                    \
                    lbuf[(CB) * NIP + 8 * i + g] = A[i][E];                                                            \
                }                                                                                                      \
                if (g == (G)) { A[KT][E] = 1.0; rscale[8 * (KT) + (G)] = -ninv; }                                      \
            }                                                                                                          \
        }
        if (w == 0) VK_PUBLISH(0, 0, 0, 0, 0)
        // one row tile of pivots; KT and the plane E are compile-time constants so that A[KT][E] is a register
        auto pivot_tile = [&](auto ktc) {
            constexpr int kt = decltype(ktc)::value;
            constexpr int ktn = (kt + 1 < NR) ? kt + 1 : kt;
            for (int h = 0; h < 4; h++) {
                // ---------------- pivot k = 8 kt + 2 h (plane 0); the next column is plane 1 of the same lane pair
                {
                    const int k = 8 * kt + 2 * h;
                    if (k >= ni) break;
                    __syncthreads();                              // multipliers of step k visible (buffer 0: k is even)
                    double nl[NR];
#pragma unroll
                    for (int i = 0; i < NR; i++) nl[i] = lbuf[8 * i + g];
                    const int src = ((2 * h) << 2) | t;
                    const double pr1 = __shfl_sync(0xffffffffu, A[kt][1], src);
                    double pr0 = __shfl_sync(0xffffffffu, A[kt][0], src);
                    if (w == kt && t == h) pr0 = 0.0;             // column k itself is final
#pragma unroll
                    for (int i = 0; i < NR; i++) A[i][1] = fma(nl[i], pr1, A[i][1]);
                    if (w == kt && k + 1 < ni) VK_PUBLISH(kt, 2 * h + 1, 1, h, 1)
#pragma unroll
                    for (int i = 0; i < NR; i++) A[i][0] = fma(nl[i], pr0, A[i][0]);
                }
                // ---------------- pivot k = 8 kt + 2 h + 1 (plane 1); the next column is plane 0 of the next lane pair / tile
                {
                    const int k = 8 * kt + 2 * h + 1;
                    if (k >= ni) break;
                    __syncthreads();                              // buffer 1: k is odd
                    double nl[NR];
#pragma unroll
                    for (int i = 0; i < NR; i++) nl[i] = lbuf[NIP + 8 * i + g];
                    const int src = ((2 * h + 1) << 2) | t;
                    const double pr0 = __shfl_sync(0xffffffffu, A[kt][0], src);
                    double pr1 = __shfl_sync(0xffffffffu, A[kt][1], src);
                    if (w == kt && t == h) pr1 = 0.0;             // column k itself is final
#pragma unroll
                    for (int i = 0; i < NR; i++) A[i][0] = fma(nl[i], pr0, A[i][0]);
                    if (k + 1 < ni) {
                        if (h < 3) { if (w == kt) VK_PUBLISH(kt, 2 * h + 2, 0, h + 1, 0) }
                        else if (kt + 1 < NR) { if (w == ktn) VK_PUBLISH(ktn, 0, 0, 0, 0) }
                    }
#pragma unroll
                    for (int i = 0; i < NR; i++) A[i][1] = fma(nl[i], pr1, A[i][1]);
                }
            }
        };
#define VK_TILE(N) if constexpr ((N) < NR) { if (8 * (N) < ni) pivot_tile(std::integral_constant<int, (N)>{}); }
        VK_TILE(0) VK_TILE(1) VK_TILE(2) VK_TILE(3) VK_TILE(4) VK_TILE(5) VK_TILE(6) VK_TILE(7)
        VK_TILE(8) VK_TILE(9) VK_TILE(10) VK_TILE(11) VK_TILE(12) VK_TILE(13) VK_TILE(14)
#undef VK_TILE
#undef VK_PUBLISH
        if (__syncthreads_or(bad)) {
            if (tid == 0) a.status[col] = VK_ERR_SINGULAR;
            return;
        }
        // ---- W_j = diag(rscale) A : stays in registers for the next layer; written out; fused forward elimination
        double *Wj = Wc + (size_t)j * NIP * NIP;
        double tv0 = 0.0, tv1 = 0.0;
        if (fuse) { tv0 = tvec[c0]; tv1 = tvec[c0 + 1]; }
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const int r = 8 * i + g;
            const double sc = rscale[r];
            A[i][0] *= sc;
            A[i][1] *= sc;
            *reinterpret_cast<double2 *>(Wj + (size_t)r * NIP + c0) = make_double2(A[i][0], A[i][1]);
            if (fuse) {
                double part = fma(A[i][0], tv0, A[i][1] * tv1);
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                part += __shfl_xor_sync(0xffffffffu, part, 2);
                if (t == 0) zpart[w * NIP + r] = part;
            }
        }
        __syncthreads();
        if (fuse && tid < NIP) {   // z_j = W_j (r_j - dn_j * z_{j-1})
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < NW; q++) acc += zpart[q * NIP + tid];
            zreg = acc;
            zc[(size_t)j * NIP + tid] = acc;
        }
        // (the next layer's first __syncthreads orders the zpart / tvec / rscale reuse)
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


#ifdef VK_TRACE
extern "C" int vk_debug_trace(long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, vk::g_trace, sizeof(long long) * 128 * 8);
}
#endif
