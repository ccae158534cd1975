// Block-tridiagonal solver: replaces ODESolver.store_bandM + scipy.linalg.solve_banded (LAPACK dgbsv), op.py:2833-2858,
// 2914, 2929.  System per column:   dn_j x_{j-1} + D_j x_j + up_j x_{j+1} = r_j   with D_j dense ni x ni and up/dn DIAGONAL
// (the transport Jacobian freezes ysum), so the Schur complement of block Thomas is a two-sided diagonal
// scaling of the previous inverse:
//     S_0 = D_0,   S_j = D_j - diag(dn_j) W_{j-1} diag(up_{j-1}),   W_j = S_j^{-1}
//     z_j = W_j (r_j - dn_j * z_{j-1}),   x_{nz-1} = z_{nz-1},   x_j = z_j - W_j (up_j * x_{j+1})
// One thread block per column marches over the layers.  W_j is formed by blocked in-register Gauss-Jordan inversion on the
// FP64 tensor pipe (see factor_kernel) and is used ONLY for the Schur update of the next layer (it never leaves the registers).
// What is stored for the solves is the block LU factorisation F_j of S_j that the same sweep produces as a by-product (8 x 8
// diagonal blocks: P_K = A_KK^{-1} on, Lt_iK = A_iK P_K below, V_Kw = P_K A_Kw above the block diagonal): x = fl(S^{-1}) t is not
// backward stable - on the reference's own HD209S systems at production dt it leaves a residual of 1e-6 |r|, which the factor
// r*dt turns into a carbon budget error of 1e-2 ... 0.35 PER STEP - whereas the block-LU solve leaves 2e-14 like LAPACK's banded
// LU (tests/test_oracle_vs_reference.py::test_solve_conserves_elements, tests/test_gpu_parity.py).  The factor is stored once and
// reused for the second Ros2 stage and for iterative refinement - the reference factorises twice.
#include "vk_factor_dev.cuh"

namespace vk {

// ------------------------------------------------------------------------------------------------------------------
// lu_solve_kernel: forward + backward sweeps over the layers with the block LU factors F_j (see the header of this file).
//   forward :  z_j = S_j^{-1} (r_j - dn_j * z_{j-1})        backward :  x_j = z_j - S_j^{-1} (up_j * x_{j+1})
// One block per column, ONE THREAD PER BLOCK ROW.  F_j (NIP x (NIP+2) doubles) arrives by a 1-D TMA bulk copy into a double-buffered
// shared-memory slot one layer ahead; the application of S_j^{-1} is a block forward / backward substitution over the NIP/8 panels:
//   forward   K = 0 ..:   publish y_K ; barrier ; rows below:  y_i -= Lt_iK y_K                       (8 FMAs per row)
//   backward  K = .. 0:   rows of K:  z_K = P_K y_K - s_K  (y_K exchanged by shuffles) ; publish ; barrier ; rows above:  s_i += V_iK z_K
// i.e. 2 NIP/8 block barriers per layer and sweep, every row reading 64 contiguous bytes of its own F row per step.
struct LuSolveArgs {
    int nz, ni;
    const double *F, *up, *dn;   // padded layouts
    const double *rhs;           // [ncol][nz][ni]
    double *x;                   // [ncol][nz][ni]
    double *z;                   // [ncol][nz][NIP] scratch
    const int *act;              // optional per-column flags (refine = auto): columns with act == 0 are skipped
    const int *fwd_done;         // optional per-column flags: z already holds the forward-eliminated vector (fused into factor_kernel)
};

template <int NIP, int NBUF>
struct LuCfg {
    static constexpr int LDF = NIP + 2;
    static constexpr int NR = NIP / 8;
    static constexpr int NT = ((NIP + 31) / 32) * 32;
    static constexpr unsigned FBYTES = (unsigned)(sizeof(double) * NIP * LDF);
    static constexpr size_t SMEM = (size_t)NBUF * FBYTES + sizeof(double) * 2 * NIP + 32;
};

template <int NIP, int NBUF>
__global__ void __launch_bounds__(LuCfg<NIP, NBUF>::NT) lu_solve_kernel(LuSolveArgs a)
{
    using C = LuCfg<NIP, NBUF>;
    constexpr int LDF = C::LDF, NR = C::NR;
    extern __shared__ __align__(128) double smem[];
    double *fbuf = smem;                            // NBUF x NIP x LDF
    double *yv = fbuf + (size_t)NBUF * NIP * LDF;   // NIP  published y_K
    double *zv = yv + NIP;                          // NIP  published z_K
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(zv + NIP);   // NBUF mbarriers
    const int col = blockIdx.x, i = threadIdx.x, lane = i & 31, pan = i >> 3;
    if (a.act && !a.act[col]) return;
    const bool live = i < NIP;
    const int nz = a.nz, ni = a.ni;
    const double *Fc = a.F + (size_t)col * nz * NIP * LDF;
    const double *upc = a.up + (size_t)col * nz * NIP;
    const double *dnc = a.dn + (size_t)col * nz * NIP;
    const double *rc = a.rhs + (size_t)col * nz * ni;
    double *xc = a.x + (size_t)col * nz * ni;
    double *zc = a.z + (size_t)col * nz * NIP;

    auto fetch = [&](int j, int v) {     // one thread: F_j -> slot v % NBUF
        void *mb = mbar + (v % NBUF);
        mbar_expect_tx(mb, C::FBYTES);
        tma_load_1d(fbuf + (size_t)(v % NBUF) * NIP * LDF, Fc + (size_t)j * NIP * LDF, C::FBYTES, mb);
    };
    // S^{-1} t for my row (t enters as my component of the right-hand side).  The 64 bytes of my F row that the NEXT step reads are
    // loaded before that step's barrier (the factors are read-only here), products in four short chains: the in-layer chain of
    // 2 NIP/8 dependent steps is what bounds a single column.
    auto apply = [&](const double *Fb, double t) -> double {
        const double *Fr = Fb + (size_t)(live ? i : 0) * LDF;
        auto loadf = [&](int K, double2 (&f)[4]) {
            const double2 *p = reinterpret_cast<const double2 *>(Fr + 8 * K);
            f[0] = p[0]; f[1] = p[1]; f[2] = p[2]; f[3] = p[3];
        };
        auto dot8 = [](const double2 (&f)[4], const double2 *y) -> double {
            const double2 y0 = y[0], y1 = y[1], y2 = y[2], y3 = y[3];
            const double a0 = fma(f[0].y, y0.y, f[0].x * y0.x), a1 = fma(f[1].y, y1.y, f[1].x * y1.x);
            const double a2 = fma(f[2].y, y2.y, f[2].x * y2.x), a3 = fma(f[3].y, y3.y, f[3].x * y3.x);
            return (a0 + a1) + (a2 + a3);
        };
        double2 f[4];
        loadf(0, f);
#pragma unroll
        for (int K = 0; K < NR; K++) {
            if (pan == K) yv[i] = t;
            __syncthreads();
            if (live && pan > K) t -= dot8(f, reinterpret_cast<const double2 *>(yv + 8 * K));
            loadf(K + 1 < NR ? K + 1 : NR - 1, f);          // next forward step, or the first backward step (panel NR-1)
        }
        double s = 0.0, z = 0.0;
#pragma unroll
        for (int K = NR - 1; K >= 0; K--) {
            if (live && pan == K) {
                const unsigned mask = 0xffu << (lane & 24);
                const int base = lane & 24;
                double2 y[4];
#pragma unroll
                for (int c = 0; c < 4; c++) y[c] = make_double2(__shfl_sync(mask, t, base + 2 * c), __shfl_sync(mask, t, base + 2 * c + 1));
                z = dot8(f, y) - s;
                zv[i] = z;
            }
            __syncthreads();
            if (live && pan < K) s += dot8(f, reinterpret_cast<const double2 *>(zv + 8 * K));
            if (K > 0) loadf(K - 1, f);
        }
        return z;
    };

    if (i == 0) {
        for (int b = 0; b < NBUF; b++) mbar_init(mbar + b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // visits: layers 0 .. nz-1 forward, nz-2 .. 0 backward; a column whose forward elimination was fused into the factorisation starts
    // with the backward part (voff = nz)
    const bool skip = a.fwd_done && a.fwd_done[col];
    const int voff = skip ? nz : 0;
    const int nvisit = 2 * nz - 1 - voff;
    auto layer_of = [&](int v) { const int vv = v + voff; return vv < nz ? vv : 2 * nz - 2 - vv; };
    if (i == 0 && nvisit > 0) fetch(layer_of(0), 0);
    int v = 0;
    // ---- forward
    double zprev = (skip && live) ? zc[(size_t)(nz - 1) * NIP + i] : 0.0;
    double rn = (i < ni) ? rc[i] : 0.0, dnn = 0.0;
    for (int j = 0; j < nz && !skip; j++, v++) {
        double t = rn - dnn * zprev;
        if (j + 1 < nz) {
            rn = (i < ni) ? rc[(size_t)(j + 1) * ni + i] : 0.0;
            dnn = live ? dnc[(size_t)(j + 1) * NIP + i] : 0.0;
        }
        __syncthreads();                                    // every row is done with the slot the next fetch overwrites
        if (NBUF == 2 && i == 0 && v + 1 < nvisit) fetch(layer_of(v + 1), v + 1);
        mbar_wait(mbar + (v % NBUF), (v / NBUF) & 1);
        const double z = apply(fbuf + (size_t)(v % NBUF) * NIP * LDF, t);
        if (live) zc[(size_t)j * NIP + i] = z;
        zprev = z;
        if (NBUF == 1) {
            __syncthreads();
            if (i == 0 && v + 1 < nvisit) fetch(layer_of(v + 1), v + 1);
        }
    }
    // ---- backward (zprev = z_{nz-1} = x_{nz-1})
    double xnext = zprev;
    if (i < ni) xc[(size_t)(nz - 1) * ni + i] = xnext;
    double upn = 0.0, zj = 0.0;
    if (nz > 1 && live) { upn = upc[(size_t)(nz - 2) * NIP + i]; zj = zc[(size_t)(nz - 2) * NIP + i]; }
    for (int j = nz - 2; j >= 0; j--, v++) {
        const double t = upn * xnext, zcur = zj;
        if (j > 0 && live) { upn = upc[(size_t)(j - 1) * NIP + i]; zj = zc[(size_t)(j - 1) * NIP + i]; }
        __syncthreads();
        if (NBUF == 2 && i == 0 && v + 1 < nvisit) fetch(layer_of(v + 1), v + 1);
        mbar_wait(mbar + (v % NBUF), (v / NBUF) & 1);
        const double sv = apply(fbuf + (size_t)(v % NBUF) * NIP * LDF, t);
        const double xv = zcur - sv;
        if (i < ni) xc[(size_t)j * ni + i] = xv;
        xnext = xv;
        if (NBUF == 1) {
            __syncthreads();
            if (i == 0 && v + 1 < nvisit) fetch(layer_of(v + 1), v + 1);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// residual  res = rhs - A x  for iterative refinement, accumulated in DOUBLE-DOUBLE (error-free products by FMA, error-free sums;
// Ogita / Rump / Oishi "Dot2"): one block per (column, layer), 4 threads per row.
// Why extended precision: the element budget of a solve is  compo^T (rhs - A x) * r * dt  (chemistry conserves elements, compo^T J = 0).
// A residual evaluated in plain fp64 carries a rounding error eps |A| |x| of its own - as large as the residual a backward-stable
// solve leaves - so fp64 refinement cannot lower the budget error (measured on the reference's HD209S system at dt = 2.4e5 s:
// 7e-6 -> 6e-6 -> 1e-5), whereas the same pass with an exact residual gains 20 - 50 x (7e-6 -> 4e-7 -> 2e-8); DESIGN.md section 4.2.
// `act` (optional): per-column flags, columns with act == 0 are skipped (refine = auto).
struct ResidArgs {
    int nz, ni, nip;
    const double *D, *up, *dn, *rhs, *x;
    double *res;
    const int *act;         // [ncol] or NULL: columns with act == 0 are skipped (refine = auto)
};
struct dd_t { double hi, lo; };
__device__ __forceinline__ void dd_fma_acc(dd_t &s, double a, double b)
{
    const double p = __dmul_rn(a, b);
    const double pe = __fma_rn(a, b, -p);                 // a b = p + pe exactly
    const double t = __dadd_rn(s.hi, p);
    const double bb = __dsub_rn(t, s.hi);
    const double e = __dadd_rn(__dsub_rn(s.hi, __dsub_rn(t, bb)), __dsub_rn(p, bb));     // s.hi + p = t + e exactly
    s.hi = t;
    s.lo = __dadd_rn(s.lo, __dadd_rn(e, pe));
}
__device__ __forceinline__ void dd_add(dd_t &s, double ohi, double olo)
{
    const double t = __dadd_rn(s.hi, ohi);
    const double bb = __dsub_rn(t, s.hi);
    const double e = __dadd_rn(__dsub_rn(s.hi, __dsub_rn(t, bb)), __dsub_rn(ohi, bb));
    s.hi = t;
    s.lo = __dadd_rn(__dadd_rn(s.lo, olo), e);
}
__global__ void __launch_bounds__(512) resid_kernel(ResidArgs a, int n_blocks)
{
    extern __shared__ double xs[];   // x_j padded
    const int nz = a.nz, ni = a.ni, nip = a.nip;
    const int tid = threadIdx.x;
    // a bounded grid walks the (column, layer) pairs: with refine = auto most launches find every column inactive, and a grid of
    // ncol * nz empty blocks costs milliseconds of block scheduling (4096 columns: 614 400 blocks)
    if (a.act) {           // nothing to do at all?  (one pass over the flags instead of one flag load per (column, layer) pair)
        const int ncol = n_blocks / nz;
        int any = 0;
        for (int c0 = tid; c0 < ncol; c0 += blockDim.x) any |= a.act[c0];
        if (!__syncthreads_or(any)) return;
    }
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int col = blk / nz, j = blk % nz;
        if (a.act && !a.act[col]) continue;
        const size_t vb = ((size_t)col * nz + j) * ni;
        __syncthreads();
        for (int i = tid; i < nip; i += blockDim.x) xs[i] = (i < ni) ? a.x[vb + i] : 0.0;
        __syncthreads();
        const int row = tid >> 2, part = tid & 3;
        const bool live = row < nip;
        const double *Dr = a.D + (((size_t)col * nz + j) * nip + (live ? row : 0)) * nip;
        dd_t acc{0.0, 0.0};
        for (int c = part * 2; live && c < nip; c += 8) {
            double2 d = *reinterpret_cast<const double2 *>(Dr + c);
            dd_fma_acc(acc, d.x, xs[c]);
            dd_fma_acc(acc, d.y, xs[c + 1]);
        }
        {
            double oh = __shfl_xor_sync(0xffffffffu, acc.hi, 1), ol = __shfl_xor_sync(0xffffffffu, acc.lo, 1);
            dd_add(acc, oh, ol);
            oh = __shfl_xor_sync(0xffffffffu, acc.hi, 2); ol = __shfl_xor_sync(0xffffffffu, acc.lo, 2);
            dd_add(acc, oh, ol);
        }
        if (part == 0 && row < ni) {
            const size_t pb = ((size_t)col * nz + j) * nip;
            if (j + 1 < nz) dd_fma_acc(acc, a.up[pb + row], a.x[vb + ni + row]);
            if (j > 0) dd_fma_acc(acc, a.dn[pb + row], a.x[vb - ni + row]);
            dd_t r{a.rhs[vb + row], 0.0};
            dd_add(r, -acc.hi, -acc.lo);
            a.res[vb + row] = __dadd_rn(r.hi, r.lo);
        }
    }
}

// refine = auto: which columns are refined at all (step size >= dt_min; every column when dt is NULL)
__global__ void refine_init_kernel(int ncol, const double *dt, double dt_min, const int *col_act, int *act)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col < ncol) act[col] = ((!dt || dt[col] >= dt_min) && (!col_act || col_act[col])) ? 1 : 0;
}
// xn = x + dx
__global__ void refine_axpy_kernel(size_t per, const double *x, const double *dx, double *xn, const int *act)
{
    const int col = blockIdx.y;
    if (act && !act[col]) return;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < per; q += (size_t)gridDim.x * blockDim.x)
        xn[col * per + q] = x[col * per + q] + dx[col * per + q];
}

// Safeguard of a refinement pass (refine = auto): the pass is kept only if it lowers the ELEMENT-WEIGHTED residual - the quantity it is
// there to protect:  E_a = sum_{j,i} compo[i][a] res[j][i],  weighed per element a against  N_a = sum compo[i][a] |x[j][i]|.
// A kept pass whose remaining element-budget error  |E_a| r dt / T_a  (T_a = atoms of element a in the column) is still above `tol`
// leaves the column active for another pass; a rejected pass (the 1-ulp sensitivity of these systems makes ~1 pass in 10 diverge at
// dt > 1e5 s) ends the refinement of that column.  One block per column; every sum is formed in a fixed order
// (thread q owns the entries q, q + 256, ...; fixed shuffle tree; warps added 0..7), so the decision - and with it the result - is
// reproducible and independent of the batch a column sits in.
struct SelectArgs {
    int nz, ni, na;
    const double *compo;       // [ni][na]
    double *res0;              // residual before the pass; receives the residual after it when the pass is kept
    const double *res1;
    const double *xn, *y;
    double *x;
    const double *dt;          // [ncol] or NULL
    double tol;
    int *act;
    int *kept, *tried;         // [ncol] counters
};
__global__ void __launch_bounds__(256) refine_select_kernel(SelectArgs a)
{
    __shared__ double red[4][8][8];      // [quantity][element][warp]
    __shared__ int take;
    const int col = blockIdx.x, tid = threadIdx.x, na = a.na;
    if (!a.act[col]) return;
    const size_t per = (size_t)a.nz * a.ni, base = col * per;
    double e0[8], e1[8], nn[8], tt[8];
    for (int q = 0; q < 8; q++) e0[q] = e1[q] = nn[q] = tt[q] = 0.0;
    for (size_t q = tid; q < per; q += 256) {
        const int i = (int)(q % a.ni);
        const double r0 = a.res0[base + q], r1 = a.res1[base + q], ax = fabs(a.x[base + q]), ay = a.y ? fabs(a.y[base + q]) : 0.0;
        for (int at = 0; at < na; at++) {
            const double w = a.compo[i * na + at];
            e0[at] = fma(w, r0, e0[at]); e1[at] = fma(w, r1, e1[at]); nn[at] = fma(w, ax, nn[at]); tt[at] = fma(w, ay, tt[at]);
        }
    }
    for (int at = 0; at < na; at++) {
        for (int off = 16; off > 0; off >>= 1) {
            e0[at] += __shfl_xor_sync(0xffffffffu, e0[at], off);
            e1[at] += __shfl_xor_sync(0xffffffffu, e1[at], off);
            nn[at] += __shfl_xor_sync(0xffffffffu, nn[at], off);
            tt[at] += __shfl_xor_sync(0xffffffffu, tt[at], off);
        }
        if ((tid & 31) == 0) { red[0][at][tid >> 5] = e0[at]; red[1][at][tid >> 5] = e1[at]; red[2][at][tid >> 5] = nn[at]; red[3][at][tid >> 5] = tt[at]; }
    }
    __syncthreads();
    if (tid == 0) {
        double w0 = 0.0, w1 = 0.0, bud1 = 0.0;
        bool finite = true;
        const double rdt = a.dt ? (1. + 1. / sqrt(2.)) * a.dt[col] : 0.0;
        for (int at = 0; at < na; at++) {
            double s0 = 0.0, s1 = 0.0, n = 0.0, t = 0.0;
            for (int w = 0; w < 8; w++) { s0 += red[0][at][w]; s1 += red[1][at][w]; n += red[2][at][w]; t += red[3][at][w]; }
            if (!(n > 0.0)) continue;
            w0 = fmax(w0, fabs(s0) / n);
            w1 = fmax(w1, fabs(s1) / n);
            if (t > 0.0) bud1 = fmax(bud1, fabs(s1) * rdt / t);
            finite = finite && (s1 == s1);
        }
        take = (finite && w1 < w0) ? 1 : 0;
        a.tried[col] += 1;
        if (take) a.kept[col] += 1;
        // another pass?  with dt and y known: while the remaining element-budget error is above the tolerance; without (component entry
        // point vk_blocktri_solve): while passes keep being accepted
        a.act[col] = (take && (!(a.dt && a.y) || bud1 > a.tol)) ? 1 : 0;
    }
    __syncthreads();
    if (take)
        for (size_t q = tid; q < per; q += 256) { a.x[base + q] = a.xn[base + q]; a.res0[base + q] = a.res1[base + q]; }
}

// ------------------------------------------------------------------------------------------------------------------
template <int NIP, int MINB>
static int launch_factor_t(vk_column *c, const double *D, const double *up, const double *dn, double *F, int *status, const double *fwd_rhs)
{
    using C = FactorCfg<NIP>;
    FactorArgs a{c->nz, c->ni, D, up, dn, F, status, c->act, nullptr, nullptr, 0, fwd_rhs, c->z, c->dt, c->opts.refine_dt_min, c->fwd_done};
    if (!fwd_rhs) a.fwd_done = nullptr;
    const size_t smem = fwd_rhs ? C::SMEM_FWD : C::SMEM;
    { int rc = ensure_smem((const void *)factor_kernel<NIP, MINB, false, NoProducer>, c->net->device, C::SMEM_FWD); if (rc) return rc; }
    factor_kernel<NIP, MINB, false, NoProducer><<<c->ncol, C::NT, smem, c->stream>>>(a, NoProducer{});
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

// F out: block LU factors of the Schur blocks, [ncol][nz][nip][nip+2]
// fwd_rhs (optional): right-hand side of the first solve; its forward elimination is fused into the sweep for the columns with
// dt < refine_dt_min (c->fwd_done[col] = 1, z in c->z; pass c->fwd_done to that solve)
int launch_factor(vk_column *c, const double *D, const double *up, const double *dn, double *F, int *status, const double *fwd_rhs)
{
    static int env_fwd = -1;
    if (env_fwd < 0) { const char *e = getenv("VK_FWD_FUSED"); env_fwd = e ? atoi(e) : 1; }
    if (!env_fwd || !c->fwd_done) fwd_rhs = nullptr;
    c->fwd_valid = 0;
    if (c->cr_now) return launch_cr_factor(c, D, up, dn, F, status);
    c->fwd_valid = fwd_rhs != nullptr;
    switch (c->nip) {
        case 48: return launch_factor_t<48, 2>(c, D, up, dn, F, status, fwd_rhs);
        case 72: return launch_factor_t<72, 2>(c, D, up, dn, F, status, fwd_rhs);      // (3 blocks per SM at <= 56 registers measured SLOWER: 25.3 vs 23.6 ms, spills)
        case 96: return launch_factor_t<96, 1>(c, D, up, dn, F, status, fwd_rhs);
        case 120: return launch_factor_t<120, 1>(c, D, up, dn, F, status, fwd_rhs);
        default: set_error("no factor kernel for this padded block size"); return VK_ERR_UNSUPPORTED;
    }
}

template <int NIP, int NBUF>
static int launch_lu_solve_t(vk_column *c, const LuSolveArgs &a)
{
    using C = LuCfg<NIP, NBUF>;
    { int rc = ensure_smem((const void *)lu_solve_kernel<NIP, NBUF>, c->net->device, C::SMEM); if (rc) return rc; }
    lu_solve_kernel<NIP, NBUF><<<c->ncol, C::NT, C::SMEM, c->stream>>>(a);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

// x = A^{-1} rhs with the stored block LU factors F ([ncol][nz][nip][nip+2]); z is scratch
int launch_solve(vk_column *c, const double *F, const double *up, const double *dn, const double *rhs, double *x, double *z, const int *act,
                 const int *fwd_done)
{
    if (c->cr_now) return launch_cr_solve(c, F, up, dn, rhs, x, act);
    LuSolveArgs a{c->nz, c->ni, F, up, dn, rhs, x, z, act, fwd_done};
    // slots of the F prefetch per block: 2 = the next layer's copy overlaps this layer's substitution inside the block (few columns:
    // nothing else hides the copy latency), 1 = more blocks per SM hide it instead.  Measured, 592 HD189 columns: 1 slot (5 blocks per
    // SM) 1.28 ms = 93 % of the measured HBM peak, 2 slots 1.87 ms; one column: 2 slots.
    static int env_nbuf = -1;
    if (env_nbuf < 0) { const char *e = getenv("VK_LU_NBUF"); env_nbuf = e ? atoi(e) : 0; }
    const int nbuf = env_nbuf ? env_nbuf : (c->ncol > 148 ? 1 : 2);       // (column groups of the pipelined host path share the SMs: 1 slot there too)
    switch (c->nip) {
        case 48: return nbuf == 1 ? launch_lu_solve_t<48, 1>(c, a) : launch_lu_solve_t<48, 2>(c, a);
        case 72: return nbuf == 1 ? launch_lu_solve_t<72, 1>(c, a) : launch_lu_solve_t<72, 2>(c, a);
        case 96: return nbuf == 1 ? launch_lu_solve_t<96, 1>(c, a) : launch_lu_solve_t<96, 2>(c, a);
        case 120: return launch_lu_solve_t<120, 1>(c, a);      // two slots of 117 KB do not fit
        default: set_error("no solve kernel for this padded block size"); return VK_ERR_UNSUPPORTED;
    }
}

int launch_residual(vk_column *c, const double *D, const double *up, const double *dn, const double *rhs, const double *x, double *res,
                    const int *act)
{
    ResidArgs a{c->nz, c->ni, c->nip, D, up, dn, rhs, x, res, act};
    const int n_blocks = c->ncol * c->nz;
    resid_kernel<<<std::min(n_blocks, 148 * 16), 512, sizeof(double) * c->nip, c->stream>>>(a, n_blocks);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

// x <- A^{-1} rhs refined.  refine > 0: that many passes x += A^{-1}(rhs - A x) with the double-double residual.  refine < 0 (AUTO): up
// to -refine... VK_REFINE_AUTO_PASSES passes on the columns whose step size is >= opts.refine_dt_min (dt_pred = their dt; NULL: every
// column), each kept only if it lowers the element-weighted residual, until the element-budget error of the solve is below
// VK_REFINE_BUDGET_TOL (refine_select_kernel).  Work vectors: c->res, c->dx, c->xn; flags c->refine_act.
#define VK_REFINE_AUTO_PASSES 4
#define VK_REFINE_BUDGET_TOL 1.0e-11
int launch_refine(vk_column *c, const double *D, const double *up, const double *dn, const double *F, const double *rhs, double *x,
                  int refine, const double *dt_pred)
{
    int rc = VK_OK;
    const size_t per = (size_t)c->nz * c->ni;
    dim3 grid((unsigned)std::min<size_t>((per + 255) / 256, c->ncol > 256 ? 4 : 32), (unsigned)c->ncol);
    if (refine > 0) {
        for (int it = 0; it < refine && rc == VK_OK; it++) {
            rc = launch_residual(c, D, up, dn, rhs, x, c->res, c->act);
            if (rc == VK_OK) rc = launch_solve(c, F, up, dn, c->res, c->dx, c->z, c->act);
            if (rc == VK_OK) {
                refine_axpy_kernel<<<grid, 256, 0, c->stream>>>(per, x, c->dx, x, c->act);
                VK_CUDA(cudaGetLastError());
            }
        }
        return rc;
    }
    if (refine == 0) return VK_OK;
    if (dt_pred && c->dt_host_max >= 0.0 && c->dt_host_max < c->opts.refine_dt_min) return VK_OK;    // no column is refined: launch nothing
    if (!c->opts.compo || c->opts.na < 1) { set_error("refine = auto needs the element composition (vk_step_opts.compo)"); return VK_ERR_INVALID; }
    int *act = c->refine_act;
    refine_init_kernel<<<(c->ncol + 127) / 128, 128, 0, c->stream>>>(c->ncol, dt_pred, c->opts.refine_dt_min, c->act, act);
    VK_CUDA(cudaGetLastError());
    if ((rc = launch_residual(c, D, up, dn, rhs, x, c->res, act))) return rc;
    const int passes = (refine == -1) ? VK_REFINE_AUTO_PASSES : -refine;
    for (int it = 0; it < passes; it++) {
        if ((rc = launch_solve(c, F, up, dn, c->res, c->dx, c->z, act))) return rc;
        refine_axpy_kernel<<<grid, 256, 0, c->stream>>>(per, x, c->dx, c->xn, act);
        VK_CUDA(cudaGetLastError());
        if ((rc = launch_residual(c, D, up, dn, rhs, c->xn, c->dx, act))) return rc;
        SelectArgs sa{c->nz, c->ni, c->opts.na, c->opts.compo, c->res, c->dx, c->xn, dt_pred ? c->y : nullptr, x, dt_pred,
                      VK_REFINE_BUDGET_TOL, act, c->refine_kept, c->refine_tried};
        refine_select_kernel<<<c->ncol, 256, 0, c->stream>>>(sa);
        VK_CUDA(cudaGetLastError());
        // ONE column (the latency path): a further pass is ~30 launches of the cyclic-reduction solve whether or not the column is still
        // being refined - ask the flag back (4 bytes, ~10 us) and stop launching once it is done (typically after the first pass)
        if (c->ncol == 1 && it + 1 < passes) {
            int still = 1;
            VK_CUDA(cudaMemcpyAsync(&still, act, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            VK_CUDA(cudaStreamSynchronize(c->stream));
            if (!still) break;
        }
    }
    return VK_OK;
}

}  // namespace vk

#include "vk_cr.inl"
