// Single-column latency path: BLOCK CYCLIC REDUCTION over the layers (BASELINE north_star bullet 2).
// Included at the end of vk_solve.cu (same translation unit: factor_kernel, the TMA helpers).
//
// Block Thomas (factor_kernel / lu_solve_kernel) is a chain of nz dependent layer steps on ONE SM: 1.04 ms to factor and 0.79 ms per solve
// for one HD189 column while 147 SMs idle.  Cyclic reduction eliminates every second layer of the current system at once - log2(nz) = 8
// levels instead of 150 steps - and every level is a batch of independent dense block operations spread over the SMs:
//
//   level l, node set I_l (I_0 = all layers), eliminated E_l = odd positions of I_l, kept I_{l+1} = even positions
//     for p in E_l      : F_p, W_p = block LU and explicit inverse of the current diagonal block D_p   (factor_kernel, block-list mode)
//                         G^L_p = W_p L_p,  G^U_p = W_p U_p                                            (DMMA GEMMs; level 0: L, U diagonal)
//     for a in I_{l+1}  : D_a -= U_a G^L_next + L_a G^U_prev ;  U'_a = -U_a G^U_next ;  L'_a = -L_a G^L_prev
//   (L_x / U_x = coupling of node x to its left / right neighbour in I_l; diagonal at level 0, dense from level 1 on).
//   solve, forward  l = 0..  : w_p = F_p^{-1} r_p (p in E_l) ;  r_a -= U_a w_next + L_a w_prev (a in I_{l+1})
//          last node         : x = F^{-1} r
//          backward l = ..0  : x_p = w_p - F_p^{-1} (L_p x_left + U_p x_right)
// As in block Thomas the explicit inverses enter only MATRIX products (the Schur complements); every application of D_p^{-1} to a vector
// goes through the block LU factors F_p (backward stable).  The dense fill-in makes cyclic reduction execute
// ~4 x the flops of block Thomas (12 n^3 of GEMM per eliminated dense node) - reported as executed flops next to the algorithmic count.
//
// Everything lives in a per-column pool of NIP x NIP blocks:  [0, nz) current diagonal blocks | [nz, 2 nz) explicit inverses |
// coupling slots (L, U of every node alive at level >= 1) | G^L, G^U scratch.   F factors: c->W (one NIP x (NIP+2) block per layer).

namespace vk {

// ---- t <- S^{-1} t with the block LU factors in shared memory: the substitution of lu_solve_kernel as a device function (one thread per row)
template <int NIP>
__device__ __forceinline__ double lu_apply_dev(const double *Fb, double t, double *yv, double *zv, int i)
{
    constexpr int LDF = NIP + 2, NR = NIP / 8;
    const int lane = i & 31, pan = i >> 3;
    const bool live = i < NIP;
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
    __syncthreads();
#pragma unroll
    for (int K = 0; K < NR; K++) {
        if (pan == K) yv[i] = t;
        __syncthreads();
        if (live && pan > K) t -= dot8(f, reinterpret_cast<const double2 *>(yv + 8 * K));
        loadf(K + 1 < NR ? K + 1 : NR - 1, f);
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
}

// ---- plan of the reduction for a given nz: node lists and job tables, built on the host once per handle
struct CrElim { int node, left, right, sL, sU; };          // eliminated node, its neighbours in I_l (-1: none), pool ids of its couplings (-1: level 0, diagonal)
struct CrKept { int node, prev, next, sL, sU, sLn, sUn; }; // kept node, the eliminated neighbours on its sides, its couplings now / at the next level
struct CrJob { int a1, b1, a2, b2, c, beta, sign; };       // pool block c = beta c + sign (a1 b1 + a2 b2);  a1 < 0: c = 0;  a2 < 0: one product
struct CrLevel { int n_elim, n_kept, elim_off, kept_off, n_gjob, gjob_off, n_kjob, kjob_off; };

}  // namespace vk

struct CrPlan {
    int nz, nip, n_levels, last_node, pool_blocks;
    std::vector<vk::CrLevel> levels;
    vk::CrElim *d_elim; vk::CrKept *d_kept; vk::CrJob *d_jobs; int *d_idx;   // device tables (d_idx: eliminated node ids per level, block-list mode)
    double *pool;          // [pool_blocks][nip][nip]
    double *rv, *wv, *xv;  // [nz][nip] working vectors of a solve
    size_t exec_flops;     // flops one factorisation executes (inversions + GEMMs)
};

namespace vk {

void cr_plan_free(CrPlan *p)
{
    if (!p) return;
    cudaFree(p->d_elim); cudaFree(p->d_kept); cudaFree(p->d_jobs); cudaFree(p->d_idx); cudaFree(p->pool);
    cudaFree(p->rv); cudaFree(p->wv); cudaFree(p->xv);
    delete p;
}

static int cr_plan_build(int nz, int nip, CrPlan **out)
{
    CrPlan *P = new CrPlan();
    P->nz = nz; P->nip = nip;
    P->d_elim = nullptr; P->d_kept = nullptr; P->d_jobs = nullptr; P->d_idx = nullptr; P->pool = nullptr; P->rv = P->wv = P->xv = nullptr;
    std::vector<int> alive(nz), sL(nz, -1), sU(nz, -1);     // current coupling slots of the alive nodes (-1: diagonal, level 0)
    for (int j = 0; j < nz; j++) alive[j] = j;
    std::vector<CrElim> elim;
    std::vector<CrKept> kept;
    std::vector<CrJob> jobs;
    int next_block = 2 * nz, max_elim = 0;
    const size_t n3 = (size_t)nip * nip * nip;
    size_t flops = 0;
    while (alive.size() > 1) {
        const int na = (int)alive.size();
        CrLevel L{};
        L.elim_off = (int)elim.size(); L.kept_off = (int)kept.size();
        std::vector<int> nxt;
        for (int q = 0; q < na; q++) {
            const int node = alive[q];
            if (q & 1) {
                elim.push_back(CrElim{node, alive[q - 1], q + 1 < na ? alive[q + 1] : -1, sL[node], sU[node]});
            } else {
                CrKept k{node, q > 0 ? alive[q - 1] : -1, q + 1 < na ? alive[q + 1] : -1, sL[node], sU[node], next_block, next_block + 1};
                next_block += 2;
                kept.push_back(k);
                nxt.push_back(node);
            }
        }
        L.n_elim = na / 2; L.n_kept = (na + 1) / 2;
        max_elim = std::max(max_elim, L.n_elim);
        flops += (size_t)L.n_elim * 2 * n3;
        P->levels.push_back(L);
        for (int q = 0; q < L.n_kept; q++) { const CrKept &k = kept[L.kept_off + q]; sL[k.node] = k.sLn; sU[k.node] = k.sUn; }
        alive = nxt;
    }
    P->n_levels = (int)P->levels.size();
    P->last_node = alive[0];
    flops += 2 * n3;
    const int g_base = next_block;                            // G^L_e = g_base + 2 e, G^U_e = g_base + 2 e + 1
    next_block += 2 * max_elim;
    P->pool_blocks = next_block;
    for (int l = 1; l < P->n_levels; l++) {                   // GEMM jobs of the dense levels
        CrLevel &L = P->levels[l];
        L.gjob_off = (int)jobs.size();
        for (int e = 0; e < L.n_elim; e++) {
            const CrElim &n = elim[L.elim_off + e];
            jobs.push_back(CrJob{nz + n.node, n.sL, -1, -1, g_base + 2 * e, 0, 1});
            jobs.push_back(CrJob{nz + n.node, n.sU, -1, -1, g_base + 2 * e + 1, 0, 1});
        }
        L.n_gjob = (int)jobs.size() - L.gjob_off;
        L.kjob_off = (int)jobs.size();
        auto epos = [&](int node) { for (int e = 0; e < L.n_elim; e++) if (elim[L.elim_off + e].node == node) return e; return -1; };
        for (int q = 0; q < L.n_kept; q++) {
            const CrKept &k = kept[L.kept_off + q];
            const int en = k.next >= 0 ? epos(k.next) : -1, ep = k.prev >= 0 ? epos(k.prev) : -1;
            if (en >= 0 && ep >= 0) jobs.push_back(CrJob{k.sU, g_base + 2 * en, k.sL, g_base + 2 * ep + 1, k.node, 1, -1});
            else if (en >= 0) jobs.push_back(CrJob{k.sU, g_base + 2 * en, -1, -1, k.node, 1, -1});
            else if (ep >= 0) jobs.push_back(CrJob{k.sL, g_base + 2 * ep + 1, -1, -1, k.node, 1, -1});
            if (en >= 0) jobs.push_back(CrJob{k.sU, g_base + 2 * en + 1, -1, -1, k.sUn, 0, -1});
            else jobs.push_back(CrJob{-1, -1, -1, -1, k.sUn, 0, 1});
            if (ep >= 0) jobs.push_back(CrJob{k.sL, g_base + 2 * ep, -1, -1, k.sLn, 0, -1});
            else jobs.push_back(CrJob{-1, -1, -1, -1, k.sLn, 0, 1});
        }
        L.n_kjob = (int)jobs.size() - L.kjob_off;
        for (int q = L.gjob_off; q < (int)jobs.size(); q++) flops += (size_t)((jobs[q].a1 >= 0) + (jobs[q].a2 >= 0)) * 2 * n3;
    }
    P->exec_flops = flops;
    std::vector<int> idx(elim.size());
    for (size_t q = 0; q < elim.size(); q++) idx[q] = elim[q].node;
    idx.push_back(P->last_node);
    cudaError_t e = cudaMalloc((void **)&P->d_elim, sizeof(CrElim) * std::max<size_t>(elim.size(), 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&P->d_kept, sizeof(CrKept) * std::max<size_t>(kept.size(), 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&P->d_jobs, sizeof(CrJob) * std::max<size_t>(jobs.size(), 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&P->d_idx, sizeof(int) * idx.size());
    if (e == cudaSuccess) e = cudaMalloc((void **)&P->pool, sizeof(double) * (size_t)P->pool_blocks * nip * nip);
    if (e == cudaSuccess) e = cudaMalloc((void **)&P->rv, sizeof(double) * (size_t)nz * nip);
    if (e == cudaSuccess) e = cudaMalloc((void **)&P->wv, sizeof(double) * (size_t)nz * nip);
    if (e == cudaSuccess) e = cudaMalloc((void **)&P->xv, sizeof(double) * (size_t)nz * nip);
    if (e == cudaSuccess) e = cudaMemcpy(P->d_elim, elim.data(), sizeof(CrElim) * elim.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(P->d_kept, kept.data(), sizeof(CrKept) * kept.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !jobs.empty()) e = cudaMemcpy(P->d_jobs, jobs.data(), sizeof(CrJob) * jobs.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(P->d_idx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cr_plan_free(P); return cuda_fail(e, "cyclic reduction plan"); }
    *out = P;
    return VK_OK;
}

// ---- level 0: the couplings are diagonal, so the Schur update of a kept node is an elementwise scaling of its neighbours' inverses
//      D_a -= up_a W_next dn_next + dn_a W_prev up_prev ;  U'_a = -up_a W_next up_next ;  L'_a = -dn_a W_prev dn_prev     (rows x cols)
__global__ void __launch_bounds__(256) cr_level0_kernel(int nip, const CrKept *kept, const double *D, const double *up, const double *dn,
                                                        double *pool, int nz)
{
    const CrKept k = kept[blockIdx.x];
    const size_t bs = (size_t)nip * nip;
    const double *Wn = k.next >= 0 ? pool + (size_t)(nz + k.next) * bs : nullptr, *Wp = k.prev >= 0 ? pool + (size_t)(nz + k.prev) * bs : nullptr;
    const double *upa = up + (size_t)k.node * nip, *dna = dn + (size_t)k.node * nip;
    const double *upn = k.next >= 0 ? up + (size_t)k.next * nip : nullptr, *dnn = k.next >= 0 ? dn + (size_t)k.next * nip : nullptr;
    const double *upp = k.prev >= 0 ? up + (size_t)k.prev * nip : nullptr, *dnp = k.prev >= 0 ? dn + (size_t)k.prev * nip : nullptr;
    double *Da = pool + (size_t)k.node * bs, *Un = pool + (size_t)k.sUn * bs, *Ln = pool + (size_t)k.sLn * bs;
    const double *D0 = D + (size_t)k.node * bs;
    for (int q = threadIdx.x; q < nip * nip; q += blockDim.x) {
        const int r = q / nip, c = q % nip;
        double d = D0[q], u = 0.0, l = 0.0;
        if (Wn) { const double w = upa[r] * Wn[q]; d -= w * dnn[c]; u = -(w * upn[c]); }
        if (Wp) { const double w = dna[r] * Wp[q]; d -= w * upp[c]; l = -(w * dnp[c]); }
        Da[q] = d; Un[q] = u; Ln[q] = l;
    }
}

// ---- dense levels: one block per job,  C = beta C + sign (A1 B1 + A2 B2)  on the FP64 tensor pipe.  Warp w owns the 8 columns [8w, 8w+8) of
// C in the DMMA accumulator layout (lane 4g+t: rows 8i+g, columns 8w+2t, 8w+2t+1); A and B are staged in shared memory with padded row
// stride (conflict-free fragment loads): a_frag = A[8i+g][4s+t], b_frag = B[4s+t][8w+g] for the k4 step s.
template <int NIP> struct CrGemmLd { static constexpr int LD = (NIP <= 96) ? NIP + 2 : NIP; };   // NIP = 120: two padded operands exceed 227 KB
template <int NIP>
__global__ void __launch_bounds__((NIP / 8) * 32) cr_gemm_kernel(const CrJob *jobs, double *pool)
{
    constexpr int NW = NIP / 8, NR = NIP / 8, LD = CrGemmLd<NIP>::LD;
    extern __shared__ __align__(16) double sm[];
    double *As = sm, *Bs = sm + (size_t)NIP * LD;
    const CrJob jb = jobs[blockIdx.x];
    const size_t bs = (size_t)NIP * NIP;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    double *C = pool + (size_t)jb.c * bs;
    double acc[NR][2];
#pragma unroll
    for (int i = 0; i < NR; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    for (int prod = 0; prod < 2; prod++) {
        const int ia = prod ? jb.a2 : jb.a1, ib = prod ? jb.b2 : jb.b1;
        if (ia < 0) break;
        const double *A = pool + (size_t)ia * bs, *B = pool + (size_t)ib * bs;
        __syncthreads();
        for (int q = tid; q < NIP * NIP / 2; q += NW * 32) {
            const int r = (2 * q) / NIP, c = (2 * q) % NIP;
            *reinterpret_cast<double2 *>(As + r * LD + c) = *reinterpret_cast<const double2 *>(A + 2 * (size_t)q);
            *reinterpret_cast<double2 *>(Bs + r * LD + c) = *reinterpret_cast<const double2 *>(B + 2 * (size_t)q);
        }
        __syncthreads();
#pragma unroll 2
        for (int s = 0; s < NIP / 4; s++) {
            const double b = Bs[(4 * s + t) * LD + 8 * w + g];
#pragma unroll
            for (int i = 0; i < NR; i++) dmma(acc[i][0], acc[i][1], As[(8 * i + g) * LD + 4 * s + t], b, acc[i][0], acc[i][1]);
        }
    }
    const double sg = (double)jb.sign;
#pragma unroll
    for (int i = 0; i < NR; i++) {
        double2 *cp = reinterpret_cast<double2 *>(C + (size_t)(8 * i + g) * NIP + 8 * w + 2 * t);
        double2 v = make_double2(sg * acc[i][0], sg * acc[i][1]);
        if (jb.beta) { const double2 c0 = *cp; v.x += c0.x; v.y += c0.y; }
        *cp = v;
    }
}

// ---- solve kernels (one thread per row, F_p staged in shared memory)
struct CrSolveArgs {
    int nz, nip, level, n;
    const CrElim *elim; const CrKept *kept;
    const double *F, *pool, *up, *dn;
    double *rv, *wv, *xv;
    const int *act;      // [1] or NULL: the column is skipped when act[0] == 0 (refine = auto below the step-size threshold, stopped column)
};

// forward, part 1:  w_p = F_p^{-1} r_p  for the eliminated nodes of the level (node = -1 in `single`: the last node, x = F^{-1} r)
template <int NIP>
__global__ void __launch_bounds__(((NIP + 31) / 32) * 32) cr_apply_kernel(CrSolveArgs a, int single)
{
    constexpr int LDF = NIP + 2;
    extern __shared__ __align__(16) double sm[];
    double *Fb = sm, *yv = sm + (size_t)NIP * LDF, *zv = yv + NIP;
    const int i = threadIdx.x;
    if (a.act && !a.act[0]) return;
    const int node = single >= 0 ? single : a.elim[blockIdx.x].node;
    const double *Fg = a.F + (size_t)node * NIP * LDF;
    for (int q = i; q < NIP * LDF / 2; q += blockDim.x) reinterpret_cast<double2 *>(Fb)[q] = reinterpret_cast<const double2 *>(Fg)[q];
    const double t = (i < NIP) ? a.rv[(size_t)node * NIP + i] : 0.0;
    __syncthreads();
    const double z = lu_apply_dev<NIP>(Fb, t, yv, zv, i);
    if (i < NIP) (single >= 0 ? a.xv : a.wv)[(size_t)node * NIP + i] = z;
}

// forward, part 2:  r_a -= U_a w_next + L_a w_prev  for the kept nodes (level 0: diagonal couplings up_a, dn_a)
template <int NIP>
__global__ void __launch_bounds__(((NIP + 31) / 32) * 32) cr_reduce_kernel(CrSolveArgs a)
{
    __shared__ double wn[NIP], wp[NIP];
    const int i = threadIdx.x;
    if (a.act && !a.act[0]) return;
    const CrKept k = a.kept[blockIdx.x];
    if (i < NIP) {
        wn[i] = k.next >= 0 ? a.wv[(size_t)k.next * NIP + i] : 0.0;
        wp[i] = k.prev >= 0 ? a.wv[(size_t)k.prev * NIP + i] : 0.0;
    }
    __syncthreads();
    if (i >= NIP) return;
    double acc = 0.0;
    if (k.sU < 0) {
        acc = a.up[(size_t)k.node * NIP + i] * wn[i] + a.dn[(size_t)k.node * NIP + i] * wp[i];
    } else {
        const double *U = a.pool + ((size_t)k.sU * NIP + i) * NIP, *L = a.pool + ((size_t)k.sL * NIP + i) * NIP;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        if (k.next >= 0)
            for (int c = 0; c < NIP; c += 2) { const double2 u = *reinterpret_cast<const double2 *>(U + c); a0 = fma(u.x, wn[c], a0); a1 = fma(u.y, wn[c + 1], a1); }
        if (k.prev >= 0)
            for (int c = 0; c < NIP; c += 2) { const double2 l = *reinterpret_cast<const double2 *>(L + c); b0 = fma(l.x, wp[c], b0); b1 = fma(l.y, wp[c + 1], b1); }
        acc = (a0 + a1) + (b0 + b1);
    }
    a.rv[(size_t)k.node * NIP + i] -= acc;
}

// backward:  x_p = w_p - F_p^{-1} (L_p x_left + U_p x_right)
template <int NIP>
__global__ void __launch_bounds__(((NIP + 31) / 32) * 32) cr_back_kernel(CrSolveArgs a)
{
    constexpr int LDF = NIP + 2;
    extern __shared__ __align__(16) double sm[];
    double *Fb = sm, *yv = sm + (size_t)NIP * LDF, *zv = yv + NIP, *xl = zv + NIP, *xr = xl + NIP;
    const int i = threadIdx.x;
    if (a.act && !a.act[0]) return;
    const CrElim e = a.elim[blockIdx.x];
    const double *Fg = a.F + (size_t)e.node * NIP * LDF;
    for (int q = i; q < NIP * LDF / 2; q += blockDim.x) reinterpret_cast<double2 *>(Fb)[q] = reinterpret_cast<const double2 *>(Fg)[q];
    if (i < NIP) {
        xl[i] = a.xv[(size_t)e.left * NIP + i];
        xr[i] = e.right >= 0 ? a.xv[(size_t)e.right * NIP + i] : 0.0;
    }
    __syncthreads();
    double t = 0.0;
    if (i < NIP) {
        if (e.sU < 0) {
            t = a.dn[(size_t)e.node * NIP + i] * xl[i] + a.up[(size_t)e.node * NIP + i] * xr[i];
        } else {
            const double *U = a.pool + ((size_t)e.sU * NIP + i) * NIP, *L = a.pool + ((size_t)e.sL * NIP + i) * NIP;
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            for (int c = 0; c < NIP; c += 2) { const double2 l = *reinterpret_cast<const double2 *>(L + c); b0 = fma(l.x, xl[c], b0); b1 = fma(l.y, xl[c + 1], b1); }
            if (e.right >= 0)
                for (int c = 0; c < NIP; c += 2) { const double2 u = *reinterpret_cast<const double2 *>(U + c); a0 = fma(u.x, xr[c], a0); a1 = fma(u.y, xr[c + 1], a1); }
            t = (a0 + a1) + (b0 + b1);
        }
    }
    const double z = lu_apply_dev<NIP>(Fb, t, yv, zv, i);
    if (i < NIP) a.xv[(size_t)e.node * NIP + i] = a.wv[(size_t)e.node * NIP + i] - z;
}

__global__ void cr_pad_rhs_kernel(int nz, int ni, int nip, const double *rhs, double *rv, const int *act)
{
    if (act && !act[0]) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nz * nip) { const int j = q / nip, i = q % nip; rv[q] = i < ni ? rhs[(size_t)j * ni + i] : 0.0; }
}
__global__ void cr_unpad_kernel(int nz, int ni, int nip, const double *xv, double *x, const int *act)
{
    if (act && !act[0]) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nz * ni) { const int j = q / ni, i = q % ni; x[q] = xv[(size_t)j * nip + i]; }
}

template <int NIP, int MINB>
static int cr_factor_t(vk_column *c, CrPlan *P, const double *D, const double *up, const double *dn, double *F, int *status)
{
    using C = FactorCfg<NIP>;
    const size_t bs = (size_t)NIP * NIP;
    { int rc = ensure_smem((const void *)factor_kernel<NIP, MINB, false, NoProducer>, c->net->device, C::SMEM); if (rc) return rc; }
    const size_t gsm = sizeof(double) * 2 * (size_t)NIP * CrGemmLd<NIP>::LD;
    { int rc = ensure_smem((const void *)cr_gemm_kernel<NIP>, c->net->device, gsm); if (rc) return rc; }
    // current diagonal blocks <- D (the original blocks stay in c->D for the refinement residual)
    VK_CUDA(cudaMemcpyAsync(P->pool, D, sizeof(double) * bs * P->nz, cudaMemcpyDeviceToDevice, c->stream));
    int idx_off = 0;
    for (int l = 0; l < P->n_levels; l++) {
        const CrLevel &L = P->levels[l];
        FactorArgs fa{1, c->ni, P->pool, nullptr, nullptr, F, status, nullptr, P->d_idx + idx_off, P->pool + bs * P->nz, 0};
        factor_kernel<NIP, MINB, false, NoProducer><<<L.n_elim, C::NT, C::SMEM, c->stream>>>(fa, NoProducer{});
        idx_off += L.n_elim;
        if (l == 0) {
            cr_level0_kernel<<<L.n_kept, 256, 0, c->stream>>>(NIP, P->d_kept + L.kept_off, D, up, dn, P->pool, P->nz);
        } else {
            cr_gemm_kernel<NIP><<<L.n_gjob, (NIP / 8) * 32, gsm, c->stream>>>(P->d_jobs + L.gjob_off, P->pool);
            cr_gemm_kernel<NIP><<<L.n_kjob, (NIP / 8) * 32, gsm, c->stream>>>(P->d_jobs + L.kjob_off, P->pool);
        }
    }
    FactorArgs fa{1, c->ni, P->pool, nullptr, nullptr, F, status, nullptr, P->d_idx + idx_off, P->pool + bs * P->nz, 0};
    factor_kernel<NIP, MINB, false, NoProducer><<<1, C::NT, C::SMEM, c->stream>>>(fa, NoProducer{});
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

template <int NIP>
static int cr_solve_t(vk_column *c, CrPlan *P, const double *F, const double *up, const double *dn, const double *rhs, double *x, const int *act)
{
    constexpr int NT = ((NIP + 31) / 32) * 32;
    const size_t asm_ = sizeof(double) * ((size_t)NIP * (NIP + 2) + 4 * NIP);
    { int rc = ensure_smem((const void *)cr_apply_kernel<NIP>, c->net->device, asm_); if (rc) return rc; }
    { int rc = ensure_smem((const void *)cr_back_kernel<NIP>, c->net->device, asm_); if (rc) return rc; }
    const int nz = P->nz;
    cr_pad_rhs_kernel<<<(nz * NIP + 255) / 256, 256, 0, c->stream>>>(nz, c->ni, NIP, rhs, P->rv, act);
    CrSolveArgs a{nz, NIP, 0, 0, nullptr, nullptr, F, P->pool, up, dn, P->rv, P->wv, P->xv, act};
    for (int l = 0; l < P->n_levels; l++) {
        const CrLevel &L = P->levels[l];
        a.level = l; a.elim = P->d_elim + L.elim_off; a.kept = P->d_kept + L.kept_off;
        cr_apply_kernel<NIP><<<L.n_elim, NT, asm_, c->stream>>>(a, -1);
        cr_reduce_kernel<NIP><<<L.n_kept, NT, 0, c->stream>>>(a);
    }
    cr_apply_kernel<NIP><<<1, NT, asm_, c->stream>>>(a, P->last_node);
    for (int l = P->n_levels - 1; l >= 0; l--) {
        const CrLevel &L = P->levels[l];
        a.level = l; a.elim = P->d_elim + L.elim_off; a.kept = P->d_kept + L.kept_off;
        cr_back_kernel<NIP><<<L.n_elim, NT, asm_, c->stream>>>(a);
    }
    cr_unpad_kernel<<<(nz * c->ni + 255) / 256, 256, 0, c->stream>>>(nz, c->ni, NIP, P->xv, x, act);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

// one column only (c->ncol == 1): the latency path
int launch_cr_factor(vk_column *c, const double *D, const double *up, const double *dn, double *F, int *status)
{
    if (!c->cr || c->cr->nz != c->nz) {
        cr_plan_free(c->cr);
        c->cr = nullptr;
        int rc = cr_plan_build(c->nz, c->nip, &c->cr);
        if (rc) return rc;
    }
    switch (c->nip) {
        case 48: return cr_factor_t<48, 2>(c, c->cr, D, up, dn, F, status);
        case 72: return cr_factor_t<72, 2>(c, c->cr, D, up, dn, F, status);
        case 96: return cr_factor_t<96, 1>(c, c->cr, D, up, dn, F, status);
        case 120: return cr_factor_t<120, 1>(c, c->cr, D, up, dn, F, status);
        default: set_error("no factor kernel for this padded block size"); return VK_ERR_UNSUPPORTED;
    }
}

int launch_cr_solve(vk_column *c, const double *F, const double *up, const double *dn, const double *rhs, double *x, const int *act)
{
    switch (c->nip) {
        case 48: return cr_solve_t<48>(c, c->cr, F, up, dn, rhs, x, act);
        case 72: return cr_solve_t<72>(c, c->cr, F, up, dn, rhs, x, act);
        case 96: return cr_solve_t<96>(c, c->cr, F, up, dn, rhs, x, act);
        case 120: return cr_solve_t<120>(c, c->cr, F, up, dn, rhs, x, act);
        default: set_error("no solve kernel for this padded block size"); return VK_ERR_UNSUPPORTED;
    }
}

}  // namespace vk
