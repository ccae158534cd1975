// Device pieces of the block-tridiagonal factorisation (see the header of vk_solve.cu): the blocked Gauss-Jordan / block-LU kernel as a
// template over how the next layer's block D_{j+1} reaches shared memory -
//   FUSED = false : D, up, dn were written to HBM by lhs_ml_kernel and are prefetched by 1-D TMA bulk copies (vk_solve.cu, vk_blocktri_solve);
//   FUSED = true  : the otherwise idle warps of the block ASSEMBLE D_{j+1} = 1/(r h) I - J_chem - J_transport in shared memory from y, k and
//                   the network tables while the column warps factor layer j (vk_chem.cu: lhs_produce): the 6.2 MB per column of D are
//                   neither written to nor read from HBM, and the separate Jacobian kernel disappears from the step.
#pragma once
#include <type_traits>

#include "vk_internal.cuh"

namespace vk {

// 1/x without the IEEE slow path: hardware approximation (2^-23) + two Newton steps (error ~1 ulp); x = 0 / inf / nan give
// inf / 0 / nan, caught by the singular-pivot flag
__device__ __forceinline__ double fast_rcp(double x)
{
#ifdef VK_EXACT_RCP
    return 1.0 / x;
#endif
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// Register layout: column warp w (< NW) owns the 8 columns [8w, 8w+8) of the block and ALL its rows, in the m8n8
// accumulator-fragment pattern stacked vertically: lane = 4*g + t holds rows 8*i + g (i < NIP/8) and columns 8w + 2t, 8w + 2t + 1
// - the C/D layout of mma.sync.m8n8k4.f64 (DMMA), so a tile of the matrix is directly a DMMA accumulator.
//
// Algorithm: BLOCKED in-place Gauss-Jordan inversion with 8-wide panels on the FP64 tensor pipe.  For panel K (rows and
// columns 8kt .. 8kt+7), with P = A_KK^{-1}:
//     columns outside the panel (warp w != kt):   V = P A_Kw ;  A_Kw <- V ;  A_iw <- A_iw - A_iK V     (i != kt)
//     panel columns (warp kt):                    A_iK <- -A_iK P  (i != kt) ;  A_KK <- P
// i.e. per panel and warp 2 + 2(NR-1) DMMAs instead of 8 x 2 NR DFMAs, one barrier per 8 pivots instead of 8, and NR 16-byte
// shared-memory loads instead of 8 NR.  Only two things cross warps: the RAW panel columns A_iK and the 8 x 8 inverse P.
// The k index of the m8n8k4 shape is mapped slot t <-> panel index 2t+s (s = which of the two k4 steps), so that the A
// operand of every product is exactly the (x, y) pair a lane already holds / loads with one LDS.128.
//
// Warp specialisation: the 8 dependent pivot steps of the 8 x 8 inverse (shuffle -> 1/x -> multiply -> FMA, ~115 clk each) are
// the critical chain of a panel.  A warp that also carries DMMAs issues in order and is held up whenever the tensor pipe of
// its sub-partition is busy with the other warps' updates, so the inverse runs on a DEDICATED warp (warp NW) that owns no
// columns: warp kt+1 updates its diagonal tile first, hands it over through shared memory (named barrier, 64 threads), goes on
// with its other tiles and publishes its raw columns; the inverse warp publishes P and ARRIVES on the panel barrier that the
// column warps wait on.  The next layer's D block and couplings are prefetched into shared memory by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier) issued one panel into the current layer.
//
// Pivoting: DIAGONAL pivots inside the 8 x 8 panel inverse.  Measured on the reference's own matrices (tests + DESIGN.md
// §4.1): for these systems (1/(r h) I - J with the loss terms on the diagonal; abundances spanning 30+ decades so that rows
// carry wildly different scales) elimination on the diagonal is 2-5 orders of magnitude MORE accurate than LAPACK-style
// partial pivoting, which lets the largest entry of a column - a scale artefact - destroy componentwise accuracy.  A zero /
// non-finite pivot sets VK_ERR_SINGULAR for the column: the step is rejected and retried with dt/2 like any failed step.
template <int NIP>
struct FactorCfg {
    static constexpr int NW = NIP / 8;           // column warps
    static constexpr int NR = NIP / 8;           // row tiles per lane
    // Warp -> sub-partition is warp id % 4 (scripts/ubench/dmma_warps.cu).  DFMA/DMUL of the panel-inverse chain share the FP64
    // pipe with DMMA, so the inverse warp gets a sub-partition of its own: warp 3 inverts, warps 7, 11, .. only keep the block
    // barriers company, and the column warps fill sub-partitions 0-2 evenly (NW is a multiple of 3).
    static constexpr bool SPREAD = (NW % 3 == 0) && (NW / 3) * 4 * 32 <= 512;
    static constexpr int NWARPS = SPREAD ? (NW / 3) * 4 : NW + 1;
    static constexpr int HELPER = SPREAD ? 3 : NW;
    static constexpr int NT = NWARPS * 32;
    static constexpr int NSYNC = (NW + 1) * 32;  // threads on the panel barrier: column warps + the inverse warp
    // doubles: dbuf[NIP*NIP] + updn[2*NIP] + mraw[2][NIP][8] + pbuf[3][64] + hbuf[2][2][64] + hraw[2][64] + mbarrier
    static constexpr size_t SMEM_DOUBLES = (size_t)NIP * NIP + 2 * NIP + 16 * NIP + 192 + 256 + 128 + 2;
    static constexpr size_t SMEM = sizeof(double) * SMEM_DOUBLES;
    static constexpr unsigned TX_BYTES = (unsigned)(sizeof(double) * ((size_t)NIP * NIP + 2 * NIP));
    // fused assembly (FUSED = true): the idle warps of the spread layout are the producers of D_{j+1}
    static constexpr int NPROD = SPREAD ? NW / 3 - 1 : 0;            // producer warps: 1 / 2 / 3 / 4 for NIP = 48 / 72 / 96 / 120
    static constexpr int NFEED = (NW + NPROD) * 32;                  // threads on the dbuf hand-over barriers
    // fused forward elimination: tvec[NIP] + zpart[NW][NIP] behind the buffers above
    static constexpr size_t FWD_OFF = (SMEM_DOUBLES + 1) & ~(size_t)1;
    static constexpr size_t SMEM_FWD = sizeof(double) * (FWD_OFF + (size_t)NIP + (size_t)NW * NIP);
};
struct NoProducer {};
// producer side of the fused kernel (defined in vk_lhs_dev.cuh, instantiated only by vk_chem.cu)
template <int NIP, class PA>
__device__ void lhs_produce(const PA &pa, int col, int nz, double *dbuf, double *updn, double *extra, int pw, int lane);

struct FactorArgs {
    int nz, ni;
    const double *D;     // [ncol][nz][NIP][NIP]
    const double *up;    // [ncol][nz][NIP]
    const double *dn;
    double *F;           // [ncol][nz][NIP][NIP+2] block LU factors of S_j for the solve sweeps (row stride NIP+2: conflict-free LDS.128)
    int *status;         // [ncol]
    const int *act;      // [ncol] or NULL: stopped columns are skipped
    // block-list mode (cyclic reduction, vk_cr.inl): nz = 1 and block b of the grid inverts / factors the block number blk_idx[b] of the
    // pool D; its explicit inverse goes to Wout (same indexing as D), a singular block is reported in status[status_idx]
    const int *blk_idx;
    double *Wout;
    int status_idx;
    // forward elimination of stage 1 fused into the sweep (optional, rhs != NULL): for columns with dt < fwd_dt_max the kernel forms
    // z_j = W_j (r_j - dn_j z_{j-1}) from the explicit inverse it holds in registers and sets fwd_done[col] = 1, so that the first
    // solve only runs its backward sweep (one read of F less).  x = fl(S^-1) t is not backward stable: it is used only
    // below the step size from which the element budget needs the block LU solve (and refinement), fwd_dt_max = refine_dt_min
    const double *rhs;   // [ncol][nz][ni]
    double *z;           // [ncol][nz][NIP]
    const double *dt;    // [ncol]
    double fwd_dt_max;
    int *fwd_done;       // [ncol]
};

// D(8x8) = A(8x4) B(4x8) + C on the FP64 tensor pipe: lane 4g+t supplies A[g][t], B[t][g], C[g][2t..2t+1]
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
        : "=d"(d0), "=d"(d1)
        : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// tile in accumulator layout (x0 = T[g][2t], x1 = T[g][2t+1])  ->  B fragments of the two k4 steps, b_s = T[2t+s][g]
__device__ __forceinline__ void to_bfrag(double x0, double x1, int g, int t, double &b0, double &b1)
{
    const int s0 = (t << 3) | (g >> 1), s1 = s0 + 4;      // lanes (2t, g>>1) and (2t+1, g>>1)
    const double p0 = __shfl_sync(0xffffffffu, x0, s0), p1 = __shfl_sync(0xffffffffu, x1, s0);
    const double q0 = __shfl_sync(0xffffffffu, x0, s1), q1 = __shfl_sync(0xffffffffu, x1, s1);
    b0 = (g & 1) ? p1 : p0;
    b1 = (g & 1) ? q1 : q0;
}

// one pivot step P (compile time) of the in-warp 8 x 8 Gauss-Jordan inverse on an accumulator-layout tile
template <int P>
__device__ __forceinline__ void gj8_step(double &x0, double &x1, int g, int t, int &bad)
{
    constexpr int HP = P >> 1, E = P & 1;
    const double mine = E ? x1 : x0;
    const double piv = __shfl_sync(0xffffffffu, mine, (P << 2) | HP);     // a[P][P]
    const double cp = __shfl_sync(0xffffffffu, mine, (g << 2) | HP);      // a[g][P]
    const double r0 = __shfl_sync(0xffffffffu, x0, (P << 2) | t);         // a[P][2t]
    const double r1 = __shfl_sync(0xffffffffu, x1, (P << 2) | t);         // a[P][2t+1]
    const double rinv = fast_rcp(piv);
    if (!(fabs(piv) > 0.0) || !(fabs(piv) < 1.0e300)) bad = 1;
    if (g == P) {
        x0 = r0 * rinv;
        x1 = r1 * rinv;
        if (t == HP) { if (E) x1 = rinv; else x0 = rinv; }
    } else {
        const double m = -cp * rinv;
        x0 = fma(m, r0, x0);
        x1 = fma(m, r1, x1);
        if (t == HP) { if (E) x1 = m; else x0 = m; }
    }
}

// named barriers (id 0 is __syncthreads): producer/consumer hand-offs between the column warps and the inverse warp
template <int ID, int NTHREADS>
__device__ __forceinline__ void bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory"); }
template <int ID, int NTHREADS>
__device__ __forceinline__ void bar_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory"); }

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VK_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra VK_DONE_%=;\n"
        "bra VK_WAIT_%=;\n"
        "VK_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (16-byte aligned, size multiple of 16)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

enum { VK_BAR_PANEL = 1, VK_BAR_TILE = 3, VK_BAR_RAW = 5, VK_BAR_COLS = 7,     // + panel parity (not COLS)
       VK_BAR_FREE = 8, VK_BAR_FULL = 9, VK_BAR_PROD = 10 };                  // fused assembly: dbuf consumed / dbuf filled / producers only

template <int NIP, int MINB, bool FUSED, class PA>
__global__ void __launch_bounds__(FactorCfg<NIP>::NT, MINB) factor_kernel(FactorArgs a, PA pa)
{
    using C = FactorCfg<NIP>;
    constexpr int NR = C::NR, NW = C::NW, NT = C::NSYNC;
    extern __shared__ __align__(128) double smem[];
    double *dbuf = smem;                 // NIP x NIP     D_j (TMA destination)
    double *updn = dbuf + NIP * NIP;     // 2 x NIP       up_{j-1}, dn_j (TMA destination)
    double *mraw = updn + 2 * NIP;       // 2 x NIP x 8   raw panel columns A_iK (double buffered by panel parity)
    double *pbuf = mraw + 16 * NIP;      // 3 x 64        P = A_KK^{-1} (the inverse warp runs up to two panels ahead)
    double *hbuf = pbuf + 192;           // 2 x 2 x 64    tiles handed to the inverse warp: [parity][0] = A_{K-1,K}, [parity][1] = A_KK
    double *hraw = hbuf + 256;           // 2 x 64        [parity of m] = A_mK, K = panel m-1 (from warp m-1)
    void *mbar = hraw + 128;             // mbarrier of the TMA prefetch

    const int col = a.blk_idx ? a.blk_idx[blockIdx.x] : blockIdx.x;
    if (a.act && !a.blk_idx && !a.act[col]) return;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int w = C::SPREAD ? wid - (wid >> 2) : wid;     // column-warp index (meaningless for the other warps)
    const int tid = w * 32 + lane;                         // thread index among the column warps
    const int nz = a.nz;
    const size_t cbase = (size_t)col * nz;
    int bad = 0;
    bool fuse = false;
    if constexpr (!FUSED) fuse = a.rhs != nullptr && a.blk_idx == nullptr && a.dt[col] < a.fwd_dt_max;
    double *tvec = smem + C::FWD_OFF;    // NIP        t_j = r_j - dn_j z_{j-1}
    double *zpart = tvec + NIP;          // NW x NIP   per column warp: its columns' share of W_j t_j

    auto prefetch = [&](int j) {         // one thread: D_j, up_{j-1}, dn_j -> shared memory
        mbar_expect_tx(mbar, (j > 0) ? C::TX_BYTES : (unsigned)(sizeof(double) * NIP * NIP));
        tma_load_1d(dbuf, a.D + (cbase + j) * NIP * NIP, (unsigned)(sizeof(double) * NIP * NIP), mbar);
        if (j > 0) {
            tma_load_1d(updn, a.up + (cbase + j - 1) * NIP, (unsigned)(sizeof(double) * NIP), mbar);
            tma_load_1d(updn + NIP, a.dn + (cbase + j) * NIP, (unsigned)(sizeof(double) * NIP), mbar);
        }
    };
    if (!FUSED && threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (!FUSED && threadIdx.x == 0) prefetch(0);

    if (C::SPREAD && (wid & 3) == 3 && wid != C::HELPER) {
        if constexpr (FUSED) {
            // producer warps: assemble D_{j+1}, up_j, dn_{j+1} in dbuf / updn while the column warps factor layer j (vk_lhs_dev.cuh)
            lhs_produce<NIP, PA>(pa, col, nz, dbuf, updn, smem + ((C::SMEM_DOUBLES + 1) & ~(size_t)1), (wid >> 2) - 1, lane);
        } else {
            // idle warps of the spread layout: only the block-wide barrier of every layer
            for (int j = 0; j < nz; j++) {
                if (__syncthreads_or(0)) return;
            }
        }
        return;
    }
    if (wid == C::HELPER) {
        // ================= panel-inverse warp =================
        // P_m = (A_mm - A_mK P_{m-1} A_Km)^{-1} with K = panel m-1 and all three tiles as they are after the update of panel m-2:
        // the whole chain P_{m-1} -> P_m stays inside this warp, the column warps only feed it tiles one panel ahead of time.
        for (int j = 0; j < nz; j++) {
            double x0 = 0.0, x1 = 0.0;       // P_{m-1} in accumulator layout = A fragments of the next product
            auto step = [&](auto parc, int m) {
                constexpr int par = decltype(parc)::value;
                bar_sync<VK_BAR_TILE + par, 64>();                     // tiles of panel m are in hbuf[par]
                const double *hb = hbuf + par * 128;
                const double2 dg = *reinterpret_cast<const double2 *>(hb + 64 + g * 8 + 2 * t);
                double s0 = dg.x, s1 = dg.y;
                if (m > 0) {
                    const double ub0 = hb[(2 * t) * 8 + g], ub1 = hb[(2 * t + 1) * 8 + g];
                    double ta0, ta1, tb0, tb1;
                    dmma(ta0, ta1, x0, ub0, 0.0, 0.0);
                    dmma(tb0, tb1, x1, ub1, 0.0, 0.0);
                    ta0 += tb0; ta1 += tb1;                            // T = P_{m-1} A_Km
                    double nb0, nb1;
                    to_bfrag(ta0, ta1, g, t, nb0, nb1);
                    bar_sync<VK_BAR_RAW + par, 64>();                  // A_mK (K = panel m-1), handed over by warp m-1
                    const double2 mr = *reinterpret_cast<const double2 *>(hraw + par * 64 + g * 8 + 2 * t);
                    double sa0, sa1, sb0, sb1;
                    dmma(sa0, sa1, mr.x, -nb0, s0, s1);
                    dmma(sb0, sb1, mr.y, -nb1, 0.0, 0.0);
                    s0 = sa0 + sb0; s1 = sa1 + sb1;
                }
                x0 = s0; x1 = s1;
                gj8_step<0>(x0, x1, g, t, bad); gj8_step<1>(x0, x1, g, t, bad); gj8_step<2>(x0, x1, g, t, bad);
                gj8_step<3>(x0, x1, g, t, bad); gj8_step<4>(x0, x1, g, t, bad); gj8_step<5>(x0, x1, g, t, bad);
                gj8_step<6>(x0, x1, g, t, bad); gj8_step<7>(x0, x1, g, t, bad);
                *reinterpret_cast<double2 *>(pbuf + (m % 3) * 64 + g * 8 + 2 * t) = make_double2(x0, x1);
                bar_arrive<VK_BAR_PANEL + par, NT>();                  // P_m published
            };
            for (int m = 0; m < NR; m += 2) {
                step(std::integral_constant<int, 0>{}, m);
                if (m + 1 < NR) step(std::integral_constant<int, 1>{}, m + 1);
            }
            if (__syncthreads_or(bad)) return;     // the one block-wide barrier of a layer (the second one is among the column warps)
        }
        return;
    }

    // ================= column warps =================
    const int c0 = 8 * w + 2 * t;        // my columns c0, c0+1 ; my rows 8*i + g
    double A[NR][2];
    auto publish_raw = [&](int buf) {
#pragma unroll
        for (int i = 0; i < NR; i++)
            *reinterpret_cast<double2 *>(mraw + ((size_t)buf * NIP + 8 * i + g) * 8 + 2 * t) = make_double2(A[i][0], A[i][1]);
    };

    // tiles the inverse warp needs for P_0 and P_1 of a layer (from warps 0 and 1), raw columns of panel 0
    auto publish_first = [&]() {
        if (w == 0) {
            *reinterpret_cast<double2 *>(hbuf + 64 + g * 8 + 2 * t) = make_double2(A[0][0], A[0][1]);
            bar_arrive<VK_BAR_TILE, 64>();
            *reinterpret_cast<double2 *>(hraw + 64 + g * 8 + 2 * t) = make_double2(A[1][0], A[1][1]);
            bar_arrive<VK_BAR_RAW + 1, 64>();
        } else if (w == 1) {
            *reinterpret_cast<double2 *>(hbuf + 128 + g * 8 + 2 * t) = make_double2(A[0][0], A[0][1]);
            *reinterpret_cast<double2 *>(hbuf + 128 + 64 + g * 8 + 2 * t) = make_double2(A[1][0], A[1][1]);
            bar_arrive<VK_BAR_TILE + 1, 64>();
        }
    };
    // ---- layer 0: S_0 = D_0
    {
        if constexpr (FUSED) bar_sync<VK_BAR_FULL, C::NFEED>(); else mbar_wait(mbar, 0);
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const double2 d = *reinterpret_cast<const double2 *>(dbuf + (8 * i + g) * NIP + c0);
            A[i][0] = d.x;
            A[i][1] = d.y;
        }
        publish_first();
        if (w == 0) publish_raw(0);
        if (a.fwd_done && tid == 0) a.fwd_done[col] = fuse ? 1 : 0;
        if (fuse && tid < NIP) tvec[tid] = (tid < a.ni) ? a.rhs[cbase * a.ni + tid] : 0.0;      // (ordered by the panel barriers of layer 0)
    }

    for (int j = 0; j < nz; j++) {
        double u0 = 0.0, u1 = 0.0;       // B fragments of my pivot-row tile of the coming panel
        if (w != 0) to_bfrag(A[0][0], A[0][1], g, t, u0, u1);
        auto panel = [&](auto ktc) {
            constexpr int kt = decltype(ktc)::value;
            constexpr int par = kt & 1;
            const double *pb = pbuf + (kt % 3) * 64;
            bar_sync<VK_BAR_PANEL + par, NT>();           // raw panel columns + P of panel kt visible
            if constexpr (FUSED) {
                if (kt == 1 && j + 1 < nz) bar_arrive<VK_BAR_FREE, C::NFEED>();   // every column warp has consumed dbuf / updn by now
            } else {
                if (kt == 1 && tid == 0 && j + 1 < nz) prefetch(j + 1);
            }
            if (w == kt) {
                // ---- panel columns: A_iK <- -A_iK P, A_KK <- P.  A operand = my own (x, y) pair, B operand = -P
                const double b0 = -pb[(2 * t) * 8 + g], b1 = -pb[(2 * t + 1) * 8 + g];
#pragma unroll
                for (int i = 0; i < NR; i++) {
                    if (i == kt) continue;
                    double d0, d1;
                    dmma(d0, d1, A[i][0], b0, 0.0, 0.0);
                    dmma(A[i][0], A[i][1], A[i][1], b1, d0, d1);
                }
                const double2 pc = *reinterpret_cast<const double2 *>(pb + g * 8 + 2 * t);
                A[kt][0] = pc.x; A[kt][1] = pc.y;
                if (a.F) {      // block LU by-product: P_K on the diagonal tile, Lt_iK = A_iK P_K (= minus the new panel column) below it
                    double *Ft = a.F + (cbase + j) * (size_t)(NIP * (NIP + 2)) + (size_t)g * (NIP + 2) + c0;
                    *reinterpret_cast<double2 *>(Ft + (size_t)(8 * kt) * (NIP + 2)) = make_double2(pc.x, pc.y);
#pragma unroll
                    for (int i = kt + 1; i < NR; i++)
                        *reinterpret_cast<double2 *>(Ft + (size_t)(8 * i) * (NIP + 2)) = make_double2(-A[i][0], -A[i][1]);
                }
                if constexpr (kt + 1 < NR) to_bfrag(A[kt + 1][0], A[kt + 1][1], g, t, u0, u1);
            } else {
                // ---- V = P A_Kw (new pivot rows of my columns)
                const double2 pa = *reinterpret_cast<const double2 *>(pb + g * 8 + 2 * t);
                double v0, v1;
                dmma(v0, v1, pa.x, u0, 0.0, 0.0);
                dmma(v0, v1, pa.y, u1, v0, v1);
                A[kt][0] = v0; A[kt][1] = v1;
                if (a.F && w > kt)   // block LU by-product: V_Kw = P_K A_Kw right of the diagonal tile
                    *reinterpret_cast<double2 *>(a.F + (cbase + j) * (size_t)(NIP * (NIP + 2)) + (size_t)(8 * kt + g) * (NIP + 2) + c0) =
                        make_double2(v0, v1);
                double nv0, nv1;
                to_bfrag(v0, v1, g, t, nv0, nv1);
                nv0 = -nv0; nv1 = -nv1;
                const double *mr = mraw + ((size_t)par * NIP + g) * 8 + 2 * t;
                auto upd = [&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    const double2 m = *reinterpret_cast<const double2 *>(mr + i * 64);
                    dmma(A[i][0], A[i][1], m.x, nv0, A[i][0], A[i][1]);
                    dmma(A[i][0], A[i][1], m.y, nv1, A[i][0], A[i][1]);
                };
                // the tiles the inverse warp is waiting for first: pivot rows of the next panel, diagonal tile of the one after
                if constexpr (kt + 1 < NR) {
                    upd(std::integral_constant<int, kt + 1>{});
                    if (w != kt + 1) to_bfrag(A[kt + 1][0], A[kt + 1][1], g, t, u0, u1);
                }
                if constexpr (kt + 2 < NR) {
                    upd(std::integral_constant<int, kt + 2>{});
                    if (w == kt + 1) {                    // A_{kt+2,K'} (K' = panel kt+1) for P_{kt+2}
                        constexpr int hq = (kt + 2) & 1;
                        *reinterpret_cast<double2 *>(hraw + hq * 64 + g * 8 + 2 * t) = make_double2(A[kt + 2][0], A[kt + 2][1]);
                        bar_arrive<VK_BAR_RAW + hq, 64>();
                    }
                    if (w == kt + 2) {
                        constexpr int hp = (kt + 2) & 1;
                        *reinterpret_cast<double2 *>(hbuf + hp * 128 + g * 8 + 2 * t) = make_double2(A[kt + 1][0], A[kt + 1][1]);
                        *reinterpret_cast<double2 *>(hbuf + hp * 128 + 64 + g * 8 + 2 * t) = make_double2(A[kt + 2][0], A[kt + 2][1]);
                        bar_arrive<VK_BAR_TILE + hp, 64>();
                    }
                }
#pragma unroll
                for (int i = 0; i < NR; i++) {
                    if (i == kt || i == kt + 1 || i == kt + 2) continue;
                    const double2 m = *reinterpret_cast<const double2 *>(mr + i * 64);
                    dmma(A[i][0], A[i][1], m.x, nv0, A[i][0], A[i][1]);
                    dmma(A[i][0], A[i][1], m.y, nv1, A[i][0], A[i][1]);
                }
                if constexpr (kt + 1 < NR) {
                    if (w == kt + 1) {
                        publish_raw((kt + 1) & 1);
                    }
                }
            }
        };
#define VK_PANEL(N) if constexpr ((N) < NR) panel(std::integral_constant<int, (N)>{});
        VK_PANEL(0) VK_PANEL(1) VK_PANEL(2) VK_PANEL(3) VK_PANEL(4) VK_PANEL(5) VK_PANEL(6) VK_PANEL(7)
        VK_PANEL(8) VK_PANEL(9) VK_PANEL(10) VK_PANEL(11) VK_PANEL(12) VK_PANEL(13) VK_PANEL(14)
#undef VK_PANEL
        if (__syncthreads_or(bad)) {
            if (tid == 0) a.status[a.blk_idx ? a.status_idx : col] = VK_ERR_SINGULAR;
            return;
        }
        if (a.Wout) {        // the explicit inverse W_j (block-list mode)
            double *Wg = a.Wout + (cbase + j) * (size_t)(NIP * NIP);
#pragma unroll
            for (int i = 0; i < NR; i++) *reinterpret_cast<double2 *>(Wg + (size_t)(8 * i + g) * NIP + c0) = make_double2(A[i][0], A[i][1]);
        }
        // ---- A now holds W_j = S_j^{-1} (it stays in the registers; the block LU factors of S_j went out tile by tile above): Schur
        // update of the NEXT layer, S_{j+1} = D_{j+1} - diag(dn_{j+1}) W_j diag(up_j) (D, up, dn already in shared memory by TMA).  Warps 0
        // and 1 hand their first two new tiles to the inverse warp as soon as they exist, so the P_0 chain of layer j+1 starts at once.
        const bool more = j + 1 < nz;
        double rnext = 0.0, tv0 = 0.0, tv1 = 0.0;
        if (fuse) {
            if (more && tid < a.ni) rnext = a.rhs[(cbase + j + 1) * a.ni + tid];
            tv0 = tvec[c0]; tv1 = tvec[c0 + 1];
        }
        double su0 = 0.0, su1 = 0.0;
        if (more) {
            if constexpr (FUSED) bar_sync<VK_BAR_FULL, C::NFEED>(); else mbar_wait(mbar, (j + 1) & 1);
            su0 = updn[c0]; su1 = updn[c0 + 1];
        }
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const int r = 8 * i + g;
            if (fuse) {
                double part = fma(A[i][0], tv0, A[i][1] * tv1);
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                part += __shfl_xor_sync(0xffffffffu, part, 2);
                if (t == 0) zpart[w * NIP + r] = part;
            }
            if (more) {
                const double2 d = *reinterpret_cast<const double2 *>(dbuf + r * NIP + c0);
                const double l = updn[NIP + r];
                A[i][0] = d.x - (l * A[i][0]) * su0;
                A[i][1] = d.y - (l * A[i][1]) * su1;
                if (i == 1) publish_first();
            }
        }
        if (more && w == 0) publish_raw(0);
        bar_sync<VK_BAR_COLS, NW * 32>();
        if (fuse && tid < NIP) {   // z_j = W_j (r_j - dn_j z_{j-1}); right-hand side of the next layer's elimination
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < NW; q++) acc += zpart[q * NIP + tid];
            a.z[(cbase + j) * NIP + tid] = acc;
            if (more) tvec[tid] = rnext - updn[NIP + tid] * acc;
        }
        // (the panel barriers of the next layer order the reuse of zpart / tvec, and the prefetch of layer j+2 that overwrites updn)
    }
}

}  // namespace vk
