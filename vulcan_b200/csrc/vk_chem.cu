// Chemistry + transport kernels: one thread block per (column, layer).
//
//   rhs_warp_kernel : chem_funs.chemdf (make_chem_funs.py:113-430) + ODESolver.diffdf / diffdf_settling / diffdf_no_mol
//                (op.py:1438-1791), optionally the Ros2 stage-2 right-hand side  f(y+k1/r) - 2/(r h) k1 (op.py:2917-2928)
//   lhs_ml_kernel : chem_funs.neg_symjac block (make_chem_funs.py:653-717) + lhs_jac_tot / _settling / _no_mol
//                (op.py:1973-2364):  D = 1/(r h) I - J_chem - J_transport, up/dn = the diagonal couplings
//   atm_pre_kernel : the parts of the transport stencil that depend on the atmosphere only (molecular-diffusion quotients,
//                thermal/gravity brackets, settling terms) evaluated ONCE per vk_set_atm with the reference's operation
//                order, so that the per-step kernels reproduce the reference bit for bit without re-doing ~25 divisions
//                per species and layer.
//
// This translation unit is compiled with -fmad=false: every expression is evaluated with one IEEE rounding per operation in
// the reference's order, so chemdf and diffdf are bit-identical to the numpy reference (tests/test_gpu_parity.py).
#include <cstdio>
#include <cstdlib>

#include "vk_internal.cuh"
#include "vk_device_math.cuh"
#include "vk_factor_dev.cuh"

namespace vk {

struct AtmLayer {  // per-column base pointers
    const double *Kzz, *vz, *dzi, *Dzz, *vs, *Tco, *g, *Ti, *Hpi, *ms, *alpha, *top_flux, *bot_flux, *bot_vdep;
};
__device__ __forceinline__ AtmLayer atm_at(const AtmDev &a, int col)
{
    AtmLayer L;
    L.Kzz = a.Kzz + col * a.cs1; L.vz = a.vz + col * a.cs1; L.dzi = a.dzi + col * a.cs1;
    L.Ti = a.Ti + col * a.cs1;   L.Hpi = a.Hpi + col * a.cs1;
    L.Dzz = a.Dzz + col * a.csn; L.vs = a.vs + col * a.csn;
    L.Tco = a.Tco + col * a.csz; L.g = a.g + col * a.csz;
    L.ms = a.ms + col * a.csi;   L.alpha = a.alpha + col * a.csi; L.top_flux = a.top_flux + col * a.csi;
    L.bot_flux = a.bot_flux + col * a.csi; L.bot_vdep = a.bot_vdep + col * a.csi;
    return L;
}
// thermal / gravity bracket: -1./Hpi[m] + ms*gx/(Navo*kb*Ti[m]) + alpha/Ti[m]*(Tco[m+1]-Tco[m])/dzi[m]   (op.py:1554-1580)
__device__ __forceinline__ double phi_br(const AtmLayer &L, int m, int i, double gx)
{
    return -1. / L.Hpi[m] + L.ms[i] * gx / (VK_NAVO * VK_KB * L.Ti[m]) + L.alpha[i] / L.Ti[m] * (L.Tco[m + 1] - L.Tco[m]) / L.dzi[m];
}

// ------------------------------------------------------------------------------------------------------------------
// atmosphere-only pieces, arrays [ncol_atm][nz][ni]:
//   Q  : Dzz[m]/dzi[m] at interface m = j (row nz-1 unused)
//   QB : interior 1./dz_ave*Dzz[j]/dzi[j]      ; j=0: 1./dzi0*(D0/dzi0)            ; top: -1./dzi_m*(Dm/dzi_m)   (the A quotient)
//   QC : interior 1./dz_ave*Dzz[j-1]/dzi[j-1]  ; j=0: -1./dzi0*(D0/dzi0) (A quot.) ; top:  1./dzi_m*(Dm/dzi_m)
//   TA/TB/TC : thermal-gravity terms; SA/SB/SC : settling terms (signs applied by the consumer exactly as op.py does)
__global__ void atm_pre_kernel(AtmDev a, int ncol_atm, AtmPre p, const int *pred)
{
    const int nz = a.nz, ni = a.ni;
    const int col = blockIdx.x / nz, j = blockIdx.x % nz;
    if (col >= ncol_atm) return;
    if (pred && !pred[col]) return;
    const AtmLayer L = atm_at(a, col);
    const double *dzi = L.dzi, *Dzz = L.Dzz, *vs = L.vs, *g = L.g;
    const size_t base = ((size_t)col * nz + j) * ni;
    const double *vmv = a.use_vm_mol ? a.vm + col * a.csv : nullptr;
    const int stl = a.use_settling && a.use_moldiff;
    for (int i = threadIdx.x; i < ni; i += blockDim.x) {
        double Q = 0, QB = 0, QC = 0, TA = 0, TB = 0, TC = 0, SA = 0, SB = 0, SC = 0;
        if (j < nz - 1) Q = Dzz[(size_t)j * ni + i] / dzi[j];
        if (a.use_vm_mol) {
            // use_vm_mol (diffdf_vm op.py:1644-1673, diffdf_settling_vm op.py:1845-1879 and their lhs twins): no thermal/gravity
            // brackets; S* hold the upwind terms of vm, T* those of vs.  vm has nz rows (`vm[-1]` = row nz-1, `vs[-1]` = row nz-2);
            // the settling variant has NO vm term in the bottom row, so S* carry the vs term there.
            if (j == 0) {
                double D0 = Dzz[i];
                QB = 1. / (dzi[0]) * (D0 / dzi[0]);
                QC = -1. / (dzi[0]) * (D0 / dzi[0]);
                const double w = stl ? vs[i] : vmv[i];
                SA = (posv(w)) / dzi[0];
                SB = (negv(w)) / dzi[0];
            } else if (j == nz - 1) {
                const int m = nz - 2;
                double Dm = Dzz[(size_t)m * ni + i];
                QB = -1. / (dzi[m]) * (Dm / dzi[m]);
                QC = 1. / (dzi[m]) * (Dm / dzi[m]);
                const double wt = vmv[(size_t)(nz - 1) * ni + i];
                SA = (negv(wt)) / dzi[m];
                SC = (posv(wt)) / dzi[m];
                TA = (negv(vs[(size_t)m * ni + i])) / dzi[m];
                TC = (posv(vs[(size_t)m * ni + i])) / dzi[m];
            } else {
                double dz_ave = 0.5 * (dzi[j - 1] + dzi[j]);
                double Dj = Dzz[(size_t)j * ni + i], Dm = Dzz[(size_t)(j - 1) * ni + i];
                QB = 1. / dz_ave * Dj / dzi[j];
                QC = 1. / dz_ave * Dm / dzi[j - 1];
                const double wj = vmv[(size_t)j * ni + i], wm = vmv[(size_t)(j - 1) * ni + i];
                SA = (posv(wj) - negv(wm)) / dz_ave;
                SB = (negv(wj)) / dz_ave;
                SC = (posv(wm)) / dz_ave;
                const double vj = vs[(size_t)j * ni + i], vmn = vs[(size_t)(j - 1) * ni + i];
                TA = (posv(vj) - negv(vmn)) / dz_ave;
                TB = (negv(vj)) / dz_ave;
                TC = (posv(vmn)) / dz_ave;
            }
        } else if (j == 0) {
            double D0 = Dzz[i];
            QB = 1. / (dzi[0]) * (D0 / dzi[0]);
            QC = -1. / (dzi[0]) * (D0 / dzi[0]);
            double br = phi_br(L, 0, i, g[0]);
            TA = 1. / (dzi[0]) * D0 / 2. * br;
            TB = TA;
            SA = (posv(vs[i])) / dzi[0];
            SB = (negv(vs[i])) / dzi[0];
        } else if (j == nz - 1) {
            const int m = nz - 2;
            double Dm = Dzz[(size_t)m * ni + i];
            QB = -1. / (dzi[m]) * (Dm / dzi[m]);
            QC = 1. / (dzi[m]) * (Dm / dzi[m]);
            double br = phi_br(L, m, i, g[nz - 1]);
            TA = 1. / (dzi[m]) * Dm / 2. * br;
            TC = TA;
            SA = (negv(vs[(size_t)m * ni + i])) / dzi[m];
            SC = (posv(vs[(size_t)m * ni + i])) / dzi[m];
        } else {
            double dz_ave = 0.5 * (dzi[j - 1] + dzi[j]);
            double Dj = Dzz[(size_t)j * ni + i], Dm = Dzz[(size_t)(j - 1) * ni + i];
            QB = 1. / dz_ave * Dj / dzi[j];
            QC = 1. / dz_ave * Dm / dzi[j - 1];
            TA = 1. / (2. * dz_ave) * (Dj * phi_br(L, j, i, g[j]) - Dm * phi_br(L, j - 1, i, g[j]));
            TB = 1. / (2. * dz_ave) * Dj * phi_br(L, j, i, g[j + 1]);
            TC = 1. / (2. * dz_ave) * Dm * phi_br(L, j - 1, i, g[j - 1]);
            double vj = vs[(size_t)j * ni + i], vm = vs[(size_t)(j - 1) * ni + i];
            SA = (posv(vj) - negv(vm)) / dz_ave;
            SB = (negv(vj)) / dz_ave;
            SC = (posv(vm)) / dz_ave;
        }
        p.Q[base + i] = Q; p.QB[base + i] = QB; p.QC[base + i] = QC;
        p.TA[base + i] = TA; p.TB[base + i] = TB; p.TC[base + i] = TC;
        p.SA[base + i] = SA; p.SB[base + i] = SB; p.SC[base + i] = SC;
    }
    if (threadIdx.x == 0) {
        // per-layer scalars: [0] A prefactor (rhs form), [1] Kzz[j]/dzi[j], [2] Kzz[j-1]/dzi[j-1], [3] B prefactor (rhs form),
        // [4] C prefactor (rhs form), [5..7] advection terms added to A, B, C, [8] -1/dz_ave, [9] 1/dz_ave
        const double *Kzz = L.Kzz, *vz = L.vz;
        double s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (j == 0) {
            s[0] = -1. / (dzi[0]) * (Kzz[0] / dzi[0]);
            s[3] = 1. / (dzi[0]) * (Kzz[0] / dzi[0]);
            s[5] = -(posv(vz[0])) / dzi[0];
            s[6] = -(negv(vz[0])) / dzi[0];
        } else if (j == nz - 1) {
            const int m = nz - 2;
            s[0] = -1. / (dzi[m]) * (Kzz[m] / dzi[m]);
            s[4] = 1. / (dzi[m]) * (Kzz[m] / dzi[m]);
            s[5] = (negv(vz[m])) / dzi[m];
            s[7] = (posv(vz[m])) / dzi[m];
        } else {
            double dz_ave = 0.5 * (dzi[j - 1] + dzi[j]);
            s[1] = Kzz[j] / dzi[j];
            s[2] = Kzz[j - 1] / dzi[j - 1];
            if (a.use_moldiff) {
                s[0] = -1. / dz_ave;
                s[3] = 1. / dz_ave * Kzz[j] / dzi[j];
                s[4] = 1. / dz_ave * Kzz[j - 1] / dzi[j - 1];
            } else {   // diffdf_no_mol writes 2./(dzi[j-1]+dzi[j]) (op.py:1474-1476)
                s[0] = -2. / (dzi[j - 1] + dzi[j]);
                s[3] = 2. / (dzi[j - 1] + dzi[j]) * Kzz[j] / dzi[j];
                s[4] = 2. / (dzi[j - 1] + dzi[j]) * Kzz[j - 1] / dzi[j - 1];
            }
            s[5] = -(posv(vz[j]) - negv(vz[j - 1])) / dz_ave;
            s[6] = -(negv(vz[j])) / dz_ave;
            s[7] = (posv(vz[j - 1])) / dz_ave;
            s[8] = -1. / dz_ave;
            s[9] = 1. / dz_ave;
        }
        for (int q = 0; q < 10; q++) p.LS[((size_t)col * nz + j) * 10 + q] = s[q];
    }
}

int launch_atm_pre(vk_column *c, int ncol_atm)
{
    atm_pre_kernel<<<ncol_atm * c->nz, 96, 0, c->stream>>>(c->atm, ncol_atm, c->atm.pre, nullptr);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}
// re-evaluation for the columns whose grid update_mu_dz has just changed (vk_steady.cu); per-column atmosphere only
int launch_atm_pre_pred(vk_column *c, const int *pred)
{
    if (c->atm.pre_cs == 0 && c->ncol > 1) { set_error("per-column atmosphere needed"); return VK_ERR_INVALID; }
    atm_pre_kernel<<<c->ncol * c->nz, 96, 0, c->stream>>>(c->atm, c->ncol, c->atm.pre, pred);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

// ---- use_vm_mol: how the consumers combine the upwind terms (pos: 0 bottom row, 1 interior, 2 top row).  Ai/Bi/Ci enter
// holding the diffusive part only.  The reference adds the terms in a different order on the two sides:
//   diffdf_vm / diffdf_settling_vm (op.py:1644-1650, 1670-1673 / 1845-1855, 1876-1879): interior  Ai += -(vm..)/dz -(vs..)/dz  is ONE
//   right-hand expression; top row vm[-1] first, then vs[-1]
//   lhs_jac_tot_vm / lhs_jac_settling_vm (op.py:2090-2118 / 2412-2442): vs terms first, then vm terms, each its own subtraction
__device__ __forceinline__ void vm_rhs_adv(const AtmPre &P, size_t pb, int pos, int st, double &Ai, double &Bi, double &Ci)
{
    if (pos == 0) {
        Ai = Ai - P.SA[pb];
        Bi = Bi - P.SB[pb];
    } else if (pos == 2) {
        Ai = Ai + P.SA[pb];
        Ci = Ci + P.SC[pb];
        if (st) {
            Ai = Ai + P.TA[pb];
            Ci = Ci + P.TC[pb];
        }
    } else if (st) {
        Ai += -P.SA[pb] - P.TA[pb];
        Bi += -P.SB[pb] - P.TB[pb];
        Ci += +P.SC[pb] + P.TC[pb];
    } else {
        Ai += -P.SA[pb];
        Bi += -P.SB[pb];
        Ci += +P.SC[pb];
    }
}
__device__ __forceinline__ void vm_lhs_adv(const AtmPre &P, size_t pb, int pos, int st, double &ta, double &tb, double &tc)
{
    if (pos == 0) {
        ta = ta - P.SA[pb];
        tb = tb - P.SB[pb];
    } else if (pos == 2) {
        if (st) {
            ta = ta + P.TA[pb];
            tc = tc + P.TC[pb];
        }
        ta = ta + P.SA[pb];
        tc = tc + P.SC[pb];
    } else {
        if (st) {
            ta = ta - P.TA[pb];
            tb = tb - P.TB[pb];
            tc = tc + P.TC[pb];
        }
        ta = ta - P.SA[pb];
        tb = tb - P.SB[pb];
        tc = tc + P.SC[pb];
    }
}
// diffusion-limited escape on the top diagonal, only in the *_vm lhs variants (op.py:2101-2107, 2425-2431)
__device__ __forceinline__ double vm_diff_lim(const AtmDev &a, const AtmLayer &L, int i, double ytop)
{
    double diff_lim = 0.0;
    for (int q = 0; q < a.n_diff_esc; q++)
        if (a.diff_esc_idx[q] == i && ytop > 0) diff_lim += L.top_flux[i] / ytop;
    return diff_lim;
}

// per-layer scalars that depend on the layer sums (one thread per block computes them)
struct LayerScal {
    double Aa, Bb, Cc;        // eddy + advection parts (diffdf form)
    double m1;                // -1./dz_ave (interior)
    double sp, sm;            // (ysp+ys0), (ys0+ysm)
    double ys0, ysp, ysm;
};

struct RhsArgs {
    NetDev net;
    AtmDev atm;
    int nz;
    const double *y;        // [ncol][nz][ni]
    const double *k;        // [ncol|1][nz][nr+1]
    size_t k_cs;
    const double *k1;       // stage 2: y_eval = y + k1/r ; NULL for stage 1
    const double *dt;       // [ncol] (stage 2)
    double *yk2_out;        // stage 2: store y + k1/r
    double *out_sum, *out_chem, *out_diff;  // any may be NULL; out_sum = chem + diff (stage 2: - 2/(r h) k1)
    const unsigned char *fix_mask;          // rows forced to zero (op.py:2896-2925)
    const int *act;                         // [ncol] or NULL: stopped columns are skipped
    int fast;                               // segmented summation of the production / loss terms instead of the reference's order
    const double *chem_in;                  // chemdf already evaluated by the emitted kernel of this network (vk_emit.cu): only the transport
                                            // stencil, the combine and the outputs are left to this kernel
};

// transport part of dy/dt of species i at layer j of column col (ODESolver.diffdf / diffdf_settling / diffdf_no_mol / *_vm, op.py:1438-1898),
// expression by expression in the reference's order; s: the layer scalars, y0v / ymv / ypv: y_i at layers j, j-1, j+1
__device__ __forceinline__ double stencil_diff(const AtmDev &atm, const AtmLayer &L, const LayerScal &s, int nz, int ni, int col, int j, int i,
                                               double y0v, double ymv, double ypv)
{
    const int md = atm.use_moldiff, st = atm.use_settling && atm.use_moldiff;
    const int vmm = atm.use_vm_mol;
    const double *dzi = L.dzi;
    const AtmPre &P = atm.pre;
    const size_t pb = ((size_t)col * atm.pre_cs + (size_t)j * ni) + i;
    double diff;
    if (j == 0) {
        if (md) {
            double Ai = P.QC[pb] * s.sp / 2. / s.ys0;
            double Bi = P.QB[pb] * s.sp / 2. / s.ysp;
            if (vmm) {
                double Cx = 0.0;
                vm_rhs_adv(P, pb, 0, st, Ai, Bi, Cx);
            } else {
                Ai = Ai + P.TA[pb];
                Bi = Bi + P.TB[pb];
                if (st) {
                    Ai = Ai - P.SA[pb];
                    Bi = Bi - P.SB[pb];
                }
            }
            diff = (s.Aa + Ai) * y0v + (s.Bb + Bi) * ypv;
        } else {
            diff = s.Aa * y0v + s.Bb * ypv;
        }
        if (atm.use_botflux) diff += (L.bot_flux[i] - y0v * L.bot_vdep[i]) / dzi[0];
    } else if (j == nz - 1) {
        if (md) {
            double Ai = P.QB[pb] * s.sm / 2. / s.ys0;
            double Ci = P.QC[pb] * s.sm / 2. / s.ysm;
            if (vmm) {
                double Bx = 0.0;
                vm_rhs_adv(P, pb, 2, st, Ai, Bx, Ci);
            } else {
                Ai = Ai - P.TA[pb];
                Ci = Ci - P.TC[pb];
                if (st) {
                    Ai = Ai + P.SA[pb];
                    Ci = Ci + P.SC[pb];
                }
            }
            diff = (s.Aa + Ai) * y0v + (s.Cc + Ci) * ymv;
        } else {
            diff = s.Aa * y0v + s.Cc * ymv;
        }
        if (atm.use_topflux) diff += L.top_flux[i] / dzi[nz - 2];
    } else {
        double t1 = s.Aa * y0v + s.Bb * ypv + s.Cc * ymv;
        if (md) {
            double Ai = s.m1 * (P.Q[pb] * s.sp / 2. + P.Q[pb - ni] * s.sm / 2.) / s.ys0;
            double Bi = P.QB[pb] * s.sp / 2. / s.ysp;
            double Ci = P.QC[pb] * s.sm / 2. / s.ysm;
            if (vmm) {
                vm_rhs_adv(P, pb, 1, st, Ai, Bi, Ci);
            } else {
                if (st) {
                    Ai = Ai - P.SA[pb];
                    Bi = Bi - P.SB[pb];
                    Ci = Ci + P.SC[pb];
                }
                Ai += P.TA[pb];
                Bi += P.TB[pb];
                Ci += -P.TC[pb];
            }
            double t2 = Ai * y0v + Bi * ypv + Ci * ymv;
            diff = t1 + t2;
        } else {
            diff = t1;
        }
    }
    return diff;
}

// ------------------------------------------------------------------------------------------------------------------
// rhs_warp_kernel: ONE WARP per (column, layer), RHS_WPB layers per block, no block-wide barrier after the tables are staged.
// The per-species production-loss sum must be added left to right in the reference's order (bit-identical RHS is the first parity
// gate), so the longest chain (H: 170 terms in NCHO) bounds the latency of a layer whatever the thread count; what matters is how
// many layers are in flight per SM and how few instructions a term costs.  Per warp: the three y rows, the 439 pair rates
// v_p = k_f prod(y) - k_r prod(y) (each lane forms whole pairs, so the separate differencing pass disappears) and the layer scalars
// live in 5.5 KB of shared memory; term descriptors are 16 bit (pair index + sign; all coefficients of the shipped networks are
// +-1, otherwise the general 32-bit descriptors are used) and are staged once per block; species are dealt to lanes longest
// chain first (vk_network_create).
#define RHS_WPB 8
struct RhsWarpSmem {    // offsets in doubles inside one warp's slab
    int ym, y0, yp, v, chem, scal, part, total;
};
__host__ __device__ inline RhsWarpSmem rhs_warp_layout(int ni, int nr, int n_seg)
{
    RhsWarpSmem L;
    L.ym = 0; L.y0 = L.ym + (ni + 2); L.yp = L.y0 + (ni + 2);
    L.v = L.yp + (ni + 2);                    // [nr/2 + 1] (+ the always-zero slot of the padding terms)
    L.chem = L.v + (nr / 2 + 2);              // [ni]
    L.scal = L.chem + ni + (ni & 1);          // LayerScal + ysum[4]
    L.part = L.scal + 16;                     // [n_seg] partial sums of the segmented summation
    L.total = L.part + n_seg + (n_seg & 1);
    return L;
}

__global__ void __launch_bounds__(RHS_WPB * 32) rhs_warp_kernel(RhsArgs A, int n_layers_total)
{
    extern __shared__ __align__(16) double sm[];
    const int ni = A.net.ni, nr = A.net.nr, nz = A.nz, npair = nr / 2;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const bool fast = A.fast != 0;
    const RhsWarpSmem SL = rhs_warp_layout(ni, nr, fast ? A.net.rhs_n_seg : 0);
    // block-shared tables: rate factors [nr+1] (uchar4), 16-bit descriptors [n_rhs] (reference order) or [32 T] (segmented, transposed)
    uchar4 *rfac = reinterpret_cast<uchar4 *>(sm + (size_t)RHS_WPB * SL.total);
    unsigned short *td = reinterpret_cast<unsigned short *>(rfac + (nr + 2));
    const bool have_chem = A.chem_in != nullptr;
    if (!have_chem) {
        for (int i = tid; i <= nr; i += blockDim.x) rfac[i] = A.net.rate_fac[i];
        if (fast) { for (int i = tid; i < 32 * A.net.rhs_flat_T; i += blockDim.x) td[i] = A.net.rhs_flat16[i]; }
        else if (A.net.rhs_unit) for (int i = tid; i < A.net.n_rhs; i += blockDim.x) td[i] = A.net.rhs_desc16[i];
        __syncthreads();
    }

    const int lay = blockIdx.x * RHS_WPB + w;
    if (lay >= n_layers_total) return;
    const int col = lay / nz, j = lay % nz;
    if (A.act && !A.act[col]) return;
    double *ws = sm + (size_t)w * SL.total;
    double *ym = ws + SL.ym, *y0 = ws + SL.y0, *yp = ws + SL.yp, *v = ws + SL.v, *chem_s = ws + SL.chem;
    LayerScal *S = reinterpret_cast<LayerScal *>(ws + SL.scal);
    double *ysum = ws + SL.scal + 10;         // [4] behind the 10 doubles of LayerScal

    const size_t base = ((size_t)col * nz + j) * ni;
    const double rr = 1. + 1. / sqrt(2.);
    for (int i = lane; i < ni; i += 32) {
        double v0 = A.y[base + i];
        double vm = (j > 0) ? A.y[base - ni + i] : 0.0;
        double vp = (j < nz - 1) ? A.y[base + ni + i] : 0.0;
        if (A.k1) {   // yk2 = y + k1/r   (op.py:2917)
            v0 = v0 + A.k1[base + i] / rr;
            if (j > 0) vm = vm + A.k1[base - ni + i] / rr;
            if (j < nz - 1) vp = vp + A.k1[base + ni + i] / rr;
            if (A.yk2_out) A.yk2_out[base + i] = v0;
        }
        y0[i] = v0; ym[i] = vm; yp[i] = vp;
    }
    if (lane == 0) { y0[ni] = A.atm.M[col * A.atm.csz + j]; y0[ni + 1] = 1.0; }
    __syncwarp();

    // ---- pair rates v_p = rate[2p+1] - rate[2p+2], rate[i] = k[i]*f0*f1*f2*f3 in written order (padding slots multiply by 1.0)
    const double *kg = A.k + col * A.k_cs + (size_t)j * (nr + 1);
    for (int p = lane; p < (have_chem ? 0 : npair); p += 32) {
        double r2[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int i = 2 * p + 1 + h;
            const uchar4 f = rfac[i];
            double x = kg[i];
            if (!A.net.has_pow) {
                x = x * y0[f.x]; x = x * y0[f.y]; x = x * y0[f.z]; x = x * y0[f.w];
            } else {
                const uchar4 pw = A.net.rate_pow[i];
                const unsigned char ff[4] = {f.x, f.y, f.z, f.w}, pp[4] = {pw.x, pw.y, pw.z, pw.w};
                for (int q = 0; q < 4; q++) {
                    const double b = y0[ff[q]];
                    const double tt = (pp[q] == 1) ? b : ((pp[q] == 2) ? b * b : pow(b, (double)pp[q]));
                    x = x * tt;
                }
            }
            r2[h] = x;
        }
        v[p] = r2[0] - r2[1];
    }
    // ---- the three layer sums (numpy association) and the layer scalars of the stencil
    {
        const int q = lane >> 3;
        const int jj = j - 1 + q;
        const bool act = (q < 3) && jj >= 0 && jj < nz;
        const double *row = (q == 0) ? ym : ((q == 1) ? y0 : yp);
        double sres;
        if (A.atm.n_gas > 0) {     // np.sum(y[:,gas_indx], axis=1) is a plain left-to-right sum (see row_sum)
            sres = 0.0;
            if (act && (lane & 7) == 0) sres = row_sum(row, ni, A.atm.n_gas, A.atm.gas_indx, nullptr);
        } else {
            sres = np_pairwise_group8(row, ni, act);
        }
        if (act && (lane & 7) == 0) ysum[q] = sres;
    }
    __syncwarp();
    const AtmLayer L = atm_at(A.atm, col);
    const int md = A.atm.use_moldiff, st = A.atm.use_settling && A.atm.use_moldiff;
    const int vmm = A.atm.use_vm_mol;
    if (lane < 3) {
        const double *ls = A.atm.pre.LS + ((size_t)col * (A.atm.pre_cs ? nz : 0) + j) * 10;
        const double ys0 = ysum[1], ysm = ysum[0], ysp = ysum[2];
        const double sp = ysp + ys0, smm = ys0 + ysm;
        if (lane == 0) {
            double x;
            if (j == 0) x = ls[0] * sp / 2. / ys0;
            else if (j == nz - 1) x = ls[0] * smm / 2. / ys0;
            else x = ls[0] * (ls[1] * sp / 2. + ls[2] * smm / 2.) / ys0;
            x += ls[5];
            S->Aa = x; S->m1 = ls[8]; S->sp = sp; S->sm = smm; S->ys0 = ys0; S->ysp = ysp; S->ysm = ysm;
        } else if (lane == 1) {
            double x = 0.0;
            if (j < nz - 1) { x = ls[3] * sp / 2. / ysp; x += ls[6]; }
            S->Bb = x;
        } else {
            double x = 0.0;
            if (j > 0) { x = ls[4] * smm / 2. / ysm; x += ls[7]; }
            S->Cc = x;
        }
    }
    __syncwarp();

    // ---- chemistry, segmented order (default): the terms of all species, in species order, are cut into 32 equal chunks; every lane sums
    // its chunk (one chain of T ~ 52 adds, a partial sum flushed at every species / chunk boundary), then the <= 5 partials of a species
    // are added in order.  Differs from the reference's left-to-right sum by the rounding of those few partial sums (checked against an
    // extended-precision sum, tests/test_gpu_parity.py::test_rhs_segmented_order); the warp no longer waits for its longest chain (H:
    // 170 dependent adds while 31 lanes idle) - 420 instead of 2450 issue slots per layer for NCHO.
    if (fast && !have_chem) {
        double *part = ws + SL.part;
        if (lane == 0) v[npair] = 0.0;
        __syncwarp();
        const int T = A.net.rhs_flat_T;
        int pslot = A.net.rhs_lane_slot0[lane];
        double acc = 0.0;
#pragma unroll 4
        for (int q = 0; q < T; q++) {
            const unsigned d = td[32 * q + lane];
            const double x = v[d & 0x3fffu];
            acc = acc + __hiloint2double(__double2hiint(x) ^ (int)((d & 0x8000u) << 16), __double2loint(x));
            if (d & 0x4000u) { part[pslot++] = acc; acc = 0.0; }
        }
        __syncwarp();
        for (int s = lane; s < ni; s += 32) {
            double chem = 0.0;
            for (int p = A.net.rhs_seg_ptr[s]; p < A.net.rhs_seg_ptr[s + 1]; p++) chem = chem + part[p];
            chem_s[s] = chem;
        }
    }
    // ---- chemistry, reference order: left-to-right sum of coef * v_p per species in network order (make_chem_funs.py:258-285)
    const int *my_sp = A.net.rhs_lane_sp + lane * VK_RHS_SPL;
    for (int slot = 0; slot < ((fast || have_chem) ? 0 : VK_RHS_SPL); slot++) {
        const int s = my_sp[slot];
        if (s < 0) break;
        const int q0 = A.net.rhs_ptr[s], q1 = A.net.rhs_ptr[s + 1];
        double chem = 0.0;
        if (A.net.rhs_unit) {
            // x - v is exactly x + (-1.0 * v): the sign is applied by flipping the sign bit of v
            auto term = [&](int q) -> double {
                const unsigned d = td[q];
                const double x = v[d & 0x7fffu];
                return __hiloint2double(__double2hiint(x) ^ (int)((d & 0x8000u) << 16), __double2loint(x));
            };
            int q = q0;
            if (q < q1) { chem = term(q); q++; }
            for (; q + 4 <= q1; q += 4) {
                const double a0 = term(q), a1 = term(q + 1), a2 = term(q + 2), a3 = term(q + 3);
                chem = chem + a0; chem = chem + a1; chem = chem + a2; chem = chem + a3;
            }
            for (; q < q1; q++) chem = chem + term(q);
        } else {
            for (int q = q0; q < q1; q++) {
                const int t = A.net.rhs_term[q];
                const double x = (double)((signed char)(t & 0xff)) * v[((t >> 8) - 1) >> 1];
                chem = (q == q0) ? x : chem + x;
            }
        }
        chem_s[s] = chem;
    }
    __syncwarp();

    // ---- transport stencil + output, species i = lane, lane + 32, ...
    const LayerScal s = *S;
    for (int i = lane; i < ni; i += 32) {
        const double diff = stencil_diff(A.atm, L, s, nz, ni, col, j, i, y0[i], ym[i], yp[i]);
        const double chem = have_chem ? A.chem_in[base + i] : chem_s[i];
        if (A.out_chem) A.out_chem[base + i] = chem;
        if (A.out_diff) A.out_diff[base + i] = diff;
        if (A.out_sum) {
            double f = chem + diff;                                   // op.py:2892 / 2918
            if (A.fix_mask && A.fix_mask[base + i]) f = 0.0;          // op.py:2904, 2924
            if (A.k1) {
                double c = 2. / (rr * A.dt[col]);
                f = f - c * A.k1[base + i];                           // op.py:2928
            }
            A.out_sum[base + i] = f;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The right-hand side behind an EMITTED chemdf kernel (vk_emit.cu): chemdf and the layer sums of y are already in HBM, what is left is
// elementwise.  layer_scal_kernel: one thread per (column, layer) forms the eddy / advection coefficients of the layer (the scalars lanes
// 0-2 of rhs_warp_kernel form); rhs_stencil_kernel: one thread per (column, layer, species) adds the transport stencil, applies the stage-2
// combine and the fixed-row mask - same expressions, same order, bit-identical to rhs_warp_kernel.
struct StencilArgs {
    AtmDev atm;
    int nz, ni, ncol;
    const double *y;        // y (stage 1) or y + k1/r (stage 2, written by the emitted kernel)
    const double *k1, *dt;  // stage 2 combine
    const double *chem, *ysum;
    LayerScal *S;           // [ncol][nz]
    double *out_sum, *out_chem, *out_diff;
    const unsigned char *fix_mask;
    const int *act;
};
__global__ void __launch_bounds__(128) layer_scal_kernel(StencilArgs A)
{
    const int lay = blockIdx.x * blockDim.x + threadIdx.x;
    const int nz = A.nz;
    if (lay >= A.ncol * nz) return;
    const int col = lay / nz, j = lay - col * nz;
    if (A.act && !A.act[col]) return;
    const double *ls = A.atm.pre.LS + ((size_t)col * (A.atm.pre_cs ? nz : 0) + j) * 10;
    const double ys0 = A.ysum[lay], ysm = (j > 0) ? A.ysum[lay - 1] : 0.0, ysp = (j < nz - 1) ? A.ysum[lay + 1] : 0.0;
    const double sp = ysp + ys0, smm = ys0 + ysm;
    LayerScal S;
    double x;
    if (j == 0) x = ls[0] * sp / 2. / ys0;
    else if (j == nz - 1) x = ls[0] * smm / 2. / ys0;
    else x = ls[0] * (ls[1] * sp / 2. + ls[2] * smm / 2.) / ys0;
    x += ls[5];
    S.Aa = x; S.m1 = ls[8]; S.sp = sp; S.sm = smm; S.ys0 = ys0; S.ysp = ysp; S.ysm = ysm;
    x = 0.0;
    if (j < nz - 1) { x = ls[3] * sp / 2. / ysp; x += ls[6]; }
    S.Bb = x;
    x = 0.0;
    if (j > 0) { x = ls[4] * smm / 2. / ysm; x += ls[7]; }
    S.Cc = x;
    A.S[lay] = S;
}
__global__ void __launch_bounds__(256) rhs_stencil_kernel(StencilArgs A)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ni = A.ni, nz = A.nz;
    if (e >= (size_t)A.ncol * nz * ni) return;
    const int lay = (int)(e / ni), i = (int)(e - (size_t)lay * ni);
    const int col = lay / nz, j = lay - col * nz;
    if (A.act && !A.act[col]) return;
    const AtmLayer L = atm_at(A.atm, col);
    const LayerScal s = A.S[lay];
    const double y0v = A.y[e];
    const double ymv = (j > 0) ? A.y[e - ni] : 0.0;
    const double ypv = (j < nz - 1) ? A.y[e + ni] : 0.0;
    const double diff = stencil_diff(A.atm, L, s, nz, ni, col, j, i, y0v, ymv, ypv);
    const double chem = A.chem[e];
    if (A.out_chem) A.out_chem[e] = chem;
    if (A.out_diff) A.out_diff[e] = diff;
    if (A.out_sum) {
        double f = chem + diff;                                   // op.py:2892 / 2918
        if (A.fix_mask && A.fix_mask[e]) f = 0.0;                 // op.py:2904, 2924
        if (A.k1) {
            const double rr = 1. + 1. / sqrt(2.);
            double c = 2. / (rr * A.dt[col]);
            f = f - c * A.k1[e];                                  // op.py:2928
        }
        A.out_sum[e] = f;
    }
}

// ------------------------------------------------------------------------------------------------------------------
struct LhsArgs {
    NetDev net;
    AtmDev atm;
    int nz;
    const double *y;
    const double *k;
    size_t k_cs;
    const double *dt;    // [ncol]
    double *D;           // [ncol][nz][ld][ld]
    double *up, *dn;     // [ncol][nz][ld]
    int ld;              // nip for the solver layout, ni for the dense diagnostic output
    const unsigned char *fix_mask;
    const int *act;
};

// ------------------------------------------------------------------------------------------------------------------
// lhs_ml_kernel: chem_funs.neg_symjac block + lhs_jac_* diagonal and couplings.
//  * One thread block walks `lpb` (1..30, chosen by the launcher) consecutive layers of a column; the network tables (distinct products, 16-bit term descriptors,
//    segment schedule - 40 KB for NCHO) are staged in shared memory ONCE per block, so the per-term work never waits on L2.
//  * The 5953 Jacobian terms of NCHO are coefficient x one of only 1603 distinct products k_r y_a y_b y_c: the products are formed
//    once per layer (phase A), a term is then one 16-bit descriptor, one product load, one multiply by +-1/2/4 (exact) and one add.
//  * k and the three y rows of the NEXT layer are fetched with cp.async while the current layer is assembled; the finished block is
//    streamed out with 16-byte stores and zeroed behind the copy.
__constant__ double c_jac_coef[8] = {1., -1., 2., -2., 4., -4., 3., -3.};

struct LhsMlSmem {      // offsets in doubles
    int kz, ym, y0, yp, dprod, part, misc, blk, tab, total_bytes;
};
static inline LhsMlSmem lhs_ml_layout(const NetDev &n, int ld)
{
    LhsMlSmem L;
    int o = 0;
    L.kz = o; o += n.nr + 2;
    L.ym = o; o += n.ni + 2; L.y0 = o; o += n.ni + 2; L.yp = o; o += n.ni + 2;
    L.dprod = o; o += n.n_uniq + 2;
    L.part = o; o += n.n_part + 2;
    L.misc = o; o += 16 + 16;       // [0..3] layer sums, [4] spare, [8..15] coefficient table, [16..31] 128 diagonal-in-pattern flags
    o += o & 1;
    L.blk = o; o += ld * ld + (ld & 1);
    L.tab = o;
    size_t bytes = sizeof(double) * (size_t)o + sizeof(uint2) * (n.n_grp + 1) + sizeof(uint2) * (n.n_multi + 1) +
                   sizeof(unsigned) * (n.n_uniq + 2) + sizeof(unsigned) * ((size_t)n.n_grp * 32) + sizeof(unsigned short) * (n.n_tt + 8);
    L.total_bytes = (int)((bytes + 15) & ~(size_t)15);
    return L;
}

__device__ __forceinline__ void cp_async8(double *dst, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int NTHR>
__global__ void __launch_bounds__(NTHR, 2) lhs_ml_kernel(LhsArgs A, LhsMlSmem SL, int blocks_per_col, int lpb, int dbg)
{
    extern __shared__ __align__(16) double sm[];
    const int ni = A.net.ni, nr = A.net.nr, nz = A.nz, ld = A.ld;
    const int col = blockIdx.x / blocks_per_col, j0 = (blockIdx.x % blocks_per_col) * lpb;
    if (A.act && !A.act[col]) return;
    const int j1 = min(j0 + lpb, nz);
    const int tid = threadIdx.x, nt = blockDim.x;
    double *kz = sm + SL.kz, *ym = sm + SL.ym, *y0 = sm + SL.y0, *yp = sm + SL.yp, *dprod = sm + SL.dprod, *part = sm + SL.part;
    double *ysum = sm + SL.misc, *ctab = sm + SL.misc + 8, *blk = sm + SL.blk;
    uint2 *grp = reinterpret_cast<uint2 *>(sm + SL.tab);
    uint2 *multi = grp + (A.net.n_grp + 1);
    unsigned *uq = reinterpret_cast<unsigned *>(multi + (A.net.n_multi + 1));
    unsigned *seg = uq + (A.net.n_uniq + 2);
    unsigned short *tt = reinterpret_cast<unsigned short *>(seg + (size_t)A.net.n_grp * 32);

    auto prefetch = [&](int j) {      // k row and the three y rows of layer j -> shared memory (cp.async, 8 bytes each)
        const double *kg = A.k + col * A.k_cs + (size_t)j * (nr + 1);
        for (int i = tid; i <= nr; i += nt) cp_async8(kz + i, kg + i);
        const size_t base = ((size_t)col * nz + j) * ni;
        for (int i = tid; i < ni; i += nt) {
            cp_async8(y0 + i, A.y + base + i);
            if (j > 0) cp_async8(ym + i, A.y + base - ni + i);
            if (j < nz - 1) cp_async8(yp + i, A.y + base + ni + i);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(j0);
    for (int i = tid; i < A.net.n_grp; i += nt) grp[i] = A.net.jac_grp[i];
    for (int i = tid; i < A.net.n_grp * 32; i += nt) seg[i] = A.net.jac_seg4[i];
    for (int i = tid; i < A.net.n_multi; i += nt) multi[i] = A.net.jac_multi[i];
    for (int i = tid; i < A.net.n_uniq; i += nt) uq[i] = A.net.jac_uniq[i];
    {   // 16-bit descriptors copied as 32-bit words (the table is 4-byte aligned and padded)
        const unsigned *src = reinterpret_cast<const unsigned *>(A.net.jac_tt);
        unsigned *dst = reinterpret_cast<unsigned *>(tt);
        for (int i = tid; i < (A.net.n_tt + 1) / 2; i += nt) dst[i] = src[i];
    }
    if (tid == 0) dprod[A.net.n_uniq] = 0.0;      // the padding product
    // The block buffer is zeroed ONCE: every layer has the same sparsity pattern, the gather assigns (never accumulates into) each
    // pattern entry, so entries outside the pattern stay zero from layer to layer.  Only the diagonal needs care: a diagonal entry that
    // is not in the chemical pattern would otherwise carry the previous layer's value into  d = c0 + blk[i][i]  -> per-species flags.
    for (int q = tid; q < ld * ld; q += nt) blk[q] = 0.0;
    unsigned char *dflag = reinterpret_cast<unsigned char *>(sm + SL.misc + 16);
    if (tid < 128) dflag[tid] = 0;
    if (tid == 0) { y0[ni + 1] = 1.0; }
    if (tid < 8) ctab[tid] = c_jac_coef[tid];     // per-lane lookups: shared memory (a constant bank would serialise)
    __syncthreads();
    for (int i = tid; i < A.net.n_grp * 32; i += nt) {
        const unsigned sg = seg[i];
        const unsigned row = sg & 0xffu, colx = (sg >> 8) & 0xffu;
        if (row != 0xffu && row == colx) dflag[row] = 1;       // whole entry or a partial of a split one: both end up in blk[row][row]
    }
    for (int i = tid; i < A.net.n_multi; i += nt) {
        const uint2 me = multi[i];
        if ((me.x & 0xffff) == (me.x >> 16)) dflag[me.x & 0xffff] = 1;
    }
    const unsigned blk_s = (unsigned)__cvta_generic_to_shared(blk);
    const unsigned blk_bytes = (unsigned)(ld * ld * sizeof(double));
    const bool bulk = (blk_bytes % 16u) == 0 && !(dbg & 16);
    const int warp = tid >> 5;
    // gather hand-out: pairs of groups dealt round-robin, the warps that own no species (no transport work) first
    const int nwarp = nt >> 5, nw_tr = (ld + 31) >> 5;
    const int gslot = (warp >= nw_tr) ? warp - nw_tr : warp + (nwarp - nw_tr);
    const AtmLayer L = atm_at(A.atm, col);
    const double *dzi = L.dzi;
    const int md = A.atm.use_moldiff, st = A.atm.use_settling && A.atm.use_moldiff;
    const int vmm = A.atm.use_vm_mol;
    const double rr = 1. + 1. / sqrt(2.);
    const double c0 = 1. / (rr * A.dt[col]);
    const AtmPre &P = A.atm.pre;

    for (int j = j0; j < j1; j++) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (tid == 0) {
            y0[ni] = A.atm.M[col * A.atm.csz + j];
            // the previous layer's block is still being read out of shared memory by the bulk store: the first writers (gather /
            // diagonal) run behind the barrier below
            if (bulk) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
        // ---- phase A: distinct products k_r y_a y_b y_c; the three layer sums
        if (!(dbg & 4)) for (int u = tid; u < A.net.n_uniq; u += nt) {
            const unsigned d = uq[u];
            double x = kz[d & 0x7ffu];
            x = x * y0[(d >> 11) & 0x7fu];
            x = x * y0[(d >> 18) & 0x7fu];
            x = x * y0[(d >> 25) & 0x7fu];
            dprod[u] = x;
        }
        if (tid >= 224) {              // last warp: the three layer sums, 8 lanes each, numpy association
            const int lane_ = tid - 224, q = lane_ >> 3;
            const int jj = j - 1 + q;
            const bool act = (q < 3) && jj >= 0 && jj < nz;
            const double *row = (q == 0) ? ym : ((q == 1) ? y0 : yp);
            double sres;
            if (A.atm.n_gas_lhs > 0) {
                sres = 0.0;
                if (act && (lane_ & 7) == 0) sres = row_sum(row, ni, A.atm.n_gas_lhs, A.atm.gas_indx_lhs, nullptr);
            } else {
                sres = np_pairwise_group8(row, ni, act);
            }
            if (act && (lane_ & 7) == 0) ysum[q] = sres;
        }
        __syncthreads();
        if (j + 1 < j1) prefetch(j + 1);           // k / y of this layer are consumed: fetch the next layer behind the assembly
        // ---- transport part of the diagonal and the couplings up/dn (op.py:1998-2040): independent of the chemical Jacobian, so the
        // threads that own a species do it FIRST (6 true divisions, ~1500 clk of latency) and then join the gather, which hands out
        // its groups through a shared counter - the other warps start gathering at once.
        const size_t base = ((size_t)col * nz + j) * ni;
        const size_t vbase = ((size_t)col * nz + j) * ld;
        double eA = 0.0, tA = 0.0, tV = 0.0;       // subtracted from the diagonal in this order after the gather
        double tE = 0.0;                           // diffusion-limited escape (top row of the *_vm variants), subtracted first
        if (tid < ld && !(dbg & 8)) {
            const int i = tid;
            if (i >= ni) {
                A.up[vbase + i] = 0.0;
                A.dn[vbase + i] = 0.0;
            } else {
                const double ys0 = ysum[1], ysm = ysum[0], ysp = ysum[2];
                const double *ls = A.atm.pre.LS + ((size_t)col * (A.atm.pre_cs ? nz : 0) + j) * 10;
                const size_t pb = ((size_t)col * A.atm.pre_cs + (size_t)j * ni) + i;
                double eB = 0.0, eC = 0.0, u = 0.0, l = 0.0;
                if (j == 0) {
                    eA = ls[0] * (ysp + ys0) / (2. * ys0) + ls[5];
                    eB = ls[3] * (ysp + ys0) / (2. * ysp) + ls[6];
                    u -= eB;
                    if (md) {
                        double ta = P.QC[pb] * (ysp + ys0) / (2. * ys0);
                        double tb = P.QB[pb] * (ysp + ys0) / (2. * ysp);
                        if (vmm) {
                            double tx = 0.0;
                            vm_lhs_adv(P, pb, 0, st, ta, tb, tx);
                        } else {
                            ta = ta + P.TA[pb];
                            tb = tb + P.TB[pb];
                            if (st) {
                                ta = ta - P.SA[pb];
                                tb = tb - P.SB[pb];
                            }
                        }
                        tA = ta;
                        u -= tb;
                    }
                    if (A.atm.use_botflux) tV = -1. * L.bot_vdep[i] / dzi[0];
                } else if (j == nz - 1) {
                    if (vmm && A.atm.n_diff_esc > 0) tE = vm_diff_lim(A.atm, L, i, y0[i]);   // top layer: nothing is prefetched over y0
                    eA = ls[0] * (ysm + ys0) / (2. * ys0) + ls[5];
                    eC = ls[4] * (ysm + ys0) / (2. * ysm) + ls[7];
                    l -= eC;
                    if (md) {
                        double ta = P.QB[pb] * (ys0 + ysm) / (2. * ys0);
                        double tc = P.QC[pb] * (ys0 + ysm) / (2. * ysm);
                        if (vmm) {
                            double tx = 0.0;
                            vm_lhs_adv(P, pb, 2, st, ta, tx, tc);
                        } else {
                            ta = ta - P.TA[pb];
                            tc = tc - P.TC[pb];
                            if (st) {
                                ta = ta + P.SA[pb];
                                tc = tc + P.SC[pb];
                            }
                        }
                        tA = ta;
                        l -= tc;
                    }
                } else {
                    eA = ls[8] * (ls[1] * (ysp + ys0) / 2. + ls[2] * (ysm + ys0) / 2.) / ys0 + ls[5];
                    eB = ls[9] * (ls[1] * (ysp + ys0) / (2. * ysp)) + ls[6];
                    eC = ls[9] * (ls[2] * (ysm + ys0) / (2. * ysm)) + ls[7];
                    u -= eB;
                    l -= eC;
                    if (md) {
                        double ta = ls[8] * (P.Q[pb] * (ysp + ys0) / 2. + P.Q[pb - ni] * (ysm + ys0) / 2.) / ys0;
                        double tb = ls[9] * (P.Q[pb] * (ysp + ys0) / (2. * ysp));
                        double tc = ls[9] * (P.Q[pb - ni] * (ysm + ys0) / (2. * ysm));
                        if (vmm) {
                            vm_lhs_adv(P, pb, 1, st, ta, tb, tc);
                        } else {
                            ta = ta + P.TA[pb];
                            tb = tb + P.TB[pb];
                            tc = tc - P.TC[pb];
                            if (st) {
                                ta = ta - P.SA[pb];
                                tb = tb - P.SB[pb];
                                tc = tc + P.SC[pb];
                            }
                        }
                        tA = ta;
                        u -= tb;
                        l -= tc;
                    }
                }
                if (A.fix_mask && A.fix_mask[base + i]) { u = 0.0; l = 0.0; }
                A.up[vbase + i] = u;
                A.dn[vbase + i] = l;
            }
        }
        // ---- phase B: groups of 32 segments (<= 16 terms each, sorted by length), two adjacent groups per warp and turn (two
        // independent accumulation chains); term q of lane l at tt[base + 32 q + l]: conflict-free 16-bit loads, warp-uniform trip
        // counts, no divergence
        {
            const int lane = tid & 31;
            for (int gI = 2 * gslot; gI < A.net.n_grp && !(dbg & 2); gI += 2 * nwarp) {
                const int gJ = gI + 1;
                const bool two = gJ < A.net.n_grp;
                const uint2 ga = grp[gI], gb = grp[two ? gJ : gI];
                const unsigned short *ta = tt + ga.x + lane, *tb = tt + gb.x + lane;
                const int na = (int)ga.y, nb = two ? (int)gb.y : 0;     // na >= nb (sorted by decreasing length)
                double acc_a = 0.0, acc_b = 0.0;
                int q = 0;
#pragma unroll 2
                for (; q < nb; q++) {
                    const unsigned da = ta[32 * q], db = tb[32 * q];
                    acc_a += ctab[da >> 13] * dprod[da & 0x1fffu];
                    acc_b += ctab[db >> 13] * dprod[db & 0x1fffu];
                }
#pragma unroll 2
                for (; q < na; q++) {
                    const unsigned da = ta[32 * q];
                    acc_a += ctab[da >> 13] * dprod[da & 0x1fffu];
                }
                {
                    const unsigned sg = seg[gI * 32 + lane];
                    const unsigned slot = sg >> 16, row = sg & 0xffu, colx = (sg >> 8) & 0xffu;
                    if (row != 0xffu) {
                        if (slot == 0xffffu) blk[row * ld + colx] = -acc_a;
                        else part[slot] = acc_a;
                    }
                }
                if (two) {
                    const unsigned sg = seg[gJ * 32 + lane];
                    const unsigned slot = sg >> 16, row = sg & 0xffu, colx = (sg >> 8) & 0xffu;
                    if (row != 0xffu) {
                        if (slot == 0xffffu) blk[row * ld + colx] = -acc_b;
                        else part[slot] = acc_b;
                    }
                }
            }
        }
        __syncthreads();
        // ---- phase C: split entries (fixed-order sum of their partials)
        for (int m = tid; m < A.net.n_multi; m += nt) {
            const uint2 me = multi[m];                 // x = row | col << 16, y = first slot | n << 16
            const int s0 = (int)(me.y & 0xffff), n = (int)(me.y >> 16);
            double acc = 0.0;
            for (int q = 0; q < n; q++) acc += part[s0 + q];
            blk[(me.x & 0xffff) * ld + (me.x >> 16)] = -acc;
        }
        __syncthreads();
        // ---- diagonal: c0 + negJ_ss - transport
        if (tid < ld && !(dbg & 8)) {
            const int i = tid;
            if (i >= ni) {   // padding: decoupled identity rows keep the padded block invertible
                blk[i * ld + i] = 1.0;
            } else {
                double d = c0 + (dflag[i] ? blk[i * ld + i] : 0.0);
                d -= tE;
                d -= eA;
                if (md) d -= tA;
                if (A.atm.use_botflux && j == 0) d -= tV;
                if (A.fix_mask && A.fix_mask[base + i]) {   // op.py:2903-2906: row -> 1/(r h) e_i
                    for (int t = 0; t < ni; t++) blk[i * ld + t] = 0.0;
                    d = c0;
                }
                blk[i * ld + i] = d;
            }
        }
        // every thread's generic-proxy writes of the block (gather, split entries, diagonal) -> visible to the async proxy, then the barrier
        if (bulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        // ---- phase D: the finished block leaves as ONE asynchronous TMA bulk store (shared -> global); nothing is re-zeroed (see the
        // set-up).  The copy is drained at the top of the next layer, behind that layer's wait for its prefetched rows.
        double *Dg = A.D + ((size_t)col * nz + j) * ld * ld;
        if (dbg & 1) {
        } else if (bulk) {
            if (tid == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(Dg), "r"(blk_s), "r"(blk_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if ((ld & 1) == 0) {
            for (int q = tid; q < ld * ld / 2; q += nt) reinterpret_cast<double2 *>(Dg)[q] = reinterpret_cast<const double2 *>(blk)[q];
        } else {
            for (int q = tid; q < ld * ld; q += nt) Dg[q] = blk[q];
        }
        // (the barrier at the top of the next iteration orders the stored block and the prefetched rows)
    }
    if (bulk && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the last store
}

// ------------------------------------------------------------------------------------------------------------------
// Fused assembly: the producer side of factor_kernel<NIP, MINB, true> (vk_factor_dev.cuh).  The NPROD otherwise idle warps of the block
// run the layer loop of lhs_ml_kernel - same tables, same phases, same expression order, hence the same D bit for bit - but write the
// finished block into the factor kernel's own D buffer in shared memory, one layer ahead of the column warps:
//     FREE  (column warps arrive at panel 1 of layer j : D_j, up_{j-1}, dn_j consumed)  ->  producers assemble D_{j+1}, up_j, dn_{j+1}
//     FULL  (producers arrive)  ->  the Schur update of layer j+1 may read them.
// up / dn still go to HBM (the solve sweeps read them: 2 x NIP doubles per layer); D goes to HBM only for the columns that are going
// to be refined (resid_kernel reads it) - store_D: 0 never, 1 always, 2 columns with dt >= dt_min.
struct LhsProdSmem {    // offsets in doubles behind the factor kernel's own shared memory
    int kz, ym, y0, yp, dprod, part, misc, trs, upprev, tab, total_bytes;
};
static inline LhsProdSmem lhs_prod_layout(const NetDev &n, int ld)
{
    LhsProdSmem L;
    int o = 0;
    L.kz = o; o += n.nr + 2;
    L.ym = o; o += n.ni + 2; L.y0 = o; o += n.ni + 2; L.yp = o; o += n.ni + 2;
    L.dprod = o; o += n.n_uniq + 2;
    L.part = o; o += n.n_part + 2;
    L.misc = o; o += 16 + 16;       // [0..3] layer sums, [8..15] coefficient table, [16..31] 128 diagonal-in-pattern flags
    L.trs = o; o += 6 * ld;         // per species: eA, tA, tV, tE, u, l of the layer being assembled
    L.upprev = o; o += ld;          // up_{j-1}
    o += o & 1;
    L.tab = o;
    size_t bytes = sizeof(double) * (size_t)o + sizeof(uint2) * (n.n_grp + 1) + sizeof(uint2) * (n.n_multi + 1) +
                   sizeof(unsigned) * (n.n_uniq + 2) + sizeof(unsigned) * ((size_t)n.n_grp * 32) + sizeof(unsigned short) * (n.n_tt + 8);
    L.total_bytes = (int)((bytes + 15) & ~(size_t)15);
    return L;
}
struct LhsProdArgs {
    LhsArgs L;
    LhsProdSmem SL;
    int store_D;
    double dt_min;
};

#ifdef VK_PROD_TRACE
__device__ long long g_prod_trace[16];
#define PTRACE(ev) do { if (col == 0 && j == 7 && tid == 0) g_prod_trace[ev] = clock64(); } while (0)
#else
#define PTRACE(ev) do { } while (0)
#endif

template <int NIP, class PA>
__device__ void lhs_produce(const PA &PR, int col, int nz, double *blk, double *updn, double *sm, int pw, int lane)
{
    using C = FactorCfg<NIP>;
    constexpr int PNT = C::NPROD * 32, ld = NIP;
    const LhsArgs &A = PR.L;
    const LhsProdSmem &SL = PR.SL;
    const int ni = A.net.ni, nr = A.net.nr;
    const int tid = pw * 32 + lane;
    double *kz = sm + SL.kz, *ym = sm + SL.ym, *y0 = sm + SL.y0, *yp = sm + SL.yp, *dprod = sm + SL.dprod, *part = sm + SL.part;
    double *ysum = sm + SL.misc, *ctab = sm + SL.misc + 8, *trs = sm + SL.trs, *upprev = sm + SL.upprev;
    uint2 *grp = reinterpret_cast<uint2 *>(sm + SL.tab);
    uint2 *multi = grp + (A.net.n_grp + 1);
    unsigned *uq = reinterpret_cast<unsigned *>(multi + (A.net.n_multi + 1));
    unsigned *seg = uq + (A.net.n_uniq + 2);
    unsigned short *tt = reinterpret_cast<unsigned short *>(seg + (size_t)A.net.n_grp * 32);
    unsigned char *dflag = reinterpret_cast<unsigned char *>(sm + SL.misc + 16);
    auto psync = [&]() { if (C::NPROD == 1) __syncwarp(); else bar_sync<VK_BAR_PROD, (PNT > 0 ? PNT : 32)>(); };

    // The atmosphere-only stencil pieces of the layer (AtmPre: 9 arrays + Q of the layer below + 10 layer scalars) are prefetched too:
    // read from global memory at the point of use they are a dozen SERIALISED DRAM round trips per layer (the loads sit behind the
    // boundary / moldiff / settling branches), which 16 co-resident warps hide in lhs_ml_kernel and two producer warps cannot.  They
    // land in the dprod region - dead between phase B of one layer and phase A of the next - and are consumed (transport part) BEFORE
    // phase A overwrites it.
    double *pre = dprod;                                   // [10][ni] Qm Q QB QC TA TB TC SA SB SC, then LS[10]
    const AtmPre &PG = A.atm.pre;
    auto prefetch = [&](int j) {      // k row, the three y rows and the stencil pieces of layer j -> shared memory (cp.async, 8 bytes each)
        const double *kg = A.k + col * A.k_cs + (size_t)j * (nr + 1);
        for (int i = tid; i <= nr; i += PNT) cp_async8(kz + i, kg + i);
        const size_t base = ((size_t)col * nz + j) * ni;
        const size_t pb0 = (size_t)col * A.atm.pre_cs + (size_t)j * ni;
        for (int i = tid; i < ni; i += PNT) {
            cp_async8(y0 + i, A.y + base + i);
            if (j > 0) cp_async8(ym + i, A.y + base - ni + i);
            if (j < nz - 1) cp_async8(yp + i, A.y + base + ni + i);
            if (j > 0) cp_async8(pre + i, PG.Q + pb0 - ni + i);
            cp_async8(pre + ni + i, PG.Q + pb0 + i);
            cp_async8(pre + 2 * ni + i, PG.QB + pb0 + i);
            cp_async8(pre + 3 * ni + i, PG.QC + pb0 + i);
            cp_async8(pre + 4 * ni + i, PG.TA + pb0 + i);
            cp_async8(pre + 5 * ni + i, PG.TB + pb0 + i);
            cp_async8(pre + 6 * ni + i, PG.TC + pb0 + i);
            cp_async8(pre + 7 * ni + i, PG.SA + pb0 + i);
            cp_async8(pre + 8 * ni + i, PG.SB + pb0 + i);
            cp_async8(pre + 9 * ni + i, PG.SC + pb0 + i);
        }
        if (tid < 10) cp_async8(pre + 10 * ni + tid, PG.LS + ((size_t)col * (A.atm.pre_cs ? nz : 0) + j) * 10 + tid);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(0);
    for (int i = tid; i < A.net.n_grp; i += PNT) grp[i] = A.net.jac_grp[i];
    for (int i = tid; i < A.net.n_grp * 32; i += PNT) seg[i] = A.net.jac_seg4[i];
    for (int i = tid; i < A.net.n_multi; i += PNT) multi[i] = A.net.jac_multi[i];
    for (int i = tid; i < A.net.n_uniq; i += PNT) uq[i] = A.net.jac_uniq[i];
    {
        const unsigned *src = reinterpret_cast<const unsigned *>(A.net.jac_tt);
        unsigned *dst = reinterpret_cast<unsigned *>(tt);
        for (int i = tid; i < (A.net.n_tt + 1) / 2; i += PNT) dst[i] = src[i];
    }
    if (tid == 0) y0[ni + 1] = 1.0;
    for (int q = tid; q < ld * ld; q += PNT) blk[q] = 0.0;         // zeroed ONCE: every layer assigns the same pattern entries (see lhs_ml_kernel)
    for (int i = tid; i < 128; i += PNT) dflag[i] = 0;
    for (int i = tid; i < ld; i += PNT) upprev[i] = 0.0;
    if (tid < 8) ctab[tid] = c_jac_coef[tid];
    psync();
    for (int i = tid; i < A.net.n_grp * 32; i += PNT) {
        const unsigned sg = seg[i];
        const unsigned row = sg & 0xffu, colx = (sg >> 8) & 0xffu;
        if (row != 0xffu && row == colx) dflag[row] = 1;
    }
    for (int i = tid; i < A.net.n_multi; i += PNT) {
        const uint2 me = multi[i];
        if ((me.x & 0xffff) == (me.x >> 16)) dflag[me.x & 0xffff] = 1;
    }
    const unsigned blk_s = (unsigned)__cvta_generic_to_shared(blk);
    const unsigned blk_bytes = (unsigned)(ld * ld * sizeof(double));
    const AtmLayer L = atm_at(A.atm, col);
    const double *dzi = L.dzi;
    const int md = A.atm.use_moldiff, st = A.atm.use_settling && A.atm.use_moldiff;
    const int vmm = A.atm.use_vm_mol;
    const double rr = 1. + 1. / sqrt(2.);
    const double dtc = A.dt[col];
    const double c0 = 1. / (rr * dtc);
    AtmPre P;                                              // the prefetched copies, indexed by species
    P.Q = pre + ni; P.QB = pre + 2 * ni; P.QC = pre + 3 * ni; P.TA = pre + 4 * ni; P.TB = pre + 5 * ni; P.TC = pre + 6 * ni;
    P.SA = pre + 7 * ni; P.SB = pre + 8 * ni; P.SC = pre + 9 * ni; P.LS = pre + 10 * ni;
    const double *Qm = pre;
    const bool storeD = A.D && (PR.store_D == 1 || (PR.store_D == 2 && dtc >= PR.dt_min));
    double *eAs = trs, *tAs = trs + ld, *tVs = trs + 2 * ld, *tEs = trs + 3 * ld, *us = trs + 4 * ld, *ls_ = trs + 5 * ld;

    for (int j = 0; j < nz; j++) {
        PTRACE(0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (tid == 0) y0[ni] = A.atm.M[col * A.atm.csz + j];
        psync();
        PTRACE(1);
        // ---- the three layer sums (last producer warp, numpy association)
        if (pw == C::NPROD - 1) {
            const int q = lane >> 3;
            const int jj = j - 1 + q;
            const bool act = (q < 3) && jj >= 0 && jj < nz;
            const double *row = (q == 0) ? ym : ((q == 1) ? y0 : yp);
            double sres;
            if (A.atm.n_gas_lhs > 0) {
                sres = 0.0;
                if (act && (lane & 7) == 0) sres = row_sum(row, ni, A.atm.n_gas_lhs, A.atm.gas_indx_lhs, nullptr);
            } else {
                sres = np_pairwise_group8(row, ni, act);
            }
            if (act && (lane & 7) == 0) ysum[q] = sres;
        }
        psync();
        PTRACE(2);
        const size_t base = ((size_t)col * nz + j) * ni;
        const size_t vbase = ((size_t)col * nz + j) * ld;
        // ---- transport part of the diagonal and the couplings (op.py:1998-2040), expression by expression as in lhs_ml_kernel
        for (int i = tid; i < ld; i += PNT) {
            double eA = 0.0, tA = 0.0, tV = 0.0, tE = 0.0, u = 0.0, l = 0.0;
            if (i < ni) {
                const double ys0 = ysum[1], ysm = ysum[0], ysp = ysum[2];
                const double *ls = P.LS;
                const size_t pb = (size_t)i;
                double eB = 0.0, eC = 0.0;
                if (j == 0) {
                    eA = ls[0] * (ysp + ys0) / (2. * ys0) + ls[5];
                    eB = ls[3] * (ysp + ys0) / (2. * ysp) + ls[6];
                    u -= eB;
                    if (md) {
                        double ta = P.QC[pb] * (ysp + ys0) / (2. * ys0);
                        double tb = P.QB[pb] * (ysp + ys0) / (2. * ysp);
                        if (vmm) {
                            double tx = 0.0;
                            vm_lhs_adv(P, pb, 0, st, ta, tb, tx);
                        } else {
                            ta = ta + P.TA[pb];
                            tb = tb + P.TB[pb];
                            if (st) {
                                ta = ta - P.SA[pb];
                                tb = tb - P.SB[pb];
                            }
                        }
                        tA = ta;
                        u -= tb;
                    }
                    if (A.atm.use_botflux) tV = -1. * L.bot_vdep[i] / dzi[0];
                } else if (j == nz - 1) {
                    if (vmm && A.atm.n_diff_esc > 0) tE = vm_diff_lim(A.atm, L, i, y0[i]);
                    eA = ls[0] * (ysm + ys0) / (2. * ys0) + ls[5];
                    eC = ls[4] * (ysm + ys0) / (2. * ysm) + ls[7];
                    l -= eC;
                    if (md) {
                        double ta = P.QB[pb] * (ys0 + ysm) / (2. * ys0);
                        double tc = P.QC[pb] * (ys0 + ysm) / (2. * ysm);
                        if (vmm) {
                            double tx = 0.0;
                            vm_lhs_adv(P, pb, 2, st, ta, tx, tc);
                        } else {
                            ta = ta - P.TA[pb];
                            tc = tc - P.TC[pb];
                            if (st) {
                                ta = ta + P.SA[pb];
                                tc = tc + P.SC[pb];
                            }
                        }
                        tA = ta;
                        l -= tc;
                    }
                } else {
                    eA = ls[8] * (ls[1] * (ysp + ys0) / 2. + ls[2] * (ysm + ys0) / 2.) / ys0 + ls[5];
                    eB = ls[9] * (ls[1] * (ysp + ys0) / (2. * ysp)) + ls[6];
                    eC = ls[9] * (ls[2] * (ysm + ys0) / (2. * ysm)) + ls[7];
                    u -= eB;
                    l -= eC;
                    if (md) {
                        double ta = ls[8] * (P.Q[pb] * (ysp + ys0) / 2. + Qm[i] * (ysm + ys0) / 2.) / ys0;
                        double tb = ls[9] * (P.Q[pb] * (ysp + ys0) / (2. * ysp));
                        double tc = ls[9] * (Qm[i] * (ysm + ys0) / (2. * ysm));
                        if (vmm) {
                            vm_lhs_adv(P, pb, 1, st, ta, tb, tc);
                        } else {
                            ta = ta + P.TA[pb];
                            tb = tb + P.TB[pb];
                            tc = tc - P.TC[pb];
                            if (st) {
                                ta = ta - P.SA[pb];
                                tb = tb - P.SB[pb];
                                tc = tc + P.SC[pb];
                            }
                        }
                        tA = ta;
                        u -= tb;
                        l -= tc;
                    }
                }
                if (A.fix_mask && A.fix_mask[base + i]) { u = 0.0; l = 0.0; }
            }
            eAs[i] = eA; tAs[i] = tA; tVs[i] = tV; tEs[i] = tE; us[i] = u; ls_[i] = l;
            A.up[vbase + i] = u;
            A.dn[vbase + i] = l;
        }
        psync();                                   // the stencil pieces are consumed: the products may overwrite them
        PTRACE(3);
        // ---- phase A: distinct products k_r y_a y_b y_c
#pragma unroll 8
        for (int u = tid; u < A.net.n_uniq; u += PNT) {
            const unsigned d = uq[u];
            double x = kz[d & 0x7ffu];
            x = x * y0[(d >> 11) & 0x7fu];
            x = x * y0[(d >> 18) & 0x7fu];
            x = x * y0[(d >> 25) & 0x7fu];
            dprod[u] = x;
        }
        if (tid == 0) {
            dprod[A.net.n_uniq] = 0.0;      // the padding product
            if (storeD) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous block has left shared memory
        }
        // everything above touched only the producers' own buffers; from here on the block buffer is written: wait until the column
        // warps have consumed D_{j-1}, up_{j-2}, dn_{j-1} (they arrive one panel into layer j-1)
        PTRACE(4);
        if (j > 0) bar_sync<VK_BAR_FREE, C::NFEED>();
        psync();
        PTRACE(5);
        // ---- phase B: groups of 32 segments (see lhs_ml_kernel), EIGHT adjacent groups per warp and turn: the producers are few warps and
        // share the SM's load/store unit with the column warps (measured: 165 clk per term step with 4 chains), so the per-term chain
        // (descriptor -> product -> multiply -> add) is hidden by independent chains of the same lane, not by other warps.  The +-1/2/4/3
        // coefficient is decoded arithmetically (one shared-memory load less per term); products by +-1, 2, 4 are exact, and x * 3.0 is the
        // same rounding as the table version
        auto coef = [](unsigned d) -> double {
            const unsigned c = d >> 13;                        // codes: 1, -1, 2, -2, 4, -4, 3, -3
            const double m = (c >> 1) == 0 ? 1.0 : ((c >> 1) == 1 ? 2.0 : ((c >> 1) == 2 ? 4.0 : 3.0));
            return (c & 1u) ? -m : m;
        };
        for (int gI = 8 * pw; gI < A.net.n_grp; gI += 8 * C::NPROD) {
            const int ng = min(8, A.net.n_grp - gI);
            const unsigned short *tp[8];
            int nn[8];
            double acc[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint2 gg = grp[gI + (u < ng ? u : 0)];
                tp[u] = tt + gg.x + lane;
                nn[u] = (u < ng) ? (int)gg.y : 0;          // nn[0] >= nn[1] >= ... (groups sorted by decreasing length)
                acc[u] = 0.0;
            }
            int q = 0;
            // staged loops: all eight chains while the shortest group lasts, then seven, ... (warp-uniform trip counts)
#define VK_GATHER_STAGE(NCH)                                                          \
            for (; q < nn[NCH - 1]; q++) {                                            \
                unsigned dd[NCH];                                                     \
                _Pragma("unroll") for (int u = 0; u < NCH; u++) dd[u] = tp[u][32 * q]; \
                _Pragma("unroll") for (int u = 0; u < NCH; u++) acc[u] += coef(dd[u]) * dprod[dd[u] & 0x1fffu]; \
            }
            VK_GATHER_STAGE(8) VK_GATHER_STAGE(7) VK_GATHER_STAGE(6) VK_GATHER_STAGE(5)
            VK_GATHER_STAGE(4) VK_GATHER_STAGE(3) VK_GATHER_STAGE(2) VK_GATHER_STAGE(1)
#undef VK_GATHER_STAGE
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (u < ng) {
                    const unsigned sg = seg[(gI + u) * 32 + lane];
                    const unsigned slot = sg >> 16, row = sg & 0xffu, colx = (sg >> 8) & 0xffu;
                    if (row != 0xffu) {
                        if (slot == 0xffffu) blk[row * ld + colx] = -acc[u];
                        else part[slot] = acc[u];
                    }
                }
            }
        }
        psync();
        PTRACE(6);
        if (j + 1 < nz) prefetch(j + 1);           // k / y rows of this layer are consumed (products formed, layer sums taken, y0 read by tE)
        // ---- phase C: split entries (fixed-order sum of their partials)
        for (int m = tid; m < A.net.n_multi; m += PNT) {
            const uint2 me = multi[m];
            const int s0 = (int)(me.y & 0xffff), n = (int)(me.y >> 16);
            double acc = 0.0;
            for (int q = 0; q < n; q++) acc += part[s0 + q];
            blk[(me.x & 0xffff) * ld + (me.x >> 16)] = -acc;
        }
        psync();
        PTRACE(7);
        // ---- diagonal: c0 + negJ_ss - transport; the couplings the Schur update of this layer reads: up_{j-1}, dn_j
        for (int i = tid; i < ld; i += PNT) {
            if (i >= ni) {
                blk[i * ld + i] = 1.0;
            } else {
                double d = c0 + (dflag[i] ? blk[i * ld + i] : 0.0);
                d -= tEs[i];
                d -= eAs[i];
                if (md) d -= tAs[i];
                if (A.atm.use_botflux && j == 0) d -= tVs[i];
                if (A.fix_mask && A.fix_mask[base + i]) {
                    for (int t = 0; t < ni; t++) blk[i * ld + t] = 0.0;
                    d = c0;
                }
                blk[i * ld + i] = d;
            }
            updn[i] = upprev[i];
            updn[ld + i] = ls_[i];
            upprev[i] = us[i];
        }
        if (storeD) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __threadfence_block();
        psync();
        if (storeD && tid == 0) {
            double *Dg = A.D + ((size_t)col * nz + j) * ld * ld;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(Dg), "r"(blk_s), "r"(blk_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        PTRACE(8);
        bar_arrive<VK_BAR_FULL, C::NFEED>();
        if (j > 0) {       // the block-wide barrier that ends factor layer j-1 (D_j is complete by then)
            if (__syncthreads_or(0)) { if (storeD && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); return; }
        }
        PTRACE(9);
    }
    __syncthreads_or(0);   // ... and the one of the last layer
    if (storeD && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int NIP, int MINB>
static int launch_factor_fused_t(vk_column *c, const LhsProdArgs &pa, double *F, int *status)
{
    using C = FactorCfg<NIP>;
    if (C::NPROD == 0) return VK_ERR_UNSUPPORTED;          // no idle warps in this block layout (NIP = 120): two-kernel path
    FactorArgs a{c->nz, c->ni, nullptr, nullptr, nullptr, F, status, c->act, nullptr, nullptr, 0};
    const size_t smem = C::SMEM + 16 + (size_t)pa.SL.total_bytes;
    if (smem > 227 * 1024) return VK_ERR_UNSUPPORTED;
    { int rc = ensure_smem((const void *)factor_kernel<NIP, MINB, true, LhsProdArgs>, c->net->device, smem); if (rc) return rc; }
    if (getenv("VK_DEBUG")) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, factor_kernel<NIP, MINB, true, LhsProdArgs>, C::NT, smem);
        fprintf(stderr, "vulcan_b200: fused factor kernel NIP %d: %zu bytes of shared memory per block, %d block(s) per SM\n", NIP, smem, nb);
    }
    factor_kernel<NIP, MINB, true, LhsProdArgs><<<c->ncol, C::NT, smem, c->stream>>>(a, pa);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

// lhs assembly + block-tridiagonal factorisation in ONE kernel.  Returns VK_ERR_UNSUPPORTED (no error message set) when the tables of
// this network do not fit next to the factor kernel's buffers: the caller then takes the two-kernel path.
int launch_factor_fused(vk_column *c, const double *y_dev, const double *dt_dev, double *D_out, double *up_out, double *dn_out, double *F,
                        int *status, int store_D)
{
    LhsProdArgs pa;
    LhsArgs &a = pa.L;
    a.net = c->net->d; a.atm = c->atm; a.nz = c->nz; a.y = y_dev; a.k = c->k; a.k_cs = c->k_cs; a.dt = dt_dev;
    a.D = D_out; a.up = up_out; a.dn = dn_out; a.ld = c->nip; a.fix_mask = c->opts.fix_mask; a.act = c->act;
    if (!a.net.lhs_ml_ok) return VK_ERR_UNSUPPORTED;
    pa.SL = lhs_prod_layout(a.net, c->nip);
    pa.store_D = store_D; pa.dt_min = c->opts.refine_dt_min;
    switch (c->nip) {
        case 48: return launch_factor_fused_t<48, 2>(c, pa, F, status);
        case 72: return launch_factor_fused_t<72, 2>(c, pa, F, status);
        case 96: return launch_factor_fused_t<96, 1>(c, pa, F, status);
        case 120: return launch_factor_fused_t<120, 1>(c, pa, F, status);
        default: return VK_ERR_UNSUPPORTED;
    }
}

// ------------------------------------------------------------------------------------------------------------------
int launch_rhs(vk_column *c, const double *y_dev, double *out_sum, double *out_chem, double *out_diff,
               const double *k1_for_rhs2, const double *dt_dev)
{
    RhsArgs a;
    a.net = c->net->d; a.atm = c->atm; a.nz = c->nz; a.y = y_dev; a.k = c->k; a.k_cs = c->k_cs;
    a.k1 = k1_for_rhs2; a.dt = dt_dev; a.yk2_out = k1_for_rhs2 ? c->yk2 : nullptr;
    a.out_sum = out_sum; a.out_chem = out_chem; a.out_diff = out_diff;
    a.fix_mask = c->opts.fix_mask; a.act = c->act;
    a.fast = ((c->opts.rhs_order == 0 || c->opts.rhs_order == 3) && a.net.rhs_flat_ok) ? 1 : 0;
    // chemistry through the emitted kernel of this network when the library has one (reference summation order, bit-identical to the
    // table-driven reference-order path); VK_EMIT=0 or rhs_order = 2 / 3 keep the table-driven kernels
    const char *ee = getenv("VK_EMIT");          // (read per call: the parity tests switch it inside one process)
    const int emit_env = ee ? atoi(ee) : 1;
    a.chem_in = nullptr;
    // emitted path: batches that share their rate coefficients (block = one layer of 128 columns, vk_emit_rt.cuh)
    if (c->net->emit && emit_env && c->opts.rhs_order < 2 && (c->k_cs == 0 || c->k_static_shared) && c->ncol >= 32) {
        const size_t nl = (size_t)c->ncol * c->nz;
        if (!c->chem_tmp) {
            VK_CUDA(cudaMalloc((void **)&c->chem_tmp, sizeof(double) * nl * c->ni));
            VK_CUDA(cudaMalloc((void **)&c->ysum_tmp, sizeof(double) * nl));
            VK_CUDA(cudaMalloc((void **)&c->scal_tmp, sizeof(LayerScal) * nl));
        }
        int rc = launch_chem_emitted(c, y_dev, k1_for_rhs2, c->chem_tmp, c->ysum_tmp, k1_for_rhs2 ? c->yk2 : nullptr);
        if (rc) return rc;
        StencilArgs sa;
        sa.atm = c->atm; sa.nz = c->nz; sa.ni = c->ni; sa.ncol = c->ncol;
        sa.y = k1_for_rhs2 ? c->yk2 : y_dev; sa.k1 = k1_for_rhs2; sa.dt = dt_dev;
        sa.chem = c->chem_tmp; sa.ysum = c->ysum_tmp; sa.S = static_cast<LayerScal *>(c->scal_tmp);
        sa.out_sum = out_sum; sa.out_chem = out_chem; sa.out_diff = out_diff; sa.fix_mask = c->opts.fix_mask; sa.act = c->act;
        layer_scal_kernel<<<(unsigned)((nl + 127) / 128), 128, 0, c->stream>>>(sa);
        rhs_stencil_kernel<<<(unsigned)((nl * c->ni + 255) / 256), 256, 0, c->stream>>>(sa);
        VK_CUDA(cudaGetLastError());
        return VK_OK;
    }
    const RhsWarpSmem SL = rhs_warp_layout(c->ni, c->nr, a.fast ? a.net.rhs_n_seg : 0);
    const size_t smem = sizeof(double) * (size_t)RHS_WPB * SL.total + sizeof(uchar4) * (c->nr + 2) +
                        sizeof(unsigned short) * (std::max(a.net.n_rhs, 32 * a.net.rhs_flat_T) + 8) + 16;
    { int rc = ensure_smem((const void *)rhs_warp_kernel, c->net->device, smem); if (rc) return rc; }
    const int n_layers = c->ncol * c->nz;
    rhs_warp_kernel<<<(n_layers + RHS_WPB - 1) / RHS_WPB, RHS_WPB * 32, smem, c->stream>>>(a, n_layers);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// lhs behind an EMITTED Jacobian kernel (vk_emit.cu: -J as dense rows in D, layer sums in ysum): one thread per (column, layer, row of the
// padded block) adds what lhs_jac_tot adds around the chemical Jacobian (op.py:1973-2444) - c0 = 1/(r dt) and the transport terms on the
// diagonal, the couplings up / dn, identity rows for fixed species and for the padding - same expressions, same order as lhs_ml_kernel.
struct LhsDiagArgs {
    AtmDev atm;
    int nz, ni, ncol, ld;
    const double *y, *dt, *ysum;
    double *D, *up, *dn;
    const unsigned char *fix_mask;
    const int *act;
};
__global__ void __launch_bounds__(256) lhs_diag_kernel(LhsDiagArgs A)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ni = A.ni, nz = A.nz, ld = A.ld;
    if (e >= (size_t)A.ncol * nz * ld) return;
    const int lay = (int)(e / ld), i = (int)(e - (size_t)lay * ld);
    const int col = lay / nz, j = lay - col * nz;
    if (A.act && !A.act[col]) return;
    double *Drow = A.D + ((size_t)lay * ld + i) * ld;
    const size_t vbase = (size_t)lay * ld;
    // padding rows ni .. ld-1: decoupled identity rows keep the padded block invertible; thread i writes column i of each (coalesced)
    for (int r = ni; r < ld; r++) A.D[((size_t)lay * ld + r) * ld + i] = (r == i) ? 1.0 : 0.0;
    if (i >= ni) {
        A.up[vbase + i] = 0.0;
        A.dn[vbase + i] = 0.0;
        return;
    }
    const double rr = 1. + 1. / sqrt(2.);
    const double c0 = 1. / (rr * A.dt[col]);
    const AtmLayer L = atm_at(A.atm, col);
    const AtmPre &P = A.atm.pre;
    const double *dzi = L.dzi;
    const int md = A.atm.use_moldiff, st = A.atm.use_settling && A.atm.use_moldiff;
    const int vmm = A.atm.use_vm_mol;
    const size_t base = (size_t)lay * ni;
    const double ys0 = A.ysum[lay], ysm = (j > 0) ? A.ysum[lay - 1] : 0.0, ysp = (j < nz - 1) ? A.ysum[lay + 1] : 0.0;
    const double *ls = A.atm.pre.LS + ((size_t)col * (A.atm.pre_cs ? nz : 0) + j) * 10;
    const size_t pb = ((size_t)col * A.atm.pre_cs + (size_t)j * ni) + i;
    double eA = 0.0, tA = 0.0, tV = 0.0, tE = 0.0;
    double eB = 0.0, eC = 0.0, u = 0.0, l = 0.0;
    if (j == 0) {
        eA = ls[0] * (ysp + ys0) / (2. * ys0) + ls[5];
        eB = ls[3] * (ysp + ys0) / (2. * ysp) + ls[6];
        u -= eB;
        if (md) {
            double ta = P.QC[pb] * (ysp + ys0) / (2. * ys0);
            double tb = P.QB[pb] * (ysp + ys0) / (2. * ysp);
            if (vmm) {
                double tx = 0.0;
                vm_lhs_adv(P, pb, 0, st, ta, tb, tx);
            } else {
                ta = ta + P.TA[pb];
                tb = tb + P.TB[pb];
                if (st) {
                    ta = ta - P.SA[pb];
                    tb = tb - P.SB[pb];
                }
            }
            tA = ta;
            u -= tb;
        }
        if (A.atm.use_botflux) tV = -1. * L.bot_vdep[i] / dzi[0];
    } else if (j == nz - 1) {
        if (vmm && A.atm.n_diff_esc > 0) tE = vm_diff_lim(A.atm, L, i, A.y[base + i]);
        eA = ls[0] * (ysm + ys0) / (2. * ys0) + ls[5];
        eC = ls[4] * (ysm + ys0) / (2. * ysm) + ls[7];
        l -= eC;
        if (md) {
            double ta = P.QB[pb] * (ys0 + ysm) / (2. * ys0);
            double tc = P.QC[pb] * (ys0 + ysm) / (2. * ysm);
            if (vmm) {
                double tx = 0.0;
                vm_lhs_adv(P, pb, 2, st, ta, tx, tc);
            } else {
                ta = ta - P.TA[pb];
                tc = tc - P.TC[pb];
                if (st) {
                    ta = ta + P.SA[pb];
                    tc = tc + P.SC[pb];
                }
            }
            tA = ta;
            l -= tc;
        }
    } else {
        eA = ls[8] * (ls[1] * (ysp + ys0) / 2. + ls[2] * (ysm + ys0) / 2.) / ys0 + ls[5];
        eB = ls[9] * (ls[1] * (ysp + ys0) / (2. * ysp)) + ls[6];
        eC = ls[9] * (ls[2] * (ysm + ys0) / (2. * ysm)) + ls[7];
        u -= eB;
        l -= eC;
        if (md) {
            double ta = ls[8] * (P.Q[pb] * (ysp + ys0) / 2. + P.Q[pb - ni] * (ysm + ys0) / 2.) / ys0;
            double tb = ls[9] * (P.Q[pb] * (ysp + ys0) / (2. * ysp));
            double tc = ls[9] * (P.Q[pb - ni] * (ysm + ys0) / (2. * ysm));
            if (vmm) {
                vm_lhs_adv(P, pb, 1, st, ta, tb, tc);
            } else {
                ta = ta + P.TA[pb];
                tb = tb + P.TB[pb];
                tc = tc - P.TC[pb];
                if (st) {
                    ta = ta - P.SA[pb];
                    tb = tb - P.SB[pb];
                    tc = tc + P.SC[pb];
                }
            }
            tA = ta;
            u -= tb;
            l -= tc;
        }
    }
    const bool fixed = A.fix_mask && A.fix_mask[base + i];
    if (fixed) { u = 0.0; l = 0.0; }
    A.up[vbase + i] = u;
    A.dn[vbase + i] = l;
    double d = c0 + Drow[i];
    d -= tE;
    d -= eA;
    if (md) d -= tA;
    if (A.atm.use_botflux && j == 0) d -= tV;
    if (fixed) {                 // op.py:2903-2906: row -> 1/(r h) e_i
        for (int t = 0; t < ni; t++) Drow[t] = 0.0;
        d = c0;
    }
    Drow[i] = d;
}

int launch_lhs(vk_column *c, const double *y_dev, const double *dt_dev, int ld, double *D_out, double *up_out, double *dn_out)
{
    LhsArgs a;
    a.net = c->net->d; a.atm = c->atm; a.nz = c->nz; a.y = y_dev; a.k = c->k; a.k_cs = c->k_cs; a.dt = dt_dev;
    a.D = D_out; a.up = up_out; a.dn = dn_out; a.ld = ld; a.fix_mask = c->opts.fix_mask; a.act = c->act;
    // emitted path (batches that share their rate coefficients): -J as straight-line code, then the elementwise diagonal / coupling kernel
    const char *ej = getenv("VK_EMIT_JAC");      // (read per call: the parity tests switch it inside one process)
    const int emit_env = ej ? atoi(ej) : 1;
    if (emit_env && emit_has_jac(c->net->emit) && (c->k_cs == 0 || c->k_static_shared) && c->ncol >= 32 && ld == c->nip && !getenv("VK_LHS_DEBUG")) {
        const size_t nl = (size_t)c->ncol * c->nz;
        if (!c->ysum_lhs_tmp) VK_CUDA(cudaMalloc((void **)&c->ysum_lhs_tmp, sizeof(double) * nl));
        int rc = launch_jac_emitted(c, y_dev, D_out, c->ysum_lhs_tmp);
        if (rc) return rc;
        LhsDiagArgs da;
        da.atm = c->atm; da.nz = c->nz; da.ni = c->ni; da.ncol = c->ncol; da.ld = ld;
        da.y = y_dev; da.dt = dt_dev; da.ysum = c->ysum_lhs_tmp; da.D = D_out; da.up = up_out; da.dn = dn_out;
        da.fix_mask = c->opts.fix_mask; da.act = c->act;
        lhs_diag_kernel<<<(unsigned)((nl * ld + 255) / 256), 256, 0, c->stream>>>(da);
        VK_CUDA(cudaGetLastError());
        return VK_OK;
    }
    if (!a.net.lhs_ml_ok) {
        set_error("network exceeds the packing limits of the Jacobian kernel (nr < 2048, ni < 127, < 8191 distinct products, coefficients in +-{1,2,3,4})");
        return VK_ERR_UNSUPPORTED;
    }
    const LhsMlSmem SL = lhs_ml_layout(a.net, ld);
    if (SL.total_bytes > 227 * 1024) { set_error("Jacobian kernel tables do not fit the shared memory of one SM"); return VK_ERR_UNSUPPORTED; }
    // threads per block: 256 (8 warps, <= 128 registers) or 512 (16 warps at <= 64 registers: twice the warps per SM to hide the latency
    // of the gather's dependent shared-memory loads); VK_LHS_NT overrides
    static int nthr = -1;
    if (nthr < 0) { const char *e = getenv("VK_LHS_NT"); nthr = e ? atoi(e) : 256; if (nthr != 512 && nthr != 384) nthr = 256; }
    { int rc = ensure_smem(nthr == 512 ? (const void *)lhs_ml_kernel<512> : (nthr == 384 ? (const void *)lhs_ml_kernel<384> : (const void *)lhs_ml_kernel<256>), c->net->device, (size_t)SL.total_bytes); if (rc) return rc; }
    // layers per block: amortise the table staging (40 KB per block) but keep >= ~8 blocks per SM in the grid
    int lpb = (int)(((long long)c->ncol * c->nz) / (8 * 148));
    lpb = std::max(1, std::min(lpb, 30));
    { const char *e = getenv("VK_LHS_LPB"); if (e) lpb = atoi(e); }
    const int bpc = (c->nz + lpb - 1) / lpb;
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("VK_LHS_DBG"); dbg = e ? atoi(e) : 0; }
    if (nthr == 512) lhs_ml_kernel<512><<<c->ncol * bpc, 512, SL.total_bytes, c->stream>>>(a, SL, bpc, lpb, dbg);
    else if (nthr == 384) lhs_ml_kernel<384><<<c->ncol * bpc, 384, SL.total_bytes, c->stream>>>(a, SL, bpc, lpb, dbg);
    else lhs_ml_kernel<256><<<c->ncol * bpc, 256, SL.total_bytes, c->stream>>>(a, SL, bpc, lpb, dbg);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

}  // namespace vk

#ifdef VK_PROD_TRACE
extern "C" int vk_debug_prod_trace(long long *out) { return (int)cudaMemcpyFromSymbol(out, vk::g_prod_trace, sizeof(long long) * 16); }
#endif
