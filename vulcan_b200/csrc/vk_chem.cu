// Chemistry + transport kernels: one thread block per (column, layer).
//
//   rhs_kernel : chem_funs.chemdf (make_chem_funs.py:113-430) + ODESolver.diffdf / diffdf_settling / diffdf_no_mol
//                (op.py:1438-1791), optionally the Ros2 stage-2 right-hand side  f(y+k1/r) - 2/(r h) k1 (op.py:2917-2928)
//   lhs_kernel : chem_funs.neg_symjac block (make_chem_funs.py:653-717) + lhs_jac_tot / _settling / _no_mol
//                (op.py:1973-2364):  D = 1/(r h) I - J_chem - J_transport, up/dn = the diagonal couplings
//
// This translation unit is compiled with -fmad=false: every expression below is evaluated with one IEEE rounding per
// operation in the reference's order, so chemdf and diffdf are bit-identical to the numpy reference (tests/test_gpu_parity.py).
// The reaction tables are staged per block in shared memory next to the layer's y and k vectors.
#include "vk_internal.cuh"
#include "vk_device_math.cuh"

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
};

__global__ void __launch_bounds__(256) rhs_kernel(RhsArgs A)
{
    extern __shared__ double sm[];
    const int ni = A.net.ni, nr = A.net.nr, nz = A.nz;
    const int col = blockIdx.x / nz, j = blockIdx.x % nz;
    const int tid = threadIdx.x, nt = blockDim.x;
    // shared layout
    double *ym = sm;                    // y[j-1]  (ni+2)
    double *y0 = ym + (ni + 2);         // y[j]    yx: [ni] = M, [ni+1] = 1.0
    double *yp = y0 + (ni + 2);         // y[j+1]
    double *kz = yp + (ni + 2);         // k[j][0..nr]
    double *rate = kz + (nr + 2);       // rate[0..nr]
    double *tmp = rate + (nr + 2);      // 3*ni scratch for gas-compacted sums
    double *ysum = tmp + 3 * ni;        // [3]

    const size_t base = ((size_t)col * nz + j) * ni;
    const double rr = 1. + 1. / sqrt(2.);
    for (int i = tid; i < ni; i += nt) {
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
    if (tid == 0) { y0[ni] = A.atm.M[col * A.atm.csz + j]; y0[ni + 1] = 1.0; }
    const double *kg = A.k + col * A.k_cs + (size_t)j * (nr + 1);
    for (int i = tid; i <= nr; i += nt) kz[i] = kg[i];
    __syncthreads();

    // ---- rates of progress: rate[i] = k[i] * f0 * f1 * f2 * f3 (written order)
    for (int i = tid + 1; i <= nr; i += nt) {
        uchar4 f = A.net.rate_fac[i];
        double v = kz[i];
        if (!A.net.has_pow) {
            v = v * y0[f.x]; v = v * y0[f.y]; v = v * y0[f.z]; v = v * y0[f.w];   // padding slots multiply by exactly 1.0
        } else {
            uchar4 p = A.net.rate_pow[i];
            unsigned char ff[4] = {f.x, f.y, f.z, f.w}, pp[4] = {p.x, p.y, p.z, p.w};
            for (int q = 0; q < 4; q++) {
                double b = y0[ff[q]];
                double t = (pp[q] == 1) ? b : ((pp[q] == 2) ? b * b : pow(b, (double)pp[q]));
                v = v * t;
            }
        }
        rate[i] = v;
    }
    // ---- layer sums for the diffusion stencil (three threads, numpy pairwise order)
    if (tid < 3) {
        int jj = j - 1 + tid;
        if (jj >= 0 && jj < nz) {
            const double *row = (tid == 0) ? ym : ((tid == 1) ? y0 : yp);
            ysum[tid] = row_sum(row, ni, A.atm.n_gas, A.atm.gas_indx, tmp + tid * ni);
        }
    }
    __syncthreads();

    const AtmLayer L = atm_at(A.atm, col);
    const double *dzi = L.dzi, *Kzz = L.Kzz, *vz = L.vz;
    const int md = A.atm.use_moldiff, st = A.atm.use_settling && A.atm.use_moldiff;
    for (int i = tid; i < ni; i += nt) {
        // ---- chemistry: left-to-right sum of coef * (rate[j] - rate[j+1]) in network order
        double chem = 0.0;
        {
            int q0 = A.net.rhs_ptr[i], q1 = A.net.rhs_ptr[i + 1];
            for (int q = q0; q < q1; q++) {
                int t = A.net.rhs_term[q];
                int jf = t >> 8;
                double coef = (double)((signed char)(t & 0xff));
                double v = rate[jf] - rate[jf + 1];
                double term = coef * v;
                chem = (q == q0) ? term : chem + term;
            }
        }
        // ---- transport
        double diff;
        const double ys0 = ysum[1], ysm = ysum[0], ysp = ysum[2];
        if (j == 0) {
            double Aa = -1. / (dzi[0]) * (Kzz[0] / dzi[0]) * (ysp + ys0) / 2. / ys0;
            double Bb = 1. / (dzi[0]) * (Kzz[0] / dzi[0]) * (ysp + ys0) / 2. / ysp;
            Aa += -(posv(vz[0])) / dzi[0];
            Bb += -(negv(vz[0])) / dzi[0];
            if (md) {
                double D0 = L.Dzz[i];
                double br = phi_br(L, 0, i, L.g[0]);
                double Ai = -1. / (dzi[0]) * (D0 / dzi[0]) * (ysp + ys0) / 2. / ys0 + 1. / (dzi[0]) * D0 / 2. * br;
                double Bi = 1. / (dzi[0]) * (D0 / dzi[0]) * (ysp + ys0) / 2. / ysp + 1. / (dzi[0]) * D0 / 2. * br;
                if (st) {
                    Ai = Ai - (posv(L.vs[i])) / dzi[0];
                    Bi = Bi - (negv(L.vs[i])) / dzi[0];
                }
                diff = (Aa + Ai) * y0[i] + (Bb + Bi) * yp[i];
            } else {
                diff = Aa * y0[i] + Bb * yp[i];
            }
            if (A.atm.use_botflux) diff += (L.bot_flux[i] - y0[i] * L.bot_vdep[i]) / dzi[0];
        } else if (j == nz - 1) {
            const int m = nz - 2;
            double Aa = -1. / (dzi[m]) * (Kzz[m] / dzi[m]) * (ys0 + ysm) / 2. / ys0;
            double Cc = 1. / (dzi[m]) * (Kzz[m] / dzi[m]) * (ys0 + ysm) / 2. / ysm;
            Aa += (negv(vz[m])) / dzi[m];
            Cc += (posv(vz[m])) / dzi[m];
            if (md) {
                double Dm = L.Dzz[(size_t)m * ni + i];
                double br = phi_br(L, m, i, L.g[nz - 1]);
                double Ai = -1. / (dzi[m]) * (Dm / dzi[m]) * (ys0 + ysm) / 2. / ys0 - 1. / (dzi[m]) * Dm / 2. * br;
                double Ci = 1. / (dzi[m]) * (Dm / dzi[m]) * (ys0 + ysm) / 2. / ysm - 1. / (dzi[m]) * Dm / 2. * br;
                if (st) {
                    Ai = Ai + (negv(L.vs[(size_t)m * ni + i])) / dzi[m];
                    Ci = Ci + (posv(L.vs[(size_t)m * ni + i])) / dzi[m];
                }
                diff = (Aa + Ai) * y0[i] + (Cc + Ci) * ym[i];
            } else {
                diff = Aa * y0[i] + Cc * ym[i];
            }
            if (A.atm.use_topflux) diff += L.top_flux[i] / dzi[m];
        } else {
            double dz_ave = 0.5 * (dzi[j - 1] + dzi[j]);
            double Aa, Bb, Cc;
            if (md) {
                Aa = -1. / dz_ave * (Kzz[j] / dzi[j] * (ysp + ys0) / 2. + Kzz[j - 1] / dzi[j - 1] * (ys0 + ysm) / 2.) / ys0;
                Bb = 1. / dz_ave * Kzz[j] / dzi[j] * (ysp + ys0) / 2. / ysp;
                Cc = 1. / dz_ave * Kzz[j - 1] / dzi[j - 1] * (ys0 + ysm) / 2. / ysm;
            } else {   // diffdf_no_mol writes 2./(dzi[j-1]+dzi[j]) (op.py:1474-1476)
                Aa = -2. / (dzi[j - 1] + dzi[j]) * (Kzz[j] / dzi[j] * (ysp + ys0) / 2. + Kzz[j - 1] / dzi[j - 1] * (ys0 + ysm) / 2.) / ys0;
                Bb = 2. / (dzi[j - 1] + dzi[j]) * Kzz[j] / dzi[j] * (ysp + ys0) / 2. / ysp;
                Cc = 2. / (dzi[j - 1] + dzi[j]) * Kzz[j - 1] / dzi[j - 1] * (ys0 + ysm) / 2. / ysm;
            }
            Aa += -(posv(vz[j]) - negv(vz[j - 1])) / dz_ave;
            Bb += -(negv(vz[j])) / dz_ave;
            Cc += (posv(vz[j - 1])) / dz_ave;
            double t1 = Aa * y0[i] + Bb * yp[i] + Cc * ym[i];
            if (md) {
                double Dj = L.Dzz[(size_t)j * ni + i], Dm = L.Dzz[(size_t)(j - 1) * ni + i];
                double Ai = -1. / dz_ave * (Dj / dzi[j] * (ysp + ys0) / 2. + Dm / dzi[j - 1] * (ys0 + ysm) / 2.) / ys0;
                double Bi = 1. / dz_ave * Dj / dzi[j] * (ysp + ys0) / 2. / ysp;
                double Ci = 1. / dz_ave * Dm / dzi[j - 1] * (ys0 + ysm) / 2. / ysm;
                if (st) {
                    double vj = L.vs[(size_t)j * ni + i], vm = L.vs[(size_t)(j - 1) * ni + i];
                    Ai = Ai - (posv(vj) - negv(vm)) / dz_ave;
                    Bi = Bi - (negv(vj)) / dz_ave;
                    Ci = Ci + (posv(vm)) / dz_ave;
                }
                Ai += 1. / (2. * dz_ave) * (Dj * phi_br(L, j, i, L.g[j]) - Dm * phi_br(L, j - 1, i, L.g[j]));
                Bi += 1. / (2. * dz_ave) * Dj * phi_br(L, j, i, L.g[j + 1]);
                Ci += -1. / (2. * dz_ave) * Dm * phi_br(L, j - 1, i, L.g[j - 1]);
                double t2 = Ai * y0[i] + Bi * yp[i] + Ci * ym[i];
                diff = t1 + t2;
            } else {
                diff = t1;
            }
        }
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
};

__global__ void __launch_bounds__(256) lhs_kernel(LhsArgs A)
{
    extern __shared__ double sm[];
    const int ni = A.net.ni, nr = A.net.nr, nz = A.nz, ld = A.ld;
    const int col = blockIdx.x / nz, j = blockIdx.x % nz;
    const int tid = threadIdx.x, nt = blockDim.x;
    double *ym = sm;
    double *y0 = ym + (ni + 2);
    double *yp = y0 + (ni + 2);
    double *kz = yp + (ni + 2);
    double *tmp = kz + (nr + 2);
    double *ysum = tmp + 3 * ni;
    double *blk = ysum + 4;            // ld*ld block

    const size_t base = ((size_t)col * nz + j) * ni;
    for (int i = tid; i < ni; i += nt) {
        y0[i] = A.y[base + i];
        ym[i] = (j > 0) ? A.y[base - ni + i] : 0.0;
        yp[i] = (j < nz - 1) ? A.y[base + ni + i] : 0.0;
    }
    if (tid == 0) { y0[ni] = A.atm.M[col * A.atm.csz + j]; y0[ni + 1] = 1.0; }
    const double *kg = A.k + col * A.k_cs + (size_t)j * (nr + 1);
    for (int i = tid; i <= nr; i += nt) kz[i] = kg[i];
    for (int q = tid; q < ld * ld; q += nt) blk[q] = 0.0;
    __syncthreads();
    if (tid < 3) {
        int jj = j - 1 + tid;
        if (jj >= 0 && jj < nz) {
            const double *row = (tid == 0) ? ym : ((tid == 1) ? y0 : yp);
            ysum[tid] = row_sum(row, ni, A.atm.n_gas_lhs, A.atm.gas_indx_lhs, tmp + tid * ni);
        }
    }
    // ---- chemical Jacobian entries: blk[s][t] = -(d f_s / d y_t)
    for (int e = tid; e < A.net.n_ent; e += nt) {
        int q0 = A.net.jac_ptr[e], q1 = A.net.jac_ptr[e + 1];
        double acc = 0.0;
        for (int q = q0; q < q1; q++) {
            uint2 t = A.net.jac_term[q];
            double coef = (double)((signed char)((t.x >> 16) & 0xff));
            double term = coef * kz[t.x & 0xffff];
            term = term * y0[t.y & 0xff];
            term = term * y0[(t.y >> 8) & 0xff];
            term = term * y0[(t.y >> 16) & 0xff];
            acc += term;
        }
        ushort2 rc = A.net.jac_rc[e];
        blk[rc.x * ld + rc.y] = -acc;
    }
    __syncthreads();
    // ---- diagonal: c0 + negJ_ss - transport;  couplings up/dn   (op.py:1998-2040)
    const AtmLayer L = atm_at(A.atm, col);
    const double *dzi = L.dzi, *Kzz = L.Kzz, *vz = L.vz;
    const int md = A.atm.use_moldiff, st = A.atm.use_settling && A.atm.use_moldiff;
    const double rr = 1. + 1. / sqrt(2.);
    const double c0 = 1. / (rr * A.dt[col]);
    const double ys0 = ysum[1], ysm = ysum[0], ysp = ysum[2];
    const size_t vbase = ((size_t)col * nz + j) * ld;
    for (int i = tid; i < ld; i += nt) {
        if (i >= ni) {   // padding: decoupled identity rows keep the padded block invertible
            blk[i * ld + i] = 1.0;
            A.up[vbase + i] = 0.0;
            A.dn[vbase + i] = 0.0;
            continue;
        }
        double d = c0 + blk[i * ld + i];
        double u = 0.0, l = 0.0;
        if (j == 0) {
            d -= -1. / (dzi[0]) * (Kzz[0] / dzi[0]) * (ysp + ys0) / (2. * ys0) - (posv(vz[0])) / dzi[0];
            u -= 1. / (dzi[0]) * (Kzz[0] / dzi[0]) * (ysp + ys0) / (2. * ysp) - (negv(vz[0])) / dzi[0];
            if (md) {
                double D0 = L.Dzz[i];
                double br = phi_br(L, 0, i, L.g[0]);
                double ta = -1. / (dzi[0]) * (D0 / dzi[0]) * (ysp + ys0) / (2. * ys0) + 1. / (dzi[0]) * D0 / 2. * br;
                double tb = 1. / (dzi[0]) * (D0 / dzi[0]) * (ysp + ys0) / (2. * ysp) + 1. / (dzi[0]) * D0 / 2. * br;
                if (st) {
                    ta = ta - (posv(L.vs[i])) / dzi[0];
                    tb = tb - (negv(L.vs[i])) / dzi[0];
                }
                d -= ta;
                if (A.atm.use_botflux) d -= -1. * L.bot_vdep[i] / dzi[0];
                u -= tb;
            } else {
                if (A.atm.use_botflux) d -= -1. * L.bot_vdep[i] / dzi[0];
            }
        } else if (j == nz - 1) {
            const int m = nz - 2;
            d -= -1. / (dzi[m]) * (Kzz[m] / dzi[m]) * (ysm + ys0) / (2. * ys0) + (negv(vz[m])) / dzi[m];
            l -= 1. / (dzi[m]) * (Kzz[m] / dzi[m]) * (ysm + ys0) / (2. * ysm) + (posv(vz[m])) / dzi[m];
            if (md) {
                double Dm = L.Dzz[(size_t)m * ni + i];
                double br = phi_br(L, m, i, L.g[nz - 1]);
                double ta = -1. / (dzi[m]) * (Dm / dzi[m]) * (ys0 + ysm) / (2. * ys0) - 1. / (dzi[m]) * Dm / 2. * br;
                double tc = 1. / (dzi[m]) * (Dm / dzi[m]) * (ys0 + ysm) / (2. * ysm) - 1. / (dzi[m]) * Dm / 2. * br;
                if (st) {
                    ta = ta + (negv(L.vs[(size_t)m * ni + i])) / dzi[m];
                    tc = tc + (posv(L.vs[(size_t)m * ni + i])) / dzi[m];
                }
                d -= ta;
                l -= tc;
            }
        } else {
            double dz_ave = 0.5 * (dzi[j - 1] + dzi[j]);
            d -= -1. / dz_ave * (Kzz[j] / dzi[j] * (ysp + ys0) / 2. + Kzz[j - 1] / dzi[j - 1] * (ysm + ys0) / 2.) / ys0 -
                 (posv(vz[j]) - negv(vz[j - 1])) / dz_ave;
            u -= 1. / dz_ave * (Kzz[j] / dzi[j] * (ysp + ys0) / (2. * ysp)) - (negv(vz[j])) / dz_ave;
            l -= 1. / dz_ave * (Kzz[j - 1] / dzi[j - 1] * (ysm + ys0) / (2. * ysm)) + (posv(vz[j - 1])) / dz_ave;
            if (md) {
                double Dj = L.Dzz[(size_t)j * ni + i], Dm = L.Dzz[(size_t)(j - 1) * ni + i];
                double ta = -1. / dz_ave * (Dj / dzi[j] * (ysp + ys0) / 2. + Dm / dzi[j - 1] * (ysm + ys0) / 2.) / ys0 +
                            1. / (2. * dz_ave) * (Dj * phi_br(L, j, i, L.g[j]) - Dm * phi_br(L, j - 1, i, L.g[j]));
                double tb = 1. / dz_ave * (Dj / dzi[j] * (ysp + ys0) / (2. * ysp)) + 1. / (2. * dz_ave) * Dj * phi_br(L, j, i, L.g[j + 1]);
                double tc = 1. / dz_ave * (Dm / dzi[j - 1] * (ysm + ys0) / (2. * ysm)) - 1. / (2. * dz_ave) * Dm * phi_br(L, j - 1, i, L.g[j - 1]);
                if (st) {
                    double vj = L.vs[(size_t)j * ni + i], vm = L.vs[(size_t)(j - 1) * ni + i];
                    ta = ta - (posv(vj) - negv(vm)) / dz_ave;
                    tb = tb - (negv(vj)) / dz_ave;
                    tc = tc + (posv(vm)) / dz_ave;
                }
                d -= ta;
                u -= tb;
                l -= tc;
            }
        }
        if (A.fix_mask && A.fix_mask[base + i]) {   // op.py:2903-2906: row -> 1/(r h) e_i
            for (int t = 0; t < ni; t++) blk[i * ld + t] = 0.0;
            d = c0; u = 0.0; l = 0.0;
        }
        blk[i * ld + i] = d;
        A.up[vbase + i] = u;
        A.dn[vbase + i] = l;
    }
    __syncthreads();
    double *Dg = A.D + ((size_t)col * nz + j) * ld * ld;
    for (int q = tid; q < ld * ld; q += nt) Dg[q] = blk[q];
}

// ------------------------------------------------------------------------------------------------------------------
int launch_rhs(vk_column *c, const double *y_dev, double *out_sum, double *out_chem, double *out_diff,
               const double *k1_for_rhs2, const double *dt_dev)
{
    RhsArgs a;
    a.net = c->net->d; a.atm = c->atm; a.nz = c->nz; a.y = y_dev; a.k = c->k; a.k_cs = c->k_cs;
    a.k1 = k1_for_rhs2; a.dt = dt_dev; a.yk2_out = k1_for_rhs2 ? c->yk2 : nullptr;
    a.out_sum = out_sum; a.out_chem = out_chem; a.out_diff = out_diff;
    a.fix_mask = c->opts.fix_mask;
    size_t smem = sizeof(double) * (3 * (c->ni + 2) + 2 * (c->nr + 2) + 3 * c->ni + 4);
    rhs_kernel<<<c->ncol * c->nz, 256, smem, c->stream>>>(a);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

int launch_lhs(vk_column *c, const double *y_dev, const double *dt_dev, int ld, double *D_out, double *up_out, double *dn_out)
{
    LhsArgs a;
    a.net = c->net->d; a.atm = c->atm; a.nz = c->nz; a.y = y_dev; a.k = c->k; a.k_cs = c->k_cs; a.dt = dt_dev;
    a.D = D_out; a.up = up_out; a.dn = dn_out; a.ld = ld; a.fix_mask = c->opts.fix_mask;
    size_t smem = sizeof(double) * (3 * (c->ni + 2) + (c->nr + 2) + 3 * c->ni + 4 + (size_t)ld * ld);
    static size_t configured = 0;
    if (smem > configured) {
        VK_CUDA(cudaFuncSetAttribute(lhs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    lhs_kernel<<<c->ncol * c->nz, 256, smem, c->stream>>>(a);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

}  // namespace vk
