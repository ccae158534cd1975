// Photolysis update: ODESolver.compute_tau / compute_flux / compute_J (op.py:2580-2786).
//
// Wavelength-parallel: one thread per (column, wavelength bin) marches along z.
//   pass 1 (top -> bottom): optical depth suffix sum (op.py:2588-2599), Beer direct beam, two-stream coefficients
//                           (op.py:2612-2677) and the downward diffuse sweep (op.py:2691-2692) which uses the upward flux
//                           LEFT BY THE PREVIOUS CALL (lagged fixed point, state lives on the device);
//   pass 2 (bottom -> top): upward sweep (op.py:2693-2694), coefficients recomputed instead of stored;
//   pass 3               : actinic flux and its relative change (op.py:2709-2737).
// J rates: one warp per (column, branch, layer) contracts aflux[z,:] with the branch cross section (trapezoid on the two
// uniform grids, op.py:2766-2786) and writes J*f_diurnal straight into the device copy of k.
// Compiled with -fmad=false (same expression order as the reference; exp() is the only non-identical primitive).
#include "vk_internal.cuh"

struct PhotoState {
    int nbin, i12, n_abs, n_photo, n_scat, n_br;
    double dbin1, dbin2, sl_angle, edd, flux_atol, f_diurnal;
    double *bins, *sflux_top, *cross_abs, *cross_abs_T, *cross_photo, *cross_scat, *cross_J, *cross_J_T;
    int *abs_idx, *photo_idx, *scat_idx, *br_rate_index;
    unsigned char *abs_is_T, *br_is_T;
    // per column state
    double *tau, *sflux, *dflux_u, *dflux_d;   // [ncol][nz+1][nbin]
    double *aflux;                             // [ncol][nz][nbin]
    double *dz, *J;                            // [ncol][nz], [ncol][n_br][nz]
    unsigned long long *change_bits;           // [ncol]
    double *pk;                                // [ncol][nz][n_abs + 2 n_scat + n_photo] packed y dz / ymix of the species the sweeps read
    int photo_same;                            // the photo species and their cross sections are the absorbers' (the usual case): one tile
    std::vector<void *> allocs;
};

namespace vk {

struct FluxArgs {
    int nz, ni, nbin, ncol;
    int n_abs, n_photo, n_scat;
    const int *abs_idx, *photo_idx, *scat_idx;
    const double *cross_abs, *cross_abs_T, *cross_photo, *cross_scat;
    const unsigned char *abs_is_T;
    const double *y, *ymix, *dz, *sflux_top, *bins;
    double sl_angle, edd, flux_atol;
    double *tau, *sflux, *dflux_u, *dflux_d, *aflux;
    unsigned long long *change_bits;
    const int *pred;           // [ncol] or NULL: only the flagged columns are updated (device-resident cadence, vk_steady.cu)
    // what the sweeps read of y / ymix, packed per (column, layer) by photo_pack_kernel: [n_abs] y*dz of the absorbers, [n_scat] y*dz of the
    // scatterers, [n_photo] ymix of the photo species, [n_scat] ymix of the scatterers - 464 contiguous bytes that every thread of a block
    // reads at the same address (one cached broadcast) instead of three dependent loads (index, y, cross section) per species and layer
    double *pk;
    int pk_ld;
    int photo_same;
};

#define FLUX_TB 128
#define FLUX_NZMAX 192
#define FLUX_PKT 16
struct Coef { double chi, xi, phi, i_u, i_d; };

__device__ __forceinline__ Coef two_stream_coef(const FluxArgs &a, const double *ym_photo, const double *ym_scat, const double *xs_photo,
                                                const double *xs_scat, double tau_j, double tau_jp, double dir_j, double dir_jp, double mu_ang)
{
    // single-scattering albedo (op.py:2621-2636); xs_*: the thread's cross sections in shared memory (stride FLUX_TB)
    double tot_abs = 0.0, tot_scat = 0.0;
    for (int s = 0; s < a.n_photo; s++) tot_abs += ym_photo[s] * xs_photo[s * FLUX_TB];
    for (int s = 0; s < a.n_scat; s++) tot_scat += ym_scat[s] * xs_scat[s * FLUX_TB];
    double w0 = tot_scat / (tot_abs + tot_scat);
    if (w0 != w0) w0 = 0.0;                                   // np.nan_to_num
    else if (isinf(w0)) w0 = (w0 > 0) ? 1.7976931348623157e308 : -1.7976931348623157e308;
    w0 = fmin(w0, 1. - 1.E-8);
    const double edd = a.edd;
    double dtau = tau_j - tau_jp;
    double sq = sqrt(1. - w0);
    double tran = exp(-1. / edd * sq * dtau);
    double zp = 0.5 * (1. + sq), zm = 0.5 * (1. - sq);
    double ll = -1. * w0 / (1. / (mu_ang * mu_ang) - 1. / (edd * edd) * (1. - w0));
    double g_p = 0.5 * (ll * (1. / edd + 1. / mu_ang));
    double g_m = 0.5 * (ll * (1. / edd - 1. / mu_ang));
    double t2 = tran * tran;
    Coef c;
    c.chi = zm * zm * t2 - zp * zp;
    c.xi = zp * zm * (1. - t2);
    c.phi = (zm * zm - zp * zp) * tran;
    c.i_u = c.phi * g_p * dir_j - (c.xi * g_m + c.chi * g_p) * dir_jp;
    c.i_d = c.phi * g_m * dir_jp - (c.chi * g_m + c.xi * g_p) * dir_j;
    return c;
}

// y * dz and ymix of the species the sweeps read, one thread per (column, layer, slot)
__global__ void __launch_bounds__(256) photo_pack_kernel(FluxArgs a)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int PL = a.pk_ld;
    if (e >= (size_t)a.ncol * a.nz * PL) return;
    const int q = (int)(e % PL);
    const size_t lay = e / PL;
    const int col = (int)(lay / a.nz);
    if (a.pred && !a.pred[col]) return;
    const double *yj = a.y + lay * a.ni, *ymj = a.ymix + lay * a.ni;
    double v;
    if (q < a.n_abs) v = yj[a.abs_idx[q]] * a.dz[lay];
    else if (q < a.n_abs + a.n_scat) v = yj[a.scat_idx[q - a.n_abs]] * a.dz[lay];
    else if (q < a.n_abs + a.n_scat + a.n_photo) v = ymj[a.photo_idx[q - a.n_abs - a.n_scat]];
    else v = ymj[a.scat_idx[q - a.n_abs - a.n_scat - a.n_photo]];
    a.pk[e] = v;
}

// one thread per (column, wavelength bin): blockIdx.y = column, the block's FLUX_TB bins keep their cross sections in shared memory
// (xs[s][tid]: each thread only ever reads its own column of the tile)
__global__ void __launch_bounds__(FLUX_TB) flux_kernel(FluxArgs a)
{
    extern __shared__ double xs[];
    const int nz = a.nz, nbin = a.nbin;
    const int col = blockIdx.y, tid = threadIdx.x;
    if (a.pred && !a.pred[col]) return;
    const int b = blockIdx.x * FLUX_TB + tid;
    const bool live = b < nbin;
    const int bb = live ? b : nbin - 1;
    double *xs_abs = xs + tid, *xs_scat = xs_abs + (size_t)a.n_abs * FLUX_TB;
    double *xs_photo = a.photo_same ? xs_abs : xs_scat + (size_t)a.n_scat * FLUX_TB;
    for (int s = 0; s < a.n_abs; s++) xs_abs[s * FLUX_TB] = a.cross_abs[(size_t)s * nbin + bb];
    for (int s = 0; s < a.n_scat; s++) xs_scat[s * FLUX_TB] = a.cross_scat[(size_t)s * nbin + bb];
    if (!a.photo_same) for (int s = 0; s < a.n_photo; s++) xs_photo[s * FLUX_TB] = a.cross_photo[(size_t)s * nbin + bb];
    double change = 0.0;
    bool has = false;
    // two-stream coefficients of the downward sweep, kept for the upward sweep (same inputs, same values: tau / sflux of the layer and the one
    // above) in thread-local memory - 4 doubles per layer, coalesced across the warp, sized by the RESIDENT threads (2.4 GB on a B200) rather
    // than by columns x bins; saves the second evaluation (an exp, a sqrt, 4 divisions and the albedo sums per bin and layer)
    double cst[FLUX_NZMAX][4];
    const bool keep = nz <= FLUX_NZMAX;
    const int PL = a.pk_ld;
    const double *pkc = a.pk + (size_t)col * nz * PL;
    // pass 1 reads the packed y rows of FLUX_PKT layers at a time from shared memory (the same 464 bytes per layer for every thread of
    // the block: a broadcast, no global-load latency inside the species loops).  Every thread of the block walks the loop (barriers);
    // threads beyond the last bin compute on the last bin's data and store nothing
    double *pkt = xs + (size_t)FLUX_TB * (a.n_abs + a.n_scat + (a.photo_same ? 0 : a.n_photo));
    double *tau = a.tau + (size_t)col * (nz + 1) * nbin + bb;
    double *sfl = a.sflux + (size_t)col * (nz + 1) * nbin + bb;
    double *du = a.dflux_u + (size_t)col * (nz + 1) * nbin + bb;
    double *dd = a.dflux_d + (size_t)col * (nz + 1) * nbin + bb;
    double *af = a.aflux + (size_t)col * nz * nbin + bb;
    const double cosz = cos(a.sl_angle), mu_ang = -1. * cos(a.sl_angle);
    const double top = a.sflux_top[bb];
    {
        // ---- pass 1: top -> bottom
        double tau_above = 0.0;
        if (live) tau[(size_t)nz * nbin] = 0.0;
        double s_above = top * exp(-1. * tau_above / cosz);
        if (live) sfl[(size_t)nz * nbin] = s_above;
        double dd_above = dd[(size_t)nz * nbin];          // stays as left by the caller (zero): dflux_d[nz] is never written
        bool any_T = false;                               // temperature-dependent cross sections (Earth: per-layer tables) - rare
        if (a.abs_is_T) for (int s = 0; s < a.n_abs; s++) any_T = any_T || a.abs_is_T[s];
        for (int j0 = nz - 1; j0 >= 0; j0 -= FLUX_PKT) {
            __syncthreads();                              // the previous tile is consumed
            for (int q = tid; q < FLUX_PKT * PL; q += FLUX_TB) {
                const int l = q / PL, e = q - l * PL;
                if (j0 - l >= 0) pkt[q] = pkc[(size_t)(j0 - l) * PL + e];
            }
            __syncthreads();
            for (int l = 0; l < FLUX_PKT && j0 - l >= 0; l++) {
                const int j = j0 - l;
                const double *pj = pkt + l * PL;          // [n_abs] y dz | [n_scat] y dz | [n_photo] ymix | [n_scat] ymix
                const double du_j = du[(size_t)j * nbin]; // (issued before the sums: its latency hides behind them)
                double tj = 0.0;
                if (!any_T) {
                    for (int s = 0; s < a.n_abs; s++) tj += pj[s] * xs_abs[s * FLUX_TB];
                } else {
                    for (int s = 0; s < a.n_abs; s++) {
                        const double f = pj[s];
                        const double cs = a.abs_is_T[s] ? a.cross_abs_T[((size_t)s * nz + j) * nbin + bb] : xs_abs[s * FLUX_TB];
                        tj += f * cs;
                    }
                }
                for (int s = 0; s < a.n_scat; s++) tj += pj[a.n_abs + s] * xs_scat[s * FLUX_TB];
                tj += tau_above;
                if (live) tau[(size_t)j * nbin] = tj;
                double sj = top * exp(-1. * tj / cosz);
                if (live) sfl[(size_t)j * nbin] = sj;
                Coef c = two_stream_coef(a, pj + a.n_abs + a.n_scat, pj + a.n_abs + a.n_scat + a.n_photo, xs_photo, xs_scat, tj, tau_above,
                                         sj * cosz, s_above * cosz, mu_ang);
                double ddj = 1. / c.chi * (c.phi * dd_above - c.xi * du_j + c.i_d / mu_ang);   // op.py:2692
                if (live) dd[(size_t)j * nbin] = ddj;
                dd_above = ddj; tau_above = tj; s_above = sj;
                if (keep) { cst[j][0] = c.chi; cst[j][1] = c.xi; cst[j][2] = c.phi; cst[j][3] = c.i_u; }
            }
        }
    }
    if (live) {
        // ---- pass 2: bottom -> top (dflux_u[0] keeps its value: zero upward flux at the bottom)
        double du_below = du[0];
        for (int j = 1; j <= nz; j++) {
            const int m = j - 1;
            Coef c;
            if (keep) {
                c.chi = cst[m][0]; c.xi = cst[m][1]; c.phi = cst[m][2]; c.i_u = cst[m][3];
            } else {
                const double *pm = pkc + (size_t)m * PL;
                double t_m = tau[(size_t)m * nbin], t_j = tau[(size_t)j * nbin];
                double s_m = sfl[(size_t)m * nbin], s_j = sfl[(size_t)j * nbin];
                c = two_stream_coef(a, pm + a.n_abs + a.n_scat, pm + a.n_abs + a.n_scat + a.n_photo, xs_photo, xs_scat, t_m, t_j,
                                    s_m * cosz, s_j * cosz, mu_ang);
            }
            double duj = 1. / c.chi * (c.phi * du_below - c.xi * dd[(size_t)j * nbin] + c.i_u / mu_ang);   // op.py:2694
            du[(size_t)j * nbin] = duj;
            du_below = duj;
        }
        // ---- pass 3: actinic flux (op.py:2709-2737)
        const double hcl = VK_HC / a.bins[b];
        for (int j = 0; j < nz; j++) {
            double ave_dir = 0.5 * (sfl[(size_t)j * nbin] + sfl[(size_t)(j + 1) * nbin]);
            double tot = ave_dir + 0.5 * (du[(size_t)j * nbin] + du[(size_t)(j + 1) * nbin] + dd[(size_t)(j + 1) * nbin] + dd[(size_t)j * nbin]) / a.edd;
            double prev = af[(size_t)j * nbin];
            double cur = tot / hcl;
            af[(size_t)j * nbin] = cur;
            if (cur > a.flux_atol) {
                double ch = fabs(cur - prev) / cur;
                if (ch == ch && (!has || ch > change)) { change = ch; has = true; }   // np.nanmax
            }
        }
    }
    unsigned long long bits = has ? (unsigned long long)__double_as_longlong(change) : 0ull;
    for (int off = 16; off > 0; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, bits, off);
        bits = (o > bits) ? o : bits;
    }
    if ((tid & 31) == 0 && bits) atomicMax(a.change_bits + col, bits);
}

struct JArgs {
    int nz, nbin, i12, n_br, ncol, nr;
    double dbin1, dbin2, f_diurnal;
    const double *aflux, *cross_J, *cross_J_T;
    const unsigned char *br_is_T;
    const int *br_rate_index;
    double *J;       // [ncol][n_br][nz]
    double *k;       // device k
    size_t k_cs;
    const int *pred;
};

// one warp per (column, branch, layer): the latency version for a few columns (7200 warps for one HD189 column)
__global__ void __launch_bounds__(256) jrate_kernel(JArgs a)
{
    const int lane = threadIdx.x & 31;
    const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t nw = (size_t)a.ncol * a.n_br * a.nz;
    if (w >= nw) return;
    const int j = (int)(w % a.nz), br = (int)((w / a.nz) % a.n_br), col = (int)(w / ((size_t)a.nz * a.n_br));
    if (a.pred && !a.pred[col]) return;
    const double *f = a.aflux + ((size_t)col * a.nz + j) * a.nbin;
    const double *c = (a.br_is_T && a.br_is_T[br]) ? a.cross_J_T + ((size_t)br * a.nz + j) * a.nbin : a.cross_J + (size_t)br * a.nbin;
    double s1 = 0.0, s2 = 0.0;
    for (int b = lane; b < a.i12; b += 32) s1 += f[b] * c[b] * a.dbin1;
    for (int b = a.i12 + lane; b < a.nbin; b += 32) s2 += f[b] * c[b] * a.dbin2;
    for (int off = 16; off > 0; off >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
    }
    if (lane == 0) {
        double v = s1;
        v -= 0.5 * (f[0] * c[0] + f[a.i12 - 1] * c[a.i12 - 1]) * a.dbin1;
        v += s2;
        v -= 0.5 * (f[a.i12] * c[a.i12] + f[a.nbin - 1] * c[a.nbin - 1]) * a.dbin2;
        a.J[((size_t)col * a.n_br + br) * a.nz + j] = v;
        const int rid = a.br_rate_index[br];
        if (rid > 0) a.k[(size_t)col * a.k_cs + (size_t)j * (a.nr + 1) + rid] = v * a.f_diurnal;   // op.py:2785-2786
    }
}

// (4 layers x 8 branches per warp: 32 accumulators, two blocks of 8 warps per SM)
// One warp per (column, tile of JR_LT layers, tile of JR_BT branches): lanes stride over the wavelength bins; the cross section of a branch
// at a bin is loaded once for the JR_LT layers, the actinic flux of a layer at a bin once for the JR_BT branches (the first version - one warp
// per (column, branch, layer) - moved 18 GB through L1 / L2 per 64 columns).  Every lane adds its bins in increasing order exactly as before,
// so J has the same bits.
#define JR_LT 4
#define JR_BT 8
__global__ void __launch_bounds__(256, 2) jrate_tile_kernel(JArgs a)
{
    const int lane = threadIdx.x & 31;
    const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int nlt = (a.nz + JR_LT - 1) / JR_LT, nbt = (a.n_br + JR_BT - 1) / JR_BT;
    const size_t nw = (size_t)a.ncol * nlt * nbt;
    if (w >= nw) return;
    const int bt = (int)(w % nbt), lt = (int)((w / nbt) % nlt), col = (int)(w / ((size_t)nbt * nlt));
    if (a.pred && !a.pred[col]) return;
    const int j0 = lt * JR_LT, br0 = bt * JR_BT;
    const int nl = min(JR_LT, a.nz - j0), nb = min(JR_BT, a.n_br - br0);
    const double *f[JR_LT];
#pragma unroll
    for (int l = 0; l < JR_LT; l++) f[l] = a.aflux + ((size_t)col * a.nz + j0 + (l < nl ? l : 0)) * a.nbin;
    bool tileT = false;
    if (a.br_is_T) for (int q = 0; q < nb; q++) tileT = tileT || a.br_is_T[br0 + q];
    auto cptr = [&](int q, int l) -> const double * {      // cross section row of branch br0 + q (at layer j0 + l when temperature dependent)
        const int br = br0 + (q < nb ? q : 0);
        return (a.br_is_T && a.br_is_T[br]) ? a.cross_J_T + ((size_t)br * a.nz + j0 + (l < nl ? l : 0)) * a.nbin : a.cross_J + (size_t)br * a.nbin;
    };
    __shared__ double vs[8][JR_LT][JR_BT];      // region-0 part of every J of the warp's tile (keeps 64 doubles out of the registers)
    double (*v)[JR_BT] = vs[threadIdx.x >> 5];
    for (int region = 0; region < 2; region++) {
        const int lo = region ? a.i12 : 0, hi = region ? a.nbin : a.i12;
        const double db = region ? a.dbin2 : a.dbin1;
        double acc[JR_LT][JR_BT];
#pragma unroll
        for (int l = 0; l < JR_LT; l++)
#pragma unroll
            for (int q = 0; q < JR_BT; q++) acc[l][q] = 0.0;
        if (!tileT) {
            const double *c[JR_BT];
#pragma unroll
            for (int q = 0; q < JR_BT; q++) c[q] = cptr(q, 0);
            for (int b = lo + lane; b < hi; b += 32) {
                double fv[JR_LT];
#pragma unroll
                for (int l = 0; l < JR_LT; l++) fv[l] = f[l][b];
#pragma unroll
                for (int q = 0; q < JR_BT; q++) {
                    const double cv = c[q][b];
#pragma unroll
                    for (int l = 0; l < JR_LT; l++) acc[l][q] += fv[l] * cv * db;
                }
            }
        } else {
            for (int b = lo + lane; b < hi; b += 32)
#pragma unroll
                for (int q = 0; q < JR_BT; q++)
#pragma unroll
                    for (int l = 0; l < JR_LT; l++) acc[l][q] += f[l][b] * cptr(q, l)[b] * db;
        }
#pragma unroll
        for (int l = 0; l < JR_LT; l++)
#pragma unroll
            for (int q = 0; q < JR_BT; q++) {
                double s = acc[l][q];
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                // trapezoid on the uniform grid of the region: the end points count half (op.py:2766-2786)
                const double *cc = cptr(q, l);
                const double e = 0.5 * (f[l][lo] * cc[lo] + f[l][hi - 1] * cc[hi - 1]) * db;
                if (lane == 0) {
                    if (region == 0) v[l][q] = s - e;
                    else { double t = v[l][q]; t += s; t -= e; v[l][q] = t; }
                }
            }
    }
    if (lane == 0) {
        for (int l = 0; l < nl; l++)
            for (int q = 0; q < nb; q++) {
                const int br = br0 + q, j = j0 + l;
                a.J[((size_t)col * a.n_br + br) * a.nz + j] = v[l][q];
                const int rid = a.br_rate_index[br];
                if (rid > 0) a.k[(size_t)col * a.k_cs + (size_t)j * (a.nr + 1) + rid] = v[l][q] * a.f_diurnal;   // op.py:2785-2786
            }
    }
}

void photo_destroy(vk_column *c)
{
    if (!c->photo) return;
    for (void *p : c->photo->allocs) cudaFree(p);
    delete c->photo;
    c->photo = nullptr;
}

template <typename T>
static int pcopy(PhotoState *p, const T *host, size_t n, T **out)
{
    void *d = nullptr;
    VK_CUDA(cudaMalloc(&d, sizeof(T) * (n ? n : 1)));
    p->allocs.push_back(d);
    if (host && n) VK_CUDA(cudaMemcpy(d, host, sizeof(T) * n, cudaMemcpyHostToDevice));
    else VK_CUDA(cudaMemset(d, 0, sizeof(T) * (n ? n : 1)));
    *out = reinterpret_cast<T *>(d);
    return VK_OK;
}

}  // namespace vk

using namespace vk;

extern "C" {

int vk_photo_setup(vk_column *c, const vk_photo_view *v)
{
    if (!c || !v || v->nbin < 4 || v->i12 < 2 || v->i12 > v->nbin - 2) { set_error("bad photo view"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    photo_destroy(c);
    PhotoState *p = new PhotoState();
    c->photo = p;
    p->nbin = v->nbin; p->i12 = v->i12; p->dbin1 = v->dbin1; p->dbin2 = v->dbin2; p->sl_angle = v->sl_angle; p->edd = v->edd;
    p->flux_atol = v->flux_atol; p->f_diurnal = v->f_diurnal;
    p->n_abs = v->n_abs; p->n_photo = v->n_photo; p->n_scat = v->n_scat; p->n_br = v->n_br;
    const size_t nb = v->nbin, nz = c->nz, ncol = c->ncol;
    int rc = VK_OK;
#define PC(field, n) if (rc == VK_OK) rc = pcopy(p, v->field, (size_t)(n), &p->field)
    PC(bins, nb); PC(sflux_top, nb);
    PC(abs_idx, v->n_abs); PC(cross_abs, v->n_abs * nb);
    PC(photo_idx, v->n_photo); PC(cross_photo, v->n_photo * nb);
    p->photo_same = (v->n_abs == v->n_photo && memcmp(v->abs_idx, v->photo_idx, sizeof(int) * v->n_abs) == 0 &&
                     memcmp(v->cross_abs, v->cross_photo, sizeof(double) * (size_t)v->n_abs * nb) == 0) ? 1 : 0;
    PC(scat_idx, v->n_scat); PC(cross_scat, v->n_scat * nb);
    PC(cross_J, v->n_br * nb); PC(br_rate_index, v->n_br);
#undef PC
    p->abs_is_T = nullptr; p->cross_abs_T = nullptr; p->br_is_T = nullptr; p->cross_J_T = nullptr;
    if (rc == VK_OK && v->abs_is_T && v->cross_abs_T) {
        rc = pcopy(p, v->abs_is_T, (size_t)v->n_abs, &p->abs_is_T);
        if (rc == VK_OK) rc = pcopy(p, v->cross_abs_T, (size_t)v->n_abs * nz * nb, &p->cross_abs_T);
    }
    if (rc == VK_OK && v->br_is_T && v->cross_J_T) {
        rc = pcopy(p, v->br_is_T, (size_t)v->n_br, &p->br_is_T);
        if (rc == VK_OK) rc = pcopy(p, v->cross_J_T, (size_t)v->n_br * nz * nb, &p->cross_J_T);
    }
    const double *nul = nullptr;
    if (rc == VK_OK) rc = pcopy(p, nul, ncol * (nz + 1) * nb, &p->tau);
    if (rc == VK_OK) rc = pcopy(p, nul, ncol * (nz + 1) * nb, &p->sflux);
    if (rc == VK_OK) rc = pcopy(p, nul, ncol * (nz + 1) * nb, &p->dflux_u);
    if (rc == VK_OK) rc = pcopy(p, nul, ncol * (nz + 1) * nb, &p->dflux_d);
    if (rc == VK_OK) rc = pcopy(p, nul, ncol * nz * nb, &p->aflux);
    if (rc == VK_OK) rc = pcopy(p, nul, ncol * nz, &p->dz);
    if (rc == VK_OK) rc = pcopy(p, nul, ncol * (size_t)v->n_br * nz, &p->J);
    const unsigned long long *nulu = nullptr;
    if (rc == VK_OK) rc = pcopy(p, nulu, ncol, &p->change_bits);
    if (rc != VK_OK) photo_destroy(c);
    else VK_CUDA(cudaDeviceSynchronize());      // memsets of the legacy stream before anything runs on the handle's non-blocking stream
    return rc;
}

int vk_photo_reset(vk_column *c)
{
    if (!c || !c->photo) { set_error("photo not set up"); return VK_ERR_INVALID; }
    PhotoState *p = c->photo;
    VK_CUDA(cudaSetDevice(c->net->device));
    const size_t n1 = (size_t)c->ncol * (c->nz + 1) * p->nbin;
    VK_CUDA(cudaMemsetAsync(p->dflux_u, 0, sizeof(double) * n1, c->stream));
    VK_CUDA(cudaMemsetAsync(p->dflux_d, 0, sizeof(double) * n1, c->stream));
    VK_CUDA(cudaMemsetAsync(p->aflux, 0, sizeof(double) * (size_t)c->ncol * c->nz * p->nbin, c->stream));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    return VK_OK;
}

}  // extern "C"

namespace vk {
__global__ void photo_begin_kernel(int ncol, const int *pred, unsigned long long *bits)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col < ncol && (!pred || pred[col])) bits[col] = 0ull;
}
__global__ void photo_end_kernel(int ncol, const int *pred, const unsigned long long *bits, double *change)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col < ncol && (!pred || pred[col])) change[col] = __longlong_as_double((long long)bits[col]);
}

// device-resident update: y, ymix, dz on the device; pred (optional) selects the columns; aflux_change_out (optional, device) receives
// var.aflux_change of the updated columns
int photo_update_device(vk_column *c, const double *y_dev, const double *ymix_dev, const double *dz_dev, const int *pred,
                        double *aflux_change_out)
{
    PhotoState *p = c->photo;
    FluxArgs a;
    a.nz = c->nz; a.ni = c->ni; a.nbin = p->nbin; a.ncol = c->ncol;
    a.n_abs = p->n_abs; a.n_photo = p->n_photo; a.n_scat = p->n_scat;
    a.abs_idx = p->abs_idx; a.photo_idx = p->photo_idx; a.scat_idx = p->scat_idx;
    a.cross_abs = p->cross_abs; a.cross_abs_T = p->cross_abs_T; a.cross_photo = p->cross_photo; a.cross_scat = p->cross_scat;
    a.abs_is_T = p->abs_is_T;
    a.y = y_dev; a.ymix = ymix_dev; a.dz = dz_dev; a.sflux_top = p->sflux_top; a.bins = p->bins;
    a.sl_angle = p->sl_angle; a.edd = p->edd; a.flux_atol = p->flux_atol;
    a.tau = p->tau; a.sflux = p->sflux; a.dflux_u = p->dflux_u; a.dflux_d = p->dflux_d; a.aflux = p->aflux;
    a.change_bits = p->change_bits; a.pred = pred;
    const int nb = (c->ncol + 127) / 128;
    photo_begin_kernel<<<nb, 128, 0, c->stream>>>(c->ncol, pred, p->change_bits);
    a.pk_ld = p->n_abs + 2 * p->n_scat + p->n_photo;
    if (!p->pk) {
        VK_CUDA(cudaMalloc((void **)&p->pk, sizeof(double) * (size_t)c->ncol * c->nz * a.pk_ld));
        p->allocs.push_back(p->pk);
    }
    a.pk = p->pk;
    const size_t npk = (size_t)c->ncol * c->nz * a.pk_ld;
    photo_pack_kernel<<<(unsigned)((npk + 255) / 256), 256, 0, c->stream>>>(a);
    a.photo_same = p->photo_same;
    const size_t smem = sizeof(double) * (FLUX_TB * (size_t)(p->n_abs + p->n_scat + (p->photo_same ? 0 : p->n_photo)) + (size_t)FLUX_PKT * a.pk_ld);
    if (smem > 200 * 1024) { set_error("too many absorbing species for the flux kernel's shared-memory tile"); return VK_ERR_UNSUPPORTED; }
    { int rc = ensure_smem((const void *)flux_kernel, c->net->device, smem); if (rc) return rc; }
    flux_kernel<<<dim3((unsigned)((p->nbin + FLUX_TB - 1) / FLUX_TB), (unsigned)c->ncol), FLUX_TB, smem, c->stream>>>(a);
    VK_CUDA(cudaGetLastError());
    JArgs ja;
    ja.nz = c->nz; ja.nbin = p->nbin; ja.i12 = p->i12; ja.n_br = p->n_br; ja.ncol = c->ncol; ja.nr = c->nr;
    ja.dbin1 = p->dbin1; ja.dbin2 = p->dbin2; ja.f_diurnal = p->f_diurnal;
    ja.aflux = p->aflux; ja.cross_J = p->cross_J; ja.cross_J_T = p->cross_J_T; ja.br_is_T = p->br_is_T;
    ja.br_rate_index = p->br_rate_index; ja.J = p->J; ja.k = c->k; ja.k_cs = c->k_cs; ja.pred = pred;
    if (c->ncol >= 16) {       // batches: layer x branch tiles per warp (fewer, longer warps); a few columns: one warp per (branch, layer)
        const size_t nwarp = (size_t)c->ncol * ((c->nz + JR_LT - 1) / JR_LT) * ((p->n_br + JR_BT - 1) / JR_BT);
        jrate_tile_kernel<<<(unsigned)((nwarp * 32 + 255) / 256), 256, 0, c->stream>>>(ja);
    } else {
        const size_t nwarp = (size_t)c->ncol * p->n_br * c->nz;
        jrate_kernel<<<(unsigned)((nwarp * 32 + 255) / 256), 256, 0, c->stream>>>(ja);
    }
    if (aflux_change_out) photo_end_kernel<<<nb, 128, 0, c->stream>>>(c->ncol, pred, p->change_bits, aflux_change_out);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}
}  // namespace vk

extern "C" {

int vk_photo_update(vk_column *c, const double *y, const double *ymix, const double *dz, double *J, double *aflux_change)
{
    if (!c || !c->photo || !y || !ymix || !dz) { set_error("photo not set up / null buffer"); return VK_ERR_INVALID; }
    if (!c->k_set) { set_error("vk_set_k must be called first"); return VK_ERR_INVALID; }
    if (c->k_cs == 0 && c->ncol > 1) { set_error("photolysis writes J into k: k must be per column"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    PhotoState *p = c->photo;
    const size_t nv = (size_t)c->ncol * c->nz * c->ni;
    VK_CUDA(cudaMemcpyAsync(c->y, y, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(c->ymix, ymix, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(p->dz, dz, sizeof(double) * c->ncol * c->nz, cudaMemcpyHostToDevice, c->stream));
    int rc = photo_update_device(c, c->y, c->ymix, p->dz, nullptr, nullptr);
    if (rc) return rc;
    if (J) VK_CUDA(cudaMemcpyAsync(J, p->J, sizeof(double) * (size_t)c->ncol * p->n_br * c->nz, cudaMemcpyDeviceToHost, c->stream));
    std::vector<unsigned long long> bits(c->ncol);
    VK_CUDA(cudaMemcpyAsync(bits.data(), p->change_bits, sizeof(unsigned long long) * c->ncol, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    if (aflux_change)
        for (int i = 0; i < c->ncol; i++) memcpy(&aflux_change[i], &bits[i], sizeof(double));
    return VK_OK;
}

int vk_photo_read(vk_column *c, double *tau, double *sflux, double *dflux_u, double *dflux_d, double *aflux)
{
    if (!c || !c->photo) { set_error("photo not set up"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    PhotoState *p = c->photo;
    const size_t n1 = sizeof(double) * (size_t)c->ncol * (c->nz + 1) * p->nbin, n0 = sizeof(double) * (size_t)c->ncol * c->nz * p->nbin;
    VK_CUDA(cudaStreamSynchronize(c->stream));
    if (tau) VK_CUDA(cudaMemcpy(tau, p->tau, n1, cudaMemcpyDeviceToHost));
    if (sflux) VK_CUDA(cudaMemcpy(sflux, p->sflux, n1, cudaMemcpyDeviceToHost));
    if (dflux_u) VK_CUDA(cudaMemcpy(dflux_u, p->dflux_u, n1, cudaMemcpyDeviceToHost));
    if (dflux_d) VK_CUDA(cudaMemcpy(dflux_d, p->dflux_d, n1, cudaMemcpyDeviceToHost));
    if (aflux) VK_CUDA(cudaMemcpy(aflux, p->aflux, n0, cudaMemcpyDeviceToHost));
    return VK_OK;
}

}  // extern "C"
