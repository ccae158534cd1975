// Condensation operators of the caller: Integration.conden (op.py:1109-1300), h2o_conden_evap_relax / nh3_conden_evap_relax
// (op.py:1340-1421) and the fix_species switch (op.py:859-893) as device kernels, so that the configurations with condensation (Jupiter, Earth)
// run in the device-resident loop as well (vk_steady.cu) instead of round-tripping var.y / var.k through the host after every accepted step.
// All three act on nz-vectors of one or two species per column; the arithmetic follows the reference expression by expression (this
// translation unit is compiled with -fmad=false), so the results are bit-identical to the reference's operators on the same inputs
// (tests/test_gpu_conden.py against tests/golden/<cfg>_conden.npz, recorded from the reference's own calls).
#include "vk_internal.cuh"
#include "vk_device_math.cuh"
#include "vk_ens_state.cuh"

struct CondenDev {       // plain data: passed to the kernels by value
    int n_re, n_relax, n_fix;
    int *re_idx, *gas_idx;              // [n_re] forward reaction id of `X -> X_l_s`, gas species
    double *m, *rho_p, *r_p;            // [n_re] molecular mass (g), particle density, particle radius
    double *sat;                        // [n_re][nz] saturation number density  sat_p / kb / Tco  (x humidity for H2O)
    unsigned char *zero_rate;           // [n_re] the relaxation operator replaces this growth reaction: k[re] = k[re+1] = 0 (op.py:1124-1126)
    int *relax_kind, *relax_gas, *relax_ice, *relax_top;   // [n_relax] 1 = H2O (op.py:1340-1376), 2 = NH3 (op.py:1378-1421); conden_top (NH3)
    double *relax_m, *relax_rho, *relax_r, *relax_sat;     // [n_relax], relax_sat [n_relax][nz]
    // fix_species switch
    int *fix_sp, *fix_all;              // [n_fix] species, 1 = frozen through the whole column (condensates; fix_species_from_coldtrap_lev = False)
    double *fix_sat_mix;                // [n_fix][nz] atm.sat_mix of the species (gas species with a cold trap) or zeros
    double start_conden_time, stop_conden_time, post_conden_rtol;
    int use_fix;
};
struct CondenState {
    CondenDev d;
    std::vector<void *> allocs;
};

namespace vk {

void conden_destroy(vk_column *c)
{
    if (!c->conden) return;
    for (void *p : c->conden->allocs) cudaFree(p);
    delete c->conden;
    c->conden = nullptr;
}

struct CondenArgs {
    int nz, ni, nr, ncol;
    CondenDev s;
    const double *y, *Dzz; size_t csn;   // y [ncol][nz][ni]; atm.Dzz [nz-1][ni] with column stride csn
    double *k; size_t k_cs;
    const int *pred;
    double *k_rows_out;                  // optional [ncol][n_re][2][nz]
};
// conden: growth / evaporation rate coefficients of the condensation reactions from the new number densities (op.py:1109-1300)
__global__ void conden_rate_kernel(CondenArgs a)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = a.ncol * a.s.n_re * a.nz;
    if (q >= total) return;
    const int j = q % a.nz, r = (q / a.nz) % a.s.n_re, col = q / (a.nz * a.s.n_re);
    if (a.pred && !a.pred[col]) return;
    double kf = 0.0, kb_ = 0.0;
    if (!a.s.zero_rate[r]) {
        const int gi = a.s.gas_idx[r];
        const double *Dzz = a.Dzz + col * a.csn;
        const double Dg = Dzz[(size_t)(j > 0 ? j - 1 : 0) * a.ni + gi];              // np.insert(Dzz[:, i], 0, Dzz[0, i])
        const double yg = a.y[((size_t)col * a.nz + j) * a.ni + gi];
        const double rp = a.s.r_p[r];
        const double rate = Dg * a.s.m[r] / a.s.rho_p[r] / (rp * rp) * (yg - a.s.sat[(size_t)r * a.nz + j]);
        kf = fmax(rate, 0.0);                                                         // positive: condensation
        kb_ = fabs(fmin(rate, 0.0));                                                  // negative: evaporation
        if (rate != rate) { kf = rate; kb_ = rate; }                                  // np.maximum / np.minimum propagate nan
    }
    const int re = a.s.re_idx[r];
    double *kc = a.k + col * a.k_cs + (size_t)j * (a.nr + 1);
    kc[re] = kf;
    kc[re + 1] = kb_;
    if (a.k_rows_out) {
        a.k_rows_out[(((size_t)col * a.s.n_re + r) * 2 + 0) * a.nz + j] = kf;
        a.k_rows_out[(((size_t)col * a.s.n_re + r) * 2 + 1) * a.nz + j] = kb_;
    }
}

struct RelaxArgs {
    int nz, ni, which;
    CondenDev s;
    double *y, *ymix;
    const double *Dzz; size_t csn;
    const double *n_0;                   // [ncol][nz]
    const double *dt;                    // [ncol] the step just taken
    int n_gas; const int *gas_indx;
    const int *pred;
};
// h2o_conden_evap_relax / nh3_conden_evap_relax: implicit-Euler relaxation of the vapour towards saturation, the difference goes to /
// comes from the condensate; then y = ymix * sum_gas(y).  One block per column.
__global__ void __launch_bounds__(256) relax_kernel(RelaxArgs a)
{
    extern __shared__ double gsum[];     // [nz] sum over the gas species of y BEFORE the operator
    const int col = blockIdx.x, tid = threadIdx.x, nz = a.nz, ni = a.ni, w = a.which;
    if (a.pred && !a.pred[col]) return;
    double *yc = a.y + (size_t)col * nz * ni, *ym = a.ymix + (size_t)col * nz * ni;
    const double *Dzz = a.Dzz + col * a.csn, *n0 = a.n_0 + (size_t)col * nz;
    const int ig = a.s.relax_gas[w], il = a.s.relax_ice[w], kind = a.s.relax_kind[w], top = a.s.relax_top[w];
    const double m = a.s.relax_m[w], rho_p = a.s.relax_rho[w], r_p = a.s.relax_r[w], dt = a.dt[col];
    const double *sat = a.s.relax_sat + (size_t)w * nz;
    for (int j = tid; j < nz; j += blockDim.x) {
        gsum[j] = row_sum(yc + (size_t)j * ni, ni, a.n_gas, a.gas_indx, nullptr);
        const double Dg = Dzz[(size_t)(j > 0 ? j - 1 : 0) * ni + ig];
        const double yg = yc[(size_t)j * ni + ig], yl = yc[(size_t)j * ni + il];
        const double tau = 1. / (Dg * m / rho_p / (r_p * r_p) * (yg - sat[j]));
        const double sat_mix = sat[j] / n0[j];
        double mg = ym[(size_t)j * ni + ig], ml = ym[(size_t)j * ni + il];
        const double y_conden = (mg + dt / tau * sat_mix) / (1. + dt / tau);
        const double raw_loss = (yg - sat[j]) * dt / tau;
        const double ice_loss = (yl != yl || raw_loss != raw_loss) ? (yl + raw_loss) : fmin(yl, raw_loss);   // np.minimum (nan propagates)
        if (tau > 0 && (kind != 2 || j <= top)) {            // condensation (NH3: not above the top of the condensation zone)
            ml += (mg - y_conden);
            mg = y_conden;
        } else if (tau < 0) {                                // evaporation
            mg += ice_loss / n0[j];
            ml -= ice_loss / n0[j];
        }
        if (kind == 2) ml = fmax(ml, 0.0);                   // op.py:1419
        ym[(size_t)j * ni + ig] = mg;
        ym[(size_t)j * ni + il] = ml;
    }
    __syncthreads();
    for (int q = tid; q < nz * ni; q += blockDim.x) yc[q] = ym[q] * gsum[q / ni];        // var.y = var.ymix * vstack(sum(y[:, gas_indx]))
}

template <typename T>
static int ccopy(CondenState *s, const T *host, size_t n, T **out)
{
    void *d = nullptr;
    VK_CUDA(cudaMalloc(&d, sizeof(T) * (n ? n : 1)));
    s->allocs.push_back(d);
    if (host && n) VK_CUDA(cudaMemcpy(d, host, sizeof(T) * n, cudaMemcpyHostToDevice));
    else VK_CUDA(cudaMemset(d, 0, sizeof(T) * (n ? n : 1)));
    *out = reinterpret_cast<T *>(d);
    return VK_OK;
}

// conden (part & 1) and the relaxation operators (part & 2) on device state; pred (optional): the columns they act on
int conden_device(vk_column *c, double *y_dev, double *ymix_dev, const double *dt_dev, const double *n0_dev, const int *pred, double *k_rows_out,
                  int part)
{
    CondenDev &s = c->conden->d;
    if ((part & 1) && s.n_re > 0) {
        CondenArgs a{c->nz, c->ni, c->nr, c->ncol, s, y_dev, c->atm.Dzz, c->atm.csn, c->k, c->k_cs, pred, k_rows_out};
        const int total = c->ncol * s.n_re * c->nz;
        conden_rate_kernel<<<(total + 127) / 128, 128, 0, c->stream>>>(a);
    }
    for (int w = 0; (part & 2) && w < s.n_relax; w++) {
        RelaxArgs r{c->nz, c->ni, w, s, y_dev, ymix_dev, c->atm.Dzz, c->atm.csn, n0_dev, dt_dev, c->atm.n_gas, c->atm.gas_indx, pred};
        relax_kernel<<<c->ncol, 256, sizeof(double) * c->nz, c->stream>>>(r);
    }
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

}  // namespace vk

using namespace vk;

extern "C" {

int vk_conden_setup(vk_column *c, const vk_conden_desc *d)
{
    if (!c || !d || d->n_re < 0 || d->n_relax < 0) { set_error("bad condensation description"); return VK_ERR_INVALID; }
    if (!c->atm_set || !c->k_set) { set_error("vk_set_atm / vk_set_k must be called first"); return VK_ERR_INVALID; }
    if (c->k_cs == 0 && c->ncol > 1) { set_error("conden writes rate coefficients into k: vk_set_k per column"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->device));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    conden_destroy(c);
    CondenState *st = new CondenState();
    c->conden = st;
    CondenDev *s = &st->d;
    const size_t nz = c->nz;
    s->n_re = d->n_re; s->n_relax = d->n_relax; s->n_fix = 0; s->use_fix = 0;
    s->start_conden_time = d->start_conden_time; s->stop_conden_time = d->stop_conden_time; s->post_conden_rtol = d->post_conden_rtol;
    for (int r = 0; r < d->n_re; r++)
        if (d->re_idx[r] < 1 || d->re_idx[r] + 1 > c->nr || d->gas_idx[r] < 0 || d->gas_idx[r] >= c->ni) { set_error("condensation reaction / species out of range"); conden_destroy(c); return VK_ERR_INVALID; }
    int rc = ccopy(st, d->re_idx, (size_t)d->n_re, &s->re_idx);
    if (rc == VK_OK) rc = ccopy(st, d->gas_idx, (size_t)d->n_re, &s->gas_idx);
    if (rc == VK_OK) rc = ccopy(st, d->m, (size_t)d->n_re, &s->m);
    if (rc == VK_OK) rc = ccopy(st, d->rho_p, (size_t)d->n_re, &s->rho_p);
    if (rc == VK_OK) rc = ccopy(st, d->r_p, (size_t)d->n_re, &s->r_p);
    if (rc == VK_OK) rc = ccopy(st, d->sat, (size_t)d->n_re * nz, &s->sat);
    if (rc == VK_OK) rc = ccopy(st, d->zero_rate, (size_t)d->n_re, &s->zero_rate);
    if (rc == VK_OK) rc = ccopy(st, d->relax_kind, (size_t)d->n_relax, &s->relax_kind);
    if (rc == VK_OK) rc = ccopy(st, d->relax_gas, (size_t)d->n_relax, &s->relax_gas);
    if (rc == VK_OK) rc = ccopy(st, d->relax_ice, (size_t)d->n_relax, &s->relax_ice);
    if (rc == VK_OK) rc = ccopy(st, d->relax_top, (size_t)d->n_relax, &s->relax_top);
    if (rc == VK_OK) rc = ccopy(st, d->relax_m, (size_t)d->n_relax, &s->relax_m);
    if (rc == VK_OK) rc = ccopy(st, d->relax_rho, (size_t)d->n_relax, &s->relax_rho);
    if (rc == VK_OK) rc = ccopy(st, d->relax_r, (size_t)d->n_relax, &s->relax_r);
    if (rc == VK_OK) rc = ccopy(st, d->relax_sat, (size_t)d->n_relax * nz, &s->relax_sat);
    if (rc != VK_OK) { conden_destroy(c); return rc; }
    VK_CUDA(cudaDeviceSynchronize());
    return VK_OK;
}

// component entry point (parity tests): conden + relaxation operators on host state.  y, ymix [ncol][nz][ni] in / out; dt [ncol];
// n_0 [ncol][nz]; k_rows out [ncol][n_re][2][nz] (the two rate coefficients of every condensation reaction) or NULL
int vk_conden_apply(vk_column *c, double *y, double *ymix, const double *dt, const double *n_0, double *k_rows)
{
    if (!c || !c->conden || !y || !ymix || !dt || !n_0) { set_error("vk_conden_setup first / null buffer"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->device));
    CondenDev &s = c->conden->d;
    const size_t nv = (size_t)c->ncol * c->nz * c->ni;
    double *dn0 = nullptr, *dk = nullptr;
    VK_CUDA(cudaMalloc((void **)&dn0, sizeof(double) * c->ncol * c->nz));
    if (k_rows && s.n_re > 0) VK_CUDA(cudaMalloc((void **)&dk, sizeof(double) * (size_t)c->ncol * s.n_re * 2 * c->nz));
    VK_CUDA(cudaMemcpyAsync(c->sol, y, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(c->ymix_out, ymix, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(c->dt, dt, sizeof(double) * c->ncol, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(dn0, n_0, sizeof(double) * c->ncol * c->nz, cudaMemcpyHostToDevice, c->stream));
    int rc = conden_device(c, c->sol, c->ymix_out, c->dt, dn0, nullptr, dk, 3);
    if (rc == VK_OK) {
        VK_CUDA(cudaMemcpyAsync(y, c->sol, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
        VK_CUDA(cudaMemcpyAsync(ymix, c->ymix_out, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
        if (dk) VK_CUDA(cudaMemcpyAsync(k_rows, dk, sizeof(double) * (size_t)c->ncol * s.n_re * 2 * c->nz, cudaMemcpyDeviceToHost, c->stream));
        VK_CUDA(cudaStreamSynchronize(c->stream));
    }
    cudaFree(dn0);
    if (dk) cudaFree(dk);
    return rc;
}

}  // extern "C"
