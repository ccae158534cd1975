// Rate coefficients on the device: replaces the reference's set-up-time ReadRate.read_rate / lim_lowT_rates / rev_rate /
// remove_rate (op.py:63-342) and the generated chem_funs.Gibbs (make_chem_funs.py:568-580, thermo/gibbs_text.txt) for ensembles whose
// columns differ in T-P: k[col][layer][1..nr] is produced where it is consumed instead of being uploaded (1 MB per column).
//   forward   a T^n exp(-E/T);  three-body with high-pressure limit  k0 / (1 + k0 M / k_inf);  the special OH + CH3 + M form
//   caps      the three low-temperature limits of Moses+2005 (use_lowT_limit_rates)
//   reverse   k_f / K_eq,  K_eq = exp(-(sum nu g/RT)) (kb T / 1e6)^(n_reac - n_prod),  g/RT = h/RT - s/R from NASA-9 (switch at 1000 K)
// One block per (column, layer): g/RT of every species once into shared memory, then one thread per reaction pair.  Compiled with
// -fmad=false and the reference's operation order; libm vs CUDA pow/exp/log leave ulp-level differences (tests: 1e-13 relative).
#include "vk_internal.cuh"

struct RateTables {
    int npair, ni;
    int *kind;          // [npair] 0 zero (photo / ion / condensation), 1 Arrhenius, 2 falloff, 3 special OH + CH3 + M
    double *arr;        // [npair][6] a, n, E, a_inf, n_inf, E_inf
    int *cap_kind;      // [npair] 0 none, 1 k0/(1 + k0 M/(2.06e-10 T^-0.4)) below cap[0] with k0 = cap[1], 2 constant cap[1] below cap[0]
    double *cap;        // [npair][2]
    unsigned char *rev; // [npair] reverse rate from the equilibrium constant (id + 1 < stop_rev_indx, not removed)
    unsigned char *fwd_removed;
    int *gptr, *gsp, *gnu, *dnu;   // Gibbs terms in written order (reactants -n, products +n, M skipped); dnu = n_reac - n_prod
    double *nasa9;      // [ni][20]
    std::vector<void *> allocs;
};

namespace vk {

void rates_destroy(vk_network *n)
{
    if (!n || !n->rates) return;
    for (void *p : n->rates->allocs) cudaFree(p);
    delete n->rates;
    n->rates = nullptr;
}

struct RateArgs {
    RateTables t;
    int nz, nr;
    const double *Tco, *M;   // [ncol or 1][nz]
    size_t tm_cs;            // column stride of Tco / M (0 = shared)
    double *k;               // [ncol or 1][nz][nr+1]
    size_t k_cs;
};

__device__ __forceinline__ double h_RT(double T, const double *a)      // gibbs_text.txt:14-15
{
    return -a[0] / (T * T) + a[1] * log(T) / T + a[2] + a[3] * T / 2. + a[4] * (T * T) / 3. + a[5] * pow(T, 3.) / 4. + a[6] * pow(T, 4.) / 5. + a[8] / T;
}
__device__ __forceinline__ double s_R(double T, const double *a)       // gibbs_text.txt:18-19
{
    return -a[0] / (T * T) / 2. - a[1] / T + a[2] * log(T) + a[3] * T + a[4] * (T * T) / 2. + a[5] * pow(T, 3.) / 3. + a[6] * pow(T, 4.) / 4. + a[9];
}

__global__ void __launch_bounds__(128) rate_kernel(RateArgs A)
{
    extern __shared__ double g[];       // g/RT per species
    const int nz = A.nz, nr = A.nr;
    const int col = blockIdx.x / nz, j = blockIdx.x % nz;
    const double T = A.Tco[col * A.tm_cs + j], M = A.M[col * A.tm_cs + j];
    for (int s = threadIdx.x; s < A.t.ni; s += blockDim.x) {
        const double *a = A.t.nasa9 + (size_t)s * 20 + ((T < 1000) ? 0 : 10);
        g[s] = h_RT(T, a) - s_R(T, a);
    }
    __syncthreads();
    double *kout = A.k + col * A.k_cs + (size_t)j * (nr + 1);
    if (threadIdx.x == 0) kout[0] = 0.0;
    for (int p = threadIdx.x; p < A.t.npair; p += blockDim.x) {
        const double *c = A.t.arr + (size_t)p * 6;
        const int kind = A.t.kind[p];
        double kf = 0.0;
        if (kind == 1 || kind == 2) {
            kf = c[0] * pow(T, c[1]) * exp(-c[2] / T);                                  // op.py:166
            if (kind == 2) {
                const double k_inf = c[3] * pow(T, c[4]) * exp(-c[5] / T);              // op.py:181
                kf = kf / (1 + kf * M / k_inf);                                        // op.py:183
            }
        } else if (kind == 3) {                                                        // op.py:197-207
            kf = 1.932E3 * pow(T, -9.88) * exp(-7544. / T) + 5.109E-11 * pow(T, -6.25) * exp(-1433. / T);
            const double k_inf = 1.031E-10 * pow(T, -0.018) * exp(16.74 / T);
            const double Fc = 0.1855 * exp(-T / 155.8) + 0.8145 * exp(-T / 1675.) + exp(-4531. / T);
            const double nn = 0.75 - 1.27 * log(Fc);
            const double lr = log(kf * M / k_inf) / nn;
            const double ff = exp(log(Fc) / (1. + lr * lr));
            kf = kf / (1 + kf * M / k_inf) * ff;
        }
        const int ck = A.t.cap_kind[p];
        if (ck) {                                                                      // op.py:320-342
            const double Tcap = A.t.cap[2 * p], val = A.t.cap[2 * p + 1];
            if (T <= Tcap) {
                if (ck == 1) { const double kinf = 2.06E-10 * pow(T, -0.4); kf = val / (1. + val * M / kinf); }
                else kf = val;
            }
        }
        double kr = 0.0;
        if (A.t.rev[p]) {                                                              // op.py:289-304
            double acc = 0.0;
            const int q0 = A.t.gptr[p], q1 = A.t.gptr[p + 1];
            for (int q = q0; q < q1; q++) {
                const int nu = A.t.gnu[q];
                const double term = (double)(nu < 0 ? -nu : nu) * g[A.t.gsp[q]];
                if (q == q0) acc = (nu < 0) ? -term : term;
                else acc = (nu < 0) ? acc - term : acc + term;
            }
            double K = exp(-(acc));
            const int dnu = A.t.dnu[p];
            if (dnu != 0) K = K * pow(VK_KB / 1.e6 * T, (double)dnu);
            kr = kf / K;
        }
        if (A.t.fwd_removed[p] & 1) kf = 0.0;                                          // op.py:311-317
        if (A.t.fwd_removed[p] & 2) kr = 0.0;
        kout[2 * p + 1] = kf;
        kout[2 * p + 2] = kr;
    }
}

template <typename T>
static int rcopy(RateTables *t, const T *host, size_t n, T **out)
{
    void *d = nullptr;
    VK_CUDA(cudaMalloc(&d, sizeof(T) * (n ? n : 1)));
    t->allocs.push_back(d);
    if (n) VK_CUDA(cudaMemcpy(d, host, sizeof(T) * n, cudaMemcpyHostToDevice));
    *out = reinterpret_cast<T *>(d);
    return VK_OK;
}

}  // namespace vk

using namespace vk;

extern "C" {

int vk_rates_set(vk_network *n, const vk_rate_desc *d)
{
    if (!n || !d || d->npair * 2 != n->d.nr || !d->kind || !d->arrhenius || !d->nasa9) { set_error("bad rate description"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(n->device));
    rates_destroy(n);
    RateTables *t = new RateTables();
    n->rates = t;
    t->npair = d->npair; t->ni = n->d.ni;
    const size_t np = d->npair, ng = d->gibbs_ptr[d->npair];
    int rc = rcopy(t, d->kind, np, &t->kind);
    if (rc == VK_OK) rc = rcopy(t, d->arrhenius, np * 6, &t->arr);
    if (rc == VK_OK) rc = rcopy(t, d->cap_kind, np, &t->cap_kind);
    if (rc == VK_OK) rc = rcopy(t, d->cap, np * 2, &t->cap);
    if (rc == VK_OK) rc = rcopy(t, d->reverse, np, &t->rev);
    if (rc == VK_OK) rc = rcopy(t, d->removed, np, &t->fwd_removed);
    if (rc == VK_OK) rc = rcopy(t, d->gibbs_ptr, np + 1, &t->gptr);
    if (rc == VK_OK) rc = rcopy(t, d->gibbs_sp, ng, &t->gsp);
    if (rc == VK_OK) rc = rcopy(t, d->gibbs_nu, ng, &t->gnu);
    if (rc == VK_OK) rc = rcopy(t, d->dnu, np, &t->dnu);
    if (rc == VK_OK) rc = rcopy(t, d->nasa9, (size_t)n->d.ni * 20, &t->nasa9);
    if (rc != VK_OK) rates_destroy(n);
    return rc;
}

int vk_compute_k(vk_column *c, const double *Tco, const double *M, int shared)
{
    if (!c || !Tco || !M) { set_error("null argument"); return VK_ERR_INVALID; }
    if (!c->net->rates) { set_error("vk_rates_set must be called first"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    const size_t per = (size_t)c->nz * (c->nr + 1);
    const size_t want_cs = shared ? 0 : per;
    const size_t ncolk = shared ? 1 : c->ncol;
    if (!c->k || c->k_cs != want_cs || !c->k_set) {
        VK_CUDA(cudaStreamSynchronize(c->stream));
        if (c->k) cudaFree(c->k);
        c->k = nullptr;
        VK_CUDA(cudaMalloc((void **)&c->k, sizeof(double) * per * ncolk));
        c->k_cs = want_cs;
    }
    double *dT = nullptr, *dM = nullptr;
    VK_CUDA(cudaMalloc((void **)&dT, sizeof(double) * ncolk * c->nz));
    VK_CUDA(cudaMalloc((void **)&dM, sizeof(double) * ncolk * c->nz));
    VK_CUDA(cudaMemcpyAsync(dT, Tco, sizeof(double) * ncolk * c->nz, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(dM, M, sizeof(double) * ncolk * c->nz, cudaMemcpyHostToDevice, c->stream));
    RateArgs a{*c->net->rates, c->nz, c->nr, dT, dM, shared ? (size_t)0 : (size_t)c->nz, c->k, c->k_cs};
    rate_kernel<<<(unsigned)(ncolk * c->nz), 128, sizeof(double) * c->ni, c->stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dT);
    cudaFree(dM);
    if (e != cudaSuccess) return cuda_fail(e, "vk_compute_k");
    c->k_set = true;
    c->k_static_shared = false;
    return VK_OK;
}

int vk_get_k(vk_column *c, double *k_host)
{
    if (!c || !k_host || !c->k_set) { set_error("no rate coefficients on the device"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    const size_t per = (size_t)c->nz * (c->nr + 1);
    VK_CUDA(cudaMemcpy(k_host, c->k, sizeof(double) * per * (c->k_cs ? c->ncol : 1), cudaMemcpyDeviceToHost));
    return VK_OK;
}

}  // extern "C"
