// Device-resident run to steady state for a batch of columns: the caller-side protocol of the reference's
// Integration.__call__ (op.py:808-935) that round 1 still ran on the host -
//   stop / conv            op.py:1018-1087   convergence test against the stored history at the st_factor look-back, end cases 1 / 2 / 3
//   photolysis cadence     op.py:818-829     update every update_photo_frq accepted steps, switch to final_update_photo_frq near convergence
//   update_mu_dz           op.py:944-984     mean molecular weight -> g, Hp, dz, zco, dzi, Hpi (hydrostatic grid), every update_frq steps
//   update_phi_esc         op.py:986-999     diffusion-limited escape flux at the top
//   save_step              op.py:1089-1105   t += dt, count += 1, history
// - per column, with no host round trip: columns of a sweep converge after different numbers of steps, every kernel of the step skips the
// columns that have stopped (vk_column::act).  One iteration = one ATTEMPTED step of every active column (vk_ens.cu); the pre-step part
// (stop test, photolysis update) runs only on "fresh" columns, i.e. after an accepted step, as in the reference where a rejected attempt
// is retried inside one_step (op.py:3091-3103).
//
// History: conv() compares the current state with y_time[indx], indx = the stored step closest to t * st_factor, at most conv_step steps
// back.  A single column keeps all conv_step states (exact).  A large batch cannot (500 states x 83 KB x 4096 columns = 170 GB): the ring
// then keeps every `stride`-th accepted state, `cap` of them, and the look-back index is rounded to a stored one - the criterion is
// evaluated against a state at most stride / 2 steps away from the reference's choice (stride = 1: identical).
#include "vk_internal.cuh"
#include "vk_ens_state.cuh"

namespace vk {

int vk_step_device_impl(vk_column *c);
int launch_clip(vk_column *c, double *y_dev, const double *ymix_in_dev, double *ymix_out_dev, int na, const double *compo_dev,
                const unsigned char *skip_dev, double pos_cut, double nega_cut, double *atom_sum_dev, double *small_dev,
                double *nega_dev, int *anyneg_dev);
int launch_ens_control(vk_column *c);
int launch_ens_apply(vk_column *c);
int launch_atm_pre_pred(vk_column *c, const int *pred);
int photo_update_device(vk_column *c, const double *y_dev, const double *ymix_dev, const double *dz_dev, const int *pred,
                        double *aflux_change_out);
int conden_device(vk_column *c, double *y_dev, double *ymix_dev, const double *dt_dev, const double *n0_dev, const int *pred, double *k_rows_out,
                  int part);

// the fix_species switch of the flagged columns (op.py:860-893): record the values the condensable species are frozen at, their cold-trap
// levels, switch the rtol read by step_size, turn the settling velocity off.  One block per column.
struct SwitchArgs {
    int nz, ni;
    SteadyDev s;
    AtmDev atm;
    const double *y, *n_0;             // state after the step (post-clip), [ncol][nz][ni]; n_0 [ncol][nz]
    unsigned char *fix_mask; double *fix_y;   // step options of the handle, [ncol][nz][ni]
};
__global__ void __launch_bounds__(128) steady_switch_kernel(SwitchArgs a)
{
    __shared__ double s_min;
    __shared__ int s_lev;
    const int col = blockIdx.x, tid = threadIdx.x, nz = a.nz, ni = a.ni;
    SteadyDev &s = a.s;
    if (!s.do_switch[col]) return;
    const double *yc = a.y + (size_t)col * nz * ni, *n0 = a.n_0 + (size_t)col * nz;
    for (int f = 0; f < s.n_fix; f++) {
        const int i = s.fix_sp[f];
        int top = nz;                                                  // fix_species_from_coldtrap_lev = False: the whole column
        if (s.fix_from_coldtrap) {
            if (s.fix_whole[f]) {
                top = nz - 1;                                          // condensates (op.py:878-879)
            } else {
                if (tid == 0) {                                        // gas species: the level of the minimum saturation mixing ratio inside the
                    const double *sm = s.fix_sat_mix + (size_t)f * nz; // region where it condenses (op.py:881-892)
                    double best = 0.0; int have = 0, lev = 0;
                    for (int j = 0; j < nz; j++)
                        if (yc[(size_t)j * ni + i] >= n0[j] * sm[j] && (!have || sm[j] < best)) { best = sm[j]; have = 1; lev = j; }
                    s_min = best; s_lev = have ? lev : 0;
                }
                __syncthreads();
                top = s_lev;
                __syncthreads();
            }
        }
        for (int j = tid; j < nz; j += blockDim.x) {
            const size_t q = ((size_t)col * nz + j) * ni + i;
            a.fix_mask[q] = (j < top) ? 1 : 0;
            a.fix_y[q] = (j < top) ? yc[(size_t)j * ni + i] : 0.0;
        }
    }
    double *vs = const_cast<double *>(a.atm.vs) + col * a.atm.csn;     // atm.vs *= 0
    for (int q = tid; q < (nz - 1) * ni; q += blockDim.x) vs[q] = 0.0 * vs[q];
    if (tid == 0) { s.fix_started[col] = 1; s.rtol_col[col] = s.post_conden_rtol; }
}

struct PreArgs {
    int nz, ni;
    SteadyDev s;
    const double *y, *ymix, *n_0, *t;
    const int *n_accept;
    const double *Kzz; size_t cs1;
};

// stop() + conv() + photolysis cadence for the fresh columns: one block per column
__global__ void __launch_bounds__(256) steady_pre_kernel(PreArgs a)
{
    __shared__ double red[8];
    __shared__ int s_indx;
    const int col = blockIdx.x, tid = threadIdx.x, nz = a.nz, ni = a.ni;
    SteadyDev &s = a.s;
    if (!s.act[col]) return;
    if (!s.fresh[col]) { if (tid == 0) s.do_photo[col] = 0; return; }
    const int count = a.n_accept[col];
    const double t = a.t[col];
    const bool test_conv = (t > s.trun_min) && (count > s.count_min);        // Python's `and` short-circuits: conv() is not even called otherwise
    double longdy = 0.0, slope = 1.0e300;
    if (test_conv) {
        if (tid == 0) {
            // indx = argmin |t_time - t * st_factor| over the accepted steps 0 .. count-1 (t_time ascending: binary search, first minimum)
            const double *tt = s.t_time + (size_t)col * s.cap_t;
            const double target = t * s.st_factor;
            int lo = 0, hi = count - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (tt[mid] < target) lo = mid + 1; else hi = mid; }
            int indx = lo;
            if (indx > 0 && fabs(tt[indx - 1] - target) <= fabs(tt[indx] - target)) indx -= 1;
            if (indx == count - 1) indx -= 1;                                  // op.py:1032
            indx = max(count - s.conv_step, indx);                             // op.py:1035
            // ring: stored indices are the multiples of `stride`; round to the nearest stored one that is still in the ring and older
            // than the current state
            const int st = s.hist_stride;
            int q = (indx + st / 2) / st;
            const int newest = (count - 2) / st, oldest = max(0, (count - 1) / st - s.hist_cap + 1);
            q = min(max(q, oldest), max(newest, 0));
            s_indx = q * st;
        }
        __syncthreads();
        const int indx = s_indx;
        const double *yh = s.hist + ((size_t)col * s.hist_cap + (indx / s.hist_stride) % s.hist_cap) * ((size_t)nz * ni);
        const double *yc = a.y + (size_t)col * nz * ni, *ym = a.ymix + (size_t)col * nz * ni;
        const double *n0 = a.n_0 + (size_t)col * nz;
        double best = 0.0;
        for (int q = tid; q < nz * ni; q += 256) {
            const int j = q / ni, i = q % ni;
            const double m = ym[q];
            double v = fabs((yc[q] - yh[q]) / n0[j]);                          // op.py:1040
            if (m < s.mtol_conv) v = 0;
            if (yc[q] < s.atol) v = 0;
            if (s.ignore_sp && s.ignore_sp[i]) v = 0;                          // conver_ignore, non-gas species (op.py:1045-1049)
            if (m > 0) { v = v / m; if (v > best || v != v) best = v; }
        }
        for (int off = 16; off > 0; off >>= 1) { const double o = __shfl_xor_sync(0xffffffffu, best, off); if (o > best || o != o) best = o; }
        if ((tid & 31) == 0) red[tid >> 5] = best;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; w++) if (red[w] > best || red[w] != red[w]) best = red[w];
            longdy = best;
            const double *Kzz = a.Kzz + col * a.cs1, *Hp = s.Hp + (size_t)col * nz;
            double smin = 1.0e300;
            for (int j = 0; j < nz - 1; j++) { const double h = 0.1 * Hp[j]; smin = fmin(smin, Kzz[j] / (h * h)); }
            slope = fmax(fmin(smin, 1.e-8), 1.e-10);                          // op.py:1028-1029
        }
    }
    if (tid == 0) {
        int end_case = 0;
        if (test_conv) {
            const double *tt = s.t_time + (size_t)col * s.cap_t;
            const double longdydt = longdy / (tt[count - 1] - tt[s_indx]);    // op.py:1055
            s.longdy[col] = longdy; s.longdydt[col] = longdydt;
            const bool ok = ((longdy < s.yconv_cri && longdydt < s.slope_cri) || (longdy < s.yconv_min && longdydt < slope)) &&
                            (s.aflux_change[col] < s.flux_cri);
            if (ok) end_case = 1;
        }
        if (!end_case && t > s.runtime) end_case = 2;
        if (!end_case && count > s.count_max) end_case = 3;
        int dop = 0;
        if (end_case) {
            s.end_case[col] = end_case;
            s.act[col] = 0;
        } else if (s.use_photo) {
            if (s.longdy[col] < s.yconv_min * 10. && s.longdydt[col] < 1.e-6) s.photo_frq[col] = s.final_frq;     // op.py:818-822
            dop = (count % s.photo_frq[col] == 0) ? 1 : 0;                                                         // op.py:824
        }
        s.do_photo[col] = dop;
    }
}

struct MuArgs {
    int nz, ni;
    SteadyDev s;
    AtmDev atm;
    const double *ymix, *ysol;     // after the clip of the accepted attempt
};
// update_mu_dz + update_phi_esc of the columns flagged by the controller: one block per column
__global__ void __launch_bounds__(128) steady_mu_dz_kernel(MuArgs a)
{
    const int col = blockIdx.x, tid = threadIdx.x, nz = a.nz, ni = a.ni;
    SteadyDev &s = a.s;
    if (!s.do_mu[col]) return;
    double *mu = s.mu + (size_t)col * nz, *Hp = s.Hp + (size_t)col * nz, *dz = s.dz + (size_t)col * nz, *zco = s.zco + (size_t)col * (nz + 1);
    const double *ym = a.ymix + (size_t)col * nz * ni;
    for (int j = tid; j < nz; j += blockDim.x) {          // mean_mass: mu += ms[i] * ymix[:, i] in species order (build_atm.py:515-520)
        double m = 0.0;
        for (int i = 0; i < ni; i++) m += s.ms[i] * ym[(size_t)j * ni + i];
        mu[j] = m;
    }
    __syncthreads();
    const AtmDev &A = a.atm;
    double *g = const_cast<double *>(A.g) + col * A.csz;
    double *dzi = const_cast<double *>(A.dzi) + col * A.cs1;
    double *Hpi = const_cast<double *>(A.Hpi) + col * A.cs1;
    const double *Tco = A.Tco + col * A.csz;
    if (tid == 0) {
        const int p0 = s.pref_indx;
        const double *pico = s.pico;
        for (int i = p0; i < nz; i++) {                    // op.py:955-964
            if (i == p0) g[i] = s.gs;
            else g[i] = s.gs * ((s.Rp / (s.Rp + zco[i])) * (s.Rp / (s.Rp + zco[i])));
            Hp[i] = VK_KB * Tco[i] / (mu[i] / VK_NAVO * g[i]);
            dz[i] = Hp[i] * log(pico[i] / pico[i + 1]);
            zco[i + 1] = zco[i] + dz[i];
        }
        for (int i = p0 - 1; i >= 0; i--) {                // op.py:966-971
            g[i] = s.gs * ((s.Rp / (s.Rp + zco[i + 1])) * (s.Rp / (s.Rp + zco[i + 1])));
            Hp[i] = VK_KB * Tco[i] / (mu[i] / VK_NAVO * g[i]);
            dz[i] = Hp[i] * log(pico[i] / pico[i + 1]);
            zco[i] = zco[i + 1] - dz[i];
        }
    }
    __syncthreads();
    for (int j = tid; j < nz - 1; j += blockDim.x) {       // op.py:974-982
        dzi[j] = 0.5 * (dz[j + 1] + dz[j]);
        if (A.use_moldiff) Hpi[j] = 0.5 * (Hp[j] + Hp[j + 1]);
    }
    // update_phi_esc (op.py:986-999): top_flux of the escaping species from the state after the clip
    if (tid < s.n_diff_esc) {
        const int i = s.diff_esc_idx[tid];
        double *tf = const_cast<double *>(A.top_flux) + col * A.csi;
        const double *Dzz = A.Dzz + col * A.csn, *ms = A.ms + col * A.csi;
        const double ytop = a.ysol[((size_t)col * nz + nz - 1) * ni + i];
        double f = -Dzz[(size_t)(nz - 2) * ni + i] * ytop * (1. / Hp[nz - 1] - ms[i] * g[nz - 1] / (VK_NAVO * VK_KB * Tco[nz - 1]));
        tf[i] = fmax(f, s.max_flux * (-1));
    }
}

template <typename T>
static int scopy(EnsState *e, const T *host, size_t n, T **out)
{
    void *d = nullptr;
    VK_CUDA(cudaMalloc(&d, sizeof(T) * (n ? n : 1)));
    e->allocs.push_back(d);
    if (host) VK_CUDA(cudaMemcpy(d, host, sizeof(T) * n, cudaMemcpyHostToDevice));
    else VK_CUDA(cudaMemset(d, 0, sizeof(T) * (n ? n : 1)));
    *out = reinterpret_cast<T *>(d);
    return VK_OK;
}

__global__ void fill_kernel(int n, double *a, double va, int *b, int vb)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) { if (a) a[q] = va; if (b) b[q] = vb; }
}
__global__ void count_active_kernel(int n, const int *act, int *out)
{
    int c = 0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) c += act[q] ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, c);
}

}  // namespace vk

using namespace vk;

extern "C" {

int vk_ens_setup_steady(vk_column *c, const vk_steady_opts *o)
{
    if (!c || !c->ens || !o || !o->pico || !o->ms || !o->zco || !o->Hp || !o->dz) { set_error("vk_ens_setup first / null steady options"); return VK_ERR_INVALID; }
    if (c->atm.cs1 == 0 && c->ncol > 1 && o->update_frq > 0) { set_error("update_mu_dz changes the grid per column: vk_set_atm with shared = 0"); return VK_ERR_INVALID; }
    if (o->use_photo && !c->photo) { set_error("use_photo needs vk_photo_setup"); return VK_ERR_INVALID; }
    if (o->use_photo && c->k_cs == 0 && c->ncol > 1) { set_error("photolysis writes J into k: vk_set_k per column"); return VK_ERR_INVALID; }
    if (o->hist_cap < 2 || o->hist_stride < 1 || o->conv_step < 1 || o->count_max < 1) { set_error("bad history / count options"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    EnsState *e = c->ens;
    SteadyDev &s = e->steady;
    const size_t ncol = c->ncol, nz = c->nz, ni = c->ni;
    s.st_factor = o->st_factor; s.mtol_conv = o->mtol_conv; s.atol = o->atol; s.yconv_cri = o->yconv_cri; s.slope_cri = o->slope_cri;
    s.yconv_min = o->yconv_min; s.flux_cri = o->flux_cri; s.trun_min = o->trun_min; s.runtime = o->runtime;
    s.conv_step = o->conv_step; s.count_min = o->count_min; s.count_max = o->count_max;
    s.use_photo = o->use_photo; s.ini_frq = o->ini_update_photo_frq; s.final_frq = o->final_update_photo_frq; s.update_frq = o->update_frq;
    s.pref_indx = o->pref_indx; s.gs = o->gs; s.Rp = o->Rp; s.max_flux = o->max_flux;
    s.hist_cap = o->hist_cap; s.hist_stride = o->hist_stride; s.cap_t = o->count_max + 2;
    s.n_diff_esc = o->n_diff_esc;
    s.use_condense = o->use_condense; s.use_fix = o->fix_species_switch; s.fix_from_coldtrap = o->fix_from_coldtrap; s.n_fix = o->n_fix;
    s.start_conden_time = o->start_conden_time; s.stop_conden_time = o->stop_conden_time; s.post_conden_rtol = o->post_conden_rtol;
    if (s.use_condense && !c->conden) { set_error("use_condense needs vk_conden_setup"); return VK_ERR_INVALID; }
    if (s.use_condense && s.use_fix && (!c->opts.fix_mask || !c->opts.fix_y)) { set_error("the fix_species switch writes the handle's fix_mask / fix_y: pass (zeroed) arrays to vk_set_step_opts"); return VK_ERR_INVALID; }
    if (s.use_condense && c->atm.csn == 0 && c->ncol > 1) { set_error("the switch zeroes atm.vs per column: vk_set_atm with shared = 0"); return VK_ERR_INVALID; }
    int rc = VK_OK;
    s.ignore_sp = nullptr; s.diff_esc_idx = nullptr;
    if (o->conv_ignore_sp) rc = scopy(e, o->conv_ignore_sp, ni, &s.ignore_sp);
    if (rc == VK_OK && o->n_diff_esc > 0) rc = scopy(e, o->diff_esc_idx, (size_t)o->n_diff_esc, &s.diff_esc_idx);
    if (rc == VK_OK) rc = scopy(e, o->pico, nz + 1, &s.pico);
    if (rc == VK_OK) rc = scopy(e, o->ms, ni, &s.ms);
    if (rc == VK_OK) rc = scopy(e, o->zco, ncol * (nz + 1), &s.zco);
    if (rc == VK_OK) rc = scopy(e, o->Hp, ncol * nz, &s.Hp);
    if (rc == VK_OK) rc = scopy(e, o->dz, ncol * nz, &s.dz);
    const double *nd = nullptr; const int *ni_ = nullptr;
    s.fix_sp = nullptr; s.fix_whole = nullptr; s.fix_sat_mix = nullptr;
    if (rc == VK_OK && s.n_fix > 0) rc = scopy(e, o->fix_sp, (size_t)s.n_fix, &s.fix_sp);
    if (rc == VK_OK && s.n_fix > 0) rc = scopy(e, o->fix_whole_column, (size_t)s.n_fix, &s.fix_whole);
    if (rc == VK_OK && s.n_fix > 0) rc = scopy(e, o->fix_sat_mix, (size_t)s.n_fix * nz, &s.fix_sat_mix);
    if (rc == VK_OK) rc = scopy(e, nd, ncol, &s.dt_used);
    if (rc == VK_OK) rc = scopy(e, nd, ncol, &s.rtol_col);
    if (rc == VK_OK) rc = scopy(e, nd, ncol * nz, &s.mu);
    if (rc == VK_OK) rc = scopy(e, nd, ncol, &s.longdy);
    if (rc == VK_OK) rc = scopy(e, nd, ncol, &s.longdydt);
    if (rc == VK_OK) rc = scopy(e, nd, ncol, &s.aflux_change);
    if (rc == VK_OK) rc = scopy(e, nd, ncol * (size_t)s.cap_t, &s.t_time);
    if (rc == VK_OK) rc = scopy(e, nd, ncol * (size_t)s.hist_cap * nz * ni, &s.hist);
    int **ints[] = {&s.end_case, &s.act, &s.fresh, &s.do_photo, &s.do_mu, &s.photo_frq, &s.n_left, &s.fix_started, &s.do_conden, &s.do_switch};
    for (int **p : ints)
        if (rc == VK_OK) rc = scopy(e, ni_, ncol, p);
    if (rc != VK_OK) return rc;
    // the cudaMemset calls above run on the legacy default stream, the handle's stream is non-blocking: without this barrier a memset
    // may land AFTER the fill kernels below (seen: act reset to 0 -> a batch that never starts)
    VK_CUDA(cudaDeviceSynchronize());
    const int nb = (int)((ncol + 127) / 128);
    fill_kernel<<<nb, 128, 0, c->stream>>>((int)ncol, s.longdy, 1.0, s.act, 1);             // store.py:36-38: longdy = longdydt = 1
    fill_kernel<<<nb, 128, 0, c->stream>>>((int)ncol, s.longdydt, 1.0, s.fresh, 1);
    fill_kernel<<<nb, 128, 0, c->stream>>>((int)ncol, nullptr, 0.0, s.photo_frq, s.ini_frq > 0 ? s.ini_frq : 1);
    fill_kernel<<<nb, 128, 0, c->stream>>>((int)ncol, s.rtol_col, e->rtol, nullptr, 0);
    VK_CUDA(cudaGetLastError());
    VK_CUDA(cudaStreamSynchronize(c->stream));
    e->steady_set = true;
    return VK_OK;
}

// one photolysis update of every active column from the resident state (vulcan.py:170-176 does one at set-up, before the loop)
int vk_ens_photo_update(vk_column *c)
{
    if (!c || !c->ens || !c->ens->steady_set || !c->photo) { set_error("steady state driver / photolysis not set up"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    SteadyDev &s = c->ens->steady;
    VK_CUDA(cudaEventRecord(c->ev0, c->stream));
    int rc = photo_update_device(c, c->y, c->ymix, s.dz, s.act, s.aflux_change);
    if (rc) return rc;
    VK_CUDA(cudaEventRecord(c->ev3, c->stream));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&c->last_ms_total, c->ev0, c->ev3);       // vk_last_kernel_ms: the update (flux_kernel + jrate_kernel) on the device
    return VK_OK;
}

int vk_ens_run_steady(vk_column *c, int max_iterations, int *n_active_left)
{
    if (!c || !c->ens || !c->ens->steady_set || max_iterations < 0) { set_error("steady state driver not set up"); return VK_ERR_INVALID; }
    if (!c->atm_set || !c->k_set) { set_error("vk_set_atm / vk_set_k must be called first"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    EnsState *e = c->ens;
    SteadyDev &s = e->steady;
    cudaEvent_t run0, run1;
    VK_CUDA(cudaEventCreate(&run0));
    VK_CUDA(cudaEventCreate(&run1));
    VK_CUDA(cudaEventRecord(run0, c->stream));
    int rc = VK_OK;
    c->act = s.act;
    for (int it = 0; it < max_iterations && rc == VK_OK; it++) {
        if (c->use_cr) {      // latency path: cyclic reduction while dt is below cr_dt_max (one small read-back per step of ONE column)
            double hdt = 0.0;
            VK_CUDA(cudaMemcpyAsync(&hdt, c->dt, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            VK_CUDA(cudaStreamSynchronize(c->stream));
            c->cr_now = hdt <= c->cr_dt_max;
            c->dt_host_max = hdt;
        } else {
            c->dt_host_max = -1.0;
        }
        PreArgs pa{c->nz, c->ni, s, c->y, c->ymix, e->n_0, e->t, e->n_accept, c->atm.Kzz, c->atm.cs1};
        steady_pre_kernel<<<c->ncol, 256, 0, c->stream>>>(pa);
        if (s.use_photo) rc = photo_update_device(c, c->y, c->ymix, s.dz, s.do_photo, s.aflux_change);
        if (rc == VK_OK) rc = vk_step_device_impl(c);
        if (rc == VK_OK) rc = launch_clip(c, c->sol, c->ymix_out, c->ymix_out, e->na, e->compo, nullptr, e->pos_cut, e->nega_cut, e->atom_sum,
                                          e->small_y, e->nega_y, e->anyneg);
        if (rc == VK_OK) rc = launch_ens_control(c);
        if (rc == VK_OK && s.use_condense) {
            // op.py:856-901 on the flagged columns: growth rates, then the switch (records the state BEFORE the relaxation operators), then relaxation
            rc = conden_device(c, c->sol, c->ymix_out, s.dt_used, e->n_0, s.do_conden, nullptr, 1);
            if (rc == VK_OK && s.use_fix) {
                SwitchArgs sa{c->nz, c->ni, s, c->atm, c->sol, e->n_0, const_cast<unsigned char *>(c->opts.fix_mask), const_cast<double *>(c->opts.fix_y)};
                steady_switch_kernel<<<c->ncol, 128, 0, c->stream>>>(sa);
                rc = launch_atm_pre_pred(c, s.do_switch);
            }
            if (rc == VK_OK) rc = conden_device(c, c->sol, c->ymix_out, s.dt_used, e->n_0, s.do_conden, nullptr, 2);
        }
        if (rc == VK_OK && s.update_frq > 0) {
            MuArgs ma{c->nz, c->ni, s, c->atm, c->ymix_out, c->sol};
            steady_mu_dz_kernel<<<c->ncol, 128, 0, c->stream>>>(ma);
            rc = launch_atm_pre_pred(c, s.do_mu);
        }
        if (rc == VK_OK) rc = launch_ens_apply(c);
        cudaError_t ce = cudaGetLastError();
        if (rc == VK_OK && ce != cudaSuccess) rc = cuda_fail(ce, "steady state loop kernels");
    }
    c->act = nullptr;
    if (rc == VK_OK) {
        cudaMemsetAsync(s.n_left, 0, sizeof(int), c->stream);
        count_active_kernel<<<1, 256, 0, c->stream>>>(c->ncol, s.act, s.n_left);
    }
    cudaEventRecord(run1, c->stream);
    { const int rcw = stream_wait(c); if (rc == VK_OK) rc = rcw; }        // (batches: blocking wait, the host thread sleeps)
    if (rc == VK_OK) cudaEventElapsedTime(&c->last_ms_total, run0, run1);
    if (rc == VK_OK && n_active_left) VK_CUDA(cudaMemcpy(n_active_left, s.n_left, sizeof(int), cudaMemcpyDeviceToHost));
    cudaEventDestroy(run0);
    cudaEventDestroy(run1);
    return rc;
}

int vk_ens_get_steady(vk_column *c, int *end_case, double *longdy, double *longdydt, double *aflux_change, double *dz, double *zco)
{
    if (!c || !c->ens || !c->ens->steady_set) { set_error("steady state driver not set up"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    SteadyDev &s = c->ens->steady;
    VK_CUDA(cudaStreamSynchronize(c->stream));
    const size_t ncol = c->ncol;
    if (end_case) VK_CUDA(cudaMemcpy(end_case, s.end_case, sizeof(int) * ncol, cudaMemcpyDeviceToHost));
    if (longdy) VK_CUDA(cudaMemcpy(longdy, s.longdy, sizeof(double) * ncol, cudaMemcpyDeviceToHost));
    if (longdydt) VK_CUDA(cudaMemcpy(longdydt, s.longdydt, sizeof(double) * ncol, cudaMemcpyDeviceToHost));
    if (aflux_change) VK_CUDA(cudaMemcpy(aflux_change, s.aflux_change, sizeof(double) * ncol, cudaMemcpyDeviceToHost));
    if (dz) VK_CUDA(cudaMemcpy(dz, s.dz, sizeof(double) * ncol * c->nz, cudaMemcpyDeviceToHost));
    if (zco) VK_CUDA(cudaMemcpy(zco, s.zco, sizeof(double) * ncol * (c->nz + 1), cudaMemcpyDeviceToHost));
    return VK_OK;
}

int vk_ens_get_fix(vk_column *c, int *fix_started, unsigned char *fix_mask, double *fix_y)
{
    if (!c || !c->ens || !c->ens->steady_set) { set_error("steady state driver not set up"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    SteadyDev &s = c->ens->steady;
    VK_CUDA(cudaStreamSynchronize(c->stream));
    const size_t ncol = c->ncol, per = (size_t)c->nz * c->ni;
    if (fix_started) VK_CUDA(cudaMemcpy(fix_started, s.fix_started, sizeof(int) * ncol, cudaMemcpyDeviceToHost));
    if ((fix_mask || fix_y) && (!c->opts.fix_mask || !c->opts.fix_y)) { set_error("no fix_mask / fix_y on this handle"); return VK_ERR_INVALID; }
    if (fix_mask) VK_CUDA(cudaMemcpy(fix_mask, c->opts.fix_mask, ncol * per, cudaMemcpyDeviceToHost));
    if (fix_y) VK_CUDA(cudaMemcpy(fix_y, c->opts.fix_y, sizeof(double) * ncol * per, cudaMemcpyDeviceToHost));
    return VK_OK;
}

}  // extern "C"
