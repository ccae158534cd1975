// Device-resident batched controller: the per-column part of Integration.__call__ / ODESolver.one_step / step_size
// (op.py:808-935, 2489-2534, 3091-3125) for `ncol` independent columns, without any host round trip inside the loop.
//
// One iteration = one ATTEMPTED Ros2 step of every column:
//   solver (vk_step_device) -> clip + loss (op.py:2447-2487) -> accept test step_ok (op.py:2489-2493)
//     accepted: t += dt, count += 1, hydrostatic rescale y = n_0 * ymix (op.py:909-914), new dt from step_size (op.py:3105-3125)
//     rejected: y restored, dt *= dt_var_min, counters (op.py:2495-2534); retried by the next iteration.  As in the reference
//               (reset_y restores y but not ymix, op.py:2530) the mixing ratios of the rejected solution are kept for the
//               significance mask of the retry.
#include "vk_ens_state.cuh"

namespace vk {

int vk_step_device_impl(vk_column *c);
int launch_clip(vk_column *c, double *y_dev, const double *ymix_in_dev, double *ymix_out_dev, int na, const double *compo_dev,
                const unsigned char *skip_dev, double pos_cut, double nega_cut, double *atom_sum_dev, double *small_dev,
                double *nega_dev, int *anyneg_dev);

void ens_destroy(vk_column *c)
{
    if (!c->ens) return;
    for (void *p : c->ens->allocs) cudaFree(p);
    delete c->ens;
    c->ens = nullptr;
}

struct CtlArgs {
    int ncol, na;
    double rtol, loss_eps, dt_min, dt_max, dt_var_min, dt_var_max;
    const double *atom_sum, *atom_ini, *delta;
    const int *anyneg, *status;
    double *atom_loss_prev, *dt, *t;
    int *accept, *n_accept, *n_reject, *n_delta, *n_nega, *n_loss;
    // steady-state driver (vk_steady.cu) or NULL / 0: columns that stopped are skipped; an accepted step records its time, flags the
    // update_mu_dz cadence (op.py:904: count % update_frq == 0 with the count BEFORE save_step) and marks the column fresh
    const int *act;
    int *fresh, *do_mu;
    double *t_time; int cap_t, update_frq;
    // condensation in the loop: flags for the operators that follow an accepted step, evaluated with the model time BEFORE save_step
    // like the reference (op.py:856, 860); rtol_col: the rtol step_size reads (vulcan_cfg.rtol, changed by the switch; step_ok keeps the
    // value frozen in its default argument, op.py:2489)
    int use_condense, use_fix;
    double start_conden_time, stop_conden_time, post_conden_rtol;
    const int *fix_started;
    int *do_conden, *do_switch;
    double *dt_used;
    const double *rtol_col;
};

__global__ void control_kernel(CtlArgs a)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= a.ncol) return;
    if (a.act && !a.act[col]) { a.accept[col] = 0; if (a.do_mu) a.do_mu[col] = 0; return; }
    const double delta = a.delta[col];
    double dloss = 0.0;
    double loss[8];
    for (int q = 0; q < a.na; q++) {
        loss[q] = (a.atom_sum[col * a.na + q] - a.atom_ini[col * a.na + q]) / a.atom_ini[col * a.na + q];   // op.py:2485
        dloss = fmax(dloss, fabs(loss[q] - a.atom_loss_prev[col * a.na + q]));
    }
    const bool neg = a.anyneg[col] != 0;
    bool ok = !neg && (dloss < a.loss_eps) && (delta <= a.rtol) && a.status[col] == 0;                       // op.py:2490
    double dt = a.dt[col];
    int acc = 0;
    if (ok) {
        acc = 1;
    } else {
        if (delta > a.rtol) a.n_delta[col] += 1;                                                              // op.py:2497-2509
        else if (neg) a.n_nega[col] += 1;
        else a.n_loss[col] += 1;
        a.n_reject[col] += 1;
        dt *= a.dt_var_min;                                                                                   // op.py:2531
        // give up (op.py:2515-2520): the reference moves on with dt_min; Integration then rescales y = n_0 * ymix with the
        // mixing ratios of the failed attempt (ymix is not restored by reset_y), i.e. the attempt is effectively accepted
        if (dt < a.dt_min) { dt = a.dt_min; acc = 2; }
    }
    bool sw = false;
    if (a.do_conden) {
        const double t_before = a.t[col];
        const int dc = (acc && a.use_condense && t_before >= a.start_conden_time && !a.fix_started[col]) ? 1 : 0;
        a.do_conden[col] = dc;
        sw = dc && a.use_fix && t_before > a.stop_conden_time;
        a.do_switch[col] = sw ? 1 : 0;
        a.dt_used[col] = dt;
    }
    if (a.fresh) a.fresh[col] = acc ? 1 : 0;
    if (a.do_mu) a.do_mu[col] = (acc && a.update_frq > 0 && (a.n_accept[col] % a.update_frq) == 0) ? 1 : 0;
    if (acc) {
        a.t[col] += dt;                                                                                        // save_step op.py:1091
        if (a.t_time && a.n_accept[col] < a.cap_t) a.t_time[(size_t)col * a.cap_t + a.n_accept[col]] = a.t[col];
        a.n_accept[col] += 1;
        {
            for (int q = 0; q < a.na; q++) a.atom_loss_prev[col * a.na + q] = loss[q];                        // backup op.py:941
            // step_size runs after the switch inside the same iteration (op.py:866, 925), so that iteration already sees post_conden_rtol
            const double rt = sw ? a.post_conden_rtol : (a.rtol_col ? a.rtol_col[col] : a.rtol);
            double d = (delta == 0) ? 0.01 * rt : delta;                                                      // step_size op.py:3113-3120
            double hf = 0.9 * sqrt(rt / d);
            hf = fmax(hf, a.dt_var_min);
            hf = fmin(hf, a.dt_var_max);
            dt = dt * hf;
            dt = fmax(dt, a.dt_min);
            dt = fmin(dt, a.dt_max);
        }
    }
    a.dt[col] = dt;
    a.accept[col] = acc;
}

struct ApplyArgs {
    int nz, ni, n_gas;
    const int *gas_indx;
    const int *accept;
    const double *sol, *ymix_new, *n_0;
    double *y, *ymix;
    const int *act;
    // history ring of the steady-state driver: the state after accepted step number c = n_accept - 1 goes to slot (c / stride) % cap
    // when c is a multiple of stride
    double *hist; int hist_cap, hist_stride;
    const int *n_accept;
};
__global__ void apply_kernel(ApplyArgs a)
{
    const int col = blockIdx.y;
    if (a.act && !a.act[col]) return;
    const int acc = a.accept[col];
    const size_t per = (size_t)a.nz * a.ni;
    double *hslot = nullptr;
    if (acc && a.hist) {
        const int cidx = a.n_accept[col] - 1;              // control_kernel has already counted this step
        if (cidx % a.hist_stride == 0) hslot = a.hist + ((size_t)col * a.hist_cap + (cidx / a.hist_stride) % a.hist_cap) * per;
    }
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < per; q += (size_t)gridDim.x * blockDim.x) {
        const size_t g = col * per + q;
        const double ym = a.ymix_new[g];
        a.ymix[g] = ym;                                    // kept even when the step is rejected (op.py:2530 quirk)
        if (acc) {
            const int j = (int)(q / a.ni), i = (int)(q % a.ni);
            bool gas = true;
            if (a.n_gas > 0) {
                gas = false;
                for (int s = 0; s < a.n_gas; s++) gas = gas || (a.gas_indx[s] == i);
            }
            const double v = gas ? a.n_0[(size_t)col * a.nz + j] * ym : a.sol[g];   // hydrostatic rescale op.py:909-914
            a.y[g] = v;
            if (hslot) hslot[q] = v;                       // save_step: y_time.append(var.y) (op.py:1093)
        }
    }
}

int launch_ens_control(vk_column *c)
{
    EnsState *e = c->ens;
    const bool st = e->steady_set && c->act;
    CtlArgs a{c->ncol, e->na, e->rtol, e->loss_eps, e->dt_min, e->dt_max, e->dt_var_min, e->dt_var_max, e->atom_sum, e->atom_ini,
              c->delta, e->anyneg, c->status, e->atom_loss_prev, c->dt, e->t, e->accept, e->n_accept, e->n_reject, e->n_delta,
              e->n_nega, e->n_loss, c->act, st ? e->steady.fresh : nullptr, st ? e->steady.do_mu : nullptr,
              st ? e->steady.t_time : nullptr, st ? e->steady.cap_t : 0, st ? e->steady.update_frq : 0,
              st ? e->steady.use_condense : 0, st ? e->steady.use_fix : 0, st ? e->steady.start_conden_time : 0.0,
              st ? e->steady.stop_conden_time : 0.0, st ? e->steady.post_conden_rtol : 0.0, st ? e->steady.fix_started : nullptr, (st && e->steady.use_condense) ? e->steady.do_conden : nullptr,
              st ? e->steady.do_switch : nullptr, st ? e->steady.dt_used : nullptr, (st && e->steady.use_condense) ? e->steady.rtol_col : nullptr};
    control_kernel<<<(c->ncol + 127) / 128, 128, 0, c->stream>>>(a);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

int launch_ens_apply(vk_column *c)
{
    EnsState *e = c->ens;
    const bool st = e->steady_set && c->act;
    ApplyArgs b{c->nz, c->ni, c->atm.n_gas, c->atm.gas_indx, e->accept, c->sol, c->ymix_out, e->n_0, c->y, c->ymix, c->act,
                st ? e->steady.hist : nullptr, st ? e->steady.hist_cap : 1, st ? e->steady.hist_stride : 1, e->n_accept};
    dim3 grid((unsigned)std::min<size_t>(((size_t)c->nz * c->ni + 255) / 256, 64), (unsigned)c->ncol);
    apply_kernel<<<grid, 256, 0, c->stream>>>(b);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

template <typename T>
static int ecopy(EnsState *e, const T *host, size_t n, T **out)
{
    void *d = nullptr;
    VK_CUDA(cudaMalloc(&d, sizeof(T) * (n ? n : 1)));
    e->allocs.push_back(d);
    if (host) VK_CUDA(cudaMemcpy(d, host, sizeof(T) * n, cudaMemcpyHostToDevice));
    else VK_CUDA(cudaMemset(d, 0, sizeof(T) * (n ? n : 1)));
    *out = reinterpret_cast<T *>(d);
    return VK_OK;
}

}  // namespace vk

using namespace vk;

extern "C" {

int vk_ens_setup(vk_column *c, const vk_ens_opts *o)
{
    if (!c || !o || !o->compo || !o->atom_ini || !o->n_0 || o->na < 1 || o->na > 8) { set_error("bad ensemble options"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    ens_destroy(c);
    EnsState *e = new EnsState();
    e->steady_set = false;
    c->ens = e;
    e->rtol = o->rtol; e->loss_eps = o->loss_eps; e->dt_min = o->dt_min; e->dt_max = o->dt_max; e->dt_var_min = o->dt_var_min;
    e->dt_var_max = o->dt_var_max; e->pos_cut = o->pos_cut; e->nega_cut = o->nega_cut; e->na = o->na;
    const size_t ncol = c->ncol;
    int rc = ecopy(e, o->compo, (size_t)c->ni * o->na, &e->compo);
    if (rc == VK_OK) rc = ecopy(e, o->atom_ini, ncol * o->na, &e->atom_ini);
    if (rc == VK_OK) rc = ecopy(e, o->n_0, ncol * c->nz, &e->n_0);
    const double *nd = nullptr; const int *ni_ = nullptr;
    if (rc == VK_OK) rc = ecopy(e, nd, ncol * o->na, &e->atom_sum);
    if (rc == VK_OK) rc = ecopy(e, nd, ncol * o->na, &e->atom_loss_prev);
    if (rc == VK_OK) rc = ecopy(e, nd, ncol, &e->small_y);
    if (rc == VK_OK) rc = ecopy(e, nd, ncol, &e->nega_y);
    if (rc == VK_OK) rc = ecopy(e, nd, ncol, &e->t);
    int **ints[] = {&e->anyneg, &e->accept, &e->n_accept, &e->n_reject, &e->n_delta, &e->n_nega, &e->n_loss};
    for (int **p : ints)
        if (rc == VK_OK) rc = ecopy(e, ni_, ncol, p);
    if (rc != VK_OK) ens_destroy(c);
    else VK_CUDA(cudaDeviceSynchronize());      // memsets of the legacy stream before anything runs on the handle's non-blocking stream
    return rc;
}

__global__ void ymix_kernel(int nz, int ni, int n_gas, const int *gas, const double *y, double *ymix)
{
    // plain (non-pairwise) row sums are sufficient to initialise the significance mask
    const size_t row = blockIdx.x;
    const double *yr = y + row * ni;
    __shared__ double s;
    if (threadIdx.x == 0) {
        double acc = 0.0;
        if (n_gas > 0) for (int q = 0; q < n_gas; q++) acc += yr[gas[q]];
        else for (int q = 0; q < ni; q++) acc += yr[q];
        s = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ni; i += blockDim.x) ymix[row * ni + i] = yr[i] / s;
}

int vk_ens_set_state(vk_column *c, const double *y, const double *dt)
{
    if (!c || !c->ens || !y || !dt) { set_error("ensemble not set up / null buffer"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    const size_t nv = (size_t)c->ncol * c->nz * c->ni;
    VK_CUDA(cudaMemcpyAsync(c->y, y, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(c->dt, dt, sizeof(double) * c->ncol, cudaMemcpyHostToDevice, c->stream));
    ymix_kernel<<<(unsigned)((size_t)c->ncol * c->nz), 64, 0, c->stream>>>(c->nz, c->ni, c->atm.n_gas, c->atm.gas_indx, c->y, c->ymix);
    VK_CUDA(cudaGetLastError());
    EnsState *e = c->ens;
    VK_CUDA(cudaMemsetAsync(e->t, 0, sizeof(double) * c->ncol, c->stream));
    VK_CUDA(cudaMemsetAsync(e->atom_loss_prev, 0, sizeof(double) * c->ncol * e->na, c->stream));
    int *ints[] = {e->n_accept, e->n_reject, e->n_delta, e->n_nega, e->n_loss};
    for (int *p : ints) VK_CUDA(cudaMemsetAsync(p, 0, sizeof(int) * c->ncol, c->stream));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    return VK_OK;
}

int vk_ens_run(vk_column *c, int n_steps)
{
    if (!c || !c->ens || n_steps < 0) { set_error("ensemble not set up"); return VK_ERR_INVALID; }
    if (!c->atm_set || !c->k_set) { set_error("vk_set_atm / vk_set_k must be called first"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    EnsState *e = c->ens;
    cudaEvent_t e0 = c->ev0, e3 = c->ev3;
    cudaEvent_t run0, run1;
    VK_CUDA(cudaEventCreate(&run0));
    VK_CUDA(cudaEventCreate(&run1));
    VK_CUDA(cudaEventRecord(run0, c->stream));
    int rc = VK_OK;
    for (int it = 0; it < n_steps && rc == VK_OK; it++) {
        if (c->use_cr) {      // latency path: cyclic reduction while dt is below cr_dt_max (one small read-back per step of ONE column)
            double hdt = 0.0;
            VK_CUDA(cudaMemcpyAsync(&hdt, c->dt, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            VK_CUDA(cudaStreamSynchronize(c->stream));
            c->cr_now = hdt <= c->cr_dt_max;
            c->dt_host_max = hdt;
        } else {
            c->dt_host_max = -1.0;
        }
        rc = vk_step_device_impl(c);
        if (rc) break;
        rc = launch_clip(c, c->sol, c->ymix_out, c->ymix_out, e->na, e->compo, nullptr, e->pos_cut, e->nega_cut, e->atom_sum,
                         e->small_y, e->nega_y, e->anyneg);
        if (rc) break;
        rc = launch_ens_control(c);
        if (rc == VK_OK) rc = launch_ens_apply(c);
    }
    (void)e0; (void)e3;
    cudaEventRecord(run1, c->stream);
    { const int rcw = stream_wait(c); if (rc == VK_OK) rc = rcw; }        // (batches: blocking wait, the host thread sleeps)
    if (rc == VK_OK) cudaEventElapsedTime(&c->last_ms_total, run0, run1);
    if (rc == VK_OK && n_steps > 0) cudaEventElapsedTime(&c->last_ms_factor, c->ev1, c->ev2);   // factor kernel of the last step
    cudaEventDestroy(run0);
    cudaEventDestroy(run1);
    return rc;
}

int vk_ens_get_state(vk_column *c, double *y, double *t, double *dt, int *n_accept, int *n_reject)
{
    if (!c || !c->ens) { set_error("ensemble not set up"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    EnsState *e = c->ens;
    VK_CUDA(cudaStreamSynchronize(c->stream));
    if (y) VK_CUDA(cudaMemcpy(y, c->y, sizeof(double) * (size_t)c->ncol * c->nz * c->ni, cudaMemcpyDeviceToHost));
    if (t) VK_CUDA(cudaMemcpy(t, e->t, sizeof(double) * c->ncol, cudaMemcpyDeviceToHost));
    if (dt) VK_CUDA(cudaMemcpy(dt, c->dt, sizeof(double) * c->ncol, cudaMemcpyDeviceToHost));
    if (n_accept) VK_CUDA(cudaMemcpy(n_accept, e->n_accept, sizeof(int) * c->ncol, cudaMemcpyDeviceToHost));
    if (n_reject) VK_CUDA(cudaMemcpy(n_reject, e->n_reject, sizeof(int) * c->ncol, cudaMemcpyDeviceToHost));
    return VK_OK;
}

}  // extern "C"
