// State of the device-resident controller (vk_ens.cu: attempted step, accept / reject, step size; vk_steady.cu: stop / conv, photolysis
// cadence, update_mu_dz, history).  Internal, not part of the C ABI.
#pragma once
#include "vk_internal.cuh"

namespace vk {
struct SteadyDev {
    double st_factor, mtol_conv, atol, yconv_cri, slope_cri, yconv_min, flux_cri, trun_min, runtime;   // vulcan_cfg, op.py:1018-1087
    int conv_step, count_min, count_max;
    int use_photo, ini_frq, final_frq, update_frq, pref_indx;
    double gs, Rp, max_flux;
    int n_diff_esc;
    unsigned char *ignore_sp;           // [ni] species left out of conv (conver_ignore, non-gas species)
    int *diff_esc_idx;
    double *pico, *ms;                  // [nz+1] pressure at the interfaces, [ni] molar masses (shared by the batch)
    double *zco, *Hp, *mu, *dz;         // [ncol][nz+1], [ncol][nz] x3
    int *end_case, *act, *fresh, *do_photo, *do_mu, *photo_frq, *n_left;   // [ncol]
    double *longdy, *longdydt, *aflux_change;                               // [ncol]
    double *t_time; int cap_t;          // [ncol][cap_t] model time after every accepted step
    double *hist; int hist_cap, hist_stride;   // [ncol][hist_cap][nz*ni] ring of accepted states
    // condensation in the loop (op.py:856-901; operators in vk_conden.cu): conden + relaxation after every accepted step from
    // start_conden_time on, until the fix_species switch (t > stop_conden_time) freezes the condensable species of the column
    int use_condense, use_fix, fix_from_coldtrap, n_fix;
    double start_conden_time, stop_conden_time, post_conden_rtol;
    int *fix_sp;                        // [n_fix] species of vulcan_cfg.fix_species
    unsigned char *fix_whole;           // [n_fix] condensates: frozen up to layer nz-2 whatever the saturation profile (op.py:878-879)
    double *fix_sat_mix;                // [n_fix][nz] atm.sat_mix (gas species)
    int *fix_started, *do_conden, *do_switch;   // [ncol]
    double *dt_used, *rtol_col;         // [ncol] step size of the accepted attempt; rtol read by step_size (post_conden_rtol after the switch)
};
}  // namespace vk

struct EnsState {
    double rtol, loss_eps, dt_min, dt_max, dt_var_min, dt_var_max, pos_cut, nega_cut;
    int na;
    double *compo, *atom_ini, *n_0;       // [ni][na], [ncol][na], [ncol][nz]
    double *atom_sum, *atom_loss_prev;    // [ncol][na]
    double *small_y, *nega_y, *t;         // [ncol]
    int *anyneg, *accept, *n_accept, *n_reject, *n_delta, *n_nega, *n_loss;   // [ncol]
    bool steady_set;
    vk::SteadyDev steady;
    std::vector<void *> allocs;
};
