/* vulcan_b200 — C ABI of the B200-native Ros2 hot path of VULCAN (exoclime/VULCAN).
 *
 * Drop-in boundary: the reference has no native interface for this path — the solver is a Python
 * class looked up by name (`getattr(op, vulcan_cfg.ode_solver)()`, vulcan.py:162-163) whose methods call the
 * sympy-generated chem_funs.py and scipy.linalg.solve_banded.  The entry points below are what a ctypes binding
 * of that class needs; each one cites the reference routine it replaces.  All pointers are HOST pointers unless the
 * name ends in `_dev`; all arrays are C-order fp64 (or int32 where stated); nothing is thrown across the ABI: every
 * function returns 0 on success or a negative vk_status, and vk_last_error() returns a message for the calling thread.
 * A handle is bound to one CUDA device and one stream and is not thread-safe.
 *
 * Layouts:  y, ymix, sol ...  [ncol][nz][ni]      k  [ncol][nz][nr+1]  (layer-major; slot 0 unused; the reference's
 * var.k is a dict {1..nr -> (nz,)} (store.py:25, op.py:166-186) — the Python host packs it)
 */
#ifndef VULCAN_B200_H
#define VULCAN_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define VK_ABI_VERSION 3

typedef enum {
    VK_OK = 0,
    VK_ERR_INVALID = -1,    /* bad argument */
    VK_ERR_CUDA = -2,       /* CUDA runtime error (no device, launch failure, out of memory) */
    VK_ERR_UNSUPPORTED = -3,/* network too large for the compiled kernel set / flag combination not built */
    VK_ERR_SINGULAR = -4    /* a diagonal block was exactly singular during factorisation */
} vk_status;

typedef struct vk_network vk_network; /* compiled reaction network on one device */
typedef struct vk_column vk_column;   /* a batch of ncol independent columns (ncol = 1: the drop-in case) */

int vk_abi_version(void);
const char *vk_last_error(void);
/* number of CUDA devices visible; <0 on error.  The product has NO CPU fallback: every compute entry point fails
 * with VK_ERR_CUDA when there is no device. */
int vk_device_count(void);

/* ---- network: replaces the generated chem_funs.py (make_chem_funs.py:113-717) --------------------------------
 * Tables come from vulcan_b200/network.py::Network.tables(): the rate-of-progress monomials, the per-species
 * production/loss lists in reference summation order, and the analytic Jacobian term lists. */
typedef struct {
    int ni, nr;               /* species, reactions (forward+reverse) */
    int maxf, maxjf;          /* factor slots per rate term / per Jacobian term */
    int n_rhs, n_ent, n_term; /* lengths of rhs_pair, jac_row, jac_k */
    const int *rate_fac;      /* [nr+1][maxf]  species index | ni = third body M | ni+1 = 1.0 */
    const int *rate_pow;      /* [nr+1][maxf] */
    const int *rhs_ptr;       /* [ni+1] */
    const int *rhs_pair;      /* [n_rhs] forward reaction id j (odd): term = coef*(rate[j]-rate[j+1]) */
    const double *rhs_coef;   /* [n_rhs] */
    const int *jac_ptr;       /* [n_ent+1] */
    const int *jac_row;       /* [n_ent] */
    const int *jac_col;       /* [n_ent] */
    const int *jac_k;         /* [n_term] */
    const double *jac_coef;   /* [n_term] */
    const int *jac_fac;       /* [n_term][maxjf] */
} vk_network_desc;

int vk_network_create(const vk_network_desc *desc, int device, vk_network **out);
void vk_network_destroy(vk_network *net);

/* ---- rate coefficients on the device: replaces ReadRate.read_rate / lim_lowT_rates / rev_rate / remove_rate (op.py:63-342)
 * and the generated chem_funs.Gibbs (make_chem_funs.py:568-580, thermo/gibbs_text.txt).  Tables come from
 * vulcan_b200/rates.py::RateTable; all arrays are indexed by reaction PAIR p (forward id 2p+1, reverse 2p+2). */
typedef struct {
    int npair;
    const int *kind;            /* [npair] 0 zero (photo/ion/condensation rows), 1 a T^n exp(-E/T), 2 + high-pressure limit, 3 OH+CH3+M */
    const double *arrhenius;    /* [npair][6] a, n, E, a_inf, n_inf, E_inf (op.py:154-163) */
    const int *cap_kind;        /* [npair] low-temperature limits (op.py:320-342): 0 none, 1 Lindemann form, 2 constant */
    const double *cap;          /* [npair][2] threshold temperature, value */
    const unsigned char *reverse; /* [npair] reverse rate k_f / K_eq (even id < stop_rev_indx, op.py:289-304) */
    const unsigned char *removed; /* [npair] bit 0: forward id in remove_list, bit 1: reverse id in remove_list (op.py:311-317) */
    const int *gibbs_ptr;       /* [npair+1] */
    const int *gibbs_sp;        /* species of every Gibbs term, written order, M skipped */
    const int *gibbs_nu;        /* -n for reactants, +n for products */
    const int *dnu;             /* [npair] n_reac - n_prod: K_eq carries (kb T / 1e6)^dnu */
    const double *nasa9;        /* [ni][20] NASA-9 coefficients, 200-1000 K then 1000-6000 K (thermo/NASA9/<species>.txt) */
} vk_rate_desc;
int vk_rates_set(vk_network *net, const vk_rate_desc *desc);

/* ---- columns --------------------------------------------------------------------------------------------- */
int vk_column_create(vk_network *net, int nz, int ncol, vk_column **out);
void vk_column_destroy(vk_column *col);

/* Transport / boundary-condition view: the store.AtmData fields read by ODESolver.diffdf* and lhs_jac_*
 * (op.py:1438-2444).  Arrays are per column ([ncol][...]) unless `shared` is non-zero (then one copy serves all). */
typedef struct {
    int shared;
    int use_moldiff, use_settling, use_topflux, use_botflux; /* vulcan_cfg flags (op.py:2869-2888, 1589-1595) */
    int n_gas;               /* species summed into ysum for diffdf / ymix (0 = all)  op.py:1505-1507, 2990-2993 */
    const int *gas_indx;     /* [n_gas] */
    int n_gas_lhs;           /* same list as used by the lhs variant (op.py:1981-1984) */
    const int *gas_indx_lhs;
    const double *Kzz, *vz, *dzi;  /* [nz-1] */
    const double *Dzz, *vs;        /* [nz-1][ni] */
    const double *Tco, *g, *M;     /* [nz] */
    const double *Ti, *Hpi;        /* [nz-1] */
    const double *ms, *alpha, *top_flux, *bot_flux, *bot_vdep; /* [ni] */
    /* ABI 2: vulcan_cfg.use_vm_mol selects diffdf_vm / lhs_jac_tot_vm or, with use_settling, diffdf_settling_vm /
     * lhs_jac_settling_vm (op.py:2879-2888; 1599-1694, 1794-1898, 2044-2119, 2366-2444) */
    int use_vm_mol;
    const double *vm;        /* [nz][ni] atm.vm, the advective velocity of molecular diffusion (build_atm.py:735-739); NULL unless use_vm_mol */
    int n_diff_esc;          /* vulcan_cfg.diff_esc: species with diffusion-limited escape; only the *_vm lhs variants read it (op.py:2101-2107) */
    const int *diff_esc_idx; /* [n_diff_esc] */
} vk_atm_view;
int vk_set_atm(vk_column *col, const vk_atm_view *atm);

/* rate coefficients: var.k packed.  shared = 1: one copy [nz][nr+1] serves every column; 0: [ncol][nz][nr+1]; 2: [ncol][nz][nr+1] with the
 * caller's promise that all rows outside the network's photolysis / ionisation / condensation sections are identical in every column (one
 * T-P profile; only the rows the run itself rewrites - J rates, condensation growth rates - differ): the emitted chemistry kernels then
 * apply to the batch.  vk_set_k_rows with per-column values for any other row withdraws the promise. */
int vk_set_k(vk_column *col, const double *k, int shared);
/* rate coefficients computed ON the device from temperature and total number density (needs vk_rates_set):
 * Tco, M [ncol][nz], or [nz] when shared != 0.  Photolysis rows are zero until vk_photo_update fills them. */
int vk_compute_k(vk_column *col, const double *Tco, const double *M, int shared);
/* read the device copy of k back: [ncol][nz][nr+1], or [nz][nr+1] when it is shared */
int vk_get_k(vk_column *col, double *k);
/* overwrite whole reactions (e.g. the photolysis J rows after compute_J, op.py:2785-2786):
 * rows[n_rows] reaction ids, vals [ncol][n_rows][nz] */
int vk_set_k_rows(vk_column *col, int n_rows, const int *rows, const double *vals);

/* step options: the vulcan_cfg / para state consulted inside Ros2.solver (op.py:2896-2970) */
typedef struct {
    double mtol, atol;               /* vulcan_cfg.mtol / atol (op.py:2949-2950) */
    int refine;                      /* iterative refinement of each linear solve with a double-double residual: 0 none, n > 0 that many
                                      * passes, -1 AUTO: on the columns with dt >= refine_dt_min up to 4 passes, each kept only if it lowers
                                      * the element-weighted residual compo^T (rhs - A x), until the element-budget error of the solve
                                      * is below 1e-11 (needs compo); -n: the same with at most n passes */
    int zero_delta_row0;             /* use_botflux or use_fix_sp_bot: delta[0] = 0 (op.py:2953) */
    int n_fix_bot;                   /* use_fix_sp_bot (op.py:2945-2946) */
    const int *fix_bot_idx;          /* [n_fix_bot] species */
    const double *fix_bot_val;       /* [ncol][n_fix_bot] = mix * n_0[0] */
    const unsigned char *delta_zero_sp; /* [ni] species whose delta is ignored (op.py:2956-2970) or NULL */
    const unsigned char *fix_mask;   /* [ncol][nz][ni] rows pinned to identity (fix_species / electrons) or NULL */
    const double *fix_y;             /* [ncol][nz][ni] values re-imposed where fix_mask (op.py:2960-2968) or NULL */
    /* ABI 3 */
    int na; const double *compo;     /* [ni][na] atoms per species (build_atm.py:18-25, thermo/all_compose.txt) or NULL */
    double refine_dt_min;            /* refine = -1: step size (s) from which a column is refined */
    int rhs_order;                   /* chemdf (chem_funs.py:931): 0 (default) = the fastest kernel this library has for the network: the EMITTED
                                      * straight-line kernel (vulcan_b200/emit.py; reference summation order, bit-identical to the generated
                                      * chemdf) when one is compiled in, else the table-driven kernel with the segmented summation (32 partial
                                      * chains per layer); 1 = reference order (emitted, else table-driven left-to-right); 2 = table-driven,
                                      * reference order; 3 = table-driven, segmented */
} vk_step_opts;
int vk_set_step_opts(vk_column *col, const vk_step_opts *opts);

/* ---- the hot path -------------------------------------------------------------------------------------------
 * vk_ros2_solve  ≙  Ros2.solver (op.py:2860-3007): one ATTEMPTED step of every column in the batch.
 *   in : y, ymix [ncol][nz][ni], dt[ncol]
 *   out: sol, ymix_out [ncol][nz][ni], delta[ncol] (para.delta), status[ncol] (0 or VK_ERR_SINGULAR)
 * Host buffers; H2D/D2H copies are part of the call (the reference's Integration reads var.y every step). */
int vk_ros2_solve(vk_column *col, const double *y, const double *ymix, const double *dt, double *sol,
                  double *ymix_out, double *delta, int *status);

/* ODESolver.clip + loss (op.py:2447-2487) on host arrays.  compo [ni][na] (na <= 8); atom_sum out [ncol][na];
 * small_y / nega_y [ncol] are ACCUMULATED like para.small_y / para.nega_y.  mtol = vulcan_cfg.mtol of the mask
 * y[(ymix < mtol) & (y < 0)] = 0 (op.py:2459); a negative value means "the mtol of vk_set_step_opts". */
int vk_clip_loss(vk_column *col, double *y, const double *ymix_in, double *ymix_out, int na, const double *compo,
                 const unsigned char *atom_skip, double pos_cut, double nega_cut, double mtol, double *atom_sum,
                 double *small_y, double *nega_y, int *any_negative);

/* ---- component entry points (parity tests, diagnostics) ---------------------------------------------------- */
/* chem_funs.chemdf (chem_funs.py:931) + ODESolver.diffdf* (op.py:1438-1898): out_chem/out_diff [ncol][nz][ni],
 * either may be NULL */
int vk_eval_rhs(vk_column *col, const double *y, double *out_chem, double *out_diff);
/* lhs_jac_* (op.py:1973-2444): D [ncol][nz][ni][ni] diagonal blocks, up/dn [ncol][nz][ni] diagonal couplings */
int vk_eval_lhs(vk_column *col, const double *y, const double *dt, double *D, double *up, double *dn);
/* block-tridiagonal factor + solve of a caller-supplied system (replaces store_bandM + solve_banded,
 * op.py:2833-2858, 2914): D [ncol][nz][ni][ni], up/dn/rhs/x [ncol][nz][ni] */
int vk_blocktri_solve(vk_column *col, const double *D, const double *up, const double *dn, const double *rhs,
                      double *x, int refine, int *status);

/* ---- photolysis: compute_tau / compute_flux / compute_J (op.py:2580-2786) ---------------------------------- */
typedef struct {
    int nbin, i12;                 /* var.nbin, var.sflux_din12_indx */
    double dbin1, dbin2;           /* trapezoid weights */
    double sl_angle, edd, flux_atol, f_diurnal; /* vulcan_cfg */
    const double *bins;            /* [nbin] wavelengths (nm) */
    const double *sflux_top;       /* [nbin] */
    int n_abs;  const int *abs_idx;  const double *cross_abs;   /* tau absorbers (photo ∪ ion species): [n_abs][nbin] */
    const unsigned char *abs_is_T; const double *cross_abs_T;   /* optional T-dependent: [n_abs][nz][nbin] */
    int n_photo; const int *photo_idx; const double *cross_photo; /* omega0 absorbers (photo_sp): [n_photo][nbin] */
    int n_scat; const int *scat_idx; const double *cross_scat;  /* Rayleigh: [n_scat][nbin] */
    int n_br;   const double *cross_J; const int *br_rate_index; /* [n_br][nbin], reaction id per branch (0 = skip) */
    const unsigned char *br_is_T; const double *cross_J_T;      /* optional [n_br][nz][nbin] */
} vk_photo_view;
int vk_photo_setup(vk_column *col, const vk_photo_view *pv);
/* one photolysis update of every column: y, ymix [ncol][nz][ni], dz [ncol][nz] ->
 * J [ncol][n_br][nz] (also written into the device copy of k as J*f_diurnal), aflux_change [ncol].
 * The diffuse-flux state (var.dflux_u of the previous call, op.py:2692) lives on the device between calls. */
int vk_photo_update(vk_column *col, const double *y, const double *ymix, const double *dz, double *J,
                    double *aflux_change);
/* optional read-back of the photolysis fields (debug / parity): any pointer may be NULL.
 * tau, sflux, dflux_u, dflux_d [ncol][nz+1][nbin]; aflux [ncol][nz][nbin] */
int vk_photo_read(vk_column *col, double *tau, double *sflux, double *dflux_u, double *dflux_d, double *aflux);
int vk_photo_reset(vk_column *col);

/* ---- device-resident ensemble driver (Integration.__call__ / one_step / step_size batched) -- */
typedef struct {
    double rtol, loss_eps, dt_min, dt_max, dt_var_min, dt_var_max; /* op.py:2489-2534, 3105-3125 */
    double pos_cut, nega_cut;
    int na; const double *compo; const double *atom_ini;          /* [ni][na], [ncol][na] */
    const double *n_0;                                            /* [ncol][nz] hydrostatic rescale (op.py:909-914) */
} vk_ens_opts;
int vk_ens_setup(vk_column *col, const vk_ens_opts *o);
int vk_ens_set_state(vk_column *col, const double *y, const double *dt);
/* n_steps loop iterations of every column entirely on the device: solver -> clip -> accept/reject(+retry next
 * iteration with dt/2) -> rescale -> step_size.  No host round trip inside. */
int vk_ens_run(vk_column *col, int n_steps);
/* accepted / rejected counters [ncol], model time t [ncol], dt [ncol], y [ncol][nz][ni]; any may be NULL */
int vk_ens_get_state(vk_column *col, double *y, double *t, double *dt, int *n_accept, int *n_reject);

/* ---- device-resident run to steady state (Integration.__call__ / stop / conv / save_step / update_mu_dz /
 * update_phi_esc, op.py:808-1105, per column and without a host round trip).  Call after vk_ens_setup + vk_ens_set_state. */
typedef struct {
    double st_factor, mtol_conv, atol, yconv_cri, slope_cri, yconv_min, flux_cri, trun_min, runtime;   /* vulcan_cfg (op.py:1018-1087) */
    int conv_step, count_min, count_max;
    const unsigned char *conv_ignore_sp;     /* [ni] species left out of conv: conver_ignore + non_gas_sp (op.py:1045-1049) or NULL */
    int use_photo, ini_update_photo_frq, final_update_photo_frq;   /* photolysis cadence (op.py:800, 818-829); needs vk_photo_setup */
    int update_frq;                          /* update_mu_dz / update_phi_esc every update_frq accepted steps (op.py:904-906); 0 = never */
    int pref_indx; double gs, Rp, max_flux;  /* atm.pref_indx, atm.gs, vulcan_cfg.Rp, vulcan_cfg.max_flux (op.py:944-999) */
    const double *pico;                      /* [nz+1] atm.pico, shared by the batch */
    const double *ms;                        /* [ni] molar masses as mean_mass reads them (build_atm.py:511-520) */
    const double *zco, *Hp, *dz;             /* [ncol][nz+1], [ncol][nz], [ncol][nz] initial atm.zco / Hp / dz */
    int n_diff_esc; const int *diff_esc_idx; /* vulcan_cfg.diff_esc */
    int hist_cap, hist_stride;               /* ring of accepted states for conv: hist_cap states, every hist_stride-th accepted step
                                              * (hist_cap = conv_step, hist_stride = 1: the reference's look-back exactly) */
    /* condensation in the loop (op.py:856-901; needs vk_conden_setup, and zeroed fix_mask / fix_y in vk_set_step_opts for the switch) */
    int use_condense;                        /* conden + relaxation operators after every accepted step from start_conden_time on */
    int fix_species_switch;                  /* vulcan_cfg.fix_species non-empty: freeze them once t > stop_conden_time (op.py:860-893) */
    int fix_from_coldtrap;                   /* vulcan_cfg.fix_species_from_coldtrap_lev */
    int n_fix; const int *fix_sp;            /* vulcan_cfg.fix_species */
    const unsigned char *fix_whole_column;   /* [n_fix] condensates (H2O_l_s, H2SO4_l, NH3_l_s, S8_l_s): frozen up to layer nz-2 (op.py:878-879) */
    const double *fix_sat_mix;               /* [n_fix][nz] atm.sat_mix of the gas species (cold-trap level, op.py:881-892); zeros for condensates */
    double start_conden_time, stop_conden_time, post_conden_rtol;
} vk_steady_opts;
int vk_ens_setup_steady(vk_column *col, const vk_steady_opts *o);
/* one photolysis update of every active column from the resident state (vulcan.py:170-176 does one at set-up, before the loop) */
int vk_ens_photo_update(vk_column *col);
/* up to max_iterations attempted steps of every column that has not stopped; *n_active_left = columns still running afterwards */
int vk_ens_run_steady(vk_column *col, int max_iterations, int *n_active_left);
/* per column: para.end_case (0 = still running, 1 converged, 2 runtime, 3 count_max), var.longdy, var.longdydt, var.aflux_change,
 * atm.dz [ncol][nz], atm.zco [ncol][nz+1]; any pointer may be NULL */
int vk_ens_get_steady(vk_column *col, int *end_case, double *longdy, double *longdydt, double *aflux_change, double *dz, double *zco);
/* result of the fix_species switch (op.py:860-893) per column: para.fix_species_start, the rows frozen (fix_mask [ncol][nz][ni]; per
 * species the count of set rows is atm.conden_min_lev) and the values they are frozen at (var.fix_y); any pointer may be NULL */
int vk_ens_get_fix(vk_column *col, int *fix_started, unsigned char *fix_mask, double *fix_y);

/* ---- condensation operators of the caller: Integration.conden (op.py:1109-1300) and h2o / nh3_conden_evap_relax
 * (op.py:1340-1421) on the device.  Tables come from the reference's containers (atm.sat_p, atm.r_p, atm.rho_p, var.conden_re_list, var.Rf). */
typedef struct {
    int n_re;                          /* condensation growth reactions `X -> X_l_s` handled by conden */
    const int *re_idx, *gas_idx;       /* [n_re] forward reaction id, gas species */
    const double *m, *rho_p, *r_p;     /* [n_re] molecular mass in g (amu / Navo as the reference writes it), atm.rho_p, atm.r_p */
    const double *sat;                 /* [n_re][nz] saturation number density sat_p / kb / Tco (x humidity for H2O) */
    const unsigned char *zero_rate;    /* [n_re] 1: the relaxation operator replaces the reaction, k[re] = k[re+1] = 0 (op.py:1124-1126) */
    int n_relax;                       /* relaxation operators in call order (op.py:896-901) */
    const int *relax_kind;             /* [n_relax] 1 = h2o_conden_evap_relax, 2 = nh3_conden_evap_relax */
    const int *relax_gas, *relax_ice;  /* [n_relax] vapour / condensate species */
    const int *relax_top;              /* [n_relax] NH3: top layer of the condensation zone, argmin(sat_mix) (op.py:1390); else nz */
    const double *relax_m, *relax_rho, *relax_r;   /* [n_relax] */
    const double *relax_sat;           /* [n_relax][nz] saturation number density */
    double start_conden_time, stop_conden_time, post_conden_rtol;   /* vulcan_cfg (device-resident loop only) */
} vk_conden_desc;
int vk_conden_setup(vk_column *col, const vk_conden_desc *d);
/* conden + the relaxation operators on host state: y, ymix [ncol][nz][ni] in / out, dt [ncol] (the step just taken), n_0 [ncol][nz];
 * k_rows out [ncol][n_re][2][nz] (also written into the device copy of k) or NULL */
int vk_conden_apply(vk_column *col, double *y, double *ymix, const double *dt, const double *n_0, double *k_rows);

/* refine = -1 bookkeeping: refinement passes kept / tried by the safeguard since the handle was created, [ncol] each, may be NULL */
int vk_refine_stats(vk_column *col, int *kept, int *tried);
/* timing of the last vk_ros2_solve / vk_ens_run on the handle's stream, measured with CUDA events (ms) */
int vk_last_kernel_ms(vk_column *col, float *ms_total, float *ms_factor);
/* profiling aid (bench.py, scripts/kernel_times.py): `reps` back-to-back launches of ONE kernel of the step on the resident state with
 * the step's own launch arguments, timed with CUDA events on the handle's stream; *ms = average per launch.
 * which: 0 Jacobian + lhs assembly, 1 right-hand side (stage 1), 2 block-tridiagonal factorisation, 3 / 4 solve sweeps (forward + backward) */
int vk_debug_time_kernel(vk_column *col, int which, int reps, float *ms);
/* raw device pointers for callers that keep state resident (torch tensors): y/ymix/sol [ncol][nz][ni] */
int vk_device_buffers(vk_column *col, void **y_dev, void **ymix_dev, void **sol_dev, void **k_dev);
int vk_stream(vk_column *col, void **cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
