// vulcan_b200 internal device structures (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/vulcan_b200.h"

#define VK_KB 1.38064852e-16   // phy_const.py:3
#define VK_NAVO 6.02214086e23  // phy_const.py:4
#define VK_HC 1.98644582e-9    // phy_const.py:8
#define VK_RHS_SPL 8            // species slots per lane in the warp-per-layer rhs kernel (ni <= 256)

namespace vk {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
#define VK_STR2(x) #x
#define VK_STR(x) VK_STR2(x)
#define VK_CUDA(call)                                              \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return vk::cuda_fail(_e, #call " at " __FILE__ ":" VK_STR(__LINE__));    \
    } while (0)

// opt-in to more than 48 KB of dynamic shared memory for `func`: the attribute belongs to the device's primary context, so it is
// tracked per (function, device) - a process may drive several devices (Ros2 / DeviceNetwork / EnsembleRunner take `device`) and
// PipelinedHostSolver launches from a thread pool (mutex inside)
int ensure_smem(const void *func, int device, size_t bytes);

// ---- compiled network on the device ---------------------------------------------------------------------------
// factor slots are bytes: species index (< 254), ni = third body M, ni+1 = constant 1.0
struct NetDev {
    int ni, nr, nip;            // nip = padded block size (multiple of 24)
    int n_ent, n_term, n_rhs, max_rhs_len;
    int has_pow;                // any exponent != 1 in the network (rare `n*X` syntax)
    const uchar4 *rate_fac;     // [nr+1] four ordered factor slots
    const uchar4 *rate_pow;     // [nr+1]
    const int *rhs_ptr;         // [ni+1]
    const int *rhs_term;        // [n_rhs]  (forward id << 8) | (coef & 0xff), coef signed 8 bit
    const int *jac_ptr;         // [n_ent+1]
    const ushort2 *jac_rc;      // [n_ent] (row, col)
    const uint2 *jac_term;      // [n_term] .x = k index | (coef8 << 16), .y = 3 factor bytes
    // warp-per-layer rhs kernel: 16-bit term descriptors (bit 15 = minus, bits 0..14 = reaction pair (id-1)/2) when every
    // stoichiometric coefficient is +-1 (rhs_unit), and the species each of the 32 lanes sums (longest chains first, then
    // greedily balanced): rhs_lane_sp[lane][VK_RHS_SPL], -1 = none
    int rhs_unit;
    const unsigned short *rhs_desc16;   // [n_rhs]
    const int *rhs_lane_sp;             // [32][VK_RHS_SPL]
    // segmented summation (rhs_order = 0, the default): the n_rhs terms in species order are cut into 32 equal chunks, one per lane
    // (T = rhs_flat_T terms each, stored transposed: term q of lane l at rhs_flat16[32 q + l]); bit 15 = minus, bit 14 = "flush the partial
    // sum after this term" (end of the species or of the chunk), bits 0..13 = reaction pair; partial sums are numbered in flat order, species
    // s owns the partials [rhs_seg_ptr[s], rhs_seg_ptr[s+1]), lane l starts at partial rhs_lane_slot0[l]
    int rhs_flat_ok, rhs_flat_T, rhs_n_seg;
    const unsigned short *rhs_flat16;   // [32 * T]
    const int *rhs_seg_ptr;             // [ni + 1]
    const int *rhs_lane_slot0;          // [32]
    // Jacobian work schedule: entries cut into segments of <= 16 terms, sorted by length (warp lanes carry equal work)
    int n_seg, n_multi, n_part;
    const uint4 *jac_seg;       // [n_seg]   x = row | col << 16, y = first term, z = n terms | slot << 16 (slot 0xffff: whole entry)
    const uint2 *jac_multi;     // [n_multi] x = row | col << 16, y = first slot | n slots << 16
    // multi-layer Jacobian kernel: the 5953 terms of NCHO share 1603 distinct products k_r y_a y_b y_c
    int n_uniq, lhs_ml_ok;      // lhs_ml_ok: packing limits hold (nr < 2048, ni + 1 < 128, n_uniq < 8192, n_term < 65536, coefficients in the table)
    const unsigned *jac_uniq;   // [n_uniq]  r | f0 << 11 | f1 << 18 | f2 << 25
    // segments in groups of 32 (one per lane), terms stored TRANSPOSED inside a group: term q of lane l at jac_tt[base_g + 32 q + l],
    // padded to the group's longest segment with a descriptor of the always-zero product n_uniq (bank-conflict free, no divergence)
    int n_grp, n_tt;
    const unsigned short *jac_tt;       // [n_tt]   product index | coefficient code << 13
    const unsigned *jac_seg4;           // [n_grp*32] row | col << 8 | slot << 16 (slot 0xffff: whole entry; row 0xff: padding lane)
    const uint2 *jac_grp;               // [n_grp]  x = base into jac_tt, y = terms per lane
};

// atmosphere-only pieces of the transport stencil, [ncol_atm][nz][ni] each (see atm_pre_kernel)
struct AtmPre {
    double *Q, *QB, *QC, *TA, *TB, *TC, *SA, *SB, *SC;
    double *LS;                 // [ncol_atm][nz][10] per-layer scalars (eddy / advection prefactors)
};

// ---- transport view on the device -------------------------------------------------------------------------------
struct AtmDev {
    int nz, ni;
    int use_moldiff, use_settling, use_topflux, use_botflux;
    int n_gas, n_gas_lhs;
    const int *gas_indx, *gas_indx_lhs;
    size_t cs1, csn, csz;       // column strides (elements) of [nz-1], [nz-1][ni] and [nz] arrays; cs_i for [ni]; 0 when shared
    size_t csi;
    const double *Kzz, *vz, *dzi, *Dzz, *vs, *Tco, *g, *M, *Ti, *Hpi, *ms, *alpha, *top_flux, *bot_flux, *bot_vdep;
    AtmPre pre;
    size_t pre_cs;              // column stride of the AtmPre arrays (0 when shared)
    // use_vm_mol (op.py:2879-2888): vm [nz][ni] with column stride csv; in this mode the AtmPre T*/S* arrays hold upwind terms
    int use_vm_mol, n_diff_esc;
    const double *vm;
    size_t csv;
    const int *diff_esc_idx;
};

struct StepOptsDev {
    double mtol, atol;
    int refine, zero_delta_row0, n_fix_bot;
    const int *fix_bot_idx;
    const double *fix_bot_val;        // [ncol][n_fix_bot]
    const unsigned char *delta_zero_sp;
    const unsigned char *fix_mask;    // [ncol][nz][ni]
    const double *fix_y;
    int rhs_order;                    // 0: segmented summation of the production / loss terms (default), 1: the reference's left-to-right order
    int na;                           // refine = auto: element composition for the safeguard
    const double *compo;              // [ni][na]
    double refine_dt_min;
};

}  // namespace vk

struct RateTables;
struct vk_network {
    int device;
    vk::NetDev d;
    std::vector<void *> allocs;
    RateTables *rates;          // on-device rate-coefficient tables (vk_rates_set), optional
    const void *emit;           // emitted chemdf kernel of this network (vk::emitted::EmitEntry, vk_emit.cu) or NULL: table-driven kernels
    unsigned long long table_hash;
};

struct PhotoState;
struct EnsState;
struct CrPlan;
struct CondenState;

struct vk_column {
    vk_network *net;
    int device;                 // copy of net->device: the network handle may be destroyed first (garbage-collection order of the bindings)
    int nz, ncol, ni, nr, nip;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1, ev2, ev3;
    cudaEvent_t ev_sync;        // blocking-sync event: host threads of big batches sleep instead of spinning while the stream drains
    float last_ms_total, last_ms_factor;
    bool last_fused;            // the last step took the fused assembly + factorisation kernel
    // state
    double *y, *ymix, *sol, *ymix_out, *k;   // [ncol][nz][ni] x4, k [ncol or 1][nz][nr+1]
    size_t k_cs;                              // column stride of k (0 = shared)
    bool k_static_shared;                     // per-column k whose rows outside the photolysis / ionisation / condensation sections are identical
                                              // in every column (vk_set_k shared = 2): the emitted chemistry kernels apply
    double *f, *k1, *k2, *yk2, *rhs, *z, *res, *dx, *xn;  // work vectors [ncol][nz][ni]
    double *ysum_lhs_tmp;       // layer sums written by the emitted Jacobian kernel
    double *chem_tmp, *ysum_tmp; void *scal_tmp;   // emitted chemdf path (allocated on first use): chemdf [ncol][nz][ni], layer sums and layer scalars [ncol][nz]
    int *fwd_done; int fwd_valid;   // [ncol] forward elimination of the first solve fused into the factorisation (launch_factor); valid for the last factorisation
    int *refine_kept, *refine_tried, *refine_act;   // [ncol] refinement passes kept / tried by the safeguard, active flags (refine = auto)
    double *D, *W;                            // D [ncol][nz][nip][nip] lhs diagonal blocks; W [ncol][nz][nip][nip+2] block LU factors of the Schur blocks
    double *up, *dn;                          // [ncol][nz][nip]
    double *dt, *delta;                       // [ncol]
    int *status;                              // [ncol]
    vk::AtmDev atm;
    bool atm_set, k_set;
    vk::StepOptsDev opts;
    std::vector<void *> atm_allocs, opt_allocs;
    // pinned host staging
    double *h_pin;
    size_t h_pin_bytes;
    PhotoState *photo;
    EnsState *ens;
    CondenState *conden;        // condensation operators on the device (vk_conden.cu), optional
    CrPlan *cr;                 // single-column latency path: plan + buffers of the block cyclic reduction (vk_cr.inl), built on first use
    double dt_host_max;         // largest step size of the batch when the host knows it for the current step (vk_ros2_solve, the one-column
                                // device loops), else < 0: refine = auto then skips its launches outright below refine_dt_min
    bool use_cr;                // this handle may solve by cyclic reduction (ncol == 1 unless VK_CR=0) ...
    bool cr_now;                // ... and does so for the current step: only while dt <= cr_dt_max.  Beyond it (cond(A) > 1e16) the extra
    double cr_dt_max;           // dense products of the reduction cost accuracy (HD209S at dt = 2.4e5 s: element budget 3.0e-4 against 1.2e-4
                                // for block Thomas, the refinement pass no longer contracts, a whole run loses 1.5e-2 of its carbon), so
                                // those steps take block Thomas.  Default 1e5 s (VK_CR_DT_MAX): HD189 / HD209S runs identical to Thomas-only
    const int *act;             // steady-state driver: columns with act == 0 have stopped and are skipped by every kernel of the step (else NULL)
    // persistent scratch of vk_clip_loss (device doubles / ints + one pinned host mirror)
    double *clip_d, *clip_h;
    int *clip_i;
};

namespace vk {
// kernels (vk_chem.cu)
int launch_rhs(vk_column *c, const double *y_dev, double *out_sum, double *out_chem, double *out_diff,
               const double *k1_for_rhs2, const double *dt_dev);
int launch_lhs(vk_column *c, const double *y_dev, const double *dt_dev, int dense_out_ni, double *D_out, double *up_out,
               double *dn_out);
int launch_factor_fused(vk_column *c, const double *y_dev, const double *dt_dev, double *D_out, double *up_out, double *dn_out, double *F,
                        int *status, int store_D);
int launch_atm_pre(vk_column *c, int ncol_atm);
int launch_atm_pre_pred(vk_column *c, const int *pred);
// kernels (vk_solve.cu)
unsigned long long network_table_hash(const vk_network_desc *d);
const void *emit_lookup(unsigned long long hash, int ni, int nr);
int launch_jac_emitted(vk_column *c, const double *y_dev, double *D_out, double *ysum_out);
bool emit_has_jac(const void *entry);
bool emit_row_is_dynamic(const void *entry, int i);
int launch_chem_emitted(vk_column *c, const double *y_dev, const double *k1, double *chem_out, double *ysum_out, double *yk2_out);
int stream_wait(vk_column *c);
int launch_factor(vk_column *c, const double *D, const double *up, const double *dn, double *F, int *status, const double *fwd_rhs = nullptr);
int launch_solve(vk_column *c, const double *F, const double *up, const double *dn, const double *rhs, double *x, double *z,
                 const int *act = nullptr, const int *fwd_done = nullptr);
int launch_residual(vk_column *c, const double *D, const double *up, const double *dn, const double *rhs, const double *x,
                    double *res, const int *act = nullptr);
int launch_cr_factor(vk_column *c, const double *D, const double *up, const double *dn, double *F, int *status);
int launch_cr_solve(vk_column *c, const double *F, const double *up, const double *dn, const double *rhs, double *x, const int *act);
void cr_plan_free(CrPlan *p);
int launch_refine(vk_column *c, const double *D, const double *up, const double *dn, const double *F, const double *rhs, double *x,
                  int refine, const double *dt_pred);
// kernels (vk_step.cu)
int launch_epilogue(vk_column *c);
}  // namespace vk
