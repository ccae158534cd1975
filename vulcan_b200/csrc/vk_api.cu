// C ABI glue (include/vulcan_b200.h): handles, host<->device staging, kernel sequencing for one attempted Ros2 step.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "vk_internal.cuh"

namespace vk {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char *what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return VK_ERR_CUDA;
}

int ensure_smem(const void *func, int device, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> done;
    if (bytes <= 48 * 1024) return VK_OK;
    std::lock_guard<std::mutex> lock(mu);
    size_t &have = done[std::make_pair(func, device)];
    if (bytes > have) {
        VK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        // co-residency of two blocks needs the full shared-memory carve-out of the SM (the default heuristic sizes it for ONE block)
        VK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        have = bytes;
    }
    return VK_OK;
}

int launch_clip(vk_column *c, double *y_dev, const double *ymix_in_dev, double *ymix_out_dev, int na, const double *compo_dev,
                const unsigned char *skip_dev, double pos_cut, double nega_cut, double *atom_sum_dev, double *small_dev,
                double *nega_dev, int *anyneg_dev);
void photo_destroy(vk_column *c);
void ens_destroy(vk_column *c);
void conden_destroy(vk_column *c);
void rates_destroy(vk_network *n);

template <typename T>
static int dev_copy(std::vector<void *> &allocs, const T *host, size_t n, const T **out)
{
    void *p = nullptr;
    VK_CUDA(cudaMalloc(&p, sizeof(T) * (n ? n : 1)));
    allocs.push_back(p);
    if (n) VK_CUDA(cudaMemcpy(p, host, sizeof(T) * n, cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const T *>(p);
    return VK_OK;
}

static int pad_block(int ni)
{
    if (ni <= 48) return 48;
    return 24 * ((ni + 23) / 24);
}

// wait for the handle's stream.  Batches: through a blocking-sync event - the calling thread sleeps.  Several ranks x several column-group
// threads per rank outnumber the host cores of a GPU box (8 ranks x 4 groups on 16 threads); spinning waiters then take the cores away from
// the threads that still have launches to issue.  One column: a spinning wait (the wake-up latency of a blocking wait is of the order of a step)
int stream_wait(vk_column *c)
{
    if (c->ncol < 32 || !c->ev_sync) { VK_CUDA(cudaStreamSynchronize(c->stream)); return VK_OK; }
    VK_CUDA(cudaEventRecord(c->ev_sync, c->stream));
    VK_CUDA(cudaEventSynchronize(c->ev_sync));
    return VK_OK;
}

// device-side sequence of one attempted step; y, ymix, dt already resident
int vk_step_device_impl(vk_column *c)
{
    int rc;
    VK_CUDA(cudaMemsetAsync(c->status, 0, sizeof(int) * c->ncol, c->stream));
    VK_CUDA(cudaEventRecord(c->ev0, c->stream));
    if ((rc = launch_rhs(c, c->y, c->f, nullptr, nullptr, nullptr, nullptr))) return rc;          // f(y_n)            op.py:2892
    // I/(r h) - J (op.py:2893) and the block LU factors F_j of the Schur blocks (c->W).  Default: lhs_ml_kernel -> D in HBM ->
    // factor_kernel.  VK_FUSED=1: ONE kernel - the block's otherwise idle warps assemble D_{j+1} in shared memory while the column warps
    // factor layer j; D reaches HBM only for the columns the refinement will read it for.  Measured on the B200:
    // bit-identical blocks, 36.1 ms against 23.1 + 14.4 ms for 4096 columns - the two producer warps are bound by the shared-memory pipe
    // they share with the column warps - and SLOWER for one column, so it is an option, not the default
    static int fused_env = -1;
    if (fused_env < 0) { const char *e = getenv("VK_FUSED"); fused_env = e ? atoi(e) : 0; }
    VK_CUDA(cudaEventRecord(c->ev1, c->stream));
    rc = VK_ERR_UNSUPPORTED;
    if (fused_env && !c->cr_now) rc = launch_factor_fused(c, c->y, c->dt, c->D, c->up, c->dn, c->W, c->status, c->opts.refine > 0 ? 1 : (c->opts.refine < 0 ? 2 : 0));
    c->last_fused = (rc == VK_OK);
    if (rc == VK_ERR_UNSUPPORTED) {
        if ((rc = launch_lhs(c, c->y, c->dt, c->nip, c->D, c->up, c->dn))) return rc;
        VK_CUDA(cudaEventRecord(c->ev1, c->stream));
        rc = launch_factor(c, c->D, c->up, c->dn, c->W, c->status, c->f);       // (+ forward elimination of the k1 solve at small dt)
    } else c->fwd_valid = 0;
    if (rc) return rc;
    VK_CUDA(cudaEventRecord(c->ev2, c->stream));
    if ((rc = launch_solve(c, c->W, c->up, c->dn, c->f, c->k1, c->z, c->act, c->fwd_valid ? c->fwd_done : nullptr))) return rc;           // k1                op.py:2914
    if ((rc = launch_refine(c, c->D, c->up, c->dn, c->W, c->f, c->k1, c->opts.refine, c->dt))) return rc;
    if ((rc = launch_rhs(c, c->y, c->rhs, nullptr, nullptr, c->k1, c->dt))) return rc;            // f(y+k1/r) - 2/(rh) k1   op.py:2917-2928
    if ((rc = launch_solve(c, c->W, c->up, c->dn, c->rhs, c->k2, c->z, c->act))) return rc;         // k2                op.py:2929
    if ((rc = launch_refine(c, c->D, c->up, c->dn, c->W, c->rhs, c->k2, c->opts.refine, c->dt))) return rc;
    if ((rc = launch_epilogue(c))) return rc;                                                       // sol, delta, ymix  op.py:2932-2993
    VK_CUDA(cudaEventRecord(c->ev3, c->stream));
    return VK_OK;
}


}  // namespace vk

using namespace vk;

extern "C" {

int vk_abi_version(void) { return VK_ABI_VERSION; }
const char *vk_last_error(void) { return g_err.c_str(); }
int vk_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDeviceCount"); return VK_ERR_CUDA; }
    return n;
}

// ---------------------------------------------------------------------------------------------------------------
int vk_network_create(const vk_network_desc *d, int device, vk_network **out)
{
    if (!d || !out) { set_error("null argument"); return VK_ERR_INVALID; }
    if (d->ni < 1 || d->ni > 253 || d->nr < 2 || d->nr > 65534) { set_error("network size out of range (ni <= 253, nr <= 65534)"); return VK_ERR_UNSUPPORTED; }
    if (d->maxf > 4 || d->maxjf > 3) { set_error("more than 4 factors per rate term / 3 per Jacobian term"); return VK_ERR_UNSUPPORTED; }
    if (pad_block(d->ni) > 120) { set_error("ni > 120: no factor kernel instantiated"); return VK_ERR_UNSUPPORTED; }
    VK_CUDA(cudaSetDevice(device));
    vk_network *n = new vk_network();
    n->device = device;
    n->rates = nullptr;
    n->table_hash = network_table_hash(d);
    n->emit = emit_lookup(n->table_hash, d->ni, d->nr);
    if (getenv("VK_DEBUG")) fprintf(stderr, "vulcan_b200: network ni %d nr %d table hash %016llx: %s\n", d->ni, d->nr, n->table_hash,
                                    n->emit ? "emitted chemdf kernel" : "table-driven chemdf");
    const int ni = d->ni, nr = d->nr;
    std::vector<uchar4> rf(nr + 1), rp(nr + 1);
    int has_pow = 0;
    for (int i = 0; i <= nr; i++) {
        unsigned char f[4], p[4];
        for (int q = 0; q < 4; q++) {
            int slot = (q < d->maxf) ? d->rate_fac[i * d->maxf + q] : ni + 1;
            int pw = (q < d->maxf) ? d->rate_pow[i * d->maxf + q] : 1;
            if (i == 0) { slot = ni + 1; pw = 1; }
            if (slot < 0 || slot > ni + 1 || pw < 1 || pw > 255) { delete n; set_error("bad rate table entry"); return VK_ERR_INVALID; }
            if (pw != 1) has_pow = 1;
            f[q] = (unsigned char)slot; p[q] = (unsigned char)pw;
        }
        rf[i] = make_uchar4(f[0], f[1], f[2], f[3]);
        rp[i] = make_uchar4(p[0], p[1], p[2], p[3]);
    }
    std::vector<int> rt(d->n_rhs);
    int max_len = 0;
    for (int s = 0; s < ni; s++) max_len = std::max(max_len, d->rhs_ptr[s + 1] - d->rhs_ptr[s]);
    for (int q = 0; q < d->n_rhs; q++) {
        double c = d->rhs_coef[q];
        int ci = (int)c;
        if ((double)ci != c || ci < -127 || ci > 127) { delete n; set_error("non-integer stoichiometric coefficient"); return VK_ERR_UNSUPPORTED; }
        rt[q] = (d->rhs_pair[q] << 8) | (ci & 0xff);
    }
    std::vector<ushort2> rc(d->n_ent);
    for (int e = 0; e < d->n_ent; e++) rc[e] = make_ushort2((unsigned short)d->jac_row[e], (unsigned short)d->jac_col[e]);
    std::vector<uint2> jt(d->n_term);
    for (int q = 0; q < d->n_term; q++) {
        double c = d->jac_coef[q];
        int ci = (int)c;
        if ((double)ci != c || ci < -127 || ci > 127) { delete n; set_error("Jacobian coefficient out of range"); return VK_ERR_UNSUPPORTED; }
        unsigned f[3];
        for (int x = 0; x < 3; x++) f[x] = (x < d->maxjf) ? (unsigned)d->jac_fac[q * d->maxjf + x] : (unsigned)(ni + 1);
        jt[q] = make_uint2((unsigned)d->jac_k[q] | ((unsigned)(ci & 0xff) << 16), f[0] | (f[1] << 8) | (f[2] << 16));
    }
    // warp-per-layer rhs kernel tables
    std::vector<unsigned short> rd16(d->n_rhs);
    int rhs_unit = (nr / 2 < 32768) ? 1 : 0;
    for (int q = 0; q < d->n_rhs; q++) {
        const int ci = (int)d->rhs_coef[q];
        if (ci != 1 && ci != -1) rhs_unit = 0;
        rd16[q] = (unsigned short)((((d->rhs_pair[q] - 1) / 2) & 0x7fff) | (ci < 0 ? 0x8000 : 0));
    }
    std::vector<int> lane_sp(32 * VK_RHS_SPL, -1);
    {
        std::vector<int> order(ni), load(32, 0), cnt(32, 0);
        for (int s = 0; s < ni; s++) order[s] = s;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return d->rhs_ptr[a + 1] - d->rhs_ptr[a] > d->rhs_ptr[b + 1] - d->rhs_ptr[b]; });
        for (int s : order) {
            int best = -1;
            for (int l = 0; l < 32; l++)
                if (cnt[l] < VK_RHS_SPL && (best < 0 || load[l] < load[best])) best = l;
            if (best < 0) { delete n; set_error("too many species for the rhs lane schedule"); return VK_ERR_UNSUPPORTED; }
            lane_sp[best * VK_RHS_SPL + cnt[best]++] = s;
            load[best] += d->rhs_ptr[s + 1] - d->rhs_ptr[s] + 4;     // + a little per-species overhead
        }
    }
    // segmented rhs summation: flat term list in species order, 32 equal chunks
    std::vector<unsigned short> flat16;
    std::vector<int> seg_ptr(ni + 1, 0), lane_slot0(32, 0);
    int flat_ok = (rhs_unit && nr / 2 < 16383) ? 1 : 0, flat_T = 0, n_segs = 0;
    if (flat_ok) {
        const int nterm = d->n_rhs;
        flat_T = std::max(1, (nterm + 31) / 32);
        flat16.assign((size_t)32 * flat_T, (unsigned short)(nr / 2));          // padding: the always-zero pair slot v[npair], no flush
        std::vector<int> sp_of(nterm);
        for (int s = 0; s < ni; s++) for (int q = d->rhs_ptr[s]; q < d->rhs_ptr[s + 1]; q++) sp_of[q] = s;
        int seg = 0, cur_sp = -1;
        std::vector<int> first_seg(ni, -1), last_seg(ni, -1);
        for (int f = 0; f < nterm; f++) {
            const int l = f / flat_T, q = f % flat_T;
            if (q == 0) lane_slot0[l] = seg;
            const int s = sp_of[f];
            if (first_seg[s] < 0) first_seg[s] = seg;
            last_seg[s] = seg;
            const bool flush = (f + 1 == nterm) || (sp_of[f + 1] != s) || (q == flat_T - 1);
            const int ci = (int)d->rhs_coef[f];
            flat16[(size_t)32 * q + l] = (unsigned short)((((d->rhs_pair[f] - 1) / 2) & 0x3fff) | (flush ? 0x4000 : 0) | (ci < 0 ? 0x8000 : 0));
            if (flush) seg++;
            (void)cur_sp;
        }
        for (int l = (nterm + flat_T - 1) / flat_T; l < 32; l++) lane_slot0[l] = seg;
        n_segs = seg;
        int run = 0;
        for (int s = 0; s < ni; s++) {           // species without terms own no partial
            seg_ptr[s] = (first_seg[s] >= 0) ? first_seg[s] : run;
            if (last_seg[s] >= 0) run = last_seg[s] + 1;
        }
        seg_ptr[ni] = n_segs;
        for (int s = ni - 1; s >= 0; s--) if (first_seg[s] < 0) seg_ptr[s] = seg_ptr[s + 1];
    }
    if (flat16.empty()) flat16.push_back(0);
    // work schedule of the Jacobian kernel: segments of <= 16 terms sorted by decreasing length
    struct Seg { unsigned rc; int q0; int len; int ent; };
    std::vector<Seg> segs;
    std::vector<uint2> multi;
    std::vector<int> ent_slot0(d->n_ent, -1);
    int n_part = 0;
    const int SEGLEN = 16;
    for (int e = 0; e < d->n_ent; e++) {
        const int q0 = d->jac_ptr[e], q1 = d->jac_ptr[e + 1];
        const unsigned rcw = (unsigned)d->jac_row[e] | ((unsigned)d->jac_col[e] << 16);
        const int nseg = std::max(1, (q1 - q0 + SEGLEN - 1) / SEGLEN);
        if (nseg > 1) {
            ent_slot0[e] = n_part;
            multi.push_back(make_uint2(rcw, (unsigned)n_part | ((unsigned)nseg << 16)));
            n_part += nseg;
        }
        for (int sI = 0; sI < nseg; sI++) {
            const int a0 = q0 + sI * SEGLEN, a1 = std::min(q1, a0 + SEGLEN);
            segs.push_back(Seg{rcw, a0, a1 - a0, (nseg > 1) ? (ent_slot0[e] + sI) : 0xffff});
        }
    }
    if (n_part >= 0xffff) { delete n; set_error("too many split Jacobian entries"); return VK_ERR_UNSUPPORTED; }
    std::stable_sort(segs.begin(), segs.end(), [](const Seg &a, const Seg &b) { return a.len > b.len; });
    std::vector<uint4> segw(segs.size());
    for (size_t q = 0; q < segs.size(); q++)
        segw[q] = make_uint4(segs[q].rc, (unsigned)segs[q].q0, (unsigned)segs[q].len | ((unsigned)segs[q].ent << 16), 0u);
    // multi-layer Jacobian kernel tables: distinct products + 16-bit term descriptors (coefficient codes, see lhs_ml_kernel)
    std::vector<unsigned> uq;
    std::vector<unsigned short> jt16(d->n_term), ttT;
    std::vector<unsigned> seg4;
    std::vector<uint2> grp;
    int ml_ok = (nr < 2048 && ni + 1 < 128 && d->n_term < 65536 && d->maxjf <= 3) ? 1 : 0;
    if (ml_ok) {
        std::vector<std::pair<unsigned, int>> keyed(d->n_term);
        for (int q = 0; q < d->n_term; q++) {
            unsigned f[3];
            for (int x = 0; x < 3; x++) f[x] = (x < d->maxjf) ? (unsigned)d->jac_fac[q * d->maxjf + x] : (unsigned)(ni + 1);
            keyed[q] = std::make_pair((unsigned)d->jac_k[q] | (f[0] << 11) | (f[1] << 18) | (f[2] << 25), q);
        }
        std::vector<std::pair<unsigned, int>> sorted = keyed;
        std::sort(sorted.begin(), sorted.end());
        std::vector<int> uidx(d->n_term);
        for (size_t q = 0; q < sorted.size(); q++) {
            if (q == 0 || sorted[q].first != sorted[q - 1].first) uq.push_back(sorted[q].first);
            uidx[sorted[q].second] = (int)uq.size() - 1;
        }
        if (uq.size() >= 8191) ml_ok = 0;
        static const int codes[8] = {1, -1, 2, -2, 4, -4, 3, -3};
        for (int q = 0; q < d->n_term && ml_ok; q++) {
            const int ci = (int)d->jac_coef[q];
            int code = -1;
            for (int x = 0; x < 8; x++) if (codes[x] == ci) code = x;
            if (code < 0) { ml_ok = 0; break; }
            jt16[q] = (unsigned short)(uidx[q] | (code << 13));
        }
        // groups of 32 segments (already sorted by decreasing length), terms transposed, padded with the zero product
        const unsigned short pad = (unsigned short)uq.size();          // dprod[n_uniq] = 0.0, code 0 (x 1.0)
        for (size_t g0 = 0; g0 < segs.size() && ml_ok; g0 += 32) {
            const int nmax = segs[g0].len;
            grp.push_back(make_uint2((unsigned)ttT.size(), (unsigned)nmax));
            const size_t base = ttT.size();
            ttT.resize(base + (size_t)32 * nmax, pad);
            for (int l = 0; l < 32; l++) {
                if (g0 + l < segs.size()) {
                    const Seg &sg = segs[g0 + l];
                    for (int q = 0; q < sg.len; q++) ttT[base + (size_t)32 * q + l] = jt16[sg.q0 + q];
                    seg4.push_back((sg.rc & 0xff) | (((sg.rc >> 16) & 0xff) << 8) | ((unsigned)sg.ent << 16));
                } else {
                    seg4.push_back(0xffu | (0xffu << 8) | (0xffffu << 16));
                }
            }
        }
    }
    if (grp.empty()) { grp.push_back(make_uint2(0, 0)); seg4.resize(32, 0xffffffffu); ttT.push_back(0); }
    if (uq.empty()) uq.push_back(0);
    NetDev &nd = n->d;
    nd.n_uniq = ml_ok ? (int)uq.size() : 0; nd.lhs_ml_ok = ml_ok;
    nd.n_grp = (int)grp.size(); nd.n_tt = (int)ttT.size();
    nd.ni = ni; nd.nr = nr; nd.nip = pad_block(ni);
    nd.n_seg = (int)segw.size(); nd.n_multi = (int)multi.size(); nd.n_part = n_part;
    nd.n_ent = d->n_ent; nd.n_term = d->n_term; nd.n_rhs = d->n_rhs; nd.max_rhs_len = max_len; nd.has_pow = has_pow;
    int rcode = VK_OK;
#define CP(vec, field) if (rcode == VK_OK) rcode = dev_copy(n->allocs, vec.data(), vec.size(), &nd.field)
    CP(rf, rate_fac); CP(rp, rate_pow); CP(rt, rhs_term); CP(rc, jac_rc); CP(jt, jac_term); CP(segw, jac_seg); CP(multi, jac_multi);
    CP(rd16, rhs_desc16); CP(lane_sp, rhs_lane_sp);
    CP(flat16, rhs_flat16); CP(seg_ptr, rhs_seg_ptr); CP(lane_slot0, rhs_lane_slot0);
    nd.rhs_flat_ok = flat_ok; nd.rhs_flat_T = flat_T; nd.rhs_n_seg = n_segs;
    CP(uq, jac_uniq); CP(ttT, jac_tt); CP(seg4, jac_seg4); CP(grp, jac_grp);
    nd.rhs_unit = rhs_unit;
#undef CP
    if (rcode == VK_OK) rcode = dev_copy(n->allocs, d->rhs_ptr, (size_t)ni + 1, &nd.rhs_ptr);
    if (rcode == VK_OK) rcode = dev_copy(n->allocs, d->jac_ptr, (size_t)d->n_ent + 1, &nd.jac_ptr);
    if (rcode != VK_OK) { vk_network_destroy(n); return rcode; }
    VK_CUDA(cudaDeviceSynchronize());      // pageable cudaMemcpy may return before its DMA has landed; the handles' streams are non-blocking
    *out = n;
    return VK_OK;
}

void vk_network_destroy(vk_network *n)
{
    if (!n) return;
    rates_destroy(n);
    for (void *p : n->allocs) cudaFree(p);
    delete n;
}

// ---------------------------------------------------------------------------------------------------------------
static void free_list(std::vector<void *> &v)
{
    for (void *p : v) cudaFree(p);
    v.clear();
}

int vk_column_create(vk_network *net, int nz, int ncol, vk_column **out)
{
    if (!net || !out || nz < 3 || ncol < 1) { set_error("bad argument"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(net->device));
    vk_column *c = new vk_column();
    memset(static_cast<void *>(c), 0, sizeof(vk_column));
    new (&c->atm_allocs) std::vector<void *>();
    new (&c->opt_allocs) std::vector<void *>();
    c->net = net; c->device = net->device; c->nz = nz; c->ncol = ncol; c->ni = net->d.ni; c->nr = net->d.nr; c->nip = net->d.nip;
    const size_t nv = (size_t)ncol * nz * c->ni, nb = (size_t)ncol * nz * c->nip * c->nip, np = (size_t)ncol * nz * c->nip;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev2);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev3);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming);
    double **vecs[] = {&c->y, &c->ymix, &c->sol, &c->ymix_out, &c->f, &c->k1, &c->k2, &c->yk2, &c->rhs, &c->res, &c->dx, &c->xn};
    for (double **v : vecs)
        if (e == cudaSuccess) e = cudaMalloc((void **)v, sizeof(double) * nv);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->z, sizeof(double) * np);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->up, sizeof(double) * np);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->dn, sizeof(double) * np);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->D, sizeof(double) * nb);
    // block LU factors of the Schur blocks, row stride nip + 2 (vk_solve.cu)
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->W, sizeof(double) * (size_t)ncol * nz * c->nip * (c->nip + 2));
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->dt, sizeof(double) * ncol);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->delta, sizeof(double) * ncol);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->status, sizeof(int) * ncol);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->refine_kept, sizeof(int) * 4 * ncol);
    if (e == cudaSuccess) e = cudaMemset(c->refine_kept, 0, sizeof(int) * 4 * ncol);
    if (e == cudaSuccess) { c->refine_tried = c->refine_kept + ncol; c->refine_act = c->refine_kept + 2 * ncol; c->fwd_done = c->refine_kept + 3 * ncol; }
    c->h_pin_bytes = sizeof(double) * (4 * nv + 4 * (size_t)ncol) + 64;
    if (e == cudaSuccess) e = cudaMallocHost((void **)&c->h_pin, c->h_pin_bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();      // the memset above is on the legacy stream, the handle's stream is non-blocking
    if (e != cudaSuccess) { cuda_fail(e, "vk_column_create allocation"); vk_column_destroy(c); return VK_ERR_CUDA; }
    c->opts.mtol = 0; c->opts.atol = 0;
    {   // one column: the latency path (block cyclic reduction over the layers, vk_cr.inl) unless VK_CR=0
        const char *e = getenv("VK_CR");
        c->use_cr = (ncol == 1) && !(e && atoi(e) == 0);
        const char *m = getenv("VK_CR_DT_MAX");
        c->cr_dt_max = m ? atof(m) : 1.0e5;
        c->cr_now = c->use_cr;
        c->dt_host_max = -1.0;
    }
    *out = c;
    return VK_OK;
}

void vk_column_destroy(vk_column *c)
{
    if (!c) return;
    if (cudaSetDevice(c->device) != cudaSuccess) cudaGetLastError();
    if (c->stream) cudaStreamSynchronize(c->stream);
    photo_destroy(c);
    ens_destroy(c);
    conden_destroy(c);
    cr_plan_free(c->cr);
    if (c->refine_kept) cudaFree(c->refine_kept);
    double *vecs[] = {c->y, c->ymix, c->sol, c->ymix_out, c->f, c->k1, c->k2, c->yk2, c->rhs, c->res, c->dx, c->xn, c->z, c->up, c->dn,
                      c->D, c->W, c->dt, c->delta, c->k, c->chem_tmp, c->ysum_tmp, static_cast<double *>(c->scal_tmp), c->ysum_lhs_tmp};
    for (double *v : vecs) if (v) cudaFree(v);
    if (c->status) cudaFree(c->status);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->clip_d) cudaFree(c->clip_d);
    if (c->clip_i) cudaFree(c->clip_i);
    if (c->clip_h) cudaFreeHost(c->clip_h);
    free_list(c->atm_allocs);
    free_list(c->opt_allocs);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev3) cudaEventDestroy(c->ev3);
    if (c->ev_sync) cudaEventDestroy(c->ev_sync);
    if (c->stream) cudaStreamDestroy(c->stream);
    c->atm_allocs.~vector();
    c->opt_allocs.~vector();
    ::operator delete(c);
}

int vk_set_atm(vk_column *c, const vk_atm_view *v)
{
    if (!c || !v) { set_error("null argument"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    free_list(c->atm_allocs);
    const int nz = c->nz, ni = c->ni;
    const size_t rep = v->shared ? 1 : (size_t)c->ncol;
    AtmDev &a = c->atm;
    a.nz = nz; a.ni = ni;
    a.use_moldiff = v->use_moldiff; a.use_settling = v->use_settling; a.use_topflux = v->use_topflux; a.use_botflux = v->use_botflux;
    a.n_gas = (v->n_gas == ni) ? 0 : v->n_gas;
    a.n_gas_lhs = (v->n_gas_lhs == ni) ? 0 : v->n_gas_lhs;
    a.cs1 = v->shared ? 0 : (size_t)(nz - 1);
    a.csn = v->shared ? 0 : (size_t)(nz - 1) * ni;
    a.csz = v->shared ? 0 : (size_t)nz;
    a.csi = v->shared ? 0 : (size_t)ni;
    int rc = VK_OK;
    a.gas_indx = nullptr; a.gas_indx_lhs = nullptr;
    if (a.n_gas > 0) rc = dev_copy(c->atm_allocs, v->gas_indx, (size_t)a.n_gas, &a.gas_indx);
    if (rc == VK_OK && a.n_gas_lhs > 0) rc = dev_copy(c->atm_allocs, v->gas_indx_lhs, (size_t)a.n_gas_lhs, &a.gas_indx_lhs);
    std::vector<double> zeros((size_t)rep * (size_t)(nz) * ni, 0.0);
#define CPA(field, n) if (rc == VK_OK) rc = dev_copy(c->atm_allocs, v->field ? v->field : zeros.data(), rep * (size_t)(n), &a.field)
    CPA(Kzz, nz - 1); CPA(vz, nz - 1); CPA(dzi, nz - 1); CPA(Ti, nz - 1); CPA(Hpi, nz - 1);
    CPA(Dzz, (size_t)(nz - 1) * ni); CPA(vs, (size_t)(nz - 1) * ni);
    CPA(Tco, nz); CPA(g, nz); CPA(M, nz);
    CPA(ms, ni); CPA(alpha, ni); CPA(top_flux, ni); CPA(bot_flux, ni); CPA(bot_vdep, ni);
    a.use_vm_mol = (v->use_vm_mol && v->use_moldiff) ? 1 : 0;      // Ros2.solver's dispatch (op.py:2879-2888)
    a.vm = nullptr; a.csv = 0; a.n_diff_esc = 0; a.diff_esc_idx = nullptr;
    if (a.use_vm_mol) {
        if (!v->vm && rc == VK_OK) { set_error("use_vm_mol needs atm.vm"); rc = VK_ERR_INVALID; }
        CPA(vm, (size_t)nz * ni);
        a.csv = v->shared ? 0 : (size_t)nz * ni;
        if (rc == VK_OK && v->n_diff_esc > 0) {
            for (int q = 0; q < v->n_diff_esc; q++)
                if (v->diff_esc_idx[q] < 0 || v->diff_esc_idx[q] >= ni) { set_error("diff_esc species out of range"); rc = VK_ERR_INVALID; }
            a.n_diff_esc = v->n_diff_esc;
            if (rc == VK_OK) rc = dev_copy(c->atm_allocs, v->diff_esc_idx, (size_t)a.n_diff_esc, &a.diff_esc_idx);
        }
    }
#undef CPA
    if (rc != VK_OK) return rc;
    // atmosphere-only stencil pieces (evaluated once per vk_set_atm)
    {
        const size_t npre = rep * (size_t)nz * ni;
        double **pp[] = {&a.pre.Q, &a.pre.QB, &a.pre.QC, &a.pre.TA, &a.pre.TB, &a.pre.TC, &a.pre.SA, &a.pre.SB, &a.pre.SC};
        for (double **q : pp) {
            void *ptr = nullptr;
            VK_CUDA(cudaMalloc(&ptr, sizeof(double) * npre));
            c->atm_allocs.push_back(ptr);
            *q = reinterpret_cast<double *>(ptr);
        }
        {
            void *ptr = nullptr;
            VK_CUDA(cudaMalloc(&ptr, sizeof(double) * rep * (size_t)nz * 10));
            c->atm_allocs.push_back(ptr);
            a.pre.LS = reinterpret_cast<double *>(ptr);
        }
        a.pre_cs = v->shared ? 0 : (size_t)nz * ni;
        VK_CUDA(cudaDeviceSynchronize());  // the copies above ran on the legacy stream (pageable sources): land them before the handle's stream reads
        if ((rc = launch_atm_pre(c, (int)rep))) return rc;
        VK_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->atm_set = true;
    return VK_OK;
}

int vk_set_k(vk_column *c, const double *k, int shared)
{
    if (!c || !k) { set_error("null argument"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    const size_t per = (size_t)c->nz * (c->nr + 1);
    const bool one = (shared == 1);
    const size_t want_cs = one ? 0 : per;
    if (!c->k || c->k_cs != want_cs || !c->k_set) {
        VK_CUDA(cudaStreamSynchronize(c->stream));
        if (c->k) cudaFree(c->k);
        c->k = nullptr;
        VK_CUDA(cudaMalloc((void **)&c->k, sizeof(double) * per * (one ? 1 : c->ncol)));
        c->k_cs = want_cs;
    }
    VK_CUDA(cudaMemcpyAsync(c->k, k, sizeof(double) * per * (one ? 1 : c->ncol), cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    c->k_set = true;
    c->k_static_shared = (shared == 2);
    return VK_OK;
}

__global__ void scatter_k_rows(double *k, size_t k_cs, int nz, int nr, int ncol_k, int n_rows, const int *rows, const double *vals)
{
    // vals [ncol_k][n_rows][nz]
    size_t n = (size_t)ncol_k * n_rows * nz;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        int j = (int)(q % nz);
        int r = (int)((q / nz) % n_rows);
        size_t col = q / ((size_t)nz * n_rows);
        k[col * k_cs + (size_t)j * (nr + 1) + rows[r]] = vals[q];
    }
}

int vk_set_k_rows(vk_column *c, int n_rows, const int *rows, const double *vals)
{
    if (!c || !c->k_set || n_rows < 0 || (n_rows && (!rows || !vals))) { set_error("bad argument / k not set"); return VK_ERR_INVALID; }
    if (n_rows == 0) return VK_OK;
    VK_CUDA(cudaSetDevice(c->net->device));
    for (int r = 0; r < n_rows; r++) {
        if (rows[r] < 1 || rows[r] > c->nr) { set_error("reaction id out of range"); return VK_ERR_INVALID; }
        // per-column values for a row the emitted kernels read from the block's shared copy: the promise of vk_set_k(shared = 2) is gone
        if (c->k_cs && c->k_static_shared && !emit_row_is_dynamic(c->net->emit, rows[r])) c->k_static_shared = false;
    }
    const int ncol_k = c->k_cs ? c->ncol : 1;
    const size_t n = (size_t)ncol_k * n_rows * c->nz;
    int *drows = nullptr; double *dvals = nullptr;
    VK_CUDA(cudaMalloc((void **)&drows, sizeof(int) * n_rows));
    cudaError_t e = cudaMalloc((void **)&dvals, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(drows, rows, sizeof(int) * n_rows, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dvals, vals, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        scatter_k_rows<<<(int)std::min<size_t>((n + 255) / 256, 1184), 256, 0, c->stream>>>(c->k, c->k_cs, c->nz, c->nr, ncol_k, n_rows, drows, dvals);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(drows); if (dvals) cudaFree(dvals);
    if (e != cudaSuccess) return cuda_fail(e, "vk_set_k_rows");
    return VK_OK;
}

int vk_set_step_opts(vk_column *c, const vk_step_opts *o)
{
    if (!c || !o) { set_error("null argument"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    free_list(c->opt_allocs);
    StepOptsDev &d = c->opts;
    memset(&d, 0, sizeof(d));
    d.mtol = o->mtol; d.atol = o->atol; d.refine = o->refine; d.zero_delta_row0 = o->zero_delta_row0; d.n_fix_bot = o->n_fix_bot;
    const size_t nv = (size_t)c->ncol * c->nz * c->ni;
    int rc = VK_OK;
    if (o->n_fix_bot > 0) {
        rc = dev_copy(c->opt_allocs, o->fix_bot_idx, (size_t)o->n_fix_bot, &d.fix_bot_idx);
        if (rc == VK_OK) rc = dev_copy(c->opt_allocs, o->fix_bot_val, (size_t)c->ncol * o->n_fix_bot, &d.fix_bot_val);
    }
    if (rc == VK_OK && o->delta_zero_sp) rc = dev_copy(c->opt_allocs, o->delta_zero_sp, (size_t)c->ni, &d.delta_zero_sp);
    if (rc == VK_OK && o->fix_mask) rc = dev_copy(c->opt_allocs, o->fix_mask, nv, &d.fix_mask);
    if (rc == VK_OK && o->fix_y) rc = dev_copy(c->opt_allocs, o->fix_y, nv, &d.fix_y);
    d.refine_dt_min = o->refine_dt_min;
    d.rhs_order = o->rhs_order;
    if (rc == VK_OK && o->compo && o->na > 0) {
        if (o->na > 8) { set_error("at most 8 elements in atom_list are supported"); return VK_ERR_UNSUPPORTED; }
        d.na = o->na;
        rc = dev_copy(c->opt_allocs, o->compo, (size_t)c->ni * o->na, &d.compo);
    }
    if (rc == VK_OK && o->refine < 0 && !d.compo) { set_error("refine = auto (-1) needs compo / na"); return VK_ERR_INVALID; }
    if (rc == VK_OK) VK_CUDA(cudaDeviceSynchronize());
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
static int ready(vk_column *c)
{
    if (!c) { set_error("null handle"); return VK_ERR_INVALID; }
    if (!c->atm_set || !c->k_set) { set_error("vk_set_atm / vk_set_k must be called first"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    return VK_OK;
}

int vk_ros2_solve(vk_column *c, const double *y, const double *ymix, const double *dt, double *sol, double *ymix_out,
                  double *delta, int *status)
{
    int rc = ready(c);
    if (rc) return rc;
    if (!y || !ymix || !dt || !sol || !ymix_out || !delta) { set_error("null buffer"); return VK_ERR_INVALID; }
    const size_t nv = (size_t)c->ncol * c->nz * c->ni;
    // caller buffers that are already page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) are used directly;
    // pageable buffers are staged through the handle's pinned area so that the copies are real asynchronous DMA
    auto pinned = [](const void *p) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    const bool direct = pinned(y) && pinned(ymix) && pinned(sol) && pinned(ymix_out);
    double *hy = c->h_pin, *hm = hy + nv, *hs = hm + nv, *ho = hs + nv, *hdt = ho + nv, *hdl = hdt + c->ncol;
    int *hst = reinterpret_cast<int *>(hdl + c->ncol);
    if (!direct) {
        memcpy(hy, y, sizeof(double) * nv);
        memcpy(hm, ymix, sizeof(double) * nv);
    }
    memcpy(hdt, dt, sizeof(double) * c->ncol);
    c->cr_now = c->use_cr && dt[0] <= c->cr_dt_max;
    c->dt_host_max = dt[0];
    for (int q = 1; q < c->ncol; q++) c->dt_host_max = std::max(c->dt_host_max, dt[q]);
    VK_CUDA(cudaMemcpyAsync(c->y, direct ? y : hy, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(c->ymix, direct ? ymix : hm, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(c->dt, hdt, sizeof(double) * c->ncol, cudaMemcpyHostToDevice, c->stream));
    if ((rc = vk_step_device_impl(c))) return rc;
    VK_CUDA(cudaMemcpyAsync(direct ? sol : hs, c->sol, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaMemcpyAsync(direct ? ymix_out : ho, c->ymix_out, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaMemcpyAsync(hdl, c->delta, sizeof(double) * c->ncol, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaMemcpyAsync(hst, c->status, sizeof(int) * c->ncol, cudaMemcpyDeviceToHost, c->stream));
    { int rcw = stream_wait(c); if (rcw) return rcw; }
    if (!direct) {
        memcpy(sol, hs, sizeof(double) * nv);
        memcpy(ymix_out, ho, sizeof(double) * nv);
    }
    memcpy(delta, hdl, sizeof(double) * c->ncol);
    if (status) memcpy(status, hst, sizeof(int) * c->ncol);
    cudaEventElapsedTime(&c->last_ms_total, c->ev0, c->ev3);
    cudaEventElapsedTime(&c->last_ms_factor, c->ev1, c->ev2);
    return VK_OK;
}

int vk_clip_loss(vk_column *c, double *y, const double *ymix_in, double *ymix_out, int na, const double *compo,
                 const unsigned char *atom_skip, double pos_cut, double nega_cut, double mtol, double *atom_sum, double *small_y,
                 double *nega_y, int *any_negative)
{
    int rc = ready(c);
    if (rc) return rc;
    if (!y || !ymix_in || !ymix_out || !compo || !atom_sum || !small_y || !nega_y || !any_negative || na < 1 || na > 8) { set_error("null buffer / na out of range"); return VK_ERR_INVALID; }
    const size_t nv = (size_t)c->ncol * c->nz * c->ni;
    // persistent scratch (one allocation per handle, not per call): compo [ni][8] | atom_sum [ncol][8] | small, nega [ncol] | skip [8] | anyneg [ncol]
    const size_t n_d = (size_t)c->ni * 8 + (size_t)c->ncol * 8 + 2 * (size_t)c->ncol;
    if (!c->clip_d) {
        VK_CUDA(cudaMalloc((void **)&c->clip_d, sizeof(double) * n_d));
        cudaError_t e = cudaMalloc((void **)&c->clip_i, sizeof(int) * ((size_t)c->ncol + 2));
        if (e == cudaSuccess) e = cudaMallocHost((void **)&c->clip_h, sizeof(double) * n_d + sizeof(int) * ((size_t)c->ncol + 2));
        if (e != cudaSuccess) { cudaFree(c->clip_d); c->clip_d = nullptr; if (c->clip_i) cudaFree(c->clip_i); c->clip_i = nullptr; return cuda_fail(e, "vk_clip_loss scratch"); }
    }
    double *d_compo = c->clip_d, *d_asum = d_compo + (size_t)c->ni * 8, *d_small = d_asum + (size_t)c->ncol * 8, *d_nega = d_small + c->ncol;
    int *d_neg = c->clip_i + 2;
    unsigned char *d_skip = reinterpret_cast<unsigned char *>(c->clip_i);
    double *h = c->clip_h, *h_compo = h, *h_asum = h_compo + (size_t)c->ni * 8, *h_small = h_asum + (size_t)c->ncol * 8, *h_nega = h_small + c->ncol;
    int *h_i = reinterpret_cast<int *>(h + n_d);
    memcpy(h_compo, compo, sizeof(double) * c->ni * na);
    memcpy(h_asum, atom_sum, sizeof(double) * c->ncol * na);
    memcpy(h_small, small_y, sizeof(double) * c->ncol);
    memcpy(h_nega, nega_y, sizeof(double) * c->ncol);
    memset(h_i, 0, 8);
    if (atom_skip) memcpy(h_i, atom_skip, (size_t)na);
    VK_CUDA(cudaMemcpyAsync(c->clip_d, h, sizeof(double) * n_d, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(c->clip_i, h_i, 8, cudaMemcpyHostToDevice, c->stream));
    // sol / ymix buffers double as staging for the clip
    VK_CUDA(cudaMemcpyAsync(c->sol, y, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(c->ymix, ymix_in, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    const double keep_mtol = c->opts.mtol;
    if (mtol >= 0) c->opts.mtol = mtol;              // op.py:2459 reads vulcan_cfg.mtol; negative: the value of vk_set_step_opts
    rc = launch_clip(c, c->sol, c->ymix, c->ymix_out, na, d_compo, atom_skip ? d_skip : nullptr, pos_cut, nega_cut, d_asum, d_small, d_nega, d_neg);
    c->opts.mtol = keep_mtol;
    if (rc) return rc;
    VK_CUDA(cudaMemcpyAsync(y, c->sol, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaMemcpyAsync(ymix_out, c->ymix_out, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaMemcpyAsync(h_asum, d_asum, sizeof(double) * ((size_t)c->ncol * 8 + 2 * (size_t)c->ncol), cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaMemcpyAsync(h_i + 2, d_neg, sizeof(int) * c->ncol, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(atom_sum, h_asum, sizeof(double) * c->ncol * na);
    memcpy(small_y, h_small, sizeof(double) * c->ncol);
    memcpy(nega_y, h_nega, sizeof(double) * c->ncol);
    memcpy(any_negative, h_i + 2, sizeof(int) * c->ncol);
    return VK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
int vk_eval_rhs(vk_column *c, const double *y, double *out_chem, double *out_diff)
{
    int rc = ready(c);
    if (rc) return rc;
    if (!y) { set_error("null buffer"); return VK_ERR_INVALID; }
    const size_t nv = (size_t)c->ncol * c->nz * c->ni;
    VK_CUDA(cudaMemcpyAsync(c->y, y, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    if ((rc = launch_rhs(c, c->y, nullptr, out_chem ? c->k1 : nullptr, out_diff ? c->k2 : nullptr, nullptr, nullptr))) return rc;
    if (out_chem) VK_CUDA(cudaMemcpyAsync(out_chem, c->k1, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
    if (out_diff) VK_CUDA(cudaMemcpyAsync(out_diff, c->k2, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    return VK_OK;
}

__global__ void unpad_system(int ni, int nip, const double *D, const double *up, const double *dn, double *Dd, double *upd, double *dnd)
{
    const size_t blk = blockIdx.x;
    for (int q = threadIdx.x; q < ni * ni; q += blockDim.x) Dd[blk * ni * ni + q] = D[blk * nip * nip + (size_t)(q / ni) * nip + q % ni];
    for (int i = threadIdx.x; i < ni; i += blockDim.x) { upd[blk * ni + i] = up[blk * nip + i]; dnd[blk * ni + i] = dn[blk * nip + i]; }
}

int vk_eval_lhs(vk_column *c, const double *y, const double *dt, double *D, double *up, double *dn)
{
    int rc = ready(c);
    if (rc) return rc;
    if (!y || !dt || !D || !up || !dn) { set_error("null buffer"); return VK_ERR_INVALID; }
    const size_t nv = (size_t)c->ncol * c->nz * c->ni;
    VK_CUDA(cudaMemcpyAsync(c->y, y, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream));
    VK_CUDA(cudaMemcpyAsync(c->dt, dt, sizeof(double) * c->ncol, cudaMemcpyHostToDevice, c->stream));
    const char *via = getenv("VK_LHS_VIA_FUSED");
    const char *ej = getenv("VK_EMIT_JAC");
    // batches that the step assembles through the EMITTED Jacobian kernel (vk_chem.cu: launch_lhs) are evaluated the same way here: padded
    // blocks, un-padded on the way out
    const bool emitted = !(via && atoi(via)) && !(ej && !atoi(ej)) && emit_has_jac(c->net->emit) && (c->k_cs == 0 || c->k_static_shared) && c->ncol >= 32;
    if ((via && atoi(via)) || emitted) {
        // parity aid: the blocks as the FUSED assembly + factorisation kernel forms them (producer warps, store_D = always), un-padded
        // on the way out - must be bit-identical to lhs_ml_kernel's (tests/test_gpu_fused.py)
        if (emitted) {
            rc = launch_lhs(c, c->y, c->dt, c->nip, c->D, c->up, c->dn);
        } else {
            VK_CUDA(cudaMemsetAsync(c->status, 0, sizeof(int) * c->ncol, c->stream));
            rc = launch_factor_fused(c, c->y, c->dt, c->D, c->up, c->dn, c->W, c->status, 1);
            if (rc == VK_ERR_UNSUPPORTED) set_error("the fused kernel does not support this network / block size");
        }
        if (rc) return rc;
        double *dD = nullptr, *dU = nullptr, *dL = nullptr;
        VK_CUDA(cudaMalloc((void **)&dD, sizeof(double) * nv * c->ni));
        VK_CUDA(cudaMalloc((void **)&dU, sizeof(double) * nv));
        VK_CUDA(cudaMalloc((void **)&dL, sizeof(double) * nv));
        unpad_system<<<(unsigned)((size_t)c->ncol * c->nz), 256, 0, c->stream>>>(c->ni, c->nip, c->D, c->up, c->dn, dD, dU, dL);
        VK_CUDA(cudaMemcpyAsync(D, dD, sizeof(double) * nv * c->ni, cudaMemcpyDeviceToHost, c->stream));
        VK_CUDA(cudaMemcpyAsync(up, dU, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
        VK_CUDA(cudaMemcpyAsync(dn, dL, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
        VK_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(dD); cudaFree(dU); cudaFree(dL);
        return VK_OK;
    }
    // dense (unpadded) output layout: ld = ni; D / up / dn buffers are large enough (nip >= ni)
    if ((rc = launch_lhs(c, c->y, c->dt, c->ni, c->D, c->up, c->dn))) return rc;
    VK_CUDA(cudaMemcpyAsync(D, c->D, sizeof(double) * nv * c->ni, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaMemcpyAsync(up, c->up, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaMemcpyAsync(dn, c->dn, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    return VK_OK;
}

__global__ void pad_system(int nz, int ni, int nip, size_t nblk, const double *Dd, const double *upd, const double *dnd, double *D,
                           double *up, double *dn)
{
    // dense [nblk][ni][ni] -> padded [nblk][nip][nip] with identity padding
    const size_t blk = blockIdx.x;
    if (blk >= nblk) return;
    for (int q = threadIdx.x; q < nip * nip; q += blockDim.x) {
        int r = q / nip, cc = q % nip;
        double v = (r < ni && cc < ni) ? Dd[blk * ni * ni + (size_t)r * ni + cc] : ((r == cc) ? 1.0 : 0.0);
        D[blk * nip * nip + q] = v;
    }
    for (int i = threadIdx.x; i < nip; i += blockDim.x) {
        up[blk * nip + i] = (i < ni) ? upd[blk * ni + i] : 0.0;
        dn[blk * nip + i] = (i < ni) ? dnd[blk * ni + i] : 0.0;
    }
}

int vk_blocktri_solve(vk_column *c, const double *D, const double *up, const double *dn, const double *rhs, double *x,
                      int refine, int *status)
{
    if (!c) { set_error("null handle"); return VK_ERR_INVALID; }
    if (!D || !up || !dn || !rhs || !x) { set_error("null buffer"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    {   // a non-sticky error left behind by an unrelated earlier runtime call must not be blamed on this one
        cudaError_t stale = cudaGetLastError();
        if (stale != cudaSuccess && getenv("VK_DEBUG")) fprintf(stderr, "vulcan_b200: stale CUDA error at vk_blocktri_solve entry: %s\n", cudaGetErrorString(stale));
    }
    const size_t nblk = (size_t)c->ncol * c->nz, nv = nblk * c->ni;
    c->cr_now = c->use_cr;
    double *Dd = nullptr, *upd = nullptr, *dnd = nullptr;
    VK_CUDA(cudaMalloc((void **)&Dd, sizeof(double) * nv * c->ni));
    cudaError_t e = cudaMalloc((void **)&upd, sizeof(double) * nv);
    if (e == cudaSuccess) e = cudaMalloc((void **)&dnd, sizeof(double) * nv);
    if (e == cudaSuccess) e = cudaMemcpyAsync(Dd, D, sizeof(double) * nv * c->ni, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(upd, up, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dnd, dn, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->f, rhs, sizeof(double) * nv, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->status, 0, sizeof(int) * c->ncol, c->stream);
    int rc = VK_OK;
    if (e == cudaSuccess) {
        pad_system<<<(unsigned)nblk, 256, 0, c->stream>>>(c->nz, c->ni, c->nip, nblk, Dd, upd, dnd, c->D, c->up, c->dn);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "vk_blocktri_solve staging");
    if (rc == VK_OK) rc = launch_factor(c, c->D, c->up, c->dn, c->W, c->status);
    if (rc == VK_OK) rc = launch_solve(c, c->W, c->up, c->dn, c->f, c->k1, c->z);
    if (rc == VK_OK) rc = launch_refine(c, c->D, c->up, c->dn, c->W, c->f, c->k1, refine, nullptr);   // refine < 0: safeguarded pass on every column
    if (rc == VK_OK) {
        e = cudaMemcpyAsync(x, c->k1, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && status) e = cudaMemcpyAsync(status, c->status, sizeof(int) * c->ncol, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "vk_blocktri_solve readback");
    }
    cudaStreamSynchronize(c->stream);
    cudaFree(Dd); if (upd) cudaFree(upd); if (dnd) cudaFree(dnd);
    return rc;
}

int vk_refine_stats(vk_column *c, int *kept, int *tried)
{
    if (!c) { set_error("null handle"); return VK_ERR_INVALID; }
    VK_CUDA(cudaSetDevice(c->net->device));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    if (kept) VK_CUDA(cudaMemcpy(kept, c->refine_kept, sizeof(int) * c->ncol, cudaMemcpyDeviceToHost));
    if (tried) VK_CUDA(cudaMemcpy(tried, c->refine_tried, sizeof(int) * c->ncol, cudaMemcpyDeviceToHost));
    return VK_OK;
}

int vk_last_kernel_ms(vk_column *c, float *ms_total, float *ms_factor)
{
    if (!c) { set_error("null handle"); return VK_ERR_INVALID; }
    if (ms_total) *ms_total = c->last_ms_total;
    if (ms_factor) *ms_factor = c->last_ms_factor;
    return VK_OK;
}

int vk_device_buffers(vk_column *c, void **y_dev, void **ymix_dev, void **sol_dev, void **k_dev)
{
    if (!c) { set_error("null handle"); return VK_ERR_INVALID; }
    if (y_dev) *y_dev = c->y;
    if (ymix_dev) *ymix_dev = c->ymix;
    if (sol_dev) *sol_dev = c->sol;
    if (k_dev) *k_dev = c->k;
    return VK_OK;
}

int vk_stream(vk_column *c, void **cuda_stream)
{
    if (!c || !cuda_stream) { set_error("null argument"); return VK_ERR_INVALID; }
    *cuda_stream = (void *)c->stream;
    return VK_OK;
}

}  // extern "C"

// profiling aid (declared in the public header): time `reps` launches of one kernel of the step on the resident state.
// which: 0 = lhs, 1 = rhs (stage 1), 2 = factor (+ the fused forward elimination of the first solve where dt allows, as in the step),
// 3 = first solve of the step (backward sweep only for the columns whose forward elimination the factorisation did), 4 = solve (forward +
// backward), 5 = fused assembly + factorisation, 6 = the emitted chemdf kernel alone
extern "C" int vk_debug_time_kernel(vk_column *c, int which, int reps, float *ms)
{
    if (!c || !ms || reps < 1) return VK_ERR_INVALID;
    VK_CUDA(cudaSetDevice(c->net->device));
    c->cr_now = c->use_cr;
    cudaEvent_t a, b;
    VK_CUDA(cudaEventCreate(&a));
    VK_CUDA(cudaEventCreate(&b));
    int rc = VK_OK;
    VK_CUDA(cudaEventRecord(a, c->stream));
    for (int r = 0; r < reps && rc == VK_OK; r++) {
        if (which == 0) rc = vk::launch_lhs(c, c->y, c->dt, c->nip, c->D, c->up, c->dn);
        else if (which == 1) rc = vk::launch_rhs(c, c->y, c->f, nullptr, nullptr, nullptr, nullptr);
        else if (which == 2) rc = vk::launch_factor(c, c->D, c->up, c->dn, c->W, c->status, c->f);
        else if (which == 5) rc = vk::launch_factor_fused(c, c->y, c->dt, c->D, c->up, c->dn, c->W, c->status, 0);
        else if (which == 6) {
            if (!c->chem_tmp) rc = vk::launch_rhs(c, c->y, c->f, nullptr, nullptr, nullptr, nullptr);      // allocates the scratch of the emitted path
            if (rc == VK_OK && !c->chem_tmp) { set_error("this handle does not take the emitted chemistry path"); rc = VK_ERR_UNSUPPORTED; }
            if (rc == VK_OK) rc = vk::launch_chem_emitted(c, c->y, nullptr, c->chem_tmp, c->ysum_tmp, nullptr);
        }
        else if (which == 7) {
            if (!c->ysum_lhs_tmp) VK_CUDA(cudaMalloc((void **)&c->ysum_lhs_tmp, sizeof(double) * (size_t)c->ncol * c->nz));
            rc = vk::launch_jac_emitted(c, c->y, c->D, c->ysum_lhs_tmp);
        }
        else if (which == 3) rc = vk::launch_solve(c, c->W, c->up, c->dn, c->f, c->k1, c->z, c->act, c->fwd_valid ? c->fwd_done : nullptr);
        else rc = vk::launch_solve(c, c->W, c->up, c->dn, c->f, c->k1, c->z);
    }
    VK_CUDA(cudaEventRecord(b, c->stream));
    VK_CUDA(cudaStreamSynchronize(c->stream));
    VK_CUDA(cudaEventElapsedTime(ms, a, b));
    *ms /= reps;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return rc;
}
