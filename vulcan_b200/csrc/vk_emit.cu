// Registry of the EMITTED chemistry kernels compiled into this library (csrc/gen/chemdf_*.cu, written by vulcan_b200/emit.py - the
// network -> CUDA generator that replaces make_chem_funs.py:113-430) and the launch of chemdf through them.
#include "vk_internal.cuh"
#include "vk_emit_rt.cuh"

#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace vk { namespace emitted {

static std::vector<EmitEntry> &registry()
{
    static std::vector<EmitEntry> r;      // constructed on first use: the registrars run during static initialisation of other units
    return r;
}

void emit_register(const EmitEntry &e) { registry().push_back(e); }

const EmitEntry *emit_find(unsigned long long hash, int ni, int nr)
{
    for (const EmitEntry &e : registry())
        if (e.hash == hash && e.ni == ni && e.nr == nr) return &e;
    return nullptr;
}

int emit_set_smem(const void *func, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> done;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(func, dev);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return 0;
    if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return 1;
    done[key] = bytes;
    return 0;
}

}  // namespace emitted

// FNV-1a (64 bit) over the tables that define chemdf - the same bytes vulcan_b200/emit.py hashes
unsigned long long network_table_hash(const vk_network_desc *d)
{
    unsigned long long h = 0xcbf29ce484222325ull;
    auto feed = [&](const void *p, size_t n) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
    };
    const int head[3] = {d->ni, d->nr, d->maxf};
    feed(head, sizeof(head));
    feed(d->rate_fac, sizeof(int) * (size_t)(d->nr + 1) * d->maxf);
    feed(d->rate_pow, sizeof(int) * (size_t)(d->nr + 1) * d->maxf);
    feed(d->rhs_ptr, sizeof(int) * (size_t)(d->ni + 1));
    feed(d->rhs_pair, sizeof(int) * (size_t)d->n_rhs);
    feed(d->rhs_coef, sizeof(double) * (size_t)d->n_rhs);
    return h;
}

const void *emit_lookup(unsigned long long hash, int ni, int nr) { return emitted::emit_find(hash, ni, nr); }

// chemdf of y (stage 1) or of y + k1/r (stage 2, also written to yk2_out) into chem_out [ncol][nz][ni], and the layer sums of that state
// into ysum_out [ncol][nz], through the emitted kernel of this network; the batch shares its rate coefficients (c->k_cs == 0)
int launch_chem_emitted(vk_column *c, const double *y_dev, const double *k1, double *chem_out, double *ysum_out, double *yk2_out)
{
    const emitted::EmitEntry *e = static_cast<const emitted::EmitEntry *>(c->net->emit);
    if (!e) { set_error("no emitted chemistry kernel for this network"); return VK_ERR_UNSUPPORTED; }
    if (c->k_cs != 0 && !c->k_static_shared) { set_error("the emitted chemistry kernel needs rate coefficients shared by the batch"); return VK_ERR_UNSUPPORTED; }
    emitted::EmitArgs a;
    a.nz = c->nz; a.ncol = c->ncol;
    a.y = y_dev; a.k1 = k1; a.yk2_out = yk2_out; a.k = c->k; a.k_cs = c->k_cs;
    a.M = c->atm.M; a.M_cs = c->atm.csz;
    a.chem = chem_out; a.ysum = ysum_out; a.n_gas = c->atm.n_gas; a.gas_indx = c->atm.gas_indx; a.act = c->act;
    if (e->fn(a, c->stream)) return cuda_fail(cudaGetLastError(), "emitted chemdf kernel");
    return VK_OK;
}

// -J of every layer as dense rows into D_out ([ncol][nz][nip][nip]) and the layer sums lhs_jac_tot reads into ysum_out, through the emitted
// Jacobian kernel of this network
int launch_jac_emitted(vk_column *c, const double *y_dev, double *D_out, double *ysum_out)
{
    const emitted::EmitEntry *e = static_cast<const emitted::EmitEntry *>(c->net->emit);
    if (!e || !e->jac) { set_error("no emitted Jacobian kernel for this network"); return VK_ERR_UNSUPPORTED; }
    if (c->k_cs != 0 && !c->k_static_shared) { set_error("the emitted Jacobian kernel needs rate coefficients shared by the batch"); return VK_ERR_UNSUPPORTED; }
    emitted::EmitJacArgs a;
    a.nz = c->nz; a.ncol = c->ncol; a.y = y_dev; a.k = c->k; a.k_cs = c->k_cs;
    a.M = c->atm.M; a.M_cs = c->atm.csz; a.D = D_out; a.ysum = ysum_out;
    a.n_gas = c->atm.n_gas_lhs; a.gas_indx = c->atm.gas_indx_lhs; a.act = c->act;
    if (e->jac(a, c->stream)) return cuda_fail(cudaGetLastError(), "emitted Jacobian kernel");
    return VK_OK;
}
bool emit_has_jac(const void *entry) { return entry && static_cast<const emitted::EmitEntry *>(entry)->jac != nullptr; }
// is k index i one of the rows the emitted kernels read per column (photolysis / ionisation / condensation)?
bool emit_row_is_dynamic(const void *entry, int i)
{
    const emitted::EmitEntry *e = static_cast<const emitted::EmitEntry *>(entry);
    if (!e) return false;
    for (int q = 0; q < e->n_dyn; q++)
        if (e->dyn[q] == i) return true;
    return false;
}

}  // namespace vk
