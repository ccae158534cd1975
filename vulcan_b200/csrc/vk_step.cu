// Ros2 stage algebra, error estimate and step bookkeeping (op.py:2932-2993, 2447-2487), one block per (column, layer).
#include <algorithm>

#include "vk_internal.cuh"
#include "vk_device_math.cuh"

namespace vk {

struct EpiArgs {
    int nz, ni;
    const double *y, *ymix, *k1, *k2, *yk2;
    double *sol, *ymix_out;
    unsigned long long *delta_bits;   // [ncol], zeroed before launch; positive doubles order like their bit patterns
    StepOptsDev o;
    int n_gas;
    const int *gas_indx;
    const int *act;
};

// one WARP per (column, layer), 4 layers per block: the layer sum (numpy's pairwise association, np_pairwise_group8) is formed by 8 lanes
// instead of one thread walking the row while 127 wait
#define EPI_WPB 4
__global__ void __launch_bounds__(EPI_WPB * 32) epilogue_kernel(EpiArgs a, int n_layers_total)
{
    extern __shared__ double sm[];
    const int nz = a.nz, ni = a.ni;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lay = blockIdx.x * EPI_WPB + w;
    if (lay >= n_layers_total) return;
    const int col = lay / nz, j = lay - col * nz;
    if (a.act && !a.act[col]) return;
    double *srow = sm + (size_t)w * (2 * ni + 2);   // ni
    double *tmp = srow + ni;                        // ni (scratch of the gas-only sum)
    const size_t base = (size_t)lay * ni;
    const double r = 1. + 1. / sqrt(2.);
    const double c1 = 3. / (2. * r), c2 = 1 / (2. * r);
    double dmax = 0.0;
    bool has = false;
    for (int i = lane; i < ni; i += 32) {
        const size_t q = base + i;
        double s = a.y[q] + c1 * a.k1[q] + c2 * a.k2[q];                   // op.py:2932
        if (j == 0)
            for (int b = 0; b < a.o.n_fix_bot; b++)
                if (a.o.fix_bot_idx[b] == i) s = a.o.fix_bot_val[(size_t)col * a.o.n_fix_bot + b];   // op.py:2946
        double d = fabs(s - a.yk2[q]);                                     // op.py:2948
        if (a.o.fix_mask && a.o.fix_y && a.o.fix_mask[q]) s = a.o.fix_y[q]; // op.py:2960-2968
        if (a.ymix[q] < a.o.mtol) d = 0;
        if (s < a.o.atol) d = 0;
        if (a.o.zero_delta_row0 && j == 0) d = 0;
        if (a.o.delta_zero_sp && a.o.delta_zero_sp[i]) d = 0;
        if (s > 0) {
            double v = d / s;                                              // op.py:2985
            if (!has || v > dmax || v != v) dmax = v;
            has = true;
        }
        a.sol[q] = s;
        srow[i] = s;
    }
    // warp max of non-negative doubles (NaN has the largest bit pattern and therefore propagates like np.amax)
    unsigned long long bits = has ? (unsigned long long)__double_as_longlong(dmax) : 0ull;
    for (int off = 16; off > 0; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, bits, off);
        bits = (o > bits) ? o : bits;
    }
    if (lane == 0 && bits) atomicMax(a.delta_bits + col, bits);
    __syncwarp();
    double ssum;                                                            // op.py:2990-2993
    if (a.n_gas > 0) {
        ssum = 0.0;
        if (lane == 0) ssum = row_sum(srow, ni, a.n_gas, a.gas_indx, tmp);
        ssum = __shfl_sync(0xffffffffu, ssum, 0);
    } else {
        ssum = np_pairwise_group8(srow, ni, lane < 8);
        ssum = __shfl_sync(0xffffffffu, ssum, 0);
    }
    for (int i = lane; i < ni; i += 32) a.ymix_out[base + i] = srow[i] / ssum;
}

int launch_epilogue(vk_column *c)
{
    EpiArgs a;
    a.nz = c->nz; a.ni = c->ni;
    a.y = c->y; a.ymix = c->ymix; a.k1 = c->k1; a.k2 = c->k2; a.yk2 = c->yk2;
    a.sol = c->sol; a.ymix_out = c->ymix_out;
    a.delta_bits = reinterpret_cast<unsigned long long *>(c->delta);
    a.o = c->opts;
    a.n_gas = c->atm.n_gas; a.gas_indx = c->atm.gas_indx; a.act = c->act;
    VK_CUDA(cudaMemsetAsync(c->delta, 0, sizeof(double) * c->ncol, c->stream));
    const int n_layers = c->ncol * c->nz;
    epilogue_kernel<<<(n_layers + EPI_WPB - 1) / EPI_WPB, EPI_WPB * 32, sizeof(double) * EPI_WPB * (2 * c->ni + 2), c->stream>>>(a, n_layers);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// ODESolver.clip + loss (op.py:2447-2487), one block per column.
struct ClipArgs {
    int nz, ni, na;
    double *y;
    const double *ymix_in;
    double *ymix_out;
    const double *compo;              // [ni][na]
    const unsigned char *atom_skip;   // [na] or NULL
    double pos_cut, nega_cut, mtol;
    int n_gas;
    const int *gas_indx;
    double *atom_sum;                 // [ncol][na]
    double *small_y, *nega_y;         // [ncol] accumulated
    int *any_negative;                // [ncol]
    const int *act;
};

__global__ void __launch_bounds__(256) clip_kernel(ClipArgs a)
{
    extern __shared__ double sm[];
    const int nz = a.nz, ni = a.ni, na = a.na;
    const int col = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    if (a.act && !a.act[col]) return;
    double *red = sm;                 // nt * (na + 2)
    double *yc = a.y + (size_t)col * nz * ni;
    const double *ymc = a.ymix_in + (size_t)col * nz * ni;
    double small = 0.0, nega = 0.0;
    double asum[8];
    for (int q = 0; q < 8; q++) asum[q] = 0.0;
    int anyneg = 0;
    for (int q = tid; q < nz * ni; q += nt) {
        double v = yc[q];
        if (v < a.pos_cut && v >= 0) small += v;
        if (v > a.nega_cut && v <= 0) nega += v;
        if (v < a.pos_cut && v >= a.nega_cut) v = 0.;
        if (ymc[q] < a.mtol && v < 0) v = 0.;
        yc[q] = v;
        if (!(v >= 0)) anyneg = 1;
        const int i = q % ni;
        for (int at = 0; at < na; at++) asum[at] += a.compo[i * na + at] * v;
    }
    red[tid] = small;
    red[nt + tid] = nega;
    for (int at = 0; at < na; at++) red[(2 + at) * nt + tid] = asum[at];
    int anyb = __syncthreads_or(anyneg);
    for (int s = nt / 2; s > 0; s >>= 1) {
        if (tid < s)
            for (int v = 0; v < na + 2; v++) red[v * nt + tid] += red[v * nt + tid + s];
        __syncthreads();
    }
    if (tid == 0) {
        a.small_y[col] += fabs(red[0]);
        a.nega_y[col] += fabs(red[nt]);
        for (int at = 0; at < na; at++)
            if (!(a.atom_skip && a.atom_skip[at])) a.atom_sum[(size_t)col * na + at] = red[(2 + at) * nt];
        a.any_negative[col] = anyb;
    }
    // ymix = y / sum_gas(y) per layer, numpy pairwise order: one warp per layer (8 lanes form the sum, the row is written coalesced)
    const int w = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int j = w; j < nz; j += nw) {
        const double *row = yc + (size_t)j * ni;
        double s;
        if (a.n_gas > 0) {
            s = 0.0;
            if (lane == 0) s = row_sum(row, ni, a.n_gas, a.gas_indx, nullptr);
        } else {
            s = np_pairwise_group8(row, ni, lane < 8);
        }
        s = __shfl_sync(0xffffffffu, s, 0);
        for (int i = lane; i < ni; i += 32) a.ymix_out[((size_t)col * nz + j) * ni + i] = row[i] / s;
    }
}

int launch_clip(vk_column *c, double *y_dev, const double *ymix_in_dev, double *ymix_out_dev, int na, const double *compo_dev,
                const unsigned char *skip_dev, double pos_cut, double nega_cut, double *atom_sum_dev, double *small_dev,
                double *nega_dev, int *anyneg_dev)
{
    if (na > 8) { set_error("at most 8 elements in atom_list are supported"); return VK_ERR_UNSUPPORTED; }
    ClipArgs a{c->nz, c->ni, na, y_dev, ymix_in_dev, ymix_out_dev, compo_dev, skip_dev, pos_cut, nega_cut, c->opts.mtol,
               c->atm.n_gas, c->atm.gas_indx, atom_sum_dev, small_dev, nega_dev, anyneg_dev, c->act};
    const int nt = 256;
    const size_t smem = sizeof(double) * (size_t)nt * (na + 2);      // <= 20 KB
    clip_kernel<<<c->ncol, nt, smem, c->stream>>>(a);
    VK_CUDA(cudaGetLastError());
    return VK_OK;
}

}  // namespace vk
