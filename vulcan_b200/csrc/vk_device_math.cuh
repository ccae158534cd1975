// Device helpers shared by the -fmad=false translation units (vk_chem.cu, vk_step.cu).
#pragma once

namespace vk {

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src) for n <= 128 contiguous doubles;
// np.sum(y, axis=1) at op.py:1505-1507, 2990-2993 reduces each row of ni species this way.
static __device__ __forceinline__ double np_pairwise_le128(const double *a, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 += a[i]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
        r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; i++) res += a[i];
    return res;
}
// n <= 256 (one level of numpy's recursion); vk_network_create rejects ni > 253
static __device__ __forceinline__ double np_pairwise(const double *a, int n)
{
    if (n <= 128) return np_pairwise_le128(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_le128(a, n2) + np_pairwise_le128(a + n2, n - n2);
}
// row sum over the gas species or over all species (`tmp` unused, kept for the call sites)
static __device__ __forceinline__ double row_sum(const double *yrow, int ni, int n_gas, const int *gas, double *tmp)
{
    if (n_gas > 0) {
        // np.sum(y[:, gas_indx], axis=1): the fancy-indexed copy is F-ordered, numpy reduces it column by column, i.e. a plain
        // left-to-right sum in gas_indx order (NOT pairwise) - pinned by the Jupiter / Earth fixtures
        double acc = yrow[gas[0]];
        for (int i = 1; i < n_gas; i++) acc += yrow[gas[i]];
        return acc;
    }
    return np_pairwise(yrow, ni);
}

// warp-cooperative version of row_sum for up to 4 rows at once: lanes 8*q .. 8*q+7 reduce row q with numpy's 8 accumulators
// (identical association: r_m = a[m] + a[8+m] + ..., ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail sequentially).
// n <= 128.  Call with all 32 lanes; returns the sum of row (lane >> 3) in every lane of that group.
static __device__ __forceinline__ double np_pairwise_group8(const double *a, int n, bool active)
{
    const int m = threadIdx.x & 7;
    double r = 0.0, res;
    if (n < 8) {
        res = 0.;
        if (active) for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    const int nb = n - (n % 8);
    if (active) {
        r = a[m];
        for (int i = 8; i < nb; i += 8) r += a[i + m];
    }
    // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) : butterfly over the 8 lanes of the group keeps exactly this association
    r = r + __shfl_xor_sync(0xffffffffu, r, 1);
    r = r + __shfl_xor_sync(0xffffffffu, r, 2);
    r = r + __shfl_xor_sync(0xffffffffu, r, 4);
    res = r;
    if (active) for (int i = nb; i < n; i++) res += a[i];
    return res;
}

static __device__ __forceinline__ double posv(double v) { return (v > 0) ? v : 0.0 * v; }  // (v>0)*v
static __device__ __forceinline__ double negv(double v) { return (v < 0) ? v : 0.0 * v; }  // (v<0)*v

}  // namespace vk
