"""TEST INFRASTRUCTURE - numpy restatement of the reference's rate-coefficient set-up, the checker for vk_compute_k.

Follows, operation by operation (numpy on numpy, so the result is expected to be BIT-IDENTICAL to the reference's var.k):
  ReadRate.read_rate   op.py:63-271   two-body / radiative `a*T**n*exp(-E/T)`; three-body with high-pressure limit
                                      `k0/(1 + k0*M/k_inf)`; three-body k0 only; the special `OH + CH3 + M -> CH3OH + M` form
                                      (op.py:197-207); condensation / photo / ion rows zero
  lim_lowT_rates       op.py:320-342  low-temperature caps (use_lowT_limit_rates)
  rev_rate             op.py:289-309  k[i] = k[i-1] / Gibbs(i-1, T) for even i < stop_rev_indx, zero above
  Gibbs                make_chem_funs.py:568-580 + thermo/gibbs_text.txt:13-27 (NASA-9 h/RT - s/R, switch at 1000 K, `(corr*T)**dnu`)
  remove_rate          op.py:311-317
Pinned by tests/test_rates.py against <cfg>_static.npz['k'] (written by the unmodified reference)."""
import numpy as np

KB = 1.38064852e-16
CORR = KB / 1.e6                     # gibbs_text.txt:1
M_SLOT = -1


def h_RT(T, a):                      # gibbs_text.txt:14-15
    return -a[0] / T**2 + a[1] * np.log(T) / T + a[2] + a[3] * T / 2. + a[4] * T**2 / 3. + a[5] * T**3 / 4. + a[6] * T**4 / 5. + a[8] / T


def s_R(T, a):                       # gibbs_text.txt:18-19
    return -a[0] / T**2 / 2. - a[1] / T + a[2] * np.log(T) + a[3] * T + a[4] * T**2 / 2. + a[5] * T**3 / 3. + a[6] * T**4 / 4. + a[9]


def g_RT(T, a_low, a_high):          # gibbs_text.txt:22-25
    return (T < 1000) * (h_RT(T, a_low) - s_R(T, a_low)) + (T >= 1000) * (h_RT(T, a_high) - s_R(T, a_high))


def gibbs_K(net, r, T, nasa9):
    """chem_funs.Gibbs(r.id, T): exp(-( -n*g(reactants, written order) + n*g(products) )) * (corr*T)**(n_reac - n_prod)."""
    acc = None
    nreac = nprod = 0
    for sign, side in ((-1, r.reac), (1, r.prod)):
        for slot, n in side:
            if slot == M_SLOT:
                continue
            term = n * g_RT(T, nasa9[slot, 0:10], nasa9[slot, 10:20])
            if acc is None:
                acc = -term if sign < 0 else term          # '-1*g' at the head of the expression
            else:
                acc = acc - term if sign < 0 else acc + term
            if sign < 0:
                nreac += n
            else:
                nprod += n
    K = np.exp(-(acc))
    if nprod - nreac != 0:
        K = K * (CORR * T) ** (nreac - nprod)
    return K


def compute_k(net, Tco, M, nasa9, remove_list=(), use_lowT_limit_rates=False):
    """-> k [nr+1, nz] like ref_session.pack_k(var) BEFORE any photolysis update (photo rows zero)."""
    from vulcan_b200.network import (SECTION_2BODY, SECTION_3BODY, SECTION_3BODY_K0, SECTION_SPECIAL, SECTION_RECOMB)
    Tco, M = np.asarray(Tco, dtype=float), np.asarray(M, dtype=float)
    nz = Tco.size
    k = np.zeros((net.nr + 1, nz))
    for r in net.reactions:
        i = r.id
        if r.section in (SECTION_2BODY, SECTION_3BODY, SECTION_3BODY_K0, SECTION_RECOMB):
            a, n, E = r.rate_cols[0], r.rate_cols[1], r.rate_cols[2]
            k[i] = a * Tco**n * np.exp(-E / Tco)
            if r.section == SECTION_3BODY and len(r.rate_cols) >= 6:
                a_inf, n_inf, E_inf = r.rate_cols[3], r.rate_cols[4], r.rate_cols[5]
                k_inf = a_inf * Tco**n_inf * np.exp(-E_inf / Tco)
                k[i] = k[i] / (1 + k[i] * M / k_inf)
        elif r.section == SECTION_SPECIAL and r.text == 'OH + CH3 + M -> CH3OH + M':
            k[i] = 1.932E3 * Tco**-9.88 * np.exp(-7544. / Tco) + 5.109E-11 * Tco**-6.25 * np.exp(-1433. / Tco)
            k_inf = 1.031E-10 * Tco**-0.018 * np.exp(16.74 / Tco)
            Fc = 0.1855 * np.exp(-Tco / 155.8) + 0.8145 * np.exp(-Tco / 1675.) + np.exp(-4531. / Tco)
            nn = 0.75 - 1.27 * np.log(Fc)
            ff = np.exp(np.log(Fc) / (1. + (np.log(k[i] * M / k_inf) / nn)**2))
            k[i] = k[i] / (1 + k[i] * M / k_inf) * ff
    if use_lowT_limit_rates:
        for r in net.reactions:
            i = r.id
            if r.text == 'H + CH3 + M -> CH4 + M':
                mask = Tco <= 277.5
                k0 = 6e-29
                kinf = 2.06E-10 * Tco**-0.4
                k[i][mask] = (k0 / (1. + k0 * M / kinf))[mask]
            elif r.text == 'H + C2H4 + M -> C2H5 + M':
                k[i][Tco <= 300] = 3.7E-30
            elif r.text == 'H + C2H5 + M -> C2H6 + M':
                k[i][Tco <= 200] = 2.49E-27
    for r in net.reactions:
        i = r.id + 1
        if i < net.stop_rev_indx and i not in remove_list:
            k[i] = k[i - 1] / gibbs_K(net, r, Tco, nasa9)
    for i in remove_list:
        k[i] = 0.
    return k
