"""Rate-coefficient tables for the device kernel (csrc/vk_rates.cu): the network's Arrhenius / falloff parameters, the special
forms and low-temperature caps the reference hard-codes, and the Gibbs (equilibrium-constant) term lists of the reverse rates.

Replaces what the reference does at set-up on the host (ReadRate.read_rate, lim_lowT_rates, rev_rate, remove_rate, op.py:63-342, and
the generated chem_funs.Gibbs) so that ensembles with per-column T-P profiles need no 1 MB-per-column upload of `k`.  The NASA-9
coefficients come from `thermo/NASA9/<species>.txt` of a VULCAN checkout (`load_nasa9`) exactly like gibbs_text.txt:6-11 reads them."""
import os

import numpy as np

from .network import (M_SLOT, SECTION_2BODY, SECTION_3BODY, SECTION_3BODY_K0, SECTION_RECOMB, SECTION_SPECIAL)

KIND_ZERO, KIND_ARRHENIUS, KIND_FALLOFF, KIND_OH_CH3 = 0, 1, 2, 3
# lim_lowT_rates (op.py:320-342): reaction text -> (cap kind, threshold temperature, value)
LOW_T_CAPS = {'H + CH3 + M -> CH4 + M': (1, 277.5, 6e-29), 'H + C2H4 + M -> C2H5 + M': (2, 300., 3.7E-30),
              'H + C2H5 + M -> C2H6 + M': (2, 200., 2.49E-27)}


def load_nasa9(species, thermo_dir, alias=None):
    """[ni, 20] from <thermo_dir>/NASA9/<species>.txt (gibbs_text.txt:6-11)."""
    alias = alias or {}
    out = np.zeros((len(species), 20))
    for i, sp in enumerate(species):
        out[i] = np.loadtxt(os.path.join(thermo_dir, "NASA9", alias.get(sp, sp) + ".txt")).flatten()[:20]
    return out


class RateTable(object):
    def __init__(self, network, nasa9, remove_list=(), use_lowT_limit_rates=False):
        net = network
        npair = net.nr // 2
        self.npair = npair
        self.kind = np.zeros(npair, dtype=np.int32)
        self.arrhenius = np.zeros((npair, 6))
        self.cap_kind = np.zeros(npair, dtype=np.int32)
        self.cap = np.zeros((npair, 2))
        self.reverse = np.zeros(npair, dtype=np.uint8)
        self.removed = np.zeros(npair, dtype=np.uint8)
        self.dnu = np.zeros(npair, dtype=np.int32)
        gptr, gsp, gnu = [0], [], []
        remove = set(int(x) for x in remove_list)
        for p, r in enumerate(net.reactions):
            assert r.id == 2 * p + 1
            if r.section in (SECTION_2BODY, SECTION_3BODY, SECTION_3BODY_K0, SECTION_RECOMB):
                self.kind[p] = KIND_ARRHENIUS
                self.arrhenius[p, :3] = r.rate_cols[:3]
                if r.section == SECTION_3BODY and len(r.rate_cols) >= 6:        # op.py:176-183
                    self.kind[p] = KIND_FALLOFF
                    self.arrhenius[p, 3:6] = r.rate_cols[3:6]
            elif r.section == SECTION_SPECIAL and r.text == 'OH + CH3 + M -> CH3OH + M':   # op.py:197
                self.kind[p] = KIND_OH_CH3
            if use_lowT_limit_rates and r.text in LOW_T_CAPS:
                ck, tcap, val = LOW_T_CAPS[r.text]
                self.cap_kind[p], self.cap[p] = ck, (tcap, val)
            nreac = nprod = 0
            for sign, side in ((-1, r.reac), (1, r.prod)):                      # make_chem_funs.py:568-575
                for slot, n in side:
                    if slot == M_SLOT:
                        continue
                    gsp.append(slot)
                    gnu.append(sign * n)
                    if sign < 0:
                        nreac += n
                    else:
                        nprod += n
            gptr.append(len(gsp))
            self.dnu[p] = nreac - nprod
            if r.id + 1 < net.stop_rev_indx and (r.id + 1) not in remove:     # op.py:291-304
                self.reverse[p] = 1
            self.removed[p] = (1 if r.id in remove else 0) | (2 if (r.id + 1) in remove else 0)
        self.gibbs_ptr = np.array(gptr, dtype=np.int32)
        self.gibbs_sp = np.array(gsp, dtype=np.int32)
        self.gibbs_nu = np.array(gnu, dtype=np.int32)
        self.nasa9 = np.ascontiguousarray(nasa9, dtype=np.float64)
        assert self.nasa9.shape == (net.ni, 20)
