"""Photolysis input tables from the reference's data files (the wire formats on the input side of the path):
the wavelength grid, the absorption / dissociation / ionisation / Rayleigh cross sections binned onto it, their temperature-
dependent versions per layer, and the stellar flux at the top of the atmosphere.  These are the arrays `vk_photo_setup` takes
(include/vulcan_b200.h) and that the reference builds at set-up time in `ReadRate.make_bins_read_cross` (op.py:495-780) and
`Atm.read_sflux` (build_atm.py:617-655); with this module the package can be fed from a `thermo/photo_cross/` tree without the
reference's Python modules.

File formats (thermo/photo_cross/):
  <sp>/<sp>_cross.csv            header line, then  lambda[nm], total absorption, photodissociation[, photoionisation]  (cm^2)
  <sp>/<sp>_branch.csv           2 header lines (the 2nd names the columns: lambda, br_ratio_1, ...), branching ratios
  <sp>/<sp>_ion_branch.csv       same for the ionisation branches
  <sp>/<sp>_cross_<T>K.csv       temperature-dependent cross sections (species in T_cross_sp), same columns as _cross.csv
  rayleigh/<sp>_scat.txt         header line, then  lambda, cross section
  thresholds.txt                 header line, then  species  threshold wavelength
  stellar flux file              header line, then  lambda, flux at the stellar surface

Interpolation follows what the reference's `scipy.interpolate.interp1d(..., bounds_error=False, fill_value=...)` evaluates for these
1-D linear cases (it delegates to `numpy.interp` and then overwrites the out-of-range points with the fill value), vectorised over
the wavelength grid instead of one Python call per bin and layer.  Checked against the tables the unmodified reference built for
the fixture configs (tests/test_photo_tables.py)."""
import os

import numpy as np

R_SUN = 6.957e10       # phy_const.py:9
AU = 1.49597871e13    # phy_const.py:8


def _interp(x, y, xnew, fill=0.0):
    """interp1d(x, y, bounds_error=False, fill_value=fill)(xnew) for 1-D linear data; fill = scalar or (below, above)"""
    x, y, xnew = np.asarray(x, dtype=float), np.asarray(y, dtype=float), np.asarray(xnew, dtype=float)
    if np.any(np.diff(x) < 0):        # interp1d sorts its abscissae (assume_sorted=False); CH3SH_branch.csv ends with two swapped rows
        ind = np.argsort(x, kind="mergesort")
        x, y = x[ind], y[ind]
    out = np.interp(xnew, x, y)
    lo, hi = (fill, fill) if np.isscalar(fill) else fill
    out = np.where(xnew < x[0], lo, out)
    out = np.where(xnew > x[-1], hi, out)
    return out


def _read_cross(path, use_ion):
    names = ["lambda", "cross", "disso", "ion"] if use_ion else ["lambda", "cross", "disso"]
    return np.genfromtxt(path, dtype=float, delimiter=",", skip_header=1, names=names)          # op.py:523-531


def _read_ratio(path):
    return np.genfromtxt(path, dtype=float, delimiter=",", skip_header=1, names=True)            # op.py:526, 536


def wavelength_grid(bin_min, bin_max, dbin1, dbin2, dbin_12trans):
    """two uniform grids joined at dbin_12trans (op.py:582-586)"""
    if bin_min <= dbin_12trans <= bin_max:
        return np.concatenate((np.arange(bin_min, dbin_12trans, dbin1), np.arange(dbin_12trans, bin_max, dbin2)))
    return np.arange(bin_min, bin_max, dbin1)


class PhotoTables(object):
    """bins, cross[sp], cross_J[(sp, br)], cross_scat[sp], cross_T[sp] (nz, nbin), cross_J_T[(sp, br)] (nz, nbin),
    cross_Jion[(sp, br)]  -  the attributes `make_bins_read_cross` leaves on `var`."""

    def __init__(self, cross_folder, species, photo_sp, n_branch, scat_sp, def_bin_min, def_bin_max, dbin1, dbin2, dbin_12trans,
                 T_cross_sp=(), Tco=None, ion_sp=(), ion_branch=None, use_ion=False):
        photo_sp, ion_sp = list(photo_sp), list(ion_sp) if use_ion else []
        absp = photo_sp + [s for s in ion_sp if s not in photo_sp]
        lab = np.genfromtxt(os.path.join(cross_folder, "thresholds.txt"), dtype=str, usecols=0)
        lmd = np.genfromtxt(os.path.join(cross_folder, "thresholds.txt"), skip_header=1)[:, 1]
        self.threshold = {l: v for l, v in zip(lab, lmd) if l in species}                        # op.py:513-517
        raw, ratio, ion_ratio, raw_T = {}, {}, {}, {}
        bin_min = bin_max = diss_max = None
        for sp in photo_sp + [s for s in ion_sp]:
            if sp in raw:
                continue
            raw[sp] = _read_cross(os.path.join(cross_folder, sp, sp + "_cross.csv"), use_ion)
            if sp in ion_sp:
                ion_ratio[sp] = _read_ratio(os.path.join(cross_folder, sp, sp + "_ion_branch.csv"))
            if sp in photo_sp:
                ratio[sp] = _read_ratio(os.path.join(cross_folder, sp, sp + "_branch.csv"))
            if sp in T_cross_sp:                                                                  # op.py:539-553
                temps = []
                for fn in os.listdir(os.path.join(cross_folder, sp)):
                    if fn.startswith(sp) and fn.endswith("K.csv"):
                        temps.append(int(fn.replace(sp, "").replace("_cross_", "").replace("K.csv", "")))
                for tt in temps:
                    raw_T[(sp, tt)] = _read_cross(os.path.join(cross_folder, sp, "%s_cross_%dK.csv" % (sp, tt)), use_ion)
                raw_T[(sp, 300)] = raw[sp]
            if raw[sp]["cross"][0] == 0 or raw[sp]["cross"][-1] == 0:
                raise IOError("Please remove the zeros in the cross file of " + sp)
            lo, hi = raw[sp]["lambda"][0], raw[sp]["lambda"][-1]
            bin_min = lo if bin_min is None else min(bin_min, lo)
            bin_max = hi if bin_max is None else max(bin_max, hi)
            diss_max = self.threshold[sp] if diss_max is None else max(diss_max, self.threshold[sp])
        bin_min = max(bin_min, def_bin_min)                                                       # op.py:575-576
        bin_max = min(bin_max, def_bin_max, diss_max)
        self.bins = bins = wavelength_grid(bin_min, bin_max, dbin1, dbin2, dbin_12trans)
        self.nbin = len(bins)
        self.dbin1, self.dbin2 = dbin1, dbin2
        self.cross, self.cross_J, self.cross_scat, self.cross_T, self.cross_J_T, self.cross_Jion = {}, {}, {}, {}, {}, {}
        br_ratio = {}
        for sp in photo_sp:                                                                       # op.py:620-637
            r = raw[sp]
            self.cross[sp] = _interp(r["lambda"], r["cross"], bins)
            diss = _interp(r["lambda"], r["disso"], bins)
            for i in range(1, n_branch[sp] + 1):
                col = ratio[sp]["br_ratio_%d" % i]
                br_ratio[(sp, i)] = _interp(ratio[sp]["lambda"], col, bins, (col[0], col[-1]))
                self.cross_J[(sp, i)] = diss * br_ratio[(sp, i)]
        for sp in [s for s in T_cross_sp if s in photo_sp]:
            self._temperature_tables(sp, raw_T, n_branch[sp], br_ratio, np.asarray(Tco, dtype=float))
        for sp in ion_sp:                                                                         # op.py:749-768
            r = raw[sp]
            if sp not in photo_sp:
                self.cross[sp] = _interp(r["lambda"], r["cross"], bins)
            ion = _interp(r["lambda"], r["ion"], bins)
            for i in range(1, ion_branch[sp] + 1):
                col = ion_ratio[sp]["br_ratio_%d" % i]
                self.cross_Jion[(sp, i)] = ion * _interp(ion_ratio[sp]["lambda"], col, bins, (col[0], col[-1]))
        for sp in scat_sp:                                                                        # op.py:771-780
            sc = np.genfromtxt(os.path.join(cross_folder, "rayleigh", sp + "_scat.txt"), dtype=float, skip_header=1, names=["lambda", "cross"])
            self.cross_scat[sp] = _interp(sc["lambda"], sc["cross"], bins)

    # ------------------------------------------------------------------------------------------------ op.py:642-745
    def _temperature_tables(self, sp, raw_T, nbr, br_ratio, Tco):
        bins, nz = self.bins, len(Tco)
        T_list = np.array(sorted(t for (s, t) in raw_T if s == sp))
        cT = np.zeros((nz, self.nbin))
        cJT = {i: np.zeros((nz, self.nbin)) for i in range(1, nbr + 1)}

        def log10_floor(v):                       # log10 with -inf replaced by -100 (op.py:675-677)
            with np.errstate(divide="ignore"):
                lg = np.log10(v)
            return np.where(np.isinf(lg), -100., lg)

        for lev, Tz in enumerate(Tco):
            below, above = T_list[T_list <= Tz], T_list[T_list > Tz]
            if len(below) and len(above):         # between two tabulated temperatures: log10 in the cross section, linear in T
                Tlow, Thigh = below.max(), above.min()
                rl, rh = raw_T[(sp, Tlow)], raw_T[(sp, Thigh)]
                inside = (bins >= max(rl["lambda"][0], rh["lambda"][0])) & (bins <= min(rl["lambda"][-1], rh["lambda"][-1]))

                def blend(key):
                    lo, hi = log10_floor(_interp(rl["lambda"], rl[key], bins)), log10_floor(_interp(rh["lambda"], rh[key], bins))
                    val = np.where(Tz == Tlow, lo, (hi - lo) / (Thigh - Tlow) * (Tz - Tlow) + lo)      # numpy.interp on two points
                    return val, np.where(val == -100., 0., 10. ** val)
                _, c = blend("cross")
                cT[lev] = np.where(inside, c, self.cross[sp])
                # (the reference never stores the 0 of a -100 result for the total cross section: `==` instead of `=`, op.py:680 -
                #  the entry keeps its initial 0, which is the same value)
                _, cj = blend("disso")
                for i in range(1, nbr + 1):
                    cJT[i][lev] = np.where(inside, cj * br_ratio[(sp, i)], self.cross_J[(sp, i)])
            else:
                edge = T_list.min() if not len(T_list[T_list < Tz]) else T_list.max()             # colder / hotter than every table
                if edge == 300:
                    cT[lev] = self.cross[sp]
                    for i in range(1, nbr + 1):
                        cJT[i][lev] = self.cross_J[(sp, i)]
                else:
                    r = raw_T[(sp, edge)]
                    inside = (bins >= r["lambda"][0]) & (bins <= r["lambda"][-1])
                    cT[lev] = np.where(inside, _interp(r["lambda"], r["cross"], bins), self.cross[sp])
                    dj = _interp(r["lambda"], r["disso"], bins)
                    for i in range(1, nbr + 1):
                        cJT[i][lev] = np.where(inside, dj * br_ratio[(sp, i)], self.cross_J[(sp, i)])
        self.cross_T[sp] = cT
        for i in range(1, nbr + 1):
            self.cross_J_T[(sp, i)] = cJT[i]


def stellar_flux_top(sflux_file, bins, r_star, orbit_radius, dbin_12trans):
    """`Atm.read_sflux` (build_atm.py:617-631): flux at the stellar surface scaled to the planet's orbit and interpolated onto the bins;
    returns (sflux_top, index of the first bin of the coarse grid or -1)."""
    raw = np.genfromtxt(sflux_file, dtype=float, skip_header=1, names=["lambda", "flux"])
    top = _interp(raw["lambda"], raw["flux"] * (r_star * R_SUN / (AU * orbit_radius)) ** 2, bins)
    hit = np.nonzero(bins == dbin_12trans)[0]
    return top, (int(hit[-1]) if len(hit) else -1)


def default_bin_range(sflux_file):
    """store.Variables: def_bin_min / def_bin_max from the stellar spectrum (store.py:79-80)"""
    raw = np.genfromtxt(sflux_file, dtype=float, skip_header=1, names=["lambda", "flux"])
    return max(raw["lambda"][0], 2.), min(raw["lambda"][-1], 700.)
