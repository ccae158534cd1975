#!/usr/bin/env python
"""TEST INFRASTRUCTURE — stages a writable scratch copy of the read-only reference.

The reference (exoclime/VULCAN, /root/reference) cannot run in place: it rewrites its
network file (make_chem_funs.py:109), writes chem_funs.py into cwd, FastChem writes
under fastchem_vulcan/{input,output} (build_atm.py:84-131), and vulcan_cfg.py:11 names
a network file that does not exist.  This script copies the tree to a scratch
directory and applies the NON-NUMERICAL shims listed in SURVEY.md §8c:

  1. stub `matplotlib` / `PIL` packages (not installed here; op.py:15-16, vulcan.py:56-57)
  2. `use_adapt_rtol = False` appended to the cfg (op.py:836,845 read it, no cfg defines it)
  3. the chosen cfg_examples/vulcan_cfg_<NAME>.py copied over vulcan_cfg.py, live plots off
  4. `make` in fastchem_vulcan/ (the only native code of the reference, setup only)
  5. `python make_chem_funs.py` (its trailing check_conserv crashes under numpy 2 at
     make_chem_funs.py:725 AFTER chem_funs.py is complete - ignored)
  6. HD209 + SNCHO_photo_network_2025: alias CH3CCH -> CH3C2H NASA9 file / compose row

Nothing under oracle/ is imported by the product package (vulcan_b200/); only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg use it.

usage: python oracle/stage_reference.py --config HD189 [--dest /tmp/vulcan_ref_HD189]
"""
import argparse
import os
import shutil
import subprocess
import sys

REF = os.environ.get("VULCAN_REFERENCE", "/root/reference")

MPL_STUB = '''"""stub: matplotlib is not installed; the reference only needs the import to succeed."""
class _Any(object):
    def __getattr__(self, name):
        return _Any()
    def __call__(self, *a, **k):
        return _Any()
    def __iter__(self):
        return iter(())
import sys as _sys, types as _types
for _n in ("pyplot", "legend", "colors", "cm", "animation"):
    _m = _types.ModuleType(__name__ + "." + _n)
    _m.__getattr__ = lambda name, _a=_Any(): _a
    _sys.modules[__name__ + "." + _n] = _m
    globals()[_n] = _m
def use(*a, **k):
    pass
'''

# per-config cfg edits: (source cfg, {name: python-literal replacement}, extra appended text)
CONFIGS = {
    "HD189": dict(src="cfg_examples/vulcan_cfg_HD189.py", edits={}, extra=""),
    "Jupiter": dict(src="cfg_examples/vulcan_cfg_Jupiter.py", edits={}, extra=""),
    # fixture-only variant: the shipped Jupiter cfg with condensation brought forward from 1e6 s to 10 s of model time and the
    # fix_species switch (op.py:862-893) from 1e8 s to 500 s, so that ~200 reference steps exercise conden / the relaxation operators,
    # reach the switch and the fixed-species rows of Ros2.solver (op.py:2896-2906, 2921-2924, 2960-2970).  (The shipped times are
    # out of reach of a fixture run: 3000 reference steps = 27 min of CPU only get to t = 6.6e5 s.)
    "JupiterFix": dict(src="cfg_examples/vulcan_cfg_Jupiter.py", edits={"start_conden_time": "10.", "stop_conden_time": "500."}, extra=""),
    # the same with fix_species_from_coldtrap_lev = False: the WHOLE column of every fixed species is frozen (op.py:2898-2899, 2962-2963)
    "JupiterFixAll": dict(src="cfg_examples/vulcan_cfg_Jupiter.py",
                          edits={"start_conden_time": "10.", "stop_conden_time": "500.", "fix_species_from_coldtrap_lev": "False"}, extra=""),
    # use_vm_mol variants ("under testing" in the reference, vulcan_cfg.py:77): upwind advective form of molecular diffusion,
    # diffdf_vm + lhs_jac_tot_vm (op.py:1599-1694, 2044-2119) and, with settling, diffdf_settling_vm + lhs_jac_settling_vm
    # (op.py:1794-1898, 2366-2444)
    "HD189vm": dict(src="cfg_examples/vulcan_cfg_HD189.py", edits={"use_vm_mol": "True"}, extra=""),
    # use_vz = True with a sign-changing vertical wind (the upwind advection terms of diffdf / lhs_jac_tot, op.py:1531-1595, 1996-2034: every
    # shipped cfg has use_vz = False, i.e. vz = 0).  vz_prof = 'file' reads a `vz` column from the atm file (build_atm.py:420-422): the
    # scratch copy gets atm_HD189_Kzz.txt + a vz column of +-150 cm/s (comparable to Kzz / H) - a test INPUT; the code is the unmodified reference
    "HD189vz": dict(src="cfg_examples/vulcan_cfg_HD189.py",
                    edits={"use_vz": "True", "vz_prof": "'file'", "atm_file": "'atm/atm_HD189_Kzz_vz_test.txt'"}, extra=""),
    # the same for the settling stencils: diffdf_settling / lhs_jac_settling (op.py:1696-1791, 2295-2364) and their use_vm_mol twins
    # (op.py:1794-1898, 2366-2444); amplitude 5 cm/s (Jupiter's Kzz / H is of that order)
    "JupiterVz": dict(src="cfg_examples/vulcan_cfg_Jupiter.py",
                      edits={"use_vz": "True", "vz_prof": "'file'", "atm_file": "'atm/Jupiter_deep_top_vz_test.txt'"}, extra=""),
    "JupiterVmVz": dict(src="cfg_examples/vulcan_cfg_Jupiter.py",
                        edits={"use_vz": "True", "vz_prof": "'file'", "atm_file": "'atm/Jupiter_deep_top_vz_test.txt'", "use_vm_mol": "True"}, extra=""),
    # the smallest shipped photochemical network (CHO_photo_network.txt, ni = 41: padded block size 48)
    "HD189cho": dict(src="cfg_examples/vulcan_cfg_HD189.py", edits={"network": "'thermo/CHO_photo_network.txt'", "atom_list": "['H', 'O', 'C']"}, extra=""),
    # thermochemistry only: a network without a photo section and use_photo = False (no compute_tau / flux / J calls at all)
    "HD189thermo": dict(src="cfg_examples/vulcan_cfg_HD189.py", edits={"network": "'thermo/NCHO_thermo_network.txt'", "use_photo": "False"}, extra=""),
    # use_moldiff = False: eddy diffusion only, diffdf_no_mol + lhs_jac_no_mol (op.py:1438-1494, 2122-2166)
    "HD189nomol": dict(src="cfg_examples/vulcan_cfg_HD189.py", edits={"use_moldiff": "False"}, extra=""),
    "JupiterVm": dict(src="cfg_examples/vulcan_cfg_Jupiter.py", edits={"use_vm_mol": "True"}, extra=""),
    # use_ion (op.py:2789-2820 compute_Jion, 2908-2911 / 2926 electron rows, 2998-3004 charge balance).  No shipped network has an
    # `# ionisation` section, so this fixture-only variant appends one to NCHO_photo_network.txt in the scratch copy (ION_TEST_*
    # below: H and H2 photo-ionisation with the shipped H / H2 ion cross sections and NASA-9 data of e, H_p, H2_p, H3_p plus four
    # ion-neutral / recombination reactions).  The network is a test input; the code that runs it is the unmodified reference.
    "HD189ion": dict(src="cfg_examples/vulcan_cfg_HD189.py",
                     edits={"use_ion": "True", "network": "'thermo/NCHO_photo_ion_test_network.txt'"}, extra=""),
    "HD209S": dict(
        src="cfg_examples/vulcan_cfg_HD189.py",
        edits={
            "atom_list": "['H', 'O', 'C', 'N', 'S']",
            "network": "'thermo/SNCHO_photo_network_2025.txt'",
            "atm_file": "'atm/atm_HD209_Kzz.txt'",
            "sflux_file": "'atm/stellar_flux/Gueymard_solar.txt'",
            "r_star": "1.203",
            "Rp": "1.38*7.1492E9",
            "orbit_radius": "0.04747",
            "gs": "936.",
            "out_name": "'HD209S.vul'",
        },
        extra="",
    ),
    "Earth": dict(
        src="cfg_examples/vulcan_cfg_Earth.py",
        edits={"network": "'thermo/NCHO_earth_photo_network.txt'"},
        extra="",  # filled in by stage() (species filtering needs the network parsed)
    ),
    # the Earth cfg AS SHIPPED: cfg_examples/vulcan_cfg_Earth.py:11 names SNCHO_full_photo_network.txt (ni = 99, nr = 1284 - the largest
    # network the reference ships a cfg for), sulphur boundary fluxes and const_mix included.  The cfg does not run as shipped: its
    # const_mix / atom_list name Ar, which SNCHO_full_photo_network.txt does not contain (build_atm.py:195 ValueError) -> Ar dropped
    "EarthS": dict(src="cfg_examples/vulcan_cfg_Earth.py",
                   edits={"atom_list": "['H', 'O', 'C', 'N', 'S']",
                          "const_mix": "{'N2':0.78, 'O2':0.20, 'H2O':1e-6, 'CO2':4E-4, 'SO2': 2e-10}"}, extra=""),
    # Earth + use_vm_mol: lhs_jac_settling_vm with the diffusion-limited escape term (diff_esc = ['H2','H'], op.py:2425-2431)
    "EarthVm": dict(
        src="cfg_examples/vulcan_cfg_Earth.py",
        edits={"network": "'thermo/NCHO_earth_photo_network.txt'", "use_vm_mol": "True"},
        extra="",
    ),
}

def write_vz_test_atm(dest, name="atm/atm_HD189_Kzz.txt", out_name="atm/atm_HD189_Kzz_vz_test.txt", amp=150.0):
    """atm/atm_HD189_Kzz.txt + a fourth column vz (cm/s): one and a half sine periods over the table, amplitude 150 cm/s, so that both
    signs and both sign changes of the upwind switch occur in the column."""
    import math
    src, dst = os.path.join(dest, name), os.path.join(dest, out_name)
    with open(src) as f:
        lines = f.read().splitlines()
    rows = [ln for ln in lines[2:] if ln.strip()]
    out = [lines[0] + "\t (cm/s)", lines[1] + "\t vz"]
    for q, ln in enumerate(rows):
        out.append("%s\t %.3E" % (ln, amp * math.sin(3.0 * math.pi * q / (len(rows) - 1))))
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")


ION_TEST_TWO_BODY = """\
9001 [ H_p + e -> H                       ]  4.00E-12    -0.640       0.0      ion test (radiative recombination)
9003 [ H2_p + e -> H + H                  ]  1.60E-08    -0.430       0.0      ion test
9005 [ H2_p + H2 -> H3_p + H              ]  2.00E-09     0.000       0.0      ion test
9007 [ H3_p + e -> H2 + H                 ]  4.00E-08    -0.500       0.0      ion test
"""
ION_TEST_IONISATION = """\
# ionisation
# id	# Reactions                                     sp		br_index #(starting from 1)
9101 [ H -> H_p + e                       ]             H       1
9103 [ H2 -> H2_p + e                     ]             H2      1
"""


def write_ion_test_network(dest):
    """NCHO_photo_network.txt + four ion reactions at the end of the two-body block + an `# ionisation` section (ids are
    renumbered by the reference's own make_chem_funs.py:35-110 when it rewrites the network file)."""
    src = os.path.join(dest, "thermo", "NCHO_photo_network.txt")
    out = os.path.join(dest, "thermo", "NCHO_photo_ion_test_network.txt")
    with open(src) as f:
        lines = f.readlines()
    k = [i for i, l in enumerate(lines) if l.startswith("# 3-body and Diss")][0]
    lines[k:k] = [ION_TEST_TWO_BODY, "\n"]
    if not lines[-1].endswith("\n"):
        lines[-1] += "\n"
    lines.append(ION_TEST_IONISATION)
    with open(out, "w") as f:
        f.writelines(lines)
    return out


COMMON_OFF = {
    "use_live_plot": "False",
    "use_live_flux": "False",
    "use_plot_end": "False",
    "use_plot_evo": "False",
    "use_save_movie": "False",
    "use_flux_movie": "False",
    "plot_TP": "False",
    "use_print_prog": "False",
}


def _edit_cfg(text, edits):
    """replace top-level `name = value` assignments (first occurrence at column 0)."""
    out = []
    done = set()
    for line in text.splitlines():
        stripped = line.split("#")[0]
        name = stripped.split("=")[0].strip() if "=" in stripped else None
        if name in edits and name not in done and line.startswith(name):
            out.append("%s = %s" % (name, edits[name]))
            done.add(name)
        else:
            out.append(line)
    for name in edits:
        if name not in done:
            out.append("%s = %s" % (name, edits[name]))
    return "\n".join(out) + "\n"


def _network_species(path):
    sp = []
    with open(path) as f:
        for line in f:
            if line.startswith("#") or "[" not in line:
                continue
            body = line.partition("[")[-1].rpartition("]")[0]
            for term in body.replace("->", "+").split("+"):
                t = term.strip()
                if not t or t == "M":
                    continue
                if "*" in t:
                    t = t.split("*")[1]
                if t not in sp:
                    sp.append(t)
    return sp


def stage(config, dest, run_codegen=True, quiet=True):
    cfg = CONFIGS[config]
    if os.path.exists(dest):
        shutil.rmtree(dest)
    ignore = shutil.ignore_patterns(".git", "demo", "plot_py", "*.gif")
    shutil.copytree(REF, dest, ignore=ignore)
    for d in ("output", "plot", "plot/movie", "fastchem_vulcan/obj", "fastchem_vulcan/output"):
        os.makedirs(os.path.join(dest, d), exist_ok=True)
    # make everything writable (the reference mount is read-only)
    subprocess.run(["chmod", "-R", "u+w", dest], check=True)

    # (1) stubs
    os.makedirs(os.path.join(dest, "_stubs", "matplotlib"), exist_ok=True)
    with open(os.path.join(dest, "_stubs", "matplotlib", "__init__.py"), "w") as f:
        f.write(MPL_STUB)

    # (6) CH3CCH alias (SNCHO_photo_network_2025 uses a name with no thermo data)
    compose = os.path.join(dest, "thermo", "all_compose.txt")
    with open(compose) as f:
        ctext = f.read()
    if not ctext.endswith("\n"):
        ctext += "\n"
    if "CH3CCH" not in ctext.split():
        for line in ctext.splitlines():
            if line.split() and line.split()[0] == "CH3C2H":
                ctext += line.replace("CH3C2H", "CH3CCH", 1) + "\n"
                break
        with open(compose, "w") as f:
            f.write(ctext)
    nasa = os.path.join(dest, "thermo", "NASA9")
    if not os.path.exists(os.path.join(nasa, "CH3CCH.txt")) and os.path.exists(os.path.join(nasa, "CH3C2H.txt")):
        shutil.copy(os.path.join(nasa, "CH3C2H.txt"), os.path.join(nasa, "CH3CCH.txt"))

    # (2)+(3) cfg
    with open(os.path.join(dest, cfg["src"])) as f:
        text = f.read()
    edits = dict(COMMON_OFF)
    edits.update(cfg["edits"])
    extra = cfg["extra"]
    if config in ("Earth", "EarthVm"):
        # BASELINE names NCHO_earth_photo_network.txt, which lacks the sulphur
        # species the shipped Earth cfg / BC file mention -> restrict to species in the network.
        sp = set(_network_species(os.path.join(dest, "thermo/NCHO_earth_photo_network.txt")))
        bc_src = os.path.join(dest, "atm/BC_bot_Earth.txt")
        bc_dst = os.path.join(dest, "atm/BC_bot_Earth_NCHO.txt")
        with open(bc_src) as f, open(bc_dst, "w") as g:
            for line in f:
                tok = line.split()
                if line.startswith("#") or not tok or tok[0] in sp:
                    g.write(line)
        edits.update({
            "atom_list": "['H', 'O', 'C', 'N']",
            "bot_BC_flux_file": "'atm/BC_bot_Earth_NCHO.txt'",
            "const_mix": "{'N2':0.78, 'O2':0.20, 'H2O':1e-6, 'CO2':4E-4, 'Ar':9.34e-3}" if "Ar" in sp
            else "{'N2':0.78, 'O2':0.20, 'H2O':1e-6, 'CO2':4E-4}",
            "use_relax": "['H2O']",
            "condense_sp": "['H2O']",
            "non_gas_sp": "['H2O_l_s']",
            "fix_species": "['H2O','H2O_l_s']",
            "r_p": "{'H2O_l_s': 0.01}",
            "rho_p": "{'H2O_l_s': 0.9}",
            "remove_list": "[]",
        })
    if config == "HD189ion":
        write_ion_test_network(dest)
    if config == "HD189vz":
        write_vz_test_atm(dest)
    if config in ("JupiterVz", "JupiterVmVz"):
        write_vz_test_atm(dest, "atm/Jupiter_deep_top.txt", "atm/Jupiter_deep_top_vz_test.txt", amp=5.0)
    text = _edit_cfg(text, edits)
    text += "\n# --- shims added by oracle/stage_reference.py (non-numerical) ---\nuse_adapt_rtol = False\n" + extra
    with open(os.path.join(dest, "vulcan_cfg.py"), "w") as f:
        f.write(text)

    out = subprocess.DEVNULL if quiet else None
    # (4) FastChem
    subprocess.run(["make"], cwd=os.path.join(dest, "fastchem_vulcan"), check=True, stdout=out, stderr=out)
    # (5) chem_funs.py
    if run_codegen:
        env = dict(os.environ, PYTHONHASHSEED="0", PYTHONPATH=os.path.join(dest, "_stubs"), OMP_NUM_THREADS="1")
        subprocess.run([sys.executable, "make_chem_funs.py"], cwd=dest, env=env, stdout=out, stderr=out)
        if not os.path.exists(os.path.join(dest, "chem_funs.py")):
            raise RuntimeError("make_chem_funs.py did not produce chem_funs.py")
        with open(os.path.join(dest, "chem_funs.py")) as f:
            if "def neg_symjac" not in f.read():
                raise RuntimeError("chem_funs.py incomplete (no neg_symjac)")
    return dest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="HD189", choices=sorted(CONFIGS))
    ap.add_argument("--dest", default=None)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    d = a.dest or "/tmp/vulcan_ref_%s" % a.config
    stage(a.config, d, quiet=not a.verbose)
    print(d)
