"""Network compiler: VULCAN reaction-network text file -> flat integer/float tables.

This is the front half of the network-to-CUDA generator that replaces the reference's
sympy pipeline (make_chem_funs.py -> chem_funs.py).  It reproduces the *semantics* of the
reference's parser, not its code:

* reaction ids: every reaction line, top to bottom, gets the next odd id 1,3,5,...; id+1 is its
  reverse (make_chem_funs.py:17-110 renumbers the file the same way; op.py:89-271 reads it).
* species index = order of first appearance scanning each reaction left to right, reactants
  then products, `M` skipped (make_chem_funs.py:160-184).
* rate of progress  v_j = k[j]*Π reactants  -  k[j+1]*Π products, factors multiplied in the order
  they are written, `M` included where written (make_chem_funs.py:221-256).
* dy_s/dt = Σ over reactions in file order, for every occurrence of s as reactant `-stoi*v_j`,
  then as product `+stoi*v_j` (make_chem_funs.py:258-285) - the summation order is kept because
  P-L cancels catastrophically near equilibrium (SURVEY.md §7 "hard parts").
* section markers (`# 3-body`, `# 3-body reactions without high-pressure rates`, `# special`,
  `# condensation`, `# radiative`, `# photo`, `# ionisation`, `# reverse stops`) follow op.py:89-136.

The Jacobian tables are derived analytically here (product rule on the monomials) instead of by
symbolic differentiation of generated source (make_chem_funs.py:653-717).

Outputs are plain numpy arrays (see `Network.tables()`), consumed by
  - vulcan_b200/_abi.py     (uploaded to the GPU through vk_network_create)
  - the CPU checker under tests/ (test infrastructure only, never imported from here)
"""
import json
import re

import numpy as np

M_SLOT = -1          # factor index standing for the third body `M` in reaction term lists

SECTION_2BODY, SECTION_3BODY, SECTION_3BODY_K0, SECTION_SPECIAL, SECTION_CONDEN, SECTION_RECOMB, \
    SECTION_PHOTO, SECTION_ION = range(8)
SECTION_NAMES = ["two-body", "three-body", "three-body-k0", "special", "condensation", "radiative", "photo", "ion"]


class Reaction(object):
    __slots__ = ("id", "text", "reac", "prod", "section", "rate_cols", "photo_sp", "branch", "tag")

    def __init__(self, rid, text, reac, prod, section, rate_cols, photo_sp=None, branch=None, tag=None):
        self.id, self.text, self.reac, self.prod = rid, text, reac, prod
        self.section, self.rate_cols, self.photo_sp, self.branch, self.tag = section, rate_cols, photo_sp, branch, tag


def _parse_side(side):
    """'H + 2*OH + M' -> [('H',1), ('OH',2), ('M',1)] in written order."""
    out = []
    for term in side.split():
        if term == "+":
            continue
        if "*" in term:
            n, name = term.split("*", 1)
            out.append((name, int(n)))
        else:
            out.append((term, 1))
    return out


class Network(object):
    """Parsed reaction network + derived evaluation tables."""

    def __init__(self, reactions, species, stop_rev_indx, name=""):
        self.reactions = reactions
        self.species = list(species)
        self.ni = len(self.species)
        self.nr = 2 * len(reactions)
        self.stop_rev_indx = stop_rev_indx
        self.name = name
        self._tables = None

    # ------------------------------------------------------------------ parsing
    @classmethod
    def from_text(cls, text, name=""):
        reactions, species, sp_index = [], [], {}
        section = SECTION_2BODY
        rid = 1
        stop_rev = None
        for line in text.splitlines():
            if line.startswith("# 3-body reactions without high-pressure rates"):
                section = SECTION_3BODY_K0
            elif line.startswith("# 3-body"):
                section = SECTION_3BODY
            elif line.startswith("# special"):
                section = SECTION_SPECIAL
            elif line.startswith("# condensation"):
                section = SECTION_CONDEN
            elif line.startswith("# radiative"):
                section = SECTION_RECOMB
            elif line.startswith("# photo"):
                section = SECTION_PHOTO
            elif line.startswith("# ionisation"):
                section = SECTION_ION
            elif line.startswith("# reverse stops"):
                stop_rev = rid
            if line.startswith("#") or not line.strip() or "[" not in line:
                continue
            body = line.partition("[")[-1].rpartition("]")[0].strip()
            cols = line.partition("]")[-1].split()
            lhs, _, rhs = body.partition("->")
            reac, prod = _parse_side(lhs), _parse_side(rhs)
            for nm, _n in reac + prod:
                if nm != "M" and nm not in sp_index:
                    sp_index[nm] = len(species)
                    species.append(nm)
            rate_cols, photo_sp, branch, tag = [], None, None, None
            if section in (SECTION_2BODY, SECTION_3BODY, SECTION_3BODY_K0, SECTION_RECOMB):
                nnum = 6 if section == SECTION_3BODY else 3
                rate_cols = [float(c) for c in cols[:nnum]]
                if cols and cols[-1] in ("He", "ex1"):
                    tag = cols[-1]
            elif section in (SECTION_PHOTO, SECTION_ION):
                photo_sp, branch = cols[0], int(cols[1])
            reactions.append(Reaction(rid, body,
                                      [(M_SLOT if nm == "M" else sp_index[nm], n) for nm, n in reac],
                                      [(M_SLOT if nm == "M" else sp_index[nm], n) for nm, n in prod],
                                      section, rate_cols, photo_sp, branch, tag))
            rid += 2
        if stop_rev is None:
            stop_rev = rid
        return cls(reactions, species, stop_rev, name)

    @classmethod
    def from_file(cls, path):
        with open(path) as f:
            return cls.from_text(f.read(), name=path.rsplit("/", 1)[-1])

    # ------------------------------------------------------------------ tables
    def tables(self):
        """Flat tables (all numpy, C-order):

        rate_fac  int32 [nr+1, MAXF]   ordered factor slots of rate term i (k index i): species index,
                                       ni = third body M, ni+1 = the constant 1.0 (padding)
        rate_pow  int32 [nr+1, MAXF]   exponent of each factor (1 unless the network writes `n*X`)
        rhs_ptr   int32 [ni+1]         CSR over species: contributions rhs_pair/rhs_coef in reference order
        rhs_pair  int32 [nnz]          forward reaction id j (odd); contribution = coef * (rate[j] - rate[j+1])
        rhs_coef  f64   [nnz]
        jac_ptr   int32 [nent+1]       CSR over non-zero Jacobian entries
        jac_row, jac_col int32 [nent]  (species s, species t):  d(dy_s/dt)/dy_t
        jac_k     int32 [nterm]        k index of the term
        jac_coef  f64   [nterm]        signed integer coefficient (stoichiometry x exponent)
        jac_fac   int32 [nterm, MAXF-1+?] remaining factor slots after differentiation (padded with ni+1)
        """
        if self._tables is not None:
            return self._tables
        ni, nr = self.ni, self.nr
        ONE = ni + 1
        maxf = 1
        for r in self.reactions:
            maxf = max(maxf, len(r.reac), len(r.prod))
        maxf = max(maxf, 4)
        rate_fac = np.full((nr + 1, maxf), ONE, dtype=np.int32)
        rate_pow = np.ones((nr + 1, maxf), dtype=np.int32)
        rhs_lists = [[] for _ in range(ni)]
        # Jacobian terms keyed by (s, t)
        jac = {}

        def add_terms(side, kidx, stoich_of_species_sign):
            """side: ordered [(slot, n)], kidx: k index of this rate term;
            stoich_of_species_sign: list of (species s, signed coef) telling how +rate(kidx) enters dy_s/dt."""
            # merged exponents per species (the monomial is k * M^m * Π y_t^{a_t})
            expo = {}
            m_pow = 0
            for slot, n in side:
                if slot == M_SLOT:
                    m_pow += n
                else:
                    expo[slot] = expo.get(slot, 0) + n
            for t, a_t in expo.items():
                # d/dy_t: a_t * y_t^(a_t-1) * Π_{f != t} y_f^{a_f} * M^m
                fac = []
                for f, a_f in sorted(expo.items()):
                    e = a_f - 1 if f == t else a_f
                    fac += [f] * e
                fac += [ni] * m_pow
                for s, c in stoich_of_species_sign:
                    jac.setdefault((s, t), []).append((kidx, float(c * a_t), tuple(fac)))

        for r in self.reactions:
            j = r.id
            for q, (slot, n) in enumerate(r.reac):
                rate_fac[j, q] = ni if slot == M_SLOT else slot
                rate_pow[j, q] = n
            for q, (slot, n) in enumerate(r.prod):
                rate_fac[j + 1, q] = ni if slot == M_SLOT else slot
                rate_pow[j + 1, q] = n
            contrib = []          # (species, signed coef) of v_j = fwd - rev in dy/dt, reference order
            for slot, n in r.reac:
                if slot != M_SLOT:
                    rhs_lists[slot].append((j, -float(n)))
                    contrib.append((slot, -n))
            for slot, n in r.prod:
                if slot != M_SLOT:
                    rhs_lists[slot].append((j, float(n)))
                    contrib.append((slot, n))
            add_terms(r.reac, j, contrib)
            add_terms(r.prod, j + 1, [(s, -c) for s, c in contrib])

        rhs_ptr = np.zeros(ni + 1, dtype=np.int32)
        rhs_pair, rhs_coef = [], []
        for s in range(ni):
            for j, c in rhs_lists[s]:
                rhs_pair.append(j)
                rhs_coef.append(c)
            rhs_ptr[s + 1] = len(rhs_pair)

        # merge duplicate (kidx, factors) terms inside an entry (e.g. H + H -> ...: two identical contributions)
        keys = sorted(jac.keys())
        jac_ptr = [0]
        jac_row, jac_col, jac_k, jac_coef, jac_facs = [], [], [], [], []
        maxjf = 1
        for (s, t) in keys:
            merged = {}
            order = []
            for kidx, c, fac in jac[(s, t)]:
                key = (kidx, fac)
                if key not in merged:
                    merged[key] = 0.0
                    order.append(key)
                merged[key] += c
            n_added = 0
            for key in order:
                c = merged[key]
                if c == 0.0:
                    continue          # e.g. A + B -> A + C : species A's net stoichiometry is zero
                jac_k.append(key[0])
                jac_coef.append(c)
                jac_facs.append(key[1])
                maxjf = max(maxjf, len(key[1]))
                n_added += 1
            if n_added:
                jac_row.append(s)
                jac_col.append(t)
                jac_ptr.append(len(jac_k))
        maxjf = max(maxjf, 3)
        jac_fac = np.full((len(jac_k), maxjf), ONE, dtype=np.int32)
        for q, fac in enumerate(jac_facs):
            jac_fac[q, :len(fac)] = fac

        self._tables = dict(
            ni=ni, nr=nr, maxf=maxf, maxjf=maxjf,
            rate_fac=rate_fac, rate_pow=rate_pow,
            rhs_ptr=rhs_ptr, rhs_pair=np.array(rhs_pair, dtype=np.int32), rhs_coef=np.array(rhs_coef, dtype=np.float64),
            jac_ptr=np.array(jac_ptr, dtype=np.int32), jac_row=np.array(jac_row, dtype=np.int32),
            jac_col=np.array(jac_col, dtype=np.int32), jac_k=np.array(jac_k, dtype=np.int32),
            jac_coef=np.array(jac_coef, dtype=np.float64), jac_fac=jac_fac,
            # k indices (forward and reverse) that the run itself rewrites per column: photolysis / ionisation rates (compute_J, compute_Jion)
            # and condensation growth rates (Integration.conden) - everything else is set once from the T-P profile
            dyn_k=np.array(sorted({r.id + d for r in self.reactions if r.section in (SECTION_PHOTO, SECTION_ION, SECTION_CONDEN)
                                   for d in (0, 1)}), dtype=np.int32),
        )
        return self._tables

    # ------------------------------------------------------------------ misc
    def photo_table(self):
        """[(species name, branch, forward reaction id)] for the photodissociation section (op.py:235-250)."""
        return [(r.photo_sp, r.branch, r.id) for r in self.reactions if r.section == SECTION_PHOTO]

    def ion_table(self):
        return [(r.photo_sp, r.branch, r.id) for r in self.reactions if r.section == SECTION_ION]

    def check_conservation(self, compo):
        """element balance of every reaction; compo: {species: {atom: n}} (make_chem_funs.py:719-747)."""
        bad = []
        for r in self.reactions:
            tot = {}
            for sign, side in ((-1, r.reac), (1, r.prod)):
                for slot, n in side:
                    if slot == M_SLOT:
                        continue
                    for atom, cnt in compo[self.species[slot]].items():
                        tot[atom] = tot.get(atom, 0) + sign * n * cnt
            if any(v != 0 for v in tot.values()):
                bad.append(r.id)
        return bad

    def to_json(self):
        return json.dumps(dict(
            name=self.name, species=self.species, stop_rev_indx=self.stop_rev_indx,
            reactions=[dict(id=r.id, text=r.text, reac=r.reac, prod=r.prod, section=r.section,
                            rate_cols=r.rate_cols, photo_sp=r.photo_sp, branch=r.branch, tag=r.tag)
                       for r in self.reactions]))

    @classmethod
    def from_json(cls, s):
        d = json.loads(s)
        rs = [Reaction(r["id"], r["text"], [tuple(x) for x in r["reac"]], [tuple(x) for x in r["prod"]],
                       r["section"], r["rate_cols"], r["photo_sp"], r["branch"], r["tag"]) for r in d["reactions"]]
        return cls(rs, d["species"], d["stop_rev_indx"], d.get("name", ""))
