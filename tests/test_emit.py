"""The network -> CUDA generator (vulcan_b200/emit.py, the counterpart of make_chem_funs.py:113-430) without a GPU: the emitted straight-line
kernel body is translated statement by statement to numpy (same IEEE operations, no contraction - the unit is compiled -fmad=false) and
evaluated on the reference's recorded states: chemdf must be BIT-IDENTICAL to the reference's generated chem_funs.chemdf.  The device run of
the same source is tests/test_gpu_parity.py::test_rhs_bit_exact[1] / test_rhs_default_kernel."""
import re

import numpy as np
import pytest

from helpers import Case, have
from vulcan_b200 import emit

CASES = [p for p in [("HD189", 0), ("HD189", 300), ("Jupiter", 30), ("Earth", 30), ("HD209S", 30), ("EarthS", 100), ("HD189cho", 30)]
         if have(p[0], "step%04d.npz" % p[1])]


def _statements(line):
    s = line.strip().replace("const double ", "").replace("double ", "")
    return [x.strip() for x in s.split(";") if x.strip()]


def _evaluate(src, y, M, k):
    """numpy interpretation of the emitted chemdf kernel(s): y [nz, ni], M [nz], k [nz, nr+1] -> chemdf [nz, ni]"""
    nz, ni = y.shape
    yx = np.concatenate([y, M[:, None]], axis=1)
    out = np.full((nz, ni), np.nan)
    ns, active = {}, False
    for line in src.splitlines():
        s = line.strip()
        if s.startswith("__global__"):
            ns, active = {}, "chemdf_" in s
            continue
        if not active or not s or s in ("{", "}") or s.startswith("}") or s.startswith(("//", "#", "VK_EMIT")):
            continue
        if s.startswith(("int launch", "const size_t", "return", "const EmitRegistrar")):
            active = False
            continue
        for stmt in _statements(s):
            if re.match(r"f\d+ = 0\.0(, f\d+ = 0\.0)*$", stmt):
                for part in stmt.split(", "):
                    ns[part.split(" = ")[0]] = np.zeros(nz)
                continue
            m = re.match(r"F\((\d+)\) = f(\d+)$", stmt)
            if m:
                out[:, int(m.group(1))] = ns["f" + m.group(2)]
                continue
            stmt = re.sub(r"KG?\((\d+)\)", r"k[:, \1]", stmt)
            stmt = re.sub(r"Y\((\d+)\)", r"yx[:, \1]", stmt)
            exec(stmt, {"k": k, "yx": yx, "pow": np.power}, ns)
    return out


class _Rows(dict):
    def __init__(self, nz):
        dict.__init__(self)
        self.nz = nz

    def __missing__(self, key):
        return np.zeros(self.nz)


def _evaluate_jac(src, y, M, k, nip):
    """numpy interpretation of the emitted Jacobian kernel: -> -J [nz, nip, nip] (rows >= ni stay zero: lhs_diag_kernel writes them)"""
    nz, ni = y.shape
    J = np.zeros((nz, nip, nip))
    rT = _Rows(nz)
    for s_ in range(ni):
        rT[s_] = y[:, s_].copy()                       # the staged y row (S(s) = rT[s])
    ns, active = {"yM": M}, False
    for line in src.splitlines():
        s = line.strip()
        if s.startswith("__global__"):
            active = "negjac_" in s
            continue
        if not active:
            continue
        if s.startswith("int launch"):
            break
        m = re.match(r"VK_EMITJ_BEGIN\((\d+), (\d+)\)", s)
        if m:
            for r in range(int(m.group(2))):
                rT[r] = np.zeros(nz)
            continue
        m = re.match(r"VK_EMITJ_ROW\((\d+), (\d+), (\d+)\)", s)
        if m:
            for t in range(nip):
                J[:, int(m.group(1)), t] = rT[t]
            continue
        s = s.split("//")[0].strip("{} ")
        if not s or s.startswith(("#", "VK_EMIT")):
            continue
        for stmt in _statements(s):
            if re.match(r"(t|a\d+_\d+(, a\d+_\d+)*)$", stmt):          # declarations without initialiser
                continue
            if re.match(r"[yc]\d+ = S\(\d+\)(, [yc]\d+ = S\(\d+\))*$", stmt):
                for part in stmt.split(", "):
                    name, src_ = part.split(" = ")
                    ns[name] = rT[int(src_[2:-1])].copy()
                continue
            stmt = re.sub(r"KG?\((\d+)\)", r"k[:, \1]", stmt)
            stmt = re.sub(r"R\((\d+)\)", r"rT[\1]", stmt)
            stmt = re.sub(r"= 0\.0$", "= zeros()", stmt)
            exec(stmt, {"k": k, "rT": rT, "zeros": lambda: np.zeros(nz)}, ns)
    return J


@pytest.mark.parametrize("tag,step", CASES)
def test_emitted_chemdf_is_bit_identical_to_the_reference(tag, step):
    c = Case(tag, step)
    src, h = emit.emit_chemdf(c.net.tables(), tag)
    chem = _evaluate(src, c.y, c.st["M"], c.k)
    assert not np.isnan(chem).any()                      # every species is stored by exactly one pass
    assert np.array_equal(chem, c.fx["chemdf"])
    assert ("launch_%016x" % h) in src and ("0x%016xull" % h) in src


@pytest.mark.parametrize("tag,step", CASES)
def test_emitted_jacobian_is_bit_identical_to_the_oracle(tag, step):
    """the emitted Jacobian kernel sums the terms of every entry left to right in the order of the oracle's vko_chemjac (the analytic
    restatement of chem_funs.symjac, pinned against the reference's sympy Jacobian in test_oracle_vs_reference.py): same bits"""
    from oracle import Oracle
    c = Case(tag, step)
    src, h = emit.emit_chemdf(c.net.tables(), tag)
    assert ("negjac_%016x" % h) in src
    nip = emit.pad_block(c.ni)
    negJ = _evaluate_jac(src, c.y, c.st["M"], c.k, nip)
    Jo = Oracle(c.net).chemjac(c.y, c.st["M"], c.k)
    assert np.array_equal(negJ[:, :c.ni, :c.ni], -Jo)
    assert not negJ[:, c.ni:, :].any() and not negJ[:, :, c.ni:].any()


def test_hash_and_registry_of_the_baseline_networks():
    for tag in ("HD189", "Jupiter", "Earth", "HD209S"):
        if not have(tag, "step0000.npz"):
            continue
        net = Case(tag, 0).net
        t = net.tables()
        assert emit.table_hash(t) == emit._fnv_fast(t)
        assert emit.has_kernel(net), "%s: tables not registered under vulcan_b200/networks (python -m vulcan_b200.emit <network file>)" % tag


def test_dynamic_rows_are_read_per_column():
    """the k indices a run rewrites per column (photolysis / ionisation / condensation sections) are emitted as KG(i) - read from the
    thread's own column - and every other index as K(i) - the block's shared copy of the row"""
    c = Case("HD189", 0)
    t = c.net.tables()
    src, h = emit.emit_chemdf(t, "HD189")
    dyn = set(int(x) for x in t["dyn_k"])
    photo = {rid + d for _, _, rid in c.net.photo_table() for d in (0, 1)}
    assert photo <= dyn and len(photo) > 0
    kg = {int(x) for x in re.findall(r"KG\((\d+)\)", src)}
    ks = {int(x) for x in re.findall(r"(?<!G)K\((\d+)\)", src.replace("KG(", "G("))}
    assert kg and kg <= dyn and not (ks & dyn)
    assert ("dyn_%016x[] = {" % h) in src


def test_terms_out_of_reaction_order_are_refused():
    t = dict(Case("HD189", 0).net.tables())
    pair = np.array(t["rhs_pair"]).copy()
    p0, p1 = t["rhs_ptr"][0], t["rhs_ptr"][1]
    assert p1 - p0 >= 2
    pair[p0], pair[p1 - 1] = pair[p1 - 1], pair[p0]
    t["rhs_pair"] = pair
    with pytest.raises(ValueError):
        emit.emit_chemdf(t, "broken")


def test_registered_tables_round_trip(tmp_path, monkeypatch):
    """python -m vulcan_b200.emit <network>: the tables written under networks/ are the ones the emitter and the hash read back"""
    c = Case("HD189", 0)
    monkeypatch.setattr(emit, "NET_DIR", str(tmp_path))
    out = emit.register_network(c.net, "roundtrip")
    z = np.load(out)
    t = c.net.tables()
    for key in emit.TABLE_KEYS:
        assert np.array_equal(np.asarray(z[key]), np.asarray(t[key])), key
    assert emit.has_kernel(c.net)          # found by hash in the (temporary) registry directory
