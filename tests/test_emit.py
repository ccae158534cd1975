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


def _evaluate(src, y, M, k):
    """numpy interpretation of the emitted kernel(s): y [nz, ni], M [nz], k [nz, nr+1] -> chemdf [nz, ni]"""
    nz, ni = y.shape
    yx = np.concatenate([y, M[:, None]], axis=1)
    out = np.full((nz, ni), np.nan)
    ns, lo = {}, 0
    for line in src.splitlines():
        s = line.strip()
        if s.startswith("__global__"):
            ns = {}
            continue
        if not s or s in ("{", "}") or s.startswith("}") or s.startswith(("//", "#", "namespace", "VK_EMIT_", "int launch", "const size_t", "return", "const EmitRegistrar", "}}")):
            continue
        s = s.replace("const double ", "").replace("double ", "")
        if re.match(r"f\d+ = 0\.0, ", s):
            for part in s.rstrip(";").split(", "):
                name, val = part.split(" = ")
                ns[name] = np.zeros(nz)
            continue
        for stmt in [x for x in s.split(";") if x.strip()]:
            stmt = stmt.strip()
            m = re.match(r"F\((\d+)\) = f(\d+)$", stmt)
            if m:
                out[:, int(m.group(1))] = ns["f" + m.group(2)]
                continue
            stmt = re.sub(r"K\((\d+)\)", lambda q: "k[:, %d]" % (lo + int(q.group(1))), stmt)
            stmt = re.sub(r"Y\((\d+)\)", r"yx[:, \1]", stmt)
            exec(stmt, {"k": k, "yx": yx, "pow": np.power}, ns)
    return out


@pytest.mark.parametrize("tag,step", CASES)
def test_emitted_chemdf_is_bit_identical_to_the_reference(tag, step):
    c = Case(tag, step)
    src, h = emit.emit_chemdf(c.net.tables(), tag)
    chem = _evaluate(src, c.y, c.st["M"], c.k)
    assert not np.isnan(chem).any()                      # every species is stored by exactly one pass
    assert np.array_equal(chem, c.fx["chemdf"])
    assert ("launch_%016x" % h) in src and ("0x%016xull" % h) in src


def test_hash_and_registry_of_the_baseline_networks():
    for tag in ("HD189", "Jupiter", "Earth", "HD209S"):
        if not have(tag, "step0000.npz"):
            continue
        net = Case(tag, 0).net
        t = net.tables()
        assert emit.table_hash(t) == emit._fnv_fast(t)
        assert emit.has_kernel(net), "%s: tables not registered under vulcan_b200/networks (python -m vulcan_b200.emit <network file>)" % tag


def test_terms_out_of_reaction_order_are_refused():
    t = dict(Case("HD189", 0).net.tables())
    pair = np.array(t["rhs_pair"]).copy()
    p0, p1 = t["rhs_ptr"][0], t["rhs_ptr"][1]
    assert p1 - p0 >= 2
    pair[p0], pair[p1 - 1] = pair[p1 - 1], pair[p0]
    t["rhs_pair"] = pair
    with pytest.raises(ValueError):
        emit.emit_chemdf(t, "broken")
