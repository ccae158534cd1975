"""Network -> CUDA code generator: the B200 counterpart of the reference's `make_chem_funs.py` (make_chem_funs.py:113-430 writes
`chem_funs.chemdf` as one Python expression per species; here the same network becomes one straight-line CUDA kernel).

    python -m vulcan_b200.emit                      regenerate vulcan_b200/csrc/gen/*.cu from vulcan_b200/networks/*.npz
    python -m vulcan_b200.emit path/to/network.txt  register a network file of the reference's grammar (writes networks/<name>.npz,
                                                    then `python -m vulcan_b200.build` compiles it into the library)

What is emitted (chemdf_<hash>): one THREAD per (column, layer).  The thread keeps dy_s/dt of every species in a named register and walks
the reaction pairs in file order:   r_f = k[2p+1] y_a y_b ..;  r_r = k[2p+2] y_c ..;  v = r_f - r_r;  f_a -= v; f_b -= v; f_c += v; ...
Every species index, stoichiometric coefficient and k index is an immediate - no descriptor loads, no decode, no indirect addressing.
Because each species' terms arrive in increasing reaction order, every f_s is summed LEFT TO RIGHT IN THE REFERENCE'S ORDER
(make_chem_funs.py:258-285): the result is bit-identical to chem_funs.chemdf, like the table-driven reference-order kernel, while the
dependent add chain of a species (H: 170 terms) is interleaved with ~1000 independent instructions of the other species.
A block takes ONE layer of 128 consecutive columns of a batch that shares its rate coefficients: the k row of the layer sits in shared
memory and is read as a broadcast with immediate offsets, y of the 128 columns is transposed in shared memory (conflict-free `ys[s][tid]`).
Runtime pieces: csrc/vk_emit_rt.cuh.  Batches with per-column k and single columns stay on the table-driven kernels.

The library looks an uploaded network up by the FNV-1a hash of its tables (vk_network_create); networks without an emitted kernel run on
the table-driven kernels (vk_chem.cu), so a drop-in still accepts any network file at run time.
"""
import glob
import json
import os
import sys

import numpy as np

from .network import Network

HERE = os.path.dirname(os.path.abspath(__file__))
NET_DIR = os.path.join(HERE, "networks")
GEN_DIR = os.path.join(HERE, "csrc", "gen")
JAC_TB = int(os.environ.get('VK_EMIT_JAC_TB', 128))     # emit_negjac: threads (columns) per block
JAC_BLOCKS = int(os.environ.get('VK_EMIT_JAC_BLOCKS', 2)) # emit_negjac: resident blocks per SM aimed at
JAC_COLD_MAX = 32     # emit_negjac: at most this many species are read from shared memory instead of registers
ROW_ACC = 24          # emit_negjac: entries of a row accumulated in registers before the row buffer is touched       # emit_negjac: forbid common sub-expressions across Jacobian rows (see there)
MAX_ACC = 84          # dy/dt accumulators held in registers per pass (2 registers each); larger networks take several passes


def table_hash(t):
    """FNV-1a (64 bit) over the tables that define chemdf; the same bytes are hashed by vk_network_create (vk_api.cu)"""
    h = 0xcbf29ce484222325
    parts = [np.array([t["ni"], t["nr"], t["maxf"]], dtype=np.int32), np.ascontiguousarray(t["rate_fac"], dtype=np.int32),
             np.ascontiguousarray(t["rate_pow"], dtype=np.int32), np.ascontiguousarray(t["rhs_ptr"], dtype=np.int32),
             np.ascontiguousarray(t["rhs_pair"], dtype=np.int32), np.ascontiguousarray(t["rhs_coef"], dtype=np.float64)]
    for a in parts:
        for b in a.tobytes():
            h = ((h ^ b) * 0x100000001b3) & 0xffffffffffffffff
    return h


def _fnv_fast(t):
    """same hash, vectorised enough for 100 kB of tables"""
    data = b"".join([np.array([t["ni"], t["nr"], t["maxf"]], dtype=np.int32).tobytes(), np.ascontiguousarray(t["rate_fac"], dtype=np.int32).tobytes(),
                     np.ascontiguousarray(t["rate_pow"], dtype=np.int32).tobytes(), np.ascontiguousarray(t["rhs_ptr"], dtype=np.int32).tobytes(),
                     np.ascontiguousarray(t["rhs_pair"], dtype=np.int32).tobytes(), np.ascontiguousarray(t["rhs_coef"], dtype=np.float64).tobytes()])
    h = 0xcbf29ce484222325
    for b in data:
        h = ((h ^ b) * 0x100000001b3) & 0xffffffffffffffff
    return h


def _coef_lit(c):
    a = abs(c)
    return "%d.0" % int(a) if a == int(a) else repr(float(a))


def pad_block(ni):
    """padded block size the factor / solve kernels are instantiated for (vk_api.cu: pad_block)"""
    for n in (48, 72, 96, 120):
        if ni <= n:
            return n
    raise ValueError("ni > 120")


def emit_negjac(t, h, w):
    """the chemical Jacobian (chem_funs.neg_symjac, make_chem_funs.py:653-717; analytic terms of Network.tables()) as straight-line code:
    one thread per (column, layer) forms the entries ROW BY ROW - every entry the left-to-right sum of its terms coef * k * prod(y), exactly
    the order of the oracle's vko_chemjac - into its own row buffer in shared memory; the finished dense row (NIP doubles, zeros included)
    leaves as one asynchronous bulk copy issued by the thread, and the next row is formed while it drains.  y of the species the terms read
    most lives in registers, the rest behind the row buffer.  Returns False when the tables carry no Jacobian."""
    if "jac_ptr" not in t or t.get("jac_ptr") is None:
        return False
    ni, nr, mjf = t["ni"], t["nr"], int(t["maxjf"])
    nip = pad_block(ni)
    jp, jr, jc = np.asarray(t["jac_ptr"]), np.asarray(t["jac_row"]), np.asarray(t["jac_col"])
    jk, jcoef, jfac = np.asarray(t["jac_k"]), np.asarray(t["jac_coef"]), np.asarray(t["jac_fac"]).reshape(-1, mjf)
    rows = [[] for _ in range(ni)]
    for e in range(len(jr)):
        rows[int(jr[e])].append(e)
    dyn = set(int(x) for x in np.asarray(t.get("dyn_k", [])).ravel())
    # hot species live in registers, cold ones (least used by the Jacobian terms) behind the thread's row buffer in shared memory; the row
    # stride is chosen so that JAC_TB threads x JAC_BLOCKS blocks fit the shared memory of an SM
    use = np.bincount(jfac[jfac < ni].ravel(), minlength=ni)
    ks_bytes = 8 * (nr + 2)
    budget = (226 * 1024) // JAC_BLOCKS - ks_bytes
    if JAC_TB * 8 * (nip + 2) > budget:
        budget = 226 * 1024 - ks_bytes
    rld = min((budget // (JAC_TB * 8)) & ~1, nip + 2 + ni)
    n_cold = max(0, min(rld - (nip + 2), ni - 8, JAC_COLD_MAX))
    rld = nip + 2 + n_cold + (n_cold & 1)
    cold = {int(sp): q for q, sp in enumerate(np.argsort(use, kind="stable")[:n_cold])}
    cb = nip + 2
    yname = lambda sl: "yM" if sl == ni else ("rT[%d]" % (cb + cold[sl]) if sl in cold else "y%d" % sl)
    blocks = max(1, min(JAC_BLOCKS, (226 * 1024) // (JAC_TB * 8 * rld + ks_bytes)))
    w("// %d of %d species in registers, %d in shared memory (%.1f %% of the factor reads); row stride %d doubles, %d block(s) per SM" % (
        ni - n_cold, ni, n_cold, 100.0 * sum(use[sp] for sp in cold) / max(1, use.sum()), rld, blocks))
    w("__global__ void __launch_bounds__(%d, %d) negjac_%016x(EmitJacArgs a)" % (JAC_TB, blocks, h))
    w("{")
    w("    VK_EMITJ_PROLOGUE(%d, %d, %d, %d, %d)" % (ni, nr, nip, rld, JAC_TB))
    w("    const double " + ", ".join("y%d = S(%d)" % (s, s) for s in range(ni) if s not in cold) + ";")
    if cold:
        w("    { const double " + ", ".join("c%d = S(%d)" % (q, sp) for sp, q in sorted(cold.items(), key=lambda x: x[1])) + ";")
        w("      " + " ".join("rT[%d] = c%d;" % (cb + q, q) for q in range(n_cold)) + " }")
    w("    VK_EMITJ_BEGIN(%d, %d)" % (ni, nip))
    prev = set()
    for s in range(ni):
        cur = set()
        w("    {   // row %d" % s)
        w("        double t;")
        ents = rows[s]
        opened = False
        for g0 in range(0, max(1, len(ents)), ROW_ACC):
            grp = ents[g0:g0 + ROW_ACC]
            if grp:
                w("        double " + ", ".join("a%d_%d" % (g0, q) for q in range(len(grp))) + ";")
            for q, e in enumerate(grp):
                acc = "a%d_%d" % (g0, q)
                cur.add(int(jc[e]))
                first = True
                for qq in range(int(jp[e]), int(jp[e + 1])):
                    c = float(jcoef[qq])
                    kx = ("KG(%d)" if int(jk[qq]) in dyn else "K(%d)") % int(jk[qq])
                    head = kx if c == 1.0 else ("-" + kx if c == -1.0 else "%s%s * %s" % ("-" if c < 0 else "", _coef_lit(c), kx))
                    stm = ["t = %s;" % head]
                    for f in range(mjf):
                        sl = int(jfac[qq, f])
                        if sl == ni + 1:
                            continue
                        stm.append("t = t * %s;" % yname(sl))
                    stm.append("%s = t;" % acc if first else "%s = %s + t;" % (acc, acc))
                    first = False
                    w("        " + " ".join(stm))
            if not opened:
                w("        VK_EMITJ_ROW_OPEN(%d)" % s)
                opened = True
            if grp:
                w("        " + " ".join("R(%d) = -a%d_%d;" % (int(jc[e]), g0, q) for q, e in enumerate(grp)))
        stale = sorted(prev - cur)
        if stale:
            w("        " + " ".join("R(%d) = 0.0;" % c for c in stale))
        w("    }")
        w("    VK_EMITJ_ROW(%d, %d, %d)" % (s, ni, nip))
        prev = cur
    w("    VK_EMITJ_END()")
    w("}")
    w("int launch_jac_%016x(const EmitJacArgs &a, cudaStream_t st)" % h)
    w("{")
    w("    const size_t smem = emitj_smem_bytes(%d, %d, %d);" % (rld, nr, JAC_TB))
    w("    VK_EMITJ_LAUNCH(negjac_%016x, %d)" % (h, JAC_TB))
    w("    return 0;")
    w("}")
    return True


def emit_chemdf(t, name):
    """CUDA source of the emitted chemdf kernel(s) of one network; t = Network.tables() (only the chemdf tables are read)"""
    ni, nr, maxf = t["ni"], t["nr"], t["maxf"]
    rate_fac, rate_pow = t["rate_fac"], t["rate_pow"]
    rhs_ptr, rhs_pair, rhs_coef = t["rhs_ptr"], t["rhs_pair"], t["rhs_coef"]
    h = _fnv_fast(t)
    npair = nr // 2
    dyn = set(int(x) for x in np.asarray(t.get("dyn_k", [])).ravel())
    # terms of every pair in the order they are applied: (species, coef), species' own order preserved (their lists are sorted by reaction)
    terms = [[] for _ in range(npair + 1)]
    for s in range(ni):
        last = -1
        for q in range(rhs_ptr[s], rhs_ptr[s + 1]):
            j = int(rhs_pair[q])
            if j < last:
                raise ValueError("species %d: production / loss terms are not in reaction order - no emitted kernel for this network" % s)
            last = j
            terms[(j - 1) // 2].append((s, float(rhs_coef[q]), q == rhs_ptr[s]))
    n_pass = (ni + MAX_ACC - 1) // MAX_ACC
    per = (ni + n_pass - 1) // n_pass
    groups = [list(range(g * per, min(ni, (g + 1) * per))) for g in range(n_pass)]
    out = []
    w = out.append
    w("// GENERATED by vulcan_b200/emit.py from network '%s' (ni = %d, nr = %d, %d reaction pairs, %d production / loss terms) - do not edit." % (
        name, ni, nr, npair, len(rhs_pair)))
    w("// chem_funs.chemdf (make_chem_funs.py:113-430) as straight-line code; runtime pieces in ../vk_emit_rt.cuh")
    w('#include "../vk_emit_rt.cuh"')
    w("namespace vk { namespace emitted {")
    w("namespace {")
    for g, group in enumerate(groups):
        gset = set(group)
        w("__global__ void __launch_bounds__(VK_EMIT_TB, 2) chemdf_%016x_p%d(EmitArgs a)" % (h, g))
        w("{")
        w("    VK_EMIT_PROLOGUE(%d, %d, %d)" % (ni, nr, 1 if g == 0 else 0))
        w("    double " + ", ".join("f%d = 0.0" % s for s in group) + ";")
        for p in range(npair):
            tl = [x for x in terms[p] if x[0] in gset]
            if not tl:
                continue
            w("    {")
            for hname, i in (("rf", 2 * p + 1), ("rr", 2 * p + 2)):
                expr = ("KG(%d)" if i in dyn else "K(%d)") % i
                stmts = []
                first = True
                for q in range(maxf):
                    sl, pw = int(rate_fac[i, q]), int(rate_pow[i, q])
                    if sl == ni + 1:
                        continue
                    yv = "Y(%d)" % sl
                    fac = yv if pw == 1 else ("(%s * %s)" % (yv, yv) if pw == 2 else "pow(%s, %d.0)" % (yv, pw))
                    if first:
                        stmts.append("double %s = %s * %s;" % (hname, expr, fac))
                        first = False
                    else:
                        stmts.append("%s = %s * %s;" % (hname, hname, fac))
                if first:
                    stmts.append("double %s = %s;" % (hname, expr))
                w("        " + " ".join(stmts))
            w("        const double v = rf - rr;")
            ups = []
            for s_, cf, is_first in tl:
                term = "v" if abs(cf) == 1.0 else "%s * v" % _coef_lit(cf)
                if is_first:
                    ups.append("f%d = %s%s;" % (s_, "-" if cf < 0 else "", term if abs(cf) == 1.0 else "(%s)" % term))
                else:
                    ups.append("f%d = f%d %s %s;" % (s_, s_, "-" if cf < 0 else "+", term))
            w("        " + " ".join(ups))
            w("    }")
        w("    VK_EMIT_STORE_BEGIN(%d)" % ni)
        for s_ in group:
            w("    F(%d) = f%d;" % (s_, s_))
        w("    VK_EMIT_STORE_END(%d, %d, %d)" % (ni, group[0], group[-1] + 1))
        w("}")
    w("int launch_%016x(const EmitArgs &a, cudaStream_t st)" % h)
    w("{")
    w("    const size_t smem = emit_smem_bytes(%d, %d);" % (ni, nr))
    for g in range(n_pass):
        w("    VK_EMIT_LAUNCH(chemdf_%016x_p%d)" % (h, g))
    w("    return 0;")
    w("}")
    have_jac = emit_negjac(t, h, w)
    dl = sorted(dyn)
    w("const int dyn_%016x[] = {%s};" % (h, ", ".join(str(x) for x in dl) if dl else "0"))
    w("const EmitRegistrar reg_%016x(0x%016xull, %d, %d, \"%s\", launch_%016x, %s, dyn_%016x, %d);" % (
        h, h, ni, nr, name, h, ("launch_jac_%016x" % h) if have_jac else "nullptr", h, len(dl)))
    w("}  // namespace")
    w("}}  // namespace vk::emitted")
    return "\n".join(out) + "\n", h


TABLE_KEYS = ("ni", "nr", "maxf", "rate_fac", "rate_pow", "rhs_ptr", "rhs_pair", "rhs_coef",
              "maxjf", "jac_ptr", "jac_row", "jac_col", "jac_k", "jac_coef", "jac_fac", "dyn_k")
HASH_KEYS = TABLE_KEYS[:8]


def register_network(net, name=None):
    """networks/<name>.npz: the chemdf tables of a compiled network (what the emitter and the hash read)"""
    name = name or os.path.splitext(os.path.basename(net.name))[0]
    t = net.tables()
    os.makedirs(NET_DIR, exist_ok=True)
    out = os.path.join(NET_DIR, name + ".npz")
    np.savez_compressed(out, **{k: np.asarray(t[k]) for k in TABLE_KEYS})
    return out


def register_network_file(path, name=None):
    if path.endswith(".json"):
        with open(path) as f:
            net = Network.from_json(f.read())
    else:
        net = Network.from_file(path)
    return register_network(net, name)


def has_kernel(net):
    """True when networks/ holds the tables of this network, i.e. the built library carries its emitted chemdf kernel"""
    h = _fnv_fast(net.tables())
    for path in glob.glob(os.path.join(NET_DIR, "*.npz")):
        z = np.load(path)
        if _fnv_fast({k: (int(z[k]) if z[k].ndim == 0 else z[k]) for k in HASH_KEYS}) == h:
            return True
    return False


def generate_all(verbose=True):
    """csrc/gen/chemdf_<name>.cu for every compiled network under networks/; returns the list of generated files"""
    os.makedirs(GEN_DIR, exist_ok=True)
    files = []
    seen = {}
    for path in sorted(glob.glob(os.path.join(NET_DIR, "*.npz"))):
        name = os.path.splitext(os.path.basename(path))[0]
        z = np.load(path)
        t = {k: (int(z[k]) if z[k].ndim == 0 else z[k]) for k in TABLE_KEYS if k in z.files}
        try:
            src, h = emit_chemdf(t, name)
        except ValueError as e:
            if verbose:
                print("emit: %s skipped (%s)" % (name, e))
            continue
        if h in seen:
            if verbose:
                print("emit: %s has the same tables as %s" % (name, seen[h]))
            continue
        seen[h] = name
        out = os.path.join(GEN_DIR, "chemdf_%s.cu" % name)
        old = open(out).read() if os.path.exists(out) else None
        if old != src:
            with open(out, "w") as f:
                f.write(src)
        files.append(out)
        if verbose:
            print("emit: %s -> %s (%d lines, hash %016x)" % (name, os.path.relpath(out, HERE), src.count("\n"), h))
    return files


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("registered", register_network_file(p))
    generate_all()
