"""CPU study for DESIGN §9 item 2 (no GPU needed): which change to the block-LU solve closes the gap to LAPACK in the element budget of
k1 at dt > 1e5 s?  numpy emulation of the solver the GPU and the oracle run (block Thomas, explicit inverse for the Schur update only,
block LU with NB x NB diagonal blocks and diagonal pivots for every application to a vector) on the reference's own systems
(fixtures HD209S-150 / -400, production dt), against the 80-bit solve of the same system.

Variants:
  base        what ships (NB = 8, diagonal pivots, reference species order)
  order=...   static species permutation applied symmetrically to every block (rows and columns): 'diag' = descending |D_jj| of the
              layer-median diagonal, 'loss' = descending chemical loss frequency
  ppiv        partial pivoting INSIDE the NB x NB diagonal block (rows of the panel only - no exchange across panels)
  refine=n    fp64 refinement passes; 'safe' keeps a correction only if the scaled residual drops
    python scripts/study_pivoting.py [HD209S:400 HD209S:150 ...]
"""
import os
import sys
import numpy as np
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
from helpers import Case          # noqa: E402
from oracle import Oracle         # noqa: E402


def inv_diag_pivots(A, ppiv):
    """Gauss-Jordan inverse of a small block; diagonal pivots, or partial pivoting among the block's own rows."""
    n = A.shape[0]
    A = A.copy()
    perm = np.arange(n)
    for k in range(n):
        if ppiv:
            p = k + int(np.argmax(np.abs(A[k:, k])))
            if p != k:
                A[[k, p]] = A[[p, k]]
                perm[[k, p]] = perm[[p, k]]
        rinv = 1.0 / A[k, k]
        A[k, :] *= rinv
        A[k, k] = rinv
        col = A[:, k].copy()
        col[k] = 0.0
        A[:, k] = np.where(np.arange(n) == k, A[:, k], 0.0)
        A -= np.outer(col, A[k, :])
    out = np.empty_like(A)
    out[:, perm] = A
    return out


def gj_blocked_inverse(S, NB, ppiv):
    """in-place blocked Gauss-Jordan inversion as factor_kernel runs it: panel K, P = A_KK^-1 (diagonal pivots),
    V = P A_Kw, A_iw -= A_iK V (all rows i != K), A_iK <- -A_iK P, A_KK <- P."""
    A = S.copy()
    n = A.shape[0]
    for a0 in range(0, n, NB):
        b0 = min(a0 + NB, n)
        K = slice(a0, b0)
        rest = np.r_[0:a0, b0:n]
        P = inv_diag_pivots(A[K, K], ppiv)
        V = P @ A[K][:, rest]
        AiK = A[rest][:, K].copy()
        A[np.ix_(rest, rest)] -= AiK @ V
        A[np.ix_(rest, np.arange(a0, b0))] = -AiK @ P
        A[np.ix_(np.arange(a0, b0), rest)] = V
        A[K, K] = P
    return A


def block_lu(S, NB, ppiv):
    A = S.copy()
    n = A.shape[0]
    for a0 in range(0, n, NB):
        b0 = min(a0 + NB, n)
        P = inv_diag_pivots(A[a0:b0, a0:b0], ppiv)
        V = P @ A[a0:b0, b0:]
        A[b0:, b0:] -= A[b0:, a0:b0] @ V
        A[b0:, a0:b0] = A[b0:, a0:b0] @ P
        A[a0:b0, b0:] = V
        A[a0:b0, a0:b0] = P
    return A


def block_lu_apply(F, NB, t):
    n = F.shape[0]
    t = t.copy()
    for a0 in range(0, n, NB):
        b0 = min(a0 + NB, n)
        t[b0:] -= F[b0:, a0:b0] @ t[a0:b0]
    for a0 in range(((n - 1) // NB) * NB, -1, -NB):
        b0 = min(a0 + NB, n)
        t[a0:b0] = F[a0:b0, a0:b0] @ t[a0:b0] - F[a0:b0, b0:] @ t[b0:]
    return t


class Solver(object):
    def __init__(self, D, up, dn, NB=8, ppiv=False, perm=None, lapack_blocks=False, winv="lapack"):
        nz, n, _ = D.shape
        self.nz, self.n, self.NB = nz, n, NB
        self.perm = np.arange(n) if perm is None else np.asarray(perm)
        p = self.perm
        self.D, self.up, self.dn = D[:, p][:, :, p], up[:, p], dn[:, p]
        self.F = []
        self.lapack_blocks = lapack_blocks
        Sinv = None
        for j in range(nz):
            S = self.D[j].copy() if j == 0 else self.D[j] - (self.dn[j][:, None] * Sinv) * self.up[j - 1][None, :]
            if j + 1 < nz:
                # the explicit inverse that only the Schur update reads: LAPACK (what the oracle's pivoted Gauss-Jordan amounts to),
                # or the device's blocked Gauss-Jordan with diagonal pivots and NO exchange across panels ("gj-diag")
                Sinv = np.linalg.inv(S) if winv == "lapack" else gj_blocked_inverse(S, NB, ppiv)
            if lapack_blocks:
                import scipy.linalg as sl
                self.F.append(sl.lu_factor(S))
            else:
                self.F.append(block_lu(S, NB, ppiv))

    def _apply(self, j, t):
        if self.lapack_blocks:
            import scipy.linalg as sl
            return sl.lu_solve(self.F[j], t)
        return block_lu_apply(self.F[j], self.NB, t)

    def solve(self, r):
        p = self.perm
        r = r[:, p]
        z = np.empty_like(r)
        for j in range(self.nz):
            t = r[j] if j == 0 else r[j] - self.dn[j] * z[j - 1]
            z[j] = self._apply(j, t)
        x = np.empty_like(r)
        x[-1] = z[-1]
        for j in range(self.nz - 2, -1, -1):
            x[j] = z[j] - self._apply(j, self.up[j] * x[j + 1])
        out = np.empty_like(x)
        out[:, p] = x
        return out


def matvec(D, up, dn, x):
    out = np.einsum("jsc,jc->js", D, x)
    out[:-1] += up[:-1] * x[1:]
    out[1:] += dn[1:] * x[:-1]
    return out


def study(tag, step):
    c = Case(tag, step)
    o = Oracle(c.net)
    atm = o.make_atm(**c.atm_kwargs())
    D, up, dn = o.lhs(atm, c.y, c.k, c.dt)
    rhs = c.fx["chemdf"] + c.fx["diffdf"]
    xt = o.blocktri_truth(D, up, dn, rhs, 3)
    compo = c.st["compo"]
    tot = (c.y[:, :, None] * compo[None]).sum(axis=(0, 1))
    bud = lambda v: (v[:, :, None] * compo[None]).sum(axis=(0, 1)) / tot
    res = lambda v: np.abs(rhs - matvec(D, up, dn, v)).max() / np.abs(rhs).max()
    # componentwise (Oettli-Prager) backward error
    absA = lambda v: matvec(np.abs(D), np.abs(up), np.abs(dn), np.abs(v)) + np.abs(rhs)
    cbe = lambda v: np.max(np.abs(rhs - matvec(D, up, dn, v)) / np.maximum(absA(v), 1e-300))
    e = lambda v: np.abs(bud(v) - bud(xt)).max()
    print("== %s-%d  dt %.3e  ni %d" % (tag, step, c.dt, c.ni))
    if "k1" in c.fx:
        print("  %-34s res %.1e  cbe %.1e  budget err %.1e" % ("LAPACK dgbsv (reference's k1)", res(c.fx["k1"]), cbe(c.fx["k1"]), e(c.fx["k1"])))
    xo = o.blocktri_solve(o.blocktri_factor(D, up, dn), up, dn, rhs)
    print("  %-34s res %.1e  cbe %.1e  budget err %.1e" % ("oracle C (what ships)", res(xo), cbe(xo), e(xo)))
    dmed = np.median(np.abs(np.einsum("jss->js", D)), axis=0)
    loss = np.median(np.abs(np.einsum("jss->js", D)) - 1.0 / ((1 + 1 / 2 ** 0.5) * c.dt), axis=0)
    orders = {"ref": None, "diag-desc": np.argsort(-dmed, kind="stable"), "diag-asc": np.argsort(dmed, kind="stable"),
              "loss-desc": np.argsort(-loss, kind="stable")}
    variants = [("base NB=8", dict()), ("NB=8, W by device GJ", dict(winv="gj")), ("NB=8 ppiv, W by device GJ+ppiv", dict(winv="gj", ppiv=True)), ("NB=8 ppiv in panel", dict(ppiv=True)), ("NB=16", dict(NB=16)), ("NB=16 ppiv", dict(NB=16, ppiv=True)),
                ("LAPACK LU per block (full ppiv)", dict(lapack_blocks=True))]
    for oname, perm in orders.items():
        for vname, kw in variants:
            if oname != "ref" and vname not in ("base NB=8", "NB=8 ppiv in panel", "NB=8, W by device GJ"):
                continue
            s = Solver(D, up, dn, perm=perm, **kw)
            x0 = s.solve(rhs)
            line = "  %-34s res %.1e  cbe %.1e  budget err %.1e" % ("%s, order %s" % (vname, oname), res(x0), cbe(x0), e(x0))
            x, best = x0.copy(), cbe(x0)
            for it in range(2):
                xn = x + s.solve(rhs - matvec(D, up, dn, x))
                line += " | r%d %.1e" % (it + 1, e(xn))
                x = xn
            # safeguarded: accept while the componentwise backward error drops
            x, hist = x0.copy(), []
            for it in range(3):
                xn = x + s.solve(rhs - matvec(D, up, dn, x))
                cn = cbe(xn)
                if cn >= best:
                    break
                x, best = xn, cn
                hist.append(e(x))
            line += " | safe(%d) %.1e" % (len(hist), e(x))
            print(line, flush=True)


if __name__ == "__main__":
    cases = sys.argv[1:] or ["HD209S:400", "HD209S:150"]
    for cs in cases:
        t, s = cs.split(":")
        study(t, int(s))
