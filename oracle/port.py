"""TEST INFRASTRUCTURE ONLY -- CPU oracle: numpy/scipy restatement of the reference hot path.

Nothing under `richmol_b200/` may import this module.  It is used by `tests/` (as the checker),
by `__graft_entry__.smoke()` and by the `cpu_baseline` / `--impl reference` legs of `bench.py`.

Each function restates the algorithm of the cited reference code (paths relative to the
CFEL-CMI/richmol tree), with the same data structures (nested dicts of scipy CSR blocks), the same
order of floating-point operations where it matters, and the same control flow:

    OracleTensor.field   richmol/field.py:1073-1142   (CarTens.field)
    OracleTensor.vec     richmol/field.py:1145-1245   (CarTens.vec, matvec_lib='scipy')
    OracleTensor.mul     richmol/field.py:932-948
    OracleTensor.add     richmol/field.py:951-1070    (add_cartens: frozen mfmat, renamed irreps)
    OracleTensor.tomat   richmol/field.py:449-569 + 659-692 (form='full')
    flat_matvec          richmol/tdse.py:340-362      (cartensvec closure)
    expmv_lanczos        richmol/tdse.py:417-486      (_expmv_lanczos)
    update_step          richmol/tdse.py:336-414      (TDSE.update, propag='internal')
    init_state           richmol/tdse.py:177-262

Pinning: `tests/test_oracle_pinning.py` checks this port against the UNMODIFIED
reference imported from /root/reference (when present) and against golden vectors generated from
it (`tests/golden/`), including the reference's own `pop_lanczos.txt` populations.
The small matrix exponential is `scipy.sparse.linalg.expm`, exactly what the reference calls
(richmol/tdse.py:474; scipy is an unpinned dependency of the reference, setup.py:53).
"""
import itertools

import numpy as np
import scipy.constants as const
from scipy.sparse import bmat, csr_matrix, diags, kron
from scipy.sparse.linalg import expm


class OracleTensor:
    """Host-only tensor with the reference data model (kmat / mmat / mfmat nested dicts)."""

    _BASIS = ("Jlist1", "Jlist2", "symlist1", "symlist2", "dim1", "dim2", "dim_k1", "dim_k2",
              "dim_m1", "dim_m2")

    def __init__(self, src=None):
        if src is not None:
            for a in self._BASIS + ("rank", "cart", "os", "kmat", "mmat"):
                if hasattr(src, a):
                    setattr(self, a, getattr(src, a))
            if not hasattr(src, "mmat") and hasattr(src, "mfmat"):
                self.mfmat = src.mfmat

    # -- field.py:1073-1142 ----------------------------------------------------------------------
    def field(self, field, thresh=None):
        fx, fy, fz = field[:3]
        f = np.array([fx, fy, fz])
        prods = {}
        for comb in itertools.product((0, 1, 2), repeat=self.rank):
            prods["".join("xyz"[c] for c in comb)] = np.prod(f[list(comb)])
        prods["0"] = 1
        if thresh is not None:
            prods = {c: v for c, v in prods.items() if abs(v) >= thresh}
        self.mfmat = {}
        if not prods:
            return
        for Jpair, m_J in self.mmat.items():
            for sympair, m_s in m_J.items():
                blk = {}
                for irrep, m_c in m_s.items():
                    dense = [prods[c] * m_c[c].toarray() for c in m_c if c in prods]
                    if len(dense) == 1:
                        mat = csr_matrix(dense[0])
                    elif dense:
                        mat = csr_matrix(sum(dense))
                    else:
                        continue
                    if thresh is not None and thresh > 0:
                        mat.data[abs(mat.data) < thresh] = 0
                        mat.eliminate_zeros()
                    if mat.nnz > 0:
                        blk[irrep] = mat
                if blk:
                    self.mfmat.setdefault(Jpair, {})[sympair] = blk

    # -- field.py:1145-1245 ----------------------------------------------------------------------
    def vec(self, vec):
        out = {}
        for Jpair in self.mfmat.keys() & self.kmat.keys():
            J1, J2 = Jpair
            mf_J, k_J = self.mfmat[Jpair], self.kmat[Jpair]
            out.setdefault(J1, {})
            for sympair in mf_J.keys() & k_J.keys():
                sym1, sym2 = sympair
                mf, km = mf_J[sympair], k_J[sympair]
                try:
                    x = vec[J2][sym2]
                except KeyError:
                    continue
                xt = x.reshape(self.dim_m2[J2][sym2], self.dim_k2[J2][sym2]).T
                acc = []
                for irrep in mf.keys() & km.keys():
                    t = km[irrep].dot(xt)                      # (dk1, dm2)
                    acc.append(mf[irrep].dot(t.T).reshape(self.dim1[J1][sym1]))
                if not acc:
                    acc = [0]
                if sym1 in out[J1]:
                    out[J1][sym1] += sum(acc)
                else:
                    out[J1][sym1] = sum(acc)
        return out

    # -- field.py:932-948 -----------------------------------------------------------------------
    def mul(self, arg):
        self.kmat = {Jp: {sp_: {ir: v * arg for ir, v in ks.items()} for sp_, ks in kJ.items()}
                     for Jp, kJ in self.kmat.items()}

    def scaled(self, arg):
        new = OracleTensor(self)
        if hasattr(self, "mfmat"):
            new.mfmat = self.mfmat
        new.mul(arg)
        return new

    # -- field.py:951-1070 ----------------------------------------------------------------------
    def add(self, other):
        for t in (self, other):
            if getattr(t, "cart", [None])[0] == "0":
                t.field([0, 0, 1])
        res = OracleTensor()
        for a in self._BASIS:
            setattr(res, a, getattr(self, a))
        res.kmat, res.mfmat = {}, {}
        for t, sfx in ((self, "_1"), (other, "_2")):
            for src, dst in ((t.kmat, res.kmat), (t.mfmat, res.mfmat)):
                for Jpair, d_J in src.items():
                    for sympair, d_s in d_J.items():
                        tgt = dst.setdefault(Jpair, {}).setdefault(sympair, {})
                        for irrep, val in d_s.items():
                            tgt[str(irrep) + sfx] = val
        return res

    # -- field.py:449-569, 659-692 (form='full') ---------------------------------------------------
    def tomat(self, cart=None):
        if cart is None:
            md = self.mfmat
            pick = lambda m_s: m_s
        else:
            md = self.mmat
            pick = lambda m_s: {ir: v[cart] for ir, v in m_s.items() if cart in v}
        blocks = {}
        for Jpair in md.keys() & self.kmat.keys():
            for sympair in md[Jpair].keys() & self.kmat[Jpair].keys():
                mm, kk = pick(md[Jpair][sympair]), self.kmat[Jpair][sympair]
                terms = [kron(mm[ir], kk[ir]) for ir in mm.keys() & kk.keys()]
                if terms:
                    blocks[(Jpair, sympair)] = sum(terms[1:], terms[0])
        rows = []
        for J1 in self.Jlist1:
            for s1 in self.symlist1[J1]:
                rows.append([blocks.get(((J1, J2), (s1, s2)),
                                        csr_matrix((self.dim1[J1][s1], self.dim2[J2][s2])))
                             for J2 in self.Jlist2 for s2 in self.symlist2[J2]])
        return csr_matrix(bmat(rows))

    @property
    def N(self):
        return sum(self.dim2[J][s] for J in self.Jlist2 for s in self.symlist2[J])


# -- tdse.py:340-362 ------------------------------------------------------------------------------
def flat_matvec(H, v):
    d, ind = {}, 0
    for J in H.Jlist2:
        d[J] = {}
        for sym in H.symlist2[J]:
            n = H.dim2[J][sym]
            d[J][sym] = v[ind:ind + n]
            ind += n
    r = H.vec(d)
    parts = []
    for J in H.Jlist2:
        for sym in H.symlist2[J]:
            if J in r and sym in r[J] and not np.isscalar(r[J][sym]):
                parts.append(r[J][sym])
            else:
                parts.append(np.zeros(H.dim2[J][sym], dtype=np.complex128))
    return np.concatenate(parts)


# -- tdse.py:417-486 ------------------------------------------------------------------------------
def expmv_lanczos(vec, fac, matvec, maxorder=100, tol=1e-15, info=None):
    """exp(fac*H) vec by the reference's Lanczos: unnormalised V[0], no re-orthogonalisation,
    stop when sum|u_k - u_{k-1}|^2 <= tol, ValueError when k reaches maxorder."""
    V, W = [vec], []
    T = np.zeros((maxorder, maxorder), dtype=vec.dtype)
    w = matvec(V[0])
    T[0, 0] = np.vdot(w, V[0])
    W.append(w - T[0, 0] * V[0])
    u_k, conv, k = V[0], 1, 1
    while k < maxorder and conv > tol:
        T[k - 1, k] = np.sqrt(sum(np.abs(W[k - 1]) ** 2))
        T[k, k - 1] = T[k - 1, k]
        if not T[k - 1, k] == 0:
            V.append(W[k - 1] / T[k - 1, k])
        else:
            v = np.ones(V[k - 1].shape, dtype=np.complex128)
            for j in range(k):
                v = v - np.vdot(V[j], v) * V[j]
            V.append(v / np.sqrt(sum(np.abs(v) ** 2)))
        w = matvec(V[k])
        T[k, k] = np.vdot(w, V[k])
        W.append(w - T[k, k] * V[k] - T[k - 1, k] * V[k - 1])
        u_prev = u_k
        e = expm(fac * T[:k + 1, :k + 1])
        u_k = sum([e[i, 0] * v_i for i, v_i in enumerate(V)])
        conv = sum(np.abs(u_k - u_prev) ** 2)
        k += 1
    if info is not None:
        info.append(k - 1)       # index of the last iteration (= matvecs - 1)
    if k == maxorder:
        raise ValueError(f"Lanczos reached maximum order of '{maxorder}' without convergence")
    return u_k


def exp_factor(dt, t_to_s=1e-12, enr_to_J=None):
    """tdse.py:336-337; default units ps and cm^-1 (tdse.py:111-143)."""
    if enr_to_J is None:
        enr_to_J = const.value("Planck constant") * 1e2 * const.value("speed of light in vacuum")
    return -1j * dt * t_to_s * enr_to_J / const.value("reduced Planck constant")


def h0_phase(H0, exp_fac):
    """tdse.py:368-373"""
    m = H0.tomat(cart="0")
    assert (m - diags(m.diagonal())).nnz == 0
    return np.exp(exp_fac / 2 * m.diagonal())


# -- tdse.py:375-414 ------------------------------------------------------------------------------
def update_step(H, vecs, exp_fac, phase=None, tol=1e-15, maxorder=100, orders=None):
    """One TDSE.update call (propag='internal').  `phase` = exp(exp_fac/2 * diag(H0)) or None."""
    mv = lambda v: flat_matvec(H, v)
    if phase is not None:
        out = vecs * phase
        if hasattr(H, "mfmat") and len(H.mfmat) > 0:
            for i, v in enumerate(out):
                out[i] = expmv_lanczos(v, exp_fac, mv, maxorder=maxorder, tol=tol, info=orders)
        elif orders is not None:
            orders.extend([0] * len(out))
        out *= phase
    else:
        out = np.empty(vecs.shape, dtype=vecs.dtype)
        for i, v in enumerate(vecs):
            out[i] = expmv_lanczos(v, exp_fac, mv, maxorder=maxorder, tol=tol, info=orders)
    return out


# -- tdse.py:177-262 ------------------------------------------------------------------------------
def init_state(H, temp=None, thresh=1e-3, enr_to_J=None):
    if enr_to_J is None:
        enr_to_J = const.value("Planck constant") * 1e2 * const.value("speed of light in vacuum")
    diag = getattr(H, "cart", [None])[0] == "0"
    if diag:
        H.field([0, 0, 1])
        enrs = H.tomat().diagonal().real.copy()
        vecs = np.eye(len(enrs))
    else:
        enrs, vecs = np.linalg.eigh(H.tomat().toarray())
    enrs = enrs - enrs[0]
    vecs = vecs.T
    if temp is None:
        pass
    elif temp == 0:
        vecs = vecs[:1]
    else:
        w = np.exp(-enrs * enr_to_J / (const.value("Boltzmann constant") * temp))
        w /= np.sum(w)
        keep = [i for i in range(len(w)) if (1 - np.sum(w[:i + 1])) > thresh]
        vecs = vecs[keep] * np.sqrt(w[keep])[:, None]
    return vecs.astype(np.complex128)
