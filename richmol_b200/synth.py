"""Synthetic rigid-rotor tensor sources: in-memory `CarTens` objects for the benchmark configurations.

The reference builds its operators with `richmol.rot` (`solve` + `LabTensor`), which needs py3nj /
spherical / quaternionic / Fortran quadrature and is a one-off input generator, not part of the
per-step hot path.  This module restates the *formulae* that generator
implements so that the benchmark inputs can be fabricated on any box:

  K-tensor  <J_b,k_b| K_w |J_k,k_k> = (-1)^|k_b| sum_s 3j(J_k w J_b; k_k s -k_b) (Us T)_{w s}
            contracted with the rotor eigenvectors            (richmol/rot/labtens.py:482-502)
  M-tensor  <J_b,m_b| M_{w,cart} |J_k,m_k> = sqrt((2J_k+1)(2J_b+1)) (-1)^|m_b|
            sum_s Ux[cart,(w,s)] 3j(J_k w J_b; m_k s -m_b)    (richmol/rot/labtens.py:504-523)
  named observables cos(theta), cos^2(theta)-1/3              (richmol/rot/labtens.py:362-385)
  rigid rotor / Watson A-reduced Hamiltonian in a Wang basis  (richmol/rot/solution.py:227-334,
            richmol/rot/basis.py:702-800), four D2-type symmetry blocks (rot/symmetry.py:179-209)

Everything here is host-side numpy and runs once per model.
"""
import itertools
import math

import numpy as np
from scipy.sparse import csr_matrix
from scipy.special import gammaln

from .field import CarTens

_EPS = np.finfo(np.complex128).eps


# ----------------------------------------------------------------------------------------------
# Wigner 3j symbols (integer angular momenta), vectorised Racah sum
# ----------------------------------------------------------------------------------------------
def wigner3j(j1, j2, j3, m1, m2, m3):
    """3j symbol for integer arguments (arrays broadcast).  The Racah sum runs over at most
    2*min(j)+1 terms; with the small tensor rank j2 <= 2 used here cancellation is negligible."""
    j1, j2, j3, m1, m2, m3 = np.broadcast_arrays(*[np.asarray(a, dtype=np.int64) for a in (j1, j2, j3, m1, m2, m3)])
    shape = j1.shape
    j1, j2, j3, m1, m2, m3 = [a.reshape(-1) for a in (j1, j2, j3, m1, m2, m3)]
    ok = ((m1 + m2 + m3 == 0) & (np.abs(m1) <= j1) & (np.abs(m2) <= j2) & (np.abs(m3) <= j3)
          & (j3 >= np.abs(j1 - j2)) & (j3 <= j1 + j2))
    lg = lambda x: gammaln(np.maximum(x, 0) + 1.0)
    lnpref = 0.5 * (lg(j1 + j2 - j3) + lg(j1 - j2 + j3) + lg(-j1 + j2 + j3) - lg(j1 + j2 + j3 + 1)
                    + lg(j1 + m1) + lg(j1 - m1) + lg(j2 + m2) + lg(j2 - m2) + lg(j3 + m3) + lg(j3 - m3))
    tmin = np.maximum(0, np.maximum(j2 - j3 - m1, j1 - j3 + m2))
    tmax = np.minimum(j1 + j2 - j3, np.minimum(j1 - m1, j2 + m2))
    res = np.zeros(j1.shape, dtype=np.float64)
    nt = int(np.max(np.where(ok, tmax - tmin, -1))) + 1 if ok.any() else 0
    for dt in range(nt):
        t = tmin + dt
        valid = ok & (t <= tmax)
        lnden = (lg(t) + lg(j3 - j2 + t + m1) + lg(j3 - j1 + t - m2) + lg(j1 + j2 - j3 - t)
                 + lg(j1 - t - m1) + lg(j2 - t + m2))
        term = np.exp(lnpref - lnden)
        sign = 1.0 - 2.0 * (t % 2)
        res += np.where(valid, sign * term, 0.0)
    res *= 1.0 - 2.0 * (np.abs(j1 - j2 - m3) % 2)
    return np.where(ok, res, 0.0).reshape(shape)


def clebsch_gordan(j1, m1, j2, m2, j3, m3):
    return ((-1.0) ** (j1 - j2 + m3)) * math.sqrt(2 * j3 + 1) * float(wigner3j(j1, j2, j3, m1, m2, -m3))


# ----------------------------------------------------------------------------------------------
# Cartesian <-> spherical tensor transformation
# ----------------------------------------------------------------------------------------------
_CART = {1: ["x", "y", "z"], 2: ["".join(p) for p in itertools.product("xyz", repeat=2)]}


def cart_to_spher(rank):
    """Returns (Us, Ux, os, cart): Us[(w,s), cart] maps Cartesian components to spherical ones,
    Ux = pinv(Us).  Rank 1: T_{1,-1} = (x - iy)/sqrt2, T_{1,0} = z, T_{1,1} = -(x + iy)/sqrt2;
    rank 2 by Clebsch-Gordan coupling of two rank-1 tensors."""
    s2 = math.sqrt(0.5)
    u1 = np.array([[s2, -1j * s2, 0], [0, 0, 1], [-s2, -1j * s2, 0]], dtype=np.complex128)
    if rank == 1:
        us = u1
        os_ = [(1, -1), (1, 0), (1, 1)]
    elif rank == 2:
        os_ = [(w, s) for w in range(3) for s in range(-w, w + 1)]
        us = np.zeros((9, 9), dtype=np.complex128)
        for i, (w, s) in enumerate(os_):
            for a in (-1, 0, 1):
                b = s - a
                if abs(b) > 1:
                    continue
                cg = clebsch_gordan(1, a, 1, b, w, s)
                us[i] += cg * np.kron(u1[a + 1], u1[b + 1])
    else:
        raise NotImplementedError(f"tensor of rank = {rank} is not implemented")
    # Us is unitary: its pseudo-inverse is exact up to rounding dust (~1e-16), which would otherwise keep
    # physically absent M diagonals alive (e.g. Delta M = +-1, +-2 for the zz component)
    us[np.abs(us) < 1e-13] = 0
    ux = np.linalg.pinv(us)
    ux[np.abs(ux) < 1e-13] = 0
    return us, ux, os_, list(_CART[rank])


# ----------------------------------------------------------------------------------------------
# rotor solutions
# ----------------------------------------------------------------------------------------------
class RotorStates:
    """Field-free rotor eigenstates: for every J and symmetry the energies and the real
    eigenvector coefficients over |J,k>, k = -J..J, plus the list of m quanta kept."""

    def __init__(self):
        self.Jlist = []
        self.sym = {}      # J -> [sym]
        self.enr = {}      # J -> sym -> energies
        self.coef = {}     # J -> sym -> (2J+1, nstates)
        self.mlist = {}    # J -> [m]
        self.label = {}    # J -> sym -> [str]


def linear_rotor(B, Jmax, D=0.0, jfilter=None, mfilter=None):
    """Linear molecule: k = 0, E = B J(J+1) - D J^2 (J+1)^2, one symmetry 'A'."""
    st = RotorStates()
    for J in range(Jmax + 1):
        if jfilter is not None and not jfilter(J):
            continue
        ms = [m for m in range(-J, J + 1) if mfilter is None or mfilter(J, m)]
        if not ms:
            continue
        c = np.zeros((2 * J + 1, 1))
        c[J, 0] = 1.0
        Jf = float(J)
        st.Jlist.append(Jf)
        st.sym[Jf] = ["A"]
        st.enr[Jf] = {"A": np.array([B * J * (J + 1) - D * (J * (J + 1)) ** 2])}
        st.coef[Jf] = {"A": c}
        st.mlist[Jf] = ms
        st.label[Jf] = {"A": [f"{J} 0 0"]}
    return st


def asymmetric_rotor(A, B, C, Jmax, watson=None, Jmin=0, mfilter=None, emax=None):
    """Asymmetric top H = A Jz^2 + B Jx^2 + C Jy^2 (I^r-type axis assignment) with optional
    Watson A-reduction quartic constants `watson = dict(DJ, DJK, DK, dJ, dK)`, diagonalised in the
    four Wang blocks (k parity) x (J + tau parity), labelled 'A', 'B1', 'B2', 'B3'."""
    w = dict(DJ=0.0, DJK=0.0, DK=0.0, dJ=0.0, dK=0.0)
    if watson:
        w.update(watson)
    st = RotorStates()
    names = {(0, 0): "A", (0, 1): "B1", (1, 0): "B2", (1, 1): "B3"}
    for J in range(Jmin, Jmax + 1):
        ms = [m for m in range(-J, J + 1) if mfilter is None or mfilter(J, m)]
        if not ms:
            continue
        dim = 2 * J + 1
        ks = np.arange(-J, J + 1)
        jj = J * (J + 1.0)
        h = np.zeros((dim, dim))
        h[np.arange(dim), np.arange(dim)] = (0.5 * (B + C) * (jj - ks ** 2) + A * ks ** 2
                                            - w["DJ"] * jj ** 2 - w["DJK"] * jj * ks ** 2 - w["DK"] * ks ** 4)
        for i, k in enumerate(ks[:-2]):
            f = math.sqrt((jj - k * (k + 1)) * (jj - (k + 1) * (k + 2)))
            v = (0.25 * (B - C) - w["dJ"] * jj - 0.5 * w["dK"] * (k ** 2 + (k + 2) ** 2)) * f
            h[i, i + 2] = h[i + 2, i] = v
        # Wang functions |K,tau> = (|K> + (-1)^tau |-K>)/sqrt2, K > 0; |0,0> = |0>
        blocks = {key: [] for key in names}
        for K in range(0, J + 1):
            for tau in ((0,) if K == 0 else (0, 1)):
                v = np.zeros(dim)
                if K == 0:
                    v[J] = 1.0
                else:
                    v[J + K] = math.sqrt(0.5)
                    v[J - K] = math.sqrt(0.5) * (-1) ** tau
                blocks[(K % 2, (J + tau) % 2)].append(v)
        Jf = float(J)
        syms, enr, coef, lab = [], {}, {}, {}
        for key in sorted(names, key=lambda kk: names[kk]):
            if not blocks[key]:
                continue
            wmat = np.array(blocks[key]).T                      # (dim, nb)
            e, u = np.linalg.eigh(wmat.T @ h @ wmat)
            # fix the sign of every eigenvector (largest component positive): deterministic inputs
            idx = np.argmax(np.abs(u), axis=0)
            u = u * np.sign(u[idx, np.arange(u.shape[1])])
            if emax is not None:
                keep = e <= emax
                e, u = e[keep], u[:, keep]
            if len(e) == 0:
                continue
            sym = names[key]
            syms.append(sym)
            enr[sym] = e
            coef[sym] = wmat @ u
            lab[sym] = [f"{J} {sym} {i}" for i in range(len(e))]
        if not syms:
            continue
        st.Jlist.append(Jf)
        st.sym[Jf] = syms
        st.enr[Jf] = enr
        st.coef[Jf] = coef
        st.mlist[Jf] = ms
        st.label[Jf] = lab
    return st


# ----------------------------------------------------------------------------------------------
# tensors
# ----------------------------------------------------------------------------------------------
def _basis_attrs(t, st):
    Jl = list(st.Jlist)
    t.Jlist1 = list(Jl)
    t.Jlist2 = list(Jl)
    t.symlist1 = {J: list(st.sym[J]) for J in Jl}
    t.symlist2 = {J: list(st.sym[J]) for J in Jl}
    t.dim_k1 = {J: {s: st.coef[J][s].shape[1] for s in st.sym[J]} for J in Jl}
    t.dim_k2 = {J: dict(v) for J, v in t.dim_k1.items()}
    t.dim_m1 = {J: {s: len(st.mlist[J]) for s in st.sym[J]} for J in Jl}
    t.dim_m2 = {J: dict(v) for J, v in t.dim_m1.items()}
    t.dim1 = {J: {s: t.dim_m1[J][s] * t.dim_k1[J][s] for s in st.sym[J]} for J in Jl}
    t.dim2 = {J: dict(v) for J, v in t.dim1.items()}
    t.quanta_m1 = {J: {s: [int(m) for m in st.mlist[J]] for s in st.sym[J]} for J in Jl}
    t.quanta_m2 = {J: {s: list(v) for s, v in d.items()} for J, d in t.quanta_m1.items()}
    t.quanta_k1 = {J: {s: [(q, float(e)) for q, e in zip(st.label[J][s], st.enr[J][s])] for s in st.sym[J]}
                   for J in Jl}
    t.quanta_k2 = {J: {s: list(v) for s, v in d.items()} for J, d in t.quanta_k1.items()}


def hamiltonian_tensor(st):
    """Field-free Hamiltonian as a rank-0 tensor: K = diag(E), M = identity (what the reference's
    LabTensor(molecule, solution) / CarTensTrove(states) hold; richmol/trove.py:154-176)."""
    t = CarTens()
    t.rank, t.cart, t.os = 0, ["0"], [(0, 0)]
    _basis_attrs(t, st)
    t.kmat, t.mmat = {}, {}
    for J in st.Jlist:
        nm = len(st.mlist[J])
        for s in st.sym[J]:
            t.kmat.setdefault((J, J), {})[(s, s)] = {0: csr_matrix(np.diag(st.enr[J][s]).astype(np.complex128))}
            t.mmat.setdefault((J, J), {})[(s, s)] = {0: {"0": csr_matrix(np.eye(nm, dtype=np.complex128))}}
    return t


def lab_tensor(arg, st, thresh=None):
    """Laboratory-frame tensor operator from a molecular-frame Cartesian tensor (`arg` = vector of 3
    or 3x3 matrix) or a named observable ('costheta', 'cos2theta')."""
    thr = _EPS if thresh is None else thresh
    t = CarTens()
    if isinstance(arg, str):
        name = arg.lower()
        if name == "costheta":
            os_, ux_val = [(1, 0)], 1.0
        elif name == "cos2theta":
            os_, ux_val = [(2, 0)], 2.0 / 3.0
        else:
            raise ValueError(f"unknown name for tensor operator: '{arg}'")
        rank, cart = 0, ["0"]
        ux = np.full((1, 1), ux_val, dtype=np.complex128)
        ust = np.ones(1, dtype=np.complex128)             # (Us T) per (w, s)
    else:
        tens = np.asarray(arg, dtype=np.float64)
        if not all(d == 3 for d in tens.shape):
            raise ValueError(f"input tensor has inappropriate shape: '{tens.shape}'")
        rank = tens.ndim
        us, ux, os_, cart = cart_to_spher(rank)
        ust = us @ tens.reshape(-1)
    t.rank, t.cart, t.os = rank, cart, os_
    _basis_attrs(t, st)
    irreps = sorted(set(w for w, _ in os_))
    sig = {w: [(i, s) for i, (ww, s) in enumerate(os_) if ww == w] for w in irreps}
    t.kmat, t.mmat = {}, {}
    wmax = max(irreps)
    ints = {J: int(round(J)) for J in st.Jlist}
    for J1 in st.Jlist:          # bra
        for J2 in st.Jlist:      # ket
            j1, j2 = ints[J1], ints[J2]
            if abs(j1 - j2) > wmax:
                continue
            m1 = np.array(st.mlist[J1])[:, None]
            m2 = np.array(st.mlist[J2])[None, :]
            k1 = np.arange(-j1, j1 + 1)[:, None]
            k2 = np.arange(-j2, j2 + 1)[None, :]
            mm, kk = {}, {}
            for w in irreps:
                if abs(j1 - j2) > w or j1 + j2 < w:
                    continue
                # primitive K (molecular frame) and M (laboratory frame) matrices
                kprim = np.zeros((2 * j1 + 1, 2 * j2 + 1), dtype=np.complex128)
                mcart = np.zeros((len(cart),) + np.broadcast(m1, m2).shape, dtype=np.complex128)
                for i, s in sig[w]:
                    if abs(ust[i]) > thr:
                        kprim += ust[i] * wigner3j(j2, w, j1, k2, s, -k1)
                    col = ux[:, i]
                    if np.any(np.abs(col) > thr):
                        tj = wigner3j(j2, w, j1, m2, s, -m1)
                        mcart += col[:, None, None] * tj[None]
                kprim *= (1.0 - 2.0 * (np.abs(k1) % 2))
                mcart *= math.sqrt((2 * j1 + 1) * (2 * j2 + 1)) * (1.0 - 2.0 * (np.abs(m1) % 2))
                mcart[np.abs(mcart) < thr] = 0
                kk[w] = kprim
                mm[w] = {c: csr_matrix(mcart[ic]) for ic, c in enumerate(cart) if np.any(mcart[ic] != 0)}
            for s1 in st.sym[J1]:
                for s2 in st.sym[J2]:
                    for w in kk:
                        if not mm[w]:
                            continue
                        me = st.coef[J1][s1].T.conj() @ kk[w] @ st.coef[J2][s2]
                        me[np.abs(me) < thr] = 0
                        if not np.any(me != 0):
                            continue
                        if not np.any(me.imag != 0):
                            me = me.real
                        t.kmat.setdefault((J1, J2), {}).setdefault((s1, s2), {})[w] = csr_matrix(me)
                        t.mmat.setdefault((J1, J2), {}).setdefault((s1, s2), {})[w] = dict(mm[w])
    return t


# ----------------------------------------------------------------------------------------------
# the benchmark molecules (BASELINE.json configs; parameters from the reference's examples)
# ----------------------------------------------------------------------------------------------
OCS_B = 0.2034394           # cm^-1 (tests/benchmarks/data/alignment_ocs: E(J=1) = 0.40687875)
OCS_DIP = [0, 0, -0.31093]  # au (examples/ocs_alignment.py:29)
OCS_POL = [[25.5778097, 0, 0], [0, 25.5778097, 0], [0, 0, 52.4651140]]   # au (:32)


def ocs(Jmax, **kw):
    st = linear_rotor(OCS_B, Jmax, **kw)
    return dict(states=st, h0=hamiltonian_tensor(st), dip=lab_tensor(OCS_DIP, st),
                pol=lab_tensor(OCS_POL, st), cos=lab_tensor("costheta", st),
                cos2=lab_tensor("cos2theta", st))


H2O_ABC = (27.8806, 14.5216, 9.2777)   # cm^-1, ground-state rotational constants (synthetic use)
H2O_WATSON = dict(DJ=1.25e-3, DJK=-5.77e-3, DK=3.25e-2, dJ=5.1e-4, dK=1.3e-3)
H2O_DIP = [0, 0, -0.7288]              # au (examples/h2o_stark_rigrot.py)
H2O_POL = [[9.1369, 0, 0], [0, 9.8701, 0], [0, 0, 9.4486]]


def h2o(Jmax, **kw):
    st = asymmetric_rotor(*H2O_ABC, Jmax, watson=H2O_WATSON, **kw)
    return dict(states=st, h0=hamiltonian_tensor(st), dip=lab_tensor(H2O_DIP, st),
                pol=lab_tensor(H2O_POL, st), cos2=lab_tensor("cos2theta", st))


H2S_ABC = (10.36, 9.02, 4.73)
H2S_POL = [[23.0, 0, 0], [0, 25.5, 0], [0, 0, 24.7]]


def h2s(Jmax, **kw):
    st = asymmetric_rotor(*H2S_ABC, Jmax, **kw)
    return dict(states=st, h0=hamiltonian_tensor(st), pol=lab_tensor(H2S_POL, st),
                cos2=lab_tensor("cos2theta", st))


def trove_style(Jmax, nk=25, seed=0, kscale=1e-2, Jmin=0):
    """Synthetic asymmetric top in the shape of a TROVE rovibrational database (BASELINE config 5; the
    structure follows the H2O fixture of the reference, tests/benchmarks/data/h2o_rchm_files_TROVE/ read by
    richmol/trove.py:130-191): for every J four C2v-like symmetries 'A1', 'A2', 'B1', 'B2' with `nk`
    rovibrational states each, a rank-1 (dipole) tensor whose M factors are the exact 3j expressions of
    `lab_tensor` and whose K factors are dense real random blocks (`default_rng(seed)`, normal * kscale).
    The selection rule couples A1<->A2 and B1<->B2 for J' - J = 0, +-1; K(J2,J1) is chosen as
    +-K(J1,J2)^T with the sign that makes the field-dressed operator Hermitian.
    Returns dict(states, h0, dip)."""
    rng = np.random.default_rng(seed)
    syms = ["A1", "A2", "B1", "B2"]
    partner = {"A1": "A2", "A2": "A1", "B1": "B2", "B2": "B1"}
    st = RotorStates()
    for J in range(Jmin, Jmax + 1):
        Jf = float(J)
        st.Jlist.append(Jf)
        st.sym[Jf] = list(syms)
        st.mlist[Jf] = list(range(-J, J + 1))
        st.enr[Jf], st.coef[Jf], st.label[Jf] = {}, {}, {}
        for i, s in enumerate(syms):
            st.enr[Jf][s] = 9.5 * J * (J + 1) + 1.0 * i + np.sort(rng.uniform(0.0, 4000.0, size=nk))
            st.coef[Jf][s] = np.zeros((2 * J + 1, nk))          # only the shape (dim_k) is used
            st.label[Jf][s] = [f"{J} {s} {v}" for v in range(nk)]
    h0 = hamiltonian_tensor(st)
    us, ux, os_, cart = cart_to_spher(1)
    t = CarTens()
    t.rank, t.cart, t.os = 1, cart, os_
    _basis_attrs(t, st)
    t.kmat, t.mmat = {}, {}
    mtab = {}
    for J1 in st.Jlist:
        for J2 in st.Jlist:
            j1, j2 = int(J1), int(J2)
            if abs(j1 - j2) > 1 or j1 + j2 < 1:
                continue
            m1 = np.array(st.mlist[J1])[:, None]
            m2 = np.array(st.mlist[J2])[None, :]
            mcart = np.zeros((3,) + np.broadcast(m1, m2).shape, dtype=np.complex128)
            for i, (w, s) in enumerate(os_):
                mcart += ux[:, i][:, None, None] * wigner3j(j2, 1, j1, m2, s, -m1)[None]
            mcart *= math.sqrt((2 * j1 + 1) * (2 * j2 + 1)) * (1.0 - 2.0 * (np.abs(m1) % 2))
            mcart[np.abs(mcart) < _EPS] = 0
            mtab[(J1, J2)] = mcart
    for (J1, J2), mc in mtab.items():
        if J1 > J2:
            continue
        # M_c(J2,J1) = sgn * M_c(J1,J2)^+ for every Cartesian component: K(J2,J1) = sgn * K(J1,J2)^T
        back = mtab[(J2, J1)]
        ref = np.conj(np.transpose(mc, (0, 2, 1)))
        i = np.unravel_index(np.argmax(np.abs(ref)), ref.shape)
        sgn = float(np.sign((back[i] / ref[i]).real))
        assert np.allclose(back, sgn * ref, atol=1e-12)
        mm = {c: csr_matrix(mc[ic]) for ic, c in enumerate(cart) if np.any(mc[ic] != 0)}
        mmb = {c: csr_matrix(back[ic]) for ic, c in enumerate(cart) if np.any(back[ic] != 0)}
        for s1 in syms:
            s2 = partner[s1]
            if J1 == J2 and s1 > s2:
                continue
            k = rng.normal(size=(nk, nk)) * kscale
            t.kmat.setdefault((J1, J2), {})[(s1, s2)] = {1: csr_matrix(k)}
            t.mmat.setdefault((J1, J2), {})[(s1, s2)] = {1: dict(mm)}
            t.kmat.setdefault((J2, J1), {})[(s2, s1)] = {1: csr_matrix(sgn * k.T)}
            t.mmat.setdefault((J2, J1), {})[(s2, s1)] = {1: dict(mmb)}
    return dict(states=st, h0=h0, dip=t)
