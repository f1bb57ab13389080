"""Generates the golden vectors in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONHASHSEED=0 python tests/golden/make_golden.py

Inputs come from the reference's own fixtures (tests/benchmarks/data/...), loaded with the
reference's `CarTensTrove`; outputs are produced by the reference's `CarTens.field`, `CarTens.vec`
and `TDSE.update` (propag='internal').  Per-state Lanczos orders are recorded by wrapping (at run
time, not in the tree) `richmol.tdse._expmv_lanczos` with a matvec counter.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402
from richmol_b200.field import CarTens  # noqa: E402
from richmol_b200.io import save_cartens  # noqa: E402

r = refshim.load()
os.chdir(refshim.REF_ROOT)

ORDERS = []
_orig = r.tdse._expmv_lanczos


def _counting(vec, fac, matvec, maxorder=100, tol=1e-15):
    n = [0]

    def mv(v):
        n[0] += 1
        return matvec(v)
    out = _orig(vec, fac, mv, maxorder=maxorder, tol=tol)
    ORDERS.append(n[0] - 1)
    return out


r.tdse._expmv_lanczos = _counting


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name), **arrays)
    print("wrote", name, {k: np.shape(v) for k, v in arrays.items()})


def g4():
    wpath = 'tests/benchmarks/data/h2o_rchm_files_TROVE/'
    wstates = wpath + 'energies_j0_j40_MARVEL_HITRAN.rchm'

    def filt(**kw):
        ok = True
        if 'J' in kw:
            ok = ok and kw['J'] <= 2
        if 'enr' in kw:
            ok = ok and kw['enr'] <= 6000
        return ok
    H0 = quiet(r.trove.CarTensTrove, wstates, bra=filt, ket=filt)
    mu = quiet(r.trove.CarTensTrove, wstates, wpath + 'matelem_MU_j<j1>_j<j2>.rchm', bra=filt, ket=filt)
    save_cartens(os.path.join(HERE, "g4_h2o_trove_h0.npz"), CarTens.from_richmol(H0))
    save_cartens(os.path.join(HERE, "g4_h2o_trove_mu.npz"), CarTens.from_richmol(mu))
    mu = mu * (-1.0) * r.convert_units.Debye_x_Vm_to_invcm()
    tdse = r.tdse.TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    vecs = tdse.init_state(H0, temp=40.0)
    keep = np.sort(np.argsort(-np.linalg.norm(vecs, axis=1), kind="stable")[:24])
    vecs = vecs[keep]                                # the 24 heaviest members (keeps the fixture small)
    N = vecs.shape[1]
    rng = np.random.default_rng(13)
    x = rng.normal(size=N) + 1j * rng.normal(size=N)
    fields = np.array([[3e7 * np.sin(0.4), 2e7 * np.sin(0.9 * i), 3e7 * np.cos(0.4) + 1e7 * np.cos(0.5 * i)]
                       for i in range(6)])
    outs, orders, mv = [], [], []
    v = vecs.copy()
    for i, E in enumerate(fields):
        mu.field(E)
        ORDERS.clear()
        v, _ = tdse.update(mu, v, H0=H0)
        orders.append(list(ORDERS))
        outs.append(v.copy())
        vd, ind = {}, 0
        for J in mu.Jlist2:
            vd[J] = {}
            for sym in mu.symlist2[J]:
                vd[J][sym] = x[ind: ind + mu.dim2[J][sym]]
                ind += mu.dim2[J][sym]
        y = mu.vec(vd)
        mv.append(np.concatenate([y[J][sym] if J in y and sym in y[J] else np.zeros(mu.dim2[J][sym])
                                  for J in mu.Jlist2 for sym in mu.symlist2[J]]))
    save("g4_h2o_trove_run.npz", fields=fields, vecs0=vecs, outs=np.array(outs), orders=np.array(orders),
         x=x, matvec=np.array(mv))


if len(sys.argv) > 1 and sys.argv[1] == "g4":      # only the newest fixture (the others are unchanged)
    g4()
    sys.exit(0)

# ---------------------------------------------------------------------------------------------
# g1: the reference's own unit test (tests/test_tdse.py:15-110): OCS, J even <= 30, m = 0
# ---------------------------------------------------------------------------------------------
path = 'tests/benchmarks/data/alignment_ocs/'


def filt_even_m0(**kw):
    ok = True
    if 'J' in kw:
        ok = ok and kw['J'] <= 30 and kw['J'] % 2 == 0
    if 'm' in kw:
        ok = ok and float(kw['m']) == 0.0
    return ok


states = path + 'matelem/ocs_energies_j0_j30.rchm'
matelem = path + 'matelem/ocs_matelem_alpha_j<j1>_j<j2>.rchm'
H0 = quiet(r.trove.CarTensTrove, states, bra=filt_even_m0, ket=filt_even_m0)
Hbar = quiet(r.trove.CarTensTrove, states, matelem, bra=filt_even_m0, ket=filt_even_m0)
save_cartens(os.path.join(HERE, "g1_ocs_h0.npz"), CarTens.from_richmol(H0))
save_cartens(os.path.join(HERE, "g1_ocs_alpha.npz"), CarTens.from_richmol(Hbar))
Hbar = Hbar * (-0.5) * r.convert_units.AUpol_x_Vm_to_invcm()
with open(path + 'field.txt') as f:
    field = np.array([[0, 0, 1e2 * float(line.split()[3])] for line in f][:500])
tdse = r.tdse.TDSE(t_end=5, dt=0.01)
vecs = tdse.init_state(H0, temp=0)
pops, raw, orders = [], [], []
for ind, _ in enumerate(tdse.time_grid()):
    Hbar.field(field[ind])
    ORDERS.clear()
    vecs, t = tdse.update(Hbar, vecs, H0=H0, matvec_lib='scipy', propag='internal')
    orders.append(list(ORDERS))
    if ind % 10 == 0:
        pops.append([round(t - 0.01, 2)] + [abs(vecs[0][j]) ** 2 for j in range(7)])
        raw.append(vecs.copy())
ref_lanczos = np.loadtxt(path + 'pop_lanczos.txt')[:len(pops)]
assert np.max(np.abs(np.round(np.array(pops)[:, 1:], 4) - ref_lanczos[:, 1:])) <= 1.0001e-4
save("g1_ocs_run.npz", field=field, pops=np.array(pops), raw=np.array(raw), final=vecs,
     orders=np.array(orders), pop_lanczos=ref_lanczos)

# ---------------------------------------------------------------------------------------------
# g2: OCS J <= 10, all m (N = 121), thermal ensemble at 1 K, tilted strong field with thresholds
# ---------------------------------------------------------------------------------------------
filt10 = lambda **kw: kw.get('J', 0) <= 10
H0 = quiet(r.trove.CarTensTrove, states, bra=filt10, ket=filt10)
Hal = quiet(r.trove.CarTensTrove, states, matelem, bra=filt10, ket=filt10)
save_cartens(os.path.join(HERE, "g2_ocs_h0.npz"), CarTens.from_richmol(H0))
save_cartens(os.path.join(HERE, "g2_ocs_alpha.npz"), CarTens.from_richmol(Hal))
Hal = Hal * (-0.5) * r.convert_units.AUpol_x_Vm_to_invcm()
tdse = r.tdse.TDSE(t_end=1, dt=0.01)
tdse.time_grid()
vecs = tdse.init_state(H0, temp=1.0)
fields = np.array([[3e9 * np.cos(0.7 * i), -2e9 * np.sin(0.3 * i + 1), 6e9 * np.exp(-((i - 4) / 3.0) ** 2)]
                   for i in range(10)])
fields[8] = [10.0, 20.0, 30.0]          # every product below thresh=1e3 -> Krylov part skipped
outs, orders = [], []
v = vecs.copy()
rng = np.random.default_rng(7)
x = rng.normal(size=vecs.shape[1]) + 1j * rng.normal(size=vecs.shape[1])
mv = []
for i, E in enumerate(fields):
    Hal.field(E, thresh=1e3)
    ORDERS.clear()
    v, _ = tdse.update(Hal, v, H0=H0)
    orders.append(list(ORDERS) if ORDERS else [0] * len(v))
    outs.append(v.copy())
    if len(Hal.mfmat) > 0:
        vd, ind = {}, 0
        for J in Hal.Jlist2:
            vd[J] = {}
            for sym in Hal.symlist2[J]:
                vd[J][sym] = x[ind: ind + Hal.dim2[J][sym]]
                ind += Hal.dim2[J][sym]
        y = Hal.vec(vd)
        mv.append(np.concatenate([y[J][sym] for J in Hal.Jlist2 for sym in Hal.symlist2[J]]))
    else:
        mv.append(np.zeros_like(x))
save("g2_ocs_run.npz", fields=fields, vecs0=vecs, outs=np.array(outs), orders=np.array(orders),
     x=x, matvec=np.array(mv), thresh=1e3)

# ---------------------------------------------------------------------------------------------
# g3: camphor J <= 3 (asymmetric top, 4 symmetries, complex M, dense K): dipole + polarisability,
#     lazy sum H = -mu.E - 1/2 alpha:EE, no H0 split for half of the steps
# ---------------------------------------------------------------------------------------------
cpath = 'tests/benchmarks/data/r-camphor_rchm_files/'
cstates = cpath + 'camphor_energies_j0_j20.rchm'
filt3 = lambda **kw: kw.get('J', 0) <= 3
H0 = quiet(r.trove.CarTensTrove, cstates, bra=filt3, ket=filt3)
mu = quiet(r.trove.CarTensTrove, cstates, cpath + 'camphor_matelem_mu_j<j1>_j<j2>.rchm', bra=filt3, ket=filt3)
al = quiet(r.trove.CarTensTrove, cstates, cpath + 'camphor_matelem_alpha_j<j1>_j<j2>.rchm', bra=filt3, ket=filt3)
for nm, t in (("h0", H0), ("mu", mu), ("alpha", al)):
    save_cartens(os.path.join(HERE, f"g3_camphor_{nm}.npz"), CarTens.from_richmol(t))
mu = mu * (-1.0) * r.convert_units.Debye_x_Vm_to_invcm()
al = al * (-0.5) * r.convert_units.AUpol_x_Vm_to_invcm()
tdse = r.tdse.TDSE(t_end=1, dt=0.01)
tdse.time_grid()
vecs = tdse.init_state(H0, temp=2.0)
N = vecs.shape[1]
rng = np.random.default_rng(11)
x = rng.normal(size=N) + 1j * rng.normal(size=N)
dc = [2e6, -1e6, 3e6]
mu.field(dc)
fields = np.array([[1e9 * np.cos(i), 2e9 * np.sin(0.5 * i), 4e9 * np.cos(0.2 * i)] for i in range(6)])
outs, orders, mv = [], [], []
v = vecs.copy()
for i, E in enumerate(fields):
    al.field(E, thresh=1e2)
    H = mu + al
    ORDERS.clear()
    if i % 2 == 0:
        v, _ = tdse.update(H, v, H0=H0)
    else:
        v, _ = tdse.update(H + H0, v)
    orders.append(list(ORDERS))
    outs.append(v.copy())
    vd, ind = {}, 0
    for J in H.Jlist2:
        vd[J] = {}
        for sym in H.symlist2[J]:
            vd[J][sym] = x[ind: ind + H.dim2[J][sym]]
            ind += H.dim2[J][sym]
    y = H.vec(vd)
    mv.append(np.concatenate([y[J][sym] if J in y and sym in y[J] else np.zeros(H.dim2[J][sym])
                              for J in H.Jlist2 for sym in H.symlist2[J]]))
save("g3_camphor_run.npz", dc=np.array(dc), fields=fields, vecs0=vecs, outs=np.array(outs),
     orders=np.array(orders), x=x, matvec=np.array(mv), thresh=1e2)

# ---------------------------------------------------------------------------------------------
# g4: H2O rovibrational (TROVE) states and dipole, the fixture of tests/benchmarks/test_water_stark.py:
#     J <= 2, enr <= 6000 cm^-1 (N = 280, real dense K blocks with dim_k up to 14, two to four symmetries per J);
#     static tilted dc field + oscillating field along Y (complex MF, all m mixed), thermal ensemble
# ---------------------------------------------------------------------------------------------
g4()
print("done")
