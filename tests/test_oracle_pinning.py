"""Pins the CPU oracle (oracle/port.py) against the golden vectors generated from the UNMODIFIED
reference (tests/golden/make_golden.py) and, when /root/reference is present, against the live
reference itself."""
import contextlib
import io
import os

import numpy as np
import pytest

from oracle import port, refshim

from helpers import AUPOL, DEBYE, EXP_FAC, golden, load, oracle_of, relerr


def test_g1_reference_unit_test_populations():
    """tests/test_tdse.py:15-110 of the reference: OCS, J even, m = 0, 500 steps."""
    g = golden("g1_ocs_run.npz")
    h0, al = oracle_of(load("g1_ocs_h0.npz")), oracle_of(load("g1_ocs_alpha.npz"))
    al.mul(-0.5)
    al.mul(AUPOL)
    vecs = port.init_state(h0, temp=0)
    phase = port.h0_phase(h0, EXP_FAC)
    k, orders = 0, []
    for i, E in enumerate(g["field"]):
        al.field(E)
        o = []
        vecs = port.update_step(al, vecs, EXP_FAC, phase=phase, orders=o)
        orders.append(o)
        if i % 10 == 0:
            assert relerr(vecs, g["raw"][k]) < 1e-12
            pops = np.abs(vecs[0][:7]) ** 2
            # the reference's own output file, 4 decimals
            assert np.max(np.abs(np.round(pops, 4) - g["pop_lanczos"][k, 1:])) <= 1.0001e-4
            k += 1
    assert np.array_equal(np.array(orders), g["orders"])
    assert relerr(vecs, g["final"]) < 1e-12


def test_g2_thermal_ensemble_with_thresholds():
    g = golden("g2_ocs_run.npz")
    h0, al = oracle_of(load("g2_ocs_h0.npz")), oracle_of(load("g2_ocs_alpha.npz"))
    al.mul(-0.5)
    al.mul(AUPOL)
    vecs = port.init_state(h0, temp=1.0)
    assert relerr(vecs, g["vecs0"]) < 1e-15
    phase = port.h0_phase(h0, EXP_FAC)
    for i, E in enumerate(g["fields"]):
        al.field(E, thresh=float(g["thresh"]))
        o = []
        vecs = port.update_step(al, vecs, EXP_FAC, phase=phase, orders=o)
        assert relerr(vecs, g["outs"][i]) < 1e-12
        assert o == list(g["orders"][i])
        y = port.flat_matvec(al, g["x"]) if al.mfmat else np.zeros_like(g["x"])
        assert relerr(y, g["matvec"][i]) < 1e-14 or not np.any(g["matvec"][i])


def test_g3_camphor_lazy_sum():
    g = golden("g3_camphor_run.npz")
    h0 = oracle_of(load("g3_camphor_h0.npz"))
    mu, al = oracle_of(load("g3_camphor_mu.npz")), oracle_of(load("g3_camphor_alpha.npz"))
    mu.mul(-1.0)
    mu.mul(DEBYE)
    al.mul(-0.5)
    al.mul(AUPOL)
    vecs = port.init_state(h0, temp=2.0)
    assert relerr(vecs, g["vecs0"]) < 1e-15
    phase = port.h0_phase(h0, EXP_FAC)
    mu.field(g["dc"])
    for i, E in enumerate(g["fields"]):
        al.field(E, thresh=float(g["thresh"]))
        H = mu.add(al)
        o = []
        if i % 2 == 0:
            vecs = port.update_step(H, vecs, EXP_FAC, phase=phase, orders=o)
        else:
            vecs = port.update_step(H.add(h0), vecs, EXP_FAC, orders=o)
        assert relerr(vecs, g["outs"][i]) < 1e-12
        assert o == list(g["orders"][i])
        assert relerr(port.flat_matvec(H, g["x"]), g["matvec"][i]) < 1e-14


def test_g4_h2o_trove_rovibrational_dipole():
    """TROVE rovibrational H2O (the fixture of the reference's tests/benchmarks/test_water_stark.py): real dense K
    blocks with dim_k up to 14, dipole in a field with a Y component (complex MF, every m mixed)."""
    g = golden("g4_h2o_trove_run.npz")
    h0, mu = oracle_of(load("g4_h2o_trove_h0.npz")), oracle_of(load("g4_h2o_trove_mu.npz"))
    assert max(max(d.values()) for d in mu.dim_k2.values()) == 14
    mu.mul(-1.0)
    mu.mul(DEBYE)
    full = port.init_state(h0, temp=40.0)
    keep = np.sort(np.argsort(-np.linalg.norm(full, axis=1), kind="stable")[:24])
    vecs = full[keep]
    assert relerr(vecs, g["vecs0"]) < 1e-15
    phase = port.h0_phase(h0, EXP_FAC)
    for i, E in enumerate(g["fields"]):
        mu.field(E)
        o = []
        vecs = port.update_step(mu, vecs, EXP_FAC, phase=phase, orders=o)
        assert relerr(vecs, g["outs"][i]) < 1e-12
        assert o == list(g["orders"][i])
        assert relerr(port.flat_matvec(mu, g["x"]), g["matvec"][i]) < 1e-14


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_port_against_live_reference():
    r = refshim.load()
    cwd = os.getcwd()
    os.chdir(refshim.REF_ROOT)
    try:
        path = 'tests/benchmarks/data/r-camphor_rchm_files/'
        filt = lambda **kw: kw.get('J', 0) <= 2
        with contextlib.redirect_stdout(io.StringIO()):
            mu = r.trove.CarTensTrove(path + 'camphor_energies_j0_j20.rchm',
                                      path + 'camphor_matelem_mu_j<j1>_j<j2>.rchm', bra=filt, ket=filt)
            h0 = r.trove.CarTensTrove(path + 'camphor_energies_j0_j20.rchm', bra=filt, ket=filt)
    finally:
        os.chdir(cwd)
    mu.mul(-DEBYE)
    om, oh = port.OracleTensor(mu), port.OracleTensor(h0)
    tdse = r.tdse.TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    v_ref = tdse.init_state(h0, temp=5.0)
    v_or = port.init_state(oh, temp=5.0)
    assert relerr(v_or, v_ref) < 1e-15
    phase = port.h0_phase(oh, EXP_FAC)
    for E in ([1e7, 0, 0], [2e7, -1e7, 3e7], [0, 0, 5e7]):
        mu.field(E)
        om.field(E)
        v_ref, _ = tdse.update(mu, v_ref, H0=h0)
        v_or = port.update_step(om, v_or, EXP_FAC, phase=phase)
        assert relerr(v_or, v_ref) < 1e-13
    assert relerr(om.tomat().toarray(), mu.tomat(form='full', repres='dense')) < 1e-15
    # the `tol` keyword (tdse.py:301-306): looser stop rule, same vectors and fewer iterations on both sides;
    # reaching `maxorder` raises the reference's ValueError (tdse.py:480-484)
    for tol in (1e-6, 1e-10, 1):
        a, _ = tdse.update(mu, v_ref, H0=h0, tol=tol)
        b = port.update_step(om, v_or, EXP_FAC, phase=phase, tol=tol)
        assert relerr(b, a) < 1e-13, tol
    with pytest.raises(AssertionError):
        tdse.update(mu, v_ref, H0=h0, tol=0)
    mu.field([5e10, 0, 5e10])
    om.field([5e10, 0, 5e10])
    with pytest.raises(ValueError, match="Lanczos reached maximum order"):
        r.tdse._expmv_lanczos(v_ref[0], EXP_FAC, lambda v: port.flat_matvec(om, v), maxorder=4)
    with pytest.raises(ValueError, match="Lanczos reached maximum order"):
        port.expmv_lanczos(v_or[0], EXP_FAC, lambda v: port.flat_matvec(om, v), maxorder=4)
    # field-dressed initial states (tdse.py:231-233: dense eigh of h0 + V), as in examples/ocs_mixed_field.py
    E = [2e7, -1e7, 3e7]
    mu.field(E)
    om.field(E)
    d_ref = tdse.init_state(h0 + mu, temp=3.0)
    d_or = port.init_state(oh.add(om), temp=3.0)
    assert d_or.shape == d_ref.shape and relerr(d_or, d_ref) < 1e-13


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_host_mirror_time_grid_and_init_state_against_live_reference():
    """richmol_b200.TDSE.time_grid / init_state (host side of the drop-in, richmol/tdse.py:146-262) against the
    unmodified reference on tensors adopted with CarTens.from_richmol: grids for several (t_start, t_end, dt,
    units), ensembles for T = 0, low T with the default and a tight partition threshold, and the error behaviour."""
    from richmol_b200 import TDSE
    from richmol_b200.field import CarTens
    r = refshim.load()
    cwd = os.getcwd()
    os.chdir(refshim.REF_ROOT)
    try:
        path = 'tests/benchmarks/data/r-camphor_rchm_files/'
        filt = lambda **kw: kw.get('J', 0) <= 3
        with contextlib.redirect_stdout(io.StringIO()):
            h0_ref = r.trove.CarTensTrove(path + 'camphor_energies_j0_j20.rchm', bra=filt, ket=filt)
    finally:
        os.chdir(cwd)
    h0 = CarTens.from_richmol(h0_ref)
    for kw in (dict(t_end=1, dt=0.01), dict(t_start=0.5, t_end=3.2, dt=0.07),
               dict(t_start=0, t_end=2000, dt=10, t_units="fs"), dict(t_end=0.3, dt=0.01, t_units="ns")):
        ours, ref = TDSE(**kw), r.tdse.TDSE(**kw)
        assert np.array_equal(ours.time_grid(), ref.time_grid())
        for a, b in zip(ours._time_grid, ref._time_grid):
            assert np.array_equal(a, b)
    ours, ref = TDSE(t_end=1, dt=0.01), r.tdse.TDSE(t_end=1, dt=0.01)
    for kw in (dict(temp=0), dict(temp=None), dict(temp=1.5), dict(temp=5.0, thresh=1e-6), dict(temp=0.2, thresh=1e-1)):
        a, b = ours.init_state(h0, **kw), ref.init_state(h0_ref, **kw)
        assert a.shape == b.shape, kw
        assert relerr(a, b) < 1e-15, kw
    with pytest.raises(AssertionError):
        ours.init_state(h0, temp=-1.0)
    with pytest.raises(AssertionError):
        ref.init_state(h0_ref, temp=-1.0)


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_host_mirror_cartesian_matrices_against_live_reference():
    """CarTens.tomat(cart=...) (richmol/field.py:449-569: sum_irrep kron(M_cart, K), block and full form) and the
    scalar algebra of the host mirror (mul / __mul__ / __rmul__, field.py:932-948, 1248-1301) against the reference,
    on the camphor dipole and polarisability (complex M, dense K, four symmetries)."""
    from richmol_b200.field import CarTens
    r = refshim.load()
    cwd = os.getcwd()
    os.chdir(refshim.REF_ROOT)
    try:
        path = 'tests/benchmarks/data/r-camphor_rchm_files/'
        filt = lambda **kw: kw.get('J', 0) <= 2
        with contextlib.redirect_stdout(io.StringIO()):
            refs = [r.trove.CarTensTrove(path + 'camphor_energies_j0_j20.rchm',
                                         path + f'camphor_matelem_{nm}_j<j1>_j<j2>.rchm', bra=filt, ket=filt)
                    for nm in ('mu', 'alpha')]
    finally:
        os.chdir(cwd)
    for ref in refs:
        ours = CarTens.from_richmol(ref)
        assert list(ours.cart) == list(ref.cart) and ours.rank == ref.rank
        for scale in (None, -0.5, 2.0 - 0.25j):
            a, b = ours, ref
            if scale is not None:
                a, b = ours * scale, ref * scale
                a2, b2 = scale * ours, scale * ref
            for cart in ref.cart:
                fa, fb = a.tomat(form='full', cart=cart), b.tomat(form='full', cart=cart)
                assert relerr(fa.toarray(), fb.toarray()) < 1e-15, (cart, scale)
                if scale is not None:
                    assert relerr(a2.tomat(form='full', cart=cart).toarray(), fb.toarray()) < 1e-15
            blk_a, blk_b = a.tomat(form='block', cart=ref.cart[0]), b.tomat(form='block', cart=ref.cart[0])
            assert blk_a.keys() == blk_b.keys()
            for Jp in blk_b:
                assert blk_a[Jp].keys() == blk_b[Jp].keys()
                for sp_ in blk_b[Jp]:
                    assert relerr(blk_a[Jp][sp_].toarray(), blk_b[Jp][sp_].toarray()) < 1e-15
        with pytest.raises(ValueError):
            ours.tomat(cart="nope")
        with pytest.raises(ValueError):
            ref.tomat(cart="nope")


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_port_lanczos_against_the_reference_function():
    """oracle.port.expmv_lanczos against richmol.tdse._expmv_lanczos itself (tdse.py:417-486) with the same matvec
    callable: the exact zero-beta branch (a state the operator does not couple: Gram-Schmidt of the all-ones vector,
    :459-465), an exact eigenvector (beta ~ 1e-16), unnormalised and complex start vectors, several `fac`."""
    r = refshim.load()
    rng = np.random.default_rng(3)
    n = 12
    A = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    A = (A + A.conj().T) / 2
    H = np.zeros((n + 1, n + 1), dtype=complex)
    H[1:, 1:] = A
    mv = lambda v: H @ v
    _, U = np.linalg.eigh(A)
    eig = np.zeros(n + 1, dtype=complex)
    eig[1:] = U[:, 0]
    starts = {"decoupled": np.eye(n + 1)[0].astype(complex), "eigenvector": eig,
              "random": 0.3 * (rng.normal(size=n + 1) + 1j * rng.normal(size=n + 1)),
              "mixed": np.concatenate([[2.0], 1e-3 * U[:, 1]]).astype(complex)}
    for name, v in starts.items():
        for fac in (EXP_FAC, 10 * EXP_FAC, -0.05j, 0.02 - 0.01j):
            info = []
            a = r.tdse._expmv_lanczos(v.copy(), fac, mv)
            b = port.expmv_lanczos(v.copy(), fac, mv, info=info)
            assert relerr(b, a) < 1e-14, (name, fac)
    info = []
    port.expmv_lanczos(starts["decoupled"].copy(), EXP_FAC, mv, info=info)
    assert info == [1]
