"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors.

Tolerances: 1e-10 relative on state vectors, populations and expectation values (the figure
BASELINE.json's north_star states); matvecs and field contractions to 1e-13; per-state Lanczos
iteration counts must be identical.
"""
import numpy as np
import pytest

from oracle import port
from richmol_b200 import TDSE, synth
from richmol_b200.tdse import expectation, populations

from helpers import (AUDIP, AUPOL, DEBYE, EXP_FAC, dict_to_flat, flat_to_dict, golden, load, oracle_of,
                     random_states, relerr)

pytestmark = pytest.mark.gpu

TOL = 1e-10


def gpu_matvec(t, x):
    import torch
    from richmol_b200 import _lib
    op = t._device()
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    yd = torch.zeros_like(xd)
    _lib.check(_lib.lib().rmb_matvec(op.handle, xd.data_ptr(), yd.data_ptr(), xd.shape[0], xd.shape[1], None))
    torch.cuda.synchronize()
    return yd.cpu().numpy()


def mfmat_diff(a, b):
    worst = 0.0
    assert a.keys() == b.keys()
    for Jp in a:
        assert a[Jp].keys() == b[Jp].keys(), Jp
        for sp_ in a[Jp]:
            assert a[Jp][sp_].keys() == b[Jp][sp_].keys(), (Jp, sp_)
            for ir in a[Jp][sp_]:
                x, y = a[Jp][sp_][ir].toarray(), b[Jp][sp_][ir].toarray()
                worst = max(worst, np.abs(x - y).max() / max(np.abs(y).max(), 1e-300))
    return worst


CASES = [
    ("ocs_pol", lambda: synth.ocs(12)["pol"], [3e8, -2e8, 9e8], 1e3),
    ("ocs_dip", lambda: synth.ocs(12)["dip"], [3e5, -2e5, 9e5], None),
    ("ocs_pol_m0", lambda: synth.ocs(16, jfilter=lambda J: J % 2 == 0, mfilter=lambda J, m: m == 0)["pol"],
     [0, 0, 1e9], None),
    ("h2o_dip", lambda: synth.h2o(6)["dip"], [1e6, -2e6, 3e6], None),
    ("h2o_pol", lambda: synth.h2o(6)["pol"], [1e9, 2e8, -3e8], 1e3),
    ("h2o_h0", lambda: synth.h2o(5)["h0"], [0, 0, 1], None),
    ("camphor_mu", lambda: load("g3_camphor_mu.npz"), [1e6, 2e6, -1e6], None),
    ("camphor_alpha", lambda: load("g3_camphor_alpha.npz"), [1e9, 2e9, -1e9], 1e2),
]


@pytest.mark.parametrize("name,build,E,thresh", CASES, ids=[c[0] for c in CASES])
def test_field_and_matvec_against_oracle(name, build, E, thresh):
    t = build()
    t.field(E, thresh=thresh)
    o = oracle_of(t)
    o.field(E, thresh=thresh)
    # K1: contracted M factors, same keys (empty blocks dropped) and values
    assert mfmat_diff(t.mfmat, o.mfmat) < 1e-14
    # K2: batched matvec, ragged batch sizes around the state tile
    N = t._basis().N
    for nst in (1, 3, 4, 9):
        x = random_states(nst, N, seed=nst)
        yo = np.array([port.flat_matvec(o, xi) for xi in x])
        assert relerr(gpu_matvec(t, x), yo) < 1e-13
    # CarTens.vec dictionary interface
    x = random_states(1, N, seed=5)[0]
    y = t.vec(flat_to_dict(t, x))
    assert relerr(dict_to_flat(t, y), port.flat_matvec(o, x)) < 1e-13


def test_g1_reference_unit_test(golden_dir):
    """The reference's only TDSE test (tests/test_tdse.py): 500 steps, populations every 10."""
    g = golden("g1_ocs_run.npz")
    h0 = load("g1_ocs_h0.npz")
    H = load("g1_ocs_alpha.npz") * (-0.5) * AUPOL
    tdse = TDSE(t_end=5, dt=0.01)
    vecs = tdse.init_state(h0, temp=0)
    k = 0
    for i, _ in enumerate(tdse.time_grid()):
        H.field(g["field"][i])
        vecs, t = tdse.update(H, vecs, H0=h0, matvec_lib='scipy', propag='internal')
        assert list(tdse.last_orders) == list(g["orders"][i])
        if i % 10 == 0:
            assert relerr(vecs, g["raw"][k]) < TOL
            pops = np.abs(vecs[0][:7]) ** 2
            assert np.max(np.abs(np.round(pops, 4) - g["pop_lanczos"][k, 1:])) <= 1.0001e-4
            assert relerr(pops, g["pops"][k, 1:]) < TOL
            k += 1
    assert abs(t - 5.0) < 1e-12
    assert relerr(vecs, g["final"]) < TOL


def test_g2_thermal_ensemble_thresholds_and_skip():
    g = golden("g2_ocs_run.npz")
    h0 = load("g2_ocs_h0.npz")
    H = load("g2_ocs_alpha.npz") * (-0.5) * AUPOL
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    vecs = tdse.init_state(h0, temp=1.0)
    assert relerr(vecs, g["vecs0"]) < 1e-15
    for i, E in enumerate(g["fields"]):
        H.field(E, thresh=float(g["thresh"]))
        vecs, _ = tdse.update(H, vecs, H0=h0)
        assert relerr(vecs, g["outs"][i]) < TOL, i
        assert list(tdse.last_orders) == list(g["orders"][i]), i
        if np.any(g["matvec"][i]):
            assert relerr(gpu_matvec(H, g["x"][None, :])[0], g["matvec"][i]) < 1e-13
        else:
            assert len(H.mfmat) == 0


def test_g3_camphor_lazy_sum_device_resident():
    import torch
    g = golden("g3_camphor_run.npz")
    h0 = load("g3_camphor_h0.npz")
    mu = load("g3_camphor_mu.npz") * (-1.0) * DEBYE
    al = load("g3_camphor_alpha.npz") * (-0.5) * AUPOL
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    vecs = torch.from_numpy(tdse.init_state(h0, temp=2.0)).cuda()
    mu.field(g["dc"])
    for i, E in enumerate(g["fields"]):
        al.field(E, thresh=float(g["thresh"]))
        H = mu + al
        if i % 2 == 0:
            vecs, _ = tdse.update(H, vecs, H0=h0)
        else:
            vecs, _ = tdse.update(H + h0, vecs)
        assert vecs.is_cuda
        assert relerr(vecs.cpu().numpy(), g["outs"][i]) < TOL, i
        assert list(tdse.last_orders) == list(g["orders"][i]), i
        assert relerr(gpu_matvec(H, g["x"][None, :])[0], g["matvec"][i]) < 1e-13


def _oracle_run(h0, H_builder, fields, vecs, thresh=None):
    oh = oracle_of(h0)
    phase = port.h0_phase(oh, EXP_FAC)
    outs, orders = [], []
    for E in fields:
        oH = H_builder(E)
        o = []
        vecs = port.update_step(oH, vecs, EXP_FAC, phase=phase, orders=o)
        outs.append(vecs.copy())
        orders.append(o)
    return outs, orders


def test_h2o_mixed_field_ensemble_against_oracle():
    """Asymmetric top, dc dipole + ac polarisability (lazy sum), Boltzmann ensemble, M-mixing."""
    m = synth.h2o(5)
    h0, dip, pol = m["h0"], m["dip"] * (-AUDIP), m["pol"] * (-0.5 * AUPOL)
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    vecs0 = tdse.init_state(h0, temp=20.0)
    assert len(vecs0) > 20
    dc = [1.2e7 * np.sin(0.6), 0.0, 1.2e7 * np.cos(0.6)]
    fields = [[0, 0, 2e9 * np.exp(-((i - 3) / 2.0) ** 2)] for i in range(6)]
    dip.field(dc)
    od, op_ = oracle_of(dip), oracle_of(pol)
    od.field(dc)

    def build(E):
        op_.field(E, thresh=1e1)
        return od.add(op_)
    outs, orders = _oracle_run(h0, build, fields, vecs0.copy())
    vecs = vecs0.copy()
    cos2 = m["cos2"]
    cos2.field([0, 0, 1])
    cos2mat = oracle_of(cos2)
    cos2mat.field([0, 0, 1])
    cm = cos2mat.tomat()
    for i, E in enumerate(fields):
        pol.field(E, thresh=1e1)
        vecs, _ = tdse.update(dip + pol, vecs, H0=h0)
        assert relerr(vecs, outs[i]) < TOL, i
        assert list(tdse.last_orders) == orders[i], i
        # K5: fused observables against the host recipe of examples/ocs_alignment.py:99-100
        ev = expectation(cos2, vecs)
        ev_ref = np.array([np.dot(np.conj(v), cm.dot(v)) for v in outs[i]])
        assert relerr(ev, ev_ref) < TOL
        assert relerr(populations(vecs), (np.abs(outs[i]) ** 2).sum(axis=0)) < TOL


def test_zero_beta_fallback_and_exact_eigenvector():
    """A start vector that H maps to zero exercises the beta == 0 branch (tdse.py:459-465)."""
    m = synth.ocs(6, mfilter=lambda J, m: m == 0)
    h0, pol = m["h0"], m["pol"] * (-0.5 * AUPOL)
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    N = pol._basis().N
    pol.field([0, 0, 1e9])
    o = oracle_of(pol)
    o.field([0, 0, 1e9])
    # eigenvector of the interaction operator: W_0 = 0 up to rounding; and the exact zero vector
    hm = o.tomat().toarray()
    w, u = np.linalg.eigh(hm)
    vecs = np.vstack([u[:, 0], np.zeros(N), random_states(1, N, seed=3)[0]]).astype(np.complex128)
    phase = port.h0_phase(oracle_of(h0), EXP_FAC)
    ref_rows, ref_orders = [], []
    for v in vecs:
        oo = []
        try:
            ref_rows.append(port.update_step(o, v[None, :].copy(), EXP_FAC, phase=phase, orders=oo)[0])
        except (ValueError, FloatingPointError):
            ref_rows.append(None)
        ref_orders.append(oo[0] if oo else None)
    out, _ = tdse.update(pol, vecs, H0=h0)
    for i, rr in enumerate(ref_rows):
        if rr is not None and np.all(np.isfinite(rr)):
            assert relerr(out[i], rr) < 1e-9, i
            assert tdse.last_orders[i] == ref_orders[i]


def test_high_order_and_maxorder():
    """Strong field: ~60 Lanczos vectors still match the oracle; a stronger one hits maxorder and
    raises ValueError like the reference (tdse.py:480-484)."""
    m = synth.ocs(24)
    h0, pol = m["h0"], m["pol"] * (-0.5 * AUPOL)
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    o, oh = oracle_of(pol), oracle_of(h0)
    vecs = random_states(2, o.N, seed=1)
    s = 5e10
    E = [0.6 * s, 0.3 * s, s]
    pol.field(E)
    o.field(E)
    orders = []
    ref = port.update_step(o, vecs.copy(), EXP_FAC, phase=port.h0_phase(oh, EXP_FAC), orders=orders)
    out, _ = tdse.update(pol, vecs, H0=h0)
    assert min(orders) > 40
    assert list(tdse.last_orders) == orders
    assert relerr(out, ref) < TOL
    s = 1e11
    pol.field([0.6 * s, 0.3 * s, s])
    with pytest.raises(ValueError, match="maximum order"):
        tdse.update(pol, vecs, H0=h0)


def test_api_error_behaviour():
    m = synth.ocs(4)
    h0, pol = m["h0"], m["pol"]
    tdse = TDSE(t_end=1, dt=0.01)
    vecs = np.zeros((1, pol._basis().N), dtype=np.complex128)
    with pytest.raises(AttributeError):          # time_grid() never called (tdse.py:21)
        pol.field([0, 0, 1e8])
        tdse.update(pol, vecs, H0=h0)
    tdse.time_grid()
    with pytest.raises(AssertionError):
        tdse.update(pol, vecs, H0=h0, tol=2.0)
    with pytest.raises(AssertionError):
        tdse.update(pol, vecs, H0=h0, propag='magic')
    fresh = synth.ocs(4)["pol"]
    with pytest.raises(AttributeError):          # no field applied and no H0 -> CarTens.vec fails
        tdse.update(fresh, vecs)
    out, _ = tdse.update(fresh, vecs + 1.0, H0=h0)   # with H0: phases only (tdse.py:377)
    ph = port.h0_phase(oracle_of(h0), EXP_FAC)
    assert relerr(out, (vecs + 1.0) * ph * ph) < 1e-14
