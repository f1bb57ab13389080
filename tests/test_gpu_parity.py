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


# ---------------------------------------------------------------------------------------------
# the BASELINE.json configurations at oracle-friendly sizes + size-independent properties
# ---------------------------------------------------------------------------------------------
def _run_both(h0, H_gpu_builder, H_oracle_builder, fields, vecs0, tdse, use_h0=True):
    oh = oracle_of(h0)
    phase = port.h0_phase(oh, EXP_FAC) if use_h0 else None
    v_gpu, v_or = vecs0.copy(), vecs0.copy()
    for i, E in enumerate(fields):
        orders = []
        v_or = port.update_step(H_oracle_builder(E), v_or, EXP_FAC, phase=phase, orders=orders)
        if use_h0:
            v_gpu, _ = tdse.update(H_gpu_builder(E), v_gpu, H0=h0)
        else:
            v_gpu, _ = tdse.update(H_gpu_builder(E), v_gpu)
        assert relerr(v_gpu, v_or) < TOL, i
        assert list(tdse.last_orders) == orders, i
    return v_gpu, v_or


def test_config1_ocs_alignment_all_m():
    """examples/ocs_alignment.py: linear rotor (dim_k = 1), Jmax = 30, all m (N = 961), T = 0, strong
    800 nm Gaussian pulse with thresh = 1e3; steps around the pulse maximum and in its screened tail."""
    m = synth.ocs(30)
    h0, H = m["h0"], m["pol"] * (-0.5) * AUPOL
    oH = oracle_of(m["pol"]).scaled(-0.5).scaled(AUPOL)
    tdse = TDSE(t_start=0, t_end=300, dt=0.01)
    tdse.time_grid()
    vecs0 = tdse.init_state(h0, temp=0)
    assert vecs0.shape == (1, 961)
    omega = 2 * np.pi * 299792458.0 / 800e-9 * 1e-12

    def field(t, fwhm=1.0):
        t0 = 2.5 * fwhm / 2
        return [0, 0, 1e10 * np.exp(-4 * np.log(2) * (t - t0) ** 2 / fwhm ** 2) * np.cos(omega * t)]
    times = [1.25 + 0.01 * i for i in range(8)] + [0.02, 299.0]

    def g(E):
        H.field(E, thresh=1e3)
        return H

    def o(E):
        oH.field(E, thresh=1e3)
        return oH
    _run_both(h0, g, o, [field(t) for t in times], vecs0, tdse)


def test_config3_ocs_mixed_field_dressed_states():
    """examples/ocs_mixed_field.py: dc field tilted in the XZ plane (M-mixing) contracted once, ac pulse
    per step, lazy sum Hdc + Hac, initial states = eigenvectors of h0 + Hdc (dense eigh), cos / cos2."""
    m = synth.ocs(8)
    h0, dip, pol = m["h0"], m["dip"], m["pol"]
    beta = 35.0 * np.pi / 180.0
    dc = 20.7e5 * 50 * np.array([np.sin(beta), 0.0, np.cos(beta)])
    Hdc = -1 * dip * dc * AUDIP
    Hac = -0.5 * pol * AUPOL
    tdse = TDSE(t_end=10, dt=0.01)
    tdse.time_grid()
    vecs0 = tdse.init_state(h0 + Hdc, temp=1.0)
    odc = oracle_of(dip)
    odc.field(dc)
    odc = odc.scaled(-1).scaled(AUDIP)
    oac = oracle_of(pol).scaled(-0.5).scaled(AUPOL)
    ref0 = port.init_state(oracle_of(h0).add(odc), temp=1.0)
    assert vecs0.shape == ref0.shape
    # eigenvectors are defined up to a phase: compare projectors
    assert relerr(np.abs(vecs0 @ ref0.conj().T), np.abs(ref0 @ ref0.conj().T)) < 1e-8
    fields = [[0, 0, 1.5e9 * np.exp(-((i - 3) / 2.0) ** 2)] for i in range(6)]

    def g(E):
        Hac.field(E, thresh=1e1)
        return Hdc + Hac

    def o(E):
        oac.field(E, thresh=1e1)
        return odc.add(oac)
    v_gpu, v_or = _run_both(h0, g, o, fields, ref0, tdse)
    for name in ("cos", "cos2"):
        O = m[name]
        oc = oracle_of(O)
        oc.field([0, 0, 1])
        cm = oc.tomat()
        ref = np.array([np.vdot(v, cm.dot(v)) for v in v_or])
        assert relerr(expectation(O, v_gpu), ref) < TOL


def test_config4_optical_centrifuge_complex_mf():
    """Rotating polarisation E0 [cos(bt^2), sin(bt^2), 0]: xx, xy, yx, yy products, complex MF with
    dm = 0, +-2 diagonals, asymmetric top; no H0 split on odd steps (full operator sum)."""
    m = synth.h2s(5)
    h0, pol = m["h0"], m["pol"] * (-0.5 * AUPOL)
    opol = oracle_of(m["pol"]).scaled(-0.5 * AUPOL)
    tdse = TDSE(t_end=10, dt=0.01)
    tdse.time_grid()
    vecs0 = tdse.init_state(h0, temp=10.0)[:40]
    fields = [[4e9 * np.cos(0.3 * i * i), 4e9 * np.sin(0.3 * i * i), 0.0] for i in range(5)]

    def g(E):
        pol.field(E)
        return pol

    def o(E):
        opol.field(E)
        return opol
    _run_both(h0, g, o, fields, vecs0, tdse)
    # without H0=: exp(-i (H0 + V) dt) through the lazy sum with the rank-0 tensor
    oh0 = oracle_of(h0)
    _run_both(h0, lambda E: g(E) + h0, lambda E: o(E).add(oh0), fields[:2], vecs0, tdse, use_h0=False)


def test_wide_k_blocks_column_chunks_and_row_tiles():
    """dim_k > 12 (several column chunks per bra block) and dim_m = 2J+1 up to 53 (state tiles)."""
    st = synth.asymmetric_rotor(*synth.H2S_ABC, 26, Jmin=24)
    pol = synth.lab_tensor(synth.H2S_POL, st)
    pol.field([1e9, -2e9, 3e9])
    o = oracle_of(pol)
    o.field([1e9, -2e9, 3e9])
    N = pol._basis().N
    assert max(pol._basis().dk) > 12
    x = random_states(3, N, seed=4)
    yo = np.array([port.flat_matvec(o, xi) for xi in x])
    assert relerr(gpu_matvec(pol, x), yo) < 1e-13


def test_complex_k_and_scalar_fallback_paths():
    """K made complex by a complex scalar (field.py:939-944) runs the complex-K kernel; RMB_MATVEC=scalar
    forces the general fallback kernel (and the unfused Lanczos path) -- both must match the oracle."""
    import os
    from richmol_b200.field import clear_device_cache
    m = synth.h2o(4)
    h0 = m["h0"]
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    vecs0 = tdse.init_state(h0, temp=30.0)[:9]
    E = [1e9, 5e8, 2e9]
    # complex K: not Hermitian any more, compare a plain matvec
    polc = m["pol"] * (0.3 - 0.7j)
    polc.field(E)
    oc = oracle_of(m["pol"]).scaled(0.3 - 0.7j)
    oc.field(E)
    x = random_states(5, polc._basis().N, seed=6)
    assert relerr(gpu_matvec(polc, x), np.array([port.flat_matvec(oc, xi) for xi in x])) < 1e-13
    # scalar fallback
    os.environ["RMB_MATVEC"] = "scalar"
    clear_device_cache()
    try:
        pol = synth.h2o(4)["pol"] * (-0.5 * AUPOL)
        opol = oracle_of(pol)

        def g(E_):
            pol.field(E_)
            return pol

        def o(E_):
            opol.field(E_)
            return opol
        _run_both(h0, g, o, [E, [0, 0, 3e9]], vecs0, tdse)
    finally:
        del os.environ["RMB_MATVEC"]
        clear_device_cache()


def test_full_size_properties_config2():
    """BASELINE config 2 at full size (N = 12 341): properties that do not need the oracle."""
    import torch
    m = synth.h2o(20)
    h0, dip, pol = m["h0"], m["dip"] * (-AUDIP), m["pol"] * (-0.5 * AUPOL)
    N = h0._basis().N
    assert N == 12341
    dip.field([3e6, 0.0, 4e6])
    pol.field([0.0, 0.0, 3e9], thresh=1e1)
    H = dip + pol
    x = random_states(6, N, seed=8)
    y = gpu_matvec(H, x)
    # linearity and Hermiticity of the assembled operator
    a, b = 0.3 - 1.1j, -0.7 + 0.2j
    assert relerr(gpu_matvec(H, (a * x[0] + b * x[1])[None, :])[0], a * y[0] + b * y[1]) < 1e-13
    assert abs(np.vdot(x[2], y[3]) - np.conj(np.vdot(x[3], y[2]))) < 1e-12 * np.abs(y).max()
    # one oracle matvec as an anchor
    oH = oracle_of(dip)
    oH.field([3e6, 0.0, 4e6])
    op_ = oracle_of(pol)
    op_.field([0.0, 0.0, 3e9], thresh=1e1)
    assert relerr(y[0], port.flat_matvec(oH.add(op_), x[0])) < 1e-13
    # propagation: unitarity (norms kept to the Lanczos tolerance), batch independence (a state gives
    # bit-identical results alone or inside a batch), time-reversal (dt -> -dt undoes the step)
    tdse = TDSE(t_end=10, dt=0.01)
    tdse.time_grid()
    vecs = torch.from_numpy(random_states(37, N, seed=9)).cuda()
    out, _ = tdse.update(H, vecs, H0=h0)
    orders = tdse.last_orders.copy()
    assert orders.min() >= 2 and orders.max() < 30
    n0 = torch.linalg.vector_norm(vecs, dim=1)
    n1 = torch.linalg.vector_norm(out, dim=1)
    assert float((n1 / n0 - 1).abs().max()) < 1e-7
    alone, _ = tdse.update(H, vecs[5:6].clone(), H0=h0)
    assert torch.equal(alone[0], out[5])
    back = TDSE(t_end=10, dt=0.01)
    back.time_grid()
    back._dt = -0.01
    rev, _ = back.update(H, out, H0=h0)
    assert float((rev - vecs).abs().max()) < 1e-7


def test_host_pipeline_with_device_side_observables():
    """numpy in / numpy out: the chunked upload-compute-download pipeline gives the same states as the
    device-resident call, and `expect=` returns the observables of the propagated states."""
    import torch
    m = synth.h2o(6)
    h0, dip, pol, cos2 = m["h0"], m["dip"] * (-AUDIP), m["pol"] * (-0.5 * AUPOL), m["cos2"]
    dip.field([4e6, 0.0, 6e6])
    pol.field([0.0, 0.0, 2e9], thresh=1e1)
    H = dip + pol
    N = h0._basis().N
    vecs = random_states(301, N, seed=12)            # > 128 states: four chunks, ragged last one
    t1 = TDSE(t_end=1, dt=0.01)
    t1.time_grid()
    out_host, _ = t1.update(H, vecs, H0=h0, expect=[cos2])
    t2 = TDSE(t_end=1, dt=0.01)
    t2.time_grid()
    out_dev, _ = t2.update(H, torch.from_numpy(vecs).cuda(), H0=h0)
    assert np.array_equal(out_host, out_dev.cpu().numpy())          # same kernels, same bits
    assert np.array_equal(t1.last_orders, t2.last_orders)
    ev = expectation(cos2, out_dev)
    assert relerr(t1.last_expect[0], ev.cpu().numpy()) < 1e-14
    oc = oracle_of(cos2)
    oc.field([0, 0, 1])
    cm = oc.tomat()
    ref = np.array([np.vdot(v, cm.dot(v)) for v in out_host[:5]])
    assert relerr(t1.last_expect[0][:5], ref) < TOL


def test_wide_k_blocks_dmma_path_propagation():
    """dim_k ~ 20 and dim_m = 81/83: the DMMA kernel (Z staged once, mma.sync m8n8k4 f64) with several
    row tiles per bra block, inside a full split-operator Lanczos step (fused <w,V_k> epilogue)."""
    st = synth.asymmetric_rotor(*synth.H2S_ABC, 41, Jmin=40)
    h0 = synth.hamiltonian_tensor(st)
    pol = synth.lab_tensor(synth.H2S_POL, st) * (-0.5 * AUPOL)
    N = pol._basis().N
    assert max(pol._basis().dk) >= 20
    opol = oracle_of(pol)
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    vecs0 = random_states(5, N, seed=21)
    E = [3e9, -2e9, 4e9]
    pol.field(E)
    opol.field(E)
    x = random_states(3, N, seed=22)
    assert relerr(gpu_matvec(pol, x), np.array([port.flat_matvec(opol, xi) for xi in x])) < 1e-13
    _run_both(h0, lambda E_: pol, lambda E_: opol, [E], vecs0, tdse)


def test_sub_batching_and_strided_states():
    """A workspace budget smaller than the ensemble forces the sub-batch loop; a leading dimension larger
    than N (rows of a wider array) must give the same states."""
    import ctypes as C
    import torch
    from richmol_b200 import _lib
    m = synth.h2o(5)
    h0, pol = m["h0"], m["pol"] * (-0.5 * AUPOL)
    pol.field([1e9, 0.0, 2e9])
    N = pol._basis().N
    vecs = random_states(45, N, seed=31)
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    ref, _ = tdse.update(pol, torch.from_numpy(vecs).cuda(), H0=h0)
    ref_orders = tdse.last_orders.copy()
    # fresh operator with a tiny budget: 7 states per sub-batch
    from richmol_b200.field import clear_device_cache
    clear_device_cache()
    pol2 = synth.h2o(5)["pol"] * (-0.5 * AUPOL)
    pol2.field([1e9, 0.0, 2e9])
    op = pol2._device()
    _lib.check(_lib.lib().rmb_set_workspace_budget(op.handle, C.c_int64(7 * 16 * 16 * int(op.basis.N * 1.2))))
    t2 = TDSE(t_end=1, dt=0.01)
    t2.time_grid()
    out, _ = t2.update(pol2, torch.from_numpy(vecs).cuda(), H0=h0)
    assert torch.equal(out, ref)
    assert np.array_equal(t2.last_orders, ref_orders)
    # strided rows through the C ABI directly
    ld = N + 13
    wide = torch.zeros((45, ld), dtype=torch.complex128, device="cuda")
    wide[:, :N] = torch.from_numpy(vecs).cuda()
    wide[:, N:] = 7.0
    phase = torch.from_numpy(port.h0_phase(oracle_of(h0), EXP_FAC)).cuda()
    orders = np.zeros(45, dtype=np.int32)
    _lib.check(_lib.lib().rmb_propagate_step(op.handle, wide.data_ptr(), 45, ld, EXP_FAC.real, EXP_FAC.imag, 1e-15,
                                             100, phase.data_ptr(), 0, orders.ctypes.data, None))
    torch.cuda.synchronize()
    assert torch.equal(wide[:, :N], ref)
    assert bool((wide[:, N:] == 7.0).all())
    clear_device_cache()


def test_propagate_many_equals_step_by_step():
    """The multi-step entry point is a loop of `update` calls: same states bit for bit (fused and batched
    paths), same skip rule, same observables."""
    for build, dyn_scale in ((lambda: synth.ocs(10), 6e9), (lambda: synth.h2o(4), 3e9)):
        m = build()
        h0, dip, pol, cos2 = m["h0"], m["dip"] * (-AUDIP), m["pol"] * (-0.5 * AUPOL), m["cos2"]
        dc = [2e6, 0.0, 3e6]
        dip.field(dc)
        nsteps = 12
        fields = np.array([[0.0, 0.1 * dyn_scale * np.sin(i), dyn_scale * np.exp(-((i - 5) / 2.5) ** 2)]
                           for i in range(nsteps)])
        fields[10] = [1.0, 2.0, 3.0]                 # screened out by thresh: phases only
        t1 = TDSE(t_end=10, dt=0.01)
        t1.time_grid()
        vecs0 = t1.init_state(h0, temp=5.0)[:7]
        v = vecs0.copy()
        ev_ref = []
        for i in range(nsteps):
            pol.field(fields[i], thresh=1e2)
            v, t = t1.update(dip + pol, v, H0=h0)
            if i % 3 == 2:
                ev_ref.append(expectation(cos2, v))
        t2 = TDSE(t_end=10, dt=0.01)
        t2.time_grid()
        out, times, ev = t2.propagate([(dip, None, None), (pol, fields, 1e2)], vecs0, H0=h0, expect=[cos2], every=3)
        assert np.array_equal(out, v)
        assert abs(times[-1] - t) < 1e-12 and len(times) == nsteps
        assert relerr(ev[:, 0, :], np.array(ev_ref)) < 1e-14
        assert np.array_equal(t2.last_orders, t1.last_orders)


def test_linear_rotor_sliding_window_kernel():
    """Large ensembles of linear rotors use the sliding-window matvec (ring of ket blocks in shared memory,
    lanes over rows): matvec, expectation and the batched Lanczos path (RMB_FUSED=0) vs the oracle and vs
    the fused single-launch path."""
    import os
    import torch
    from richmol_b200.field import clear_device_cache
    m = synth.ocs(14)
    h0, cos2 = m["h0"], m["cos2"]
    dc = [3e6, 0.0, 5e6]
    ac = [2e9, -1e9, 6e9]                      # general polarisation: complex MF, five diagonals
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    N = h0._basis().N
    vecs = random_states(150, N, seed=41)      # ragged: 150 = 18 tiles of 8 + 6

    def build():
        mm = synth.ocs(14)
        Hdc, Hac = mm["dip"] * (-AUDIP), mm["pol"] * (-0.5 * AUPOL)
        Hdc.field(dc)
        Hac.field(ac, thresh=1e2)
        return Hdc + Hac, mm["cos2"]
    H, c2 = build()
    odc, oac = oracle_of(m["dip"]).scaled(-AUDIP), oracle_of(m["pol"]).scaled(-0.5 * AUPOL)
    odc.field(dc)
    oac.field(ac, thresh=1e2)
    oH = odc.add(oac)
    # matvec (>= 64 states: sliding window) against the oracle
    y = gpu_matvec(H, vecs)
    yo = np.array([port.flat_matvec(oH, v) for v in vecs[:12]])
    assert relerr(y[:12], yo) < 1e-13
    y_small = np.vstack([gpu_matvec(H, vecs[i:i + 10]) for i in range(0, 150, 10)])   # tiled kernel
    assert relerr(y, y_small) < 1e-14
    for _ in range(5):                         # the window pipeline is asynchronous: repeat
        assert np.array_equal(gpu_matvec(H, vecs), y)
    # a new field rebuilds the per-block entry lists (fewer surviving diagonals: Z-polarised)
    Hz = m["pol"] * (-0.5 * AUPOL)
    for f in ([1e9, 2e9, 3e9], [0.0, 0.0, 4e9], [1e9, 0.0, 4e9]):
        Hz.field(f, thresh=1e2)
        oz = oracle_of(m["pol"]).scaled(-0.5 * AUPOL)
        oz.field(f, thresh=1e2)
        yz = gpu_matvec(Hz, vecs)
        assert relerr(yz[:5], np.array([port.flat_matvec(oz, v) for v in vecs[:5]])) < 1e-13
    # fewer blocks than the window is wide
    for jm in (1, 2, 3):
        ms = synth.ocs(jm)
        Hd, Hp = ms["dip"] * (-AUDIP), ms["pol"] * (-0.5 * AUPOL)
        Hd.field(dc)
        Hp.field(ac, thresh=1e2)
        od, op_ = oracle_of(ms["dip"]).scaled(-AUDIP), oracle_of(ms["pol"]).scaled(-0.5 * AUPOL)
        od.field(dc)
        op_.field(ac, thresh=1e2)
        vs = random_states(40, ms["h0"]._basis().N, seed=jm)
        ys = gpu_matvec(Hd + Hp, vs)
        osum = od.add(op_)
        assert relerr(ys, np.array([port.flat_matvec(osum, v) for v in vs])) < 1e-13
    # the 4-state tile used when the ring of an 8-state tile does not fit (large Jmax)
    os.environ["RMB_LIN_T"] = "4"
    clear_device_cache()
    try:
        H4, _ = build()
        for _ in range(3):
            assert relerr(gpu_matvec(H4, vecs), y) < 1e-14
    finally:
        del os.environ["RMB_LIN_T"]
        clear_device_cache()
    # expectation through the fused epilogue of the sliding-window kernel
    oc = oracle_of(cos2)
    oc.field([0, 0, 1])
    cm = oc.tomat()
    ev = expectation(c2, vecs)
    assert relerr(ev[:12], np.array([np.vdot(v, cm.dot(v)) for v in vecs[:12]])) < 1e-12
    # Lanczos: fused single-launch path vs batched path with the sliding-window matvec
    out_fused, _ = tdse.update(H, vecs, H0=h0)
    orders_fused = tdse.last_orders.copy()
    os.environ["RMB_FUSED"] = "0"
    clear_device_cache()
    try:
        H2, _ = build()
        t2 = TDSE(t_end=1, dt=0.01)
        t2.time_grid()
        out_lin, _ = t2.update(H2, vecs, H0=h0)
        assert np.array_equal(t2.last_orders, orders_fused)
        assert relerr(out_lin, out_fused) < 1e-12
        ref = port.update_step(oH, vecs[:6].copy(), EXP_FAC, phase=port.h0_phase(oracle_of(h0), EXP_FAC))
        assert relerr(out_lin[:6], ref) < TOL
    finally:
        del os.environ["RMB_FUSED"]
        clear_device_cache()


def test_tiled_matvec_multi_tile_ctas_and_field_changes(monkeypatch):
    """A CTA of the tiled matvec walks several state tiles of one item (RMB_MV2_TILES forces four): ragged last
    unit, descriptor `nnz` refreshed when the surviving diagonals change between launches on the same operator,
    and whole tiles that converge before their neighbours (states ordered by their Lanczos order)."""
    monkeypatch.setenv("RMB_MV2_TILES", "4")
    m = synth.h2o(5)
    h0, dip, pol = m["h0"], m["dip"] * (-AUDIP), m["pol"] * (-0.5 * AUPOL)
    N = h0._basis().N
    # (a) matvec, 150 states (tiles of <= 32 states: units of four tiles and a ragged rest), three polarisations
    x = random_states(150, N, seed=11)
    o = oracle_of(dip)
    for E in ([0, 0, 2e6], [1e6, 0, 2e6], [1e6, -3e6, 2e6], [0, 0, 1e6]):
        dip.field(E)
        o.field(E)
        yo = np.array([port.flat_matvec(o, xi) for xi in x])
        assert relerr(gpu_matvec(dip, x), yo) < 1e-13, E
    # (b) propagation: strong-field and field-free-like states in separate tiles
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    rng = np.random.default_rng(5)
    weak = np.zeros((70, N), dtype=np.complex128)
    weak[np.arange(70), rng.integers(0, N, 70)] = 1e-4          # tiny norm: converges in fewer iterations
    vecs0 = np.concatenate([random_states(80, N, seed=3), weak])
    fields = [[2e7 * np.sin(0.5), 0.0, 2e7 * np.cos(0.5)]] * 2
    od, op_ = oracle_of(dip), oracle_of(pol)
    pol.field([0, 0, 1.5e9], thresh=1e1)
    op_.field([0, 0, 1.5e9], thresh=1e1)

    def build(E):
        od.field(E)
        return od.add(op_)
    outs, orders = _oracle_run(h0, build, fields, vecs0.copy())
    assert min(orders[0]) < max(orders[0])                      # tiles do finish at different iterations
    vecs = vecs0.copy()
    for i, E in enumerate(fields):
        dip.field(E)
        vecs, _ = tdse.update(dip + pol, vecs, H0=h0)
        assert relerr(vecs, outs[i]) < TOL, i
        assert list(tdse.last_orders) == orders[i], i
