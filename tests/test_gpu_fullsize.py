"""GPU parity at the sizes, batches and kernel routes the benchmark runs:

  * g4: TROVE rovibrational H2O golden from the unmodified reference (tiled + DMMA items in one operator);
  * BASELINE config 2 at full size: the bench's operator, its 500-state batch and tiles, >= 10 bench field steps,
    32 sampled rows against `port.update_step` at 1e-10 with identical Lanczos orders;
  * the OCS Jmax = 60 bench operator with a 512-state batch through `k_matvec_lin`;
  * one full-band H2S step at Jmax >= 40 with 32 states through `k_matvec_dmma`;
  * the bench's own self-check helper on a small workload.

Every comparison follows richmol/tdse.py:417-486 (`_expmv_lanczos`) on every sampled row via the pinned port.
"""
import os
import sys

import numpy as np
import pytest

from oracle import port
from richmol_b200 import TDSE, synth

from helpers import AUDIP, AUPOL, DEBYE, EXP_FAC, golden, load, oracle_of, random_states, relerr
from test_gpu_parity import gpu_matvec

pytestmark = pytest.mark.gpu

TOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    return bench


def test_g4_h2o_trove_rovibrational_dipole():
    """Golden g4 (tests/golden/make_golden.py): TROVE rovibrational H2O, J <= 2, real dense K blocks with dim_k up
    to 14 (tiled and DMMA items in one operator), dipole in a field with a Y component (complex MF)."""
    g = golden("g4_h2o_trove_run.npz")
    h0 = load("g4_h2o_trove_h0.npz")
    mu = load("g4_h2o_trove_mu.npz") * (-1.0) * DEBYE
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    vecs = np.array(g["vecs0"])
    for i, E in enumerate(g["fields"]):
        mu.field(E)
        assert relerr(gpu_matvec(mu, g["x"][None, :])[0], g["matvec"][i]) < 1e-13, i
        vecs, _ = tdse.update(mu, vecs, H0=h0)
        assert relerr(vecs, g["outs"][i]) < TOL, i
        assert list(tdse.last_orders) == list(g["orders"][i]), i


def _propagate_and_compare(w, m, nloc, pick, steps, step0=0):
    """The bench step for `steps` field steps on the first `nloc` rows (device-resident, in place); the rows
    `pick` are compared with the port after every step.  Returns the worst relative error."""
    import torch
    b = _bench()
    tdse = TDSE(t_end=1e6, dt=b.DT)
    tdse._time_grid = (None, b._Endless(b.DT), None)
    rows = w.rows(m, 0, nloc)
    v = torch.from_numpy(rows).cuda()
    tensors = [t["tensor"] for t in m["terms"]]
    ots = []
    for t in m["terms"]:
        ot = oracle_of(t["tensor"])
        if t["static"] is not None:
            ot.field(list(t["static"]))
        ots.append(ot)
    phase = port.h0_phase(oracle_of(m["h0"]), EXP_FAC)
    ref = rows[pick].copy()
    worst = 0.0
    for i in range(step0, step0 + steps):
        for t, ot in zip(m["terms"], ots):
            if t["static"] is None:
                f = w.field(t["name"], i)
                kw = {} if t["thresh"] is None else dict(thresh=t["thresh"])
                t["tensor"].field(f, **kw)
                ot.field(f, **kw)
        v, _ = tdse.update(b.hamiltonian(tensors), v, H0=m["h0"], inplace=True)
        H = ots[0]
        for ot in ots[1:]:
            H = H.add(ot)
        orders = []
        ref = port.update_step(H, ref, EXP_FAC, phase=phase, orders=orders)
        got = v[torch.as_tensor(pick, device=v.device)].cpu().numpy()
        err = relerr(got, ref)
        worst = max(worst, err)
        assert err < TOL, (i, err)
        assert [int(o) for o in tdse.last_orders[pick]] == orders, i
    return worst


def test_config2_full_size_bench_batch_against_oracle():
    """BASELINE config 2 exactly as `bench.py --workload h2o` runs it: Jmax = 20 (N = 12 341), the 500-state
    Boltzmann batch (multi-tile CTAs, NC = 12 items), bench field steps 0-4 and 195-199 (start and end of the
    ramp); 32 sampled rows against the port, identical orders (a flipped stop decision costs ~1e-9)."""
    b = _bench()
    w = b.WORKLOADS["h2o"]()
    m = b.build_model(w)
    assert m["h0"]._basis().N == 12341
    pick = sorted(set(int(x) for x in np.linspace(0, 499, 32)))
    _propagate_and_compare(w, m, 500, pick, 5, 0)
    _propagate_and_compare(w, m, 500, pick[::2], 5, 195)


def test_ocs_jmax60_lin_kernel_512_states_against_oracle():
    """`bench.py --workload ocs_batch` operator (OCS Jmax = 60, N = 3721, dc dipole + ac polarisability) with a
    512-state batch: the sliding-window kernel k_matvec_lin inside the batched Lanczos loop."""
    import ctypes as C
    from richmol_b200 import _lib
    b = _bench()
    w = b.WORKLOADS["ocs_batch"]()
    m = b.build_model(w)
    assert m["h0"]._basis().N == 3721
    pick = sorted(set(int(x) for x in np.linspace(0, 511, 24)))
    _propagate_and_compare(w, m, 512, pick, 6, 40)
    op = b.hamiltonian([t["tensor"] for t in m["terms"]])._device()
    info = (C.c_int64 * 8)()
    _lib.check(_lib.lib().rmb_operator_info(op.handle, info))
    assert info[3] in (4, 8) and 512 >= 4 * info[3]          # routed to k_matvec_lin


def test_h2s_full_band_dmma_step_32_states():
    """H2S Jmax = 40, all J (N = 91 881, dim_k up to 21): one optical-centrifuge step (complex MF, Delta m = 0, +-2)
    with 32 states; bra blocks with dim_k > 12 run through k_matvec_dmma, the rest through k_matvec_tiled."""
    import ctypes as C
    import torch
    from richmol_b200 import _lib
    m = synth.h2s(40)
    h0, pol = m["h0"], m["pol"] * (-0.5 * AUPOL)
    N = h0._basis().N
    E = [3e9 * np.cos(0.3), 3e9 * np.sin(0.3), 0.0]
    pol.field(E)
    opol = oracle_of(pol)
    opol.field(E)
    op = pol._device()
    info = (C.c_int64 * 8)()
    _lib.check(_lib.lib().rmb_operator_info(op.handle, info))
    assert info[1] > 0 and info[5] > 12
    rng = np.random.default_rng(3)
    idx = rng.integers(0, N, size=32)
    rows = np.zeros((32, N), dtype=np.complex128)
    rows[np.arange(32), idx] = 1.0
    rows[:8] = random_states(8, N, seed=31)
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    out, _ = tdse.update(pol, torch.from_numpy(rows).cuda(), H0=h0)
    pick = [0, 3, 9, 17, 31]
    orders = []
    ref = port.update_step(opol, rows[pick], EXP_FAC, phase=port.h0_phase(oracle_of(h0), EXP_FAC), orders=orders)
    assert relerr(out[pick].cpu().numpy(), ref) < TOL
    assert [int(o) for o in tdse.last_orders[pick]] == orders


def test_bench_self_check_helper():
    """`bench.parity_check` (the `parity` object of every bench line) on the smallest workload."""
    import torch
    b = _bench()
    w = b.WORKLOADS["ocs_align"]()
    m = b.build_model(w)
    tdse = TDSE(t_end=1e6, dt=b.DT)
    tdse._time_grid = (None, b._Endless(b.DT), None)
    ctx = dict(world=1, rank=0, local=0, dev=torch.device("cuda", 0))
    p = b.parity_check(w, m, tdse, w.rows(m, 0, 1), ctx, nsteps=3)
    assert p["ok"] and p["orders_equal"] and p["parity_max_rel"] < TOL


def test_reference_fielded_tensor_is_adopted_with_its_field():
    """a tensor carrying both `mmat` and `mfmat` without the private `_rmb_field` (the state the
    reference's own `H.field(...)` leaves, examples/ocs_alignment.py:93-96) keeps its interaction."""
    import types
    m = synth.ocs(8)
    h0, H = m["h0"], m["pol"] * (-0.5 * AUPOL)
    E = [0.0, 0.0, 6e9]
    o = oracle_of(H)
    o.field(E, thresh=1e3)
    foreign = types.SimpleNamespace(**{a: getattr(H, a) for a in (
        "Jlist1", "Jlist2", "symlist1", "symlist2", "dim1", "dim2", "dim_k1", "dim_k2", "dim_m1", "dim_m2",
        "rank", "cart", "os", "kmat", "mmat")})
    foreign.mfmat = o.mfmat
    foreign.tomat = None
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    vecs = tdse.init_state(h0, temp=2.0)
    out, _ = tdse.update(foreign, vecs, H0=h0)
    orders = []
    ref = port.update_step(o, vecs, EXP_FAC, phase=port.h0_phase(oracle_of(h0), EXP_FAC), orders=orders)
    assert max(orders) >= 2                                   # the interaction ran
    assert relerr(out, ref) < TOL
    assert list(tdse.last_orders) == orders


def test_expectation_small_batch_of_linear_rotor_with_forced_scalar_rows(monkeypatch):
    """`rmb_expectation` with a batch below 4*lin_T on an operator that is not covered by the tiled
    kernels must take the unfused matvec + dot path, not fail."""
    from richmol_b200.field import clear_device_cache
    from richmol_b200.tdse import expectation
    monkeypatch.setenv("RMB_MATVEC", "scalar")
    clear_device_cache()
    try:
        m = synth.ocs(10)
        cos2 = m["cos2"]
        cos2.field([0, 0, 1])
        v = random_states(3, cos2._basis().N, seed=4)
        ev = expectation(cos2, v)
        oc = oracle_of(cos2)
        oc.field([0, 0, 1])
        cm = oc.tomat()
        ref = np.array([np.vdot(x, cm.dot(x)) for x in v])
        assert relerr(ev, ref) < 1e-12
    finally:
        monkeypatch.delenv("RMB_MATVEC")
        clear_device_cache()


def test_external_propagator_is_refused():
    m = synth.ocs(4)
    H = m["pol"] * (-0.5 * AUPOL)
    H.field([0, 0, 1e9])
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    v = tdse.init_state(m["h0"], temp=0)
    with pytest.raises(NotImplementedError, match="Expokit"):
        tdse.update(H, v, H0=m["h0"], propag="external")
    with pytest.raises(AssertionError):
        tdse.update(H, v, H0=m["h0"], propag="other")


def test_dressed_init_state_with_device_eigensolver():
    """`init_state(h0 + Hdc, temp, device_eigh=True)` -- same energies / Boltzmann weights as the host
    path (richmol/tdse.py:231-257) and every row an eigenvector of the dressed Hamiltonian."""
    m = synth.ocs(12)
    Hdc = -1 * m["dip"] * AUDIP
    Hdc.field([1.2e6, 0.0, 1.7e6])
    H = m["h0"] + Hdc
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    a = tdse.init_state(H, temp=1.0)
    b = tdse.init_state(H, temp=1.0, device_eigh=True)
    assert a.shape == b.shape
    assert np.allclose(np.linalg.norm(a, axis=1), np.linalg.norm(b, axis=1), rtol=1e-10, atol=1e-14)
    Hm = H.tomat(form="full", repres="dense")
    for row in b:
        u = row / np.linalg.norm(row)
        assert np.linalg.norm(Hm @ u - np.vdot(u, Hm @ u) * u) < 1e-9 * np.abs(Hm).max()


def test_small_exponential_against_scipy_expm():
    """a4 (no direct test): `warp_expm_col0`, the routine behind `expm(fac*T_k)[:, 0]` (richmol/tdse.py:474),
    through `rmb_small_expm` against `scipy.linalg.expm` -- all three paths: registers (n <= 32, one sub-step), registers with
    sub-steps (half-width of the spectrum times |fac| up to ~40), shared memory (n > 32); a large common shift of the diagonal
    (the isotropic polarisability term) that the routine splits off as a scalar phase; a complex `fac` with a damping part."""
    import ctypes as C
    from scipy.linalg import expm
    from richmol_b200 import _lib
    rng = np.random.default_rng(11)
    lib = _lib.lib()
    cases = [(2, 0.0, 30.0, -1.88e-3j), (9, -1400.0, 500.0, -1.88e-3j), (10, -1400.0, 500.0, -5e-2j), (32, 200.0, 3000.0, -1.88e-3j),
             (33, -900.0, 400.0, -1.88e-3j), (60, 0.0, 2000.0, -4e-3j), (12, -50.0, 300.0, -2e-3j - 1e-4), (5, 0.0, 0.0, -1.88e-3j)]
    for n, shift, spread, fac in cases:
        nmat = 6
        alpha = shift + spread * (rng.random((nmat, n)) - 0.5) + 1e-14 * rng.standard_normal((nmat, n)) * 1j
        beta = np.zeros((nmat, n))
        beta[:, 1:] = 0.3 * spread * rng.random((nmat, n - 1))
        beta[0, n // 2] = 0.0                                   # a decoupled block (the zero-beta case of the recurrence)
        out = np.zeros((nmat, n), dtype=np.complex128)
        a = np.ascontiguousarray(alpha, dtype=np.complex128)
        _lib.check(lib.rmb_small_expm(nmat, n, a.ctypes.data_as(C.c_void_p), beta.ctypes.data_as(C.c_void_p),
                                      float(np.real(fac)), float(np.imag(fac)), out.ctypes.data_as(C.c_void_p), None))
        for i in range(nmat):
            T = np.diag(alpha[i]) + np.diag(beta[i, 1:], 1) + np.diag(beta[i, 1:], -1)
            ref = expm(fac * T)[:, 0]
            assert np.abs(out[i] - ref).max() < 2e-13 * max(1.0, np.abs(ref).max()), (n, shift, spread, fac, i)
