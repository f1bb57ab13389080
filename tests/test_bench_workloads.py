"""CPU checks of the benchmark's host logic: workload definitions, sharding modes, the synthetic TROVE-style
generator, the byte-compiled reference arm (`oracle/_ref`) and the adoption of reference-fielded tensors."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import port, refshim
from richmol_b200 import TDSE, synth
from richmol_b200.field import CarTens
from richmol_b200.tdse import _as_cartens

from helpers import AUPOL, EXP_FAC, oracle_of, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workload_table_covers_every_baseline_config():
    import json
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    have = {w.config for w in bench.WORKLOADS.values() if w.config is not None}
    assert have == set(range(len(cfgs)))
    assert bench.DEFAULT == "h2s" and bench.WORKLOADS["h2s"].config == 3      # largest single-GPU configuration
    assert set(bench.ALSO_DEFAULT) | {bench.DEFAULT} == set(bench.WORKLOADS)


@pytest.mark.parametrize("name", ["h2o", "ocs_mixed", "asym", "ocs_align"])
def test_sharding_modes(name):
    w = bench.WORKLOADS[name]()
    for world in (1, 2, 8):
        spans = [w.bounds(r, world) for r in range(world)]
        if w.scaling == "weak":
            assert spans == [(r * w.nstates, (r + 1) * w.nstates) for r in range(world)]
            assert w.total_states(world) == world * w.nstates
        elif w.scaling == "strong":
            assert spans[0][0] == 0 and spans[-1][1] == w.nstates
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
        else:
            assert spans == [(0, w.nstates)] * world


def test_small_workload_models_and_rows():
    w = bench.WORKLOADS["ocs_align"]()
    m = bench.build_model(w)
    rows = w.rows(m, 0, 1)
    assert rows.shape == (1, 961) and rows[0, 0] == 1.0
    f = w.field("ac", 150)
    assert f[0] == 0 and f[1] == 0 and abs(f[2]) <= 1e10
    # the port runs one step of it (what `cpu_baseline` and the parity self-check do)
    o = oracle_of(m["terms"][0]["tensor"])
    o.field(f, thresh=1e3)
    orders = []
    out = port.update_step(o, rows, EXP_FAC, phase=port.h0_phase(oracle_of(m["h0"]), EXP_FAC), orders=orders)
    assert abs(np.linalg.norm(out) - 1) < 1e-7 and orders[0] >= 2


def test_dressed_rows_match_init_state_of_the_port():
    """bench.dressed_rows = TDSE.init_state(h0 + Hdc, temp) rows (richmol/tdse.py:231-257) on a small OCS."""
    from richmol_b200 import convert_units as cu
    m = synth.ocs(6)
    Hdc = -1 * m["dip"] * cu.AUdip_x_Vm_to_invcm()
    dc = [1.2e6, 0.0, 1.7e6]
    Hdc.field(dc)
    model = dict(h0=m["h0"], terms=[dict(name="dc", tensor=Hdc, static=dc, thresh=None)])
    rows = bench.dressed_rows(model, 10, 1.0)
    oh, od = oracle_of(m["h0"]), oracle_of(Hdc)
    od.field(dc)
    ref = port.init_state(oh.add(od), temp=1.0, thresh=1e-12)
    n = min(10, len(ref))
    # same Boltzmann weights row by row; eigenvectors are only defined up to rotations inside degenerate
    # subspaces, so each row is checked to be an eigenvector of the dressed Hamiltonian instead
    Hm = oh.add(od).tomat().toarray()
    for a, b in zip(rows[:n], ref[:n]):
        assert abs(np.linalg.norm(a) - np.linalg.norm(b)) < 1e-12
        u = a / np.linalg.norm(a)
        assert np.linalg.norm(Hm @ u - np.vdot(u, Hm @ u) * u) < 1e-10 * np.abs(Hm).max()


def test_trove_style_operator_is_hermitian_and_dense_k():
    m = synth.trove_style(3, nk=5, seed=1)
    dip = m["dip"]
    b = dip._basis()
    assert b.N == sum((2 * J + 1) * 4 * 5 for J in range(4)) and set(b.dk) == {5}
    o = oracle_of(dip)
    o.field([0.3, -0.5, 0.7])
    H = o.tomat().toarray()
    assert np.abs(H - H.conj().T).max() < 1e-15 and np.abs(H).max() > 1e-3
    for kJ in dip.kmat.values():
        for ks in kJ.values():
            for k in ks.values():
                assert k.nnz == 25 and not np.iscomplexobj(k.toarray())


def test_byte_compiled_reference_runs_a_bench_step_without_the_source_tree():
    """oracle/_ref (oracle/build_ref.py) imported in a process that cannot see /root/reference: the unmodified
    reference's TDSE.update on a synthetic tensor agrees with the port (what `--impl reference` times on the box)."""
    if not refshim.built():
        if not refshim.available():
            pytest.skip("neither the reference tree nor oracle/_ref is present")
        from oracle import build_ref
        assert build_ref.build()
    code = r"""
import os, sys, numpy as np
sys.path.insert(0, %r)
from oracle import refshim, port
assert not refshim.available() and refshim.built()
r = refshim.load()
assert r.tdse.__file__.endswith('.pyc')
from richmol_b200 import synth, convert_units as cu
m = synth.ocs(6)
H = m['pol'] * (-0.5 * cu.AUpol_x_Vm_to_invcm())
rH, rh0 = refshim.to_reference(r, H), refshim.to_reference(r, m['h0'])
tdse = r.tdse.TDSE(t_end=1, dt=0.01); tdse.time_grid()
v = tdse.init_state(rh0, temp=1.0)
E = [1e8, 0, 5e9]
rH.field(E, thresh=1e3)
out, _ = tdse.update(rH, H0=rh0, vecs=v)
o = port.OracleTensor(H); o.field(E, thresh=1e3)
fac = port.exp_factor(0.01)
ref = port.update_step(o, v, fac, phase=port.h0_phase(port.OracleTensor(m['h0']), fac))
print('ERR', np.abs(out - ref).max() / np.abs(ref).max())
""" % ROOT
    env = dict(os.environ, RICHMOL_REFERENCE="/nonexistent", PYTHONHASHSEED="0")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    err = float(res.stdout.strip().split("ERR")[-1])
    assert err < 1e-13


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_reference_tensor_after_its_own_field_call_keeps_the_field():
    """`ref.field(E)` leaves `mmat` and `mfmat`; the adopted tensor must carry that field."""
    r = refshim.load()
    m = synth.ocs(4)
    ref = refshim.to_reference(r, m["pol"] * (-0.5 * AUPOL))
    assert not _as_cartens(ref, {})._has_field()
    ref.field([0, 0, 5e9], thresh=1e3)
    ours = _as_cartens(ref, {})
    assert ours._has_field() and not ours._krylov_skippable()
    ref.field([0, 0, 1.0], thresh=1e3)                  # every product screened out: empty mfmat
    ours = _as_cartens(ref, {})
    assert ours._has_field() and ours._krylov_skippable()
    prod = ref * [0, 0, 5e9]                            # CarTens * field (field.py:1259-1262)
    assert _as_cartens(prod, {})._has_field()


def test_adoption_cache_holds_its_key_objects():
    import types
    m = synth.ocs(3)
    H = m["pol"]
    mk = lambda: types.SimpleNamespace(**{a: getattr(H, a) for a in (
        "Jlist1", "Jlist2", "symlist1", "symlist2", "dim1", "dim2", "dim_k1", "dim_k2", "dim_m1", "dim_m2",
        "rank", "cart", "os", "kmat", "mmat")})
    cache = {}
    a = mk()
    ca = _as_cartens(a, cache)
    assert _as_cartens(a, cache) is ca and cache["adopt"][0] is a
    b = mk()
    assert _as_cartens(b, cache) is not ca


def test_scaled_parts_are_shared_between_clones():
    m = synth.ocs(3)
    pol = m["pol"]
    p0 = pol._parts()[0][0]
    a = (-0.5 * pol) * [0, 0, 1e9]
    b = (-0.5 * pol) * [0, 0, 2e9]
    assert a._parts()[0][0] is b._parts()[0][0]          # one packed part -> one device operator
    assert a._parts()[0][0] is not p0
    assert np.allclose(a._parts()[0][0].kpool, -0.5 * p0.kpool)
    assert a._parts()[0][1].serial != b._parts()[0][1].serial


def test_external_propagator_is_refused():
    m = synth.ocs(2)
    tdse = TDSE(t_end=1, dt=0.01)
    tdse.time_grid()
    v = tdse.init_state(m["h0"], temp=0)
    H = m["pol"] * 1.0
    H.field([0, 0, 1e9])
    with pytest.raises(NotImplementedError, match="Expokit"):
        tdse.update(H, v, H0=m["h0"], propag="external")
