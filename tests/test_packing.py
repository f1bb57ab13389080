"""Host logic without a GPU: the packed block tables evaluate to the oracle's results."""
import numpy as np
import pytest

from oracle import packed_eval, port
from richmol_b200 import synth
from richmol_b200.packing import field_products

from helpers import AUDIP, AUPOL, load, oracle_of, random_states, relerr


def packed_matvec(t, x):
    parts = t._parts()
    return packed_eval.matvec(t._basis(), [p for p, _, _ in parts], [fs for _, fs, _ in parts], x)


CASES = [
    ("ocs_pol", lambda: synth.ocs(6)["pol"], [3e8, -2e8, 9e8], 1e3),
    ("ocs_dip", lambda: synth.ocs(6)["dip"], [3e5, -2e5, 9e5], None),
    ("ocs_pol_m0", lambda: synth.ocs(8, jfilter=lambda J: J % 2 == 0, mfilter=lambda J, m: m == 0)["pol"],
     [0, 0, 1e9], None),
    ("h2o_dip", lambda: synth.h2o(4)["dip"], [1e6, -2e6, 3e6], None),
    ("h2o_pol", lambda: synth.h2o(4)["pol"], [1e9, 2e8, -3e8], 1e3),
    ("h2o_h0", lambda: synth.h2o(3)["h0"], [0, 0, 1], None),
    ("camphor_mu", lambda: load("g3_camphor_mu.npz"), [1e6, 2e6, -1e6], None),
    ("camphor_alpha", lambda: load("g3_camphor_alpha.npz"), [1e9, 2e9, -1e9], 1e2),
    ("h2o_trove_mu", lambda: load("g4_h2o_trove_mu.npz"), [3e6, -2e6, 1e6], None),
]


@pytest.mark.parametrize("name,build,E,thresh", CASES, ids=[c[0] for c in CASES])
def test_packed_tables_match_oracle(name, build, E, thresh):
    t = build()
    t.field(E, thresh=thresh)
    o = oracle_of(t)
    o.field(E, thresh=thresh)
    x = random_states(3, t._basis().N, seed=1)
    y = packed_matvec(t, x)
    yo = np.array([port.flat_matvec(o, xi) for xi in x])
    assert relerr(y, yo) < 1e-14
    assert t._basis().N == o.N


def test_sum_scaling_and_snapshot_semantics():
    m = synth.h2o(3)
    dip, pol = m["dip"], m["pol"]
    dip.field([1e6, 0, 2e6])
    pol.field([1e9, 0, 3e9], thresh=1e2)
    H = dip * (-AUDIP) + pol * (-0.5 * AUPOL)
    o1, o2 = oracle_of(dip), oracle_of(pol)
    o1.field([1e6, 0, 2e6])
    o2.field([1e9, 0, 3e9], thresh=1e2)
    oH = o1.scaled(-AUDIP).add(o2.scaled(-0.5 * AUPOL))
    x = random_states(2, H._basis().N, seed=2)
    yo = np.array([port.flat_matvec(oH, xi) for xi in x])
    assert relerr(packed_matvec(H, x), yo) < 1e-14
    # subtraction = addition of the negated tensor
    D = dip * (-AUDIP) - pol * (0.5 * AUPOL)
    assert relerr(packed_matvec(D, x), yo) < 1e-14
    # the sum is frozen: a later field() on an operand does not change it (field.py:1029-1068, T5)
    pol.field([5e9, 5e9, 5e9])
    assert relerr(packed_matvec(H, x), yo) < 1e-14
    # renamed irrep keys as in the reference
    assert set(next(iter(next(iter(H.kmat.values())).values())).keys()) <= {"1_1", "0_2", "2_2"}
    # nested sums flatten with nested suffixes
    h0 = m["h0"]
    HH = H + h0
    oHH = oH.add(oracle_of(h0))
    yo2 = np.array([port.flat_matvec(oHH, xi) for xi in x])
    assert relerr(packed_matvec(HH, x), yo2) < 1e-14
    assert [s for _, _, s in HH._parts()] == ["_1_1", "_2_1", "_2"]


def test_field_product_screening():
    cart = ["xx", "xy", "xz", "yx", "yy", "yz", "zx", "zy", "zz"]
    f, dropped = field_products(cart, [0, 0, 1e5], 1e3)
    assert not dropped and f[-1] == 1e10 and np.count_nonzero(f) == 1
    f, dropped = field_products(cart, [10, 20, 30], 1e3)
    assert dropped and not f.any()
    f, dropped = field_products(["0"], [0, 0, 1], None)
    assert not dropped and f[0] == 1
    f, dropped = field_products(["0"], [0, 0, 1], 10.0)     # the "0" product is screened too (T4)
    assert dropped
    with pytest.raises(IndexError):
        field_products(cart, 1.0, None)


def test_identical_m_tables_are_stored_once():
    pol = synth.h2o(4)["pol"]
    pol.field([1e9, 0, 0])
    p = pol._parts()[0][0]
    assert len(p.tb_dm1) < len(p.pr_bra) / 2


def test_mul_creates_fresh_dictionaries():
    pol = synth.ocs(3)["pol"]
    k_before = pol.kmat
    pol.mul(2.0)
    assert pol.kmat is not k_before
    a = k_before[(1.0, 1.0)][("A", "A")][2].toarray()
    b = pol.kmat[(1.0, 1.0)][("A", "A")][2].toarray()
    assert np.allclose(b, 2 * a)
    with pytest.raises(TypeError):
        pol.mul("x")
    with pytest.raises(TypeError):
        pol + 1.0


def test_vec_without_field_raises_attribute_error():
    pol = synth.ocs(2)["pol"]
    assert not hasattr(pol, "mfmat")
    with pytest.raises(AttributeError):
        pol.vec({0.0: {"A": np.ones(1, dtype=complex)}})
    with pytest.raises(AttributeError):
        pol.tomat(form="full")


def test_add_requires_same_basis():
    a, b = synth.ocs(2)["pol"], synth.ocs(3)["pol"]
    a.field([0, 0, 1.0])
    b.field([0, 0, 1.0])
    with pytest.raises(ValueError):
        a + b


@pytest.mark.parametrize("what,E,thresh", [("pol", [0, 0, 2e9], 1e1), ("pol", [3e8, -2e8, 9e8], 1e3),
                                           ("dip", [3e5, 0.0, 9e5], None), ("sum", None, None)])
def test_linear_rotor_entry_lists_and_merge(what, E, thresh):
    """The entry lists of the sliding-window matvec (k_lin_entries: one entry per surviving diagonal, pairs on the
    same ket block and diagonal merged) evaluate to the oracle's matvec; merging removes entries whenever two
    products couple the same blocks (rank 0 and rank 2 of a polarisability, operands of a sum)."""
    m = synth.ocs(8)
    if what == "sum":
        dip, pol = m["dip"] * (-AUDIP), m["pol"] * (-0.5 * AUPOL)
        dip.field([2e7 * np.sin(0.6), 0.0, 2e7 * np.cos(0.6)])
        pol.field([0, 0, 2e9], thresh=1e1)
        t = dip + pol
        od, op_ = oracle_of(dip), oracle_of(pol)
        od.field([2e7 * np.sin(0.6), 0.0, 2e7 * np.cos(0.6)])
        op_.field([0, 0, 2e9], thresh=1e1)
        o = od.add(op_)
    else:
        t = m[what]
        t.field(E, thresh=thresh)
        o = oracle_of(t)
        o.field(E, thresh=thresh)
    parts = t._parts()
    basis = t._basis()
    lists = packed_eval.lin_entry_lists(basis, [p for p, _, _ in parts], [fs for _, fs, _ in parts])
    x = random_states(3, basis.N, seed=2)
    yo = np.array([port.flat_matvec(o, xi) for xi in x])
    assert relerr(packed_eval.lin_matvec(basis, lists, x), yo) < 1e-13
    # no two entries of a bra block share (ket block, diagonal)
    for ents in lists:
        keys = [(b2, doff) for b2, doff, _ in ents]
        assert len(keys) == len(set(keys))
    if what in ("pol", "sum"):
        # unmerged count = surviving diagonals over all products: strictly more than the merged lists hold
        unmerged = 0
        for part, fs, _ in parts:
            val = packed_eval.contract_field(part, fs)
            for p in range(len(part.pr_bra)):
                tt = int(part.pr_table[p])
                dm1, nd, e0 = int(basis.dm[int(part.pr_bra[p])]), int(part.tb_nd[tt]), int(part.tb_off[tt])
                unmerged += int(np.any(val[e0:e0 + dm1 * nd].reshape(dm1, nd) != 0, axis=0).sum())
        assert sum(len(e) for e in lists) < unmerged
