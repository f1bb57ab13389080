"""Shared helpers of the test-suite: model builders and oracle drivers."""
import os

import numpy as np

from oracle import port
from richmol_b200 import convert_units as cu
from richmol_b200.io import load_cartens

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXP_FAC = port.exp_factor(0.01)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def load(name):
    return load_cartens(os.path.join(GOLDEN, name))


def random_states(n, N, seed=0, normalise=True):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, N)) + 1j * rng.normal(size=(n, N))
    if normalise:
        v /= np.linalg.norm(v, axis=1)[:, None]
    return v


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    d = np.abs(a - b).max() if a.size else 0.0
    s = max(np.abs(b).max() if b.size else 0.0, 1e-300)
    return d / s


def oracle_of(t):
    """OracleTensor sharing the host dictionaries of a richmol_b200 CarTens."""
    return port.OracleTensor(t)


def flat_to_dict(t, x):
    d, ind = {}, 0
    for J in t.Jlist2:
        d[J] = {}
        for sym in t.symlist2[J]:
            d[J][sym] = x[ind: ind + t.dim2[J][sym]]
            ind += t.dim2[J][sym]
    return d


def dict_to_flat(t, y):
    return np.concatenate([
        y[J][sym] if J in y and sym in y[J] else np.zeros(t.dim2[J][sym], dtype=np.complex128)
        for J in t.Jlist2 for sym in t.symlist2[J]])


AUPOL = cu.AUpol_x_Vm_to_invcm()
AUDIP = cu.AUdip_x_Vm_to_invcm()
DEBYE = cu.Debye_x_Vm_to_invcm()
