"""Device generator of the laboratory-frame tensor factors.

The reference builds `LabTensor` M and K factors in `richmol/rot/labtens.py:411-638` with nested Python loops and
one py3nj call per 3j symbol (O(J^4) interpreter work).  Here every (J1, J2, omega) block is one CUDA launch
(`rmb_threej_band`, include/richmol_b200.h): the Racah sum of the 3j symbol is evaluated per matrix element on the
GPU.  Formulae (same as `richmol_b200.synth`, which is the host restatement used for CPU-side inputs):

    M_{omega,cart}[m1, m2] = sqrt((2J1+1)(2J2+1)) (-1)^|m1| sum_s Ux[cart,(omega,s)] 3j(J2 omega J1; m2 s -m1)   (:504-523)
    Kprim_omega[k1, k2]    = (-1)^|k1| sum_s (Us T)_{omega s} 3j(J2 omega J1; k2 s -k1)                          (:482-502)
"""
import math

import numpy as np
from scipy.sparse import csr_matrix

from . import _lib
from .synth import _EPS, cart_to_spher


def threej_band(j1, j2, omega, coef, pref=1.0):
    """out[c, a, b] = pref (-1)^|a-j1| sum_s coef[c, s+omega] 3j(j2 omega j1; b-j2, s, -(a-j1)) on the GPU."""
    import ctypes as C
    _lib.require_device()
    coef = np.ascontiguousarray(np.atleast_2d(coef), dtype=np.complex128)
    if coef.shape[1] != 2 * omega + 1:
        raise ValueError(f"coef must have {2 * omega + 1} columns (sigma = -omega..omega)")
    out = np.empty((coef.shape[0], 2 * j1 + 1, 2 * j2 + 1), dtype=np.complex128)
    _lib.check(_lib.lib().rmb_threej_band(int(j1), int(j2), int(omega), coef.shape[0], coef.ctypes.data, float(pref),
                                          out.ctypes.data, None))
    return out


def m_tensor(rank, J1, J2, thresh=None):
    """{omega: {cart: csr (2J1+1 x 2J2+1)}} -- the M factors of a rank-`rank` Cartesian tensor between bra J1 and ket J2
    (all m), zero components dropped, as `richmol_b200.synth.lab_tensor` / the reference's LabTensor hold them."""
    thr = _EPS if thresh is None else thresh
    us, ux, os_, cart = cart_to_spher(rank)
    j1, j2 = int(round(J1)), int(round(J2))
    out = {}
    for w in sorted(set(w for w, _ in os_)):
        if abs(j1 - j2) > w or j1 + j2 < w:
            continue
        coef = np.zeros((len(cart), 2 * w + 1), dtype=np.complex128)
        for i, (ww, s) in enumerate(os_):
            if ww == w:
                coef[:, s + w] = ux[:, i]
        m = threej_band(j1, j2, w, coef, math.sqrt((2 * j1 + 1) * (2 * j2 + 1)))
        m[np.abs(m) < thr] = 0
        out[w] = {c: csr_matrix(m[ic]) for ic, c in enumerate(cart) if np.any(m[ic] != 0)}
    return out


def k_primitive(tens, J1, J2, thresh=None):
    """{omega: array (2J1+1 x 2J2+1)} -- primitive K factors over |J,k> of the molecular-frame tensor `tens` (vector of 3 or
    3x3 matrix), before the contraction with the rotor eigenvectors."""
    thr = _EPS if thresh is None else thresh
    tens = np.asarray(tens, dtype=np.float64)
    us, ux, os_, cart = cart_to_spher(tens.ndim)
    ust = us @ tens.reshape(-1)
    j1, j2 = int(round(J1)), int(round(J2))
    out = {}
    for w in sorted(set(w for w, _ in os_)):
        if abs(j1 - j2) > w or j1 + j2 < w:
            continue
        coef = np.zeros((1, 2 * w + 1), dtype=np.complex128)
        for i, (ww, s) in enumerate(os_):
            if ww == w and abs(ust[i]) > thr:
                coef[0, s + w] = ust[i]
        out[w] = threej_band(j1, j2, w, coef, 1.0)[0]
    return out
