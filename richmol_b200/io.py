"""Compact `.npz` container for `CarTens` objects (used for the committed test fixtures).

The reference persists tensors in HDF5 through h5py (`CarTens.store` / `CarTens.read`,
richmol/field.py:1304-1848); h5py is not available in this environment and HDF5 I/O is outside
the hot-path scope, so fixtures use this flat numpy schema instead:

    meta           json: rank, cart, os, blocks [(J, sym, dim_m, dim_k)], quanta_m, quanta_k
    k_keys         (nk, 5) objects   J1, J2, sym1, sym2, irrep
    k_<i>          dense K block
    m_keys         (nm, 6) objects   J1, J2, sym1, sym2, irrep, cart
    m_<i>_{data,indices,indptr,shape}   CSR M block
"""
import json

import numpy as np
from scipy.sparse import csr_matrix


def _jsonable(x):
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    return x


def save_cartens(path, tens):
    meta = dict(
        rank=int(tens.rank), cart=list(tens.cart), os=[list(map(int, o)) for o in tens.os],
        blocks=[[float(J), sym, int(tens.dim_m2[J][sym]), int(tens.dim_k2[J][sym])]
                for J in tens.Jlist2 for sym in tens.symlist2[J]],
        quanta_m=[_jsonable(tens.quanta_m2[J][sym]) for J in tens.Jlist2 for sym in tens.symlist2[J]],
        quanta_k=[_jsonable(tens.quanta_k2[J][sym]) for J in tens.Jlist2 for sym in tens.symlist2[J]],
    )
    arrays = {"meta": np.array(json.dumps(meta))}
    k_keys, m_keys = [], []
    for (J1, J2), kJ in tens.kmat.items():
        for (s1, s2), ks in kJ.items():
            for irrep, val in ks.items():
                arrays[f"k_{len(k_keys)}"] = val.toarray() if hasattr(val, "toarray") else np.asarray(val)
                k_keys.append([float(J1), float(J2), s1, s2, int(irrep)])
    for (J1, J2), mJ in tens.mmat.items():
        for (s1, s2), ms in mJ.items():
            for irrep, mc in ms.items():
                for cart, val in mc.items():
                    m = csr_matrix(val)
                    i = len(m_keys)
                    arrays[f"m_{i}_data"], arrays[f"m_{i}_indices"] = m.data, m.indices
                    arrays[f"m_{i}_indptr"], arrays[f"m_{i}_shape"] = m.indptr, np.array(m.shape)
                    m_keys.append([float(J1), float(J2), s1, s2, int(irrep), cart])
    arrays["k_keys"] = np.array(json.dumps(k_keys))
    arrays["m_keys"] = np.array(json.dumps(m_keys))
    np.savez_compressed(path, **arrays)


def load_cartens(path, name=None):
    from .field import CarTens
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    t = CarTens()
    t.rank, t.cart = meta["rank"], list(meta["cart"])
    t.os = [tuple(o) for o in meta["os"]]
    Jlist, symlist, dim_m, dim_k, qm, qk = [], {}, {}, {}, {}, {}
    for (J, sym, dm, dk), m_, k_ in zip(meta["blocks"], meta["quanta_m"], meta["quanta_k"]):
        if J not in symlist:
            Jlist.append(J)
            symlist[J], dim_m[J], dim_k[J], qm[J], qk[J] = [], {}, {}, {}, {}
        symlist[J].append(sym)
        dim_m[J][sym], dim_k[J][sym] = dm, dk
        qm[J][sym] = list(m_)
        qk[J][sym] = [tuple(q) for q in k_]
    for side in ("1", "2"):
        setattr(t, "Jlist" + side, list(Jlist))
        setattr(t, "symlist" + side, {J: list(v) for J, v in symlist.items()})
        setattr(t, "dim_m" + side, {J: dict(v) for J, v in dim_m.items()})
        setattr(t, "dim_k" + side, {J: dict(v) for J, v in dim_k.items()})
        setattr(t, "dim" + side, {J: {s: dim_m[J][s] * dim_k[J][s] for s in symlist[J]} for J in Jlist})
        setattr(t, "quanta_m" + side, {J: {s: list(v) for s, v in d.items()} for J, d in qm.items()})
        setattr(t, "quanta_k" + side, {J: {s: list(v) for s, v in d.items()} for J, d in qk.items()})
    t.kmat, t.mmat = {}, {}
    for i, (J1, J2, s1, s2, irrep) in enumerate(json.loads(str(z["k_keys"]))):
        t.kmat.setdefault((J1, J2), {}).setdefault((s1, s2), {})[irrep] = csr_matrix(z[f"k_{i}"])
    for i, (J1, J2, s1, s2, irrep, cart) in enumerate(json.loads(str(z["m_keys"]))):
        m = csr_matrix((z[f"m_{i}_data"], z[f"m_{i}_indices"], z[f"m_{i}_indptr"]),
                       shape=tuple(z[f"m_{i}_shape"]))
        t.mmat.setdefault((J1, J2), {}).setdefault((s1, s2), {}).setdefault(irrep, {})[cart] = m
    return t
