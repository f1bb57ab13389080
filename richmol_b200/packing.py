"""Flattening of the CarTens nested-dict data model into the block tables of the C ABI.

The reference walks `kmat[(J1,J2)][(sym1,sym2)][irrep]` / `mfmat[...]` dictionaries on every
matvec (richmol/field.py:1212-1243); here that walk happens once, at packing time.

Layout produced (see include/richmol_b200.h, rmb_part_desc):
  * basis blocks in `for J in Jlist2 for sym in symlist2[J]` order (richmol/tdse.py:343-348),
    block index = im*dim_k + ik;
  * one *product* per (Jpair, sympair, irrep) key present in both the K and the M dictionaries;
  * K factors as dense row-major (dk1 x dk2) blocks in one pool (real if every K is real);
  * M factors as ELL tables over the union sparsity pattern of all Cartesian components, with one
    coefficient plane per Cartesian component; identical M tables (they only depend on
    (J1, J2, irrep), not on symmetry) are stored once.
"""
import hashlib

import numpy as np
import scipy.sparse as sp


class Basis:
    """(J, sym) block layout of the flat state vector."""

    def __init__(self, Jlist, symlist, dim_m, dim_k):
        self.blocks = [(J, sym) for J in Jlist for sym in symlist[J]]
        self.index = {b: i for i, b in enumerate(self.blocks)}
        self.dm = np.array([dim_m[J][sym] for J, sym in self.blocks], dtype=np.int32)
        self.dk = np.array([dim_k[J][sym] for J, sym in self.blocks], dtype=np.int32)
        self.off = np.zeros(len(self.blocks) + 1, dtype=np.int64)
        np.cumsum(self.dm.astype(np.int64) * self.dk.astype(np.int64), out=self.off[1:])
        self.N = int(self.off[-1])

    def key(self):
        k = self.__dict__.get("_key")
        if k is None:
            k = (tuple(self.blocks), self.dm.tobytes(), self.dk.tobytes())
            self.__dict__["_key"] = k
        return k

    @classmethod
    def of(cls, tens, side=2):
        s = str(side)
        return cls(getattr(tens, "Jlist" + s), getattr(tens, "symlist" + s),
                   getattr(tens, "dim_m" + s), getattr(tens, "dim_k" + s))


def _to_coo(mat):
    m = sp.coo_matrix(mat)
    m.sum_duplicates()
    return m


def _dense(mat):
    return mat.toarray() if sp.issparse(mat) else np.asarray(mat)


class PackedPart:
    """Immutable flat tables of one tensor operator (one term of a sum)."""

    def __init__(self):
        self.cart = []
        self.keys = []          # per product: (Jpair, sympair, irrep)
        self.table_shape = []   # per table: (dm1, dm2)

    # -- construction ------------------------------------------------------------------------
    @classmethod
    def build(cls, basis, kmat, mdict, cart, static=False):
        """`mdict` is `mmat` (leaf: ...[irrep][cart] -> matrix) or, with static=True, an `mfmat`
        (...[irrep] -> matrix) which is packed as a single pseudo-Cartesian component."""
        self = cls()
        self.cart = ["_mf"] if static else list(cart)
        ncart = len(self.cart)
        cart_index = {c: i for i, c in enumerate(self.cart)}
        pr_bra, pr_ket, pr_table, pr_koff = [], [], [], []
        kblocks, kcomplex = [], False
        koff = 0
        tables = {}   # digest -> table id
        tb_dm1, tb_dm2, tb_nd, tb_off = [], [], [], [0]
        ent_col, ent_coef = [], []
        for Jpair, kmat_J in kmat.items():
            if Jpair not in mdict:
                continue
            m_J = mdict[Jpair]
            for sympair, kmat_s in kmat_J.items():
                if sympair not in m_J:
                    continue
                m_s = m_J[sympair]
                b1 = basis.index.get((Jpair[0], sympair[0]))
                b2 = basis.index.get((Jpair[1], sympair[1]))
                if b1 is None or b2 is None:
                    continue
                dm1, dk1 = int(basis.dm[b1]), int(basis.dk[b1])
                dm2, dk2 = int(basis.dm[b2]), int(basis.dk[b2])
                for irrep, kval in kmat_s.items():
                    if irrep not in m_s:
                        continue
                    K = _dense(kval)
                    if K.shape != (dk1, dk2):
                        raise ValueError(
                            f"K matrix for {Jpair} {sympair} irrep {irrep} has shape {K.shape}, "
                            f"basis expects {(dk1, dk2)}")
                    mc = {"_mf": m_s[irrep]} if static else m_s[irrep]
                    coos = []
                    h = hashlib.blake2b(digest_size=16)
                    h.update(np.array([dm1, dm2], dtype=np.int64).tobytes())
                    for c in self.cart:
                        if c not in mc:
                            continue
                        m = _to_coo(mc[c])
                        if m.shape != (dm1, dm2):
                            raise ValueError(
                                f"M matrix for {Jpair} {sympair} irrep {irrep} cart {c} has shape "
                                f"{m.shape}, basis expects {(dm1, dm2)}")
                        coos.append((cart_index[c], m))
                        h.update(c.encode())
                        h.update(m.row.astype(np.int32).tobytes())
                        h.update(m.col.astype(np.int32).tobytes())
                        h.update(m.data.astype(np.complex128).tobytes())
                    dig = h.digest()
                    t = tables.get(dig)
                    if t is None:
                        t = len(tb_dm1)
                        tables[dig] = t
                        col, coef, nd = _ell_table(dm1, dm2, coos, ncart)
                        tb_dm1.append(dm1)
                        tb_dm2.append(dm2)
                        tb_nd.append(nd)
                        tb_off.append(tb_off[-1] + dm1 * nd)
                        ent_col.append(col)
                        ent_coef.append(coef)
                        self.table_shape.append((dm1, dm2))
                    pr_bra.append(b1)
                    pr_ket.append(b2)
                    pr_table.append(t)
                    pr_koff.append(koff)
                    koff += dk1 * dk2
                    if np.iscomplexobj(K) and np.any(K.imag != 0):
                        kcomplex = True
                    kblocks.append(K.reshape(-1))
                    self.keys.append((Jpair, sympair, irrep))
        self.ncart = ncart
        self.pr_bra = np.array(pr_bra, dtype=np.int32)
        self.pr_ket = np.array(pr_ket, dtype=np.int32)
        self.pr_table = np.array(pr_table, dtype=np.int32)
        self.pr_koff = np.array(pr_koff, dtype=np.int64)
        self.k_is_complex = bool(kcomplex)
        if kblocks:
            kp = np.concatenate(kblocks)
            kp = kp.astype(np.complex128) if kcomplex else np.ascontiguousarray(kp.real, dtype=np.float64)
        else:
            kp = np.zeros(0, dtype=np.float64)
        self.kpool = np.ascontiguousarray(kp)
        self.tb_dm1 = np.array(tb_dm1, dtype=np.int32)
        self.tb_dm2 = np.array(tb_dm2, dtype=np.int32)
        self.tb_nd = np.array(tb_nd, dtype=np.int32)
        self.tb_off = np.array(tb_off, dtype=np.int64)
        nent = int(tb_off[-1])
        self.ent_col = (np.concatenate(ent_col) if ent_col else np.zeros(0)).astype(np.int32)
        self.ent_coef = np.zeros((ncart, nent), dtype=np.complex128)
        for t, coef in enumerate(ent_coef):
            self.ent_coef[:, tb_off[t]:tb_off[t + 1]] = coef
        self.nent = nent
        self.static = static
        return self

    def scaled(self, factor):
        """Same tables with every K multiplied by `factor` (CarTens.mul, field.py:932-948).
        Memoised per factor: the per-step pattern `H = -0.5 * pol * E` of the reference's examples
        then maps to ONE packed part (and so to one device operator, which is keyed by part identity)
        instead of a new table upload every step."""
        memo = self.__dict__.setdefault("_scaled_memo", {})
        fkey = complex(factor)
        hit = memo.get(fkey)
        if hit is not None:
            return hit
        new = PackedPart()
        new.__dict__.update(self.__dict__)
        kp = self.kpool * factor
        if np.iscomplexobj(kp) and np.any(kp.imag != 0):
            new.kpool = np.ascontiguousarray(kp, dtype=np.complex128)
            new.k_is_complex = True
        else:
            new.kpool = np.ascontiguousarray(kp.real, dtype=np.float64)
            new.k_is_complex = False
        new.__dict__["_scaled_memo"] = {}
        if len(memo) >= 8:
            memo.pop(next(iter(memo)))
        memo[fkey] = new
        return new

    # -- helpers -----------------------------------------------------------------------------
    def mf_to_dict(self, values, drop_zero=True):
        """Rebuilds the nested `mfmat` dictionary from contracted entry values (nent complex)."""
        out = {}
        cache = {}
        for (Jpair, sympair, irrep), t in zip(self.keys, self.pr_table):
            t = int(t)
            if t not in cache:
                dm1, dm2 = self.table_shape[t]
                nd = int(self.tb_nd[t])
                sl = slice(int(self.tb_off[t]), int(self.tb_off[t + 1]))
                col = self.ent_col[sl]
                val = values[sl]
                rows = np.repeat(np.arange(dm1), nd)
                mask = col >= 0
                if drop_zero:
                    mask &= val != 0
                m = sp.csr_matrix((val[mask], (rows[mask], col[mask])), shape=(dm1, dm2))
                cache[t] = m
            m = cache[t]
            if m.nnz > 0 or not drop_zero:
                out.setdefault(Jpair, {}).setdefault(sympair, {})[irrep] = m
        return out


def _ell_table(dm1, dm2, coos, ncart):
    """Diagonal-aligned ELL table over the union pattern: slot j of every row holds the entry on
    diagonal (col - row) = diag[j] (or -1 where that diagonal leaves the matrix), so a diagonal that
    vanishes after the field contraction vanishes for the whole table.
    Returns (col[dm1*nd], coef[ncart, dm1*nd], nd)."""
    if coos:
        rows = np.concatenate([m.row.astype(np.int64) for _, m in coos])
        cols = np.concatenate([m.col.astype(np.int64) for _, m in coos])
    else:
        rows = cols = np.zeros(0, dtype=np.int64)
    diags = np.unique(cols - rows)
    nd = max(1, len(diags))
    col = np.full(dm1 * nd, -1, dtype=np.int32)
    coef = np.zeros((ncart, dm1 * nd), dtype=np.complex128)
    if len(diags):
        ent_all = rows * nd + np.searchsorted(diags, cols - rows)
        col[ent_all] = cols.astype(np.int32)
        for ci, m in coos:
            r, c = m.row.astype(np.int64), m.col.astype(np.int64)
            np.add.at(coef[ci], r * nd + np.searchsorted(diags, c - r), m.data)
    return col, coef, nd


_AXIS = {"x": 0, "y": 1, "z": 2}
_CART_IDX = {}


def field_products(cart, field, thresh):
    """Products of field components per Cartesian label with the product screening of
    CarTens.field (richmol/field.py:1094-1105).  Returns (fprod[ncart], all_dropped)."""
    try:
        fx, fy, fz = field[:3]
        f = (float(fx), float(fy), float(fz))
    except (TypeError, IndexError, ValueError):
        raise IndexError(
            "field variable must be an iterable with three items which represent field's X, Y, "
            "and Z components") from None
    key = tuple(cart)
    idx = _CART_IDX.get(key)
    if idx is None:
        idx = [None if c == "0" else tuple(_AXIS[ch] for ch in c) for c in cart]
        _CART_IDX[key] = idx
    fprod = [0.0] * len(idx)
    kept = 0
    for i, ax in enumerate(idx):
        val = 1.0
        if ax is not None:
            for a in ax:
                val *= f[a]
        if thresh is not None and not abs(val) >= thresh:
            continue
        fprod[i] = val
        kept += 1
    return np.array(fprod, dtype=np.float64), kept == 0
