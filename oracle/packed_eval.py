"""TEST INFRASTRUCTURE ONLY -- numpy evaluation of the *packed* block tables.

Evaluates exactly what the CUDA kernels are specified to compute from the tables of
`richmol_b200.packing.PackedPart` (field contraction K1 and block matvec K2), so that the packer
can be validated against the oracle (`oracle/port.py`) on a box without a GPU.
"""
import numpy as np


def contract_field(part, fstate):
    """MF entry values: sum_c fprod[c] * coef[c, e], element threshold (field.py:1122-1139)."""
    if fstate.all_dropped:
        return np.zeros(part.nent, dtype=np.complex128)
    val = (fstate.fprod[:, None] * part.ent_coef).sum(axis=0)
    if fstate.thresh > 0:
        val[np.abs(val) < fstate.thresh] = 0
    return val


def matvec(basis, parts, fstates, x):
    """y = sum_products (MF (x) K) x for a flat vector or a (nstates, N) batch."""
    x2 = np.atleast_2d(x)
    y = np.zeros_like(x2, dtype=np.complex128)
    for part, fs in zip(parts, fstates):
        val = contract_field(part, fs)
        for p in range(len(part.pr_bra)):
            b1, b2, t = int(part.pr_bra[p]), int(part.pr_ket[p]), int(part.pr_table[p])
            dm1, dk1 = int(basis.dm[b1]), int(basis.dk[b1])
            dm2, dk2 = int(basis.dm[b2]), int(basis.dk[b2])
            nd = int(part.tb_nd[t])
            e0 = int(part.tb_off[t])
            col = part.ent_col[e0:e0 + dm1 * nd].reshape(dm1, nd)
            mf = val[e0:e0 + dm1 * nd].reshape(dm1, nd)
            K = part.kpool[int(part.pr_koff[p]):int(part.pr_koff[p]) + dk1 * dk2].reshape(dk1, dk2)
            X = x2[:, basis.off[b2]:basis.off[b2 + 1]].reshape(-1, dm2, dk2)
            Z = np.zeros((x2.shape[0], dm1, dk2), dtype=np.complex128)
            for j in range(nd):
                ok = col[:, j] >= 0
                Z[:, ok, :] += mf[ok, j][None, :, None] * X[:, col[ok, j], :]
            y[:, basis.off[b1]:basis.off[b1 + 1]] += (Z @ K.T).reshape(x2.shape[0], -1)
    return y if np.ndim(x) == 2 else y[0]


def lin_entry_lists(basis, parts, fstates):
    """Specification of `k_lin_entries` (linear rotors, dim_k = 1 everywhere): per bra block the list of
    (ket block, diagonal offset, values[dm1]) with values = K * MF along one surviving diagonal of one block
    product; pairs that read the same ket block along the same diagonal are merged (their values added).
    A diagonal survives when any of its MF entries is non-zero after the field contraction (the table mask
    of `k_field_contract`); its offset is col - row of the first row whose entry lies inside the ket block."""
    nb = len(basis.dm)
    assert all(int(d) == 1 for d in basis.dk)
    lists = [[] for _ in range(nb)]
    for part, fs in zip(parts, fstates):
        val = contract_field(part, fs)
        for p in range(len(part.pr_bra)):
            b1, b2, t = int(part.pr_bra[p]), int(part.pr_ket[p]), int(part.pr_table[p])
            dm1, nd, e0 = int(basis.dm[b1]), int(part.tb_nd[t]), int(part.tb_off[t])
            col = part.ent_col[e0:e0 + dm1 * nd].reshape(dm1, nd)
            mf = val[e0:e0 + dm1 * nd].reshape(dm1, nd)
            k = part.kpool[int(part.pr_koff[p])]
            for j in range(nd):
                if not np.any(mf[:, j] != 0):
                    continue                                   # diagonal screened out by the field
                rows = np.nonzero(col[:, j] >= 0)[0]
                doff = int(col[rows[0], j] - rows[0]) if len(rows) else 0
                v = np.where(col[:, j] >= 0, k * mf[:, j], 0.0)
                for ent in lists[b1]:
                    if ent[0] == b2 and ent[1] == doff:
                        ent[2] = ent[2] + v                    # merged pair
                        break
                else:
                    lists[b1].append([b2, doff, v])
    return lists


def lin_matvec(basis, lists, x):
    """Specification of `k_matvec_lin`: y[b1][r] = sum_entries value[r] * x[ket][clamp(r + doff)] (the value is
    zero wherever the diagonal leaves the ket block, so the clamped element does not contribute)."""
    x2 = np.atleast_2d(x)
    y = np.zeros_like(x2, dtype=np.complex128)
    for b1, ents in enumerate(lists):
        dm1 = int(basis.dm[b1])
        r = np.arange(dm1)
        for b2, doff, v in ents:
            dm2 = int(basis.dm[b2])
            c = np.clip(r + doff, 0, dm2 - 1)
            y[:, basis.off[b1]:basis.off[b1 + 1]] += v[None, :] * x2[:, basis.off[b2] + c]
    return y if np.ndim(x) == 2 else y[0]
