"""`CarTens` -- laboratory-frame Cartesian tensor operator, B200-native.

Mirrors the public interface of the reference class `richmol.field.CarTens`
(richmol/field.py:25-1855) for everything on the TDSE hot path:

    field        richmol/field.py:1073-1142   -> K1 field-contraction kernel
    vec          richmol/field.py:1145-1245   -> K2 batched block matvec kernel
    mul          richmol/field.py:932-948
    add_cartens  richmol/field.py:951-1070    (lazy sum, frozen `mfmat`, renamed irrep keys)
    __mul__/__add__/__sub__                    richmol/field.py:1248-1301
    tomat / full_form                          richmol/field.py:449-569, 659-692 (host, set-up only)

The data model (attribute names, nested dictionaries of scipy CSR matrices) is the reference's
(richmol/field.py:58-170), so tensors built by richmol itself can be adopted with
`CarTens.from_richmol(obj)`.  The numerical work of `field`, `vec` and of the propagator runs on the
GPU through `librichmol_b200.so`; there is no CPU fallback.
"""
import copy
import itertools
from collections import OrderedDict
from collections.abc import Mapping

import numpy as np
import scipy.sparse as sp

from . import _lib
from .packing import Basis, PackedPart, field_products

_SCALARS = (int, float, complex, np.integer, np.floating, np.complexfloating)

_BASIS_ATTRS = ("Jlist1", "Jlist2", "symlist1", "symlist2", "dim1", "dim2", "dim_k1", "dim_k2",
                "dim_m1", "dim_m2")
_DATA_ATTRS = _BASIS_ATTRS + ("quanta_k1", "quanta_k2", "quanta_m1", "quanta_m2", "rank", "cart",
                              "os", "kmat", "mmat")


class FieldState:
    """Result of one `CarTens.field()` call: screened field products + element threshold."""
    _serial = itertools.count(1)

    def __init__(self, fprod, thresh, all_dropped):
        self.fprod = np.ascontiguousarray(fprod, dtype=np.float64)
        self.thresh = 0.0 if thresh is None else float(thresh)
        self.all_dropped = bool(all_dropped)
        self.serial = next(FieldState._serial)


class DeviceOperator:
    """Python owner of one `rmb_operator` handle (block tables resident in HBM)."""

    def __init__(self, basis, parts):
        import ctypes as C
        _lib.require_device()
        lib = _lib.lib()
        self.basis = basis
        self.parts = list(parts)      # PackedPart objects (kept alive: the cache is keyed by id)
        self.N = basis.N
        descs = (_lib.PartDesc * max(1, len(parts)))()
        keep = []
        for d, p in zip(descs, parts):
            coef = np.ascontiguousarray(p.ent_coef).view(np.float64)
            kp = p.kpool.view(np.float64) if p.k_is_complex else p.kpool
            keep += [coef, kp]
            d.ncart, d.nprod = p.ncart, len(p.pr_bra)
            d.pr_bra, d.pr_ket = _lib.ptr(p.pr_bra, C.c_int32), _lib.ptr(p.pr_ket, C.c_int32)
            d.pr_table, d.pr_koff = _lib.ptr(p.pr_table, C.c_int32), _lib.ptr(p.pr_koff, C.c_int64)
            d.k_is_complex = int(p.k_is_complex)
            d.kpool, d.kpool_len = _lib.ptr(kp, C.c_double), len(p.kpool)
            d.ntables = len(p.tb_dm1)
            d.tb_dm1, d.tb_dm2 = _lib.ptr(p.tb_dm1, C.c_int32), _lib.ptr(p.tb_dm2, C.c_int32)
            d.tb_nd, d.tb_off = _lib.ptr(p.tb_nd, C.c_int32), _lib.ptr(p.tb_off, C.c_int64)
            d.ent_col, d.ent_coef = _lib.ptr(p.ent_col, C.c_int32), _lib.ptr(coef, C.c_double)
        od = _lib.OperatorDesc()
        od.nblocks = len(basis.blocks)
        od.blk_off, od.blk_dm = _lib.ptr(basis.off, C.c_int64), _lib.ptr(basis.dm, C.c_int32)
        od.blk_dk = _lib.ptr(basis.dk, C.c_int32)
        od.nparts, od.parts = len(parts), descs
        h = C.c_void_p()
        _lib.check(lib.rmb_operator_create(C.byref(od), C.byref(h)))
        self.handle = h
        self._applied = [0] * len(parts)   # serial of the FieldState each part currently holds
        self._destroy = lib.rmb_operator_destroy

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._destroy(h)

    def apply_fields(self, fstates, stream=None):
        import ctypes as C
        lib = _lib.lib()
        for i, fs in enumerate(fstates):
            if fs is None:
                raise AttributeError(
                    "you need to multiply tensor with field before applying it to a vector")
            if self._applied[i] != fs.serial:
                _lib.check(lib.rmb_operator_set_field(
                    self.handle, i, _lib.ptr(fs.fprod, C.c_double), fs.thresh, int(fs.all_dropped), stream))
                self._applied[i] = fs.serial

    def get_mf(self, part, stream=None):
        lib = _lib.lib()
        n = lib.rmb_operator_nentries(self.handle, part)
        out = np.zeros(max(n, 0), dtype=np.complex128)
        if n > 0:
            _lib.check(lib.rmb_operator_get_mf(self.handle, part, out.ctypes.data, stream))
        return out

    def counters(self):
        import ctypes as C
        out = (C.c_int64 * 4)()
        _lib.check(_lib.lib().rmb_get_counters(self.handle, out))
        return dict(launches=out[0], matvec_launches=out[1], iterations=out[2], state_matvecs=out[3])


_OP_CACHE = OrderedDict()
_OP_CACHE_MAX = 16


def device_operator(basis, parts):
    """Returns the (cached) device operator for an ordered list of PackedPart objects."""
    key = tuple(id(p) for p in parts)
    op = _OP_CACHE.get(key)
    if op is None:
        op = DeviceOperator(basis, parts)
        _OP_CACHE[key] = op
        while len(_OP_CACHE) > _OP_CACHE_MAX:
            _OP_CACHE.popitem(last=False)
    else:
        _OP_CACHE.move_to_end(key)
    return op


def clear_device_cache():
    _OP_CACHE.clear()


def _copy_nested(d, depth):
    if depth == 0 or not isinstance(d, Mapping):
        return d
    return {k: _copy_nested(v, depth - 1) for k, v in d.items()}


class CarTens:
    """General class for laboratory-frame Cartesian tensor operator (B200 device-backed).

    Attributes follow richmol/field.py:58-170: `rank`, `cart`, `os`, `Jlist1/2`, `symlist1/2`,
    `dim1/2`, `dim_m1/2`, `dim_k1/2`, `quanta_m1/2`, `quanta_k1/2`, `kmat`, `mmat`, `mfmat`.
    """

    def __init__(self, filename=None, name=None, **kwargs):
        if filename is not None:
            from . import io as _io
            self.__dict__.update(_io.load_cartens(filename, name=name).__dict__)

    # -- adoption of reference objects ---------------------------------------------------------
    @classmethod
    def from_richmol(cls, obj):
        """Adopts a `richmol.field.CarTens` (or any object exposing the same data model).
        A frozen sum (no `mmat`, only `mfmat`; field.py:1012-1015) becomes a static operator."""
        self = cls()
        for a in _DATA_ATTRS:
            if hasattr(obj, a):
                v = getattr(obj, a)
                depth = {"kmat": 3, "mmat": 4}.get(a, 0)
                setattr(self, a, _copy_nested(v, depth) if depth else copy.copy(v))
        if hasattr(obj, "mfmat") and hasattr(obj, "mmat") and hasattr(obj, "_rmb_field"):
            self.field(*obj._rmb_field)
        elif hasattr(obj, "mfmat"):
            # a frozen sum, or a reference tensor after its own `H.field(...)` (the canonical loop of
            # examples/ocs_alignment.py:93-96): the contracted `mfmat` IS the operator; an empty one (every
            # field product screened out, field.py:1107-1112) makes the Krylov part skippable (tdse.py:377)
            self._static_mf = _copy_nested(obj.mfmat, 3)
            self._fstate = FieldState([1.0], None, len(obj.mfmat) == 0)
        return self

    # -- internals -------------------------------------------------------------------------------
    def _basis(self):
        b = self.__dict__.get("_basis_cache")
        if b is None:
            b = Basis.of(self, 2)
            if Basis.of(self, 1).key() != b.key():
                raise ValueError("bra and ket basis sets differ: the propagator needs bra == ket "
                                 "(richmol/tdse.py:344-358)")
            self.__dict__["_basis_cache"] = b
        return b

    def _parts(self):
        """[(PackedPart, FieldState, key suffix)] -- one entry for a plain tensor, several for a sum."""
        if "_sum_parts" in self.__dict__:
            return self._sum_parts
        p = self.__dict__.get("_packed")
        if p is None:
            if "_static_mf" in self.__dict__:
                p = PackedPart.build(self._basis(), self.kmat, self._static_mf, None, static=True)
            else:
                p = PackedPart.build(self._basis(), self.kmat, self.mmat, self.cart)
            self.__dict__["_packed"] = p
        return [(p, self.__dict__.get("_fstate"), "")]

    def _device(self, stream=None):
        parts = self._parts()
        op = device_operator(self._basis(), [p for p, _, _ in parts])
        op.apply_fields([fs for _, fs, _ in parts], stream)
        return op

    def _invalidate(self):
        for a in ("_packed", "_basis_cache", "_mfmat_cache"):
            self.__dict__.pop(a, None)

    def _has_field(self):
        """hasattr(self, 'mfmat') without touching the device."""
        return all(fs is not None for _, fs, _ in self._parts())

    def _krylov_skippable(self):
        """True when `mfmat` is known to be empty without asking the device (every field product
        screened out; richmol/field.py:1107-1112 and richmol/tdse.py:377)."""
        return all(fs is not None and fs.all_dropped for _, fs, _ in self._parts())

    # -- mfmat: contracted M factors, fetched from the device on demand ---------------------------
    @property
    def mfmat(self):
        parts = self._parts()
        if any(fs is None for _, fs, _ in parts):
            raise AttributeError("'CarTens' object has no attribute 'mfmat'")
        token = tuple(fs.serial for _, fs, _ in parts)
        cached = self.__dict__.get("_mfmat_cache")
        if cached is not None and cached[0] == token:
            return cached[1]
        out = {}
        if not all(fs.all_dropped for _, fs, _ in parts):
            op = self._device()
            for i, (p, fs, suffix) in enumerate(parts):
                if fs.all_dropped:
                    continue
                d = p.mf_to_dict(op.get_mf(i))
                for Jpair, dJ in d.items():
                    for sympair, ds in dJ.items():
                        tgt = out.setdefault(Jpair, {}).setdefault(sympair, {})
                        for irrep, m in ds.items():
                            tgt[_rename(irrep, suffix)] = m
        self.__dict__["_mfmat_cache"] = (token, out)
        return out

    # -- field --------------------------------------------------------------------------------
    def field(self, field, thresh=None):
        """In-place multiplication of tensor with field (richmol/field.py:1073-1142).

        The contraction MF = sum_cart (prod E_c) M_cart runs on the GPU the next time the operator is
        applied; product screening (|prod| < thresh dropped, also the "0" product) is done here."""
        if "_sum_parts" in self.__dict__ or ("_static_mf" in self.__dict__ and "mmat" not in self.__dict__):
            raise AttributeError("'CarTens' object has no attribute 'mmat'")   # field.py:1115 on a sum
        if "_static_mf" in self.__dict__:
            # adopted together with a contracted mfmat: a new field replaces it (field.py:1107)
            self.__dict__.pop("_static_mf")
            self.__dict__.pop("_packed", None)
        fprod, all_dropped = field_products(self.cart, field, thresh)
        self.__dict__["_fstate"] = FieldState(fprod, thresh, all_dropped)
        self.__dict__["_rmb_field"] = (list(field[:3]), thresh)
        self.__dict__.pop("_mfmat_cache", None)

    # -- vec ----------------------------------------------------------------------------------
    def vec(self, vec, matvec_lib='scipy'):
        """Computes product of tensor with vector (richmol/field.py:1145-1245) on the GPU.

        `vec[J][sym]` -> array of length dim2[J][sym]; returns the same structure over the bra
        blocks the tensor connects to.  `matvec_lib` is accepted for compatibility ('scipy',
        'numba', 'cupy' all map to the CUDA kernel)."""
        if any(fs is None for _, fs, _ in self._parts()):
            raise AttributeError(
                "you need to multiply tensor with field before applying it to a vector")
        if not isinstance(vec, Mapping):
            raise TypeError(f"bad argument type for `vec`: '{type(vec)}'")
        assert (matvec_lib in ['scipy', 'numba', 'cupy']), \
            f"bad argument for `matvec_lib`: '{matvec_lib}' (must be 'scipy', 'numba', 'cupy')"
        import torch
        basis = self._basis()
        x = np.zeros(basis.N, dtype=np.complex128)
        present = set()
        for i, (J, sym) in enumerate(basis.blocks):
            try:
                x[basis.off[i]:basis.off[i + 1]] = np.asarray(vec[J][sym]).reshape(-1)
                present.add(i)
            except KeyError:
                pass
        op = self._device()
        xd = torch.from_numpy(x).cuda()
        yd = torch.empty_like(xd)
        _lib.check(_lib.lib().rmb_matvec(op.handle, xd.data_ptr(), yd.data_ptr(), 1, basis.N,
                                         _stream_ptr()))
        y = yd.cpu().numpy()
        touched = set()
        for p, fs, _ in self._parts():
            if fs.all_dropped:
                continue
            for b1, b2 in zip(p.pr_bra, p.pr_ket):
                if int(b2) in present:
                    touched.add(int(b1))
        out = {}
        for i in sorted(touched):
            J, sym = basis.blocks[i]
            out.setdefault(J, {})[sym] = y[basis.off[i]:basis.off[i + 1]].copy()
        return out

    # -- scalar multiplication ------------------------------------------------------------------
    def mul(self, arg):
        """In-place multiplication of tensor with a scalar `arg` (richmol/field.py:932-948)."""
        if isinstance(arg, bool) or not isinstance(arg, _SCALARS):
            raise TypeError(f"bad argument type for `arg` : '{type(arg)}'") from None
        if "_sum_parts" in self.__dict__:
            self.__dict__["_sum_parts"] = [(p.scaled(arg), fs, sfx) for p, fs, sfx in self._sum_parts]
            self.__dict__.pop("_kmat_sum", None)
            return
        packed = self.__dict__.get("_packed")
        # fresh dictionaries: sums built earlier keep the old K (snapshot semantics of field.py:1029-1036)
        self.kmat = {
            Jpair: {sympair: {key: val * arg for key, val in ks.items()} for sympair, ks in kJ.items()}
            for Jpair, kJ in self.kmat.items()
        }
        if packed is not None:
            self.__dict__["_packed"] = packed.scaled(arg)

    # -- sums -----------------------------------------------------------------------------------
    def add_cartens(self, arg):
        """Adds two tensors together (richmol/field.py:951-1070): a lazy sum whose `mfmat` is frozen
        at the time of the addition and whose irrep keys are renamed '<irrep>_1' / '<irrep>_2'."""
        if not isinstance(arg, CarTens):
            raise TypeError(f"bad argument type for `arg`: '{type(arg)}'") from None
        # same check as the reference (field.py:967-986), done on the cached block layout first
        b1, b2 = self.__dict__.get("_basis_cache"), arg.__dict__.get("_basis_cache")
        if b1 is None or b2 is None or (b1 is not b2 and b1.key() != b2.key()):
            bad = [a for a in _BASIS_ATTRS if getattr(self, a) != getattr(arg, a)]
            if bad:
                raise ValueError(
                    f"tensors defined with respect to different basis sets (differ in {bad})") from None
        for t in (self, arg):
            try:
                if t.cart[0] == "0":
                    t.field([0, 0, 1])
            except AttributeError:
                pass
        res = CarTens()
        res.__dict__.update({k: v for k, v in self.__dict__.items() if not k.startswith("_")})
        for a in ("os", "rank", "cart", "mmat", "kmat"):
            res.__dict__.pop(a, None)
        parts = []
        for t, sfx in ((self, "_1"), (arg, "_2")):
            for p, fs, s in t._parts():
                if fs is None:
                    raise AttributeError("'CarTens' object has no attribute 'mfmat'")
                parts.append((p, fs, s + sfx))
        res.__dict__["_sum_parts"] = parts
        res.__dict__["_kmat_src"] = [(self.kmat, "_1"), (arg.kmat, "_2")]
        if "_basis_cache" in self.__dict__:
            res.__dict__["_basis_cache"] = self.__dict__["_basis_cache"]
        return res

    def __getattr__(self, name):
        # `kmat` of a sum is assembled lazily (merged dictionaries with renamed irrep keys)
        if name == "kmat" and "_kmat_src" in self.__dict__:
            merged = self.__dict__.get("_kmat_sum")
            if merged is None:
                merged = {}
                for src, sfx in self.__dict__["_kmat_src"]:
                    for Jpair, kJ in src.items():
                        for sympair, ks in kJ.items():
                            tgt = merged.setdefault(Jpair, {}).setdefault(sympair, {})
                            for irrep, val in ks.items():
                                tgt[_rename(irrep, sfx)] = val
                self.__dict__["_kmat_sum"] = merged
            return merged
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    def _clone(self):
        res = CarTens()
        for k, v in self.__dict__.items():
            if k in ("kmat", "_static_mf"):
                res.__dict__[k] = _copy_nested(v, 3)
            elif k == "mmat":
                res.__dict__[k] = _copy_nested(v, 4)
            elif k == "_mfmat_cache":
                continue
            elif k in ("_packed", "_basis_cache"):
                res.__dict__[k] = v      # immutable: shared, so clones map to the same device operator
            else:
                res.__dict__[k] = copy.copy(v)
        return res

    def __mul__(self, arg):
        """Multiplication with a scalar (`mul`) or with a field (`field`) -- richmol/field.py:1248-1275."""
        if isinstance(arg, _SCALARS) and not isinstance(arg, bool):
            res = self._clone()
            res.mul(arg)
        elif isinstance(arg, (np.ndarray, list, tuple)):
            res = self._clone()
            res.field(arg)
        elif isinstance(arg, dict):
            res = self._clone()   # the reference discards the product as well (field.py:1263-1266)
            res.vec(arg)
        else:
            raise TypeError(
                f"unsupported operand type(s) for '*': '{self.__class__.__name__}' and "
                f"'{type(arg)}'") from None
        return res

    def __add__(self, arg):
        if isinstance(arg, CarTens):
            return self.add_cartens(arg)
        raise TypeError(
            f"unsupported operand type(s) for '+': '{self.__class__.__name__}' and "
            f"'{type(arg)}'") from None

    def __sub__(self, arg):
        if isinstance(arg, CarTens):
            return self.add_cartens(arg * (-1))
        raise TypeError(
            f"unsupported operand type(s) for '-': '{self.__class__.__name__}' and "
            f"'{type(arg)}'") from None

    __rmul__ = __mul__
    __radd__ = __add__
    __rsub__ = __sub__

    # -- matrix representation (set-up only: observables, H0 diagonal, init_state) ---------------
    def tomat(self, form='block', repres='csr_matrix', thresh=None, cart=None):
        """Matrix representation of the tensor (richmol/field.py:449-569): `cart=None` gives the
        potential sum_irrep kron(MF, K) (MF fetched from the device), otherwise the given
        Cartesian component sum_irrep kron(M_cart, K)."""
        assert (form in ('block', 'full')), f"`form` unknown: '{form}' (use 'block', 'full')"
        if cart is None:
            try:
                mdict = self.mfmat
            except AttributeError:
                raise AttributeError(
                    "specify Cartesian component `cart` of tensor or multiply tensor with field "
                    "before computing its its matrix representation") from None
            pick = lambda m_s: m_s
        else:
            if cart not in self.cart:
                raise ValueError(
                    f"specified Cartesian component '{cart}' is not contained in tensor "
                    f"components '{self.cart}'") from None
            mdict = self.mmat
            pick = lambda m_s: {irrep: val[cart] for irrep, val in m_s.items() if cart in val}
        kmat = self.kmat
        mat = {}
        for Jpair in mdict.keys() & kmat.keys():
            for sympair in mdict[Jpair].keys() & kmat[Jpair].keys():
                mm = pick(mdict[Jpair][sympair])
                kk = kmat[Jpair][sympair]
                terms = [sp.kron(mm[irrep], kk[irrep]) for irrep in mm.keys() & kk.keys()]
                if terms:
                    me = terms[0]
                    for t in terms[1:]:
                        me = me + t
                    mat.setdefault(Jpair, {})[sympair] = me
        if thresh is not None and thresh > 0:
            for mat_J in mat.values():
                for sympair, m in mat_J.items():
                    m = m.tocoo()
                    keep = np.abs(m.data) > thresh
                    mat_J[sympair] = sp.csr_matrix((m.data[keep], (m.row[keep], m.col[keep])), shape=m.shape)
        if form == 'block':
            for mat_J in mat.values():
                for sympair, m in mat_J.items():
                    mat_J[sympair] = m.toarray() if repres == 'dense' else getattr(sp, repres)(m)
            return mat
        return self.full_form(mat, repres, thresh)

    def full_form(self, mat, repres='csr_matrix', thresh=None):
        """Block representation -> 2D matrix (richmol/field.py:659-692).  Assembled from the COO triplets of the
        blocks that exist (the reference builds a dense zero matrix for every missing block pair, which at
        N ~ 10^6 is the most expensive thing it ever does)."""
        roff, coff, ind = {}, {}, 0
        for J1 in self.Jlist1:
            for sym1 in self.symlist1[J1]:
                roff[(J1, sym1)] = ind
                ind += self.dim1[J1][sym1]
        nrow, ind = ind, 0
        for J2 in self.Jlist2:
            for sym2 in self.symlist2[J2]:
                coff[(J2, sym2)] = ind
                ind += self.dim2[J2][sym2]
        ncol = ind
        rows, cols, vals = [], [], []
        for (J1, J2), mat_J in mat.items():
            for (sym1, sym2), m in mat_J.items():
                if (J1, sym1) not in roff or (J2, sym2) not in coff:
                    continue
                m = sp.coo_matrix(m)
                rows.append(m.row.astype(np.int64) + roff[(J1, sym1)])
                cols.append(m.col.astype(np.int64) + coff[(J2, sym2)])
                vals.append(m.data)
        if vals:
            res = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                shape=(nrow, ncol))
        else:
            res = sp.coo_matrix((nrow, ncol))
        return res.toarray() if repres == 'dense' else getattr(sp, repres)(res)


def _rename(irrep, suffix):
    return irrep if not suffix else str(irrep) + suffix


def _stream_ptr():
    """cudaStream_t of torch's current stream on the current device (the raw-handle query: `torch.cuda.current_stream()`
    builds a Stream object every call, ~13 us, three times per single-state step)."""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


def filter(obj, bra=lambda **kw: True, ket=lambda **kw: True, thresh=None):
    """State filters are applied by the tensor sources (`richmol_b200.synth`, `from_richmol`);
    the reference's module-level `filter` (richmol/field.py:1859) is a loader-side utility and is
    out of scope of the hot path."""
    raise NotImplementedError("state filters are applied when the tensor is built (see synth.py)")
