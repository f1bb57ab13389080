"""`TDSE` -- time propagation of wavepackets / thermal ensembles, B200-native.

Mirrors `richmol.tdse.TDSE` (richmol/tdse.py:27-414): same constructor keywords, `time_grid`,
`init_state`, and `update(H, vecs, H0=, matvec_lib=, propag=, tol=)` returning `(vecs2, time)`.
`update` runs the reference's split-operator step and its in-house Lanczos exponential
(`_expmv_lanczos`, richmol/tdse.py:417-486) for all states at once on the GPU through
`rmb_propagate_step` (include/richmol_b200.h); recurrences and the stopping rule are the
reference's, so per-state iteration counts match.

`vecs` may be
  * a numpy array `(nstates, N)` complex128 -> a new numpy array is returned (host <-> device
    copies inside the call, like any numpy caller of the reference), or
  * a CUDA `torch.Tensor` of the same shape/dtype -> a CUDA tensor is returned and the ensemble
    stays resident in HBM between steps (pass `inplace=True` to update it in place).
"""
import functools

import numpy as np
import scipy.constants as const
from scipy.sparse import csr_matrix, diags, issparse

from . import _lib, convert_units
from .field import CarTens, _stream_ptr


def update_counter(func):
    """Counts calls and returns the upper edge of the current time interval with the result
    (richmol/tdse.py:11-24)."""
    @functools.wraps(func)
    def wrapper(self, *args, **kwargs):
        vecs = func(self, *args, **kwargs)
        name = func.__name__
        icall = self._ncalls.setdefault(name, 0)
        time = self._time_grid[1][icall]
        self._ncalls[name] += 1
        return vecs, time
    return wrapper


def _as_cartens(obj, cache):
    if isinstance(obj, CarTens):
        return obj
    if hasattr(obj, "kmat") and (hasattr(obj, "mmat") or hasattr(obj, "mfmat")):
        # a richmol.field.CarTens: adopt it (static operators are re-packed when their mfmat changes)
        # the entry keeps `obj` and the dictionaries alive, so their ids cannot be recycled while it is live
        mf, km = getattr(obj, "mfmat", None), obj.kmat
        hit = cache.get("adopt")
        if hit is None or hit[0] is not obj or hit[1] is not mf or hit[2] is not km:
            hit = (obj, mf, km, CarTens.from_richmol(obj))
            cache["adopt"] = hit
        return hit[3]
    raise TypeError(f"bad operator type: '{type(obj)}'")


class TDSE():
    """Class for time-propagator (see richmol/tdse.py:27-143 for the keyword semantics)."""

    def __init__(self, **kwargs):
        self._ncalls = dict()   # per instance (the reference shares one dict across instances)
        self._cache = dict()

        if 't_start' in kwargs:
            assert (type(kwargs['t_start']) in [int, float]), \
                f"starting time `t_start` has bad type: '{type(kwargs['t_start'])}', (use 'int', 'float')"
            self._t_start = kwargs['t_start']
        else:
            self._t_start = 0

        assert ('t_end' in kwargs), "terminal time `t_end` not found in kwargs"
        assert (type(kwargs['t_end']) in [int, float]), \
            f"terminal time `t_end` has bad type: '{type(kwargs['t_end'])}', (use 'int', 'float')"
        assert (kwargs['t_end'] > self._t_start), \
            f"terminal time `t_end` has bad value: '{kwargs['t_end']}', (must be > '{self._t_start}')"
        self._t_end = kwargs['t_end']

        assert ('dt' in kwargs), "time step `dt` not found in kwargs"
        assert (type(kwargs['dt']) in [int, float]), \
            f"time step `dt` has bad type: '{type(kwargs['dt'])}', (use 'int', 'float')"
        assert (kwargs['dt'] <= (self._t_end - self._t_start)), \
            f"time step `dt` has bad value : '{kwargs['dt']}', (must be <= '{self._t_end - self._t_start}')"
        self._dt = kwargs['dt']

        t_units = {'fs': 1e-15, 'ps': 1e-12, 'ns': 1e-9, 'au': const.value("atomic unit of time")}
        if 't_units' in kwargs:
            assert (type(kwargs['t_units']) == str), \
                f"time units `t_units` has bad type: '{type(kwargs['t_units'])}', (use 'str')"
            assert (kwargs['t_units'] in t_units), \
                f"time units `t_units` has bad value: '{kwargs['t_units']}', (use 'fs', 'ps', 'ns', 'au')"
            self._t_to_s = t_units[kwargs['t_units']]
        else:
            self._t_to_s = t_units['ps']

        enr_units = {'invcm': 1 / convert_units.J_to_invcm(),
                     'mhz': convert_units.MHz_to_invcm() / convert_units.J_to_invcm()}
        if 'enr_units' in kwargs:
            assert (type(kwargs['enr_units']) == str), \
                f"energy units `enr_units` has bad type: '{type(kwargs['enr_units'])}', (use 'str')"
            assert (kwargs['enr_units'].lower() in enr_units), \
                f"energy units `enr_units` has bad value: '{kwargs['enr_units']}', (use 'invcm', 'MHz')"
            self._enr_to_J = enr_units[kwargs['enr_units'].lower()]
        else:
            self._enr_to_J = enr_units['invcm']

    # ------------------------------------------------------------------------------------------
    def time_grid(self, grid_type='equidistant'):
        """Generates time-grid to propagate on; returns interval centres (richmol/tdse.py:146-174)."""
        assert (grid_type.lower() in ['equidistant']), \
            f"time grid type `grid_type` has bad value: '{grid_type}', (use 'equidistant')"
        grid_size = int((self._t_end - self._t_start) / self._dt)
        t_1 = np.linspace(self._t_start, self._t_end, num=grid_size, endpoint=False)
        t_2 = t_1 + self._dt
        t_c = t_1 + self._dt / 2
        self._time_grid = (t_1, t_2, t_c)
        return t_c

    # ------------------------------------------------------------------------------------------
    def init_state(self, H, temp=None, thresh=1e-3, zpe=None, sparse=False, device_eigh=False):
        """Initial state vectors: eigenfunctions of `H`, Boltzmann-weighted for temp > 0
        (richmol/tdse.py:177-262).  Rows are sqrt(w_i)-scaled; the kept prefix is cut in basis
        order by cumulative weight.  Host-side, once per run.

        `device_eigh=True` (extension): the dense diagonalisation of a field-dressed `H`
        (richmol/tdse.py:231-233, `np.linalg.eigh` of the full matrix) runs on the GPU.  Eigenvalues agree to
        rounding; eigenvectors are only defined up to a phase (and a rotation inside degenerate subspaces), so
        the rows span the same eigenspaces but are not bit-identical to the host path."""
        if temp is not None:
            assert (temp >= 0), f"temperature `temp` has negative value: '{temp}'"
        assert (thresh >= 0), f"partition function threshold `thresh` has negative or zero value: '{thresh}'"

        H = _as_cartens(H, self._cache)
        H_is_diag = False
        try:
            if H.cart[0] == "0":
                H.field([0, 0, 1])
                H_is_diag = True
        except AttributeError:
            pass

        if not H._has_field():
            raise AttributeError("hamiltonian `H` has inappropriate type (must be Hamiltonian)") from None
        if H_is_diag:
            # field([0,0,1]) on a rank-0 tensor gives MF = 1 * M['0'] (field.py:1099), so the
            # potential equals the '0' component; taking it from the host tables keeps this
            # set-up step independent of the device
            enrs = H.tomat(form="full", repres="csr_matrix", cart="0").diagonal()
            vecs = diags(np.ones(len(enrs)), format='csr')
        else:
            hmat = H.tomat(form="full", repres="dense")
            if device_eigh:
                import torch
                _lib.require_device()
                e_d, v_d = torch.linalg.eigh(torch.from_numpy(np.ascontiguousarray(hmat)).cuda())
                enrs, vecs = e_d.cpu().numpy(), v_d.cpu().numpy()
            else:
                enrs, vecs = np.linalg.eigh(hmat)
            vecs = csr_matrix(vecs)
        enrs = np.array(enrs.real if np.iscomplexobj(enrs) else enrs, dtype=np.float64)

        if zpe is not None:
            assert (zpe <= abs(enrs[0])), \
                f"zero-point energy `zpe` has a value that is larger than the lowest energy: " \
                f"zpe = '{zpe}' > emin = '{enrs[0]}'"
        else:
            zpe = enrs[0]

        enrs -= zpe
        vecs = vecs.transpose().tocsr()
        if temp is None:
            pass
        elif temp == 0:
            vecs = vecs.getrow(0)
        else:
            enrs *= self._enr_to_J
            beta = 1.0 / (const.value("Boltzmann constant") * temp)
            weights = np.exp(-beta * enrs)
            weights /= np.sum(weights)
            if len(weights) <= 20000:
                csum = np.array([np.sum(weights[: i + 1]) for i in range(len(weights))])
            else:
                csum = np.cumsum(weights)
            inds = [i for i in range(len(weights)) if (1 - csum[i]) > thresh]
            sqrt_weights = np.sqrt(weights[inds])
            vecs = vecs[inds].multiply(sqrt_weights[:, None])

        if not sparse:
            vecs = vecs.toarray()
        return vecs.astype(np.complex128)

    # ------------------------------------------------------------------------------------------
    def _exp_fac(self):
        key = (self._dt, self._t_to_s, self._enr_to_J)
        hit = self._cache.get("exp_fac")
        if hit is None or hit[0] != key:
            hit = (key, -1j * self._dt * self._t_to_s * self._enr_to_J / const.value("reduced Planck constant"))
            self._cache["exp_fac"] = hit
        return hit[1]

    def _h0_phase(self, H0, exp_fac):
        """exp(exp_fac/2 * diag(H0)), cached on first use like the reference (richmol/tdse.py:368-373)."""
        if '_exp_fac_H0' not in self.__dict__:
            H0_mat = H0.tomat(form='full', cart='0')
            assert ((H0_mat - diags(H0_mat.diagonal())).nnz == 0), \
                "field-free Hamiltonian `H0` is not diagonal -> split-operator approach not implemented"
            self._exp_fac_H0 = np.exp(exp_fac / 2 * H0_mat.diagonal())
        return self._exp_fac_H0

    def _h0_phase_device(self, device):
        import torch
        hit = self._cache.get("phase_dev")
        if hit is None or hit[0] is not self._exp_fac_H0 or hit[1].device != device:
            t = torch.from_numpy(np.ascontiguousarray(self._exp_fac_H0, dtype=np.complex128)).to(device)
            hit = (self._exp_fac_H0, t)
            self._cache["phase_dev"] = hit
        return hit[1]

    @property
    def last_orders(self):
        """Per-state index of the last Lanczos iteration of the latest `update` (= matvecs - 1).  For
        device-resident ensembles the values are copied back asynchronously; reading them waits for
        the step to finish."""
        o = getattr(self, "_orders", None)
        if isinstance(o, tuple):
            pin, nst, stream = o
            if isinstance(stream, int):                   # raw cudaStream_t of the step (no Stream object per step)
                import torch
                stream = torch.cuda.ExternalStream(stream) if stream else torch.cuda.default_stream()
            stream.synchronize()
            o = pin[:nst].numpy().copy()
            self._orders = o
        return o

    @update_counter
    def update(self, H, vecs, **kwargs):
        """Propagates vectors by one time-step (richmol/tdse.py:265-414).

        Kwargs: `H0` (field-free Hamiltonian -> split-operator step), `matvec_lib` (accepted,
        ignored: the CUDA path is the only one), `propag` ('internal'; 'external' = Expokit raises
        NotImplementedError), `tol` (default 1e-15), and the extensions
        `inplace` (CUDA tensors only), `out` (numpy result buffer, e.g. pinned memory) and `expect`
        (numpy path: list of observables whose per-state expectation values of the propagated states are
        left in `tdse.last_expect[(iobs, istate)]`, computed on the device before the download)."""
        import ctypes as C

        if 'H0' in kwargs:
            H0 = kwargs['H0']
            assert (isinstance(H0, CarTens) or (hasattr(H0, 'kmat') and hasattr(H0, 'tomat'))), \
                f"field-free Hamiltonian `H0` is of bad (sub-)class: '{type(H0)}', " \
                f"(must be (sub-class of) '{CarTens}')"
        else:
            H0 = None

        if 'propag' in kwargs:
            assert (type(kwargs['propag']) is str), \
                f"propagator 'propag' has bad type: '{type(kwargs['propag'])}', (must be 'str')"
            assert (kwargs['propag'] in ['external', 'internal']), \
                f"propagator 'propag' has bad value: '{kwargs['propag']}', (use 'external', 'internal')"
            if kwargs['propag'] == 'external':
                # richmol/tdse.py:385-392,405-412 -> pyexpokit.zhexpv (Fortran ZHEXPV: m = 12, adaptive
                # sub-steps, Pade(6,6)).  A different algorithm with different iterates; answering it with the
                # internal Lanczos would silently return another result under the reference's own keyword.
                raise NotImplementedError(
                    "propag='external' (Expokit ZHEXPV) is not provided by richmol_b200; use the default "
                    "propag='internal' (richmol's in-house Lanczos, reproduced to 1e-10)")

        if 'tol' in kwargs:
            assert (type(kwargs['tol']) in [int, float]), \
                f"tolerance `tol` has bad type: '{type(kwargs['tol'])}', (must be 'int', 'float')"
            assert (kwargs['tol'] > 0 and kwargs['tol'] <= 1), \
                f"tolerance `tol` has bad value: '{kwargs['tol']}', (must be > 0 and <= 1)"
            tol = kwargs['tol']
        else:
            tol = 1e-15

        exp_fac = self._exp_fac()
        H = _as_cartens(H, self._cache)
        lib = _lib.lib()
        stream = _stream_ptr()

        # with H0 the Krylov part only runs if the tensor has a (non-empty) mfmat (tdse.py:377);
        # without H0 the reference needs mfmat and fails in CarTens.vec otherwise
        has_field = H._has_field()
        if H0 is not None:
            phase = self._h0_phase(H0, exp_fac)
            skip = (not has_field) or H._krylov_skippable()
        else:
            phase = None
            skip = False
            if not has_field:
                raise AttributeError(
                    "you need to multiply tensor with field before applying it to a vector")
        N = H._basis().N

        op = None if skip else H._device(stream)
        if op is None:
            # phases only: any handle of this basis will do; build a product-free operator
            op = self._cache.get("phase_only_op")
            if op is None or op.N != N:
                from .field import DeviceOperator
                op = DeviceOperator(H._basis(), [])
                self._cache["phase_only_op"] = op
        self._orders = None

        try:
            import torch
            is_tensor = isinstance(vecs, torch.Tensor)
        except ImportError:   # pragma: no cover
            is_tensor = False

        if is_tensor:
            if not vecs.is_cuda or vecs.dtype != torch.complex128 or vecs.dim() != 2:
                raise TypeError("device `vecs` must be a 2D CUDA tensor of dtype complex128")
            if vecs.shape[1] != N:
                raise ValueError(f"vecs has {vecs.shape[1]} columns, the basis has dimension {N}")
            out = vecs if kwargs.get('inplace', False) else vecs.clone()
            if not out.is_contiguous():
                raise ValueError("device `vecs` must be contiguous")
            nst = out.shape[0]
            # pinned buffer: the order copy is enqueued, not waited for (see `last_orders`)
            pin = self._cache.get("orders_pin")
            if pin is None or pin.numel() < nst:
                pin = torch.zeros(max(nst, 1), dtype=torch.int32).pin_memory()
                self._cache["orders_pin"] = pin
            ph_ptr = self._h0_phase_device(out.device).data_ptr() if phase is not None else None
            status = lib.rmb_propagate_step(
                op.handle, out.data_ptr(), nst, N, exp_fac.real, exp_fac.imag, float(tol), 100,
                ph_ptr, int(skip), pin.data_ptr(), stream)
            self._orders = (pin, nst, int(stream))
            _lib.check(status)
            return out

        if issparse(vecs):
            vecs = vecs.toarray()
        vin = np.ascontiguousarray(vecs, dtype=np.complex128)
        if vin.ndim != 2 or vin.shape[1] != N:
            raise ValueError(f"vecs must have shape (nstates, {N}), got {vin.shape}")
        nst = vin.shape[0]
        vout = kwargs.get('out')
        if vout is None:
            vout = np.empty_like(vin)
        elif (not isinstance(vout, np.ndarray) or vout.shape != vin.shape or vout.dtype != np.complex128
              or not vout.flags.c_contiguous):
            raise ValueError("`out` must be a C-contiguous complex128 array of the shape of `vecs`")
        orders = np.zeros(nst, dtype=np.int32)
        ph_dev = None
        if phase is not None:
            import torch
            ph_dev = self._h0_phase_device(torch.device("cuda", torch.cuda.current_device()))   # cached on the device
        # extension: observables of the propagated states evaluated on the device before the download
        obs = [_as_cartens(O, {}) for O in kwargs.get('expect', ())]
        for O in obs:
            if not O._has_field() and getattr(O, "cart", [None])[0] == "0":
                O.field([0, 0, 1])
        obs_ops = [O._device(stream) for O in obs]      # referenced until the call returns (LRU eviction)
        handles = (C.c_void_p * max(1, len(obs)))(*[o.handle for o in obs_ops])
        expv = np.zeros((len(obs), nst), dtype=np.complex128)
        status = lib.rmb_propagate_step_host_obs(
            op.handle, vin.ctypes.data, vout.ctypes.data, nst, N, exp_fac.real, exp_fac.imag,
            float(tol), 100, ph_dev.data_ptr() if ph_dev is not None else None, int(skip),
            orders.ctypes.data, len(obs), handles, expv.ctypes.data, stream)
        self._orders = orders
        self.last_expect = expv
        _lib.check(status)
        return vout


    # ------------------------------------------------------------------------------------------
    def propagate(self, terms, vecs, H0=None, tol=1e-15, expect=(), every=1):
        """Extension: many `update` steps in one call (the time loop of examples/ocs_alignment.py:89-100
        without returning to Python after every step).

        terms : list of (tensor, fields, thresh) -- the Hamiltonian is the sum of the tensors;
                `fields` is an array (nsteps, 3) of field vectors at the interval centres for a
                time-dependent term (thresh as in `CarTens.field`), or None for a static term whose field
                was applied beforehand (e.g. the dc part of ocs_mixed_field.py).
        expect : observables evaluated every `every` steps on the propagated states.
        Returns (vecs, times, expvals[nsteps // every, len(expect), nstates]); `vecs` lives where the input
        lives (numpy array or CUDA tensor).  Each step is exactly one `update(sum(terms), vecs, H0=H0)`."""
        import ctypes as C
        import torch
        from .field import device_operator
        from .packing import field_products
        terms = [(_as_cartens(t, {}), f, th) for (t, f, th) in terms]
        nsteps = None
        for _, f, _ in terms:
            if f is not None:
                f = np.asarray(f, dtype=np.float64)
                if f.ndim != 2 or f.shape[1] < 3:
                    raise IndexError("fields must be an array (nsteps, 3) of X, Y, Z components")
                if nsteps is not None and len(f) != nsteps:
                    raise ValueError("all time-dependent terms need the same number of steps")
                nsteps = len(f)
        if nsteps is None:
            raise ValueError("at least one term must be time dependent")
        first = terms[0][0]
        basis = first._basis()
        parts, dyn, static_states = [], [], []
        for it, (t, f, th) in enumerate(terms):
            if t._basis().key() != basis.key():
                raise ValueError("tensors defined with respect to different basis sets")
            p = t._parts()
            if len(p) != 1:
                raise TypeError("terms must be plain tensors (not sums)")
            parts.append(p[0][0])
            if f is None:
                if p[0][1] is None:
                    raise AttributeError("a static term needs its field applied beforehand")
                static_states.append(p[0][1])
            else:
                dyn.append(it)
                static_states.append(None)
        stream = _stream_ptr()
        op = device_operator(basis, parts)
        ndyn = len(dyn)
        fprod = np.zeros((nsteps, ndyn, 16), dtype=np.float64)
        dropped = np.zeros((nsteps, ndyn), dtype=np.int32)
        thresh = np.zeros(ndyn, dtype=np.float64)
        for j, it in enumerate(dyn):
            t, f, th = terms[it]
            thresh[j] = 0.0 if th is None else float(th)
            if len(t.cart) > 16:
                raise NotImplementedError("tensors with more than 16 Cartesian components")
            f = np.asarray(f, dtype=np.float64)
            for i in range(nsteps):
                fp, dr = field_products(t.cart, f[i], th)
                fprod[i, j, :len(fp)] = fp
                dropped[i, j] = int(dr)
        # static parts: apply their field once; dynamic parts get a placeholder state (overwritten per step)
        from .field import FieldState
        op.apply_fields([fs if fs is not None else FieldState(np.zeros(len(terms[k][0].cart)), None, True)
                         for k, fs in enumerate(static_states)], stream)
        for it in dyn:
            op._applied[it] = -1          # the C loop leaves its own field in these parts
        exp_fac = self._exp_fac()
        phase_dev = None
        is_tensor = isinstance(vecs, torch.Tensor)
        if is_tensor:
            if not vecs.is_cuda or vecs.dtype != torch.complex128 or vecs.dim() != 2:
                raise TypeError("device `vecs` must be a 2D CUDA tensor of dtype complex128")
            work = vecs.clone().contiguous()
        else:
            work = torch.from_numpy(np.ascontiguousarray(vecs, dtype=np.complex128)).cuda()
        if work.shape[1] != basis.N:
            raise ValueError(f"vecs has {work.shape[1]} columns, the basis has dimension {basis.N}")
        if H0 is not None:
            self._h0_phase(_as_cartens(H0, self._cache), exp_fac)
            phase_dev = self._h0_phase_device(work.device)
        obs = [_as_cartens(O, {}) for O in expect]
        for O in obs:
            if not O._has_field() and getattr(O, "cart", [None])[0] == "0":
                O.field([0, 0, 1])
        obs_ops = [O._device(stream) for O in obs]      # referenced until the call returns (LRU eviction)
        handles = (C.c_void_p * max(1, len(obs)))(*[o.handle for o in obs_ops])
        nst = work.shape[0]
        nout = nsteps // max(1, every)
        expv = torch.zeros((max(nout, 1), max(len(obs), 1), nst), dtype=torch.complex128, device=work.device)
        dyn_arr = np.array(dyn, dtype=np.int32)
        orders = torch.zeros(max(nst, 1), dtype=torch.int32).pin_memory()
        status = _lib.lib().rmb_propagate_many(
            op.handle, work.data_ptr(), nst, basis.N, nsteps, exp_fac.real, exp_fac.imag, float(tol), 100,
            phase_dev.data_ptr() if phase_dev is not None else None, ndyn, _lib.ptr(dyn_arr, C.c_int32),
            _lib.ptr(fprod, C.c_double), _lib.ptr(thresh, C.c_double), _lib.ptr(dropped, C.c_int32),
            len(obs), handles, max(1, every), expv.data_ptr(), orders.data_ptr(), stream)
        self._orders = (orders, nst, torch.cuda.current_stream(work.device))
        _lib.check(status)
        icall = self._ncalls.setdefault('update', 0)
        times = np.array([self._time_grid[1][icall + i] for i in range(nsteps)])
        self._ncalls['update'] = icall + nsteps
        expv = expv[:nout, :len(obs)]
        if is_tensor:
            return work, times, expv
        return work.cpu().numpy(), times, expv.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# K5: observables on device-resident ensembles (user code in examples/ocs_alignment.py:99-100,
# examples/ocs_mixed_field.py:112-117, tests/test_tdse.py:66 does this on the host per state)
# ---------------------------------------------------------------------------------------------
def expectation(O, vecs):
    """Per-state <v_i| O |v_i> (complex, shape (nstates,)).  `O` is a CarTens with a field applied
    (rank-0 observables such as cos2theta: `O.field([0, 0, 1])`).  `vecs`: numpy array or CUDA
    tensor; the result lives where `vecs` lives."""
    import torch
    if not isinstance(O, CarTens):
        O = CarTens.from_richmol(O)
    try:
        if not hasattr(O, "mfmat") and O.cart[0] == "0":
            O.field([0, 0, 1])
    except AttributeError:
        pass
    is_tensor = isinstance(vecs, torch.Tensor)
    v = vecs if is_tensor else torch.from_numpy(np.ascontiguousarray(vecs, dtype=np.complex128)).cuda()
    if not v.is_cuda or v.dtype != torch.complex128 or v.dim() != 2 or not v.is_contiguous():
        raise TypeError("`vecs` must be a contiguous 2D complex128 array / CUDA tensor")
    stream = _stream_ptr()
    op = O._device(stream)
    if v.shape[1] != op.N:
        raise ValueError(f"vecs has {v.shape[1]} columns, the basis has dimension {op.N}")
    out = torch.empty(v.shape[0], dtype=torch.complex128, device=v.device)
    _lib.check(_lib.lib().rmb_expectation(op.handle, v.data_ptr(), v.shape[0], v.shape[1],
                                          out.data_ptr(), stream))
    return out if is_tensor else out.cpu().numpy()


def populations(vecs):
    """Ensemble populations sum_i |v_i[j]|^2 (shape (N,), float64)."""
    import torch
    is_tensor = isinstance(vecs, torch.Tensor)
    v = vecs if is_tensor else torch.from_numpy(np.ascontiguousarray(vecs, dtype=np.complex128)).cuda()
    if not v.is_cuda or v.dtype != torch.complex128 or v.dim() != 2 or not v.is_contiguous():
        raise TypeError("`vecs` must be a contiguous 2D complex128 array / CUDA tensor")
    out = torch.empty(v.shape[1], dtype=torch.float64, device=v.device)
    _lib.check(_lib.lib().rmb_populations(v.data_ptr(), v.shape[0], v.shape[1], v.shape[1],
                                          out.data_ptr(), _stream_ptr()))
    return out if is_tensor else out.cpu().numpy()
