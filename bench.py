#!/usr/bin/env python
"""Benchmark of the TDSE hot path (BASELINE.json metric: ensemble state-timesteps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload h2s|h2o|ocs_align|ocs_mixed|ocs_batch|asym] [--also LIST|none]

One *step* of every workload is the loop body of the reference's examples
(examples/ocs_alignment.py:89-100, examples/ocs_mixed_field.py:102-117):

    for every time-dependent term:  term.field(E(t), thresh)          # CarTens.field
    vecs, t = tdse.update(sum(terms), vecs, H0=h0)                    # split-operator Lanczos step
    <cos^2 theta> of the ensemble                                     # (+ one all-reduce at N > 1)

The headline workload (default `h2s`) is the largest single-GPU BASELINE configuration (configs[3]); the
JSON line also carries a `workloads` array with the same measurements (value, e2e, roofline, parity
self-check) for the other BASELINE configurations, selected with `--also` (default: all of them):

    ocs_align  configs[0]  OCS linear rotor Jmax=30, one state (T = 0), 1e10 V/m 800 nm pulse      (replicas only)
    h2o        configs[1]  H2O Watson-A rigid rotor Jmax=20, 500-state Boltzmann shard per GPU     (weak)
    ocs_mixed  configs[2]  OCS Jmax=60, tilted dc field + ac pulse, dc-dressed states at 1 K       (strong)
    h2s        configs[3]  H2S optical centrifuge Jmax=80 (N = 708 561), 64 states per GPU         (weak)
    asym       configs[4]  TROVE-style asymmetric top J <= 100 (N = 1 020 100), fixed 2048-state batch (strong)
    ocs_batch  (extra)     OCS Jmax=60 with an 8192-state batch: the HBM-bound linear-rotor matvec  (strong)

All inputs are synthetic (richmol_b200/synth.py); arithmetic is complex128 throughout.  After the timed
loop every workload re-propagates its first steps from the initial rows and compares a sample of rows with
the CPU oracle (`parity`: max relative error, equality of the per-state Lanczos orders).

`--impl reference` times the UNMODIFIED reference (richmol.field.CarTens.field + richmol.tdse.TDSE.update,
byte-compiled into oracle/_ref by oracle/build_ref.py; the numpy port oracle/port.py if that is missing) for
the same step on all host cores, each step a bounded sample of the ensemble.
"""
import argparse
import json
import os
import pickle
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "state-timesteps/sec"
UNIT = "state-steps/s"
DT = 0.01            # ps
C_LIGHT = 299792458.0
OMEGA_800 = 2 * np.pi * C_LIGHT / 800e-9 * 1e-12      # 1/ps


# ------------------------------------------------------------------------------------------------
# workload definitions (shared by the GPU arm and the CPU arms)
# ------------------------------------------------------------------------------------------------
class Workload:
    """name, BASELINE config index, scaling mode, states, and the three callables below.

    build()        -> dict(h0, cos2, terms=[dict(name, tensor, static_field | None, thresh)])
    field(name, i) -> field vector of the time-dependent term `name` at step i
    rows(m, lo, hi)-> initial rows [lo, hi) of the ensemble (numpy, complex128)
    """
    name = ""
    config = None
    scaling = "weak"          # weak: `nstates` per GPU; strong: `nstates` in total; replicas: every rank all
    nstates = 1
    cpu_states_per_core = 1
    check_rows = 3
    text = ""

    def total_states(self, world):
        return self.nstates * world if self.scaling == "weak" else self.nstates

    def bounds(self, rank, world):
        from richmol_b200.ensemble import shard_bounds
        if self.scaling == "weak":
            return rank * self.nstates, (rank + 1) * self.nstates
        if self.scaling == "replicas":
            return 0, self.nstates
        return shard_bounds(self.nstates, rank, world)


def _boltzmann_rows(h0, lo, hi, temp):
    """Rows [lo, hi) of the Boltzmann ensemble sqrt(w_i)|i> in basis order (what TDSE.init_state builds,
    richmol/tdse.py:250-257), without the prefix truncation."""
    import scipy.constants as const
    from richmol_b200 import convert_units as cu
    enr = h0.tomat(form="full", cart="0").diagonal().real
    enr = enr - enr[0]
    w = np.exp(-enr / cu.J_to_invcm() / (const.value("Boltzmann constant") * temp))
    w /= w.sum()
    N = len(enr)
    idx = np.arange(lo, hi) % N
    v = np.zeros((hi - lo, N), dtype=np.complex128)
    v[np.arange(hi - lo), idx] = np.sqrt(w[idx])
    return v


def _basis_rows(N, lo, hi, seed):
    """Unit basis states |J,k,m> at seeded random positions (rows lo..hi of one fixed sequence)."""
    idx = np.random.default_rng(seed).integers(0, N, size=max(hi, 1))[lo:hi]
    v = np.zeros((hi - lo, N), dtype=np.complex128)
    v[np.arange(hi - lo), idx] = 1.0
    return v


class OcsAlign(Workload):
    name, config, scaling, nstates = "ocs_align", 0, "replicas", 1
    check_rows = 1
    text = ("ocs_align: OCS linear rotor Jmax=30 N={N}, one state (T=0), -1/2 alpha:EE with the 1e10 V/m 800 nm "
            "Gaussian pulse of examples/ocs_alignment.py (FWHM 10 ps, window at the peak), thresh 1e3")

    def build(self):
        from richmol_b200 import convert_units as cu, synth
        m = synth.ocs(30)
        H = -1 / 2 * m["pol"] * cu.AUpol_x_Vm_to_invcm()
        return dict(h0=m["h0"], cos2=m["cos2"], terms=[dict(name="ac", tensor=H, static=None, thresh=1e3)])

    def field(self, name, i):
        fwhm = 10.0
        t = 11.0 + (i + 0.5) * DT                       # t0 = 12.5 ps: the window sits on the rising edge / peak
        t0 = 2.5 * fwhm / 2
        return [0, 0, 1e10 * np.exp(-4 * np.log(2) * (t - t0) ** 2 / fwhm ** 2) * np.cos(OMEGA_800 * t)]

    def rows(self, m, lo, hi):
        v = np.zeros((1, m["h0"]._basis().N), dtype=np.complex128)
        v[0, 0] = 1.0                                   # init_state(h0, temp=0): the ground state
        return v[lo:hi]


class H2O(Workload):
    name, config, scaling, nstates = "h2o", 1, "weak", 500
    cpu_states_per_core = 4
    text = ("h2o: H2O rigid rotor Watson-A Jmax=20 N={N}, {S}-state Boltzmann (300 K) shard per GPU, dc dipole "
            "(tilted 35 deg, ramp) + z-polarised ac polarisability")

    def build(self):
        from richmol_b200 import convert_units as cu, synth
        m = synth.h2o(20)
        return dict(h0=m["h0"], cos2=m["cos2"], terms=[
            dict(name="dc", tensor=m["dip"] * (-cu.AUdip_x_Vm_to_invcm()), static=None, thresh=None),
            dict(name="ac", tensor=m["pol"] * (-0.5 * cu.AUpol_x_Vm_to_invcm()), static=None, thresh=1e1)])

    def field(self, name, i):
        t = (i + 0.5) * DT
        if name == "dc":
            beta = 35.0 * np.pi / 180.0
            ramp = 0.5 + 0.5 * min(1.0, i / 200.0)
            return list(5e6 * ramp * np.array([np.sin(beta), 0.0, np.cos(beta)]))
        return [0.0, 0.0, 3e9 * np.exp(-4 * np.log(2) * (t - 1.0) ** 2) * np.cos(OMEGA_800 * t)]

    def rows(self, m, lo, hi):
        return _boltzmann_rows(m["h0"], lo, hi, 300.0)


class OcsMixed(Workload):
    name, config, scaling, nstates = "ocs_mixed", 2, "strong", 64
    cpu_states_per_core = 1
    text = ("ocs_mixed: OCS Jmax=60 N={N}, static dc 20.7 kV/cm tilted 35 deg (contracted once) + 1.5e9 V/m ac pulse "
            "(examples/ocs_mixed_field.py, window at the peak, thresh 1e1), {S} dc-dressed eigenstates "
            "Boltzmann-weighted at 1 K (init_state(h0 + Hdc, temp=1), thresh lowered to keep {S} rows), ensemble "
            "split over the GPUs")

    def build(self):
        from richmol_b200 import convert_units as cu, synth
        m = synth.ocs(60)
        dc = 20.7 * 1000 * 100
        beta = 35.0 * np.pi / 180.0
        rot = np.array([[np.cos(beta), 0, np.sin(beta)], [0, 1, 0], [-np.sin(beta), 0, np.cos(beta)]])
        dc_field = list(np.dot(rot, [dc, 0, 0]))
        Hdc = -1 * m["dip"] * cu.AUdip_x_Vm_to_invcm()
        Hac = -0.5 * m["pol"] * cu.AUpol_x_Vm_to_invcm()
        return dict(h0=m["h0"], cos2=m["cos2"], cos=m["cos"], terms=[
            dict(name="dc", tensor=Hdc, static=dc_field, thresh=None),
            dict(name="ac", tensor=Hac, static=None, thresh=1e1)])

    def field(self, name, i):
        t = 690.0 + (i + 0.5) * DT
        t0, fwhm = 700.0, 2 * 600.0 / 2.5
        return [0, 0, 1.5e9 * np.exp(-4 * np.log(2) * (t - t0) ** 2 / fwhm ** 2) * np.cos(OMEGA_800 * t)]

    def rows(self, m, lo, hi):
        rows = m.get("_rows")
        if rows is None:
            # dressed states: eigenvectors of h0 + Hdc (richmol/tdse.py:231-233), Boltzmann weights at 1 K in
            # basis order of the eigenvalues; the first `nstates` rows
            rows = dressed_rows(m, self.nstates, 1.0)
            m["_rows"] = rows
        return rows[lo:hi]


class H2S(Workload):
    name, config, scaling, nstates = "h2s", 3, "weak", 64
    cpu_states_per_core = 1
    check_rows = 2
    text = ("h2s: H2S rigid asymmetric top Jmax=80 N={N}, optical centrifuge -1/2 alpha:EE with "
            "E = 3e9 V/m [cos(b t^2), sin(b t^2), 0] (complex MF, Delta m = 0, +-2), {S} unit basis states per GPU")

    def build(self):
        from richmol_b200 import convert_units as cu, synth
        m = synth.h2s(80)
        return dict(h0=m["h0"], cos2=m["cos2"], terms=[
            dict(name="ac", tensor=m["pol"] * (-0.5 * cu.AUpol_x_Vm_to_invcm()), static=None, thresh=None)])

    def field(self, name, i):
        b = 0.02
        return [3e9 * np.cos(b * i * i), 3e9 * np.sin(b * i * i), 0.0]

    def rows(self, m, lo, hi):
        return _basis_rows(m["h0"]._basis().N, lo, hi, 4)


class Asym(Workload):
    name, config, scaling, nstates = "asym", 4, "strong", 2048
    cpu_states_per_core = 1
    check_rows = 2
    text = ("asym: TROVE-style asymmetric top J<=100, 4 symmetries x 25 states per J, N={N}, rank-1 tensor with dense "
            "random real K (seed 0) and exact 3j M, tilted 1e8 V/m field rotating in XZ; fixed batch of {S} unit basis "
            "states split over the GPUs (BASELINE's 8192-state batch = 134 GB of Psi is cut to {S} so that the "
            "default run stays within minutes; --asym-states 8192 runs it, sub-batched)")

    def build(self):
        from richmol_b200 import convert_units as cu, synth
        m = synth.trove_style(100)
        return dict(h0=m["h0"], cos2=None, terms=[
            dict(name="dc", tensor=m["dip"] * (-cu.AUdip_x_Vm_to_invcm() * 10.0), static=None, thresh=None)])

    def field(self, name, i):
        a = 0.6 + 0.01 * i
        return [1e8 * np.sin(a), 0.0, 1e8 * np.cos(a)]

    def rows(self, m, lo, hi):
        return _basis_rows(m["h0"]._basis().N, lo, hi, 5)


class OcsBatch(Workload):
    name, config, scaling, nstates = "ocs_batch", None, "strong", 8192
    cpu_states_per_core = 4
    text = ("ocs_batch: OCS linear rotor Jmax=60 N={N}, fixed batch of {S} Boltzmann (300 K) rows split over the GPUs, "
            "tilted dc dipole ramp + ac polarisability (M-mixing): the HBM-bound H.Psi case (k_matvec_lin)")

    def build(self):
        from richmol_b200 import convert_units as cu, synth
        m = synth.ocs(60)
        return dict(h0=m["h0"], cos2=m["cos2"], terms=[
            dict(name="dc", tensor=m["dip"] * (-cu.AUdip_x_Vm_to_invcm()), static=None, thresh=None),
            dict(name="ac", tensor=m["pol"] * (-0.5 * cu.AUpol_x_Vm_to_invcm()), static=None, thresh=1e1)])

    field = H2O.field

    def rows(self, m, lo, hi):
        return _boltzmann_rows(m["h0"], lo, hi, 300.0)


WORKLOADS = {w.name: w for w in (OcsAlign, H2O, OcsMixed, H2S, Asym, OcsBatch)}
DEFAULT = "h2s"
ALSO_DEFAULT = ["ocs_align", "h2o", "ocs_mixed", "ocs_batch", "asym"]


def dressed_rows(m, nstates, temp):
    """First `nstates` rows of TDSE.init_state(h0 + Hdc, temp) (richmol/tdse.py:231-257: eigenvectors of the
    dressed Hamiltonian, rows scaled by sqrt(Boltzmann weight)) without the cumulative-weight cut.  Host-only
    (the matrices come from `tomat(cart=...)`), so both arms of the benchmark get the same rows."""
    import scipy.constants as const
    from richmol_b200 import convert_units as cu
    dc = [t for t in m["terms"] if t["static"] is not None][0]
    H = m["h0"].tomat(form="full", repres="dense", cart="0").astype(np.complex128)
    for c, f in zip("xyz", dc["static"]):
        if f != 0:
            H = H + f * dc["tensor"].tomat(form="full", repres="dense", cart=c)
    enr, vec = np.linalg.eigh(H)
    enr = (enr - enr[0]) / cu.J_to_invcm()
    wgt = np.exp(-enr / (const.value("Boltzmann constant") * temp))
    wgt /= wgt.sum()
    return np.ascontiguousarray((vec[:, :nstates] * np.sqrt(wgt[:nstates])[None, :]).T.astype(np.complex128))


def build_model(w):
    m = w.build()
    for t in m["terms"]:
        if t["static"] is not None:
            t["tensor"].field(list(t["static"]))                # contracted once (examples/ocs_mixed_field.py:91)
    if m.get("cos2") is not None:
        m["cos2"].field([0, 0, 1])
    return m


def hamiltonian(tensors):
    H = tensors[0]
    for t in tensors[1:]:
        H = H + t
    return H


# ------------------------------------------------------------------------------------------------
# CPU arms: the reference (oracle/_ref, unmodified) or its numpy port, rows split across processes
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_kind():
    from oracle import refshim
    return "reference" if refshim.root() is not None else "port"


def _cpu_init(path, kind):
    """Worker set-up: the model pickled by the parent -> tensors of the arm under test."""
    os.environ.setdefault("PYTHONHASHSEED", "0")
    with open(path, "rb") as f:
        wname, m = pickle.load(f)
    w = WORKLOADS[wname]()
    _W.update(w=w, kind=kind)
    if kind == "reference":
        from oracle import refshim
        r = refshim.load()
        conv = lambda t: refshim.to_reference(r, t)
        _W["h0"] = conv(m["h0"])
        _W["terms"] = []
        for t in m["terms"]:
            rt = conv(t["tensor"])
            if t["static"] is not None:
                rt.field(list(t["static"]))
            _W["terms"].append((t["name"], rt, t["static"] is None, t["thresh"]))
        tdse = r.tdse.TDSE(t_end=1e6, dt=DT)
        tdse._time_grid = (None, _Endless(DT), None)              # open-ended grid for the benchmark
        tdse._exp_fac_H0 = m["phase"]                             # the reference's own cache slot (tdse.py:368-373)
        _W["tdse"] = tdse
        if m.get("cos2mat") is not None:
            _W["cos2"] = m["cos2mat"]
    else:
        from oracle import port
        _W["port"] = port
        _W["h0"] = port.OracleTensor(m["h0"])
        _W["terms"] = []
        for t in m["terms"]:
            ot = port.OracleTensor(t["tensor"])
            if t["static"] is not None:
                ot.field(list(t["static"]))
            _W["terms"].append((t["name"], ot, t["static"] is None, t["thresh"]))
        if m.get("cos2mat") is not None:
            _W["cos2"] = m["cos2mat"]
        _W["fac"] = port.exp_factor(DT)
        _W["phase"] = m["phase"]
    return True


def _cpu_steps(args):
    """`nsteps` steps of the example loop on `rows`; returns (rows, per-step orders or None, seconds)."""
    rows, step0, nsteps, want_orders = args
    w = _W["w"]
    v = rows
    all_orders = []
    t0 = time.perf_counter()
    for s in range(step0, step0 + nsteps):
        for name, t, dyn, thresh in _W["terms"]:
            if dyn:
                t.field(w.field(name, s), thresh=thresh) if thresh is not None else t.field(w.field(name, s))
        if _W["kind"] == "reference":
            H = hamiltonian([t for _, t, _, _ in _W["terms"]])
            v, _ = _W["tdse"].update(H, H0=_W["h0"], vecs=v, matvec_lib="scipy")
        else:
            port = _W["port"]
            ts = [t for _, t, _, _ in _W["terms"]]
            H = ts[0]
            for t in ts[1:]:
                H = H.add(t)
            orders = [] if want_orders else None
            v = port.update_step(H, v, _W["fac"], phase=_W["phase"], orders=orders)
            all_orders.append(orders)
        if "cos2" in _W:
            sum(np.dot(np.conj(x), _W["cos2"].dot(x)) for x in v) + 1 / 3
    return v, (all_orders if want_orders else None), time.perf_counter() - t0


class _Endless:
    def __init__(self, dt):
        self.dt = dt

    def __getitem__(self, i):
        return (i + 1) * self.dt


def _dump_model(w, m):
    import types
    attrs = ("Jlist1", "Jlist2", "symlist1", "symlist2", "dim1", "dim2", "dim_k1", "dim_k2", "dim_m1", "dim_m2",
             "quanta_k1", "quanta_k2", "quanta_m1", "quanta_m2", "rank", "cart", "os", "kmat", "mmat")
    plain = lambda t: None if t is None else types.SimpleNamespace(**{a: getattr(t, a) for a in attrs if a in t.__dict__})
    fd, path = tempfile.mkstemp(prefix="rmb_model_", suffix=".pkl")
    slim = {k: plain(v) for k, v in m.items() if k in ("h0", "cos2")}
    slim["terms"] = [dict(t, tensor=plain(t["tensor"])) for t in m["terms"]]
    # One-time set-up quantities, computed here from the host tables and handed to the workers: the reference
    # builds them with `tomat(form='full')`, whose `full_form` allocates a DENSE zero matrix for every missing
    # block pair (richmol/field.py:677-686): N^2 * 8 bytes in total, i.e. hours at N = 708 561.  They are set-up,
    # not part of the step; the reference itself caches the phase vector in `_exp_fac_H0` (richmol/tdse.py:368-373).
    from oracle import port
    slim["phase"] = np.exp(port.exp_factor(DT) / 2 * m["h0"].tomat(form="full", cart="0").diagonal())
    slim["cos2mat"] = m["cos2"].tomat(form="full", cart="0") if m.get("cos2") is not None else None
    with os.fdopen(fd, "wb") as f:
        pickle.dump((w.name, slim), f, protocol=pickle.HIGHEST_PROTOCOL)
    return path


def cpu_run(w, m, steps, warmup, step0=0, one_core=True, kind=None):
    """Times `steps` steps of the reference algorithm on a sample of the ensemble using every host core (the
    reference's own scale-out pattern: rows of `vecs` split across processes,
    docs/source/notebooks/tdse_mpi.ipynb:268-272), and one process alone for the 1-core figure."""
    import multiprocessing as mp
    kind = kind or _cpu_kind()
    cores = os.cpu_count() or 1
    spc = w.cpu_states_per_core
    total = w.total_states(1)
    nproc = max(1, min(cores, total // spc if total >= spc else 1))
    spc = min(spc, total)
    nst = nproc * spc
    rows = w.rows(m, 0, nst)
    chunks = [rows[i * spc:(i + 1) * spc] for i in range(nproc)]
    path = _dump_model(w, m)
    ctx = mp.get_context("spawn")
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS",
                                            "PYTHONHASHSEED")}
    for k in saved:                      # one BLAS thread per worker process (inherited on spawn)
        os.environ[k] = "0" if k == "PYTHONHASHSEED" else "1"
    try:
        pool = ctx.Pool(nproc, initializer=_cpu_init, initargs=(path, kind))
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    one = None
    with pool:
        pool.map(_cpu_steps, [(c[:0], 0, 0, False) for c in chunks])          # workers up, models loaded
        if one_core:
            # one process alone (the others idle): the reference as a user runs it, single-threaded
            r = pool.apply(_cpu_steps, ((chunks[0][:1], step0, 1, False),))
            one = 1.0 / r[2]
        if warmup > 0:
            res = pool.map(_cpu_steps, [(c, step0, warmup, False) for c in chunks])
            chunks = [r[0] for r in res]
        t0 = time.perf_counter()
        pool.map(_cpu_steps, [(c, step0 + warmup, steps, False) for c in chunks])
        dt = time.perf_counter() - t0
    os.remove(path)
    what = ("UNMODIFIED reference: richmol CarTens.field + TDSE.update (byte code in oracle/_ref)" if kind == "reference"
            else "numpy/scipy port of richmol CarTens.field/vec + TDSE.update (oracle/port.py)")
    sample = (f"{nst} of {total} states x {steps} steps ({warmup} warm-up), {nproc} processes x {spc} states, {what}")
    return dict(value=nst * steps / dt, cores=nproc, kind=kind, sample=sample, ms_per_step=dt / steps * 1e3,
                one_core=one)


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks and throttle reasons DURING the timed region (B200_PROFILING.md): an NVML polling
    thread (10 ms period); falls back to an `nvidia-smi -lms` subprocess if pynvml is unavailable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.thread = self.proc = None
        self.stop_flag = False

    def _poll(self):
        import pynvml
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        while not self.stop_flag:
            self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            try:
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            time.sleep(0.01)

    def start(self):
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            # the physical index of the visible device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                self.index = int(vis.split(",")[self.index])
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            try:
                self.path = f"/tmp/rmb_clocks_{os.getpid()}.csv"
                self.f = open(self.path, "w")
                self.proc = subprocess.Popen(
                    ["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                     "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "20",
                     "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
            except OSError:
                self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.f.close()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                try:
                    self.samples.append(float(p[0]))
                    self.max_mhz = float(p[1])
                except (ValueError, IndexError):
                    continue
                for n, v in zip(names, p[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            os.remove(self.path)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
_PEAKS = {}


def measured_peaks():
    """HBM: MEASURED_PEAKS.json (driver-written) else the B200_PROFILING.md fallback.  FP64: neither file has
    an entry, so the DFMA / DMMA peaks are measured in this run on this GPU (rmb_fp64_peak)."""
    if _PEAKS:
        return _PEAKS
    import ctypes as C
    from richmol_b200 import _lib
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        _PEAKS.update(hbm_gbs=float(p["hbm_gbs"]), source="MEASURED_PEAKS.json (of measured)")
    except Exception:
        _PEAKS.update(hbm_gbs=6650.0, source="B200_PROFILING.md fallback (of fallback)")
    a, b = C.c_double(), C.c_double()
    _lib.check(_lib.lib().rmb_fp64_peak(C.byref(a), C.byref(b), None))
    _PEAKS.update(dfma_tflops=a.value, dmma_tflops=b.value, fp64_tflops=max(a.value, b.value),
                  fp64_source="rmb_fp64_peak: register-resident DFMA / DMMA loops measured in this run on this GPU "
                              "(MEASURED_PEAKS.json and B200_PROFILING.md have no FP64 entry)")
    return _PEAKS


def ncu_traffic(workload):
    """DRAM bytes (read + write) of one full-batch matvec launch from the committed ncu capture; None if there is none."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_matvec_traffic.json")))[workload]
        return int(t["dram_bytes_read"] + t["dram_bytes_write"])
    except Exception:
        return None


def op_info(op):
    """flops / operator bytes per state-matvec for the field currently applied (formulae of SURVEY.md 8d;
    only the M diagonals that survive the field contraction are counted, as in the reference's CSR), and the
    kernel routing of the operator."""
    import ctypes as C
    from richmol_b200 import _lib
    fl, by = C.c_double(), C.c_double()
    _lib.check(_lib.lib().rmb_operator_work(op.handle, C.byref(fl), C.byref(by), None))
    r = (C.c_int64 * 8)()
    _lib.check(_lib.lib().rmb_operator_info(op.handle, r))
    return {"flops_per_state": fl.value, "op_bytes": by.value, "tiled": r[0], "dmma": r[1], "scalar": r[2],
            "lin_T": r[3], "fused": r[4], "dk_max": r[5]}


def gpu_workload(w, args, headline, ctx):
    """Device-resident timing, roofline of the matvec, end-to-end timing and the parity self-check of one
    workload.  Returns the dict that becomes the headline line / one entry of `workloads`."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from richmol_b200 import TDSE, _lib
    from richmol_b200.tdse import expectation
    world, rank, dev = ctx["world"], ctx["rank"], ctx["dev"]
    lib = _lib.lib()

    t_build = time.perf_counter()
    m = build_model(w)
    h0, cos2 = m["h0"], m.get("cos2")
    N = h0._basis().N
    lo, hi = w.bounds(rank, world)
    rows = w.rows(m, lo, hi)
    nloc = len(rows)
    total = w.total_states(world) if w.scaling != "replicas" else w.nstates
    t_build = time.perf_counter() - t_build
    dyn = [t for t in m["terms"] if t["static"] is None]
    tensors = [t["tensor"] for t in m["terms"]]

    def new_tdse():
        tdse = TDSE(t_end=1e6, dt=DT)
        tdse._time_grid = (None, _Endless(DT), None)              # open-ended grid for the benchmark
        return tdse

    tdse = new_tdse()
    # ensemble <cos^2 theta>: per-step shard sums collected in a small device ring, all-reduced every OBS_EVERY steps (double
    # buffered, asynchronous): the path's only collective, off the critical path of the latency-bound workloads
    OBS_EVERY = 16
    obs_ring = [torch.zeros(OBS_EVERY, dtype=torch.complex128, device=dev) for _ in range(2)]
    obs_slot = [0, 0]                                              # ring in use, next slot
    pending = [None]

    def apply_fields(i):
        for t in dyn:
            if t["thresh"] is None:
                t["tensor"].field(w.field(t["name"], i))
            else:
                t["tensor"].field(w.field(t["name"], i), thresh=t["thresh"])

    def step(i, v):
        apply_fields(i)
        v, _ = tdse.update(hamiltonian(tensors), v, H0=h0, inplace=True)
        if cos2 is not None:
            ev = expectation(cos2, v)
            ring, slot = obs_ring[obs_slot[0]], obs_slot[1]
            torch.sum(ev, dim=0, keepdim=True, out=ring[slot:slot + 1])     # <cos^2 theta> - 1/3 of this shard
            obs_slot[1] = slot + 1
            if obs_slot[1] == OBS_EVERY:
                flush_obs()
        return v

    def flush_obs():
        if pending[0] is not None:
            pending[0].wait()                                      # the previous block's all-reduce, overlapped with 16 steps
            pending[0] = None
        if obs_slot[1] > 0 and world > 1 and w.scaling != "replicas":
            pending[0] = dist.all_reduce(torch.view_as_real(obs_ring[obs_slot[0]]), async_op=True)
        obs_slot[0] ^= 1
        obs_slot[1] = 0

    def barrier():
        flush_obs()
        if pending[0] is not None:
            pending[0].wait()
            pending[0] = None
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    vecs = torch.from_numpy(rows).to(dev)
    steps, warmup = args.steps, max(args.warmup, 3)
    for i in range(warmup):
        vecs = step(i, vecs)
    barrier()
    if not headline:
        # secondary workloads choose their own length: >= 200 steps or >= 0.5 s of timed work (<= 2000 steps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vecs = step(warmup, vecs)
        e1.record()
        barrier()
        est = torch.tensor([max(e0.elapsed_time(e1), 1e-3)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(est, op=dist.ReduceOp.MAX)
        steps = int(min(2000, max(3, min(200, 30e3 / est.item()), 500.0 / est.item())))
        warmup += 1
    op = hamiltonian(tensors)._device()
    ops = [op] + ([cos2._device()] if cos2 is not None else [])
    cnt0 = [o.counters() for o in ops]
    ms_, n_ = C.c_double(), C.c_int64()
    ms_o, n_o = C.c_double(), C.c_int64()
    for o in ops:
        lib.rmb_matvec_timing(o.handle, 1, C.byref(ms_), C.byref(n_))  # enable + reset
    sampler = ClockSampler(ctx["local"])
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warmup, warmup + steps):
        vecs = step(i, vecs)
    flush_obs()                                                    # the last (partial) block of observables is reduced inside the timed region
    if pending[0] is not None:
        pending[0].wait()
        pending[0] = None
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    lib.rmb_matvec_timing(op.handle, 0, C.byref(ms_), C.byref(n_))
    for o in ops[1:]:
        lib.rmb_matvec_timing(o.handle, 0, C.byref(ms_o), C.byref(n_o))   # the observable's matvec (<cos^2 theta>)
    cnt1 = [o.counters() for o in ops]
    launches = sum(b["launches"] - a["launches"] for a, b in zip(cnt0, cnt1))
    state_mv = cnt1[0]["state_matvecs"] - cnt0[0]["state_matvecs"]
    mv_launches = max(1, n_.value)
    counted = total if w.scaling != "replicas" else w.nstates * world
    value = counted * steps / (ms * 1e-3)

    # ---- roofline of the H.Psi matvec: 32*N bytes per state-matvec + operator bytes once per
    # launch; flops from the block tables
    info = op_info(op)
    peaks = measured_peaks()
    fp64_peak = peaks["fp64_tflops"]
    ai = info["flops_per_state"] / (32.0 * N)
    ridge = fp64_peak * 1e3 / peaks["hbm_gbs"]
    fused_step = bool(info["fused"]) and n_.value == 0           # the whole step is one k_lanczos_fused launch
    if fused_step:
        # no separate matvec launch to time: the algorithmic matvec work over the whole step time
        mv_s = ms * 1e-3
        state_mv = sum(int(o) + 1 for o in tdse.last_orders) * steps if tdse.last_orders is not None else 0
        mv_launches = steps
        kernel = ("k_lanczos_fused (the whole split-operator Lanczos step of a small linear rotor in ONE launch, one CTA per "
                  "state; latency-bound: the figures are the step's algorithmic matvec work over the whole step time)")
    else:
        mv_s = ms_.value * 1e-3
        if info["lin_T"] and nloc >= 4 * info["lin_T"]:
            kernel = "k_matvec_lin (H.Psi of a linear rotor: sliding window of ket blocks in shared memory, fused <w,V_k>)"
        elif info["dmma"] and info["tiled"]:
            kernel = (f"k_matvec_dmma (DMMA mma.sync.m8n8k4.f64, {info['dmma']} items with dim_k > 12) + k_matvec_tiled "
                      f"({info['tiled']} items): one H.Psi = both launches, timed together")
        elif info["dmma"]:
            kernel = "k_matvec_dmma (H.Psi with wide K blocks on the FP64 tensor pipe, DMMA mma.sync.m8n8k4.f64)"
        else:
            kernel = ("k_matvec_tiled (H.Psi: warpgroup-specialised, TMA-staged ket rows, fused MF(x)K block products, "
                      "fused <w,V_k>)")
    alg_bytes = 32.0 * N * state_mv + info["op_bytes"] * mv_launches
    alg_flops = info["flops_per_state"] * state_mv
    gbs = alg_bytes / mv_s / 1e9 if mv_s > 0 else 0.0
    tfs = alg_flops / mv_s / 1e12 if mv_s > 0 else 0.0
    hbm = {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
           "peak_source": peaks["source"]}
    fp64 = {"achieved": tfs, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfs / fp64_peak,
            "peak_source": peaks["fp64_source"], "dfma_tflops": peaks["dfma_tflops"],
            "dmma_tflops": peaks["dmma_tflops"]}
    # right of the ridge the FP64 pipe bounds the kernel ("tensor": on B200 the FP64 tensor pipe and the FMA pipe
    # have the same peak); linear rotors (dim_k = 1) are HBM-bound.  Both fractions are reported.
    compute_bound = ai > 1.25 * ridge
    head = fp64 if compute_bound else hbm
    roofline = {
        "kernel": kernel, "bound": "tensor" if compute_bound else "hbm", "achieved": head["achieved"],
        "peak": head["peak"], "unit": head["unit"], "frac": head["frac"], "traffic": ncu_traffic(w.name),
        "peak_source": head["peak_source"], "hbm": hbm, "fp64": fp64, "arithmetic_intensity": ai, "ridge": ridge,
        "flops_per_state_matvec": info["flops_per_state"], "bytes_per_state_matvec": 32.0 * N,
        "operator_bytes_per_launch": info["op_bytes"], "launches": int(mv_launches),
        "avg_launch_us": mv_s / mv_launches * 1e6,
        # H.Psi launches of the Hamiltonian plus the one of the observable (same kernels, another operator handle)
        "share_of_step": (mv_s * 1e3 + (0.0 if fused_step else ms_o.value)) / ms,
        "share_of_step_hamiltonian_only": mv_s * 1e3 / ms,
        "matvecs_per_state_step": state_mv / max(1, nloc * steps),
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE full-batch launch of this kernel from the committed "
                        "ncu --set full capture (profiles/r02_matvec_traffic.json, tools/round_profile.sh); not measured "
                        "inside the timed run (a run under ncu is never a bench value); null where no capture exists",
    }

    del vecs
    torch.cuda.empty_cache()
    # ---- end to end through the public API with HOST buffers (numpy in, numpy out)
    e2e = e2e_run(w, m, args, new_tdse(), rows, ctx, counted, steps)

    # ---- small ensembles: the same loop through the multi-step entry point (TDSE.propagate: one call, numpy in / out once,
    #      observable every step) -- what a user of examples/ocs_alignment.py would call instead of 30 000 update() calls
    multi = None
    if fused_step and rank == 0:
        try:
            nst_m = 2000
            tm = new_tdse()
            terms = [(t["tensor"], None if t["static"] is not None else np.array([w.field(t["name"], i) for i in range(nst_m)]),
                      t["thresh"]) for t in m["terms"]]
            tm.propagate(terms, rows, H0=h0, expect=[cos2] if cos2 is not None else [])       # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tm.propagate(terms, rows, H0=h0, expect=[cos2] if cos2 is not None else [])
            dtm = time.perf_counter() - t0
            multi = {"value": nloc * nst_m / dtm, "unit": UNIT, "steps": nst_m,
                     "note": "TDSE.propagate(terms, vecs, H0=, expect=[cos2]): all steps in one call, host arrays in and out"}
        except Exception as e:
            multi = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- parity self-check: the first steps again from the initial rows, a sample of rows against the oracle
    parity = parity_check(w, m, new_tdse(), rows, ctx) if rank == 0 and not args.no_parity else None
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    S = w.nstates
    return {
        "name": w.name, "baseline_config": w.config, "value": value, "unit": UNIT, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "scaling": {"replicas": "weak"}.get(w.scaling, w.scaling),
        "config": {"workload": w.text.format(N=N, S=S), "states_total": counted, "states_this_gpu": nloc,
                   "hilbert_dim": N, "dt_ps": DT,
                   "parallelism": ("replicas only (a single state does not shard)" if w.scaling == "replicas"
                                   else f"ensemble-sharded x{world} ({w.scaling})"),
                   "l2": f"Psi + Krylov vectors of this GPU: ~{nloc * N * 16 * 8 / 1e9:.2f} GB "
                         + ("(larger than the 126 MB L2)" if nloc * N * 16 * 8 > 126e6 else
                            "(SMALLER than the 126 MB L2: a latency-bound configuration, L2-resident by nature)"),
                   "model_build_s": round(t_build, 1)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "parity": parity,
        **({"multi_step_call": multi} if multi is not None else {}),
    }


def e2e_run(w, m, args, tdse, rows, ctx, counted, steps_dev):
    """Same step through TDSE.update with numpy arrays in pinned host memory: every step copies the
    ensemble host->device and the propagated ensemble + observable device->host."""
    import torch
    import torch.distributed as dist
    world, dev = ctx["world"], ctx["dev"]
    cos2 = m.get("cos2")
    dyn = [t for t in m["terms"] if t["static"] is None]
    tensors = [t["tensor"] for t in m["terms"]]
    full = len(rows)
    cap = max(1, int(2e9 // (rows.shape[1] * 16)))               # <= 2 GB per pinned buffer
    if full > cap:
        rows = rows[:cap]
        counted = counted * cap / full
    pin_a = torch.empty(rows.shape, dtype=torch.complex128).pin_memory()
    pin_b = torch.empty(rows.shape, dtype=torch.complex128).pin_memory()
    a, b = pin_a.numpy(), pin_b.numpy()
    a[...] = rows
    nbytes = rows.size * 16
    # ~1 s of work, at least 3 and at most 50 steps
    nsteps = int(max(3, min(50, steps_dev)))

    def step(i, src, dst):
        for t in dyn:
            if t["thresh"] is None:
                t["tensor"].field(w.field(t["name"], i))
            else:
                t["tensor"].field(w.field(t["name"], i), thresh=t["thresh"])
        tdse.update(hamiltonian(tensors), src, H0=m["h0"], out=dst,
                    expect=[cos2] if cos2 is not None else [])     # host array in -> host arrays out
        return complex(tdse.last_expect.sum())

    for i in range(2):
        step(i, a, b)
        a, b = b, a
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(2, 2 + nsteps):
        step(i, a, b)
        a, b = b, a
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return {"value": counted * nsteps / float(dt.item()), "unit": UNIT, "steps": nsteps,
            "states_this_gpu": len(rows),
            "h2d_bytes_per_step": int(nbytes + rows.shape[1] * 16),
            "d2h_bytes_per_step": int(nbytes + rows.shape[0] * 16 * (1 if cos2 is not None else 0)),
            "note": "numpy (pinned) in/out through TDSE.update(..., expect=[cos2]); includes the host-side "
                    "field products, the H2D/D2H copies of the ensemble (chunked, overlapped with the kernels) "
                    "and the per-step observable"}


def parity_check(w, m, tdse, rows, ctx, nsteps=2):
    """Re-propagates the first `nsteps` bench steps of the FULL batch of this GPU from the initial rows (same
    operator, same batch, same tiles as the timed loop) and compares `check_rows` sampled rows and their Lanczos
    orders with the CPU oracle port (tdse.py:417-486 semantics) on the same inputs."""
    import torch
    from oracle import port
    dev = ctx["dev"]
    nloc = len(rows)
    pick = sorted(set(int(x) for x in np.linspace(0, nloc - 1, min(w.check_rows, nloc))))
    dyn = [t for t in m["terms"] if t["static"] is None]
    tensors = [t["tensor"] for t in m["terms"]]
    ots = []
    for t in m["terms"]:
        ot = port.OracleTensor(t["tensor"])
        if t["static"] is not None:
            ot.field(list(t["static"]))
        ots.append(ot)
    fac = port.exp_factor(DT)
    phase = np.exp(fac / 2 * m["h0"].tomat(form="full", cart="0").diagonal())      # tdse.py:368-373
    v = torch.from_numpy(rows).to(dev)
    ref = rows[pick].copy()
    worst, orders_equal = 0.0, True
    t0 = time.perf_counter()
    for i in range(nsteps):
        for t, ot in zip(m["terms"], ots):
            if t["static"] is None:
                f = w.field(t["name"], i)
                if t["thresh"] is None:
                    t["tensor"].field(f)
                    ot.field(f)
                else:
                    t["tensor"].field(f, thresh=t["thresh"])
                    ot.field(f, thresh=t["thresh"])
        v, _ = tdse.update(hamiltonian(tensors), v, H0=m["h0"], inplace=True)
        H = ots[0]
        for ot in ots[1:]:
            H = H.add(ot)
        orders = []
        ref = port.update_step(H, ref, fac, phase=phase, orders=orders)
        got = v[pick].cpu().numpy()
        worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
        orders_equal = orders_equal and [int(o) for o in tdse.last_orders[pick]] == [int(o) for o in orders]
    return {"rows": pick, "steps": nsteps, "parity_max_rel": worst, "orders_equal": bool(orders_equal),
            "tolerance": 1e-10, "ok": bool(worst < 1e-10 and orders_equal),
            "oracle": "oracle/port.py (pinned to the unmodified reference)", "seconds": round(time.perf_counter() - t0, 1)}


def gpu_main(args, cpu_also=None):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from richmol_b200.ensemble import bind_to_gpu_numa
    cpus = bind_to_gpu_numa(local)                      # before any pinned allocation (first touch on the GPU's node)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = dict(world=world, rank=rank, local=local, dev=torch.device("cuda", local), cpu_also=cpu_also or {})
    head_w = WORKLOADS[args.workload]()
    if args.asym_states:
        Asym.nstates = args.asym_states
    head = gpu_workload(head_w, args, True, ctx)
    cpu_also = ctx.get("cpu_also", {})
    others = []
    for name in args.also:
        if name == args.workload:
            continue
        try:
            r = gpu_workload(WORKLOADS[name](), args, False, ctx)
        except Exception as e:        # a secondary workload must not take the headline line down with it
            r = {"name": name, "error": f"{type(e).__name__}: {e}"[:300]} if rank == 0 else None
            if world > 1:
                raise
        if r is not None:
            if name in cpu_also:
                r["cpu_baseline"] = cpu_also[name]
            others.append(r)
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return None
    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": head["steps"],
        "warmup": head["warmup"], "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": head["scaling"], "vs_baseline": None, "dtype": "c128 (f64)", "data": "synthetic",
        "config": head["config"], "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
        "roofline": head["roofline"], "parity": head["parity"],
        "workloads": [dict(r, n_gpus=world) for r in others],
    }
    line["config"]["cpu_affinity"] = (f"rank 0 bound to {len(cpus)} cores of its GPU's NUMA node (NVML ideal affinity)"
                                      if cpus else "not bound")
    line["config"]["baseline_config"] = head["baseline_config"]
    if "multi_step_call" in head:
        line["multi_step_call"] = head["multi_step_call"]
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument("--also", default=",".join(ALSO_DEFAULT),
                    help="comma-separated secondary workloads reported in `workloads` ('none' for the headline only)")
    ap.add_argument("--asym-states", type=int, default=0, help="batch of the `asym` workload (default 2048)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--cpu-kind", default=None, choices=["reference", "port"])
    args = ap.parse_args()
    args.also = [] if args.also in ("none", "") else [x for x in args.also.split(",") if x]
    for x in args.also:
        if x not in WORKLOADS:
            ap.error(f"unknown workload '{x}'")
    rank = int(os.environ.get("RANK", "0"))
    w = WORKLOADS[args.workload]()

    if args.impl == "reference":
        if rank != 0:
            return
        m = build_model(w)
        N = m["h0"]._basis().N
        r = cpu_run(w, m, args.steps, args.warmup, one_core=False, kind=args.cpu_kind)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": {"replicas": "weak"}.get(w.scaling, w.scaling), "vs_baseline": None,
            "dtype": "c128 (f64)", "data": "synthetic",
            "config": {"workload": w.text.format(N=N, S=w.nstates), "hilbert_dim": N, "dt_ps": DT,
                       "baseline_config": w.config,
                       "parallelism": f"host processes x{r['cores']}, rows of the ensemble split across them",
                       "sample": r["sample"]},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    cpu, cpu_also = None, {}
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process (workers are spawned, not forked)
        def baseline(wl, steps):
            m = build_model(wl)
            r = cpu_run(wl, m, steps, 1, kind=args.cpu_kind)
            return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                    "one_core": {"value": r["one_core"], "unit": UNIT, "cores": 1,
                                 "sample": "one process alone, 1 state x 1 step of the same workload"}}
        cpu = baseline(w, 2)
        # the light secondary workloads get their own CPU figure (seconds each); the heavy ones (asym) do not
        for name in args.also:
            if name != args.workload and name in ("ocs_align", "h2o", "ocs_mixed", "ocs_batch"):
                cpu_also[name] = baseline(WORKLOADS[name](), 3)
    line = gpu_main(args, cpu_also)
    if line is not None:
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))


if __name__ == "__main__":
    main()
