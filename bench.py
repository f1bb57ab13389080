#!/usr/bin/env python
"""Benchmark of the TDSE hot path (BASELINE.json metric: ensemble state-timesteps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload h2o|ocs]

Workload (BASELINE.json configs[1]): H2O rigid rotor (Watson A-reduction, D2-type symmetry blocks),
Jmax = 20 (N = 12 341), 500-state Boltzmann shard per GPU, H(t) = H0 - mu.E_dc(t) - 1/2 alpha:E_ac E_ac
with a tilted (M-mixing) dc field ramp and a z-polarised 800 nm Gaussian pulse; one *step* =
`Hdc.field(E_dc(t)); Hac.field(E_ac(t)); tdse.update(Hdc + Hac, H0=h0, vecs=vecs)` followed by the
ensemble expectation value <cos^2 theta> (+ one NCCL all-reduce of the observables at N > 1).
All inputs are synthetic (richmol_b200/synth.py); arithmetic is complex128 throughout.

`--impl reference` times the reference's CPU algorithm for the same step (the numpy/scipy port in
oracle/port.py -- the reference itself is pure Python and cannot travel to the GPU box) on all host
cores, each step a bounded sample of the ensemble.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "state-timesteps/sec"
UNIT = "state-steps/s"
DT = 0.01            # ps
NSTATES = 500        # states per GPU (weak scaling: the ensemble grows with the number of GPUs)
NSTATES_BY_WORKLOAD = {"h2o": 500, "ocs": 8192}   # the linear-rotor case needs a batch larger than L2
TEMP = 300.0         # K, Boltzmann weights of the shard rows


# ------------------------------------------------------------------------------------------------
# workload definition (shared by the GPU arm and the CPU arm)
# ------------------------------------------------------------------------------------------------
WORKLOAD_TEXT = {
    "h2o": "h2o: H2O rigid rotor Watson-A Jmax=20 N={N}, {S}-state Boltzmann shard per GPU, dc dipole (tilted, "
           "ramp) + ac polarisability, split-operator Lanczos step + <cos2theta>",
    "ocs": "ocs: OCS linear rotor Jmax=60 N={N}, {S}-state shard per GPU, tilted dc dipole + ac polarisability "
           "(M-mixing), split-operator Lanczos step + <cos2theta>",
}


def fields_at(step):
    """dc: 50 kV/cm tilted 35 deg in the XZ plane, ramped; ac: 800 nm Gaussian pulse along Z."""
    t = (step + 0.5) * DT
    beta = 35.0 * np.pi / 180.0
    ramp = 0.5 + 0.5 * min(1.0, step / 200.0)
    dc = 5e6 * ramp * np.array([np.sin(beta), 0.0, np.cos(beta)])
    omega = 2 * np.pi * 299792458.0 / 800e-9 * 1e-12
    t0, fwhm = 1.0, 1.0
    ac = np.array([0.0, 0.0, 3e9 * np.exp(-4 * np.log(2) * (t - t0) ** 2 / fwhm ** 2) * np.cos(omega * t)])
    return dc, ac


def build_model(workload):
    from richmol_b200 import convert_units as cu, synth
    if workload == "h2o":
        m = synth.h2o(20)
    elif workload == "ocs":
        m = synth.ocs(60)
    else:
        raise ValueError(workload)
    m["Hdc"] = m["dip"] * (-cu.AUdip_x_Vm_to_invcm())
    m["Hac"] = m["pol"] * (-0.5 * cu.AUpol_x_Vm_to_invcm())
    return m


def ensemble_rows(h0, first, count):
    """Rows [first, first+count) of the Boltzmann ensemble sqrt(w_i)|i> in basis order
    (what TDSE.init_state builds, richmol/tdse.py:250-257), without the prefix truncation."""
    import scipy.constants as const
    enr = h0.tomat(form="full", cart="0").diagonal().real
    enr = enr - enr[0]
    from richmol_b200 import convert_units as cu
    w = np.exp(-enr / cu.J_to_invcm() / (const.value("Boltzmann constant") * TEMP))
    w /= w.sum()
    N = len(enr)
    idx = (first + np.arange(count)) % N
    v = np.zeros((count, N), dtype=np.complex128)
    v[np.arange(count), idx] = np.sqrt(w[idx])
    return v


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on host cores, rows split across processes
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(workload):
    from oracle import port
    m = build_model(workload)
    _W["port"] = port
    _W["h0"] = port.OracleTensor(m["h0"])
    _W["dc"] = port.OracleTensor(m["Hdc"])
    _W["ac"] = port.OracleTensor(m["Hac"])
    c2 = port.OracleTensor(m["cos2"])
    c2.field([0, 0, 1])
    _W["cos2"] = c2.tomat()
    fac = port.exp_factor(DT)
    _W["fac"] = fac
    _W["phase"] = port.h0_phase(_W["h0"], fac)


def _cpu_steps(args):
    rows, step0, nsteps = args
    port = _W["port"]
    v = rows
    ev = 0.0
    for s in range(step0, step0 + nsteps):
        dc, ac = fields_at(s)
        _W["dc"].field(dc)
        _W["ac"].field(ac, thresh=1e1)
        H = _W["dc"].add(_W["ac"])
        v = port.update_step(H, v, _W["fac"], phase=_W["phase"])
        ev = sum(np.dot(np.conj(x), _W["cos2"].dot(x)) for x in v) + 1 / 3
    return v, ev


def cpu_run(workload, h0, steps, warmup, states_per_core, step0=0):
    """Times `steps` steps of the reference algorithm on a sample of the ensemble using every host
    core (the reference's own scale-out pattern: rows of `vecs` split across processes,
    docs/source/notebooks/tdse_mpi.ipynb:268-272).  Returns (state-steps/s, cores, sample text)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    nst = cores * states_per_core
    rows = ensemble_rows(h0, 0, nst)
    chunks = [rows[i * states_per_core:(i + 1) * states_per_core] for i in range(cores)]
    ctx = mp.get_context("spawn")
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:                      # one BLAS thread per worker process (inherited on spawn)
        os.environ[k] = "1"
    pool = ctx.Pool(cores, initializer=_cpu_init, initargs=(workload,))
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    with pool:
        if warmup > 0:
            res = pool.map(_cpu_steps, [(c, step0, warmup) for c in chunks])
            chunks = [r[0] for r in res]
        t0 = time.perf_counter()
        pool.map(_cpu_steps, [(c, step0 + warmup, steps) for c in chunks])
        dt = time.perf_counter() - t0
    sample = (f"{nst} of {NSTATES} states x {steps} steps ({warmup} warm-up), {cores} processes x "
              f"{states_per_core} states, numpy/scipy port of richmol CarTens.field/vec + "
              f"TDSE.update (oracle/port.py)")
    return nst * steps / dt, cores, sample, dt / steps * 1e3


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
def gpu_run(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from richmol_b200 import TDSE, _lib
    from richmol_b200.tdse import expectation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    m = build_model(args.workload)
    h0, Hdc, Hac, cos2 = m["h0"], m["Hdc"], m["Hac"], m["cos2"]
    cos2.field([0, 0, 1])
    N = h0._basis().N
    rows = ensemble_rows(h0, rank * NSTATES, NSTATES)          # this rank's shard (weak scaling)
    tdse = TDSE(t_end=1e6, dt=DT)
    tdse.time_grid = lambda *a, **k: None                      # open-ended grid for the benchmark
    tdse._time_grid = (None, _Endless(DT), None)
    vecs = torch.from_numpy(rows).to(dev)
    obs = torch.zeros(1, dtype=torch.complex128, device=dev)

    def step(i, v):
        dc, ac = fields_at(i)
        Hdc.field(dc)
        Hac.field(ac, thresh=1e1)
        v, _ = tdse.update(Hdc + Hac, v, H0=h0, inplace=True)
        ev = expectation(cos2, v)
        torch.sum(ev, dim=0, keepdim=True, out=obs)            # ensemble <cos^2 theta> - 1/3 of this shard
        if world > 1:
            dist.all_reduce(torch.view_as_real(obs))           # the path's only collective
        return v

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        vecs = step(i, vecs)
    op = (Hdc + Hac)._device()
    lib = _lib.lib()
    cnt0 = op.counters()
    c2op = cos2._device()
    cnt0c = c2op.counters()
    ms_ = C.c_double()
    n_ = C.c_int64()
    lib.rmb_matvec_timing(op.handle, 1, C.byref(ms_), C.byref(n_))     # enable + reset
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, args.warmup + args.steps):
        vecs = step(i, vecs)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    lib.rmb_matvec_timing(op.handle, 0, C.byref(ms_), C.byref(n_))
    cnt1, cnt1c = op.counters(), c2op.counters()
    launches = (cnt1["launches"] - cnt0["launches"]) + (cnt1c["launches"] - cnt0c["launches"])
    state_mv = cnt1["state_matvecs"] - cnt0["state_matvecs"]
    mv_launches = max(1, n_.value)
    value = world * NSTATES * args.steps / (ms * 1e-3)

    # roofline of the H.Psi matvec: 32*N bytes per state-matvec + operator bytes once
    # per launch; flops from the block tables
    info = op_info(op)
    alg_bytes = 32.0 * N * state_mv + info["op_bytes"] * mv_launches
    alg_flops = info["flops_per_state"] * state_mv
    mv_s = ms_.value * 1e-3
    peaks = measured_peaks()
    gbs = alg_bytes / mv_s / 1e9 if mv_s > 0 else 0.0
    tfs = alg_flops / mv_s / 1e12 if mv_s > 0 else 0.0
    ai = info["flops_per_state"] / (32.0 * N)
    fp64_peak = 37.2     # TFLOP/s, DMMA microbenchmark on this pool's B200 (DFMA: 35.3), profiles/r01_fp64_peak.txt
    ridge = fp64_peak * 1e3 / peaks["hbm_gbs"]
    # The matvec of an asymmetric top (dense K blocks) sits right of the ridge: it is bounded by the FP64 pipe
    # ("tensor": on B200 the FP64 tensor pipe and the FMA pipe have the same peak); linear rotors (dim_k = 1)
    # are HBM-bound.  Both fractions are reported, the headline one follows the arithmetic intensity.
    hbm = {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
           "peak_source": peaks["source"]}
    fp64 = {"achieved": tfs, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfs / fp64_peak,
            "peak_source": "FP64 DMMA microbenchmark on this pool's B200, profiles/r01_fp64_peak.txt (of measured)"}
    # (a workload within 25% of the ridge, like the OCS linear rotor, is reported against HBM: its operator
    # reads come from L2 and the state vectors are the only compulsory DRAM traffic)
    compute_bound = ai > 1.25 * ridge
    head = fp64 if compute_bound else hbm
    kernel = ("k_matvec_tiled (H.Psi: warpgroup-specialised, TMA-staged ket rows, fused MF(x)K block products, fused <w,V_k>)"
              if args.workload == "h2o" else
              "k_matvec_lin (H.Psi of a linear rotor: sliding window of ket blocks in shared memory, fused <w,V_k>)")
    roofline = {
        "kernel": kernel,
        "bound": "tensor" if compute_bound else "hbm", "achieved": head["achieved"], "peak": head["peak"],
        "unit": head["unit"], "frac": head["frac"], "traffic": ncu_traffic(args.workload),
        "peak_source": head["peak_source"], "hbm": hbm, "fp64": fp64,
        "arithmetic_intensity": ai, "ridge": ridge, "flops_per_state_matvec": info["flops_per_state"],
        "bytes_per_state_matvec": 32.0 * N, "operator_bytes_per_launch": info["op_bytes"],
        "launches": int(mv_launches), "avg_launch_us": mv_s / mv_launches * 1e6,
        "share_of_step": mv_s * 1e3 / ms, "matvecs_per_state_step": state_mv / (NSTATES * args.steps),
    }

    # ---- end to end through the public API with HOST buffers (numpy in, numpy out)
    e2e = e2e_run(args, tdse, Hdc, Hac, h0, cos2, rows, world, dev)

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return None
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64)", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload].format(N=N, S=NSTATES),
                   "states_per_gpu": NSTATES, "hilbert_dim": N, "dt_ps": DT,
                   "parallelism": f"ensemble-sharded x{world}",
                   "l2": "working set (Psi + Krylov vectors, ~0.6 GB per GPU) larger than the 126 MB L2"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
    }


class _Endless:
    def __init__(self, dt):
        self.dt = dt

    def __getitem__(self, i):
        return (i + 1) * self.dt


def op_info(op):
    """flops / operator bytes per state-matvec for the field currently applied (formulae of SURVEY.md 8d;
    only the M diagonals that survive the field contraction are counted, as in the reference's CSR)."""
    import ctypes as C
    from richmol_b200 import _lib
    fl, by = C.c_double(), C.c_double()
    _lib.check(_lib.lib().rmb_operator_work(op.handle, C.byref(fl), C.byref(by), None))
    return {"flops_per_state": fl.value, "op_bytes": by.value}


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one full matvec launch from the committed ncu capture
    (profiles/), bytes per launch; None if no capture exists for this workload."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_matvec_traffic.json")))[workload]
        return int(t["dram_bytes_read"] + t["dram_bytes_write"])
    except Exception:
        return None


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm_gbs": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json (of measured)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


def e2e_run(args, tdse, Hdc, Hac, h0, cos2, rows, world, dev):
    """Same step through TDSE.update with numpy arrays in pinned host memory: every step copies the
    ensemble host->device and the propagated ensemble + observable device->host."""
    import torch
    import torch.distributed as dist
    from richmol_b200.tdse import expectation
    pin_a = torch.empty(rows.shape, dtype=torch.complex128).pin_memory()
    pin_b = torch.empty(rows.shape, dtype=torch.complex128).pin_memory()
    a, b = pin_a.numpy(), pin_b.numpy()
    a[...] = rows
    nsteps = max(3, min(args.steps, 10))

    def step(i, src, dst):
        dc, ac = fields_at(i)
        Hdc.field(dc)
        Hac.field(ac, thresh=1e1)
        tdse.update(Hdc + Hac, src, H0=h0, out=dst, expect=[cos2])   # host array in -> host arrays out
        return complex(tdse.last_expect[0].sum())

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
    nbytes = rows.size * 16
    return {"value": world * NSTATES * nsteps / float(dt.item()), "unit": UNIT, "steps": nsteps,
            "h2d_bytes_per_step": int(nbytes + rows.shape[1] * 16),
            "d2h_bytes_per_step": int(nbytes + rows.shape[0] * 16),
            "note": "numpy (pinned) in/out through TDSE.update(..., expect=[cos2]); includes the host-side "
                    "field products, the H2D/D2H copies of the ensemble (chunked, overlapped with the kernels) "
                    "and the per-step observable"}


def main():
    global NSTATES
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="h2o", choices=["h2o", "ocs"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-states-per-core", type=int, default=4)
    args = ap.parse_args()
    NSTATES = NSTATES_BY_WORKLOAD[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        m = build_model(args.workload)
        v, cores, sample, ms_step = cpu_run(args.workload, m["h0"], args.steps, min(args.warmup, 1),
                                            args.cpu_states_per_core)
        N = m["h0"]._basis().N
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload].format(N=N, S=NSTATES),
                       "states_per_gpu": NSTATES, "hilbert_dim": N, "dt_ps": DT,
                       "parallelism": f"host processes x{cores}, rows of the ensemble split across them",
                       "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process (workers are spawned, not forked)
        m = build_model(args.workload)
        v, cores, sample, _ = cpu_run(args.workload, m["h0"], 6, 1, args.cpu_states_per_core)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    line = gpu_run(args)
    if line is not None:
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))


if __name__ == "__main__":
    main()
