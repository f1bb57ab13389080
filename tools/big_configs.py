"""BASELINE configs 4 and 5 at (near) full size: H2S optical centrifuge Jmax=80 (N = 708 561) and a J <= 100
asymmetric top (N ~ 1.37e6).  Reports state-steps/s and the matvec FP64 rate; one oracle matvec as anchor."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import port
from richmol_b200 import TDSE, _lib, convert_units as cu, synth


def run(name, Jmax, nstates, nsteps, check=True):
    t0 = time.time()
    st = synth.asymmetric_rotor(*synth.H2S_ABC, Jmax)
    h0 = synth.hamiltonian_tensor(st)
    pol = synth.lab_tensor(synth.H2S_POL, st) * (-0.5 * cu.AUpol_x_Vm_to_invcm())
    N = pol._basis().N
    print(f"{name}: built N={N} in {time.time()-t0:.1f}s, max dim_k {max(pol._basis().dk)}", flush=True)
    tdse = TDSE(t_end=1e6, dt=0.01)
    tdse.time_grid()
    rng = np.random.default_rng(0)
    idx = rng.integers(0, N, size=nstates)
    v = torch.zeros((nstates, N), dtype=torch.complex128, device="cuda")
    v[torch.arange(nstates), torch.from_numpy(idx).cuda()] = 1.0

    def field(i):   # optical centrifuge: rotating linear polarisation
        b = 0.02
        return [3e9 * np.cos(b * i * i), 3e9 * np.sin(b * i * i), 0.0]
    t0 = time.time()
    pol.field(field(1))
    op = pol._device()
    print(f"  packed + uploaded in {time.time()-t0:.1f}s", flush=True)
    if check:
        o = port.OracleTensor(pol)
        o.field(field(1))
        x = rng.normal(size=N) + 1j * rng.normal(size=N)
        t0 = time.time()
        yo = port.flat_matvec(o, x)
        tc = time.time() - t0
        xd = torch.from_numpy(x[None, :]).cuda()
        yd = torch.zeros_like(xd)
        _lib.check(_lib.lib().rmb_matvec(op.handle, xd.data_ptr(), yd.data_ptr(), 1, N, None))
        err = np.abs(yd.cpu().numpy()[0] - yo).max() / np.abs(yo).max()
        print(f"  matvec vs oracle: rel err {err:.2e} (oracle {tc:.2f}s per state-matvec on one core)", flush=True)
        assert err < 1e-12
    for i in range(2):
        pol.field(field(i))
        v, _ = tdse.update(pol, v, H0=h0, inplace=True)
    lib = _lib.lib()
    ms_, n_ = C.c_double(), C.c_int64()
    lib.rmb_matvec_timing(op.handle, 1, C.byref(ms_), C.byref(n_))
    c0 = op.counters()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(2, 2 + nsteps):
        pol.field(field(i))
        v, _ = tdse.update(pol, v, H0=h0, inplace=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    lib.rmb_matvec_timing(op.handle, 0, C.byref(ms_), C.byref(n_))
    c1 = op.counters()
    fl, by = C.c_double(), C.c_double()
    lib.rmb_operator_work(op.handle, C.byref(fl), C.byref(by), None)
    smv = c1["state_matvecs"] - c0["state_matvecs"]
    print(f"  {nstates} states x {nsteps} steps: {nstates*nsteps/dt:.1f} state-steps/s ({dt/nsteps*1e3:.1f} ms/step), "
          f"orders {tdse.last_orders.min()}..{tdse.last_orders.max()}, {smv/(nstates*nsteps):.2f} matvecs/state-step; "
          f"matvec {fl.value/1e9:.3f} Gflop/state, {fl.value*smv/(ms_.value*1e-3)/1e12:.2f} TFLOP/s FP64, "
          f"share of step {ms_.value/(dt*1e3):.2f}, norm drift {abs(float((v.abs()**2).sum(1).max())-1):.1e}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "4"
    if which == "4":
        run("config4 H2S centrifuge Jmax=80", 80, 64, 5)
    elif which == "5":
        run("config5 asym top Jmax=100", 100, 128, 3)
    else:
        run("medium Jmax=40", 40, 64, 5)
