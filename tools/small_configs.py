"""Latency of the small BASELINE configurations (1: OCS alignment T=0; 3: OCS mixed field, 1 K ensemble):
steps per second through TDSE.update with a device-resident ensemble, vs the CPU oracle port."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from richmol_b200 import TDSE, synth, convert_units as cu
from oracle import port


def run(name, Jmax, temp, mixed, nsteps=300, cpu_steps=10):
    m = synth.ocs(Jmax)
    h0 = m["h0"]
    Hac = m["pol"] * (-0.5 * cu.AUpol_x_Vm_to_invcm())
    tdse = TDSE(t_end=1000, dt=0.01)
    tdse.time_grid()
    if mixed:
        beta = 35.0 * np.pi / 180.0
        dc = 20.7e5 * np.array([np.sin(beta), 0.0, np.cos(beta)])
        Hdc = m["dip"] * (-cu.AUdip_x_Vm_to_invcm())
        Hdc.field(dc)
    vecs0 = tdse.init_state(h0, temp=temp)
    omega = 2 * np.pi * 299792458.0 / 800e-9 * 1e-12
    field = lambda i: [0, 0, 1e10 * np.exp(-4 * np.log(2) * ((i + 0.5) * 0.01 - 1.25) ** 2) * np.cos(omega * (i + 0.5) * 0.01)]
    v = torch.from_numpy(vecs0).cuda()
    def step(i, v):
        Hac.field(field(i), thresh=1e3)
        H = (Hdc + Hac) if mixed else Hac
        v, _ = tdse.update(H, v, H0=h0, inplace=True)
        return v
    for i in range(100, 110):
        v = step(i, v)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(110, 110 + nsteps):
        v = step(i, v)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    orders = tdse.last_orders
    # CPU oracle
    oh, oac = port.OracleTensor(h0), port.OracleTensor(Hac)
    if mixed:
        odc = port.OracleTensor(Hdc); odc.field(dc)
    fac = port.exp_factor(0.01); ph = port.h0_phase(oh, fac)
    vc = vecs0.copy()
    t0 = time.perf_counter()
    for i in range(110, 110 + cpu_steps):
        oac.field(field(i), thresh=1e3)
        H = odc.add(oac) if mixed else oac
        vc = port.update_step(H, vc, fac, phase=ph)
    dtc = time.perf_counter() - t0
    # the same steps through the multi-step entry point (one call)
    flds = np.array([field(i) for i in range(110, 110 + nsteps)])
    tm = TDSE(t_end=1000, dt=0.01)
    tm.time_grid()
    terms = ([(Hdc, None, None)] if mixed else []) + [(Hac, flds, 1e3)]
    vm = torch.from_numpy(vecs0).cuda()
    tm.propagate(terms, vm, H0=h0, expect=[m["cos2"]])          # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vm2, times, ev = tm.propagate(terms, vm, H0=h0, expect=[m["cos2"]])
    torch.cuda.synchronize()
    dtm = time.perf_counter() - t0
    print(f"{name}: multi-step call with <cos2> every step: {nsteps/dtm:.0f} steps/s ({dtm/nsteps*1e6:.0f} us/step), "
          f"ratio vs CPU port {nsteps/dtm/(cpu_steps/dtc):.0f}x")
    print(f"{name}: N={vecs0.shape[1]} states={len(vecs0)} orders {orders.min()}..{orders.max()}  GPU {nsteps/dt:.0f} steps/s "
          f"({len(vecs0)*nsteps/dt:.0f} state-steps/s, {dt/nsteps*1e6:.0f} us/step)  CPU port 1 core {cpu_steps/dtc:.1f} steps/s  ratio {nsteps/dt/(cpu_steps/dtc):.0f}x")


run("config1 OCS alignment Jmax=30 T=0", 30, 0, False)
run("config3 OCS mixed field Jmax=60 T=1K", 60, 1.0, True, nsteps=200, cpu_steps=3)
