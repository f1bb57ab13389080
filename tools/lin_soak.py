"""Soak of the sliding-window matvec k_matvec_lin: N launches of the same product per
instantiation (8- and 4-state tiles, two state groups and the single-state-group variant RMB_LIN_G1=1 that failed
intermittently in round 1), every result compared bit for bit with the first one and the first one with the oracle.

    python tools/lin_soak.py [launches=1000] [states=512] [modes=T8,T4,T8G1,T4G1]
Also the driver of the racecheck / synccheck runs (a handful of launches under compute-sanitizer)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import port
from richmol_b200 import _lib, convert_units as cu, synth
from richmol_b200.field import clear_device_cache

nlaunch = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nstates = int(sys.argv[2]) if len(sys.argv) > 2 else 512
modes = (sys.argv[3] if len(sys.argv) > 3 else "T8,T4,T8G1,T4G1").split(",")

m = synth.ocs(60)
dip = m["dip"] * (-cu.AUdip_x_Vm_to_invcm())
pol = m["pol"] * (-0.5 * cu.AUpol_x_Vm_to_invcm())
Edc, Eac = [2.9e6, 0.0, 4.1e6], [0.0, 0.0, 2.5e9]
N = dip._basis().N
rng = np.random.default_rng(7)
x = rng.normal(size=(nstates, N)) + 1j * rng.normal(size=(nstates, N))
od, op_ = port.OracleTensor(dip), port.OracleTensor(pol)
od.field(Edc)
op_.field(Eac, thresh=1e1)
oH = od.add(op_)
pick = [0, nstates // 2, nstates - 1]
yo = np.array([port.flat_matvec(oH, x[i]) for i in pick])
xd = torch.from_numpy(x).cuda()
bad_total = 0
for mode in modes:
    os.environ["RMB_LIN_T"] = "4" if mode.startswith("T4") else "8"
    os.environ["RMB_LIN_G1"] = "1" if mode.endswith("G1") else "0"
    clear_device_cache()
    dip.field(Edc)
    pol.field(Eac, thresh=1e1)
    H = dip + pol
    op = H._device()
    lib = _lib.lib()
    y0 = torch.zeros_like(xd)
    _lib.check(lib.rmb_matvec(op.handle, xd.data_ptr(), y0.data_ptr(), nstates, N, None))
    torch.cuda.synchronize()
    err = np.abs(y0[pick].cpu().numpy() - yo).max() / np.abs(yo).max()
    y = torch.zeros_like(xd)
    nbad = torch.zeros((), dtype=torch.int64, device="cuda")
    worst = torch.zeros((), dtype=torch.float64, device="cuda")
    t0 = time.time()
    for i in range(nlaunch):
        _lib.check(lib.rmb_matvec(op.handle, xd.data_ptr(), y.data_ptr(), nstates, N, None))
        d = (torch.view_as_real(y) != torch.view_as_real(y0)).any()
        nbad += d
        worst = torch.maximum(worst, (y - y0).abs().max())
    torch.cuda.synchronize()
    print(f"{mode}: first launch vs oracle rel err {err:.2e}; {nlaunch} launches, {int(nbad)} differ from the first "
          f"(max abs diff {float(worst):.3e}), {time.time() - t0:.1f}s", flush=True)
    bad_total += int(nbad) + (0 if err < 1e-12 else 1)
sys.exit(1 if bad_total else 0)
