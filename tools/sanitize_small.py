"""Small drivers for compute-sanitizer (racecheck / synccheck / memcheck): one launch of each kernel family on
problems small enough for the tools' slow-down.  usage: python tools/sanitize_small.py gemm|fused|tiled|recur"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import port
from richmol_b200 import TDSE, convert_units as cu, synth

which = sys.argv[1]
fac = port.exp_factor(0.01)
tdse = TDSE(t_end=1, dt=0.01)
tdse.time_grid()
rng = np.random.default_rng(1)
if which == "gemm":            # dim_k 13..14: k_matvec_dmma (DMMA) + the batched Lanczos kernels
    st = synth.asymmetric_rotor(*synth.H2S_ABC, 27, Jmin=26)
    h0 = synth.hamiltonian_tensor(st)
    H = synth.lab_tensor(synth.H2S_POL, st) * (-0.5 * cu.AUpol_x_Vm_to_invcm())
    E, nst = [3e9, -2e9, 4e9], 4
elif which == "tiled":         # dim_k <= 6: k_matvec_tiled
    m = synth.h2o(5)
    h0, H = m["h0"], m["pol"] * (-0.5 * cu.AUpol_x_Vm_to_invcm())
    E, nst = [3e9, 1e9, 4e9], 6
else:                          # linear rotor: k_lanczos_fused (whole step in one launch)
    m = synth.ocs(10)
    h0, H = m["h0"], m["pol"] * (-0.5 * cu.AUpol_x_Vm_to_invcm())
    E, nst = [0, 0, 8e9], 3
    if which == "recur":
        os.environ["RMB_FUSED"] = "0"
N = h0._basis().N
v = rng.normal(size=(nst, N)) + 1j * rng.normal(size=(nst, N))
v /= np.linalg.norm(v, axis=1)[:, None]
H.field(E)
out, _ = tdse.update(H, torch.from_numpy(v).cuda(), H0=h0)
o = port.OracleTensor(H)
o.field(E)
orders = []
ref = port.update_step(o, v, fac, phase=port.h0_phase(port.OracleTensor(h0), fac), orders=orders)
err = np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()
print(f"{which}: N={N}, {nst} states, rel err {err:.2e}, orders {list(tdse.last_orders)} vs {orders}")
sys.exit(0 if err < 1e-10 and list(tdse.last_orders) == orders else 1)
