"""Where the host spends a device-resident bench step (cProfile over the step loop of bench.py):
    python tools/host_profile.py [h2o|ocs] [steps]
Time inside the ctypes calls that wait for the GPU (event / stream synchronisation) shows up under
`rmb_propagate_step`; everything else is Python / launch overhead the GPU may or may not hide."""
import cProfile
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from richmol_b200 import TDSE
from richmol_b200.tdse import expectation

wl = sys.argv[1] if len(sys.argv) > 1 else "ocs"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bench.NSTATES = bench.NSTATES_BY_WORKLOAD[wl]
m = bench.build_model(wl)
h0, Hdc, Hac, cos2 = m["h0"], m["Hdc"], m["Hac"], m["cos2"]
cos2.field([0, 0, 1])
tdse = TDSE(t_end=1e6, dt=bench.DT)
tdse.time_grid = lambda *a, **k: None
tdse._time_grid = (None, bench._Endless(bench.DT), None)
v = torch.from_numpy(bench.ensemble_rows(h0, 0, bench.NSTATES)).cuda()
obs = torch.zeros(1, dtype=torch.complex128, device="cuda")
tsplit = {"field": 0.0, "update": 0.0, "expect": 0.0}


def step(i, v):
    t0 = time.perf_counter()
    dc, ac = bench.fields_at(i)
    Hdc.field(dc)
    Hac.field(ac, thresh=1e1)
    H = Hdc + Hac
    t1 = time.perf_counter()
    v, _ = tdse.update(H, v, H0=h0, inplace=True)
    t2 = time.perf_counter()
    ev = expectation(cos2, v)
    torch.sum(ev, dim=0, keepdim=True, out=obs)
    t3 = time.perf_counter()
    tsplit["field"] += t1 - t0
    tsplit["update"] += t2 - t1
    tsplit["expect"] += t3 - t2
    return v


for i in range(3):
    v = step(i, v)
torch.cuda.synchronize()
for k in tsplit:
    tsplit[k] = 0.0
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for i in range(3, 3 + steps):
    v = step(i, v)
pr.disable()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f"{wl}: {wall / steps * 1e3:.3f} ms per step (host wall clock), host split per step (ms):",
      {k: round(x / steps * 1e3, 3) for k, x in tsplit.items()})
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
