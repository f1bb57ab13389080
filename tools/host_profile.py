"""Where the host spends a bench step (cProfile over the step loop):
    python tools/host_profile.py [workload] [steps] [numpy|device]
Time inside the ctypes calls that wait for the GPU shows up under the `rmb_*` entry points; everything else is
Python / launch overhead the GPU may or may not hide."""
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

wl = sys.argv[1] if len(sys.argv) > 1 else "ocs_align"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
mode = sys.argv[3] if len(sys.argv) > 3 else "numpy"
w = bench.WORKLOADS[wl]()
m = bench.build_model(w)
tdse = TDSE(t_end=1e6, dt=bench.DT)
tdse._time_grid = (None, bench._Endless(bench.DT), None)
rows = w.rows(m, 0, w.nstates)
tensors = [t["tensor"] for t in m["terms"]]
cos2 = m.get("cos2")
if mode == "numpy":
    a = torch.empty(rows.shape, dtype=torch.complex128).pin_memory().numpy()
    b = torch.empty(rows.shape, dtype=torch.complex128).pin_memory().numpy()
    a[...] = rows
else:
    v = torch.from_numpy(rows).cuda()


def step(i):
    global a, b, v
    for t in m["terms"]:
        if t["static"] is None:
            kw = {} if t["thresh"] is None else dict(thresh=t["thresh"])
            t["tensor"].field(w.field(t["name"], i), **kw)
    H = bench.hamiltonian(tensors)
    if mode == "numpy":
        tdse.update(H, a, H0=m["h0"], out=b, expect=[cos2] if cos2 is not None else [])
        a, b = b, a
    else:
        v, _ = tdse.update(H, v, H0=m["h0"], inplace=True)
        if cos2 is not None:
            expectation(cos2, v)


for i in range(5):
    step(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for i in range(5, 5 + steps):
    step(i)
pr.disable()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f"{wl} ({mode}): {wall / steps * 1e6:.1f} us per step (host wall clock)")
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
