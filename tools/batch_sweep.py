"""Device-resident step time of a bench workload as a function of the batch size (states per launch):
    python tools/batch_sweep.py [workload] n1 n2 ...
Shows what a chunk of the host-buffer pipeline (or a sub-batch of a large ensemble) costs."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from richmol_b200 import TDSE
from richmol_b200.tdse import expectation

wl = sys.argv[1] if len(sys.argv) > 1 else "h2o"
sizes = [int(a) for a in sys.argv[2:]] or [500, 250, 168, 126, 84]
w = bench.WORKLOADS[wl]()
m = bench.build_model(w)
tdse = TDSE(t_end=1e6, dt=bench.DT)
tdse._time_grid = (None, bench._Endless(bench.DT), None)
rows = w.rows(m, 0, max(sizes))
tensors = [t["tensor"] for t in m["terms"]]
for n in sizes:
    v = torch.from_numpy(rows[:n].copy()).cuda()

    def step(i, v, obs=True):
        for t in m["terms"]:
            if t["static"] is None:
                kw = {} if t["thresh"] is None else dict(thresh=t["thresh"])
                t["tensor"].field(w.field(t["name"], i), **kw)
        v, _ = tdse.update(bench.hamiltonian(tensors), v, H0=m["h0"], inplace=True)
        if obs and m.get("cos2") is not None:
            expectation(m["cos2"], v)
        return v
    for i in range(3):
        v = step(i, v)
    for obs in (False, True):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 10
        for i in range(3, 3 + K):
            v = step(i, v, obs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"{wl} batch {n} {'with' if obs else 'no'} <cos2>: {ms:.3f} ms/step, {n / ms * 1e3:.0f} state-steps/s, "
              f"{ms / n * 1e3:.2f} us per state", flush=True)
