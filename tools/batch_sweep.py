"""Device-resident step time of the bench workload as a function of the batch size (states per launch):
    python tools/batch_sweep.py [h2o|ocs] n1 n2 ...
Shows what a chunk of the host-buffer pipeline (or a sub-batch of a large ensemble) costs."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from richmol_b200 import TDSE

wl = sys.argv[1] if len(sys.argv) > 1 else "h2o"
sizes = [int(a) for a in sys.argv[2:]] or [500, 250, 168, 126, 84]
m = bench.build_model(wl)
h0, Hdc, Hac = m["h0"], m["Hdc"], m["Hac"]
tdse = TDSE(t_end=1e6, dt=bench.DT)
tdse.time_grid = lambda *a, **k: None
tdse._time_grid = (None, bench._Endless(bench.DT), None)
rows = bench.ensemble_rows(h0, 0, max(sizes))
for n in sizes:
    v = torch.from_numpy(rows[:n].copy()).cuda()
    def step(i, v):
        dc, ac = bench.fields_at(i)
        Hdc.field(dc)
        Hac.field(ac, thresh=1e1)
        v, _ = tdse.update(Hdc + Hac, v, H0=h0, inplace=True)
        return v
    for i in range(3):
        v = step(i, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 30
    for i in range(3, 3 + K):
        v = step(i, v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"{wl} batch {n}: {ms:.3f} ms/step, {n / ms * 1e3:.0f} state-steps/s, {ms / n * 1e3:.2f} us per state")
