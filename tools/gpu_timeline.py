"""GPU timeline of device-resident bench steps (CUPTI through torch.profiler; no profiler-timed number is a bench value):
    python tools/gpu_timeline.py [workload] [steps] [nstates]
Prints busy time per kernel, the idle time of the device between the first and the last kernel of the timed steps, and
the largest gaps with the kernels on either side."""
import collections
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from richmol_b200 import TDSE
from richmol_b200.tdse import expectation

wl = sys.argv[1] if len(sys.argv) > 1 else "h2s"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = bench.WORKLOADS[wl]()
m = bench.build_model(w)
tdse = TDSE(t_end=1e6, dt=bench.DT)
tdse._time_grid = (None, bench._Endless(bench.DT), None)
rows = w.rows(m, 0, int(sys.argv[3]) if len(sys.argv) > 3 else w.nstates)
tensors = [t["tensor"] for t in m["terms"]]
cos2 = m.get("cos2")
v = torch.from_numpy(rows).cuda()


def step(i):
    global v
    for t in m["terms"]:
        if t["static"] is None:
            kw = {} if t["thresh"] is None else dict(thresh=t["thresh"])
            t["tensor"].field(w.field(t["name"], i), **kw)
    v, _ = tdse.update(bench.hamiltonian(tensors), v, H0=m["h0"], inplace=True)
    if cos2 is not None:
        expectation(cos2, v)


for i in range(4):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(4, 4 + steps):
    step(i)
e1.record()
torch.cuda.synchronize()
print(f"{wl}: unprofiled {e0.elapsed_time(e1) / steps:.3f} ms/step")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(4 + steps, 4 + 2 * steps):
        step(i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy = collections.Counter()
cnt = collections.Counter()
gaps = []
end = t0
for e in ev:
    name = e.name.split("(")[0][:48]
    busy[name] += e.time_range.end - e.time_range.start
    cnt[name] += 1
    if e.time_range.start > end:
        gaps.append((e.time_range.start - end, prev, name))
    if e.time_range.end > end:
        end, prev = e.time_range.end, name
span = t1 - t0
tot = sum(busy.values())
print(f"span {span / steps / 1e3:.3f} ms/step, kernels {tot / steps / 1e3:.3f} ms/step (overlap counted twice), "
      f"idle {sum(g[0] for g in gaps) / steps / 1e3:.3f} ms/step in {len(gaps) / steps:.1f} gaps/step")
for k, t in busy.most_common(12):
    print(f"  {k:50s} {cnt[k] / steps:6.1f}/step {t / steps / 1e3:9.3f} ms/step  avg {t / cnt[k]:9.1f} us")
agg = collections.Counter()
for g, a, b in gaps:
    agg[(a, b)] += g
print("idle by (kernel before -> kernel after), ms/step:")
for (a, b), g in agg.most_common(10):
    print(f"  {g / steps / 1e3:8.3f}  {a} -> {b}")
