"""Per-phase cycle counts of k_lanczos_fused (library built with -DRMB_FUSED_TRACE):
    EXTRA=-DRMB_FUSED_TRACE tools/ab_build.sh WORKTREE ftrace
    RMB_LIB=build_variants/lib_ftrace.so python tools/fused_trace.py ocs_align"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from richmol_b200 import TDSE

wl = sys.argv[1] if len(sys.argv) > 1 else "ocs_align"
w = bench.WORKLOADS[wl]()
m = bench.build_model(w)
tdse = TDSE(t_end=1e6, dt=bench.DT)
tdse._time_grid = (None, bench._Endless(bench.DT), None)
rows = w.rows(m, 0, w.nstates)
tensors = [t["tensor"] for t in m["terms"]]
v = torch.from_numpy(rows).cuda()
for i in list(range(0, 6)) + [500, 501, 1000, 1001]:
    for t in m["terms"]:
        if t["static"] is None:
            kw = {} if t["thresh"] is None else dict(thresh=t["thresh"])
            t["tensor"].field(w.field(t["name"], i), **kw)
    v, _ = tdse.update(bench.hamiltonian(tensors), v, H0=m["h0"], inplace=True)
    torch.cuda.synchronize()
    print("step", i, "orders", tdse.last_orders[:4], flush=True)
