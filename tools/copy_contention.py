"""Does concurrent PCIe traffic slow the propagation kernels?  Device-resident h2s steps (batch n) alone, and
with continuous pinned H2D / D2H copies on two other streams (what the chunks of the host-buffer pipeline see)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from richmol_b200 import TDSE
from richmol_b200.tdse import expectation

wl = sys.argv[1] if len(sys.argv) > 1 else "h2s"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
w = bench.WORKLOADS[wl]()
m = bench.build_model(w)
tdse = TDSE(t_end=1e6, dt=bench.DT)
tdse._time_grid = (None, bench._Endless(bench.DT), None)
v = torch.from_numpy(w.rows(m, 0, n)).cuda()
tensors = [t["tensor"] for t in m["terms"]]
nb = 136 * 1024 * 1024
h_in = torch.empty(nb, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nb, dtype=torch.uint8).pin_memory()
d_a = torch.empty(nb, dtype=torch.uint8, device="cuda")
d_b = torch.empty(nb, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def step(i, v):
    for t in m["terms"]:
        if t["static"] is None:
            t["tensor"].field(w.field(t["name"], i))
    v, _ = tdse.update(bench.hamiltonian(tensors), v, H0=m["h0"], inplace=True)
    expectation(m["cos2"], v)
    return v


def timed(mode, K=10):
    global v
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 3 + K):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
        if mode.startswith("d2h_"):
            piece = int(mode.split("_")[1]) * 1024 * 1024
            with torch.cuda.stream(s2):
                for o in range(0, nb, piece):
                    h_out[o:o + piece].copy_(d_b[o:o + piece], non_blocking=True)
        v = step(i, v)
    e1.record()
    torch.cuda.synchronize()
    print(f"{wl} batch {n}, concurrent copies: {mode:5s}: {e0.elapsed_time(e1) / K:.3f} ms/step", flush=True)


for i in range(3):
    v = step(i, v)
for mode in ("none", "d2h", "d2h_16", "d2h_4", "d2h_1", "none"):
    timed(mode)
