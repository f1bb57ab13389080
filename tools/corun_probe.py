"""Would co-running chunks recover the batch efficiency?  Two (or three) independent operator handles, each
propagating its share of the h2o ensemble device-resident on its own stream from its own host thread."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from richmol_b200 import TDSE
from richmol_b200.tdse import expectation

wl = sys.argv[1] if len(sys.argv) > 1 else "h2o"
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = bench.WORKLOADS[wl]()
total = w.nstates
models = [bench.build_model(w) for _ in range(nthreads)]
K = 40


def worker(idx, out):
    m = models[idx]
    n = total // nthreads
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        tdse = TDSE(t_end=1e6, dt=bench.DT)
        tdse._time_grid = (None, bench._Endless(bench.DT), None)
        v = torch.from_numpy(w.rows(m, idx * n, (idx + 1) * n)).cuda()
        tensors = [t["tensor"] for t in m["terms"]]

        def step(i, v):
            for t in m["terms"]:
                if t["static"] is None:
                    kw = {} if t["thresh"] is None else dict(thresh=t["thresh"])
                    t["tensor"].field(w.field(t["name"], i), **kw)
            v, _ = tdse.update(bench.hamiltonian(tensors), v, H0=m["h0"], inplace=True)
            expectation(m["cos2"], v)
            return v
        for i in range(3):
            v = step(i, v)
        stream.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        for i in range(3, 3 + K):
            v = step(i, v)
        stream.synchronize()
        out[idx] = time.perf_counter() - t0


barrier = threading.Barrier(nthreads)
out = [0.0] * nthreads
ths = [threading.Thread(target=worker, args=(i, out)) for i in range(nthreads)]
for t in ths:
    t.start()
for t in ths:
    t.join()
dt = max(out)
print(f"{wl}: {nthreads} co-running batches of {total // nthreads} states: {dt / K * 1e3:.3f} ms/step, "
      f"{total * K / dt:.0f} state-steps/s aggregate")
