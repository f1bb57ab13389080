"""Profiling driver: a few full-batch H.Psi launches (rmb_matvec) on a bench operator.

    python tools/matvec_probe.py [workload of bench.py] [nstates] [step]

Prints the kernel-only time of each launch (CUDA events around the matvec kernels, rmb_matvec_timing) and the
algorithmic FP64 / HBM rates; run it under `ncu -k regex:k_matvec_...` to capture a kernel."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from richmol_b200 import _lib
from richmol_b200.field import _stream_ptr

workload = sys.argv[1] if len(sys.argv) > 1 else "ocs_batch"
w = bench.WORKLOADS[workload]()
nst = int(sys.argv[2]) if len(sys.argv) > 2 else w.nstates
step = int(sys.argv[3]) if len(sys.argv) > 3 else 100
m = bench.build_model(w)
for t in m["terms"]:
    if t["static"] is None:
        kw = {} if t["thresh"] is None else dict(thresh=t["thresh"])
        t["tensor"].field(w.field(t["name"], step), **kw)
H = bench.hamiltonian([t["tensor"] for t in m["terms"]])
op = H._device()
N = H._basis().N
x = torch.randn(nst, N, dtype=torch.complex128, device="cuda")
y = torch.empty_like(x)
lib = _lib.lib()
info = bench.op_info(op)
print("routing", {k: info[k] for k in ("tiled", "dmma", "scalar", "lin_T", "fused", "dk_max")}, "N", N, "states", nst)
ms, cnt = C.c_double(), C.c_int64()
for it in range(4):
    lib.rmb_matvec_timing(op.handle, 1, C.byref(ms), C.byref(cnt))       # enable + reset
    _lib.check(lib.rmb_matvec(op.handle, x.data_ptr(), y.data_ptr(), nst, N, _stream_ptr()))
    torch.cuda.synchronize()
    lib.rmb_matvec_timing(op.handle, 1, C.byref(ms), C.byref(cnt))
    print(f"matvec kernels {ms.value:.3f} ms ({cnt.value} launch): {info['flops_per_state'] * nst / ms.value / 1e9:.2f} TFLOP/s, "
          f"{(32.0 * N * nst + info['op_bytes']) / ms.value / 1e6:.0f} GB/s algorithmic")
