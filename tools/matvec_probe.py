"""Profiling driver: a few full-batch H.Psi launches (rmb_matvec) on a bench operator.

    python tools/matvec_probe.py [ocs|h2o] [nstates]

Prints the kernel-only time of each launch (CUDA events around the matvec kernel, rmb_matvec_timing); run it
under `ncu -k regex:k_matvec_...` to capture the kernel."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from richmol_b200 import _lib
from richmol_b200.field import _stream_ptr

workload = sys.argv[1] if len(sys.argv) > 1 else "ocs"
nst = int(sys.argv[2]) if len(sys.argv) > 2 else (8192 if workload == "ocs" else 500)
m = bench.build_model(workload)
dc, ac = bench.fields_at(100)
m["Hdc"].field(dc)
m["Hac"].field(ac, thresh=1e1)
H = m["Hdc"] + m["Hac"]
op = H._device()
N = H._basis().N
x = torch.randn(nst, N, dtype=torch.complex128, device="cuda")
y = torch.empty_like(x)
lib = _lib.lib()
ms, cnt = C.c_double(), C.c_int64()
for it in range(4):
    lib.rmb_matvec_timing(op.handle, 1, C.byref(ms), C.byref(cnt))       # enable + reset
    _lib.check(lib.rmb_matvec(op.handle, x.data_ptr(), y.data_ptr(), nst, N, _stream_ptr()))
    torch.cuda.synchronize()
    lib.rmb_matvec_timing(op.handle, 1, C.byref(ms), C.byref(cnt))
    print("matvec kernel ms", ms.value, "launches", cnt.value)
