"""Profiling driver for the linear-rotor matvec: a few rmb_matvec launches on the OCS bench operator."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from richmol_b200 import _lib
from richmol_b200.field import _stream_ptr

nst = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = bench.build_model("ocs")
dc, ac = bench.fields_at(100)
m["Hdc"].field(dc)
m["Hac"].field(ac)
H = m["Hdc"] + m["Hac"]
op = H._device()
N = H._basis().N
x = torch.randn(nst, N, dtype=torch.complex128, device="cuda")
y = torch.empty_like(x)
lib = _lib.lib()
import ctypes as C
ms, cnt = C.c_double(), C.c_int64()
for it in range(4):
    lib.rmb_matvec_timing(op.handle, 1, C.byref(ms), C.byref(cnt))       # enable + reset
    _lib.check(lib.rmb_matvec(op.handle, x.data_ptr(), y.data_ptr(), nst, N, _stream_ptr()))
    torch.cuda.synchronize()
    lib.rmb_matvec_timing(op.handle, 1, C.byref(ms), C.byref(cnt))
    print("matvec kernel ms", ms.value, "launches", cnt.value)
