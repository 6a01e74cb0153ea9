"""Experiment: host<->device copy rates of a numpy array page-locked in place (cudaHostRegister, what
TransferSession.pin_host_fields does) against a cudaHostAlloc block (pinned_empty), both directions."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from cassiopee_b200 import _lib

ctx = _lib.default_context()
n = 85 * 1000 * 1000
d = _lib.DeviceArray(ctx, n)
a = np.ones(n)
b = _lib.pinned_empty(n); b[:] = 1.
_lib.host_register(a)
for name, h in (("registered numpy", a), ("cudaHostAlloc", b)):
    for what, f in (("H2D", lambda: ctx.upload_async(d.ptr, h)), ("D2H", lambda: ctx.download_async(h, d.ptr))):
        f(); ctx.sync()
        t0 = time.perf_counter()
        for _ in range(3): f()
        ctx.sync()
        dt = (time.perf_counter() - t0) / 3
        print("%-18s %s %.2f ms  %.1f GB/s" % (name, what, dt * 1e3, h.nbytes / dt / 1e9))
_lib.host_unregister(a)
