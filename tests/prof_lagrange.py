"""Ad-hoc timing helper (not a test): setInterpData orders 3/5 on C2a (optionally strided receptors)
and the order-3/5 transfer kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cassiopee_b200 import _lib, workloads as W

order = int(sys.argv[1]); stride = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = _lib.default_context()
Bg, Oc = W.c2(1.0)
zone = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z)
xr, yr, zr = Oc.x[::stride].copy(), Oc.y[::stride].copy(), Oc.z[::stride].copy()
n = xr.size
dx = _lib.DeviceArray.from_host(ctx, xr); dy = _lib.DeviceArray.from_host(ctx, yr); dz = _lib.DeviceArray.from_host(ctx, zr)
for it in range(2):
    ctx.sync(); t0 = time.perf_counter()
    res = _lib.set_interp_data_dev(ctx, n, dx.ptr, dy.ptr, dz.ptr, [zone], order, 1, 1, 1)
    ctx.sync(); dt = time.perf_counter() - t0
    print("order %d: %d receptors, setInterpData %.1f ms -> %.3g receptor pts/s" % (order, n, dt * 1e3, n / dt), flush=True)
out = res.fetch()
rcv, dnr, typ, cf = out[0][0], out[1][0], out[2][0], out[3][0]
print("types", np.unique(typ, return_counts=True))
plan = _lib.Plan(ctx, Bg.ni, Bg.nj, Bg.nk, n, dnr, rcv, typ, cf)
nv = 5
fD = W.conservative_fields(Bg.x, Bg.y, Bg.z, nv)
dD = [_lib.DeviceArray.from_host(ctx, f) for f in fD]; dR = [_lib.DeviceArray(ctx, n) for _ in range(nv)]
pD = [d.ptr for d in dD]; pR = [d.ptr for d in dR]
for _ in range(3): plan.transfer_dev(pD, pR)
ctx.sync(); ctx.event_record(0)
for _ in range(10): plan.transfer_dev(pD, pR)
ctx.event_record(1); ms = ctx.event_elapsed_ms() / 10
B = plan.alg_bytes(nv)
print("order %d transfer: %.3f ms/step, B_alg %.3f GB -> %.1f GB/s" % (order, ms, B / 1e9, B / ms / 1e6))
