cd $GRAFT_REPO_ROOT
python tests/prof_lagrange.py 3 1 2>&1 | tail -5
python tests/prof_lagrange.py 5 16 2>&1 | tail -5
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np
from cassiopee_b200 import _lib, workloads as W
ctx = _lib.default_context()
Bg, Oc = W.c2(1.0)
t0 = time.perf_counter(); zone = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z, interpDataType=0); ctx.sync()
print("cart zone create %.3f s" % (time.perf_counter() - t0))
n = Oc.npts
dx = _lib.DeviceArray.from_host(ctx, Oc.x); dy = _lib.DeviceArray.from_host(ctx, Oc.y); dz = _lib.DeviceArray.from_host(ctx, Oc.z)
for order in (2, 3):
    for it in range(3):
        ctx.sync(); t0 = time.perf_counter()
        res = _lib.set_interp_data_dev(ctx, n, dx.ptr, dy.ptr, dz.ptr, [zone], order, 1, 1, 1)
        ctx.sync(); dt = time.perf_counter() - t0
        print("InterpCart order %d: %d receptors %.2f ms -> %.3g pts/s (search kernel %.2f ms)" % (order, n, dt * 1e3, n / dt, res.stats()["search_ms"]), flush=True)
        res = None
PY
