"""Ad-hoc timing helper (not a test): phases of the host-pointer setInterpData path on C2a."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cassiopee_b200 import _lib, workloads as W
ctx = _lib.default_context()
Bg, Oc = W.c2(1.0)
for it in range(3):
    ctx.sync(); t0 = time.perf_counter()
    zone = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z)
    ctx.sync(); t1 = time.perf_counter()
    res = _lib.set_interp_data(ctx, Oc.x, Oc.y, Oc.z, [zone], 2, 1, 1, 1)
    ctx.sync(); t2 = time.perf_counter()
    out = res.fetch()
    t3 = time.perf_counter()
    del res, zone, out
    ctx.sync(); t4 = time.perf_counter()
    print("zone %.1f ms | set_interp_data(host) %.1f ms | fetch %.1f ms | free %.1f ms | total %.1f ms" %
          ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3), flush=True)
