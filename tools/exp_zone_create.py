"""Experiment: cost of creating the device donor zone (upload + ADT build) of one C4 block (100^3 nodes -> 99^3 centres)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cassiopee_b200 import _lib, workloads as W
ctx = _lib.default_context()
g = W.cart('b', (0, 0, 0), (0.01, 0.01, 0.01), (100, 100, 100)).centers()
z = _lib.Zone(ctx, g.ni, g.nj, g.nk, g.x, g.y, g.z)   # warm-up (module load)
for it in range(5):
    ctx.sync(); t0 = time.perf_counter()
    z = _lib.Zone(ctx, g.ni, g.nj, g.nk, g.x, g.y, g.z)
    ctx.sync(); dt = time.perf_counter() - t0
    i = z.info()
    print("zone create %.2f ms (device ADT build %.2f ms, depth %d, %d cells)" % (dt * 1e3, i["build_ms"], i["depth"], i["ncells"]), flush=True)
t0 = time.perf_counter()
zc = _lib.Zone(ctx, g.ni, g.nj, g.nk, g.x, g.y, g.z, interpDataType=0); ctx.sync()
print("cart zone create %.2f ms" % ((time.perf_counter() - t0) * 1e3))
