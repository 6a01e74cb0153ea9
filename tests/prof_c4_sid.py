"""Ad-hoc profiling helper (not a test): phases of one receptor zone's setInterpData in the C4 case."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cassiopee_b200 import _lib, workloads as W

ctx = _lib.default_context()
blocks = W.c4_layout((2, 2, 2), 100)
grids = [W.c4_block_grid(b).centers() for b in blocks]
zones = [_lib.Zone(ctx, g.ni, g.nj, g.nk, g.x, g.y, g.z) for g in grids]
rb = blocks[0]; g = grids[0]
sel = np.nonzero(W.c4_receptor_mask(rb, blocks) == 2.)[0]
xr, yr, zr = g.x[sel].copy(), g.y[sel].copy(), g.z[sel].copy()
donors = zones[1:]
for it in range(4):
    ctx.sync(); t0 = time.perf_counter()
    res = _lib.set_interp_data(ctx, xr, yr, zr, donors, 2, 1, 1, 1)
    t1 = time.perf_counter()
    out = res.fetch()
    t2 = time.perf_counter()
    st = res.stats()
    del res
    t3 = time.perf_counter()
    print("n=%d call %.2f ms (search kernel %.2f, extrap %.2f) fetch %.2f ms destroy %.2f ms orphans %d ext %d" %
          (sel.size, (t1 - t0) * 1e3, st['search_ms'], st['extrap_ms'], (t2 - t1) * 1e3, (t3 - t2) * 1e3, out[5].size,
           sum(e.size for e in out[4])))
