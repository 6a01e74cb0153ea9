"""Experiment: cost of the scattered receptor stores of the C2a transfer (mode 0, in place) against the same
kernel writing a dense block in original entry order (mode 1) and in the plan's sorted order (mode 2, coalesced)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cassiopee_b200 import _lib, workloads as W
ctx = _lib.default_context()
Bg, Oc = W.c2(1.0)
zone = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z)
n = Oc.npts
out = _lib.set_interp_data(ctx, Oc.x, Oc.y, Oc.z, [zone], 2, 1, 1, 1).fetch()
rcv, dnr, typ, cf = out[0][0], out[1][0], out[2][0], out[3][0]
plan = _lib.Plan(ctx, Bg.ni, Bg.nj, Bg.nk, n, dnr, rcv, typ, cf)
nv = 5
fD = W.conservative_fields(Bg.x, Bg.y, Bg.z, nv)
dD = [_lib.DeviceArray.from_host(ctx, f) for f in fD]
dR = [_lib.DeviceArray(ctx, n) for _ in range(nv)]
dense = _lib.DeviceArray(ctx, n * nv)
pD = [d.ptr for d in dD]; pR = [d.ptr for d in dR]
B = plan.alg_bytes(nv)
sets = {"mode0_inplace_scatter": _lib.PlanSet(ctx, nv, [(plan, 0, pD, pR, None, 0)]),
        "mode1_dense_entry_order": _lib.PlanSet(ctx, nv, [(plan, 1, pD, None, dense.ptr, n)]),
        "mode2_dense_sorted_order": _lib.PlanSet(ctx, nv, [(plan, 2, pD, None, dense.ptr, n)])}
for name, ps in sets.items():
    for _ in range(5): ps.run()
    ctx.sync(); ctx.event_record(0)
    for _ in range(20): ps.run()
    ctx.event_record(1); ms = ctx.event_elapsed_ms() / 20
    print("%-28s %.4f ms  %.1f GB/s (B_alg)" % (name, ms, B / ms / 1e6), flush=True)
