#!/usr/bin/env python
"""bench.py — chimera interpolation hot path on B200 (driver contract).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one setInterpTransfers pass over all (donor zone, receptor zone)
plans of the workload with the interpData already resident in HBM.
  N=1  workload C2 (BASELINE.json configs[1]): Cartesian background 256^3 +
       cylinder O-grid (513,129,257), order 2, ~17 M receptors, 5 variables.
  N>1  the same C2 pair once per GPU ("N bodies in a row", weak scaling), the
       neighbouring backgrounds overlapping by 8 cells and exchanging their
       fringes over NCCL (bench_c2n.py); in the same run the C4/C5 workload
       (configs[3],[4]: 64-zone cluster at N=8, 8 zones of 100^3 per GPU,
       7 variables, 100-iteration loop) is measured and reported under "c4_c5"
       (bench_c4.py; alone with --workload c4).
metric = setInterpTransfers algorithmic GB/s (B_alg of SURVEY.md 8(d) / time);
setInterpData receptor pts/s is reported in the same line under
"set_interp_data".  The reference arm times the reference's own C++
(oracle/_ref/libcassref.so, compiled from /root/reference) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "setInterpTransfers_algorithmic_GBps"
UNIT = "GB/s"
NVARS_C2 = 5


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, as extracted from an
    `ncu --set full` capture of this very command by tools/ncu_traffic.py into profiles/ncu_traffic.json (committed
    next to the .ncu-rep summary it came from).  None when no capture of that workload is on file: never a constant
    typed into the source."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        e = json.load(open(p))[key]
        return int(e["bytes_per_launch"]), "%s (%s, kernel %s)" % (e["source"], e["captured"], e["kernel"])
    except Exception:
        return None, None


def config_c2(nvars, scale, ngpus=1):
    """`config` of the C2 lines: static description only, so that both arms (ours, --impl reference) emit the same
    dict for the same command; measured counts go under "stats"."""
    if ngpus == 1:
        wl = ("C2a: cart 256^3 background -> cylinder O-grid (513,129,257) receptors, order 2, nature=1 penalty=1 "
              "extrap=1, %d vars" % nvars) if scale == 1.0 else "C2a scaled x%g, %d vars" % (scale, nvars)
        par = "1 GPU"
    else:
        wl = ("C2a per GPU (cart 256^3 background -> cylinder O-grid (513,129,257) receptors, order 2, nature=1 "
              "penalty=1 extrap=1, %d vars), %d bodies in a row, neighbouring backgrounds overlap by 8 cells and "
              "exchange 2-layer fringes GPU to GPU; C4/C5 cluster under c4_c5" % (nvars, ngpus))
        if scale != 1.0: wl = "scaled x%g: " % scale + wl
        par = "one body (2 zones) per GPU by .Solver#Param/proc"
    return {"workload": wl, "nvars": nvars, "n_gpus": ngpus,
            "l2": "inputs larger than L2 (about 2.3 GB of plan + field traffic per GPU and step, L2 126 MB); no flush",
            "parallelism": par}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------
# workload C2 (single GPU)
# --------------------------------------------------------------------------
def build_c2(scale):
    from cassiopee_b200 import workloads as W
    Bg, Oc = W.c2(scale)
    return Bg, Oc


def alg_bytes(n, ncoef, nvars, distinct):
    """B_alg of SURVEY.md 8(d): entries*(8*ncf+12) + 8*N*n + 8*N*U."""
    return 8 * ncoef + 12 * n + 8 * nvars * n + 8 * nvars * distinct


def run_c2(args):
    from cassiopee_b200 import _lib
    from cassiopee_b200 import workloads as W
    ctx = _lib.default_context(0)
    t0 = time.perf_counter()
    Bg, Oc = build_c2(args.scale)
    t_gen = time.perf_counter() - t0
    nrcv = Oc.npts
    # donor zone on the device (= the ADT hook): upload + level-synchronous ADT build
    t0 = time.perf_counter()
    zone = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z)
    t_zone = time.perf_counter() - t0
    zinfo = zone.info()

    # ---- setInterpData: device-resident receptor coordinates -----------------
    dxr = _lib.DeviceArray.from_host(ctx, Oc.x); dyr = _lib.DeviceArray.from_host(ctx, Oc.y)
    dzr = _lib.DeviceArray.from_host(ctx, Oc.z)
    sid_ms = []
    res = None
    for it in range(args.sid_warmup + args.sid_steps):
        res = None
        ctx.sync()
        l0 = ctx.launches()
        t0 = time.perf_counter()
        res = _lib.set_interp_data_dev(ctx, nrcv, dxr.ptr, dyr.ptr, dzr.ptr, [zone], 2, 1, 1, 1)
        ctx.sync()
        dt = time.perf_counter() - t0
        if it >= args.sid_warmup:
            sid_ms.append(dt * 1e3)
        sid_launches = ctx.launches() - l0
    st = res.stats()
    # end to end through the host-pointer C ABI: H2D receptor coordinates + ADT build + search + D2H of the 7-list
    if args.quick:
        t0 = time.perf_counter()
        out = res.fetch()
        sid_e2e_s = time.perf_counter() - t0
    else:
        # host buffers are page-locked (pinned) numpy arrays; one untimed warm-up call, then the timed call
        hx = [_lib.pinned_copy(a) for a in (Bg.x, Bg.y, Bg.z)]
        hr = [_lib.pinned_copy(a) for a in (Oc.x, Oc.y, Oc.z)]
        res = None
        e2e_calls = []
        for it in range(3):
            out = None
            ctx.sync()
            t0 = time.perf_counter()
            zone_e = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, hx[0], hx[1], hx[2])
            t1 = time.perf_counter()
            res_e = _lib.set_interp_data(ctx, hr[0], hr[1], hr[2], [zone_e], 2, 1, 1, 1)
            t2 = time.perf_counter()
            out = res_e.fetch(pinned=True)
            sid_e2e_s = time.perf_counter() - t0
            e2e_calls.append({"total_s": sid_e2e_s, "zone_upload_adt_s": t1 - t0, "search_s": t2 - t1,
                              "fetch_s": sid_e2e_s - (t2 - t0)})
            del zone_e, res_e
        del hx, hr
    rcv, dnr, typ, coefs = out[0][0], out[1][0], out[2][0], out[3][0]
    n = rcv.size
    sid = {"value": nrcv / (np.mean(sid_ms) * 1e-3), "unit": "receptor pts/s", "receptors": int(nrcv),
           "ms": float(np.mean(sid_ms)), "search_kernel_ms": float(st["search_ms"]),
           "nodes_visited_per_query": float(st["nodes_visited"]) / max(1, nrcv),
           "adt_build_ms": float(zinfo["build_ms"]), "adt_depth": int(zinfo["depth"]),
           "adt_cells": int(zinfo["ncells"]), "zone_create_s": t_zone,
           "e2e": {"value": nrcv / sid_e2e_s, "unit": "receptor pts/s",
                   "includes": "pinned host buffers: H2D donor+receptor coords, ADT build, search, packing, D2H of "
                               "index/type/coef lists (3rd call; pinned output pool and device memory pool warm)",
                   "calls": (e2e_calls if not args.quick else None),
                   "h2d_bytes": int(8 * 3 * (Bg.npts + nrcv)), "d2h_bytes": int(12 * n + 8 * coefs.size)},
           "interpolated": int(n), "orphans": int(out[5].size), "gpu_launches": int(sid_launches)}
    res = None

    # ---- transfers: plan + device-resident fields -----------------------------
    nvars = args.nvars or NVARS_C2
    plan = _lib.Plan(ctx, Bg.ni, Bg.nj, Bg.nk, nrcv, dnr, rcv, typ, coefs)
    pinfo = plan.info()
    B = alg_bytes(n, coefs.size, nvars, pinfo["distinct"])
    fD = W.conservative_fields(Bg.x, Bg.y, Bg.z, nvars)
    dD = [_lib.DeviceArray.from_host(ctx, f) for f in fD]
    dR = [_lib.DeviceArray(ctx, nrcv) for _ in range(nvars)]
    pD = [d.ptr for d in dD]; pR = [d.ptr for d in dR]
    for _ in range(max(3, args.warmup)):
        plan.transfer_dev(pD, pR)
    ctx.sync()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = ctx.launches()
    ctx.event_record(0)
    for _ in range(args.steps):
        plan.transfer_dev(pD, pR)
    ctx.event_record(1)
    ms_total = ctx.event_elapsed_ms()
    ctx.sync()
    launches = ctx.launches() - l0
    ms_step = ms_total / args.steps
    # keep the GPU busy a little longer so nvidia-smi gets samples under load
    t_end = time.perf_counter() + 0.6
    while time.perf_counter() < t_end:
        for _ in range(20):
            plan.transfer_dev(pD, pR)
        ctx.sync()
    clocks = sampler.stop()
    value = B / (ms_step * 1e-3) / 1e9

    # ---- end to end through the public API (host buffers, copies inside) ------
    if args.quick:
        e2e = None; cpu = None
    else:
        hD = [_lib.pinned_copy(f) for f in fD]
        hR = [_lib.pinned_empty(nrcv) for _ in range(nvars)]
        for a in hR:
            a[:] = 0.
        ne2e = max(2, min(args.steps, 5))
        plan.transfer(hD, hR); plan.transfer(hD, hR)     # first call: the plan's donor list and staging blocks
        t0 = time.perf_counter()
        for _ in range(ne2e):
            plan.transfer(hD, hR)
        e2e_s = (time.perf_counter() - t0) / ne2e
        e2e = {"value": B / e2e_s / 1e9, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(8 * nvars * plan.upload_points()), "d2h_bytes_per_step": int(8 * nvars * n),
               "donor_points_uploaded": int(plan.upload_points()), "donor_points": int(Bg.npts),
               "api": "cx_transfer (connector._setInterpTransfers boundary), pinned host fields in and out; only the "
                      "donor points the plan reads cross the bus: their index list when that is < 1/4 of the zone, else "
                      "their bounding box as one strided copy per variable when it holds < 85 % of the zone"}
        # ---- CPU baseline: the reference's own transfer loops on the host cores ----
        cpu = cpu_baseline_transfer(fD, nrcv, Bg, rcv, dnr, typ, coefs, B, nvars)

    extra = None if args.quick else extra_legs(ctx, Bg, Oc, zone, dxr, dyr, dzr, pD, nvars)
    peak, how = measured_peaks()
    traffic, traffic_src = ncu_traffic("c2a_%dvars" % nvars) if args.scale == 1.0 else (None, None)
    # the multi-GPU config of BASELINE.json (C4/C5) at its N=1 point: 8 zones of 100^3 on this GPU, so that the
    # per-N lines of a scaling run can be compared like with like
    c4 = None
    if not args.quick and args.scale == 1.0:
        del plan, dD, dR, zone, dxr, dyr, dzr
        import gc
        gc.collect()
        from bench_c4 import measure_c4
        try:
            c4 = measure_c4(args, ctx, None, 1, 0, 1)
        except Exception as e:
            c4 = {"unavailable": str(e)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_c2(nvars, args.scale),
        "stats": {"receptors": int(n), "donor_nodes": int(Bg.npts), "alg_bytes_per_step": int(B),
                  "distinct_donor_nodes": int(pinfo["distinct"])},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": value, "peak": peak, "unit": "GB/s", "frac": value / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": how, "kernel": "k_transfer2s<0>",
                     "frac_of_nominal_8TBps": value / 8000.0},
        "cpu_baseline": cpu,
        "set_interp_data": sid,
        "other_configs": extra,
        "c4_c5": c4,
    }
    print(json.dumps(line))


# GJ flop count of SURVEY.md 8(d) (full in-place inverse + 3 RHS, as the reference's kgaussj does)
def gj_flops(n):
    return n * (2 * (n + 3) * (n - 1) + (n + 3))


def extra_legs(ctx, Bg, Oc, zone, dxr, dyr, dzr, pD, nvars):
    """Device-resident timings of the other BASELINE configs on the C2 geometry (informational; parity for them
    is in tests/): C2b = curvilinear donor (receptors = background nodes inside the annulus <- O-grid);
    C3 = Lagrange order 3 (all receptors) and order 5 (every 16th receptor) + their transfers;
    interpDataType=0 = the reference's closed-form path for Cartesian donors (InterpCart)."""
    from cassiopee_b200 import _lib
    out = {}
    nrcv = Oc.npts
    # --- InterpCart, order 2 and 3
    t0 = time.perf_counter()
    zc = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z, interpDataType=0)
    ctx.sync()
    t_zc = time.perf_counter() - t0
    for order in (2, 3):
        ms = []
        for it in range(4):
            ctx.sync(); t0 = time.perf_counter()
            r = _lib.set_interp_data_dev(ctx, nrcv, dxr.ptr, dyr.ptr, dzr.ptr, [zc], order, 1, 1, 1)
            ctx.sync()
            if it > 0:
                ms.append((time.perf_counter() - t0) * 1e3)
            r = None
        out["interpcart_order%d" % order] = {"value": nrcv / (np.mean(ms) * 1e-3), "unit": "receptor pts/s",
                                              "ms": float(np.mean(ms)), "receptors": int(nrcv), "zone_create_s": t_zc}
    del zc
    # --- C2b: curvilinear donor (the O-grid), receptors = background nodes inside the annulus
    r2 = Bg.x * Bg.x + Bg.y * Bg.y
    selb = np.nonzero((r2 > 0.25) & (r2 < 6.25))[0]
    xb, yb, zb = [np.ascontiguousarray(a[selb]) for a in (Bg.x, Bg.y, Bg.z)]
    t0 = time.perf_counter()
    zo = _lib.Zone(ctx, Oc.ni, Oc.nj, Oc.nk, Oc.x, Oc.y, Oc.z)
    ctx.sync()
    t_zo = time.perf_counter() - t0
    io = zo.info()
    dxb = _lib.DeviceArray.from_host(ctx, xb); dyb = _lib.DeviceArray.from_host(ctx, yb); dzb = _lib.DeviceArray.from_host(ctx, zb)
    ms = []
    for it in range(3):
        ctx.sync(); t0 = time.perf_counter()
        r = _lib.set_interp_data_dev(ctx, selb.size, dxb.ptr, dyb.ptr, dzb.ptr, [zo], 2, 1, 1, 1)
        ctx.sync()
        if it > 0:
            ms.append((time.perf_counter() - t0) * 1e3)
        if it < 2:
            r = None
    st = r.stats()
    o = r.fetch()
    r = None
    planb = _lib.Plan(ctx, Oc.ni, Oc.nj, Oc.nk, selb.size, o[1][0], o[0][0], o[2][0], o[3][0])
    fO = [np.ascontiguousarray(Oc.x + 2 * Oc.y + 3 * Oc.z)] + [np.ascontiguousarray(np.sin(Oc.x * (v + 1)) + Oc.z) for v in range(nvars - 1)]
    dO = [_lib.DeviceArray.from_host(ctx, f) for f in fO]
    dRb = [_lib.DeviceArray(ctx, selb.size) for _ in range(nvars)]
    pO = [d.ptr for d in dO]; pRb = [d.ptr for d in dRb]
    for _ in range(3):
        planb.transfer_dev(pO, pRb)
    ctx.sync(); ctx.event_record(0)
    for _ in range(10):
        planb.transfer_dev(pO, pRb)
    ctx.event_record(1)
    tms = ctx.event_elapsed_ms() / 10
    Bb = planb.alg_bytes(nvars)
    lin = dRb[0].download()
    ok = np.zeros(selb.size, bool); ok[o[0][0]] = True
    ext = np.zeros(selb.size, bool); ext[o[4][0]] = True
    err = float(np.abs(lin - (xb + 2 * yb + 3 * zb))[ok & ~ext].max()) if (ok & ~ext).any() else None
    out["c2b_curvilinear_donor"] = {
        "set_interp_data": {"value": selb.size / (np.mean(ms) * 1e-3), "unit": "receptor pts/s", "ms": float(np.mean(ms)),
                            "receptors": int(selb.size), "interpolated": int(o[0][0].size), "extrapolated": int(o[4][0].size),
                            "orphans": int(o[5].size), "search_kernel_ms": float(st["search_ms"]),
                            "nodes_visited_per_query": float(st["nodes_visited"]) / selb.size,
                            "adt_build_ms": float(io["build_ms"]), "adt_depth": int(io["depth"]), "zone_create_s": t_zo},
        "transfer": {"value": Bb / (tms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": tms, "alg_bytes_per_step": int(Bb),
                     "nvars": nvars, "linear_field_max_abs_error": err}}
    del planb, dO, dRb, zo, dxb, dyb, dzb
    # --- C3: Lagrange orders 3 and 5 through the ADT donor
    for order, stride in ((3, 1), (5, 16)):
        if stride == 1:
            dx, dy, dz, n = dxr, dyr, dzr, nrcv
        else:
            # 1/16 of the O-grid as one block of consecutive k-planes: neighbouring receptors share Lagrange stencils
            # exactly like in the full case (a strided sample would hide that and the per-stencil solve)
            m = nrcv // stride
            dx = _lib.DeviceArray.from_host(ctx, np.ascontiguousarray(Oc.x[:m]))
            dy = _lib.DeviceArray.from_host(ctx, np.ascontiguousarray(Oc.y[:m]))
            dz = _lib.DeviceArray.from_host(ctx, np.ascontiguousarray(Oc.z[:m]))
            n = dx.n
        ctx.sync(); t0 = time.perf_counter()
        r = _lib.set_interp_data_dev(ctx, n, dx.ptr, dy.ptr, dz.ptr, [zone], order, 1, 1, 1)
        ctx.sync()
        dt = time.perf_counter() - t0
        o = r.fetch()
        rcv, dnr, typ, cf = o[0][0], o[1][0], o[2][0], o[3][0]
        r = None
        plan = _lib.Plan(ctx, Bg.ni, Bg.nj, Bg.nk, n, dnr, rcv, typ, cf)
        dR = [_lib.DeviceArray(ctx, n) for _ in range(nvars)]
        pR = [d.ptr for d in dR]
        for _ in range(3):
            plan.transfer_dev(pD, pR)
        ctx.sync(); ctx.event_record(0)
        for _ in range(10):
            plan.transfer_dev(pD, pR)
        ctx.event_record(1)
        tms = ctx.event_elapsed_ms() / 10
        B = plan.alg_bytes(nvars)
        out["lagrange_order%d" % order] = {
            "set_interp_data": {"value": n / dt, "unit": "receptor pts/s", "ms": dt * 1e3, "receptors": int(n),
                                "sample": "all O-grid nodes" if stride == 1 else "the first 1/%d of the O-grid nodes (consecutive k-planes)" % stride,
                                "gj_gflops_reference_count": gj_flops(order ** 3) * n / dt / 1e9,
                                "types": {int(k): int(v) for k, v in zip(*np.unique(typ, return_counts=True))}},
            "transfer": {"value": B / (tms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": tms, "alg_bytes_per_step": int(B),
                         "nvars": nvars}}
        del plan, dR
    return out


def cpu_baseline_transfer(fD, nrcv, Bg, rcv, dnr, typ, coefs, B, nvars, reps=2):
    try:
        from oracle import oracle as O
        fR = [np.zeros(nrcv) for _ in range(nvars)]
        O.transfer(fD, fR, Bg.ni, Bg.nj, rcv, dnr, typ, coefs)
        t0 = time.perf_counter()
        for _ in range(reps):
            O.transfer(fD, fR, Bg.ni, Bg.nj, rcv, dnr, typ, coefs)
        dt = (time.perf_counter() - t0) / reps
        return {"value": B / dt / 1e9, "unit": UNIT, "cores": int(min(nvars, O.num_threads())), "kind": "reference",
                "host_cores": os.cpu_count(), "ms_per_step": dt * 1e3,
                "sample": "full workload, %d timed passes of the reference's commonInterpTransfers_indirect.h loop "
                          "(parallel over variables only: min(nvars, threads) threads)" % reps}
    except Exception as e:  # oracle missing on this box
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}


# --------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# --------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from cassiopee_b200 import workloads as W
    nthreads = os.cpu_count() or 1
    O.set_num_threads(nthreads)
    c4 = None
    if True:
        Bg, Oc = W.c2(args.scale)
        nvars = args.nvars or NVARS_C2
        # (1) setInterpData of the FULL workload with the reference's own code: serial ADT build of the whole background
        # (InterpAdt.cpp:284-433; a few seconds for 16.6 M cells), then the search of every O-grid receptor on all host
        # threads.  Its lists ARE the interpData the transfer loop below runs on: same config as the GPU arm.
        t0 = time.perf_counter()
        zone = O.RefZone(Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z)
        t_adt = time.perf_counter() - t0
        t0 = time.perf_counter()
        out = O.set_interp_data(Oc.x, Oc.y, Oc.z, [zone], 2, 1, 1, 1)
        t_search = time.perf_counter() - t0
        rcv, dnr, typ, coefs = out[0][0], out[1][0], out[2][0], out[3][0]
        sid = {"value": Oc.npts / (t_adt + t_search), "unit": "receptor pts/s", "receptors": int(Oc.npts),
               "adt_build_s": t_adt, "search_s": t_search, "cores": nthreads, "sample": "full C2a workload",
               "note": "hook=None semantics: serial ADT build inside (InterpAdt.cpp:284-433)"}
        other = {}
        try:
            for order, st_ in ((3, 64), (5, 512)):
                m = Oc.npts // st_
                t0 = time.perf_counter()
                O.set_interp_data(Oc.x[:m].copy(), Oc.y[:m].copy(), Oc.z[:m].copy(), [zone], order, 1, 1, 1)
                dtl = time.perf_counter() - t0
                other["lagrange_order%d" % order] = {"value": m / dtl, "unit": "receptor pts/s", "receptors": int(m),
                                                     "s": dtl, "cores": nthreads,
                                                     "sample": "the first 1/%d of the O-grid nodes, full background ADT, "
                                                               "ADT build not included" % st_}
            t0 = time.perf_counter()
            zc = O.RefZone(Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z, interpDataType=0)
            O.set_interp_data(Oc.x, Oc.y, Oc.z, [zc], 2, 1, 1, 1)
            dtc = time.perf_counter() - t0
            other["interpcart_order2"] = {"value": Oc.npts / dtc, "unit": "receptor pts/s", "receptors": int(Oc.npts),
                                          "s": dtc, "cores": nthreads, "sample": "full C2a, InterpCart donor (no tree)"}
            del zc
        except Exception as e:
            other["unavailable"] = str(e)
        del zone, out
        # (2) the reference's transfer loop on that interpData, full size, the steps and warm-ups asked for
        n = rcv.size
        uniq = distinct_nodes(dnr, typ, Bg.ni, Bg.nj, Bg.npts)
        fD = W.conservative_fields(Bg.x, Bg.y, Bg.z, nvars)
        B = alg_bytes(n, coefs.size, nvars, uniq)
        fR = [np.zeros(Oc.npts) for _ in range(nvars)]
        warm = max(3, args.warmup)
        for _ in range(warm):
            O.transfer(fD, fR, Bg.ni, Bg.nj, rcv, dnr, typ, coefs)
        steps = max(1, args.steps)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.transfer(fD, fR, Bg.ni, Bg.nj, rcv, dnr, typ, coefs)
        dt = (time.perf_counter() - t0) / steps
        value = B / dt / 1e9
        cfg = config_c2(nvars, args.scale, args.gpus)
        sample = ("full C2a workload: interpData computed by the reference's own setInterpData (ADT %.1f s + search %.1f s), "
                  "%d timed passes of its transfer loop after %d warm-ups" % (t_adt, t_search, steps, warm))
        stats = {"receptors": int(n), "donor_nodes": int(Bg.npts), "alg_bytes_per_step": int(B),
                 "distinct_donor_nodes": int(uniq)}
        if args.gpus > 1:
            # the N-GPU workload is N such bodies; the host's rate does not depend on how many it is given,
            # so the bounded sample is one body (the fringe exchange has no CPU counterpart in one process)
            sample = ("bounded sample = 1 of the %d bodies of the N-GPU workload (the host's rate does not depend on how "
                      "many bodies it is given): " % args.gpus) + sample
            del fD, fR, coefs, dnr, rcv, typ
            try:
                cv, dt4, st4, cfg4, sid4, nv4 = reference_c4(args, O, W, nthreads)
                c4 = {"value": cv, "unit": UNIT, "ms_per_step": dt4 * 1e3, "config": cfg4, "set_interp_data": sid4,
                      "cores": int(min(nv4, nthreads))}
            except Exception as e:
                c4 = {"unavailable": str(e)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "stats": stats,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(min(nvars, nthreads)), "kind": "reference",
                             "host_cores": os.cpu_count(),
                             "sample": sample + "; transfer threads = min(nvars, host threads)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "set_interp_data": sid, "gpu_launches": 0}
    if c4 is not None:
        line["c4_c5"] = c4
    line["other_configs"] = other
    print(json.dumps(line))


def distinct_nodes(dnr, typ, imd, jmd, npts=None):
    """U of SURVEY.md 8(d) on the host: distinct donor nodes of the expanded stencils (a mark array over the donor
    zone when its size is given, else numpy.unique)."""
    ij = imd * jmd
    if npts is not None:
        mark = np.zeros(int(npts), bool)
        for t, o in ((1, 1), (2, 2), (3, 3), (5, 5)):
            m = dnr[typ == t].astype(np.int64)
            if m.size == 0:
                continue
            for kk in range(o):
                for jj in range(o):
                    for ii in range(o):
                        mark[m + ii + jj * imd + kk * ij] = True
        return int(mark.sum())
    sel = []
    for t, o in ((1, 1), (2, 2), (3, 3), (5, 5)):
        m = dnr[typ == t].astype(np.int64)
        if m.size == 0:
            continue
        for kk in range(o):
            for jj in range(o):
                for ii in range(o):
                    sel.append(m + ii + jj * imd + kk * ij)
    if not sel:
        return 0
    return int(np.unique(np.concatenate(sel)).size)


def reference_c4(args, O, W, nthreads):
    from bench_c4 import reference_c4 as impl
    return impl(args, O, W, nthreads)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the C2 grids (debug only)")
    ap.add_argument("--nvars", type=int, default=0)
    ap.add_argument("--sid-steps", type=int, default=3, dest="sid_steps")
    ap.add_argument("--sid-warmup", type=int, default=1, dest="sid_warmup")
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c2n", "c4"])
    ap.add_argument("--quick", action="store_true", help="skip the e2e and CPU-baseline legs (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly one JSON line: anything a library prints on fd 1 (e.g. NCCL's version banner) is sent
    # to stderr, and print() keeps writing to the real stdout
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)
    if args.workload == "c4":
        from bench_c4 import run_c4
        return run_c4(args)
    if (args.gpus > 1 or world > 1 or args.workload == "c2n") and args.workload != "c2":
        from bench_c2n import run_c2n
        return run_c2n(args)
    return run_c2(args)


if __name__ == "__main__":
    main()
