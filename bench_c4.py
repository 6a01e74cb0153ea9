"""bench_c4.py — multi-GPU leg of bench.py: workload C4/C5 of BASELINE.json.

64-zone overset cluster (4x4x4 blocks of 100^3 points, 8-cell overlap, per-block
sub-cell shifts), receptors = outer 2 cell layers lying inside another block,
zones placed on the GPUs through `.Solver#Param/proc` (Distributor2), weak
scaling at 8 zones per GPU (N=1: 2x2x2 ... N=8: 4x4x4).  One process per GPU:
distributed setInterpData (receptor owner computes, donor zones replicated,
ID_* subregions routed to the donor owners), then the steady-state
setInterpTransfers loop of 7 variables with the interpData reused: one fused
transfer launch + one packed NCCL send/recv per rank pair + one scatter launch
per step.  Timing: barrier + device sync on both sides, CUDA events per rank,
max over ranks; value = algorithmic bytes of ALL ranks / that time.
"""
import json
import os
import time

import numpy as np

NBLOCKS = {1: (2, 2, 2), 2: (4, 2, 2), 4: (4, 4, 2), 8: (4, 4, 4)}


def layout_for(ngpus, n=100):
    from cassiopee_b200 import workloads as W
    nb = NBLOCKS.get(ngpus)
    if nb is None:
        nb = (2 * ngpus, 2, 2)
    return W.c4_layout(nb, n=n, overlap=8, seed=1), nb


def build_local(blocks, owners, rank, nvars):
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Converter as C
    from cassiopee_b200 import workloads as W
    zones = []
    for b, p in zip(blocks, owners):
        if p != rank:
            continue
        n = b['n']
        z = G.cart(b['origin'], (b['h'],) * 3, (n, n, n), name=b['name'])
        C._initVars(z, 'centers:cellN', W.c4_receptor_mask(b, blocks, depth=2))
        for i, v in enumerate(W.VARS7[:nvars]):
            C._initVars(z, 'centers:' + v,
                        lambda x, y, zz, i=i: 1. + 0.1 * (i + 1) * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + 0.05 * (i + 1) * zz)
        zones.append(z)
    return C.newPyTree(['Base'] + zones)


def config_c4(ngpus, nb, nblk, nvars):
    """Static description of the C4/C5 line (both arms emit the same dict; measured counts go under "stats")."""
    return {"workload": "C4/C5: %dx%dx%d blocks of %d^3 pts (8-cell overlap, sub-cell shifts), receptors = outer 2 cell "
                        "layers inside another block, order 2, %d vars, interpData reused" % (nb[0], nb[1], nb[2], nblk, nvars),
            "zones": nb[0] * nb[1] * nb[2], "zones_per_gpu": nb[0] * nb[1] * nb[2] // max(1, ngpus), "nvars": nvars,
            "n_gpus": ngpus,
            "l2": "per-GPU working set about 110 MB of algorithmic traffic; L2 not flushed (steady-state loop of a solver)",
            "parallelism": "zones over %d GPU(s) by .Solver#Param/proc (rcb)" % ngpus}


def linear_field_check(sess, tR, tD, var, loc, ctx, comm):
    """Correctness of the distributed step at full size, through a size-independent property: with the first variable
    set to F = 1 + x + 2y + 3z on every zone of every rank (donor side and receptor side), one step must leave F at
    every INTERPOLATED receptor point (order-2 weights reproduce linear fields exactly).  Extrapolated points (the
    ExtrapPointList of the ID_* subregions, gathered from the donor owners) are reported apart: their weights, moved
    onto the valid vertices of the donor cell, are not linear-exact by construction (InterpData.cpp:699-947).  Orphans
    are never written.  A wrong index list, exchange buffer or scatter shows up as an error of the order of the cell
    size.  The fields are restored afterwards.  Returns max |error| over all ranks and the number of values compared."""
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Converter as C
    F = lambda x, y, z: 1. + x + 2. * y + 3. * z
    rname = ('centers:' + var) if loc == 'centers' else var
    g = (lambda v: comm.allgather_obj(v)) if comm is not None else (lambda v: [v])
    ext = {}
    for zd in Internal.getZones(tD):
        for sr in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
            e = Internal.getNodeFromName1(sr, 'ExtrapPointList')
            if sr[0].startswith('ID_') and e is not None and e[1].size:
                ext.setdefault(Internal.getValue(sr), []).append(np.asarray(e[1]))
    extAll = {}
    for d in g({k: np.concatenate(v) for k, v in ext.items()}):
        for k, v in d.items(): extAll.setdefault(k, []).append(v)
    saved = []
    for z in Internal.getZones(tD):
        a = C.getFieldArray(z, var); saved.append((a, a.copy(order='K')))
        x, y, zz = C.getCoords(z)
        a[...] = F(x, y, zz)
    exact = []
    for z in Internal.getZones(tR):
        a = C.getFieldArray(z, rname)
        if not any(a is s[0] for s in saved): saved.append((a, a.copy(order='K')))
        x, y, zz = C.getCoords(z)
        if loc == 'centers': x, y, zz = [C.node2CenterArray(v) for v in (x, y, zz)]
        e = F(x, y, zz)
        a[...] = e
        isx = np.zeros(a.size, bool)
        if z[0] in extAll: isx[np.concatenate(extAll[z[0]])] = True
        exact.append((a, e, C.getFieldArray(z, ('centers:cellN') if loc == 'centers' else 'cellN') == 2.,
                      isx.reshape(a.shape, order='F')))
    sess.upload_donor_fields(); ctx.sync()               # donors hold F everywhere
    for a, e, m, isx in exact:
        a[m] = -777.                                     # receptor points start from a wrong value
    sess.upload_receptor_fields()
    sess.step(); sess.download()
    err = 0.; n = 0; nbad = 0; errx = 0.; nx = 0; north = 0
    for a, e, m, isx in exact:
        written = m & ~(np.abs(a + 777.) < 1e-9)         # orphans are never written
        north += int((m & ~written).sum())
        d = np.abs(a - e)
        di = d[written & ~isx]; dx = d[written & isx]
        if di.size:
            err = max(err, float(di.max())); n += int(di.size); nbad += int((di > 1e-9).sum())
        if dx.size:
            errx = max(errx, float(dx.max())); nx += int(dx.size)
    for a, v in saved:
        a[...] = v
    sess.upload_receptor_fields(); sess.upload_donor_fields()
    ctx.sync()
    errs = g(err); ns = g(n); bad = g(nbad); errxs = g(errx); nxs = g(nx); norths = g(north)
    return {"kind": "F = 1+x+2y+3z on all zones of all ranks, one distributed step, |F_interpolated - F| at the "
                    "interpolated receptor points", "max_abs_err": max(errs), "receptor_values_checked": int(sum(ns)),
            "values_off_by_more_than_1e-9": int(sum(bad)), "ok": bool(max(errs) < 1e-9),
            "extrapolated_points": {"n": int(sum(nxs)), "max_abs_err": max(errxs),
                                    "note": "not linear-exact by construction; reported, not judged"},
            "orphans_never_written": int(sum(norths))}


def measure_c4(args, ctx, comm, ngpus, rank, world):
    """C4/C5 on the ranks of this job; returns the result dict on rank 0 (None elsewhere)."""
    from bench import ClockSampler, measured_peaks, ncu_traffic
    from cassiopee_b200 import _lib, Distributor2 as D2, workloads as W
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Connector.Mpi as Xmpi
    nvars = int(os.environ.get("CX_C4_NVARS", "7"))
    nblk = int(os.environ.get("CX_C4_N", "100"))
    blocks, nb = layout_for(ngpus, nblk)
    centres = [np.array(b['origin']) + 0.5 * (b['n'] - 1) * b['h'] for b in blocks]
    owners = D2.distribute_arrays([b['n'] ** 3 for b in blocks], ngpus, centres, 'rcb')['distrib']
    if world == 1 and ngpus > 1:
        owners = [0] * len(blocks)        # --workload c4 on one GPU: all zones local
    tl = build_local(blocks, owners, rank, nvars)
    tcl = C.node2Center(tl)
    zoneOrder = [b['name'] for b in blocks]
    names = W.VARS7[:nvars]

    def barrier():
        ctx.sync()
        if comm is not None:
            comm.barrier()

    # CUDA module load / first-launch costs are not part of the metric: run the smoke-sized case once
    _warm(ctx)
    # ---- distributed setInterpData ---------------------------------------------
    # called twice: the first call also pays the growth of the device memory pool (physical allocation of
    # ~150 MB per donor zone); a solver with moving grids calls it every few steps with the pool warm
    sid_calls = []
    for call in range(3):
        strip_interp_data(tcl)
        barrier()
        t0 = time.perf_counter()
        Xmpi._setInterpData(tl, tcl, order=2, nature=1, penalty=1, extrap=1, loc='centers', storage='inverse', sameBase=1,
                            itype='chimera',
                            comm=comm, zoneOrder=zoneOrder, ctx=ctx, verbose=0)
        barrier()
        sid_calls.append(time.perf_counter() - t0)
    sid_s = min(sid_calls[1:])       # best of the two warm calls; every call is listed in `calls_s`
    nrcv_local = sum(int((C.getFieldArray(z, 'centers:cellN') == 2.).sum()) for z in Internal.getZones(tl))

    # ---- steady-state transfers --------------------------------------------------
    sess = Xmpi.TransferSession(tl, tcl, names, comm=comm, ctx=ctx)
    use_graph = os.environ.get("CX_OVERLAP", "1") != "0"   # comm/compute overlap
    sess.step(max(3, args.warmup), use_graph)
    barrier()
    sampler = ClockSampler(ctx.device)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    barrier()
    ctx.event_record(0)
    sess.step(args.steps, use_graph)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms()
    barrier()
    launches = ctx.launches() - l0
    # C5: the 100-iteration loop with the interpData reused (first 5 calls skipped)
    sess.step(5, use_graph)
    barrier()
    ctx.event_record(0)
    sess.step(100, use_graph)
    ctx.event_record(1)
    ms100 = ctx.event_elapsed_ms()
    barrier()
    # keep the GPUs busy so nvidia-smi samples clocks under load (same count on every rank: NCCL pairs must match)
    sess.step(2000, use_graph)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: host donor fields -> device, step, receptor fields -> host ----
    sess.pin_host_fields()
    ne2e = max(2, min(args.steps, 5))
    e2e_seq = os.environ.get("CX_E2E", "host") == "seq"      # A/B: the three separate calls instead of transfer_host()
    sess.transfer_host(); sess.transfer_host()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(ne2e):
        if e2e_seq:
            sess.upload_donor_fields()
            sess.step()
            h2d_e, d2h = sess.h2d_bytes, sess.download()
        else:
            h2d_e, d2h = sess.transfer_host()
    barrier()
    e2e_s = (time.perf_counter() - t0) / ne2e
    # phase breakdown of one more step (a sync after every phase; diagnostic, outside the timed region)
    barrier()
    tp = [time.perf_counter()]
    sess.upload_donor_fields(); ctx.sync(); tp.append(time.perf_counter())
    sess.step(); ctx.sync(); tp.append(time.perf_counter())
    sess.download(); tp.append(time.perf_counter())
    e2e_phases = [1e3 * (tp[i + 1] - tp[i]) for i in range(3)]
    sess.unpin_host_fields()

    def gather(v):
        return comm.allgather_obj(v) if comm is not None else [v]

    check = linear_field_check(sess, tl, tcl, names[0], 'centers', ctx, comm)
    B = sum(gather(sess.alg_bytes()))
    ms_max = max(gather(ms))
    ms100_max = max(gather(ms100))
    e2e_max = max(gather(e2e_s))
    e2e_ph = np.max(np.array(gather(e2e_phases)), axis=0)
    sid_max = max(gather(sid_s))
    sid_first = max(gather(sid_calls[0]))
    sid_calls_max = np.max(np.array(gather(sid_calls)), axis=0)
    nrcv = sum(gather(nrcv_local))
    nccl_b = sum(gather(sess.nccl_bytes()))
    h2d = sum(gather(h2d_e)); d2h = sum(gather(d2h))
    nplans = sum(gather(len(sess.plans)))
    transport = sess.transport
    launches_all = sum(gather(launches))
    sess.close()
    if rank != 0:
        return None
    cpu = None
    if ngpus == 1 and world == 1 and not args.quick:
        cpu = cpu_baseline_c4(args, blocks, nvars, tree=tcl)
    ms_step = ms_max / args.steps
    value = B / (ms_step * 1e-3) / 1e9
    peak, how = measured_peaks()
    return {
        "value": value, "unit": "GB/s", "n_gpus": ngpus, "steps": args.steps, "ms_per_step": ms_step,
        "config": config_c4(ngpus, nb, nblk, nvars),
        "stats": {"zones": len(blocks), "receptors": int(nrcv), "plans": int(nplans), "alg_bytes_per_step": int(B),
                  "exchange_bytes_per_step": int(nccl_b),
                  "transport": {"p2p": "peer memory: remote plans store into the receptor rank's IPC-mapped buffers over "
                                       "NVLink, flag per rank pair (no NCCL call in the step)",
                                "nccl": "grouped ncclSend/ncclRecv of one packed buffer per rank pair",
                                "none": "single rank"}[transport]},
        "check": check, "cpu_baseline": cpu,
        "c5_100_iterations": {"ms_total": ms100_max, "ms_per_step": ms100_max / 100.,
                              "GBps": B / (ms100_max / 100. * 1e-3) / 1e9},
        "clocks": clocks,
        "e2e": {"value": B / e2e_max / 1e9, "unit": "GB/s", "ms_per_step": e2e_max * 1e3,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "phases_ms_max_over_ranks": {"upload": float(e2e_ph[0]), "step": float(e2e_ph[1]), "download": float(e2e_ph[2])},
                "api": "TransferSession.transfer_host(): host trees in, host trees out; the donor points the plans read go "
                       "up, the receptor points they wrote come back, pipelined per variable over three streams "
                       "(the phases are those of the three separate calls upload_donor_fields / step / download)"},
        "gpu_launches": int(launches_all), "comm_overlap": bool(use_graph),
        "roofline": {"bound": "hbm", "achieved": value / ngpus, "peak": peak, "unit": "GB/s per GPU",
                     "frac": value / ngpus / peak,
                     "traffic": (ncu_traffic("c4_%dvars" % nvars)[0] if (ngpus == 1 and nblk == 100) else None),
                     "traffic_source": (ncu_traffic("c4_%dvars" % nvars)[1] if (ngpus == 1 and nblk == 100) else None),
                     "peak_source": how,
                     "kernel": "k_transfer_set<2> (+ at N>1: remote plans into peer memory, flag exchange, k_scatter_set)"},
        "set_interp_data": {"value": nrcv / sid_max, "unit": "receptor pts/s", "receptors": int(nrcv),
                            "s": sid_max, "s_first_call": sid_first,
                            "calls_s_max_over_ranks": [float(v) for v in sid_calls_max],
                            "note": "best of the two warm calls (device memory pool warm), max over ranks; every call "
                                    "is listed beside it",
                            "includes": "donor-zone replication, device ADT builds, search, "
                            "routing of ID_* subregions to donor owners (host trees in/out)"},
    }


def run_c4(args):
    """`--workload c4`: the C4/C5 leg alone, printed as the bench line."""
    from bench import METRIC, UNIT
    from cassiopee_b200 import _lib, Mpi as Cmpi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ngpus = max(args.gpus, world)
    ctx = _lib.default_context()
    comm = Cmpi.Comm(ctx) if world > 1 else None
    os.environ.setdefault("CX_C4_NVARS", str(args.nvars or 7))
    r = measure_c4(args, ctx, comm, ngpus, rank, world)
    if rank != 0:
        return
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": ngpus, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    for k in ("config", "stats", "check", "clocks", "e2e", "gpu_launches", "comm_overlap", "roofline", "set_interp_data",
              "c5_100_iterations"):
        line[k] = r[k]
    line["cpu_baseline"] = r.get("cpu_baseline")
    print(json.dumps(line))


def strip_interp_data(t):
    """Remove the ID_* subregions of a tree (so that setInterpData can be called again on it)."""
    import cassiopee_b200.Internal as Internal
    for z in Internal.getZones(t):
        z[2][:] = [c for c in z[2] if not (c[3] == 'ZoneSubRegion_t' and c[0].startswith('ID_'))]


def _warm(ctx):
    from cassiopee_b200 import _lib, workloads as W
    a = W.cart('w', (0, 0, 0), (0.1, 0.1, 0.1), (9, 9, 9))
    z = _lib.Zone(ctx, 9, 9, 9, a.x, a.y, a.z)
    r = _lib.set_interp_data(ctx, a.x[:64] + 0.03, a.y[:64] + 0.02, a.z[:64] + 0.01, [z], 2, 1, 1, 1).fetch()
    p = _lib.Plan(ctx, 9, 9, 9, 64, r[1][0], r[0][0], r[2][0], r[3][0])
    p.transferD([a.x])
    ctx.sync()


def _sample_case(blocks, nvars, nrz=2):
    """Bounded sample for the CPU arm: the first nrz receptor zones with their
    bbox-intersecting donor zones."""
    from cassiopee_b200 import workloads as W
    grids = {}
    def grid(b):
        if b['name'] not in grids:
            grids[b['name']] = W.c4_block_grid(b).centers()
        return grids[b['name']]
    out = []
    for rb in blocks[:nrz]:
        g = grid(rb)
        cellN = W.c4_receptor_mask(rb, blocks, depth=2)
        sel = np.nonzero(cellN == 2.)[0]
        bb = g.bbox()
        donors = [b for b in blocks if b is not rb and W.bbox_intersect(bb, grid(b).bbox())]
        out.append((rb, g, sel, donors))
    return out, grid


def cpu_baseline_c4(args, blocks, nvars, nrz=2, tree=None):
    try:
        return reference_c4_impl(args, blocks, nvars, nrz, tree)[0]
    except Exception as e:
        return {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}


def reference_c4_impl(args, blocks, nvars, nrz=2, tree=None):
    """The reference's C++ (oracle/_ref) on a bounded sample of C4 on the host cores.  tree: the donor tree the device
    path filled (N=1): the ID_* subregions of the sampled receptor zones are compared with the reference's lists, list
    by list, at full size."""
    from oracle import oracle as O
    from cassiopee_b200 import workloads as W
    from bench import alg_bytes, distinct_nodes
    sample, grid = _sample_case(blocks, nvars, nrz)
    zones = {}
    t0 = time.perf_counter()
    for rb, g, sel, donors in sample:
        for b in donors:
            if b['name'] not in zones:
                d = grid(b)
                zones[b['name']] = O.RefZone(d.ni, d.nj, d.nk, d.x, d.y, d.z)
    t_adt = time.perf_counter() - t0
    plans = []; plans_named = []
    t0 = time.perf_counter()
    nrcv = 0
    for rb, g, sel, donors in sample:
        out = O.set_interp_data(g.x[sel], g.y[sel], g.z[sel], [zones[b['name']] for b in donors], 2, 1, 1, 1)
        nrcv += sel.size
        for k, b in enumerate(donors):
            if out[0][k].size:
                plans.append((grid(b), sel[out[0][k]].astype(np.int32), out[1][k], out[2][k], out[3][k], g.npts))
                plans_named.append(plans[-1] + (rb['name'],))
    t_search = time.perf_counter() - t0
    B = 0
    fields = {}
    for d, rcv, dnr, typ, cf, nR in plans:
        B += alg_bytes(rcv.size, cf.size, nvars, distinct_nodes(dnr, typ, d.ni, d.nj))
        if d.name not in fields:
            fields[d.name] = W.conservative_fields(d.x, d.y, d.z, nvars)
    fR = [np.zeros(plans[0][5]) for _ in range(nvars)] if plans else []
    steps = max(1, min(args.steps, 20))
    def one_pass():
        for d, rcv, dnr, typ, cf, nR in plans:
            O.transfer(fields[d.name], fR, d.ni, d.nj, rcv, dnr, typ, cf)
    one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = (time.perf_counter() - t0) / steps
    value = B / dt / 1e9
    nth = os.cpu_count() or 1
    desc = ("C4 bounded sample: %d receptor zones (%d receptors) with their %d bbox-intersecting donor zones, one ADT "
            "per donor zone built once (hooks, as Connector/Mpi.py:950-983), %d plans, %d vars" %
            (nrz, nrcv, len(zones), len(plans), nvars))
    cpu = {"value": value, "unit": "GB/s", "cores": int(min(nvars, nth)), "kind": "reference", "host_cores": nth,
           "ms_per_step": dt * 1e3, "sample": desc}
    if tree is not None:
        import cassiopee_b200.Internal as Internal
        zd = {z[0]: z for z in Internal.getZones(tree)}
        nsub = 0; same = True
        # plans were appended receptor zone by receptor zone, donors in list order, empty pairs skipped
        for d, rcv, dnr, typ, cf, nR, rname in plans_named:
            sr = Internal.getNodeFromName1(zd[d.name], 'ID_' + rname) if d.name in zd else None
            if sr is None:
                same = False; continue
            get = lambda nm: Internal.getNodeFromName1(sr, nm)[1]
            same = same and np.array_equal(get('PointListDonor'), rcv) and np.array_equal(get('PointList'), dnr) \
                and np.array_equal(get('InterpolantsType'), typ) and np.array_equal(get('InterpolantsDonor'), cf)
            nsub += 1
        cpu["parity_vs_device_at_full_size"] = {"subregions_compared": nsub, "receptors": int(nrcv),
                                                "lists_identical": bool(same)}
    sid = {"value": nrcv / (t_adt + t_search), "unit": "receptor pts/s", "receptors": int(nrcv), "adt_build_s": t_adt,
           "search_s": t_search, "cores": nth}
    return cpu, sid, dt, steps, desc, B


def reference_c4(args, O, W, nthreads):
    nvars = args.nvars or 7
    blocks, nb = layout_for(args.gpus, int(os.environ.get("CX_C4_N", "100")))
    cpu, sid, dt, steps, desc, B = reference_c4_impl(args, blocks, nvars)
    cfg = {"workload": desc, "nvars": nvars, "alg_bytes_per_step": int(B)}
    return cpu["value"], dt, steps, cfg, sid, nvars
