"""bench_c2n.py — multi-GPU leg of bench.py: BASELINE.json configs[1] (C2) weak-scaled.

N=1 of bench.py is C2a: one Cartesian background 256^3 + one cylinder O-grid
(17.0 M receptors).  For N GPUs the same pair is instantiated once per GPU
("N bodies in a row"), consecutive backgrounds overlap by 8 cells and exchange
their 2-layer fringes, so the path has its real exchange step: zones are
placed by `.Solver#Param/proc` (each body's two zones on one GPU, what
Distributor2 does to minimise communication), the distributed
Connector.Mpi._setInterpData replicates the neighbouring backgrounds and
routes the ID_* subregions to the donor owners, and every transfer step runs
the local plans, the remote plans into per-peer send buffers, one grouped
NCCL send/recv and one scatter launch.  Per-GPU work is fixed (scaling:
"weak"); value = algorithmic bytes of ALL ranks / max-over-ranks time.

The C4/C5 configs (64-zone cluster, 7 variables, 100-iteration loop) are
measured in the same run and reported under "c4_c5" (bench_c4.measure_c4).
"""
import json
import os
import time

import numpy as np


def build_local(bodies, rank, nvars):
    import cassiopee_b200.Converter as C
    from cassiopee_b200 import workloads as W
    zones = []
    for b in bodies:
        if b['r'] != rank:
            continue
        for z in W.c2n_zones(b):
            for i, v in enumerate(W.VARS7[:nvars]):
                C._initVars(z, v, lambda x, y, zz, i=i: 1. + 0.1 * (i + 1) * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
                            + 0.05 * (i + 1) * zz)
            zones.append(z)
    return C.newPyTree(['Base'] + zones)


def run_c2n(args):
    from bench import METRIC, UNIT, ClockSampler, measured_peaks, NVARS_C2, config_c2
    from bench_c4 import _warm, measure_c4, strip_interp_data, linear_field_check
    from cassiopee_b200 import _lib, Mpi as Cmpi, Distributor2 as D2, workloads as W
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Connector.Mpi as Xmpi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ngpus = max(args.gpus, world)
    ctx = _lib.default_context()
    comm = Cmpi.Comm(ctx) if world > 1 else None
    nvars = args.nvars or NVARS_C2
    bodies = W.c2n_layout(ngpus if world > 1 else 1, args.scale)
    t0 = time.perf_counter()
    t = build_local(bodies, rank, nvars)
    for z in Internal.getZones(t):
        D2._setProc(z, rank)
    t_gen = time.perf_counter() - t0
    zoneOrder = [n for b in bodies for n in (b['bg'], b['og'])]
    names = W.VARS7[:nvars]

    def barrier():
        ctx.sync()
        if comm is not None:
            comm.barrier()

    def gather(v):
        return comm.allgather_obj(v) if comm is not None else [v]

    _warm(ctx)
    # ---- distributed setInterpData (receptor owner computes; neighbour backgrounds replicated) ----
    # called twice: the first call also pays the growth of the device memory pool (see bench_c4.py)
    sid_calls = []
    for call in range(2):
        strip_interp_data(t)
        barrier()
        t0 = time.perf_counter()
        Xmpi._setInterpData(t, t, order=2, nature=1, penalty=1, extrap=1, loc='nodes', storage='inverse', sameBase=1,
                            itype='chimera',
                            comm=comm, zoneOrder=zoneOrder, ctx=ctx, verbose=0)
        barrier()
        sid_calls.append(time.perf_counter() - t0)
    sid_s = sid_calls[-1]
    nrcv_local = sum(int((C.getFieldArray(z, 'cellN') == 2.).sum()) for z in Internal.getZones(t))
    norphan = 0
    for z in Internal.getZones(t):
        for s in Internal.getNodesFromType1(z, 'ZoneSubRegion_t'):
            o = Internal.getNodeFromName1(s, 'OrphanPointList')
            if o is not None:
                norphan = max(norphan, int(o[1].size))

    # ---- steady-state transfers ------------------------------------------------------------------
    sess = Xmpi.TransferSession(t, t, names, comm=comm, ctx=ctx)
    sess.step(max(3, args.warmup))
    barrier()
    sampler = ClockSampler(ctx.device)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    barrier()
    ctx.event_record(0)
    sess.step(args.steps)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms()
    barrier()
    launches = ctx.launches() - l0
    sess.step(600)          # keep the GPUs busy so nvidia-smi samples clocks under load
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: host donor fields -> device, step, receptor fields -> host (tree arrays pinned in place) ----
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

    check = linear_field_check(sess, t, t, names[0], 'nodes', ctx, comm)
    B = sum(gather(sess.alg_bytes()))
    ms_max = max(gather(ms))
    e2e_max = max(gather(e2e_s))
    e2e_ph = np.max(np.array(gather(e2e_phases)), axis=0)
    sid_max = max(gather(sid_s))
    sid_first = max(gather(sid_calls[0]))
    nrcv = sum(gather(nrcv_local))
    nccl_b = sum(gather(sess.nccl_bytes()))
    h2d = sum(gather(h2d_e)); d2h = sum(gather(d2h))
    nplans = sum(gather(len(sess.plans)))
    transport = sess.transport
    launches_all = sum(gather(launches))
    norphan = sum(gather(norphan))
    # free the C2 session before the C4/C5 leg
    sess.close()
    del sess, t
    import gc
    gc.collect()
    c4 = measure_c4(args, ctx, comm, ngpus, rank, world) if not args.quick else None
    if rank != 0:
        return
    ms_step = ms_max / args.steps
    value = B / (ms_step * 1e-3) / 1e9
    peak, how = measured_peaks()
    b0 = bodies[0]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpus, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_c2(nvars, args.scale, ngpus),
        "stats": {"zones": 2 * len(bodies), "receptors": int(nrcv), "plans": int(nplans), "alg_bytes_per_step": int(B),
                  "exchange_bytes_per_step": int(nccl_b), "orphans": int(norphan),
                  "transport": {"p2p": "peer memory: remote plans store into the receptor rank's IPC-mapped buffers over "
                                       "NVLink, flag per rank pair (no NCCL call in the step)",
                                "nccl": "grouped ncclSend/ncclRecv of one packed buffer per rank pair",
                                "none": "single rank"}[transport]},
        "check": check,
        "clocks": clocks,
        "e2e": {"value": B / e2e_max / 1e9, "unit": UNIT, "ms_per_step": e2e_max * 1e3,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "phases_ms_max_over_ranks": {"upload": float(e2e_ph[0]), "step": float(e2e_ph[1]), "download": float(e2e_ph[2])},
                "api": "Connector.Mpi.TransferSession.transfer_host(): host trees in, host trees out (tree arrays "
                       "pinned in place, copies inside the timed region); the donor points the plans read go up, the "
                       "receptor points they wrote come back, pipelined per variable over three streams (the phases are "
                       "those of the three separate calls upload_donor_fields / step / download)"},
        "gpu_launches": int(launches_all),
        "roofline": {"bound": "hbm", "achieved": value / ngpus, "peak": peak, "unit": "GB/s per GPU",
                     "frac": value / ngpus / peak, "traffic": None, "peak_source": how,
                     "kernel": "k_transfer2s<0> (local C2a plan) + k_transfer_set<2> (fringe plans) + NCCL send/recv + "
                               "k_scatter_set"},
        "cpu_baseline": None,
        "set_interp_data": {"value": nrcv / sid_max, "unit": "receptor pts/s", "receptors": int(nrcv), "s": sid_max,
                            "s_first_call": sid_first,
                            "note": "second call (device memory pool warm); the first call is reported beside it",
                            "includes": "host trees in/out: donor-zone replication, device ADT builds, search, packing, "
                                        "D2H of the lists, routing of ID_* subregions to donor owners",
                            "tree_build_s": t_gen},
        "c4_c5": c4,
    }
    print(json.dumps(line))
