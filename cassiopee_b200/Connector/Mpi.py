"""Distributed chimera path: zones partitioned across ranks (one process per GPU).

Mirror of the reference's Connector/Connector/Mpi.py for this path:
  _setInterpData     :858-1022  receptor owner computes; intersecting donor zones
                                are replicated first (Cmpi._addXZones, :941), the
                                resulting ID_* subregions are routed to the donor
                                zone's owner (Cmpi.sendRecv, :1012)
  _setInterpTransfers :396-440  donor owner runs setInterpTransfersD; local pairs
                                are scattered in place, remote ones are sent and
                                scattered on arrival
Zone placement is the `.Solver#Param/proc` node (Distributor2).  Bulk data
moves with the library's NCCL communicator (device) or gloo (CPU tests);
see cassiopee_b200/Mpi.py.  TransferSession is the device-resident steady
state: all plans of a rank fused in one launch, one packed NCCL message per
rank pair and iteration, one scatter launch.
"""
import os
import time
import sys as _sys
import numpy
from .. import Internal
from .. import Converter as C
from .. import _lib
from ..workloads import bbox_intersect
from . import OversetData as X


def _zoneBBox(z):
    """Axis-aligned bounding box of a structured zone.  For a 3-D zone the extremes of the (one-to-one) grid mapping
    lie on its six boundary faces, which is also what the reference's K_COMPGEOM::boundingBoxStruct relies on
    (KCore/KCore/CompGeom/compGeom.cpp:446-...), so only the faces are scanned (6 n^2 instead of n^3 values)."""
    x, y, zz = C.getCoords(z)
    if x.ndim == 3 and min(x.shape) >= 3:
        lo = []; hi = []
        for a in (x, y, zz):
            faces = (a[0], a[-1], a[:, 0], a[:, -1], a[:, :, 0], a[:, :, -1])
            lo.append(min(float(f.min()) for f in faces)); hi.append(max(float(f.max()) for f in faces))
        return lo + hi
    return [float(x.min()), float(y.min()), float(zz.min()), float(x.max()), float(y.max()), float(zz.max())]


def _receptorBBox(z, loc):
    return _zoneBBox(z)


def _makeZone(name, ni, nj, nk, x, y, z, cellN):
    zn = Internal.newZone(name, ni, nj, nk)
    gc = Internal.createChild(zn, Internal.__GridCoordinates__, 'GridCoordinates_t')
    shp = (ni, nj, nk)
    for nm, a in (('CoordinateX', x), ('CoordinateY', y), ('CoordinateZ', z)):
        Internal.createChild(gc, nm, 'DataArray_t', numpy.asfortranarray(a.reshape(shp, order='F')))
    fs = Internal.createChild(zn, Internal.__FlowSolutionNodes__, 'FlowSolution_t')
    Internal.createChild(fs, 'cellN', 'DataArray_t', numpy.asfortranarray(cellN.reshape(shp, order='F')))
    return zn


def _baseNames(t):
    """zone name -> name of its CGNSBase_t (None for zones outside a base), Connector/Mpi.py:925-928"""
    out = {}
    for b in Internal.getBases(t):
        for z in Internal.getZones(b): out[z[0]] = b[0]
    return out


def computeDonorGraph(aR, aD, comm, loc='nodes', sameName=1, zoneOrder=None, sameBase=1):
    """Global donor metadata and, per local receptor zone, the ordered list of
    bbox-intersecting donor zones (what Connector/Mpi.py:950-983 builds with
    getIntersectingDomains + hooks).  Order = zoneOrder (names) if given, else
    (rank, position) order, so results do not depend on who computes.
    sameBase=0: only donors of ANOTHER base are kept (Connector/Mpi.py:930-937)."""
    local = []
    boxes = {}          # zone object -> bbox (a tree that is both receptor and donor is reduced once)
    basesD = _baseNames(aD); basesR = _baseNames(aR)
    for z in Internal.getZones(aD):
        _, ni, nj, nk, _ = Internal.getZoneDim(z)
        boxes[id(z)] = _zoneBBox(z)
        local.append(dict(name=z[0], dims=(ni, nj, nk), bbox=boxes[id(z)], rank=comm.rank, base=basesD.get(z[0])))
    allmeta = comm.allgather_obj(local)
    meta = [m for part in allmeta for m in part]
    if zoneOrder is not None:
        pos = {n: i for i, n in enumerate(zoneOrder)}
        meta.sort(key=lambda m: pos[m['name']])
    donorLists = {}
    for z in Internal.getZones(aR):
        bb = boxes[id(z)] if id(z) in boxes else _receptorBBox(z, loc)
        baseR = basesR.get(z[0], basesD.get(z[0]))
        donorLists[z[0]] = [m['name'] for m in meta
                            if not (sameName == 1 and m['name'] == z[0]) and bbox_intersect(bb, m['bbox'])
                            and not (sameBase == 0 and m.get('base') == baseR)]
    return meta, donorLists


def _addXZones(aD, meta, donorLists, comm):
    """Replicate the remote donor zones this rank needs (coordinates + cellN):
    analogue of Cmpi._addXZones(aD, graph, variables=['cellN'])."""
    byname = {m['name']: m for m in meta}
    need = sorted({n for lst in donorLists.values() for n in lst if byname[n]['rank'] != comm.rank})
    allneed = comm.allgather_obj(need)
    localz = {z[0]: z for z in Internal.getZones(aD)}
    send = {}
    for p, lst in enumerate(allneed):
        if p == comm.rank:
            continue
        parts = []
        for n in lst:
            if n in localz:
                z = localz[n]
                x, y, zz = C.getCoords(z)
                c = C.getFieldArray(z, 'cellN')
                npts = x.size
                parts += [x.reshape(npts, order='F'), y.reshape(npts, order='F'), zz.reshape(npts, order='F'),
                          c.reshape(npts, order='F') if c is not None else numpy.ones(npts)]
        if parts:
            send[p] = numpy.concatenate(parts)
    recv_sizes = {}
    for n in need:
        m = byname[n]
        recv_sizes[m['rank']] = recv_sizes.get(m['rank'], 0) + 4 * m['dims'][0] * m['dims'][1] * m['dims'][2]
    got = comm.exchange_host(send, recv_sizes, dtype=numpy.float64)
    added = []
    off = {p: 0 for p in recv_sizes}
    for n in need:     # sorted, same order as the sender's packing
        m = byname[n]; p = m['rank']
        ni, nj, nk = m['dims']; npts = ni * nj * nk
        buf = got[p]; o = off[p]
        zn = _makeZone(n, ni, nj, nk, buf[o:o + npts], buf[o + npts:o + 2 * npts], buf[o + 2 * npts:o + 3 * npts],
                       buf[o + 3 * npts:o + 4 * npts])
        off[p] = o + 4 * npts
        added.append(zn)
    return added


def _addXZonesDevice(aD, meta, donorLists, comm, ctx, interpDataType=1, typeOf=None):
    """Device form of _addXZones: the local donor zones go to the GPU once (coordinates + cellN + ADT = the hook),
    and the zones other ranks need travel GPU to GPU over NCCL together with their ADT replica
    (cx_zone_exchange), so nothing is staged through host memory and no tree is rebuilt on the receiver.
    Returns (bare container zones of the replicated donors, {zone name: device hook})"""
    byname = {m['name']: m for m in meta}
    need = sorted({n for lst in donorLists.values() for n in lst if byname[n]['rank'] != comm.rank})
    allneed = comm.allgather_obj(need)
    localz = {z[0]: z for z in Internal.getZones(aD)}
    wanted = {n for lst in donorLists.values() for n in lst if n in localz}
    requested = {n for p, lst in enumerate(allneed) if p != comm.rank for n in lst if n in localz}
    be = _lib.DeviceBackend(ctx)
    hooks = {}
    names = sorted(wanted | requested)
    specs = []
    for n in names:
        z = localz[n]
        _, ni, nj, nk, _ = Internal.getZoneDim(z)
        x, y, zz = [a.reshape(-1, order='F') for a in C.getCoords(z)]
        c = C.getFieldArray(z, 'cellN')
        specs.append((ni, nj, nk, x, y, zz, c.reshape(-1, order='F') if c is not None else None,
                      typeOf[n] if typeOf is not None else interpDataType))
    for n, h in zip(names, be.make_zones(specs)):      # one batch: the ADT builds run concurrently on the device
        hooks[n] = h
    mine = {}
    for n in requested:
        mi, md = hooks[n].meta()
        mine[n] = (mi.tolist(), md.tolist())
    zmeta = {}
    for part in comm.allgather_obj(mine):
        zmeta.update(part)
    sends = [(hooks[n], p) for p, lst in enumerate(allneed) if p != comm.rank for n in lst if n in localz]
    recvs = [(byname[n]['rank'], numpy.asarray(zmeta[n][0], numpy.int32), numpy.asarray(zmeta[n][1], numpy.float64))
             for n in need]
    got = _lib.zone_exchange(ctx, sends, recvs)
    added = []
    for n, hz in zip(need, got):
        ni, nj, nk = byname[n]['dims']
        added.append(Internal.newZone(n, ni, nj, nk))
        hooks[n] = hz
    return added, hooks


_FIELDS = (('PointList', numpy.int32), ('PointListDonor', numpy.int32), ('InterpolantsType', numpy.int32),
           ('InterpolantsDonor', numpy.float64), ('ExtrapPointList', numpy.int32), ('OrphanPointList', numpy.int32))


def _setInterpData(aR, aD, order=2, penalty=1, nature=0, extrap=1, method='lagrangian', loc='nodes',
                   storage='direct', interpDataType=1, hook=None, cartesian=False, sameBase=0,
                   topTreeRcv=None, topTreeDnr=None, sameName=1, verbose=2, dim=3, itype='abutting',
                   comm=None, zoneOrder=None, backend=None, ctx=None):
    """Distributed setInterpData (signature and defaults of Connector/Mpi.py:858-862).  itype='chimera' with
    storage='inverse' is the path built here: on return every local donor
    zone of aD carries the ID_<receptor> subregions of all receptor zones
    (local or remote) it feeds.  aR / aD hold the zones owned by this rank.
    sameBase=0 (the reference's default, Connector/Mpi.py:858-862, :930-937): a zone is only interpolated from
    zones of OTHER bases; sameBase=1: from every intersecting zone.  Like the reference (Connector/Mpi.py:950-983: the
    hook of a zone of a base named 'CARTESIAN' is None, i.e. interpDataType 0) the donor zones of a 'CARTESIAN' base are
    located in closed form (InterpCart) whatever `interpDataType` says; `cartesian` itself only selects a compressed
    transport of such zones in the reference's Cmpi._addXZones and changes nothing here (zones travel GPU to GPU)."""
    if itype == 'abutting':
        # the reference's default: match (ghost-cell) connectivity only, which has no device path; like
        # OversetData._setInterpData it is an error only when the zones actually carry such connectivity
        for z in Internal.getZones(aR):
            if Internal.getNodesFromType2(z, 'GridConnectivity1to1_t') or Internal.getNodesFromType2(z, 'GridConnectivity_t'):
                raise NotImplementedError("Mpi._setInterpData: abutting (ghost-cell) interpolation data has no device "
                                          "path; call with itype='chimera'.")
        return None
    if itype != 'chimera':
        raise NotImplementedError("Mpi._setInterpData: itype=%r has no device path (itype='chimera')." % (itype,))
    if storage != 'inverse':
        raise NotImplementedError("Mpi._setInterpData: only storage='inverse' is distributed.")
    cartBase = {z[0] for b in Internal.getBases(aD) if b[0] == 'CARTESIAN' for z in Internal.getZones(b)}
    if comm is None or comm.size == 1:
        if cartBase and not isinstance(interpDataType, list):
            interpDataType = [0 if z[0] in cartBase else interpDataType for z in Internal.getZones(aD)]
        donorLists = computeDonorGraph(aR, aD, _Single(), loc, sameName, zoneOrder, sameBase)[1]
        # receptor zones without any candidate donor are skipped, like `if dnrZones != []` (Connector/Mpi.py:975)
        rcvs = [z for z in Internal.getZones(aR) if donorLists.get(z[0])]
        if not rcvs:
            return None
        return X._setInterpDataChimera(rcvs, aD, order=order, penalty=penalty, nature=nature, extrap=extrap,
                                       method=method, loc=loc, storage=storage, interpDataType=interpDataType,
                                       hook=hook, verbose=verbose, sameName=sameName, dim=dim, itype=itype,
                                       ctx=ctx, backend=backend, donorLists=donorLists)
    prof = os.environ.get("CX_PROFILE") and comm.rank == 0
    t0 = time.perf_counter()
    hadCellN = {z[0]: C.isNamePresent(z, 'cellN') for z in Internal.getZones(aD)}
    X._addCellN__(aD, loc='nodes')
    meta, donorLists = computeDonorGraph(aR, aD, comm, loc, sameName, zoneOrder, sameBase)
    byname = {m['name']: m for m in meta}
    t1 = time.perf_counter()
    # CX_ZONE_REPL=host: replicate through host memory like the first version (A/B); default: GPU to GPU
    devrepl = (backend is None and getattr(comm, 'device', False) and not isinstance(interpDataType, list)
               and os.environ.get("CX_ZONE_REPL", "device") != "host")
    # zones of a base named 'CARTESIAN' (on any rank: the base travels in the metadata) are InterpCart donors
    typeOf = None
    if any(m.get('base') == 'CARTESIAN' for m in meta) and not isinstance(interpDataType, list):
        typeOf = {m['name']: (0 if m.get('base') == 'CARTESIAN' else interpDataType) for m in meta}
    hooks = None
    if devrepl:
        added, hooks = _addXZonesDevice(aD, meta, donorLists, comm, ctx or comm.ctx, interpDataType, typeOf)
    else:
        added = _addXZones(aD, meta, donorLists, comm)
    t2 = time.perf_counter()
    # working donor list: local zones + replicated ones, in the global order
    zonesL = {z[0]: z for z in Internal.getZones(aD)}
    work = []
    for m in meta:
        if m['name'] in zonesL: work.append(zonesL[m['name']])
    addedByName = {z[0]: z for z in added}
    for m in meta:
        if m['name'] in addedByName: work.append(addedByName[m['name']])
    order_pos = {m['name']: i for i, m in enumerate(meta)}
    work.sort(key=lambda z: order_pos[z[0]])
    rcvs = [z for z in Internal.getZones(aR) if donorLists.get(z[0])]
    if typeOf is not None:
        interpDataType = [typeOf[z[0]] for z in work]
    if not rcvs:
        pass                       # nothing to interpolate here; the routing collectives below still run
    elif devrepl:
        X._setInterpDataChimera(rcvs, work, order=order, penalty=penalty, nature=nature, extrap=extrap, method=method,
                                loc=loc, storage='inverse', interpDataType=interpDataType,
                                hook=[hooks.get(z[0]) for z in work], verbose=verbose, sameName=sameName, dim=dim,
                                itype='chimera', donorLists=donorLists, ctx=ctx or comm.ctx, backend=backend,
                                deviceOnly={z[0] for z in added}, refreshHooks=False)
    else:
        X._setInterpDataChimera(rcvs, work, order=order, penalty=penalty, nature=nature, extrap=extrap, method=method,
                                loc=loc, storage='inverse', interpDataType=interpDataType, hook=None, verbose=verbose,
                                sameName=sameName, dim=dim, itype='chimera', donorLists=donorLists, ctx=ctx,
                                backend=backend)
    t3 = time.perf_counter()
    # route the subregions created under replicated zones to their owners
    outMeta = {}
    outData = {}
    for z in added:
        p = byname[z[0]]['rank']
        for s in Internal.getNodesFromType1(z, 'ZoneSubRegion_t'):
            if not s[0].startswith('ID_'):
                continue
            rec = dict(donor=z[0], rcv=Internal.getValue(s),
                       centers=Internal.getNodeFromName1(s, 'GridLocation') is not None, sizes=[])
            for name, dt in _FIELDS:
                node = Internal.getNodeFromName1(s, name)
                a = numpy.zeros(0, dt) if node is None else numpy.ascontiguousarray(node[1], dtype=dt)
                rec['sizes'].append(int(a.size))
                outData.setdefault(p, []).append(a.view(numpy.uint8))
            outMeta.setdefault(p, []).append(rec)
    allMeta = comm.allgather_obj(outMeta)
    send = {p: numpy.concatenate(parts) if parts else numpy.zeros(0, numpy.uint8) for p, parts in outData.items()}
    recv_sizes = {}
    for q, om in enumerate(allMeta):
        if q == comm.rank or comm.rank not in om:
            continue
        tot = 0
        for rec in om[comm.rank]:
            for (name, dt), n in zip(_FIELDS, rec['sizes']):
                tot += n * numpy.dtype(dt).itemsize
        recv_sizes[q] = tot
    got = comm.exchange_host(send, recv_sizes, dtype=numpy.uint8)
    for q in sorted(recv_sizes):
        buf = got[q]; o = 0
        for rec in allMeta[q][comm.rank]:
            arrs = []
            for (name, dt), n in zip(_FIELDS, rec['sizes']):
                nb = n * numpy.dtype(dt).itemsize
                arrs.append(buf[o:o + nb].view(dt).copy()); o += nb
            zd = zonesL[rec['donor']]
            X._createInterpRegion__(zd, rec['rcv'], arrs[0], arrs[1], arrs[3], arrs[2], numpy.array([], numpy.float64),
                                    arrs[4], arrs[5], tag='Donor', loc='centers' if rec['centers'] else 'nodes')
    for z in Internal.getZones(aD):
        if hadCellN.get(z[0], 1) == -1: C._rmVars(z, ['cellN'])
    if prof:
        print("[Mpi._setInterpData rank0] graph %.3fs addXZones %.3fs (%d zones) chimera %.3fs routing %.3fs" %
              (t1 - t0, t2 - t1, len(added), t3 - t2, time.perf_counter() - t3), file=_sys.stderr, flush=True)
    return None


def setInterpData2(tR, tD, order=2, loc='centers', cartesian=False, comm=None, ctx=None, backend=None):
    """Interpolation data between two distributed trees; returns the donor tree (Connector/Connector/Mpi.py:1025-1030)."""
    aD = Internal.copyRef(tD)
    aR = Internal.copyRef(tR)
    _setInterpData2(aR, aD, order=order, loc=loc, cartesian=cartesian, comm=comm, ctx=ctx, backend=backend)
    return aD


def _setInterpData2(tR, tD, order=2, loc='centers', cartesian=False, comm=None, ctx=None, backend=None):
    """Distributed form of Connector.PyTree._setInterpData2 (Connector/Connector/Mpi.py:1031-1105): tR / tD hold the
    receptor / donor zones this rank owns.  Previous subregions of tD are dropped, a receptor zone without cellN is
    interpolated as a whole (cellN = 2 for the call), every bbox-intersecting donor zone of tD is a candidate whatever
    its base (the two trees are different trees), and the ID_* subregions end up under the donor zones' owners.  Same
    fixed arguments as the reference: nature=1, penalty=1, extrap=1, storage='inverse', sameName=0, itype='chimera'."""
    varcelln = 'cellN' if loc == 'nodes' else 'centers:cellN'
    for z in Internal.getZones(tD):
        z[2][:] = [c for c in z[2] if c[3] != 'ZoneSubRegion_t' and c[0] != 'GridCoordinates#Init']
    temporary = [z for z in Internal.getZones(tR) if C.isNamePresent(z, varcelln) == -1]
    for z in temporary: C._initVars(z, varcelln, 2.)
    _setInterpData(tR, tD, order=order, penalty=1, nature=1, extrap=1, loc=loc, storage='inverse',
                   interpDataType=0 if cartesian else 1, sameName=0, sameBase=1, itype='chimera', verbose=0, comm=comm, ctx=ctx,
                   backend=backend)
    for z in temporary: C._rmVars(z, [varcelln])
    return None


class _Single:
    rank = 0
    size = 1

    def allgather_obj(self, o):
        return [o]


def _rawF(a):
    """True when the array's memory IS its Fortran-order flattening (the layout of the device copy), so that it can be
    copied raw; a C-contiguous 2-D/3-D array is not (it goes through a.reshape(-1, order='F'))."""
    return a.flags.f_contiguous or a.ndim == 1


def _rcvFieldName(v, loc):
    return ('centers:' + v) if loc == 'centers' else v


def setInterpTransfers(aR, aD, variables=[], cellNVariable='',
                       variablesIBC=['Density', 'MomentumX', 'MomentumY', 'MomentumZ', 'EnergyStagnationDensity'],
                       bcType=0, varType=1, graph=None, procDict=None, type='ALLD',
                       Gamma=1.4, Cv=1.7857142857142865, MuS=1.e-08, Cs=0.3831337844872463, Ts=1.0, alpha=1.,
                       comm=None, backend=None, ctx=None):
    """Copy variant (Connector/Mpi.py:383-394): the receptor tree is copied by reference, the fields are updated."""
    tp = Internal.copyRef(aR)
    _setInterpTransfers(tp, aD, variables=variables, cellNVariable=cellNVariable, variablesIBC=variablesIBC,
                        bcType=bcType, varType=varType, compact=0, graph=graph, procDict=procDict, type=type,
                        Gamma=Gamma, Cv=Cv, MuS=MuS, Cs=Cs, Ts=Ts, alpha=alpha, comm=comm, backend=backend, ctx=ctx)
    return tp


def _setInterpTransfers(aR, aD, variables=[], cellNVariable='',
                        variablesIBC=['Density', 'MomentumX', 'MomentumY', 'MomentumZ', 'EnergyStagnationDensity'],
                        bcType=0, varType=1, compact=0, graph=None, procDict=None, type='ALLD',
                        Gamma=1.4, Cv=1.7857142857142865, MuS=1.e-08, Cs=0.3831337844872463, Ts=1.0, alpha=1.,
                        storage=1, comm=None, backend=None, ctx=None):
    """Distributed setInterpTransfers on host trees (reference signature and semantics,
    Connector/Mpi.py:396-440): donor-side dense transfer per ID_* subregion,
    in-place scatter for local receptor zones, send + scatter for remote ones.
    procDict (zone name -> rank) is used when given, else the owners are gathered; `graph` is not needed (the
    exchange is sized from the subregions); the IBC arguments only concern IBCD_* subregions, which are not on
    the device path."""
    infos = X._setInterpTransfersD(aD, variables=variables, cellNVariable=cellNVariable, varType=varType,
                                   compact=compact, ctx=ctx, backend=backend)
    zr = {z[0]: z for z in Internal.getZones(aR)}
    owners = {}
    if comm is not None and comm.size > 1:
        if procDict is not None:
            owners = {n: int(p) for n, p in procDict.items() if n not in zr}
        else:
            for p, names in enumerate(comm.allgather_obj(sorted(zr.keys()))):
                for n in names: owners[n] = p
    outMeta, outData = {}, {}
    for rcvName, arr, ListRcv, loc in infos:
        names = arr[0].split(',') if arr[0] else []
        if rcvName in zr:
            _scatter(zr[rcvName], names, arr[1], ListRcv, loc)
        elif rcvName in owners:
            p = owners[rcvName]
            outMeta.setdefault(p, []).append(dict(rcv=rcvName, names=names, n=int(ListRcv.size), loc=loc))
            outData.setdefault(p, []).append(numpy.ascontiguousarray(ListRcv, dtype=numpy.int32).view(numpy.uint8))
            outData[p].append(numpy.ascontiguousarray(arr[1], dtype=numpy.float64).reshape(-1).view(numpy.uint8))
    if comm is None or comm.size == 1:
        return None
    allMeta = comm.allgather_obj(outMeta)
    send = {p: numpy.concatenate(parts) for p, parts in outData.items()}
    recv_sizes = {}
    for q, om in enumerate(allMeta):
        if q != comm.rank and comm.rank in om:
            recv_sizes[q] = sum(4 * r['n'] + 8 * r['n'] * len(r['names']) for r in om[comm.rank])
    got = comm.exchange_host(send, recv_sizes, dtype=numpy.uint8)
    for q in sorted(recv_sizes):
        buf = got[q]; o = 0
        for r in allMeta[q][comm.rank]:
            n = r['n']; nv = len(r['names'])
            lst = buf[o:o + 4 * n].view(numpy.int32); o += 4 * n
            vals = buf[o:o + 8 * n * nv].view(numpy.float64).reshape(nv, n); o += 8 * n * nv
            _scatter(zr[r['rcv']], r['names'], vals, lst, r['loc'])
    return None


def _scatter(z, names, vals, ListRcv, loc):
    """C._setPartialFields analogue: fieldR[v][ListRcv] = vals[v]."""
    for v, name in enumerate(names):
        f = C.getFieldArray(z, _rcvFieldName(name, loc))
        if f is None:
            continue
        f.reshape(-1, order='F')[ListRcv] = vals[v]


# ---------------------------------------------------------------------------
# device-resident steady state
# ---------------------------------------------------------------------------
class TransferSession:
    """All ID_* plans of this rank resident on the GPU: donor fields
    (FlowSolution of aD zones) and receptor fields (container at the
    subregion's GridLocation in aR zones) are uploaded once; step() runs one
    fused transfer launch per interpolation type, one grouped NCCL
    send/recv (one packed buffer per peer rank) and one scatter launch."""

    def __init__(self, aR, aD, variables, comm=None, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.comm = comm
        self.variables = list(variables)
        nv = len(self.variables)
        ctx = self.ctx
        rank = comm.rank if comm is not None else 0
        size = comm.size if comm is not None else 1
        self.zr = {z[0]: z for z in Internal.getZones(aR)}
        owners = {}
        if size > 1:
            for p, names in enumerate(comm.allgather_obj(sorted(self.zr.keys()))):
                for n in names: owners[n] = p
        else:
            owners = {n: 0 for n in self.zr}
        # resident fields
        self.dD, self.dR, self.hR = {}, {}, {}
        self._rcvIdx = {}      # receptor zone -> index lists written by the plans (local and remote) that feed it
        self.h2d_bytes = 0
        for zd in Internal.getZones(aD):
            arrs = [C.getFieldArray(zd, v) for v in self.variables]
            if any(a is None for a in arrs):
                continue
            self.dD[zd[0]] = [_lib.DeviceArray.from_host(ctx, a.reshape(-1, order='F')) for a in arrs]
            self.h2d_bytes += sum(a.nbytes for a in arrs)
        self._aD = aD
        self.plans = []
        self._donorPlans = {}  # donor zone -> its plans (local and remote): what the zone has to provide per step
        items = []
        ritems = []
        remote = {}       # peer -> list of (donorName, rcvName, plan)
        self.alg_fixed = 0; self.n_entries = 0; self.distinct = 0
        for zd in Internal.getZones(aD):
            if zd[0] not in self.dD:
                continue
            _, imd, jmd, kmd, _ = Internal.getZoneDim(zd)
            for s in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
                if not s[0].startswith('ID_') or Internal.getNodeFromName1(s, 'InterpolantsDonor') is None:
                    continue
                rname = Internal.getValue(s)
                loc = 'centers' if Internal.getNodeFromName1(s, 'GridLocation') is not None else 'nodes'
                ListDonor = Internal.getNodeFromName1(s, 'PointList')[1]
                ListRcv = Internal.getNodeFromName1(s, 'PointListDonor')[1]
                typ = Internal.getNodeFromName1(s, 'InterpolantsType')[1]
                cf = Internal.getNodeFromName1(s, 'InterpolantsDonor')[1]
                if rname in self.zr:
                    zr = self.zr[rname]
                    nptsR = int(numpy.prod(C.fieldShape(zr, loc)))
                    self._residentR(zr, loc)
                    self._rcvIdx.setdefault(rname, []).append(numpy.asarray(ListRcv))
                    plan = _lib.Plan(ctx, imd, jmd, kmd, nptsR, ListDonor, ListRcv, typ, cf)
                    items.append((plan, 0, [d.ptr for d in self.dD[zd[0]]], [d.ptr for d in self.dR[rname]], None, 0))
                elif rname in owners:
                    nmax = int(ListRcv.max()) + 1 if ListRcv.size else 1
                    plan = _lib.Plan(ctx, imd, jmd, kmd, nmax, ListDonor, ListRcv, typ, cf)
                    remote.setdefault(owners[rname], []).append((zd[0], rname, plan, ListRcv, loc))
                else:
                    continue
                self.plans.append(plan)
                self._donorPlans.setdefault(zd[0], []).append(plan)
                info = plan.info()
                self.alg_fixed += info['alg_fixed']; self.n_entries += info['n']; self.distinct += info['distinct']
        # one block (nvars, n_pair) per remote plan, in (donor, receptor) name order inside each peer's message.
        # The blocks are written in the plan's own donor-sorted entry order (mode 2: coalesced stores, also
        # over NVLink), so the receptor rank gets the index list in that order (once, here).
        self.sendcnt = {}
        outMeta, outIdx = {}, {}
        rlayout = {}      # peer -> list of (donorName, plan, offset in doubles)
        for p, lst in remote.items():
            lst.sort(key=lambda t: (t[0], t[1]))
            self.sendcnt[p] = sum(t[2].n for t in lst) * nv
            o = 0
            for dn, rn, plan, ListRcv, loc in lst:
                rlayout.setdefault(p, []).append((dn, plan, o))
                outMeta.setdefault(p, []).append(dict(donor=dn, rcv=rn, n=int(plan.n), loc=loc))
                outIdx.setdefault(p, []).append(plan.sorted_rcv().view(numpy.uint8))
                o += plan.n * nv
        self.planset = _lib.PlanSet(ctx, nv, items)          # local pairs, in place
        # receive side: index lists are exchanged once
        self.recvcnt = {}
        rlists = {}       # peer -> list of (n, device index list, offset in doubles, receptor zone name)
        self._keep = []
        if size > 1:
            allMeta = comm.allgather_obj(outMeta)
            send = {p: numpy.concatenate(parts) for p, parts in outIdx.items()}
            recv_sizes = {q: sum(4 * r['n'] for r in om[rank]) for q, om in enumerate(allMeta)
                          if q != rank and rank in om}
            got = comm.exchange_host(send, recv_sizes, dtype=numpy.uint8)
            for q in sorted(recv_sizes):
                recs = allMeta[q][rank]
                self.recvcnt[q] = sum(r['n'] for r in recs) * nv
                o = 0; ob = 0
                for r in recs:
                    n = r['n']
                    lst = got[q][ob:ob + 4 * n].view(numpy.int32).copy(); ob += 4 * n
                    self._residentR(self.zr[r['rcv']], r['loc'])
                    self._rcvIdx.setdefault(r['rcv'], []).append(lst)
                    dl = _lib.DeviceArray.from_host(ctx, lst)
                    self._keep.append(dl)
                    rlists.setdefault(q, []).append((n, dl, o, r['rcv']))
                    o += n * nv
        self.peers = sorted(set(self.sendcnt) | set(self.recvcnt))
        self.transport = 'none'
        if size > 1:
            # collective decision: every rank goes through the same sequence of exchanges, also without peers
            want = os.environ.get("CX_TRANSPORT", "p2p")
            ok = self._setup_p2p(rlayout, rlists, nv) if want == 'p2p' else False
            if ok:
                self.transport = 'p2p'
            else:
                self._setup_nccl(rlayout, rlists, nv)
                self.transport = 'nccl'
        else:
            self.planset_remote = None
            self.scatterset = _lib.ScatterSet(ctx, nv, [])
            self._step = _lib.Step(ctx, None, self.planset, self.scatterset, [], [], [], [], [])
        self._setup_sparse_download(nv)
        ctx.sync()

    def _setup_sparse_download(self, nv):
        """Receptor zones of which a step writes less than half the points are downloaded sparsely: the written values
        are packed on the device (cx_gather_dev), copied into a page-locked block and stored into the host arrays by
        index (cx_host_scatter) -- what Connector/Mpi.py:396-440 does with C._setPartialFields.  CX_SPARSE_D2H=0: always
        whole fields."""
        self._sparse = {}
        if os.environ.get("CX_SPARSE_D2H", "1") == "0":
            return
        for name, lists in self._rcvIdx.items():
            arrs = self.hR.get(name)
            if not arrs or not all(_rawF(a) for a in arrs):
                continue
            idx = numpy.unique(numpy.concatenate(lists)).astype(numpy.int32)
            if idx.size == 0 or 2 * idx.size >= arrs[0].size:
                continue
            self._sparse[name] = dict(idx=idx, didx=_lib.DeviceArray.from_host(self.ctx, idx),
                                      dbuf=_lib.DeviceArray(self.ctx, nv * idx.size),
                                      hbuf=_lib.pinned_empty(nv * idx.size))

    def _setup_nccl(self, rlayout, rlists, nv):
        """Remote plans -> per-peer send buffers -> grouped ncclSend/ncclRecv -> scatter."""
        ctx = self.ctx
        self.sendbuf = {p: _lib.DeviceArray(ctx, max(c, 1)) for p, c in self.sendcnt.items()}
        self.recvbuf = {q: _lib.DeviceArray(ctx, max(c, 1)) for q, c in self.recvcnt.items()}
        ritems = [(plan, 2, [d.ptr for d in self.dD[dn]], None, self.sendbuf[p].ptr + 8 * o, plan.n)
                  for p, lst in rlayout.items() for dn, plan, o in lst]
        sitems = [(n, dl, self.recvbuf[q].ptr + 8 * o, n, [d.ptr for d in self.dR[rn]])
                  for q, lst in rlists.items() for n, dl, o, rn in lst]
        self.planset_remote = _lib.PlanSet(ctx, nv, ritems) if ritems else None
        self.scatterset = _lib.ScatterSet(ctx, nv, sitems)
        sp = [self.sendbuf[p].ptr if p in self.sendbuf else 0 for p in self.peers]
        sc = [self.sendcnt.get(p, 0) for p in self.peers]
        rp = [self.recvbuf[p].ptr if p in self.recvbuf else 0 for p in self.peers]
        rc = [self.recvcnt.get(p, 0) for p in self.peers]
        self._step = _lib.Step(ctx, self.planset_remote, self.planset, self.scatterset, self.peers, sp, sc, rp, rc)

    def _setup_p2p(self, rlayout, rlists, nv):
        """Peer-memory transport: this rank's receive buffers (two per peer, by iteration parity) and its flag
        array are exported with CUDA IPC; the remote plans of the donor ranks store into them directly."""
        ctx, comm = self.ctx, self.comm
        rank = comm.rank
        try:
            self.recvbuf2 = {q: [_lib.DeviceArray(ctx, max(self.recvcnt.get(q, 0), 1)) for _ in (0, 1)] for q in self.peers}
            self.flags = _lib.DeviceArray(ctx, max(len(self.peers), 1), numpy.int32)
            ctx.upload(self.flags.ptr, numpy.zeros(max(len(self.peers), 1), numpy.int32))
            mine = {q: (_lib.ipc_export(self.recvbuf2[q][0].ptr), _lib.ipc_export(self.recvbuf2[q][1].ptr),
                        _lib.ipc_export(self.flags.ptr + 4 * k)) for k, q in enumerate(self.peers)}
            err = None
        except _lib.CxError as e:
            mine, err = None, str(e)
        allx = comm.allgather_obj(mine)
        if any(x is None for x in allx):
            if err and os.environ.get("CX_PROFILE"):
                print("[TransferSession] peer-memory transport unavailable (%s); using NCCL" % err, file=_sys.stderr, flush=True)
            return False
        self._ipc_descs = []                    # one entry per successful ipc_open (the mappings are reference counted)

        def _open(d):
            q = _lib.ipc_open(ctx, d)
            self._ipc_descs.append(d)
            return q
        try:
            rbuf = {p: [_open(allx[p][rank][par]) for par in (0, 1)] for p in self.peers}
            pflags = [_open(allx[p][rank][2]) for p in self.peers]
        except _lib.CxError as e:
            if os.environ.get("CX_PROFILE"):
                print("[TransferSession] cudaIpcOpenMemHandle failed (%s); using NCCL" % e, file=_sys.stderr, flush=True)
            rbuf = None
        if not all(comm.allgather_obj(rbuf is not None)):
            for d in self._ipc_descs:             # NCCL fallback: drop what this rank did map
                _lib.ipc_close(d)
            self._ipc_descs = []
            comm.barrier()
            return False
        ps2, ss2 = [], []
        for par in (0, 1):
            ritems = [(plan, 2, [d.ptr for d in self.dD[dn]], None, rbuf[p][par] + 8 * o, plan.n)
                      for p, lst in rlayout.items() for dn, plan, o in lst]
            sitems = [(n, dl, self.recvbuf2[q][par].ptr + 8 * o, n, [d.ptr for d in self.dR[rn]])
                      for q, lst in rlists.items() for n, dl, o, rn in lst]
            ps2.append(_lib.PlanSet(ctx, nv, ritems) if ritems else None)
            ss2.append(_lib.ScatterSet(ctx, nv, sitems) if sitems else None)
        self.planset_remote = ps2[0]
        self._ps2, self._ss2 = ps2, ss2
        self._step = _lib.StepP2P(ctx, ps2, self.planset, ss2, pflags, self.flags)
        ctx.sync()
        comm.barrier()      # every rank's flags are zeroed and mapped before anyone signals
        return True

    def close(self):
        """Collective: unmap the peers' buffers before anybody frees them (peer-memory transport), then drop the
        device objects of this session.  Call it before deleting the session when another one will follow."""
        if getattr(self, 'transport', 'none') == 'p2p':
            self.ctx.sync()
            self.comm.barrier()                    # nobody is still storing into a peer
            self._step = None; self._ps2 = None; self.planset_remote = None; self._hostio = None
            for d in getattr(self, '_ipc_descs', []):
                _lib.ipc_close(d)
            self._ipc_descs = []
            self.comm.barrier()                    # every mapping is gone: the exporters may free
            self.recvbuf2 = None; self.flags = None; self._ss2 = None
        self.transport = 'closed'

    def _residentR(self, zr, loc):
        if zr[0] in self.dR:
            return
        arrs = [C.getFieldArray(zr, _rcvFieldName(v, loc)) for v in self.variables]
        self.hR[zr[0]] = arrs
        self.dR[zr[0]] = [_lib.DeviceArray.from_host(self.ctx, a.reshape(-1, order='F')) for a in arrs]

    def alg_bytes(self):
        """B_alg of SURVEY.md 8(d) summed over this rank's plans."""
        nv = len(self.variables)
        return self.alg_fixed + 8 * nv * (self.n_entries + self.distinct)

    def nccl_bytes(self):
        """Bytes this rank ships to other ranks per iteration (NCCL messages or peer-memory stores)."""
        return 8 * sum(self.sendcnt.values())

    def pin_host_fields(self):
        """Page-lock the tree's donor and receptor field arrays in place, so that the per-iteration
        upload / download run as asynchronous copies at bus speed (idempotent)."""
        if getattr(self, '_pinned', None) is not None:
            return
        self._pinned = []
        seen = set()
        arrs = [a for lst in self.hR.values() for a in lst]
        for zd in Internal.getZones(self._aD):
            if zd[0] in self.dD:
                arrs += [C.getFieldArray(zd, v) for v in self.variables]
        for a in arrs:
            if id(a) in seen or not _rawF(a):
                continue
            seen.add(id(a))
            try:
                if _lib.host_register(a):
                    self._pinned.append(a)
            except _lib.CxError:
                pass

    def unpin_host_fields(self):
        for a in getattr(self, '_pinned', None) or []:
            _lib.host_unregister(a)
        self._pinned = None

    def upload_donor_fields(self):
        """Host tree -> device, asynchronous on the context's stream (the next step() is ordered after it)."""
        for zd in Internal.getZones(self._aD):
            if zd[0] in self.dD and zd[0] in self._donorPlans:      # a zone nobody interpolates from stays where it is
                for d, v in zip(self.dD[zd[0]], self.variables):
                    a = C.getFieldArray(zd, v)
                    if _rawF(a):
                        self.ctx.upload_async(d.ptr, a)
                    else:
                        d.upload(numpy.ascontiguousarray(a.reshape(-1, order='F')))

    def upload_receptor_fields(self):
        """Host tree -> device for the receptor fields (the device copies are otherwise only written by step())."""
        for name, arrs in self.hR.items():
            for d, a in zip(self.dR[name], arrs):
                d.upload(numpy.ascontiguousarray(a.reshape(-1, order='F')))

    def step(self, niter=1, overlap=True):
        """niter iterations (asynchronous): remote plans -> NCCL exchange + scatter on the
        communication stream, overlapped with the local in-place plans."""
        self._step.run(niter, overlap)

    def _setup_hostio(self):
        """The copies of transfer_host(): per donor zone the points its plans read (whole fields when that is more than
        half the zone), per receptor zone the points the step writes (the sparse-download lists)."""
        nv = len(self.variables)
        io = _lib.HostIO(self.ctx, nv)
        for zd in Internal.getZones(self._aD):
            plans = self._donorPlans.get(zd[0])
            if not plans or zd[0] not in self.dD:
                continue
            arrs = [C.getFieldArray(zd, v) for v in self.variables]
            if not all(_rawF(a) for a in arrs):
                return None
            npts = arrs[0].size
            idx = None; box = None
            if os.environ.get("CX_SPARSE_H2D", "1") != "0":
                mask = _lib.plan_donor_mask(self.ctx, plans, npts)
                idx = numpy.flatnonzero(mask).astype(numpy.int32)
                if 4 * idx.size >= npts:      # a host gather over most of the array is slower than the plain copy:
                    # the bounding box of the points read goes up as one strided copy per variable instead
                    if idx.size and arrs[0].ndim == 3 and all(a.flags.f_contiguous for a in arrs):
                        m3 = mask.reshape(arrs[0].shape, order='F').astype(bool)
                        ii = numpy.flatnonzero(m3.any(axis=(1, 2))); jj = numpy.flatnonzero(m3.any(axis=(0, 2)))
                        kk = numpy.flatnonzero(m3.any(axis=(0, 1)))
                        b = (int(ii[0]), int(ii[-1]), int(jj[0]), int(jj[-1]), int(kk[0]), int(kk[-1]))
                        if (b[1] - b[0] + 1) * (b[3] - b[2] + 1) * (b[5] - b[4] + 1) <= 0.85 * npts:
                            box = b
                    idx = None
            if box is not None:
                io.add_box('up', arrs, [d.ptr for d in self.dD[zd[0]]], box)
            else:
                io.add('up', arrs, [d.ptr for d in self.dD[zd[0]]], idx)
        for name, arrs in self.hR.items():
            if not all(_rawF(a) for a in arrs):
                return None
            sp = self._sparse.get(name)
            io.add('down', arrs, [d.ptr for d in self.dR[name]], sp['idx'] if sp is not None else None)
        return io

    def transfer_host(self, pipelined=True):
        """One iteration, host trees in and host trees out: the donor values the plans read go up, the plans and the
        exchange run, the receptor values they wrote come back into the tree's arrays (what a call of the reference's
        Xmpi.__setInterpTransfers + C._setPartialFields does).  Pipelined per variable over three streams (peer-memory
        transport or a single rank), so uploads, kernels and downloads overlap.  Collective: every rank calls it.
        Returns (bytes up, bytes down)."""
        if getattr(self, '_hostio', None) is None:
            self._hostio = self._setup_hostio() or False
        if self._hostio is False:           # some field is not a Fortran-flat float64 array: plain sequence
            self.upload_donor_fields()
            if pipelined and self._step.splittable():
                # the same sequence of sub-iterations as the ranks that pipeline (the flag protocol counts them)
                for v in range(len(self.variables)):
                    self._step.run_vars(v, 1)
            else:
                self.step()
            nb = self.download()
            return self.h2d_bytes, nb
        self._step.run_host(self._hostio, pipelined)
        return self._hostio.bytes()

    def download(self):
        """Copy the receptor fields back into the host tree (in place); returns the bytes moved."""
        nb = 0
        slow = []
        for name, arrs in self.hR.items():
            sp = self._sparse.get(name)
            if sp is not None:
                n = sp['idx'].size
                _lib.gather_dev(self.ctx, sp['dbuf'].ptr, n, n, sp['didx'].ptr, [d.ptr for d in self.dR[name]])
                self.ctx.download_async(sp['hbuf'], sp['dbuf'].ptr)
                nb += sp['hbuf'].nbytes
                continue
            for d, a in zip(self.dR[name], arrs):
                if _rawF(a):
                    self.ctx.download_async(a, d.ptr)
                else:
                    slow.append((d, a))
                nb += a.nbytes
        self.ctx.sync()
        for name, sp in self._sparse.items():
            n = sp['idx'].size
            for v, a in enumerate(self.hR[name]):
                _lib.host_scatter(a, sp['idx'], sp['hbuf'][v * n:(v + 1) * n])
        for d, a in slow:
            a[...] = d.download().reshape(a.shape, order='F')
        return nb
