"""Compact solver transfers: the flattened donor tree (`Parameter_int` / `Parameter_real` under the tc root) and the
transfer the Fast solver calls on it every sub-iteration.

Reference: Connector/Connector/compactTransfers.py:16 (`miseAPlatDonorTree__`, the packer) and
Connector/Connector/setInterpTransfers.cpp:761-1410 (`connector.___setInterpTransfers`, the reader loop; per-type
arithmetic in commonInterpTransfers_reorder_5eq.h / _6eq.h), called from Connector/Connector/Mpi.py:496-530
(`__setInterpTransfers`).

Scope here (everything else raises): chimera `ID_*` racs of structured zones, cell-centred receptors (the reader
refuses vertices itself, setInterpTransfers.cpp:1104-1109), steady racs (no `#` time levels), no IBC regions, one
process (graph=None: every rac goes to "proc 0", NbP2P = 1), 5 or 6 equations, implicit / global explicit time
stepping (exploc = 0).  The `param_int` / `param_real` layout is the reference's, index for index, so a tree packed
here is read by the reference's reader and vice versa (oracle/ref_driver.cpp:ref_transfer_compact is that reader
with the reference's own arithmetic fragments #included; tests/ compare against it).

The arithmetic runs in the CUDA plan kernels (csrc/cx_transfer.cu) through one transfer plan per rac, built once per
`param_int` and cached.  Differences of the compact reader with `_setInterpTransfers` that are reproduced:
type 1 is a pure injection (the stored coefficient is ignored, commonInterpTransfers_reorder_5eq.h:33-45); sums
start from the first product instead of 0 (only the sign of an all-zero result can differ: -0.0 vs +0.0)."""
import ctypes
import os
import weakref

import numpy

from .. import Internal
from .. import _lib

# indices into a zone's .Solver#ownData/Parameter_int (Connector/Connector/param_solver.h:78-119)
NIJK = 0
IFLOW = 27
NDIMDX = 41
LEVEL = 53

NTAB_INT = 18          # per-rac tables of param_int (compactTransfers.py:70)


def _sizecf(t):
    """SIZECF, structured donors (Connector/Connector/connector.h:27-37)."""
    try:
        return {1: 1, 2: 8, 22: 4, 3: 9, 44: 12, 4: 8, 5: 15}[int(t)]
    except KeyError:
        raise NotImplementedError("compact transfers: interpolation type %d has no device path." % int(t))


def miseAPlatDonorTree__(t, tc, graph=None, list_graph=None, nbpts_linelets=0):
    """Flatten the interpolation data of tc into `Parameter_int` / `Parameter_real` (created under the tc root) in the
    layout of compactTransfers.py:16-1242 and re-point every subregion's PointList / PointListDonor /
    InterpolantsType / InterpolantsDonor at its slice of those arrays, entries regrouped by interpolation type
    (`triMultiType`, :1244-: types in order of first appearance, stable inside a type).  Zone numbers (NoD, NoR) are
    positions in Internal.getZones(t), which must list the same zones as tc in the same order."""
    if graph is not None or list_graph is not None:
        raise NotImplementedError("miseAPlatDonorTree__: one process only (graph=None).")
    zones = Internal.getZones(t)
    zones_tc = Internal.getZones(tc)
    if [z[0] for z in zones] != [z[0] for z in zones_tc]:
        raise ValueError("miseAPlatDonorTree__: t and tc must hold the same zones in the same order.")
    numero = {z[0]: c for c, z in enumerate(zones)}
    neq_base = {}
    for b in Internal.getBases(tc) or [tc]:
        g = Internal.getNodeFromName2(b, 'GoverningEquations')
        model = 'NSLaminar' if g is None else Internal.getValue(g)
        if model not in ('NSLaminar', 'Euler', 'NSTurbulent'):
            raise NotImplementedError("miseAPlatDonorTree__: GoverningEquations %s has no device path." % model)
        for z in Internal.getZones(b):
            gz = Internal.getNodeFromName2(z, 'GoverningEquations')
            mz = model if gz is None else Internal.getValue(gz)
            neq_base[z[0]] = 6 if mz == 'NSTurbulent' else 5
    racs = []                     # (subregion, NoD, neq)
    for c, z in enumerate(zones_tc):
        zt = Internal.getNodeFromType1(z, 'ZoneType_t')
        if zt is not None and Internal.getValue(zt) != 'Structured':
            raise NotImplementedError("miseAPlatDonorTree__: unstructured donor zone %s." % z[0])
        for s in Internal.getNodesFromType1(z, 'ZoneSubRegion_t'):
            if '#' in s[0]:
                raise NotImplementedError("miseAPlatDonorTree__: unsteady racs (%s) have no device path." % s[0])
            if s[0][0:2] == 'IB':
                raise NotImplementedError("miseAPlatDonorTree__: IBC regions (%s) have no device path." % s[0])
            if Internal.getNodeFromName1(s, 'InterpolantsDonor') is None: continue
            if Internal.getNodeFromName1(s, 'RotationAngle') is not None:
                raise NotImplementedError("miseAPlatDonorTree__: periodicity by rotation (%s)." % s[0])
            racs.append((s, c, neq_base[z[0]]))
    nrac = len(racs)
    # sizes (compactTransfers.py:297-311): one destination, no time level
    sizeNbD = sum(Internal.getNodeFromName1(s, 'PointListDonor')[1].size for s, _, _ in racs)
    sizeNb = sum(Internal.getNodeFromName1(s, 'PointList')[1].size for s, _, _ in racs)
    ntyp = [numpy.unique(Internal.getNodeFromName1(s, 'InterpolantsType')[1]).size for s, _, _ in racs]
    sizeType = sizeNbD + sum(n + 1 for n in ntyp)
    sizeI = sizeNbD + sizeNb + sizeType
    sizeR = sum(Internal.getNodeFromName1(s, 'InterpolantsDonor')[1].size for s, _, _ in racs)
    NbP2P = 1 if nrac else 0
    sizeproc = 4 + NTAB_INT * nrac + sizeI
    size_int = 2 + NbP2P + (sizeproc if nrac else 0)
    param_int = numpy.zeros(size_int + 2, dtype=Internal.E_NpyInt)
    param_real = numpy.zeros(max(sizeR, 1), dtype=numpy.float64)
    Internal.createUniqueChild(tc, 'Parameter_int', 'DataArray_t', param_int)
    Internal.createUniqueChild(tc, 'Parameter_real', 'DataArray_t', param_real)
    param_int[0] = 0
    param_int[1] = NbP2P
    param_int[2] = 0              # no IBC graph
    param_int[3] = 0              # no ID graph (one process)
    if nrac == 0:
        return 1
    shift_graph = 4
    pt_ech = NbP2P + shift_graph
    param_int[shift_graph] = pt_ech
    param_int[pt_ech:pt_ech + 4] = (0, nrac, 0, 0)      # destination, racs, unsteady racs, time levels
    iadr2 = pt_ech + 4
    iadr_ptr = iadr2 + NTAB_INT * nrac
    o_lstD = o_typ = o_lst = o_cf = 0
    npass = 1
    for k, (s, NoD, neq) in enumerate(racs):
        nodes = {n: Internal.getNodeFromName1(s, n) for n in ('PointList', 'PointListDonor', 'InterpolantsType',
                                                              'InterpolantsDonor')}
        pl, pld, ty, cf = [nodes[n][1] for n in ('PointList', 'PointListDonor', 'InterpolantsType', 'InterpolantsDonor')]
        n = pld.size
        if pl.size != n:
            raise NotImplementedError("miseAPlatDonorTree__: type-0 (point cloud) entries in %s." % s[0])
        zRname = Internal.getValue(s)
        if zRname not in numero:
            raise ValueError("miseAPlatDonorTree__: receptor zone %s of %s is not in t." % (zRname, s[0]))
        # types in order of first appearance; stable regrouping (triMultiType)
        first = {}
        for v in ty.tolist():
            if v not in first: first[v] = len(first)
        order_types = sorted(first, key=first.get)
        if any(v in (44, 5) for v in order_types): npass = 2
        szc = numpy.array([_sizecf(v) for v in ty.tolist()], dtype=numpy.int64)
        if int(szc.sum()) != cf.size:
            raise ValueError("miseAPlatDonorTree__: %s: coefficient count does not match the types." % s[0])
        cstart = numpy.concatenate(([0], numpy.cumsum(szc)[:-1]))
        rank = numpy.array([first[v] for v in ty.tolist()], dtype=numpy.int64)
        perm = numpy.argsort(rank, kind='stable')
        lstD = iadr_ptr + o_lstD
        ptTy = iadr_ptr + sizeNbD + o_typ
        lst = iadr_ptr + sizeNbD + sizeType + o_lst
        ntl = len(order_types)
        param_int[ptTy] = ntl
        for q, v in enumerate(order_types):
            param_int[ptTy + 1 + q] = int((ty == v).sum())
        shift_typ = 1 + ntl
        param_int[ptTy + shift_typ: ptTy + shift_typ + n] = ty[perm]
        param_int[lst: lst + n] = pl[perm]
        param_int[lstD: lstD + n] = pld[perm]
        idx = numpy.concatenate([numpy.arange(cstart[e], cstart[e] + szc[e]) for e in perm]) if n else numpy.zeros(0, numpy.int64)
        param_real[o_cf: o_cf + cf.size] = cf[idx]
        iadr = iadr2 + k
        loc = Internal.getNodeFromName1(s, 'GridLocation')
        tab = {0: n, 1: cf.size, 2: ntl, 3: -1, 4: 0, 5: NoD, 6: lst, 7: ptTy, 8: o_cf,
               9: 1 if (loc is not None and Internal.getValue(loc) == 'CellCenter') else 0, 10: n, 11: numero[zRname],
               12: lstD, 13: min(5, neq) if Internal.getNodeFromName1(s, 'RANSLES') is not None else neq, 14: 0, 15: 1,
               16: 0, 17: iadr_ptr + sizeNbD + sizeType + sizeNb}
        for col, val in tab.items():
            param_int[iadr + nrac * col] = val
        # the subregion now lives in the flattened arrays
        nodes['PointList'][1] = param_int[lst: lst + n]
        nodes['InterpolantsType'][1] = param_int[ptTy + shift_typ: ptTy + shift_typ + n]
        nodes['PointListDonor'][1] = param_int[lstD: lstD + n]
        nodes['InterpolantsDonor'][1] = param_real[o_cf: o_cf + cf.size]
        o_lstD += n; o_lst += n; o_typ += n + ntl + 1; o_cf += cf.size
    return npass


# ---------------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------------
_COMPACT_CACHE = {}


def _racTable(param_int, no_transfert):
    """Header of one exchange of param_int (setInterpTransfers.cpp:908-918)."""
    nbcomIBC = int(param_int[2])
    nbcomID = int(param_int[3 + nbcomIBC])
    shift_graph = nbcomIBC + nbcomID + 3
    ech = int(param_int[no_transfert + shift_graph])
    nrac, nrac_inst, timelevel = int(param_int[ech + 1]), int(param_int[ech + 2]), int(param_int[ech + 3])
    return ech, nrac, nrac_inst, timelevel


def _impliLocal(nzones, zparams, dtloc, nstep, threadmax_sdm):
    """Zones modified at sub-iteration nstep (setInterpTransfers.cpp:930-966)."""
    out = [0] * nzones
    nssiter = int(dtloc[0]); shift_omp = int(dtloc[11])
    omp = dtloc[shift_omp:]
    nbtask = int(omp[nstep - 1]); ptiter = int(omp[nssiter + nstep - 1])
    for ntask in range(nbtask):
        out[int(omp[ptiter + ntask * (6 + threadmax_sdm * 7)])] = 1
    maxlevel = int(dtloc[9]); it_cycl_lbm = int(dtloc[10])
    max_it = int(2 ** (maxlevel - 1)) if maxlevel >= 1 else 0
    level_next_it = maxlevel
    if it_cycl_lbm != max_it - 1: level_next_it = int(dtloc[13 + it_cycl_lbm + 1])
    for nd in range(nzones):
        if int(zparams[nd][IFLOW]) == 4:
            out[nd] = 1 if int(zparams[nd][LEVEL]) <= level_next_it + 1 else 0
    return out


def _hostView(first, count):
    """count doubles starting at the first element of numpy array `first` (the compact solution block)."""
    addr = first.__array_interface__['data'][0]
    buf = (ctypes.c_double * count).from_address(addr)
    return numpy.ctypeslib.as_array(buf)


def _numThreads():
    """__NUMTHREADS__ of the reader (the stride of dtloc's task records): OMP_NUM_THREADS like an OpenMP runtime
    reads it, else the host's core count."""
    v = os.environ.get("OMP_NUM_THREADS", "")
    try:
        return max(1, int(v.split(',')[0]))
    except ValueError:
        return os.cpu_count() or 1


def ___setInterpTransfers(zones, zonesD, vars, dtloc, param_int, param_real, it_target, varType, type_transfert,
                          no_transfert, nstep, nitmax, rk, exploc, num_passage, isWireModel=0, ctx=None,
                          threadmax_sdm=None):
    """connector.___setInterpTransfers (setInterpTransfers.cpp:761-1410): in-place transfers on compacted zones
    driven by the flattened tc.  zones / zonesD: receptor (t) and donor (tc) zones, index-aligned; vars[0] names the
    first variable of the compact solution (centres of t, nodes of tc); every zone of `zones` carries
    .Solver#ownData/Parameter_int.  Only the racs whose donor zone was modified at sub-iteration nstep (dtloc) are
    transferred, like the reference."""
    if exploc == 1:
        raise NotImplementedError("___setInterpTransfers: local explicit time stepping (exploc=1) has no device path.")
    if type_transfert == 1:
        return None                       # IBC pass only: nothing on this path
    if isWireModel not in (0, None):
        raise NotImplementedError("___setInterpTransfers: wire model transfers have no device path.")
    if 1 <= varType <= 3: nvars = 5
    elif varType in (11, 21, 31): nvars = 6
    else:
        raise NotImplementedError("___setInterpTransfers: varType %d (LBM / hybrid transfers) has no device path." % varType)
    ctx = ctx or _lib.default_context()
    ech, nrac, nrac_inst, timelevel = _racTable(param_int, no_transfert)
    if nrac_inst > 0:
        raise NotImplementedError("___setInterpTransfers: unsteady racs have no device path.")
    zparams = []
    for z in zones:
        own = Internal.getNodeFromName1(z, '.Solver#ownData')
        pi = None if own is None else Internal.getNodeFromName1(own, 'Parameter_int')
        if pi is None:
            raise TypeError("___setInterpTransfers: zone %s has no .Solver#ownData/Parameter_int." % z[0])
        zparams.append(pi[1])
    if threadmax_sdm is None: threadmax_sdm = _numThreads()
    impli = _impliLocal(len(zones), zparams, dtloc, nstep, threadmax_sdm)
    varname = vars[0]

    def first(z, container):
        sol = Internal.getNodeFromName1(z, container)
        node = None if sol is None else Internal.getNodeFromName1(sol, varname)
        if node is None:
            raise TypeError("___setInterpTransfers: %s not found in %s of zone %s." % (varname, container, z[0]))
        return node[1]

    plans = _compactPlans(ctx, param_int, param_real, ech, nrac, timelevel, zparams)
    ctx.batch_begin()
    try:
        lastD = None
        for irac in range(nrac):
            shift_rac = ech + 4 + timelevel * 2 + irac
            NoD = int(param_int[shift_rac + nrac * 5])
            if impli[NoD] != 1: continue
            if int(param_int[shift_rac + nrac * 3]) >= 0:
                raise NotImplementedError("___setInterpTransfers: IBC racs have no device path.")
            if int(param_int[shift_rac + nrac * 9]) == 0:
                raise NotImplementedError("___setInterpTransfers: vertex-located racs (the reference refuses them too, "
                                          "setInterpTransfers.cpp:1104-1109).")
            if int(param_int[shift_rac + nrac * 14]) != 0:
                raise NotImplementedError("___setInterpTransfers: periodicity by rotation has no device path.")
            NoR = int(param_int[shift_rac + nrac * 11])
            nvars_loc = int(param_int[shift_rac + nrac * 13])
            if nvars_loc not in (5, 6) or nvars_loc > nvars:
                raise NotImplementedError("___setInterpTransfers: %d equations on a rac (LBM / hybrid)." % nvars_loc)
            ndD = int(zparams[NoD][NDIMDX]); ndR = int(zparams[NoR][NDIMDX])
            if lastD is not None and lastD != NoD: ctx.batch_flush()
            lastD = NoD
            bD = _hostView(first(zonesD[NoD], Internal.__FlowSolutionNodes__), nvars_loc * ndD)
            bR = _hostView(first(zones[NoR], Internal.__FlowSolutionCenters__), nvars_loc * ndR)
            fD = [bD[eq * ndD:(eq + 1) * ndD] for eq in range(nvars_loc)]
            fR = [bR[eq * ndR:(eq + 1) * ndR] for eq in range(nvars_loc)]
            plans[irac].transfer(fD, fR)
    finally:
        ctx.batch_end()
    return None


def _compactPlans(ctx, param_int, param_real, ech, nrac, timelevel, zparams):
    """One device transfer plan per rac of an exchange, cached on the identity of param_int / param_real."""
    key = (ctx.device, id(param_int), id(param_real), ech)
    plans = _COMPACT_CACHE.get(key)
    if plans is not None:
        return plans
    plans = []
    for irac in range(nrac):
        shift_rac = ech + 4 + timelevel * 2 + irac
        n = int(param_int[shift_rac + nrac * 10])
        NoD = int(param_int[shift_rac + nrac * 5]); NoR = int(param_int[shift_rac + nrac * 11])
        pos = int(param_int[shift_rac + nrac * 7])
        ntl = int(param_int[pos])
        counts = [int(v) for v in param_int[pos + 1: pos + 1 + ntl]]
        types = numpy.ascontiguousarray(param_int[pos + 1 + ntl: pos + 1 + ntl + n], dtype=numpy.int32)
        p6 = int(param_int[shift_rac + nrac * 6]); p12 = int(param_int[shift_rac + nrac * 12])
        donorPts = numpy.ascontiguousarray(param_int[p6: p6 + n], dtype=numpy.int32)
        rcvPts = numpy.ascontiguousarray(param_int[p12: p12 + n], dtype=numpy.int32)
        # the reader walks whole type ranges (ntype[1..]); inside a range every entry has the range's type
        tt = numpy.repeat(types[numpy.concatenate(([0], numpy.cumsum(counts)[:-1])).astype(int)], counts) if n else types
        if sum(counts) != n or not numpy.array_equal(tt, types):
            raise ValueError("___setInterpTransfers: rac %d: entries are not regrouped by type." % irac)
        if (types == 0).any() or (types == 4).any():
            raise NotImplementedError("___setInterpTransfers: point-cloud / tetra donors have no device path.")
        szc = numpy.array([_sizecf(v) for v in types.tolist()], dtype=numpy.int64)
        p8 = int(param_int[shift_rac + nrac * 8])
        coefs = numpy.array(param_real[p8: p8 + int(szc.sum())], dtype=numpy.float64, copy=True)
        # type 1 = pure injection: the stored weight is not applied (commonInterpTransfers_reorder_5eq.h:33-45)
        if (types == 1).any():
            cstart = numpy.concatenate(([0], numpy.cumsum(szc)[:-1]))
            coefs[cstart[types == 1]] = 1.
        imd = int(zparams[NoD][NIJK]); jmd = int(zparams[NoD][NIJK + 1]); ndD = int(zparams[NoD][NDIMDX])
        kmd = ndD // (imd * jmd)
        plans.append(_lib.Plan(ctx, imd, jmd, kmd, int(zparams[NoR][NDIMDX]), donorPts, rcvPts, types, coefs))
    _COMPACT_CACHE[key] = plans
    for a in (param_int, param_real):
        try:
            weakref.finalize(a, _COMPACT_CACHE.pop, key, None)
        except TypeError:
            pass
    return plans


def __setInterpTransfers(zones, zonesD, vars, dtloc, param_int, param_real, type_transfert, it_target,
                         nstep, nitmax, rk, exploc, num_passage, varType=1, compact=1,
                         graph=None, procDict=None, isWireModel_int=0, ctx=None):
    """Connector.Mpi.__setInterpTransfers (Connector/Connector/Mpi.py:496-530), intra-process part: every exchange of
    param_int whose destination is this process goes through ___setInterpTransfers."""
    nbcomIBC = int(param_int[2])
    shift_graph = nbcomIBC + int(param_int[3 + nbcomIBC]) + 3
    for comm_P2P in range(1, int(param_int[1]) + 1):
        pt_ech = int(param_int[comm_P2P + shift_graph])
        if int(param_int[pt_ech]) != 0 or graph is not None:
            raise NotImplementedError("__setInterpTransfers: exchanges with other processes have no device path here; "
                                      "use Connector.Mpi.TransferSession.")
        ___setInterpTransfers(zones, zonesD, vars, dtloc, param_int, param_real, it_target, varType, type_transfert,
                              comm_P2P, nstep, nitmax, rk, exploc, num_passage, max(isWireModel_int, 0), ctx=ctx)
    return None


def compactFields(z, variables, container, nghost=0):
    """Stand-in for the Fast solver's own compaction (FastC `_compact`): re-allocate the listed variables of one
    container of zone z behind one another in a single block (variable eq at first + eq*ndimdx) and write NIJK /
    NDIMDX into .Solver#ownData/Parameter_int (created with 128 entries, IFLOW = 1, LEVEL = 1, if absent)."""
    sol = Internal.getNodeFromName1(z, container)
    if sol is None:
        raise ValueError("compactFields: zone %s has no %s." % (z[0], container))
    nodes = [Internal.getNodeFromName1(sol, v) for v in variables]
    if any(n is None for n in nodes):
        raise ValueError("compactFields: variables missing in %s of %s." % (container, z[0]))
    shp = nodes[0][1].shape
    nd = int(numpy.prod(shp))
    block = numpy.empty(nd * len(nodes), dtype=numpy.float64)
    for eq, n in enumerate(nodes):
        view = block[eq * nd:(eq + 1) * nd].reshape(shp, order='F')
        view[...] = n[1]
        n[1] = view
    own = Internal.getNodeFromName1(z, '.Solver#ownData')
    if own is None:
        own = Internal.createChild(z, '.Solver#ownData', 'UserDefinedData_t')
    pi = Internal.getNodeFromName1(own, 'Parameter_int')
    if pi is None:
        arr = numpy.zeros(128, dtype=Internal.E_NpyInt)
        arr[IFLOW] = 1; arr[LEVEL] = 1
        own[2].append(['Parameter_int', arr, [], 'DataArray_t'])
        pi = own[2][-1]
    dims = list(shp) + [1] * (3 - len(shp))
    pi[1][NIJK:NIJK + 3] = dims
    pi[1][NDIMDX] = nd
    return block
