"""Connector.PyTree chimera path on the device: setInterpData / setInterpTransfers.

Host-side mirror (same names, arguments, storage layout, error behaviour) of
/root/reference/Cassiopee/Connector/Connector/OversetData.py:
  setInterpData :1133, _setInterpData :1147, _setInterpDataChimera :1174,
  _createInterpRegion__ :1752, setInterpTransfers :1924, _setInterpTransfers :1962,
  setInterpTransfersD :2167, _setInterpTransfersD :2179, _adaptForRANSLES__ :1702.
The C entry points the reference calls (connector.setInterpData,
connector._setInterpTransfers, connector._setInterpTransfersD,
connector.getInterpolatedPoints) are replaced by the CUDA library through
cassiopee_b200._lib (include/cassiopee_b200.h).  Everything that is not the
structured-donor 'lagrangian' chimera path raises NotImplementedError — there
is no CPU fallback.

Differences that do not change results:
* device ADTs are built once per call for all donor zones (the reference
  rebuilds every donor ADT for every receptor zone when hook=None,
  OversetData.py:1223-1264 x setInterpData.cpp:772-818);
* the local->zone index remap uses numpy fancy indexing instead of the
  per-element Python loops at OversetData.py:1287-1313;
* transfer plans (device-resident interpData) are cached per ID_* subregion.
"""
import weakref
import sys as _sys
import numpy
from .. import Internal
from .. import Converter as C
from .. import _lib
from . import Connector


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
def _addCellN__(t, loc='nodes'):
    """Add cellN=1 at loc on zones that lack it (OversetData.py _addCellN__)."""
    var = 'cellN' if loc == 'nodes' else 'centers:cellN'
    for z in Internal.getZones(t):
        if C.isNamePresent(z, var) == -1:
            C._initVars(z, var, 1.)
    return None


def _zoneArray(z, withCellN=True):
    """Converter array ['x,y,z,cellN', (4,npts), ni,nj,nk] of a zone at nodes."""
    _, ni, nj, nk, _ = Internal.getZoneDim(z)
    x, y, zc = C.getCoords(z)
    n = ni * nj * nk
    rows = [x.reshape(n, order='F'), y.reshape(n, order='F'), zc.reshape(n, order='F')]
    names = 'x,y,z'
    if withCellN:
        c = C.getFieldArray(z, 'cellN')
        if c is not None:
            rows.append(c.reshape(n, order='F'))
            names += ',cellN'
    return [names, numpy.ascontiguousarray(numpy.stack(rows)), ni, nj, nk]


def _governingModel(node, default='None'):
    g = Internal.getNodeFromName2(node, 'GoverningEquations')
    return default if g is None else Internal.getValue(g)


# (model that makes a pair mixed, flag node put on the subregion): Connector/Connector/OversetData.py:1741-1748
_MIXED_MODEL_FLAGS = (('NSTurbulent', 'RANSLES'), ('LBMLaminar', 'NSLBM'))


def _adaptForRANSLES__(tR, tD, dictOfModels=None):
    """Flag the subregions whose donor and receptor zones solve different equation sets (RANS next to LES, Navier-Stokes
    next to LBM) with a `RANSLES` / `NSLBM` child, like Connector/Connector/OversetData.py:1702-1750.  A zone's model is
    its own GoverningEquations node, else its base's, else 'None' (such zones never flag anything)."""
    if dictOfModels is None:
        dictOfModels = {}
        for tree in (tR, tD):
            bases = Internal.getBases(tree)
            groups = [(b, _governingModel(b)) for b in bases] if bases != [] else [(tree, 'None')]
            for holder, inherited in groups:
                dictOfModels.update({z[0]: _governingModel(z, inherited) for z in Internal.getZones(holder)})
    receptors = {z[0] for z in Internal.getZones(tR)}
    one = lambda: numpy.ones(1, dtype=Internal.E_NpyInt)
    for zd in Internal.getZones(tD):
        md = dictOfModels.get(zd[0], 'None')
        if md == 'None': continue
        for s in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
            zrname = Internal.getValue(s)
            mr = dictOfModels.get(zrname, 'None') if zrname in receptors else 'None'
            if mr == 'None' or mr == md: continue
            for model, flag in _MIXED_MODEL_FLAGS:
                if model in (mr, md): Internal.createUniqueChild(s, flag, 'DataArray_t', one())
    return None


_REGION_PREFIX = {'chimera': 'ID_', 'abutting': 'ID_', 'ibc': 'IBCD_', 'ibc2': '2_IBCD_'}


def _createInterpRegion__(z, zname, pointlist, pointlistdonor, interpCoef, interpType, interpVol, indicesExtrap,
                          indicesOrphan, EXDir=[], pointlistdonorExtC=None, tag='Receiver', loc='centers',
                          itype='chimera', prefix=None, RotationAngle=None, RotationCenter=None):
    """Store one (donor, receptor) interpData block as a ZoneSubRegion_t of zone z.  Node names, their order and the
    way a second block for the same pair is merged are the storage contract of
    Connector/Connector/OversetData.py:1752-1884; here they are one table (`payload`) and one merge rule per node."""
    if loc == 'faces':
        raise NotImplementedError("_createInterpRegion__: loc='faces' is not on the device path.")
    name = (_REGION_PREFIX[itype] if prefix is None else prefix) + zname
    if RotationAngle is not None: RotationAngle[2] = []     # the DimensionalUnits child is not stored
    extC = None
    if pointlistdonorExtC is not None: extC = 'PointListDonorExtC' if tag == 'Receiver' else 'PointListExtC'
    # (node name, value, CGNS type, stored when) in storage order
    payload = [('PointList', pointlist, 'IndexArray_t', True),
               ('PointListDonor', pointlistdonor, 'IndexArray_t', True),
               (extC, pointlistdonorExtC, 'IndexArray_t', extC is not None),
               ('InterpolantsDonor', interpCoef, 'DataArray_t', True),
               ('InterpolantsVol', interpVol, 'DataArray_t', interpVol.size > 0),
               ('InterpolantsType', interpType, 'DataArray_t', True),
               ('OrphanPointList', indicesOrphan, 'DataArray_t', indicesOrphan.size > 0),
               ('ExtrapPointList', indicesExtrap, 'DataArray_t', indicesExtrap.size > 0)]
    found = Internal.getNodesFromName1(z, name)
    if found == []:
        region = Internal.createChild(z, name, 'ZoneSubRegion_t', value=zname)
        Internal._createChild(region, 'ZoneRole', 'DataArray_t', value=tag)
    else:
        region = found[0]
    if Internal.getNodesFromName1(region, 'InterpolantsDonor') == []:
        # first block of this pair (the region may already exist, e.g. created by a periodicity pass)
        if loc == 'centers':
            Internal._createChild(region, 'GridLocation', 'GridLocation_t', value='CellCenter')
        for nm, val, typ, keep in payload:
            if keep: region[2].append([nm, numpy.unique(val) if nm == 'OrphanPointList' else val, [], typ])
        region[2] += [n for n in (RotationAngle, RotationCenter) if n is not None]
        return None
    # further block: lists are appended to the nodes that exist (a node absent from the first block stays absent),
    # the orphan list is replaced by the new one, both ExtC flavours receive the new ExtC list
    incoming = {nm: val for nm, val, _, _ in payload if nm is not None}
    for nm in ('PointListExtC', 'PointListDonorExtC'): incoming[nm] = pointlistdonorExtC
    for node in region[2]:
        nm = node[0]
        if nm not in incoming or incoming[nm] is None: continue
        if nm == 'OrphanPointList': node[1] = numpy.unique(indicesOrphan)
        elif nm == 'InterpolantsVol' and interpVol.size == 0: continue
        else: node[1] = numpy.concatenate((node[1], incoming[nm]))
    return None


# ---------------------------------------------------------------------------
# 0. callers either side of the path, host side
# ---------------------------------------------------------------------------
def _prepBackend(backend, ctx=None):
    """The compute backend of the cellN preparers: the CUDA library unless a test injects another one."""
    if backend is None:
        backend = _lib.DeviceBackend(ctx or _lib.default_context())
    return backend


def _applyBCOverlaps(a, depth=2, loc='centers', val=2, cellNName='cellN', checkCellN=True, ctx=None, backend=None):
    """Connector.PyTree._applyBCOverlaps for structured zones (Connector/PyTree.py:1600-1645, :1699-1717):
    cellN (created =1 if absent, unless checkCellN=False: a zone without it is then left alone) is set to `val` in the
    `depth` layers next to every overset window of the zone: the non doubly-defined BCOverlap windows
    (GridConnectivity_t of type 'Overset' with a PointRange) and the BC_t windows whose family carries a
    `.Solver#Overlap` node.  The windows of a zone are applied by connector.applyBCOverlapStruct =
    cx_apply_bc_overlaps (csrc/cx_prepare.cu), one upload / download per zone."""
    if loc not in ('nodes', 'centers'):
        raise ValueError("applyBCOverlapsStruct: invalid location.")
    var = cellNName if loc == 'nodes' else 'centers:' + cellNName

    def families(node, out):            # Family_t nodes at any depth (they sit under the bases)
        for c in node[2]:
            if c[3] == 'Family_t':
                if Internal.getNodeFromName1(c, '.Solver#Overlap') is not None: out.append(c[0])
            elif c[3] in ('CGNSBase_t', 'CGNSTree_t'):
                families(c, out)
        return out
    oversetFamNames = families(a, []) if isinstance(a, list) and len(a) == 4 and isinstance(a[2], list) else []

    def window(node):
        r = Internal.getNodeFromType1(node, 'IndexRange_t')
        if r is None:
            print("Warning: applyBCOverlaps: BCOverlap is ill-defined.")
            return None
        w = r[1]
        return ((int(w[0, 0]), int(w[1, 0]), int(w[2, 0])), (int(w[0, 1]), int(w[1, 1]), int(w[2, 1])))

    for z in Internal.getZones(a):
        if C.isNamePresent(z, var) == -1:
            if not checkCellN: continue
            C._initVars(z, var, 1.)
        windows = []
        for o in Internal.getNodesFromType2(z, 'GridConnectivity_t'):
            n = Internal.getNodeFromType1(o, 'GridConnectivityType_t')
            if n is None or Internal.getValue(n) != 'Overset':
                continue
            ud = Internal.getNodeFromName1(o, 'UserDefinedData')
            if ud is not None and Internal.getNodeFromName1(ud, 'doubly_defined') is not None:
                continue
            w = window(o)
            if w is not None: windows.append(w)
        if oversetFamNames:
            for bc in Internal.getNodesFromType2(z, 'BC_t'):
                fam = Internal.getNodeFromType1(bc, 'FamilyName_t')
                if fam is not None and Internal.getValue(fam) in oversetFamNames:
                    w = window(bc)
                    if w is not None: windows.append(w)
        if windows:
            _prepBackend(backend, ctx).apply_bc_overlaps(C.getFieldArray(z, var), windows, depth, loc, val)
    return None


def applyBCOverlaps(t, depth=2, loc='centers', val=2, cellNName='cellN', ctx=None, backend=None):
    a = Internal.copyRef(t)
    for z in Internal.getZones(a):      # the cellN array is modified in place: give the copy its own
        var = cellNName if loc == 'nodes' else 'centers:' + cellNName
        c = C.getFieldArray(z, var)
        if c is not None:
            C._initVars(z, var, numpy.array(c, order='F', copy=True))
    _applyBCOverlaps(a, depth=depth, loc=loc, val=val, cellNName=cellNName, ctx=ctx, backend=backend)
    return a


def _setHoleInterpolatedPoints(a, depth=2, dir=0, loc='centers', cellNName='cellN', addGC=True, ctx=None, backend=None):
    """Connector.PyTree._setHoleInterpolatedPoints for structured zones (Connector/PyTree.py:1766-1783,
    Connector.py:168-192): cellN=2 on the `depth` layers around the blanked (cellN=0) points, by
    connector._getOversetHolesInterpCellCenters / Nodes = cx_hole_fringe (csrc/cx_prepare.cu; dir 0: cross, dir 1:
    star); depth<0 grows the fringe inwards (0 and 1 exchanged before and after, Connector.py:179-182).  The
    reference's second pass on ghost cells (addGC) only propagates through abutting (match) connectivity, which this
    path does not carry."""
    if depth == 0: return None
    if dir not in (0, 1):
        raise NotImplementedError("setHoleInterpolatedPoints: dir=%d (diamond / octahedron stencils) is not available." % dir)
    var = cellNName if loc == 'nodes' else 'centers:' + cellNName
    for z in Internal.getZones(a):
        c = C.getFieldArray(z, var)
        if c is None: continue
        _prepBackend(backend, ctx).hole_fringe(c, depth, dir)
    return None


def setHoleInterpolatedPoints(a, depth=2, dir=0, loc='centers', cellNName='cellN', ctx=None, backend=None):
    """Set cellN=2 around cellN=0 points."""
    b = Internal.copyRef(a)
    var = cellNName if loc == 'nodes' else 'centers:' + cellNName
    for z in Internal.getZones(b):
        c = C.getFieldArray(z, var)
        if c is not None:
            C._initVars(z, var, numpy.array(c, order='F', copy=True))
    _setHoleInterpolatedPoints(b, depth=depth, dir=dir, loc=loc, cellNName=cellNName, ctx=ctx, backend=backend)
    return b


def getIntersectingDomains(t, t2=None, method='AABB', taabb=None, tobb=None, taabb2=None, tobb2=None, tol=1e-10):
    """Dictionary zoneName -> names of the zones (of t2, or of t) whose axis-aligned bounding box
    intersects it (OversetData.py:115-215, method 'AABB'; compBoundingBoxIntersection compGeom.cpp:1055-1070)."""
    if method != 'AABB':
        raise NotImplementedError("getIntersectingDomains: only method='AABB' is available.")
    def boxes(tt):
        out = []
        for z in Internal.getZones(tt):
            x, y, zz = C.getCoords(z)
            out.append((z[0], (x.min(), y.min(), zz.min(), x.max(), y.max(), zz.max())))
        return out
    b1 = boxes(t)
    b2 = boxes(t2) if t2 else b1
    res = {}
    for n1, a in b1:
        res[n1] = [n2 for n2, b in b2 if n2 != n1 and not (
            a[0] > b[3] + tol or a[3] < b[0] - tol or a[1] > b[4] + tol or a[4] < b[1] - tol or
            a[2] > b[5] + tol or a[5] < b[2] - tol)]
    return res


# ---------------------------------------------------------------------------
# 1. setInterpData
# ---------------------------------------------------------------------------
def setInterpData(tR, tD, order=2, penalty=1, nature=0, extrap=1,
                  method='lagrangian', loc='nodes', storage='direct',
                  interpDataType=1, hook=None, verbose=2,
                  topTreeRcv=None, topTreeDnr=None, sameName=1, dim=3, itype='both', dictOfModels=None):
    """Compute and store overset interpolation data."""
    aR = Internal.copyRef(tR)
    aD = Internal.copyRef(tD)
    _setInterpData(aR, aD, order=order, penalty=penalty, nature=nature, extrap=extrap,
                   method=method, loc=loc, storage=storage, interpDataType=interpDataType,
                   hook=hook, verbose=verbose,
                   topTreeRcv=topTreeRcv, topTreeDnr=topTreeDnr, sameName=sameName, dim=dim, itype=itype,
                   dictOfModels=dictOfModels)
    if storage == 'direct': return aR
    else: return aD


def setInterpData2(aR, aD, order=2, loc='centers', cartesian=False):
    """Interpolation data between two different trees, every default the solver set-up uses (nature=1, penalty=1,
    extrap=1, storage='inverse', sameName=0); returns the donor tree (Connector/Connector/OversetData.py:1083-1087)."""
    aD = Internal.copyRef(aD)
    aR = Internal.copyRef(aR)
    _setInterpData2(aR, aD, order=order, loc=loc, cartesian=cartesian)
    return aD


def _setInterpData2(aR, aD, order=2, loc='centers', cartesian=False):
    """In-place form (OversetData.py:1089-1101).  A receptor tree without cellN is interpolated as a whole: cellN = 2
    is put on it for the call and removed afterwards.  cartesian=True: donors located in closed form
    (interpDataType=0)."""
    varcelln = 'cellN' if loc == 'nodes' else 'centers:cellN'
    temporary = C.isNamePresent(aR, varcelln) == -1
    if temporary: C._initVars(aR, varcelln, 2.)
    _setInterpData(aR, aD, order=order, penalty=1, nature=1, extrap=1, method='lagrangian', loc=loc,
                   storage='inverse', interpDataType=0 if cartesian else 1, sameName=0, itype='chimera')
    if temporary: C._rmVars(aR, [varcelln])
    return None


def _setInterpData(aR, aD, order=2, penalty=1, nature=0, extrap=1,
                   method='lagrangian', loc='nodes', storage='direct',
                   interpDataType=1, hook=None, verbose=2,
                   topTreeRcv=None, topTreeDnr=None, sameName=1, dim=3, itype='both', dictOfModels=None):
    # The abutting (ghost-cell) part of itype='both' only acts on zones that
    # carry match/near-match GridConnectivity nodes
    # (_setInterpDataForGhostCellsStruct__, OversetData.py:1153-1162); it has
    # no device path.
    if itype != 'chimera':
        for z in Internal.getZones(aR):
            if Internal.getNodesFromType2(z, 'GridConnectivity1to1_t') or Internal.getNodesFromType2(z, 'GridConnectivity_t'):
                raise NotImplementedError("setInterpData: abutting (ghost-cell) interpolation data has no device "
                                          "path; call with itype='chimera'.")
        if itype == 'abutting':
            return None
    if itype != 'abutting':
        _setInterpDataChimera(aR, aD, order=order, penalty=penalty, nature=nature, extrap=extrap, verbose=verbose,
                              method=method, loc=loc, storage=storage, interpDataType=interpDataType, hook=hook,
                              topTreeRcv=topTreeRcv, topTreeDnr=topTreeDnr, sameName=sameName, dim=dim, itype=itype,
                              dictOfModels=dictOfModels)
    return None


# receptor points in flight per window of _setInterpDataChimera (the per-receptor tables of a queued search stay on the
# device until its lists are fetched)
_WINDOW_PTS = 48 * 1000 * 1000


def _setInterpDataChimera(aR, aD, order=2, penalty=1, nature=0, extrap=1,
                          method='lagrangian', loc='nodes', storage='direct',
                          interpDataType=1, hook=None, verbose=2,
                          topTreeRcv=None, topTreeDnr=None, sameName=1, dim=3, itype='both', dictOfModels=None,
                          donorLists=None, ctx=None, backend=None, deviceOnly=None, refreshHooks=True):
    """donorLists (extension, optional): dict receptorZoneName -> list of donor
    zone names to consider, in that order — what Connector/Mpi.py:950-983 does
    with bbox-intersecting donors; None = all donors like the sequential API.
    deviceOnly (extension, used by Connector.Mpi): names of donor zones that exist only as device hooks
    (replicated from another rank's GPU): their tree node is a bare container for the ID_* subregions.
    refreshHooks=False: the hooks were built from the current cellN, no need to upload it again; entries of
    `hook` may then be None (built here on first use)."""
    deviceOnly = deviceOnly or ()
    locR = loc
    if loc == 'nodes': cellNPresent = C.isNamePresent(aR, 'cellN')
    elif loc == 'centers': cellNPresent = C.isNamePresent(aR, 'centers:cellN')
    else: raise ValueError("setInterpData: invalid loc.")
    if cellNPresent == -1 or itype == 'abutting': return None
    if method == 'conservative':
        if loc != 'centers':
            raise ValueError("setInterpData: conservative type is available only for loc='centers'.")
        raise NotImplementedError("setInterpData: method='conservative' has no device path.")
    if method != 'lagrangian':
        if method == 'leastsquares':
            raise NotImplementedError("setInterpData: method='leastsquares' has no device path.")
        raise ValueError("setInterpData__: %s: not a valid interpolation method." % method)
    if order not in (2, 3, 4, 5):
        print("Warning: setInterpData: unknown interpolation order. Set to 2nd order.")
        order = 2

    dictIsCellNPresent = {}
    for zd in Internal.getZones(aD):
        if zd[0] in deviceOnly:
            dictIsCellNPresent[zd[0]] = 1
            continue
        dictIsCellNPresent[zd[0]] = C.isNamePresent(zd, 'cellN')
        _addCellN__(zd, loc='nodes')

    zonesRcv = Internal.getZones(aR)
    zonesDnrL = Internal.getZones(aD)
    if backend is None:
        backend = _lib.DeviceBackend(ctx or _lib.default_context())

    # device donor zones: the hook list given by the caller, or built once here
    if hook is not None:
        if not isinstance(hook, list):
            raise TypeError("setInterpData: hook must be a list of hooks.")
        if len(hook) != len(zonesDnrL):
            raise TypeError("setInterpData: size of list of hooks must be equal to the number of donor zones.")
        allHooksL = hook[:]
        for h, zd in zip(allHooksL, zonesDnrL):   # cellN may have changed since the hook was built
            if not refreshHooks or h is None or zd[0] in deviceOnly:
                continue
            c = C.getFieldArray(zd, 'cellN')
            backend.set_celln(h, c.reshape(-1, order='F') if c is not None else None)
    else:
        allHooksL = [None] * len(zonesDnrL)

    # interpDataType: 1 (ADT) or 0 (regular Cartesian donor, InterpCart), one value or one per donor zone of aD
    # (OversetData.py:1214-1221: a scalar is broadcast).  A hook built by the caller carries its own type.
    if isinstance(interpDataType, list):
        if len(interpDataType) != len(zonesDnrL):
            raise TypeError("setInterpData: size of interpDataType must be equal to the number of donor zones.")
        typesL = list(interpDataType)
    else:
        typesL = [interpDataType] * len(zonesDnrL)
    for ty in typesL:
        if ty not in (0, 1):
            raise TypeError("setInterpData: InterpDataType must be 0 for CART or 1 for ADT.")

    def hookArgs(i):
        zd = zonesDnrL[i]
        _, ni_, nj_, nk_, _ = Internal.getZoneDim(zd)
        xd, yd, zzd = [a.reshape(-1, order='F') for a in C.getCoords(zd)]
        cd = C.getFieldArray(zd, 'cellN')
        cdf = cd.reshape(-1, order='F') if cd is not None else None
        return (ni_, nj_, nk_, xd, yd, zzd, cdf, typesL[i])

    def beginHooks(ids):
        """Start the hooks not built yet among `ids`: coordinates and cellN go up zone after zone, every ADT build starts
        on a side stream as soon as its zone is up (backend.make_zones_begin); finishHooks waits for the builds."""
        todo = [i for i in ids if allHooksL[i] is None]
        if not todo:
            return []
        started = backend.make_zones_begin([hookArgs(i) for i in todo])
        for i, h in zip(todo, started):
            allHooksL[i] = h
        return started

    def finishHooks(started):
        if started:
            backend.make_zones_finish(started)

    import os as _os, time as _time
    _prof = _os.environ.get("CX_PROFILE")
    _tt = dict(hooks=0., pts=0., search=0., store=0.)

    def finishZone(z, zonesDnr, sel, cellN, pending):
        """Etape 2 (end) and 3 for one receptor zone: fetch the lists of its search, index remap, storage."""
        _t2 = _time.perf_counter()
        resInterp = pending.fetch()
        _t3 = _time.perf_counter()
        _tt['search'] += _t3 - _t2
        nzonesDnr = len(zonesDnr)
        nbinterpolated = sel.size
        nbextrapolated = sum(r.shape[0] for r in resInterp[4])
        nborphan = resInterp[5].size
        nbinterpolated = nbinterpolated - nbextrapolated - nborphan
        if verbose != 0:
            print('Zone %s: interpolated=%d ; extrapolated=%d ; orphan=%d' % (z[0], nbinterpolated, nbextrapolated, nborphan))
            if nborphan > 0: print('Warning: zone %s has %d orphan points !' % (z[0], nborphan))
        # local position -> index in the receptor zone (indcell), vectorised; when every point of the zone is a receptor
        # the map is the identity and the lists are kept as they are
        indcells = sel.astype(Internal.E_NpyInt)
        identity = sel.size == cellN.size
        if nborphan > 0:
            if not identity: resInterp[5] = indcells[resInterp[5]]
            if verbose == 3:     # force cellN#Orphan = -1 at the orphan points (OversetData.py:1287-1294)
                pre = 'centers:' if locR == 'centers' else ''
                if Internal.getNodeFromName2(z, 'cellN#Orphan') is None:
                    C._initVars(z, pre + 'cellN#Orphan', numpy.array(cellN, dtype=numpy.float64, order='F', copy=True))
                co = C.getFieldArray(z, pre + 'cellN#Orphan')
                co.reshape(-1, order='F')[resInterp[5]] = -1.
        for noz in range(nzonesDnr):
            if resInterp[0][noz].size > 0 and not identity:
                resInterp[0][noz] = indcells[resInterp[0][noz]]
            if resInterp[4][noz].size > 0:
                if not identity: resInterp[4][noz] = indcells[resInterp[4][noz]]
                if verbose == 3:     # force cellN#Orphan = -2 at the extrapolated points (OversetData.py:1313-1319)
                    pre = 'centers:' if locR == 'centers' else ''
                    if Internal.getNodeFromName2(z, 'cellN#Orphan') is None:
                        C._initVars(z, pre + 'cellN#Orphan', numpy.array(cellN, dtype=numpy.float64, order='F', copy=True))
                    C.getFieldArray(z, pre + 'cellN#Orphan').reshape(-1, order='F')[resInterp[4][noz]] = -2.

        # Etape 3: storage in the tree
        for noz in range(nzonesDnr):
            if resInterp[0][noz].size > 0:
                extrapPts = resInterp[4][noz] if resInterp[4][noz].size > 0 else numpy.array([], dtype=Internal.E_NpyInt)
                if storage == 'direct':
                    _createInterpRegion__(z, zonesDnr[noz][0], resInterp[0][noz], resInterp[1][noz],
                                          resInterp[3][noz], resInterp[2][noz], numpy.array([], numpy.float64),
                                          extrapPts, resInterp[5], tag='Receiver', loc=locR)
                else:
                    _createInterpRegion__(zonesDnr[noz], z[0], resInterp[1][noz], resInterp[0][noz],
                                          resInterp[3][noz], resInterp[2][noz], numpy.array([], numpy.float64),
                                          extrapPts, resInterp[5], tag='Donor', loc=locR)
        _tt['store'] += _time.perf_counter() - _t3

    # The receptor zones go through in windows: the hooks the window needs are started first (coordinates up, ADT builds
    # in flight on side streams), the receptor points of every zone of the window are extracted meanwhile (Etape 1), the
    # builds are finished, every search is queued on the device without waiting for the previous one, and the lists are
    # fetched and stored zone after zone in the reference's order.  A window closes at _WINDOW_PTS points (the
    # per-receptor tables of a search stay on the device until its fetch; a zone counts with all its points).
    def donorsOf(z):
        idx = list(range(len(zonesDnrL)))
        if sameName == 1:
            for cL, zo in enumerate(zonesDnrL):
                if zo[0] == z[0]:
                    idx.pop(cL)
                    break
        if donorLists is not None and z[0] in donorLists:
            names = donorLists[z[0]]
            byname = {zonesDnrL[i][0]: i for i in idx}
            idx = [byname[n] for n in names if n in byname]
        return idx

    try:
        pos = 0
        while pos < len(zonesRcv):
            cand = []; inflight = 0
            while pos < len(zonesRcv) and (not cand or inflight < _WINDOW_PTS):
                z = zonesRcv[pos]; pos += 1
                if C.isNamePresent(z, 'cellN' if loc == 'nodes' else 'centers:cellN') == -1:
                    continue
                cand.append((z, donorsOf(z)))
                inflight += int(numpy.prod(C.fieldShape(z, locR)))
            _t1 = _time.perf_counter()
            started = beginHooks(sorted(set(i for _, idx in cand for i in idx)))
            _tt['hooks'] += _time.perf_counter() - _t1
            jobs = []
            for z, idx in cand:
                _t0 = _time.perf_counter()
                zonesDnr = [zonesDnrL[i] for i in idx]

                # Etape 1: points to interpolate (cellN == 2): getInterpolatedPoints keeps the points with
                # |cellN-2| <= 1e-15 in source order and records their index ('indcell'); for loc='centers' the
                # centre coordinates (node2Center formula) are only evaluated for those cells.  Done by the backend
                # (on the device: cx_receptors_create), which keeps the coordinates resident for the search.
                x, y, zc = C.getCoords(z)
                cellN = C.getFieldArray(z, 'cellN' if locR == 'nodes' else 'centers:cellN')
                _, ni_r, nj_r, nk_r, _ = Internal.getZoneDim(z)
                # a tree interpolated from itself at the nodes: the zone is resident as a donor hook already (same node
                # coordinates, same node cellN), its receptor points are taken from there and nothing is uploaded
                own = None
                if locR == 'nodes' and hasattr(backend, 'select_receptors_from_hook'):
                    for i, zo in enumerate(zonesDnrL):
                        if zo is z and allHooksL[i] is not None and zo[0] not in deviceOnly:
                            own = allHooksL[i]
                            break
                if own is not None:
                    rcvPts = backend.select_receptors_from_hook(own)
                else:
                    rcvPts = backend.select_receptors(ni_r, nj_r, nk_r, x.reshape(-1, order='F'), y.reshape(-1, order='F'),
                                                      zc.reshape(-1, order='F'), cellN.reshape(-1, order='F'), locR)
                sel = rcvPts.indcell
                _tt['pts'] += _time.perf_counter() - _t0
                if len(zonesDnr) == 0:
                    if sel.size > 0:
                        raise TypeError("setInterpData: no valid donor zone found.")
                    continue
                if sel.size == 0:
                    continue
                jobs.append((z, zonesDnr, idx, rcvPts, sel, cellN))
            # Etape 2: interpolation data on the device
            _t1 = _time.perf_counter()
            finishHooks(started)
            _t2 = _time.perf_counter()
            _tt['hooks'] += _t2 - _t1
            pend = [backend.submit_interp_data_rcv(rcvPts, [allHooksL[i] for i in idx], order, nature, penalty, extrap)
                    for (z, zonesDnr, idx, rcvPts, sel, cellN) in jobs]
            _tt['search'] += _time.perf_counter() - _t2
            for (z, zonesDnr, idx, rcvPts, sel, cellN), pending in zip(jobs, pend):
                finishZone(z, zonesDnr, sel, cellN, pending)
            del jobs, pend
    finally:
        if _prof:
            print("[_setInterpDataChimera] receptor extraction %.3fs, donor hooks (upload+ADT) %.3fs, search+fetch %.3fs, "
                  "remap+store %.3fs" % (_tt['pts'], _tt['hooks'], _tt['search'], _tt['store']), file=_sys.stderr, flush=True)
        if storage != 'direct':
            _adaptForRANSLES__(aR, aD, dictOfModels=dictOfModels)
        for zd in Internal.getZones(aD):
            if dictIsCellNPresent[zd[0]] == -1: C._rmVars(zd, ['cellN'])
    return None


# ---------------------------------------------------------------------------
# 2. setInterpTransfers
# ---------------------------------------------------------------------------
_PLAN_CACHE = {}


def _getPlan(ctx, zd, nptsR, ListDonor, ListRcv, DonorType, Coefs):
    """Device-resident interpData of one subregion, cached on the identity of
    its four numpy arrays (reused across the 100s of calls of a solver loop)."""
    key = (ctx.device, id(Coefs), id(ListDonor), id(ListRcv), id(DonorType), Coefs.size, ListRcv.size, nptsR)
    p = _PLAN_CACHE.get(key)
    if p is None:
        _, imd, jmd, kmd, _ = Internal.getZoneDim(zd)
        p = _lib.Plan(ctx, imd, jmd, kmd, nptsR, ListDonor, ListRcv, DonorType, Coefs)
        _PLAN_CACHE[key] = p
        for a in (Coefs, ListDonor, ListRcv, DonorType):
            try:
                weakref.finalize(a, _PLAN_CACHE.pop, key, None)
            except TypeError:
                pass
    return p


def _selectVariables(zr, zd, variables, cellNVariable, loc):
    """Names to interpolate and whether the strict cellN rule applies
    (setInterpTransfers.cpp:352-432): listed names present on both sides, or,
    for an empty list, every name common to the donor's FlowSolution and the
    receptor's container at loc; cellNVariable present on both is excluded and
    handled by the strict rule."""
    locR = 'centers' if loc == 1 else 'nodes'
    varsD = C.getVarNames(zd, 'nodes')
    varsR = C.getVarNames(zr, locR)
    if not varsD:
        raise TypeError("_setInterpTransfers: no field found in donor zones.")
    if not varsR:
        raise TypeError("_setInterpTransfers: no field found in receiver zones.")
    poscd = cellNVariable in varsD
    poscr = cellNVariable in varsR
    cellnD = cellnR = False
    if isinstance(variables, list) and len(variables) > 0:
        names = []
        for v in variables:
            if not isinstance(v, str):
                continue
            if v == cellNVariable and poscd: cellnD = True
            if v == cellNVariable and poscr: cellnR = True
            if v in varsD and v in varsR:
                names.append(v)
    else:
        names = [v for v in varsR if v in varsD]
        cellnD, cellnR = poscd, poscr
    strict = cellnD and cellnR
    if strict:
        names = [v for v in names if v != cellNVariable]
    return names, strict, locR


def setInterpTransfers(aR, topTreeD, variables=[], cellNVariable='cellN',
                       variablesIBC=['Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature'],
                       bcType=0, varType=2, storage=-1, compact=0, compactD=0,
                       Gamma=1.4, Cv=1.7857142857142865, MuS=1.e-08,
                       Cs=0.3831337844872463, Ts=1.0):
    """Transfer variables once interpolation data has been computed."""
    tR = Internal.copyRef(aR)
    _setInterpTransfers(tR, topTreeD, variables=variables, variablesIBC=variablesIBC,
                        bcType=bcType, varType=varType, storage=storage, compact=compact, compactD=compactD,
                        cellNVariable=cellNVariable, Gamma=Gamma, Cv=Cv, MuS=MuS, Cs=Cs, Ts=Ts)
    return tR


def _rotationOf(s):
    """(dirR, theta) of an ID_* subregion (OversetData.py:2000-2006, setInterpTransfers.cpp:281-285): the first
    non-zero component of the RotationAngle child selects the axis; the value goes to cos/sin as it is."""
    node = Internal.getNodeFromName1(s, 'RotationAngle')
    if node is None:
        return 0, 0.
    ang = numpy.asarray(node[1], dtype=numpy.float64).reshape(-1)
    for d in range(3):
        if abs(ang[d]) > 0.:
            return d + 1, float(ang[d])
    return 0, 0.


def _rotationTriplets(zr, locR, nvars):
    """Positions rotated by includeTransfers.h: VelocityX/Y/Z and MomentumX/Y/Z are looked up in the receptor
    container's variable string (setInterpTransfers.cpp:401-410) and used as indices into the list of
    transferred fields (vectOfRcvFields), exactly like the reference."""
    varsR = C.getVarNames(zr, locR)
    trip = []
    for base in ('Velocity', 'Momentum'):
        pos = [varsR.index(base + a) if (base + a) in varsR else -1 for a in 'XYZ']
        if min(pos) > -1:
            if max(pos) >= nvars:
                raise ValueError("_setInterpTransfers: %sX/Y/Z sit at positions %s of the receptor fields but only %d "
                                 "fields are transferred (the reference indexes out of bounds here)." % (base, pos, nvars))
            trip.append(tuple(pos))
    return trip


def _compactFields(first, nvars, ndimdx, what):
    """compact=1 (getfromzoneDcompact.h / getfromzoneRcompact.h): only the first variable is looked up, field eq
    lives at first + eq*ndimdx doubles.  The views are checked against the memory block that owns the first
    array, so non-compacted trees raise instead of reading out of bounds."""
    if first is None:
        raise TypeError("_setInterpTransfers: compact=1: first variable not found in the %s zone." % what)
    if first.dtype != numpy.float64 or first.size != ndimdx or not (first.flags.f_contiguous or first.flags.c_contiguous):
        raise TypeError("_setInterpTransfers: compact=1: %s field must be a contiguous float64 array of the zone size." % what)
    owner = first
    while isinstance(owner.base, numpy.ndarray):
        owner = owner.base
    lo = owner.__array_interface__['data'][0]
    hi = lo + owner.nbytes
    a0 = first.__array_interface__['data'][0]
    if not owner.flags.c_contiguous and not owner.flags.f_contiguous:
        raise ValueError("_setInterpTransfers: compact=1: the %s fields are not compacted." % what)
    if a0 < lo or a0 + 8 * ndimdx * nvars > hi:
        raise ValueError("_setInterpTransfers: compact=1: the %s fields are not compacted (%d fields of %d values "
                         "expected behind the first variable)." % (what, nvars, ndimdx))
    flat = owner.reshape(-1, order='A').view(numpy.float64) if owner.dtype == numpy.float64 else None
    if flat is None:
        raise TypeError("_setInterpTransfers: compact=1: %s block is not float64." % what)
    o0 = (a0 - lo) // 8
    return [flat[o0 + eq * ndimdx: o0 + (eq + 1) * ndimdx] for eq in range(nvars)]


def _transferOneCompact(ctx, zr, zd, s, variables, varType, ListRcv, ListDonor):
    """connector._setInterpTransfers, compact != 0 (setInterpTransfers.cpp:474-497): 5 fields (1 <= varType <= 3)
    or 6, contiguous behind variables[0] on both sides; no cellN rule, no rotation (the reference warns)."""
    location = Internal.getNodeFromName1(s, 'GridLocation')
    if location is not None: location = Internal.getValue(location)
    loc = 1 if location == 'CellCenter' else 0
    if not isinstance(variables, list) or len(variables) == 0 or not isinstance(variables[0], str):
        raise TypeError("_setInterpTransfers: compact=1 needs the name of the first variable.")
    nvars = 5 if 1 <= varType <= 3 else 6
    # containers as getfromzoneDcompact.h:2-3 / getfromzoneRcompact.h:5-6 select them
    solD = Internal.getNodeFromName1(zd, Internal.__FlowSolutionCenters__ if loc == 0 else Internal.__FlowSolutionNodes__)
    solR = Internal.getNodeFromName1(zr, Internal.__FlowSolutionNodes__ if loc == 0 else Internal.__FlowSolutionCenters__)
    nD = None if solD is None else Internal.getNodeFromName1(solD, variables[0])
    nR = None if solR is None else Internal.getNodeFromName1(solR, variables[0])
    _, imd, jmd, kmd, _ = Internal.getZoneDim(zd)
    ndimdxD = imd * jmd * kmd
    ndimdxR = int(numpy.prod(C.fieldShape(zr, 'centers' if loc == 1 else 'nodes')))
    fD = _compactFields(None if nD is None else nD[1], nvars, ndimdxD, 'donor')
    fR = _compactFields(None if nR is None else nR[1], nvars, ndimdxR, 'receiver')
    Coefs = Internal.getNodeFromName1(s, 'InterpolantsDonor')[1]
    DonorType = Internal.getNodeFromName1(s, 'InterpolantsType')[1]
    plan = _getPlan(ctx, zd, ndimdxR, ListDonor, ListRcv, DonorType, Coefs)
    plan.transfer(fD, fR)
    if _rotationOf(s)[0] != 0:
        print("Warning: _setInterpTransfers: no correction for periodicity by rotation can be applied in compacted version.")


def _transferOne(ctx, zr, zd, s, variables, cellNVariable, ListRcv, ListDonor, varType=2, compact=0):
    if compact != 0:
        return _transferOneCompact(ctx, zr, zd, s, variables, varType, ListRcv, ListDonor)
    location = Internal.getNodeFromName1(s, 'GridLocation')
    if location is not None: location = Internal.getValue(location)
    loc = 1 if location == 'CellCenter' else 0
    dirR, theta = _rotationOf(s)
    Coefs = Internal.getNodeFromName1(s, 'InterpolantsDonor')[1]
    DonorType = Internal.getNodeFromName1(s, 'InterpolantsType')[1]
    names, strict, locR = _selectVariables(zr, zd, variables, cellNVariable, loc)
    nptsR = int(numpy.prod(C.fieldShape(zr, locR)))
    plan = _getPlan(ctx, zd, nptsR, ListDonor, ListRcv, DonorType, Coefs)
    trip = _rotationTriplets(zr, locR, len(names)) if dirR != 0 else []
    if names:
        fD = [C.getFieldArray(zd, v) for v in names]
        fR = [C.getFieldArray(zr, (('centers:' + v) if locR == 'centers' else v)) for v in names]
        if trip:
            plan.transfer_rot(fD, fR, dirR, theta, trip)
        else:
            plan.transfer(fD, fR)
    if strict:
        cD = C.getFieldArray(zd, cellNVariable)
        cR = C.getFieldArray(zr, ('centers:' + cellNVariable) if locR == 'centers' else cellNVariable)
        plan.transfer_celln(cD, cR)


def _setInterpTransfers(aR, topTreeD, variables=[], cellNVariable='',
                        variablesIBC=['Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature'],
                        bcType=0, varType=2, storage=-1, compact=0, compactD=0,
                        Gamma=1.4, Cv=1.7857142857142865, MuS=1.e-08, Cs=0.3831337844872463, Ts=1.0, ctx=None):
    ctx = ctx or _lib.default_context()
    # every subregion of a donor zone reads the same donor arrays: one upload per array for the whole call
    ctx.batch_begin()
    try:
        _setInterpTransfersBody(ctx, aR, topTreeD, variables, cellNVariable, varType, storage, compact)
    finally:
        ctx.batch_end()
    return None


def _setInterpTransfersBody(ctx, aR, topTreeD, variables, cellNVariable, varType, storage, compact):
    # direct storage: data held by the receptor zones (OversetData.py:1968-2033)
    if storage < 1:
        znd = {z[0]: z for z in Internal.getZones(topTreeD)}
        for zr in Internal.getZones(aR):
            for s in Internal.getNodesFromType1(zr, 'ZoneSubRegion_t'):
                sname = s[0][0:2]
                if sname == 'ID' and variables is not None:
                    idn = Internal.getNodeFromName1(s, 'InterpolantsDonor')
                    if idn is not None:
                        zoneRole = Internal.getValue(Internal.getNodeFromName2(s, 'ZoneRole'))
                        if zoneRole == 'Receiver':
                            ListRcv = Internal.getNodeFromName1(s, 'PointList')[1]
                            ListDonor = Internal.getNodeFromName1(s, 'PointListDonor')[1]
                            zd = znd[Internal.getValue(s)]
                            _transferOne(ctx, zr, zd, s, variables, cellNVariable, ListRcv, ListDonor, varType, compact)
    # inverse storage: data held by the donor zones (OversetData.py:2036-2098)
    if storage != 0:
        znr = {z[0]: z for z in Internal.getZones(aR)}
        for zd in Internal.getZones(topTreeD):
            ctx.batch_flush()          # the previous donor zone's fields are not read again
            for s in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
                sname = s[0][0:2]
                if sname == 'ID' and variables is not None:
                    idn = Internal.getNodeFromName1(s, 'InterpolantsDonor')
                    if idn is not None:
                        zoneRole = Internal.getValue(Internal.getNodeFromName2(s, 'ZoneRole'))
                        if zoneRole == 'Donor':
                            ListDonor = Internal.getNodeFromName1(s, 'PointList')[1]
                            ListRcv = Internal.getNodeFromName1(s, 'PointListDonor')[1]
                            zr = znr.get(Internal.getValue(s), None)
                            if zr is not None:
                                _transferOne(ctx, zr, zd, s, variables, cellNVariable, ListRcv, ListDonor, varType, compact)
    return None


def setInterpTransfersD(topTreeD, variables=[], cellNVariable='',
                        variablesIBC=['Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature'],
                        bcType=0, varType=2, compact=0, compactD=0,
                        Gamma=1.4, Cv=1.7857142857142865, MuS=1.e-08,
                        Cs=0.3831337844872463, Ts=1.0, extract=0, alpha=1.):
    tD = Internal.copyRef(topTreeD)
    return _setInterpTransfersD(tD, variables=variables, cellNVariable=cellNVariable, variablesIBC=variablesIBC,
                                bcType=bcType, varType=varType, compact=compact, compactD=compactD,
                                Gamma=Gamma, Cv=Cv, MuS=MuS, Cs=Cs, Ts=Ts, extract=extract)


def _setInterpTransfersD(topTreeD, variables=[], cellNVariable='',
                         variablesIBC=['Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature'],
                         bcType=0, varType=2, compact=0, compactD=0,
                         Gamma=1.4, Cv=1.7857142857142865, MuS=1.e-08,
                         Cs=0.3831337844872463, Ts=1.0, extract=0, alpha=1., ctx=None, backend=None):
    """Donor-side transfer: list of [rcvZoneName, array(nvars,nrcv), ListRcv, loc]
    (OversetData.py:2179-2270; connector._setInterpTransfersD,
    setInterpTransfersD.cpp:106-303: listed variables present in the donor's
    FlowSolution, or all of them but cellNVariable for an empty list)."""
    if backend is None:
        ctx = ctx or _lib.default_context()
        ctx.batch_begin()          # a donor zone's fields are uploaded once for all its subregions
        try:
            return _setInterpTransfersDBody(topTreeD, variables, cellNVariable, varType, compact, ctx, backend)
        finally:
            ctx.batch_end()
    return _setInterpTransfersDBody(topTreeD, variables, cellNVariable, varType, compact, ctx, backend)


def _setInterpTransfersDBody(topTreeD, variables, cellNVariable, varType, compact, ctx, backend):
    infos = []
    for zd in Internal.getZones(topTreeD):
        if backend is None: ctx.batch_flush()
        for s in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
            if s[0][0:2] == 'ID' and variables is not None:
                idn = Internal.getNodeFromName1(s, 'InterpolantsDonor')
                if idn is None:
                    continue
                if Internal.getValue(Internal.getNodeFromName2(s, 'ZoneRole')) != 'Donor':
                    continue
                location = Internal.getNodeFromName1(s, 'GridLocation')
                if location is not None: location = Internal.getValue(location)
                loc = 'centers' if location == 'CellCenter' else 'nodes'
                Coefs = idn[1]
                DonorType = Internal.getNodeFromName1(s, 'InterpolantsType')[1]
                ListDonor = Internal.getNodeFromName1(s, 'PointList')[1]
                ListRcv = Internal.getNodeFromName1(s, 'PointListDonor')[1]
                if compact != 0:
                    # setInterpTransfersD.cpp:232-235, :282-296, getfromzonecompactD.h: 5 (1 <= varType <= 3) or 6
                    # fields behind variables[0] in the donor's FlowSolution; the names given label the array
                    if backend is not None:
                        raise NotImplementedError("_setInterpTransfersD: compact=1 needs the device backend.")
                    if not isinstance(variables, list) or len(variables) == 0 or not isinstance(variables[0], str):
                        raise TypeError("_setInterpTransfersD: compact=1 needs the name of the first variable.")
                    nvc = 5 if 1 <= varType <= 3 else 6
                    _, imd, jmd, kmd, _ = Internal.getZoneDim(zd)
                    solD = Internal.getNodeFromName1(zd, Internal.__FlowSolutionNodes__)
                    nD = None if solD is None else Internal.getNodeFromName1(solD, variables[0])
                    fDc = _compactFields(None if nD is None else nD[1], nvc, imd * jmd * kmd, 'donor')
                    nmax = int(ListRcv.max()) + 1 if ListRcv.size else 1
                    plan = _getPlan(ctx, zd, nmax, ListDonor, ListRcv, DonorType, Coefs)
                    arr = plan.transferD(fDc)
                    infos.append([Internal.getValue(s), [','.join(variables), arr, ListRcv.size, 1, 1], ListRcv, loc])
                    continue
                varsD = C.getVarNames(zd, 'nodes')
                if isinstance(variables, list) and len(variables) > 0:
                    names = [v for v in variables if v in varsD and v != cellNVariable]
                    withCellN = cellNVariable in variables and cellNVariable in varsD
                else:
                    names = [v for v in varsD if v != cellNVariable]
                    withCellN = False
                nmax = int(ListRcv.max()) + 1 if ListRcv.size else 1
                if backend is None:
                    plan = _getPlan(ctx, zd, nmax, ListDonor, ListRcv, DonorType, Coefs)
                else:
                    _, imd, jmd, kmd, _ = Internal.getZoneDim(zd)
                    plan = backend.make_plan(imd, jmd, kmd, nmax, ListDonor, ListRcv, DonorType, Coefs)
                arr = plan.transferD([C.getFieldArray(zd, v) for v in names])
                varString = ','.join(names)
                if withCellN:
                    # strict cellN rule evaluated donor side (setInterpTransfersD.cpp:159-180)
                    cR = numpy.zeros(nmax)
                    plan.transfer_celln(C.getFieldArray(zd, cellNVariable), cR)
                    arr = numpy.vstack([arr, cR[ListRcv][None, :]])
                    varString = (varString + ',' if varString else '') + cellNVariable
                infos.append([Internal.getValue(s), [varString, arr, ListRcv.size, 1, 1], ListRcv, loc])
    return infos
