"""Converter.PyTree helpers the chimera path needs on the host side.

Restated minimal subset (plain numpy on CGNS/Python trees):

* node2Center      Converter/Converter/PyTree.py (C.node2Center) +
                   KCore/KCore/Loc/node2Center.cpp:118-146  (0.125*sum of 8 nodes,
                   left-associated in the order ind0..ind7)
* _initVars / _addVars / _rmVars / isNamePresent / getField-style accessors
* newPyTree, getAllFields helpers

Fields live in ``FlowSolution`` (nodes) and ``FlowSolution#Centers`` (centres),
one Fortran-ordered (ni,nj,nk) float64 numpy per variable.
"""
import numpy
from . import Internal


def newPyTree(args=None):
    """newPyTree(['Base', z1, z2, 'Base2', z3]) like Converter.PyTree.newPyTree."""
    t = Internal.newCGNSTree()
    base = None
    for a in (args or []):
        if isinstance(a, str):
            base = Internal.newCGNSBase(a, 3, 3, parent=t)
        elif isinstance(a, list):
            if base is None:
                base = Internal.newCGNSBase('Base', 3, 3, parent=t)
            if Internal.isStdNode(a) == -1:
                base[2].append(a)
            else:
                for z in a:
                    base[2].append(z)
    return t


def _container(z, loc, create=False):
    name = Internal.__FlowSolutionCenters__ if loc == 'centers' else Internal.__FlowSolutionNodes__
    c = Internal.getNodeFromName1(z, name)
    if c is None and create:
        c = Internal.createChild(z, name, 'FlowSolution_t')
        if loc == 'centers':
            Internal.createChild(c, 'GridLocation', 'GridLocation_t', 'CellCenter')
    return c


def _splitName(var):
    if var.startswith('centers:'):
        return 'centers', var[8:]
    if var.startswith('nodes:'):
        return 'nodes', var[6:]
    return 'nodes', var


def isNamePresent(t, var):
    """1 if var is in all zones, 0 if in some, -1 if in none (C.isNamePresent)."""
    loc, name = _splitName(var)
    zones = Internal.getZones(t)
    if not zones:
        return -1
    n = 0
    for z in zones:
        c = _container(z, loc)
        if c is not None and Internal.getNodeFromName1(c, name) is not None:
            n += 1
    if n == len(zones):
        return 1
    return 0 if n > 0 else -1


def getCoords(z):
    gc = Internal.getNodeFromName1(z, Internal.__GridCoordinates__)
    return [Internal.getNodeFromName1(gc, n)[1] for n in ('CoordinateX', 'CoordinateY', 'CoordinateZ')]


def getFieldArray(z, var):
    loc, name = _splitName(var)
    c = _container(z, loc)
    if c is None:
        return None
    n = Internal.getNodeFromName1(c, name)
    return None if n is None else n[1]


def fieldShape(z, loc):
    _, ni, nj, nk, _ = Internal.getZoneDim(z)
    if loc == 'centers':
        return (max(ni - 1, 1), max(nj - 1, 1), max(nk - 1, 1))
    return (ni, nj, nk)


def _initVars(t, var, value=0.):
    """Set a field: constant, ndarray, or callable f(x,y,z) of the coordinates at loc."""
    loc, name = _splitName(var)
    for z in Internal.getZones(t):
        shp = fieldShape(z, loc)
        if callable(value):
            x, y, zc = getCoords(z)
            if loc == 'centers':
                x, y, zc = [node2CenterArray(a) for a in (x, y, zc)]
            a = numpy.asfortranarray(numpy.asarray(value(x, y, zc), dtype=numpy.float64).reshape(shp, order='F'))
        elif isinstance(value, numpy.ndarray):
            a = numpy.asfortranarray(value.astype(numpy.float64).reshape(shp, order='F'))
        else:
            a = numpy.full(shp, float(value), dtype=numpy.float64, order='F')
        c = _container(z, loc, create=True)
        n = Internal.getNodeFromName1(c, name)
        if n is None:
            Internal.createChild(c, name, 'DataArray_t', a)
        else:
            n[1] = a
    return None


def initVars(t, var, value=0.):
    tp = Internal.copyRef(t)
    _initVars(tp, var, value)
    return tp


def _rmVars(t, variables):
    if isinstance(variables, str):
        variables = [variables]
    for z in Internal.getZones(t):
        for v in variables:
            loc, name = _splitName(v)
            c = _container(z, loc)
            if c is not None:
                c[2][:] = [n for n in c[2] if n[0] != name]
    return None


def getVarNames(z, loc='nodes'):
    c = _container(z, loc)
    if c is None:
        return []
    return [n[0] for n in c[2] if n[3] == 'DataArray_t']


def node2CenterArray(a):
    """(ni,nj,nk) node array -> (ni-1,nj-1,nk-1) centre array; 3-D formula of
    node2Center.cpp:118-146, summation order ind0..ind7."""
    a = numpy.asarray(a)
    ni, nj, nk = a.shape
    if nj == 1 and nk == 1:
        return numpy.asfortranarray(0.5 * (a[:-1] + a[1:]))      # curve along i (node2Center.cpp:80-95)
    if nk == 1:
        s = a[:-1, :-1, :] + a[1:, :-1, :]
        s = s + a[:-1, 1:, :]
        s = s + a[1:, 1:, :]
        return numpy.asfortranarray(0.25 * s)
    s = a[:-1, :-1, :-1] + a[1:, :-1, :-1]
    s = s + a[:-1, 1:, :-1]
    s = s + a[1:, 1:, :-1]
    s = s + a[:-1, :-1, 1:]
    s = s + a[1:, :-1, 1:]
    s = s + a[:-1, 1:, 1:]
    s = s + a[1:, 1:, 1:]
    return numpy.asfortranarray(0.125 * s)


def node2CenterSelected(a, sel):
    """Centre values of the cells `sel` only (flat i-fastest cell indices), same
    summation order as node2CenterArray, hence bit-identical to
    node2CenterArray(a).reshape(-1, order='F')[sel]."""
    a = numpy.asarray(a)
    ni, nj, nk = a.shape
    f = a.reshape(-1, order='F')
    sel = numpy.asarray(sel, dtype=numpy.int64)
    nic, njc = max(ni - 1, 1), max(nj - 1, 1)
    k = sel // (nic * njc)
    j = (sel - k * nic * njc) // nic
    i = sel - j * nic - k * nic * njc
    n0 = i + ni * j + ni * nj * k
    if nj == 1 and nk == 1:
        return 0.5 * (f[sel] + f[sel + 1])
    if nk == 1:
        s = f[n0] + f[n0 + 1]
        s = s + f[n0 + ni]
        s = s + f[n0 + ni + 1]
        return 0.25 * s
    nij = ni * nj
    s = f[n0] + f[n0 + 1]
    s = s + f[n0 + ni]
    s = s + f[n0 + ni + 1]
    s = s + f[n0 + nij]
    s = s + f[n0 + nij + 1]
    s = s + f[n0 + nij + ni]
    s = s + f[n0 + nij + ni + 1]
    return 0.125 * s


def node2Center(t):
    """C.node2Center(t): tree of the centre meshes.  Coordinates become the
    cell centres; centre fields (FlowSolution#Centers) become node fields of
    the new zones; node fields are averaged to centres.  This is the ``tc``
    convention of the reference (Apps/Apps/Fast/MB.py:101-103)."""
    tp = Internal.copyRef(t)
    for z in Internal.getZones(tp):
        _, ni, nj, nk, _ = Internal.getZoneDim(z)
        x, y, zc = getCoords(z)
        xc, yc, zcc = [node2CenterArray(a) for a in (x, y, zc)]
        nic, njc, nkc = xc.shape
        v = numpy.zeros((3, 3), dtype=Internal.E_NpyInt, order='F')
        v[:, 0] = (nic, njc, nkc)
        v[:, 1] = (max(nic - 1, 1), max(njc - 1, 1), max(nkc - 1, 1))
        z[1] = v
        gc = Internal.getNodeFromName1(z, Internal.__GridCoordinates__)
        for nm, a in (('CoordinateX', xc), ('CoordinateY', yc), ('CoordinateZ', zcc)):
            Internal.getNodeFromName1(gc, nm)[1] = a
        fsn = Internal.getNodeFromName1(z, Internal.__FlowSolutionNodes__)
        fsc = Internal.getNodeFromName1(z, Internal.__FlowSolutionCenters__)
        newfields = []
        if fsn is not None:
            for n in fsn[2]:
                if n[3] == 'DataArray_t':
                    newfields.append([n[0], node2CenterArray(n[1]), [], 'DataArray_t'])
        if fsc is not None:
            names = [f[0] for f in newfields]
            for n in fsc[2]:
                if n[3] == 'DataArray_t' and n[0] not in names:
                    # a copy, like the reference (getFields(..., api=1) + setFields build new arrays,
                    # Converter/PyTree.py:6400-6414): tc's fields must not alias t's, or a transfer that writes a
                    # receptor zone of t would change what the next subregion reads from the same zone of tc
                    newfields.append([n[0], numpy.array(n[1], order='F', copy=True), [], 'DataArray_t'])
        z[2][:] = [c for c in z[2] if c[0] not in (Internal.__FlowSolutionNodes__, Internal.__FlowSolutionCenters__)]
        if newfields:
            c = Internal.createChild(z, Internal.__FlowSolutionNodes__, 'FlowSolution_t')
            c[2].extend(newfields)
    return tp


def _addBC2Zone(z, bndName, bndType, wrange):
    """C._addBC2Zone for structured zones, windows given as [imin,imax,jmin,jmax,kmin,kmax] (1-based).
    'BCOverlap' -> ZoneGridConnectivity/GridConnectivity_t with GridConnectivityType 'Overset' and a
    PointRange (what Connector.applyBCOverlaps looks for, Connector/PyTree.py:1609-1633);
    anything else -> ZoneBC/BC_t with a PointRange."""
    r = numpy.zeros((3, 2), dtype=Internal.E_NpyInt, order='F')
    r[0, :] = wrange[0:2]; r[1, :] = wrange[2:4]; r[2, :] = wrange[4:6]
    if bndType == 'BCOverlap':
        zgc = Internal.getNodeFromName1(z, 'ZoneGridConnectivity')
        if zgc is None:
            zgc = Internal.createChild(z, 'ZoneGridConnectivity', 'ZoneGridConnectivity_t')
        gc = Internal.createChild(zgc, bndName, 'GridConnectivity_t', value=z[0])
        gc[2].append(['PointRange', r, [], 'IndexRange_t'])
        Internal.createChild(gc, 'GridConnectivityType', 'GridConnectivityType_t', value='Overset')
    else:
        zbc = Internal.getNodeFromName1(z, 'ZoneBC')
        if zbc is None:
            zbc = Internal.createChild(z, 'ZoneBC', 'ZoneBC_t')
        bc = Internal.createChild(zbc, bndName, 'BC_t', value=bndType)
        bc[2].append(['PointRange', r, [], 'IndexRange_t'])
    return None


def createHook(a, function='None', ctx=None, backend=None):
    """C.createHook(z, 'adt') (Converter/Converter/PyTree.py:7198-7209, hook.cpp:578-632): the ADT of one
    structured zone, built once and passed to setInterpData(hook=[...]) so that it is not rebuilt per call.
    Here the hook is the zone resident on the device with its ADT replica (cassiopee_b200._lib.Zone);
    function 'cart' (extension) gives the closed-form Cartesian donor of interpDataType=0 instead.
    cellN is not part of a hook: setInterpData refreshes it from the donor tree at every call."""
    from . import _lib
    if function not in ('adt', 'cart'):
        raise NotImplementedError("createHook: only function='adt' (and 'cart') has a device hook.")
    zones = Internal.getZones(a)
    if len(zones) != 1:
        raise TypeError("createHook: 'adt' hooks are built per zone: one zone expected, got %d." % len(zones))
    z = zones[0]
    _, ni, nj, nk, _ = Internal.getZoneDim(z)
    x, y, zz = [c.reshape(-1, order='F') for c in getCoords(z)]
    if backend is None:
        backend = _lib.DeviceBackend(ctx or _lib.default_context())
    if function == 'cart':
        return backend.make_zone(ni, nj, nk, x, y, zz, None, interpDataType=0)
    return backend.make_zone(ni, nj, nk, x, y, zz, None)


def freeHook(hook):
    """C.freeHook: release the device memory of a hook now (it is also released when the object dies)."""
    fin = getattr(hook, '_fin', None)
    if fin is not None:
        fin()
    elif hasattr(hook, 'free'):
        hook.free()
    return None


# -- file I/O: the reference's 'bin_pickle' format (Converter/Converter/PyTree.py convertPyTree2File /
#    convertFile2PyTree with format='bin_pickle': the tree pickled as is).  Enough to carry a donor tree tc with
#    its ID_* subregions from a preparation run to a solver run; CGNS/HDF5 files are out
#    of scope (no HDF5 in this image).
def _formatOf(fileName, format):
    if format is not None:
        return format
    ext = fileName.rsplit('.', 1)[-1].lower() if '.' in fileName else ''
    return {'pickle': 'bin_pickle', 'pkl': 'bin_pickle', 'ref': 'bin_pickle'}.get(ext, ext)


def convertPyTree2File(t, fileName, format=None):
    """Write a pyTree to a file (format 'bin_pickle' only)."""
    import pickle
    fmt = _formatOf(fileName, format)
    if fmt != 'bin_pickle':
        raise NotImplementedError("convertPyTree2File: only format='bin_pickle' is available (got %r)." % fmt)
    with open(fileName, 'wb') as f:
        pickle.dump(t, f, protocol=pickle.HIGHEST_PROTOCOL)
    return None


def convertFile2PyTree(fileName, format=None):
    """Read a pyTree from a file (format 'bin_pickle' only)."""
    import pickle
    fmt = _formatOf(fileName, format)
    if fmt != 'bin_pickle':
        raise NotImplementedError("convertFile2PyTree: only format='bin_pickle' is available (got %r)." % fmt)
    with open(fileName, 'rb') as f:
        try:
            t = pickle.load(f, encoding='latin1')
        except TypeError:
            t = pickle.load(f)
    if not (isinstance(t, list) and len(t) == 4):
        raise TypeError("convertFile2PyTree: %s does not hold a pyTree." % fileName)
    return t
