"""Post.extractMesh on the device (the other consumer of the K_INTERP search).

Mirror of Post/Post/PyTree.py:160-230 + Post/Post/extractMesh.cpp:42-470 for the part that shares the chimera hot
path: structured 3-D donor zones with node-located fields.  Every node of the extraction mesh is located with
K_INTERP::getInterpolationCell(nature=0, penalty=0) over the donor zones in order, then, if not found, with
getExtrapolationCell (constraint 40, extrapOrder 1 — the defaults, the only values the device search implements);
all donor fields are interpolated with the resulting stencil (compInterpolatedValues = the transfer kernels);
points without donor keep 0.  Donor centre fields would first go through center2Node (mode 'robust') or be
extracted on the centre mesh (mode 'accurate'): not on the device path, NotImplementedError.
"""
import numpy
from . import Internal
from . import Converter as C
from . import _lib


def extractMesh(t, extractionMesh, order=2, extrapOrder=1, constraint=40., tol=1.e-6, hook=None, mode='robust',
                ctx=None):
    """Extract the solution on a given mesh.
    Usage: extractMesh(t, extractMesh, order, extrapOrder, constraint, tol, hook)"""
    te = Internal.copyRef(extractionMesh)
    for z in Internal.getZones(te):        # the copy gets its own FlowSolution container
        z[2][:] = [c for c in z[2] if c[0] != Internal.__FlowSolutionNodes__]
    _extractMesh(t, te, order, extrapOrder, constraint, tol, hook, mode, ctx=ctx)
    return te


def _extractMesh(t, extractionMesh, order=2, extrapOrder=1, constraint=40., tol=1.e-6, hook=None, mode='robust',
                 ctx=None):
    if extrapOrder != 1 or constraint != 40.:
        raise NotImplementedError("extractMesh: the device search implements extrapOrder=1, constraint=40. only.")
    if order not in (2, 3, 5):
        raise ValueError("extractMesh: order must be 2, 3 or 5.")
    zonesD = Internal.getZones(t)
    if len(zonesD) == 0:
        return None
    for zd in zonesD:
        if C.getVarNames(zd, 'centers'):
            raise NotImplementedError("extractMesh: donor fields located at centres (center2Node) have no device path.")
    names = C.getVarNames(zonesD[0], 'nodes')
    for zd in zonesD[1:]:
        if len(C.getVarNames(zd, 'nodes')) != len(names):
            raise TypeError("extractMesh: all arrays must have the same number of variables.")
        if set(C.getVarNames(zd, 'nodes')) != set(names):
            raise TypeError("extractMesh: donor zones must carry the same variables.")
    if hook is not None:
        if not isinstance(hook, list):
            raise TypeError("_extractMesh: hook must be a list of hooks on ADTs.")
        if len(hook) != len(zonesD):
            raise TypeError("extractMesh: size of list of hooks must be equal to the number of donor zones.")
    ctx = ctx or _lib.default_context()
    hooks = []
    dims = []
    for i, zd in enumerate(zonesD):
        _, ni, nj, nk, _ = Internal.getZoneDim(zd)
        dims.append((ni, nj, nk))
        if hook is not None:
            hooks.append(hook[i])
        else:
            x, y, z = [a.reshape(-1, order='F') for a in C.getCoords(zd)]
            c = C.getFieldArray(zd, 'cellN')
            hooks.append(_lib.Zone(ctx, ni, nj, nk, x, y, z, c.reshape(-1, order='F') if c is not None else None))
    for ze in Internal.getZones(extractionMesh):
        x, y, z = [numpy.ascontiguousarray(a.reshape(-1, order='F')) for a in C.getCoords(ze)]
        npts = x.size
        res = _lib.set_interp_data(ctx, x, y, z, hooks, order, 0, 0, 1).fetch()
        out = [numpy.zeros(npts) for _ in names]
        for noz, zd in enumerate(zonesD):
            rcv, dnr, typ, cf = res[0][noz], res[1][noz], res[2][noz], res[3][noz]
            if rcv.size == 0:
                continue
            ni, nj, nk = dims[noz]
            plan = _lib.Plan(ctx, ni, nj, nk, npts, dnr, rcv, typ, cf)
            plan.transfer([C.getFieldArray(zd, v) for v in names], out)
        shp = C.fieldShape(ze, 'nodes')
        for v, a in zip(names, out):
            C._initVars(ze, v, a.reshape(shp, order='F'))
    return None
