"""Zone -> processor placement, consumed by the distributed chimera path.

The reference writes one integer per zone, ``.Solver#Param/proc``
(Distributor2/Distributor2/PyTree.py:70-101) and reads it back with
getProcDict (:268).  That node is the contract honoured here; the balancing
algorithms themselves (METIS k-way graph.cpp:298, genetic, gradient) are out of
scope (64 integers, not a hot path).  Two simple stand-ins are
provided so synthetic cases can be placed:
  'fast' : greedy balance of the point counts (largest zone first to the
           least loaded processor);
  'rcb'  : recursive coordinate bisection of the zone bounding-box centres,
           which yields the spatially compact parts a graph partitioner gives
           on block-structured clusters (minimises cross-GPU pairs).
"""
import numpy
from . import Internal
from . import Converter as C


def _nbpts(z):
    d = Internal.getZoneDim(z)
    return d[1] * d[2] * d[3]


def _centre(z):
    x, y, zz = C.getCoords(z)
    return numpy.array([0.5 * (x.min() + x.max()), 0.5 * (y.min() + y.max()), 0.5 * (zz.min() + zz.max())])


def _rcb(idx, ctr, w, nparts, first, out):
    if nparts == 1:
        for i in idx:
            out[i] = first
        return
    nl = nparts // 2
    pts = ctr[idx]
    ax = int(numpy.argmax(pts.max(axis=0) - pts.min(axis=0)))
    order = sorted(idx, key=lambda i: (ctr[i][ax], i))
    tot = sum(w[i] for i in order)
    target = tot * nl / float(nparts)
    acc = 0.; cut = 0
    for k, i in enumerate(order):
        if acc + 0.5 * w[i] > target and k > 0:
            break
        acc += w[i]; cut = k + 1
    cut = min(max(cut, 1), len(order) - 1) if len(order) > 1 else 1
    _rcb(order[:cut], ctr, w, nl, first, out)
    _rcb(order[cut:], ctr, w, nparts - nl, first + nl, out)


def distribute_arrays(nbPts, NProc, centres=None, algorithm='fast', prescribed=None):
    n = len(nbPts)
    out = [-1] * n
    if prescribed:
        for i, p in prescribed.items():
            out[i] = p
    free = [i for i in range(n) if out[i] < 0]
    if algorithm == 'rcb' and centres is not None and len(free) >= NProc:
        _rcb(free, numpy.asarray(centres), nbPts, NProc, 0, out)
    else:
        load = [0.] * NProc
        for i in range(n):
            if out[i] >= 0: load[out[i]] += nbPts[i]
        for i in sorted(free, key=lambda i: (-nbPts[i], i)):
            p = int(numpy.argmin(load))
            out[i] = p; load[p] += nbPts[i]
    load = [0.] * NProc
    for i in range(n): load[out[i]] += nbPts[i]
    mean = sum(load) / NProc
    return {'distrib': out, 'varMin': (min(load) - mean) / mean if mean else 0.,
            'varMax': (max(load) - mean) / mean if mean else 0., 'nptsPerProc': load}


def _distribute(t, NProc, prescribed=None, algorithm='fast'):
    """Write .Solver#Param/proc on every zone (in place); returns stats like the reference's dict."""
    zones = Internal.getZones(t)
    pres = {}
    if prescribed:
        for i, z in enumerate(zones):
            if z[0] in prescribed: pres[i] = prescribed[z[0]]
    out = distribute_arrays([_nbpts(z) for z in zones], NProc, [_centre(z) for z in zones], algorithm, pres)
    for z, p in zip(zones, out['distrib']):
        _setProc(z, p)
    return out


def distribute(t, NProc, prescribed=None, algorithm='fast'):
    tp = Internal.copyRef(t)
    out = _distribute(tp, NProc, prescribed, algorithm)
    return tp, out


def _setProc(z, p):
    param = Internal.getNodeFromName1(z, '.Solver#Param')
    if param is None:
        param = ['.Solver#Param', None, [], 'UserDefinedData_t']
        z[2].append(param)
    v = numpy.zeros((1, 1), dtype=Internal.E_NpyInt); v[0, 0] = p
    node = Internal.getNodeFromName1(param, 'proc')
    if node is not None: node[1] = v
    else: param[2].append(['proc', v, [], 'DataArray_t'])


def getProc(z):
    param = Internal.getNodeFromName1(z, '.Solver#Param')
    if param is None: return -1
    node = Internal.getNodeFromName1(param, 'proc')
    return -1 if node is None else int(node[1].flat[0])


def getProcDict(t):
    """{zoneName: proc} (Distributor2/PyTree.py:268)."""
    return {z[0]: getProc(z) for z in Internal.getZones(t)}
