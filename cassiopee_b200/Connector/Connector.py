"""Array-level Connector API on the device (mirror of
/root/reference/Cassiopee/Connector/Connector/Connector.py:425-446, :600-606).

A Converter "array" is ``[varString, ndarray(nfld, npts) float64 C-order, ni, nj, nk]``
(structured) — the object the reference's C entry points receive
(SURVEY.md 8(b)/(c)).
"""
import numpy
from .. import _lib


def _pos(varString, names):
    v = varString.split(',')
    for n in names:
        if n in v:
            return v.index(n)
    return -1


def _posxyz(varString):
    px = _pos(varString, ['x', 'CoordinateX'])
    py = _pos(varString, ['y', 'CoordinateY'])
    pz = _pos(varString, ['z', 'CoordinateZ'])
    return px, py, pz


def _poscelln(varString):
    return _pos(varString, ['cellN', 'cellnf', 'cellNF'])


def getInterpolatedPoints__(a):
    """connector.getInterpolatedPoints (getInterpolatedPoints.cpp:1675-1732):
    keep points with |cellN-2| <= 1e-15, in source order, append 'indcell' (the
    original index, stored as a double like the reference does)."""
    if isinstance(a[0], list):
        return [getInterpolatedPoints__(i) for i in a]
    posc = _poscelln(a[0])
    if posc == -1:
        raise TypeError("getInterpolatedPoints: array must contain cellN.")
    f = a[1]
    t = f[posc, :] - 2.
    sel = numpy.nonzero((t >= -1.e-15) & (t <= 1.e-15))[0]
    out = numpy.empty((f.shape[0] + 1, sel.size), dtype=numpy.float64)
    out[:-1, :] = f[:, sel]
    out[-1, :] = sel
    return [a[0] + ',indcell', out, sel.size, 1, 1]


def createHooks__(zonesD, ctx=None, interpDataType=1):
    """One device donor per array (analogue of [C.createHook(z,'adt') ...]); interpDataType 0 makes a
    regular Cartesian donor located in closed form (InterpCart) instead of an ADT."""
    ctx = ctx or _lib.default_context()
    hooks = []
    types = interpDataType if isinstance(interpDataType, list) else [interpDataType] * len(zonesD)
    for a, ty in zip(zonesD, types):
        px, py, pz = _posxyz(a[0])
        pc = _poscelln(a[0])
        if px == -1 or py == -1 or pz == -1:
            raise TypeError("setInterpData: 2nd argument is not valid.")
        f = a[1]
        hooks.append(_lib.Zone(ctx, a[2], a[3], a[4], numpy.ascontiguousarray(f[px]), numpy.ascontiguousarray(f[py]),
                               numpy.ascontiguousarray(f[pz]),
                               numpy.ascontiguousarray(f[pc]) if pc != -1 else None, interpDataType=ty))
    return hooks


def _dwDonorSlots(donor, typ):
    """The donor list as connector.setInterpDataDW stores it (setInterpData.cpp:1655-1659): a coincident (type 1) entry
    writes its donor index to the NEXT slot of an array initialised to -1, every other entry to its own slot, in entry
    order; the array is then cut to the number of entries.  Slot p therefore holds donor[p] unless entry p is of type 1,
    in which case it holds donor[p-1] if entry p-1 is of type 1 too, else -1."""
    m = donor.size
    out = numpy.full(m, -1, dtype=donor.dtype)
    t1 = typ == 1
    out[~t1] = donor[~t1]
    if m > 1:
        both = t1[1:] & t1[:-1]
        out[1:][both] = donor[:-1][both]
    return out


def _setInterpDataDW(interpPts, zonesD, order, nature, penalty, hook, ctx=None):
    if len(interpPts) != len(zonesD):
        raise TypeError("setInterpDataDW: 1st and 2nd arguments must be lists of the same size.")
    if order not in (2, 3, 5):
        print("Warning: setInterpDataDW: unknown interpolation order. Set to 2nd order.")
        order = 2
    ctx = ctx or _lib.default_context()
    pts = []
    for a in interpPts:
        px, py, pz = _posxyz(a[0])
        if px == -1 or py == -1 or pz == -1:
            raise TypeError("setInterpDataDW: 1st arg must contain coordinates.")
        if _pos(a[0], ['EXdir']) != -1:
            raise NotImplementedError("setInterpDataDW: EX points have no device path.")
        f = a[1]
        pts.append((numpy.ascontiguousarray(f[px]), numpy.ascontiguousarray(f[py]), numpy.ascontiguousarray(f[pz])))
    if hook is None:
        zones = createHooks__(zonesD, ctx, 1)
    else:
        if not isinstance(hook, list) or len(hook) != len(zonesD):
            raise TypeError("setInterpDataDW: hook must be a list of one hook per donor zone.")
        zones = hook
    res = _lib.set_interp_data_dw(ctx, pts, zones, order=order, nature=nature, penalty=penalty)
    out = res.fetch()
    out[1] = [_dwDonorSlots(d, t) for d, t in zip(out[1], out[2])]
    return out


def setInterpData__(interpPts, zonesD, order=2, penalty=1, extrap=1, nature=0, method='lagrangian',
                    interpDataType=1, hook=None, dim=3, ctx=None):
    """connector.setInterpData on the device.  Returns the reference's list
    [rcvInd1D]*nz, [donorInd1D]*nz, [donorType]*nz, [coefs]*nz, [extrap]*nz, orphan, [EXdir]
    (Connector.py:415-439, setInterpData.cpp:1272) or None without receptors.
    hook: None (device ADTs are built for this call) or a list of device
    zones from createHooks__, one per donor, same order."""
    if method != 'lagrangian':
        if method == 'leastsquares':
            raise NotImplementedError("setInterpData: method='leastsquares' has no device path.")
        raise ValueError("setInterpData__: %s: not a valid interpolation method." % method)
    if isinstance(interpPts[0], list):
        # list of arrays = double wall (Connector.py:427-429): array noz holds the receptor points as donor zone noz
        # must see them; connector.setInterpDataDW(interpPts, zonesD, order, nature, penalty, hook)
        if interpPts[0][1].shape[1] == 0:
            return None
        return _setInterpDataDW(interpPts, zonesD, order, nature, penalty, hook, ctx)
    if interpPts[1].shape[1] == 0:
        return None
    types = interpDataType if isinstance(interpDataType, list) else [interpDataType] * len(zonesD)
    for t in types:
        if t not in (0, 1):
            raise TypeError("setInterpData: InterpDataType must be 0 for CART or 1 for ADT.")
    if len(zonesD) == 0:
        raise TypeError("setInterpData: no valid donor zone found.")
    ctx = ctx or _lib.default_context()
    px, py, pz = _posxyz(interpPts[0])
    if px == -1 or py == -1 or pz == -1:
        raise TypeError("setInterpData: 1st arg must contain coordinates.")
    if hook is None:
        zones = createHooks__(zonesD, ctx, types)
    else:
        if not isinstance(hook, list):
            raise TypeError("setInterpData: hook must be a list of hooks.")
        if len(hook) != len(zonesD):
            raise TypeError("setInterpData: size of list of hooks must be equal to the number of donor zones.")
        zones = hook
    f = interpPts[1]
    res = _lib.set_interp_data(ctx, numpy.ascontiguousarray(f[px]), numpy.ascontiguousarray(f[py]),
                               numpy.ascontiguousarray(f[pz]), zones, order=order, nature=nature, penalty=penalty,
                               extrap=extrap)
    out = res.fetch()
    out.append(res.stats())   # extra 8th entry: device statistics (ignored by callers of the 7-list)
    return out
