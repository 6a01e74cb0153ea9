"""Synthetic structured grids used as inputs of the chimera path.

Only the two formulas BASELINE.json names are restated, as zones of a
CGNS/Python tree:

* cart      -> Generator/Generator/cart.cpp:72-84      x = xo + i*hi
* cylinder  -> Generator/Generator/cylinder.cpp:41-96  alpha = t1 + i*delta*(t2-t1),
               x = xo + (j*hR+R1)*cos(alpha), pi = 4*atan(1.)

The arrays are inputs to both the oracle and the CUDA path, so only the shape
and the formula matter (not bit-parity with libm's cos/sin).
"""
import math
import numpy
from . import Internal


def _zone_from_xyz(name, x, y, z, ni, nj, nk):
    zn = Internal.newZone(name, ni, nj, nk)
    gc = Internal.createChild(zn, Internal.__GridCoordinates__, 'GridCoordinates_t')
    for nm, a in (('CoordinateX', x), ('CoordinateY', y), ('CoordinateZ', z)):
        Internal.createChild(gc, nm, 'DataArray_t', numpy.asfortranarray(a.reshape((ni, nj, nk), order='F')))
    return zn


def cart(Xo, H, N, name='cart'):
    """Cartesian structured zone: Xo origin, H steps, N = (ni,nj,nk)."""
    ni, nj, nk = [int(v) for v in N]
    xo, yo, zo = [float(v) for v in Xo]
    hi, hj, hk = [float(v) for v in H]
    i = numpy.arange(ni, dtype=numpy.float64)
    j = numpy.arange(nj, dtype=numpy.float64)
    k = numpy.arange(nk, dtype=numpy.float64)
    x = numpy.empty((ni, nj, nk), order='F'); y = numpy.empty((ni, nj, nk), order='F')
    z = numpy.empty((ni, nj, nk), order='F')
    x[:, :, :] = (xo + i * hi)[:, None, None]
    y[:, :, :] = (yo + j * hj)[None, :, None]
    z[:, :, :] = (zo + k * hk)[None, None, :]
    return _zone_from_xyz(name, x, y, z, ni, nj, nk)


def cylinder(Xo, R1, R2, tetas, tetae, H, N, name='cylinder'):
    """Portion of cylinder, i angular, j radial, k axial (cylinder.cpp:41-96)."""
    ni, nj, nk = [int(v) for v in N]
    xo, yo, zo = [float(v) for v in Xo]
    pi = 4 * math.atan(1.)
    t1 = tetas * pi / 180.
    t2 = tetae * pi / 180.
    hk = 0. if nk == 1 else H / (nk - 1)
    hR = 0. if nj == 1 else (R2 - R1) / (nj - 1)
    delta = 0. if ni == 1 else 1. / (ni - 1.)
    dt2mt1 = delta * (t2 - t1)
    ca = numpy.array([math.cos(t1 + i * dt2mt1) for i in range(ni)])
    sa = numpy.array([math.sin(t1 + i * dt2mt1) for i in range(ni)])
    coef = numpy.arange(nj, dtype=numpy.float64) * hR + R1
    x = numpy.empty((ni, nj, nk), order='F'); y = numpy.empty((ni, nj, nk), order='F')
    z = numpy.empty((ni, nj, nk), order='F')
    x[:, :, :] = (xo + coef[None, :] * ca[:, None])[:, :, None]
    y[:, :, :] = (yo + coef[None, :] * sa[:, None])[:, :, None]
    z[:, :, :] = (zo + numpy.arange(nk, dtype=numpy.float64) * hk)[None, None, :]
    return _zone_from_xyz(name, x, y, z, ni, nj, nk)
