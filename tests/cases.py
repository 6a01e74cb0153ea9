"""Seeded synthetic overset cases shared by the CPU and GPU tests (shapes follow
SURVEY.md 8(d), scaled down so the oracle finishes in seconds)."""
import numpy as np
import cassiopee_b200.Generator as G
import cassiopee_b200.Converter as C


def flat(z):
    return [a.reshape(-1, order='F') for a in C.getCoords(z)]


def dims(z):
    from cassiopee_b200 import Internal
    d = Internal.getZoneDim(z)
    return d[1], d[2], d[3]


def c1(variant, n=31, m=21):
    """Two overlapping Cartesian zones, centres meshes: returns (donor tuple A, donor tuple B)
    as (ni,nj,nk,x,y,z)."""
    h = 1. / (n - 1)
    A = C.node2Center(G.cart((0, 0, 0), (h, h, h), (n, n, n)))
    o = int(0.3 * (n - 1))
    if variant == 'a':
        B = G.cart((o * h, o * h, o * h), (h, h, h), (m, m, m))
    elif variant == 'b':
        B = G.cart(((o + 0.5) * h, o * h, o * h), (h, h, h), (m, m, m))
    else:
        B = G.cart((0.3037, 0.2911, 0.3173), (0.9 * h, 0.9 * h, 0.9 * h), (m, m, m))
    B = C.node2Center(B)
    return (n - 1, n - 1, n - 1) + tuple(flat(A)), (m - 1, m - 1, m - 1) + tuple(flat(B))


def c2(nb=32, nc=(65, 17, 33)):
    """Cartesian background + cylinder O-grid (seam duplicated)."""
    Bg = G.cart((-3, -3, 0), (6. / (nb - 1), 6. / (nb - 1), 2. / (nb - 1)), (nb, nb, nb))
    Oc = G.cylinder((0, 0, 0), 0.5, 2.5, 360., 0., 2., nc)
    return (nb, nb, nb) + tuple(flat(Bg)), tuple(nc) + tuple(flat(Oc))


def fields(x, y, z, nvars=5, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for v in range(nvars):
        f = 1. + 0.1 * (v + 1) * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + 0.05 * (v + 1) * z
        out.append(f + 1e-3 * rng.standard_normal(x.size))
    return out


def circle_zone(name='circle', centre=(0.5, 0.5, 0.), R=0.05, N=20):
    """Structured curve zone (N,1,1) like Geom.circle((0.5,0.5,0), 0.05, N=20) of
    Connector/test/setInterpDataPT_t1.py (points at equal angles, first = last)."""
    import numpy as np
    import cassiopee_b200.Internal as Internal
    t = np.linspace(0., 2. * np.pi, N)
    z = Internal.newZone(name, N, 1, 1)
    gc = Internal.createChild(z, Internal.__GridCoordinates__, 'GridCoordinates_t')
    for nm, v in (('CoordinateX', centre[0] + R * np.cos(t)), ('CoordinateY', centre[1] + R * np.sin(t)),
                  ('CoordinateZ', np.full(N, centre[2]))):
        Internal.createChild(gc, nm, 'DataArray_t', np.asfortranarray(v.reshape((N, 1, 1))))
    return z


def check_curve_receptors(backend_kw, order, loc, storage, pen, nat):
    """One combination of the reference's setInterpDataPT_t1.py sweep: circle receptors in a 101 x 101 x 5 Cartesian
    donor; the subregion's lists must be the oracle's for the points the tree layer extracts."""
    import numpy as np
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.OversetData as X
    from oracle import oracle as O
    a = G.cart((0, 0, -0.2), (0.01, 0.01, 0.1), (101, 101, 5), name='cart')
    pts = circle_zone()
    tR = C.newPyTree(['R', pts]); tD = C.newPyTree(['D', a])
    C._initVars(tR, 'cellN', 2.); C._initVars(tR, 'centers:cellN', 2.)
    X._setInterpDataChimera(tR, tD, order=order, penalty=pen, nature=nat, loc=loc, storage=storage, verbose=0, **backend_kw)
    zr = Internal.getZones(tR)[0]; zd = Internal.getZones(tD)[0]
    holder, sname = (zr, 'ID_cart') if storage == 'direct' else (zd, 'ID_circle')
    s = Internal.getNodeFromName1(holder, sname)
    assert s is not None
    xr, yr, zz = C.getCoords(zr)
    if loc == 'centers':
        xr, yr, zz = [C.node2CenterArray(v) for v in (xr, yr, zz)]
    p = [np.ascontiguousarray(v.reshape(-1, order='F')) for v in (xr, yr, zz)]
    assert p[0].size == (20 if loc == 'nodes' else 19)
    xd, yd, zdd = [v.reshape(-1, order='F') for v in C.getCoords(zd)]
    ref = O.set_interp_data(p[0], p[1], p[2], [O.RefZone(101, 101, 5, xd, yd, zdd, np.ones(xd.size))],
                            order=order, nature=nat, penalty=pen, extrap=1)
    pl = Internal.getNodeFromName1(s, 'PointList')[1]; pld = Internal.getNodeFromName1(s, 'PointListDonor')[1]
    rcv, dnr = (pl, pld) if storage == 'direct' else (pld, pl)
    assert np.array_equal(rcv, ref[0][0]) and np.array_equal(dnr, ref[1][0])
    assert np.array_equal(Internal.getNodeFromName1(s, 'InterpolantsType')[1], ref[2][0])
    assert np.array_equal(Internal.getNodeFromName1(s, 'InterpolantsDonor')[1], ref[3][0])
    assert rcv.size == p[0].size
    return ref
