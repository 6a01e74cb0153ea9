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
