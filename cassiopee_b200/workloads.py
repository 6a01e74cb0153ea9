"""Synthetic overset workloads of BASELINE.json / SURVEY.md 8(d), as plain numpy
arrays (i fastest), shared by bench.py and the tests.  All grids follow the
reference generators' formulas (cassiopee_b200.Generator); seeds are fixed.

C1  two overlapping Cartesian zones (~1.0 M pts), three placement variants
C2  Cartesian background 256^3 + cylinder O-grid (513,129,257): ~17 M receptors
C4  4x4x4 blocks of 100^3 points with 8-cell overlap and per-block sub-cell
    shifts; receptors = outer 2 cell layers whose centre lies in another block
"""
import numpy as np
from . import Generator as G
from . import Converter as C


def _flat(z):
    return [np.ascontiguousarray(a.reshape(-1, order='F')) for a in C.getCoords(z)]


class Grid:
    """Structured zone as flat coordinate arrays."""

    def __init__(self, name, ni, nj, nk, x, y, z):
        self.name = name
        self.ni, self.nj, self.nk = int(ni), int(nj), int(nk)
        self.x, self.y, self.z = x, y, z

    @property
    def npts(self):
        return self.ni * self.nj * self.nk

    def bbox(self):
        return np.array([self.x.min(), self.y.min(), self.z.min(), self.x.max(), self.y.max(), self.z.max()])

    def centers(self):
        """Centre mesh (node2Center of the coordinates)."""
        out = []
        for a in (self.x, self.y, self.z):
            c = C.node2CenterArray(a.reshape((self.ni, self.nj, self.nk), order='F'))
            out.append(np.ascontiguousarray(c.reshape(-1, order='F')))
        return Grid(self.name, self.ni - 1, self.nj - 1, self.nk - 1, *out)


def cart(name, Xo, H, N):
    z = G.cart(Xo, H, N, name=name)
    return Grid(name, N[0], N[1], N[2], *_flat(z))


def cylinder(name, Xo, R1, R2, t1, t2, H, N):
    z = G.cylinder(Xo, R1, R2, t1, t2, H, N, name=name)
    return Grid(name, N[0], N[1], N[2], *_flat(z))


def c1(variant='c'):
    """A = cart 91^3, B = cart 63^3; returns node grids."""
    h = 1. / 90
    A = cart('A', (0, 0, 0), (h, h, h), (91, 91, 91))
    if variant == 'a':
        B = cart('B', (27 * h, 27 * h, 27 * h), (h, h, h), (63, 63, 63))
    elif variant == 'b':
        B = cart('B', (27.5 * h, 27 * h, 27 * h), (h, h, h), (63, 63, 63))
    else:
        B = cart('B', (0.3037, 0.2911, 0.3173), (0.9 * h, 0.9 * h, 0.9 * h), (63, 63, 63))
    return A, B


def c2(scale=1.0):
    """Background Bg = cart 256^3 over [-3,3]^2x[0,2]; near-body O-grid
    cylinder((0,0,0),0.5,2.5,360,0,2,(513,129,257)); scale<1 shrinks both."""
    nb = max(8, int(round(256 * scale)))
    nc = (max(9, int(round(512 * scale)) + 1), max(5, int(round(128 * scale)) + 1), max(5, int(round(256 * scale)) + 1))
    Bg = cart('Bg', (-3, -3, 0), (6. / (nb - 1), 6. / (nb - 1), 2. / (nb - 1)), (nb, nb, nb))
    Oc = cylinder('O', (0, 0, 0), 0.5, 2.5, 360., 0., 2., nc)
    return Bg, Oc


def conservative_fields(x, y, z, nvars=5, seed=0):
    """Analytic + seeded noise fields on donor nodes."""
    rng = np.random.default_rng(seed)
    out = []
    for v in range(nvars):
        f = 1. + 0.1 * (v + 1) * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + 0.05 * (v + 1) * z
        f += 1.e-3 * rng.standard_normal(x.size)
        out.append(np.ascontiguousarray(f))
    return out


VARS7 = ['Density', 'MomentumX', 'MomentumY', 'MomentumZ', 'EnergyStagnationDensity',
         'TurbulentSANuTildeDensity', 'Scalar']


def c4_layout(nblocks=(4, 4, 4), n=100, overlap=8, seed=1):
    """Block origins of the C4 cluster: block (a,b,c) = cart((a*L+sx, b*L+sy, c*L+sz), h, n^3),
    h=0.01, L=(n-overlap)*h, sub-cell shift s in [0,h)^3 from default_rng(seed)."""
    h = 0.01
    L = (n - overlap) * h
    rng = np.random.default_rng(seed)
    blocks = []
    for c in range(nblocks[2]):
        for b in range(nblocks[1]):
            for a in range(nblocks[0]):
                s = rng.uniform(0., h, 3)
                blocks.append(dict(name='blk_%d_%d_%d' % (a, b, c), abc=(a, b, c),
                                   origin=(a * L + s[0], b * L + s[1], c * L + s[2]), h=h, n=n))
    return blocks


def c4_block_grid(blk):
    n = blk['n']; h = blk['h']
    return cart(blk['name'], blk['origin'], (h, h, h), (n, n, n))


def c4_receptor_mask(blk, blocks, depth=2):
    """centers:cellN of one block: 2 on the outer `depth` cell layers whose
    centre lies inside another block's bounding box (what applyBCOverlaps
    depth=2 marks next to BCOverlap windows), 1 elsewhere.  Returns a flat
    (nc^3) array, i fastest."""
    n = blk['n']; h = blk['h']; nc = n - 1
    o = np.array(blk['origin'])
    idx = np.arange(nc)
    ctr = [o[d] + (idx + 0.5) * h for d in range(3)]
    edge = (idx < depth) | (idx >= nc - depth)
    outer = edge[:, None, None] | edge[None, :, None] | edge[None, None, :]
    inside = np.zeros((nc, nc, nc), bool)
    for ob in blocks:
        if ob is blk:
            continue
        oo = np.array(ob['origin']); ext = (ob['n'] - 1) * ob['h']
        m = [(ctr[d] > oo[d]) & (ctr[d] < oo[d] + ext) for d in range(3)]
        if not (m[0].any() and m[1].any() and m[2].any()):
            continue
        inside |= m[0][:, None, None] & m[1][None, :, None] & m[2][None, None, :]
    cellN = np.where(outer & inside, 2., 1.)
    return np.ascontiguousarray(cellN.reshape(-1, order='F'))


def bbox_intersect(b1, b2, tol=1.e-6):
    """K_COMPGEOM::compBoundingBoxIntersection semantics (compGeom.cpp:1055-1070)."""
    return not (b1[0] > b2[3] + tol or b1[3] < b2[0] - tol or b1[1] > b2[4] + tol or b1[4] < b2[1] - tol or
                b1[2] > b2[5] + tol or b1[5] < b2[2] - tol)


# ---------------------------------------------------------------------------
# C2 weak-scaled over GPUs: one C2a pair per body, backgrounds overlap their neighbours
# ---------------------------------------------------------------------------
def c2n_layout(nbodies, scale=1.0, overlap=8, seed=2):
    """N bodies in a row along x.  Body r = the C2 pair translated by cx_r: background
    Bg_r = cart((-3+cx_r,-3,0), (6/(nb-1), 6/(nb-1), 2/(nb-1)), nb^3) and near-body O-grid
    O_r = cylinder((cx_r,0,0), 0.5, 2.5, 360, 0, 2, nc).  Consecutive backgrounds overlap by
    `overlap` cells (+ a sub-cell shift from default_rng(seed), zero for body 0, so that one
    body alone is exactly C2a)."""
    nb = max(8, int(round(256 * scale)))
    nc = (max(9, int(round(512 * scale)) + 1), max(5, int(round(128 * scale)) + 1), max(5, int(round(256 * scale)) + 1))
    hx = 6. / (nb - 1)
    Lx = (nb - 1 - overlap) * hx
    rng = np.random.default_rng(seed)
    out = []
    for r in range(nbodies):
        sx = 0. if r == 0 else float(rng.uniform(0., hx))
        out.append(dict(r=r, n=nbodies, cx=r * Lx + sx, nb=nb, nc=nc, bg='Bg_%d' % r, og='O_%d' % r))
    return out


def c2n_zones(body, depth=2):
    """The two zones (CGNS/Python trees) of one body with node cellN: the O-grid is all receptors
    (cellN=2, like C2a); the background's `depth` node layers next to a neighbouring background are
    receptors (what applyBCOverlaps marks next to a BCOverlap window), 1 elsewhere."""
    nb = body['nb']; cx = body['cx']
    h = (6. / (nb - 1), 6. / (nb - 1), 2. / (nb - 1))
    bg = G.cart((-3. + cx, -3., 0.), h, (nb, nb, nb), name=body['bg'])
    og = G.cylinder((cx, 0., 0.), 0.5, 2.5, 360., 0., 2., body['nc'], name=body['og'])
    c = np.ones((nb, nb, nb), order='F')
    if body['r'] > 0:
        c[:depth, :, :] = 2.
    if body['r'] < body['n'] - 1:
        c[nb - depth:, :, :] = 2.
    C._initVars(bg, 'cellN', c)
    C._initVars(og, 'cellN', 2.)
    return bg, og
