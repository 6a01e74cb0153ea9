"""TEST INFRASTRUCTURE: compute backend built on the oracle, injected into the
Python host layer so its tree/partition logic can be tested without a GPU.
The product only ever constructs cassiopee_b200._lib.DeviceBackend."""
import numpy as np


class OraclePlan:
    def __init__(self, imd, jmd, kmd, nptsR, donor, rcv, typ, coefs):
        self.a = (imd, jmd, donor, rcv, typ, coefs)
        self.n = len(typ)

    def transferD(self, fields):
        from oracle import oracle as O
        imd, jmd, donor, rcv, typ, coefs = self.a
        return O.transferD([np.ascontiguousarray(f.reshape(-1, order='F')) for f in fields], imd, jmd, donor, typ, coefs)


class OracleBackend:
    def make_zone(self, ni, nj, nk, x, y, z, cellN=None, interpDataType=1):
        from oracle import oracle as O
        return O.RefZone(ni, nj, nk, x, y, z, cellN, interpDataType=interpDataType)

    def make_zones(self, specs):
        return [self.make_zone(*sp[:7], interpDataType=sp[7]) for sp in specs]

    def make_zones_begin(self, specs):
        return self.make_zones(specs)

    def make_zones_finish(self, zones):
        pass

    def set_celln(self, hook, cellN):
        pass

    def set_interp_data(self, xr, yr, zr, hooks, order, nature, penalty, extrap):
        from oracle import oracle as O
        return O.set_interp_data(xr, yr, zr, hooks, order, nature, penalty, extrap)

    def select_receptors(self, ni, nj, nk, x, y, z, cellN, loc):
        """numpy restatement of getInterpolatedPoints (+ node2Center of the selected cells)."""
        import cassiopee_b200.Converter as C
        tcn = cellN - 2.
        sel = np.nonzero((tcn >= -1.e-15) & (tcn <= 1.e-15))[0]
        shp = (ni, nj, nk)
        if loc == 'nodes':
            pts = [np.ascontiguousarray(a[sel]) for a in (x, y, z)]
        else:
            pts = [C.node2CenterSelected(a.reshape(shp, order='F'), sel) for a in (x, y, z)]
        return OracleReceptors(sel.astype(np.int32), pts)

    def set_interp_data_rcv(self, rcv, hooks, order, nature, penalty, extrap):
        return self.set_interp_data(rcv.pts[0], rcv.pts[1], rcv.pts[2], hooks, order, nature, penalty, extrap)

    def submit_interp_data_rcv(self, rcv, hooks, order, nature, penalty, extrap):
        res = self.set_interp_data_rcv(rcv, hooks, order, nature, penalty, extrap)

        class _Done:
            def fetch(self_):
                return res
        return _Done()

    def make_plan(self, *a):
        return OraclePlan(*a)

    def apply_bc_overlaps(self, cellN, windows, depth, loc, val):
        from oracle import oracle as O
        for a, b in windows:
            O.apply_bc_overlap(cellN, a, b, depth, loc, val)

    def hole_fringe(self, cellN, depth, dir):
        """The reference's own searchMaskInterpolatedCellsStructOpt; depth < 0 as Connector.py:176-182 does it."""
        from oracle import oracle as O
        if depth < 0:
            cellN[...] = 1. - cellN + (cellN > 1.5) * 3.
            O.hole_fringe(cellN, -depth, dir)
            cellN[...] = 1. - cellN + (cellN > 1.5) * 3.
        else:
            O.hole_fringe(cellN, depth, dir)


class OracleReceptors:
    def __init__(self, indcell, pts):
        self.indcell = indcell; self.pts = pts; self.n = indcell.size
