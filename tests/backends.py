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

    def set_celln(self, hook, cellN):
        pass

    def set_interp_data(self, xr, yr, zr, hooks, order, nature, penalty, extrap):
        from oracle import oracle as O
        return O.set_interp_data(xr, yr, zr, hooks, order, nature, penalty, extrap)

    def make_plan(self, *a):
        return OraclePlan(*a)
