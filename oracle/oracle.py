"""ORACLE — TEST INFRASTRUCTURE ONLY.

Never imported by the product path (cassiopee_b200/).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
use it, and only as the checker / the CPU baseline.

ctypes front-end of oracle/_ref/libcassref.so = the reference's own K_INTERP
C++ compiled unchanged from /root/reference (recipe: oracle/Makefile) + the C
restatement of four Fortran routines (oracle/lagrange_restated.c) + a thin
driver (oracle/ref_driver.cpp).  Also holds the numpy restatement of the small
host-side pieces of the path:

* getInterpolatedPoints   Connector/Connector/getInterpolatedPoints.cpp:1675-1732
* node2centerStruct       KCore/KCore/Loc/node2Center.cpp:118-146
* packing per donor zone  Connector/Connector/setInterpData.cpp:1183-1252
* transfer (numpy form)   Connector/Connector/commonInterpTransfers.h:1-151

Parity status: order-2 search/coefficients = unmodified reference code, pinned
against SURVEY.md Appendix C (KAT-0) in tests/test_oracle.py; order 3/5 =
reference C++ + transliterated Fortran (flagged in DESIGN.md).
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libcassref.so")
_REFROOT = "/root/reference/Cassiopee"

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Build oracle/_ref/*.so when the reference tree is present (builder
    container only; the GPU box uses the prebuilt files)."""
    if os.path.exists(_LIB) and not force:
        return True
    if not os.path.isdir(_REFROOT):
        return os.path.exists(_LIB)
    subprocess.check_call(["make", "-s", "-j8", "-C", _HERE, "all"])
    return os.path.exists(_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        if not os.path.exists(_LIB):
            raise RuntimeError("oracle/_ref/libcassref.so missing and /root/reference absent")
        L = ctypes.CDLL(_LIB)
        L.ref_zone_create.restype = ctypes.c_void_p
        L.ref_zone_create.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4 + [ctypes.c_int]
        L.ref_zone_free.argtypes = [ctypes.c_void_p]
        L.ref_zone_bbox.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ref_zone_has_adt.argtypes = [ctypes.c_void_p]
        L.ref_candidates.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_int]
        L.ref_set_interp_data.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p] + \
            [ctypes.c_int] * 4 + [ctypes.c_void_p] * 6
        L.ref_transfer.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_int]
        L.ref_transferD.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ref_transfer_celln.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.ref_rotate.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p] + [ctypes.c_double] * 3 + \
            [ctypes.c_int] * 6
        L.ref_cell_volume.restype = ctypes.c_double
        L.ref_cell_volume.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


def num_threads():
    return lib().ref_num_threads()


def set_num_threads(n):
    lib().ref_set_num_threads(int(n))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32).reshape(-1)


class RefZone:
    """Donor zone held by the reference code: coordinates, cellN and its ADT
    (InterpAdt::buildStructAdt, KCore/KCore/Interp/InterpAdt.cpp:284-433)."""

    def __init__(self, ni, nj, nk, x, y, z, cellN=None, interpDataType=1):
        """interpDataType 1: InterpAdt; 0: InterpCart built like setInterpData.cpp:790-804."""
        self.ni, self.nj, self.nk = int(ni), int(nj), int(nk)
        self._x, self._y, self._z = _f64(x), _f64(y), _f64(z)
        n = self.ni * self.nj * self.nk
        assert self._x.size == n
        self._c = _f64(cellN) if cellN is not None else None
        self.h = lib().ref_zone_create(self.ni, self.nj, self.nk, self._x.ctypes.data, self._y.ctypes.data,
                                       self._z.ctypes.data, self._c.ctypes.data if self._c is not None else None,
                                       1 if interpDataType == 1 else 2)
        if not lib().ref_zone_has_adt(self.h):
            raise RuntimeError("reference ADT build failed")

    def bbox(self):
        bb = np.zeros(6)
        lib().ref_zone_bbox(self.h, bb.ctypes.data)
        return bb

    def candidates(self, x, y, z, alphaTol=0.0, cap=1 << 16):
        out = np.zeros(cap, dtype=np.int32)
        n = lib().ref_candidates(self.h, float(x), float(y), float(z), float(alphaTol), out.ctypes.data, cap)
        return out[:min(n, cap)].copy()

    def volume(self, indnode):
        return lib().ref_cell_volume(self.h, int(indnode))

    def free(self):
        if self.h:
            lib().ref_zone_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def set_interp_data_raw(xr, yr, zr, zones, order=2, nature=0, penalty=1, extrap=1):
    """Per-receptor results of the reference loop (setInterpData.cpp:978-1114).
    Returns dict(noblk,type,dind,isext,cf[n,16],vol)."""
    xr, yr, zr = _f64(xr), _f64(yr), _f64(zr)
    n = xr.size
    hs = (ctypes.c_void_p * len(zones))(*[z.h for z in zones])
    noblk = np.zeros(n, np.int32); typ = np.zeros(n, np.int32); dind = np.zeros(n, np.int32)
    isext = np.zeros(n, np.int32); cf = np.zeros((n, 16)); vol = np.zeros(n)
    rc = lib().ref_set_interp_data(n, xr.ctypes.data, yr.ctypes.data, zr.ctypes.data, len(zones), hs,
                                   order, nature, penalty, extrap, noblk.ctypes.data, typ.ctypes.data,
                                   dind.ctypes.data, isext.ctypes.data, cf.ctypes.data, vol.ctypes.data)
    if rc != 0:
        raise RuntimeError("ref_set_interp_data failed rc=%d" % rc)
    return dict(noblk=noblk, type=typ, dind=dind, isext=isext, cf=cf, vol=vol)


_NCF = {1: 1, 2: 8, 3: 9, 44: 12, 5: 15, 22: 4, 4: 8}


def pack_per_zone(raw, nz):
    """Re-create connector.setInterpData's return value
    [rcv]*nz,[donor]*nz,[type]*nz,[coefs]*nz,[extrap]*nz,orphan from the
    per-receptor table (setInterpData.cpp:1183-1272): entries per donor zone in
    ascending receptor position."""
    rcv, dnr, typ, cfs, ext = [], [], [], [], []
    for z in range(nz):
        sel = np.nonzero(raw["noblk"] == z + 1)[0]
        rcv.append(sel.astype(np.int32))
        dnr.append(raw["dind"][sel].astype(np.int32))
        t = raw["type"][sel].astype(np.int32)
        typ.append(t)
        ncf = np.array([_NCF[int(v)] for v in t], dtype=np.int64) if sel.size else np.zeros(0, np.int64)
        tot = int(ncf.sum())
        out = np.zeros(tot)
        if tot:
            off = np.concatenate(([0], np.cumsum(ncf)[:-1]))
            # ragged gather
            idx = np.repeat(np.arange(sel.size), ncf)
            k = np.arange(tot) - np.repeat(off, ncf)
            out[:] = raw["cf"][sel][idx, k]
        cfs.append(out)
        ext.append(sel[raw["isext"][sel] == 1].astype(np.int32))
    orphan = np.nonzero(raw["noblk"] == 0)[0].astype(np.int32)
    return [rcv, dnr, typ, cfs, ext, orphan, []]


def set_interp_data(xr, yr, zr, zones, order=2, nature=0, penalty=1, extrap=1):
    raw = set_interp_data_raw(xr, yr, zr, zones, order, nature, penalty, extrap)
    return pack_per_zone(raw, len(zones))


def set_interp_data_dw(pts, zones, order=2, nature=0, penalty=1):
    """connector.setInterpDataDW (setInterpData.cpp:1346-1890): pts[noz] = (x, y, z) of the receptors as donor zone noz
    sees them.  The reference's own getInterpolationCellDW / getExtrapolationCellDW per receptor (ref_set_interp_data_dw),
    then the reference's storage statement by statement (:1627-1737), including its slot rule for type 1: the donor
    index of a coincident entry goes to slot size+1 of an array initialised to -1 (:1655-1659), so it is overwritten by
    the next entry of the zone unless that one is coincident too."""
    nz = len(zones)
    assert len(pts) == nz
    arrs = [[_f64(a) for a in p] for p in pts]
    n = arrs[0][0].size
    hs = (ctypes.c_void_p * nz)(*[z.h for z in zones])
    noblk = np.zeros(n, np.int32); typ = np.zeros(n, np.int32); dind = np.zeros(n, np.int32)
    isext = np.zeros(n, np.int32); cf = np.zeros((n, 16))
    f = lib().ref_set_interp_data_dw
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5
    rc = f(n, _ptrs([a[0] for a in arrs]), _ptrs([a[1] for a in arrs]), _ptrs([a[2] for a in arrs]), nz, hs, order, nature,
           penalty, noblk.ctypes.data, typ.ctypes.data, dind.ctypes.data, isext.ctypes.data, cf.ctypes.data)
    if rc != 0:
        raise RuntimeError("ref_set_interp_data_dw failed rc=%d" % rc)
    rcvL = [[] for _ in range(nz)]; typL = [[] for _ in range(nz)]; cfL = [[] for _ in range(nz)]
    extL = [[] for _ in range(nz)]; orphan = []
    dnrA = [np.full(2 * n + 2, -1, np.int32) for _ in range(nz)]
    size = [0] * nz
    for ind in range(n):
        b = int(noblk[ind])
        if b == 0:
            orphan.append(ind); continue
        z = b - 1
        t = int(typ[ind])
        typL[z].append(t); rcvL[z].append(ind)
        cfL[z].extend(cf[ind, :_NCF[t]])
        if t == 1: dnrA[z][size[z] + 1] = dind[ind]
        else: dnrA[z][size[z]] = dind[ind]
        size[z] += 1
        if isext[ind] == 1: extL[z].append(ind)
    return [[np.array(r, np.int32) for r in rcvL], [dnrA[z][:size[z]].copy() for z in range(nz)],
            [np.array(t, np.int32) for t in typL], [np.array(c, np.float64) for c in cfL],
            [np.array(e, np.int32) for e in extL], np.array(orphan, np.int32), []]


def get_interpolated_points(cellN):
    """getInterpolatedPoints.cpp:1717-1724: positions with |cellN-2| <= 1e-15,
    source order."""
    c = np.asarray(cellN).reshape(-1)
    t = c - 2.0
    return np.nonzero((t >= -1e-15) & (t <= 1e-15))[0]


def node2center(x, ni, nj, nk):
    """K_LOC::node2centerStruct (KCore/KCore/Loc/node2Center.cpp:118-146), 3-D
    case: 0.125*(sum of the 8 nodes, left-assoc in the order
    ind0..ind7 = (i,j,k),(i+1,j,k),(i,j+1,k),(i+1,j+1,k),(..k+1))."""
    a = np.asarray(x, dtype=np.float64).reshape((nk, nj, ni))
    s = a[:-1, :-1, :-1] + a[:-1, :-1, 1:]
    s = s + a[:-1, 1:, :-1]
    s = s + a[:-1, 1:, 1:]
    s = s + a[1:, :-1, :-1]
    s = s + a[1:, :-1, 1:]
    s = s + a[1:, 1:, :-1]
    s = s + a[1:, 1:, 1:]
    return (0.125 * s).reshape(-1)


def _ptrs(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def transfer(fieldsD, fieldsR, imd, jmd, rcv, dnr, typ, coefs):
    """In-place reference transfer (commonInterpTransfers_indirect.h)."""
    rcv, dnr, typ = _i32(rcv), _i32(dnr), _i32(typ)
    coefs = _f64(coefs)
    for a in list(fieldsD) + list(fieldsR):
        assert a.dtype == np.float64 and a.flags.c_contiguous or a.flags.f_contiguous
    lib().ref_transfer(len(fieldsD), _ptrs(fieldsD), _ptrs(fieldsR), int(imd), int(jmd), rcv.size,
                       rcv.ctypes.data, dnr.ctypes.data, typ.ctypes.data, coefs.ctypes.data, coefs.size)


def transferD(fieldsD, imd, jmd, dnr, typ, coefs):
    """Reference _setInterpTransfersD arithmetic: dense (nvars,nrcv)."""
    dnr, typ = _i32(dnr), _i32(typ)
    coefs = _f64(coefs)
    out = np.zeros((len(fieldsD), typ.size))
    lib().ref_transferD(len(fieldsD), _ptrs(fieldsD), out.ctypes.data, int(imd), int(jmd), typ.size,
                        dnr.ctypes.data, typ.ctypes.data, coefs.ctypes.data, coefs.size)
    return out


def transfer_celln(cellND, cellNR, imd, jmd, rcv, dnr, typ, coefs):
    rcv, dnr, typ = _i32(rcv), _i32(dnr), _i32(typ)
    coefs = _f64(coefs)
    lib().ref_transfer_celln(cellND.ctypes.data, cellNR.ctypes.data, int(imd), int(jmd), rcv.size,
                             rcv.ctypes.data, dnr.ctypes.data, typ.ctypes.data, coefs.ctypes.data)


def hole_fringe(cellN, depth, dir=0):
    """The reference's searchMaskInterpolatedCellsStructOpt (getInterpolatedPoints.cpp:370-, compiled unchanged) on a
    Fortran-ordered (im,jm,km) cellN array, in place."""
    assert cellN.dtype == np.float64 and cellN.flags.f_contiguous
    im, jm, km = (list(cellN.shape) + [1, 1])[:3]
    f = lib().ref_hole_fringe
    f.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]
    f.restype = None
    f(im, jm, km, int(depth), int(dir), cellN.ctypes.data)


def apply_bc_overlap(cellN, minIndex, maxIndex, depth, loc, val=2):
    """connector.applyBCOverlapStruct (applyBCOverlaps.cpp:27-125, restated in oracle/ref_driver.cpp) on a
    Fortran-ordered (im,jm,km) cellN array, in place; raises TypeError for the windows the reference refuses."""
    assert cellN.dtype == np.float64 and cellN.flags.f_contiguous
    im, jm, km = (list(cellN.shape) + [1, 1])[:3]
    f = lib().ref_apply_bc_overlap
    f.argtypes = [ctypes.c_int] * 12 + [ctypes.c_void_p]
    f.restype = ctypes.c_int
    rc = f(im, jm, km, int(minIndex[0]), int(minIndex[1]), int(minIndex[2]), int(maxIndex[0]), int(maxIndex[1]),
           int(maxIndex[2]), int(depth), 0 if loc == 'nodes' else 1, int(val), cellN.ctypes.data)
    if rc != 0:
        raise TypeError("applyBCOverlaps: indices of structured window are not valid.")


def transfer_compact(roR, roD, zparams, dtloc, param_int, param_real, vartype=1, type_transfert=0, no_transfert=1,
                     nstep=1, threadmax_sdm=1):
    """The reference's ___setInterpTransfers reader (setInterpTransfers.cpp:761-1410) with its own arithmetic fragments,
    chimera subset (oracle/ref_driver.cpp:ref_transfer_compact).  roR[nd] / roD[nd]: compact solution blocks (numpy
    float64, nvars*NDIMDX values) of the receptor / donor zones, updated in place; zparams[nd]: the zones'
    Parameter_int (int32).  Returns the number of racs transferred."""
    n = len(roR)
    assert len(roD) == n and len(zparams) == n
    zp = [np.ascontiguousarray(z, dtype=np.int32) for z in zparams]
    R = (ctypes.c_void_p * n)(*[a.ctypes.data for a in roR])
    D = (ctypes.c_void_p * n)(*[a.ctypes.data for a in roD])
    Z = (ctypes.c_void_p * n)(*[a.ctypes.data for a in zp])
    dt = np.ascontiguousarray(dtloc, dtype=np.int32); pi = np.ascontiguousarray(param_int, dtype=np.int32)
    pr = np.ascontiguousarray(param_real, dtype=np.float64)
    f = lib().ref_transfer_compact
    f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.c_void_p] + [ctypes.c_int] * 5
    f.restype = ctypes.c_int
    rc = f(n, R, D, Z, dt.ctypes.data, pi.ctypes.data, pr.ctypes.data, int(vartype), int(type_transfert),
           int(no_transfert), int(nstep), int(threadmax_sdm))
    if rc < 0:
        raise RuntimeError("ref_transfer_compact: rac outside the chimera subset")
    return rc


def rotate(fieldsR, rcv, angles, posv=(-1, -1, -1), posm=(-1, -1, -1)):
    """Reference periodicity-by-rotation correction (includeTransfers.h) in place on the receptor fields."""
    rcv = _i32(rcv)
    lib().ref_rotate(len(fieldsR), _ptrs(fieldsR), rcv.size, rcv.ctypes.data, float(angles[0]), float(angles[1]),
                     float(angles[2]), *[int(p) for p in posv], *[int(p) for p in posm])


def transfer_numpy(fieldsD, fieldsR, imd, jmd, rcv, dnr, typ, coefs):
    """Independent numpy restatement of Appendix B / commonInterpTransfers.h
    (accumulation order kk outer, ii inner; products left to right; no FMA)."""
    rcv = np.asarray(rcv); dnr = np.asarray(dnr); typ = np.asarray(typ)
    ncf = np.array([_NCF[int(t)] for t in typ], dtype=np.int64)
    off = np.concatenate(([0], np.cumsum(ncf)))[:-1]
    imdjmd = imd * jmd
    for fD, fR in zip(fieldsD, fieldsR):
        fD = fD.reshape(-1, order="A"); fRv = fR.reshape(-1, order="A")
        for t, o in ((1, 0), (2, 2), (3, 3), (5, 5)):
            m = np.nonzero(typ == t)[0]
            if m.size == 0:
                continue
            d0 = dnr[m].astype(np.int64); c0 = off[m]
            if t == 1:
                fRv[rcv[m]] = coefs[c0] * fD[d0]
                continue
            val = np.zeros(m.size)
            for kk in range(o):
                for jj in range(o):
                    for ii in range(o):
                        f = fD[d0 + ii + jj * imd + kk * imdjmd]
                        if t == 2:
                            val = val + coefs[c0 + ii + 2 * jj + 4 * kk] * f
                        else:
                            val = val + coefs[c0 + ii] * coefs[c0 + jj + o] * coefs[c0 + kk + 2 * o] * f
            fRv[rcv[m]] = val
