"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of tests/hostemu/libhostemu.so (the
device header compiled for the host, see hostemu.cpp)."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostemu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "hostemu.cpp")
        hdr = os.path.join(_HERE, "..", "..", "cassiopee_b200", "csrc", "cx_core.cuh")
        if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                                   src, "-o", _SO])
        L = ctypes.CDLL(_SO)
        L.emu_zone_create.restype = ctypes.c_void_p
        L.emu_zone_create.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5
        L.emu_zone_create_cart.restype = ctypes.c_void_p
        L.emu_zone_create_cart.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4
        L.emu_cart_o3.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6
        L.emu_cart_o4.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6
        L.emu_zone_free.argtypes = [ctypes.c_void_p]
        L.emu_zone_depth.argtypes = [ctypes.c_void_p]
        L.emu_candidates.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4 + [ctypes.c_void_p, ctypes.c_int]
        L.emu_set_interp_data.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p] + \
            [ctypes.c_int] * 3 + [ctypes.c_void_p] * 6
        L.emu_cell_volume.restype = ctypes.c_double
        L.emu_cell_volume.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


def zone_bbox(ni, nj, nk, x, y, z):
    """Boundary-cell bbox (K_COMPGEOM::boundingBoxStruct) in numpy."""
    out = []
    m = np.zeros((nk, nj, ni), bool)
    for sl in (slice(0, 2), slice(-2, None)):
        m[sl, :, :] = True; m[:, sl, :] = True; m[:, :, sl] = True
    m = m.reshape(-1)
    lo = [a[m].min() for a in (x, y, z)]; hi = [a[m].max() for a in (x, y, z)]
    if nk == 1:                      # InterpAdt.cpp:293-301: a z = constant plane gets a thickness of 1
        assert abs(hi[2] - lo[2]) < 1e-8
        hi[2] = lo[2] + 1.
    return np.array(lo + hi)


class EmuZone:
    def __init__(self, ni, nj, nk, x, y, z, cellN=None, interpDataType=1):
        self.x = np.ascontiguousarray(x, np.float64).reshape(-1); self.y = np.ascontiguousarray(y, np.float64).reshape(-1)
        self.z = np.ascontiguousarray(z, np.float64).reshape(-1)
        self.c = np.ascontiguousarray(cellN, np.float64).reshape(-1) if cellN is not None else None
        if interpDataType == 0:
            self.h = lib().emu_zone_create_cart(ni, nj, nk, self.x.ctypes.data, self.y.ctypes.data, self.z.ctypes.data,
                                                self.c.ctypes.data if self.c is not None else None)
            return
        self.bb = zone_bbox(ni, nj, nk, self.x, self.y, self.z)
        self.h = lib().emu_zone_create(ni, nj, nk, self.x.ctypes.data, self.y.ctypes.data, self.z.ctypes.data,
                                       self.c.ctypes.data if self.c is not None else None, self.bb.ctypes.data)

    def depth(self):
        return lib().emu_zone_depth(self.h)

    def candidates(self, x, y, z, alphaTol=0.0, cap=1 << 16):
        out = np.zeros(cap, np.int32)
        n = lib().emu_candidates(self.h, x, y, z, alphaTol, out.ctypes.data, cap)
        return out[:min(n, cap)].copy()

    def volume(self, ind):
        return lib().emu_cell_volume(self.h, int(ind))

    def __del__(self):
        try: lib().emu_zone_free(self.h)
        except Exception: pass


def cart_o3(zone, xr, yr, zr, order=3):
    xr = np.ascontiguousarray(xr, np.float64).reshape(-1); yr = np.ascontiguousarray(yr, np.float64).reshape(-1)
    zr = np.ascontiguousarray(zr, np.float64).reshape(-1)
    n = xr.size
    found = np.zeros(n, np.int32); dind = np.zeros(n, np.int32); cf = np.zeros((n, 3 * order))
    (lib().emu_cart_o4 if order == 4 else lib().emu_cart_o3)(zone.h, n, xr.ctypes.data, yr.ctypes.data, zr.ctypes.data, found.ctypes.data, dind.ctypes.data,
                      cf.ctypes.data)
    return found, dind, cf


def set_interp_data_raw(xr, yr, zr, zones, nature=0, penalty=1, extrap=1):
    xr = np.ascontiguousarray(xr, np.float64).reshape(-1); yr = np.ascontiguousarray(yr, np.float64).reshape(-1)
    zr = np.ascontiguousarray(zr, np.float64).reshape(-1)
    n = xr.size
    hs = (ctypes.c_void_p * len(zones))(*[z.h for z in zones])
    noblk = np.zeros(n, np.int32); typ = np.zeros(n, np.int32); dind = np.zeros(n, np.int32)
    isext = np.zeros(n, np.int32); cf = np.zeros((n, 8)); st = np.zeros(4, np.int64)
    lib().emu_set_interp_data(n, xr.ctypes.data, yr.ctypes.data, zr.ctypes.data, len(zones), hs, nature, penalty,
                              extrap, noblk.ctypes.data, typ.ctypes.data, dind.ctypes.data, isext.ctypes.data,
                              cf.ctypes.data, st.ctypes.data)
    return dict(noblk=noblk, type=typ, dind=dind, isext=isext, cf=cf, stats=st)
