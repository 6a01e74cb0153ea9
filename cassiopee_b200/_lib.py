"""ctypes binding of csrc/libcassiopee_b200.so (C ABI: include/cassiopee_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device
can be opened, every compute entry point raises.  numpy host arrays stay owned
by Python; device mirrors are owned by the handle objects below.
"""
import ctypes
import math
import os
import weakref
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("CX_LIBPATH") or os.path.join(_HERE, "csrc", "libcassiopee_b200.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_ll = ctypes.c_longlong
_d = ctypes.c_double
_sz = ctypes.c_size_t

_SIGS = {
    "cx_last_error": (ctypes.c_char_p, []),
    "cx_version": (_i, []),
    "cx_device_count": (_i, [_vp]),
    "cx_ctx_create": (_i, [_i, _vp]),
    "cx_ctx_destroy": (_i, [_vp]),
    "cx_ctx_sync": (_i, [_vp]),
    "cx_ctx_launches": (_ll, [_vp]),
    "cx_ctx_info": (_i, [_vp, _vp, _vp, _vp]),
    "cx_event_record": (_i, [_vp, _i]),
    "cx_event_elapsed_ms": (_i, [_vp, _vp]),
    "cx_flush_l2": (_i, [_vp]),
    "cx_dev_alloc": (_i, [_vp, _sz, _vp]),
    "cx_dev_free": (_i, [_vp, _vp]),
    "cx_dev_upload": (_i, [_vp, _vp, _vp, _sz]),
    "cx_dev_download": (_i, [_vp, _vp, _vp, _sz]),
    "cx_dev_memset": (_i, [_vp, _vp, _i, _sz]),
    "cx_dev_upload_async": (_i, [_vp, _vp, _vp, _sz]),
    "cx_dev_download_async": (_i, [_vp, _vp, _vp, _sz]),
    "cx_host_register": (_i, [_vp, _sz]),
    "cx_host_unregister": (_i, [_vp]),
    "cx_pinned_alloc": (_i, [_sz, _vp]),
    "cx_pinned_free": (_i, [_vp]),
    "cx_zone_create": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cx_zone_create_noadt": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cx_zones_build_adt": (_i, [_vp, _i, _vp]),
    "cx_zone_adt_begin": (_i, [_vp]),
    "cx_zone_create_cart": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cx_zone_set_celln": (_i, [_vp, _vp]),
    "cx_zone_info": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "cx_zone_candidates": (_i, [_vp, _d, _d, _d, _d, _vp, _i, _vp]),
    "cx_zone_export_adt": (_i, [_vp, _vp, _sz]),
    "cx_zone_destroy": (_i, [_vp]),
    "cx_set_interp_data": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "cx_set_interp_data_dev": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "cx_receptors_create": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "cx_receptors_from_zone": (_i, [_vp, _vp, _vp]),
    "cx_receptors_count": (_i, [_vp, _vp]),
    "cx_receptors_fetch": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "cx_receptors_destroy": (_i, [_vp]),
    "cx_set_interp_data_rcv": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "cx_ctx_search_slot": (_i, [_vp, _i]),
    "cx_set_interp_data_dw": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "cx_result_sizes": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "cx_result_fetch_zone": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "cx_result_fetch_orphans": (_i, [_vp, _vp]),
    "cx_result_fetch_all": (_i, [_vp] * 7),
    "cx_result_fetch_raw": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "cx_result_stats": (_i, [_vp, _vp]),
    "cx_result_destroy": (_i, [_vp]),
    "cx_plan_create": (_i, [_vp, _i, _i, _i, _ll, _i, _vp, _vp, _vp, _vp, _ll, _vp]),
    "cx_transfer": (_i, [_vp, _i, _vp, _vp]),
    "cx_transferD": (_i, [_vp, _i, _vp, _vp]),
    "cx_transfer_rot": (_i, [_vp, _i, _vp, _vp, _i, _d, _d, _i, _vp]),
    "cx_batch_begin": (_i, [_vp, _ll]),
    "cx_batch_flush": (_i, [_vp]),
    "cx_batch_end": (_i, [_vp, _vp]),
    "cx_rotate_dev": (_i, [_vp, _i, _d, _d, _vp, _vp, _vp]),
    "cx_transfer_celln": (_i, [_vp, _vp, _vp]),
    "cx_transfer_dev": (_i, [_vp, _i, _vp, _vp]),
    "cx_transferD_dev": (_i, [_vp, _i, _vp, _vp, _ll]),
    "cx_transfer_celln_dev": (_i, [_vp, _vp, _vp]),
    "cx_scatter_dev": (_i, [_vp, _i, _vp, _ll, _i, _vp, _vp]),
    "cx_gather_dev": (_i, [_vp, _i, _vp, _ll, _i, _vp, _vp]),
    "cx_apply_bc_overlaps": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, ctypes.c_double, _vp]),
    "cx_apply_bc_overlap_dev": (_i, [_vp] + [_i] * 11 + [ctypes.c_double, _vp]),
    "cx_hole_fringe": (_i, [_vp] + [_i] * 5 + [_vp]),
    "cx_hole_fringe_dev": (_i, [_vp] + [_i] * 5 + [_vp, _vp]),
    "cx_host_scatter": (_i, [_vp, _vp, _vp, _ll]),
    "cx_plan_info": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "cx_plan_upload_points": (_i, [_vp, _vp]),
    "cx_plan_sorted_rcv": (_i, [_vp, _vp]),
    "cx_ipc_export": (_i, [_vp, _vp]),
    "cx_ipc_open": (_i, [_vp, _vp, _vp]),
    "cx_ipc_close": (_i, [_vp]),
    "cx_ipc_close_all": (_i, []),
    "cx_plan_destroy": (_i, [_vp]),
    "cx_planset_create": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cx_planset_run": (_i, [_vp]),
    "cx_planset_destroy": (_i, [_vp]),
    "cx_scatterset_create": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cx_scatterset_run": (_i, [_vp]),
    "cx_scatterset_destroy": (_i, [_vp]),
    "cx_step_create": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cx_step_create_p2p": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "cx_step_run": (_i, [_vp, _i, _i]),
    "cx_step_run_vars": (_i, [_vp, _i, _i, _i, _i]),
    "cx_step_splittable": (_i, [_vp]),
    "cx_hostio_create": (_i, [_vp, _i, _vp]),
    "cx_hostio_add": (_i, [_vp, _i, _ll, _vp, _vp, _vp]),
    "cx_hostio_add_box": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "cx_hostio_bytes": (_i, [_vp, _vp, _vp]),
    "cx_step_run_host": (_i, [_vp, _vp, _i]),
    "cx_hostio_destroy": (_i, [_vp]),
    "cx_plan_mark_donors": (_i, [_vp, _vp]),
    "cx_host_gather": (_i, [_vp, _vp, _vp, _ll]),
    "cx_step_destroy": (_i, [_vp]),
    "cx_comm_unique_id": (_i, [_vp]),
    "cx_comm_init": (_i, [_vp, _i, _i, _vp]),
    "cx_comm_destroy": (_i, [_vp]),
    "cx_comm_barrier": (_i, [_vp]),
    "cx_exchange_dev": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "cx_exchange_bytes_dev": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "cx_zone_meta": (_i, [_vp, _vp, _vp]),
    "cx_zone_exchange": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
}

EXPORTED_SYMBOLS = sorted(_SIGS.keys())


class CxError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library (loudly: no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise CxError("cassiopee_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIBPATH)
        L = ctypes.CDLL(LIBPATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().cx_last_error().decode(errors="replace")
        if rc == -4:
            raise NotImplementedError(msg)
        if rc == -1:
            raise TypeError(msg)
        raise CxError("cassiopee_b200 (status %d): %s" % (rc, msg))


def f64(a):
    a = np.asarray(a)
    if a.dtype != np.float64:
        raise TypeError("cassiopee_b200: arrays must be float64")
    if not (a.flags.c_contiguous or a.flags.f_contiguous):
        raise TypeError("cassiopee_b200: arrays must be C- or F-contiguous (no copy is made)")
    return a


def i32(a):
    a = np.asarray(a)
    if a.dtype != np.int32 or not (a.flags.c_contiguous or a.flags.f_contiguous):
        a = np.ascontiguousarray(a, dtype=np.int32)
    return a


def ptr(a):
    return a.ctypes.data if a is not None else None


def ptr_array(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


class Context:
    """One CUDA device + stream (cx_ctx)."""

    def __init__(self, device=0):
        h = _vp()
        check(lib().cx_ctx_create(int(device), ctypes.byref(h)))
        self.h = h
        self.device = int(device)
        self._fin = weakref.finalize(self, lib().cx_ctx_destroy, h)

    def sync(self):
        check(lib().cx_ctx_sync(self.h))

    def launches(self):
        return int(lib().cx_ctx_launches(self.h))

    def info(self):
        sm = _i(); fr = _sz(); tot = _sz()
        check(lib().cx_ctx_info(self.h, ctypes.byref(sm), ctypes.byref(fr), ctypes.byref(tot)))
        return dict(sm_count=sm.value, free_bytes=fr.value, total_bytes=tot.value)

    def event_record(self, which):
        check(lib().cx_event_record(self.h, int(which)))

    def event_elapsed_ms(self):
        ms = ctypes.c_float()
        check(lib().cx_event_elapsed_ms(self.h, ctypes.byref(ms)))
        return float(ms.value)

    def batch_begin(self, budget_bytes=0):
        """Open a batch of host-pointer transfers: donor fields are uploaded once per host array (cx_batch_begin)."""
        check(lib().cx_batch_begin(self.h, int(budget_bytes)))

    def batch_flush(self):
        check(lib().cx_batch_flush(self.h))

    def batch_end(self):
        n = _ll()
        check(lib().cx_batch_end(self.h, ctypes.byref(n)))
        return n.value

    def flush_l2(self):
        check(lib().cx_flush_l2(self.h))

    # raw device arrays of doubles / ints
    def alloc(self, nbytes):
        p = _vp()
        check(lib().cx_dev_alloc(self.h, int(nbytes), ctypes.byref(p)))
        return p.value

    def free(self, dptr):
        if dptr:
            lib().cx_dev_free(self.h, _vp(dptr))

    def upload(self, dptr, host):
        check(lib().cx_dev_upload(self.h, _vp(dptr), ptr(host), host.nbytes))

    def download(self, host, dptr):
        check(lib().cx_dev_download(self.h, ptr(host), _vp(dptr), host.nbytes))

    def upload_async(self, dptr, host):
        """No synchronisation (order with sync()); host should be pinned or registered."""
        check(lib().cx_dev_upload_async(self.h, _vp(dptr), ptr(host), host.nbytes))

    def download_async(self, host, dptr):
        check(lib().cx_dev_download_async(self.h, ptr(host), _vp(dptr), host.nbytes))


_default_ctx = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get("CX_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        n = _i()
        check(lib().cx_device_count(ctypes.byref(n)))
        if n.value > 0:
            device = device % n.value
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class DeviceArray:
    """float64/int32 array resident in HBM (cx_dev_alloc)."""

    def __init__(self, ctx, n, dtype=np.float64):
        self.ctx = ctx
        self.n = int(n)
        self.dtype = np.dtype(dtype)
        self.ptr = ctx.alloc(self.n * self.dtype.itemsize)
        self._fin = weakref.finalize(self, lib().cx_dev_free, ctx.h, _vp(self.ptr))

    @classmethod
    def from_host(cls, ctx, a):
        a = np.asarray(a)
        d = cls(ctx, a.size, a.dtype)
        d.upload(a)
        return d

    def upload(self, a):
        a = np.asarray(a)
        assert a.size == self.n and a.dtype == self.dtype
        if not (a.flags.c_contiguous or a.flags.f_contiguous):
            a = np.ascontiguousarray(a)
        self.ctx.upload(self.ptr, a)

    def download(self, out=None):
        if out is None:
            out = np.empty(self.n, dtype=self.dtype)
        self.ctx.download(out, self.ptr)
        return out


class Zone:
    """Donor zone resident on the device with its ADT replica (cx_zone) — the
    analogue of the reference's ``C.createHook(z, 'adt')``."""

    def __init__(self, ctx, ni, nj, nk, x, y, z, cellN=None, interpDataType=1, defer_adt=False):
        """interpDataType 1: ADT replica (InterpAdt); 0: regular Cartesian donor located in closed form
        (InterpCart, setInterpData.cpp:790-804) — no tree is built.  defer_adt: the tree is built later, with those of
        the other zones of a batch (build_adts)."""
        self.ctx = ctx
        self.ni, self.nj, self.nk = int(ni), int(nj), int(nk)
        self.interpDataType = int(interpDataType)
        x, y, z = f64(x), f64(y), f64(z)
        n = self.ni * self.nj * self.nk
        if x.size != n or y.size != n or z.size != n:
            raise TypeError("setInterpData: 2nd argument is not valid.")
        if self.interpDataType not in (0, 1):
            raise TypeError("setInterpData: InterpDataType must be 0 for CART or 1 for ADT.")
        c = f64(cellN) if cellN is not None else None
        h = _vp()
        create = lib().cx_zone_create_cart if self.interpDataType == 0 else (
            lib().cx_zone_create_noadt if defer_adt else lib().cx_zone_create)
        check(create(ctx.h, self.ni, self.nj, self.nk, ptr(x), ptr(y), ptr(z), ptr(c), ctypes.byref(h)))
        self.h = h
        self._fin = weakref.finalize(self, lib().cx_zone_destroy, h)

    @classmethod
    def _from_handle(cls, ctx, h, dims, topo):
        """Wrap a zone handle created by the library (cx_zone_exchange)."""
        z = cls.__new__(cls)
        z.ctx = ctx
        z.ni, z.nj, z.nk = [int(d) for d in dims]
        z.interpDataType = 1 if topo == 1 else 0
        z.h = h
        z._fin = weakref.finalize(z, lib().cx_zone_destroy, h)
        return z

    def meta(self):
        """(ints[6], doubles[9]) a peer needs to adopt this zone's device data (cx_zone_meta)."""
        mi = np.zeros(6, np.int32); md = np.zeros(9)
        check(lib().cx_zone_meta(self.h, ptr(mi), ptr(md)))
        return mi, md

    def set_celln(self, cellN):
        c = f64(cellN) if cellN is not None else None
        check(lib().cx_zone_set_celln(self.h, ptr(c)))

    def info(self):
        bb = np.zeros(6); nc = _i(); dp = _i(); ms = ctypes.c_float()
        check(lib().cx_zone_info(self.h, ptr(bb), ctypes.byref(nc), ctypes.byref(dp), ctypes.byref(ms)))
        return dict(bbox=bb, ncells=nc.value, depth=dp.value, build_ms=float(ms.value))

    def candidates(self, x, y, z, alphaTol=0.0, cap=1 << 16):
        out = np.zeros(cap, np.int32); n = _i()
        check(lib().cx_zone_candidates(self.h, float(x), float(y), float(z), float(alphaTol), ptr(out), cap,
                                       ctypes.byref(n)))
        return out[:min(n.value, cap)].copy()

    def export_adt(self):
        nc = self.info()["ncells"]
        dt = np.dtype([("bb", np.float64, 6), ("sub", np.float64, 6), ("mid", np.float64), ("rlo", np.float64),
                       ("left", np.int32), ("right", np.int32), ("dim", np.int32), ("ind", np.int32)])
        assert dt.itemsize == 128
        out = np.zeros(nc, dtype=dt)
        check(lib().cx_zone_export_adt(self.h, ptr(out), out.nbytes))
        return out


def build_adts(ctx, zones):
    """ADTs of the zones created with defer_adt=True, built concurrently on the device (cx_zones_build_adt)."""
    if not zones:
        return
    hs = (ctypes.c_void_p * len(zones))(*[z.h.value for z in zones])
    check(lib().cx_zones_build_adt(ctx.h, len(zones), hs))


def zone_exchange(ctx, sends, recvs):
    """Device-to-device replication of donor zones over NCCL (cx_zone_exchange).
    sends: list of (Zone, peer); recvs: list of (peer, meta_i, meta_d) with the owner's Zone.meta().
    Returns the received Zone objects in the order of recvs."""
    ns, nr = len(sends), len(recvs)
    sz = (ctypes.c_void_p * max(ns, 1))(*[z.h.value for z, _ in sends])
    sp = (ctypes.c_int * max(ns, 1))(*[int(p) for _, p in sends])
    rp = (ctypes.c_int * max(nr, 1))(*[int(p) for p, _, _ in recvs])
    mi = np.ascontiguousarray(np.concatenate([np.asarray(m, np.int32) for _, m, _ in recvs]) if nr else np.zeros(6, np.int32))
    md = np.ascontiguousarray(np.concatenate([np.asarray(m, np.float64) for _, _, m in recvs]) if nr else np.zeros(9))
    out = (ctypes.c_void_p * max(nr, 1))()
    check(lib().cx_zone_exchange(ctx.h, ns, sz, sp, nr, rp, ptr(mi), ptr(md), out))
    zones = []
    for i, (p, m_i, m_d) in enumerate(recvs):
        zones.append(Zone._from_handle(ctx, _vp(out[i]), m_i[:3], int(m_i[5])))
    return zones


class InterpResult:
    """Output of one setInterpData call (cx_result)."""

    def __init__(self, h, nz, nrcv):
        self.h = h
        self.nz = nz
        self.nrcv = nrcv
        self._fin = weakref.finalize(self, lib().cx_result_destroy, h)

    def sizes(self):
        cnt = np.zeros(self.nz, np.int32); nco = np.zeros(self.nz, np.int32); nex = np.zeros(self.nz, np.int32)
        no = _i()
        check(lib().cx_result_sizes(self.h, ptr(cnt), ptr(nco), ptr(nex), ctypes.byref(no)))
        return cnt, nco, nex, no.value

    def fetch(self, pinned=False):
        """The reference's 7-list [rcv]*nz,[donor]*nz,[type]*nz,[coefs]*nz,[extrap]*nz,orphan,[]
        (Connector/Connector/setInterpData.cpp:1272).  pinned=True returns numpy arrays backed by
        page-locked memory from the reuse pool (fast D2H, no first-touch page faults)."""
        cnt, nco, nex, no = self.sizes()
        big = 1 << 16
        mk = (lambda n, dt: pinned_empty(n, dt) if (pinned and n >= big) else np.empty(n, dt))
        ne, nc, nx = int(cnt.sum()), int(nco.sum()), int(nex.sum())
        # one block per kind of list for the whole call (six copies instead of five per donor zone); the per-zone
        # lists are consecutive slices of them
        r = mk(ne, np.int32); d = mk(ne, np.int32); t = mk(ne, np.int32); c = mk(nc, np.float64)
        e = np.empty(nx, np.int32); orphan = np.empty(no, np.int32)
        check(lib().cx_result_fetch_all(self.h, ptr(r), ptr(d), ptr(t), ptr(c), ptr(e), ptr(orphan)))
        eo = np.concatenate(([0], np.cumsum(cnt))); co = np.concatenate(([0], np.cumsum(nco)))
        xo = np.concatenate(([0], np.cumsum(nex)))
        R = [r[eo[z]:eo[z + 1]] for z in range(self.nz)]
        D = [d[eo[z]:eo[z + 1]] for z in range(self.nz)]
        T = [t[eo[z]:eo[z + 1]] for z in range(self.nz)]
        C = [c[co[z]:co[z + 1]] for z in range(self.nz)]
        E = [e[xo[z]:xo[z + 1]] for z in range(self.nz)]
        return [R, D, T, C, E, orphan, []]

    def raw(self):
        st = self.stats()
        ncf = int(st["ncfmax"])
        n = self.nrcv
        noblk = np.zeros(n, np.int32); typ = np.zeros(n, np.int32); dind = np.zeros(n, np.int32)
        isext = np.zeros(n, np.int32); cf = np.zeros((ncf, n))
        check(lib().cx_result_fetch_raw(self.h, ptr(noblk), ptr(typ), ptr(dind), ptr(isext), ptr(cf)))
        return dict(noblk=noblk, type=typ, dind=dind, isext=isext, cf=cf.T.copy())

    def stats(self):
        out = np.zeros(8)
        check(lib().cx_result_stats(self.h, ptr(out)))
        return dict(search_ms=out[0], extrap_ms=out[1], nodes_visited=out[2], candidates=out[3], ncfmax=out[4])


def set_interp_data(ctx, xr, yr, zr, zones, order=2, nature=0, penalty=1, extrap=1):
    """connector.setInterpData on host receptor coordinates (numpy)."""
    xr, yr, zr = f64(xr).reshape(-1), f64(yr).reshape(-1), f64(zr).reshape(-1)
    n = xr.size
    hs = (ctypes.c_void_p * len(zones))(*[z.h.value for z in zones])
    h = _vp()
    check(lib().cx_set_interp_data(ctx.h, n, ptr(xr), ptr(yr), ptr(zr), len(zones), hs, int(order), int(nature),
                                   int(penalty), int(extrap), ctypes.byref(h)))
    return InterpResult(h, len(zones), n)


def set_interp_data_dw(ctx, pts, zones, order=2, nature=0, penalty=1):
    """connector.setInterpDataDW: pts[noz] = (x, y, z) host arrays of the receptors as donor zone noz sees them."""
    if len(pts) != len(zones):
        raise TypeError("setInterpDataDW: one receptor array per donor zone is expected.")
    arrs = [[f64(a).reshape(-1) for a in p] for p in pts]
    n = arrs[0][0].size
    for p in arrs:
        if any(a.size != n for a in p):
            raise TypeError("setInterpDataDW: receptor arrays must have the same number of points.")
    hs = (ctypes.c_void_p * len(zones))(*[z.h.value for z in zones])
    h = _vp()
    check(lib().cx_set_interp_data_dw(ctx.h, n, ptr_array([ptr(p[0]) for p in arrs]), ptr_array([ptr(p[1]) for p in arrs]),
                                      ptr_array([ptr(p[2]) for p in arrs]), len(zones), hs, int(order), int(nature),
                                      int(penalty), ctypes.byref(h)))
    res = InterpResult(h, len(zones), n)
    res._keep = arrs
    return res


class Receptors:
    """Receptor points of one zone extracted on the device (cx_receptors): the points with cellN == 2 at `loc`,
    their index in the zone ('indcell') and their coordinates (cell centres for loc='centers')."""

    def __init__(self, ctx, ni, nj, nk, x, y, z, cellN, loc):
        if loc not in ('nodes', 'centers'):
            raise ValueError("setInterpData: invalid loc.")
        self.ctx = ctx
        x, y, z, cellN = f64(x), f64(y), f64(z), f64(cellN)
        nn = int(ni) * int(nj) * int(nk)
        npts = nn if loc == 'nodes' else max(int(ni) - 1, 1) * max(int(nj) - 1, 1) * max(int(nk) - 1, 1)
        if x.size != nn or y.size != nn or z.size != nn or cellN.size != npts:
            raise TypeError("getInterpolatedPoints: coordinates / cellN do not match the zone dimensions.")
        h = _vp()
        check(lib().cx_receptors_create(ctx.h, int(ni), int(nj), int(nk), ptr(x), ptr(y), ptr(z), ptr(cellN),
                                        0 if loc == 'nodes' else 1, ctypes.byref(h)))
        self.h = h
        self._fin = weakref.finalize(self, lib().cx_receptors_destroy, h)
        n = _i()
        check(lib().cx_receptors_count(h, ctypes.byref(n)))
        self.n = n.value
        self._ind = None

    @classmethod
    def from_zone(cls, ctx, zone):
        """Receptor points (node cellN == 2) of a zone that is resident as a donor hook: nothing is uploaded."""
        r = cls.__new__(cls)
        r.ctx = ctx
        h = _vp()
        check(lib().cx_receptors_from_zone(ctx.h, zone.h, ctypes.byref(h)))
        r.h = h
        r._fin = weakref.finalize(r, lib().cx_receptors_destroy, h)
        n = _i()
        check(lib().cx_receptors_count(h, ctypes.byref(n)))
        r.n = n.value
        r._ind = None
        return r

    @property
    def indcell(self):
        if self._ind is None:
            self._ind = np.empty(self.n, np.int32)
            check(lib().cx_receptors_fetch(self.h, ptr(self._ind), None, None, None))
        return self._ind

    def coords(self):
        x = np.empty(self.n); y = np.empty(self.n); z = np.empty(self.n)
        check(lib().cx_receptors_fetch(self.h, None, ptr(x), ptr(y), ptr(z)))
        return x, y, z


def set_interp_data_rcv(ctx, rcv, zones, order=2, nature=0, penalty=1, extrap=1):
    hs = (ctypes.c_void_p * len(zones))(*[z.h.value for z in zones])
    h = _vp()
    check(lib().cx_set_interp_data_rcv(ctx.h, rcv.h, len(zones), hs, int(order), int(nature), int(penalty),
                                       int(extrap), ctypes.byref(h)))
    return InterpResult(h, len(zones), rcv.n)


def set_interp_data_dev(ctx, n, dxr, dyr, dzr, zones, order=2, nature=0, penalty=1, extrap=1):
    hs = (ctypes.c_void_p * len(zones))(*[z.h.value for z in zones])
    h = _vp()
    check(lib().cx_set_interp_data_dev(ctx.h, int(n), _vp(dxr), _vp(dyr), _vp(dzr), len(zones), hs, int(order),
                                       int(nature), int(penalty), int(extrap), ctypes.byref(h)))
    return InterpResult(h, len(zones), int(n))


class Plan:
    """Device-resident interpData of one ID_* subregion (cx_plan)."""

    def __init__(self, ctx, imd, jmd, kmd, nptsR, donor, rcv, typ, coefs):
        self.ctx = ctx
        donor, rcv, typ = i32(donor).reshape(-1), i32(rcv).reshape(-1), i32(typ).reshape(-1)
        coefs = f64(coefs).reshape(-1)
        if not (donor.size == rcv.size == typ.size):
            raise TypeError("_setInterpTransfers: PointList, PointListDonor and InterpolantsType differ in size")
        self.n = typ.size
        self.imd, self.jmd, self.kmd, self.nptsR = int(imd), int(jmd), int(kmd), int(nptsR)
        self.nptsD = self.imd * self.jmd * self.kmd
        h = _vp()
        check(lib().cx_plan_create(ctx.h, self.imd, self.jmd, self.kmd, self.nptsR, self.n, ptr(donor), ptr(rcv),
                                   ptr(typ), ptr(coefs), coefs.size, ctypes.byref(h)))
        self.h = h
        self._fin = weakref.finalize(self, lib().cx_plan_destroy, h)

    def _check_fields(self, fields, npts, what):
        out = []
        for a in fields:
            a = f64(a)
            if a.size != npts:
                raise TypeError("_setInterpTransfers: %s field has %d values, zone has %d" % (what, a.size, npts))
            out.append(a)
        return out

    def transfer(self, fieldsD, fieldsR):
        """In place on host numpy arrays (connector._setInterpTransfers)."""
        fD = self._check_fields(fieldsD, self.nptsD, "donor")
        fR = self._check_fields(fieldsR, self.nptsR, "receptor")
        assert len(fD) == len(fR)
        check(lib().cx_transfer(self.h, len(fD), ptr_array([ptr(a) for a in fD]), ptr_array([ptr(a) for a in fR])))

    def transfer_rot(self, fieldsD, fieldsR, dirR, theta, triplets):
        """transfer() followed by the periodicity-by-rotation correction (includeTransfers.h): triplets =
        list of (ix, iy, iz) positions in the variable list; cos/sin by the host libm like the reference."""
        fD = self._check_fields(fieldsD, self.nptsD, "donor")
        fR = self._check_fields(fieldsR, self.nptsR, "receptor")
        assert len(fD) == len(fR)
        trip = np.ascontiguousarray(np.asarray(triplets, np.int32).reshape(-1))
        check(lib().cx_transfer_rot(self.h, len(fD), ptr_array([ptr(a) for a in fD]), ptr_array([ptr(a) for a in fR]),
                                    int(dirR), math.cos(theta), math.sin(theta), trip.size // 3, ptr(trip)))

    def rotate_dev(self, dirR, theta, dfx, dfy, dfz):
        check(lib().cx_rotate_dev(self.h, int(dirR), math.cos(theta), math.sin(theta), _vp(dfx), _vp(dfy), _vp(dfz)))

    def transferD(self, fieldsD):
        """Dense (nvars, n) host array (connector._setInterpTransfersD)."""
        fD = self._check_fields(fieldsD, self.nptsD, "donor")
        out = np.empty((len(fD), self.n))
        check(lib().cx_transferD(self.h, len(fD), ptr_array([ptr(a) for a in fD]), ptr(out)))
        return out

    def transfer_celln(self, cellND, cellNR):
        cD = self._check_fields([cellND], self.nptsD, "donor")[0]
        cR = self._check_fields([cellNR], self.nptsR, "receptor")[0]
        check(lib().cx_transfer_celln(self.h, ptr(cD), ptr(cR)))

    def transfer_dev(self, dptrsD, dptrsR):
        check(lib().cx_transfer_dev(self.h, len(dptrsD), ptr_array(dptrsD), ptr_array(dptrsR)))

    def transferD_dev(self, dptrsD, dout, ld):
        check(lib().cx_transferD_dev(self.h, len(dptrsD), ptr_array(dptrsD), _vp(dout), int(ld)))

    def transfer_celln_dev(self, dD, dR):
        check(lib().cx_transfer_celln_dev(self.h, _vp(dD), _vp(dR)))

    def sorted_rcv(self):
        """Receptor indices in the plan's own entry order (the order of a mode-2 dense block)."""
        out = np.empty(self.n, np.int32)
        check(lib().cx_plan_sorted_rcv(self.h, ptr(out)))
        return out

    def info(self):
        n = _i(); ng = _i(); af = _ll(); di = _ll()
        check(lib().cx_plan_info(self.h, ctypes.byref(n), ctypes.byref(ng), ctypes.byref(af), ctypes.byref(di)))
        return dict(n=n.value, ngroups=ng.value, alg_fixed=af.value, distinct=di.value)

    def upload_points(self):
        """Donor points per variable a host-pointer transfer of this plan moves to the device."""
        n = _ll()
        check(lib().cx_plan_upload_points(self.h, ctypes.byref(n)))
        return n.value

    def alg_bytes(self, nvars):
        """B_alg of SURVEY.md 8(d) for one call with nvars variables."""
        i = self.info()
        return i["alg_fixed"] + 8 * nvars * (i["n"] + i["distinct"])


def _celln_dims(cellN):
    if cellN.dtype != np.float64 or not cellN.flags.f_contiguous:
        raise TypeError("cellN must be a Fortran-ordered float64 array")
    return (list(cellN.shape) + [1, 1])[:3]


def apply_bc_overlaps(ctx, cellN, windows, depth, loc, val):
    """connector.applyBCOverlapStruct for every (minIndex, maxIndex) window of one zone, on a Fortran-ordered
    (im,jm,km) float64 array, in place (device: one upload, one launch per window, one download)."""
    im, jm, km = _celln_dims(cellN)
    win = np.ascontiguousarray([list(a) + list(b) for a, b in windows], dtype=np.int32)
    check(lib().cx_apply_bc_overlaps(ctx.h, im, jm, km, len(windows), ptr(win), int(depth), 0 if loc == 'nodes' else 1,
                                     float(val), ptr(cellN)))


def hole_fringe(ctx, cellN, depth, dir):
    """connector._getOversetHolesInterp* on a Fortran-ordered (im,jm,km) float64 array, in place (device); depth < 0:
    the fringe grows into the blanked region."""
    im, jm, km = _celln_dims(cellN)
    check(lib().cx_hole_fringe(ctx.h, im, jm, km, int(depth), int(dir), ptr(cellN)))


def gather_dev(ctx, d_out, ld, n, d_idx, field_ptrs):
    """out[v*ld + e] = field[v][idx[e]] on the device (asynchronous on the context's stream)."""
    nv = len(field_ptrs)
    tab = (ctypes.c_void_p * nv)(*[int(p) for p in field_ptrs])
    check(lib().cx_gather_dev(ctx.h, nv, _vp(d_out), int(ld), int(n), _vp(d_idx), tab))


def host_scatter(dst, idx, src):
    """dst.flat[idx[e]] = src[e] for an array whose memory is its Fortran-order flattening (threaded on the host)."""
    assert dst.dtype == np.float64 and src.dtype == np.float64 and idx.dtype == np.int32
    assert dst.flags.f_contiguous or dst.ndim == 1
    check(lib().cx_host_scatter(ptr(dst), ptr(idx), ptr(src), int(idx.size)))


class HostIO:
    """The copies of a host-to-host step (cx_hostio): per zone and direction, whole fields or listed points only."""

    def __init__(self, ctx, nvars):
        self.ctx = ctx
        self.nvars = int(nvars)
        h = _vp()
        check(lib().cx_hostio_create(ctx.h, self.nvars, ctypes.byref(h)))
        self.h = h
        self._keep = []
        self._fin = weakref.finalize(self, lib().cx_hostio_destroy, h)

    def add(self, direction, host_arrays, dev_ptrs, idx=None):
        """direction 'up' | 'down'; host_arrays: nvars Fortran-flat float64 arrays (kept alive here); idx: int32 points
        to move (None: the whole arrays)."""
        assert len(host_arrays) == self.nvars and len(dev_ptrs) == self.nvars
        for a in host_arrays:
            if a.dtype != np.float64 or not (a.flags.f_contiguous or a.ndim == 1):
                raise TypeError("HostIO: fields must be Fortran-flat float64 arrays")
        n = int(idx.size) if idx is not None else int(host_arrays[0].size)
        if idx is not None:
            idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._keep.append(list(host_arrays))
        check(lib().cx_hostio_add(self.h, 0 if direction == 'up' else 1, n, ptr(idx) if idx is not None else None,
                                  ptr_array([a.ctypes.data for a in host_arrays]), ptr_array(list(dev_ptrs))))

    def add_box(self, direction, host_arrays, dev_ptrs, box):
        """The same for the index box (i0, i1, j0, j1, k0, k1), inclusive, of 3-D Fortran-ordered arrays: one strided
        copy per variable."""
        assert len(host_arrays) == self.nvars and len(dev_ptrs) == self.nvars
        for a in host_arrays:
            if a.dtype != np.float64 or a.ndim != 3 or not a.flags.f_contiguous:
                raise TypeError("HostIO.add_box: fields must be 3-D Fortran-ordered float64 arrays")
        im, jm, km = host_arrays[0].shape
        b = np.ascontiguousarray(box, dtype=np.int32)
        self._keep.append(list(host_arrays))
        check(lib().cx_hostio_add_box(self.h, 0 if direction == 'up' else 1, im, jm, km, ptr(b),
                                      ptr_array([a.ctypes.data for a in host_arrays]), ptr_array(list(dev_ptrs))))

    def bytes(self):
        a = _ll(); b = _ll()
        check(lib().cx_hostio_bytes(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value


def plan_donor_mask(ctx, plans, nptsD):
    """uint8 mask (host) of the donor points read by `plans` (all of one donor zone of nptsD points)."""
    d = DeviceArray(ctx, nptsD, np.uint8)
    ctx.upload(d.ptr, np.zeros(nptsD, np.uint8))
    for p in plans:
        check(lib().cx_plan_mark_donors(p.h, _vp(d.ptr)))
    return d.download()


def host_register(a):
    """Page-lock a contiguous numpy array in place (cudaHostRegister); returns True if this call pinned it."""
    rc = lib().cx_host_register(ptr(a), a.nbytes)
    if rc < 0:
        check(rc)
    return rc == 0


def host_unregister(a):
    lib().cx_host_unregister(ptr(a))


class _PinnedBlock:
    def __init__(self, nbytes):
        p = _vp()
        check(lib().cx_pinned_alloc(int(nbytes), ctypes.byref(p)))
        self.ptr = p.value
        self.nbytes = int(nbytes)
        self._fin = weakref.finalize(self, lib().cx_pinned_free, _vp(self.ptr))


# Page-locked blocks released by the garbage collector are kept for reuse (cudaHostAlloc of GB-sized
# blocks costs hundreds of ms); CX_PINNED_POOL_MB bounds what is kept (default 8192).
_PINNED_FREE = []
_PINNED_KEEP = {}


def _pinned_release(key):
    blk = _PINNED_KEEP.pop(key, None)
    if blk is None:
        return
    cap = int(os.environ.get("CX_PINNED_POOL_MB", "8192")) << 20
    if sum(b.nbytes for b in _PINNED_FREE) + blk.nbytes <= cap:
        _PINNED_FREE.append(blk)


def _pinned_take(nbytes):
    best = None
    for i, b in enumerate(_PINNED_FREE):
        if b.nbytes >= nbytes and b.nbytes <= 2 * nbytes + (1 << 20) and (best is None or b.nbytes < _PINNED_FREE[best].nbytes):
            best = i
    if best is not None:
        return _PINNED_FREE.pop(best)
    return _PinnedBlock(nbytes)


def pinned_empty(n, dtype=np.float64):
    """numpy array backed by page-locked host memory (cudaHostAlloc); when the array
    (and its views) are collected the block goes back to a reuse pool."""
    dtype = np.dtype(dtype)
    nbytes = max(1, int(n)) * dtype.itemsize
    blk = _pinned_take(nbytes)
    buf = (ctypes.c_byte * nbytes).from_address(blk.ptr)
    a = np.frombuffer(buf, dtype=dtype, count=int(n))
    _PINNED_KEEP[id(buf)] = blk
    weakref.finalize(buf, _pinned_release, id(buf))
    return a


def pinned_copy(a):
    a = np.asarray(a)
    out = pinned_empty(a.size, a.dtype)
    out[:] = a.reshape(-1, order='A')
    return out


class PlanSet:
    """Every plan of a rank fused into one launch per interpolation type
    (cx_planset).  items: list of (plan, mode, donorFieldPtrs, receptorFieldPtrs|None, densePtr|None, ld)."""

    def __init__(self, ctx, nvars, items):
        self.ctx = ctx
        self.items = items     # keep plans / arrays alive
        n = len(items)
        plans = (ctypes.c_void_p * max(n, 1))(*[it[0].h.value for it in items])
        modes = (ctypes.c_int * max(n, 1))(*[int(it[1]) for it in items])
        self._tabs = []
        fD = (ctypes.c_void_p * max(n, 1))()
        fR = (ctypes.c_void_p * max(n, 1))()
        dn = (ctypes.c_void_p * max(n, 1))()
        ld = (ctypes.c_longlong * max(n, 1))()
        for i, it in enumerate(items):
            td = ptr_array(it[2]); self._tabs.append(td)
            fD[i] = ctypes.cast(td, ctypes.c_void_p).value
            if it[1] == 0:
                tr = ptr_array(it[3]); self._tabs.append(tr)
                fR[i] = ctypes.cast(tr, ctypes.c_void_p).value
            else:
                dn[i] = it[4]; ld[i] = int(it[5])
        h = _vp()
        check(lib().cx_planset_create(ctx.h, n, plans, int(nvars), modes, fD, fR, dn, ld, ctypes.byref(h)))
        self.h = h
        self._fin = weakref.finalize(self, lib().cx_planset_destroy, h)

    def run(self):
        check(lib().cx_planset_run(self.h))


class ScatterSet:
    """Receptor-side scatter of all received dense blocks in one launch (cx_scatterset).
    items: list of (n, rcvIndexDeviceArray, inPtr, ld, receptorFieldPtrs)."""

    def __init__(self, ctx, nvars, items):
        self.ctx = ctx
        self.items = items
        n = len(items)
        cnt = (ctypes.c_int * max(n, 1))(*[int(it[0]) for it in items])
        rcv = (ctypes.c_void_p * max(n, 1))(*[it[1].ptr for it in items])
        din = (ctypes.c_void_p * max(n, 1))(*[it[2] for it in items])
        ld = (ctypes.c_longlong * max(n, 1))(*[int(it[3]) for it in items])
        self._tabs = []
        fR = (ctypes.c_void_p * max(n, 1))()
        for i, it in enumerate(items):
            tr = ptr_array(it[4]); self._tabs.append(tr)
            fR[i] = ctypes.cast(tr, ctypes.c_void_p).value
        h = _vp()
        check(lib().cx_scatterset_create(ctx.h, n, int(nvars), cnt, rcv, din, ld, fR, ctypes.byref(h)))
        self.h = h
        self._fin = weakref.finalize(self, lib().cx_scatterset_destroy, h)

    def run(self):
        check(lib().cx_scatterset_run(self.h))


class _Pending:
    def __init__(self, f):
        self._f = f

    def fetch(self):
        return self._f()


class DeviceBackend:
    """Compute backend of the Python layer = the CUDA library.  (The distributed
    host logic takes a backend object so the CPU tests can inject the oracle;
    the product only ever constructs this one.)"""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()

    def make_zone(self, ni, nj, nk, x, y, z, cellN=None, interpDataType=1):
        return Zone(self.ctx, ni, nj, nk, x, y, z, cellN, interpDataType)

    def make_zones(self, specs):
        """Hooks of several donor zones, specs = [(ni, nj, nk, x, y, z, cellN, interpDataType)]: the uploads run zone
        after zone, the ADT builds of the whole batch run concurrently on the device (cx_zones_build_adt)."""
        zones = self.make_zones_begin(specs)
        self.make_zones_finish(zones)
        return zones

    def make_zones_begin(self, specs):
        """Uploads zone after zone, every ADT build started on a side stream as soon as its zone is up; the zones can
        serve as receptor sources (coordinates, cellN) at once, as donors after make_zones_finish."""
        zones = []
        for sp in specs:
            z = Zone(self.ctx, *sp[:7], interpDataType=sp[7], defer_adt=True)
            check(lib().cx_zone_adt_begin(z.h))      # builds while the next zone's coordinates go up
            zones.append(z)
        return zones

    def make_zones_finish(self, zones):
        build_adts(self.ctx, zones)

    def set_celln(self, hook, cellN):
        hook.set_celln(cellN)

    def set_interp_data(self, xr, yr, zr, hooks, order, nature, penalty, extrap):
        res = set_interp_data(self.ctx, xr, yr, zr, hooks, order, nature, penalty, extrap)
        return res.fetch(pinned=True)

    def select_receptors(self, ni, nj, nk, x, y, z, cellN, loc):
        """Receptor points (cellN == 2 at loc) of one zone, extracted and kept on the device."""
        return Receptors(self.ctx, ni, nj, nk, x, y, z, cellN, loc)

    def select_receptors_from_hook(self, hook):
        """The receptor points of a zone whose donor hook `hook` holds the same node coordinates and node cellN."""
        return Receptors.from_zone(self.ctx, hook)

    def set_interp_data_rcv(self, rcv, hooks, order, nature, penalty, extrap):
        # large lists land in page-locked arrays of the reuse pool (they become the tree's numpy arrays)
        return set_interp_data_rcv(self.ctx, rcv, hooks, order, nature, penalty, extrap).fetch(pinned=True)

    def submit_interp_data_rcv(self, rcv, hooks, order, nature, penalty, extrap):
        """Queue the search of one receptor zone on the device and return at once; .fetch() of the returned object
        waits for it and gives the lists."""
        self._slot = getattr(self, '_slot', 0) % 8 + 1          # round robin over the side streams
        check(lib().cx_ctx_search_slot(self.ctx.h, self._slot))
        try:
            res = set_interp_data_rcv(self.ctx, rcv, hooks, order, nature, penalty, extrap)
        finally:
            check(lib().cx_ctx_search_slot(self.ctx.h, 0))
        return _Pending(lambda: res.fetch(pinned=True))

    def make_plan(self, imd, jmd, kmd, nptsR, donor, rcv, typ, coefs):
        return Plan(self.ctx, imd, jmd, kmd, nptsR, donor, rcv, typ, coefs)

    def apply_bc_overlaps(self, cellN, windows, depth, loc, val):
        apply_bc_overlaps(self.ctx, cellN, windows, depth, loc, val)

    def hole_fringe(self, cellN, depth, dir):
        hole_fringe(self.ctx, cellN, depth, dir)


def ipc_export(dptr):
    """72-byte CUDA IPC descriptor (handle of the containing allocation + offset) of a device pointer."""
    buf = ctypes.create_string_buffer(72)
    check(lib().cx_ipc_export(_vp(dptr), buf))
    return bytes(buf.raw)


def ipc_open(ctx, desc):
    """Map a peer process's device memory exported with ipc_export; returns the device pointer here."""
    p = _vp()
    check(lib().cx_ipc_open(ctx.h, ctypes.c_char_p(desc), ctypes.byref(p)))
    return p.value


def ipc_close(desc):
    lib().cx_ipc_close(ctypes.c_char_p(desc))


class _StepRun:
    """What both kinds of step (cx_step) can run."""

    def run(self, niter=1, overlap=True):
        check(lib().cx_step_run(self.h, int(niter), 1 if overlap else 0))

    def run_host(self, io, pipelined=True):
        check(lib().cx_step_run_host(self.h, io.h, 1 if pipelined else 0))

    def run_vars(self, v0, nv, overlap=True):
        """One iteration restricted to the variables [v0, v0 + nv) (cx_step_run_vars)."""
        check(lib().cx_step_run_vars(self.h, 1, 1 if overlap else 0, int(v0), int(nv)))

    def splittable(self):
        return bool(lib().cx_step_splittable(self.h))


class StepP2P(_StepRun):
    """One solver iteration of a rank with the peer-memory transport (cx_step_create_p2p): remote plans store into
    the peers' IPC-mapped receive buffers, flags instead of messages, scatter, overlapped with the local plans."""

    def __init__(self, ctx, ps_remote2, ps_local, ss2, peer_flag_ptrs, my_flags):
        self.ctx = ctx
        self._keep = (ps_remote2, ps_local, ss2, my_flags)
        n = len(peer_flag_ptrs)
        h = _vp()
        hv = lambda o: o.h if o is not None else None
        check(lib().cx_step_create_p2p(ctx.h, hv(ps_remote2[0]), hv(ps_remote2[1]), hv(ps_local), hv(ss2[0]), hv(ss2[1]),
                                       n, ptr_array(peer_flag_ptrs) if n else None, _vp(my_flags.ptr), ctypes.byref(h)))
        self.h = h
        self._fin = weakref.finalize(self, lib().cx_step_destroy, h)



class Step(_StepRun):
    """One iteration of a rank (cx_step): remote plans -> NCCL exchange + scatter (second
    stream) overlapped with the local plans."""

    def __init__(self, ctx, ps_remote, ps_local, scatterset, peers, send_ptrs, send_counts, recv_ptrs, recv_counts):
        self.keep = (ps_remote, ps_local, scatterset)
        n = len(peers)
        h = _vp()
        check(lib().cx_step_create(ctx.h, ps_remote.h if ps_remote is not None else None,
                                   ps_local.h if ps_local is not None else None,
                                   scatterset.h if scatterset is not None else None, n,
                                   (ctypes.c_int * max(n, 1))(*peers), ptr_array(send_ptrs or [0]),
                                   (ctypes.c_longlong * max(n, 1))(*send_counts), ptr_array(recv_ptrs or [0]),
                                   (ctypes.c_longlong * max(n, 1))(*recv_counts), ctypes.byref(h)))
        self.h = h
        self._fin = weakref.finalize(self, lib().cx_step_destroy, h)

