"""GPU parity tests (run with -m gpu on the B200): CUDA path through the C ABI
vs the oracle (= the reference's own K_INTERP C++ compiled from
/root/reference, oracle/_ref/libcassref.so) on the same seeded inputs.
Bar: donor indices, receptor indices, types, extrapolation and orphan lists
bit-identical and identically ordered; coefficients and transferred fields
bit-identical as well (the device code reproduces the reference's un-fused
fp64 expression order; tolerance stated by BASELINE.json is 1e-12 relative)."""
import numpy as np
import pytest
from tests import cases

pytestmark = pytest.mark.gpu


def _zones(ctx, dz, cellNs=None, types=None):
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    gz, oz = [], []
    for i, (ni, nj, nk, x, y, z) in enumerate(dz):
        c = cellNs[i] if cellNs else None
        t = types[i] if types else 1
        gz.append(_lib.Zone(ctx, ni, nj, nk, x, y, z, c, interpDataType=t))
        oz.append(O.RefZone(ni, nj, nk, x, y, z, c, interpDataType=t))
    return gz, oz


def _compare(ctx, dz, pts, order=2, nature=1, penalty=1, extrap=1, cellNs=None, rtol=0.0, types=None):
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    gz, oz = _zones(ctx, dz, cellNs, types)
    for g, o in zip(gz, oz):
        assert np.array_equal(g.info()["bbox"], o.bbox())
    res = _lib.set_interp_data(ctx, pts[0], pts[1], pts[2], gz, order, nature, penalty, extrap)
    got = res.fetch()
    ref = O.set_interp_data(pts[0], pts[1], pts[2], oz, order, nature, penalty, extrap)
    for z in range(len(dz)):
        assert np.array_equal(got[0][z], ref[0][z]), "receptor list zone %d" % z
        assert np.array_equal(got[1][z], ref[1][z]), "donor list zone %d" % z
        assert np.array_equal(got[2][z], ref[2][z]), "type list zone %d" % z
        assert np.array_equal(got[4][z], ref[4][z]), "extrap list zone %d" % z
        if rtol == 0.0:
            assert np.array_equal(got[3][z], ref[3][z]), "coefficients zone %d" % z
        else:
            assert got[3][z].shape == ref[3][z].shape
            assert np.allclose(got[3][z], ref[3][z], rtol=rtol, atol=rtol)
    assert np.array_equal(got[5], ref[5]), "orphans"
    return got, ref, gz, oz


@pytest.mark.parametrize("variant", ["a", "b", "c"])
def test_c1_cart_pairs(ctx, variant):
    A, B = cases.c1(variant)
    _compare(ctx, [A], B[3:])
    _compare(ctx, [B], A[3:])          # mostly extrapolated / orphans


def test_adt_candidates_match_reference(ctx):
    A, B = cases.c1("a", n=17, m=9)
    gz, oz = _zones(ctx, [A])
    rng = np.random.default_rng(3)
    pts = list(zip(B[3][:50], B[4][:50], B[5][:50])) + [tuple(p) for p in rng.uniform(-0.1, 1.1, (50, 3))]
    for (x, y, z) in pts:
        for alpha in (0.0, 7.5):
            assert np.array_equal(gz[0].candidates(x, y, z, alpha), oz[0].candidates(x, y, z, alpha))


def test_c2_cart_cylinder(ctx):
    Bg, Oc = cases.c2()
    _compare(ctx, [Bg], Oc[3:])
    _compare(ctx, [Oc], Bg[3:])


@pytest.mark.parametrize("nature", [0, 1])
@pytest.mark.parametrize("penalty", [0, 1])
def test_multizone_holes(ctx, nature, penalty):
    Bg, Oc = cases.c2()
    rng = np.random.default_rng(0)
    cb = np.ones(Bg[3].size); cb[rng.random(cb.size) < 0.05] = 0.; cb[rng.random(cb.size) < 0.05] = 2.
    co = np.ones(Oc[3].size); co[rng.random(co.size) < 0.05] = 0.; co[rng.random(co.size) < 0.05] = 2.
    P = rng.uniform(-3.3, 3.3, (3, 20000)); P[2] = rng.uniform(-0.2, 2.2, 20000)
    _compare(ctx, [Bg, Oc], P, nature=nature, penalty=penalty, cellNs=[cb, co])
    _compare(ctx, [Oc, Bg], P, nature=nature, penalty=penalty, extrap=0, cellNs=[co, cb])


def test_empty_and_tiny(ctx):
    from cassiopee_b200 import _lib
    A, B = cases.c1("c", n=9, m=5)
    gz, oz = _zones(ctx, [A])
    res = _lib.set_interp_data(ctx, np.zeros(0), np.zeros(0), np.zeros(0), gz)
    got = res.fetch()
    assert got[0][0].size == 0 and got[5].size == 0
    _compare(ctx, [A], (B[3][:1], B[4][:1], B[5][:1]))


@pytest.mark.parametrize("nvars", [1, 5, 7])
def test_transfer_matches_reference(ctx, nvars):
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    Bg, Oc = cases.c2()
    got, ref, gz, oz = _compare(ctx, [Bg], Oc[3:])
    rcv, dnr, typ, cf = got[0][0], got[1][0], got[2][0], got[3][0]
    fD = cases.fields(Bg[3], Bg[4], Bg[5], nvars)
    nR = Oc[3].size
    fR_g = [np.zeros(nR) for _ in range(nvars)]
    fR_o = [np.zeros(nR) for _ in range(nvars)]
    plan = _lib.Plan(ctx, Bg[0], Bg[1], Bg[2], nR, dnr, rcv, typ, cf)
    plan.transfer(fD, fR_g)
    O.transfer(fD, fR_o, Bg[0], Bg[1], rcv, dnr, typ, cf)
    for a, b in zip(fR_g, fR_o):
        assert np.array_equal(a, b)
    dense = plan.transferD(fD)
    assert np.array_equal(dense, O.transferD(fD, Bg[0], Bg[1], dnr, typ, cf))
    # device-resident variant
    dD = [_lib.DeviceArray.from_host(ctx, f) for f in fD]
    dR = [_lib.DeviceArray.from_host(ctx, np.zeros(nR)) for _ in range(nvars)]
    plan.transfer_dev([d.ptr for d in dD], [d.ptr for d in dR])
    ctx.sync()
    for d, b in zip(dR, fR_o):
        assert np.array_equal(d.download(), b)
    info = plan.info()
    assert info["n"] == rcv.size and info["distinct"] > 0


def test_transfer_mixed_types_and_celln(ctx):
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    A, B = cases.c1("a")          # type 1 (coincident) entries
    got, ref, gz, oz = _compare(ctx, [A], B[3:])
    A2, B2 = cases.c1("c")
    rng = np.random.default_rng(5)
    rcv, dnr, typ, cf = got[0][0], got[1][0], got[2][0], got[3][0]
    assert (typ == 1).any()
    nD = A[3].size; nR = B[3].size
    fD = cases.fields(A[3], A[4], A[5], 5)
    fRg = [np.zeros(nR) for _ in range(5)]; fRo = [np.zeros(nR) for _ in range(5)]
    plan = _lib.Plan(ctx, A[0], A[1], A[2], nR, dnr, rcv, typ, cf)
    plan.transfer(fD, fRg)
    O.transfer(fD, fRo, A[0], A[1], rcv, dnr, typ, cf)
    for a, b in zip(fRg, fRo):
        assert np.array_equal(a, b)
    cD = np.ones(nD); cD[rng.random(nD) < 0.2] = 0.; cD[rng.random(nD) < 0.2] = 2.
    cRg = np.full(nR, 2.); cRo = np.full(nR, 2.)
    plan.transfer_celln(cD, cRg)
    O.transfer_celln(cD, cRo, A[0], A[1], rcv, dnr, typ, cf)
    assert np.array_equal(cRg, cRo)


def test_linear_field_exact(ctx):
    """Known answer of Connector/test/setInterpTransfersPT_t1.py: F=x+2y+3z is
    reproduced by the interpolation."""
    from cassiopee_b200 import _lib
    Bg, Oc = cases.c2()
    gz, _ = _zones(ctx, [Bg])
    res = _lib.set_interp_data(ctx, Oc[3], Oc[4], Oc[5], gz, 2, 1, 1, 1).fetch()
    F = Bg[3] + 2 * Bg[4] + 3 * Bg[5]
    out = np.zeros(Oc[3].size)
    plan = _lib.Plan(ctx, Bg[0], Bg[1], Bg[2], out.size, res[1][0], res[0][0], res[2][0], res[3][0])
    plan.transfer([F], [out])
    exact = Oc[3] + 2 * Oc[4] + 3 * Oc[5]
    assert np.abs(out - exact)[res[0][0]].max() < 1e-12


def test_pytree_api_inverse_storage(ctx):
    import __graft_entry__ as g
    g.smoke()


@pytest.mark.parametrize("order", [3, 4, 5])
def test_lagrange_orders_cart_and_cylinder(ctx, order):
    """Orders 3/5 (directional 9/15 coefficients).  The oracle for these orders
    is the reference C++ plus the C transliteration of four Fortran routines
    (oracle/lagrange_restated.c)."""
    Bg, Oc = cases.c2()
    rng = np.random.default_rng(11)
    sel = rng.choice(Oc[3].size, 1500 if order == 5 else 8000, replace=False)
    sel.sort()
    got, ref, gz, oz = _compare(ctx, [Bg], (Oc[3][sel], Oc[4][sel], Oc[5][sel]), order=order)
    assert (got[2][0] == (44 if order == 4 else order)).any()
    r = np.sqrt(Bg[3] ** 2 + Bg[4] ** 2)
    m = np.nonzero((r > 0.5) & (r < 2.5))[0]
    sel = rng.choice(m, 1000 if order == 5 else 6000, replace=False)
    sel.sort()
    _compare(ctx, [Oc], (Bg[3][sel], Bg[4][sel], Bg[5][sel]), order=order)


@pytest.mark.parametrize("order", [3, 4, 5])
@pytest.mark.parametrize("nature", [0, 1])
def test_lagrange_multizone_holes(ctx, order, nature):
    Bg, Oc = cases.c2()
    rng = np.random.default_rng(1)
    cb = np.ones(Bg[3].size); cb[rng.random(cb.size) < 0.03] = 0.; cb[rng.random(cb.size) < 0.03] = 2.
    co = np.ones(Oc[3].size); co[rng.random(co.size) < 0.03] = 0.; co[rng.random(co.size) < 0.03] = 2.
    npt = 1200 if order == 5 else 6000
    P = rng.uniform(-3.3, 3.3, (3, npt)); P[2] = rng.uniform(-0.2, 2.2, npt)
    _compare(ctx, [Bg, Oc], P, order=order, nature=nature, cellNs=[cb, co])
    _compare(ctx, [Oc, Bg], P, order=order, nature=nature, extrap=0, cellNs=[co, cb])


@pytest.mark.parametrize("order", [3, 4, 5])
def test_lagrange_coincident_and_transfer(ctx, order):
    """C1a: receptors coincide with donor nodes -> type 1 collapse of the
    directional weights; then order-3/5 transfers vs the reference loops and
    linear-field exactness (Connector/test/setInterpTransfersPT_t1.py)."""
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    A, B = cases.c1("a", n=17, m=9)
    _compare(ctx, [A], B[3:], order=order)
    A, B = cases.c1("c", n=17, m=9)
    got, ref, gz, oz = _compare(ctx, [A], B[3:], order=order)
    rcv, dnr, typ, cf = got[0][0], got[1][0], got[2][0], got[3][0]
    fD = cases.fields(A[3], A[4], A[5], 5) + [A[3] + 2 * A[4] + 3 * A[5]]
    nR = B[3].size
    fRg = [np.zeros(nR) for _ in fD]; fRo = [np.zeros(nR) for _ in fD]
    plan = _lib.Plan(ctx, A[0], A[1], A[2], nR, dnr, rcv, typ, cf)
    plan.transfer(fD, fRg)
    O.transfer(fD, fRo, A[0], A[1], rcv, dnr, typ, cf)
    for a, b in zip(fRg, fRo):
        assert np.array_equal(a, b)
    exact = B[3] + 2 * B[4] + 3 * B[5]
    ext = np.zeros(nR, bool); ext[got[4][0]] = True
    ok = np.zeros(nR, bool); ok[rcv] = True
    assert np.abs(fRg[-1] - exact)[ok & ~ext].max() < 1e-10


@pytest.mark.parametrize("dims", [(33, 31, 9), (64, 16, 8), (7, 5, 4), (401, 3, 3)])
@pytest.mark.parametrize("density", [3.0, 0.05])
def test_transfer_synthetic_plans_staged_and_gather(ctx, dims, density):
    """Random type-1/2/3 plans straight into the transfer kernels: odd row pitch (16-byte misaligned
    row segments in the shared-memory staged kernel), dense and sparse donor spans (staged tiles and
    direct-gather tiles in one launch), stencils touching the last node of the zone, non-contiguous and
    contiguous receptor lists (host scatter vs in-place download).  Bit-exact vs the reference loops."""
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    ni, nj, nk = dims
    rng = np.random.default_rng(ni * 1000 + nj)
    nD = ni * nj * nk
    n = max(8, int(density * nD))
    nR = 2 * n + 5
    for contiguous in (False, True):
        typ = rng.choice(np.array([1, 2, 2, 2, 2, 3, 3, 5, 44], np.int32), n).astype(np.int32)
        if ni < 5 or nj < 5 or nk < 5:
            typ[typ == 5] = 3
        if ni < 4 or nj < 4 or nk < 4:
            typ[typ == 44] = 3
        if ni < 3 or nj < 3 or nk < 3:
            typ[typ == 3] = 2
        o = np.where(typ == 1, 1, np.where(typ == 44, 4, typ))
        i = (rng.random(n) * (ni - o + 1)).astype(np.int64); j = (rng.random(n) * (nj - o + 1)).astype(np.int64)
        k = (rng.random(n) * (nk - o + 1)).astype(np.int64)
        i[:3] = ni - o[:3]; j[:3] = nj - o[:3]; k[:3] = nk - o[:3]          # stencil ends on the last node
        dnr = (i + ni * j + ni * nj * k).astype(np.int32)
        rcv = (np.arange(n) + 3 if contiguous else np.sort(rng.choice(nR, n, replace=False))).astype(np.int32)
        ncf = np.where(typ == 1, 1, np.where(typ == 2, 8, np.where(typ == 3, 9, np.where(typ == 44, 12, 15))))
        cf = rng.standard_normal(int(ncf.sum()))
        for nvars in (1, 2, 5):
            fD = [rng.standard_normal(nD) for _ in range(nvars)]
            fRg = [np.full(nR, -7.) for _ in range(nvars)]; fRo = [np.full(nR, -7.) for _ in range(nvars)]
            plan = _lib.Plan(ctx, ni, nj, nk, nR, dnr, rcv, typ, cf)
            plan.transfer(fD, fRg)
            O.transfer(fD, fRo, ni, nj, rcv, dnr, typ, cf)
            for a, b in zip(fRg, fRo):
                assert np.array_equal(a, b)
            assert np.array_equal(plan.transferD(fD), O.transferD(fD, ni, nj, dnr, typ, cf))
            dD = [_lib.DeviceArray.from_host(ctx, f) for f in fD]
            dR = [_lib.DeviceArray.from_host(ctx, np.full(nR, -7.)) for _ in range(nvars)]
            plan.transfer_dev([d.ptr for d in dD], [d.ptr for d in dR])
            ctx.sync()
            for d, b in zip(dR, fRo):
                assert np.array_equal(d.download(), b)
            # the same plan through the fused plan-set launch, in place and dense
            dR2 = [_lib.DeviceArray.from_host(ctx, np.full(nR, -7.)) for _ in range(nvars)]
            dense = _lib.DeviceArray(ctx, n * nvars)
            ps = _lib.PlanSet(ctx, nvars, [(plan, 0, [d.ptr for d in dD], [d.ptr for d in dR2], None, 0),
                                           (plan, 1, [d.ptr for d in dD], None, dense.ptr, n)])
            ps.run(); ctx.sync()
            for d, b in zip(dR2, fRo):
                assert np.array_equal(d.download(), b)
            assert np.array_equal(dense.download().reshape(nvars, n), O.transferD(fD, ni, nj, dnr, typ, cf))


def test_host_pointer_transfer_sparse_upload(ctx, monkeypatch):
    """cx_transfer on a donor zone of more than 2^20 points of which the plan reads a part: only the donor points of
    the stencils cross the bus (host gather, packed copy, device store by index); same fields bit for bit as the
    reference loops and as the whole-field upload, for type 1/2/3/5 entries and stencils on the zone's last nodes."""
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    ni, nj, nk = 129, 97, 88
    nD = ni * nj * nk
    assert nD >= 1 << 20
    rng = np.random.default_rng(12)
    n = 90000
    nR = 2 * n + 11
    typ = rng.choice(np.array([1, 2, 2, 2, 3, 5], np.int32), n).astype(np.int32)
    o = np.where(typ == 1, 1, typ)
    # receptors fed from a slab of the donor zone (k < 16) plus a few stencils that end on the very last node
    i = (rng.random(n) * (ni - o + 1)).astype(np.int64); j = (rng.random(n) * (nj - o + 1)).astype(np.int64)
    k = (rng.random(n) * (16 - o + 1)).astype(np.int64)
    i[:4] = ni - o[:4]; j[:4] = nj - o[:4]; k[:4] = nk - o[:4]
    dnr = (i + ni * j + ni * nj * k).astype(np.int32)
    rcv = np.sort(rng.choice(nR, n, replace=False)).astype(np.int32)
    ncf = np.where(typ == 1, 1, np.where(typ == 2, 8, np.where(typ == 3, 9, 15)))
    cf = rng.standard_normal(int(ncf.sum()))
    fD = [rng.standard_normal(nD) for _ in range(3)]
    want = [np.full(nR, -7.) for _ in range(3)]
    O.transfer(fD, want, ni, nj, rcv, dnr, typ, cf)
    plan = _lib.Plan(ctx, ni, nj, nk, nR, dnr, rcv, typ, cf)
    up = plan.upload_points()
    assert 0 < up < 0.25 * nD                                  # the slab only
    for rep in range(2):                                       # second call: the lists and staging blocks are reused
        got = [np.full(nR, -7.) for _ in range(3)]
        plan.transfer(fD, got)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    assert np.array_equal(plan.transferD(fD), O.transferD(fD, ni, nj, dnr, typ, cf))
    # a plan that reads a block of the zone (more than a quarter of it, its bounding box under 85 %): the box goes up as
    # one strided copy per variable
    typ3 = rng.choice(np.array([1, 2, 3], np.int32), n).astype(np.int32)
    o3 = np.where(typ3 == 1, 1, typ3)
    i3 = 20 + (rng.random(n) * (70 - o3 + 1)).astype(np.int64); j3 = 5 + (rng.random(n) * (80 - o3 + 1)).astype(np.int64)
    k3 = (rng.random(n) * (nk - o3 + 1)).astype(np.int64)
    i3[:2] = 20; j3[:2] = 5; k3[:2] = 0; i3[2:4] = 20 + 70 - o3[2:4]; j3[2:4] = 5 + 80 - o3[2:4]; k3[2:4] = nk - o3[2:4]
    dnr3 = (i3 + ni * j3 + ni * nj * k3).astype(np.int32)
    cf3 = rng.standard_normal(int(np.where(typ3 == 1, 1, np.where(typ3 == 2, 8, 9)).sum()))
    plan3 = _lib.Plan(ctx, ni, nj, nk, nR, dnr3, rcv, typ3, cf3)
    assert plan3.upload_points() == 70 * 80 * nk and 0.25 * nD < plan3.upload_points() < 0.85 * nD
    want3 = [np.full(nR, -7.) for _ in range(3)]
    O.transfer(fD, want3, ni, nj, rcv, dnr3, typ3, cf3)
    for rep in range(2):
        got = [np.full(nR, -7.) for _ in range(3)]
        plan3.transfer(fD, got)
        for a, b in zip(got, want3):
            assert np.array_equal(a, b)
    monkeypatch.setenv("CX_SPARSE_H2D", "0")                   # read once per process: only new plans could differ
    # a plan that reads most of the zone keeps the whole-field upload
    k2 = (rng.random(n) * (nk - o + 1)).astype(np.int64)
    i2 = (np.arange(n) * 7 % (ni - 5)).astype(np.int64); j2 = (np.arange(n) * 13 % (nj - 5)).astype(np.int64)
    typ2 = np.full(n, 5, np.int32)
    k2 = (np.arange(n) * 3 % (nk - 5)).astype(np.int64)
    dnr2 = (i2 + ni * j2 + ni * nj * k2).astype(np.int32)
    cf2 = rng.standard_normal(15 * n)
    plan2 = _lib.Plan(ctx, ni, nj, nk, nR, dnr2, rcv, typ2, cf2)
    got = [np.full(nR, -7.) for _ in range(3)]; want2 = [np.full(nR, -7.) for _ in range(3)]
    plan2.transfer(fD, got)
    O.transfer(fD, want2, ni, nj, rcv, dnr2, typ2, cf2)
    for a, b in zip(got, want2):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("order", [2, 3, 4])
def test_interpcart_donors(ctx, order):
    """interpDataType=0 (regular Cartesian donors located in closed form, K_INTERP::InterpCart): single donor,
    Cartesian + ADT donors mixed in both list orders, holes, nature/penalty, extrapolation and orphans."""
    for variant in ("a", "b", "c"):
        A, B = cases.c1(variant)
        got, ref, _, _ = _compare(ctx, [A], B[3:], order=order, types=[0])
        assert got[0][0].size == B[3].size
        got, ref, _, _ = _compare(ctx, [B], A[3:], order=order, types=[0])
        assert got[5].size > 0
    Bg, Oc = cases.c2()
    rng = np.random.default_rng(11)
    cb = np.ones(Bg[3].size); cb[rng.random(cb.size) < 0.05] = 0.; cb[rng.random(cb.size) < 0.05] = 2.
    co = np.ones(Oc[3].size); co[rng.random(co.size) < 0.05] = 0.; co[rng.random(co.size) < 0.05] = 2.
    P = rng.uniform(-3.3, 3.3, (3, 20000)); P[2] = rng.uniform(-0.2, 2.2, 20000)
    next_ = 0
    for nature in (0, 1):
        for penalty in (0, 1):
            got, _, _, _ = _compare(ctx, [Bg], P, order, nature, penalty, 1, [cb], types=[0])
            next_ += got[4][0].size
            _compare(ctx, [Bg, Oc], P, order, nature, penalty, 1, [cb, co], types=[0, 1])
            _compare(ctx, [Oc, Bg], P, order, nature, penalty, 0, [co, cb], types=[1, 0])
    assert next_ > 0


def test_interpcart_order5_falls_to_extrapolation(ctx):
    """The reference has no order-5 search in a Cartesian InterpData (Interp2.cpp:464-467 returns -1): every
    receptor goes through the order-2 extrapolation pass.  Small case (the reference prints one line per query)."""
    A, B = cases.c1("c", n=13, m=7)
    got, ref, _, _ = _compare(ctx, [A], B[3:], order=5, types=[0])
    assert got[4][0].size == got[0][0].size and (got[2][0] == 2).all()


def test_interpcart_pytree_api(ctx):
    """X.setInterpData(..., interpDataType=0) through the tree layer, transfers of a linear field exact."""
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.PyTree as X
    a = G.cart((0, 0, 0), (0.1, 0.1, 0.1), (11, 11, 11), name='A')
    b = G.cart((0.23, 0.31, 0.17), (0.05, 0.05, 0.05), (9, 9, 9), name='B')
    C._initVars(a, 'F', lambda x, y, z: x + 2 * y + 3 * z)
    C._initVars(b, 'F', 0.)
    C._initVars(b, 'cellN', 2.)
    for order in (2, 3):
        C._initVars(b, 'F', 0.)
        tD = X.setInterpData(C.newPyTree(['R', b]), C.newPyTree(['D', a]), order=order, nature=1, loc='nodes',
                             storage='inverse', interpDataType=0, itype='chimera', verbose=0)
        s = Internal.getNodeFromName(tD, 'ID_B')
        assert s is not None and (Internal.getNodeFromName1(s, 'InterpolantsType')[1] == order).all()
        tR = X.setInterpTransfers(C.newPyTree(['R', b]), tD, variables=['F'], storage=1)
        zb = Internal.getZones(tR)[0]
        x, y, z = C.getCoords(zb)
        assert np.abs(C.getFieldArray(zb, 'F') - (x + 2 * y + 3 * z)).max() < 1e-12


@pytest.mark.parametrize("loc", ["nodes", "centers"])
def test_receptor_extraction_on_device(ctx, loc):
    """cx_receptors_create (getInterpolatedPoints + node2Center of the selected cells on the device) against the
    numpy restatement used by the CPU tests; then setInterpData on the resident receptors == host-pointer call."""
    from cassiopee_b200 import _lib
    from tests.backends import OracleBackend
    Bg, Oc = cases.c2()
    ni, nj, nk, x, y, z = Oc
    rng = np.random.default_rng(9)
    npts = ni * nj * nk if loc == "nodes" else (ni - 1) * (nj - 1) * (nk - 1)
    c = np.ones(npts); c[rng.random(npts) < 0.3] = 2.; c[rng.random(npts) < 0.1] = 0.
    c[rng.random(npts) < 0.05] = 2. + 1.e-15; c[rng.random(npts) < 0.05] = 2. + 4.e-15
    dev = _lib.DeviceBackend(ctx).select_receptors(ni, nj, nk, x, y, z, c, loc)
    ref = OracleBackend().select_receptors(ni, nj, nk, x, y, z, c, loc)
    assert dev.n == ref.n and dev.n > 0
    assert np.array_equal(dev.indcell, ref.indcell)
    for a, b in zip(dev.coords(), ref.pts):
        assert np.array_equal(a, b)
    gz, _ = _zones(ctx, [Bg])
    got = _lib.set_interp_data_rcv(ctx, dev, gz, 2, 1, 1, 1).fetch()
    exp = _lib.set_interp_data(ctx, ref.pts[0], ref.pts[1], ref.pts[2], gz, 2, 1, 1, 1).fetch()
    for q in range(5):
        assert np.array_equal(got[q][0], exp[q][0])
    assert np.array_equal(got[5], exp[5])
    # no receptor at all
    empty = _lib.DeviceBackend(ctx).select_receptors(ni, nj, nk, x, y, z, np.ones(npts), loc)
    assert empty.n == 0 and empty.indcell.size == 0


def test_committed_golden_fixtures(ctx):
    """The CUDA path against the committed fixtures of tests/golden (outputs of the reference's compiled code,
    written by tests/golden/make_golden.py): no oracle library involved."""
    import os
    from cassiopee_b200 import _lib
    from tests.golden.make_golden import golden_cases
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name, (dz, pts, kw) in golden_cases().items():
        kw = dict(kw)
        types = kw.pop("types", None) or [1] * len(dz)
        zones = [_lib.Zone(ctx, *z, interpDataType=t) for z, t in zip(dz, types)]
        pts = [np.ascontiguousarray(p) for p in pts]
        got = _lib.set_interp_data(ctx, pts[0], pts[1], pts[2], zones, **kw).fetch()
        ref = np.load(os.path.join(gold, name + ".npz"))
        for z in range(len(dz)):
            for q, key in enumerate(("rcv", "dnr", "typ", "cf", "ext")):
                assert np.array_equal(got[q][z], ref["%s%d" % (key, z)]), (name, key, z)
        assert np.array_equal(got[5], ref["orphan"]), name


@pytest.mark.parametrize("order", [2, 3])
def test_post_extract_mesh(ctx, order):
    """Post.extractMesh = getInterpolationCell(nature 0, penalty 0) -> getExtrapolationCell -> interpolated values of
    every donor field, zeros where nothing is found; against the same composition of the reference functions."""
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Post as P
    from oracle import oracle as O
    a = G.cart((0, 0, 0), (0.1, 0.1, 0.1), (11, 11, 11), name='A')
    b = G.cylinder((0.5, 0.5, 0.), 0.1, 0.45, 360., 0., 1., (33, 7, 9), name='B')
    names = ['F', 'G']
    for z in (a, b):
        C._initVars(z, 'F', lambda x, y, zz: x + 2 * y + 3 * zz)
        C._initVars(z, 'G', lambda x, y, zz: np.sin(3 * x) * np.cos(2 * y) + zz * zz)
    e = G.cart((-1.83, 0.21, 0.11), (0.17, 0.05, 0.09), (19, 13, 9), name='E')
    te = P.extractMesh(C.newPyTree(['D', a, b]), C.newPyTree(['E', e]), order=order)
    ze = Internal.getZones(te)[0]
    assert C.getFieldArray(e, 'F') is None                                   # the input mesh is untouched
    xe, ye, zz = [v.reshape(-1, order='F') for v in C.getCoords(e)]
    oz = []
    for z in (a, b):
        d = Internal.getZoneDim(z)
        oz.append(O.RefZone(d[1], d[2], d[3], *[v.reshape(-1, order='F') for v in C.getCoords(z)]))
    ref = O.set_interp_data(xe, ye, zz, oz, order, 0, 0, 1)
    exp = [np.zeros(xe.size) for _ in names]
    for noz, z in enumerate((a, b)):
        d = Internal.getZoneDim(z)
        O.transfer([C.getFieldArray(z, v).reshape(-1, order='F') for v in names], exp, d[1], d[2],
                   ref[0][noz], ref[1][noz], ref[2][noz], ref[3][noz])
    found = np.zeros(xe.size, bool)
    for noz in range(2):
        found[ref[0][noz]] = True
    assert found.sum() > 200 and (~found).sum() > 200
    for v, r in zip(names, exp):
        got = C.getFieldArray(ze, v).reshape(-1, order='F')
        assert np.array_equal(got, r), v
        assert (got[~found] == 0.).all()


def test_two_dimensional_donors_type22(ctx):
    """nk = 1 zones (z = constant): setInterpData (type 22 / type 1, extrapolation, holes, mixed with order 3
    requests that fall back to extrapolation) and type-22 transfers + strict cellN, bit-exact vs the reference."""
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    import cassiopee_b200.Generator as G
    a = G.cart((0, 0, 0.25), (0.1, 0.1, 1.), (11, 11, 1))
    b = G.cylinder((0.5, 0.5, 0.25), 0.1, 0.6, 360., 0., 1., (33, 9, 1))
    A = (11, 11, 1) + tuple(cases.flat(a)); B = (33, 9, 1) + tuple(cases.flat(b))
    rng = np.random.default_rng(0)
    P = rng.uniform(-0.3, 1.3, (3, 6000)); P[2] = 0.25
    P[2][:50] = 0.3
    P[0][50:250] = np.round(P[0][50:250], 1); P[1][50:250] = np.round(P[1][50:250], 1)
    ca = np.ones(A[3].size); ca[rng.random(ca.size) < 0.06] = 0.; ca[rng.random(ca.size) < 0.06] = 2.
    cb = np.ones(B[3].size); cb[rng.random(cb.size) < 0.06] = 0.; cb[rng.random(cb.size) < 0.06] = 2.
    got, ref, gz, oz = _compare(ctx, [A], P)
    typ = got[2][0]
    assert (typ == 22).sum() > 1000 and (typ == 1).sum() > 10 and got[4][0].size > 100
    for nature in (0, 1):
        for penalty in (0, 1):
            _compare(ctx, [A, B], P, 2, nature, penalty, 1, [ca, cb])
            _compare(ctx, [B, A], P, 2, nature, penalty, 0, [cb, ca])
    _compare(ctx, [A, B], P, 3, 1, 1, 1, [ca, cb])          # order 3 needs nk >= 3: everything is extrapolated (type 22)
    # transfers
    rcv, dnr, cf = got[0][0], got[1][0], got[3][0]
    nR = P[0].size
    fD = cases.fields(A[3], A[4], A[5], 5) + [A[3] + 2 * A[4]]
    fRg = [np.zeros(nR) for _ in fD]; fRo = [np.zeros(nR) for _ in fD]
    plan = _lib.Plan(ctx, 11, 11, 1, nR, dnr, rcv, typ, cf)
    plan.transfer(fD, fRg)
    O.transfer(fD, fRo, 11, 11, rcv, dnr, typ, cf)
    for x, y in zip(fRg, fRo):
        assert np.array_equal(x, y)
    assert np.array_equal(plan.transferD(fD), O.transferD(fD, 11, 11, dnr, typ, cf))
    ext = np.zeros(nR, bool); ext[got[4][0]] = True
    ok = np.zeros(nR, bool); ok[rcv] = True
    assert np.abs(fRg[-1] - (P[0] + 2 * P[1]))[ok & ~ext].max() < 1e-12
    cRg = np.full(nR, 2.); cRo = np.full(nR, 2.)
    plan.transfer_celln(ca, cRg)
    O.transfer_celln(ca, cRo, 11, 11, rcv, dnr, typ, cf)
    assert np.array_equal(cRg, cRo)


def test_two_dimensional_pytree_api(ctx):
    """X.setInterpData / setInterpTransfers on 2-D zones, receptors at centres (node2Center of quads)."""
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.PyTree as X
    a = G.cart((0, 0, 0), (0.1, 0.1, 1.), (21, 21, 1), name='A')
    b = G.cart((0.53, 0.41, 0), (0.07, 0.06, 1.), (9, 8, 1), name='B')
    t = C.newPyTree(['Base', a, b])
    C._initVars(t, 'centers:F', lambda x, y, z: x + 2 * y)
    C._initVars(t, 'centers:cellN', 1.)
    zb = Internal.getZones(t)[1]
    C._initVars(zb, 'centers:cellN', 2.)
    exact = C.getFieldArray(zb, 'centers:F').copy()
    C._initVars(zb, 'centers:F', 0.)
    tc = C.node2Center(t)
    tc = X.setInterpData(t, tc, order=2, nature=1, loc='centers', storage='inverse', itype='chimera', verbose=0)
    s = Internal.getNodeFromName1(Internal.getZones(tc)[0], 'ID_B')
    assert s is not None
    assert set(np.unique(Internal.getNodeFromName1(s, 'InterpolantsType')[1])) <= {1, 22}
    assert Internal.getNodeFromName1(s, 'PointListDonor')[1].size == 8 * 7
    t2 = X.setInterpTransfers(t, tc, variables=['F'], storage=1)
    got = C.getFieldArray(Internal.getZones(t2)[1], 'centers:F')
    assert np.abs(got - exact).max() < 1e-12


def _rotation_case(rcv_shuffle):
    """Cart pair with velocity + momentum fields on both sides and interp data in inverse storage."""
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.PyTree as X
    h = 1. / 16
    a = G.cart((0, 0, 0), (h, h, h), (17, 17, 17), name='A')
    b = G.cart((0.27, 0.22, 0.31), (0.7 * h, 0.7 * h, 0.7 * h), (11, 11, 11), name='B')
    t = C.newPyTree(['Base', a, b])
    names = ['Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature', 'MomentumX', 'MomentumY', 'MomentumZ']
    for i, v in enumerate(names):
        C._initVars(t, 'centers:' + v, lambda x, y, z, i=i: 1. + 0.3 * i + np.sin(3 * x + i) + 2. * y * y + 3. * z)
    C._initVars(t, 'centers:cellN', 1.)
    zb = Internal.getZones(t)[1]
    cn = np.ones(C.fieldShape(zb, 'centers'), order='F')
    if rcv_shuffle:
        cn[::2, 1::2, :] = 2.          # scattered receptors: host-scatter path of cx_transfer
    else:
        cn[...] = 2.                   # every cell: one ascending run, downloaded in place
    C._initVars(zb, 'centers:cellN', cn)
    tc = C.node2Center(t)
    tc = X.setInterpData(t, tc, order=2, nature=1, penalty=1, extrap=1, loc='centers', storage='inverse',
                         sameName=1, itype='chimera', verbose=0)
    return t, tc, names


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("rcv_shuffle", [False, True])
def test_periodic_rotation_after_transfer(ctx, axis, rcv_shuffle):
    """RotationAngle child of the ID_* subregion (setInterpTransfers.cpp:281-285, :468-472, includeTransfers.h):
    velocity and momentum triplets rotated at the receptors; against the reference's own fragment."""
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.PyTree as X
    from oracle import oracle as O
    t, tc, names = _rotation_case(rcv_shuffle)
    zdA = Internal.getZones(tc)[0]
    zb = Internal.getZones(t)[1]
    s = Internal.getNodeFromName1(zdA, 'ID_B')
    ang = np.zeros(3); ang[axis] = 0.37 * (axis + 1)
    Internal.createChild(s, 'RotationAngle', 'DataArray_t', ang)
    rcv = Internal.getNodeFromName1(s, 'PointListDonor')[1]
    dnr = Internal.getNodeFromName1(s, 'PointList')[1]
    typ = Internal.getNodeFromName1(s, 'InterpolantsType')[1]
    cf = Internal.getNodeFromName1(s, 'InterpolantsDonor')[1]
    assert rcv.size > 100
    fD = [C.getFieldArray(zdA, v) for v in names]
    fR = [np.array(C.getFieldArray(zb, 'centers:' + v), order='F', copy=True) for v in names]
    O.transfer(fD, fR, 16, 16, rcv, dnr, typ, cf)
    O.rotate(fR, rcv, ang, posv=(1, 2, 3), posm=(5, 6, 7))
    t2 = X.setInterpTransfers(t, tc, variables=names, storage=1)
    zb2 = Internal.getZones(t2)[1]
    for v, r in zip(names, fR):
        assert np.array_equal(C.getFieldArray(zb2, 'centers:' + v), r), v
    # device-resident variant: transfer, then the rotation alone at the plan's receptors
    from cassiopee_b200 import _lib
    plan = _lib.Plan(ctx, 16, 16, 16, fR[0].size, dnr, rcv, typ, cf)
    dD = [_lib.DeviceArray.from_host(ctx, np.ascontiguousarray(f.reshape(-1, order='F'))) for f in fD]
    f0 = [np.ascontiguousarray(C.getFieldArray(zb, 'centers:' + v).reshape(-1, order='F')) for v in names]
    dR = [_lib.DeviceArray.from_host(ctx, f) for f in f0]
    plan.transfer_dev([d.ptr for d in dD], [d.ptr for d in dR])
    plan.rotate_dev(axis + 1, ang[axis], dR[1].ptr, dR[2].ptr, dR[3].ptr)
    plan.rotate_dev(axis + 1, ang[axis], dR[5].ptr, dR[6].ptr, dR[7].ptr)
    ctx.sync()
    for d, r in zip(dR, fR):
        assert np.array_equal(d.download(), r.reshape(-1, order='F'))


def test_compact_fields_transfer(ctx):
    """_setInterpTransfers(compact=1) (setInterpTransfers.cpp:474-497, getfromzone{D,R}compact.h): 5 fields
    (varType 1..3) or 6 lying one behind the other behind variables[0]; same arithmetic as the general path."""
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.PyTree as X
    from oracle import oracle as O
    t, tc, _ = _rotation_case(True)
    zdA = Internal.getZones(tc)[0]
    zb = Internal.getZones(t)[1]
    s = Internal.getNodeFromName1(zdA, 'ID_B')
    rcv = Internal.getNodeFromName1(s, 'PointListDonor')[1]
    dnr = Internal.getNodeFromName1(s, 'PointList')[1]
    typ = Internal.getNodeFromName1(s, 'InterpolantsType')[1]
    cf = Internal.getNodeFromName1(s, 'InterpolantsDonor')[1]
    for varType, nv in ((2, 5), (21, 6)):
        names = ['Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature', 'MomentumX'][:nv]
        # compact the receptor fields (FlowSolution#Centers of B) and the donor fields (FlowSolution of tc's A)
        for z, cont, shp in ((zb, Internal.__FlowSolutionCenters__, (10, 10, 10)), (zdA, Internal.__FlowSolutionNodes__, (16, 16, 16))):
            c = Internal.getNodeFromName1(z, cont)
            block = np.empty(shp + (nv,), order='F')
            for i, v in enumerate(names):
                n = Internal.getNodeFromName1(c, v)
                block[..., i] = n[1]
                n[1] = block[..., i]
        fD = [C.getFieldArray(zdA, v) for v in names]
        fR = [np.array(C.getFieldArray(zb, 'centers:' + v), order='F', copy=True) for v in names]
        O.transfer(fD, fR, 16, 16, rcv, dnr, typ, cf)
        X._setInterpTransfers(t, tc, variables=['Density'], storage=1, compact=1, varType=varType)
        for v, r in zip(names, fR):
            assert np.array_equal(C.getFieldArray(zb, 'centers:' + v), r), v
    # donor-side variant (setInterpTransfersD.cpp:282-296): same 6 compacted fields, dense (nvars, nrcv) array
    infos = X._setInterpTransfersD(tc, variables=['Density'], compact=1, varType=21)
    info = [i for i in infos if i[0] == 'B'][0]
    refD = O.transferD(fD, 16, 16, dnr, typ, cf)
    assert info[1][0] == 'Density' and np.array_equal(info[1][1], refD) and np.array_equal(info[2], rcv)
    # non-compacted trees raise instead of reading out of bounds
    t3, tc3, _ = _rotation_case(True)
    with pytest.raises(ValueError):
        X._setInterpTransfers(t3, tc3, variables=['Density'], storage=1, compact=1)


@pytest.mark.parametrize("dims,density", [((37, 23, 9), 4.0), ((64, 64, 6), 2.0), ((41, 29, 7), 0.02)])
def test_transfer_pipelined_kernel(ctx, dims, density, monkeypatch):
    """Persistent pipelined type-2 kernel (k_transfer2p, experimental) forced on small plans: odd row pitch, a type-1 group in
    front of the type-2 group (index lists not 16-byte aligned), partial last tile, dense spans (rows staged) and
    sparse spans (direct gather), 1..8 variables (9 falls back to the per-tile kernel).  Bit-exact vs the
    reference loops, in place and dense."""
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    monkeypatch.setenv("CX_TRANSFER_PIPE", "1")
    monkeypatch.setenv("CX_PIPE_MIN_TILES", "1")
    ni, nj, nk = dims
    rng = np.random.default_rng(ni * 100 + nk)
    nD = ni * nj * nk
    n = int(density * nD) + 3
    nR = n + 11
    typ = rng.choice(np.array([1, 2, 2, 2, 2, 2, 2, 2], np.int32), n).astype(np.int32)
    typ[:5] = 1                                   # type-2 group starts at an odd offset
    o = np.where(typ == 1, 1, 2)
    i = (rng.random(n) * (ni - o + 1)).astype(np.int64); j = (rng.random(n) * (nj - o + 1)).astype(np.int64)
    k = (rng.random(n) * (nk - o + 1)).astype(np.int64)
    i[5:8] = ni - 2; j[5:8] = nj - 2; k[5:8] = nk - 2      # stencil ends on the last node of the zone
    dnr = (i + ni * j + ni * nj * k).astype(np.int32)
    rcv = rng.permutation(nR)[:n].astype(np.int32)
    ncf = np.where(typ == 1, 1, 8)
    cf = rng.standard_normal(int(ncf.sum()))
    plan = _lib.Plan(ctx, ni, nj, nk, nR, dnr, rcv, typ, cf)
    for nvars in (1, 5, 7, 8, 9):
        fD = [rng.standard_normal(nD) for _ in range(nvars)]
        fRo = [np.full(nR, -7.) for _ in range(nvars)]
        O.transfer(fD, fRo, ni, nj, rcv, dnr, typ, cf)
        dD = [_lib.DeviceArray.from_host(ctx, f) for f in fD]
        dR = [_lib.DeviceArray.from_host(ctx, np.full(nR, -7.)) for _ in range(nvars)]
        plan.transfer_dev([d.ptr for d in dD], [d.ptr for d in dR])
        ctx.sync()
        for d, b in zip(dR, fRo):
            assert np.array_equal(d.download(), b)
        dense = _lib.DeviceArray(ctx, n * nvars)
        plan.transferD_dev([d.ptr for d in dD], dense.ptr, n)
        ctx.sync()
        assert np.array_equal(dense.download().reshape(nvars, n), O.transferD(fD, ni, nj, dnr, typ, cf))


@pytest.mark.parametrize("order", [2, 3, 5])
@pytest.mark.parametrize("loc,storage", [('nodes', 'direct'), ('centers', 'direct'), ('centers', 'inverse'), ('nodes', 'inverse')])
def test_reference_script_setInterpTransfersPT_t1_t2(ctx, order, loc, storage):
    """The reference's own known-answer scripts Connector/test/setInterpTransfersPT_t1.py / _t2.py through the
    PyTree API: an 11^3 Cartesian donor with F = x + 2y + 3z at its nodes, an identical grid as receptor (all of
    its nodes / centres have cellN = 2), orders 2, 3 and 5, direct and inverse storage; every interpolated value
    reproduces the linear field, and lists / values equal the oracle's (= the reference's K_INTERP + transfer loop)."""
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.PyTree as X
    from oracle import oracle as O
    ni = nj = nk = 11
    h = 10. / (ni - 1)
    F = lambda x, y, z: x + 2. * y + 3. * z
    m = G.cart((0, 0, 0), (h, h, 1), (ni, nj, nk), name='donor')
    a = G.cart((0, 0, 0), (h, h, 1), (ni, nj, nk), name='extraction')
    tD = C.newPyTree(['Dnr', m]); tR = C.newPyTree(['Rcv', a])
    C._initVars(tD, 'F', F)
    fld = 'F' if loc == 'nodes' else 'centers:F'
    C._initVars(tR, 'cellN' if loc == 'nodes' else 'centers:cellN', 2.)
    res = X.setInterpData(tR, tD, order=order, loc=loc, storage=storage, verbose=0)
    if storage == 'direct': tR = res
    else: tD = res
    C._initVars(tR, fld, 0.)
    tR2 = X.setInterpTransfers(tR, tD, variables=['F'])
    zr = Internal.getZones(tR2)[0]; zd = Internal.getZones(tD)[0]
    got = C.getFieldArray(zr, fld).reshape(-1, order='F')
    # receptor points as the reference takes them (nodes, or node2Center of the receptor zone)
    xr, yr, zrr = C.getCoords(Internal.getZones(tR)[0])
    if loc == 'centers':
        xr, yr, zrr = [C.node2CenterArray(v) for v in (xr, yr, zrr)]
    pts = [np.ascontiguousarray(v.reshape(-1, order='F')) for v in (xr, yr, zrr)]
    xd, yd, zdd = [v.reshape(-1, order='F') for v in C.getCoords(zd)]
    # the tree layer hands the donor over with cellN = 1 added (OversetData.py:1189-1193)
    ref = O.set_interp_data(pts[0], pts[1], pts[2], [O.RefZone(ni, nj, nk, xd, yd, zdd, np.ones(xd.size))],
                            order=order, nature=0, penalty=1, extrap=1)
    holder = Internal.getZones(tR)[0] if storage == 'direct' else zd
    s = Internal.getNodeFromName1(holder, 'ID_donor' if storage == 'direct' else 'ID_extraction')
    assert s is not None
    pl = Internal.getNodeFromName1(s, 'PointList')[1]; pld = Internal.getNodeFromName1(s, 'PointListDonor')[1]
    rcv, dnr = (pl, pld) if storage == 'direct' else (pld, pl)
    assert np.array_equal(rcv, ref[0][0]) and np.array_equal(dnr, ref[1][0])
    assert np.array_equal(Internal.getNodeFromName1(s, 'InterpolantsType')[1], ref[2][0])
    assert np.array_equal(Internal.getNodeFromName1(s, 'InterpolantsDonor')[1], ref[3][0])
    assert rcv.size > 0.9 * got.size
    fD = [np.ascontiguousarray(C.getFieldArray(zd, 'F').reshape(-1, order='F'))]
    fR = [np.zeros(got.size)]
    O.transfer(fD, fR, ni, nj, ref[0][0], ref[1][0], ref[2][0], ref[3][0])
    assert np.array_equal(got, fR[0])
    exact = F(*pts)
    # orders 2 and 3 reproduce the linear field to rounding; the reference's order-5 weights carry its 12-digit
    # literals (inv24 = 0.041666666667, Interp2.cpp:640-658): its own error on this case is 1.4e-9
    assert np.abs(got - exact)[rcv].max() < (1e-11 if order < 5 else 1e-8)


@pytest.mark.parametrize("loc", ['nodes', 'centers'])
@pytest.mark.parametrize("order", [2, 3, 5])
def test_reference_script_setInterpDataPT_t1_curve_receptors(ctx, loc, order):
    """Connector/test/setInterpDataPT_t1.py on the device: circle (curve zone) receptors at nodes and at centres,
    both storages, penalty and nature sweeps, orders 2 / 3 / 5; lists identical to the oracle's."""
    for storage in ('direct', 'inverse'):
        for pen in (0, 1):
            for nat in (0, 1):
                cases.check_curve_receptors(dict(ctx=ctx), order, loc, storage, pen, nat)


def _tetra_like_coefs(rng, typ):
    """Coefficient list whose type-2 entries look like the reference's tetrahedron weights (InterpData.cpp:621-631:
    at most 4 distinct doubles among the 8), with degenerate ones (1, 2, 3 distinct values, +0.0 next to -0.0) and
    ~20 % of entries with 8 distinct values (what an extrapolated entry may hold) mixed in."""
    out = []
    for t in typ:
        if t == 1:
            out.append(rng.standard_normal(1)); continue
        r = rng.random()
        if r < 0.2:
            out.append(rng.standard_normal(8)); continue
        nd = 4 if r < 0.8 else int(rng.integers(1, 4))
        vals = rng.standard_normal(nd)
        if r > 0.97:
            vals[0] = 0.
            if nd > 1: vals[1] = -0.
        pick = rng.integers(0, nd, 8)
        pick[:nd] = rng.permutation(nd)
        out.append(vals[pick])
    return np.concatenate(out)


@pytest.mark.parametrize("layout", ["1", "5"])
def test_transfer_packed_type2_groups(ctx, layout, monkeypatch):
    """Packed type-2 groups (4 distinct doubles + 16-bit selector per entry, cx_transfer.cu:load_cf2) next to raw
    ones in the same plan: every launch flavour (single plan in place / dense, plan set in its three output modes,
    strict cellN) bit-exact vs the reference loops, and identical to the same plan built with CX_PLAN_PACK=0."""
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    monkeypatch.setenv("CX_PLAN_SORT", layout)
    ni, nj, nk = 67, 29, 23
    rng = np.random.default_rng(77 + int(layout))
    nD = ni * nj * nk
    cells = np.array([(i, j, k) for k in range(1, nk - 1) for j in range(0, nj - 1) for i in (0, 1, 2, 40, 41, 64, 65)], np.int64)
    cells = cells[rng.permutation(cells.shape[0])]
    n = cells.shape[0]
    typ = np.full(n, 2, np.int32); typ[rng.random(n) < 0.05] = 1
    dnr = (cells[:, 0] + ni * cells[:, 1] + ni * nj * cells[:, 2]).astype(np.int32)
    nR = n + 9
    rcv = np.sort(rng.permutation(nR)[:n]).astype(np.int32)
    cf = _tetra_like_coefs(rng, typ)
    monkeypatch.setenv("CX_PLAN_PACK", "1")
    plan = _lib.Plan(ctx, ni, nj, nk, nR, dnr, rcv, typ, cf)
    monkeypatch.setenv("CX_PLAN_PACK", "0")
    plan_raw = _lib.Plan(ctx, ni, nj, nk, nR, dnr, rcv, typ, cf)
    monkeypatch.delenv("CX_PLAN_PACK")
    assert plan.info()['ngroups'] == plan_raw.info()['ngroups'] + 1          # one more group: packed + raw type 2
    for nvars in (1, 5, 7, 16):
        fD = [rng.standard_normal(nD) for _ in range(nvars)]
        fRo = [np.full(nR, -7.) for _ in range(nvars)]
        O.transfer(fD, fRo, ni, nj, rcv, dnr, typ, cf)
        refD = O.transferD(fD, ni, nj, dnr, typ, cf)
        dD = [_lib.DeviceArray.from_host(ctx, f) for f in fD]
        for pl in (plan, plan_raw):
            dR = [_lib.DeviceArray.from_host(ctx, np.full(nR, -7.)) for _ in range(nvars)]
            pl.transfer_dev([d.ptr for d in dD], [d.ptr for d in dR])
            ctx.sync()
            for d, b in zip(dR, fRo):
                assert np.array_equal(d.download(), b)
            dense = _lib.DeviceArray(ctx, n * nvars)
            pl.transferD_dev([d.ptr for d in dD], dense.ptr, n)
            ctx.sync()
            assert np.array_equal(dense.download().reshape(nvars, n), refD)
            dR2 = [_lib.DeviceArray.from_host(ctx, np.full(nR, -7.)) for _ in range(nvars)]
            dense1 = _lib.DeviceArray(ctx, n * nvars); dense2 = _lib.DeviceArray(ctx, n * nvars)
            ps = _lib.PlanSet(ctx, nvars, [(pl, 0, [d.ptr for d in dD], [d.ptr for d in dR2], None, 0),
                                           (pl, 1, [d.ptr for d in dD], None, dense1.ptr, n),
                                           (pl, 2, [d.ptr for d in dD], None, dense2.ptr, n)])
            ps.run(); ctx.sync()
            for d, b in zip(dR2, fRo):
                assert np.array_equal(d.download(), b)
            assert np.array_equal(dense1.download().reshape(nvars, n), refD)
            srt = pl.sorted_rcv()
            d2 = dense2.download().reshape(nvars, n)
            for v in range(nvars):
                got = np.full(nR, -7.); got[srt] = d2[v]
                assert np.array_equal(got, fRo[v])
    cD = np.ones(nD); cD[rng.random(nD) < 0.2] = 0.; cD[rng.random(nD) < 0.2] = 2.
    cRg = np.full(nR, 2.); cRo = np.full(nR, 2.)
    plan.transfer_celln(cD, cRg)
    O.transfer_celln(cD, cRo, ni, nj, rcv, dnr, typ, cf)
    assert np.array_equal(cRg, cRo)


@pytest.mark.parametrize("shape", ["xslab", "yslab", "zslab", "volume", "sparse"])
def test_transfer_box_layout(ctx, shape, monkeypatch):
    """Box layout of the type-2 groups (k_transfer2b / the layout-3 branch of k_transfer_set, forced with
    CX_PLAN_SORT=5): fringe slabs 3 cells thick normal to each axis (the C4 shape: odd row pitch 99), a dense volume
    and a sparse cloud, a type-1 group in front, stencils ending on the last node, 1..16 variables; single-plan
    launches (in place, dense) and the fused plan-set launch in its three output modes.  Bit-exact vs the
    reference loops."""
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    monkeypatch.setenv("CX_PLAN_SORT", "5")
    ni, nj, nk = (99, 41, 37) if shape != "volume" else (45, 33, 21)
    rng = np.random.default_rng(len(shape) * 7 + ni)
    nD = ni * nj * nk
    if shape == "xslab":
        cells = [(i, j, k) for k in range(2, nk - 3) for j in range(1, nj - 1) for i in (91, 92, 93)]
    elif shape == "yslab":
        cells = [(i, j, k) for k in range(0, nk - 1) for j in (5, 6, 7) for i in range(0, ni - 1)]
    elif shape == "zslab":
        cells = [(i, j, k) for k in (nk - 4, nk - 3, nk - 2) for j in range(0, nj - 1) for i in range(3, ni - 1)]
    elif shape == "volume":
        cells = [(i, j, k) for k in range(nk - 1) for j in range(nj - 1) for i in range(ni - 1)]
    else:
        cells = [tuple(c) for c in np.stack([rng.integers(0, ni - 1, 3000), rng.integers(0, nj - 1, 3000),
                                             rng.integers(0, nk - 1, 3000)], axis=1)]
    cells = np.array(cells, np.int64)
    rep = 2 if shape != "sparse" else 1
    cells = np.repeat(cells, rep, axis=0)
    cells = cells[rng.permutation(cells.shape[0])]
    n = cells.shape[0]
    typ = np.full(n, 2, np.int32)
    typ[rng.random(n) < 0.05] = 1
    typ[:3] = 1
    dnr = (cells[:, 0] + ni * cells[:, 1] + ni * nj * cells[:, 2]).astype(np.int32)
    dnr[3] = (ni - 2) + ni * (nj - 2) + ni * nj * (nk - 2); typ[3] = 2          # stencil ends on the last node
    nR = n + 17
    rcv = np.sort(rng.permutation(nR)[:n]).astype(np.int32)
    ncf = np.where(typ == 1, 1, 8)
    cf = rng.standard_normal(int(ncf.sum()))
    plan = _lib.Plan(ctx, ni, nj, nk, nR, dnr, rcv, typ, cf)
    for nvars in (1, 5, 7, 16):
        fD = [rng.standard_normal(nD) for _ in range(nvars)]
        fRo = [np.full(nR, -7.) for _ in range(nvars)]
        O.transfer(fD, fRo, ni, nj, rcv, dnr, typ, cf)
        refD = O.transferD(fD, ni, nj, dnr, typ, cf)
        dD = [_lib.DeviceArray.from_host(ctx, f) for f in fD]
        dR = [_lib.DeviceArray.from_host(ctx, np.full(nR, -7.)) for _ in range(nvars)]
        plan.transfer_dev([d.ptr for d in dD], [d.ptr for d in dR])
        ctx.sync()
        for d, b in zip(dR, fRo):
            assert np.array_equal(d.download(), b)
        dense = _lib.DeviceArray(ctx, n * nvars)
        plan.transferD_dev([d.ptr for d in dD], dense.ptr, n)
        ctx.sync()
        assert np.array_equal(dense.download().reshape(nvars, n), refD)
        dR2 = [_lib.DeviceArray.from_host(ctx, np.full(nR, -7.)) for _ in range(nvars)]
        dense1 = _lib.DeviceArray(ctx, n * nvars); dense2 = _lib.DeviceArray(ctx, n * nvars)
        ps = _lib.PlanSet(ctx, nvars, [(plan, 0, [d.ptr for d in dD], [d.ptr for d in dR2], None, 0),
                                       (plan, 1, [d.ptr for d in dD], None, dense1.ptr, n),
                                       (plan, 2, [d.ptr for d in dD], None, dense2.ptr, n)])
        ps.run(); ctx.sync()
        for d, b in zip(dR2, fRo):
            assert np.array_equal(d.download(), b)
        assert np.array_equal(dense1.download().reshape(nvars, n), refD)
        srt = plan.sorted_rcv()
        d2 = dense2.download().reshape(nvars, n)
        for v in range(nvars):
            got = np.full(nR, -7.); got[srt] = d2[v]
            assert np.array_equal(got, fRo[v])


def test_set_interp_transfers_batch_keeps_sequential_semantics(ctx):
    """_setInterpTransfers uploads every donor array once per call (cx_batch_begin/end).  Three node-located zones in
    ONE tree (donor tree == receptor tree): receptor fields are donor fields of the next subregions, so the loop is
    sequential like the reference's (a subregion reads what the previous ones wrote).  Checked against the reference
    loops applied in the same subregion order, and the upload count is the number of (zone, variable) arrays + the
    re-uploads forced by the writes, not subregions x variables."""
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.PyTree as X
    from oracle import oracle as O
    h = 0.05
    za = G.cart((0, 0, 0), (h, h, h), (21, 17, 13), name='A')
    zb = G.cart((0.31, 0.12, 0.07), (0.9 * h, 0.9 * h, 0.9 * h), (19, 15, 11), name='B')
    zc = G.cart((0.52, 0.22, 0.11), (1.1 * h, 1.1 * h, 1.1 * h), (15, 13, 9), name='C')
    t = C.newPyTree(['Base', za, zb, zc])
    names = ['Density', 'MomentumX', 'EnergyStagnationDensity']
    rng = np.random.default_rng(5)
    for z in Internal.getZones(t):
        shp = C.fieldShape(z, 'nodes')
        for v in names:
            C._initVars(z, v, np.asfortranarray(rng.standard_normal(shp)))
        c = np.ones(shp, order='F')
        c[rng.random(shp) < 0.3] = 2.
        C._initVars(z, 'cellN', c)
    t = X.setInterpData(t, t, order=2, nature=0, penalty=0, extrap=1, loc='nodes', storage='inverse', sameName=1,
                        itype='chimera', verbose=0)
    before = {(z[0], v): np.array(C.getFieldArray(z, v), order='F', copy=True) for z in Internal.getZones(t) for v in names}
    nsub = sum(len([s for s in Internal.getNodesFromType1(z, 'ZoneSubRegion_t') if s[0].startswith('ID_')])
               for z in Internal.getZones(t))
    assert nsub >= 4
    l0 = ctx.launches()
    t2 = X.setInterpTransfers(t, t, variables=names, storage=1)
    # reference loops, same order: donor zones in tree order, their subregions in node order (OversetData.py:2036-2098)
    cur = {k: v.copy(order='F') for k, v in before.items()}
    for zd in Internal.getZones(t):
        _, imd, jmd, kmd, _ = Internal.getZoneDim(zd)
        for s in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
            if not s[0].startswith('ID_'):
                continue
            zr = Internal.getValue(s)
            fD = [cur[(zd[0], v)].reshape(-1, order='F') for v in names]
            fR = [cur[(zr, v)].reshape(-1, order='F') for v in names]
            O.transfer(fD, fR, imd, jmd, Internal.getNodeFromName1(s, 'PointListDonor')[1],
                       Internal.getNodeFromName1(s, 'PointList')[1], Internal.getNodeFromName1(s, 'InterpolantsType')[1],
                       Internal.getNodeFromName1(s, 'InterpolantsDonor')[1])
    changed = 0
    for z in Internal.getZones(t2):
        for v in names:
            got = C.getFieldArray(z, v)
            assert np.array_equal(got, cur[(z[0], v)]), (z[0], v)
            changed += int(not np.array_equal(got, before[(z[0], v)]))
    assert changed >= 6 and ctx.launches() > l0


@pytest.mark.gpu
def test_transfer_session_sparse_download_and_c_ordered_fields(ctx, monkeypatch):
    """TransferSession.upload_donor_fields / step / download on host trees: receptor zones of which the step writes
    only a fringe are downloaded sparsely (device gather + index store on the host), receptor or donor arrays that are
    C-contiguous 3-D arrays go through the Fortran-order staging path (raw copies would permute them), and both give
    what _setInterpTransfers gives on the same trees; the untouched receptor values keep their host content."""
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.PyTree as X
    import cassiopee_b200.Connector.Mpi as Xmpi
    h = 1. / 16
    def build():
        a = G.cart((0, 0, 0), (h, h, h), (17, 15, 13), name='A')
        b = G.cart((0.31, 0.22, 0.18), (0.8 * h, 0.8 * h, 0.8 * h), (15, 13, 11), name='B')
        t = C.newPyTree(['Base', a, b])
        names = ['Density', 'MomentumX', 'EnergyStagnationDensity']
        rng = np.random.default_rng(3)
        for z in Internal.getZones(t):
            shp = C.fieldShape(z, 'centers')
            for v in names:
                C._initVars(z, 'centers:' + v, np.asfortranarray(rng.standard_normal(shp)))
        C._initVars(t, 'centers:cellN', 1.)
        za, zb = Internal.getZones(t)
        cb = np.ones(C.fieldShape(zb, 'centers'), order='F'); cb[:2] = 2.; cb[:, :, -2:] = 2.     # B: a fringe only
        C._initVars(zb, 'centers:cellN', cb)
        ca = np.ones(C.fieldShape(za, 'centers'), order='F'); ca[2:, 1:, 1:] = 2.                  # A: most of the zone
        C._initVars(za, 'centers:cellN', ca)
        tc = C.node2Center(t)
        tc = X.setInterpData(t, tc, order=2, nature=0, penalty=1, extrap=1, loc='centers', storage='inverse',
                             sameName=1, itype='chimera', verbose=0)
        return t, tc, names
    t, tc, names = build()
    want = X.setInterpTransfers(t, tc, variables=names, storage=1)
    for c_order in (False, True):
        t, tc, names = build()
        if c_order:
            for tree, pre in ((t, 'centers:'), (tc, '')):
                for z in Internal.getZones(tree):
                    cont = Internal.getNodeFromName1(z, Internal.__FlowSolutionCenters__ if pre else Internal.__FlowSolutionNodes__)
                    node = Internal.getNodeFromName1(cont, names[1])
                    node[1] = np.ascontiguousarray(node[1])              # same values, C-ordered memory
                    assert not node[1].flags.f_contiguous
        sess = Xmpi.TransferSession(t, tc, names, comm=None, ctx=ctx)
        assert ('B' in sess._sparse) == (not c_order) and 'A' not in sess._sparse
        before = {z[0]: [C.getFieldArray(z, 'centers:' + v).copy() for v in names] for z in Internal.getZones(t)}
        sess.upload_donor_fields(); sess.step(); nb = sess.download()
        for z, zw in zip(Internal.getZones(t), Internal.getZones(want)):
            for v in names:
                assert np.array_equal(C.getFieldArray(z, 'centers:' + v), C.getFieldArray(zw, 'centers:' + v)), (z[0], v, c_order)
        if not c_order:
            full = sum(a.nbytes for z in Internal.getZones(t) for a in before[z[0]])
            assert nb < full                                             # B's fields were not downloaded whole
        sess.close()
        # the same iteration as ONE host-to-host call (sparse uploads of the donor points the plans read, per-variable
        # pipeline over three streams), pipelined or not, on fresh trees
        for pipelined in (True, False):
            t, tc, names = build()
            if c_order:
                node = Internal.getNodeFromName1(Internal.getNodeFromName1(Internal.getZones(tc)[0], Internal.__FlowSolutionNodes__), names[1])
                node[1] = np.ascontiguousarray(node[1])
            sess = Xmpi.TransferSession(t, tc, names, comm=None, ctx=ctx)
            # poison the device copies of the donor fields: only what transfer_host uploads may be read
            for lst in sess.dD.values():
                for d in lst: d.upload(np.full(d.n, np.nan))
            up, down = sess.transfer_host(pipelined=pipelined)
            for z, zw in zip(Internal.getZones(t), Internal.getZones(want)):
                for v in names:
                    assert np.array_equal(C.getFieldArray(z, 'centers:' + v), C.getFieldArray(zw, 'centers:' + v)), (z[0], v, c_order, pipelined)
            if not c_order:
                fullD = sum(C.getFieldArray(z, v).nbytes for z in Internal.getZones(tc) for v in names)
                assert 0 < up < fullD and 0 < down < full      # B reads a part of A only; B is written on a fringe only
            up2, down2 = sess.transfer_host(pipelined=pipelined)         # idempotent (donor fields live in tc)
            assert (up2, down2) == (up, down)
            for z, zw in zip(Internal.getZones(t), Internal.getZones(want)):
                for v in names:
                    assert np.array_equal(C.getFieldArray(z, 'centers:' + v), C.getFieldArray(zw, 'centers:' + v))
            sess.close()


# ---- cellN preparers either side of the path, device vs the reference's routines ----------
@pytest.mark.parametrize("shape", [(23, 17, 11), (40, 9, 1), (1, 1, 1), (97, 101, 33)])
def test_hole_fringe_matches_reference(ctx, shape):
    """cx_hole_fringe against K_CONNECTOR::searchMaskInterpolatedCellsStructOpt (getInterpolatedPoints.cpp:370-,
    compiled unchanged): cross and star stencils, depths 1-3, 2-D and 3-D, depth < 0 (Connector.py:176-182)."""
    from cassiopee_b200 import _lib
    from tests.backends import OracleBackend
    rng = np.random.default_rng(sum(shape))
    be = OracleBackend()
    for dir_ in (0, 1):
        for depth in (1, 2, 3, -1, -2):
            c = np.ones(shape, order='F')
            r = rng.random(shape)
            c[r < 0.02] = 0.; c[(r > 0.02) & (r < 0.05)] = 2.; c[(r > 0.05) & (r < 0.06)] = 1. + 1e-16
            got = c.copy(order='F'); ref = c.copy(order='F')
            _lib.hole_fringe(ctx, got, depth, dir_)
            be.hole_fringe(ref, depth, dir_)
            assert np.array_equal(got, ref), (shape, dir_, depth)
            if depth > 0 and c.size > 1: assert (got != c).any()
    with pytest.raises(NotImplementedError):
        _lib.hole_fringe(ctx, np.ones(shape, order='F'), 2, 2)


def test_apply_bc_overlaps_matches_reference(ctx):
    """cx_apply_bc_overlaps against the statements of applyBCOverlaps.cpp:61-112 (oracle/ref_driver.cpp), nodes and
    centres, every face, depths up to the zone size, blanked cells kept, invalid windows refused."""
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    rng = np.random.default_rng(8)
    im, jm, km = 14, 9, 6
    for loc in ('nodes', 'centers'):
        e = 0 if loc == 'nodes' else 1                      # centres: the window is given on the node grid
        faces = [((1, 1, 1), (1, jm + e, km + e)), ((im + e, 1, 1), (im + e, jm + e, km + e)),
                 ((1, 1, 1), (im + e, 1, km + e)), ((1, jm + e, 1), (im + e, jm + e, km + e)),
                 ((1, 1, 1), (im + e, jm + e, 1)), ((1, 1, km + e), (im + e, jm + e, km + e)),
                 ((3, jm + e, 2), (8, jm + e, 5)), ((im + e, 2, 2), (im + e, 5, 4))]
        for depth in (1, 2, 3, 20):
            for val in (2, 3):
                c = np.ones((im, jm, km), order='F'); c[rng.random(c.shape) < 0.1] = 0.
                for w in faces:
                    got = c.copy(order='F'); ref = c.copy(order='F')
                    _lib.apply_bc_overlaps(ctx, got, [w], depth, loc, val)
                    O.apply_bc_overlap(ref, w[0], w[1], depth, loc, val)
                    assert np.array_equal(got, ref), (loc, depth, w)
                got = c.copy(order='F'); ref = c.copy(order='F')
                _lib.apply_bc_overlaps(ctx, got, faces, depth, loc, val)
                for w in faces: O.apply_bc_overlap(ref, w[0], w[1], depth, loc, val)
                assert np.array_equal(got, ref)
        bad = ((0, 1, 1), (1, jm, km))
        c = np.ones((im, jm, km), order='F')
        with pytest.raises(TypeError): _lib.apply_bc_overlaps(ctx, c, [bad], 2, loc, 2)
        with pytest.raises(TypeError): O.apply_bc_overlap(c, bad[0], bad[1], 2, loc, 2)
        assert (c == 1.).all()


def test_preparers_through_the_tree_layer(ctx):
    """applyBCOverlaps + setHoleInterpolatedPoints (device backend) give the tree the oracle backend gives, and the
    cellN they leave drives setInterpData."""
    import cassiopee_b200.Connector.PyTree as XP
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Internal as Internal
    from tests.backends import OracleBackend
    a = G.cart((0, 0, 0), (0.1, 0.1, 0.1), (21, 21, 21), name='A')
    b = G.cart((1.45, 0.22, 0.31), (0.1, 0.1, 0.1), (18, 14, 12), name='B')
    t = C.newPyTree(['Base', a, b])
    zA, zB = Internal.getZones(t)
    C._addBC2Zone(zA, 'ov1', 'BCOverlap', [21, 21, 1, 21, 1, 21])
    C._addBC2Zone(zB, 'ov1', 'BCOverlap', [1, 1, 1, 14, 1, 12])
    cB = np.ones((17, 13, 11), order='F'); cB[8:11, 5:8, 4:7] = 0.
    C._initVars(zB, 'centers:cellN', cB)
    out = []
    for be in (None, OracleBackend()):
        t2 = XP.applyBCOverlaps(t, depth=2, loc='centers', backend=be)
        t2 = XP.setHoleInterpolatedPoints(t2, depth=2, dir=1, loc='centers', backend=be)
        out.append([C.getFieldArray(z, 'centers:cellN') for z in Internal.getZones(t2)])
    for g, r in zip(*out):
        assert np.array_equal(g, r)
    assert (out[0][1] == 2.).sum() > 2 * 13 * 11 and (out[0][1] == 0.).sum() == 27


def test_batched_hooks_build_the_same_trees(ctx):
    """DeviceBackend.make_zones (uploads zone after zone, every ADT build started on a side stream while the next zone
    goes up, a long first batch of levels whatever the zone size) gives node for node the tree Zone() builds alone:
    a zone above the 2^21-cell threshold of the one-level-per-round-trip regime, small zones, a 2-D zone, a Cartesian
    (InterpCart) zone in the middle of the batch."""
    from cassiopee_b200 import _lib
    rng = np.random.default_rng(21)

    def grid(ni, nj, nk, h, o, wobble):
        i, j, k = np.meshgrid(np.arange(ni), np.arange(nj), np.arange(nk), indexing='ij')
        x = o[0] + h * i + wobble * np.sin(0.7 * j + 0.3 * k); y = o[1] + h * j + wobble * np.sin(0.5 * k + 0.2 * i)
        z = o[2] + h * k + (wobble * np.sin(0.9 * i + 0.4 * j) if nk > 1 else 0.)
        return [a.reshape(-1, order='F') for a in (x, y, z)]
    specs = []
    for (ni, nj, nk, ty) in ((131, 131, 131, 1), (23, 19, 17, 1), (40, 33, 1, 1), (21, 21, 21, 0), (64, 48, 40, 1)):
        x, y, z = grid(ni, nj, nk, 0.1, rng.random(3), 0.02 if ty == 1 else 0.)
        c = np.ones(ni * nj * nk); c[rng.random(c.size) < 0.05] = 0.
        specs.append((ni, nj, nk, x, y, z, c, ty))
    assert (131 - 1) ** 3 > 1 << 21
    batch = _lib.DeviceBackend(ctx).make_zones(specs)
    for sp, zb in zip(specs, batch):
        alone = _lib.Zone(ctx, *sp[:7], interpDataType=sp[7])
        ia, ib = alone.info(), zb.info()
        assert ia["depth"] == ib["depth"] and np.array_equal(ia["bbox"], ib["bbox"])
        if sp[7] == 1:
            assert np.array_equal(alone.export_adt().view(np.uint8), zb.export_adt().view(np.uint8))
    # and they answer a search like the ones built alone
    pts = [rng.random(5000) * 2. for _ in range(3)]
    ra = _lib.set_interp_data(ctx, *pts, [batch[1], batch[4]], 2, 1, 1, 1).fetch()
    rb = _lib.set_interp_data(ctx, *pts, [_lib.Zone(ctx, *specs[1][:7]), _lib.Zone(ctx, *specs[4][:7])], 2, 1, 1, 1).fetch()
    for k in range(5):
        for a, b in zip(ra[k], rb[k]): assert np.array_equal(a, b)


def test_transfer_session_box_upload(ctx):
    """transfer_host() when the plans of a donor zone read more than a quarter of it: the bounding box of the donor
    points goes up as one strided copy per variable (HostIO.add_box).  The device copies of the donor fields are
    poisoned first, so the result can only be right if that box holds everything the plans read."""
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.PyTree as X
    import cassiopee_b200.Connector.Mpi as Xmpi
    h = 1. / 20
    a = G.cart((0, 0, 0), (h, h, h), (31, 21, 17), name='A')
    b = G.cart((0.52, 0.11, 0.09), (0.7 * h, 0.7 * h, 0.7 * h), (25, 21, 17), name='B')     # B inside the right part of A
    t = C.newPyTree(['Base', a, b])
    names = ['Density', 'MomentumX']
    rng = np.random.default_rng(31)
    for z in Internal.getZones(t):
        for v in names:
            C._initVars(z, 'centers:' + v, np.asfortranarray(rng.standard_normal(C.fieldShape(z, 'centers'))))
    C._initVars(t, 'centers:cellN', 1.)
    C._initVars(Internal.getZones(t)[1], 'centers:cellN', 2.)
    tc = C.node2Center(t)
    tc = X.setInterpData(t, tc, order=2, nature=0, penalty=1, extrap=1, loc='centers', storage='inverse', sameName=1,
                         itype='chimera', verbose=0)
    want = X.setInterpTransfers(t, tc, variables=names, storage=1)
    sess = Xmpi.TransferSession(t, tc, names, comm=None, ctx=ctx)
    for lst in sess.dD.values():
        for d in lst: d.upload(np.full(d.n, np.nan))
    up, down = sess.transfer_host()
    zA = Internal.getZones(tc)[0]
    full = sum(C.getFieldArray(zA, v).nbytes for v in names)
    assert 0.25 * full < up < 0.86 * full, (up, full)            # a box of A, not its index list, not the whole zone
    for z, zw in zip(Internal.getZones(t), Internal.getZones(want)):
        for v in names:
            assert np.array_equal(C.getFieldArray(z, 'centers:' + v), C.getFieldArray(zw, 'centers:' + v)), (z[0], v)
    sess.close()
