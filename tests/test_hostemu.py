"""CPU tests of the DEVICE logic: cassiopee_b200/csrc/cx_core.cuh (the header
the CUDA kernels are built from) compiled for the host by tests/hostemu and
compared with the oracle: level-synchronous ADT build (with a shuffled
insertion-claim order), DFS candidate order, jump walk / fallback, 24-tetra
coefficients, cellN validity, cross-zone selection, extrapolation.  Bit-exact."""
import numpy as np
import pytest
from tests import cases
from tests.hostemu import emu as E
from oracle import oracle as O


def _compare(dz, pts, nature=1, penalty=1, extrap=1, cellNs=None):
    oz, ez = [], []
    for i, z in enumerate(dz):
        c = cellNs[i] if cellNs else None
        oz.append(O.RefZone(*z, cellN=c)); ez.append(E.EmuZone(*z, cellN=c))
        assert np.array_equal(oz[-1].bbox(), ez[-1].bb)
    ro = O.set_interp_data_raw(pts[0], pts[1], pts[2], oz, 2, nature, penalty, extrap)
    re = E.set_interp_data_raw(pts[0], pts[1], pts[2], ez, nature, penalty, extrap)
    for k in ("noblk", "type", "dind", "isext"):
        assert np.array_equal(ro[k], re[k]), k
    cfo = ro["cf"][:, :8].copy(); cfe = re["cf"].copy()
    t1 = ro["type"] == 1
    cfo[t1, 1:] = 0; cfe[t1, 1:] = 0
    assert np.array_equal(cfo, cfe)
    return ro


@pytest.mark.parametrize("variant", ["a", "b", "c"])
def test_c1_small(variant):
    A, B = cases.c1(variant, n=17, m=11)
    r = _compare([A], B[3:])
    assert (r["noblk"] == 1).all()
    r = _compare([B], A[3:])
    assert r["isext"].sum() > 0


def test_candidate_lists():
    A, B = cases.c1("a", n=13, m=7)
    oz = O.RefZone(*A); ez = E.EmuZone(*A)
    rng = np.random.default_rng(3)
    pts = list(zip(B[3][:60], B[4][:60], B[5][:60])) + [tuple(p) for p in rng.uniform(-0.1, 1.1, (60, 3))]
    for (x, y, z) in pts:
        for alpha in (0.0, 7.5):
            assert np.array_equal(ez.candidates(x, y, z, alpha), oz.candidates(x, y, z, alpha))


def test_cylinder_and_multizone_holes():
    Bg, Oc = cases.c2(nb=20, nc=(33, 9, 17))
    _compare([Bg], Oc[3:])
    _compare([Oc], Bg[3:])
    rng = np.random.default_rng(0)
    cb = np.ones(Bg[3].size); cb[rng.random(cb.size) < 0.05] = 0.; cb[rng.random(cb.size) < 0.05] = 2.
    co = np.ones(Oc[3].size); co[rng.random(co.size) < 0.05] = 0.; co[rng.random(co.size) < 0.05] = 2.
    P = rng.uniform(-3.3, 3.3, (3, 4000)); P[2] = rng.uniform(-0.2, 2.2, 4000)
    for nature in (0, 1):
        for penalty in (0, 1):
            _compare([Bg, Oc], P, nature, penalty, 1, [cb, co])
            _compare([Oc, Bg], P, nature, penalty, 0, [co, cb])


def test_volume_matches():
    Bg, Oc = cases.c2(nb=12, nc=(17, 7, 9))
    oz = O.RefZone(*Oc); ez = E.EmuZone(*Oc)
    for ind in (0, 5, 17 * 3 + 4, 17 * 7 * 2 + 17 * 2 + 3):
        assert oz.volume(ind) == ez.volume(ind)


def _compare_types(dz, pts, types, nature=1, penalty=1, extrap=1, cellNs=None):
    """Same as _compare with a per-zone interpDataType (0: InterpCart, 1: InterpAdt)."""
    oz, ez = [], []
    for i, z in enumerate(dz):
        c = cellNs[i] if cellNs else None
        oz.append(O.RefZone(*z, cellN=c, interpDataType=types[i])); ez.append(E.EmuZone(*z, cellN=c, interpDataType=types[i]))
    ro = O.set_interp_data_raw(pts[0], pts[1], pts[2], oz, 2, nature, penalty, extrap)
    re = E.set_interp_data_raw(pts[0], pts[1], pts[2], ez, nature, penalty, extrap)
    for k in ("noblk", "type", "dind", "isext"):
        assert np.array_equal(ro[k], re[k]), k
    cfo = ro["cf"][:, :8].copy(); cfe = re["cf"].copy()
    t1 = ro["type"] == 1
    cfo[t1, 1:] = 0; cfe[t1, 1:] = 0
    assert np.array_equal(cfo, cfe)
    return ro


def test_cartesian_donors_interpcart_order2_and_extrapolation():
    """interpDataType=0: closed-form Cartesian donor search (InterpCart) — trilinear weights, cellN tests,
    cross-zone choice against an ADT donor, extrapolation over the candidate box."""
    for variant in ("a", "b", "c"):
        A, B = cases.c1(variant, n=17, m=11)
        r = _compare_types([A], B[3:], [0])
        assert (r["noblk"] == 1).all()
        r = _compare_types([B], A[3:], [0])
        assert (r["noblk"] == 0).sum() > 0
    Bg, Oc = cases.c2(nb=20, nc=(33, 9, 17))
    rng = np.random.default_rng(1)
    cb = np.ones(Bg[3].size); cb[rng.random(cb.size) < 0.05] = 0.; cb[rng.random(cb.size) < 0.05] = 2.
    co = np.ones(Oc[3].size); co[rng.random(co.size) < 0.05] = 0.; co[rng.random(co.size) < 0.05] = 2.
    P = rng.uniform(-3.3, 3.3, (3, 4000)); P[2] = rng.uniform(-0.2, 2.2, 4000)
    next_ = 0
    for nature in (0, 1):
        for penalty in (0, 1):
            r = _compare_types([Bg], P, [0], nature, penalty, 1, [cb])
            next_ += int(r["isext"].sum())
            _compare_types([Bg, Oc], P, [0, 1], nature, penalty, 1, [cb, co])
            _compare_types([Oc, Bg], P, [1, 0], nature, penalty, 0, [co, cb])
    assert next_ > 0


@pytest.mark.parametrize("order", [3, 4])
def test_cartesian_donors_interpcart_order3_order4_weights(order):
    A, B = cases.c1("c", n=17, m=11)
    rng = np.random.default_rng(2)
    P = np.concatenate([np.array(B[3:]), rng.uniform(-0.05, 1.05, (3, 3000))], axis=1)
    oz = O.RefZone(*A, interpDataType=0); ez = E.EmuZone(*A, interpDataType=0)
    ro = O.set_interp_data_raw(P[0], P[1], P[2], [oz], order, 1, 1, 0)
    found, dind, cf = E.cart_o3(ez, P[0], P[1], P[2], order)
    hit = ro["noblk"] == 1
    assert np.array_equal(hit, found == 1) and hit.sum() > 1000 and (~hit).sum() > 100
    tt = hit & (ro["type"] == (44 if order == 4 else 3))
    assert tt.sum() > 1000
    assert np.array_equal(ro["dind"][tt], dind[tt])
    assert np.array_equal(ro["cf"][tt][:, :3 * order], cf[tt])


def _zones2d():
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Converter as C
    a = G.cart((0, 0, 0.25), (0.1, 0.1, 1.), (11, 11, 1))
    b = G.cylinder((0.5, 0.5, 0.25), 0.1, 0.6, 360., 0., 1., (33, 9, 1))
    A = (11, 11, 1) + tuple(cases.flat(a)); B = (33, 9, 1) + tuple(cases.flat(b))
    return A, B


def test_two_dimensional_donors_type22():
    """nk = 1 donors in a z = constant plane: extruded hexa search, coefficients folded onto 4 nodes (type 22),
    quad area for the cross-zone choice, 2-D cellN tests, extrapolation, type-1 collapse."""
    A, B = _zones2d()
    rng = np.random.default_rng(0)
    P = rng.uniform(-0.3, 1.3, (3, 3000)); P[2] = 0.25
    P[2][:50] = 0.3                                   # off the plane: never interpolated (Interp2.cpp:204-207)
    P[0][50:150] = np.round(P[0][50:150], 1); P[1][50:150] = np.round(P[1][50:150], 1)   # on donor nodes: type 1
    ca = np.ones(A[3].size); ca[rng.random(ca.size) < 0.06] = 0.; ca[rng.random(ca.size) < 0.06] = 2.
    cb = np.ones(B[3].size); cb[rng.random(cb.size) < 0.06] = 0.; cb[rng.random(cb.size) < 0.06] = 2.
    r = _compare([A], P)
    assert (r["type"] == 22).sum() > 500 and (r["type"] == 1).sum() > 10 and r["isext"].sum() > 100
    for nature in (0, 1):
        for penalty in (0, 1):
            _compare([A, B], P, nature, penalty, 1, [ca, cb])
            _compare([B, A], P, nature, penalty, 0, [cb, ca])
