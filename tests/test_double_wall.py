"""Double wall: connector.setInterpDataDW = the ordinary search with one receptor point per donor
zone (the point projected onto that zone's walls by the caller).  CPU: the oracle (the reference's own
getInterpolationCellDW / getExtrapolationCellDW driven by a restatement of setInterpDataDW's loop) against the ordinary
oracle and the slot rule for type 1; GPU: cx_set_interp_data_dw and Connector.setInterpData__ with a list of receptor
arrays against that oracle."""
import numpy as np
import pytest
from tests import cases


def _zones_and_points(seed=0, n=600):
    A, B = cases.c1("c", n=15, m=11)
    rng = np.random.default_rng(seed)
    lo = np.array([min(A[3].min(), B[3].min()), min(A[4].min(), B[4].min()), min(A[5].min(), B[5].min())])
    hi = np.array([max(A[3].max(), B[3].max()), max(A[4].max(), B[4].max()), max(A[5].max(), B[5].max())])
    p = lo[:, None] - 0.05 + rng.random((3, n)) * (hi - lo + 0.1)[:, None]
    p[:, :40] = np.stack([A[3][100:140], A[4][100:140], A[5][100:140]])        # on donor nodes: type 1
    p[:, 40:60] = np.stack([B[3][7:27], B[4][7:27], B[5][7:27]])
    return A, B, p


def test_oracle_double_wall_reduces_to_the_ordinary_search():
    from oracle import oracle as O
    from cassiopee_b200.Connector.Connector import _dwDonorSlots
    A, B, p = _zones_and_points()
    za = O.RefZone(*A[:3], *A[3:], None); zb = O.RefZone(*B[:3], *B[3:], None)
    for nature in (0, 1):
        for penalty in (0, 1):
            ref = O.set_interp_data(p[0], p[1], p[2], [za, zb], 2, nature, penalty, 1)
            dw = O.set_interp_data_dw([tuple(p), tuple(p)], [za, zb], 2, nature, penalty)
            for k in (0, 2, 3, 4):
                for a, b in zip(ref[k], dw[k]):
                    assert np.array_equal(a, b), k
            assert np.array_equal(ref[5], dw[5])
            assert sum(int((t == 1).sum()) for t in ref[2]) >= 40
            for d, t, q in zip(ref[1], ref[2], dw[1]):
                assert np.array_equal(_dwDonorSlots(d, t), q)
    # a zone that sees the receptors far away takes none of them
    far = (p[0] + 100., p[1], p[2])
    dw = O.set_interp_data_dw([tuple(p), far], [za, zb], 2, 1, 1)
    only = O.set_interp_data(p[0], p[1], p[2], [za], 2, 1, 1, 1)
    assert dw[0][1].size == 0 and np.array_equal(dw[0][0], only[0][0]) and np.array_equal(dw[3][0], only[3][0])


def test_type1_slot_rule():
    """setInterpData.cpp:1655-1659 replayed literally against the vectorised form."""
    from cassiopee_b200.Connector.Connector import _dwDonorSlots
    rng = np.random.default_rng(5)
    for m in (0, 1, 2, 7, 200):
        typ = rng.choice(np.array([1, 1, 2, 22], np.int32), m).astype(np.int32)
        dnr = rng.integers(0, 10 ** 6, m).astype(np.int32)
        lit = np.full(2 * m + 2, -1, np.int32)
        for e in range(m):
            if typ[e] == 1: lit[e + 1] = dnr[e]
            else: lit[e] = dnr[e]
        assert np.array_equal(_dwDonorSlots(dnr, typ), lit[:m])


@pytest.mark.gpu
def test_double_wall_matches_reference(ctx):
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    A, B, p = _zones_and_points(seed=3, n=4000)
    rng = np.random.default_rng(9)
    cA = np.ones(A[3].size); cA[rng.random(cA.size) < 0.05] = 0.; cA[rng.random(cA.size) < 0.05] = 2.
    gz = [_lib.Zone(ctx, *A[:3], *A[3:], cA), _lib.Zone(ctx, *B[:3], *B[3:], None)]
    oz = [O.RefZone(*A[:3], *A[3:], cA), O.RefZone(*B[:3], *B[3:], None)]
    h = A[3][1] - A[3][0]
    pa = tuple(p)                                                       # zone A sees the points where they are,
    pb = (p[0] + 0.37 * h * rng.standard_normal(p.shape[1]), p[1] - 0.21 * h, p[2] + 0.5 * h * rng.random(p.shape[1]))  # B displaced
    for nature in (0, 1):
        for penalty in (0, 1):
            got = _lib.set_interp_data_dw(ctx, [pa, pb], gz, 2, nature, penalty).fetch()
            ref = O.set_interp_data_dw([pa, pb], oz, 2, nature, penalty)
            from cassiopee_b200.Connector.Connector import _dwDonorSlots
            for z in range(2):
                assert np.array_equal(got[0][z], ref[0][z]), "receptors zone %d" % z
                assert np.array_equal(_dwDonorSlots(got[1][z], got[2][z]), ref[1][z]), "donors zone %d" % z
                assert np.array_equal(got[2][z], ref[2][z]) and np.array_equal(got[3][z], ref[3][z])
                assert np.array_equal(got[4][z], ref[4][z])
            assert np.array_equal(got[5], ref[5])
            assert got[0][0].size > 100 and got[0][1].size > 100 and sum(e.size for e in got[4]) > 10
    # the displaced zone really changes the outcome
    same = _lib.set_interp_data_dw(ctx, [pa, pa], gz, 2, 1, 1).fetch()
    assert not np.array_equal(same[0][1], got[0][1])
    # Lagrange orders through the same per-zone coordinates (order 5 on fewer points: one 125 x 125 solve per stencil)
    for order, m in ((3, 4000), (5, 600)):
        qa = tuple(a[:m] for a in pa); qb = tuple(a[:m] for a in pb)
        got = _lib.set_interp_data_dw(ctx, [qa, qb], gz, order, 1, 1).fetch()
        ref = O.set_interp_data_dw([qa, qb], oz, order, 1, 1)
        for z in range(2):
            assert np.array_equal(got[0][z], ref[0][z]), (order, "receptors", z)
            assert np.array_equal(_dwDonorSlots(got[1][z], got[2][z]), ref[1][z]), (order, "donors", z)
            assert np.array_equal(got[2][z], ref[2][z]) and np.array_equal(got[3][z], ref[3][z]), (order, z)
            assert np.array_equal(got[4][z], ref[4][z])
        assert np.array_equal(got[5], ref[5])
        assert sum(int((t == order).sum()) for t in got[2]) > 50
    with pytest.raises(NotImplementedError):
        _lib.set_interp_data_dw(ctx, [pa, pb], gz, 4, 1, 1)


@pytest.mark.gpu
def test_double_wall_through_the_array_api(ctx):
    """Connector.setInterpData__ with a LIST of receptor arrays (Connector.py:427-429), hooks given or built."""
    import cassiopee_b200.Connector.Connector as XC
    from oracle import oracle as O
    A, B, p = _zones_and_points(seed=4, n=1500)
    zonesD = [['x,y,z,cellN', np.stack([A[3], A[4], A[5], np.ones(A[3].size)]), A[0], A[1], A[2]],
              ['x,y,z', np.stack([B[3], B[4], B[5]]), B[0], B[1], B[2]]]
    pb = p + np.array([[0.013], [-0.02], [0.007]])
    pts = [['x,y,z', np.ascontiguousarray(p), p.shape[1], 1, 1], ['x,y,z', np.ascontiguousarray(pb), p.shape[1], 1, 1]]
    oz = [O.RefZone(*A[:3], *A[3:], np.ones(A[3].size)), O.RefZone(*B[:3], *B[3:], None)]
    ref = O.set_interp_data_dw([tuple(p), tuple(pb)], oz, 2, 0, 1)
    for hook in (None, XC.createHooks__(zonesD, ctx)):
        got = XC.setInterpData__(pts, zonesD, order=2, penalty=1, nature=0, hook=hook, ctx=ctx)
        for k in range(5):
            for a, b in zip(got[k], ref[k]):
                assert np.array_equal(a, b), k
        assert np.array_equal(got[5], ref[5])
    empty = [['x,y,z', np.zeros((3, 0)), 0, 1, 1]] * 2
    assert XC.setInterpData__(empty, zonesD, ctx=ctx) is None
