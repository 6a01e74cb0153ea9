"""CPU tests of the Python host layer (tree model, storage layout, argument
semantics, error behaviour) with the oracle injected as compute backend."""
import numpy as np
import pytest
import cassiopee_b200.Internal as Internal
import cassiopee_b200.Converter as C
import cassiopee_b200.Generator as G
import cassiopee_b200.Distributor2 as D2
import cassiopee_b200.Connector.OversetData as X
from tests.backends import OracleBackend


def _case():
    h = 1. / 12
    a = G.cart((0, 0, 0), (h, h, h), (13, 13, 13), name='A')
    b = G.cart((0.31, 0.28, 0.33), (0.8 * h, 0.8 * h, 0.8 * h), (9, 9, 9), name='B')
    t = C.newPyTree(['Base', a, b])
    C._initVars(t, 'centers:Density', lambda x, y, z: 1. + x + 2 * y + 3 * z)
    C._initVars(t, 'centers:cellN', 1.)
    C._initVars(Internal.getZones(t)[1], 'centers:cellN', 2.)
    return t


def test_generators_follow_reference_formulas():
    z = G.cart((1., 2., 3.), (0.5, 0.25, 0.125), (4, 3, 2))
    x, y, zz = C.getCoords(z)
    assert x.shape == (4, 3, 2) and x.flags.f_contiguous
    assert x[3, 0, 0] == 1. + 3 * 0.5 and y[0, 2, 0] == 2. + 2 * 0.25 and zz[0, 0, 1] == 3. + 0.125
    c = G.cylinder((0, 0, 0), 0.5, 2.5, 360., 0., 2., (9, 5, 3))
    x, y, zz = C.getCoords(c)
    assert abs(np.hypot(x[:, 0, 0], y[:, 0, 0]) - 0.5).max() < 1e-15
    assert abs(np.hypot(x[:, -1, 0], y[:, -1, 0]) - 2.5).max() < 1e-14
    assert zz[0, 0, -1] == 2.0
    assert abs(x[0, :, :] - x[-1, :, :]).max() < 1e-14      # seam duplicated


def test_tree_model_and_node2center():
    t = _case()
    zs = Internal.getZones(t)
    assert [z[0] for z in zs] == ['A', 'B']
    assert Internal.getZoneDim(zs[0]) == ['Structured', 13, 13, 13, 3]
    tc = C.node2Center(t)
    zc = Internal.getZones(tc)[0]
    assert Internal.getZoneDim(zc)[1:4] == [12, 12, 12]
    assert C.getFieldArray(zc, 'Density').shape == (12, 12, 12)      # centre field became a node field of tc
    # tc's fields are copies, like the reference's node2Center (Converter/PyTree.py:6400-6414): a transfer that writes
    # t must not change what the next subregion reads from tc
    assert not np.shares_memory(C.getFieldArray(zs[0], 'centers:Density'), C.getFieldArray(zc, 'Density'))
    assert np.array_equal(C.getFieldArray(zs[0], 'centers:Density'), C.getFieldArray(zc, 'Density'))
    n = Internal.createNode('s', 'DataArray_t', 'Donor')
    assert Internal.getValue(n) == 'Donor' and n[1].dtype.char == 'c'


def test_set_interp_data_inverse_storage_layout():
    t = _case()
    tc = C.node2Center(t)
    tD = Internal.copyRef(tc); tR = Internal.copyRef(t)
    X._setInterpDataChimera(tR, tD, order=2, nature=1, penalty=1, extrap=1, loc='centers', storage='inverse',
                            sameName=1, verbose=0, backend=OracleBackend())
    zA, zB = Internal.getZones(tD)
    s = Internal.getNodeFromName1(zA, 'ID_B')
    assert s is not None and s[3] == 'ZoneSubRegion_t' and Internal.getValue(s) == 'B'
    names = [c[0] for c in s[2]]
    assert names[:6] == ['ZoneRole', 'GridLocation', 'PointList', 'PointListDonor', 'InterpolantsDonor',
                         'InterpolantsType']
    assert Internal.getValue(s[2][0]) == 'Donor' and Internal.getValue(s[2][1]) == 'CellCenter'
    pl = Internal.getNodeFromName1(s, 'PointList')[1]; pld = Internal.getNodeFromName1(s, 'PointListDonor')[1]
    typ = Internal.getNodeFromName1(s, 'InterpolantsType')[1]; cf = Internal.getNodeFromName1(s, 'InterpolantsDonor')[1]
    assert pl.dtype == np.int32 and pld.dtype == np.int32 and typ.dtype == np.int32 and cf.dtype == np.float64
    assert pl.size == pld.size == typ.size == 8 ** 3        # every centre of B is a receptor
    assert np.array_equal(pld, np.arange(8 ** 3))           # ascending receptor index
    assert cf.size == sum({1: 1, 2: 8}[int(v)] for v in typ)
    assert Internal.getNodeFromName1(zB, 'ID_A') is None     # A has no cellN==2 points
    assert C.getFieldArray(zA, 'cellN') is not None          # tc carries the centres' cellN as a node field
    # sameName=1 drops the donor named like the receptor
    tD2 = Internal.copyRef(tc); tR2 = Internal.copyRef(t)
    C._initVars(tR2, 'centers:cellN', 2.)
    X._setInterpDataChimera(tR2, tD2, order=2, nature=1, loc='centers', storage='inverse', sameName=1, verbose=0,
                            backend=OracleBackend())
    for zd in Internal.getZones(tD2):
        assert Internal.getNodeFromName1(zd, 'ID_' + zd[0]) is None


def test_direct_storage_and_orphans():
    t = _case()
    tc = C.node2Center(t)
    tR = Internal.copyRef(t); tD = Internal.copyRef(tc)
    C._initVars(tR, 'centers:cellN', 2.)       # A's centres mostly lie outside B -> orphans without extrapolation
    C._initVars(tD, 'cellN', 1.)               # donors all valid
    X._setInterpDataChimera(tR, tD, order=2, nature=1, extrap=0, loc='centers', storage='direct', verbose=0,
                            backend=OracleBackend())
    zA = Internal.getZones(tR)[0]
    s = Internal.getNodeFromName1(zA, 'ID_B')
    assert Internal.getValue(Internal.getNodeFromName1(s, 'ZoneRole')) == 'Receiver'
    orph = Internal.getNodeFromName1(s, 'OrphanPointList')[1]
    assert orph.size > 0 and np.array_equal(orph, np.unique(orph))
    rcv = Internal.getNodeFromName1(s, 'PointList')[1]
    assert np.intersect1d(rcv, orph).size == 0 and rcv.size + orph.size == 12 ** 3
    # verbose=3 also marks the orphans with -1 in a copy of cellN (OversetData.py:1287-1294)
    tR3 = Internal.copyRef(t); tD3 = Internal.copyRef(tc)
    C._initVars(tR3, 'centers:cellN', 2.); C._initVars(tD3, 'cellN', 1.)
    X._setInterpDataChimera(tR3, tD3, order=2, nature=1, extrap=0, loc='centers', storage='direct', verbose=3,
                            backend=OracleBackend())
    zA3 = Internal.getZones(tR3)[0]
    co = C.getFieldArray(zA3, 'centers:cellN#Orphan').reshape(-1, order='F')
    assert np.array_equal(np.nonzero(co == -1.)[0], orph) and np.all(co[rcv] == 2.)
    assert np.all(C.getFieldArray(zA3, 'centers:cellN') == 2.)        # cellN itself is untouched
    # ... and the extrapolated points with -2 (OversetData.py:1313-1319)
    tR4 = Internal.copyRef(t); tD4 = Internal.copyRef(tc)
    C._initVars(tR4, 'centers:cellN', 2.); C._initVars(tD4, 'cellN', 1.)
    X._setInterpDataChimera(tR4, tD4, order=2, nature=1, extrap=1, loc='centers', storage='direct', verbose=3,
                            backend=OracleBackend())
    zA4 = Internal.getZones(tR4)[0]
    s4 = Internal.getNodeFromName1(zA4, 'ID_B')
    ext = Internal.getNodeFromName1(s4, 'ExtrapPointList')[1]
    co4 = C.getFieldArray(zA4, 'centers:cellN#Orphan').reshape(-1, order='F')
    assert ext.size > 0 and np.array_equal(np.nonzero(co4 == -2.)[0], np.sort(ext))
    o4 = Internal.getNodeFromName1(s4, 'OrphanPointList')
    if o4 is not None: assert np.array_equal(np.nonzero(co4 == -1.)[0], o4[1])


def test_argument_errors_match_reference():
    t = _case(); tc = C.node2Center(t)
    with pytest.raises(ValueError):
        X._setInterpDataChimera(t, tc, loc='faces', backend=OracleBackend())
    with pytest.raises(ValueError):
        X._setInterpDataChimera(t, tc, loc='centers', method='spline', backend=OracleBackend())
    with pytest.raises(NotImplementedError):
        X._setInterpDataChimera(t, tc, loc='centers', method='leastsquares', backend=OracleBackend())
    with pytest.raises(TypeError):
        X._setInterpDataChimera(t, tc, loc='centers', hook=[1, 2, 3], backend=OracleBackend())
    # no cellN on the receptor tree -> nothing to do (OversetData.py:1184)
    t2 = C.newPyTree(['Base', G.cart((0, 0, 0), (1, 1, 1), (3, 3, 3))])
    assert X._setInterpDataChimera(t2, t2, loc='nodes', backend=OracleBackend()) is None


def test_variable_selection_rules():
    t = _case(); tc = C.node2Center(t)
    zr = Internal.getZones(t)[1]; zd = Internal.getZones(tc)[0]
    C._initVars(zd, 'Extra', 3.)
    names, strict, loc = X._selectVariables(zr, zd, [], 'cellN', 1)
    assert 'Density' in names and 'cellN' not in names and 'Extra' not in names and strict and loc == 'centers'
    names, strict, _ = X._selectVariables(zr, zd, ['Density', 'Missing'], 'cellN', 1)
    assert names == ['Density'] and not strict
    names, strict, _ = X._selectVariables(zr, zd, ['Density', 'cellN'], 'cellN', 1)
    assert names == ['Density'] and strict


def test_distributor2_proc_nodes():
    zones = [G.cart((i, 0, 0), (0.1, 0.1, 0.1), (5, 5, 5), name='z%d' % i) for i in range(6)]
    t = C.newPyTree(['Base'] + zones)
    out = D2._distribute(t, 3, algorithm='fast')
    d = D2.getProcDict(t)
    assert sorted(d.keys()) == ['z%d' % i for i in range(6)] and set(d.values()) == {0, 1, 2}
    assert max(out['nptsPerProc']) == min(out['nptsPerProc'])
    z0 = Internal.getZones(t)[0]
    node = Internal.getNodeFromName1(Internal.getNodeFromName1(z0, '.Solver#Param'), 'proc')
    assert node[1].shape == (1, 1) and node[1].dtype == np.int32
    D2._distribute(t, 2, algorithm='rcb')
    d = D2.getProcDict(t)
    assert [d['z%d' % i] for i in range(6)] == [0, 0, 0, 1, 1, 1]


def test_apply_bc_overlaps_and_intersecting_domains():
    """applyBCOverlaps marks the `depth` cell layers next to BCOverlap windows (the receptors of a
    chimera assembly); getIntersectingDomains gives the bbox-intersecting donor lists."""
    import cassiopee_b200.Connector.PyTree as XP
    a = G.cart((0, 0, 0), (0.1, 0.1, 0.1), (11, 11, 11), name='A')
    b = G.cart((0.75, 0.2, 0.2), (0.1, 0.1, 0.1), (8, 7, 6), name='B')
    c = G.cart((5, 5, 5), (0.1, 0.1, 0.1), (4, 4, 4), name='C')
    t = C.newPyTree(['Base', a, b, c])
    zA, zB, zC = Internal.getZones(t)
    C._addBC2Zone(zA, 'ov1', 'BCOverlap', [11, 11, 1, 11, 1, 11])     # imax face of A
    C._addBC2Zone(zB, 'ov1', 'BCOverlap', [1, 1, 1, 7, 1, 6])         # imin face of B
    C._addBC2Zone(zB, 'wall', 'BCWall', [1, 8, 1, 1, 1, 6])
    be = OracleBackend()
    t2 = XP.applyBCOverlaps(t, depth=2, loc='centers', backend=be)
    cA = C.getFieldArray(Internal.getZones(t2)[0], 'centers:cellN')
    cB = C.getFieldArray(Internal.getZones(t2)[1], 'centers:cellN')
    assert cA.shape == (10, 10, 10) and (cA[8:, :, :] == 2).all() and (cA[:8, :, :] == 1).all()
    assert (cB[:2, :, :] == 2).all() and (cB[2:, :, :] == 1).all()
    assert C.getFieldArray(zA, 'centers:cellN') is None               # the input tree is untouched
    t3 = XP.applyBCOverlaps(t, depth=1, loc='nodes', backend=be)
    nA = C.getFieldArray(Internal.getZones(t3)[0], 'cellN')
    assert (nA[10, :, :] == 2).all() and (nA[:10, :, :] == 1).all()
    d = XP.getIntersectingDomains(t)
    assert d == {'A': ['B'], 'B': ['A'], 'C': []}
    # blanked cells stay blanked
    C._initVars(zA, 'centers:cellN', 1.)
    C.getFieldArray(zA, 'centers:cellN')[9, 0, 0] = 0.
    XP._applyBCOverlaps(zA, depth=2, loc='centers', backend=be)
    assert C.getFieldArray(zA, 'centers:cellN')[9, 0, 0] == 0. and C.getFieldArray(zA, 'centers:cellN')[9, 1, 0] == 2.


def test_hooks_and_interpdatatype_through_the_tree_layer():
    """hook=[C.createHook(zd,'adt') ...] gives the same tree as hook=None; interpDataType=0 (Cartesian donors,
    InterpCart) goes through the same layer; bad hook lists / types raise like the reference."""
    t = _case(); tc = C.node2Center(t)
    be = OracleBackend()
    ref = Internal.copyRef(tc)
    X._setInterpDataChimera(Internal.copyRef(t), ref, order=2, nature=1, loc='centers', storage='inverse', verbose=0,
                            backend=be)
    hooks = [C.createHook(z, 'adt', backend=be) for z in Internal.getZones(tc)]
    got = Internal.copyRef(tc)
    X._setInterpDataChimera(Internal.copyRef(t), got, order=2, nature=1, loc='centers', storage='inverse', verbose=0,
                            backend=be, hook=hooks)
    for za, zb in zip(Internal.getZones(ref), Internal.getZones(got)):
        sa = Internal.getNodesFromType1(za, 'ZoneSubRegion_t'); sb = Internal.getNodesFromType1(zb, 'ZoneSubRegion_t')
        assert [s[0] for s in sa] == [s[0] for s in sb]
        for a, b in zip(sa, sb):
            for name in ('PointList', 'PointListDonor', 'InterpolantsType', 'InterpolantsDonor'):
                assert np.array_equal(Internal.getNodeFromName1(a, name)[1], Internal.getNodeFromName1(b, name)[1])
    for h in hooks:
        C.freeHook(h)
    with pytest.raises(TypeError):
        X._setInterpDataChimera(Internal.copyRef(t), Internal.copyRef(tc), loc='centers', storage='inverse', verbose=0,
                                backend=be, hook=hooks[:1])
    with pytest.raises(TypeError):
        X._setInterpDataChimera(Internal.copyRef(t), Internal.copyRef(tc), loc='centers', storage='inverse', verbose=0,
                                backend=be, interpDataType=2)
    with pytest.raises(NotImplementedError):
        C.createHook(Internal.getZones(tc)[0], 'nodes', backend=be)
    # Cartesian donors: trilinear weights, type 2 everywhere the ADT path also finds a cell
    cart = Internal.copyRef(tc)
    X._setInterpDataChimera(Internal.copyRef(t), cart, order=2, nature=1, loc='centers', storage='inverse', verbose=0,
                            backend=be, interpDataType=0)
    sa = Internal.getNodeFromName1(Internal.getZones(ref)[0], 'ID_B'); sc = Internal.getNodeFromName1(Internal.getZones(cart)[0], 'ID_B')
    assert np.array_equal(Internal.getNodeFromName1(sa, 'PointListDonor')[1], Internal.getNodeFromName1(sc, 'PointListDonor')[1])
    cf = Internal.getNodeFromName1(sc, 'InterpolantsDonor')[1]; ty = Internal.getNodeFromName1(sc, 'InterpolantsType')[1]
    assert (ty == 2).all() and np.allclose(cf.reshape(-1, 8).sum(axis=1), 1., atol=1e-13)


def _hole_fringe_loops(c, depth, dir):
    """Restatement with plain loops of searchMaskInterpolatedCellsStructOpt, 3-D
    (Connector/Connector/getInterpolatedPoints.cpp:444-494 cross, :572-650 star)."""
    im, jm, km = c.shape
    out = c.copy()
    for k in range(km):
        for j in range(jm):
            for i in range(im):
                if abs(c[i, j, k] - 1.) > 1e-15: continue
                idx = []
                if dir == 0:
                    for d in range(-depth, 0): idx.append((i, j, max(0, k + d)))
                    for d in range(-depth, 0): idx.append((i, max(0, j + d), k))
                    for d in range(-depth, 0): idx.append((max(0, i + d), j, k))
                    for d in range(1, depth + 1): idx.append((min(i + d, im - 1), j, k))
                    for d in range(1, depth + 1): idx.append((i, min(j + d, jm - 1), k))
                    for d in range(1, depth + 1): idx.append((i, j, min(k + d, km - 1)))
                else:
                    def comp(v, off, n):
                        w = v + off
                        return v if (w < 0 or w > n - 1) else w
                    for kd in list(range(-depth, 0)) + list(range(1, depth + 1)):
                        kk = comp(k, kd, km)
                        for jd in (-1, 0, 1):
                            for id_ in (-1, 0, 1):
                                idx.append((comp(i, id_ * abs(kd), im), comp(j, jd * abs(kd), jm), kk))
                    for d in range(1, depth + 1):
                        for jd in (-1, 1):
                            jj = comp(j, jd * d, jm)
                            idx += [(comp(i, -d, im), jj, k), (i, jj, k), (comp(i, d, im), jj, k)]
                        idx += [(comp(i, -d, im), j, k), (comp(i, d, im), j, k)]
                if any(abs(c[q]) <= 1e-15 for q in idx):
                    out[i, j, k] = 2.
    return out


def test_set_hole_interpolated_points():
    import cassiopee_b200.Connector.PyTree as XP
    rng = np.random.default_rng(4)
    for shape in ((9, 8, 7), (12, 5, 1)):
        for dir_ in (0, 1):
            for depth in (1, 2, 3):
                c = np.ones(shape, order='F')
                c[rng.random(shape) < 0.03] = 0.
                c[rng.random(shape) < 0.03] = 2.
                z = G.cart((0, 0, 0), (1, 1, 1), (shape[0] + 1, shape[1] + 1, shape[2] + 1), name='Z')
                C._initVars(z, 'centers:cellN', c)
                z2 = XP.setHoleInterpolatedPoints(z, depth=depth, dir=dir_, loc='centers', backend=OracleBackend())
                got = C.getFieldArray(Internal.getZones(z2)[0], 'centers:cellN')
                assert np.array_equal(got, _hole_fringe_loops(c, depth, dir_)), (shape, dir_, depth)
                assert np.array_equal(C.getFieldArray(z, 'centers:cellN'), c)      # input untouched
    # depth < 0: the fringe grows into the blanked region
    c = np.ones((7, 7, 7), order='F'); c[2:5, 2:5, 2:5] = 0.
    z = G.cart((0, 0, 0), (1, 1, 1), (8, 8, 8), name='Z'); C._initVars(z, 'centers:cellN', c)
    XP._setHoleInterpolatedPoints(z, depth=-1, dir=0, loc='centers', backend=OracleBackend())
    got = C.getFieldArray(z, 'centers:cellN')
    assert got[3, 3, 3] == 0. and got[2, 3, 3] == 2. and got[1, 3, 3] == 1. and got[2, 2, 2] == 2.
    with pytest.raises(NotImplementedError):
        XP._setHoleInterpolatedPoints(z, depth=2, dir=2, loc='centers', backend=OracleBackend())


def test_donor_tree_round_trip_on_disk(tmp_path):
    """tc with its ID_* subregions written and read back ('bin_pickle', a format of the reference's Converter):
    the interpData survives bit for bit and drives the same transfers."""
    t = _case(); tc = C.node2Center(t)
    be = OracleBackend()
    X._setInterpDataChimera(t, tc, order=2, nature=1, loc='centers', storage='inverse', verbose=0, backend=be)
    f = str(tmp_path / "tc.pickle")
    C.convertPyTree2File(tc, f)
    tc2 = C.convertFile2PyTree(f)
    for za, zb in zip(Internal.getZones(tc), Internal.getZones(tc2)):
        sa = Internal.getNodesFromType1(za, 'ZoneSubRegion_t'); sb = Internal.getNodesFromType1(zb, 'ZoneSubRegion_t')
        assert [s[0] for s in sa] == [s[0] for s in sb] and len(sa) > 0 or len(sb) == 0
        for a, b in zip(sa, sb):
            assert Internal.getValue(a) == Internal.getValue(b)
            for ca, cb in zip(a[2], b[2]):
                assert ca[0] == cb[0] and ca[3] == cb[3]
                if isinstance(ca[1], np.ndarray):
                    assert ca[1].dtype == cb[1].dtype and np.array_equal(ca[1], cb[1])
    with pytest.raises(NotImplementedError):
        C.convertPyTree2File(tc, str(tmp_path / "tc.cgns"))


def test_rotation_and_compact_helpers():
    """Host logic of the periodicity-by-rotation correction and of compact=1 (no compute)."""
    t = _case(); tc = C.node2Center(t)
    zr = Internal.getZones(t)[1]
    s = ['ID_B', None, [], 'ZoneSubRegion_t']
    assert X._rotationOf(s) == (0, 0.)
    Internal.createChild(s, 'RotationAngle', 'DataArray_t', np.array([0., -0.25, 0.5]))
    assert X._rotationOf(s) == (2, -0.25)          # first non-zero component (setInterpTransfers.cpp:281-285)
    for v in ('VelocityX', 'VelocityY', 'VelocityZ'):
        C._initVars(zr, 'centers:' + v, 1.)
    names = C.getVarNames(zr, 'centers')
    trip = X._rotationTriplets(zr, 'centers', len(names))
    assert trip == [tuple(names.index('Velocity' + a) for a in 'XYZ')]
    with pytest.raises(ValueError):
        X._rotationTriplets(zr, 'centers', 2)      # the reference would index vectOfRcvFields out of bounds
    # compact=1: fields must sit one behind the other in one block
    block = np.arange(5 * 8, dtype=np.float64)
    views = X._compactFields(block[:8], 5, 8, 'donor')
    assert len(views) == 5 and all(v.size == 8 for v in views)
    assert views[3][0] == 24. and views[3].__array_interface__['data'][0] == block.__array_interface__['data'][0] + 24 * 8
    blockF = np.zeros((2, 2, 2, 6), order='F')
    viewsF = X._compactFields(blockF[:, :, :, 0], 6, 8, 'receiver')
    viewsF[5][:] = 7.
    assert blockF[0, 0, 0, 5] == 7. and blockF[1, 1, 1, 5] == 7. and blockF[..., :5].max() == 0.
    with pytest.raises(ValueError):
        X._compactFields(np.zeros(8), 5, 8, 'donor')     # a stand-alone array is not a compacted block
    with pytest.raises(ValueError):
        X._compactFields(block[16:24], 5, 8, 'donor')    # not enough fields behind the first one
    with pytest.raises(TypeError):
        X._compactFields(None, 5, 8, 'donor')


@pytest.mark.parametrize("loc,storage", [('nodes', 'direct'), ('centers', 'inverse')])
@pytest.mark.parametrize("order", [2, 3, 5])
def test_reference_script_storage_layouts(loc, storage, order):
    """Host half of Connector/test/setInterpTransfersPT_t1.py / _t2.py (identical 11^3 Cartesian grids): where the
    ID_* subregion lands, its role, and that its lists are the oracle's for the receptor points the tree layer
    extracts (nodes, or node2Center of the receptor zone)."""
    from oracle import oracle as O
    ni = nj = nk = 11
    m = G.cart((0, 0, 0), (1, 1, 1), (ni, nj, nk), name='donor')
    a = G.cart((0, 0, 0), (1, 1, 1), (ni, nj, nk), name='extraction')
    tD = C.newPyTree(['Dnr', m]); tR = C.newPyTree(['Rcv', a])
    C._initVars(tR, 'cellN' if loc == 'nodes' else 'centers:cellN', 2.)
    X._setInterpDataChimera(tR, tD, order=order, loc=loc, storage=storage, verbose=0, backend=OracleBackend())
    zr = Internal.getZones(tR)[0]; zd = Internal.getZones(tD)[0]
    holder, other, sname, role = (zr, zd, 'ID_donor', 'Receiver') if storage == 'direct' else (zd, zr, 'ID_extraction', 'Donor')
    s = Internal.getNodeFromName1(holder, sname)
    assert s is not None and not Internal.getNodesFromType1(other, 'ZoneSubRegion_t')
    assert Internal.getValue(Internal.getNodeFromName1(s, 'ZoneRole')) == role
    assert (Internal.getNodeFromName1(s, 'GridLocation') is not None) == (loc == 'centers')
    xr, yr, zz = C.getCoords(zr)
    if loc == 'centers':
        xr, yr, zz = [C.node2CenterArray(v) for v in (xr, yr, zz)]
    pts = [np.ascontiguousarray(v.reshape(-1, order='F')) for v in (xr, yr, zz)]
    xd, yd, zdd = [v.reshape(-1, order='F') for v in C.getCoords(zd)]
    ref = O.set_interp_data(pts[0], pts[1], pts[2], [O.RefZone(ni, nj, nk, xd, yd, zdd, np.ones(xd.size))],
                            order=order, nature=0, penalty=1, extrap=1)
    pl = Internal.getNodeFromName1(s, 'PointList')[1]; pld = Internal.getNodeFromName1(s, 'PointListDonor')[1]
    rcv, dnr = (pl, pld) if storage == 'direct' else (pld, pl)
    assert np.array_equal(rcv, ref[0][0]) and np.array_equal(dnr, ref[1][0])
    assert np.array_equal(Internal.getNodeFromName1(s, 'InterpolantsType')[1], ref[2][0])
    assert np.array_equal(Internal.getNodeFromName1(s, 'InterpolantsDonor')[1], ref[3][0])
    assert rcv.size == pts[0].size and ref[5].size == 0
    # coincident nodes collapse to type 1 for every order; cell centres keep the order's own type
    assert set(np.unique(ref[2][0])) == ({1} if loc == 'nodes' else {order})
    assert C.isNamePresent(zd, 'cellN') == -1          # the cellN added for the call is removed again


@pytest.mark.parametrize("loc", ['nodes', 'centers'])
@pytest.mark.parametrize("order", [2, 3, 5])
def test_reference_script_setInterpDataPT_t1_curve_receptors(loc, order):
    """Connector/test/setInterpDataPT_t1.py: a circle (structured curve, 20 points) as receptor zone, nodes and
    centres (19 segment mid-points, node2Center.cpp:80-95), both storages, penalty and nature sweeps."""
    from tests import cases
    for storage in ('direct', 'inverse'):
        for pen in (0, 1):
            for nat in (0, 1):
                cases.check_curve_receptors(dict(backend=OracleBackend()), order, loc, storage, pen, nat)


def test_distributed_entry_same_base_semantics():
    """Connector.Mpi._setInterpData keeps the reference's sameBase rule (Connector/Mpi.py:858-862, :930-937):
    with the default sameBase=0 a zone is interpolated only from zones of another base."""
    import cassiopee_b200.Connector.Mpi as Xmpi
    def trees(two_bases):
        t = _case()
        if two_bases:
            za, zb = Internal.getZones(t)
            t = C.newPyTree(['BaseA', za, 'BaseB', zb])
        return t, C.node2Center(t)
    for two_bases, sameBase, expect in ((False, 0, False), (False, 1, True), (True, 0, True), (True, 1, True)):
        t, tc = trees(two_bases)
        Xmpi._setInterpData(t, tc, order=2, nature=1, loc='centers', storage='inverse', sameBase=sameBase, itype='chimera', verbose=0,
                            comm=None, backend=OracleBackend())
        zdA = [z for z in Internal.getZones(tc) if z[0] == 'A'][0]
        assert (Internal.getNodeFromName1(zdA, 'ID_B') is not None) == expect, (two_bases, sameBase)
    # cartesian=True only names a transport of the reference's Cmpi._addXZones: accepted, same result
    t, tc = trees(True)
    Xmpi._setInterpData(t, tc, order=2, nature=1, loc='centers', storage='inverse', sameBase=1, cartesian=True, itype='chimera',
                        verbose=0, comm=None, backend=OracleBackend())
    assert Internal.getNodeFromName1([z for z in Internal.getZones(tc) if z[0] == 'A'][0], 'ID_B') is not None
    # the reference's defaults (storage='direct', itype='abutting'): nothing to do without match connectivity, and the
    # combinations without a device path say so
    t, tc = trees(True)
    assert Xmpi._setInterpData(t, tc, comm=None, backend=OracleBackend()) is None
    assert Internal.getNodeFromName1([z for z in Internal.getZones(tc) if z[0] == 'A'][0], 'ID_B') is None
    with pytest.raises(NotImplementedError):
        Xmpi._setInterpData(t, tc, loc='centers', itype='chimera', comm=None, backend=OracleBackend())      # storage='direct'
    with pytest.raises(NotImplementedError):
        Xmpi._setInterpData(t, tc, loc='centers', storage='inverse', itype='chimeraOld', comm=None, backend=OracleBackend())


def test_create_interp_region_node_order_and_concatenation():
    """_createInterpRegion__ (OversetData.py:1752-1884): child names and order, the ExtC / volume / rotation
    children, and the concatenation rule when the subregion already exists."""
    z = Internal.newZone('D', 3, 3, 3)
    i32 = lambda *v: np.array(v, dtype=np.int32)
    rot = ['RotationAngle', np.array([0., 0., 30.]), [['DimensionalUnits', None, [], 'DimensionalUnits_t']], 'DataArray_t']
    ctr = ['RotationCenter', np.zeros(3), [], 'DataArray_t']
    X._createInterpRegion__(z, 'R', i32(5, 6), i32(0, 1), np.arange(16.), i32(2, 2), np.array([.1, .2]), i32(1),
                            i32(7, 7, 3), pointlistdonorExtC=i32(9, 9), tag='Donor', loc='centers',
                            RotationAngle=rot, RotationCenter=ctr)
    s = Internal.getNodeFromName1(z, 'ID_R')
    assert Internal.getValue(s) == 'R'
    assert [c[0] for c in s[2]] == ['ZoneRole', 'GridLocation', 'PointList', 'PointListDonor', 'PointListExtC',
                                    'InterpolantsDonor', 'InterpolantsVol', 'InterpolantsType', 'OrphanPointList',
                                    'ExtrapPointList', 'RotationAngle', 'RotationCenter']
    assert Internal.getNodeFromName1(s, 'RotationAngle')[2] == []            # DimensionalUnits child removed
    assert np.array_equal(Internal.getNodeFromName1(s, 'OrphanPointList')[1], [3, 7])
    X._createInterpRegion__(z, 'R', i32(8), i32(2), np.arange(8.) + 100, i32(2), np.array([.3]), i32(2), i32(4),
                            pointlistdonorExtC=i32(10), tag='Donor', loc='centers')
    assert len(Internal.getNodesFromName1(z, 'ID_R')) == 1
    assert np.array_equal(Internal.getNodeFromName1(s, 'PointList')[1], [5, 6, 8])
    assert np.array_equal(Internal.getNodeFromName1(s, 'PointListDonor')[1], [0, 1, 2])
    assert np.array_equal(Internal.getNodeFromName1(s, 'PointListExtC')[1], [9, 9, 10])
    assert Internal.getNodeFromName1(s, 'InterpolantsDonor')[1].size == 24
    assert np.array_equal(Internal.getNodeFromName1(s, 'InterpolantsVol')[1], [.1, .2, .3])
    assert np.array_equal(Internal.getNodeFromName1(s, 'InterpolantsType')[1], [2, 2, 2])
    assert np.array_equal(Internal.getNodeFromName1(s, 'OrphanPointList')[1], [4])       # replaced, not appended
    assert np.array_equal(Internal.getNodeFromName1(s, 'ExtrapPointList')[1], [1, 2])
    # receiver side names the extra list PointListDonorExtC; prefix overrides the ID_/IBCD_ rule
    z2 = Internal.newZone('R', 3, 3, 3)
    X._createInterpRegion__(z2, 'D', i32(1), i32(2), np.ones(8), i32(2), np.array([]), i32(), i32(),
                            pointlistdonorExtC=i32(3), tag='Receiver', loc='nodes', prefix='XX_')
    s2 = Internal.getNodeFromName1(z2, 'XX_D')
    assert [c[0] for c in s2[2]] == ['ZoneRole', 'PointList', 'PointListDonor', 'PointListDonorExtC',
                                     'InterpolantsDonor', 'InterpolantsType']
    with pytest.raises(NotImplementedError):
        X._createInterpRegion__(z2, 'D', i32(1), i32(2), np.ones(8), i32(2), np.array([]), i32(), i32(), loc='faces')


def test_set_interp_data2_wrappers(monkeypatch):
    """setInterpData2 / _setInterpData2 (OversetData.py:1083-1101) and their distributed forms (Connector/Mpi.py:1025-1105):
    fixed arguments (nature=1, penalty=1, extrap=1, inverse storage, sameName=0), cellN = 2 put on a receptor tree that
    has none and removed afterwards, previous subregions of the donor tree dropped by the distributed form, the inputs
    of the copying forms untouched.  Same trees as Connector/test/setInterpData2PT_m1.py (cart 11^3 donors,
    cart 21^3 receptors of half the spacing)."""
    import cassiopee_b200.Connector.PyTree as XP
    import cassiopee_b200.Connector.Mpi as Xmpi
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    monkeypatch.setattr(_lib, 'DeviceBackend', lambda ctx: OracleBackend())
    monkeypatch.setattr(_lib, 'default_context', lambda *a: None)
    aD = G.cart((0, 0, 0), (1, 1, 1), (11, 11, 11), name='dnr')
    aR = G.cart((0, 0, 0), (0.5, 0.5, 0.5), (21, 21, 21), name='rcv')
    tD = C.newPyTree(['BD', aD]); tR = C.newPyTree(['BR', aR])
    out = XP.setInterpData2(tR, tD, order=2, loc='centers')
    assert C.isNamePresent(tR, 'centers:cellN') == -1 and not Internal.getNodesFromType2(tD, 'ZoneSubRegion_t')
    s = Internal.getNodeFromName1(Internal.getZones(out)[0], 'ID_rcv')
    assert s is not None and Internal.getValue(Internal.getNodeFromName1(s, 'ZoneRole')) == 'Donor'
    # oracle on the same points: all 20^3 centres of the receptor zone, donors = the nodes of the donor tree
    xr, yr, zr = [np.ascontiguousarray(C.node2CenterArray(v).reshape(-1, order='F')) for v in C.getCoords(aR)]
    xd, yd, zd = [v.reshape(-1, order='F') for v in C.getCoords(aD)]
    ref = O.set_interp_data(xr, yr, zr, [O.RefZone(11, 11, 11, xd, yd, zd, np.ones(xd.size))], order=2, nature=1,
                            penalty=1, extrap=1)
    for name, k in (('PointList', 1), ('PointListDonor', 0), ('InterpolantsType', 2), ('InterpolantsDonor', 3)):
        assert np.array_equal(Internal.getNodeFromName1(s, name)[1], ref[k][0]), name
    # in-place distributed form, single rank: stale subregions of the donor tree are dropped first
    tD2 = Internal.copyRef(out); tR2 = C.newPyTree(['BR', aR])
    Internal.getZones(tD2)[0][2].append(['ID_old', np.zeros(1), [], 'ZoneSubRegion_t'])
    Xmpi._setInterpData2(tR2, tD2, loc='centers', cartesian=False, comm=None, backend=OracleBackend())
    subs = Internal.getNodesFromType1(Internal.getZones(tD2)[0], 'ZoneSubRegion_t')
    assert [n[0] for n in subs] == ['ID_rcv']
    assert np.array_equal(Internal.getNodeFromName1(subs[0], 'InterpolantsDonor')[1], ref[3][0])
    assert C.isNamePresent(tR2, 'centers:cellN') == -1
    # a receptor tree that carries its own cellN keeps it, and only its cellN == 2 points are interpolated
    tR3 = C.newPyTree(['BR', G.cart((0, 0, 0), (0.5, 0.5, 0.5), (21, 21, 21), name='rcv')])
    cn = np.ones((20, 20, 20), order='F'); cn[:2] = 2.
    C._initVars(tR3, 'centers:cellN', cn)
    out3 = Xmpi.setInterpData2(tR3, tD, loc='centers', comm=None, backend=OracleBackend())
    s3 = Internal.getNodeFromName1(Internal.getZones(out3)[0], 'ID_rcv')
    assert Internal.getNodeFromName1(s3, 'PointListDonor')[1].size == 2 * 20 * 20
    assert C.isNamePresent(tR3, 'centers:cellN') == 1


def test_receptor_zones_in_windows(monkeypatch):
    """_setInterpDataChimera queues the searches of a window of receptor zones before fetching any: with a window of
    one zone, or of all zones, the tree is the same (three mutually overlapping zones, every zone receptor and donor;
    hooks built in batches, some given by the caller)."""
    h = 1. / 10
    def build():
        zs = [G.cart((0, 0, 0), (h, h, h), (11, 11, 11), name='A'),
              G.cart((0.45, 0.12, 0.08), (0.9 * h, 0.9 * h, 0.9 * h), (10, 10, 10), name='B'),
              G.cart((0.21, 0.47, 0.33), (0.8 * h, 0.8 * h, 0.8 * h), (9, 9, 9), name='C')]
        t = C.newPyTree(['Base'] + zs)
        C._initVars(t, 'centers:cellN', 1.)
        for z in Internal.getZones(t):
            c = C.getFieldArray(z, 'centers:cellN'); c[:2] = 2.; c[:, -2:] = 2.
        return t, C.node2Center(t)
    be = OracleBackend()
    trees = []
    for window, hooks in ((1, False), (10 ** 9, False), (5, True)):
        monkeypatch.setattr(X, '_WINDOW_PTS', window)
        t, tc = build()
        hook = None
        if hooks:      # the caller gives one hook, the others are built in the window's batch
            zd = Internal.getZones(tc)[1]
            _, ni, nj, nk, _ = Internal.getZoneDim(zd)
            xd, yd, zzd = [a.reshape(-1, order='F') for a in C.getCoords(zd)]
            hook = [None, be.make_zone(ni, nj, nk, xd, yd, zzd, C.getFieldArray(zd, 'cellN').reshape(-1, order='F')), None]
        X._setInterpDataChimera(t, tc, order=2, nature=1, loc='centers', storage='inverse', verbose=0, backend=be,
                                hook=hook, refreshHooks=False)
        trees.append(tc)
    names = ('PointList', 'PointListDonor', 'InterpolantsDonor', 'InterpolantsType', 'ExtrapPointList', 'OrphanPointList')
    ref = trees[0]
    nsub = 0
    for other in trees[1:]:
        for za, zb in zip(Internal.getZones(ref), Internal.getZones(other)):
            sa = [s for s in Internal.getNodesFromType1(za, 'ZoneSubRegion_t') if s[0].startswith('ID_')]
            sb = [s for s in Internal.getNodesFromType1(zb, 'ZoneSubRegion_t') if s[0].startswith('ID_')]
            assert [s[0] for s in sa] == [s[0] for s in sb]
            for a, b in zip(sa, sb):
                for n in names:
                    na, nb = Internal.getNodeFromName1(a, n), Internal.getNodeFromName1(b, n)
                    assert (na is None) == (nb is None)
                    if na is not None: assert np.array_equal(na[1], nb[1]), (za[0], a[0], n)
                nsub += 1
    assert nsub >= 8


def test_cartesian_base_zones_are_interpcart_donors():
    """Xmpi._setInterpData (single process): the donor zones of a base named 'CARTESIAN' are located in closed form
    (interpDataType 0), the others through their ADT, like the reference's hooks[z] = None rule (Connector/Mpi.py:950-983);
    cartesian=True only names a transport of the reference and is accepted."""
    import cassiopee_b200.Connector.Mpi as Xmpi
    h = 1. / 10
    def build():
        a = G.cart((0, 0, 0), (h, h, h), (11, 11, 11), name='A')
        b = G.cart((0.45, 0.12, 0.08), (0.9 * h, 0.9 * h, 0.9 * h), (10, 10, 10), name='B')
        c = G.cart((0.21, 0.47, 0.33), (0.8 * h, 0.8 * h, 0.8 * h), (9, 9, 9), name='C')
        t = C.newPyTree(['CARTESIAN', a, 'Body', b, c])
        C._initVars(t, 'centers:cellN', 1.)
        for z in Internal.getZones(t):
            cn = C.getFieldArray(z, 'centers:cellN'); cn[:2] = 2.; cn[:, -2:] = 2.
        return t, C.node2Center(t)

    class Recording(OracleBackend):
        def __init__(self): self.types = {}
        def make_zone(self, ni, nj, nk, x, y, z, cellN=None, interpDataType=1):
            self.types[(ni, nj, nk)] = interpDataType
            return OracleBackend.make_zone(self, ni, nj, nk, x, y, z, cellN, interpDataType=interpDataType)

    be = Recording()
    t, tc = build()
    Xmpi._setInterpData(t, tc, order=2, nature=1, penalty=1, extrap=1, loc='centers', storage='inverse', sameBase=1,
                        cartesian=True, itype='chimera', verbose=0, comm=None, backend=be)
    assert be.types == {(10, 10, 10): 0, (9, 9, 9): 1, (8, 8, 8): 1}
    t2, tc2 = build()
    X._setInterpDataChimera(t2, tc2, order=2, nature=1, penalty=1, extrap=1, loc='centers', storage='inverse', verbose=0,
                            interpDataType=[0, 1, 1], backend=OracleBackend())
    n = 0
    for za, zb in zip(Internal.getZones(tc), Internal.getZones(tc2)):
        sa = [s for s in Internal.getNodesFromType1(za, 'ZoneSubRegion_t') if s[0].startswith('ID_')]
        sb = [s for s in Internal.getNodesFromType1(zb, 'ZoneSubRegion_t') if s[0].startswith('ID_')]
        assert [s[0] for s in sa] == [s[0] for s in sb]
        for a, b in zip(sa, sb):
            for nm in ('PointList', 'PointListDonor', 'InterpolantsDonor', 'InterpolantsType'):
                assert np.array_equal(Internal.getNodeFromName1(a, nm)[1], Internal.getNodeFromName1(b, nm)[1])
            n += 1
    assert n >= 4


def test_apply_bc_overlaps_overset_families_and_checkcelln():
    """BC_t windows whose family carries .Solver#Overlap count like BCOverlap windows (Connector/PyTree.py:1634-1650);
    checkCellN=False leaves a zone without cellN alone."""
    import cassiopee_b200.Connector.PyTree as XP
    a = G.cart((0, 0, 0), (0.1, 0.1, 0.1), (11, 9, 7), name='A')
    b = G.cart((2, 0, 0), (0.1, 0.1, 0.1), (6, 6, 6), name='B')
    t = C.newPyTree(['Base', a, b])
    base = Internal.getBases(t)[0]
    fam = Internal.createChild(base, 'OVS', 'Family_t')
    Internal.createChild(fam, '.Solver#Overlap', 'UserDefinedData_t')
    Internal.createChild(base, 'WALLS', 'Family_t')
    zA = Internal.getZones(t)[0]
    C._addBC2Zone(zA, 'ovfam', 'FamilySpecified', [1, 11, 1, 1, 1, 7])       # jmin face, overset by family
    bc = Internal.getNodesFromType2(zA, 'BC_t')[0]
    Internal.createChild(bc, 'FamilyName', 'FamilyName_t', value='OVS')
    C._addBC2Zone(zA, 'wall', 'FamilySpecified', [11, 11, 1, 9, 1, 7])       # imax face, an ordinary family
    Internal.createChild(Internal.getNodesFromType2(zA, 'BC_t')[1], 'FamilyName', 'FamilyName_t', value='WALLS')
    be = OracleBackend()
    XP._applyBCOverlaps(t, depth=2, loc='centers', backend=be)
    cA = C.getFieldArray(zA, 'centers:cellN')
    assert (cA[:, :2, :] == 2).all() and (cA[:, 2:, :] == 1).all()
    cB = C.getFieldArray(Internal.getZones(t)[1], 'centers:cellN')
    assert cB is not None and (cB == 1).all()                                # created, no window
    t2 = C.newPyTree(['Base', G.cart((0, 0, 0), (1, 1, 1), (4, 4, 4), name='Z')])
    XP._applyBCOverlaps(t2, depth=1, loc='centers', checkCellN=False, backend=be)
    assert C.getFieldArray(Internal.getZones(t2)[0], 'centers:cellN') is None
