"""GPU tests at BASELINE.json's full sizes (run with -m gpu on the B200).

Parity gate against the oracle (= the reference's own C++, oracle/_ref/libcassref.so) on BASELINE's configs:
  * C1a/b/c at 91^3 / 63^3, both directions, every receptor, order 2 (+ orders 3 and 5 for B <- A);
  * C2a (256^3 background donor) and C2b (O-grid 513 x 129 x 257 donor): the oracle builds the FULL-SIZE donor ADT
    (a few seconds, printed) and searches a strided sample of >= 1 M receptors plus every receptor lying within
    1e-6 of a donor grid plane or on the donor bounding box; the device runs ALL receptors and its packed lists,
    restricted to the sample, must be bit-identical (orders 2 and 3; order 5 on >= 30 k receptors);
  * InterpCart (interpDataType=0): C2a, all 17.0 M receptors;
  * the 8-zone share of C4 through Connector.Mpi._setInterpData, device backend vs oracle backend: every child of
    every ID_* subregion identical;
plus size-independent properties (ascending lists, partition of unity, linear field exact to 1e-12, idempotent and
linear transfers, dense / in-place / host / device variants bit-identical)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2a(ctx):
    from cassiopee_b200 import workloads as W
    Bg, Oc = W.c2(1.0)
    return Bg, Oc


_NCF = {1: 1, 2: 8, 3: 9, 44: 12, 5: 15, 22: 4, 4: 8}


def _restrict(packed, sel, n):
    """The packed 7-list of a call on n receptors, restricted to the receptors `sel` (ascending) and renumbered
    by their position in sel: what the same call on the points sel alone returns (receptors are independent)."""
    look = np.full(n, -1, np.int64)
    look[sel] = np.arange(sel.size)
    nz = len(packed[0])
    R, D, T, Cf, E = [], [], [], [], []
    lut = np.zeros(64, np.int64)
    for k, v in _NCF.items():
        lut[k] = v
    for z in range(nz):
        r = packed[0][z]
        keep = look[r] >= 0
        ncf = lut[packed[2][z]]
        R.append(look[r[keep]].astype(np.int32)); D.append(packed[1][z][keep]); T.append(packed[2][z][keep])
        Cf.append(packed[3][z][np.repeat(keep, ncf)])
        e = look[packed[4][z]]
        E.append(e[e >= 0].astype(np.int32))
    o = look[packed[5]]
    return [R, D, T, Cf, E, o[o >= 0].astype(np.int32), []]


def _assert_same(got, ref, what):
    for z in range(len(ref[0])):
        for q, nm in enumerate(("receptors", "donors", "types", "coefficients", "extrapolated")):
            assert got[q][z].shape == ref[q][z].shape, (what, nm, z, got[q][z].shape, ref[q][z].shape)
            if not np.array_equal(got[q][z], ref[q][z]):
                bad = np.nonzero(got[q][z] != ref[q][z])[0]
                raise AssertionError("%s: %s of zone %d differ at %d entries, first %s" % (what, nm, z, bad.size, bad[:5]))
    assert np.array_equal(got[5], ref[5]), (what, "orphans")


@pytest.mark.parametrize("variant", ["a", "b", "c"])
def test_c1_fullsize_bit_exact_vs_reference(ctx, variant):
    """BASELINE configs[0] at its stated size: A = cart 91^3, B = cart 63^3, centre meshes, both directions, every
    receptor (A <- B exercises extrapolation and orphans on 386 k points)."""
    import time
    from cassiopee_b200 import _lib, workloads as W
    from oracle import oracle as O
    A, B = W.c1(variant)
    Ac, Bc = A.centers(), B.centers()
    assert Ac.npts == 729000 and Bc.npts == 238328
    t0 = time.perf_counter()
    oa = O.RefZone(Ac.ni, Ac.nj, Ac.nk, Ac.x, Ac.y, Ac.z); ob = O.RefZone(Bc.ni, Bc.nj, Bc.nk, Bc.x, Bc.y, Bc.z)
    t_adt = time.perf_counter() - t0
    ga = _lib.Zone(ctx, Ac.ni, Ac.nj, Ac.nk, Ac.x, Ac.y, Ac.z); gb = _lib.Zone(ctx, Bc.ni, Bc.nj, Bc.nk, Bc.x, Bc.y, Bc.z)
    t0 = time.perf_counter()
    for order in (2, 3, 5):
        st = 1 if order < 5 else 6          # order 5 costs the oracle 0.1 ms per receptor: every 6th receptor (39.7 k)
        px, py, pz = [np.ascontiguousarray(a[::st]) for a in (Bc.x, Bc.y, Bc.z)]
        got = _lib.set_interp_data(ctx, px, py, pz, [ga], order, 1, 1, 1).fetch()
        ref = O.set_interp_data(px, py, pz, [oa], order, 1, 1, 1)
        _assert_same(got, ref, "C1%s B<-A order %d" % (variant, order))
        assert got[0][0].size == px.size
    got = _lib.set_interp_data(ctx, Ac.x, Ac.y, Ac.z, [gb], 2, 1, 1, 1).fetch()
    ref = O.set_interp_data(Ac.x, Ac.y, Ac.z, [ob], 2, 1, 1, 1)
    _assert_same(got, ref, "C1%s A<-B order 2" % variant)
    assert got[4][0].size > 100000 and got[5].size > 300000          # the extrapolation / orphan path did run
    # transfers of the 5 conservative variables, B <- A and A <- B, against the reference loop
    fA = W.conservative_fields(Ac.x, Ac.y, Ac.z, 5, seed=0)
    got = _lib.set_interp_data(ctx, Bc.x, Bc.y, Bc.z, [ga], 2, 1, 1, 1).fetch()
    rcv, dnr, typ, cf = got[0][0], got[1][0], got[2][0], got[3][0]
    plan = _lib.Plan(ctx, Ac.ni, Ac.nj, Ac.nk, Bc.npts, dnr, rcv, typ, cf)
    outG = [np.zeros(Bc.npts) for _ in range(5)]; outR = [np.zeros(Bc.npts) for _ in range(5)]
    plan.transfer(fA, outG)
    O.transfer(fA, outR, Ac.ni, Ac.nj, rcv, dnr, typ, cf)
    for a, b in zip(outG, outR):
        assert np.array_equal(a, b)
    print("C1%s full size: oracle ADTs %.2f s, oracle + device searches %.1f s" % (variant, t_adt, time.perf_counter() - t0))


def _near_plane(v, v0, h, tol=1.e-6):
    f = (v - v0) / h
    return np.abs(f - np.rint(f)) * h < tol


def _c2_sample(n, stride, extra):
    m = np.zeros(n, bool)
    m[::stride] = True
    m |= extra
    return np.nonzero(m)[0]


def test_c2a_fullsize_adt_bit_exact_vs_reference(ctx, c2a):
    """BASELINE configs[1]/[2], direction (2a): all 17.0 M O-grid nodes <- the 256^3 background through the ADT path.
    Oracle = the reference's serial buildStructAdt on the full 16.6 M-cell donor + its search on a strided sample
    (every 17th receptor, 1.0 M) united with every receptor within 1e-6 of a donor grid plane and the two node planes
    lying on the donor bounding box (z = 0 and z = 2)."""
    import time
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    Bg, Oc = c2a
    n = Oc.npts
    t0 = time.perf_counter()
    oz = O.RefZone(Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z)
    t_adt = time.perf_counter() - t0
    gz = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z)
    assert np.array_equal(gz.info()["bbox"], oz.bbox())
    hx, hy, hz = Bg.x[1] - Bg.x[0], Bg.y[Bg.ni] - Bg.y[0], Bg.z[Bg.ni * Bg.nj] - Bg.z[0]
    near = _near_plane(Oc.x, Bg.x[0], hx) | _near_plane(Oc.y, Bg.y[0], hy) | _near_plane(Oc.z, Bg.z[0], hz)
    sel = _c2_sample(n, 17, near)
    assert sel.size >= 1000000 and near.sum() >= 2 * Oc.ni * Oc.nj
    xs, ys, zs = Oc.x[sel], Oc.y[sel], Oc.z[sel]
    times = {}
    for order, sub in ((2, None), (3, None), (5, 24)):
        s = sel if sub is None else sel[::sub]
        full = _lib.set_interp_data(ctx, Oc.x, Oc.y, Oc.z, [gz], order, 1, 1, 1).fetch() if order != 5 else None
        t0 = time.perf_counter()
        ref = O.set_interp_data(Oc.x[s], Oc.y[s], Oc.z[s], [oz], order, 1, 1, 1)
        times[order] = (s.size, time.perf_counter() - t0)
        if full is None:        # order 5 at 17 M receptors is tens of seconds even on the device: the sample only
            got = _lib.set_interp_data(ctx, Oc.x[s], Oc.y[s], Oc.z[s], [gz], order, 1, 1, 1).fetch()
        else:
            assert full[0][0].size == n and full[5].size == 0
            got = _restrict(full, s, n)
        _assert_same(got, ref, "C2a order %d" % order)
        assert s.size >= 30000
    print("C2a full size: oracle ADT of %d cells %.2f s; oracle searches %s" % ((Bg.ni - 1) ** 3, t_adt, times))


def test_c2b_fullsize_adt_bit_exact_vs_reference(ctx, c2a):
    """Direction (2b): background nodes inside the annulus 0.5 < r < 2.5 <- the curvilinear O-grid (513 x 129 x 257,
    seam duplicated, ADT depth 49).  Sample: every 9th receptor + the background planes z = 0 / z = 2 (on the donor
    bounding box) + the nodes within 1e-6 of a radial or axial donor grid surface."""
    import time
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    Bg, Oc = c2a
    r = np.sqrt(Bg.x * Bg.x + Bg.y * Bg.y)
    inside = np.nonzero((r > 0.5) & (r < 2.5))[0]
    xr, yr, zr = Bg.x[inside], Bg.y[inside], Bg.z[inside]
    n = inside.size
    assert n > 8000000
    t0 = time.perf_counter()
    oz = O.RefZone(Oc.ni, Oc.nj, Oc.nk, Oc.x, Oc.y, Oc.z)
    t_adt = time.perf_counter() - t0
    gz = _lib.Zone(ctx, Oc.ni, Oc.nj, Oc.nk, Oc.x, Oc.y, Oc.z)
    assert np.array_equal(gz.info()["bbox"], oz.bbox())
    rr = r[inside]
    near = _near_plane(rr, 0.5, 2.0 / (Oc.nj - 1)) | _near_plane(zr, 0., 2.0 / (Oc.nk - 1)) | (np.abs(yr) < 1.e-6)
    sel = _c2_sample(n, 9, near)
    assert sel.size >= 900000
    times = {}
    for order, sub in ((2, None), (3, 4), (5, 30)):
        s = sel if sub is None else sel[::sub]
        t0 = time.perf_counter()
        ref = O.set_interp_data(xr[s], yr[s], zr[s], [oz], order, 1, 1, 1)
        times[order] = (s.size, time.perf_counter() - t0)
        if order == 2:
            full = _lib.set_interp_data(ctx, xr, yr, zr, [gz], order, 1, 1, 1).fetch()
            got = _restrict(full, s, n)
        else:
            got = _lib.set_interp_data(ctx, xr[s], yr[s], zr[s], [gz], order, 1, 1, 1).fetch()
        _assert_same(got, ref, "C2b order %d" % order)
        assert s.size >= 30000
    print("C2b full size: oracle ADT of %d cells %.2f s; oracle searches %s" %
          ((Oc.ni - 1) * (Oc.nj - 1) * (Oc.nk - 1), t_adt, times))


def _subregions(tc):
    import cassiopee_b200.Internal as Internal
    out = {}
    for zd in Internal.getZones(tc):
        for s in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
            out[(zd[0], s[0])] = [(c[0], c[3], None if c[1] is None else np.array(c[1])) for c in s[2]]
    return out


def _c4_share(nb=(2, 2, 2), n=100):
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Generator as G
    from cassiopee_b200 import workloads as W
    blocks = W.c4_layout(nb, n=n, overlap=8, seed=1)
    zones = []
    for b in blocks:
        z = G.cart(b['origin'], (b['h'],) * 3, (b['n'],) * 3, name=b['name'])
        C._initVars(z, 'centers:cellN', W.c4_receptor_mask(b, blocks, depth=2))
        zones.append(z)
    t = C.newPyTree(['Base'] + zones)
    return t, C.node2Center(t), [b['name'] for b in blocks]


def test_c4_share_tree_level_bit_exact_vs_reference(ctx):
    """The per-GPU share of BASELINE configs[3] (8 zones of 100^3, 2-layer fringes, sub-cell shifts) through the tree
    layer Connector.Mpi._setInterpData: device backend vs oracle backend (the reference's C++ behind the same host
    code), every child of every ID_* subregion compared (names, types, order, values)."""
    import time
    import cassiopee_b200.Connector.Mpi as Xmpi
    from tests.backends import OracleBackend
    kw = dict(order=2, nature=1, penalty=1, extrap=1, loc='centers', storage='inverse', sameBase=1, comm=None, verbose=0,
              itype='chimera')
    t, tc, names = _c4_share()
    t0 = time.perf_counter()
    Xmpi._setInterpData(t, tc, zoneOrder=names, ctx=ctx, **kw)
    t_dev = time.perf_counter() - t0
    t2, tc2, _ = _c4_share()
    t0 = time.perf_counter()
    Xmpi._setInterpData(t2, tc2, zoneOrder=names, backend=OracleBackend(), **kw)
    t_ref = time.perf_counter() - t0
    got, ref = _subregions(tc), _subregions(tc2)
    assert set(got) == set(ref) and len(ref) >= 24
    nrcv = 0
    for k in ref:
        assert [(c[0], c[1]) for c in got[k]] == [(c[0], c[1]) for c in ref[k]], k
        for cg, cr in zip(got[k], ref[k]):
            if cr[2] is None:
                assert cg[2] is None
            else:
                assert cg[2].dtype == cr[2].dtype and np.array_equal(cg[2], cr[2]), (k, cr[0])
            if cr[0] == 'PointListDonor':
                nrcv += cr[2].size
    assert nrcv > 400000
    print("C4 share (8 x 100^3): %d subregions, %d receptors; device %.3f s, oracle backend %.1f s" %
          (len(ref), nrcv, t_dev, t_ref))


def test_c2a_fullsize_interpcart_bit_exact_vs_reference(ctx, c2a):
    from cassiopee_b200 import _lib
    from oracle import oracle as O
    Bg, Oc = c2a
    gz = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z, interpDataType=0)
    oz = O.RefZone(Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z, interpDataType=0)
    for order in (2, 3):
        got = _lib.set_interp_data(ctx, Oc.x, Oc.y, Oc.z, [gz], order, 1, 1, 1).fetch()
        ref = O.set_interp_data(Oc.x, Oc.y, Oc.z, [oz], order, 1, 1, 1)
        assert got[0][0].size == Oc.npts
        for q in range(5):
            assert np.array_equal(got[q][0], ref[q][0]), (order, q)
        assert np.array_equal(got[5], ref[5])


def test_c2a_fullsize_adt_path_properties(ctx, c2a):
    from cassiopee_b200 import _lib
    Bg, Oc = c2a
    n = Oc.npts
    zone = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z)
    out = _lib.set_interp_data(ctx, Oc.x, Oc.y, Oc.z, [zone], 2, 1, 1, 1).fetch()
    rcv, dnr, typ, cf = out[0][0], out[1][0], out[2][0], out[3][0]
    assert rcv.size == n and out[5].size == 0 and out[4][0].size == 0        # all found, none extrapolated
    assert np.array_equal(rcv, np.arange(n, dtype=np.int32))                  # ascending receptor order
    assert set(np.unique(typ)) <= {1, 2}
    t2 = typ == 2
    ncf = np.where(t2, 8, 1)
    off = np.concatenate([[0], np.cumsum(ncf)[:-1]])
    assert cf.size == ncf.sum()
    c8 = cf[(off[t2][:, None] + np.arange(8)[None, :])]
    assert np.abs(c8.sum(axis=1) - 1.).max() < 1e-12 and c8.min() > -1e-3    # partition of unity, inside the cell
    # the donor cell is the one the closed-form Cartesian search finds, for points strictly inside a cell
    zc = _lib.Zone(ctx, Bg.ni, Bg.nj, Bg.nk, Bg.x, Bg.y, Bg.z, interpDataType=0)
    oc = _lib.set_interp_data(ctx, Oc.x, Oc.y, Oc.z, [zc], 2, 1, 1, 1).fetch()
    h = (Bg.x[1] - Bg.x[0], Bg.y[Bg.ni] - Bg.y[0], Bg.z[Bg.ni * Bg.nj] - Bg.z[0])
    fr = [((a - a0) / hh) % 1. for a, a0, hh in ((Oc.x, Bg.x[0], h[0]), (Oc.y, Bg.y[0], h[1]), (Oc.z, Bg.z[0], h[2]))]
    interior = t2 & (oc[2][0] == 2)
    for f in fr:
        interior &= (f > 1e-6) & (f < 1. - 1e-6)
    assert interior.sum() > 0.5 * n
    assert np.array_equal(dnr[interior], oc[1][0][interior])
    # transfers: linear field exact, idempotent, linear in the donor field, variants agree
    plan = _lib.Plan(ctx, Bg.ni, Bg.nj, Bg.nk, n, dnr, rcv, typ, cf)
    F = Bg.x + 2 * Bg.y + 3 * Bg.z
    rng = np.random.default_rng(0)
    G = rng.standard_normal(Bg.npts)
    oF = np.zeros(n); oG = np.zeros(n); oS = np.zeros(n)
    plan.transfer([F, G], [oF, oG])
    assert np.abs(oF - (Oc.x + 2 * Oc.y + 3 * Oc.z)).max() < 1e-12
    oF2 = oF.copy(); oG2 = np.zeros(n)
    plan.transfer([F, G], [oF2, oG2])
    assert np.array_equal(oF2, oF) and np.array_equal(oG2, oG)
    plan.transfer([F + G], [oS])
    assert np.abs(oS - (oF + oG)).max() < 1e-12 * max(1., np.abs(oS).max())
    dense = plan.transferD([F, G])
    assert np.array_equal(dense[0], oF) and np.array_equal(dense[1], oG)      # rcv is the identity here
    dD = [_lib.DeviceArray.from_host(ctx, a) for a in (F, G)]
    dR = [_lib.DeviceArray(ctx, n) for _ in range(2)]
    plan.transfer_dev([d.ptr for d in dD], [d.ptr for d in dR]); ctx.sync()
    assert np.array_equal(dR[0].download(), oF) and np.array_equal(dR[1].download(), oG)
    # checksum of checksums: the order of evaluation inside a launch does not matter
    assert float(np.sum(oG.view(np.uint64) % 1000003)) == float(np.sum(dR[1].download().view(np.uint64) % 1000003))


def test_c4_cluster_fullsize_properties(ctx):
    """8 zones of 100^3 (the per-GPU share of C4): distributed code path with one rank, linear field exact on
    every interpolated receptor, both plan-set launch and per-plan host-pointer transfers agree."""
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Internal as Internal
    import cassiopee_b200.Connector.Mpi as Xmpi
    import cassiopee_b200.Connector.PyTree as X
    from cassiopee_b200 import workloads as W
    blocks = W.c4_layout((2, 2, 2), n=100, overlap=8, seed=1)
    zones = []
    for b in blocks:
        z = G.cart(b['origin'], (b['h'],) * 3, (b['n'],) * 3, name=b['name'])
        C._initVars(z, 'centers:cellN', W.c4_receptor_mask(b, blocks, depth=2))
        C._initVars(z, 'centers:F', lambda x, y, zz: x + 2 * y + 3 * zz)
        zones.append(z)
    t = C.newPyTree(['Base'] + zones)
    tc = C.node2Center(t)
    for z in Internal.getZones(tc):         # node2Center shares the centre arrays: give the donor tree its own F
        C._initVars(z, 'F', np.array(C.getFieldArray(z, 'F'), order='F', copy=True))
    Xmpi._setInterpData(t, tc, order=2, nature=1, penalty=1, extrap=1, loc='centers', storage='inverse', sameBase=1, comm=None,
                        itype='chimera',
                        zoneOrder=[b['name'] for b in blocks], ctx=ctx, verbose=0)
    nrcv = 0
    orphans = {}
    extrap = {}
    for zd in Internal.getZones(tc):
        for s in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
            pld = Internal.getNodeFromName1(s, 'PointListDonor')[1]
            assert np.array_equal(pld, np.sort(pld)) and np.unique(pld).size == pld.size
            o = Internal.getNodeFromName1(s, 'OrphanPointList')
            if o is not None:
                orphans[Internal.getValue(s)] = o[1]
            x = Internal.getNodeFromName1(s, 'ExtrapPointList')
            if x is not None and x[1].size:
                extrap.setdefault(Internal.getValue(s), []).append(x[1])
            nrcv += pld.size
    norph = sum(v.size for v in orphans.values())
    mask_total = sum(int((C.getFieldArray(z, 'centers:cellN') == 2.).sum()) for z in Internal.getZones(t))
    assert nrcv + norph == mask_total and nrcv > 400000 and norph < 0.001 * mask_total
    for z in Internal.getZones(t):          # orphans keep their value: take them out of the receptor mask
        if z[0] in orphans:
            C.getFieldArray(z, 'centers:cellN').reshape(-1, order='F')[orphans[z[0]]] = 1.
    # receptors zeroed, then transferred back: exact for a linear field (extrapolated receptors, next to the outer
    # boundary of the cluster, are not linear-exact by construction of getExtrapolationCoeffForCell: left out)
    exact = {}; keep = {}
    next_ = 0
    for z in Internal.getZones(t):
        f = C.getFieldArray(z, 'centers:F'); m = C.getFieldArray(z, 'centers:cellN') == 2.
        exact[z[0]] = f.copy(); f[m] = 0.
        k = np.ones(f.size, bool)
        for lst in extrap.get(z[0], []):
            k[lst] = False; next_ += lst.size
        keep[z[0]] = k.reshape(f.shape, order='F')
    assert next_ < 0.2 * nrcv
    t2 = X.setInterpTransfers(t, tc, variables=['F'], storage=1)
    for z in Internal.getZones(t2):
        assert np.abs(C.getFieldArray(z, 'centers:F') - exact[z[0]])[keep[z[0]]].max() < 1e-12
    got1 = {z[0]: C.getFieldArray(z, 'centers:F').copy() for z in Internal.getZones(t2)}
    for z in Internal.getZones(t):
        f = C.getFieldArray(z, 'centers:F'); f[C.getFieldArray(z, 'centers:cellN') == 2.] = 0.
    sess = Xmpi.TransferSession(t, tc, ['F'], comm=None, ctx=ctx)
    sess.step(2); sess.download()
    for z in Internal.getZones(t):      # fused plan-set launch == per-plan host-pointer calls, bit for bit
        assert np.array_equal(C.getFieldArray(z, 'centers:F'), got1[z[0]])
