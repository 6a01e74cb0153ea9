"""Compact solver transfers: the packer `miseAPlatDonorTree__` (reference layout of
param_int / param_real, compactTransfers.py:16) and the reader `___setInterpTransfers`
(setInterpTransfers.cpp:761-1410).

Checker = oracle/ref_driver.cpp:ref_transfer_compact, the reference's reader loop restated for chimera racs with the
reference's own arithmetic fragments commonInterpTransfers_reorder_5eq.h / _6eq.h #included.  CPU tests: a tree
packed here is read correctly by that reader (layout index for index) and gives what the plain transfer loop gives
on the unpacked interpData (type 1 = injection aside).  GPU tests: the device path on the same flattened tree."""
import numpy as np
import pytest

import cassiopee_b200.Internal as Internal
import cassiopee_b200.Converter as C
import cassiopee_b200.Generator as G
import cassiopee_b200.Connector.OversetData as X
import cassiopee_b200.Connector.compactTransfers as CT
from tests.backends import OracleBackend

VARS5 = ['Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature']
VARS6 = VARS5 + ['TurbulentSANuTilde']


def _tree(kind, neq):
    """Three overlapping Cartesian zones with cell-centred fields.  'coincident': B's centres fall on A's centres
    (type-1 entries); 'shifted': generic type-2 entries.  C is interpolated from both."""
    h = 1. / 12
    a = G.cart((0, 0, 0), (h, h, h), (13, 13, 13), name='A')
    if kind == 'coincident':
        b = G.cart((4 * h, 3 * h, 2 * h), (h, h, h), (8, 8, 8), name='B')
    else:
        b = G.cart((0.31, 0.28, 0.33), (0.8 * h, 0.8 * h, 0.8 * h), (9, 9, 9), name='B')
    c = G.cart((0.52, 0.18, 0.21), (1.1 * h, 1.1 * h, 1.1 * h), (7, 7, 7), name='C')
    t = C.newPyTree(['Base', a, b, c])
    if neq == 6:
        Internal.createChild(Internal.getBases(t)[0], 'FlowEquationSet', 'FlowEquationSet_t',
                             children=[['GoverningEquations', np.frombuffer(b'NSTurbulent', dtype='S1').copy(), [],
                                        'GoverningEquations_t']])
    names = VARS5 if neq == 5 else VARS6
    rng = np.random.default_rng(neq)
    for z in Internal.getZones(t):
        shp = C.fieldShape(z, 'centers')
        for v in names:
            C._initVars(z, 'centers:' + v, np.asfortranarray(rng.standard_normal(shp)))
    C._initVars(t, 'centers:cellN', 1.)
    zb, zc = Internal.getZones(t)[1:]
    C._initVars(zb, 'centers:cellN', 2.)
    cn = np.ones(C.fieldShape(zc, 'centers'), order='F'); cn[:2] = 2.; cn[:, -2:] = 2.
    C._initVars(zc, 'centers:cellN', cn)
    return t, names


def _prepare(kind, neq, backend=None, ctx=None):
    t, names = _tree(kind, neq)
    tc = C.node2Center(t)
    X._setInterpDataChimera(t, tc, order=2, nature=1, penalty=1, extrap=1, loc='centers', storage='inverse', sameName=1,
                            verbose=0, backend=backend, ctx=ctx)
    if neq == 6:
        b = Internal.getBases(tc)[0]
        if Internal.getNodeFromName2(b, 'GoverningEquations') is None:
            Internal.createChild(b, 'FlowEquationSet', 'FlowEquationSet_t',
                                 children=[['GoverningEquations', np.frombuffer(b'NSTurbulent', dtype='S1').copy(), [],
                                            'GoverningEquations_t']])
    raw = {}
    for zd in Internal.getZones(tc):
        for s in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
            raw[(zd[0], s[0])] = tuple(np.array(Internal.getNodeFromName1(s, n)[1], copy=True) for n in
                                       ('PointList', 'PointListDonor', 'InterpolantsType', 'InterpolantsDonor'))
    blocksR = [CT.compactFields(z, names, Internal.__FlowSolutionCenters__) for z in Internal.getZones(t)]
    blocksD = [CT.compactFields(z, names, Internal.__FlowSolutionNodes__) for z in Internal.getZones(tc)]
    npass = CT.miseAPlatDonorTree__(t, tc, graph=None)
    return t, tc, names, raw, blocksR, blocksD, npass


def _dtloc(nzones, modified, nssiter=3, nthreads=4):
    """dtloc as the reader walks it (setInterpTransfers.cpp:930-966): task lists per sub-iteration; every
    sub-iteration modifies the zones of `modified`."""
    rec = 6 + nthreads * 7
    shift_omp = 40
    a = np.zeros(shift_omp + 2 * nssiter + nssiter * len(modified) * rec + 8, np.int32)
    a[0] = nssiter; a[9] = 1; a[10] = 0; a[11] = shift_omp
    pt = 2 * nssiter
    for it in range(nssiter):
        a[shift_omp + it] = len(modified)
        a[shift_omp + nssiter + it] = pt
        for k, nd in enumerate(modified):
            a[shift_omp + pt + k * rec] = nd
        pt += len(modified) * rec
    return a


@pytest.mark.parametrize("kind,neq", [('coincident', 5), ('shifted', 5), ('shifted', 6)])
def test_packer_layout_is_read_by_the_reference_reader(kind, neq):
    from oracle import oracle as O
    t, tc, names, raw, blocksR, blocksD, npass = _prepare(kind, neq, backend=OracleBackend())
    pi = Internal.getNodeFromName1(tc, 'Parameter_int')[1]; pr = Internal.getNodeFromName1(tc, 'Parameter_real')[1]
    zones = Internal.getZones(t); zonesD = Internal.getZones(tc)
    nrac = sum(len(Internal.getNodesFromType1(z, 'ZoneSubRegion_t')) for z in zonesD)
    assert nrac >= 2 and npass == 1
    # header (compactTransfers.py:375-420, reader setInterpTransfers.cpp:908-918)
    assert pi[0] == 0 and pi[1] == 1 and pi[2] == 0 and pi[3] == 0
    ech = pi[1 + 3]
    assert tuple(pi[ech:ech + 4]) == (0, nrac, 0, 0)
    # the subregions now alias the flattened arrays, regrouped by type (stable)
    for zd in zonesD:
        for s in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
            pl0, pld0, ty0, cf0 = raw[(zd[0], s[0])]
            ty = Internal.getNodeFromName1(s, 'InterpolantsType')[1]
            assert np.shares_memory(ty, pi) and np.shares_memory(Internal.getNodeFromName1(s, 'InterpolantsDonor')[1], pr)
            seen = []
            for v in ty0.tolist():
                if v not in seen: seen.append(v)
            perm = np.argsort(np.array([seen.index(v) for v in ty0.tolist()]), kind='stable')
            assert np.array_equal(ty, ty0[perm])
            assert np.array_equal(Internal.getNodeFromName1(s, 'PointList')[1], pl0[perm])
            assert np.array_equal(Internal.getNodeFromName1(s, 'PointListDonor')[1], pld0[perm])
    if kind == 'coincident':
        assert any((v[2] == 1).any() for v in raw.values())
    # reference reader on the flattened tree == plain reference loop on the unpacked lists (type 1: injection)
    zparams = [Internal.getNodeFromName2(z, 'Parameter_int')[1] for z in zones]
    want = [b.copy() for b in blocksR]
    nd = {z[0]: i for i, z in enumerate(zones)}
    for zd in zonesD:                               # rac order = donor zones in tree order, subregions in node order
        _, imd, jmd, kmd, _ = Internal.getZoneDim(zd)
        for s in Internal.getNodesFromType1(zd, 'ZoneSubRegion_t'):
            pl0, pld0, ty0, cf0 = raw[(zd[0], s[0])]
            cf1 = cf0.copy()
            ncf = np.array([{1: 1, 2: 8}[int(v)] for v in ty0]); st = np.concatenate(([0], np.cumsum(ncf)[:-1]))
            cf1[st[ty0 == 1]] = 1.
            iD = nd[zd[0]]; iR = nd[Internal.getValue(s)]
            nD = blocksD[iD].size // neq; nR = want[iR].size // neq
            fD = [blocksD[iD][e * nD:(e + 1) * nD] for e in range(neq)]
            fR = [want[iR][e * nR:(e + 1) * nR] for e in range(neq)]
            O.transfer(fD, fR, imd, jmd, pld0, pl0, ty0, cf1)
    got = [b.copy() for b in blocksR]
    done = O.transfer_compact(got, blocksD, zparams, _dtloc(3, [0, 1, 2]), pi, pr, vartype=2 if neq == 5 else 21,
                              nstep=2, threadmax_sdm=4)
    assert done == nrac
    for g, w, b in zip(got, want, blocksR):
        assert np.array_equal(g, w) and not np.array_equal(g, b) or np.array_equal(w, b)
    # only the racs whose DONOR zone was modified at this sub-iteration are transferred
    got2 = [b.copy() for b in blocksR]
    done2 = O.transfer_compact(got2, blocksD, zparams, _dtloc(3, [1]), pi, pr, vartype=2 if neq == 5 else 21, nstep=1,
                               threadmax_sdm=4)
    assert done2 == len(Internal.getNodesFromType1(zonesD[1], 'ZoneSubRegion_t'))


def test_compact_argument_errors():
    t, tc, names, raw, blocksR, blocksD, npass = _prepare('shifted', 5, backend=OracleBackend())
    pi = Internal.getNodeFromName1(tc, 'Parameter_int')[1]; pr = Internal.getNodeFromName1(tc, 'Parameter_real')[1]
    zones = Internal.getZones(t); zonesD = Internal.getZones(tc)
    with pytest.raises(NotImplementedError):
        CT.___setInterpTransfers(zones, zonesD, names, _dtloc(3, [0]), pi, pr, 0, 2, 0, 1, 1, 3, 3, 1, 1, 0, ctx=object())
    with pytest.raises(NotImplementedError):
        CT.___setInterpTransfers(zones, zonesD, names, _dtloc(3, [0]), pi, pr, 0, 4, 0, 1, 1, 3, 3, 0, 1, 0, ctx=object())
    with pytest.raises(NotImplementedError):
        CT.miseAPlatDonorTree__(t, tc, graph={'procDict': {}})
    # the IBC-only pass has nothing to do on this path
    assert CT.___setInterpTransfers(zones, zonesD, names, _dtloc(3, [0]), pi, pr, 0, 2, 1, 1, 1, 3, 3, 0, 1, 0, ctx=object()) is None


@pytest.mark.gpu
@pytest.mark.parametrize("kind,neq", [('coincident', 5), ('shifted', 5), ('shifted', 6)])
def test_compact_transfers_device_vs_reference_reader(ctx, kind, neq):
    """___setInterpTransfers on the device == the reference reader (its own reorder fragments) on the same flattened
    tree, every rac, 5 and 6 equations, type-1 injection included; then the sub-iteration filter of dtloc and a second
    call (plans cached on param_int)."""
    from oracle import oracle as O
    t, tc, names, raw, blocksR, blocksD, npass = _prepare(kind, neq, ctx=ctx)
    pi = Internal.getNodeFromName1(tc, 'Parameter_int')[1]; pr = Internal.getNodeFromName1(tc, 'Parameter_real')[1]
    zones = Internal.getZones(t); zonesD = Internal.getZones(tc)
    zparams = [Internal.getNodeFromName2(z, 'Parameter_int')[1] for z in zones]
    vt = 2 if neq == 5 else 21
    for modified, nstep in (([0, 1, 2], 2), ([1], 1), ([0, 2], 3)):
        dt = _dtloc(3, modified)
        want = [b.copy() for b in blocksR]
        O.transfer_compact(want, blocksD, zparams, dt, pi, pr, vartype=vt, nstep=nstep, threadmax_sdm=4)
        before = [b.copy() for b in blocksR]
        CT.___setInterpTransfers(zones, zonesD, names, dt, pi, pr, 0, vt, 0, 1, nstep, 3, 3, 0, 1, 0, ctx=ctx,
                                 threadmax_sdm=4)
        for g, w in zip(blocksR, want):
            assert np.array_equal(g, w)
        # the tree's arrays are the compact blocks (views): the result is visible through the tree
        for z, b in zip(zones, blocksR):
            assert np.shares_memory(C.getFieldArray(z, 'centers:' + names[0]), b)
        for b, v in zip(blocksR, before):
            b[:] = v
    # Connector.Mpi.__setInterpTransfers, intra-process part
    dt = _dtloc(3, [0, 1, 2])
    want = [b.copy() for b in blocksR]
    O.transfer_compact(want, blocksD, zparams, dt, pi, pr, vartype=vt, nstep=1, threadmax_sdm=4)
    import os
    os.environ["OMP_NUM_THREADS"] = "4"
    try:
        getattr(CT, '__setInterpTransfers')(zones, zonesD, names, dt, pi, pr, 0, 0, 1, 3, 3, 0, 1, varType=vt, ctx=ctx)
    finally:
        del os.environ["OMP_NUM_THREADS"]
    for g, w in zip(blocksR, want):
        assert np.array_equal(g, w)
