"""Known-answer tests that pin the order-3/5 part of the oracle (oracle/lagrange_restated.c, a C restatement of four
Fortran routines the image cannot compile) independently of the transliteration itself:

* kgaussj: pivot SEQUENCE on matrices with tied maxima, derived by hand from the text of
  KCore/KCore/Linear/GaussjF.for:34-50 (scan rows j then columns k, `.GE.` keeps the LAST maximum), and exact
  solutions of systems whose pivots are powers of two;
* compinterpolatedptinrefelt on an affinely sheared stencil: the stable frame of
  Cp_coord_in_stable_frame_3dF.for is then the affine frame itself, so (ksi, eta, zeta) of a point
  O + a u + b v + c w are (a, b, c) -- closed form, no reference run needed -- and the `abs(ksi) >= max + 1e-3`
  exit (CompInterpolatedPtInRefElementF.for:286-299) flips exactly at the threshold;
* real curvilinear stencils (cylinder O-grid of BASELINE config 2): the same polynomial map solved with mpmath at 60
  digits (frame by the formulas of Coord_in_ref_frame_3dF.for:23-127, i.e. M^-1 t) bounds the double-precision answer
  by condition number x epsilon.

None of this compares the restatement with a gfortran build (there is none here), so SURVEY row a9 stays flagged
"pinned to the transliteration"; what these tests rule out is a wrong tie rule, a wrong frame, a wrong monomial /
right-hand-side order or a wrong threshold."""
import ctypes
import numpy as np
import pytest

from oracle import oracle as O


def _gaussj(A, B):
    """kgaussj_ on copies (column-major like the Fortran); returns (A^-1, X, ok, indxr, indxc)."""
    L = O.lib()
    n, m = A.shape[0], B.shape[1]
    a = np.asfortranarray(A, dtype=np.float64).copy(order='F'); b = np.asfortranarray(B, dtype=np.float64).copy(order='F')
    ok = ctypes.c_int(0); nn = ctypes.c_int(n); mm = ctypes.c_int(m)
    ic = np.zeros(n, np.int32); ir = np.zeros(n, np.int32); ip = np.zeros(n, np.int32)
    vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    L.kgaussj_.argtypes = [ctypes.c_void_p] * 8
    L.kgaussj_(vp(a), ctypes.addressof(nn), vp(b), ctypes.addressof(mm), ctypes.addressof(ok), vp(ic), vp(ir), vp(ip))
    return a, b, ok.value, ir, ic


def test_gaussj_tie_keeps_last_maximum_diagonal():
    # |a| = 4 at (1,1), (2,2), (3,3): the scan of GaussjF.for:34-44 with .GE. ends on (3,3); among rows/columns 1,2
    # the tie (1,1) / (2,2) ends on (2,2); then (1,1).  A strict '>' would give 1,2,3.
    A = np.array([[4., 1., 0.], [1., 4., 0.], [0., 0., 4.]])
    B = np.array([[5.], [5.], [4.]])
    inv, x, ok, ir, ic = _gaussj(A, B)
    assert ok == 1
    assert list(ir) == [3, 2, 1] and list(ic) == [3, 2, 1]
    # every pivot inverse is exact (1/4, 1/4, 1/3.75 applied to 3.75): exact solution
    assert np.array_equal(x[:, 0], [1., 1., 1.])
    assert np.allclose(inv @ A, np.eye(3), atol=1e-15)


def test_gaussj_tie_off_diagonal_swaps_rows():
    # maxima 2 at (1,2) and (2,1): last in scan order is row 2, column 1 -> rows 2 and 1 are swapped first
    A = np.array([[1., 2.], [2., 0.]])
    B = np.array([[4.], [2.]])
    inv, x, ok, ir, ic = _gaussj(A, B)
    assert ok == 1
    assert (ir[0], ic[0]) == (2, 1) and (ir[1], ic[1]) == (2, 2)
    assert np.array_equal(x[:, 0], [1., 1.5])           # pivots 2 and 2: exact
    assert np.array_equal(inv, np.array([[0., 0.5], [0.5, -0.25]]))


def test_gaussj_sign_and_pivoted_rows_are_skipped():
    # |-8| wins first at (2,3) (row 2, column 3); afterwards ROW 3 (= the pivoted column index) is excluded from the
    # search (ipiv(j).NE.1 tests the row index) as well as column 3: the 7 at (3,1) is never a candidate in step 2,
    # which must pick the last maximum among rows {1,2} x columns {1,2}
    A = np.array([[2., 1., 0.], [2., 2., -8.], [7., 0., 1.]])
    B = np.eye(3)
    inv, x, ok, ir, ic = _gaussj(A, B)
    assert ok == 1
    assert (ir[0], ic[0]) == (2, 3)
    # after the swap rows are: [2,1,0], [7,0,1], [2,2,-8]/-8 ...; the second search runs on rows 1,2, columns 1,2
    assert ic[1] in (1, 2) and ir[1] in (1, 2)
    assert np.allclose(inv @ A, np.eye(3), atol=1e-14) and np.allclose(x, inv, atol=0)


def test_gaussj_singular_exit():
    A = np.array([[1., 2.], [2., 4.]])
    inv, x, ok, ir, ic = _gaussj(A, np.ones((2, 1)))
    assert ok == 0          # 'singular matrix in gaussj, 2': a(icol,icol) == 0 after the first elimination


def _ref_elt(xt, yt, zt, ni, nj, ic, jc, kc, p, n1d):
    """compinterpolatedptinrefelt_ (1-based stencil start ic,jc,kc) -> (ksi, eta, zeta, erreur)."""
    L = O.lib()
    n3d = n1d ** 3
    ci = ctypes.c_int; cd = ctypes.c_double
    args_i = [ci(xt.size), ci(ic), ci(jc), ci(kc), ci(ni), ci(nj)]
    x0, y0, z0 = cd(p[0]), cd(p[1]), cd(p[2])
    c1, c3 = ci(n1d), ci(n3d)
    xi = np.zeros(n3d); yi = np.zeros(n3d); zi = np.zeros(n3d); base = np.zeros(n1d)
    A1 = np.zeros(n3d * n3d); B1 = np.zeros(n3d * 3)
    ixc = np.zeros(n3d, np.int32); ixr = np.zeros(n3d, np.int32); ipv = np.zeros(n3d, np.int32)
    xx, yy, zz, err = cd(0), cd(0), cd(0), ci(0)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ad = ctypes.addressof
    L.compinterpolatedptinrefelt_.argtypes = [ctypes.c_void_p] * 27
    L.compinterpolatedptinrefelt_(vp(xt), vp(yt), vp(zt), *[ad(a) for a in args_i], ad(x0), ad(y0), ad(z0), ad(c1), ad(c3),
                                  vp(xi), vp(yi), vp(zi), vp(base), vp(A1), vp(B1), vp(ixc), vp(ixr), vp(ipv),
                                  ad(xx), ad(yy), ad(zz), ad(err))
    return xx.value, yy.value, zz.value, err.value


def _affine_block(n, O0, u, v, w):
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing='ij')
    P = O0[None, None, None, :] + i[..., None] * u + j[..., None] * v + k[..., None] * w
    f = lambda a: np.ascontiguousarray(a.reshape(-1, order='F'))
    return f(P[..., 0]), f(P[..., 1]), f(P[..., 2])


@pytest.mark.parametrize("n1d", [3, 5])
def test_ref_element_on_sheared_stencil_closed_form(n1d):
    rng = np.random.default_rng(n1d)
    O0 = np.array([0.3, -1.2, 2.5])
    u = np.array([0.11, 0.02, -0.01]); v = np.array([0.03, 0.09, 0.015]); w = np.array([-0.02, 0.025, 0.13])   # sheared
    n = n1d + 2
    xt, yt, zt = _affine_block(n, O0, u, v, w)
    ic = jc = kc = 2                                # 1-based stencil start: nodes 1..n1d (0-based)
    c = 1 + (n1d + 1) // 2 - 1                      # 0-based index of the stencil's centre node in the block
    centre = O0 + c * (u + v + w)
    half = n1d // 2
    for _ in range(20):
        abc = rng.uniform(-0.95 * half, 0.95 * half, 3)
        p = centre + abc[0] * u + abc[1] * v + abc[2] * w
        ksi, eta, zeta, err = _ref_elt(xt, yt, zt, n, n, ic, jc, kc, p, n1d)
        assert err == 0
        assert np.allclose([ksi, eta, zeta], abc, rtol=0, atol=(2e-12 if n1d == 3 else 2e-9))
    # the exit at abs(ksi) >= n1d/2 + 0.001 (integer division)
    for a, want in ((half + 0.0011, 1), (half + 0.0009, 0), (-(half + 0.0011), 1)):
        p = centre + a * u + 0.1 * v - 0.2 * w
        assert _ref_elt(xt, yt, zt, n, n, ic, jc, kc, p, n1d)[3] == want


def _mp_ref_elt(X, Y, Z, p, n1d, mp):
    """The polynomial map of CompInterpolatedPtInRefElementF.for in exact-ish arithmetic (mpmath): stable frame
    (centre node, averaged edge vectors), coordinates M^-1 (x - O), monomial matrix, base right-hand sides, solve,
    evaluate at the receptor.  X, Y, Z: (n1d,n1d,n1d) arrays indexed [i,j,k]."""
    c = (n1d + 1) // 2 - 1
    P = lambda i, j, k: mp.matrix([mp.mpf(float(X[i, j, k])), mp.mpf(float(Y[i, j, k])), mp.mpf(float(Z[i, j, k]))])
    Oc = P(c, c, c)
    ui = sum((P(c + 1, c + dj, c + dk) - P(c, c + dj, c + dk) for dj in (0, 1) for dk in (0, 1)), mp.zeros(3, 1)) / 4
    uj = sum((P(c + di, c + 1, c + dk) - P(c + di, c, c + dk) for di in (0, 1) for dk in (0, 1)), mp.zeros(3, 1)) / 4
    uk = sum((P(c + di, c + dj, c + 1) - P(c + di, c + dj, c) for di in (0, 1) for dj in (0, 1)), mp.zeros(3, 1)) / 4
    M = mp.matrix(3, 3)
    for r in range(3):
        M[r, 0] = ui[r]; M[r, 1] = uj[r]; M[r, 2] = uk[r]
    Minv = M ** -1
    loc = lambda q: Minv * (q - Oc)
    n3d = n1d ** 3
    A = mp.matrix(n3d, n3d); B = mp.matrix(n3d, 3)
    base = [mp.mpf(b) for b in ((-1, 0, 1) if n1d == 3 else (-2, -1, 0, 1, 2))]
    row = 0
    for k in range(n1d):
        for j in range(n1d):
            for i in range(n1d):
                q = loc(P(i, j, k))
                col = 0
                for kk in range(n1d):
                    for jj in range(n1d):
                        for ii in range(n1d):
                            A[row, col] = q[0] ** ii * q[1] ** jj * q[2] ** kk
                            col += 1
                B[row, 0] = base[i]; B[row, 1] = base[j]; B[row, 2] = base[k]
                row += 1
    cols = [mp.lu_solve(A, B[:, d]) for d in range(3)]
    S = mp.matrix(n3d, 3)
    for d in range(3):
        for r in range(n3d):
            S[r, d] = cols[d][r]
    q = loc(mp.matrix([mp.mpf(float(p[0])), mp.mpf(float(p[1])), mp.mpf(float(p[2]))]))
    out = [mp.mpf(0)] * 3
    col = 0
    for kk in range(n1d):
        for jj in range(n1d):
            for ii in range(n1d):
                mono = q[0] ** ii * q[1] ** jj * q[2] ** kk
                for d in range(3):
                    out[d] += S[col, d] * mono
                col += 1
    Af = np.array([[float(A[r, c2]) for c2 in range(n3d)] for r in range(n3d)])
    return [float(o) for o in out], np.linalg.cond(Af)


@pytest.mark.parametrize("n1d,nst", [(3, 40), (5, 3)])
def test_ref_element_on_cylinder_stencils_vs_mpmath(n1d, nst):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 60
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Converter as C
    z = G.cylinder((0., 0., 0.), 0.5, 2.5, 360., 0., 2., (65, 17, 33))       # the O-grid of config 2, coarsened
    X, Y, Z = C.getCoords(z)
    ni, nj, nk = X.shape
    xt, yt, zt = [np.ascontiguousarray(a.reshape(-1, order='F')) for a in (X, Y, Z)]
    rng = np.random.default_rng(100 + n1d)
    worst = 0.
    for _ in range(nst):
        i0 = int(rng.integers(0, ni - n1d)); j0 = int(rng.integers(0, nj - n1d)); k0 = int(rng.integers(0, nk - n1d))
        sl = (slice(i0, i0 + n1d), slice(j0, j0 + n1d), slice(k0, k0 + n1d))
        c = (n1d + 1) // 2 - 1
        # receptor: a random point of the stencil's centre cell (trilinear blend of its 8 nodes)
        t = rng.random(3)
        p = np.zeros(3)
        for di in (0, 1):
            for dj in (0, 1):
                for dk in (0, 1):
                    wgt = (t[0] if di else 1 - t[0]) * (t[1] if dj else 1 - t[1]) * (t[2] if dk else 1 - t[2])
                    p += wgt * np.array([X[i0 + c + di, j0 + c + dj, k0 + c + dk], Y[i0 + c + di, j0 + c + dj, k0 + c + dk],
                                         Z[i0 + c + di, j0 + c + dj, k0 + c + dk]])
        got = _ref_elt(xt, yt, zt, ni, nj, i0 + 1, j0 + 1, k0 + 1, p, n1d)
        want, cond = _mp_ref_elt(X[sl], Y[sl], Z[sl], p, n1d, mp)
        err = max(abs(g - w) for g, w in zip(got[:3], want))
        assert got[3] == 0
        assert err <= 64 * cond * np.finfo(float).eps, (err, cond)
        # the receptor sits in the centre cell: reference coordinates in [0, 1] (up to the curvature of the cell)
        assert all(-0.05 <= w <= 1.05 for w in want)
        worst = max(worst, err)
    print("order %d: worst |double - mpmath| = %.3g" % (n1d, worst))
