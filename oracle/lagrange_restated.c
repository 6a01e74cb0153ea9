/* ORACLE — TEST INFRASTRUCTURE ONLY. Never imported, linked or called by the
 * product path (cassiopee_b200/). Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, as the checker.
 *
 * C restatement of the four Fortran routines the reference's order-3/5
 * Lagrange path needs (the image has no gfortran, so they cannot be compiled
 * from /root/reference). Exported with the gfortran calling convention
 * (trailing underscore, every argument by reference) so the reference's own
 * C++ (KCore/KCore/Interp/Interp2.cpp:26-38, :564-570) links against them
 * unchanged inside oracle/_ref/libcassref.so.
 *
 * Follows, statement by statement:
 *   compinterpolatedptinrefelt  KCore/KCore/Interp/CompInterpolatedPtInRefElementF.for:25-300
 *   cp_coord_in_stable_frame_3d KCore/KCore/Interp/Cp_coord_in_stable_frame_3dF.for:25-174
 *   coord_in_ref_frame_3d       KCore/KCore/Interp/Coord_in_ref_frame_3dF.for:23-127
 *   determinant (ni=3 branch)   KCore/KCore/Interp/Coord_in_ref_frame_3dF.for:131-228
 *   kgaussj                     KCore/KCore/Linear/GaussjF.for:1-103
 *   ind_interp                  KCore/KCore/Interp/IndInterpF.h:50-54
 *
 * Arrays keep Fortran column-major indexing: A(i,j) -> a[(i-1) + (j-1)*n].
 * Expressions are evaluated left to right as written (gfortran without
 * -ffast-math does not reassociate); compile WITHOUT -ffast-math and with
 * -ffp-contract=off.  E_Int is 32-bit int (no E_DOUBLEINT), REAL_E is double.
 *
 * Parity note: this file is a transliteration, not the reference binary, so
 * order-3/5 coefficient parity is pinned to "reference C++ + this C".
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

typedef int E_Int;
typedef double E_Float;

/* IndInterpF.h:50-54 (1-based result) */
static inline E_Int ind_interp(E_Int n1d, E_Int i, E_Int j, E_Int k)
{
  return i + (j - 1) * n1d + (k - 1) * n1d * n1d;
}

/* Coord_in_ref_frame_3dF.for:205-211 */
static E_Float determinant3(const E_Float mat[9])
{
#define M(i, j) mat[(i - 1) + (j - 1) * 3]
  return M(1, 1) * M(2, 2) * M(3, 3) + M(2, 1) * M(3, 2) * M(1, 3) +
         M(3, 1) * M(1, 2) * M(2, 3) - M(1, 1) * M(3, 2) * M(2, 3) -
         M(2, 1) * M(1, 2) * M(3, 3) - M(3, 1) * M(2, 2) * M(1, 3);
#undef M
}

/* Coord_in_ref_frame_3dF.for:23-127; pts is (3,4) column-major */
static void coord_in_ref_frame_3d(const E_Float pts[12], E_Float x1, E_Float y1,
                                  E_Float z1, E_Float* x2, E_Float* y2,
                                  E_Float* z2)
{
#define PTS(i, j) pts[(i - 1) + (j - 1) * 3]
#define MAT(i, j) mat[(i - 1) + (j - 1) * 3]
  E_Float t[3], u[3], v[3], w[3], mat[9], invdet;
  for (int i = 1; i <= 3; i++)
  {
    u[i - 1] = PTS(i, 2) - PTS(i, 1);
    v[i - 1] = PTS(i, 3) - PTS(i, 1);
    w[i - 1] = PTS(i, 4) - PTS(i, 1);
  }
  t[0] = x1 - PTS(1, 1);
  t[1] = y1 - PTS(2, 1);
  t[2] = z1 - PTS(3, 1);
  for (int i = 1; i <= 3; i++)
  {
    MAT(i, 1) = u[i - 1]; MAT(i, 2) = v[i - 1]; MAT(i, 3) = w[i - 1];
  }
  invdet = determinant3(mat);
  invdet = 1.0 / invdet;
  for (int i = 1; i <= 3; i++)
  {
    MAT(i, 1) = t[i - 1]; MAT(i, 2) = v[i - 1]; MAT(i, 3) = w[i - 1];
  }
  *x2 = determinant3(mat);
  *x2 = *x2 * invdet;
  for (int i = 1; i <= 3; i++)
  {
    MAT(i, 1) = u[i - 1]; MAT(i, 2) = t[i - 1]; MAT(i, 3) = w[i - 1];
  }
  *y2 = determinant3(mat);
  *y2 = *y2 * invdet;
  for (int i = 1; i <= 3; i++)
  {
    MAT(i, 1) = u[i - 1]; MAT(i, 2) = v[i - 1]; MAT(i, 3) = t[i - 1];
  }
  *z2 = determinant3(mat);
  *z2 = *z2 * invdet;
#undef PTS
#undef MAT
}

/* Cp_coord_in_stable_frame_3dF.for:25-174 */
static void cp_coord_in_stable_frame_3d(E_Int n1d, E_Int n3d, E_Float* xi,
                                        E_Float* yi, E_Float* zi, E_Float* xp,
                                        E_Float* yp, E_Float* zp)
{
  E_Int i0 = (n1d + 1) / 2, j0 = i0, k0 = i0;
  E_Float pt_O[3], u_i[3] = {0., 0., 0.}, u_j[3] = {0., 0., 0.},
                   u_k[3] = {0., 0., 0.};
  E_Float pts[12], xtmp, ytmp, ztmp;
  E_Int id, id1, id2;
#define X(i) xi[(i)-1]
#define Y(i) yi[(i)-1]
#define Z(i) zi[(i)-1]
  id = ind_interp(n1d, i0, j0, k0);
  pt_O[0] = X(id); pt_O[1] = Y(id); pt_O[2] = Z(id);
  for (int dk = 0; dk <= 1; dk++)
    for (int dj = 0; dj <= 1; dj++)
    {
      id1 = ind_interp(n1d, i0, j0 + dj, k0 + dk);
      id2 = ind_interp(n1d, i0 + 1, j0 + dj, k0 + dk);
      u_i[0] = u_i[0] + X(id2) - X(id1);
      u_i[1] = u_i[1] + Y(id2) - Y(id1);
      u_i[2] = u_i[2] + Z(id2) - Z(id1);
    }
  for (int i = 0; i < 3; i++) u_i[i] = u_i[i] * 0.25;
  for (int dk = 0; dk <= 1; dk++)
    for (int di = 0; di <= 1; di++)
    {
      id1 = ind_interp(n1d, i0 + di, j0, k0 + dk);
      id2 = ind_interp(n1d, i0 + di, j0 + 1, k0 + dk);
      u_j[0] = u_j[0] + X(id2) - X(id1);
      u_j[1] = u_j[1] + Y(id2) - Y(id1);
      u_j[2] = u_j[2] + Z(id2) - Z(id1);
    }
  for (int i = 0; i < 3; i++) u_j[i] = u_j[i] * 0.25;
  for (int dj = 0; dj <= 1; dj++)
    for (int di = 0; di <= 1; di++)
    {
      id1 = ind_interp(n1d, i0 + di, j0 + dj, k0);
      id2 = ind_interp(n1d, i0 + di, j0 + dj, k0 + 1);
      u_k[0] = u_k[0] + X(id2) - X(id1);
      u_k[1] = u_k[1] + Y(id2) - Y(id1);
      u_k[2] = u_k[2] + Z(id2) - Z(id1);
    }
  for (int i = 0; i < 3; i++) u_k[i] = u_k[i] * 0.25;
  for (int i = 0; i < 3; i++)
  {
    pts[i + 0] = pt_O[i];
    pts[i + 3] = pt_O[i] + u_i[i];
    pts[i + 6] = pt_O[i] + u_j[i];
    pts[i + 9] = pt_O[i] + u_k[i];
  }
  for (int i = 1; i <= n3d; i++)
  {
    coord_in_ref_frame_3d(pts, X(i), Y(i), Z(i), &xtmp, &ytmp, &ztmp);
    X(i) = xtmp; Y(i) = ytmp; Z(i) = ztmp;
  }
  coord_in_ref_frame_3d(pts, *xp, *yp, *zp, &xtmp, &ytmp, &ztmp);
  *xp = xtmp; *yp = ytmp; *zp = ztmp;
#undef X
#undef Y
#undef Z
}

/* GaussjF.for — Numerical-Recipes full-pivot Gauss-Jordan; a(n,n), b(n,m)
 * column-major; `>=` keeps the LAST maximum (GaussjF.for:38). */
void kgaussj_(E_Float* a, const E_Int* pn, E_Float* b, const E_Int* pm,
              E_Int* ok, E_Int* indxc, E_Int* indxr, E_Int* ipiv)
{
  const E_Int n = *pn, m = *pm;
#define A(i, j) a[((i)-1) + (size_t)((j)-1) * n]
#define B(i, j) b[((i)-1) + (size_t)((j)-1) * n]
  E_Int icol = 0, irow = 0;
  E_Float big, dum, pivinv;
  *ok = 1;
  for (E_Int j = 1; j <= n; j++) ipiv[j - 1] = 0;
  for (E_Int i = 1; i <= n; i++)
  {
    big = 0.;
    for (E_Int j = 1; j <= n; j++)
    {
      if (ipiv[j - 1] != 1)
      {
        for (E_Int k = 1; k <= n; k++)
        {
          if (ipiv[k - 1] == 0)
          {
            if (fabs(A(j, k)) >= big)
            {
              big = fabs(A(j, k));
              irow = j;
              icol = k;
            }
          }
          else if (ipiv[k - 1] > 1)
          {
            printf(" singular matrix in gaussj, 1\n");
            *ok = 0;
            return;
          }
        }
      }
    }
    ipiv[icol - 1] = ipiv[icol - 1] + 1;
    if (irow != icol)
    {
      for (E_Int l = 1; l <= n; l++)
      {
        dum = A(irow, l); A(irow, l) = A(icol, l); A(icol, l) = dum;
      }
      for (E_Int l = 1; l <= m; l++)
      {
        dum = B(irow, l); B(irow, l) = B(icol, l); B(icol, l) = dum;
      }
    }
    indxr[i - 1] = irow;
    indxc[i - 1] = icol;
    if (A(icol, icol) == 0.)
    {
      printf(" singular matrix in gaussj, 2\n");
      *ok = 0;
      return;
    }
    pivinv = 1. / A(icol, icol);
    A(icol, icol) = 1.;
    for (E_Int l = 1; l <= n; l++) A(icol, l) = A(icol, l) * pivinv;
    for (E_Int l = 1; l <= m; l++) B(icol, l) = B(icol, l) * pivinv;
    for (E_Int ll = 1; ll <= n; ll++)
    {
      if (ll != icol)
      {
        dum = A(ll, icol);
        A(ll, icol) = 0.;
        for (E_Int l = 1; l <= n; l++) A(ll, l) = A(ll, l) - A(icol, l) * dum;
        for (E_Int l = 1; l <= m; l++) B(ll, l) = B(ll, l) - B(icol, l) * dum;
      }
    }
  }
  for (E_Int l = n; l >= 1; l--)
  {
    if (indxr[l - 1] != indxc[l - 1])
    {
      for (E_Int k = 1; k <= n; k++)
      {
        dum = A(k, indxr[l - 1]);
        A(k, indxr[l - 1]) = A(k, indxc[l - 1]);
        A(k, indxc[l - 1]) = dum;
      }
    }
  }
#undef A
#undef B
}

/* CompInterpolatedPtInRefElementF.for:25-300 */
void compinterpolatedptinrefelt_(
    const E_Float* xt, const E_Float* yt, const E_Float* zt, const E_Int* npts,
    const E_Int* pic, const E_Int* pjc, const E_Int* pkc, const E_Int* pni,
    const E_Int* pnj, const E_Float* x0, const E_Float* y0, const E_Float* z0,
    const E_Int* pn1d, const E_Int* pn3d, E_Float* x_interp, E_Float* y_interp,
    E_Float* z_interp, E_Float* base, E_Float* A1, E_Float* B1, E_Int* indxc,
    E_Int* indxr, E_Int* ipiv, E_Float* xx, E_Float* yy, E_Float* zz,
    E_Int* erreur)
{
  (void)npts;
  const E_Int ic = *pic, jc = *pjc, kc = *pkc, ni = *pni, nj = *pnj;
  const E_Int n1d = *pn1d, n3d = *pn3d;
  const E_Int ninj = ni * nj;
  const E_Int max = n1d / 2;
  const E_Float eps = 0.001;
  const E_Float thres = max + eps;
  E_Float xp = *x0, yp = *y0, zp = *z0;
  E_Float xp1, yp1, zp1, mi, mj, mk, m;
  E_Int id_pt, id_col, ind, ok;
#define AA(i, j) A1[((i)-1) + (size_t)((j)-1) * n3d]
#define BB(i, j) B1[((i)-1) + (size_t)((j)-1) * n3d]
  *erreur = 0;
  const E_Int ind0 = ic - 1 + (jc - 1) * ni + (kc - 1) * ninj;
  id_pt = 1;
  for (E_Int k = 1; k <= n1d; k++)
    for (E_Int j = 1; j <= n1d; j++)
      for (E_Int i = 1; i <= n1d; i++)
      {
        ind = ind0 + i - 1 + (j - 1) * ni + (k - 1) * ninj;
        x_interp[id_pt - 1] = xt[ind];
        y_interp[id_pt - 1] = yt[ind];
        z_interp[id_pt - 1] = zt[ind];
        id_pt = id_pt + 1;
      }
  cp_coord_in_stable_frame_3d(n1d, n3d, x_interp, y_interp, z_interp, &xp, &yp,
                              &zp);
  for (id_pt = 1; id_pt <= n3d; id_pt++)
  {
    id_col = 1;
    zp1 = 1;
    for (E_Int k = 1; k <= n1d; k++)
    {
      yp1 = 1;
      for (E_Int j = 1; j <= n1d; j++)
      {
        xp1 = 1;
        for (E_Int i = 1; i <= n1d; i++)
        {
          AA(id_pt, id_col) = xp1 * yp1 * zp1;
          id_col = id_col + 1;
          xp1 = xp1 * x_interp[id_pt - 1];
        }
        yp1 = yp1 * y_interp[id_pt - 1];
      }
      zp1 = zp1 * z_interp[id_pt - 1];
    }
  }
  if (n1d == 2) { base[0] = 0.; base[1] = 1.; }
  else if (n1d == 3) { base[0] = -1.; base[1] = 0.; base[2] = 1.; }
  else if (n1d == 4) { base[0] = -1.; base[1] = 0.; base[2] = 1.; base[3] = 2.; }
  else if (n1d == 5)
  {
    base[0] = -2.; base[1] = -1.; base[2] = 0.; base[3] = 1.; base[4] = 2.;
  }
  else
  {
    printf("CompInterpolatedPtInRefElement: bad number of interpolation points\n");
    exit(1);
  }
  id_pt = 1;
  for (E_Int k = 1; k <= n1d; k++)
    for (E_Int j = 1; j <= n1d; j++)
      for (E_Int i = 1; i <= n1d; i++)
      {
        BB(id_pt, 1) = base[i - 1];
        BB(id_pt, 2) = base[j - 1];
        BB(id_pt, 3) = base[k - 1];
        id_pt = id_pt + 1;
      }
  const E_Int three = 3;
  kgaussj_(A1, &n3d, B1, &three, &ok, indxc, indxr, ipiv);

  *xx = 0.; *yy = 0.; *zz = 0.;
  id_pt = 1;
  mk = 1;
  for (E_Int k = 1; k <= n1d; k++)
  {
    mj = 1;
    for (E_Int j = 1; j <= n1d; j++)
    {
      mi = 1;
      for (E_Int i = 1; i <= n1d; i++)
      {
        m = mi * mj * mk;
        *xx = *xx + BB(id_pt, 1) * m;
        *yy = *yy + BB(id_pt, 2) * m;
        *zz = *zz + BB(id_pt, 3) * m;
        id_pt = id_pt + 1;
        mi = mi * xp;
      }
      mj = mj * yp;
    }
    mk = mk * zp;
  }
  if (n1d == 2)
  {
    if (*xx < eps || *yy < eps || *zz < eps || *xx >= thres || *yy >= thres ||
        *zz >= thres)
      *erreur = 1;
  }
  else
  {
    if (fabs(*xx) >= thres || fabs(*yy) >= thres || fabs(*zz) >= thres)
      *erreur = 1;
  }
#undef AA
#undef BB
}

/* TETRA-donor volume (KCore/KCore/Metric, Fortran): not on the structured
 * hot path; only referenced by Interp.cpp so the symbol must exist. */
void k6compvoloftetracell_(const E_Int* npts, const E_Int* i1, const E_Int* i2,
                           const E_Int* i3, const E_Int* i4, const E_Float* xt,
                           const E_Float* yt, const E_Float* zt, E_Float* vol)
{
  (void)npts; (void)i1; (void)i2; (void)i3; (void)i4; (void)xt; (void)yt; (void)zt;
  (void)vol;
  fprintf(stderr, "oracle: TETRA donors are out of scope (k6compvoloftetracell)\n");
  abort();
}
