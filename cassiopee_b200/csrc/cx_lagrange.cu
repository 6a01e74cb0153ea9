// cx_lagrange.cu — order-3 / order-4 / order-5 Lagrange interpolation data on the device.
//
// Reference path (per receptor, per donor zone):
//   K_INTERP::getInterpolationData, O3ABC / O5ABC branches   KCore/KCore/Interp/Interp2.cpp:344-394, :452-503
//   CELLN(order) validity macro                              Interp2.cpp:40-67
//   K_INTERP::compLagrangeCoefs                              Interp2.cpp:517-665
//   compinterpolatedptinrefelt (27/125-point polynomial map) KCore/KCore/Interp/CompInterpolatedPtInRefElementF.for:25-300
//   cp_coord_in_stable_frame_3d / coord_in_ref_frame_3d      Cp_coord_in_stable_frame_3dF.for:25-174, Coord_in_ref_frame_3dF.for:23-228
//   kgaussj (full-pivot Gauss-Jordan, '>=' keeps the LAST max) KCore/KCore/Linear/GaussjF.for:29-100
//   cross-zone choice / type-1 collapse                      commonTypesForExtrapAndInterp.h, commonCheckCoefs.h
//
// Device pipeline per donor zone z (host loop over the ordered zone list):
//   k_lag_search   one thread per receptor: order-2 cell search (cx_core.cuh)
//   k_lag_solve    one thread GROUP per found receptor (a warp for 27x27, a
//                  256-thread CTA for 125x125): gather stencil, stable frame,
//                  monomial matrix in shared memory, Gauss-Jordan with the
//                  reference's pivot sequence, evaluate (ksi,eta,zeta)
//   k_lag_finish   one thread per receptor: Lagrange weights, cellN check,
//                  order-2 fallback, volume, running best over zones
// then k_lag_collapse applies the coincident (type 1) rule.
//
// Arithmetic: only the columns that are still un-pivoted and the 3 right-hand
// sides are updated during elimination.  Pivoted columns (the in-place inverse
// the Fortran routine also builds) are never read again by the pivot search
// (ipiv(k).EQ.0 filter) nor by the RHS update, so the pivot sequence and every
// value that reaches (ksi,eta,zeta) are bit-identical while the work halves.
#include "cx_internal.h"
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include <cub/device/device_radix_sort.cuh>

using namespace cx;

namespace {

struct LagState {
  int n;
  int* found;      // [n] 0/1 for the current zone
  int* ijk;        // [3][n] 1-based cell of the order-2 search
  double* tcf;     // [15][n] per-receptor tmpCf, persistent across zones (Interp.cpp:113)
  double* ksi;     // [3][n]
  int* err;        // [n]
  double* best;    // [n]
  int* noblk; int* type; int* dind;
  double* cf;      // [ncfmax][n]
  int ncfmax;
};

__global__ void k_lag_init(LagState S)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= S.n) return;
  S.best[r] = CX_MAXF; S.noblk[r] = 0; S.type[r] = 0; S.dind[r] = -1;
  for (int k = 0; k < 15; k++) S.tcf[(size_t)k * S.n + r] = 0.;
  for (int k = 0; k < S.ncfmax; k++) S.cf[(size_t)k * S.n + r] = 0.;
}

__global__ void __launch_bounds__(128)
k_lag_search(LagState S, const double* __restrict__ xr, const double* __restrict__ yr, const double* __restrict__ zr,
             const ZoneView* __restrict__ zones, int noz, int order, unsigned long long* __restrict__ flags, int cap,
             unsigned long long* gstats)
{
  // search scratch in dynamic shared memory (cx_core.cuh: Scratch), column = thread
  extern __shared__ __align__(16) double lag_sm[];
  Scratch SC;
  SC.vtx = lag_sm + threadIdx.x;
  SC.stk = reinterpret_cast<int*>(lag_sm + 24 * 128) + threadIdx.x;
  SC.ss = 128; SC.cap = cap;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= S.n) return;
  const ZoneView& Z = zones[noz];
  int ic = 0, jc = 0, kc = 0, found = 0;
  double cf[8];
  SearchStats st; st.nvisit = 0; st.ncand = 0; st.ntet = 0;
  // Interp2.cpp:347-351 / :455-459: zones too thin for the stencil return -1
  if (!(Z.ni < order || Z.nj < order || Z.nk < order))
  {
    if (Z.topo == 0)
    {
      // Cartesian donor (Interp2.cpp:384-393): order 3 has closed-form direction weights (no polynomial solve);
      // order 5 is not implemented by the reference for Cartesian InterpData (Interp2.cpp:464-467: returns -1)
      if (order == 3)
      {
        double c9[9];
        if (cart_search_o3(Z, xr[r], yr[r], zr[r], ic, jc, kc, c9) == 1)
        {
          found = 2;                  // final coefficients: skip the solve, k_lag_finish only tests and ranks
          for (int k = 0; k < 9; k++) S.tcf[(size_t)k * S.n + r] = c9[k];
        }
      }
      else if (order == 4)
      {
        double c12[12];
        if (cart_search_o4(Z, xr[r], yr[r], zr[r], ic, jc, kc, c12) == 1)
        {
          found = 2;
          for (int k = 0; k < 12; k++) S.tcf[(size_t)k * S.n + r] = c12[k];
        }
      }
    }
    else
    {
      found = search_interp_cell_s(Z, xr[r], yr[r], zr[r], ic, jc, kc, cf, SC, &st);
      if (found == 1)
        for (int k = 0; k < 8; k++) S.tcf[(size_t)k * S.n + r] = cf[k];
    }
  }
  S.found[r] = found;
  S.ijk[r] = ic; S.ijk[(size_t)S.n + r] = jc; S.ijk[2 * (size_t)S.n + r] = kc;
  flags[r] = (found == 1) ? 1ull : 0ull;
  if (gstats && st.nvisit) atomicAdd(&gstats[0], (unsigned long long)st.nvisit);
}

// 1-based start (ics,jcs,kcs) of the Lagrange stencil from the 1-based cell
// of the order-2 search (Interp2.cpp:359-366 for O3, :469-485 for O5)
__device__ __forceinline__ void stencil_start(int order, const ZoneView& Z, int ic, int jc, int kc, int& ics, int& jcs,
                                              int& kcs)
{
  if (order == 3)
  {
    ic -= 1; jc -= 1; kc -= 1;
    if (ic < 1) ic = 1;
    if (jc < 1) jc = 1;
    if (kc < 1) kc = 1;
  }
  else if (order == 4)
  {
    // Interp2.cpp:411-416: shift the donor by one, re-centre next to the max boundary
    if (ic > 1) ic -= 1;
    if (ic + 3 > Z.ni) ic -= ic + 3 - Z.ni;
    if (jc > 1) jc -= 1;
    if (jc + 3 > Z.nj) jc -= jc + 3 - Z.nj;
    if (kc > 1) kc -= 1;
    if (kc + 3 > Z.nk) kc -= kc + 3 - Z.nk;
  }
  else
  {
    if (ic >= Z.ni - 1) ic = Z.ni - 4; else ic = ic - 2;
    if (jc >= Z.nj - 1) jc = Z.nj - 4; else jc = jc - 2;
    if (kc >= Z.nk - 1) kc = Z.nk - 4; else kc = kc - 2;
    if (ic < 1) ic = 1;
    if (jc < 1) jc = 1;
    if (kc < 1) kc = 1;
  }
  ics = ic; jcs = jc; kcs = kc;
}

__device__ __forceinline__ double det3(const double* m)
{ // Coord_in_ref_frame_3dF.for:205-211, column-major m(i,j) = m[(i-1)+(j-1)*3]
#define MM(i, j) m[(i - 1) + (j - 1) * 3]
  return MM(1, 1) * MM(2, 2) * MM(3, 3) + MM(2, 1) * MM(3, 2) * MM(1, 3) + MM(3, 1) * MM(1, 2) * MM(2, 3) -
         MM(1, 1) * MM(3, 2) * MM(2, 3) - MM(2, 1) * MM(1, 2) * MM(3, 3) - MM(3, 1) * MM(2, 2) * MM(1, 3);
#undef MM
}

// coord_in_ref_frame_3d (Coord_in_ref_frame_3dF.for:23-127); pts(3,4) column-major
__device__ __forceinline__ void ref_frame(const double* pts, double x1, double y1, double z1, double& x2, double& y2,
                                          double& z2)
{
  double u[3], v[3], w[3], t[3], m[9];
  for (int i = 0; i < 3; i++) { u[i] = pts[i + 3] - pts[i]; v[i] = pts[i + 6] - pts[i]; w[i] = pts[i + 9] - pts[i]; }
  t[0] = x1 - pts[0]; t[1] = y1 - pts[1]; t[2] = z1 - pts[2];
  for (int i = 0; i < 3; i++) { m[i] = u[i]; m[i + 3] = v[i]; m[i + 6] = w[i]; }
  double invdet = det3(m);
  invdet = 1.0 / invdet;
  for (int i = 0; i < 3; i++) { m[i] = t[i]; m[i + 3] = v[i]; m[i + 6] = w[i]; }
  x2 = det3(m); x2 = x2 * invdet;
  for (int i = 0; i < 3; i++) { m[i] = u[i]; m[i + 3] = t[i]; m[i + 6] = w[i]; }
  y2 = det3(m); y2 = y2 * invdet;
  for (int i = 0; i < 3; i++) { m[i] = u[i]; m[i + 3] = v[i]; m[i + 6] = t[i]; }
  z2 = det3(m); z2 = z2 * invdet;
}

template <bool WARP>
__device__ __forceinline__ void gsync()
{
  if (WARP) __syncwarp(); else __syncthreads();
}

// One thread group (NT threads) solves the system of one STENCIL and evaluates it for the nr receptors rl[0..nr)
// that share it.  The frame and the matrix are built from the stencil nodes alone
// (CompInterpolatedPtInRefElementF.for:92-139, Cp_coord_in_stable_frame_3dF.for), the receptor only enters at the
// evaluation (:144-186, :242-263), so receptors with the same stencil get bit for bit what one solve each gives them.
//   N1 = points per direction (3 or 5), N = N1^3 rows, LD = padded row length
//   (N columns + 3 RHS, odd number of doubles => conflict-free row-strided
//   access).  Shared layout per group: A[N][LD], px/py/pz[N], small scratch.
template <int N1, int NT, bool WARP>
__device__ void lag_solve_group(const ZoneView& Z, int ics, int jcs, int kcs, const LagState& S, const int* __restrict__ rl,
                                int nr, const double* __restrict__ xr, const double* __restrict__ yr,
                                const double* __restrict__ zr, double* sm, int tid)
{
  constexpr int N = N1 * N1 * N1;
  constexpr int LD = (N + 3) | 1;
  double* A = sm;                 // N*LD
  double* px = A + N * LD;        // N   stencil coordinates (stable frame)
  double* py = px + N;
  double* pz = py + N;
  double* sc = pz + N;            // 16: pts(3,4) + transformed receptor
  constexpr int RED = WARP ? 0 : 64;         // cross-warp reduction scratch (none for one warp per system)
  volatile double* red_v = sc + 16;          // 2 x NT/32 partial maxima (double-buffered, up to 32 warps)
  volatile int* red_i = (volatile int*)(red_v + RED);  // 2 x NT/32 indices
  int* ipiv = (int*)(red_v + RED) + RED;     // N flags
  int* ctl = ipiv + N + 3;                   // [0]=irow [1]=icol [2]=singular

  const int ni = Z.ni, ninj = Z.ni * Z.nj;
  const int ind0 = ics - 1 + (jcs - 1) * ni + (kcs - 1) * ninj;
  // gather the stencil (CompInterpolatedPtInRefElementF.for:77-96)
  for (int p = tid; p < N; p += NT)
  {
    int i = p % N1, j = (p / N1) % N1, k = p / (N1 * N1);
    int ind = ind0 + i + j * ni + k * ninj;
    px[p] = Z.x[ind]; py[p] = Z.y[ind]; pz[p] = Z.z[ind];
    ipiv[p] = 0;
  }
  gsync<WARP>();
  // stable frame (Cp_coord_in_stable_frame_3dF.for:77-150): computed by one thread
  if (tid == 0)
  {
    const int c0 = (N1 + 1) / 2;   // 1-based centre index
#define IDP(i, j, k) ((i - 1) + (j - 1) * N1 + (k - 1) * N1 * N1)
    double ui[3] = {0., 0., 0.}, uj[3] = {0., 0., 0.}, uk[3] = {0., 0., 0.};
    int idc = IDP(c0, c0, c0);
    double O[3] = {px[idc], py[idc], pz[idc]};
    for (int dk = 0; dk <= 1; dk++) for (int dj = 0; dj <= 1; dj++)
    {
      int a = IDP(c0, c0 + dj, c0 + dk), b = IDP(c0 + 1, c0 + dj, c0 + dk);
      ui[0] = ui[0] + px[b] - px[a]; ui[1] = ui[1] + py[b] - py[a]; ui[2] = ui[2] + pz[b] - pz[a];
    }
    for (int dk = 0; dk <= 1; dk++) for (int di = 0; di <= 1; di++)
    {
      int a = IDP(c0 + di, c0, c0 + dk), b = IDP(c0 + di, c0 + 1, c0 + dk);
      uj[0] = uj[0] + px[b] - px[a]; uj[1] = uj[1] + py[b] - py[a]; uj[2] = uj[2] + pz[b] - pz[a];
    }
    for (int dj = 0; dj <= 1; dj++) for (int di = 0; di <= 1; di++)
    {
      int a = IDP(c0 + di, c0 + dj, c0), b = IDP(c0 + di, c0 + dj, c0 + 1);
      uk[0] = uk[0] + px[b] - px[a]; uk[1] = uk[1] + py[b] - py[a]; uk[2] = uk[2] + pz[b] - pz[a];
    }
#undef IDP
    for (int i = 0; i < 3; i++)
    {
      ui[i] = ui[i] * 0.25; uj[i] = uj[i] * 0.25; uk[i] = uk[i] * 0.25;
      sc[i] = O[i]; sc[i + 3] = O[i] + ui[i]; sc[i + 6] = O[i] + uj[i]; sc[i + 9] = O[i] + uk[i];
    }
    ctl[2] = 0;
  }
  gsync<WARP>();
  for (int p = tid; p < N; p += NT)
  {
    double a, b, c;
    ref_frame(sc, px[p], py[p], pz[p], a, b, c);
    // monomial row (CompInterpolatedPtInRefElementF.for:121-139) and RHS (:167-180)
    double* row = A + p * LD;
    double zp1 = 1;
    int col = 0;
    for (int k = 0; k < N1; k++)
    {
      double yp1 = 1;
      for (int j = 0; j < N1; j++)
      {
        double xp1 = 1;
        for (int i = 0; i < N1; i++)
        {
          row[col] = xp1 * yp1 * zp1;
          col++;
          xp1 = xp1 * a;
        }
        yp1 = yp1 * b;
      }
      zp1 = zp1 * c;
    }
    const double b0 = (N1 == 5) ? -2. : -1.;     // base(1): -1 for 3 and 4 points, -2 for 5 (CompInterpolatedPtInRefElementF.for:144-163)
    int i = p % N1, j = (p / N1) % N1, k = p / (N1 * N1);
    row[N] = b0 + i; row[N + 1] = b0 + j; row[N + 2] = b0 + k;
  }
  gsync<WARP>();

  // --- kgaussj (GaussjF.for:29-100) -----------------------------------------
  // thread -> (row, column segment): SEG segments per row.  The full-pivot search of step s+1 is fused
  // into the elimination of step s: while a thread updates its part of a row it tracks that part's
  // maximum |a| over the still un-pivoted columns (ascending columns, '>=' => last maximum, exactly the
  // scan order of GaussjF.for:34-44), so each search is only a cross-thread arg-max.
  constexpr int SEG = (NT >= 8 * N) ? 8 : (NT >= 4 * N) ? 4 : (NT >= 2 * N) ? 2 : 1;
  constexpr int ROWS = NT / SEG;
  const int row = tid % ROWS;
  const int seg = tid / ROWS;
  constexpr int CHUNK = (N + SEG - 1) / SEG;
  const int c0 = seg * CHUNK;
  const int c1 = (c0 + CHUNK < N) ? c0 + CHUNK : N;
  const bool rowOn = row < N && seg < SEG;
  static_assert(CHUNK <= 32, "live-column mask is one 32-bit word per thread");
  unsigned live = 0u;          // bit i set <=> column c0+i of this thread's segment is still un-pivoted
  if (rowOn) live = (c1 - c0 >= 32) ? 0xffffffffu : ((1u << (c1 - c0)) - 1u);
  double tmax = -1.; int tidx = -1;
  if (rowOn)
  {
    const double* ar = A + row * LD;
    for (int k = c0; k < c1; k++)
    {
      double v = fabs(ar[k]);
      if (v >= tmax) { tmax = v; tidx = row * N + k; }
    }
  }
  for (int step = 0; step < N; step++)
  {
    // rows whose index is an already pivoted column are excluded (ipiv(j).NE.1)
    double big = -1.; int bidx = -1;
    if (rowOn && ipiv[row] != 1) { big = tmax; bidx = tidx; }
    // arg-max: larger value, then larger (row,col) index == the LAST maximum of the reference's scan
#define CX_ARGMAX_STEP(o)                                                     \
    {                                                                         \
      double ov = __shfl_down_sync(0xffffffffu, big, o);                      \
      int oi = __shfl_down_sync(0xffffffffu, bidx, o);                        \
      if (ov > big || (ov == big && oi > bidx)) { big = ov; bidx = oi; }      \
    }
    CX_ARGMAX_STEP(16) CX_ARGMAX_STEP(8) CX_ARGMAX_STEP(4) CX_ARGMAX_STEP(2) CX_ARGMAX_STEP(1)
    if (WARP)
    {
      bidx = __shfl_sync(0xffffffffu, bidx, 0);
    }
    else
    {
      // one barrier: every warp publishes its winner, then every warp reduces the NT/32 winners itself
      const int par = step & 1;       // double-buffered so the next step's writes cannot race with late readers
      if ((tid & 31) == 0) { red_v[par * 32 + (tid >> 5)] = big; red_i[par * 32 + (tid >> 5)] = bidx; }
      __syncthreads();
      int ln = tid & 31;
      big = (ln < NT / 32) ? red_v[par * 32 + ln] : -2.;
      bidx = (ln < NT / 32) ? red_i[par * 32 + ln] : -1;
      CX_ARGMAX_STEP(16) CX_ARGMAX_STEP(8) CX_ARGMAX_STEP(4) CX_ARGMAX_STEP(2) CX_ARGMAX_STEP(1)
      bidx = __shfl_sync(0xffffffffu, bidx, 0);
    }
#undef CX_ARGMAX_STEP
    // bidx < 0 can only happen with NaNs everywhere; treat like the singular exit
    if (bidx < 0) { if (tid == 0) ctl[2] = 1; gsync<WARP>(); break; }
    const int irow = bidx / N, icol = bidx % N;
    const double piv = A[irow * LD + icol];      // a(icol,icol) after the row interchange
    if (icol >= c0 && icol < c1) live &= ~(1u << (icol - c0));
    gsync<WARP>();                               // everybody has read the pivot before rows change
    if (piv == 0.) { if (tid == 0) ctl[2] = 1; gsync<WARP>(); break; }   // 'singular matrix in gaussj, 2'
    const double pivinv = 1. / piv;
    if (tid == 0) ipiv[icol] = ipiv[icol] + 1;
    // row interchange irow <-> icol and scaling of the pivot row in one pass (GaussjF.for:51-77)
    for (int l = tid; l < N + 3; l += NT)
    {
      double* a = A + irow * LD + l; double* b = A + icol * LD + l;
      double x = *a, y = *b;
      *b = (l == icol) ? 1. * pivinv : x * pivinv;
      if (irow != icol) *a = y;
    }
    gsync<WARP>();
    // eliminate column icol from every other row (GaussjF.for:78-89); only live columns and the 3 RHS
    // are updated; the maxima for the next search are tracked on the fly
    tmax = -1.; tidx = -1;
    if (rowOn && row != icol)
    {
      double* ar = A + row * LD;
      const double* ap = A + icol * LD;
      const double dum = ar[icol];
      for (unsigned mk = live; mk; mk &= mk - 1)
      {
        int l = c0 + __ffs(mk) - 1;
        double nv = ar[l] - ap[l] * dum;
        ar[l] = nv;
        double v = fabs(nv);
        if (v >= tmax) { tmax = v; tidx = row * N + l; }
      }
      if (seg == SEG - 1)
      {
        ar[N] = ar[N] - ap[N] * dum; ar[N + 1] = ar[N + 1] - ap[N + 1] * dum; ar[N + 2] = ar[N + 2] - ap[N + 2] * dum;
      }
    }
    gsync<WARP>();
    // a(ll,icol) = 0 for ll != icol happens implicitly: column icol is never read again
  }
  gsync<WARP>();
  // evaluate the polynomial map at every receptor of the stencil (CompInterpolatedPtInRefElementF.for:242-263):
  // three lanes per receptor (ksi, eta, zeta), ten receptors per warp and pass
  const double thres = (double)(N1 / 2) + 0.001;
  const int lane = tid & 31, wq = lane / 3, comp = lane - 3 * wq;
  constexpr int PER_PASS = 10 * (NT / 32);
  for (int base = 0; base < nr; base += PER_PASS)
  {
    const int q = base + (tid >> 5) * 10 + wq;
    const bool on = lane < 30 && q < nr;
    double acc = 0.;
    int r = 0;
    if (on)
    {
      r = rl[q];
      double xp, yp, zp;
      ref_frame(sc, xr[r], yr[r], zr[r], xp, yp, zp);
      int id = 0;
      double mk = 1;
      for (int k = 0; k < N1; k++)
      {
        double mj = 1;
        for (int j = 0; j < N1; j++)
        {
          double mi = 1;
          for (int i = 0; i < N1; i++)
          {
            double m = mi * mj * mk;
            acc = acc + A[id * LD + N + comp] * m;
            id++;
            mi = mi * xp;
          }
          mj = mj * yp;
        }
        mk = mk * zp;
      }
      S.ksi[(size_t)comp * S.n + r] = acc;
    }
    const unsigned bad = __ballot_sync(0xffffffffu, on && fabs(acc) >= thres);
    if (on && comp == 0) S.err[r] = ((bad >> (3 * wq)) & 7u) ? 1 : 0;
  }
  gsync<WARP>();
}

template <int N1> struct LagCfg {
  static constexpr int N = N1 * N1 * N1;
  static constexpr int LD = (N + 3) | 1;
  // doubles of shared memory per group
  static constexpr int RED = (N1 == 3) ? 0 : 64;      // order 3 runs one warp per system
  static constexpr int SMEM_D = N * LD + 3 * N + 16 + RED + ((RED + N + 3 + 4) * 4 + 7) / 8;
};

// Distinct stencils of the found receptors: key = linear index of the stencil's first node
__global__ void k_lag_keys(LagState S, int nlist, const int* __restrict__ list, const ZoneView* __restrict__ zones, int noz,
                           int order, unsigned* __restrict__ key)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nlist) return;
  const int r = list[a];
  const ZoneView& Z = zones[noz];
  int ics, jcs, kcs;
  stencil_start(order, Z, S.ijk[r], S.ijk[(size_t)S.n + r], S.ijk[2 * (size_t)S.n + r], ics, jcs, kcs);
  key[a] = (unsigned)(ics - 1 + (jcs - 1) * Z.ni + (kcs - 1) * Z.ni * Z.nj);
}
__global__ void k_lag_heads(int nlist, const unsigned* __restrict__ key, unsigned long long* __restrict__ f)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < nlist) f[a] = (a == 0 || key[a] != key[a - 1]) ? 1ull : 0ull;
}
__global__ void k_lag_ustart(int nlist, const unsigned long long* __restrict__ f, const unsigned long long* __restrict__ pos,
                             int* __restrict__ ustart, int nu)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a == 0) ustart[nu] = nlist;
  if (a < nlist && f[a]) ustart[(int)pos[a]] = a;
}

// Stencil u of the sorted list: receptors rl[ustart[u] .. ustart[u+1])
__device__ __forceinline__ void stencil_of(const LagState& S, const ZoneView& Z, int order, int r0, int& ics, int& jcs, int& kcs)
{
  stencil_start(order, Z, S.ijk[r0], S.ijk[(size_t)S.n + r0], S.ijk[2 * (size_t)S.n + r0], ics, jcs, kcs);
}

// order 3: one warp per stencil, 4 stencils per CTA
__global__ void __launch_bounds__(128)
k_lag_solve3(LagState S, int nu, const int* __restrict__ ustart, const int* __restrict__ rl, const double* __restrict__ xr,
             const double* __restrict__ yr, const double* __restrict__ zr, const ZoneView* __restrict__ zones, int noz)
{
  extern __shared__ double smem[];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u = blockIdx.x * 4 + w;
  if (u >= nu) return;
  const int a0 = ustart[u], a1 = ustart[u + 1];
  const ZoneView& Z = zones[noz];
  int ics, jcs, kcs;
  stencil_of(S, Z, 3, rl[a0], ics, jcs, kcs);
  lag_solve_group<3, 32, true>(Z, ics, jcs, kcs, S, rl + a0, a1 - a0, xr, yr, zr, smem + (size_t)w * LagCfg<3>::SMEM_D, lane);
}

// order 4: one 256-thread CTA per stencil, matrix (64 x 67 doubles) in shared memory
__global__ void __launch_bounds__(256)
k_lag_solve4(LagState S, int nu, const int* __restrict__ ustart, const int* __restrict__ rl, const double* __restrict__ xr,
             const double* __restrict__ yr, const double* __restrict__ zr, const ZoneView* __restrict__ zones, int noz)
{
  extern __shared__ double smem[];
  for (int u = blockIdx.x; u < nu; u += gridDim.x)
  {
    const int a0 = ustart[u], a1 = ustart[u + 1];
    const ZoneView& Z = zones[noz];
    int ics, jcs, kcs;
    stencil_of(S, Z, 4, rl[a0], ics, jcs, kcs);
    lag_solve_group<4, 256, false>(Z, ics, jcs, kcs, S, rl + a0, a1 - a0, xr, yr, zr, smem, threadIdx.x);
    __syncthreads();
  }
}

// order 5: one 512-thread CTA per stencil, matrix (125 x 129 doubles) in shared memory (CX_LAG5=512)
__global__ void __launch_bounds__(512)
k_lag_solve5(LagState S, int nu, const int* __restrict__ ustart, const int* __restrict__ rl, const double* __restrict__ xr,
             const double* __restrict__ yr, const double* __restrict__ zr, const ZoneView* __restrict__ zones, int noz)
{
  extern __shared__ double smem[];
  for (int u = blockIdx.x; u < nu; u += gridDim.x)
  {
    const int a0 = ustart[u], a1 = ustart[u + 1];
    const ZoneView& Z = zones[noz];
    int ics, jcs, kcs;
    stencil_of(S, Z, 5, rl[a0], ics, jcs, kcs);
    lag_solve_group<5, 512, false>(Z, ics, jcs, kcs, S, rl + a0, a1 - a0, xr, yr, zr, smem, threadIdx.x);
    __syncthreads();
  }
}

// order 5 with a 1024-thread CTA per stencil (8 column segments of 16 per row, 64 registers): twice the warps
// per SM for the pivot step's dependent chain; the default (1.90 s vs 1.99 s per 1.06 M receptors; A/B:
// CX_LAG5=512).  Keeping the matrix in registers instead (three variants, profiles/r01_e_lagrange5.md) did
// not help: the pivot step is bound by instruction issue and the two barriers, not by shared memory.
__global__ void __launch_bounds__(1024, 1)
k_lag_solve5w(LagState S, int nu, const int* __restrict__ ustart, const int* __restrict__ rl, const double* __restrict__ xr,
              const double* __restrict__ yr, const double* __restrict__ zr, const ZoneView* __restrict__ zones, int noz)
{
  extern __shared__ double smem[];
  for (int u = blockIdx.x; u < nu; u += gridDim.x)
  {
    const int a0 = ustart[u], a1 = ustart[u + 1];
    const ZoneView& Z = zones[noz];
    int ics, jcs, kcs;
    stencil_of(S, Z, 5, rl[a0], ics, jcs, kcs);
    lag_solve_group<5, 1024, false>(Z, ics, jcs, kcs, S, rl + a0, a1 - a0, xr, yr, zr, smem, threadIdx.x);
    __syncthreads();
  }
}

// CELLN(order) (Interp2.cpp:40-67) on the per-receptor tmpCf; (ic,jc,kc) 0-based stencil start
__device__ __forceinline__ bool celln_valid_lag(const ZoneView& Z, int O, int ic, int jc, int kc, const double* cf,
                                                int nature)
{
  if (Z.cellN == nullptr) return true;
  const int ni = Z.ni, nij = Z.ni * Z.nj;
  if (nature == 0)
  {
    double val = 1.;
    for (int kk = 0; kk < O; kk++) for (int jj = 0; jj < O; jj++) for (int ii = 0; ii < O; ii++)
    {
      double c0 = Z.cellN[(ic + ii) + (jc + jj) * ni + (kc + kk) * nij];
      int d = cf[ii] * cf[jj + O] * cf[kk + O * 2] > 1.e-12;
      val *= c0 - d + 1;
    }
    return !eqzero(val, CX_EPS_GEOM);
  }
  double val = 0.;
  for (int kk = 0; kk < O; kk++) for (int jj = 0; jj < O; jj++) for (int ii = 0; ii < O; ii++)
  {
    double c0 = Z.cellN[(ic + ii) + (jc + jj) * ni + (kc + kk) * nij];
    val += fabs(cf[ii] * cf[jj + O] * cf[kk + O * 2]) * fabs(1. - c0);
  }
  return eqzero(val, CX_EPS_GEOM);
}

// Lagrange weights with the reference's literal constants (Interp2.cpp:580-658)
__device__ __forceinline__ void lagrange_weights(int order, double ksi, double eta, double zeta, double* cf)
{
  double ksip = 1. + ksi, ksim = 1. - ksi, etap = 1. + eta, etam = 1. - eta, zetap = 1. + zeta, zetam = 1. - zeta;
  if (order == 3)
  {
    cf[0] = -0.5 * ksi * ksim; cf[1] = ksip * ksim; cf[2] = 0.5 * ksi * ksip;
    cf[3] = -0.5 * eta * etam; cf[4] = etap * etam; cf[5] = 0.5 * eta * etap;
    cf[6] = -0.5 * zeta * zetam; cf[7] = zetap * zetam; cf[8] = 0.5 * zeta * zetap;
  }
  else if (order == 4)
  {
    // Interp2.cpp O4ABC branch (literal inv6, ONE_HALF)
    const double inv6 = 0.166666666667;
    double ksim2 = 2. - ksi, etam2 = 2. - eta, zetam2 = 2. - zeta;
    cf[0] = -inv6 * ksim2 * ksim * ksi;
    cf[1] = 0.5 * ksim2 * ksim * ksip;
    cf[2] = 0.5 * ksim2 * ksi * ksip;
    cf[3] = -inv6 * ksim * ksi * ksip;
    cf[4] = -inv6 * etam2 * etam * eta;
    cf[5] = 0.5 * etam2 * etam * etap;
    cf[6] = 0.5 * etam2 * eta * etap;
    cf[7] = -inv6 * etam * eta * etap;
    cf[8] = -inv6 * zetam2 * zetam * zeta;
    cf[9] = 0.5 * zetam2 * zetam * zetap;
    cf[10] = 0.5 * zetam2 * zeta * zetap;
    cf[11] = -inv6 * zetam * zeta * zetap;
  }
  else
  {
    const double inv24 = 0.041666666667, inv6 = 0.166666666667, inv4 = 0.25;
    double ksip2 = 2. + ksi, ksim2 = 2. - ksi, etap2 = 2. + eta, etam2 = 2. - eta, zetap2 = 2. + zeta, zetam2 = 2. - zeta;
    cf[0] = inv24 * ksi * ksip * ksim * ksim2;
    cf[1] = -inv6 * ksi * ksip2 * ksim * ksim2;
    cf[2] = inv4 * ksim2 * ksim * ksip * ksip2;
    cf[3] = inv6 * ksi * ksim2 * ksip * ksip2;
    cf[4] = -inv24 * ksip2 * ksip * ksi * ksim;
    cf[5] = inv24 * eta * etap * etam * etam2;
    cf[6] = -inv6 * eta * etap2 * etam * etam2;
    cf[7] = inv4 * etam2 * etam * etap * etap2;
    cf[8] = inv6 * eta * etam2 * etap * etap2;
    cf[9] = -inv24 * etap2 * etap * eta * etam;
    cf[10] = inv24 * zeta * zetap * zetam * zetam2;
    cf[11] = -inv6 * zeta * zetap2 * zetam * zetam2;
    cf[12] = inv4 * zetam2 * zetam * zetap * zetap2;
    cf[13] = inv6 * zeta * zetam2 * zetap * zetap2;
    cf[14] = -inv24 * zetap2 * zetap * zeta * zetam;
  }
}

__global__ void __launch_bounds__(128)
k_lag_finish(LagState S, const ZoneView* __restrict__ zones, int noz, int order, int nature, int penalty)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= S.n) return;
  if (S.found[r] == 2)
  {
    // Cartesian donor, order 3 or 4 (Interp2.cpp:386-392, :441-447): stencil start and weights come from the search
    const ZoneView& Z = zones[noz];
    const size_t n = S.n;
    const int O = order, nc = 3 * order;
    const int ic = S.ijk[r] - 1, jc = S.ijk[n + r] - 1, kc = S.ijk[2 * n + r] - 1;
    const int tind = ic + jc * Z.ni + kc * Z.ni * Z.nj;
    const int isBorder = (ic == 0 || ic == Z.ni - O || jc == 0 || jc == Z.nj - O || kc == 0 || kc == Z.nk - O);
    double cf[12];
    for (int k = 0; k < nc; k++) cf[k] = S.tcf[(size_t)k * n + r];
    if (!celln_valid_lag(Z, O, ic, jc, kc, cf, nature)) return;
    double vol = cell_volume(Z, tind);
    if (penalty == 1 && isBorder == 1) vol += 1.e3;
    if (vol < S.best[r] + CX_EPS_GEOM)
    {
      S.best[r] = vol; S.type[r] = (O == 4) ? 44 : 3; S.dind[r] = tind; S.noblk[r] = noz + 1;
      for (int k = 0; k < S.ncfmax; k++) S.cf[(size_t)k * n + r] = k < nc ? cf[k] : 0.;
    }
    return;
  }
  if (S.found[r] != 1) return;
  const ZoneView& Z = zones[noz];
  const size_t n = S.n;
  const int O = order, ncf = 3 * order;
  int ic1 = S.ijk[r], jc1 = S.ijk[n + r], kc1 = S.ijk[2 * n + r];   // 1-based order-2 cell
  int ics, jcs, kcs;
  stencil_start(order, Z, ic1, jc1, kc1, ics, jcs, kcs);
  int ic = ics - 1, jc = jcs - 1, kc = kcs - 1;
  int isBorder = (ic == 0 || ic == Z.ni - O || jc == 0 || jc == Z.nj - O || kc == 0 || kc == Z.nk - O);
  int type = (order == 4) ? 44 : order;
  int tind = ic + jc * Z.ni + kc * Z.ni * Z.nj;
  double cf[15];
  for (int k = 0; k < 15; k++) cf[k] = S.tcf[(size_t)k * n + r];
  const int corr = S.err[r];
  if (order == 3 || order == 4)
  {
    // O3 / O4: weights first (when well approximated), then the cellN test (Interp2.cpp:369-372, :418-425)
    if (corr == 0) lagrange_weights(order, S.ksi[r], S.ksi[n + r], S.ksi[2 * n + r], cf);
    if (!celln_valid_lag(Z, order, ic, jc, kc, cf, nature))
    {
      for (int k = 0; k < ncf; k++) S.tcf[(size_t)k * n + r] = cf[k];
      return;
    }
  }
  else
  {
    // O5: cellN test BEFORE the weights, on whatever tmpCf holds (Interp2.cpp:490-492)
    if (!celln_valid_lag(Z, 5, ic, jc, kc, cf, nature)) return;
    if (corr == 0) lagrange_weights(5, S.ksi[r], S.ksi[n + r], S.ksi[2 * n + r], cf);
  }
  if (corr == 1)
  {
    // badly approximated point: order-2 result of the same search (Interp2.cpp:374-382, :493-502)
    ic = ic1 - 1; jc = jc1 - 1; kc = kc1 - 1;
    isBorder = (ic == 0 || ic == Z.ni - 2 || jc == 0 || jc == Z.nj - 2 || kc == 0 || kc == Z.nk - 2);
    type = 2; tind = ic + jc * Z.ni + kc * Z.ni * Z.nj;
  }
  for (int k = 0; k < ncf; k++) S.tcf[(size_t)k * n + r] = cf[k];
  // commonTypesForExtrapAndInterp.h:7-36
  double vol = cell_volume(Z, tind);
  if (penalty == 1 && isBorder == 1) vol += 1.e3;
  if (vol < S.best[r] + CX_EPS_GEOM)
  {
    S.best[r] = vol; S.type[r] = type; S.dind[r] = tind; S.noblk[r] = noz + 1;
    for (int k = 0; k < S.ncfmax; k++) S.cf[(size_t)k * n + r] = cf[k];
  }
}

// commonCheckCoefs.h: exactly one stencil weight above 1e-10 -> type 1
__global__ void k_lag_collapse(LagState S, const ZoneView* __restrict__ zones)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= S.n) return;
  int nb = S.noblk[r];
  if (nb <= 0) return;
  const ZoneView& Z = zones[nb - 1];
  const size_t n = S.n;
  int type = S.type[r], dind = S.dind[r];
  const int ni = Z.ni, ninj = Z.ni * Z.nj;
  int cnt = 0, indNonZero = -1; double coefNonZero = 0.;
  if (type == 2)
  {
    for (int kk = 0; kk < 2; kk++) for (int jj = 0; jj < 2; jj++) for (int ii = 0; ii < 2; ii++)
    {
      double c = S.cf[(size_t)(ii + jj * 2 + kk * 4) * n + r];
      if (fabs(c) > 1.e-10) { cnt++; indNonZero = dind + ii + jj * ni + kk * ninj; coefNonZero = c; }
    }
  }
  else if (type == 3 || type == 5)      // no case 44 in commonCheckCoefs.h: order-4 stencils are never collapsed
  {
    const int O = type;
    for (int kk = 0; kk < O; kk++) for (int jj = 0; jj < O; jj++) for (int ii = 0; ii < O; ii++)
    {
      double c = S.cf[(size_t)ii * n + r] * S.cf[(size_t)(jj + O) * n + r] * S.cf[(size_t)(kk + 2 * O) * n + r];
      if (fabs(c) > 1.e-10) { cnt++; indNonZero = dind + ii + jj * ni + kk * ninj; coefNonZero = c; }
    }
  }
  if (cnt == 1) { S.type[r] = 1; S.dind[r] = indNonZero; S.cf[r] = coefNonZero; }
}

__global__ void k_iota_lag(int n, int* __restrict__ a)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) a[e] = e;
}

__global__ void k_list(int n, const unsigned long long* __restrict__ f, const unsigned long long* __restrict__ pos,
                       int* __restrict__ out)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if (f[r]) out[(int)pos[r]] = r;
}

}  // namespace

int cx_lagrange_pipeline(cx_ctx* ctx, int order, int n, const double* d_xr0, const double* d_yr0, const double* d_zr0,
                         int nz, const ZoneView* d_zones, const ZoneView* h_zones, int nature, int penalty, int* noblk,
                         int* type, int* dind, double* cf, int ncfmax, unsigned long long* d_stats, const RcvPerZone* h_pz)
{
  // h_pz (host copy of the double-wall table, or nullptr): donor zone noz sees the receptors at h_pz[noz]; the zones go
  // through one after the other, so each zone's launches simply get that zone's coordinate arrays
  cudaStream_t s = ctx->stream;
  LagState S;
  size_t N = n;
  S.n = n; S.noblk = noblk; S.type = type; S.dind = dind; S.cf = cf; S.ncfmax = ncfmax;
  CX_MALLOC(S.found, sizeof(int) * N, s);
  CX_MALLOC(S.ijk, sizeof(int) * 3 * N, s);
  CX_MALLOC(S.tcf, sizeof(double) * 15 * N, s);
  CX_MALLOC(S.ksi, sizeof(double) * 3 * N, s);
  CX_MALLOC(S.err, sizeof(int) * N, s);
  CX_MALLOC(S.best, sizeof(double) * N, s);
  unsigned long long *d_f, *d_p; int *d_list, *d_slist, *d_ustart; unsigned *d_key, *d_skey;
  CX_MALLOC(d_f, sizeof(unsigned long long) * N, s);
  CX_MALLOC(d_p, sizeof(unsigned long long) * N, s);
  CX_MALLOC(d_list, sizeof(int) * N, s);
  CX_MALLOC(d_slist, sizeof(int) * N, s);
  CX_MALLOC(d_ustart, sizeof(int) * (N + 1), s);
  CX_MALLOC(d_key, sizeof(unsigned) * N, s);
  CX_MALLOC(d_skey, sizeof(unsigned) * N, s);
  void* d_tmp = nullptr; size_t tmp_cap = 0;
  // CX_LAG_DEDUP=0: one solve per receptor, like the reference (A/B)
  const char* dd_env = getenv("CX_LAG_DEDUP");
  const bool dedup = !(dd_env && dd_env[0] == '0');
  const int TB = 128;
  long long lag_solves = 0, lag_found = 0;
  k_lag_init<<<cx_grid(n, 256), 256, 0, s>>>(S);
  ctx->launches++;
  const size_t smem3 = sizeof(double) * LagCfg<3>::SMEM_D * 4;
  const size_t smem4 = sizeof(double) * LagCfg<4>::SMEM_D;
  const size_t smem5 = sizeof(double) * LagCfg<5>::SMEM_D;
  const char* lag5_env = getenv("CX_LAG5");
  const bool lag5_wide = !(lag5_env && strcmp(lag5_env, "512") == 0);   // A/B: CX_LAG5=512 = one 512-thread CTA
  if (order == 5)
  {
    CX_CUDA(cudaFuncSetAttribute(k_lag_solve5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));
    CX_CUDA(cudaFuncSetAttribute(k_lag_solve5w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));
  }
  for (int noz = 0; noz < nz; noz++)
  {
    const double* d_xr = h_pz ? h_pz[noz].x : d_xr0; const double* d_yr = h_pz ? h_pz[noz].y : d_yr0;
    const double* d_zr = h_pz ? h_pz[noz].z : d_zr0;
    const int cap = std::max(4, h_zones[noz].depth + 2);
    const size_t smsearch = (size_t)TB * (24 * sizeof(double) + (size_t)cap * sizeof(int));
    if (smsearch > 200 * 1024) { cx_set_error("setInterpData: ADT depth %d needs %zu bytes of shared memory per CTA", cap - 2, smsearch); return -1; }
    CX_CUDA(cudaFuncSetAttribute(k_lag_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smsearch));
    k_lag_search<<<cx_grid(n, TB), TB, smsearch, s>>>(S, d_xr, d_yr, d_zr, d_zones, noz, order, d_f, cap, d_stats);
    ctx->launches++;
    unsigned long long tot = 0;
    CX_CHECK(cx_exclusive_scan_u64(ctx, d_f, d_p, n, &tot));
    int nlist = (int)tot;
    if (nlist == 0 && h_zones[noz].topo != 0) continue;
    if (nlist > 0)
    {
      k_list<<<cx_grid(n, 256), 256, 0, s>>>(n, d_f, d_p, d_list);
      // one system per DISTINCT stencil: receptors sorted by stencil (stable radix sort), run starts = stencils
      int nu = nlist;
      const int* rl = d_list;
      if (dedup)
      {
        k_lag_keys<<<cx_grid(nlist, 256), 256, 0, s>>>(S, nlist, d_list, d_zones, noz, order, d_key);
        const unsigned kmax = (unsigned)h_zones[noz].ni * (unsigned)h_zones[noz].nj * (unsigned)h_zones[noz].nk;
        int bits = 1; while (bits < 32 && (1ull << bits) <= (unsigned long long)kmax) bits++;
        size_t need = 0;
        CX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, d_key, d_skey, d_list, d_slist, nlist, 0, bits, s));
        if (need > tmp_cap) { if (d_tmp) cudaFreeAsync(d_tmp, s); CX_MALLOC(d_tmp, need, s); tmp_cap = need; }
        CX_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_key, d_skey, d_list, d_slist, nlist, 0, bits, s));
        k_lag_heads<<<cx_grid(nlist, 256), 256, 0, s>>>(nlist, d_skey, d_f);
        unsigned long long tu = 0;
        CX_CHECK(cx_exclusive_scan_u64(ctx, d_f, d_p, nlist, &tu));
        nu = (int)tu;
        k_lag_ustart<<<cx_grid(nlist, 256), 256, 0, s>>>(nlist, d_f, d_p, d_ustart, nu);
        rl = d_slist;
        ctx->launches += 4;
      }
      else
        k_iota_lag<<<cx_grid(nlist + 1, 256), 256, 0, s>>>(nlist + 1, d_ustart);
      if (order == 3)
        k_lag_solve3<<<cx_grid(nu, 4), 128, smem3, s>>>(S, nu, d_ustart, rl, d_xr, d_yr, d_zr, d_zones, noz);
      else if (order == 4)
      {
        int grid = nu < ctx->sm_count * 24 ? nu : ctx->sm_count * 24;
        k_lag_solve4<<<grid, 256, smem4, s>>>(S, nu, d_ustart, rl, d_xr, d_yr, d_zr, d_zones, noz);
      }
      else
      {
        int grid = nu < ctx->sm_count * 4 ? nu : ctx->sm_count * 4;
        if (lag5_wide) k_lag_solve5w<<<grid, 1024, smem5, s>>>(S, nu, d_ustart, rl, d_xr, d_yr, d_zr, d_zones, noz);
        else k_lag_solve5<<<grid, 512, smem5, s>>>(S, nu, d_ustart, rl, d_xr, d_yr, d_zr, d_zones, noz);
      }
      lag_solves += nu; lag_found += nlist;
    }
    k_lag_finish<<<cx_grid(n, TB), TB, 0, s>>>(S, d_zones, noz, order, nature, penalty);
    ctx->launches += 3;
  }
  k_lag_collapse<<<cx_grid(n, 256), 256, 0, s>>>(S, d_zones);
  ctx->launches++;
  CX_CUDA(cudaStreamSynchronize(s));
  CX_CUDA(cudaGetLastError());
  cudaFreeAsync(S.found, s); cudaFreeAsync(S.ijk, s); cudaFreeAsync(S.tcf, s); cudaFreeAsync(S.ksi, s); cudaFreeAsync(S.err, s); cudaFreeAsync(S.best, s);
  cudaFreeAsync(d_f, s); cudaFreeAsync(d_p, s); cudaFreeAsync(d_list, s);
  cudaFreeAsync(d_slist, s); cudaFreeAsync(d_ustart, s); cudaFreeAsync(d_key, s); cudaFreeAsync(d_skey, s);
  if (d_tmp) cudaFreeAsync(d_tmp, s);
  if (getenv("CX_PROFILE")) fprintf(stderr, "[cx] lagrange order %d: %lld found receptors, %lld systems solved\n", order, lag_found, lag_solves);
  return 0;
}
