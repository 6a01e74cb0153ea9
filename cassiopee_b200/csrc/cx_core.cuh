// cx_core.cuh — per-receptor chimera search / coefficient logic (fp64, no FMA).
//
// Everything here runs per thread in registers/local memory on sm_100a
// (CX_D = __device__).  tests/hostemu compiles the same header with g++
// (CX_D empty) purely to unit-test the logic on the GPU-less builder box; the
// product never runs it on the CPU.
//
// What the reference computes, re-stated for one CUDA thread per receptor:
//   ADT range query, DFS order node/right/left   KCore/KCore/Interp/InterpAdt.cpp:658-929
//   jump walk + ordered fallback                 KCore/KCore/Interp/InterpAdt.cpp:1252-1400
//   hexa -> 15 points, 24 centre tetrahedra      KCore/KCore/Interp/InterpData.cpp:56-134, :141-379, :383-425, :496-685
//   3x3 cofactor tetra solve                     KCore/KCore/Interp/InterpData.cpp:429-485
//   extrapolation over all 24 tetrahedra         KCore/KCore/Interp/InterpData.cpp:699-947, InterpAdt.cpp:1069-1182
//   order-2 zone result + cellN validity         KCore/KCore/Interp/Interp2.cpp:185-342, :84-177
//   hexa volume for cross-zone selection         KCore/KCore/Metric/compVolOfStructCell.cpp:91-218
// Design differences (B200-first): the ADT is a flat array of 128-byte nodes
// with the split thresholds of each node precomputed at build time, so the
// query stack holds node ids only (4 B/level instead of the reference's
// 13-double frames); the candidate std::list is never materialised — the DFS
// is a resumable iterator and list membership of the jump cell is decided by
// walking that cell's own insertion path.
//
// Arithmetic contract: IEEE fp64, expression order as in the
// reference, compiled with -fmad=false.  Thresholds are the reference's.
#pragma once
#include <math.h>
#include <limits.h>
#include <stddef.h>

#ifdef __CUDACC__
#define CX_D __device__ __forceinline__
#define CX_DN static __device__ __noinline__
#define CX_TBL static __device__ const
#else
#define CX_D static inline
#define CX_DN static
#define CX_TBL static const
#endif

#define CX_NULL 0x7fffffff

namespace cx {

// constants: KCore/KCore/Def/DefCplusPlusConst.cpp:39-62, InterpData.cpp:30-32
#define CX_EPS_DET 1.e-16
#define CX_EPS_TETRA 1.e-4
#define CX_EPS_GEOM 1.e-8
#define CX_MAXF 1.7976931348623157e308

struct AdtNode {   // 128 B = one cache line; the first 64 bytes are all a traversal step reads at a non-candidate node
  double mid;      // descend left  iff a[dim] <= mid
  double rlo;      // descend right iff b[dim] >= rlo
  int left, right; // node ids (CX_NULL = none)
  int dim;         // depth % 6
  // Bounding boxes quantised to 16 bits per coordinate on the zone's bounding box and rounded outwards
  // (conservative; lo x,y,z then hi x,y,z): qb = this cell (pre-filter of the exact candidate test), ql / qr = the
  // whole subtree of the left / right child, so that a query decides whether to descend into a child WITHOUT
  // fetching the child
  unsigned short qb[6], ql[6], qr[6];
  // ---- second half: read for candidates only
  double bb[6];    // xmin,ymin,zmin,xmax,ymax,zmax of the cell
  int ind;         // donor node index i + ni*j + ni*nj*k of the cell corner
  unsigned short qs[6];   // this node's whole subtree (build only)
};
static_assert(sizeof(AdtNode) == 128, "AdtNode must stay one 128-byte line");

// quantisation of coordinate d (0..2) on the zone bounding box: q = (v - org) * scl, the box maps to [250, 65250]
#ifdef __CUDACC__
#define CX_HD __host__ __device__ __forceinline__
#else
#define CX_HD static inline
#endif
CX_HD void quant_params(const double* bbox, int d, double& org, double& scl)
{
  double ext = bbox[d + 3] - bbox[d];
  if (!(ext > 0.)) ext = 1.;
  scl = 65000.0 / ext;
  org = bbox[d] - 250.0 / scl;
}
CX_D unsigned short quant_lo(double v, double org, double scl)
{ // two quanta of slack cover the 1e-8 inflation of the candidate test and every rounding of this formula
  double t = floor((v - org) * scl) - 2.;
  return (unsigned short)(t < 0. ? 0. : (t > 65535. ? 65535. : t));
}
CX_D unsigned short quant_hi(double v, double org, double scl)
{
  double t = ceil((v - org) * scl) + 2.;
  return (unsigned short)(t < 0. ? 0. : (t > 65535. ? 65535. : t));
}

struct ZoneView {      // device view of one donor zone
  const double* x; const double* y; const double* z; const double* cellN;
  const AdtNode* nodes;
  int ni, nj, nk;
  int ncells;
  int depth;           // max node depth (stack bound)
  double bbox[6];      // xmin,ymin,zmin,xmax,ymax,zmax (boundary-cell bbox; InterpCart: first point .. first + (n-1)*h)
  // topo 1: ADT (InterpAdt), topo 0: regular Cartesian donor located in closed form (InterpCart,
  // KCore/KCore/Interp/InterpCart.cpp:42-66): h = spacing, hinv = 1/h, hs2 = h/2; origin and end = bbox
  int topo;
  double h[3], hinv[3], hs2[3];
  double qorg[3], qscl[3];   // quantisation of the subtree boxes (quant_params of bbox)
};

// double -> int conversion with x86 cvttsd2si semantics (out of range / NaN ->
// INT_MIN), which is what the reference's E_Int(...) casts do on the CPU.
CX_D int trunc_int(double v)
{
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
  return (int)v;
}
CX_D double fsign(double a) { return (a < 0.0) ? -1.0 : 1.0; }  // DefFunction.h:145-148
CX_D int imin(int a, int b) { return a < b ? a : b; }
CX_D int imax(int a, int b) { return a > b ? a : b; }
CX_D bool eqzero(double v, double p) { return v >= -p && v <= p; }   // DefFunction.h:189-192

// ---------------------------------------------------------------------------
// ADT depth-first iterator (node, right subtree, left subtree).
// ---------------------------------------------------------------------------
#define CX_STACK 192
struct AdtQuery {
  double a[6], b[6];    // query box of InterpAdt.cpp:678-682
  double px, py, pz, alphaTol;
  int stack[CX_STACK];
  int sp;
  int prune;            // 1: skip subtrees whose bounding box cannot hold a candidate (alphaTol == 0 only)
  int tlo[3], thi[3];   // the point on the quantisation grid, one quantum of slack either side
};

// No node of a subtree can pass node_is_candidate (alphaTol = 0) when the point is outside the subtree box
// inflated by 1e-8: fl(xmax_c + 1e-8) is monotone in xmax_c, so x < xmax_c + 1e-8 implies x < sub_xmax + 1e-8.
// The quantised box contains that inflated box (quant_lo / quant_hi round outwards with slack), so skipping a
// subtree whose quantised box misses the (equally slackened) quantised point leaves the reference's candidate
// list and its order unchanged while cutting the visited nodes several-fold.  Codes clamped to 0 / 65535 stand for
// "unbounded": a point that passed zone_reject lies inside the zone box, i.e. strictly between them.
CX_D void quant_point(const ZoneView& Z, double x, double y, double z, int* tlo, int* thi)
{
  const double t0 = (x - Z.qorg[0]) * Z.qscl[0], t1 = (y - Z.qorg[1]) * Z.qscl[1], t2 = (z - Z.qorg[2]) * Z.qscl[2];
  tlo[0] = (int)floor(t0) - 1; thi[0] = (int)ceil(t0) + 1;
  tlo[1] = (int)floor(t1) - 1; thi[1] = (int)ceil(t1) + 1;
  tlo[2] = (int)floor(t2) - 1; thi[2] = (int)ceil(t2) + 1;
}
CX_D bool qbox_may_hold(const unsigned short* q, const int* tlo, const int* thi)
{
  return thi[0] >= (int)q[0] && thi[1] >= (int)q[1] && thi[2] >= (int)q[2] &&
         tlo[0] <= (int)q[3] && tlo[1] <= (int)q[4] && tlo[2] <= (int)q[5];
}

CX_D bool zone_reject(const ZoneView& Z, double x, double y, double z, double alphaTol)
{ // InterpAdt.cpp:662-672
  double dx = 0.1 * alphaTol * (Z.bbox[3] - Z.bbox[0]);
  double dy = 0.1 * alphaTol * (Z.bbox[4] - Z.bbox[1]);
  double dz = 0.1 * alphaTol * (Z.bbox[5] - Z.bbox[2]);
  if (x < Z.bbox[0] - dx) return true;
  if (y < Z.bbox[1] - dy) return true;
  if (z < Z.bbox[2] - dz) return true;
  if (x > Z.bbox[3] + dx) return true;
  if (y > Z.bbox[4] + dy) return true;
  if (z > Z.bbox[5] + dz) return true;
  return Z.ncells <= 0;
}

CX_D void query_init(AdtQuery& q, const ZoneView& Z, double x, double y, double z, double alphaTol)
{
  q.a[0] = Z.bbox[0]; q.a[1] = Z.bbox[1]; q.a[2] = Z.bbox[2];
  q.a[3] = x; q.a[4] = y; q.a[5] = z;
  q.b[0] = x; q.b[1] = y; q.b[2] = z;
  q.b[3] = Z.bbox[3]; q.b[4] = Z.bbox[4]; q.b[5] = Z.bbox[5];
  q.px = x; q.py = y; q.pz = z; q.alphaTol = alphaTol;
  q.stack[0] = 0; q.sp = 1;
  q.prune = (alphaTol == 0.0) ? 1 : 0;
  quant_point(Z, x, y, z, q.tlo, q.thi);
}

CX_D bool node_is_candidate(const AdtNode& n, double x, double y, double z, double alphaTol)
{ // InterpAdt.cpp:746-766
  double dx = alphaTol * (n.bb[3] - n.bb[0]) + CX_EPS_GEOM;
  double dy = alphaTol * (n.bb[4] - n.bb[1]) + CX_EPS_GEOM;
  double dz = alphaTol * (n.bb[5] - n.bb[2]) + CX_EPS_GEOM;
  return (x < n.bb[3] + dx && x > n.bb[0] - dx &&
          y < n.bb[4] + dy && y > n.bb[1] - dy &&
          z < n.bb[5] + dz && z > n.bb[2] - dz);
}

// Next candidate cell (its donor corner index) in the reference's list order,
// or -1 when the traversal is exhausted.  *nodeId receives the ADT node id.
CX_D int query_next(AdtQuery& q, const AdtNode* __restrict__ nodes, int* nodeId, int* nvisit)
{
  while (q.sp > 0)
  {
    int id = q.stack[--q.sp];
    const AdtNode& n = nodes[id];
    (*nvisit)++;
    int d = n.dim;
    // left pushed first, right second => right popped first (InterpAdt.cpp:768-925)
    if (q.a[d] <= n.mid && n.left != CX_NULL && (!q.prune || qbox_may_hold(n.ql, q.tlo, q.thi))) q.stack[q.sp++] = n.left;
    if (q.b[d] >= n.rlo && n.right != CX_NULL && (!q.prune || qbox_may_hold(n.qr, q.tlo, q.thi))) q.stack[q.sp++] = n.right;
    if ((!q.prune || qbox_may_hold(n.qb, q.tlo, q.thi)) && node_is_candidate(n, q.px, q.py, q.pz, q.alphaTol))
    { *nodeId = id; return n.ind; }
  }
  return -1;
}

// Is cell `cellId` (ADT node id) in the candidate list of query q?  It is iff
// the DFS reaches its node (every ancestor test on its insertion path passes)
// and its inflated bbox contains the point.
CX_D bool query_contains(const AdtQuery& q, const AdtNode* __restrict__ nodes, int cellId)
{
  const AdtNode& t = nodes[cellId];
  if (!node_is_candidate(t, q.px, q.py, q.pz, q.alphaTol)) return false;
  int id = 0;
  while (id != cellId)
  {
    const AdtNode& n = nodes[id];
    int d = n.dim;
    if (t.bb[d] <= n.mid)   // insertion rule, InterpAdt.cpp:537-547
    {
      if (!(q.a[d] <= n.mid)) return false;
      id = n.left;
    }
    else
    {
      if (!(q.b[d] >= n.rlo)) return false;
      id = n.right;
    }
    if (id == CX_NULL) return false;  // cannot happen for a built tree
  }
  return true;
}

// ---------------------------------------------------------------------------
// Cell geometry
// ---------------------------------------------------------------------------
struct Hexa { double x[15], y[15], z[15]; };

// InterpData::coordHexa (InterpData.cpp:56-134); ic,jc,kc 1-based.  nk == 1 (2-D zone in a z = constant plane):
// the top face is the bottom face lifted by 1 in z (:96-102).
CX_D void coord_hexa(const ZoneView& Z, int ind, int& ic, int& jc, int& kc, Hexa& H)
{
  const int ni = Z.ni, nij = Z.ni * Z.nj;
  kc = ind / nij + 1;
  int kcnij = kc * nij, kcmnij = (kc - 1) * nij;
  jc = (ind - kcmnij) / ni + 1;
  int jcni = jc * ni, jcmni = (jc - 1) * ni;
  ic = ind - jcmni - kcmnij + 1;
  int i0 = ind;
  int i1 = ic + jcmni + kcmnij;
  int i2 = ic - 1 + jcni + kcmnij;
  int i3 = ic + jcni + kcmnij;
  int i4 = ic - 1 + jcmni + kcnij;
  int i5 = ic + jcmni + kcnij;
  int i6 = ic - 1 + jcni + kcnij;
  int i7 = ic + jcni + kcnij;
  if (Z.nk > 1)
  {
    const int idx[8] = {i0, i1, i2, i3, i4, i5, i6, i7};
#pragma unroll
    for (int v = 0; v < 8; v++) { H.x[v] = Z.x[idx[v]]; H.y[v] = Z.y[idx[v]]; H.z[v] = Z.z[idx[v]]; }
  }
  else
  {
    const int idx[4] = {i0, i1, i2, i3};
#pragma unroll
    for (int v = 0; v < 4; v++)
    {
      H.x[v] = Z.x[idx[v]]; H.y[v] = Z.y[idx[v]]; H.z[v] = Z.z[idx[v]];
      H.x[v + 4] = H.x[v]; H.y[v + 4] = H.y[v]; H.z[v + 4] = H.z[v] + 1.;
    }
  }
#define CX_FC(o, a, b, c, d)                                  \
  H.x[o] = 0.25 * (H.x[a] + H.x[b] + H.x[c] + H.x[d]);        \
  H.y[o] = 0.25 * (H.y[a] + H.y[b] + H.y[c] + H.y[d]);        \
  H.z[o] = 0.25 * (H.z[a] + H.z[b] + H.z[c] + H.z[d]);
  CX_FC(8, 0, 1, 2, 3)
  CX_FC(9, 0, 2, 4, 6)
  CX_FC(10, 4, 5, 6, 7)
  CX_FC(11, 1, 3, 5, 7)
  CX_FC(12, 2, 3, 6, 7)
  CX_FC(13, 0, 1, 4, 5)
#undef CX_FC
  H.x[14] = 0.5 * (H.x[8] + H.x[10]);
  H.y[14] = 0.5 * (H.y[8] + H.y[10]);
  H.z[14] = 0.5 * (H.z[8] + H.z[10]);
}

// InterpData::coeffInterpTetra (InterpData.cpp:429-485)
CX_D void tetra_solve(double x, double y, double z, const Hexa& H, int p, int q, int r, int s,
                      double& xi, double& yi, double& zi)
{
  double xp = H.x[p], yp = H.y[p], zp = H.z[p];
  double a11 = H.x[q] - xp, a12 = H.x[r] - xp, a13 = H.x[s] - xp;
  double a21 = H.y[q] - yp, a22 = H.y[r] - yp, a23 = H.y[s] - yp;
  double a31 = H.z[q] - zp, a32 = H.z[r] - zp, a33 = H.z[s] - zp;
  double c11 = a22 * a33 - a32 * a23;
  double c12 = -(a21 * a33 - a31 * a23);
  double c13 = a21 * a32 - a31 * a22;
  double c21 = -(a12 * a33 - a32 * a13);
  double c22 = a11 * a33 - a31 * a13;
  double c23 = -(a11 * a32 - a31 * a12);
  double c31 = a12 * a23 - a22 * a13;
  double c32 = -(a11 * a23 - a21 * a13);
  double c33 = a11 * a22 - a21 * a12;
  double det = a11 * c11 + a12 * c12 + a13 * c13;
  if (fabs(det) < CX_EPS_DET) { xi = -10.; yi = -10.; zi = -10.; }
  else
  {
    double xpm = x - xp, ypm = y - yp, zpm = z - zp;
    det = 1.0 / det;
    xi = (c11 * xpm + c21 * ypm + c31 * zpm) * det;
    yi = (c12 * xpm + c22 * ypm + c32 * zpm) * det;
    zi = (c13 * xpm + c23 * ypm + c33 * zpm) * det;
  }
}

// Lookup tables of the 24-tetrahedra decomposition (InterpData.cpp:158-211).
CX_TBL signed char T_neighbour[24] = {1,2,4,0,3,5,3,0,6,2,1,7,5,0,6,4,7,1,7,4,2,6,5,3};
CX_TBL signed char T_center[32] = {13,8,9,13,13,8,11,13,12,8,9,12,12,8,11,12,10,13,9,
                                   10,13,10,11,13,12,10,9,12,12,10,11,12};
CX_TBL signed char T_indss[32] = {0,3,5,6,1,2,4,7,2,4,7,1,3,5,6,0,4,7,1,2,
                                  5,6,0,3,6,0,3,5,7,1,2,4};
CX_TBL signed char T_indfc[24] = {0,1,3,2,4,0,2,6,4,5,7,6,5,1,3,7,6,7,3,2,4,5,1,0};
CX_TBL signed char T_indtr[64] = {5,13,5,13,7,15,7,15, 5,13,5,13,7,15,7,15,
                                  23,21,19,17,23,21,19,17, 4,12,6,14,4,12,6,14,
                                  0,0,2,2,8,8,10,10, 3,1,3,1,11,9,11,9,
                                  22,22,18,18,20,20,16,16, 22,22,18,18,20,20,16,16};
CX_TBL signed char T_indtd[96] = {0,1,8,14, 1,3,8,14, 2,3,8,14, 0,2,8,14, 0,4,9,14, 0,2,9,14,
                                  2,6,9,14, 4,6,9,14, 4,5,10,14, 5,7,10,14, 6,7,10,14, 4,6,10,14,
                                  1,5,11,14, 1,3,11,14, 3,7,11,14, 5,7,11,14, 6,7,12,14, 3,7,12,14,
                                  2,3,12,14, 2,6,12,14, 4,5,13,14, 1,5,13,14, 0,1,13,14, 0,4,13,14};

// tetra (xi,yi,zi) -> 8 hexa weights (InterpData.cpp:302-328, :621-631)
CX_D void tetra_to_hexa(double xi, double yi, double zi, int indq, int indr, int inds, double* cf)
{
  double cf0 = 0.125 * (1.0 - xi - yi - zi);
#pragma unroll
  for (int i = 0; i < 8; i++) cf[i] = cf0;
  double tl = xi * 0.25;
  int l = (indq - 8) * 4;
  cf[T_indfc[l]] += tl; cf[T_indfc[l + 1]] += tl; cf[T_indfc[l + 2]] += tl; cf[T_indfc[l + 3]] += tl;
  cf[indr] += yi;
  cf[inds] += zi;
}

CX_D bool tetra_accept(double xi, double yi, double zi)
{ // InterpData.cpp:297-298
  return (xi > -CX_EPS_TETRA) && (yi > -CX_EPS_TETRA) && (zi > -CX_EPS_TETRA) &&
         (xi + yi + zi < 1.0 + 3 * CX_EPS_TETRA);
}

// InterpData::getCellJump (InterpData.cpp:383-425)
CX_D bool cell_jump(double x, double y, double z, const Hexa& H, int& isomm, double& xi, double& yi, double& zi)
{
  tetra_solve(x, y, z, H, 14, 11, 12, 10, xi, yi, zi);
  if (fabs(xi) <= 1.0 + CX_EPS_GEOM && fabs(yi) <= 1.0 + CX_EPS_GEOM && fabs(zi) <= 1.0 + CX_EPS_GEOM)
  {
    isomm = (xi >= 0 ? 1 : 0) + (yi >= 0 ? 2 : 0) + (zi >= 0 ? 4 : 0);
    return true;
  }
  return false;
}

// Shared tail of getCoeffInterpHexa / coeffInterpHexa: flip signs by corner,
// predict the tetrahedron, then scan the 24 in the reference's order
// (InterpData.cpp:242-378, :563-684).
CX_D bool hexa_from_corner(double x, double y, double z, const Hexa& H, int isomm,
                           double xi, double yi, double zi, double* cf)
{
  if (!(isomm & 1)) xi = -xi;
  if (!(isomm & 2)) yi = -yi;
  if (!(isomm & 4)) zi = -zi;
  int is1 = trunc_int((1.0 + fsign(xi - yi)) * 0.5);
  int is2 = trunc_int((1.0 + fsign(yi - zi)) * 0.5);
  int is3 = trunc_int((1.0 + fsign(zi - xi)) * 0.5);
  int itri = T_indtr[isomm + 8 * is1 + 16 * is2 + 32 * is3];
  int indr = T_indtd[0 + itri * 4], inds = T_indtd[1 + itri * 4], indq = T_indtd[2 + itri * 4];
  tetra_solve(x, y, z, H, 14, indq, indr, inds, xi, yi, zi);
  if (tetra_accept(xi, yi, zi)) { tetra_to_hexa(xi, yi, zi, indq, indr, inds, cf); return true; }
  for (int ibsom = 0; ibsom < 4; ibsom++)
  {
    int isom = T_indss[ibsom + 4 * isomm];
    indr = isom;
    for (int its = 0; its < 3; its++)
    {
      inds = T_neighbour[its + isom * 3];
      for (int itq = 0; itq < 2; itq++)
      {
        indq = T_center[itq + its + isom * 4];
        tetra_solve(x, y, z, H, 14, indq, indr, inds, xi, yi, zi);
        if (tetra_accept(xi, yi, zi)) { tetra_to_hexa(xi, yi, zi, indq, indr, inds, cf); return true; }
      }
    }
  }
  return false;
}

// InterpData::coeffInterpHexa (InterpData.cpp:141-379)
CX_D bool coeff_interp_hexa(double x, double y, double z, const Hexa& H, double* cf)
{
  double xi, yi, zi;
  tetra_solve(x, y, z, H, 14, 11, 12, 10, xi, yi, zi);
  int isomm = 0;
  if (fabs(xi) <= 1.0 + CX_EPS_GEOM && fabs(yi) <= 1.0 + CX_EPS_GEOM && fabs(zi) <= 1.0 + CX_EPS_GEOM)
    isomm = (xi >= 0 ? 1 : 0) + (yi >= 0 ? 2 : 0) + (zi >= 0 ? 4 : 0);
  return hexa_from_corner(x, y, z, H, isomm, xi, yi, zi, cf);
}

// K_METRIC::compVolOfStructCell3D with indcell=-1 (compVolOfStructCell.cpp:91-218)
CX_D double cell_volume(const ZoneView& Z, int indnode)
{
  const int ni = Z.ni, nj = Z.nj, nk = Z.nk, ninj = ni * nj;
  int k = indnode / ninj;
  int j = indnode / ni - k * nj;
  int i = indnode - k * ninj - j * ni;
  if (i == ni - 1) indnode = indnode - 1;
  if (j == nj - 1) indnode = indnode - ni;
  if (k == nk - 1) indnode = indnode - ninj;
  int iA = indnode, iB = iA + 1, iC = iB + ni, iD = iA + ni;
  int iE = iA + ninj, iF = iB + ninj, iG = iC + ninj, iH = iD + ninj;
  const double* X = Z.x; const double* Y = Z.y; const double* W = Z.z;
  double xA = X[iA], yA = Y[iA], zA = W[iA], xB = X[iB], yB = Y[iB], zB = W[iB];
  double xC = X[iC], yC = Y[iC], zC = W[iC], xD = X[iD], yD = Y[iD], zD = W[iD];
  double xE = X[iE], yE = Y[iE], zE = W[iE], xF = X[iF], yF = Y[iF], zF = W[iF];
  double xG = X[iG], yG = Y[iG], zG = W[iG], xH = X[iH], yH = Y[iH], zH = W[iH];
  double xcc, ycc, zcc, xv1, yv1, zv1, xv2, yv2, zv2, nx, ny, nz;
#define CX_FACE(P, Q, R, S, V1a, V1b, V2a, V2b, out)                              \
  xcc = x##P + x##Q + x##R + x##S; ycc = y##P + y##Q + y##R + y##S; zcc = z##P + z##Q + z##R + z##S; \
  xv1 = x##V1a - x##V1b; yv1 = y##V1a - y##V1b; zv1 = z##V1a - z##V1b;           \
  xv2 = x##V2a - x##V2b; yv2 = y##V2a - y##V2b; zv2 = z##V2a - z##V2b;           \
  nx = yv1 * zv2 - zv1 * yv2; ny = zv1 * xv2 - xv1 * zv2; nz = xv1 * yv2 - yv1 * xv2; \
  double out = xcc * nx + ycc * ny + zcc * nz;
  CX_FACE(A, B, C, D, C, A, D, B, v1)
  CX_FACE(E, F, G, H, G, E, H, F, v2)
  CX_FACE(A, D, H, E, D, E, H, A, v3)
  CX_FACE(B, F, G, C, C, F, G, B, v4)
  CX_FACE(A, B, F, E, B, E, A, F, v5)
  CX_FACE(D, C, G, H, C, H, D, G, v6)
#undef CX_FACE
  const double ONE_24TH = 0.125 * 0.33333333333333333333;
  return fabs(v2 - v1 + v4 - v3 + v6 - v5) * ONE_24TH;
}

// K_METRIC::compVolOfStructCell2D (Metric/compVolOfStructCell.cpp:29-89) with indnode given: signed area of the quad
CX_D double cell_area_2d(const ZoneView& Z, int indnode)
{
  const int ni = Z.ni;
  const int ind1 = indnode, ind2 = ind1 + ni, ind3 = ind1 + 1, ind4 = ind2 + 1;
  const double* X = Z.x; const double* Y = Z.y; const double* W = Z.z;
  double l1x = X[ind1] - X[ind2], l1y = Y[ind1] - Y[ind2], l1z = W[ind1] - W[ind2];
  double l2x = X[ind1] - X[ind3], l2y = Y[ind1] - Y[ind3], l2z = W[ind1] - W[ind3];
  double surf1x = (l1y * l2z - l1z * l2y), surf1y = (l1z * l2x - l1x * l2z), surf1z = (l1x * l2y - l1y * l2x);
  double surface1 = sqrt(surf1x * surf1x + surf1y * surf1y + surf1z * surf1z);
  double l3x = X[ind4] - X[ind3], l3y = Y[ind4] - Y[ind3], l3z = W[ind4] - W[ind3];
  double l4x = X[ind4] - X[ind2], l4y = Y[ind4] - Y[ind2], l4z = W[ind4] - W[ind2];
  double surf2x = (l3y * l4z - l3z * l4y), surf2y = (l3z * l4x - l3x * l4z), surf2z = (l3x * l4y - l3y * l4x);
  double surface2 = sqrt(surf2x * surf2x + surf2y * surf2y + surf2z * surf2z);
  double ps = surf1x * surf2x + surf1y * surf2y + surf1z * surf2z;
  return 0.5 * fsign(ps) * (surface1 + surface2);
}

struct SearchStats { int nvisit; int ncand; int ntet; };

// InterpAdt::searchInterpolationCellStruct (InterpAdt.cpp:1252-1400).
// Returns 1 and (ic,jc,kc) 1-based + cf[8] when found, else 0.
CX_DN int search_interp_cell(const ZoneView& Z, double x, double y, double z,
                             int& ic, int& jc, int& kc, double* cf, SearchStats* st)
{
  if (zone_reject(Z, x, y, z, 0.0)) return 0;
  AdtQuery q;
  query_init(q, Z, x, y, z, 0.0);
  int nid, nvisit = 0;
  int first = query_next(q, Z.nodes, &nid, &nvisit);
  if (st) { st->nvisit += nvisit; }
  if (first < 0) return 0;
  const int ni = Z.ni, nj = Z.nj, nk = Z.nk, nij = ni * nj;
  Hexa H;
  bool JUMP = false;
  int ind = first, isomm = 0;
  double xi = 0, yi = 0, zi = 0;
  int erased = -1;
  for (int i = 0; i < 8; i++)
  {
    coord_hexa(Z, ind, ic, jc, kc, H);
    if (!cell_jump(x, y, z, H, isomm, xi, yi, zi))
    {
      if (i == 7) break;
      int is = imin(8, trunc_int((xi + fsign(xi)) * 0.5)); is = imax(-8, is);
      int js = imin(8, trunc_int((yi + fsign(yi)) * 0.5)); js = imax(-8, js);
      int ks = imin(8, trunc_int((zi + fsign(zi)) * 0.5)); ks = imax(-8, ks);
      ind = ind + is + js * ni + ks * nij;
      kc = ind / nij;
      int kcnij = kc * nij;
      jc = (ind - kcnij) / ni;
      int jcni = jc * ni;
      ic = ind - jcni - kcnij;
      if (ic < 0 || ic > ni - 2 || jc < 0 || jc > nj - 2 || kc < 0 || kc > nk - 2) break;
    }
    else
    {
      // list membership of the jump cell (InterpAdt.cpp:1342-1353)
      int cid = (ic - 1) + (ni - 1) * ((jc - 1) + (nj - 1) * (kc - 1));
      JUMP = (ind == first) ? true : query_contains(q, Z.nodes, cid);
      break;
    }
  }
  if (JUMP)
  {
    if (hexa_from_corner(x, y, z, H, isomm, xi, yi, zi, cf)) return 1;
    erased = ind;
  }
  // ordered fallback over the whole candidate list (InterpAdt.cpp:1378-1397)
  query_init(q, Z, x, y, z, 0.0);
  nvisit = 0;
  int c;
  while ((c = query_next(q, Z.nodes, &nid, &nvisit)) >= 0)
  {
    if (c == erased) continue;
    coord_hexa(Z, c, ic, jc, kc, H);
    if (coeff_interp_hexa(x, y, z, H, cf)) { if (st) st->nvisit += nvisit; return 1; }
  }
  if (st) st->nvisit += nvisit;
  return 0;
}

// ---------------------------------------------------------------------------
// Regular Cartesian donors (interpDataType=0, K_INTERP::InterpCart): no tree, the cell comes from
// the point's offset to the first node.  Offsets (_ioff,_joff,_koff) are 1 for zones built by
// setInterpData (setInterpData.cpp:790-804: InterpCart(ni,nj,nk,hi,hj,hk,x0,y0,z0)).
// ---------------------------------------------------------------------------
CX_D bool cart_outside(const ZoneView& Z, double x, double y, double z)
{ // InterpCart.cpp:76-82
  const double EPS = CX_EPS_GEOM;
  return (x < Z.bbox[0] - EPS) || (x > Z.bbox[3] + EPS) || (y < Z.bbox[1] - EPS) || (y > Z.bbox[4] + EPS) ||
         (z < Z.bbox[2] - EPS) || (z > Z.bbox[5] + EPS);
}

// InterpCart::searchInterpolationCellCartO2 (InterpCart.cpp:71-125): trilinear weights, (ic,jc,kc) 1-based
CX_D int cart_search_o2(const ZoneView& Z, double x, double y, double z, int& ic, int& jc, int& kc, double* cf)
{
  if (cart_outside(Z, x, y, z)) return 0;
  const double xmin = Z.bbox[0], ymin = Z.bbox[1], zmin = Z.bbox[2];
  ic = trunc_int((x - xmin) * Z.hinv[0]) + 1;
  jc = trunc_int((y - ymin) * Z.hinv[1]) + 1;
  kc = trunc_int((z - zmin) * Z.hinv[2]) + 1;
  ic = imin(ic, Z.ni - 1); ic = imax(ic, 1);
  jc = imin(jc, Z.nj - 1); jc = imax(jc, 1);
  kc = imin(kc, Z.nk - 1); kc = imax(kc, 1);
  double x0 = xmin + (ic - 1) * Z.h[0];
  double y0 = ymin + (jc - 1) * Z.h[1];
  double z0 = zmin + (kc - 1) * Z.h[2];
  double x1 = (x - x0) * Z.hinv[0];
  double x2 = (y - y0) * Z.hinv[1];
  double x3 = (z - z0) * Z.hinv[2];
  double t1 = (1. - x1), t2 = (1. - x2), t3 = (1. - x3);
  double t1t2 = t1 * t2, x2t3 = x2 * t3, x1t2 = x1 * t2, x2x3 = x2 * x3;
  cf[0] = t1t2 * t3;
  cf[1] = x1t2 * t3;
  cf[2] = t1 * x2t3;
  cf[3] = x1 * x2t3;
  if (Z.nk <= 2 && fabs(Z.bbox[5] - zmin) < CX_EPS_GEOM) { cf[4] = 0.; cf[5] = 0.; cf[6] = 0.; cf[7] = 0.; }
  else
  {
    cf[4] = t1t2 * x3;
    cf[5] = x1t2 * x3;
    cf[6] = t1 * x2x3;
    cf[7] = x1 * x2x3;
  }
  return 1;
}

// one direction of InterpCart::searchInterpolationCellCartO3 (InterpCart.cpp:151-266, 288-318):
// 3-node stencil around the cell (shifted left unless the point is in the right half), 1-D Lagrange weights
CX_D void cart_o3_dir(double x, double xmin, double h, double hs2, int n, int& ic, double* l)
{
  double x_i = xmin + (ic - 1) * h, x_ip = x_i + h;
  double a, b, c;
  int icHO;
  if (x - x_i >= hs2)
  {
    if (ic >= n - 1) { icHO = ic - 1; a = x_i - h; b = x_i; c = x_ip; }
    else { icHO = ic; a = x_i; b = x_ip; c = b + h; }
  }
  else
  {
    if (ic < 3) { icHO = ic; a = x_i; b = x_ip; c = b + h; }
    else { icHO = ic - 1; a = x_i - h; b = x_i; c = x_ip; }
  }
  double hc = b - a, hcp = c - b, hcpp = c - a;
  double inv_hc = 1. / hc, inv_hcp = 1. / hcp, inv_hcpp = 1. / hcpp;
  l[0] = (x - b) * (x - c) * inv_hc * inv_hcpp;
  l[1] = -(x - a) * (x - c) * inv_hc * inv_hcp;
  l[2] = (x - a) * (x - b) * inv_hcp * inv_hcpp;
  ic = icHO;
}

// InterpCart::searchInterpolationCellCartO3: cf[9] by direction, (ic,jc,kc) = 1-based start of the 3x3x3 stencil
CX_D int cart_search_o3(const ZoneView& Z, double x, double y, double z, int& ic, int& jc, int& kc, double* cf)
{
  if (cart_outside(Z, x, y, z)) return 0;
  const double xmin = Z.bbox[0], ymin = Z.bbox[1], zmin = Z.bbox[2];
  ic = trunc_int((x - xmin) * Z.hinv[0]) + 1;
  jc = trunc_int((y - ymin) * Z.hinv[1]) + 1;
  kc = trunc_int((z - zmin) * Z.hinv[2]) + 1;
  ic = imax(ic, 1); ic = imin(ic, Z.ni - 1);
  jc = imax(jc, 1); jc = imin(jc, Z.nj - 1);
  kc = imax(kc, 1); kc = imin(kc, Z.nk - 1);
  cart_o3_dir(x, xmin, Z.h[0], Z.hs2[0], Z.ni, ic, cf);
  cart_o3_dir(y, ymin, Z.h[1], Z.hs2[1], Z.nj, jc, cf + 3);
  cart_o3_dir(z, zmin, Z.h[2], Z.hs2[2], Z.nk, kc, cf + 6);
  return 1;
}

// one direction of InterpCart::searchInterpolationCellCartO4 (InterpCart.cpp:347-412): 4-node stencil shifted one
// node to the left of the cell and pulled back inside at the max boundary, cubic 1-D Lagrange weights
CX_D void cart_o4_dir(double x, double xmin, double h, int n, int& ic, double* l)
{
  if (ic > 1) ic -= 1;
  if (ic + 3 > n) ic -= ic + 3 - n;
  double x_i = xmin + (ic - 1) * h, x_ip = x_i + h, x_ipp = x_i + 2 * h, x_ippp = x_i + 3 * h;
  double norm = -6. * h * h * h;
  l[0] = (x - x_ip) * (x - x_ipp) * (x - x_ippp) / norm;
  norm = 2. * h * h * h;
  l[1] = (x - x_i) * (x - x_ipp) * (x - x_ippp) / norm;
  norm = -2. * h * h * h;
  l[2] = (x - x_i) * (x - x_ip) * (x - x_ippp) / norm;
  norm = 6. * h * h * h;
  l[3] = (x - x_i) * (x - x_ip) * (x - x_ipp) / norm;
}

// InterpCart::searchInterpolationCellCartO4: cf[12] by direction, (ic,jc,kc) = 1-based start of the 4x4x4 stencil
CX_D int cart_search_o4(const ZoneView& Z, double x, double y, double z, int& ic, int& jc, int& kc, double* cf)
{
  if (cart_outside(Z, x, y, z)) return 0;
  const double xmin = Z.bbox[0], ymin = Z.bbox[1], zmin = Z.bbox[2];
  ic = trunc_int((x - xmin) * Z.hinv[0]) + 1;
  jc = trunc_int((y - ymin) * Z.hinv[1]) + 1;
  kc = trunc_int((z - zmin) * Z.hinv[2]) + 1;
  ic = imax(ic, 1); ic = imin(ic, Z.ni - 1);
  jc = imax(jc, 1); jc = imin(jc, Z.nj - 1);
  kc = imax(kc, 1); kc = imin(kc, Z.nk - 1);
  cart_o4_dir(x, xmin, Z.h[0], Z.ni, ic, cf);
  cart_o4_dir(y, ymin, Z.h[1], Z.nj, jc, cf + 4);
  cart_o4_dir(z, zmin, Z.h[2], Z.nk, kc, cf + 8);
  return 1;
}

// InterpCart::getListOfCandidateCells (InterpCart.cpp:423-474): the box of cells within 0.1*alphaTol*h of the
// point, enumerated k-outer / i-inner.  Returns the number of candidates (0: outside) and the 1-based bounds.
CX_D int cart_candidate_box(const ZoneView& Z, double x, double y, double z, double alphaTol, int* lo, int* hi)
{
  const double dx = 0.1 * alphaTol * Z.h[0], dy = 0.1 * alphaTol * Z.h[1], dz = 0.1 * alphaTol * Z.h[2];
  const double xmin = Z.bbox[0], ymin = Z.bbox[1], zmin = Z.bbox[2];
  if (x < xmin - dx) return 0;
  if (y < ymin - dy) return 0;
  if (Z.nk > 1 && z < zmin - dz) return 0;
  if (x > Z.bbox[3] + dx) return 0;
  if (y > Z.bbox[4] + dy) return 0;
  if (Z.nk > 1 && z > Z.bbox[5] + dz) return 0;
  double x0 = x - dx, x1 = x + dx, y0 = y - dy, y1 = y + dy, z0 = z - dz, z1 = z + dz;
  int icmin = trunc_int((x0 - xmin) * Z.hinv[0]) + 1, jcmin = trunc_int((y0 - ymin) * Z.hinv[1]) + 1;
  int kcmin = trunc_int((z0 - zmin) * Z.hinv[2]) + 1;
  int icmax = trunc_int((x1 - xmin) * Z.hinv[0]) + 1, jcmax = trunc_int((y1 - ymin) * Z.hinv[1]) + 1;
  int kcmax = trunc_int((z1 - zmin) * Z.hinv[2]) + 1;
  if (icmin < 1 && icmax < 1) return 0;
  if (jcmin < 1 && jcmax < 1) return 0;
  if (kcmin < 1 && kcmax < 1) return 0;
  if (icmin > Z.ni && icmax > Z.ni) return 0;
  if (jcmin > Z.nj && jcmax > Z.nj) return 0;
  if (kcmin > Z.nk && kcmax > Z.nk) return 0;
  icmin = imin(Z.ni - 1, icmin); icmin = imax(1, icmin); icmax = imin(Z.ni - 1, icmax); icmax = imax(1, icmax);
  jcmin = imin(Z.nj - 1, jcmin); jcmin = imax(1, jcmin); jcmax = imin(Z.nj - 1, jcmax); jcmax = imax(1, jcmax);
  kcmin = imin(Z.nk - 1, kcmin); kcmin = imax(1, kcmin); kcmax = imin(Z.nk - 1, kcmax); kcmax = imax(1, kcmax);
  lo[0] = icmin; lo[1] = jcmin; lo[2] = kcmin; hi[0] = icmax; hi[1] = jcmax; hi[2] = kcmax;
  int ni_ = icmax - icmin + 1, nj_ = jcmax - jcmin + 1, nk_ = kcmax - kcmin + 1;
  if (ni_ <= 0 || nj_ <= 0 || nk_ <= 0) return 0;
  return ni_ * nj_ * nk_;
}

// cellN validity of an order-2 stencil (Interp2.cpp:269-298)
CX_D bool celln_valid_o2(const ZoneView& Z, int ic, int jc, int kc, const double* cf, int nature)
{
  if (Z.cellN == nullptr) return true;
  const int ni = Z.ni, nij = Z.ni * Z.nj;
  if (nature == 0)
  {
    double val = 1.;
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
      const int ii = q & 1, jj = (q >> 1) & 1, kk = q >> 2;
      double c0 = Z.cellN[(ic + ii) + (jc + jj) * ni + (kc + kk) * nij];
      int d = cf[q] > 1.e-12;
      val *= c0 - d + 1;
    }
    return !eqzero(val, CX_EPS_GEOM);
  }
  double val = 0.;
#pragma unroll
  for (int q = 0; q < 8; q++)
  {
    const int ii = q & 1, jj = (q >> 1) & 1, kk = q >> 2;
    double c0 = Z.cellN[(ic + ii) + (jc + jj) * ni + (kc + kk) * nij];
    val += fabs(cf[q]) * fabs(1. - c0);
  }
  return eqzero(val, CX_EPS_GEOM);
}

// cellN validity of a 2-D order-2 stencil, coefficients already folded onto 4 nodes (Interp2.cpp:312-338)
CX_D bool celln_valid_o2_2d(const ZoneView& Z, int ic, int jc, const double* cf, int nature)
{
  if (Z.cellN == nullptr) return true;
  const int ni = Z.ni;
  if (nature == 0)
  {
    double val = 1.;
    for (int jj = 0; jj < 2; jj++) for (int ii = 0; ii < 2; ii++)
    {
      double c0 = Z.cellN[(ic + ii) + (jc + jj) * ni];
      int d = cf[ii + jj * 2] > 1.e-12;
      val *= c0 - d + 1;
    }
    return !eqzero(val, CX_EPS_GEOM);
  }
  double val = 0.;
  for (int jj = 0; jj < 2; jj++) for (int ii = 0; ii < 2; ii++)
  {
    double c0 = Z.cellN[(ic + ii) + (jc + jj) * ni];
    val += fabs(cf[ii + jj * 2]) * fabs(1. - c0);
  }
  return eqzero(val, CX_EPS_GEOM);
}

// InterpData::getExtrapolationCoeffForCell (InterpData.cpp:699-947), 3-D and 2-D (nk == 1: 4 nodes).
// (ic,jc,kc) 1-based on entry.  cf holds the last tetra's weights on return 0.
CX_DN int extrap_coeff_for_cell(const ZoneView& Z, double x, double y, double z, int ic, int jc, int kc,
                                double* cf, int nature, double constraint, double& diff_coeff)
{
  const int ni = Z.ni, nij = Z.ni * Z.nj;
  ic = ic - 1; jc = jc - 1; kc = imax(0, kc - 1);
  int ind = ic + jc * ni + kc * nij;
  int d0, d1, d2;
  Hexa H;
  coord_hexa(Z, ind, d0, d1, d2, H);
  double cfSumMin = CX_MAXF, cfSumMin2 = CX_MAXF;
  double cfSav[8], cfSav2[8];
  for (int i = 0; i < 8; i++) { cfSav[i] = 0.; cfSav2[i] = 0.; }
  bool valid = false;
  // validity of the 8 vertices does not depend on the tetrahedron
  const int nkk = (Z.nk == 1) ? 1 : 2;       // dim == 2: the 4 nodes of the plane only (InterpData.cpp:828-850)
  double val = 1.;
  if (Z.cellN != nullptr)
  {
    for (int kk = 0; kk < nkk; kk++) for (int jj = 0; jj < 2; jj++) for (int ii = 0; ii < 2; ii++)
    {
      double c0 = Z.cellN[(ic + ii) + (jc + jj) * ni + (kc + kk) * nij];
      if (nature == 0) val *= c0; else val *= c0 * (c0 - 2.);
    }
  }
  bool cellValid = !eqzero(val, CX_EPS_GEOM);
  const int isomm = 0;
  for (int ibsom = 0; ibsom < 4; ibsom++)
  {
    int isom = T_indss[ibsom + 4 * isomm];
    int indr = isom;
    for (int its = 0; its < 3; its++)
    {
      int inds = T_neighbour[its + isom * 3];
      for (int itq = 0; itq < 2; itq++)
      {
        int indq = T_center[itq + its + isom * 4];
        double xi, yi, zi;
        tetra_solve(x, y, z, H, 14, indq, indr, inds, xi, yi, zi);
        // InterpData.cpp:788-797 (xi*0.25 added four times)
        double cf0 = 0.125 * (1.0 - xi - yi - zi);
        for (int i = 0; i < 8; i++) cf[i] = cf0;
        for (int i = 0; i < 4; i++) cf[T_indfc[i + (indq - 8) * 4]] += xi * 0.25;
        cf[indr] += yi; cf[inds] += zi;
        double cfSum = fabs(cf[0]) + fabs(cf[1]) + fabs(cf[2]) + fabs(cf[3]) +
                       fabs(cf[4]) + fabs(cf[5]) + fabs(cf[6]) + fabs(cf[7]);
        if (cfSum < cfSumMin2) { cfSumMin2 = cfSum; for (int i = 0; i < 8; i++) cfSav2[i] = cf[i]; }
        if ((xi + yi + zi <= 1.0) && (xi >= 0.) && (yi >= 0.) && (zi >= 0.))
        { cfSumMin2 = 0.; for (int i = 0; i < 8; i++) cfSav2[i] = cf[i]; }
        if (cellValid && cfSum < cfSumMin)
        { cfSumMin = cfSum; for (int i = 0; i < 8; i++) cfSav[i] = cf[i]; valid = true; }
      }
    }
  }
  if (valid && cfSumMin < constraint)
  {
    double mn = CX_MAXF, mx = -CX_MAXF;
    for (int i = 0; i < 8; i++)
    {
      cf[i] = cfSav[i];
      if (fabs(cf[i]) > 1.e-10) { mn = fmin(mn, cf[i]); mx = fmax(mx, cf[i]); }
    }
    diff_coeff = mx - mn;
    return 1;
  }
  else if (cfSumMin2 < constraint)
  {
    double mn = CX_MAXF, mx = -CX_MAXF;
    for (int i = 0; i < 8; i++)
    {
      cf[i] = cfSav2[i];
      if (fabs(cf[i]) > 1.e-10) { mn = fmin(mn, cf[i]); mx = fmax(mx, cf[i]); }
    }
    diff_coeff = mx - mn;
    double loc[8], sum;
    for (int i = 0; i < 8; i++) loc[i] = 1.;
    if (Z.cellN != nullptr)
    {
      sum = 0.;
      int icell = 0;
      for (int kk = 0; kk < nkk; kk++) for (int jj = 0; jj < 2; jj++) for (int ii = 0; ii < 2; ii++)
      {
        double c0 = Z.cellN[(ic + ii) + (jc + jj) * ni + (kc + kk) * nij];
        if (nature == 1) loc[icell] = 1. - fabs(c0 - 1.);
        else if (nature == 0) loc[icell] = fmin(1., c0);
        sum += loc[icell];
        icell++;
      }
    }
    else sum = 8.;
    if (sum < 1.e-15) return 0;
    const int ncells = (Z.nk == 1) ? 4 : 8;
    if (Z.nk == 1)                            // 2-D: the 8 coefficients are folded onto the plane first (:927-934)
      for (int i = 0; i < 4; i++) { cf[i] += cf[i + 4]; cf[i + 4] = 0.; }
    double sum2 = 0.;
    for (int i = 0; i < ncells; i++) sum2 += (1. - loc[i]) * cf[i];
    double report = sum2 / sum;
    for (int i = 0; i < ncells; i++) cf[i] = (cf[i] + report) * loc[i];
    return 1;
  }
  return 0;
}

// InterpAdt::searchExtrapolationCellStruct (InterpAdt.cpp:1069-1182), extrapOrder=1.
CX_DN int search_extrap_cell(const ZoneView& Z, double x, double y, double z, int& ic, int& jc, int& kc,
                             double* cf, int nature, double constraint, SearchStats* st)
{
  for (int i = 0; i < 8; i++) cf[i] = 0.;
  double alphaTol = fmax((2 * constraint - 5.) * 0.1, 0.);
  if (Z.topo == 0)
  {
    // InterpCart::searchExtrapolationCellCart (InterpCart.cpp:478-571), extrapOrder=1: same fold over the box
    int lo[3], hi[3];
    if (cart_candidate_box(Z, x, y, z, alphaTol, lo, hi) == 0) return 0;
    double saved_sum = CX_MAXF, saved_diff = CX_MAXF, diff = CX_MAXF;
    int found = 0, icsav = 0, jcsav = 0, kcsav = 0, ncand = 0;
    double cfs[8], cft[8];
    for (int i = 0; i < 8; i++) { cfs[i] = 0.; cft[i] = 0.; }
    for (int k = lo[2]; k <= hi[2]; k++) for (int j = lo[1]; j <= hi[1]; j++) for (int i = lo[0]; i <= hi[0]; i++)
    {
      ncand++;
      int ret = extrap_coeff_for_cell(Z, x, y, z, i, j, k, cft, nature, constraint, diff);
      double s = fabs(cft[0]) + fabs(cft[1]) + fabs(cft[2]) + fabs(cft[3]) +
                 fabs(cft[4]) + fabs(cft[5]) + fabs(cft[6]) + fabs(cft[7]);
      if (ret == 1 && s < saved_sum + 1.e-6 && diff < saved_diff)
      {
        saved_diff = diff; found = 1; saved_sum = s;
        icsav = i; jcsav = j; kcsav = k;
        for (int m = 0; m < 8; m++) cfs[m] = cft[m];
      }
    }
    if (st) st->ncand += ncand;
    ic = icsav; jc = jcsav; kc = kcsav;
    for (int m = 0; m < 8; m++) cf[m] = cfs[m];
    return found;
  }
  if (zone_reject(Z, x, y, z, alphaTol)) return 0;
  AdtQuery q;
  query_init(q, Z, x, y, z, alphaTol);
  const int ni = Z.ni, nij = Z.ni * Z.nj;
  double saved_sum = CX_MAXF, saved_diff = CX_MAXF, diff = CX_MAXF;
  int found = 0, icsav = 0, jcsav = 0, kcsav = 0;
  double cfs[8], cft[8];
  for (int i = 0; i < 8; i++) { cfs[i] = 0.; cft[i] = 0.; }
  int c, nid, nvisit = 0, ncand = 0;
  while ((c = query_next(q, Z.nodes, &nid, &nvisit)) >= 0)
  {
    ncand++;
    int k = c / nij, j = (c - k * nij) / ni, i = c - j * ni - k * nij;
    k++; j++; i++;
    int ret = extrap_coeff_for_cell(Z, x, y, z, i, j, k, cft, nature, constraint, diff);
    double s = fabs(cft[0]) + fabs(cft[1]) + fabs(cft[2]) + fabs(cft[3]) +
               fabs(cft[4]) + fabs(cft[5]) + fabs(cft[6]) + fabs(cft[7]);
    if (ret == 1 && s < saved_sum + 1.e-6 && diff < saved_diff)
    {
      saved_diff = diff; found = 1; saved_sum = s;
      icsav = i; jcsav = j; kcsav = k;
      for (int m = 0; m < 8; m++) cfs[m] = cft[m];
    }
  }
  if (st) { st->nvisit += nvisit; st->ncand += ncand; }
  if (ncand == 0) return 0;   // empty list: cf stays zero, InterpAdt.cpp:1100
  ic = icsav; jc = jcsav; kc = kcsav;
  for (int m = 0; m < 8; m++) cf[m] = cfs[m];
  return found;
}

// ---------------------------------------------------------------------------
// Order-2 search with the dynamically indexed per-thread state OUT of local memory.
//
// The functions above keep the DFS stack (int[192]), the query box a[6]/b[6] and the 15 points of the cell
// under test (Hexa, 45 doubles, indexed by table look-ups) in per-thread arrays, i.e. in local memory: ptxas
// reported a 1904-byte stack frame for k_interp_o2 and ncu 20.5 GB of DRAM writes for 1.4 GB of results
// (profiles/r01_f_summary.md).  They remain in use for the rare extrapolation pass only.  Here
//   * the stack and the 8 vertices of the cell live in a Scratch area addressed with a stride (shared memory on
//     the device, column = thread, stride = threads per CTA: conflict-free; a plain array with stride 1 on the host);
//   * the query box is rebuilt per node from registers (a = (zone min, p), b = (p, zone max));
//   * face centres are recomputed from the scratch vertices when a tetrahedron needs one (same expression and
//     summation order as coord_hexa, so the same value), the cell centre is kept in three registers;
//   * the 8 hexa weights are produced with predicated adds instead of table-indexed stores.
// Every floating-point operation and its operand order is unchanged, so results are bit-identical to the
// functions above (tests/hostemu runs both against the oracle).
// ---------------------------------------------------------------------------
struct Scratch {
  int* stk;      // DFS stack, entry l at stk[l * ss]
  double* vtx;   // cell vertices: component c (0 x, 1 y, 2 z) of vertex v at vtx[(c * 8 + v) * ss]
  int ss;        // stride between consecutive entries
  int cap;       // stack capacity (>= depth + 2 of every zone searched)
};
struct AdtQueryS { double px, py, pz, alphaTol; int sp; int prune; int tlo[3], thi[3]; };

CX_D void query_init_s(AdtQueryS& q, const ZoneView& Z, const Scratch& S, double x, double y, double z, double alphaTol)
{
  q.px = x; q.py = y; q.pz = z; q.alphaTol = alphaTol;
  S.stk[0] = 0; q.sp = 1;
  q.prune = (alphaTol == 0.0) ? 1 : 0;
  quant_point(Z, x, y, z, q.tlo, q.thi);
}

// query_next with the stack in scratch and a[d], b[d] of InterpAdt.cpp:678-682 selected from registers
CX_D int query_next_s(AdtQueryS& q, const ZoneView& Z, const Scratch& S, int* nodeId, int* nvisit)
{
  const AdtNode* __restrict__ nodes = Z.nodes;
  while (q.sp > 0)
  {
    int id = S.stk[(--q.sp) * S.ss];
    const AdtNode& n = nodes[id];
    (*nvisit)++;
    const int d = n.dim;
    const int dd = d < 3 ? d : d - 3;
    const double pv = dd == 0 ? q.px : (dd == 1 ? q.py : q.pz);
    const double av = d < 3 ? Z.bbox[dd] : pv;          // a = (xmin, ymin, zmin, x, y, z)
    const double bv = d < 3 ? pv : Z.bbox[dd + 3];      // b = (x, y, z, xmax, ymax, zmax)
    const int nl = n.left, nr = n.right;
    // a child whose subtree box (kept in this node) cannot hold a candidate is never fetched
    if (av <= n.mid && nl != CX_NULL && q.sp < S.cap && (!q.prune || qbox_may_hold(n.ql, q.tlo, q.thi)))
      S.stk[(q.sp++) * S.ss] = nl;
    if (bv >= n.rlo && nr != CX_NULL && q.sp < S.cap && (!q.prune || qbox_may_hold(n.qr, q.tlo, q.thi)))
      S.stk[(q.sp++) * S.ss] = nr;
    if ((!q.prune || qbox_may_hold(n.qb, q.tlo, q.thi)) && node_is_candidate(n, q.px, q.py, q.pz, q.alphaTol))
    { *nodeId = id; return n.ind; }
  }
  return -1;
}

CX_D bool query_contains_s(const AdtQueryS& q, const ZoneView& Z, int cellId)
{
  const AdtNode* __restrict__ nodes = Z.nodes;
  const AdtNode& t = nodes[cellId];
  if (!node_is_candidate(t, q.px, q.py, q.pz, q.alphaTol)) return false;
  int id = 0;
  while (id != cellId)
  {
    const AdtNode& n = nodes[id];
    const int d = n.dim;
    const int dd = d < 3 ? d : d - 3;
    const double pv = dd == 0 ? q.px : (dd == 1 ? q.py : q.pz);
    if (t.bb[d] <= n.mid)
    {
      const double av = d < 3 ? Z.bbox[dd] : pv;
      if (!(av <= n.mid)) return false;
      id = n.left;
    }
    else
    {
      const double bv = d < 3 ? pv : Z.bbox[dd + 3];
      if (!(bv >= n.rlo)) return false;
      id = n.right;
    }
    if (id == CX_NULL) return false;
  }
  return true;
}

// InterpData::coeffInterpTetra on explicit points P (origin), Q, R, S
CX_D void tetra_solve_p(double x, double y, double z, double xp, double yp, double zp, double xq, double yq, double zq,
                        double xr, double yr, double zr, double xs, double ys, double zs,
                        double& xi, double& yi, double& zi)
{
  double a11 = xq - xp, a12 = xr - xp, a13 = xs - xp;
  double a21 = yq - yp, a22 = yr - yp, a23 = ys - yp;
  double a31 = zq - zp, a32 = zr - zp, a33 = zs - zp;
  double c11 = a22 * a33 - a32 * a23;
  double c12 = -(a21 * a33 - a31 * a23);
  double c13 = a21 * a32 - a31 * a22;
  double c21 = -(a12 * a33 - a32 * a13);
  double c22 = a11 * a33 - a31 * a13;
  double c23 = -(a11 * a32 - a31 * a12);
  double c31 = a12 * a23 - a22 * a13;
  double c32 = -(a11 * a23 - a21 * a13);
  double c33 = a11 * a22 - a21 * a12;
  double det = a11 * c11 + a12 * c12 + a13 * c13;
  if (fabs(det) < CX_EPS_DET) { xi = -10.; yi = -10.; zi = -10.; }
  else
  {
    double xpm = x - xp, ypm = y - yp, zpm = z - zp;
    det = 1.0 / det;
    xi = (c11 * xpm + c21 * ypm + c31 * zpm) * det;
    yi = (c12 * xpm + c22 * ypm + c32 * zpm) * det;
    zi = (c13 * xpm + c23 * ypm + c33 * zpm) * det;
  }
}

// InterpData::coordHexa (InterpData.cpp:56-134) + the centre tetrahedron 14,11,12,10 of getCellJump /
// coeffInterpHexa (:383-399, :214-222): the 8 vertices go to scratch, the centre point 14 stays in registers.
CX_D void cell_load_jump(const ZoneView& Z, int ind, const Scratch& S, double x, double y, double z, int& ic, int& jc,
                         int& kc, double& cx, double& cy, double& cz, double& xi, double& yi, double& zi)
{
  const int ni = Z.ni, nij = Z.ni * Z.nj;
  kc = ind / nij + 1;
  int kcnij = kc * nij, kcmnij = (kc - 1) * nij;
  jc = (ind - kcmnij) / ni + 1;
  int jcni = jc * ni, jcmni = (jc - 1) * ni;
  ic = ind - jcmni - kcmnij + 1;
  double vx[8], vy[8], vz[8];
  {
    const int i0 = ind, i1 = ic + jcmni + kcmnij, i2 = ic - 1 + jcni + kcmnij, i3 = ic + jcni + kcmnij;
    vx[0] = Z.x[i0]; vy[0] = Z.y[i0]; vz[0] = Z.z[i0];
    vx[1] = Z.x[i1]; vy[1] = Z.y[i1]; vz[1] = Z.z[i1];
    vx[2] = Z.x[i2]; vy[2] = Z.y[i2]; vz[2] = Z.z[i2];
    vx[3] = Z.x[i3]; vy[3] = Z.y[i3]; vz[3] = Z.z[i3];
    if (Z.nk > 1)
    {
      const int i4 = ic - 1 + jcmni + kcnij, i5 = ic + jcmni + kcnij, i6 = ic - 1 + jcni + kcnij, i7 = ic + jcni + kcnij;
      vx[4] = Z.x[i4]; vy[4] = Z.y[i4]; vz[4] = Z.z[i4];
      vx[5] = Z.x[i5]; vy[5] = Z.y[i5]; vz[5] = Z.z[i5];
      vx[6] = Z.x[i6]; vy[6] = Z.y[i6]; vz[6] = Z.z[i6];
      vx[7] = Z.x[i7]; vy[7] = Z.y[i7]; vz[7] = Z.z[i7];
    }
    else
    {
#pragma unroll
      for (int v = 0; v < 4; v++) { vx[v + 4] = vx[v]; vy[v + 4] = vy[v]; vz[v + 4] = vz[v] + 1.; }
    }
  }
#pragma unroll
  for (int v = 0; v < 8; v++)
  {
    S.vtx[v * S.ss] = vx[v]; S.vtx[(8 + v) * S.ss] = vy[v]; S.vtx[(16 + v) * S.ss] = vz[v];
  }
#define CX_FCR(a, b, c, d, ox, oy, oz)            \
  const double ox = 0.25 * (vx[a] + vx[b] + vx[c] + vx[d]); \
  const double oy = 0.25 * (vy[a] + vy[b] + vy[c] + vy[d]); \
  const double oz = 0.25 * (vz[a] + vz[b] + vz[c] + vz[d]);
  CX_FCR(0, 1, 2, 3, x8, y8, z8)
  CX_FCR(4, 5, 6, 7, x10, y10, z10)
  CX_FCR(1, 3, 5, 7, x11, y11, z11)
  CX_FCR(2, 3, 6, 7, x12, y12, z12)
#undef CX_FCR
  cx = 0.5 * (x8 + x10); cy = 0.5 * (y8 + y10); cz = 0.5 * (z8 + z10);
  tetra_solve_p(x, y, z, cx, cy, cz, x11, y11, z11, x12, y12, z12, x10, y10, z10, xi, yi, zi);
}

// tetrahedron (14, q, r, s): q a face centre (8..13), r and s vertices (0..7), all read from scratch
CX_D void tetra_solve_s(double x, double y, double z, const Scratch& S, double cx, double cy, double cz, int q, int r,
                        int s, double& xi, double& yi, double& zi)
{
  // vertices of face q in the summation order of coordHexa, 4 bits each: 8:(0,1,2,3) 9:(0,2,4,6) 10:(4,5,6,7)
  // 11:(1,3,5,7) 12:(2,3,6,7) 13:(0,1,4,5)
  const unsigned long long lo = 0x7531765464203210ull;
  const unsigned hi = 0x54107632u;
  const unsigned f = (q < 12) ? (unsigned)(lo >> (16 * (q - 8))) : (hi >> (16 * (q - 12)));
  const int a = f & 7, b = (f >> 4) & 7, c = (f >> 8) & 7, d = (f >> 12) & 7;
  const double* vx = S.vtx; const double* vy = S.vtx + 8 * S.ss; const double* vz = S.vtx + 16 * S.ss;
  const int ss = S.ss;
  const double xq = 0.25 * (vx[a * ss] + vx[b * ss] + vx[c * ss] + vx[d * ss]);
  const double yq = 0.25 * (vy[a * ss] + vy[b * ss] + vy[c * ss] + vy[d * ss]);
  const double zq = 0.25 * (vz[a * ss] + vz[b * ss] + vz[c * ss] + vz[d * ss]);
  tetra_solve_p(x, y, z, cx, cy, cz, xq, yq, zq, vx[r * ss], vy[r * ss], vz[r * ss], vx[s * ss], vy[s * ss], vz[s * ss],
                xi, yi, zi);
}

// tetra (xi,yi,zi) -> 8 hexa weights with predicated adds in the order of tetra_to_hexa
CX_D void tetra_to_hexa_s(double xi, double yi, double zi, int indq, int indr, int inds, double* cf)
{
  const double cf0 = 0.125 * (1.0 - xi - yi - zi);
  const double tl = xi * 0.25;
  // vertices of face indq (T_indfc rows) as bit masks: 8: 0x0f, 9: 0x55, 10: 0xf0, 11: 0xaa, 12: 0xcc, 13: 0x33
  const unsigned fm = (unsigned)((0x33ccaaf0550full >> (8 * (indq - 8))) & 0xffu);
#pragma unroll
  for (int i = 0; i < 8; i++)
  {
    double t = cf0;
    if ((fm >> i) & 1u) t += tl;
    if (i == indr) t += yi;
    if (i == inds) t += zi;
    cf[i] = t;
  }
}

CX_D bool hexa_from_corner_s(double x, double y, double z, const Scratch& S, double cx, double cy, double cz, int isomm,
                             double xi, double yi, double zi, double* cf)
{
  if (!(isomm & 1)) xi = -xi;
  if (!(isomm & 2)) yi = -yi;
  if (!(isomm & 4)) zi = -zi;
  int is1 = trunc_int((1.0 + fsign(xi - yi)) * 0.5);
  int is2 = trunc_int((1.0 + fsign(yi - zi)) * 0.5);
  int is3 = trunc_int((1.0 + fsign(zi - xi)) * 0.5);
  int itri = T_indtr[isomm + 8 * is1 + 16 * is2 + 32 * is3];
  int indr = T_indtd[0 + itri * 4], inds = T_indtd[1 + itri * 4], indq = T_indtd[2 + itri * 4];
  tetra_solve_s(x, y, z, S, cx, cy, cz, indq, indr, inds, xi, yi, zi);
  if (tetra_accept(xi, yi, zi)) { tetra_to_hexa_s(xi, yi, zi, indq, indr, inds, cf); return true; }
  for (int ibsom = 0; ibsom < 4; ibsom++)
  {
    int isom = T_indss[ibsom + 4 * isomm];
    indr = isom;
    for (int its = 0; its < 3; its++)
    {
      inds = T_neighbour[its + isom * 3];
      for (int itq = 0; itq < 2; itq++)
      {
        indq = T_center[itq + its + isom * 4];
        tetra_solve_s(x, y, z, S, cx, cy, cz, indq, indr, inds, xi, yi, zi);
        if (tetra_accept(xi, yi, zi)) { tetra_to_hexa_s(xi, yi, zi, indq, indr, inds, cf); return true; }
      }
    }
  }
  return false;
}

// InterpAdt::searchInterpolationCellStruct (InterpAdt.cpp:1252-1400), scratch version of search_interp_cell
CX_D int search_interp_cell_s(const ZoneView& Z, double x, double y, double z, int& ic, int& jc, int& kc, double* cf,
                              const Scratch& S, SearchStats* st)
{
  if (zone_reject(Z, x, y, z, 0.0)) return 0;
  AdtQueryS q;
  query_init_s(q, Z, S, x, y, z, 0.0);
  int nid, nvisit = 0;
  int first = query_next_s(q, Z, S, &nid, &nvisit);
  if (st) { st->nvisit += nvisit; }
  if (first < 0) return 0;
  const int ni = Z.ni, nj = Z.nj, nk = Z.nk, nij = ni * nj;
  bool JUMP = false;
  int ind = first, isomm = 0;
  double xi = 0, yi = 0, zi = 0, cx = 0, cy = 0, cz = 0;
  int erased = -1;
  for (int i = 0; i < 8; i++)
  {
    cell_load_jump(Z, ind, S, x, y, z, ic, jc, kc, cx, cy, cz, xi, yi, zi);
    if (!(fabs(xi) <= 1.0 + CX_EPS_GEOM && fabs(yi) <= 1.0 + CX_EPS_GEOM && fabs(zi) <= 1.0 + CX_EPS_GEOM))
    {
      if (i == 7) break;
      int is = imin(8, trunc_int((xi + fsign(xi)) * 0.5)); is = imax(-8, is);
      int js = imin(8, trunc_int((yi + fsign(yi)) * 0.5)); js = imax(-8, js);
      int ks = imin(8, trunc_int((zi + fsign(zi)) * 0.5)); ks = imax(-8, ks);
      ind = ind + is + js * ni + ks * nij;
      kc = ind / nij;
      int kcnij = kc * nij;
      jc = (ind - kcnij) / ni;
      int jcni = jc * ni;
      ic = ind - jcni - kcnij;
      if (ic < 0 || ic > ni - 2 || jc < 0 || jc > nj - 2 || kc < 0 || kc > nk - 2) break;
    }
    else
    {
      isomm = (xi >= 0 ? 1 : 0) + (yi >= 0 ? 2 : 0) + (zi >= 0 ? 4 : 0);
      int cid = (ic - 1) + (ni - 1) * ((jc - 1) + (nj - 1) * (kc - 1));
      JUMP = (ind == first) ? true : query_contains_s(q, Z, cid);
      break;
    }
  }
  if (JUMP)
  {
    if (hexa_from_corner_s(x, y, z, S, cx, cy, cz, isomm, xi, yi, zi, cf)) return 1;
    erased = ind;
  }
  // ordered fallback over the whole candidate list (InterpAdt.cpp:1378-1397)
  query_init_s(q, Z, S, x, y, z, 0.0);
  nvisit = 0;
  int c;
  while ((c = query_next_s(q, Z, S, &nid, &nvisit)) >= 0)
  {
    if (c == erased) continue;
    cell_load_jump(Z, c, S, x, y, z, ic, jc, kc, cx, cy, cz, xi, yi, zi);
    isomm = 0;                                                 // coeffInterpHexa (InterpData.cpp:214-241)
    if (fabs(xi) <= 1.0 + CX_EPS_GEOM && fabs(yi) <= 1.0 + CX_EPS_GEOM && fabs(zi) <= 1.0 + CX_EPS_GEOM)
      isomm = (xi >= 0 ? 1 : 0) + (yi >= 0 ? 2 : 0) + (zi >= 0 ? 4 : 0);
    if (hexa_from_corner_s(x, y, z, S, cx, cy, cz, isomm, xi, yi, zi, cf)) { if (st) st->nvisit += nvisit; return 1; }
  }
  if (st) st->nvisit += nvisit;
  return 0;
}

// ---------------------------------------------------------------------------
// Per-receptor drivers (bodies of the search kernels)
// ---------------------------------------------------------------------------
struct RcvOut { int noblk, type, dind; double cf[8]; };

// commonCheckCoefs.h, type 2 / type 22 branches: exactly one |coef| > 1e-10 -> type 1
CX_D void collapse_type2(const ZoneView& Z, int& type, int& dind, double* cf)
{
  const int ni = Z.ni, ninj = Z.ni * Z.nj;
  int nb = 0, indNonZero = -1;
  double coefNonZero = 0.;
  if (type == 22)
  {
    for (int jj = 0; jj < 2; jj++) for (int ii = 0; ii < 2; ii++)
    {
      double c = cf[ii + jj * 2];
      if (fabs(c) > 1.e-10) { nb++; indNonZero = dind + ii + jj * ni; coefNonZero = c; }
    }
    if (nb == 1) { type = 1; dind = indNonZero; cf[0] = coefNonZero; }
    return;
  }
#pragma unroll
  for (int q = 0; q < 8; q++)
  {
    const int ii = q & 1, jj = (q >> 1) & 1, kk = q >> 2;
    double c = cf[q];
    if (fabs(c) > 1.e-10) { nb++; indNonZero = dind + ii + jj * ni + kk * ninj; coefNonZero = c; }
  }
  if (nb == 1) { type = 1; dind = indNonZero; cf[0] = coefNonZero; }
}

// K_INTERP::getInterpolationCell for O2CF (Interp.cpp:81-165 +
// Interp2.cpp:252-299 + commonTypesForExtrapAndInterp.h + commonCheckCoefs.h):
// zones scanned in list order; `vol < best + 1e-8` lets a later zone within
// 1e-8 replace an earlier one.
CX_D void interp_o2_receptor(double x, double y, double z, int nz, const ZoneView* zones, int nature, int penalty,
                             RcvOut& o, SearchStats* st)
{
  double best = CX_MAXF;
  o.noblk = 0; o.type = 0; o.dind = -1;
  double tcf[8];
  for (int i = 0; i < 8; i++) { o.cf[i] = 0.; tcf[i] = 0.; }
  for (int noz = 0; noz < nz; noz++)
  {
    const ZoneView& Z = zones[noz];
    int ic, jc, kc;
    if (Z.nk == 1)
    {
      // 2-D donor (Interp2.cpp:204-207, :300-340): z must be the zone's plane; coefficients folded onto the 4 nodes
      if (fabs(z - Z.bbox[2]) > CX_EPS_GEOM) continue;
      int found2 = search_interp_cell(Z, x, y, z, ic, jc, kc, tcf, st);
      if (found2 < 1) continue;
      ic -= 1; jc -= 1;
      int isBorder2 = (ic == 0 || ic == Z.ni - 2 || jc == 0 || jc == Z.nj - 2);
      int tind2 = ic + jc * Z.ni;
      for (int i = 0; i < 4; i++) { tcf[i] += tcf[i + 4]; tcf[i + 4] = 0.; }
      if (!celln_valid_o2_2d(Z, ic, jc, tcf, nature)) continue;
      double vol2 = cell_area_2d(Z, tind2);
      if (penalty == 1 && isBorder2 == 1) vol2 += 1.e3;
      if (vol2 < best + CX_EPS_GEOM)
      {
        best = vol2; o.type = 22; o.dind = tind2; o.noblk = noz + 1;
        for (int i = 0; i < 4; i++) { o.cf[i] = tcf[i] + tcf[i + 4]; o.cf[i + 4] = 0.; }
      }
      continue;
    }
    int found = (Z.topo == 0) ? cart_search_o2(Z, x, y, z, ic, jc, kc, tcf)
                              : search_interp_cell(Z, x, y, z, ic, jc, kc, tcf, st);
    if (found < 1) continue;
    ic -= 1; jc -= 1; kc -= 1;
    int isBorder = (ic == 0 || ic == Z.ni - 2 || jc == 0 || jc == Z.nj - 2 || kc == 0 || kc == Z.nk - 2);
    int tind = ic + jc * Z.ni + kc * Z.ni * Z.nj;
    if (!celln_valid_o2(Z, ic, jc, kc, tcf, nature)) continue;
    double vol = cell_volume(Z, tind);
    if (penalty == 1 && isBorder == 1) vol += 1.e3;
    if (vol < best + CX_EPS_GEOM)
    {
      best = vol; o.type = 2; o.dind = tind; o.noblk = noz + 1;
      for (int i = 0; i < 8; i++) o.cf[i] = tcf[i];
    }
  }
  if (o.noblk > 0) collapse_type2(zones[o.noblk - 1], o.type, o.dind, o.cf);
}

// Double wall (K_INTERP::getInterpolationCellDW / getExtrapolationCellDW, Interp.cpp:176-255, :362-): the receptor
// seen from donor zone noz is (x[r], y[r], z[r]) of entry noz of this table -- the point projected onto that zone's
// walls by the caller; everything else is the ordinary search.  nullptr: one point for every zone.
struct RcvPerZone { const double* x; const double* y; const double* z; };

// interp_o2_receptor with the search state in scratch (the version the kernels run)
CX_D void interp_o2_receptor_s(double x, double y, double z, int nz, const ZoneView* zones, int nature, int penalty,
                               RcvOut& o, const Scratch& S, SearchStats* st, const RcvPerZone* pz = nullptr, int r = 0)
{
  double best = CX_MAXF;
  o.noblk = 0; o.type = 0; o.dind = -1;
  double tcf[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { o.cf[i] = 0.; tcf[i] = 0.; }
  for (int noz = 0; noz < nz; noz++)
  {
    const ZoneView& Z = zones[noz];
    if (pz) { x = pz[noz].x[r]; y = pz[noz].y[r]; z = pz[noz].z[r]; }
    int ic, jc, kc;
    if (Z.nk == 1)
    {
      if (fabs(z - Z.bbox[2]) > CX_EPS_GEOM) continue;
      int found2 = search_interp_cell_s(Z, x, y, z, ic, jc, kc, tcf, S, st);
      if (found2 < 1) continue;
      ic -= 1; jc -= 1;
      int isBorder2 = (ic == 0 || ic == Z.ni - 2 || jc == 0 || jc == Z.nj - 2);
      int tind2 = ic + jc * Z.ni;
#pragma unroll
      for (int i = 0; i < 4; i++) { tcf[i] += tcf[i + 4]; tcf[i + 4] = 0.; }
      if (!celln_valid_o2_2d(Z, ic, jc, tcf, nature)) continue;
      double vol2 = cell_area_2d(Z, tind2);
      if (penalty == 1 && isBorder2 == 1) vol2 += 1.e3;
      if (vol2 < best + CX_EPS_GEOM)
      {
        best = vol2; o.type = 22; o.dind = tind2; o.noblk = noz + 1;
#pragma unroll
        for (int i = 0; i < 4; i++) { o.cf[i] = tcf[i] + tcf[i + 4]; o.cf[i + 4] = 0.; }
      }
      continue;
    }
    int found = (Z.topo == 0) ? cart_search_o2(Z, x, y, z, ic, jc, kc, tcf)
                              : search_interp_cell_s(Z, x, y, z, ic, jc, kc, tcf, S, st);
    if (found < 1) continue;
    ic -= 1; jc -= 1; kc -= 1;
    int isBorder = (ic == 0 || ic == Z.ni - 2 || jc == 0 || jc == Z.nj - 2 || kc == 0 || kc == Z.nk - 2);
    int tind = ic + jc * Z.ni + kc * Z.ni * Z.nj;
    if (!celln_valid_o2(Z, ic, jc, kc, tcf, nature)) continue;
    double vol = cell_volume(Z, tind);
    if (penalty == 1 && isBorder == 1) vol += 1.e3;
    if (vol < best + CX_EPS_GEOM)
    {
      best = vol; o.type = 2; o.dind = tind; o.noblk = noz + 1;
#pragma unroll
      for (int i = 0; i < 8; i++) o.cf[i] = tcf[i];
    }
  }
  if (o.noblk > 0) collapse_type2(zones[o.noblk - 1], o.type, o.dind, o.cf);
}

// K_INTERP::getExtrapolationCell (Interp.cpp:282-360, Interp2.cpp:84-137):
// constraint=40, extrapOrder=1, order forced to O2CF.
CX_D void extrap_receptor(double x, double y, double z, int nz, const ZoneView* zones, int nature, int penalty,
                          RcvOut& o, SearchStats* st)
{
  double best = CX_MAXF, sumCoefMin = CX_MAXF;
  o.noblk = 0; o.type = 0; o.dind = -1;
  double tcf[8];
  for (int i = 0; i < 8; i++) { o.cf[i] = 0.; tcf[i] = 0.; }
  const double constraint = 40.;
  for (int noz = 0; noz < nz; noz++)
  {
    const ZoneView& Z = zones[noz];
    int ic, jc, kc;
    int found = search_extrap_cell(Z, x, y, z, ic, jc, kc, tcf, nature, constraint, st);
    if (found < 1) continue;
    ic -= 1; jc -= 1; kc -= 1;
    if (Z.nk == 1)
    {
      // Interp2.cpp:125-136 + commonTypesForExtrapAndInterp.h:72-105
      int isBorder2 = (ic == 0 || ic == Z.ni - 2 || jc == 0 || jc == Z.nj - 2);
      int tind2 = ic + jc * Z.ni;
      for (int i = 0; i < 4; i++) { tcf[i] += tcf[i + 4]; tcf[i + 4] = 0.; }
      double vol2 = cell_area_2d(Z, tind2);
      if (penalty == 1 && isBorder2 == 1) vol2 += 1.e3;
      if (vol2 < best + CX_EPS_GEOM)
      {
        double sumCf = 0.;
        for (int i = 0; i < 4; i++) sumCf += fabs(tcf[i]);
        if (sumCf < sumCoefMin)
        {
          sumCoefMin = sumCf;
          best = vol2; o.type = 22; o.dind = tind2; o.noblk = noz + 1;
          for (int i = 0; i < 4; i++) { o.cf[i] = tcf[i] + tcf[i + 4]; o.cf[i + 4] = 0.; }
        }
      }
      continue;
    }
    int isBorder = (ic == 0 || ic == Z.ni - 2 || jc == 0 || jc == Z.nj - 2 || kc == 0 || kc == Z.nk - 2);
    int tind = ic + jc * Z.ni + kc * Z.ni * Z.nj;
    double vol = cell_volume(Z, tind);
    if (penalty == 1 && isBorder == 1) vol += 1.e3;
    if (vol < best + CX_EPS_GEOM)
    {
      double sumCf = 0.;
      for (int i = 0; i < 8; i++) sumCf += fabs(tcf[i]);
      if (sumCf < sumCoefMin)
      {
        sumCoefMin = sumCf;
        best = vol; o.type = 2; o.dind = tind; o.noblk = noz + 1;
        for (int i = 0; i < 8; i++) o.cf[i] = tcf[i];
      }
    }
  }
  if (o.noblk > 0) collapse_type2(zones[o.noblk - 1], o.type, o.dind, o.cf);
}

// ---------------------------------------------------------------------------
// Level-synchronous ADT build, per-cell steps (see cx_adt.cu)
// ---------------------------------------------------------------------------
struct ZoneBox { double lo[6], hi[6]; };   // region of the root for the 6 keys
struct BuildView {
  double* keys; double* lo; double* hi;    // [6][ncells]
  int* cur; unsigned char* br; AdtNode* nodes;
  // children slots of the build, child[2*p] = left, child[2*p+1] = right (CX_NULL = free): a compact 8-byte-per-node
  // array that the atomics and the winner look-ups of every level hit instead of the 128-byte nodes (133 MB instead of
  // 2.1 GB for 16.6 M cells: it stays in L2); copied into the nodes once the levels are done (adt_set_children)
  int* child;
  int* depth;                              // [ncells] depth at which each cell was placed
  int ncells;
  ZoneBox zb;
};

CX_D void freeze_node(AdtNode& nd, int depth, double lo, double hi)
{
  // thresholds the query compares against at this node (InterpAdt.cpp:772-784)
  double mid = 0.5 * (hi + lo);
  double h2 = 2 * mid - lo;
  nd.mid = mid;
  nd.rlo = 0.5 * (h2 + lo);
  nd.dim = depth % 6;
}

// cell c (currently below parent cur[c]) evaluates level L of the insertion
// descent (InterpAdt.cpp:532-641): returns the parent and the branch taken.
CX_D void adt_claim(const BuildView& B, int L, int c, int& p, int& right)
{
  p = B.cur[c];
  int d = L % 6;
  size_t o = (size_t)d * B.ncells + c;
  double l, h;
  if (L < 6) { l = B.zb.lo[d]; h = B.zb.hi[d]; }
  else { l = B.lo[o]; h = B.hi[o]; }
  double key = B.keys[o];
  double mid = 0.5 * (h + l);               // InterpAdt.cpp:537
  if (key <= mid) { h = mid; right = 0; }   // :539-542
  else { h = 2 * mid - l; l = 0.5 * (h + l); right = 1; }   // :545-547
  B.lo[o] = l; B.hi[o] = h;
  B.br[c] = (unsigned char)right;
}

// after all claims of level L: true when c won its slot (placed at depth L+1)
CX_D bool adt_resolve(const BuildView& B, int L, int c)
{
  int p = B.cur[c];
  int w = B.child[2 * (size_t)p + (B.br[c] ? 1 : 0)];
  if (w == c)
  {
    int depth = L + 1;
    int d = depth % 6;
    double l, h;
    if (depth < 6) { l = B.zb.lo[d]; h = B.zb.hi[d]; }
    else { size_t o = (size_t)d * B.ncells + c; l = B.lo[o]; h = B.hi[o]; }
    freeze_node(B.nodes[c], depth, l, h);
    B.depth[c] = depth;
    return true;
  }
  B.cur[c] = w;
  return false;
}

CX_D void adt_set_children(const BuildView& B, int c)
{
  // a free slot holds CX_NULL or the byte-fill 0x7f7f7f7f of the device build (both above any cell id)
  const int l = B.child[2 * (size_t)c], r = B.child[2 * (size_t)c + 1];
  B.nodes[c].left = l >= 0x7f7f7f7f ? CX_NULL : l; B.nodes[c].right = r >= 0x7f7f7f7f ? CX_NULL : r;
}

// bbox keys of cell c (InterpAdt.cpp:310-388)
// nk == 1 (InterpAdt.cpp:391-431): the 4 nodes of the quad give x,y; the z keys are the zone's plane and plane + 1
// (zmin2d, zmax2d = _zmin, _zmax of the InterpAdt)
CX_D void adt_cell_keys(int ni, int nj, int nk, const double* x, const double* y, const double* z, int c,
                        int ncells, double* keys, AdtNode* nodes, const double* zbbox, double zmin2d = 0., double zmax2d = 0.)
{
  int nic = ni - 1, njc = nj - 1;
  int k = c / (nic * njc);
  int j = (c - k * nic * njc) / nic;
  int i = c - j * nic - k * nic * njc;
  int nij = ni * nj;
  int ind = i + ni * j + nij * k;
  double xmin = 1.e20, ymin = 1.e20, zmin = 1.e20, xmax = -1.e20, ymax = -1.e20, zmax = -1.e20;
  if (nk > 1)
  {
    const int off[8] = {0, 1, ni, nij, ni + 1, nij + 1, ni + nij, ni + nij + 1};
    for (int v = 0; v < 8; v++)
    {
      int n = ind + off[v];
      double xv = x[n], yv = y[n], zv = z[n];
      xmin = xv < xmin ? xv : xmin; ymin = yv < ymin ? yv : ymin; zmin = zv < zmin ? zv : zmin;
      xmax = xv > xmax ? xv : xmax; ymax = yv > ymax ? yv : ymax; zmax = zv > zmax ? zv : zmax;
    }
  }
  else
  {
    const int off[4] = {0, 1, ni, ni + 1};
    for (int v = 0; v < 4; v++)
    {
      int n = ind + off[v];
      double xv = x[n], yv = y[n];
      xmin = xv < xmin ? xv : xmin; ymin = yv < ymin ? yv : ymin;
      xmax = xv > xmax ? xv : xmax; ymax = yv > ymax ? yv : ymax;
    }
    zmin = zmin2d; zmax = zmax2d;
  }
  size_t n = ncells;
  keys[c] = xmin; keys[n + c] = ymin; keys[2 * n + c] = zmin;
  keys[3 * n + c] = xmax; keys[4 * n + c] = ymax; keys[5 * n + c] = zmax;
  AdtNode nd;
  nd.bb[0] = xmin; nd.bb[1] = ymin; nd.bb[2] = zmin; nd.bb[3] = xmax; nd.bb[4] = ymax; nd.bb[5] = zmax;
  for (int d = 0; d < 3; d++)
  {
    double org, scl;
    quant_params(zbbox, d, org, scl);
    nd.qs[d] = quant_lo(nd.bb[d] - CX_EPS_GEOM, org, scl); nd.qs[d + 3] = quant_hi(nd.bb[d + 3] + CX_EPS_GEOM, org, scl);
    nd.qb[d] = nd.qs[d]; nd.qb[d + 3] = nd.qs[d + 3];
    nd.ql[d] = 65535; nd.ql[d + 3] = 0; nd.qr[d] = 65535; nd.qr[d + 3] = 0;      // no child yet: empty boxes
  }
  nd.mid = 0.; nd.rlo = 0.; nd.left = CX_NULL; nd.right = CX_NULL; nd.dim = 0; nd.ind = ind;
  nodes[c] = nd;
}

// bottom-up pass (deepest level first): subtree box = own box U children's boxes; the children's boxes are also
// copied into the parent (ql, qr)
CX_D void adt_subtree_box(AdtNode* nodes, int c)
{
  AdtNode& n = nodes[c];
  for (int side = 0; side < 2; side++)
  {
    int ch = side ? n.right : n.left;
    if (ch == CX_NULL) continue;
    const AdtNode& m = nodes[ch];
    unsigned short* dst = side ? n.qr : n.ql;
    for (int d = 0; d < 3; d++)
    {
      dst[d] = m.qs[d]; dst[d + 3] = m.qs[d + 3];
      if (m.qs[d] < n.qs[d]) n.qs[d] = m.qs[d];
      if (m.qs[d + 3] > n.qs[d + 3]) n.qs[d + 3] = m.qs[d + 3];
    }
  }
}

}  // namespace cx
