// cx_receptors.cu — receptor points of one zone, extracted on the device.
//
// Replaces, for structured zones,
//   connector.getInterpolatedPoints   Connector/Connector/getInterpolatedPoints.cpp:1675-1732
//       keep the points with |cellN - 2| <= 1e-15 (K_FUNC::fEqualZero, E_CUTOFF) in source order and record
//       their index ('indcell')
//   K_LOC::node2centerStruct          KCore/KCore/Loc/node2Center.cpp:118-146
//       centre = 0.125*(n0+n1+...+n7), summed left to right in that order — evaluated only for the selected
//       cells (Connector/Connector/OversetData.py:1246-1257 computes every centre, then filters)
// by flag -> exclusive scan -> gather, so that the coordinates handed to the donor search are already resident.
#include "cx_internal.h"
#include "../../include/cassiopee_b200.h"

struct cx_receptors {
  cx_ctx* ctx;
  int n;                 // selected points
  int npts;              // points of the zone at `loc`
  int* d_ind;            // [n] index in the zone at loc (ascending)
  double* d_x; double* d_y; double* d_z;   // [n]
};

namespace {
__global__ void k_rcv_flag(int n, const double* __restrict__ cellN, unsigned long long* __restrict__ f)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t = cellN[i] - 2.;
  f[i] = (t >= -1.e-15 && t <= 1.e-15) ? 1ull : 0ull;
}

// loc 0: the selected nodes; loc 1: centres of the selected cells (node dims ni,nj,nk)
__global__ void k_rcv_gather(int n, int loc, int ni, int nj, int nk, const unsigned long long* __restrict__ f,
                             const unsigned long long* __restrict__ pos, const double* __restrict__ x,
                             const double* __restrict__ y, const double* __restrict__ z, int* __restrict__ ind,
                             double* __restrict__ xo, double* __restrict__ yo, double* __restrict__ zo)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n || !f[c]) return;
  const size_t o = (size_t)pos[c];
  ind[o] = c;
  if (loc == 0) { xo[o] = x[c]; yo[o] = y[c]; zo[o] = z[c]; return; }
  const double* a[3] = {x, y, z};
  double* b[3] = {xo, yo, zo};
  if (nj == 1)
  {
    // curve along i (node2Center.cpp:80-95): 0.5*(n0+n1)
    for (int q = 0; q < 3; q++) b[q][o] = 0.5 * (a[q][c] + a[q][c + 1]);
    return;
  }
  const int nic = ni - 1, njc = nj - 1, nij = ni * nj;
  const int k = c / (nic * njc), j = (c - k * nic * njc) / nic, i = c - j * nic - k * nic * njc;
  const int n0 = i + j * ni + k * nij;
  if (nk == 1)
  {
    // 2-D zone (node2Center.cpp:98-117): 0.25*(n0+n1+n2+n3)
    for (int q = 0; q < 3; q++)
    {
      const double* p = a[q] + n0;
      double s = p[0] + p[1];
      s = s + p[ni];
      s = s + p[ni + 1];
      b[q][o] = 0.25 * s;
    }
    return;
  }
  for (int q = 0; q < 3; q++)
  {
    const double* p = a[q] + n0;
    double s = p[0] + p[1];
    s = s + p[ni];
    s = s + p[ni + 1];
    s = s + p[nij];
    s = s + p[nij + 1];
    s = s + p[nij + ni];
    s = s + p[nij + ni + 1];
    b[q][o] = 0.125 * s;
  }
}
}  // namespace

extern "C" {

int cx_receptors_create(cx_ctx* c, int ni, int nj, int nk, const double* x, const double* y, const double* z,
                        const double* cellN, int loc, cx_receptors** out)
{
  if (loc != 0 && loc != 1) { cx_set_error("setInterpData: invalid loc."); return -1; }
  if (ni < 1 || nj < 1 || nk < 1) { cx_set_error("getInterpolatedPoints: invalid zone dimensions"); return -1; }
  // loc 'nodes': any structured layout is just a list of points (curves and point clouds like the circle of
  // Connector/test/setInterpDataPT_t1.py included); loc 'centers': 3-D zones, 2-D zones with nk = 1 and curves
  // along i (node2Center.cpp:42-72 also knows the j / k curves and the ni = 1 / nj = 1 surfaces)
  const bool curve = ni >= 2 && nj == 1 && nk == 1;
  if (loc == 1 && !curve && (ni < 2 || nj < 2))
  { cx_set_error("receptor zones at centers must be 3D, 2D with nk=1, or curves along i"); return -4; }
  CX_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  const size_t nn = (size_t)ni * nj * nk;
  const size_t np = loc == 0 ? nn : curve ? (size_t)(ni - 1) : (size_t)(ni - 1) * (nj - 1) * (size_t)(nk > 1 ? nk - 1 : 1);
  if (nn > 2000000000ull) { cx_set_error("receptor zone too large for int32 indices"); return -1; }
  cx_receptors* R = new cx_receptors;
  R->ctx = c; R->n = 0; R->npts = (int)np; R->d_ind = nullptr; R->d_x = R->d_y = R->d_z = nullptr;
  double *dx, *dy, *dz, *dc; unsigned long long *f, *p;
  CX_MALLOC(dx, sizeof(double) * nn, s); CX_MALLOC(dy, sizeof(double) * nn, s); CX_MALLOC(dz, sizeof(double) * nn, s);
  CX_MALLOC(dc, sizeof(double) * np, s);
  CX_MALLOC(f, sizeof(unsigned long long) * np, s); CX_MALLOC(p, sizeof(unsigned long long) * np, s);
  CX_CHECK(cx_upload(c, dc, cellN, sizeof(double) * np));
  CX_CHECK(cx_upload(c, dx, x, sizeof(double) * nn));
  CX_CHECK(cx_upload(c, dy, y, sizeof(double) * nn));
  CX_CHECK(cx_upload(c, dz, z, sizeof(double) * nn));
  k_rcv_flag<<<cx_grid((long long)np, 256), 256, 0, s>>>((int)np, dc, f);
  c->launches++;
  unsigned long long tot = 0;
  int rc = cx_exclusive_scan_u64(c, f, p, (int)np, &tot);
  if (rc == 0)
  {
    R->n = (int)tot;
    const size_t m = tot ? (size_t)tot : 1;
    cudaError_t e = cx_malloc_async((void**)&R->d_ind, sizeof(int) * m, s);
    if (e == cudaSuccess) e = cx_malloc_async((void**)&R->d_x, sizeof(double) * m, s);
    if (e == cudaSuccess) e = cx_malloc_async((void**)&R->d_y, sizeof(double) * m, s);
    if (e == cudaSuccess) e = cx_malloc_async((void**)&R->d_z, sizeof(double) * m, s);
    if (e != cudaSuccess) { cx_set_error("receptor extraction: %s", cudaGetErrorString(e)); rc = -2; }
  }
  if (rc == 0 && tot > 0)
  {
    k_rcv_gather<<<cx_grid((long long)np, 256), 256, 0, s>>>((int)np, loc, ni, nj, nk, f, p, dx, dy, dz, R->d_ind, R->d_x,
                                                             R->d_y, R->d_z);
    c->launches++;
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc == 0 && e != cudaSuccess) { cx_set_error("receptor extraction: %s", cudaGetErrorString(e)); rc = -2; }
  cudaFreeAsync(dx, s); cudaFreeAsync(dy, s); cudaFreeAsync(dz, s); cudaFreeAsync(dc, s); cudaFreeAsync(f, s); cudaFreeAsync(p, s);
  if (rc != 0) { cx_receptors_destroy(R); return rc; }
  *out = R;
  return 0;
}

// The receptor points of a zone that is ALSO resident as a donor hook (a tree interpolated from itself, loc='nodes':
// the node coordinates and the node cellN of the hook are the receptor zone's own): nothing is uploaded.
int cx_receptors_from_zone(cx_ctx* c, cx_zone* zn, cx_receptors** out)
{
  if (zn->ctx != c) { cx_set_error("getInterpolatedPoints: zone and context differ"); return -1; }
  CX_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  const size_t np = (size_t)zn->npts;
  cx_receptors* R = new cx_receptors;
  R->ctx = c; R->n = 0; R->npts = (int)np; R->d_ind = nullptr; R->d_x = R->d_y = R->d_z = nullptr;
  unsigned long long *f, *p;
  CX_MALLOC(f, sizeof(unsigned long long) * np, s); CX_MALLOC(p, sizeof(unsigned long long) * np, s);
  k_rcv_flag<<<cx_grid((long long)np, 256), 256, 0, s>>>((int)np, zn->d_cellN, f);
  c->launches++;
  unsigned long long tot = 0;
  int rc = cx_exclusive_scan_u64(c, f, p, (int)np, &tot);
  if (rc == 0)
  {
    R->n = (int)tot;
    const size_t m = tot ? (size_t)tot : 1;
    cudaError_t e = cx_malloc_async((void**)&R->d_ind, sizeof(int) * m, s);
    if (e == cudaSuccess) e = cx_malloc_async((void**)&R->d_x, sizeof(double) * m, s);
    if (e == cudaSuccess) e = cx_malloc_async((void**)&R->d_y, sizeof(double) * m, s);
    if (e == cudaSuccess) e = cx_malloc_async((void**)&R->d_z, sizeof(double) * m, s);
    if (e != cudaSuccess) { cx_set_error("receptor extraction: %s", cudaGetErrorString(e)); rc = -2; }
  }
  if (rc == 0 && tot > 0)
  {
    k_rcv_gather<<<cx_grid((long long)np, 256), 256, 0, s>>>((int)np, 0, zn->ni, zn->nj, zn->nk, f, p, zn->d_x, zn->d_y, zn->d_z,
                                                             R->d_ind, R->d_x, R->d_y, R->d_z);
    c->launches++;
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc == 0 && e != cudaSuccess) { cx_set_error("receptor extraction: %s", cudaGetErrorString(e)); rc = -2; }
  cudaFreeAsync(f, s); cudaFreeAsync(p, s);
  if (rc != 0) { cx_receptors_destroy(R); return rc; }
  *out = R;
  return 0;
}

int cx_receptors_count(cx_receptors* R, int* n) { *n = R->n; return 0; }

int cx_receptors_fetch(cx_receptors* R, int* indcell, double* x, double* y, double* z)
{
  if (R->n == 0) return 0;
  cudaStream_t s = R->ctx->stream;
  const size_t n = (size_t)R->n;
  if (indcell) CX_CUDA(cudaMemcpyAsync(indcell, R->d_ind, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  if (x) CX_CUDA(cudaMemcpyAsync(x, R->d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  if (y) CX_CUDA(cudaMemcpyAsync(y, R->d_y, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  if (z) CX_CUDA(cudaMemcpyAsync(z, R->d_z, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  CX_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int cx_set_interp_data_rcv(cx_ctx* c, cx_receptors* R, int nz, cx_zone* const* zones, int order, int nature,
                           int penalty, int extrap, cx_result** out)
{
  if (R->ctx != c) { cx_set_error("setInterpData: receptors and context differ"); return -1; }
  return cx_set_interp_data_dev(c, R->n, R->d_x, R->d_y, R->d_z, nz, zones, order, nature, penalty, extrap, out);
}

int cx_receptors_destroy(cx_receptors* R)
{
  if (!R) return 0;
  cudaStream_t s = R->ctx->stream;
  void* ptrs[] = {R->d_ind, R->d_x, R->d_y, R->d_z};
  for (void* p : ptrs) if (p) cudaFreeAsync(p, s);
  delete R;
  return 0;
}

}  // extern "C"
