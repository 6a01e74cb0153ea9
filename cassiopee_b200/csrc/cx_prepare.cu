// cx_prepare.cu — the cellN preparers either side of the path, on the device.
//
// Replaces, for structured zones,
//   connector.applyBCOverlapStruct            Connector/Connector/applyBCOverlaps.cpp:27-125
//       cellN = val (2) in the `depth` layers next to a BCOverlap window, except where cellN == 0
//   connector._getOversetHolesInterpNodes /
//   connector._getOversetHolesInterpCellCenters (searchMaskInterpolatedCellsStructOpt)
//                                             Connector/Connector/getInterpolatedPoints.cpp:370-650
//       a point with cellN == 1 becomes 2 when a point of its stencil has cellN == 0; dir 0: cross (+-1..depth along
//       each axis, indices clamped to the zone), dir 1: star (the 8 / 26 directions, a component that leaves the
//       zone falls back to the point's own index); evaluated on the input values (the reference works on a copy).
// Both exist as device-resident entry points (the cellN of a zone hook) and as host-pointer entry points
// (upload, one launch, download).
#include "cx_internal.h"
#include "../../include/cassiopee_b200.h"
#include <algorithm>

namespace {

__device__ __forceinline__ bool is_zero(double v) { return v >= -1.e-15 && v <= 1.e-15; }      // K_FUNC::fEqualZero
__device__ __forceinline__ bool is_one(double v) { v -= 1.; return v >= -1.e-15 && v <= 1.e-15; } // K_FUNC::fEqual(v,1)

// depth < 0 (Connector.py:176-182): the routine runs on 1 - cellN + (cellN > 1.5)*3 (blanked and computed points
// exchanged, 2 kept) and its result goes through the same map again; the map is applied on the fly, SW = 1
__device__ __forceinline__ double swap01(double c) { return 1. - c + (c > 1.5 ? 1. : 0.) * 3.; }

template <int SW>
__device__ __forceinline__ double ldc(const double* __restrict__ cellN, long long i) { return SW ? swap01(cellN[i]) : cellN[i]; }

template <int SW>
__global__ void k_hole_fringe(int im, int jm, int km, int depth, int dir, const double* __restrict__ cellN_,
                              double* __restrict__ out)
{
  struct V { const double* p; __device__ double operator[](long long i) const { return ldc<SW>(p, i); } } cellN{cellN_};
  const long long n = (long long)im * jm * km;
  long long ind = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ind >= n) return;
  const double c = cellN[ind];
  double res = c;
  if (is_one(c))
  {
    const int imjm = im * jm;
    const int k = (int)(ind / imjm), j = (int)((ind - (long long)k * imjm) / im), i = (int)(ind - (long long)k * imjm - (long long)j * im);
    bool hit = false;
    if (dir == 0)
    {
      for (int d = 1; d <= depth && !hit; d++)
      {
        const int il = max(0, i - d), ih = min(i + d, im - 1), jl = max(0, j - d), jh = min(j + d, jm - 1);
        const int kl = max(0, k - d), kh = min(k + d, km - 1);
        hit = is_zero(cellN[il + (long long)j * im + (long long)k * imjm]) || is_zero(cellN[ih + (long long)j * im + (long long)k * imjm]) ||
              is_zero(cellN[i + (long long)jl * im + (long long)k * imjm]) || is_zero(cellN[i + (long long)jh * im + (long long)k * imjm]);
        if (!hit && km > 1)
          hit = is_zero(cellN[i + (long long)j * im + (long long)kl * imjm]) || is_zero(cellN[i + (long long)j * im + (long long)kh * imjm]);
      }
    }
    else
    {
      // star: (i + id*d, j + jd*d, k + kd*d), id, jd, kd in {-1,0,1} not all zero; a component out of the zone -> own index
      for (int d = 1; d <= depth && !hit; d++)
        for (int kd = (km > 1 ? -1 : 0); kd <= (km > 1 ? 1 : 0) && !hit; kd++)
        {
          int kk = k + kd * d; if (kk < 0 || kk > km - 1) kk = k;
          for (int jd = -1; jd <= 1 && !hit; jd++)
          {
            int jj = j + jd * d; if (jj < 0 || jj > jm - 1) jj = j;
            for (int id = -1; id <= 1; id++)
            {
              if (id == 0 && jd == 0 && kd == 0) continue;
              int ii = i + id * d; if (ii < 0 || ii > im - 1) ii = i;
              if (is_zero(cellN[ii + (long long)jj * im + (long long)kk * imjm])) { hit = true; break; }
            }
          }
        }
    }
    if (hit) res = 2.;
  }
  out[ind] = SW ? swap01(res) : res;
}

// window [i0,i1] x [j0,j1] x [k0,k1], 0-based inclusive
__global__ void k_apply_window(int im, int jm, int i0, int i1, int j0, int j1, int k0, int k1, double val, double* __restrict__ cellN)
{
  const int ni = i1 - i0 + 1, nj = j1 - j0 + 1, nk = k1 - k0 + 1;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)ni * nj * nk) return;
  const int k = (int)(t / ((long long)ni * nj)), j = (int)((t - (long long)k * ni * nj) / ni), i = (int)(t - (long long)k * ni * nj - (long long)j * ni);
  const long long ind = (i0 + i) + (long long)(j0 + j) * im + (long long)(k0 + k) * im * jm;
  if (cellN[ind] != 0.) cellN[ind] = val;
}

// the window arithmetic of applyBCOverlaps.cpp:61-96 (1-based in, 1-based inclusive out); < 0: invalid window
int overlap_window(int im, int jm, int km, int depth, int loc, int& imin, int& jmin, int& kmin, int& imax, int& jmax, int& kmax)
{
  const int shift = loc == 0 ? 1 : 0;       // loc 0 = nodes
  if (loc == 0) { if (imin < 1 || imax > im || jmin < 1 || jmax > jm || kmin < 1 || kmax > km) return -1; }
  else { if (imin < 1 || imax > im + 1 || jmin < 1 || jmax > jm + 1 || kmin < 1 || kmax > km + 1) return -1; }
  auto rng = [&](int& a, int& b, int n) {
    if (a == b)
    {
      if (a == 1) { a = 1; b = std::min(depth, n); }
      else { a = std::max(a - depth + shift, 1); b = b - 1 + shift; }
    }
    else b = b - 1 + shift;
  };
  rng(imin, imax, im); rng(jmin, jmax, jm); rng(kmin, kmax, km);
  return 0;
}

}  // namespace

extern "C" {

int cx_hole_fringe_dev(cx_ctx* c, int im, int jm, int km, int depth, int dir, const double* d_in, double* d_out)
{
  if (dir != 0 && dir != 1) { cx_set_error("setHoleInterpolatedPoints: dir=%d (diamond / octahedron stencils) has no device path", dir); return -4; }
  if (im < 1 || jm < 1 || km < 1) { cx_set_error("setHoleInterpolatedPoints: invalid dimensions"); return -1; }
  CX_CUDA(cudaSetDevice(c->device));
  const long long n = (long long)im * jm * km;
  if (depth >= 0) k_hole_fringe<0><<<cx_grid(n, 256), 256, 0, c->stream>>>(im, jm, km, depth, dir, d_in, d_out);
  else k_hole_fringe<1><<<cx_grid(n, 256), 256, 0, c->stream>>>(im, jm, km, -depth, dir, d_in, d_out);
  c->launches++;
  return 0;
}

int cx_hole_fringe(cx_ctx* c, int im, int jm, int km, int depth, int dir, double* cellN)
{
  if (dir != 0 && dir != 1) { cx_set_error("setHoleInterpolatedPoints: dir=%d (diamond / octahedron stencils) has no device path", dir); return -4; }
  CX_CUDA(cudaSetDevice(c->device));
  const size_t nb = sizeof(double) * (size_t)im * jm * km;
  double *d_in, *d_out;
  CX_MALLOC(d_in, nb, c->stream); CX_MALLOC(d_out, nb, c->stream);
  CX_CHECK(cx_upload(c, d_in, cellN, nb));
  int rc = cx_hole_fringe_dev(c, im, jm, km, depth, dir, d_in, d_out);
  if (rc == 0)
  {
    CX_CUDA(cudaMemcpyAsync(cellN, d_out, nb, cudaMemcpyDeviceToHost, c->stream));
    CX_CUDA(cudaStreamSynchronize(c->stream));
    CX_CUDA(cudaGetLastError());
  }
  cudaFreeAsync(d_in, c->stream); cudaFreeAsync(d_out, c->stream);
  return rc;
}

int cx_apply_bc_overlap_dev(cx_ctx* c, int im, int jm, int km, int imin, int jmin, int kmin, int imax, int jmax, int kmax,
                            int depth, int loc, double val, double* d_cellN)
{
  if (overlap_window(im, jm, km, depth, loc, imin, jmin, kmin, imax, jmax, kmax) < 0)
  { cx_set_error("applyBCOverlaps: indices of structured window are not valid."); return -1; }
  // a window that ends beyond the array (centres: imax may be im+1 before the shift) is clipped like the loop bounds are
  imax = std::min(imax, im); jmax = std::min(jmax, jm); kmax = std::min(kmax, km);
  if (imax < imin || jmax < jmin || kmax < kmin) return 0;
  CX_CUDA(cudaSetDevice(c->device));
  const long long n = (long long)(imax - imin + 1) * (jmax - jmin + 1) * (kmax - kmin + 1);
  k_apply_window<<<cx_grid(n, 256), 256, 0, c->stream>>>(im, jm, imin - 1, imax - 1, jmin - 1, jmax - 1, kmin - 1, kmax - 1, val, d_cellN);
  c->launches++;
  return 0;
}

int cx_apply_bc_overlaps(cx_ctx* c, int im, int jm, int km, int nwin, const int* win, int depth, int loc, double val,
                         double* cellN)
{
  CX_CUDA(cudaSetDevice(c->device));
  // refuse before touching anything, like the reference does window by window before its loop
  for (int w = 0; w < nwin; w++)
  {
    int a[6]; for (int q = 0; q < 6; q++) a[q] = win[6 * w + q];
    if (overlap_window(im, jm, km, depth, loc, a[0], a[1], a[2], a[3], a[4], a[5]) < 0)
    { cx_set_error("applyBCOverlaps: indices of structured window are not valid."); return -1; }
  }
  const size_t nb = sizeof(double) * (size_t)im * jm * km;
  double* d;
  CX_MALLOC(d, nb, c->stream);
  CX_CHECK(cx_upload(c, d, cellN, nb));
  int rc = 0;
  for (int w = 0; w < nwin && rc == 0; w++)
    rc = cx_apply_bc_overlap_dev(c, im, jm, km, win[6 * w], win[6 * w + 1], win[6 * w + 2], win[6 * w + 3], win[6 * w + 4],
                                 win[6 * w + 5], depth, loc, val, d);
  if (rc == 0)
  {
    CX_CUDA(cudaMemcpyAsync(cellN, d, nb, cudaMemcpyDeviceToHost, c->stream));
    CX_CUDA(cudaStreamSynchronize(c->stream));
    CX_CUDA(cudaGetLastError());
  }
  cudaFreeAsync(d, c->stream);
  return rc;
}

}  // extern "C"
