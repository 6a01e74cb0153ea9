// cx_search.cu — setInterpData on the device: one thread per receptor.
//
// Replaces the OpenMP receptor loop of K_CONNECTOR::setInterpData
// (Connector/Connector/setInterpData.cpp:978-1114) and everything below it:
//   K_INTERP::getInterpolationCell   KCore/KCore/Interp/Interp.cpp:81-165
//   K_INTERP::getInterpolationData   KCore/KCore/Interp/Interp2.cpp:185-342 (O2CF branch)
//   cross-zone choice                KCore/KCore/Interp/commonTypesForExtrapAndInterp.h:1-36
//   coincident collapse (type 1)     KCore/KCore/Interp/commonCheckCoefs.h:1-93
//   K_INTERP::getExtrapolationCell   KCore/KCore/Interp/Interp.cpp:282-360, Interp2.cpp:84-137
// and the per-donor-zone packing (setInterpData.cpp:1183-1272), done here by a
// flag/scan/scatter so every zone's entries come out in ascending receptor
// order exactly like the reference's static-schedule chunk concatenation.
#include "cx_internal.h"
#include <stdlib.h>
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

using namespace cx;

namespace {

struct RcvTable {
  int* noblk; int* type; int* dind; int* isext;
  double* cf;   // [ncfmax][n]
  int n, ncfmax;
};

__device__ __forceinline__ void warp_add_stats(unsigned long long* g, const SearchStats& st)
{
  unsigned long long a = st.nvisit, b = st.ncand, c = st.ntet;
  for (int o = 16; o > 0; o >>= 1)
  {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
    c += __shfl_down_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&g[0], a); atomicAdd(&g[1], b); atomicAdd(&g[2], c); }
}

// Order-2 interpolation, all donor zones fused in one launch.  Per-thread scratch in dynamic shared memory
// (cx_core.cuh: Scratch): 24 doubles of cell vertices + `cap` stack entries per thread, column = thread.
// blockDim.x = CX_SEARCH_TB (64 or 128); shared memory, not registers, bounds the residency.
template <int TB>
__global__ void __launch_bounds__(TB, 512 / TB)
k_interp_o2(int n, const double* __restrict__ xr, const double* __restrict__ yr, const double* __restrict__ zr,
            int nz, const ZoneView* __restrict__ zones, int nature, int penalty, RcvTable T, int cap,
            unsigned long long* gstats, const RcvPerZone* __restrict__ pz)
{
  extern __shared__ __align__(16) double o2_sm[];
  Scratch S;
  S.vtx = o2_sm + threadIdx.x;
  S.stk = reinterpret_cast<int*>(o2_sm + 24 * TB) + threadIdx.x;
  S.ss = TB; S.cap = cap;
  int r = blockIdx.x * TB + threadIdx.x;
  SearchStats st; st.nvisit = 0; st.ncand = 0; st.ntet = 0;
  if (r < n)
  {
    RcvOut o;
    interp_o2_receptor_s(xr[r], yr[r], zr[r], nz, zones, nature, penalty, o, S, &st, pz, r);
    T.noblk[r] = o.noblk; T.type[r] = o.type; T.dind[r] = o.dind; T.isext[r] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) T.cf[(size_t)i * n + r] = o.cf[i];
  }
  if (gstats) warp_add_stats(gstats, st);
}

// Extrapolation pass over the receptors that found no donor: ONE WARP per receptor.
// The reference evaluates all 24 tetrahedra of every candidate cell and folds the candidates
// sequentially (InterpAdt.cpp:1113-1147).  Here lane 0 walks the ADT and hands out batches of up to 32
// candidates, the lanes evaluate one candidate each (extrap_coeff_for_cell, the expensive part), and the
// fold `ret==1 && sum < saved+1e-6 && diff < saved_diff` is then replayed in candidate order with
// shuffles, so the winner is the reference's.
__device__ __forceinline__ int search_extrap_cell_warp(const ZoneView& Z, double x, double y, double z, int& ic, int& jc,
                                                       int& kc, double* cf, int nature, double constraint,
                                                       int* s_cand, int lane, SearchStats* st)
{
  for (int i = 0; i < 8; i++) cf[i] = 0.;
  double alphaTol = fmax((2 * constraint - 5.) * 0.1, 0.);
  const bool cart = Z.topo == 0;
  // Cartesian donor (InterpCart::searchExtrapolationCellCart): the candidates are the cells of a small box
  // around the point, k-outer / i-inner, handed to the lanes in that order, 32 at a time
  int clo[3] = {0, 0, 0}, chi[3] = {0, 0, 0}, ctotal = 0, cbase = 0;
  if (cart)
  {
    ctotal = cart_candidate_box(Z, x, y, z, alphaTol, clo, chi);
    if (ctotal == 0) return 0;
  }
  else if (zone_reject(Z, x, y, z, alphaTol)) return 0;
  AdtQuery q;
  if (!cart && lane == 0) query_init(q, Z, x, y, z, alphaTol);
  const int ni = Z.ni, nij = Z.ni * Z.nj;
  double saved_sum = CX_MAXF, saved_diff = CX_MAXF;
  int found = 0, icsav = 0, jcsav = 0, kcsav = 0, ncand = 0;
  double cfs[8];
  for (int i = 0; i < 8; i++) cfs[i] = 0.;
  bool done = false;
  while (!done)
  {
    int cnt = 0;
    if (cart)
    {
      cnt = ctotal - cbase < 32 ? ctotal - cbase : 32;
      if (cbase + cnt >= ctotal) cnt |= 0x100;
    }
    else if (lane == 0)
    {
      int c, nid, nvisit = 0;
      while (cnt < 32 && (c = query_next(q, Z.nodes, &nid, &nvisit)) >= 0) s_cand[cnt++] = c;
      if (st) st->nvisit += nvisit;
      if (cnt < 32) cnt |= 0x100;        // traversal exhausted
    }
    __syncwarp();
    cnt = __shfl_sync(0xffffffffu, cnt, 0);
    done = (cnt & 0x100) != 0;
    cnt &= 0xff;
    ncand += cnt;
    int ret = 0, i = 0, j = 0, k = 0;
    double sum = 0., diff = CX_MAXF, cft[8];
    if (lane < cnt)
    {
      if (cart)
      {
        const int bi = chi[0] - clo[0] + 1, bj = chi[1] - clo[1] + 1;
        const int c = cbase + lane;
        k = c / (bi * bj); j = (c - k * bi * bj) / bi; i = c - j * bi - k * bi * bj;
        i += clo[0] - 1; j += clo[1] - 1; k += clo[2] - 1;
      }
      else
      {
        int c = s_cand[lane];
        k = c / nij; j = (c - k * nij) / ni; i = c - j * ni - k * nij;
      }
      k++; j++; i++;
      ret = extrap_coeff_for_cell(Z, x, y, z, i, j, k, cft, nature, constraint, diff);
      sum = fabs(cft[0]) + fabs(cft[1]) + fabs(cft[2]) + fabs(cft[3]) + fabs(cft[4]) + fabs(cft[5]) + fabs(cft[6]) +
            fabs(cft[7]);
    }
    __syncwarp();
    for (int l = 0; l < cnt; l++)
    {
      int rl = __shfl_sync(0xffffffffu, ret, l);
      double sl = __shfl_sync(0xffffffffu, sum, l);
      double dl = __shfl_sync(0xffffffffu, diff, l);
      if (rl == 1 && sl < saved_sum + 1.e-6 && dl < saved_diff)      // uniform across the warp
      {
        saved_diff = dl; saved_sum = sl; found = 1;
        icsav = __shfl_sync(0xffffffffu, i, l); jcsav = __shfl_sync(0xffffffffu, j, l);
        kcsav = __shfl_sync(0xffffffffu, k, l);
        for (int m = 0; m < 8; m++) cfs[m] = __shfl_sync(0xffffffffu, cft[m], l);
      }
    }
    cbase += cnt;
  }
  if (st && lane == 0) st->ncand += ncand;
  if (ncand == 0) return 0;
  ic = icsav; jc = jcsav; kc = kcsav;
  for (int m = 0; m < 8; m++) cf[m] = cfs[m];
  return found;
}

__global__ void __launch_bounds__(128)
k_extrap(const int* __restrict__ nlist_p, const int* __restrict__ list, const double* __restrict__ xr, const double* __restrict__ yr,
         const double* __restrict__ zr, int nz, const ZoneView* __restrict__ zones, int nature, int penalty,
         RcvTable T, unsigned long long* gstats, const RcvPerZone* __restrict__ pz)
{
  // the list length lives on the device (the host never waits for the interpolation pass): the warps stride over it
  __shared__ int s_cand[4][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nlist = *nlist_p;
  SearchStats st; st.nvisit = 0; st.ncand = 0; st.ntet = 0;
  for (int a = blockIdx.x * 4 + w; a < nlist; a += gridDim.x * 4)      // warp-uniform
  {
    const int r = list[a];
    double x = xr[r], y = yr[r], z = zr[r];
    // K_INTERP::getExtrapolationCell (Interp.cpp:282-360): constraint=40, extrapOrder=1, order forced to O2CF
    double best = CX_MAXF, sumCoefMin = CX_MAXF;
    int noblk = 0, type = 0, dind = -1;
    double cf[8], tcf[8];
    for (int i = 0; i < 8; i++) { cf[i] = 0.; tcf[i] = 0.; }
    for (int noz = 0; noz < nz; noz++)
    {
      const ZoneView& Z = zones[noz];
      if (pz) { x = pz[noz].x[r]; y = pz[noz].y[r]; z = pz[noz].z[r]; }
      int ic, jc, kc;
      int found = search_extrap_cell_warp(Z, x, y, z, ic, jc, kc, tcf, nature, 40., s_cand[w], lane, &st);
      if (found < 1) continue;
      ic -= 1; jc -= 1; kc -= 1;
      if (Z.nk == 1)
      {
        // 2-D donor: type 22, coefficients folded onto the plane (Interp2.cpp:125-136, commonTypesForExtrapAndInterp.h:72-105)
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
            best = vol2; type = 22; dind = tind2; noblk = noz + 1;
            for (int i = 0; i < 4; i++) { cf[i] = tcf[i] + tcf[i + 4]; cf[i + 4] = 0.; }
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
          best = vol; type = 2; dind = tind; noblk = noz + 1;
          for (int i = 0; i < 8; i++) cf[i] = tcf[i];
        }
      }
    }
    if (noblk > 0 && lane == 0)
    {
      collapse_type2(zones[noblk - 1], type, dind, cf);
      T.noblk[r] = noblk; T.type[r] = type; T.dind[r] = dind; T.isext[r] = 1;
      for (int i = 0; i < 8; i++) T.cf[(size_t)i * T.n + r] = cf[i];
    }
  }
  if (gstats) warp_add_stats(gstats, st);
}

// ---- extrapolation, second form: the walks of 16 receptors run side by side -------------------------------------
// In k_extrap lane 0 walks the tree while 31 lanes wait: 63 % of the issued instructions ran on one lane (ncu,
// profiles/r03_*: 12 active threads per instruction, fp64 pipe 11 % busy, 8 resident warps per SM at 228 registers).
// Here a warp takes a group of XG receptors and ONE donor zone: lanes 0..XG-1 walk their own receptor's tree
// and buffer up to CX_XB candidates each in shared memory (in the reference's list order), then the whole warp
// evaluates the buffered candidates of one receptor after the other, 32 at a time, and replays the reference's
// sequential fold (InterpAdt.cpp:1113-1147) in candidate order; rounds repeat until every walk is exhausted.  The
// per-(receptor, zone) winner goes to a table; k_extrap_combine then runs the cross-zone choice of
// getExtrapolationCell (Interp.cpp:282-360) zone after zone, exactly as k_extrap does.
#define CX_XB 64
struct ExtrapZoneRes { int found, ic, jc, kc; double cf[8]; };   // per (zone, failing receptor of the chunk)

template <int MINB, int XG>
__global__ void __launch_bounds__(128, MINB)
k_extrap_zone(const int* __restrict__ nlist_p, const int* __restrict__ list, int chunk0, int chunkN,
              const double* __restrict__ xr, const double* __restrict__ yr, const double* __restrict__ zr, int nz,
              const ZoneView* __restrict__ zones, int nature, ExtrapZoneRes* __restrict__ out, unsigned long long* gstats,
              const RcvPerZone* __restrict__ pz)
{
  __shared__ int s_buf[4][XG][CX_XB];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nlist = min(*nlist_p, chunk0 + chunkN);
  const int ngroups = (max(nlist - chunk0, 0) + XG - 1) / XG;
  const long long nitems = (long long)ngroups * nz;
  SearchStats st; st.nvisit = 0; st.ncand = 0; st.ntet = 0;
  const double constraint = 40.;
  const double alphaTol = fmax((2 * constraint - 5.) * 0.1, 0.);
  for (long long item = (long long)blockIdx.x * 4 + w; item < nitems; item += (long long)gridDim.x * 4)   // warp-uniform
  {
    const int g = (int)(item / nz), noz = (int)(item - (long long)g * nz);
    const ZoneView& Z = zones[noz];
    const bool cart = Z.topo == 0;
    const int a = chunk0 + g * XG + lane;
    const bool mine = lane < XG && a < nlist;
    double x = 0., y = 0., z = 0.;
    if (mine)
    {
      const int r = list[a];
      if (pz) { x = pz[noz].x[r]; y = pz[noz].y[r]; z = pz[noz].z[r]; }
      else { x = xr[r]; y = yr[r]; z = zr[r]; }
    }
    // walk state of this lane's receptor
    AdtQuery q;
    int clo[3] = {0, 0, 0}, chi[3] = {0, 0, 0}, ctotal = 0, cbase = 0;
    bool walking = mine;
    if (mine)
    {
      if (cart) { ctotal = cart_candidate_box(Z, x, y, z, alphaTol, clo, chi); if (ctotal == 0) walking = false; }
      else if (zone_reject(Z, x, y, z, alphaTol)) walking = false;
      else query_init(q, Z, x, y, z, alphaTol);
    }
    // fold state of this lane's receptor (InterpAdt.cpp:1113-1147)
    double saved_sum = CX_MAXF, saved_diff = CX_MAXF;
    int found = 0, icsav = 0, jcsav = 0, kcsav = 0, ncand = 0;
    double cfs[8];
    for (int i = 0; i < 8; i++) cfs[i] = 0.;
    const int ni = Z.ni, nij = Z.ni * Z.nj;
    while (true)
    {
      // every walking lane buffers its next candidates, in list order
      int cnt = 0;
      if (walking)
      {
        if (cart)
        {
          // InterpCart::searchExtrapolationCellCart: the cells of a small box around the point, k-outer / i-inner
          const int bi = chi[0] - clo[0] + 1, bj = chi[1] - clo[1] + 1;
          while (cnt < CX_XB && cbase < ctotal)
          {
            const int c = cbase++;
            const int k = c / (bi * bj), j = (c - k * bi * bj) / bi, i = c - j * bi - k * bi * bj;
            s_buf[w][lane][cnt++] = (i + clo[0] - 1) + (j + clo[1] - 1) * ni + (k + clo[2] - 1) * nij;
          }
          if (cbase >= ctotal) walking = false;
        }
        else
        {
          int c, nid, nvisit = 0;
          while (cnt < CX_XB && (c = query_next(q, Z.nodes, &nid, &nvisit)) >= 0) s_buf[w][lane][cnt++] = c;
          st.nvisit += nvisit;
          if (cnt < CX_XB) walking = false;          // traversal exhausted
        }
      }
      __syncwarp();
      unsigned have = __ballot_sync(0xffffffffu, cnt > 0);
      if (have == 0) { if (__ballot_sync(0xffffffffu, walking) == 0) break; else continue; }
      // the warp evaluates the buffered candidates of one receptor after the other
      while (have)
      {
        const int rl = __ffs(have) - 1; have &= have - 1;
        const int n_rl = __shfl_sync(0xffffffffu, cnt, rl);
        const double px = __shfl_sync(0xffffffffu, x, rl), py = __shfl_sync(0xffffffffu, y, rl), pz = __shfl_sync(0xffffffffu, z, rl);
        double f_sum = __shfl_sync(0xffffffffu, saved_sum, rl), f_diff = __shfl_sync(0xffffffffu, saved_diff, rl);
        int f_new = 0, f_i = 0, f_j = 0, f_k = 0;
        double f_cf[8];
        for (int m = 0; m < 8; m++) f_cf[m] = 0.;
        for (int b0 = 0; b0 < n_rl; b0 += 32)
        {
          const int nb = min(32, n_rl - b0);
          int ret = 0, i = 0, j = 0, k = 0;
          double sum = 0., diff = CX_MAXF, cft[8];
          if (lane < nb)
          {
            const int c = s_buf[w][rl][b0 + lane];
            k = c / nij; j = (c - k * nij) / ni; i = c - j * ni - k * nij;
            k++; j++; i++;
            ret = extrap_coeff_for_cell(Z, px, py, pz, i, j, k, cft, nature, constraint, diff);
            sum = fabs(cft[0]) + fabs(cft[1]) + fabs(cft[2]) + fabs(cft[3]) + fabs(cft[4]) + fabs(cft[5]) + fabs(cft[6]) +
                  fabs(cft[7]);
          }
          __syncwarp();
          for (int l = 0; l < nb; l++)
          {
            const int rl_ = __shfl_sync(0xffffffffu, ret, l);
            const double sl = __shfl_sync(0xffffffffu, sum, l);
            const double dl = __shfl_sync(0xffffffffu, diff, l);
            if (rl_ == 1 && sl < f_sum + 1.e-6 && dl < f_diff)      // uniform across the warp
            {
              f_diff = dl; f_sum = sl; f_new = 1;
              f_i = __shfl_sync(0xffffffffu, i, l); f_j = __shfl_sync(0xffffffffu, j, l); f_k = __shfl_sync(0xffffffffu, k, l);
              for (int m = 0; m < 8; m++) f_cf[m] = __shfl_sync(0xffffffffu, cft[m], l);
            }
          }
        }
        if (lane == rl)
        {
          ncand += n_rl;
          if (f_new)
          {
            saved_sum = f_sum; saved_diff = f_diff; found = 1; icsav = f_i; jcsav = f_j; kcsav = f_k;
            for (int m = 0; m < 8; m++) cfs[m] = f_cf[m];
          }
        }
      }
      __syncwarp();
    }
    if (mine)
    {
      st.ncand += ncand;
      ExtrapZoneRes& o = out[(size_t)noz * chunkN + (size_t)(a - chunk0)];
      o.found = (ncand > 0) ? found : 0; o.ic = icsav; o.jc = jcsav; o.kc = kcsav;
      for (int m = 0; m < 8; m++) o.cf[m] = cfs[m];
    }
  }
  if (gstats) warp_add_stats(gstats, st);
}

// K_INTERP::getExtrapolationCell (Interp.cpp:282-360: constraint=40, extrapOrder=1, order forced to O2CF): the
// choice between the donor zones, zone after zone, from the per-zone winners of k_extrap_zone
__global__ void k_extrap_combine(const int* __restrict__ nlist_p, const int* __restrict__ list, int chunk0, int chunkN,
                                 int nz, const ZoneView* __restrict__ zones, int penalty,
                                 const ExtrapZoneRes* __restrict__ res, RcvTable T)
{
  const int nlist = min(*nlist_p, chunk0 + chunkN);
  for (int a = chunk0 + blockIdx.x * blockDim.x + threadIdx.x; a < nlist; a += gridDim.x * blockDim.x)
  {
    const int r = list[a];
    double best = CX_MAXF, sumCoefMin = CX_MAXF;
    int noblk = 0, type = 0, dind = -1;
    double cf[8], tcf[8];
    for (int i = 0; i < 8; i++) { cf[i] = 0.; tcf[i] = 0.; }
    for (int noz = 0; noz < nz; noz++)
    {
      const ExtrapZoneRes& o = res[(size_t)noz * chunkN + (size_t)(a - chunk0)];
      if (o.found < 1) continue;
      const ZoneView& Z = zones[noz];
      for (int i = 0; i < 8; i++) tcf[i] = o.cf[i];
      const int ic = o.ic - 1, jc = o.jc - 1, kc = o.kc - 1;
      if (Z.nk == 1)
      {
        // 2-D donor: type 22, coefficients folded onto the plane (Interp2.cpp:125-136, commonTypesForExtrapAndInterp.h:72-105)
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
            best = vol2; type = 22; dind = tind2; noblk = noz + 1;
            for (int i = 0; i < 4; i++) { cf[i] = tcf[i] + tcf[i + 4]; cf[i + 4] = 0.; }
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
          best = vol; type = 2; dind = tind; noblk = noz + 1;
          for (int i = 0; i < 8; i++) cf[i] = tcf[i];
        }
      }
    }
    if (noblk > 0)
    {
      collapse_type2(zones[noblk - 1], type, dind, cf);
      T.noblk[r] = noblk; T.type[r] = type; T.dind[r] = dind; T.isext[r] = 1;
      for (int i = 0; i < 8; i++) T.cf[(size_t)i * T.n + r] = cf[i];
    }
  }
}

__device__ __forceinline__ int ncf_of_type(int t)
{ // SIZECF, Connector/Connector/connector.h:27-37 (structured donors)
  return t == 1 ? 1 : t == 2 ? 8 : t == 3 ? 9 : t == 5 ? 15 : t == 44 ? 12 : t == 22 ? 4 : 0;
}

// per-zone totals in one pass: cnt[z*3+0] entries, +1 coefficients, +2 extrapolated; cnt[nz*3] orphans
// (lanes of a warp with the same zone are reduced with match/reduce before the atomics)
__global__ void k_zone_hist(int n, int nz, const int* __restrict__ noblk, const int* __restrict__ type,
                            const int* __restrict__ isext, unsigned long long* __restrict__ cnt)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  bool on = r < n;
  unsigned live = __ballot_sync(0xffffffffu, on);
  if (!on) return;
  int b = noblk[r];
  unsigned ncf = b > 0 ? (unsigned)ncf_of_type(type[r]) : 0u;
  unsigned ext = (b > 0 && isext[r] == 1) ? 1u : 0u;
  unsigned grp = __match_any_sync(live, b);
  unsigned s2 = __reduce_add_sync(grp, ncf);
  unsigned s3 = __reduce_add_sync(grp, ext);
  if ((int)(threadIdx.x & 31) != __ffs(grp) - 1) return;
  if (b > 0)
  {
    atomicAdd(&cnt[(b - 1) * 3 + 0], (unsigned long long)__popc(grp));
    atomicAdd(&cnt[(b - 1) * 3 + 1], (unsigned long long)s2);
    if (s3) atomicAdd(&cnt[(b - 1) * 3 + 2], (unsigned long long)s3);
  }
  else atomicAdd(&cnt[nz * 3], (unsigned long long)__popc(grp));
}

// receptors the interpolation pass left without a donor, appended in any order (every one is extrapolated on its own)
__global__ void k_collect_fail(int n, const int* __restrict__ noblk, int* __restrict__ fail, int* __restrict__ nfail)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool f = r < n && noblk[r] == 0;
  const unsigned m = __ballot_sync(0xffffffffu, f);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(nfail, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (f) fail[base + __popc(m & ((1u << lane) - 1u))] = r;
}

// ---- packing: ONE pass for all donor zones ------------------------------------------------------------------
// A stable sort of the receptors by donor zone number (0 = orphan) gives, zone after zone, the entries in ascending
// receptor order (setInterpData.cpp:1183-1252: the static-schedule chunks concatenated); one 64-bit scan over that
// order gives every entry its coefficient offset (low 40 bits) and its rank among the extrapolated entries (high
// bits), and one scatter writes the ragged lists of all zones.
__global__ void k_pack_keys(int n, const int* __restrict__ noblk, unsigned* __restrict__ keys, int* __restrict__ vals)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  keys[r] = (unsigned)noblk[r]; vals[r] = r;
}

__global__ void k_pack_flags(int n, const int* __restrict__ order, RcvTable T, unsigned long long* __restrict__ f)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = order[i];
  unsigned long long v = 0;
  if (T.noblk[r] > 0) v = (unsigned long long)ncf_of_type(T.type[r]) | (T.isext[r] == 1 ? (1ull << 40) : 0ull);
  f[i] = v;
}

__global__ void k_pack_scatter(int n, int nz, const int* __restrict__ order, RcvTable T,
                               const unsigned long long* __restrict__ pos, const unsigned long long* __restrict__ cnt,
                               int* __restrict__ rcv, int* __restrict__ dnr, int* __restrict__ typ,
                               double* __restrict__ coefs, int* __restrict__ ext, int* __restrict__ orphan)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = order[i];
  if (T.noblk[r] == 0) { orphan[i] = r; return; }         // key 0 sorts first: position == rank among the orphans
  const int e = i - (int)cnt[(size_t)nz * 3];
  const unsigned long long p = pos[i];
  const size_t c = (size_t)(p & ((1ull << 40) - 1ull));
  const int t = T.type[r];
  rcv[e] = r; dnr[e] = T.dind[r]; typ[e] = t;
  const int ncf = ncf_of_type(t);
  for (int k = 0; k < ncf; k++) coefs[c + k] = T.cf[(size_t)k * n + r];
  if (T.isext[r] == 1) ext[(size_t)(p >> 40)] = r;
}

}  // namespace

struct cx_result_dev {
  // packed outputs, zones concatenated
  int* rcv; int* dnr; int* typ; double* coefs; int* ext; int* orphan;
  std::vector<long long> e_off, c_off, x_off;
};

extern int cx_lagrange_pipeline(cx_ctx* ctx, int order, int n, const double* d_xr, const double* d_yr,
                                const double* d_zr, int nz, const ZoneView* d_zones, const ZoneView* h_zones,
                                int nature, int penalty, int* noblk, int* type, int* dind, double* cf, int ncfmax,
                                unsigned long long* d_stats, const RcvPerZone* h_pz);

int cx_search_run(cx_ctx* ctx, int n, const double* d_xr, const double* d_yr, const double* d_zr, int nz,
                  const ZoneView* d_zones, const ZoneView* h_zones, int order, int nature, int penalty, int extrap,
                  cx_result* R, cx_result_dev* D, const RcvPerZone* d_pz, const RcvPerZone* h_pz)
{
  cudaStream_t s = ctx->stream;
  if (d_pz && order == 4) { cx_set_error("setInterpDataDW: orders 2, 3 and 5 (the reference knows no order 4 here)"); return -4; }
  const int ncfmax = (order == 3) ? 9 : (order == 4) ? 12 : (order == 5) ? 15 : 8;   // setInterpData.cpp:624-652
  R->nrcv = n; R->nz = nz; R->ncfmax = ncfmax;
  R->cnt.assign(nz, 0); R->ncoef.assign(nz, 0); R->nextrap.assign(nz, 0); R->norphan = 0;
  for (int i = 0; i < 4; i++) R->stats[i] = 0;
  D->e_off.assign(nz + 1, 0); D->c_off.assign(nz + 1, 0); D->x_off.assign(nz + 1, 0);
  D->rcv = D->dnr = D->typ = D->ext = D->orphan = nullptr; D->coefs = nullptr;
  R->d_noblk = R->d_type = R->d_dind = R->d_isext = nullptr; R->d_cf = nullptr;
  R->pending = 0; R->h_counts = nullptr; R->h_counts_own = 0;
  for (int i = 0; i < 5; i++) R->ev[i] = nullptr;
  if (n == 0) return 0;   // (the zone table of the call, if any, is released with the result)
  size_t N = n;
  for (int i = 0; i < 5; i++) CX_CUDA(cudaEventCreate(&R->ev[i]));
  CX_MALLOC(R->d_noblk, sizeof(int) * N, s);
  CX_MALLOC(R->d_type, sizeof(int) * N, s);
  CX_MALLOC(R->d_dind, sizeof(int) * N, s);
  CX_MALLOC(R->d_isext, sizeof(int) * N, s);
  CX_MALLOC(R->d_cf, sizeof(double) * N * ncfmax, s);
  unsigned long long* d_stats;
  CX_MALLOC(d_stats, sizeof(unsigned long long) * 4, s);
  CX_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(unsigned long long) * 4, s));
  RcvTable T{R->d_noblk, R->d_type, R->d_dind, R->d_isext, R->d_cf, n, ncfmax};
  const int TB = 128;

  CX_CUDA(cudaEventRecord(R->ev[0], s));
  if (order == 2)
  {
    int cap = 4;
    for (int z = 0; z < nz; z++) if (h_zones[z].topo != 0) cap = std::max(cap, h_zones[z].depth + 2);
    static int tbs = 0;        // CX_SEARCH_TB: threads per CTA of the search kernel (64 or 128)
    if (!tbs) { const char* e = getenv("CX_SEARCH_TB"); tbs = e ? atoi(e) : 128; if (tbs != 64 && tbs != 128) tbs = 128; }
    const size_t smb = (size_t)tbs * (24 * sizeof(double) + (size_t)cap * sizeof(int));
    if (smb > 200 * 1024) { cx_set_error("setInterpData: ADT depth %d needs %zu bytes of shared memory per CTA", cap - 2, smb); return -1; }
    if (tbs == 64)
    {
      CX_CUDA(cudaFuncSetAttribute(k_interp_o2<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
      k_interp_o2<64><<<cx_grid(n, 64), 64, smb, s>>>(n, d_xr, d_yr, d_zr, nz, d_zones, nature, penalty, T, cap, d_stats, d_pz);
    }
    else
    {
      CX_CUDA(cudaFuncSetAttribute(k_interp_o2<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
      k_interp_o2<128><<<cx_grid(n, 128), 128, smb, s>>>(n, d_xr, d_yr, d_zr, nz, d_zones, nature, penalty, T, cap, d_stats, d_pz);
    }
    ctx->launches++;
  }
  else
  {
    CX_CUDA(cudaMemsetAsync(R->d_isext, 0, sizeof(int) * N, s));
    CX_CHECK(cx_lagrange_pipeline(ctx, order, n, d_xr, d_yr, d_zr, nz, d_zones, h_zones, nature, penalty,
                                  R->d_noblk, R->d_type, R->d_dind, R->d_cf, ncfmax, d_stats, h_pz));
  }
  CX_CUDA(cudaEventRecord(R->ev[1], s));

  // ---- extrapolation of the receptors without a donor; nothing here waits for the device ----
  int* d_fail = nullptr; int* d_nfail = nullptr;
  CX_CUDA(cudaEventRecord(R->ev[2], s));
  const int TB2 = 256;
  if (extrap == 1)
  {
    CX_MALLOC(d_fail, sizeof(int) * N, s);
    CX_MALLOC(d_nfail, sizeof(int), s);
    CX_CUDA(cudaMemsetAsync(d_nfail, 0, sizeof(int), s));
    k_collect_fail<<<cx_grid(n, TB2), TB2, 0, s>>>(n, R->d_noblk, d_fail, d_nfail);
    static int xmode = -1;       // CX_EXTRAP=warp: the first form (one warp per receptor, lane 0 walks); default: grouped
    if (xmode < 0) { const char* e = getenv("CX_EXTRAP"); xmode = (e && e[0] == 'w') ? 1 : 0; }
    if (xmode == 1)
    {
      const int gx = std::max(1, std::min(cx_grid(n, 4), ctx->sm_count * 16));
      k_extrap<<<gx, 128, 0, s>>>(d_nfail, d_fail, d_xr, d_yr, d_zr, nz, d_zones, nature, penalty, T, d_stats, d_pz);
      ctx->launches += 2;
    }
    else
    {
      // the failing receptors go through in chunks of at most CHUNK (the per-(zone, receptor) table is sized by the
      // chunk, not by the unknown number of failures: the count stays on the device, chunks beyond it do nothing)
      const int CHUNK = (int)std::min<long long>(n, std::max<long long>(1 << 16, (512ll << 20) / ((long long)sizeof(ExtrapZoneRes) * nz)));
      ExtrapZoneRes* d_res;
      CX_MALLOC(d_res, sizeof(ExtrapZoneRes) * (size_t)CHUNK * nz, s);
      // CX_XMINB = resident CTAs per SM (2 | 4 | 6: 246 / 128 / 80 registers, the last two with ~0.5 KB of spills),
      // CX_XG = receptors per warp (16 | 8 | 4).  The kernel is bound by the latency of each warp's dependent chain:
      // more resident warps and shorter chains win over fewer spills and wider walks (measured: gpurun_out/r3_xg_*).
      static int minb = 0, xg = 0;
      if (!minb)
      {
        const char* e = getenv("CX_XMINB"); minb = e ? atoi(e) : 4; if (minb != 2 && minb != 4 && minb != 6) minb = 4;
        const char* g = getenv("CX_XG"); xg = g ? atoi(g) : 16; if (xg != 16 && xg != 8 && xg != 4) xg = 16;
      }
      const long long items = (long long)((CHUNK + xg - 1) / xg) * nz;
      const int gz = (int)std::max(1ll, std::min((items + 3) / 4, (long long)ctx->sm_count * 32));
      const int gc = std::max(1, std::min(cx_grid(CHUNK, 128), ctx->sm_count * 8));
      for (int c0 = 0; c0 < n; c0 += CHUNK)
      {
#define CX_XZ(MB, G) k_extrap_zone<MB, G><<<gz, 128, 0, s>>>(d_nfail, d_fail, c0, CHUNK, d_xr, d_yr, d_zr, nz, d_zones, nature, d_res, d_stats, d_pz)
        if (xg == 16) { if (minb == 2) CX_XZ(2, 16); else if (minb == 4) CX_XZ(4, 16); else CX_XZ(6, 16); }
        else if (xg == 8) { if (minb == 2) CX_XZ(2, 8); else if (minb == 4) CX_XZ(4, 8); else CX_XZ(6, 8); }
        else { if (minb == 2) CX_XZ(2, 4); else if (minb == 4) CX_XZ(4, 4); else CX_XZ(6, 4); }
#undef CX_XZ
        k_extrap_combine<<<gc, 128, 0, s>>>(d_nfail, d_fail, c0, CHUNK, nz, d_zones, penalty, d_res, T);
        ctx->launches += 2;
      }
      ctx->launches += 1;
      cudaFreeAsync(d_res, s);
    }
    cudaFreeAsync(d_fail, s); cudaFreeAsync(d_nfail, s);
  }
  CX_CUDA(cudaEventRecord(R->ev[3], s));

  // ---- packing, all donor zones in one pass; outputs sized by their upper bounds (the counts stay on the device
  // until cx_search_finish reads them) ----
  unsigned long long *d_f, *d_p, *d_cnt;
  CX_MALLOC(d_f, sizeof(unsigned long long) * N, s);
  CX_MALLOC(d_p, sizeof(unsigned long long) * N, s);
  CX_MALLOC(d_cnt, sizeof(unsigned long long) * (size_t)(nz * 3 + 1), s);
  CX_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * (size_t)(nz * 3 + 1), s));
  k_zone_hist<<<cx_grid(n, TB2), TB2, 0, s>>>(n, nz, R->d_noblk, R->d_type, R->d_isext, d_cnt);
  unsigned *k_in, *k_out; int *v_in, *v_out;
  CX_MALLOC(k_in, sizeof(unsigned) * N, s); CX_MALLOC(k_out, sizeof(unsigned) * N, s);
  CX_MALLOC(v_in, sizeof(int) * N, s); CX_MALLOC(v_out, sizeof(int) * N, s);
  k_pack_keys<<<cx_grid(n, TB2), TB2, 0, s>>>(n, R->d_noblk, k_in, v_in);
  int bits = 1; while ((1 << bits) <= nz) bits++;
  size_t tmp_bytes = 0; void* tmp = nullptr;
  CX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, n, 0, bits, s));
  CX_MALLOC(tmp, tmp_bytes, s);
  CX_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, n, 0, bits, s));
  k_pack_flags<<<cx_grid(n, TB2), TB2, 0, s>>>(n, v_out, T, d_f);
  CX_CHECK(cx_exclusive_scan_u64(ctx, d_f, d_p, n, nullptr));
  CX_MALLOC(D->rcv, sizeof(int) * N, s);
  CX_MALLOC(D->dnr, sizeof(int) * N, s);
  CX_MALLOC(D->typ, sizeof(int) * N, s);
  CX_MALLOC(D->ext, sizeof(int) * N, s);
  CX_MALLOC(D->orphan, sizeof(int) * N, s);
  CX_MALLOC(D->coefs, sizeof(double) * N * ncfmax, s);
  k_pack_scatter<<<cx_grid(n, TB2), TB2, 0, s>>>(n, nz, v_out, T, d_p, d_cnt, D->rcv, D->dnr, D->typ, D->coefs, D->ext, D->orphan);
  ctx->launches += 5;
  cudaFreeAsync(k_in, s); cudaFreeAsync(k_out, s); cudaFreeAsync(v_in, s); cudaFreeAsync(v_out, s); cudaFreeAsync(tmp, s);
  cudaFreeAsync(d_f, s); cudaFreeAsync(d_p, s);
  // counts and statistics come back through page-locked memory, so the copy does not stall the host either
  const size_t ncnt = (size_t)nz * 3 + 1 + 4;
  R->h_counts = (unsigned long long*)cx_pin_take(ctx, sizeof(unsigned long long) * ncnt, &R->h_counts_own);
  if (!R->h_counts) { cx_set_error("setInterpData: no page-locked memory for the counts"); return -2; }
  CX_CUDA(cudaMemcpyAsync(R->h_counts, d_cnt, sizeof(unsigned long long) * (ncnt - 4), cudaMemcpyDeviceToHost, s));
  CX_CUDA(cudaMemcpyAsync(R->h_counts + (ncnt - 4), d_stats, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, s));
  CX_CUDA(cudaEventRecord(R->ev[4], s));
  cudaFreeAsync(d_stats, s); cudaFreeAsync(d_cnt, s);
  R->pending = 1;
  return 0;
}

// wait for a launched search and read its counts (every accessor of a result goes through here)
int cx_search_finish(cx_result* R, cx_result_dev* D)
{
  if (!R->pending) return 0;
  R->pending = 0;
  const int nz = R->nz;
  CX_CUDA(cudaEventSynchronize(R->ev[4]));
  CX_CUDA(cudaGetLastError());
  const unsigned long long* hc = R->h_counts;
  long long eo = 0, co = 0, xo = 0;
  for (int z = 0; z < nz; z++)
  {
    D->e_off[z] = eo; D->c_off[z] = co; D->x_off[z] = xo;
    R->cnt[z] = (int)hc[z * 3]; R->ncoef[z] = (int)hc[z * 3 + 1]; R->nextrap[z] = (int)hc[z * 3 + 2];
    eo += R->cnt[z]; co += R->ncoef[z]; xo += R->nextrap[z];
  }
  D->e_off[nz] = eo; D->c_off[nz] = co; D->x_off[nz] = xo;
  R->norphan = (int)hc[(size_t)nz * 3];
  for (int i = 0; i < 4; i++) R->stats[i] = hc[(size_t)nz * 3 + 1 + i];
  cudaEventElapsedTime(&R->search_ms, R->ev[0], R->ev[1]);
  cudaEventElapsedTime(&R->extrap_ms, R->ev[2], R->ev[3]);
  cx_pin_give(R->ctx, R->h_counts, R->h_counts_own);
  R->h_counts = nullptr;
  if (R->d_zones_owned) { cudaFreeAsync(R->d_zones_owned, R->ctx->stream); R->d_zones_owned = nullptr; }
  if (R->h_zones) { cx_pin_give(R->ctx, R->h_zones, R->h_zones_own); R->h_zones = nullptr; }
  for (int i = 0; i < 5; i++) { cudaEventDestroy(R->ev[i]); R->ev[i] = nullptr; }
  return 0;
}
