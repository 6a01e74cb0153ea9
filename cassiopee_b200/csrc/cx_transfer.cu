// cx_transfer.cu — setInterpTransfers on the device.
//
// Replaces the variable-parallel CPU loops of
//   Connector/Connector/commonInterpTransfers_indirect.h:1-39 (in-place, indR = rcvPts[n])
//   Connector/Connector/commonInterpTransfers_direct.h        (dense (nvars,nrcv), indR = n)
//   Connector/Connector/commonInterpTransfers.h:1-151         (per-type stencils, Appendix B of SURVEY.md)
//   Connector/Connector/commonCellNTransfersStrict.h          (strict cellN rule)
// with receptor-parallel kernels.  The reference walks the ragged coefficient
// array with a running pointer (stride SIZECF(type), connector.h:27-37); the
// plan replaces that by grouping the entries by type once (stable) and storing
// each group's coefficients slot-major ([ncf][m]) so that a warp's coefficient
// loads are fully coalesced.  The variable loop is innermost over SoA fields.
// Accumulation order is the reference's (kk outer, ii inner, val=0; val+=c*f)
// and the kernels are compiled without FMA contraction, so results are
// bit-identical to the CPU code, not merely within 1e-12.
//
// Roofline: HBM.  Algorithmic bytes per entry: 8*ncf + 4 (donor) + 4 (rcv) + 4 (type)
// + 8*nvars (receptor write) + 8*nvars*U/n (each distinct donor node read once).
#include "cx_internal.h"
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include <vector>

#define CX_MAXVARS 16

struct FieldPtrs { const double* d[CX_MAXVARS]; double* r[CX_MAXVARS]; };

namespace {

template <int O> struct Sten { };

// MODE 0: scatter to fieldR[v][rcv[e]];  MODE 1: dense out[v*ld + pos[e]]
template <int TYPE, int MODE>
__global__ void __launch_bounds__(256)
k_transfer(int m, int nvars, FieldPtrs P, int imd, int imdjmd, const int* __restrict__ donor,
           const int* __restrict__ rcv, const int* __restrict__ pos, const double* __restrict__ cfT,
           double* __restrict__ dense, long long ld)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int d0 = donor[e];
  size_t o;
  if (MODE == 0) o = (size_t)rcv[e]; else o = (size_t)pos[e];
  if (TYPE == 1)
  {
    double c0 = cfT[e];
    for (int v = 0; v < nvars; v++)
    {
      double val = c0 * __ldg(P.d[v] + d0);
      if (MODE == 0) P.r[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
  }
  else if (TYPE == 2)
  {
    double c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = cfT[(size_t)k * m + e];
    for (int v = 0; v < nvars; v++)
    {
      const double* f = P.d[v] + d0;
      double val = 0.;
      val += c[0] * __ldg(f);
      val += c[1] * __ldg(f + 1);
      val += c[2] * __ldg(f + imd);
      val += c[3] * __ldg(f + imd + 1);
      val += c[4] * __ldg(f + imdjmd);
      val += c[5] * __ldg(f + imdjmd + 1);
      val += c[6] * __ldg(f + imdjmd + imd);
      val += c[7] * __ldg(f + imdjmd + imd + 1);
      if (MODE == 0) P.r[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
  }
  else
  {
    constexpr int O = (TYPE == 3) ? 3 : 5;
    double c[3 * O];
#pragma unroll
    for (int k = 0; k < 3 * O; k++) c[k] = cfT[(size_t)k * m + e];
    for (int v = 0; v < nvars; v++)
    {
      const double* f = P.d[v] + d0;
      double val = 0.;
#pragma unroll
      for (int kk = 0; kk < O; kk++)
#pragma unroll
        for (int jj = 0; jj < O; jj++)
        {
          const double* g = f + jj * imd + kk * imdjmd;
#pragma unroll
          for (int ii = 0; ii < O; ii++)
            val += c[ii] * c[jj + O] * c[kk + 2 * O] * __ldg(g + ii);   // left to right, as commonInterpTransfers.h:85,141
        }
      if (MODE == 0) P.r[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
  }
}

// strict cellN rule (commonCellNTransfersStrict.h): product of cellN*(2-cellN)
// over stencil nodes whose coefficient exceeds 1e-11; > 0.5 -> 1 else 0
template <int TYPE>
__global__ void k_transfer_celln(int m, const double* __restrict__ cellND, double* __restrict__ cellNR, int imd,
                                 int imdjmd, const int* __restrict__ donor, const int* __restrict__ rcv,
                                 const double* __restrict__ cfT)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int d0 = donor[e];
  double val = 1.;
  if (TYPE == 1) { double c = cellND[d0]; val = c * (2. - c); }
  else if (TYPE == 2)
  {
    for (int kk = 0; kk < 2; kk++) for (int jj = 0; jj < 2; jj++) for (int ii = 0; ii < 2; ii++)
    {
      double cf = cfT[(size_t)(ii + 2 * jj + 4 * kk) * m + e];
      if (fabs(cf) > 1.e-11) { double c = cellND[d0 + ii + jj * imd + kk * imdjmd]; val *= c * (2. - c); }
    }
  }
  else
  {
    constexpr int O = (TYPE == 3) ? 3 : 5;
    for (int kk = 0; kk < O; kk++) for (int jj = 0; jj < O; jj++) for (int ii = 0; ii < O; ii++)
    {
      double cf = cfT[(size_t)ii * m + e] * cfT[(size_t)(jj + O) * m + e] * cfT[(size_t)(kk + 2 * O) * m + e];
      if (fabs(cf) > 1.e-11) { double c = cellND[d0 + ii + jj * imd + kk * imdjmd]; val *= c * (2. - c); }
    }
  }
  cellNR[rcv[e]] = (val > 0.5) ? 1. : 0.;
}

}  // namespace

static int ncf_host(int t) { return t == 1 ? 1 : t == 2 ? 8 : t == 3 ? 9 : t == 5 ? 15 : -1; }
static int sten_host(int t) { return t == 1 ? 1 : t == 2 ? 8 : t == 3 ? 27 : t == 5 ? 125 : 0; }

namespace {
// plan build: counting sort of one type group by donor index (so a warp's
// stencil gathers come from adjacent donor cells) + slot-major coefficients
__global__ void k_plan_hist(int n, int T, const int* __restrict__ type, const int* __restrict__ donor,
                            unsigned long long* __restrict__ cnt)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || type[e] != T) return;
  atomicAdd(&cnt[donor[e]], 1ull);
}
__global__ void k_plan_place(int n, int T, int ncf, int m, int sorted, const int* __restrict__ type,
                             const int* __restrict__ donor, const int* __restrict__ rcv,
                             const unsigned long long* __restrict__ off, const double* __restrict__ coefs,
                             unsigned long long* cursor, const unsigned long long* rank,
                             int* __restrict__ od, int* __restrict__ orc, int* __restrict__ op, double* __restrict__ cfT)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || type[e] != T) return;
  int d0 = donor[e];
  int slot = sorted ? (int)atomicAdd(&cursor[d0], 1ull) : (int)rank[e];
  od[slot] = d0; orc[slot] = rcv[e]; op[slot] = e;
  const double* c = coefs + off[e];
  for (int k = 0; k < ncf; k++) cfT[(size_t)k * m + slot] = c[k];
}
__global__ void k_plan_flag(int n, int T, const int* __restrict__ type, unsigned long long* __restrict__ f)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) f[e] = (type[e] == T) ? 1ull : 0ull;
}
}  // namespace

static int plan_sort_mode()
{
  const char* s = getenv("CX_PLAN_SORT");
  return (s && s[0] == '0') ? 0 : 1;
}

int cx_plan_build(cx_ctx* ctx, int imd, int jmd, long long nptsD, long long nptsR, int n, const int* donor,
                  const int* rcv, const int* type, const double* coefs, long long ncoefs, cx_plan* P)
{
  P->ctx = ctx; P->imd = imd; P->jmd = jmd; P->nptsD = (int)nptsD; P->nptsR = (int)nptsR; P->n = n;
  P->d_donor = P->d_rcv = P->d_pos = nullptr; P->alg_bytes_fixed = 0; P->distinct_nodes = -1; P->last_ms = 0.f;
  P->groups.clear();
  if (n == 0) { P->ngroups = 0; return 0; }
  const int types[4] = {1, 2, 3, 5};
  std::vector<unsigned long long> off((size_t)n);
  long long run = 0;
  int cnt[4] = {0, 0, 0, 0};
  for (int e = 0; e < n; e++)
  {
    off[e] = (unsigned long long)run;
    int nc = ncf_host(type[e]);
    if (nc < 0) { cx_set_error("transfer plan: interpolation type %d is not supported (structured donors: 1,2,3,5)", type[e]); return -1; }
    if (donor[e] < 0 || (long long)donor[e] >= nptsD) { cx_set_error("transfer plan: donor index %d out of range", donor[e]); return -1; }
    if (rcv[e] < 0 || (long long)rcv[e] >= nptsR) { cx_set_error("transfer plan: receptor index %d out of range", rcv[e]); return -1; }
    run += nc;
    for (int g = 0; g < 4; g++) if (type[e] == types[g]) cnt[g]++;
  }
  if (run != ncoefs) { cx_set_error("transfer plan: coefficient array has %lld entries, types need %lld", ncoefs, run); return -1; }
  P->alg_bytes_fixed = 8 * run + 12ll * n;
  cudaStream_t s = ctx->stream;
  const size_t N = n;
  int *r_dn, *r_rc, *r_ty; unsigned long long *r_off, *d_cnt, *d_scan; double* r_cf;
  CX_CUDA(cudaMalloc(&r_dn, sizeof(int) * N)); CX_CUDA(cudaMalloc(&r_rc, sizeof(int) * N));
  CX_CUDA(cudaMalloc(&r_ty, sizeof(int) * N)); CX_CUDA(cudaMalloc(&r_off, sizeof(unsigned long long) * N));
  CX_CUDA(cudaMalloc(&r_cf, sizeof(double) * (size_t)std::max(run, 1ll)));
  const int sorted = plan_sort_mode();
  const size_t nscan = sorted ? (size_t)nptsD : N;
  CX_CUDA(cudaMalloc(&d_cnt, sizeof(unsigned long long) * nscan));
  CX_CUDA(cudaMalloc(&d_scan, sizeof(unsigned long long) * nscan));
  CX_CUDA(cudaMemcpyAsync(r_dn, donor, sizeof(int) * N, cudaMemcpyHostToDevice, s));
  CX_CUDA(cudaMemcpyAsync(r_rc, rcv, sizeof(int) * N, cudaMemcpyHostToDevice, s));
  CX_CUDA(cudaMemcpyAsync(r_ty, type, sizeof(int) * N, cudaMemcpyHostToDevice, s));
  CX_CUDA(cudaMemcpyAsync(r_off, off.data(), sizeof(unsigned long long) * N, cudaMemcpyHostToDevice, s));
  CX_CUDA(cudaMemcpyAsync(r_cf, coefs, sizeof(double) * (size_t)run, cudaMemcpyHostToDevice, s));
  CX_CUDA(cudaMalloc(&P->d_donor, sizeof(int) * N));
  CX_CUDA(cudaMalloc(&P->d_rcv, sizeof(int) * N));
  CX_CUDA(cudaMalloc(&P->d_pos, sizeof(int) * N));
  const int TB = 256;
  int first = 0;
  for (int g = 0; g < 4; g++)
  {
    if (cnt[g] == 0) continue;
    cx_plan::Group G; G.type = types[g]; G.ncf = ncf_host(types[g]); G.sten = sten_host(types[g]);
    G.n = cnt[g]; G.first = first; G.d_cf = nullptr;
    CX_CUDA(cudaMalloc(&G.d_cf, sizeof(double) * (size_t)G.ncf * G.n));
    if (sorted)
    {
      CX_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * nscan, s));
      k_plan_hist<<<cx_grid(n, TB), TB, 0, s>>>(n, G.type, r_ty, r_dn, d_cnt);
      CX_CHECK(cx_exclusive_scan_u64(ctx, d_cnt, d_scan, (int)nscan, nullptr));
    }
    else
    {
      k_plan_flag<<<cx_grid(n, TB), TB, 0, s>>>(n, G.type, r_ty, d_cnt);
      CX_CHECK(cx_exclusive_scan_u64(ctx, d_cnt, d_scan, n, nullptr));
    }
    k_plan_place<<<cx_grid(n, TB), TB, 0, s>>>(n, G.type, G.ncf, G.n, sorted, r_ty, r_dn, r_rc, r_off, r_cf, d_scan,
                                              d_scan, P->d_donor + first, P->d_rcv + first, P->d_pos + first, G.d_cf);
    ctx->launches += 2;
    P->groups.push_back(G);
    first += G.n;
  }
  P->ngroups = (int)P->groups.size();
  CX_CUDA(cudaStreamSynchronize(s));
  CX_CUDA(cudaGetLastError());
  cudaFree(r_dn); cudaFree(r_rc); cudaFree(r_ty); cudaFree(r_off); cudaFree(r_cf); cudaFree(d_cnt); cudaFree(d_scan);
  return 0;
}

// mode 0: in-place scatter into receptor fields; mode 1: dense (nvars, ld) output
int cx_plan_launch(cx_plan* P, int nvars, const double* const* dD, double* const* dR, double* dense, long long ld,
                   int mode, cudaStream_t s)
{
  cx_ctx* ctx = P->ctx;
  const int TB = 256;
  for (int v0 = 0; v0 < nvars; v0 += CX_MAXVARS)
  {
    int nv = std::min(CX_MAXVARS, nvars - v0);
    FieldPtrs F;
    memset(&F, 0, sizeof(F));
    for (int v = 0; v < nv; v++) { F.d[v] = dD[v0 + v]; F.r[v] = dR ? dR[v0 + v] : nullptr; }
    double* dn = dense ? dense + (size_t)v0 * ld : nullptr;
    for (auto& G : P->groups)
    {
      int grid = cx_grid(G.n, TB);
      const int* dn_ = P->d_donor + G.first; const int* rc_ = P->d_rcv + G.first; const int* ps_ = P->d_pos + G.first;
      int imd = P->imd, ij = P->imd * P->jmd;
#define CX_LAUNCH(T)                                                                                           \
      if (mode == 0) k_transfer<T, 0><<<grid, TB, 0, s>>>(G.n, nv, F, imd, ij, dn_, rc_, ps_, G.d_cf, dn, ld); \
      else k_transfer<T, 1><<<grid, TB, 0, s>>>(G.n, nv, F, imd, ij, dn_, rc_, ps_, G.d_cf, dn, ld);
      switch (G.type)
      {
        case 1: CX_LAUNCH(1) break;
        case 2: CX_LAUNCH(2) break;
        case 3: CX_LAUNCH(3) break;
        case 5: CX_LAUNCH(5) break;
      }
#undef CX_LAUNCH
      ctx->launches++;
    }
  }
  return 0;
}

int cx_plan_launch_celln(cx_plan* P, const double* dCellND, double* dCellNR, cudaStream_t s)
{
  const int TB = 256;
  for (auto& G : P->groups)
  {
    int grid = cx_grid(G.n, TB);
    const int* dn_ = P->d_donor + G.first; const int* rc_ = P->d_rcv + G.first;
    int imd = P->imd, ij = P->imd * P->jmd;
    switch (G.type)
    {
      case 1: k_transfer_celln<1><<<grid, TB, 0, s>>>(G.n, dCellND, dCellNR, imd, ij, dn_, rc_, G.d_cf); break;
      case 2: k_transfer_celln<2><<<grid, TB, 0, s>>>(G.n, dCellND, dCellNR, imd, ij, dn_, rc_, G.d_cf); break;
      case 3: k_transfer_celln<3><<<grid, TB, 0, s>>>(G.n, dCellND, dCellNR, imd, ij, dn_, rc_, G.d_cf); break;
      case 5: k_transfer_celln<5><<<grid, TB, 0, s>>>(G.n, dCellND, dCellNR, imd, ij, dn_, rc_, G.d_cf); break;
    }
    P->ctx->launches++;
  }
  return 0;
}
