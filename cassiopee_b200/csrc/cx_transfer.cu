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
#include "../../include/cassiopee_b200.h"
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include <vector>

#define CX_MAXVARS 16

struct FieldPtrs { const double* d[CX_MAXVARS]; double* r[CX_MAXVARS]; };

namespace {

template <int O> struct Sten { };

// One entry of one plan.  MODE 0: scatter to fieldR[v][rcv[e]];  MODE 1: dense out[v*ld + pos[e]].
// PD/PR: anything indexable by variable (kernel-parameter struct or device pointer table).
template <int TYPE, int MODE, class PD, class PR>
__device__ __forceinline__ void transfer_entry(int e, int m, int nvars, const PD& pd, const PR& pr, int imd, int imdjmd,
                                               const int* __restrict__ donor, const int* __restrict__ idx,
                                               const double* __restrict__ cfT, double* __restrict__ dense,
                                               long long ld)
{
  const int d0 = donor[e];
  const size_t o = (size_t)idx[e];
  if (TYPE == 1)
  {
    double c0 = cfT[e];
    for (int v = 0; v < nvars; v++)
    {
      double val = c0 * __ldg(pd[v] + d0);
      if (MODE == 0) pr[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
  }
  else if (TYPE == 2)
  {
    double c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = cfT[(size_t)k * m + e];
    for (int v = 0; v < nvars; v++)
    {
      const double* f = pd[v] + d0;
      double val = 0.;
      val += c[0] * __ldg(f);
      val += c[1] * __ldg(f + 1);
      val += c[2] * __ldg(f + imd);
      val += c[3] * __ldg(f + imd + 1);
      val += c[4] * __ldg(f + imdjmd);
      val += c[5] * __ldg(f + imdjmd + 1);
      val += c[6] * __ldg(f + imdjmd + imd);
      val += c[7] * __ldg(f + imdjmd + imd + 1);
      if (MODE == 0) pr[v][o] = val; else dense[(size_t)v * ld + o] = val;
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
      const double* f = pd[v] + d0;
      double val = 0.;
#pragma unroll 1
      for (int kk = 0; kk < O; kk++)
#pragma unroll
        for (int jj = 0; jj < O; jj++)
        {
          const double* g = f + jj * imd + kk * imdjmd;
#pragma unroll
          for (int ii = 0; ii < O; ii++)
            val += c[ii] * c[jj + O] * c[kk + 2 * O] * __ldg(g + ii);   // left to right, as commonInterpTransfers.h:85,141
        }
      if (MODE == 0) pr[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
  }
}

struct PDk { const double* const* p; __device__ __forceinline__ const double* operator[](int v) const { return p[v]; } };
struct PRk { double* const* p; __device__ __forceinline__ double* operator[](int v) const { return p[v]; } };

template <int TYPE, int MODE>
__global__ void __launch_bounds__(256)
k_transfer(int m, int nvars, FieldPtrs P, int imd, int imdjmd, const int* __restrict__ donor,
           const int* __restrict__ rcv, const int* __restrict__ pos, const double* __restrict__ cfT,
           double* __restrict__ dense, long long ld)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  PDk pd{P.d}; PRk pr{P.r};
  transfer_entry<TYPE, MODE>(e, m, nvars, pd, pr, imd, imdjmd, donor, MODE == 0 ? rcv : pos, cfT, dense, ld);
}

// Plan set: every (plan, type group) of a rank in ONE launch.  Block b works on
// descriptor blk2desc[b]; field pointers come from per-zone device tables.
struct SetDesc {
  int m, type, imd, imdjmd, mode, first_block;
  const int* donor; const int* idx; const double* cfT;
  const double* const* dptr;   // device table: nvars donor field pointers
  double* const* rptr;         // device table: nvars receptor field pointers (mode 0)
  double* dense; long long ld; // mode 1
};

template <int TYPE>
__global__ void __launch_bounds__(256)
k_transfer_set(const SetDesc* __restrict__ descs, const int* __restrict__ blk2desc, int nvars)
{
  const SetDesc D = descs[blk2desc[blockIdx.x]];
  int e = (blockIdx.x - D.first_block) * blockDim.x + threadIdx.x;
  if (e >= D.m) return;
  PDk pd{D.dptr}; PRk pr{D.rptr};
  if (D.mode == 0) transfer_entry<TYPE, 0>(e, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.donor, D.idx, D.cfT, D.dense, D.ld);
  else transfer_entry<TYPE, 1>(e, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.donor, D.idx, D.cfT, D.dense, D.ld);
}

// Receptor-side batched scatter of received dense buffers: fieldR[v][rcv[e]] = in[v*ld + e]
struct ScatDesc { int n, first_block; const int* rcv; const double* in; long long ld; double* const* rptr; };
__global__ void __launch_bounds__(256)
k_scatter_set(const ScatDesc* __restrict__ descs, const int* __restrict__ blk2desc, int nvars)
{
  const ScatDesc D = descs[blk2desc[blockIdx.x]];
  int e = (blockIdx.x - D.first_block) * blockDim.x + threadIdx.x;
  if (e >= D.n) return;
  int r = D.rcv[e];
  for (int v = 0; v < nvars; v++) D.rptr[v][r] = D.in[(size_t)v * D.ld + e];
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
  CX_MALLOC(r_dn, sizeof(int) * N, s); CX_MALLOC(r_rc, sizeof(int) * N, s);
  CX_MALLOC(r_ty, sizeof(int) * N, s); CX_MALLOC(r_off, sizeof(unsigned long long) * N, s);
  CX_MALLOC(r_cf, sizeof(double) * (size_t)std::max(run, 1ll), s);
  const int sorted = plan_sort_mode();
  const size_t nscan = sorted ? (size_t)nptsD : N;
  CX_MALLOC(d_cnt, sizeof(unsigned long long) * nscan, s);
  CX_MALLOC(d_scan, sizeof(unsigned long long) * nscan, s);
  CX_CUDA(cudaMemcpyAsync(r_dn, donor, sizeof(int) * N, cudaMemcpyHostToDevice, s));
  CX_CUDA(cudaMemcpyAsync(r_rc, rcv, sizeof(int) * N, cudaMemcpyHostToDevice, s));
  CX_CUDA(cudaMemcpyAsync(r_ty, type, sizeof(int) * N, cudaMemcpyHostToDevice, s));
  CX_CUDA(cudaMemcpyAsync(r_off, off.data(), sizeof(unsigned long long) * N, cudaMemcpyHostToDevice, s));
  CX_CUDA(cudaMemcpyAsync(r_cf, coefs, sizeof(double) * (size_t)run, cudaMemcpyHostToDevice, s));
  CX_MALLOC(P->d_donor, sizeof(int) * N, s);
  CX_MALLOC(P->d_rcv, sizeof(int) * N, s);
  CX_MALLOC(P->d_pos, sizeof(int) * N, s);
  const int TB = 256;
  int first = 0;
  for (int g = 0; g < 4; g++)
  {
    if (cnt[g] == 0) continue;
    cx_plan::Group G; G.type = types[g]; G.ncf = ncf_host(types[g]); G.sten = sten_host(types[g]);
    G.n = cnt[g]; G.first = first; G.d_cf = nullptr;
    CX_MALLOC(G.d_cf, sizeof(double) * (size_t)G.ncf * G.n, s);
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
  cudaFreeAsync(r_dn, s); cudaFreeAsync(r_rc, s); cudaFreeAsync(r_ty, s); cudaFreeAsync(r_off, s); cudaFreeAsync(r_cf, s); cudaFreeAsync(d_cnt, s); cudaFreeAsync(d_scan, s);
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


// ---------------------------------------------------------------------------
// plan set / scatter set
// ---------------------------------------------------------------------------
struct cx_planset {
  cx_ctx* ctx;
  int nvars, ndesc;
  int nblocks[4];                       // per interpolation type 1,2,3,5
  SetDesc* d_descs[4]; int* d_blk2desc[4];
  void** d_tables; size_t ntables;      // device pointer tables owned by the set
};

struct cx_scatterset {
  cx_ctx* ctx;
  int nvars, nblocks, ndesc;
  ScatDesc* d_descs; int* d_blk2desc;
  std::vector<void*> owned;
};

static int upload_ptr_table(const double* const* host, int nvars, double*** out)
{
  double** d;
  CX_CUDA(cudaMalloc(&d, sizeof(double*) * (size_t)nvars));
  CX_CUDA(cudaMemcpy(d, host, sizeof(double*) * (size_t)nvars, cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

extern "C" {

// modes[p]: 0 in-place into d_fieldR[p] (nvars device pointers), 1 dense into d_dense[p] with leading dim ld[p]
int cx_planset_create(cx_ctx* ctx, int nplans, cx_plan* const* plans, int nvars, const int* modes,
                      const double* const* const* d_fieldD, double* const* const* d_fieldR, double* const* d_dense,
                      const long long* ld, cx_planset** out)
{
  CX_CUDA(cudaSetDevice(ctx->device));
  cx_planset* S = new cx_planset;
  S->ctx = ctx; S->nvars = nvars; S->d_tables = nullptr; S->ntables = 0; S->ndesc = 0;
  std::vector<SetDesc> descs[4]; std::vector<int> b2d[4]; std::vector<void*> tables;
  const int TB = 256;
  int nb[4] = {0, 0, 0, 0};
  for (int p = 0; p < nplans; p++)
  {
    cx_plan* P = plans[p];
    if (P->n == 0) continue;
    double** td = nullptr; double** tr = nullptr;
    CX_CHECK(upload_ptr_table(d_fieldD[p], nvars, &td)); tables.push_back(td);
    if (modes[p] == 0) { CX_CHECK(upload_ptr_table((const double* const*)d_fieldR[p], nvars, &tr)); tables.push_back(tr); }
    for (auto& G : P->groups)
    {
      int t = G.type == 1 ? 0 : G.type == 2 ? 1 : G.type == 3 ? 2 : 3;
      SetDesc D;
      D.m = G.n; D.type = G.type; D.imd = P->imd; D.imdjmd = P->imd * P->jmd; D.mode = modes[p]; D.first_block = nb[t];
      D.donor = P->d_donor + G.first; D.idx = (modes[p] == 0 ? P->d_rcv : P->d_pos) + G.first; D.cfT = G.d_cf;
      D.dptr = td; D.rptr = tr; D.dense = modes[p] == 1 ? d_dense[p] : nullptr; D.ld = modes[p] == 1 ? ld[p] : 0;
      int blocks = cx_grid(G.n, TB);
      for (int b = 0; b < blocks; b++) b2d[t].push_back((int)descs[t].size());
      descs[t].push_back(D);
      nb[t] += blocks;
      S->ndesc++;
    }
  }
  for (int t = 0; t < 4; t++)
  {
    S->nblocks[t] = nb[t]; S->d_descs[t] = nullptr; S->d_blk2desc[t] = nullptr;
    if (nb[t] == 0) continue;
    CX_CUDA(cudaMalloc(&S->d_descs[t], sizeof(SetDesc) * descs[t].size()));
    CX_CUDA(cudaMalloc(&S->d_blk2desc[t], sizeof(int) * b2d[t].size()));
    CX_CUDA(cudaMemcpy(S->d_descs[t], descs[t].data(), sizeof(SetDesc) * descs[t].size(), cudaMemcpyHostToDevice));
    CX_CUDA(cudaMemcpy(S->d_blk2desc[t], b2d[t].data(), sizeof(int) * b2d[t].size(), cudaMemcpyHostToDevice));
  }
  S->ntables = tables.size();
  S->d_tables = (void**)malloc(sizeof(void*) * (tables.size() + 1));
  for (size_t i = 0; i < tables.size(); i++) S->d_tables[i] = tables[i];
  *out = S;
  return 0;
}

static int planset_run_on(cx_planset* S, cudaStream_t s)
{
  if (S->nblocks[0]) { k_transfer_set<1><<<S->nblocks[0], 256, 0, s>>>(S->d_descs[0], S->d_blk2desc[0], S->nvars); S->ctx->launches++; }
  if (S->nblocks[1]) { k_transfer_set<2><<<S->nblocks[1], 256, 0, s>>>(S->d_descs[1], S->d_blk2desc[1], S->nvars); S->ctx->launches++; }
  if (S->nblocks[2]) { k_transfer_set<3><<<S->nblocks[2], 256, 0, s>>>(S->d_descs[2], S->d_blk2desc[2], S->nvars); S->ctx->launches++; }
  if (S->nblocks[3]) { k_transfer_set<5><<<S->nblocks[3], 256, 0, s>>>(S->d_descs[3], S->d_blk2desc[3], S->nvars); S->ctx->launches++; }
  return 0;
}

int cx_planset_run(cx_planset* S) { return planset_run_on(S, S->ctx->stream); }

int cx_planset_destroy(cx_planset* S)
{
  if (!S) return 0;
  for (int t = 0; t < 4; t++) { cudaFree(S->d_descs[t]); cudaFree(S->d_blk2desc[t]); }
  for (size_t i = 0; i < S->ntables; i++) cudaFree(S->d_tables[i]);
  free(S->d_tables);
  delete S;
  return 0;
}

// nsets dense buffers (one per incoming message part): d_in[i] holds (nvars, ld[i]) values for the n[i]
// receptor indices d_rcv[i] of the zone whose nvars field pointers are d_fieldR[i]
int cx_scatterset_create(cx_ctx* ctx, int nsets, int nvars, const int* n, const int* const* d_rcv,
                         const double* const* d_in, const long long* ld, double* const* const* d_fieldR,
                         cx_scatterset** out)
{
  CX_CUDA(cudaSetDevice(ctx->device));
  cx_scatterset* S = new cx_scatterset;
  S->ctx = ctx; S->nvars = nvars; S->d_descs = nullptr; S->d_blk2desc = nullptr;
  std::vector<ScatDesc> descs; std::vector<int> b2d;
  int nb = 0;
  for (int i = 0; i < nsets; i++)
  {
    if (n[i] == 0) continue;
    double** tr = nullptr;
    CX_CHECK(upload_ptr_table((const double* const*)d_fieldR[i], nvars, &tr));
    S->owned.push_back(tr);
    ScatDesc D; D.n = n[i]; D.first_block = nb; D.rcv = d_rcv[i]; D.in = d_in[i]; D.ld = ld[i]; D.rptr = tr;
    int blocks = cx_grid(n[i], 256);
    for (int b = 0; b < blocks; b++) b2d.push_back((int)descs.size());
    descs.push_back(D);
    nb += blocks;
  }
  S->nblocks = nb; S->ndesc = (int)descs.size();
  if (nb > 0)
  {
    CX_CUDA(cudaMalloc(&S->d_descs, sizeof(ScatDesc) * descs.size()));
    CX_CUDA(cudaMalloc(&S->d_blk2desc, sizeof(int) * b2d.size()));
    CX_CUDA(cudaMemcpy(S->d_descs, descs.data(), sizeof(ScatDesc) * descs.size(), cudaMemcpyHostToDevice));
    CX_CUDA(cudaMemcpy(S->d_blk2desc, b2d.data(), sizeof(int) * b2d.size(), cudaMemcpyHostToDevice));
  }
  *out = S;
  return 0;
}

static int scatterset_run_on(cx_scatterset* S, cudaStream_t s)
{
  if (S->nblocks == 0) return 0;
  k_scatter_set<<<S->nblocks, 256, 0, s>>>(S->d_descs, S->d_blk2desc, S->nvars);
  S->ctx->launches++;
  return 0;
}

int cx_scatterset_run(cx_scatterset* S) { return scatterset_run_on(S, S->ctx->stream); }

// ---- one solver iteration of a rank -------------------------------------------------
//   main stream : [wait previous exchange] remote plans (dense blocks -> send buffers) | fork | local plans (in place)
//   comm stream :                                                       wait fork -> NCCL send/recv -> scatter | join
// so the NVLink exchange and the scatter of iteration i overlap the local transfers of iteration i.
extern int cx_exchange_on_stream(cx_ctx* ctx, cudaStream_t st, int npeers, const int* peers, const double* const* d_send,
                                 const long long* sendcounts, double* const* d_recv, const long long* recvcounts);

struct cx_step {
  cx_ctx* ctx; cx_planset* ps_remote; cx_planset* ps_local; cx_scatterset* ss;
  std::vector<int> peers; std::vector<const double*> sp; std::vector<double*> rp;
  std::vector<long long> sc, rc;
  int overlap;
};

int cx_step_create(cx_ctx* ctx, cx_planset* ps_remote, cx_planset* ps_local, cx_scatterset* ss, int npeers,
                   const int* peers, const double* const* d_send, const long long* sendcounts, double* const* d_recv,
                   const long long* recvcounts, cx_step** out)
{
  cx_step* S = new cx_step;
  S->ctx = ctx; S->ps_remote = ps_remote; S->ps_local = ps_local; S->ss = ss; S->overlap = 1;
  for (int p = 0; p < npeers; p++)
  {
    S->peers.push_back(peers[p]); S->sp.push_back(d_send[p]); S->rp.push_back(d_recv[p]);
    S->sc.push_back(sendcounts[p]); S->rc.push_back(recvcounts[p]);
  }
  *out = S;
  return 0;
}

int cx_step_run(cx_step* S, int niter, int overlap)
{
  cx_ctx* c = S->ctx;
  cudaStream_t s = c->stream, cs = overlap ? c->comm_stream : c->stream;
  const bool comm = !S->peers.empty();
  for (int it = 0; it < niter; it++)
  {
    if (comm && overlap) CX_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));   // send buffers free again
    if (S->ps_remote) CX_CHECK(planset_run_on(S->ps_remote, s));
    if (comm)
    {
      if (overlap) { CX_CUDA(cudaEventRecord(c->ev_fork, s)); CX_CUDA(cudaStreamWaitEvent(cs, c->ev_fork, 0)); }
      CX_CHECK(cx_exchange_on_stream(c, cs, (int)S->peers.size(), S->peers.data(), S->sp.data(), S->sc.data(),
                                     S->rp.data(), S->rc.data()));
      if (S->ss) CX_CHECK(scatterset_run_on(S->ss, cs));
      if (overlap) CX_CUDA(cudaEventRecord(c->ev_join, cs));
    }
    if (S->ps_local) CX_CHECK(planset_run_on(S->ps_local, s));
  }
  if (comm && overlap) CX_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));     // join: later work on the main stream sees the scatter
  return 0;
}

int cx_step_destroy(cx_step* S)
{
  if (!S) return 0;
  delete S;
  return 0;
}

int cx_scatterset_destroy(cx_scatterset* S)
{
  if (!S) return 0;
  cudaFree(S->d_descs); cudaFree(S->d_blk2desc);
  for (void* p : S->owned) cudaFree(p);
  delete S;
  return 0;
}

}  // extern "C"
