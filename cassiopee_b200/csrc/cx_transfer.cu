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
#include <cub/device/device_radix_sort.cuh>

#define CX_MAXVARS 16
#define CX_BIG_PLAN (1 << 20)   // entries; plan sets launch such plans on their own
#ifndef CX_SET_MINB
#define CX_SET_MINB 4            // resident CTAs per SM the plan-set kernels are compiled for
#endif
#ifndef CX_SET_VU
#define CX_SEL_PACKED2 102        // build-time code of the packable type-2 entries (cx_plan::Group::sel)
#define CX_SET_VU 1              // variables gathered per pass by the plan-set kernels' direct-gather path
#endif

struct FieldPtrs { const double* d[CX_MAXVARS]; double* r[CX_MAXVARS]; };

namespace {

template <int O> struct Sten { };

// Type-2 coefficients of entry e.  The 8 weights the reference's tetrahedron formula produces take at most 4
// distinct values (InterpData.cpp:621-631: cf0 on the four vertices off the tetrahedron's face, cf0+xi/4 on the
// face, two of those again + yi / + zi), so a plan stores such entries as the 4 distinct doubles (slot-major, like
// the raw layout) + a 16-bit selector word (2 bits per vertex): 34 bytes instead of 64.  The values are the same
// doubles bit for bit, so the arithmetic is untouched.  pat == nullptr: raw group, 8 slots.
__device__ __forceinline__ void load_cf2(double* c, const double* __restrict__ cfT, const unsigned short* __restrict__ pat,
                                         int m, int e)
{
  if (pat == nullptr)
  {
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = cfT[(size_t)k * m + e];
  }
  else
  {
    const double v0 = cfT[e], v1 = cfT[(size_t)m + e], v2 = cfT[(size_t)2 * m + e], v3 = cfT[(size_t)3 * m + e];
    const unsigned p = pat[e];
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
      const unsigned sl = (p >> (2 * k)) & 3u;
      c[k] = (sl & 2u) ? ((sl & 1u) ? v3 : v2) : ((sl & 1u) ? v1 : v0);
    }
  }
}

// One entry of one plan.  MODE 0: scatter to fieldR[v][rcv[e]];  MODE 1: dense out[v*ld + pos[e]].
// PD/PR: anything indexable by variable (kernel-parameter struct or device pointer table).
template <int TYPE, int MODE, class PD, class PR, int VU = 1>
__device__ __forceinline__ void transfer_entry(int e, int m, int nvars, const PD& pd, const PR& pr, int imd, int imdjmd,
                                               const int* __restrict__ donor, const int* __restrict__ idx,
                                               const double* __restrict__ cfT, double* __restrict__ dense,
                                               long long ld, const unsigned short* __restrict__ pat = nullptr)
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
  else if (TYPE == 22)
  {
    // 2-D order 2 (commonInterpTransfers.h:56-73): 4 nodes of one plane
    double c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) c[k] = cfT[(size_t)k * m + e];
    for (int v = 0; v < nvars; v++)
    {
      const double* f = pd[v] + d0;
      double val = 0.;
      val += c[0] * __ldg(f);
      val += c[1] * __ldg(f + 1);
      val += c[2] * __ldg(f + imd);
      val += c[3] * __ldg(f + imd + 1);
      if (MODE == 0) pr[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
  }
  else if (TYPE == 2)
  {
    double c[8];
    load_cf2(c, cfT, pat, m, e);
    int v = 0;
    if (VU == 2)
    {
      // two variables per pass: 16 independent gathers in flight (small, latency-bound plans)
      for (; v + 1 < nvars; v += 2)
      {
        const double* f = pd[v] + d0; const double* g = pd[v + 1] + d0;
        const double f0 = __ldg(f), f1 = __ldg(f + 1), f2 = __ldg(f + imd), f3 = __ldg(f + imd + 1);
        const double f4 = __ldg(f + imdjmd), f5 = __ldg(f + imdjmd + 1), f6 = __ldg(f + imdjmd + imd), f7 = __ldg(f + imdjmd + imd + 1);
        const double g0 = __ldg(g), g1 = __ldg(g + 1), g2 = __ldg(g + imd), g3 = __ldg(g + imd + 1);
        const double g4 = __ldg(g + imdjmd), g5 = __ldg(g + imdjmd + 1), g6 = __ldg(g + imdjmd + imd), g7 = __ldg(g + imdjmd + imd + 1);
        double va = 0., vb = 0.;
        va += c[0] * f0; va += c[1] * f1; va += c[2] * f2; va += c[3] * f3;
        va += c[4] * f4; va += c[5] * f5; va += c[6] * f6; va += c[7] * f7;
        vb += c[0] * g0; vb += c[1] * g1; vb += c[2] * g2; vb += c[3] * g3;
        vb += c[4] * g4; vb += c[5] * g5; vb += c[6] * g6; vb += c[7] * g7;
        if (MODE == 0) { pr[v][o] = va; pr[v + 1][o] = vb; }
        else { dense[(size_t)v * ld + o] = va; dense[(size_t)(v + 1) * ld + o] = vb; }
      }
    }
    for (; v < nvars; v++)
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
    constexpr int O = (TYPE == 3) ? 3 : (TYPE == 44) ? 4 : 5;
    double c[3 * O];
#pragma unroll
    for (int k = 0; k < 3 * O; k++) c[k] = cfT[(size_t)k * m + e];
    for (int v = 0; v < nvars; v++)
    {
      const double* f = pd[v] + d0;
      double val = 0.;
#pragma unroll 1
      for (int kk = 0; kk < O; kk++)
      {
        double ck = c[2 * O];                              // c[kk + 2*O] without indexing registers dynamically
#pragma unroll
        for (int q = 1; q < O; q++) ck = (kk == q) ? c[2 * O + q] : ck;
#pragma unroll
        for (int jj = 0; jj < O; jj++)
        {
          const double* g = f + jj * imd + kk * imdjmd;
#pragma unroll
          for (int ii = 0; ii < O; ii++)
            val += c[ii] * c[jj + O] * ck * __ldg(g + ii);   // left to right, as commonInterpTransfers.h:85,141
        }
      }
      if (MODE == 0) pr[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
  }
}

// ---- shared-memory staging of a tile's donor rows (type 2) -----------------------------------
// The plan's entries are sorted by donor index, so the 256 entries of a CTA read four short contiguous
// row segments of each donor field: rows (j,k), (j+1,k), (j,k+1), (j+1,k+1), nodes [d_lo, d_hi+1].  Instead of
// 8 dependent 8-byte gathers per entry and variable, one elected thread asks the TMA unit for the four
// segments (cp.async.bulk global->shared, completion on an mbarrier), two variables in flight per CTA, and
// the entries then read their 8 stencil values from shared memory.  Every byte of the segment is fetched
// once per CTA, coalesced, with no register or scoreboard cost while it is in flight.  Tiles whose donors
// span more than CX_SPAN nodes (row wraps at the edge of the overlap region, sparse fringes) take the
// direct-gather path.  Arithmetic (operand order, no FMA) is the same on both paths.
// Measured on C2a: tiles of 128 entries 0.576 ms, 256 entries 0.489 ms, 512 entries 0.511 ms; a span buffer of
// 512 nodes (33 KB per CTA, L1 squeezed to 28 KB) 0.583 ms, 384 or 256 nodes 0.49 ms; 3 variables staged per CTA
// (CX_NBUF=3) 0.493 ms, 4: 0.60 ms; the same segments through 16-byte cp.async + commit groups (LSU path, two
// barriers per variable) instead of cp.async.bulk: 0.528 ms; streaming (`__stcs`) receptor stores: 0.4886 vs 0.4887 ms.
#define CX_SPAN 256

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init()
{ asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "CX_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra CX_DONE;\n"
               "bra CX_WAIT;\n"
               "CX_DONE:\n"
               "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

#ifndef CX_NBUF
#define CX_NBUF 2                 // variables staged per CTA (one being read + CX_NBUF-1 in flight)
#endif
struct StageSmem {
  double buf[CX_NBUF][4][CX_SPAN + 4];
  unsigned long long bar[CX_NBUF];
};

// One tile (<= blockDim.x entries starting at t0, all of one plan/type-2 group) of the CTA.
// Returns false (for the whole CTA) if the tile cannot be staged; the caller then gathers directly.
template <int MODE, class PD, class PR>
__device__ __forceinline__ bool transfer_tile2_staged(StageSmem& S, int t0, int m, int nvars, const PD& pd, const PR& pr,
                                                      int imd, int imdjmd, int nptsD, const int* __restrict__ donor,
                                                      const int* __restrict__ idx, const double* __restrict__ cfT,
                                                      double* __restrict__ dense, long long ld, int known_lo = -1,
                                                      int known_hi = -1, const unsigned short* __restrict__ pat = nullptr)
{
  const int tid = threadIdx.x;
  const int e = t0 + tid;
  const int last = min(t0 + (int)blockDim.x, m) - 1;
  // sorted ascending inside a group; plan sets pass the tile's range from their block table
  const int d_lo = known_lo >= 0 ? known_lo : donor[t0], d_hi = known_lo >= 0 ? known_hi : donor[last];
  const int off1 = imd, off2 = imdjmd, off3 = imdjmd + imd;
  // 16-byte aligned segment starts (element index even) and lengths (even), per row
  const int s0 = d_lo & 1, s1 = (d_lo + off1) & 1, s2 = (d_lo + off2) & 1, s3 = (d_lo + off3) & 1;
  const int span = d_hi - d_lo + 2;                      // nodes d_lo .. d_hi+1
  const int len0 = (span + s0 + 1) & ~1, len1 = (span + s1 + 1) & ~1, len2 = (span + s2 + 1) & ~1, len3 = (span + s3 + 1) & ~1;
  if (span + 2 > CX_SPAN || d_lo + off3 - s3 + len3 > nptsD || d_lo + off2 - s2 + len2 > nptsD ||
      d_lo + off1 - s1 + len1 > nptsD || d_lo - s0 + len0 > nptsD)
    return false;
  if (tid == 0) { for (int q = 0; q < CX_NBUF; q++) mbar_init(&S.bar[q], 1); mbar_fence_init(); }
  __syncthreads();
  const unsigned bytes = 8u * (unsigned)(len0 + len1 + len2 + len3);
  auto issue = [&](int v) {
    const int st = v % CX_NBUF;
    const double* f = pd[v];
    mbar_expect_tx(&S.bar[st], bytes);
    bulk_g2s(&S.buf[st][0][0], f + (d_lo - s0), 8u * len0, &S.bar[st]);
    bulk_g2s(&S.buf[st][1][0], f + (d_lo + off1 - s1), 8u * len1, &S.bar[st]);
    bulk_g2s(&S.buf[st][2][0], f + (d_lo + off2 - s2), 8u * len2, &S.bar[st]);
    bulk_g2s(&S.buf[st][3][0], f + (d_lo + off3 - s3), 8u * len3, &S.bar[st]);
  };
  if (tid == 0) for (int q = 0; q < CX_NBUF - 1 && q < nvars; q++) issue(q);
  int l = 0; size_t o = 0;
  double c[8];
  const bool act = e < m;
  if (act)
  {
    l = donor[e] - d_lo; o = (size_t)idx[e];
    load_cf2(c, cfT, pat, m, e);
  }
  for (int v = 0; v < nvars; v++)
  {
    // variable v+NBUF-1 goes into the buffer released by the barrier that ended iteration v-1
    if (tid == 0 && v + CX_NBUF - 1 < nvars) issue(v + CX_NBUF - 1);
    mbar_wait(&S.bar[v % CX_NBUF], (unsigned)((v / CX_NBUF) & 1));
    if (act)
    {
      const double(*b)[CX_SPAN + 4] = S.buf[v % CX_NBUF];
      double val = 0.;
      val += c[0] * b[0][l + s0];
      val += c[1] * b[0][l + s0 + 1];
      val += c[2] * b[1][l + s1];
      val += c[3] * b[1][l + s1 + 1];
      val += c[4] * b[2][l + s2];
      val += c[5] * b[2][l + s2 + 1];
      val += c[6] * b[3][l + s3];
      val += c[7] * b[3][l + s3 + 1];
      if (MODE == 0) pr[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
    __syncthreads();
  }
  return true;
}

// Type 3 (3 x 3 x 3 stencils, direction weights cf[ii]*cf[jj+O]*cf[kk+2O]): the same row staging with 9 row
// segments per variable, two variables in flight, spans up to 256 nodes (0.97 vs 1.07 ms on C2a order 3).  The
// template also instantiates for O = 5 (25 rows, one variable at a time) but type 5 keeps the direct gather:
// staging measured slower there (5.5 vs 3.5 ms).  Accumulation order is commonInterpTransfers.h:85,141.
template <int O> struct StageCfg { };
template <> struct StageCfg<3> { static constexpr int NSTAGE = 2, SPAN = 256; };
template <> struct StageCfg<5> { static constexpr int NSTAGE = 1, SPAN = 160; };
template <int O> struct StageSmemO {
  double buf[StageCfg<O>::NSTAGE][O * O][StageCfg<O>::SPAN + 4];
  unsigned long long bar[2];
};

template <int O, int MODE, class PD, class PR>
__device__ __forceinline__ bool transfer_tileO_staged(StageSmemO<O>& S, int t0, int m, int nvars, const PD& pd,
                                                      const PR& pr, int imd, int imdjmd, int nptsD,
                                                      const int* __restrict__ donor, const int* __restrict__ idx,
                                                      const double* __restrict__ cfT, double* __restrict__ dense,
                                                      long long ld)
{
  constexpr int NSTAGE = StageCfg<O>::NSTAGE, SPAN = StageCfg<O>::SPAN, NROW = O * O;
  const int tid = threadIdx.x;
  const int e = t0 + tid;
  const int last = min(t0 + (int)blockDim.x, m) - 1;
  const int d_lo = donor[t0], d_hi = donor[last];
  const int span = d_hi - d_lo + O;                      // nodes d_lo .. d_hi+O-1 of every row
  const int offmax = (O - 1) * imd + (O - 1) * imdjmd;
  if (span + 2 > SPAN || d_lo + offmax + span + 2 > nptsD) return false;
  if (tid == 0) { mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1); mbar_fence_init(); }
  __syncthreads();
  auto issue = [&](int v) {      // executed by warp 0
    const int st = v % NSTAGE;
    const double* f = pd[v];
    unsigned bytes = 0;
    for (int r = 0; r < NROW; r++) bytes += 8u * (unsigned)((span + ((d_lo + (r % O) * imd + (r / O) * imdjmd) & 1) + 1) & ~1);
    if (tid == 0) mbar_expect_tx(&S.bar[st], bytes);
    __syncwarp();
    for (int r = tid; r < NROW; r += 32)
    {
      const int g = d_lo + (r % O) * imd + (r / O) * imdjmd;
      const int len = (span + (g & 1) + 1) & ~1;
      bulk_g2s(&S.buf[st][r][0], f + (g & ~1), 8u * len, &S.bar[st]);
    }
  };
  if (tid < 32) issue(0);
  int l = 0; size_t o = 0;
  double c[3 * O];
  const bool act = e < m;
  if (act)
  {
    l = donor[e] - d_lo; o = (size_t)idx[e];
#pragma unroll
    for (int k = 0; k < 3 * O; k++) c[k] = cfT[(size_t)k * m + e];
  }
  for (int v = 0; v < nvars; v++)
  {
    if (NSTAGE == 2 && tid < 32 && v + 1 < nvars) issue(v + 1);
    mbar_wait(&S.bar[v % NSTAGE], (unsigned)((v / NSTAGE) & 1));
    if (act)
    {
      const double(*b)[SPAN + 4] = S.buf[v % NSTAGE];
      double val = 0.;
#pragma unroll 1
      for (int kk = 0; kk < O; kk++)
      {
        double ck = c[2 * O];                              // c[kk + 2*O] without indexing registers dynamically
#pragma unroll
        for (int q = 1; q < O; q++) ck = (kk == q) ? c[2 * O + q] : ck;
#pragma unroll
        for (int jj = 0; jj < O; jj++)
        {
          const double* g = &b[jj + O * kk][l + ((d_lo + jj * imd + kk * imdjmd) & 1)];
#pragma unroll
          for (int ii = 0; ii < O; ii++)
            val += c[ii] * c[jj + O] * ck * g[ii];
        }
      }
      if (MODE == 0) pr[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
    __syncthreads();
    if (NSTAGE == 1 && tid < 32 && v + 1 < nvars) issue(v + 1);
  }
  return true;
}

// ---- tiled layout (dense overlaps): entries grouped by donor-cell box, receptor order inside ---------
// When many receptors fall in each donor cell neighbourhood (volume overlaps such as a near-body grid inside
// a background), the donor-sorted layout above coalesces the stencil reads but scatters every 8-byte
// receptor store into its own 32-byte sector: ncu shows the L2 tag pipeline (lts__t_tag_requests) as the
// busiest unit, fed mostly by those partial-sector stores.  The tiled layout keeps BOTH sides local: the
// plan's type-2 entries are grouped by the CX_TX x CX_TY box of donor cells (one k plane) they fall in and
// keep their ascending receptor order inside a box.  A CTA owns one box: the box's (CX_TX+1) x (CX_TY+1) x 2
// donor nodes of every variable arrive in shared memory through cp.async.bulk row segments (one mbarrier),
// the entries then read their stencils from shared memory, and a warp's 32 consecutive receptors are stored
// as full sectors.  Arithmetic is unchanged (same operand order, no FMA).
#define CX_TX 32
#define CX_TY 8
#define CX_ROWP 34                                   // row pitch in shared memory (doubles): CX_TX + 1 nodes + 1 alignment shift
#define CX_BOXV (2 * (CX_TY + 1) * CX_ROWP)          // doubles per variable
struct TileRef { int tile, start; };                 // box id, first entry; entry count = next.start - start

template <int MODE, class PD, class PR>
__device__ __forceinline__ void transfer_box2(double* __restrict__ sm, unsigned long long* bar, int tile, int start,
                                              int count, int m, int nvars, const PD& pd, const PR& pr, int imd, int jmd,
                                              int imdjmd, int nptsD, int ntx, int nty, const int* __restrict__ donor,
                                              const int* __restrict__ idx, const double* __restrict__ cfT,
                                              double* __restrict__ dense, long long ld)
{
  const int tid = threadIdx.x;
  const int tx = tile % ntx, ty = (tile / ntx) % nty, k = tile / (ntx * nty);
  const int i0 = tx * CX_TX, j0 = ty * CX_TY;
  const int ilen = min(CX_TX + 1, imd - i0), jlen = min(CX_TY + 1, jmd - j0);
  const int base = i0 + j0 * imd + k * imdjmd;
  if (base + (jlen - 1) * imd + imdjmd + ilen + 2 > nptsD)
  {
    // the 16-byte rounding of the last row segment could run past the end of the field: gather directly
    for (int c = tid; c < count; c += blockDim.x)
      transfer_entry<2, MODE>(start + c, m, nvars, pd, pr, imd, imdjmd, donor, idx, cfT, dense, ld);
    return;
  }
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (tid < 32)
  {
    const int nseg = 2 * jlen;
    int tot = 0;
    for (int sg = tid; sg < nseg; sg += 32)
    {
      const int g = base + (sg % jlen) * imd + (sg / jlen) * imdjmd;
      tot += (ilen + (g & 1) + 1) & ~1;
    }
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (tid == 0) mbar_expect_tx(bar, 8u * (unsigned)tot * (unsigned)nvars);
    __syncwarp();
    for (int sg = tid; sg < nseg; sg += 32)
    {
      const int r = sg % jlen, p = sg / jlen;
      const int g = base + r * imd + p * imdjmd;
      const int len = (ilen + (g & 1) + 1) & ~1;
      for (int v = 0; v < nvars; v++)
        bulk_g2s(sm + (size_t)v * CX_BOXV + (p * (CX_TY + 1) + r) * CX_ROWP, pd[v] + (g & ~1), 8u * len, bar);
    }
  }
  int c = tid;
  int d0 = 0; size_t o = 0; double cf[8];
  if (c < count)
  {
    const int e = start + c;
    d0 = donor[e]; o = (size_t)idx[e];
#pragma unroll
    for (int q = 0; q < 8; q++) cf[q] = cfT[(size_t)q * m + e];
  }
  mbar_wait(bar, 0u);
  while (c < count)
  {
    const int rel = d0 - base;
    const int rj = rel / imd, ri = rel - rj * imd;
    const int g00 = base + rj * imd;
    const int s00 = g00 & 1, s01 = (g00 + imd) & 1, s10 = (g00 + imdjmd) & 1, s11 = (g00 + imdjmd + imd) & 1;
    const double* b0 = sm + rj * CX_ROWP + ri;
    for (int v = 0; v < nvars; v++)
    {
      const double* b = b0 + (size_t)v * CX_BOXV;
      const double* b1 = b + (CX_TY + 1) * CX_ROWP;
      double val = 0.;
      val += cf[0] * b[s00];
      val += cf[1] * b[s00 + 1];
      val += cf[2] * b[CX_ROWP + s01];
      val += cf[3] * b[CX_ROWP + s01 + 1];
      val += cf[4] * b1[s10];
      val += cf[5] * b1[s10 + 1];
      val += cf[6] * b1[CX_ROWP + s11];
      val += cf[7] * b1[CX_ROWP + s11 + 1];
      if (MODE == 0) pr[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
    c += blockDim.x;
    if (c < count)
    {
      const int e = start + c;
      d0 = donor[e]; o = (size_t)idx[e];
#pragma unroll
      for (int q = 0; q < 8; q++) cf[q] = cfT[(size_t)q * m + e];
    }
  }
}

// ---- box layout (layout 3): entries grouped by a 3-D box of donor cells sized for the plan's shape -----
// Fringe plans (the 2-layer overlap slabs of a multi-block case such as C4) reference a donor region that is
// only 3-4 nodes thick in one direction.  Sorted by donor index, a CTA's 256 entries then span ~100 short rows
// and every entry gathers its 8 corners for every variable with dependent 8-byte loads (ncu, round 1: 44.7
// long-scoreboard stalls per issue, issue slots 12 % busy).  Here the type-2 entries of a plan are grouped by
// boxes of TI x TJ x TK donor cells chosen from the bounding box of the plan's donor cells (thin directions
// whole, the rest split so that a box holds ~256 cells and at most CX_BOX_NODES nodes), in ascending receptor
// order inside a box.  A CTA owns one box: all (TI+1)(TJ+1)(TK+1) nodes of EVERY variable are requested at
// once with 8-byte cp.async (LDGSTS: no register staging, one memory round trip for the whole box while the
// threads fetch their entries' indices and coefficients), each node once instead of once per referencing
// entry, and the stencils are then read from shared memory (row pitch odd: conflict-free for thin rows).
// cp.async.bulk.tensor is not usable here: a tensor map needs global strides that are multiples of 16 bytes
// and the donor rows of a centre mesh of a 100^3 block are 99 * 8 = 792 bytes apart; 1-D bulk copies need
// 16-byte aligned sources, which rows 3-4 nodes long at odd pitch do not have.
// Arithmetic (operand order, no FMA) is the same as on every other path.
#define CX_BOX_NODES 768                              // shared-memory doubles per variable and box (pitch included)
#define CX_BOX_CELLS 256                              // target donor cells per box
struct BoxTile { int start, base, dims, pad_; };      // first entry; donor node index of the box corner; node counts i | j<<8 | k<<16

__device__ __forceinline__ void cp_async8(double* dst, const double* src)
{ asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all()
{ asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int MODE, class PD, class PR>
__device__ __forceinline__ void transfer_box3(double* __restrict__ sm, int start, int count, int base, int dims, int m,
                                              int nvars, const PD& pd, const PR& pr, int imd, int imdjmd,
                                              const int* __restrict__ donor, const int* __restrict__ idx,
                                              const double* __restrict__ cfT, double* __restrict__ dense, long long ld,
                                              const unsigned short* __restrict__ pat = nullptr)
{
  const int tid = threadIdx.x;
  const int ilen = dims & 255, jlen = (dims >> 8) & 255, klen = dims >> 16;
  const int ip = ilen | 1;                             // odd row pitch
  const int pk = ip * jlen;                            // plane pitch
  const int vs = pk * klen;                            // doubles per variable
  const int ij = ilen * jlen, nn = ij * klen;
  for (int q = tid; q < nn; q += blockDim.x)
  {
    const int lk = q / ij, r = q - lk * ij, lj = r / ilen, li = r - lj * ilen;
    const size_t g = (size_t)base + li + (size_t)lj * imd + (size_t)lk * imdjmd;
    double* d = sm + lk * pk + lj * ip + li;
    for (int v = 0; v < nvars; v++) cp_async8(d + (size_t)v * vs, pd[v] + g);
  }
  int c = tid;
  int d0 = 0; size_t o = 0; double cf[8];
  if (c < count)
  {
    const int e = start + c;
    d0 = donor[e]; o = (size_t)idx[e];
    load_cf2(cf, cfT, pat, m, e);
  }
  cp_async_wait_all();
  __syncthreads();
  while (c < count)
  {
    const int rel = d0 - base;
    const int lk = rel / imdjmd, r = rel - lk * imdjmd, lj = r / imd, li = r - lj * imd;
    const double* b = sm + lk * pk + lj * ip + li;
    for (int v = 0; v < nvars; v++, b += vs)
    {
      double val = 0.;
      val += cf[0] * b[0];
      val += cf[1] * b[1];
      val += cf[2] * b[ip];
      val += cf[3] * b[ip + 1];
      val += cf[4] * b[pk];
      val += cf[5] * b[pk + 1];
      val += cf[6] * b[pk + ip];
      val += cf[7] * b[pk + ip + 1];
      if (MODE == 0) pr[v][o] = val; else dense[(size_t)v * ld + o] = val;
    }
    c += blockDim.x;
    if (c < count)
    {
      const int e = start + c;
      d0 = donor[e]; o = (size_t)idx[e];
      load_cf2(cf, cfT, pat, m, e);
    }
  }
}

struct PDk { const double* const* p; __device__ __forceinline__ const double* operator[](int v) const { return p[v]; } };
struct PRk { double* const* p; __device__ __forceinline__ double* operator[](int v) const { return p[v]; } };

template <int TYPE, int MODE>
__global__ void __launch_bounds__(256)
k_transfer(int m, int nvars, FieldPtrs P, int imd, int imdjmd, const int* __restrict__ donor,
           const int* __restrict__ rcv, const int* __restrict__ pos, const double* __restrict__ cfT,
           double* __restrict__ dense, long long ld, const unsigned short* __restrict__ pat)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  PDk pd{P.d}; PRk pr{P.r};
  transfer_entry<TYPE, MODE>(e, m, nvars, pd, pr, imd, imdjmd, donor, MODE == 0 ? rcv : pos, cfT, dense, ld, pat);
}

// type 2 with shared-memory staging (falls back to the direct gather tile by tile)
template <int MODE>
__global__ void __launch_bounds__(256)
k_transfer2s(int m, int nvars, FieldPtrs P, int imd, int imdjmd, int nptsD, const int* __restrict__ donor,
             const int* __restrict__ rcv, const int* __restrict__ pos, const double* __restrict__ cfT,
             double* __restrict__ dense, long long ld, const unsigned short* __restrict__ pat)
{
  extern __shared__ __align__(16) double dyn_sm[];
  StageSmem& S = *reinterpret_cast<StageSmem*>(dyn_sm);
  PDk pd{P.d}; PRk pr{P.r};
  const int t0 = blockIdx.x * blockDim.x;
  if (transfer_tile2_staged<MODE>(S, t0, m, nvars, pd, pr, imd, imdjmd, nptsD, donor, MODE == 0 ? rcv : pos, cfT, dense, ld,
                                  -1, -1, pat))
    return;
  int e = t0 + threadIdx.x;
  if (e >= m) return;
  transfer_entry<2, MODE>(e, m, nvars, pd, pr, imd, imdjmd, donor, MODE == 0 ? rcv : pos, cfT, dense, ld, pat);
}

// Measured and rejected (round 2, C2a, 17.0 M receptors, 5 variables; gpurun_out/r2_t2v*.out): an all-TMA variant of
// the kernel above -- the sorted group cut into variable tiles (<= 256 entries AND <= 192 donor nodes, so that every
// tile stages), indices / selector words / coefficient slots / the four row segments of EVERY variable requested at
// once with cp.async.bulk (two round trips per CTA instead of 2 + nvars), one mbarrier per variable -- ran 0.549 ms
// against 0.499 ms (raw 8-slot coefficients: 0.66 ms; tiles of <= 128 nodes: 0.65 ms; the issue spread over the
// CTA's warps: 0.577 ms).  The time follows the NUMBER of bulk copies (27-31 per tile of ~1-2 KB each, i.e. one copy
// per 60-80 SM cycles), not the bytes: small non-tensor bulk copies are bound by the TMA unit's per-request cost, so
// the plan data stays on plain coalesced loads and only the row segments (20 copies per tile) go through TMA.

// ---- persistent, software-pipelined type-2 kernel (donor-sorted groups; EXPERIMENTAL, CX_TRANSFER_PIPE=1) ----
// Measured on C2a (17.0 M receptors, B200): 5 variables 0.62 ms (3 stages, 2-3 consumer groups) against 0.49 ms
// for k_transfer2s; 7 variables 0.88 ms (only 2 stages fit) against 0.67 ms.  With the producer on a consumer
// warp (first version) 0.83-0.95 ms: the issue of ~30 bulk copies per tile sat on every iteration's critical
// path.  One CTA per SM keeps at most 2 tiles (120 KB) in flight, less than the 6 resident CTAs of the
// per-tile kernel do with plain loads + 2 staged variables, so the per-tile kernel stays the default.
// k_transfer2s hides the latency of a tile's plan data (indices, 8 coefficient slots) and of its first donor
// rows only behind OTHER resident CTAs.  Here one CTA per SM walks over its tiles and a single elected thread
// keeps the NEXT tiles in flight: everything a tile needs -- donor and receptor indices, the 8 coefficient
// slots, and the four donor row segments of every variable -- arrives in shared memory through cp.async.bulk
// copies that complete on the stage's mbarrier, one to two tiles ahead of the arithmetic (2-3 stages of
// 2.1 + 16.5 + 8.3*nvars KB).  No register or scoreboard is tied up while the bytes are in flight, and the
// consumers only ever wait on shared memory.  Source addresses are rounded down to 16 bytes (the element
// shift is applied when reading).  Tiles whose donors span more than CX_SPAN nodes stage their plan data
// only and gather the fields directly.  Arithmetic (operand order, no FMA) is unchanged.
#define CX_PIPE_MAXV 8
#define CX_PIPE_IDX 264                       // ints per staged index list (256 + shift + rounding)
#define CX_PIPE_CF 258                        // doubles per staged coefficient slot
#define CX_PIPE_ROW (CX_SPAN + 4)             // doubles per staged row segment
struct PipeRange { int d_lo, d_hi; };         // donor range of a tile; d_lo < 0: rows not stageable

__host__ __device__ inline size_t pipe_stage_bytes(int nvars)
{
  return 2 * sizeof(int) * CX_PIPE_IDX + 8 * sizeof(double) * CX_PIPE_CF + (size_t)nvars * 4 * sizeof(double) * CX_PIPE_ROW;
}

__global__ void k_pipe_ranges(int ntiles, int m, int imd, int imdjmd, int nptsD, const int* __restrict__ donor,
                              PipeRange* __restrict__ out)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const int t0 = t * 256, last = min(t0 + 256, m) - 1;
  int d_lo = donor[t0], d_hi = donor[last];
  const int span = d_hi - d_lo + 2;
  // same conditions as transfer_tile2_staged: span fits, rounded segments stay inside the field
  if (span + 2 > CX_SPAN || d_lo + imdjmd + imd + span + 2 > nptsD) d_lo = -1;
  out[t].d_lo = d_lo; out[t].d_hi = d_hi;
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// blockDim.x = 256 * NG + 32: NG consumer groups of 256 threads (thread = entry of the tile, the variables are
// dealt round-robin to the groups) and one producer warp that never computes, so that issuing the next tiles
// is off the consumers' critical path.  full[st]: bulk copies landed (tx count); empty[st]: every consumer warp
// is done with the stage.  No CTA-wide barrier inside the loop.
template <int MODE>
__global__ void __launch_bounds__(800, 1)
k_transfer2p(int m, int nvars, FieldPtrs P, int imd, int imdjmd, const int* __restrict__ donor,
             const int* __restrict__ idx, const double* __restrict__ cfT, const PipeRange* __restrict__ ranges,
             int ntiles, int nst, double* __restrict__ dense, long long ld)
{
  extern __shared__ __align__(128) unsigned char psm[];
  unsigned long long* full = reinterpret_cast<unsigned long long*>(psm);         // [nst]
  unsigned long long* empty = full + 4;                                           // [nst]
  PipeRange* rng = reinterpret_cast<PipeRange*>(psm + 64);                        // [nst] range of the staged tile
  const size_t stage_bytes = pipe_stage_bytes(nvars);
  const int tid = threadIdx.x;
  const int ncons = (int)blockDim.x - 32;            // consumer threads
  const int off1 = imd, off2 = imdjmd, off3 = imdjmd + imd;
  if (tid == 0)
  {
    for (int q = 0; q < nst; q++) { mbar_init(&full[q], 1); mbar_init(&empty[q], ncons >> 5); }
    mbar_fence_init();
  }
  __syncthreads();

  if (tid >= ncons)
  {
    // ---- producer warp (one lane issues) ----
    if (tid == ncons)
    {
      int it = 0;
      PipeRange Rn = ranges[blockIdx.x];             // range of the next tile to issue, fetched one tile ahead
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++)
      {
        const int st = it % nst;
        const PipeRange R = Rn;
        if (tile + (int)gridDim.x < ntiles) Rn = ranges[tile + gridDim.x];
        if (it >= nst) mbar_wait(&empty[st], (unsigned)(((it / nst) - 1) & 1));
        unsigned char* sp = psm + 128 + (size_t)st * stage_bytes;
        int* dS = reinterpret_cast<int*>(sp);
        int* iS = dS + CX_PIPE_IDX;
        double* cS = reinterpret_cast<double*>(sp + 2 * sizeof(int) * CX_PIPE_IDX);
        double* rS = cS + 8 * CX_PIPE_CF;
        const int t0 = tile * 256, cnt = min(256, m - t0);
        rng[st] = R;
        const int sd = (int)((reinterpret_cast<uintptr_t>(donor + t0) >> 2) & 3), nd = (cnt + sd + 3) & ~3;
        const int si = (int)((reinterpret_cast<uintptr_t>(idx + t0) >> 2) & 3), ni = (cnt + si + 3) & ~3;
        unsigned bytes = 4u * (unsigned)(nd + ni);
        for (int k = 0; k < 8; k++) bytes += 8u * (unsigned)((cnt + (int)(((size_t)k * m + t0) & 1) + 1) & ~1);
        int len[4] = {0, 0, 0, 0}, g[4] = {0, 0, 0, 0};
        if (R.d_lo >= 0)
        {
          const int span = R.d_hi - R.d_lo + 2;
          g[0] = R.d_lo; g[1] = R.d_lo + off1; g[2] = R.d_lo + off2; g[3] = R.d_lo + off3;
          for (int r = 0; r < 4; r++) { len[r] = (span + (g[r] & 1) + 1) & ~1; bytes += 8u * (unsigned)len[r] * (unsigned)nvars; }
        }
        mbar_expect_tx(&full[st], bytes);
        bulk_g2s(dS, donor + t0 - sd, 4u * nd, &full[st]);
        bulk_g2s(iS, idx + t0 - si, 4u * ni, &full[st]);
        for (int k = 0; k < 8; k++)
        {
          const size_t o = (size_t)k * m + t0;
          const int sk = (int)(o & 1);
          bulk_g2s(cS + k * CX_PIPE_CF, cfT + (o - sk), 8u * (unsigned)((cnt + sk + 1) & ~1), &full[st]);
        }
        if (R.d_lo >= 0)
          for (int v = 0; v < nvars; v++)
            for (int r = 0; r < 4; r++)
              bulk_g2s(rS + (size_t)(v * 4 + r) * CX_PIPE_ROW, P.d[v] + (g[r] & ~1), 8u * (unsigned)len[r], &full[st]);
      }
    }
    return;
  }

  // ---- consumers ----
  const int el = tid & 255, vg = tid >> 8, nvg = ncons >> 8;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++)
  {
    const int st = it % nst;
    mbar_wait(&full[st], (unsigned)((it / nst) & 1));
    const int t0 = tile * 256, cnt = min(256, m - t0);
    if (el < cnt && vg < nvars)
    {
      const unsigned char* sp = psm + 128 + (size_t)st * stage_bytes;
      const int* dS = reinterpret_cast<const int*>(sp);
      const int* iS = dS + CX_PIPE_IDX;
      const double* cS = reinterpret_cast<const double*>(sp + 2 * sizeof(int) * CX_PIPE_IDX);
      const double* rS = cS + 8 * CX_PIPE_CF;
      const PipeRange R = rng[st];
      const int sd = (int)((reinterpret_cast<uintptr_t>(donor + t0) >> 2) & 3);
      const int si = (int)((reinterpret_cast<uintptr_t>(idx + t0) >> 2) & 3);
      const int d0 = dS[el + sd];
      const size_t o = (size_t)iS[el + si];
      double c[8];
#pragma unroll
      for (int k = 0; k < 8; k++) c[k] = cS[k * CX_PIPE_CF + el + (int)(((size_t)k * m + t0) & 1)];
      if (R.d_lo >= 0)
      {
        const int l = d0 - R.d_lo;
        const int l0 = l + (R.d_lo & 1), l1 = l + ((R.d_lo + off1) & 1) + CX_PIPE_ROW;
        const int l2 = l + ((R.d_lo + off2) & 1) + 2 * CX_PIPE_ROW, l3 = l + ((R.d_lo + off3) & 1) + 3 * CX_PIPE_ROW;
        for (int v = vg; v < nvars; v += nvg)
        {
          const double* b = rS + (size_t)v * 4 * CX_PIPE_ROW;
          double val = 0.;
          val += c[0] * b[l0];
          val += c[1] * b[l0 + 1];
          val += c[2] * b[l1];
          val += c[3] * b[l1 + 1];
          val += c[4] * b[l2];
          val += c[5] * b[l2 + 1];
          val += c[6] * b[l3];
          val += c[7] * b[l3 + 1];
          if (MODE == 0) P.r[v][o] = val; else dense[(size_t)v * ld + o] = val;
        }
      }
      else
      {
        for (int v = vg; v < nvars; v += nvg)
        {
          const double* f = P.d[v] + d0;
          double val = 0.;
          val += c[0] * __ldg(f);
          val += c[1] * __ldg(f + 1);
          val += c[2] * __ldg(f + off1);
          val += c[3] * __ldg(f + off1 + 1);
          val += c[4] * __ldg(f + off2);
          val += c[5] * __ldg(f + off2 + 1);
          val += c[6] * __ldg(f + off3);
          val += c[7] * __ldg(f + off3 + 1);
          if (MODE == 0) P.r[v][o] = val; else dense[(size_t)v * ld + o] = val;
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[st]);     // this warp no longer reads the stage
  }
}

// types 3 / 5 with shared-memory staging (falls back to the direct gather tile by tile)
template <int O, int MODE>
__global__ void __launch_bounds__(256)
k_transferOs(int m, int nvars, FieldPtrs P, int imd, int imdjmd, int nptsD, const int* __restrict__ donor,
             const int* __restrict__ rcv, const int* __restrict__ pos, const double* __restrict__ cfT,
             double* __restrict__ dense, long long ld)
{
  extern __shared__ __align__(16) double dyn_sm[];
  StageSmemO<O>& S = *reinterpret_cast<StageSmemO<O>*>(dyn_sm);
  PDk pd{P.d}; PRk pr{P.r};
  const int t0 = blockIdx.x * blockDim.x;
  if (transfer_tileO_staged<O, MODE>(S, t0, m, nvars, pd, pr, imd, imdjmd, nptsD, donor, MODE == 0 ? rcv : pos, cfT, dense, ld))
    return;
  int e = t0 + threadIdx.x;
  if (e >= m) return;
  transfer_entry<O, MODE>(e, m, nvars, pd, pr, imd, imdjmd, donor, MODE == 0 ? rcv : pos, cfT, dense, ld);
}

// type 2, tiled layout: one CTA per non-empty donor box
template <int MODE>
__global__ void __launch_bounds__(256)
k_transfer2t(const TileRef* __restrict__ tiles, int m, int nvars, FieldPtrs P, int imd, int jmd, int nptsD, int ntx,
             int nty, const int* __restrict__ donor, const int* __restrict__ rcv, const int* __restrict__ pos,
             const double* __restrict__ cfT, double* __restrict__ dense, long long ld)
{
  extern __shared__ __align__(16) double dyn_sm[];
  __shared__ unsigned long long bar;
  PDk pd{P.d}; PRk pr{P.r};
  const TileRef t = tiles[blockIdx.x];
  const int count = tiles[blockIdx.x + 1].start - t.start;
  transfer_box2<MODE>(dyn_sm, &bar, t.tile, t.start, count, m, nvars, pd, pr, imd, jmd, imd * jmd, nptsD, ntx, nty, donor,
                      MODE == 0 ? rcv : pos, cfT, dense, ld);
}

// type 2, box layout: one CTA per non-empty donor box
template <int MODE>
__global__ void __launch_bounds__(256)
k_transfer2b(const BoxTile* __restrict__ tiles, int m, int nvars, FieldPtrs P, int imd, int imdjmd,
             const int* __restrict__ donor, const int* __restrict__ rcv, const int* __restrict__ pos,
             const double* __restrict__ cfT, double* __restrict__ dense, long long ld,
             const unsigned short* __restrict__ pat)
{
  extern __shared__ __align__(16) double dyn_sm[];
  PDk pd{P.d}; PRk pr{P.r};
  const BoxTile t = tiles[blockIdx.x];
  const int count = tiles[blockIdx.x + 1].start - t.start;
  transfer_box3<MODE>(dyn_sm, t.start, count, t.base, t.dims, m, nvars, pd, pr, imd, imdjmd, donor, MODE == 0 ? rcv : pos,
                      cfT, dense, ld, pat);
}

// Plan set: every (plan, type group) of a rank in ONE launch.  Block b works on
// descriptor blk2desc[b]; field pointers come from per-zone device tables.
struct SetDesc {
  int m, type, imd, imdjmd, mode, first_block, nptsD;
  int layout, jmd, ntx, nty;                 // layout 2: tiled, blocks = boxes
  int pad_;
  const TileRef* tiles;
  const int* donor; const int* idx; const double* cfT;
  const unsigned short* pat;   // packed type-2 group: selector words (else nullptr)
  double* dense; long long ld; // mode 1
  const double* dp[CX_MAXVARS];  // donor field pointers (inline: no dependent table fetch)
  double* rp[CX_MAXVARS];        // receptor field pointers (mode 0)
};
static_assert(sizeof(SetDesc) % 8 == 0, "SetDesc is copied to shared memory in 8-byte words");
// one record per block of a plan-set launch: its descriptor and, for donor-sorted type-2 groups, the donor
// index range of the block's tile (so the staging can start without reading the donor list first)
struct BlkRec { int desc, d_lo, d_hi, start, count, base, dims, pad_; };   // start..dims: box layout (layout 3) blocks

template <int TYPE, bool TILED>
__global__ void __launch_bounds__(256, TILED ? 2 : CX_SET_MINB)
k_transfer_set(const SetDesc* __restrict__ descs, const BlkRec* __restrict__ blk, int nvars, int staged, int v0)
{
  // the descriptor (with its field-pointer tables) goes through shared memory once per CTA, copied by 8-byte
  // words in one round trip: the per-variable pointer fetches inside the entry loop are shared-memory reads
  __shared__ __align__(16) SetDesc sD;
  const BlkRec br = blk[blockIdx.x];
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(descs + br.desc);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&sD);
    if (threadIdx.x < sizeof(SetDesc) / 8) dst[threadIdx.x] = src[threadIdx.x];
  }
  __syncthreads();
  if (v0)
  {
    // a sub-range [v0, v0 + nvars) of the set's variables (per-variable pipeline of a host-to-host step): the dense
    // block of a remote plan is (nvars_of_the_set, ld), variable-major
    if (threadIdx.x == 0 && sD.dense) sD.dense += (size_t)v0 * sD.ld;
    __syncthreads();
  }
  const SetDesc& D = sD;
  PDk pd{sD.dp + v0}; PRk pr{sD.rp + v0};
  if (TYPE == 2 && D.type == 1)
  {
    // coincident (type 1) entries ride in the type-2 launch: small plans are launch-latency bound, one launch less
    int e1 = (blockIdx.x - D.first_block) * blockDim.x + threadIdx.x;
    if (e1 >= D.m) return;
    if (D.mode == 0) transfer_entry<1, 0, PDk, PRk, 1>(e1, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.donor, D.idx, D.cfT, D.dense, D.ld);
    else transfer_entry<1, 1, PDk, PRk, 1>(e1, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.donor, D.idx, D.cfT, D.dense, D.ld);
    return;
  }
  if (TYPE == 2)
  {
    extern __shared__ __align__(16) double dyn_sm[];
    __shared__ unsigned long long bar;
    if (TILED && D.layout == 2)
    {
      const int b = blockIdx.x - D.first_block;
      const TileRef t = D.tiles[b];
      const int count = D.tiles[b + 1].start - t.start;
      if (staged)
      {
        if (D.mode == 0) transfer_box2<0>(dyn_sm, &bar, t.tile, t.start, count, D.m, nvars, pd, pr, D.imd, D.jmd, D.imdjmd, D.nptsD, D.ntx, D.nty, D.donor, D.idx, D.cfT, D.dense, D.ld);
        else transfer_box2<1>(dyn_sm, &bar, t.tile, t.start, count, D.m, nvars, pd, pr, D.imd, D.jmd, D.imdjmd, D.nptsD, D.ntx, D.nty, D.donor, D.idx, D.cfT, D.dense, D.ld);
      }
      else
        for (int c = threadIdx.x; c < count; c += blockDim.x)
        {
          if (D.mode == 0) transfer_entry<2, 0>(t.start + c, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.donor, D.idx, D.cfT, D.dense, D.ld);
          else transfer_entry<2, 1>(t.start + c, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.donor, D.idx, D.cfT, D.dense, D.ld);
        }
      return;
    }
    if (D.layout == 3)
    {
      if (D.mode == 0) transfer_box3<0>(dyn_sm, br.start, br.count, br.base, br.dims, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.donor, D.idx, D.cfT, D.dense, D.ld, D.pat);
      else transfer_box3<1>(dyn_sm, br.start, br.count, br.base, br.dims, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.donor, D.idx, D.cfT, D.dense, D.ld, D.pat);
      return;
    }
    if (staged && D.layout == 1)
    {
      StageSmem& S = *reinterpret_cast<StageSmem*>(dyn_sm);
      const int t0 = (blockIdx.x - D.first_block) * blockDim.x;
      bool done = (D.mode == 0)
        ? transfer_tile2_staged<0>(S, t0, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.nptsD, D.donor, D.idx, D.cfT, D.dense, D.ld, br.d_lo, br.d_hi, D.pat)
        : transfer_tile2_staged<1>(S, t0, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.nptsD, D.donor, D.idx, D.cfT, D.dense, D.ld, br.d_lo, br.d_hi, D.pat);
      if (done) return;
    }
  }
  if (TYPE == 3)
  {
    constexpr int O = 3;
    extern __shared__ __align__(16) double dyn_sm[];
    if (staged && D.layout == 1)
    {
      StageSmemO<O>& S = *reinterpret_cast<StageSmemO<O>*>(dyn_sm);
      const int t0 = (blockIdx.x - D.first_block) * blockDim.x;
      bool done = (D.mode == 0)
        ? transfer_tileO_staged<O, 0>(S, t0, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.nptsD, D.donor, D.idx, D.cfT, D.dense, D.ld)
        : transfer_tileO_staged<O, 1>(S, t0, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.nptsD, D.donor, D.idx, D.cfT, D.dense, D.ld);
      if (done) return;
    }
  }
  int e = (blockIdx.x - D.first_block) * blockDim.x + threadIdx.x;
  if (e >= D.m) return;
  if (D.mode == 0) transfer_entry<TYPE, 0, PDk, PRk, CX_SET_VU>(e, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.donor, D.idx, D.cfT, D.dense, D.ld, D.pat);
  else transfer_entry<TYPE, 1, PDk, PRk, CX_SET_VU>(e, D.m, nvars, pd, pr, D.imd, D.imdjmd, D.donor, D.idx, D.cfT, D.dense, D.ld, D.pat);
}

// Measured and rejected (round 2, C4 step, 8 zones of 100^3, 7 variables; gpurun_out/r2_lean.out): a lean variant of the
// kernel above for the type-1 / type-2 table -- thread = entry, direct gathers only, no shared memory, the descriptor
// read straight from global memory, 40 registers = 6 resident CTAs per SM instead of 4 -- ran 56.7 us against 56.0 us
// (8 CTAs per SM, with spills: 62.6 us).  Half again as many resident warps change nothing: the step is not bound by
// latency x occupancy but by the sector-granular traffic of the fringe slabs (188 MB of DRAM traffic for 108 MB of
// algorithmic bytes, 1.45 M partial-sector receptor stores that each cost a fetch-on-write; profiles/r02_*).

// donor index range of every block's tile (donor-sorted type-2 groups of a plan set), once at creation
__global__ void k_block_ranges(int nblocks, const SetDesc* __restrict__ descs, BlkRec* __restrict__ blk, int tb)
{
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const SetDesc& D = descs[blk[b].desc];
  if (D.type == 2 && D.layout == 3)
  {
    const BoxTile* T = reinterpret_cast<const BoxTile*>(D.tiles) + (b - D.first_block);
    blk[b].start = T[0].start; blk[b].count = T[1].start - T[0].start; blk[b].base = T[0].base; blk[b].dims = T[0].dims;
    return;
  }
  if (D.type != 2 || D.layout != 1) return;
  const int t0 = (b - D.first_block) * tb;
  const int last = min(t0 + tb, D.m) - 1;
  blk[b].d_lo = D.donor[t0]; blk[b].d_hi = D.donor[last];
}

// Receptor-side batched scatter of received dense buffers: fieldR[v][rcv[e]] = in[v*ld + e]
struct ScatDesc { int n, first_block; const int* rcv; const double* in; long long ld; double* const* rptr; };
__global__ void __launch_bounds__(256)
k_scatter_set(const ScatDesc* __restrict__ descs, const int* __restrict__ blk2desc, int v0, int nvars)
{
  const ScatDesc D = descs[blk2desc[blockIdx.x]];
  int e = (blockIdx.x - D.first_block) * blockDim.x + threadIdx.x;
  if (e >= D.n) return;
  int r = D.rcv[e];
  for (int v = v0; v < v0 + nvars; v++) D.rptr[v][r] = D.in[(size_t)v * D.ld + e];
}

// strict cellN rule (commonCellNTransfersStrict.h): product of cellN*(2-cellN)
// over stencil nodes whose coefficient exceeds 1e-11; > 0.5 -> 1 else 0
template <int TYPE>
__global__ void k_transfer_celln(int m, const double* __restrict__ cellND, double* __restrict__ cellNR, int imd,
                                 int imdjmd, const int* __restrict__ donor, const int* __restrict__ rcv,
                                 const double* __restrict__ cfT, const unsigned short* __restrict__ pat = nullptr)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int d0 = donor[e];
  double val = 1.;
  if (TYPE == 1) { double c = cellND[d0]; val = c * (2. - c); }
  else if (TYPE == 22)
  {
    for (int jj = 0; jj < 2; jj++) for (int ii = 0; ii < 2; ii++)
    {
      double cf = cfT[(size_t)(ii + 2 * jj) * m + e];
      if (fabs(cf) > 1.e-11) { double c = cellND[d0 + ii + jj * imd]; val *= c * (2. - c); }
    }
  }
  else if (TYPE == 2)
  {
    double c8[8];
    load_cf2(c8, cfT, pat, m, e);
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
      const int ii = q & 1, jj = (q >> 1) & 1, kk = q >> 2;
      double cf = c8[q];
      if (fabs(cf) > 1.e-11) { double c = cellND[d0 + ii + jj * imd + kk * imdjmd]; val *= c * (2. - c); }
    }
  }
  else
  {
    constexpr int O = (TYPE == 3) ? 3 : (TYPE == 44) ? 4 : 5;
    for (int kk = 0; kk < O; kk++) for (int jj = 0; jj < O; jj++) for (int ii = 0; ii < O; ii++)
    {
      double cf = cfT[(size_t)ii * m + e] * cfT[(size_t)(jj + O) * m + e] * cfT[(size_t)(kk + 2 * O) * m + e];
      if (fabs(cf) > 1.e-11) { double c = cellND[d0 + ii + jj * imd + kk * imdjmd]; val *= c * (2. - c); }
    }
  }
  cellNR[rcv[e]] = (val > 0.5) ? 1. : 0.;
}

}  // namespace

static int ncf_host(int t) { return t == 1 ? 1 : t == 2 ? 8 : t == 22 ? 4 : t == 3 ? 9 : t == 44 ? 12 : t == 5 ? 15 : -1; }
static int sten_host(int t) { return t == 1 ? 1 : t == 2 ? 8 : t == 22 ? 4 : t == 3 ? 27 : t == 44 ? 64 : t == 5 ? 125 : 0; }

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
__global__ void k_iota(int n, int* __restrict__ a)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) a[e] = e;
}
__global__ void k_plan_flag(int n, int T, const int* __restrict__ type, unsigned long long* __restrict__ f)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) f[e] = (type[e] == T) ? 1ull : 0ull;
}
// tiled layout: sort key = donor box of the entry (entries of other types sort last), value = entry
__global__ void k_tile_keys(int n, int T, const int* __restrict__ type, const int* __restrict__ donor, int imd, int jmd,
                            int ntx, int nty, unsigned nboxes, unsigned* __restrict__ key, int* __restrict__ val)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  unsigned kk = nboxes;
  if (type[e] == T)
  {
    int d = donor[e];
    int i = d % imd, j = (d / imd) % jmd, k = d / (imd * jmd);
    kk = (unsigned)((k * nty + j / CX_TY) * ntx + i / CX_TX);
  }
  key[e] = kk; val[e] = e;
}
// donor-sorted layout: sort key = donor index of the entry (entries of other types sort last), value = entry
__global__ void k_donor_keys(int n, int T, const int* __restrict__ type, const int* __restrict__ donor, unsigned sentinel,
                             unsigned* __restrict__ key, int* __restrict__ val)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  key[e] = (type[e] == T) ? (unsigned)donor[e] : sentinel;
  val[e] = e;
}
// box layout: sort key = box of the entry's donor cell, boxes of (ti,tj,tk) cells laid from the corner (i0,j0,k0)
// of the bounding box of the group's donor cells (entries of other types sort last), value = entry
struct BoxGrid { int i0, j0, k0, i1, j1, k1, ti, tj, tk, nbx, nby, nbz; };
__global__ void k_box_keys(int n, int T, const int* __restrict__ type, const int* __restrict__ donor, int imd, int jmd,
                           BoxGrid B, unsigned nboxes, unsigned* __restrict__ key, int* __restrict__ val)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  unsigned kk = nboxes;
  if (type[e] == T)
  {
    int d = donor[e];
    int i = d % imd, j = (d / imd) % jmd, k = d / (imd * jmd);
    kk = (unsigned)((((k - B.k0) / B.tk) * B.nby + (j - B.j0) / B.tj) * B.nbx + (i - B.i0) / B.ti);
  }
  key[e] = kk; val[e] = e;
}
__global__ void k_box_table(int m, const unsigned* __restrict__ key, const unsigned long long* __restrict__ scan,
                            BoxTile* __restrict__ tiles, int ntiles, BoxGrid B, int imd, int imdjmd)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) { BoxTile t; t.start = m; t.base = 0; t.dims = 0; t.pad_ = 0; tiles[ntiles] = t; }
  if (s >= m) return;
  if (s == 0 || key[s] != key[s - 1])
  {
    const int b = (int)key[s];
    const int bx = b % B.nbx, by = (b / B.nbx) % B.nby, bz = b / (B.nbx * B.nby);
    const int ci = B.i0 + bx * B.ti, cj = B.j0 + by * B.tj, ck = B.k0 + bz * B.tk;     // first donor cell of the box
    const int ilen = min(B.ti, B.i1 - ci + 1) + 1, jlen = min(B.tj, B.j1 - cj + 1) + 1, klen = min(B.tk, B.k1 - ck + 1) + 1;
    BoxTile t; t.start = s; t.base = ci + cj * imd + ck * imdjmd; t.dims = ilen | (jlen << 8) | (klen << 16); t.pad_ = 0;
    tiles[scan[s]] = t;
  }
}
__global__ void k_tile_heads(int m, const unsigned* __restrict__ key, unsigned long long* __restrict__ f)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < m) f[s] = (s == 0 || key[s] != key[s - 1]) ? 1ull : 0ull;
}
__global__ void k_tile_table(int m, const unsigned* __restrict__ key, const unsigned long long* __restrict__ scan,
                             TileRef* __restrict__ tiles, int ntiles)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) { tiles[ntiles].tile = -1; tiles[ntiles].start = m; }
  if (s >= m) return;
  if (s == 0 || key[s] != key[s - 1]) { TileRef t; t.tile = (int)key[s]; t.start = s; tiles[scan[s]] = t; }
}
__global__ void k_plan_place_perm(int m, int ncf, const int* __restrict__ perm, const int* __restrict__ donor,
                                  const int* __restrict__ rcv, const unsigned long long* __restrict__ off,
                                  const double* __restrict__ coefs, int* __restrict__ od, int* __restrict__ orc,
                                  int* __restrict__ op, double* __restrict__ cfT, unsigned short* __restrict__ pat)
{
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= m) return;
  int e = perm[slot];
  od[slot] = donor[e]; orc[slot] = rcv[e]; op[slot] = e;
  const double* c = coefs + off[e];
  if (pat == nullptr)
  {
    for (int k = 0; k < ncf; k++) cfT[(size_t)k * m + slot] = c[k];
    return;
  }
  // packed type-2 entry (the host has checked that its 8 weights take at most 4 distinct bit patterns): distinct
  // values in order of first appearance + 2-bit selector per vertex
  long long val[4] = {0, 0, 0, 0};
  int nd = 0; unsigned p = 0;
#pragma unroll
  for (int k = 0; k < 8; k++)
  {
    const long long b = __double_as_longlong(c[k]);
    int sl = -1;
#pragma unroll
    for (int q = 0; q < 4; q++) if (sl < 0 && q < nd && val[q] == b) sl = q;
    if (sl < 0)
    {
      sl = nd < 4 ? nd : 3;
#pragma unroll
      for (int q = 0; q < 4; q++) if (q == sl) val[q] = b;
      nd++;
    }
    p |= (unsigned)sl << (2 * k);
  }
#pragma unroll
  for (int q = 0; q < 4; q++) cfT[(size_t)q * m + slot] = __longlong_as_double(val[q]);
  pat[slot] = (unsigned short)p;
}
}  // namespace

// CX_TRANSFER_STAGING=0 disables the shared-memory staged type-2 kernel (A/B measurements)
int cx_transfer_staging()
{
  static int v = -1;
  if (v < 0) { const char* s = getenv("CX_TRANSFER_STAGING"); v = (s && s[0] == '0') ? 0 : 1; }
  return v;
}

// CX_PLAN_SORT: 0 = keep the entries in list order, 1 = donor-sorted layout for every group,
// 2 = tiled layout for the type-2 groups that are dense enough (experimental), 4 = donor-sorted with the
// counting sort / atomic cursor of the first version (A/B), unset = 1
static int plan_sort_mode()
{
  const char* s = getenv("CX_PLAN_SORT");
  if (!s || !s[0]) return -1;
  return s[0] - '0';
}

// Try the tiled layout for the type-2 group G (m entries).  On success fills G.d_tiles/ntiles and the
// regrouped arrays and returns 1; returns 0 if the group is too sparse (caller falls back), <0 on error.
static int build_tiled_group(cx_ctx* ctx, cx_plan* P, cx_plan::Group& G, int n, int kmd, const int* r_ty, const int* r_dn,
                             const int* r_rc, const unsigned long long* r_off, const double* r_cf, int first, int force)
{
  cudaStream_t s = ctx->stream;
  const int TB = 256;
  const int ntx = (std::max(P->imd - 1, 1) + CX_TX - 1) / CX_TX, nty = (std::max(P->jmd - 1, 1) + CX_TY - 1) / CX_TY;
  const long long nboxes_ll = (long long)ntx * nty * std::max(kmd - 1, 1);
  if (nboxes_ll >= 0x7fffffffll) return 0;
  const unsigned nboxes = (unsigned)nboxes_ll;
  unsigned *k_in, *k_out; int *v_in, *v_out; void* tmp = nullptr; size_t tmp_bytes = 0;
  CX_MALLOC(k_in, sizeof(unsigned) * (size_t)n, s); CX_MALLOC(k_out, sizeof(unsigned) * (size_t)n, s);
  CX_MALLOC(v_in, sizeof(int) * (size_t)n, s); CX_MALLOC(v_out, sizeof(int) * (size_t)n, s);
  k_tile_keys<<<cx_grid(n, TB), TB, 0, s>>>(n, G.sel, r_ty, r_dn, P->imd, P->jmd, ntx, nty, nboxes, k_in, v_in);
  int bits = 1; while ((1ull << bits) <= (unsigned long long)nboxes) bits++;
  // stable LSD radix sort (CUB): entries keep their list order (ascending receptors) inside a box
  CX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, n, 0, bits, s));
  CX_MALLOC(tmp, tmp_bytes, s);
  CX_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, n, 0, bits, s));
  const int m = G.n;
  unsigned long long *f, *sc, total = 0;
  CX_MALLOC(f, sizeof(unsigned long long) * (size_t)m, s); CX_MALLOC(sc, sizeof(unsigned long long) * (size_t)m, s);
  k_tile_heads<<<cx_grid(m, TB), TB, 0, s>>>(m, k_out, f);
  CX_CHECK(cx_exclusive_scan_u64(ctx, f, sc, m, &total));
  ctx->launches += 3;
  int ok = 0;
  if (force || (long long)m >= 96ll * (long long)total)
  {
    G.ntiles = (int)total; G.ntx = ntx; G.nty = nty; G.layout = 2;
    CX_CUDA(cudaMalloc(&G.d_tiles, sizeof(TileRef) * ((size_t)total + 1)));
    k_tile_table<<<cx_grid(m, TB), TB, 0, s>>>(m, k_out, sc, (TileRef*)G.d_tiles, (int)total);
    k_plan_place_perm<<<cx_grid(m, TB), TB, 0, s>>>(m, G.ncf, v_out, r_dn, r_rc, r_off, r_cf, P->d_donor + first,
                                                   P->d_rcv + first, P->d_pos + first, G.d_cf, G.d_pat);
    ctx->launches += 2;
    ok = 1;
  }
  cudaFreeAsync(k_in, s); cudaFreeAsync(k_out, s); cudaFreeAsync(v_in, s); cudaFreeAsync(v_out, s);
  cudaFreeAsync(tmp, s); cudaFreeAsync(f, s); cudaFreeAsync(sc, s);
  return ok;
}

// Box layout for the type-2 group G (see transfer_box3).  bb = i0,i1,j0,j1,k0,k1 of the group's donor cells (lowest
// corner node of each stencil).  Returns 1 and fills the group when the boxes are filled well enough (or force),
// 0 if the group is too sparse (caller falls back to the donor-sorted layout), < 0 on error.
static int build_box_group(cx_ctx* ctx, cx_plan* P, cx_plan::Group& G, int n, const int* bb, const int* r_ty,
                           const int* r_dn, const int* r_rc, const unsigned long long* r_off, const double* r_cf, int first,
                           int force)
{
  cudaStream_t s = ctx->stream;
  const int TB = 256;
  const int ei = bb[1] - bb[0] + 1, ej = bb[3] - bb[2] + 1, ek = bb[5] - bb[4] + 1;     // extents in cells
  BoxGrid B;
  B.i0 = bb[0]; B.i1 = bb[1]; B.j0 = bb[2]; B.j1 = bb[3]; B.k0 = bb[4]; B.k1 = bb[5];
  // rows as long as possible (up to 32 cells), the thinner of the two other directions whole or ~sqrt of the rest
  B.ti = std::min(ei, 32);
  int rem = std::max(1, CX_BOX_CELLS / B.ti);
  int rt = 1; while ((rt + 1) * (rt + 1) <= rem) rt++;
  if (ek <= ej) { B.tk = std::min(ek, std::max(rt, 1)); B.tj = std::min(ej, std::max(1, rem / B.tk)); }
  else { B.tj = std::min(ej, std::max(rt, 1)); B.tk = std::min(ek, std::max(1, rem / B.tj)); }
  auto vs_of = [&]() { return ((B.ti + 1) | 1) * (B.tj + 1) * (B.tk + 1); };
  while (vs_of() > CX_BOX_NODES) { if (B.tj >= B.tk && B.tj > 1) B.tj--; else if (B.tk > 1) B.tk--; else B.ti--; }
  if (B.tj > 254 || B.tk > 254) return 0;
  B.nbx = (ei + B.ti - 1) / B.ti; B.nby = (ej + B.tj - 1) / B.tj; B.nbz = (ek + B.tk - 1) / B.tk;
  const long long nboxes_ll = (long long)B.nbx * B.nby * B.nbz;
  if (nboxes_ll >= 0x7fffffffll) return 0;
  const unsigned nboxes = (unsigned)nboxes_ll;
  unsigned *k_in, *k_out; int *v_in, *v_out; void* tmp = nullptr; size_t tmp_bytes = 0;
  CX_MALLOC(k_in, sizeof(unsigned) * (size_t)n, s); CX_MALLOC(k_out, sizeof(unsigned) * (size_t)n, s);
  CX_MALLOC(v_in, sizeof(int) * (size_t)n, s); CX_MALLOC(v_out, sizeof(int) * (size_t)n, s);
  k_box_keys<<<cx_grid(n, TB), TB, 0, s>>>(n, G.sel, r_ty, r_dn, P->imd, P->jmd, B, nboxes, k_in, v_in);
  int bits = 1; while ((1ull << bits) <= (unsigned long long)nboxes) bits++;
  // stable LSD radix sort (CUB): entries keep their list order (ascending receptors) inside a box
  CX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, n, 0, bits, s));
  CX_MALLOC(tmp, tmp_bytes, s);
  CX_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, n, 0, bits, s));
  const int m = G.n;
  unsigned long long *f, *sc, total = 0;
  CX_MALLOC(f, sizeof(unsigned long long) * (size_t)m, s); CX_MALLOC(sc, sizeof(unsigned long long) * (size_t)m, s);
  k_tile_heads<<<cx_grid(m, TB), TB, 0, s>>>(m, k_out, f);
  CX_CHECK(cx_exclusive_scan_u64(ctx, f, sc, m, &total));
  ctx->launches += 3;
  int ok = 0;
  // worth it when a box's entries reuse its nodes: at least ~1 entry per 4 staged nodes on average
  if (force || (long long)m * 4 >= (long long)total * vs_of())
  {
    G.ntiles = (int)total; G.layout = 3; G.box_vs = vs_of();
    CX_CUDA(cudaMalloc(&G.d_tiles, sizeof(BoxTile) * ((size_t)total + 1)));
    k_box_table<<<cx_grid(m, TB), TB, 0, s>>>(m, k_out, sc, (BoxTile*)G.d_tiles, (int)total, B, P->imd, P->imd * P->jmd);
    k_plan_place_perm<<<cx_grid(m, TB), TB, 0, s>>>(m, G.ncf, v_out, r_dn, r_rc, r_off, r_cf, P->d_donor + first,
                                                   P->d_rcv + first, P->d_pos + first, G.d_cf, G.d_pat);
    ctx->launches += 2;
    ok = 1;
  }
  cudaFreeAsync(k_in, s); cudaFreeAsync(k_out, s); cudaFreeAsync(v_in, s); cudaFreeAsync(v_out, s);
  cudaFreeAsync(tmp, s); cudaFreeAsync(f, s); cudaFreeAsync(sc, s);
  return ok;
}

// Donor-sorted layout with a STABLE sort (CUB LSD radix sort on the donor index): entries that share a donor cell
// keep their list order, i.e. ascending receptor index, so the lanes that handle them store to neighbouring
// receptor addresses and the warp's store touches fewer 32-byte sectors than with an atomic-cursor placement.
static int build_sorted_group(cx_ctx* ctx, cx_plan* P, cx_plan::Group& G, int n, const int* r_ty, const int* r_dn,
                              const int* r_rc, const unsigned long long* r_off, const double* r_cf, int first)
{
  cudaStream_t s = ctx->stream;
  const int TB = 256;
  unsigned *k_in, *k_out; int *v_in, *v_out; void* tmp = nullptr; size_t tmp_bytes = 0;
  CX_MALLOC(k_in, sizeof(unsigned) * (size_t)n, s); CX_MALLOC(k_out, sizeof(unsigned) * (size_t)n, s);
  CX_MALLOC(v_in, sizeof(int) * (size_t)n, s); CX_MALLOC(v_out, sizeof(int) * (size_t)n, s);
  const unsigned sentinel = (unsigned)P->nptsD;
  k_donor_keys<<<cx_grid(n, TB), TB, 0, s>>>(n, G.sel, r_ty, r_dn, sentinel, k_in, v_in);
  int bits = 1; while ((1ull << bits) <= (unsigned long long)sentinel) bits++;
  CX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, n, 0, bits, s));
  CX_MALLOC(tmp, tmp_bytes, s);
  CX_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, n, 0, bits, s));
  k_plan_place_perm<<<cx_grid(G.n, TB), TB, 0, s>>>(G.n, G.ncf, v_out, r_dn, r_rc, r_off, r_cf, P->d_donor + first,
                                                     P->d_rcv + first, P->d_pos + first, G.d_cf, G.d_pat);
  ctx->launches += 3;
  cudaFreeAsync(k_in, s); cudaFreeAsync(k_out, s); cudaFreeAsync(v_in, s); cudaFreeAsync(v_out, s); cudaFreeAsync(tmp, s);
  if (G.type == 2)
  {
    // per-tile donor ranges for the pipelined kernel (kept in d_tiles: layout 1 has no box table)
    G.ntiles = cx_grid(G.n, 256);
    CX_CUDA(cudaMalloc(&G.d_tiles, sizeof(PipeRange) * (size_t)G.ntiles));
    k_pipe_ranges<<<cx_grid(G.ntiles, TB), TB, 0, s>>>(G.ntiles, G.n, P->imd, P->imd * P->jmd, P->nptsD, P->d_donor + first,
                                                       (PipeRange*)G.d_tiles);
    ctx->launches++;
  }
  return 0;
}

int cx_plan_build(cx_ctx* ctx, int imd, int jmd, long long nptsD, long long nptsR, int n, const int* donor,
                  const int* rcv, const int* type, const double* coefs, long long ncoefs, cx_plan* P)
{
  P->ctx = ctx; P->imd = imd; P->jmd = jmd; P->nptsD = (int)nptsD; P->nptsR = (int)nptsR; P->n = n;
  P->d_donor = P->d_rcv = P->d_pos = P->d_seq = nullptr; P->alg_bytes_fixed = 0; P->distinct_nodes = -1; P->last_ms = 0.f;
  P->groups.clear();
  if (n == 0) { P->ngroups = 0; return 0; }
  // group order: type 1, packed type 2, raw type 2, 3, 5, 44, 22
  const int types[7] = {1, 2, 2, 3, 5, 44, 22};
  const int sels[7] = {1, CX_SEL_PACKED2, 2, 3, 5, 44, 22};
  const int mode_env = plan_sort_mode();
  // packed type-2 groups (load_cf2) for the donor-sorted and box layouts.  Default: the plans a plan set fuses
  // (< CX_BIG_PLAN entries; C4 step 57.6 -> 56.3 us) -- on the 17 M-entry C2a plan the 22 % fewer bytes do not pay
  // (0.499 vs 0.490 ms: that launch is bound by latency and LSU wavefronts, not by DRAM).  CX_PLAN_PACK=0 / 1 forces.
  const char* pk_env = getenv("CX_PLAN_PACK");
  const bool pack_on = pk_env && pk_env[0] ? pk_env[0] != '0' : n < CX_BIG_PLAN;
  const bool pack2 = pack_on && (mode_env < 0 || mode_env == 1 || mode_env == 5);
  std::vector<unsigned long long> off((size_t)n);
  std::vector<int> sel((size_t)n);
  long long run = 0;
  int cnt[7] = {0, 0, 0, 0, 0, 0, 0};
  int bb2[6] = {INT_MAX, -1, INT_MAX, -1, INT_MAX, -1};      // donor cells of the type-2 entries: i0,i1,j0,j1,k0,k1
  for (int e = 0; e < n; e++)
  {
    off[e] = (unsigned long long)run;
    int nc = ncf_host(type[e]);
    if (nc < 0) { cx_set_error("transfer plan: interpolation type %d is not supported (structured donors: 1,2,22,3,44,5)", type[e]); return -1; }
    if (donor[e] < 0 || (long long)donor[e] >= nptsD) { cx_set_error("transfer plan: donor index %d out of range", donor[e]); return -1; }
    if (rcv[e] < 0 || (long long)rcv[e] >= nptsR) { cx_set_error("transfer plan: receptor index %d out of range", rcv[e]); return -1; }
    if (run + nc > ncoefs) { cx_set_error("transfer plan: coefficient array has %lld entries, types need more", ncoefs); return -1; }
    int sl = type[e];
    if (type[e] == 2)
    {
      const int d = donor[e], ci = d % imd, cj = (d / imd) % jmd, ck = d / (imd * jmd);
      bb2[0] = std::min(bb2[0], ci); bb2[1] = std::max(bb2[1], ci); bb2[2] = std::min(bb2[2], cj);
      bb2[3] = std::max(bb2[3], cj); bb2[4] = std::min(bb2[4], ck); bb2[5] = std::max(bb2[5], ck);
      if (pack2)
      {
        // packable: the 8 weights take at most 4 distinct bit patterns (always true for the reference's
        // tetrahedron weights, InterpData.cpp:621-631; extrapolated entries may not be)
        long long b[8], val[4]; int nd = 0;
        memcpy(b, coefs + run, sizeof(b));
        for (int k = 0; k < 8 && nd <= 4; k++)
        {
          int q = 0;
          while (q < nd && q < 4 && val[q] != b[k]) q++;
          if (q == nd || q == 4) { if (nd < 4) val[nd] = b[k]; nd++; }
        }
        if (nd <= 4) sl = CX_SEL_PACKED2;
      }
    }
    sel[e] = sl;
    run += nc;
    for (int g = 0; g < 7; g++) if (sl == sels[g]) cnt[g]++;
  }
  if (run != ncoefs) { cx_set_error("transfer plan: coefficient array has %lld entries, types need %lld", ncoefs, run); return -1; }
  P->alg_bytes_fixed = 8 * run + 12ll * n;
  cudaStream_t s = ctx->stream;
  const size_t N = n;
  int *r_dn, *r_rc, *r_ty; unsigned long long *r_off, *d_cnt, *d_scan; double* r_cf;
  CX_MALLOC(r_dn, sizeof(int) * N, s); CX_MALLOC(r_rc, sizeof(int) * N, s);
  CX_MALLOC(r_ty, sizeof(int) * N, s); CX_MALLOC(r_off, sizeof(unsigned long long) * N, s);
  CX_MALLOC(r_cf, sizeof(double) * (size_t)std::max(run, 1ll), s);
  const int sorted = mode_env == 0 ? 0 : 1;
  const int kmd = (int)(nptsD / ((long long)imd * jmd));
  const size_t nscan = sorted ? (size_t)nptsD : N;
  CX_MALLOC(d_cnt, sizeof(unsigned long long) * nscan, s);
  CX_MALLOC(d_scan, sizeof(unsigned long long) * nscan, s);
  CX_CHECK(cx_upload(ctx, r_dn, donor, sizeof(int) * N));
  CX_CHECK(cx_upload(ctx, r_rc, rcv, sizeof(int) * N));
  CX_CHECK(cx_upload(ctx, r_ty, sel.data(), sizeof(int) * N));
  CX_CHECK(cx_upload(ctx, r_off, off.data(), sizeof(unsigned long long) * N));
  CX_CHECK(cx_upload(ctx, r_cf, coefs, sizeof(double) * (size_t)run));
  // + 64 bytes: the pipelined kernel's bulk copies round a tile's source range outwards to 16 bytes
  CX_MALLOC(P->d_donor, sizeof(int) * N + 64, s);
  CX_MALLOC(P->d_rcv, sizeof(int) * N + 64, s);
  CX_MALLOC(P->d_pos, sizeof(int) * N + 64, s);
  CX_MALLOC(P->d_seq, sizeof(int) * N + 64, s);
  k_iota<<<cx_grid(n, 256), 256, 0, s>>>(n, P->d_seq);
  const int TB = 256;
  int first = 0;
  for (int g = 0; g < 7; g++)
  {
    if (cnt[g] == 0) continue;
    cx_plan::Group G; G.type = types[g]; G.sel = sels[g]; G.ncf = ncf_host(types[g]); G.sten = sten_host(types[g]);
    G.n = cnt[g]; G.first = first; G.d_cf = nullptr; G.d_pat = nullptr;
    G.layout = sorted ? 1 : 0; G.d_tiles = nullptr; G.ntiles = 0; G.ntx = G.nty = 0; G.box_vs = 0;
    if (G.sel == CX_SEL_PACKED2)
    {
      CX_MALLOC(G.d_cf, sizeof(double) * 4 * (size_t)G.n + 64, s);
      CX_MALLOC(G.d_pat, sizeof(unsigned short) * (size_t)G.n + 64, s);
    }
    else
    CX_MALLOC(G.d_cf, sizeof(double) * (size_t)G.ncf * G.n + 64, s);
    if (G.type == 2 && (mode_env == 2 || mode_env == 3) && kmd >= 2 && G.n >= 4096)
    {
      int rc = build_tiled_group(ctx, P, G, n, kmd, r_ty, r_dn, r_rc, r_off, r_cf, first, mode_env == 3);
      if (rc < 0) return rc;
      if (rc == 1) { P->groups.push_back(G); first += G.n; continue; }
    }
    // box layout: default for the plans a plan set fuses (below CX_BIG_PLAN entries) when their boxes fill well;
    // CX_PLAN_SORT=5 forces it for every 3-D type-2 group, =1 keeps the donor-sorted layout
    if (G.type == 2 && kmd >= 2 && (mode_env == 5 || (mode_env < 0 && G.n >= 1024 && n < CX_BIG_PLAN)))
    {
      int rc = build_box_group(ctx, P, G, n, bb2, r_ty, r_dn, r_rc, r_off, r_cf, first, mode_env == 5);
      if (rc < 0) return rc;
      if (rc == 1) { P->groups.push_back(G); first += G.n; continue; }
    }
    if (sorted && mode_env != 4)
    {
      CX_CHECK(build_sorted_group(ctx, P, G, n, r_ty, r_dn, r_rc, r_off, r_cf, first));
      P->groups.push_back(G);
      first += G.n;
      continue;
    }
    if (sorted)       // CX_PLAN_SORT=4: counting sort with an atomic cursor (order inside a donor cell is arbitrary)
    {
      CX_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * nscan, s));
      k_plan_hist<<<cx_grid(n, TB), TB, 0, s>>>(n, G.sel, r_ty, r_dn, d_cnt);
      CX_CHECK(cx_exclusive_scan_u64(ctx, d_cnt, d_scan, (int)nscan, nullptr));
    }
    else
    {
      k_plan_flag<<<cx_grid(n, TB), TB, 0, s>>>(n, G.sel, r_ty, d_cnt);
      CX_CHECK(cx_exclusive_scan_u64(ctx, d_cnt, d_scan, n, nullptr));
    }
    k_plan_place<<<cx_grid(n, TB), TB, 0, s>>>(n, G.sel, G.ncf, G.n, sorted, r_ty, r_dn, r_rc, r_off, r_cf, d_scan,
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

// CX_TRANSFER_PIPE=1 selects the persistent pipelined type-2 kernel (experimental: measured slower than the
// per-tile kernel on C2a, 0.62 vs 0.49 ms for 5 variables; read at every launch so tests can switch it)
static int cx_transfer_pipelined()
{
  const char* e = getenv("CX_TRANSFER_PIPE");
  return (e && e[0] == '1') ? 1 : 0;
}

static int pipe_smem_optin()
{
  static int done = 0;
  if (!done)
  {
    const int mx = 224 * 1024;
    CX_CUDA(cudaFuncSetAttribute(k_transfer2p<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CX_CUDA(cudaFuncSetAttribute(k_transfer2p<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    done = 1;
  }
  return 0;
}

static int tiled_smem_optin()
{
  static int done = 0;
  if (!done)
  {
    const int mx = (int)(sizeof(double) * (CX_BOXV > CX_BOX_NODES ? CX_BOXV : CX_BOX_NODES) * CX_MAXVARS);
    CX_CUDA(cudaFuncSetAttribute(k_transfer2b<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CX_CUDA(cudaFuncSetAttribute(k_transfer2b<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CX_CUDA(cudaFuncSetAttribute(k_transfer_set<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CX_CUDA(cudaFuncSetAttribute(k_transfer2t<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CX_CUDA(cudaFuncSetAttribute(k_transfer2t<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CX_CUDA(cudaFuncSetAttribute(k_transfer_set<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    done = 1;
  }
  return 0;
}


int cx_plan_launch(cx_plan* P, int nvars, const double* const* dD, double* const* dR, double* dense, long long ld,
                   int mode_in, cudaStream_t s)
{
  // mode 0: in place; 1: dense, original entry order (pos); 2: dense, the plan's own sorted order (coalesced stores)
  const int mode = mode_in == 0 ? 0 : 1;
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
      static int tb2s = 0;       // CX_T2S_TB: threads per CTA of the staged type-2 kernel (experiment; default 256)
      if (!tb2s) { const char* e = getenv("CX_T2S_TB"); tb2s = e ? atoi(e) : 256; if (tb2s != 128 && tb2s != 256) tb2s = 256; }
      const int* dn_ = P->d_donor + G.first; const int* rc_ = P->d_rcv + G.first;
      const int* ps_ = (mode_in == 2 ? P->d_seq : P->d_pos) + G.first;
      int imd = P->imd, ij = P->imd * P->jmd;
#define CX_LAUNCH(T)                                                                                           \
      if (mode == 0) k_transfer<T, 0><<<grid, TB, 0, s>>>(G.n, nv, F, imd, ij, dn_, rc_, ps_, G.d_cf, dn, ld, G.d_pat); \
      else k_transfer<T, 1><<<grid, TB, 0, s>>>(G.n, nv, F, imd, ij, dn_, rc_, ps_, G.d_cf, dn, ld, G.d_pat);
      bool staged = G.layout != 0 && cx_transfer_staging();
      for (int v = 0; v < nv && staged; v++) if (((uintptr_t)F.d[v]) & 15) staged = false;
      switch (G.type)
      {
        case 1: CX_LAUNCH(1) break;
        case 2:
          if (staged && G.layout == 2)
          {
            const size_t smb = sizeof(double) * CX_BOXV * (size_t)nv;
            CX_CHECK(tiled_smem_optin());
            const TileRef* tl = (const TileRef*)G.d_tiles;
            if (mode == 0) k_transfer2t<0><<<G.ntiles, TB, smb, s>>>(tl, G.n, nv, F, imd, P->jmd, P->nptsD, G.ntx, G.nty, dn_, rc_, ps_, G.d_cf, dn, ld);
            else k_transfer2t<1><<<G.ntiles, TB, smb, s>>>(tl, G.n, nv, F, imd, P->jmd, P->nptsD, G.ntx, G.nty, dn_, rc_, ps_, G.d_cf, dn, ld);
          }
          else if (G.layout == 3)
          {
            const size_t smb = sizeof(double) * (size_t)G.box_vs * (size_t)nv;
            CX_CHECK(tiled_smem_optin());
            const BoxTile* tl = (const BoxTile*)G.d_tiles;
            if (mode == 0) k_transfer2b<0><<<G.ntiles, TB, smb, s>>>(tl, G.n, nv, F, imd, ij, dn_, rc_, ps_, G.d_cf, dn, ld, G.d_pat);
            else k_transfer2b<1><<<G.ntiles, TB, smb, s>>>(tl, G.n, nv, F, imd, ij, dn_, rc_, ps_, G.d_cf, dn, ld, G.d_pat);
          }
          else if (staged && G.layout == 1)
          {
            const int g2 = cx_grid(G.n, tb2s);
            int nst = 0;
            const char* mt = getenv("CX_PIPE_MIN_TILES");          // tests lower the threshold
            const int min_tiles = mt ? atoi(mt) : 4 * ctx->sm_count;
            if (cx_transfer_pipelined() && !G.d_pat && G.d_tiles && nv <= CX_PIPE_MAXV && G.ntiles >= min_tiles)
            {
              nst = (128 + 3 * pipe_stage_bytes(nv) <= 200 * 1024) ? 3 : (128 + 2 * pipe_stage_bytes(nv) <= 220 * 1024) ? 2 : 0;
            }
            if (nst)
            {
              const size_t smb = 128 + (size_t)nst * pipe_stage_bytes(nv);
              CX_CHECK(pipe_smem_optin());
              const int gp = std::min(G.ntiles, ctx->sm_count);
              const PipeRange* rg = (const PipeRange*)G.d_tiles;
              static int png = 0;    // CX_PIPE_GROUPS: consumer groups of 256 threads per CTA (1..3; + one producer warp)
              if (!png) { const char* e = getenv("CX_PIPE_GROUPS"); png = e ? atoi(e) : 2; if (png < 1 || png > 3) png = 2; }
              const int tbp = 256 * std::min(png, nv) + 32;
              if (mode == 0) k_transfer2p<0><<<gp, tbp, smb, s>>>(G.n, nv, F, imd, ij, dn_, rc_, G.d_cf, rg, G.ntiles, nst, dn, ld);
              else k_transfer2p<1><<<gp, tbp, smb, s>>>(G.n, nv, F, imd, ij, dn_, ps_, G.d_cf, rg, G.ntiles, nst, dn, ld);
            }
            else if (mode == 0) k_transfer2s<0><<<g2, tb2s, sizeof(StageSmem), s>>>(G.n, nv, F, imd, ij, P->nptsD, dn_, rc_, ps_, G.d_cf, dn, ld, G.d_pat);
            else k_transfer2s<1><<<g2, tb2s, sizeof(StageSmem), s>>>(G.n, nv, F, imd, ij, P->nptsD, dn_, rc_, ps_, G.d_cf, dn, ld, G.d_pat);
          }
          else { CX_LAUNCH(2) }
          break;
        case 3:
          if (staged && G.layout == 1)
          {
            if (mode == 0) k_transferOs<3, 0><<<grid, TB, sizeof(StageSmemO<3>), s>>>(G.n, nv, F, imd, ij, P->nptsD, dn_, rc_, ps_, G.d_cf, dn, ld);
            else k_transferOs<3, 1><<<grid, TB, sizeof(StageSmemO<3>), s>>>(G.n, nv, F, imd, ij, P->nptsD, dn_, rc_, ps_, G.d_cf, dn, ld);
          }
          else { CX_LAUNCH(3) }
          break;
        case 44: CX_LAUNCH(44) break;
        case 22: CX_LAUNCH(22) break;
        case 5: CX_LAUNCH(5) break;   // 125-point stencils: staging measured slower (5.5 vs 3.5 ms on C2a), direct gather
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
      case 2: k_transfer_celln<2><<<grid, TB, 0, s>>>(G.n, dCellND, dCellNR, imd, ij, dn_, rc_, G.d_cf, G.d_pat); break;
      case 3: k_transfer_celln<3><<<grid, TB, 0, s>>>(G.n, dCellND, dCellNR, imd, ij, dn_, rc_, G.d_cf); break;
      case 44: k_transfer_celln<44><<<grid, TB, 0, s>>>(G.n, dCellND, dCellNR, imd, ij, dn_, rc_, G.d_cf); break;
      case 22: k_transfer_celln<22><<<grid, TB, 0, s>>>(G.n, dCellND, dCellNR, imd, ij, dn_, rc_, G.d_cf); break;
      case 5: k_transfer_celln<5><<<grid, TB, 0, s>>>(G.n, dCellND, dCellNR, imd, ij, dn_, rc_, G.d_cf); break;
    }
    P->ctx->launches++;
  }
  return 0;
}

// Periodicity by rotation (Connector/Connector/includeTransfers.h:1-60): after a transfer the vector (fx,fy,fz) of
// every receptor of the plan is rotated by theta about axis dirR (1: x, 2: y, 3: z); same expressions, no FMA.
// dense != 0: the three arrays are dense (n) blocks of one transfer (any common entry order); else they are
// receptor fields addressed through the plan's receptor indices.
namespace {
__global__ void k_rotate(int n, const int* __restrict__ rcv, int dirR, double ct, double st, double* __restrict__ fx,
                         double* __restrict__ fy, double* __restrict__ fz)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const size_t r = rcv ? (size_t)rcv[e] : (size_t)e;
  if (dirR == 1)
  {
    double v1 = fy[r], v2 = fz[r];
    fy[r] = ct * v1 - st * v2;
    fz[r] = st * v1 + ct * v2;
  }
  else if (dirR == 2)
  {
    double v1 = fx[r], v2 = fz[r];
    fx[r] = ct * v1 + st * v2;
    fz[r] = -st * v1 + ct * v2;
  }
  else if (dirR == 3)
  {
    double v1 = fx[r], v2 = fy[r];
    fx[r] = ct * v1 - st * v2;
    fy[r] = st * v1 + ct * v2;
  }
}
}  // namespace

int cx_plan_rotate_on(cx_plan* P, int dirR, double ct, double st, double* fx, double* fy, double* fz, int dense,
                      cudaStream_t s)
{
  if (P->n == 0 || dirR == 0) return 0;
  k_rotate<<<cx_grid(P->n, 256), 256, 0, s>>>(P->n, dense ? nullptr : P->d_rcv, dirR, ct, st, fx, fy, fz);
  P->ctx->launches++;
  CX_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// plan set / scatter set
// ---------------------------------------------------------------------------
struct cx_planset {
  cx_ctx* ctx;
  int nvars, ndesc, staged, tiled, boxed;
  // plans too large to gain anything from fusion run as their own launches (leaner kernel, more CTAs per SM)
  struct Big { cx_plan* P; int mode; std::vector<const double*> dD; std::vector<double*> dR; double* dense; long long ld; };
  std::vector<Big> big;
  size_t smem, smem3, smem5;            // dynamic shared memory of the type-2 / 3 / 5 launches
  int nblocks[6];                       // per interpolation type 1,2,3,5,44,22
  SetDesc* d_descs[6]; BlkRec* d_blk2desc[6];
  void** d_tables; size_t ntables;      // (unused: pointer tables are inline in the descriptors)
};

struct cx_scatterset {
  cx_ctx* ctx;
  int nvars, nblocks, ndesc;
  ScatDesc* d_descs; int* d_blk2desc;
  std::vector<void*> owned;
};

// Descriptor / pointer tables: host -> device ON the context's stream and complete on return.  (A plain cudaMemcpy from
// pageable memory may return before its DMA has landed, and the non-blocking streams the kernels run on are not
// ordered behind the legacy stream it uses.)
static int upload_table(cx_ctx* ctx, void* dptr, const void* host, size_t bytes)
{
  if (bytes == 0) return 0;
  CX_CUDA(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static int upload_ptr_table(cx_ctx* ctx, const double* const* host, int nvars, double*** out)
{
  double** d;
  CX_CUDA(cudaMalloc(&d, sizeof(double*) * (size_t)nvars));
  CX_CHECK(upload_table(ctx, d, host, sizeof(double*) * (size_t)nvars));
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
  S->staged = cx_transfer_staging(); S->smem = 0; S->smem3 = 0; S->smem5 = 0; S->tiled = 0; S->boxed = 0;
  if (nvars > CX_MAXVARS) { cx_set_error("plan set: at most %d variables per set", CX_MAXVARS); delete S; return -1; }
  std::vector<SetDesc> descs[6]; std::vector<BlkRec> b2d[6]; std::vector<void*> tables;
  const int TB = 256;
  int nb[6] = {0, 0, 0, 0, 0, 0};
  bool has2 = false;
  for (int p = 0; p < nplans; p++)
    if (plans[p]->n > 0 && plans[p]->n < CX_BIG_PLAN)
      for (auto& G : plans[p]->groups) if (G.type == 2) has2 = true;
  for (int p = 0; p < nplans; p++)
  {
    cx_plan* P = plans[p];
    if (P->n == 0) continue;
    if (P->n >= CX_BIG_PLAN)
    {
      cx_planset::Big B;
      B.P = P; B.mode = modes[p]; B.dense = modes[p] != 0 ? d_dense[p] : nullptr; B.ld = modes[p] != 0 ? ld[p] : 0;
      B.dD.assign(d_fieldD[p], d_fieldD[p] + nvars);
      if (modes[p] == 0) B.dR.assign(d_fieldR[p], d_fieldR[p] + nvars);
      S->big.push_back(B);
      continue;
    }
    for (auto& G : P->groups)
    {
      int t = G.type == 1 ? (has2 ? 1 : 0) : G.type == 2 ? 1 : G.type == 3 ? 2 : G.type == 5 ? 3 : G.type == 44 ? 4 : 5;
      SetDesc D;
      D.m = G.n; D.type = G.type; D.imd = P->imd; D.imdjmd = P->imd * P->jmd; D.mode = modes[p] == 0 ? 0 : 1; D.first_block = nb[t];
      D.nptsD = P->nptsD; D.layout = G.layout; D.jmd = P->jmd; D.ntx = G.ntx; D.nty = G.nty;
      D.tiles = (const TileRef*)G.d_tiles;
      if (G.type == 2 && G.layout == 2) { S->smem = std::max(S->smem, sizeof(double) * CX_BOXV * (size_t)nvars); S->tiled = 1; }
      if (G.type == 2 && G.layout == 1) S->smem = std::max(S->smem, sizeof(StageSmem));
      if (G.type == 2 && G.layout == 3) { S->smem = std::max(S->smem, sizeof(double) * (size_t)G.box_vs * (size_t)nvars); S->boxed = 1; }
      if (G.type == 3 && G.layout == 1) S->smem3 = sizeof(StageSmemO<3>);
      for (int v = 0; v < nvars; v++) if (((uintptr_t)d_fieldD[p][v]) & 15) S->staged = 0;
      D.donor = P->d_donor + G.first;
      D.idx = (modes[p] == 0 ? P->d_rcv : modes[p] == 1 ? P->d_pos : P->d_seq) + G.first; D.cfT = G.d_cf;
      D.pat = G.d_pat; D.pad_ = 0; D.dense = modes[p] != 0 ? d_dense[p] : nullptr; D.ld = modes[p] != 0 ? ld[p] : 0;
      for (int v = 0; v < CX_MAXVARS; v++)
      {
        D.dp[v] = v < nvars ? d_fieldD[p][v] : nullptr;
        D.rp[v] = (v < nvars && modes[p] == 0) ? d_fieldR[p][v] : nullptr;
      }
      int blocks = (G.type == 2 && (G.layout == 2 || G.layout == 3)) ? G.ntiles : cx_grid(G.n, TB);
      for (int b = 0; b < blocks; b++) b2d[t].push_back(BlkRec{(int)descs[t].size(), -1, -1, 0, 0, 0, 0, 0});
      descs[t].push_back(D);
      nb[t] += blocks;
      S->ndesc++;
    }
  }
  // Measured and rejected (round 2, third session): loading the plan data of a set with an L2 evict_last policy
  // (createpolicy.fractional.L2::evict_last + ld.global.nc.L2::cache_hint), so that the 38 MB of plans of the C4 step
  // stay resident between the iterations of a loop while the fields stream through: 59.6 us per step either way
  // (gpurun_out/r3_keep_*.json) -- the step waits for the sector-granular field traffic, not for the plans.
  for (int t = 0; t < 6; t++)
  {
    S->nblocks[t] = nb[t]; S->d_descs[t] = nullptr; S->d_blk2desc[t] = nullptr;
    if (nb[t] == 0) continue;
    CX_CUDA(cudaMalloc(&S->d_descs[t], sizeof(SetDesc) * descs[t].size()));
    CX_CUDA(cudaMalloc(&S->d_blk2desc[t], sizeof(BlkRec) * b2d[t].size()));
    CX_CHECK(upload_table(ctx, S->d_descs[t], descs[t].data(), sizeof(SetDesc) * descs[t].size()));
    CX_CHECK(upload_table(ctx, S->d_blk2desc[t], b2d[t].data(), sizeof(BlkRec) * b2d[t].size()));
    if (t == 1)
    {
      k_block_ranges<<<cx_grid(nb[t], 256), 256, 0, ctx->stream>>>(nb[t], S->d_descs[t], S->d_blk2desc[t], TB);
      CX_CUDA(cudaStreamSynchronize(ctx->stream));
      CX_CUDA(cudaGetLastError());
    }
  }
  S->ntables = tables.size();
  S->d_tables = (void**)malloc(sizeof(void*) * (tables.size() + 1));
  for (size_t i = 0; i < tables.size(); i++) S->d_tables[i] = tables[i];
  *out = S;
  return 0;
}

static int planset_run_on(cx_planset* S, cudaStream_t s, int v0 = 0, int nv = -1)
{
  if (nv < 0) nv = S->nvars - v0;
  for (auto& B : S->big)
    CX_CHECK(cx_plan_launch(B.P, nv, B.dD.data() + v0, B.mode == 0 ? B.dR.data() + v0 : nullptr,
                            B.dense ? B.dense + (size_t)v0 * B.ld : nullptr, B.ld, B.mode, s));
  if (S->smem > 48 * 1024) CX_CHECK(tiled_smem_optin());
  if (S->nblocks[0]) { k_transfer_set<1, false><<<S->nblocks[0], 256, 0, s>>>(S->d_descs[0], S->d_blk2desc[0], nv, S->staged, v0); S->ctx->launches++; }
  if (S->nblocks[1])
  {
    if (S->tiled) k_transfer_set<2, true><<<S->nblocks[1], 256, S->smem, s>>>(S->d_descs[1], S->d_blk2desc[1], nv, S->staged, v0);
    else k_transfer_set<2, false><<<S->nblocks[1], 256, S->smem, s>>>(S->d_descs[1], S->d_blk2desc[1], nv, S->staged, v0);
    S->ctx->launches++;
  }
  if (S->nblocks[2]) { k_transfer_set<3, false><<<S->nblocks[2], 256, S->smem3, s>>>(S->d_descs[2], S->d_blk2desc[2], nv, S->staged, v0); S->ctx->launches++; }
  if (S->nblocks[3]) { k_transfer_set<5, false><<<S->nblocks[3], 256, S->smem5, s>>>(S->d_descs[3], S->d_blk2desc[3], nv, S->staged, v0); S->ctx->launches++; }
  if (S->nblocks[4]) { k_transfer_set<44, false><<<S->nblocks[4], 256, 0, s>>>(S->d_descs[4], S->d_blk2desc[4], nv, S->staged, v0); S->ctx->launches++; }
  if (S->nblocks[5]) { k_transfer_set<22, false><<<S->nblocks[5], 256, 0, s>>>(S->d_descs[5], S->d_blk2desc[5], nv, S->staged, v0); S->ctx->launches++; }
  return 0;
}

int cx_planset_run(cx_planset* S) { return planset_run_on(S, S->ctx->stream); }

int cx_planset_destroy(cx_planset* S)
{
  if (!S) return 0;
  for (int t = 0; t < 6; t++) { cudaFree(S->d_descs[t]); cudaFree(S->d_blk2desc[t]); }
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
    CX_CHECK(upload_ptr_table(ctx, (const double* const*)d_fieldR[i], nvars, &tr));
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
    CX_CHECK(upload_table(ctx, S->d_descs, descs.data(), sizeof(ScatDesc) * descs.size()));
    CX_CHECK(upload_table(ctx, S->d_blk2desc, b2d.data(), sizeof(int) * b2d.size()));
  }
  *out = S;
  return 0;
}

static int scatterset_run_on(cx_scatterset* S, cudaStream_t s, int v0 = 0, int nv = -1)
{
  if (S->nblocks == 0) return 0;
  if (nv < 0) nv = S->nvars - v0;
  k_scatter_set<<<S->nblocks, 256, 0, s>>>(S->d_descs, S->d_blk2desc, v0, nv);
  S->ctx->launches++;
  return 0;
}

int cx_scatterset_run(cx_scatterset* S) { return scatterset_run_on(S, S->ctx->stream); }

// ---- one solver iteration of a rank -------------------------------------------------
//   main stream : fork | local plans (in place)                                                     | join
//   comm stream :      wait fork -> remote plans (dense blocks -> send buffers) -> NCCL send/recv -> scatter |
// so the remote plans, the NVLink exchange and the scatter of iteration i overlap its local transfers.
extern int cx_exchange_on_stream(cx_ctx* ctx, cudaStream_t st, int npeers, const int* peers, const double* const* d_send,
                                 const long long* sendcounts, double* const* d_recv, const long long* recvcounts);

int cx_p2p_signal_on(cudaStream_t s, int npeers, int* const* d_peer_flags, int iter);
int cx_p2p_wait_on(cudaStream_t s, int npeers, const int* d_my_flags, int iter);
int cx_p2p_signal_wait_on(cudaStream_t s, int npeers, int* const* d_peer_flags, const int* d_my_flags, int iter);

struct cx_step {
  cx_ctx* ctx; cx_planset* ps_remote; cx_planset* ps_local; cx_scatterset* ss;
  // peer-memory transport (cx_p2p.cu): plan sets / scatter sets per buffer parity, flags
  int p2p; cx_planset* ps_remote2[2]; cx_scatterset* ss2[2]; int npeers_p2p; int** d_peer_flags; int* d_my_flags; int iter;
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
  S->p2p = 0; S->d_peer_flags = nullptr; S->d_my_flags = nullptr; S->iter = 0; S->npeers_p2p = 0;
  for (int p = 0; p < npeers; p++)
  {
    S->peers.push_back(peers[p]); S->sp.push_back(d_send[p]); S->rp.push_back(d_recv[p]);
    S->sc.push_back(sendcounts[p]); S->rc.push_back(recvcounts[p]);
  }
  *out = S;
  return 0;
}

// Peer-memory step.  ps_remote[par] writes into the peers' receive buffers of parity par (IPC-mapped),
// ss[par] scatters this rank's receive buffers of parity par; peer_flags[p] = address (in peer p's memory) of the
// flag this rank raises, d_my_flags[p] = the flag peer p raises here.
int cx_step_create_p2p(cx_ctx* ctx, cx_planset* ps_remote0, cx_planset* ps_remote1, cx_planset* ps_local,
                       cx_scatterset* ss0, cx_scatterset* ss1, int npeers, int* const* peer_flags, int* d_my_flags,
                       cx_step** out)
{
  CX_CUDA(cudaSetDevice(ctx->device));
  cx_step* S = new cx_step;
  S->ctx = ctx; S->ps_remote = nullptr; S->ps_local = ps_local; S->ss = nullptr; S->overlap = 1;
  S->p2p = 1; S->ps_remote2[0] = ps_remote0; S->ps_remote2[1] = ps_remote1; S->ss2[0] = ss0; S->ss2[1] = ss1;
  S->npeers_p2p = npeers; S->d_my_flags = d_my_flags; S->iter = 0; S->d_peer_flags = nullptr;
  if (npeers > 32) { cx_set_error("peer-memory step: at most 32 peers"); delete S; return -1; }
  if (npeers > 0)
  {
    CX_CUDA(cudaMalloc(&S->d_peer_flags, sizeof(int*) * (size_t)npeers));
    CX_CHECK(upload_table(ctx, S->d_peer_flags, peer_flags, sizeof(int*) * (size_t)npeers));
  }
  *out = S;
  return 0;
}

static int step_run_p2p(cx_step* S, int niter, int overlap, int v0, int nv)
{
  cx_ctx* c = S->ctx;
  cudaStream_t s = c->stream, cs = overlap ? c->comm_stream : c->stream;
  const bool comm = S->npeers_p2p > 0;
  for (int it = 0; it < niter; it++)
  {
    const int iter = ++S->iter, par = iter & 1;
    if (comm && overlap) { CX_CUDA(cudaEventRecord(c->ev_fork, s)); CX_CUDA(cudaStreamWaitEvent(cs, c->ev_fork, 0)); }
    if (S->ps_remote2[par]) CX_CHECK(planset_run_on(S->ps_remote2[par], comm ? cs : s, v0, nv));
    if (comm)
    {
      CX_CHECK(cx_p2p_signal_wait_on(cs, S->npeers_p2p, S->d_peer_flags, S->d_my_flags, iter));
      c->launches += 1;
      if (S->ss2[par]) CX_CHECK(scatterset_run_on(S->ss2[par], cs, v0, nv));
      if (overlap) CX_CUDA(cudaEventRecord(c->ev_join, cs));
    }
    if (S->ps_local) CX_CHECK(planset_run_on(S->ps_local, s, v0, nv));
    if (comm && overlap) CX_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
  }
  return 0;
}

// One iteration restricted to the variables [v0, v0 + nv) of the step's sets (nv < 0: all from v0).  With the
// peer-memory transport a sub-range is a complete iteration of the flag protocol (every rank must issue the same
// sequence of sub-ranges); the NCCL transport ships whole packed buffers, so it only takes the full range.
int cx_step_run_vars(cx_step* S, int niter, int overlap, int v0, int nv)
{
  if (S->p2p) return step_run_p2p(S, niter, overlap, v0, nv);
  cx_ctx* c = S->ctx;
  cudaStream_t s = c->stream, cs = overlap ? c->comm_stream : c->stream;
  const bool comm = !S->peers.empty();
  if (comm && (v0 != 0 || (nv >= 0 && S->ps_local && nv != S->ps_local->nvars)))
  { cx_set_error("cx_step_run_vars: the NCCL transport exchanges all variables at once"); return -1; }
  for (int it = 0; it < niter; it++)
  {
    // fork at the iteration boundary: everything queued on the main stream so far (the solver's update of the
    // donor fields) precedes the remote plans; the send buffers are free again because the previous
    // iteration's exchange ran on the same communication stream
    if (comm && overlap) { CX_CUDA(cudaEventRecord(c->ev_fork, s)); CX_CUDA(cudaStreamWaitEvent(cs, c->ev_fork, 0)); }
    if (S->ps_remote) CX_CHECK(planset_run_on(S->ps_remote, comm ? cs : s, v0, nv));
    if (comm)
    {
      CX_CHECK(cx_exchange_on_stream(c, cs, (int)S->peers.size(), S->peers.data(), S->sp.data(), S->sc.data(),
                                     S->rp.data(), S->rc.data()));
      if (S->ss) CX_CHECK(scatterset_run_on(S->ss, cs, v0, nv));
      if (overlap) CX_CUDA(cudaEventRecord(c->ev_join, cs));
    }
    if (S->ps_local) CX_CHECK(planset_run_on(S->ps_local, s, v0, nv));
    // join: the iteration ends when both the local plans and the scatter of the received values are done
    if (comm && overlap) CX_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
  }
  return 0;
}

int cx_step_run(cx_step* S, int niter, int overlap) { return cx_step_run_vars(S, niter, overlap, 0, -1); }

// 1 when the step can be split into sub-ranges of variables (single rank, or peer-memory transport)
int cx_step_splittable(cx_step* S) { return (S->p2p || S->peers.empty()) ? 1 : 0; }

int cx_step_destroy(cx_step* S)
{
  if (!S) return 0;
  if (S->d_peer_flags) cudaFree(S->d_peer_flags);
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
