// cx_adt.cu — device build of an exact replica of the reference's ADT.
//
// The reference inserts the cells one by one (k,j,i loops, i fastest) into a
// 6-key alternating digital tree rooted at the zone's boundary-cell bounding
// box (KCore/KCore/Interp/InterpAdt.cpp:284-433 buildStructAdt, :506-646
// insert).  The position of a cell in that tree is a pure function of the
// cells inserted before it: it descends by the key/midpoint rule until it
// meets an empty slot.  Equivalently, level by level: among the not-yet-placed
// cells that reach the same empty slot, the one with the smallest insertion
// rank takes it and the others descend through it.  That gives a
// level-synchronous, fully parallel build:
//
//   claim(L)   every active cell evaluates the reference's midpoint recurrence
//              for key L%6 on its own region interval and atomicMin's its rank
//              into the child slot it falls into (warp-aggregated);
//   resolve(L) the winner is placed (its split thresholds are frozen into the
//              node), the others make the winner their new parent.
//
// Node id == insertion rank == cell id, so no allocation is needed.  The
// midpoint recurrences (mid=0.5*(hi+lo); right: hi=2*mid-lo, lo=0.5*(hi+lo))
// are evaluated exactly as the reference does, in fp64 without FMA, because
// the query prunes with the very same numbers (InterpAdt.cpp:768-925).
#include "cx_internal.h"
#include <vector>

using namespace cx;

namespace {

struct BuildScratch { int* actA; int* actB; };

__global__ void k_cell_keys(int ni, int nj, int nk, const double* __restrict__ x, const double* __restrict__ y,
                            const double* __restrict__ z, double* __restrict__ keys, AdtNode* __restrict__ nodes,
                            int ncells, ZoneBox zb, double zmin2d, double zmax2d)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const double zbbox[6] = {zb.lo[0], zb.lo[1], zb.lo[2], zb.hi[3], zb.hi[4], zb.hi[5]};
  adt_cell_keys(ni, nj, nk, x, y, z, c, ncells, keys, nodes, zbbox, zmin2d, zmax2d);
}

__global__ void k_root(AdtNode* nodes, ZoneBox zb)
{
  freeze_node(nodes[0], 0, zb.lo[0], zb.hi[0]);
}

__global__ void k_init_active(int* act, int* cur, int n)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  act[a] = a + 1;
  cur[a + 1] = 0;
}

// cnt[L] = number of active cells at level L (device resident: a batch of levels runs without a host round trip,
// the grids are sized by the count known at the last synchronisation, an upper bound)
__global__ void k_claim(int L, int nact_known, const int* __restrict__ cnt, const int* __restrict__ act, BuildView B)
{
  const int nact = nact_known >= 0 ? nact_known : cnt[L];      // known on the host for the first level of a batch
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  bool on = a < nact;
  int c = 0, p = 0, right = 0;
  if (on) { c = act[a]; adt_claim(B, L, c, p, right); }
  // warp-aggregated atomicMin of the insertion rank on the child slot
  unsigned live = __ballot_sync(0xffffffffu, on);
  if (!on) return;
  int slotKey = p * 2 + right;
  unsigned grp = __match_any_sync(live, slotKey);
  int mn = __reduce_min_sync(grp, c);
  int leader = __ffs(grp) - 1;
  if ((threadIdx.x & 31) == leader)
  {
    atomicMin(&B.child[2 * (size_t)p + (right ? 1 : 0)], mn);
  }
}

__global__ void k_set_children(int ncells, BuildView B)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncells) adt_set_children(B, c);
}

// subtree boxes of the cells placed at one depth: `list` = the cells k_resolve placed at that depth (any order)
__global__ void k_subtree(int count, const int* __restrict__ list, AdtNode* nodes)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= count) return;
  adt_subtree_box(nodes, list ? list[a] : 0);
}

// placed / pcnt: the cells placed at this level are appended to `placed` at n0 - nact + (running count of the level),
// so that the list ends up grouped by depth (the subtree pass then touches each cell once instead of scanning all
// cells once per level)
__global__ void k_resolve(int L, int nact_known, int* __restrict__ cnt, const int* __restrict__ act, BuildView B,
                          int* __restrict__ next, int n0, int* __restrict__ placed, int* __restrict__ pcnt)
{
  const int nact = nact_known >= 0 ? nact_known : cnt[L];
  int* counter = cnt + L + 1;
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  bool keep = false, put = false;
  if (a < nact) { c = act[a]; keep = !adt_resolve(B, L, c); put = !keep; }
  {
    const unsigned mp = __ballot_sync(0xffffffffu, put);
    if (mp)
    {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == __ffs(mp) - 1) base = atomicAdd(&pcnt[L], __popc(mp));
      base = __shfl_sync(0xffffffffu, base, __ffs(mp) - 1);
      if (put) placed[n0 - nact + base + __popc(mp & ((1u << lane) - 1u))] = c;
    }
  }
  // warp-aggregated compaction: one atomicAdd per warp and the lanes keep their relative order, so the active
  // list stays (piecewise) ascending and the next level's accesses to keys/lo/hi[c] stay coalesced
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(&counter[0], __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (keep) next[base + __popc(m & ((1u << lane) - 1u))] = c;
}

}  // namespace

// Zone bounding box over the cells touching the 6 boundary faces
// (K_COMPGEOM::boundingBoxStruct, KCore/KCore/CompGeom/compGeom.cpp:446-...),
// i.e. min/max over the nodes with i in {0,1,ni-2,ni-1} or j ... or k ...
// Host side: the coordinates arrive from the host and the surface is O(n^(2/3)).
void cx_zone_bbox_host(int ni, int nj, int nk, const double* x, const double* y, const double* z, double* bb)
{
  double mn[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
  double mx[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
  const double* arr[3] = {x, y, z};
  auto visit = [&](int i, int j, int k) {
    size_t n = (size_t)i + (size_t)ni * j + (size_t)ni * nj * k;
    for (int d = 0; d < 3; d++)
    {
      double v = arr[d][n];
      if (v < mn[d]) mn[d] = v;
      if (v > mx[d]) mx[d] = v;
    }
  };
  auto onb = [](int a, int n) { return a <= 1 || a >= n - 2; };
  for (int k = 0; k < nk; k++)
    for (int j = 0; j < nj; j++)
    {
      bool jk = onb(j, nj) || onb(k, nk);
      if (jk) { for (int i = 0; i < ni; i++) visit(i, j, k); }
      else
      {
        visit(0, j, k); visit(ni - 1, j, k);
        if (ni > 2) { visit(1, j, k); visit(ni - 2, j, k); }
      }
    }
  bb[0] = mn[0]; bb[1] = mn[1]; bb[2] = mn[2]; bb[3] = mx[0]; bb[4] = mx[1]; bb[5] = mx[2];
}

// One zone's build as a small state machine, so that the builds of several zones can be interleaved by the host and
// overlap on the device (each on its own stream): begin -> { enqueue_batch -> finish_batch } until done -> subtree -> end.
// A 10^6-cell zone is ~100 dependent launches of ~20 us: alone it leaves the GPU idle most of the time.
namespace {
struct AdtBuild {
  cx_zone* z; cx_ctx* ctx; cudaStream_t s;
  BuildView B; BuildScratch S;
  int *d_placed, *d_pcnt, *d_cnt;
  int* h_cnt; int h_own;                 // page-locked: the active counts of a batch of levels
  int *act, *next;
  int nact, n0, L, depth, nb;
  std::vector<int> cntAll;               // cntAll[L] = active cells at the start of level L
  cudaEvent_t ev0, ev1;
  bool done;
  static constexpr int BATCH = 8, BATCH_MAX = 64, MAXL = 4096, TB = 256;
  int batch = BATCH;                     // levels queued per host round trip (small zones)
  bool force_batch = false;              // queue `batch` levels whatever the size (cx_adt_begin: the build overlaps an upload)

  int begin(cx_zone* zone, cudaStream_t stream)
  {
    z = zone; ctx = z->ctx; s = stream; done = false; h_cnt = nullptr; ev0 = ev1 = nullptr;
    const int ncells = z->ncells;
    const size_t n = ncells;
    CX_MALLOC(z->d_nodes, sizeof(AdtNode) * n, s);
    B.nodes = z->d_nodes; B.ncells = ncells;
    CX_MALLOC(B.keys, sizeof(double) * 6 * n, s);
    CX_MALLOC(B.lo, sizeof(double) * 6 * n, s);
    CX_MALLOC(B.hi, sizeof(double) * 6 * n, s);
    CX_MALLOC(B.cur, sizeof(int) * n, s);
    CX_MALLOC(B.br, n, s);
    CX_MALLOC(B.depth, sizeof(int) * n, s);
    CX_CUDA(cudaMemsetAsync(B.depth, 0, sizeof(int) * n, s));
    CX_MALLOC(B.child, sizeof(int) * 2 * n, s);
    CX_CUDA(cudaMemsetAsync(B.child, 0x7f, sizeof(int) * 2 * n, s));      // 0x7f7f7f7f: larger than any cell id
    CX_MALLOC(S.actA, sizeof(int) * n, s);
    CX_MALLOC(S.actB, sizeof(int) * n, s);
    CX_MALLOC(d_placed, sizeof(int) * n, s);
    ZoneBox zb;
    for (int d = 0; d < 6; d++) { zb.lo[d] = z->bbox[d % 3]; zb.hi[d] = z->bbox[3 + d % 3]; }
    B.zb = zb;
    CX_CUDA(cudaEventCreate(&ev0)); CX_CUDA(cudaEventCreate(&ev1));
    CX_CUDA(cudaEventRecord(ev0, s));
    k_cell_keys<<<cx_grid(n, TB), TB, 0, s>>>(z->ni, z->nj, z->nk, z->d_x, z->d_y, z->d_z, B.keys, z->d_nodes, ncells,
                                              zb, z->bbox[2], z->bbox[5]);
    k_root<<<1, 1, 0, s>>>(z->d_nodes, zb);
    ctx->launches += 2;
    nact = ncells - 1;
    if (nact > 0) { k_init_active<<<cx_grid(nact, TB), TB, 0, s>>>(S.actA, B.cur, nact); ctx->launches++; }
    act = S.actA; next = S.actB;
    L = 0; depth = 0;
    CX_MALLOC(d_cnt, sizeof(int) * (MAXL + BATCH_MAX + 2), s);
    CX_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * (MAXL + BATCH_MAX + 2), s));
    h_cnt = (int*)cx_pin_take(ctx, sizeof(int) * (BATCH_MAX + 2), &h_own);
    if (!h_cnt) { cx_set_error("ADT build: no page-locked memory"); return -2; }
    h_cnt[0] = nact;
    CX_CUDA(cudaMemcpyAsync(d_cnt, h_cnt, sizeof(int), cudaMemcpyHostToDevice, s));
    CX_MALLOC(d_pcnt, sizeof(int) * (MAXL + BATCH_MAX + 2), s);
    CX_CUDA(cudaMemsetAsync(d_pcnt, 0, sizeof(int) * (MAXL + BATCH_MAX + 2), s));
    n0 = nact;                        // cells to place below the root
    cntAll.assign(1, nact);
    if (nact <= 0) done = true;
    return 0;
  }

  // small zones: levels run in batches of 8 without a host round trip (the per-level synchronisation costs as
  // much as the level's kernels: 941 192 cells 2.2 ms); large zones keep one level per round trip, so that no
  // kernel is launched over a stale, too large active count (16.6 M cells: 39.2 ms, batched 41.9 ms)
  int enqueue_batch()
  {
    if (done) return 0;
    const int grid = cx_grid(nact, TB);
    nb = (nact > (1 << 21) && !force_batch) ? 1 : batch;
    for (int q = 0; q < nb; q++)
    {
      k_claim<<<grid, TB, 0, s>>>(L + q, q == 0 ? nact : -1, d_cnt, act, B);
      k_resolve<<<grid, TB, 0, s>>>(L + q, q == 0 ? nact : -1, d_cnt, act, B, next, n0, d_placed, d_pcnt);
      ctx->launches += 2;
      int* t = act; act = next; next = t;
    }
    CX_CUDA(cudaMemcpyAsync(h_cnt, d_cnt + L, sizeof(int) * (nb + 1), cudaMemcpyDeviceToHost, s));
    return 0;
  }

  int finish_batch()
  {
    if (done) return 0;
    CX_CUDA(cudaStreamSynchronize(s));
    int used = nb;
    for (int q = 0; q < nb; q++)
    {
      if (h_cnt[q + 1] < h_cnt[q]) depth = L + q + 1;   // somebody was placed at depth L+q+1
      if (h_cnt[q + 1] == 0) { used = q + 1; break; }
    }
    for (int q = 1; q <= used; q++) cntAll.push_back(h_cnt[q]);
    nact = h_cnt[used];
    // an odd number of levels used out of the batch: act/next were swapped nb times, the live list is `used` swaps in
    if ((nb - used) & 1) { int* t = act; act = next; next = t; }
    L += used;
    if (L > MAXL) { cx_set_error("ADT deeper than 4096 levels (degenerate/duplicate cells)"); return -3; }
    if (nact <= 0) done = true;
    return 0;
  }

  // subtree bounding boxes, deepest level first: the cells of depth d were placed during level d - 1 and sit at
  // placed[n0 - cntAll[d-1] .. n0 - cntAll[d]); depth 0 is the root
  int subtree()
  {
    cudaFreeAsync(d_cnt, s); cudaFreeAsync(d_pcnt, s);
    k_set_children<<<cx_grid(z->ncells, TB), TB, 0, s>>>(z->ncells, B);
    ctx->launches++;
    for (int lv = depth - 1; lv >= 1; lv--)
    {
      const int first = n0 - cntAll[lv - 1], count = cntAll[lv - 1] - cntAll[lv];
      if (count <= 0) continue;
      k_subtree<<<cx_grid(count, TB), TB, 0, s>>>(count, d_placed + first, z->d_nodes);
      ctx->launches++;
    }
    if (depth >= 1) { k_subtree<<<1, 32, 0, s>>>(1, nullptr, z->d_nodes); ctx->launches++; }
    CX_CUDA(cudaEventRecord(ev1, s));
    return 0;
  }

  int end()
  {
    CX_CUDA(cudaStreamSynchronize(s));
    CX_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&z->build_ms, ev0, ev1);
    z->depth = depth;
    z->levels = L;
    release();
    if (depth + 2 > CX_STACK)
    {
      cx_set_error("ADT depth %d exceeds the traversal stack (%d)", depth, CX_STACK);
      return -3;
    }
    return 0;
  }

  void release()
  {
    if (!z) return;
    cudaFreeAsync(B.keys, s); cudaFreeAsync(B.lo, s); cudaFreeAsync(B.hi, s); cudaFreeAsync(B.cur, s); cudaFreeAsync(B.br, s); cudaFreeAsync(B.depth, s); cudaFreeAsync(B.child, s);
    cudaFreeAsync(S.actA, s); cudaFreeAsync(S.actB, s); cudaFreeAsync(d_placed, s);
    if (h_cnt) { cx_pin_give(ctx, h_cnt, h_own); h_cnt = nullptr; }
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    ev0 = ev1 = nullptr;
    z = nullptr;
  }
};
}  // namespace

int cx_build_adt(cx_zone* z)
{
  AdtBuild b;
  int rc = b.begin(z, z->ctx->stream);
  while (rc == 0 && !b.done) { rc = b.enqueue_batch(); if (rc == 0) rc = b.finish_batch(); }
  if (rc == 0) rc = b.subtree();
  if (rc == 0) return b.end();
  cudaStreamSynchronize(z->ctx->stream);
  b.release();
  return rc;
}

// A build that overlaps what the host does next (the upload of the following zones of a batch of hooks): the first 48
// levels of a small zone are queued in one go on a side stream (grids sized by the zone, the active counts stay on
// the device: a 100^3 zone has 35 levels), cx_adt_finish picks the build up again, queues what is left and ends it.
int cx_adt_begin(cx_zone* z, cudaStream_t s)
{
  if (z->adt_pending || z->d_nodes) { cx_set_error("ADT build: the zone already has a tree or a build in flight"); return -1; }
  AdtBuild* b = new AdtBuild;
  int rc = b->begin(z, s);
  // (a large zone pays ~7 % for launches over a stale active count -- 39.2 vs 41.9 ms at 16.6 M cells -- and gains the
  // whole upload of the next zone)
  if (rc == 0) { b->batch = 48; b->force_batch = true; rc = b->enqueue_batch(); b->batch = AdtBuild::BATCH; b->force_batch = false; }
  if (rc != 0) { cudaStreamSynchronize(s); b->release(); delete b; return rc; }
  z->adt_pending = b;
  return 0;
}

int cx_adt_finish(cx_zone* z)
{
  AdtBuild* b = (AdtBuild*)z->adt_pending;
  if (!b) return 0;
  z->adt_pending = nullptr;
  int rc = 0;
  if (!b->done) rc = b->finish_batch();
  while (rc == 0 && !b->done) { rc = b->enqueue_batch(); if (rc == 0) rc = b->finish_batch(); }
  if (rc == 0) rc = b->subtree();
  if (rc == 0) rc = b->end();
  else { cudaStreamSynchronize(b->s); b->release(); }
  delete b;
  return rc;
}

// The ADTs of several zones at once: zone q is built on stream streams[q % nstreams]; the host walks the zones round
// after round (a batch of levels each), so the device always has the short kernels of the other zones to run while
// one zone waits for its counts.
int cx_build_adt_many(int nz, cx_zone* const* zones, int nstreams, cudaStream_t* streams)
{
  std::vector<AdtBuild> b(nz);
  int rc = 0;
  for (int q = 0; q < nz && rc == 0; q++) rc = b[q].begin(zones[q], streams[q % nstreams]);
  bool any = true;
  while (rc == 0 && any)
  {
    any = false;
    for (int q = 0; q < nz && rc == 0; q++) if (!b[q].done) rc = b[q].enqueue_batch();
    for (int q = 0; q < nz && rc == 0; q++) if (!b[q].done) { rc = b[q].finish_batch(); any = any || !b[q].done; }
  }
  for (int q = 0; q < nz && rc == 0; q++) rc = b[q].subtree();
  for (int q = 0; q < nz; q++)
  {
    if (rc == 0) rc = b[q].end();
    else if (b[q].z) { cudaStreamSynchronize(b[q].s); b[q].release(); }
  }
  return rc;
}
