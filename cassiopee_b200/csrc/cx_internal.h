// cx_internal.h — host-side objects behind the C ABI (include/cassiopee_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "cx_core.cuh"

void cx_set_error(const char* fmt, ...);

#define CX_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      cx_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(e_), \
                   cudaGetErrorString(e_));                                            \
      return -2;                                                                       \
    }                                                                                  \
  } while (0)

#define CX_CHECK(rc_)                 \
  do {                                \
    int r__ = (rc_);                  \
    if (r__ != 0) return r__;         \
  } while (0)

struct cx_ctx {
  int device;
  int sm_count;
  cudaStream_t stream;
  cudaStream_t comm_stream;     // NCCL exchange + scatter overlap the local transfers
  cudaEvent_t ev_fork, ev_join;
  cudaEvent_t ev0, ev1;
  // pinned staging buffer for host<->device copies
  void* pinned; size_t pinned_bytes;
  cudaEvent_t ev_stage[2];      // staging buffer reuse (cx_upload)
  // counters (kernels launched by this library since creation)
  long long launches;
  // last timed kernel durations (ms), written by entry points that time themselves
  float last_kernel_ms;
  void* nccl;   // cx_comm*
  // host-pointer transfers (cx_api.cu): one grow-only device scratch shared by every plan of the context, and the
  // donor fields uploaded during a batch of transfers (cx_batch_begin .. cx_batch_end), keyed by host pointer
  double* d_tfields; size_t tfields_cap;
  double* d_tdense; size_t tdense_cap;
  double* d_tstage; size_t tstage_cap;   // packed donor values of a sparse upload
  void* fcache;   // cx_field_cache*
  // free page-locked blocks (CX_PIN_BLOCK bytes each) for the zone tables and the counts of searches in flight
  void* pin_pool;   // std::vector<void*>*
  // side streams for work that overlaps across zones (batched ADT builds), created on first use
  cudaStream_t aux[8]; int naux; unsigned aux_next;
  int search_slot;   // 0: searches run on the main stream; k > 0: on side stream k-1 (cx_ctx_search_slot)
};
#define CX_PIN_BLOCK 16384
// a page-locked block of at least `bytes` (from the context's pool when it fits a pool block); give it back with the
// same `own` flag once the device is done with it
void* cx_pin_take(cx_ctx* c, size_t bytes, int* own);
void cx_pin_give(cx_ctx* c, void* p, int own);
int cx_host_threads();

struct cx_zone {
  cx_ctx* ctx;
  int ni, nj, nk, npts, ncells;
  double* d_x; double* d_y; double* d_z; double* d_cellN;   // device coordinates + cellN
  cx::AdtNode* d_nodes;
  int depth;
  int levels;
  double bbox[6];
  float build_ms;
  int topo;           // 1: ADT donor, 0: regular Cartesian donor (InterpCart), no tree
  double h[3];        // topo 0: spacings
  void* adt_pending;  // a build started by cx_adt_begin and not yet finished (AdtBuild*)
};

struct cx_result {
  cx_ctx* ctx;
  int nrcv, nz, ncfmax;
  // per-receptor table (device)
  int* d_noblk; int* d_type; int* d_dind; int* d_isext;
  double* d_cf;       // [ncfmax][nrcv] SoA
  // per-zone compaction
  std::vector<int> cnt, ncoef, nextrap;
  int norphan;
  // search statistics (sum over receptors): nodes visited, candidates, tetra solves
  unsigned long long stats[4];
  float search_ms, extrap_ms, total_ms;
  // a launched search whose counts have not been read yet (cx_search_finish): nothing in cx_search_run waits for
  // the device, so several receptor zones can be in flight
  int pending;
  unsigned long long* h_counts; int h_counts_own;   // page-locked: [nz*3+1] counts + 4 statistics
  void* h_zones; int h_zones_own;                   // page-locked copy of the ZoneView table the upload reads
  cudaEvent_t ev[5];                                // search begin/end, extrapolation begin/end, all done
  void* d_zones_owned;                              // the ZoneView table of the call (freed when the search is done)
  void* d_dw_owned;                                 // double wall: per-zone receptor coordinates + their table
};

struct cx_plan {
  cx_ctx* ctx;
  int imd, jmd, nptsD, nptsR;
  int n;                // entries
  int ngroups;          // type groups
  // entries regrouped by type (stable), SoA coefficients per group
  // layout 0: list order, 1: sorted by donor index, 2: tiled by donor box (type 2 only; d_tiles = ntiles+1 TileRef)
  // layout 3: boxes of donor cells sized per plan (type 2 only; d_tiles = ntiles+1 BoxTile, box_vs = doubles of shared memory per variable)
  // d_pat != nullptr: packed type-2 group (cx_transfer.cu:load_cf2): d_cf holds the 4 distinct doubles of every entry
  // slot-major, d_pat the 16-bit selector word (2 bits per vertex); sel = code of the group's entries in the build's
  // type array (the type, or CX_SEL_PACKED2 for the packable type-2 entries)
  struct Group { int type, ncf, sten, n, first; double* d_cf; int layout; void* d_tiles; int ntiles, ntx, nty, box_vs;
                 unsigned short* d_pat; int sel; };
  std::vector<Group> groups;
  int* d_donor; int* d_rcv; int* d_pos;   // regrouped order; pos = original entry position
  int* d_seq;                             // 0..n-1 (dense output in the plan's own order)
  long long alg_bytes_fixed;              // sum over entries of (8*ncf + 12)
  long long distinct_nodes;               // U (computed on demand)
  float last_ms;
};

// pageable host -> device at bus speed (cx_api.cu); ordered on ctx->stream, host buffer reusable on return
int cx_upload(cx_ctx* ctx, void* dptr, const void* host, size_t bytes);

// scan utilities (cx_scan.cu)
int cx_exclusive_scan_u64(cx_ctx* ctx, const unsigned long long* d_in, unsigned long long* d_out, int n,
                          unsigned long long* h_total);

// ADT build (cx_adt.cu)
int cx_build_adt(cx_zone* z);
int cx_build_adt_many(int nz, cx_zone* const* zones, int nstreams, cudaStream_t* streams);
// start the build of one zone on stream s without waiting for it (the first levels are queued in one go), finish it later
int cx_adt_begin(cx_zone* z, cudaStream_t s);
int cx_adt_finish(cx_zone* z);

// stream-ordered allocation (pool keeps freed blocks: release threshold = max, set at ctx creation);
// pointers obtained here may also be released with plain cudaFree
// Large blocks (> CX_ASYNC_MAX_MB; default: never) go through plain cudaMalloc: first-touch growth of the
// stream-ordered pool is slow for multi-GB blocks; both kinds may be released with cudaFreeAsync/cudaFree.
size_t cx_async_max_bytes();
static inline cudaError_t cx_malloc_async(void** p, size_t bytes, cudaStream_t s)
{
  if (bytes > cx_async_max_bytes()) return cudaMalloc(p, bytes);
  return cudaMallocAsync(p, bytes ? bytes : 8, s);
}
#define CX_MALLOC(ptr, bytes, stream) CX_CUDA(cx_malloc_async((void**)&(ptr), (bytes), (stream)))

static inline int cx_grid(long long n, int block) { return (int)((n + block - 1) / block); }
