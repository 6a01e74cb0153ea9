// cx_api.cu — C ABI of libcassiopee_b200.so (see include/cassiopee_b200.h).
#include "cx_internal.h"
#include <unordered_map>
#include "../../include/cassiopee_b200.h"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>
#include <time.h>

using namespace cx;

static thread_local char g_err[1024] = "";

void cx_set_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

size_t cx_async_max_bytes()
{
  static size_t v = 0;
  if (!v)
  {
    const char* e = getenv("CX_ASYNC_MAX_MB");
    v = (size_t)(e ? atoll(e) : (1 << 20)) << 20;
    if (!v) v = 1;
  }
  return v;
}

// Host -> device copy of a PAGEABLE buffer at bus speed.  cudaMemcpyAsync from pageable memory is staged by the
// driver through a single-threaded bounce copy (5-12 GB/s measured here: 2.1 ms for the 24 MB of one 100^3 zone's
// coordinates, 0.1 s for the 0.5 GB of one C2 donor zone).  Here a team of OpenMP threads (persistent in the runtime:
// no thread start per call) copies chunks into two page-locked staging buffers of the context while the previous
// chunk is on the bus; chunks are sized so that even a 2 MB array is pipelined.  Page-locked or registered sources,
// and small ones, go straight to cudaMemcpyAsync.  The copy is ordered on the context's stream; the host buffer may
// be reused when the function returns.
int cx_upload(cx_ctx* c, void* dptr, const void* host, size_t bytes)
{
  const size_t CHUNK_MAX = 32u << 20;
  if (bytes == 0) return 0;
  bool direct = bytes < (1u << 20);
  if (!direct)
  {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) == cudaSuccess) direct = (at.type == cudaMemoryTypeHost);
    else cudaGetLastError();
  }
  if (direct)
  {
    CX_CUDA(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
  }
  if (c->pinned_bytes < 2 * CHUNK_MAX)
  {
    if (c->pinned) { cudaFreeHost(c->pinned); c->pinned = nullptr; c->pinned_bytes = 0; }
    CX_CUDA(cudaHostAlloc(&c->pinned, 2 * CHUNK_MAX, cudaHostAllocDefault));
    c->pinned_bytes = 2 * CHUNK_MAX;
  }
  if (!c->ev_stage[0])
    for (int i = 0; i < 2; i++) CX_CUDA(cudaEventCreateWithFlags(&c->ev_stage[i], cudaEventDisableTiming));
  static int nt = 0;
  if (!nt)
  {
    const char* e = getenv("CX_UPLOAD_THREADS");
    nt = e ? atoi(e) : std::min(cx_host_threads(), 8);
    if (nt < 1) nt = 1;
  }
  // about 4 chunks per call (staging of chunk i+1 overlaps the DMA of chunk i), 1..32 MB each, 64-byte multiples
  size_t CHUNK = std::min(CHUNK_MAX, std::max((size_t)1 << 20, (bytes / 4 + 63) & ~(size_t)63));
  size_t off = 0;
  int i = 0;
  while (off < bytes)
  {
    const size_t nb = std::min(CHUNK, bytes - off);
    char* buf = (char*)c->pinned + (size_t)(i & 1) * CHUNK_MAX;
    if (i >= 2) CX_CUDA(cudaEventSynchronize(c->ev_stage[i & 1]));     // the copy that last used this buffer is done
    const char* src = (const char*)host + off;
    if (nt == 1) memcpy(buf, src, nb);
    else
    {
#pragma omp parallel for num_threads(nt) schedule(static)
      for (int t = 0; t < nt; t++)
      {
        size_t a = nb * t / nt, b = nb * (t + 1) / nt;
        memcpy(buf + a, src + a, b - a);
      }
    }
    CX_CUDA(cudaMemcpyAsync((char*)dptr + off, buf, nb, cudaMemcpyHostToDevice, c->stream));
    CX_CUDA(cudaEventRecord(c->ev_stage[i & 1], c->stream));
    off += nb; i++;
  }
  // the staging buffers are reused by the next call: wait for the last two copies
  CX_CUDA(cudaEventSynchronize(c->ev_stage[0]));
  if (i >= 2) CX_CUDA(cudaEventSynchronize(c->ev_stage[1]));
  return 0;
}

// from other translation units
struct cx_result_dev {
  int* rcv; int* dnr; int* typ; double* coefs; int* ext; int* orphan;
  std::vector<long long> e_off, c_off, x_off;
};
void cx_zone_bbox_host(int ni, int nj, int nk, const double* x, const double* y, const double* z, double* bb);
int cx_search_run(cx_ctx* ctx, int n, const double* d_xr, const double* d_yr, const double* d_zr, int nz,
                  const ZoneView* d_zones, const ZoneView* h_zones, int order, int nature, int penalty, int extrap,
                  cx_result* R, cx_result_dev* D, const RcvPerZone* d_pz, const RcvPerZone* h_pz);
int cx_search_finish(cx_result* R, cx_result_dev* D);
int cx_plan_build(cx_ctx* ctx, int imd, int jmd, long long nptsD, long long nptsR, int n, const int* donor,
                  const int* rcv, const int* type, const double* coefs, long long ncoefs, cx_plan* P);
int cx_plan_launch(cx_plan* P, int nvars, const double* const* dD, double* const* dR, double* dense, long long ld,
                   int mode, cudaStream_t s);
int cx_plan_launch_celln(cx_plan* P, const double* dCellND, double* dCellNR, cudaStream_t s);
int cx_plan_rotate_on(cx_plan* P, int dirR, double ct, double st, double* fx, double* fy, double* fz, int dense,
                      cudaStream_t s);

struct cx_result_full { cx_result R; cx_result_dev D; };
static inline int result_ready(cx_result* R) { return cx_search_finish(R, &((cx_result_full*)R)->D); }

void* cx_pin_take(cx_ctx* c, size_t bytes, int* own)
{
  void* p = nullptr;
  std::vector<void*>* pool = (std::vector<void*>*)c->pin_pool;
  if (bytes <= CX_PIN_BLOCK && pool)
  {
    *own = 0;
    if (!pool->empty()) { p = pool->back(); pool->pop_back(); return p; }
    if (cudaHostAlloc(&p, CX_PIN_BLOCK, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
  }
  *own = 1;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}

void cx_pin_give(cx_ctx* c, void* p, int own)
{
  if (!p) return;
  if (own || !c->pin_pool) cudaFreeHost(p);
  else ((std::vector<void*>*)c->pin_pool)->push_back(p);
}

// threads of the host-side loops (staging copies, index gathers / scatters): CX_HOST_THREADS, else the cores of the
// box shared between the ranks of the job (LOCAL_WORLD_SIZE), at most 16.  Given to the parallel regions explicitly:
// torchrun exports OMP_NUM_THREADS=1.
int cx_host_threads()
{
  static int nt = 0;
  if (!nt)
  {
    const char* e = getenv("CX_HOST_THREADS");
    if (e) nt = atoi(e);
    else
    {
      unsigned hw = std::thread::hardware_concurrency();
      const char* lw = getenv("LOCAL_WORLD_SIZE");
      int ranks = lw ? std::max(1, atoi(lw)) : 1;
      nt = (int)std::min<unsigned>(16u, std::max(1u, (hw ? hw : 4u) / (unsigned)ranks));
    }
    if (nt < 1) nt = 1;
  }
  return nt;
}

extern "C" {

const char* cx_last_error(void) { return g_err; }
int cx_version(void) { return 100; }

int cx_device_count(int* n)
{
  CX_CUDA(cudaGetDeviceCount(n));
  return 0;
}

int cx_ctx_create(int device, cx_ctx** out)
{
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
  {
    cx_set_error("no CUDA device available (%s); libcassiopee_b200 has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return -2;
  }
  if (device < 0 || device >= ndev) { cx_set_error("device %d out of range (0..%d)", device, ndev - 1); return -1; }
  CX_CUDA(cudaSetDevice(device));
  cx_ctx* c = new cx_ctx;
  memset(c, 0, sizeof(*c));
  c->device = device;
  cudaDeviceProp prop;
  CX_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
    {
      unsigned long long thr = ~0ull;   // keep freed blocks cached in the pool
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  CX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    // the communication stream carries short kernels (remote plans, NCCL, scatter) that must not queue behind the
    // thousands of CTAs of a large local transfer on the main stream: highest priority
    int lo = 0, hi = 0;
    CX_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CX_CUDA(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
  }
  CX_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CX_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  CX_CUDA(cudaEventCreate(&c->ev0));
  CX_CUDA(cudaEventCreate(&c->ev1));
  c->pinned = nullptr; c->pinned_bytes = 0; c->launches = 0; c->nccl = nullptr;
  c->pin_pool = new std::vector<void*>;
  *out = c;
  return 0;
}

static void cx_ctx_free_transfer_scratch(cx_ctx* c);
int cx_ctx_destroy(cx_ctx* c)
{
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->nccl) cx_comm_destroy(c);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->pin_pool)
  {
    for (void* p : *(std::vector<void*>*)c->pin_pool) cudaFreeHost(p);
    delete (std::vector<void*>*)c->pin_pool;
  }
  cx_ctx_free_transfer_scratch(c);
  for (int i = 0; i < 2; i++) if (c->ev_stage[i]) cudaEventDestroy(c->ev_stage[i]);
  for (int i = 0; i < c->naux; i++) cudaStreamDestroy(c->aux[i]);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
  cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join);
  cudaStreamDestroy(c->comm_stream);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int cx_ctx_sync(cx_ctx* c)
{
  CX_CUDA(cudaSetDevice(c->device));
  CX_CUDA(cudaStreamSynchronize(c->stream));
  CX_CUDA(cudaGetLastError());
  return 0;
}

long long cx_ctx_launches(cx_ctx* c) { return c->launches; }

int cx_ctx_info(cx_ctx* c, int* sm_count, size_t* free_bytes, size_t* total_bytes)
{
  CX_CUDA(cudaSetDevice(c->device));
  if (sm_count) *sm_count = c->sm_count;
  size_t f, t;
  CX_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return 0;
}

int cx_event_record(cx_ctx* c, int which)
{
  CX_CUDA(cudaEventRecord(which == 0 ? c->ev0 : c->ev1, c->stream));
  return 0;
}

int cx_event_elapsed_ms(cx_ctx* c, float* ms)
{
  CX_CUDA(cudaEventSynchronize(c->ev1));
  CX_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return 0;
}

static void* g_flush = nullptr;
int cx_flush_l2(cx_ctx* c)
{
  const size_t bytes = 256ull << 20;
  if (!g_flush) CX_CUDA(cudaMalloc(&g_flush, bytes));
  CX_CUDA(cudaMemsetAsync(g_flush, 1, bytes, c->stream));
  return 0;
}

int cx_dev_alloc(cx_ctx* c, size_t bytes, void** dptr)
{
  CX_CUDA(cudaSetDevice(c->device));
  CX_CUDA(cudaMalloc(dptr, bytes ? bytes : 8));
  return 0;
}
int cx_dev_free(cx_ctx* c, void* dptr) { (void)c; CX_CUDA(cudaFree(dptr)); return 0; }
int cx_dev_upload(cx_ctx* c, void* dptr, const void* host, size_t bytes)
{
  CX_CHECK(cx_upload(c, dptr, host, bytes));
  CX_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int cx_dev_download(cx_ctx* c, void* host, const void* dptr, size_t bytes)
{
  CX_CUDA(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
  CX_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int cx_dev_memset(cx_ctx* c, void* dptr, int value, size_t bytes)
{
  CX_CUDA(cudaMemsetAsync(dptr, value, bytes, c->stream));
  return 0;
}
// no synchronisation: the caller orders with cx_ctx_sync (host memory should be pinned or registered)
int cx_dev_upload_async(cx_ctx* c, void* dptr, const void* host, size_t bytes)
{
  CX_CUDA(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, c->stream));
  return 0;
}
int cx_dev_download_async(cx_ctx* c, void* host, const void* dptr, size_t bytes)
{
  CX_CUDA(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}
// page-lock a caller-owned host range in place (numpy arrays of a tree that is transferred every iteration)
int cx_host_register(void* hptr, size_t bytes)
{
  cudaError_t e = cudaHostRegister(hptr, bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 1; }
  if (e != cudaSuccess) { cudaGetLastError(); cx_set_error("cudaHostRegister: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}
int cx_host_unregister(void* hptr)
{
  cudaError_t e = cudaHostUnregister(hptr);
  if (e != cudaSuccess) cudaGetLastError();
  return 0;
}
int cx_pinned_alloc(size_t bytes, void** hptr) { CX_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 8, cudaHostAllocDefault)); return 0; }
int cx_pinned_free(void* hptr) { CX_CUDA(cudaFreeHost(hptr)); return 0; }

/* ---- zones --------------------------------------------------------------- */
static double wall_ms()
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int zone_create_impl(cx_ctx* c, int ni, int nj, int nk, const double* x, const double* y, const double* z,
                            const double* cellN, bool defer_adt, cx_zone** out)
{
  if (ni < 2 || nj < 2 || nk < 1) { cx_set_error("setInterpData: structured donor zones must be 3D or 2D with z=constant."); return -1; }
  if ((long long)ni * nj * nk > 2000000000ll) { cx_set_error("donor zone too large for int32 indices"); return -1; }
  CX_CUDA(cudaSetDevice(c->device));
  const bool prof = getenv("CX_PROFILE_ZONE") != nullptr;
  double t0 = wall_ms();
  cx_zone* zn = new cx_zone;
  memset(zn, 0, sizeof(*zn));
  zn->ctx = c; zn->ni = ni; zn->nj = nj; zn->nk = nk;
  zn->npts = ni * nj * nk;
  zn->ncells = (ni - 1) * (nj - 1) * std::max(nk - 1, 1);
  zn->topo = 1;
  size_t nb = sizeof(double) * (size_t)zn->npts;
  // x, y, z, cellN in one block (one allocation instead of four), from the stream-ordered pool: the blocks of the
  // previous call's hooks are reused instead of going back to the driver (cudaFree of a 30-150 MB block synchronises
  // the device and unmaps; when Python collected the old hooks in the middle of a call that showed as 30-40 ms outliers)
  CX_MALLOC(zn->d_x, 4 * nb, c->stream);
  zn->d_y = zn->d_x + zn->npts; zn->d_z = zn->d_y + zn->npts; zn->d_cellN = zn->d_z + zn->npts;
  double t1 = wall_ms();
  CX_CHECK(cx_upload(c, zn->d_x, x, nb));
  CX_CHECK(cx_upload(c, zn->d_y, y, nb));
  CX_CHECK(cx_upload(c, zn->d_z, z, nb));
  double t2 = wall_ms();
  cx_zone_bbox_host(ni, nj, nk, x, y, z, zn->bbox);
  if (nk == 1)
  {
    // 2-D zone (InterpAdt.cpp:293-301): one z = constant plane, given a fictitious thickness of 1
    if (fabs(zn->bbox[5] - zn->bbox[2]) < 1.e-8) zn->bbox[5] = zn->bbox[2] + 1.;
    else
    {
      cx_set_error("setInterpData: structured donor zones must be 3D or 2D with z=constant.");
      cudaFree(zn->d_x); delete zn;
      return -1;
    }
  }
  double t3 = wall_ms();
  *out = zn;
  int rc = cx_zone_set_celln(zn, cellN);
  double t4 = wall_ms();
  if (rc == 0 && !defer_adt) rc = cx_build_adt(zn);
  if (rc != 0) { cx_zone_destroy(zn); *out = nullptr; return rc; }
  if (prof)
    fprintf(stderr, "[cx_zone_create %dx%dx%d] alloc %.2f ms, coord upload %.2f, bbox %.2f, cellN %.2f, ADT %.2f (device %.2f), total %.2f\n",
            ni, nj, nk, t1 - t0, t2 - t1, t3 - t2, t4 - t3, wall_ms() - t4, zn->build_ms, wall_ms() - t0);
  return 0;
}

int cx_zone_create(cx_ctx* c, int ni, int nj, int nk, const double* x, const double* y, const double* z,
                   const double* cellN, cx_zone** out)
{
  return zone_create_impl(c, ni, nj, nk, x, y, z, cellN, false, out);
}

// the same without the tree: coordinates and cellN go up, the ADT is built later by cx_zones_build_adt together with
// those of the other zones of a batch (a search on a zone without its ADT is refused)
int cx_zone_create_noadt(cx_ctx* c, int ni, int nj, int nk, const double* x, const double* y, const double* z,
                         const double* cellN, cx_zone** out)
{
  return zone_create_impl(c, ni, nj, nk, x, y, z, cellN, true, out);
}

// Start the ADT build of a zone created with cx_zone_create_noadt on a side stream and return at once: the build
// then overlaps the upload of the next zones of a batch of hooks.  cx_zones_build_adt finishes it.
int cx_zone_adt_begin(cx_zone* zn)
{
  cx_ctx* c = zn->ctx;
  if (zn->topo == 0) return 0;
  CX_CUDA(cudaSetDevice(c->device));
  const int k = (int)(c->aux_next++ % 8);
  while (c->naux <= k) { CX_CUDA(cudaStreamCreateWithFlags(&c->aux[c->naux], cudaStreamNonBlocking)); c->naux++; }
  return cx_adt_begin(zn, c->aux[k]);
}

// ADTs of a batch of zones created with cx_zone_create_noadt, built concurrently (one side stream per zone, up to 8);
// builds started with cx_zone_adt_begin are finished; zones that have their tree already, and Cartesian (InterpCart)
// zones, are skipped
int cx_zones_build_adt(cx_ctx* c, int nz, cx_zone* const* zones)
{
  CX_CUDA(cudaSetDevice(c->device));
  std::vector<cx_zone*> todo;
  int rcp = 0;
  for (int i = 0; i < nz; i++)
  {
    if (!zones[i] || zones[i]->ctx != c) { cx_set_error("cx_zones_build_adt: zone %d does not belong to this context", i); return -1; }
    if (zones[i]->adt_pending) { int rc = cx_adt_finish(zones[i]); if (rc != 0 && rcp == 0) rcp = rc; }
    else if (zones[i]->topo != 0 && !zones[i]->d_nodes) todo.push_back(zones[i]);
  }
  if (rcp != 0) return rcp;
  if (todo.empty()) return 0;
  if (todo.size() == 1) return cx_build_adt(todo[0]);
  const int want = (int)std::min<size_t>(8, todo.size());
  while (c->naux < want)
  {
    CX_CUDA(cudaStreamCreateWithFlags(&c->aux[c->naux], cudaStreamNonBlocking));
    c->naux++;
  }
  CX_CUDA(cudaStreamSynchronize(c->stream));      // the coordinates are up
  return cx_build_adt_many((int)todo.size(), todo.data(), want, c->aux);
}

// Regular Cartesian donor located in closed form (interpDataType=0): the analogue of
// new K_INTERP::InterpCart(ni,nj,nk,hi,hj,hk,x0,y0,z0) in Connector/Connector/setInterpData.cpp:790-804
// (spacings from the first nodes of each direction).  Coordinates and cellN are still resident: the
// cell volume, the cellN tests and the extrapolation read them like the reference does.
int cx_zone_create_cart(cx_ctx* c, int ni, int nj, int nk, const double* x, const double* y, const double* z,
                        const double* cellN, cx_zone** out)
{
  if (ni < 2 || nj < 2 || nk < 2) { cx_set_error("Cartesian (InterpCart) donor zones must be 3D on the device path"); return -4; }
  if ((long long)ni * nj * nk > 2000000000ll) { cx_set_error("donor zone too large for int32 indices"); return -1; }
  CX_CUDA(cudaSetDevice(c->device));
  cx_zone* zn = new cx_zone;
  memset(zn, 0, sizeof(*zn));
  zn->ctx = c; zn->ni = ni; zn->nj = nj; zn->nk = nk;
  zn->npts = ni * nj * nk;
  zn->ncells = (ni - 1) * (nj - 1) * (nk - 1);
  zn->topo = 0;
  zn->h[0] = x[1] - x[0]; zn->h[1] = y[ni] - y[0]; zn->h[2] = z[(size_t)ni * nj] - z[0];
  zn->bbox[0] = x[0]; zn->bbox[1] = y[0]; zn->bbox[2] = z[0];
  zn->bbox[3] = zn->bbox[0] + (ni - 1) * zn->h[0];
  zn->bbox[4] = zn->bbox[1] + (nj - 1) * zn->h[1];
  zn->bbox[5] = zn->bbox[2] + (nk - 1) * zn->h[2];
  size_t nb = sizeof(double) * (size_t)zn->npts;
  CX_MALLOC(zn->d_x, 4 * nb, c->stream);
  zn->d_y = zn->d_x + zn->npts; zn->d_z = zn->d_y + zn->npts; zn->d_cellN = zn->d_z + zn->npts;
  CX_CHECK(cx_upload(c, zn->d_x, x, nb));
  CX_CHECK(cx_upload(c, zn->d_y, y, nb));
  CX_CHECK(cx_upload(c, zn->d_z, z, nb));
  *out = zn;
  int rc = cx_zone_set_celln(zn, cellN);
  if (rc == 0 && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = -2;
  if (rc != 0) { cx_zone_destroy(zn); *out = nullptr; return rc; }
  return 0;
}

namespace { __global__ void k_fill(double* p, double v, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; } }

int cx_zone_set_celln(cx_zone* zn, const double* cellN)
{
  cx_ctx* c = zn->ctx;
  if (cellN)
  {
    CX_CHECK(cx_upload(c, zn->d_cellN, cellN, sizeof(double) * (size_t)zn->npts));
    CX_CUDA(cudaStreamSynchronize(c->stream));
  }
  else { k_fill<<<cx_grid(zn->npts, 256), 256, 0, c->stream>>>(zn->d_cellN, 1.0, zn->npts); c->launches++; }
  return 0;
}

int cx_zone_info(cx_zone* zn, double bbox[6], int* ncells, int* depth, float* build_ms)
{
  if (bbox) for (int i = 0; i < 6; i++) bbox[i] = zn->bbox[i];
  if (ncells) *ncells = zn->ncells;
  if (depth) *depth = zn->depth;
  if (build_ms) *build_ms = zn->build_ms;
  return 0;
}

static ZoneView make_view(const cx_zone* zn)
{
  ZoneView V;
  V.x = zn->d_x; V.y = zn->d_y; V.z = zn->d_z; V.cellN = zn->d_cellN; V.nodes = zn->d_nodes;
  V.ni = zn->ni; V.nj = zn->nj; V.nk = zn->nk; V.ncells = zn->ncells; V.depth = zn->depth;
  for (int i = 0; i < 6; i++) V.bbox[i] = zn->bbox[i];
  V.topo = zn->topo;
  for (int i = 0; i < 3; i++)
  {
    V.h[i] = zn->h[i];
    V.hinv[i] = zn->topo == 0 ? 1.0 / zn->h[i] : 0.;     // InterpCart.cpp:51-56
    V.hs2[i] = 0.5 * zn->h[i];
    quant_params(zn->bbox, i, V.qorg[i], V.qscl[i]);
  }
  return V;
}

namespace {
__global__ void k_candidates(ZoneView Z, double x, double y, double z, double alphaTol, int* out, int cap, int* n)
{
  int cnt = 0;
  if (!zone_reject(Z, x, y, z, alphaTol))
  {
    AdtQuery q;
    query_init(q, Z, x, y, z, alphaTol);
    int c, nid, nv = 0;
    while ((c = query_next(q, Z.nodes, &nid, &nv)) >= 0) { if (cnt < cap) out[cnt] = c; cnt++; }
  }
  *n = cnt;
}
}

int cx_zone_candidates(cx_zone* zn, double x, double y, double z, double alphaTol, int* out, int cap, int* n)
{
  cx_ctx* c = zn->ctx;
  if (zn->topo == 0) { cx_set_error("candidate lists are an ADT query; this zone is a Cartesian (InterpCart) donor"); return -1; }
  int *d_out, *d_n;
  CX_CUDA(cudaMalloc(&d_out, sizeof(int) * (size_t)std::max(cap, 1)));
  CX_CUDA(cudaMalloc(&d_n, sizeof(int)));
  k_candidates<<<1, 1, 0, c->stream>>>(make_view(zn), x, y, z, alphaTol, d_out, cap, d_n);
  c->launches++;
  CX_CUDA(cudaMemcpyAsync(n, d_n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CX_CUDA(cudaStreamSynchronize(c->stream));
  int m = std::min(*n, cap);
  if (m > 0) CX_CUDA(cudaMemcpy(out, d_out, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost));
  cudaFree(d_out); cudaFree(d_n);
  return 0;
}

int cx_zone_export_adt(cx_zone* zn, void* nodes_out, size_t bytes)
{
  if (zn->topo == 0) { cx_set_error("Cartesian (InterpCart) donor zones have no ADT"); return -1; }
  size_t need = sizeof(AdtNode) * (size_t)zn->ncells;
  if (bytes < need) { cx_set_error("export buffer too small: need %zu bytes", need); return -1; }
  CX_CUDA(cudaMemcpy(nodes_out, zn->d_nodes, need, cudaMemcpyDeviceToHost));
  return 0;
}

int cx_zone_destroy(cx_zone* zn)
{
  if (!zn) return 0;
  if (zn->adt_pending) cx_adt_finish(zn);
  // d_y, d_z, d_cellN live in d_x's block; pool blocks go back to the pool in stream order, the few allocations too
  // large for it (cx_malloc_async) to the driver
  const size_t bx = 4 * sizeof(double) * (size_t)zn->npts, bn = sizeof(AdtNode) * (size_t)zn->ncells;
  if (zn->d_x) { if (bx > cx_async_max_bytes()) cudaFree(zn->d_x); else cudaFreeAsync(zn->d_x, zn->ctx->stream); }
  if (zn->d_nodes) { if (bn > cx_async_max_bytes()) cudaFree(zn->d_nodes); else cudaFreeAsync(zn->d_nodes, zn->ctx->stream); }
  delete zn;
  return 0;
}

/* ---- setInterpData --------------------------------------------------------- */
static int set_interp_data_impl(cx_ctx* c, int nrcv, const double* d_xr, const double* d_yr, const double* d_zr,
                                const RcvPerZone* d_pz, const RcvPerZone* h_pz, void* dw_block, int nz,
                                cx_zone* const* zones, int order, int nature, int penalty, int extrap, cx_result** out);

int cx_set_interp_data_dev(cx_ctx* c, int nrcv, const double* d_xr, const double* d_yr, const double* d_zr, int nz,
                           cx_zone* const* zones, int order, int nature, int penalty, int extrap, cx_result** out)
{
  return set_interp_data_impl(c, nrcv, d_xr, d_yr, d_zr, nullptr, nullptr, nullptr, nz, zones, order, nature, penalty, extrap, out);
}

// Double wall: connector.setInterpDataDW(receiverArrays, donorArrays, order, nature, penalty, hook)
// (Connector/Connector/setInterpData.cpp:1346-1890): receiver array noz holds the receptor points as donor zone noz
// must see them (projected onto its walls by changeWall); extrapolation is always on, donors are ADT zones.
int cx_set_interp_data_dw(cx_ctx* c, int nrcv, const double* const* xr, const double* const* yr, const double* const* zr,
                          int nz, cx_zone* const* zones, int order, int nature, int penalty, cx_result** out)
{
  if (order != 2 && order != 3 && order != 5) { cx_set_error("setInterpDataDW: order %d (2, 3 or 5)", order); return -4; }
  if (nz <= 0) { cx_set_error("setInterpDataDW: no valid donor zone found."); return -1; }
  CX_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)std::max(nrcv, 1);
  double* blk = nullptr;
  CX_CUDA(cudaMalloc(&blk, sizeof(double) * 3 * n * nz + sizeof(RcvPerZone) * (size_t)nz));
  std::vector<RcvPerZone> pz(nz);
  for (int i = 0; i < nz; i++)
  {
    pz[i].x = blk + (3 * (size_t)i) * n; pz[i].y = blk + (3 * (size_t)i + 1) * n; pz[i].z = blk + (3 * (size_t)i + 2) * n;
    if (nrcv > 0)
    {
      int rcu = cx_upload(c, (void*)pz[i].x, xr[i], sizeof(double) * (size_t)nrcv);
      if (rcu == 0) rcu = cx_upload(c, (void*)pz[i].y, yr[i], sizeof(double) * (size_t)nrcv);
      if (rcu == 0) rcu = cx_upload(c, (void*)pz[i].z, zr[i], sizeof(double) * (size_t)nrcv);
      if (rcu != 0) { cudaFree(blk); return rcu; }
    }
  }
  RcvPerZone* d_pz = (RcvPerZone*)(blk + 3 * n * nz);
  { int rcu = cx_upload(c, d_pz, pz.data(), sizeof(RcvPerZone) * (size_t)nz); if (rcu != 0) { cudaFree(blk); return rcu; } }
  int rc = set_interp_data_impl(c, nrcv, pz[0].x, pz[0].y, pz[0].z, d_pz, pz.data(), blk, nz, zones, order, nature, penalty, 1, out);
  if (rc != 0) cudaFree(blk);
  return rc;
}

static int set_interp_data_impl(cx_ctx* c, int nrcv, const double* d_xr, const double* d_yr, const double* d_zr,
                                const RcvPerZone* d_pz, const RcvPerZone* h_pz, void* dw_block, int nz,
                                cx_zone* const* zones, int order, int nature, int penalty, int extrap, cx_result** out)
{
  if (order != 2 && order != 3 && order != 4 && order != 5)
  { cx_set_error("setInterpData: order %d not supported on the device (2, 3, 4 or 5)", order); return -4; }
  if (nz <= 0) { cx_set_error("setInterpData: no valid donor zone found."); return -1; }
  CX_CUDA(cudaSetDevice(c->device));
  std::vector<ZoneView> hv(nz);
  for (int i = 0; i < nz; i++)
  {
    if (!zones[i] || (zones[i]->topo != 0 && (!zones[i]->d_nodes || zones[i]->adt_pending))) { cx_set_error("setInterpData: donor zone %d has no ADT", i); return -1; }
    hv[i] = make_view(zones[i]);
  }
  // the zone table goes up from a page-locked block of the context's pool and belongs to the result until the search
  // is done; the call itself returns as soon as everything is queued (cx_search_finish, reached
  // through every accessor of the result, waits and reads the counts)
  // searches of different receptor zones may be spread over the side streams (cx_ctx_search_slot): every helper
  // below takes the context's stream, so it is swapped for the duration of the call
  struct StreamSwap {
    cx_ctx* c; cudaStream_t saved;
    StreamSwap(cx_ctx* c_, cudaStream_t s) : c(c_), saved(c_->stream) { c->stream = s; }
    ~StreamSwap() { c->stream = saved; }
  };
  cudaStream_t run_on = c->stream;
  if (c->search_slot > 0)
  {
    const int k = (c->search_slot - 1) % 8;
    while (c->naux <= k) { CX_CUDA(cudaStreamCreateWithFlags(&c->aux[c->naux], cudaStreamNonBlocking)); c->naux++; }
    run_on = c->aux[k];
  }
  StreamSwap swap(c, run_on);
  ZoneView* d_zones;
  CX_MALLOC(d_zones, sizeof(ZoneView) * (size_t)nz, c->stream);
  cx_result_full* F = new cx_result_full;
  F->R.ctx = c; F->R.d_zones_owned = d_zones; F->R.d_dw_owned = nullptr;
  F->R.h_zones = cx_pin_take(c, sizeof(ZoneView) * (size_t)nz, &F->R.h_zones_own);
  if (!F->R.h_zones) { cx_set_error("setInterpData: no page-locked memory for the zone table"); delete F; return -2; }
  memcpy(F->R.h_zones, hv.data(), sizeof(ZoneView) * (size_t)nz);
  CX_CUDA(cudaMemcpyAsync(d_zones, F->R.h_zones, sizeof(ZoneView) * (size_t)nz, cudaMemcpyHostToDevice, c->stream));
  int rc = cx_search_run(c, nrcv, d_xr, d_yr, d_zr, nz, d_zones, hv.data(), order, nature, penalty, extrap, &F->R, &F->D, d_pz, h_pz);
  if (rc != 0) { cx_result_destroy(&F->R); return rc; }
  F->R.d_dw_owned = dw_block;       // the per-zone receptor coordinates live as long as the result
  *out = &F->R;
  return 0;
}

int cx_set_interp_data(cx_ctx* c, int nrcv, const double* xr, const double* yr, const double* zr, int nz,
                       cx_zone* const* zones, int order, int nature, int penalty, int extrap, cx_result** out)
{
  CX_CUDA(cudaSetDevice(c->device));
  double *dx = nullptr, *dy = nullptr, *dz = nullptr;
  size_t nb = sizeof(double) * (size_t)std::max(nrcv, 1);
  CX_CUDA(cudaMalloc(&dx, nb)); CX_CUDA(cudaMalloc(&dy, nb)); CX_CUDA(cudaMalloc(&dz, nb));
  if (nrcv > 0)
  {
    CX_CHECK(cx_upload(c, dx, xr, sizeof(double) * (size_t)nrcv));
    CX_CHECK(cx_upload(c, dy, yr, sizeof(double) * (size_t)nrcv));
    CX_CHECK(cx_upload(c, dz, zr, sizeof(double) * (size_t)nrcv));
  }
  int rc = cx_set_interp_data_dev(c, nrcv, dx, dy, dz, nz, zones, order, nature, penalty, extrap, out);
  cudaStreamSynchronize(c->stream);
  cudaFree(dx); cudaFree(dy); cudaFree(dz);
  return rc;
}

// The searches issued next run on side stream `slot` - 1 (slot 1..8) instead of the main stream, so that the searches
// of several receptor zones overlap on the device (their extrapolation passes are long chains of short work); slot 0
// goes back to the main stream.  The inputs of such a search (receptor coordinates, donor hooks) must be complete
// when it is issued (cx_receptors_create and the zone constructors return with their data on the device).
int cx_ctx_search_slot(cx_ctx* c, int slot)
{
  if (slot < 0 || slot > 8) { cx_set_error("cx_ctx_search_slot: slot 0..8"); return -1; }
  c->search_slot = slot;
  return 0;
}

int cx_result_sizes(cx_result* R, int* cnt, int* ncoef, int* nextrap, int* norphan)
{
  CX_CHECK(result_ready(R));
  for (int z = 0; z < R->nz; z++)
  {
    if (cnt) cnt[z] = R->cnt[z];
    if (ncoef) ncoef[z] = R->ncoef[z];
    if (nextrap) nextrap[z] = R->nextrap[z];
  }
  if (norphan) *norphan = R->norphan;
  return 0;
}

int cx_result_fetch_zone(cx_result* R, int z, int* rcv, int* donor, int* type, double* coefs, int* extrap)
{
  cx_result_dev& D = ((cx_result_full*)R)->D;
  CX_CHECK(result_ready(R));
  if (z < 0 || z >= R->nz) { cx_set_error("zone %d out of range", z); return -1; }
  size_t n = (size_t)R->cnt[z];
  if (n > 0)
  {
    if (rcv) CX_CUDA(cudaMemcpy(rcv, D.rcv + D.e_off[z], sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (donor) CX_CUDA(cudaMemcpy(donor, D.dnr + D.e_off[z], sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (type) CX_CUDA(cudaMemcpy(type, D.typ + D.e_off[z], sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (coefs) CX_CUDA(cudaMemcpy(coefs, D.coefs + D.c_off[z], sizeof(double) * (size_t)R->ncoef[z], cudaMemcpyDeviceToHost));
  }
  if (extrap && R->nextrap[z] > 0)
    CX_CUDA(cudaMemcpy(extrap, D.ext + D.x_off[z], sizeof(int) * (size_t)R->nextrap[z], cudaMemcpyDeviceToHost));
  return 0;
}

int cx_result_fetch_orphans(cx_result* R, int* orphan)
{
  cx_result_dev& D = ((cx_result_full*)R)->D;
  CX_CHECK(result_ready(R));
  if (R->norphan > 0) CX_CUDA(cudaMemcpy(orphan, D.orphan, sizeof(int) * (size_t)R->norphan, cudaMemcpyDeviceToHost));
  return 0;
}

// every list of the call in six copies: rcv / donor / type hold sum(cnt) entries, zones one behind the other in call
// order, coefs sum(ncoef), extrap sum(nextrap), orphan norphan (cx_result_sizes gives the split); the destinations
// may be page-locked (then the copies overlap each other)
int cx_result_fetch_all(cx_result* R, int* rcv, int* donor, int* type, double* coefs, int* extrap, int* orphan)
{
  cx_result_dev& D = ((cx_result_full*)R)->D;
  CX_CHECK(result_ready(R));
  if (R->nrcv == 0) return 0;
  cudaStream_t s = R->ctx->stream;
  const size_t eo = (size_t)D.e_off[R->nz], co = (size_t)D.c_off[R->nz], xo = (size_t)D.x_off[R->nz];
  if (eo && rcv) CX_CUDA(cudaMemcpyAsync(rcv, D.rcv, sizeof(int) * eo, cudaMemcpyDeviceToHost, s));
  if (eo && donor) CX_CUDA(cudaMemcpyAsync(donor, D.dnr, sizeof(int) * eo, cudaMemcpyDeviceToHost, s));
  if (eo && type) CX_CUDA(cudaMemcpyAsync(type, D.typ, sizeof(int) * eo, cudaMemcpyDeviceToHost, s));
  if (co && coefs) CX_CUDA(cudaMemcpyAsync(coefs, D.coefs, sizeof(double) * co, cudaMemcpyDeviceToHost, s));
  if (xo && extrap) CX_CUDA(cudaMemcpyAsync(extrap, D.ext, sizeof(int) * xo, cudaMemcpyDeviceToHost, s));
  if (R->norphan > 0 && orphan) CX_CUDA(cudaMemcpyAsync(orphan, D.orphan, sizeof(int) * (size_t)R->norphan, cudaMemcpyDeviceToHost, s));
  CX_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int cx_result_fetch_raw(cx_result* R, int* noblk, int* type, int* dind, int* isext, double* cf)
{
  CX_CHECK(result_ready(R));
  size_t n = (size_t)R->nrcv;
  if (n == 0) return 0;
  if (noblk) CX_CUDA(cudaMemcpy(noblk, R->d_noblk, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (type) CX_CUDA(cudaMemcpy(type, R->d_type, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (dind) CX_CUDA(cudaMemcpy(dind, R->d_dind, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (isext) CX_CUDA(cudaMemcpy(isext, R->d_isext, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (cf) CX_CUDA(cudaMemcpy(cf, R->d_cf, sizeof(double) * n * R->ncfmax, cudaMemcpyDeviceToHost));
  return 0;
}

int cx_result_stats(cx_result* R, double* out)
{
  CX_CHECK(result_ready(R));
  out[0] = R->search_ms; out[1] = R->extrap_ms;
  out[2] = (double)R->stats[0]; out[3] = (double)R->stats[1]; out[4] = (double)R->ncfmax;
  out[5] = (double)R->stats[2]; out[6] = 0; out[7] = 0;
  return 0;
}

int cx_result_destroy(cx_result* R)
{
  if (!R) return 0;
  cx_result_full* F = (cx_result_full*)R;
  result_ready(R);
  cudaStream_t s = R->ctx->stream;
  if (R->d_zones_owned) cudaFreeAsync(R->d_zones_owned, s);
  if (R->h_zones) { cx_pin_give(R->ctx, R->h_zones, R->h_zones_own); R->h_zones = nullptr; }
  if (R->d_dw_owned) cudaFree(R->d_dw_owned);
  for (int i = 0; i < 5; i++) if (R->ev[i]) cudaEventDestroy(R->ev[i]);
  void* ptrs[] = {R->d_noblk, R->d_type, R->d_dind, R->d_isext, R->d_cf, F->D.rcv, F->D.dnr, F->D.typ, F->D.coefs,
                  F->D.ext, F->D.orphan};
  for (void* p : ptrs) if (p) cudaFreeAsync(p, s);
  delete F;
  return 0;
}

/* ---- transfers ------------------------------------------------------------- */
struct cx_plan_full {
  cx_plan P;
  // scratch for the host-pointer entry points
  double* h_dense; size_t hdense_cap;   // pinned
  std::vector<int> h_rcv;               // original-order receptor indices (host scatter)
  bool rcv_contiguous;                  // h_rcv[e] == h_rcv[0] + e
  std::vector<cudaEvent_t> evk, evd;    // per-variable kernel-done / download-done events (host-pointer pipeline)
  // sparse upload of the donor fields (host-pointer pipeline, large zones): the donor points the plan reads
  int didx_state = 0;                   // 0: not looked at yet, 1: sparse (h_didx / d_didx), 2: whole fields, 3: box
  int box[6] = {0, 0, 0, 0, 0, 0};      // state 3: i0, i1, j0, j1, k0, k1 (inclusive) of the donor points read
  std::vector<int> h_didx; int* d_didx = nullptr;
  double* h_stage = nullptr; size_t hstage_cap = 0;   // pinned, nvars x h_didx.size()
};

int cx_plan_create(cx_ctx* c, int imd, int jmd, int kmd, long long nptsR, int n, const int* donor, const int* rcv,
                   const int* type, const double* coefs, long long ncoefs, cx_plan** out)
{
  CX_CUDA(cudaSetDevice(c->device));
  cx_plan_full* F = new cx_plan_full;
  F->h_dense = nullptr; F->hdense_cap = 0;
  int rc = cx_plan_build(c, imd, jmd, (long long)imd * jmd * kmd, nptsR, n, donor, rcv, type, coefs, ncoefs, &F->P);
  if (rc != 0) { cx_plan_destroy(&F->P); return rc; }
  F->h_rcv.assign(rcv, rcv + n);
  F->rcv_contiguous = n > 0;
  for (int e = 1; e < n && F->rcv_contiguous; e++) if (rcv[e] != rcv[0] + e) F->rcv_contiguous = false;
  *out = &F->P;
  return 0;
}

static int ensure(double** p, size_t* cap, size_t need, bool pinned)
{
  if (*cap >= need) return 0;
  if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; }
  if (pinned) CX_CUDA(cudaHostAlloc((void**)p, need, cudaHostAllocDefault));
  else CX_CUDA(cudaMalloc((void**)p, need));
  *cap = need;
  return 0;
}

// Host-pointer transfers are PCIe-bound (8*nvars*(nptsD + n) bytes cross the bus for ~0.5 ms of kernel), so they
// run as a per-variable pipeline over two streams: variable v's upload + kernel on the main stream, its download
// on the copy stream as soon as the kernel's event fires.  Uploads and downloads then overlap (the bus is full
// duplex) and the total tends to max(H2D, D2H) instead of their sum.  The per-variable launches re-read the
// plan's coefficients nvars times from HBM, which costs ~0.25 ms per variable at C2a size against ~2.5 ms of bus
// time per variable.
// Donor fields uploaded during a batch of host-pointer transfers (one _setInterpTransfers call over a tree): the k
// subregions of a donor zone read the same host arrays, which then cross the bus once instead of k times.  A host
// array a transfer writes into (receptor field) is dropped from the cache, so a later subregion that reads it as a
// donor field sees the updated values, like the reference's sequential loop over the subregions.
struct cx_field_cache {
  struct Buf { double* d; size_t bytes; };
  std::unordered_map<const void*, Buf> map;
  size_t bytes = 0, budget = (size_t)8 << 30;
  long long uploads = 0;
  bool on = false;
};

static int field_cache_clear(cx_ctx* c)
{
  cx_field_cache* FC = (cx_field_cache*)c->fcache;
  if (!FC) return 0;
  for (auto& kv : FC->map) cudaFreeAsync(kv.second.d, c->stream);
  FC->map.clear(); FC->bytes = 0;
  return 0;
}

static void field_cache_written(cx_ctx* c, int nvars, double* const* fieldR)
{
  cx_field_cache* FC = (cx_field_cache*)c->fcache;
  if (!FC || !FC->on || FC->map.empty()) return;
  for (int v = 0; v < nvars; v++)
  {
    auto it = FC->map.find(fieldR[v]);
    if (it != FC->map.end()) { cudaFreeAsync(it->second.d, c->stream); FC->bytes -= it->second.bytes; FC->map.erase(it); }
  }
}

int cx_batch_begin(cx_ctx* c, long long budget_bytes)
{
  CX_CUDA(cudaSetDevice(c->device));
  if (!c->fcache) c->fcache = new cx_field_cache;
  cx_field_cache* FC = (cx_field_cache*)c->fcache;
  CX_CHECK(field_cache_clear(c));
  if (budget_bytes > 0) FC->budget = (size_t)budget_bytes;
  FC->on = true; FC->uploads = 0;
  return 0;
}

// drop the cached fields (called between donor zones: their fields are not read again); the batch stays open
int cx_batch_flush(cx_ctx* c)
{
  CX_CUDA(cudaSetDevice(c->device));
  return field_cache_clear(c);
}

int cx_batch_end(cx_ctx* c, long long* uploads)
{
  CX_CUDA(cudaSetDevice(c->device));
  cx_field_cache* FC = (cx_field_cache*)c->fcache;
  if (uploads) *uploads = FC ? FC->uploads : 0;
  CX_CHECK(field_cache_clear(c));
  if (FC) FC->on = false;
  return 0;
}

static void cx_ctx_free_transfer_scratch(cx_ctx* c)
{
  field_cache_clear(c);
  delete (cx_field_cache*)c->fcache; c->fcache = nullptr;
  if (c->d_tfields) cudaFree(c->d_tfields);
  if (c->d_tdense) cudaFree(c->d_tdense);
  if (c->d_tstage) cudaFree(c->d_tstage);
  c->d_tfields = c->d_tdense = c->d_tstage = nullptr; c->tfields_cap = c->tdense_cap = c->tstage_cap = 0;
}

static int pipeline_events(cx_plan_full* F, int nvars)
{
  while ((int)F->evk.size() < nvars)
  {
    cudaEvent_t a, b;
    CX_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CX_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    F->evk.push_back(a); F->evd.push_back(b);
  }
  return 0;
}

// Periodicity by rotation (includeTransfers.h): the vector triplets (indices into the variable list) are rotated
// on the device, in their dense blocks, before they are downloaded.
struct RotSpec { int dirR; double ct, st; int ntrip; const int* trip; };

namespace {
__global__ void k_unpack1(int n, const double* __restrict__ in, const int* __restrict__ idx, double* __restrict__ out)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[idx[e]] = in[e];
}
}

// Which donor points does the plan read?  For a large donor zone of which the plan reads less than 1/4 (fringe
// plans), only those cross the bus: gathered on the host by an OpenMP team into a page-locked block, copied, and
// stored by index into the device copy of the field (the rest of that copy is never read).  Decided once per plan
// (mark kernel + one mask download + a host compaction).  CX_SPARSE_H2D=0: never.  Measured on C2a, where the plan
// reads 53 % of the background: 19.7 ms per call against 17.5 ms with whole fields (gpurun_out/r3_e2e_*.json) -- the
// host gather walks the whole array at less than the DMA engine's rate -- hence the 1/4.
static int plan_donor_list(cx_plan_full* F)
{
  cx_plan* P = &F->P; cx_ctx* c = P->ctx;
  F->didx_state = 2;
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("CX_SPARSE_H2D"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
  if (!enabled || P->nptsD < (1 << 20)) return 0;
  unsigned char* d_mark;
  CX_MALLOC(d_mark, (size_t)P->nptsD, c->stream);
  CX_CUDA(cudaMemsetAsync(d_mark, 0, (size_t)P->nptsD, c->stream));
  CX_CHECK(cx_plan_mark_donors(P, d_mark));
  std::vector<unsigned char> mark((size_t)P->nptsD);
  CX_CUDA(cudaMemcpyAsync(mark.data(), d_mark, (size_t)P->nptsD, cudaMemcpyDeviceToHost, c->stream));
  CX_CUDA(cudaStreamSynchronize(c->stream));
  cudaFreeAsync(d_mark, c->stream);
  size_t nd = 0;
  for (size_t i = 0; i < mark.size(); i++) nd += mark[i];
  if (nd == 0) return 0;
  if (4 * nd >= (size_t)P->nptsD)
  {
    // too dense for the index list: the bounding box of the points read (a near-body grid inside a background reads a
    // block of it) goes up as ONE strided copy per variable when it holds less than 85 % of the zone
    const int imd = P->imd, jmd = P->jmd, kmd = (int)(P->nptsD / ((long long)imd * jmd));
    int b[6] = {imd, -1, jmd, -1, kmd, -1};
    for (int k = 0; k < kmd; k++)
      for (int j = 0; j < jmd; j++)
      {
        const unsigned char* row = mark.data() + ((size_t)k * jmd + j) * imd;
        int lo = -1, hi = -1;
        for (int i = 0; i < imd; i++) if (row[i]) { if (lo < 0) lo = i; hi = i; }
        if (lo < 0) continue;
        b[0] = std::min(b[0], lo); b[1] = std::max(b[1], hi);
        b[2] = std::min(b[2], j); b[3] = std::max(b[3], j); b[4] = std::min(b[4], k); b[5] = std::max(b[5], k);
      }
    const double vol = (double)(b[1] - b[0] + 1) * (b[3] - b[2] + 1) * (b[5] - b[4] + 1);
    static int box_on = -1;
    if (box_on < 0) { const char* e = getenv("CX_BOX_H2D"); box_on = (e && atoi(e) == 0) ? 0 : 1; }
    if (box_on && vol <= 0.85 * (double)P->nptsD) { for (int q = 0; q < 6; q++) F->box[q] = b[q]; F->didx_state = 3; }
    return 0;
  }
  F->h_didx.resize(nd);
  size_t k = 0;
  for (size_t i = 0; i < mark.size(); i++) if (mark[i]) F->h_didx[k++] = (int)i;
  CX_CUDA(cudaMalloc(&F->d_didx, sizeof(int) * nd));
  CX_CHECK(cx_upload(c, F->d_didx, F->h_didx.data(), sizeof(int) * nd));
  F->didx_state = 1;
  return 0;
}

// dst[v]: host destination of variable v's dense (n) block
static int transfer_pipeline(cx_plan_full* F, int nvars, const double* const* fieldD, double* const* dst,
                             const RotSpec* rot = nullptr)
{
  cx_plan* P = &F->P; cx_ctx* c = P->ctx;
  CX_CUDA(cudaSetDevice(c->device));
  const size_t per = (size_t)P->nptsD, n = (size_t)P->n;
  // scratch of the context, not of the plan (a tree with k subregions per donor zone would otherwise hold k
  // copies of the zone's fields); calls are synchronous, so one buffer serves every plan
  cx_field_cache* FC = (cx_field_cache*)c->fcache;
  const bool batch = FC && FC->on;
  if (!batch) CX_CHECK(ensure(&c->d_tfields, &c->tfields_cap, sizeof(double) * per * nvars, false));
  if (!batch && F->didx_state == 0) CX_CHECK(plan_donor_list(F));
  const bool sparse = !batch && F->didx_state == 1;
  const size_t nd = F->h_didx.size();
  if (sparse)
  {
    CX_CHECK(ensure(&F->h_stage, &F->hstage_cap, sizeof(double) * nd * nvars, true));
    CX_CHECK(ensure(&c->d_tstage, &c->tstage_cap, sizeof(double) * nd * nvars, false));
  }
  CX_CHECK(ensure(&c->d_tdense, &c->tdense_cap, sizeof(double) * n * nvars, false));
  double* const d_dense = c->d_tdense;
  CX_CHECK(pipeline_events(F, nvars));
  cudaStream_t sa = c->stream, sb = c->comm_stream;
  std::vector<char> hold(nvars, 0);        // downloads deferred until the rotation has run
  if (rot && rot->dirR != 0)
    for (int t = 0; t < 3 * rot->ntrip; t++)
    {
      if (rot->trip[t] < 0 || rot->trip[t] >= nvars) { cx_set_error("cx_transfer_rot: variable index out of range"); return -2; }
      hold[rot->trip[t]] = 1;
    }
  for (int v = 0; v < nvars; v++)
  {
    double* dv = nullptr;
    if (batch)
    {
      // donor field already on the device from an earlier subregion of the same batch?
      auto it = FC->map.find(fieldD[v]);
      if (it != FC->map.end() && it->second.bytes == sizeof(double) * per) dv = it->second.d;
      else
      {
        if (it != FC->map.end()) { cudaFreeAsync(it->second.d, sa); FC->bytes -= it->second.bytes; FC->map.erase(it); }
        if (FC->bytes + sizeof(double) * per > FC->budget) CX_CHECK(field_cache_clear(c));
        CX_MALLOC(dv, sizeof(double) * per, sa);
        CX_CUDA(cudaMemcpyAsync(dv, fieldD[v], sizeof(double) * per, cudaMemcpyHostToDevice, sa));
        FC->map[fieldD[v]] = cx_field_cache::Buf{dv, sizeof(double) * per};
        FC->bytes += sizeof(double) * per;
        FC->uploads++;
      }
    }
    else if (sparse)
    {
      // the team gathers variable v while the copies of the variables before it are on the bus
      dv = c->d_tfields + per * v;
      double* hs = F->h_stage + nd * v; double* ds = c->d_tstage + nd * v;
      cx_host_gather(hs, F->h_didx.data(), fieldD[v], (long long)nd);
      CX_CUDA(cudaMemcpyAsync(ds, hs, sizeof(double) * nd, cudaMemcpyHostToDevice, sa));
      k_unpack1<<<cx_grid((long long)nd, 256), 256, 0, sa>>>((int)nd, ds, F->d_didx, dv);
      c->launches++;
    }
    else if (!batch && F->didx_state == 3)
    {
      // the box of the donor points read, as one strided copy (rows of the box, plane after plane)
      dv = c->d_tfields + per * v;
      const int* b = F->box;
      cudaMemcpy3DParms p;
      memset(&p, 0, sizeof(p));
      p.srcPtr = make_cudaPitchedPtr((void*)fieldD[v], sizeof(double) * (size_t)P->imd, sizeof(double) * (size_t)P->imd, (size_t)P->jmd);
      p.dstPtr = make_cudaPitchedPtr((void*)dv, sizeof(double) * (size_t)P->imd, sizeof(double) * (size_t)P->imd, (size_t)P->jmd);
      p.srcPos = make_cudaPos(sizeof(double) * (size_t)b[0], (size_t)b[2], (size_t)b[4]);
      p.dstPos = p.srcPos;
      p.extent = make_cudaExtent(sizeof(double) * (size_t)(b[1] - b[0] + 1), (size_t)(b[3] - b[2] + 1), (size_t)(b[5] - b[4] + 1));
      p.kind = cudaMemcpyHostToDevice;
      CX_CUDA(cudaMemcpy3DAsync(&p, sa));
    }
    else
    {
      dv = c->d_tfields + per * v;
      CX_CUDA(cudaMemcpyAsync(dv, fieldD[v], sizeof(double) * per, cudaMemcpyHostToDevice, sa));
    }
    const double* dvp = dv;
    CX_CHECK(cx_plan_launch(P, 1, &dvp, nullptr, d_dense + n * v, (long long)n, 1, sa));
    if (hold[v]) continue;
    CX_CUDA(cudaEventRecord(F->evk[v], sa));
    CX_CUDA(cudaStreamWaitEvent(sb, F->evk[v], 0));
    CX_CUDA(cudaMemcpyAsync(dst[v], d_dense + n * v, sizeof(double) * n, cudaMemcpyDeviceToHost, sb));
    CX_CUDA(cudaEventRecord(F->evd[v], sb));
  }
  if (rot && rot->dirR != 0)
  {
    for (int t = 0; t < rot->ntrip; t++)
    {
      const int* tr = rot->trip + 3 * t;
      CX_CHECK(cx_plan_rotate_on(P, rot->dirR, rot->ct, rot->st, d_dense + n * tr[0], d_dense + n * tr[1],
                                 d_dense + n * tr[2], 1, sa));
    }
    for (int v = 0; v < nvars; v++)
    {
      if (!hold[v]) continue;
      CX_CUDA(cudaEventRecord(F->evk[v], sa));
      CX_CUDA(cudaStreamWaitEvent(sb, F->evk[v], 0));
      CX_CUDA(cudaMemcpyAsync(dst[v], d_dense + n * v, sizeof(double) * n, cudaMemcpyDeviceToHost, sb));
      CX_CUDA(cudaEventRecord(F->evd[v], sb));
    }
  }
  return 0;
}

int cx_transferD(cx_plan* P, int nvars, const double* const* fieldD, double* out)
{
  cx_plan_full* F = (cx_plan_full*)P; cx_ctx* c = P->ctx;
  if (P->n == 0 || nvars == 0) return 0;
  std::vector<double*> dst(nvars);
  for (int v = 0; v < nvars; v++) dst[v] = out + (size_t)P->n * v;
  CX_CHECK(transfer_pipeline(F, nvars, fieldD, dst.data()));
  CX_CUDA(cudaStreamSynchronize(c->comm_stream));
  CX_CUDA(cudaStreamSynchronize(c->stream));
  CX_CUDA(cudaGetLastError());
  return 0;
}

static int transfer_host(cx_plan* P, int nvars, const double* const* fieldD, double* const* fieldR, const RotSpec* rot);

int cx_transfer(cx_plan* P, int nvars, const double* const* fieldD, double* const* fieldR)
{
  return transfer_host(P, nvars, fieldD, fieldR, nullptr);
}

int cx_transfer_rot(cx_plan* P, int nvars, const double* const* fieldD, double* const* fieldR, int dirR, double ctheta,
                    double stheta, int ntrip, const int* trip)
{
  if (dirR < 0 || dirR > 3) { cx_set_error("cx_transfer_rot: dirR must be 0..3"); return -2; }
  RotSpec rot{dirR, ctheta, stheta, ntrip, trip};
  return transfer_host(P, nvars, fieldD, fieldR, &rot);
}

int cx_rotate_dev(cx_plan* P, int dirR, double ctheta, double stheta, double* d_fx, double* d_fy, double* d_fz)
{
  if (dirR < 1 || dirR > 3) { cx_set_error("cx_rotate_dev: dirR must be 1..3"); return -2; }
  if (P->n == 0) return 0;
  return cx_plan_rotate_on(P, dirR, ctheta, stheta, d_fx, d_fy, d_fz, 0, P->ctx->stream);
}

static int transfer_host(cx_plan* P, int nvars, const double* const* fieldD, double* const* fieldR, const RotSpec* rot)
{
  cx_plan_full* F = (cx_plan_full*)P; cx_ctx* c = P->ctx;
  if (P->n == 0 || nvars == 0) return 0;
  const int n = P->n;
  std::vector<double*> dst(nvars);
  if (F->rcv_contiguous)
  {
    // receptor indices are one ascending run (every point of the zone, or one slab): the dense block of a
    // variable IS the receptor field's slice, so it is downloaded in place without a host scatter
    for (int v = 0; v < nvars; v++) dst[v] = fieldR[v] + F->h_rcv[0];
    CX_CHECK(transfer_pipeline(F, nvars, fieldD, dst.data(), rot));
    CX_CUDA(cudaStreamSynchronize(c->comm_stream));
    CX_CUDA(cudaStreamSynchronize(c->stream));
    CX_CUDA(cudaGetLastError());
    field_cache_written(c, nvars, fieldR);
    return 0;
  }
  CX_CHECK(ensure(&F->h_dense, &F->hdense_cap, sizeof(double) * (size_t)n * nvars, true));
  for (int v = 0; v < nvars; v++) dst[v] = F->h_dense + (size_t)n * v;
  CX_CHECK(transfer_pipeline(F, nvars, fieldD, dst.data(), rot));
  // host scatter of variable v while the later variables are still on the bus; receptors of one plan are
  // distinct, so entry ranges can go to different threads
  const int* rcv = F->h_rcv.data();
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::min<unsigned>(hw ? hw : 1, 16);
  if (n < (1 << 18)) nt = 1;
  for (int v = 0; v < nvars; v++)
  {
    CX_CUDA(cudaEventSynchronize(F->evd[v]));
    double* fr = fieldR[v];
    const double* src = F->h_dense + (size_t)v * n;
    auto work = [&](int t) {
      int e0 = (int)((long long)n * t / nt), e1 = (int)((long long)n * (t + 1) / nt);
      for (int e = e0; e < e1; e++) fr[rcv[e]] = src[e];
    };
    if (nt == 1) work(0);
    else
    {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) th.emplace_back(work, t);
      for (auto& x : th) x.join();
    }
  }
  CX_CUDA(cudaStreamSynchronize(c->comm_stream));
  CX_CUDA(cudaStreamSynchronize(c->stream));
  CX_CUDA(cudaGetLastError());
  field_cache_written(c, nvars, fieldR);
  return 0;
}

int cx_transfer_celln(cx_plan* P, const double* cellND, double* cellNR)
{
  cx_ctx* c = P->ctx;
  if (P->n == 0) return 0;
  CX_CUDA(cudaSetDevice(c->device));
  double *dD, *dR;
  CX_CUDA(cudaMalloc(&dD, sizeof(double) * (size_t)P->nptsD));
  CX_CUDA(cudaMalloc(&dR, sizeof(double) * (size_t)P->nptsR));
  CX_CUDA(cudaMemcpyAsync(dD, cellND, sizeof(double) * (size_t)P->nptsD, cudaMemcpyHostToDevice, c->stream));
  CX_CUDA(cudaMemcpyAsync(dR, cellNR, sizeof(double) * (size_t)P->nptsR, cudaMemcpyHostToDevice, c->stream));
  CX_CHECK(cx_plan_launch_celln(P, dD, dR, c->stream));
  CX_CUDA(cudaMemcpyAsync(cellNR, dR, sizeof(double) * (size_t)P->nptsR, cudaMemcpyDeviceToHost, c->stream));
  CX_CUDA(cudaStreamSynchronize(c->stream));
  CX_CUDA(cudaGetLastError());
  cudaFree(dD); cudaFree(dR);
  return 0;
}

int cx_transfer_dev(cx_plan* P, int nvars, const double* const* d_fieldD, double* const* d_fieldR)
{
  if (P->n == 0 || nvars == 0) return 0;
  return cx_plan_launch(P, nvars, d_fieldD, d_fieldR, nullptr, 0, 0, P->ctx->stream);
}

int cx_transferD_dev(cx_plan* P, int nvars, const double* const* d_fieldD, double* d_out, long long ld)
{
  if (P->n == 0 || nvars == 0) return 0;
  return cx_plan_launch(P, nvars, d_fieldD, nullptr, d_out, ld, 1, P->ctx->stream);
}

int cx_transfer_celln_dev(cx_plan* P, const double* d_cellND, double* d_cellNR)
{
  if (P->n == 0) return 0;
  return cx_plan_launch_celln(P, d_cellND, d_cellNR, P->ctx->stream);
}

namespace {
struct RPtrs { double* r[16]; };
__global__ void k_scatter(int n, int nvars, const double* __restrict__ in, long long ld, const int* __restrict__ rcv, RPtrs R)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int r = rcv[e];
  for (int v = 0; v < nvars; v++) R.r[v][r] = in[(size_t)v * ld + e];
}
__global__ void k_mark(int m, int sten_o, int sten_k, int imd, int imdjmd, const int* __restrict__ donor, unsigned char* mark)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  int d0 = donor[e];
  for (int kk = 0; kk < sten_k; kk++) for (int jj = 0; jj < sten_o; jj++) for (int ii = 0; ii < sten_o; ii++)
    mark[d0 + ii + jj * imd + kk * imdjmd] = 1;
}
__global__ void k_count(long long n, const unsigned char* mark, unsigned long long* cnt)
{
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned v = (i < n && mark[i]) ? 1u : 0u;
  unsigned b = __ballot_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(cnt, (unsigned long long)__popc(b));
}
}

int cx_scatter_dev(cx_ctx* c, int nvars, const double* d_in, long long ld, int n, const int* d_rcv, double* const* d_fieldR)
{
  if (n == 0) return 0;
  for (int v0 = 0; v0 < nvars; v0 += 16)
  {
    int nv = std::min(16, nvars - v0);
    RPtrs R; for (int v = 0; v < nv; v++) R.r[v] = d_fieldR[v0 + v];
    k_scatter<<<cx_grid(n, 256), 256, 0, c->stream>>>(n, nv, d_in + (size_t)v0 * ld, ld, d_rcv, R);
    c->launches++;
  }
  return 0;
}

namespace {
__global__ void k_gather(int n, int nvars, double* __restrict__ out, long long ld, const int* __restrict__ idx, RPtrs R)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int r = idx[e];
  for (int v = 0; v < nvars; v++) out[(size_t)v * ld + e] = R.r[v][r];
}
}

// out[v*ld + e] = field[v][idx[e]]: the receptor values of a zone packed for a sparse download
int cx_gather_dev(cx_ctx* c, int nvars, double* d_out, long long ld, int n, const int* d_idx, double* const* d_field)
{
  if (n == 0) return 0;
  for (int v0 = 0; v0 < nvars; v0 += 16)
  {
    int nv = std::min(16, nvars - v0);
    RPtrs R; for (int v = 0; v < nv; v++) R.r[v] = d_field[v0 + v];
    k_gather<<<cx_grid(n, 256), 256, 0, c->stream>>>(n, nv, d_out + (size_t)v0 * ld, ld, d_idx, R);
    c->launches++;
  }
  return 0;
}

// host side of a sparse download: dst[idx[e]] = src[e] (OpenMP team; idx holds distinct indices)
int cx_host_scatter(double* dst, const int* idx, const double* src, long long n)
{
  if (n < (1 << 15))
  {
    for (long long e = 0; e < n; e++) dst[idx[e]] = src[e];
    return 0;
  }
  const int nt = cx_host_threads();
#pragma omp parallel for num_threads(nt) schedule(static)
  for (long long e = 0; e < n; e++) dst[idx[e]] = src[e];
  return 0;
}

// receptor indices in the plan's own (regrouped, sorted) entry order: what a receptor rank needs to scatter a
// dense block written with mode 2
int cx_plan_sorted_rcv(cx_plan* P, int* out)
{
  if (P->n == 0) return 0;
  CX_CUDA(cudaMemcpy(out, P->d_rcv, sizeof(int) * (size_t)P->n, cudaMemcpyDeviceToHost));
  return 0;
}

// mark[node] = 1 for every donor node the plan reads (the stencil of each entry); mark: nptsD bytes on the device,
// zeroed by the caller (several plans of one donor zone accumulate into the same mask); asynchronous
int cx_plan_mark_donors(cx_plan* P, unsigned char* d_mark)
{
  cx_ctx* c = P->ctx;
  for (auto& G : P->groups)
  {
    if (G.n == 0) continue;
    int o = G.type == 1 ? 1 : (G.type == 2 || G.type == 22) ? 2 : G.type == 3 ? 3 : G.type == 44 ? 4 : 5;
    k_mark<<<cx_grid(G.n, 256), 256, 0, c->stream>>>(G.n, o, G.type == 22 ? 1 : o, P->imd, P->imd * P->jmd, P->d_donor + G.first, d_mark);
    c->launches++;
  }
  return 0;
}

// host side of a sparse upload: dst[e] = src[idx[e]]
int cx_host_gather(double* dst, const int* idx, const double* src, long long n)
{
  if (n < (1 << 15))
  {
    for (long long e = 0; e < n; e++) dst[e] = src[idx[e]];
    return 0;
  }
  const int nt = cx_host_threads();
#pragma omp parallel for num_threads(nt) schedule(static)
  for (long long e = 0; e < n; e++) dst[e] = src[idx[e]];
  return 0;
}

// ---- host trees in, host trees out: one iteration of a step with its copies --------------------------------
// What Connector/Mpi.py:396-440 (__setInterpTransfers + C._setPartialFields) does per call with host arrays: the
// donor values a step reads go up, the receptor values it wrote come back.  An item is one zone's fields in one
// direction: whole arrays, or (idx != NULL) only the listed points, packed through a page-locked block (host
// gather / device scatter on the way up, device gather / host scatter on the way down).  The run is pipelined
// per variable over three streams: upload of variable v+1 | plans + exchange of v | download of v-1, so both
// directions of the bus are busy at once.
struct cx_hostio {
  cx_ctx* ctx; int nvars;
  cudaStream_t up, dn;
  std::vector<cudaEvent_t> evU, evS, evD;
  struct Item {
    int dir; long long n; bool sparse;
    bool boxed = false; int dims[2] = {0, 0}; int box[6] = {0, 0, 0, 0, 0, 0};   // whole-array item cut to a box
    std::vector<int> idx; int* d_idx; double* hbuf; double* dbuf;
    std::vector<double*> h; std::vector<double*> d;
  };
  std::vector<Item> items;
  long long h2d_bytes, d2h_bytes;
};

int cx_hostio_create(cx_ctx* ctx, int nvars, cx_hostio** out)
{
  CX_CUDA(cudaSetDevice(ctx->device));
  if (nvars < 1 || nvars > 16) { cx_set_error("cx_hostio_create: 1..16 variables"); return -1; }
  cx_hostio* io = new cx_hostio;
  io->ctx = ctx; io->nvars = nvars; io->h2d_bytes = io->d2h_bytes = 0;
  CX_CUDA(cudaStreamCreateWithFlags(&io->up, cudaStreamNonBlocking));
  CX_CUDA(cudaStreamCreateWithFlags(&io->dn, cudaStreamNonBlocking));
  io->evU.resize(nvars); io->evS.resize(nvars); io->evD.resize(nvars);
  for (int v = 0; v < nvars; v++)
  {
    CX_CUDA(cudaEventCreateWithFlags(&io->evU[v], cudaEventDisableTiming));
    CX_CUDA(cudaEventCreateWithFlags(&io->evS[v], cudaEventDisableTiming));
    CX_CUDA(cudaEventCreateWithFlags(&io->evD[v], cudaEventDisableTiming));
  }
  *out = io;
  return 0;
}

// dir 0: host -> device before the step; 1: device -> host after it.  n: points moved per variable (the length of
// h_idx, or of the whole arrays when h_idx == NULL); h_field / d_field: nvars pointers each.
int cx_hostio_add(cx_hostio* io, int dir, long long n, const int* h_idx, double* const* h_field, double* const* d_field)
{
  if (n <= 0) return 0;
  cx_ctx* c = io->ctx;
  CX_CUDA(cudaSetDevice(c->device));
  io->items.emplace_back();
  cx_hostio::Item& I = io->items.back();
  I.dir = dir; I.n = n; I.sparse = h_idx != nullptr; I.d_idx = nullptr; I.hbuf = nullptr; I.dbuf = nullptr;
  I.h.assign(h_field, h_field + io->nvars); I.d.assign(d_field, d_field + io->nvars);
  if (I.sparse)
  {
    I.idx.assign(h_idx, h_idx + n);
    CX_CUDA(cudaMalloc(&I.d_idx, sizeof(int) * (size_t)n));
    CX_CHECK(cx_upload(c, I.d_idx, I.idx.data(), sizeof(int) * (size_t)n));
    CX_CUDA(cudaStreamSynchronize(c->stream));
    CX_CUDA(cudaMalloc(&I.dbuf, sizeof(double) * (size_t)n * io->nvars));
    CX_CUDA(cudaHostAlloc(&I.hbuf, sizeof(double) * (size_t)n * io->nvars, cudaHostAllocDefault));
  }
  (dir == 0 ? io->h2d_bytes : io->d2h_bytes) += 8ll * n * io->nvars;
  return 0;
}

// the same for the box [i0,i1] x [j0,j1] x [k0,k1] (inclusive) of (imd, jmd, kmd) arrays: one strided copy per variable
int cx_hostio_add_box(cx_hostio* io, int dir, int imd, int jmd, int kmd, const int* box, double* const* h_field,
                      double* const* d_field)
{
  if (box[0] < 0 || box[1] >= imd || box[2] < 0 || box[3] >= jmd || box[4] < 0 || box[5] >= kmd || box[0] > box[1] ||
      box[2] > box[3] || box[4] > box[5]) { cx_set_error("cx_hostio_add_box: box outside the array"); return -1; }
  io->items.emplace_back();
  cx_hostio::Item& I = io->items.back();
  I.dir = dir; I.sparse = false; I.boxed = true; I.d_idx = nullptr; I.hbuf = nullptr; I.dbuf = nullptr;
  I.dims[0] = imd; I.dims[1] = jmd;
  for (int q = 0; q < 6; q++) I.box[q] = box[q];
  I.n = (long long)(box[1] - box[0] + 1) * (box[3] - box[2] + 1) * (box[5] - box[4] + 1);
  I.h.assign(h_field, h_field + io->nvars); I.d.assign(d_field, d_field + io->nvars);
  (dir == 0 ? io->h2d_bytes : io->d2h_bytes) += 8ll * I.n * io->nvars;
  return 0;
}

static int hostio_box_copy(const cx_hostio::Item& I, int v, cudaStream_t s)
{
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  const size_t pitch = sizeof(double) * (size_t)I.dims[0];
  cudaPitchedPtr hp = make_cudaPitchedPtr((void*)I.h[v], pitch, pitch, (size_t)I.dims[1]);
  cudaPitchedPtr dp = make_cudaPitchedPtr((void*)I.d[v], pitch, pitch, (size_t)I.dims[1]);
  p.srcPtr = I.dir == 0 ? hp : dp; p.dstPtr = I.dir == 0 ? dp : hp;
  p.srcPos = make_cudaPos(sizeof(double) * (size_t)I.box[0], (size_t)I.box[2], (size_t)I.box[4]);
  p.dstPos = p.srcPos;
  p.extent = make_cudaExtent(sizeof(double) * (size_t)(I.box[1] - I.box[0] + 1), (size_t)(I.box[3] - I.box[2] + 1),
                             (size_t)(I.box[5] - I.box[4] + 1));
  p.kind = I.dir == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  CX_CUDA(cudaMemcpy3DAsync(&p, s));
  return 0;
}

int cx_hostio_bytes(cx_hostio* io, long long* h2d, long long* d2h)
{
  if (h2d) *h2d = io->h2d_bytes;
  if (d2h) *d2h = io->d2h_bytes;
  return 0;
}

extern int cx_step_run_vars(cx_step* S, int niter, int overlap, int v0, int nv);
extern int cx_step_splittable(cx_step* S);

int cx_step_run_host(cx_step* S, cx_hostio* io, int pipelined)
{
  cx_ctx* c = io->ctx;
  CX_CUDA(cudaSetDevice(c->device));
  const int nv = io->nvars;
  const int chunk = (pipelined && cx_step_splittable(S)) ? 1 : nv;
  // the copies of this call must follow whatever is queued on the main stream (e.g. the previous step)
  CX_CUDA(cudaEventRecord(c->ev_fork, c->stream));
  CX_CUDA(cudaStreamWaitEvent(io->up, c->ev_fork, 0));
  for (int v0 = 0; v0 < nv; v0 += chunk)
  {
    for (auto& I : io->items)
    {
      if (I.dir != 0) continue;
      if (I.sparse)
      {
        for (int v = v0; v < v0 + chunk; v++) cx_host_gather(I.hbuf + (size_t)v * I.n, I.idx.data(), I.h[v], I.n);
        CX_CUDA(cudaMemcpyAsync(I.dbuf + (size_t)v0 * I.n, I.hbuf + (size_t)v0 * I.n, sizeof(double) * (size_t)I.n * chunk,
                                cudaMemcpyHostToDevice, io->up));
        RPtrs R; for (int v = 0; v < chunk; v++) R.r[v] = I.d[v0 + v];
        k_scatter<<<cx_grid(I.n, 256), 256, 0, io->up>>>((int)I.n, chunk, I.dbuf + (size_t)v0 * I.n, I.n, I.d_idx, R);
        c->launches++;
      }
      else if (I.boxed)
        for (int v = v0; v < v0 + chunk; v++) CX_CHECK(hostio_box_copy(I, v, io->up));
      else
        for (int v = v0; v < v0 + chunk; v++)
          CX_CUDA(cudaMemcpyAsync(I.d[v], I.h[v], sizeof(double) * (size_t)I.n, cudaMemcpyHostToDevice, io->up));
    }
    CX_CUDA(cudaEventRecord(io->evU[v0], io->up));
    CX_CUDA(cudaStreamWaitEvent(c->stream, io->evU[v0], 0));
    CX_CHECK(cx_step_run_vars(S, 1, 1, v0, chunk));
    CX_CUDA(cudaEventRecord(io->evS[v0], c->stream));
    CX_CUDA(cudaStreamWaitEvent(io->dn, io->evS[v0], 0));
    for (auto& I : io->items)
    {
      if (I.dir != 1) continue;
      if (I.sparse)
      {
        RPtrs R; for (int v = 0; v < chunk; v++) R.r[v] = I.d[v0 + v];
        k_gather<<<cx_grid(I.n, 256), 256, 0, io->dn>>>((int)I.n, chunk, I.dbuf + (size_t)v0 * I.n, I.n, I.d_idx, R);
        c->launches++;
        CX_CUDA(cudaMemcpyAsync(I.hbuf + (size_t)v0 * I.n, I.dbuf + (size_t)v0 * I.n, sizeof(double) * (size_t)I.n * chunk,
                                cudaMemcpyDeviceToHost, io->dn));
      }
      else if (I.boxed)
        for (int v = v0; v < v0 + chunk; v++) CX_CHECK(hostio_box_copy(I, v, io->dn));
      else
        for (int v = v0; v < v0 + chunk; v++)
          CX_CUDA(cudaMemcpyAsync(I.h[v], I.d[v], sizeof(double) * (size_t)I.n, cudaMemcpyDeviceToHost, io->dn));
    }
    CX_CUDA(cudaEventRecord(io->evD[v0], io->dn));
  }
  for (int v0 = 0; v0 < nv; v0 += chunk)
  {
    CX_CUDA(cudaEventSynchronize(io->evD[v0]));
    for (auto& I : io->items)
      if (I.dir == 1 && I.sparse)
        for (int v = v0; v < v0 + chunk; v++) cx_host_scatter(I.h[v], I.idx.data(), I.hbuf + (size_t)v * I.n, I.n);
  }
  CX_CUDA(cudaStreamSynchronize(c->stream));
  CX_CUDA(cudaGetLastError());
  return 0;
}

int cx_hostio_destroy(cx_hostio* io)
{
  if (!io) return 0;
  cudaStreamSynchronize(io->up); cudaStreamSynchronize(io->dn);
  for (auto& I : io->items) { if (I.d_idx) cudaFree(I.d_idx); if (I.dbuf) cudaFree(I.dbuf); if (I.hbuf) cudaFreeHost(I.hbuf); }
  for (auto e : io->evU) cudaEventDestroy(e);
  for (auto e : io->evS) cudaEventDestroy(e);
  for (auto e : io->evD) cudaEventDestroy(e);
  cudaStreamDestroy(io->up); cudaStreamDestroy(io->dn);
  delete io;
  return 0;
}

int cx_plan_info(cx_plan* P, int* n, int* ngroups, long long* alg_fixed, long long* distinct)
{
  if (n) *n = P->n;
  if (ngroups) *ngroups = P->ngroups;
  if (alg_fixed) *alg_fixed = P->alg_bytes_fixed;
  if (distinct)
  {
    if (P->distinct_nodes < 0 && P->n > 0)
    {
      cx_ctx* c = P->ctx;
      unsigned char* mark; unsigned long long* cnt;
      CX_CUDA(cudaMalloc(&mark, (size_t)P->nptsD));
      CX_CUDA(cudaMalloc(&cnt, sizeof(unsigned long long)));
      CX_CUDA(cudaMemsetAsync(mark, 0, (size_t)P->nptsD, c->stream));
      CX_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), c->stream));
      CX_CHECK(cx_plan_mark_donors(P, mark));
      k_count<<<cx_grid(P->nptsD, 256), 256, 0, c->stream>>>(P->nptsD, mark, cnt);
      c->launches++;
      unsigned long long h = 0;
      CX_CUDA(cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
      CX_CUDA(cudaStreamSynchronize(c->stream));
      cudaFree(mark); cudaFree(cnt);
      P->distinct_nodes = (long long)h;
    }
    *distinct = P->n > 0 ? P->distinct_nodes : 0;
  }
  return 0;
}

// donor points per variable the host-pointer transfers of this plan move to the device: the whole zone, or the
// points the plan reads when the sparse upload applies (decided here if no transfer has run yet)
int cx_plan_upload_points(cx_plan* P, long long* npts)
{
  cx_plan_full* F = (cx_plan_full*)P;
  if (F->didx_state == 0 && P->n > 0) CX_CHECK(plan_donor_list(F));
  *npts = F->didx_state == 1 ? (long long)F->h_didx.size()
        : F->didx_state == 3 ? (long long)(F->box[1] - F->box[0] + 1) * (F->box[3] - F->box[2] + 1) * (F->box[5] - F->box[4] + 1)
        : (long long)P->nptsD;
  return 0;
}

int cx_plan_destroy(cx_plan* P)
{
  if (!P) return 0;
  cx_plan_full* F = (cx_plan_full*)P;
  for (auto& G : P->groups) { cudaFree(G.d_cf); if (G.d_tiles) cudaFree(G.d_tiles); if (G.d_pat) cudaFree(G.d_pat); }
  cudaFree(P->d_donor); cudaFree(P->d_rcv); cudaFree(P->d_pos); cudaFree(P->d_seq);
  if (F->h_dense) cudaFreeHost(F->h_dense);
  if (F->h_stage) cudaFreeHost(F->h_stage);
  if (F->d_didx) cudaFree(F->d_didx);
  for (cudaEvent_t e : F->evk) cudaEventDestroy(e);
  for (cudaEvent_t e : F->evd) cudaEventDestroy(e);
  delete F;
  return 0;
}

}  // extern "C"
