// cx_nccl.cu — cross-GPU donor->receptor value exchange.
//
// The reference ships interpolated values between ranks as pickled numpy
// objects with mpi4py isend/recv following a {src:{dst:[zones]}} graph
// (Converter/Converter/Mpi4py.py:149-168, used by Connector/Mpi.py:396-440).
// Here each rank pair exchanges ONE packed device buffer per iteration
// (nvars x sum of nrcv of the pair's plans, written directly by the donor-side
// transferD kernels) with grouped ncclSend/ncclRecv over NVLink; index lists
// are exchanged once at plan creation by the Python layer.
// NCCL is dlopen'ed so single-GPU use has no NCCL dependency.
#include "cx_internal.h"
#include "../../include/cassiopee_b200.h"
#include <dlfcn.h>
#include <string.h>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };   // ncclDataType_t: ncclDouble = 8

struct NcclApi {
  void* h;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl = {nullptr};

static int load_nccl()
{
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (int i = 0; i < 2 && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { cx_set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return -5; }
#define SYM(f, n) *(void**)(&g_nccl.f) = dlsym(h, n); if (!g_nccl.f) { cx_set_error("NCCL symbol %s missing", n); return -5; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
  SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.h = h;
  return 0;
}

struct cx_comm { ncclComm_t comm; int rank, nranks; double* d_tmp; };

#define CX_NCCL(call)                                                             \
  do {                                                                            \
    ncclResult_t r_ = (call);                                                     \
    if (r_ != 0) { cx_set_error("%s:%d NCCL error %s", __FILE__, __LINE__, g_nccl.GetErrorString(r_)); return -5; } \
  } while (0)

extern "C" {

int cx_comm_unique_id(char* id128)
{
  CX_CHECK(load_nccl());
  ncclUniqueId id;
  CX_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return 0;
}

int cx_comm_init(cx_ctx* ctx, int rank, int nranks, const char* id128)
{
  CX_CHECK(load_nccl());
  CX_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  cx_comm* c = new cx_comm;
  c->rank = rank; c->nranks = nranks; c->d_tmp = nullptr;
  CX_NCCL(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
  CX_CUDA(cudaMalloc(&c->d_tmp, sizeof(double) * 8));
  ctx->nccl = c;
  return 0;
}

int cx_comm_destroy(cx_ctx* ctx)
{
  cx_comm* c = (cx_comm*)ctx->nccl;
  if (!c) return 0;
  g_nccl.CommDestroy(c->comm);
  cudaFree(c->d_tmp);
  delete c;
  ctx->nccl = nullptr;
  return 0;
}

int cx_comm_barrier(cx_ctx* ctx)
{
  cx_comm* c = (cx_comm*)ctx->nccl;
  if (!c) { cx_set_error("communicator not initialised"); return -5; }
  CX_CUDA(cudaMemsetAsync(c->d_tmp, 0, sizeof(double) * 2, ctx->stream));
  CX_NCCL(g_nccl.AllReduce(c->d_tmp, c->d_tmp + 1, 1, ncclFloat64, 0 /*ncclSum*/, c->comm, ctx->stream));
  CX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int cx_exchange_on_stream(cx_ctx* ctx, cudaStream_t st, int npeers, const int* peers, const double* const* d_send,
                          const long long* sendcounts, double* const* d_recv, const long long* recvcounts)
{
  cx_comm* c = (cx_comm*)ctx->nccl;
  if (!c) { cx_set_error("communicator not initialised"); return -5; }
  CX_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < npeers; p++)
  {
    if (sendcounts[p] > 0) CX_NCCL(g_nccl.Send(d_send[p], (size_t)sendcounts[p], ncclFloat64, peers[p], c->comm, st));
    if (recvcounts[p] > 0) CX_NCCL(g_nccl.Recv(d_recv[p], (size_t)recvcounts[p], ncclFloat64, peers[p], c->comm, st));
  }
  CX_NCCL(g_nccl.GroupEnd());
  return 0;
}

int cx_exchange_dev(cx_ctx* ctx, int npeers, const int* peers, const double* const* d_send,
                    const long long* sendcounts, double* const* d_recv, const long long* recvcounts)
{
  return cx_exchange_on_stream(ctx, ctx->stream, npeers, peers, d_send, sendcounts, d_recv, recvcounts);
}

int cx_exchange_bytes_dev(cx_ctx* ctx, int npeers, const int* peers, const void* const* d_send,
                          const long long* sendbytes, void* const* d_recv, const long long* recvbytes)
{
  cx_comm* c = (cx_comm*)ctx->nccl;
  if (!c) { cx_set_error("communicator not initialised"); return -5; }
  CX_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < npeers; p++)
  {
    if (sendbytes[p] > 0) CX_NCCL(g_nccl.Send(d_send[p], (size_t)sendbytes[p], 0 /*ncclInt8*/, peers[p], c->comm, ctx->stream));
    if (recvbytes[p] > 0) CX_NCCL(g_nccl.Recv(d_recv[p], (size_t)recvbytes[p], 0 /*ncclInt8*/, peers[p], c->comm, ctx->stream));
  }
  CX_NCCL(g_nccl.GroupEnd());
  return 0;
}

// ---- device-to-device replication of donor zones ------------------------------------------------------------
// The device form of Cmpi._addXZones(aD, graph, variables=['cellN']) (Converter/Converter/Mpi.py, called by
// Connector/Connector/Mpi.py:941): a rank that needs a remote donor zone receives the owner's resident block
// (x, y, z, cellN) AND its ADT replica over NCCL/NVLink, so the zone is ready to be queried without a host
// round trip and without rebuilding the tree.  meta_i[6*i] = ni, nj, nk, depth, levels, topo; meta_d[9*i] =
// bbox[6], h[3] (from cx_zone_meta on the owner, exchanged by the caller).  Between one pair of ranks the
// zones must be listed in the same order on both sides.
int cx_zone_meta(cx_zone* z, int* meta_i, double* meta_d)
{
  meta_i[0] = z->ni; meta_i[1] = z->nj; meta_i[2] = z->nk; meta_i[3] = z->depth; meta_i[4] = z->levels; meta_i[5] = z->topo;
  for (int i = 0; i < 6; i++) meta_d[i] = z->bbox[i];
  for (int i = 0; i < 3; i++) meta_d[6 + i] = z->h[i];
  return 0;
}

int cx_zone_exchange(cx_ctx* ctx, int nsend, cx_zone* const* send_zones, const int* send_peers, int nrecv,
                     const int* recv_peers, const int* meta_i, const double* meta_d, cx_zone** out)
{
  cx_comm* c = (cx_comm*)ctx->nccl;
  if (!c) { cx_set_error("communicator not initialised"); return -5; }
  CX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  for (int i = 0; i < nrecv; i++) out[i] = nullptr;
  for (int i = 0; i < nrecv; i++)
  {
    cx_zone* zn = new cx_zone;
    memset(zn, 0, sizeof(*zn));
    const int* mi = meta_i + 6 * i; const double* md = meta_d + 9 * i;
    zn->ctx = ctx; zn->ni = mi[0]; zn->nj = mi[1]; zn->nk = mi[2]; zn->depth = mi[3]; zn->levels = mi[4]; zn->topo = mi[5];
    zn->npts = zn->ni * zn->nj * zn->nk;
    zn->ncells = (zn->ni - 1) * (zn->nj - 1) * (zn->nk > 1 ? zn->nk - 1 : 1);
    for (int q = 0; q < 6; q++) zn->bbox[q] = md[q];
    for (int q = 0; q < 3; q++) zn->h[q] = md[6 + q];
    out[i] = zn;
    // stream-ordered pool: the blocks of the previous call's replicas are reused instead of going back to the driver
    cudaError_t e = cx_malloc_async((void**)&zn->d_x, 4 * sizeof(double) * (size_t)zn->npts, st);
    if (e == cudaSuccess && zn->topo == 1)
      e = cx_malloc_async((void**)&zn->d_nodes, sizeof(cx::AdtNode) * (size_t)zn->ncells, st);
    if (e != cudaSuccess)
    {
      cx_set_error("cx_zone_exchange: %s", cudaGetErrorString(e));
      for (int q = 0; q <= i; q++) { cudaFree(out[q]->d_x); cudaFree(out[q]->d_nodes); delete out[q]; out[q] = nullptr; }
      return -2;
    }
    zn->d_y = zn->d_x + zn->npts; zn->d_z = zn->d_y + zn->npts; zn->d_cellN = zn->d_z + zn->npts;
  }
  auto transfer = [&]() -> int {
    CX_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < nsend; i++)
    {
      cx_zone* z = send_zones[i];
      CX_NCCL(g_nccl.Send(z->d_x, 4 * (size_t)z->npts, ncclFloat64, send_peers[i], c->comm, st));
      if (z->topo == 1)
        CX_NCCL(g_nccl.Send(z->d_nodes, sizeof(cx::AdtNode) * (size_t)z->ncells, 0 /*ncclInt8*/, send_peers[i], c->comm, st));
    }
    for (int i = 0; i < nrecv; i++)
    {
      cx_zone* z = out[i];
      CX_NCCL(g_nccl.Recv(z->d_x, 4 * (size_t)z->npts, ncclFloat64, recv_peers[i], c->comm, st));
      if (z->topo == 1)
        CX_NCCL(g_nccl.Recv(z->d_nodes, sizeof(cx::AdtNode) * (size_t)z->ncells, 0 /*ncclInt8*/, recv_peers[i], c->comm, st));
    }
    CX_NCCL(g_nccl.GroupEnd());
    CX_CUDA(cudaStreamSynchronize(st));
    CX_CUDA(cudaGetLastError());
    return 0;
  };
  const int rc = transfer();
  if (rc != 0)       // the caller never sees half-filled zones
    for (int q = 0; q < nrecv; q++) { cudaFree(out[q]->d_x); cudaFree(out[q]->d_nodes); delete out[q]; out[q] = nullptr; }
  return rc;
}

}  // extern "C"
