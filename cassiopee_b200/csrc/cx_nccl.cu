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

}  // extern "C"
