// cx_p2p.cu — peer-memory transport of the cross-GPU donor->receptor values.
//
// With NCCL the step of a rank is  remote plans -> send buffers -> ncclSend/ncclRecv -> scatter  and its cost at
// C4 sizes is the latency of the grouped send/recv (7 peers at 8 GPUs), not the bytes.  Here the donor-side
// transfer kernel itself is the sender: the receive buffers of every rank are exported with CUDA IPC, mapped
// by the donor ranks, and the remote plans write their dense (nvars, n_pair) blocks straight into the
// receptor rank's buffer with coalesced NVLink stores (entries are written in the plan's sorted order; the
// receptor rank got the matching index list once at set-up).  Completion is a flag per rank pair:
//     donor rank   : remote plans (stores to peer memory) ; k_p2p_signal: fence.sys, flag[peer][me] = iter
//     receptor rank: k_p2p_wait spins until flag[me][peer] >= iter for every peer ; scatter
// Receive buffers are double-buffered by iteration parity.  A rank only writes buffer (i+2)%2 of a peer after it
// has itself seen that peer's flag i+1, which the peer raises after its scatter of iteration i on the same
// stream, so no acknowledgement message is needed (flags are exchanged both ways for every pair, also when one
// direction carries no data).
#include "cx_internal.h"
#include "../../include/cassiopee_b200.h"
#include <dlfcn.h>
#include <map>
#include <string>
#include <string.h>

namespace {
__global__ void k_p2p_signal(int npeers, int* const* __restrict__ peer_flags, int iter)
{
  int p = threadIdx.x;
  if (p >= npeers) return;
  __threadfence_system();                       // the remote plans' stores (previous kernel) are performed; order the flag after them
  *((volatile int*)peer_flags[p]) = iter;
}
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// A peer that never signals (crashed rank, mismatched iteration counts) must not hang the GPU: after
// CX_P2P_TIMEOUT_NS the kernel traps and the error surfaces at the next synchronisation.
#define CX_P2P_TIMEOUT_NS 30000000000ull
__global__ void k_p2p_wait(int npeers, const int* my_flags, int iter)
{
  int p = threadIdx.x;
  if (p < npeers)
  {
    const volatile int* f = my_flags + p;
    const unsigned long long t0 = global_ns();
    while (*f < iter)
    {
      __nanosleep(100);
      if (global_ns() - t0 > CX_P2P_TIMEOUT_NS) __trap();
    }
  }
  __threadfence_system();
}

// signal + wait in one launch (one launch less on the exchange's critical path): every peer's flag is raised
// before any of ours is awaited, so two ranks that run this kernel at the same time cannot wait on each other
__global__ void k_p2p_signal_wait(int npeers, int* const* __restrict__ peer_flags, const int* my_flags, int iter)
{
  int p = threadIdx.x;
  if (p < npeers)
  {
    __threadfence_system();
    *((volatile int*)peer_flags[p]) = iter;
    const volatile int* f = my_flags + p;
    const unsigned long long t0 = global_ns();
    while (*f < iter)
    {
      __nanosleep(100);
      if (global_ns() - t0 > CX_P2P_TIMEOUT_NS) __trap();
    }
  }
  __threadfence_system();
}

typedef int (*cuMemGetAddressRange_t)(unsigned long long*, size_t*, unsigned long long);
cuMemGetAddressRange_t g_range = nullptr;
int load_range()
{
  if (g_range) return 0;
  void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { cx_set_error("cannot dlopen libcuda.so.1: %s", dlerror()); return -5; }
  g_range = (cuMemGetAddressRange_t)dlsym(h, "cuMemGetAddressRange_v2");
  if (!g_range) { cx_set_error("cuMemGetAddressRange_v2 missing"); return -5; }
  return 0;
}
struct Opened { void* base; int refs; };
std::map<std::string, Opened> g_opened;          // IPC handle bytes -> mapped base (a handle can be opened once per process) + users
}  // namespace

extern "C" {

// handle72 = 64-byte cudaIpcMemHandle_t of the allocation that contains dptr + 8-byte offset of dptr inside it
int cx_ipc_export(void* dptr, void* handle72)
{
  CX_CHECK(load_range());
  unsigned long long base = 0; size_t size = 0;
  if (g_range(&base, &size, (unsigned long long)(uintptr_t)dptr) != 0) { cx_set_error("cuMemGetAddressRange failed"); return -2; }
  cudaIpcMemHandle_t h;
  CX_CUDA(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
  memcpy(handle72, &h, 64);
  unsigned long long off = (unsigned long long)(uintptr_t)dptr - base;
  memcpy((char*)handle72 + 64, &off, 8);
  return 0;
}

int cx_ipc_open(cx_ctx* ctx, const void* handle72, void** dptr)
{
  CX_CUDA(cudaSetDevice(ctx->device));
  std::string key((const char*)handle72, 64);
  void* base = nullptr;
  auto it = g_opened.find(key);
  if (it != g_opened.end()) { base = it->second.base; it->second.refs++; }
  else
  {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle72, 64);
    CX_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    g_opened[key] = Opened{base, 1};
  }
  unsigned long long off;
  memcpy(&off, (const char*)handle72 + 64, 8);
  *dptr = (char*)base + off;
  return 0;
}

// drop one use of an imported allocation; it is unmapped when its last user closes it (two sessions can hold
// descriptors that resolve to the same allocation)
int cx_ipc_close(const void* handle72)
{
  std::string key((const char*)handle72, 64);
  auto it = g_opened.find(key);
  if (it == g_opened.end()) return 0;
  if (--it->second.refs > 0) return 0;
  cudaIpcCloseMemHandle(it->second.base);
  g_opened.erase(it);
  return 0;
}

int cx_ipc_close_all(void)
{
  for (auto& kv : g_opened) cudaIpcCloseMemHandle(kv.second.base);
  g_opened.clear();
  return 0;
}

}  // extern "C"

// used by cx_step_run (cx_transfer.cu, declared there inside its extern "C" block)
extern "C" int cx_p2p_signal_on(cudaStream_t s, int npeers, int* const* d_peer_flags, int iter)
{
  if (npeers <= 0) return 0;
  k_p2p_signal<<<1, 32, 0, s>>>(npeers, d_peer_flags, iter);
  return 0;
}
extern "C" int cx_p2p_signal_wait_on(cudaStream_t s, int npeers, int* const* d_peer_flags, const int* d_my_flags, int iter)
{
  if (npeers <= 0) return 0;
  k_p2p_signal_wait<<<1, 32, 0, s>>>(npeers, d_peer_flags, d_my_flags, iter);
  return 0;
}
extern "C" int cx_p2p_wait_on(cudaStream_t s, int npeers, const int* d_my_flags, int iter)
{
  if (npeers <= 0) return 0;
  k_p2p_wait<<<1, 32, 0, s>>>(npeers, d_my_flags, iter);
  return 0;
}
