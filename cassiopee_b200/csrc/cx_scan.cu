// cx_scan.cu — exclusive prefix sum of 64-bit counters (three-pass, block
// tiles of 2048).  Used for the per-donor-zone compaction of the per-receptor
// table (ascending receptor order per zone, Connector/Connector/setInterpData.cpp:1183-1252)
// and for the ragged coefficient offsets (SIZECF, connector.h:27-37): one packed
// 64-bit add carries (entry count << 32 | coefficient count).
#include "cx_internal.h"

namespace {
constexpr int TB = 256;
constexpr int IPT = 8;
constexpr int TILE = TB * IPT;

__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long* sh)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  unsigned long long t = 0;
  if (w == 0)
  {
    t = (l < TB / 32) ? sh[l] : 0ull;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (l == 0) sh[0] = t;
  }
  __syncthreads();
  t = sh[0];
  __syncthreads();
  return t;
}

__global__ void k_tile_sums(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ sums, int n)
{
  __shared__ unsigned long long sh[TB / 32];
  long long base = (long long)blockIdx.x * TILE;
  unsigned long long v = 0;
  for (int i = 0; i < IPT; i++)
  {
    long long idx = base + i * TB + threadIdx.x;
    if (idx < n) v += in[idx];
  }
  unsigned long long t = block_sum(v, sh);
  if (threadIdx.x == 0) sums[blockIdx.x] = t;
}

// single block: exclusive scan of the tile sums, in place; total to sums[nb]
__global__ void k_scan_sums(unsigned long long* sums, int nb)
{
  __shared__ unsigned long long sh[TB];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += TB)
  {
    int i = base + threadIdx.x;
    unsigned long long v = (i < nb) ? sums[i] : 0ull;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < TB; o <<= 1)
    {
      unsigned long long t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0ull;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    unsigned long long incl = sh[threadIdx.x];
    if (i < nb) sums[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == TB - 1) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[nb] = carry;
}

__global__ void k_tile_scan(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out,
                            const unsigned long long* __restrict__ sums, int n)
{
  // thread owns IPT consecutive items (blocked arrangement)
  __shared__ unsigned long long sh[TB];
  long long base = (long long)blockIdx.x * TILE + (long long)threadIdx.x * IPT;
  unsigned long long v[IPT];
  unsigned long long tsum = 0;
  for (int i = 0; i < IPT; i++)
  {
    long long idx = base + i;
    v[i] = (idx < n) ? in[idx] : 0ull;
    tsum += v[i];
  }
  sh[threadIdx.x] = tsum;
  __syncthreads();
  for (int o = 1; o < TB; o <<= 1)
  {
    unsigned long long t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0ull;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  unsigned long long run = sums[blockIdx.x] + sh[threadIdx.x] - tsum;
  for (int i = 0; i < IPT; i++)
  {
    long long idx = base + i;
    if (idx < n) out[idx] = run;
    run += v[i];
  }
}
}  // namespace

int cx_exclusive_scan_u64(cx_ctx* ctx, const unsigned long long* d_in, unsigned long long* d_out, int n,
                          unsigned long long* h_total)
{
  if (n <= 0) { if (h_total) *h_total = 0; return 0; }
  int nb = (n + TILE - 1) / TILE;
  unsigned long long* d_sums;
  CX_MALLOC(d_sums, sizeof(unsigned long long) * (nb + 1), ctx->stream);
  k_tile_sums<<<nb, TB, 0, ctx->stream>>>(d_in, d_sums, n);
  k_scan_sums<<<1, TB, 0, ctx->stream>>>(d_sums, nb);
  k_tile_scan<<<nb, TB, 0, ctx->stream>>>(d_in, d_out, d_sums, n);
  ctx->launches += 3;
  if (h_total)
  {
    CX_CUDA(cudaMemcpyAsync(h_total, d_sums + nb, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  CX_CUDA(cudaGetLastError());
  cudaFreeAsync(d_sums, ctx->stream);
  return 0;
}
