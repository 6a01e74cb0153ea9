// Experiment: cost of first-touch device allocations on B200 (cudaMalloc vs stream-ordered pool growth).
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
  cudaFree(0);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 0);
  unsigned long long thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  std::vector<void*> keep;
  for (size_t mb : {8, 128, 1024})
  {
    double t0 = now(); void* p; cudaMalloc(&p, mb << 20); double t1 = now(); keep.push_back(p);
    printf("cudaMalloc %zu MB: %.3f ms\n", mb, (t1 - t0) * 1e3);
  }
  for (size_t mb : {8, 128, 1024, 4096})
  {
    double t0 = now(); void* p; cudaMallocAsync(&p, mb << 20, s); cudaStreamSynchronize(s); double t1 = now();
    keep.push_back(p);
    printf("cudaMallocAsync (pool growth) %zu MB: %.3f ms\n", mb, (t1 - t0) * 1e3);
  }
  // free the 4 GB block back to the pool, then allocate 128 MB pieces from the reserved pool
  cudaFreeAsync(keep.back(), s); keep.pop_back(); cudaStreamSynchronize(s);
  double t0 = now();
  for (int i = 0; i < 20; i++) { void* p; cudaMallocAsync(&p, 128 << 20, s); keep.push_back(p); }
  cudaStreamSynchronize(s);
  printf("20 x cudaMallocAsync 128 MB from a primed pool: %.3f ms total\n", (now() - t0) * 1e3);
  t0 = now();
  for (int i = 0; i < 20; i++) { void* p; cudaMalloc(&p, 128 << 20); keep.push_back(p); }
  printf("20 x cudaMalloc 128 MB: %.3f ms total\n", (now() - t0) * 1e3);
  return 0;
}
