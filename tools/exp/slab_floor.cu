// slab_floor.cu — experiment: how fast can the B200 move the memory footprint of one C4 transfer step at all?
// The step reads, per donor zone (99^3 doubles per variable, 7 variables), three 4-node-thick slabs normal to x, y, z
// and writes three 2-cell-thick receptor slabs, plus a streaming read of the plan (indices + coefficients).
// Each component is timed alone with an "ideal" access kernel (one thread per 8-byte element, grid-wide, no
// arithmetic, no indices): x-normal slabs touch 1-2 sectors of every 792-byte row, y/z-normal slabs are contiguous.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/slab_floor tools/exp/slab_floor.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int N = 99, NF = 56;   // nodes per direction, fields (8 zones x 7 vars)

// slab of thickness T starting at S along axis AX (0: x, 1: y, 2: z): element e of the slab -> node index
template <int AX> __device__ __forceinline__ int slab_node(int e, int S, int T)
{
  if (AX == 0) { int t = e % T, r = e / T; return (S + t) + N * r; }                 // r = j + N*k
  if (AX == 1) { int i = e % N, r = e / N, t = r % T, k = r / T; return i + N * (S + t) + N * N * k; }
  int i = e % N, r = e / N, j = r % N, t = r / N; return i + N * j + N * N * (S + t);
}
template <int AX> __global__ void k_read(double* const* f, int S, int T, double* sink)
{
  const int per = T * N * N;
  double acc = 0.;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < (long long)per * NF; g += (long long)gridDim.x * blockDim.x)
  {
    int fi = (int)(g / per), e = (int)(g - (long long)fi * per);
    acc += __ldg(f[fi] + slab_node<AX>(e, S, T));
  }
  if (acc == 1.2345e-300) sink[0] = acc;
}
template <int AX> __global__ void k_write(double* const* f, int S, int T)
{
  const int per = T * N * N;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < (long long)per * NF; g += (long long)gridDim.x * blockDim.x)
  {
    int fi = (int)(g / per), e = (int)(g - (long long)fi * per);
    f[fi][slab_node<AX>(e, S, T)] = 1.0;
  }
}
__global__ void k_stream(const double* p, size_t n, double* sink)
{
  double acc = 0.;
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n; g += (size_t)gridDim.x * blockDim.x) acc += __ldg(p + g);
  if (acc == 1.2345e-300) sink[0] = acc;
}
__global__ void k_flush(double* p, size_t n) { for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n; g += (size_t)gridDim.x * blockDim.x) p[g] = 0.; }

int main()
{
  std::vector<double*> h(NF);
  for (int i = 0; i < NF; i++) { CK(cudaMalloc(&h[i], sizeof(double) * N * N * N)); CK(cudaMemset(h[i], 0, sizeof(double) * N * N * N)); }
  double** d; CK(cudaMalloc(&d, sizeof(double*) * NF)); CK(cudaMemcpy(d, h.data(), sizeof(double*) * NF, cudaMemcpyHostToDevice));
  double* sink; CK(cudaMalloc(&sink, 8));
  const size_t nplan = 38000000 / 8; double* plan; CK(cudaMalloc(&plan, nplan * 8)); CK(cudaMemset(plan, 0, nplan * 8));
  const size_t nfl = 256u << 20 >> 3; double* fl; CK(cudaMalloc(&fl, nfl * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int G = 148 * 8, B = 256;
  auto timeit = [&](const char* name, double bytes, auto launch) {
    float best = 1e9f, cold = 0.f;
    for (int it = 0; it < 6; it++)
    {
      k_flush<<<G, B>>>(fl, nfl);                 // evict L2 (256 MB written)
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it == 0) cold = ms; if (ms < best) best = ms;
    }
    printf("%-22s useful %7.2f MB  best %7.2f us  -> %7.1f GB/s useful (first %.2f us)\n", name, bytes / 1e6, best * 1e3, bytes / best / 1e6, cold * 1e3);
  };
  const double slabR = 8.0 * 4 * N * N * NF / 3, slabW = 8.0 * 2 * N * N * NF / 3;   // a third of the fields' slabs per direction... (all 56 fields are touched here: scale below)
  (void)slabR; (void)slabW;
  // every field has one slab per direction in the real step (each zone has 3 face neighbours); time all 56 fields per direction
  timeit("read  x-slab (4 thick)", 8.0 * 4 * N * N * NF, [&] { k_read<0><<<G, B>>>(d, 91, 4, sink); });
  timeit("read  x-slab (3 thick)", 8.0 * 3 * N * N * NF, [&] { k_read<0><<<G, B>>>(d, 91, 3, sink); });
  timeit("read  y-slab (4 thick)", 8.0 * 4 * N * N * NF, [&] { k_read<1><<<G, B>>>(d, 91, 4, sink); });
  timeit("read  z-slab (4 thick)", 8.0 * 4 * N * N * NF, [&] { k_read<2><<<G, B>>>(d, 91, 4, sink); });
  timeit("write x-slab (2 thick)", 8.0 * 2 * N * N * NF, [&] { k_write<0><<<G, B>>>(d, 0, 2); });
  timeit("write y-slab (2 thick)", 8.0 * 2 * N * N * NF, [&] { k_write<1><<<G, B>>>(d, 0, 2); });
  timeit("write z-slab (2 thick)", 8.0 * 2 * N * N * NF, [&] { k_write<2><<<G, B>>>(d, 0, 2); });
  timeit("stream plan (38 MB)", 8.0 * nplan, [&] { k_stream<<<G, B>>>(plan, nplan, sink); });
  timeit("all of the above", 8.0 * (4 * 3 + 2 * 3) * N * N * NF + 8.0 * nplan, [&] {
    k_read<0><<<G, B>>>(d, 91, 4, sink); k_read<1><<<G, B>>>(d, 91, 4, sink); k_read<2><<<G, B>>>(d, 91, 4, sink);
    k_write<0><<<G, B>>>(d, 0, 2); k_write<1><<<G, B>>>(d, 0, 2); k_write<2><<<G, B>>>(d, 0, 2); k_stream<<<G, B>>>(plan, nplan, sink); });
  return 0;
}
