// Peer-store microbenchmark (one process, 2 GPUs, peer access): what do scattered runs of L consecutive elements
// cost when the destination is the OTHER GPU's memory (NVLink stores) compared with local HBM?  This is the access
// pattern of the fused partition + shuffle (join.cu jpart1_kernel<true>): a tile of staged rows is written out as
// one run per combined bucket, run length = tile rows / (ranks x radix buckets).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/peer_store_bench.cu -o tools/peer_store_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 31; x *= 0x9E3779B97F4A7C15ull; x ^= x >> 29; return x; }

// every thread stores `ESZ` bytes; consecutive threads fill a run of L elements; run r lands at slot perm(r) of the output
template <typename T>
__global__ void scatter_runs(T* __restrict__ dst, long long nelem, int logL, long long nruns_mask) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nelem; i += stride) {
    const long long run = i >> logL, within = i & ((1ll << logL) - 1);
    const long long slot = (long long)(mix((uint64_t)run) & (uint64_t)nruns_mask);
    T v; memset(&v, 0, sizeof(T)); *reinterpret_cast<uint32_t*>(&v) = (uint32_t)i;
    dst[(slot << logL) + within] = v;
  }
}

template <typename T>
static int run(const char* name, T* dst, long long bytes, int logL, cudaStream_t s) {
  const long long nelem = bytes / sizeof(T);
  const long long nruns = nelem >> logL;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(a, s);
    scatter_runs<T><<<148 * 8, 512, 0, s>>>(dst, nelem, logL, nruns - 1);
    cudaEventRecord(b, s);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep && ms < best) best = ms;
  }
  printf("%-6s elem %2zu B  run %5lld B : %8.1f GB/s\n", name, sizeof(T), (long long)sizeof(T) << logL, bytes / (best * 1e-3) / 1e9);
  return 0;
}

int main() {
  int n = 0; CK(cudaGetDeviceCount(&n));
  const long long bytes = 4ll << 30;       // power of two
  void *local = nullptr, *peer = nullptr;
  CK(cudaSetDevice(0)); CK(cudaMalloc(&local, bytes));
  if (n > 1) { CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, bytes)); CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); }
  cudaStream_t s; CK(cudaStreamCreate(&s));
  for (int pass = 0; pass < (n > 1 ? 2 : 1); pass++) {
    void* d = pass ? peer : local;
    const char* nm = pass ? "peer" : "local";
    for (int logL : {2, 3, 4, 5, 7, 10}) if (run<uint64_t>(nm, (uint64_t*)d, bytes, logL, s)) return 1;
    for (int logL : {2, 3, 4, 5, 7, 10}) if (run<uint32_t>(nm, (uint32_t*)d, bytes / 2, logL, s)) return 1;
    for (int logL : {1, 2, 3, 4, 6, 9}) if (run<ulonglong2>(nm, (ulonglong2*)d, bytes, logL, s)) return 1;
  }
  if (n > 1) {   // plain peer copy for reference
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(a, s); cudaMemcpyPeerAsync(peer, 1, local, 0, bytes, s); cudaEventRecord(b, s); cudaEventSynchronize(b); }
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("cudaMemcpyPeerAsync 4 GiB: %.1f GB/s\n", bytes / (ms * 1e-3) / 1e9);
  }
  return 0;
}
