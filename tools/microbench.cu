// Micro-benchmarks behind the design choices in DESIGN.md §4 (run on a B200 via gpurun; prints cycles
// per warp-instruction per SM-resident warp set and derived rows/cycle/SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// mode: 0 LDS.64+STS.64 RMW f64 add, 1 LDS.128+STS.128 RMW, 2 atomicAdd u32 smem, 3 atomicAdd u64 smem, 4 atomicAdd f64 smem,
//       5 match_any only, 6 match_any + rank loop + RMW 128, 7 smem byte election + RMW 128,
//       8 u32 atomic ticket + rank loop + RMW 128 (what gb_shared_kernel does), 9 loop overhead only,
//       10 ticket + 3x LDS.128 + STS.128 + STS.128 (the six-aggregate record update)
template <int MODE>
__global__ void smem_kernel(int iters, int groups, u64* out, long long* cycles) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // per-warp private table of `groups` 16-byte records
  ulonglong2* tab = reinterpret_cast<ulonglong2*>(sm) + (size_t)warp * groups;
  volatile uint8_t* own = reinterpret_cast<volatile uint8_t*>(sm) + (size_t)(blockDim.x >> 5) * groups * 16 + (size_t)warp * groups;
  for (int i = lane; i < groups; i += 32) tab[i] = make_ulonglong2(0, 0);
  __syncthreads();
  uint32_t s = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
  u64 acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    int g = (int)(((u64)rng(s) * (u64)groups) >> 32);
    double x = __hiloint2double(0x40000000 | (s & 0xFFFFF), s);
    if (MODE == 0) {
      double* p = reinterpret_cast<double*>(tab) + g;   // 8-byte records
      *p += x;
    } else if (MODE == 1) {
      ulonglong2 a = tab[g];
      a.x = (u64)__double_as_longlong(__longlong_as_double((long long)a.x) + x); a.y += 1;
      tab[g] = a;
    } else if (MODE == 2) {
      atomicAdd(reinterpret_cast<unsigned int*>(tab) + g, 1u);
    } else if (MODE == 3) {
      atomicAdd(reinterpret_cast<u64*>(tab) + g, 1ull);
    } else if (MODE == 4) {
      atomicAdd(reinterpret_cast<double*>(tab) + g, x);
    } else if (MODE == 5) {
      acc += __match_any_sync(0xFFFFFFFFu, g);
    } else if (MODE == 6) {
      unsigned peers = __match_any_sync(0xFFFFFFFFu, g);
      int rank = __popc(peers & ((1u << lane) - 1u));
      int maxr = __reduce_max_sync(0xFFFFFFFFu, rank);
      for (int r = 0; r <= maxr; r++) {
        if (rank == r) { ulonglong2 a = tab[g]; a.x = (u64)__double_as_longlong(__longlong_as_double((long long)a.x) + x); a.y += 1; tab[g] = a; }
        __syncwarp();
      }
    } else if (MODE == 8 || MODE == 10) {
      uint32_t* cnt = reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(own)) ;   // reuse the byte area (groups bytes >= 4*groups/4..): see smem size
      uint32_t old = atomicAdd(&cnt[g >> 2], 1u);     // 4 groups share a counter word here: only the cost matters
      __syncwarp();
      uint32_t now = *reinterpret_cast<volatile uint32_t*>(&cnt[g >> 2]);
      int rr = (int)(now - old - 1u);
      int maxr = __reduce_max_sync(0xFFFFFFFFu, rr);
      for (int r = 0; r <= maxr; r++) {
        if (rr == r) {
          ulonglong2 a = tab[g];
          a.x = (u64)__double_as_longlong(__longlong_as_double((long long)a.x) + x); a.y += 1;
          if (MODE == 10) {
            ulonglong2 b = tab[(g + 1) % groups], m = tab[(g + 2) % groups];
            b.x += 1; acc += m.x;
            tab[(g + 1) % groups] = b;
          }
          tab[g] = a;
        }
        if (r < maxr) __syncwarp();
      }
    } else if (MODE == 9) {
      acc += g + (u64)x;
    } else if (MODE == 7) {
      bool pending = true;
      while (__any_sync(0xFFFFFFFFu, pending)) {
        if (pending) own[g] = (uint8_t)lane;
        __syncwarp();
        if (pending && own[g] == (uint8_t)lane) { ulonglong2 a = tab[g]; a.x = (u64)__double_as_longlong(__longlong_as_double((long long)a.x) + x); a.y += 1; tab[g] = a; pending = false; }
        __syncwarp();
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x1234) out[0] = acc;
  if (lane == 0 && warp == 0) out[1 + blockIdx.x] = tab[0].x;
}

// global RED.ADD.F64 into `groups` addresses (stride 64 B)
__global__ void red_kernel(double* tab, int groups, int iters, long long* cycles) {
  uint32_t s = 99991u + (blockIdx.x * blockDim.x + threadIdx.x) * 7919u;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) { int g = rng(s) % groups; atomicAdd(&tab[(size_t)g * 8], 1.0); }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run_smem(const char* name, int warps, int groups, int iters) {
  u64* out; long long* cyc;
  CHECK(cudaMalloc(&out, 8 * 1024)); CHECK(cudaMalloc(&cyc, 8 * 1024));
  size_t smem = (size_t)warps * groups * 17 + 64;
  CHECK(cudaFuncSetAttribute(smem_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  smem_kernel<MODE><<<148, warps * 32, smem>>>(iters, groups, out, cyc);
  CHECK(cudaDeviceSynchronize());
  long long h[148]; CHECK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
  double per_warp_instr = avg / iters;                 // cycles between two iterations of one warp
  double rows_per_cycle = (double)warps * 32 / per_warp_instr;
  printf("%-34s warps=%2d groups=%5d  %7.1f cyc/iter/warp  -> %6.2f rows/cycle/SM  (%.3f cyc/row)\n", name, warps, groups, per_warp_instr, rows_per_cycle, 1.0 / rows_per_cycle);
  cudaFree(out); cudaFree(cyc);
  return 0;
}

int main() {
  const int iters = 20000;
  for (int groups : {8, 1024}) {
    for (int warps : {4, 8, 16}) {
      if ((size_t)warps * groups * 17 > 220000) continue;
      run_smem<0>("LDS.64/STS.64 RMW", warps, groups, iters);
      run_smem<1>("LDS.128/STS.128 RMW", warps, groups, iters);
      run_smem<2>("ATOMS.ADD u32", warps, groups, iters);
      run_smem<3>("ATOMS.ADD u64", warps, groups, iters);
      run_smem<4>("atomicAdd f64 (smem)", warps, groups, iters);
      run_smem<5>("MATCH.ANY only", warps, groups, iters);
      run_smem<6>("match_any + rank rounds + RMW128", warps, groups, iters);
      run_smem<7>("smem byte election + RMW128", warps, groups, iters);
      run_smem<8>("u32 ticket + rank rounds + RMW128", warps, groups, iters);
      run_smem<10>("u32 ticket + 3xLDS128 + 2xSTS128", warps, groups, iters);
      run_smem<9>("loop overhead only", warps, groups, iters);
    }
  }
  for (int groups : {1000, 100000, 10000000}) {
    double* tab; long long* cyc;
    CHECK(cudaMalloc(&tab, (size_t)groups * 64)); CHECK(cudaMemset(tab, 0, (size_t)groups * 64)); CHECK(cudaMalloc(&cyc, 8 * 4096));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int it = 2000;
    red_kernel<<<148 * 8, 256>>>(tab, groups, 10, cyc);
    cudaEventRecord(a);
    red_kernel<<<148 * 8, 256>>>(tab, groups, it, cyc);
    cudaEventRecord(b);
    CHECK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)148 * 8 * 256 * it;
    printf("RED.ADD.F64 global, %8d addresses (64 B apart): %.2f G updates/s\n", groups, ops / ms / 1e6);
    cudaFree(tab); cudaFree(cyc);
  }
  return 0;
}
