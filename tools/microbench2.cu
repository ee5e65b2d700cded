// Marginal throughput of shared-memory operations with random addresses (CTA-shared table), B200.
// Each loop iteration issues K independent operations; the index arithmetic is measured separately (op 0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define K 8

template <int OP>
__global__ void tput_kernel(int iters, int log_groups, u64* out, long long* cycles) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int groups = 1 << log_groups;
  for (int i = threadIdx.x; i < groups * 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  __syncthreads();
  uint32_t s = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
  u64 acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      s = s * 1664525u + 1013904223u;
      const uint32_t g = (s * 2654435761u) >> (32 - log_groups);
      if (OP == 0) acc += g;
      if (OP == 1) atomicAdd(reinterpret_cast<uint32_t*>(sm) + g, 1u);
      if (OP == 2) acc += atomicAdd(reinterpret_cast<uint32_t*>(sm) + g, 1u);
      if (OP == 3) acc += reinterpret_cast<volatile uint32_t*>(sm)[g];
      if (OP == 4) acc += reinterpret_cast<volatile u64*>(sm)[g];
      if (OP == 5) { ulonglong2 v; asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(sm) + g * 16)); acc += v.x + v.y; }
      if (OP == 6) reinterpret_cast<volatile u64*>(sm)[g] = s;
      if (OP == 7) { asm volatile("st.shared.v2.u64 [%0], {%1,%2};" :: "r"((uint32_t)__cvta_generic_to_shared(sm) + g * 16), "l"((u64)s), "l"((u64)k)); }
      if (OP == 8) atomicMax(reinterpret_cast<uint32_t*>(sm) + g, s);
      if (OP == 9) atomicAdd(reinterpret_cast<u64*>(sm) + g, 1ull);
      if (OP == 10) acc += atomicExch(reinterpret_cast<uint32_t*>(sm) + g, s);
      if (OP == 11) atomicAdd(reinterpret_cast<float*>(sm) + g, 1.0f);
      if (OP == 12) acc += __match_any_sync(0xFFFFFFFFu, g & 31);
      if (OP == 13) acc += __reduce_max_sync(0xFFFFFFFFu, g);
      if (OP == 14) acc += __ballot_sync(0xFFFFFFFFu, g & 1);
      if (OP == 15) { __syncwarp(); acc += g; }
      if (OP == 16) acc += __shfl_xor_sync(0xFFFFFFFFu, g, 1);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x1234567) out[0] = acc;
}

template <int OP>
void run(const char* name, int warps, int log_groups, int iters) {
  u64* out; long long* cyc;
  cudaMalloc(&out, 64); cudaMalloc(&cyc, 8 * 256);
  size_t smem = (size_t)(1 << log_groups) * 16;
  cudaFuncSetAttribute(tput_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tput_kernel<OP><<<148, warps * 32, smem>>>(iters, log_groups, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
  double per_instr_sm = avg / ((double)iters * K * warps);     // SM cycles per warp-instruction
  printf("%-28s warps=%2d groups=%5d  %6.2f SM-cycles per warp-op  (%.3f cyc/row)\n", name, warps, 1 << log_groups, per_instr_sm, per_instr_sm / 32);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int iters = 4000;
  for (int lg : {3, 10, 13}) {
    for (int warps : {8, 32}) {
      run<0>("index arithmetic only", warps, lg, iters);
      run<1>("RED.shared.add.u32", warps, lg, iters);
      run<2>("ATOMS.ADD.u32 (returns)", warps, lg, iters);
      run<8>("RED.shared.max.u32", warps, lg, iters);
      run<10>("ATOMS.EXCH.u32 (returns)", warps, lg, iters);
      run<11>("atomicAdd f32 (smem)", warps, lg, iters);
      run<9>("atomicAdd u64 (smem)", warps, lg, iters);
      run<3>("LDS.32", warps, lg, iters);
      run<4>("LDS.64", warps, lg, iters);
      run<5>("LDS.128", warps, lg, iters);
      run<6>("STS.64", warps, lg, iters);
      run<7>("STS.128", warps, lg, iters);
      if (lg == 10) {
        run<12>("MATCH.ANY (<=32 distinct)", warps, lg, iters);
        run<13>("REDUX.MAX", warps, lg, iters);
        run<14>("VOTE.BALLOT", warps, lg, iters);
        run<15>("WARPSYNC", warps, lg, iters);
        run<16>("SHFL", warps, lg, iters);
      }
    }
  }
  return 0;
}
