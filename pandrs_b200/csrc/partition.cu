// Hash partitioning of rows by key tuple: dest = mix(hash(key)) mod nparts, plus a permutation that
// groups the row ids by destination.  This is the local half of the multi-GPU shuffle (one process per
// GPU, NCCL all-to-all of the permuted columns); pandrs has no counterpart (no comms backend,
// SURVEY.md §5) — the only reference-side notion is PartitionStrategy::Hash
// (src/distributed/core/partition.rs:11-18).
#include <algorithm>

#include "groupby_kernels.cuh"

#define PART_THREADS 256
#define PART_ITEMS 8
#define PART_MAX 1024

struct PartParams { KeySpec ks; long long n; int nparts; u64* counts; u64* cursor; long long* perm; };

template <int NW>
__device__ __forceinline__ int part_dest(const KeySpec& ks, long long row, int nparts) {
  u64 w[NW];
  if (load_key_generic<NW>(ks, row, w)) return 0;   // the NULL-key group lives on rank 0
  return (int)(pdrs_mix64(key_hash<NW>(w) ^ 0x5851F42D4C957F2Dull) % (u64)nparts);
}

template <int NW>
__global__ void __launch_bounds__(PART_THREADS) part_count_kernel(const PartParams p) {
  __shared__ unsigned int hist[PART_MAX];
  for (int i = threadIdx.x; i < p.nparts; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < p.n; r += (long long)gridDim.x * blockDim.x)
    atomicAdd(&hist[part_dest<NW>(p.ks, r, p.nparts)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < p.nparts; i += blockDim.x) if (hist[i]) atomicAdd(&p.counts[i], (u64)hist[i]);
}

__global__ void part_scan_kernel(const u64* counts, u64* cursor, int nparts) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { u64 s = 0; for (int i = 0; i < nparts; i++) { cursor[i] = s; s += counts[i]; } }
}

template <int NW>
__global__ void __launch_bounds__(PART_THREADS) part_scatter_kernel(const PartParams p) {
  __shared__ unsigned int hist[PART_MAX];
  __shared__ u64 base[PART_MAX];
  const long long tile = (long long)PART_THREADS * PART_ITEMS;
  for (long long t0 = (long long)blockIdx.x * tile; t0 < p.n; t0 += (long long)gridDim.x * tile) {
    for (int i = threadIdx.x; i < p.nparts; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    int dest[PART_ITEMS];
    unsigned int off[PART_ITEMS];
#pragma unroll
    for (int j = 0; j < PART_ITEMS; j++) {
      long long r = t0 + (long long)j * PART_THREADS + threadIdx.x;
      dest[j] = -1;
      if (r < p.n) { dest[j] = part_dest<NW>(p.ks, r, p.nparts); off[j] = atomicAdd(&hist[dest[j]], 1u); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.nparts; i += blockDim.x) if (hist[i]) base[i] = atomicAdd(&p.cursor[i], (u64)hist[i]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PART_ITEMS; j++) {
      long long r = t0 + (long long)j * PART_THREADS + threadIdx.x;
      if (dest[j] >= 0) p.perm[base[dest[j]] + off[j]] = r;
    }
    __syncthreads();
  }
}

extern "C" int32_t pdrs_hash_partition(pdrs_ctx* c, const pdrs_col* keys, int32_t nkeys, int32_t nparts, int64_t* perm_dev, int64_t* counts_host) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!keys || nkeys < 1 || nkeys > PDRS_MAX_KEYS || nparts < 1 || nparts > PART_MAX || !counts_host)
    return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_hash_partition: bad argument (nkeys %d, nparts %d)", nkeys, nparts);
  PDRS_CUDA(c, cudaSetDevice(c->device));
  const int64_t n = keys[0].len;
  for (int k = 0; k < nkeys; k++) if (keys[k].len != n) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "key column %d length mismatch", k);
  if (n && !perm_dev) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_hash_partition: perm_dev is NULL");
  std::vector<ColView> kv(nkeys);
  for (int k = 0; k < nkeys; k++) PDRS_TRY(pdrs_view_col(c, &keys[k], &kv[k]));
  KeySpec ks;
  PDRS_TRY(pdrs_build_keyspec(c, kv.data(), nkeys, &ks));
  DevBuf cnt;
  PDRS_TRY(cnt.alloc(c, (size_t)2 * nparts * 8, true));
  PartParams p{ks, n, nparts, cnt.as<u64>(), cnt.as<u64>() + nparts, (long long*)perm_dev};
  if (n > 0) {
    int g = pdrs_grid_for(c, n, PART_THREADS);
    int g2 = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 8, (n + PART_THREADS * PART_ITEMS - 1) / (PART_THREADS * PART_ITEMS)));
    switch (ks.nwords) {
      case 1: part_count_kernel<1><<<g, PART_THREADS, 0, c->stream>>>(p); break;
      case 2: part_count_kernel<2><<<g, PART_THREADS, 0, c->stream>>>(p); break;
      default: part_count_kernel<3><<<g, PART_THREADS, 0, c->stream>>>(p); break;
    }
    part_scan_kernel<<<1, 32, 0, c->stream>>>(p.counts, p.cursor, nparts);
    switch (ks.nwords) {
      case 1: part_scatter_kernel<1><<<g2, PART_THREADS, 0, c->stream>>>(p); break;
      case 2: part_scatter_kernel<2><<<g2, PART_THREADS, 0, c->stream>>>(p); break;
      default: part_scatter_kernel<3><<<g2, PART_THREADS, 0, c->stream>>>(p); break;
    }
    c->stats.kernel_launches += 3;
    PDRS_CUDA(c, cudaGetLastError());
  }
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, p.counts, (size_t)std::min(nparts, 64) * 8, cudaMemcpyDeviceToHost, c->stream));
  std::vector<int64_t> big;
  if (nparts > 64) { big.resize(nparts); PDRS_CUDA(c, cudaMemcpyAsync(big.data(), p.counts, (size_t)nparts * 8, cudaMemcpyDeviceToHost, c->stream)); }
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < nparts; i++) counts_host[i] = nparts > 64 ? big[i] : c->pinned_scalars[i];
  return PDRS_OK;
}
