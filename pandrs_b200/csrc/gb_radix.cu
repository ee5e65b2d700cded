// High-cardinality groupby: radix-partitioned rows + L2-resident regions of the global table.
//
// With millions of groups the global table (16 B header + 64 B state per slot) is far larger than the 126 MB
// L2, and updating it in row order costs one or more DRAM round trips per atomic.  The rows (key, value, flags)
// are therefore first partitioned on the TOP bits of the key hash; the table slot is the top bits of the same
// hash (GTable::shift), so bucket b owns one contiguous region of the table, sized to stay in L2.  The update
// kernel then walks the partitioned rows in order (tiles handed out by a global counter, so the rows in flight
// are always one contiguous window = one or two buckets) and its atomics are served by the L2.
// Same state and finalisation as gb_global_kernel (aggregation.rs:500-754).
#include <algorithm>

#include "groupby_kernels.cuh"

#define GR_THREADS 256
#define GR_ITEMS 8
#define GR_TILE (GR_THREADS * GR_ITEMS)
#define GR_MAX_BUCKETS 1024

struct GrSrc {            // one 64-bit key column (variant k1) + the value column of this pass
  const u64* keys; const uint8_t* knull;
  const u64* vals; const uint8_t* vnull;
  const uint8_t* fbits; const uint8_t* fnull;
  long long n;
  int compat_nulls;
};

// flags: bit 0 = value is NULL, bit 1 = key is NULL.  Returns false for rows the filter drops.
__device__ __forceinline__ bool gr_load(const GrSrc& s, long long i, u64* key, u64* val, uint32_t* flags) {
  if (s.fbits) {
    if (!pdrs_bit(s.fbits, i)) return false;
    if (s.fnull && pdrs_bit(s.fnull, i)) return false;
  }
  uint32_t f = 0;
  *key = __ldcs(s.keys + i);
  if (s.knull && pdrs_bit(s.knull, i)) { f |= 2u; *key = 0; }
  *val = s.vals ? __ldcs(s.vals + i) : 0ull;
  if (!s.vals) f |= 1u;
  else if (s.vnull && pdrs_bit(s.vnull, i)) { if (s.compat_nulls) *val = 0; else f |= 1u; }
  *flags = f;
  return true;
}
__device__ __forceinline__ uint32_t gr_bucket(u64 key, uint32_t flags, int log_nb) {
  if (flags & 2u) return 0;
  u64 w[1] = {key};
  return (uint32_t)(key_hash<1>(w) >> (64 - log_nb));
}

__global__ void __launch_bounds__(GR_THREADS) gr_hist_kernel(GrSrc src, int log_nb, u64* __restrict__ hist) {
  __shared__ uint32_t sh[GR_MAX_BUCKETS];
  const int nb = 1 << log_nb;
  for (long long t0 = (long long)blockIdx.x * GR_TILE; t0 < src.n; t0 += (long long)gridDim.x * GR_TILE) {
    for (int i = threadIdx.x; i < nb; i += GR_THREADS) sh[i] = 0;
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < GR_ITEMS; j++) {
      const long long i = t0 + (long long)j * GR_THREADS + threadIdx.x;
      u64 key, val;
      uint32_t fl;
      if (i < src.n && gr_load(src, i, &key, &val, &fl)) atomicAdd(&sh[gr_bucket(key, fl, log_nb)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += GR_THREADS) if (sh[i]) atomicAdd(&hist[i], (u64)sh[i]);
    __syncthreads();
  }
}

__global__ void gr_scan_kernel(const u64* __restrict__ hist, u64* __restrict__ cursor, int nb) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    u64 s = 0;
    for (int i = 0; i < nb; i++) { cursor[i] = s; s += hist[i]; }
    cursor[nb] = s;
  }
}

__global__ void __launch_bounds__(GR_THREADS, 4) gr_scatter_kernel(GrSrc src, int log_nb, u64* __restrict__ cursor, u64* __restrict__ out_keys,
                                                                   u64* __restrict__ out_vals, uint8_t* __restrict__ out_flags) {
  extern __shared__ __align__(16) unsigned char gsm[];
  u64* st_key = reinterpret_cast<u64*>(gsm);                         // [GR_TILE]
  u64* st_val = st_key + GR_TILE;                                    // [GR_TILE]
  uint32_t* st_dst = reinterpret_cast<uint32_t*>(st_val + GR_TILE);  // [GR_TILE]
  uint8_t* st_fl = reinterpret_cast<uint8_t*>(st_dst + GR_TILE);     // [GR_TILE]
  __shared__ uint32_t hist[GR_MAX_BUCKETS], lbase[GR_MAX_BUCKETS];
  __shared__ u64 gbase[GR_MAX_BUCKETS];
  __shared__ uint32_t total;
  const int nb = 1 << log_nb;
  for (long long t0 = (long long)blockIdx.x * GR_TILE; t0 < src.n; t0 += (long long)gridDim.x * GR_TILE) {
    for (int i = threadIdx.x; i < nb; i += GR_THREADS) hist[i] = 0;
    __syncthreads();
    u64 key[GR_ITEMS], val[GR_ITEMS];
    uint32_t bkt[GR_ITEMS], rank[GR_ITEMS], fl[GR_ITEMS];
#pragma unroll
    for (int j = 0; j < GR_ITEMS; j++) {
      const long long i = t0 + (long long)j * GR_THREADS + threadIdx.x;
      bkt[j] = 0xFFFFFFFFu;
      if (i < src.n && gr_load(src, i, &key[j], &val[j], &fl[j])) { bkt[j] = gr_bucket(key[j], fl[j], log_nb); rank[j] = atomicAdd(&hist[bkt[j]], 1u); }
    }
    __syncthreads();
    if (threadIdx.x < 32) {            // exclusive scan of the bucket counts by one warp
      uint32_t run = 0;
      for (int b0 = 0; b0 < nb; b0 += 32) {
        const int b = b0 + threadIdx.x;
        uint32_t c = b < nb ? hist[b] : 0, incl = c;
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)threadIdx.x >= d) incl += t; }
        if (b < nb) lbase[b] = run + incl - c;
        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
      }
      if (threadIdx.x == 0) total = run;
    }
    for (int b = threadIdx.x; b < nb; b += GR_THREADS) gbase[b] = hist[b] ? atomicAdd(&cursor[b], (u64)hist[b]) : 0ull;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < GR_ITEMS; j++) {
      if (bkt[j] == 0xFFFFFFFFu) continue;
      const uint32_t pos = lbase[bkt[j]] + rank[j];
      st_key[pos] = key[j];
      st_val[pos] = val[j];
      st_fl[pos] = (uint8_t)fl[j];
      st_dst[pos] = (uint32_t)(gbase[bkt[j]] + rank[j]);
    }
    __syncthreads();
    for (uint32_t pos = threadIdx.x; pos < total; pos += GR_THREADS) {
      const uint32_t d = st_dst[pos];
      out_keys[d] = st_key[pos];
      out_vals[d] = st_val[pos];
      out_flags[d] = st_fl[pos];
    }
    __syncthreads();
  }
}

// Update of the global table from the partitioned rows, tiles in order.
template <typename VT, int FLAGS>
__global__ void __launch_bounds__(256) gr_update_kernel(const u64* __restrict__ pkeys, const u64* __restrict__ pvals, const uint8_t* __restrict__ pflags,
                                                        long long n, GTable gt, int count_rows, u64* tile_ctr) {
  using T = ValTraits<VT>;
  __shared__ long long sh_tile;
  const long long ntiles = (n + GR_TILE - 1) / GR_TILE;
  for (;;) {
    if (threadIdx.x == 0) sh_tile = (long long)atomicAdd(tile_ctr, 1ull);
    __syncthreads();
    const long long tile = sh_tile;
    __syncthreads();
    if (tile >= ntiles) break;
    const long long lo = tile * GR_TILE;
#pragma unroll 2
    for (int j = 0; j < GR_ITEMS; j++) {       // warp-uniform trip count: g_find_or_insert is warp-synchronous
      const long long i = lo + (long long)j * GR_THREADS + threadIdx.x;
      const bool active = i < n;
      u64 w[1] = {0};
      uint32_t fl = 0;
      VT v = VT(0);
      if (active) { w[0] = __ldcs(pkeys + i); fl = __ldcs(pflags + i); v = T::from_bits(__ldcs(pvals + i)); }
      const bool knull = fl & 2u;
      long long gs = g_find_or_insert<1>(gt, w, active && !knull);
      if (active && knull) { gs = gt.slots; if (!(ld_cg_u64(&gt.hdr[gs].rowsw) & GB_FULL)) atomicOr(&gt.hdr[gs].rowsw, GB_FULL); }
      if (active && gs >= 0) g_update_row<VT, FLAGS>(gt, gs, count_rows != 0, !(fl & 1u), v);
    }
  }
}

// One aggregation pass over a 64-bit key column through the partitioned path.  `gp` carries the table, the value
// column and the filter like for gb_global_kernel.  Returns PDRS_ERR_UNSUPPORTED when the path does not apply.
int32_t gb_radix_pass(pdrs_ctx* c, const GbParams& gp, int is_int, int flags, float* kernel_ms) {
  const long long n = gp.n;
  if (n >= (1ll << 32)) return PDRS_ERR_UNSUPPORTED;
  const size_t table_bytes = (size_t)(gp.gt.slots + 1) * (sizeof(GHdr) + (gp.gt.st ? sizeof(GState) : 0));
  int log_nb = 2;
  while (log_nb < 10 && (table_bytes >> log_nb) > (40ull << 20)) log_nb++;
  if (gp.gt.slots < (1ll << log_nb)) return PDRS_ERR_UNSUPPORTED;
  const int nb = 1 << log_nb;
  GrSrc src{reinterpret_cast<const u64*>(gp.ks.c[0].data), gp.ks.c[0].nulls, reinterpret_cast<const u64*>(gp.val), gp.vnull, gp.fbits, gp.fnull, n, gp.compat_nulls};
  DevBuf hist, pkeys, pvals, pflags;
  PDRS_TRY(hist.alloc(c, (size_t)(2 * nb + 4) * 8, true));
  u64* h = hist.as<u64>();
  const int ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 4, (n + GR_TILE - 1) / GR_TILE));
  gr_hist_kernel<<<ctas, GR_THREADS, 0, c->stream>>>(src, log_nb, h);
  gr_scan_kernel<<<1, 32, 0, c->stream>>>(h, h + nb, nb);
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 8, h + 2 * nb, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  const long long m = c->pinned_scalars[8];
  PDRS_TRY(pkeys.alloc(c, (size_t)std::max<long long>(m, 1) * 8));
  PDRS_TRY(pvals.alloc(c, (size_t)std::max<long long>(m, 1) * 8));
  PDRS_TRY(pflags.alloc(c, (size_t)std::max<long long>(m, 1) + 16));
  const size_t smem = (size_t)GR_TILE * 21 + 16;
  PDRS_CUDA(c, cudaFuncSetAttribute(gr_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: every call
  gr_scatter_kernel<<<ctas, GR_THREADS, smem, c->stream>>>(src, log_nb, h + nb, pkeys.as<u64>(), pvals.as<u64>(), pflags.as<uint8_t>());
  PDRS_CUDA(c, cudaGetLastError());
  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
  u64* ctr = h + 2 * nb + 2;      // zeroed with the histogram
  const int uctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 8, (m + GR_TILE - 1) / GR_TILE));
  if (m > 0) {
    if (!is_int) {
      if (flags == GB_SUM) gr_update_kernel<double, GB_SUM><<<uctas, 256, 0, c->stream>>>(pkeys.as<u64>(), pvals.as<u64>(), pflags.as<uint8_t>(), m, gp.gt, gp.count_rows, ctr);
      else gr_update_kernel<double, GB_ALL><<<uctas, 256, 0, c->stream>>>(pkeys.as<u64>(), pvals.as<u64>(), pflags.as<uint8_t>(), m, gp.gt, gp.count_rows, ctr);
    } else {
      if (flags == GB_SUM) gr_update_kernel<long long, GB_SUM><<<uctas, 256, 0, c->stream>>>(pkeys.as<u64>(), pvals.as<u64>(), pflags.as<uint8_t>(), m, gp.gt, gp.count_rows, ctr);
      else gr_update_kernel<long long, GB_ALL><<<uctas, 256, 0, c->stream>>>(pkeys.as<u64>(), pvals.as<u64>(), pflags.as<uint8_t>(), m, gp.gt, gp.count_rows, ctr);
    }
    PDRS_CUDA(c, cudaGetLastError());
  }
  c->stats.kernel_launches += 4;
  if (c->opt_timing) {
    PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
    PDRS_CUDA(c, cudaEventSynchronize(c->ev_b));
    PDRS_CUDA(c, cudaEventElapsedTime(kernel_ms, c->ev_a, c->ev_b));
  }
  return PDRS_OK;
}
