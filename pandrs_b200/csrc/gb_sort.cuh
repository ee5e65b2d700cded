// Device-wide exclusive scan and a STABLE least-significant-digit radix sort of 32-bit payloads by a gathered key (8 bits per pass),
// shared by the row-list grouping (gb_rows.cu) and the dictionary encoder (ingest.cu).  Hand-written: no CUB / thrust.
//   rs_hist_kernel     per-tile (8192 elements) digit histograms, stored digit-major
//   scan_exclusive     one exclusive scan over [256][tiles] = where every (digit, tile) run starts
//   rs_scatter_kernel  every warp walks its 1024 elements IN ORDER, 32 per step, ranks them inside the step with MATCH.ANY and
//                      bumps its private per-digit cursor: no atomics on the output side and equal digits keep their order
#pragma once
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32, RS_WROWS = 1024, RS_TILE = RS_WARPS * RS_WROWS;   // 8192 rows per CTA

// ---------------------------------------------------------------- exclusive scan (any length): tile sums -> scan of the sums -> add
template <typename T>
__global__ void scan_tile_sums_kernel(const T* __restrict__ in, long long n, int per, T* __restrict__ sums) {
  // tile b = elements [b * per * 256, ...): thread t owns `per` consecutive elements
  const long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * per;
  T s = 0;
  for (int i = 0; i < per; i++) if (base + i < n) s += in[base + i];
  __shared__ T sh[32];
  for (int d = 16; d; d >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { T t = 0; for (int w = 0; w < (int)blockDim.x / 32; w++) t += sh[w]; sums[blockIdx.x] = t; }
}
// one CTA: exclusive scan of `m` values in place (thread t owns a contiguous run), total -> *total
template <typename T>
__global__ void scan_single_kernel(T* __restrict__ v, long long m, T* __restrict__ total) {
  const int nt = blockDim.x, t = threadIdx.x;
  const long long per = (m + nt - 1) / nt, lo = min(m, t * per), hi = min(m, lo + per);
  T s = 0;
  for (long long i = lo; i < hi; i++) s += v[i];
  __shared__ T sh[1024];
  sh[t] = s;
  __syncthreads();
  if (t == 0) { T run = 0; for (int i = 0; i < nt; i++) { const T x = sh[i]; sh[i] = run; run += x; } if (total) *total = run; }
  __syncthreads();
  T run = sh[t];
  for (long long i = lo; i < hi; i++) { const T x = v[i]; v[i] = run; run += x; }
}
template <typename T, typename TO>
__global__ void scan_apply_kernel(const T* in, long long n, int per, const T* __restrict__ sums, TO* out) {   // in / out may alias
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * per;
  T s = 0;
  for (int i = 0; i < per; i++) if (base + i < n) s += in[base + i];
  T incl = s;
  for (int d = 1; d < 32; d <<= 1) { const T o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
  __shared__ T sh[32];
  if (lane == 31) sh[warp] = incl;
  __syncthreads();
  T wpre = 0;
  for (int w = 0; w < warp; w++) wpre += sh[w];
  T run = sums[blockIdx.x] + wpre + incl - s;
  for (int i = 0; i < per; i++) if (base + i < n) { const T x = in[base + i]; out[base + i] = (TO)run; run += x; }
}
// out[0 .. n) = exclusive prefix sums of in[0 .. n) (out may alias in when T == TO); *total_dev (optional) = the grand total
template <typename T, typename TO>
inline int32_t scan_exclusive(pdrs_ctx* c, const T* in, long long n, TO* out, T* total_dev) {
  if (n <= 0) { if (total_dev) PDRS_CUDA(c, cudaMemsetAsync(total_dev, 0, sizeof(T), c->stream)); return PDRS_OK; }
  const int per = 8;
  const long long tile = 256ll * per, nb = (n + tile - 1) / tile;
  DevBuf sums;
  PDRS_TRY(sums.alloc(c, (size_t)nb * sizeof(T)));
  scan_tile_sums_kernel<T><<<(unsigned)nb, 256, 0, c->stream>>>(in, n, per, sums.as<T>());
  scan_single_kernel<T><<<1, 1024, 0, c->stream>>>(sums.as<T>(), nb, total_dev);
  scan_apply_kernel<T, TO><<<(unsigned)nb, 256, 0, c->stream>>>(in, n, per, sums.as<T>(), out);
  c->stats.kernel_launches += 3;
  PDRS_CUDA(c, cudaGetLastError());
  return PDRS_OK;
}

// ---------------------------------------------------------------- stable LSD radix sort of 32-bit payloads by a gathered key
// element i of a pass: payload p = pin ? pin[i] : i, key = keys[p], digit = (key >> shift) & 255
template <typename KT>
__device__ __forceinline__ uint32_t rs_digit(const KT* __restrict__ keys, const uint32_t* __restrict__ pin, long long i, int shift, uint32_t* payload) {
  const uint32_t p = pin ? pin[i] : (uint32_t)i;
  *payload = p;
  return (uint32_t)(keys[p] >> shift) & 255u;
}
template <typename KT>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const KT* __restrict__ keys, const uint32_t* __restrict__ pin, long long n, int shift, long long ntiles,
                                                               uint32_t* __restrict__ hist /*[256][ntiles]*/) {
  __shared__ uint32_t h[256];
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = tile * RS_TILE;
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
      if (base + i < n) { uint32_t p; atomicAdd(&h[rs_digit<KT>(keys, pin, base + i, shift, &p)], 1u); }
    }
    __syncthreads();
    hist[(long long)threadIdx.x * ntiles + tile] = h[threadIdx.x];
    __syncthreads();
  }
}
template <typename KT, bool OUT64>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const KT* __restrict__ keys, const uint32_t* __restrict__ pin, long long n, int shift, long long ntiles,
                                                                  const uint32_t* __restrict__ gbase /*[256][ntiles] scanned*/, uint32_t* __restrict__ pout, long long* __restrict__ pout64) {
  __shared__ uint32_t wh[RS_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (int w = 0; w < RS_WARPS; w++) wh[w][threadIdx.x] = 0;
    __syncthreads();
    const long long wbase = tile * RS_TILE + (long long)warp * RS_WROWS;
    // the warp's digit counts
    for (int s = 0; s < RS_WROWS / 32; s++) {
      const long long i = wbase + 32 * s + lane;
      if (i < n) { uint32_t p; atomicAdd(&wh[warp][rs_digit<KT>(keys, pin, i, shift, &p)], 1u); }
    }
    __syncthreads();
    // digit d: where the rows of warp 0, 1, ... of this tile go
    {
      const int d = threadIdx.x;
      uint32_t run = gbase[(long long)d * ntiles + tile];
      for (int w = 0; w < RS_WARPS; w++) { const uint32_t t = wh[w][d]; wh[w][d] = run; run += t; }
    }
    __syncthreads();
    // in order: 32 rows per step, ranked inside the warp (rows of one digit keep their order)
    for (int s = 0; s < RS_WROWS / 32; s++) {
      const long long i = wbase + 32 * s + lane;
      const bool act = i < n;
      uint32_t p = 0;
      const uint32_t d = act ? rs_digit<KT>(keys, pin, i, shift, &p) : 256u + lane;
      const uint32_t m = __match_any_sync(0xFFFFFFFFu, d);
      const int leader = __ffs(m) - 1;
      uint32_t b = 0;
      if (act && lane == leader) { b = wh[warp][d]; wh[warp][d] = b + __popc(m); }
      b = __shfl_sync(0xFFFFFFFFu, b, leader);
      if (act) {
        const uint32_t dst = b + __popc(m & ((1u << lane) - 1u));
        if (OUT64) pout64[dst] = (long long)p; else pout[dst] = p;
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

// Sorts the payloads 0 .. n-1 (or `first_in`) by keys[payload], stably, looking at key bits [0, bits).  The last pass writes
// out64 (when given) or leaves the result in *result (one of the two u32 work buffers).
template <typename KT>
inline int32_t radix_sort_by_key(pdrs_ctx* c, const KT* keys, const uint32_t* first_in, long long n, int bits, uint32_t* buf0, uint32_t* buf1,
                          long long* out64, const uint32_t** result) {
  const int npass = std::max(1, (bits + 7) / 8);
  const long long ntiles = (n + RS_TILE - 1) / RS_TILE;
  DevBuf hist;
  PDRS_TRY(hist.alloc(c, (size_t)256 * ntiles * 4));
  const unsigned grid = (unsigned)std::min<long long>(ntiles, (long long)c->sm_count * 16);
  const uint32_t* in = first_in;
  uint32_t* bufs[2] = {buf0, buf1};
  int nb = in == buf0 ? 1 : 0;
  for (int ps = 0; ps < npass; ps++) {
    const bool last = ps + 1 == npass;
    rs_hist_kernel<KT><<<grid, RS_THREADS, 0, c->stream>>>(keys, in, n, 8 * ps, ntiles, hist.as<uint32_t>());
    c->stats.kernel_launches++;
    PDRS_TRY((scan_exclusive<uint32_t, uint32_t>(c, hist.as<uint32_t>(), 256 * ntiles, hist.as<uint32_t>(), nullptr)));
    if (last && out64) rs_scatter_kernel<KT, true><<<grid, RS_THREADS, 0, c->stream>>>(keys, in, n, 8 * ps, ntiles, hist.as<uint32_t>(), nullptr, out64);
    else rs_scatter_kernel<KT, false><<<grid, RS_THREADS, 0, c->stream>>>(keys, in, n, 8 * ps, ntiles, hist.as<uint32_t>(), bufs[nb], nullptr);
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
    if (!(last && out64)) { in = bufs[nb]; nb ^= 1; }
  }
  if (result) *result = in;
  return PDRS_OK;
}

}  // namespace
