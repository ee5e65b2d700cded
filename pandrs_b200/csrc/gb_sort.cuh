// Device-wide exclusive scan and a STABLE least-significant-digit radix sort of (key, 32-bit payload) pairs (8 bits per pass),
// shared by the row-list grouping (gb_rows.cu) and the dictionary encoder (ingest.cu).  Hand-written: no CUB / thrust.
//   rs_hist_kernel     per-tile (8192 elements) digit histograms, stored digit-major
//   scan_exclusive     one exclusive scan over [256][tiles] = where every (digit, tile) run starts
//   rs_scatter_kernel  every warp walks its 1024 elements IN ORDER, 32 per step, ranks them inside the step with MATCH.ANY and
//                      bumps its private per-digit cursor (no atomics, equal digits keep their order); the tile is staged in shared
//                      memory in output order and written out as contiguous runs
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

// ---------------------------------------------------------------- stable LSD radix sort of (key, 32-bit payload) pairs
// Keys travel WITH the payloads (sequential reads in every pass; a gathered key cost one random DRAM sector per element and pass:
// 20 ms per 1e9 rows just for the histogram).  A pass over one 8192-element tile:
//   per-warp digit counts -> tile-local exclusive prefix per (digit, warp) -> every warp walks its 1024 elements IN ORDER, 32 per
//   step, ranks them inside the step with MATCH.ANY and bumps its private cursor (stable, no atomics) -> the element is staged in
//   shared memory at its tile-local position -> the tile is written out with consecutive threads on consecutive staged elements:
//   the run of a digit (32 elements on average) goes to consecutive addresses, i.e. whole sectors instead of one 4-byte store each.
template <typename KT> __device__ __forceinline__ uint32_t rs_dig(KT k, int shift) { return (uint32_t)(k >> shift) & 255u; }

template <typename KT>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const KT* __restrict__ keys, long long n, int shift, long long ntiles, uint32_t* __restrict__ hist /*[256][ntiles]*/) {
  __shared__ uint32_t h[256];
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = tile * RS_TILE;
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS)
      if (base + i < n) atomicAdd(&h[rs_dig<KT>(keys[base + i], shift)], 1u);
    __syncthreads();
    hist[(long long)threadIdx.x * ntiles + tile] = h[threadIdx.x];
    __syncthreads();
  }
}
template <typename KT> constexpr size_t rs_scatter_smem() { return (size_t)RS_TILE * (sizeof(KT) + 4) + (size_t)RS_WARPS * 256 * 4 + 256 * 4 + 64; }

// pay_in == nullptr: payload = element index.  keys_out may be nullptr (last pass).  OUT64: payloads are written as i64.
template <typename KT, bool OUT64>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const KT* __restrict__ keys_in, const uint32_t* __restrict__ pay_in, long long n, int shift, long long ntiles,
                                                                  const uint32_t* __restrict__ gbase /*[256][ntiles] scanned*/, KT* __restrict__ keys_out,
                                                                  uint32_t* __restrict__ pay_out, long long* __restrict__ pay_out64) {
  extern __shared__ __align__(16) unsigned char rs_sm[];
  KT* st_k = reinterpret_cast<KT*>(rs_sm);                               // [RS_TILE] staged keys, tile-local order
  uint32_t* st_p = reinterpret_cast<uint32_t*>(st_k + RS_TILE);          // [RS_TILE] staged payloads
  uint32_t (*wh)[256] = reinterpret_cast<uint32_t (*)[256]>(st_p + RS_TILE);   // [RS_WARPS][256] counts, then cursors
  uint32_t* gdelta = reinterpret_cast<uint32_t*>(wh + RS_WARPS);        // [256] global position - tile-local position of a digit's run
  __shared__ uint32_t wsum[RS_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (int w = 0; w < RS_WARPS; w++) wh[w][threadIdx.x] = 0;
    __syncthreads();
    const long long tbase = tile * RS_TILE, wbase = tbase + (long long)warp * RS_WROWS;
    const int tcount = (int)min((long long)RS_TILE, n - tbase);
    for (int s = 0; s < RS_WROWS / 32; s++) {
      const long long i = wbase + 32 * s + lane;
      if (i < n) atomicAdd(&wh[warp][rs_dig<KT>(keys_in[i], shift)], 1u);
    }
    __syncthreads();
    {   // digit d = threadIdx.x: exclusive prefix over digits (tile-local start of the digit's run), then over the warps inside it
      const int d = threadIdx.x;
      uint32_t tot = 0;
      for (int w = 0; w < RS_WARPS; w++) tot += wh[w][d];
      uint32_t incl = tot;
      for (int k = 1; k < 32; k <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, k); if (lane >= k) incl += o; }
      if (lane == 31) wsum[warp] = incl;
      __syncthreads();
      uint32_t wpre = 0;
      for (int w = 0; w < warp; w++) wpre += wsum[w];
      uint32_t run = wpre + incl - tot;
      gdelta[d] = gbase[(long long)d * ntiles + tile] - run;
      for (int w = 0; w < RS_WARPS; w++) { const uint32_t t = wh[w][d]; wh[w][d] = run; run += t; }
    }
    __syncthreads();
    for (int s = 0; s < RS_WROWS / 32; s++) {
      const long long i = wbase + 32 * s + lane;
      const bool act = i < n;
      KT k = 0;
      if (act) k = keys_in[i];
      const uint32_t p = act ? (pay_in ? pay_in[i] : (uint32_t)i) : 0u;
      const uint32_t d = act ? rs_dig<KT>(k, shift) : 256u + lane;
      const uint32_t m = __match_any_sync(0xFFFFFFFFu, d);
      const int leader = __ffs(m) - 1;
      uint32_t b = 0;
      if (act && lane == leader) { b = wh[warp][d]; wh[warp][d] = b + __popc(m); }
      b = __shfl_sync(0xFFFFFFFFu, b, leader);
      if (act) {
        const uint32_t pos = b + __popc(m & ((1u << lane) - 1u));
        st_k[pos] = k;
        st_p[pos] = p;
      }
      __syncwarp();
    }
    __syncthreads();
    for (int pos = threadIdx.x; pos < tcount; pos += RS_THREADS) {
      const KT k = st_k[pos];
      const uint32_t g = (uint32_t)pos + gdelta[rs_dig<KT>(k, shift)];
      if (keys_out) keys_out[g] = k;
      if (OUT64) pay_out64[g] = (long long)st_p[pos]; else pay_out[g] = st_p[pos];
    }
    __syncthreads();
  }
}

// Sorts the pairs (keys[i], pay[i]) (pay == nullptr: payload i) by key bits [0, bits), stably.  kbuf / pbuf: two work buffers each
// (n elements; unused when one pass suffices and out64 is given).  The last pass writes the payloads to out64 (as i64) when given,
// else leaves them in *result_pay (one of pbuf).  The sorted keys are not kept.
template <typename KT>
inline int32_t radix_sort_pairs(pdrs_ctx* c, const KT* keys, const uint32_t* pay, long long n, int bits, KT* kbuf0, KT* kbuf1, uint32_t* pbuf0, uint32_t* pbuf1,
                                long long* out64, const uint32_t** result_pay) {
  const int npass = std::max(1, (bits + 7) / 8);
  const long long ntiles = (n + RS_TILE - 1) / RS_TILE;
  DevBuf hist;
  PDRS_TRY(hist.alloc(c, (size_t)256 * ntiles * 4));
  const size_t smem = rs_scatter_smem<KT>();
  PDRS_CUDA(c, cudaFuncSetAttribute(rs_scatter_kernel<KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     // per device: every call
  PDRS_CUDA(c, cudaFuncSetAttribute(rs_scatter_kernel<KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)std::min<long long>(ntiles, (long long)c->sm_count * 16);
  const KT* kin = keys;
  const uint32_t* pin = pay;
  KT* kb[2] = {kbuf0, kbuf1};
  uint32_t* pb2[2] = {pbuf0, pbuf1};
  int nb = (pin == pbuf0 || kin == kbuf0) ? 1 : 0;
  for (int ps = 0; ps < npass; ps++) {
    const bool last = ps + 1 == npass;
    rs_hist_kernel<KT><<<grid, RS_THREADS, 0, c->stream>>>(kin, n, 8 * ps, ntiles, hist.as<uint32_t>());
    c->stats.kernel_launches++;
    pdrs_trace(c, "  sort: histogram");
    PDRS_TRY((scan_exclusive<uint32_t, uint32_t>(c, hist.as<uint32_t>(), 256 * ntiles, hist.as<uint32_t>(), nullptr)));
    pdrs_trace(c, "  sort: scan");
    if (last && out64) rs_scatter_kernel<KT, true><<<grid, RS_THREADS, smem, c->stream>>>(kin, pin, n, 8 * ps, ntiles, hist.as<uint32_t>(), nullptr, nullptr, out64);
    else rs_scatter_kernel<KT, false><<<grid, RS_THREADS, smem, c->stream>>>(kin, pin, n, 8 * ps, ntiles, hist.as<uint32_t>(), last ? nullptr : kb[nb], pb2[nb], nullptr);
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
    pdrs_trace(c, "  sort: scatter");
    if (!(last && out64)) { kin = kb[nb]; pin = pb2[nb]; nb ^= 1; }
  }
  if (result_pay) *result_pay = pin;
  return PDRS_OK;
}

}  // namespace
