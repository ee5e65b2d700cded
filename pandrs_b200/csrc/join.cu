// Hash join (inner / left) on one key column: build a global open-addressing table on the right
// frame, probe it with the left frame, and materialise (left_row, right_row) index pairs with a
// two-pass count / scan / write so that the output is left-row-major like the reference.
//
// Replaces the build / probe loops of OptimizedDataFrame::join_impl
// (src/optimized/split_dataframe/join.rs:107-208):
//   BUILD  :107-142  right_key_to_indices[key].push(i), NULL keys skipped
//   PROBE  :146-208  for every left row in order: all matching right rows in ascending order,
//                    (i, None) when nothing matches and the join is Left; NULL left keys are
//                    skipped entirely, also for Left (:152)
//
//   jpart_*             (large inputs) both sides are radix-partitioned on the top bits of the key hash into NB
//                       buckets of (key, row id); the table slot is a monotone function of the same hash, so one
//                       bucket owns one contiguous <= 32 MB region of the table and build / probe of a bucket run
//                       out of the 126 MB L2 instead of DRAM.  Pairs then come out in bucket order.
//   join_build_kernel   one CAS claims a 16-byte slot {key, head}; rows with the same key are chained
//                       through next[] (head carries a MULTI flag so unique keys never touch next[])
//   join_probe_kernel   one pass over the left keys: stash[i] = head word of the matching slot (or
//                       NOMATCH / NULLKEY) and a per-CTA count of output rows
//   join_write_kernel   per-CTA base offsets from the scan of those counts, block-level exclusive
//                       scans inside each CTA's contiguous row range, ordered writes of the pairs
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"

typedef unsigned long long u64;

struct JSlot { u64 key; long long head; };
static constexpr long long J_EMPTY = -1, J_BUSY = -2, J_NOMATCH = -1, J_NULLKEY = -3;
static constexpr long long J_MULTI = 1ll << 62;
#define JOIN_THREADS 256
#define JOIN_ITEMS 4

struct JKeyCol { const void* data; const uint8_t* nulls; int dtype; };

__device__ __forceinline__ u64 jhash(u64 k) {
  u64 h = k * 0x9E3779B97F4A7C15ull;
  h ^= h >> 32;
  return h * 0xD6E8FEB86659FD93ull;
}

// slot = floor(hash_hi32 * slots / 2^32): monotone in the hash, so the rows of one radix bucket (top hash bits)
// fall into one contiguous region of the table; `slots` need not be a power of two
__device__ __forceinline__ u64 jslot(u64 key, u64 slots) { return (u64)__umulhi((uint32_t)(jhash(key) >> 32), (uint32_t)slots); }   // slots < 2^32

// `Some(v) -> v.to_string()` equality restated on the physical values (join.rs:112-139)
__device__ __forceinline__ bool jload_key(const JKeyCol& c, long long row, u64* k) {
  if (c.nulls && pdrs_bit(c.nulls, row)) return false;
  switch (c.dtype) {
    case PDRS_I64: *k = (u64)__ldcs((const long long*)c.data + row); break;
    case PDRS_F64: {
      double d = __ldcs((const double*)c.data + row);
      *k = (d != d) ? 0x7FF8000000000000ull : (u64)__double_as_longlong(d);
      break;
    }
    case PDRS_I32: *k = (u64)(uint32_t)__ldcs((const int*)c.data + row); break;
    case PDRS_DICT_U32: *k = (u64)__ldcs((const uint32_t*)c.data + row); break;
    default: *k = pdrs_bit((const uint8_t*)c.data, row); break;
  }
  return true;
}

__device__ __forceinline__ ulonglong2 ld_volatile_v2(const void* p) {
  ulonglong2 r;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ u64 ld_volatile_u64(const void* p) {
  u64 r;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
  return r;
}

// Rows come either straight from a key column (row id = position) or from a radix-partitioned copy
// (canonical 64-bit keys + original row ids, NULL keys already dropped).
struct JSrc { JKeyCol col; const u64* pkeys; const uint32_t* prows; };
__device__ __forceinline__ bool jsrc_load(const JSrc& s, long long i, u64* key, long long* row) {
  if (s.pkeys) { *key = __ldcs(s.pkeys + i); *row = (long long)__ldcs(s.prows + i); return true; }
  *row = i;
  return jload_key(s.col, i, key);
}

// Insertion: the slot is claimed by a CAS on the key word itself (all-ones = empty), so there is no "being
// published" state, nobody ever waits and no fence is needed (the probe runs in a later kernel).  Rows with the
// same key are chained: head <- row (exchange), next[row] <- old head; next[] is pre-filled with -1 so that
// unique keys never write it, and a second row of a key sets the MULTI flag of the head.
// The one key whose bit pattern is all-ones lives in the reserved slot `slots` (never reached by probing).
static constexpr u64 J_EMPTY_KEY = ~0ull;
#define JB_TILE 2048
__global__ void __launch_bounds__(256) join_build_kernel(JSlot* tab, u64 slots, long long* next, JSrc src, long long n, u64* fail) {
  // tiles are handed out in order by a global counter (fail[1]): whatever the relative speed of the CTAs, the rows
  // in flight form one contiguous window, i.e. they stay inside one or two radix buckets = L2-resident table regions
  __shared__ long long sh_tile;
  const long long ntiles = (n + JB_TILE - 1) / JB_TILE;
  for (;;) {
    if (threadIdx.x == 0) sh_tile = (long long)atomicAdd(&fail[1], 1ull);
    __syncthreads();
    const long long tile = sh_tile;
    __syncthreads();
    if (tile >= ntiles) break;
    const long long hi = min(n, (tile + 1) * JB_TILE);
    for (long long i = tile * JB_TILE + threadIdx.x; i < hi; i += blockDim.x) {
      u64 key = 0;
      long long r = 0;
      if (!jsrc_load(src, i, &key, &r)) continue;
      u64 slot = jslot(key, slots);
      bool done = false;
      if (key == J_EMPTY_KEY) slot = slots;
      for (u64 probe = 0; probe <= slots; probe++) {
        u64 k = key == J_EMPTY_KEY ? key : __ldcg(&tab[slot].key);
        if (k == J_EMPTY_KEY && key != J_EMPTY_KEY) k = atomicCAS(&tab[slot].key, J_EMPTY_KEY, key), k = (k == J_EMPTY_KEY) ? key : k;
        if (k == key) {
          const long long old = (long long)atomicExch(reinterpret_cast<u64*>(&tab[slot].head), (u64)r);
          if (old != J_EMPTY) {      // not the first row of this key: link, and flag the head
            next[r] = old & ~J_MULTI;
            atomicOr(reinterpret_cast<u64*>(&tab[slot].head), (u64)J_MULTI);
          }
          done = true;
          break;
        }
        slot = slot + 1 >= slots ? 0 : slot + 1;
      }
      if (!done) atomicAdd(fail, 1ull);
    }
  }
}

// L2 cache policies: the table region of the current radix bucket must stay in L2 (evict_last) while the key /
// stash / output streams pass through it once (evict_first).
__device__ __forceinline__ u64 l2_policy_evict_last() { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 l2_policy_evict_first() { u64 p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 ld_stream_u64(const u64* a, u64 pol) { u64 v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(pol)); return v; }
__device__ __forceinline__ void st_stream_u64(long long* a, long long v, u64 pol) { asm volatile("st.global.L1::no_allocate.L2::cache_hint.u64 [%0], %1, %2;" :: "l"(a), "l"(v), "l"(pol) : "memory"); }

// Probing reads one 32-byte sector (two slots) per step: `slots` is even and pairs are sector-aligned.
// Returns the head word of the matching slot or J_NOMATCH.
__device__ __forceinline__ long long jprobe(const JSlot* tab, u64 slots, u64 key, u64 pol) {
  if (key == J_EMPTY_KEY) { const long long h = (long long)__ldg(reinterpret_cast<const u64*>(&tab[slots].head)); return h == J_EMPTY ? J_NOMATCH : h; }
  const uint32_t nslots = (uint32_t)slots;
  const uint32_t home = (uint32_t)jslot(key, slots);
  uint32_t pair = home & ~1u;
  bool skip_even = home & 1u;          // the even slot of the first pair precedes the home slot: not part of the probe sequence
  for (uint32_t step = 0; step <= nslots; step += 2) {
    ulonglong4 s;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0, %1, %2, %3}, [%4], %5;" : "=l"(s.x), "=l"(s.y), "=l"(s.z), "=l"(s.w) : "l"(tab + pair), "l"(pol));
    if (!skip_even) {
      if (s.x == key) return (long long)s.y;
      if (s.x == J_EMPTY_KEY) return J_NOMATCH;
    }
    if (s.z == key) return (long long)s.w;
    if (s.z == J_EMPTY_KEY) return J_NOMATCH;
    skip_even = false;
    pair = pair + 2 >= nslots ? 0 : pair + 2;
  }
  return J_NOMATCH;
}

__device__ __forceinline__ long long jcount(long long stash, const long long* next, int left_join) {
  if (stash == J_NULLKEY) return 0;
  if (stash == J_NOMATCH) return left_join ? 1 : 0;
  if (!(stash & J_MULTI)) return 1;
  long long c = 0;
  for (long long r = stash & ~J_MULTI; r >= 0; r = __ldg(next + r)) c++;
  return c;
}

// Rows are processed in tiles of JOIN_TILE consecutive positions, tile t by CTA t mod grid: the output stays in
// position order (tile offsets come from a scan of the per-tile counts) and, on the partitioned path, all CTAs
// work inside the same radix bucket at any time (L2-resident table region).
#define JOIN_TILE (JOIN_THREADS * JOIN_ITEMS * 4)

__global__ void __launch_bounds__(JOIN_THREADS) join_probe_kernel(const JSlot* __restrict__ tab, u64 slots, const long long* __restrict__ next, JSrc src, long long n,
                                                                  int left_join, long long* __restrict__ stash, u64* __restrict__ tile_counts, u64* __restrict__ tile_ctr) {
  __shared__ u64 wsum[JOIN_THREADS / 32];
  const long long ntiles = (n + JOIN_TILE - 1) / JOIN_TILE;
  const u64 pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  __shared__ long long sh_tile;
  for (;;) {
    // tiles in order from a global counter: the probes in flight stay inside one or two radix buckets
    if (threadIdx.x == 0) sh_tile = (long long)atomicAdd(tile_ctr, 1ull);
    __syncthreads();
    const long long tile = sh_tile;
    if (tile >= ntiles) break;
    const long long lo = tile * JOIN_TILE, hi = min(n, lo + JOIN_TILE);
    u64 cnt = 0;
    for (long long i0 = lo + threadIdx.x; i0 < hi; i0 += (long long)JOIN_THREADS * JOIN_ITEMS) {
      u64 key[JOIN_ITEMS];
      bool ok[JOIN_ITEMS];
#pragma unroll
      for (int j = 0; j < JOIN_ITEMS; j++) {
        long long i = i0 + (long long)j * JOIN_THREADS;
        ok[j] = false;
        if (i < hi) { if (src.pkeys) { key[j] = ld_stream_u64(src.pkeys + i, pol_stream); ok[j] = true; } else ok[j] = jload_key(src.col, i, &key[j]); }
      }
      long long st[JOIN_ITEMS];
#pragma unroll
      for (int j = 0; j < JOIN_ITEMS; j++) st[j] = ok[j] ? jprobe(tab, slots, key[j], pol_keep) : J_NULLKEY;
#pragma unroll
      for (int j = 0; j < JOIN_ITEMS; j++) {
        long long i = i0 + (long long)j * JOIN_THREADS;
        if (i < hi) { st_stream_u64(stash + i, st[j], pol_stream); cnt += (u64)jcount(st[j], next, left_join); }
      }
    }
    for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { u64 t = 0; for (int w = 0; w < JOIN_THREADS / 32; w++) t += wsum[w]; tile_counts[tile] = t; }
    __syncthreads();
  }
}

__global__ void join_scan_kernel(u64* v, int n, u64* total) {   // exclusive scan of the per-tile counts (n = rows / 4096)
  __shared__ u64 wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + threadIdx.x;
    u64 x = i < n ? v[i] : 0, incl = x;
    for (int d = 1; d < 32; d <<= 1) { u64 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      u64 w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0, wi = w;
      for (int d = 1; d < 32; d <<= 1) { u64 t = __shfl_up_sync(0xFFFFFFFFu, wi, d); if (lane >= d) wi += t; }
      wsum[lane] = wi - w;   // exclusive
      if (lane == 31) wsum[31] = wi - w, v[n + 1] = wi;   // scratch: block total
    }
    __syncthreads();
    if (i < n) v[i] = carry + wsum[warp] + incl - x;
    __syncthreads();
    carry += v[n + 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(JOIN_THREADS) join_write_kernel(const long long* __restrict__ stash, const long long* __restrict__ next, long long n, int left_join,
                                                                  const u64* __restrict__ tile_offsets, const uint32_t* __restrict__ prows,
                                                                  long long* __restrict__ out_l, long long* __restrict__ out_r) {
  __shared__ u64 wsum[JOIN_THREADS / 32];
  __shared__ u64 sh_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ntiles = (n + JOIN_TILE - 1) / JOIN_TILE;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long lo = tile * JOIN_TILE, hi = min(n, lo + JOIN_TILE);
    if (threadIdx.x == 0) sh_base = tile_offsets[tile];
    __syncthreads();
    // thread t owns JOIN_ITEMS consecutive rows of every chunk, so positions are monotone in the row position
    for (long long c0 = lo; c0 < hi; c0 += (long long)JOIN_THREADS * JOIN_ITEMS) {
      long long st[JOIN_ITEMS];
      u64 cn[JOIN_ITEMS], mine = 0;
#pragma unroll
      for (int j = 0; j < JOIN_ITEMS; j++) {
        long long i = c0 + (long long)threadIdx.x * JOIN_ITEMS + j;
        st[j] = i < hi ? __ldcs(stash + i) : J_NULLKEY;
        cn[j] = (u64)jcount(st[j], next, left_join);
        mine += cn[j];
      }
      u64 incl = mine;
      for (int d = 1; d < 32; d <<= 1) { u64 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
      if (lane == 31) wsum[warp] = incl;
      __syncthreads();
      u64 wbase = 0, total = 0;
      for (int w = 0; w < JOIN_THREADS / 32; w++) { u64 t = wsum[w]; if (w < warp) wbase += t; total += t; }
      u64 pos = sh_base + wbase + incl - mine;
#pragma unroll
      for (int j = 0; j < JOIN_ITEMS; j++) {
        long long i = c0 + (long long)threadIdx.x * JOIN_ITEMS + j;
        if (cn[j] == 0) continue;
        if (prows) i = (long long)__ldg(prows + i);      // partitioned probe side: position -> original left row
        if (st[j] == J_NOMATCH) { out_l[pos] = i; out_r[pos] = -1; pos++; }
        else if (!(st[j] & J_MULTI)) { out_l[pos] = i; out_r[pos] = st[j]; pos++; }
        else {
          const u64 p0 = pos;
          for (long long r = st[j] & ~J_MULTI; r >= 0; r = __ldg(next + r)) {   // chain order is arbitrary:
            u64 q = pos++;                                                     // insert in ascending right row (join.rs:158-161)
            while (q > p0 && out_r[q - 1] > r) { out_r[q] = out_r[q - 1]; q--; }
            out_r[q] = r;
          }
          for (u64 q = p0; q < pos; q++) out_l[q] = i;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) sh_base += total;
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------- radix partitioning of one join side
#define JP_THREADS 256
#define JP_ITEMS 8
#define JP_TILE (JP_THREADS * JP_ITEMS)
#define JP_MAX_BUCKETS 1024

__device__ __forceinline__ uint32_t jbucket(u64 key, int log_nb) { return (uint32_t)(jhash(key) >> (64 - log_nb)); }

__global__ void __launch_bounds__(JP_THREADS) jpart_hist_kernel(JKeyCol col, long long n, int log_nb, u64* __restrict__ hist) {
  __shared__ uint32_t sh[JP_MAX_BUCKETS];
  const int nb = 1 << log_nb;
  for (long long t0 = (long long)blockIdx.x * JP_TILE; t0 < n; t0 += (long long)gridDim.x * JP_TILE) {
    for (int i = threadIdx.x; i < nb; i += JP_THREADS) sh[i] = 0;
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < JP_ITEMS; j++) {
      const long long i = t0 + (long long)j * JP_THREADS + threadIdx.x;
      u64 key;
      if (i < n && jload_key(col, i, &key)) atomicAdd(&sh[jbucket(key, log_nb)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += JP_THREADS) if (sh[i]) atomicAdd(&hist[i], (u64)sh[i]);
    __syncthreads();
  }
}

__global__ void jpart_scan_kernel(const u64* __restrict__ hist, u64* __restrict__ cursor, u64* __restrict__ starts, int nb) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    u64 s = 0;
    for (int i = 0; i < nb; i++) { cursor[i] = s; starts[i] = s; s += hist[i]; }
    starts[nb] = s;
  }
}

// One tile of 4096 rows per iteration: bucket histogram in shared memory (the returned ticket is the row's rank
// inside its bucket), one global reservation per bucket, rows staged in shared memory in bucket order, then
// written out in coalesced runs.
__global__ void __launch_bounds__(JP_THREADS, 4) jpart_scatter_kernel(JKeyCol col, long long n, int log_nb, u64* __restrict__ cursor,
                                                                   u64* __restrict__ out_keys, uint32_t* __restrict__ out_rows) {
  extern __shared__ __align__(16) unsigned char jsm[];
  u64* st_key = reinterpret_cast<u64*>(jsm);                       // [JP_TILE]
  uint32_t* st_row = reinterpret_cast<uint32_t*>(st_key + JP_TILE);      // [JP_TILE]
  uint32_t* st_dst = st_row + JP_TILE;                                   // [JP_TILE]
  __shared__ uint32_t hist[JP_MAX_BUCKETS], lbase[JP_MAX_BUCKETS];
  __shared__ u64 gbase[JP_MAX_BUCKETS];
  __shared__ uint32_t total;
  const int nb = 1 << log_nb;
  for (long long t0 = (long long)blockIdx.x * JP_TILE; t0 < n; t0 += (long long)gridDim.x * JP_TILE) {
    for (int i = threadIdx.x; i < nb; i += JP_THREADS) hist[i] = 0;
    __syncthreads();
    u64 key[JP_ITEMS];
    uint32_t bkt[JP_ITEMS], rank[JP_ITEMS];
#pragma unroll
    for (int j = 0; j < JP_ITEMS; j++) {
      const long long i = t0 + (long long)j * JP_THREADS + threadIdx.x;
      bkt[j] = 0xFFFFFFFFu;
      if (i < n && jload_key(col, i, &key[j])) { bkt[j] = jbucket(key[j], log_nb); rank[j] = atomicAdd(&hist[bkt[j]], 1u); }
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
    for (int b = threadIdx.x; b < nb; b += JP_THREADS) gbase[b] = hist[b] ? atomicAdd(&cursor[b], (u64)hist[b]) : 0ull;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < JP_ITEMS; j++) {
      if (bkt[j] == 0xFFFFFFFFu) continue;
      const uint32_t pos = lbase[bkt[j]] + rank[j];
      st_key[pos] = key[j];
      st_row[pos] = (uint32_t)(t0 + (long long)j * JP_THREADS + threadIdx.x);
      st_dst[pos] = (uint32_t)(gbase[bkt[j]] + rank[j]);
    }
    __syncthreads();
    for (uint32_t pos = threadIdx.x; pos < total; pos += JP_THREADS) {
      const uint32_t d = st_dst[pos];
      out_keys[d] = st_key[pos];
      out_rows[d] = st_row[pos];
    }
    __syncthreads();
  }
}

struct JPart { DevBuf keys, rows; long long n = 0; };
static int32_t jpartition(pdrs_ctx* c, const JKeyCol& col, long long n, int log_nb, JPart* out) {
  const int nb = 1 << log_nb;
  DevBuf hist;
  PDRS_TRY(hist.alloc(c, (size_t)(3 * nb + 2) * 8, true));
  u64* h = hist.as<u64>();
  int ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 4, (n + JP_TILE - 1) / JP_TILE));   // 4 CTAs/SM: 32 KB staging each
  jpart_hist_kernel<<<ctas, JP_THREADS, 0, c->stream>>>(col, n, log_nb, h);
  jpart_scan_kernel<<<1, 32, 0, c->stream>>>(h, h + nb, h + 2 * nb, nb);
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 8, h + 3 * nb, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  out->n = c->pinned_scalars[8];
  PDRS_TRY(out->keys.alloc(c, (size_t)std::max<long long>(out->n, 1) * 8));
  PDRS_TRY(out->rows.alloc(c, (size_t)std::max<long long>(out->n, 1) * 4));
  const size_t smem = (size_t)JP_TILE * 16;
  static bool attr_set = false;
  if (!attr_set) { PDRS_CUDA(c, cudaFuncSetAttribute(jpart_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; }
  jpart_scatter_kernel<<<ctas, JP_THREADS, smem, c->stream>>>(col, n, log_nb, h + nb, out->keys.as<u64>(), out->rows.as<uint32_t>());
  c->stats.kernel_launches += 3;
  PDRS_CUDA(c, cudaGetLastError());
  return PDRS_OK;
}

struct pdrs_join_result {
  pdrs_ctx* ctx = nullptr;
  int64_t n = 0;
  DevBuf left, right;
};

extern "C" {

int32_t pdrs_join_pairs(pdrs_ctx* c, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how, pdrs_join_result** out) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!left_key || !right_key || !out) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_join_pairs: NULL argument");
  if (how != PDRS_INNER && how != PDRS_LEFT) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "join type %d is not implemented (Inner and Left are)", how);
  if (left_key->dtype != right_key->dtype)   // join.rs:98-104
    return pdrs_fail(c, PDRS_ERR_TYPE_MISMATCH, "join key columns have different types (%d vs %d)", left_key->dtype, right_key->dtype);
  PDRS_CUDA(c, cudaSetDevice(c->device));
  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
  c->stats.main_kernel_ms = 0; c->stats.total_ms = 0;
  ColView lv, rv;
  PDRS_TRY(pdrs_view_col(c, left_key, &lv));
  PDRS_TRY(pdrs_view_col(c, right_key, &rv));
  const int64_t nl = lv.len, nr = rv.len;
  auto* res = new pdrs_join_result();
  res->ctx = c;
  struct Guard { pdrs_join_result* r; ~Guard() { delete r; } } guard{res};

  const long long slots = std::max<long long>(1024, 2 * nr);           // load factor <= 1/2; any even size (see jslot)
  if (slots >= (1ll << 32) - 2) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "join: build side too large (%lld rows)", (long long)nr);
  const size_t table_bytes = (size_t)(slots + 2) * sizeof(JSlot);      // + the reserved slot of the all-ones key
  // Large tables: radix-partition both sides so that one bucket's table region (<= 32 MB) stays in L2.
  // join_algo: 0 auto, 1 = always the direct (left-row-major output) path, 2 = always partitioned.
  int log_nb = 0;
  while (log_nb < 10 && (table_bytes >> log_nb) > (40ull << 20)) log_nb++;      // measured: 25-50 MB regions are L2-resident, fewer buckets partition faster
  if (c->opt_join_log_nb > 0) log_nb = (int)c->opt_join_log_nb;
  bool radix = log_nb > 1 && nl < (1ll << 32) && nr < (1ll << 32);
  if (c->opt_join_algo == 1) radix = false;
  if (c->opt_join_algo == 2 && nl < (1ll << 32) && nr < (1ll << 32)) { radix = true; log_nb = std::max(log_nb, 2); }
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  auto mark = [&](const char* name) {
    if (c->opt_timing < 2) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c->stream); marks.push_back({name, e});
  };
  mark("start");
  DevBuf tab, next, counts, fail;
  PDRS_TRY(tab.alloc(c, table_bytes));
  PDRS_CUDA(c, cudaMemsetAsync(tab.p, 0xFF, table_bytes, c->stream));   // head = -1 (EMPTY)
  PDRS_TRY(next.alloc(c, (size_t)std::max<int64_t>(nr, 1) * 8));
  PDRS_CUDA(c, cudaMemsetAsync(next.p, 0xFF, (size_t)std::max<int64_t>(nr, 1) * 8, c->stream));      // -1 = end of chain
  PDRS_TRY(fail.alloc(c, 32, true));      // [0] failed inserts, [1] build tile counter, [2] probe tile counter
  c->stats.table_slots = slots;
  JKeyCol rc{rv.data, rv.nulls, rv.dtype}, lc{lv.data, lv.nulls, lv.dtype};
  JPart lp, rp;
  JSrc rsrc{rc, nullptr, nullptr}, lsrc{lc, nullptr, nullptr};
  long long nl_eff = nl, nr_eff = nr;
  if (radix) {
    mark("memset");
    PDRS_TRY(jpartition(c, rc, nr, log_nb, &rp));
    mark("partition build side");
    PDRS_TRY(jpartition(c, lc, nl, log_nb, &lp));
    mark("partition probe side");
    rsrc.pkeys = rp.keys.as<u64>(); rsrc.prows = rp.rows.as<uint32_t>(); nr_eff = rp.n;
    lsrc.pkeys = lp.keys.as<u64>(); lsrc.prows = lp.rows.as<uint32_t>(); nl_eff = lp.n;
  }
  if (nr_eff > 0) {
    join_build_kernel<<<(int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 8, (nr_eff + 255) / 256)), 256, 0, c->stream>>>(tab.as<JSlot>(), (u64)slots, next.as<long long>(), rsrc, nr_eff, fail.as<u64>());
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  mark("build");
  const long long ntiles = (nl_eff + JOIN_TILE - 1) / JOIN_TILE;
  const int ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * (c->opt_join_ctas_per_sm > 0 ? c->opt_join_ctas_per_sm : 8), ntiles));
  PDRS_TRY(counts.alloc(c, (size_t)(ntiles + 4) * 8, true));
  DevBuf stash;
  PDRS_TRY(stash.alloc(c, (size_t)std::max<int64_t>(nl_eff, 1) * 8));
  u64* cc = counts.as<u64>();
  if (nl_eff > 0) {
    if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
    join_probe_kernel<<<ctas, JOIN_THREADS, 0, c->stream>>>(tab.as<JSlot>(), (u64)slots, next.as<long long>(), lsrc, nl_eff, how == PDRS_LEFT, stash.as<long long>(), cc, fail.as<u64>() + 2);
    if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
    join_scan_kernel<<<1, 1024, 0, c->stream>>>(cc, (int)ntiles, cc + ntiles + 2);
    c->stats.kernel_launches += 2;
    PDRS_CUDA(c, cudaGetLastError());
  }
  mark("probe+scan");
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, cc + ntiles + 2, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 1, fail.p, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->pinned_scalars[1] != 0) return pdrs_fail(c, PDRS_ERR_CUDA, "join build: hash table insertion failed for %lld rows", (long long)c->pinned_scalars[1]);
  const int64_t M = nl_eff > 0 ? c->pinned_scalars[0] : 0;
  res->n = M;
  PDRS_TRY(res->left.alloc(c, (size_t)std::max<int64_t>(M, 1) * 8));
  PDRS_TRY(res->right.alloc(c, (size_t)std::max<int64_t>(M, 1) * 8));
  if (M > 0) {
    join_write_kernel<<<ctas, JOIN_THREADS, 0, c->stream>>>(stash.as<long long>(), next.as<long long>(), nl_eff, how == PDRS_LEFT, cc, lsrc.prows,
                                                            res->left.as<long long>(), res->right.as<long long>());
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  mark("write");
  c->stats.groupby_algo_used = radix ? 2 : 1;
  if (!marks.empty()) {
    cudaStreamSynchronize(c->stream);
    for (size_t i = 1; i < marks.size(); i++) { float ms = 0; cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second); fprintf(stderr, "[pdrs join] %-22s %8.3f ms\n", marks[i].first, ms); }
    for (auto& m : marks) cudaEventDestroy(m.second);
  }
  if (c->opt_timing) {
    PDRS_CUDA(c, cudaEventRecord(c->ev_t1, c->stream));
    PDRS_CUDA(c, cudaEventSynchronize(c->ev_t1));
    PDRS_CUDA(c, cudaEventElapsedTime(&c->stats.total_ms, c->ev_t0, c->ev_t1));
    if (nl_eff > 0) PDRS_CUDA(c, cudaEventElapsedTime(&c->stats.main_kernel_ms, c->ev_a, c->ev_b));
  } else {
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  guard.r = nullptr;
  *out = res;
  return PDRS_OK;
}

int64_t pdrs_join_len(const pdrs_join_result* r) { return r ? r->n : -1; }
int32_t pdrs_join_indices(const pdrs_join_result* r, int64_t* left_out, int64_t* right_out) {
  if (!r) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = r->ctx;
  if (r->n == 0) return PDRS_OK;
  if (!left_out || !right_out) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_join_indices: NULL output");
  PDRS_CUDA(c, cudaMemcpyAsync(left_out, r->left.p, (size_t)r->n * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaMemcpyAsync(right_out, r->right.p, (size_t)r->n * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}
const int64_t* pdrs_join_left_dev(const pdrs_join_result* r) { return r ? r->left.as<int64_t>() : nullptr; }
const int64_t* pdrs_join_right_dev(const pdrs_join_result* r) { return r ? r->right.as<int64_t>() : nullptr; }
void pdrs_join_result_free(pdrs_join_result* r) {
  if (!r) return;
  cudaSetDevice(r->ctx->device);
  delete r;
}

}  // extern "C"
