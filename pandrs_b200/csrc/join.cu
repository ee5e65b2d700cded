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
//   join_build_kernel   one CAS on the key word claims a slot, the row id is exchanged into the head word; a head that
//                       was already set means the build keys are not unique
//   jdup_*              duplicate build keys only: the rows of every key are collected into one segment
//                       [count, row, row, ...] of a CSR array (count / reserve / fill passes over the build side) and
//                       every segment is sorted ascending ONCE (join.rs:158-161 emits the matches of a key in ascending
//                       right row); the head word then holds the segment's offset.  No chains, no per-probe sorting:
//                       a key with k build rows costs O(k log^2 k) once, not O(k^2) per probe row.
//   join_probe_kernel   one pass over the left keys: stash[i] = head word of the matching slot (or
//                       NOMATCH / NULLKEY) and a per-CTA count of output rows
//   join_write_kernel   per-CTA base offsets from the scan of those counts, block-level exclusive
//                       scans inside each CTA's contiguous row range, ordered writes of the pairs
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <vector>

#include "common.cuh"

typedef unsigned long long u64;

// Structure of arrays: a probe reads one 32-byte sector = the 4 keys of a bucket; the head word is read on a hit only.
struct JTab { u64* keys; uint32_t* heads; u64 slots; uint32_t nbmul; };   // slots is a multiple of 4; index `slots` is the reserved slot of the all-ones key
// nbmul: 1 = one table for all keys.  NB > 1 = the table of ONE radix bucket out of NB (bucket-at-a-time join): the bucket is
// floor(hash * NB / 2^32), so inside a bucket the low 32 bits of hash * NB are uniform again and pick the slot.
// head word: unique build keys: the build row of the key (rows < 2^32 - 1); duplicate build keys: offset of the key's
// segment [count, rows ascending ...] in the CSR array; all ones = none
static constexpr uint32_t JH_EMPTY = 0xFFFFFFFFu;
__device__ __forceinline__ long long jhead_decode(uint32_t h) { return (long long)h; }
static constexpr long long J_NOMATCH = -1, J_NULLKEY = -3;
#define JOIN_THREADS 256
#define JOIN_ITEMS 4

struct JKeyCol { const void* data; const uint8_t* nulls; int dtype; };

// Top 32 bits of (fold(k) * odd constant): the fold brings the high word into the low one, so keys that differ only
// in high bits still spread; the multiply carries every low bit into the top word.  3 integer multiplies.
__device__ __forceinline__ uint32_t jhash32(u64 k) {
  const uint32_t lo = (uint32_t)k ^ (uint32_t)(k >> 32), hi = (uint32_t)(k >> 32);
  constexpr uint32_t CL = 0x7F4A7C15u, CH = 0x9E3779B9u;
  return __umulhi(lo, CL) + lo * CH + hi * CL;
}

// slot = floor(hash_hi32 * slots / 2^32): monotone in the hash, so the rows of one radix bucket (top hash bits)
// fall into one contiguous region of the table; `slots` need not be a power of two
__device__ __forceinline__ u64 jslot(u64 key, u64 slots, uint32_t nbmul = 1u) { return (u64)__umulhi(jhash32(key) * nbmul, (uint32_t)slots); }   // slots < 2^32

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
// Partitioned layouts: flat (bucket after bucket, cap == 0) or padded (bucket b owns positions [b * cap, b * cap + cnt[b]),
// cap a multiple of every tile size, so that a tile never straddles two buckets).
// Multi-GPU exchange path (pdrs_xjoin_*): a bucket is split into 2^log_srcs sub-buckets, one per source rank
// (sub-bucket = bucket << log_srcs | source), cap / cnt describe sub-buckets, prows are row numbers LOCAL to the
// source rank and row0[source] turns them into global row numbers when pairs are emitted.
// Staged exchange (two steps: shuffle by rank, then a local radix partition that mixes the sources): prows of the left
// side are (source rank << row_shift | local row) and are decoded by value.
struct JSrc { JKeyCol col; const u64* pkeys; const uint32_t* prows; long long cap; const u64* cnt; int log_nb; const long long* row0; int log_srcs; int row_shift; };
__device__ __forceinline__ long long jsrc_row_add(const JSrc& s, long long pos) {
  return (s.row0 && !s.row_shift) ? __ldg(s.row0 + ((pos / s.cap) & ((1ll << s.log_srcs) - 1))) : 0ll;
}
__device__ __forceinline__ long long jsrc_left_row(const JSrc& s, uint32_t enc, long long radd) {
  if (s.row_shift) return __ldg(s.row0 + (enc >> s.row_shift)) + (long long)(enc & ((1u << s.row_shift) - 1u));
  return (long long)enc + radd;
}
// end of the valid positions of the tile [lo, lo + tile)
__device__ __forceinline__ long long jsrc_tile_hi(const JSrc& s, long long lo, long long tile, long long n) {
  if (s.cap == 0) return min(n, lo + tile);
  const long long b = lo / s.cap;
  const long long c = min((long long)__ldg(s.cnt + b), s.cap);
  return min(lo + tile, b * s.cap + c);
}
__device__ __forceinline__ bool jsrc_load(const JSrc& s, long long i, u64* key, long long* row) {
  if (s.pkeys) { *key = __ldcs(s.pkeys + i); *row = (long long)__ldcs(s.prows + i); return true; }
  *row = i;
  return jload_key(s.col, i, key);
}

// L2 cache policies: the table region of the current radix bucket must stay in L2 (evict_last) while the key /
// stash / output streams pass through it once (evict_first).
__device__ __forceinline__ u64 l2_policy_evict_last() { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 l2_policy_evict_first() { u64 p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 ld_stream_u64(const u64* a, u64 pol) { u64 v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(pol)); return v; }
__device__ __forceinline__ void st_stream_u64(long long* a, long long v, u64 pol) { asm volatile("st.global.L1::no_allocate.L2::cache_hint.u64 [%0], %1, %2;" :: "l"(a), "l"(v), "l"(pol) : "memory"); }

// Cold table regions: the first probes (or inserts) of a radix bucket miss L2 and would fetch the region from DRAM one
// random 32-byte sector at a time.  Instead every tile streams the matching slice of the NEXT bucket's region (keys
// and head words) into L2 with full-line loads and the evict_last policy; the loaded values are not used.
__device__ __forceinline__ void jprefetch_next_region(const JTab& t, const JSrc& s, long long lo, long long tile, u64 pol, int tid, int nthreads) {
  if (s.cap == 0 || s.log_nb == 0) return;
  const long long b = lo / s.cap, nb = 1ll << s.log_nb;
  if (b + 1 >= nb) return;
  const long long cnt = max(1ll, min((long long)__ldg(s.cnt + b), s.cap));
  const long long p0 = lo - b * s.cap, p1 = min(p0 + tile, cnt);
  if (p0 >= cnt) return;
  const int sh = 32 - s.log_nb;
  const u64 s0 = (u64)__umulhi((uint32_t)((b + 1) << sh), (uint32_t)t.slots) & ~3ull;
  const u64 s1 = b + 2 >= nb ? t.slots : ((u64)__umulhi((uint32_t)((b + 2) << sh), (uint32_t)t.slots) & ~3ull);
  u64 a = s0 + (u64)((double)(s1 - s0) * ((double)p0 / (double)cnt));
  u64 e = s0 + (u64)((double)(s1 - s0) * ((double)p1 / (double)cnt)) + 16;
  a &= ~15ull;                                   // 16 slots = one 128-byte line of keys (and of heads)
  if (e > s1) e = s1;
  // one prefetch per 64 bytes: 16 slots = 128 B of keys + 64 B of heads per lane and step; no destination registers
  for (u64 x = a + 16ull * tid; x < e; x += 16ull * nthreads) {
    asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(t.keys + x));
    asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(t.keys + x + 8));
    asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(t.heads + x));
  }
}

// Insertion: the slot is claimed by a CAS on the key word itself (all-ones = empty), so there is no "being
// published" state, nobody ever waits and no fence is needed (the probe runs in a later kernel).  The row id is
// exchanged into the head word; an exchange that returns another row means the build keys are NOT unique
// (counted in fail[3]: the host then builds the CSR segments below and the head words are rewritten).
// The one key whose bit pattern is all-ones lives in the reserved slot `slots` (never reached by probing).
static constexpr u64 J_EMPTY_KEY = ~0ull;
#define JB_TILE 256    // rows per ticket: ~4700 resident warps x 256 rows keep the window of rows in flight inside one or two radix buckets
#define JB_ITEMS 4
// MODE 0: insert (claim + head exchange).  MODE 1: count the rows of every key in its head word.  MODE 2: fill the CSR
// segments (head word = segment offset, csr[offset] counts the rows placed so far and ends up as the segment length).
template <int MODE>
__global__ void __launch_bounds__(256) join_build_kernel(JTab t, uint32_t* __restrict__ csr, JSrc src, long long n, u64* fail) {
  // Every warp works on its own.  Tiles are handed out in order by a global counter (fail[1]): whatever the relative
  // speed of the warps, the rows in flight form one contiguous window, i.e. they stay inside one or two radix
  // buckets = L2-resident table regions.  A lane keeps JB_ITEMS insertions in flight (claim CAS, then head exchange).
  const long long ntiles = (n + JB_TILE - 1) / JB_TILE;
  const u64 slots = t.slots;
  const u64 pol_keep = l2_policy_evict_last();
  const int lane = threadIdx.x & 31;
  for (;;) {
    long long tile = 0;
    if (lane == 0) tile = (long long)atomicAdd(&fail[1], 1ull);
    tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
    if (tile >= ntiles) break;
    const long long hi = jsrc_tile_hi(src, tile * JB_TILE, JB_TILE, n);
    if (MODE == 0) jprefetch_next_region(t, src, tile * JB_TILE, JB_TILE, pol_keep, lane, 32);
    for (long long i0 = tile * JB_TILE + lane; i0 < hi; i0 += 32 * JB_ITEMS) {
      u64 key[JB_ITEMS], slot[JB_ITEMS], seen[JB_ITEMS];
      long long r[JB_ITEMS];
      uint32_t live = 0;
#pragma unroll
      for (int j = 0; j < JB_ITEMS; j++) {
        const long long i = i0 + 32 * j;
        key[j] = 0; r[j] = 0;
        if (i < hi && jsrc_load(src, i, &key[j], &r[j])) live |= 1u << j;
        slot[j] = key[j] == J_EMPTY_KEY ? slots : (jslot(key[j], slots, t.nbmul) & ~3ull);   // the probe sequence starts at the first slot of the home bucket
      }
      // claim: CAS straight away (the home slot is free for most keys at load factor 1/3); MODE > 0: the key is there, find it
      uint32_t todo = live;
      for (u64 probe = 0; todo && probe <= slots; probe++) {
#pragma unroll
        for (int j = 0; j < JB_ITEMS; j++)
          if ((todo >> j) & 1u) seen[j] = key[j] == J_EMPTY_KEY ? key[j] : (MODE == 0 ? atomicCAS(&t.keys[slot[j]], J_EMPTY_KEY, key[j]) : ld_volatile_u64(&t.keys[slot[j]]));
#pragma unroll
        for (int j = 0; j < JB_ITEMS; j++) {
          if (!((todo >> j) & 1u)) continue;
          if ((MODE == 0 && seen[j] == J_EMPTY_KEY) || seen[j] == key[j]) todo &= ~(1u << j);
          else slot[j] = slot[j] + 1 >= slots ? 0 : slot[j] + 1;
        }
      }
      if (todo) { atomicAdd(fail, (u64)__popc(todo)); live &= ~todo; }
      if (MODE == 0) {
        // the lane whose CAS claimed the slot owns its head word: a plain store (measured: the build is bound by the SMs'
        // atomic issue rate, ~50 G returning atomics / s - one atomic per row instead of two).  A CAS that found the key
        // already there = duplicate build keys: the head words are rebuilt by the CSR passes, nothing to store here.
        uint32_t dups = 0;
#pragma unroll
        for (int j = 0; j < JB_ITEMS; j++) {
          if (!((live >> j) & 1u)) continue;
          if (key[j] == J_EMPTY_KEY) { if (atomicExch(&t.heads[slot[j]], (uint32_t)r[j]) != JH_EMPTY) dups++; }      // reserved slot: no CAS on a key word
          else if (seen[j] == J_EMPTY_KEY) t.heads[slot[j]] = (uint32_t)r[j];
          else dups++;
        }
        if (dups) atomicAdd(&fail[3], (u64)dups);       // duplicate build keys: the single-pass probe does not apply
      } else if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < JB_ITEMS; j++) if ((live >> j) & 1u) atomicAdd(&t.heads[slot[j]], 1u);
      } else {
        uint32_t off[JB_ITEMS], pos[JB_ITEMS];
#pragma unroll
        for (int j = 0; j < JB_ITEMS; j++) if ((live >> j) & 1u) off[j] = t.heads[slot[j]];
#pragma unroll
        for (int j = 0; j < JB_ITEMS; j++) if ((live >> j) & 1u) pos[j] = atomicAdd(&csr[off[j]], 1u);
#pragma unroll
        for (int j = 0; j < JB_ITEMS; j++) if ((live >> j) & 1u) csr[off[j] + 1u + pos[j]] = (uint32_t)r[j];
      }
    }
  }
}

// Duplicate build keys, between the count and fill passes: every occupied slot (head word = number of rows of its key)
// reserves a segment of count + 1 words in the CSR array (one warp-aggregated atomic per 32 slots); the head word
// becomes the segment offset and the first word of the segment (the fill counter, finally the length) is cleared.
__global__ void __launch_bounds__(256) jdup_reserve_kernel(JTab t, uint32_t* __restrict__ csr, u64* __restrict__ cursor) {
  const int lane = threadIdx.x & 31;
  const long long total = (long long)t.slots + 1;           // + the reserved slot of the all-ones key
  for (long long s0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; s0 < total; s0 += (long long)gridDim.x * blockDim.x) {
    const long long s = s0 + lane;
    uint32_t c = 0;
    if (s < total) { c = t.heads[s]; if (s < (long long)t.slots && t.keys[s] == J_EMPTY_KEY) c = 0; }
    if (s == (long long)t.slots && c == JH_EMPTY) c = 0;   // (not reached: the count pass runs on zeroed head words)
    uint32_t need = c ? c + 1u : 0u, incl = need;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
    const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
    u64 base = 0;
    if (lane == 0 && tot) base = atomicAdd(cursor, (u64)tot);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (s < total) {
      if (c) { const uint32_t off = (uint32_t)(base + incl - need); t.heads[s] = off; csr[off] = 0u; }
      else t.heads[s] = JH_EMPTY;
    }
  }
}

// Sort every segment ascending (join.rs:158-161: the matches of a key come out in ascending right row).  Segments of up to
// 32 rows: one thread, insertion sort.  Longer ones are appended to a work list for jdup_sort_long_kernel.
__global__ void __launch_bounds__(256) jdup_sort_short_kernel(JTab t, uint32_t* __restrict__ csr, u64* __restrict__ cursor /* [1] long segments, [2] padded length of the longest */,
                                                              uint2* __restrict__ longs, long long longs_cap) {
  const long long total = (long long)t.slots + 1;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (long long)gridDim.x * blockDim.x) {
    const uint32_t off = t.heads[s];
    if (off == JH_EMPTY) continue;
    const uint32_t c = csr[off];
    if (c <= 1) continue;
    uint32_t* a = csr + off + 1;
    if (c <= 32) {
      for (uint32_t i = 1; i < c; i++) {
        const uint32_t x = a[i];
        uint32_t q = i;
        while (q > 0 && a[q - 1] > x) { a[q] = a[q - 1]; q--; }
        a[q] = x;
      }
    } else {
      const u64 at = atomicAdd(&cursor[1], 1ull);
      if ((long long)at < longs_cap) longs[at] = make_uint2(off + 1, c);
      uint32_t m = 64;
      while (m < c) m <<= 1;
      atomicMax(&cursor[2], (u64)m);
    }
  }
}
// One CTA per long segment: bitonic sort of the segment padded with all-ones to a power of two - in shared memory up to
// 8192 rows, else in this CTA's slice of `scratch` (global memory, L2-resident for all but absurd segment lengths).
#define JDS_NT 512
#define JDS_SMEM_ROWS 8192
__global__ void __launch_bounds__(JDS_NT) jdup_sort_long_kernel(uint32_t* __restrict__ csr, const uint2* __restrict__ longs, long long nlong, uint32_t* __restrict__ scratch, u64 scratch_stride) {
  __shared__ uint32_t sh[JDS_SMEM_ROWS];
  for (long long w = blockIdx.x; w < nlong; w += gridDim.x) {
    const uint2 seg = longs[w];
    uint32_t* a = csr + seg.x;
    const uint32_t c = seg.y;
    uint32_t m = 64;
    while (m < c) m <<= 1;
    uint32_t* b = m <= JDS_SMEM_ROWS ? sh : scratch + (u64)blockIdx.x * scratch_stride;
    for (uint32_t i = threadIdx.x; i < m; i += JDS_NT) b[i] = i < c ? a[i] : 0xFFFFFFFFu;
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t i = threadIdx.x; i < m; i += JDS_NT) {
          const uint32_t l = i ^ j;
          if (l > i) {
            const uint32_t x = b[i], y = b[l];
            if ((x > y) == ((i & k) == 0)) { b[i] = y; b[l] = x; }
          }
        }
        __syncthreads();
      }
    }
    for (uint32_t i = threadIdx.x; i < c; i += JDS_NT) a[i] = b[i];
    __syncthreads();
  }
}

__device__ __forceinline__ ulonglong4 ld_bucket(const u64* p, u64 pol) {
  ulonglong4 s;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0, %1, %2, %3}, [%4], %5;" : "=l"(s.x), "=l"(s.y), "=l"(s.z), "=l"(s.w) : "l"(p), "l"(pol));
  return s;
}
// index (0..3) of `key` in a bucket, 4 = the bucket ends the probe sequence, 5 = go on with the next bucket.
// Insertion always takes the first empty slot of the sequence, so the occupied slots of a bucket form a prefix:
// the sequence ends inside this bucket iff its last slot is empty.
__device__ __forceinline__ int jbucket_find(const ulonglong4& b, u64 key) {
  int f = b.w == J_EMPTY_KEY ? 4 : 5;
  f = b.w == key ? 3 : f;
  f = b.z == key ? 2 : f;
  f = b.y == key ? 1 : f;
  f = b.x == key ? 0 : f;
  return f;
}
// Probing reads one 32-byte sector (the four keys of a bucket) per step.  Returns the head word of the matching
// slot or J_NOMATCH.
__device__ __forceinline__ long long jprobe_from(const JTab& t, uint32_t bkt, u64 key, u64 pol) {
  const uint32_t nslots = (uint32_t)t.slots;
  for (uint32_t step = 0; step <= nslots; step += 4) {
    const int f = jbucket_find(ld_bucket(t.keys + bkt, pol), key);
    if (f < 4) return jhead_decode(__ldg(t.heads + bkt + f));
    if (f == 4) return J_NOMATCH;
    bkt = bkt + 4 >= nslots ? 0 : bkt + 4;
  }
  return J_NOMATCH;
}
__device__ __forceinline__ long long jprobe(const JTab& t, u64 key, u64 pol) {
  if (key == J_EMPTY_KEY) { const uint32_t h = __ldg(&t.heads[t.slots]); return h == JH_EMPTY ? J_NOMATCH : jhead_decode(h); }
  return jprobe_from(t, (uint32_t)jslot(key, t.slots, t.nbmul) & ~3u, key, pol);
}

// pairs a probe row emits; csr != NULL: the build keys are not unique and the stash holds a segment offset
__device__ __forceinline__ long long jcount(long long stash, const uint32_t* __restrict__ csr, int left_join) {
  if (stash == J_NULLKEY) return 0;
  if (stash == J_NOMATCH) return left_join ? 1 : 0;
  return csr ? (long long)__ldg(csr + stash) : 1ll;
}

// Rows are processed in tiles of JOIN_TILE consecutive positions, tile t by CTA t mod grid: the output stays in
// position order (tile offsets come from a scan of the per-tile counts) and, on the partitioned path, all CTAs
// work inside the same radix bucket at any time (L2-resident table region).
#define JOIN_TILE (JOIN_THREADS * JOIN_ITEMS * 4)

__global__ void __launch_bounds__(JOIN_THREADS) join_probe_kernel(JTab tab, const uint32_t* __restrict__ csr, JSrc src, long long n,
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
    const long long lo = tile * JOIN_TILE, hi = jsrc_tile_hi(src, lo, JOIN_TILE, n);
    jprefetch_next_region(tab, src, lo, JOIN_TILE, pol_keep, threadIdx.x, blockDim.x);
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
      for (int j = 0; j < JOIN_ITEMS; j++) st[j] = ok[j] ? jprobe(tab, key[j], pol_keep) : J_NULLKEY;
#pragma unroll
      for (int j = 0; j < JOIN_ITEMS; j++) {
        long long i = i0 + (long long)j * JOIN_THREADS;
        if (i < hi) { st_stream_u64(stash + i, st[j], pol_stream); cnt += (u64)jcount(st[j], csr, left_join); }
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

__global__ void __launch_bounds__(JOIN_THREADS) join_write_kernel(const long long* __restrict__ stash, const uint32_t* __restrict__ csr, long long n, int left_join,
                                                                  const u64* __restrict__ tile_offsets, JSrc src,
                                                                  long long* __restrict__ out_l, long long* __restrict__ out_r) {
  const uint32_t* __restrict__ prows = src.prows;
  __shared__ u64 wsum[JOIN_THREADS / 32];
  __shared__ u64 sh_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ntiles = (n + JOIN_TILE - 1) / JOIN_TILE;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long lo = tile * JOIN_TILE, hi = jsrc_tile_hi(src, lo, JOIN_TILE, n);
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
        cn[j] = (u64)jcount(st[j], csr, left_join);
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
        if (prows) i = jsrc_left_row(src, __ldg(prows + i), jsrc_row_add(src, i));      // partitioned probe side: position -> original left row
        if (st[j] == J_NOMATCH) { out_l[pos] = i; out_r[pos] = -1; pos++; }
        else if (!csr) { out_l[pos] = i; out_r[pos] = st[j]; pos++; }
        else {                                    // the key's segment: its build rows in ascending order (join.rs:158-161)
          const uint32_t* __restrict__ seg = csr + st[j] + 1;
          for (u64 q = 0; q < cn[j]; q++) { out_l[pos] = i; out_r[pos] = (long long)__ldg(seg + q); pos++; }
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

__device__ __forceinline__ uint32_t jbucket(u64 key, int log_nb) { return log_nb ? jhash32(key) >> (32 - log_nb) : 0u; }
// Destination rank of a key on the multi-GPU exchange path: a second, independent multiplicative hash, so that the
// rows a rank receives still spread over the whole range of jhash32 (= over all of its radix buckets and table slots).
__device__ __forceinline__ uint32_t jhash_rank(u64 k) { return (uint32_t)(((k ^ (k >> 29)) * 0xD6E8FEB86659FD93ull) >> 32); }
// Where the one-pass partition writes: one (keys, rows) buffer per destination rank.  Single GPU: world = 1, the
// buffers are local.  Exchange path: keys[r] / rows[r] are rank r's receive buffers mapped into this process (CUDA
// IPC), so the partition pass IS the shuffle - every bucket run is stored straight through NVLink; combined bucket
// = rank << log_nb | radix bucket; inside rank r's buffers this rank `me` owns sub-bucket (bucket << log_world | me).
struct JXDst { u64* keys[8]; uint32_t* rows[8]; int log_world; int me; uint32_t row_add; };
// Input of the local partition on the staged exchange path: (key, row) records received from the peers, one padded
// region of `cap` positions per source rank holding cnt[source] rows.
struct JStaged { const u64* keys; const uint32_t* rows; const u64* cnt; long long cap; };

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

// ---------------------------------------------------------------- one-pass radix partition into padded buckets
// No histogram pass: bucket b owns the fixed range [b * cap, (b + 1) * cap) of the output (cap = expected rows per
// bucket + slack; the hash spreads distinct keys evenly, heavy duplicates can overflow -> the caller falls back to
// the exact two-pass partition above).  Per tile of 8192 rows: shared-memory histogram whose atomics return the
// row's rank, one global reservation per bucket, rows staged in shared memory in bucket order, then every bucket's
// run (8192 / nb rows) is written out contiguously.  2 CTAs per SM overlap each other's barriers.
#define JQ_NT 512
#define JQ_ITEMS 8
#define JQ_TILE (JQ_NT * JQ_ITEMS)
#define JQ_SMEM ((size_t)JQ_TILE * 16 + 1056 * 4 + 1024 * 8 + JQ_TILE)
template <bool XCHG, bool STAGED = false>
__global__ void __launch_bounds__(JQ_NT, 2) jpart1_kernel(JKeyCol col, long long n, int log_nb, long long cap, u64* __restrict__ cursor,
                                                          const JXDst x, u64* __restrict__ overflow, const JStaged sin = JStaged{}) {
  u64* __restrict__ out_keys = x.keys[0];
  uint32_t* __restrict__ out_rows = x.rows[0];
  extern __shared__ __align__(16) unsigned char jsm[];
  u64* st_key = reinterpret_cast<u64*>(jsm);                                 // [JQ_TILE] staged keys
  uint32_t* st_row = reinterpret_cast<uint32_t*>(st_key + JQ_TILE);         // [JQ_TILE] staged row ids
  uint32_t* st_dst = st_row + JQ_TILE;                                      // [JQ_TILE] output position of the staged row
  uint32_t* H = st_dst + JQ_TILE;                                           // [1024 + 32] bucket counts of the tile
  uint2* HD = reinterpret_cast<uint2*>(H + 1056);                           // [1024] {offset in the staging area, output position of the first row}
  uint8_t* st_rk = reinterpret_cast<uint8_t*>(HD + 1024);                   // [JQ_TILE] destination rank of the staged row (exchange path)
  const int lw = XCHG ? x.log_world : 0;
  const uint32_t nbmask = (1u << log_nb) - 1u;
  __shared__ uint32_t wsum[JQ_NT / 32];
  __shared__ uint32_t sh_total;
  const int nb = 1 << log_nb;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool fast = !STAGED && col.dtype == PDRS_I64 && !col.nulls;
  const long long tstride = (long long)gridDim.x * JQ_TILE;
  uint32_t srow[STAGED ? JQ_ITEMS : 1];
  long long t0 = (long long)blockIdx.x * JQ_TILE;
  u64 key[JQ_ITEMS];
  // plain Int64 keys: the loads of a tile are issued one stage ahead (in flight while the previous tile is written out)
  auto load_tile = [&](long long tb) {
#pragma unroll
    for (int j = 0; j < JQ_ITEMS; j++) { const long long i = tb + (long long)j * JQ_NT + tid; key[j] = i < n ? (u64)__ldcs((const long long*)col.data + i) : 0ull; }
  };
  if (fast && t0 < n) load_tile(t0);
  for (; t0 < n; t0 += tstride) {
    for (int i = tid; i < 1056; i += JQ_NT) H[i] = 0;
    __syncthreads();
    uint32_t br[JQ_ITEMS];            // bucket << 16 | rank inside the bucket (tile <= 4096 rows); all ones = no row
#pragma unroll
    for (int j = 0; j < JQ_ITEMS; j++) {
      const long long i = t0 + (long long)j * JQ_NT + tid;
      br[j] = 0xFFFFFFFFu;
      bool live = i < n;
      if (STAGED) {      // a tile never straddles two source regions (cap is a multiple of the tile)
        const long long r = i / sin.cap;
        live = live && (i - r * sin.cap) < (long long)__ldg(sin.cnt + r);
        if (live) { key[j] = __ldcs(sin.keys + i); srow[j] = __ldcs(sin.rows + i); }
      } else if (!fast) live = live && jload_key(col, i, &key[j]);
      if (live) {
        uint32_t b = jbucket(key[j], log_nb);
        if (XCHG && lw) b |= (jhash_rank(key[j]) >> (32 - lw)) << log_nb;
        br[j] = (b << 16) | atomicAdd(&H[b], 1u);
      }
    }
    __syncthreads();
    // first output position of combined bucket cb (see JXDst); single GPU: cb * cap
    auto obase = [&](uint32_t cb) -> u64 { return XCHG ? (u64)((((cb & nbmask) << lw) | (uint32_t)x.me)) * (u64)cap : (u64)cb * (u64)cap; };
    {   // exclusive scan of the bucket counts (thread t owns buckets 2t, 2t + 1) + one global reservation per bucket
      const uint32_t c0 = H[2 * tid], c1 = H[2 * tid + 1], c = c0 + c1;
      uint32_t incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
      if (lane == 31) wsum[warp] = incl;
      uint32_t g0 = 0xFFFFFFFFu, g1 = 0xFFFFFFFFu;
      if (c0) { const u64 at = atomicAdd(&cursor[2 * tid], (u64)c0); if (at + c0 > (u64)cap) atomicAdd(overflow, 1ull); else g0 = (uint32_t)(obase(2 * tid) + at); }
      if (c1) { const u64 at = atomicAdd(&cursor[2 * tid + 1], (u64)c1); if (at + c1 > (u64)cap) atomicAdd(overflow, 1ull); else g1 = (uint32_t)(obase(2 * tid + 1) + at); }
      __syncthreads();
      uint32_t ws = lane < JQ_NT / 32 ? wsum[lane] : 0u;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, ws, d); if (lane >= d) ws += o; }
      const uint32_t wprefix = __shfl_sync(0xFFFFFFFFu, ws, (warp + 31) & 31);
      const uint32_t excl = (warp ? wprefix : 0u) + incl - c;
      HD[2 * tid] = make_uint2(excl, g0);
      HD[2 * tid + 1] = make_uint2(excl + c0, g1);
      if (tid == JQ_NT - 1) sh_total = excl + c;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < JQ_ITEMS; j++) {
      if (br[j] == 0xFFFFFFFFu) continue;
      const uint2 hd = HD[br[j] >> 16];
      const uint32_t rank = br[j] & 0xFFFFu, pos = hd.x + rank;
      st_key[pos] = key[j];
      st_row[pos] = STAGED ? srow[j] : (uint32_t)(t0 + (long long)j * JQ_NT + tid);
      st_dst[pos] = hd.y == 0xFFFFFFFFu ? 0xFFFFFFFFu : hd.y + rank;       // all ones: the bucket overflowed, the caller discards this partitioning
      if (XCHG) st_rk[pos] = (uint8_t)(br[j] >> (16 + log_nb));
    }
    if (fast && t0 + tstride < n) load_tile(t0 + tstride);
    __syncthreads();
    const uint32_t total = sh_total;
    if (XCHG && log_nb == 0) {
      // Staged exchange: one long run per destination rank.  Peer stores move whole 128-byte lines, so the rows are
      // handed to the lanes by DESTINATION position: lane l writes positions = l (mod 32), every warp store covers
      // aligned lines (2 for the keys, 1 for the row ids) and no line is sent twice.
      const int R = 1 << lw;
      for (int r = 0; r < R; r++) {
        const uint2 hd = HD[r];
        const uint32_t len = (r + 1 < R ? HD[r + 1].x : total) - hd.x;
        if (hd.y == 0xFFFFFFFFu || len == 0) continue;
        const uint32_t mis = hd.y & 31u;
        u64* __restrict__ dk = x.keys[r] + (hd.y - mis);
        uint32_t* __restrict__ dr = x.rows[r] + (hd.y - mis);
        for (uint32_t q = tid; q < len + mis; q += JQ_NT) {
          if (q < mis) continue;
          dk[q] = st_key[hd.x + q - mis];
          dr[q] = st_row[hd.x + q - mis] + x.row_add;
        }
      }
    } else
    for (uint32_t pos = tid; pos < total; pos += JQ_NT) {     // consecutive staged rows of a bucket go to consecutive output rows
      const uint32_t d = st_dst[pos];
      if (d == 0xFFFFFFFFu) continue;
      if (XCHG) {     // consecutive staged rows of a combined bucket form one contiguous run in the destination rank's buffer (NVLink stores)
        const int r = st_rk[pos];
        x.keys[r][d] = st_key[pos];
        x.rows[r][d] = st_row[pos] + x.row_add;
      } else {
        out_keys[d] = st_key[pos];
        out_rows[d] = st_row[pos];
      }
    }
    __syncthreads();
  }
}

// Exchange path: after the partition pass every rank tells its peers how many rows it stored into each of its
// sub-buckets (cursor[cb] of this rank -> cnt[(bucket << log_world) | me] of rank cb >> log_nb).
struct JXCnt { u64* cnt[8]; };
__global__ void jx_publish_counts_kernel(const u64* __restrict__ cursor, const JXCnt dst, int log_nb, int log_world, int me, long long cap) {
  const int cb = blockIdx.x * blockDim.x + threadIdx.x;
  if (cb >= (1 << (log_nb + log_world))) return;
  const int r = cb >> log_nb, b = cb & ((1 << log_nb) - 1);
  const u64 c = cursor[cb];
  dst.cnt[r][(b << log_world) | me] = c < (u64)cap ? c : (u64)cap;
}

// ---------------------------------------------------------------- single-pass probe + emit (unique build keys)
// With at most one match per probe row the output position of a row is a prefix count of match flags: ballots
// inside the warp, a scan of 8 warp totals, and one atomicAdd per tile on the global output cursor.  No stash, no
// second pass over the probe side.  Pairs come out in tile order of arrival (the ABI leaves the order open; parity
// is checked after the canonical sort).  out_l / out_r must hold one entry per probe row.
#define JE_THREADS 256
#define JE_HALF 4                            // probes a thread keeps in flight
#define JE_ITEMS 8                           // rows per lane per compaction step
#define JE_UNIT (32 * JE_ITEMS)              // rows per warp per compaction step
#define JE_TILE (JE_UNIT * 8)                // rows per tile ticket (one warp)
__device__ __forceinline__ uint32_t ld_keep_u32(const uint32_t* a, u64 pol) { uint32_t v; asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol)); return v; }
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* a, u64 pol) { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol)); return v; }
// Every WARP works on its own: tiles of 2048 positions are handed out in order by a global counter (so the probes in
// flight stay inside one or two radix buckets = L2-resident table regions), and there is no block-level barrier on
// the path - a warp's chain of dependent memory round trips (keys -> bucket -> head word -> output reservation)
// overlaps with the chains of the 23 other warps of the SM.
__global__ void __launch_bounds__(JE_THREADS, 3) join_probe_emit_kernel(JTab tab, JSrc src, long long n, int left_join,
                                                                      long long* __restrict__ out_l, long long* __restrict__ out_r,
                                                                      u64* __restrict__ out_cursor, u64* __restrict__ tile_ctr, int tile_units = 8) {
  constexpr uint32_t NONE = 0xFFFFFFFFu;
  const long long tile_rows = (long long)JE_UNIT * tile_units;      // rows per ticket
  const long long ntiles = (n + tile_rows - 1) / tile_rows;
  const u64 pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  const int lane = threadIdx.x & 31;
  const uint32_t nslots = (uint32_t)tab.slots, nbmul = tab.nbmul;
  for (;;) {
    long long tile = 0;
    if (lane == 0) tile = (long long)atomicAdd(tile_ctr, 1ull);
    tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
    if (tile >= ntiles) break;
    const long long tlo = tile * tile_rows, hi = jsrc_tile_hi(src, tlo, tile_rows, n);
    jprefetch_next_region(tab, src, tlo, tile_rows, pol_keep, lane, 32);
#pragma unroll 1
    for (long long lo = tlo; lo < hi; lo += JE_UNIT) {
      uint32_t res[JE_ITEMS];             // matching right row (build rows < 2^32 on this path), NONE = no match
      uint32_t lrow[JE_ITEMS];
      uint32_t okmask = 0;
      if (src.pkeys && lo + JE_UNIT <= hi) {
        // ---- fast path: a full unit of partitioned rows.  All 16 stream loads first, then the bucket loads of 4 rows
        //      at a time, then their head words; the all-ones key (= the empty marker) takes the general path.
        u64 key[JE_ITEMS];
        const u64* kp = src.pkeys + lo + lane;
        const uint32_t* rp = src.prows + lo + lane;
#pragma unroll
        for (int j = 0; j < JE_ITEMS; j++) { key[j] = ld_stream_u64(kp + j * 32, pol_stream); lrow[j] = ld_stream_u32(rp + j * 32, pol_stream); }
        okmask = (1u << JE_ITEMS) - 1u;
        uint32_t pend = 0, special = 0;       // rows whose probe sequence goes on past the home bucket / all-ones keys
#pragma unroll
        for (int half = 0; half < JE_ITEMS / JE_HALF; half++) {
          ulonglong4 bk[JE_HALF];
          uint32_t home[JE_HALF];
#pragma unroll
          for (int jj = 0; jj < JE_HALF; jj++) {
            home[jj] = __umulhi(jhash32(key[half * JE_HALF + jj]) * nbmul, nslots) & ~3u;
            bk[jj] = ld_bucket(tab.keys + home[jj], pol_keep);
          }
#pragma unroll
          for (int jj = 0; jj < JE_HALF; jj++) {
            const int j = half * JE_HALF + jj;
            const int f = jbucket_find(bk[jj], key[j]);
            res[j] = NONE;
            if (f < 4) res[j] = ld_keep_u32(tab.heads + home[jj] + f, pol_keep);   // the row (unique keys: no flag bit)
            if (f == 5) pend |= 1u << j;
            if (key[j] == J_EMPTY_KEY) special |= 1u << j;
          }
        }
        pend &= ~special;
        // longer probe sequences (a few % of the rows): every lane resolves its first two pending rows per round, so a
        // round costs one dependent round trip for the whole warp, not one per row
        uint32_t steps = 0;                    // 4 bits per row: buckets examined beyond the home bucket
        while (__any_sync(0xFFFFFFFFu, pend != 0)) {
          const int j0 = pend ? __ffs(pend) - 1 : -1;
          const uint32_t p2 = pend & (pend - 1u);
          const int j1 = p2 ? __ffs(p2) - 1 : -1;
          u64 k0 = 0, k1 = 0;
#pragma unroll
          for (int j = 0; j < JE_ITEMS; j++) { if (j == j0) k0 = key[j]; if (j == j1) k1 = key[j]; }
          uint32_t b0 = 0, b1 = 0;
          ulonglong4 q0 = make_ulonglong4(0, 0, 0, J_EMPTY_KEY), q1 = q0;
          if (j0 >= 0) {
            const uint32_t st = ((steps >> (4 * j0)) & 15u) + 1u;
            b0 = (uint32_t)(((u64)(__umulhi(jhash32(k0) * nbmul, nslots) & ~3u) + 4ull * st) % nslots);
            q0 = ld_bucket(tab.keys + b0, pol_keep);
          }
          if (j1 >= 0) {
            const uint32_t st = ((steps >> (4 * j1)) & 15u) + 1u;
            b1 = (uint32_t)(((u64)(__umulhi(jhash32(k1) * nbmul, nslots) & ~3u) + 4ull * st) % nslots);
            q1 = ld_bucket(tab.keys + b1, pol_keep);
          }
          const int f0 = j0 >= 0 ? jbucket_find(q0, k0) : 4, f1 = j1 >= 0 ? jbucket_find(q1, k1) : 4;
          uint32_t r0 = NONE, r1 = NONE;
          if (f0 < 4) r0 = ld_keep_u32(tab.heads + b0 + f0, pol_keep);
          if (f1 < 4) r1 = ld_keep_u32(tab.heads + b1 + f1, pol_keep);
#pragma unroll
          for (int j = 0; j < JE_ITEMS; j++) { if (j == j0 && f0 < 4) res[j] = r0; if (j == j1 && f1 < 4) res[j] = r1; }
          if (j0 >= 0) { if (f0 != 5) pend &= ~(1u << j0); else if (((steps >> (4 * j0)) & 15u) == 14u) { pend &= ~(1u << j0); special |= 1u << j0; } else steps += 1u << (4 * j0); }
          if (j1 >= 0) { if (f1 != 5) pend &= ~(1u << j1); else if (((steps >> (4 * j1)) & 15u) == 14u) { pend &= ~(1u << j1); special |= 1u << j1; } else steps += 1u << (4 * j1); }
        }
        if (special) {                         // all-ones keys, probe sequences longer than 16 buckets: the general routine
#pragma unroll
          for (int j = 0; j < JE_ITEMS; j++) if ((special >> j) & 1u) { const long long h = jprobe(tab, key[j], pol_keep); res[j] = h == J_NOMATCH ? NONE : (uint32_t)h; }
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < JE_ITEMS; j++) {
          const long long i = lo + (long long)j * 32 + lane;
          uint32_t r = NONE, lr = (uint32_t)i;
          if (i < hi) {
            u64 key;
            bool ok;
            if (src.pkeys) { key = __ldcs(src.pkeys + i); lr = __ldcs(src.prows + i); ok = true; }
            else ok = jload_key(src.col, i, &key);
            if (ok) { okmask |= 1u << j; const long long h = jprobe(tab, key, pol_keep); r = h == J_NOMATCH ? NONE : (uint32_t)h; }
          }
#pragma unroll
          for (int jj = 0; jj < JE_ITEMS; jj++) if (jj == j) { res[jj] = r; lrow[jj] = lr; }
        }
      }
      // compaction inside the warp: slab j = the 32 rows of item j; one output reservation per unit
      uint32_t emit = 0, wcnt = 0;
      uint32_t woff[JE_ITEMS];
#pragma unroll
      for (int j = 0; j < JE_ITEMS; j++) {
        const bool e = ((okmask >> j) & 1u) && (left_join || res[j] != NONE);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, e);
        if (e) emit |= 1u << j;
        woff[j] = wcnt + __popc(m & ((1u << lane) - 1u));
        wcnt += __popc(m);
      }
      u64 base = 0;
      if (lane == 0 && wcnt) base = atomicAdd(out_cursor, (u64)wcnt);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      long long* ol = out_l + base;
      long long* orr = out_r + base;
      const long long radd = jsrc_row_add(src, lo);     // a unit never straddles two sub-buckets
#pragma unroll
      for (int j = 0; j < JE_ITEMS; j++) {
        if (!((emit >> j) & 1u)) continue;
        st_stream_u64(ol + woff[j], jsrc_left_row(src, lrow[j], radd), pol_stream);
        st_stream_u64(orr + woff[j], res[j] == NONE ? -1ll : (long long)res[j], pol_stream);
      }
    }
  }
}

// ---------------------------------------------------------------- Right / Outer: unmatched rows of the right frame
// join.rs:211-224: after the pairs of the left rows, every right row that no pair references is appended as
// (None, r) in ascending r - including right rows whose key is NULL (they never match).
__global__ void jmark_matched_kernel(const long long* __restrict__ out_r, long long m, uint8_t* __restrict__ matched) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (long long)gridDim.x * blockDim.x) {
    const long long r = out_r[j];
    if (r >= 0) matched[r] = 1;
  }
}
#define JU_TILE 4096
__global__ void __launch_bounds__(256) junmatched_count_kernel(const uint8_t* __restrict__ matched, long long nr, u64* __restrict__ tile_counts) {
  __shared__ uint32_t sh;
  for (long long tile = blockIdx.x; tile * JU_TILE < nr; tile += gridDim.x) {
    if (threadIdx.x == 0) sh = 0;
    __syncthreads();
    uint32_t c = 0;
    for (long long i = tile * JU_TILE + threadIdx.x; i < min(nr, (tile + 1) * JU_TILE); i += 256) c += matched[i] ? 0u : 1u;
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sh, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_counts[tile] = sh;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) junmatched_write_kernel(const uint8_t* __restrict__ matched, long long nr, const u64* __restrict__ tile_offsets,
                                                               long long* __restrict__ out_l, long long* __restrict__ out_r) {
  __shared__ uint32_t wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long tile = blockIdx.x; tile * JU_TILE < nr; tile += gridDim.x) {
    // thread t owns 16 consecutive rows of the tile: positions stay ascending in r
    const long long r0 = tile * JU_TILE + (long long)threadIdx.x * 16;
    uint32_t mask = 0;
    for (int k = 0; k < 16; k++) if (r0 + k < nr && !matched[r0 + k]) mask |= 1u << k;
    const uint32_t mine = __popc(mask);
    uint32_t incl = mine;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < warp; w++) wbase += wsum[w];
    u64 pos = tile_offsets[tile] + wbase + incl - mine;
    for (int k = 0; k < 16; k++) if ((mask >> k) & 1u) { out_l[pos] = -1; out_r[pos] = r0 + k; pos++; }
    __syncthreads();
  }
}

// ================================================================ bucket-at-a-time join (unique build keys, hash-uniform keys)
// The radix join above keeps ONE table for all buckets (3.6 GB for 1e8 build rows): it has to be cleared with a memset,
// every region is fetched from DRAM when its bucket starts and written back when it ends.  Here both sides are
// partitioned into NB buckets as before, but then the buckets are joined ONE AT A TIME in a table region that is reused
// for every bucket: clear the region (24 MB: stays in L2), insert the bucket's build rows, probe with the bucket's probe
// rows, next bucket.  The table never touches DRAM.  Two streams with one region each alternate, so that the clear +
// build of bucket b + 1 (latency bound, small) fills the tail of the probe of bucket b.
//
// Payload mode (pdrs_join_gather): up to two 8-byte columns of the BUILD side travel with their rows - through the
// partition (a second / third round of the same tile through the same staging buffer, NULL -> 0) and into 32-byte table
// slots {key, row, p0, p1} = one L2 sector: a hit returns the right row AND its payload values in the sector the key
// comparison needed anyway.  That replaces the per-column gathers of join.rs:290-552 (random 8-byte DRAM reads over the
// whole right column, once per PAIR) by sequential traffic (once per BUILD ROW) plus nothing.
#define JP_MAXPAY 2
struct JPay { const u64* src[JP_MAXPAY]; const uint8_t* nulls[JP_MAXPAY]; u64* dst[JP_MAXPAY]; int n; };

__device__ __forceinline__ uint32_t jbucket_nb(u64 key, uint32_t nb) { return __umulhi(jhash32(key), nb); }

// One-pass partition into `nb` (any number <= 1024) padded buckets of (key, row) [+ payload columns]: tiles of 8192 rows
// (1024 threads x 8), bucket histogram in shared memory (the atomic's return value ranks the row inside its bucket), one
// global reservation per bucket and tile, rows staged in bucket order, every bucket's run written out contiguously.
#define JR_NT 1024
#define JR_ITEMS 8
#define JR_TILE (JR_NT * JR_ITEMS)
#define JR_SMEM ((size_t)JR_TILE * 16 + 1056 * 4 + 1024 * 8)
__global__ void __launch_bounds__(JR_NT, 1) jpartp_kernel(JKeyCol col, long long n, uint32_t nb, long long cap, u64* __restrict__ cursor,
                                                          u64* __restrict__ out_keys, uint32_t* __restrict__ out_rows, u64* __restrict__ overflow, const JPay pay) {
  extern __shared__ __align__(16) unsigned char jsm[];
  u64* st_key = reinterpret_cast<u64*>(jsm);                                 // [JR_TILE] staged keys (then: staged payload values)
  uint32_t* st_row = reinterpret_cast<uint32_t*>(st_key + JR_TILE);         // [JR_TILE] staged row ids
  uint32_t* st_dst = st_row + JR_TILE;                                      // [JR_TILE] output position of the staged row
  uint32_t* H = st_dst + JR_TILE;                                           // [1024 + 32] bucket counts of the tile
  uint2* HD = reinterpret_cast<uint2*>(H + 1056);                           // [1024] {offset in the staging area, output position of the first row}
  __shared__ uint32_t wsum[JR_NT / 32];
  __shared__ uint32_t sh_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool fast = col.dtype == PDRS_I64 && !col.nulls;
  const long long tstride = (long long)gridDim.x * JR_TILE;
  long long t0 = (long long)blockIdx.x * JR_TILE;
  u64 key[JR_ITEMS];
  auto load_tile = [&](long long tb) {
#pragma unroll
    for (int j = 0; j < JR_ITEMS; j++) { const long long i = tb + (long long)j * JR_NT + tid; key[j] = i < n ? (u64)__ldcs((const long long*)col.data + i) : 0ull; }
  };
  if (fast && t0 < n) load_tile(t0);
  for (; t0 < n; t0 += tstride) {
    for (int i = tid; i < 1056; i += JR_NT) H[i] = 0;
    __syncthreads();
    uint32_t br[JR_ITEMS];            // bucket << 16 | rank inside the bucket (tile <= 8192 rows); all ones = no row
#pragma unroll
    for (int j = 0; j < JR_ITEMS; j++) {
      const long long i = t0 + (long long)j * JR_NT + tid;
      br[j] = 0xFFFFFFFFu;
      bool live = i < n;
      if (!fast) live = live && jload_key(col, i, &key[j]);
      if (live) { const uint32_t b = jbucket_nb(key[j], nb); br[j] = (b << 16) | atomicAdd(&H[b], 1u); }
    }
    __syncthreads();
    {   // exclusive scan of the bucket counts (thread t owns bucket t) + one global reservation per bucket
      const uint32_t c = H[tid];
      uint32_t incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
      if (lane == 31) wsum[warp] = incl;
      uint32_t g = 0xFFFFFFFFu;
      if (c) { const u64 at = atomicAdd(&cursor[tid], (u64)c); if (at + c > (u64)cap) atomicAdd(overflow, 1ull); else g = (uint32_t)((u64)tid * (u64)cap + at); }
      __syncthreads();
      uint32_t ws = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, ws, d); if (lane >= d) ws += o; }
      const uint32_t wprefix = __shfl_sync(0xFFFFFFFFu, ws, (warp + 31) & 31);
      const uint32_t excl = (warp ? wprefix : 0u) + incl - c;
      HD[tid] = make_uint2(excl, g);
      if (tid == JR_NT - 1) sh_total = excl + c;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < JR_ITEMS; j++) {
      if (br[j] == 0xFFFFFFFFu) continue;
      const uint2 hd = HD[br[j] >> 16];
      const uint32_t rank = br[j] & 0xFFFFu, pos = hd.x + rank;
      st_key[pos] = key[j];
      st_row[pos] = (uint32_t)(t0 + (long long)j * JR_NT + tid);
      st_dst[pos] = hd.y == 0xFFFFFFFFu ? 0xFFFFFFFFu : hd.y + rank;       // all ones: the bucket overflowed, the caller discards this partitioning
    }
    if (fast && t0 + tstride < n) load_tile(t0 + tstride);
    __syncthreads();
    const uint32_t total = sh_total;
    for (uint32_t pos = tid; pos < total; pos += JR_NT) {     // consecutive staged rows of a bucket go to consecutive output rows
      const uint32_t d = st_dst[pos];
      if (d == 0xFFFFFFFFu) continue;
      out_keys[d] = st_key[pos];
      out_rows[d] = st_row[pos];
    }
    for (int p = 0; p < pay.n; p++) {      // payload columns: the same tile, the same positions, through the same staging buffer
      __syncthreads();
#pragma unroll
      for (int j = 0; j < JR_ITEMS; j++) {
        if (br[j] == 0xFFFFFFFFu) continue;
        const long long i = t0 + (long long)j * JR_NT + tid;
        u64 v = __ldcs(pay.src[p] + i);
        if (pay.nulls[p] && pdrs_bit(pay.nulls[p], i)) v = 0ull;          // a NULL source value yields the type default (join.rs:290-361)
        st_key[HD[br[j] >> 16].x + (br[j] & 0xFFFFu)] = v;
      }
      __syncthreads();
      u64* __restrict__ dst = pay.dst[p];
      for (uint32_t pos = tid; pos < total; pos += JR_NT) {
        const uint32_t d = st_dst[pos];
        if (d != 0xFFFFFFFFu) dst[d] = st_key[pos];
      }
    }
    __syncthreads();
  }
}

// ---- 32-byte slots {key, row, p0, p1}; linear probing, one slot = one sector; slot `nslots` belongs to the all-ones key
struct JFatTab { ulonglong4* slots; uint32_t nslots; uint32_t nbmul; };
static constexpr u64 JF_NOROW = ~0ull;
// fail[0] rows that found no slot, fail[3] rows whose key was already there (duplicate build keys: the caller falls back)
__global__ void __launch_bounds__(256) jfat_build_kernel(JFatTab t, const u64* __restrict__ pkeys, const uint32_t* __restrict__ prows, const u64* __restrict__ p0,
                                                         const u64* __restrict__ p1, const u64* __restrict__ cnt, long long cap, u64* __restrict__ fail) {
  const long long n = min((long long)__ldg(cnt), cap);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const u64 key = __ldcs(pkeys + i), row = (u64)__ldcs(prows + i);
    const u64 a = p0 ? __ldcs(p0 + i) : 0ull, b = p1 ? __ldcs(p1 + i) : 0ull;
    if (key == J_EMPTY_KEY) {
      ulonglong4* s = t.slots + t.nslots;
      if (atomicCAS(&s->y, JF_NOROW, row) != JF_NOROW) atomicAdd(&fail[3], 1ull); else { s->z = a; s->w = b; }
      continue;
    }
    uint32_t s = __umulhi(jhash32(key) * t.nbmul, t.nslots), probe = 0;
    for (; probe <= t.nslots; probe++) {
      const u64 old = atomicCAS(&t.slots[s].x, J_EMPTY_KEY, key);
      if (old == J_EMPTY_KEY) { t.slots[s].y = row; t.slots[s].z = a; t.slots[s].w = b; break; }
      if (old == key) { atomicAdd(&fail[3], 1ull); break; }
      s = s + 1 == t.nslots ? 0 : s + 1;
    }
    if (probe > t.nslots) atomicAdd(&fail[0], 1ull);
  }
}

// Probe + emit of one bucket: every warp takes tickets of 512 probe rows (in order), works through them in units of 128
// (4 slot loads per lane in flight), follows longer probe sequences in warp-synchronous rounds, compacts the pairs
// with ballots and reserves the output with one atomic per unit.  Emits (left row, right row [, p0, p1]); Left join:
// (left row, -1 [, 0, 0]).
#define JF_ITEMS 4
#define JF_UNIT (32 * JF_ITEMS)
#define JF_TICKET (4 * JF_UNIT)
__global__ void __launch_bounds__(256, 4) jfat_probe_emit_kernel(JFatTab t, const u64* __restrict__ pkeys, const uint32_t* __restrict__ prows, const u64* __restrict__ cnt,
                                                               long long cap, int left_join, int npay, long long* __restrict__ out_l, long long* __restrict__ out_r,
                                                               u64* __restrict__ out_p0, u64* __restrict__ out_p1, u64* __restrict__ out_cursor, u64* __restrict__ tile_ctr) {
  constexpr uint32_t NONE = 0xFFFFFFFFu;
  const long long n = min((long long)__ldg(cnt), cap);
  const long long ntickets = (n + JF_TICKET - 1) / JF_TICKET;
  const u64 pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  const int lane = threadIdx.x & 31;
  const uint32_t nslots = t.nslots, nbmul = t.nbmul;
  const u64* tab = reinterpret_cast<const u64*>(t.slots);
  for (;;) {
    long long ticket = 0;
    if (lane == 0) ticket = (long long)atomicAdd(tile_ctr, 1ull);
    ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
    if (ticket >= ntickets) break;
#pragma unroll 1
    for (long long lo = ticket * JF_TICKET; lo < min(n, (ticket + 1) * JF_TICKET); lo += JF_UNIT) {
      u64 key[JF_ITEMS], pa[JF_ITEMS], pb[JF_ITEMS];
      uint32_t lrow[JF_ITEMS], slot[JF_ITEMS], rrow[JF_ITEMS];
      uint32_t okmask = 0, pend = 0;
#pragma unroll
      for (int j = 0; j < JF_ITEMS; j++) {
        const long long i = lo + j * 32 + lane;
        key[j] = 0; lrow[j] = 0;
        if (i < n) { key[j] = ld_stream_u64(pkeys + i, pol_stream); lrow[j] = ld_stream_u32(prows + i, pol_stream); okmask |= 1u << j; }
      }
      {
        ulonglong4 v[JF_ITEMS];
#pragma unroll
        for (int j = 0; j < JF_ITEMS; j++) {
          slot[j] = key[j] == J_EMPTY_KEY ? nslots : __umulhi(jhash32(key[j]) * nbmul, nslots);
          v[j] = ld_bucket(tab + 4ull * slot[j], pol_keep);
        }
#pragma unroll
        for (int j = 0; j < JF_ITEMS; j++) {
          rrow[j] = NONE; pa[j] = 0; pb[j] = 0;
          if (!((okmask >> j) & 1u)) continue;
          bool hit = v[j].x == key[j];
          if (key[j] == J_EMPTY_KEY) hit = v[j].y != JF_NOROW;                 // the reserved slot of the all-ones key
          if (hit) { rrow[j] = (uint32_t)v[j].y; pa[j] = v[j].z; pb[j] = v[j].w; }
          else if (v[j].x != J_EMPTY_KEY) pend |= 1u << j;
        }
      }
      // rows displaced from their home slot: every lane follows the probe sequence of its first pending row per round
      uint32_t steps = 0;
      while (__any_sync(0xFFFFFFFFu, pend != 0)) {
        const int j0 = pend ? __ffs(pend) - 1 : -1;
        u64 k0 = 0;
        uint32_t s0 = 0;
#pragma unroll
        for (int j = 0; j < JF_ITEMS; j++) if (j == j0) { k0 = key[j]; s0 = slot[j]; }
        if (j0 >= 0) {
          s0 = s0 + 1 >= nslots ? 0 : s0 + 1;
          const ulonglong4 q = ld_bucket(tab + 4ull * s0, pol_keep);
          const bool hit = q.x == k0, end = q.x == J_EMPTY_KEY || ++steps > nslots;
#pragma unroll
          for (int j = 0; j < JF_ITEMS; j++) if (j == j0) { slot[j] = s0; if (hit) { rrow[j] = (uint32_t)q.y; pa[j] = q.z; pb[j] = q.w; } }
          if (hit || end) { pend &= ~(1u << j0); steps = 0; }
        }
      }
      // compaction inside the warp: slab j = the 32 rows of item j; one output reservation per unit
      uint32_t emit = 0, wcnt = 0, woff[JF_ITEMS];
#pragma unroll
      for (int j = 0; j < JF_ITEMS; j++) {
        const bool e = ((okmask >> j) & 1u) && (left_join || rrow[j] != NONE);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, e);
        if (e) emit |= 1u << j;
        woff[j] = wcnt + __popc(m & ((1u << lane) - 1u));
        wcnt += __popc(m);
      }
      u64 base = 0;
      if (lane == 0 && wcnt) base = atomicAdd(out_cursor, (u64)wcnt);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
#pragma unroll
      for (int j = 0; j < JF_ITEMS; j++) {
        if (!((emit >> j) & 1u)) continue;
        const u64 o = base + woff[j];
        st_stream_u64(out_l + o, (long long)lrow[j], pol_stream);
        st_stream_u64(out_r + o, rrow[j] == NONE ? -1ll : (long long)rrow[j], pol_stream);
        if (npay > 0) st_stream_u64(reinterpret_cast<long long*>(out_p0) + o, (long long)pa[j], pol_stream);
        if (npay > 1) st_stream_u64(reinterpret_cast<long long*>(out_p1) + o, (long long)pb[j], pol_stream);
      }
    }
  }
}

struct pdrs_join_result {
  pdrs_ctx* ctx = nullptr;
  int64_t n = 0;
  int64_t cap = 0;                    // pairs the left / right arrays have room for (single-pass path; rounds append)
  DevBuf left, right;
  int npay = 0;                       // pdrs_join_gather: materialised columns of the right frame
  int pay_dtype[PDRS_MAX_VALS] = {};
  DevBuf pay[PDRS_MAX_VALS];
};

// pairs of several join results, one after the other (the rounds of pdrs_join_pairs_dist); takes ownership of the parts
int32_t pdrs_join_result_concat(pdrs_ctx* c, pdrs_join_result** parts, int n, pdrs_join_result** out) {
  auto* res = new pdrs_join_result();
  res->ctx = c;
  int64_t total = 0;
  for (int i = 0; i < n; i++) total += parts[i]->n;
  int32_t st = res->left.alloc(c, (size_t)std::max<int64_t>(total, 1) * 8);
  if (st == PDRS_OK) st = res->right.alloc(c, (size_t)std::max<int64_t>(total, 1) * 8);
  int64_t at = 0;
  for (int i = 0; i < n && st == PDRS_OK; i++) {
    if (parts[i]->n) {
      if (cudaMemcpyAsync(res->left.as<int64_t>() + at, parts[i]->left.p, (size_t)parts[i]->n * 8, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess ||
          cudaMemcpyAsync(res->right.as<int64_t>() + at, parts[i]->right.p, (size_t)parts[i]->n * 8, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess)
        st = pdrs_fail(c, PDRS_ERR_CUDA, "join: concatenating the rounds failed");
    }
    at += parts[i]->n;
  }
  for (int i = 0; i < n; i++) { delete parts[i]; parts[i] = nullptr; }
  if (st != PDRS_OK) { delete res; return st; }
  res->n = total;
  cudaStreamSynchronize(c->stream);
  *out = res;
  return PDRS_OK;
}

pdrs_join_result* pdrs_join_result_new(pdrs_ctx* c) { auto* r = new pdrs_join_result(); r->ctx = c; return r; }

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
  PDRS_CUDA(c, cudaFuncSetAttribute(jpart_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: every call
  jpart_scatter_kernel<<<ctas, JP_THREADS, smem, c->stream>>>(col, n, log_nb, h + nb, out->keys.as<u64>(), out->rows.as<uint32_t>());
  c->stats.kernel_launches += 3;
  PDRS_CUDA(c, cudaGetLastError());
  return PDRS_OK;
}

// One-pass partition into padded buckets; *ok = false when a bucket overflowed (heavily duplicated keys).
// `staged` (exchange path): the input is the (key, row) records received from the peers - n positions in padded
// source regions holding `rows_in` rows - instead of a key column.
static int32_t jpartition1(pdrs_ctx* c, const JKeyCol& col, long long n, int log_nb, JPart* out, DevBuf* counts, long long* cap_out, bool* ok,
                           const JStaged* staged = nullptr, long long rows_in = 0) {
  const int nb = 1 << log_nb;
  const long long nrows = staged ? rows_in : n;
  long long cap = nrows / nb + nrows / (nb * 32ll) + 65536;           // ~3% + 64K rows of slack per bucket
  cap = (cap + JQ_TILE - 1) / JQ_TILE * JQ_TILE;             // multiple of every tile size used downstream
  *cap_out = cap;
  PDRS_TRY(counts->alloc(c, (size_t)(nb + 2) * 8, true));    // [nb] bucket cursors, [nb] overflow count
  PDRS_TRY(out->keys.alloc(c, (size_t)nb * cap * 8));
  PDRS_TRY(out->rows.alloc(c, (size_t)nb * cap * 4));
  out->n = (long long)nb * cap;
  if ((unsigned long long)nb * (unsigned long long)cap >= (1ull << 32)) { *ok = false; return PDRS_OK; }   // 32-bit output positions
  const size_t smem = JQ_SMEM;
  // the attribute is per device (several contexts / GPUs may live in one process): set it on every call
  PDRS_CUDA(c, cudaFuncSetAttribute(jpart1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PDRS_CUDA(c, cudaFuncSetAttribute(jpart1_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 2, (n + JQ_TILE - 1) / JQ_TILE));
  JXDst x{};
  x.keys[0] = out->keys.as<u64>(); x.rows[0] = out->rows.as<uint32_t>();
  if (staged) jpart1_kernel<false, true><<<ctas, JQ_NT, smem, c->stream>>>(col, n, log_nb, cap, counts->as<u64>(), x, counts->as<u64>() + nb, *staged);
  else jpart1_kernel<false><<<ctas, JQ_NT, smem, c->stream>>>(col, n, log_nb, cap, counts->as<u64>(), x, counts->as<u64>() + nb);
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 8, counts->as<u64>() + nb, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  *ok = c->pinned_scalars[8] == 0;
  return PDRS_OK;
}


// Build the table from `rsrc`, probe it with `lsrc` and materialise the pairs into `res` (left / right arrays).
// nl_out = number of probe rows (capacity of the single-pass output); nl_eff / nr_eff = positions to scan (padded
// partition layouts scan cap rows per bucket); nr_rows = build rows (size of the CSR array of the duplicate-key case).
// phases: 1 = build, 2 = probe (3 = both).  `bst` carries what the probe needs to know about the build (duplicate keys: the CSR
// segments) when the two run in different calls (exchange join in rounds: one build, several probes that APPEND to `res`;
// cap_hint = pairs to make room for when the arrays are first allocated).
struct JBuildState { bool dups = false; DevBuf csr; };
static int32_t jbuild_probe(pdrs_ctx* c, const JTab& jt, DevBuf& fail, const JSrc& rsrc, long long nr_eff, int64_t nr_rows, const JSrc& lsrc, long long nl_eff,
                            int64_t nl, bool radix, int how, pdrs_join_result* res, int64_t* M_out, const std::function<void(const char*)>& mark,
                            int phases = 3, JBuildState* bst = nullptr, int64_t cap_hint = 0) {
  DevBuf counts;
  JBuildState local_state;
  if (!bst) bst = &local_state;
  DevBuf& csr_buf = bst->csr;
  const int bctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 8, (nr_eff + 255) / 256));
  if (phases & 1) {
  bst->dups = false;
  if (nr_eff > 0) {
    join_build_kernel<0><<<bctas, 256, 0, c->stream>>>(jt, nullptr, rsrc, nr_eff, fail.as<u64>());
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  mark("build");
  // fail[0] failed inserts, fail[3] build rows whose key was already in the table
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, fail.p, 32, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->pinned_scalars[0] != 0) return pdrs_fail(c, PDRS_ERR_CUDA, "join build: hash table insertion failed for %lld rows", (long long)c->pinned_scalars[0]);
  bst->dups = c->pinned_scalars[3] != 0;
  if (bst->dups) {
    // Duplicate build keys: head word = number of rows of the key (count pass) -> offset of its segment [length, rows ...] in
    // the CSR array (reserve) -> rows filled in (fill pass) -> every segment sorted ascending, once.
    if (nr_rows >= (1ll << 31) - 8) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "join: %lld build rows with duplicate keys (limit 2^31)", (long long)nr_rows);
    DevBuf longs, scratch;
    PDRS_TRY(csr_buf.alloc(c, (size_t)(2 * nr_rows + 8) * 4));
    const long long longs_cap = nr_rows / 33 + 1;
    PDRS_TRY(longs.alloc(c, (size_t)longs_cap * 8));
    PDRS_CUDA(c, cudaMemsetAsync(jt.heads, 0, (size_t)(jt.slots + 4) * 4, c->stream));
    PDRS_CUDA(c, cudaMemsetAsync(fail.as<u64>() + 5, 0, 24, c->stream));          // [5] CSR cursor, [6] long segments, [7] longest (padded)
    PDRS_CUDA(c, cudaMemsetAsync(fail.as<u64>() + 1, 0, 8, c->stream));
    join_build_kernel<1><<<bctas, 256, 0, c->stream>>>(jt, nullptr, rsrc, nr_eff, fail.as<u64>());
    const int sg = pdrs_grid_for(c, (long long)jt.slots + 1, 256);
    jdup_reserve_kernel<<<sg, 256, 0, c->stream>>>(jt, csr_buf.as<uint32_t>(), fail.as<u64>() + 5);
    PDRS_CUDA(c, cudaMemsetAsync(fail.as<u64>() + 1, 0, 8, c->stream));
    join_build_kernel<2><<<bctas, 256, 0, c->stream>>>(jt, csr_buf.as<uint32_t>(), rsrc, nr_eff, fail.as<u64>());
    jdup_sort_short_kernel<<<sg, 256, 0, c->stream>>>(jt, csr_buf.as<uint32_t>(), fail.as<u64>() + 5, longs.as<uint2>(), longs_cap);
    c->stats.kernel_launches += 4;
    PDRS_CUDA(c, cudaGetLastError());
    PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, fail.p, 64, cudaMemcpyDeviceToHost, c->stream));
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->pinned_scalars[0] != 0) return pdrs_fail(c, PDRS_ERR_CUDA, "join build: %lld rows did not find their key again", (long long)c->pinned_scalars[0]);
    const long long nlong = std::min<long long>(c->pinned_scalars[6], longs_cap), mmax = c->pinned_scalars[7];
    if (nlong > 0) {
      const int lg = (int)std::min<long long>(nlong, (long long)c->sm_count * 2);
      if (mmax > JDS_SMEM_ROWS) PDRS_TRY(scratch.alloc(c, (size_t)lg * (size_t)mmax * 4));
      jdup_sort_long_kernel<<<lg, JDS_NT, 0, c->stream>>>(csr_buf.as<uint32_t>(), longs.as<uint2>(), nlong, scratch.as<uint32_t>(), (u64)mmax);
      c->stats.kernel_launches++;
      PDRS_CUDA(c, cudaGetLastError());
    }
    mark("duplicate keys: segments");
  }
  }   // phases & 1
  if (!(phases & 2)) { *M_out = res ? res->n : 0; return PDRS_OK; }
  const bool dups = bst->dups;
  const uint32_t* csr = dups ? csr_buf.as<uint32_t>() : nullptr;
  // unique build keys (the usual dimension-table join): single-pass probe + emit.  Needs one output slot per probe row.
  const bool single_pass = radix && c->opt_join_emit != 2 && !dups;   // small inputs keep the reference's left-row-major order
  int64_t M = 0;
  if (single_pass) {
    // the pairs of this probe are appended behind the res->n pairs already there (0 except in the later rounds of an exchange join)
    const int64_t base = res->n, need = base + std::max<int64_t>(nl, 1);
    if (res->cap < need) {
      const int64_t ncap = std::max<int64_t>(need, cap_hint);
      DevBuf nl_buf, nr_buf;
      PDRS_TRY(nl_buf.alloc(c, (size_t)ncap * 8));
      PDRS_TRY(nr_buf.alloc(c, (size_t)ncap * 8));
      if (base > 0) {
        PDRS_CUDA(c, cudaMemcpyAsync(nl_buf.p, res->left.p, (size_t)base * 8, cudaMemcpyDeviceToDevice, c->stream));
        PDRS_CUDA(c, cudaMemcpyAsync(nr_buf.p, res->right.p, (size_t)base * 8, cudaMemcpyDeviceToDevice, c->stream));
      }
      res->left = std::move(nl_buf); res->right = std::move(nr_buf); res->cap = ncap;
    }
    // output cursor = pairs so far, probe tile tickets from 0
    c->pinned_scalars[16] = base; c->pinned_scalars[17] = 0;
    PDRS_CUDA(c, cudaMemcpyAsync(fail.as<u64>() + 2, c->pinned_scalars + 16, 8, cudaMemcpyHostToDevice, c->stream));
    PDRS_CUDA(c, cudaMemcpyAsync(fail.as<u64>() + 4, c->pinned_scalars + 17, 8, cudaMemcpyHostToDevice, c->stream));
    M = base;
    if (nl_eff > 0) {
      const long long ntiles = (nl_eff + JE_TILE - 1) / JE_TILE;
      const int ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * (c->opt_join_ctas_per_sm > 0 ? c->opt_join_ctas_per_sm : 8), ntiles));
      if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
      join_probe_emit_kernel<<<ctas, JE_THREADS, 0, c->stream>>>(jt, lsrc, nl_eff, how == PDRS_LEFT, res->left.as<long long>(), res->right.as<long long>(),
                                                                 fail.as<u64>() + 2, fail.as<u64>() + 4);
      if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
      c->stats.kernel_launches++;
      PDRS_CUDA(c, cudaGetLastError());
      PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, fail.as<u64>() + 2, 8, cudaMemcpyDeviceToHost, c->stream));
      PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
      M = c->pinned_scalars[0];
    }
    res->n = M;
    mark("probe+emit");
  } else {
  const long long ntiles = (nl_eff + JOIN_TILE - 1) / JOIN_TILE;
  const int ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * (c->opt_join_ctas_per_sm > 0 ? c->opt_join_ctas_per_sm : 8), ntiles));
  PDRS_TRY(counts.alloc(c, (size_t)(ntiles + 4) * 8, true));
  DevBuf stash;
  PDRS_TRY(stash.alloc(c, (size_t)std::max<int64_t>(nl_eff, 1) * 8));
  u64* cc = counts.as<u64>();
  if (nl_eff > 0) {
    PDRS_CUDA(c, cudaMemsetAsync(fail.as<u64>() + 2, 0, 8, c->stream));          // tile tickets from 0 (the buffer may have served an earlier probe)
    if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
    join_probe_kernel<<<ctas, JOIN_THREADS, 0, c->stream>>>(jt, csr, lsrc, nl_eff, how == PDRS_LEFT, stash.as<long long>(), cc, fail.as<u64>() + 2);
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
  M = nl_eff > 0 ? c->pinned_scalars[0] : 0;
  res->n = M;
  PDRS_TRY(res->left.alloc(c, (size_t)std::max<int64_t>(M, 1) * 8));
  PDRS_TRY(res->right.alloc(c, (size_t)std::max<int64_t>(M, 1) * 8));
  if (M > 0) {
    join_write_kernel<<<ctas, JOIN_THREADS, 0, c->stream>>>(stash.as<long long>(), csr, nl_eff, how == PDRS_LEFT, cc, lsrc,
                                                            res->left.as<long long>(), res->right.as<long long>());
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  }
  *M_out = M;
  return PDRS_OK;
}


static int32_t join_aux_streams(pdrs_ctx* c) {
  if (c->aux_stream[0]) return PDRS_OK;
  for (int s = 0; s < 2; s++) {
    PDRS_CUDA(c, cudaStreamCreateWithFlags(&c->aux_stream[s], cudaStreamNonBlocking));
    PDRS_CUDA(c, cudaEventCreateWithFlags(&c->aux_done[s], cudaEventDisableTiming));
  }
  PDRS_CUDA(c, cudaEventCreateWithFlags(&c->aux_fork, cudaEventDisableTiming));
  return PDRS_OK;
}

// Bucket-at-a-time join (see the comment above jpartp_kernel).  *done = false: not applicable (a padded bucket overflowed =
// skewed keys, duplicate build keys, sizes) - nothing was emitted and the caller takes the one-table path.
static int32_t join_bucketwise(pdrs_ctx* c, const ColView& lv, const ColView& rv, int how, const ColView* pay, int npay, pdrs_join_result* res, bool* done,
                               const std::function<void(const char*)>& mark) {
  *done = false;
  const int64_t nl = lv.len, nr = rv.len;
  if (nl >= (1ll << 32) - 1 || nr >= (1ll << 32) - 1 || nr < 1 || nl < 1 || npay > JP_MAXPAY) return PDRS_OK;
  const bool fat = npay > 0;
  const long long spk = c->opt_join_slots_mult > 0 ? c->opt_join_slots_mult : 3;                   // slots per build key
  const size_t slot_bytes = fat ? 32 : 12;
  const size_t region_target = (size_t)(c->opt_join_region_mb > 0 ? c->opt_join_region_mb : 24) << 20;
  long long NB = (long long)(((size_t)nr * spk * slot_bytes + region_target - 1) / region_target);
  if (c->opt_join_log_nb > 0) NB = 1ll << c->opt_join_log_nb;
  NB = std::max<long long>(2, std::min<long long>(1024, NB));
  auto cap_for = [&](long long n) {
    const double m = (double)n / (double)NB;
    const long long cap = (long long)(m + m / 32.0 + 6.0 * std::sqrt(m + 1.0)) + 1024;
    return (cap + 4095) / 4096 * 4096;
  };
  const long long rcap = cap_for(nr), lcap = cap_for(nl);
  if ((unsigned long long)NB * rcap >= (1ull << 32) - 1 || (unsigned long long)NB * lcap >= (1ull << 32) - 1) return PDRS_OK;   // 32-bit output positions
  PDRS_TRY(join_aux_streams(c));
  // ---- partition both sides (one pass each; the build side carries its payload columns)
  DevBuf ctr, rk, rr, lk, lr, pk[JP_MAXPAY];
  PDRS_TRY(ctr.alloc(c, (size_t)(2 * NB + 8 + 8 * NB) * 8, true));     // [NB] build cursors, [NB] probe cursors, [8] {overflow, out cursor}, [NB][8] per-bucket counters
  u64* rcur = ctr.as<u64>();
  u64* lcur = rcur + NB;
  u64* glob = lcur + NB;                 // [0] overflowed reservations, [1] output cursor
  u64* bctr = glob + 8;                  // per bucket: [0] failed inserts, [1] build tickets, [3] duplicate keys, [4] probe tickets
  PDRS_TRY(rk.alloc(c, (size_t)NB * rcap * 8));
  PDRS_TRY(rr.alloc(c, (size_t)NB * rcap * 4));
  PDRS_TRY(lk.alloc(c, (size_t)NB * lcap * 8));
  PDRS_TRY(lr.alloc(c, (size_t)NB * lcap * 4));
  JPay jp{};
  jp.n = npay;
  for (int p = 0; p < npay; p++) {
    PDRS_TRY(pk[p].alloc(c, (size_t)NB * rcap * 8));
    jp.src[p] = reinterpret_cast<const u64*>(pay[p].data); jp.nulls[p] = pay[p].nulls; jp.dst[p] = pk[p].as<u64>();
  }
  PDRS_CUDA(c, cudaFuncSetAttribute(jpartp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JR_SMEM));
  const JKeyCol rc{rv.data, rv.nulls, rv.dtype}, lc{lv.data, lv.nulls, lv.dtype};
  jpartp_kernel<<<(int)std::max<long long>(1, std::min<long long>(c->sm_count, (nr + JR_TILE - 1) / JR_TILE)), JR_NT, JR_SMEM, c->stream>>>(
      rc, nr, (uint32_t)NB, rcap, rcur, rk.as<u64>(), rr.as<uint32_t>(), glob, jp);
  mark("partition build side");
  jpartp_kernel<<<(int)std::max<long long>(1, std::min<long long>(c->sm_count, (nl + JR_TILE - 1) / JR_TILE)), JR_NT, JR_SMEM, c->stream>>>(
      lc, nl, (uint32_t)NB, lcap, lcur, lk.as<u64>(), lr.as<uint32_t>(), glob, JPay{});
  c->stats.kernel_launches += 2;
  PDRS_CUDA(c, cudaGetLastError());
  mark("partition probe side");
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 8, glob, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->pinned_scalars[8] != 0) return PDRS_OK;                       // skewed keys: a bucket overflowed its padded range
  // ---- two table regions, one per stream
  const long long slots = (spk * rcap + 3) / 4 * 4;
  const size_t region_bytes = fat ? (size_t)(slots + 1) * 32 : (size_t)(slots + 4) * 12;
  DevBuf region[2];
  for (int s = 0; s < 2; s++) PDRS_TRY(region[s].alloc(c, region_bytes + 256));
  PDRS_TRY(res->left.alloc(c, (size_t)nl * 8));
  PDRS_TRY(res->right.alloc(c, (size_t)nl * 8));
  for (int p = 0; p < npay; p++) PDRS_TRY(res->pay[p].alloc(c, (size_t)nl * 8));
  c->stats.table_slots = slots;
  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
  PDRS_CUDA(c, cudaEventRecord(c->aux_fork, c->stream));
  for (int s = 0; s < 2; s++) PDRS_CUDA(c, cudaStreamWaitEvent(c->aux_stream[s], c->aux_fork, 0));
  const int bgrid = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 8, (rcap + 255) / 256));
  const int tile_units = 2;                                            // compact probe: tickets of 512 rows
  const long long ptickets = fat ? (lcap + JF_TICKET - 1) / JF_TICKET : (lcap + JE_UNIT * tile_units - 1) / (JE_UNIT * tile_units);
  const int pgrid = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * (c->opt_join_ctas_per_sm > 0 ? c->opt_join_ctas_per_sm : 4), (ptickets + 7) / 8));
  for (long long b = 0; b < NB; b++) {
    const int s = (int)(b & 1);
    cudaStream_t st = c->aux_stream[s];
    u64* bc = bctr + 8 * b;
    PDRS_CUDA(c, cudaMemsetAsync(region[s].p, 0xFF, region_bytes, st));      // key = all ones (EMPTY), head / row = all ones (none)
    if (fat) {
      const JFatTab ft{region[s].as<ulonglong4>(), (uint32_t)slots, (uint32_t)NB};
      jfat_build_kernel<<<bgrid, 256, 0, st>>>(ft, rk.as<u64>() + b * rcap, rr.as<uint32_t>() + b * rcap, pk[0].as<u64>() + b * rcap,
                                              npay > 1 ? pk[1].as<u64>() + b * rcap : nullptr, rcur + b, rcap, bc);
      jfat_probe_emit_kernel<<<pgrid, 256, 0, st>>>(ft, lk.as<u64>() + b * lcap, lr.as<uint32_t>() + b * lcap, lcur + b, lcap, how == PDRS_LEFT, npay,
                                                   res->left.as<long long>(), res->right.as<long long>(), res->pay[0].as<u64>(), res->pay[1].as<u64>(), glob + 1, bc + 4);
    } else {
      const JTab jt{region[s].as<u64>(), reinterpret_cast<uint32_t*>(region[s].as<u64>() + slots + 4), (u64)slots, (uint32_t)NB};
      JSrc rs{}, ls{};
      rs.pkeys = rk.as<u64>() + b * rcap; rs.prows = rr.as<uint32_t>() + b * rcap; rs.cap = rcap; rs.cnt = rcur + b;
      ls.pkeys = lk.as<u64>() + b * lcap; ls.prows = lr.as<uint32_t>() + b * lcap; ls.cap = lcap; ls.cnt = lcur + b;
      join_build_kernel<0><<<bgrid, 256, 0, st>>>(jt, nullptr, rs, rcap, bc);
      join_probe_emit_kernel<<<pgrid, JE_THREADS, 0, st>>>(jt, ls, lcap, how == PDRS_LEFT, res->left.as<long long>(), res->right.as<long long>(), glob + 1, bc + 4, tile_units);
    }
  }
  c->stats.kernel_launches += 2 * NB;
  PDRS_CUDA(c, cudaGetLastError());
  for (int s = 0; s < 2; s++) {
    PDRS_CUDA(c, cudaEventRecord(c->aux_done[s], c->aux_stream[s]));
    PDRS_CUDA(c, cudaStreamWaitEvent(c->stream, c->aux_done[s], 0));
  }
  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
  mark("buckets: clear + build + probe");
  std::vector<u64> hc((size_t)8 * NB + 8);
  PDRS_CUDA(c, cudaMemcpyAsync(hc.data(), glob, hc.size() * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  u64 failed = 0, dups = 0;
  for (long long b = 0; b < NB; b++) { failed += hc[8 + 8 * b]; dups += hc[8 + 8 * b + 3]; }
  if (failed) return pdrs_fail(c, PDRS_ERR_CUDA, "join build: hash table insertion failed for %llu rows", (unsigned long long)failed);
  if (dups) {                                                          // duplicate build keys: the one-table path builds the CSR segments
    res->left.release(); res->right.release();
    for (int p = 0; p < npay; p++) res->pay[p].release();
    return PDRS_OK;
  }
  res->n = (int64_t)hc[1];
  *done = true;
  return PDRS_OK;
}

extern "C" {

}  // extern "C"

// pdrs_join_pairs (nrcols == 0) and pdrs_join_gather (columns of the right frame materialised along with the pairs)
static int32_t join_run(pdrs_ctx* c, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how, const pdrs_col* rcols, int32_t nrcols, pdrs_join_result** out) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!left_key || !right_key || !out || nrcols < 0 || nrcols > PDRS_MAX_VALS || (nrcols && !rcols)) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_join: bad argument");
  if (how < PDRS_INNER || how > PDRS_OUTER) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "unknown join type %d", how);
  for (int k = 0; k < nrcols; k++)
    if (rcols[k].len != right_key->len) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "right column %d has %lld rows, the key column %lld", k, (long long)rcols[k].len, (long long)right_key->len);
  const int32_t how_req = how;
  how = (how == PDRS_RIGHT) ? PDRS_INNER : (how == PDRS_OUTER ? PDRS_LEFT : how);   // Right / Outer = Inner / Left + the unmatched right rows
  if (left_key->dtype != right_key->dtype)   // join.rs:98-104
    return pdrs_fail(c, PDRS_ERR_TYPE_MISMATCH, "join key columns have different types (%d vs %d)", left_key->dtype, right_key->dtype);
  PDRS_CUDA(c, cudaSetDevice(c->device));
  pdrs_settle_frees(c);       // the multi-GB buffers of the previous call are reusable once the host has seen their frees (common.cuh)
  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
  c->stats.main_kernel_ms = 0; c->stats.total_ms = 0;
  ColView lv, rv;
  PDRS_TRY(pdrs_view_col(c, left_key, &lv));
  PDRS_TRY(pdrs_view_col(c, right_key, &rv));
  const int64_t nl = lv.len, nr = rv.len;
  auto* res = new pdrs_join_result();
  res->ctx = c;
  struct Guard { pdrs_join_result* r; ~Guard() { delete r; } } guard{res};

  const long long slots = (std::max<long long>(1024, (c->opt_join_slots_mult > 0 ? c->opt_join_slots_mult : 3) * nr) + 3) / 4 * 4;   // load factor <= 1/3: 4-slot buckets overflow for ~5% of the keys (see jslot)
  if (slots >= (1ll << 32) - 8) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "join: build side too large (%lld rows)", (long long)nr);
  const size_t table_bytes = (size_t)(slots + 4) * 12;                  // u64 keys[slots + 4] + u32 heads[slots + 4]; slot `slots` is reserved for the all-ones key
  // Large tables: radix-partition both sides so that one bucket's table region (<= 32 MB) stays in L2.
  // join_algo: 0 auto, 1 = always the direct (left-row-major output) path, 2 = always partitioned.
  int log_nb = 0;
  while (log_nb < 10 && (table_bytes >> log_nb) > (40ull << 20)) log_nb++;      // measured: 25-50 MB regions are L2-resident, fewer buckets partition faster
  if (c->opt_join_log_nb > 0) log_nb = (int)c->opt_join_log_nb;
  bool radix = log_nb > 1 && nl < (1ll << 32) && nr < (1ll << 32);
  if (c->opt_join_algo == 1) radix = false;
  if (c->opt_join_algo == 2 && nl < (1ll << 32) && nr < (1ll << 32)) { radix = true; log_nb = std::max(log_nb, 2); }
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  std::function<void(const char*)> mark = [&](const char* name) {
    if (c->opt_timing < 2) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c->stream); marks.push_back({name, e});
  };
  mark("start");
  // Bucket-at-a-time path: unique, hash-uniform build keys (the usual dimension-table join).  Up to JP_MAXPAY 8-byte
  // columns of the right frame travel with the build rows and come out with the pairs (pdrs_join_gather).
  bool fused_pay = nrcols > 0 && nrcols <= JP_MAXPAY && how_req <= PDRS_LEFT;
  for (int k = 0; k < nrcols; k++) fused_pay = fused_pay && (rcols[k].dtype == PDRS_I64 || rcols[k].dtype == PDRS_F64);
  bool bucketwise_done = false;
  long long nl_eff = nl, nr_eff = nr;
  int64_t M = 0;
  // Measured (1e9 x 1e8, profiles/join_phases_r02.txt): for pairs only the one-table path is faster (the per-bucket launches do
  // not overlap enough to pay for themselves, 21.2 vs 18.9 ms); with payload columns the bucket-at-a-time path wins (31.8 vs
  // 41.9 ms with gathers).  join_bucketwise: 1 = auto (payload columns only), 2 = always, 0 = never.
  if (radix && (c->opt_join_bucketwise == 2 || (c->opt_join_bucketwise == 1 && fused_pay)) && c->opt_join_part != 2 && c->opt_join_emit != 2) {
    std::vector<ColView> pv(fused_pay ? nrcols : 0);
    for (size_t k = 0; k < pv.size(); k++) PDRS_TRY(pdrs_view_col(c, &rcols[k], &pv[k]));
    PDRS_TRY(join_bucketwise(c, lv, rv, how, pv.data(), (int)pv.size(), res, &bucketwise_done, mark));
    if (bucketwise_done) { M = res->n; c->stats.groupby_algo_used = 3; if (fused_pay) { res->npay = nrcols; for (int k = 0; k < nrcols; k++) res->pay_dtype[k] = rcols[k].dtype; } }
  }
  if (!bucketwise_done) {
  DevBuf tab, fail;
  PDRS_TRY(tab.alloc(c, table_bytes));
  PDRS_CUDA(c, cudaMemsetAsync(tab.p, 0xFF, table_bytes, c->stream));   // key = all ones (EMPTY), head = -1 (EMPTY)
  JTab jt{tab.as<u64>(), reinterpret_cast<uint32_t*>(tab.as<u64>() + slots + 4), (u64)slots, 1u};
  PDRS_TRY(fail.alloc(c, 64, true));      // [0] failed inserts, [1] build tile counter, [2] probe tile counter (two-pass) / output cursor (single pass), [3] duplicate build keys, [4] probe tile counter (single pass)
  c->stats.table_slots = slots;
  JKeyCol rc{rv.data, rv.nulls, rv.dtype}, lc{lv.data, lv.nulls, lv.dtype};
  JPart lp, rp;
  DevBuf lcnt, rcnt;
  JSrc rsrc{rc, nullptr, nullptr, 0, nullptr, 0}, lsrc{lc, nullptr, nullptr, 0, nullptr, 0};
  if (radix) {
    mark("memset");
    // one-pass partition into padded buckets; exact two-pass partition when a bucket overflows (skewed keys)
    bool ok1 = c->opt_join_part != 2;
    long long rcap = 0, lcap = 0;
    if (ok1) PDRS_TRY(jpartition1(c, rc, nr, log_nb, &rp, &rcnt, &rcap, &ok1));
    if (ok1) { rsrc.cap = rcap; rsrc.cnt = rcnt.as<u64>(); rsrc.log_nb = c->opt_join_prefetch ? log_nb : 0; }
    else { rp = JPart(); PDRS_TRY(jpartition(c, rc, nr, log_nb, &rp)); }
    mark("partition build side");
    bool ok2 = c->opt_join_part != 2;
    if (ok2) PDRS_TRY(jpartition1(c, lc, nl, log_nb, &lp, &lcnt, &lcap, &ok2));
    if (ok2) { lsrc.cap = lcap; lsrc.cnt = lcnt.as<u64>(); lsrc.log_nb = c->opt_join_prefetch ? log_nb : 0; }
    else { lp = JPart(); PDRS_TRY(jpartition(c, lc, nl, log_nb, &lp)); }
    mark("partition probe side");
    rsrc.pkeys = rp.keys.as<u64>(); rsrc.prows = rp.rows.as<uint32_t>(); nr_eff = rp.n;
    lsrc.pkeys = lp.keys.as<u64>(); lsrc.prows = lp.rows.as<uint32_t>(); nl_eff = lp.n;
  }
  PDRS_TRY(jbuild_probe(c, jt, fail, rsrc, nr_eff, nr, lsrc, nl_eff, nl, radix, how, res, &M, mark));
  mark("write");
  c->stats.groupby_algo_used = radix ? 2 : 1;
  }
  if (how_req == PDRS_RIGHT || how_req == PDRS_OUTER) {
    DevBuf matched, ucounts;
    PDRS_TRY(matched.alloc(c, (size_t)std::max<int64_t>(nr, 1), true));
    const long long utiles = (nr + JU_TILE - 1) / JU_TILE;
    PDRS_TRY(ucounts.alloc(c, (size_t)(utiles + 4) * 8, true));
    const int ug = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 8, utiles));
    if (M > 0) jmark_matched_kernel<<<pdrs_grid_for(c, M, 256), 256, 0, c->stream>>>(res->right.as<long long>(), M, matched.as<uint8_t>());
    long long U = 0;
    if (nr > 0) {
      junmatched_count_kernel<<<ug, 256, 0, c->stream>>>(matched.as<uint8_t>(), nr, ucounts.as<u64>());
      join_scan_kernel<<<1, 1024, 0, c->stream>>>(ucounts.as<u64>(), (int)utiles, ucounts.as<u64>() + utiles + 2);
      PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, ucounts.as<u64>() + utiles + 2, 8, cudaMemcpyDeviceToHost, c->stream));
      PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
      U = c->pinned_scalars[0];
    }
    if (U > 0) {
      DevBuf nl2, nr2;
      PDRS_TRY(nl2.alloc(c, (size_t)(M + U) * 8));
      PDRS_TRY(nr2.alloc(c, (size_t)(M + U) * 8));
      if (M > 0) {
        PDRS_CUDA(c, cudaMemcpyAsync(nl2.p, res->left.p, (size_t)M * 8, cudaMemcpyDeviceToDevice, c->stream));
        PDRS_CUDA(c, cudaMemcpyAsync(nr2.p, res->right.p, (size_t)M * 8, cudaMemcpyDeviceToDevice, c->stream));
      }
      junmatched_write_kernel<<<ug, 256, 0, c->stream>>>(matched.as<uint8_t>(), nr, ucounts.as<u64>(), nl2.as<long long>() + M, nr2.as<long long>() + M);
      res->left = std::move(nl2);
      res->right = std::move(nr2);
      res->n = M + U;
    }
    c->stats.kernel_launches += 4;
    PDRS_CUDA(c, cudaGetLastError());
    mark("unmatched right rows");
  }
  // columns of the right frame that did not travel with the build rows: gathered by right row, type default for a missing
  // side or a NULL source value, no null mask (join.rs:290-552)
  if (nrcols > 0 && res->npay == 0) {
    for (int k = 0; k < nrcols; k++) {
      const int esz = rcols[k].dtype == PDRS_BOOL_BITS ? 1 : pdrs_dtype_bytes(rcols[k].dtype);
      PDRS_TRY(res->pay[k].alloc(c, (size_t)std::max<int64_t>(res->n, 1) * esz));
      res->pay_dtype[k] = rcols[k].dtype;
      PDRS_TRY(pdrs_gather(c, &rcols[k], res->right.as<int64_t>(), PDRS_MEM_DEVICE, res->n, res->pay[k].p, PDRS_MEM_DEVICE));
    }
    res->npay = nrcols;
    mark("gather right columns");
  }
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

extern "C" {

int32_t pdrs_join_pairs(pdrs_ctx* c, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how, pdrs_join_result** out) {
  return join_run(c, left_key, right_key, how, nullptr, 0, out);
}

int32_t pdrs_join_gather(pdrs_ctx* c, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how, const pdrs_col* right_cols, int32_t n_right_cols,
                         pdrs_join_result** out) {
  return join_run(c, left_key, right_key, how, right_cols, n_right_cols, out);
}

// ---------------------------------------------------------------- multi-GPU join: fused partition + shuffle over peer memory
// One pdrs_xjoin per rank (= per GPU / process).  Its receive area holds, for both join sides, (key, row) rows in
// padded sub-buckets [radix bucket][source rank] plus one row count per sub-bucket.  pdrs_xjoin_shuffle runs the
// one-pass partition kernel with the peers' receive areas as its output (CUDA IPC mappings, stores go through
// NVLink / NVSwitch): there is no send buffer and no separate all-to-all.  Fused layout: the receiver probes the
// sub-buckets as they arrived; staged layout (default for > 1 rank): regions [source rank] only, written with
// line-aligned stores, and the receiver runs the ordinary local radix partition over them (see pdrs_xjoin_create).
}  // extern "C"
struct XSide { size_t keys = 0, rows = 0, cnt = 0; long long cap = 0; };      // byte offsets inside the receive area
struct pdrs_xjoin {
  pdrs_ctx* ctx = nullptr;
  int rank = 0, world = 1, log_world = 0;
  int log_nb = 0;                       // radix buckets of the local join
  int sh_log_nb = 0;                    // radix bits applied by the shuffle kernel itself (fused mode: log_nb, staged mode: 0)
  int row_shift = 0;                    // staged mode: left rows travel as (source rank << row_shift | local row)
  int64_t total_right = 0;
  int64_t max_left = 0, max_right = 0;  // rows per rank the receive areas were sized for (pdrs_xjoin_create)
  XSide L, R;
  size_t bytes = 0;
  void* base = nullptr;                 // this rank's receive area (cudaMalloc: exportable through CUDA IPC)
  void* peer[8] = {};                   // receive areas of all ranks as seen from this process (peer[rank] == base)
  bool ipc_open[8] = {};
  bool attached = false, shuffled = false;
  // the hash table of the build side outlives a call: later rounds of pdrs_join_pairs_dist shuffle and probe more left rows only
  DevBuf tab, failb;
  JTab jt{};
  JBuildState bst;
  bool table_ready = false;
  int64_t nr_built = 0;
};
extern "C" {

int32_t pdrs_xjoin_create(pdrs_ctx* c, int32_t rank, int32_t world, int64_t max_left_rows, int64_t max_right_rows, int64_t total_right_rows, pdrs_xjoin** out) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!out || rank < 0 || rank >= world || world > 8 || (world & (world - 1)) || max_left_rows < 0 || max_right_rows < 0)
    return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_xjoin_create: world must be 1, 2, 4 or 8 and 0 <= rank < world");
  if (max_left_rows >= (1ll << 32) || total_right_rows >= (1ll << 31))
    return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_xjoin_create: more than 2^32 left rows per rank or 2^31 right rows in total");
  PDRS_CUDA(c, cudaSetDevice(c->device));
  auto* x = new pdrs_xjoin();
  x->ctx = c; x->rank = rank; x->world = world; x->total_right = total_right_rows;
  x->max_left = max_left_rows; x->max_right = max_right_rows;
  while ((1 << x->log_world) < world) x->log_world++;
  // radix buckets: the table region of one bucket of the rows this rank RECEIVES (~ total / world) stays L2-resident
  const size_t table_bytes = (size_t)(3 * (total_right_rows / world + 1) + 8) * 12;
  while (x->log_nb + x->log_world < 10 && (table_bytes >> x->log_nb) > (40ull << 20)) x->log_nb++;
  if (c->opt_join_log_nb > 0) x->log_nb = (int)std::min<int64_t>(c->opt_join_log_nb, 10 - x->log_world);
  // Two ways to get the rows into radix buckets on the right GPU (xjoin_mode: 0 auto, 1 fused, 2 staged):
  //  fused   the shuffle kernel partitions by (rank, radix bucket) at once: one pass, but a tile of 4096 rows is cut into
  //          world x 2^log_nb runs, and peer stores move whole 128-byte lines (tools/peer_store_bench.cu: 32-byte runs reach
  //          219 GB/s, 64-byte runs 438 GB/s, >= 128 bytes 716 GB/s over NVLink) - only good while runs stay >= 128 bytes
  //  staged  the shuffle kernel partitions by rank only (runs of 4096 / world rows = full NVLink speed), the receiver
  //          runs the ordinary local radix partition over what arrived (one more HBM pass, ~6 ms per 1e9 rows).  Left rows
  //          then lose their position-implied source, so they travel as (source << row_shift | local row).
  x->row_shift = 32 - x->log_world;
  const bool staged_ok = x->log_world > 0 && max_left_rows < (1ll << x->row_shift);
  const bool staged = c->opt_xjoin_mode == 2 ? staged_ok : (c->opt_xjoin_mode == 1 ? false : (staged_ok && x->log_world > 0));
  if (!staged) x->row_shift = 0;
  x->sh_log_nb = staged ? 0 : x->log_nb;
  const long long sb = 1ll << (x->sh_log_nb + x->log_world);         // sub-buckets per rank and side = combined buckets per source
  auto cap_for = [&](int64_t n) {
    const double m = (double)n / (double)sb;                         // rows one source sends to one sub-bucket (hash-uniform keys)
    long long cap = (long long)(m + m / 32.0 + 6.0 * std::sqrt(m + 1.0)) + 256;
    return (cap + JQ_TILE - 1) / JQ_TILE * JQ_TILE;                  // multiple of every tile size used downstream
  };
  x->L.cap = cap_for(max_left_rows); x->R.cap = cap_for(max_right_rows);
  if ((unsigned long long)sb * x->L.cap >= (1ull << 32) - 1 || (unsigned long long)sb * x->R.cap >= (1ull << 32) - 1) {
    delete x;
    return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_xjoin_create: receive area exceeds 32-bit positions");
  }
  size_t off = 0;
  auto place = [&](size_t n) { const size_t at = off; off += (n + 255) / 256 * 256; return at; };
  for (XSide* s : {&x->R, &x->L}) { s->keys = place((size_t)sb * s->cap * 8); s->rows = place((size_t)sb * s->cap * 4); s->cnt = place((size_t)sb * 8); }
  x->bytes = off;
  const cudaError_t e = cudaMalloc(&x->base, x->bytes);
  if (e != cudaSuccess) { delete x; return pdrs_fail(c, PDRS_ERR_OOM, "pdrs_xjoin_create: cudaMalloc(%zu) failed: %s", off, cudaGetErrorString(e)); }
  x->peer[rank] = x->base;
  x->attached = world == 1;
  *out = x;
  return PDRS_OK;
}

int64_t pdrs_xjoin_bytes(const pdrs_xjoin* x) { return x ? (int64_t)x->bytes : -1; }
void* pdrs_xjoin_base(const pdrs_xjoin* x) { return x ? x->base : nullptr; }

int32_t pdrs_xjoin_ipc_handle(pdrs_xjoin* x, uint8_t* handle64) {
  if (!x || !handle64) return PDRS_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  PDRS_CUDA(x->ctx, cudaSetDevice(x->ctx->device));
  PDRS_CUDA(x->ctx, cudaIpcGetMemHandle(&h, x->base));
  memcpy(handle64, &h, 64);
  return PDRS_OK;
}

// handles: world x 64 bytes, entry r = what rank r's pdrs_xjoin_ipc_handle returned (all_gather by the caller)
int32_t pdrs_xjoin_attach_ipc(pdrs_xjoin* x, const uint8_t* handles) {
  if (!x || !handles) return PDRS_ERR_BAD_ARG;
  PDRS_CUDA(x->ctx, cudaSetDevice(x->ctx->device));
  for (int r = 0; r < x->world; r++) {
    if (r == x->rank || x->ipc_open[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * r, 64);
    PDRS_CUDA(x->ctx, cudaIpcOpenMemHandle(&x->peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    x->ipc_open[r] = true;
  }
  x->attached = true;
  return PDRS_OK;
}

// same-process peers (one process driving several contexts / devices with peer access enabled): bases[r] = pdrs_xjoin_base of rank r
int32_t pdrs_xjoin_attach_ptrs(pdrs_xjoin* x, void* const* bases) {
  if (!x || !bases) return PDRS_ERR_BAD_ARG;
  for (int r = 0; r < x->world; r++) if (r != x->rank) x->peer[r] = bases[r];
  x->attached = true;
  return PDRS_OK;
}

// Partition both key columns by (destination rank, radix bucket) straight into the peers' receive areas and publish
// the sub-bucket counts.  Returns after this rank's stores are complete; the caller then runs a barrier over all
// ranks (and another one before the next shuffle, so that nobody overwrites an area that is still being read).
// Rows travel as (key, row): right rows carry GLOBAL row numbers (right_row0 + local row, < 2^31), left rows local
// ones (the receiver adds the source's left_row0).  Returns PDRS_ERR_UNSUPPORTED when a sub-bucket overflowed
// (heavily duplicated / skewed keys): the caller falls back to the all_to_all path.
int32_t pdrs_xjoin_shuffle(pdrs_xjoin* x, const pdrs_col* left_key, const pdrs_col* right_key, int64_t right_row0) {
  if (!x || !left_key) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = x->ctx;
  if (!x->attached) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_xjoin_shuffle: peers are not attached");
  // right_key == NULL: only left rows travel; the build side (and its hash table) of the previous shuffle stays (rounds)
  if (!right_key && !x->shuffled) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_xjoin_shuffle: no build side has been shuffled yet");
  if (right_key && left_key->dtype != right_key->dtype)
    return pdrs_fail(c, PDRS_ERR_TYPE_MISMATCH, "join key columns have different types (%d vs %d)", left_key->dtype, right_key->dtype);
  PDRS_CUDA(c, cudaSetDevice(c->device));
  ColView lv, rv;
  PDRS_TRY(pdrs_view_col(c, left_key, &lv));
  if (right_key) { PDRS_TRY(pdrs_view_col(c, right_key, &rv)); x->table_ready = false; }
  if (right_key && (right_row0 < 0 || right_row0 + rv.len > x->total_right)) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_xjoin_shuffle: right rows beyond total_right_rows");
  // the padded regions and (staged layout) the row encoding were sized for the row counts given to pdrs_xjoin_create
  if (lv.len > x->max_left || rv.len > x->max_right)
    return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_xjoin_shuffle: %lld left / %lld right rows exceed the %lld / %lld given to pdrs_xjoin_create",
                     (long long)lv.len, (long long)rv.len, (long long)x->max_left, (long long)x->max_right);
  if (x->row_shift && lv.len >= (1ll << x->row_shift))
    return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_xjoin_shuffle: %lld left rows do not fit the staged row encoding (2^%d)", (long long)lv.len, x->row_shift);
  const int ncb = 1 << (x->sh_log_nb + x->log_world);
  DevBuf cur;
  PDRS_TRY(cur.alloc(c, (size_t)(2 * ncb + 8) * 8, true));          // [ncb] right cursors, [ncb] left cursors, [1] overflow
  u64* ovf = cur.as<u64>() + 2 * ncb;
  PDRS_CUDA(c, cudaFuncSetAttribute(jpart1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JQ_SMEM));   // per device: every call
  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
  for (int side = right_key ? 0 : 1; side < 2; side++) {
    const ColView& v = side ? lv : rv;
    const XSide& s = side ? x->L : x->R;
    u64* cursor = cur.as<u64>() + side * ncb;
    JXDst d{};
    JXCnt dc{};
    for (int r = 0; r < x->world; r++) {
      d.keys[r] = reinterpret_cast<u64*>((char*)x->peer[r] + s.keys);
      d.rows[r] = reinterpret_cast<uint32_t*>((char*)x->peer[r] + s.rows);
      dc.cnt[r] = reinterpret_cast<u64*>((char*)x->peer[r] + s.cnt);
    }
    d.log_world = x->log_world; d.me = x->rank;
    d.row_add = side ? (x->row_shift ? (uint32_t)x->rank << x->row_shift : 0u) : (uint32_t)right_row0;
    const JKeyCol col{v.data, v.nulls, v.dtype};
    if (v.len > 0) {
      const int ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 2, (v.len + JQ_TILE - 1) / JQ_TILE));
      jpart1_kernel<true><<<ctas, JQ_NT, JQ_SMEM, c->stream>>>(col, v.len, x->sh_log_nb, s.cap, cursor, d, ovf);
      c->stats.kernel_launches++;
    }
    jx_publish_counts_kernel<<<(ncb + 255) / 256, 256, 0, c->stream>>>(cursor, dc, x->sh_log_nb, x->log_world, x->rank, s.cap);
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_t1, c->stream));
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 8, ovf, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->opt_timing) { PDRS_CUDA(c, cudaEventElapsedTime(&c->stats.total_ms, c->ev_t0, c->ev_t1)); c->stats.main_kernel_ms = c->stats.total_ms; }
  x->shuffled = true;
  if (c->pinned_scalars[8] != 0) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_xjoin_shuffle: a sub-bucket overflowed its padded range (skewed keys)");
  return PDRS_OK;
}

// Local join of the received rows (after the barrier): same build / probe kernels as pdrs_join_pairs on the radix
// path.  left_row0[world] = global number of the first left row of every rank.  Pairs are in GLOBAL row numbers;
// this rank returns the pairs of the keys whose rank hash maps to it.  Inner and Left only.
}  // extern "C"
// Build (unless the table of this build side exists already) and probe; the pairs are appended to `res`.
int32_t pdrs_xjoin_local_append(pdrs_xjoin* x, int32_t how, const int64_t* left_row0, pdrs_join_result* res, int64_t cap_hint) {
  pdrs_ctx* c = x->ctx;
  if (how != PDRS_INNER && how != PDRS_LEFT) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_xjoin_local: only Inner and Left joins are sharded");
  if (!x->shuffled) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_xjoin_local: nothing was shuffled");
  PDRS_CUDA(c, cudaSetDevice(c->device));
  pdrs_settle_frees(c);
  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
  c->stats.main_kernel_ms = 0; c->stats.total_ms = 0;
  const long long sb = 1ll << (x->sh_log_nb + x->log_world);
  char* base = (char*)x->base;
  // rows received per side
  std::vector<u64> hc((size_t)2 * sb);
  PDRS_CUDA(c, cudaMemcpyAsync(hc.data(), base + x->R.cnt, (size_t)sb * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaMemcpyAsync(hc.data() + sb, base + x->L.cnt, (size_t)sb * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  int64_t nr = 0, nl = 0;
  for (long long i = 0; i < sb; i++) { nr += (int64_t)hc[i]; nl += (int64_t)hc[sb + i]; }
  DevBuf row0;
  PDRS_TRY(row0.alloc(c, 64));
  PDRS_CUDA(c, cudaMemcpyAsync(row0.p, left_row0, (size_t)x->world * 8, cudaMemcpyHostToDevice, c->stream));
  JSrc rsrc{}, lsrc{};
  rsrc.pkeys = reinterpret_cast<const u64*>(base + x->R.keys); rsrc.prows = reinterpret_cast<const uint32_t*>(base + x->R.rows);
  rsrc.cap = x->R.cap; rsrc.cnt = reinterpret_cast<const u64*>(base + x->R.cnt);
  lsrc.pkeys = reinterpret_cast<const u64*>(base + x->L.keys); lsrc.prows = reinterpret_cast<const uint32_t*>(base + x->L.rows);
  lsrc.cap = x->L.cap; lsrc.cnt = reinterpret_cast<const u64*>(base + x->L.cnt);
  lsrc.row0 = row0.as<long long>(); lsrc.log_srcs = x->log_world; lsrc.row_shift = x->row_shift;
  long long nr_eff = sb * x->R.cap, nl_eff = sb * x->L.cap;
  const bool staged = x->row_shift && x->log_nb > 1;
  // staged mode: the ordinary one-pass radix partition of the single-GPU join, reading the received records
  auto partition_side = [&](int side, JPart& part, DevBuf& cnt) -> int32_t {
    const XSide& s = side ? x->L : x->R;
    JSrc& src = side ? lsrc : rsrc;
    const JStaged st{src.pkeys, src.prows, src.cnt, s.cap};
    long long cap = 0;
    bool ok = true;
    PDRS_TRY(jpartition1(c, JKeyCol{}, sb * s.cap, x->log_nb, &part, &cnt, &cap, &ok, &st, side ? nl : nr));
    if (!ok) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_xjoin_local: a radix bucket overflowed its padded range (skewed keys)");
    src.pkeys = part.keys.as<u64>(); src.prows = part.rows.as<uint32_t>(); src.cap = cap; src.cnt = cnt.as<u64>();
    src.log_nb = c->opt_join_prefetch ? x->log_nb : 0;
    (side ? nl_eff : nr_eff) = part.n;
    return PDRS_OK;
  };
  int64_t M = 0;
  JPart rp;               // (alive until the end of the call: a multi-GB free in the middle of it would slow the allocations that follow)
  DevBuf rcnt;
  if (!x->table_ready) {
    const long long slots = (std::max<long long>(1024, (c->opt_join_slots_mult > 0 ? c->opt_join_slots_mult : 3) * nr) + 3) / 4 * 4;
    if (slots >= (1ll << 32) - 8) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "join: build side too large (%lld rows)", (long long)nr);
    const size_t table_bytes = (size_t)(slots + 4) * 12;
    // the table of the previous call is reused when it is large enough (a free immediately followed by an allocation of the same
    // multi-GB size is exactly what the pool handles badly, see pdrs_settle_frees)
    if (x->tab.bytes < table_bytes || x->tab.bytes > table_bytes + table_bytes / 4) {
      x->tab.release();
      pdrs_settle_frees(c);
      PDRS_TRY(x->tab.alloc(c, table_bytes));
    }
    PDRS_CUDA(c, cudaMemsetAsync(x->tab.p, 0xFF, table_bytes, c->stream));
    x->jt = JTab{x->tab.as<u64>(), reinterpret_cast<uint32_t*>(x->tab.as<u64>() + slots + 4), (u64)slots, 1u};
    if (!x->failb.p) PDRS_TRY(x->failb.alloc(c, 64));
    PDRS_CUDA(c, cudaMemsetAsync(x->failb.p, 0, 64, c->stream));
    x->bst.csr.release();
    x->bst.dups = false;
    pdrs_settle_frees(c);
    if (staged) PDRS_TRY(partition_side(0, rp, rcnt));
    PDRS_TRY(jbuild_probe(c, x->jt, x->failb, rsrc, nr_eff, nr, lsrc, 0, 0, true, how, res, &M, [](const char*) {}, 1, &x->bst));
    x->table_ready = true;
    x->nr_built = nr;
  }
  c->stats.table_slots = (int64_t)x->jt.slots;
  JPart lp;
  DevBuf lcnt;
  if (staged) PDRS_TRY(partition_side(1, lp, lcnt));
  if (!x->bst.dups || res->n == 0) {
    PDRS_TRY(jbuild_probe(c, x->jt, x->failb, rsrc, 0, x->nr_built, lsrc, nl_eff, nl, true, how, res, &M, [](const char*) {}, 2, &x->bst, cap_hint));
  } else {
    // duplicate build keys take the count / scan / write path, which sizes its own arrays: probe into a temporary, then append
    pdrs_join_result tmp;
    tmp.ctx = c;
    PDRS_TRY(jbuild_probe(c, x->jt, x->failb, rsrc, 0, x->nr_built, lsrc, nl_eff, nl, true, how, &tmp, &M, [](const char*) {}, 2, &x->bst));
    const int64_t need = res->n + tmp.n;
    DevBuf nl_buf, nr_buf;
    PDRS_TRY(nl_buf.alloc(c, (size_t)std::max<int64_t>(need, 1) * 8));
    PDRS_TRY(nr_buf.alloc(c, (size_t)std::max<int64_t>(need, 1) * 8));
    if (res->n) {
      PDRS_CUDA(c, cudaMemcpyAsync(nl_buf.p, res->left.p, (size_t)res->n * 8, cudaMemcpyDeviceToDevice, c->stream));
      PDRS_CUDA(c, cudaMemcpyAsync(nr_buf.p, res->right.p, (size_t)res->n * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (tmp.n) {
      PDRS_CUDA(c, cudaMemcpyAsync(nl_buf.as<int64_t>() + res->n, tmp.left.p, (size_t)tmp.n * 8, cudaMemcpyDeviceToDevice, c->stream));
      PDRS_CUDA(c, cudaMemcpyAsync(nr_buf.as<int64_t>() + res->n, tmp.right.p, (size_t)tmp.n * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    res->left = std::move(nl_buf); res->right = std::move(nr_buf); res->n = need; res->cap = need;
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  c->stats.groupby_algo_used = 2;
  if (c->opt_timing) {
    PDRS_CUDA(c, cudaEventRecord(c->ev_t1, c->stream));
    PDRS_CUDA(c, cudaEventSynchronize(c->ev_t1));
    PDRS_CUDA(c, cudaEventElapsedTime(&c->stats.total_ms, c->ev_t0, c->ev_t1));
  } else {
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return PDRS_OK;
}
extern "C" {

int32_t pdrs_xjoin_local(pdrs_xjoin* x, int32_t how, const int64_t* left_row0, pdrs_join_result** out) {
  if (!x || !out || !left_row0) return PDRS_ERR_BAD_ARG;
  auto* res = new pdrs_join_result();
  res->ctx = x->ctx;
  const int32_t rc = pdrs_xjoin_local_append(x, how, left_row0, res, 0);
  if (rc != PDRS_OK) { delete res; return rc; }
  *out = res;
  return PDRS_OK;
}

void pdrs_xjoin_destroy(pdrs_xjoin* x) {
  if (!x) return;
  cudaSetDevice(x->ctx->device);
  cudaStreamSynchronize(x->ctx->stream);
  for (int r = 0; r < x->world; r++) if (x->ipc_open[r]) cudaIpcCloseMemHandle(x->peer[r]);
  if (x->base) cudaFree(x->base);
  delete x;
}

int64_t pdrs_join_len(const pdrs_join_result* r) { return r ? r->n : -1; }
int32_t pdrs_join_indices(const pdrs_join_result* r, int64_t* left_out, int64_t* right_out) {
  if (!r) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = r->ctx;
  if (r->n == 0) return PDRS_OK;
  if (!left_out || !right_out) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_join_indices: NULL output");
  PDRS_TRY(pdrs_copy_to_host(c, left_out, r->left.p, (size_t)r->n * 8));
  PDRS_TRY(pdrs_copy_to_host(c, right_out, r->right.p, (size_t)r->n * 8));
  return PDRS_OK;
}
int32_t pdrs_join_right_col(const pdrs_join_result* r, int32_t k, void* out_host) {
  if (!r || k < 0 || k >= r->npay) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = r->ctx;
  if (r->n == 0) return PDRS_OK;
  if (!out_host) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_join_right_col: NULL output");
  const int esz = r->pay_dtype[k] == PDRS_BOOL_BITS ? 1 : pdrs_dtype_bytes(r->pay_dtype[k]);
  return pdrs_copy_to_host(c, out_host, r->pay[k].p, (size_t)r->n * esz);
}
const void* pdrs_join_right_col_dev(const pdrs_join_result* r, int32_t k) { return (r && k >= 0 && k < r->npay) ? r->pay[k].p : nullptr; }
const int64_t* pdrs_join_left_dev(const pdrs_join_result* r) { return r ? r->left.as<int64_t>() : nullptr; }
const int64_t* pdrs_join_right_dev(const pdrs_join_result* r) { return r ? r->right.as<int64_t>() : nullptr; }
void pdrs_join_result_free(pdrs_join_result* r) {
  if (!r) return;
  pdrs_ctx* c = r->ctx;
  cudaSetDevice(c->device);
  delete r;
  pdrs_settle_frees(c);
}

}  // extern "C"
