// Groupby-aggregate kernels for sm_100a.
//
// Replaces the two hot loops of the reference (grouping.rs:62-104 row -> group, aggregation.rs:507-742
// per-group per-aggregate gathers) with ONE streaming pass over the key and value columns:
//
//   gb_shared_kernel   low cardinality.  A CTA-shared open-addressing key table in shared memory maps a
//                      key to a dense group id; every WARP owns private accumulator planes (16-byte
//                      records, optionally replicated NG times across lane groups) that are updated with
//                      plain LDS.128/STS.128 read-modify-writes.  Two lanes of a warp that hit the same
//                      record are serialised by rank (__match_any_sync on the record index), so the
//                      per-row path has no atomics at all: 64-bit shared atomics cost ~2 cycles per lane
//                      (f64 add is a CAS loop) and that alone exceeds the per-row cycle budget.
//                      Rows are staged through registers with 128-bit loads, one unit (512 rows per warp)
//                      ahead of the unit being aggregated.  Keys that do not fit the CTA table spill to
//                      the global table.
//   gb_global_kernel   high cardinality: every row goes to a global open-addressing table; slots are
//                      claimed with one 64-bit CAS on the header word, accumulators are updated with
//                      no-return reductions (RED.ADD.F64 / RED.ADD.64 / RED.MAX.64).
//   gb_finalize_kernel compacts occupied slots into dense output arrays, decodes the packed keys and
//                      evaluates the aggregates with the reference's formulas (aggregation.rs:500-754).
//
// Per-group state is (rows, n, pivot, S1 = sum(x - pivot), S2 = sum((x - pivot)^2), min, max, isum); the
// pivot is the first finite value seen for the group, which keeps the one-pass variance as accurate as
// the reference's two-pass formula (aggregation.rs:881-903) without reading the column twice.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------- flags / layout
enum { GB_SUM = 0, GB_ALL = 1 };        // which statistics a pass maintains: {rows,n,sum} or everything
#define GB_MAX_WARPS 8

typedef unsigned long long u64;

static constexpr u64 GB_BUSY = 1ull << 63;
static constexpr u64 GB_FULL = 1ull << 62;
static constexpr u64 GB_CNT_MASK = GB_FULL - 1;
static constexpr u64 GB_PIV_X = 0x7FF8C0DEC0DE0001ull;   // pivot bits are stored XOR this (0 = unset)
static constexpr u64 GB_SIGN = 1ull << 63;

enum { CNT_NGROUPS = 0, CNT_OVERFLOW = 1, CNT_SPILLED = 2, CNT_OUT = 3, CNT_SPIN_FAIL = 4, CNT_KMINC = 5 /* ~(min key ^ SIGN) */, CNT_KMAX = 6 /* max key ^ SIGN */,
       CNT_SPILLBUF = 7 /* rows appended to the spill side buffer */, CNT_N = 8 };

struct KeyColDev {
  const void* data;
  const uint8_t* nulls;
  long long null_alias;
  long long offset;    // range compression of multi-key tuples: the part holds (value - offset) in `bits` bits (0 / natural width otherwise)
  int dtype, word, shift, bits;
  int nword, nshift;   // where the "part is NULL" flag lives (nword < 0: none)
};
struct KeySpec {
  KeyColDev c[PDRS_MAX_KEYS];
  int nkeys, nwords;
  int single_null;     // nkeys == 1: a NULL key goes to the dedicated NULL group (no flag bits)
};

struct GHdr { u64 key0; u64 rowsw; };   // rowsw: BUSY | FULL | row count
struct GState {
  u64 n, pivotx; double S1, S2;         // sector 0
  u64 mnc, mxo, isum, pad;              // sector 1 (mnc = ~ord(min), mxo = ord(max); 0 = unset)
};
struct GTable {
  GHdr* hdr;                     // [slots + 1]; index `slots` is the NULL-key group of single-key groupbys
  u64* kw1;                      // [slots + 1] when nwords > 1
  u64* kw2;                      // [slots + 1] when nwords > 2
  GState* st;                    // [slots + 1] state of the value column of this pass (may be NULL)
  u64 mask;                      // slots - 1 (slots is a power of two)
  int shift;                     // 64 - log2(slots): slot = hash >> shift, so that the radix buckets (top hash bits) own contiguous regions
  long long slots;
  u64* counters;                 // CNT_*
  // Skew fallback of the tile-sort kernel (one-word keys): rows whose key got no register accumulator are appended
  // here as (key word, value bits) instead of being applied to the table with ~6 atomics each; the host aggregates
  // the buffer - the tail of the distribution, hot keys removed - in a second pass.  NULL when unused.
  u64* spill_k; u64* spill_v; long long spill_cap;
};

struct GbParams {
  KeySpec ks;
  const void* val;               // value column of this pass (NULL: count only)
  const uint8_t* vnull;
  const uint8_t* fbits;          // optional row filter (BOOL_BITS) and its null bitmap
  const uint8_t* fnull;
  long long n;
  GTable gt;
  int count_rows;                // this pass owns the group row counts
  int compat_nulls;              // filter present + compat_filter_nulls: NULL values count as 0 (data_ops.rs:64-71)
  int sh_cap, sh_slots, sh_log_slots, sh_ng;   // shared-memory kernel geometry
  int sh_dense;                  // single I64 key: id = key - sh_dense_base when 0 <= id < sh_cap (no key table)
  long long sh_dense_base;
  // tile-sort kernel over hash-partitioned rows (gb_tsort.cu): partition q owns rows [q * part_cap, q * part_cap + part_cnt[q])
  // of part_keys / part_vals (no NULLs, filter already applied); part_bits = log2(number of partitions)
  const u64* part_keys; const u64* part_vals; const uint8_t* part_flags /* 1 = value is NULL; may be NULL */; const u64* part_cnt; long long part_cap; int part_bits;
  int part_n;                       // number of partitions (2^part_bits hash partitions + the chunks of the overflow side area)
  int part_cpp, part_chunk_tiles;   // work items: every partition is cut into part_cpp chunks of part_chunk_tiles tiles (one flush per chunk)
  int ts_heavy;                     // tile-sort kernel: segments longer than this are reduced by the whole warp
  int ts_team;                      // tile-sort kernel: use the variant with the team-of-8 reduce
  int ts_mid;                       // ... longer than this (and up to ts_heavy) by a team of 8 lanes, shorter ones by their owner thread
  int ts_generic;                   // tile-sort kernel: keys are generic tuples packed into one word (not one Int64 column)
  // Hot keys of a skewed distribution (the keys the cardinality sample met >= 6 times; at most 1800): an open-addressing set of
  // packed key words (all ones = empty).  The partitioned path routes their rows around the hash partitions (gb_part.cu).
  const uint32_t* hot_tab; int hot_log_slots; float hot_frac;      // hot_frac = share of the sampled rows that carry a hot key
};
// The set holds a 32-bit TAG per key (0 = empty), not the key: a false positive (2^-31 per probe) only sends a row of a cold key
// to the side area, which aggregates any key - 4 bytes per slot keep the partition kernel at two CTAs per SM.
#define GB_HOT_LOG_SLOTS 12
__host__ __device__ __forceinline__ uint32_t gb_hot_slot(u64 k, int log_slots) {
  const uint32_t lo = (uint32_t)k ^ (uint32_t)(k >> 32);
  return (lo * 0x9E3779B1u + (uint32_t)(k >> 32) * 0x85EBCA6Bu) >> (32 - log_slots);
}
__host__ __device__ __forceinline__ uint32_t gb_hot_tag(u64 k) {
  const uint32_t lo = (uint32_t)k, hi = (uint32_t)(k >> 32);
  return ((lo * 0xC2B2AE35u) ^ (hi * 0x27D4EB2Fu) ^ (lo >> 15)) | 1u;
}

// ---------------------------------------------------------------- small device helpers
// Loads of words that other threads publish (slot headers, keys, pivots): volatile, so that a spin on a
// BUSY slot really re-reads memory (a plain __ldcg may legally be hoisted out of the retry loop).
__device__ __forceinline__ u64 ld_cg_u64(const u64* p) {
  u64 r;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ ulonglong2 ld_cg_hdr(const GHdr* p) {
  ulonglong2 r;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
  return r;
}

// streaming 128-bit load that the compiler may not sink below the aggregation of the previous unit
__device__ __forceinline__ ulonglong2 ld_stream_v2(const void* p) {
  ulonglong2 r;
  asm volatile("ld.global.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ ulonglong2 lds_volatile_v2(const void* p) {
  ulonglong2 r;
  asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return r;
}

template <int NW>
__device__ __forceinline__ u64 key_hash(const u64 (&w)[NW]) {
  u64 h = w[0] * 0x9E3779B97F4A7C15ull;
  if (NW > 1) h += (w[1] ^ (w[1] >> 29)) * 0xC2B2AE3D27D4EB4Full;
  if (NW > 2) h += (w[2] ^ (w[2] >> 31)) * 0x165667B19E3779F9ull;
  h ^= h >> 32;
  return h * 0xD6E8FEB86659FD93ull;
}

// 64 bitmap bits of rows [64*chunk, 64*chunk+64).  Bitmaps handed to the kernels cover ceil(n/8) bytes
// rounded up to 8 and are 8-byte aligned (pdrs_view_col guarantees both).
__device__ __forceinline__ u64 load_bits64(const uint8_t* bits, long long chunk) {
  return __ldg(reinterpret_cast<const u64*>(bits) + chunk);
}

// Packs the key tuple of one row.  Returns true when the row belongs to the dedicated NULL group.
template <int NW>
__device__ __noinline__ bool load_key_generic(const KeySpec& ks, long long row, u64 (&w)[NW]) {
#pragma unroll
  for (int i = 0; i < NW; i++) w[i] = 0;
  for (int k = 0; k < ks.nkeys; k++) {
    const KeyColDev& c = ks.c[k];
    bool isnull = c.nulls && pdrs_bit(c.nulls, row);
    u64 v = 0;
    if (!isnull) {
      switch (c.dtype) {
        case PDRS_I64: v = (u64)__ldg((const long long*)c.data + row) - (u64)c.offset; break;
        case PDRS_F64: {
          double d = __ldg((const double*)c.data + row);
          v = (d != d) ? 0x7FF8000000000000ull : (u64)__double_as_longlong(d);   // all NaNs print "NaN"
          break;
        }
        case PDRS_I32: v = ((u64)(long long)__ldg((const int*)c.data + row) - (u64)c.offset) & 0xFFFFFFFFull; break;
        case PDRS_DICT_U32: {
          uint32_t id = __ldg((const uint32_t*)c.data + row);
          if ((long long)id == c.null_alias) isnull = true; else v = (u64)id - (u64)c.offset;
          break;
        }
        case PDRS_BOOL_BITS: v = pdrs_bit((const uint8_t*)c.data, row); break;
      }
    }
    if (isnull) {
      if (ks.single_null) return true;
      if (c.nword >= 0) {
#pragma unroll
        for (int i = 0; i < NW; i++) if (i == c.nword) w[i] |= 1ull << c.nshift;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NW; i++) if (i == c.word) w[i] |= v << c.shift;
    }
  }
  return false;
}

// The same packing, fully inlined with a compile-time bound on the key loop, for kernels whose parameter block must stay in the
// constant bank: load_key_generic() takes the KeySpec by ADDRESS, and a kernel that passes (part of) its parameter struct by
// address gets the whole struct copied to local memory and read back with LDL in its hot loop.
template <int NW>
__device__ __forceinline__ bool load_key_inline(const KeySpec& ks, long long row, u64 (&w)[NW]) {
#pragma unroll
  for (int i = 0; i < NW; i++) w[i] = 0;
  bool nullgroup = false;
#pragma unroll
  for (int k = 0; k < PDRS_MAX_KEYS; k++) {
    if (k >= ks.nkeys) break;
    const KeyColDev& c = ks.c[k];
    bool isnull = c.nulls && pdrs_bit(c.nulls, row);
    u64 v = 0;
    if (!isnull) {
      switch (c.dtype) {
        case PDRS_I64: v = (u64)__ldg((const long long*)c.data + row) - (u64)c.offset; break;
        case PDRS_F64: { const double d = __ldg((const double*)c.data + row); v = (d != d) ? 0x7FF8000000000000ull : (u64)__double_as_longlong(d); break; }
        case PDRS_I32: v = ((u64)(long long)__ldg((const int*)c.data + row) - (u64)c.offset) & 0xFFFFFFFFull; break;
        case PDRS_DICT_U32: { const uint32_t id = __ldg((const uint32_t*)c.data + row); if ((long long)id == c.null_alias) isnull = true; else v = (u64)id - (u64)c.offset; break; }
        default: v = pdrs_bit((const uint8_t*)c.data, row); break;
      }
    }
    if (isnull) {
      if (ks.single_null) nullgroup = true;
      else if (c.nword >= 0) {
#pragma unroll
        for (int i = 0; i < NW; i++) if (i == c.nword) w[i] |= 1ull << c.nshift;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NW; i++) if (i == c.word) w[i] |= v << c.shift;
    }
  }
  return nullgroup;
}

// ---------------------------------------------------------------- global table
// Slots are claimed with one CAS on the header word (0 -> BUSY), then the key words are written and the
// header is published (FULL).  A thread that meets a BUSY slot must NOT spin on it: the publisher may be a
// lane of the same warp, and a lane spinning inside a divergent loop can starve it.  So a probe attempt
// returns G_RETRY instead, and the retry loop is warp-synchronous (every iteration reconverges at
// __any_sync, which lets the publisher finish).  All 32 lanes of a warp must call g_find_or_insert together.
static constexpr long long G_RETRY = -2;

template <int NW>
__device__ __forceinline__ long long g_try_insert(const GTable& t, const u64 (&w)[NW], u64& slot, u64& probe) {
  while (probe <= t.mask) {
    ulonglong2 h = ld_cg_hdr(&t.hdr[slot]);
    if (h.y == 0) {
      u64 old = atomicCAS(&t.hdr[slot].rowsw, 0ull, GB_BUSY);
      if (old == 0) {
        t.hdr[slot].key0 = w[0];
        if (NW > 1) t.kw1[slot] = w[1];
        if (NW > 2) t.kw2[slot] = w[2];
        __threadfence();
        atomicExch(&t.hdr[slot].rowsw, GB_FULL);
        atomicAdd(&t.counters[CNT_NGROUPS], 1ull);
        return (long long)slot;
      }
      h.y = old;
      if (!(old & GB_BUSY)) h.x = ld_cg_u64(&t.hdr[slot].key0);
    }
    if (h.y & GB_BUSY) return G_RETRY;     // somebody is publishing this slot: look again next round
    bool match = h.x == w[0];
    if (NW > 1) match = match && ld_cg_u64(&t.kw1[slot]) == w[1];
    if (NW > 2) match = match && ld_cg_u64(&t.kw2[slot]) == w[2];
    if (match) return (long long)slot;
    slot = (slot + 1) & t.mask;
    probe++;
  }
  atomicAdd(&t.counters[CNT_OVERFLOW], 1ull);
  return -1;
}

// Returns the slot of the key (claiming a free one if needed), or -1 on overflow / for inactive lanes.
template <int NW>
__device__ __forceinline__ long long g_find_or_insert(const GTable& t, const u64 (&w)[NW], bool active) {
  u64 slot = key_hash<NW>(w) >> t.shift, probe = 0;
  long long res = -1;
  bool pending = active;
  int rounds = 0;
  {
    // Fail fast when the table fills up (cardinality underestimated by the sample, e.g. Zipf-skewed keys): linear
    // probing degenerates into scans of the whole table long before it is full, and the attempt is thrown away by
    // the host anyway (it retries with 4x the slots).  One 16-byte load of {groups, overflow flag} per call.
    const ulonglong2 c01 = ld_cg_hdr(reinterpret_cast<const GHdr*>(t.counters + CNT_NGROUPS));
    const bool full = c01.y != 0 || c01.x > (u64)t.slots - ((u64)t.slots >> 2);
    if (__any_sync(0xFFFFFFFFu, full)) {          // warp-uniform decision: the retry loop below is warp-synchronous
      if (full && c01.y == 0) atomicAdd(&t.counters[CNT_OVERFLOW], 1ull);   // (also by lanes without a row: rows of this warp are dropped)
      return -1;
    }
  }
  while (__any_sync(0xFFFFFFFFu, pending)) {
    if (pending) {
      long long r = g_try_insert<NW>(t, w, slot, probe);
      if (r != G_RETRY) { res = r; pending = false; }
    }
    if (++rounds > (1 << 22)) { if (pending) atomicAdd(&t.counters[CNT_SPIN_FAIL], 1ull); break; }
  }
  return res;
}

// Global pivot of a group: first caller sets it, everybody gets the same value back.
__device__ __forceinline__ double g_pivot(GState* s, double candidate) {
  u64 px = ld_cg_u64(&s->pivotx);
  if (px == 0) {
    u64 mine = (u64)__double_as_longlong(candidate) ^ GB_PIV_X;
    u64 old = atomicCAS(&s->pivotx, 0ull, mine);
    px = old ? old : mine;
  }
  return __longlong_as_double((long long)(px ^ GB_PIV_X));
}

template <typename VT> struct ValTraits;
template <> struct ValTraits<double> {
  static constexpr bool is_int = false;
  __device__ static __forceinline__ double to_f64(double v) { return v; }
  __device__ static __forceinline__ u64 ord(double v) { return pdrs_ord_f64(v); }
  __device__ static __forceinline__ bool orderable(double v) { return v == v; }   // NaN is ignored by f64::min/max
  __device__ static __forceinline__ double min_init() { return __longlong_as_double(0x7FF0000000000000ll); }
  __device__ static __forceinline__ double max_init() { return __longlong_as_double((long long)0xFFF0000000000000ull); }
  __device__ static __forceinline__ double from_bits(u64 b) { return __longlong_as_double((long long)b); }
  __device__ static __forceinline__ u64 to_bits(double v) { return (u64)__double_as_longlong(v); }
};
template <> struct ValTraits<long long> {
  static constexpr bool is_int = true;
  __device__ static __forceinline__ double to_f64(long long v) { return (double)v; }
  __device__ static __forceinline__ u64 ord(long long v) { return (u64)v ^ GB_SIGN; }
  __device__ static __forceinline__ bool orderable(long long) { return true; }
  __device__ static __forceinline__ long long min_init() { return 0x7FFFFFFFFFFFFFFFll; }
  __device__ static __forceinline__ long long max_init() { return (long long)0x8000000000000000ull; }
  __device__ static __forceinline__ long long from_bits(u64 b) { return (long long)b; }
  __device__ static __forceinline__ u64 to_bits(long long v) { return (u64)v; }
};

__device__ __forceinline__ bool is_finite_f64(double d) { return fabs(d) <= 1.7976931348623157e308; }

// One row straight into the global table (high-cardinality path and shared-table spills).
template <typename VT, int FLAGS>
__device__ __forceinline__ void g_update_row(const GTable& t, long long slot, bool count_row, bool valid, VT v) {
  using T = ValTraits<VT>;
  if (count_row) atomicAdd(&t.hdr[slot].rowsw, 1ull);
  if (!valid || !t.st) return;
  GState* s = &t.st[slot];
  atomicAdd(&s->n, 1ull);
  double x = T::to_f64(v);
  if (T::is_int) atomicAdd(&s->isum, (u64)v);
  if (FLAGS == GB_ALL || !T::is_int) {
    double c = 0.0;
    if (FLAGS == GB_ALL) {
      if (is_finite_f64(x)) c = g_pivot(s, x);
      else { u64 px = ld_cg_u64(&s->pivotx); c = px ? __longlong_as_double((long long)(px ^ GB_PIV_X)) : 0.0; }
    }
    double d = x - c;
    atomicAdd(&s->S1, d);
    if (FLAGS == GB_ALL) atomicAdd(&s->S2, d * d);
  }
  if (FLAGS == GB_ALL && T::orderable(v)) {
    u64 o = T::ord(v);
    if (~o > ld_cg_u64(&s->mnc)) atomicMax(&s->mnc, ~o);
    if (o > ld_cg_u64(&s->mxo)) atomicMax(&s->mxo, o);
  }
}

// A pre-aggregated batch (n values with pivot c, S1, S2 relative to c) into the global table.
template <int FLAGS, bool IS_INT>
__device__ __forceinline__ void g_update_batch(const GTable& t, long long slot, u64 rows, u64 n, double c, bool have_c,
                                               double S1, double S2, u64 isum, u64 mnc, u64 mxo) {
  if (rows) atomicAdd(&t.hdr[slot].rowsw, rows);
  if (!t.st || n == 0) return;
  GState* s = &t.st[slot];
  atomicAdd(&s->n, n);
  if (IS_INT) atomicAdd(&s->isum, isum);
  if (FLAGS == GB_ALL || !IS_INT) {
    double a1 = S1, a2 = S2;
    if (FLAGS == GB_ALL) {
      double C;
      if (have_c) C = g_pivot(s, c);
      else { u64 px = ld_cg_u64(&s->pivotx); C = px ? __longlong_as_double((long long)(px ^ GB_PIV_X)) : 0.0; c = 0.0; }   // batch of non-finite values only
      double dl = c - C, nn = (double)n;                 // re-base the batch from its pivot c to the group pivot C
      a1 = S1 + nn * dl;
      a2 = S2 + 2.0 * dl * S1 + nn * dl * dl;
    }
    atomicAdd(&s->S1, a1);
    if (FLAGS == GB_ALL) atomicAdd(&s->S2, a2);
  }
  if (FLAGS == GB_ALL) {
    if (mnc && mnc > ld_cg_u64(&s->mnc)) atomicMax(&s->mnc, mnc);
    if (mxo && mxo > ld_cg_u64(&s->mxo)) atomicMax(&s->mxo, mxo);
  }
}

// ---------------------------------------------------------------- shared-memory layout
// Measured on B200 (tools/microbench2.cu, profiles/microbench_r01.md): a random shared-memory access costs
// ~2.6 SM-cycles per warp instruction per 4 bytes (LDS.32 2.6, LDS.64 5.0, LDS.128 9.3, same for STS),
// a native 32-bit shared atomic costs the same 2.6, 64-bit / f64 shared atomics are CAS loops (18-31) and
// MATCH.ANY costs 45.  The per-row shared-memory footprint is therefore kept minimal:
//
// CTA-shared   key table u64[NW][S] + id table u32[S] (0 = empty, SH_BUSY, else id + 1)   (not in dense mode)
//              META[cap+1]  {pivot^tag, f32 min bound, f32 max bound}   read-mostly, one LDS.128 per row  (ALL)
//              EXACT[cap+1] {~ord(min), ord(max)}                        touched only when a bound is beaten (ALL)
//              NUL[cap+1]   u32 NULL values (n = rows - nulls), native atomic, ~5% of the rows
//              misc u32[4]
// per warp     ACC[E]  f64 {S1}, i64 {isum}                       (SUM)        E = (cap + 1) * ng
//                      {S1, S2} (+ ISUM[E] for i64 values)        (ALL)
//              CNT[E]  u32 rows: bumped with the native atomic; the returned ticket ranks the lanes of the
//                      warp that hit the same record in this batch (re-read after __syncwarp gives the count)
// S1 / S2 are plain LDS/STS read-modify-writes, one rank per round, so the per-row path has no 64-bit atomics.
template <typename VT, int FLAGS> struct ShPlanes {
  static constexpr bool IS_INT = ValTraits<VT>::is_int;
  static constexpr int ACC_BYTES = FLAGS == GB_SUM ? 8 : 16;
  static constexpr int REC_BYTES = ACC_BYTES + ((FLAGS == GB_ALL && IS_INT) ? 8 : 0) + 4;   // per warp per record: sums + CNT
  static constexpr int CTA_BYTES = (FLAGS == GB_ALL ? 32 : 0) + 4;                            // per group: [META + EXACT] + NUL
};
__host__ __device__ inline size_t gb_sh_fixed_bytes(int nw, int slots, int cap, int cta_bytes, int dense) {
  size_t b = dense ? 0 : ((size_t)8 * nw * slots + (size_t)4 * slots);
  b = (b + 15) / 16 * 16;
  b += ((size_t)cta_bytes * (cap + 1) + 15) / 16 * 16;
  return b + 16;
}
// E = (cap + 1) * ng records + 1 trash record (lanes with nothing to count bump that one: no branches)
__host__ __device__ inline size_t gb_sh_warp_bytes(int rec_bytes, int cap, int ng) {
  const size_t E1 = (size_t)(cap + 1) * ng + 1;
  return ((size_t)rec_bytes * E1 + 15) / 16 * 16;
}

static constexpr uint32_t SH_BUSY = 0xFFFFFFFFu;

// key -> dense id through the CTA-shared key table; -1 = table full (spill).  Same no-spin protocol as the
// global table: all 32 lanes of a warp call sh_lookup together.
static constexpr int SH_RETRY = -2;
template <int NW>
__device__ __forceinline__ int sh_try(u64* ktab_key, uint32_t* ktab_id, uint32_t* misc, int S, int cap, const u64 (&w)[NW], uint32_t& slot, int& probe) {
  while (probe < S) {
    // id first, key second (the publisher writes key -> fence -> id), both loads in flight together
    uint32_t idw = *reinterpret_cast<volatile uint32_t*>(&ktab_id[slot]);
    u64 kk[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) kk[i] = *reinterpret_cast<volatile u64*>(&ktab_key[i * S + slot]);
    if (idw == 0) {
      if (*reinterpret_cast<volatile uint32_t*>(&misc[0]) >= (uint32_t)cap) return -1;
      uint32_t old = atomicCAS(&ktab_id[slot], 0u, SH_BUSY);
      if (old == 0) {
        uint32_t nid = atomicAdd(&misc[0], 1u);
        if (nid >= (uint32_t)cap) {           // table is full: give the slot back, spill the row
          *reinterpret_cast<volatile uint32_t*>(&ktab_id[slot]) = 0;
          return -1;
        }
#pragma unroll
        for (int i = 0; i < NW; i++) *reinterpret_cast<volatile u64*>(&ktab_key[i * S + slot]) = w[i];
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t*>(&ktab_id[slot]) = nid + 1;
        return (int)nid;
      }
      return SH_RETRY;   // lost the race: the winner is publishing (or gave the slot back), look again
    }
    if (idw == SH_BUSY) return SH_RETRY;
    bool match = true;
#pragma unroll
    for (int i = 0; i < NW; i++) match = match && (kk[i] == w[i]);
    if (match) return (int)idw - 1;
    slot = (slot + 1) & (S - 1);
    probe++;
  }
  return -1;
}
template <int NW>
__device__ __noinline__ int sh_lookup(u64* ktab_key, uint32_t* ktab_id, uint32_t* misc, int S, int log_slots, int cap, const u64 (&w)[NW], bool active) {
  uint32_t slot = (uint32_t)(key_hash<NW>(w) >> (64 - log_slots));
  int probe = 0, res = -1, rounds = 0;
  bool pending = active;
  while (__any_sync(0xFFFFFFFFu, pending)) {
    if (pending) {
      int r = sh_try<NW>(ktab_key, ktab_id, misc, S, cap, w, slot, probe);
      if (r != SH_RETRY) { res = r; pending = false; }
    }
    if (++rounds > (1 << 20)) break;   // res stays -1: the row spills
  }
  return res;
}

// One unit of a warp: 256 rows = GB_Q batches of 32; lane owns rows base + 64*j + 2*lane + h, batch q = 2*j + h.
// Bitmap words are kept raw (the u32 word holding this lane's two bits of each 64-row chunk) and decoded
// when the unit is aggregated, so that loading the NEXT unit never waits on a load.
#define GB_Q 8
#define GB_UNIT_ROWS (32 * GB_Q)
template <typename VT> struct GbUnit {
  u64 k[GB_Q];                 // KM != 1 only: the 64-bit key column
  u64 v[GB_Q];                 // value bits
  uint32_t vnw[GB_Q / 2];      // NULL-value bitmap words
  uint32_t knw[GB_Q / 2];      // NULL-key bitmap words (KM != 1)
  uint32_t fw[GB_Q / 2];       // filter: value & ~null
  uint32_t act;                // partial unit only: bit q = row is in range
};

template <int KM, bool FULL, typename VT>
__device__ __forceinline__ void gb_load_unit(const GbParams& p, long long base, int lane, GbUnit<VT>& u) {
  const long long n = p.n;
  const u64* keys = reinterpret_cast<const u64*>(p.ks.c[0].data) + base + 2 * lane;
  const u64* vals = reinterpret_cast<const u64*>(p.val) + base + 2 * lane;
  const long long w0 = (base >> 5) + (lane >> 4);     // u32 bitmap word of this lane in chunk 0
  u.act = (1u << GB_Q) - 1u;
  if (FULL) {
#pragma unroll
    for (int j = 0; j < GB_Q / 2; j++) {
      ulonglong2 kk = make_ulonglong2(0, 0), vv = make_ulonglong2(0, 0);
      if (KM != 1) kk = ld_stream_v2(keys + 64 * j);
      if (p.val) vv = ld_stream_v2(vals + 64 * j);
      u.k[2 * j] = kk.x; u.k[2 * j + 1] = kk.y;
      u.v[2 * j] = vv.x; u.v[2 * j + 1] = vv.y;
    }
  } else {                                     // the last, partial unit
    u.act = 0;
#pragma unroll
    for (int j = 0; j < GB_Q / 2; j++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const bool inb = base + 64 * j + 2 * lane + h < n;
        u.k[2 * j + h] = (KM != 1 && inb) ? __ldg(keys + 64 * j + h) : 0ull;
        u.v[2 * j + h] = (p.val && inb) ? __ldg(vals + 64 * j + h) : 0ull;
        if (inb) u.act |= 1u << (2 * j + h);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < GB_Q / 2; j++) {
    const bool inb = FULL || base + 64 * j < n;
    u.vnw[j] = (p.vnull && inb) ? __ldg(reinterpret_cast<const uint32_t*>(p.vnull) + w0 + 2 * j) : 0u;
    u.knw[j] = (KM != 1 && p.ks.c[0].nulls && inb) ? __ldg(reinterpret_cast<const uint32_t*>(p.ks.c[0].nulls) + w0 + 2 * j) : 0u;
    uint32_t f = 0xFFFFFFFFu;
    if (p.fbits) {   // filter: Some(true) rows only (data_ops.rs:49-55)
      f = inb ? __ldg(reinterpret_cast<const uint32_t*>(p.fbits) + w0 + 2 * j) : 0u;
      if (p.fnull && inb) f &= ~__ldg(reinterpret_cast<const uint32_t*>(p.fnull) + w0 + 2 * j);
    }
    u.fw[j] = f;
  }
}

// ---------------------------------------------------------------- the shared-memory kernel
// Rare paths live in __noinline__ functions so that the unrolled hot loop stays inside the instruction cache.
template <int NW, typename VT, int FLAGS>
__device__ __noinline__ void gb_spill_rows(const GTable gt, u64 w0, u64 w1, u64 w2, bool spill, bool count_row, bool valid, VT v) {
  if (NW == 1 && gt.spill_k) {     // rows with a value go to the side buffer (warp-aggregated append); NULL values only count: below
    const bool tobuf = spill && valid;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, tobuf);
    if (m) {
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      u64 at = 0;
      if (lane == leader) { at = atomicAdd(&gt.counters[CNT_SPILLBUF], (u64)__popc(m)); atomicAdd(&gt.counters[CNT_SPILLED], (u64)__popc(m)); }
      at = __shfl_sync(0xFFFFFFFFu, at, leader) + (u64)__popc(m & ((1u << lane) - 1u));
      if (tobuf && at < (u64)gt.spill_cap) { gt.spill_k[at] = w0; gt.spill_v[at] = ValTraits<VT>::to_bits(v); spill = false; }
    }
  }
  u64 w[NW];
  w[0] = w0;
  if (NW > 1) w[1] = w1;
  if (NW > 2) w[2] = w2;
  long long gs = g_find_or_insert<NW>(gt, w, spill);
  if (spill && gs >= 0) g_update_row<VT, FLAGS>(gt, gs, count_row, valid, v);
  if (spill) atomicAdd(&gt.counters[CNT_SPILLED], 1ull);
}

// A value reached the f32 bound of its group: update the exact minimum / maximum (64-bit shared atomicMax is a
// CAS loop, but this runs O(log rows) times per group) and refresh the bounds.
template <typename VT>
__device__ __noinline__ void gb_minmax_slow(ulonglong2* META, ulonglong2* EXACT, int id, VT v) {
  using T = ValTraits<VT>;
  if (!T::orderable(v)) return;
  {
    const u64 o = ~T::ord(v);
    const u64 was = atomicMax(&EXACT[id].x, o);
    const u64 best = ~(was > o ? was : o);     // ord() of the current exact minimum
    float b;
    if (T::is_int) b = __ll2float_rn((long long)(best ^ GB_SIGN)); else b = __double2float_rn(pdrs_unord_f64(best));
    reinterpret_cast<volatile uint32_t*>(&META[id].y)[0] = __float_as_uint(b);   // bound = rn(some earlier exact min) >= rn(exact min)
  }
  {
    const u64 o = T::ord(v);
    const u64 was = atomicMax(&EXACT[id].y, o);
    const u64 best = was > o ? was : o;
    float b;
    if (T::is_int) b = __ll2float_rn((long long)(best ^ GB_SIGN)); else b = __double2float_rn(pdrs_unord_f64(best));
    reinterpret_cast<volatile uint32_t*>(&META[id].y)[1] = __float_as_uint(b);   // bound = rn(some earlier exact max) <= rn(exact max)
  }
}

// The CTA pivot of a group is the first finite value with its mantissa LSB forced to 1 (any value near the data
// works as a pivot), so that 0 bits mean "unset" and the stored word IS the pivot: no decoding on the hot path.
static __device__ __noinline__ u64 gb_pivot_set(ulonglong2* META, int id, double x) {
  const u64 mine = (u64)__double_as_longlong(x) | 1ull;
  const u64 was = atomicCAS(&META[id].x, 0ull, mine);
  return was ? was : mine;
}

// ---- explicit shared-memory accessors on 32-bit shared addresses (keeps the address arithmetic of the hot
//      loop to one IMAD per access and the access width explicit)
__device__ __forceinline__ uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t sm_atom_add32(uint32_t a, uint32_t v) { uint32_t r; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(a), "r"(v) : "memory"); return r; }
__device__ __forceinline__ void sm_red_add32(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t sm_ld32(uint32_t a) { uint32_t r; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ u64 sm_ld64(uint32_t a) { u64 r; asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(r) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ void sm_st64(uint32_t a, u64 v) { asm volatile("st.volatile.shared.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ ulonglong2 sm_ld128(uint32_t a) { ulonglong2 r; asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ void sm_st128(uint32_t a, ulonglong2 v) { asm volatile("st.volatile.shared.v2.u64 [%0], {%1, %2};" :: "r"(a), "l"(v.x), "l"(v.y) : "memory"); }

// Everything the per-unit body needs (kept in one struct so that the body can be instantiated several times:
// full units in the main loop, the single partial unit after it).
template <typename VT, int FLAGS> struct GbShared {
  u64* ktab_key; uint32_t* ktab_id; uint32_t* misc;
  ulonglong2* META; ulonglong2* EXACT;
  uint32_t a_meta, a_nul, a_acc, a_isum, a_cnt;   // 32-bit shared addresses of META, NUL and this warp's ACC / ISUM / CNT
  int S, cap, NG, E, rep, lane;
  u64 dense_base;
};

// PLAIN: no row filter, no NULL keys, no compat_filter_nulls, full unit -> every row is active.
template <int NW, int KM, typename VT, int FLAGS, bool PLAIN>
__device__ __forceinline__ void gb_unit_body(const GbParams& p, const GbShared<VT, FLAGS>& sh, const GbUnit<VT>& cur, long long base) {
  using T = ValTraits<VT>;
  constexpr bool IS_INT = T::is_int;
  constexpr bool ALL = FLAGS == GB_ALL;
  constexpr bool dense = KM == 2;
  constexpr int ACCB = ALL ? 16 : 8;
  const int lane = sh.lane, sh2 = (2 * lane) & 31;
  uint32_t arec[GB_Q];     // rec * 4 (byte offset into CNT; ACC offset = arec * ACCB / 4)
  int ids[GB_Q];
  u64 vb[GB_Q];            // value bits (0 where compat_filter_nulls turns a NULL into a default)
  double dd[GB_Q];
  uint32_t old[GB_Q];
  uint32_t updmask = 0, spillmask = 0, slowmask = 0, nullkey = 0;
  // -- phase 1: key -> record, bump the row counters (tickets), pivot / min / max bounds
#pragma unroll
  for (int q = 0; q < GB_Q; q++) {
    const int j = q >> 1, h = q & 1;
    bool active = true, knull = false;
    bool vnull = !p.val || ((cur.vnw[j] >> (sh2 + h)) & 1u);
    vb[q] = cur.v[q];
    if (!PLAIN) {
      active = ((cur.act >> q) & 1u) && ((cur.fw[j] >> (sh2 + h)) & 1u);
      knull = (cur.knw[j] >> (sh2 + h)) & 1u;
      if (p.compat_nulls && vnull && p.val) { vnull = false; vb[q] = 0; }   // filter + compat_filter_nulls (data_ops.rs:64-71)
    }
    int id = -1;
    if (dense) {                       // small dense integer keys: direct-mapped
      const u64 off = cur.k[q] - sh.dense_base;
      if (active && off < (u64)sh.cap) id = (int)off;
    } else {                           // CTA-shared key table
      u64 w[NW];
#pragma unroll
      for (int i = 0; i < NW; i++) w[i] = 0;
      if (KM != 1) w[0] = cur.k[q];
      else if (active) knull = load_key_generic<NW>(p.ks, base + 64 * j + 2 * lane + h, w);
      id = sh_lookup<NW>(sh.ktab_key, sh.ktab_id, sh.misc, sh.S, p.sh_log_slots, sh.cap, w, active && !knull);
    }
    if (!PLAIN || KM == 1) { if (active && knull) { id = sh.cap; nullkey = 1; } }
    if (active && id < 0) spillmask |= 1u << q;
    ids[q] = id;
    arec[q] = (uint32_t)(id >= 0 ? id * sh.NG + sh.rep : sh.E) * 4u;
    old[q] = sm_atom_add32(sh.a_cnt + arec[q], 1u);
    if (id >= 0 && vnull) sm_red_add32(sh.a_nul + (uint32_t)id * 4u, 1u);
    const bool upd = id >= 0 && !vnull;
    if (upd) updmask |= 1u << q;
    const VT v = T::from_bits(vb[q]);
    const double x = T::to_f64(v);
    dd[q] = x;
    if (ALL) {                         // pivot and min / max bounds: CTA-shared, read-mostly, one LDS.128
      const ulonglong2 meta = sm_ld128(sh.a_meta + (uint32_t)(id >= 0 ? id : 0) * 16u);
      dd[q] = x - __longlong_as_double((long long)meta.x);   // unset pivot = 0 bits = +0.0
      // f32 bounds hold round-to-nearest(exact min / max); rounding is monotonic, so a value below the exact
      // minimum always satisfies xf <= bound (values equal to the bound after rounding take the slow path too)
      const float xf = IS_INT ? __ll2float_rn((long long)T::to_bits(v)) : __double2float_rn(x);
      const bool slow = xf <= __uint_as_float((uint32_t)meta.y) || xf >= __uint_as_float((uint32_t)(meta.y >> 32)) || !((uint32_t)meta.x & 1u);
      if (upd && slow) slowmask |= 1u << q;
    }
  }
  if (nullkey && !*reinterpret_cast<volatile uint32_t*>(&sh.misc[1])) *reinterpret_cast<volatile uint32_t*>(&sh.misc[1]) = 1u;
  if (ALL && slowmask) {               // rare: first value of a group (pivot) or a value at / beyond a bound
#pragma unroll
    for (int q = 0; q < GB_Q; q++) if ((slowmask >> q) & 1u) {
      const VT v = T::from_bits(vb[q]);
      const double x = T::to_f64(v);
      u64 px = *reinterpret_cast<volatile u64*>(&sh.META[ids[q]].x);
      if (px == 0 && is_finite_f64(x)) px = gb_pivot_set(sh.META, ids[q], x);   // first finite value of the group in this CTA
      dd[q] = x - __longlong_as_double((long long)px);
      gb_minmax_slow<VT>(sh.META, sh.EXACT, ids[q], v);
    }
  }
  if (__any_sync(0xFFFFFFFFu, spillmask != 0)) {   // rare: keys that do not fit this CTA's table go to the global table
#pragma unroll
    for (int q = 0; q < GB_Q; q++) {
      u64 w[NW];
#pragma unroll
      for (int i = 0; i < NW; i++) w[i] = 0;
      const bool sp = (spillmask >> q) & 1u;
      if (KM != 1) w[0] = cur.k[q];
      else if (sp) load_key_generic<NW>(p.ks, base + 64 * (q >> 1) + 2 * lane + (q & 1), w);
      const bool vnull = !p.val || (!p.compat_nulls && ((cur.vnw[q >> 1] >> (sh2 + (q & 1))) & 1u));
      gb_spill_rows<NW, VT, FLAGS>(p.gt, w[0], NW > 1 ? w[NW > 1 ? 1 : 0] : 0ull, NW > 2 ? w[NW > 2 ? 2 : 0] : 0ull, sp, p.count_rows != 0,
                                   !vnull, T::from_bits(vb[q]));
    }
  }
  // -- phase 2: ranks from the tickets (unique per record across the whole unit)
  __syncwarp();
  uint32_t latemask = 0;
#pragma unroll
  for (int q = 0; q < GB_Q; q++) {
    const uint32_t now = sm_ld32(sh.a_cnt + arec[q]);
    if (now - old[q] != 1u) latemask |= 1u << q;     // not the last ticket of its record
  }
  latemask &= updmask;
  // -- phase 3a: the row holding the LAST ticket of its record does a plain read-modify-write
  const uint32_t firstmask = updmask & ~latemask;
  if (!ALL) {
    u64 a[GB_Q];
#pragma unroll
    for (int q = 0; q < GB_Q; q++) if ((firstmask >> q) & 1u) a[q] = sm_ld64(sh.a_acc + arec[q] * 2u);
#pragma unroll
    for (int q = 0; q < GB_Q; q++) if ((firstmask >> q) & 1u) {
      if (IS_INT) a[q] += vb[q];
      else a[q] = (u64)__double_as_longlong(__longlong_as_double((long long)a[q]) + dd[q]);
      sm_st64(sh.a_acc + arec[q] * 2u, a[q]);
    }
  } else {
    ulonglong2 a[GB_Q];
#pragma unroll
    for (int q = 0; q < GB_Q; q++) if ((firstmask >> q) & 1u) a[q] = sm_ld128(sh.a_acc + arec[q] * 4u);
#pragma unroll
    for (int q = 0; q < GB_Q; q++) if ((firstmask >> q) & 1u) {
      a[q].x = (u64)__double_as_longlong(__longlong_as_double((long long)a[q].x) + dd[q]);
      a[q].y = (u64)__double_as_longlong(fma(dd[q], dd[q], __longlong_as_double((long long)a[q].y)));
      sm_st128(sh.a_acc + arec[q] * 4u, a[q]);
      if (IS_INT) sm_st64(sh.a_isum + arec[q] * 2u, sm_ld64(sh.a_isum + arec[q] * 2u) + vb[q]);
    }
  }
  // -- phase 3b: the other rows of a record (same warp, same unit: ~10% of the rows at 1000 groups) add with
  //    64-bit shared atomics (CAS loops) once the plain updates are done
  if (__any_sync(0xFFFFFFFFu, latemask != 0)) {
    __syncwarp();
    const char* accg = reinterpret_cast<const char*>(__cvta_shared_to_generic(sh.a_acc));
    const char* isumg = reinterpret_cast<const char*>(__cvta_shared_to_generic(sh.a_isum));
#pragma unroll
    for (int q = 0; q < GB_Q; q++) if ((latemask >> q) & 1u) {
      if (!ALL) {
        if (IS_INT) atomicAdd((u64*)(accg + arec[q] * 2u), vb[q]);
        else atomicAdd((double*)(accg + arec[q] * 2u), dd[q]);
      } else {
        atomicAdd((double*)(accg + arec[q] * 4u), dd[q]);
        atomicAdd((double*)(accg + arec[q] * 4u + 8), dd[q] * dd[q]);
        if (IS_INT) atomicAdd((u64*)(isumg + arec[q] * 2u), vb[q]);
      }
    }
  }
  __syncwarp();
  (void)ACCB;
}

// KM: 0 = one 64-bit key column through the CTA-shared key table, 1 = generic packed key tuple,
//     2 = one 64-bit key column with small dense integer keys (direct-mapped group ids, no key table)
template <int NW, int KM, typename VT, int FLAGS>
__global__ void __launch_bounds__(GB_MAX_WARPS * 32, 1) gb_shared_kernel(const GbParams p) {
  using T = ValTraits<VT>;
  using L = ShPlanes<VT, FLAGS>;
  constexpr bool IS_INT = T::is_int;
  constexpr bool ALL = FLAGS == GB_ALL;
  constexpr bool dense = KM == 2;
  extern __shared__ __align__(16) unsigned char smem[];
  const int S = p.sh_slots, cap = p.sh_cap, NG = p.sh_ng;
  const int E = (cap + 1) * NG;              // record E is the per-warp trash record
  u64* ktab_key = reinterpret_cast<u64*>(smem);
  uint32_t* ktab_id = reinterpret_cast<uint32_t*>(smem + (size_t)8 * NW * S);
  const size_t keys_bytes = dense ? 0 : (((size_t)8 * NW * S + (size_t)4 * S + 15) / 16 * 16);
  ulonglong2* META = reinterpret_cast<ulonglong2*>(smem + keys_bytes);                 // ALL only
  ulonglong2* EXACT = META + (cap + 1);                                               // ALL only
  uint32_t* NUL = reinterpret_cast<uint32_t*>(smem + keys_bytes + (size_t)(L::CTA_BYTES - 4) * (cap + 1));
  const size_t fixed = gb_sh_fixed_bytes(NW, S, cap, L::CTA_BYTES, dense);
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + fixed - 16);   // [0] groups in this CTA, [1] NULL group seen
  const size_t warp_bytes = gb_sh_warp_bytes(L::REC_BYTES, cap, NG);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  unsigned char* wbase = smem + fixed + (size_t)warp * warp_bytes;

  // ---- init
  if (!dense) for (int i = threadIdx.x; i < S; i += blockDim.x) ktab_id[i] = 0;
  if (threadIdx.x < 4) misc[threadIdx.x] = 0;
  for (int i = threadIdx.x; i <= cap; i += blockDim.x) {
    NUL[i] = 0;
    if (ALL) {
      const float inf = __int_as_float(0x7f800000);
      META[i] = make_ulonglong2(0ull, ((u64)__float_as_uint(-inf) << 32) | (u64)__float_as_uint(inf));   // .y = {lo: min bound, hi: max bound}
      EXACT[i] = make_ulonglong2(0ull, 0ull);
    }
  }
  for (size_t i = lane; i < warp_bytes / 8; i += 32) reinterpret_cast<u64*>(wbase)[i] = 0;
  __syncthreads();

  GbShared<VT, FLAGS> sh;
  sh.ktab_key = ktab_key; sh.ktab_id = ktab_id; sh.misc = misc; sh.META = META; sh.EXACT = EXACT;
  // per-warp arrays over E + 1 records: ACC (8 or 16 B), [ISUM 8 B], CNT u32
  sh.a_meta = sm_addr(META); sh.a_nul = sm_addr(NUL);
  sh.a_acc = sm_addr(wbase);
  sh.a_isum = sm_addr(wbase + (size_t)L::ACC_BYTES * (E + 1));
  sh.a_cnt = sm_addr(wbase + (size_t)(L::REC_BYTES - 4) * (E + 1));
  sh.S = S; sh.cap = cap; sh.NG = NG; sh.E = E; sh.rep = lane & (NG - 1); sh.lane = lane;
  sh.dense_base = (u64)p.sh_dense_base;

  const long long n = p.n;
  const long long full_units = n / GB_UNIT_ROWS;
  const long long gwarp = (long long)blockIdx.x * nwarps + warp, total_warps = (long long)gridDim.x * nwarps;

  // Full units, double-buffered through registers: the loads of unit u + total_warps are issued before unit u is
  // aggregated, and nothing of the next unit is touched until then.
  // Units are double-buffered through registers: the loads of unit u + total_warps are issued before unit u is
  // aggregated, and nothing of the next unit is touched until then.  The common case (no filter, no NULL keys)
  // runs a leaner body over the full units; everything else, and the partial last unit, takes the general one.
  const long long total_units = (n + GB_UNIT_ROWS - 1) / GB_UNIT_ROWS;
  const bool plain = !p.fbits && !p.compat_nulls && (KM == 1 || !p.ks.c[0].nulls);
  GbUnit<VT> cur, nxt;
  long long u = gwarp;
  if (plain) {
    if (u < full_units) gb_load_unit<KM, true, VT>(p, u * GB_UNIT_ROWS, lane, cur);
#pragma unroll 1
    for (; u < full_units; u += total_warps) {
      if (u + total_warps < full_units) gb_load_unit<KM, true, VT>(p, (u + total_warps) * GB_UNIT_ROWS, lane, nxt);
      gb_unit_body<NW, KM, VT, FLAGS, true>(p, sh, cur, u * GB_UNIT_ROWS);
      cur = nxt;
    }
  }
  if (u < total_units) {
    if (u < full_units) gb_load_unit<KM, true, VT>(p, u * GB_UNIT_ROWS, lane, cur); else gb_load_unit<KM, false, VT>(p, u * GB_UNIT_ROWS, lane, cur);
  }
#pragma unroll 1
  for (; u < total_units; u += total_warps) {
    const long long un = u + total_warps;
    if (un < total_units) {
      if (un < full_units) gb_load_unit<KM, true, VT>(p, un * GB_UNIT_ROWS, lane, nxt); else gb_load_unit<KM, false, VT>(p, un * GB_UNIT_ROWS, lane, nxt);
    }
    gb_unit_body<NW, KM, VT, FLAGS, false>(p, sh, cur, u * GB_UNIT_ROWS);
    cur = nxt;
  }
  __syncthreads();

  // ---- flush: reduce the warps' private records per group, then one batch update per (CTA, group)
  const int nid = dense ? cap : S;
  for (int s0 = 0; s0 <= nid; s0 += blockDim.x) {     // warp-uniform trip count (g_find_or_insert is warp-synchronous)
    const int s = s0 + threadIdx.x;
    int id = -1;
    if (s == nid) { if (misc[1]) id = cap; }
    else if (s < nid) {
      if (dense) id = s;
      else { uint32_t idw = ktab_id[s]; if (idw != 0 && idw != SH_BUSY) id = (int)idw - 1; }
    }
    bool have = id >= 0;
    if (!have) id = 0;
    u64 rows = 0, isum = 0;
    const u64 nulls = NUL[id];
    double S1 = 0.0, S2 = 0.0;
    for (int wq = 0; wq < nwarps; wq++) {
      const unsigned char* qb = smem + fixed + (size_t)wq * warp_bytes;
      const uint32_t* QC = reinterpret_cast<const uint32_t*>(qb + (size_t)(L::REC_BYTES - 4) * (E + 1));
      for (int r = 0; r < NG; r++) {
        const int e = id * NG + r;
        rows += QC[e];
        if (!ALL) {
          const u64 a = reinterpret_cast<const u64*>(qb)[e];
          if (IS_INT) isum += a; else S1 += __longlong_as_double((long long)a);
        } else {
          const ulonglong2 a = reinterpret_cast<const ulonglong2*>(qb)[e];
          S1 += __longlong_as_double((long long)a.x);
          S2 += __longlong_as_double((long long)a.y);
          if (IS_INT) isum += reinterpret_cast<const u64*>(qb + (size_t)16 * (E + 1))[e];
        }
      }
    }
    if (have && rows == 0) have = false;       // dense mode: ids nobody hit
    u64 w[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = 0;
    if (have && id != cap) {
      if (dense) w[0] = sh.dense_base + (u64)id;
      else {
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = ktab_key[i * S + s];
      }
    }
    long long gs = g_find_or_insert<NW>(p.gt, w, have && id != cap);
    if (have && id == cap) { gs = p.gt.slots; if (!(ld_cg_u64(&p.gt.hdr[gs].rowsw) & GB_FULL)) atomicOr(&p.gt.hdr[gs].rowsw, GB_FULL); }
    if (!have || gs < 0) continue;
    u64 mnc = 0, mxo = 0, px = 0;
    if (ALL) { px = META[id].x; mnc = EXACT[id].x; mxo = EXACT[id].y; }
    if (p.count_rows) atomicAdd(&p.gt.hdr[gs].rowsw, rows);
    g_update_batch<FLAGS, IS_INT>(p.gt, gs, 0ull, rows - nulls, __longlong_as_double((long long)px), px != 0, S1, S2, isum, mnc, mxo);
  }
}

// ---------------------------------------------------------------- the global-table kernel
template <int NW, int KM, typename VT, int FLAGS>
__global__ void __launch_bounds__(256) gb_global_kernel(const GbParams p) {
  using T = ValTraits<VT>;
  constexpr int R = 4;
  const long long n = p.n;
  const VT* __restrict__ vals = reinterpret_cast<const VT*>(p.val);
  const long long* __restrict__ keys64 = reinterpret_cast<const long long*>(p.ks.c[0].data);
  const uint8_t* knull0 = p.ks.c[0].nulls;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // warp-uniform loop bounds: g_find_or_insert is warp-synchronous
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 - lane < n; i0 += stride * R) {
    u64 k0[R];
    VT v[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
      long long row = i0 + j * stride;
      bool inb = row < n;
      if (KM == 0) k0[j] = inb ? (u64)__ldcs(keys64 + row) : 0ull;
      v[j] = (vals && inb) ? __ldcs(vals + row) : VT(0);
    }
#pragma unroll
    for (int j = 0; j < R; j++) {
      long long row = i0 + j * stride;
      bool active = row < n;
      if (active && p.fbits) {
        if (!pdrs_bit(p.fbits, row)) active = false;
        else if (p.fnull && pdrs_bit(p.fnull, row)) active = false;
      }
      bool valid = active && vals && !(p.vnull && pdrs_bit(p.vnull, row));
      VT vv = v[j];
      if (p.compat_nulls && active && vals && !valid) { valid = true; vv = VT(0); }
      u64 w[NW];
#pragma unroll
      for (int i = 0; i < NW; i++) w[i] = 0;
      bool knull = false;
      if (active) {
        if (KM == 0) { w[0] = k0[j]; if (knull0) knull = pdrs_bit(knull0, row); }
        else knull = load_key_generic<NW>(p.ks, row, w);
      }
      long long gs = g_find_or_insert<NW>(p.gt, w, active && !knull);
      if (active && knull) { gs = p.gt.slots; if (!(ld_cg_u64(&p.gt.hdr[gs].rowsw) & GB_FULL)) atomicOr(&p.gt.hdr[gs].rowsw, GB_FULL); }
      if (active && gs >= 0) g_update_row<VT, FLAGS>(p.gt, gs, p.count_rows != 0, valid, vv);
    }
  }
  (void)sizeof(T);
}

// ---------------------------------------------------------------- sampling (cardinality estimate)
// Inserts the keys of `nblocks` evenly spread runs of 256 rows into a scratch table; CNT_NGROUPS then
// holds the number of distinct sampled keys.
template <int NW, int KM>
__global__ void gb_sample_kernel(const GbParams p, long long nblocks, long long block_stride_rows) {
  u64 kmax = 0, kminc = 0;
  for (long long b = blockIdx.x; b < nblocks; b += gridDim.x) {
    long long row = b * block_stride_rows + threadIdx.x;
    const bool inb = row < p.n;
    u64 w[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = 0;
    bool knull = false;
    if (inb) {
      if (KM == 0) { w[0] = (u64)__ldg(reinterpret_cast<const long long*>(p.ks.c[0].data) + row); knull = p.ks.c[0].nulls && pdrs_bit(p.ks.c[0].nulls, row); }
      else knull = load_key_generic<NW>(p.ks, row, w);
    }
    if (KM == 0 && inb && !knull) { const u64 o = w[0] ^ GB_SIGN; kmax = max(kmax, o); kminc = max(kminc, ~o); }
    const long long gs = g_find_or_insert<NW>(p.gt, w, inb && !knull);
    if (inb && !knull && gs >= 0) atomicAdd(&p.gt.hdr[gs].rowsw, 1ull);      // occurrences in the sample: hot keys of a skewed distribution
  }
  if (KM == 0) {   // range of the sampled keys (dense-key fast path)
    for (int d = 16; d; d >>= 1) { kmax = max(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, d)); kminc = max(kminc, __shfl_xor_sync(0xFFFFFFFFu, kminc, d)); }
    if ((threadIdx.x & 31) == 0 && kminc) { atomicMax(&p.gt.counters[CNT_KMAX], kmax); atomicMax(&p.gt.counters[CNT_KMINC], kminc); }
  }
}

// ---------------------------------------------------------------- launcher interface (gb_inst.cu)
struct GbCfg {
  int nw, km, vt /*0 f64, 1 i64*/, flags;
  int warps, ctas;
};
int32_t pdrs_build_keyspec(pdrs_ctx* c, const ColView* kv, int nkeys, KeySpec* ks);

// one translation unit per key variant (gb_inst_*.cu) so that the build parallelises
#define GB_DECLARE_VARIANT(tag)                                                                                     \
  cudaError_t gb_launch_shared_##tag(const GbCfg& c, const GbParams& p, size_t smem, cudaStream_t s);               \
  cudaError_t gb_launch_global_##tag(const GbCfg& c, const GbParams& p, cudaStream_t s);                            \
  cudaError_t gb_launch_sample_##tag(const GbParams& p, long long nblocks, long long stride_rows, int ctas, cudaStream_t s);
struct GbPart;   // gb_radix.cu
GB_DECLARE_VARIANT(k1)   // NW = 1, one 64-bit key column, direct loads
GB_DECLARE_VARIANT(g1)   // NW = 1, generic packing
GB_DECLARE_VARIANT(g2)   // NW = 2
GB_DECLARE_VARIANT(g3)   // NW = 3
