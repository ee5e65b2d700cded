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
#define GB_R 16                         // rows per lane per unit
#define GB_UNIT_ROWS (32 * GB_R)
#define GB_MAX_WARPS 8

typedef unsigned long long u64;

static constexpr u64 GB_BUSY = 1ull << 63;
static constexpr u64 GB_FULL = 1ull << 62;
static constexpr u64 GB_CNT_MASK = GB_FULL - 1;
static constexpr u64 GB_PIV_X = 0x7FF8C0DEC0DE0001ull;   // pivot bits are stored XOR this (0 = unset)
static constexpr u64 GB_SIGN = 1ull << 63;

enum { CNT_NGROUPS = 0, CNT_OVERFLOW = 1, CNT_SPILLED = 2, CNT_OUT = 3, CNT_SPIN_FAIL = 4, CNT_KMINC = 5 /* ~(min key ^ SIGN) */, CNT_KMAX = 6 /* max key ^ SIGN */, CNT_N = 8 };

struct KeyColDev {
  const void* data;
  const uint8_t* nulls;
  long long null_alias;
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
  u64 mask;                      // slots - 1
  long long slots;
  u64* counters;                 // CNT_*
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
};

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
        case PDRS_I64: v = (u64)__ldg((const long long*)c.data + row); break;
        case PDRS_F64: {
          double d = __ldg((const double*)c.data + row);
          v = (d != d) ? 0x7FF8000000000000ull : (u64)__double_as_longlong(d);   // all NaNs print "NaN"
          break;
        }
        case PDRS_I32: v = (u64)(uint32_t)__ldg((const int*)c.data + row); break;
        case PDRS_DICT_U32: {
          uint32_t id = __ldg((const uint32_t*)c.data + row);
          if ((long long)id == c.null_alias) isnull = true; else v = id;
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
  u64 slot = (key_hash<NW>(w) >> 20) & t.mask, probe = 0;
  long long res = -1;
  bool pending = active;
  int rounds = 0;
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
//              misc u32[4]
// per warp     ACC[E]  f64 {S1}, i64 {isum}                       (SUM)        E = (cap + 1) * ng
//                      {S1, S2} (+ ISUM[E] for i64 values)        (ALL)
//              CNT[E]  u32 rows: bumped with the native atomic; the returned ticket ranks the lanes of the
//                      warp that hit the same record in this batch (re-read after __syncwarp gives the count)
//              NUL[E]  u32 NULL values (n = rows - nulls)
// S1 / S2 are plain LDS/STS read-modify-writes, one rank per round, so the per-row path has no 64-bit atomics.
template <typename VT, int FLAGS> struct ShPlanes {
  static constexpr bool IS_INT = ValTraits<VT>::is_int;
  static constexpr int ACC_BYTES = FLAGS == GB_SUM ? 8 : 16;
  static constexpr int REC_BYTES = ACC_BYTES + ((FLAGS == GB_ALL && IS_INT) ? 8 : 0) + 8;   // per warp per record
  static constexpr int CTA_BYTES = FLAGS == GB_ALL ? 32 : 0;                                 // META + EXACT per group
};
__host__ __device__ inline size_t gb_sh_fixed_bytes(int nw, int slots, int cap, int cta_bytes, int dense) {
  size_t b = dense ? 0 : ((size_t)8 * nw * slots + (size_t)4 * slots);
  b = (b + 15) / 16 * 16;
  b += (size_t)cta_bytes * (cap + 1);
  return b + 16;
}
__host__ __device__ inline size_t gb_sh_warp_bytes(int rec_bytes, int cap, int ng) {
  const size_t E = (size_t)(cap + 1) * ng;
  return ((size_t)rec_bytes * E + 15) / 16 * 16;
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

// One unit of a warp: 512 rows, lane owns rows base + 64*j + 2*lane + {0,1}, j = 0..7.
// fl[j / 4] holds, 6 bits per j: {active(2), NULL value(2), NULL key(2)} of this lane's two rows.
template <typename VT> struct GbUnit {
  u64 k[GB_R];     // KM != 1 only: the 64-bit key column
  u64 v[GB_R];     // value bits
  uint32_t fl[2];
};

template <int KM, typename VT>
__device__ __forceinline__ void gb_load_unit(const GbParams& p, long long base, int lane, GbUnit<VT>& u) {
  const long long n = p.n;
  const char* keys = reinterpret_cast<const char*>(p.ks.c[0].data);
  const char* vals = reinterpret_cast<const char*>(p.val);
#pragma unroll
  for (int j = 0; j < GB_R / 2; j++) {
    const long long r0 = base + 64 * j + 2 * lane;
    ulonglong2 kk = make_ulonglong2(0, 0), vv = make_ulonglong2(0, 0);
    if (r0 + 1 < n) {
      if (KM != 1) kk = ld_stream_v2(keys + 8 * r0);
      if (vals) vv = ld_stream_v2(vals + 8 * r0);
    } else if (r0 < n) {
      if (KM != 1) kk.x = __ldg(reinterpret_cast<const u64*>(keys) + r0);
      if (vals) vv.x = __ldg(reinterpret_cast<const u64*>(vals) + r0);
    }
    u.k[2 * j] = kk.x; u.k[2 * j + 1] = kk.y;
    u.v[2 * j] = vv.x; u.v[2 * j + 1] = vv.y;
  }
  u.fl[0] = u.fl[1] = 0;
#pragma unroll
  for (int j = 0; j < GB_R / 2; j++) {
    const long long c0 = base + 64 * j;
    u64 m = 0, vn = 0, kn = 0;
    if (c0 < n) {
      const long long rem = n - c0, chunk = c0 >> 6;
      m = rem >= 64 ? ~0ull : ((1ull << rem) - 1ull);
      if (p.fbits) {   // filter: Some(true) rows only (data_ops.rs:49-55)
        m &= load_bits64(p.fbits, chunk);
        if (p.fnull) m &= ~load_bits64(p.fnull, chunk);
      }
      if (p.vnull) vn = load_bits64(p.vnull, chunk);
      if (KM != 1 && p.ks.c[0].nulls) kn = load_bits64(p.ks.c[0].nulls, chunk);
    }
    if (!vals) vn = ~0ull;
    const uint32_t a2 = (uint32_t)(m >> (2 * lane)) & 3u, v2 = (uint32_t)(vn >> (2 * lane)) & 3u, k2 = (uint32_t)(kn >> (2 * lane)) & 3u;
    u.fl[j >> 2] |= (a2 | (v2 << 2) | (k2 << 4)) << (6 * (j & 3));
  }
  if (p.compat_nulls && vals) {   // filter + compat_filter_nulls: NULL values count as 0 (data_ops.rs:64-71)
#pragma unroll
    for (int jj = 0; jj < GB_R; jj++) {
      const uint32_t bit = 1u << (6 * ((jj >> 1) & 3) + 2 + (jj & 1));
      if (u.fl[jj >> 3] & bit) { u.fl[jj >> 3] &= ~bit; u.v[jj] = 0; }
    }
  }
}

// ---------------------------------------------------------------- the shared-memory kernel
// Rare paths live in __noinline__ functions so that the unrolled hot loop stays inside the instruction cache.
template <int NW, typename VT, int FLAGS>
__device__ __noinline__ void gb_spill_rows(const GTable gt, u64 w0, u64 w1, u64 w2, bool spill, bool count_row, bool valid, VT v) {
  u64 w[NW];
  w[0] = w0;
  if (NW > 1) w[1] = w1;
  if (NW > 2) w[2] = w2;
  long long gs = g_find_or_insert<NW>(gt, w, spill);
  if (spill && gs >= 0) g_update_row<VT, FLAGS>(gt, gs, count_row, valid, v);
  if (spill) atomicAdd(&gt.counters[CNT_SPILLED], 1ull);
}

// A value beat the f32 bound of its group: update the exact minimum / maximum (64-bit shared atomicMax is a
// CAS loop, but this runs O(log rows) times per group) and tighten the bound.
template <typename VT>
__device__ __noinline__ void gb_minmax_slow(ulonglong2* META, ulonglong2* EXACT, int id, VT v, bool lo, bool hi) {
  using T = ValTraits<VT>;
  if (!T::orderable(v)) return;
  if (lo) {
    const u64 o = ~T::ord(v);
    const u64 was = atomicMax(&EXACT[id].x, o);
    const u64 best = ~(was > o ? was : o);     // ord() of the current exact minimum
    float b;
    if (T::is_int) b = __ll2float_rn((long long)(best ^ GB_SIGN)); else b = __double2float_rn(pdrs_unord_f64(best));
    reinterpret_cast<volatile uint32_t*>(&META[id].y)[0] = __float_as_uint(b);   // bound = rn(some earlier exact min) >= rn(exact min)
  }
  if (hi) {
    const u64 o = T::ord(v);
    const u64 was = atomicMax(&EXACT[id].y, o);
    const u64 best = was > o ? was : o;
    float b;
    if (T::is_int) b = __ll2float_rn((long long)(best ^ GB_SIGN)); else b = __double2float_rn(pdrs_unord_f64(best));
    reinterpret_cast<volatile uint32_t*>(&META[id].y)[1] = __float_as_uint(b);   // bound = rn(some earlier exact max) <= rn(exact max)
  }
}

static __device__ __noinline__ u64 gb_pivot_set(ulonglong2* META, int id, double x) {
  const u64 mine = (u64)__double_as_longlong(x) ^ GB_PIV_X;
  const u64 was = atomicCAS(&META[id].x, 0ull, mine);
  return was ? was : mine;
}

#define GB_Q 4   // batches (of 32 rows) that share one ticket phase

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
  const int E = (cap + 1) * NG;
  u64* ktab_key = reinterpret_cast<u64*>(smem);
  uint32_t* ktab_id = reinterpret_cast<uint32_t*>(smem + (size_t)8 * NW * S);
  const size_t keys_bytes = dense ? 0 : (((size_t)8 * NW * S + (size_t)4 * S + 15) / 16 * 16);
  ulonglong2* META = reinterpret_cast<ulonglong2*>(smem + keys_bytes);                 // ALL only
  ulonglong2* EXACT = META + (cap + 1);                                               // ALL only
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + keys_bytes + (size_t)L::CTA_BYTES * (cap + 1));   // [0] groups in this CTA, [1] NULL group seen
  const size_t fixed = gb_sh_fixed_bytes(NW, S, cap, L::CTA_BYTES, dense);
  const size_t warp_bytes = gb_sh_warp_bytes(L::REC_BYTES, cap, NG);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  unsigned char* wbase = smem + fixed + (size_t)warp * warp_bytes;
  // per-warp arrays: ACC (8 or 16 B), [ISUM 8 B], CNT u32, NUL u32
  u64* ACC1 = reinterpret_cast<u64*>(wbase);                 // SUM: one 8-byte word per record
  ulonglong2* ACC2 = reinterpret_cast<ulonglong2*>(wbase);   // ALL: {S1, S2}
  u64* ISUM = reinterpret_cast<u64*>(wbase + (size_t)L::ACC_BYTES * E);
  uint32_t* CNT = reinterpret_cast<uint32_t*>(wbase + (size_t)(L::REC_BYTES - 8) * E);
  uint32_t* NUL = CNT + E;

  // ---- init
  if (!dense) for (int i = threadIdx.x; i < S; i += blockDim.x) ktab_id[i] = 0;
  if (threadIdx.x < 4) misc[threadIdx.x] = 0;
  if (ALL) {
    const float inf = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i <= cap; i += blockDim.x) {
      META[i] = make_ulonglong2(0ull, ((u64)__float_as_uint(-inf) << 32) | (u64)__float_as_uint(inf));   // .y = {lo: min bound, hi: max bound}
      EXACT[i] = make_ulonglong2(0ull, 0ull);
    }
  }
  for (size_t i = lane; i < warp_bytes / 8; i += 32) reinterpret_cast<u64*>(wbase)[i] = 0;
  __syncthreads();

  const long long n = p.n;
  const long long total_units = (n + GB_UNIT_ROWS - 1) / GB_UNIT_ROWS;
  const long long gwarp = (long long)blockIdx.x * nwarps + warp, total_warps = (long long)gridDim.x * nwarps;
  const int rep = lane & (NG - 1);
  const u64 dense_base = (u64)p.sh_dense_base;
  const bool count_rows = p.count_rows != 0;

  // The aggregation (shared-memory bound, >1 cycle per row per SM) dominates a unit by far, so the loads of a
  // unit are not double-buffered through registers; the next unit of this warp is pulled into L2 instead.
  GbUnit<VT> cur;
  for (long long u = gwarp; u < total_units; u += total_warps) {
    const long long base = u * GB_UNIT_ROWS;
    gb_load_unit<KM, VT>(p, base, lane, cur);
    if (u + total_warps < total_units) {
      const long long nb = (u + total_warps) * GB_UNIT_ROWS + 16 * lane;   // 32 lanes x 128 B = one unit of an 8-byte column
      if (nb < n) {
        if (KM != 1) asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(p.ks.c[0].data) + 8 * nb));
        if (p.val) asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(p.val) + 8 * nb));
      }
    }

#pragma unroll
    for (int g0 = 0; g0 < GB_R; g0 += GB_Q) {
      int eidx[GB_Q];          // record index of the row, -1: nothing to add to the sums
      int rr[GB_Q];
      VT vv[GB_Q];
      double dd[GB_Q];
      uint32_t old[GB_Q];
      u64 wsp[GB_Q][NW];
      uint32_t spillmask = 0;
      // -- phase 1: key -> record, bump the row counters (tickets), pivot / min / max
#pragma unroll
      for (int q = 0; q < GB_Q; q++) {
        const int jj = g0 + q;
        const uint32_t fl = cur.fl[jj >> 3];
        constexpr int dummy = 0; (void)dummy;
        const int sh = 6 * ((jj >> 1) & 3) + (jj & 1);
        const bool active = (fl >> sh) & 1u;
        const bool valid = active && !((fl >> (sh + 2)) & 1u);
        bool knull = (fl >> (sh + 4)) & 1u;
        const VT v = T::from_bits(cur.v[jj]);
        u64 w[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = 0;
        if (KM != 1) w[0] = cur.k[jj];
        else if (active) knull = load_key_generic<NW>(p.ks, base + 64 * (jj >> 1) + 2 * lane + (jj & 1), w);
        int id = -1;
        if (dense) {                       // small dense integer keys: direct-mapped
          const u64 off = w[0] - dense_base;
          if (active && off < (u64)cap) id = (int)off;
        } else {                           // CTA-shared key table
          id = sh_lookup<NW>(ktab_key, ktab_id, misc, S, p.sh_log_slots, cap, w, active && !knull);
        }
        if (active && knull) { id = cap; if (!*reinterpret_cast<volatile uint32_t*>(&misc[1])) *reinterpret_cast<volatile uint32_t*>(&misc[1]) = 1u; }
        if (active && id < 0) spillmask |= 1u << q;
#pragma unroll
        for (int i = 0; i < NW; i++) wsp[q][i] = w[i];
        const int e = id * NG + rep;
        old[q] = 0;
        if (id >= 0) {
          old[q] = atomicAdd(&CNT[e], 1u);
          if (!valid) atomicAdd(&NUL[e], 1u);
        }
        const bool upd = id >= 0 && valid;
        eidx[q] = upd ? e : -1;
        vv[q] = v;
        double x = T::to_f64(v);
        dd[q] = x;
        if (ALL && upd) {                  // pivot and min / max bounds: CTA-shared, read-mostly, one LDS.128
          const ulonglong2 meta = lds_volatile_v2(&META[id]);
          u64 px = meta.x;
          if (px == 0 && is_finite_f64(x)) px = gb_pivot_set(META, id, x);   // first finite value of the group in this CTA
          dd[q] = x - (px ? __longlong_as_double((long long)(px ^ GB_PIV_X)) : 0.0);
          // f32 bounds hold round-to-nearest(exact min / max); rounding is monotonic, so a value below the exact
          // minimum always satisfies xf <= bound (values equal to the bound after rounding take the slow path too)
          const float xf = IS_INT ? __ll2float_rn((long long)T::to_bits(v)) : __double2float_rn(x);
          const bool lo = xf <= __uint_as_float((uint32_t)meta.y), hi = xf >= __uint_as_float((uint32_t)(meta.y >> 32));
          if (lo || hi) gb_minmax_slow<VT>(META, EXACT, id, v, lo, hi);
        }
      }
      if (__any_sync(0xFFFFFFFFu, spillmask != 0)) {   // rare: keys that do not fit this CTA's table go to the global table
#pragma unroll
        for (int q = 0; q < GB_Q; q++) {
          const int jj = g0 + q;
          const bool valid = !((cur.fl[jj >> 3] >> (6 * ((jj >> 1) & 3) + (jj & 1) + 2)) & 1u);
          gb_spill_rows<NW, VT, FLAGS>(p.gt, wsp[q][0], NW > 1 ? wsp[q][NW > 1 ? 1 : 0] : 0ull, NW > 2 ? wsp[q][NW > 2 ? 2 : 0] : 0ull,
                                       (spillmask >> q) & 1u, count_rows, valid, vv[q]);
        }
      }
      // -- phase 2: ranks from the tickets (unique per record across the whole group of GB_Q batches)
      __syncwarp();
      int mr = 0;
#pragma unroll
      for (int q = 0; q < GB_Q; q++) {
        rr[q] = -1;
        if (eidx[q] >= 0) { rr[q] = (int)(*reinterpret_cast<volatile uint32_t*>(&CNT[eidx[q]]) - old[q] - 1u); mr = max(mr, rr[q]); }
      }
      const int maxr = __reduce_max_sync(0xFFFFFFFFu, mr);
      // -- phase 3: plain read-modify-write of this warp's sums, one rank per round
      for (int r = 0; r <= maxr; r++) {
        if (!ALL) {
          u64 a[GB_Q];
#pragma unroll
          for (int q = 0; q < GB_Q; q++) if (rr[q] == r) a[q] = ACC1[eidx[q]];
#pragma unroll
          for (int q = 0; q < GB_Q; q++) if (rr[q] == r) {
            if (IS_INT) a[q] += (u64)T::to_bits(vv[q]);
            else a[q] = (u64)__double_as_longlong(__longlong_as_double((long long)a[q]) + dd[q]);
            ACC1[eidx[q]] = a[q];
          }
        } else {
          ulonglong2 a[GB_Q];
#pragma unroll
          for (int q = 0; q < GB_Q; q++) if (rr[q] == r) a[q] = ACC2[eidx[q]];
#pragma unroll
          for (int q = 0; q < GB_Q; q++) if (rr[q] == r) {
            a[q].x = (u64)__double_as_longlong(__longlong_as_double((long long)a[q].x) + dd[q]);
            a[q].y = (u64)__double_as_longlong(__longlong_as_double((long long)a[q].y) + dd[q] * dd[q]);
            ACC2[eidx[q]] = a[q];
            if (IS_INT) ISUM[eidx[q]] += (u64)T::to_bits(vv[q]);
          }
        }
        __syncwarp();
      }
    }
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
    u64 rows = 0, nulls = 0, isum = 0;
    double S1 = 0.0, S2 = 0.0;
    for (int wq = 0; wq < nwarps; wq++) {
      const unsigned char* qb = smem + fixed + (size_t)wq * warp_bytes;
      const uint32_t* QC = reinterpret_cast<const uint32_t*>(qb + (size_t)(L::REC_BYTES - 8) * E);
      for (int r = 0; r < NG; r++) {
        const int e = id * NG + r;
        rows += QC[e];
        nulls += QC[E + e];
        if (!ALL) {
          const u64 a = reinterpret_cast<const u64*>(qb)[e];
          if (IS_INT) isum += a; else S1 += __longlong_as_double((long long)a);
        } else {
          const ulonglong2 a = reinterpret_cast<const ulonglong2*>(qb)[e];
          S1 += __longlong_as_double((long long)a.x);
          S2 += __longlong_as_double((long long)a.y);
          if (IS_INT) isum += reinterpret_cast<const u64*>(qb + (size_t)16 * E)[e];
        }
      }
    }
    if (have && rows == 0) have = false;       // dense mode: ids nobody hit
    u64 w[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = 0;
    if (have && id != cap) {
      if (dense) w[0] = dense_base + (u64)id;
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
    g_update_batch<FLAGS, IS_INT>(p.gt, gs, 0ull, rows - nulls, px ? __longlong_as_double((long long)(px ^ GB_PIV_X)) : 0.0, px != 0, S1, S2, isum, mnc, mxo);
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
    g_find_or_insert<NW>(p.gt, w, inb && !knull);
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
GB_DECLARE_VARIANT(k1)   // NW = 1, one 64-bit key column, direct loads
GB_DECLARE_VARIANT(g1)   // NW = 1, generic packing
GB_DECLARE_VARIANT(g2)   // NW = 2
GB_DECLARE_VARIANT(g3)   // NW = 3
