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

enum { CNT_NGROUPS = 0, CNT_OVERFLOW = 1, CNT_SPILLED = 2, CNT_OUT = 3, CNT_SPIN_FAIL = 4, CNT_N = 8 };

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
};

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ u64 ld_cg_u64(const u64* p) { return __ldcg(p); }
__device__ __forceinline__ ulonglong2 ld_cg_hdr(const GHdr* p) { return __ldcg(reinterpret_cast<const ulonglong2*>(p)); }

// streaming 128-bit load that the compiler may not sink below the aggregation of the previous unit
__device__ __forceinline__ ulonglong2 ld_stream_v2(const void* p) {
  ulonglong2 r;
  asm volatile("ld.global.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
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
__device__ __forceinline__ bool load_key_generic(const KeySpec& ks, long long row, u64 (&w)[NW]) {
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
// Returns the slot of the key (claiming a free one if INSERT), or -1 on overflow / not found.
template <int NW, bool INSERT = true>
__device__ __forceinline__ long long g_find_or_insert(const GTable& t, const u64 (&w)[NW]) {
  u64 slot = (key_hash<NW>(w) >> 20) & t.mask;
  int spins = 0;
  for (u64 probe = 0; probe <= t.mask;) {
    ulonglong2 h = ld_cg_hdr(&t.hdr[slot]);
    if (h.y == 0) {
      if (!INSERT) return -1;
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
    if (h.y & GB_BUSY) {             // another thread is publishing this slot: look again
      if (++spins > (1 << 22)) { atomicAdd(&t.counters[CNT_SPIN_FAIL], 1ull); return -1; }
      continue;
    }
    bool match = h.x == w[0];
    if (NW > 1) match = match && ld_cg_u64(&t.kw1[slot]) == w[1];
    if (NW > 2) match = match && ld_cg_u64(&t.kw2[slot]) == w[2];
    if (match) return (long long)slot;
    slot = (slot + 1) & t.mask;
    probe++;
  }
  if (INSERT) atomicAdd(&t.counters[CNT_OVERFLOW], 1ull);
  return -1;
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
// CTA-shared: key table u64[NW][S], id table u32[S] (0 = empty, SH_BUSY, else id + 1), misc u32[4].
// Per warp: NPL planes of ulonglong2[E], E = (cap + 1) * ng  (+1: the NULL-key group).
//   f64 SUM: P0 = {S1, cnt}            cnt = rows | n << 32
//   i64 SUM: P0 = {isum, cnt}
//   f64 ALL: P0 = {S1, S2}  P1 = {cnt, pivotx}  P2 = {min, max}
//   i64 ALL: P0 = {S1, S2}  P1 = {cnt, pivotx}  P2 = {min, max}  P3 = {isum, -}
template <typename VT, int FLAGS> struct ShPlanes {
  static constexpr int NPL = FLAGS == GB_SUM ? 1 : (ValTraits<VT>::is_int ? 4 : 3);
};
__host__ __device__ inline size_t gb_sh_fixed_bytes(int nw, int slots) { return ((size_t)8 * nw * slots + (size_t)4 * slots + 16 + 15) / 16 * 16; }
__host__ __device__ inline size_t gb_sh_warp_bytes(int npl, int cap, int ng) { return (size_t)npl * 16 * (size_t)(cap + 1) * ng; }

static constexpr uint32_t SH_BUSY = 0xFFFFFFFFu;

// key -> dense id through the CTA-shared key table; -1 = table full (spill)
template <int NW>
__device__ __forceinline__ int sh_lookup(u64* ktab_key, uint32_t* ktab_id, uint32_t* misc, int S, int log_slots, int cap, const u64 (&w)[NW]) {
  uint32_t slot = (uint32_t)(key_hash<NW>(w) >> (64 - log_slots));
  int spins = 0;
  for (int probe = 0; probe < S;) {
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
      idw = SH_BUSY;   // lost the race: the winner is publishing, look again
    }
    if (idw == SH_BUSY) { if (++spins > (1 << 20)) return -1; continue; }
    bool match = true;
#pragma unroll
    for (int i = 0; i < NW; i++) match = match && (kk[i] == w[i]);
    if (match) return (int)idw - 1;
    slot = (slot + 1) & (S - 1);
    probe++;
  }
  return -1;
}

// One unit of a warp: 512 rows, lane owns rows base + 64*j + 2*lane + {0,1}, j = 0..7.
template <typename VT> struct GbUnit {
  u64 k[GB_R];     // KM == 0 only: the 64-bit key column
  u64 v[GB_R];     // value bits
  u64 act[GB_R / 2], vn[GB_R / 2], kn[GB_R / 2];   // per 64-row chunk: active rows, NULL values, NULL keys
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
      if (KM == 0) kk = ld_stream_v2(keys + 8 * r0);
      if (vals) vv = ld_stream_v2(vals + 8 * r0);
    } else if (r0 < n) {
      if (KM == 0) kk.x = __ldg(reinterpret_cast<const u64*>(keys) + r0);
      if (vals) vv.x = __ldg(reinterpret_cast<const u64*>(vals) + r0);
    }
    u.k[2 * j] = kk.x; u.k[2 * j + 1] = kk.y;
    u.v[2 * j] = vv.x; u.v[2 * j + 1] = vv.y;
  }
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
      if (KM == 0 && p.ks.c[0].nulls) kn = load_bits64(p.ks.c[0].nulls, chunk);
    }
    if (!vals) vn = ~0ull;
    u.act[j] = m; u.vn[j] = vn; u.kn[j] = kn;
  }
}

// ---------------------------------------------------------------- the shared-memory kernel
template <int NW, int KM /*0: one 64-bit key column, direct loads; 1: generic packing*/, typename VT, int FLAGS>
__global__ void __launch_bounds__(GB_MAX_WARPS * 32, 1) gb_shared_kernel(const GbParams p) {
  using T = ValTraits<VT>;
  constexpr int NPL = ShPlanes<VT, FLAGS>::NPL;
  constexpr bool IS_INT = T::is_int;
  extern __shared__ __align__(16) unsigned char smem[];
  const int S = p.sh_slots, cap = p.sh_cap, NG = p.sh_ng;
  const int E = (cap + 1) * NG;
  u64* ktab_key = reinterpret_cast<u64*>(smem);
  uint32_t* ktab_id = reinterpret_cast<uint32_t*>(smem + (size_t)8 * NW * S);
  uint32_t* misc = ktab_id + S;   // [0] = number of groups in this CTA, [1] = NULL group seen
  const size_t fixed = gb_sh_fixed_bytes(NW, S);
  const size_t warp_bytes = gb_sh_warp_bytes(NPL, cap, NG);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  ulonglong2* P0 = reinterpret_cast<ulonglong2*>(smem + fixed + (size_t)warp * warp_bytes);
  ulonglong2* P1 = P0 + E;
  ulonglong2* P2 = P1 + E;
  ulonglong2* P3 = P2 + E;

  // ---- init
  for (int i = threadIdx.x; i < S; i += blockDim.x) ktab_id[i] = 0;
  if (threadIdx.x < 4) misc[threadIdx.x] = 0;
  for (int e = lane; e < E; e += 32) {
    P0[e] = make_ulonglong2(0, 0);
    if (FLAGS == GB_ALL) {
      P1[e] = make_ulonglong2(0, 0);
      P2[e] = make_ulonglong2(T::to_bits(T::min_init()), T::to_bits(T::max_init()));
      if (IS_INT) P3[e] = make_ulonglong2(0, 0);
    }
  }
  __syncthreads();

  const long long n = p.n;
  const long long total_units = (n + GB_UNIT_ROWS - 1) / GB_UNIT_ROWS;
  const long long gwarp = (long long)blockIdx.x * nwarps + warp, total_warps = (long long)gridDim.x * nwarps;
  const int rep = lane & (NG - 1);
  const bool lane_private = NG == 32;
  const unsigned lt_mask = (1u << lane) - 1u;

  GbUnit<VT> cur, nxt;
  if (gwarp < total_units) gb_load_unit<KM, VT>(p, gwarp * GB_UNIT_ROWS, lane, cur);
  for (long long u = gwarp; u < total_units; u += total_warps) {
    const long long base = u * GB_UNIT_ROWS;
    if (u + total_warps < total_units) gb_load_unit<KM, VT>(p, (u + total_warps) * GB_UNIT_ROWS, lane, nxt);

#pragma unroll
    for (int jj = 0; jj < GB_R; jj++) {
      const int j = jj >> 1, h = jj & 1;
      const int bitpos = 2 * lane + h;
      const long long row = base + 64 * j + bitpos;
      const bool active = (cur.act[j] >> bitpos) & 1;
      bool valid = active && !((cur.vn[j] >> bitpos) & 1);
      VT v = T::from_bits(cur.v[jj]);
      if (p.compat_nulls && active && p.val && !valid) { valid = true; v = VT(0); }
      u64 w[NW];
      bool knull = false;
      if (KM == 0) {
        w[0] = cur.k[jj];
        knull = (cur.kn[j] >> bitpos) & 1;
      } else if (active) {
        knull = load_key_generic<NW>(p.ks, row, w);
      }
      // -- key -> dense id through the CTA-shared key table
      int id = -1;
      if (active) {
        if (knull) { id = cap; if (!*reinterpret_cast<volatile uint32_t*>(&misc[1])) *reinterpret_cast<volatile uint32_t*>(&misc[1]) = 1u; }
        else {
          id = sh_lookup<NW>(ktab_key, ktab_id, misc, S, p.sh_log_slots, cap, w);
          if (id < 0) {   // rare: the key does not fit this CTA's table
            long long gs = g_find_or_insert<NW>(p.gt, w);
            if (gs >= 0) g_update_row<VT, FLAGS>(p.gt, gs, p.count_rows != 0, valid, v);
            atomicAdd(&p.gt.counters[CNT_SPILLED], 1ull);
          }
        }
      }
      // -- plain read-modify-write of this warp's records; lanes of the warp that hit the same record
      //    take turns by rank
      const int e = id >= 0 ? id * NG + rep : -(1 + lane);
      int rank = 0, maxr = 0;
      if (!lane_private) {
        unsigned peers = __match_any_sync(0xFFFFFFFFu, e);
        rank = __popc(peers & lt_mask);
        maxr = __reduce_max_sync(0xFFFFFFFFu, id >= 0 ? rank : 0);
      }
      const double x = T::to_f64(v);
      for (int r = 0; r <= maxr; r++) {
        if (id >= 0 && rank == r) {
          if (FLAGS == GB_SUM) {
            ulonglong2 a = P0[e];
            a.y += 1ull + (valid ? (1ull << 32) : 0ull);
            if (valid) {
              if (IS_INT) a.x += (u64)T::to_bits(v);
              else a.x = (u64)__double_as_longlong(__longlong_as_double((long long)a.x) + x);
            }
            P0[e] = a;
          } else {
            ulonglong2 b = P1[e];
            b.x += 1ull + (valid ? (1ull << 32) : 0ull);
            if (valid) {
              ulonglong2 a = P0[e], m = P2[e];
              const bool unset = b.y == 0;
              const bool fin = is_finite_f64(x);
              if (unset && fin) b.y = (u64)__double_as_longlong(x) ^ GB_PIV_X;
              const double piv = b.y ? __longlong_as_double((long long)(b.y ^ GB_PIV_X)) : 0.0;
              const double d = x - piv;
              a.x = (u64)__double_as_longlong(__longlong_as_double((long long)a.x) + d);
              a.y = (u64)__double_as_longlong(__longlong_as_double((long long)a.y) + d * d);
              P0[e] = a;
              const VT mn = T::from_bits(m.x), mx = T::from_bits(m.y);
              if (v < mn || v > mx) {
                if (v < mn) m.x = T::to_bits(v);
                if (v > mx) m.y = T::to_bits(v);
                P2[e] = m;
              }
              if (IS_INT) { ulonglong2 s = P3[e]; s.x += (u64)T::to_bits(v); P3[e] = s; }
            }
            P1[e] = b;
          }
        }
        if (!lane_private) __syncwarp();
      }
    }
    cur = nxt;
  }
  __syncthreads();

  // ---- flush: reduce the warps' private records per group, then one batch update per (CTA, group)
  for (int s = threadIdx.x; s <= S; s += blockDim.x) {
    int id;
    if (s == S) { if (!misc[1]) continue; id = cap; }
    else { uint32_t idw = ktab_id[s]; if (idw == 0 || idw == SH_BUSY) continue; id = (int)idw - 1; }
    u64 rows = 0, nv = 0, isum = 0;
    double S1 = 0.0, S2 = 0.0, c = 0.0;
    bool have_c = false;
    VT mn = T::min_init(), mx = T::max_init();
    for (int wq = 0; wq < nwarps; wq++) {
      const ulonglong2* Q0 = reinterpret_cast<const ulonglong2*>(smem + fixed + (size_t)wq * warp_bytes);
      for (int r = 0; r < NG; r++) {
        const int e = id * NG + r;
        if (FLAGS == GB_SUM) {
          ulonglong2 a = Q0[e];
          rows += a.y & 0xFFFFFFFFull; nv += a.y >> 32;
          if (IS_INT) isum += a.x; else S1 += __longlong_as_double((long long)a.x);
        } else {
          ulonglong2 a = Q0[e], b = Q0[E + e], m = Q0[2 * E + e];
          const u64 n2 = b.x >> 32;
          rows += b.x & 0xFFFFFFFFull;
          if (n2) {
            const double s1 = __longlong_as_double((long long)a.x), s2 = __longlong_as_double((long long)a.y);
            const bool hc2 = b.y != 0;
            const double c2 = hc2 ? __longlong_as_double((long long)(b.y ^ GB_PIV_X)) : 0.0;
            if (!have_c && hc2) {    // adopt the pivot; what was accumulated so far had pivot 0 (non-finite values only)
              const double dl = 0.0 - c2, nn = (double)nv;
              S2 = S2 + 2.0 * dl * S1 + nn * dl * dl; S1 = S1 + nn * dl;
              c = c2; have_c = true;
            }
            const double dl = (hc2 ? c2 : 0.0) - c, nn = (double)n2;
            S1 += s1 + nn * dl;
            S2 += s2 + 2.0 * dl * s1 + nn * dl * dl;
            nv += n2;
            const VT t0 = T::from_bits(m.x), t1 = T::from_bits(m.y);
            if (t0 < mn) mn = t0;
            if (t1 > mx) mx = t1;
            if (IS_INT) isum += Q0[3 * E + e].x;
          }
        }
      }
    }
    long long gs;
    if (id == cap) { gs = p.gt.slots; if (!(ld_cg_u64(&p.gt.hdr[gs].rowsw) & GB_FULL)) atomicOr(&p.gt.hdr[gs].rowsw, GB_FULL); }
    else {
      u64 w[NW];
#pragma unroll
      for (int i = 0; i < NW; i++) w[i] = ktab_key[i * S + s];
      gs = g_find_or_insert<NW>(p.gt, w);
    }
    if (gs < 0) continue;
    u64 mnc = 0, mxo = 0;
    if (FLAGS == GB_ALL) {
      if (T::orderable(mn) && mn != T::min_init()) mnc = ~T::ord(mn);
      if (T::orderable(mx) && mx != T::max_init()) mxo = T::ord(mx);
    }
    if (rows && !p.count_rows) rows = 0;
    if (rows) atomicAdd(&p.gt.hdr[gs].rowsw, rows);
    g_update_batch<FLAGS, IS_INT>(p.gt, gs, 0ull, nv, c, have_c, S1, S2, isum, mnc, mxo);
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
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * R) {
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
      if (row >= n) continue;
      if (p.fbits) {
        if (!pdrs_bit(p.fbits, row)) continue;
        if (p.fnull && pdrs_bit(p.fnull, row)) continue;
      }
      bool valid = vals && !(p.vnull && pdrs_bit(p.vnull, row));
      VT vv = v[j];
      if (p.compat_nulls && vals && !valid) { valid = true; vv = VT(0); }
      u64 w[NW];
      bool knull = false;
      if (KM == 0) { w[0] = k0[j]; if (knull0) knull = pdrs_bit(knull0, row); }
      else knull = load_key_generic<NW>(p.ks, row, w);
      long long gs = knull ? p.gt.slots : g_find_or_insert<NW>(p.gt, w);
      if (knull && !(ld_cg_u64(&p.gt.hdr[gs].rowsw) & GB_FULL)) atomicOr(&p.gt.hdr[gs].rowsw, GB_FULL);
      if (gs >= 0) g_update_row<VT, FLAGS>(p.gt, gs, p.count_rows != 0, valid, vv);
    }
  }
  (void)sizeof(T);
}

// ---------------------------------------------------------------- sampling (cardinality estimate)
// Inserts the keys of `nblocks` evenly spread runs of 256 rows into a scratch table; CNT_NGROUPS then
// holds the number of distinct sampled keys.
template <int NW, int KM>
__global__ void gb_sample_kernel(const GbParams p, long long nblocks, long long block_stride_rows) {
  for (long long b = blockIdx.x; b < nblocks; b += gridDim.x) {
    long long row = b * block_stride_rows + threadIdx.x;
    if (row >= p.n) continue;
    u64 w[NW];
    bool knull = false;
    if (KM == 0) { w[0] = (u64)__ldg(reinterpret_cast<const long long*>(p.ks.c[0].data) + row); knull = p.ks.c[0].nulls && pdrs_bit(p.ks.c[0].nulls, row); }
    else knull = load_key_generic<NW>(p.ks, row, w);
    if (!knull) g_find_or_insert<NW>(p.gt, w);
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
