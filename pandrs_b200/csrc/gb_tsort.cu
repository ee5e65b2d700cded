// Low / medium-cardinality groupby (tens to ~2000 groups, one 64-bit key column): tile sort + register reduction.
//
// Replaces the row loop of grouping.rs:62-104 and the per-group gathers of aggregation.rs:507-742.
//
// Measured on B200 (profiles/microbench2_r01.txt): a random shared-memory access costs ~2.6 SM-cycles per warp
// instruction per 4 bytes and 64-bit shared atomics are CAS loops, so updating {S1, S2, min, max, n} per row in
// shared memory costs >= 50 cycles per 32 rows, while the HBM roofline allows ~22.  This kernel touches shared
// memory per row only to SORT the values of a tile by group:
//
//   phase 1  key -> group id (direct-mapped for dense integer keys, else a CTA-shared key table); one native
//            32-bit shared atomic per row on the tile histogram, whose return value is the row's rank in its group
//   phase 2  exclusive scan of the histogram (1024 threads, two barriers)
//   phase 3  value -> sorted[offset[group] + rank]                      (one LDS.32 + one STS.64 per row)
//   phase 4  every thread OWNS one (or two) groups for the whole kernel: it walks its segment of the sorted tile
//            and keeps rows / n / pivot / S1 / S2 / min / max / isum in REGISTERS - all six aggregates cost the
//            same single sequential read.  Segments longer than ts_heavy (128) rows (skewed keys, few groups) are reduced by the
//            owner's whole warp with shuffles.
//
// Keys are loaded with 128-bit streaming loads one tile ahead (registers); values arrive through one bulk
// asynchronous copy (cp.async.bulk, completion on an mbarrier) per tile into a staging buffer, issued a whole
// tile period before they are consumed.  At the end every thread flushes its groups into the global table as
// one pre-aggregated batch (same state and finalisation as the other groupby kernels).
#include <algorithm>

#include "groupby_kernels.cuh"

#ifdef TS_PROFILE
__device__ unsigned long long ts_prof[8];
#define TSP_MARK(k) do { const long long _t = clock64(); tacc[k] += (unsigned long long)(_t - tprev); tprev = _t; } while (0)
#else
#define TSP_MARK(k)
#endif

namespace {

constexpr int TS_T = 8192;               // rows per tile (4096 in the two-CTAs-per-SM geometry of 256 threads)
template <int NT> struct TsGeom { static constexpr int TT = NT == 256 ? 4096 : 8192; static constexpr int CTAS = NT == 256 ? 2 : 1; };

// Group id -> histogram position.  XORs the low bits of the id into the warp field so that consecutive ids (dense
// keys, or first-seen order under skew) are owned by different warps, while the lanes of a warp own consecutive
// positions.  An involution: the same function maps a position back to its id.
template <int NT>
__device__ __forceinline__ uint32_t ts_perm(uint32_t g) { return g ^ ((g << 5) & (uint32_t)((NT / 32 - 1) << 5)); }

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(cnt) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int RPT> struct TsRows {
  u64 k[RPT];
  uint32_t vnw, knw, fw;   // lane l holds bitmap word (l % RPT) of the warp's 32 * RPT rows: value NULLs, key NULLs, filter & ~filter NULLs
  uint32_t fl[RPT / 2];    // partitioned input: the flag bytes of rows 2 j, 2 j + 1 (bit 0 / bit 8: value is NULL)
  uint32_t act;            // bit q: row q of this lane is inside the column / partition
  uint32_t knm;            // generic (packed) keys: bit q: the key of row q is NULL (single-column keys only)
};

// One unit of work of a CTA: a tile of up to TS_T consecutive rows.  Plain input: tiles blockIdx, blockIdx + grid, ...
// of the columns.  Partitioned input: the tiles of partitions blockIdx, blockIdx + grid, ...; `last` marks the last
// tile of a partition (the CTA then flushes its groups and starts over with empty tables).
struct TsItem { long long base; long long part; int valid; int last; };   // partitioned input: `part` = chunk index
// tiles [t0, t1) of chunk `ch` (partition ch / cpp, chunk ch % cpp of it) that hold rows
template <int TT>
__device__ __forceinline__ bool ts_chunk_range(const GbParams& p, long long ch, long long* pbase, long long* cnt, long long* t0, long long* t1) {
  const long long q = ch / p.part_cpp, k = ch % p.part_cpp;
  *cnt = min((long long)__ldg(p.part_cnt + q), p.part_cap);
  *pbase = q * p.part_cap;
  const long long tiles = (*cnt + TT - 1) / TT;
  *t0 = k * p.part_chunk_tiles;
  *t1 = min(tiles, *t0 + p.part_chunk_tiles);
  return *t0 < *t1;
}
template <bool PART, int TT>
__device__ __forceinline__ TsItem ts_first_item(const GbParams& p) {
  TsItem it;
  it.part = blockIdx.x; it.last = 0; it.base = 0; it.valid = 0;
  if (!PART) {
    it.base = (long long)blockIdx.x * TT;
    it.valid = it.base < p.n ? (int)min((long long)TT, p.n - it.base) : 0;
    return it;
  }
  const long long nchunks = (long long)p.part_n * p.part_cpp;
  for (; it.part < nchunks; it.part += gridDim.x) {
    long long pb, c, t0, t1;
    if (ts_chunk_range<TT>(p, it.part, &pb, &c, &t0, &t1)) { it.base = pb + t0 * TT; it.valid = (int)min((long long)TT, c - t0 * TT); it.last = t0 + 1 == t1; return it; }
  }
  return it;
}
template <bool PART, int TT>
__device__ __forceinline__ TsItem ts_next_item(const GbParams& p, const TsItem& cur) {
  TsItem it = cur;
  if (!PART) {
    it.base = cur.base + (long long)gridDim.x * TT;
    it.valid = it.base < p.n ? (int)min((long long)TT, p.n - it.base) : 0;
    return it;
  }
  long long pb, c, t0, t1;
  if (!cur.last) {
    ts_chunk_range<TT>(p, cur.part, &pb, &c, &t0, &t1);
    const long long t = (cur.base - pb) / TT + 1;
    it.base = cur.base + TT; it.valid = (int)min((long long)TT, c - t * TT); it.last = t + 1 == t1;
    return it;
  }
  const long long nchunks = (long long)p.part_n * p.part_cpp;
  it.valid = 0; it.last = 0;
  for (it.part = cur.part + gridDim.x; it.part < nchunks; it.part += gridDim.x) {
    if (ts_chunk_range<TT>(p, it.part, &pb, &c, &t0, &t1)) { it.base = pb + t0 * TT; it.valid = (int)min((long long)TT, c - t0 * TT); it.last = t0 + 1 == t1; return it; }
  }
  return it;
}

// Rows of one warp in one tile: load j, lane l -> rows wbase + 64 j + 2 l + {0, 1} (one 128-bit load).
// Partitioned input: the tile's memory always exists (partitions are padded to whole tiles), rows >= valid are masked.
template <int RPT, bool PLAIN, bool PART, bool GENERIC, bool PACK32>
__device__ __forceinline__ void ts_load(const GbParams& p, const TsItem& it, int warp, int lane, TsRows<RPT>& r) {
  const long long wbase = it.base + (long long)warp * (32 * RPT);
  const int wfirst = warp * (32 * RPT);           // first row of the warp inside the tile
  const u64* keys = (PART ? p.part_keys : reinterpret_cast<const u64*>(p.ks.c[0].data)) + wbase + 2 * lane;
  r.knm = 0;
  if (PACK32) {           // one or two 4-byte key columns without null bitmaps: RAW loads only (two rows of column k, chunk j in
                          // r.k[k * RPT / 2 + j]); nothing is consumed here, the tuple is packed at the start of phase 1
    r.act = 0;
#pragma unroll
    for (int q = 0; q < RPT; q++) { r.k[q] = 0; if (wfirst + 64 * (q >> 1) + 2 * lane + (q & 1) < it.valid) r.act |= 1u << q; }
#pragma unroll
    for (int k = 0; k < 2; k++) {
      if (k < p.ks.nkeys) {
        const uint32_t* d = reinterpret_cast<const uint32_t*>(p.ks.c[k].data) + wbase + 2 * lane;
#pragma unroll
        for (int j = 0; j < RPT / 2; j++) {
          if ((r.act >> (2 * j + 1)) & 1u) { const uint2 t = __ldcs(reinterpret_cast<const uint2*>(d + 64 * j)); r.k[k * (RPT / 2) + j] = (u64)t.x | ((u64)t.y << 32); }
          else if ((r.act >> (2 * j)) & 1u) r.k[k * (RPT / 2) + j] = (u64)__ldg(d + 64 * j);
        }
      }
    }
  } else if (GENERIC) {   // any key tuple that packs into one 64-bit word (dictionary ids, i32, bool, pairs of them)
    // column by column, two adjacent rows per load (the packing of load_key_generic, restated without per-row calls)
    r.act = 0;
#pragma unroll
    for (int q = 0; q < RPT; q++) { r.k[q] = 0; if (wfirst + 64 * (q >> 1) + 2 * lane + (q & 1) < it.valid) r.act |= 1u << q; }
    for (int k = 0; k < p.ks.nkeys; k++) {
      const KeyColDev c = p.ks.c[k];
#pragma unroll
      for (int j = 0; j < RPT / 2; j++) {
        const long long row0 = wbase + 64 * j + 2 * lane;
        const bool in0 = (r.act >> (2 * j)) & 1u, in1 = (r.act >> (2 * j + 1)) & 1u;
        u64 v0 = 0, v1 = 0;
        bool n0 = false, n1 = false;
        if (c.nulls && in0) { const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(c.nulls) + (row0 >> 5)); n0 = (w >> (row0 & 31)) & 1u; n1 = (w >> ((row0 & 31) + 1)) & 1u; }
        switch (c.dtype) {
          case PDRS_I64: case PDRS_F64: {
            const u64* d = reinterpret_cast<const u64*>(c.data) + row0;
            if (in1) { const ulonglong2 t = ld_stream_v2(d); v0 = t.x; v1 = t.y; } else if (in0) v0 = __ldg(d);
            if (c.dtype == PDRS_I64) { if (in0) v0 -= (u64)c.offset; if (in1) v1 -= (u64)c.offset; }      // range-compressed tuples (offset 0 otherwise)
            if (c.dtype == PDRS_F64) {       // all NaNs print "NaN"
              if (__longlong_as_double((long long)v0) != __longlong_as_double((long long)v0)) v0 = 0x7FF8000000000000ull;
              if (__longlong_as_double((long long)v1) != __longlong_as_double((long long)v1)) v1 = 0x7FF8000000000000ull;
            }
            break;
          }
          case PDRS_I32: case PDRS_DICT_U32: {
            const uint32_t* d = reinterpret_cast<const uint32_t*>(c.data) + row0;
            if (in1) { const uint2 t = __ldg(reinterpret_cast<const uint2*>(d)); v0 = t.x; v1 = t.y; } else if (in0) v0 = __ldg(d);
            if (c.dtype == PDRS_DICT_U32) { if ((long long)v0 == c.null_alias) n0 = true; if ((long long)v1 == c.null_alias) n1 = true; if (in0) v0 -= (u64)c.offset; if (in1) v1 -= (u64)c.offset; }
            else {
              if (in0) v0 = ((u64)(long long)(int)(uint32_t)v0 - (u64)c.offset) & 0xFFFFFFFFull;
              if (in1) v1 = ((u64)(long long)(int)(uint32_t)v1 - (u64)c.offset) & 0xFFFFFFFFull;
            }
            break;
          }
          default: {                           // PDRS_BOOL_BITS
            if (in0) { const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(c.data) + (row0 >> 5)); v0 = (w >> (row0 & 31)) & 1u; v1 = (w >> ((row0 & 31) + 1)) & 1u; }
            break;
          }
        }
        if (n0) { if (p.ks.single_null) r.knm |= 1u << (2 * j); else if (c.nword >= 0) r.k[2 * j] |= 1ull << c.nshift; }
        else r.k[2 * j] |= v0 << c.shift;
        if (n1) { if (p.ks.single_null) r.knm |= 1u << (2 * j + 1); else if (c.nword >= 0) r.k[2 * j + 1] |= 1ull << c.nshift; }
        else r.k[2 * j + 1] |= v1 << c.shift;
      }
    }
  } else if (PART || wfirst + 32 * RPT <= it.valid) {
#pragma unroll
    for (int j = 0; j < RPT / 2; j++) {
      const ulonglong2 kk = ld_stream_v2(keys + 64 * j);
      r.k[2 * j] = kk.x; r.k[2 * j + 1] = kk.y;
    }
  }
  if (GENERIC || PACK32) {
  } else if (wfirst + 32 * RPT <= it.valid) r.act = (1u << RPT) - 1u;
  else {
    r.act = 0;
#pragma unroll
    for (int q = 0; q < RPT; q++) {
      const bool inb = wfirst + 64 * (q >> 1) + 2 * lane + (q & 1) < it.valid;
      if (!PART) r.k[q] = inb ? __ldg(keys + 64 * (q >> 1) + (q & 1)) : 0ull;
      if (inb) r.act |= 1u << q;
    }
  }
  r.vnw = 0;
  if (PART) {
#pragma unroll
    for (int j = 0; j < RPT / 2; j++) r.fl[j] = p.part_flags ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p.part_flags + wbase + 2 * lane) + 32 * j) : 0u;
  }
  if (!PART) {
    const long long w = (wbase >> 5) + (lane & (RPT - 1));
    const bool inb = w * 32 < p.n;           // bitmaps cover ceil(n / 64) * 8 bytes (pdrs_view_col)
    r.vnw = (p.vnull && inb) ? __ldg(reinterpret_cast<const uint32_t*>(p.vnull) + w) : 0u;
    if (!PLAIN) {
      r.knw = (!GENERIC && !PACK32 && p.ks.c[0].nulls && inb) ? __ldg(reinterpret_cast<const uint32_t*>(p.ks.c[0].nulls) + w) : 0u;
      uint32_t f = 0xFFFFFFFFu;
      if (p.fbits) {   // filter keeps Some(true) rows only (data_ops.rs:49-55)
        f = inb ? __ldg(reinterpret_cast<const uint32_t*>(p.fbits) + w) : 0u;
        if (p.fnull && inb) f &= ~__ldg(reinterpret_cast<const uint32_t*>(p.fnull) + w);
      }
      r.fw = f;
    }
  }
}

template <typename VT, int FLAGS> struct TsAcc {
  uint32_t rows, n;
  u64 piv;            // pivot bits (LSB forced to 1), 0 = unset
  double S1, S2;
  VT mn, mx;
  u64 isum;
};

// NaN operands never win a comparison, i.e. they are ignored like f64::min / f64::max do (aggregation.rs:649-674)
template <typename VT, int FLAGS>
__device__ __forceinline__ void ts_add(double& S1, double& S2, VT& mn, VT& mx, u64& isum, double pv, u64 bits) {
  using T = ValTraits<VT>;
  const VT v = T::from_bits(bits);
  if (T::is_int) isum += bits;
  if (FLAGS == GB_ALL) {
    const double d = T::to_f64(v) - pv;
    S1 += d;
    S2 = fma(d, d, S2);
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
  } else if (!T::is_int) {
    S1 += T::to_f64(v);
  }
}

// Top 32 bits of (fold(k) * odd constant); the radix partitions take the top bits, the CTA key table the bits below.
__device__ __host__ __forceinline__ uint32_t ts_hash32(u64 k) {
  const uint32_t lo = (uint32_t)k ^ (uint32_t)(k >> 32), hi = (uint32_t)(k >> 32);
  constexpr uint32_t CL = 0x7F4A7C15u, CH = 0x9E3779B9u;
  return (uint32_t)(((u64)lo * CL) >> 32) + lo * CH + hi * CL;
}

// CTA key table of the hashed modes: open addressing over S slots {key, id + 1}; ids are handed out in first-seen
// order.  Same no-spin protocol as sh_try / sh_lookup (groupby_kernels.cuh): every lane of the warp calls together.
__device__ __noinline__ int ts_insert(u64* ktab_key, uint32_t* ktab_id, uint32_t* misc, int S, int cap, u64 key, uint32_t slot, bool active) {
  int probe = 0, res = -1, rounds = 0;
  bool pending = active;
  while (__any_sync(0xFFFFFFFFu, pending)) {
    if (pending) {
      int r = SH_RETRY;
      while (probe < S) {
        const uint32_t idw = *reinterpret_cast<volatile uint32_t*>(&ktab_id[slot]);
        const u64 kk = *reinterpret_cast<volatile u64*>(&ktab_key[slot]);
        if (idw == 0) {
          if (*reinterpret_cast<volatile uint32_t*>(&misc[0]) >= (uint32_t)cap) { r = -1; break; }
          const uint32_t old = atomicCAS(&ktab_id[slot], 0u, SH_BUSY);
          if (old == 0) {
            const uint32_t nid = atomicAdd(&misc[0], 1u);
            if (nid >= (uint32_t)cap) { *reinterpret_cast<volatile uint32_t*>(&ktab_id[slot]) = 0; r = -1; break; }   // table full: give the slot back, spill the row
            *reinterpret_cast<volatile u64*>(&ktab_key[slot]) = key;
            __threadfence_block();
            *reinterpret_cast<volatile uint32_t*>(&ktab_id[slot]) = nid + 1;
            r = (int)nid;
          }
          break;                       // lost the race: look at the slot again next round
        }
        if (idw == SH_BUSY) break;
        if (kk == key) { r = (int)idw - 1; break; }
        slot = (slot + 1) & (S - 1);
        probe++;
      }
      if (probe >= S) r = -1;
      if (r != SH_RETRY) { res = r; pending = false; }
    }
    if (++rounds > (1 << 20)) break;
  }
  return res;
}

// KMODE: 0 = direct-mapped ids (small dense integer keys), 1 = CTA key table, 2 = CTA key table over hash-partitioned rows,
//        3 = CTA key table over generic key tuples packed into one 64-bit word (load_key_generic),
//        4 = the same for one or two 4-byte key columns without null bitmaps (raw loads stay in flight, packed in phase 1).
// Histogram word of a group in a tile: low 16 bits = rows with a value (after the scan: offset of the group's
// segment), high 16 bits = rows whose value is NULL.  One native atomic per row serves both counts.
template <int NT, typename VT, int FLAGS, int GPT, int KMODE, bool PLAIN, bool TEAM>
__global__ void __launch_bounds__(NT, TsGeom<NT>::CTAS) gb_tsort_kernel(const GbParams p) {
  using T = ValTraits<VT>;
  constexpr bool IS_INT = T::is_int;
  constexpr bool ALL = FLAGS == GB_ALL;
  constexpr bool DENSE = KMODE == 0, PART = KMODE == 2, PACK32 = KMODE == 4, GENERIC = KMODE == 3 || KMODE == 4;
  constexpr int TT = TsGeom<NT>::TT;
  constexpr int RPT = TT / NT, WROWS = 32 * RPT, NWARPS = NT / 32;
  constexpr int NP = NT * GPT, NPAD = NP + 32;
  constexpr uint32_t TRASH = NP + 8;                                 // histogram slot of rows that are not aggregated here
  extern __shared__ __align__(128) unsigned char smem[];
  u64* sorted = reinterpret_cast<u64*>(smem);                        // [TT]
  u64* stage = sorted + TT;                                        // [TT]
  uint32_t* H = reinterpret_cast<uint32_t*>(stage + TT);           // [2][NPAD]
  uint32_t* wsum = H + 2 * NPAD;                                     // [32]
  uint32_t* misc = wsum + 32;                                        // [4]  0: ids handed out
  u64* mbar = reinterpret_cast<u64*>(misc + 4);                      // [2]
  u64* ktab_key = mbar + 2;                                          // [S]  (hashed keys only)
  const int S = p.sh_slots, cap = p.sh_cap;
  uint32_t* ktab_id = reinterpret_cast<uint32_t*>(ktab_key + S);     // [S]
  u64* idkey = reinterpret_cast<u64*>(ktab_id + S);                  // [cap + 1] id -> key (hashed keys only)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t a_mbar = sm_addr(mbar), a_stage = sm_addr(stage), a_sorted = sm_addr(sorted);
  const u64* vals = PART ? p.part_vals : reinterpret_cast<const u64*>(p.val);
  const u64 dense_base = (u64)p.sh_dense_base;
  const int slot_lsh = PART ? p.part_bits : 0, slot_rsh = 32 - p.sh_log_slots;   // key table slot = (hash32 << lsh) >> rsh

  for (int i = tid; i < 2 * NPAD; i += NT) H[i] = 0;
  if (!DENSE) for (int i = tid; i < S; i += NT) ktab_id[i] = 0;
  if (tid < 4) misc[tid] = 0;
  if (tid == 0) { mbar_init(a_mbar, 1); fence_proxy_async(); }
  __syncthreads();

  TsAcc<VT, FLAGS> acc[GPT];
  auto reset_acc = [&]() {
#pragma unroll
    for (int s = 0; s < GPT; s++) { acc[s].rows = 0; acc[s].n = 0; acc[s].piv = 0; acc[s].S1 = 0.0; acc[s].S2 = 0.0; acc[s].mn = T::min_init(); acc[s].mx = T::max_init(); acc[s].isum = 0; }
  };
  reset_acc();

  // values of a tile -> stage: one bulk copy when the whole tile exists in memory, a plain copy loop otherwise
  auto tile_bulk = [&](const TsItem& it) { return PART || it.valid == TT; };
  auto issue_vals = [&](const TsItem& it) {
    if (tile_bulk(it)) {
      if (tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(a_mbar, TT * 8);
#pragma unroll
        for (int c = 0; c < 4; c++) bulk_g2s(a_stage + c * (TT * 2), vals + it.base + c * (TT / 4), TT * 2, a_mbar);
      }
    } else {
      for (int i = tid; i < TT; i += NT) stage[i] = i < it.valid ? __ldg(vals + it.base + i) : 0ull;
    }
  };
  // one pre-aggregated batch per (CTA, group) into the global table
  auto flush = [&]() {
    __syncthreads();
#pragma unroll
    for (int s = 0; s < GPT; s++) {
      const uint32_t gid = ts_perm<NT>((uint32_t)(tid + s * NT));
      const bool have = acc[s].rows != 0;
      const bool nullgroup = have && gid == (uint32_t)cap;
      u64 w[1] = {0};
      if (have && !nullgroup) w[0] = DENSE ? dense_base + gid : idkey[gid];
      long long gs = g_find_or_insert<1>(p.gt, w, have && !nullgroup);
      if (nullgroup) { gs = p.gt.slots; if (!(ld_cg_u64(&p.gt.hdr[gs].rowsw) & GB_FULL)) atomicOr(&p.gt.hdr[gs].rowsw, GB_FULL); }
      if (!have || gs < 0) continue;
      if (p.count_rows) atomicAdd(&p.gt.hdr[gs].rowsw, (u64)acc[s].rows);
      u64 mnc = 0, mxo = 0;
      if (ALL && acc[s].n) { mnc = ~T::ord(acc[s].mn); mxo = T::ord(acc[s].mx); }
      g_update_batch<FLAGS, IS_INT>(p.gt, gs, 0ull, (u64)acc[s].n, __longlong_as_double((long long)acc[s].piv), acc[s].piv != 0, acc[s].S1, acc[s].S2, acc[s].isum, mnc, mxo);
    }
  };

  TsItem cur = ts_first_item<PART, TT>(p);
  TsRows<RPT> r;
  if (cur.valid) { issue_vals(cur); ts_load<RPT, PLAIN, PART, GENERIC && !PACK32, PACK32>(p, cur, warp, lane, r); }
  uint32_t tma_phase = 0;
  const int sh2 = (2 * lane) & 31;
  const uint32_t vm0 = 1u << sh2, vm1 = 2u << sh2;      // this lane's two bits in a bitmap word
  int b = 0;
#pragma unroll 1
#ifdef TS_PROFILE
  unsigned long long tacc[6] = {0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#endif
  for (; cur.valid; b ^= 1) {
    const TsItem nxt = ts_next_item<PART, TT>(p, cur);
    uint32_t* Hc = H + b * NPAD;
    const uint32_t a_h = sm_addr(Hc);
    const bool tile_tma = tile_bulk(cur);
    // ---- phase 1: group ids, tile histogram (the atomic's return value ranks the row inside its group)
    uint32_t pack[RPT];
    uint32_t skipmask = 0, zeromask = 0;
    if (PACK32) {        // raw 4-byte key columns -> the packed key tuple of every row (the layout of load_key_generic)
      u64 t[RPT];
      uint32_t knm = 0;
      const KeyColDev c0 = p.ks.c[0], c1 = p.ks.c[1];
      const bool two = p.ks.nkeys == 2;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        const int j = q >> 1, h = q & 1;
        u64 word = 0;
        const uint32_t a = (uint32_t)(r.k[j] >> (32 * h));
        if (c0.dtype == PDRS_DICT_U32 && (long long)a == c0.null_alias) { if (p.ks.single_null) knm |= 1u << q; else if (c0.nword >= 0) word |= 1ull << c0.nshift; }
        else word |= (u64)a << c0.shift;
        if (two) {
          const uint32_t bq = (uint32_t)(r.k[RPT / 2 + j] >> (32 * h));
          if (c1.dtype == PDRS_DICT_U32 && (long long)bq == c1.null_alias) { if (c1.nword >= 0) word |= 1ull << c1.nshift; }
          else word |= (u64)bq << c1.shift;
        }
        t[q] = word;
      }
#pragma unroll
      for (int q = 0; q < RPT; q++) r.k[q] = t[q];
      r.knm = knm;
    }
    const bool fullw = __all_sync(0xFFFFFFFFu, r.act == (1u << RPT) - 1u);
    if (PLAIN && DENSE && fullw) {
      // every row is inside the column, no filter, no NULL keys, direct-mapped ids: ~16 instructions per row
      bool bad = false;
#pragma unroll
      for (int j = 0; j < RPT / 2; j++) {
        const uint32_t vw = __shfl_sync(0xFFFFFFFFu, r.vnw, 2 * j + (lane >> 4));
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int q = 2 * j + h;
          const u64 off = r.k[q] - dense_base;
          const bool ok = off < (u64)cap;
          const bool vnull = (vw & (h ? vm1 : vm0)) != 0;
          const uint32_t pp = ok ? ts_perm<NT>((uint32_t)off) : TRASH;
          const uint32_t old = sm_atom_add32(a_h + pp * 4u, vnull ? 0x10000u : 1u);
          bad = bad || !ok;
          if (!ok || vnull) skipmask |= 1u << q;
          pack[q] = pp | (old << 16);
        }
      }
      if (__any_sync(0xFFFFFFFFu, bad)) {   // rare: key outside the dense range -> global table
#pragma unroll 1
        for (int q = 0; q < RPT; q++) {
          u64 kq = 0;
#pragma unroll
          for (int qq = 0; qq < RPT; qq++) if (qq == q) kq = r.k[qq];
          const bool sp = kq - dense_base >= (u64)cap;
          const long long row = cur.base + warp * WROWS + 64 * (q >> 1) + 2 * lane + (q & 1);
          u64 vb = 0;
          bool vnull = false;
          if (sp) { vb = __ldg(vals + row); vnull = p.vnull && pdrs_bit(p.vnull, row); }
          gb_spill_rows<1, VT, FLAGS>(p.gt, kq, 0ull, 0ull, sp, p.count_rows != 0, !vnull, T::from_bits(vb));
        }
      }
    } else if (PLAIN && !DENSE) {
      // no filter, no NULL keys; key -> id through the CTA key table: up to three lock-free probes per row
      // (the table runs at a load factor <= 1/4, so > 90% of the rows hit their home slot), insertion of new keys
      // and longer probe sequences in the warp-synchronous slow path
      uint32_t slowmask = 0, pend = 0, ids[RPT];
      // home slot of every row: one lock-free probe each, all loads in flight together
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        const uint32_t slot = (ts_hash32(r.k[q]) << slot_lsh) >> slot_rsh;
        const uint32_t idw = *reinterpret_cast<volatile uint32_t*>(&ktab_id[slot]);
        const u64 kk = *reinterpret_cast<volatile u64*>(&ktab_key[slot]);
        const bool live = (r.act >> q) & 1u;
        bool hit = idw != 0 && idw != SH_BUSY && kk == r.k[q];
        ids[q] = hit ? idw - 1 : 0xFFFFFFFFu;
        if (GENERIC && ((r.knm >> q) & 1u)) { hit = true; ids[q] = (uint32_t)cap; }      // the NULL-key group
        if (live && !hit) { if (idw == 0 || idw == SH_BUSY) slowmask |= 1u << q; else pend |= 1u << q; }   // new key -> insertion; else go on probing
      }
      // rows displaced from their home slot (~10% at load factor 1/4): every lane follows the probe sequence of its
      // first pending row per round, so a round costs the warp one shared-memory round trip, not one per row
      {
        uint32_t dist = 1;
        int cur = -1;
        while (__any_sync(0xFFFFFFFFu, pend != 0)) {
          const int j = pend ? __ffs(pend) - 1 : -1;
          if (j != cur) { cur = j; dist = 1; }
          u64 kq = 0;
#pragma unroll
          for (int qq = 0; qq < RPT; qq++) if (qq == j) kq = r.k[qq];
          if (j >= 0) {
            const uint32_t slot = (((ts_hash32(kq) << slot_lsh) >> slot_rsh) + dist) & (uint32_t)(S - 1);
            const uint32_t idw = *reinterpret_cast<volatile uint32_t*>(&ktab_id[slot]);
            const u64 kk = *reinterpret_cast<volatile u64*>(&ktab_key[slot]);
            if (idw == 0 || idw == SH_BUSY || dist >= 64u) { slowmask |= 1u << j; pend &= ~(1u << j); }
            else if (kk == kq) {
#pragma unroll
              for (int qq = 0; qq < RPT; qq++) if (qq == j) ids[qq] = idw - 1;
              pend &= ~(1u << j);
            } else dist++;
          }
        }
      }
      // new keys (every row of the first tile of a partition): each lane inserts the key of its FIRST unresolved row
      // through the warp-synchronous routine, then probes its other unresolved rows again without locks - after one
      // round most keys of the tile are in the table, so a cold tile costs 2-3 rounds, not one insertion per row
      uint32_t spillmask = 0;
      while (__any_sync(0xFFFFFFFFu, slowmask != 0)) {
        const int j = slowmask ? __ffs(slowmask) - 1 : -1;
        u64 kq = 0;
#pragma unroll
        for (int qq = 0; qq < RPT; qq++) if (qq == j) kq = r.k[qq];
        const int id = ts_insert(ktab_key, ktab_id, misc, S, cap, kq, (ts_hash32(kq) << slot_lsh) >> slot_rsh, j >= 0);
        if (j >= 0) {
          if (id >= 0) {
            idkey[id] = kq;            // benign duplicate stores of the same value
#pragma unroll
            for (int qq = 0; qq < RPT; qq++) if (qq == j) ids[qq] = (uint32_t)id;
          } else spillmask |= 1u << j;
          slowmask &= ~(1u << j);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < RPT; q++) {
          if (!((slowmask >> q) & 1u)) continue;
          uint32_t slot = (ts_hash32(r.k[q]) << slot_lsh) >> slot_rsh;
#pragma unroll 1
          for (int pr = 0; pr < 4; pr++) {
            const uint32_t idw = *reinterpret_cast<volatile uint32_t*>(&ktab_id[slot]);
            const u64 kk = *reinterpret_cast<volatile u64*>(&ktab_key[slot]);
            if (idw == 0 || idw == SH_BUSY) break;                       // not there (yet): stays for the next round
            if (kk == r.k[q]) { ids[q] = idw - 1; slowmask &= ~(1u << q); break; }
            slot = (slot + 1) & (uint32_t)(S - 1);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < RPT / 2; j++) {
        const uint32_t vw = PART ? 0u : __shfl_sync(0xFFFFFFFFu, r.vnw, 2 * j + (lane >> 4));
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int q = 2 * j + h;
          const bool ok = ids[q] != 0xFFFFFFFFu && ((r.act >> q) & 1u);
          const bool vnull = PART ? ((r.fl[j] >> (8 * h)) & 1u) != 0 : (vw & (h ? vm1 : vm0)) != 0;
          const uint32_t pp = ok ? ts_perm<NT>(ids[q]) : TRASH;
          const uint32_t old = sm_atom_add32(a_h + pp * 4u, vnull ? 0x10000u : 1u);
          if (!ok || vnull) skipmask |= 1u << q;
          pack[q] = pp | (old << 16);
        }
      }
      if (__any_sync(0xFFFFFFFFu, spillmask != 0)) {   // rare: CTA key table full -> global table
#pragma unroll 1
        for (int q = 0; q < RPT; q++) {
          const bool sp = (spillmask >> q) & 1u;
          const long long row = cur.base + warp * WROWS + 64 * (q >> 1) + 2 * lane + (q & 1);
          u64 vb = 0, kq = 0;
          bool vnull = false;
#pragma unroll
          for (int qq = 0; qq < RPT; qq++) if (qq == q) kq = r.k[qq];
          if (sp) { vb = __ldg(vals + row); vnull = PART ? (p.part_flags && (p.part_flags[row] & 1)) : (p.vnull && pdrs_bit(p.vnull, row)); }
          gb_spill_rows<1, VT, FLAGS>(p.gt, kq, 0ull, 0ull, sp, p.count_rows != 0, !vnull, T::from_bits(vb));
        }
      }
    } else {
      uint32_t spillmask = 0;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        const int j = q >> 1, h = q & 1;
        const int wsrc = 2 * j + (lane >> 4);                                   // bitmap word of this lane's chunk-j rows
        bool vnull = (__shfl_sync(0xFFFFFFFFu, r.vnw, wsrc) & (h ? vm1 : vm0)) != 0;
        bool active = (r.act >> q) & 1u, knull = false;
        if (!PLAIN) {
          const uint32_t fword = __shfl_sync(0xFFFFFFFFu, r.fw, wsrc), kword = __shfl_sync(0xFFFFFFFFu, r.knw, wsrc);   // every lane shuffles: no short-circuit
          active = active && (fword & (h ? vm1 : vm0)) != 0;
          knull = GENERIC ? ((r.knm >> q) & 1u) != 0 : (kword & (h ? vm1 : vm0)) != 0;
          if (p.compat_nulls && vnull) { vnull = false; zeromask |= 1u << q; }   // filter + compat_filter_nulls (data_ops.rs:64-71)
        }
        uint32_t gid = 0;
        bool ok;
        if (DENSE) {
          const u64 off = r.k[q] - dense_base;
          ok = off < (u64)cap;
          gid = (uint32_t)off;
        } else {
          const int id = ts_insert(ktab_key, ktab_id, misc, S, cap, r.k[q], (ts_hash32(r.k[q]) << slot_lsh) >> slot_rsh, active && !knull);
          ok = active && !knull && id >= 0;
          if (ok) { gid = (uint32_t)id; idkey[id] = r.k[q]; }
        }
        if (!PLAIN) { ok = ok && !knull; if (knull) { gid = (uint32_t)cap; ok = true; } }
        if (active && !ok) spillmask |= 1u << q;
        ok = ok && active;
        const uint32_t pp = ok ? ts_perm<NT>(gid) : TRASH;
        const uint32_t old = sm_atom_add32(a_h + pp * 4u, vnull ? 0x10000u : 1u);
        if (!ok || vnull) skipmask |= 1u << q;
        pack[q] = pp | (old << 16);
      }
      if (__any_sync(0xFFFFFFFFu, spillmask != 0)) {   // rare: key outside the dense range / CTA key table full -> global table
#pragma unroll 1
        for (int q = 0; q < RPT; q++) {
          const bool sp = (spillmask >> q) & 1u;
          const long long row = cur.base + warp * WROWS + 64 * (q >> 1) + 2 * lane + (q & 1);
          u64 vb = 0, kq = 0;
          bool vnull = false;
#pragma unroll
          for (int qq = 0; qq < RPT; qq++) if (qq == q) kq = r.k[qq];
          if (sp) {
            vb = __ldg(vals + row);
            vnull = !PART && p.vnull && pdrs_bit(p.vnull, row);
            if (vnull && p.compat_nulls) { vnull = false; vb = 0; }
          }
          gb_spill_rows<1, VT, FLAGS>(p.gt, kq, 0ull, 0ull, sp, p.count_rows != 0, !vnull, T::from_bits(vb));
        }
      }
    }
    // keys (and bitmap words) of this CTA's next tile: in flight during phases 2-4
    if (nxt.valid) ts_load<RPT, PLAIN, PART, GENERIC && !PACK32, PACK32>(p, nxt, warp, lane, r);
    TSP_MARK(0);
    __syncthreads();
    TSP_MARK(1);
    // ---- phase 2: exclusive scan of the histogram (thread t owns entries [t * GPT, t * GPT + GPT))
    {
      uint32_t c[GPT], tsum = 0;
#pragma unroll
      for (int s = 0; s < GPT; s++) { c[s] = Hc[tid * GPT + s]; tsum += c[s] & 0xFFFFu; }
      uint32_t incl = tsum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
      if (lane == 31) wsum[warp] = incl;
      __syncthreads();
      uint32_t ws = lane < NWARPS ? wsum[lane] : 0u;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, ws, d); if (lane >= d) ws += o; }
      const uint32_t wprefix = __shfl_sync(0xFFFFFFFFu, ws, (warp + 31) & 31);
      uint32_t run = (warp ? wprefix : 0u) + incl - tsum;
#pragma unroll
      for (int s = 0; s < GPT; s++) { Hc[tid * GPT + s] = run | (c[s] & 0xFFFF0000u); run += c[s] & 0xFFFFu; }
      if (tid == NT - 1) Hc[NP] = run;
      // the other buffer was last read in phase 4 of the previous tile: clear it for the next one
#pragma unroll
      for (int s = 0; s < GPT; s++) H[(b ^ 1) * NPAD + tid + s * NT] = 0;
      __syncthreads();
    }
    TSP_MARK(2);
    // ---- phase 3: scatter the values into group order
    {
      uint32_t pos[RPT];
#pragma unroll
      for (int q = 0; q < RPT; q++) pos[q] = (sm_ld32(a_h + (pack[q] & 0xFFFFu) * 4u) & 0xFFFFu) + (pack[q] >> 16);
      if (tile_tma) { while (!mbar_try_wait(a_mbar, tma_phase)) {} tma_phase ^= 1u; }
      const uint32_t a_src = a_stage + (uint32_t)(warp * WROWS + 2 * lane) * 8u;
#pragma unroll
      for (int j = 0; j < RPT / 2; j++) {
        const ulonglong2 vv = sm_ld128(a_src + 64 * 8 * j);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int q = 2 * j + h;
          u64 v = h ? vv.y : vv.x;
          if (!PLAIN) { if ((zeromask >> q) & 1u) v = 0; }
          if (!((skipmask >> q) & 1u)) sm_st64(a_sorted + pos[q] * 8u, v);
        }
      }
    }
    TSP_MARK(3);
    __syncthreads();
    TSP_MARK(4);
    if (nxt.valid) issue_vals(nxt);   // stage is free: values of the next tile
    // ---- phase 4: every thread reduces the segments of the groups it owns
    bool heavy[GPT], medium[GPT];
    uint32_t hoff[GPT], hend[GPT];
#pragma unroll
    for (int s = 0; s < GPT; s++) {
      const int pidx = tid + s * NT;
      const uint32_t h0 = Hc[pidx];
      const uint32_t off = h0 & 0xFFFFu, end = Hc[pidx + 1] & 0xFFFFu;
      const uint32_t len = end - off;
      acc[s].rows += len + (h0 >> 16);
      acc[s].n += len;
      hoff[s] = off; hend[s] = end;
      if (ALL && len && acc[s].piv == 0) {     // pivot = first finite value of the group seen by this CTA
        for (uint32_t i = off; i < end; i++) {
          const double x = T::to_f64(T::from_bits(sorted[i]));
          if (is_finite_f64(x)) { acc[s].piv = (u64)__double_as_longlong(x) | 1ull; break; }
        }
      }
      heavy[s] = len > (uint32_t)p.ts_heavy;
      medium[s] = TEAM && !heavy[s] && len > (uint32_t)p.ts_mid;
      if (!heavy[s] && !medium[s]) {
        const double pv = __longlong_as_double((long long)acc[s].piv);
        double S1 = acc[s].S1, S2 = acc[s].S2;
        VT mn = acc[s].mn, mx = acc[s].mx;
        u64 isum = acc[s].isum;
        uint32_t i = off;
#pragma unroll 1
        for (; i + 4 <= end; i += 4) {
          const u64 x0 = sorted[i], x1 = sorted[i + 1], x2 = sorted[i + 2], x3 = sorted[i + 3];
          ts_add<VT, FLAGS>(S1, S2, mn, mx, isum, pv, x0);
          ts_add<VT, FLAGS>(S1, S2, mn, mx, isum, pv, x1);
          ts_add<VT, FLAGS>(S1, S2, mn, mx, isum, pv, x2);
          ts_add<VT, FLAGS>(S1, S2, mn, mx, isum, pv, x3);
        }
#pragma unroll 1
        for (; i < end; i++) ts_add<VT, FLAGS>(S1, S2, mn, mx, isum, pv, sorted[i]);
        acc[s].S1 = S1; acc[s].S2 = S2; acc[s].mn = mn; acc[s].mx = mx; acc[s].isum = isum;
      }
    }
    // segments of ts_mid .. ts_heavy rows: teams of 8 lanes, four groups of the warp at a time (a whole-warp reduce
    // would spend more on the 5 x 8 shuffles than on the rows; one lane alone would serialise ~100 rows)
#pragma unroll
    for (int s = 0; s < GPT; s++) {
      unsigned mm = TEAM ? __ballot_sync(0xFFFFFFFFu, medium[s]) : 0u;
      while (TEAM && mm) {
        const int team = lane >> 3, rr = lane & 7;
        const int src = __fns(mm, 0, team + 1);                     // owner lane of this team's group, -1 if none
        const bool on = src >= 0 && src < 32;
        const int srcl = on ? src : 0;
        const uint32_t o = __shfl_sync(0xFFFFFFFFu, hoff[s], srcl), e = __shfl_sync(0xFFFFFFFFu, hend[s], srcl);     // every lane shuffles
        const u64 pvb = __shfl_sync(0xFFFFFFFFu, acc[s].piv, srcl);
        const double pv = __longlong_as_double((long long)pvb);
        double S1 = 0.0, S2 = 0.0;
        VT mn = T::min_init(), mx = T::max_init();
        u64 isum = 0;
        uint32_t i = o + rr;
        const uint32_t ee = on ? e : 0u;
#pragma unroll 1
        for (; i + 8 < ee; i += 16) { const u64 x0 = sorted[i], x1 = sorted[i + 8]; ts_add<VT, FLAGS>(S1, S2, mn, mx, isum, pv, x0); ts_add<VT, FLAGS>(S1, S2, mn, mx, isum, pv, x1); }
        if (i < ee) ts_add<VT, FLAGS>(S1, S2, mn, mx, isum, pv, sorted[i]);
#pragma unroll
        for (int d = 4; d; d >>= 1) {
          if (ALL || !IS_INT) S1 += __shfl_xor_sync(0xFFFFFFFFu, S1, d);
          if (IS_INT) isum += __shfl_xor_sync(0xFFFFFFFFu, isum, d);
          if (ALL) {
            S2 += __shfl_xor_sync(0xFFFFFFFFu, S2, d);
            const VT omn = T::from_bits(__shfl_xor_sync(0xFFFFFFFFu, T::to_bits(mn), d)), omx = T::from_bits(__shfl_xor_sync(0xFFFFFFFFu, T::to_bits(mx), d));
            mn = omn < mn ? omn : mn;
            mx = omx > mx ? omx : mx;
          }
        }
        // hand the team results to the owner lanes: the k-th selected owner reads lane 8 k
        const int rank = __popc(mm & ((1u << lane) - 1u));
        const bool mine = ((mm >> lane) & 1u) && rank < 4;
        const int from = mine ? rank * 8 : lane;
        const double rS1 = __shfl_sync(0xFFFFFFFFu, S1, from), rS2 = __shfl_sync(0xFFFFFFFFu, S2, from);
        const u64 risum = __shfl_sync(0xFFFFFFFFu, isum, from);
        const VT rmn = T::from_bits(__shfl_sync(0xFFFFFFFFu, T::to_bits(mn), from)), rmx = T::from_bits(__shfl_sync(0xFFFFFFFFu, T::to_bits(mx), from));
        if (mine) {
          acc[s].S1 += rS1; acc[s].S2 += rS2; acc[s].isum += risum;
          if (ALL) { acc[s].mn = rmn < acc[s].mn ? rmn : acc[s].mn; acc[s].mx = rmx > acc[s].mx ? rmx : acc[s].mx; }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) mm &= mm - 1;
      }
    }
#pragma unroll
    for (int s = 0; s < GPT; s++) {
      unsigned hm = __ballot_sync(0xFFFFFFFFu, heavy[s]);
      while (hm) {
        const int src = __ffs(hm) - 1;
        hm &= hm - 1;
        const uint32_t o = __shfl_sync(0xFFFFFFFFu, hoff[s], src), e = __shfl_sync(0xFFFFFFFFu, hend[s], src);
        const u64 pvb = __shfl_sync(0xFFFFFFFFu, acc[s].piv, src);
        const double pv = __longlong_as_double((long long)pvb);
        double S1 = 0.0, S2 = 0.0;
        VT mn = T::min_init(), mx = T::max_init();
        u64 isum = 0;
        for (uint32_t i = o + lane; i < e; i += 32) ts_add<VT, FLAGS>(S1, S2, mn, mx, isum, pv, sorted[i]);
#pragma unroll
        for (int d = 16; d; d >>= 1) {
          if (ALL || !IS_INT) S1 += __shfl_xor_sync(0xFFFFFFFFu, S1, d);
          if (IS_INT) isum += __shfl_xor_sync(0xFFFFFFFFu, isum, d);
          if (ALL) {
            S2 += __shfl_xor_sync(0xFFFFFFFFu, S2, d);
            const VT omn = T::from_bits(__shfl_xor_sync(0xFFFFFFFFu, T::to_bits(mn), d)), omx = T::from_bits(__shfl_xor_sync(0xFFFFFFFFu, T::to_bits(mx), d));
            mn = omn < mn ? omn : mn;
            mx = omx > mx ? omx : mx;
          }
        }
        if (lane == src) {
          acc[s].S1 += S1; acc[s].S2 += S2; acc[s].isum += isum;
          if (ALL) { acc[s].mn = mn < acc[s].mn ? mn : acc[s].mn; acc[s].mx = mx > acc[s].mx ? mx : acc[s].mx; }
        }
      }
    }
    // no barrier here: the next tile's phase 1 only touches the other histogram buffer, and its scatter into
    // `sorted` comes after that tile's first barrier, which every thread reaches only after this phase
    TSP_MARK(5);
    if (PART && cur.last) {              // end of a partition: flush its groups, start over with empty tables
      flush();
      __syncthreads();
      reset_acc();
      for (int i = tid; i < S; i += NT) ktab_id[i] = 0;
      if (tid == 0) misc[0] = 0;
      __syncthreads();
    }
    cur = nxt;
  }
#ifdef TS_PROFILE
  if (lane == 0) for (int k = 0; k < 6; k++) atomicAdd(&ts_prof[k], tacc[k]);
#endif
  if (!PART) flush();
}

template <int NT, typename VT, int FLAGS, int GPT, bool TEAM>
cudaError_t ts_launch5(const GbParams& p, int ctas, size_t smem, cudaStream_t s) {
  const bool generic = p.ts_generic != 0;
  const bool plain = !p.fbits && !p.compat_nulls && (generic || !p.ks.c[0].nulls);
  auto k = gb_tsort_kernel<NT, VT, FLAGS, GPT, 0, true, TEAM>;
  if (p.part_keys) k = gb_tsort_kernel<NT, VT, FLAGS, GPT, 2, true, TEAM>;
  else if (generic) {
    bool pack32 = p.ks.nkeys <= 2;
    for (int i = 0; i < p.ks.nkeys; i++) pack32 = pack32 && !p.ks.c[i].nulls && (p.ks.c[i].dtype == PDRS_I32 || p.ks.c[i].dtype == PDRS_DICT_U32) && p.ks.c[i].offset == 0 && p.ks.c[i].bits == 32;
    if (pack32) k = plain ? gb_tsort_kernel<NT, VT, FLAGS, GPT, 4, true, TEAM> : gb_tsort_kernel<NT, VT, FLAGS, GPT, 4, false, TEAM>;
    else k = plain ? gb_tsort_kernel<NT, VT, FLAGS, GPT, 3, true, TEAM> : gb_tsort_kernel<NT, VT, FLAGS, GPT, 3, false, TEAM>;
  }
  else if (p.sh_dense) k = plain ? gb_tsort_kernel<NT, VT, FLAGS, GPT, 0, true, TEAM> : gb_tsort_kernel<NT, VT, FLAGS, GPT, 0, false, TEAM>;
  else k = plain ? gb_tsort_kernel<NT, VT, FLAGS, GPT, 1, true, TEAM> : gb_tsort_kernel<NT, VT, FLAGS, GPT, 1, false, TEAM>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<ctas, NT, smem, s>>>(p);
  return cudaGetLastError();
}
// team: the variant with the team-of-8 reduce for segments of ts_mid .. ts_heavy rows (few hundred groups or fewer, skew);
// with ~1000 uniform groups no segment is that long and the leaner variant is ~2% faster
template <typename VT, int FLAGS>
cudaError_t ts_launch2(const GbParams& p, int nt, int gpt, int ctas, size_t smem, cudaStream_t s) {
  const bool team = p.ts_team != 0;
  (void)nt;
  if (gpt == 2) return team ? ts_launch5<512, VT, FLAGS, 2, true>(p, ctas, smem, s) : ts_launch5<512, VT, FLAGS, 2, false>(p, ctas, smem, s);
  return team ? ts_launch5<512, VT, FLAGS, 4, true>(p, ctas, smem, s) : ts_launch5<512, VT, FLAGS, 4, false>(p, ctas, smem, s);
}

}  // namespace

// Geometry: cap + 1 group ids over nt * gpt owners (nt threads, gpt groups per thread).  Returns false when the
// group count does not fit.  nt_pref: 0 = auto, else 512 / 1024.  Hashed keys: the key table gets >= 4 slots per id
// when shared memory allows (load factor <= 1/4), else >= 2.
bool gb_tsort_geometry(long long cap, bool dense, int smem_budget, int nt_pref, int* nt, int* gpt, int* slots, size_t* smem) {
  if (cap + 1 > 2048) return false;
  (void)nt_pref;                                          // 1024 x 1 group and 2 x 256 threads with 4096-row tiles were measured: no faster
  const int t = 512;
  const int g = cap + 1 <= 1024 ? 1024 / t : 2048 / t;
  const size_t np = (size_t)t * g;
  const size_t tile = t == 256 ? 4096 : TS_T;
  const size_t budget = t == 256 ? (size_t)smem_budget / 2 - 1024 : (size_t)smem_budget;
  const size_t fixed = tile * 16 + 2 * (np + 32) * 4 + 32 * 4 + 16 + 16;
  long long S = 0;
  if (!dense) {
    S = 64;
    while (S < 4 * cap) S <<= 1;
    while (S > 2 * cap && fixed + (size_t)S * 12 + (size_t)(cap + 1) * 8 > budget) S >>= 1;
  }
  const size_t bytes = fixed + (size_t)S * 12 + (dense ? 0 : (size_t)(cap + 1) * 8);
  if (bytes > budget) return false;
  *nt = t; *gpt = g; *slots = (int)S; *smem = bytes;
  return true;
}
long long gb_tsort_tile_rows_for(int nt) { return nt == 256 ? 4096 : TS_T; }
long long gb_tsort_tile_rows() { return TS_T; }

cudaError_t gb_tsort_launch(const GbParams& p, int is_int, int flags, int nt, int gpt, int ctas, size_t smem, cudaStream_t s) {
  if (!is_int) return flags == GB_SUM ? ts_launch2<double, GB_SUM>(p, nt, gpt, ctas, smem, s) : ts_launch2<double, GB_ALL>(p, nt, gpt, ctas, smem, s);
  return flags == GB_SUM ? ts_launch2<long long, GB_SUM>(p, nt, gpt, ctas, smem, s) : ts_launch2<long long, GB_ALL>(p, nt, gpt, ctas, smem, s);
}

#ifdef TS_PROFILE
extern "C" int pdrs_debug_tsprof(unsigned long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, ts_prof, sizeof(ts_prof));
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(ts_prof, z, sizeof(z)); }
  return 0;
}
#endif
