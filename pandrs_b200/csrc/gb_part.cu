// High-cardinality groupby (thousands to tens of millions of groups, one 64-bit key column): hash-partition the
// (key, value) rows until every partition holds at most ~600 groups, then run the tile-sort kernel (gb_tsort.cu)
// partition by partition.  Replaces grouping.rs:62-104 + aggregation.rs:507-742 like the other groupby kernels.
//
// Why not a global hash table: with millions of groups every row costs 3-5 L2 atomics on random addresses
// (rows, n, S1, S2, min / max), and the L2 atomic units, not HBM, bound the kernel (measured 70 ms for 1e9 rows
// and 1e7 groups, profiles/bench_r01.json "groupby_all6_10m").  After partitioning, all rows of a group sit in one
// partition, a partition's groups fit the register accumulators of one CTA, and the only global-table traffic
// is one pre-aggregated batch per group.
//
//   gp_part_kernel<true>   level 1: columns (+ value NULL bitmap, row filter) -> 2^bits1 padded buckets of
//                          (key, value, flag byte); rows whose value is NULL only count (Count includes NULLs,
//                          aggregation.rs:743) and travel with flag bit 0 set
//   gp_part_kernel<false>  level 2: every level-1 bucket -> 2^bits2 sub-buckets; partition id = top bits1+bits2
//                          bits of the key hash
// Both levels are ONE pass (no histogram pass): bucket b owns a fixed padded range of the output, a tile's rows
// are ranked by a shared-memory histogram, one global reservation per bucket and tile, rows staged in shared
// memory in bucket order and written out as contiguous runs.  A bucket that fills up (hot keys) sends its further runs
// to a side area behind the buckets, which is aggregated as extra partitions; only when that overflows too does the
// caller fall back (tile-sort over the columns with a spill buffer, or the global-table path).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "groupby_kernels.cuh"
#include "gb_final.cuh"

bool gb_tsort_geometry(long long cap, bool dense, int smem_budget, int nt_pref, int* nt, int* gpt, int* slots, size_t* smem);
long long gb_tsort_tile_rows();
cudaError_t gb_tsort_launch(const GbParams& p, int is_int, int flags, int nt, int gpt, int ctas, size_t smem, cudaStream_t s);
int32_t gb_estimate_groups_i64(pdrs_ctx* c, const u64* keys, long long n, long long sample_rows, long long* est_out);   // groupby.cu

namespace {

constexpr int GP_NT = 512, GP_ITEMS = 8, GP_TILE = GP_NT * GP_ITEMS;

__device__ __forceinline__ uint32_t gp_hash32(u64 k) {     // = ts_hash32 (gb_tsort.cu)
  const uint32_t lo = (uint32_t)k ^ (uint32_t)(k >> 32), hi = (uint32_t)(k >> 32);
  constexpr uint32_t CL = 0x7F4A7C15u, CH = 0x9E3779B9u;
  return __umulhi(lo, CL) + lo * CH + hi * CL;
}

struct GpIn {
  // level 1: the columns
  const u64* keys; const u64* vals; const uint8_t* vnull; const uint8_t* fbits; const uint8_t* fnull;
  long long n;
  int compat_nulls;
  // level 2: padded buckets of level 1
  const u64* pkeys; const u64* pvals; const uint8_t* pflags; const u64* pcnt; long long pcap;
  // level 1, GENERIC: any key tuple that packs into one 64-bit word (dictionary ids, i32, bool and pairs of them, no NULLs)
  KeySpec ks;
  // level 1: hot keys (open-addressing set, all ones = empty) whose rows go straight to the side area
  const uint32_t* hot_tab; int hot_log_slots;
};

template <bool FROM_COLS, bool GENERIC = false>
__global__ void __launch_bounds__(GP_NT, 2) gp_part_kernel(GpIn in, int hash_shr, int local_bits, long long cap_out, u64* __restrict__ cursor,
                                                          u64* __restrict__ out_keys, u64* __restrict__ out_vals, uint8_t* __restrict__ out_flags,
                                                          u64* __restrict__ overflow, u64* __restrict__ side = nullptr, long long side_base = 0, long long side_cap = 0) {
  extern __shared__ __align__(16) unsigned char gsm[];
  ulonglong2* st_kv = reinterpret_cast<ulonglong2*>(gsm);      // [GP_TILE] staged (key, value)
  uint32_t* st_dst = reinterpret_cast<uint32_t*>(st_kv + GP_TILE);   // [GP_TILE] output position of the staged row (< 2^31), bit 31 = value is NULL
  uint32_t* H = st_dst + GP_TILE;                         // [256 + 32] bucket counts of the tile; entry nb = the rows of hot keys
  uint2* HD = reinterpret_cast<uint2*>(H + 288);          // [288] {offset of the bucket in the staging area, output position of its first row}
  uint32_t* hot = reinterpret_cast<uint32_t*>(HD + 288);  // [2^hot_log_slots] tags of the hot keys (level 1 only)
  const bool use_hot = FROM_COLS && in.hot_tab != nullptr && side != nullptr;
  const int hot_mask = (1 << in.hot_log_slots) - 1;
  if (use_hot) { for (int i = threadIdx.x; i <= hot_mask; i += GP_NT) hot[i] = in.hot_tab[i]; }
  __shared__ uint32_t sh_total;
  __shared__ uint32_t wsum[GP_NT / 32];
  const int nb = 1 << local_bits;
  const uint32_t lmask = (uint32_t)nb - 1u;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // level 2: blockIdx.y = level-1 bucket, its tiles are strided over gridDim.x
  long long t0 = (long long)blockIdx.x * GP_TILE, tstride = (long long)gridDim.x * GP_TILE;
  long long lim = in.n, base0 = 0;
  if (!FROM_COLS) { lim = min((long long)__ldg(in.pcnt + blockIdx.y), in.pcap); base0 = (long long)blockIdx.y * in.pcap; }
  u64 key[GP_ITEMS], val[GP_ITEMS];
  // Bitmap words of a tile: the 32 rows of item j of a warp share one 32-bit word, lane j holds it (value NULLs,
  // filter, filter NULLs).  Level 2: the raw flag byte of every row.  Nothing here is consumed at load time, so the
  // loads of a tile stay in flight while the previous tile is written out.
  uint32_t wv = 0, wf = 0xFFFFFFFFu, wfn = 0, fl[FROM_COLS ? 1 : GP_ITEMS];
  auto load_tile = [&](long long tb) {
#pragma unroll
    for (int j = 0; j < GP_ITEMS; j++) {
      const long long i = tb + (long long)j * GP_NT + tid;
      const bool inb = i < lim;
      if (FROM_COLS) {
        if (GENERIC) { u64 w[1] = {0ull}; if (inb) load_key_inline<1>(in.ks, i, w); key[j] = w[0]; }
        else key[j] = inb ? __ldcs(in.keys + i) : 0ull;
        val[j] = inb ? __ldcs(in.vals + i) : 0ull;
      } else {
        key[j] = inb ? __ldcs(in.pkeys + base0 + i) : 0ull;
        val[j] = inb ? __ldcs(in.pvals + base0 + i) : 0ull;
        fl[j] = (inb && in.pflags) ? (uint32_t)__ldcs(in.pflags + base0 + i) : 0u;
      }
    }
    if (FROM_COLS) {
      wv = 0; wf = 0xFFFFFFFFu; wfn = 0;
      const long long row0 = tb + (long long)(lane & (GP_ITEMS - 1)) * GP_NT + warp * 32;
      if (row0 < lim) {                     // bitmaps cover ceil(n / 64) * 8 bytes (pdrs_view_col)
        if (in.vnull) wv = __ldg(reinterpret_cast<const uint32_t*>(in.vnull) + (row0 >> 5));
        if (in.fbits) wf = __ldg(reinterpret_cast<const uint32_t*>(in.fbits) + (row0 >> 5));
        if (in.fnull) wfn = __ldg(reinterpret_cast<const uint32_t*>(in.fnull) + (row0 >> 5));
      }
    }
  };
  if (t0 < lim) load_tile(t0);
  for (; t0 < lim; t0 += tstride) {
    if (tid < 288) H[tid] = 0;
    __syncthreads();
    uint32_t br[GP_ITEMS];            // local bucket << 16 | rank inside the bucket; all ones = no row
    uint32_t vnullmask = 0;           // rows whose value is NULL: they only count (aggregation.rs:743)
#pragma unroll
    for (int j = 0; j < GP_ITEMS; j++) {
      br[j] = 0xFFFFFFFFu;
      bool live = t0 + (long long)j * GP_NT + tid < lim;
      if (FROM_COLS) {
        const uint32_t keep = __shfl_sync(0xFFFFFFFFu, wf & ~wfn, j);        // filter keeps Some(true) rows only (data_ops.rs:49-55)
        const uint32_t vn = __shfl_sync(0xFFFFFFFFu, wv, j);
        live = live && ((keep >> lane) & 1u);
        if ((vn >> lane) & 1u) { if (in.compat_nulls) val[j] = 0; else vnullmask |= 1u << j; }
      } else if (fl[j] & 1u) vnullmask |= 1u << j;
      if (!live) continue;
      uint32_t b = (gp_hash32(key[j]) >> hash_shr) & lmask;
      if (use_hot) {        // a hot key's rows bypass the hash buckets (they would overflow one): virtual bucket nb = the side area
        uint32_t hs = gb_hot_slot(key[j], in.hot_log_slots);
        const uint32_t tag = gb_hot_tag(key[j]);
        for (;;) { const uint32_t hk = hot[hs]; if (hk == tag) { b = (uint32_t)nb; break; } if (hk == 0u) break; hs = (hs + 1) & (uint32_t)hot_mask; }
      }
      br[j] = (b << 16) | atomicAdd(&H[b], 1u);
    }
    __syncthreads();
    {   // exclusive scan of the bucket counts (thread b owns bucket b) + one global reservation per bucket
      const uint32_t c = tid <= nb ? H[tid] : 0u;            // (H[nb] = rows of hot keys, 0 when there are none)
      uint32_t incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
      if (lane == 31) wsum[warp] = incl;
      uint32_t g = 0xFFFFFFFFu;
      if (c && tid == nb) {      // rows of hot keys: one reservation in the side area
        const u64 at2 = atomicAdd(&side[0], (u64)c);
        if (at2 + c <= (u64)side_cap) g = (uint32_t)((u64)side_base + at2); else atomicAdd(overflow, 1ull);
      } else if (c) {
        // output bucket: level 1 = the local bucket; level 2 = (level-1 bucket of this CTA's rows) * 2^bits + local bucket
        const u64 q = FROM_COLS ? (u64)tid : (((u64)blockIdx.y << local_bits) | (u64)tid);
        const u64 at = atomicAdd(&cursor[q], (u64)c);
        if (at + c > (u64)cap_out) {
          // The bucket is full (hot keys).  Level 1 parks the run in the side area behind the buckets - the caller
          // aggregates it as extra partitions; side[0] = its cursor, side[1 + q] = where the rows of bucket q end (the
          // one reservation that straddles the end of the range records it; later ones start beyond the range).
          bool placed = false;
          if (side) {
            if (at < (u64)cap_out) side[1 + q] = at;
            const u64 at2 = atomicAdd(&side[0], (u64)c);
            if (at2 + c <= (u64)side_cap) { g = (uint32_t)((u64)side_base + at2); placed = true; }
          }
          if (!placed) atomicAdd(overflow, 1ull);                      // the caller discards this partitioning
        } else g = (uint32_t)(q * (u64)cap_out + at);                 // < 2^31 (checked by the caller)
      }
      __syncthreads();
      uint32_t ws = lane < GP_NT / 32 ? wsum[lane] : 0u;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, ws, d); if (lane >= d) ws += o; }
      const uint32_t wprefix = __shfl_sync(0xFFFFFFFFu, ws, (warp + 31) & 31);
      const uint32_t excl = (warp ? wprefix : 0u) + incl - c;
      if (tid < 288) HD[tid] = make_uint2(excl, g);
      if (tid == GP_NT - 1) sh_total = excl + c;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < GP_ITEMS; j++) {
      if (br[j] == 0xFFFFFFFFu) continue;
      const uint2 hd = HD[br[j] >> 16];
      const uint32_t rank = br[j] & 0xFFFFu, pos = hd.x + rank;
      st_kv[pos] = make_ulonglong2(key[j], val[j]);
      st_dst[pos] = hd.y == 0xFFFFFFFFu ? 0xFFFFFFFFu : ((hd.y + rank) | (((vnullmask >> j) & 1u) << 31));
    }
    if (t0 + tstride < lim) load_tile(t0 + tstride);      // next tile: in flight during the write-out below
    __syncthreads();
    const uint32_t total = sh_total;
    for (uint32_t pos = tid; pos < total; pos += GP_NT) {   // consecutive staged rows of a bucket go to consecutive output rows
      const uint32_t dd = st_dst[pos];
      if (dd == 0xFFFFFFFFu) continue;
      const uint32_t d = dd & 0x7FFFFFFFu;
      const ulonglong2 kv = st_kv[pos];
      out_keys[d] = kv.x;
      out_vals[d] = kv.y;
      if (out_flags) out_flags[d] = (uint8_t)(dd >> 31);
    }
    __syncthreads();
  }
}

// final row counts of the partitions: hash buckets (clamped to where their rows end) followed by the chunks of the side area
__global__ void gp_counts_kernel(const u64* __restrict__ cursor, const u64* __restrict__ side, int nb, long long cap, int nside, u64* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nb) { u64 c = cursor[q]; if (c > (u64)cap) c = min((u64)cap, side[1 + q]); out[q] = c; }
  else if (q < nb + nside) {
    const long long total = (long long)side[0], j = q - nb;
    out[q] = (u64)max(0ll, min(cap, total - j * cap));
  }
}

// ---------------------------------------------------------------- hash aggregation of COMPLETE partitions, straight to the result
// After hash partitioning a group lives in exactly one partition.  When that partition received all of its rows (no run of it
// was parked in a side area: hot keys), the CTA that aggregates it holds the FINAL state of each of its groups and can write
// them to the result directly - one output reservation per partition - instead of adding one batch per group to the global
// table (~10 L2 atomics per group, a second scan of the table by the finalisation kernel).  Few rows per group (heavy-tailed
// key tuples: 3 - 20 rows per group) are also where the tile-sort kernel is weakest: every row of a partition's first tile is
// a new key.  So this kernel keeps one open-addressing table per partition in shared memory - keys + state planes, ~600
// groups in 2048 slots - and updates it with shared-memory atomics (~150 SM-cycles per 32 rows: affordable next to the three
// passes over HBM the partitioning costs).  Per tile of 2048 rows: (A) find / claim the slot of every row, (B) the first
// finite value of a group becomes its pivot (plain store, any winner is fine), (C) rows / n / S1 / S2 / min / max / isum atomics.
// A partition whose table fills up spills rows to the global table and then flushes its groups there as well (a group must
// not be reported twice); partitions flagged incomplete are left to the tile-sort kernel, which flushes to the table.
struct GpDirect {
  FinParams fin;            // result arrays with room for `cap` groups
  long long cap;
  u64* cursor;              // [0] groups written straight to the result, [1] partition tickets
  int dv;                   // index of the value column this pass aggregates
};
constexpr int GH_NT = 512, GH_ITEMS = 4, GH_TILE = GH_NT * GH_ITEMS;
static size_t gh_smem(int S, bool is_int) { return (size_t)(S + 1) * (is_int ? 64 : 56) + 64; }     // keys, pivot, S1, S2, min, max (+ isum) + rows, n

template <typename VT, int FLAGS>
__global__ void __launch_bounds__(GH_NT, 1) gp_hash_agg_kernel(const GbParams p, const GpDirect d, const u64* __restrict__ part_cnt, const uint8_t* __restrict__ part_complete, int log_s) {
  using T = ValTraits<VT>;
  constexpr bool IS_INT = T::is_int, ALL = FLAGS == GB_ALL;
  constexpr u64 EMPTY = ~0ull;
  extern __shared__ __align__(16) unsigned char gsm[];
  const int S = 1 << log_s, S1n = S + 1;                         // slot S belongs to the key whose packed word is all ones
  u64* keys = reinterpret_cast<u64*>(gsm);
  u64* piv = keys + S1n;
  double* aS1 = reinterpret_cast<double*>(piv + S1n);
  double* aS2 = aS1 + S1n;
  u64* amn = reinterpret_cast<u64*>(aS2 + S1n);
  u64* amx = amn + S1n;
  u64* aisum = amx + S1n;                                         // Int64 values only (the plane does not exist otherwise)
  uint32_t* arows = reinterpret_cast<uint32_t*>(IS_INT ? aisum + S1n : aisum);
  uint32_t* an = arows + S1n;
  __shared__ long long sh_q;
  __shared__ uint32_t sh_groups, sh_spilled, sh_out;
  __shared__ u64 sh_base;
  const int tid = threadIdx.x;
  const int limit = S - (S >> 3);                                // claim at most 7/8 of the slots
  for (;;) {
    if (tid == 0) sh_q = (long long)atomicAdd(&d.cursor[1], 1ull);
    __syncthreads();
    const long long q = sh_q;
    if (q >= p.part_n) break;
    const long long cnt = min((long long)__ldg(part_cnt + q), p.part_cap);
    if (cnt == 0) { __syncthreads(); continue; }
    for (int i = tid; i < S1n; i += GH_NT) { keys[i] = EMPTY; piv[i] = 0; aS1[i] = 0.0; aS2[i] = 0.0; amn[i] = 0; amx[i] = 0; if (IS_INT) aisum[i] = 0; arows[i] = 0; an[i] = 0; }
    if (tid == 0) { sh_groups = 0; sh_spilled = part_complete[q] ? 0u : 1u; sh_out = 0; }      // incomplete partition: its groups go through the global table
    __syncthreads();
    const long long base = q * p.part_cap;
    for (long long t0 = 0; t0 < cnt; t0 += GH_TILE) {
      u64 key[GH_ITEMS], vb[GH_ITEMS];
      int slot[GH_ITEMS];
      uint32_t live = 0, vnull = 0, spill = 0;
#pragma unroll
      for (int j = 0; j < GH_ITEMS; j++) {
        const long long i = t0 + (long long)j * GH_NT + tid;
        key[j] = 0; vb[j] = 0; slot[j] = -1;
        if (i < cnt) {
          live |= 1u << j;
          key[j] = __ldcs(p.part_keys + base + i);
          vb[j] = __ldcs(p.part_vals + base + i);
          if (p.part_flags && (__ldcs(p.part_flags + base + i) & 1)) vnull |= 1u << j;
        }
      }
      // ---- (A) slot of every row
#pragma unroll
      for (int j = 0; j < GH_ITEMS; j++) {
        if (!((live >> j) & 1u)) continue;
        if (key[j] == EMPTY) { slot[j] = S; if (keys[S] == EMPTY) { if (atomicCAS(&keys[S], EMPTY, 0ull) == EMPTY) atomicAdd(&sh_groups, 1u); } continue; }
        uint32_t s = (gp_hash32(key[j]) << p.part_bits) >> (32 - log_s);
        for (int probe = 0; probe < S; probe++) {
          u64 k = *reinterpret_cast<volatile u64*>(&keys[s]);
          if (k == EMPTY) {
            if (*reinterpret_cast<volatile uint32_t*>(&sh_groups) >= (uint32_t)limit) break;          // table full: the row spills
            k = atomicCAS(&keys[s], EMPTY, key[j]);
            if (k == EMPTY) { atomicAdd(&sh_groups, 1u); k = key[j]; }
          }
          if (k == key[j]) { slot[j] = (int)s; break; }
          s = (s + 1) & (uint32_t)(S - 1);
        }
        if (slot[j] < 0) spill |= 1u << j;
      }
      if (__any_sync(0xFFFFFFFFu, spill != 0)) {        // rare: more groups than the table holds -> global table, row by row
        if (spill) sh_spilled = 1;
#pragma unroll
        for (int j = 0; j < GH_ITEMS; j++)
          gb_spill_rows<1, VT, FLAGS>(p.gt, key[j], 0ull, 0ull, (spill >> j) & 1u, p.count_rows != 0, !((vnull >> j) & 1u), T::from_bits(vb[j]));
      }
      __syncthreads();
      // ---- (B) pivot = some finite value of the group (all racers belong to this tile; everybody reads it after the barrier)
      if (ALL) {
#pragma unroll
        for (int j = 0; j < GH_ITEMS; j++) {
          if (slot[j] < 0 || ((vnull >> j) & 1u)) continue;
          const double x = T::to_f64(T::from_bits(vb[j]));
          if (is_finite_f64(x) && *reinterpret_cast<volatile u64*>(&piv[slot[j]]) == 0) *reinterpret_cast<volatile u64*>(&piv[slot[j]]) = (u64)__double_as_longlong(x) | 1ull;
        }
        __syncthreads();
      }
      // ---- (C) accumulate
#pragma unroll
      for (int j = 0; j < GH_ITEMS; j++) {
        if (slot[j] < 0) continue;
        const int s = slot[j];
        atomicAdd(&arows[s], 1u);
        if ((vnull >> j) & 1u) continue;
        atomicAdd(&an[s], 1u);
        const VT v = T::from_bits(vb[j]);
        if (IS_INT) atomicAdd(&aisum[s], vb[j]);
        if (ALL) {
          const double dd = T::to_f64(v) - __longlong_as_double((long long)piv[s]);
          atomicAdd(&aS1[s], dd);
          atomicAdd(&aS2[s], dd * dd);
          if (T::orderable(v)) {
            const u64 o = T::ord(v);
            if (~o > *reinterpret_cast<volatile u64*>(&amn[s])) atomicMax(&amn[s], ~o);
            if (o > *reinterpret_cast<volatile u64*>(&amx[s])) atomicMax(&amx[s], o);
          }
        } else if (!IS_INT) atomicAdd(&aS1[s], T::to_f64(v));
      }
      __syncthreads();
    }
    // ---- the partition is done: its groups are final
    const uint32_t ng = sh_groups;
    if (tid == 0) {
      u64 b = ~0ull;
      if (!sh_spilled) { b = atomicAdd(&d.cursor[0], (u64)ng); if (b + ng > (u64)d.cap) { atomicAdd(&d.cursor[0], (u64)0 - (u64)ng); b = ~0ull; } }
      sh_base = b;
    }
    __syncthreads();
    const u64 obase = sh_base;
    for (int s0 = 0; s0 < S1n; s0 += GH_NT) {            // uniform trip count: the table insertion below is warp-synchronous
      const int s = s0 + tid;
      const bool have = s < S1n && keys[s] != EMPTY;
      GState st;
      st.n = 0; st.pivotx = 0; st.S1 = 0; st.S2 = 0; st.mnc = 0; st.mxo = 0; st.isum = 0; st.pad = 0;
      u64 rows = 0, kw = 0;
      if (have) {
        rows = arows[s]; kw = s == S ? EMPTY : keys[s];
        st.n = an[s]; st.S1 = aS1[s]; st.S2 = aS2[s]; st.mnc = amn[s]; st.mxo = amx[s]; st.isum = IS_INT ? aisum[s] : 0ull;
        st.pivotx = piv[s] ? (piv[s] ^ GB_PIV_X) : 0ull;
      }
      if (obase != ~0ull) {
        if (!have) continue;
        const long long o = (long long)(obase + atomicAdd(&sh_out, 1u));
        const u64 w[PDRS_MAX_WORDS] = {kw, 0ull, 0ull};
        fin_write_group(d.fin, o, w, false, rows, [&](int v) -> const GState* { return v == d.dv ? &st : nullptr; });
      } else {                                             // spilled rows / result capacity: through the global table like everybody else
        u64 w[1] = {kw};
        const long long gs = g_find_or_insert<1>(p.gt, w, have);
        if (!have || gs < 0) continue;
        if (p.count_rows) atomicAdd(&p.gt.hdr[gs].rowsw, rows);
        g_update_batch<FLAGS, IS_INT>(p.gt, gs, 0ull, st.n, fin_pivot(st.pivotx), st.pivotx != 0, st.S1, st.S2, st.isum, st.mnc, st.mxo);
      }
    }
    __syncthreads();
  }
}

template <typename VT, int FLAGS>
static cudaError_t gh_launch(const GbParams& p, const GpDirect& d, const u64* cnt, const uint8_t* complete, int log_s, int ctas, cudaStream_t s) {
  const size_t smem = gh_smem(1 << log_s, ValTraits<VT>::is_int);
  auto k = gp_hash_agg_kernel<VT, FLAGS>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<ctas, GH_NT, smem, s>>>(p, d, cnt, complete, log_s);
  return cudaGetLastError();
}

// Cardinality from the first level-1 bucket, exact for any key distribution: the bucket holds ALL rows of 1 / nb1 of the hash
// space; the rows whose NEXT `sub_bits` hash bits are zero are all rows of 1 / (nb1 * 2^sub_bits) of it.  Their distinct keys are
// counted exactly in a scratch table (the whole bucket is scanned - sequential, cheap - but only the slice is inserted).
__global__ void gp_slice_distinct_kernel(const u64* __restrict__ keys, long long n, int used_bits, int sub_bits, GTable t) {
  const int lane = threadIdx.x & 31;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i - lane < n; i += (long long)gridDim.x * blockDim.x) {
    u64 w[1] = {0ull};
    bool in = false;
    if (i < n) {
      w[0] = __ldcs(keys + i);
      in = sub_bits == 0 || ((gp_hash32(w[0]) << used_bits) >> (32 - sub_bits)) == 0u;
    }
    if (__any_sync(0xFFFFFFFFu, in)) g_find_or_insert<1>(t, w, in);
  }
}

// which partitions are complete: hash partitions none of whose rows (nor of its level-1 bucket's rows) went to a side area.
// out_hash[q] = rows for the hash-aggregation kernel (complete partitions: groups straight to the result; incomplete ones: flushed
// to the global table), out_ts[q] = rows for the tile-sort kernel (the chunks of the side area: hot keys); cnt = the clamped
// counts (gp_counts_kernel) or the raw cursors when there is no side area.
__global__ void gp_split_counts_kernel(const u64* __restrict__ cnt, const u64* __restrict__ cur1, long long cap1, int bits2, const u64* __restrict__ cur2, long long cap2,
                                       int nparts, int nall, int use_hash, u64* __restrict__ out_hash, u64* __restrict__ out_ts, uint8_t* __restrict__ out_complete) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nall) return;
  const u64 c = cnt[q];
  const bool hashed = use_hash && q < nparts;            // every hash partition goes to the hash kernel; the side-area chunks to the tile-sort kernel
  bool complete = hashed;
  if (complete && cur2) complete = cur2[q] <= (u64)cap2 && cur1[q >> bits2] <= (u64)cap1;
  else if (complete) complete = cur1[q] <= (u64)cap1;
  out_hash[q] = hashed ? c : 0ull;
  out_ts[q] = hashed ? 0ull : c;
  out_complete[q] = complete ? 1 : 0;
}

long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

}  // namespace

// One aggregation pass through the partitioned path.  Returns PDRS_ERR_UNSUPPORTED when it does not apply or a
// bucket overflowed (the caller then takes the global-table path; *dirty = the table was touched: never, today).
// est_refined (may be NULL): the cardinality estimate of the caller comes from 2^18 rows spread over the whole input and is
// far too low for heavy-tailed keys (Zipf tuples: most sampled keys are hot ones).  After level 1 the first bucket
// holds ALL rows of 1 / 2^bits1 of the key space, so a sample of it sees 2^bits1 times more of that slice: when
// groups(bucket 0) x 2^bits1 exceeds the estimate by more than 1.5x the pass stops, *est_refined is set, and the
// caller starts over with the better estimate (right-sized table, right number of partitions).
struct GbDirectOut { const FinParams* fin; long long cap; u64* cursor; int dv; };   // (declared in groupby.cu as well)
int32_t gb_part_pass(pdrs_ctx* c, const GbParams& gp, int is_int, int flags, long long est_groups, float* kernel_ms, bool* dirty, bool* skewed,
                     long long* est_refined, const GbDirectOut* direct) {
  *dirty = false;
  *skewed = false;
  if (est_refined) *est_refined = 0;
  // timing = 2: CUDA-event time of every phase on stderr
  struct Marks {
    pdrs_ctx* c; std::vector<std::pair<const char*, cudaEvent_t>> ev;
    void operator()(const char* name) { if (c->opt_timing < 2) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c->stream); ev.push_back({name, e}); }
    ~Marks() {
      if (ev.empty()) return;
      cudaStreamSynchronize(c->stream);
      for (size_t i = 1; i < ev.size(); i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second); fprintf(stderr, "[pdrs groupby part] %-28s %8.3f ms\n", ev[i].first, ms); }
      for (auto& e : ev) cudaEventDestroy(e.second);
    }
  } mark{c, {}};
  mark("start");
  const long long n = gp.n;
  const long long T = gb_tsort_tile_rows();
  const bool generic = !(gp.ks.nkeys == 1 && gp.ks.c[0].dtype == PDRS_I64);
  if (gp.ks.nwords != 1 || !gp.val || n < (1 << 20)) return PDRS_ERR_UNSUPPORTED;
  for (int k = 0; k < gp.ks.nkeys; k++)
    if (gp.ks.c[k].nulls || (gp.ks.c[k].dtype == PDRS_DICT_U32 && gp.ks.c[k].null_alias >= 0)) return PDRS_ERR_UNSUPPORTED;   // NULL keys: other paths
  // partitions of <= ~600 expected groups (the tile-sort kernel holds 1023 ids with 512 threads x 2 groups).  Few
  // partitions are fine: the aggregation kernel cuts every partition into chunks of tiles, one work item each.
  // Few rows per group (and room to write groups straight to the result): the shared-memory hash kernel aggregates the
  // partitions - its table holds up to 3584 groups (f64 values; 1792 for Int64 values) and a partition may be smaller than a tile.
  // Measured (profiles/c4_phases_r02.txt, 5e8 rows): uniform keys, 20 rows per group: hash kernel 14.2 ms vs tile sort 16.4 ms; Zipf
  // tuples (moderately hot keys inside the partitions): its shared-memory atomics serialise, 45 ms vs 18 ms - so with hot keys
  // around it is used only where the tile-sort kernel cannot go (more than ~700 groups per partition: near-unique tuples).
  const bool use_hash = direct && c->opt_part_hash != 2 &&
                        (c->opt_part_hash == 1 || (n / std::max<long long>(est_groups, 1) < 64 && (!gp.hot_tab || (est_groups >> 16) > 700)));
  const long long hash_groups_max = is_int ? 1400 : 2800;
  int bits = 1;
  while (bits < 16 && (est_groups >> bits) > 600) bits++;
  if ((est_groups >> bits) > (use_hash ? hash_groups_max : 700) || (n >> bits) < (use_hash ? 1024 : T / 2)) return PDRS_ERR_UNSUPPORTED;
  const int bits1 = bits <= 8 ? bits : (bits + 1) / 2, bits2 = bits - bits1;
  const long long nb1 = 1ll << bits1, nparts = 1ll << bits;
  // a bucket holds whole groups: with m groups per bucket its size varies by ~1/sqrt(m) -> 6 sigma of slack
  auto padded = [&](long long buckets, long long extra) {
    const double m = std::max(1.0, (double)est_groups / (double)buckets);
    const double mean = (double)n / (double)buckets;
    return round_up((long long)(mean * (1.0 + std::max(1.0 / 32.0, 6.0 / std::sqrt(m)))) + extra, T);
  };
  const long long cap1 = padded(nb1, 65536);
  const long long cap2 = bits2 ? padded(nparts, 8192) : cap1;
  // single level: a side area of up to n / 4 rows behind the buckets takes the runs of buckets that fill up (hot keys of
  // a skewed distribution); it is aggregated as `nside` more partitions of cap1 rows
  // Two levels: level 2 gets a side area of its own behind its partitions; the rows level 1 parked are copied to its
  // front (they belong to many level-1 buckets, so level 2 cannot re-partition them; they are dominated by the hot keys
  // anyway), and the whole area is aggregated as extra partitions of cap2 rows.
  long long nside = 0;
  if (c->opt_part_side != 0) {
    long long side_rows = n / 4;
    if (gp.hot_tab) side_rows = std::max<long long>(side_rows, (long long)((double)n * std::min(0.95, (double)gp.hot_frac * 1.15 + 0.05)) + n / 16);
    nside = std::max<long long>(1, (side_rows + cap1 - 1) / cap1);
    while (nside > 0 && (unsigned long long)(nb1 + nside) * cap1 >= (1ull << 31)) nside--;
  }
  if ((unsigned long long)(nb1 + nside) * cap1 >= (1ull << 31) || (unsigned long long)nparts * cap2 >= (1ull << 31)) return PDRS_ERR_UNSUPPORTED;   // 31-bit output positions
  const size_t need = (size_t)(nb1 + nside) * cap1 * 17 + (bits2 ? ((size_t)nparts * cap2 + (nside ? (size_t)n / 2 : 0)) * 17 : 0);
  size_t free_b = 0;
  PDRS_TRY(pdrs_mem_available(c, &free_b));
  if (need + (2ull << 30) > free_b) return PDRS_ERR_UNSUPPORTED;

  int ts_nt = 0, ts_gpt = 0, ts_slots = 0;
  size_t ts_smem = 0;
  const long long ts_cap = 1023;
  if (!gb_tsort_geometry(ts_cap, false, c->smem_optin, 512, &ts_nt, &ts_gpt, &ts_slots, &ts_smem)) return PDRS_ERR_UNSUPPORTED;

  DevBuf cnt, k1, v1, f1, k2, v2, f2, side, pcnt, side2, pcnt2;
  const bool has_flags = gp.vnull != nullptr && !gp.compat_nulls;
  PDRS_TRY(cnt.alloc(c, (size_t)(nb1 + nparts + 8) * 8, true));     // [nb1] level-1 cursors, [nparts] level-2 cursors, [1] overflow
  u64* cur1 = cnt.as<u64>();
  u64* cur2 = cur1 + nb1;
  u64* ovf = cur2 + nparts;
  PDRS_TRY(k1.alloc(c, (size_t)(nb1 + nside) * cap1 * 8));
  PDRS_TRY(v1.alloc(c, (size_t)(nb1 + nside) * cap1 * 8));
  if (has_flags) PDRS_TRY(f1.alloc(c, (size_t)(nb1 + nside) * cap1 + 64));
  if (nside) {
    PDRS_TRY(side.alloc(c, (size_t)(nb1 + 2) * 8));
    PDRS_CUDA(c, cudaMemsetAsync(side.p, 0xFF, (size_t)(nb1 + 2) * 8, c->stream));     // bucket ends: none recorded
    PDRS_CUDA(c, cudaMemsetAsync(side.p, 0, 8, c->stream));                             // side cursor
    PDRS_TRY(pcnt.alloc(c, (size_t)(nb1 + nside + 2) * 8, true));
  }
  u64* sidep = nside ? side.as<u64>() : nullptr;
  const long long side_base = nb1 * cap1, side_cap = nside * cap1;
  const size_t smem = (size_t)GP_TILE * 20 + 288 * 4 + 288 * 8 + ((size_t)4 << GB_HOT_LOG_SLOTS);
  // the attribute is per device (several contexts / GPUs may live in one process): set it on every call
  PDRS_CUDA(c, cudaFuncSetAttribute(gp_part_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PDRS_CUDA(c, cudaFuncSetAttribute(gp_part_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PDRS_CUDA(c, cudaFuncSetAttribute(gp_part_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
  GpIn in{};
  in.keys = reinterpret_cast<const u64*>(gp.ks.c[0].data); in.vals = reinterpret_cast<const u64*>(gp.val); in.vnull = gp.vnull;
  in.fbits = gp.fbits; in.fnull = gp.fnull; in.n = n; in.compat_nulls = gp.compat_nulls;
  const int ctas1 = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 2, (n + GP_TILE - 1) / GP_TILE));
  in.ks = gp.ks;
  const bool side_single = nside && gp.hot_tab != nullptr;      // hot keys known: their rows go to the side area, which one tile-sort pass aggregates as a single input
  if (side_single) { in.hot_tab = gp.hot_tab; in.hot_log_slots = gp.hot_log_slots; }
  if (generic) gp_part_kernel<true, true><<<ctas1, GP_NT, smem, c->stream>>>(in, 32 - bits1, bits1, cap1, cur1, k1.as<u64>(), v1.as<u64>(), has_flags ? f1.as<uint8_t>() : nullptr, ovf, sidep, side_base, side_cap);
  else gp_part_kernel<true><<<ctas1, GP_NT, smem, c->stream>>>(in, 32 - bits1, bits1, cap1, cur1, k1.as<u64>(), v1.as<u64>(), has_flags ? f1.as<uint8_t>() : nullptr, ovf, sidep, side_base, side_cap);
  if (nside) { gp_counts_kernel<<<(int)((nb1 + nside + 255) / 256), 256, 0, c->stream>>>(cur1, sidep, (int)nb1, cap1, (int)nside, pcnt.as<u64>()); c->stats.kernel_launches++; }
  c->stats.kernel_launches++;
  mark("level-1 partition");
  long long side_rows1 = 0;         // rows level 1 parked in its side area
  if ((est_refined && nb1 >= 4) || (bits2 && nside)) {
    PDRS_CUDA(c, cudaGetLastError());
    PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 8, ovf, 8, cudaMemcpyDeviceToHost, c->stream));
    PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 9, cur1, 8, cudaMemcpyDeviceToHost, c->stream));
    if (nside) PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 10, sidep, 8, cudaMemcpyDeviceToHost, c->stream));
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->pinned_scalars[8] != 0) { *skewed = true; return PDRS_ERR_UNSUPPORTED; }
    if (nside) side_rows1 = std::min<long long>((long long)c->pinned_scalars[10], side_cap);
  }
  if (est_refined && nb1 >= 4) {
    // Not bucket 0: the packed tuple of all-minimum parts is the word 0, hash(0) = 0, and under skew that is the HOTTEST key - its
    // bucket overflows into the side area and shows only part of its rows.  The bucket with the fewest rows has no hot key
    // and holds all of its rows.
    std::vector<u64> hcur((size_t)nb1);
    PDRS_CUDA(c, cudaMemcpyAsync(hcur.data(), cur1, (size_t)nb1 * 8, cudaMemcpyDeviceToHost, c->stream));
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    long long qmin = 0;
    for (long long q = 1; q < nb1; q++) if (hcur[q] < hcur[qmin]) qmin = q;
    const long long cnt0 = std::min<long long>((long long)hcur[qmin], cap1);
    const u64* kq = k1.as<u64>() + qmin * cap1;
    long long est0 = 0;
    // (the uniform-model inversion of a row sample is far too low for Zipf tuples: count a hash slice of the bucket exactly)
    int sub_bits = 0;
    while (sub_bits < 10 && (cnt0 >> sub_bits) > (1ll << 21)) sub_bits++;
    {
      DevBuf sh, sc;
      long long slots = 1024;
      while (slots < 4 * ((cnt0 >> sub_bits) + 1024)) slots <<= 1;
      PDRS_TRY(sh.alloc(c, (size_t)(slots + 1) * sizeof(GHdr), true));
      PDRS_TRY(sc.alloc(c, CNT_N * 8, true));
      GTable st{};
      st.hdr = sh.as<GHdr>(); st.mask = (u64)slots - 1; st.slots = slots; st.counters = sc.as<u64>();
      int lg = 0;
      while ((1ll << lg) < slots) lg++;
      st.shift = 64 - lg;
      if (cnt0 > 0) gp_slice_distinct_kernel<<<pdrs_grid_for(c, cnt0, 256), 256, 0, c->stream>>>(kq, cnt0, bits1, sub_bits, st);
      c->stats.kernel_launches++;
      PDRS_CUDA(c, cudaGetLastError());
      PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 9, st.counters + CNT_NGROUPS, 8, cudaMemcpyDeviceToHost, c->stream));
      PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
      est0 = (long long)c->pinned_scalars[9] << sub_bits;
    }
    mark("cardinality from one bucket");
    const long long refined = std::min<long long>(est0 * nb1 + est0 * nb1 / 32, n);
    if (refined > est_groups + est_groups / 2) { *est_refined = refined; return PDRS_ERR_UNSUPPORTED; }
  }
  const u64 *pk = k1.as<u64>(), *pv = v1.as<u64>(), *pc = nside ? pcnt.as<u64>() : cur1;
  const uint8_t* pf = has_flags ? f1.as<uint8_t>() : nullptr;
  long long pcap = cap1;
  long long nside2 = 0;
  if (bits2) {
    if (nside) {
      nside2 = (side_rows1 + n / 4 + cap2 - 1) / cap2;
      while (nside2 > 0 && (unsigned long long)(nparts + nside2) * cap2 >= (1ull << 31)) nside2--;
      if (nside2 * cap2 < side_rows1) { *skewed = true; return PDRS_ERR_UNSUPPORTED; }
    }
    const long long side_base2 = nparts * cap2, side_cap2 = nside2 * cap2;
    PDRS_TRY(k2.alloc(c, (size_t)(nparts + nside2) * cap2 * 8));
    PDRS_TRY(v2.alloc(c, (size_t)(nparts + nside2) * cap2 * 8));
    if (has_flags) PDRS_TRY(f2.alloc(c, (size_t)(nparts + nside2) * cap2 + 64));
    u64* side2p = nullptr;
    if (nside2) {
      PDRS_TRY(side2.alloc(c, (size_t)(nparts + 2) * 8));
      PDRS_TRY(pcnt2.alloc(c, (size_t)(nparts + nside2 + 2) * 8, true));
      side2p = side2.as<u64>();
      PDRS_CUDA(c, cudaMemsetAsync(side2.p, 0xFF, (size_t)(nparts + 2) * 8, c->stream));       // partition ends: none recorded
      c->pinned_scalars[11] = side_rows1;                                                        // level 2 appends behind level 1's rows
      PDRS_CUDA(c, cudaMemcpyAsync(side2.p, c->pinned_scalars + 11, 8, cudaMemcpyHostToDevice, c->stream));
      if (side_rows1 > 0) {
        PDRS_CUDA(c, cudaMemcpyAsync(k2.as<u64>() + side_base2, k1.as<u64>() + side_base, (size_t)side_rows1 * 8, cudaMemcpyDeviceToDevice, c->stream));
        PDRS_CUDA(c, cudaMemcpyAsync(v2.as<u64>() + side_base2, v1.as<u64>() + side_base, (size_t)side_rows1 * 8, cudaMemcpyDeviceToDevice, c->stream));
        if (has_flags) PDRS_CUDA(c, cudaMemcpyAsync(f2.as<uint8_t>() + side_base2, f1.as<uint8_t>() + side_base, (size_t)side_rows1, cudaMemcpyDeviceToDevice, c->stream));
      }
    }
    GpIn in2{};
    in2.pkeys = k1.as<u64>(); in2.pvals = v1.as<u64>(); in2.pflags = pf; in2.pcnt = pc; in2.pcap = cap1;      // pc: level-1 counts (clamped when a side area exists)
    const int gx = (int)std::max<long long>(1, std::min<long long>((cap1 + GP_TILE - 1) / GP_TILE, std::max<long long>(1, (long long)c->sm_count * 2 * 2 / nb1)));
    gp_part_kernel<false><<<dim3(gx, (unsigned)nb1), GP_NT, smem, c->stream>>>(in2, 32 - bits, bits2, cap2, cur2, k2.as<u64>(), v2.as<u64>(), has_flags ? f2.as<uint8_t>() : nullptr, ovf,
                                                                                side2p, side_base2, side_cap2);
    c->stats.kernel_launches++;
    if (nside2) { gp_counts_kernel<<<(int)((nparts + nside2 + 255) / 256), 256, 0, c->stream>>>(cur2, side2p, (int)nparts, cap2, (int)nside2, pcnt2.as<u64>()); c->stats.kernel_launches++; }
    pk = k2.as<u64>(); pv = v2.as<u64>(); pc = nside2 ? pcnt2.as<u64>() : cur2; pcap = cap2;
    pf = has_flags ? f2.as<uint8_t>() : nullptr;
  }
  mark("level-2 partition");
  PDRS_CUDA(c, cudaGetLastError());
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 8, ovf, 8, cudaMemcpyDeviceToHost, c->stream));
  // rows in the final side area (hot keys, runs of full buckets) and where it lies
  const u64* side_cursor = bits2 ? (nside2 ? side2.as<u64>() : nullptr) : sidep;
  const long long side_at = bits2 ? nparts * cap2 : side_base, side_room = bits2 ? nside2 * cap2 : side_cap;
  c->pinned_scalars[13] = 0;
  if (side_cursor) PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 13, side_cursor, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  const long long side_rows_final = std::min<long long>((long long)c->pinned_scalars[13], side_room);
  if (c->pinned_scalars[8] != 0) { *skewed = true; return PDRS_ERR_UNSUPPORTED; }      // skewed keys: a bucket overflowed its padded range
  if (bits2) { k1.release(); v1.release(); f1.release(); }                        // only the last level is read below
  GbParams tp = gp;
  tp.ts_generic = 0;               // the partitions hold packed key words
  tp.part_keys = pk; tp.part_vals = pv; tp.part_flags = pf; tp.part_cnt = pc; tp.part_cap = pcap; tp.part_bits = bits;
  const long long nparts_all = nparts + (bits2 ? nside2 : nside);
  tp.part_n = (int)nparts_all;
  {   // work items: ~16 chunks per CTA, each chunk a run of whole tiles of one partition
    const long long tiles_cap = pcap / T, want = 16ll * c->sm_count;
    long long cpp = std::max<long long>(1, std::min<long long>(tiles_cap, (want + nparts_all - 1) / nparts_all));
    tp.part_chunk_tiles = (int)((tiles_cap + cpp - 1) / cpp);
    tp.part_cpp = (int)((tiles_cap + tp.part_chunk_tiles - 1) / tp.part_chunk_tiles);
  }
  tp.ts_heavy = c->opt_tsort_heavy > 0 ? (int)c->opt_tsort_heavy : 256;
  tp.ts_mid = c->opt_tsort_mid > 0 ? (int)c->opt_tsort_mid : 48;
  tp.ts_team = (est_groups >> bits) < 64;          // a partition with few groups has long segments
  tp.sh_cap = (int)ts_cap; tp.sh_slots = ts_slots; tp.sh_dense = 0; tp.sh_dense_base = 0;
  int lg = 0;
  while ((1 << lg) < ts_slots) lg++;
  tp.sh_log_slots = lg;
  // Complete partitions (all rows of their groups are in them): aggregated by the shared-memory hash kernel, groups written
  // straight to the result.  Chosen when a group has few rows (the tile-sort kernel pays for every new key of a partition's
  // first tile); the tile-sort kernel then only sees the incomplete partitions and the side-area chunks.
  DevBuf split;
  if (direct) {
    PDRS_TRY(split.alloc(c, (size_t)(2 * nparts_all + 2) * 8 + (size_t)nparts_all + 64));
    u64* cnt_hash = split.as<u64>();
    u64* cnt_ts = cnt_hash + nparts_all;
    uint8_t* complete = reinterpret_cast<uint8_t*>(cnt_ts + nparts_all + 2);
    gp_split_counts_kernel<<<(int)((nparts_all + 255) / 256), 256, 0, c->stream>>>(pc, cur1, cap1, bits2, bits2 ? cur2 : nullptr, cap2, (int)nparts, (int)nparts_all, use_hash ? 1 : 0, cnt_hash, cnt_ts, complete);
    c->stats.kernel_launches++;
    if (use_hash) {
      GpDirect gd{};
      gd.fin = *direct->fin; gd.cap = direct->cap; gd.cursor = direct->cursor; gd.dv = direct->dv;
      int log_s = 8;
      while (log_s < (is_int ? 11 : 12) && (1ll << log_s) < 3 * std::max<long long>(1, est_groups >> bits)) log_s++;
      const int hctas = (int)std::min<long long>(c->sm_count, nparts);
      GbParams hp = tp;
      cudaError_t e;
      if (!is_int) e = flags == GB_SUM ? gh_launch<double, GB_SUM>(hp, gd, cnt_hash, complete, log_s, hctas, c->stream) : gh_launch<double, GB_ALL>(hp, gd, cnt_hash, complete, log_s, hctas, c->stream);
      else e = flags == GB_SUM ? gh_launch<long long, GB_SUM>(hp, gd, cnt_hash, complete, log_s, hctas, c->stream) : gh_launch<long long, GB_ALL>(hp, gd, cnt_hash, complete, log_s, hctas, c->stream);
      PDRS_CUDA(c, e);
      c->stats.kernel_launches++;
      tp.part_cnt = cnt_ts;
      mark("hash aggregation (direct)");
    }
  }
  if (!use_hash && side_single) {      // tile-sort over the hash partitions only (their chunk counts), then the side area below
    tp.part_n = (int)nparts;
    PDRS_CUDA(c, gb_tsort_launch(tp, is_int, flags, ts_nt, ts_gpt, (int)std::min<long long>(c->sm_count, nparts * tp.part_cpp), ts_smem, c->stream));
    c->stats.kernel_launches++;
    mark("tile-sort aggregation (partitions)");
  }
  if (use_hash || side_single) {
    // the side area as ONE input of the tile-sort kernel: hot keys + the runs of buckets that filled up - few distinct keys, many
    // rows each (its sweet spot); every CTA takes a run of tiles and adds one batch per group to the global table at its end
    if (side_rows_final > 0) {
      GbParams sp = tp;
      sp.part_keys = pk + side_at; sp.part_vals = pv + side_at; sp.part_flags = pf ? pf + side_at : nullptr;
      sp.part_cnt = side_cursor; sp.part_cap = side_room; sp.part_n = 1;
      const long long tiles = (side_rows_final + T - 1) / T;
      const int sctas = (int)std::min<long long>(c->sm_count, tiles);
      sp.part_chunk_tiles = (int)((tiles + sctas - 1) / sctas);
      sp.part_cpp = (int)((tiles + sp.part_chunk_tiles - 1) / sp.part_chunk_tiles);
      sp.ts_team = c->opt_tsort_team == 1 ? 1 : 0;
      int s_nt = ts_nt, s_gpt = ts_gpt, s_slots = ts_slots;
      size_t s_smem = ts_smem;
      if (gb_tsort_geometry(2047, false, c->smem_optin, 512, &s_nt, &s_gpt, &s_slots, &s_smem)) {      // room for every hot key (<= 1800) + some more
        sp.sh_cap = 2047; sp.sh_slots = s_slots;
        int l2 = 0;
        while ((1 << l2) < s_slots) l2++;
        sp.sh_log_slots = l2;
      } else { s_nt = ts_nt; s_gpt = ts_gpt; s_smem = ts_smem; }
      PDRS_CUDA(c, gb_tsort_launch(sp, is_int, flags, s_nt, s_gpt, sctas, s_smem, c->stream));
      c->stats.kernel_launches++;
    }
  } else {
    PDRS_CUDA(c, gb_tsort_launch(tp, is_int, flags, ts_nt, ts_gpt, (int)std::min<long long>(c->sm_count, nparts_all * tp.part_cpp), ts_smem, c->stream));
    c->stats.kernel_launches++;
  }
  mark("tile-sort aggregation");
  if (c->opt_timing) {
    PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
    PDRS_CUDA(c, cudaEventSynchronize(c->ev_b));
    PDRS_CUDA(c, cudaEventElapsedTime(kernel_ms, c->ev_a, c->ev_b));
  }
  return PDRS_OK;
}
