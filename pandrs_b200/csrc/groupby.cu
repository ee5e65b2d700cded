// Host orchestration of the groupby-aggregate path: key packing, cardinality estimate, algorithm
// choice, one aggregation pass per value column, finalisation, partial states and their merge.
//
// Replaces group_by_with_options (split_dataframe/group/grouping.rs:38-115) + GroupBy::aggregate
// (group/aggregation.rs:763-871) + calculate_aggregation (aggregation.rs:500-754) and the Aggregate arm
// of LazyFrame::execute (optimized/lazy.rs:186-404).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "groupby_kernels.cuh"
#include "gb_few.cuh"
#include "gb_final.cuh"
#include "gb_pack.cuh"
#include "comm.cuh"

// ---------------------------------------------------------------- finalisation (formulas + key decoding: gb_final.cuh)
__global__ void gb_finalize_kernel(const FinParams p) {
  const long long total = p.gt.slots + 1;
  const int lane = threadIdx.x & 31;
  for (long long s0 = (long long)blockIdx.x * blockDim.x; s0 < total; s0 += (long long)gridDim.x * blockDim.x) {
    const long long s = s0 + threadIdx.x;
    u64 rw = 0;
    if (s < total) rw = p.gt.hdr[s].rowsw;
    const bool full = (rw & GB_FULL) != 0;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, full);
    if (!m) continue;
    u64 basepos = 0;
    if (lane == 0) basepos = atomicAdd(&p.gt.counters[CNT_OUT], (u64)__popc(m));
    basepos = __shfl_sync(0xFFFFFFFFu, basepos, 0);
    if (!full) continue;
    const long long o = (long long)(basepos + __popc(m & ((1u << lane) - 1u)));
    u64 w[PDRS_MAX_WORDS] = {0, 0, 0};
    const bool nullgroup = s == p.gt.slots;
    if (!nullgroup) {
      w[0] = p.gt.hdr[s].key0;
      if (p.ks.nwords > 1) w[1] = p.gt.kw1[s];
      if (p.ks.nwords > 2) w[2] = p.gt.kw2[s];
    }
    fin_write_group(p, o, w, nullgroup, rw & GB_CNT_MASK, [&](int v) -> const GState* { return p.vals[v].st ? &p.vals[v].st[s] : nullptr; });
  }
}

// ---------------------------------------------------------------- merge of partial states
struct MergeParams {
  KeySpec ks;
  long long n;
  GTable gt;
  int nvals;
  const u64* states[PDRS_MAX_VALS];
  GState* st[PDRS_MAX_VALS];
};
template <int NW>
__global__ void gb_merge_kernel(const MergeParams p) {
  const int lane = threadIdx.x & 31;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i - lane < p.n; i += (long long)gridDim.x * blockDim.x) {
    const bool inb = i < p.n;
    u64 w[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) w[k] = 0;
    bool knull = inb ? load_key_generic<NW>(p.ks, i, w) : false;
    long long gs = g_find_or_insert<NW>(p.gt, w, inb && !knull);
    if (inb && knull) { gs = p.gt.slots; if (!(ld_cg_u64(&p.gt.hdr[gs].rowsw) & GB_FULL)) atomicOr(&p.gt.hdr[gs].rowsw, GB_FULL); }
    if (!inb || gs < 0) continue;
    for (int v = 0; v < p.nvals; v++) {
      const u64* q = p.states[v] + 8 * i;
      GTable t = p.gt;
      t.st = p.st[v];
      const bool have_c = q[2] != 0;
      g_update_batch<GB_ALL, true>(t, gs, v == 0 ? q[0] : 0ull, q[1], fin_pivot(q[2]), have_c,
                                   __longlong_as_double((long long)q[3]), __longlong_as_double((long long)q[4]), q[7], q[5], q[6]);
    }
    if (p.nvals == 0) { /* keys only */ }
  }
}

// ---------------------------------------------------------------- result object
struct pdrs_groupby_result {
  pdrs_ctx* ctx = nullptr;
  int64_t n_groups = 0;
  int nkeys = 0, nvals = 0, naggs = 0;
  int key_dtype[PDRS_MAX_KEYS] = {0, 0, 0, 0};
  DevBuf key_vals[PDRS_MAX_KEYS], key_nulls[PDRS_MAX_KEYS], rows;
  DevBuf validn[PDRS_MAX_VALS], states[PDRS_MAX_VALS];
  DevBuf aggs[PDRS_MAX_AGGS];
};

static int key_out_bytes(int dtype) {
  switch (dtype) { case PDRS_I64: case PDRS_F64: return 8; case PDRS_I32: case PDRS_DICT_U32: return 4; default: return 1; }
}

// ---------------------------------------------------------------- key layout
// Smallest and largest value of every integer key column (signed; dictionary ids as non-negative numbers; values under
// NULL bits are included, which can only widen the range).  out[2 k] = min, out[2 k + 1] = max, pre-set to +max / -max.
__global__ void gb_key_range_kernel(const KeySpec ks, long long n, long long* __restrict__ out) {
  long long mn[PDRS_MAX_KEYS], mx[PDRS_MAX_KEYS];
  for (int k = 0; k < PDRS_MAX_KEYS; k++) { mn[k] = 0x7FFFFFFFFFFFFFFFll; mx[k] = -0x7FFFFFFFFFFFFFFFll - 1; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    for (int k = 0; k < ks.nkeys; k++) {
      long long v;
      switch (ks.c[k].dtype) {
        case PDRS_I64: v = __ldcs((const long long*)ks.c[k].data + i); break;
        case PDRS_I32: v = (long long)__ldcs((const int*)ks.c[k].data + i); break;
        case PDRS_DICT_U32: v = (long long)__ldcs((const uint32_t*)ks.c[k].data + i); break;
        default: continue;
      }
      mn[k] = min(mn[k], v); mx[k] = max(mx[k], v);
    }
  }
  for (int k = 0; k < ks.nkeys; k++) {
    for (int d = 16; d; d >>= 1) { mn[k] = min(mn[k], __shfl_xor_sync(0xFFFFFFFFu, mn[k], d)); mx[k] = max(mx[k], __shfl_xor_sync(0xFFFFFFFFu, mx[k], d)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(out + 2 * k, mn[k]); atomicMax(out + 2 * k + 1, mx[k]); }
  }
}

// Multi-key tuples that need more than one 64-bit word in their natural widths (an Int64 part takes a whole word) but
// whose VALUE RANGES fit one word together - (i32, i64) ids, (returnflag, linestatus, day) ... - are packed as
// (value - column minimum) in ceil(log2(range)) bits each.  One word = the tile-sort / partitioned kernels instead of
// the global table.  Costs one extra read of the key columns; exact ranges, so no row can fall outside its field.
// Used for pdrs_groupby_agg / _partial only: hashes that must agree between ranks (pdrs_hash_partition) and the merge of
// partial states keep the natural layout.
static int32_t compress_keyspec(pdrs_ctx* c, const ColView* kv, int nkeys, long long n, KeySpec* ks) {
  if (nkeys < 2 || ks->nwords < 2 || n < (1 << 16) || c->opt_key_compress == 0) return PDRS_OK;
  for (int k = 0; k < nkeys; k++) if (kv[k].dtype == PDRS_F64) return PDRS_OK;
  DevBuf rng;
  PDRS_TRY(rng.alloc(c, 2 * PDRS_MAX_KEYS * 8));
  long long init[2 * PDRS_MAX_KEYS];
  for (int k = 0; k < PDRS_MAX_KEYS; k++) { init[2 * k] = 0x7FFFFFFFFFFFFFFFll; init[2 * k + 1] = -0x7FFFFFFFFFFFFFFFll - 1; }
  PDRS_CUDA(c, cudaMemcpyAsync(rng.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  gb_key_range_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(*ks, n, rng.as<long long>());
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  long long h[2 * PDRS_MAX_KEYS];
  PDRS_CUDA(c, cudaMemcpyAsync(h, rng.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  int bits[PDRS_MAX_KEYS], total = 0;
  for (int k = 0; k < nkeys; k++) {
    if (kv[k].dtype == PDRS_BOOL_BITS) bits[k] = 1;
    else {
      const unsigned long long range = (unsigned long long)h[2 * k + 1] - (unsigned long long)h[2 * k];     // max - min, exact in 64 bits
      if (h[2 * k + 1] < h[2 * k]) return PDRS_OK;
      int b = 1;
      while (b < 64 && (range >> b) != 0) b++;
      const int natural = kv[k].dtype == PDRS_I64 ? 64 : 32;
      bits[k] = std::min(b, natural);
    }
    total += bits[k] + ((kv[k].nulls || (kv[k].dtype == PDRS_DICT_U32 && kv[k].null_alias >= 0)) ? 1 : 0);
  }
  if (total > 64) return PDRS_OK;
  int used = 0;
  for (int k = 0; k < nkeys; k++) {
    KeyColDev& d = ks->c[k];
    d.word = 0; d.shift = used; d.bits = bits[k]; used += bits[k];
    const int natural = kv[k].dtype == PDRS_I64 ? 64 : (kv[k].dtype == PDRS_BOOL_BITS ? 1 : 32);
    d.offset = (kv[k].dtype == PDRS_BOOL_BITS || bits[k] == natural) ? 0 : h[2 * k];
    if (kv[k].dtype == PDRS_DICT_U32 && d.null_alias >= 0 && bits[k] != natural) d.offset = h[2 * k];
  }
  for (int k = 0; k < nkeys; k++) {
    KeyColDev& d = ks->c[k];
    d.nword = -1; d.nshift = 0;
    if (kv[k].nulls || (kv[k].dtype == PDRS_DICT_U32 && kv[k].null_alias >= 0)) { d.nword = 0; d.nshift = used++; }
  }
  ks->nwords = 1;
  return PDRS_OK;
}

// compat_filter_nulls: the reference's filter() replaces the NULLs of EVERY column of the kept rows by the type default
// and drops the masks (data_ops.rs:64-108, parallel.rs:177-231) - key columns included: a NULL Int64 key joins group "0",
// a NULL string key the group of the empty string.  The fused filter -> groupby does the same by grouping on a
// defaulted copy of every key column that carries a null bitmap (one extra pass over that column, this case only).
template <typename T>
__global__ void gb_key_default_kernel(const T* __restrict__ src, const uint8_t* __restrict__ nulls, long long n, T* __restrict__ out, T dflt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = pdrs_bit(nulls, i) ? dflt : src[i];
}
__global__ void gb_key_default_bits_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ nulls, long long nwords, uint32_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (long long)gridDim.x * blockDim.x) out[i] = src[i] & ~nulls[i];
}
static int32_t default_null_keys(pdrs_ctx* c, ColView* v) {
  if (!v->nulls) return PDRS_OK;
  const long long n = v->len;
  DevBuf out;
  if (v->dtype == PDRS_BOOL_BITS) {
    const long long nwords = (n + 63) / 64 * 2;                      // bitmaps cover ceil(n / 64) * 8 bytes (pdrs_view_col)
    PDRS_TRY(out.alloc(c, (size_t)nwords * 4 + 64, true));
    if (n) gb_key_default_bits_kernel<<<pdrs_grid_for(c, nwords, 256), 256, 0, c->stream>>>((const uint32_t*)v->data, (const uint32_t*)v->nulls, nwords, out.as<uint32_t>());
  } else {
    const int esz = pdrs_dtype_bytes(v->dtype);
    PDRS_TRY(out.alloc(c, (size_t)std::max<long long>(n, 1) * esz + 64));
    const int g = pdrs_grid_for(c, n, 256);
    if (n) switch (v->dtype) {
      case PDRS_I64: case PDRS_F64: gb_key_default_kernel<u64><<<g, 256, 0, c->stream>>>((const u64*)v->data, v->nulls, n, out.as<u64>(), 0ull); break;       // 0 / +0.0
      case PDRS_I32: gb_key_default_kernel<uint32_t><<<g, 256, 0, c->stream>>>((const uint32_t*)v->data, v->nulls, n, out.as<uint32_t>(), 0u); break;
      default: gb_key_default_kernel<uint32_t><<<g, 256, 0, c->stream>>>((const uint32_t*)v->data, v->nulls, n, out.as<uint32_t>(), (uint32_t)c->opt_empty_string_id); break;   // ""
    }
  }
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  v->own_data = std::move(out);
  v->data = v->own_data.p;
  v->nulls = nullptr;
  v->own_nulls.release();
  return PDRS_OK;
}

// Exact cardinality for ANY key distribution from one scan of the key columns: the keys whose hash falls into 1 / 2^slice_bits of
// the hash space are counted exactly in a scratch table (every row is read, only the slice is inserted).  The row sample above,
// inverted under a uniform model, underestimates heavy-tailed key tuples (Zipf) several times over - and a wrong estimate costs
// a whole partition pass.
__global__ void gb_slice_distinct_kernel(const KeySpec ks, long long n, int slice_bits, GTable t) {
  const int lane = threadIdx.x & 31;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i - lane < n; i += (long long)gridDim.x * blockDim.x) {
    u64 w[1] = {0ull};
    bool in = false;
    if (i < n) {
      in = !load_key_inline<1>(ks, i, w);
      const uint32_t lo = (uint32_t)w[0] ^ (uint32_t)(w[0] >> 32);
      in = in && (slice_bits == 0 || ((lo * 0x85EBCA6Bu) >> (32 - slice_bits)) == 0u);
    }
    if (__any_sync(0xFFFFFFFFu, in)) g_find_or_insert<1>(t, w, in);      // (its fail-fast check reads a hot counter line: not for warps with nothing to insert)
  }
}

// keys of the cardinality sample that occurred at least `min_count` times: out[0] = how many, then (key word, occurrences) pairs
__global__ void gb_hot_keys_kernel(const GHdr* __restrict__ hdr, long long slots, u64 min_count, int cap, u64* __restrict__ out) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (long long)gridDim.x * blockDim.x) {
    const u64 rw = hdr[s].rowsw;
    if (!(rw & GB_FULL) || (rw & GB_CNT_MASK) < min_count) continue;
    const u64 at = atomicAdd(&out[0], 1ull);
    if (at < (u64)cap) { out[1 + 2 * at] = hdr[s].key0; out[2 + 2 * at] = rw & GB_CNT_MASK; }
  }
}

// A typed predicate as a Boolean bitmask (general paths; the few-groups kernel evaluates it inside its scan): one 32-bit
// word per thread; ANDed with an optional Boolean filter column (value & ~null).
__global__ void gb_pred_mask_kernel(const u64* __restrict__ col, const uint8_t* __restrict__ cnull, int is_f64, int op, long long ival, double fval,
                                    const uint32_t* __restrict__ fbits, const uint32_t* __restrict__ fnull, long long n, uint32_t* __restrict__ out) {
  const long long nwords = (n + 31) / 32;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (long long)gridDim.x * blockDim.x) {
    uint32_t m = 0;
    for (int b = 0; b < 32; b++) {
      const long long i = w * 32 + b;
      if (i >= n) break;
      const u64 bits = __ldcs(col + i);
      bool k;
      if (is_f64) { const double x = __longlong_as_double((long long)bits);
        k = op == PDRS_CMP_LT ? x < fval : op == PDRS_CMP_LE ? x <= fval : op == PDRS_CMP_GT ? x > fval : op == PDRS_CMP_GE ? x >= fval : op == PDRS_CMP_EQ ? x == fval : x != fval; }
      else { const long long x = (long long)bits;
        k = op == PDRS_CMP_LT ? x < ival : op == PDRS_CMP_LE ? x <= ival : op == PDRS_CMP_GT ? x > ival : op == PDRS_CMP_GE ? x >= ival : op == PDRS_CMP_EQ ? x == ival : x != ival; }
      if (k) m |= 1u << b;
    }
    if (cnull) m &= ~reinterpret_cast<const uint32_t*>(cnull)[w];
    if (fbits) m &= fbits[w];
    if (fnull) m &= ~fnull[w];
    out[w] = m;
  }
}

int32_t pdrs_build_keyspec(pdrs_ctx* c, const ColView* kv, int nkeys, KeySpec* ks) {
  memset(ks, 0, sizeof(*ks));
  ks->nkeys = nkeys;
  if (nkeys == 1) {
    ks->single_null = 1;
    ks->nwords = 1;
    KeyColDev& d = ks->c[0];
    d.data = kv[0].data; d.nulls = kv[0].nulls; d.null_alias = kv[0].null_alias; d.dtype = kv[0].dtype;
    d.word = 0; d.shift = 0; d.nword = -1; d.nshift = 0;
    d.bits = (kv[0].dtype == PDRS_I64 || kv[0].dtype == PDRS_F64) ? 64 : (kv[0].dtype == PDRS_BOOL_BITS ? 1 : 32);
    return PDRS_OK;
  }
  int used[PDRS_MAX_WORDS + 2] = {0, 0, 0, 0, 0};
  int nwords = 0;
  auto place = [&](int bits, int* word, int* shift) -> bool {
    for (int w = 0; w < PDRS_MAX_WORDS; w++) {
      if (used[w] + bits <= 64) { *word = w; *shift = used[w]; used[w] += bits; nwords = std::max(nwords, w + 1); return true; }
    }
    return false;
  };
  bool ok = true;
  for (int pass = 0; pass < 2 && ok; pass++) {      // 64-bit parts first, then the narrow ones
    for (int k = 0; k < nkeys && ok; k++) {
      int bits = (kv[k].dtype == PDRS_I64 || kv[k].dtype == PDRS_F64) ? 64 : (kv[k].dtype == PDRS_BOOL_BITS ? 1 : 32);
      if ((bits == 64) != (pass == 0)) continue;
      KeyColDev& d = ks->c[k];
      d.data = kv[k].data; d.nulls = kv[k].nulls; d.null_alias = kv[k].null_alias; d.dtype = kv[k].dtype; d.bits = bits;
      ok = place(bits, &d.word, &d.shift);
    }
  }
  for (int k = 0; k < nkeys && ok; k++) {
    KeyColDev& d = ks->c[k];
    d.nword = -1;
    if (kv[k].nulls || (kv[k].dtype == PDRS_DICT_U32 && kv[k].null_alias >= 0)) ok = place(1, &d.nword, &d.nshift);
  }
  if (!ok) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "key tuple does not fit %d x 64 bits", PDRS_MAX_WORDS);
  ks->nwords = nwords;
  ks->single_null = 0;
  return PDRS_OK;
}

int32_t gb_radix_pass(pdrs_ctx* c, const GbParams& gp, int is_int, int flags, float* kernel_ms);   // gb_radix.cu
struct GbDirectOut { const FinParams* fin; long long cap; u64* cursor; int dv; };   // finished groups written straight to the result (defined in gb_part.cu too)
int32_t gb_part_pass(pdrs_ctx* c, const GbParams& gp, int is_int, int flags, long long est_groups, float* kernel_ms, bool* dirty, bool* skewed,
                     long long* est_refined, const GbDirectOut* direct = nullptr);   // gb_part.cu
// gb_tsort.cu
bool gb_tsort_geometry(long long cap, bool dense, int smem_budget, int nt_pref, int* nt, int* gpt, int* slots, size_t* smem);
long long gb_tsort_tile_rows();
long long gb_tsort_tile_rows_for(int nt);
cudaError_t gb_tsort_launch(const GbParams& p, int is_int, int flags, int nt, int gpt, int ctas, size_t smem, cudaStream_t s);

// ---------------------------------------------------------------- dispatch helpers
static int variant_of(const KeySpec& ks) {
  if (ks.nkeys == 1 && ks.c[0].dtype == PDRS_I64) return 0;   // k1
  return ks.nwords;                                            // g1, g2, g3
}
static cudaError_t launch_shared(int variant, const GbCfg& cfg, const GbParams& p, size_t smem, cudaStream_t s) {
  switch (variant) {
    case 0: return gb_launch_shared_k1(cfg, p, smem, s);
    case 1: return gb_launch_shared_g1(cfg, p, smem, s);
    case 2: return gb_launch_shared_g2(cfg, p, smem, s);
    default: return gb_launch_shared_g3(cfg, p, smem, s);
  }
}
static cudaError_t launch_global(int variant, const GbCfg& cfg, const GbParams& p, cudaStream_t s) {
  switch (variant) {
    case 0: return gb_launch_global_k1(cfg, p, s);
    case 1: return gb_launch_global_g1(cfg, p, s);
    case 2: return gb_launch_global_g2(cfg, p, s);
    default: return gb_launch_global_g3(cfg, p, s);
  }
}
static cudaError_t launch_sample(int variant, const GbParams& p, long long nb, long long stride, int ctas, cudaStream_t s) {
  switch (variant) {
    case 0: return gb_launch_sample_k1(p, nb, stride, ctas, s);
    case 1: return gb_launch_sample_g1(p, nb, stride, ctas, s);
    case 2: return gb_launch_sample_g2(p, nb, stride, ctas, s);
    default: return gb_launch_sample_g3(p, nb, stride, ctas, s);
  }
}

static long long pow2ceil(long long x) { long long p = 1; while (p < x) p <<= 1; return p; }
static int ilog2(long long x) { int l = 0; while ((1ll << l) < x) l++; return l; }

struct TableMem {
  DevBuf hdr, kw1, kw2, counters;
  GTable t{};
};
static int32_t alloc_table(pdrs_ctx* c, long long slots, int nwords, TableMem* tm) {
  PDRS_TRY(tm->hdr.alloc(c, (size_t)(slots + 1) * sizeof(GHdr), true));
  if (nwords > 1) PDRS_TRY(tm->kw1.alloc(c, (size_t)(slots + 1) * 8));
  if (nwords > 2) PDRS_TRY(tm->kw2.alloc(c, (size_t)(slots + 1) * 8));
  PDRS_TRY(tm->counters.alloc(c, CNT_N * 8, true));
  tm->t.hdr = tm->hdr.as<GHdr>();
  tm->t.kw1 = tm->kw1.as<u64>();
  tm->t.kw2 = tm->kw2.as<u64>();
  tm->t.st = nullptr;
  tm->t.mask = (u64)slots - 1;
  tm->t.shift = 64 - ilog2(slots);
  tm->t.slots = slots;
  tm->t.counters = tm->counters.as<u64>();
  return PDRS_OK;
}
static int32_t read_counters(pdrs_ctx* c, const TableMem& tm, u64* out /*[CNT_N + 1]: + NULL-group flag*/) {
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, tm.t.counters, CNT_N * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + CNT_N, &tm.t.hdr[tm.t.slots].rowsw, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i <= CNT_N; i++) out[i] = (u64)c->pinned_scalars[i];
  return PDRS_OK;
}

// d distinct keys among s uniformly sampled rows -> number of groups under a uniform model
static double invert_distinct(double d, double s) {
  if (d >= s) return 1e18;
  double lo = d, hi = 1e18;
  for (int it = 0; it < 200; it++) {
    double mid = std::sqrt(lo * hi);
    double f = mid * (1.0 - std::exp(-s / mid));
    if (f < d) lo = mid; else hi = mid;
    if (hi / lo < 1.0001) break;
  }
  return std::sqrt(lo * hi);
}

// Estimated number of distinct values among n Int64 keys on the device, from blocks of 256 consecutive rows spread over
// the array (all rows when n <= sample_rows: then the count is exact).
int32_t gb_estimate_groups_i64(pdrs_ctx* c, const u64* keys, long long n, long long sample_rows, long long* est_out) {
  *est_out = 0;
  if (n <= 0) return PDRS_OK;
  const long long s_rows = std::min<long long>(n, std::max<long long>(4096, sample_rows));
  long long nb = (s_rows + 255) / 256, stride = std::max<long long>(256, n / nb);
  if (s_rows >= n) { nb = (n + 255) / 256; stride = 256; }
  TableMem stm;
  PDRS_TRY(alloc_table(c, pow2ceil(4 * nb * 256), 1, &stm));
  GbParams sp{};
  sp.ks.nkeys = 1; sp.ks.nwords = 1; sp.ks.single_null = 1;
  sp.ks.c[0].data = keys; sp.ks.c[0].dtype = PDRS_I64; sp.ks.c[0].bits = 64; sp.ks.c[0].nword = -1; sp.ks.c[0].null_alias = -1;
  sp.n = n;
  sp.gt = stm.t;
  PDRS_CUDA(c, launch_sample(0, sp, nb, stride, (int)std::min<long long>(nb, c->sm_count * 8), c->stream));
  c->stats.kernel_launches++;
  u64 cn[CNT_N + 1];
  PDRS_TRY(read_counters(c, stm, cn));
  const double d = (double)cn[CNT_NGROUPS], s = (double)std::min<long long>(n, nb * 256);
  *est_out = s_rows >= n ? (long long)d : (long long)std::min<double>((double)n, invert_distinct(d, s) * 1.05 + 1.0);
  return PDRS_OK;
}

struct PassPlan { int val; int flags; int is_int; };

enum { MODE_AGG = 0, MODE_PARTIAL = 1 };

static int32_t groupby_run(pdrs_ctx* c, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals,
                           const pdrs_agg* aggs, int32_t naggs, const pdrs_col* filter, int mode, int partial_all,
                           pdrs_groupby_result** out, const pdrs_pred* pred = nullptr) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!out || !keys || nkeys < 1 || nkeys > PDRS_MAX_KEYS || nvals < 0 || nvals > PDRS_MAX_VALS || naggs < 0 || naggs > PDRS_MAX_AGGS ||
      (nvals && !vals) || (naggs && !aggs))
    return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_groupby: bad argument (nkeys %d, nvals %d, naggs %d)", nkeys, nvals, naggs);
  PDRS_CUDA(c, cudaSetDevice(c->device));
  pdrs_settle_frees(c);
  const int64_t n = keys[0].len;
  for (int k = 0; k < nkeys; k++) if (keys[k].len != n) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "key column %d has %lld rows, expected %lld", k, (long long)keys[k].len, (long long)n);
  for (int v = 0; v < nvals; v++) if (vals[v].len != n) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "value column %d has %lld rows, expected %lld", v, (long long)vals[v].len, (long long)n);
  if (filter && (filter->dtype != PDRS_BOOL_BITS || filter->len != n)) return pdrs_fail(c, PDRS_ERR_TYPE_MISMATCH, "filter must be a Boolean column of the same length");
  if (pred && ((pred->col.dtype != PDRS_I64 && pred->col.dtype != PDRS_F64) || pred->col.len != n || pred->op < PDRS_CMP_LT || pred->op > PDRS_CMP_NE))
    return pdrs_fail(c, PDRS_ERR_TYPE_MISMATCH, "predicate: needs an Int64 / Float64 column of the same length and a pdrs_cmp_op");

  // ---- which statistics does each value column need
  int need[PDRS_MAX_VALS];   // -1 unused, GB_SUM, GB_ALL
  for (int v = 0; v < nvals; v++) need[v] = mode == MODE_PARTIAL ? (partial_all ? GB_ALL : GB_SUM) : -1;
  for (int a = 0; a < naggs; a++) {
    const pdrs_agg& ag = aggs[a];
    if (ag.op < PDRS_SUM || ag.op > PDRS_VAR) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "aggregate %d: unknown op %d", a, ag.op);
    if (ag.op == PDRS_COUNT) { if (ag.value_col >= nvals) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "aggregate %d: value column %d out of range", a, ag.value_col); continue; }
    if (ag.value_col < 0 || ag.value_col >= nvals) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "aggregate %d: value column %d out of range", a, ag.value_col);
    const int dt = vals[ag.value_col].dtype;
    if (dt != PDRS_I64 && dt != PDRS_F64)   // aggregation.rs:748-752
      return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "aggregate %d: op %d is not supported on a column of dtype %d (only Count is)", a, ag.op, dt);
    const int f = (ag.op == PDRS_SUM || ag.op == PDRS_MEAN) ? GB_SUM : GB_ALL;
    need[ag.value_col] = std::max(need[ag.value_col], f);
  }
  if (mode == MODE_PARTIAL) for (int v = 0; v < nvals; v++) if (vals[v].dtype != PDRS_I64 && vals[v].dtype != PDRS_F64) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "partial aggregation needs Int64/Float64 value columns");

  auto* res = new pdrs_groupby_result();
  res->ctx = c; res->nkeys = nkeys; res->nvals = nvals; res->naggs = naggs;
  for (int k = 0; k < nkeys; k++) res->key_dtype[k] = keys[k].dtype;
  struct Guard { pdrs_groupby_result* r; ~Guard() { delete r; } } guard{res};

  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
  c->stats.groupby_algo_used = 0; c->stats.retries = 0; c->stats.est_groups = 0; c->stats.table_slots = 0; c->stats.spilled_rows = 0;
  c->stats.main_kernel_ms = 0; c->stats.total_ms = 0;

  // ---- views
  std::vector<ColView> kv(nkeys), vv(nvals);
  ColView fv;
  for (int k = 0; k < nkeys; k++) PDRS_TRY(pdrs_view_col(c, &keys[k], &kv[k]));
  for (int v = 0; v < nvals; v++) if (need[v] >= 0) PDRS_TRY(pdrs_view_col(c, &vals[v], &vv[v]));
  ColView pcv;
  if (filter) PDRS_TRY(pdrs_view_col(c, filter, &fv));
  if (pred) PDRS_TRY(pdrs_view_col(c, &pred->col, &pcv));
  const bool filtered = filter || pred;
  if (filtered && c->opts.compat_filter_nulls) for (int k = 0; k < nkeys; k++) PDRS_TRY(default_null_keys(c, &kv[k]));

  KeySpec ks;
  PDRS_TRY(pdrs_build_keyspec(c, kv.data(), nkeys, &ks));
  PDRS_TRY(compress_keyspec(c, kv.data(), nkeys, n, &ks));
  const int variant = variant_of(ks);

  std::vector<PassPlan> passes;
  for (int v = 0; v < nvals; v++) if (need[v] >= 0) passes.push_back({v, need[v], vals[v].dtype == PDRS_I64});
  if (passes.empty()) passes.push_back({-1, GB_SUM, 0});

  GbParams base{};
  base.ks = ks;
  base.n = n;
  base.fbits = filter ? (const uint8_t*)fv.data : nullptr;
  base.fnull = filter ? fv.nulls : nullptr;
  base.compat_nulls = (filtered && c->opts.compat_filter_nulls) ? 1 : 0;

  // ---- few-groups kernel (gb_few.cu): one-word key tuples without NULLs, sum / mean / count only, <= 8 value columns
  bool few_eligible = c->opt_few != 0 && ks.nwords == 1 && n >= 4096 && (int)passes.size() <= GF_MAXV &&
                      (c->opts.groupby_algo == PDRS_GB_AUTO || c->opts.groupby_algo == PDRS_GB_FEW);
  for (int k = 0; k < nkeys; k++) if (kv[k].nulls || (kv[k].dtype == PDRS_DICT_U32 && kv[k].null_alias >= 0)) few_eligible = false;
  for (auto& pp : passes) if (pp.flags != GB_SUM) few_eligible = false;
  std::vector<u64> few_keys;
  DevBuf hot_dev;

  // ---- cardinality estimate
  long long est = c->opts.groups_hint > 0 ? c->opts.groups_hint : 0;
  bool dense_ok = false;
  long long dense_base = 0, dense_range = 0;
  if (n > 0 && (est == 0 || (est <= GF_MAXG && few_eligible))) {
    const long long hint = est;
    long long s_rows = std::min<long long>(n, std::max<long long>(4096, c->opt_sample_rows));
    long long nb = (s_rows + 255) / 256;
    long long stride = std::max<long long>(256, n / nb);
    if (s_rows >= n) { nb = (n + 255) / 256; stride = 256; }
    TableMem stm;
    PDRS_TRY(alloc_table(c, pow2ceil(4 * nb * 256), ks.nwords, &stm));
    GbParams sp = base;
    sp.gt = stm.t;
    PDRS_CUDA(c, launch_sample(variant, sp, nb, stride, (int)std::min<long long>(nb, c->sm_count * 8), c->stream));
    c->stats.kernel_launches++;
    u64 cn[CNT_N + 1];
    PDRS_TRY(read_counters(c, stm, cn));
    double d = (double)cn[CNT_NGROUPS], s = (double)std::min<long long>(n, nb * 256);
    if (s_rows >= n) est = (long long)d;
    else est = (long long)std::min<double>((double)n, invert_distinct(d, s) * 1.05 + 1.0);
    if (est < 1) est = 1;
    if (hint > 0) est = hint;
    c->stats.est_groups = est;
    // <= 16 groups and only sum / mean / count: the few-groups kernel scans all value columns at once (gb_few.cu); it needs the
    // group keys up front - the sample table holds them
    if (est <= GF_MAXG && few_eligible) {
      DevBuf kb;
      PDRS_TRY(kb.alloc(c, (GF_MAXG + 2) * 8, true));
      PDRS_CUDA(c, gb_few_collect_keys(stm.t, kb.as<u64>(), c->sm_count * 4, c->stream));
      c->stats.kernel_launches++;
      u64 hk[GF_MAXG + 2];
      PDRS_CUDA(c, cudaMemcpyAsync(hk, kb.p, sizeof(hk), cudaMemcpyDeviceToHost, c->stream));
      PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
      if (hk[0] >= 1 && hk[0] <= GF_MAXG) { few_keys.assign(hk + 1, hk + 1 + hk[0]); std::sort(few_keys.begin(), few_keys.end()); }
    }
    // hot keys of a skewed distribution (one-word keys, high cardinality): the partitioned path routes their rows around the
    // hash partitions, so that no partition overflows and every partition's groups can be written straight to the result
    if (ks.nwords == 1 && est > 2047 && s_rows < n && c->opt_part != 0 && c->opt_part_hot != 0) {
      const int HOTCAP = 4096;
      DevBuf hb;
      PDRS_TRY(hb.alloc(c, (size_t)(2 * HOTCAP + 2) * 8, true));
      gb_hot_keys_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(stm.t.hdr, stm.t.slots, 6ull, HOTCAP, hb.as<u64>());
      c->stats.kernel_launches++;
      std::vector<u64> hh((size_t)2 * HOTCAP + 2);
      PDRS_CUDA(c, cudaMemcpyAsync(hh.data(), hb.p, hh.size() * 8, cudaMemcpyDeviceToHost, c->stream));
      PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
      const size_t nh = (size_t)std::min<u64>(hh[0], HOTCAP);
      std::vector<std::pair<u64, u64>> hk;      // (occurrences, key)
      for (size_t i = 0; i < nh; i++) hk.push_back({hh[2 + 2 * i], hh[1 + 2 * i]});
      std::sort(hk.begin(), hk.end(), [](const std::pair<u64, u64>& a, const std::pair<u64, u64>& b) { return a.first > b.first; });
      if (hk.size() > 1800) hk.resize(1800);
      if (!hk.empty()) {
        std::vector<uint32_t> tab((size_t)1 << GB_HOT_LOG_SLOTS, 0u);
        u64 hot_rows = 0;
        for (auto& kc : hk) {
          uint32_t sl = gb_hot_slot(kc.second, GB_HOT_LOG_SLOTS);
          while (tab[sl] != 0u) sl = (sl + 1) & ((1u << GB_HOT_LOG_SLOTS) - 1u);
          tab[sl] = gb_hot_tag(kc.second);
          hot_rows += kc.first;
        }
        PDRS_TRY(hot_dev.alloc(c, tab.size() * 4));
        PDRS_CUDA(c, cudaMemcpyAsync(hot_dev.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, c->stream));
        PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
        base.hot_tab = hot_dev.as<uint32_t>(); base.hot_log_slots = GB_HOT_LOG_SLOTS; base.hot_frac = (float)((double)hot_rows / s);
      }
    }
    if (variant == 0 && cn[CNT_KMINC]) {     // small dense integer keys: direct-mapped group ids, no key table
      const long long kmin = (long long)(~cn[CNT_KMINC] ^ GB_SIGN), kmax = (long long)(cn[CNT_KMAX] ^ GB_SIGN);
      const unsigned long long range = (unsigned long long)kmax - (unsigned long long)kmin + 1ull;
      if (range <= 60000ull && (long long)range <= 2 * est + 64) { dense_ok = true; dense_base = kmin; dense_range = (long long)range; }
    }
  }
  if (est < 1) est = 1;
  bool est_exact = false;
  if (variant == 1 && ks.nwords == 1 && est > 2047 && n >= (1ll << 22) && c->opts.groups_hint == 0 && c->opt_part != 0 && c->opts.groupby_algo == PDRS_GB_AUTO) {
    // packed multi-column / dictionary keys headed for the partitioned path: these are the heavy-tailed ones in practice
    int sb = 0;
    while ((n >> sb) > (1ll << 21)) sb++;
    TableMem stm;
    PDRS_TRY(alloc_table(c, pow2ceil(4 * ((n >> sb) + 1024)), 1, &stm));
    gb_slice_distinct_kernel<<<pdrs_grid_for(c, n, 256), 256, 0, c->stream>>>(ks, n, sb, stm.t);
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
    u64 cn2[CNT_N + 1];
    PDRS_TRY(read_counters(c, stm, cn2));
    const long long exact = (long long)cn2[CNT_NGROUPS] << sb;
    est = std::max<long long>(1, exact + exact / 32);
    c->stats.est_groups = est;
    est_exact = true;
  }

  pdrs_trace(c, "gb: views + estimate");
  // ---- geometry of the shared-memory kernel per pass; fall back to the global table when it does not fit
  auto shared_geometry = [&](const PassPlan& pp, GbParams* gp, GbCfg* cfg, size_t* smem) -> bool {
    const int rec_bytes = (pp.flags == GB_SUM ? 8 : 16) + ((pp.flags == GB_ALL && pp.is_int) ? 8 : 0) + 4;   // ShPlanes::REC_BYTES
    const int cta_bytes = (pp.flags == GB_ALL ? 32 : 0) + 4;                                                   // ShPlanes::CTA_BYTES
    const bool dense = variant == 0 && dense_ok && c->opt_dense != 0;
    long long cap = dense ? dense_range + std::max<long long>(dense_range / 16, 8) : est + std::max<long long>(est / 16, 8);
    cap = (cap + 7) / 8 * 8;
    if (cap > 60000) return false;
    long long S = dense ? 0 : std::max<long long>(64, pow2ceil(cap + cap / 2));
    const size_t fixed = gb_sh_fixed_bytes(ks.nwords, (int)S, (int)cap, cta_bytes, dense);
    const size_t budget = (size_t)c->smem_optin;
    int warps = c->opt_warps > 0 ? (int)std::min<int64_t>(c->opt_warps, GB_MAX_WARPS) : GB_MAX_WARPS;
    int ng = 0;
    for (;;) {
      for (int g = 32; g >= 1; g >>= 1) {
        if (c->opt_ng > 0 && g > c->opt_ng) continue;
        if (fixed + (size_t)warps * gb_sh_warp_bytes(rec_bytes, (int)cap, g) <= budget) { ng = g; break; }
      }
      if (ng || warps <= 4 || c->opt_warps > 0) break;
      warps--;
    }
    if (!ng) return false;
    gp->sh_cap = (int)cap; gp->sh_slots = (int)S; gp->sh_log_slots = S ? ilog2(S) : 0; gp->sh_ng = ng;
    gp->sh_dense = dense ? 1 : 0; gp->sh_dense_base = dense_base;
    cfg->warps = warps;
    cfg->ctas = c->sm_count * (c->opt_ctas_per_sm > 0 ? (int)c->opt_ctas_per_sm : 1);
    const long long units = (n + GB_UNIT_ROWS - 1) / GB_UNIT_ROWS;
    cfg->ctas = (int)std::max<long long>(1, std::min<long long>(cfg->ctas, (units + warps - 1) / warps));
    *smem = fixed + (size_t)warps * gb_sh_warp_bytes(rec_bytes, (int)cap, ng);
    return true;
  };

  // ---- tile-sort kernel (gb_tsort.cu): one 64-bit key column, tens to ~2000 groups
  int ts_nt = 0, ts_gpt = 0, ts_slots = 0;
  size_t ts_smem = 0;
  bool ts_dense = false;
  long long ts_cap = 0;
  bool ts_fit = false;
  const bool ts_generic = variant == 1;     // any key tuple that packs into one 64-bit word (dictionary ids, i32, bool, pairs of them)
  if ((variant == 0 || ts_generic) && c->opt_tsort != 0 && est >= c->opt_tsort_min_groups && n >= 4 * gb_tsort_tile_rows() && n < (1ll << 38) &&
      (c->opts.groupby_algo == PDRS_GB_AUTO || c->opts.groupby_algo == PDRS_GB_TILESORT)) {
    ts_dense = dense_ok && c->opt_dense != 0;
    ts_cap = ts_dense ? dense_range + std::max<long long>(dense_range / 16, 8) : est + std::max<long long>(est / 16, 8);
    // one group per thread when the keys seen by the sample fit 1023 ids (later keys outside the range spill)
    const long long ts_seen = ts_dense ? dense_range : est;
    // (the sampled estimate is inflated by ~5%: up to 1100 / 2150 estimated groups still try the smaller geometry; keys
    //  that do not fit spill to the global table)
    if (ts_cap > 1023 && ts_seen <= (ts_dense ? 1023 : 1100)) ts_cap = 1023;
    if (ts_cap > 2047 && ts_seen <= (ts_dense ? 2047 : 2150)) ts_cap = 2047;
    ts_fit = gb_tsort_geometry(ts_cap, ts_dense, c->smem_optin, (int)c->opt_tsort_threads, &ts_nt, &ts_gpt, &ts_slots, &ts_smem);
  }
  if (c->opts.groupby_algo == PDRS_GB_TILESORT && !ts_fit && c->opts.groups_hint > 0)
    return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "groupby_algo=TILESORT: needs one Int64 key column and at most ~2000 groups (estimated %lld)", est);

  // result arrays for `cap` groups + the finalisation parameters that point at them (state pointers are filled in later)
  FinParams fp{};
  fp.ks = ks;
  // keep > 0: the arrays already hold `keep` finished groups (written straight to the result) that move to the new arrays
  auto alloc_outputs = [&](size_t cap, size_t keep = 0) -> int32_t {
    auto renew = [&](DevBuf& b, size_t elem) -> int32_t {
      DevBuf nb;
      PDRS_TRY(nb.alloc(c, cap * elem));
      if (keep && b.p) PDRS_CUDA(c, cudaMemcpyAsync(nb.p, b.p, keep * elem, cudaMemcpyDeviceToDevice, c->stream));
      b = std::move(nb);
      return PDRS_OK;
    };
    for (int k = 0; k < nkeys; k++) {
      PDRS_TRY(renew(res->key_vals[k], key_out_bytes(keys[k].dtype)));
      PDRS_TRY(renew(res->key_nulls[k], 1));
      fp.key_out[k] = res->key_vals[k].p;
      fp.key_null_out[k] = res->key_nulls[k].as<uint8_t>();
    }
    PDRS_TRY(renew(res->rows, 8));
    fp.rows_out = res->rows.as<long long>();
    fp.nvals = nvals;
    for (int v = 0; v < nvals; v++) { fp.vals[v].st = nullptr; fp.vals[v].is_int = vals[v].dtype == PDRS_I64; fp.vals[v].flags = std::max(need[v], 0); fp.vals[v].validn_out = nullptr; fp.vals[v].states_out = nullptr; }
    for (size_t i = 0; i < passes.size(); i++) {
      const int v = passes[i].val;
      if (v < 0) continue;
      PDRS_TRY(renew(res->validn[v], 8));
      fp.vals[v].validn_out = res->validn[v].as<long long>();
      if (mode == MODE_PARTIAL) {
        PDRS_TRY(renew(res->states[v], 64));
        fp.vals[v].states_out = res->states[v].as<u64>();
      }
    }
    fp.naggs = naggs;
    for (int a = 0; a < naggs; a++) {
      PDRS_TRY(renew(res->aggs[a], 8));
      fp.aggs[a].val = aggs[a].op == PDRS_COUNT ? -1 : aggs[a].value_col;
      fp.aggs[a].op = aggs[a].op;
      fp.aggs[a].out = res->aggs[a].as<double>();
    }
    return PDRS_OK;
  };
  // Partitioned path with ONE value column: finished groups of complete partitions go straight to the result (gb_part.cu);
  // direct_cap = room in the result arrays, direct_cur[0] = groups written that way.
  DevBuf direct_cur;
  long long direct_cap = 0, direct_groups = 0;
  bool direct_off = c->opt_part_direct == 0;

  int algo = c->opts.groupby_algo;
  if (algo == PDRS_GB_DENSE || algo == PDRS_GB_TILESORT) algo = PDRS_GB_SHARED;
  long long slots_mult = 1;
  TableMem tm;
  std::vector<DevBuf> states(passes.size());
  u64 cn[CNT_N + 1] = {0};
  bool radix_used = false, restart = false;
  bool part_ok = c->opt_part != 0 && c->opts.groupby_algo == PDRS_GB_AUTO;
  bool reestimated = est_exact;  // the partitioned path may correct the sampled cardinality estimate once (gb_part_pass)
  bool ts_skew = false;          // the tile-sort kernel runs as the skew fallback: its spills go to a side buffer + a second pass
  bool few_off = false;          // the few-groups kernel met too many keys the sample had not seen
  DevBuf pred_mask;
  (void)radix_used;
  for (int attempt = 0;; attempt++) {
    if (attempt > 6) return pdrs_fail(c, PDRS_ERR_OOM, "groupby: hash table kept overflowing after %d retries", attempt);
    bool use_shared = algo != PDRS_GB_GLOBAL;
    std::vector<GbParams> gps(passes.size(), base);
    std::vector<GbCfg> cfgs(passes.size());
    std::vector<size_t> smems(passes.size(), 0);
    if (use_shared) {
      for (size_t i = 0; i < passes.size(); i++) {
        cfgs[i] = GbCfg{ks.nwords, variant == 0 ? 0 : 1, passes[i].is_int, passes[i].flags, 0, 0};
        if (!shared_geometry(passes[i], &gps[i], &cfgs[i], &smems[i])) { use_shared = false; break; }
      }
    }
    if (!use_shared && !ts_fit && algo == PDRS_GB_SHARED && attempt == 0 && c->opts.groups_hint > 0)
      return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "groupby_algo=SHARED: %lld groups do not fit shared memory", est);
    long long slots = std::max<long long>(1024, pow2ceil(2 * est + 1024)) * slots_mult;
    tm = TableMem();
    PDRS_TRY(alloc_table(c, slots, ks.nwords, &tm));
    c->stats.table_slots = slots;
    c->stats.groupby_algo_used = use_shared ? PDRS_GB_SHARED : PDRS_GB_GLOBAL;
    bool few_ok = !few_keys.empty() && !few_off && gb_few_smem((int)few_keys.size(), passes[0].val >= 0 ? (int)passes.size() : 0, true) <= (size_t)c->smem_optin;
    if (n > 0 && few_ok) {
      GfParams fp{};
      fp.base = base;
      fp.base.gt = tm.t;
      fp.base.count_rows = 1;
      fp.nv = passes[0].val >= 0 ? (int)passes.size() : 0;
      for (int i = 0; i < fp.nv; i++) {
        PDRS_TRY(states[i].alloc(c, (size_t)(slots + 1) * sizeof(GState), true));
        fp.val[i] = vv[passes[i].val].data; fp.vnull[i] = vv[passes[i].val].nulls; fp.st[i] = states[i].as<GState>(); fp.is_int[i] = passes[i].is_int;
        if (fp.vnull[i] && !base.compat_nulls) fp.any_vnull = 1;
      }
      fp.ng = (int)few_keys.size();
      for (int g = 0; g < fp.ng; g++) fp.gkey[g] = few_keys[g];
      if (pred) { fp.pcol = pcv.data; fp.pnull = pcv.nulls; fp.pdtype = pcv.dtype; fp.pop = pred->op; fp.pival = pred->ival; fp.pfval = pred->fval; }
      const size_t smem = gb_few_smem(fp.ng, fp.nv, fp.any_vnull != 0);
      const long long units = (n + 63) / 64;
      const int ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * std::max<size_t>(1, std::min<size_t>(3, (size_t)c->smem_optin / std::max<size_t>(smem, 1))), (units + 7) / 8));
      if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
      PDRS_CUDA(c, gb_few_launch(fp, ctas, smem, c->stream));
      c->stats.kernel_launches++;
      c->stats.groupby_algo_used = PDRS_GB_FEW;
      if (c->opt_timing) {
        PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
        PDRS_CUDA(c, cudaEventSynchronize(c->ev_b));
        float ms = 0;
        PDRS_CUDA(c, cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
        c->stats.main_kernel_ms += ms;
      }
    } else if (n > 0) {
      if (pred && !pred_mask.p) {      // general paths take the predicate as a Boolean mask (one more pass over that column)
        const long long nwords = (n + 63) / 64 * 2;
        PDRS_TRY(pred_mask.alloc(c, (size_t)nwords * 4 + 64, true));
        gb_pred_mask_kernel<<<pdrs_grid_for(c, (n + 31) / 32, 256), 256, 0, c->stream>>>(reinterpret_cast<const u64*>(pcv.data), pcv.nulls, pcv.dtype == PDRS_F64, pred->op, pred->ival, pred->fval,
            reinterpret_cast<const uint32_t*>(base.fbits), reinterpret_cast<const uint32_t*>(base.fnull), n, pred_mask.as<uint32_t>());
        c->stats.kernel_launches++;
        PDRS_CUDA(c, cudaGetLastError());
        base.fbits = pred_mask.as<uint8_t>(); base.fnull = nullptr;
        for (auto& g : gps) { g.fbits = base.fbits; g.fnull = nullptr; }
      }
      for (size_t i = 0; i < passes.size(); i++) {
        GbParams& gp = gps[i];
        gp.gt = tm.t;
        gp.count_rows = i == 0;
        if (passes[i].val >= 0) {
          PDRS_TRY(states[i].alloc(c, (size_t)(slots + 1) * sizeof(GState), true));
          gp.gt.st = states[i].as<GState>();
          gp.val = vv[passes[i].val].data;
          gp.vnull = vv[passes[i].val].nulls;
        }
        // high cardinality, one 64-bit key column: hash-partition the rows, tile-sort kernel per partition (gb_part.cu)
        if ((variant == 0 || variant == 1) && part_ok && !ts_fit && algo != PDRS_GB_GLOBAL && passes[i].val >= 0 && est > 2047) {
          float ms = 0;
          bool dirty = false, skewed = false;
          long long est_new = 0;
          GbDirectOut dout{};
          if (passes.size() == 1 && !direct_off) {
            direct_cap = std::min<long long>(n, 4 * est + (1 << 20));      // heavy-tailed tuples: the estimate is often low; partitions that find no room flush to the table
            PDRS_TRY(alloc_outputs((size_t)direct_cap));
            PDRS_TRY(direct_cur.alloc(c, 64, true));
            fp.gt = gp.gt;
            dout.fin = &fp; dout.cap = direct_cap; dout.cursor = direct_cur.as<u64>(); dout.dv = passes[i].val;
          }
          int32_t rs = gb_part_pass(c, gp, passes[i].is_int, passes[i].flags, est, &ms, &dirty, &skewed, reestimated ? nullptr : &est_new, dout.fin ? &dout : nullptr);
          if (rs == PDRS_OK) { c->stats.main_kernel_ms += ms; c->stats.groupby_algo_used = PDRS_GB_PARTITIONED; continue; }
          direct_cap = 0;
          if (rs != PDRS_ERR_UNSUPPORTED) return rs;
          if (est_new > 0) {        // the first bucket shows far more groups than the sample predicted: start over with that estimate
            est = est_new; c->stats.est_groups = est; reestimated = true; restart = true;
            break;
          }
          part_ok = false;
          // Skewed keys (a hash bucket overflowed its padded range): a few hot keys carry most of the rows.  The
          // tile-sort kernel keeps the first 2047 keys every CTA meets (the hot ones, with high probability) in its
          // register accumulators and sends the rows of the remaining keys to the global table one by one.
          if (skewed && est <= 1000000 && c->opt_tsort != 0 && n >= 4 * gb_tsort_tile_rows() && n < (1ll << 38)) {
            ts_dense = false; ts_cap = 2047;
            ts_fit = gb_tsort_geometry(ts_cap, false, c->smem_optin, (int)c->opt_tsort_threads, &ts_nt, &ts_gpt, &ts_slots, &ts_smem);
            ts_skew = ts_fit;
          }
          if (dirty || i > 0) { restart = true; break; }     // the table already holds counts of this attempt: start over without this path
        }
        // high cardinality, one 64-bit key column: radix-partitioned rows + L2-resident table regions (gb_radix.cu)
        const size_t tbytes = (size_t)(slots + 1) * (sizeof(GHdr) + (gp.gt.st ? sizeof(GState) : 0));
        if (!use_shared && variant == 0 && c->opt_radix != 0 && tbytes > (64ull << 20) && n >= (1 << 20)) {
          float ms = 0;
          int32_t rs = gb_radix_pass(c, gp, passes[i].is_int, passes[i].flags, &ms);
          if (rs == PDRS_OK) { c->stats.main_kernel_ms += ms; c->stats.groupby_algo_used = PDRS_GB_GLOBAL; radix_used = true; continue; }
          if (rs != PDRS_ERR_UNSUPPORTED) return rs;
        }
        const bool use_ts = ts_fit && algo != PDRS_GB_GLOBAL && passes[i].val >= 0;
        DevBuf spill_k, spill_v;
        if (use_ts && ts_skew && c->opt_spillbuf != 0) {
          const long long cap = std::max<long long>(n / 3, 1 << 20);
          size_t free_b = 0;
          PDRS_TRY(pdrs_mem_available(c, &free_b));
          if ((size_t)cap * 16 + (4ull << 30) < free_b) {
            PDRS_TRY(spill_k.alloc(c, (size_t)cap * 8));
            PDRS_TRY(spill_v.alloc(c, (size_t)cap * 8));
            gp.gt.spill_k = spill_k.as<u64>(); gp.gt.spill_v = spill_v.as<u64>(); gp.gt.spill_cap = cap;
            PDRS_CUDA(c, cudaMemsetAsync(gp.gt.counters + CNT_SPILLBUF, 0, 8, c->stream));
          }
        }
        if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
        if (use_ts) {
          gp.sh_cap = (int)ts_cap; gp.sh_slots = ts_slots; gp.sh_log_slots = ts_slots ? ilog2(ts_slots) : 0;
          gp.sh_dense = ts_dense ? 1 : 0; gp.sh_dense_base = dense_base;
          gp.ts_heavy = c->opt_tsort_heavy > 0 ? (int)c->opt_tsort_heavy : 256;
          gp.ts_generic = ts_generic ? 1 : 0;
          gp.ts_mid = c->opt_tsort_mid > 0 ? (int)c->opt_tsort_mid : 48;
          gp.ts_team = (est < 500 || c->opt_tsort_team == 1) && c->opt_tsort_team != 2;   // long segments are expected (skewed keys spill over to the whole-warp reduce at ts_heavy)
          const long long trows = gb_tsort_tile_rows_for(ts_nt), tiles = (n + trows - 1) / trows;
          PDRS_CUDA(c, gb_tsort_launch(gp, passes[i].is_int, passes[i].flags, ts_nt, ts_gpt, (int)std::min<long long>((long long)c->sm_count * (ts_nt == 256 ? 2 : 1), tiles), ts_smem, c->stream));
          c->stats.groupby_algo_used = PDRS_GB_TILESORT;
        } else if (use_shared) PDRS_CUDA(c, launch_shared(variant, cfgs[i], gp, smems[i], c->stream));
        else {
          GbCfg cfg{ks.nwords, variant == 0 ? 0 : 1, passes[i].is_int, passes[i].flags, 8, 0};
          cfg.ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 8, (n + 1023) / 1024));
          PDRS_CUDA(c, launch_global(variant, cfg, gp, c->stream));
        }
        if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
        c->stats.kernel_launches++;
        if (c->opt_timing) {
          PDRS_CUDA(c, cudaEventSynchronize(c->ev_b));
          float ms = 0;
          PDRS_CUDA(c, cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
          c->stats.main_kernel_ms += ms;
        }
        if (gp.gt.spill_k) {
          // second pass of the skew fallback: the side buffer holds the rows of the keys outside the per-CTA hot sets
          // as (key word, value) - an Int64-keyed groupby into the SAME table; without the hot keys the hash
          // partitions are balanced, so the partitioned path applies (global-table kernel when it does not)
          PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 9, gp.gt.counters + CNT_SPILLBUF, 8, cudaMemcpyDeviceToHost, c->stream));
          PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
          const long long ns = std::min<long long>((long long)c->pinned_scalars[9], gp.gt.spill_cap);
          if (ns > 0) {
            GbParams sp = gp;
            memset(&sp.ks, 0, sizeof(sp.ks));
            sp.ks.nkeys = 1; sp.ks.nwords = 1; sp.ks.single_null = 1;
            sp.ks.c[0].data = gp.gt.spill_k; sp.ks.c[0].dtype = PDRS_I64; sp.ks.c[0].bits = 64; sp.ks.c[0].nword = -1; sp.ks.c[0].null_alias = -1;
            sp.val = gp.gt.spill_v; sp.vnull = nullptr; sp.fbits = nullptr; sp.fnull = nullptr; sp.compat_nulls = 0; sp.n = ns;
            sp.gt.spill_k = nullptr; sp.gt.spill_v = nullptr; sp.gt.spill_cap = 0;
            sp.ts_generic = 0;
            float ms = 0;
            bool dirty = false, skewed = false;
            int32_t rs = gb_part_pass(c, sp, passes[i].is_int, passes[i].flags, est, &ms, &dirty, &skewed, nullptr);
            if (rs != PDRS_OK && rs != PDRS_ERR_UNSUPPORTED) return rs;
            if (rs == PDRS_ERR_UNSUPPORTED) {
              GbCfg cfg{1, 0, passes[i].is_int, passes[i].flags, 8, 0};
              cfg.ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 8, (ns + 1023) / 1024));
              if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
              PDRS_CUDA(c, launch_global(0, cfg, sp, c->stream));
              c->stats.kernel_launches++;
              if (c->opt_timing) {
                PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
                PDRS_CUDA(c, cudaEventSynchronize(c->ev_b));
                PDRS_CUDA(c, cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
              }
            }
            c->stats.main_kernel_ms += ms;
          }
        }
      }
    }
    if (restart) { restart = false; for (auto& s : states) s.release(); direct_cap = 0; attempt--; continue; }
    PDRS_TRY(read_counters(c, tm, cn));
    c->stats.spilled_rows = (int64_t)cn[CNT_SPILLED];
    if (cn[CNT_OVERFLOW] == 0 && cn[CNT_SPIN_FAIL] == 0) break;
    c->stats.retries++;
    slots_mult *= 4;
    few_off = true;
    direct_cap = 0;
    if ((use_shared || ts_fit) && (long long)cn[CNT_SPILLED] > n / 16) algo = PDRS_GB_GLOBAL;
    for (auto& s : states) s.release();
  }

  pdrs_trace(c, "gb: aggregation passes");
  // ---- finalise: the groups of the table are appended behind the groups that were written straight to the result
  const int64_t Gt = (int64_t)cn[CNT_NGROUPS] + ((cn[CNT_N] & GB_FULL) ? 1 : 0);
  if (direct_cap > 0) {
    PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 12, direct_cur.p, 8, cudaMemcpyDeviceToHost, c->stream));
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    direct_groups = c->pinned_scalars[12];
    if (direct_groups + Gt > direct_cap) PDRS_TRY(alloc_outputs((size_t)(direct_groups + Gt), (size_t)direct_groups));      // the table holds more groups than expected: larger arrays, the direct groups move over
    PDRS_CUDA(c, cudaMemcpyAsync(tm.t.counters + CNT_OUT, direct_cur.p, 8, cudaMemcpyDeviceToDevice, c->stream));
  } else {
    PDRS_TRY(alloc_outputs((size_t)std::max<int64_t>(Gt, 1)));
  }
  const int64_t G = Gt + direct_groups;
  res->n_groups = G;
  fp.gt = tm.t;
  for (size_t i = 0; i < passes.size(); i++) if (passes[i].val >= 0) fp.vals[passes[i].val].st = states[i].as<GState>();
  if (Gt > 0) {
    int g = pdrs_grid_for(c, tm.t.slots + 1, 256);
    gb_finalize_kernel<<<g, 256, 0, c->stream>>>(fp);
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  if (c->opt_timing) {
    PDRS_CUDA(c, cudaEventRecord(c->ev_t1, c->stream));
    PDRS_CUDA(c, cudaEventSynchronize(c->ev_t1));
    PDRS_CUDA(c, cudaEventElapsedTime(&c->stats.total_ms, c->ev_t0, c->ev_t1));
  } else {
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  pdrs_trace(c, "gb: finalise");
  guard.r = nullptr;
  *out = res;
  return PDRS_OK;
}

// Tail of every merge of partial states: read the table counters, allocate the result arrays, finalise.
static std::vector<int> keys_dtypes(const pdrs_col* keys, int nkeys) { std::vector<int> d(nkeys); for (int k = 0; k < nkeys; k++) d[k] = keys[k].dtype; return d; }
static int32_t finish_merged(pdrs_ctx* c, const KeySpec& ks, TableMem& tm, std::vector<DevBuf>& states, const int* key_dtype, int nkeys, int nvals,
                             const int32_t* val_is_int, const pdrs_agg* aggs, int naggs, pdrs_groupby_result* res) {
  u64 cn[CNT_N + 1];
  PDRS_TRY(read_counters(c, tm, cn));
  if (cn[CNT_OVERFLOW] || cn[CNT_SPIN_FAIL]) return pdrs_fail(c, PDRS_ERR_CUDA, "merge: hash table overflow");
  const int64_t G = (int64_t)cn[CNT_NGROUPS] + ((cn[CNT_N] & GB_FULL) ? 1 : 0);
  res->n_groups = G;
  FinParams fp{};
  fp.gt = tm.t; fp.ks = ks;
  const size_t Galloc = (size_t)std::max<int64_t>(G, 1);
  for (int k = 0; k < nkeys; k++) {
    PDRS_TRY(res->key_vals[k].alloc(c, Galloc * key_out_bytes(key_dtype[k])));
    PDRS_TRY(res->key_nulls[k].alloc(c, Galloc));
    fp.key_out[k] = res->key_vals[k].p;
    fp.key_null_out[k] = res->key_nulls[k].as<uint8_t>();
  }
  PDRS_TRY(res->rows.alloc(c, Galloc * 8));
  fp.rows_out = res->rows.as<long long>();
  fp.nvals = nvals;
  for (int v = 0; v < nvals; v++) {
    fp.vals[v].st = states[v].as<GState>();
    fp.vals[v].is_int = val_is_int ? val_is_int[v] : 0;
    fp.vals[v].flags = GB_ALL;
    PDRS_TRY(res->validn[v].alloc(c, Galloc * 8));
    fp.vals[v].validn_out = res->validn[v].as<long long>();
    PDRS_TRY(res->states[v].alloc(c, Galloc * 64));
    fp.vals[v].states_out = res->states[v].as<u64>();
  }
  fp.naggs = naggs;
  for (int a = 0; a < naggs; a++) {
    PDRS_TRY(res->aggs[a].alloc(c, Galloc * 8));
    fp.aggs[a].val = aggs[a].op == PDRS_COUNT ? -1 : aggs[a].value_col;
    fp.aggs[a].op = aggs[a].op;
    fp.aggs[a].out = res->aggs[a].as<double>();
  }
  if (G > 0) {
    gb_finalize_kernel<<<pdrs_grid_for(c, tm.t.slots + 1, 256), 256, 0, c->stream>>>(fp);
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}

static bool stream_eligible(pdrs_ctx* c, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals, const pdrs_col* filter, const pdrs_pred* pred);
static int32_t groupby_stream(pdrs_ctx* c, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals, const pdrs_agg* aggs, int32_t naggs,
                              const pdrs_col* filter, const pdrs_pred* pred, pdrs_groupby_result** out, int partial = 0);

extern "C" {

int32_t pdrs_groupby_agg(pdrs_ctx* ctx, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals,
                         const pdrs_agg* aggs, int32_t naggs, const pdrs_col* filter, pdrs_groupby_result** out) {
  if (stream_eligible(ctx, keys, nkeys, vals, nvals, filter, nullptr)) return groupby_stream(ctx, keys, nkeys, vals, nvals, aggs, naggs, filter, nullptr, out);
  return groupby_run(ctx, keys, nkeys, vals, nvals, aggs, naggs, filter, MODE_AGG, 0, out);
}

int32_t pdrs_groupby_agg_where(pdrs_ctx* ctx, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals,
                               const pdrs_agg* aggs, int32_t naggs, const pdrs_col* filter, const pdrs_pred* pred, pdrs_groupby_result** out) {
  if (stream_eligible(ctx, keys, nkeys, vals, nvals, filter, pred)) return groupby_stream(ctx, keys, nkeys, vals, nvals, aggs, naggs, filter, pred, out);
  return groupby_run(ctx, keys, nkeys, vals, nvals, aggs, naggs, filter, MODE_AGG, 0, out, pred);
}

int32_t pdrs_groupby_partial(pdrs_ctx* ctx, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals,
                             const pdrs_col* filter, int32_t all_stats, pdrs_groupby_result** out) {
  return groupby_run(ctx, keys, nkeys, vals, nvals, nullptr, 0, filter, MODE_PARTIAL, all_stats, out);
}

int32_t pdrs_groupby_merge(pdrs_ctx* c, const pdrs_col* keys, int32_t nkeys, const uint64_t* const* states_dev, const int32_t* val_is_int,
                           int32_t nvals, int64_t n_state_rows, const pdrs_agg* aggs, int32_t naggs, pdrs_groupby_result** out) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!out || !keys || nkeys < 1 || nkeys > PDRS_MAX_KEYS || nvals < 0 || nvals > PDRS_MAX_VALS || naggs < 0 || naggs > PDRS_MAX_AGGS || n_state_rows < 0)
    return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_groupby_merge: bad argument");
  PDRS_CUDA(c, cudaSetDevice(c->device));
  for (int k = 0; k < nkeys; k++) if (keys[k].len != n_state_rows) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "merge: key column %d length mismatch", k);
  for (int a = 0; a < naggs; a++) if (aggs[a].op != PDRS_COUNT && (aggs[a].value_col < 0 || aggs[a].value_col >= nvals)) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "merge: aggregate %d: bad value column", a);
  auto* res = new pdrs_groupby_result();
  res->ctx = c; res->nkeys = nkeys; res->nvals = nvals; res->naggs = naggs;
  for (int k = 0; k < nkeys; k++) res->key_dtype[k] = keys[k].dtype;
  struct Guard { pdrs_groupby_result* r; ~Guard() { delete r; } } guard{res};
  std::vector<ColView> kv(nkeys);
  for (int k = 0; k < nkeys; k++) PDRS_TRY(pdrs_view_col(c, &keys[k], &kv[k]));
  KeySpec ks;
  PDRS_TRY(pdrs_build_keyspec(c, kv.data(), nkeys, &ks));
  const int64_t n = n_state_rows;
  TableMem tm;
  const long long slots = std::max<long long>(1024, pow2ceil(2 * n + 16));
  PDRS_TRY(alloc_table(c, slots, ks.nwords, &tm));
  std::vector<DevBuf> states(nvals);
  MergeParams mp{};
  mp.ks = ks; mp.n = n; mp.gt = tm.t; mp.nvals = nvals;
  for (int v = 0; v < nvals; v++) {
    PDRS_TRY(states[v].alloc(c, (size_t)(slots + 1) * sizeof(GState), true));
    mp.states[v] = (const u64*)states_dev[v];
    mp.st[v] = states[v].as<GState>();
  }
  if (n > 0) {
    int g = pdrs_grid_for(c, n, 256);
    switch (ks.nwords) {
      case 1: gb_merge_kernel<1><<<g, 256, 0, c->stream>>>(mp); break;
      case 2: gb_merge_kernel<2><<<g, 256, 0, c->stream>>>(mp); break;
      default: gb_merge_kernel<3><<<g, 256, 0, c->stream>>>(mp); break;
    }
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  PDRS_TRY(finish_merged(c, ks, tm, states, keys_dtypes(keys, nkeys).data(), nkeys, nvals, val_is_int, aggs, naggs, res));
  guard.r = nullptr;
  *out = res;
  return PDRS_OK;
}

}  // extern "C"

// ================================================================ multi-GPU groupby behind one call (pdrs_groupby_agg_dist)
// One rank = one process = one GPU; the communicator (comm.cu) carries NCCL.  Semantics = the single-frame groupby of the
// reference (grouping.rs:38-115 + aggregation.rs:500-903) applied to the UNION of the ranks' rows:
//   every rank aggregates its own rows into mergeable states (the pdrs_groupby_partial kernels), packs one row per group
//   [key words | NULL-group flag | group rows | 8 state words per value column], and then
//   replicated  (few groups: <= groups_cap per rank) ONE fixed-size ncclAllGather of the packed rows, a merge kernel over the
//               ranks' rows (Chan-style re-basing of S1 / S2 to a common pivot, like pdrs_groupby_merge) and the usual
//               finalisation - every rank returns ALL groups.  No row ever crosses NVLink, no host round trip in between.
//   sharded     (many groups) the packed rows are scattered by destination rank = hash(key) mod ranks, exchanged with one
//               all-to-all (grouped ncclSend / ncclRecv) and merged by their owner - every rank returns the groups it owns.
//               Only one row per (rank, group) crosses NVLink, not the input rows.
// The key layout is the natural one with a NULL flag reserved for EVERY key part, so that ranks agree on it whatever
// null bitmaps their shards happen to carry.
template <int NW>
__device__ __forceinline__ int dist_dest(const u64 (&w)[NW], bool nullgroup, int world) {
  if (nullgroup) return 0;                       // the NULL-key group lives on rank 0 (like pdrs_hash_partition)
  return (int)(pdrs_mix64(key_hash<NW>(w) ^ 0x5851F42D4C957F2Dull) % (u64)world);
}
// MODE 0: row j -> out[8 + j * stride] (replicated; out[0] = G).  MODE 1: count the rows per destination rank.
// MODE 2: scatter into out[(off[dest] + ticket) * stride].  MODE 3: row j -> out[j * stride] (plain append, no header).
template <int NW, int MODE>
__global__ void dist_pack_kernel(const DistPack p, u64* __restrict__ out, u64* __restrict__ counts, const u64* __restrict__ off) {
  if (MODE == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = (u64)p.G;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < p.G; j += (long long)gridDim.x * blockDim.x) {
    if (MODE == 0 && j >= p.cap) break;
    u64 w[NW];
    const bool ng = dist_row_words<NW>(p, j, w);
    if (MODE == 1) { atomicAdd(&counts[dist_dest<NW>(w, ng, p.world)], 1ull); continue; }
    long long at = j;
    if (MODE == 2) { const int d = dist_dest<NW>(w, ng, p.world); at = (long long)(off[d] + atomicAdd(&counts[d], 1ull)); }
    u64* q = out + (MODE == 0 ? 8 : 0) + at * p.stride;   // MODE 3: at = j
    q[0] = w[0]; q[1] = NW > 1 ? w[NW > 1 ? 1 : 0] : 0ull; q[2] = NW > 2 ? w[NW > 2 ? 2 : 0] : 0ull;
    q[3] = ng ? 1ull : 0ull;
    q[4] = (u64)p.rows[j];
    for (int v = 0; v < p.nvals; v++) {
      const u64* s = p.states[v] + 8 * j;
#pragma unroll
      for (int i = 0; i < 8; i++) q[5 + 8 * v + i] = s[i];
    }
  }
}

struct DistMerge {
  const u64* buf; long long nrows; long long cap; long long block;      // block > 0: rank r's rows start at buf[r * block + 8], count = buf[r * block]
  int stride, nvals;
  GTable gt;
  GState* st[PDRS_MAX_VALS];
};
template <int NW>
__global__ void dist_merge_kernel(const DistMerge p) {
  const int lane = threadIdx.x & 31;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i - lane < p.nrows; i += (long long)gridDim.x * blockDim.x) {
    bool inb = i < p.nrows;
    const u64* q = p.buf + i * p.stride;
    if (p.block > 0 && inb) {
      const long long r = i / p.cap, j = i - r * p.cap;
      inb = j < (long long)p.buf[r * p.block];
      q = p.buf + r * p.block + 8 + j * p.stride;
    }
    u64 w[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) w[k] = inb ? q[k] : 0ull;
    const bool knull = inb && q[3] != 0;
    long long gs = g_find_or_insert<NW>(p.gt, w, inb && !knull);
    if (knull) { gs = p.gt.slots; if (!(ld_cg_u64(&p.gt.hdr[gs].rowsw) & GB_FULL)) atomicOr(&p.gt.hdr[gs].rowsw, GB_FULL); }
    if (!inb || gs < 0) continue;
    atomicAdd(&p.gt.hdr[gs].rowsw, q[4]);
    for (int v = 0; v < p.nvals; v++) {
      const u64* s = q + 5 + 8 * v;
      GTable t = p.gt;
      t.st = p.st[v];
      g_update_batch<GB_ALL, true>(t, gs, 0ull, s[1], fin_pivot(s[2]), s[2] != 0, __longlong_as_double((long long)s[3]), __longlong_as_double((long long)s[4]), s[7], s[5], s[6]);
    }
  }
}

static int32_t dist_merge_launch(pdrs_ctx* c, int nw, const DistMerge& mp) {
  if (mp.nrows <= 0) return PDRS_OK;
  const int g = pdrs_grid_for(c, mp.nrows, 256);
  switch (nw) {
    case 1: dist_merge_kernel<1><<<g, 256, 0, c->stream>>>(mp); break;
    case 2: dist_merge_kernel<2><<<g, 256, 0, c->stream>>>(mp); break;
    default: dist_merge_kernel<3><<<g, 256, 0, c->stream>>>(mp); break;
  }
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  return PDRS_OK;
}
template <int MODE>
static int32_t dist_pack_launch(pdrs_ctx* c, int nw, const DistPack& pk, u64* out, u64* counts, const u64* off) {
  const int g = pdrs_grid_for(c, std::max<long long>(pk.G, 1), 256);
  switch (nw) {
    case 1: dist_pack_kernel<1, MODE><<<g, 256, 0, c->stream>>>(pk, out, counts, off); break;
    case 2: dist_pack_kernel<2, MODE><<<g, 256, 0, c->stream>>>(pk, out, counts, off); break;
    default: dist_pack_kernel<3, MODE><<<g, 256, 0, c->stream>>>(pk, out, counts, off); break;
  }
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  return PDRS_OK;
}

// ================================================================ chunked groupby over HOST columns (out-of-core inputs)
// SURVEY.md 8(f) row 4 / src/large/mod.rs (ChunkedDataFrame: a frame larger than memory is processed chunk by chunk): the same
// groupby(keys).agg(...) as pdrs_groupby_agg, for columns that live in HOST memory.  Instead of staging whole columns (which caps
// the input at what fits in HBM next to the work areas, and leaves the GPU idle during the transfer) the rows are cut into chunks:
//   chunk i + 1 travels host -> device (staging engine of stage.cu: pinned sources by direct DMA, pageable sources through the
//   worker threads' pinned slots) WHILE chunk i is aggregated into mergeable per-group states (the kernels of pdrs_groupby_partial);
//   the states of all chunks are appended as packed rows [key words | NULL-group flag | rows | 8 words per value column] and merged
//   at the end exactly like the ranks' states of pdrs_groupby_agg_dist (Chan-style re-basing of S1 / S2; compaction in between
//   when the appended rows outgrow the groups).  Device memory: two chunk buffers per column + the states, whatever n is.
static bool stream_eligible(pdrs_ctx* c, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals, const pdrs_col* filter, const pdrs_pred* pred) {
  if (!c || !keys || nkeys < 1 || nkeys > PDRS_MAX_KEYS || nvals < 0 || nvals > PDRS_MAX_VALS || (nvals && !vals)) return false;
  if (c->opt_stream_rows <= 0 || keys[0].len < c->opt_stream_rows) return false;
  const int64_t n = keys[0].len;
  for (int k = 0; k < nkeys; k++) if (keys[k].mem != PDRS_MEM_HOST || keys[k].len != n || keys[k].dtype < 0 || keys[k].dtype > PDRS_I32) return false;
  for (int v = 0; v < nvals; v++) if (vals[v].mem != PDRS_MEM_HOST || vals[v].len != n || vals[v].dtype < 0 || vals[v].dtype > PDRS_I32) return false;
  if (filter && (filter->mem != PDRS_MEM_HOST || filter->len != n || filter->dtype != PDRS_BOOL_BITS)) return false;
  if (pred && (pred->col.mem != PDRS_MEM_HOST || pred->col.len != n)) return false;
  return true;
}

namespace {
struct StreamCol {              // one host column and its two device chunk buffers
  const pdrs_col* h = nullptr;
  DevBuf data[2], nulls[2];
  size_t data_cap = 0, null_cap = 0;
};
}  // namespace

// partial: 0 = the caller's aggregates; 1 / 2 = mergeable states of ALL value columns instead ({rows, n, sum} / everything), the
// result pdrs_groupby_partial would give (used by pdrs_groupby_agg_dist for its local step)
static int32_t groupby_stream(pdrs_ctx* c, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals, const pdrs_agg* aggs, int32_t naggs,
                              const pdrs_col* filter, const pdrs_pred* pred, pdrs_groupby_result** out, int partial) {
  if (partial) { aggs = nullptr; naggs = 0; }
  if (!out || naggs < 0 || naggs > PDRS_MAX_AGGS || (naggs && !aggs)) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_groupby: bad argument");
  PDRS_CUDA(c, cudaSetDevice(c->device));
  const int64_t n = keys[0].len;
  bool all_stats = false;
  for (int a = 0; a < naggs; a++) {
    if (aggs[a].op < PDRS_SUM || aggs[a].op > PDRS_VAR) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "aggregate %d: unknown op %d", a, aggs[a].op);
    if (aggs[a].op == PDRS_COUNT) { if (aggs[a].value_col >= nvals) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "aggregate %d: value column %d out of range", a, aggs[a].value_col); continue; }
    if (aggs[a].value_col < 0 || aggs[a].value_col >= nvals) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "aggregate %d: value column %d out of range", a, aggs[a].value_col);
    const int dt = vals[aggs[a].value_col].dtype;
    if (dt != PDRS_I64 && dt != PDRS_F64) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "aggregate %d: op %d is not supported on a column of dtype %d (only Count is)", a, aggs[a].op, dt);
    if (aggs[a].op != PDRS_SUM && aggs[a].op != PDRS_MEAN) all_stats = true;
  }
  if (pred && ((pred->col.dtype != PDRS_I64 && pred->col.dtype != PDRS_F64) || pred->op < PDRS_CMP_LT || pred->op > PDRS_CMP_NE))
    return pdrs_fail(c, PDRS_ERR_TYPE_MISMATCH, "predicate: needs an Int64 / Float64 column of the same length and a pdrs_cmp_op");
  // only the value columns an aggregate reads travel
  std::vector<int> vmap(nvals, -1);
  std::vector<const pdrs_col*> used;
  if (partial) {
    all_stats = partial == 2;
    for (int v = 0; v < nvals; v++) {
      if (vals[v].dtype != PDRS_I64 && vals[v].dtype != PDRS_F64) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "partial aggregation needs Int64/Float64 value columns");
      vmap[v] = v; used.push_back(&vals[v]);
    }
  }
  for (int a = 0; a < naggs; a++) if (aggs[a].op != PDRS_COUNT && vmap[aggs[a].value_col] < 0) { vmap[aggs[a].value_col] = (int)used.size(); used.push_back(&vals[aggs[a].value_col]); }
  const int nvs = (int)used.size();
  std::vector<pdrs_agg> ag(aggs, aggs + naggs);
  for (auto& a : ag) if (a.op != PDRS_COUNT) a.value_col = vmap[a.value_col];
  std::vector<int32_t> val_is_int(std::max(nvs, 1));
  for (int v = 0; v < nvs; v++) val_is_int[v] = used[v]->dtype == PDRS_I64;

  long long chunk = c->opt_stream_chunk_rows > 0 ? c->opt_stream_chunk_rows : (1ll << 26);
  chunk = std::max<long long>(1 << 16, (chunk + 65535) / 65536 * 65536);      // bitmap words and 128-bit loads stay aligned
  const long long nchunks = (n + chunk - 1) / chunk;

  const int ncols = nkeys + nvs + (filter ? 1 : 0) + (pred ? 1 : 0);
  std::vector<StreamCol> cols(ncols);
  {
    int i = 0;
    for (int k = 0; k < nkeys; k++) cols[i++].h = &keys[k];
    for (int v = 0; v < nvs; v++) cols[i++].h = used[v];
    if (filter) cols[i++].h = filter;
    if (pred) cols[i++].h = &pred->col;
  }
  const long long crow = std::min<long long>(chunk, n);
  for (auto& sc : cols) {
    sc.data_cap = (sc.h->dtype == PDRS_BOOL_BITS ? (size_t)(crow + 7) / 8 : (size_t)crow * pdrs_dtype_bytes(sc.h->dtype)) + 128;
    sc.null_cap = (size_t)(crow + 7) / 8 + 128;
    for (int b = 0; b < 2 && b < nchunks; b++) {
      PDRS_TRY(sc.data[b].alloc(c, sc.data_cap, sc.h->dtype == PDRS_BOOL_BITS));
      if (sc.h->null_bits) PDRS_TRY(sc.nulls[b].alloc(c, sc.null_cap, true));
    }
  }
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));      // the buffers exist before the staging streams write into them

  auto rows_of = [&](long long i) { return std::min<long long>(chunk, n - i * chunk); };
  // queue the copies of chunk i into buffer i & 1 (its previous user, chunk i - 2, has been aggregated: the caller synchronised)
  auto issue = [&](long long i) -> int32_t {
    const long long r0 = i * chunk, rows = rows_of(i);
    const int b = (int)(i & 1);
    bool cleared = false;
    for (auto& sc : cols) {
      const bool bits = sc.h->dtype == PDRS_BOOL_BITS;
      const size_t need = (size_t)(rows + 7) / 8;
      long long have = 0;
      if (sc.h->null_bits) have = std::max<long long>(0, std::min<long long>((long long)need, sc.h->null_len - r0 / 8));
      // a partial last chunk / a short null mask must not see the bits of the chunk that used the buffer before
      if (rows < chunk || (sc.h->null_bits && (size_t)have < need)) {
        if (bits) PDRS_CUDA(c, cudaMemsetAsync(sc.data[b].p, 0, sc.data_cap, c->stream));
        if (sc.h->null_bits) PDRS_CUDA(c, cudaMemsetAsync(sc.nulls[b].p, 0, sc.null_cap, c->stream));
        cleared = true;
      }
    }
    if (cleared) PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& sc : cols) {
      const bool bits = sc.h->dtype == PDRS_BOOL_BITS;
      const size_t esz = bits ? 0 : (size_t)pdrs_dtype_bytes(sc.h->dtype);
      const size_t nb = bits ? (size_t)(rows + 7) / 8 : (size_t)rows * esz;
      const char* src = (const char*)sc.h->data + (bits ? (size_t)(r0 / 8) : (size_t)r0 * esz);
      PDRS_TRY(pdrs_stage_copy_async(c, sc.data[b].p, src, nb));
      if (sc.h->null_bits) {
        const long long have = std::max<long long>(0, std::min<long long>((rows + 7) / 8, sc.h->null_len - r0 / 8));
        if (have > 0) PDRS_TRY(pdrs_stage_copy_async(c, sc.nulls[b].p, sc.h->null_bits + r0 / 8, (size_t)have));
      }
    }
    return PDRS_OK;
  };
  auto dev_col = [&](const StreamCol& sc, long long i) {
    pdrs_col d = *sc.h;
    const int b = (int)(i & 1);
    d.mem = PDRS_MEM_DEVICE;
    d.len = rows_of(i);
    d.data = sc.data[b].p;
    d.null_bits = sc.h->null_bits ? sc.nulls[b].as<uint8_t>() : nullptr;
    d.null_len = sc.h->null_bits ? (int64_t)sc.null_cap : 0;
    return d;
  };

  // the key layout every chunk's states are packed in: natural widths, a NULL flag for every key part
  KeySpec ks;
  {
    std::vector<ColView> kv(nkeys);
    for (int k = 0; k < nkeys; k++) { kv[k].dtype = keys[k].dtype; kv[k].len = 0; kv[k].nulls = reinterpret_cast<const uint8_t*>(1); kv[k].null_alias = -1; }
    PDRS_TRY(pdrs_build_keyspec(c, kv.data(), nkeys, &ks));
    for (int k = 0; k < nkeys; k++) { ks.c[k].data = nullptr; ks.c[k].nulls = nullptr; ks.c[k].null_alias = -1; kv[k].nulls = nullptr; }
  }
  const int NW = ks.nwords, stride = 5 + 8 * nvs;
  DevBuf acc;
  long long acc_rows = 0, acc_cap = 0;
  auto reserve = [&](long long rows) -> int32_t {
    if (rows <= acc_cap) return PDRS_OK;
    const long long ncap = std::max<long long>(rows + rows / 2, 1 << 16);
    DevBuf nb;
    PDRS_TRY(nb.alloc(c, (size_t)ncap * stride * 8));
    if (acc_rows) PDRS_CUDA(c, cudaMemcpyAsync(nb.p, acc.p, (size_t)acc_rows * stride * 8, cudaMemcpyDeviceToDevice, c->stream));
    acc = std::move(nb);
    acc_cap = ncap;
    return PDRS_OK;
  };
  auto append = [&](const pdrs_groupby_result* part) -> int32_t {
    if (part->n_groups == 0) return PDRS_OK;
    PDRS_TRY(reserve(acc_rows + part->n_groups));
    DistPack pk{};
    pk.ks = ks; pk.nvals = nvs; pk.G = part->n_groups; pk.cap = part->n_groups; pk.stride = stride; pk.world = 1; pk.rows = part->rows.as<long long>();
    for (int k = 0; k < nkeys; k++) { pk.key_vals[k] = part->key_vals[k].p; pk.key_null[k] = part->key_nulls[k].as<uint8_t>(); }
    for (int v = 0; v < nvs; v++) pk.states[v] = part->states[v].as<u64>();
    PDRS_TRY(dist_pack_launch<3>(c, NW, pk, acc.as<u64>() + acc_rows * stride, nullptr, nullptr));
    acc_rows += part->n_groups;
    return PDRS_OK;
  };
  // merge the appended rows: `fin` = with the caller's aggregates (the result), else states only (compaction)
  auto merge = [&](bool fin, pdrs_groupby_result* res) -> int32_t {
    res->ctx = c; res->nkeys = nkeys; res->nvals = nvs; res->naggs = fin ? naggs : 0;
    for (int k = 0; k < nkeys; k++) res->key_dtype[k] = keys[k].dtype;
    TableMem tm;
    std::vector<DevBuf> states(nvs);
    DistMerge mp{};
    mp.stride = stride; mp.nvals = nvs;
    const long long slots = std::max<long long>(1024, pow2ceil(2 * acc_rows + 16));
    PDRS_TRY(alloc_table(c, slots, NW, &tm));
    for (int v = 0; v < nvs; v++) { PDRS_TRY(states[v].alloc(c, (size_t)(slots + 1) * sizeof(GState), true)); mp.st[v] = states[v].as<GState>(); }
    mp.gt = tm.t;
    mp.buf = acc.as<u64>(); mp.nrows = acc_rows; mp.cap = 0; mp.block = 0;
    PDRS_TRY(dist_merge_launch(c, NW, mp));
    return finish_merged(c, ks, tm, states, res->key_dtype, nkeys, nvs, val_is_int.data(), fin ? ag.data() : nullptr, fin ? naggs : 0, res);
  };

  std::vector<pdrs_col> kc(nkeys), vc(std::max(nvs, 1));
  pdrs_pred pd{};
  if (pred) pd = *pred;
  float kernel_ms = 0;
  int algo = 0;
  long long spilled = 0, est = 0, last_groups = 0;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  if (c->opt_timing) { PDRS_CUDA(c, cudaEventCreate(&t0)); PDRS_CUDA(c, cudaEventCreate(&t1)); PDRS_CUDA(c, cudaEventRecord(t0, c->stream)); }
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } evg{t0, t1};
  int32_t st = issue(0);
  for (long long i = 0; i < nchunks && st == PDRS_OK; i++) {
    st = pdrs_stage_join(c, c->stream);                       // chunk i is on its way; the aggregation below is ordered behind it
    if (st == PDRS_OK && i + 1 < nchunks) st = issue(i + 1);  // chunk i + 1 travels while chunk i is aggregated
    if (st != PDRS_OK) break;
    int j = 0;
    for (int k = 0; k < nkeys; k++) kc[k] = dev_col(cols[j++], i);
    for (int v = 0; v < nvs; v++) vc[v] = dev_col(cols[j++], i);
    pdrs_col fc{};
    if (filter) fc = dev_col(cols[j++], i);
    if (pred) pd.col = dev_col(cols[j++], i);
    pdrs_groupby_result* part = nullptr;
    st = groupby_run(c, kc.data(), nkeys, vc.data(), nvs, nullptr, 0, filter ? &fc : nullptr, MODE_PARTIAL, all_stats ? 1 : 0, &part, pred ? &pd : nullptr);
    if (st != PDRS_OK) break;
    kernel_ms += c->stats.main_kernel_ms; algo = c->stats.groupby_algo_used; spilled += c->stats.spilled_rows; est = std::max<long long>(est, c->stats.est_groups);
    last_groups = part->n_groups;
    st = append(part);
    cudaStreamSynchronize(c->stream);       // the buffers of chunk i are free again (and `part` may go)
    delete part;
    // compaction: the appended rows are re-merged into one row per group once they outgrow the groups
    if (st == PDRS_OK && i + 1 < nchunks && acc_rows > std::max<long long>(4 * last_groups, c->opt_stream_compact_rows)) {
      pdrs_groupby_result tmp;
      st = merge(false, &tmp);
      if (st == PDRS_OK) { acc_rows = 0; st = append(&tmp); cudaStreamSynchronize(c->stream); }
    }
  }
  if (st != PDRS_OK) { pdrs_stage_join(c, c->stream); cudaStreamSynchronize(c->stream); return st; }
  auto* res = new pdrs_groupby_result();
  st = merge(!partial, res);
  if (st != PDRS_OK) { delete res; return st; }
  // the caller's value column numbering
  if (nvs != nvals) {
    DevBuf vn[PDRS_MAX_VALS], vs[PDRS_MAX_VALS];
    for (int v = 0; v < nvals; v++) if (vmap[v] >= 0) { vn[v] = std::move(res->validn[vmap[v]]); vs[v] = std::move(res->states[vmap[v]]); }
    for (int v = 0; v < PDRS_MAX_VALS; v++) { res->validn[v] = std::move(vn[v]); res->states[v] = std::move(vs[v]); }
  }
  res->nvals = nvals;
  c->stats.main_kernel_ms = kernel_ms; c->stats.groupby_algo_used = algo; c->stats.spilled_rows = spilled; c->stats.est_groups = est;
  if (c->opt_timing) {
    PDRS_CUDA(c, cudaEventRecord(t1, c->stream));
    PDRS_CUDA(c, cudaEventSynchronize(t1));
    PDRS_CUDA(c, cudaEventElapsedTime(&c->stats.total_ms, t0, t1));
  }
  *out = res;
  return PDRS_OK;
}

extern "C" int32_t pdrs_groupby_agg_dist(pdrs_comm* cm, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals, const pdrs_agg* aggs, int32_t naggs,
                                         const pdrs_col* filter, const pdrs_pred* pred, int32_t result_mode, pdrs_groupby_result** out) {
  if (!cm) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = cm->ctx;
  if (!out || !keys || nkeys < 1 || nkeys > PDRS_MAX_KEYS || nvals < 0 || nvals > PDRS_MAX_VALS || naggs < 0 || naggs > PDRS_MAX_AGGS || (naggs && !aggs) ||
      result_mode < 0 || result_mode > 2)
    return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_groupby_agg_dist: bad argument");
  bool all_stats = false;
  for (int a = 0; a < naggs; a++) {
    if (aggs[a].op < PDRS_SUM || aggs[a].op > PDRS_VAR) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "aggregate %d: unknown op %d", a, aggs[a].op);
    if (aggs[a].op != PDRS_COUNT && (aggs[a].value_col < 0 || aggs[a].value_col >= nvals)) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "aggregate %d: value column out of range", a);
    if (aggs[a].op != PDRS_COUNT && vals[aggs[a].value_col].dtype != PDRS_I64 && vals[aggs[a].value_col].dtype != PDRS_F64)
      return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "aggregate %d: op %d needs an Int64 / Float64 column", a, aggs[a].op);
    if (aggs[a].op == PDRS_MIN || aggs[a].op == PDRS_MAX || aggs[a].op == PDRS_STD || aggs[a].op == PDRS_VAR) all_stats = true;
  }
  // value columns that no aggregate reads travel as nothing: only Int64 / Float64 columns have states
  std::vector<pdrs_col> nv;
  std::vector<int> vmap(nvals, -1);
  for (int v = 0; v < nvals; v++) if (vals[v].dtype == PDRS_I64 || vals[v].dtype == PDRS_F64) { vmap[v] = (int)nv.size(); nv.push_back(vals[v]); }
  std::vector<pdrs_agg> ag(aggs, aggs + naggs);
  for (auto& a : ag) if (a.op != PDRS_COUNT) a.value_col = vmap[a.value_col];
  const int nvs = (int)nv.size();
  // ---- local partial aggregation (every rank; the kernels of pdrs_groupby_partial)
  pdrs_trace(c, nullptr);
  pdrs_groupby_result* part = nullptr;
  if (stream_eligible(c, keys, nkeys, nv.data(), nvs, filter, pred))      // host shards: chunk by chunk through the staging engine
    PDRS_TRY(groupby_stream(c, keys, nkeys, nv.data(), nvs, nullptr, 0, filter, pred, &part, all_stats ? 2 : 1));
  else
    PDRS_TRY(groupby_run(c, keys, nkeys, nv.data(), nvs, nullptr, 0, filter, MODE_PARTIAL, all_stats ? 1 : 0, &part, pred));
  struct PartGuard { pdrs_groupby_result* r; ~PartGuard() { if (r) { cudaSetDevice(r->ctx->device); delete r; } } } pguard{part};
  pdrs_trace(c, "dist: local partial");
  const float local_ms = c->stats.main_kernel_ms;
  const int local_algo = c->stats.groupby_algo_used;
  // ---- the layout all ranks agree on
  KeySpec ks;
  {
    std::vector<ColView> kv(nkeys);
    for (int k = 0; k < nkeys; k++) { kv[k].dtype = keys[k].dtype; kv[k].len = 0; kv[k].nulls = reinterpret_cast<const uint8_t*>(1); kv[k].null_alias = -1; }
    PDRS_TRY(pdrs_build_keyspec(c, kv.data(), nkeys, &ks));
    for (int k = 0; k < nkeys; k++) { ks.c[k].data = nullptr; ks.c[k].nulls = nullptr; ks.c[k].null_alias = -1; kv[k].nulls = nullptr; }
  }
  const int NW = ks.nwords, stride = 5 + 8 * nvs, world = cm->world;
  DistPack pk{};
  pk.ks = ks; pk.nvals = nvs; pk.G = part->n_groups; pk.stride = stride; pk.world = world; pk.rows = part->rows.as<long long>();
  for (int k = 0; k < nkeys; k++) { pk.key_vals[k] = part->key_vals[k].p; pk.key_null[k] = part->key_nulls[k].as<uint8_t>(); }
  for (int v = 0; v < nvs; v++) pk.states[v] = part->states[v].as<u64>();
  auto* res = new pdrs_groupby_result();
  res->ctx = c; res->nkeys = nkeys; res->nvals = nvs; res->naggs = naggs;
  for (int k = 0; k < nkeys; k++) res->key_dtype[k] = keys[k].dtype;
  struct Guard { pdrs_groupby_result* r; ~Guard() { delete r; } } guard{res};
  std::vector<int32_t> val_is_int(std::max(nvs, 1));
  for (int v = 0; v < nvs; v++) val_is_int[v] = nv[v].dtype == PDRS_I64;
  auto grow = [&](DevBuf& b, size_t bytes) -> int32_t { if (b.bytes < bytes) { b.release(); PDRS_TRY(b.alloc(c, bytes + bytes / 4)); } return PDRS_OK; };
  bool sharded = result_mode == 2;
  TableMem tm;
  std::vector<DevBuf> states(nvs);
  DistMerge mp{};
  mp.stride = stride; mp.nvals = nvs;
  auto make_table = [&](long long rows) -> int32_t {
    const long long slots = std::max<long long>(1024, pow2ceil(2 * rows + 16));
    PDRS_TRY(alloc_table(c, slots, NW, &tm));
    for (int v = 0; v < nvs; v++) { PDRS_TRY(states[v].alloc(c, (size_t)(slots + 1) * sizeof(GState), true)); mp.st[v] = states[v].as<GState>(); }
    mp.gt = tm.t;
    return PDRS_OK;
  };
  if (!sharded) {
    const long long cap = cm->groups_cap, block = 8 + cap * stride;
    pk.cap = cap;
    PDRS_TRY(grow(cm->send, (size_t)block * 8));
    PDRS_TRY(grow(cm->recv, (size_t)block * 8 * world));
    PDRS_TRY(dist_pack_launch<0>(c, NW, pk, cm->send.as<u64>(), nullptr, nullptr));
    pdrs_trace(c, "dist: pack");
    PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
    PDRS_TRY(pdrs_comm_allgather(cm, cm->send.p, cm->recv.p, (size_t)block * 8));
    PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
    cm->last_exchange_bytes = (int64_t)block * 8 * (world - 1);
    pdrs_trace(c, "dist: all-gather");
    // merge right away (same stream, no host round trip); the group counts are checked afterwards
    PDRS_TRY(make_table((long long)world * cap));
    mp.buf = cm->recv.as<u64>(); mp.nrows = (long long)world * cap; mp.cap = cap; mp.block = block;
    PDRS_TRY(dist_merge_launch(c, NW, mp));
    std::vector<u64> counts((size_t)world);
    PDRS_CUDA(c, cudaMemcpy2DAsync(counts.data(), 8, cm->recv.p, (size_t)block * 8, 8, (size_t)world, cudaMemcpyDeviceToHost, c->stream));
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    bool fits = true;
    pdrs_trace(c, "dist: table + merge + counts");
    for (int r = 0; r < world; r++) fits = fits && (long long)counts[r] <= cap;       // every rank sees the same counts: the decision is collective
    if (!fits) {
      if (result_mode == 1) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_groupby_agg_dist: a rank holds more than groups_cap = %lld groups; use the sharded result mode", cap);
      sharded = true;
      for (auto& s : states) s.release();
      tm = TableMem();
    }
  }
  if (sharded) {
    // rows by destination rank: count, offsets, scatter
    DevBuf cnt;
    PDRS_TRY(cnt.alloc(c, (size_t)(3 * world + 2) * 8, true));
    u64* dcount = cnt.as<u64>();
    u64* dcur = dcount + world;
    u64* doff = dcur + world;
    std::vector<u64> hc((size_t)world, 0), off((size_t)world + 1, 0);
    if (pk.G > 0) PDRS_TRY(dist_pack_launch<1>(c, NW, pk, nullptr, dcount, nullptr));
    PDRS_CUDA(c, cudaMemcpyAsync(hc.data(), dcount, (size_t)world * 8, cudaMemcpyDeviceToHost, c->stream));
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int r = 0; r < world; r++) off[r + 1] = off[r] + hc[r];
    PDRS_CUDA(c, cudaMemcpyAsync(doff, off.data(), (size_t)world * 8, cudaMemcpyHostToDevice, c->stream));
    PDRS_TRY(grow(cm->send, (size_t)std::max<long long>(pk.G, 1) * stride * 8));
    if (pk.G > 0) PDRS_TRY(dist_pack_launch<2>(c, NW, pk, cm->send.as<u64>(), dcur, doff));
    std::vector<u64> all((size_t)world * world);
    PDRS_TRY(pdrs_comm_allgather_host(cm, hc.data(), all.data(), (size_t)world * 8));     // all[src][dst]
    std::vector<size_t> soff(world), sb(world), roff(world), rb(world);
    size_t T = 0;
    int64_t to_peers = 0;
    for (int r = 0; r < world; r++) {
      soff[r] = (size_t)off[r] * stride * 8; sb[r] = (size_t)hc[r] * stride * 8;
      roff[r] = T * stride * 8; rb[r] = (size_t)all[(size_t)r * world + cm->rank] * stride * 8;
      T += (size_t)all[(size_t)r * world + cm->rank];
      if (r != cm->rank) to_peers += (int64_t)sb[r];
    }
    PDRS_TRY(grow(cm->recv, std::max<size_t>(T, 1) * stride * 8));
    PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
    PDRS_TRY(pdrs_comm_alltoallv(cm, cm->send.p, soff.data(), sb.data(), cm->recv.p, roff.data(), rb.data()));
    PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
    cm->last_exchange_bytes = to_peers;
    PDRS_TRY(make_table((long long)T));
    mp.buf = cm->recv.as<u64>(); mp.nrows = (long long)T; mp.cap = 0; mp.block = 0;
    PDRS_TRY(dist_merge_launch(c, NW, mp));
  }
  PDRS_TRY(finish_merged(c, ks, tm, states, res->key_dtype, nkeys, nvs, val_is_int.data(), ag.data(), naggs, res));
  pdrs_trace(c, "dist: finalise");
  PDRS_CUDA(c, cudaEventElapsedTime(&cm->last_exchange_ms, c->ev_a, c->ev_b));
  c->stats.main_kernel_ms = local_ms;
  c->stats.groupby_algo_used = local_algo;
  guard.r = nullptr;
  *out = res;
  return PDRS_OK;
}

extern "C" {

int64_t pdrs_groupby_n_groups(const pdrs_groupby_result* r) { return r ? r->n_groups : -1; }

static int32_t copy_out(const pdrs_groupby_result* r, void* dst, const DevBuf& src, size_t bytes) {
  pdrs_ctx* c = r->ctx;
  if (!dst) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "NULL output pointer");
  if (!src.p) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "result does not hold the requested array");
  if (bytes == 0) return PDRS_OK;
  PDRS_CUDA(c, cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}

int32_t pdrs_groupby_key(const pdrs_groupby_result* r, int32_t k, void* out_values, uint8_t* out_is_null) {
  if (!r || k < 0 || k >= r->nkeys) return PDRS_ERR_BAD_ARG;
  PDRS_TRY(copy_out(r, out_values, r->key_vals[k], (size_t)r->n_groups * key_out_bytes(r->key_dtype[k])));
  if (out_is_null) PDRS_TRY(copy_out(r, out_is_null, r->key_nulls[k], (size_t)r->n_groups));
  return PDRS_OK;
}
int32_t pdrs_groupby_agg_values(const pdrs_groupby_result* r, int32_t a, double* out) {
  if (!r || a < 0 || a >= r->naggs) return PDRS_ERR_BAD_ARG;
  return copy_out(r, out, r->aggs[a], (size_t)r->n_groups * 8);
}
int32_t pdrs_groupby_group_rows(const pdrs_groupby_result* r, int64_t* out) {
  if (!r) return PDRS_ERR_BAD_ARG;
  return copy_out(r, out, r->rows, (size_t)r->n_groups * 8);
}
int32_t pdrs_groupby_valid_n(const pdrs_groupby_result* r, int32_t v, int64_t* out) {
  if (!r || v < 0 || v >= r->nvals) return PDRS_ERR_BAD_ARG;
  return copy_out(r, out, r->validn[v], (size_t)r->n_groups * 8);
}
const void* pdrs_groupby_key_dev(const pdrs_groupby_result* r, int32_t k) { return (r && k >= 0 && k < r->nkeys) ? r->key_vals[k].p : nullptr; }
const uint8_t* pdrs_groupby_key_null_dev(const pdrs_groupby_result* r, int32_t k) { return (r && k >= 0 && k < r->nkeys) ? r->key_nulls[k].as<uint8_t>() : nullptr; }
const double* pdrs_groupby_agg_dev(const pdrs_groupby_result* r, int32_t a) { return (r && a >= 0 && a < r->naggs) ? r->aggs[a].as<double>() : nullptr; }
const int64_t* pdrs_groupby_group_rows_dev(const pdrs_groupby_result* r) { return r ? r->rows.as<int64_t>() : nullptr; }
const uint64_t* pdrs_groupby_states_dev(const pdrs_groupby_result* r, int32_t v) { return (r && v >= 0 && v < r->nvals) ? r->states[v].as<uint64_t>() : nullptr; }
void pdrs_groupby_result_free(pdrs_groupby_result* r) {
  if (!r) return;
  pdrs_ctx* c = r->ctx;
  cudaSetDevice(c->device);
  delete r;
  pdrs_settle_frees(c);
}

}  // extern "C"
