// Row lists per group: the grouping half of the groupby on its own, for the callers that need the ROWS of every group and not
// (only) an aggregate of them.
//
// Replaces, for typed key columns,
//   OptimizedDataFrame::par_groupby                      split_dataframe/group/grouping.rs:124-331   (row ids per group, ascending;
//                                                        the caller builds one sub-frame per group with filter_by_indices,
//                                                        data_ops.rs:124-211 = pdrs_gather over the permutation returned here)
//   the row lists GroupBy holds (`groups: HashMap<Vec<String>, Vec<usize>>`, grouping.rs:62-104) as far as the order-dependent
//   aggregates need them: AggregateOp::First / Last / Median, group/aggregation.rs:585-624, 703-742.
//
// Device algorithm (no CPU fallback):
//   1. the groups (keys, sizes) from the groupby kernels (count-only pdrs_groupby_agg); exclusive scan of the sizes -> offsets
//   2. gr_table_build_kernel  key words of the G groups -> a read-only open-addressing table {key -> group number}
//      gr_assign_kernel       row -> group number with plain cached loads (no atomics; the table of a few thousand groups is L1 / L2)
//   3. a STABLE least-significant-digit radix sort of (group number, row number), 8 bits per pass (gb_sort.cuh)
//      -> rows[offsets[g] .. offsets[g + 1]) = the rows of group g in ascending order, exactly the reference's Vec<usize>.
// Median sorts (ord(value), position) with the same passes (8 value digits, then the group digits) and reads the middle element(s).
#include <algorithm>

#include "groupby_kernels.cuh"
#include "gb_final.cuh"
#include "gb_pack.cuh"
#include "gb_sort.cuh"

int32_t pdrs_build_keyspec(pdrs_ctx* c, const ColView* kv, int nkeys, KeySpec* ks);   // groupby.cu

struct pdrs_group_rows {
  pdrs_ctx* ctx = nullptr;
  int64_t n_groups = 0, n_rows = 0;
  int nkeys = 0;
  int key_dtype[PDRS_MAX_KEYS] = {0, 0, 0, 0};
  DevBuf key_vals[PDRS_MAX_KEYS], key_nulls[PDRS_MAX_KEYS];
  DevBuf sizes;     // i64 [G]
  DevBuf offsets;   // i64 [G + 1]
  DevBuf rows;      // i64 [n_rows]
};

namespace {

// ---------------------------------------------------------------- row -> slot -> group number
// The groups are known before the rows are looked at (count-only groupby): group j = entry j of its key arrays.  A read-only
// open-addressing table {key words -> j} is built from those G entries; a row then finds its group number with plain cached loads -
// no atomics, no claim protocol, and for a few thousand groups the table lives in L1 / L2.
struct LookupTab { u64* k0; u64* k1; u64* k2; uint32_t* gid; u64 mask; int shift; uint32_t* null_gid; };   // gid: all ones = empty slot
static constexpr uint32_t GR_EMPTY = 0xFFFFFFFFu;

template <int NW>
__global__ void gr_table_build_kernel(const DistPack pk, const LookupTab t) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < pk.G; j += (long long)gridDim.x * blockDim.x) {
    u64 w[NW];
    if (dist_row_words<NW>(pk, j, w)) { *t.null_gid = (uint32_t)j; continue; }       // the NULL-key group of a single-key grouping
    u64 slot = key_hash<NW>(w) >> t.shift;
    for (;;) {       // the keys are distinct: a slot is claimed by one CAS on its group number, the key words are plain stores
      if (atomicCAS(&t.gid[slot], GR_EMPTY, (uint32_t)j) == GR_EMPTY) { t.k0[slot] = w[0]; if (NW > 1) t.k1[slot] = w[NW > 1 ? 1 : 0]; if (NW > 2) t.k2[slot] = w[NW > 2 ? 2 : 0]; break; }
      slot = (slot + 1) & t.mask;
    }
  }
}
struct AssignParams { KeySpec ks; long long n; LookupTab t; uint32_t* gid_of_row; unsigned long long* missing; };
template <int NW>
__global__ void __launch_bounds__(256) gr_assign_kernel(const AssignParams p) {
  const uint32_t null_gid = *p.t.null_gid;
  unsigned long long miss = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    u64 w[NW];
    uint32_t g = GR_EMPTY;
    if (load_key_inline<NW>(p.ks, i, w)) g = null_gid;       // (inlined packer: the parameter block stays in the constant bank, no local copy of the KeySpec)
    else {
      u64 slot = key_hash<NW>(w) >> p.t.shift;
      for (u64 probe = 0; probe <= p.t.mask; probe++) {
        const uint32_t cand = __ldg(p.t.gid + slot);
        if (cand == GR_EMPTY) break;
        bool match = __ldg(p.t.k0 + slot) == w[0];
        if (NW > 1) match = match && __ldg(p.t.k1 + slot) == w[NW > 1 ? 1 : 0];
        if (NW > 2) match = match && __ldg(p.t.k2 + slot) == w[NW > 2 ? 2 : 0];
        if (match) { g = cand; break; }
        slot = (slot + 1) & p.t.mask;
      }
    }
    if (g == GR_EMPTY) { miss++; g = 0; }
    p.gid_of_row[i] = g;
  }
  if (miss) atomicAdd(p.missing, miss);
}

// ---------------------------------------------------------------- First / Last / Median over the row lists
// aggregation.rs:605-624 / 723-742: the value of the group's first (last) row as f64; NULL there -> 0.0
template <typename VT>
__global__ void gr_first_last_kernel(const VT* __restrict__ val, const uint8_t* __restrict__ vnull, const long long* __restrict__ off, const long long* __restrict__ rows,
                                     long long G, int last, double* __restrict__ out) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (long long)gridDim.x * blockDim.x) {
    double r = 0.0;
    if (off[g + 1] > off[g]) {
      const long long row = rows[last ? off[g + 1] - 1 : off[g]];
      if (!(vnull && pdrs_bit(vnull, row))) r = (double)val[row];
    }
    out[g] = r;
  }
}
__global__ void gr_gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, long long n, uint32_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = src[idx[i]];
}
// position j of the grouped order: its group number and the order-preserving image of its value (NULL -> all ones: sorts last)
template <typename VT>
__global__ void gr_median_keys_kernel(const VT* __restrict__ val, const uint8_t* __restrict__ vnull, const long long* __restrict__ off, const long long* __restrict__ rows,
                                      long long G, long long n, u64* __restrict__ vkey, uint32_t* __restrict__ gid, unsigned long long* __restrict__ validn) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    long long lo = 0, hi = G;            // largest g with off[g] <= j
    while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (off[mid] <= j) lo = mid; else hi = mid; }
    const long long row = rows[j];
    u64 k = ~0ull;
    if (!(vnull && pdrs_bit(vnull, row))) {
      k = ValTraits<VT>::ord(val[row]);    // i64::MAX also maps to all ones and may interleave with the NULLs: see the pick kernel
      atomicAdd(&validn[lo], 1ull);
    }
    vkey[j] = k;
    gid[j] = (uint32_t)lo;
  }
}
// aggregation.rs:585-604 / 703-722: middle element of the sorted non-NULL values, or the mean of the two middle ones
template <typename VT>
__global__ void gr_median_pick_kernel(const VT* __restrict__ val, const uint8_t* __restrict__ vnull, const long long* __restrict__ off, const long long* __restrict__ rows,
                                      const uint32_t* __restrict__ order, const unsigned long long* __restrict__ validn, long long G, double* __restrict__ out) {
  // a NULL row among the first m sorted positions can only be tied with values whose key is all ones (i64::MAX): it stands for one
  auto pick = [&](long long pos) -> VT {
    const long long row = rows[order[pos]];
    if (vnull && pdrs_bit(vnull, row)) return ValTraits<VT>::from_bits(ValTraits<VT>::is_int ? 0x7FFFFFFFFFFFFFFFull : 0x7FFFFFFFFFFFFFFFull);
    return val[row];
  };
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (long long)gridDim.x * blockDim.x) {
    const long long m = (long long)validn[g];
    double r = 0.0;
    if (m > 0) {
      const long long mid = m / 2;
      const VT b = pick(off[g] + mid);
      if (m % 2 == 0) {
        const VT a = pick(off[g] + mid - 1);
        if (ValTraits<VT>::is_int) r = (double)(long long)((u64)a + (u64)b) / 2.0;     // (values[mid - 1] + values[mid]) as f64 / 2.0
        else r = ((double)a + (double)b) / 2.0;
      } else r = (double)b;
    }
    out[g] = r;
  }
}

int ceil_log2(long long x) { int l = 0; while ((1ll << l) < x) l++; return l; }
long long pow2ceil_ll(long long x) { long long p = 1; while (p < x) p <<= 1; return p; }

pdrs_col view_as_col(const ColView& v) {
  pdrs_col d{};
  d.dtype = v.dtype; d.mem = PDRS_MEM_DEVICE; d.data = v.data; d.null_bits = v.nulls; d.len = v.len; d.null_alias = v.null_alias;
  d.null_len = v.nulls ? ((v.len + 7) / 8 + 7) / 8 * 8 : 0;
  return d;
}

}  // namespace

extern "C" {

int32_t pdrs_groupby_rows(pdrs_ctx* c, const pdrs_col* keys, int32_t nkeys, pdrs_group_rows** out) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!out || !keys || nkeys < 1 || nkeys > PDRS_MAX_KEYS) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_groupby_rows: bad argument (nkeys %d)", nkeys);
  PDRS_CUDA(c, cudaSetDevice(c->device));
  pdrs_settle_frees(c);
  pdrs_trace(c, nullptr);
  const int64_t n = keys[0].len;
  for (int k = 0; k < nkeys; k++) if (keys[k].len != n) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "key column %d has %lld rows, expected %lld", k, (long long)keys[k].len, (long long)n);
  if (n >= (1ll << 32) - 1) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_groupby_rows: at most 2^32 - 2 rows per call");
  auto* res = new pdrs_group_rows();
  res->ctx = c; res->nkeys = nkeys; res->n_rows = n;
  for (int k = 0; k < nkeys; k++) res->key_dtype[k] = keys[k].dtype;
  struct Guard { pdrs_group_rows* r; ~Guard() { delete r; } } guard{res};
  std::vector<ColView> kv(nkeys);
  for (int k = 0; k < nkeys; k++) PDRS_TRY(pdrs_view_col(c, &keys[k], &kv[k]));
  // 1. the groups: keys and sizes from the count-only groupby (the staged views are handed on as device columns: nothing is copied
  //    twice); group j of the row lists = entry j of that result
  long long G = 0;
  pdrs_groupby_result* cnt = nullptr;
  if (n > 0) {
    std::vector<pdrs_col> dk(nkeys);
    for (int k = 0; k < nkeys; k++) dk[k] = view_as_col(kv[k]);
    const int32_t saved = c->opts.compat_filter_nulls;
    c->opts.compat_filter_nulls = 0;
    const int32_t st = pdrs_groupby_agg(c, dk.data(), nkeys, nullptr, 0, nullptr, 0, nullptr, &cnt);
    c->opts.compat_filter_nulls = saved;
    PDRS_TRY(st);
    G = pdrs_groupby_n_groups(cnt);
  }
  struct CntGuard { pdrs_groupby_result* r; ~CntGuard() { if (r) pdrs_groupby_result_free(r); } } cguard{cnt};
  pdrs_trace(c, "rows: views + group count");
  res->n_groups = G;
  const size_t Galloc = (size_t)std::max<long long>(G, 1);
  KeySpec ks;
  PDRS_TRY(pdrs_build_keyspec(c, kv.data(), nkeys, &ks));
  DistPack pk{};
  pk.ks = ks; pk.G = G;
  for (int k = 0; k < nkeys; k++) {
    const int kb = (keys[k].dtype == PDRS_I64 || keys[k].dtype == PDRS_F64) ? 8 : (keys[k].dtype == PDRS_BOOL_BITS ? 1 : 4);
    PDRS_TRY(res->key_vals[k].alloc(c, Galloc * kb));
    PDRS_TRY(res->key_nulls[k].alloc(c, Galloc));
    if (G > 0) {
      PDRS_CUDA(c, cudaMemcpyAsync(res->key_vals[k].p, pdrs_groupby_key_dev(cnt, k), (size_t)G * kb, cudaMemcpyDeviceToDevice, c->stream));
      PDRS_CUDA(c, cudaMemcpyAsync(res->key_nulls[k].p, pdrs_groupby_key_null_dev(cnt, k), (size_t)G, cudaMemcpyDeviceToDevice, c->stream));
    }
    pk.key_vals[k] = res->key_vals[k].p;
    pk.key_null[k] = res->key_nulls[k].as<uint8_t>();
  }
  PDRS_TRY(res->sizes.alloc(c, Galloc * 8));
  PDRS_TRY(res->offsets.alloc(c, (Galloc + 1) * 8, true));
  PDRS_TRY(res->rows.alloc(c, (size_t)std::max<int64_t>(n, 1) * 8));
  if (G > 0) {
    PDRS_CUDA(c, cudaMemcpyAsync(res->sizes.p, pdrs_groupby_group_rows_dev(cnt), (size_t)G * 8, cudaMemcpyDeviceToDevice, c->stream));
    PDRS_TRY((scan_exclusive<long long, long long>(c, res->sizes.as<long long>(), G, res->offsets.as<long long>(), res->offsets.as<long long>() + G)));
    // 2. key words -> group number, then row -> group number
    const long long slots = std::max<long long>(1024, pow2ceil_ll(2 * G + 16));
    if (slots >= (1ll << 32) - 1) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_groupby_rows: too many groups (%lld)", G);
    DevBuf k0, k1, k2, tg, misc, gid_of_row;
    PDRS_TRY(k0.alloc(c, (size_t)slots * 8));
    if (ks.nwords > 1) PDRS_TRY(k1.alloc(c, (size_t)slots * 8));
    if (ks.nwords > 2) PDRS_TRY(k2.alloc(c, (size_t)slots * 8));
    PDRS_TRY(tg.alloc(c, (size_t)slots * 4));
    PDRS_CUDA(c, cudaMemsetAsync(tg.p, 0xFF, (size_t)slots * 4, c->stream));
    PDRS_TRY(misc.alloc(c, 16, true));        // [0] the NULL-key group's number (u32), [1] rows whose key was not found (u64)
    PDRS_TRY(gid_of_row.alloc(c, (size_t)n * 4));
    LookupTab lt{k0.as<u64>(), k1.as<u64>(), k2.as<u64>(), tg.as<uint32_t>(), (u64)slots - 1, 64 - ceil_log2(slots), misc.as<uint32_t>()};
    AssignParams ap{ks, n, lt, gid_of_row.as<uint32_t>(), reinterpret_cast<unsigned long long*>(misc.as<u64>() + 1)};
    const int gb = pdrs_grid_for(c, G, 256), ga = pdrs_grid_for(c, n, 256);
    switch (ks.nwords) {
      case 1: gr_table_build_kernel<1><<<gb, 256, 0, c->stream>>>(pk, lt); gr_assign_kernel<1><<<ga, 256, 0, c->stream>>>(ap); break;
      case 2: gr_table_build_kernel<2><<<gb, 256, 0, c->stream>>>(pk, lt); gr_assign_kernel<2><<<ga, 256, 0, c->stream>>>(ap); break;
      default: gr_table_build_kernel<3><<<gb, 256, 0, c->stream>>>(pk, lt); gr_assign_kernel<3><<<ga, 256, 0, c->stream>>>(ap); break;
    }
    c->stats.kernel_launches += 2;
    PDRS_CUDA(c, cudaGetLastError());
    PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, misc.as<u64>() + 1, 8, cudaMemcpyDeviceToHost, c->stream));
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    pdrs_trace(c, "rows: offsets, table, row -> group");
    if (c->pinned_scalars[0] != 0) return pdrs_fail(c, PDRS_ERR_CUDA, "pdrs_groupby_rows: %lld rows carry a key the grouping did not report", (long long)c->pinned_scalars[0]);
    // 3. stable sort of the row numbers by group number
    DevBuf kb0, kb1, pb0, pb1;
    const int bits = ceil_log2(std::max<long long>(G, 2));
    if (bits > 8) { PDRS_TRY(kb0.alloc(c, (size_t)n * 4)); PDRS_TRY(kb1.alloc(c, (size_t)n * 4)); PDRS_TRY(pb0.alloc(c, (size_t)n * 4)); PDRS_TRY(pb1.alloc(c, (size_t)n * 4)); }
    PDRS_TRY((radix_sort_pairs<uint32_t>(c, gid_of_row.as<uint32_t>(), nullptr, n, bits, kb0.as<uint32_t>(), kb1.as<uint32_t>(), pb0.as<uint32_t>(), pb1.as<uint32_t>(),
                                         res->rows.as<long long>(), nullptr)));
  }
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  pdrs_trace(c, "rows: radix sort");
  guard.r = nullptr;
  *out = res;
  return PDRS_OK;
}

int64_t pdrs_group_rows_n_groups(const pdrs_group_rows* r) { return r ? r->n_groups : -1; }
int64_t pdrs_group_rows_n_rows(const pdrs_group_rows* r) { return r ? r->n_rows : -1; }

static int32_t gr_copy_out(const pdrs_group_rows* r, void* dst, const DevBuf& src, size_t bytes) {
  pdrs_ctx* c = r->ctx;
  if (!dst) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "NULL output pointer");
  if (bytes == 0) return PDRS_OK;
  PDRS_CUDA(c, cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}
int32_t pdrs_group_rows_key(const pdrs_group_rows* r, int32_t k, void* out_values, uint8_t* out_is_null) {
  if (!r || k < 0 || k >= r->nkeys) return PDRS_ERR_BAD_ARG;
  const int dt = r->key_dtype[k];
  const int kb = (dt == PDRS_I64 || dt == PDRS_F64) ? 8 : (dt == PDRS_BOOL_BITS ? 1 : 4);
  PDRS_TRY(gr_copy_out(r, out_values, r->key_vals[k], (size_t)r->n_groups * kb));
  if (out_is_null) PDRS_TRY(gr_copy_out(r, out_is_null, r->key_nulls[k], (size_t)r->n_groups));
  return PDRS_OK;
}
int32_t pdrs_group_rows_offsets(const pdrs_group_rows* r, int64_t* out) {
  if (!r) return PDRS_ERR_BAD_ARG;
  return gr_copy_out(r, out, r->offsets, (size_t)(r->n_groups + 1) * 8);
}
int32_t pdrs_group_rows_ids(const pdrs_group_rows* r, int64_t* out) {
  if (!r) return PDRS_ERR_BAD_ARG;
  return gr_copy_out(r, out, r->rows, (size_t)r->n_rows * 8);
}
const int64_t* pdrs_group_rows_offsets_dev(const pdrs_group_rows* r) { return r ? r->offsets.as<int64_t>() : nullptr; }
const int64_t* pdrs_group_rows_ids_dev(const pdrs_group_rows* r) { return r ? r->rows.as<int64_t>() : nullptr; }

int32_t pdrs_group_rows_agg(pdrs_group_rows* r, const pdrs_col* val, int32_t op, double* out_host) {
  if (!r || !val || !out_host) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = r->ctx;
  PDRS_CUDA(c, cudaSetDevice(c->device));
  if (op != PDRS_MEDIAN && op != PDRS_FIRST && op != PDRS_LAST) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_group_rows_agg: op %d is not Median / First / Last", op);
  if (val->dtype != PDRS_I64 && val->dtype != PDRS_F64)      // aggregation.rs:748-752
    return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "op %d is not supported on a column of dtype %d", op, val->dtype);
  if (val->len != r->n_rows) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "value column has %lld rows, expected %lld", (long long)val->len, (long long)r->n_rows);
  const long long G = r->n_groups, n = r->n_rows;
  if (G == 0) return PDRS_OK;
  ColView v;
  PDRS_TRY(pdrs_view_col(c, val, &v));
  DevBuf outd;
  PDRS_TRY(outd.alloc(c, (size_t)G * 8));
  const long long* off = r->offsets.as<long long>();
  const long long* rows = r->rows.as<long long>();
  const bool is_int = val->dtype == PDRS_I64;
  const int gg = pdrs_grid_for(c, G, 256);
  if (op == PDRS_FIRST || op == PDRS_LAST) {
    if (is_int) gr_first_last_kernel<long long><<<gg, 256, 0, c->stream>>>((const long long*)v.data, v.nulls, off, rows, G, op == PDRS_LAST, outd.as<double>());
    else gr_first_last_kernel<double><<<gg, 256, 0, c->stream>>>((const double*)v.data, v.nulls, off, rows, G, op == PDRS_LAST, outd.as<double>());
    c->stats.kernel_launches++;
  } else {
    DevBuf vkey, gid, validn, b0, b1;
    PDRS_TRY(vkey.alloc(c, (size_t)n * 8));
    PDRS_TRY(gid.alloc(c, (size_t)n * 4));
    PDRS_TRY(validn.alloc(c, (size_t)G * 8, true));
    PDRS_TRY(b0.alloc(c, (size_t)n * 4));
    PDRS_TRY(b1.alloc(c, (size_t)n * 4));
    const int gn = pdrs_grid_for(c, n, 256);
    if (is_int) gr_median_keys_kernel<long long><<<gn, 256, 0, c->stream>>>((const long long*)v.data, v.nulls, off, rows, G, n, vkey.as<u64>(), gid.as<uint32_t>(), validn.as<unsigned long long>());
    else gr_median_keys_kernel<double><<<gn, 256, 0, c->stream>>>((const double*)v.data, v.nulls, off, rows, G, n, vkey.as<u64>(), gid.as<uint32_t>(), validn.as<unsigned long long>());
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
    // sort by value (8 passes over the 64-bit images), then - stably - by group number
    DevBuf k0, k1, gk;
    PDRS_TRY(k0.alloc(c, (size_t)n * 8));
    PDRS_TRY(k1.alloc(c, (size_t)n * 8));
    const uint32_t* by_val = nullptr;
    PDRS_TRY((radix_sort_pairs<u64>(c, vkey.as<u64>(), nullptr, n, 64, k0.as<u64>(), k1.as<u64>(), b0.as<uint32_t>(), b1.as<uint32_t>(), nullptr, &by_val)));
    PDRS_TRY(gk.alloc(c, (size_t)n * 4));
    gr_gather_u32_kernel<<<gn, 256, 0, c->stream>>>(gid.as<uint32_t>(), by_val, n, gk.as<uint32_t>());
    c->stats.kernel_launches++;
    const uint32_t* order = nullptr;
    PDRS_TRY((radix_sort_pairs<uint32_t>(c, gk.as<uint32_t>(), by_val, n, ceil_log2(std::max<long long>(G, 2)), k0.as<uint32_t>(), k1.as<uint32_t>(), b0.as<uint32_t>(), b1.as<uint32_t>(),
                                         nullptr, &order)));
    if (is_int) gr_median_pick_kernel<long long><<<gg, 256, 0, c->stream>>>((const long long*)v.data, v.nulls, off, rows, order, validn.as<unsigned long long>(), G, outd.as<double>());
    else gr_median_pick_kernel<double><<<gg, 256, 0, c->stream>>>((const double*)v.data, v.nulls, off, rows, order, validn.as<unsigned long long>(), G, outd.as<double>());
    c->stats.kernel_launches++;
  }
  PDRS_CUDA(c, cudaGetLastError());
  PDRS_CUDA(c, cudaMemcpyAsync(out_host, outd.p, (size_t)G * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}

void pdrs_group_rows_free(pdrs_group_rows* r) {
  if (!r) return;
  pdrs_ctx* c = r->ctx;
  cudaSetDevice(c->device);
  delete r;
  pdrs_settle_frees(c);
}

}  // extern "C"
