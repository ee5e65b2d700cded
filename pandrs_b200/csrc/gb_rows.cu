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
//   1. exact number of groups from the groupby kernels (count-only pdrs_groupby_agg)            -> table size
//   2. gr_assign_kernel   row -> slot of its key tuple in an open-addressing table (same key packing, same claim protocol as the
//                         global-table groupby kernel); slot counts by warp-aggregated atomics
//   3. gr_compact_kernel  slots -> dense group numbers, decoded key columns, group sizes; exclusive scan -> offsets
//   4. a STABLE least-significant-digit radix sort of the row numbers by group number, 8 bits per pass
//      (rs_hist_kernel / scan / rs_scatter_kernel: per-tile digit histograms, one global exclusive scan in digit-major order,
//      then every warp walks its rows in order and ranks them inside the warp with MATCH.ANY - no atomics, order preserved)
//      -> rows[offsets[g] .. offsets[g + 1]) = the rows of group g in ascending order, exactly the reference's Vec<usize>.
// Median sorts (ord(value), position) with the same passes (8 value digits, then the group digits) and reads the middle element(s).
#include <algorithm>

#include "groupby_kernels.cuh"
#include "gb_final.cuh"
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
struct AssignParams { KeySpec ks; long long n; GTable gt; uint32_t* slot_of_row; };

template <int NW>
__global__ void __launch_bounds__(256) gr_assign_kernel(const AssignParams p) {
  const int lane = threadIdx.x & 31;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i - lane < p.n; i += (long long)gridDim.x * blockDim.x) {
    const bool inb = i < p.n;
    u64 w[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) w[k] = 0;
    const bool knull = inb ? load_key_generic<NW>(p.ks, i, w) : false;
    long long gs = g_find_or_insert<NW>(p.gt, w, inb && !knull);
    if (knull) { gs = p.gt.slots; if (!(ld_cg_u64(&p.gt.hdr[gs].rowsw) & GB_FULL)) atomicOr(&p.gt.hdr[gs].rowsw, GB_FULL); }
    const bool ok = inb && gs >= 0;
    // group sizes: one atomic per distinct slot of the warp's 32 rows
    const unsigned long long tag = ok ? (unsigned long long)gs : ~0ull - (unsigned long long)lane;
    const uint32_t m = __match_any_sync(0xFFFFFFFFu, tag);
    if (ok) {
      if (lane == __ffs(m) - 1) atomicAdd(&p.gt.hdr[gs].rowsw, (u64)__popc(m));
      p.slot_of_row[i] = (uint32_t)gs;
    }
  }
}

struct CompactParams { FinParams fp; uint32_t* gid_of_slot; };
__global__ void gr_compact_kernel(const CompactParams p) {
  const GTable& gt = p.fp.gt;
  const long long total = gt.slots + 1;
  const int lane = threadIdx.x & 31;
  for (long long s0 = (long long)blockIdx.x * blockDim.x; s0 < total; s0 += (long long)gridDim.x * blockDim.x) {
    const long long s = s0 + threadIdx.x;
    u64 rw = 0;
    if (s < total) rw = gt.hdr[s].rowsw;
    const bool full = (rw & GB_FULL) != 0;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, full);
    if (!m) continue;
    u64 basepos = 0;
    if (lane == 0) basepos = atomicAdd(&gt.counters[CNT_OUT], (u64)__popc(m));
    basepos = __shfl_sync(0xFFFFFFFFu, basepos, 0);
    if (!full) continue;
    const long long o = (long long)(basepos + __popc(m & ((1u << lane) - 1u)));
    u64 w[PDRS_MAX_WORDS] = {0, 0, 0};
    const bool nullgroup = s == gt.slots;
    if (!nullgroup) {
      w[0] = gt.hdr[s].key0;
      if (p.fp.ks.nwords > 1) w[1] = gt.kw1[s];
      if (p.fp.ks.nwords > 2) w[2] = gt.kw2[s];
    }
    fin_write_group(p.fp, o, w, nullgroup, rw & GB_CNT_MASK, [](int) -> const GState* { return nullptr; });
    p.gid_of_slot[s] = (uint32_t)o;
  }
}
__global__ void gr_slot_to_gid_kernel(uint32_t* __restrict__ slot_of_row, long long n, const uint32_t* __restrict__ gid_of_slot) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) slot_of_row[i] = gid_of_slot[slot_of_row[i]];
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
  const int64_t n = keys[0].len;
  for (int k = 0; k < nkeys; k++) if (keys[k].len != n) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "key column %d has %lld rows, expected %lld", k, (long long)keys[k].len, (long long)n);
  if (n >= (1ll << 32) - 1) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_groupby_rows: at most 2^32 - 2 rows per call");
  auto* res = new pdrs_group_rows();
  res->ctx = c; res->nkeys = nkeys; res->n_rows = n;
  for (int k = 0; k < nkeys; k++) res->key_dtype[k] = keys[k].dtype;
  struct Guard { pdrs_group_rows* r; ~Guard() { delete r; } } guard{res};
  std::vector<ColView> kv(nkeys);
  for (int k = 0; k < nkeys; k++) PDRS_TRY(pdrs_view_col(c, &keys[k], &kv[k]));
  // 1. exact group count (the staged views are handed on as device columns: nothing is copied twice)
  long long G0 = 0;
  if (n > 0) {
    std::vector<pdrs_col> dk(nkeys);
    for (int k = 0; k < nkeys; k++) dk[k] = view_as_col(kv[k]);
    pdrs_groupby_result* cnt = nullptr;
    const int32_t saved = c->opts.compat_filter_nulls;
    c->opts.compat_filter_nulls = 0;
    const int32_t st = pdrs_groupby_agg(c, dk.data(), nkeys, nullptr, 0, nullptr, 0, nullptr, &cnt);
    c->opts.compat_filter_nulls = saved;
    PDRS_TRY(st);
    G0 = pdrs_groupby_n_groups(cnt);
    pdrs_groupby_result_free(cnt);
  }
  KeySpec ks;
  PDRS_TRY(pdrs_build_keyspec(c, kv.data(), nkeys, &ks));
  // 2. row -> slot
  const long long slots = std::max<long long>(1024, pow2ceil_ll(2 * G0 + 16));
  if (slots >= (1ll << 32) - 1) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_groupby_rows: too many groups (%lld)", G0);
  DevBuf hdr, kw1, kw2, counters, slot_of_row, gid_of_slot;
  PDRS_TRY(hdr.alloc(c, (size_t)(slots + 1) * sizeof(GHdr), true));
  if (ks.nwords > 1) PDRS_TRY(kw1.alloc(c, (size_t)(slots + 1) * 8));
  if (ks.nwords > 2) PDRS_TRY(kw2.alloc(c, (size_t)(slots + 1) * 8));
  PDRS_TRY(counters.alloc(c, CNT_N * 8, true));
  PDRS_TRY(slot_of_row.alloc(c, (size_t)std::max<int64_t>(n, 1) * 4));
  PDRS_TRY(gid_of_slot.alloc(c, (size_t)(slots + 1) * 4));
  GTable gt{};
  gt.hdr = hdr.as<GHdr>(); gt.kw1 = kw1.as<u64>(); gt.kw2 = kw2.as<u64>(); gt.st = nullptr;
  gt.mask = (u64)slots - 1; gt.shift = 64 - ceil_log2(slots); gt.slots = slots; gt.counters = counters.as<u64>();
  if (n > 0) {
    AssignParams ap{ks, n, gt, slot_of_row.as<uint32_t>()};
    const int g = pdrs_grid_for(c, n, 256);
    switch (ks.nwords) {
      case 1: gr_assign_kernel<1><<<g, 256, 0, c->stream>>>(ap); break;
      case 2: gr_assign_kernel<2><<<g, 256, 0, c->stream>>>(ap); break;
      default: gr_assign_kernel<3><<<g, 256, 0, c->stream>>>(ap); break;
    }
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, gt.counters, CNT_N * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + CNT_N, &gt.hdr[slots].rowsw, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->pinned_scalars[CNT_OVERFLOW] || c->pinned_scalars[CNT_SPIN_FAIL]) return pdrs_fail(c, PDRS_ERR_CUDA, "pdrs_groupby_rows: group table overflow");
  const int64_t G = c->pinned_scalars[CNT_NGROUPS] + (((u64)c->pinned_scalars[CNT_N] & GB_FULL) ? 1 : 0);
  res->n_groups = G;
  // 3. dense group numbers, keys, sizes, offsets
  const size_t Galloc = (size_t)std::max<int64_t>(G, 1);
  CompactParams cp{};
  cp.fp.gt = gt; cp.fp.ks = ks; cp.fp.nvals = 0; cp.fp.naggs = 0;
  for (int k = 0; k < nkeys; k++) {
    const int kb = (keys[k].dtype == PDRS_I64 || keys[k].dtype == PDRS_F64) ? 8 : (keys[k].dtype == PDRS_BOOL_BITS ? 1 : 4);
    PDRS_TRY(res->key_vals[k].alloc(c, Galloc * kb));
    PDRS_TRY(res->key_nulls[k].alloc(c, Galloc));
    cp.fp.key_out[k] = res->key_vals[k].p;
    cp.fp.key_null_out[k] = res->key_nulls[k].as<uint8_t>();
  }
  PDRS_TRY(res->sizes.alloc(c, Galloc * 8));
  PDRS_TRY(res->offsets.alloc(c, (Galloc + 1) * 8, true));
  PDRS_TRY(res->rows.alloc(c, (size_t)std::max<int64_t>(n, 1) * 8));
  cp.fp.rows_out = res->sizes.as<long long>();
  cp.gid_of_slot = gid_of_slot.as<uint32_t>();
  if (G > 0) {
    gr_compact_kernel<<<pdrs_grid_for(c, slots + 1, 256), 256, 0, c->stream>>>(cp);
    gr_slot_to_gid_kernel<<<pdrs_grid_for(c, n, 256), 256, 0, c->stream>>>(slot_of_row.as<uint32_t>(), n, gid_of_slot.as<uint32_t>());
    c->stats.kernel_launches += 2;
    PDRS_CUDA(c, cudaGetLastError());
    PDRS_TRY((scan_exclusive<long long, long long>(c, res->sizes.as<long long>(), G, res->offsets.as<long long>(), res->offsets.as<long long>() + G)));
    // 4. stable sort of the row numbers by group number
    DevBuf b0, b1;
    const int bits = ceil_log2(std::max<long long>(G, 2));
    if (bits > 8) { PDRS_TRY(b0.alloc(c, (size_t)n * 4)); PDRS_TRY(b1.alloc(c, (size_t)n * 4)); }
    PDRS_TRY((radix_sort_by_key<uint32_t>(c, slot_of_row.as<uint32_t>(), nullptr, n, bits, b0.as<uint32_t>(), b1.as<uint32_t>(), res->rows.as<long long>(), nullptr)));
  }
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
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
    const uint32_t* by_val = nullptr;
    PDRS_TRY((radix_sort_by_key<u64>(c, vkey.as<u64>(), nullptr, n, 64, b0.as<uint32_t>(), b1.as<uint32_t>(), nullptr, &by_val)));
    const uint32_t* order = nullptr;
    PDRS_TRY((radix_sort_by_key<uint32_t>(c, gid.as<uint32_t>(), by_val, n, ceil_log2(std::max<long long>(G, 2)), b0.as<uint32_t>(), b1.as<uint32_t>(), nullptr, &order)));
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
