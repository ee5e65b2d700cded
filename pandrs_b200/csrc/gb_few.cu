// Very low cardinality (<= 16 groups), several value columns, sum / mean / count: ONE scan over the key column(s), the
// row filter and ALL value columns (BASELINE.json configs[4]: Q1-style filter -> groupby(returnflag, linestatus) with
// sums and means of five f64 columns).
//
// Replaces filter (data_ops.rs:37-121) + group_by_with_options (grouping.rs:38-115) + calculate_aggregation
// Sum / Mean / Count (aggregation.rs:507-530, 625-648, 743) / the Aggregate arm of LazyFrame::execute (lazy.rs:179-404).
//
// The other groupby kernels make one pass per value column (keys and filter are read again every time: 5 passes for
// Q1).  With a handful of groups nothing needs sorting or hashing:
//   * the group keys are known up front (the cardinality sample of the host has seen them); a row finds its group id
//     with a branch-free compare chain against <= 16 kernel parameters; a key the sample missed goes to the global table
//     row by row (same spill path as the other kernels),
//   * every LANE owns private accumulators in shared memory: acc[group][column][lane] - a row's update is a plain
//     LDS.64 / add / STS.64 on an address whose bank is the lane's own, so there are no conflicts and no atomics,
//   * a lane loads 2 consecutive rows of every column with one 128-bit streaming load; ~100 bytes are in flight per lane,
//     24 warps per SM.
// Algorithmic bytes per row = key bytes + 8 per value column + 1/8 per bitmap (+ 8 when a typed predicate is
// evaluated from its column instead of a precomputed Boolean column, lazy.rs:170-182).
#include <algorithm>

#include "groupby_kernels.cuh"
#include "gb_few.cuh"

namespace {

constexpr int GF_NT = 256, GF_ROWS = 2;                 // rows per lane and unit; a warp's unit = 64 consecutive rows (one 128-bit load per 8-byte column)
constexpr uint32_t GF_RMASK = (1u << GF_ROWS) - 1u;
constexpr int GF_UNIT = 32 * GF_ROWS;

__device__ __forceinline__ bool gf_pred(const GfParams& p, u64 bits) {
  if (p.pdtype == PDRS_F64) {
    const double x = __longlong_as_double((long long)bits), c = p.pfval;
    switch (p.pop) { case PDRS_CMP_LT: return x < c; case PDRS_CMP_LE: return x <= c; case PDRS_CMP_GT: return x > c; case PDRS_CMP_GE: return x >= c;
                     case PDRS_CMP_EQ: return x == c; default: return x != c; }
  }
  const long long x = (long long)bits, c = p.pival;
  switch (p.pop) { case PDRS_CMP_LT: return x < c; case PDRS_CMP_LE: return x <= c; case PDRS_CMP_GT: return x > c; case PDRS_CMP_GE: return x >= c;
                   case PDRS_CMP_EQ: return x == c; default: return x != c; }
}

// load_key_generic (groupby_kernels.cuh) restated for this kernel, fully inlined: the shared routine is __noinline__ and takes the
// KeySpec by address, which would make the compiler copy the whole 1 KB parameter block to LOCAL memory and read it back in the
// scan (measured: 160 B/row of extra L2 traffic, the kernel ran at a third of its speed).  One-word tuples, no NULL keys here.
__device__ __forceinline__ u64 gf_key_word(const KeySpec& ks, long long row) {
  u64 w = 0;
#pragma unroll
  for (int k = 0; k < PDRS_MAX_KEYS; k++) {
    if (k >= ks.nkeys) break;
    const KeyColDev& c = ks.c[k];
    u64 v = 0;
    switch (c.dtype) {
      case PDRS_I64: v = (u64)__ldg((const long long*)c.data + row) - (u64)c.offset; break;
      case PDRS_F64: { const double d = __ldg((const double*)c.data + row); v = (d != d) ? 0x7FF8000000000000ull : (u64)__double_as_longlong(d); break; }
      case PDRS_I32: v = ((u64)(long long)__ldg((const int*)c.data + row) - (u64)c.offset) & 0xFFFFFFFFull; break;
      case PDRS_DICT_U32: v = (u64)__ldg((const uint32_t*)c.data + row) - (u64)c.offset; break;
      default: v = pdrs_bit((const uint8_t*)c.data, row); break;
    }
    w |= v << c.shift;
  }
  return w;
}

// KM: 0 = one Int64 key column, 1 = one or two raw 4-byte key columns (i32 / dictionary ids, no NULLs), 2 = any key tuple
// that packs into one word (load_key_generic).  NV = number of value columns (compile time: the loads are unrolled).
template <int KM, int NV>
__global__ void __launch_bounds__(GF_NT) gb_few_kernel(const GfParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = GF_NT / 32;
  const int ng = p.ng, ne = ng * (NV > 0 ? NV : 1);
  // per warp: acc u64[ne][32], rows u32[ng][32], cnt u32[ne][32] (only when a value column has NULLs)
  const size_t warp_bytes = (size_t)ne * 256 + (size_t)ng * 128 + (p.any_vnull ? (size_t)ne * 128 : 0);
  unsigned char* wb = smem + (size_t)warp * warp_bytes;
  u64* acc = reinterpret_cast<u64*>(wb);
  uint32_t* rows = reinterpret_cast<uint32_t*>(wb + (size_t)ne * 256);
  uint32_t* cnt = rows + ng * 32;
  for (int i = lane; i < ne * 32; i += 32) { acc[i] = 0; if (p.any_vnull) cnt[i] = 0; }
  for (int i = lane; i < ng * 32; i += 32) rows[i] = 0;
  __syncwarp();
  const GbParams& b = p.base;
  const long long n = b.n, nunits = (n + GF_UNIT - 1) / GF_UNIT;
  const long long gwarp = (long long)blockIdx.x * nwarps + warp, twarps = (long long)gridDim.x * nwarps;
  const KeyColDev c0 = b.ks.c[0], c1 = b.ks.c[1];
#pragma unroll 1
  for (long long u = gwarp; u < nunits; u += twarps) {
    const long long r0 = u * GF_UNIT + GF_ROWS * lane;          // this lane's first row
    const bool full = (u + 1) * GF_UNIT <= n;
    u64 key[GF_ROWS], val[NV > 0 ? NV : 1][GF_ROWS], pv[GF_ROWS];
    uint32_t act = 0;                                            // bit q: row q takes part
    uint32_t vn[NV > 0 ? NV : 1];                                // bit q: value of column v is NULL in row q
    if (full) {
      // ---- loads: 128-bit, nothing consumed before everything is issued
      if (KM == 0) {
        const ulonglong2 a = ld_stream_v2(reinterpret_cast<const u64*>(c0.data) + r0);
        key[0] = a.x; key[1] = a.y;
      } else if (KM == 1) {
        const u64 a = __ldcs(reinterpret_cast<const u64*>(reinterpret_cast<const uint32_t*>(c0.data) + r0));
        const u64 c = b.ks.nkeys == 2 ? __ldcs(reinterpret_cast<const u64*>(reinterpret_cast<const uint32_t*>(c1.data) + r0)) : 0ull;
        key[0] = ((a & 0xFFFFFFFFull) << c0.shift) | (b.ks.nkeys == 2 ? ((c & 0xFFFFFFFFull) << c1.shift) : 0ull);
        key[1] = ((a >> 32) << c0.shift) | (b.ks.nkeys == 2 ? ((c >> 32) << c1.shift) : 0ull);
      } else {
#pragma unroll
        for (int q = 0; q < GF_ROWS; q++) key[q] = gf_key_word(b.ks, r0 + q);
      }
#pragma unroll
      for (int v = 0; v < NV; v++) {
        const ulonglong2 a = ld_stream_v2(reinterpret_cast<const u64*>(p.val[v]) + r0);
        val[v][0] = a.x; val[v][1] = a.y;
      }
      if (p.pcol) {
        const ulonglong2 a = ld_stream_v2(reinterpret_cast<const u64*>(p.pcol) + r0);
        pv[0] = a.x; pv[1] = a.y;
      }
      act = GF_RMASK;
    } else {
#pragma unroll
      for (int q = 0; q < GF_ROWS; q++) {
        const long long r = r0 + q;
        key[q] = 0; pv[q] = 0;
#pragma unroll
        for (int v = 0; v < NV; v++) val[v][q] = 0;
        if (r >= n) continue;
        act |= 1u << q;
        key[q] = gf_key_word(b.ks, r);
#pragma unroll
        for (int v = 0; v < NV; v++) val[v][q] = __ldg(reinterpret_cast<const u64*>(p.val[v]) + r);
        if (p.pcol) pv[q] = __ldg(reinterpret_cast<const u64*>(p.pcol) + r);
      }
    }
    // ---- bitmaps: the bits of this lane's rows sit in one 32-bit word (bitmaps cover ceil(n / 64) * 8 bytes)
    const long long bw = r0 >> 5;
    const int bs = (int)(r0 & 31);
    const bool inb = r0 < n;
    if (b.fbits) {     // filter keeps Some(true) rows only (data_ops.rs:49-55)
      uint32_t f = inb ? (__ldg(reinterpret_cast<const uint32_t*>(b.fbits) + bw) >> bs) & GF_RMASK : 0u;
      if (b.fnull && inb) f &= ~((__ldg(reinterpret_cast<const uint32_t*>(b.fnull) + bw) >> bs) & GF_RMASK);
      act &= f;
    }
    if (p.pcol) {      // typed predicate evaluated from its column; a NULL compares as "not kept" (Some(true) only)
      uint32_t f = 0;
#pragma unroll
      for (int q = 0; q < GF_ROWS; q++) if (gf_pred(p, pv[q])) f |= 1u << q;
      if (p.pnull && inb) f &= ~((__ldg(reinterpret_cast<const uint32_t*>(p.pnull) + bw) >> bs) & GF_RMASK);
      act &= f;
    }
#pragma unroll
    for (int v = 0; v < NV; v++) {
      vn[v] = (p.vnull[v] && inb) ? (__ldg(reinterpret_cast<const uint32_t*>(p.vnull[v]) + bw) >> bs) & GF_RMASK : 0u;
      if (b.compat_nulls && vn[v]) {      // filter + compat_filter_nulls: NULL values count as 0 (data_ops.rs:64-71)
#pragma unroll
        for (int q = 0; q < GF_ROWS; q++) if ((vn[v] >> q) & 1u) val[v][q] = 0;
        vn[v] = 0;
      }
    }
    // ---- group ids: compare chain against the known keys
    int gid[GF_ROWS];
    uint32_t spill = 0;
#pragma unroll
    for (int q = 0; q < GF_ROWS; q++) {
      int g = -1;
#pragma unroll
      for (int k = GF_MAXG - 1; k >= 0; k--) if (k < ng && key[q] == p.gkey[k]) g = k;
      gid[q] = g;
      if (((act >> q) & 1u) && g < 0) spill |= 1u << q;
    }
    // ---- lane-private accumulators
#pragma unroll
    for (int q = 0; q < GF_ROWS; q++) {
      if (!((act >> q) & 1u) || gid[q] < 0) continue;
      rows[gid[q] * 32 + lane]++;
#pragma unroll
      for (int v = 0; v < NV; v++) {
        if ((vn[v] >> q) & 1u) continue;
        const int e = (gid[q] * NV + v) * 32 + lane;
        const u64 a = acc[e];
        acc[e] = p.is_int[v] ? a + val[v][q] : (u64)__double_as_longlong(__longlong_as_double((long long)a) + __longlong_as_double((long long)val[v][q]));
        if (p.any_vnull) cnt[e]++;
      }
    }
    if (__any_sync(0xFFFFFFFFu, spill != 0)) {     // rare: a key the sample never saw -> global table, row by row
#pragma unroll
      for (int q = 0; q < GF_ROWS; q++) {          // (unrolled: a dynamic row index would push key[] / val[][] into local memory)
        const bool sp = (spill >> q) & 1u;
        u64 w[1] = {key[q]};
        const long long gs = g_find_or_insert<1>(b.gt, w, sp);
        if (!sp || gs < 0) continue;
        atomicAdd(&b.gt.hdr[gs].rowsw, 1ull);
        atomicAdd(&b.gt.counters[CNT_SPILLED], 1ull);
#pragma unroll
        for (int v = 0; v < NV; v++) {
          if ((vn[v] >> q) & 1u) continue;
          GState* s = &p.st[v][gs];
          atomicAdd(&s->n, 1ull);
          if (p.is_int[v]) atomicAdd(&s->isum, val[v][q]); else atomicAdd(&s->S1, __longlong_as_double((long long)val[v][q]));
        }
      }
    }
  }
  __syncwarp();
  // ---- flush: lane g of every warp sums the 32 lane-private copies of group g and adds one batch per (warp, group, column)
  {
    const int g = lane;
    const bool have_g = g < ng;
    u64 r = 0;
    if (have_g) for (int l = 0; l < 32; l++) r += rows[g * 32 + l];
    const bool have = have_g && r != 0;
    u64 w[1] = {have_g ? p.gkey[g] : 0ull};
    const long long gs = g_find_or_insert<1>(b.gt, w, have);
    if (have && gs >= 0) {
      atomicAdd(&b.gt.hdr[gs].rowsw, r);
#pragma unroll
      for (int v = 0; v < NV; v++) {
        u64 a = 0, nvld = 0;
        double s1 = 0.0;
        for (int l = 0; l < 32; l++) {
          const int e = (g * NV + v) * 32 + l;
          if (p.is_int[v]) a += acc[e]; else s1 += __longlong_as_double((long long)acc[e]);
          if (p.any_vnull) nvld += cnt[e];
        }
        if (!p.any_vnull) nvld = r;
        GTable t = b.gt;
        t.st = p.st[v];
        if (p.is_int[v]) g_update_batch<GB_SUM, true>(t, gs, 0ull, nvld, 0.0, false, 0.0, 0.0, a, 0ull, 0ull);
        else g_update_batch<GB_SUM, false>(t, gs, 0ull, nvld, 0.0, false, s1, 0.0, 0ull, 0ull, 0ull);
      }
    }
  }
}

template <int KM>
cudaError_t gf_launch_nv(const GfParams& p, int ctas, size_t smem, cudaStream_t s) {
  auto k = gb_few_kernel<KM, 0>;
  switch (p.nv) {
    case 0: k = gb_few_kernel<KM, 0>; break;
    case 1: k = gb_few_kernel<KM, 1>; break;
    case 2: k = gb_few_kernel<KM, 2>; break;
    case 3: k = gb_few_kernel<KM, 3>; break;
    case 4: k = gb_few_kernel<KM, 4>; break;
    case 5: k = gb_few_kernel<KM, 5>; break;
    case 6: k = gb_few_kernel<KM, 6>; break;
    case 7: k = gb_few_kernel<KM, 7>; break;
    default: k = gb_few_kernel<KM, 8>; break;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<ctas, GF_NT, smem, s>>>(p);
  return cudaGetLastError();
}

// the distinct keys of the cardinality sample (scratch table of gb_sample_kernel), at most GF_MAXG + 1 of them
__global__ void gf_collect_keys_kernel(const GHdr* __restrict__ hdr, long long slots, u64* __restrict__ out /* [0] count, [1..] keys */) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (long long)gridDim.x * blockDim.x) {
    if (!(hdr[s].rowsw & GB_FULL)) continue;
    const u64 at = atomicAdd(&out[0], 1ull);
    if (at < GF_MAXG + 1) out[1 + at] = hdr[s].key0;
  }
}

}  // namespace

size_t gb_few_smem(int ng, int nv, bool any_vnull) {
  const size_t ne = (size_t)ng * (nv > 0 ? nv : 1);
  return (size_t)(GF_NT / 32) * (ne * 256 + (size_t)ng * 128 + (any_vnull ? ne * 128 : 0));
}

cudaError_t gb_few_collect_keys(const GTable& sample, u64* out_dev, int ctas, cudaStream_t s) {
  gf_collect_keys_kernel<<<ctas, 256, 0, s>>>(sample.hdr, sample.slots, out_dev);
  return cudaGetLastError();
}

cudaError_t gb_few_launch(const GfParams& p, int ctas, size_t smem, cudaStream_t s) {
  const KeySpec& ks = p.base.ks;
  int km = 2;
  if (ks.nkeys == 1 && ks.c[0].dtype == PDRS_I64 && !ks.c[0].nulls && ks.c[0].offset == 0 && ks.c[0].shift == 0) km = 0;
  else if (ks.nkeys <= 2) {
    bool raw = true;
    for (int i = 0; i < ks.nkeys; i++) raw = raw && !ks.c[i].nulls && (ks.c[i].dtype == PDRS_I32 || ks.c[i].dtype == PDRS_DICT_U32) && ks.c[i].offset == 0 && ks.c[i].bits == 32 && ks.c[i].null_alias < 0;
    if (raw) km = 1;
  }
  switch (km) {
    case 0: return gf_launch_nv<0>(p, ctas, smem, s);
    case 1: return gf_launch_nv<1>(p, ctas, smem, s);
    default: return gf_launch_nv<2>(p, ctas, smem, s);
  }
}
