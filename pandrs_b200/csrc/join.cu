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
//   join_build_kernel   one CAS claims a 16-byte slot {key, head}; rows with the same key are chained
//                       through next[] (head carries a MULTI flag so unique keys never touch next[])
//   join_probe_kernel   one pass over the left keys: stash[i] = head word of the matching slot (or
//                       NOMATCH / NULLKEY) and a per-CTA count of output rows
//   join_write_kernel   per-CTA base offsets from the scan of those counts, block-level exclusive
//                       scans inside each CTA's contiguous row range, ordered writes of the pairs
#include <algorithm>
#include <cstring>

#include "common.cuh"

typedef unsigned long long u64;

struct JSlot { u64 key; long long head; };
static constexpr long long J_EMPTY = -1, J_BUSY = -2, J_NOMATCH = -1, J_NULLKEY = -3;
static constexpr long long J_MULTI = 1ll << 62;
#define JOIN_THREADS 256
#define JOIN_ITEMS 4

struct JKeyCol { const void* data; const uint8_t* nulls; int dtype; };

__device__ __forceinline__ u64 jhash(u64 k) {
  u64 h = k * 0x9E3779B97F4A7C15ull;
  h ^= h >> 32;
  return h * 0xD6E8FEB86659FD93ull;
}

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

// One insertion attempt; never waits on a BUSY slot (the claimer may be a lane of the same warp): returns
// false to be called again after the warp has reconverged.
__device__ __forceinline__ bool jtry_insert(JSlot* tab, u64 mask, long long* next, u64 key, long long r, u64& slot, u64& probe, u64* fail) {
  while (probe <= mask) {
    ulonglong2 s = ld_volatile_v2(&tab[slot]);
    long long head = (long long)s.y;
    if (head == J_EMPTY) {
      long long old = (long long)atomicCAS(reinterpret_cast<u64*>(&tab[slot].head), (u64)J_EMPTY, (u64)J_BUSY);
      if (old == J_EMPTY) {
        tab[slot].key = key;
        next[r] = -1;
        __threadfence();
        atomicExch(reinterpret_cast<u64*>(&tab[slot].head), (u64)r);
        return true;
      }
      head = old;
      if (head != J_BUSY) s.x = ld_volatile_u64(&tab[slot].key);
    }
    if (head == J_BUSY) return false;
    if (s.x == key) {
      long long old = (long long)atomicExch(reinterpret_cast<u64*>(&tab[slot].head), (u64)(r | J_MULTI));
      next[r] = old & ~J_MULTI;
      return true;
    }
    slot = (slot + 1) & mask;
    probe++;
  }
  atomicAdd(fail, 1ull);
  return true;
}

__global__ void __launch_bounds__(256) join_build_kernel(JSlot* tab, u64 mask, long long* next, JKeyCol col, long long n, u64* fail) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r - lane < n; r += (long long)gridDim.x * blockDim.x) {
    u64 key = 0;
    bool pending = r < n && jload_key(col, r, &key);
    u64 slot = (jhash(key) >> 20) & mask, probe = 0;
    int rounds = 0;
    while (__any_sync(0xFFFFFFFFu, pending)) {      // warp-synchronous retry loop
      if (pending && jtry_insert(tab, mask, next, key, r, slot, probe, fail)) pending = false;
      if (++rounds > (1 << 22)) { if (pending) atomicAdd(fail, 1ull); break; }
    }
  }
}

__device__ __forceinline__ long long jprobe(const JSlot* tab, u64 mask, u64 key) {
  u64 slot = (jhash(key) >> 20) & mask;
  for (u64 probe = 0; probe <= mask; probe++) {
    ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2*>(&tab[slot]));
    if ((long long)s.y == J_EMPTY) return J_NOMATCH;
    if (s.x == key) return (long long)s.y;
    slot = (slot + 1) & mask;
  }
  return J_NOMATCH;
}

__device__ __forceinline__ long long jcount(long long stash, const long long* next, int left_join) {
  if (stash == J_NULLKEY) return 0;
  if (stash == J_NOMATCH) return left_join ? 1 : 0;
  if (!(stash & J_MULTI)) return 1;
  long long c = 0;
  for (long long r = stash & ~J_MULTI; r >= 0; r = __ldg(next + r)) c++;
  return c;
}

// rows [lo, hi) of CTA b: contiguous, so that the output stays left-row-major
__device__ __forceinline__ void cta_range(long long n, long long* lo, long long* hi) {
  const long long chunk = (long long)JOIN_THREADS * JOIN_ITEMS;
  const long long nchunks = (n + chunk - 1) / chunk;
  const long long per = (nchunks + gridDim.x - 1) / gridDim.x;
  *lo = min(n, (long long)blockIdx.x * per * chunk);
  *hi = min(n, *lo + per * chunk);
}

__global__ void __launch_bounds__(JOIN_THREADS) join_probe_kernel(const JSlot* __restrict__ tab, u64 mask, const long long* __restrict__ next, JKeyCol col, long long n,
                                                                  int left_join, long long* __restrict__ stash, u64* __restrict__ cta_counts) {
  __shared__ u64 sh_total;
  if (threadIdx.x == 0) sh_total = 0;
  __syncthreads();
  long long lo, hi;
  cta_range(n, &lo, &hi);
  u64 cnt = 0;
  for (long long i0 = lo + threadIdx.x; i0 < hi; i0 += (long long)JOIN_THREADS * JOIN_ITEMS) {
    u64 key[JOIN_ITEMS];
    bool ok[JOIN_ITEMS];
#pragma unroll
    for (int j = 0; j < JOIN_ITEMS; j++) { long long i = i0 + (long long)j * JOIN_THREADS; ok[j] = i < hi && jload_key(col, i, &key[j]); }
    long long st[JOIN_ITEMS];
#pragma unroll
    for (int j = 0; j < JOIN_ITEMS; j++) st[j] = ok[j] ? jprobe(tab, mask, key[j]) : J_NULLKEY;
#pragma unroll
    for (int j = 0; j < JOIN_ITEMS; j++) {
      long long i = i0 + (long long)j * JOIN_THREADS;
      if (i < hi) { __stcs(stash + i, st[j]); cnt += (u64)jcount(st[j], next, left_join); }
    }
  }
  for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&sh_total, cnt);
  __syncthreads();
  if (threadIdx.x == 0) cta_counts[blockIdx.x] = sh_total;
}

__global__ void join_scan_kernel(u64* v, int n, u64* total) {   // n <= a few thousand CTAs
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

__global__ void __launch_bounds__(JOIN_THREADS) join_write_kernel(const long long* __restrict__ stash, const long long* __restrict__ next, long long n, int left_join,
                                                                  const u64* __restrict__ cta_offsets, long long* __restrict__ out_l, long long* __restrict__ out_r) {
  __shared__ u64 wsum[JOIN_THREADS / 32];
  __shared__ u64 sh_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long lo, hi;
  cta_range(n, &lo, &hi);
  if (threadIdx.x == 0) sh_base = cta_offsets[blockIdx.x];
  __syncthreads();
  // thread t owns JOIN_ITEMS consecutive rows of every chunk, so positions are monotone in the row id
  for (long long c0 = lo; c0 < hi; c0 += (long long)JOIN_THREADS * JOIN_ITEMS) {
    long long st[JOIN_ITEMS];
    u64 cn[JOIN_ITEMS], mine = 0;
#pragma unroll
    for (int j = 0; j < JOIN_ITEMS; j++) {
      long long i = c0 + (long long)threadIdx.x * JOIN_ITEMS + j;
      st[j] = i < hi ? __ldcs(stash + i) : J_NULLKEY;
      cn[j] = (u64)jcount(st[j], next, left_join);
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
      if (st[j] == J_NOMATCH) { out_l[pos] = i; out_r[pos] = -1; pos++; }
      else if (!(st[j] & J_MULTI)) { out_l[pos] = i; out_r[pos] = st[j]; pos++; }
      else {
        const u64 p0 = pos;
        for (long long r = st[j] & ~J_MULTI; r >= 0; r = __ldg(next + r)) {   // chain order is arbitrary:
          u64 q = pos++;                                                     // insert in ascending right row (join.rs:158-161)
          while (q > p0 && out_r[q - 1] > r) { out_r[q] = out_r[q - 1]; q--; }
          out_r[q] = r;
        }
        for (u64 q = p0; q < pos; q++) out_l[q] = i;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) sh_base += total;
    __syncthreads();
  }
}

struct pdrs_join_result {
  pdrs_ctx* ctx = nullptr;
  int64_t n = 0;
  DevBuf left, right;
};

extern "C" {

int32_t pdrs_join_pairs(pdrs_ctx* c, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how, pdrs_join_result** out) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!left_key || !right_key || !out) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_join_pairs: NULL argument");
  if (how != PDRS_INNER && how != PDRS_LEFT) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "join type %d is not implemented (Inner and Left are)", how);
  if (left_key->dtype != right_key->dtype)   // join.rs:98-104
    return pdrs_fail(c, PDRS_ERR_TYPE_MISMATCH, "join key columns have different types (%d vs %d)", left_key->dtype, right_key->dtype);
  PDRS_CUDA(c, cudaSetDevice(c->device));
  if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
  c->stats.main_kernel_ms = 0; c->stats.total_ms = 0;
  ColView lv, rv;
  PDRS_TRY(pdrs_view_col(c, left_key, &lv));
  PDRS_TRY(pdrs_view_col(c, right_key, &rv));
  const int64_t nl = lv.len, nr = rv.len;
  auto* res = new pdrs_join_result();
  res->ctx = c;
  struct Guard { pdrs_join_result* r; ~Guard() { delete r; } } guard{res};

  long long slots = 1024;
  while (slots < 2 * nr) slots <<= 1;
  DevBuf tab, next, counts, fail;
  PDRS_TRY(tab.alloc(c, (size_t)slots * sizeof(JSlot)));
  PDRS_CUDA(c, cudaMemsetAsync(tab.p, 0xFF, (size_t)slots * sizeof(JSlot), c->stream));   // head = -1 (EMPTY)
  PDRS_TRY(next.alloc(c, (size_t)std::max<int64_t>(nr, 1) * 8));
  PDRS_TRY(fail.alloc(c, 8, true));
  c->stats.table_slots = slots;
  JKeyCol rc{rv.data, rv.nulls, rv.dtype}, lc{lv.data, lv.nulls, lv.dtype};
  if (nr > 0) {
    join_build_kernel<<<pdrs_grid_for(c, nr, 256), 256, 0, c->stream>>>(tab.as<JSlot>(), (u64)slots - 1, next.as<long long>(), rc, nr, fail.as<u64>());
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  const long long chunk = (long long)JOIN_THREADS * JOIN_ITEMS;
  int ctas = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 8, (nl + chunk - 1) / chunk));
  PDRS_TRY(counts.alloc(c, (size_t)(ctas + 4) * 8, true));
  DevBuf stash;
  PDRS_TRY(stash.alloc(c, (size_t)std::max<int64_t>(nl, 1) * 8));
  u64* cc = counts.as<u64>();
  if (nl > 0) {
    if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_a, c->stream));
    join_probe_kernel<<<ctas, JOIN_THREADS, 0, c->stream>>>(tab.as<JSlot>(), (u64)slots - 1, next.as<long long>(), lc, nl, how == PDRS_LEFT, stash.as<long long>(), cc);
    if (c->opt_timing) PDRS_CUDA(c, cudaEventRecord(c->ev_b, c->stream));
    join_scan_kernel<<<1, 1024, 0, c->stream>>>(cc, ctas, cc + ctas + 2);
    c->stats.kernel_launches += 2;
    PDRS_CUDA(c, cudaGetLastError());
  }
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, cc + ctas + 2, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars + 1, fail.p, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->pinned_scalars[1] != 0) return pdrs_fail(c, PDRS_ERR_CUDA, "join build: hash table insertion failed for %lld rows", (long long)c->pinned_scalars[1]);
  const int64_t M = nl > 0 ? c->pinned_scalars[0] : 0;
  res->n = M;
  PDRS_TRY(res->left.alloc(c, (size_t)std::max<int64_t>(M, 1) * 8));
  PDRS_TRY(res->right.alloc(c, (size_t)std::max<int64_t>(M, 1) * 8));
  if (M > 0) {
    join_write_kernel<<<ctas, JOIN_THREADS, 0, c->stream>>>(stash.as<long long>(), next.as<long long>(), nl, how == PDRS_LEFT, cc, res->left.as<long long>(), res->right.as<long long>());
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
  }
  if (c->opt_timing) {
    PDRS_CUDA(c, cudaEventRecord(c->ev_t1, c->stream));
    PDRS_CUDA(c, cudaEventSynchronize(c->ev_t1));
    PDRS_CUDA(c, cudaEventElapsedTime(&c->stats.total_ms, c->ev_t0, c->ev_t1));
    if (nl > 0) PDRS_CUDA(c, cudaEventElapsedTime(&c->stats.main_kernel_ms, c->ev_a, c->ev_b));
  } else {
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  guard.r = nullptr;
  *out = res;
  return PDRS_OK;
}

int64_t pdrs_join_len(const pdrs_join_result* r) { return r ? r->n : -1; }
int32_t pdrs_join_indices(const pdrs_join_result* r, int64_t* left_out, int64_t* right_out) {
  if (!r) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = r->ctx;
  if (r->n == 0) return PDRS_OK;
  if (!left_out || !right_out) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_join_indices: NULL output");
  PDRS_CUDA(c, cudaMemcpyAsync(left_out, r->left.p, (size_t)r->n * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaMemcpyAsync(right_out, r->right.p, (size_t)r->n * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}
const int64_t* pdrs_join_left_dev(const pdrs_join_result* r) { return r ? r->left.as<int64_t>() : nullptr; }
const int64_t* pdrs_join_right_dev(const pdrs_join_result* r) { return r ? r->right.as<int64_t>() : nullptr; }
void pdrs_join_result_free(pdrs_join_result* r) {
  if (!r) return;
  cudaSetDevice(r->ctx->device);
  delete r;
}

}  // extern "C"
