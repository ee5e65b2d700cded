// Columnar ingest in front of the path: Arrow validity bitmaps -> pandrs null masks, and dictionary
// encoding of Arrow string arrays on the device.
//
// Replaces
//   the per-element `arr.is_null(i)` loops of ArrowConverter::arrow_array_to_series          src/arrow_integration.rs:160-225
//   StringColumn::with_nulls -> GLOBAL_STRING_POOL.add_strings / StringPool::get_or_insert    src/column/string_column.rs:95-112,
//   (one RwLock-protected HashMap<Arc<str>, u32> lookup per row)                              src/column/string_pool.rs:28-52
//
// Dictionary encoding:
//   de_hash_kernel      row -> two independent 64-bit hashes of its bytes (a NULL row is the empty string) -> slot of the pair in an
//                       open-addressing table (claim protocol of the groupby's global table), atomicMin of the row number per slot
//   de_compact_kernel   occupied slots -> (first row, slot) pairs
//   radix sort          of the distinct strings by their first row (gb_sort.cuh) = the order in which a row loop over
//                       get_or_insert hands out ids
//   de_rank_kernel      slot -> id; rows -> ids
//   de_verify_kernel    BYTES of every row against the bytes of its id's first row: a hash collision fails the call instead of
//                       producing a wrong id
#include <algorithm>

#include "groupby_kernels.cuh"
#include "gb_sort.cuh"

struct pdrs_dict {
  pdrs_ctx* ctx = nullptr;
  int64_t len = 0, n_unique = 0;
  DevBuf ids;          // u32 [len]
  DevBuf first_rows;   // i64 [n_unique]
  DevBuf nulls;        // pandrs null mask, when the array had a validity bitmap
  bool has_nulls = false;
};

namespace {

__device__ __forceinline__ bool arrow_valid(const uint8_t* validity, long long bit) { return (validity[bit >> 3] >> (bit & 7)) & 1; }

// one thread per output byte (8 rows)
__global__ void validity_to_nulls_kernel(const uint8_t* __restrict__ validity, long long bit_offset, long long len, uint8_t* __restrict__ out, unsigned long long* __restrict__ n_nulls) {
  const long long nb = (len + 7) / 8;
  unsigned long long cnt = 0;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
    const long long rows = min(8ll, len - 8 * b);
    const uint32_t keep = (1u << rows) - 1u;
    uint32_t v = 0xFFu;
    if (validity) {
      const long long pos = bit_offset + 8 * b;
      const int sh = (int)(pos & 7);
      v = (uint32_t)validity[pos >> 3] >> sh;
      if (sh && rows > 8 - sh) v |= (uint32_t)validity[(pos >> 3) + 1] << (8 - sh);
    }
    const uint32_t nulls = ~v & keep;
    out[b] = (uint8_t)nulls;
    cnt += __popc(nulls);
  }
  for (int d = 16; d; d >>= 1) cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, d);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_nulls, cnt);
}

struct DictIn {
  const void* offsets; int off64;
  const uint8_t* bytes;
  const uint8_t* validity; long long bit_offset;
  long long n;
};
__device__ __forceinline__ void de_range(const DictIn& in, long long i, long long* lo, long long* hi) {
  if (in.validity && !arrow_valid(in.validity, in.bit_offset + i)) { *lo = 0; *hi = 0; return; }      // NULL = the empty string
  if (in.off64) { *lo = reinterpret_cast<const long long*>(in.offsets)[i]; *hi = reinterpret_cast<const long long*>(in.offsets)[i + 1]; }
  else { *lo = reinterpret_cast<const int*>(in.offsets)[i]; *hi = reinterpret_cast<const int*>(in.offsets)[i + 1]; }
}

// the offsets must describe the byte array before any kernel follows them: ascending, inside [0, nbytes]
__global__ void de_check_offsets_kernel(const DictIn in, long long nbytes, unsigned long long* __restrict__ bad) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < in.n; i += (long long)gridDim.x * blockDim.x) {
    long long lo, hi;
    if (in.off64) { lo = reinterpret_cast<const long long*>(in.offsets)[i]; hi = reinterpret_cast<const long long*>(in.offsets)[i + 1]; }
    else { lo = reinterpret_cast<const int*>(in.offsets)[i]; hi = reinterpret_cast<const int*>(in.offsets)[i + 1]; }
    if (lo < 0 || hi < lo || hi > nbytes) atomicAdd(bad, 1ull);
  }
}

struct HashParams { DictIn in; GTable gt; uint32_t* first; uint32_t* slot_of_row; };
__global__ void __launch_bounds__(256) de_hash_kernel(const HashParams p) {
  const int lane = threadIdx.x & 31;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i - lane < p.in.n; i += (long long)gridDim.x * blockDim.x) {
    const bool inb = i < p.in.n;
    u64 w[2] = {0, 0};
    if (inb) {
      long long lo, hi;
      de_range(p.in, i, &lo, &hi);
      u64 h1 = 0xCBF29CE484222325ull ^ (u64)(hi - lo), h2 = 0x9AE16A3B2F90404Full + (u64)(hi - lo) * 0xC2B2AE3D27D4EB4Full;
      for (long long b = lo; b < hi; b++) {
        const u64 ch = p.in.bytes[b];
        h1 = (h1 ^ ch) * 0x100000001B3ull;
        h2 = (h2 + ch + 1) * 0x9E3779B97F4A7C15ull;
        h2 ^= h2 >> 29;
      }
      w[0] = pdrs_mix64(h1);
      w[1] = pdrs_mix64(h2 ^ 0x5851F42D4C957F2Dull);
    }
    const long long gs = g_find_or_insert<2>(p.gt, w, inb);
    if (inb && gs >= 0) { atomicMin(&p.first[gs], (uint32_t)i); p.slot_of_row[i] = (uint32_t)gs; }
  }
}
__global__ void de_compact_kernel(const GTable gt, const uint32_t* __restrict__ first, uint32_t* __restrict__ cfirst, uint32_t* __restrict__ cslot) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < gt.slots; s += (long long)gridDim.x * blockDim.x) {
    if (gt.hdr[s].rowsw & GB_FULL) {
      const u64 o = atomicAdd(&gt.counters[CNT_OUT], 1ull);
      cfirst[o] = first[s];
      cslot[o] = (uint32_t)s;
    }
  }
}
__global__ void de_rank_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ cfirst, const uint32_t* __restrict__ cslot, long long D,
                               uint32_t* __restrict__ id_of_slot, long long* __restrict__ first_rows) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < D; r += (long long)gridDim.x * blockDim.x) {
    const uint32_t u = order[r];
    id_of_slot[cslot[u]] = (uint32_t)r;
    first_rows[r] = (long long)cfirst[u];
  }
}
__global__ void de_map_kernel(uint32_t* __restrict__ ids, long long n, const uint32_t* __restrict__ map) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) ids[i] = map[ids[i]];
}
__global__ void de_verify_kernel(const DictIn in, const uint32_t* __restrict__ ids, const long long* __restrict__ first_rows, unsigned long long* __restrict__ bad) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < in.n; i += (long long)gridDim.x * blockDim.x) {
    const long long f = first_rows[ids[i]];
    if (f == i) continue;
    long long lo, hi, flo, fhi;
    de_range(in, i, &lo, &hi);
    de_range(in, f, &flo, &fhi);
    bool same = hi - lo == fhi - flo && f < i;
    for (long long b = 0; same && b < hi - lo; b++) same = in.bytes[lo + b] == in.bytes[flo + b];
    if (!same) atomicAdd(bad, 1ull);
  }
}

int ilog2c(long long x) { int l = 0; while ((1ll << l) < x) l++; return l; }
long long p2c(long long x) { long long p = 1; while (p < x) p <<= 1; return p; }

// device copy of a caller buffer (host or device), padded by `pad` zero bytes
int32_t stage_bytes(pdrs_ctx* c, const void* src, int32_t mem, size_t bytes, size_t pad, DevBuf* own, const void** out) {
  if (mem == PDRS_MEM_DEVICE && pad == 0) { *out = src; return PDRS_OK; }
  PDRS_TRY(own->alloc(c, bytes + pad + 8, pad != 0));
  if (bytes) PDRS_CUDA(c, cudaMemcpyAsync(own->p, src, bytes, mem == PDRS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  *out = own->p;
  return PDRS_OK;
}

}  // namespace

extern "C" {

int32_t pdrs_arrow_validity_to_nulls(pdrs_ctx* c, const uint8_t* validity, int32_t validity_mem, int64_t bit_offset, int64_t len,
                                     uint8_t* out_null_bits, int32_t out_mem, int64_t* n_nulls_host) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (len < 0 || bit_offset < 0 || (!out_null_bits && len)) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_arrow_validity_to_nulls: bad argument");
  PDRS_CUDA(c, cudaSetDevice(c->device));
  if (n_nulls_host) *n_nulls_host = 0;
  if (len == 0) return PDRS_OK;
  const size_t nb = (size_t)(len + 7) / 8;
  DevBuf vown, oown, cnt;
  const void* vd = nullptr;
  if (validity) {
    const size_t first = (size_t)(bit_offset >> 3), vbytes = (size_t)((bit_offset & 7) + len + 7) / 8;
    PDRS_TRY(stage_bytes(c, validity + first, validity_mem, vbytes, validity_mem == PDRS_MEM_DEVICE ? 0 : 8, &vown, &vd));
    bit_offset &= 7;
  }
  uint8_t* od = out_null_bits;
  if (out_mem == PDRS_MEM_HOST) { PDRS_TRY(oown.alloc(c, nb)); od = oown.as<uint8_t>(); }
  PDRS_TRY(cnt.alloc(c, 8, true));
  validity_to_nulls_kernel<<<pdrs_grid_for(c, (int64_t)nb, 256), 256, 0, c->stream>>>((const uint8_t*)vd, bit_offset, len, od, cnt.as<unsigned long long>());
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  if (out_mem == PDRS_MEM_HOST) PDRS_CUDA(c, cudaMemcpyAsync(out_null_bits, od, nb, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, cnt.p, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  if (n_nulls_host) *n_nulls_host = c->pinned_scalars[0];
  return PDRS_OK;
}

int32_t pdrs_dict_encode(pdrs_ctx* c, const void* offsets, int32_t offsets_are_64, const uint8_t* bytes, int64_t nbytes,
                         const uint8_t* validity, int64_t bit_offset, int64_t len, int32_t mem, pdrs_dict** out) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!out || len < 0 || nbytes < 0 || bit_offset < 0 || (len && !offsets) || (nbytes && !bytes)) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_dict_encode: bad argument");
  if (len >= (1ll << 32) - 1) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_dict_encode: at most 2^32 - 2 rows per call");
  PDRS_CUDA(c, cudaSetDevice(c->device));
  auto* res = new pdrs_dict();
  res->ctx = c; res->len = len;
  struct Guard { pdrs_dict* r; ~Guard() { delete r; } } guard{res};
  PDRS_TRY(res->ids.alloc(c, (size_t)std::max<int64_t>(len, 1) * 4));
  if (len == 0) { PDRS_TRY(res->first_rows.alloc(c, 8)); guard.r = nullptr; *out = res; return PDRS_OK; }
  DevBuf o_own, b_own, v_own;
  DictIn in{};
  in.off64 = offsets_are_64 ? 1 : 0; in.n = len;
  PDRS_TRY(stage_bytes(c, offsets, mem, (size_t)(len + 1) * (offsets_are_64 ? 8 : 4), 0, &o_own, &in.offsets));
  const void* bd = nullptr;
  PDRS_TRY(stage_bytes(c, bytes, mem, (size_t)nbytes, 0, &b_own, &bd));
  in.bytes = (const uint8_t*)bd;
  if (validity) {
    const void* vd = nullptr;
    const size_t first = (size_t)(bit_offset >> 3), vbytes = (size_t)((bit_offset & 7) + len + 7) / 8;
    PDRS_TRY(stage_bytes(c, validity + first, mem, vbytes, mem == PDRS_MEM_DEVICE ? 0 : 8, &v_own, &vd));
    in.validity = (const uint8_t*)vd;
    in.bit_offset = bit_offset & 7;
    PDRS_TRY(res->nulls.alloc(c, (size_t)(len + 7) / 8 + 64, true));
    DevBuf cnt;
    PDRS_TRY(cnt.alloc(c, 8, true));
    validity_to_nulls_kernel<<<pdrs_grid_for(c, (len + 7) / 8, 256), 256, 0, c->stream>>>(in.validity, in.bit_offset, len, res->nulls.as<uint8_t>(), cnt.as<unsigned long long>());
    c->stats.kernel_launches++;
    res->has_nulls = true;
  }
  {
    DevBuf badoff;
    PDRS_TRY(badoff.alloc(c, 8, true));
    de_check_offsets_kernel<<<pdrs_grid_for(c, len, 256), 256, 0, c->stream>>>(in, nbytes, badoff.as<unsigned long long>());
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
    PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, badoff.p, 8, cudaMemcpyDeviceToHost, c->stream));
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->pinned_scalars[0]) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_dict_encode: %lld offsets are descending or outside the %lld value bytes", (long long)c->pinned_scalars[0], (long long)nbytes);
  }
  DevBuf slot_of_row;
  PDRS_TRY(slot_of_row.alloc(c, (size_t)len * 4));
  // table: first for at most ~n / 10 distinct strings, then for any number of them
  long long D = -1, slots = 0;
  DevBuf hdr, kw1, counters, first;
  GTable gt{};
  for (int attempt = 0; attempt < 2 && D < 0; attempt++) {
    slots = attempt == 0 ? p2c(len / 8 + 1024) : p2c(2 * len + 16);
    PDRS_TRY(hdr.alloc(c, (size_t)(slots + 1) * sizeof(GHdr), true));
    PDRS_TRY(kw1.alloc(c, (size_t)(slots + 1) * 8));
    PDRS_TRY(counters.alloc(c, CNT_N * 8, true));
    PDRS_TRY(first.alloc(c, (size_t)(slots + 1) * 4));
    PDRS_CUDA(c, cudaMemsetAsync(first.p, 0xFF, (size_t)(slots + 1) * 4, c->stream));
    gt = GTable{};
    gt.hdr = hdr.as<GHdr>(); gt.kw1 = kw1.as<u64>(); gt.kw2 = nullptr; gt.st = nullptr;
    gt.mask = (u64)slots - 1; gt.shift = 64 - ilog2c(slots); gt.slots = slots; gt.counters = counters.as<u64>();
    HashParams hp{in, gt, first.as<uint32_t>(), slot_of_row.as<uint32_t>()};
    de_hash_kernel<<<pdrs_grid_for(c, len, 256), 256, 0, c->stream>>>(hp);
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
    PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, gt.counters, CNT_N * 8, cudaMemcpyDeviceToHost, c->stream));
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->pinned_scalars[CNT_SPIN_FAIL]) return pdrs_fail(c, PDRS_ERR_CUDA, "pdrs_dict_encode: table contention");
    if (!c->pinned_scalars[CNT_OVERFLOW]) D = c->pinned_scalars[CNT_NGROUPS];
  }
  if (D < 0) return pdrs_fail(c, PDRS_ERR_CUDA, "pdrs_dict_encode: hash table overflow");
  res->n_unique = D;
  DevBuf cfirst, cslot, b0, b1, kb0, kb1, id_of_slot, bad;
  PDRS_TRY(cfirst.alloc(c, (size_t)D * 4));
  PDRS_TRY(cslot.alloc(c, (size_t)D * 4));
  PDRS_TRY(b0.alloc(c, (size_t)D * 4));
  PDRS_TRY(b1.alloc(c, (size_t)D * 4));
  PDRS_TRY(kb0.alloc(c, (size_t)D * 4));
  PDRS_TRY(kb1.alloc(c, (size_t)D * 4));
  PDRS_TRY(id_of_slot.alloc(c, (size_t)(slots + 1) * 4));
  PDRS_TRY(res->first_rows.alloc(c, (size_t)D * 8));
  PDRS_TRY(bad.alloc(c, 8, true));
  de_compact_kernel<<<pdrs_grid_for(c, slots, 256), 256, 0, c->stream>>>(gt, first.as<uint32_t>(), cfirst.as<uint32_t>(), cslot.as<uint32_t>());
  c->stats.kernel_launches++;
  const uint32_t* order = nullptr;
  PDRS_TRY((radix_sort_pairs<uint32_t>(c, cfirst.as<uint32_t>(), nullptr, D, ilog2c(std::max<long long>(len, 2)), kb0.as<uint32_t>(), kb1.as<uint32_t>(), b0.as<uint32_t>(), b1.as<uint32_t>(),
                                       nullptr, &order)));
  de_rank_kernel<<<pdrs_grid_for(c, D, 256), 256, 0, c->stream>>>(order, cfirst.as<uint32_t>(), cslot.as<uint32_t>(), D, id_of_slot.as<uint32_t>(), res->first_rows.as<long long>());
  PDRS_CUDA(c, cudaMemcpyAsync(res->ids.p, slot_of_row.p, (size_t)len * 4, cudaMemcpyDeviceToDevice, c->stream));
  de_map_kernel<<<pdrs_grid_for(c, len, 256), 256, 0, c->stream>>>(res->ids.as<uint32_t>(), len, id_of_slot.as<uint32_t>());
  de_verify_kernel<<<pdrs_grid_for(c, len, 256), 256, 0, c->stream>>>(in, res->ids.as<uint32_t>(), res->first_rows.as<long long>(), bad.as<unsigned long long>());
  c->stats.kernel_launches += 3;
  PDRS_CUDA(c, cudaGetLastError());
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, bad.p, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->pinned_scalars[0]) return pdrs_fail(c, PDRS_ERR_UNSUPPORTED, "pdrs_dict_encode: %lld rows collide with a different string under both hashes", (long long)c->pinned_scalars[0]);
  guard.r = nullptr;
  *out = res;
  return PDRS_OK;
}

int64_t pdrs_dict_n_unique(const pdrs_dict* d) { return d ? d->n_unique : -1; }
int32_t pdrs_dict_ids(const pdrs_dict* d, uint32_t* out_host) {
  if (!d || (!out_host && d->len)) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = d->ctx;
  if (d->len) PDRS_CUDA(c, cudaMemcpyAsync(out_host, d->ids.p, (size_t)d->len * 4, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}
const uint32_t* pdrs_dict_ids_dev(const pdrs_dict* d) { return d ? d->ids.as<uint32_t>() : nullptr; }
int32_t pdrs_dict_first_rows(const pdrs_dict* d, int64_t* out_host) {
  if (!d || (!out_host && d->n_unique)) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = d->ctx;
  if (d->n_unique) PDRS_CUDA(c, cudaMemcpyAsync(out_host, d->first_rows.p, (size_t)d->n_unique * 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}
const uint8_t* pdrs_dict_nulls_dev(const pdrs_dict* d) { return (d && d->has_nulls) ? d->nulls.as<uint8_t>() : nullptr; }
int32_t pdrs_dict_remap(pdrs_dict* d, const uint32_t* new_ids_host) {
  if (!d || (!new_ids_host && d->n_unique)) return PDRS_ERR_BAD_ARG;
  pdrs_ctx* c = d->ctx;
  PDRS_CUDA(c, cudaSetDevice(c->device));
  if (d->len == 0 || d->n_unique == 0) return PDRS_OK;
  DevBuf m;
  PDRS_TRY(m.alloc(c, (size_t)d->n_unique * 4));
  PDRS_CUDA(c, cudaMemcpyAsync(m.p, new_ids_host, (size_t)d->n_unique * 4, cudaMemcpyHostToDevice, c->stream));
  de_map_kernel<<<pdrs_grid_for(c, d->len, 256), 256, 0, c->stream>>>(d->ids.as<uint32_t>(), d->len, m.as<uint32_t>());
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}
void pdrs_dict_free(pdrs_dict* d) {
  if (!d) return;
  cudaSetDevice(d->ctx->device);
  delete d;
}

}  // extern "C"
