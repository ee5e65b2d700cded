// Context, error reporting, device memory, column upload, gather/filter and synthetic generators.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>

#include "common.cuh"

static thread_local std::string g_create_err;

int32_t pdrs_fail(pdrs_ctx* ctx, int32_t code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_err = buf;
  return code;
}

void pdrs_trace(pdrs_ctx* c, const char* label) {
  if (!c || !c->opt_trace) return;
  cudaStreamSynchronize(c->stream);
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  const double t = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  if (label) fprintf(stderr, "[pdrs trace] %-34s %8.3f ms\n", label, t - c->trace_t0);
  c->trace_t0 = t;
}

int32_t pdrs_mem_available(pdrs_ctx* c, size_t* bytes) {
  size_t free_b = 0, total_b = 0;
  PDRS_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
    uint64_t reserved = 0, used = 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
      free_b += (size_t)(reserved - used);
  } else cudaGetLastError();
  *bytes = free_b;
  return PDRS_OK;
}

int32_t pdrs_copy_to_host(pdrs_ctx* c, void* dst_host, const void* src_dev, size_t bytes) {
  if (bytes == 0) return PDRS_OK;
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));           // the result is complete before other streams read it
  if (bytes >= (64ull << 20) && c->opt_stage_threads >= 0) {
    PDRS_TRY(pdrs_stage_copy_async(c, dst_host, src_dev, bytes, 1));
    PDRS_TRY(pdrs_stage_join(c, c->stream));
  } else PDRS_CUDA(c, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}

int pdrs_dtype_bytes(int32_t dtype) {
  switch (dtype) {
    case PDRS_I64: case PDRS_F64: return 8;
    case PDRS_DICT_U32: case PDRS_I32: return 4;
    default: return 0;  // BOOL_BITS is bit-packed
  }
}

int32_t DevBuf::alloc(pdrs_ctx* c, size_t n, bool zero) {
  release();
  ctx = c;
  bytes = n ? n : 1;
  PDRS_CUDA(c, cudaMallocAsync(&p, bytes, c->stream));
  if (zero) PDRS_CUDA(c, cudaMemsetAsync(p, 0, bytes, c->stream));
  return PDRS_OK;
}
void DevBuf::release() {
  if (p && ctx) { cudaFreeAsync(p, ctx->stream); if (bytes >= (256ull << 20)) ctx->big_free_pending = true; }
  p = nullptr;
  bytes = 0;
}

extern "C" {

int32_t pdrs_abi_version(void) { return PDRS_ABI_VERSION; }

int32_t pdrs_ctx_create(const pdrs_options* opts, pdrs_ctx** out) {
  if (!out) return pdrs_fail(nullptr, PDRS_ERR_BAD_ARG, "pdrs_ctx_create: out is NULL");
  pdrs_ctx* c = new pdrs_ctx();
  if (opts) c->opts = *opts;
  c->device = c->opts.device;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    delete c;
    return pdrs_fail(nullptr, PDRS_ERR_CUDA, "no CUDA device available (%s); libpandrs_b200 has no CPU fallback",
                     cudaGetErrorString(e));
  }
  const int dev = c->device;
  if (dev < 0 || dev >= ndev) { delete c; return pdrs_fail(nullptr, PDRS_ERR_BAD_ARG, "device %d out of range (0..%d)", dev, ndev - 1); }
  if ((e = cudaSetDevice(c->device)) != cudaSuccess) { delete c; return pdrs_fail(nullptr, PDRS_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e)); }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, c->device);
  if (prop.major != 10) { delete c; return pdrs_fail(nullptr, PDRS_ERR_UNSUPPORTED, "libpandrs_b200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor); }
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (c->opts.stream) { c->stream = (cudaStream_t)c->opts.stream; c->own_stream = false; }
  else { cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking); c->own_stream = true; }
  cudaEventCreate(&c->ev_a); cudaEventCreate(&c->ev_b);
  cudaEventCreate(&c->ev_t0); cudaEventCreate(&c->ev_t1);
  cudaMallocHost((void**)&c->pinned_scalars, 64 * sizeof(int64_t));
  // keep freed blocks cached in the stream-ordered pool instead of returning them to the driver
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = c;
  return PDRS_OK;
}

void pdrs_ctx_destroy(pdrs_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  pdrs_stage_destroy(c);
  if (c->flush_buf) cudaFree(c->flush_buf);
  if (c->pinned_scalars) cudaFreeHost(c->pinned_scalars);
  cudaEventDestroy(c->ev_a); cudaEventDestroy(c->ev_b);
  cudaEventDestroy(c->ev_t0); cudaEventDestroy(c->ev_t1);
  for (int s = 0; s < 2; s++) { if (c->aux_stream[s]) cudaStreamDestroy(c->aux_stream[s]); if (c->aux_done[s]) cudaEventDestroy(c->aux_done[s]); }
  if (c->aux_fork) cudaEventDestroy(c->aux_fork);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* pdrs_last_error(pdrs_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int32_t pdrs_sync(pdrs_ctx* c) {
  if (!c) return PDRS_ERR_BAD_ARG;
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}

int32_t pdrs_get_stats(pdrs_ctx* c, pdrs_stats* out) {
  if (!c || !out) return PDRS_ERR_BAD_ARG;
  *out = c->stats;
  return PDRS_OK;
}

int32_t pdrs_set_option(pdrs_ctx* c, const char* name, int64_t value) {
  if (!c || !name) return PDRS_ERR_BAD_ARG;
  if (!strcmp(name, "groupby_algo")) c->opts.groupby_algo = (int32_t)value;
  else if (!strcmp(name, "groups_hint")) c->opts.groups_hint = value;
  else if (!strcmp(name, "warps")) c->opt_warps = value;
  else if (!strcmp(name, "sample_rows")) c->opt_sample_rows = value;
  else if (!strcmp(name, "ctas_per_sm")) c->opt_ctas_per_sm = value;
  else if (!strcmp(name, "compat_filter_nulls")) c->opts.compat_filter_nulls = (int32_t)value;
  else if (!strcmp(name, "ng")) c->opt_ng = value;
  else if (!strcmp(name, "join_algo")) c->opt_join_algo = value;
  else if (!strcmp(name, "join_log_nb")) c->opt_join_log_nb = value;
  else if (!strcmp(name, "join_ctas_per_sm")) c->opt_join_ctas_per_sm = value;
  else if (!strcmp(name, "join_part")) c->opt_join_part = value;
  else if (!strcmp(name, "join_emit")) c->opt_join_emit = value;
  else if (!strcmp(name, "join_prefetch")) c->opt_join_prefetch = value;
  else if (!strcmp(name, "xjoin_mode")) c->opt_xjoin_mode = value;
  else if (!strcmp(name, "spillbuf")) c->opt_spillbuf = value;
  else if (!strcmp(name, "part_side")) c->opt_part_side = value;
  else if (!strcmp(name, "key_compress")) c->opt_key_compress = value;
  else if (!strcmp(name, "join_slots_mult")) c->opt_join_slots_mult = value;
  else if (!strcmp(name, "join_bucketwise")) c->opt_join_bucketwise = value;
  else if (!strcmp(name, "join_region_mb")) c->opt_join_region_mb = value;
  else if (!strcmp(name, "timing")) c->opt_timing = value;
  else if (!strcmp(name, "dense")) c->opt_dense = value;
  else if (!strcmp(name, "radix")) c->opt_radix = value;
  else if (!strcmp(name, "tsort")) c->opt_tsort = value;
  else if (!strcmp(name, "part")) c->opt_part = value;
  else if (!strcmp(name, "tsort_heavy")) c->opt_tsort_heavy = value;
  else if (!strcmp(name, "tsort_mid")) c->opt_tsort_mid = value;
  else if (!strcmp(name, "tsort_team")) c->opt_tsort_team = value;
  else if (!strcmp(name, "tsort_threads")) c->opt_tsort_threads = value;
  else if (!strcmp(name, "tsort_min_groups")) c->opt_tsort_min_groups = value;
  else if (!strcmp(name, "compat_empty_string_id")) c->opt_empty_string_id = value;
  else if (!strcmp(name, "few")) c->opt_few = value;
  else if (!strcmp(name, "part_direct")) c->opt_part_direct = value;
  else if (!strcmp(name, "part_hot")) c->opt_part_hot = value;
  else if (!strcmp(name, "part_hash")) c->opt_part_hash = value;
  else if (!strcmp(name, "stage_threads")) { if (c->stager && value != c->opt_stage_threads) pdrs_stage_destroy(c); c->opt_stage_threads = value; }
  else if (!strcmp(name, "trace")) c->opt_trace = value;
  else if (!strcmp(name, "xjoin_round_rows")) c->opt_xjoin_round_rows = value;
  else if (!strcmp(name, "stream_rows")) c->opt_stream_rows = value;
  else if (!strcmp(name, "stream_chunk_rows")) c->opt_stream_chunk_rows = value;
  else if (!strcmp(name, "stream_compact_rows")) c->opt_stream_compact_rows = value;
  else return pdrs_fail(c, PDRS_ERR_BAD_ARG, "unknown option '%s'", name);
  return PDRS_OK;
}

int32_t pdrs_dev_alloc(pdrs_ctx* c, int64_t bytes, void** out) {
  if (!c || !out || bytes < 0) return PDRS_ERR_BAD_ARG;
  PDRS_CUDA(c, cudaSetDevice(c->device));
  PDRS_CUDA(c, cudaMalloc(out, bytes ? bytes : 1));
  return PDRS_OK;
}
int32_t pdrs_dev_free(pdrs_ctx* c, void* p) {
  if (!c) return PDRS_ERR_BAD_ARG;
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  PDRS_CUDA(c, cudaFree(p));
  return PDRS_OK;
}
int32_t pdrs_host_alloc(pdrs_ctx* c, int64_t bytes, void** out) {
  if (!c || !out || bytes < 0) return PDRS_ERR_BAD_ARG;
  PDRS_CUDA(c, cudaMallocHost(out, bytes ? bytes : 1));
  return PDRS_OK;
}
int32_t pdrs_host_free(pdrs_ctx* c, void* p) {
  if (!c) return PDRS_ERR_BAD_ARG;
  PDRS_CUDA(c, cudaFreeHost(p));
  return PDRS_OK;
}
int32_t pdrs_memcpy(pdrs_ctx* c, void* dst, const void* src, int64_t bytes, int32_t kind) {
  if (!c || bytes < 0) return PDRS_ERR_BAD_ARG;
  cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  PDRS_CUDA(c, cudaMemcpyAsync(dst, src, bytes, k, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}

int32_t pdrs_flush_l2(pdrs_ctx* c) {
  if (!c) return PDRS_ERR_BAD_ARG;
  if (!c->flush_buf) {
    c->flush_bytes = 512ull << 20;  // 4x the 126 MB L2
    PDRS_CUDA(c, cudaMalloc(&c->flush_buf, c->flush_bytes));
  }
  PDRS_CUDA(c, cudaMemsetAsync(c->flush_buf, 0x5A, c->flush_bytes, c->stream));
  return PDRS_OK;
}

int32_t pdrs_timer_begin(pdrs_ctx* c) {
  if (!c) return PDRS_ERR_BAD_ARG;
  PDRS_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
  return PDRS_OK;
}
int32_t pdrs_timer_end(pdrs_ctx* c, float* ms) {
  if (!c || !ms) return PDRS_ERR_BAD_ARG;
  PDRS_CUDA(c, cudaEventRecord(c->ev_t1, c->stream));
  PDRS_CUDA(c, cudaEventSynchronize(c->ev_t1));
  PDRS_CUDA(c, cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
  return PDRS_OK;
}

// ---- column upload ----
static int64_t data_bytes(const pdrs_col* c) {
  if (c->dtype == PDRS_BOOL_BITS) return (c->len + 7) / 8;
  return c->len * (int64_t)pdrs_dtype_bytes(c->dtype);
}

int32_t pdrs_col_upload(pdrs_ctx* c, const pdrs_col* h, pdrs_col* d) {
  if (!c || !h || !d) return PDRS_ERR_BAD_ARG;
  if (h->len < 0 || h->dtype < 0 || h->dtype > PDRS_I32) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_col_upload: bad column");
  *d = *h;
  d->mem = PDRS_MEM_DEVICE;
  int64_t nb = data_bytes(h);
  int64_t padded = ((nb + 63) / 64) * 64 + 64;
  void* dd = nullptr;
  PDRS_CUDA(c, cudaMalloc(&dd, padded));
  PDRS_CUDA(c, cudaMemsetAsync((char*)dd + (padded - 128 > 0 ? padded - 128 : 0), 0, padded > 128 ? 128 : padded, c->stream));
  cudaMemcpyKind k = h->mem == PDRS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (nb) PDRS_CUDA(c, cudaMemcpyAsync(dd, h->data, nb, k, c->stream));
  d->data = dd;
  d->null_bits = nullptr;
  d->null_len = 0;
  if (h->null_bits) {
    int64_t need = (h->len + 7) / 8;
    int64_t pad = ((need + 63) / 64) * 64 + 64;
    void* nn = nullptr;
    PDRS_CUDA(c, cudaMalloc(&nn, pad));
    PDRS_CUDA(c, cudaMemsetAsync(nn, 0, pad, c->stream));   // short masks: missing bytes mean "not NULL"
    int64_t have = h->null_len < need ? h->null_len : need;
    if (have > 0) PDRS_CUDA(c, cudaMemcpyAsync(nn, h->null_bits, have, k, c->stream));
    d->null_bits = (const uint8_t*)nn;
    d->null_len = pad;
  }
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}

int32_t pdrs_col_free(pdrs_ctx* c, pdrs_col* d) {
  if (!c || !d) return PDRS_ERR_BAD_ARG;
  if (d->mem != PDRS_MEM_DEVICE) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "pdrs_col_free: not a device column");
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  if (d->data) cudaFree((void*)d->data);
  if (d->null_bits) cudaFree((void*)d->null_bits);
  d->data = nullptr;
  d->null_bits = nullptr;
  return PDRS_OK;
}

}  // extern "C"

// Borrow a device column as-is, or stage a host column in stream-ordered scratch for this call.
int32_t pdrs_view_col(pdrs_ctx* c, const pdrs_col* col, ColView* v) {
  if (!col) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "NULL column");
  if (col->len < 0 || col->dtype < 0 || col->dtype > PDRS_I32) return pdrs_fail(c, PDRS_ERR_BAD_ARG, "bad column (dtype %d, len %lld)", col->dtype, (long long)col->len);
  v->dtype = col->dtype;
  v->len = col->len;
  v->null_alias = col->dtype == PDRS_DICT_U32 ? col->null_alias : -1;
  int64_t need = (col->len + 7) / 8;
  const int64_t need8 = (need + 7) / 8 * 8;
  if (col->mem == PDRS_MEM_DEVICE) {
    // the kernels use 128-bit loads on data and 64-bit loads on bitmaps: borrow when the caller's
    // buffers allow it, otherwise stage an aligned / padded copy (device to device)
    int64_t nbd = col->dtype == PDRS_BOOL_BITS ? need : col->len * (int64_t)pdrs_dtype_bytes(col->dtype);
    const bool bits_col = col->dtype == PDRS_BOOL_BITS;
    if (((uintptr_t)col->data & 15) || (bits_col && (nbd % 8))) {
      PDRS_TRY(v->own_data.alloc(c, (size_t)nbd + 64, bits_col));
      if (nbd) PDRS_CUDA(c, cudaMemcpyAsync(v->own_data.p, col->data, nbd, cudaMemcpyDeviceToDevice, c->stream));
      v->data = v->own_data.p;
    } else {
      v->data = col->data;
    }
    v->nulls = col->null_bits;
    if (col->null_bits && (((uintptr_t)col->null_bits & 7) || col->null_len < need8)) {
      PDRS_TRY(v->own_nulls.alloc(c, (size_t)need8 + 64, true));
      int64_t have = col->null_len < need ? col->null_len : need;
      if (have > 0) PDRS_CUDA(c, cudaMemcpyAsync(v->own_nulls.p, col->null_bits, have, cudaMemcpyDeviceToDevice, c->stream));
      v->nulls = (const uint8_t*)v->own_nulls.p;
    }
    return PDRS_OK;
  }
  int64_t nb = col->dtype == PDRS_BOOL_BITS ? need : col->len * (int64_t)pdrs_dtype_bytes(col->dtype);
  PDRS_TRY(v->own_data.alloc(c, (size_t)nb + 64, col->dtype == PDRS_BOOL_BITS));
  if (nb >= (64ll << 20) && c->opt_stage_threads >= 0) {
    // large host column: through the staging engine (stage.cu) - a pageable source moves at PCIe speed instead of the driver's
    // ~11 GB/s bounce path.  The stream-ordered allocation must exist before other streams write into it.
    PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
    PDRS_TRY(pdrs_stage_copy_async(c, v->own_data.p, col->data, (size_t)nb));
    PDRS_TRY(pdrs_stage_join(c, c->stream));
  } else if (nb) PDRS_CUDA(c, cudaMemcpyAsync(v->own_data.p, col->data, nb, cudaMemcpyHostToDevice, c->stream));
  v->data = v->own_data.p;
  if (col->null_bits) {
    PDRS_TRY(v->own_nulls.alloc(c, (size_t)need + 64, true));
    int64_t have = col->null_len < need ? col->null_len : need;
    if (have > 0) PDRS_CUDA(c, cudaMemcpyAsync(v->own_nulls.p, col->null_bits, have, cudaMemcpyHostToDevice, c->stream));
    v->nulls = (const uint8_t*)v->own_nulls.p;
  }
  return PDRS_OK;
}

// ---- synthetic generators (same arithmetic as oracle/pandrs_oracle.cpp) ----
__global__ void synth_keys_kernel(int64_t* out, int64_t n, int64_t row0, uint64_t seed, uint64_t card, int scramble) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t k = pdrs_mix64(seed * 0x100000001B3ULL + (uint64_t)(row0 + i)) % card;
    out[i] = (int64_t)(scramble ? pdrs_mix64(k ^ 0xA5A5A5A5DEADBEEFULL) : k);
  }
}
__global__ void synth_vals_kernel(double* out, int64_t n, int64_t row0, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t r = pdrs_mix64((seed + 1) * 0x100000001B3ULL + (uint64_t)(row0 + i));
    out[i] = (double)(r >> 11) * (1000.0 / 9007199254740992.0);
  }
}
__global__ void synth_nulls_kernel(uint8_t* out, int64_t n, int64_t row0, uint64_t seed, uint32_t per_million) {
  int64_t nb = (n + 7) / 8;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
    uint32_t byte = 0;
    for (int j = 0; j < 8 && b * 8 + j < n; j++) {
      uint64_t r = pdrs_mix64((seed + 2) * 0x100000001B3ULL + (uint64_t)(row0 + b * 8 + j));
      if ((uint32_t)(r % 1000000ULL) < per_million) byte |= 1u << j;
    }
    out[b] = (uint8_t)byte;
  }
}
// Join keys: unique -> a bijective mix of the row number (build side); else a uniform draw from
// [0, domain) pushed through the same mix (probe side), SURVEY.md §8d C3.
__global__ void synth_join_keys_kernel(int64_t* out, int64_t n, int64_t row0, uint64_t seed, uint64_t domain, int unique) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t id = unique ? (uint64_t)(row0 + i) : pdrs_mix64((seed + 3) * 0x100000001B3ULL + (uint64_t)(row0 + i)) % domain;
    out[i] = (int64_t)(id * 0x9E3779B97F4A7C15ULL);   // odd multiplier: bijective mod 2^64
  }
}

// ---- gather / filter ----
template <typename T>
__global__ void gather_kernel(const T* __restrict__ src, const uint8_t* __restrict__ nulls, const int64_t* __restrict__ idx, int64_t n, T* __restrict__ out, T dflt) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = idx[j];
    T v = dflt;
    if (i >= 0 && !(nulls && pdrs_bit(nulls, i))) v = src[i];
    out[j] = v;
  }
}
__global__ void gather_bits_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ nulls, const int64_t* __restrict__ idx, int64_t n, uint8_t* __restrict__ out) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = idx[j];
    out[j] = (i >= 0 && !(nulls && pdrs_bit(nulls, i))) ? (uint8_t)pdrs_bit(src, i) : 0;
  }
}

// Two-pass compaction of "Some(true)" rows: per-CTA counts -> exclusive scan (single CTA) -> ordered write.
__global__ void filter_count_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ nulls, int64_t nwords, int64_t n, unsigned long long* __restrict__ cta_counts) {
  __shared__ unsigned long long sh;
  if (threadIdx.x == 0) sh = 0;
  __syncthreads();
  int64_t per = (nwords + gridDim.x - 1) / gridDim.x;
  int64_t lo = blockIdx.x * per, hi = min(nwords, lo + per);
  unsigned long long c = 0;
  for (int64_t w = lo + threadIdx.x; w < hi; w += blockDim.x) {
    uint32_t m = bits[w] & (nulls ? ~nulls[w] : 0xFFFFFFFFu);
    int64_t rem = n - w * 32;
    if (rem < 32) m &= (rem <= 0) ? 0u : ((1u << rem) - 1u);
    c += __popc(m);
  }
  atomicAdd(&sh, c);
  __syncthreads();
  if (threadIdx.x == 0) cta_counts[blockIdx.x] = sh;
}
__global__ void scan_small_kernel(unsigned long long* v, int n, unsigned long long* total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long s = 0;
    for (int i = 0; i < n; i++) { unsigned long long t = v[i]; v[i] = s; s += t; }
    *total = s;
  }
}
__global__ void filter_write_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ nulls, int64_t nwords, int64_t n, const unsigned long long* __restrict__ cta_offsets, int64_t* __restrict__ out) {
  // one warp walks the CTA's word range in order so that the output stays ascending
  __shared__ unsigned long long base;
  if (threadIdx.x == 0) base = cta_offsets[blockIdx.x];
  __syncthreads();
  if (threadIdx.x >= 32) return;
  int lane = threadIdx.x;
  int64_t per = (nwords + gridDim.x - 1) / gridDim.x;
  int64_t lo = blockIdx.x * per, hi = min(nwords, lo + per);
  unsigned long long off = base;
  for (int64_t w0 = lo; w0 < hi; w0 += 32) {
    int64_t w = w0 + lane;
    uint32_t m = 0;
    if (w < hi) {
      m = bits[w] & (nulls ? ~nulls[w] : 0xFFFFFFFFu);
      int64_t rem = n - w * 32;
      if (rem < 32) m &= (rem <= 0) ? 0u : ((1u << rem) - 1u);
    }
    int c = __popc(m), incl = c;
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    unsigned long long o = off + (unsigned long long)(incl - c);
    while (m) { int b = __ffs(m) - 1; m &= m - 1; out[o++] = w * 32 + b; }
    off += (unsigned long long)__shfl_sync(0xFFFFFFFFu, incl, 31);
  }
}

int pdrs_grid_for(pdrs_ctx* c, int64_t n, int threads) {
  int64_t blocks = (n + threads - 1) / threads;
  int64_t cap = (int64_t)c->sm_count * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

extern "C" {

int32_t pdrs_synth_keys(pdrs_ctx* c, int64_t* out, int64_t n, int64_t row0, uint64_t seed, uint64_t card, int32_t scramble) {
  if (!c || !out || n < 0 || card == 0) return PDRS_ERR_BAD_ARG;
  synth_keys_kernel<<<pdrs_grid_for(c, n, 256), 256, 0, c->stream>>>(out, n, row0, seed, card, scramble);
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  return PDRS_OK;
}
int32_t pdrs_synth_vals(pdrs_ctx* c, double* out, int64_t n, int64_t row0, uint64_t seed) {
  if (!c || !out || n < 0) return PDRS_ERR_BAD_ARG;
  synth_vals_kernel<<<pdrs_grid_for(c, n, 256), 256, 0, c->stream>>>(out, n, row0, seed);
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  return PDRS_OK;
}
int32_t pdrs_synth_nulls(pdrs_ctx* c, uint8_t* out, int64_t n, int64_t row0, uint64_t seed, uint32_t per_million) {
  if (!c || !out || n < 0) return PDRS_ERR_BAD_ARG;
  synth_nulls_kernel<<<pdrs_grid_for(c, (n + 7) / 8, 256), 256, 0, c->stream>>>(out, n, row0, seed, per_million);
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  return PDRS_OK;
}
int32_t pdrs_synth_join_keys(pdrs_ctx* c, int64_t* out, int64_t n, int64_t row0, uint64_t seed, uint64_t domain, int32_t unique) {
  if (!c || !out || n < 0 || (!unique && domain == 0)) return PDRS_ERR_BAD_ARG;
  synth_join_keys_kernel<<<pdrs_grid_for(c, n, 256), 256, 0, c->stream>>>(out, n, row0, seed, domain, unique);
  c->stats.kernel_launches++;
  PDRS_CUDA(c, cudaGetLastError());
  return PDRS_OK;
}

int32_t pdrs_gather(pdrs_ctx* c, const pdrs_col* col, const int64_t* idx, int32_t idx_mem, int64_t n, void* out, int32_t out_mem) {
  if (!c || !col || (!idx && n) || (!out && n) || n < 0) return PDRS_ERR_BAD_ARG;
  ColView v;
  PDRS_TRY(pdrs_view_col(c, col, &v));
  DevBuf didx, dout;
  const int64_t* di = idx;
  if (idx_mem == PDRS_MEM_HOST) {
    PDRS_TRY(didx.alloc(c, (size_t)n * 8));
    if (n) PDRS_CUDA(c, cudaMemcpyAsync(didx.p, idx, n * 8, cudaMemcpyHostToDevice, c->stream));
    di = didx.as<int64_t>();
  }
  int esz = col->dtype == PDRS_BOOL_BITS ? 1 : pdrs_dtype_bytes(col->dtype);
  void* dout_p = out;
  if (out_mem == PDRS_MEM_HOST) { PDRS_TRY(dout.alloc(c, (size_t)n * esz)); dout_p = dout.p; }
  if (n) {
    int g = pdrs_grid_for(c, n, 256);
    switch (col->dtype) {
      case PDRS_I64: gather_kernel<int64_t><<<g, 256, 0, c->stream>>>((const int64_t*)v.data, v.nulls, di, n, (int64_t*)dout_p, 0); break;
      case PDRS_F64: gather_kernel<double><<<g, 256, 0, c->stream>>>((const double*)v.data, v.nulls, di, n, (double*)dout_p, 0.0); break;
      case PDRS_I32: gather_kernel<int32_t><<<g, 256, 0, c->stream>>>((const int32_t*)v.data, v.nulls, di, n, (int32_t*)dout_p, 0); break;
      case PDRS_DICT_U32: gather_kernel<uint32_t><<<g, 256, 0, c->stream>>>((const uint32_t*)v.data, v.nulls, di, n, (uint32_t*)dout_p, 0xFFFFFFFFu); break;
      case PDRS_BOOL_BITS: gather_bits_kernel<<<g, 256, 0, c->stream>>>((const uint8_t*)v.data, v.nulls, di, n, (uint8_t*)dout_p); break;
    }
    c->stats.kernel_launches++;
    PDRS_CUDA(c, cudaGetLastError());
    if (out_mem == PDRS_MEM_HOST) PDRS_CUDA(c, cudaMemcpyAsync(out, dout_p, (size_t)n * esz, cudaMemcpyDeviceToHost, c->stream));
  }
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PDRS_OK;
}

int32_t pdrs_filter_indices(pdrs_ctx* c, const pdrs_col* mask, int64_t* out_idx_dev, int64_t* n_out_host) {
  if (!c || !mask || !n_out_host) return PDRS_ERR_BAD_ARG;
  if (mask->dtype != PDRS_BOOL_BITS) return pdrs_fail(c, PDRS_ERR_TYPE_MISMATCH, "filter column must be Boolean");
  ColView v;
  PDRS_TRY(pdrs_view_col(c, mask, &v));
  int64_t n = mask->len, nwords = (n + 31) / 32;
  *n_out_host = 0;
  if (n == 0) return PDRS_OK;
  int ctas = (int)std::min<int64_t>((int64_t)c->sm_count * 4, std::max<int64_t>(1, nwords / 64));
  DevBuf counts;
  PDRS_TRY(counts.alloc(c, (size_t)(ctas + 1) * 8));
  auto* cc = counts.as<unsigned long long>();
  filter_count_kernel<<<ctas, 256, 0, c->stream>>>((const uint32_t*)v.data, (const uint32_t*)v.nulls, nwords, n, cc);
  scan_small_kernel<<<1, 32, 0, c->stream>>>(cc, ctas, cc + ctas);
  filter_write_kernel<<<ctas, 32, 0, c->stream>>>((const uint32_t*)v.data, (const uint32_t*)v.nulls, nwords, n, cc, out_idx_dev);
  c->stats.kernel_launches += 3;
  PDRS_CUDA(c, cudaGetLastError());
  PDRS_CUDA(c, cudaMemcpyAsync(c->pinned_scalars, cc + ctas, 8, cudaMemcpyDeviceToHost, c->stream));
  PDRS_CUDA(c, cudaStreamSynchronize(c->stream));
  *n_out_host = c->pinned_scalars[0];
  return PDRS_OK;
}

}  // extern "C"
