// Internal declarations shared by the translation units of libpandrs_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/pandrs_b200.h"

#define PDRS_MAX_KEYS 4
#define PDRS_MAX_WORDS 3
#define PDRS_MAX_VALS 16
#define PDRS_MAX_AGGS 64

struct PdrsStager;   // stage.cu

struct pdrs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  pdrs_options opts{};
  pdrs_stats stats{};
  int sm_count = 148;
  int smem_optin = 0;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  cudaStream_t aux_stream[2] = {nullptr, nullptr};     // bucket-at-a-time join: two streams alternate over the radix buckets
  cudaEvent_t aux_fork = nullptr, aux_done[2] = {nullptr, nullptr};
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  int64_t* pinned_scalars = nullptr;   // small pinned host area for D2H of counters (64 x i64)
  // tunables (pdrs_set_option)
  int64_t opt_warps = 0;               // 0 = auto (warps per CTA of the shared-memory groupby kernel)
  int64_t opt_sample_rows = 1 << 18;
  int64_t opt_ctas_per_sm = 0;
  int64_t opt_ng = 0;                  // 0 = auto (accumulator replicas per warp)
  int64_t opt_join_algo = 0;           // 0 = auto
  int64_t opt_join_log_nb = 0;         // 0 = auto (log2 of the number of radix buckets)
  int64_t opt_join_ctas_per_sm = 0;    // 0 = auto
  int64_t opt_join_part = 0;           // 0 = auto (one-pass padded partition, exact two-pass on overflow), 2 = always two-pass
  int64_t opt_join_slots_mult = 0;     // table slots per build row (0 = default 3)
  int64_t opt_join_bucketwise = 1;     // radix join: one reused, L2-resident table region per bucket (0 = one table for all buckets)
  int64_t opt_join_region_mb = 0;      // ... size of that region (0 = default 24 MB)
  int64_t opt_join_prefetch = 1;       // stream the next radix bucket's table region into L2 ahead of its first probes
  int64_t opt_join_emit = 0;           // 0 = auto (single-pass probe + emit when the build keys are unique), 2 = always count / scan / write
  int64_t opt_key_compress = 1;        // pack multi-key tuples by value range when that brings them down to one 64-bit word
  int64_t opt_part_side = 1;           // partitioned groupby: runs of full buckets (hot keys) go to a side area that is aggregated as extra partitions
  int64_t opt_spillbuf = 1;            // skew fallback of the tile-sort kernel: spilled rows go to a side buffer + second pass (0: straight to the global table)
  int64_t opt_xjoin_mode = 0;          // exchange join: 0 = auto, 1 = fused (rank x radix bucket in one pass), 2 = staged (shuffle by rank, local radix partition)
  int64_t opt_timing = 1;              // record CUDA-event times in pdrs_stats
  int64_t opt_radix = 1;               // allow the radix-partitioned high-cardinality groupby path
  int64_t opt_dense = 1;               // allow the direct-mapped path for small dense integer keys
  int64_t opt_tsort = 1;               // allow the tile-sort kernel (gb_tsort.cu)
  int64_t opt_part = 1;                // allow the hash-partitioned high-cardinality path (gb_part.cu)
  int64_t opt_tsort_heavy = 0;         // 0 = default 256
  int64_t opt_tsort_mid = 0;           // 0 = default 48
  int64_t opt_tsort_team = 0;          // 0 = auto (estimated groups < 500), 1 = always, 2 = never
  int64_t opt_tsort_threads = 0;       // 0 = auto, else 512 / 1024 threads per CTA
  int64_t opt_empty_string_id = 0xFFFFFFFFll;   // compat_filter_nulls: the dictionary id NULL string keys turn into (the pool id of ""; default = the id pdrs_gather fills in for "")
  int64_t opt_part_hot = 1;            // partitioned groupby: rows of the sample's hot keys bypass the hash partitions (skewed keys)
  int64_t opt_part_direct = 1;         // partitioned groupby: complete partitions write their groups straight to the result
  int64_t opt_part_hash = 0;           // ... aggregated by the shared-memory hash kernel: 0 = auto (few rows per group), 1 = always, 2 = never (tile sort)
  int64_t opt_few = 1;                 // allow the few-groups multi-column kernel (gb_few.cu)
  int64_t opt_tsort_min_groups = 2;    // a single group serialises the tile histogram: the per-warp shared tables win
  int64_t opt_stage_threads = 0;       // staging engine (stage.cu): worker threads, 0 = auto (<= 8), -1 = plain cudaMemcpyAsync from the caller's memory
  int64_t opt_stream_rows = 1 << 25;   // groupby over HOST columns with at least this many rows runs chunk by chunk (0 = never)
  int64_t opt_stream_chunk_rows = 0;   // rows per chunk of that path (0 = default 2^26)
  int64_t opt_stream_compact_rows = 1 << 22;   // ... appended state rows beyond which (and beyond 4x the groups of a chunk) they are re-merged
  int64_t opt_xjoin_round_rows = 0;    // pdrs_join_pairs_dist: left rows per round (0 = as many as the staged row encoding allows)
  int64_t opt_trace = 0;               // 1: print host-clock phase marks of the collective operators to stderr (each mark synchronises the stream)
  double trace_t0 = 0.0;
  PdrsStager* stager = nullptr;
  bool big_free_pending = false;       // a block of >= 256 MB went back to the pool since the last synchronisation (pdrs_settle_frees)
};
// Multi-GB blocks given back with cudaFreeAsync are only cheap to get again once the host has SEEN the stream pass the frees
// (measured, tools/bisect_join.py: back-to-back joins without a synchronisation in between took 40 - 80 ms instead of 18.9 ms -
// the pool grew by fresh driver allocations instead of reusing its blocks).  Operators call this before they allocate.
inline void pdrs_settle_frees(pdrs_ctx* c) { if (c->big_free_pending) { cudaStreamSynchronize(c->stream); c->big_free_pending = false; } }
void pdrs_trace(pdrs_ctx* c, const char* label);     // ctx.cu; no-op unless opt_trace

int32_t pdrs_fail(pdrs_ctx* ctx, int32_t code, const char* fmt, ...);
// stage.cu: asynchronous host -> device copies through a pool of staging threads (pageable sources) / direct DMA (pinned sources)
// d2h = 1: the other direction (dst = host, src = device); the data is in place after pdrs_stage_join + a stream synchronisation
int32_t pdrs_stage_copy_async(pdrs_ctx* c, void* dst, const void* src, size_t bytes, int d2h = 0);
// device -> host copy of a finished result: through the staging engine when large, plain cudaMemcpyAsync otherwise; returns synchronised
int32_t pdrs_copy_to_host(pdrs_ctx* c, void* dst_host, const void* src_dev, size_t bytes);
int32_t pdrs_stage_join(pdrs_ctx* c, cudaStream_t consumer);
void pdrs_stage_destroy(pdrs_ctx* c);

#define PDRS_CUDA(ctx, call)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return pdrs_fail((ctx), _e == cudaErrorMemoryAllocation ? PDRS_ERR_OOM : PDRS_ERR_CUDA,  \
                       "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define PDRS_TRY(call)                 \
  do {                                 \
    int32_t _s = (call);               \
    if (_s != PDRS_OK) return _s;      \
  } while (0)

// Stream-ordered device memory owned by a call; freed (stream-ordered) when the holder dies.
struct DevBuf {
  pdrs_ctx* ctx = nullptr;
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
  DevBuf& operator=(DevBuf&& o) noexcept { release(); ctx = o.ctx; p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; return *this; }
  ~DevBuf() { release(); }
  int32_t alloc(pdrs_ctx* c, size_t n, bool zero = false);
  void release();
  template <class T> T* as() const { return (T*)p; }
};

// A device view of one input column (uploaded on demand when the caller passed host memory).
struct ColView {
  int32_t dtype = 0;
  const void* data = nullptr;
  const uint8_t* nulls = nullptr;   // covers >= ceil(len/8) + 32 bytes when owned
  int64_t len = 0;
  int64_t null_alias = -1;
  DevBuf own_data, own_nulls;
};
int32_t pdrs_view_col(pdrs_ctx* ctx, const pdrs_col* c, ColView* out);
int pdrs_dtype_bytes(int32_t dtype);
// device memory a stream-ordered allocation can still get: what the driver reports free PLUS what the context's memory pool holds
// in reserve (blocks freed by earlier calls stay cached in the pool and do not show up as free)
int32_t pdrs_mem_available(pdrs_ctx* c, size_t* bytes);
int pdrs_grid_for(pdrs_ctx* c, int64_t n, int threads);

// ---- device helpers ----
#ifdef __CUDACC__
__host__ __device__ __forceinline__ uint64_t pdrs_mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__device__ __forceinline__ bool pdrs_bit(const uint8_t* bits, int64_t i) { return (bits[i >> 3] >> (i & 7)) & 1; }
// order-preserving maps to unsigned; 0 is reserved as "empty" (only reachable by NaN payloads, which are skipped)
__device__ __forceinline__ uint64_t pdrs_ord_f64(double v) {
  uint64_t u = (uint64_t)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
__device__ __forceinline__ double pdrs_unord_f64(uint64_t o) {
  uint64_t u = (o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFULL) : ~o;
  return __longlong_as_double((long long)u);
}
#endif
