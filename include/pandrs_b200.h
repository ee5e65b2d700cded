/* pandrs_b200.h — C ABI of libpandrs_b200.so: the B200 (sm_100a) execution path for pandrs's
 * groupby-aggregate and inner/left hash-join hot paths.
 *
 * pandrs (cool-japan/pandrs) has no FFI/plugin interface for this path (its `src/gpu` is a cudarc
 * shell for dense matrix ops, src/gpu/operations.rs:17-39), so the boundary is introduced at the
 * Rust method level: the functions below are what `extern "C"` blocks in src/groupby,
 * src/optimized/split_dataframe/{group,join}.rs and src/optimized/lazy.rs would bind (see
 * INTEGRATION.md for the Rust side).  Each entry point names the reference code it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types; no exceptions or panics cross the ABI.
 *  - every function returns a pdrs_status (0 = ok, <0 = error); pdrs_last_error() gives the text.
 *    (reference convention: Result<T, Error>, GPU errors -> Error::Computation, src/gpu/mod.rs:206-210)
 *  - inputs are BORROWED for the duration of the call (the Rust side keeps its Arc<[T]> alive);
 *    results are owned by the library and released only by the matching *_free().
 *  - column memory layout is pandrs's own (zero-copy from Arc<[i64]>/Arc<[f64]>/Arc<[u32]>):
 *    null bitmap LSB-first, bit SET = NULL, may be shorter than ceil(len/8) (missing bytes = not NULL)
 *    (src/column/int64_column.rs:72-81, src/core/column.rs:163-177).
 *  - a pdrs_ctx is bound to one device and one stream and is NOT re-entrant; use one per thread
 *    (reference: global Mutex<Option<GpuManager>>, src/gpu/mod.rs:249-251).
 *  - there is NO CPU fallback: an unsupported dtype/op returns PDRS_ERR_UNSUPPORTED.
 */
#ifndef PANDRS_B200_H
#define PANDRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDRS_ABI_VERSION 1

typedef enum pdrs_status {
  PDRS_OK = 0,
  PDRS_ERR_BAD_ARG = -1,        /* Error::InvalidInput / IndexOutOfBounds */
  PDRS_ERR_TYPE_MISMATCH = -2,  /* Error::ColumnTypeMismatch (join.rs:98-104) */
  PDRS_ERR_OOM = -3,
  PDRS_ERR_CUDA = -4,           /* Error::Computation(String) */
  PDRS_ERR_NCCL = -5,
  PDRS_ERR_UNSUPPORTED = -6     /* Error::OperationFailed (aggregation.rs:748-752) */
} pdrs_status;

/* ColumnType (src/core/column.rs:9-14) + I32 as a physical extension for packed multi-key groupbys */
typedef enum pdrs_dtype {
  PDRS_I64 = 0,       /* Int64Column.data: Arc<[i64]>      (src/column/int64_column.rs:9-13)   */
  PDRS_F64 = 1,       /* Float64Column.data: Arc<[f64]>    (src/column/float64_column.rs:9-13) */
  PDRS_DICT_U32 = 2,  /* StringColumn.indices: Arc<[u32]>, GLOBAL_STRING_POOL ids (string_column.rs:26-32) */
  PDRS_BOOL_BITS = 3, /* BooleanColumn.data: BitMask, LSB-first, 1 = true (core/column.rs:72-156)  */
  PDRS_I32 = 4
} pdrs_dtype;

/* AggregateOp, same order as src/optimized/split_dataframe/group/types.rs:11-34 */
typedef enum pdrs_agg_op {
  PDRS_SUM = 0, PDRS_MEAN = 1, PDRS_MIN = 2, PDRS_MAX = 3, PDRS_COUNT = 4, PDRS_STD = 5, PDRS_VAR = 6,
  PDRS_MEDIAN = 7, PDRS_FIRST = 8, PDRS_LAST = 9   /* order-dependent: computed from the row lists, pdrs_group_rows_agg() */
} pdrs_agg_op;

/* JoinType (src/optimized/split_dataframe/join.rs:11-20) */
typedef enum pdrs_join_type { PDRS_INNER = 0, PDRS_LEFT = 1, PDRS_RIGHT = 2, PDRS_OUTER = 3 } pdrs_join_type;

typedef enum pdrs_mem { PDRS_MEM_HOST = 0, PDRS_MEM_DEVICE = 1 } pdrs_mem;

/* One borrowed column.  For PDRS_MEM_DEVICE columns `data` must be 16-byte aligned and
 * `null_len` must cover ceil(len/8) bytes (pdrs_col_upload() guarantees both). */
typedef struct pdrs_col {
  int32_t dtype;            /* pdrs_dtype */
  int32_t mem;              /* pdrs_mem: where data/null_bits live */
  const void* data;
  const uint8_t* null_bits; /* NULL = no NULLs */
  int64_t null_len;         /* bytes available at null_bits */
  int64_t len;              /* rows */
  int64_t null_alias;       /* PDRS_DICT_U32 keys: pool id of the literal string "NULL" (it merges with the
                               NULL group, grouping.rs:69-98), or -1 */
} pdrs_col;

typedef struct pdrs_agg {
  int32_t value_col;        /* index into the `vals` array of pdrs_groupby_agg */
  int32_t op;               /* pdrs_agg_op */
} pdrs_agg;

/* A typed row predicate `column <op> constant` (Int64 / Float64 column): what a caller would otherwise turn into a Boolean
 * column first and pass as `filter` (LazyFrame filter step, src/optimized/lazy.rs:170-182; data_ops.rs:37-62).  A NULL
 * in the column keeps the row out, like a NULL in a Boolean filter column. */
typedef enum pdrs_cmp_op { PDRS_CMP_LT = 0, PDRS_CMP_LE = 1, PDRS_CMP_GT = 2, PDRS_CMP_GE = 3, PDRS_CMP_EQ = 4, PDRS_CMP_NE = 5 } pdrs_cmp_op;
typedef struct pdrs_pred {
  pdrs_col col;             /* PDRS_I64 or PDRS_F64 */
  int32_t op;               /* pdrs_cmp_op */
  int32_t reserved;
  int64_t ival;             /* constant for PDRS_I64 columns */
  double fval;              /* constant for PDRS_F64 columns */
} pdrs_pred;

typedef enum pdrs_groupby_algo {
  PDRS_GB_AUTO = 0,
  PDRS_GB_SHARED = 1,       /* per-warp tables in shared memory, spill to the global table */
  PDRS_GB_GLOBAL = 2,       /* global open-addressing table only (high cardinality) */
  PDRS_GB_DENSE = 3,        /* direct-mapped shared tables for small dense integer key ranges */
  PDRS_GB_TILESORT = 4,     /* tile sort in shared memory + per-thread register aggregation (tens to ~2000 groups) */
  PDRS_GB_PARTITIONED = 5,  /* reported only: hash-partitioned rows + tile sort per partition (thousands to millions of groups) */
  PDRS_GB_FEW = 6           /* reported only: <= 16 groups, sum / mean / count of up to 8 value columns in ONE scan (lane-private accumulators) */
} pdrs_groupby_algo;

typedef struct pdrs_options {
  int32_t device;           /* CUDA ordinal */
  int32_t groupby_algo;     /* pdrs_groupby_algo; AUTO picks from a sampled cardinality estimate */
  int64_t groups_hint;      /* expected number of groups, 0 = estimate from a sample */
  void* stream;             /* cudaStream_t to run on, NULL = the context creates its own */
  int32_t compat_filter_nulls; /* 1 = a fused row filter turns NULLs into defaults like data_ops.rs:64-71 */
  int32_t reserved0;
  int64_t reserved[4];
} pdrs_options;

typedef struct pdrs_stats {   /* filled by the last groupby/join call on the context */
  int64_t kernel_launches;    /* kernels launched by the library since context creation */
  int32_t groupby_algo_used;  /* pdrs_groupby_algo actually run */
  int32_t retries;            /* table-overflow retries */
  int64_t est_groups;         /* sampled cardinality estimate (0 when a hint was given) */
  int64_t table_slots;        /* global table capacity used */
  int64_t spilled_rows;       /* rows that bypassed the shared-memory tables */
  float main_kernel_ms;       /* CUDA-event time of the dominant kernel of the last call */
  float total_ms;             /* CUDA-event time of the whole device-side call */
} pdrs_stats;

typedef struct pdrs_ctx pdrs_ctx;
typedef struct pdrs_groupby_result pdrs_groupby_result;
typedef struct pdrs_join_result pdrs_join_result;

/* ---- context (replaces src/gpu/mod.rs:249-282 get_gpu_manager / GpuConfig) ---- */
int32_t pdrs_abi_version(void);
int32_t pdrs_ctx_create(const pdrs_options* opts /* may be NULL */, pdrs_ctx** out);
void pdrs_ctx_destroy(pdrs_ctx* ctx);
const char* pdrs_last_error(pdrs_ctx* ctx /* NULL = creation errors */);
int32_t pdrs_sync(pdrs_ctx* ctx);
int32_t pdrs_get_stats(pdrs_ctx* ctx, pdrs_stats* out);
int32_t pdrs_set_option(pdrs_ctx* ctx, const char* name, int64_t value);

/* ---- device-resident columns (replaces src/gpu/memory_pool.rs:124-383 for this path) ----
 * pdrs_col_upload copies a host column into 256-byte aligned device memory, pads the null bitmap to
 * ceil(len/8) (+32) zero bytes, and returns a PDRS_MEM_DEVICE descriptor owned by the context. */
int32_t pdrs_col_upload(pdrs_ctx* ctx, const pdrs_col* host_col, pdrs_col* dev_col_out);
int32_t pdrs_col_free(pdrs_ctx* ctx, pdrs_col* dev_col);
/* pinned host staging buffers for callers that want overlapped H2D (bench e2e leg) */
int32_t pdrs_host_alloc(pdrs_ctx* ctx, int64_t bytes, void** out);
int32_t pdrs_host_free(pdrs_ctx* ctx, void* p);

/* ---- groupby-aggregate ----
 * Replaces OptimizedDataFrame::group_by_with_options (group/grouping.rs:38-115) fused with
 * GroupBy::aggregate / par_aggregate (group/aggregation.rs:763-871, 22-182) and the Aggregate arm of
 * LazyFrame::execute (src/optimized/lazy.rs:186-404).
 *  keys[nkeys]   key columns (I64, I32, DICT_U32, BOOL_BITS, F64 by text-equivalent bit pattern)
 *  vals[nvals]   value columns (I64 or F64; other dtypes only under PDRS_COUNT)
 *  aggs[naggs]   (value_col, op); every aggregate is returned as f64 (aggregation.rs:865)
 *  filter        optional BOOL_BITS column: only rows where it is Some(true) take part
 *                (data_ops.rs:49-55 fused in front of the aggregation), NULL = no filter
 * HOST columns of >= "stream_rows" rows (pdrs_set_option, default 2^25) are processed CHUNK BY CHUNK (src/large/mod.rs): chunk
 * i + 1 travels to the device (pageable sources through a pool of staging threads with pinned buffers, pinned sources by direct
 * DMA) while chunk i is aggregated into mergeable states, so the input may be larger than device memory.
 * Semantics follow SURVEY.md §9.1-9.2: NULL keys form one group, Count = group size incl. NULL values,
 * empty/all-NULL -> 0.0, Min/Max sentinel collapse, Std/Var with n-1.  Group order is unspecified. */
int32_t pdrs_groupby_agg(pdrs_ctx* ctx, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals,
                         int32_t nvals, const pdrs_agg* aggs, int32_t naggs, const pdrs_col* filter,
                         pdrs_groupby_result** out);
/* The same with a typed predicate evaluated inside the scan (filter and pred may both be given: rows must pass both).
 * Fuses the comparison that would have produced the Boolean filter column: +8 bytes per row instead of a separate pass
 * (the 48 B/row variant of configs[4]). */
int32_t pdrs_groupby_agg_where(pdrs_ctx* ctx, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals,
                               const pdrs_agg* aggs, int32_t naggs, const pdrs_col* filter, const pdrs_pred* pred,
                               pdrs_groupby_result** out);
int64_t pdrs_groupby_n_groups(const pdrs_groupby_result* r);
/* copy-out to host: key k as its physical type (i64/f64/u32/i32; BOOL as one byte per group) + 1 byte
 * per group that is 1 where the key part is NULL (the Rust side prints "NULL", grouping.rs:74) */
int32_t pdrs_groupby_key(const pdrs_groupby_result* r, int32_t k, void* out_values, uint8_t* out_is_null);
int32_t pdrs_groupby_agg_values(const pdrs_groupby_result* r, int32_t a, double* out);
int32_t pdrs_groupby_group_rows(const pdrs_groupby_result* r, int64_t* out);          /* = Count */
int32_t pdrs_groupby_valid_n(const pdrs_groupby_result* r, int32_t value_col, int64_t* out);
/* device pointers of the same arrays (valid until the result is freed) */
const void* pdrs_groupby_key_dev(const pdrs_groupby_result* r, int32_t k);
const uint8_t* pdrs_groupby_key_null_dev(const pdrs_groupby_result* r, int32_t k);
const double* pdrs_groupby_agg_dev(const pdrs_groupby_result* r, int32_t a);
const int64_t* pdrs_groupby_group_rows_dev(const pdrs_groupby_result* r);
void pdrs_groupby_result_free(pdrs_groupby_result* r);

/* Partial (mergeable) aggregation states for the multi-GPU path (no reference counterpart: pandrs has
 * no comms backend).  A state row is 8 x 64-bit words per (group, value column), opaque
 * to the caller:  [rows, n, pivot^tag, S1, S2, ~ord(min), ord(max), isum]  with S1 = sum(x - pivot),
 * S2 = sum((x - pivot)^2), so that states of the same group merge exactly like Chan's update.
 * pdrs_groupby_partial() runs the same kernels as pdrs_groupby_agg() but keeps the states
 * (all_stats = 0: rows/n/sum only; 1: also min/max/variance terms);
 * pdrs_groupby_merge() groups state rows by key again, merges them and finalises `aggs`. */
int32_t pdrs_groupby_partial(pdrs_ctx* ctx, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals,
                             int32_t nvals, const pdrs_col* filter, int32_t all_stats,
                             pdrs_groupby_result** out);
const uint64_t* pdrs_groupby_states_dev(const pdrs_groupby_result* r, int32_t value_col); /* [n_groups][8] */
int32_t pdrs_groupby_merge(pdrs_ctx* ctx, const pdrs_col* keys, int32_t nkeys, const uint64_t* const* states_dev,
                           const int32_t* val_is_int, int32_t nvals, int64_t n_state_rows,
                           const pdrs_agg* aggs, int32_t naggs, pdrs_groupby_result** out);
/* destination rank of every row, dest = mix(hash(key tuple)) mod nparts (NULL-key group -> 0), as a
 * permutation perm_dev[nrows] that groups the row ids by destination, and counts_host[nparts]. */
int32_t pdrs_hash_partition(pdrs_ctx* ctx, const pdrs_col* keys, int32_t nkeys, int32_t nparts,
                            int64_t* perm_dev, int64_t* counts_host);

/* ---- row lists per group ----
 * Replaces OptimizedDataFrame::par_groupby (group/grouping.rs:124-331) up to the sub-frame construction, and the row lists of
 * GroupBy (`groups: HashMap<Vec<String>, Vec<usize>>`, grouping.rs:62-104) for the order-dependent aggregates:
 *   ids[offsets[g] .. offsets[g + 1]) = the rows of group g in ASCENDING order (what the reference pushes, grouping.rs:103, 188),
 *   keys as in pdrs_groupby_key (the caller labels a NULL part "NA" and joins the parts with "_", grouping.rs:158-186).
 * One sub-frame per group = one pdrs_gather per column over the WHOLE permutation `ids` (filter_by_indices, data_ops.rs:124-211)
 * and a slice [offsets[g], offsets[g + 1]) of the gathered column per group.  At most 2^32 - 2 rows per call.  Group order is
 * unspecified (HashMap). */
typedef struct pdrs_group_rows pdrs_group_rows;
int32_t pdrs_groupby_rows(pdrs_ctx* ctx, const pdrs_col* keys, int32_t nkeys, pdrs_group_rows** out);
int64_t pdrs_group_rows_n_groups(const pdrs_group_rows* r);
int64_t pdrs_group_rows_n_rows(const pdrs_group_rows* r);
int32_t pdrs_group_rows_key(const pdrs_group_rows* r, int32_t k, void* out_values, uint8_t* out_is_null);
int32_t pdrs_group_rows_offsets(const pdrs_group_rows* r, int64_t* out /* n_groups + 1 */);
int32_t pdrs_group_rows_ids(const pdrs_group_rows* r, int64_t* out /* n_rows */);
const int64_t* pdrs_group_rows_offsets_dev(const pdrs_group_rows* r);
const int64_t* pdrs_group_rows_ids_dev(const pdrs_group_rows* r);
/* AggregateOp::Median / First / Last of an Int64 / Float64 column over the row lists (group/aggregation.rs:585-624, 703-742):
 * First / Last = the value of the group's first / last row as f64, 0.0 when it is NULL; Median = middle of the sorted non-NULL
 * values (mean of the two middle ones for an even count; Int64: wrapping i64 sum, then / 2.0), 0.0 for an all-NULL group.
 * Groups holding NaN: unspecified (the reference sorts with partial_cmp(..).unwrap_or(Equal)).  out_host[n_groups]. */
int32_t pdrs_group_rows_agg(pdrs_group_rows* r, const pdrs_col* val, int32_t op, double* out_host);
void pdrs_group_rows_free(pdrs_group_rows* r);

/* ---- columnar ingest (the step in front of the path) ----
 * Arrow validity bitmap -> pandrs null mask.  Arrow: bit SET = valid, row i at bit (bit_offset + i), LSB first; pandrs: bit SET =
 * NULL, row i at bit i (src/column/common/utils create_bitmask, core/column.rs:163-177).  Replaces the per-element is_null() loops
 * of ArrowConverter::arrow_array_to_series (src/arrow_integration.rs:160-225).  out_null_bits: ceil(len / 8) bytes (trailing bits
 * of the last byte are 0).  validity == NULL means "all valid" (an all-zero mask).  n_nulls_host (optional) = number of NULL rows. */
int32_t pdrs_arrow_validity_to_nulls(pdrs_ctx* ctx, const uint8_t* validity, int32_t validity_mem, int64_t bit_offset, int64_t len,
                                     uint8_t* out_null_bits, int32_t out_mem, int64_t* n_nulls_host);
/* Dictionary encoding of an Arrow Utf8 (32-bit offsets) / LargeUtf8 (64-bit offsets) array on the device: what
 * StringColumn::with_nulls -> GLOBAL_STRING_POOL.add_strings (src/column/string_column.rs:95-112, string_pool.rs:28-52) does on the
 * host under an RwLock.  ids[i] = dense id of row i's string, handed out in order of FIRST OCCURRENCE (the order get_or_insert assigns
 * in a row loop); a NULL row is the empty string (its placeholder in StringColumn::with_nulls) and keeps its bit in the null mask.
 * first_rows[id] = the first row carrying that string (the host reads the text from its own copy of the array and interns it;
 * pdrs_dict_remap then rewrites the ids into the ids of an existing pool).  Strings are compared through two independent 64-bit
 * hashes (the offsets are validated first: ascending, inside the value bytes - PDRS_ERR_BAD_ARG otherwise); a verification pass compares the BYTES of every row with those of its id's first row and the call fails with
 * PDRS_ERR_UNSUPPORTED should they ever differ - ids are never wrong.  offsets / bytes / validity live where `mem` says. */
typedef struct pdrs_dict pdrs_dict;
int32_t pdrs_dict_encode(pdrs_ctx* ctx, const void* offsets /* len + 1 */, int32_t offsets_are_64, const uint8_t* bytes, int64_t nbytes,
                         const uint8_t* validity, int64_t bit_offset, int64_t len, int32_t mem, pdrs_dict** out);
int64_t pdrs_dict_n_unique(const pdrs_dict* d);
int32_t pdrs_dict_ids(const pdrs_dict* d, uint32_t* out_host /* len */);
const uint32_t* pdrs_dict_ids_dev(const pdrs_dict* d);
int32_t pdrs_dict_first_rows(const pdrs_dict* d, int64_t* out_host /* n_unique */);
const uint8_t* pdrs_dict_nulls_dev(const pdrs_dict* d);     /* pandrs null mask of the column (NULL when validity was NULL) */
int32_t pdrs_dict_remap(pdrs_dict* d, const uint32_t* new_ids_host /* n_unique */);
void pdrs_dict_free(pdrs_dict* d);

/* ---- hash join ----
 * Replaces the build/probe part of OptimizedDataFrame::join_impl (split_dataframe/join.rs:107-208).
 * Emits index pairs (left_row, right_row); -1 stands for None: right_row = -1 for a left row without a match
 * (Left / Outer), left_row = -1 for the right rows that no pair references, appended in ascending right row after
 * the other pairs (Right / Outer, join.rs:211-224; right rows with a NULL key are among them).
 * NULL keys never match and are dropped from BOTH sides, also for Left (join.rs:112,152).
 * Row order: small inputs keep the reference's order (left-row-major, ascending right row).  Large inputs
 * take the radix-partitioned path and emit the same multiset of pairs in an unspecified order; parity is
 * defined after a canonical sort of the pairs (BASELINE.json north_star).  With unique build keys the
 * result arrays are sized for one pair per left row; pdrs_join_len() gives the number of valid pairs. */
int32_t pdrs_join_pairs(pdrs_ctx* ctx, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how,
                        pdrs_join_result** out);
/* pdrs_join_pairs + the materialisation of columns of the RIGHT frame in one call: replaces the build / probe loops AND
 * the per-column gather loops join_impl runs for the right frame's non-key columns (join.rs:107-208 + 290-552):
 *   out column k, pair j = right_row[j] < 0 || right_cols[k][right_row[j]] is NULL ? type default : right_cols[k][right_row[j]]
 * (no null mask, defaults as in pdrs_gather).  Same pairs, same order rules as pdrs_join_pairs.  Inner / Left joins of large
 * inputs on unique build keys carry up to two Int64 / Float64 columns WITH the build rows - through the radix partition
 * and inside 32-byte hash-table slots {key, row, c0, c1} - so a match yields its column values in the L2 sector the key
 * comparison read anyway; everything else is gathered by right row after the join.  The results are identical. */
int32_t pdrs_join_gather(pdrs_ctx* ctx, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how,
                         const pdrs_col* right_cols, int32_t n_right_cols, pdrs_join_result** out);
int32_t pdrs_join_right_col(const pdrs_join_result* r, int32_t k, void* out_host);   /* pdrs_join_len() values of the column's physical type (BOOL: one byte each) */
const void* pdrs_join_right_col_dev(const pdrs_join_result* r, int32_t k);
int64_t pdrs_join_len(const pdrs_join_result* r);
int32_t pdrs_join_indices(const pdrs_join_result* r, int64_t* left_out, int64_t* right_out); /* host copy-out */
const int64_t* pdrs_join_left_dev(const pdrs_join_result* r);
const int64_t* pdrs_join_right_dev(const pdrs_join_result* r);
void pdrs_join_result_free(pdrs_join_result* r);

/* ---- multi-GPU join: fused radix partition + shuffle over NVLink peer memory ----
 * No reference counterpart (pandrs has no comms backend; PartitionStrategy::Hash is only an enum,
 * src/distributed/core/partition.rs:11-18).  Semantics = join_impl (join.rs:107-208) applied to the union of the ranks'
 * rows, Inner / Left; rank r returns the pairs of the keys whose rank hash maps to r, in GLOBAL row numbers.
 * One pdrs_xjoin per rank (process / GPU).  Protocol, identical on every rank:
 *   create(rank, world, max rows per rank of each side, total right rows)    same arguments => same layout everywhere
 *   ipc_handle -> all_gather of the 64-byte handles (torch.distributed / MPI) -> attach_ipc       [once]
 *   shuffle(left_key, right_key, right_row0)   the one-pass partition kernel stores every (key, row) run straight into the
 *                                              destination rank's receive area (CUDA IPC mapping, NVLink stores)
 *   <barrier over all ranks>
 *   local(how, left_row0[world]) -> pdrs_join_result   build / probe of the received sub-buckets
 *   <barrier before the next shuffle>
 * world must be 1, 2, 4 or 8 (one NVSwitch domain).  attach_ptrs replaces the IPC exchange when all ranks live in
 * one process.  PDRS_ERR_UNSUPPORTED from shuffle = a padded sub-bucket overflowed (skewed keys): fall back to
 * pdrs_hash_partition + all_to_all + pdrs_join_pairs. */
typedef struct pdrs_xjoin pdrs_xjoin;
int32_t pdrs_xjoin_create(pdrs_ctx* ctx, int32_t rank, int32_t world, int64_t max_left_rows, int64_t max_right_rows,
                          int64_t total_right_rows, pdrs_xjoin** out);
int64_t pdrs_xjoin_bytes(const pdrs_xjoin* x);                       /* size of the receive area */
void* pdrs_xjoin_base(const pdrs_xjoin* x);                          /* device pointer of the receive area */
int32_t pdrs_xjoin_ipc_handle(pdrs_xjoin* x, uint8_t* handle64);     /* cudaIpcMemHandle_t of the receive area */
int32_t pdrs_xjoin_attach_ipc(pdrs_xjoin* x, const uint8_t* handles /* world x 64 bytes */);
int32_t pdrs_xjoin_attach_ptrs(pdrs_xjoin* x, void* const* bases /* world */);
int32_t pdrs_xjoin_shuffle(pdrs_xjoin* x, const pdrs_col* left_key, const pdrs_col* right_key, int64_t right_row0);
int32_t pdrs_xjoin_local(pdrs_xjoin* x, int32_t how, const int64_t* left_row0 /* world */, pdrs_join_result** out);
void pdrs_xjoin_destroy(pdrs_xjoin* x);

/* ---- multi-GPU operators: one process per GPU, NCCL over NVLink / NVSwitch ----
 * No reference counterpart (pandrs has no comms backend; PartitionStrategy::Hash is only an enum,
 * src/distributed/core/partition.rs:11-18).  Semantics = the single-frame operator of the reference applied to the union of
 * the ranks' rows.  NCCL is bound at run time (dlopen): single-GPU users never load it.
 *   rank 0: pdrs_comm_unique_id(id) -> the host broadcasts the 128 bytes (MPI, TCP, torch.distributed ...) -> every rank:
 *   pdrs_comm_init(ctx, nranks, rank, id, &comm).  All *_dist calls are collective: every rank calls them in the same order. */
typedef struct pdrs_comm pdrs_comm;
int32_t pdrs_comm_unique_id(uint8_t* id128 /* 128 bytes */);
int32_t pdrs_comm_init(pdrs_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t* id128, pdrs_comm** out);
int32_t pdrs_comm_rank(const pdrs_comm* comm);
int32_t pdrs_comm_size(const pdrs_comm* comm);
int32_t pdrs_comm_set_option(pdrs_comm* comm, const char* name, int64_t value);   /* "groups_cap": state rows per rank in the replicated groupby (default 4096) */
int32_t pdrs_comm_barrier(pdrs_comm* comm);
/* CUDA-event time of the exchange step (all-gather / all-to-all / peer-store shuffle) of the last *_dist call and the bytes this
 * rank sent to other ranks in it (NVLink roofline: bytes / time against the per-direction peer bandwidth) */
int32_t pdrs_comm_last_exchange(const pdrs_comm* comm, float* ms, int64_t* bytes_to_peers);
void pdrs_comm_destroy(pdrs_comm* comm);
/* groupby(keys).agg(...) over the union of the ranks' rows (same arguments as pdrs_groupby_agg_where on every rank's shard).
 * Every rank aggregates its own rows into mergeable states, then
 *   result_mode 1 (replicated)  one fixed-size all-gather of the per-group states + merge: EVERY rank returns ALL groups
 *                               (low cardinality: at most "groups_cap" groups per rank);
 *   result_mode 2 (sharded)     the states are exchanged by hash(key) mod ranks with one all-to-all and merged by their owner:
 *                               every rank returns the groups it owns (high cardinality; the NULL-key group lives on rank 0);
 *   result_mode 0 (auto)        replicated when every rank's group count fits, else sharded (the choice is collective).
 * Only per-group states cross NVLink, never input rows.  pdrs_get_stats() reports the local aggregation kernel. */
int32_t pdrs_groupby_agg_dist(pdrs_comm* comm, const pdrs_col* keys, int32_t nkeys, const pdrs_col* vals, int32_t nvals,
                              const pdrs_agg* aggs, int32_t naggs, const pdrs_col* filter, const pdrs_pred* pred,
                              int32_t result_mode, pdrs_groupby_result** out);
/* Inner / Left join over the union of the ranks' rows: the pdrs_xjoin protocol above (IPC handle exchange, shuffle through
 * NVLink peer stores, barrier, local join) behind one collective call.  left_row0 / right_row0 = global number of this rank's
 * first row; max_* = the largest shard of any rank (same values on every rank; the receive areas are set up for them once and
 * reused).  Rank r returns the pairs of the keys whose rank hash maps to r, in GLOBAL row numbers.
 * PDRS_ERR_UNSUPPORTED (on every rank): skewed keys overflowed a padded region. */
int32_t pdrs_join_pairs_dist(pdrs_comm* comm, const pdrs_col* left_key, const pdrs_col* right_key, int32_t how, int64_t left_row0,
                             int64_t right_row0, int64_t max_left_rows, int64_t max_right_rows, int64_t total_right_rows,
                             pdrs_join_result** out);

/* Replaces the materialisation loops of join_impl (join.rs:290-552) and filter_by_indices
 * (data_ops.rs:124-211): out[j] = idx[j] < 0 || col[idx[j]] is NULL ? type default : col[idx[j]];
 * the output carries no null mask.  DICT_U32 default is 0xFFFFFFFF (the empty string ""), BOOL_BITS
 * gathers into one byte per row.  idx/out live where idx_mem/out_mem say. */
int32_t pdrs_gather(pdrs_ctx* ctx, const pdrs_col* col, const int64_t* idx, int32_t idx_mem, int64_t n,
                    void* out, int32_t out_mem);

/* Replaces the index-building half of OptimizedDataFrame::filter (data_ops.rs:37-62):
 * ascending row ids where the Boolean column is Some(true).  out_idx_dev must hold mask->len entries. */
int32_t pdrs_filter_indices(pdrs_ctx* ctx, const pdrs_col* mask, int64_t* out_idx_dev, int64_t* n_out_host);

/* ---- synthetic inputs for tests/benches (same counter-based arithmetic as oracle/pandrs_oracle.cpp) ---- */
int32_t pdrs_synth_keys(pdrs_ctx* ctx, int64_t* out_dev, int64_t n, int64_t row0, uint64_t seed, uint64_t card, int32_t scramble);
int32_t pdrs_synth_vals(pdrs_ctx* ctx, double* out_dev, int64_t n, int64_t row0, uint64_t seed);
int32_t pdrs_synth_nulls(pdrs_ctx* ctx, uint8_t* out_dev, int64_t n, int64_t row0, uint64_t seed, uint32_t per_million);
int32_t pdrs_synth_join_keys(pdrs_ctx* ctx, int64_t* out_dev, int64_t n, int64_t row0, uint64_t seed, uint64_t domain, int32_t unique);
/* raw device memory helpers for hosts without a CUDA runtime binding of their own */
int32_t pdrs_dev_alloc(pdrs_ctx* ctx, int64_t bytes, void** out);
int32_t pdrs_dev_free(pdrs_ctx* ctx, void* p);
int32_t pdrs_memcpy(pdrs_ctx* ctx, void* dst, const void* src, int64_t bytes, int32_t kind /*0 h2d,1 d2h,2 d2d*/);
/* writes > L2-size bytes so the next timed call starts with a cold L2 */
int32_t pdrs_flush_l2(pdrs_ctx* ctx);
/* CUDA-event timing on the context stream for callers without a CUDA binding: begin/end return ms */
int32_t pdrs_timer_begin(pdrs_ctx* ctx);
int32_t pdrs_timer_end(pdrs_ctx* ctx, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* PANDRS_B200_H */
