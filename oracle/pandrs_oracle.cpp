// pandrs CPU ORACLE — TEST INFRASTRUCTURE ONLY.
//
// A CPU restatement of the pandrs (cool-japan/pandrs) groupby-aggregate, hash-join and
// boolean-filter algorithms, following the reference Rust files line by line.  It is the
// checker for the CUDA path in pandrs_b200/csrc: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  Nothing in pandrs_b200/
// imports, links or calls it, and the product library has no CPU fallback.
//
// Parity pinning: the Rust crate cannot be built in this image (no cargo/rustc), so this
// restatement is pinned by the reference's own known-answer tests (tests/test_oracle_golden.py):
//   src/dataframe/pandas_compat/groupby.rs:480-617, tests/optimized_join_test.rs:6-237,
//   tests/concurrency_test.rs:351-384, tests/optimized_custom_aggregation_test.rs:49,
//   src/dataframe/pandas_compat/merge.rs:320-410 (ordering), SURVEY.md §9.6.
// Everything those tests do not cover (NULL keys/values, Std on the optimized frame, multi-key,
// duplicate-key fan-out) is "parity pinned by restatement, not by reference tests".
//
// Reference files followed (paths relative to the pandrs tree):
//   src/column/int64_column.rs:65-84, float64_column.rs:65-83, boolean_column.rs:72-91,
//   src/core/column.rs:112-129,163-177        -> col_is_null / get_* (null bit = 1, LSB first,
//                                                short masks mean "not NULL")
//   src/optimized/split_dataframe/group/grouping.rs:38-115   -> orc_groupby grouping loop
//   src/optimized/split_dataframe/group/grouping.rs:124-331  -> orc_par_groupby ("NA" / "_" labels, ascending row lists)
//   src/optimized/split_dataframe/group/aggregation.rs:500-754,875-903 -> calc_agg / variance
//   src/optimized/split_dataframe/group/aggregation.rs:22-182,763-871  -> serial / parallel drivers
//   src/optimized/lazy.rs:186-404                                      -> mode ORC_MODE_LAZY
//   src/optimized/split_dataframe/join.rs:76-555                       -> orc_join / orc_gather
//   src/optimized/split_dataframe/data_ops.rs:37-211                   -> orc_filter_indices
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

extern "C" {

enum { ORC_I64 = 0, ORC_F64 = 1, ORC_DICT_U32 = 2, ORC_BOOL_BITS = 3, ORC_I32 = 4 };
// AggregateOp discriminants in the order of group/types.rs:11-34
enum { ORC_SUM = 0, ORC_MEAN = 1, ORC_MIN = 2, ORC_MAX = 3, ORC_COUNT = 4, ORC_STD = 5, ORC_VAR = 6, ORC_MEDIAN = 7, ORC_FIRST = 8, ORC_LAST = 9 };
enum { ORC_MODE_AGGREGATE = 0, ORC_MODE_PAR_AGGREGATE = 1, ORC_MODE_LAZY = 2,
       // NOT a reference mode: the same grouping, but f64 Sum / Mean / Std / Var accumulated in 80-bit long double (two-pass
       // variance).  tests/_util.py uses it to measure the reference's OWN rounding error, so that the 1e-12 tolerance can be
       // asserted relative to the result instead of relative to max|x|.
       ORC_MODE_EXACT = 3 };
enum { ORC_INNER = 0, ORC_LEFT = 1, ORC_RIGHT = 2, ORC_OUTER = 3 };

typedef struct {
  int32_t dtype;
  int32_t _pad;
  const void* data;
  const uint8_t* null_bits;  // may be NULL; bit set = NULL (int64_column.rs:72-81)
  int64_t null_len;          // bytes in null_bits; bytes past the end read as "not NULL"
  int64_t len;
  const char* const* pool;   // ORC_DICT_U32 only: id -> NUL-terminated string
  int64_t pool_len;
} orc_col;
}

namespace {

inline bool col_is_null(const orc_col& c, int64_t i) {
  if (!c.null_bits) return false;
  int64_t byte_idx = i / 8;
  int bit_idx = (int)(i % 8);
  return byte_idx < c.null_len && (c.null_bits[byte_idx] & (1u << bit_idx)) != 0;
}

// Rust `impl Display for f64` prints the shortest digits that round-trip, never in exponent form.
std::string f64_display(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  char buf[512];
  auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
  return std::string(buf, r.ptr);
}

// `Some(v) -> v.to_string()`, `None -> "NULL"`  (grouping.rs:69-98)
std::string key_part(const orc_col& c, int64_t row) {
  if (col_is_null(c, row)) return "NULL";
  switch (c.dtype) {
    case ORC_I64: return std::to_string(((const int64_t*)c.data)[row]);
    case ORC_I32: return std::to_string(((const int32_t*)c.data)[row]);
    case ORC_F64: return f64_display(((const double*)c.data)[row]);
    case ORC_DICT_U32: {
      uint32_t id = ((const uint32_t*)c.data)[row];
      if (c.pool && (int64_t)id < c.pool_len) return std::string(c.pool[id]);
      return "#" + std::to_string(id);  // no pool supplied: ids stand for distinct strings
    }
    case ORC_BOOL_BITS: {
      const uint8_t* b = (const uint8_t*)c.data;
      return ((b[row / 8] >> (row % 8)) & 1) ? "true" : "false";
    }
  }
  return "NULL";
}

// SipHash-1-3, the hasher behind Rust's std HashMap; only iteration order depends on it in the
// reference, but the port keeps it so that the CPU baseline pays the same per-row hashing cost.
struct Sip13 {
  uint64_t v0, v1, v2, v3, tail = 0; size_t ntail = 0, length = 0;
  Sip13() { v0 = 0x736f6d6570736575ULL; v1 = 0x646f72616e646f6dULL; v2 = 0x6c7967656e657261ULL; v3 = 0x7465646279746573ULL; }
  static inline uint64_t rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
  inline void round() {
    v0 += v1; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32);
    v2 += v3; v3 = rotl(v3, 16); v3 ^= v2;
    v0 += v3; v3 = rotl(v3, 21); v3 ^= v0;
    v2 += v1; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32);
  }
  void write(const uint8_t* p, size_t n) {
    length += n;
    for (size_t i = 0; i < n; i++) {
      tail |= (uint64_t)p[i] << (8 * ntail);
      if (++ntail == 8) { v3 ^= tail; round(); v0 ^= tail; tail = 0; ntail = 0; }
    }
  }
  uint64_t finish() {
    uint64_t b = ((uint64_t)length << 56) | tail;
    v3 ^= b; round(); v0 ^= b; v2 ^= 0xff; round(); round(); round();
    return v0 ^ v1 ^ v2 ^ v3;
  }
};

struct VecStrHash {
  size_t operator()(const std::vector<std::string>& k) const {
    Sip13 h; uint64_t n = k.size(); h.write((const uint8_t*)&n, 8);   // Vec<T>: len prefix
    for (auto& s : k) { h.write((const uint8_t*)s.data(), s.size()); uint8_t ff = 0xff; h.write(&ff, 1); }  // str: bytes + 0xff
    return (size_t)h.finish();
  }
};
struct StrHash {
  size_t operator()(const std::string& s) const {
    Sip13 h; h.write((const uint8_t*)s.data(), s.size()); uint8_t ff = 0xff; h.write(&ff, 1); return (size_t)h.finish();
  }
};

// aggregation.rs:881-903
double calculate_variance(const std::vector<double>& values) {
  if (values.empty()) return 0.0;
  double n = (double)values.size();
  double s = 0.0; for (double v : values) s += v;
  double mean = s / n;
  double ssd = 0.0; for (double v : values) { double d = v - mean; ssd += d * d; }
  return values.size() > 1 ? ssd / (n - 1.0) : 0.0;
}

// Rust f64::min / f64::max: a NaN operand is ignored.
inline double rust_fmin(double a, double b) { return std::fmin(a, b); }
inline double rust_fmax(double a, double b) { return std::fmax(a, b); }

// aggregation.rs:500-754.  ok=false stands for Err(OperationFailed).
// ORC_MODE_EXACT: long double accumulation of the same quantities (Int64 sums stay exact integers)
double calc_agg_exact(const orc_col& col, int op, const std::vector<size_t>& rows) {
  auto get = [&](size_t i) -> long double {
    switch (col.dtype) {
      case ORC_I64: return (long double)((const int64_t*)col.data)[i];
      case ORC_I32: return (long double)((const int32_t*)col.data)[i];
      default: return (long double)((const double*)col.data)[i];
    }
  };
  long double s = 0.0L; int64_t c = 0;
  for (size_t i : rows) if (!col_is_null(col, i)) { s += get(i); c++; }
  if (c == 0) return 0.0;
  if (op == ORC_SUM) return (double)s;
  const long double mean = s / (long double)c;
  if (op == ORC_MEAN) return (double)mean;
  long double ssd = 0.0L;
  for (size_t i : rows) if (!col_is_null(col, i)) { const long double d = get(i) - mean; ssd += d * d; }
  const long double var = c > 1 ? ssd / (long double)(c - 1) : 0.0L;
  return op == ORC_STD ? (double)sqrtl(var) : (double)var;
}

double calc_agg(const orc_col& col, int op, const std::vector<size_t>& rows, bool lazy, bool* ok, bool exact = false) {
  *ok = true;
  if (op == ORC_COUNT) return (double)rows.size();             // :743  (group size, NULLs included)
  if (exact && (op == ORC_SUM || op == ORC_MEAN || op == ORC_STD || op == ORC_VAR) && col.dtype == ORC_F64) return calc_agg_exact(col, op, rows);
  if (exact && (op == ORC_STD || op == ORC_VAR) && (col.dtype == ORC_I64 || col.dtype == ORC_I32)) return calc_agg_exact(col, op, rows);
  if (lazy && (op == ORC_STD || op == ORC_VAR)) { *ok = false; return 0.0; }  // lazy.rs:377-382
  if (col.dtype == ORC_I64 || col.dtype == ORC_I32) {
    auto get = [&](size_t i) -> int64_t { return col.dtype == ORC_I64 ? ((const int64_t*)col.data)[i] : (int64_t)((const int32_t*)col.data)[i]; };
    switch (op) {
      case ORC_SUM: { uint64_t s = 0; for (size_t i : rows) if (!col_is_null(col, i)) s += (uint64_t)get(i); return (double)(int64_t)s; }  // :507-515 (wrapping, release build)
      case ORC_MEAN: { uint64_t s = 0; int64_t c = 0; for (size_t i : rows) if (!col_is_null(col, i)) { s += (uint64_t)get(i); c++; }
                       return c > 0 ? (double)(int64_t)s / (double)c : 0.0; }                                                         // :516-530
      case ORC_MIN: { int64_t m = INT64_MAX; for (size_t i : rows) if (!col_is_null(col, i)) m = std::min(m, get(i)); return m == INT64_MAX ? 0.0 : (double)m; }  // :531-543
      case ORC_MAX: { int64_t m = INT64_MIN; for (size_t i : rows) if (!col_is_null(col, i)) m = std::max(m, get(i)); return m == INT64_MIN ? 0.0 : (double)m; }  // :544-556
      case ORC_STD: case ORC_VAR: {
        std::vector<double> v; for (size_t i : rows) if (!col_is_null(col, i)) v.push_back((double)get(i));
        if (v.empty()) return 0.0; double var = calculate_variance(v); return op == ORC_STD ? std::sqrt(var) : var; }                // :557-584
      case ORC_MEDIAN: {                                                                                                            // :585-604
        std::vector<int64_t> v; for (size_t i : rows) if (!col_is_null(col, i)) v.push_back(get(i));
        if (v.empty()) return 0.0;
        std::sort(v.begin(), v.end());
        size_t mid = v.size() / 2;
        if (v.size() % 2 == 0) return (double)(int64_t)((uint64_t)v[mid - 1] + (uint64_t)v[mid]) / 2.0;   // i64 add (wrapping in a release build), then as f64 / 2.0
        return (double)v[mid]; }
      case ORC_FIRST: { if (rows.empty()) return 0.0; size_t i = rows.front(); return col_is_null(col, i) ? 0.0 : (double)get(i); }  // :605-614
      case ORC_LAST: { if (rows.empty()) return 0.0; size_t i = rows.back(); return col_is_null(col, i) ? 0.0 : (double)get(i); }    // :615-624
    }
  } else if (col.dtype == ORC_F64) {
    const double* d = (const double*)col.data;
    switch (op) {
      case ORC_SUM: { double s = 0.0; for (size_t i : rows) if (!col_is_null(col, i)) s += d[i]; return s; }                          // :625-633
      case ORC_MEAN: { double s = 0.0; int64_t c = 0; for (size_t i : rows) if (!col_is_null(col, i)) { s += d[i]; c++; } return c > 0 ? s / (double)c : 0.0; }  // :634-648
      case ORC_MIN: { double m = INFINITY; for (size_t i : rows) if (!col_is_null(col, i)) m = rust_fmin(m, d[i]); return m == INFINITY ? 0.0 : m; }   // :649-661
      case ORC_MAX: { double m = -INFINITY; for (size_t i : rows) if (!col_is_null(col, i)) m = rust_fmax(m, d[i]); return m == -INFINITY ? 0.0 : m; } // :662-674
      case ORC_STD: case ORC_VAR: {
        std::vector<double> v; for (size_t i : rows) if (!col_is_null(col, i)) v.push_back(d[i]);
        if (v.empty()) return 0.0; double var = calculate_variance(v); return op == ORC_STD ? std::sqrt(var) : var; }                // :675-702
      case ORC_MEDIAN: {                                                                                                            // :703-722
        std::vector<double> v; for (size_t i : rows) if (!col_is_null(col, i)) v.push_back(d[i]);
        if (v.empty()) return 0.0;
        std::stable_sort(v.begin(), v.end());           // sort_by(partial_cmp().unwrap_or(Equal)): a total order only without NaN
        size_t mid = v.size() / 2;
        if (v.size() % 2 == 0) return (v[mid - 1] + v[mid]) / 2.0;
        return v[mid]; }
      case ORC_FIRST: { if (rows.empty()) return 0.0; size_t i = rows.front(); return col_is_null(col, i) ? 0.0 : d[i]; }            // :723-732
      case ORC_LAST: { if (rows.empty()) return 0.0; size_t i = rows.back(); return col_is_null(col, i) ? 0.0 : d[i]; }              // :733-742
    }
  }
  *ok = false;  // :748-752  String/Boolean value columns support Count only
  return 0.0;
}

struct GroupByResult {
  std::vector<std::vector<std::string>> keys;   // per group: key tuple (strings)
  std::vector<int64_t> first_row;               // per group: smallest row index
  std::vector<int64_t> group_rows;              // per group: size
  std::vector<std::vector<double>> aggs;        // per aggregate: one f64 per group
  int error = 0;                                // 1 = OperationFailed
};

struct JoinResult { std::vector<int64_t> left, right; };

}  // namespace

extern "C" {

// grouping.rs:38-115 + aggregation.rs:763-871 (mode 0), :22-182 (mode 1), lazy.rs:186-404 (mode 2).
// Groups are emitted in first-appearance order (the reference's HashMap order is unspecified).
void* orc_groupby(const orc_col* keys, int nkeys, const orc_col* vals, const int32_t* agg_col,
                  const int32_t* agg_op, int naggs, int64_t nrows, int mode, int nthreads) {
  auto* res = new GroupByResult();
  std::unordered_map<std::vector<std::string>, size_t, VecStrHash> index;
  std::vector<std::vector<size_t>> groups;
  for (int64_t row = 0; row < nrows; row++) {                 // HOT LOOP 1, serial (grouping.rs:62-104)
    std::vector<std::string> key; key.reserve(nkeys);
    for (int k = 0; k < nkeys; k++) key.push_back(key_part(keys[k], row));
    auto it = index.find(key);
    size_t g;
    if (it == index.end()) { g = groups.size(); index.emplace(key, g); groups.emplace_back(); res->keys.push_back(std::move(key)); res->first_row.push_back(row); }
    else g = it->second;
    groups[g].push_back((size_t)row);
  }
  size_t G = groups.size();
  res->group_rows.resize(G);
  for (size_t g = 0; g < G; g++) res->group_rows[g] = (int64_t)groups[g].size();
  res->aggs.assign(naggs, std::vector<double>(G, 0.0));
  bool lazy = mode == ORC_MODE_LAZY;
  std::atomic<int> err{0};
  auto do_group = [&](size_t g) {                              // HOT LOOP 2 (aggregation.rs:802-807)
    for (int a = 0; a < naggs; a++) {
      bool ok; double v = calc_agg(vals[agg_col[a]], agg_op[a], groups[g], lazy, &ok, mode == ORC_MODE_EXACT);
      if (!ok) { if (mode == ORC_MODE_PAR_AGGREGATE) v = 0.0; else err = 1; }   // aggregation.rs:114-117 vs :748-752
      res->aggs[a][g] = v;
    }
  };
  // par_aggregate thresholds, aggregation.rs:36-46
  bool use_parallel = mode == ORC_MODE_PAR_AGGREGATE && nthreads > 1 && (G >= 10 || (nrows >= 10000 && G > 3));
  if (use_parallel) {
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) pool.emplace_back([&] { for (;;) { size_t g0 = next.fetch_add(64); if (g0 >= G) break; for (size_t g = g0; g < std::min(G, g0 + 64); g++) do_group(g); } });
    for (auto& th : pool) th.join();
  } else {
    for (size_t g = 0; g < G; g++) do_group(g);
  }
  res->error = err;
  return res;
}
int orc_gb_error(void* h) { return ((GroupByResult*)h)->error; }
int64_t orc_gb_ngroups(void* h) { return (int64_t)((GroupByResult*)h)->keys.size(); }
void orc_gb_first_rows(void* h, int64_t* out) { auto* r = (GroupByResult*)h; std::copy(r->first_row.begin(), r->first_row.end(), out); }
void orc_gb_group_rows(void* h, int64_t* out) { auto* r = (GroupByResult*)h; std::copy(r->group_rows.begin(), r->group_rows.end(), out); }
void orc_gb_agg(void* h, int a, double* out) { auto* r = (GroupByResult*)h; std::copy(r->aggs[a].begin(), r->aggs[a].end(), out); }
const char* orc_gb_key(void* h, int64_t g, int k) { return ((GroupByResult*)h)->keys[g][k].c_str(); }
void orc_gb_free(void* h) { delete (GroupByResult*)h; }

// grouping.rs:124-331 (par_groupby): label = parts joined with "_", NULL -> "NA" (:158-186); rows pushed in row order (the parallel
// branch merges its chunk maps in chunk order, :265-282, so the lists are ascending too).  Groups in first-appearance order.
struct ParGroups { std::vector<std::string> labels; std::vector<std::vector<int64_t>> rows; };
void* orc_par_groupby(const orc_col* keys, int nkeys, int64_t nrows) {
  auto* res = new ParGroups();
  std::unordered_map<std::string, size_t, StrHash> index;
  for (int64_t row = 0; row < nrows; row++) {
    std::string label;
    for (int k = 0; k < nkeys; k++) {
      if (k) label += "_";
      label += col_is_null(keys[k], row) ? std::string("NA") : key_part(keys[k], row);
    }
    auto it = index.find(label);
    size_t g;
    if (it == index.end()) { g = res->labels.size(); index.emplace(label, g); res->labels.push_back(label); res->rows.emplace_back(); }
    else g = it->second;
    res->rows[g].push_back(row);
  }
  return res;
}
int64_t orc_pg_ngroups(void* h) { return (int64_t)((ParGroups*)h)->labels.size(); }
const char* orc_pg_label(void* h, int64_t g) { return ((ParGroups*)h)->labels[g].c_str(); }
int64_t orc_pg_size(void* h, int64_t g) { return (int64_t)((ParGroups*)h)->rows[g].size(); }
void orc_pg_rows(void* h, int64_t g, int64_t* out) { auto& r = ((ParGroups*)h)->rows[g]; std::copy(r.begin(), r.end(), out); }
void orc_pg_free(void* h) { delete (ParGroups*)h; }

// join.rs:107-224.  Pairs in reference order: left-row-major, matches in ascending right row,
// then (Right/Outer) unmatched right rows.  -1 stands for None.
void* orc_join(const orc_col* left, const orc_col* right, int how) {
  auto* res = new JoinResult();
  std::unordered_map<std::string, std::vector<size_t>, StrHash> right_key_to_indices;
  for (int64_t i = 0; i < right->len; i++)                           // BUILD :107-142 (NULL keys skipped)
    if (!col_is_null(*right, i)) right_key_to_indices[key_part(*right, i)].push_back((size_t)i);
  for (int64_t i = 0; i < left->len; i++) {                          // PROBE :146-208
    if (col_is_null(*left, i)) continue;                             // NULL left keys dropped even for Left (:152)
    auto it = right_key_to_indices.find(key_part(*left, i));
    if (it != right_key_to_indices.end()) { for (size_t r : it->second) { res->left.push_back(i); res->right.push_back((int64_t)r); } }
    else if (how == ORC_LEFT || how == ORC_OUTER) { res->left.push_back(i); res->right.push_back(-1); }
  }
  if (how == ORC_RIGHT || how == ORC_OUTER) {                        // :211-224
    std::vector<char> matched(right->len, 0);
    for (int64_t r : res->right) if (r >= 0) matched[r] = 1;
    for (int64_t i = 0; i < right->len; i++) if (!matched[i]) { res->left.push_back(-1); res->right.push_back(i); }
  }
  return res;
}
int64_t orc_join_len(void* h) { return (int64_t)((JoinResult*)h)->left.size(); }
void orc_join_pairs(void* h, int64_t* l, int64_t* r) { auto* j = (JoinResult*)h; std::copy(j->left.begin(), j->left.end(), l); std::copy(j->right.begin(), j->right.end(), r); }
void orc_join_free(void* h) { delete (JoinResult*)h; }

// join.rs:290-361,475-552 / data_ops.rs:124-211: gather with the type default for a missing side or a
// NULL source value; the output carries no null mask.  out is i64/f64/u32/u8(bool as byte) by dtype.
void orc_gather(const orc_col* col, const int64_t* idx, int64_t n, void* out) {
  for (int64_t j = 0; j < n; j++) {
    int64_t i = idx[j];
    bool missing = i < 0 || col_is_null(*col, i);
    switch (col->dtype) {
      case ORC_I64: ((int64_t*)out)[j] = missing ? 0 : ((const int64_t*)col->data)[i]; break;
      case ORC_F64: ((double*)out)[j] = missing ? 0.0 : ((const double*)col->data)[i]; break;
      case ORC_I32: ((int32_t*)out)[j] = missing ? 0 : ((const int32_t*)col->data)[i]; break;
      case ORC_DICT_U32: ((uint32_t*)out)[j] = missing ? 0xFFFFFFFFu : ((const uint32_t*)col->data)[i]; break;  // 0xFFFFFFFF = "" (empty string)
      case ORC_BOOL_BITS: ((uint8_t*)out)[j] = missing ? 0 : ((((const uint8_t*)col->data)[i / 8] >> (i % 8)) & 1); break;
    }
  }
}

// data_ops.rs:37-62: rows where the Boolean column is Some(true).
int64_t orc_filter_indices(const orc_col* mask, int64_t* out) {
  int64_t n = 0; const uint8_t* b = (const uint8_t*)mask->data;
  for (int64_t i = 0; i < mask->len; i++) if (!col_is_null(*mask, i) && ((b[i / 8] >> (i % 8)) & 1)) out[n++] = i;
  return n;
}

// "Idealised CPU" comparator: typed i64 keys in per-thread flat hash tables, one thread
// per core, merged at the end.  NOT the reference algorithm; reported beside it for fairness only.
// Computes sum/count/min/max/sumsq of one f64 column keyed by one i64 column (no NULL keys).
double orc_ideal_groupby_checksum(const int64_t* keys, const double* vals, const uint8_t* vnull, int64_t n, int nthreads, int64_t* ngroups_out) {
  struct Acc { int64_t key; int64_t rows, cnt; double sum, sq, mn, mx; bool used; };
  auto run = [&](int64_t lo, int64_t hi, std::vector<Acc>& tab) {
    size_t cap = 1 << 12; tab.assign(cap, Acc{0, 0, 0, 0, 0, INFINITY, -INFINITY, false}); size_t used = 0;
    auto insert = [&](std::vector<Acc>& t, int64_t k) -> Acc& { size_t m = t.size() - 1; size_t s = ((uint64_t)k * 0x9E3779B97F4A7C15ULL >> 20) & m; while (t[s].used && t[s].key != k) s = (s + 1) & m; return t[s]; };
    for (int64_t i = lo; i < hi; i++) {
      if (used * 2 > cap) { std::vector<Acc> nt(cap * 2, Acc{0, 0, 0, 0, 0, INFINITY, -INFINITY, false}); for (auto& a : tab) if (a.used) insert(nt, a.key) = a; tab.swap(nt); cap *= 2; }
      Acc& a = insert(tab, keys[i]);
      if (!a.used) { a.used = true; a.key = keys[i]; used++; }
      a.rows++;
      if (!(vnull && (vnull[i / 8] >> (i % 8) & 1))) { double v = vals[i]; a.cnt++; a.sum += v; a.sq += v * v; a.mn = std::fmin(a.mn, v); a.mx = std::fmax(a.mx, v); }
    }
  };
  std::vector<std::vector<Acc>> tabs(nthreads);
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++) pool.emplace_back([&, t] { run(n * t / nthreads, n * (t + 1) / nthreads, tabs[t]); });
  for (auto& th : pool) th.join();
  std::unordered_map<int64_t, Acc> merged;
  for (auto& tab : tabs) for (auto& a : tab) if (a.used) { auto& m = merged[a.key]; if (!m.used) m = a; else { m.rows += a.rows; m.cnt += a.cnt; m.sum += a.sum; m.sq += a.sq; m.mn = std::fmin(m.mn, a.mn); m.mx = std::fmax(m.mx, a.mx); } }
  double cs = 0; for (auto& kv : merged) cs += kv.second.sum;
  *ngroups_out = (int64_t)merged.size();
  return cs;
}

// Counter-based synthetic generators shared with the CUDA side (pandrs_b200/csrc/synth.cu uses the
// same arithmetic), so CPU oracle and GPU see identical inputs without shipping data.
static inline uint64_t splitmix64(uint64_t x) { x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31); }
uint64_t orc_splitmix64(uint64_t x) { return splitmix64(x); }
void orc_synth_keys(int64_t* out, int64_t n, int64_t row0, uint64_t seed, uint64_t card, int scramble) {
  for (int64_t i = 0; i < n; i++) { uint64_t k = splitmix64(seed * 0x100000001B3ULL + (uint64_t)(row0 + i)) % card; out[i] = (int64_t)(scramble ? splitmix64(k ^ 0xA5A5A5A5DEADBEEFULL) : k); }
}
void orc_synth_vals(double* out, int64_t n, int64_t row0, uint64_t seed) {
  for (int64_t i = 0; i < n; i++) { uint64_t r = splitmix64((seed + 1) * 0x100000001B3ULL + (uint64_t)(row0 + i)); out[i] = (double)(r >> 11) * (1000.0 / 9007199254740992.0); }
}
void orc_synth_join_keys(int64_t* out, int64_t n, int64_t row0, uint64_t seed, uint64_t domain, int unique) {
  for (int64_t i = 0; i < n; i++) { uint64_t id = unique ? (uint64_t)(row0 + i) : splitmix64((seed + 3) * 0x100000001B3ULL + (uint64_t)(row0 + i)) % domain; out[i] = (int64_t)(id * 0x9E3779B97F4A7C15ULL); }
}
void orc_synth_nulls(uint8_t* out, int64_t n, int64_t row0, uint64_t seed, uint32_t per_million) {   // n, row0 multiples of 8
  for (int64_t b = 0; b < (n + 7) / 8; b++) { uint8_t byte = 0; for (int j = 0; j < 8 && b * 8 + j < n; j++) { uint64_t r = splitmix64((seed + 2) * 0x100000001B3ULL + (uint64_t)(row0 + b * 8 + j)); if ((uint32_t)(r % 1000000ULL) < per_million) byte |= (uint8_t)(1u << j); } out[b] = byte; }
}

}  // extern "C"
