// pandrs CPU ORACLE, typed-key variant — TEST INFRASTRUCTURE ONLY (same rules as pandrs_oracle.cpp).
//
// The string oracle (pandrs_oracle.cpp) restates the reference literally: `to_string()` keys, one
// `HashMap<Vec<String>, Vec<usize>>`, serial grouping.  That is ~10 M rows/s and cannot check the CUDA path at
// the sizes BASELINE.json names (1e8 - 1e9 rows).  This file computes THE SAME RESULTS with typed keys:
//
//   * equal key strings <=> equal typed tuples: i64 / i32 / dictionary ids / bools print bijectively
//     (grouping.rs:69-98); f64 keys compare by Display text = by bit pattern with all NaNs equal; a NULL part
//     prints "NULL", and a dictionary id whose string is literally "NULL" merges with it (null_alias).
//   * every group is owned by ONE thread (hash(key) mod threads) that walks all rows in ascending order, so the
//     per-group arithmetic is the reference's, operation for operation: sequential `+=` in row order
//     (aggregation.rs:507-515, 625-633), `sum / count` (:516-530, 634-648), f64::min / max folds with NaN
//     operands ignored and the sentinel collapse to 0.0 (:531-556, 649-674), Count = group size (:743) and the
//     two-pass variance of calculate_variance (:881-903): mean = sum / n first, then sum((x - mean)^2) / (n - 1)
//     in a second walk over the rows.  Results are therefore BIT-IDENTICAL to pandrs_oracle.cpp, which
//     tests/test_oracle_golden.py asserts on random inputs (that is what pins this file).
//   * rows can come from arrays (any inputs) or straight from the counter-based generators shared with the CUDA
//     side (orc_synth_*): 1e9 synthetic rows need no host memory.
//
// Join: typed build table (key -> ascending list of right rows, join.rs:107-142), probe in left-row order
// (join.rs:146-208).  Besides the pairs it returns an order-independent checksum of the pair multiset, so that a
// 1e9 x 1e8 join can be compared with the CUDA path without sorting 5e8 pairs on the host.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

extern "C" {
enum { TO_I64 = 0, TO_F64 = 1, TO_DICT_U32 = 2, TO_BOOL_BITS = 3, TO_I32 = 4 };
typedef struct {
  int32_t dtype;
  int32_t _pad;
  const void* data;
  const uint8_t* null_bits;
  int64_t null_len;
  int64_t len;
  const char* const* pool;
  int64_t pool_len;
} orc_col;   // identical to pandrs_oracle.cpp
}

namespace {

inline uint64_t splitmix64(uint64_t x) { x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31); }

inline bool col_is_null(const orc_col& c, int64_t i) {      // int64_column.rs:72-81 (short masks: not NULL)
  if (!c.null_bits) return false;
  const int64_t b = i >> 3;
  return b < c.null_len && ((c.null_bits[b] >> (i & 7)) & 1);
}

constexpr int MAXK = 4;
struct Key { uint64_t w[MAXK]; uint32_t nullmask; };
inline bool key_eq(const Key& a, const Key& b, int nk) {
  if (a.nullmask != b.nullmask) return false;
  for (int k = 0; k < nk; k++) if (a.w[k] != b.w[k]) return false;
  return true;
}
inline uint64_t key_hash(const Key& a, int nk) {
  uint64_t h = 0x243F6A8885A308D3ULL + a.nullmask;
  for (int k = 0; k < nk; k++) h = splitmix64(h ^ a.w[k]);
  return h;
}

struct Acc {
  Key key;
  int64_t first_row, rows, n;
  double sum;          // f64 values: sequential f64 sum; i64 values: (unused)
  uint64_t isum;       // i64 values: wrapping sum
  double fsum;         // i64 values: sequential sum of `v as f64` (calculate_variance input)
  double mn, mx;
  int64_t imn, imx;
  double mean, ssd;
  long double xsum, xssd, xsum2;   // (xsum2 = the 80-bit mean) NOT reference arithmetic: the same sums in 80-bit, so that callers can measure the reference's rounding error
  bool used;
};

struct Table {
  std::vector<Acc> t;
  size_t used = 0;
  int nk;
  explicit Table(int nkeys) : nk(nkeys) { t.resize(1 << 10); for (auto& a : t) a.used = false; }
  Acc* find(const Key& k, uint64_t h) {
    const size_t m = t.size() - 1;
    size_t s = (h >> 7) & m;
    while (t[s].used && !key_eq(t[s].key, k, nk)) s = (s + 1) & m;
    return &t[s];
  }
  void grow() {
    std::vector<Acc> old;
    old.swap(t);
    t.resize(old.size() * 2);
    for (auto& a : t) a.used = false;
    for (auto& a : old) if (a.used) *find(a.key, key_hash(a.key, nk)) = a;
  }
  Acc* upsert(const Key& k, uint64_t h, int64_t row) {
    if ((used + 1) * 10 > t.size() * 6) grow();
    Acc* a = find(k, h);
    if (!a->used) {
      a->used = true; a->key = k; a->first_row = row; a->rows = 0; a->n = 0; a->sum = 0.0; a->isum = 0; a->fsum = 0.0;
      a->mn = INFINITY; a->mx = -INFINITY; a->imn = INT64_MAX; a->imx = INT64_MIN; a->mean = 0.0; a->ssd = 0.0; a->xsum = 0.0L; a->xssd = 0.0L;
      used++;
    }
    return a;
  }
};

// ---- row sources -------------------------------------------------------------------------------------------
struct ArraySource {
  const orc_col* keys; const int64_t* null_alias; int nkeys;
  const orc_col* val; const orc_col* filter;
  int compat_nulls; int64_t empty_id;
  int64_t n;
  bool val_int;
  inline bool keep(int64_t i) const {          // data_ops.rs:49-55: Some(true) rows only
    if (!filter) return true;
    if (col_is_null(*filter, i)) return false;
    return (((const uint8_t*)filter->data)[i >> 3] >> (i & 7)) & 1;
  }
  inline void key(int64_t i, Key& k) const {
    k.nullmask = 0;
    for (int c = 0; c < MAXK; c++) k.w[c] = 0;
    for (int c = 0; c < nkeys; c++) {
      const orc_col& col = keys[c];
      bool isnull = col_is_null(col, i);
      uint64_t v = 0;
      if (!isnull) {
        switch (col.dtype) {
          case TO_I64: v = (uint64_t)((const int64_t*)col.data)[i]; break;
          case TO_I32: v = (uint64_t)(int64_t)((const int32_t*)col.data)[i]; break;
          case TO_DICT_U32: v = ((const uint32_t*)col.data)[i]; if (null_alias && (int64_t)v == null_alias[c]) isnull = true; break;
          case TO_F64: { double d = ((const double*)col.data)[i]; if (d != d) v = 0x7FF8000000000000ULL; else memcpy(&v, &d, 8); break; }
          case TO_BOOL_BITS: v = (((const uint8_t*)col.data)[i >> 3] >> (i & 7)) & 1; break;
        }
      } else if (filter && compat_nulls) {
        // data_ops.rs:64-71 / parallel.rs:177-231: filter() replaced the NULL by the type default before the groupby saw it
        isnull = false;
        v = col.dtype == TO_DICT_U32 ? (uint64_t)empty_id : 0;      // 0, 0.0 (bits 0), false, ""
      }
      if (isnull) k.nullmask |= 1u << c; else k.w[c] = v;
    }
  }
  // value of row i: returns false when NULL
  inline bool value(int64_t i, double* f, int64_t* iv) const {
    if (col_is_null(*val, i)) {
      if (filter && compat_nulls) { *f = 0.0; *iv = 0; return true; }
      return false;
    }
    if (val_int) { *iv = val->dtype == TO_I64 ? ((const int64_t*)val->data)[i] : (int64_t)((const int32_t*)val->data)[i]; *f = (double)*iv; }
    else { *f = ((const double*)val->data)[i]; *iv = 0; }
    return true;
  }
};

// a % d for a fixed d without a division (Lemire, Kaser, Kurz 2019): M = ceil(2^128 / d)
struct FastMod {
  unsigned __int128 M; uint64_t d;
  explicit FastMod(uint64_t dd = 1) : M(~(unsigned __int128)0 / dd + 1), d(dd) {}
  inline uint64_t mod(uint64_t a) const {
    if (d == 1) return 0;
    const unsigned __int128 low = M * a;
    const uint64_t lo = (uint64_t)low, hi = (uint64_t)(low >> 64);
    return (uint64_t)((((unsigned __int128)lo * d >> 64) + (unsigned __int128)hi * d) >> 64);
  }
};

struct SynthSource {      // BASELINE.json configs[1]: orc_synth_keys / _vals / _nulls without materialising them
  int64_t n; uint64_t seed, card; int scramble; uint32_t per_million; int64_t row0; FastMod fm, fmillion{1000000};
  const orc_col* val = nullptr; const orc_col* filter = nullptr;
  int nkeys = 1;
  bool val_int = false;
  inline bool keep(int64_t) const { return true; }
  inline void key(int64_t i, Key& k) const {
    uint64_t v = fm.mod(splitmix64(seed * 0x100000001B3ULL + (uint64_t)(row0 + i)));
    if (scramble) v = splitmix64(v ^ 0xA5A5A5A5DEADBEEFULL);
    k.nullmask = 0; k.w[0] = v; k.w[1] = k.w[2] = k.w[3] = 0;
  }
  inline bool value(int64_t i, double* f, int64_t* iv) const {
    *iv = 0;
    if (per_million) {
      const uint64_t r = splitmix64((seed + 2) * 0x100000001B3ULL + (uint64_t)(row0 + i));
      if ((uint32_t)fmillion.mod(r) < per_million) return false;
    }
    const uint64_t r = splitmix64((seed + 1) * 0x100000001B3ULL + (uint64_t)(row0 + i));
    *f = (double)(r >> 11) * (1000.0 / 9007199254740992.0);
    return true;
  }
};

struct TGResult {
  int nkeys = 0;
  std::vector<Acc> groups;
  bool val_int = false, has_val = false;
};

template <class Src>
TGResult* typed_groupby(const Src& src, bool has_val, int nthreads) {
  auto* res = new TGResult();
  res->nkeys = src.nkeys; res->val_int = src.val_int; res->has_val = has_val;
  if (nthreads < 1) nthreads = 1;
  const int nk = src.nkeys;
  std::vector<Table*> tabs(nthreads);
  std::vector<std::thread> pool;
  const bool val_int = src.val_int;
  for (int t = 0; t < nthreads; t++) {
    pool.emplace_back([&, t] {
      Table* tab = new Table(nk);
      tabs[t] = tab;
      Key k;
      // Owned rows go through a short FIFO so that the table slot of a row is prefetched a few rows before it is
      // updated (the tables of a 1e7-group run do not fit any cache); FIFO order keeps the rows in ascending order.
      constexpr int RING = 16;
      struct Pend { Key k; uint64_t h; int64_t i; };
      Pend ring[RING];
      int head = 0, cnt = 0;
      // pass 1 (grouping.rs:62-104 + aggregation.rs:507-674): rows in ascending order, this thread's keys only
      auto apply1 = [&](const Pend& p) {
        Acc* a = tab->upsert(p.k, p.h, p.i);
        a->rows++;
        if (!has_val) return;
        double f; int64_t iv;
        if (!src.value(p.i, &f, &iv)) return;
        a->n++;
        a->xsum += (long double)f;
        if (val_int) { a->isum += (uint64_t)iv; a->fsum += f; a->imn = std::min(a->imn, iv); a->imx = std::max(a->imx, iv); }
        else { a->sum += f; a->mn = std::fmin(a->mn, f); a->mx = std::fmax(a->mx, f); }
      };
      for (int64_t i = 0; i < src.n; i++) {
        if (!src.keep(i)) continue;
        src.key(i, k);
        const uint64_t h = key_hash(k, nk);
        if ((int)(((h >> 32) * (uint64_t)nthreads) >> 32) != t) continue;      // owner thread of the key (no division)
        if (cnt == RING) { apply1(ring[head]); head = (head + 1) % RING; cnt--; }
        Pend& p = ring[(head + cnt) % RING];
        p.k = k; p.h = h; p.i = i; cnt++;
        __builtin_prefetch(&tab->t[(h >> 7) & (tab->t.size() - 1)], 1);
      }
      for (; cnt; cnt--) { apply1(ring[head]); head = (head + 1) % RING; }
      if (!has_val) return;
      // calculate_variance (aggregation.rs:881-903): mean = (sequential sum of the f64 values) / n, then a second walk
      for (auto& a : tab->t) if (a.used && a.n > 0) { a.mean = (val_int ? a.fsum : a.sum) / (double)a.n; a.xsum2 = a.xsum / (long double)a.n; }
      auto apply2 = [&](const Pend& p) {
        double f; int64_t iv;
        if (!src.value(p.i, &f, &iv)) return;
        Acc* a = tab->find(p.k, p.h);
        const double d = f - a->mean;
        a->ssd += d * d;
        const long double xd = (long double)f - a->xsum2;
        a->xssd += xd * xd;
      };
      head = 0;
      for (int64_t i = 0; i < src.n; i++) {
        if (!src.keep(i)) continue;
        src.key(i, k);
        const uint64_t h = key_hash(k, nk);
        if ((int)(((h >> 32) * (uint64_t)nthreads) >> 32) != t) continue;
        if (cnt == RING) { apply2(ring[head]); head = (head + 1) % RING; cnt--; }
        Pend& p = ring[(head + cnt) % RING];
        p.k = k; p.h = h; p.i = i; cnt++;
        __builtin_prefetch(&tab->t[(h >> 7) & (tab->t.size() - 1)], 1);
      }
      for (; cnt; cnt--) { apply2(ring[head]); head = (head + 1) % RING; }
    });
  }
  for (auto& th : pool) th.join();
  size_t total = 0;
  for (Table* tab : tabs) total += tab->used;
  res->groups.reserve(total);
  for (Table* tab : tabs) { for (auto& a : tab->t) if (a.used) res->groups.push_back(a); delete tab; }
  return res;
}

}  // namespace

extern "C" {

// keys[nkeys] (+ null_alias[nkeys] or NULL), one value column (may be NULL: counts only), optional Boolean filter.
void* orc_typed_groupby(const orc_col* keys, const int64_t* null_alias, int nkeys, const orc_col* val, const orc_col* filter,
                        int compat_nulls, int64_t empty_id, int nthreads) {
  ArraySource s{keys, null_alias, nkeys, val, filter, compat_nulls, empty_id, keys[0].len, val && (val->dtype == TO_I64 || val->dtype == TO_I32)};
  return typed_groupby(s, val != nullptr, nthreads);
}
void* orc_typed_groupby_synth(int64_t n, int64_t row0, uint64_t seed, uint64_t card, int scramble, uint32_t null_per_million, int nthreads) {
  SynthSource s{n, seed, card, scramble, null_per_million, row0, FastMod(card)};
  return typed_groupby(s, true, nthreads);
}
int64_t orc_tg_ngroups(void* h) { return (int64_t)((TGResult*)h)->groups.size(); }
void orc_tg_key(void* h, int k, uint64_t* vals, uint8_t* isnull) {
  auto* r = (TGResult*)h;
  for (size_t g = 0; g < r->groups.size(); g++) { vals[g] = r->groups[g].key.w[k]; isnull[g] = (r->groups[g].key.nullmask >> k) & 1; }
}
void orc_tg_rows(void* h, int64_t* rows, int64_t* valid_n, int64_t* first_row) {
  auto* r = (TGResult*)h;
  for (size_t g = 0; g < r->groups.size(); g++) { rows[g] = r->groups[g].rows; valid_n[g] = r->groups[g].n; first_row[g] = r->groups[g].first_row; }
}
// sum, mean, min, max, std, var per group with the reference's formulas and edge cases
void orc_tg_aggs(void* h, double* sum, double* mean, double* mn, double* mx, double* sd, double* var) {
  auto* r = (TGResult*)h;
  for (size_t g = 0; g < r->groups.size(); g++) {
    const Acc& a = r->groups[g];
    if (!r->has_val || a.n == 0) { sum[g] = mean[g] = mn[g] = mx[g] = sd[g] = var[g] = 0.0; continue; }
    if (r->val_int) {
      sum[g] = (double)(int64_t)a.isum;
      mean[g] = (double)(int64_t)a.isum / (double)a.n;
      mn[g] = a.imn == INT64_MAX ? 0.0 : (double)a.imn;
      mx[g] = a.imx == INT64_MIN ? 0.0 : (double)a.imx;
    } else {
      sum[g] = a.sum;
      mean[g] = a.sum / (double)a.n;
      mn[g] = a.mn == INFINITY ? 0.0 : a.mn;
      mx[g] = a.mx == -INFINITY ? 0.0 : a.mx;
    }
    const double v = a.n > 1 ? a.ssd / ((double)a.n - 1.0) : 0.0;
    var[g] = v; sd[g] = std::sqrt(v);
  }
}
// the same sum / mean / std / var from 80-bit accumulators (two-pass variance around the 80-bit mean)
void orc_tg_exact(void* h, double* sum, double* mean, double* sd, double* var) {
  auto* r = (TGResult*)h;
  for (size_t g = 0; g < r->groups.size(); g++) {
    const Acc& a = r->groups[g];
    if (!r->has_val || a.n == 0) { sum[g] = mean[g] = sd[g] = var[g] = 0.0; continue; }
    sum[g] = (double)a.xsum; mean[g] = (double)(a.xsum / (long double)a.n);
    const long double v = a.n > 1 ? a.xssd / (long double)(a.n - 1) : 0.0L;
    var[g] = (double)v; sd[g] = (double)sqrtl(v);
  }
}
void orc_tg_free(void* h) { delete (TGResult*)h; }

// ---------------------------------------------------------------------------------------------------- join
// Typed restatement of join.rs:107-224 for one integer-like key column (I64 / I32 / DICT_U32 / BOOL / F64 bits).
// left / right may be NULL: the side then comes from orc_synth_join_keys(n, row0 = 0, seed, domain, unique).
struct JSynth { int64_t n; uint64_t seed, domain; int unique; };
}

namespace {
struct JSide {
  const orc_col* col; JSynth s;
  int64_t n() const { return col ? col->len : s.n; }
  inline bool key(int64_t i, uint64_t* k) const {
    if (!col) {
      const uint64_t id = s.unique ? (uint64_t)i : splitmix64((s.seed + 3) * 0x100000001B3ULL + (uint64_t)i) % s.domain;
      *k = id * 0x9E3779B97F4A7C15ULL;
      return true;
    }
    if (col_is_null(*col, i)) return false;
    switch (col->dtype) {
      case TO_I64: *k = (uint64_t)((const int64_t*)col->data)[i]; break;
      case TO_I32: *k = (uint64_t)(int64_t)((const int32_t*)col->data)[i]; break;
      case TO_DICT_U32: *k = ((const uint32_t*)col->data)[i]; break;
      case TO_F64: { double d = ((const double*)col->data)[i]; if (d != d) *k = 0x7FF8000000000000ULL; else memcpy(k, &d, 8); break; }
      default: *k = (((const uint8_t*)col->data)[i >> 3] >> (i & 7)) & 1; break;
    }
    return true;
  }
};
struct JSlot { uint64_t key; int64_t head, tail; };      // head / tail of the ascending chain of right rows; head < 0 = empty
struct JTable {
  std::vector<JSlot> t;
  size_t used = 0;
  JTable() { t.assign(1 << 10, JSlot{0, -1, -1}); }
  JSlot* find(uint64_t k) {
    const size_t m = t.size() - 1;
    size_t s = (splitmix64(k) >> 9) & m;
    while (t[s].head >= 0 && t[s].key != k) s = (s + 1) & m;
    return &t[s];
  }
  void grow() {
    std::vector<JSlot> old; old.swap(t);
    t.assign(old.size() * 2, JSlot{0, -1, -1});
    for (auto& e : old) if (e.head >= 0) *find(e.key) = e;
  }
};
struct TJResult { int64_t m = 0; uint64_t checksum = 0, checksum_unmatched = 0; int64_t sum_left = 0, sum_right = 0, unmatched_left = 0; std::vector<int64_t> left, right; };
inline uint64_t pair_mix(int64_t l, int64_t r) { return ((uint64_t)l * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)r * 0xC2B2AE3D27D4EB4FULL); }
}

extern "C" {
// how: 0 inner, 1 left (right / outer: the string oracle covers them).  want_pairs: also materialise the pairs in the
// reference's order.  checksum = sum over pairs of (l * A) ^ (r * B) mod 2^64, r = -1 for None.
void* orc_typed_join(const orc_col* left, const JSynth* lsynth, const orc_col* right, const JSynth* rsynth, int how, int want_pairs, int nthreads) {
  JSide L{left, lsynth ? *lsynth : JSynth{}}, R{right, rsynth ? *rsynth : JSynth{}};
  if (nthreads < 1) nthreads = 1;
  const int64_t nr = R.n(), nl = L.n();
  std::vector<int64_t> next((size_t)std::max<int64_t>(nr, 1), -1);
  std::vector<JTable> tabs(nthreads);
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) pool.emplace_back([&, t] {      // BUILD :107-142, this thread's keys, ascending right rows
      JTable& tab = tabs[t];
      for (int64_t i = 0; i < nr; i++) {
        uint64_t k;
        if (!R.key(i, &k)) continue;
        if ((int)(((splitmix64(k ^ 0x5bd1e995) >> 32) * (uint64_t)nthreads) >> 32) != t) continue;
        if ((tab.used + 1) * 10 > tab.t.size() * 6) tab.grow();
        JSlot* s = tab.find(k);
        if (s->head < 0) { s->key = k; s->head = s->tail = i; tab.used++; }
        else { next[s->tail] = i; s->tail = i; }
      }
    });
    for (auto& th : pool) th.join();
  }
  auto* res = new TJResult();
  std::vector<TJResult> part(nthreads);
  for (int pass = 0; pass < (want_pairs ? 2 : 1); pass++) {
    std::vector<int64_t> base(nthreads + 1, 0);
    if (pass == 1) { for (int t = 0; t < nthreads; t++) base[t + 1] = base[t] + part[t].m; res->left.resize(base[nthreads]); res->right.resize(base[nthreads]); }
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) pool.emplace_back([&, t, pass] {  // PROBE :146-208, a contiguous range of left rows
      TJResult p;
      int64_t at = base[t];
      for (int64_t i = nl * t / nthreads; i < nl * (t + 1) / nthreads; i++) {
        uint64_t k;
        if (!L.key(i, &k)) continue;                                      // NULL left keys are dropped, also for Left (:152)
        JTable& tab = tabs[(int)(((splitmix64(k ^ 0x5bd1e995) >> 32) * (uint64_t)nthreads) >> 32)];
        const JSlot* s = tab.find(k);
        if (s->head >= 0) {
          for (int64_t r = s->head; r >= 0; r = next[r]) {
            p.m++; p.checksum += pair_mix(i, r); p.sum_left += i; p.sum_right += r;
            if (pass == 1) { res->left[at] = i; res->right[at] = r; at++; }
          }
        } else if (how == 1) {
          p.m++; p.checksum += pair_mix(i, -1); p.checksum_unmatched += pair_mix(i, -1); p.sum_left += i; p.unmatched_left++;
          if (pass == 1) { res->left[at] = i; res->right[at] = -1; at++; }
        }
      }
      part[t] = p;
    });
    for (auto& th : pool) th.join();
  }
  for (auto& p : part) { res->m += p.m; res->checksum += p.checksum; res->checksum_unmatched += p.checksum_unmatched; res->sum_left += p.sum_left; res->sum_right += p.sum_right; res->unmatched_left += p.unmatched_left; }
  return res;
}
int64_t orc_tj_len(void* h) { return ((TJResult*)h)->m; }
void orc_tj_stats(void* h, uint64_t* checksum, uint64_t* checksum_unmatched, int64_t* sum_left, int64_t* sum_right, int64_t* unmatched_left) {
  auto* r = (TJResult*)h; *checksum = r->checksum; *checksum_unmatched = r->checksum_unmatched; *sum_left = r->sum_left; *sum_right = r->sum_right; *unmatched_left = r->unmatched_left;
}
void orc_tj_pairs(void* h, int64_t* l, int64_t* r) { auto* j = (TJResult*)h; std::copy(j->left.begin(), j->left.end(), l); std::copy(j->right.begin(), j->right.end(), r); }
void orc_tj_free(void* h) { delete (TJResult*)h; }

}  // extern "C"
