// Finalisation shared by the table scan (gb_finalize_kernel, groupby.cu) and the kernels that write finished groups straight to the
// result (gb_part.cu): key decoding + the reference's aggregate formulas (aggregation.rs:500-754).
#pragma once
#include "groupby_kernels.cuh"

struct FinVal { const GState* st; int is_int; int flags; long long* validn_out; u64* states_out; };
struct FinAgg { int val; int op; double* out; };
struct FinParams {
  GTable gt;
  KeySpec ks;
  void* key_out[PDRS_MAX_KEYS];
  uint8_t* key_null_out[PDRS_MAX_KEYS];
  long long* rows_out;
  int nvals, naggs;
  FinVal vals[PDRS_MAX_VALS];
  FinAgg aggs[PDRS_MAX_AGGS];
};

__device__ __forceinline__ double fin_pivot(u64 px) { return px ? __longlong_as_double((long long)(px ^ GB_PIV_X)) : 0.0; }

// aggregation.rs:500-754: every aggregate is an f64; empty / all-NULL -> 0.0; Min/Max sentinel collapse.
__device__ __forceinline__ double fin_eval(const GState& s, int is_int, int op, u64 rows) {
  const double n = (double)s.n;
  const double c = fin_pivot(s.pivotx);
  switch (op) {
    case PDRS_COUNT: return (double)rows;                                     // :743 group size, NULLs included
    case PDRS_SUM:
      if (s.n == 0) return 0.0;
      return is_int ? (double)(long long)s.isum : (s.S1 + n * c);              // :507-515 / :625-633
    case PDRS_MEAN:
      if (s.n == 0) return 0.0;
      return is_int ? (double)(long long)s.isum / n : (c + s.S1 / n);          // :516-530 / :634-648
    case PDRS_MIN: {
      if (s.mnc == 0) return 0.0;
      u64 o = ~s.mnc;
      if (is_int) return (double)(long long)(o ^ GB_SIGN);                     // :531-543 (i64::MAX never stored: collapses to 0.0)
      double v = pdrs_unord_f64(o);
      return v == __longlong_as_double(0x7FF0000000000000ll) ? 0.0 : v;        // :649-661 (min == +INF -> 0.0)
    }
    case PDRS_MAX: {
      if (s.mxo == 0) return 0.0;
      if (is_int) return (double)(long long)(s.mxo ^ GB_SIGN);                 // :544-556
      { double v = pdrs_unord_f64(s.mxo); return v == __longlong_as_double((long long)0xFFF0000000000000ull) ? 0.0 : v; }   // :662-674 (max == -INF -> 0.0)
    }
    case PDRS_STD: case PDRS_VAR: {                                            // :557-584 / :675-702 / :881-903
      if (s.n <= 1) return 0.0;
      double var = (s.S2 - s.S1 * s.S1 / n) / (n - 1.0);
      if (var < 0.0) var = 0.0;
      return op == PDRS_STD ? sqrt(var) : var;
    }
  }
  return 0.0;
}

// Writes group `o` of the result: decoded key parts, group size, and - per value column - valid count / mergeable state / aggregates.
// st_of(v) returns the state of value column v (or nullptr: no state, only Count is defined).
template <class StateOf>
__device__ __forceinline__ void fin_write_group(const FinParams& p, long long o, const u64 (&w)[PDRS_MAX_WORDS], bool nullgroup, u64 rows, StateOf st_of) {
  for (int k = 0; k < p.ks.nkeys; k++) {
    const KeyColDev& c = p.ks.c[k];
    bool isnull = nullgroup;
    if (!nullgroup && c.nword >= 0) isnull = (w[c.nword] >> c.nshift) & 1;
    u64 v = 0;
    if (!isnull) { v = w[c.word] >> c.shift; if (c.bits < 64) v &= (1ull << c.bits) - 1ull; v += (u64)c.offset; }
    switch (c.dtype) {
      case PDRS_I64: case PDRS_F64: reinterpret_cast<u64*>(p.key_out[k])[o] = v; break;
      case PDRS_I32: case PDRS_DICT_U32: reinterpret_cast<uint32_t*>(p.key_out[k])[o] = (uint32_t)v; break;
      case PDRS_BOOL_BITS: reinterpret_cast<uint8_t*>(p.key_out[k])[o] = (uint8_t)v; break;
    }
    p.key_null_out[k][o] = isnull ? 1 : 0;
  }
  p.rows_out[o] = (long long)rows;
  for (int v = 0; v < p.nvals; v++) {
    const GState* sp = st_of(v);
    GState st;
    if (sp) st = *sp; else { st.n = 0; st.pivotx = 0; st.S1 = 0; st.S2 = 0; st.mnc = 0; st.mxo = 0; st.isum = 0; st.pad = 0; }
    if (p.vals[v].validn_out) p.vals[v].validn_out[o] = (long long)st.n;
    if (p.vals[v].states_out) {
      u64* q = p.vals[v].states_out + 8 * o;
      q[0] = rows; q[1] = st.n; q[2] = st.pivotx; q[3] = (u64)__double_as_longlong(st.S1);
      q[4] = (u64)__double_as_longlong(st.S2); q[5] = st.mnc; q[6] = st.mxo; q[7] = st.isum;
    }
  }
  for (int a = 0; a < p.naggs; a++) {
    const FinAgg& ag = p.aggs[a];
    const GState* sp = ag.val >= 0 ? st_of(ag.val) : nullptr;
    if (!sp) { ag.out[o] = ag.op == PDRS_COUNT ? (double)rows : 0.0; continue; }
    ag.out[o] = fin_eval(*sp, p.vals[ag.val].is_int, ag.op, rows);
  }
}
