// Packing of the typed key arrays of a groupby RESULT (one entry per group + one NULL byte per group and key part) back into key
// words of a given KeySpec layout: used by the collective / chunked groupby (packed state rows) and by the row lists (gb_rows.cu:
// group number of a key tuple).
#pragma once
#include "groupby_kernels.cuh"

struct DistPack {
  KeySpec ks;                                   // layout only (data / nulls unused)
  const void* key_vals[PDRS_MAX_KEYS];          // typed key arrays of the partial result
  const uint8_t* key_null[PDRS_MAX_KEYS];       // one byte per group
  const long long* rows;
  const u64* states[PDRS_MAX_VALS];             // [G][8]
  int nvals;
  long long G, cap;                             // cap: rows that fit the destination (replicated mode)
  int stride;                                   // u64 words per packed row = 5 + 8 * nvals
  int world;
};

template <int NW>
__device__ __forceinline__ bool dist_row_words(const DistPack& p, long long j, u64 (&w)[NW]) {
#pragma unroll
  for (int i = 0; i < NW; i++) w[i] = 0;
  for (int k = 0; k < p.ks.nkeys; k++) {
    const KeyColDev& c = p.ks.c[k];
    const bool isnull = p.key_null[k][j] != 0;
    u64 v = 0;
    switch (c.dtype) {
      case PDRS_I64: v = reinterpret_cast<const u64*>(p.key_vals[k])[j]; break;
      case PDRS_F64: { v = reinterpret_cast<const u64*>(p.key_vals[k])[j]; const double d = __longlong_as_double((long long)v); if (d != d) v = 0x7FF8000000000000ull; break; }
      case PDRS_I32: v = (u64)(long long)reinterpret_cast<const int*>(p.key_vals[k])[j] & 0xFFFFFFFFull; break;
      case PDRS_DICT_U32: v = reinterpret_cast<const uint32_t*>(p.key_vals[k])[j]; break;
      default: v = reinterpret_cast<const uint8_t*>(p.key_vals[k])[j] ? 1ull : 0ull; break;
    }
    if (isnull) {
      if (p.ks.single_null) return true;
      if (c.nword >= 0) {
#pragma unroll
        for (int i = 0; i < NW; i++) if (i == c.nword) w[i] |= 1ull << c.nshift;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NW; i++) if (i == c.word) w[i] |= v << c.shift;
    }
  }
  return false;
}
