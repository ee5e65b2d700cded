// Interface of the few-groups / many-value-columns kernel (gb_few.cu).
#pragma once
#include "groupby_kernels.cuh"

#define GF_MAXV 8        // value columns per scan
#define GF_MAXG 16       // groups

struct GfParams {
  GbParams base;                       // keys, n, filter, compat_nulls, global table (st unused)
  int nv;
  const void* val[GF_MAXV];
  const uint8_t* vnull[GF_MAXV];
  GState* st[GF_MAXV];                 // state array of every value column in the global table
  int is_int[GF_MAXV];
  int any_vnull;
  int ng;
  u64 gkey[GF_MAXG];                   // packed key words of the groups the sample has seen
  // optional typed predicate `column <op> constant`, evaluated in the scan (NULL = not kept)
  const void* pcol; const uint8_t* pnull; int pdtype, pop; long long pival; double pfval;
};

size_t gb_few_smem(int ng, int nv, bool any_vnull);
cudaError_t gb_few_collect_keys(const GTable& sample, u64* out_dev, int ctas, cudaStream_t s);
cudaError_t gb_few_launch(const GfParams& p, int ctas, size_t smem, cudaStream_t s);
