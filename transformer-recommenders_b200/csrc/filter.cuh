// Survivor storage of the scoring filter (xr_score_filter -> xr_filter_finalize).
#pragma once

#include <stdint.h>

namespace xr {

struct FilterOut {
  float* b_scores;    // [u][n_sub][cap_b]   sub-bucket = (catalog split, column group) of the scoring kernel
  int32_t* b_rows;    // [u][n_sub][cap_b]   local catalog rows
  int32_t* b_count;   // [u][n_sub]          survivors its owner saw; beyond cap_b they went to the overflow list
  float* o_scores;    // [u][ovf_cap]        overflow list of the query (shared by its sub-buckets)
  int32_t* o_rows;    // [u][ovf_cap]
  int32_t* o_count;   // [u]                 zeroed by the caller; keeps counting past ovf_cap
  int n_sub, cap_b, ovf_cap;
};

// one survivor of the lane that owns the sub-bucket (b_scores / b_rows point at it, bcount = its fill count
// in a register): a plain store; only a full sub-bucket takes an atomic, on the query's overflow list
__device__ __forceinline__ void filter_keep(const FilterOut& fo, int row, float* b_scores, int32_t* b_rows,
                                            int& bcount, float score, int cat_row) {
  if (bcount < fo.cap_b) {
    b_scores[bcount] = score;
    b_rows[bcount] = cat_row;
  } else {
    const int slot = atomicAdd(fo.o_count + row, 1);
    if (slot < fo.ovf_cap) {
      fo.o_scores[(long long)row * fo.ovf_cap + slot] = score;
      fo.o_rows[(long long)row * fo.ovf_cap + slot] = cat_row;
    }
  }
  ++bcount;
}

}  // namespace xr
