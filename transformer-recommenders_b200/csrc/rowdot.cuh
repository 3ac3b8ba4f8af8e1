// Gather-dot of one query row against table rows (D = 384): the arithmetic of xr_logits_sampled, shared
// with the retrieval finalize kernel so that both produce the same bits for the same (query, row).
#pragma once

#include "common.cuh"

namespace xr {

constexpr int ROW_THREADS = 256;
constexpr int FD = 384;
template <typename T>
struct RowVec;   // 16-byte vector of a row -> floats
template <>
struct RowVec<__nv_bfloat16> {
  static constexpr int E = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(p));
    const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[2 * k] = __uint_as_float(w[k] << 16);
      o[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
    }
  }
};
template <>
struct RowVec<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(p));
    o[0] = __int_as_float(v.x); o[1] = __int_as_float(v.y);
    o[2] = __int_as_float(v.z); o[3] = __int_as_float(v.w);
  }
};

// reduce 4 per-lane partials over the warp: lanes 0 / 8 / 16 / 24 end up with totals a / b / c / d
__device__ __forceinline__ float warp_sum4(float a, float b, float c, float d, int lane) {
  const bool hi16 = lane & 16, hi8 = lane & 8;
  float p = (hi16 ? c : a) + __shfl_xor_sync(0xffffffffu, hi16 ? a : c, 16);
  float q = (hi16 ? d : b) + __shfl_xor_sync(0xffffffffu, hi16 ? b : d, 16);
  float r = (hi8 ? q : p) + __shfl_xor_sync(0xffffffffu, hi8 ? p : q, 8);
  r += __shfl_xor_sync(0xffffffffu, r, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// one query row: out[j] = q . table[idx[j]] (* qi * table_inv) for j < c.  `out` may be global or
// shared memory.  All ROW_THREADS threads call it together.
template <typename T>
__device__ __forceinline__ void sampled_logits384_row(const T* __restrict__ q_row,
                                                      const T* __restrict__ table,
                                                      const int64_t* __restrict__ idx_row,
                                                      int64_t n_table_rows, int64_t c, float qi,
                                                      const float* __restrict__ table_inv,
                                                      float* out, int64_t c_lo = 0,
                                                      int64_t c_hi = INT64_MAX,
                                                      int64_t dense_base = 0) {
  constexpr int E = RowVec<T>::E, VECS = FD / E, IT = (VECS + 31) / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = ROW_THREADS / 32;
  float qr[IT][E];
#pragma unroll
  for (int t = 0; t < IT; ++t) {
    const int v = lane + 32 * t;
    if (v < VECS) RowVec<T>::load(q_row + v * E, qr[t]);
    else
#pragma unroll
      for (int k = 0; k < E; ++k) qr[t][k] = 0.f;
  }
  if (c_hi < c) c = c_hi;   // this block's slice of the candidates [c_lo, c_hi)
  for (int64_t j0 = c_lo + (int64_t)warp * 4; j0 < c; j0 += nwarp * 4) {
    int64_t row[4];
    bool valid[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = j0 + u;
      row[u] = j < c ? (idx_row ? idx_row[j] : dense_base + j) : 0;   // dense (M,C,D): row i*C + j
      valid[u] = row[u] >= 0 && row[u] < n_table_rows;
      if (!valid[u]) row[u] = 0;
    }
    float x[4][IT][E];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < IT; ++t) {
        const int v = lane + 32 * t;
        if (v < VECS && valid[u]) RowVec<T>::load(table + row[u] * FD + v * E, x[u][t]);   // invalid: no bytes read
        else
#pragma unroll
          for (int k = 0; k < E; ++k) x[u][t][k] = 0.f;
      }
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < IT; ++t)
#pragma unroll
        for (int k = 0; k < E; ++k) dot[u] = fmaf(qr[t][k], x[u][t][k], dot[u]);
    const float tot = warp_sum4(dot[0], dot[1], dot[2], dot[3], lane);
    if ((lane & 7) == 0) {
      const int u = lane >> 3;
      const int64_t j = j0 + u;
      int64_t r_sel = row[0];
      bool v_sel = valid[0];
#pragma unroll
      for (int uu = 1; uu < 4; ++uu)   // static indexing keeps row[] / valid[] in registers
        if (u == uu) {
          r_sel = row[uu];
          v_sel = valid[uu];
        }
      if (j < c) {
        float scale = qi;
        if (table_inv) scale *= table_inv[r_sel];
        out[j] = v_sel ? tot * scale : CUDART_NAN_F;
      }
    }
  }
}


}  // namespace xr
