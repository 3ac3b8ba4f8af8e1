// L2 normalisation of one 384-wide bf16 row by one warp (losses.py:206-208 / index.py:47): 16-byte loads,
// ONE summation order shared by xr_normalize_rows (bf16 -> bf16, D = 384) and the step's three-operand
// launch, so the module path and the sync-free step produce the same bits.
#pragma once

#include "common.cuh"

namespace xr {

// returns 1 / max(||x||, eps); y (nullable) receives the normalised row
__device__ __forceinline__ float normalize_row384_bf16(const __nv_bfloat16* __restrict__ x,
                                                       __nv_bfloat16* __restrict__ y, float eps, int lane) {
  constexpr int VPR = 384 * 2 / 16;   // 48 sixteen-byte vectors per row: lanes 0-31, then lanes 0-15
  const int4* src = reinterpret_cast<const int4*>(x);
  int4 v[2];
  v[0] = __ldg(src + lane);
  v[1] = lane < VPR - 32 ? __ldg(src + 32 + lane) : make_int4(0, 0, 0, 0);
  float f[2][8];
  float ss = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const uint32_t w[4] = {(uint32_t)v[t].x, (uint32_t)v[t].y, (uint32_t)v[t].z, (uint32_t)v[t].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[t][2 * k] = __uint_as_float(w[k] << 16);
      f[t][2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
    }
  }
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) ss = fmaf(f[t][k], f[t][k], ss);
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
  if (y) {
    int4* dst = reinterpret_cast<int4*>(y);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (t == 1 && lane >= VPR - 32) break;
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(f[t][2 * k] * inv, f[t][2 * k + 1] * inv);
        o[k] = *reinterpret_cast<const uint32_t*>(&h);
      }
      dst[t * 32 + lane] = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
    }
  }
  return inv;
}

}  // namespace xr
