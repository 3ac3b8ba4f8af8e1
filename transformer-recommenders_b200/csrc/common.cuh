// Shared helpers for the xfmr_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/xfmr_b200.h"

namespace xr {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define XR_CHECK_ARG(cond, ...)  \
  do {                           \
    if (!(cond)) {               \
      xr::set_error(__VA_ARGS__); \
      return XR_E_INVALID;       \
    }                            \
  } while (0)

#define XR_CUDA(expr)                                 \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return xr::cuda_fail(_e, #expr); \
  } while (0)

#define XR_LAUNCH_CHECK(name)                                   \
  do {                                                          \
    cudaError_t _e = cudaGetLastError();                        \
    if (_e != cudaSuccess) return xr::cuda_fail(_e, "launch " name); \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();       // SMs the persistent kernels may fill (physical count minus xr_reserve_sms)
int sm_count_max();   // physical SM count: workspace sizing

constexpr int kWarp = 32;

// per-row slots (doubles) the loss passes leave for rowloss_reduce_kernel: the six per-row loss
// values (losses.py:352-372, 420-543) and the LogitsStatistics ingredients (losses.py:383-405)
constexpr int ROW_SLOTS = 20;
enum RowSlot {
  S_ALIGN = 0, S_CONTR, S_INFONCE, S_NCE, S_HINGE, S_LOGISTIC,
  S_DENS, S_POS, S_NCOUNT, S_NSUM, S_NSQ, S_NMIN, S_NMAX, S_USED
};
constexpr size_t kRowlossPartialBytes = 16384;   // scratch of the two-launch row-slot reduction
int launch_rowloss_reduce(const double* row_out, int64_t m, int64_t c, int n_hard, double* losses_out,
                          double* stats_out, cudaStream_t s, const int* dyn_m_cn, double* partial);
int launch_rowloss_reduce2(const double* row_out, const double* row_out2, int64_t m, int64_t c, double* losses_out,
                           double* stats_out, double* losses_out2, cudaStream_t s, const int* dyn_m_cn,
                           double* partial);

// ---- small device helpers --------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit accesses: data touched once should not displace the L1
__device__ __forceinline__ int4 ld_stream16(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

// load element j of a row that is fp32 or bf16
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}

// order-preserving map float -> uint32 (ascending); -0.0 canonicalised to +0.0, NaN largest
__device__ __forceinline__ uint32_t float_key(float f) {
  f += 0.0f;
  uint32_t u = __float_as_uint(f);
  if (f != f) return 0xFFFFFFFFu;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(u);
}

}  // namespace xr
