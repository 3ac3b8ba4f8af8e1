// Per-row candidate paths (HBM / L2 bound byte streams, no tensor cores by design):
//   dense   (M,C,D) candidate tensor  — the reference's own API shape, losses.py:195 bmm is a
//           batched GEMV at 0.5 FLOP/B; kept for drop-in compatibility.
//   sampled (M,C) candidate INDICES into the item table — BASELINE config 3 (K sampled
//           negatives per positive); the gather is fused into the dot product so the
//           (M,C,D) tensor of models.py:408-410 never exists.
// Both share one kernel pair; the only difference is how candidate (i,j) is addressed.
#include "common.cuh"

namespace xr {

template <typename T>
struct Vec8;  // 8 elements per lane-iteration (16 B for bf16, 32 B for fp32)

template <typename T>
__device__ __forceinline__ void load_vec4(const T* p, float (&o)[4]);
template <>
__device__ __forceinline__ void load_vec4<float>(const float* p, float (&o)[4]) {
  const int4 v = ld_stream16(p);
  o[0] = __int_as_float(v.x); o[1] = __int_as_float(v.y);
  o[2] = __int_as_float(v.z); o[3] = __int_as_float(v.w);
}
template <>
__device__ __forceinline__ void load_vec4<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[4]) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
  o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}

constexpr int ROW_THREADS = 256;

// logits[i,j] = q_i . cand(i,j) [* q_inv[i] * c_inv(i,j)];  block per query row, warp per
// candidate, lanes stride the embedding dimension in 4-element vectors.
// cand_inv_out (nullable): 1/max(||cand(i,j)||, eps) computed in the same pass (dense cosine).
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS)
row_logits_kernel(const T* __restrict__ q, const T* __restrict__ cand_base,
                  const int64_t* __restrict__ cand_idx, int64_t n_table_rows, int64_t m,
                  int64_t c, int dim, const float* __restrict__ q_inv,
                  const float* __restrict__ table_inv, float* __restrict__ cand_inv_out, float eps,
                  float* __restrict__ logits, int64_t ld) {
  extern __shared__ float s_q[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = ROW_THREADS / 32;
  for (int64_t i = blockIdx.x; i < m; i += gridDim.x) {
    __syncthreads();
    for (int d = threadIdx.x; d < dim; d += ROW_THREADS) s_q[d] = to_f32(q[i * dim + d]);
    __syncthreads();
    const float qi = q_inv ? q_inv[i] : 1.f;
    for (int64_t j = warp; j < c; j += nwarp) {
      int64_t row = cand_idx ? cand_idx[i * c + j] : (i * c + j);
      const bool valid = !cand_idx || (row >= 0 && row < n_table_rows);
      if (!valid) row = 0;
      const T* p = cand_base + row * dim;
      float dot = 0.f, ss = 0.f;
      for (int d = lane * 4; d < dim; d += 128) {
        float v[4];
        load_vec4<T>(p + d, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          dot = fmaf(s_q[d + k], v[k], dot);
          ss = fmaf(v[k], v[k], ss);
        }
      }
      dot = warp_sum(dot);
      float scale = qi;
      if (cand_inv_out) {
        ss = warp_sum(ss);
        const float ci = 1.0f / fmaxf(sqrtf(ss), eps);
        if (lane == 0) cand_inv_out[i * c + j] = ci;
        scale *= ci;
      } else if (table_inv) {
        scale *= table_inv[row];
      }
      if (lane == 0) logits[i * ld + j] = valid ? dot * scale : CUDART_NAN_F;
    }
  }
}

// dq_i = sum_j g[i,j] * cinv(i,j) * cand(i,j)   (+ cosine chain rule with qhat, q_inv)
// block per query row; each thread owns 4 consecutive embedding elements of up to 2 strips.
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS)
row_dq_kernel(const float* __restrict__ g, int64_t ld, const T* __restrict__ qhat /* RAW q; qhat = q * q_inv */,
              const T* __restrict__ cand_base, const int64_t* __restrict__ cand_idx,
              int64_t n_table_rows, int64_t m, int64_t c, int dim,
              const float* __restrict__ q_inv, const float* __restrict__ table_inv,
              const float* __restrict__ cand_inv, int cosine, float* __restrict__ dq) {
  // threads are arranged as (group, vec): vec covers the embedding dim in 4-element vectors,
  // groups split the candidate loop; partial sums are reduced through shared memory.
  extern __shared__ float s_red[];  // [groups][dim]
  const int nvec = dim / 4;
  const int groups = ROW_THREADS / nvec > 0 ? ROW_THREADS / nvec : 1;
  const int vec = threadIdx.x % nvec, grp = threadIdx.x / nvec;
  const bool active = grp < groups && threadIdx.x < groups * nvec;
  __shared__ float s_part[ROW_THREADS / 32];
  for (int64_t i = blockIdx.x; i < m; i += gridDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
      for (int64_t j = grp; j < c; j += groups) {
        float w = g[i * ld + j];
        if (w == 0.f) continue;  // masked candidates contribute nothing: skip their bytes
        int64_t row = cand_idx ? cand_idx[i * c + j] : (i * c + j);
        if (cand_idx && (row < 0 || row >= n_table_rows)) continue;
        if (cand_inv) w *= cand_inv[i * c + j];
        else if (table_inv) w *= table_inv[row];
        float v[4];
        load_vec4<T>(cand_base + row * dim + vec * 4, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fmaf(w, v[k], acc[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) s_red[grp * dim + vec * 4 + k] = acc[k];
    }
    __syncthreads();
    // fixed-order reduction over groups (deterministic)
    float dotgq = 0.f;
    for (int d = threadIdx.x; d < dim; d += ROW_THREADS) {
      float s = 0.f;
      for (int gq = 0; gq < groups; ++gq) s += s_red[gq * dim + d];
      s_red[d] = s;  // group 0's strip now holds the total
      if (cosine) dotgq = fmaf(s, to_f32(qhat[i * dim + d]) * q_inv[i], dotgq);
    }
    if (cosine) {
      dotgq = warp_sum(dotgq);
      if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = dotgq;
    }
    __syncthreads();
    float s_dot = 0.f;
    if (cosine) {
#pragma unroll
      for (int w = 0; w < ROW_THREADS / 32; ++w) s_dot += s_part[w];  // fixed order
    }
    for (int d = threadIdx.x; d < dim; d += ROW_THREADS) {
      float s = s_red[d];
      if (cosine) s = q_inv[i] * (s - s_dot * to_f32(qhat[i * dim + d]) * q_inv[i]);
      dq[i * dim + d] = s;
    }
    __syncthreads();
  }
}

static inline int row_grid(int64_t m) {
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(m < cap ? (m < 1 ? 1 : m) : cap);
}

template <typename T>
static int launch_row_logits(const void* q, const void* base, const int64_t* idx, int64_t nrows,
                             int64_t m, int64_t c, int64_t dim, const float* q_inv,
                             const float* table_inv, float* cand_inv_out, float eps, float* logits,
                             int64_t ld, cudaStream_t s) {
  row_logits_kernel<T><<<row_grid(m), ROW_THREADS, dim * sizeof(float), s>>>(
      (const T*)q, (const T*)base, idx, nrows, m, c, (int)dim, q_inv, table_inv, cand_inv_out, eps,
      logits, ld);
  XR_LAUNCH_CHECK("row_logits");
  return XR_OK;
}

template <typename T>
static int launch_row_dq(const float* g, int64_t ld, const void* qhat, const void* base,
                         const int64_t* idx, int64_t nrows, int64_t m, int64_t c, int64_t dim,
                         const float* q_inv, const float* table_inv, const float* cand_inv,
                         int cosine, float* dq, cudaStream_t s) {
  const int nvec = (int)(dim / 4);
  const int groups = ROW_THREADS / nvec > 0 ? ROW_THREADS / nvec : 1;
  row_dq_kernel<T><<<row_grid(m), ROW_THREADS, (size_t)groups * dim * sizeof(float), s>>>(
      g, ld, (const T*)qhat, (const T*)base, idx, nrows, m, c, (int)dim, q_inv, table_inv,
      cand_inv, cosine, dq);
  XR_LAUNCH_CHECK("row_dq");
  return XR_OK;
}

static int check_row_args(const char* who, int64_t m, int64_t c, int64_t dim, int dtype,
                          const void* a, const void* b) {
  XR_CHECK_ARG(m >= 0 && c >= 0 && dim > 0, "%s: bad sizes", who);
  XR_CHECK_ARG(dim % 4 == 0 && dim <= 1024, "%s: dim must be a multiple of 4 and <= 1024", who);
  XR_CHECK_ARG(dtype == XR_F32 || dtype == XR_BF16, "%s: bad dtype", who);
  XR_CHECK_ARG((uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0, "%s: buffers must be 16B aligned",
               who);
  return XR_OK;
}

}  // namespace xr

using namespace xr;

extern "C" int xr_logits_dense(const void* q, const void* cand, int64_t m, int64_t c, int64_t dim,
                               int dtype, const float* q_inv_norm, float* cand_inv_norm_out,
                               float eps, float* logits, int64_t ld, void* stream) {
  XR_CHECK_ARG(q && cand && logits && ld >= c, "xr_logits_dense: bad arguments");
  int rc = check_row_args("xr_logits_dense", m, c, dim, dtype, q, cand);
  if (rc) return rc;
  if (m == 0 || c == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  return dtype == XR_F32
             ? launch_row_logits<float>(q, cand, nullptr, 0, m, c, dim, q_inv_norm, nullptr,
                                        cand_inv_norm_out, eps, logits, ld, s)
             : launch_row_logits<__nv_bfloat16>(q, cand, nullptr, 0, m, c, dim, q_inv_norm,
                                                nullptr, cand_inv_norm_out, eps, logits, ld, s);
}

extern "C" int xr_logits_sampled(const void* q, const void* table, int64_t n_rows,
                                 const int64_t* cand_idx, int64_t m, int64_t c, int64_t dim,
                                 int dtype, const float* table_inv_norm, const float* q_inv_norm,
                                 float* logits, int64_t ld, void* stream) {
  XR_CHECK_ARG(q && table && cand_idx && logits && ld >= c && n_rows > 0,
               "xr_logits_sampled: bad arguments");
  int rc = check_row_args("xr_logits_sampled", m, c, dim, dtype, q, table);
  if (rc) return rc;
  if (m == 0 || c == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  return dtype == XR_F32
             ? launch_row_logits<float>(q, table, cand_idx, n_rows, m, c, dim, q_inv_norm,
                                        table_inv_norm, nullptr, 0.f, logits, ld, s)
             : launch_row_logits<__nv_bfloat16>(q, table, cand_idx, n_rows, m, c, dim, q_inv_norm,
                                                table_inv_norm, nullptr, 0.f, logits, ld, s);
}

extern "C" int xr_dq_dense(const float* dlogits, int64_t ld, const void* q, const void* cand,
                           int64_t m, int64_t c, int64_t dim, int dtype, int cosine,
                           const float* q_inv_norm, const float* cand_inv_norm, float* dq,
                           void* stream) {
  XR_CHECK_ARG(dlogits && q && cand && dq && ld >= c, "xr_dq_dense: bad arguments");
  XR_CHECK_ARG(!cosine || (q_inv_norm && cand_inv_norm), "xr_dq_dense: cosine needs the norms");
  int rc = check_row_args("xr_dq_dense", m, c, dim, dtype, q, cand);
  if (rc) return rc;
  if (m == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  return dtype == XR_F32
             ? launch_row_dq<float>(dlogits, ld, q, cand, nullptr, 0, m, c, dim, q_inv_norm,
                                    nullptr, cosine ? cand_inv_norm : nullptr, cosine, dq, s)
             : launch_row_dq<__nv_bfloat16>(dlogits, ld, q, cand, nullptr, 0, m, c, dim,
                                            q_inv_norm, nullptr,
                                            cosine ? cand_inv_norm : nullptr, cosine, dq, s);
}

extern "C" int xr_dq_sampled(const float* dlogits, int64_t ld, const void* q, const void* table,
                             int64_t n_rows, const int64_t* cand_idx, int64_t m, int64_t c,
                             int64_t dim, int dtype, const float* table_inv_norm,
                             const float* q_inv_norm, float* dq, void* stream) {
  XR_CHECK_ARG(dlogits && q && table && cand_idx && dq && ld >= c, "xr_dq_sampled: bad arguments");
  const int cosine = table_inv_norm != nullptr;
  XR_CHECK_ARG(!cosine || q_inv_norm, "xr_dq_sampled: cosine needs q_inv_norm");
  int rc = check_row_args("xr_dq_sampled", m, c, dim, dtype, q, table);
  if (rc) return rc;
  if (m == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  return dtype == XR_F32
             ? launch_row_dq<float>(dlogits, ld, q, table, cand_idx, n_rows, m, c, dim,
                                    q_inv_norm, table_inv_norm, nullptr, cosine, dq, s)
             : launch_row_dq<__nv_bfloat16>(dlogits, ld, q, table, cand_idx, n_rows, m, c, dim,
                                            q_inv_norm, table_inv_norm, nullptr, cosine, dq, s);
}
