// Per-row candidate paths (HBM / L2 bound byte streams, no tensor cores by design):
//   dense   (M,C,D) candidate tensor  — the reference's own API shape, losses.py:195 bmm is a
//           batched GEMV at 0.5 FLOP/B; kept for drop-in compatibility.
//   sampled (M,C) candidate INDICES into the item table — BASELINE config 3 (K sampled
//           negatives per positive); the gather is fused into the dot product so the
//           (M,C,D) tensor of models.py:408-410 never exists.
// Both share one kernel pair; the only difference is how candidate (i,j) is addressed.
#include "common.cuh"
#include "rowloss.cuh"
#include "rowdot.cuh"

#include <cstdint>

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

// ---- fast path for the sampled candidates of BASELINE config 3 (D = 384) -------------------------
// The table (67 MB bf16 / 134 MB fp32) mostly lives in L2; what bounds these kernels is how many
// row reads are in flight.  Block per query row, warp per group of 4 candidates: every lane
// issues its 16-byte vectors of all 4 rows before using any (8 loads in flight for bf16, 12 for
// fp32), the query sits in registers, and the 4 dot products are reduced together (6 shuffles).
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS)
sampled_logits384_kernel(const T* __restrict__ q, const T* __restrict__ table,
                         const int64_t* __restrict__ cand_idx, int64_t n_table_rows, int64_t m,
                         int64_t c, const float* __restrict__ q_inv,
                         const float* __restrict__ table_inv, float* __restrict__ logits, int64_t ld) {
  // few query rows with many candidates each (the retrieval re-score: 256 x 2,048): gridDim.y
  // blocks share a row, each taking a contiguous slice of its candidates (a multiple of 32)
  const int64_t per = ((c + gridDim.y - 1) / gridDim.y + 31) / 32 * 32;
  const int64_t c_lo = (int64_t)blockIdx.y * per;
  for (int64_t i = blockIdx.x; i < m; i += gridDim.x)
    sampled_logits384_row<T>(q + i * FD, table, cand_idx ? cand_idx + i * c : nullptr, n_table_rows, c,
                             q_inv ? q_inv[i] : 1.f, table_inv, logits + i * ld, c_lo, c_lo + per, i * c);
}

// dq_i = sum_j g[j] * tinv(j) * table[idx[j]]  (+ cosine chain rule): warps split the
// candidates (4 rows in flight each), lanes own fixed 16-byte column slices, partials are
// folded across warps through shared memory in a fixed order (deterministic).  `g` may be global
// or shared memory; s_red [ROW_THREADS/32][FD] and s_part [ROW_THREADS/32] are block scratch.
template <typename T>
__device__ __forceinline__ void sampled_dq384_row(const float* g, const T* __restrict__ q_row,
                                                  const T* __restrict__ table,
                                                  const int64_t* __restrict__ idx_row,
                                                  int64_t n_table_rows, int64_t c, float q_inv_i,
                                                  const float* __restrict__ table_inv, int cosine,
                                                  float* __restrict__ dq_row,
                                                  float (*s_red)[FD], float* s_part,
                                                  int64_t dense_base = 0) {
  constexpr int E = RowVec<T>::E, VECS = FD / E, IT = (VECS + 31) / 32;
  constexpr int NW = ROW_THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[IT][E];
#pragma unroll
  for (int t = 0; t < IT; ++t)
#pragma unroll
    for (int k = 0; k < E; ++k) acc[t][k] = 0.f;
  for (int64_t j0 = (int64_t)warp * 4; j0 < c; j0 += NW * 4) {
    float w[4];
    int64_t row[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = j0 + u;
      w[u] = j < c ? g[j] : 0.f;
      row[u] = (j < c && w[u] != 0.f) ? (idx_row ? idx_row[j] : dense_base + j) : -1;
      if (row[u] < 0 || row[u] >= n_table_rows) {   // masked candidates: skip their bytes
        w[u] = 0.f;
        row[u] = -1;
      } else if (table_inv) {
        w[u] *= table_inv[row[u]];
      }
    }
    if (w[0] == 0.f && w[1] == 0.f && w[2] == 0.f && w[3] == 0.f) continue;
    float x[4][IT][E];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < IT; ++t) {
        const int v = lane + 32 * t;
        if (v < VECS && row[u] >= 0) RowVec<T>::load(table + row[u] * FD + v * E, x[u][t]);
        else
#pragma unroll
          for (int k = 0; k < E; ++k) x[u][t][k] = 0.f;
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)   // ascending j within the warp: fixed order
#pragma unroll
      for (int t = 0; t < IT; ++t)
#pragma unroll
        for (int k = 0; k < E; ++k) acc[t][k] = fmaf(w[u], x[u][t][k], acc[t][k]);
  }
#pragma unroll
  for (int t = 0; t < IT; ++t) {
    const int v = lane + 32 * t;
    if (v < VECS)
#pragma unroll
      for (int k = 0; k < E; ++k) s_red[warp][v * E + k] = acc[t][k];
  }
  __syncthreads();
  float dotgq = 0.f;
  for (int d = threadIdx.x; d < FD; d += ROW_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < NW; ++w2) s += s_red[w2][d];   // fixed order
    s_red[0][d] = s;
    if (cosine) dotgq = fmaf(s, to_f32(q_row[d]) * q_inv_i, dotgq);
  }
  if (cosine) {
    dotgq = warp_sum(dotgq);
    if (lane == 0) s_part[warp] = dotgq;
  }
  __syncthreads();
  float s_dot = 0.f;
  if (cosine) {
#pragma unroll
    for (int w2 = 0; w2 < NW; ++w2) s_dot += s_part[w2];
  }
  for (int d = threadIdx.x; d < FD; d += ROW_THREADS) {
    float s = s_red[0][d];
    if (cosine) s = q_inv_i * (s - s_dot * to_f32(q_row[d]) * q_inv_i);
    dq_row[d] = s;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(ROW_THREADS)
sampled_dq384_kernel(const float* __restrict__ g, int64_t ld, const T* __restrict__ q,
                     const T* __restrict__ table, const int64_t* __restrict__ cand_idx,
                     int64_t n_table_rows, int64_t m, int64_t c, const float* __restrict__ q_inv,
                     const float* __restrict__ table_inv, int cosine, float* __restrict__ dq) {
  __shared__ float s_red[ROW_THREADS / 32][FD];
  __shared__ float s_part[ROW_THREADS / 32];
  for (int64_t i = blockIdx.x; i < m; i += gridDim.x)
    sampled_dq384_row<T>(g + i * ld, q + i * FD, table, cand_idx ? cand_idx + i * c : nullptr,
                         n_table_rows, c, cosine ? q_inv[i] : 1.f, table_inv, cosine, dq + i * FD, s_red,
                         s_part, i * c);
}

// ---- the sampled-candidate step in ONE pass (BASELINE config 3) --------------------------------
// logits (fused gather + dot) -> EmbedLoss pipeline -> dL/dq for one query row per block, the
// row's C logits and dL/dlogits living in shared memory: no (M, C) tensor reaches HBM and the
// step is one launch (+ the fixed-order reduction) instead of three.  The second sweep over the
// row's candidates (for dq) only touches candidates with a non-zero weight and finds them in L2.
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS, 4)
sampled_step384_kernel(const T* __restrict__ q, const T* __restrict__ table,
                       const int64_t* __restrict__ cand_idx, int64_t n_table_rows, int64_t m,
                       int64_t c, const float* __restrict__ q_inv,
                       const float* __restrict__ table_inv, xr_loss_config cfg, int grad_kind,
                       float grad_scale, float* __restrict__ dq, double* __restrict__ row_out) {
  static_assert(ROW_THREADS == RL_THREADS, "the row helpers assume one block shape");
  extern __shared__ float s_dyn[];   // logits[c] | dlogits[c] | hard-mining mask bytes[c]
  float* s_logit = s_dyn;
  float* s_g = s_dyn + c;
  uint8_t* s_mask = reinterpret_cast<uint8_t*>(s_g + c);
  __shared__ float s_red[ROW_THREADS / 32][FD];
  __shared__ float s_part[ROW_THREADS / 32];
  __shared__ double s_d[RL_WARPS];
  __shared__ float s_f[RL_WARPS];
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_sel[4];
  __shared__ int s_warp_cnt[RL_WARPS];
  const int cosine = table_inv != nullptr;
  for (int64_t i = blockIdx.x; i < m; i += gridDim.x) {
    const float qi = q_inv ? q_inv[i] : 1.f;
    sampled_logits384_row<T>(q + i * FD, table, cand_idx + i * c, n_table_rows, c, qi, table_inv, s_logit);
    __syncthreads();
    rowloss_row(s_logit, c, /*target column*/ 0, cfg, grad_kind, grad_scale, dq ? s_g : nullptr,
                row_out + i * ROW_SLOTS, s_mask, RowLossScratch{s_d, s_f, s_hist, s_sel, s_warp_cnt});
    __syncthreads();
    if (dq)
      sampled_dq384_row<T>(s_g, q + i * FD, table, cand_idx + i * c, n_table_rows, c, qi, table_inv,
                           cosine, dq + i * FD, s_red, s_part);
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
  if (dim == FD && dtype == XR_BF16 && !q_inv_norm && !cand_inv_norm_out) {
    // dot logits of a bf16 (M, C, 384) tensor: the 16-byte-load / 4-rows-in-flight kernel of the
    // sampled path with implicit rows i*C + j (the generic kernel's 8-byte bf16 loads reach 0.5 of HBM)
    sampled_logits384_kernel<__nv_bfloat16><<<dim3((unsigned)row_grid(m), 1), ROW_THREADS, 0, s>>>(
        (const __nv_bfloat16*)q, (const __nv_bfloat16*)cand, nullptr, m * c, m, c, nullptr, nullptr,
        logits, ld);
    XR_LAUNCH_CHECK("dense_logits384");
    return XR_OK;
  }
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
  if (dim == FD) {   // BASELINE config 3 shape: rows in flight instead of one row per warp
    // fill the machine when there are fewer rows than ~8 blocks per SM: split the candidates
    int gy = 1;
    const int64_t want = (int64_t)sm_count() * 8;
    if (m < want && c >= 256) {
      gy = (int)((want + m - 1) / m);
      const int max_gy = (int)(c / 128);   // at least 128 candidates per block
      if (gy > max_gy) gy = max_gy;
      if (gy < 1) gy = 1;
    }
    const dim3 grid((unsigned)row_grid(m), (unsigned)gy);
    if (dtype == XR_F32)
      sampled_logits384_kernel<float><<<grid, ROW_THREADS, 0, s>>>(
          (const float*)q, (const float*)table, cand_idx, n_rows, m, c, q_inv_norm, table_inv_norm,
          logits, ld);
    else
      sampled_logits384_kernel<__nv_bfloat16><<<grid, ROW_THREADS, 0, s>>>(
          (const __nv_bfloat16*)q, (const __nv_bfloat16*)table, cand_idx, n_rows, m, c, q_inv_norm,
          table_inv_norm, logits, ld);
    XR_LAUNCH_CHECK("sampled_logits384");
    return XR_OK;
  }
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
  if (dim == FD && !cosine) {   // 16-byte loads, 4 candidate rows in flight per warp, zero weights skipped
    if (dtype == XR_BF16)
      sampled_dq384_kernel<__nv_bfloat16><<<row_grid(m), ROW_THREADS, 0, s>>>(
          dlogits, ld, (const __nv_bfloat16*)q, (const __nv_bfloat16*)cand, nullptr, m * c, m, c, nullptr,
          nullptr, 0, dq);
    else
      sampled_dq384_kernel<float><<<row_grid(m), ROW_THREADS, 0, s>>>(
          dlogits, ld, (const float*)q, (const float*)cand, nullptr, m * c, m, c, nullptr, nullptr, 0, dq);
    XR_LAUNCH_CHECK("dense_dq384");
    return XR_OK;
  }
  return dtype == XR_F32
             ? launch_row_dq<float>(dlogits, ld, q, cand, nullptr, 0, m, c, dim, q_inv_norm,
                                    nullptr, cosine ? cand_inv_norm : nullptr, cosine, dq, s)
             : launch_row_dq<__nv_bfloat16>(dlogits, ld, q, cand, nullptr, 0, m, c, dim,
                                            q_inv_norm, nullptr,
                                            cosine ? cand_inv_norm : nullptr, cosine, dq, s);
}

// dL/d candidate_embed of a DENSE (M, C, D) candidate tensor (the reference's own API form; its trainer never
// asks for it: the table is frozen, models.py:251-253).  Dot logits (losses.py:195): dcand[i][c] = w q_i.
// Cosine logits (losses.py:206-208): dcand[i][c] = w (q^_i - cos c^_ic) / max(|c_ic|, eps), cos = the logit.
// One warp per candidate row, 16-byte stores when D % 4 == 0; HBM-bound (writes M C D fp32).
template <typename T>
__global__ void __launch_bounds__(256)
dcand_dense_kernel(const float* __restrict__ dlogits, const float* __restrict__ logits, int64_t ld,
                   const T* __restrict__ q, const T* __restrict__ cand, int64_t rows, int64_t c, int dim, int cosine,
                   const float* __restrict__ q_inv, const float* __restrict__ cand_inv, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const int64_t i = r / c, j = r % c;
    const float w = dlogits[i * ld + j];
    const T* qr = q + i * dim;
    const T* cr = cand + r * dim;
    float* o = out + r * dim;
    if (!cosine) {
      for (int d = lane; d < dim; d += 32) o[d] = w * to_f32(qr[d]);
    } else {
      const float cosv = logits[i * ld + j], qi = q_inv[i], ci = cand_inv[r];
      for (int d = lane; d < dim; d += 32) o[d] = w * (to_f32(qr[d]) * qi - cosv * to_f32(cr[d]) * ci) * ci;
    }
  }
}

extern "C" int xr_dcand_dense(const float* dlogits, const float* logits, int64_t ld, const void* q, const void* cand,
                              int64_t m, int64_t c, int64_t dim, int dtype, int cosine, const float* q_inv_norm,
                              const float* cand_inv_norm, float* dcand, void* stream) {
  XR_CHECK_ARG(dlogits && q && cand && dcand && ld >= c, "xr_dcand_dense: bad arguments");
  XR_CHECK_ARG(!cosine || (logits && q_inv_norm && cand_inv_norm), "xr_dcand_dense: cosine needs the logits and norms");
  int rc = check_row_args("xr_dcand_dense", m, c, dim, dtype, q, cand);
  if (rc) return rc;
  if (m == 0 || c == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  const int64_t rows = m * c;
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (dtype == XR_F32)
    dcand_dense_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(dlogits, logits, ld, (const float*)q, (const float*)cand,
                                                               rows, c, (int)dim, cosine, q_inv_norm, cand_inv_norm, dcand);
  else
    dcand_dense_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>(
        dlogits, logits, ld, (const __nv_bfloat16*)q, (const __nv_bfloat16*)cand, rows, c, (int)dim, cosine,
        q_inv_norm, cand_inv_norm, dcand);
  XR_LAUNCH_CHECK("dcand_dense");
  return XR_OK;
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
  if (dim == FD) {
    if (dtype == XR_F32)
      sampled_dq384_kernel<float><<<row_grid(m), ROW_THREADS, 0, s>>>(
          dlogits, ld, (const float*)q, (const float*)table, cand_idx, n_rows, m, c, q_inv_norm,
          table_inv_norm, cosine, dq);
    else
      sampled_dq384_kernel<__nv_bfloat16><<<row_grid(m), ROW_THREADS, 0, s>>>(
          dlogits, ld, (const __nv_bfloat16*)q, (const __nv_bfloat16*)table, cand_idx, n_rows, m, c,
          q_inv_norm, table_inv_norm, cosine, dq);
    XR_LAUNCH_CHECK("sampled_dq384");
    return XR_OK;
  }
  return dtype == XR_F32
             ? launch_row_dq<float>(dlogits, ld, q, table, cand_idx, n_rows, m, c, dim,
                                    q_inv_norm, table_inv_norm, nullptr, cosine, dq, s)
             : launch_row_dq<__nv_bfloat16>(dlogits, ld, q, table, cand_idx, n_rows, m, c, dim,
                                            q_inv_norm, table_inv_norm, nullptr, cosine, dq, s);
}

// ---- BASELINE config 3 in one pass: logits + EmbedLoss pipeline + dL/dq per query row -----------
extern "C" size_t xr_sampled_step_workspace_bytes(int64_t m) {
  return ((size_t)(m > 0 ? m : 1) * ROW_SLOTS * sizeof(double) + 255) / 256 * 256 + kRowlossPartialBytes;
}

extern "C" int xr_sampled_step(const void* q, const void* table, int64_t n_rows,
                               const int64_t* cand_idx, int64_t m, int64_t c, int64_t dim, int dtype,
                               const float* table_inv_norm, const float* q_inv_norm,
                               const xr_loss_config* cfg, int grad_kind, float grad_scale, float* dq,
                               double* losses_out, double* stats_out, void* workspace,
                               void* stream) {
  XR_CHECK_ARG(q && table && cand_idx && cfg && workspace && n_rows > 0, "xr_sampled_step: bad arguments");
  XR_CHECK_ARG(dim == FD, "xr_sampled_step: this build is specialised for dim = %d", FD);
  XR_CHECK_ARG(c >= 1 && c <= 8192, "xr_sampled_step: 1 <= candidates per row <= 8192");
  XR_CHECK_ARG(grad_kind < XR_NUM_LOSSES && (grad_kind < 0 || dq), "xr_sampled_step: bad grad_kind / dq");
  XR_CHECK_ARG((table_inv_norm == nullptr) == (q_inv_norm == nullptr),
               "xr_sampled_step: cosine needs both inverse norms");
  int rc = check_row_args("xr_sampled_step", m, c, dim, dtype, q, table);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  double* row_out = (double*)workspace;
  if (m > 0) {
    const size_t smem = (size_t)c * 9;   // two float arrays + the mask bytes
    float* dq_arg = grad_kind >= 0 ? dq : nullptr;
    if (dtype == XR_F32) {
      static size_t conf = 0;
      if (smem > conf) {
        XR_CUDA(cudaFuncSetAttribute(sampled_step384_kernel<float>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 9));
        conf = 8192 * 9;
      }
      sampled_step384_kernel<float><<<row_grid(m), ROW_THREADS, smem, s>>>(
          (const float*)q, (const float*)table, cand_idx, n_rows, m, c, q_inv_norm, table_inv_norm,
          *cfg, grad_kind, grad_scale, dq_arg, row_out);
    } else {
      static size_t conf = 0;
      if (smem > conf) {
        XR_CUDA(cudaFuncSetAttribute(sampled_step384_kernel<__nv_bfloat16>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 9));
        conf = 8192 * 9;
      }
      sampled_step384_kernel<__nv_bfloat16><<<row_grid(m), ROW_THREADS, smem, s>>>(
          (const __nv_bfloat16*)q, (const __nv_bfloat16*)table, cand_idx, n_rows, m, c, q_inv_norm,
          table_inv_norm, *cfg, grad_kind, grad_scale, dq_arg, row_out);
    }
    XR_LAUNCH_CHECK("sampled_step384");
  }
  if (losses_out || stats_out)
    return launch_rowloss_reduce(row_out, m, c, cfg->num_hard_negatives, losses_out, stats_out, s, nullptr,
                                 (double*)((uint8_t*)workspace +
                                           ((size_t)(m > 0 ? m : 1) * ROW_SLOTS * sizeof(double) + 255) / 256 * 256));
  return XR_OK;
}
