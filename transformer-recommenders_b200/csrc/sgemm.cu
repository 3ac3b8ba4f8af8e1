// fp32-exact contraction path (CUDA cores, fp32 FFMA, fp32 accumulate).
//
// Used where north_star asks for 1e-5 relative agreement with the reference's fp32 arithmetic
// (BASELINE config 1) and as the in-GPU cross-check of the tcgen05 kernels:
//   xr_logits_pool : [rowdot(q,pos) | Q.Neg^T]   == models.py:408-410 + losses.py:195
//   xr_dq_pool     : dQ = G[:,neg] . Neg + g_pos * pos   (autograd of the above)
//   xr_scores      : Q . Catalog^T with optional cosine scaling (index.py:47, 244-254)
// One tiled kernel: C[M,N] = A[M,K] . op(B), 128x128x16 tiles, 8x8 register micro-tiles.
#include "common.cuh"

namespace xr {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, GEMM_THREADS = 256;
constexpr int PAD = 4;

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&o)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&o)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[4]) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
  o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}

// B_KN = false: B is [N,K] row-major (C = A.B^T);  true: B is [K,N] row-major (C = A.B)
// VEC: all operand rows are 4-element aligned (pointer, leading dim); else guarded scalar loads
template <typename TA, typename TB, bool B_KN, bool VEC>
__global__ void __launch_bounds__(GEMM_THREADS)
sgemm_kernel(const TA* __restrict__ A, int64_t lda, const TB* __restrict__ B, int64_t ldb,
             float* __restrict__ C, int64_t ldc, int M, int N, int K,
             const float* __restrict__ row_scale, const float* __restrict__ col_scale) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // A tile: 128 rows x 16 k  -> thread loads rows (tid>>2), (tid>>2)+64 at k = (tid&3)*4
  const int a_r = tid >> 2, a_k = (tid & 3) * 4;
  // B tile NT: same shape as A.  B tile NN: 16 k x 128 n -> k = tid>>5 (+8), n = (tid&31)*4
  const int b_k = tid >> 5, b_n = (tid & 31) * 4;

  float ra[2][4], rb[2][4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = m0 + a_r + h * 64, k = k0 + a_k;
#pragma unroll
      for (int j = 0; j < 4; ++j) ra[h][j] = 0.f;
      if (r < M) {
        const TA* p = A + (int64_t)r * lda + k;
        if (VEC && k + 3 < K) {
          load4<TA>(p, ra[h]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (k + j < K) ra[h][j] = to_f32(p[j]);
        }
      }
    }
    if (!B_KN) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = n0 + a_r + h * 64, k = k0 + a_k;
#pragma unroll
        for (int j = 0; j < 4; ++j) rb[h][j] = 0.f;
        if (r < N) {
          const TB* p = B + (int64_t)r * ldb + k;
          if (VEC && k + 3 < K) {
            load4<TB>(p, rb[h]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (k + j < K) rb[h][j] = to_f32(p[j]);
          }
        }
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = k0 + b_k + h * 8, n = n0 + b_n;
#pragma unroll
        for (int j = 0; j < 4; ++j) rb[h][j] = 0.f;
        if (k < K) {
          const TB* p = B + (int64_t)k * ldb + n;
          if (VEC && n + 3 < N) {
            load4<TB>(p, rb[h]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n + j < N) rb[h][j] = to_f32(p[j]);
          }
        }
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < 4; ++j) As[a_k + j][a_r + h * 64] = ra[h][j];
    if (!B_KN) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[a_k + j][a_r + h * 64] = rb[h][j];
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        *reinterpret_cast<float4*>(&Bs[b_k + h * 8][b_n]) =
            make_float4(rb[h][0], rb[h][1], rb[h][2], rb[h][3]);
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);  // next tile's global loads overlap this tile's FFMAs
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
    const float rs = row_scale ? row_scale[r] : 1.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (c >= N) continue;
      float v = acc[i][j];
      if (row_scale) v *= rs;
      if (col_scale) v *= col_scale[c];
      C[(int64_t)r * ldc + c] = v;
    }
  }
}

template <typename TA, typename TB, bool B_KN>
static int launch_sgemm(const TA* A, int64_t lda, const TB* B, int64_t ldb, float* C, int64_t ldc,
                        int64_t M, int64_t N, int64_t K, const float* row_scale,
                        const float* col_scale, cudaStream_t s) {
  if (M == 0 || N == 0) return XR_OK;
  XR_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "sgemm: size overflow");
  const int ea = sizeof(TA) == 4 ? 16 : 8, eb = sizeof(TB) == 4 ? 16 : 8;
  const bool vec = ((uintptr_t)A % ea == 0) && ((uintptr_t)B % eb == 0) && (lda % 4 == 0) &&
                   (ldb % 4 == 0);
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));
  XR_CHECK_ARG(grid.y <= 65535, "sgemm: too many row tiles (split the call)");
  if (vec)
    sgemm_kernel<TA, TB, B_KN, true><<<grid, GEMM_THREADS, 0, s>>>(
        A, lda, B, ldb, C, ldc, (int)M, (int)N, (int)K, row_scale, col_scale);
  else
    sgemm_kernel<TA, TB, B_KN, false><<<grid, GEMM_THREADS, 0, s>>>(
        A, lda, B, ldb, C, ldc, (int)M, (int)N, (int)K, row_scale, col_scale);
  XR_LAUNCH_CHECK("sgemm");
  return XR_OK;
}

// out[i*ld_out] = a_i . b_i, accumulated in EXACTLY the order sgemm_kernel uses for one output
// element (a single fmaf chain over k = 0..dim-1).  In the reference the positive and the
// negatives of a row come out of one bmm, so a pool entry that is the row's own positive item
// has a bit-identical logit and the strict '<' of losses.py:292 masks it; keeping the two
// reduction orders identical preserves that tie.  One thread per row (M x dim FMAs: negligible).
template <typename T>
__global__ void rowdot_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t m,
                              int64_t dim, float* __restrict__ out, int64_t ld_out) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m;
       r += (int64_t)gridDim.x * blockDim.x) {
    const T* pa = a + r * dim;
    const T* pb = b + r * dim;
    float s = 0.f;
    for (int64_t c = 0; c < dim; ++c) s = fmaf(to_f32(pa[c]), to_f32(pb[c]), s);
    out[r * ld_out] = s;
  }
}

// dq_i (+)= g_pos_i * pos_i, then the cosine chain rule dq = inv_norm * (g - (g.qhat) qhat)
// (one warp per row; dq holds G_neg . Neg on entry)
template <typename T>
__global__ void dq_finalize_kernel(float* __restrict__ dq, const float* __restrict__ g_pos,
                                   int64_t ld_g, const T* __restrict__ pos,
                                   const T* __restrict__ qhat, const float* __restrict__ q_inv_norm,
                                   int64_t m, int64_t dim, int cosine) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < m; r += nwarps) {
    const float gp = g_pos[r * ld_g];
    float dotgq = 0.f;
    for (int64_t c = lane; c < dim; c += 32) {
      const float g = fmaf(gp, to_f32(pos[r * dim + c]), dq[r * dim + c]);
      dq[r * dim + c] = g;
      if (cosine) dotgq = fmaf(g, to_f32(qhat[r * dim + c]), dotgq);
    }
    if (cosine) {
      dotgq = warp_sum(dotgq);
      const float inv = q_inv_norm[r];
      for (int64_t c = lane; c < dim; c += 32) {
        const float g = dq[r * dim + c];
        dq[r * dim + c] = inv * (g - dotgq * to_f32(qhat[r * dim + c]));
      }
    }
  }
}

static inline int warp_grid(int64_t rows) {
  int64_t blocks = (rows * 32 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace xr

using namespace xr;

extern "C" int xr_logits_pool(const void* q, const void* pos, const void* neg, int64_t m,
                              int64_t cn, int64_t dim, int dtype, float* logits, int64_t ld,
                              void* stream) {
  XR_CHECK_ARG(q && pos && logits && (neg || cn == 0), "xr_logits_pool: null pointer");
  XR_CHECK_ARG(m >= 0 && cn >= 0 && dim > 0 && ld >= cn + 1, "xr_logits_pool: bad sizes");
  if (m == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  // layout: columns [0,cn) negatives, column cn the positive (keeps the GEMM operands aligned;
  // callers pass target_mode = explicit "last" to xr_rowloss)
  if (dtype == XR_F32) {
    rowdot_kernel<float><<<(unsigned)((m + 127) / 128), 128, 0, s>>>((const float*)q, (const float*)pos, m, dim,
                                                      logits + cn, ld);
    XR_LAUNCH_CHECK("rowdot");
    return launch_sgemm<float, float, false>((const float*)q, dim, (const float*)neg, dim, logits,
                                             ld, m, cn, dim, nullptr, nullptr, s);
  } else if (dtype == XR_BF16) {
    rowdot_kernel<__nv_bfloat16><<<(unsigned)((m + 127) / 128), 128, 0, s>>>(
        (const __nv_bfloat16*)q, (const __nv_bfloat16*)pos, m, dim, logits + cn, ld);
    XR_LAUNCH_CHECK("rowdot");
    return launch_sgemm<__nv_bfloat16, __nv_bfloat16, false>(
        (const __nv_bfloat16*)q, dim, (const __nv_bfloat16*)neg, dim, logits, ld, m, cn, dim,
        nullptr, nullptr, s);
  }
  XR_CHECK_ARG(false, "xr_logits_pool: bad dtype");
}

extern "C" int xr_dq_pool(const float* dlogits, int64_t ld, const void* q, const void* pos,
                          const void* neg, int64_t m, int64_t cn, int64_t dim, int dtype,
                          int cosine, const float* q_inv_norm, float* dq, void* stream) {
  XR_CHECK_ARG(dlogits && q && pos && dq && (neg || cn == 0), "xr_dq_pool: null pointer");
  XR_CHECK_ARG(!cosine || q_inv_norm, "xr_dq_pool: cosine needs q_inv_norm");
  if (m == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  int rc;
  if (cn == 0) {
    XR_CUDA(cudaMemsetAsync(dq, 0, sizeof(float) * m * dim, s));
  }
  if (dtype == XR_F32) {
    if (cn > 0) {
      rc = launch_sgemm<float, float, true>(dlogits, ld, (const float*)neg, dim, dq, dim, m, dim,
                                            cn, nullptr, nullptr, s);
      if (rc) return rc;
    }
    dq_finalize_kernel<float><<<warp_grid(m), 256, 0, s>>>(
        dq, dlogits + cn, ld, (const float*)pos, (const float*)q, q_inv_norm, m, dim, cosine);
  } else if (dtype == XR_BF16) {
    if (cn > 0) {
      rc = launch_sgemm<float, __nv_bfloat16, true>(dlogits, ld, (const __nv_bfloat16*)neg, dim,
                                                    dq, dim, m, dim, cn, nullptr, nullptr, s);
      if (rc) return rc;
    }
    dq_finalize_kernel<__nv_bfloat16><<<warp_grid(m), 256, 0, s>>>(
        dq, dlogits + cn, ld, (const __nv_bfloat16*)pos, (const __nv_bfloat16*)q, q_inv_norm, m,
        dim, cosine);
  } else {
    XR_CHECK_ARG(false, "xr_dq_pool: bad dtype");
  }
  XR_LAUNCH_CHECK("dq_finalize");
  return XR_OK;
}

extern "C" int xr_scores(const void* q, int64_t u, const void* catalog, int64_t n, int64_t dim,
                         int dtype, const float* q_inv_norm, const float* cat_inv_norm,
                         float* scores, int64_t ld, void* stream) {
  XR_CHECK_ARG(q && catalog && scores, "xr_scores: null pointer");
  XR_CHECK_ARG(u >= 0 && n >= 0 && dim > 0 && ld >= n, "xr_scores: bad sizes");
  cudaStream_t s = as_stream(stream);
  if (dtype == XR_F32)
    return launch_sgemm<float, float, false>((const float*)q, dim, (const float*)catalog, dim,
                                             scores, ld, u, n, dim, q_inv_norm, cat_inv_norm, s);
  if (dtype == XR_BF16)
    return launch_sgemm<__nv_bfloat16, __nv_bfloat16, false>(
        (const __nv_bfloat16*)q, dim, (const __nv_bfloat16*)catalog, dim, scores, ld, u, n, dim,
        q_inv_norm, cat_inv_norm, s);
  XR_CHECK_ARG(false, "xr_scores: bad dtype");
}
