// Family 1: embedding-row gathers and the index plumbing around them.
//
// Replaces nn.Embedding.forward at xfmr_rec/models.py:336-338, :400, :406, the boolean-mask
// compactions at models.py:392-416 and the normalisations inside losses.py:206-208.
// HBM-bound byte movers: 128-bit accesses, several independent loads in flight per thread,
// grids sized in multiples of the SM count.
#include "common.cuh"
#include "rownorm.cuh"

namespace xr {

// ------------------------------------------------------------------------------------------
// gather: one 16-byte OUTPUT vector per thread-iteration, UNROLL iterations in flight.
//   same dtype : 16B load -> 16B store (bit-exact copy)
//   f32 -> bf16: 2 x 16B loads -> 1 x 16B store (round to nearest even)
// Row r of the output comes from table row idx[sel ? sel[r] : r].
// ------------------------------------------------------------------------------------------
template <int MODE /*0 copy, 1 f32->bf16*/, int UNROLL>
__global__ void __launch_bounds__(256)
gather_rows_vec_kernel(const char* __restrict__ table, int64_t n_rows, int64_t row_bytes_in,
                       const int64_t* __restrict__ idx, const int64_t* __restrict__ sel,
                       int64_t n_out, int vec_per_row, char* __restrict__ out,
                       int32_t* __restrict__ err_flag) {
  const int64_t total = n_out * vec_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; v0 < total; v0 += stride * UNROLL) {
    int4 a[UNROLL], b[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t v = v0 + u * stride;
      ok[u] = v < total;
      a[u] = make_int4(0, 0, 0, 0);
      b[u] = a[u];
      if (ok[u]) {
        const int64_t r = v / vec_per_row;
        const int c = (int)(v - r * vec_per_row);
        int64_t src = idx[sel ? sel[r] : r];
        if (src < 0 || src >= n_rows) {
          if (err_flag) *err_flag = 1;
        } else {
          const char* p = table + src * row_bytes_in;
          if (MODE == 0) {
            a[u] = __ldg(reinterpret_cast<const int4*>(p) + c);
          } else {
            a[u] = __ldg(reinterpret_cast<const int4*>(p) + 2 * c);
            b[u] = __ldg(reinterpret_cast<const int4*>(p) + 2 * c + 1);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (!ok[u]) continue;
      const int64_t v = v0 + u * stride;
      int4 o;
      if (MODE == 0) {
        o = a[u];
      } else {
        const float* fa = reinterpret_cast<const float*>(&a[u]);
        const float* fb = reinterpret_cast<const float*>(&b[u]);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(fa[0], fa[1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(fa[2], fa[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(fb[0], fb[1]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(fb[2], fb[3]);
        o.x = *reinterpret_cast<int*>(&h0);
        o.y = *reinterpret_cast<int*>(&h1);
        o.z = *reinterpret_cast<int*>(&h2);
        o.w = *reinterpret_cast<int*>(&h3);
      }
      st_stream16(reinterpret_cast<int4*>(out) + v, o);
    }
  }
}

// scalar fallback for rows that are not 16-byte multiples / aligned
template <typename TI, typename TO>
__global__ void gather_rows_scalar_kernel(const TI* __restrict__ table, int64_t n_rows,
                                          int64_t dim, const int64_t* __restrict__ idx,
                                          const int64_t* __restrict__ sel, int64_t n_out,
                                          TO* __restrict__ out, int32_t* __restrict__ err_flag) {
  const int64_t total = n_out * dim;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / dim, c = e - r * dim;
    const int64_t src = idx[sel ? sel[r] : r];
    if (src < 0 || src >= n_rows) {
      if (err_flag) *err_flag = 1;
      out[e] = TO(0.f);
    } else {
      out[e] = TO(to_f32(table[src * dim + c]));
    }
  }
}

__global__ void scatter_rows_kernel(const float* __restrict__ src, int64_t n_src, int vec_per_row,
                                    const int64_t* __restrict__ sel, float* __restrict__ dst,
                                    int64_t n_dst_rows) {
  const int64_t total = n_src * vec_per_row;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = v / vec_per_row;
    const int c = (int)(v - r * vec_per_row);
    const int64_t d = sel[r];
    if (d < 0 || d >= n_dst_rows) continue;
    const int4 x = ld_stream16(reinterpret_cast<const int4*>(src) + v);
    reinterpret_cast<int4*>(dst)[d * vec_per_row + c] = x;
  }
}

// one warp per row: rownz[r] = any(x != 0)   (models.py:343)
template <typename T>
__global__ void row_nonzero_kernel(const T* __restrict__ x, int64_t n_rows, int64_t dim,
                                   uint8_t* __restrict__ rownz) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    bool nz = false;
    for (int64_t c = lane; c < dim; c += 32) nz |= (to_f32(x[r * dim + c]) != 0.0f);
    const unsigned any = __ballot_sync(0xffffffffu, nz);
    if (lane == 0) rownz[r] = any ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------
// stable compaction of the SeqBatch positions (models.py:343, 390, 398, 404, 413-416)
// three small kernels: per-chunk counts -> scan of chunk counts -> ordered write
// ------------------------------------------------------------------------------------------
constexpr int kChunk = 2048;        // positions per block
constexpr int kCompactThreads = 256;  // 8 positions per thread

__device__ __forceinline__ void position_flags(const int64_t* history_idx, const int64_t* pos_idx,
                                               const uint8_t* rownz, int64_t n_table_rows,
                                               int64_t p, bool& a, bool& q) {
  const int64_t h = history_idx[p];
  a = rownz ? (h >= 0 && h < n_table_rows && rownz[h] != 0) : (h != 0);
  q = a && (pos_idx[p] != 0);
}

__global__ void __launch_bounds__(kCompactThreads)
compact_count_kernel(const int64_t* __restrict__ history_idx, const int64_t* __restrict__ pos_idx,
                     const uint8_t* __restrict__ rownz, int64_t n_table_rows, int64_t n_pos,
                     int64_t* __restrict__ chunk_counts /* [2][nchunks] */, int64_t nchunks) {
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kChunk;
  int ca = 0, cq = 0;
  for (int i = threadIdx.x; i < kChunk; i += kCompactThreads) {
    const int64_t p = base + i;
    if (p < n_pos) {
      bool a, q;
      position_flags(history_idx, pos_idx, rownz, n_table_rows, p, a, q);
      ca += a;
      cq += q;
    }
  }
  ca = (int)warp_sum((float)ca);
  cq = (int)warp_sum((float)cq);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_cnt[0], ca);
    atomicAdd(&s_cnt[1], cq);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    chunk_counts[blockIdx.x] = s_cnt[0];
    chunk_counts[nchunks + blockIdx.x] = s_cnt[1];
  }
}

// single block: exclusive scan of the chunk counts in place, totals to counts[0..1]
__global__ void compact_scan_kernel(int64_t* __restrict__ chunk_counts, int64_t nchunks,
                                    int64_t* __restrict__ counts) {
  if (threadIdx.x < 2) {
    int64_t* c = chunk_counts + threadIdx.x * nchunks;
    int64_t run = 0;
    for (int64_t i = 0; i < nchunks; ++i) {
      const int64_t v = c[i];
      c[i] = run;
      run += v;
    }
    counts[threadIdx.x] = run;
  }
}

__global__ void __launch_bounds__(kCompactThreads)
compact_write_kernel(const int64_t* __restrict__ history_idx, const int64_t* __restrict__ pos_idx,
                     const uint8_t* __restrict__ rownz, int64_t n_table_rows, int64_t n_pos,
                     const int64_t* __restrict__ chunk_offsets, int64_t nchunks,
                     uint8_t* __restrict__ attn, int64_t* __restrict__ sel_attn,
                     int64_t* __restrict__ sel_pos, uint8_t* __restrict__ pos_mask,
                     int64_t* __restrict__ inv_pos) {
  // each thread owns 8 CONSECUTIVE positions so that a block-level scan of per-thread counts
  // yields a stable (ascending-position) compaction
  constexpr int PER = kChunk / kCompactThreads;
  __shared__ int s_a[kCompactThreads], s_q[kCompactThreads];
  const int64_t base = (int64_t)blockIdx.x * kChunk + (int64_t)threadIdx.x * PER;
  bool fa[PER], fq[PER];
  int ca = 0, cq = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int64_t p = base + i;
    fa[i] = fq[i] = false;
    if (p < n_pos) {
      position_flags(history_idx, pos_idx, rownz, n_table_rows, p, fa[i], fq[i]);
      attn[p] = fa[i];
    }
    ca += fa[i];
    cq += fq[i];
  }
  s_a[threadIdx.x] = ca;
  s_q[threadIdx.x] = cq;
  __syncthreads();
  // Hillis-Steele inclusive scan over 256 per-thread counts
  for (int off = 1; off < kCompactThreads; off <<= 1) {
    int va = 0, vq = 0;
    if ((int)threadIdx.x >= off) {
      va = s_a[threadIdx.x - off];
      vq = s_q[threadIdx.x - off];
    }
    __syncthreads();
    s_a[threadIdx.x] += va;
    s_q[threadIdx.x] += vq;
    __syncthreads();
  }
  int64_t oa = chunk_offsets[blockIdx.x] + s_a[threadIdx.x] - ca;
  int64_t oq = chunk_offsets[nchunks + blockIdx.x] + s_q[threadIdx.x] - cq;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int64_t p = base + i;
    if (fa[i]) {
      sel_attn[oa] = p;
      pos_mask[oa] = fq[i];
      ++oa;
    }
    if (fq[i]) {
      if (inv_pos) inv_pos[p] = oq;
      sel_pos[oq++] = p;
    } else if (inv_pos && p < n_pos) {
      inv_pos[p] = -1;
    }
  }
}

// The same compaction as ONE kernel when every chunk's block is resident at once (nchunks <= SM
// count: up to ~300k positions): each block scans its chunk, publishes its two totals as one 64-bit
// word (bit 63 = valid), waits for the words of the blocks before it, and writes its rows at the
// resulting offsets; the last block writes the two totals.  Same ordered result as the three-kernel
// form (count / scan / write), two launches fewer on the train step's critical path.
__global__ void __launch_bounds__(kCompactThreads)
compact_fused_kernel(const int64_t* __restrict__ history_idx, const int64_t* __restrict__ pos_idx,
                     const uint8_t* __restrict__ rownz, int64_t n_table_rows, int64_t n_pos,
                     unsigned long long* __restrict__ chunk_tot /* [nchunks], zeroed */,
                     uint8_t* __restrict__ attn, int64_t* __restrict__ sel_attn,
                     int64_t* __restrict__ sel_pos, uint8_t* __restrict__ pos_mask,
                     int64_t* __restrict__ inv_pos, int64_t* __restrict__ counts) {
  constexpr int PER = kChunk / kCompactThreads;
  __shared__ int s_a[kCompactThreads], s_q[kCompactThreads];
  __shared__ unsigned long long s_before;
  const int64_t base = (int64_t)blockIdx.x * kChunk + (int64_t)threadIdx.x * PER;
  bool fa[PER], fq[PER];
  int ca = 0, cq = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int64_t p = base + i;
    fa[i] = fq[i] = false;
    if (p < n_pos) {
      position_flags(history_idx, pos_idx, rownz, n_table_rows, p, fa[i], fq[i]);
      attn[p] = fa[i];
    }
    ca += fa[i];
    cq += fq[i];
  }
  s_a[threadIdx.x] = ca;
  s_q[threadIdx.x] = cq;
  if (threadIdx.x == 0) s_before = 0ull;
  __syncthreads();
  for (int off = 1; off < kCompactThreads; off <<= 1) {   // Hillis-Steele inclusive scan
    int va = 0, vq = 0;
    if ((int)threadIdx.x >= off) {
      va = s_a[threadIdx.x - off];
      vq = s_q[threadIdx.x - off];
    }
    __syncthreads();
    s_a[threadIdx.x] += va;
    s_q[threadIdx.x] += vq;
    __syncthreads();
  }
  const unsigned long long mine =
      ((unsigned long long)s_a[kCompactThreads - 1] << 32) | (unsigned long long)s_q[kCompactThreads - 1];
  if (threadIdx.x == 0) {
    // one 64-bit store carries data + valid bit: no separate flag, no fence ordering to get wrong
    atomicExch(chunk_tot + blockIdx.x, mine | 0x8000000000000000ull);
  }
  // totals of the chunks before this one (all blocks are resident: the spin cannot starve anyone)
  unsigned long long acc = 0ull;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += kCompactThreads) {
    unsigned long long v;
    do {
      v = *reinterpret_cast<volatile unsigned long long*>(chunk_tot + b);
    } while (!(v >> 63));
    acc += v & 0x7FFFFFFFFFFFFFFFull;   // the (a, q) halves never carry into each other: a, q < 2^31
  }
  if (acc) atomicAdd(&s_before, acc);
  __syncthreads();
  const unsigned long long before = s_before;
  int64_t oa = (int64_t)(before >> 32) + s_a[threadIdx.x] - ca;
  int64_t oq = (int64_t)(before & 0xFFFFFFFFull) + s_q[threadIdx.x] - cq;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int64_t p = base + i;
    if (fa[i]) {
      sel_attn[oa] = p;
      pos_mask[oa] = fq[i];
      ++oa;
    }
    if (fq[i]) {
      if (inv_pos) inv_pos[p] = oq;
      sel_pos[oq++] = p;
    } else if (inv_pos && p < n_pos) {
      inv_pos[p] = -1;
    }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    const unsigned long long tot = before + mine;
    counts[0] = (int64_t)(tot >> 32);
    counts[1] = (int64_t)(tot & 0xFFFFFFFFull);
  }
}

// dst[p,:] = inv[p] >= 0 ? cast(src[inv[p],:] * *scale) : 0  — the whole backward of the query
// compaction (models.py:392, 415) in one pass: zero fill + scatter + grad_output scale + cast.
template <typename TO>
__global__ void __launch_bounds__(256)
scatter_scaled_kernel(const float* __restrict__ src, const int64_t* __restrict__ inv,
                      const float* __restrict__ scale, int64_t n_dst_rows, int vec_per_row,
                      TO* __restrict__ dst) {
  const float sc = scale ? *scale : 1.0f;
  const int64_t total = n_dst_rows * vec_per_row;   // vec = 8 output elements
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = v / vec_per_row;
    const int c = (int)(v - r * vec_per_row);
    const int64_t s = inv[r];
    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (s >= 0) {
      const int4 a = ld_stream16(reinterpret_cast<const int4*>(src + (s * vec_per_row + c) * 8));
      const int4 b = ld_stream16(reinterpret_cast<const int4*>(src + (s * vec_per_row + c) * 8 + 4));
      x[0] = __int_as_float(a.x) * sc; x[1] = __int_as_float(a.y) * sc;
      x[2] = __int_as_float(a.z) * sc; x[3] = __int_as_float(a.w) * sc;
      x[4] = __int_as_float(b.x) * sc; x[5] = __int_as_float(b.y) * sc;
      x[6] = __int_as_float(b.z) * sc; x[7] = __int_as_float(b.w) * sc;
    }
    if (sizeof(TO) == 2) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(x[0], x[1]), h1 = __floats2bfloat162_rn(x[2], x[3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(x[4], x[5]), h3 = __floats2bfloat162_rn(x[6], x[7]);
      int4 o;
      o.x = *reinterpret_cast<int*>(&h0); o.y = *reinterpret_cast<int*>(&h1);
      o.z = *reinterpret_cast<int*>(&h2); o.w = *reinterpret_cast<int*>(&h3);
      st_stream16(reinterpret_cast<int4*>(dst) + v, o);
    } else {
      int4 o0 = make_int4(__float_as_int(x[0]), __float_as_int(x[1]), __float_as_int(x[2]), __float_as_int(x[3]));
      int4 o1 = make_int4(__float_as_int(x[4]), __float_as_int(x[5]), __float_as_int(x[6]), __float_as_int(x[7]));
      st_stream16(reinterpret_cast<int4*>(dst) + 2 * v, o0);
      st_stream16(reinterpret_cast<int4*>(dst) + 2 * v + 1, o1);
    }
  }
}

// one warp per row: L2 normalise (losses.py:206-208 / index.py:47)
template <typename TI, typename TO>
__global__ void normalize_rows_kernel(const TI* __restrict__ x, int64_t n_rows, int64_t dim,
                                      float eps, TO* __restrict__ y,
                                      float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const TI* xr_ = x + r * dim;
    float ss = 0.f;
    for (int64_t c = lane; c < dim; c += 32) {
      const float v = to_f32(xr_[c]);
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    if (inv_norm && lane == 0) inv_norm[r] = inv;
    if (y) {
      for (int64_t c = lane; c < dim; c += 32) y[r * dim + c] = TO(to_f32(xr_[c]) * inv);
    }
  }
}

// D = 384, bf16 -> bf16: the vectorised row routine the sync-free step uses as well (same bits)
__global__ void __launch_bounds__(256)
normalize_rows384_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t n_rows, float eps,
                              __nv_bfloat16* __restrict__ y, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const float inv = normalize_row384_bf16(x + r * 384, y ? y + r * 384 : nullptr, eps, lane);
    if (inv_norm && lane == 0) inv_norm[r] = inv;
  }
}

static inline int grid_for(int64_t work_items, int threads, int per_sm = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;  // grid-stride loops: a whole number of waves
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace xr

using namespace xr;

extern "C" int xr_gather_rows(const void* table, int64_t n_rows, int64_t dim, int table_dtype,
                              const int64_t* idx, const int64_t* sel, int64_t n_out, void* out,
                              int out_dtype, int32_t* err_flag, void* stream) {
  XR_CHECK_ARG(table && idx && out, "xr_gather_rows: null pointer");
  XR_CHECK_ARG(n_rows > 0 && dim > 0 && n_out >= 0, "xr_gather_rows: bad sizes");
  XR_CHECK_ARG((table_dtype == XR_F32 || table_dtype == XR_BF16) &&
                   (out_dtype == XR_F32 || out_dtype == XR_BF16),
               "xr_gather_rows: bad dtype");
  XR_CHECK_ARG(!(table_dtype == XR_BF16 && out_dtype == XR_F32),
               "xr_gather_rows: bf16 -> f32 widening gather is not provided");
  if (n_out == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  const int64_t in_bytes = dim * (table_dtype == XR_F32 ? 4 : 2);
  const int64_t out_bytes = dim * (out_dtype == XR_F32 ? 4 : 2);
  const bool aligned = ((uintptr_t)table % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                       (out_bytes % 16 == 0) && (in_bytes % 16 == 0);
  if (aligned) {
    const int vpr = (int)(out_bytes / 16);
    const int64_t total = n_out * vpr;
    const int grid = grid_for((total + 3) / 4, 256);
    if (table_dtype == out_dtype)
      gather_rows_vec_kernel<0, 4><<<grid, 256, 0, s>>>((const char*)table, n_rows, in_bytes, idx,
                                                        sel, n_out, vpr, (char*)out, err_flag);
    else
      gather_rows_vec_kernel<1, 4><<<grid, 256, 0, s>>>((const char*)table, n_rows, in_bytes, idx,
                                                        sel, n_out, vpr, (char*)out, err_flag);
  } else {
    const int grid = grid_for(n_out * dim, 256);
    if (table_dtype == XR_F32 && out_dtype == XR_F32)
      gather_rows_scalar_kernel<float, float><<<grid, 256, 0, s>>>(
          (const float*)table, n_rows, dim, idx, sel, n_out, (float*)out, err_flag);
    else if (table_dtype == XR_BF16 && out_dtype == XR_BF16)
      gather_rows_scalar_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(
          (const __nv_bfloat16*)table, n_rows, dim, idx, sel, n_out, (__nv_bfloat16*)out, err_flag);
    else
      gather_rows_scalar_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>(
          (const float*)table, n_rows, dim, idx, sel, n_out, (__nv_bfloat16*)out, err_flag);
  }
  XR_LAUNCH_CHECK("gather_rows");
  return XR_OK;
}

extern "C" int xr_scatter_rows(const float* src, int64_t n_src, int64_t dim, const int64_t* sel,
                               float* dst, int64_t n_dst_rows, void* stream) {
  XR_CHECK_ARG(src && sel && dst, "xr_scatter_rows: null pointer");
  XR_CHECK_ARG(dim > 0 && dim % 4 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0,
               "xr_scatter_rows: dim must be a multiple of 4 and buffers 16B aligned");
  if (n_src == 0) return XR_OK;
  const int vpr = (int)(dim / 4);
  scatter_rows_kernel<<<grid_for(n_src * vpr, 256), 256, 0, as_stream(stream)>>>(
      src, n_src, vpr, sel, dst, n_dst_rows);
  XR_LAUNCH_CHECK("scatter_rows");
  return XR_OK;
}

extern "C" int xr_scatter_scaled(const float* src, const int64_t* inv_pos, const float* scale,
                                 int64_t n_dst_rows, int64_t dim, void* dst, int dst_dtype,
                                 void* stream) {
  XR_CHECK_ARG(src && inv_pos && dst, "xr_scatter_scaled: null pointer");
  XR_CHECK_ARG(dim > 0 && dim % 8 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0,
               "xr_scatter_scaled: dim must be a multiple of 8 and buffers 16B aligned");
  if (n_dst_rows == 0) return XR_OK;
  const int vpr = (int)(dim / 8);
  const int grid = grid_for(n_dst_rows * vpr, 256);
  if (dst_dtype == XR_BF16)
    scatter_scaled_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(
        src, inv_pos, scale, n_dst_rows, vpr, (__nv_bfloat16*)dst);
  else if (dst_dtype == XR_F32)
    scatter_scaled_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(src, inv_pos, scale,
                                                                      n_dst_rows, vpr, (float*)dst);
  else
    XR_CHECK_ARG(false, "xr_scatter_scaled: bad dtype");
  XR_LAUNCH_CHECK("scatter_scaled");
  return XR_OK;
}

extern "C" int xr_row_nonzero(const void* table, int64_t n_rows, int64_t dim, int dtype,
                              uint8_t* rownz, void* stream) {
  XR_CHECK_ARG(table && rownz && n_rows > 0 && dim > 0, "xr_row_nonzero: bad arguments");
  const int grid = grid_for(n_rows * 32, 256);
  if (dtype == XR_F32)
    row_nonzero_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)table, n_rows,
                                                                   dim, rownz);
  else if (dtype == XR_BF16)
    row_nonzero_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)table, n_rows, dim, rownz);
  else
    XR_CHECK_ARG(false, "xr_row_nonzero: bad dtype");
  XR_LAUNCH_CHECK("row_nonzero");
  return XR_OK;
}

extern "C" size_t xr_compact_workspace_bytes(int64_t n_pos) {
  const int64_t nchunks = (n_pos + kChunk - 1) / kChunk;
  return (size_t)(2 * (nchunks > 0 ? nchunks : 1)) * sizeof(int64_t);
}

extern "C" int xr_compact_positions(const int64_t* history_idx, const int64_t* pos_idx,
                                    const uint8_t* rownz, int64_t n_table_rows, int64_t n_pos,
                                    uint8_t* attn, int64_t* sel_attn, int64_t* sel_pos,
                                    uint8_t* pos_mask, int64_t* inv_pos, int64_t* counts,
                                    void* workspace, void* stream) {
  XR_CHECK_ARG(history_idx && pos_idx && attn && sel_attn && sel_pos && pos_mask && counts &&
                   workspace,
               "xr_compact_positions: null pointer");
  XR_CHECK_ARG(n_pos >= 0, "xr_compact_positions: bad size");
  cudaStream_t s = as_stream(stream);
  if (n_pos == 0) {
    XR_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int64_t), s));
    return XR_OK;
  }
  const int64_t nchunks = (n_pos + kChunk - 1) / kChunk;
  int64_t* cc = (int64_t*)workspace;
  if (nchunks <= sm_count()) {   // every block resident at once: single-pass form
    XR_CUDA(cudaMemsetAsync(cc, 0, (size_t)nchunks * sizeof(int64_t), s));
    compact_fused_kernel<<<(unsigned)nchunks, kCompactThreads, 0, s>>>(
        history_idx, pos_idx, rownz, n_table_rows, n_pos, reinterpret_cast<unsigned long long*>(cc), attn,
        sel_attn, sel_pos, pos_mask, inv_pos, counts);
    XR_LAUNCH_CHECK("compact_fused");
    return XR_OK;
  }
  compact_count_kernel<<<(unsigned)nchunks, kCompactThreads, 0, s>>>(
      history_idx, pos_idx, rownz, n_table_rows, n_pos, cc, nchunks);
  XR_LAUNCH_CHECK("compact_count");
  compact_scan_kernel<<<1, 32, 0, s>>>(cc, nchunks, counts);
  XR_LAUNCH_CHECK("compact_scan");
  compact_write_kernel<<<(unsigned)nchunks, kCompactThreads, 0, s>>>(
      history_idx, pos_idx, rownz, n_table_rows, n_pos, cc, nchunks, attn, sel_attn, sel_pos,
      pos_mask, inv_pos);
  XR_LAUNCH_CHECK("compact_write");
  return XR_OK;
}

extern "C" int xr_normalize_rows(const void* x, int64_t n_rows, int64_t dim, int x_dtype,
                                 float eps, void* y, int y_dtype, float* inv_norm, void* stream) {
  XR_CHECK_ARG(x && (y || inv_norm), "xr_normalize_rows: null pointer");
  XR_CHECK_ARG(n_rows >= 0 && dim > 0, "xr_normalize_rows: bad sizes");
  if (n_rows == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(n_rows * 32, 256);
  if (x_dtype == XR_F32 && y_dtype == XR_F32)
    normalize_rows_kernel<float, float><<<grid, 256, 0, s>>>((const float*)x, n_rows, dim, eps,
                                                             (float*)y, inv_norm);
  else if (x_dtype == XR_F32 && y_dtype == XR_BF16)
    normalize_rows_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>(
        (const float*)x, n_rows, dim, eps, (__nv_bfloat16*)y, inv_norm);
  else if (x_dtype == XR_BF16 && y_dtype == XR_BF16 && dim == 384 && (uintptr_t)x % 16 == 0 &&
           (uintptr_t)y % 16 == 0)
    normalize_rows384_bf16_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, n_rows, eps,
                                                       (__nv_bfloat16*)y, inv_norm);
  else if (x_dtype == XR_BF16 && y_dtype == XR_BF16)
    normalize_rows_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(
        (const __nv_bfloat16*)x, n_rows, dim, eps, (__nv_bfloat16*)y, inv_norm);
  else if (x_dtype == XR_BF16 && y_dtype == XR_F32)
    normalize_rows_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>(
        (const __nv_bfloat16*)x, n_rows, dim, eps, (float*)y, inv_norm);
  else
    XR_CHECK_ARG(false, "xr_normalize_rows: bad dtype");
  XR_LAUNCH_CHECK("normalize_rows");
  return XR_OK;
}
