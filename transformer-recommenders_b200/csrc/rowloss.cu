// The EmbedLoss pipeline on materialised logits, one thread block per row:
//   check_target          xfmr_rec/losses.py:233-261
//   mask_false_negatives  losses.py:283-292   (strict '<' against the target logit)
//   mine_hard_negatives   losses.py:311-330   (exact radix select; ties -> lower index)
//   loss() bodies         losses.py:352-372, 420-543 — all seven in one pass
//   LogitsStatistics      losses.py:383-405   — one device block, one D2H copy instead of 9 syncs
// plus dL/dlogits of one selected loss.  Rows stay L1/L2 resident across the passes.
// Row sums are written to a workspace and folded by a single block in a fixed order, so the
// results are run-to-run deterministic (no floating-point atomics).
#include "rowloss.cuh"

namespace xr {

__global__ void __launch_bounds__(RL_THREADS)
rowloss_kernel(const float* __restrict__ logits, int64_t m, int64_t c, int64_t ld, int target_mode,
               const int64_t* __restrict__ target, xr_loss_config cfg, uint32_t loss_mask,
               int grad_kind, float grad_scale, float* __restrict__ dlogits,
               double* __restrict__ row_out /* [m][ROW_SLOTS] */,
               uint8_t* __restrict__ maskbuf /* [gridDim.x][c] when hard mining */,
               int32_t* __restrict__ err_flag) {
  __shared__ double s_d[RL_WARPS];
  __shared__ float s_f[RL_WARPS];
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_sel[4];  // prefix key, remaining quota, scratch
  __shared__ int s_warp_cnt[RL_WARPS];

  const int tid = threadIdx.x;
  const bool hard = cfg.num_hard_negatives > 0 && (int64_t)cfg.num_hard_negatives < c;
  uint8_t* mrow = hard ? maskbuf + (int64_t)blockIdx.x * c : nullptr;

  for (int64_t i = blockIdx.x; i < m; i += gridDim.x) {
    const float* lrow = logits + i * ld;
    int64_t ti = 0;
    if (target_mode == XR_TARGET_DIAGONAL) ti = i;
    else if (target_mode == XR_TARGET_EXPLICIT) ti = target[i];
    else if (target_mode == 3 /* last */) ti = c - 1;
    if (ti < 0 || ti >= c) {  // torch.gather would raise; flag it and keep memory safe
      if (tid == 0 && err_flag) *err_flag = 1;
      ti = 0;
    }
    rowloss_row(lrow, c, ti, cfg, grad_kind, grad_scale, dlogits ? dlogits + i * ld : nullptr,
                row_out + i * ROW_SLOTS, mrow, RowLossScratch{s_d, s_f, s_hist, s_sel, s_warp_cnt});
    __syncthreads();
  }
}
// fold the per-row slots into the 7 loss sums and the stats block.  Two launches, both in a fixed order
// (deterministic): RR_BLOCKS blocks reduce contiguous row ranges into partials (the 1.9 MB of row slots at
// configs[1] took 46 us through ONE SM), one block folds the partials.
constexpr int RR_BLOCKS = 64, RR_THREADS = 256, RR_NS = 13;   // 6 losses + dens + pos sum/sq + ncount + nsum + nsq
constexpr int RR_STRIDE = RR_NS + 4;                          // + pos min/max, neg min/max
__global__ void __launch_bounds__(RR_THREADS)
rowloss_partial_kernel(const double* __restrict__ row_out, int64_t m, int64_t c, int n_hard,
                       double* __restrict__ partial, const int* __restrict__ dyn_m_cn,
                       const double* __restrict__ row_out2) {
  if (blockIdx.y == 1) {   // second logit family of a one-pass monitoring launch: its own rows and partials
    row_out = row_out2;
    partial += (size_t)RR_BLOCKS * RR_STRIDE;
  }
  if (dyn_m_cn) {   // shape known only on the device (sync-free step): {rows, pool size}
    m = dyn_m_cn[0];
    c = (int64_t)dyn_m_cn[1] + 1;
  }
  __shared__ double s_sum[RR_THREADS / 32][RR_NS];
  __shared__ double s_mm[RR_THREADS / 32][4];
  double acc[RR_NS];
#pragma unroll
  for (int k = 0; k < RR_NS; ++k) acc[k] = 0.0;
  double pmin = CUDART_INF, pmax = -CUDART_INF, nmin = CUDART_INF, nmax = -CUDART_INF;
  double num_neg = (double)(c - 1);
  if (n_hard > 0 && (double)n_hard < num_neg) num_neg = (double)n_hard;  // losses.py:387-389
  const int64_t per = (m + RR_BLOCKS - 1) / RR_BLOCKS;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = lo + per < m ? lo + per : m;
  for (int64_t i = lo + threadIdx.x; i < hi; i += RR_THREADS) {
    const double* o = row_out + i * ROW_SLOTS;
    acc[0] += o[S_ALIGN]; acc[1] += o[S_CONTR]; acc[2] += o[S_INFONCE]; acc[3] += o[S_NCE];
    acc[4] += o[S_HINGE]; acc[5] += o[S_LOGISTIC];
    acc[6] += o[S_DENS] / (num_neg + 1e-9);
    acc[7] += o[S_POS]; acc[8] += o[S_POS] * o[S_POS];
    acc[9] += o[S_NCOUNT]; acc[10] += o[S_NSUM]; acc[11] += o[S_NSQ];
    pmin = fmin(pmin, o[S_POS]); pmax = fmax(pmax, o[S_POS]);
    if (o[S_NCOUNT] > 0) { nmin = fmin(nmin, o[S_NMIN]); nmax = fmax(nmax, o[S_NMAX]); }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < RR_NS; ++k) acc[k] = warp_sum(acc[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pmin = fmin(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
    pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
    nmin = fmin(nmin, __shfl_xor_sync(0xffffffffu, nmin, o));
    nmax = fmax(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < RR_NS; ++k) s_sum[warp][k] = acc[k];
    s_mm[warp][0] = pmin; s_mm[warp][1] = pmax; s_mm[warp][2] = nmin; s_mm[warp][3] = nmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[RR_NS];
    for (int k = 0; k < RR_NS; ++k) tot[k] = 0.0;
    for (int w = 0; w < RR_THREADS / 32; ++w) {
      for (int k = 0; k < RR_NS; ++k) tot[k] += s_sum[w][k];
      pmin = fmin(pmin, s_mm[w][0]); pmax = fmax(pmax, s_mm[w][1]);
      nmin = fmin(nmin, s_mm[w][2]); nmax = fmax(nmax, s_mm[w][3]);
    }
    double* o = partial + (size_t)blockIdx.x * RR_STRIDE;
    for (int k = 0; k < RR_NS; ++k) o[k] = tot[k];
    o[RR_NS] = pmin; o[RR_NS + 1] = pmax; o[RR_NS + 2] = nmin; o[RR_NS + 3] = nmax;
  }
}

__global__ void rowloss_reduce_kernel(const double* __restrict__ partial, int64_t m, int64_t c, int n_hard,
                                      double* __restrict__ losses_out, double* __restrict__ stats_out,
                                      const int* __restrict__ dyn_m_cn, double* __restrict__ losses_out2) {
  if (blockIdx.x == 1) {   // second family: losses only
    partial += (size_t)RR_BLOCKS * RR_STRIDE;
    losses_out = losses_out2;
    stats_out = nullptr;
  }
  if (dyn_m_cn) {
    m = dyn_m_cn[0];
    c = (int64_t)dyn_m_cn[1] + 1;
  }
  // lane k owns column k of the partials (sums, then pos min / max, neg min / max): 64 independent loads per
  // lane in a fixed order instead of one thread walking 64 x 17 values (38 us measured)
  __shared__ double s_col[RR_STRIDE];
  const int k = threadIdx.x;
  if (k < RR_STRIDE) {
    const bool is_min = k == RR_NS || k == RR_NS + 2, is_max = k == RR_NS + 1 || k == RR_NS + 3;
    double a = is_min ? CUDART_INF : (is_max ? -CUDART_INF : 0.0);
#pragma unroll 16
    for (int b = 0; b < RR_BLOCKS; ++b) {   // fixed order
      const double x = partial[(size_t)b * RR_STRIDE + k];
      a = is_min ? fmin(a, x) : (is_max ? fmax(a, x) : a + x);
    }
    s_col[k] = a;
  }
  __syncwarp();
  if (threadIdx.x != 0) return;
  double num_neg = (double)(c - 1);
  if (n_hard > 0 && (double)n_hard < num_neg) num_neg = (double)n_hard;
  double tot[RR_NS];
  for (int j = 0; j < RR_NS; ++j) tot[j] = s_col[j];
  const double pmin = s_col[RR_NS], pmax = s_col[RR_NS + 1], nmin = s_col[RR_NS + 2], nmax = s_col[RR_NS + 3];
  if (losses_out) {
    losses_out[XR_LOSS_ALIGNMENT] = tot[0];
    losses_out[XR_LOSS_CONTRASTIVE] = tot[1];
    losses_out[XR_LOSS_ALIGNMENT_CONTRASTIVE] = tot[0] + tot[1];
    losses_out[XR_LOSS_INFONCE] = tot[2];
    losses_out[XR_LOSS_NCE] = tot[3];
    losses_out[XR_LOSS_PAIRWISE_HINGE] = tot[4];
    losses_out[XR_LOSS_PAIRWISE_LOGISTIC] = tot[5];
  }
  if (stats_out) {
    stats_out[0] = tot[6];
    stats_out[1] = (double)m;
    stats_out[2] = tot[7]; stats_out[3] = tot[8]; stats_out[4] = pmin; stats_out[5] = pmax;
    stats_out[6] = tot[9]; stats_out[7] = tot[10]; stats_out[8] = tot[11];
    stats_out[9] = nmin; stats_out[10] = nmax; stats_out[11] = num_neg;
    for (int k = 12; k < XR_STATS_SLOTS; ++k) stats_out[k] = 0.0;
  }
}

// shared with the tensor-core all-losses pass (fused_loss_sm100.cu), which fills the same row slots.
// `partial`: kRowlossPartialBytes of scratch.
int launch_rowloss_reduce(const double* row_out, int64_t m, int64_t c, int n_hard, double* losses_out,
                          double* stats_out, cudaStream_t s, const int* dyn_m_cn, double* partial) {
  rowloss_partial_kernel<<<RR_BLOCKS, RR_THREADS, 0, s>>>(row_out, m, c, n_hard, partial, dyn_m_cn, nullptr);
  XR_LAUNCH_CHECK("rowloss_partial");
  rowloss_reduce_kernel<<<1, 32, 0, s>>>(partial, m, c, n_hard, losses_out, stats_out, dyn_m_cn, nullptr);
  XR_LAUNCH_CHECK("rowloss_reduce");
  return XR_OK;
}

// two logit families at once (one-pass monitoring): row_out / losses_out + stats of the first, row_out2 /
// losses_out2 of the second; `partial` holds both families' partials (2 * RR_BLOCKS * RR_STRIDE doubles)
int launch_rowloss_reduce2(const double* row_out, const double* row_out2, int64_t m, int64_t c, double* losses_out,
                           double* stats_out, double* losses_out2, cudaStream_t s, const int* dyn_m_cn,
                           double* partial) {
  static_assert(2 * RR_BLOCKS * RR_STRIDE * sizeof(double) <= kRowlossPartialBytes * 2, "scratch");
  rowloss_partial_kernel<<<dim3(RR_BLOCKS, 2), RR_THREADS, 0, s>>>(row_out, m, c, 0, partial, dyn_m_cn, row_out2);
  XR_LAUNCH_CHECK("rowloss_partial");
  rowloss_reduce_kernel<<<2, 32, 0, s>>>(partial, m, c, 0, losses_out, stats_out, dyn_m_cn, losses_out2);
  XR_LAUNCH_CHECK("rowloss_reduce");
  return XR_OK;
}

static inline int rl_grid(int64_t m) {
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(m < cap ? (m < 1 ? 1 : m) : cap);
}

}  // namespace xr

using namespace xr;

extern "C" size_t xr_rowloss_workspace_bytes(int64_t m, int64_t c, int num_hard_negatives) {
  size_t b = (size_t)(m > 0 ? m : 1) * ROW_SLOTS * sizeof(double) + 256 + kRowlossPartialBytes;
  if (num_hard_negatives > 0 && num_hard_negatives < c)
    b += (size_t)rl_grid(m) * (size_t)c + 256;
  return b;
}

extern "C" int xr_rowloss(const float* logits, int64_t m, int64_t c, int64_t ld, int target_mode,
                          const int64_t* target, const xr_loss_config* cfg, uint32_t loss_mask,
                          int grad_kind, float grad_scale, float* dlogits, double* losses_out,
                          double* stats_out, int32_t* err_flag, void* workspace, void* stream) {
  XR_CHECK_ARG(logits && cfg && workspace, "xr_rowloss: null pointer");
  XR_CHECK_ARG(m >= 0 && c >= 1 && ld >= c, "xr_rowloss: bad sizes (m=%lld c=%lld ld=%lld)",
               (long long)m, (long long)c, (long long)ld);
  XR_CHECK_ARG(target_mode >= 0 && target_mode <= 3, "xr_rowloss: bad target_mode");
  XR_CHECK_ARG(target_mode != XR_TARGET_EXPLICIT || target, "xr_rowloss: explicit target is null");
  XR_CHECK_ARG(target_mode != XR_TARGET_DIAGONAL || m <= c,
               "xr_rowloss: diagonal targets need num_candidates >= batch");
  XR_CHECK_ARG(grad_kind < XR_NUM_LOSSES, "xr_rowloss: bad grad_kind");
  XR_CHECK_ARG(grad_kind < 0 || dlogits, "xr_rowloss: dlogits is null");
  cudaStream_t s = as_stream(stream);
  double* row_out = (double*)workspace;
  double* partial = (double*)((uint8_t*)workspace + (((size_t)(m > 0 ? m : 1) * ROW_SLOTS * sizeof(double) + 255) / 256) * 256);
  uint8_t* maskbuf = (uint8_t*)partial + kRowlossPartialBytes;
  if (m > 0) {
    rowloss_kernel<<<rl_grid(m), RL_THREADS, 0, s>>>(logits, m, c, ld, target_mode, target, *cfg,
                                                     loss_mask, grad_kind, grad_scale, dlogits,
                                                     row_out, maskbuf, err_flag);
    XR_LAUNCH_CHECK("rowloss");
  }
  if (losses_out || stats_out) {
    int rc = launch_rowloss_reduce(row_out, m, c, cfg->num_hard_negatives, losses_out, stats_out, s, nullptr, partial);
    if (rc) return rc;
  }
  return XR_OK;
}
