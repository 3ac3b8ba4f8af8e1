// One row of the EmbedLoss pipeline on its logits (shared by rowloss_kernel, which reads them from
// global memory, and the single-pass sampled-candidate kernel, which keeps them in shared memory):
//   mask_false_negatives  xfmr_rec/losses.py:283-292
//   mine_hard_negatives   losses.py:311-330   (exact radix select; ties -> lower index)
//   loss() bodies         losses.py:352-372, 420-543 -- all seven
//   LogitsStatistics      losses.py:383-405   ingredients
//   dL/dlogits of one selected loss
#pragma once
#include "common.cuh"

namespace xr {

constexpr int RL_THREADS = 256;
constexpr int RL_WARPS = RL_THREADS / 32;

__device__ __forceinline__ double block_sum(double v, double* s_buf) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
#pragma unroll
  for (int w = 0; w < RL_WARPS; ++w) r += s_buf[w];
  return r;
}
__device__ __forceinline__ float block_max(float v, float* s_buf) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s_buf[0];
#pragma unroll
  for (int w = 1; w < RL_WARPS; ++w) r = fmaxf(r, s_buf[w]);
  return r;
}
__device__ __forceinline__ float block_min(float v, float* s_buf) {
  v = warp_min(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s_buf[0];
#pragma unroll
  for (int w = 1; w < RL_WARPS; ++w) r = fminf(r, s_buf[w]);
  return r;
}

__device__ __forceinline__ float softplusf(float x) {
  return fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + __expf(-x)); }

struct RowLossScratch {   // block-shared scratch the caller declares
  double* s_d;         // [RL_WARPS]
  float* s_f;          // [RL_WARPS]
  unsigned* s_hist;    // [256]
  unsigned* s_sel;     // [4]
  int* s_warp_cnt;     // [RL_WARPS]
};

// All RL_THREADS threads of the block call this together.  lrow: the row's c logits; ti: target
// column; mrow: c bytes of scratch when hard mining is on; grow (nullable): dL/dlogits output;
// row_slots: ROW_SLOTS doubles for rowloss_reduce_kernel.
__device__ __forceinline__ void rowloss_row(const float* lrow, int64_t c, int64_t ti,
                                            const xr_loss_config& cfg, int grad_kind,
                                            float grad_scale, float* grow, double* row_slots,
                                            uint8_t* mrow, const RowLossScratch& sc) {
  double* s_d = sc.s_d;
  float* s_f = sc.s_f;
  unsigned* s_hist = sc.s_hist;
  unsigned* s_sel = sc.s_sel;
  int* s_warp_cnt = sc.s_warp_cnt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool round_bf16 = cfg.logits_bf16 != 0;
  const float scale = cfg.scale, margin = cfg.margin;
  const bool hard = cfg.num_hard_negatives > 0 && (int64_t)cfg.num_hard_negatives < c;
  auto val = [&](int64_t j) -> float {
    const float v = lrow[j];
    return round_bf16 ? bf16_round(v) : v;
  };
  const float t = val(ti);
  auto valid0 = [&](int64_t j, float v) -> bool {
    return cfg.mask_false_negatives ? (v < t) : (j != ti);
  };

  // ---- hard-negative mining: exact top-n_hard of the valid negatives ------------------------
  if (hard) {
    // count valid negatives first; if they fit, every valid negative survives (:326-328)
    int cnt = 0;
    for (int64_t j = tid; j < c; j += RL_THREADS) cnt += valid0(j, val(j));
    const double nvalid = block_sum((double)cnt, s_d);
    if (nvalid <= (double)cfg.num_hard_negatives) {
      for (int64_t j = tid; j < c; j += RL_THREADS) mrow[j] = valid0(j, val(j));
    } else {
      // MSB-first radix select of the n_hard-th largest key among valid negatives
      unsigned prefix = 0, prefix_mask = 0, want = (unsigned)cfg.num_hard_negatives;
      for (int shift = 24; shift >= 0; shift -= 8) {
        for (int b = tid; b < 256; b += RL_THREADS) s_hist[b] = 0;
        __syncthreads();
        for (int64_t j = tid; j < c; j += RL_THREADS) {
          const float v = val(j);
          if (!valid0(j, v)) continue;
          const unsigned k = float_key(v);
          if ((k & prefix_mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
          unsigned acc = 0;
          int b = 255;
          for (; b > 0; --b) {
            if (acc + s_hist[b] >= want) break;
            acc += s_hist[b];
          }
          s_sel[0] = (unsigned)b;
          s_sel[1] = want - acc;  // still to take inside bucket b
        }
        __syncthreads();
        prefix |= s_sel[0] << shift;
        prefix_mask |= 255u << shift;
        want = s_sel[1];
        __syncthreads();
      }
      // keys > prefix all survive; of the keys == prefix the first `want` by index survive
      const unsigned tau = prefix;
      int64_t taken = 0;  // equals taken so far (uniform across the block)
      for (int64_t base = 0; base < c; base += RL_THREADS) {
        const int64_t j = base + tid;
        bool v0 = false, eq = false, gt = false;
        if (j < c) {
          const float v = val(j);
          v0 = valid0(j, v);
          const unsigned k = float_key(v);
          eq = v0 && k == tau;
          gt = v0 && k > tau;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < RL_WARPS; ++w) {
          if (w < warp) before += s_warp_cnt[w];
          total += s_warp_cnt[w];
        }
        const int64_t rank = taken + before + __popc(bal & ((1u << lane) - 1u));
        if (j < c) mrow[j] = gt || (eq && rank < (int64_t)want);
        taken += total;
        __syncthreads();
      }
    }
    __syncthreads();
  }
  auto is_valid = [&](int64_t j, float v) -> bool { return hard ? (mrow[j] != 0) : valid0(j, v); };

  // ---- pass 1: count, extrema of the kept set ------------------------------------------------
  int cnt = 0;
  float vmax = -CUDART_INF_F, vmin = CUDART_INF_F;
  for (int64_t j = tid; j < c; j += RL_THREADS) {
    const float v = val(j);
    if (is_valid(j, v)) {
      ++cnt;
      vmax = fmaxf(vmax, v);
      vmin = fminf(vmin, v);
    }
  }
  const double n_valid = block_sum((double)cnt, s_d);
  const float neg_max = block_max(vmax, s_f);
  const float neg_min = block_min(vmin, s_f);
  auto scaled = [&](float v) -> float {
    const float z = __fmul_rn(v, scale);   // a separately rounded product, as torch's `logits * scale`
    return round_bf16 ? bf16_round(z) : z;  // autocast: `logits * scale` is a bf16 op (:486)
  };
  // reference max of the softmax set {valid} U {target}
  float zmax = scaled(t);
  if (n_valid > 0) zmax = fmaxf(zmax, fmaxf(scaled(neg_max), scaled(neg_min)));
  const float den = (float)n_valid + 1e-9f;
  // rounded on its own (torch: `(1 - margin) * target` is a tensor op): contracted into an FMA with the
  // subtraction below, a negative whose logit EQUALS tm would get x != 0 and flip the hinge
  const float tm = __fmul_rn(t, 1.0f - margin);

  // ---- pass 2: sums ----------------------------------------------------------------------------
  double a_exp = 0, a_sp = 0, a_hinge = 0, a_logi = 0, a_dlogi = 0, a_dhinge = 0, a_contr = 0,
         a_sum = 0, a_sq = 0;
  for (int64_t j = tid; j < c; j += RL_THREADS) {
    const float v = val(j);
    if (!is_valid(j, v)) continue;
    a_exp += (double)__expf(scaled(v) - zmax);
    a_sp += (double)softplusf(v);
    const float x = v - tm;
    a_hinge += (double)fmaxf(x, 0.f);
    a_dhinge += x > 0.f ? 1.0 : 0.0;
    a_logi += (double)softplusf(x);
    a_dlogi += (double)sigmoidf(x);
    a_contr += (double)fmaxf(v - 1.0f + margin, 0.f);
    a_sum += (double)v;
    a_sq += (double)v * (double)v;
  }
  const double z_neg = block_sum(a_exp, s_d);
  const double sum_sp = block_sum(a_sp, s_d);
  const double sum_hinge = block_sum(a_hinge, s_d);
  const double sum_dhinge = block_sum(a_dhinge, s_d);
  const double sum_logi = block_sum(a_logi, s_d);
  const double sum_dlogi = block_sum(a_dlogi, s_d);
  const double sum_contr = block_sum(a_contr, s_d);
  const double sum_v = block_sum(a_sum, s_d);
  const double sum_sq = block_sum(a_sq, s_d);

  const double z_all = z_neg + (double)__expf(scaled(t) - zmax);
  const double lse = (double)zmax + log(z_all);
  if (tid == 0) {
    double* o = row_slots;
    o[S_ALIGN] = 1.0 - (double)t;
    o[S_CONTR] = sum_contr / (double)den;
    o[S_INFONCE] = lse - (double)scaled(t);
    o[S_NCE] = (double)softplusf(-t) + sum_sp / (double)den;
    o[S_HINGE] = sum_hinge / (double)den;
    o[S_LOGISTIC] = sum_logi / (double)den;
    o[S_DENS] = n_valid;
    o[S_POS] = (double)t;
    o[S_NCOUNT] = n_valid;
    o[S_NSUM] = sum_v;
    o[S_NSQ] = sum_sq;
    o[S_NMIN] = (double)neg_min;
    o[S_NMAX] = (double)neg_max;
  }

  // ---- pass 3: dL/dlogits of the selected loss ------------------------------------------------
  if (grow && grad_kind >= 0) {
    const float inv_den = 1.0f / den;
    const float inv_z = (float)(1.0 / z_all);
    for (int64_t j = tid; j < c; j += RL_THREADS) {
      const float v = val(j);
      const bool ok = is_valid(j, v);
      float g = 0.f;
      switch (grad_kind) {
        case XR_LOSS_INFONCE:
          if (ok || j == ti) g = scale * (__expf(scaled(v) - zmax) * inv_z - (j == ti ? 1.f : 0.f));
          break;
        case XR_LOSS_NCE:
          if (ok) g = sigmoidf(v) * inv_den;
          if (j == ti) g += -sigmoidf(-t);
          break;
        case XR_LOSS_PAIRWISE_HINGE:
          if (ok) g = (v - tm > 0.f) ? inv_den : 0.f;
          if (j == ti) g = __fadd_rn(g, __fmul_rn(__fmul_rn(-(1.0f - margin), (float)sum_dhinge), inv_den));
          break;
        case XR_LOSS_PAIRWISE_LOGISTIC:
          if (ok) g = sigmoidf(v - tm) * inv_den;
          if (j == ti) g = __fadd_rn(g, __fmul_rn(__fmul_rn(-(1.0f - margin), (float)sum_dlogi), inv_den));
          break;
        case XR_LOSS_ALIGNMENT:
          if (j == ti) g = -1.f;
          break;
        case XR_LOSS_CONTRASTIVE:
        case XR_LOSS_ALIGNMENT_CONTRASTIVE:
          if (ok) g = (v - 1.0f + margin > 0.f) ? inv_den : 0.f;
          if (grad_kind == XR_LOSS_ALIGNMENT_CONTRASTIVE && j == ti) g += -1.f;
          break;
        default:
          break;
      }
      grow[j] = g * grad_scale;
    }
  }
}

}  // namespace xr
