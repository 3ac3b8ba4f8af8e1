// Sequence encoder of the reference (xfmr_rec/models.py:51-102, 306-345): a HuggingFace BertModel built with
// is_decoder=True (causal self-attention) fed `inputs_embeds` from the frozen item table.  This file holds
// the hand-written kernels around the encoder's linear layers:
//
//   embed_ln        x0 = LayerNorm(table[idx] + position_emb[l] + token_type_emb[0])   -- the history gather
//                   of models.py:336-338 fused into the first layer's input, + the attention mask of :343
//   attention       causal + key-padding multi-head self-attention, head_dim 32, forward and backward
//                   (deterministic: one pass per query for dQ, one pass per key for dK / dV, no atomics)
//   gelu            exact (erf) GELU forward / backward on the FFN's intermediate activations
//   add_ln          LayerNorm(y + residual) forward / backward (BertSelfOutput / BertOutput)
//
// The linear layers themselves (QKV, attention output, FFN up / down and their weight / input gradients)
// are plain GEMMs and go through cuBLAS (torch.nn.functional.linear) in xfmr_rec_b200/encoder.py.
// Activations are fp32 or bf16 (template parameter); all reductions are fp32.
#include "common.cuh"
#include "philox.cuh"

namespace xr {

namespace enc {
constexpr int H = 384;           // hidden size (all-MiniLM-L6-v2, params.py:11)
constexpr int HD = 32;           // head dim (12 heads)
constexpr int PER = H / 32;      // hidden elements per lane of a warp-per-token kernel
constexpr int MAX_L = 384;      // the backward keeps Q, K, V, dO of one head in shared memory (206 KB at 384)
}  // namespace enc

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// ---- dropout (HF BertConfig defaults: hidden_dropout_prob = attention_probs_dropout_prob = 0.1, active in
// training mode; models.py:92-101 builds the BertModel with them).  Masks are a pure function of
// (seed, step counter, site, element coordinates) through Philox4x32-10, so the backward pass recomputes them
// instead of storing them, and CUDA-graph replays draw fresh masks because {seed, counter} live in device
// memory (`rng`; a null pointer or p == 0 switches the site off).
struct DropCtx {
  uint32_t k0, k1, site, ctr, thr;
  float scale;
  bool on;
};
// bits: 16 (hidden sites: keep iff r16 >= thr, p_eff = thr / 65536) or 8 (attention: p_eff = thr / 256)
__device__ __forceinline__ DropCtx make_drop(const int64_t* __restrict__ rng, float p, int site, int bits) {
  DropCtx d;
  d.on = rng != nullptr && p > 0.f;
  d.k0 = d.k1 = d.ctr = d.thr = 0;
  d.site = (uint32_t)site;
  d.scale = 1.f;
  if (d.on) {
    const uint64_t seed = (uint64_t)rng[0], ctr = (uint64_t)rng[1];
    d.k0 = (uint32_t)seed;
    d.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(ctr >> 32);
    d.ctr = (uint32_t)ctr;
    const float full = bits == 16 ? 65536.f : 256.f;
    d.thr = (uint32_t)(p * full + 0.5f);
    if (d.thr >= (uint32_t)full) d.thr = (uint32_t)full - 1;
    d.scale = full / (full - (float)d.thr);
  }
  return d;
}
// hidden-state sites (warp per token, lane-strided columns c = lane + 32 k): bit k of the result = keep
// element k of this lane.  Two Philox calls per lane and token give its twelve 16-bit draws.
__device__ __forceinline__ uint32_t drop_bits(int64_t tok, int lane, const DropCtx& d) {
  if (!d.on) return 0xFFFFFFFFu;
  uint32_t bits = 0;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const U4 r = philox((uint32_t)tok, (uint32_t)((uint64_t)tok >> 32) ^ (uint32_t)(lane | (q << 8)), d.site, d.ctr,
                        d.k0, d.k1);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int k = 8 * q + h;
      if (k < enc::PER && ((w[h >> 1] >> (16 * (h & 1))) & 0xFFFFu) >= d.thr) bits |= 1u << k;
    }
  }
  return bits;
}
__device__ __forceinline__ void drop_apply(float (&x)[enc::PER], uint32_t bits, const DropCtx& d) {
  if (!d.on) return;
#pragma unroll
  for (int k = 0; k < enc::PER; ++k) x[k] = (bits >> k) & 1u ? x[k] * d.scale : 0.f;
}
__device__ __forceinline__ void drop_row(float (&x)[enc::PER], int64_t tok, int lane, const DropCtx& d) {
  drop_apply(x, drop_bits(tok, lane, d), d);
}
// attention-probability site: the 16 x 16 block (query tile, key block) of one (sequence, head) draws 256
// bytes from 16 Philox calls laid out so that BOTH fragment layouts of the tensor-core kernels (queries along
// rows in the forward / dQ pass, keys along rows in the dK / dV pass) find the eight elements a lane holds in
// ONE call: call = ((r % 8) / 2) * 4 + (c % 8) / 2, byte = (r % 2) * 8 + (r / 8) * 4 + (c % 2) * 2 + c / 8 for
// the element at (query r, key c) of the block.
struct DropBlock {
  uint32_t w[4];
};
__device__ __forceinline__ DropBlock attn_drop_call(const DropCtx& d, uint32_t bh, uint32_t q_tile, uint32_t k_tile,
                                                    int call) {
  const U4 r = philox(bh, (q_tile << 16) | k_tile, d.site | ((uint32_t)call << 8), d.ctr, d.k0, d.k1);
  return DropBlock{{r.x, r.y, r.z, r.w}};
}
__device__ __forceinline__ bool attn_keep(const DropBlock& b, int slot, const DropCtx& d) {
  return ((b.w[slot >> 2] >> (8 * (slot & 3))) & 0xFFu) >= d.thr;
}
// forward / dQ-pass fragment: a lane's eight elements are bytes (g % 2) * 8 + [0, 8) of its call: pick that
// half once (static register indexing afterwards); local slot = (e >> 1) * 4 + (e & 1) * 2 + nt
__device__ __forceinline__ uint64_t attn_row_half(const DropBlock& b, int g) {
  const uint32_t lo = (g & 1) ? b.w[2] : b.w[0], hi = (g & 1) ? b.w[3] : b.w[1];
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ bool attn_keep8(uint64_t half, int local_slot, const DropCtx& d) {
  return ((uint32_t)(half >> (8 * local_slot)) & 0xFFu) >= d.thr;
}
// one element at a time (the fp32 kernels): query i, key j of (sequence, head) bh
__device__ __forceinline__ float attn_drop_elem(const DropCtx& d, uint32_t bh, int i, int j) {
  if (!d.on) return 1.f;
  const int r = i & 15, c = j & 15;
  const DropBlock b = attn_drop_call(d, bh, (uint32_t)(i >> 4), (uint32_t)(j >> 4), ((r & 7) >> 1) * 4 + ((c & 7) >> 1));
  return attn_keep(b, (r & 1) * 8 + (r >> 3) * 4 + (c & 1) * 2 + (c >> 3), d) ? d.scale : 0.f;
}

// warp-per-token LayerNorm over x[PER] (lane-strided columns c = lane + 32 k), HF semantics: biased variance
__device__ __forceinline__ void ln_row(const float (&x)[enc::PER], const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, int lane, float (&y)[enc::PER],
                                       float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < enc::PER; ++k) s += x[k];
  mean = warp_sum(s) * (1.0f / enc::H);
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < enc::PER; ++k) {
    const float d = x[k] - mean;
    v = fmaf(d, d, v);
  }
  rstd = rsqrtf(warp_sum(v) * (1.0f / enc::H) + eps);
#pragma unroll
  for (int k = 0; k < enc::PER; ++k) {
    const int c = lane + 32 * k;
    y[k] = (x[k] - mean) * rstd * gamma[c] + beta[c];
  }
}

// ---- embeddings: gather + position + token type + LayerNorm ------------------------------------------------
__global__ void __launch_bounds__(256)
enc_embed_ln_fwd_kernel(const float* __restrict__ table, int64_t n_table_rows, const int64_t* __restrict__ idx,
                        const float* __restrict__ pos_emb, const float* __restrict__ type_emb,
                        const float* __restrict__ gamma, const float* __restrict__ beta, int64_t n_tok, int seq_len,
                        float eps, float* __restrict__ out, __nv_bfloat16* __restrict__ out_lp,
                        float* __restrict__ stats, uint8_t* __restrict__ mask, int32_t* __restrict__ err_flag,
                        const int64_t* __restrict__ rng, float drop_p, int site) {
  const DropCtx dc = make_drop(rng, drop_p, site, 16);
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < n_tok; t += nwarps) {
    int64_t row = idx[t];
    if (row < 0 || row >= n_table_rows) {
      if (err_flag) *err_flag = 1;
      row = 0;
    }
    const int l = (int)(t % seq_len);
    float x[enc::PER], y[enc::PER];
    bool nz = false;
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) {
      const int c = lane + 32 * k;
      const float e = __ldg(table + row * enc::H + c);
      nz |= (e != 0.f);
      x[k] = e + pos_emb[(int64_t)l * enc::H + c] + type_emb[c];
    }
    nz = __any_sync(0xffffffffu, nz);     // models.py:343: (inputs_embeds != 0).any(-1)
    float mean, rstd;
    ln_row(x, gamma, beta, eps, lane, y, mean, rstd);
    drop_row(y, t, lane, dc);        // BertEmbeddings: dropout(LayerNorm(...))
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) out[t * enc::H + lane + 32 * k] = y[k];
    if (out_lp)
#pragma unroll
      for (int k = 0; k < enc::PER; ++k) out_lp[t * enc::H + lane + 32 * k] = __float2bfloat16_rn(y[k]);
    if (lane == 0) {
      stats[2 * t] = mean;
      stats[2 * t + 1] = rstd;
      mask[t] = nz ? 1 : 0;
    }
  }
}

// LayerNorm backward of one token row: dx (returned in registers) + this warp's contributions to dgamma / dbeta
__device__ __forceinline__ void ln_row_bwd(const float (&x)[enc::PER], const float (&dy)[enc::PER],
                                           const float* __restrict__ gamma, float mean, float rstd, int lane,
                                           float (&dx)[enc::PER], float (&dg)[enc::PER], float (&db)[enc::PER]) {
  float xh[enc::PER], dxh[enc::PER];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < enc::PER; ++k) {
    const int c = lane + 32 * k;
    xh[k] = (x[k] - mean) * rstd;
    dxh[k] = dy[k] * gamma[c];
    s1 += dxh[k];
    s2 = fmaf(dxh[k], xh[k], s2);
    dg[k] += dy[k] * xh[k];
    db[k] += dy[k];
  }
  s1 = warp_sum(s1) * (1.0f / enc::H);
  s2 = warp_sum(s2) * (1.0f / enc::H);
#pragma unroll
  for (int k = 0; k < enc::PER; ++k) dx[k] = rstd * (dxh[k] - s1 - xh[k] * s2);
}

constexpr int LNB_BLOCKS = 888;   // partial rows of the parameter-gradient reductions (six blocks per SM: the
                                  // backward kernels recompute dropout masks, two blocks per SM left them latency-bound)

// fold the warps' column sums (registers; vector v of NV) into the block's partial rows part[block][v][H]:
// fixed order inside the block
template <int NV>
__device__ __forceinline__ void ln_param_partials(const float (&acc)[NV][enc::PER],
                                                  float* __restrict__ part /* [gridDim.x][NV][H] */) {
  __shared__ float s_acc[8][NV][enc::H];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) s_acc[warp][v][lane + 32 * k] = acc[v][k];
  __syncthreads();
  for (int e = threadIdx.x; e < NV * enc::H; e += blockDim.x) {
    const int v = e / enc::H, c = e % enc::H;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += s_acc[w][v][c];
    part[((size_t)blockIdx.x * NV + v) * enc::H + c] = a;
  }
}

// out_v[c] = sum over the n_part partial rows part[p][v][c]; a block owns 32 columns of one vector, its 32 warps
// take every 32nd partial row, then one fixed-order sum over the warps: deterministic, ~30 loads per thread
constexpr int FOLD_WARPS = 32;
__global__ void __launch_bounds__(FOLD_WARPS * 32)
enc_fold_partials_kernel(const float* __restrict__ part, int n_part, int n_vec, int width, float* __restrict__ out0,
                         float* __restrict__ out1, float* __restrict__ out2) {
  __shared__ float s[FOLD_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = blockIdx.y, c = blockIdx.x * 32 + lane;
  float a = 0.f;
  if (c < width)
#pragma unroll 8
    for (int p = warp; p < n_part; p += FOLD_WARPS) a += part[((size_t)p * n_vec + v) * width + c];
  s[warp][lane] = a;
  __syncthreads();
  if (warp == 0 && c < width) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < FOLD_WARPS; ++w) t += s[w][lane];
    float* out = v == 0 ? out0 : (v == 1 ? out1 : out2);
    if (out) out[c] = t;
  }
}
static inline void launch_fold(const float* part, int n_part, int n_vec, int width, float* o0, float* o1, float* o2,
                               cudaStream_t s) {
  enc_fold_partials_kernel<<<dim3((width + 31) / 32, n_vec), FOLD_WARPS * 32, 0, s>>>(part, n_part, n_vec, width, o0, o1,
                                                                                     o2);
}

// backward of embed_ln: dgamma, dbeta (partials), d position_emb[l] (sum over the batch), d token_type_emb[0]
// (sum over all tokens).  The table is frozen (models.py:251-253): no gradient for it.  dx rows are written
// to dx_out (scratch) so that the column reductions over the batch run as a second, coalesced kernel.
__global__ void __launch_bounds__(256)
enc_embed_ln_bwd_kernel(const float* __restrict__ table, int64_t n_table_rows, const int64_t* __restrict__ idx,
                        const float* __restrict__ pos_emb, const float* __restrict__ type_emb,
                        const float* __restrict__ gamma, const float* __restrict__ stats,
                        const float* __restrict__ dout, const __nv_bfloat16* __restrict__ dout_lp, int64_t n_tok,
                        int seq_len, float* __restrict__ dx_out, float* __restrict__ part,
                        const int64_t* __restrict__ rng, float drop_p, int site) {
  const DropCtx dc = make_drop(rng, drop_p, site, 16);
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float acc[2][enc::PER] = {};
  for (int64_t t = warp; t < n_tok; t += nwarps) {
    int64_t row = idx[t];
    if (row < 0 || row >= n_table_rows) row = 0;
    const int l = (int)(t % seq_len);
    float x[enc::PER], dy[enc::PER], dx[enc::PER];
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) {
      const int c = lane + 32 * k;
      x[k] = __ldg(table + row * enc::H + c) + pos_emb[(int64_t)l * enc::H + c] + type_emb[c];
      dy[k] = (dout ? dout[t * enc::H + c] : 0.f) + (dout_lp ? __bfloat162float(dout_lp[t * enc::H + c]) : 0.f);
    }
    drop_row(dy, t, lane, dc);       // the same mask as the forward
    ln_row_bwd(x, dy, gamma, stats[2 * t], stats[2 * t + 1], lane, dx, acc[0], acc[1]);
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) dx_out[t * enc::H + lane + 32 * k] = dx[k];
  }
  ln_param_partials<2>(acc, part);
}

// dpos[l][c] = sum_b dx[b][l][c]  (fixed order over b); one thread per (l, c)
__global__ void enc_sum_over_batch_kernel(const float* __restrict__ dx, int batch, int seq_len,
                                          float* __restrict__ dpos) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)seq_len * enc::H) return;
  float s = 0.f;
  for (int b = 0; b < batch; ++b) s += dx[(int64_t)b * seq_len * enc::H + e];
  dpos[e] = s;
}
// dtype0[c] = sum_l dpos[l][c]
__global__ void enc_sum_rows_kernel(const float* __restrict__ x, int rows, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= enc::H) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += x[(int64_t)r * enc::H + c];
  out[c] = s;
}

// ---- LayerNorm(y + residual): BertSelfOutput / BertOutput ------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
enc_add_ln_fwd_kernel(const T* __restrict__ y, const float* __restrict__ bias, const float* __restrict__ res,
                      const float* __restrict__ gamma, const float* __restrict__ beta, int64_t n_tok, float eps,
                      float* __restrict__ out, __nv_bfloat16* __restrict__ out_lp, float* __restrict__ stats,
                      const int64_t* __restrict__ rng, float drop_p, int site) {
  const DropCtx dc = make_drop(rng, drop_p, site, 16);
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < n_tok; t += nwarps) {
    float x[enc::PER], o[enc::PER];
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) {
      const int c = lane + 32 * k;
      x[k] = to_f32(y[t * enc::H + c]) + (bias ? bias[c] : 0.f);
    }
    drop_row(x, t, lane, dc);        // BertSelfOutput / BertOutput: LayerNorm(dropout(dense(h)) + residual)
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) x[k] += res[t * enc::H + lane + 32 * k];
    float mean, rstd;
    ln_row(x, gamma, beta, eps, lane, o, mean, rstd);
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) out[t * enc::H + lane + 32 * k] = o[k];
    if (out_lp)
#pragma unroll
      for (int k = 0; k < enc::PER; ++k) out_lp[t * enc::H + lane + 32 * k] = __float2bfloat16_rn(o[k]);
    if (lane == 0) {
      stats[2 * t] = mean;
      stats[2 * t + 1] = rstd;
    }
  }
}

// dx = dLN (the gradient of BOTH y and the residual); partials of dgamma, dbeta and of the dense layer's bias
// gradient (the column sums of dx).  The upstream gradient is dout (fp32) + dout_lp (bf16), either may be null.
template <typename T>
__global__ void __launch_bounds__(256, 2)   // <= 128 registers: with the dropout recompute ptxas took 141 -> one block per SM
enc_add_ln_bwd_kernel(const T* __restrict__ y, const float* __restrict__ bias, const float* __restrict__ res,
                      const float* __restrict__ gamma, const float* __restrict__ stats,
                      const float* __restrict__ dout, const __nv_bfloat16* __restrict__ dout_lp, int64_t n_tok,
                      float* __restrict__ dx_out, T* __restrict__ dy_out, float* __restrict__ part,
                      const int64_t* __restrict__ rng, float drop_p, int site) {
  const DropCtx dc = make_drop(rng, drop_p, site, 16);
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float acc[3][enc::PER] = {};
  for (int64_t t = warp; t < n_tok; t += nwarps) {
    float x[enc::PER], dy[enc::PER], dx[enc::PER];
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) {
      const int c = lane + 32 * k;
      x[k] = to_f32(y[t * enc::H + c]) + (bias ? bias[c] : 0.f);
      dy[k] = (dout ? dout[t * enc::H + c] : 0.f) + (dout_lp ? __bfloat162float(dout_lp[t * enc::H + c]) : 0.f);
    }
    const uint32_t keep = drop_bits(t, lane, dc);
    drop_apply(x, keep, dc);         // rebuild the forward's LayerNorm input with the same mask
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) x[k] += res[t * enc::H + lane + 32 * k];
    ln_row_bwd(x, dy, gamma, stats[2 * t], stats[2 * t + 1], lane, dx, acc[0], acc[1]);
    float dyd[enc::PER];             // through the dropout: gradient of dense(h) + bias
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) dyd[k] = dx[k];
    drop_apply(dyd, keep, dc);
#pragma unroll
    for (int k = 0; k < enc::PER; ++k) {
      const int c = lane + 32 * k;
      dx_out[t * enc::H + c] = dx[k];                 // gradient of the residual stream (fp32)
      dy_out[t * enc::H + c] = from_f32<T>(dyd[k]);   // gradient of the linear layer's output (its dtype)
      acc[2][k] += dyd[k];
    }
  }
  ln_param_partials<3>(acc, part);
}

// ---- exact GELU (BertIntermediate, hidden_act = "gelu") ----------------------------------------------------
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float v) {
  const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * v * v);
  return cdf + v * pdf;
}
// bf16 activations: erf through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below a bf16 ulp): one
// exponential, one reciprocal and five FMAs instead of erff's ~30 instructions -- the GELU kernels are
// instruction-bound otherwise (39 M elements per launch) -- and the backward's exp(-x^2 / 2) IS that exponential.
__device__ __forceinline__ void gelu_fast_parts(float v, float& cdf, float& e) {
  const float z = fabsf(v) * 0.70710678118654752f;
  float t;   // MUFU reciprocal (2^-22 relative error: far below the polynomial's 1.5e-7 and a bf16 ulp)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  e = __expf(-z * z);                                    // = exp(-v^2 / 2)
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f),
                              0.254829592f);
  const float erf_abs = 1.0f - poly * e;
  cdf = 0.5f * (1.0f + copysignf(erf_abs, v));
}
template <typename T>
__device__ __forceinline__ float gelu_of(float v) {
  if (sizeof(T) == 2) {
    float cdf, e;
    gelu_fast_parts(v, cdf, e);
    return v * cdf;
  }
  return gelu_f(v);
}
template <typename T>
__device__ __forceinline__ float gelu_grad_of(float v) {
  if (sizeof(T) == 2) {
    float cdf, e;
    gelu_fast_parts(v, cdf, e);
    return fmaf(v * 0.3989422804014327f, e, cdf);
  }
  return gelu_grad_f(v);
}
// 16 bytes per thread and access (VEC = 16 / sizeof(T) elements); the scalar tail covers n % VEC and
// unaligned buffers (vec_ok = 0)
template <typename T>
__global__ void __launch_bounds__(256)
enc_gelu_fwd_kernel(const T* __restrict__ x, int64_t n, T* __restrict__ y, int vec_ok) {
  constexpr int VEC = 16 / sizeof(T);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t nv = vec_ok ? n / VEC : 0;
  for (int64_t i = tid; i < nv; i += nth) {
    int4 raw = ld_stream16(x + i * VEC);
    if (sizeof(T) == 2) {   // bf16: unpack / pack two elements per instruction
      uint32_t* w = reinterpret_cast<uint32_t*>(&raw);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float lo = __uint_as_float(w[k] << 16), hi = __uint_as_float(w[k] & 0xFFFF0000u);
        const __nv_bfloat162 r = __floats2bfloat162_rn(gelu_of<T>(lo), gelu_of<T>(hi));
        w[k] = *reinterpret_cast<const uint32_t*>(&r);
      }
    } else {
      T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
      for (int k = 0; k < VEC; ++k) e[k] = from_f32<T>(gelu_of<T>(to_f32(e[k])));
    }
    *reinterpret_cast<int4*>(y + i * VEC) = raw;
  }
  for (int64_t i = nv * VEC + tid; i < n; i += nth) y[i] = from_f32<T>(gelu_of<T>(to_f32(x[i])));
}
template <typename T>
__global__ void __launch_bounds__(256)
enc_gelu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, int64_t n, T* __restrict__ dx, int vec_ok) {
  constexpr int VEC = 16 / sizeof(T);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t nv = vec_ok ? n / VEC : 0;
  for (int64_t i = tid; i < nv; i += nth) {
    int4 rx = ld_stream16(x + i * VEC), rd = ld_stream16(dy + i * VEC);
    if (sizeof(T) == 2) {   // bf16: unpack / pack two elements per instruction
      const uint32_t* wx = reinterpret_cast<const uint32_t*>(&rx);
      uint32_t* wd = reinterpret_cast<uint32_t*>(&rd);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x0 = __uint_as_float(wx[k] << 16), x1 = __uint_as_float(wx[k] & 0xFFFF0000u);
        const float d0 = __uint_as_float(wd[k] << 16), d1 = __uint_as_float(wd[k] & 0xFFFF0000u);
        const __nv_bfloat162 r = __floats2bfloat162_rn(d0 * gelu_grad_of<T>(x0), d1 * gelu_grad_of<T>(x1));
        wd[k] = *reinterpret_cast<const uint32_t*>(&r);
      }
    } else {
      const T* ex = reinterpret_cast<const T*>(&rx);
      T* ed = reinterpret_cast<T*>(&rd);
#pragma unroll
      for (int k = 0; k < VEC; ++k) ed[k] = from_f32<T>(to_f32(ed[k]) * gelu_grad_of<T>(to_f32(ex[k])));
    }
    st_stream16(dx + i * VEC, rd);
  }
  for (int64_t i = nv * VEC + tid; i < n; i += nth) dx[i] = from_f32<T>(to_f32(dy[i]) * gelu_grad_of<T>(to_f32(x[i])));
}

// column sums of a (rows, width) matrix (a linear layer's bias gradient): block = 8 warps x 32 lanes, a lane owns
// VEC consecutive columns, warp w of row slice blockIdx.y takes rows y * 8 + w, + 8 * gridDim.y, ...; partial rows
// part[y][width] folded by enc_fold_partials_kernel.  Fixed order everywhere: deterministic.
constexpr int COLSUM_SLICES = 192;   // row slices = grid.y (64 left the kernel latency-bound: 384 blocks, 33 % warps active)
template <typename T>
__global__ void __launch_bounds__(256)
enc_colsum_kernel(const T* __restrict__ x, int64_t rows, int width, float* __restrict__ part) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float s[8][32 * VEC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + lane) * VEC;
  float a[VEC] = {};
  if (c0 < width)
    for (int64_t r = (int64_t)blockIdx.y * 8 + warp; r < rows; r += 8 * (int64_t)gridDim.y) {
      int4 raw = ld_stream16(x + r * width + c0);
      const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int k = 0; k < VEC; ++k) a[k] += to_f32(e[k]);
    }
#pragma unroll
  for (int k = 0; k < VEC; ++k) s[warp][lane * VEC + k] = a[k];
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * VEC; e += blockDim.x) {
    const int c = blockIdx.x * 32 * VEC + e;
    if (c >= width) continue;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s[w][e];
    part[(size_t)blockIdx.y * width + c] = t;
  }
}

// ---- causal multi-head self-attention, head_dim 32 ---------------------------------------------------------
// qkv: (B, L, 3 * H) rows = [Q | K | V], each H = n_heads * 32; keymask (B, L) bytes (models.py:343).  One
// block per (sequence, head): K and V of the head live in shared memory (rows padded to 33 floats: lane j
// reads row j), a warp owns one query at a time: lane = key for the scores, lane = dim for the output.
// Query i attends keys j <= i with keymask[j] (BertModel is_decoder=True: causal mask AND padding mask);
// scores are scaled by 1/sqrt(32).  lse (B, n_heads, L): log-sum-exp of every query row, for the backward.
template <typename T>
__global__ void __launch_bounds__(256)
enc_attn_fwd_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ keymask, int seq_len, int n_heads,
                    T* __restrict__ ctx, float* __restrict__ lse, const int64_t* __restrict__ rng, float drop_p,
                    int site) {
  const DropCtx dc = make_drop(rng, drop_p, site, 8);
  extern __shared__ float sm[];
  float* sK = sm;                                   // [L][33]
  float* sV = sm + (size_t)seq_len * 33;            // [L][33]
  __shared__ uint8_t s_mask[enc::MAX_L];
  const int b = blockIdx.x / n_heads, h = blockIdx.x % n_heads;
  const int hid = n_heads * enc::HD;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const T* base = qkv + (int64_t)b * seq_len * 3 * hid;
  for (int e = threadIdx.x; e < seq_len * enc::HD; e += blockDim.x) {
    const int j = e >> 5, d = e & 31;
    sK[j * 33 + d] = to_f32(base[(int64_t)j * 3 * hid + hid + h * enc::HD + d]);
    sV[j * 33 + d] = to_f32(base[(int64_t)j * 3 * hid + 2 * hid + h * enc::HD + d]);
  }
  for (int j = threadIdx.x; j < seq_len; j += blockDim.x) s_mask[j] = keymask[(int64_t)b * seq_len + j];
  __syncthreads();
  const float scale = 0.17677669529663687f;   // 1 / sqrt(32)
  for (int i = warp; i < seq_len; i += nw) {
    const float qd = to_f32(base[(int64_t)i * 3 * hid + h * enc::HD + lane]) * scale;   // lane = dim
    float m = -CUDART_INF_F, l = 0.f, o = 0.f;                                          // o: lane = dim
    for (int j0 = 0; j0 <= i; j0 += 32) {
      const int j = j0 + lane;
      const bool ok = j <= i && s_mask[j < seq_len ? j : 0];
      float s = 0.f;
      const float* kr = sK + (size_t)(j < seq_len ? j : 0) * 33;
#pragma unroll
      for (int d = 0; d < 32; ++d) s = fmaf(__shfl_sync(0xffffffffu, qd, d), kr[d], s);
      s = ok ? s : -CUDART_INF_F;
      const float mnew = fmaxf(m, warp_max(s));
      if (mnew == -CUDART_INF_F) continue;          // no valid key so far
      const float p = ok ? __expf(s - mnew) : 0.f;
      const float corr = __expf(m - mnew);          // exp(-inf) = 0 on the first valid chunk
      l = l * corr + warp_sum(p);
      o *= corr;
      const float pd = p * attn_drop_elem(dc, blockIdx.x, i, j);   // dropout on the probabilities, not on their sum
      const int jn = min(32, i + 1 - j0);
      for (int jj = 0; jj < jn; ++jj) o = fmaf(__shfl_sync(0xffffffffu, pd, jj), sV[(size_t)(j0 + jj) * 33 + lane], o);
      m = mnew;
    }
    const int64_t t = (int64_t)b * seq_len + i;
    ctx[t * hid + h * enc::HD + lane] = from_f32<T>(l > 0.f ? o / l : 0.f);
    if (lane == 0) lse[((int64_t)b * n_heads + h) * seq_len + i] = l > 0.f ? m + __logf(l) : CUDART_INF_F;
  }
}

// backward, pass A (warp per query): delta_i = dO_i . O_i, dQ_i = scale * sum_j dS_ij K_j with
// dS_ij = P_ij (dO_i . V_j - delta_i), P_ij = exp(scale q_i . k_j - lse_i)
// pass B (warp per key): dK_j = scale * sum_i dS_ij Q_i, dV_j = sum_i P_ij dO_i.  Both recompute P; no atomics.
template <typename T>
__global__ void __launch_bounds__(256)
enc_attn_bwd_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ keymask, const T* __restrict__ ctx,
                    const T* __restrict__ dctx, const float* __restrict__ lse, int seq_len, int n_heads,
                    T* __restrict__ dqkv, const int64_t* __restrict__ rng, float drop_p, int site) {
  const DropCtx dc = make_drop(rng, drop_p, site, 8);
  extern __shared__ float sm[];
  float* sQ = sm;                                   // [L][33] scaled queries
  float* sK = sQ + (size_t)seq_len * 33;
  float* sV = sK + (size_t)seq_len * 33;
  float* sdO = sV + (size_t)seq_len * 33;
  float* s_lse = sdO + (size_t)seq_len * 33;        // [L]
  float* s_delta = s_lse + seq_len;                 // [L]
  __shared__ uint8_t s_mask[enc::MAX_L];
  const int b = blockIdx.x / n_heads, h = blockIdx.x % n_heads;
  const int hid = n_heads * enc::HD;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float scale = 0.17677669529663687f;
  const T* base = qkv + (int64_t)b * seq_len * 3 * hid;
  T* dbase = dqkv + (int64_t)b * seq_len * 3 * hid;
  for (int e = threadIdx.x; e < seq_len * enc::HD; e += blockDim.x) {
    const int j = e >> 5, d = e & 31;
    const int64_t t = (int64_t)b * seq_len + j;
    sQ[j * 33 + d] = to_f32(base[(int64_t)j * 3 * hid + h * enc::HD + d]) * scale;
    sK[j * 33 + d] = to_f32(base[(int64_t)j * 3 * hid + hid + h * enc::HD + d]);
    sV[j * 33 + d] = to_f32(base[(int64_t)j * 3 * hid + 2 * hid + h * enc::HD + d]);
    sdO[j * 33 + d] = to_f32(dctx[t * hid + h * enc::HD + d]);
  }
  for (int j = threadIdx.x; j < seq_len; j += blockDim.x) {
    s_mask[j] = keymask[(int64_t)b * seq_len + j];
    s_lse[j] = lse[((int64_t)b * n_heads + h) * seq_len + j];
  }
  __syncthreads();
  // delta_i = dO_i . O_i  (warp per query, lane = dim)
  for (int i = warp; i < seq_len; i += nw) {
    const int64_t t = (int64_t)b * seq_len + i;
    const float v = sdO[i * 33 + lane] * to_f32(ctx[t * hid + h * enc::HD + lane]);
    const float d = warp_sum(v);
    if (lane == 0) s_delta[i] = d;
  }
  __syncthreads();
  // pass A: dQ
  for (int i = warp; i < seq_len; i += nw) {
    const float lse_i = s_lse[i], delta_i = s_delta[i];
    float dq = 0.f;                                                   // lane = dim
    if (lse_i != CUDART_INF_F) {
      for (int j0 = 0; j0 <= i; j0 += 32) {
        const int j = j0 + lane;
        const int jc = j < seq_len ? j : 0;
        const bool ok = j <= i && s_mask[jc];
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) {
          s = fmaf(sQ[i * 33 + d], sK[jc * 33 + d], s);
          dp = fmaf(sdO[i * 33 + d], sV[jc * 33 + d], dp);
        }
        const float ds = ok ? __expf(s - lse_i) * (dp * attn_drop_elem(dc, blockIdx.x, i, j) - delta_i) : 0.f;   // lane = key
        const int jn = min(32, i + 1 - j0);
        for (int jj = 0; jj < jn; ++jj) dq = fmaf(__shfl_sync(0xffffffffu, ds, jj), sK[(size_t)(j0 + jj) * 33 + lane], dq);
      }
    }
    dbase[(int64_t)i * 3 * hid + h * enc::HD + lane] = from_f32<T>(dq * scale);
  }
  // pass B: dK, dV (warp per key j; queries i >= j in chunks of 32, lane = query)
  for (int j = warp; j < seq_len; j += nw) {
    float dk = 0.f, dv = 0.f;                                         // lane = dim
    if (s_mask[j]) {
      for (int i0 = j; i0 < seq_len; i0 += 32) {
        const int i = i0 + lane;
        const int ic = i < seq_len ? i : 0;
        const bool ok = i < seq_len && s_lse[ic] != CUDART_INF_F;
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) {
          s = fmaf(sQ[ic * 33 + d], sK[j * 33 + d], s);
          dp = fmaf(sdO[ic * 33 + d], sV[j * 33 + d], dp);
        }
        const float p = ok ? __expf(s - s_lse[ic]) : 0.f;             // lane = query
        const float f = attn_drop_elem(dc, blockIdx.x, ic, j);
        const float ds = p * (dp * f - s_delta[ic]);
        const float pdrop = p * f;
        const int in = min(32, seq_len - i0);
        for (int ii = 0; ii < in; ++ii) {
          const float pb = __shfl_sync(0xffffffffu, pdrop, ii), dsb = __shfl_sync(0xffffffffu, ds, ii);
          dv = fmaf(pb, sdO[(size_t)(i0 + ii) * 33 + lane], dv);
          dk = fmaf(dsb, sQ[(size_t)(i0 + ii) * 33 + lane], dk);      // sQ already carries the 1/sqrt(d) scale
        }
      }
    }
    dbase[(int64_t)j * 3 * hid + hid + h * enc::HD + lane] = from_f32<T>(dk);
    dbase[(int64_t)j * 3 * hid + 2 * hid + h * enc::HD + lane] = from_f32<T>(dv);
  }
}

// ---- bf16 attention on the tensor cores ----------------------------------------------------------------------
// Same arithmetic as the kernels above for bf16 activations, with the four contractions of a head
// (Q K^T, P V; dO V^T, dS K, P^T dO, dS^T Q) issued as warp-level m16n8k16 bf16 MMAs with fp32 accumulators.
// A head is 32 wide: Q K^T is two k-steps, so a head never fills a 128-row tcgen05 tile worth its TMEM
// round trip for the softmax; the whole layer is 7.8 GFLOP forward and the kernel is bound by issuing
// exp / max / pack around the MMAs.  One block per (sequence, head); a warp owns 16 queries (forward, dQ pass)
// or 16 keys (dK / dV pass).  The probabilities feed the second MMA straight from the accumulator registers
// (the C fragment of two adjacent n8 tiles IS the A fragment of the next k16 step).  Deterministic: no atomics.
namespace enc {
constexpr int RS = 40;   // row stride (bf16) of the [L][32] tiles: 80 bytes -> the eight 16-byte rows of an 8x8
                         // ldmatrix block fall in eight different 16-byte bank groups
__host__ __device__ constexpr int pad16(int l) { return (l + 15) & ~15; }
}  // namespace enc

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// A fragment (16 rows x 32 dims = two k-steps) of a row-major tile whose first row is `rows`
__device__ __forceinline__ void load_a_rows(const __nv_bfloat16* rows, int g, int t, uint32_t (&a)[2][4]) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    a[ks][0] = lds32(rows + g * enc::RS + 16 * ks + 2 * t);
    a[ks][1] = lds32(rows + (g + 8) * enc::RS + 16 * ks + 2 * t);
    a[ks][2] = lds32(rows + g * enc::RS + 16 * ks + 2 * t + 8);
    a[ks][3] = lds32(rows + (g + 8) * enc::RS + 16 * ks + 2 * t + 8);
  }
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const __nv_bfloat16* row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const __nv_bfloat16* row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// c[nt] (16 x 8) += A (16 x 32) . X[n0 + 8 nt + ..][0..32]^T for nt = 0, 1: X row-major, its rows are the n index.
// One ldmatrix.x4 per n tile: the four 8x8 blocks (rows n0 + 8 nt .., columns 0-7 / 8-15 / 16-23 / 24-31) are
// the B fragments {b0, b1} of the two k-steps.
__device__ __forceinline__ void mma_rows(float (&c)[2][4], const uint32_t (&a)[2][4], const __nv_bfloat16* x, int n0,
                                         int lane) {
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    uint32_t b[4];
    ldmatrix_x4(b, x + (n0 + 8 * nt + (lane & 7)) * enc::RS + 8 * (lane >> 3));
    mma16816(c[nt], a[0], b[0], b[1]);
    mma16816(c[nt], a[1], b[2], b[3]);
  }
}
// acc[dt] (16 x 8 columns 8 dt ..) += A (16 x 16, k = rows k0 .. k0+15 of X) . X[k0 ..][8 dt ..]: X row-major with
// the k index as its rows -- the transposing ldmatrix hands out the {k, k+1} pairs the B fragment wants, so no
// transposed copy of X is kept.  One ldmatrix.x4.trans serves two column tiles.
__device__ __forceinline__ void mma_cols(float (&acc)[4][4], const uint32_t (&a)[4], const __nv_bfloat16* x, int k0,
                                         int lane) {
  const int m = lane >> 3;
#pragma unroll
  for (int pr = 0; pr < 2; ++pr) {
    uint32_t b[4];
    ldmatrix_x4_trans(b, x + (k0 + (lane & 7) + 8 * (m & 1)) * enc::RS + 16 * pr + 8 * (m >> 1));
    mma16816(acc[2 * pr], a, b[0], b[1]);
    mma16816(acc[2 * pr + 1], a, b[2], b[3]);
  }
}
// tile order of a block's warps: heaviest tiles first, alternating direction so every warp gets a similar sum
__device__ __forceinline__ int tile_of(int round, int warp, int nw, int n_tiles, bool heavy_last) {
  const int k = (round & 1) ? round * nw + (nw - 1 - warp) : round * nw + warp;
  if (k >= n_tiles) return -1;
  return heavy_last ? n_tiles - 1 - k : k;
}

// one head's [L][32] slice of a (rows, ld) bf16 matrix -> row-major tile; rows L .. lp-1 are zero so that masked
// probabilities never meet a NaN inside an MMA
__device__ __forceinline__ void fill_tile(const __nv_bfloat16* __restrict__ src, int64_t ld, int seq_len, int lp,
                                          __nv_bfloat16* rowmajor) {
  for (int e = threadIdx.x; e < lp * 4; e += blockDim.x) {
    const int j = e >> 2, c = e & 3;
    int4 v = make_int4(0, 0, 0, 0);
    if (j < seq_len) v = ld_stream16(src + (int64_t)j * ld + 8 * c);
    *reinterpret_cast<int4*>(rowmajor + j * enc::RS + 8 * c) = v;
  }
}

__global__ void __launch_bounds__(256)
enc_attn_fwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const uint8_t* __restrict__ keymask, int seq_len,
                        int n_heads, __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse,
                        const int64_t* __restrict__ rng, float drop_p, int site) {
  const DropCtx dc = make_drop(rng, drop_p, site, 8);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lp = enc::pad16(seq_len);
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [lp][RS]
  __nv_bfloat16* sK = sQ + lp * enc::RS;                            // [lp][RS]
  __nv_bfloat16* sV = sK + lp * enc::RS;                            // [lp][RS]
  uint8_t* s_mask = reinterpret_cast<uint8_t*>(sV + lp * enc::RS);  // [lp]
  const int b = blockIdx.x / n_heads, h = blockIdx.x % n_heads;
  const int hid = n_heads * enc::HD;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* base = qkv + (int64_t)b * seq_len * 3 * hid + h * enc::HD;
  // one pass over the head's rows with all three loads of a row segment in flight (three separate fill loops
  // left the prologue latency-bound: ncu showed the kernel stalled on global loads, not on its MMAs)
#pragma unroll 2
  for (int e = threadIdx.x; e < lp * 4; e += blockDim.x) {
    const int j = e >> 2, c = e & 3;
    int4 vq = make_int4(0, 0, 0, 0), vk = vq, vv = vq;
    if (j < seq_len) {
      const __nv_bfloat16* src = base + (int64_t)j * 3 * hid + 8 * c;
      vq = ld_stream16(src);
      vk = ld_stream16(src + hid);
      vv = ld_stream16(src + 2 * hid);
    }
    *reinterpret_cast<int4*>(sQ + j * enc::RS + 8 * c) = vq;
    *reinterpret_cast<int4*>(sK + j * enc::RS + 8 * c) = vk;
    *reinterpret_cast<int4*>(sV + j * enc::RS + 8 * c) = vv;
  }
  for (int j = threadIdx.x; j < lp; j += blockDim.x) s_mask[j] = j < seq_len ? keymask[(int64_t)b * seq_len + j] : 0;
  __syncthreads();
  const float scale = 0.17677669529663687f;   // 1 / sqrt(32)
  const int n_tiles = lp >> 4;
  for (int round = 0;; ++round) {
    const int tile = tile_of(round, warp, nw, n_tiles, true);
    if (tile < 0) break;
    const int i0 = tile << 4;
    uint32_t qa[2][4];
    load_a_rows(sQ + i0 * enc::RS, g, t, qa);
    float m[2] = {-CUDART_INF_F, -CUDART_INF_F}, l[2] = {0.f, 0.f};
    float o[4][4] = {};
    for (int kb = 0; kb <= i0; kb += 16) {
      float s[2][4] = {};
      mma_rows(s, qa, sK, kb, lane);
      float mx[2] = {-CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kb + 8 * nt + 2 * t + (e & 1), qi = i0 + g + 8 * (e >> 1);
          const bool ok = key <= qi && s_mask[key];
          s[nt][e] = ok ? s[nt][e] * scale : -CUDART_INF_F;
          mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
        }
      float corr[2], mnew[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        mnew[r] = fmaxf(m[r], mx[r]);
        corr[r] = mnew[r] == -CUDART_INF_F ? 1.f : __expf(m[r] - mnew[r]);
        m[r] = mnew[r];
      }
      float p[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          p[nt][e] = mnew[r] == -CUDART_INF_F ? 0.f : __expf(s[nt][e] - mnew[r]);
        }
#pragma unroll
      for (int r = 0; r < 2; ++r)
        l[r] = l[r] * corr[r] + p[0][2 * r] + p[0][2 * r + 1] + p[1][2 * r] + p[1][2 * r + 1];
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        o[dt][0] *= corr[0]; o[dt][1] *= corr[0];
        o[dt][2] *= corr[1]; o[dt][3] *= corr[1];
      }
      if (dc.on) {   // dropout on the probabilities that reach P V; the softmax sum l keeps all of them
        const uint64_t half =
            attn_row_half(attn_drop_call(dc, blockIdx.x, (uint32_t)tile, (uint32_t)(kb >> 4), (g >> 1) * 4 + t), g);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            p[nt][e] = attn_keep8(half, (e >> 1) * 4 + (e & 1) * 2 + nt, dc) ? p[nt][e] * dc.scale : 0.f;
      }
      const uint32_t pa[4] = {pack_bf16x2(p[0][0], p[0][1]), pack_bf16x2(p[0][2], p[0][3]),
                              pack_bf16x2(p[1][0], p[1][1]), pack_bf16x2(p[1][2], p[1][3])};
      mma_cols(o, pa, sV, kb, lane);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
      const int qi = i0 + g + 8 * r;
      if (qi >= seq_len) continue;
      const float inv = l[r] > 0.f ? 1.f / l[r] : 0.f;
      const int64_t tok = (int64_t)b * seq_len + qi;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt)
        *reinterpret_cast<uint32_t*>(ctx + tok * hid + h * enc::HD + 8 * dt + 2 * t) =
            pack_bf16x2(o[dt][2 * r] * inv, o[dt][2 * r + 1] * inv);
      if (t == 0) lse[((int64_t)b * n_heads + h) * seq_len + qi] = l[r] > 0.f ? m[r] + __logf(l[r]) : CUDART_INF_F;
    }
  }
}

__global__ void __launch_bounds__(256)
enc_attn_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const uint8_t* __restrict__ keymask,
                        const __nv_bfloat16* __restrict__ ctx, const __nv_bfloat16* __restrict__ dctx,
                        const float* __restrict__ lse, int seq_len, int n_heads, __nv_bfloat16* __restrict__ dqkv,
                        const int64_t* __restrict__ rng, float drop_p, int site) {
  const DropCtx dc = make_drop(rng, drop_p, site, 8);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lp = enc::pad16(seq_len);
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // four [lp][RS] tiles
  __nv_bfloat16* sK = sQ + lp * enc::RS;
  __nv_bfloat16* sV = sK + lp * enc::RS;
  __nv_bfloat16* sdO = sV + lp * enc::RS;
  float* s_lse = reinterpret_cast<float*>(sdO + lp * enc::RS);      // [lp]
  float* s_delta = s_lse + lp;                               // [lp]
  uint8_t* s_mask = reinterpret_cast<uint8_t*>(s_delta + lp);
  const int b = blockIdx.x / n_heads, h = blockIdx.x % n_heads;
  const int hid = n_heads * enc::HD;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float scale = 0.17677669529663687f;
  const __nv_bfloat16* base = qkv + (int64_t)b * seq_len * 3 * hid + h * enc::HD;
  __nv_bfloat16* dbase = dqkv + (int64_t)b * seq_len * 3 * hid + h * enc::HD;
  const __nv_bfloat16* dO = dctx + (int64_t)b * seq_len * hid + h * enc::HD;
  const __nv_bfloat16* O = ctx + (int64_t)b * seq_len * hid + h * enc::HD;
  // one pass over the head's rows with all five loads of a row segment in flight; delta_i = dO_i . O_i falls
  // out of the same pass (four lanes hold a row: two shuffles).  Separate fill loops and a warp-per-row delta
  // loop left the prologue latency-bound: ncu showed 4.1 long-scoreboard stalls per issued instruction.
  // (lp * 4 is a multiple of 64: whole warps enter or skip an iteration, the shuffles are safe)
#pragma unroll 2
  for (int e = threadIdx.x; e < lp * 4; e += blockDim.x) {
    const int j = e >> 2, c = e & 3;
    int4 vq = make_int4(0, 0, 0, 0), vk = vq, vv = vq, vd = vq, vo = vq;
    if (j < seq_len) {
      const __nv_bfloat16* src = base + (int64_t)j * 3 * hid + 8 * c;
      vq = ld_stream16(src);
      vk = ld_stream16(src + hid);
      vv = ld_stream16(src + 2 * hid);
      vd = ld_stream16(dO + (int64_t)j * hid + 8 * c);
      vo = ld_stream16(O + (int64_t)j * hid + 8 * c);
    }
    *reinterpret_cast<int4*>(sQ + j * enc::RS + 8 * c) = vq;
    *reinterpret_cast<int4*>(sK + j * enc::RS + 8 * c) = vk;
    *reinterpret_cast<int4*>(sV + j * enc::RS + 8 * c) = vv;
    *reinterpret_cast<int4*>(sdO + j * enc::RS + 8 * c) = vd;
    const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&vd);
    const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&vo);
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = __bfloat1622float2(hd[k]), o2 = __bfloat1622float2(ho[k]);
      part = fmaf(a.x, o2.x, part);
      part = fmaf(a.y, o2.y, part);
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (c == 0) s_delta[j] = part;
  }
  for (int j = threadIdx.x; j < lp; j += blockDim.x) {
    s_mask[j] = j < seq_len ? keymask[(int64_t)b * seq_len + j] : 0;
    s_lse[j] = j < seq_len ? lse[((int64_t)b * n_heads + h) * seq_len + j] : CUDART_INF_F;
  }
  __syncthreads();
  const int n_tiles = lp >> 4;
  // pass A: dQ of 16 queries per warp
  for (int round = 0;; ++round) {
    const int tile = tile_of(round, warp, nw, n_tiles, true);
    if (tile < 0) break;
    const int i0 = tile << 4;
    uint32_t qa[2][4], doa[2][4];
    load_a_rows(sQ + i0 * enc::RS, g, t, qa);
    load_a_rows(sdO + i0 * enc::RS, g, t, doa);
    const float lse_r[2] = {s_lse[i0 + g], s_lse[i0 + g + 8]}, delta_r[2] = {s_delta[i0 + g], s_delta[i0 + g + 8]};
    float dq[4][4] = {};
    for (int kb = 0; kb <= i0; kb += 16) {
      float s[2][4] = {}, dp[2][4] = {};
      mma_rows(s, qa, sK, kb, lane);
      mma_rows(dp, doa, sV, kb, lane);
      uint64_t half = 0;
      if (dc.on)
        half = attn_row_half(attn_drop_call(dc, blockIdx.x, (uint32_t)tile, (uint32_t)(kb >> 4), (g >> 1) * 4 + t), g);
      float ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          const int key = kb + 8 * nt + 2 * t + (e & 1), qi = i0 + g + 8 * r;
          const bool ok = key <= qi && s_mask[key] && lse_r[r] != CUDART_INF_F;
          float dpe = dp[nt][e];       // dL/dP through the dropout: the forward's mask, recomputed
          if (dc.on) dpe = attn_keep8(half, (e >> 1) * 4 + (e & 1) * 2 + nt, dc) ? dpe * dc.scale : 0.f;
          ds[nt][e] = ok ? __expf(s[nt][e] * scale - lse_r[r]) * (dpe - delta_r[r]) : 0.f;
        }
      const uint32_t dsa[4] = {pack_bf16x2(ds[0][0], ds[0][1]), pack_bf16x2(ds[0][2], ds[0][3]),
                               pack_bf16x2(ds[1][0], ds[1][1]), pack_bf16x2(ds[1][2], ds[1][3])};
      mma_cols(dq, dsa, sK, kb, lane);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = i0 + g + 8 * r;
      if (qi >= seq_len) continue;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt)
        *reinterpret_cast<uint32_t*>(dbase + (int64_t)qi * 3 * hid + 8 * dt + 2 * t) =
            pack_bf16x2(dq[dt][2 * r] * scale, dq[dt][2 * r + 1] * scale);
    }
  }
  // pass B: dK, dV of 16 keys per warp (transposed problem: rows = keys, columns = queries)
  for (int round = 0;; ++round) {
    const int tile = tile_of(round, warp, nw, n_tiles, false);
    if (tile < 0) break;
    const int j0 = tile << 4;
    uint32_t ka[2][4], va[2][4];
    load_a_rows(sK + j0 * enc::RS, g, t, ka);
    load_a_rows(sV + j0 * enc::RS, g, t, va);
    const bool key_ok[2] = {s_mask[j0 + g] != 0, s_mask[j0 + g + 8] != 0};
    float dk[4][4] = {}, dv[4][4] = {};
    for (int qb = j0; qb < lp; qb += 16) {
      float st[2][4] = {}, dpt[2][4] = {};
      mma_rows(st, ka, sQ, qb, lane);
      mma_rows(dpt, va, sdO, qb, lane);
      // transposed fragment: this lane holds (query 8 nt + 2 t + (e & 1), key g + 8 (e >> 1)) of the block, all in
      // Philox call t * 4 + g / 2 (see attn_drop_call)
      DropBlock db{};
      if (dc.on) db = attn_drop_call(dc, blockIdx.x, (uint32_t)(qb >> 4), (uint32_t)tile, t * 4 + (g >> 1));
      float p[2][4], ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          const int qi = qb + 8 * nt + 2 * t + (e & 1), key = j0 + g + 8 * r;
          const float lq = s_lse[qi];
          const bool ok = key <= qi && key_ok[r] && lq != CUDART_INF_F;
          const float pn = ok ? __expf(st[nt][e] * scale - lq) : 0.f;
          float f = 1.f;
          if (dc.on)   // byte (e & 1) * 8 + nt * 4 + (g & 1) * 2 + (e >> 1): word index static, shift per lane
            f = ((db.w[(e & 1) * 2 + nt] >> (8 * ((g & 1) * 2 + (e >> 1)))) & 0xFFu) >= dc.thr ? dc.scale : 0.f;
          ds[nt][e] = pn * (dpt[nt][e] * f - s_delta[qi]);
          p[nt][e] = pn * f;          // the dropped probabilities: what multiplied V in the forward
        }
      const uint32_t pa[4] = {pack_bf16x2(p[0][0], p[0][1]), pack_bf16x2(p[0][2], p[0][3]),
                              pack_bf16x2(p[1][0], p[1][1]), pack_bf16x2(p[1][2], p[1][3])};
      const uint32_t dsa[4] = {pack_bf16x2(ds[0][0], ds[0][1]), pack_bf16x2(ds[0][2], ds[0][3]),
                               pack_bf16x2(ds[1][0], ds[1][1]), pack_bf16x2(ds[1][2], ds[1][3])};
      mma_cols(dv, pa, sdO, qb, lane);
      mma_cols(dk, dsa, sQ, qb, lane);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = j0 + g + 8 * r;
      if (key >= seq_len) continue;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        *reinterpret_cast<uint32_t*>(dbase + (int64_t)key * 3 * hid + hid + 8 * dt + 2 * t) =
            pack_bf16x2(dk[dt][2 * r] * scale, dk[dt][2 * r + 1] * scale);
        *reinterpret_cast<uint32_t*>(dbase + (int64_t)key * 3 * hid + 2 * hid + 8 * dt + 2 * t) =
            pack_bf16x2(dv[dt][2 * r], dv[dt][2 * r + 1]);
      }
    }
  }
}

static size_t attn_mma_smem(int seq_len, bool bwd) {
  const size_t lp = enc::pad16(seq_len);
  if (!bwd) return 3 * lp * enc::RS * 2 + lp;
  return 4 * lp * enc::RS * 2 + 2 * lp * 4 + lp;
}

static inline int enc_grid(int64_t warps_needed) {
  int64_t blocks = (warps_needed + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace xr

using namespace xr;

extern "C" size_t xr_enc_ln_workspace_bytes(int64_t n_tok) {
  // partial rows of the parameter-gradient reductions + one (n_tok, H) fp32 scratch for the embedding backward
  return (size_t)LNB_BLOCKS * 3 * enc::H * 4 + (size_t)(n_tok > 0 ? n_tok : 1) * enc::H * 4 + 512;
}

extern "C" int xr_enc_embed_ln_fwd(const float* table, int64_t n_table_rows, const int64_t* idx,
                                   const float* pos_emb, const float* type_emb, const float* gamma,
                                   const float* beta, int64_t batch, int64_t seq_len, int64_t dim, float eps,
                                   float* out, void* out_bf16, float* stats, uint8_t* mask, int32_t* err_flag,
                                   const int64_t* rng, float drop_p, int site, void* stream) {
  XR_CHECK_ARG(table && idx && pos_emb && type_emb && gamma && beta && out && stats && mask,
               "xr_enc_embed_ln_fwd: null pointer");
  XR_CHECK_ARG(dim == enc::H, "xr_enc_embed_ln_fwd: this build is specialised for hidden size %d", enc::H);
  XR_CHECK_ARG(batch >= 0 && seq_len >= 1 && n_table_rows > 0, "xr_enc_embed_ln_fwd: bad sizes");
  const int64_t n_tok = batch * seq_len;
  if (n_tok == 0) return XR_OK;
  enc_embed_ln_fwd_kernel<<<enc_grid(n_tok), 256, 0, as_stream(stream)>>>(
      table, n_table_rows, idx, pos_emb, type_emb, gamma, beta, n_tok, (int)seq_len, eps, out, (__nv_bfloat16*)out_bf16,
      stats, mask, err_flag, rng, drop_p, site);
  XR_LAUNCH_CHECK("enc_embed_ln_fwd");
  return XR_OK;
}

extern "C" int xr_enc_embed_ln_bwd(const float* table, int64_t n_table_rows, const int64_t* idx,
                                   const float* pos_emb, const float* type_emb, const float* gamma,
                                   const float* stats, const float* dout, const void* dout_bf16, int64_t batch,
                                   int64_t seq_len, int64_t dim, float* dpos, float* dtype0, float* dgamma,
                                   float* dbeta, void* workspace, const int64_t* rng, float drop_p, int site,
                                   void* stream) {
  XR_CHECK_ARG(table && idx && pos_emb && type_emb && gamma && stats && (dout || dout_bf16) && dpos && dtype0 &&
                   dgamma && dbeta && workspace,
               "xr_enc_embed_ln_bwd: null pointer");
  XR_CHECK_ARG(dim == enc::H && batch >= 1 && seq_len >= 1, "xr_enc_embed_ln_bwd: bad sizes");
  cudaStream_t s = as_stream(stream);
  const int64_t n_tok = batch * seq_len;
  float* part = (float*)workspace;
  float* dx = part + (size_t)LNB_BLOCKS * 3 * enc::H;
  enc_embed_ln_bwd_kernel<<<LNB_BLOCKS, 256, 0, s>>>(table, n_table_rows, idx, pos_emb, type_emb, gamma, stats, dout,
                                                     (const __nv_bfloat16*)dout_bf16, n_tok, (int)seq_len, dx, part,
                                                     rng, drop_p, site);
  XR_LAUNCH_CHECK("enc_embed_ln_bwd");
  launch_fold(part, LNB_BLOCKS, 2, enc::H, dgamma, dbeta, nullptr, s);
  XR_LAUNCH_CHECK("enc_fold_partials");
  const int64_t n = seq_len * enc::H;
  enc_sum_over_batch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dx, (int)batch, (int)seq_len, dpos);
  XR_LAUNCH_CHECK("enc_sum_over_batch");
  enc_sum_rows_kernel<<<(enc::H + 127) / 128, 128, 0, s>>>(dpos, (int)seq_len, dtype0);
  XR_LAUNCH_CHECK("enc_sum_rows");
  return XR_OK;
}

extern "C" int xr_enc_add_ln_fwd(const void* y, int y_dtype, const float* bias, const float* residual,
                                 const float* gamma, const float* beta, int64_t n_tok, int64_t dim, float eps,
                                 float* out, void* out_bf16, float* stats, const int64_t* rng, float drop_p,
                                 int site, void* stream) {
  XR_CHECK_ARG(y && residual && gamma && beta && out && stats, "xr_enc_add_ln_fwd: null pointer");
  XR_CHECK_ARG(dim == enc::H && n_tok >= 0, "xr_enc_add_ln_fwd: bad sizes");
  if (n_tok == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  if (y_dtype == XR_F32)
    enc_add_ln_fwd_kernel<float><<<enc_grid(n_tok), 256, 0, s>>>((const float*)y, bias, residual, gamma, beta, n_tok, eps,
                                                                 out, (__nv_bfloat16*)out_bf16, stats, rng, drop_p, site);
  else if (y_dtype == XR_BF16)
    enc_add_ln_fwd_kernel<__nv_bfloat16><<<enc_grid(n_tok), 256, 0, s>>>(
        (const __nv_bfloat16*)y, bias, residual, gamma, beta, n_tok, eps, out, (__nv_bfloat16*)out_bf16, stats, rng,
        drop_p, site);
  else
    XR_CHECK_ARG(false, "xr_enc_add_ln_fwd: bad dtype");
  XR_LAUNCH_CHECK("enc_add_ln_fwd");
  return XR_OK;
}

extern "C" int xr_enc_add_ln_bwd(const void* y, int y_dtype, const float* bias, const float* residual,
                                 const float* gamma, const float* stats, const float* dout, const void* dout_bf16,
                                 int64_t n_tok, int64_t dim, float* dresidual, void* dy, float* dbias, float* dgamma,
                                 float* dbeta, void* workspace, const int64_t* rng, float drop_p, int site,
                                 void* stream) {
  XR_CHECK_ARG(y && residual && gamma && stats && (dout || dout_bf16) && dresidual && dy && dgamma && dbeta && workspace,
               "xr_enc_add_ln_bwd: null pointer");
  XR_CHECK_ARG(dim == enc::H && n_tok >= 1, "xr_enc_add_ln_bwd: bad sizes");
  cudaStream_t s = as_stream(stream);
  float* part = (float*)workspace;
  if (y_dtype == XR_F32)
    enc_add_ln_bwd_kernel<float><<<LNB_BLOCKS, 256, 0, s>>>((const float*)y, bias, residual, gamma, stats, dout,
                                                            (const __nv_bfloat16*)dout_bf16, n_tok, dresidual,
                                                            (float*)dy, part, rng, drop_p, site);
  else if (y_dtype == XR_BF16)
    enc_add_ln_bwd_kernel<__nv_bfloat16><<<LNB_BLOCKS, 256, 0, s>>>(
        (const __nv_bfloat16*)y, bias, residual, gamma, stats, dout, (const __nv_bfloat16*)dout_bf16, n_tok, dresidual,
        (__nv_bfloat16*)dy, part, rng, drop_p, site);
  else
    XR_CHECK_ARG(false, "xr_enc_add_ln_bwd: bad dtype");
  XR_LAUNCH_CHECK("enc_add_ln_bwd");
  launch_fold(part, LNB_BLOCKS, 3, enc::H, dgamma, dbeta, dbias, s);
  XR_LAUNCH_CHECK("enc_fold_partials");
  return XR_OK;
}

extern "C" int xr_enc_gelu(const void* x, const void* dy, int64_t n, int dtype, void* out, void* stream) {
  XR_CHECK_ARG(x && out && n >= 0, "xr_enc_gelu: bad arguments");
  if (n == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  int64_t blocks = (n / 4 + 255) / 256 + 1;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const int vec_ok = (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)out) % 16) == 0;
  if (dtype == XR_F32) {
    if (dy) enc_gelu_bwd_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)x, (const float*)dy, n, (float*)out, vec_ok);
    else enc_gelu_fwd_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)x, n, (float*)out, vec_ok);
  } else if (dtype == XR_BF16) {
    using bf = __nv_bfloat16;
    if (dy) enc_gelu_bwd_kernel<bf><<<(unsigned)blocks, 256, 0, s>>>((const bf*)x, (const bf*)dy, n, (bf*)out, vec_ok);
    else enc_gelu_fwd_kernel<bf><<<(unsigned)blocks, 256, 0, s>>>((const bf*)x, n, (bf*)out, vec_ok);
  } else {
    XR_CHECK_ARG(false, "xr_enc_gelu: bad dtype");
  }
  XR_LAUNCH_CHECK("enc_gelu");
  return XR_OK;
}

extern "C" size_t xr_enc_colsum_workspace_bytes(int64_t width) {
  return (size_t)COLSUM_SLICES * (size_t)(width > 0 ? width : 1) * 4 + 256;
}

extern "C" int xr_enc_colsum(const void* x, int dtype, int64_t rows, int64_t width, float* out, void* workspace,
                             void* stream) {
  XR_CHECK_ARG(x && out && workspace, "xr_enc_colsum: null pointer");
  XR_CHECK_ARG(rows >= 0 && width >= 1 && width <= (1 << 20), "xr_enc_colsum: bad sizes");
  const int vec = dtype == XR_BF16 ? 8 : 4;
  XR_CHECK_ARG(dtype == XR_BF16 || dtype == XR_F32, "xr_enc_colsum: bad dtype");
  XR_CHECK_ARG(width % vec == 0 && (uintptr_t)x % 16 == 0, "xr_enc_colsum: rows must be whole 16-byte vectors");
  cudaStream_t s = as_stream(stream);
  float* part = (float*)workspace;
  const dim3 grid((unsigned)((width / vec + 31) / 32), COLSUM_SLICES);
  if (dtype == XR_BF16) enc_colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, rows, (int)width, part);
  else enc_colsum_kernel<float><<<grid, 256, 0, s>>>((const float*)x, rows, (int)width, part);
  XR_LAUNCH_CHECK("enc_colsum");
  launch_fold(part, COLSUM_SLICES, 1, (int)width, out, nullptr, nullptr, s);
  XR_LAUNCH_CHECK("enc_fold_partials");
  return XR_OK;
}

static int enc_attention_launch_mma(const void* qkv, const uint8_t* keymask, const void* ctx, const void* dctx,
                                    float* lse, int64_t batch, int seq_len, int n_heads, void* out, cudaStream_t s,
                                    const int64_t* rng, float drop_p, int site) {
  using bf = __nv_bfloat16;
  const unsigned grid = (unsigned)(batch * n_heads);
  static bool configured = false;
  if (!configured) {
    XR_CUDA(cudaFuncSetAttribute(enc_attn_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)attn_mma_smem(enc::MAX_L, false)));
    XR_CUDA(cudaFuncSetAttribute(enc_attn_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)attn_mma_smem(enc::MAX_L, true)));
    configured = true;
  }
  if (!dctx) {
    enc_attn_fwd_mma_kernel<<<grid, 256, attn_mma_smem(seq_len, false), s>>>((const bf*)qkv, keymask, seq_len, n_heads,
                                                                            (bf*)out, lse, rng, drop_p, site);
    XR_LAUNCH_CHECK("enc_attn_fwd_mma");
  } else {
    enc_attn_bwd_mma_kernel<<<grid, 256, attn_mma_smem(seq_len, true), s>>>(
        (const bf*)qkv, keymask, (const bf*)ctx, (const bf*)dctx, lse, seq_len, n_heads, (bf*)out, rng, drop_p, site);
    XR_LAUNCH_CHECK("enc_attn_bwd_mma");
  }
  return XR_OK;
}

template <typename T>
static int enc_attention_launch(const void* qkv, const uint8_t* keymask, const void* ctx, const void* dctx,
                                float* lse, int64_t batch, int seq_len, int n_heads, void* out, cudaStream_t s,
                                const int64_t* rng, float drop_p, int site) {
  const unsigned grid = (unsigned)(batch * n_heads);
  if (!dctx) {
    const size_t smem = (size_t)seq_len * 33 * 2 * 4;
    static size_t conf = 0;
    if (smem > 48 * 1024 && smem > conf) {
      XR_CUDA(cudaFuncSetAttribute(enc_attn_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(enc::MAX_L * 33 * 2 * 4)));
      conf = enc::MAX_L * 33 * 2 * 4;
    }
    enc_attn_fwd_kernel<T><<<grid, 256, smem, s>>>((const T*)qkv, keymask, seq_len, n_heads, (T*)out, lse, rng, drop_p,
                                                   site);
    XR_LAUNCH_CHECK("enc_attn_fwd");
  } else {
    const size_t smem = ((size_t)seq_len * 33 * 4 + 2 * (size_t)seq_len) * 4;
    static size_t conf = 0;
    if (smem > 48 * 1024 && smem > conf) {
      XR_CUDA(cudaFuncSetAttribute(enc_attn_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)((enc::MAX_L * 33 * 4 + 2 * enc::MAX_L) * 4)));
      conf = (enc::MAX_L * 33 * 4 + 2 * enc::MAX_L) * 4;
    }
    enc_attn_bwd_kernel<T><<<grid, 256, smem, s>>>((const T*)qkv, keymask, (const T*)ctx, (const T*)dctx, lse, seq_len,
                                                   n_heads, (T*)out, rng, drop_p, site);
    XR_LAUNCH_CHECK("enc_attn_bwd");
  }
  return XR_OK;
}

extern "C" int xr_enc_attention(const void* qkv, const uint8_t* keymask, const void* ctx, const void* dctx,
                                float* lse, int64_t batch, int64_t seq_len, int64_t n_heads, int64_t head_dim,
                                int dtype, void* out, const int64_t* rng, float drop_p, int site, void* stream) {
  XR_CHECK_ARG(qkv && keymask && lse && out, "xr_enc_attention: null pointer");
  XR_CHECK_ARG(head_dim == enc::HD, "xr_enc_attention: this build is specialised for head_dim = %d", enc::HD);
  XR_CHECK_ARG(batch >= 0 && seq_len >= 1 && seq_len <= enc::MAX_L && n_heads >= 1,
               "xr_enc_attention: needs 1 <= seq_len <= %d", enc::MAX_L);
  XR_CHECK_ARG(!dctx || ctx, "xr_enc_attention: the backward needs ctx");
  if (batch == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  if (dtype == XR_F32)
    return enc_attention_launch<float>(qkv, keymask, ctx, dctx, lse, batch, (int)seq_len, (int)n_heads, out, s, rng,
                                       drop_p, site);
  if (dtype == XR_BF16) {
    // the tensor-core kernels read 16-byte row segments: every head slice starts 64 B into a 16 B-aligned row
    XR_CHECK_ARG(((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)ctx | (uintptr_t)dctx) % 16 == 0,
                 "xr_enc_attention: bf16 buffers must be 16-byte aligned");
    return enc_attention_launch_mma(qkv, keymask, ctx, dctx, lse, batch, (int)seq_len, (int)n_heads, out, s, rng,
                                    drop_p, site);
  }
  XR_CHECK_ARG(false, "xr_enc_attention: bad dtype");
  return XR_E_INVALID;
}
