// Family 2: fused query x candidate contraction + loss + gradient on the 5th-gen tensor cores.
//
// Replaces, for the shared-negative-pool candidate set the trainer builds
// (xfmr_rec/models.py:398-416), the whole of EmbedLoss.forward (xfmr_rec/losses.py:150-155):
// compute_logits (:195) -> mask_false_negatives (:283-292) -> loss (:479-488 InfoNCE,
// :498-511 NCE, :520-543 pairwise hinge / logistic, :338-372 + :436-469 contrastive on
// pre-normalised rows) AND its backward w.r.t. the query, in one pass over the pool:
//
//     S  = Q_blk . Neg_tile^T          tcgen05.mma  (A, B from shared memory, D in TMEM)
//     W  = f(S, t_row)                 epilogue warps: tcgen05.ld -> mask / exp / sigmoid ->
//                                      bf16 -> tcgen05.st back into TMEM (aliasing S)
//     dQ += W . Neg_tile               tcgen05.mma  (A = W from TMEM, B = the SAME smem tile,
//                                      now addressed MN-major)
//
// The M x C logits never leave the SM.  Because every per-row normaliser factors out of the
// sum over candidates (softmax with a known reference maximum, weighted means), no online
// rescaling of the accumulator is needed: partial dQ and three scalars per row are written once
// per work item and folded by a small finalize kernel in a fixed order (deterministic, no
// floating-point atomics).
//
// One CTA per SM (640 threads): warp 0 = TMA producer, warp 1 = score-MMA issuer (+ TMEM owner),
// warp 2 = gradient-MMA issuer, warps 4-19 = epilogue (4 TMEM lane quadrants x 4 column groups).
// TMEM (512 columns): dQ accumulator 384 | S 64 | W double buffer 2 x 32 (packed bf16).  Without the
// gradient pass the last 128 columns are an S double buffer.  Shared memory: Q tile 96 KB + a ring
// of 8 pairs of 64x64 bf16 sub-tiles (128 KB) fed by TMA with 128-byte swizzle.
//
// Measured facts the structure follows (profiles/microbench/): the tcgen05.mma queue is ~4 deep,
// so an issuer must never wait for an MMA it has just issued; shared-memory ingest by TMA is
// ~35 B/cycle/SM whatever the ring depth or multicast; a commit -> waiter hand-off costs ~260
// cycles, an mbarrier arrive -> waiter hand-off ~130.
#include "common.cuh"
#include "sm100.cuh"
#include "filter.cuh"
#include "rownorm.cuh"

#include <cstring>

namespace xr {

using namespace sm100;

namespace fk {
constexpr int BM = 128;             // query rows per CTA tile (UMMA M)
constexpr int BN = 64;              // candidates per tile (UMMA N of the score MMA)
constexpr int D = 384;              // embedding dim (all-MiniLM-L6-v2, params.py:11)
constexpr int KB = D / 64;          // 64-column (128-byte) k-blocks per row
constexpr int SLOTS = 16;           // ring of [64 cand x 64 col] sub-tiles ...
constexpr int PAIRS = SLOTS / 2;    // ... managed as 8 pairs (two adjacent k-blocks, 16 KB)
constexpr int SUB_BYTES = BN * 64 * 2;
constexpr int QSUB_BYTES = BM * 64 * 2;
constexpr int Q_BYTES = KB * QSUB_BYTES;
constexpr int RING_BYTES = SLOTS * SUB_BYTES;
constexpr int BAR_OFF = Q_BYTES + RING_BYTES;
constexpr int NBARS = 2 * PAIRS + 12;
constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;  // + manual 1024 B alignment slack
constexpr int EPI_WARPS = 16;           // 4 TMEM lane quadrants x 4 column groups
constexpr int CG = 4;
constexpr int THREADS = 128 + EPI_WARPS * 32;   // producer, 2 MMA issuers, 1 spare, 16 epilogue
constexpr int TMEM_COLS = 512;
constexpr int COL_O = 0, COL_S = 384;   // S: one 64-column buffer with the gradient pass, two without
constexpr int COL_W = 448;              // W: two 32-column buffers of packed bf16 weights (gradient pass)
constexpr int KIND_DIAG = 100;      // extract q_i . pos_i from the diagonal of Q_blk . Pos_blk^T
constexpr int KIND_GMAX = 101;      // retrieval: max score of every 16-column group (no loss)
constexpr int KIND_ALL_DOT = 102;   // forward of every dot-family loss + LogitsStatistics in one pass
constexpr int KIND_ALL_COS = 103;   // forward of the cosine-family losses (+ statistics of the cosine logits)
constexpr int KIND_FILTER = 104;    // retrieval: append every (score, row) with score >= thresh[u] (no loss)
constexpr int NSCAL = 4;
constexpr int NSCAL_ALL = 12;       // per (item, column group, row) scalars of the ALL kinds
}  // namespace fk

// shape + plan of one launch when they are only known on the device (xr_pool_step: the row counts
// come out of the compaction kernel; nothing is read back to the host)
struct FusedDyn {
  int m, cn, nt_count, spl, tiles_per_split, n_items, rb_count, pad;
};

struct FusedParams {
  const FusedDyn* dyn;   // nullable: overrides the seven fields below
  int m, cn, nt_count, spl, tiles_per_split, n_items;
  int mask_fn, logits_bf16, with_grad;
  float scale, margin;
  const float* t;      // [m] target logits (raw fp32)           (loss modes)
  const float* zref;   // [m] softmax reference maximum, nullable (defaults to scaled t)
  float* t_out;        // [m]                                      (KIND_DIAG)
  float* part_o;       // [slots][96 four-column groups][128 rows][4]  partial dQ of every segment
  float* part_s;       // [slots][CG][128][NSCAL]
  float* part_all;     // [slots][CG][128][NSCAL_ALL]              (KIND_ALL_*; dot family of a MON train launch)
  // MON train launches (the train kernel also accumulates what compute_losses logs, trainer.py:250-263):
  float* part_all_cos;      // [slots][CG][128][NSCAL_ALL] cosine-family sums
  const float* mon_inv_q;   // [m]  1 / max(|q_i|, eps)
  const float* mon_inv_n;   // [cn rounded up to a tile] 1 / max(|n_j|, eps)
  const float* mon_t_cos;   // [m]  cosine target logits (t_i / |q_i|) / |pos_i|
  float* gmax;         // [m][gmax_ld] group maxima                (KIND_GMAX)
  long long gmax_ld;
  int tile_stride;     // KIND_GMAX: only every tile_stride-th 64-row catalog tile is scored (0 = 1)
  const float* thresh; // KIND_FILTER: thresh[u * thresh_stride]
  long long thresh_stride;
  FilterOut fo;        // KIND_FILTER: survivor storage (filter.cuh)
  int rb_count;
  int* hang_flag;
  long long* dbg;      // STATS builds: per-tile timestamps of CTA 0 (profiling aid)
  int ctrl_low;        // profiling aid: control roles on hardware warps 0-3 (the round-1 placement)
  int ablate;          // STATS builds: 1 skip TMA loads, 2 skip epilogue math, 4 skip score MMAs, 8 skip dQ MMAs, 16 skip epilogue TMEM ld/st
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// per-element weight / loss term; returns w (goes to the dQ MMA), adds to the row scalars
template <int KIND>
__device__ __forceinline__ float elem(float l, bool valid, float tm, float zref2, float scale2,
                                      float scale, float margin, bool round_scaled, float& cnt,
                                      float& sum_a, float& sum_w) {
  float w = 0.f;
  if (KIND == XR_LOSS_INFONCE) {
    float z2;
    if (round_scaled) z2 = bf16_round(l * scale) * kLog2e;
    else z2 = l * scale2;
    w = valid ? ex2f(z2 - zref2) : 0.f;
    sum_w += w;
  } else if (KIND == XR_LOSS_NCE || KIND == XR_LOSS_PAIRWISE_LOGISTIC) {
    const float x = (KIND == XR_LOSS_NCE) ? l : l - tm;
    const float u = ex2f(-fabsf(x) * kLog2e);          // exp(-|x|) in (0,1]
    const float r = rcpf(1.0f + u);
    const float sp = fmaxf(x, 0.f) + lg2f(1.0f + u) * kLn2;
    const float sg = x >= 0.f ? r : u * r;
    w = valid ? sg : 0.f;
    sum_a += valid ? sp : 0.f;
    sum_w += w;
    cnt += valid ? 1.f : 0.f;
  } else if (KIND == XR_LOSS_PAIRWISE_HINGE) {
    const float x = l - tm;
    const bool on = valid && x > 0.f;
    w = on ? 1.f : 0.f;
    sum_a += on ? x : 0.f;
    sum_w += w;
    cnt += valid ? 1.f : 0.f;
  } else if (KIND == XR_LOSS_CONTRASTIVE || KIND == XR_LOSS_ALIGNMENT_CONTRASTIVE) {
    const float x = l - 1.0f + margin;
    const bool on = valid && x > 0.f;
    w = on ? 1.f : 0.f;
    sum_a += on ? x : 0.f;
    cnt += valid ? 1.f : 0.f;
  }
  return w;
}

// one column group (16 logits of one row) of one tile: logits -> packed bf16 weights + scalars
template <int KIND, bool RBF, bool FULL>
__device__ __forceinline__ void group_math(const uint32_t (&v)[16], uint32_t (&pk)[8], int ncols,
                                           float t_eff, float tm, float zref2, float scale2,
                                           float scale, float margin, bool round_scaled, float& cnt,
                                           float& sum_a, float& sum_w) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    float l0 = __uint_as_float(v[j]), l1 = __uint_as_float(v[j + 1]);
    if (RBF) {   // autocast: the bmm output is bf16 (losses.py:195 under trainer.py:450)
      const __nv_bfloat162 h = __floats2bfloat162_rn(l0, l1);
      const uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
      l0 = __uint_as_float(u << 16);
      l1 = __uint_as_float(u & 0xFFFF0000u);
    }
    bool v0 = l0 < t_eff, v1 = l1 < t_eff;   // strict '<' (losses.py:292); t_eff = +inf if unmasked
    if (!FULL) {
      v0 = v0 && (j < ncols);
      v1 = v1 && (j + 1 < ncols);
    }
    const float w0 = elem<KIND>(l0, v0, tm, zref2, scale2, scale, margin, round_scaled, cnt, sum_a, sum_w);
    const float w1 = elem<KIND>(l1, v1, tm, zref2, scale2, scale, margin, round_scaled, cnt, sum_a, sum_w);
    const __nv_bfloat162 pr = __floats2bfloat162_rn(w0, w1);
    pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&pr);
  }
}

// per-row accumulators of the ALL kinds (trainer.py:250-263 evaluates LogitsStatistics and all
// seven losses every step; here ONE pass over the logits of a family feeds all of them)
struct AllAcc {
  float cnt, s_exp, s_sp, s_hinge, s_logi, s_contr, s_v, s_sq, vmin, vmax;
  __device__ __forceinline__ void reset() {
    cnt = s_exp = s_sp = s_hinge = s_logi = s_contr = s_v = s_sq = 0.f;
    vmin = CUDART_INF_F;
    vmax = -CUDART_INF_F;
  }
};

// one column group (16 logits of one row) for the ALL kinds.  softplus(x) = relu(x) + ln(1 + e^-|x|);
// the 16 factors (1 + e^-|x|) lie in (1, 2], so their PRODUCT (<= 65536) is taken first and one
// lg2 per group replaces sixteen (the MUFU unit bounds this epilogue).
template <bool COS, bool RBF, bool FULL, bool WITH_EXP = true>
__device__ __forceinline__ void group_all(const uint32_t (&v)[16], int ncols, float t_eff, float tm,
                                          float zref2, float scale2, float scale, float margin,
                                          bool round_scaled, AllAcc& a) {
  // masked-out logits are replaced by a huge negative value ONCE, after which every term below
  // vanishes by itself (2^-huge = 0, relu(-huge) = 0, 1 + 0 = 1): this epilogue is bound by issue
  // slots (ncu: 79 % issue-active), so the per-term selects this saves are what matters.  The sums run on
  // PAIRS of logits with the packed fp32x2 instructions of sm_100 (add / mul / fma on two floats per issue
  // slot; each half rounds exactly as the scalar instruction does); selects, min / max and MUFU stay scalar.
  constexpr float kDead = -1.0e30f;
  if (COS) {   // LogitsStatistics is defined on the dot logits (losses.py:383-386): no neg stats here
    float2 contr = make_float2(0.f, 0.f);
    const float2 shift = make_float2(margin - 1.0f, margin - 1.0f);
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float l0 = __uint_as_float(v[j]), l1 = __uint_as_float(v[j + 1]);
      bool ok0 = l0 < t_eff, ok1 = l1 < t_eff;
      if (!FULL) {
        ok0 = ok0 && (j < ncols);
        ok1 = ok1 && (j + 1 < ncols);
      }
      a.cnt += ok0 ? 1.f : 0.f;
      a.cnt += ok1 ? 1.f : 0.f;
      const float2 x = __fadd2_rn(make_float2(ok0 ? l0 : kDead, ok1 ? l1 : kDead), shift);
      contr = __fadd2_rn(contr, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
    }
    a.s_contr += contr.x + contr.y;
    return;
  }
  float2 prod_nce = make_float2(1.f, 1.f), prod_logi = make_float2(1.f, 1.f), relu_nce = make_float2(0.f, 0.f);
  float2 sv = make_float2(0.f, 0.f), ssq = make_float2(0.f, 0.f), sexp = make_float2(0.f, 0.f);
  float2 shinge = make_float2(0.f, 0.f);
  const float2 one2 = make_float2(1.f, 1.f), sc2 = make_float2(scale2, scale2), nz2 = make_float2(-zref2, -zref2);
  const float2 ntm2 = make_float2(-tm, -tm);
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    float l0 = __uint_as_float(v[j]), l1 = __uint_as_float(v[j + 1]);
    if (RBF) {   // autocast: the bmm output is bf16 (losses.py:195 under trainer.py:450)
      const __nv_bfloat162 h = __floats2bfloat162_rn(l0, l1);
      const uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
      l0 = __uint_as_float(u << 16);
      l1 = __uint_as_float(u & 0xFFFF0000u);
    }
    bool ok0 = l0 < t_eff, ok1 = l1 < t_eff;
    if (!FULL) {
      ok0 = ok0 && (j < ncols);
      ok1 = ok1 && (j + 1 < ncols);
    }
    const float2 le = make_float2(ok0 ? l0 : kDead, ok1 ? l1 : kDead);
    a.cnt += ok0 ? 1.f : 0.f;
    a.cnt += ok1 ? 1.f : 0.f;
    const float2 lm = make_float2(ok0 ? l0 : 0.f, ok1 ? l1 : 0.f);
    sv = __fadd2_rn(sv, lm);
    ssq = __ffma2_rn(lm, lm, ssq);
    a.vmin = fminf(a.vmin, fminf(ok0 ? l0 : CUDART_INF_F, ok1 ? l1 : CUDART_INF_F));
    a.vmax = fmaxf(a.vmax, fmaxf(le.x, le.y));
    if (WITH_EXP) {   // (a MON train launch takes the softmax sum from the train epilogue's own weights)
      float2 z;
      if (round_scaled) z = __fadd2_rn(make_float2(bf16_round(le.x * scale) * kLog2e, bf16_round(le.y * scale) * kLog2e), nz2);
      else z = __ffma2_rn(le, sc2, nz2);
      sexp = __fadd2_rn(sexp, make_float2(ex2f(z.x), ex2f(z.y)));
    }
    prod_nce = __fmul2_rn(prod_nce, __fadd2_rn(one2, make_float2(ex2f(-fabsf(le.x) * kLog2e), ex2f(-fabsf(le.y) * kLog2e))));
    relu_nce = __fadd2_rn(relu_nce, make_float2(fmaxf(le.x, 0.f), fmaxf(le.y, 0.f)));
    const float2 x = __fadd2_rn(le, ntm2);
    prod_logi = __fmul2_rn(prod_logi, __fadd2_rn(one2, make_float2(ex2f(-fabsf(x.x) * kLog2e), ex2f(-fabsf(x.y) * kLog2e))));
    shinge = __fadd2_rn(shinge, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
  }
  a.s_v += sv.x + sv.y;
  a.s_sq += ssq.x + ssq.y;
  a.s_exp += sexp.x + sexp.y;
  a.s_hinge += shinge.x + shinge.y;
  // the two half-products are <= 2^8 each: their product stays far inside fp32 range
  a.s_sp += (relu_nce.x + relu_nce.y) + lg2f(prod_nce.x * prod_nce.y) * kLn2;
  a.s_logi += lg2f(prod_logi.x * prod_logi.y) * kLn2;   // + s_hinge (the relu part) in the finalize
}

template <int KIND, bool RBF, int DBG, bool MON = false>
__global__ void __launch_bounds__(fk::THREADS, 1)
fused_pool_kernel(const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_b, const FusedParams p) {
  using namespace fk;
  constexpr bool STATS = DBG == 2;     // wait counters + timeline (profiling aid)
  constexpr bool ABL = DBG != 0;       // ablation switches (timing experiments)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;   // SWIZZLE_128B operands need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t q_smem = base;
  const uint32_t ring = base + Q_BYTES;
  const uint32_t bars = base + BAR_OFF;
  // ring of PAIRS: two adjacent 64x64 sub-tiles (k-blocks 2j, 2j+1 of one candidate tile)
  auto bar_full = [&](uint32_t s) { return bars + 8u * s; };
  auto bar_empty = [&](uint32_t s) { return bars + 8u * (PAIRS + s); };
  const uint32_t bar_q_full = bars + 8u * (2 * PAIRS + 0);
  const uint32_t bar_q_empty = bars + 8u * (2 * PAIRS + 1);
  auto bar_s_full = [&](int b) { return bars + 8u * (2 * PAIRS + 2 + b); };   // S(tile) complete
  auto bar_p_full = [&](int b) { return bars + 8u * (2 * PAIRS + 4 + b); };   // W(tile) stored
  auto bar_s_read = [&](int b) { return bars + 8u * (2 * PAIRS + 6 + b); };   // S(tile) is in registers
  const uint32_t bar_o_full = bars + 8u * (2 * PAIRS + 8);
  const uint32_t bar_o_empty = bars + 8u * (2 * PAIRS + 9);
  auto bar_w_free = [&](int b) { return bars + 8u * (2 * PAIRS + 10 + b); };  // dQ(tile) consumed W
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + NBARS * 8);

  // `warp` is the ROLE index (0 TMA producer, 1 score issuer + TMEM owner, 2 gradient issuer, 3 spare,
  // 4..19 epilogue).  The warp scheduler serves the highest warp id first among eligible warps, so the
  // control roles sit on hardware warps 16-19: the epilogue's instruction bursts must not starve the MMA
  // issue stream (an interrupted stream costs ~200 cycles of tensor pipe each time).  Epilogue warp w
  // keeps w % 4 = its TMEM lane quadrant either way.
  const int hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = p.ctrl_low ? hw_warp : (hw_warp >= EPI_WARPS ? hw_warp - EPI_WARPS : hw_warp + 4);
  constexpr bool diag = (KIND == KIND_DIAG);
  // shape and plan: launch constants, or read from device memory (sync-free step)
  FusedDyn sh{p.m, p.cn, p.nt_count, p.spl, p.tiles_per_split, p.n_items, p.rb_count, 0};
  if (p.dyn) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(p.dyn));
    const int4 b = __ldg(reinterpret_cast<const int4*>(p.dyn) + 1);
    sh = FusedDyn{a.x, a.y, a.z, a.w, b.x, b.y, b.z, 0};
  }
  constexpr bool all_kind = (KIND == KIND_ALL_DOT || KIND == KIND_ALL_COS);
  constexpr bool retr = (KIND == KIND_GMAX || KIND == KIND_FILTER);   // retrieval scoring: no loss, no gradient
  const int ts = (KIND == KIND_GMAX && p.tile_stride > 1) ? p.tile_stride : 1;
  // Loss kinds ("pool" kinds: train + all-losses) run STREAM-K: the rb_count x nt_count (row block, tile)
  // pairs form one linear range, split evenly over the CTAs, so every CTA sweeps the same number of tiles
  // (+-1) whatever M and C are (the (row block, split) items of round 1 left a ragged second wave: 134
  // tiles on the busiest CTA against 127.7 on average at configs[1]).  A CTA's range crosses row-block
  // boundaries; each piece inside one row block is a SEGMENT with its own partial sums in slot
  // blockIdx.x + rb -- unique, because along the linear order either the CTA or the row block advances.
  constexpr bool pool = !diag && !retr;
  long long sk_lo = 0, sk_hi = 0;
  if (pool) {
    const long long TT = (long long)sh.rb_count * sh.nt_count;
    const long long G = TT < (long long)gridDim.x ? TT : (long long)gridDim.x;
    if ((long long)blockIdx.x < G) {
      sk_lo = TT * blockIdx.x / G;
      sk_hi = TT * (blockIdx.x + 1) / G;
    }
  }
  const bool grad = !diag && !retr && !all_kind && p.with_grad;
  // S buffers: ONE with the gradient pass (the other 64 columns hold the W double buffer), two without
  const int nsb_shift = grad ? 0 : 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PAIRS; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_q_full, 1);
    mbar_init(bar_q_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_s_full(b), 1);
      // every epilogue thread waits and arrives for itself: measured faster than one polling /
      // arriving lane per warp plus __syncwarp (reconvergence sits on the critical path)
      mbar_init(bar_p_full(b), EPI_WARPS * 32);
      mbar_init(bar_s_read(b), EPI_WARPS * 32);
      mbar_init(bar_w_free(b), 1);
    }
    mbar_init(bar_o_full, 1);
    mbar_init(bar_o_empty, EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_b);
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_ptr_smem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  auto gtime = [] {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (long long)t;
  };
  const bool dbg0 = ABL && p.dbg && blockIdx.x == 0 && lane == 0;
  if (dbg0 && warp == 1) {
    p.dbg[512 + 0] = clock64();
    p.dbg[512 + 1] = gtime();
  }

  // work items: (row block, split of the candidate tiles); identical iteration in every role
  // the tiles of an item are visited in a rotated order that depends on the row block, so the
  // CTAs sweeping the same candidate range do not hammer the same L2 lines in lockstep
  // (stream-K segments of neighbouring CTAs start ~nt/G tiles apart, so CTAs do not sweep the pool in
  //  lockstep; retrieval row blocks WANT to share catalog tiles through L2)
  auto rot_tile = [&](int t0, int, int, int tl) { return t0 + tl; };
  auto item_tiles = [&](int item, int& rb, int& t0, int& t1) {
    if (diag) {
      rb = item;
      t0 = 2 * rb;
      t1 = t0 + 2;
    } else if (retr) {
      // row block fastest: CTAs that share a catalog range run side by side, so the second
      // query block finds the catalog tiles in L2 instead of re-reading HBM
      rb = item % sh.rb_count;
      const int sp = item / sh.rb_count;
      t0 = sp * sh.tiles_per_split;
      t1 = min(sh.nt_count, t0 + sh.tiles_per_split);
    } else {
      rb = item / sh.spl;
      const int sp = item - rb * sh.spl;
      t0 = sp * sh.tiles_per_split;
      t1 = min(sh.nt_count, t0 + sh.tiles_per_split);
    }
  };

  struct WorkIt {
    long long x;
    int item;
  };
  auto work_next = [&](WorkIt& w, int& rb, int& t0, int& t1, int& slot) -> bool {
    if (pool) {
      if (w.x >= sk_hi) return false;
      rb = (int)(w.x / sh.nt_count);
      t0 = (int)(w.x - (long long)rb * sh.nt_count);
      const long long rem = sk_hi - w.x;
      t1 = (long long)t0 + rem < (long long)sh.nt_count ? t0 + (int)rem : sh.nt_count;
      slot = (int)blockIdx.x + rb;
      w.x += t1 - t0;
      return true;
    }
    if (w.item >= sh.n_items) return false;
    item_tiles(w.item, rb, t0, t1);
    slot = w.item;
    w.item += gridDim.x;
    return true;
  };
  WorkIt wk{sk_lo, (int)blockIdx.x};
  int rb, t0, t1, slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // the whole warp walks the loop (warp-uniform control flow); one elected lane issues
    uint32_t g = 0, it = 0;
    for (; work_next(wk, rb, t0, t1, slot); ++it) {
      mbar_wait<STATS>(bar_q_empty, (it & 1) ^ 1, p.hang_flag, 1);
      if (elect_one()) {
        mbar_expect_tx(bar_q_full, Q_BYTES);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(q_smem + kb * QSUB_BYTES, &tmap_q, bar_q_full, kb * 64, rb * BM);
      }
      __syncwarp();
      if (dbg0 && it < 6) p.dbg[520 + it * 8 + 4] = clock64();
      for (int tl = 0; tl < t1 - t0; ++tl) {
        const int t = rot_tile(t0, t1 - t0, rb, tl);
#pragma unroll 1
        for (int pr = 0; pr < KB / 2; ++pr, ++g) {
          const uint32_t s = g & (PAIRS - 1);
          mbar_wait<STATS>(bar_empty(s), ((g / PAIRS) & 1) ^ 1, p.hang_flag, 2);
          if (elect_one()) {
            if (ABL && (p.ablate & 1)) {
              mbar_arrive(bar_full(s));
            } else {
              mbar_expect_tx(bar_full(s), 2 * SUB_BYTES);
              tma_load_2d(ring + s * 2 * SUB_BYTES, &tmap_b, bar_full(s), pr * 128, t * ts * BN);
              tma_load_2d(ring + s * 2 * SUB_BYTES + SUB_BYTES, &tmap_b, bar_full(s), pr * 128 + 64, t * ts * BN);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ========================= score MMA issuer: S = Q . Neg^T =========================
    // The tcgen05.mma queue is only ~4 instructions deep (profiles/microbench/umma_mix.cu), so an
    // issuer that waits for the completion of an MMA it has just issued drains the tensor pipe.
    // Neither issuer does: S(t+1) needs the epilogue to have LOADED S(t) into registers
    // (bar_s_read, ~100 cycles after S(t) completes, hidden behind dQ(t-1)), dQ(t) needs W(t),
    // which the epilogue finished a whole S pass earlier.  Steady-state pipe order:
    //     S(t) dQ(t-1) S(t+1) dQ(t) ...
    // warp-uniform control flow; descriptors live in uniform registers, one elected lane issues
    constexpr uint32_t idesc_s = umma_idesc_bf16(BM, BN, 0, 0);    // both operands K-major
    const uint64_t q_desc0 = umma_desc_sw128(q_smem, 16, 1024);
    const uint64_t ring_k_desc0 = umma_desc_sw128(ring, 16, 1024);
    uint32_t g = 0, tt = 0, it = 0;
    for (; work_next(wk, rb, t0, t1, slot); ++it) {
      const int T = t1 - t0;
      mbar_wait<STATS>(bar_q_full, it & 1, p.hang_flag, 4);
      if (dbg0 && it < 6) p.dbg[520 + it * 8 + 0] = clock64();
      for (int tl = 0; tl < T; ++tl) {
        const uint32_t tile = tt + tl;
        const uint32_t use = tile >> nsb_shift;            // how often this S buffer was used before
        const int sb = tile & ((1 << nsb_shift) - 1);
        if (ABL && p.dbg && blockIdx.x == 0 && lane == 0 && tile < 64) p.dbg[tile * 8 + 0] = clock64();
        // TWO elect blocks per tile (pairs 0+1, then pair 2): every interruption of the MMA issue
        // stream (warp reconvergence, descriptor set-up, a wait) costs ~200 cycles of tensor pipe
        // (profiles/microbench/umma_mix.cu).  The waits stay OUTSIDE the blocks, warp-uniform: a
        // wait loop inside an elect block makes ptxas build every descriptor in vector registers
        // (16+ R2UR per pair instead of UIADD3.64 on a uniform base).
        auto issue_pair = [&](int pr) {
          const uint32_t s = (g + pr) & (PAIRS - 1);
          const uint64_t a0 = q_desc0 + (uint64_t)(pr * ((2 * QSUB_BYTES) >> 4));
          const uint64_t b0 = ring_k_desc0 + (uint64_t)(s * ((2 * SUB_BYTES) >> 4));
          if (!(ABL && (p.ablate & 4)))
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss(tmem + COL_S + sb * BN, a0 + h * (QSUB_BYTES >> 4) + 2 * k,
                      b0 + h * (SUB_BYTES >> 4) + 2 * k, idesc_s, (pr | h | k) ? 1u : 0u);
          if (!grad) umma_commit(bar_empty(s));   // forward only: the pair is free after S
        };
        if (use >= 1) mbar_wait<STATS>(bar_s_read(sb), (use - 1) & 1, p.hang_flag, 6);
        mbar_wait<STATS>(bar_full(g & (PAIRS - 1)), (g / PAIRS) & 1, p.hang_flag, 7);
        mbar_wait<STATS>(bar_full((g + 1) & (PAIRS - 1)), ((g + 1) / PAIRS) & 1, p.hang_flag, 7);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1   // runtime pair index: keeps ptxas from hoisting 24 descriptors into vector registers
          for (int pr = 0; pr < 2; ++pr) issue_pair(pr);
        }
        __syncwarp();
        // (peeking at the third pair to issue the whole tile in one block measured slower: the
        // extra try_wait + shuffle cost more than the saved reconvergence)
        mbar_wait<STATS>(bar_full((g + 2) & (PAIRS - 1)), ((g + 2) / PAIRS) & 1, p.hang_flag, 7);
        tc_fence_after();
        if (elect_one()) {
          int pr2 = 2;
          asm volatile("" : "+r"(pr2));   // opaque to the optimiser for the same reason
          issue_pair(pr2);
          umma_commit(bar_s_full(sb));
          if (tl == T - 1) umma_commit(bar_q_empty);   // Q is only read by the score MMAs
        }
        __syncwarp();
        g += KB / 2;
        if (ABL && p.dbg && blockIdx.x == 0 && lane == 0 && tile < 64) p.dbg[tile * 8 + 1] = clock64();
        if (dbg0 && it < 6 && tl == T - 1) p.dbg[520 + it * 8 + 1] = clock64();
      }
      tt += T;
    }
    if (dbg0) {
      p.dbg[512 + 2] = clock64();
      p.dbg[512 + 3] = gtime();
    }
  } else if (warp == 2) {
    // ========================= gradient MMA issuer: dQ += W . Neg =========================
    // A = W from TMEM (8 columns per K=16 step), B = the tile's pairs addressed MN-major.
    // pair-outer order: a pair goes back to the producer as soon as ITS four K-steps are done, so
    // the reload of the tile's 48 KB (shared-memory ingest is ~35 B/cycle/SM,
    // profiles/microbench/tma_stream.cu) starts two thirds of a dQ pass earlier
    if (grad) {
      constexpr uint32_t idesc_o = umma_idesc_bf16(BM, 128, 0, 1);
      // MN-major view of a pair: two 64-column atoms SUB_BYTES apart, 8-row groups of 1 KB
      const uint64_t ring_mn_desc0 = umma_desc_sw128(ring, SUB_BYTES, 1024);
      uint32_t g = 0, tt = 0, it = 0;
      for (; work_next(wk, rb, t0, t1, slot); ++it) {
        const int T = t1 - t0;
        mbar_wait<STATS>(bar_o_empty, (it & 1) ^ 1, p.hang_flag, 5);   // epilogue drained the last dQ
        if (dbg0 && it < 6) p.dbg[520 + it * 8 + 5] = clock64();
        for (int tl = 0; tl < T; ++tl, g += KB / 2) {
          const uint32_t tile = tt + tl;
          const int wb = tile & 1;
          mbar_wait<STATS>(bar_p_full(wb), (tile >> 1) & 1, p.hang_flag, 3);
          tc_fence_after();
          if (ABL && p.dbg && blockIdx.x == 0 && lane == 0 && tile < 64) p.dbg[tile * 8 + 5] = clock64();
          if (elect_one()) {
            const uint32_t a_tmem = tmem + COL_W + wb * (BN / 2);
#pragma unroll
            for (int pr = 0; pr < KB / 2; ++pr) {
              const uint32_t s = (g + pr) & (PAIRS - 1);
              if (!(ABL && (p.ablate & 8)))
#pragma unroll
              for (int ks = 0; ks < BN / 16; ++ks) {
                const uint64_t bdesc = ring_mn_desc0 + (uint64_t)(s * ((2 * SUB_BYTES) >> 4) + ks * (2048 >> 4));
                umma_ts(tmem + COL_O + pr * 128, a_tmem + ks * 8, bdesc, idesc_o,
                        (tl == 0 && ks == 0) ? 0u : 1u);
              }
              umma_commit(bar_empty(s));
            }
            umma_commit(bar_w_free(wb));
            if (tl == T - 1) umma_commit(bar_o_full);
          }
          __syncwarp();
        }
        tt += T;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue warps ==============================
    // 16 warps = 4 TMEM lane quadrants x 4 column groups of 16 logits: enough warps per scheduler
    // to hide the MUFU / conversion latencies behind each other
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may touch
    const int cg = (warp - 4) >> 2;                  // column group
    const int r_local = quad * 32 + lane;
    const uint32_t tmem_lane = tmem + ((uint32_t)(quad * 32) << 16);
    const bool round_scaled = RBF && p.scale != 1.0f;
    const float scale2 = p.scale * kLog2e;
    uint32_t tt = 0, it = 0;
    for (; work_next(wk, rb, t0, t1, slot); ++it) {
      const int T = t1 - t0;
      const int row = rb * BM + r_local;
      const bool row_ok = row < sh.m;
      float t = 0.f, tm = 0.f, zref2 = 0.f, t_eff = 0.f;
      if (!diag && !retr && row_ok) {
        t = p.t[row];
        if (RBF) t = bf16_round(t);
        tm = __fmul_rn(t, 1.0f - p.margin);   // rounded on its own: see rowloss.cuh
        float zr;
        if (p.zref) zr = p.zref[row];
        else zr = round_scaled ? bf16_round(t * p.scale) : t * p.scale;
        zref2 = zr * kLog2e;
      }
      t_eff = p.mask_fn ? t : CUDART_INF_F;
      if (KIND == KIND_FILTER)   // a row past m never passes; thresholds are group maxima (never NaN)
        t_eff = row_ok ? __ldg(p.thresh + (long long)row * p.thresh_stride) : CUDART_INF_F;
      float cnt = 0.f, sum_a = 0.f, sum_w = 0.f, diag_val = 0.f;
      AllAcc acc;
      acc.reset();
      // MON: the cosine family is evaluated on the SAME scores, l_cos = (s / |q|) / |n| (losses.py:206-208 with
      // fp32 inverse norms), against the cosine target built with the same association, so a pool entry equal
      // to the row's positive still ties exactly
      AllAcc acc_c;
      acc_c.reset();
      float mon_iq = 0.f, mon_tc_eff = CUDART_INF_F;
      if (MON && row_ok) {
        mon_iq = p.mon_inv_q[row];
        if (p.mask_fn) mon_tc_eff = p.mon_t_cos[row];
      }
      // KIND_FILTER: this lane's sub-bucket for the item = (row, catalog split, column group); its fill
      // count lives in a register (see score_gmax2_sm100.cu)
      const int sub = KIND == KIND_FILTER ? (slot / sh.rb_count) * CG + cg : 0;
      float* b_scores = nullptr;
      int32_t* b_rows = nullptr;
      int bcount = 0;
      if (KIND == KIND_FILTER && row_ok) {
        const long long off = ((long long)row * p.fo.n_sub + sub) * p.fo.cap_b;
        b_scores = p.fo.b_scores + off;
        b_rows = p.fo.b_rows + off;
      }
      for (int tl = 0; tl < T; ++tl) {
        const uint32_t tile = tt + tl;
        const uint32_t use = tile >> nsb_shift;
        const int sb = tile & ((1 << nsb_shift) - 1);
        const int wb = tile & 1;
        mbar_wait<STATS>(bar_s_full(sb), use & 1, p.hang_flag, 8);
        tc_fence_after();
        const bool dbg_on = ABL && p.dbg && blockIdx.x == 0 && warp == 4 && lane == 0 && tile < 64;
        if (dbg_on) p.dbg[tile * 8 + 2] = clock64();
        uint32_t v[16];
        if (ABL && (p.ablate & 16)) {   // no TMEM traffic from the epilogue at all
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0x3c003c00u + tile;
        } else {
          tmem_ld16(tmem_lane + COL_S + sb * BN + cg * 16, v);
          tmem_wait_ld();
        }
        // the logits are in registers: hand the S buffer back to the score issuer right away
        const bool late_read = ABL && (p.ablate & 64);
        if (!late_read) {
          tc_fence_before();
          mbar_arrive(bar_s_read(sb));
        }
        if (dbg_on) p.dbg[tile * 8 + 3] = clock64();
        if (diag) {
          // tile tl holds pos rows [rb*128 + tl*64, +64): the diagonal entry of local row r is
          // column r - tl*64 of tile tl = r/64, owned by column group (r%64)/16
          if ((r_local >> 6) == tl && ((r_local & 63) >> 4) == cg) {
            const int c = r_local & 15;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c == j) diag_val = __uint_as_float(v[j]);
          }
        } else if (KIND == KIND_GMAX) {
          // storage column 4 t + cg = catalog rows [64 t ts + 16 cg, +16): natural order for ts = 1
          const int ncols = sh.cn - (t0 + tl) * ts * BN - cg * 16;
          float mx = -CUDART_INF_F;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < ncols) mx = fmaxf(mx, __uint_as_float(v[j]));
          if (row_ok) p.gmax[(long long)row * p.gmax_ld + (long long)(t0 + tl) * CG + cg] = mx;
        } else if (KIND == KIND_FILTER) {
          // one test for the 16 scores: almost no lane-tile holds a survivor at the sample's threshold
          const int c0 = (t0 + tl) * BN + cg * 16;
          const int lim = sh.cn - c0;               // rows past cn are TMA zero fill, not catalog rows
          float mx = -CUDART_INF_F;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
          if (mx >= t_eff && lim > 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float sc = __uint_as_float(v[j]);
              if (sc >= t_eff && j < lim) filter_keep(p.fo, row, b_scores, b_rows, bcount, sc, c0 + j);
            }
          }
        } else if (all_kind) {
          const int ncols = sh.cn - rot_tile(t0, T, rb, tl) * BN - cg * 16;
          if (ncols >= 16)
            group_all<KIND == KIND_ALL_COS, RBF, true>(v, ncols, t_eff, tm, zref2, scale2, p.scale,
                                                       p.margin, round_scaled, acc);
          else
            group_all<KIND == KIND_ALL_COS, RBF, false>(v, ncols, t_eff, tm, zref2, scale2, p.scale,
                                                        p.margin, round_scaled, acc);
        } else {
          const int ncols = sh.cn - rot_tile(t0, T, rb, tl) * BN - cg * 16;   // valid candidates in this group
          uint32_t pk[8];
          if (ABL && (p.ablate & 2)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = v[j];
          } else if (ncols >= 16)
            group_math<KIND, RBF, true>(v, pk, ncols, t_eff, tm, zref2, scale2, p.scale, p.margin,
                                        round_scaled, cnt, sum_a, sum_w);
          else
            group_math<KIND, RBF, false>(v, pk, ncols, t_eff, tm, zref2, scale2, p.scale, p.margin,
                                         round_scaled, cnt, sum_a, sum_w);
          if (grad) {
            // W (16 bf16 = 8 packed columns per group) goes to its own double buffer, so the S
            // buffer never waits for a gradient MMA; buffer wb was last read by dQ(tile - 2)
            if (dbg_on) p.dbg[tile * 8 + 6] = clock64();
            if (tile >= 2 && !(ABL && (p.ablate & 32))) {
              mbar_wait<STATS>(bar_w_free(wb), ((tile - 2) >> 1) & 1, p.hang_flag, 10);
              tc_fence_after();
            }
            if (!(ABL && (p.ablate & 16))) {
              tmem_st8(tmem_lane + COL_W + wb * (BN / 2) + cg * 8, pk);
              tmem_wait_st();
            }
            if (dbg_on) p.dbg[tile * 8 + 7] = clock64();
            tc_fence_before();
            mbar_arrive(bar_p_full(wb));
          }
          if (MON) {   // after W is on its way to the gradient MMA: the monitoring sums are off the S -> W -> dQ chain
            const int c0 = rot_tile(t0, T, rb, tl) * BN + cg * 16;
            uint32_t vc[16];
            const float4* ninv = reinterpret_cast<const float4*>(p.mon_inv_n + c0);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float4 w4 = __ldg(ninv + jj);
              vc[4 * jj + 0] = __float_as_uint(__fmul_rn(__fmul_rn(__uint_as_float(v[4 * jj + 0]), mon_iq), w4.x));
              vc[4 * jj + 1] = __float_as_uint(__fmul_rn(__fmul_rn(__uint_as_float(v[4 * jj + 1]), mon_iq), w4.y));
              vc[4 * jj + 2] = __float_as_uint(__fmul_rn(__fmul_rn(__uint_as_float(v[4 * jj + 2]), mon_iq), w4.z));
              vc[4 * jj + 3] = __float_as_uint(__fmul_rn(__fmul_rn(__uint_as_float(v[4 * jj + 3]), mon_iq), w4.w));
            }
            if (ncols >= 16) {
              group_all<false, RBF, true, KIND != XR_LOSS_INFONCE>(v, ncols, t_eff, tm, zref2, scale2, p.scale, p.margin,
                                                                  round_scaled, acc);
              group_all<true, false, true>(vc, ncols, mon_tc_eff, 0.f, 0.f, 0.f, 1.f, p.margin, false, acc_c);
            } else {
              group_all<false, RBF, false, KIND != XR_LOSS_INFONCE>(v, ncols, t_eff, tm, zref2, scale2, p.scale, p.margin,
                                                                   round_scaled, acc);
              group_all<true, false, false>(vc, ncols, mon_tc_eff, 0.f, 0.f, 0.f, 1.f, p.margin, false, acc_c);
            }
          }
        }
        if (late_read) {
          tc_fence_before();
          mbar_arrive(bar_s_read(sb));
        }
        if (dbg_on) p.dbg[tile * 8 + 4] = clock64();
      }
      if (diag) {
        if (row_ok && (r_local & 63) >> 4 == cg) p.t_out[row] = diag_val;
      } else if (retr) {
        // group maxima / survivors were written tile by tile; the filter leaves its sub-bucket's count
        if (KIND == KIND_FILTER && row_ok) p.fo.b_count[(long long)row * p.fo.n_sub + sub] = bcount;
      } else if (all_kind) {
        if (row_ok) {
          float4* ds = reinterpret_cast<float4*>(
              p.part_all + (((size_t)slot * CG + cg) * BM + r_local) * NSCAL_ALL);
          ds[0] = make_float4(acc.cnt, acc.s_exp, acc.s_sp, acc.s_hinge);
          ds[1] = make_float4(acc.s_logi, acc.s_contr, acc.s_v, acc.s_sq);
          ds[2] = make_float4(acc.vmin, acc.vmax, 0.f, 0.f);
        }
      } else {
        if (grad) {
          mbar_wait<STATS>(bar_o_full, it & 1, p.hang_flag, 9);
          tc_fence_after();
          if (dbg0 && warp == 4 && it < 6) p.dbg[520 + it * 8 + 2] = clock64();
          // the partial dQ leaves in a BLOCKED layout [4-column group (96)][row (128)][4 floats]: lane =
          // row, so the 16-byte stores of a warp are 512 contiguous bytes.  (Row-major rows scattered
          // every store instruction over 32 lines: 22k cycles per drain, 12 % of the kernel, measured.)
          float* dst = p.part_o + (size_t)slot * BM * D + (size_t)r_local * 4;
#pragma unroll 1
          for (int c = 0; c < D / CG / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_lane + COL_O + cg * (D / CG) + c * 32, o);
            tmem_wait_ld();
            if (row_ok) {
              const int c4 = (cg * (D / CG) + c * 32) / 4;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(dst + (size_t)(c4 + j) * (BM * 4)) =
                    make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (dbg0 && warp == 4 && it < 6) p.dbg[520 + it * 8 + 3] = clock64();
          if (lane == 0) mbar_arrive(bar_o_empty);
        }
        if (row_ok) {
          float* ds = p.part_s + (((size_t)slot * CG + cg) * BM + r_local) * NSCAL;
          *reinterpret_cast<float4*>(ds) = make_float4(cnt, sum_a, sum_w, 0.f);
        }
        if (MON && row_ok) {
          const size_t off = (((size_t)slot * CG + cg) * BM + r_local) * NSCAL_ALL;
          float4* dd = reinterpret_cast<float4*>(p.part_all + off);
          // an InfoNCE train launch does not recompute the softmax sum: it IS the train epilogue's sum of
          // weights (same exponentials, same mask); the other train kinds accumulate it with the rest
          dd[0] = make_float4(acc.cnt, KIND == XR_LOSS_INFONCE ? sum_w : acc.s_exp, acc.s_sp, acc.s_hinge);
          dd[1] = make_float4(acc.s_logi, acc.s_contr, acc.s_v, acc.s_sq);
          dd[2] = make_float4(acc.vmin, acc.vmax, 0.f, 0.f);
          float4* dc = reinterpret_cast<float4*>(p.part_all_cos + off);
          dc[0] = make_float4(acc_c.cnt, 0.f, 0.f, 0.f);
          dc[1] = make_float4(0.f, acc_c.s_contr, 0.f, 0.f);
          dc[2] = make_float4(CUDART_INF_F, -CUDART_INF_F, 0.f, 0.f);
        }
      }
      tt += T;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// stream-K bookkeeping shared by the kernel above and the finalize kernels: CTA c of G sweeps the linear
// tile range [TT c / G, TT (c + 1) / G); the CTA that holds linear tile x is the largest c whose range
// starts at or before x
__host__ __device__ inline int sk_owner(long long x, long long TT, long long G) {
  return (int)(((x + 1) * G + TT - 1) / TT - 1);
}
struct SkRange {
  int c_first, c_last;
};
__device__ __forceinline__ SkRange sk_segments(int rb, int nt, int rb_count, int grid) {
  const long long TT = (long long)rb_count * nt;
  const long long G = TT < (long long)grid ? TT : (long long)grid;
  return SkRange{sk_owner((long long)rb * nt, TT, G), sk_owner((long long)(rb + 1) * nt - 1, TT, G)};
}

// ---- finalize: fold the per-segment partials, apply the per-row normalisers and the positive's
//      term, chain through the query normalisation for the cosine kinds.  A block takes FR consecutive
//      rows: phase A sums their partial dQ over the segments of their row block (fixed order:
//      deterministic) straight from the blocked layout -- 16-byte loads, 256 contiguous bytes per
//      half-warp -- into shared memory; phase B finishes the rows, one warp per row ------------------
constexpr int FR = 16;
__global__ void __launch_bounds__(256)
fused_finalize_kernel(const float* __restrict__ part_o, const float* __restrict__ part_s,
                      const float* __restrict__ t_raw, const float* __restrict__ zref,
                      const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ pos,
                      const float* __restrict__ q_inv, int m, int cn, int grid, int kind, int logits_bf16,
                      float scale, float margin, float grad_scale, float* __restrict__ dq,
                      float* __restrict__ row_loss, const FusedDyn* __restrict__ dyn,
                      const int64_t* __restrict__ inv_pos, const int64_t* __restrict__ sel_pos,
                      int64_t n_pos, void* __restrict__ dtok, int dtok_bf16) {
  using namespace fk;
  __shared__ float s_o[FR][D + 4];
  __shared__ float s_sc[FR][4];
  if (dyn) {
    m = dyn->m;
    cn = dyn->cn;
  }
  const int rb_count = (m + BM - 1) / BM, nt = (cn + BN - 1) / BN;
  // scatter mode (xr_pool_step): row i of the compacted order goes to position sel_pos[i] of the
  // encoder output's own layout and dtype; positions that hold no row get a zero gradient row -- the
  // autograd of token_embeddings[mask][pos_mask] (models.py:392, 415) folded into this kernel
  const bool scatter = dtok != nullptr;
  const bool want_o = dq || scatter;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool cosine = kind == XR_LOSS_CONTRASTIVE || kind == XR_LOSS_ALIGNMENT_CONTRASTIVE;
  constexpr int NP = D / 64;   // each lane owns the column PAIRS {64 c + 2 lane, +1}
  const int n_groups = (m + FR - 1) / FR;
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int row0 = g * FR;
    const int rb = row0 / BM, rl0 = row0 % BM;
    const SkRange sr = sk_segments(rb, nt, rb_count, grid);
    // ---- phase A ----
    if (want_o) {
      // a thread owns 6 of the 96 column quads of one row: all six loads of a segment are in flight together
      // (one load at a time left this phase waiting on L2 latency: ncu, 15.7 long-scoreboard stalls per issue)
      const int r = threadIdx.x & (FR - 1);
      constexpr int NQ = (D / 4) / (256 / FR);   // 6
      float4 acc[NQ];
#pragma unroll
      for (int k = 0; k < NQ; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = sr.c_first; c <= sr.c_last; ++c) {   // fixed order: deterministic
        const float* seg = part_o + (size_t)(c + rb) * BM * D + (size_t)(rl0 + r) * 4;
        float4 v[NQ];
#pragma unroll
        for (int k = 0; k < NQ; ++k)
          v[k] = *reinterpret_cast<const float4*>(seg + (size_t)(threadIdx.x / FR + k * (256 / FR)) * BM * 4);
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
          acc[k].x += v[k].x; acc[k].y += v[k].y; acc[k].z += v[k].z; acc[k].w += v[k].w;
        }
      }
#pragma unroll
      for (int k = 0; k < NQ; ++k)
        *reinterpret_cast<float4*>(&s_o[r][(threadIdx.x / FR + k * (256 / FR)) * 4]) = acc[k];
    }
    if (threadIdx.x < FR) {
      float cnt = 0.f, sum_a = 0.f, sum_w = 0.f;
      for (int c = sr.c_first; c <= sr.c_last; ++c) {
#pragma unroll
        for (int cg = 0; cg < CG; ++cg) {
          const float4 v = *reinterpret_cast<const float4*>(
              part_s + (((size_t)(c + rb) * CG + cg) * BM + rl0 + threadIdx.x) * NSCAL);
          cnt += v.x; sum_a += v.y; sum_w += v.z;
        }
      }
      s_sc[threadIdx.x][0] = cnt;
      s_sc[threadIdx.x][1] = sum_a;
      s_sc[threadIdx.x][2] = sum_w;
    }
    __syncthreads();
    // ---- phase B ----
    for (int r = warp; r < FR; r += 8) {
      const int64_t i = row0 + r;
      if (i >= m) break;
      const float cnt = s_sc[r][0], sum_a = s_sc[r][1], sum_w = s_sc[r][2];
      float t = t_raw[i];
      if (logits_bf16) t = bf16_round(t);
      const float den = cnt + 1e-9f;
      float loss = 0.f, co = 0.f, cp = 0.f;   // dq = co * O + cp * pos
      switch (kind) {
        case XR_LOSS_INFONCE: {
          const float st = (logits_bf16 && scale != 1.0f) ? bf16_round(t * scale) : t * scale;
          const float zr = zref ? zref[i] : st;
          const float et = __expf(st - zr);
          const float zall = sum_w + et;
          loss = zr + __logf(zall) - st;
          co = scale / zall;
          cp = scale * (et / zall - 1.0f);
          break;
        }
        case XR_LOSS_NCE: {
          const float at = fabsf(t);
          loss = fmaxf(-t, 0.f) + log1pf(__expf(-at)) + sum_a / den;
          co = 1.0f / den;
          cp = -1.0f / (1.0f + __expf(t));
          break;
        }
        case XR_LOSS_PAIRWISE_HINGE:
        case XR_LOSS_PAIRWISE_LOGISTIC:
          loss = sum_a / den;
          co = 1.0f / den;
          cp = -(1.0f - margin) * sum_w / den;
          break;
        case XR_LOSS_CONTRASTIVE:
          loss = sum_a / den;
          co = 1.0f / den;
          break;
        case XR_LOSS_ALIGNMENT_CONTRASTIVE:
          loss = sum_a / den + (1.0f - t);
          co = 1.0f / den;
          cp = -1.0f;
          break;
        default:
          break;
      }
      if (lane == 0 && row_loss) row_loss[i] = loss;
      if (want_o) {
        float2 gr[NP];
        float dot = 0.f;
        const __nv_bfloat162* pos2 = reinterpret_cast<const __nv_bfloat162*>(pos + i * D);
        const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(q + i * D);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          const float2 o = *reinterpret_cast<const float2*>(&s_o[r][c * 64 + 2 * lane]);
          const float2 pv = __bfloat1622float2(pos2[c * 32 + lane]);
          gr[c].x = co * o.x + cp * pv.x;
          gr[c].y = co * o.y + cp * pv.y;
          if (cosine) {
            const float2 qv = __bfloat1622float2(q2[c * 32 + lane]);
            dot = fmaf(gr[c].x, qv.x, dot);
            dot = fmaf(gr[c].y, qv.y, dot);
          }
        }
        if (cosine) {
          dot = warp_sum(dot);
          const float inv = q_inv[i];
#pragma unroll
          for (int c = 0; c < NP; ++c) {
            const float2 qv = __bfloat1622float2(q2[c * 32 + lane]);
            gr[c].x = inv * (gr[c].x - dot * qv.x);
            gr[c].y = inv * (gr[c].y - dot * qv.y);
          }
        }
        if (!scatter) {
          float2* dst = reinterpret_cast<float2*>(dq + i * D);
#pragma unroll
          for (int c = 0; c < NP; ++c) dst[c * 32 + lane] = make_float2(gr[c].x * grad_scale, gr[c].y * grad_scale);
        } else {
          const int64_t where = sel_pos[i];
          if (dtok_bf16) {
            __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(dtok) + where * D);
#pragma unroll
            for (int c = 0; c < NP; ++c)
              dst[c * 32 + lane] = __floats2bfloat162_rn(gr[c].x * grad_scale, gr[c].y * grad_scale);
          } else {
            float2* dst = reinterpret_cast<float2*>(reinterpret_cast<float*>(dtok) + where * D);
#pragma unroll
            for (int c = 0; c < NP; ++c) dst[c * 32 + lane] = make_float2(gr[c].x * grad_scale, gr[c].y * grad_scale);
          }
        }
      }
    }
    __syncthreads();
  }
  if (scatter) {   // zero rows for the positions that hold no row
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = gw; it < n_pos; it += nw) {
      if (inv_pos[it] >= 0) continue;
      if (dtok_bf16) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(dtok) + it * D);
#pragma unroll
        for (int c = 0; c < NP; ++c) dst[c * 32 + lane] = 0u;
      } else {
        float2* dst = reinterpret_cast<float2*>(reinterpret_cast<float*>(dtok) + it * D);
#pragma unroll
        for (int c = 0; c < NP; ++c) dst[c * 32 + lane] = make_float2(0.f, 0.f);
      }
    }
  }
}

// ---- finalize of the ALL kinds: fold the per-item scalars of every row (fixed order) into the row
//      slots rowloss_reduce_kernel sums -- the same slots, formulas and float/double mix as
//      rowloss_kernel (rowloss.cu), which is the materialised-logits implementation of this pass ----
__global__ void __launch_bounds__(256)
fused_finalize_all_kernel(const float* __restrict__ part_all, const float* __restrict__ t_raw,
                          const float* __restrict__ zref, int m, int cn, int grid, int cosine, int logits_bf16,
                          float scale, float margin, double* __restrict__ row_out,
                          const FusedDyn* __restrict__ dyn, const float* __restrict__ part_all2 = nullptr,
                          const float* __restrict__ t_raw2 = nullptr, double* __restrict__ row_out2 = nullptr) {
  using namespace fk;
  if (dyn) {
    m = dyn->m;
    cn = dyn->cn;
  }
  if (blockIdx.y == 1) {   // one-pass monitoring: the cosine family's sums, targets and rows
    part_all = part_all2;
    t_raw = t_raw2;
    zref = nullptr;
    cosine = 1;
    logits_bf16 = 0;
    row_out = row_out2;
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int rb = i / BM, rl = i % BM;
  const SkRange sr = sk_segments(rb, (cn + BN - 1) / BN, (m + BM - 1) / BM, grid);
  double cnt = 0, s_exp = 0, s_sp = 0, s_hinge = 0, s_logi = 0, s_contr = 0, s_v = 0, s_sq = 0;
  float vmin = CUDART_INF_F, vmax = -CUDART_INF_F;
  for (int seg = sr.c_first; seg <= sr.c_last; ++seg) {   // the segments of this row block, fixed order
    const size_t item = (size_t)(seg + rb);
#pragma unroll
    for (int cg = 0; cg < CG; ++cg) {
      const float4* src = reinterpret_cast<const float4*>(part_all + ((item * CG + cg) * BM + rl) * NSCAL_ALL);
      const float4 a = src[0], b = src[1], c = src[2];
      cnt += a.x; s_exp += a.y; s_sp += a.z; s_hinge += a.w;
      s_logi += b.x; s_contr += b.y; s_v += b.z; s_sq += b.w;
      vmin = fminf(vmin, c.x);
      vmax = fmaxf(vmax, c.y);
    }
  }
  float t = t_raw[i];
  if (logits_bf16) t = bf16_round(t);
  const float den = (float)cnt + 1e-9f;
  double* o = row_out + (size_t)i * ROW_SLOTS;
  o[S_ALIGN] = 1.0 - (double)t;
  o[S_CONTR] = s_contr / (double)den;
  double infonce = 0, nce = 0, hinge = 0, logi = 0;
  if (!cosine) {
    const float st = (logits_bf16 && scale != 1.0f) ? bf16_round(t * scale) : t * scale;
    const float zr = zref ? zref[i] : st;
    infonce = (double)zr + log(s_exp + (double)__expf(st - zr)) - (double)st;
    nce = (double)(fmaxf(-t, 0.f) + log1pf(__expf(-fabsf(t)))) + s_sp / (double)den;
    hinge = s_hinge / (double)den;
    logi = (s_hinge + s_logi) / (double)den;
  }
  o[S_INFONCE] = infonce;
  o[S_NCE] = nce;
  o[S_HINGE] = hinge;
  o[S_LOGISTIC] = logi;
  o[S_DENS] = cnt;
  o[S_POS] = (double)t;
  o[S_NCOUNT] = cnt;
  o[S_NSUM] = s_v;
  o[S_NSQ] = s_sq;
  o[S_NMIN] = (double)vmin;
  o[S_NMAX] = (double)vmax;
}

// single block, fixed order: loss = sum_i row_loss[i]   (double accumulation)
__global__ void __launch_bounds__(1024)
sum_rows_kernel(const float* __restrict__ row_loss, int64_t m, double* __restrict__ out,
                float* __restrict__ out_f32, const FusedDyn* __restrict__ dyn) {
  __shared__ double s[32];
  if (dyn) m = dyn->m;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) acc += (double)row_loss[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s[w];
    out[0] = tot;
    if (out_f32) out_f32[0] = (float)tot;
  }
}

// reference maximum when the false-negative mask is off (losses.py:283-287): by Cauchy-Schwarz
// |s * q.n| <= |s| * ||q|| * max_j ||n_j||, and the target itself is in the set.
__global__ void negnorm_max_kernel(const __nv_bfloat16* __restrict__ neg, int64_t cn,
                                   unsigned* __restrict__ out_bits, const FusedDyn* __restrict__ dyn) {
  if (dyn) cn = dyn->cn;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float best = 0.f;
  for (int64_t r = warp; r < cn; r += nwarps) {
    float ss = 0.f;
    for (int c = lane; c < fk::D; c += 32) {
      const float v = __bfloat162float(neg[r * fk::D + c]);
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    best = fmaxf(best, sqrtf(ss));
  }
  if (lane == 0) atomicMax(out_bits, __float_as_uint(best));   // non-negative floats order as uints
}
__global__ void zref_bound_kernel(const __nv_bfloat16* __restrict__ q, const float* __restrict__ t,
                                  const unsigned* __restrict__ maxnorm_bits, int64_t m, float scale,
                                  int logits_bf16, float* __restrict__ zref,
                                  const FusedDyn* __restrict__ dyn) {
  if (dyn) m = dyn->m;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float nmax = __uint_as_float(*maxnorm_bits);
  for (int64_t r = warp; r < m; r += nwarps) {
    float ss = 0.f;
    for (int c = lane; c < fk::D; c += 32) {
      const float v = __bfloat162float(q[r * fk::D + c]);
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    float tr = t[r];
    if (logits_bf16) tr = bf16_round(tr);
    float st = tr * scale;
    if (logits_bf16 && scale != 1.0f) st = bf16_round(st);
    if (lane == 0) zref[r] = fmaxf(st, fabsf(scale) * sqrtf(ss) * nmax * 1.01f);
  }
}

// ---- host side ------------------------------------------------------------------------------------
static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int make_tmap_bf16_rows(CUtensorMap* out, const void* base, int64_t rows, int64_t cols,
                        int64_t ld_elems, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return XR_E_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld_elems * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                         gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%lld", (int)r,
              (long long)rows, (long long)cols);
    return XR_E_CUDA;
  }
  return XR_OK;
}

// Loss kinds: stream-K (see the kernel).  Nothing to search for: the plan is the shape itself.
//   rb row blocks x nt candidate tiles = `tiles` units, split evenly over grid = min(#SMs, tiles) CTAs;
//   segment slots (partial sums) = grid + rb - 1.  The same integer arithmetic runs in the kernels.
struct FusedPlan {
  int rb, nt, grid, slots;
  long long tiles;
};
static inline FusedPlan make_plan(long long m, long long cn, int n_sm) {
  FusedPlan pl;
  pl.rb = (int)((m + fk::BM - 1) / fk::BM);
  pl.nt = (int)((cn + fk::BN - 1) / fk::BN);
  if (pl.rb < 0) pl.rb = 0;
  if (pl.nt < 0) pl.nt = 0;
  pl.tiles = (long long)pl.rb * pl.nt;
  pl.grid = (int)(pl.tiles < n_sm ? pl.tiles : n_sm);
  if (pl.grid < 1) pl.grid = 1;
  pl.slots = pl.grid + (pl.rb > 0 ? pl.rb - 1 : 0);
  return pl;
}
// upper bound of make_plan(m, cn).slots over all m <= m_max (any cn)
static long long max_plan_slots(long long m_max, int n_sm) {
  return (long long)n_sm + (m_max + fk::BM - 1) / fk::BM;
}

// retrieval scoring (KIND_GMAX / KIND_FILTER) writes no partial sums: (query block, split of the catalog
// tiles) work items, a few per SM whatever the number of query blocks
struct GmaxPlan {
  int rb, nt, spl, tps, n_items;
};
static inline GmaxPlan make_gmax_plan(long long u, long long n, int n_sm) {
  GmaxPlan pl;
  pl.rb = (int)((u + fk::BM - 1) / fk::BM);
  pl.nt = (int)((n + fk::BN - 1) / fk::BN);
  long long spl = ((long long)n_sm * 4 + pl.rb - 1) / pl.rb;
  if (spl > pl.nt) spl = pl.nt;
  if (spl < 1) spl = 1;
  pl.tps = (int)((pl.nt + spl - 1) / spl);
  pl.spl = (pl.nt + pl.tps - 1) / pl.tps;
  pl.n_items = pl.rb * pl.spl;
  return pl;
}

// device-side shape record for the sync-free step, from the row counts the compaction left on the
// device (M_a pool rows, M rows)
struct PlanHook {
  FusedDyn* dyn_main;
  FusedDyn* dyn_diag;
  __device__ void operator()(int m_a, int m) const {
    if (threadIdx.x == 0) {
      const int rb = m > 0 ? (m + fk::BM - 1) / fk::BM : 0, nt = m_a > 0 ? (m_a + fk::BN - 1) / fk::BN : 0;
      *dyn_main = FusedDyn{m, m_a, nt, 1, nt, 0, rb, 0};
      *dyn_diag = FusedDyn{m, m, 0, 1, 2, rb, rb, 0};
    }
  }
};
__global__ void fused_plan_kernel(const int64_t* __restrict__ counts, PlanHook hook) {
  hook((int)counts[0], (int)counts[1]);
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// optional per-launch timing of the main fused kernel (bench.py's roofline leg): a ring of
// CUDA event pairs recorded on the launching stream; nothing is synchronised until it is read.
bool g_ctrl_low = false;   // profiling aid (xr_fused_wait_stats bit 3)
static bool g_wait_stats = false;
static int g_ablate = 0;
static bool g_timeline = false;
static long long* g_dbg_dev = nullptr;
static long long g_dbg_host[64 * 8 + 64];   // per-tile stamps of CTA 0 + its kernel / item level stamps
static unsigned long long g_wait_host[16];
constexpr int kProfRing = 512;
static bool g_prof_on = false;
static cudaEvent_t g_prof_ev[kProfRing][2];
static bool g_prof_made = false;
static int g_prof_n = 0;

template <int KIND, bool RBF, int DBG = 0>
static int launch_fused1(const CUtensorMap& tq, const CUtensorMap& tb, const FusedParams& p,
                         int grid, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    XR_CUDA(cudaFuncSetAttribute(fused_pool_kernel<KIND, RBF, DBG>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, fk::SMEM_BYTES));
    configured = true;
  }
  FusedParams pp = p;
  pp.ctrl_low = g_ctrl_low;
  fused_pool_kernel<KIND, RBF, DBG><<<grid, fk::THREADS, fk::SMEM_BYTES, s>>>(tq, tb, pp);
  XR_LAUNCH_CHECK("fused_pool_kernel");
  return XR_OK;
}
// train kernel that also accumulates the monitoring sums of both logit families (dot-family train losses)
template <int KIND, bool RBF>
static int launch_fused_mon1(const CUtensorMap& tq, const CUtensorMap& tb, const FusedParams& p, int grid,
                             cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    XR_CUDA(cudaFuncSetAttribute(fused_pool_kernel<KIND, RBF, 0, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, fk::SMEM_BYTES));
    configured = true;
  }
  FusedParams pp = p;
  pp.ctrl_low = g_ctrl_low;
  fused_pool_kernel<KIND, RBF, 0, true><<<grid, fk::THREADS, fk::SMEM_BYTES, s>>>(tq, tb, pp);
  XR_LAUNCH_CHECK("fused_pool_kernel<MON>");
  return XR_OK;
}
template <int KIND>
static int launch_fused_mon(const CUtensorMap& tq, const CUtensorMap& tb, const FusedParams& p, int grid,
                            cudaStream_t s) {
  return p.logits_bf16 ? launch_fused_mon1<KIND, true>(tq, tb, p, grid, s)
                       : launch_fused_mon1<KIND, false>(tq, tb, p, grid, s);
}

template <int KIND>
static int launch_fused(const CUtensorMap& tq, const CUtensorMap& tb, const FusedParams& p,
                        int grid, cudaStream_t s) {
  return p.logits_bf16 ? launch_fused1<KIND, true>(tq, tb, p, grid, s)
                       : launch_fused1<KIND, false>(tq, tb, p, grid, s);
}

}  // namespace xr

using namespace xr;

namespace xr {
int launch_score_gmax2(const void* q, int64_t u, const void* catalog, int64_t n, int tile_stride, float* gmax,
                       int64_t ld, int* hang_flag, cudaStream_t s, int ablate);
int launch_score_filter2(const void* q, int64_t u, const void* catalog, int64_t n, const float* thresh,
                         int64_t thresh_stride, const FilterOut& fo, int* hang_flag, cudaStream_t s);
int gmax2_splits(int64_t u, int64_t n);
}
extern "C" int xr_fused_available(void) { return 3; }  // bit 0: fused loss, bit 1: fused retrieval scoring

// workspace carve-up shared by xr_fused_pool_loss (exact shape) and xr_pool_step (bounds)
struct FusedWs {
  float *t, *zref, *rl, *part_s, *part_o;
  int* flags;
  FusedDyn* dyn;   // [2]: main, diag (used by the step only)
  size_t bytes;
};
static FusedWs carve_fused_ws(void* workspace, long long m_max, long long max_items /* segment slots */) {
  FusedWs w;
  uint8_t* p = (uint8_t*)workspace;
  w.t = (float*)p;            p += align256((size_t)m_max * 4);
  w.zref = (float*)p;         p += align256((size_t)m_max * 4);
  w.rl = (float*)p;           p += align256((size_t)m_max * 4);
  w.part_s = (float*)p;       p += align256((size_t)max_items * fk::CG * fk::BM * fk::NSCAL * 4);
  w.part_o = (float*)p;       p += align256((size_t)max_items * fk::BM * fk::D * 4);
  w.flags = (int*)p;          p += 256;
  w.dyn = (FusedDyn*)p;       p += 256;
  w.bytes = (size_t)(p - (uint8_t*)workspace);
  return w;
}

extern "C" size_t xr_fused_pool_workspace_bytes(int64_t m, int64_t cn, int64_t dim) {
  if (dim != fk::D || m <= 0 || cn <= 0) return 512;
  const FusedPlan pl = make_plan(m, cn, sm_count_max());
  return carve_fused_ws(nullptr, m, pl.slots).bytes;
}

// The launch sequence of the fused loss: diagonal pass (target logits) -> [softmax reference
// bound] -> fused contraction/loss/gradient -> finalize -> row sum.  With `dyn` == nullptr the
// shape (m, cn) is exact; otherwise (m, cn) are upper bounds that size tensor maps and grids and
// the kernels read the real shape and plan from ws.dyn (written by fused_plan_kernel earlier on
// the same stream).
// one-pass monitoring (xr_pool_step_compute_mon): the train kernel also accumulates the sums behind every
// loss of both logit families and LogitsStatistics; these are the extra buffers and outputs
struct MonArgs {
  float *part_dot, *part_cos;     // [slots][CG][128][NSCAL_ALL] each
  float *inv_q, *inv_n, *t_cos;   // [n_pos], [n_pos + 64], [n_pos]
  double *row_out, *row_out2;     // [n_pos][ROW_SLOTS] per logit family
  double* scratch;                // 2 x kRowlossPartialBytes: the reductions' partials
  double *losses_dot, *losses_cos, *stats;
  long long n_pos;
};

// inverse row norms for the cosine family of a MON launch (losses.py:206-208: norms clamped at eps), the same
// summation order as xr_normalize_rows.  job 0: row r of q and pos -> inv_q[r], t_cos[r] = (t[r] inv_q) inv_pos;
// job 1: row r of neg -> inv_n[r].  Warp per row, rows up to the device-side counts.
__global__ void __launch_bounds__(256)
mon_inv_norms_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ pos,
                     const __nv_bfloat16* __restrict__ neg, const float* __restrict__ t, long long m, long long cn,
                     const FusedDyn* __restrict__ dyn, float eps, float* __restrict__ inv_q,
                     float* __restrict__ inv_n, float* __restrict__ t_cos) {
  if (dyn) {
    m = dyn->m;
    cn = dyn->cn;
  }
  const int job = blockIdx.y;
  const long long rows = job == 0 ? m : cn;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    if (job == 0) {
      const float iq = normalize_row384_bf16(q + r * fk::D, nullptr, eps, lane);
      const float ip = normalize_row384_bf16(pos + r * fk::D, nullptr, eps, lane);
      if (lane == 0) {
        inv_q[r] = iq;
        t_cos[r] = __fmul_rn(__fmul_rn(t[r], iq), ip);
      }
    } else {
      const float in = normalize_row384_bf16(neg + r * fk::D, nullptr, eps, lane);
      if (lane == 0) inv_n[r] = in;
    }
  }
}

struct StepScatter {   // xr_pool_step: write dL/dtok straight into the (n_pos, D) layout
  const int64_t* inv_pos;
  const int64_t* sel_pos;
  int64_t n_pos;
  void* dtok;
  int dtok_bf16;
};
static int fused_launch_all(const void* q, const void* pos, const void* neg, long long m,
                            long long cn, int loss_kind, const xr_loss_config* cfg,
                            const float* q_inv_norm, float grad_scale, float* dq, double* loss_out,
                            float* row_loss, const FusedWs& ws, bool dynamic, cudaStream_t s,
                            const StepScatter* sc = nullptr, const MonArgs* mon = nullptr) {
  const int n_sm = sm_count();
  const FusedPlan pl = make_plan(m, cn, n_sm);
  const FusedDyn* dyn_main = dynamic ? ws.dyn : nullptr;
  const FusedDyn* dyn_diag = dynamic ? ws.dyn + 1 : nullptr;
  XR_CUDA(cudaMemsetAsync(ws.flags, 0, 256, s));

  CUtensorMap tq, tp, tn;
  int rc;
  if ((rc = make_tmap_bf16_rows(&tq, q, m, fk::D, fk::D, fk::BM))) return rc;
  if ((rc = make_tmap_bf16_rows(&tp, pos, m, fk::D, fk::D, fk::BN))) return rc;
  if ((rc = make_tmap_bf16_rows(&tn, neg, cn, fk::D, fk::D, fk::BN))) return rc;

  // pass 1: target logits q_i . pos_i from the diagonal of Q_blk . Pos_blk^T, through the SAME
  // MMA instruction sequence as every pool logit, so a pool entry equal to the row's positive
  // ties exactly and the strict '<' of losses.py:292 masks it (as one bmm does in the reference)
  FusedParams pd{};
  pd.dyn = dyn_diag;
  pd.m = (int)m; pd.cn = (int)m; pd.nt_count = 0; pd.spl = 1; pd.tiles_per_split = 2;
  pd.n_items = pl.rb; pd.rb_count = pl.rb; pd.t_out = ws.t; pd.hang_flag = ws.flags;
  if ((rc = launch_fused<fk::KIND_DIAG>(tq, tp, pd, pl.rb < n_sm ? pl.rb : n_sm, s))) return rc;

  const float* zref = nullptr;
  if ((loss_kind == XR_LOSS_INFONCE || mon) && !cfg->mask_false_negatives) {   // (the monitored InfoNCE needs it too)
    unsigned* nmax = (unsigned*)(ws.flags + 8);
    negnorm_max_kernel<<<n_sm * 4, 256, 0, s>>>((const __nv_bfloat16*)neg, cn, nmax, dyn_main);
    XR_LAUNCH_CHECK("negnorm_max");
    zref_bound_kernel<<<n_sm * 4, 256, 0, s>>>((const __nv_bfloat16*)q, ws.t, nmax, m, cfg->scale,
                                               cfg->logits_bf16, ws.zref, dyn_main);
    XR_LAUNCH_CHECK("zref_bound");
    zref = ws.zref;
  }
  if (mon) {   // inverse norms + cosine targets (needs the diagonal pass above)
    long long gx = (m * 32 + 255) / 256;
    if (gx > (long long)n_sm * 4) gx = (long long)n_sm * 4;
    mon_inv_norms_kernel<<<dim3((unsigned)gx, 2), 256, 0, s>>>(
        (const __nv_bfloat16*)q, (const __nv_bfloat16*)pos, (const __nv_bfloat16*)neg, ws.t, m, cn, dyn_main, 1e-8f,
        mon->inv_q, mon->inv_n, mon->t_cos);
    XR_LAUNCH_CHECK("mon_inv_norms");
  }

  FusedParams p{};
  p.dyn = dyn_main;
  p.m = (int)m; p.cn = (int)cn; p.nt_count = pl.nt; p.rb_count = pl.rb;
  p.mask_fn = cfg->mask_false_negatives; p.logits_bf16 = cfg->logits_bf16;
  p.with_grad = dq != nullptr || (sc && sc->dtok); p.scale = cfg->scale; p.margin = cfg->margin;
  p.t = ws.t; p.zref = zref; p.part_o = ws.part_o; p.part_s = ws.part_s; p.hang_flag = ws.flags;
  if (mon) {
    p.part_all = mon->part_dot; p.part_all_cos = mon->part_cos;
    p.mon_inv_q = mon->inv_q; p.mon_inv_n = mon->inv_n; p.mon_t_cos = mon->t_cos;
  }
  // dynamic: the tile count is only known on the device; surplus CTAs get an empty range and exit.
  // The finalize kernels derive the segments of a row block from (m, cn, grid) exactly as the kernel does.
  const int grid = dynamic ? n_sm : pl.grid;
  const bool prof = g_prof_on && g_prof_n < kProfRing;
  if (prof) cudaEventRecord(g_prof_ev[g_prof_n][0], s);
  switch (loss_kind) {
    case XR_LOSS_INFONCE:
      if (mon) {
        rc = launch_fused_mon<XR_LOSS_INFONCE>(tq, tn, p, grid, s);
        break;
      }
      if (g_wait_stats || g_ablate || g_timeline) {   // profiling aid
        if (!g_dbg_dev) XR_CUDA(cudaMalloc(&g_dbg_dev, sizeof(g_dbg_host)));
        p.dbg = g_dbg_dev;
        p.ablate = g_ablate;
        XR_CUDA(cudaMemsetAsync(g_dbg_dev, 0, sizeof(g_dbg_host), s));
        if (g_wait_stats) rc = launch_fused1<XR_LOSS_INFONCE, true, 2>(tq, tn, p, grid, s);
        else rc = launch_fused1<XR_LOSS_INFONCE, true, 1>(tq, tn, p, grid, s);
      }
      else rc = launch_fused<XR_LOSS_INFONCE>(tq, tn, p, grid, s);
      break;
    case XR_LOSS_NCE:
      rc = mon ? launch_fused_mon<XR_LOSS_NCE>(tq, tn, p, grid, s) : launch_fused<XR_LOSS_NCE>(tq, tn, p, grid, s);
      break;
    case XR_LOSS_PAIRWISE_HINGE:
      rc = mon ? launch_fused_mon<XR_LOSS_PAIRWISE_HINGE>(tq, tn, p, grid, s)
               : launch_fused<XR_LOSS_PAIRWISE_HINGE>(tq, tn, p, grid, s);
      break;
    case XR_LOSS_PAIRWISE_LOGISTIC:
      rc = mon ? launch_fused_mon<XR_LOSS_PAIRWISE_LOGISTIC>(tq, tn, p, grid, s)
               : launch_fused<XR_LOSS_PAIRWISE_LOGISTIC>(tq, tn, p, grid, s);
      break;
    case XR_LOSS_CONTRASTIVE: rc = launch_fused<XR_LOSS_CONTRASTIVE>(tq, tn, p, grid, s); break;
    default: rc = launch_fused<XR_LOSS_ALIGNMENT_CONTRASTIVE>(tq, tn, p, grid, s); break;
  }
  if (prof) cudaEventRecord(g_prof_ev[g_prof_n++][1], s);
  if (rc) return rc;

  float* rl = row_loss ? row_loss : ws.rl;
  fused_finalize_kernel<<<n_sm * 4, 256, 0, s>>>(
      ws.part_o, ws.part_s, ws.t, zref, (const __nv_bfloat16*)q, (const __nv_bfloat16*)pos,
      q_inv_norm, (int)m, (int)cn, grid, loss_kind, cfg->logits_bf16, cfg->scale, cfg->margin, grad_scale,
      dq, rl, dyn_main, sc ? sc->inv_pos : nullptr, sc ? sc->sel_pos : nullptr, sc ? sc->n_pos : 0,
      sc ? sc->dtok : nullptr, sc ? sc->dtok_bf16 : 0);
  XR_LAUNCH_CHECK("fused_finalize");
  // loss_out[1] (if the caller left room) receives the fp32 copy the loss module returns
  sum_rows_kernel<<<1, 1024, 0, s>>>(rl, m, loss_out, reinterpret_cast<float*>(loss_out + 1), dyn_main);
  XR_LAUNCH_CHECK("sum_rows");
  if (mon) {
    // the monitoring sums of both families -> the row slots rowloss_kernel produces -> losses[7] / stats[16]
    const int fblocks = (int)((m + 255) / 256);
    fused_finalize_all_kernel<<<dim3(fblocks, 2), 256, 0, s>>>(mon->part_dot, ws.t, zref, (int)m, (int)cn, grid, 0,
                                                               cfg->logits_bf16, cfg->scale, cfg->margin, mon->row_out,
                                                               dyn_main, mon->part_cos, mon->t_cos, mon->row_out2);
    XR_LAUNCH_CHECK("fused_finalize_all");
    if ((rc = launch_rowloss_reduce2(mon->row_out, mon->row_out2, m, cn + 1, mon->losses_dot, mon->stats,
                                     mon->losses_cos, s, reinterpret_cast<const int*>(dyn_main), mon->scratch)))
      return rc;
  }
  if (g_wait_stats || g_timeline) {
    XR_CUDA(cudaMemcpyAsync(g_wait_host, ws.flags + 16, sizeof(g_wait_host), cudaMemcpyDeviceToHost, s));
    if (g_dbg_dev) XR_CUDA(cudaMemcpyAsync(g_dbg_host, g_dbg_dev, sizeof(g_dbg_host), cudaMemcpyDeviceToHost, s));
    XR_CUDA(cudaStreamSynchronize(s));
  }
  return XR_OK;
}

static int check_fused_device(const char* who) {
  int dev = 0, major = 0;
  XR_CUDA(cudaGetDevice(&dev));
  XR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("%s: needs an sm_100 device (tcgen05/TMEM), found sm_%d", who, major * 10);
    return XR_E_UNSUPPORTED;
  }
  return XR_OK;
}

static int check_fused_kind(const char* who, int loss_kind, const xr_loss_config* cfg) {
  const bool cosine = loss_kind == XR_LOSS_CONTRASTIVE || loss_kind == XR_LOSS_ALIGNMENT_CONTRASTIVE;
  XR_CHECK_ARG(cfg->num_hard_negatives == 0, "%s: hard-negative mining needs the materialised path", who);
  XR_CHECK_ARG(loss_kind == XR_LOSS_INFONCE || loss_kind == XR_LOSS_NCE ||
                   loss_kind == XR_LOSS_PAIRWISE_HINGE || loss_kind == XR_LOSS_PAIRWISE_LOGISTIC ||
                   cosine,
               "%s: unsupported loss kind %d", who, loss_kind);
  XR_CHECK_ARG(loss_kind != XR_LOSS_INFONCE || cfg->scale > 0.f, "%s: InfoNCE needs scale > 0", who);
  return XR_OK;
}

extern "C" int xr_fused_pool_loss(const void* q, const void* pos, const void* neg, int64_t m,
                                  int64_t cn, int64_t dim, int loss_kind,
                                  const xr_loss_config* cfg, const float* q_inv_norm,
                                  float grad_scale, float* dq, double* loss_out, float* row_loss,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  XR_CHECK_ARG(q && pos && neg && cfg && loss_out && workspace, "xr_fused_pool_loss: null pointer");
  XR_CHECK_ARG(dim == fk::D, "xr_fused_pool_loss: this build is specialised for dim = %d", fk::D);
  XR_CHECK_ARG(m > 0 && cn > 0 && m < (1ll << 30) && cn < (1ll << 30),
               "xr_fused_pool_loss: bad sizes");
  XR_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)pos % 16 == 0) && ((uintptr_t)neg % 16 == 0),
               "xr_fused_pool_loss: operands must be 16-byte aligned");
  int rc;
  if ((rc = check_fused_kind("xr_fused_pool_loss", loss_kind, cfg))) return rc;
  const bool cosine = loss_kind == XR_LOSS_CONTRASTIVE || loss_kind == XR_LOSS_ALIGNMENT_CONTRASTIVE;
  XR_CHECK_ARG(!cosine || !dq || q_inv_norm, "xr_fused_pool_loss: cosine kinds need q_inv_norm");
  XR_CHECK_ARG(workspace_bytes >= xr_fused_pool_workspace_bytes(m, cn, dim),
               "xr_fused_pool_loss: workspace too small");
  if ((rc = check_fused_device("xr_fused_pool_loss"))) return rc;
  const FusedPlan pl = make_plan(m, cn, sm_count());
  const FusedWs ws = carve_fused_ws(workspace, m, pl.slots);
  return fused_launch_all(q, pos, neg, m, cn, loss_kind, cfg, q_inv_norm, grad_scale, dq, loss_out,
                          row_loss, ws, false, as_stream(stream));
}

// ---- train loss + gradient + the monitoring of BOTH logit families from one pass (exact shapes) ------------
// The module-path form of xr_pool_step_compute_mon: what trainer.py:213-264 (compute_losses) needs for one
// batch -- InfoNCE loss and dL/dq, losses_dot[7], losses_cos[7], stats[16] -- from ONE tensor-core pass over
// the pool.  Workspace: xr_fused_pool_loss_mon_workspace_bytes(m, cn, dim).
struct MonWs {
  float *part_dot, *part_cos, *inv_q, *inv_n, *t_cos;
  double *row_out, *row_out2, *scratch;
  size_t bytes;
};
static MonWs carve_mon_ws(void* base, long long m, long long cn, long long slots) {
  MonWs w;
  uint8_t* p = (uint8_t*)base;
  const size_t part = align256((size_t)slots * fk::CG * fk::BM * fk::NSCAL_ALL * 4);
  w.part_dot = (float*)p;   p += part;
  w.part_cos = (float*)p;   p += part;
  w.inv_q = (float*)p;      p += align256((size_t)m * 4);
  w.inv_n = (float*)p;      p += align256((size_t)(cn + 64) * 4);
  w.t_cos = (float*)p;      p += align256((size_t)m * 4);
  w.row_out = (double*)p;   p += align256((size_t)m * ROW_SLOTS * 8);
  w.row_out2 = (double*)p;  p += align256((size_t)m * ROW_SLOTS * 8);
  w.scratch = (double*)p;   p += 2 * kRowlossPartialBytes;
  w.bytes = (size_t)(p - (uint8_t*)base);
  return w;
}

extern "C" size_t xr_fused_pool_loss_mon_workspace_bytes(int64_t m, int64_t cn, int64_t dim) {
  if (dim != fk::D || m <= 0 || cn <= 0) return 512;
  const FusedPlan pl = make_plan(m, cn, sm_count_max());
  return carve_fused_ws(nullptr, m, pl.slots).bytes + carve_mon_ws(nullptr, m, cn, pl.slots).bytes;
}

extern "C" int xr_fused_pool_loss_mon(const void* q, const void* pos, const void* neg, int64_t m, int64_t cn,
                                      int64_t dim, int loss_kind, const xr_loss_config* cfg, float grad_scale,
                                      float* dq, double* loss_out, double* losses_dot, double* losses_cos,
                                      double* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
  XR_CHECK_ARG(q && pos && neg && cfg && loss_out && losses_dot && losses_cos && stats_out && workspace,
               "xr_fused_pool_loss_mon: null pointer");
  XR_CHECK_ARG(dim == fk::D, "xr_fused_pool_loss_mon: this build is specialised for dim = %d", fk::D);
  XR_CHECK_ARG(m > 0 && cn > 0 && m < (1ll << 30) && cn < (1ll << 30), "xr_fused_pool_loss_mon: bad sizes");
  XR_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)pos % 16 == 0) && ((uintptr_t)neg % 16 == 0) &&
                   (uintptr_t)workspace % 256 == 0,
               "xr_fused_pool_loss_mon: operands must be 16-byte aligned, the workspace 256-byte aligned");
  XR_CHECK_ARG(loss_kind == XR_LOSS_INFONCE || loss_kind == XR_LOSS_NCE || loss_kind == XR_LOSS_PAIRWISE_HINGE ||
                   loss_kind == XR_LOSS_PAIRWISE_LOGISTIC,
               "xr_fused_pool_loss_mon: the train loss must be of the dot family");
  XR_CHECK_ARG(cfg->num_hard_negatives == 0 && cfg->scale > 0.f,
               "xr_fused_pool_loss_mon: needs scale > 0 and no hard-negative mining");
  XR_CHECK_ARG(workspace_bytes >= xr_fused_pool_loss_mon_workspace_bytes(m, cn, dim),
               "xr_fused_pool_loss_mon: workspace too small");
  int rc;
  if ((rc = check_fused_device("xr_fused_pool_loss_mon"))) return rc;
  const FusedPlan pl = make_plan(m, cn, sm_count_max());
  const FusedWs ws = carve_fused_ws(workspace, m, pl.slots);
  const MonWs mw = carve_mon_ws((uint8_t*)workspace + ws.bytes, m, cn, pl.slots);
  const MonArgs mon{mw.part_dot, mw.part_cos, mw.inv_q, mw.inv_n, mw.t_cos, mw.row_out, mw.row_out2, mw.scratch,
                    losses_dot, losses_cos, stats_out, m};
  return fused_launch_all(q, pos, neg, m, cn, loss_kind, cfg, nullptr, grad_scale, dq, loss_out, nullptr, ws,
                          false, as_stream(stream), nullptr, &mon);
}

// ---- every loss of one logit family + LogitsStatistics in ONE pass over the pool ----------------
// trainer.py:250-263 evaluates LogitsStatistics and all seven losses on every training step (eight
// logit computations in the reference).  Forward only: diagonal pass -> [softmax reference bound]
// -> fused_pool_kernel<KIND_ALL_*> -> fused_finalize_all_kernel -> rowloss_reduce_kernel, filling
// the same losses[7] / stats[16] blocks as xr_rowloss does from materialised logits.
// the per-item scalars (CG x 128 x NSCAL_ALL floats) alias the partial-dQ region (128 x 384 floats),
// which the forward-only ALL kinds never write; the row slots follow the fused workspace
static_assert(fk::CG * fk::NSCAL_ALL <= fk::D, "ALL-kind scalars must fit the partial-dQ region");
static size_t fused_all_extra_bytes(long long m) { return align256((size_t)m * ROW_SLOTS * 8) + kRowlossPartialBytes; }

extern "C" size_t xr_fused_pool_all_workspace_bytes(int64_t m, int64_t cn, int64_t dim) {
  if (dim != fk::D || m <= 0 || cn <= 0) return 512;
  const FusedPlan pl = make_plan(m, cn, sm_count_max());
  return carve_fused_ws(nullptr, m, pl.slots).bytes + fused_all_extra_bytes(m);
}

// (m, cn) exact, or -- `dynamic` -- upper bounds with the real shape and plan in ws.dyn (written by
// fused_plan_kernel earlier on the stream), exactly as fused_launch_all does for the train loss
static int fused_all_launch(const void* q, const void* pos, const void* neg, long long m, long long cn,
                            int cosine, const xr_loss_config* cfg, const FusedWs& ws, bool dynamic,
                            double* row_out, double* losses_out, double* stats_out, cudaStream_t s) {
  const int n_sm = sm_count();
  const FusedPlan pl = make_plan(m, cn, n_sm);
  const FusedDyn* dyn_main = dynamic ? ws.dyn : nullptr;
  const FusedDyn* dyn_diag = dynamic ? ws.dyn + 1 : nullptr;
  XR_CUDA(cudaMemsetAsync(ws.flags, 0, 256, s));
  int rc;
  CUtensorMap tq, tp, tn;
  if ((rc = make_tmap_bf16_rows(&tq, q, m, fk::D, fk::D, fk::BM))) return rc;
  if ((rc = make_tmap_bf16_rows(&tp, pos, m, fk::D, fk::D, fk::BN))) return rc;
  if ((rc = make_tmap_bf16_rows(&tn, neg, cn, fk::D, fk::D, fk::BN))) return rc;
  FusedParams pd{};
  pd.dyn = dyn_diag;
  pd.m = (int)m; pd.cn = (int)m; pd.nt_count = 0; pd.spl = 1; pd.tiles_per_split = 2;
  pd.n_items = pl.rb; pd.rb_count = pl.rb; pd.t_out = ws.t; pd.hang_flag = ws.flags;
  if ((rc = launch_fused<fk::KIND_DIAG>(tq, tp, pd, pl.rb < n_sm ? pl.rb : n_sm, s))) return rc;
  const float* zref = nullptr;
  if (!cosine && !cfg->mask_false_negatives) {
    unsigned* nmax = (unsigned*)(ws.flags + 8);
    negnorm_max_kernel<<<n_sm * 4, 256, 0, s>>>((const __nv_bfloat16*)neg, cn, nmax, dyn_main);
    XR_LAUNCH_CHECK("negnorm_max");
    zref_bound_kernel<<<n_sm * 4, 256, 0, s>>>((const __nv_bfloat16*)q, ws.t, nmax, m, cfg->scale,
                                               cfg->logits_bf16, ws.zref, dyn_main);
    XR_LAUNCH_CHECK("zref_bound");
    zref = ws.zref;
  }
  FusedParams p{};
  p.dyn = dyn_main;
  p.m = (int)m; p.cn = (int)cn; p.nt_count = pl.nt; p.rb_count = pl.rb;
  p.mask_fn = cfg->mask_false_negatives; p.logits_bf16 = cfg->logits_bf16;
  p.with_grad = 0; p.scale = cfg->scale; p.margin = cfg->margin;
  p.t = ws.t; p.zref = zref; p.part_all = ws.part_o; p.hang_flag = ws.flags;
  const int grid = dynamic ? n_sm : pl.grid;
  const bool prof = g_prof_on && g_prof_n < kProfRing && !dynamic;
  if (prof) cudaEventRecord(g_prof_ev[g_prof_n][0], s);
  rc = cosine ? launch_fused<fk::KIND_ALL_COS>(tq, tn, p, grid, s)
              : launch_fused<fk::KIND_ALL_DOT>(tq, tn, p, grid, s);
  if (prof) cudaEventRecord(g_prof_ev[g_prof_n++][1], s);
  if (rc) return rc;
  fused_finalize_all_kernel<<<(int)((m + 255) / 256), 256, 0, s>>>(
      ws.part_o, ws.t, zref, (int)m, (int)cn, grid, cosine, cfg->logits_bf16, cfg->scale, cfg->margin, row_out,
      dyn_main);
  XR_LAUNCH_CHECK("fused_finalize_all");
  // (row_out spans the upper bound m rows; the reduction's scratch follows it)
  return launch_rowloss_reduce(row_out, m, cn + 1, 0, losses_out, stats_out, s,
                               reinterpret_cast<const int*>(dyn_main),
                               reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(row_out) + align256((size_t)m * ROW_SLOTS * 8)));
}

extern "C" int xr_fused_pool_all(const void* q, const void* pos, const void* neg, int64_t m,
                                 int64_t cn, int64_t dim, int cosine, const xr_loss_config* cfg,
                                 double* losses_out, double* stats_out, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  XR_CHECK_ARG(q && pos && neg && cfg && losses_out && workspace, "xr_fused_pool_all: null pointer");
  XR_CHECK_ARG(dim == fk::D, "xr_fused_pool_all: this build is specialised for dim = %d", fk::D);
  XR_CHECK_ARG(m > 0 && cn > 0 && m < (1ll << 30) && cn < (1ll << 30), "xr_fused_pool_all: bad sizes");
  XR_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)pos % 16 == 0) && ((uintptr_t)neg % 16 == 0),
               "xr_fused_pool_all: operands must be 16-byte aligned");
  XR_CHECK_ARG(cfg->num_hard_negatives == 0, "xr_fused_pool_all: hard-negative mining needs the materialised path");
  XR_CHECK_ARG(cosine || cfg->scale > 0.f, "xr_fused_pool_all: InfoNCE needs scale > 0");
  XR_CHECK_ARG(workspace_bytes >= xr_fused_pool_all_workspace_bytes(m, cn, dim),
               "xr_fused_pool_all: workspace too small");
  int rc;
  if ((rc = check_fused_device("xr_fused_pool_all"))) return rc;
  const FusedWs ws = carve_fused_ws(workspace, m, make_plan(m, cn, sm_count_max()).slots);
  double* row_out = (double*)((uint8_t*)workspace + ws.bytes);
  return fused_all_launch(q, pos, neg, m, cn, cosine, cfg, ws, false, row_out, losses_out, stats_out,
                          as_stream(stream));
}

// ---- the whole scoring-and-loss step, sync-free -------------------------------------------------
// compute_embeds (models.py:388-416) + EmbedLoss.forward (losses.py:128-155) + the backward to the
// encoder output, for one SeqBatch of n_pos = B*L positions, without a single device->host copy:
// the row counts stay on the device (fused_plan_kernel plans the tensor-core kernel there), every
// buffer is sized by n_pos, so the call sequence is static and CUDA-graph capturable.
__global__ void __launch_bounds__(256)
step_gather_kernel(const char* __restrict__ tok, int tok_f32, const char* __restrict__ table,
                   int64_t n_table_rows, const int64_t* __restrict__ pos_idx,
                   const int64_t* __restrict__ neg_idx, const int64_t* __restrict__ sel_attn,
                   const int64_t* __restrict__ sel_pos, const int64_t* __restrict__ counts,
                   int64_t n_pos, char* __restrict__ q_out, char* __restrict__ pos_out,
                   char* __restrict__ neg_out, int32_t* __restrict__ err_flag) {
  // job 0: q = bf16(tok[sel_pos])   job 1: pos = table[pos_idx[sel_pos]]   job 2: neg = table[neg_idx[sel_attn]]
  // rows [count, round_up(count, 128)) are zero-filled: the tensor maps cover n_pos rows, and a
  // stale row inside the last tile would reach the gradient MMA as 0 x garbage
  constexpr int VPR = fk::D * 2 / 16;   // 16-byte vectors per bf16 row
  const int job = blockIdx.y;
  const int64_t cnt = job == 2 ? counts[0] : counts[1];
  int64_t padded = (cnt + 127) / 128 * 128;
  if (padded > n_pos) padded = n_pos;
  const int64_t* sel = job == 2 ? sel_attn : sel_pos;
  const int64_t* idx = job == 1 ? pos_idx : neg_idx;
  char* out = job == 0 ? q_out : (job == 1 ? pos_out : neg_out);
  const int64_t total = padded * VPR;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int64_t r = v / VPR;
    const int c = (int)(v - r * VPR);
    int4 o = make_int4(0, 0, 0, 0);
    if (r < cnt) {
      const int64_t position = sel[r];
      if (job == 0) {
        if (tok_f32) {
          const int4 a = ld_stream16(reinterpret_cast<const int4*>(tok + position * (fk::D * 4)) + 2 * c);
          const int4 b = ld_stream16(reinterpret_cast<const int4*>(tok + position * (fk::D * 4)) + 2 * c + 1);
          const float* fa = reinterpret_cast<const float*>(&a);
          const float* fb = reinterpret_cast<const float*>(&b);
          __nv_bfloat162 h0 = __floats2bfloat162_rn(fa[0], fa[1]), h1 = __floats2bfloat162_rn(fa[2], fa[3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(fb[0], fb[1]), h3 = __floats2bfloat162_rn(fb[2], fb[3]);
          o.x = *reinterpret_cast<int*>(&h0); o.y = *reinterpret_cast<int*>(&h1);
          o.z = *reinterpret_cast<int*>(&h2); o.w = *reinterpret_cast<int*>(&h3);
        } else {
          o = ld_stream16(reinterpret_cast<const int4*>(tok + position * (fk::D * 2)) + c);
        }
      } else {
        const int64_t src = idx[position];
        if (src < 0 || src >= n_table_rows) {
          if (err_flag) *err_flag = 1;
        } else {
          o = __ldg(reinterpret_cast<const int4*>(table + src * (fk::D * 2)) + c);
        }
      }
    }
    st_stream16(reinterpret_cast<int4*>(out) + v, o);
  }
}

// row-normalised bf16 copies of the step's three operand buffers in ONE launch (losses.py:206-208: both
// sides normalised, norms clamped at eps): job 0 q -> qn (+ 1/||q|| for the chain rule), job 1 pos -> pn,
// job 2 neg -> nn.  Warp per row; only the rows the tensor maps can reach are touched (up to the next
// multiple of 128 past the device-side row counts).
__global__ void __launch_bounds__(256)
step_normalize3_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ pos,
                       const __nv_bfloat16* __restrict__ neg, const FusedDyn* __restrict__ dyn,
                       int64_t n_pos, float eps, __nv_bfloat16* __restrict__ qn,
                       __nv_bfloat16* __restrict__ pn, __nv_bfloat16* __restrict__ nn,
                       float* __restrict__ inv_q) {
  const int job = blockIdx.y;
  const __nv_bfloat16* x = job == 0 ? q : (job == 1 ? pos : neg);
  __nv_bfloat16* y = job == 0 ? qn : (job == 1 ? pn : nn);
  int64_t rows = ((int64_t)(job == 2 ? dyn->cn : dyn->m) + 127) / 128 * 128;
  if (rows > n_pos) rows = n_pos;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float inv = normalize_row384_bf16(x + r * fk::D, y + r * fk::D, eps, lane);
    if (job == 0 && lane == 0) inv_q[r] = inv;
  }
}
static int launch_step_normalize3(const __nv_bfloat16* q, const __nv_bfloat16* pos, const __nv_bfloat16* neg,
                                  const FusedDyn* dyn, int64_t n_pos, __nv_bfloat16* qn, __nv_bfloat16* pn,
                                  __nv_bfloat16* nn, float* inv_q, cudaStream_t s) {
  int64_t gx = (n_pos * 32 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 4;
  if (gx > cap) gx = cap;
  step_normalize3_kernel<<<dim3((unsigned)gx, 3), 256, 0, s>>>(q, pos, neg, dyn, n_pos, 1e-8f, qn, pn, nn, inv_q);
  XR_LAUNCH_CHECK("step_normalize3");
  return XR_OK;
}

struct StepWs {
  int64_t *sel_attn, *sel_pos, *inv_pos;
  uint8_t *attn, *pos_mask;
  void* compact_ws;
  __nv_bfloat16 *q, *pos, *neg;
  int32_t* err;
  FusedWs fused;
  size_t bytes;
};
static StepWs carve_step_ws(void* workspace, long long n_pos) {
  StepWs w;
  uint8_t* p = (uint8_t*)workspace;
  const size_t n = (size_t)n_pos;
  w.sel_attn = (int64_t*)p;   p += align256(n * 8);
  w.sel_pos = (int64_t*)p;    p += align256(n * 8);
  w.inv_pos = (int64_t*)p;    p += align256(n * 8);
  w.attn = p;                 p += align256(n);
  w.pos_mask = p;             p += align256(n);
  w.compact_ws = p;           p += align256(xr_compact_workspace_bytes(n_pos));
  w.q = (__nv_bfloat16*)p;    p += align256(n * fk::D * 2);
  w.pos = (__nv_bfloat16*)p;  p += align256(n * fk::D * 2);
  w.neg = (__nv_bfloat16*)p;  p += align256(n * fk::D * 2);
  w.err = (int32_t*)p;        p += 256;
  const size_t off = (size_t)(p - (uint8_t*)workspace);
  w.fused = carve_fused_ws(workspace ? p : nullptr, n_pos, max_plan_slots(n_pos, sm_count_max()));
  w.bytes = off + w.fused.bytes;
  return w;
}

extern "C" size_t xr_pool_step_workspace_bytes(int64_t n_pos, int64_t dim) {
  if (dim != fk::D || n_pos <= 0) return 512;
  return carve_step_ws(nullptr, n_pos).bytes;
}

// extra buffers of the cosine kinds and of the monitoring pass: row-normalised operand copies
struct MonitorWs {
  __nv_bfloat16 *qn, *pn, *nn;
  float *inv, *inv_q;   // inv: scratch for pos / neg norms; inv_q: 1/||q|| kept for the cosine chain rule
  double* row_out;
  float *part_dot, *part_cos, *inv_n, *t_cos;   // one-pass monitoring (xr_pool_step_compute_mon)
  double *row_out2, *scratch2;
  size_t bytes;
};
static MonitorWs carve_monitor_ws(void* base, long long n_pos) {
  MonitorWs w;
  uint8_t* p = (uint8_t*)base;
  const size_t n = (size_t)n_pos;
  w.qn = (__nv_bfloat16*)p;   p += align256(n * fk::D * 2);
  w.pn = (__nv_bfloat16*)p;   p += align256(n * fk::D * 2);
  w.nn = (__nv_bfloat16*)p;   p += align256(n * fk::D * 2);
  w.inv = (float*)p;          p += align256(n * 4);
  w.inv_q = (float*)p;        p += align256(n * 4);
  w.row_out = (double*)p;     p += align256(n * ROW_SLOTS * 8) + kRowlossPartialBytes;
  const size_t part = align256((size_t)max_plan_slots(n_pos, sm_count_max()) * fk::CG * fk::BM * fk::NSCAL_ALL * 4);
  w.part_dot = (float*)p;     p += part;
  w.part_cos = (float*)p;     p += part;
  w.inv_n = (float*)p;        p += align256((n + 64) * 4);
  w.t_cos = (float*)p;        p += align256(n * 4);
  w.row_out2 = (double*)p;    p += align256(n * ROW_SLOTS * 8);
  w.scratch2 = (double*)p;    p += 2 * kRowlossPartialBytes;
  w.bytes = (size_t)(p - (uint8_t*)base);
  return w;
}

static size_t step_ws_bytes_with_normalised(long long n_pos) {
  return carve_step_ws(nullptr, n_pos).bytes + carve_monitor_ws(nullptr, n_pos).bytes;
}

// The step in two phases, so that a caller with two alternating step objects can overlap the INGEST
// of batch i+1 (index compaction, plan, the three gathers -- the only part that touches the batch's
// inputs) with the COMPUTE of batch i on another stream.  `tok` may be pinned HOST memory (UVA): the
// gather then pulls only the M selected rows over PCIe ("zero-copy": M x D x 2 bytes instead of a
// B x L x D x 2 byte copy of the whole encoder output).
extern "C" int xr_pool_step_ingest(const int64_t* history_idx, const int64_t* pos_idx,
                                   const int64_t* neg_idx, int64_t n_pos, const void* tok,
                                   int tok_dtype, const void* table_bf16, const uint8_t* rownz,
                                   int64_t n_table_rows, int64_t dim, int64_t* counts,
                                   int32_t* err_flag, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  XR_CHECK_ARG(history_idx && pos_idx && neg_idx && tok && table_bf16 && counts && workspace,
               "xr_pool_step_ingest: null pointer");
  XR_CHECK_ARG(dim == fk::D, "xr_pool_step_ingest: this build is specialised for dim = %d", fk::D);
  XR_CHECK_ARG(n_pos > 0 && n_pos < (1ll << 30) && n_table_rows > 0, "xr_pool_step_ingest: bad sizes");
  XR_CHECK_ARG(tok_dtype == XR_F32 || tok_dtype == XR_BF16, "xr_pool_step_ingest: bad dtype");
  XR_CHECK_ARG(((uintptr_t)tok % 16 == 0) && ((uintptr_t)table_bf16 % 16 == 0) &&
                   ((uintptr_t)workspace % 256 == 0),
               "xr_pool_step_ingest: buffers must be 16-byte aligned (workspace 256)");
  XR_CHECK_ARG(workspace_bytes >= xr_pool_step_workspace_bytes(n_pos, dim),
               "xr_pool_step_ingest: workspace too small");
  int rc;
  if ((rc = check_fused_device("xr_pool_step_ingest"))) return rc;
  cudaStream_t s = as_stream(stream);
  const StepWs w = carve_step_ws(workspace, n_pos);
  const int n_sm = sm_count();
  // 1 + 2. positions -> row lists + counts (models.py:343, 390, 398, 404, 413-416), then the plan
  //        of the tensor-core kernel for the real (M, M_a): 64 threads, all on the device
  if ((rc = xr_compact_positions(history_idx, pos_idx, rownz, n_table_rows, n_pos, w.attn, w.sel_attn,
                                 w.sel_pos, w.pos_mask, w.inv_pos, counts, w.compact_ws, stream)))
    return rc;
  fused_plan_kernel<<<1, 32, 0, s>>>(counts, PlanHook{w.fused.dyn, w.fused.dyn + 1});
  XR_LAUNCH_CHECK("fused_plan");
  // 3. the three gathers in one launch (models.py:392+415, :400+416, :406), bf16 operands
  const int64_t per_job = (n_pos * (fk::D * 2 / 16) + 255) / 256;
  int gx = (int)(per_job < (int64_t)n_sm * 4 ? per_job : (int64_t)n_sm * 4);
  if (gx < 1) gx = 1;
  step_gather_kernel<<<dim3(gx, 3), 256, 0, s>>>(
      (const char*)tok, tok_dtype == XR_F32, (const char*)table_bf16, n_table_rows, pos_idx, neg_idx,
      w.sel_attn, w.sel_pos, counts, n_pos, (char*)w.q, (char*)w.pos, (char*)w.neg, err_flag);
  XR_LAUNCH_CHECK("step_gather");
  return XR_OK;
}

extern "C" int xr_pool_step_compute(int64_t n_pos, int64_t dim, int loss_kind, const xr_loss_config* cfg,
                                    float grad_scale, void* dtok, int dtok_dtype, double* loss_out,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  XR_CHECK_ARG(cfg && loss_out && workspace, "xr_pool_step_compute: null pointer");
  XR_CHECK_ARG(dim == fk::D && n_pos > 0 && n_pos < (1ll << 30), "xr_pool_step_compute: bad sizes");
  XR_CHECK_ARG(!dtok || dtok_dtype == XR_F32 || dtok_dtype == XR_BF16, "xr_pool_step_compute: bad dtype");
  XR_CHECK_ARG(((uintptr_t)dtok % 16 == 0) && ((uintptr_t)workspace % 256 == 0),
               "xr_pool_step_compute: buffers must be 16-byte aligned (workspace 256)");
  int rc;
  if ((rc = check_fused_kind("xr_pool_step_compute", loss_kind, cfg))) return rc;
  const bool cosine = loss_kind == XR_LOSS_CONTRASTIVE || loss_kind == XR_LOSS_ALIGNMENT_CONTRASTIVE;
  XR_CHECK_ARG(workspace_bytes >= (cosine ? step_ws_bytes_with_normalised(n_pos)
                                          : xr_pool_step_workspace_bytes(n_pos, dim)),
               "xr_pool_step_compute: workspace too small (the cosine kinds need "
               "xr_pool_step_monitor_workspace_bytes)");
  if ((rc = check_fused_device("xr_pool_step_compute"))) return rc;
  const StepWs w = carve_step_ws(workspace, n_pos);
  // 4. fused contraction + loss + dL/dtok: the finalize kernel writes the gradient in the encoder
  //    output's layout (zero rows for unselected positions: autograd of
  //    token_embeddings[mask][pos_mask], models.py:392, 415)
  const StepScatter sc{w.inv_pos, w.sel_pos, n_pos, dtok, dtok_dtype == XR_BF16};
  if (!cosine)
    return fused_launch_all(w.q, w.pos, w.neg, n_pos, n_pos, loss_kind, cfg, nullptr, grad_scale, nullptr,
                            loss_out, nullptr, w.fused, true, as_stream(stream), dtok ? &sc : nullptr);
  // cosine kinds (CCL, losses.py:206-208, 338-372): row-normalised copies of the three operands (all
  // n_pos rows: static shapes), 1/||q|| kept for the chain rule through the query normalisation
  const MonitorWs mw = carve_monitor_ws((uint8_t*)workspace + w.bytes, n_pos);
  if ((rc = launch_step_normalize3(w.q, w.pos, w.neg, w.fused.dyn, n_pos, mw.qn, mw.pn, mw.nn, mw.inv_q,
                                   as_stream(stream))))
    return rc;
  xr_loss_config ccfg = *cfg;
  ccfg.logits_bf16 = 0;   // cosine logits stay fp32 under autocast
  return fused_launch_all(mw.qn, mw.pn, mw.nn, n_pos, n_pos, loss_kind, &ccfg, mw.inv_q, grad_scale, nullptr,
                          loss_out, nullptr, w.fused, true, as_stream(stream), dtok ? &sc : nullptr);
}

extern "C" int xr_pool_step(const int64_t* history_idx, const int64_t* pos_idx,
                            const int64_t* neg_idx, int64_t n_pos, const void* tok, int tok_dtype,
                            const void* table_bf16, const uint8_t* rownz, int64_t n_table_rows,
                            int64_t dim, int loss_kind, const xr_loss_config* cfg, float grad_scale,
                            void* dtok, int dtok_dtype, double* loss_out, int64_t* counts,
                            int32_t* err_flag, void* workspace, size_t workspace_bytes,
                            void* stream) {
  XR_CHECK_ARG(cfg && loss_out, "xr_pool_step: null pointer");
  int rc;
  if ((rc = check_fused_kind("xr_pool_step", loss_kind, cfg))) return rc;   // before any launch
  {
    const bool cosine = loss_kind == XR_LOSS_CONTRASTIVE || loss_kind == XR_LOSS_ALIGNMENT_CONTRASTIVE;
    XR_CHECK_ARG(!cosine || (dim == fk::D && n_pos > 0 && workspace_bytes >= step_ws_bytes_with_normalised(n_pos)),
                 "xr_pool_step: the cosine kinds need a workspace of xr_pool_step_monitor_workspace_bytes");
  }
  if ((rc = xr_pool_step_ingest(history_idx, pos_idx, neg_idx, n_pos, tok, tok_dtype, table_bf16, rownz,
                                n_table_rows, dim, counts, err_flag, workspace, workspace_bytes, stream)))
    return rc;
  return xr_pool_step_compute(n_pos, dim, loss_kind, cfg, grad_scale, dtok, dtok_dtype, loss_out,
                              workspace, workspace_bytes, stream);
}

// ---- the monitoring half of compute_losses inside the same sync-free sequence --------------------
// trainer.py:250-263 logs LogitsStatistics and all seven losses every step.  Called right after
// xr_pool_step on the SAME stream and workspace, this runs both all-losses passes on the operands the
// step gathered (dot: q / pos / neg as they are; cosine: their row-normalised copies,
// losses.py:206-208), still without a device->host copy: CUDA-graph capturable together with the step.
extern "C" size_t xr_pool_step_monitor_workspace_bytes(int64_t n_pos, int64_t dim) {
  if (dim != fk::D || n_pos <= 0) return 512;
  return step_ws_bytes_with_normalised(n_pos);
}

extern "C" int xr_pool_step_monitor(int64_t n_pos, int64_t dim, const xr_loss_config* cfg,
                                    double* losses_dot, double* losses_cos, double* stats_out,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  XR_CHECK_ARG(cfg && losses_dot && losses_cos && stats_out && workspace, "xr_pool_step_monitor: null pointer");
  XR_CHECK_ARG(dim == fk::D && n_pos > 0 && n_pos < (1ll << 30), "xr_pool_step_monitor: bad sizes");
  XR_CHECK_ARG(cfg->num_hard_negatives == 0 && cfg->scale > 0.f,
               "xr_pool_step_monitor: needs num_hard_negatives == 0 and scale > 0");
  XR_CHECK_ARG(workspace_bytes >= xr_pool_step_monitor_workspace_bytes(n_pos, dim) &&
                   (uintptr_t)workspace % 256 == 0,
               "xr_pool_step_monitor: workspace too small or misaligned");
  int rc;
  if ((rc = check_fused_device("xr_pool_step_monitor"))) return rc;
  cudaStream_t s = as_stream(stream);
  const StepWs w = carve_step_ws(workspace, n_pos);
  const MonitorWs mw = carve_monitor_ws((uint8_t*)workspace + w.bytes, n_pos);
  // dot family + statistics on the gathered operands
  if ((rc = fused_all_launch(w.q, w.pos, w.neg, n_pos, n_pos, 0, cfg, w.fused, true, mw.row_out,
                             losses_dot, stats_out, s)))
    return rc;
  // cosine family on the row-normalised copies (all n_pos rows: the buffers are sized by n_pos and
  // rows past the counts are masked by the kernels)
  if ((rc = launch_step_normalize3(w.q, w.pos, w.neg, w.fused.dyn, n_pos, mw.qn, mw.pn, mw.nn, mw.inv, s))) return rc;
  xr_loss_config ccfg = *cfg;
  ccfg.logits_bf16 = 0;   // cosine logits stay fp32 under autocast
  return fused_all_launch(mw.qn, mw.pn, mw.nn, n_pos, n_pos, 1, &ccfg, w.fused, true, mw.row_out,
                          losses_cos, nullptr, s);
}

// The compute phase WITH the monitoring folded into the train kernel: one tensor-core pass produces the train
// loss, dL/dtok and the sums behind all seven losses + LogitsStatistics (the dot family from the scores, the
// cosine family from the same scores times fp32 inverse norms) -- instead of the train pass plus the two
// all-losses passes of xr_pool_step_monitor.  InfoNCE train loss only (the trainer's default); the dot-family
// numbers are bit-identical to xr_pool_step_monitor's, the cosine family agrees to bf16 tolerance (it no
// longer rounds the normalised operands to bf16).
extern "C" int xr_pool_step_compute_mon(int64_t n_pos, int64_t dim, int loss_kind, const xr_loss_config* cfg,
                                        float grad_scale, void* dtok, int dtok_dtype, double* loss_out,
                                        double* losses_dot, double* losses_cos, double* stats_out,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  XR_CHECK_ARG(cfg && loss_out && losses_dot && losses_cos && stats_out && workspace,
               "xr_pool_step_compute_mon: null pointer");
  XR_CHECK_ARG(dim == fk::D && n_pos > 0 && n_pos < (1ll << 30), "xr_pool_step_compute_mon: bad sizes");
  XR_CHECK_ARG((loss_kind == XR_LOSS_INFONCE || loss_kind == XR_LOSS_NCE || loss_kind == XR_LOSS_PAIRWISE_HINGE ||
                loss_kind == XR_LOSS_PAIRWISE_LOGISTIC) &&
                   cfg->num_hard_negatives == 0 && cfg->scale > 0.f,
               "xr_pool_step_compute_mon: serves the dot-family train losses (InfoNCE, NCE, pairwise hinge / "
               "logistic) with scale > 0 and no hard-negative mining");
  XR_CHECK_ARG(!dtok || dtok_dtype == XR_F32 || dtok_dtype == XR_BF16, "xr_pool_step_compute_mon: dtok must be fp32 or bf16");
  XR_CHECK_ARG(workspace_bytes >= xr_pool_step_monitor_workspace_bytes(n_pos, dim) && (uintptr_t)workspace % 256 == 0,
               "xr_pool_step_compute_mon: workspace too small or misaligned");
  int rc;
  if ((rc = check_fused_device("xr_pool_step_compute_mon"))) return rc;
  const StepWs w = carve_step_ws(workspace, n_pos);
  const MonitorWs mw = carve_monitor_ws((uint8_t*)workspace + w.bytes, n_pos);
  const StepScatter sc{w.inv_pos, w.sel_pos, n_pos, dtok, dtok_dtype == XR_BF16};
  const MonArgs mon{mw.part_dot, mw.part_cos, mw.inv_q, mw.inv_n, mw.t_cos, mw.row_out, mw.row_out2, mw.scratch2,
                    losses_dot, losses_cos, stats_out, n_pos};
  return fused_launch_all(w.q, w.pos, w.neg, n_pos, n_pos, loss_kind, cfg, nullptr, grad_scale, nullptr, loss_out,
                          nullptr, w.fused, true, as_stream(stream), dtok ? &sc : nullptr, &mon);
}

// ---- retrieval: group maxima of Q . Cat^T on the tensor cores -----------------------------------
// gmax[u, g] = max_{c in group g} q_u . cat_c   (rows >= n give -inf).  The k-th largest group maximum
// of a row is a lower bound of its k-th largest score (index.py:244-254 semantics, exact).  With
// tile_stride = s > 1 only every s-th tile of T rows is scored (T = 128 for u > 128, the CTA-pair kernel;
// 64 otherwise): the SAMPLE that gives xr_score_filter its thresholds.  Storage column
// (T / 16) * t + g = catalog rows [T * t * s + 16 g, +16): natural order [16 c, 16 c + 16) for s = 1.
static bool g_gmax_single = false;   // profiling aid: force the single-CTA retrieval kernel
static int retr_tile_rows(int64_t u) { return (u > fk::BM && !g_gmax_single) ? 128 : fk::BN; }

extern "C" int64_t xr_score_groupmax_ld(int64_t u, int64_t n, int64_t tile_stride) {
  if (tile_stride < 1) tile_stride = 1;
  const int64_t T = retr_tile_rows(u);
  const int64_t nt = ((n + T - 1) / T + tile_stride - 1) / tile_stride;
  return nt * (T / 16);
}

static int* retr_hang_flag() {   // device word for the bounded-wait diagnostics
  static int* hang = nullptr;
  if (!hang) {
    if (cudaMalloc(&hang, 256) != cudaSuccess) return nullptr;
    cudaMemset(hang, 0, 256);
  }
  return hang;
}

static int check_retr_args(const char* who, const void* q, int64_t u, const void* catalog, int64_t n,
                           int64_t dim) {
  XR_CHECK_ARG(q && catalog, "%s: null pointer", who);
  XR_CHECK_ARG(dim == fk::D, "%s: this build is specialised for dim = %d", who, fk::D);
  XR_CHECK_ARG(u > 0 && n > 0 && u < (1ll << 30) && n < (1ll << 31) - 4096, "%s: bad sizes", who);
  XR_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)catalog % 16 == 0),
               "%s: operands must be 16-byte aligned", who);
  return check_fused_device(who);
}

extern "C" int xr_score_groupmax(const void* q, int64_t u, const void* catalog, int64_t n,
                                 int64_t dim, int64_t tile_stride, float* gmax, int64_t ld, void* stream) {
  int rc;
  if ((rc = check_retr_args("xr_score_groupmax", q, u, catalog, n, dim))) return rc;
  XR_CHECK_ARG(gmax && tile_stride >= 1 && tile_stride <= 4096, "xr_score_groupmax: bad arguments");
  XR_CHECK_ARG(ld >= xr_score_groupmax_ld(u, n, tile_stride) && ld % 2 == 0 && (uintptr_t)gmax % 8 == 0,
               "xr_score_groupmax: ld must be even and >= xr_score_groupmax_ld(u, n, tile_stride), gmax 8-byte aligned");
  cudaStream_t s = as_stream(stream);
  int* hang = retr_hang_flag();
  XR_CHECK_ARG(hang, "xr_score_groupmax: out of device memory");
  const bool prof = g_prof_on && g_prof_n < kProfRing;
  if (prof) cudaEventRecord(g_prof_ev[g_prof_n][0], s);
  if (retr_tile_rows(u) == 128) {   // CTA pairs: 256 queries per pair, half the catalog bytes per SM
    rc = launch_score_gmax2(q, u, catalog, n, (int)tile_stride, gmax, ld, hang, s, g_ablate);
  } else {
    const int n_sm = sm_count();
    const int64_t nt = ((n + fk::BN - 1) / fk::BN + tile_stride - 1) / tile_stride;   // sampled tiles
    const GmaxPlan pl = make_gmax_plan(u, nt * fk::BN, n_sm);
    CUtensorMap tq, tc;
    if ((rc = make_tmap_bf16_rows(&tq, q, u, dim, dim, fk::BM))) return rc;
    if ((rc = make_tmap_bf16_rows(&tc, catalog, n, dim, dim, fk::BN))) return rc;
    FusedParams p{};
    p.m = (int)u; p.cn = (int)n; p.nt_count = pl.nt; p.spl = pl.spl; p.tiles_per_split = pl.tps;
    p.n_items = pl.n_items; p.rb_count = pl.rb; p.gmax = gmax; p.gmax_ld = ld; p.tile_stride = (int)tile_stride;
    p.hang_flag = hang;
    rc = launch_fused1<fk::KIND_GMAX, false>(tq, tc, p, pl.n_items < n_sm ? pl.n_items : n_sm, s);
  }
  if (prof) cudaEventRecord(g_prof_ev[g_prof_n++][1], s);
  return rc;
}

// ---- retrieval: threshold filter in the scoring epilogue ---------------------------------------------
// Every (score, local row) with score >= thresh[u * thresh_stride] is kept.  Storage: n_sub sub-buckets
// of cap_b slots per query, one per (catalog split of the scoring plan, column group) -- each owned by a
// single lane, filled without atomics -- plus one overflow list of ovf_cap slots per query for sub-buckets
// that run full (clustered catalogs).  b_count[u][s] = survivors sub-bucket s saw (> cap_b: the excess is
// in the overflow list), o_count[u] (zeroed by the caller) keeps counting past ovf_cap, which is how the
// consumer detects a loss.  With thresholds from a sample (xr_score_groupmax with tile_stride = s: the
// (k+28)-th largest group maximum of the sample is <= the (k+28)-th largest score of the catalog) about
// (k+28) * s rows survive per query: the catalog is read once and ~KBs per query are written.
extern "C" int xr_score_filter_layout(int64_t u, int64_t n, int64_t expected_survivors, int64_t* n_sub,
                                      int64_t* cap_b) {
  XR_CHECK_ARG(u > 0 && n > 0 && n_sub && cap_b && expected_survivors >= 0, "xr_score_filter_layout: bad arguments");
  int64_t subs;
  if (retr_tile_rows(u) == 128) subs = 4ll * gmax2_splits(u, n);
  else subs = (int64_t)fk::CG * make_gmax_plan(u, n, sm_count()).spl;
  *n_sub = subs;
  // four times the expected fill (+ slack), multiples of 8 slots: random catalogs never spill
  int64_t c = (4 * expected_survivors + subs - 1) / subs + 8;
  c = (c + 7) / 8 * 8;
  if (c < 16) c = 16;
  if (c > 4096) c = 4096;
  *cap_b = c;
  return XR_OK;
}

extern "C" int xr_score_filter(const void* q, int64_t u, const void* catalog, int64_t n, int64_t dim,
                               const float* thresh, int64_t thresh_stride, float* bucket_scores,
                               int32_t* bucket_rows, int32_t* bucket_count, int64_t n_sub, int64_t cap_b,
                               float* ovf_scores, int32_t* ovf_rows, int32_t* ovf_count, int64_t ovf_cap,
                               void* stream) {
  int rc;
  if ((rc = check_retr_args("xr_score_filter", q, u, catalog, n, dim))) return rc;
  XR_CHECK_ARG(thresh && bucket_scores && bucket_rows && bucket_count && ovf_scores && ovf_rows && ovf_count &&
                   n_sub >= 1 && cap_b >= 1 && cap_b <= (1 << 20) && ovf_cap >= 1 && ovf_cap < (1ll << 30) &&
                   thresh_stride >= 0,
               "xr_score_filter: bad arguments");
  cudaStream_t s = as_stream(stream);
  int* hang = retr_hang_flag();
  XR_CHECK_ARG(hang, "xr_score_filter: out of device memory");
  const FilterOut fo{bucket_scores, bucket_rows, bucket_count, ovf_scores, ovf_rows, ovf_count,
                     (int)n_sub, (int)cap_b, (int)ovf_cap};
  const bool prof = g_prof_on && g_prof_n < kProfRing;
  if (prof) cudaEventRecord(g_prof_ev[g_prof_n][0], s);
  if (retr_tile_rows(u) == 128) {
    rc = launch_score_filter2(q, u, catalog, n, thresh, thresh_stride, fo, hang, s);
  } else {
    const int n_sm = sm_count();
    const GmaxPlan pl = make_gmax_plan(u, n, n_sm);
    XR_CHECK_ARG(n_sub == (int64_t)fk::CG * pl.spl, "xr_score_filter: n_sub must be %d for this (u, n) (xr_score_filter_layout)",
                 fk::CG * pl.spl);
    CUtensorMap tq, tc;
    if ((rc = make_tmap_bf16_rows(&tq, q, u, dim, dim, fk::BM))) return rc;
    if ((rc = make_tmap_bf16_rows(&tc, catalog, n, dim, dim, fk::BN))) return rc;
    FusedParams p{};
    p.m = (int)u; p.cn = (int)n; p.nt_count = pl.nt; p.spl = pl.spl; p.tiles_per_split = pl.tps;
    p.n_items = pl.n_items; p.rb_count = pl.rb; p.hang_flag = hang;
    p.thresh = thresh; p.thresh_stride = thresh_stride; p.fo = fo;
    rc = launch_fused1<fk::KIND_FILTER, false>(tq, tc, p, pl.n_items < n_sm ? pl.n_items : n_sm, s);
  }
  if (prof) cudaEventRecord(g_prof_ev[g_prof_n++][1], s);
  return rc;
}

// profiling aid (not part of the product path): summed wait cycles per barrier tag of the LAST
// xr_fused_pool_loss call while enabled.  tags: 1 q_empty(producer) 2 ring-empty(producer)
// 3 p_full(gradient issuer) 4 q_full 5 o_empty 6 s_free(score issuer) 7 ring-full(score issuer)
// 8 s_full(epilogue warps) 9 o_full(epilogue warps)
// per-tile timestamps of CTA 0 from the last stats-enabled call: 64 tiles x 8 slots
// [0] score issue start [1] score issued+committed [2] epilogue woke on s_full [3] tcgen05.ld done
// [4] weights stored + p_full arrive [5] gradient issuer woke on p_full
extern "C" int xr_fused_timeline(long long* out512_host) {
  memcpy(out512_host, g_dbg_host, sizeof(g_dbg_host));   // 576 entries
  return XR_OK;
}

extern "C" int xr_fused_wait_stats(int enable, unsigned long long* out16_host) {
  g_wait_stats = (enable & 1) != 0;
  g_timeline = (enable & 2) != 0;   // per-tile timestamps without the wait counters
  g_gmax_single = (enable & 4) != 0;
  g_ctrl_low = (enable & 8) != 0;
  g_ablate = enable >> 8;   // ablation mask for timing experiments (results are garbage)
  if (out16_host) memcpy(out16_host, g_wait_host, sizeof(g_wait_host));
  return XR_OK;
}

extern "C" int xr_fused_profile(int enable) {
  if (enable && !g_prof_made) {
    for (int i = 0; i < kProfRing; ++i)
      for (int j = 0; j < 2; ++j) XR_CUDA(cudaEventCreate(&g_prof_ev[i][j]));
    g_prof_made = true;
  }
  g_prof_on = enable != 0;
  g_prof_n = 0;
  return XR_OK;
}

// durations (ms) of the main fused kernel launches recorded since xr_fused_profile(1);
// synchronises on the recorded events.  Returns the number written (<= max_n) or <0.
extern "C" int xr_fused_profile_read(float* ms_out_host, int max_n) {
  int n = g_prof_n < max_n ? g_prof_n : max_n;
  for (int i = 0; i < n; ++i) {
    XR_CUDA(cudaEventSynchronize(g_prof_ev[i][1]));
    XR_CUDA(cudaEventElapsedTime(&ms_out_host[i], g_prof_ev[i][0], g_prof_ev[i][1]));
  }
  g_prof_n = 0;
  return n;
}

