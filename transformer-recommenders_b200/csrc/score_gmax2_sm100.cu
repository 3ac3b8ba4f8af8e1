// Retrieval scoring on CTA PAIRS (tcgen05 cta_group::2), two epilogues over the same MMA pipeline
// (index.py:244-254 semantics, exact; fused_loss_sm100.cu holds the single-CTA versions for U <= 128):
//   MODE_GMAX    gmax[u, g] = max over the 16 catalog rows of group g of q_u . cat_c, optionally over
//                every `tile_stride`-th 128-row tile only (the SAMPLE whose (k+E)-th largest maximum is
//                a lower bound of the (k+E)-th largest score of the whole catalog);
//   MODE_FILTER  every (score, row) with score >= thresh[u] is kept: the (U, N) score matrix never reaches
//                HBM, and neither does anything of size U x N / 16.  Survivors go to SUB-BUCKETS owned by
//                one (row, catalog split, column group) each, i.e. by exactly one lane of one warp, whose
//                fill count lives in a register: no atomic, no L2 round trip on the scoring path (a
//                returning global atomic per survivor cost 0.2 ms of a 1.6 ms kernel).  A full sub-bucket
//                spills to the query's overflow list, the only place an atomic is left.
//
// Why pairs: shared-memory ingest by TMA is ~35 B/cycle/SM whatever the ring depth or multicast
// (profiles/microbench/tma_stream.cu), and a 128-query CTA needs 48 KB per 64 candidates = 1,400
// cycles against 1,150 cycles of MMA work: the single-CTA kernel is ingest-bound.  With
// cta_group::2 one MMA covers 256 queries (128 per CTA) x 128 candidates, and each CTA stages only
// ITS half of the candidate tile (B is split along N across the pair): per SM the same 48 KB now
// feed 1,536 cycles of MMA work on 128 x 128 outputs, at N = 128 where the SS MMA runs at pipe rate.
//
// Per CTA (640 threads): warp 0 TMA producer (own Q rows, own half of every catalog tile; every
// load signals the LEADER CTA's mbarrier), warp 1 of the leader issues the MMAs for both SMs,
// warps 4-19 epilogue over the CTA's own 128 TMEM lanes.  TMEM: 4 S buffers of 128 columns.
#include "common.cuh"
#include "sm100.cuh"
#include "filter.cuh"

namespace xr {

using namespace sm100;

extern bool g_ctrl_low;

namespace g2 {
constexpr int BM = 128;            // query rows per CTA (256 per pair)
constexpr int BNH = 64;            // catalog rows per CTA per tile (128 per pair)
constexpr int BN = 2 * BNH;
constexpr int D = 384;
constexpr int KB = D / 64;
constexpr int PAIRS = 8;           // ring of 16 KB slots: two adjacent 64x64 k-blocks
constexpr int SUB_BYTES = BNH * 64 * 2;
constexpr int QSUB_BYTES = BM * 64 * 2;
constexpr int Q_BYTES = KB * QSUB_BYTES;
constexpr int RING_BYTES = PAIRS * 2 * SUB_BYTES;
constexpr int NSB = 4;             // S buffers (128 TMEM columns each)
constexpr int BAR_OFF = Q_BYTES + RING_BYTES;
constexpr int NBARS = 2 * PAIRS + 2 + 2 * NSB;
constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 128 + EPI_WARPS * 32;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // shared::cluster address of the pair's EVEN CTA
}  // namespace g2

constexpr int MODE_GMAX = 0, MODE_FILTER = 1;

struct Gmax2Params {
  int u, n, nt_count, tile_stride, spl, tiles_per_split, n_items, qb_count;
  float* gmax;           // MODE_GMAX: (u, gmax_ld), natural order: column 8 t + g = rows [128 t s + 16 g, +16)
  long long gmax_ld;
  const float* thresh;   // MODE_FILTER: thresh[u * thresh_stride]
  long long thresh_stride;
  FilterOut fo;          // MODE_FILTER: survivor storage (filter.cuh)
  int* hang_flag;
  int ctrl_low;
  int ablate;   // DBG instantiation only: 1 skip catalog TMA, 2 skip epilogue TMEM loads, 4 skip gmax stores, 8 skip MMAs
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes go to the mbarrier of the pair's leader (even) CTA
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst_smem, const void* tmap, uint32_t bar, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(dst_smem),
      "l"(tmap), "r"(bar & g2::PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem of each CTA: its 128 rows] . B[smem of each CTA: its N/2 rows]
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all MMAs issued so far -> one arrive on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          bar),
      "h"((uint16_t)3)
      : "memory");
}
// arrive on the LEADER's copy of a barrier, from either CTA of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  // relaxed at cluster scope: a release.cluster arrive flushes L1 (~1,250 cycles, microbench
  // tma_stream.cu); all this arrive publishes is "my tcgen05.ld has completed" (wait::ld before it)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & g2::PEER_MASK)
               : "memory");
}

template <int MODE, bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2::THREADS, 1)
score_gmax2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                   const Gmax2Params p) {
  using namespace g2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t q_smem = base, ring = base + Q_BYTES, bars = base + BAR_OFF;
  auto bar_full = [&](uint32_t s) { return bars + 8u * s; };              // leader: pair s landed in BOTH CTAs
  auto bar_empty = [&](uint32_t s) { return bars + 8u * (PAIRS + s); };   // each CTA: its slot s is free
  const uint32_t bar_q_full = bars + 8u * (2 * PAIRS);                    // leader
  const uint32_t bar_q_empty = bars + 8u * (2 * PAIRS + 1);               // each CTA
  auto bar_s_full = [&](int b) { return bars + 8u * (2 * PAIRS + 2 + b); };        // each CTA
  auto bar_s_free = [&](int b) { return bars + 8u * (2 * PAIRS + 2 + NSB + b); };  // leader
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + NBARS * 8);

  // `warp` is the ROLE index; control roles on the highest hardware warps (see fused_loss_sm100.cu)
  const int hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = p.ctrl_low ? hw_warp : (hw_warp >= EPI_WARPS ? hw_warp - EPI_WARPS : hw_warp + 4);
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PAIRS; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_q_full, 1);
    mbar_init(bar_q_empty, 1);
    for (int b = 0; b < NSB; ++b) {
      mbar_init(bar_s_full(b), 1);
      mbar_init(bar_s_free(b), 2 * EPI_WARPS);   // one elected arrive per epilogue warp of BOTH CTAs
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_c);
  }
  if (warp == 1) tmem_alloc2(smem_u32((const void*)tmem_ptr_smem), 512);
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;

  // work items of a CTA pair: (block of 256 queries, split of the 128-candidate tiles); query block
  // fastest so that pairs sharing a catalog range run side by side (second read comes from L2)
  auto item_tiles = [&](int item, int& qb, int& t0, int& t1) {
    qb = item % p.qb_count;
    const int sp = item / p.qb_count;
    t0 = sp * p.tiles_per_split;
    t1 = min(p.nt_count, t0 + p.tiles_per_split);
  };

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    uint32_t g = 0, it = 0;
    for (int item = cluster_id; item < p.n_items; item += n_clusters, ++it) {
      int qb, t0, t1;
      item_tiles(item, qb, t0, t1);
      mbar_wait(bar_q_empty, (it & 1) ^ 1, p.hang_flag, 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(bar_q_full, 2 * Q_BYTES);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d_2sm(q_smem + kb * QSUB_BYTES, &tmap_q, bar_q_full, kb * 64, qb * 2 * BM + (int)rank * BM);
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t) {
#pragma unroll 1
        for (int pr = 0; pr < KB / 2; ++pr, ++g) {
          const uint32_t s = g & (PAIRS - 1);
          mbar_wait(bar_empty(s), ((g / PAIRS) & 1) ^ 1, p.hang_flag, 2);
          if (elect_one()) {
            if (DBG && (p.ablate & 1)) {
              if (leader) mbar_arrive(bar_full(s));
            } else {
              if (leader) mbar_expect_tx(bar_full(s), 2 * 2 * SUB_BYTES);
              const int row0 = t * p.tile_stride * BN + (int)rank * BNH;
              tma_load_2d_2sm(ring + s * 2 * SUB_BYTES, &tmap_c, bar_full(s), pr * 128, row0);
              tma_load_2d_2sm(ring + s * 2 * SUB_BYTES + SUB_BYTES, &tmap_c, bar_full(s), pr * 128 + 64, row0);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA, for both SMs) =====================
    // waits stay outside the elect blocks and the pair index is a run-time value, so ptxas keeps
    // the descriptors in uniform registers (see fused_loss_sm100.cu)
    constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);
    const uint64_t q_desc0 = umma_desc_sw128(q_smem, 16, 1024);
    const uint64_t ring_desc0 = umma_desc_sw128(ring, 16, 1024);
    uint32_t g = 0, tile = 0, it = 0;
    for (int item = cluster_id; item < p.n_items; item += n_clusters, ++it) {
      int qb, t0, t1;
      item_tiles(item, qb, t0, t1);
      const int T = t1 - t0;
      mbar_wait(bar_q_full, it & 1, p.hang_flag, 4);
      for (int tl = 0; tl < T; ++tl, ++tile) {
        const int sb = tile % NSB;
        const uint32_t use = tile / NSB;
        if (use >= 1) mbar_wait(bar_s_free(sb), (use - 1) & 1, p.hang_flag, 6);
        auto issue_pair = [&](int pr) {
          const uint32_t s = (g + pr) & (PAIRS - 1);
          const uint64_t a0 = q_desc0 + (uint64_t)(pr * ((2 * QSUB_BYTES) >> 4));
          const uint64_t b0 = ring_desc0 + (uint64_t)(s * ((2 * SUB_BYTES) >> 4));
          if (!(DBG && (p.ablate & 8)))
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss2(tmem + sb * BN, a0 + h * (QSUB_BYTES >> 4) + 2 * k, b0 + h * (SUB_BYTES >> 4) + 2 * k,
                       idesc, (pr | h | k) ? 1u : 0u);
          umma_commit2(bar_empty(s));   // the pair's slot is free in both CTAs once these MMAs are done
        };
        mbar_wait(bar_full(g & (PAIRS - 1)), (g / PAIRS) & 1, p.hang_flag, 7);
        mbar_wait(bar_full((g + 1) & (PAIRS - 1)), ((g + 1) / PAIRS) & 1, p.hang_flag, 7);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1
          for (int pr = 0; pr < 2; ++pr) issue_pair(pr);
        }
        __syncwarp();
        mbar_wait(bar_full((g + 2) & (PAIRS - 1)), ((g + 2) / PAIRS) & 1, p.hang_flag, 7);
        tc_fence_after();
        if (elect_one()) {
          int pr2 = 2;
          asm volatile("" : "+r"(pr2));
          issue_pair(pr2);
          umma_commit2(bar_s_full(sb));
          if (tl == T - 1) umma_commit2(bar_q_empty);
        }
        __syncwarp();
        g += KB / 2;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (both CTAs) ================================
    const int quad = warp & 3, cg = (warp - 4) >> 2;   // TMEM lane quadrant, 32-column group
    const int r_local = quad * 32 + lane;
    const uint32_t tmem_lane = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t tile = 0;
    for (int item = cluster_id; item < p.n_items; item += n_clusters) {
      int qb, t0, t1;
      item_tiles(item, qb, t0, t1);
      const int row = qb * 2 * BM + (int)rank * BM + r_local;
      const bool row_ok = row < p.u;
      float* out_row = MODE == MODE_GMAX ? p.gmax + (long long)row * p.gmax_ld : nullptr;
      // a row past u never passes the filter; NaN thresholds cannot occur (they are group maxima)
      const float th = (MODE == MODE_FILTER && row_ok) ? __ldg(p.thresh + (long long)row * p.thresh_stride)
                                                       : CUDART_INF_F;
      // this lane's sub-bucket for the item: (row, split of the catalog, column group)
      const int sub = (item / p.qb_count) * 4 + cg;
      float* b_scores = nullptr;
      int32_t* b_rows = nullptr;
      int bcount = 0;
      if (MODE == MODE_FILTER && row_ok) {
        const long long off = ((long long)row * p.fo.n_sub + sub) * p.fo.cap_b;
        b_scores = p.fo.b_scores + off;
        b_rows = p.fo.b_rows + off;
      }
      for (int t = t0; t < t1; ++t, ++tile) {
        const int sb = tile % NSB;
        mbar_wait(bar_s_full(sb), (tile / NSB) & 1, p.hang_flag, 8);
        tc_fence_after();
        uint32_t v[32];
        if (DBG && (p.ablate & 2)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tile + j;
        } else {
          tmem_ld32(tmem_lane + sb * BN + cg * 32, v);
          tmem_wait_ld();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(bar_s_free(sb));   // the logits are in registers
        const int c0 = t * p.tile_stride * BN + cg * 32;       // first catalog row of this column group
        if (MODE == MODE_GMAX) {
          float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (c0 + j < p.n) m0 = fmaxf(m0, __uint_as_float(v[j]));
            if (c0 + 16 + j < p.n) m1 = fmaxf(m1, __uint_as_float(v[16 + j]));
          }
          // natural order: storage column 8 t + 2 cg + h holds the rows [c0 + 16 h, +16), so that the
          // (maximum desc, column asc) order of xr_topk is (maximum desc, first row asc); the four
          // column-group warps of a row fill one 32-byte sector per tile between them
          if (row_ok && !(DBG && (p.ablate & 4)))
            *reinterpret_cast<float2*>(out_row + 8ll * t + 2 * cg) = make_float2(m0, m1);
        } else {
          // one test for the 32 scores keeps the common case (no survivor) at one instruction per score
          const int lim = p.n - c0;                 // rows past n are TMA zero fill, not catalog rows
          if (lim < 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j >= lim) v[j] = 0xFF800000u;     // -inf: never a survivor (thresholds are > -inf or all-pass)
          }
          float mx = __uint_as_float(v[0]);
#pragma unroll
          for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
          if (mx >= th && lim > 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float sc = __uint_as_float(v[j]);
              if (sc >= th && j < lim) filter_keep(p.fo, row, b_scores, b_rows, bcount, sc, c0 + j);
            }
          }
        }
      }
      if (MODE == MODE_FILTER && row_ok) p.fo.b_count[(long long)row * p.fo.n_sub + sub] = bcount;
    }
  }
  tc_fence_before();
  cluster_sync_all();   // neither CTA may leave (or free TMEM) while its peer can still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem, 512);
  }
}

int make_tmap_bf16_rows(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld_elems,
                        int box_rows);

// sampled tiles of 128 catalog rows: every tile_stride-th one
int gmax2_tiles(int64_t n, int tile_stride) {
  const int64_t nt = (n + g2::BN - 1) / g2::BN;
  return (int)((nt + tile_stride - 1) / tile_stride);
}

// work items of the CTA pairs: query blocks x splits of the (sampled) tiles, whole waves of pairs
static void plan_gmax2(Gmax2Params& p, int64_t u, int64_t n, int tile_stride, int n_clusters) {
  using namespace g2;
  p.u = (int)u; p.n = (int)n;
  p.tile_stride = tile_stride;
  p.qb_count = (int)((u + 2 * BM - 1) / (2 * BM));
  p.nt_count = gmax2_tiles(n, tile_stride);
  int best_spl = 1;
  long long best = -1;
  const int max_spl = p.nt_count < 4096 ? p.nt_count : 4096;
  for (int sp = 1; sp <= max_spl; ++sp) {
    const int tps = (p.nt_count + sp - 1) / sp;
    const int s_eff = (p.nt_count + tps - 1) / tps;
    const long long items = (long long)p.qb_count * s_eff;
    const long long waves = (items + n_clusters - 1) / n_clusters;
    const long long cost = waves * (tps + 4);
    if (best < 0 || cost < best) {
      best = cost;
      best_spl = s_eff;
    }
    if (items > 8LL * n_clusters) break;
  }
  p.tiles_per_split = (p.nt_count + best_spl - 1) / best_spl;
  p.spl = (p.nt_count + p.tiles_per_split - 1) / p.tiles_per_split;
  p.n_items = p.qb_count * p.spl;
}

template <int MODE>
static int launch_gmax2_mode(const void* q, int64_t u, const void* catalog, int64_t n, Gmax2Params& p,
                             cudaStream_t s, int ablate) {
  using namespace g2;
  const int n_clusters = sm_count() / 2;
  CUtensorMap tq, tc;
  int rc;
  if ((rc = make_tmap_bf16_rows(&tq, q, u, D, D, BM))) return rc;
  if ((rc = make_tmap_bf16_rows(&tc, catalog, n, D, D, BNH))) return rc;
  static bool configured = false;
  if (!configured) {
    XR_CUDA(cudaFuncSetAttribute(score_gmax2_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    XR_CUDA(cudaFuncSetAttribute(score_gmax2_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int pairs = p.n_items < n_clusters ? p.n_items : n_clusters;
  p.ablate = ablate;
  p.ctrl_low = g_ctrl_low;
  if (ablate) score_gmax2_kernel<MODE, true><<<2 * pairs, THREADS, SMEM_BYTES, s>>>(tq, tc, p);   // timing experiments
  else score_gmax2_kernel<MODE, false><<<2 * pairs, THREADS, SMEM_BYTES, s>>>(tq, tc, p);
  XR_LAUNCH_CHECK("score_gmax2_kernel");
  return XR_OK;
}

// host: called by xr_score_groupmax for u > 128
int launch_score_gmax2(const void* q, int64_t u, const void* catalog, int64_t n, int tile_stride, float* gmax,
                       int64_t ld, int* hang_flag, cudaStream_t s, int ablate) {
  Gmax2Params p{};
  plan_gmax2(p, u, n, tile_stride, sm_count() / 2);
  p.gmax = gmax; p.gmax_ld = ld; p.hang_flag = hang_flag;
  return launch_gmax2_mode<MODE_GMAX>(q, u, catalog, n, p, s, ablate);
}

// catalog splits of the pair kernel for (u, n): the number of sub-buckets per query is 4 x this
int gmax2_splits(int64_t u, int64_t n) {
  Gmax2Params p{};
  plan_gmax2(p, u, n, 1, sm_count() / 2);
  return p.spl;
}

// host: called by xr_score_filter for u > 128
int launch_score_filter2(const void* q, int64_t u, const void* catalog, int64_t n, const float* thresh,
                         int64_t thresh_stride, const FilterOut& fo, int* hang_flag, cudaStream_t s) {
  Gmax2Params p{};
  plan_gmax2(p, u, n, 1, sm_count() / 2);
  if (fo.n_sub != p.spl * 4) {
    set_error("xr_score_filter: n_sub must be %d for this (u, n) (xr_score_filter_layout)", p.spl * 4);
    return XR_E_INVALID;
  }
  p.thresh = thresh; p.thresh_stride = thresh_stride;
  p.fo = fo;
  p.hang_flag = hang_flag;
  return launch_gmax2_mode<MODE_FILTER>(q, u, catalog, n, p, s, 0);
}

}  // namespace xr
