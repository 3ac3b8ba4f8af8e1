// SeqBatch construction on the device: the step immediately BEFORE the scoring-and-loss path.
//
// Replaces, for a whole batch in one launch, SeqDataset.__getitem__ + collate
// (xfmr_rec/data.py:669-805):
//   sample_sequence   data.py:669-689  positions 0..len-2; if more than max_seq_length of them, a
//                                      uniform subset WITHOUT replacement, sorted
//   sample_positives  data.py:691-721  per sampled position: uniform over the positively-labelled
//                                      events among the next `pos_lookahead` (0 = all later) events;
//                                      none -> 0 (padding)
//   sample_negatives  data.py:723-747  seq_len items uniform over (all items - the user's history),
//                                      without replacement when enough candidates exist, with
//                                      replacement otherwise; empty candidate set -> all items
//   collate           data.py:787-805  right-padding with 0
// The reference does this in Python per example (set differences over the whole catalog per row).
// Here: one thread block per sequence, counter-based Philox4x32-10 randomness keyed by
// (seed, step, dataset row) -- every output element is a pure function of those, so the result is
// run-to-run deterministic, independent of the batch composition, and the numpy oracle
// (oracle/xfmr_oracle.py:seq_sample_batch) reproduces it bit for bit.
//
// Exactness of the distributions (no modulo bias beyond 2^-64 * n, no approximation):
//   * subset of positions: every position gets a 32-bit random key, the L smallest (ties -> lower
//     position) are taken -- a uniformly random L-subset -- by an MSB-first radix select in shared
//     memory and an ordered compaction (positions come out sorted, as np.sort does);
//   * positives: rank r uniform in [0, #positives in the window), located by binary search on the
//     per-history inclusive prefix count of positive labels;
//   * negatives: rank r uniform in [0, N - |history|) mapped to the r-th item NOT in the (sorted,
//     unique) history by binary search -- no rejection against the history; "without replacement"
//     is the sequential rule "redraw while the value was already taken", evaluated in parallel
//     rounds where, among equal draws of one round, the lowest sequence position keeps the value.
#include "common.cuh"
#include "philox.cuh"

#include <climits>

namespace xr {

namespace sq {
constexpr int THREADS = 256;
constexpr int STREAM_POSITIONS = 0, STREAM_POSITIVES = 1, STREAM_NEGATIVES = 2;
constexpr int EMPTY = -1;

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Philox4x32-10: philox.cuh (shared with the encoder's dropout)
using xr::U4;
using xr::philox;
// uniform integer in [0, n): high 64 bits of (64 random bits) x n -- bias < n / 2^64
__device__ __forceinline__ uint64_t uniform_below(const U4& r, uint64_t n) {
  const uint64_t u = ((uint64_t)r.x << 32) | (uint64_t)r.y;
  return __umul64hi(u, n);
}
__device__ __forceinline__ uint32_t hash32(uint32_t v) {
  v ^= v >> 16; v *= 0x7FEB352Du; v ^= v >> 15; v *= 0x846CA68Bu; v ^= v >> 16;
  return v;
}
}  // namespace sq

struct SeqSampleParams {
  const int64_t* hist_off;      // [H+1] CSR offsets of the histories
  const int64_t* items;         // [total] item idx (1-based; 0 = padding never appears)
  const uint8_t* labels;        // [total] positive-interaction flags
  const int32_t* pos_prefix;    // [total] inclusive count of positive labels within the history
  const int64_t* uniq_off;      // [H+1] CSR offsets of the sorted unique item lists
  const int64_t* uniq_items;    // [totalU] ascending, unique per history
  const int64_t* row_hist;      // [R] dataset row -> history (duplicate_rows, data.py:618-636); nullable = identity
  const int64_t* rows;          // [B] dataset rows of this batch
  int64_t n_batch, n_items;
  int max_len, lookahead;
  uint64_t seed, step;
  int64_t* hist_out;            // [B][max_len]
  int64_t* pos_out;             // [B][max_len]
  int64_t* neg_out;             // [B][max_len]
  int32_t* len_out;             // [B]
  int table_size;               // power of two >= 2 * max_len
};

// dynamic shared memory: sel[max_len] | neg[max_len] | attempt[max_len] | slot[max_len] |
//                        state[max_len] (0 done, 1 pending, 2 claiming) | vals[table] | owner[table]
__global__ void __launch_bounds__(sq::THREADS)
seq_sample_kernel(const SeqSampleParams p) {
  using namespace sq;
  extern __shared__ int32_t sm[];
  int32_t* sel = sm;
  int32_t* negv = sel + p.max_len;
  int32_t* attempt = negv + p.max_len;
  int32_t* slot = attempt + p.max_len;
  int32_t* state = slot + p.max_len;
  int32_t* vals = state + p.max_len;
  int32_t* owner = vals + p.table_size;
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_sel[2];
  __shared__ int s_warp_cnt[THREADS / 32];
  __shared__ int s_flag;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = p.max_len;
  for (int64_t b = blockIdx.x; b < p.n_batch; b += gridDim.x) {
    const int64_t row = p.rows[b];
    const int64_t h = p.row_hist ? p.row_hist[row] : row;
    const int64_t s0 = p.hist_off[h];
    const int n = (int)(p.hist_off[h + 1] - s0);
    const int64_t* items = p.items + s0;
    const int32_t* pre = p.pos_prefix + s0;
    const uint64_t key = splitmix64(splitmix64(p.seed ^ splitmix64(p.step)) ^ (uint64_t)row);
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);

    // ---- sample_sequence (data.py:669-689) ------------------------------------------------------
    const int cand = n > 0 ? n - 1 : 0;
    const int seq_len = cand < L ? cand : L;
    if (cand <= L) {
      for (int i = tid; i < seq_len; i += THREADS) sel[i] = i;
    } else {
      auto rkey = [&](int pos) { return philox((uint32_t)pos, 0u, STREAM_POSITIONS, 0u, k0, k1).x; };
      // MSB-first radix select of the L-th SMALLEST 32-bit key
      unsigned prefix = 0, prefix_mask = 0, want = (unsigned)L;
      for (int shift = 24; shift >= 0; shift -= 8) {
        for (int bkt = tid; bkt < 256; bkt += THREADS) s_hist[bkt] = 0;
        __syncthreads();
        for (int pos = tid; pos < cand; pos += THREADS) {
          const unsigned k = rkey(pos);
          if ((k & prefix_mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
          unsigned acc = 0;
          int bkt = 0;
          for (; bkt < 255; ++bkt) {
            if (acc + s_hist[bkt] >= want) break;
            acc += s_hist[bkt];
          }
          s_sel[0] = (unsigned)bkt;
          s_sel[1] = want - acc;   // still to take inside this bucket
        }
        __syncthreads();
        prefix |= s_sel[0] << shift;
        prefix_mask |= 255u << shift;
        want = s_sel[1];
        __syncthreads();
      }
      // keys < prefix are all in; of the keys == prefix the first `want` by position are in;
      // ordered compaction keeps the positions ascending (np.sort, data.py:689)
      const unsigned tau = prefix;
      int taken_eq = 0, written = 0;   // uniform across the block
      for (int base = 0; base < cand; base += THREADS) {
        const int pos = base + tid;
        bool lt = false, eq = false;
        if (pos < cand) {
          const unsigned k = rkey(pos);
          lt = k < tau;
          eq = k == tau;
        }
        // rank of this thread among the equals of the chunk
        const unsigned bal_eq = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal_eq);
        __syncthreads();
        int before_eq = 0, total_eq = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) {
          if (w < warp) before_eq += s_warp_cnt[w];
          total_eq += s_warp_cnt[w];
        }
        const int rank_eq = taken_eq + before_eq + __popc(bal_eq & ((1u << lane) - 1u));
        const bool in = lt || (eq && rank_eq < (int)want);
        __syncthreads();
        const unsigned bal_in = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal_in);
        __syncthreads();
        int before_in = 0, total_in = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) {
          if (w < warp) before_in += s_warp_cnt[w];
          total_in += s_warp_cnt[w];
        }
        if (in) sel[written + before_in + __popc(bal_in & ((1u << lane) - 1u))] = pos;
        taken_eq += total_eq;
        written += total_in;
        __syncthreads();
      }
    }
    __syncthreads();

    // ---- history + sample_positives (data.py:691-721) -------------------------------------------
    for (int i = tid; i < L; i += THREADS) {
      int64_t hv = 0, pv = 0;
      if (i < seq_len) {
        const int idx = sel[i];
        hv = items[idx];
        const int start = idx + 1;
        const int end = p.lookahead > 0 ? min(n, start + p.lookahead) : n;   // [start, end)
        const int base_cnt = pre[start - 1];
        const int cnt = end > start ? pre[end - 1] - base_cnt : 0;
        if (cnt > 0) {
          const U4 r = philox((uint32_t)i, 0u, STREAM_POSITIVES, 0u, k0, k1);
          const int want_cnt = base_cnt + 1 + (int)uniform_below(r, (uint64_t)cnt);
          int lo = start, hi = end - 1;   // first j with pre[j] >= want_cnt (that event is positive)
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (pre[mid] >= want_cnt) hi = mid;
            else lo = mid + 1;
          }
          pv = items[lo];
        }
      }
      p.hist_out[b * L + i] = hv;
      p.pos_out[b * L + i] = pv;
    }

    // ---- sample_negatives (data.py:723-747) -----------------------------------------------------
    const int64_t u0 = p.uniq_off[h];
    int n_uniq = (int)(p.uniq_off[h + 1] - u0);
    const int64_t* uniq = p.uniq_items + u0;
    int64_t n_cand = p.n_items - n_uniq;
    if (n_cand <= 0) {   // history covers the catalog: candidates = all items (data.py:741-742)
      n_cand = p.n_items;
      n_uniq = 0;
    }
    const bool with_replacement = n_cand < (int64_t)seq_len;
    auto draw = [&](int i, int att) -> int32_t {
      const U4 r = philox((uint32_t)i, (uint32_t)att, STREAM_NEGATIVES, 0u, k0, k1);
      const int64_t rank = (int64_t)uniform_below(r, (uint64_t)n_cand);
      // the rank-th item (0-based) not in the history: smallest k with uniq[k] - 1 - k > rank
      int lo = 0, hi = n_uniq;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (uniq[mid] - 1 - mid > rank) hi = mid;
        else lo = mid + 1;
      }
      return (int32_t)(rank + 1 + lo);
    };
    if (with_replacement) {
      for (int i = tid; i < seq_len; i += THREADS) negv[i] = draw(i, 0);
    } else {
      const unsigned tmask = (unsigned)p.table_size - 1u;
      for (int s = tid; s < p.table_size; s += THREADS) {
        vals[s] = EMPTY;
        owner[s] = INT_MAX;
      }
      for (int i = tid; i < seq_len; i += THREADS) {
        attempt[i] = 0;
        state[i] = 1;
      }
      __syncthreads();
      for (;;) {
        // phase A: draw; a value already in the table was taken in an earlier round -> redraw
        if (tid == 0) s_flag = 0;
        __syncthreads();
        for (int i = tid; i < seq_len; i += THREADS) {
          if (state[i] == 0) continue;
          const int32_t v = draw(i, attempt[i]);
          negv[i] = v;
          unsigned s = hash32((uint32_t)v) & tmask;
          bool found = false;
          while (vals[s] != EMPTY) {
            if (vals[s] == v) {
              found = true;
              break;
            }
            s = (s + 1) & tmask;
          }
          if (found) {
            attempt[i] += 1;
            state[i] = 1;
          } else {
            state[i] = 2;
          }
          s_flag = 1;
        }
        __syncthreads();
        if (!s_flag) break;
        // phase B: claim; among equal draws of this round the lowest position owns the value
        for (int i = tid; i < seq_len; i += THREADS) {
          if (state[i] != 2) continue;
          const int32_t v = negv[i];
          unsigned s = hash32((uint32_t)v) & tmask;
          for (;;) {
            const int32_t old = atomicCAS(&vals[s], EMPTY, v);
            if (old == EMPTY || old == v) break;
            s = (s + 1) & tmask;
          }
          atomicMin(&owner[s], i);
          slot[i] = (int32_t)s;
        }
        __syncthreads();
        // phase C: the owner keeps it, everyone else redraws
        for (int i = tid; i < seq_len; i += THREADS) {
          if (state[i] != 2) continue;
          if (owner[slot[i]] == i) {
            state[i] = 0;
          } else {
            attempt[i] += 1;
            state[i] = 1;
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
    for (int i = tid; i < L; i += THREADS) p.neg_out[b * L + i] = i < seq_len ? (int64_t)negv[i] : 0;
    if (tid == 0 && p.len_out) p.len_out[b] = seq_len;
    __syncthreads();
  }
}

}  // namespace xr

using namespace xr;

extern "C" int xr_seq_sample_batch(const int64_t* hist_off, const int64_t* items,
                                   const uint8_t* labels, const int32_t* pos_prefix,
                                   const int64_t* uniq_off, const int64_t* uniq_items,
                                   const int64_t* row_hist, const int64_t* rows, int64_t n_batch,
                                   int64_t n_items, int max_seq_length, int pos_lookahead,
                                   uint64_t seed, uint64_t step, int64_t* history_out,
                                   int64_t* pos_out, int64_t* neg_out, int32_t* seq_len_out,
                                   void* stream) {
  XR_CHECK_ARG(hist_off && items && labels && pos_prefix && uniq_off && uniq_items && rows &&
                   history_out && pos_out && neg_out,
               "xr_seq_sample_batch: null pointer");
  XR_CHECK_ARG(n_batch >= 0 && n_items > 0 && n_items < (1ll << 31), "xr_seq_sample_batch: bad sizes");
  XR_CHECK_ARG(max_seq_length > 0 && max_seq_length <= 4096 && pos_lookahead >= 0,
               "xr_seq_sample_batch: max_seq_length must be in [1, 4096], pos_lookahead >= 0");
  if (n_batch == 0) return XR_OK;
  SeqSampleParams p{};
  p.hist_off = hist_off; p.items = items; p.labels = labels; p.pos_prefix = pos_prefix;
  p.uniq_off = uniq_off; p.uniq_items = uniq_items; p.row_hist = row_hist; p.rows = rows;
  p.n_batch = n_batch; p.n_items = n_items; p.max_len = max_seq_length; p.lookahead = pos_lookahead;
  p.seed = seed; p.step = step; p.hist_out = history_out; p.pos_out = pos_out; p.neg_out = neg_out;
  p.len_out = seq_len_out;
  int ts = 64;
  while (ts < 2 * max_seq_length) ts <<= 1;
  p.table_size = ts;
  const size_t smem = (size_t)(5 * max_seq_length + 2 * ts) * sizeof(int32_t);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    XR_CUDA(cudaFuncSetAttribute(seq_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(n_batch < cap ? n_batch : cap);
  seq_sample_kernel<<<grid, sq::THREADS, smem, as_stream(stream)>>>(p);
  XR_LAUNCH_CHECK("seq_sample");
  return XR_OK;
}
