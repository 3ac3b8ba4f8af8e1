// Family 3: exact top-k under the total order (score descending, column ascending).
//
// This is the exact computation that the reference's ANN search approximates
// (xfmr_rec/index.py:244-251 `.limit(top_k)`, :465-467) with the north_star tie rule
// "ties broken by lower item id" (oracle: stable descending sort).
//
// HBM-bound design: every score is read ONCE with 128-bit loads.  A block streams one chunk
// of one row; a score is appended to a shared-memory candidate buffer only if its 64-bit key
//     key = (order_preserving(score) << 32) | (0xFFFFFFFF - column)
// beats the block's running k-th best key, so after a short warm-up almost no element leaves
// the compare.  When the buffer fills, a bitonic sort keeps the best k and raises the
// threshold.  Chunks of a row are merged by the same kernel running over the partial keys.
// Selection on a strict total order makes single-GPU, chunked and sharded results identical.
#include "common.cuh"
#include "filter.cuh"

namespace xr {

constexpr int TK_THREADS = 256;
constexpr int TK_CAP = 2048;        // candidate buffer (keys)
constexpr int TK_PER_ITER = 1024;   // worst-case appends per iteration
constexpr int TK_MAX_K = TK_CAP - TK_PER_ITER;

__device__ __forceinline__ uint64_t make_key(float score, uint32_t col) {
  return ((uint64_t)float_key(score) << 32) | (uint64_t)(0xFFFFFFFFu - col);
}

// descending bitonic sort of s_keys[0..n2) (n2 = power of two), all threads of the block
__device__ __forceinline__ void bitonic_desc(uint64_t* s_keys, int n2) {
  for (int k2 = 2; k2 <= n2; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += TK_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = s_keys[i], b = s_keys[ixj];
          const bool desc = (i & k2) == 0;
          if (desc ? (a < b) : (a > b)) {
            s_keys[i] = b;
            s_keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

struct TopkState {
  uint64_t* keys;   // [TK_CAP] shared
  int* count;       // shared
  uint64_t* tau;    // shared: current k-th best key (0 = accept everything)
};

// keep the best k of the buffered keys, update the threshold; block-uniform call
__device__ __forceinline__ void compact(const TopkState& st, int k) {
  __syncthreads();
  const int n = *st.count;
  int n2 = 2;
  while (n2 < n) n2 <<= 1;
  for (int i = n + threadIdx.x; i < n2; i += TK_THREADS) st.keys[i] = 0ull;
  __syncthreads();
  bitonic_desc(st.keys, n2);
  if (threadIdx.x == 0) {
    if (n >= k) {
      *st.count = k;
      *st.tau = st.keys[k - 1];
    }
  }
  __syncthreads();
}

// Threshold refresh WITHOUT a sort: MSB-first radix select (8-bit digits) of the k-th largest
// buffered key, then an unordered filter of the survivors.  ~8x cheaper than the bitonic sort;
// the single ordered sort happens once, on k keys, when the block emits its result.
// Keys are unique (the low word is the column), so exactly k keys survive.  Block-uniform call.
__device__ __forceinline__ void compact_select(const TopkState& st, int k) {
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_pick[2];
  __syncthreads();
  const int n = *st.count;
  if (n <= k) return;
  constexpr int PER = TK_CAP / TK_THREADS;
  uint64_t my[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = threadIdx.x + j * TK_THREADS;
    my[j] = i < n ? st.keys[i] : 0ull;   // 0 never matches a live prefix bucket that is selected
  }
  uint64_t prefix = 0ull, mask = 0ull;
  unsigned want = (unsigned)k;
  const int lane = threadIdx.x & 31;
  for (int shift = 56; shift >= 0; shift -= 8) {
    s_hist[threadIdx.x] = 0u;   // TK_THREADS == 256 bins
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = threadIdx.x + j * TK_THREADS;
      if (i < n && (my[j] & mask) == prefix) atomicAdd(&s_hist[(unsigned)(my[j] >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // warp 0 walks the 256 bins from the top
      unsigned c[8], tot = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        c[b] = s_hist[255 - (lane * 8 + b)];
        tot += c[b];
      }
      unsigned incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      unsigned acc = incl - tot;
      if (acc < want && want <= incl) {
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if (acc < want && want <= acc + c[b]) {
            s_pick[0] = 255u - (unsigned)(lane * 8 + b);
            s_pick[1] = want - acc;
          }
          acc += c[b];
        }
      }
    }
    __syncthreads();
    prefix |= (uint64_t)s_pick[0] << shift;
    mask |= 0xFFull << shift;
    want = s_pick[1];
  }
  const uint64_t kth = prefix;   // all 64 bits resolved: the k-th largest key itself
  __syncthreads();
  if (threadIdx.x == 0) {
    *st.count = 0;
    *st.tau = kth;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = threadIdx.x + j * TK_THREADS;
    if (i < n && my[j] >= kth) st.keys[atomicAdd(st.count, 1)] = my[j];
  }
  __syncthreads();
}

// warp-aggregated append of keys that beat the threshold
__device__ __forceinline__ void offer(const TopkState& st, bool take, uint64_t key) {
  const unsigned bal = __ballot_sync(0xffffffffu, take);
  if (bal == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(st.count, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (take) st.keys[base + __popc(bal & ((1u << lane) - 1u))] = key;
}

// ---- generic (slow) path: SRC 0 fp32 scores, 1 uint64 keys, 2 fp32 scores + int64 ids ----------
// one block-wide barrier per 1024 elements; used for merges, unaligned rows and as the overflow
// fallback of the streaming fast path below.
template <int SRC>
__device__ __forceinline__ void slow_range(const TopkState& st, const float* srow,
                                           const int64_t* irow, const uint64_t* krow, int64_t lo,
                                           int64_t hi, int k) {
  const int64_t len = hi > lo ? hi - lo : 0;
  const int64_t iters = (len + TK_PER_ITER - 1) / TK_PER_ITER;
  for (int64_t it = 0; it < iters; ++it) {
    // make room: an iteration can append at most TK_PER_ITER keys (block-uniform branch)
    if (*st.count > TK_CAP - TK_PER_ITER) compact_select(st, k);
    const uint64_t tau = *st.tau;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t e = lo + it * TK_PER_ITER + r * TK_THREADS + threadIdx.x;
      const bool in = e < hi;
      uint64_t key = 0;
      if (in) {
        if (SRC == 0) key = make_key(srow[e], (uint32_t)e);
        else if (SRC == 1) key = krow[e];
        else {
          const int64_t id = irow[e];
          key = id < 0 ? 0ull : make_key(srow[e], (uint32_t)id);
        }
      }
      offer(st, in && key > tau, key);
    }
    __syncthreads();
  }
}

template <int SRC>
__global__ void __launch_bounds__(TK_THREADS)
topk_generic_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids,
                    const uint64_t* __restrict__ in_keys, int64_t n, int64_t ld, int64_t chunk_len,
                    int k, uint64_t* __restrict__ out_keys /* [u][P][k] */) {
  __shared__ uint64_t s_keys[TK_CAP];
  __shared__ int s_count;
  __shared__ uint64_t s_tau;
  TopkState st{s_keys, &s_count, &s_tau};
  const int64_t u = blockIdx.y, p = blockIdx.x, P = gridDim.x;
  const int64_t lo = p * chunk_len;
  const int64_t hi = lo + chunk_len < n ? lo + chunk_len : n;
  if (threadIdx.x == 0) {
    s_count = 0;
    s_tau = 0ull;
  }
  __syncthreads();
  slow_range<SRC>(st, SRC != 1 ? scores + u * ld : nullptr, SRC == 2 ? ids + u * ld : nullptr,
                  SRC == 1 ? in_keys + u * ld : nullptr, lo, hi, k);
  compact_select(st, k);
  compact(st, k);   // ordered emit: a sort of <= k keys
  const int cnt = s_count < k ? s_count : k;
  uint64_t* o = out_keys + (u * P + p) * (int64_t)k;
  for (int i = threadIdx.x; i < k; i += TK_THREADS) o[i] = i < cnt ? s_keys[i] : 0ull;
}

// ---- all-gather + merge as ONE kernel over peer memory (sharded retrieval) ----------------------
// Every rank leaves its per-shard (U, k) {score, id} lists in a buffer its peers can address over
// NVLink / NVSwitch (CUDA IPC / symmetric memory); after a cross-GPU barrier each rank's merge
// kernel LOADS the G lists of a query straight from the peers -- volatile-free plain loads through
// the peer mappings -- while building its selection keys: no NCCL all-gather, no staging copy, no
// permute.  Block per query; the candidate e in [0, G k) lives at peer e / k, slot e % k.
constexpr int TK_MAX_PEERS = 16;
struct PeerLists {
  const float* scores[TK_MAX_PEERS];     // each (U, k) fp32, row-major
  const int64_t* ids[TK_MAX_PEERS];      // each (U, k) int64 global ids, -1 = none
};
__global__ void __launch_bounds__(TK_THREADS)
topk_merge_peers_kernel(const PeerLists peers, int n_peers, int k, uint64_t* __restrict__ out_keys) {
  __shared__ uint64_t s_keys[TK_CAP];
  __shared__ int s_count;
  __shared__ uint64_t s_tau;
  TopkState st{s_keys, &s_count, &s_tau};
  const int64_t u = blockIdx.x;
  if (threadIdx.x == 0) {
    s_count = 0;
    s_tau = 0ull;
  }
  __syncthreads();
  const int64_t total = (int64_t)n_peers * k;
  const int64_t iters = (total + TK_PER_ITER - 1) / TK_PER_ITER;
  for (int64_t it = 0; it < iters; ++it) {
    if (*st.count > TK_CAP - TK_PER_ITER) compact_select(st, k);
    const uint64_t tau = *st.tau;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t e = it * TK_PER_ITER + r * TK_THREADS + threadIdx.x;
      const bool in = e < total;
      uint64_t key = 0;
      if (in) {
        const int g = (int)(e / k), j = (int)(e - (int64_t)g * k);
        // peer memory is written by another GPU between launches: bypass the (non-coherent) L1
        const int64_t id = __ldcg(peers.ids[g] + u * k + j);
        const float sc = __ldcg(peers.scores[g] + u * k + j);
        key = id < 0 ? 0ull : make_key(sc, (uint32_t)id);
      }
      offer(st, in && key > tau, key);
    }
    __syncthreads();
  }
  compact_select(st, k);
  compact(st, k);
  const int cnt = s_count < k ? s_count : k;
  uint64_t* o = out_keys + u * (int64_t)k;
  for (int i = threadIdx.x; i < k; i += TK_THREADS) o[i] = i < cnt ? s_keys[i] : 0ull;
}

// ---- streaming fast path over fp32 scores (16-byte aligned rows) --------------------------------
// Super-blocks of 8 x 2048 scores with NO barrier inside: a score is compared against the float
// threshold (one FSETP); the rare survivors build their 64-bit key, re-check it exactly and claim
// a buffer slot with a shared-memory atomic.  If a super-block overflows the buffer (cold start,
// adversarial input) its appends are discarded and it is replayed through the barrier-per-
// iteration path, so the result is exact for every input.
constexpr int TK_VEC_PER_THREAD = 2;                                  // float4 loads in flight
constexpr int TK_FAST_ITER = TK_THREADS * 4 * TK_VEC_PER_THREAD;      // 2048 scores
constexpr int TK_SUPER = 8 * TK_FAST_ITER;                            // 16384 scores
constexpr int TK_BOOT = 1024;                                         // bootstrap sample
constexpr int TK_COMPACT_AT = 640;                                    // compact early: sorts stay <= 1024 wide

__global__ void __launch_bounds__(TK_THREADS, 6)   // 40 registers: 6 blocks/SM hide the compaction phases
topk_stream_kernel(const float* __restrict__ scores, int64_t n, int64_t ld, int64_t chunk_len,
                   int k, uint64_t* __restrict__ out_keys /* [u][P][k] */) {
  __shared__ uint64_t s_keys[TK_CAP];
  __shared__ int s_count;
  __shared__ int s_ovf;
  __shared__ uint64_t s_tau;
  TopkState st{s_keys, &s_count, &s_tau};
  const int64_t u = blockIdx.y, p = blockIdx.x, P = gridDim.x;
  const int64_t lo = p * chunk_len;
  const int64_t hi = lo + chunk_len < n ? lo + chunk_len : n;
  const float* srow = scores + u * ld;
  if (threadIdx.x == 0) {
    s_count = 0;
    s_ovf = 0;
    s_tau = 0ull;
  }
  __syncthreads();

  // bootstrap: the first TK_CAP scores go straight into the buffer; one sort gives a threshold
  // good enough (k-th of 2048) for the streaming loop, instead of a cold start with tau = 0
  int64_t pos0 = lo;
  {
    const int64_t b_hi = lo + TK_BOOT < hi ? lo + TK_BOOT : hi;
    for (int64_t e = lo + threadIdx.x; e < b_hi; e += TK_THREADS)
      s_keys[e - lo] = make_key(srow[e], (uint32_t)e);
    if (threadIdx.x == 0) s_count = (int)(b_hi - lo);
    compact_select(st, k);
    pos0 = b_hi;
  }
  // super-blocks grow geometrically (2K, 4K, 8K, 16K scores) while the threshold tightens
  int64_t sb_len = TK_FAST_ITER;
  for (int64_t sb = pos0; sb < hi; sb += sb_len, sb_len = sb_len < TK_SUPER ? sb_len * 2 : TK_SUPER) {
    const int64_t sb_hi = sb + sb_len < hi ? sb + sb_len : hi;
    const int64_t full_hi = sb + (sb_hi - sb) / TK_FAST_ITER * TK_FAST_ITER;   // whole iterations
    const uint64_t tau = s_tau;
    const int count0 = s_count;
    // float form of the threshold: a score strictly below it can never enter the top-k
    const float tau_f = tau == 0ull ? -CUDART_INF_F : key_float((uint32_t)(tau >> 32));
    __syncthreads();   // everyone has the snapshot before anyone appends
#pragma unroll 2
    for (int64_t base = sb; base < full_hi; base += TK_FAST_ITER) {
      int4 r[TK_VEC_PER_THREAD];
#pragma unroll
      for (int q = 0; q < TK_VEC_PER_THREAD; ++q)
        r[q] = ld_stream16(srow + base + (q * TK_THREADS + threadIdx.x) * 4);
#pragma unroll
      for (int q = 0; q < TK_VEC_PER_THREAD; ++q) {
        const float v[4] = {__int_as_float(r[q].x), __int_as_float(r[q].y), __int_as_float(r[q].z),
                            __int_as_float(r[q].w)};
        // one test for the whole vector keeps the common case at ~1 instruction per score
        const bool any = !(v[0] < tau_f) | !(v[1] < tau_f) | !(v[2] < tau_f) | !(v[3] < tau_f);
        if (any) {                                      // rare after warm-up (NaN passes: ranks first)
          const int64_t e = base + (q * TK_THREADS + threadIdx.x) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!(v[j] < tau_f)) {
              const uint64_t key = make_key(v[j], (uint32_t)(e + j));
              if (key > tau) {
                const int pos = atomicAdd(&s_count, 1);
                if (pos < TK_CAP) s_keys[pos] = key;
                else s_ovf = 1;
              }
            }
          }
        }
      }
    }
    __syncthreads();
    const int ovf = s_ovf;
    const int cnt = s_count;
    __syncthreads();
    if (ovf) {                                           // block-uniform
      if (threadIdx.x == 0) {
        s_count = count0;                                // drop this super-block's appends
        s_ovf = 0;
      }
      __syncthreads();
      slow_range<0>(st, srow, nullptr, nullptr, sb, full_hi, k);
    } else if (cnt > (k <= 320 ? TK_COMPACT_AT : TK_CAP / 2 + k / 2)) {
      compact_select(st, k);
    }
    if (full_hi < sb_hi) slow_range<0>(st, srow, nullptr, nullptr, full_hi, sb_hi, k);   // ragged tail
  }
  compact_select(st, k);
  compact(st, k);   // ordered emit: a sort of <= k keys
  const int cnt = s_count < k ? s_count : k;
  uint64_t* o = out_keys + (u * P + p) * (int64_t)k;
  for (int i = threadIdx.x; i < k; i += TK_THREADS) o[i] = i < cnt ? s_keys[i] : 0ull;
}

__global__ void topk_emit_kernel(const uint64_t* __restrict__ keys, int64_t total,
                                 int64_t col_offset, float* __restrict__ out_scores,
                                 int64_t* __restrict__ out_idx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    if (key == 0ull) {
      out_scores[i] = -CUDART_INF_F;
      out_idx[i] = -1;
    } else {
      out_scores[i] = key_float((uint32_t)(key >> 32));
      out_idx[i] = (int64_t)(0xFFFFFFFFu - (uint32_t)key) + col_offset;
    }
  }
}

// one warp per query row: scores[u, id - col_offset] = -inf for the row's exclusion list
__global__ void mask_excluded_kernel(float* __restrict__ scores, int64_t u, int64_t n, int64_t ld,
                                     int64_t col_offset, const int64_t* __restrict__ offs,
                                     const int64_t* __restrict__ ids) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < u; r += nwarps) {
    for (int64_t e = offs[r] + lane; e < offs[r + 1]; e += 32) {
      const int64_t col = ids[e] - col_offset;
      if (col >= 0 && col < n) scores[r * ld + col] = -CUDART_INF_F;
    }
  }
}

// scores[u, j] = -inf where ids[u, j] is in row u's exclusion list or invalid (< 0 / >= n_valid);
// used on the re-scored candidate lists of the fused retrieval path (index.py:239-247 prefilter)
__global__ void mask_excluded_ids_kernel(float* __restrict__ scores, const int64_t* __restrict__ ids,
                                         int64_t u, int64_t c, int64_t ld, int64_t id_lo,
                                         int64_t id_hi, const int64_t* __restrict__ offs,
                                         const int64_t* __restrict__ excl) {
  const int64_t total = u * c;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / c, j = e - r * c;
    const int64_t id = ids[r * c + j];
    bool dead = id < id_lo || id >= id_hi;
    if (!dead && offs)
      for (int64_t x = offs[r]; x < offs[r + 1]; ++x) dead |= (excl[x] == id);
    if (dead) scores[r * ld + j] = -CUDART_INF_F;
  }
}

// group ids (U, kg) from the top-k over the group maxima -> the 16 catalog rows of every group:
// local row numbers for the re-score gather (clamped into [0, n)) and global ids (-1 = no such row)
__global__ void groups_to_rows_kernel(const int64_t* __restrict__ gi, int64_t total, int64_t n,
                                      int64_t row_offset, int64_t* __restrict__ cols,
                                      int64_t* __restrict__ ids) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = gi[e >> 4];
    const int64_t c = g * 16 + (e & 15);          // group g = catalog rows [16 g, 16 g + 16)
    const bool ok = g >= 0 && c < n;
    cols[e] = ok ? c : -1;                        // -1: the re-score gather reads nothing for it
    ids[e] = ok ? c + row_offset : -1;
  }
}

// ---- survivors of the scoring filter -> exact top-k (block per query) --------------------------------
// The survivors are ALL rows whose tensor-core score is >= the query's threshold.  They are gathered
// into shared memory as 64-bit keys (score desc, row asc: a total order, so the result does not depend on
// which lane stored what where), ONE radix select keeps the k_sel = k + max_excl best, the query's
// exclusion list is dropped (index.py:239-247) and the rest is sorted.  The reported scores are the
// tensor-core scores themselves: threshold, selection and ranking use one arithmetic, so no margin between
// two arithmetics is needed anywhere.
// Exactness: with k_sel selected and at most max_excl of them excluded, k non-excluded rows remain and every
// row left out ranks below all of them.  When fewer than k_sel rows survived (all were selected) the result
// is exact if k of them are not excluded -- or if the threshold was -inf (nothing was filtered): otherwise
// flag 4 tells the caller to take another path.
constexpr int FF_CAP = 8192;   // keys held in shared memory at a time (64 KB)
constexpr int FF_PCH = 4096;   // sub-buckets whose fill-count prefix sums are held at a time

// MSB-first radix select (8-bit digits) of the `want`-th largest of keys[0, n) (unique keys, want <= n);
// block-uniform call; s_hist[256], s_pick[2] are block scratch
__device__ __forceinline__ uint64_t radix_select_smem(const uint64_t* keys, int n, unsigned want,
                                                      unsigned* s_hist, unsigned* s_pick) {
  uint64_t prefix = 0ull, mask = 0ull;
  const int lane = threadIdx.x & 31;
  for (int shift = 56; shift >= 0; shift -= 8) {
    s_hist[threadIdx.x] = 0u;   // TK_THREADS == 256 bins
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += TK_THREADS) {
      const uint64_t key = keys[i];
      if ((key & mask) == prefix) atomicAdd(&s_hist[(unsigned)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // warp 0 walks the 256 bins from the top
      unsigned c[8], tot = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        c[b] = s_hist[255 - (lane * 8 + b)];
        tot += c[b];
      }
      unsigned incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      unsigned acc = incl - tot;
      if (acc < want && want <= incl) {
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if (acc < want && want <= acc + c[b]) {
            s_pick[0] = 255u - (unsigned)(lane * 8 + b);
            s_pick[1] = want - acc;
          }
          acc += c[b];
        }
      }
    }
    __syncthreads();
    prefix |= (uint64_t)s_pick[0] << shift;
    mask |= 0xFFull << shift;
    want = s_pick[1];
  }
  return prefix;
}

__global__ void __launch_bounds__(TK_THREADS)
filter_finalize_kernel(const FilterOut fo, const float* __restrict__ thresh, int64_t thresh_stride, int k_sel,
                       int k, int64_t row_offset, const int64_t* __restrict__ offs,
                       const int64_t* __restrict__ excl, int64_t max_excl, float* __restrict__ out_scores,
                       int64_t* __restrict__ out_idx, int32_t* __restrict__ flags) {
  extern __shared__ uint64_t ff_keys[];          // [FF_CAP] gathered keys | [TK_MAX_K] selected keys
  uint64_t* s_sel = ff_keys + FF_CAP;
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_pick[2];
  __shared__ int s_scan[TK_THREADS / 32];
  __shared__ int s_pref[FF_PCH + 1];
  __shared__ int s_fill, s_cnt, s_alive, s_total;
  const int64_t u = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ovf_raw = fo.o_count[u];
  const int n_ovf = (int)(ovf_raw < fo.ovf_cap ? ovf_raw : fo.ovf_cap);
  if (threadIdx.x == 0) {
    s_fill = 0;
    s_alive = 0;
    s_total = 0;
    int bad = ovf_raw > fo.ovf_cap ? 1 : 0;                            // survivors were dropped
    if (offs && offs[u + 1] - offs[u] > max_excl) bad |= 2;            // k_sel assumed fewer exclusions
    if (bad) atomicOr(flags, bad);
  }
  __syncthreads();

  // keep the k_sel best of ff_keys[0, s_fill) at the front (block-uniform)
  auto shrink = [&]() {
    const int fill = s_fill;
    if (fill <= k_sel) return;
    const uint64_t kth = radix_select_smem(ff_keys, fill, (unsigned)k_sel, s_hist, s_pick);
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < fill; i += TK_THREADS) {
      const uint64_t key = ff_keys[i];
      if (key >= kth) s_sel[atomicAdd(&s_cnt, 1)] = key;               // exactly k_sel keys (they are unique)
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k_sel; i += TK_THREADS) ff_keys[i] = s_sel[i];
    if (threadIdx.x == 0) s_fill = k_sel;
    __syncthreads();
  };
  // copy `total` keys, key j = src(j), behind the current fill; shrinks whenever the buffer runs full
  // (block-uniform; every thread takes every 256th key: independent loads, full memory parallelism)
  auto append = [&](int total, auto src) {
    int pos = 0;
    while (pos < total) {
      if (s_fill + (total - pos) > FF_CAP && s_fill > k_sel) shrink();
      const int fill = s_fill;
      int take = total - pos;
      if (take > FF_CAP - fill) take = FF_CAP - fill;
      for (int j = threadIdx.x; j < take; j += TK_THREADS) ff_keys[fill + j] = src(pos + j);
      __syncthreads();
      if (threadIdx.x == 0) s_fill = fill + take;
      __syncthreads();
      pos += take;
    }
  };

  // sub-buckets, FF_PCH at a time: exclusive prefix of their (clamped) fill counts in shared memory, then
  // the flattened (sub-bucket, entry) range is copied with a binary search per key
  int my_total = 0;
  for (int c0 = 0; c0 < fo.n_sub; c0 += FF_PCH) {
    const int nch = fo.n_sub - c0 < FF_PCH ? fo.n_sub - c0 : FF_PCH;
    const int it = (nch + TK_THREADS - 1) / TK_THREADS;                // consecutive sub-buckets per thread
    int cnt[FF_PCH / TK_THREADS];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < FF_PCH / TK_THREADS; ++j) {
      const int sb = threadIdx.x * it + j;
      int c = 0;
      if (j < it && sb < nch) {
        c = __ldcg(fo.b_count + u * fo.n_sub + c0 + sb);
        my_total += c;
        if (c > fo.cap_b) c = fo.cap_b;                                // beyond cap_b: in the overflow list
      }
      cnt[j] = c;
      mine += c;
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < TK_THREADS / 32; ++w) {
      if (w < warp) before += s_scan[w];
      total += s_scan[w];
    }
    int run = before + incl - mine;
#pragma unroll
    for (int j = 0; j < FF_PCH / TK_THREADS; ++j) {
      const int sb = threadIdx.x * it + j;
      if (j < it && sb < nch) s_pref[sb] = run;
      run += cnt[j];
    }
    if (threadIdx.x == 0) s_pref[nch] = total;
    __syncthreads();
    const float* bs = fo.b_scores + (u * fo.n_sub + c0) * fo.cap_b;
    const int32_t* br = fo.b_rows + (u * fo.n_sub + c0) * fo.cap_b;
    append(total, [&](int j) {
      int lo = 0, hi = nch;                                            // last sub-bucket with s_pref[sb] <= j
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_pref[mid] <= j) lo = mid;
        else hi = mid;
      }
      const long long at = (long long)lo * fo.cap_b + (j - s_pref[lo]);
      return make_key(__ldcg(bs + at), (uint32_t)__ldcg(br + at));
    });
  }
  {  // the overflow list
    const float* cs = fo.o_scores + u * fo.ovf_cap;
    const int32_t* cr = fo.o_rows + u * fo.ovf_cap;
    append(n_ovf, [&](int j) { return make_key(__ldcg(cs + j), (uint32_t)__ldcg(cr + j)); });
  }
  my_total = __reduce_add_sync(0xffffffffu, my_total);
  if (lane == 0 && my_total) atomicAdd(&s_total, my_total);
  shrink();
  const int c = s_fill;                                                // min(k_sel, survivors stored)
  // exclusion list, then the final order
  const int64_t x0 = offs ? offs[u] : 0, x1 = offs ? offs[u + 1] : 0;
  int n2 = 2;
  while (n2 < c) n2 <<= 1;
  int alive = 0;
  for (int i = threadIdx.x; i < n2; i += TK_THREADS) {
    uint64_t key = 0ull;
    if (i < c) {
      key = ff_keys[i];
      const int64_t id = (int64_t)(0xFFFFFFFFu - (uint32_t)key) + row_offset;
      bool dead = false;
      for (int64_t x = x0; x < x1; ++x) dead |= (excl[x] == id);
      // the low word of the final key is the GLOBAL id: shards then merge under one total order
      key = dead ? 0ull : ((key & 0xFFFFFFFF00000000ull) | (uint64_t)(0xFFFFFFFFu - (uint32_t)id));
      alive += dead ? 0 : 1;
    }
    s_sel[i] = key;
  }
  alive = __reduce_add_sync(0xffffffffu, alive);
  if (lane == 0 && alive) atomicAdd(&s_alive, alive);
  __syncthreads();
  if (threadIdx.x == 0 && s_total <= k_sel && s_alive < k && thresh[u * thresh_stride] > -CUDART_INF_F)
    atomicOr(flags, 4);   // too few non-excluded survivors to vouch for the rows below the threshold
  bitonic_desc(s_sel, n2);
  for (int i = threadIdx.x; i < k; i += TK_THREADS) {
    const uint64_t key = i < n2 ? s_sel[i] : 0ull;
    if (key == 0ull) {
      out_scores[u * k + i] = -CUDART_INF_F;
      out_idx[u * k + i] = -1;
    } else {
      out_scores[u * k + i] = key_float((uint32_t)(key >> 32));
      out_idx[u * k + i] = (int64_t)(0xFFFFFFFFu - (uint32_t)key);
    }
  }
}

// ---- k-th largest of every row (the thresholds of the scoring filter): block per row, 4-pass radix
//      select on the order-preserving 32-bit keys; the row (tens of KB) is re-read from L2 per pass ------
constexpr int KTH_THREADS = 1024;   // one block per row: with 256 threads the few hundred blocks of a search left
                                    // the GPU at 21 % warps active and every pass walked 76 dependent loads
__global__ void __launch_bounds__(KTH_THREADS)
kth_largest_kernel(const float* __restrict__ x, int64_t n, int64_t ld, int kth, float* __restrict__ out) {
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_pick[2];
  const float* row = x + (int64_t)blockIdx.x * ld;
  const int lane = threadIdx.x & 31;
  if (kth > n) {   // fewer values than the rank asked for: nothing bounds the k-th score from below
    if (threadIdx.x == 0) out[blockIdx.x] = -CUDART_INF_F;
    return;
  }
  uint32_t prefix = 0u, mask = 0u;
  unsigned want = (unsigned)kth;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (threadIdx.x < 256) s_hist[threadIdx.x] = 0u;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n; i += KTH_THREADS) {
      const uint32_t key = float_key(__ldg(row + i));
      if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned c[8], tot = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        c[b] = s_hist[255 - (lane * 8 + b)];
        tot += c[b];
      }
      unsigned incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      unsigned acc = incl - tot;
      if (acc < want && want <= incl) {
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if (acc < want && want <= acc + c[b]) {
            s_pick[0] = 255u - (unsigned)(lane * 8 + b);
            s_pick[1] = want - acc;
          }
          acc += c[b];
        }
      }
    }
    __syncthreads();
    prefix |= s_pick[0] << shift;
    mask |= 0xFFu << shift;
    want = s_pick[1];
  }
  if (threadIdx.x == 0) out[blockIdx.x] = key_float(prefix);
}

// ---- retrieval metrics: metrics.py:62-79, one thread per user ---------------------------------
__global__ void retrieval_metrics_kernel(const int64_t* __restrict__ rec, int64_t u, int64_t k,
                                         const int64_t* __restrict__ toffs,
                                         const int64_t* __restrict__ tids, int64_t top_k,
                                         float* __restrict__ out, uint8_t* __restrict__ valid) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= u) return;
  const int64_t t0 = toffs[r], t1 = toffs[r + 1];
  // number of DISTINCT targets (metrics.py:70 builds a set)
  int n_total = 0;
  for (int64_t a = t0; a < t1; ++a) {
    bool dup = false;
    for (int64_t b = t0; b < a; ++b) dup |= (tids[b] == tids[a]);
    n_total += !dup;
  }
  float* o = out + r * 7;
  if (n_total == 0) {  // metrics.py:62-63 returns {}
    valid[r] = 0;
    for (int j = 0; j < 7; ++j) o[j] = 0.f;
    return;
  }
  valid[r] = 1;
  // ranked list = rec padded to top_k (:65-68) followed by the targets it missed (:72); only the
  // first top_k positions matter, and a missed target can enter them only when k < top_k is
  // padded — padding ("" items) never hits, so hits come from rec[0..min(k,top_k)) alone.
  const int64_t kk = k < top_k ? k : top_k;
  double dcg = 0.0, ap_acc = 0.0, rr = 0.0;
  int n_hit = 0;
  long long pairs = 0;  // (hit before miss) pairs for the AUROC
  int misses_seen = 0;
  for (int64_t p = 0; p < kk; ++p) {
    const int64_t id = rec[r * k + p];
    bool hit = false;
    if (id >= 0)
      for (int64_t a = t0; a < t1; ++a) hit |= (tids[a] == id);
    if (hit) {
      ++n_hit;
      dcg += 1.0 / log2((double)p + 2.0);
      ap_acc += (double)n_hit / (double)(p + 1);
      if (rr == 0.0) rr = 1.0 / (double)(p + 1);
    } else {
      ++misses_seen;
    }
  }
  // AUROC over the top_k positions: fraction of (hit, miss) pairs ranked hit-first
  const int n_miss = (int)top_k - n_hit;  // padded tail positions are misses
  {
    int misses_before = 0;
    long long hits_after_miss = 0;
    for (int64_t p = 0; p < kk; ++p) {
      const int64_t id = rec[r * k + p];
      bool hit = false;
      if (id >= 0)
        for (int64_t a = t0; a < t1; ++a) hit |= (tids[a] == id);
      if (hit) hits_after_miss += misses_before;
      else ++misses_before;
    }
    pairs = (long long)n_hit * (long long)n_miss - hits_after_miss;
  }
  double idcg = 0.0;
  const int ideal = n_total < (int)top_k ? n_total : (int)top_k;
  for (int p = 0; p < ideal; ++p) idcg += 1.0 / log2((double)p + 2.0);
  o[0] = (float)(idcg > 0 ? dcg / idcg : 0.0);
  o[1] = (float)(n_hit > 0 ? ap_acc / n_hit : 0.0);
  o[2] = (float)((n_hit == 0 || n_miss == 0) ? 0.0 : (double)pairs / ((double)n_hit * n_miss));
  o[3] = (float)((double)n_hit / (double)top_k);
  o[4] = (float)((double)n_hit / (double)n_total);
  o[5] = n_hit > 0 ? 1.f : 0.f;
  o[6] = (float)rr;
  (void)misses_seen;
}

static int stream_blocks_per_sm() {
  // what the hardware really keeps resident (registers, shared memory), not an assumed figure: a
  // grid sized for 8 blocks/SM on a kernel that fits 4 runs a ragged third wave (measured: 0.54
  // instead of ~0.7 of the DRAM peak)
  static int occ = 0;
  if (occ == 0) {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, topk_stream_kernel, TK_THREADS, 0) != cudaSuccess ||
        v < 1)
      v = 4;
    occ = v;
  }
  return occ;
}

static int pick_chunks(int64_t u, int64_t n) {
  // at most ONE wave of resident blocks whenever the rows allow it (never a ragged extra wave);
  // chunks no shorter than 64k scores (and few enough partial lists that the merge stays short)
  const int64_t resident = (int64_t)sm_count() * stream_blocks_per_sm();
  int64_t want = resident / (u > 0 ? u : 1);
  int64_t maxp = (n + 65535) / 65536;
  if (want > maxp) want = maxp;
  if (want < 1) want = 1;
  if (want > 296) want = 296;
  return (int)want;
}

}  // namespace xr

using namespace xr;

extern "C" size_t xr_topk_workspace_bytes(int64_t u, int64_t n, int64_t k) {
  const int P = pick_chunks(u, n);
  // stage-1 partial keys + final keys
  return (size_t)(u > 0 ? u : 1) * (size_t)(P + 1) * (size_t)k * sizeof(uint64_t) + 256;
}

extern "C" int xr_topk(const float* scores, int64_t u, int64_t n, int64_t ld, int64_t k,
                       int64_t col_offset, float* out_scores, int64_t* out_idx, void* workspace,
                       size_t workspace_bytes, void* stream) {
  XR_CHECK_ARG(scores && out_scores && out_idx && workspace, "xr_topk: null pointer");
  XR_CHECK_ARG(u >= 0 && n >= 0 && ld >= n, "xr_topk: bad sizes");
  XR_CHECK_ARG(k >= 1 && k <= TK_MAX_K, "xr_topk: k must be in [1, %d]", TK_MAX_K);
  XR_CHECK_ARG(n < (1ll << 32) && u <= 65535, "xr_topk: n must be < 2^32 and u <= 65535 per call");
  XR_CHECK_ARG(workspace_bytes >= xr_topk_workspace_bytes(u, n, k), "xr_topk: workspace too small");
  if (u == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  const int P = pick_chunks(u, n);
  int64_t chunk = (n + P - 1) / P;
  chunk = (chunk + 3) / 4 * 4;  // chunk starts stay 16-byte aligned
  if (chunk < 4) chunk = 4;
  uint64_t* partial = (uint64_t*)workspace;
  uint64_t* final_keys = partial + (size_t)u * P * k;
  const bool vec = ((uintptr_t)scores % 16 == 0) && (ld % 4 == 0);
  dim3 grid((unsigned)P, (unsigned)u);
  uint64_t* stage1_out = P == 1 ? final_keys : partial;
  if (vec)
    topk_stream_kernel<<<grid, TK_THREADS, 0, s>>>(scores, n, ld, chunk, (int)k, stage1_out);
  else
    topk_generic_kernel<0><<<grid, TK_THREADS, 0, s>>>(scores, nullptr, nullptr, n, ld, chunk,
                                                       (int)k, stage1_out);
  XR_LAUNCH_CHECK("topk_stream");
  if (P > 1) {
    const int64_t pk = (int64_t)P * k;
    topk_generic_kernel<1><<<dim3(1, (unsigned)u), TK_THREADS, 0, s>>>(
        nullptr, nullptr, partial, pk, pk, pk, (int)k, final_keys);
    XR_LAUNCH_CHECK("topk_merge_chunks");
  }
  const int64_t total = u * k;
  topk_emit_kernel<<<(unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0,
                     s>>>(final_keys, total, col_offset, out_scores, out_idx);
  XR_LAUNCH_CHECK("topk_emit");
  return XR_OK;
}

extern "C" size_t xr_topk_merge_workspace_bytes(int64_t u, int64_t k) {
  return (size_t)(u > 0 ? u : 1) * (size_t)k * sizeof(uint64_t) + 256;
}

extern "C" int xr_topk_merge(const float* scores, const int64_t* ids, int64_t u, int64_t gk,
                             int64_t k, float* out_scores, int64_t* out_idx, void* workspace,
                             void* stream) {
  XR_CHECK_ARG(scores && ids && out_scores && out_idx && workspace, "xr_topk_merge: null pointer");
  XR_CHECK_ARG(u >= 0 && gk >= 0 && k >= 1 && k <= TK_MAX_K && u <= 65535,
               "xr_topk_merge: bad sizes");
  if (u == 0) return XR_OK;
  cudaStream_t s = as_stream(stream);
  uint64_t* final_keys = (uint64_t*)workspace;
  topk_generic_kernel<2><<<dim3(1, (unsigned)u), TK_THREADS, 0, s>>>(
      scores, ids, nullptr, gk, gk, gk > 0 ? gk : 1, (int)k, final_keys);
  XR_LAUNCH_CHECK("topk_merge");
  const int64_t total = u * k;
  topk_emit_kernel<<<(unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0,
                     s>>>(final_keys, total, 0, out_scores, out_idx);
  XR_LAUNCH_CHECK("topk_emit");
  return XR_OK;
}

extern "C" int xr_topk_merge_peers(const void* const* peer_scores_host, const void* const* peer_ids_host,
                                   int n_peers, int64_t u, int64_t k, float* out_scores,
                                   int64_t* out_idx, void* workspace, void* stream) {
  XR_CHECK_ARG(peer_scores_host && peer_ids_host && out_scores && out_idx && workspace,
               "xr_topk_merge_peers: null pointer");
  XR_CHECK_ARG(n_peers >= 1 && n_peers <= TK_MAX_PEERS, "xr_topk_merge_peers: 1 <= n_peers <= %d", TK_MAX_PEERS);
  XR_CHECK_ARG(u >= 0 && k >= 1 && k <= TK_MAX_K, "xr_topk_merge_peers: bad sizes");
  if (u == 0) return XR_OK;
  PeerLists pl{};
  for (int g = 0; g < n_peers; ++g) {
    XR_CHECK_ARG(peer_scores_host[g] && peer_ids_host[g], "xr_topk_merge_peers: null peer pointer");
    pl.scores[g] = (const float*)peer_scores_host[g];
    pl.ids[g] = (const int64_t*)peer_ids_host[g];
  }
  cudaStream_t s = as_stream(stream);
  uint64_t* final_keys = (uint64_t*)workspace;
  topk_merge_peers_kernel<<<(unsigned)u, TK_THREADS, 0, s>>>(pl, n_peers, (int)k, final_keys);
  XR_LAUNCH_CHECK("topk_merge_peers");
  const int64_t total = u * k;
  topk_emit_kernel<<<(unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0,
                     s>>>(final_keys, total, 0, out_scores, out_idx);
  XR_LAUNCH_CHECK("topk_emit");
  return XR_OK;
}

extern "C" int xr_mask_excluded(float* scores, int64_t u, int64_t n, int64_t ld,
                                int64_t col_offset, const int64_t* excl_offsets,
                                const int64_t* excl_ids, void* stream) {
  XR_CHECK_ARG(scores && excl_offsets && (excl_ids || u == 0), "xr_mask_excluded: null pointer");
  if (u == 0) return XR_OK;
  int64_t blocks = (u * 32 + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  mask_excluded_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      scores, u, n, ld, col_offset, excl_offsets, excl_ids);
  XR_LAUNCH_CHECK("mask_excluded");
  return XR_OK;
}

extern "C" int xr_mask_excluded_ids(float* scores, const int64_t* ids, int64_t u, int64_t c,
                                    int64_t ld, int64_t id_lo, int64_t id_hi,
                                    const int64_t* excl_offsets, const int64_t* excl_ids,
                                    void* stream) {
  XR_CHECK_ARG(scores && ids && ld >= c, "xr_mask_excluded_ids: bad arguments");
  if (u == 0 || c == 0) return XR_OK;
  int64_t blocks = (u * c + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  mask_excluded_ids_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      scores, ids, u, c, ld, id_lo, id_hi, excl_offsets, excl_ids);
  XR_LAUNCH_CHECK("mask_excluded_ids");
  return XR_OK;
}

extern "C" int xr_groups_to_rows(const int64_t* group_ids, int64_t u, int64_t kg, int64_t n,
                                 int64_t row_offset, int64_t* cols, int64_t* ids, void* stream) {
  XR_CHECK_ARG(group_ids && cols && ids && u >= 0 && kg >= 0 && n > 0, "xr_groups_to_rows: bad arguments");
  const int64_t total = u * kg * 16;
  if (total == 0) return XR_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  groups_to_rows_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(group_ids, total, n, row_offset,
                                                                         cols, ids);
  XR_LAUNCH_CHECK("groups_to_rows");
  return XR_OK;
}

extern "C" int xr_filter_finalize(int64_t u, int64_t n, const float* bucket_scores, const int32_t* bucket_rows,
                                  const int32_t* bucket_count, int64_t n_sub, int64_t cap_b,
                                  const float* ovf_scores, const int32_t* ovf_rows, const int32_t* ovf_count,
                                  int64_t ovf_cap, const float* thresh, int64_t thresh_stride, int64_t k_sel,
                                  int64_t k, int64_t row_offset, const int64_t* excl_offsets,
                                  const int64_t* excl_ids, int64_t max_excl, float* out_scores,
                                  int64_t* out_idx, int32_t* flags, void* stream) {
  XR_CHECK_ARG(bucket_scores && bucket_rows && bucket_count && ovf_scores && ovf_rows && ovf_count && thresh &&
                   out_scores && out_idx && flags,
               "xr_filter_finalize: null pointer");
  XR_CHECK_ARG(u >= 0 && n > 0 && n_sub >= 1 && cap_b >= 1 && ovf_cap >= 1 && k >= 1 && k_sel >= k &&
                   k_sel <= TK_MAX_K,
               "xr_filter_finalize: needs 1 <= k <= k_sel <= %d", TK_MAX_K);
  XR_CHECK_ARG(max_excl >= 0 && k_sel >= k + max_excl, "xr_filter_finalize: k_sel must be >= k + max_excl");
  XR_CHECK_ARG(row_offset >= 0 && row_offset + n <= (1ll << 32), "xr_filter_finalize: global ids must be < 2^32");
  XR_CHECK_ARG(!excl_offsets || excl_ids, "xr_filter_finalize: excl_offsets without excl_ids");
  if (u == 0) return XR_OK;
  const FilterOut fo{const_cast<float*>(bucket_scores), const_cast<int32_t*>(bucket_rows),
                     const_cast<int32_t*>(bucket_count), const_cast<float*>(ovf_scores),
                     const_cast<int32_t*>(ovf_rows), const_cast<int32_t*>(ovf_count), (int)n_sub, (int)cap_b,
                     (int)ovf_cap};
  constexpr int kSmem = (FF_CAP + TK_MAX_K) * 8;
  static bool configured = false;
  if (!configured) {
    XR_CUDA(cudaFuncSetAttribute(filter_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  filter_finalize_kernel<<<(unsigned)u, TK_THREADS, kSmem, as_stream(stream)>>>(
      fo, thresh, thresh_stride, (int)k_sel, (int)k, row_offset, excl_offsets, excl_ids, max_excl, out_scores,
      out_idx, flags);
  XR_LAUNCH_CHECK("filter_finalize");
  return XR_OK;
}

extern "C" int xr_kth_largest(const float* x, int64_t u, int64_t n, int64_t ld, int64_t kth, float* out,
                              void* stream) {
  XR_CHECK_ARG(x && out && u >= 0 && n >= 0 && ld >= n && kth >= 1, "xr_kth_largest: bad arguments");
  if (u == 0) return XR_OK;
  kth_largest_kernel<<<(unsigned)u, KTH_THREADS, 0, as_stream(stream)>>>(x, n, ld, (int)kth, out);
  XR_LAUNCH_CHECK("kth_largest");
  return XR_OK;
}

// ---- the whole local search as one call ------------------------------------------------------------
namespace xr {
constexpr int64_t kFilterOvfCap = 8192;   // overflow slots per query (sub-buckets that run full spill here)
constexpr int64_t kFilterTarget = 4096;   // survivors per query the sample stride aims at
constexpr int64_t kFilterSampleRows = 163840;   // rows the threshold sample may score on a small shard
constexpr int kFilterMargin = 28;         // rank positions of slack between the two score arithmetics
struct ScoreTopkPlan {
  int64_t kk, k_sel, stride, ld_s, n_sub, cap_b;
  size_t off_gmax, off_thr, off_bs, off_br, off_bc, off_os, off_or, off_oc, bytes;
};
static size_t al256(size_t x) { return (x + 255) / 256 * 256; }
static ScoreTopkPlan plan_score_topk(int64_t u, int64_t n, int64_t k, int64_t max_excl) {
  ScoreTopkPlan pl{};
  // threshold = the (k + slack)-th largest group maximum of the sample, whatever the exclusion lists:
  // ~ (k + slack) * stride rows survive, of which at most max_excl are excluded afterwards (the slack only
  // keeps a query with a few excluded top rows off the slow path)
  pl.kk = k + kFilterMargin;
  pl.k_sel = k + max_excl;
  // sample stride (in tiles of the scoring kernel): ~kk * stride survivors per query, and the sample keeps
  // at least 4 kk groups so that its kk-th maximum exists
  int64_t s = kFilterTarget / pl.kk;
  if (s > 32) s = 32;
  // small shards: a denser sample (>= ~160k rows stay cheap next to the launch latency) gives a tighter
  // threshold, so the filter epilogue and the finalize see proportionally fewer survivors
  if (s > n / kFilterSampleRows) s = n / kFilterSampleRows;
  if (s < 1) s = 1;
  const int64_t groups = (n + 15) / 16;
  while (s > 1 && groups / s < 4 * pl.kk) --s;
  if (s < 1) s = 1;
  pl.stride = s;
  pl.ld_s = (xr_score_groupmax_ld(u, n, s) + 1) / 2 * 2;
  // tiny catalogs: the threshold may be -inf (fewer sample groups than kk) and EVERY row survives
  const int64_t expect = groups < 4 * pl.kk * s ? n : pl.kk * s;
  xr_score_filter_layout(u, n, expect, &pl.n_sub, &pl.cap_b);
  size_t o = 0;
  pl.off_gmax = o; o += al256((size_t)u * pl.ld_s * 4);
  pl.off_thr = o;  o += al256((size_t)u * 4);
  pl.off_bs = o;   o += al256((size_t)u * pl.n_sub * pl.cap_b * 4);
  pl.off_br = o;   o += al256((size_t)u * pl.n_sub * pl.cap_b * 4);
  pl.off_bc = o;   o += al256((size_t)u * pl.n_sub * 4);
  pl.off_os = o;   o += al256((size_t)u * kFilterOvfCap * 4);
  pl.off_or = o;   o += al256((size_t)u * kFilterOvfCap * 4);
  pl.off_oc = o;   o += al256((size_t)u * 4);
  pl.bytes = o;
  return pl;
}
}  // namespace xr

extern "C" size_t xr_score_topk_workspace_bytes(int64_t u, int64_t n, int64_t k, int64_t max_excl) {
  if (u <= 0 || n <= 0 || k < 1 || max_excl < 0 || k + max_excl > TK_MAX_K) return 256;
  return plan_score_topk(u, n, k, max_excl).bytes;
}

extern "C" int xr_score_topk(const void* q, int64_t u, const void* catalog, int64_t n, int64_t dim, int64_t k,
                             int64_t row_offset, const int64_t* excl_offsets, const int64_t* excl_ids,
                             int64_t max_excl, float* out_scores, int64_t* out_idx, int32_t* flags,
                             void* workspace, size_t workspace_bytes, void* stream) {
  XR_CHECK_ARG(q && catalog && out_scores && out_idx && flags && workspace, "xr_score_topk: null pointer");
  XR_CHECK_ARG(u > 0 && u <= 65535 && n > 0 && k >= 1 && max_excl >= 0, "xr_score_topk: bad sizes (u <= 65535 per call)");
  XR_CHECK_ARG(k + max_excl <= TK_MAX_K, "xr_score_topk: k + max_excl must be <= %d", TK_MAX_K);
  XR_CHECK_ARG((uintptr_t)workspace % 256 == 0, "xr_score_topk: workspace must be 256-byte aligned");
  const ScoreTopkPlan pl = plan_score_topk(u, n, k, max_excl);
  XR_CHECK_ARG(workspace_bytes >= pl.bytes, "xr_score_topk: workspace too small");
  uint8_t* w = (uint8_t*)workspace;
  float* gmax = (float*)(w + pl.off_gmax);
  float* thr = (float*)(w + pl.off_thr);
  float* bs = (float*)(w + pl.off_bs);
  int32_t* br = (int32_t*)(w + pl.off_br);
  int32_t* bc = (int32_t*)(w + pl.off_bc);
  float* os = (float*)(w + pl.off_os);
  int32_t* orows = (int32_t*)(w + pl.off_or);
  int32_t* oc = (int32_t*)(w + pl.off_oc);
  cudaStream_t s = as_stream(stream);
  int rc;
  // 1. thresholds: the kk-th largest group maximum of a strided sample of the shard (-inf when the sample
  //    has fewer groups: every row then survives; the sub-buckets are sized for that regime)
  const int64_t ld_used = xr_score_groupmax_ld(u, n, pl.stride);
  if ((rc = xr_score_groupmax(q, u, catalog, n, dim, pl.stride, gmax, pl.ld_s, stream))) return rc;
  if ((rc = xr_kth_largest(gmax, u, ld_used, pl.ld_s, pl.kk, thr, stream))) return rc;
  // 2. one pass over the shard: survivors of the filter
  XR_CUDA(cudaMemsetAsync(oc, 0, (size_t)u * 4, s));
  if ((rc = xr_score_filter(q, u, catalog, n, dim, thr, 1, bs, br, bc, pl.n_sub, pl.cap_b, os, orows, oc,
                            kFilterOvfCap, stream)))
    return rc;
  // 3. survivors -> exact top-k
  return xr_filter_finalize(u, n, bs, br, bc, pl.n_sub, pl.cap_b, os, orows, oc, kFilterOvfCap, thr, 1, pl.k_sel, k,
                            row_offset, excl_offsets, excl_ids, max_excl, out_scores, out_idx, flags, stream);
}

extern "C" int xr_retrieval_metrics(const int64_t* rec, int64_t u, int64_t k,
                                    const int64_t* tgt_offsets, const int64_t* tgt_ids,
                                    int64_t top_k, float* out, uint8_t* valid, void* stream) {
  XR_CHECK_ARG(rec && tgt_offsets && out && valid, "xr_retrieval_metrics: null pointer");
  XR_CHECK_ARG(u >= 0 && k >= 0 && top_k >= 1, "xr_retrieval_metrics: bad sizes");
  if (u == 0) return XR_OK;
  retrieval_metrics_kernel<<<(unsigned)((u + 127) / 128), 128, 0, as_stream(stream)>>>(
      rec, u, k, tgt_offsets, tgt_ids, top_k, out, valid);
  XR_LAUNCH_CHECK("retrieval_metrics");
  return XR_OK;
}
