// Micro-benchmark: the fused kernel's per-tile MMA stream in isolation (no TMA, no epilogue):
//   24 x SS  M=128 N=64  K=16 (S = Q . Neg^T, both operands K-major, one accumulator)
//   12 x TS  M=128 N=128 K=16 (dQ += W . Neg, A from TMEM, B MN-major, three accumulators)
// issued by one thread, in the kernel's order and in a few alternatives.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_mix umma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transformer-recommenders_b200/csrc/sm100.cuh"
using namespace xr::sm100;

constexpr int QSUB = 128 * 64 * 2, SUB = 64 * 64 * 2;

// mode 0: S only   1: dQ only (MN-major B)   2: S then dQ per tile (kernel order)
// mode 3: dQ only with K-major B (N=128)     4: per pair: 8 S + 4 dQ interleaved
// mode 5: S then dQ, dQ with ks-outer order (3 accumulators round-robin)
__global__ void __launch_bounds__(128, 1) k(long long* out, int mode, int tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(8) unsigned long long dummy[8];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&dummy[i]), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_ptr), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  const uint32_t q_smem = base, ring = base + 6 * QSUB;
  constexpr uint32_t idesc_s = xr::sm100::umma_idesc_bf16(128, 64, 0, 0);
  constexpr uint32_t idesc_o = xr::sm100::umma_idesc_bf16(128, 128, 0, 1);
  constexpr uint32_t idesc_ok = xr::sm100::umma_idesc_bf16(128, 128, 0, 0);
  const uint64_t q_desc0 = umma_desc_sw128(q_smem, 16, 1024);
  const uint64_t ring_k = umma_desc_sw128(ring, 16, 1024);
  const uint64_t ring_mn = umma_desc_sw128(ring, SUB, 1024);
  if (warp == 1) {
    long long t0 = 0, t1 = 0;
    unsigned long long n0 = 0, n1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
      t0 = clock64();
      if (elect_one()) {
        for (int t = 0; t < tiles; ++t) {
          const int b = t & 1;
          const uint32_t g = (uint32_t)t * 3;
          auto S = [&](int pr) {
            const uint32_t s = (g + pr) & 7;
            const uint64_t a0 = q_desc0 + (uint64_t)(pr * ((2 * QSUB) >> 4));
            const uint64_t b0 = ring_k + (uint64_t)(s * ((2 * SUB) >> 4));
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_ss(tmem + 384 + b * 64, a0 + h * (QSUB >> 4) + 2 * kk, b0 + h * (SUB >> 4) + 2 * kk,
                        idesc_s, (pr | h | kk) ? 1u : 0u);
          };
          auto D = [&](int pr, int ks) {
            const uint32_t s = (g + pr) & 7;
            if (mode == 3)
              umma_ts(tmem + pr * 128, tmem + 384 + b * 64 + ks * 16, ring_k + (uint64_t)(s * ((2 * SUB) >> 4)) + 2 * ks,
                      idesc_ok, 1u);
            else
              umma_ts(tmem + pr * 128, tmem + 384 + b * 64 + ks * 16,
                      ring_mn + (uint64_t)(s * ((2 * SUB) >> 4) + ks * (2048 >> 4)), idesc_o, 1u);
          };
          if (mode == 0 || mode == 2 || mode == 5)
            for (int pr = 0; pr < 3; ++pr) S(pr);
          if (mode == 1 || mode == 2 || mode == 3)
            for (int pr = 0; pr < 3; ++pr)
              for (int ks = 0; ks < 4; ++ks) D(pr, ks);
          if (mode == 5)
            for (int ks = 0; ks < 4; ++ks)
              for (int pr = 0; pr < 3; ++pr) D(pr, ks);
          if (mode == 6) {   // kernel order WITH the kernel's commits
            for (int pr = 0; pr < 3; ++pr) S(pr);
            umma_commit(smem_u32(&dummy[0]));
            for (int pr = 0; pr < 3; ++pr) {
              for (int ks = 0; ks < 4; ++ks) D(pr, ks);
              umma_commit(smem_u32(&dummy[1 + pr]));
            }
          }
          if (mode == 7) {   // one commit per tile
            for (int pr = 0; pr < 3; ++pr) S(pr);
            for (int pr = 0; pr < 3; ++pr)
              for (int ks = 0; ks < 4; ++ks) D(pr, ks);
            umma_commit(smem_u32(&dummy[0]));
          }
          if (mode == 8) {   // commit after every pair of S too (forward-only release pattern)
            for (int pr = 0; pr < 3; ++pr) { S(pr); umma_commit(smem_u32(&dummy[4 + pr])); }
            umma_commit(smem_u32(&dummy[0]));
            for (int pr = 0; pr < 3; ++pr) {
              for (int ks = 0; ks < 4; ++ks) D(pr, ks);
              umma_commit(smem_u32(&dummy[1 + pr]));
            }
          }
          if (mode == 4)
            for (int pr = 0; pr < 3; ++pr) {
              S(pr);
              for (int ks = 0; ks < 4; ++ks) D(pr, ks);
            }
        }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1, nullptr, 0);
      t1 = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
    }
    if (threadIdx.x == 32) { out[blockIdx.x] = t1 - t0; out[148 + blockIdx.x] = (long long)(n1 - n0); }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// queue depth probe: issue n dependent SS N=64 MMAs back to back on an idle pipe and time the ISSUE
// (clock before the first, clock after the last instruction left the thread), then wait.
__global__ void __launch_bounds__(128, 1) qdepth(long long* out, int n) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_ptr), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  constexpr uint32_t idesc_s = xr::sm100::umma_idesc_bf16(128, 64, 0, 0);
  const uint64_t a0 = umma_desc_sw128(base, 16, 1024);
  const uint64_t b0 = umma_desc_sw128(base + 16384, 16, 1024);
  if (warp == 1) {
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      if (elect_one()) {
        t0 = clock64();
        for (int i = 0; i < n; ++i) umma_ss(tmem + 384, a0 + 2 * (i & 3), b0 + 2 * (i & 3), idesc_s, 1u);
        t1 = clock64();
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1, nullptr, 0);
      t2 = clock64();
      t0 = __shfl_sync(0xffffffffu, t0, 0); t1 = __shfl_sync(0xffffffffu, t1, 0);
    }
    if (threadIdx.x == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// The fused kernel's ISSUE STRUCTURE around the same MMA stream: warp-uniform loop, per pair an
// mbarrier wait (barriers completed in advance by a helper warp, so no wait ever blocks), fence,
// elect, 8 MMAs, commit, __syncwarp.  variant 0: as in the kernel; 1: waits hoisted (all three pair
// barriers polled before the first MMA of the tile); 2: no waits at all (upper bound).
__global__ void __launch_bounds__(128, 1) k_struct(long long* out, int variant, int tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar, full[8], empty[8], sfull[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(&sfull[0]), 1); mbar_init(smem_u32(&sfull[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_ptr), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  const uint32_t q_smem = base, ring = base + 6 * QSUB;
  constexpr uint32_t idesc_s = xr::sm100::umma_idesc_bf16(128, 64, 0, 0);
  const uint64_t q_desc0 = umma_desc_sw128(q_smem, 16, 1024);
  const uint64_t ring_k = umma_desc_sw128(ring, 16, 1024);
  if (warp == 2) {
    // helper "producer": keeps every full barrier one phase ahead (arrives as soon as empty fires)
    for (uint32_t g = 0; g < (uint32_t)tiles * 3 && (variant < 2 || variant == 6 || variant == 10); ++g) {
      const uint32_t s = g & 7;
      mbar_wait(smem_u32(&empty[s]), ((g / 8) & 1) ^ 1, nullptr, 0);
      if (lane == 0) mbar_arrive(smem_u32(&full[s]));
      __syncwarp();
    }
  } else if (warp == 1 && variant < 6) {
    long long t0 = clock64();
    uint32_t g = 0;
    for (int t = 0; t < tiles; ++t) {
      const int b = t & 1;
      if (variant == 1)
        for (int pr = 0; pr < 3; ++pr) mbar_wait(smem_u32(&full[(g + pr) & 7]), ((g + pr) / 8) & 1, nullptr, 0);
      tc_fence_after();
#pragma unroll 1
      for (int pr = 0; pr < 3; ++pr, ++g) {
        const uint32_t s = g & 7;
        if (variant == 0) mbar_wait(smem_u32(&full[s]), (g / 8) & 1, nullptr, 0);
        if (variant != 5) tc_fence_after();
        if (elect_one()) {
          const uint64_t a0 = q_desc0 + (uint64_t)(pr * ((2 * QSUB) >> 4));
          const uint64_t b0 = ring_k + (uint64_t)(s * ((2 * SUB) >> 4));
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ss(tmem + 384 + b * 64, a0 + h * (QSUB >> 4) + 2 * kk, b0 + h * (SUB >> 4) + 2 * kk,
                      idesc_s, (pr | h | kk) ? 1u : 0u);
          if (variant != 3 && variant != 5) umma_commit(smem_u32(&empty[s]));
        }
        if (variant != 4 && variant != 5) __syncwarp();
      }
      if (variant != 3 && variant != 5) {
        if (elect_one()) umma_commit(smem_u32(&sfull[b]));
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  } else if (warp == 3 && variant >= 6) {
    // ONE thread runs the whole issue loop: no elect / __syncwarp between MMA groups, and the
    // barrier a group needs was polled right after the previous group was issued (while those
    // MMAs were still queued), so the MMA stream never stops for a wait that is already satisfied
    if (variant >= 10 ? elect_one() : (lane == 0)) {
      long long t0 = clock64();
      uint32_t g = 0;
      if (variant == 6 || variant == 10) mbar_wait(smem_u32(&full[0]), 0, nullptr, 0);
      for (int t = 0; t < tiles; ++t) {
        const int b = t & 1;
#pragma unroll
        for (int pr = 0; pr < 3; ++pr, ++g) {
          const uint32_t s = g & 7;
          if (variant == 6 || variant == 8 || variant == 10) tc_fence_after();
          const uint64_t a0 = q_desc0 + (uint64_t)(pr * ((2 * QSUB) >> 4));
          const uint64_t b0 = ring_k + (uint64_t)(s * ((2 * SUB) >> 4));
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ss(tmem + 384 + b * 64, a0 + h * (QSUB >> 4) + 2 * kk, b0 + h * (SUB >> 4) + 2 * kk,
                      idesc_s, (pr | h | kk) ? 1u : 0u);
          if (variant != 9) umma_commit(smem_u32(&empty[s]));
          if (pr == 2 && variant != 9) umma_commit(smem_u32(&sfull[b]));
          // poll the NEXT group's barrier now
          const uint32_t gn = g + 1;
          if ((variant == 6 || variant == 10) && gn < (uint32_t)tiles * 3) mbar_wait(smem_u32(&full[gn & 7]), (gn / 8) & 1, nullptr, 0);
        }
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0, nullptr, 0);
      out[blockIdx.x] = clock64() - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// candidate structure for the fused kernel's score issuer: the whole role runs in ONE elected thread
template <bool WAITS>
__global__ void __launch_bounds__(128, 1) k_elected(long long* out, int tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar, full[8], empty[8], sfull[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(&sfull[0]), 1); mbar_init(smem_u32(&sfull[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_ptr), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  const uint32_t q_smem = base, ring = base + 6 * QSUB;
  constexpr uint32_t idesc_s = xr::sm100::umma_idesc_bf16(128, 64, 0, 0);
  const uint64_t q_desc0 = umma_desc_sw128(q_smem, 16, 1024);
  const uint64_t ring_k = umma_desc_sw128(ring, 16, 1024);
  const uint32_t full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]), sfull0 = smem_u32(&sfull[0]);
  if (warp == 2 && WAITS) {
    for (uint32_t g = 0; g < (uint32_t)tiles * 3; ++g) {
      const uint32_t s = g & 7;
      mbar_wait(empty0 + 8 * s, ((g / 8) & 1) ^ 1, nullptr, 0);
      if (lane == 0) mbar_arrive(full0 + 8 * s);
      __syncwarp();
    }
  } else if (warp == 1) {
    if (elect_one()) {
      long long t0 = clock64();
      uint32_t g = 0;
      if (WAITS) mbar_wait(full0, 0, nullptr, 0);
      for (int t = 0; t < tiles; ++t) {
        const int b = t & 1;
#pragma unroll
        for (int pr = 0; pr < 3; ++pr, ++g) {
          const uint32_t s = g & 7;
          if (WAITS) tc_fence_after();
          const uint64_t a0 = q_desc0 + (uint64_t)(pr * ((2 * QSUB) >> 4));
          const uint64_t b0 = ring_k + (uint64_t)(s * ((2 * SUB) >> 4));
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ss(tmem + 384 + b * 64, a0 + h * (QSUB >> 4) + 2 * kk, b0 + h * (SUB >> 4) + 2 * kk,
                      idesc_s, (pr | h | kk) ? 1u : 0u);
          umma_commit(empty0 + 8 * s);
          if (pr == 2) umma_commit(sfull0 + 8 * b);
          const uint32_t gn = g + 1;
          if (WAITS && gn < (uint32_t)tiles * 3) mbar_wait(full0 + 8 * (gn & 7), (gn / 8) & 1, nullptr, 0);
        }
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0, nullptr, 0);
      out[blockIdx.x] = clock64() - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  {
    long long* d; cudaMalloc(&d, 148 * 8);
    const int smem = 6 * QSUB + 16 * SUB + 1024;
    cudaFuncSetAttribute(k_elected<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_elected<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int w = 0; w < 2; ++w) {
      if (w) k_elected<true><<<148, 128, smem>>>(d, 400); else k_elected<false><<<148, 128, smem>>>(d, 400);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("S-only, whole issuer role in one ELECTED thread, %s: %7.1f cycles/tile  [%s]\n",
             w ? "polled waits + commits" : "commits only          ", avg / 400, cudaGetErrorString(e));
    }
  }
  {
    long long* d; cudaMalloc(&d, 148 * 8);
    const int smem = 6 * QSUB + 16 * SUB + 1024;
    cudaFuncSetAttribute(k_struct, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* nm[] = {"kernel issue structure (wait/elect per pair)", "waits hoisted to the tile start",
                        "no waits, no producer (commits + elect per pair)", "no waits, no commits (elect + syncwarp per pair)",
                        "no waits, commits, no syncwarp per pair", "elect per pair only (no fence/commit/syncwarp)",
                        "ONE issuing thread, next barrier polled behind each MMA group",
                        "ONE thread: MMAs + commits only", "ONE thread: fence + MMAs + commits", "ONE thread: MMAs only", "ONE ELECTED thread: polled waits + fence + MMAs + commits",
                        "ONE ELECTED thread: MMAs + commits"};
    for (int v = 6; v < 12; ++v) {
      k_struct<<<148, 128, smem>>>(d, v, 400);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("S-only, %-46s: %7.1f cycles/tile  [%s]\n", nm[v], avg / 400, cudaGetErrorString(e));
    }
  }
  {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(qdepth, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int n : {1, 2, 4, 8, 12, 16, 24, 32, 48, 64, 96, 128}) {
      qdepth<<<1, 128, 64 * 1024>>>(d, n);
      cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("issue %3d MMAs (N=64 SS): issue took %6lld cycles, all complete after %6lld cycles\n", n, h[0], h[1]);
    }
  }

  long long* d; cudaMalloc(&d, 2 * 148 * 8);
  const int smem = 6 * QSUB + 16 * SUB + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"S only (24 SS N=64)", "dQ only (12 TS N=128, B MN-major)", "S then dQ (kernel order)",
                         "dQ only, B K-major", "per pair: 8 S + 4 dQ", "S then dQ, dQ ks-outer",
                         "kernel order + 4 commits/tile", "kernel order + 1 commit/tile", "kernel order + 7 commits/tile"};
  for (int grid : {148})
    for (int mode = 0; mode < 9; ++mode) {
      const int tiles = 4000;
      k<<<grid, 128, smem>>>(d, mode, tiles);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[296]; cudaMemcpy(h, d, 296 * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
      double ns = 0; for (int i = 0; i < grid; ++i) ns += h[148 + i]; ns /= grid;
      printf("grid=%3d %-36s: %7.1f cycles/tile  %7.1f ns/tile -> %6.0f MHz  [%s]\n", grid, names[mode], avg / tiles, ns / tiles, avg / ns * 1e3, cudaGetErrorString(e));
    }
  return 0;
}
