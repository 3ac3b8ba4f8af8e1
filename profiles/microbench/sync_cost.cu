// Micro-benchmark: cost of the synchronisation primitives the fused kernel chains per tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o sync_cost sync_cost.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transformer-recommenders_b200/csrc/sm100.cuh"
using namespace xr::sm100;

// mode 0: commit -> wait round trip, no MMA in flight (latency of one tcgen05.commit)
// mode 1: N back-to-back commits to N barriers, wait for the last (throughput of commits)
// mode 2: one N=64 MMA then commit -> wait (latency incl. one MMA)
// mode 3: tcgen05.fence::after_thread_sync x N
// mode 4: try_wait on an already completed barrier x N
// mode 5: ping-pong between two warps through plain mbarrier arrive / try_wait (hand-off latency)
// mode 6: ping-pong where one direction is a tcgen05.commit (issuer -> epilogue hand-off)
__global__ void __launch_bounds__(640, 1) k(long long* out, int mode, int iters, int pollers) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bars[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_ptr), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  constexpr uint32_t idesc = xr::sm100::umma_idesc_bf16(128, 64, 0, 0);
  const uint64_t a0 = umma_desc_sw128(base, 16, 1024);
  const uint64_t b0 = umma_desc_sw128(base + 16384, 16, 1024);
  long long t0 = 0, t1 = 0;
  if (warp >= 4) {
    // background pollers: spin on a barrier that only completes when the measuring warp is done
    if (warp - 4 < pollers) mbar_wait(smem_u32(&bars[40]), 0, nullptr, 0);
  } else if (warp == 1) {
    if (mode == 0 || mode == 2) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (elect_one()) {
          if (mode == 2) umma_ss(tmem, a0, b0, idesc, 0u);
          umma_commit(smem_u32(&bars[0]));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bars[0]), i & 1, nullptr, 0);
      }
      t1 = clock64();
    } else if (mode == 1) {
      t0 = clock64();
      for (int rep = 0; rep < iters; ++rep) {
        if (elect_one())
          for (int i = 0; i < 32; ++i) umma_commit(smem_u32(&bars[i]));
        __syncwarp();
        mbar_wait(smem_u32(&bars[31]), rep & 1, nullptr, 0);
      }
      t1 = clock64();
    } else if (mode == 3) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) tc_fence_after();
      t1 = clock64();
    } else if (mode == 4) {
      if (lane == 0) mbar_arrive(smem_u32(&bars[0]));
      __syncwarp();
      t0 = clock64();
      int acc = 0;
      for (int i = 0; i < iters; ++i) acc += mbar_try_wait(smem_u32(&bars[0]), 0);
      t1 = clock64();
      if (acc == -1) out[1] = acc;
    } else if (mode == 5 || mode == 6) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (mode == 5) { if (lane == 0) mbar_arrive(smem_u32(&bars[0])); }
        else if (elect_one()) umma_commit(smem_u32(&bars[0]));
        __syncwarp();
        mbar_wait(smem_u32(&bars[1]), i & 1, nullptr, 0);
      }
      t1 = clock64();
    }
    if (lane == 0) out[0] = t1 - t0;
    if (lane == 0) mbar_arrive(smem_u32(&bars[40]));
  } else if (warp == 2 && (mode == 5 || mode == 6)) {
    for (int i = 0; i < iters; ++i) {
      mbar_wait(smem_u32(&bars[0]), i & 1, nullptr, 0);
      if (lane == 0) mbar_arrive(smem_u32(&bars[1]));
      __syncwarp();
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[] = {"commit -> wait round trip (idle pipe)", "32 back-to-back commits, wait last (per commit)",
                         "1 MMA(N=64) + commit -> wait", "tcgen05.fence::after_thread_sync",
                         "try_wait on completed barrier", "warp ping-pong, mbarrier both ways (round trip)",
                         "warp ping-pong, commit one way (round trip)"};
  for (int pollers : {0, 4, 16})
  for (int mode = 0; mode < 7; ++mode) {
    const int iters = 1000;
    if (pollers && mode != 0 && mode != 5 && mode != 6) continue;
    printf("[%2d polling warps] ", pollers);
    k<<<1, 640, 64 * 1024>>>(d, mode, iters, pollers);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / iters / (mode == 1 ? 32 : 1);
    printf("%-52s: %8.1f cycles  [%s]\n", names[mode], per, cudaGetErrorString(e));
  }
  return 0;
}
