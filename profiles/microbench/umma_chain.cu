// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS mode, SWIZZLE_128B) as a
// function of N and of the number of independent accumulators interleaved in the issue stream.
// Answers: is a chain of dependent accumulating MMAs latency-bound at small N?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_chain umma_chain.cu && ./umma_chain
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transformer-recommenders_b200/csrc/sm100.cuh"
using namespace xr::sm100;

template <int N, int CHAINS, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) chain_kernel(long long* out, int iters) {
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
  if (warp == 1) {
    constexpr uint32_t idesc = xr::sm100::umma_idesc_bf16(128, N, 0, 0);
    const uint64_t a0 = umma_desc_sw128(base, 16, 1024);               // 128 rows x 128 B
    const uint64_t b0 = umma_desc_sw128(base + 16384, 16, 1024);       // up to 256 rows x 128 B
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
          for (int c = 0; c < CHAINS; ++c) {
            if (A_TMEM) umma_ts(tmem + c * N, tmem + 448, b0 + 2 * (i & 3), idesc, 1u);
            else umma_ss(tmem + c * N, a0 + 2 * (i & 3), b0 + 2 * (i & 3), idesc, 1u);
          }
        }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1, nullptr, 0);
      t1 = clock64();
    }
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int CHAINS, bool A_TMEM>
void run(const char* name, int iters) {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  auto k = chain_kernel<N, CHAINS, A_TMEM>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<148, 128, 100 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double per = avg / ((double)iters * CHAINS);
  printf("%-34s N=%3d chains=%d : %7.1f cycles/MMA  (ideal %3d)  -> %5.1f%% of pipe peak   [%s]\n", name, N,
         CHAINS, per, N / 2, 100.0 * (N / 2) / per, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int it = 4096;
  run<64, 1, false>("SS dependent chain", it);
  run<64, 2, false>("SS 2 interleaved accumulators", it);
  run<128, 1, false>("SS dependent chain", it);
  run<128, 2, false>("SS 2 interleaved accumulators", it);
  run<256, 1, false>("SS dependent chain", it);
  run<64, 1, true>("TS (A in TMEM) dependent chain", it);
  run<64, 2, true>("TS 2 interleaved accumulators", it);
  run<128, 1, true>("TS dependent chain", it);
  run<128, 3, true>("TS 3 interleaved accumulators", it);
  return 0;
}
