// Micro-benchmark: latency and per-SM throughput of the TMA stream the fused kernel depends on.
// 148 persistent CTAs each stream 16 KB "pairs" (two 64x64 bf16 boxes, SWIZZLE_128B) of an
// L2-resident [C x 384] bf16 matrix through a ring of R slots; the consumer releases a slot as
// soon as it is full, so R x 16 KB are in flight per SM.  Reports cycles per pair and the issue ->
// complete latency, for three access patterns:
//   lockstep : every CTA reads the same tile sequence at the same time (the fused kernel before
//              the rotation), rotated : start tile depends on the CTA, disjoint : own rows per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu && ./tma_stream
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../transformer-recommenders_b200/csrc/sm100.cuh"
using namespace xr::sm100;

constexpr int PAIR_BYTES = 16384;

template <int R>
__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap tmap, int n_tiles, int pairs_per_cta, int mode,
              long long* out_cycles, long long* out_lat) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ __align__(8) unsigned long long bars[2 * R];
  __shared__ long long ts[R];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[R + s]), 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  int t_start = 0, t_lo = 0, t_n = n_tiles;
  if (mode == 1) t_start = (blockIdx.x * 29) % n_tiles;
  if (mode == 2) {
    t_n = n_tiles / gridDim.x;
    t_lo = blockIdx.x * t_n;
  }
  const long long t0 = clock64();
  if (warp == 0) {
    for (int g = 0; g < pairs_per_cta; ++g) {
      const int s = g % R;
      mbar_wait(smem_u32(&bars[R + s]), ((g / R) & 1) ^ 1, nullptr, 0);
      if (lane == 0) {
        const int tile = t_lo + (t_start + g / 3) % t_n;
        const int pr = g % 3;
        ts[s] = clock64();
        mbar_expect_tx(smem_u32(&bars[s]), PAIR_BYTES);
        tma_load_2d(base + s * PAIR_BYTES, &tmap, smem_u32(&bars[s]), pr * 128, tile * 64);
        tma_load_2d(base + s * PAIR_BYTES + 8192, &tmap, smem_u32(&bars[s]), pr * 128 + 64, tile * 64);
      }
      __syncwarp();
    }
  } else {
    long long lat = 0;
    for (int g = 0; g < pairs_per_cta; ++g) {
      const int s = g % R;
      mbar_wait(smem_u32(&bars[s]), (g / R) & 1, nullptr, 0);
      if (lane == 0) {
        lat += clock64() - ts[s];
        mbar_arrive(smem_u32(&bars[R + s]));
      }
      __syncwarp();
    }
    if (lane == 0) {
      out_cycles[blockIdx.x] = clock64() - t0;
      out_lat[blockIdx.x] = lat;
    }
  }
}


// ---- cluster variant: CS CTAs share the stream; CTA r loads 1/CS of every pair and multicasts it
//      to all CTAs of the cluster.  A slot is refilled when every CTA of the cluster released it.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                               int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int R, int CS, bool USE_MC = true, bool REMOTE = true>
__global__ void __launch_bounds__(64, 1)
stream_mc_kernel(const __grid_constant__ CUtensorMap tmap /* box 64 cols x (128/CS) rows */, int n_tiles,
                 int pairs_per_cta, long long* out_cycles, long long* out_lat) {
  // here a "pair" is a [128 rows x 64 cols] 16 KB block made of CS row slices
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ __align__(8) unsigned long long bars[2 * R];
  __shared__ long long ts[R];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[R + s]), CS);
    }
    fence_barrier_init();
  }
  cluster_sync_all();
  const long long t0 = clock64();
  if (warp == 0) {
    for (int g = 0; g < pairs_per_cta; ++g) {
      const int s = g % R;
      mbar_wait(smem_u32(&bars[R + s]), ((g / R) & 1) ^ 1, nullptr, 0);
      if (lane == 0) {
        const int tile = (g / 6) % n_tiles, kb = g % 6;
        ts[s] = clock64();
        mbar_expect_tx(smem_u32(&bars[s]), PAIR_BYTES);
        if (USE_MC)
          tma_load_2d_mc(base + s * PAIR_BYTES + rank * (PAIR_BYTES / CS), &tmap, smem_u32(&bars[s]),
                         kb * 64, tile * 128 + rank * (128 / CS), (uint16_t)((1u << CS) - 1));
        else
          tma_load_2d(base + s * PAIR_BYTES, &tmap, smem_u32(&bars[s]), kb * 64, tile * 128);
      }
      __syncwarp();
    }
  } else {
    long long lat = 0;
    for (int g = 0; g < pairs_per_cta; ++g) {
      const int s = g % R;
      mbar_wait(smem_u32(&bars[s]), (g / R) & 1, nullptr, 0);
      if (lane == 0) lat += clock64() - ts[s];
      if (REMOTE) {
        if (lane < CS) mbar_arrive_remote(smem_u32(&bars[R + s]), lane);
      } else if (lane == 0) {
        mbar_arrive(smem_u32(&bars[R + s]));
      }
      __syncwarp();
    }
    if (lane == 0) {
      out_cycles[blockIdx.x] = clock64() - t0;
      out_lat[blockIdx.x] = lat;
    }
  }
  cluster_sync_all();
}

// ---- 1-D bulk copies (cp.async.bulk, no tensor map): 16 KB contiguous per pair ------------------
template <int R>
__global__ void __launch_bounds__(64, 1)
stream_bulk_kernel(const uint8_t* src, long long bytes_total, int pairs_per_cta, long long* out_cycles,
                   long long* out_lat) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ __align__(8) unsigned long long bars[2 * R];
  __shared__ long long ts[R];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[R + s]), 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const long long npairs_total = bytes_total / PAIR_BYTES;
  const long long t0 = clock64();
  if (warp == 0) {
    for (int g = 0; g < pairs_per_cta; ++g) {
      const int s = g % R;
      mbar_wait(smem_u32(&bars[R + s]), ((g / R) & 1) ^ 1, nullptr, 0);
      if (lane == 0) {
        const uint8_t* p = src + ((long long)(g + blockIdx.x * 29) % npairs_total) * PAIR_BYTES;
        ts[s] = clock64();
        mbar_expect_tx(smem_u32(&bars[s]), PAIR_BYTES);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                base + s * PAIR_BYTES),
            "l"(p), "r"(PAIR_BYTES), "r"(smem_u32(&bars[s]))
            : "memory");
      }
      __syncwarp();
    }
  } else {
    long long lat = 0;
    for (int g = 0; g < pairs_per_cta; ++g) {
      const int s = g % R;
      mbar_wait(smem_u32(&bars[s]), (g / R) & 1, nullptr, 0);
      if (lane == 0) {
        lat += clock64() - ts[s];
        mbar_arrive(smem_u32(&bars[R + s]));
      }
      __syncwarp();
    }
    if (lane == 0) {
      out_cycles[blockIdx.x] = clock64() - t0;
      out_lat[blockIdx.x] = lat;
    }
  }
}

static CUtensorMap make_map(void* base, long long rows, int box_rows = 64) {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<xr::PFN_encodeTiled>(ptr);
  CUtensorMap m;
  const cuuint64_t gdim[2] = {384, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {768};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

template <int R>
void run(const CUtensorMap& map, int n_tiles, int mode, int grid) {
  long long *dc, *dl;
  cudaMalloc(&dc, grid * 8);
  cudaMalloc(&dl, grid * 8);
  const int pairs = 3 * 400;
  auto k = stream_kernel<R>;
  const int smem = R * PAIR_BYTES + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) k<<<grid, 64, smem>>>(map, n_tiles, pairs, mode, dc, dl);
  cudaError_t e = cudaDeviceSynchronize();
  long long hc[148], hl[148];
  cudaMemcpy(hc, dc, grid * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(hl, dl, grid * 8, cudaMemcpyDeviceToHost);
  double c = 0, l = 0;
  for (int i = 0; i < grid; ++i) { c += hc[i]; l += hl[i]; }
  c /= grid; l /= grid;
  const char* names[] = {"lockstep", "rotated ", "disjoint"};
  printf("%s grid=%3d ring=%d pairs (%3d KB in flight): %7.1f cycles/pair  %6.1f B/cyc/SM  latency %7.0f cyc  [%s]\n",
         names[mode], grid, R, R * 16, c / pairs, PAIR_BYTES / (c / pairs), l / pairs, cudaGetErrorString(e));
  cudaFree(dc); cudaFree(dl);
}

template <int R, int CS, bool USE_MC = true, bool REMOTE = true>
void run_mc(void* d, long long rows, int grid) {
  CUtensorMap map = make_map(d, rows, 128 / CS);
  long long *dc, *dl;
  cudaMalloc(&dc, grid * 8);
  cudaMalloc(&dl, grid * 8);
  const int pairs = 6 * 200;
  auto k = stream_mc_kernel<R, CS, USE_MC, REMOTE>;
  const int smem = R * PAIR_BYTES + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const int nt = (int)(rows / 128);
  cudaError_t e = cudaSuccess;
  for (int rep = 0; rep < 2; ++rep) e = cudaLaunchKernelEx(&cfg, k, map, nt, pairs, dc, dl);
  cudaError_t e2 = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = e2;
  long long hc[148], hl[148];
  cudaMemcpy(hc, dc, grid * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(hl, dl, grid * 8, cudaMemcpyDeviceToHost);
  double c = 0, l = 0;
  for (int i = 0; i < grid; ++i) { c += hc[i]; l += hl[i]; }
  c /= grid; l /= grid;
  printf("mc=%d remote_arrive=%d cluster=%d grid=%3d ring=%d: %7.1f cycles/16KB-block per CTA  %6.1f B/cyc/SM landed  latency %7.0f cyc  [%s]\n",
         (int)USE_MC, (int)REMOTE, CS, grid, R, c / pairs, PAIR_BYTES / (c / pairs), l / pairs, cudaGetErrorString(e));
  cudaFree(dc); cudaFree(dl);
}

template <int R>
void run_bulk(void* d, long long bytes, int grid) {
  long long *dc, *dl;
  cudaMalloc(&dc, grid * 8);
  cudaMalloc(&dl, grid * 8);
  const int pairs = 1200;
  auto k = stream_bulk_kernel<R>;
  const int smem = R * PAIR_BYTES + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) k<<<grid, 64, smem>>>((const uint8_t*)d, bytes, pairs, dc, dl);
  cudaError_t e = cudaDeviceSynchronize();
  long long hc[148], hl[148];
  cudaMemcpy(hc, dc, grid * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(hl, dl, grid * 8, cudaMemcpyDeviceToHost);
  double c = 0, l = 0;
  for (int i = 0; i < grid; ++i) { c += hc[i]; l += hl[i]; }
  c /= grid; l /= grid;
  printf("1-D bulk 16KB grid=%3d ring=%d: %7.1f cycles/pair  %6.1f B/cyc/SM  latency %7.0f cyc  [%s]\n", grid, R,
         c / pairs, PAIR_BYTES / (c / pairs), l / pairs, cudaGetErrorString(e));
  cudaFree(dc); cudaFree(dl);
}

int main() {
  const long long rows = 12672;   // 198 tiles of 64 rows (the bench workload's pool)
  void* d;
  cudaMalloc(&d, rows * 768);
  cudaMemset(d, 0, rows * 768);
  CUtensorMap map = make_map(d, rows);
  const int nt = (int)(rows / 64);
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (grid == 1 && mode) continue;
      run<1>(map, nt, mode, grid);
      run<2>(map, nt, mode, grid);
      run<3>(map, nt, mode, grid);
      run<4>(map, nt, mode, grid);
      run<6>(map, nt, mode, grid);
      run<8>(map, nt, mode, grid);
      run<12>(map, nt, mode, grid);
    }
  }
  run_mc<4, 1, false, false>(d, rows, 148);
  run_mc<4, 1, false, true>(d, rows, 148);
  run_mc<4, 1, true, false>(d, rows, 148);
  run_mc<4, 1>(d, rows, 148);
  run_mc<8, 1>(d, rows, 148);
  run_mc<4, 2>(d, rows, 148);
  run_mc<8, 2>(d, rows, 148);
  run_mc<4, 4>(d, rows, 148);
  run_mc<8, 4>(d, rows, 148);
  run_bulk<2>(d, rows * 768, 148);
  run_bulk<4>(d, rows * 768, 148);
  run_bulk<8>(d, rows * 768, 148);
  run_bulk<8>(d, rows * 768, 1);
  return 0;
}
