// Micro-benchmark: issue rate / throughput of tcgen05.mma for the shapes the attend kernel uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cosnet_b200/csrc tools/umma_bench.cu -o tools/umma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace coattn;

// mode: 0 = SS, 1 = TS.  N: 64/128/256.  alt: number of distinct D tiles cycled through (1 = dependent chain)
template <int MODE, int N, int ALT>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_16(128, N, true);
    const uint64_t adesc = make_sdesc_k_sw128(smem_u32(smem));
    const uint64_t bdesc = make_sdesc_k_sw128(smem_u32(smem + 32768));
    uint32_t phase = 0;
    // warm-up
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) umma_ss(tmem, adesc, bdesc, idesc, 1); else umma_ts(tmem, tmem + 448, bdesc, idesc, 1);
    }
    umma_commit(&bar); mbar_wait(&bar, phase, 1); phase ^= 1;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t d = tmem + ((it * 16 + k) % ALT) * N;
        if (MODE == 0) umma_ss(d, adesc + 2 * (k & 3), bdesc + 2 * (k & 3), idesc, 1);
        else umma_ts(d, tmem + 448 + 8 * (k & 3), bdesc + 2 * (k & 3), idesc, 1);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar); mbar_wait(&bar, phase, 2); phase ^= 1;
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// attend-like pattern per "tile": 16 x (TS, N=NS) into S buffer, CS commits, 4*(64/16)/(4) ... PV: NPV x (TS, N=256), CP commits
template <int NS, int CS, int CP>
__global__ void __launch_bounds__(128, 1) pattern(long long* out, int tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, junk[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&junk[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc_s = make_idesc_16(128, NS, true);
    constexpr uint32_t idesc_o = make_idesc_16(128, 256, true);
    const uint64_t bdesc = make_sdesc_k_sw128(smem_u32(smem + 32768));
    const long long t0 = clock64();
    for (int it = 0; it < tiles; ++it) {
      const uint32_t tS = tmem + 256 + (it & 1) * NS;
      constexpr int KS = 16 * 64 / NS;   // same flops per tile regardless of NS: NS=64 -> 16 mma, NS=128 -> 16 (K=256) but tile is 2x
#pragma unroll
      for (int k = 0; k < 16; ++k) umma_ts(tS, tmem + 384 + 8 * (k & 7), bdesc + 2 * (k & 3), idesc_s, k > 0);
#pragma unroll
      for (int c = 0; c < CS; ++c) umma_commit(&junk[c]);
#pragma unroll
      for (int k = 0; k < NS / 16; ++k) umma_ts(tmem, tS + 8 * (k & 3), bdesc + 2 * (k & 3), idesc_o, 1);
#pragma unroll
      for (int c = 0; c < CP; ++c) umma_commit(&junk[2 + c]);
      (void)KS;
    }
    umma_commit(&bar); mbar_wait(&bar, 0, 2);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t2 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
template <int NS, int CS, int CP>
void run_pattern(const char* name) {
  long long* d; cudaMalloc(&d, 16);
  const int tiles = 64;
  auto k = pattern<NS, CS, CP>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<148, 128, 100 * 1024>>>(d, tiles);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-40s %8.1f cyc/tile (ideal %d) %s\n", name, (double)h[0] / tiles, 16 * NS / 2 + NS / 16 * 128, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

template <int MODE, int N, int ALT>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 16);
  const int iters = 64;
  auto k = bench<MODE, N, ALT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<grid, 128, 100 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double n_mma = iters * 16.0;
  printf("%-28s grid %3d  issue %7.1f cyc/mma   complete %7.1f cyc/mma   (ideal %5.1f)  %s\n", name, grid, h[0] / n_mma,
         h[1] / n_mma, 128.0 * N / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run_pattern<64, 0, 0>("pattern NS=64  no commits");
  run_pattern<64, 1, 1>("pattern NS=64  1+1 commits");
  run_pattern<64, 2, 2>("pattern NS=64  2+2 commits");
  run_pattern<128, 0, 0>("pattern NS=128 no commits");
  run_pattern<128, 1, 1>("pattern NS=128 1+1 commits");
  run_pattern<128, 2, 2>("pattern NS=128 2+2 commits");
  for (int grid : {148}) {
    run<0, 64, 1>("SS N=64  same D", grid);
    run<0, 64, 2>("SS N=64  2 D tiles", grid);
    run<1, 64, 1>("TS N=64  same D", grid);
    run<1, 64, 2>("TS N=64  2 D tiles", grid);
    run<0, 128, 1>("SS N=128 same D", grid);
    run<1, 128, 1>("TS N=128 same D", grid);
    run<1, 128, 2>("TS N=128 2 D tiles", grid);
    run<0, 256, 1>("SS N=256 same D", grid);
    run<1, 256, 1>("TS N=256 same D", grid);
  }
  return 0;
}
