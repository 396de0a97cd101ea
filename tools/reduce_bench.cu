// Micro-benchmark for the "single S" design of the attend kernel (VERDICT r1, item 3 i; SURVEY.md 7.3-1 option b):
// if every S tile were computed once, the Z_b side would have to be reduced ACROSS CTAs: each (256-row query tile,
// 128-column key tile) contributes a [256 channels x 128 positions] fp32 partial to Z_b, added into global memory.
// This measures what the memory system sustains for exactly that: every CTA repeatedly issues
//   cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.f32   (TMA reduce-add of a [256 x 128] fp32 box, 128 KB)
// into rotating tiles of an [N * 256][Lp] fp32 array (N = 32, Lp = 3840: 126 MB, the Z_b of the headline batch).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/reduce_bench tools/reduce_bench.cu && tools/reduce_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWS>
__global__ void __launch_bounds__(128, 1) reduce_kernel(const __grid_constant__ CUtensorMap tmap, int tiles_x, int tiles_y, int iters,
                                                        int in_flight) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* tile = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < ROWS * 128; i += blockDim.x) tile[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const int num_tiles = tiles_x * tiles_y;
    for (int it = 0; it < iters; ++it) {
      // neighbouring CTAs hit DIFFERENT accumulator tiles (as the CTAs of one wave would: same key tile j of different
      // samples / query tiles -> different samples' Z_b), and a CTA walks along the key tiles of its sample
      const int t = (int)((blockIdx.x * 131u + (unsigned)it) % (unsigned)num_tiles);
      const int x = (t % tiles_x) * 128, y = (t / tiles_x) * ROWS;
      asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                   ::"l"(&tmap), "r"(x), "r"(y), "r"((uint32_t)__cvta_generic_to_shared(tile)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (in_flight == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else if (in_flight == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  const int N = 32, C = 256, Lp = 3840;
  float* buf;
  const size_t bytes = (size_t)N * C * Lp * 4;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int rows : {256, 128}) {
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)Lp, (cuuint64_t)N * C};
    const cuuint64_t strides[1] = {(cuuint64_t)Lp * 4};
    const cuuint32_t box[2] = {128, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int smem = rows * 128 * 4;
    auto kern = rows == 256 ? reduce_kernel<256> : reduce_kernel<128>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int in_flight : {1, 2, 4}) {
      const int iters = 200;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      kern<<<sms, 128, smem>>>(tm, Lp / 128, N * C / rows, 20, in_flight);
      cudaEventRecord(e0);
      kern<<<sms, 128, smem>>>(tm, Lp / 128, N * C / rows, iters, in_flight);
      cudaEventRecord(e1);
      cudaError_t ce = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double tb = (double)sms * iters * smem / (ms * 1e-3) / 1e12;
      printf("reduce-add f32 box [%d ch x 128 pos] (%d KB), %d CTAs, %d in flight per CTA: %.3f ms, %.2f TB/s of partials (%s)\n", rows,
             smem / 1024, sms, in_flight, ms, tb, cudaGetErrorString(ce));
    }
  }
  // what the single-S attend kernel would need at the headline shape: per (256-row query tile, 128-column key tile) ONE
  // 128 KB partial per CTA PAIR, every 3 x 1024 cycles of MMA (S, P V_b, A P instead of today's 2 x 2048 for two passes)
  printf("needed by a single-S attend kernel at its MMA pace: 74 CTA pairs x 128 KB / (3072 cycles / 1.9 GHz) = %.1f TB/s\n",
         74.0 * 131072 / (3072 / 1.9e9) / 1e12);
  return 0;
}
