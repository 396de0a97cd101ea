// Micro-benchmark behind the stand-alone gate / concat epilogue (gate_kernel, north_star item 3): what does a PERSISTENT TMA COPY
// PIPELINE sustain on [rows][L = 3600] fp32 planes (row stride 14 400 B, the layout of Z, V and cat), as a function of the box
// shape?  A 32 KB box can be [256 rows x 32 positions] (every channel of 32 positions: what a one-CTA gate tile needs, 128 B
// per row), [128 x 64], [64 x 128] or [32 x 256] (1 KB per row).  One elected thread per CTA: TMA load -> shared memory ->
// TMA store, SLOTS-deep ring, loads SLOTS - 1 tiles ahead; tiles are dealt round-robin along the position axis first, like the
// items of the gate kernel.  Bytes counted: read + written.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_copy_bench tools/tma_copy_bench.cu && tools/tma_copy_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kTile = 32768;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32, 1) copy_kernel(const __grid_constant__ CUtensorMap src, const __grid_constant__ CUtensorMap dst,
                                                     int box_w, int box_h, int tiles_x, int num_tiles, int slots) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)slots * kTile);
  if (threadIdx.x != 0) return;
  for (int s = 0; s < slots; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  auto load = [&](int k, int t) {
    const int s = k % slots;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(kTile) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(smem + (size_t)s * kTile)), "l"(&src), "r"(s32(&full[s])), "r"((t % tiles_x) * box_w), "r"((t / tiles_x) * box_h)
                 : "memory");
  };
  int t_load = blockIdx.x, k_load = 0;
  for (; k_load < slots - 1 && t_load < num_tiles; ++k_load, t_load += gridDim.x) load(k_load, t_load);
  int k = 0;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++k) {
    if (t_load < num_tiles) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the slot of tile k - 1 has been read by its store
      load(k_load, t_load);
      ++k_load; t_load += gridDim.x;
    }
    const int s = k % slots;
    const uint32_t parity = (k / slots) & 1;
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                   : "=r"(ok) : "r"(s32(&full[s])), "r"(parity) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(&dst), "r"(s32(smem + (size_t)s * kTile)), "r"((t % tiles_x) * box_w), "r"((t / tiles_x) * box_h) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int rows = 2 * 32 * 256 * 2, L = 3600;      // 472 MB in, 472 MB out: the byte count of the gate epilogue at batch 32
  const size_t bytes = (size_t)rows * L * 4;
  float *a, *b;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaMemset(a, 1, bytes); cudaMemset(b, 0, bytes);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  {      // reference: cudaMemcpyAsync device to device of the same bytes
    for (int i = 0; i < 3; ++i) cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D of %.0f MB: %.3f ms, %.2f TB/s (read + written)\n", bytes / 1e6, ms / 10, 2.0 * bytes / (ms / 10) / 1e9);
  }
  for (int slots : {3, 6}) {
    for (int box_w : {32, 64, 128, 256}) {
      const int box_h = kTile / 4 / box_w;
      CUtensorMap ts, td;
      const cuuint64_t dims[2] = {(cuuint64_t)L, (cuuint64_t)rows};
      const cuuint64_t strides[1] = {(cuuint64_t)L * 4};
      const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
      const cuuint32_t estr[2] = {1, 1};
      for (auto pr : {std::pair<CUtensorMap*, float*>{&ts, a}, {&td, b}}) {
        const CUresult r = enc(pr.first, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, pr.second, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      }
      const int tiles_x = (L + box_w - 1) / box_w, num_tiles = tiles_x * (rows / box_h);
      const int smem = slots * kTile + 64;
      cudaFuncSetAttribute(copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      for (int i = 0; i < 2; ++i) copy_kernel<<<sms, 32, smem>>>(ts, td, box_w, box_h, tiles_x, num_tiles, slots);
      cudaEventRecord(e0);
      for (int i = 0; i < 5; ++i) copy_kernel<<<sms, 32, smem>>>(ts, td, box_w, box_h, tiles_x, num_tiles, slots);
      cudaEventRecord(e1);
      const cudaError_t err = cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("TMA copy, box [%3d rows x %3d positions] (%4d B per row), %d slots, %d CTAs: %.3f ms, %.2f TB/s (%s)\n", box_h, box_w, box_w * 4,
             slots, sms, ms / 5, 2.0 * bytes / (ms / 5) / 1e9, cudaGetErrorString(err));
    }
  }
  return 0;
}
