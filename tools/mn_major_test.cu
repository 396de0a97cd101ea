// Descriptor experiment: tcgen05.mma with an MN-major B operand (N contiguous in memory), SWIZZLE_128B.
//   D[m][n] = sum_k A[m][k] * B[k][n]      A: [128][K] row-major (K-major operand), B: [K][128] row-major (MN-major operand)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cosnet_b200/csrc tools/mn_major_test.cu -o tools/mn_major_test
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace coattn;

constexpr int M = 128, N = 128, K = 128;

__device__ __forceinline__ void warp_mbar_wait(uint64_t* bar, uint32_t parity, int lane, int tag) {
  if (lane == 0) mbar_wait(bar, parity, tag);
  __syncwarp();
}

__device__ __forceinline__ uint64_t make_sdesc_mn_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// variant: 0 = B MN-major (the experiment), 1 = A MN-major as well (A given as [K][M])
template <int VARIANT>
__global__ void __launch_bounds__(192, 1) kern(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                                               float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // K-major: 2 k-blocks x [128 rows x 128 B] = 32 KB ; MN-major: 2 m-chunks x [K=128 rows x 128 B]
  uint8_t* sB = smem + 32768;         // MN-major: 2 n-chunks x [K = 128 rows x 128 B] = 32 KB
  __shared__ uint64_t full, done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&full, 1); mbar_init(&done, 1); fence_mbar_init(); }
  if (warp == 5) { tmem_alloc(&slot, 128); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 4 && lane == 0) {
    mbar_arrive_expect_tx(&full, 65536);
    if (VARIANT == 0) {
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(sA + kb * 16384, &tm_a, &full, kb * 64, 0);       // box {64 k, 128 m}
    } else {
      for (int mc = 0; mc < 2; ++mc) tma_load_2d(sA + mc * 16384, &tm_a, &full, mc * 64, 0);       // box {64 m, 128 k}
    }
    for (int nc = 0; nc < 2; ++nc) tma_load_2d(sB + nc * 16384, &tm_b, &full, nc * 64, 0);         // box {64 n, 128 k}
  } else if (warp == 5) {
    warp_mbar_wait(&full, 0, lane, 1);
    tc_fence_after();
    if (elect_one()) {
      // idesc: bf16 x bf16 -> f32, M128 N128; b_major (bit 16) = 1 (MN); a_major (bit 15) = VARIANT
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)VARIANT << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      for (int ks = 0; ks < K / 16; ++ks) {
        uint64_t ad;
        if (VARIANT == 0) ad = make_sdesc_k_sw128(smem_u32(sA + (ks >> 2) * 16384 + (ks & 3) * 32));
        else ad = make_sdesc_mn_sw128(smem_u32(sA + ks * 2048), 16384, 1024);
        const uint64_t bd = make_sdesc_mn_sw128(smem_u32(sB + ks * 2048), 16384, 1024);   // 16 K-rows = 2048 B per step
        umma_ss(tmem, ad, bd, idesc, ks > 0);
      }
      umma_commit(&done);
    }
    __syncwarp();
  } else if (warp < 4) {
    warp_mbar_wait(&done, 0, lane, 2);
    tc_fence_after();
    const int m = warp * 32 + lane;
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
      tmem_ld_wait();
      for (int k = 0; k < 32; ++k) out[m * N + ch * 32 + k] = __uint_as_float(v[k]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  std::vector<__nv_bfloat16> hA(M * K), hAt(K * M), hB(K * N);
  std::vector<float> fA(M * K), fB(K * N);
  srand(1);
  for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) hAt[k * M + m] = hA[m * K + k];
  for (int i = 0; i < K * N; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  std::vector<float> ref(M * N, 0.f);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[k * N + n]; ref[m * N + n] = (float)s; }
  __nv_bfloat16 *dA, *dAt, *dB; float* dO;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dAt, M * K * 2); cudaMalloc(&dB, K * N * 2); cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dAt, hAt.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), K * N * 2, cudaMemcpyHostToDevice);
  auto mk = [&](CUtensorMap* t, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows}; cuuint64_t str[1] = {cols * 2}; cuuint32_t box[2] = {64, box_rows}; cuuint32_t es[2] = {1, 1};
    return enc(t, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  for (int variant = 0; variant < 2; ++variant) {
    CUtensorMap ta, tb;
    if (variant == 0) mk(&ta, dA, M, K, 128); else mk(&ta, dAt, K, M, 128);
    mk(&tb, dB, K, N, 128);
    cudaMemset(dO, 0, M * N * 4);
    if (variant == 0) { cudaFuncSetAttribute(kern<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000); kern<0><<<1, 192, 70000>>>(ta, tb, dO); }
    else { cudaFuncSetAttribute(kern<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000); kern<1><<<1, 192, 70000>>>(ta, tb, dO); }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(M * N);
    cudaMemcpy(out.data(), dO, M * N * 4, cudaMemcpyDeviceToHost);
    double num = 0, den = 0; 
    for (int i = 0; i < M * N; ++i) { num += (out[i] - ref[i]) * (double)(out[i] - ref[i]); den += (double)ref[i] * ref[i]; }
    printf("variant %d (%s): %s  rel-L2 = %.3e   out[0..3] = %f %f %f %f  ref = %f %f %f %f\n", variant,
           variant == 0 ? "A K-major, B MN-major" : "A MN-major, B MN-major", cudaGetErrorString(e), sqrt(num / den), out[0], out[1], out[2],
           out[3], ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
