// Backward of the co-attention block (what autograd does through rgbd_segmentation_RAA.py:158-187, train.py:599).
//
// The softmax matrices are NOT kept from the forward pass (the reference keeps S_row and S_column, 104 MB per
// sample and modality at L = 3600): S is recomputed from the 16-bit operands and the saved log-sum-exp vectors.
//
//   bwd_stats / bwd_planes   d_cat_a/b, Z, mask, g  ->  delta_a, delta_b, d_gate, max |dZ|; then the scaled 16-bit planes
//                 s dZ_a, s dZ_b [C][Lp] (fp16 with an fp16 forward, bf16 with a bf16 forward) and the passthrough gradients
//   bwd_flash     (bwd_flash_kernel.cuh) everything of size L x L, flash style: S, dP_a, dP_b are recomputed tile by tile
//                 in TMEM and consumed there -- dQ = dS B^T, dA += P_b dZ_b^T (and dB with counterpart gradients)
//   gemm_nt  x2   dA += dQ W;  dW += dQ^T A^T
//
// with P_a[i,j] = exp(S[i,j] - lse_a[i]) (softmax over j, :165) and P_b[i,j] = exp(S[i,j] - lse_b[j]) (:164).
// Gradient semantics follow the reference: the B-side gate mask is a constant (:178-182) and, with
// no_grad_for_counterpart (:144-148), V_b receives no gradient.
#pragma once
#include <type_traits>

#include "coattn_kernels.cuh"

namespace coattn {

// ==============================================================================================
// gemm_nt: D[b][m][n] = sum_k A[b][m][k] * B[b][n][k]     (16-bit K-major operands, fp32 accumulate in TMEM)
// one 128x128 output tile per CTA, K streamed in 64-element blocks through a 3-stage TMA ring; two CTAs per SM so
// one tile's epilogue overlaps the other's main loop.
// ==============================================================================================
constexpr int kGemmStages = 3;
constexpr int kGemmStageBytes = 2 * 128 * 128;   // A 16 KB + B 16 KB
constexpr int kGemmSmemBytes = kGemmStages * kGemmStageBytes + 1024 + 128;

enum GemmMode : int {
  kGemmStoreF32 = 0,        // out0[(b*rows0 + m)*ld0 + n] = d                               (m < m_valid)
  kGemmStore16Both = 1,     // out0 (bf16 [m][n]) and out1 (bf16 transposed [n][m])
  kGemmAddF32T = 2,         // out0[(b*rows0 + n)*ld0 + m] += d   transposed, m < m_valid       (dA contributions)
  kGemmAtomicF32 = 3        // red.add(out0[m*ld0 + n], d)        no batch offset, K may be split    (dW, reduced over b)
};

struct GemmParams {
  void* out0;
  void* out1;
  int64_t ld0, ld1;        // leading dimensions in elements
  int64_t rows0, rows1;    // rows per batch entry of out0 / out1
  int a_rows_per_batch;    // row offset between batch entries in the A / B tensor maps
  int b_rows_per_batch;
  int num_kb;              // K / 64 handled by one CTA
  int k_split;             // CTAs per batch entry along K (atomic mode only); blockIdx.z = b * k_split + slice
  int m_valid;             // rows m >= m_valid are not stored
  uint32_t idesc;          // instruction descriptor (M128 N128, operand formats)
};

template <int MODE>
__global__ void __launch_bounds__(kNumThreads, 2)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmap_a,  // [batch*a_rows][K], box {64, 128}
               const __grid_constant__ CUtensorMap tmap_b,  // [batch*b_rows][K], box {64, 128}
               GemmParams p) {
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGemmStages * kGemmStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kGemmStages;
  uint64_t* d_full = empty + kGemmStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int n0 = blockIdx.y * 128;
  const int b = blockIdx.z / p.k_split;
  const int kb0 = (blockIdx.z % p.k_split) * p.num_kb;   // first 64-element k-block of this CTA's K slice

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kGemmStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(d_full, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      const int arow = b * p.a_rows_per_batch + m0;
      const int brow = b * p.b_rows_per_batch + n0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % kGemmStages;
        const uint32_t ph = (kb / kGemmStages) & 1;
        mbar_wait(empty + s, ph ^ 1, 30);
        mbar_arrive_expect_tx(full + s, kGemmStageBytes);
        tma_load_2d(smem + s * kGemmStageBytes, &tmap_a, full + s, (kb0 + kb) * 64, arow);
        tma_load_2d(smem + s * kGemmStageBytes + 16384, &tmap_b, full + s, (kb0 + kb) * 64, brow);
      }
    }
  } else if (warp == kMmaWarp) {
    const uint32_t base = smem_u32(smem);
    for (int kb = 0; kb < p.num_kb; ++kb) {
      const int s = kb % kGemmStages;
      const uint32_t ph = (kb / kGemmStages) & 1;
      warp_mbar_wait(full + s, ph, lane, 31);
      tc_fence_after();
      const uint64_t ad0 = make_sdesc_k_sw128(base + s * kGemmStageBytes);
      const uint64_t bd0 = make_sdesc_k_sw128(base + s * kGemmStageBytes + 16384);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem, ad0 + 2 * k, bd0 + 2 * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(empty + s);
        if (kb == p.num_kb - 1) umma_commit(d_full);
      }
      __syncwarp();
    }
  } else {
    warp_mbar_wait(d_full, 0, lane, 32);
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const bool valid = m < p.m_valid;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v[32];
      tmem_ld32(taddr + ch * 32, v);
      tmem_ld_wait();
      const int n = n0 + ch * 32;
      if constexpr (MODE == kGemmStoreF32) {
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out0) + ((int64_t)b * p.rows0 + m) * p.ld0 + n);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3]));
        }
      } else if constexpr (MODE == kGemmStore16Both) {
        unsigned short* o0 = static_cast<unsigned short*>(p.out0) + ((int64_t)b * p.rows0 + m) * p.ld0 + n;
        unsigned short* o1 = static_cast<unsigned short*>(p.out1) + ((int64_t)b * p.rows1 + n) * p.ld1 + m;
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) pk[q] = pack_bf16x2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
        uint4* d4 = reinterpret_cast<uint4*>(o0);
#pragma unroll
        for (int q = 0; q < 4; ++q) d4[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
#pragma unroll
        for (int k = 0; k < 32; ++k)
          o1[(int64_t)k * p.ld1] = (unsigned short)((k & 1) ? (pk[k >> 1] >> 16) : (pk[k >> 1] & 0xFFFFu));
      } else if constexpr (MODE == kGemmAddF32T) {
        if (valid) {
          // all 32 loads first: with a run-time stride the compiler cannot prove the 32 addresses distinct and would
          // otherwise serialise load -> add -> store (32 dependent memory round trips per chunk)
          float* o = static_cast<float*>(p.out0) + ((int64_t)b * p.rows0 + n) * p.ld0 + m;
          float old[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) old[k] = o[(int64_t)k * p.ld0];
#pragma unroll
          for (int k = 0; k < 32; ++k) o[(int64_t)k * p.ld0] = old[k] + __uint_as_float(v[k]);
        }
      } else {
        // vectorised reductions (4 floats per operation): the row of a thread is contiguous in the output
        float* o = static_cast<float*>(p.out0) + (int64_t)m * p.ld0 + n;
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + k), "f"(__uint_as_float(v[k])),
                       "f"(__uint_as_float(v[k + 1])), "f"(__uint_as_float(v[k + 2])), "f"(__uint_as_float(v[k + 3]))
                       : "memory");
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// ==============================================================================================
// bwd_prep: everything that is per position (no L x L work).  One block = 32 positions x 256 channels
// (warp = 32 channels, lane = position, so every global access is a coalesced 128-byte row segment).
//   d_ta   = (sum_c dZag Z_a) m_a (1 - m_a)          gate logit gradient, A side only (:177-182)
//   dZ_a   = dZag m_a + g d_ta                        dZ_b = dZbg m_b
//   delta  = sum_c dZ_x Z_x
//   d_gate_w += sum_i d_ta[i] Z_a[:, i]               d_gate_b += sum_i d_ta[i]
//   dA      = d_cat_a[:, C:2C]                        (passthrough half of the concat, :186)
// Two passes, because the 16-bit gradient planes are SCALED: with fp16 forward operands the gradient operands of the
// flash sweeps are fp16 as well (one operand format per tcgen05 product; 11 significant bits instead of bf16's 8, which
// is what the gradient parity rides on), and fp16 needs the values in range -- every dZ is multiplied by ONE power of two
// s chosen from max |dZ| over the whole call.
//   PASS 0  reads d_cat (gated halves), Z, mask: d_ta -> buffer, delta, d_gate, max |dZ| (one atomicMax per block)
//   PASS 1  reads d_cat, mask, d_ta, the maximum: writes s dZ_a, s dZ_b as 16-bit planes [N][C][Lp] (the orientation every
//           consumer takes) and the passthrough gradients
// With bf16 forward operands the planes are bf16 and s = 1.
// ==============================================================================================
struct BwdPrepParams {
  const float* d_cat_a;   // [N][2C][L]
  const float* d_cat_b;   // [N][2C][L] or null (depth: the B branch is gradient dead)
  const float* z;         // [2][N][C][L]
  const float* mask;      // [2][N][L]
  const float* gate_w;    // [C]
  unsigned short* dza16;  // [N][C][Lp] 16-bit, scaled
  unsigned short* dzb16;  // [N][C][Lp] 16-bit, scaled
  float* d_vb;            // [N][C][L] or null: initialised with the passthrough gradient d_cat_b[:, C:2C]
  float* delta;           // [2][N][L]  (unscaled)
  float* d_ta;            // [N][L] scratch between the passes
  unsigned* absmax;       // [1] bits of max |dZ| (caller zeroes)
  float* d_gate_w;        // [C]   (accumulated with atomics; caller zeroes)
  float* d_gate_b;        // [1] or null
  float* d_va;            // [N][C][L]  initialised with the passthrough gradient
  int N, L, Lp;
  int cat_ch;             // channels per sample of d_cat_a / d_cat_b: 2C (concat), or C (COATTN_FLAG_GATED_ONLY: the gated half
                          // only -- no passthrough gradient, d_va / d_vb start from zero)
};

constexpr int kBwdPrepThreads = 256;
constexpr int kBwdPrepPos = 32;

// The power of two that brings max |dZ| into [2^-5, 2^-4): products with O(1) features summed over 256 channels and
// multiplied by softmax weights <= 1 then stay far inside the fp16 range, and values down to 2^-20 of the maximum are
// still representable.  bits = 0 (all gradients zero) -> 1.
__device__ __forceinline__ float grad_scale_from_absmax(unsigned bits) {
  const float m = __uint_as_float(bits);
  if (!(m > 0.f) || !(m < 3.0e38f)) return 1.0f;
  int e;
  (void)frexpf(m, &e);          // m = f 2^e, f in [0.5, 1)
  return ldexpf(1.0f, -e - 4);
}

// sum over the 32 lanes of v[k] for every k: after the five exchange steps lane l holds the total of element
// k = bitreverse5(l) ... expressed here simply as "the element this lane ends up with"; 31 shuffles instead of 160
__device__ __forceinline__ float warp_multi_reduce32(float (&v)[32], int lane, int& owner) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const bool up = lane & 16;
    const float send = up ? v[k] : v[k + 16];
    const float keep = up ? v[k + 16] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool up = lane & 8;
    const float send = up ? v[k] : v[k + 8];
    const float keep = up ? v[k + 8] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool up = lane & 4;
    const float send = up ? v[k] : v[k + 4];
    const float keep = up ? v[k + 4] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool up = lane & 2;
    const float send = up ? v[k] : v[k + 2];
    const float keep = up ? v[k + 2] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  // bit b of the lane selected the upper half at the step of width b: element index = the lane bits read as written
  owner = (lane & 16) | (lane & 8) | (lane & 4) | (lane & 2) | (lane & 1);
  return v[0];
}

// PASS 0: statistics
__global__ void __launch_bounds__(kBwdPrepThreads) bwd_stats_kernel(BwdPrepParams p) {
  pdl_wait();
  __shared__ float red[8][kBwdPrepPos];
  __shared__ float redm[8];
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int wrp = threadIdx.x >> 5;      // channels [32 wrp, 32 wrp + 32)
  const int l = blockIdx.x * kBwdPrepPos + lane;
  const bool valid = l < p.L;
  const int c0 = wrp * 32;
  const float* dca = p.d_cat_a + ((size_t)n * p.cat_ch + c0) * p.L + l;
  const float* za = p.z + ((size_t)n * kC + c0) * p.L + l;
  const bool has_b = p.d_cat_b != nullptr;
  const float ma = valid ? __ldg(p.mask + (size_t)n * p.L + l) : 0.f;
  const float mb = (valid && has_b) ? __ldg(p.mask + (size_t)(p.N + n) * p.L + l) : 0.f;

  // ---- A side: this thread's 32 channels of dZag and Z_a stay in registers
  float g[32], zz[32];
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    g[k] = valid ? __ldcs(dca + (size_t)k * p.L) : 0.f;
    zz[k] = valid ? __ldcs(za + (size_t)k * p.L) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) acc = fmaf(g[k], zz[k], acc);
  red[wrp][lane] = acc;
  __syncthreads();
  float dta = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) dta += red[w][lane];
  dta *= ma * (1.f - ma);
  __syncthreads();
  if (wrp == 0 && valid) p.d_ta[(size_t)n * p.L + l] = dta;
  float dl = 0.f, amax = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float dz = fmaf(__ldg(p.gate_w + c0 + k), dta, g[k] * ma);
    dl = fmaf(dz, zz[k], dl);
    amax = fmaxf(amax, fabsf(dz));
    zz[k] *= dta;                                // d_gate_w contribution of this position
  }
  red[wrp][lane] = dl;
  int owner;
  const float gsum = warp_multi_reduce32(zz, lane, owner);
  atomicAdd(p.d_gate_w + c0 + owner, gsum);
  __syncthreads();
  if (wrp == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    if (valid) p.delta[(size_t)n * p.L + l] = t;
    if (p.d_gate_b != nullptr) {
      float s = dta;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) atomicAdd(p.d_gate_b, s);
    }
  }
  __syncthreads();
  // ---- B side: dZ_b = dZbg * m_b (the mask is a constant), delta_b
  if (has_b) {
    dl = 0.f;
    const float* dcb = p.d_cat_b + ((size_t)n * p.cat_ch + c0) * p.L + l;
    const float* zb = p.z + ((size_t)(p.N + n) * kC + c0) * p.L + l;
#pragma unroll
    for (int k = 0; k < 32; ++k) { g[k] = valid ? __ldcs(dcb + (size_t)k * p.L) : 0.f; zz[k] = valid ? __ldcs(zb + (size_t)k * p.L) : 0.f; }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float dz = g[k] * mb;
      dl = fmaf(dz, zz[k], dl);
      amax = fmaxf(amax, fabsf(dz));
    }
    red[wrp][lane] = dl;
    __syncthreads();
    if (wrp == 0 && valid) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][lane];
      p.delta[(size_t)(p.N + n) * p.L + l] = t;
    }
  }
  // ---- max |dZ| of the block -> one atomic (the bits of a non-negative float order like the float)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
  if (lane == 0) redm[wrp] = amax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = redm[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, redm[w]);
    if (m > 0.f) atomicMax(p.absmax, __float_as_uint(m));
  }
}

// PASS 1: scaled 16-bit planes + passthrough gradients.  BF16: bf16 planes, no scaling.
template <bool BF16>
__global__ void __launch_bounds__(kBwdPrepThreads) bwd_planes_kernel(BwdPrepParams p) {
  pdl_wait();
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int wrp = threadIdx.x >> 5;
  const int l = blockIdx.x * kBwdPrepPos + lane;
  const bool valid = l < p.L;
  const int c0 = wrp * 32;
  const bool has_b = p.d_cat_b != nullptr;
  const float s = BF16 ? 1.0f : grad_scale_from_absmax(__ldg(p.absmax));
  const float ma = valid ? __ldg(p.mask + (size_t)n * p.L + l) : 0.f;
  const float mb = (valid && has_b) ? __ldg(p.mask + (size_t)(p.N + n) * p.L + l) : 0.f;
  const float dta = valid ? __ldg(p.d_ta + (size_t)n * p.L + l) : 0.f;
  float g[32];
  {
    const float* dca = p.d_cat_a + ((size_t)n * p.cat_ch + c0) * p.L + l;
    unsigned short* dza = p.dza16 + ((size_t)n * kC + c0) * p.Lp + l;
    const bool pass = p.cat_ch == 2 * kC;
#pragma unroll
    for (int k = 0; k < 32; ++k) g[k] = valid ? __ldcs(dca + (size_t)k * p.L) : 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k)       // positions >= L get 0 (g = 0, dta = 0)
      dza[(size_t)k * p.Lp] = cvt16<BF16>(fmaf(__ldg(p.gate_w + c0 + k), dta, g[k] * ma) * s);
    // passthrough half of the concat -> gradient of V_a
    if (valid) {
      const float* src = dca + (size_t)kC * p.L;
      float* dst = p.d_va + ((size_t)n * kC + c0) * p.L + l;
#pragma unroll
      for (int k = 0; k < 32; ++k) g[k] = pass ? __ldcs(src + (size_t)k * p.L) : 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) __stcs(dst + (size_t)k * p.L, g[k]);
    }
  }
  if (has_b) {
    const float* dcb = p.d_cat_b + ((size_t)n * p.cat_ch + c0) * p.L + l;
    unsigned short* dzb = p.dzb16 + ((size_t)n * kC + c0) * p.Lp + l;
#pragma unroll
    for (int k = 0; k < 32; ++k) g[k] = valid ? __ldcs(dcb + (size_t)k * p.L) : 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) dzb[(size_t)k * p.Lp] = cvt16<BF16>(g[k] * mb * s);
    if (p.d_vb != nullptr && valid) {
      const float* src = dcb + (size_t)kC * p.L;
      float* dst = p.d_vb + ((size_t)n * kC + c0) * p.L + l;
#pragma unroll
      for (int k = 0; k < 32; ++k) g[k] = (p.cat_ch == 2 * kC) ? __ldcs(src + (size_t)k * p.L) : 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) __stcs(dst + (size_t)k * p.L, g[k]);
    }
  } else if (p.d_vb != nullptr && valid) {
    // depth modality: the B branch is gradient dead, nothing reads dZ_b or delta_b
    float* dst = p.d_vb + ((size_t)n * kC + c0) * p.L + l;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) dst[(size_t)k * p.L] = 0.f;
  }
}

// Start of a backward call, ONE launch instead of four memsets and two small kernels (each ~2.5 us of launch gap in a call
// whose small kernels add up to a quarter of its time): grid = C blocks of C threads.
//   Wt [C_in][C_out] bf16 = W^T   (B operand of dA += dQ W: dA[i][c] = sum_d dQt[i][d] Wt[c][d])
//   W16 [C_out][C_in] = W in the forward's operand format (w16 != null: the feature cast, which would have done it, is skipped)
//   d_w, d_gate_w, d_gate_b, absmax = 0   (accumulated with atomics by the kernels that follow)
template <bool BF16>
__global__ void bwd_init_kernel(const float* __restrict__ w, unsigned short* __restrict__ wt, unsigned short* __restrict__ w16,
                                float* __restrict__ d_w, float* __restrict__ d_gate_w, float* __restrict__ d_gate_b,
                                unsigned* __restrict__ absmax) {
  pdl_wait();
  const int c = blockIdx.x;
  const int d = threadIdx.x;
  wt[c * kC + d] = cvt16<true>(w[d * kC + c]);
  if (w16 != nullptr) w16[c * kC + d] = cvt16<BF16>(w[c * kC + d]);
  d_w[c * kC + d] = 0.f;
  if (c == 0) {
    d_gate_w[d] = 0.f;
    if (d == 0) {
      if (d_gate_b != nullptr) *d_gate_b = 0.f;
      *absmax = 0u;
    }
  }
}

}  // namespace coattn
