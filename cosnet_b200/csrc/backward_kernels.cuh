// Backward of the co-attention block (what autograd does through rgbd_segmentation_RAA.py:158-187, train.py:599).
//
// The softmax matrices are NOT kept from the forward pass (the reference keeps S_row and S_column, 104 MB per
// sample and modality at L = 3600): S is recomputed from the 16-bit operands and the saved log-sum-exp vectors.
//
//   bwd_prep      d_cat_a/b, Z, mask, g  ->  dZ_a, dZ_b (bf16 planes [C][Lp]), delta_a, delta_b, d_gate, dA init
//   bwd_tile      per 128x128 tile: S = Q^T B, dP_a = dZ_a^T B, dP_b = A^T dZ_b in TMEM (operands read MN-major from
//                 the channel-major planes), combined on the fly into
//                 dS = P_a (dP_a - delta_a) + P_b (dP_b - delta_b) and P_b            (bf16, [L, L], transient)
//   gemm_nt  x4   dQ = dS B^T;  dA += P_b dZ_b^T;  dA += dQ W;  dW += dQ^T A^T
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
// dZ_a / dZ_b leave as bf16 planes [N][C][Lp] -- the orientation every consumer takes (MN-major operands of the
// tile kernel, K-major operands of the position-contracting GEMMs).
// ==============================================================================================
struct BwdPrepParams {
  const float* d_cat_a;   // [N][2C][L]
  const float* d_cat_b;   // [N][2C][L] or null (depth: the B branch is gradient dead)
  const float* z;         // [2][N][C][L]
  const float* mask;      // [2][N][L]
  const float* gate_w;    // [C]
  unsigned short* dza16;  // [N][C][Lp] bf16
  unsigned short* dzb16;  // [N][C][Lp] bf16
  float* d_vb;            // [N][C][L] or null: initialised with the passthrough gradient d_cat_b[:, C:2C]
  float* delta;           // [2][N][L]
  float* d_gate_w;        // [C]   (accumulated with atomics; caller zeroes)
  float* d_gate_b;        // [1] or null
  float* d_va;            // [N][C][L]  initialised with the passthrough gradient
  int N, L, Lp;
};

constexpr int kBwdPrepThreads = 256;
constexpr int kBwdPrepPos = 32;

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

__global__ void __launch_bounds__(kBwdPrepThreads) bwd_prep_kernel(BwdPrepParams p) {
  __shared__ float red[8][kBwdPrepPos];
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int wrp = threadIdx.x >> 5;      // channels [32 wrp, 32 wrp + 32)
  const int l = blockIdx.x * kBwdPrepPos + lane;
  const bool valid = l < p.L;
  const int c0 = wrp * 32;
  const float* dca = p.d_cat_a + ((size_t)n * 2 * kC + c0) * p.L + l;
  const float* za = p.z + ((size_t)n * kC + c0) * p.L + l;
  const bool has_b = p.d_cat_b != nullptr;
  const float ma = valid ? __ldg(p.mask + (size_t)n * p.L + l) : 0.f;
  const float mb = (valid && has_b) ? __ldg(p.mask + (size_t)(p.N + n) * p.L + l) : 0.f;

  // ---- A side, pass 1: this thread's 32 channels of dZag and Z_a stay in registers
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
  // ---- pass 2: dZ_a, delta_a partial, d_gate_w partials
  float dl = 0.f;
  unsigned short* dza = p.dza16 + ((size_t)n * kC + c0) * p.Lp + l;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float dz = fmaf(__ldg(p.gate_w + c0 + k), dta, g[k] * ma);
    dl = fmaf(dz, zz[k], dl);
    dza[(size_t)k * p.Lp] = cvt16<true>(dz);     // positions >= L get 0 (g = 0, dta = 0)
    zz[k] *= dta;                                // d_gate_w contribution of this position
  }
  red[wrp][lane] = dl;
  int owner;
  const float gsum = warp_multi_reduce32(zz, lane, owner);
  atomicAdd(p.d_gate_w + c0 + owner, gsum);
  // passthrough half of the concat -> gradient of V_a
  {
    const float* src = p.d_cat_a + ((size_t)n * 2 * kC + kC + c0) * p.L + l;
    float* dst = p.d_va + ((size_t)n * kC + c0) * p.L + l;
    if (valid) {
      float t[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) t[k] = __ldcs(src + (size_t)k * p.L);
#pragma unroll
      for (int k = 0; k < 32; ++k) __stcs(dst + (size_t)k * p.L, t[k]);
    }
  }
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
  dl = 0.f;
  unsigned short* dzb = p.dzb16 + ((size_t)n * kC + c0) * p.Lp + l;
  if (has_b) {
    const float* dcb = p.d_cat_b + ((size_t)n * 2 * kC + c0) * p.L + l;
    const float* zb = p.z + ((size_t)(p.N + n) * kC + c0) * p.L + l;
#pragma unroll
    for (int k = 0; k < 32; ++k) { g[k] = valid ? __ldcs(dcb + (size_t)k * p.L) : 0.f; zz[k] = valid ? __ldcs(zb + (size_t)k * p.L) : 0.f; }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float dz = g[k] * mb;
      dl = fmaf(dz, zz[k], dl);
      dzb[(size_t)k * p.Lp] = cvt16<true>(dz);
    }
    if (p.d_vb != nullptr && valid) {
      const float* src = dcb + (size_t)kC * p.L;
      float* dst = p.d_vb + ((size_t)n * kC + c0) * p.L + l;
#pragma unroll
      for (int k = 0; k < 32; ++k) g[k] = __ldcs(src + (size_t)k * p.L);
#pragma unroll
      for (int k = 0; k < 32; ++k) __stcs(dst + (size_t)k * p.L, g[k]);
    }
  } else {
    // depth modality: the B branch is gradient dead, nothing reads dZ_b or delta_b
    if (p.d_vb != nullptr && valid) {
      float* dst = p.d_vb + ((size_t)n * kC + c0) * p.L + l;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) dst[(size_t)k * p.L] = 0.f;
    }
  }
  if (!has_b) return;
  red[wrp][lane] = dl;
  __syncthreads();
  if (wrp == 0 && valid) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    p.delta[(size_t)(p.N + n) * p.L + l] = t;
  }
}

// ==============================================================================================
// bwd_tile: the three [L, L] products of the backward and their elementwise combination in ONE persistent kernel.
// Work item = one 128 x 128 tile (i-tile, j-tile) of one sample; nothing of size L x L is ever written in fp32:
//   S    = Q[:, i]^T B[:, j]      (forward operand format, so that exp(S - lse) is exactly the forward's softmax)
//   dP_a = dZ_a[:, i]^T B[:, j]   (bf16)
//   dP_b = A[:, i]^T dZ_b[:, j]   (bf16; HAS_B only)
// All six operands are read as MN-major tiles straight from channel-major planes [N][C][Lp] (positions contiguous,
// channels = K), the orientation the forward pass and bwd_prep produce -- nothing is transposed for this kernel.
//   dS = P_a (dP_a - delta_a[i]) + P_b (dP_b - delta_b[j]),   P_a = exp(S - lse_a[i]),  P_b = exp(S - lse_b[j])
// written as bf16: dS, P_b (HAS_B) and P_a (counterpart gradients only).  TMEM: three 128-column accumulators.
// One CTA per SM walks the tiles (i fastest, so the CTAs running together share a handful of j-tiles in L2); the TMA
// ring runs continuously across tiles, so the 384 KB of operand tiles of the next item stream in while the eight
// epilogue warps (two per TMEM lane quadrant, 64 columns each) are still combining the current one.
// ==============================================================================================
constexpr int kTileThreads = 320;            // warps 0-7 epilogue, 8 TMA producer, 9 MMA issuer
constexpr int kTileProducerWarp = 8;
constexpr int kTileMmaWarp = 9;
constexpr int kTileRingBytes = 12 * 16384;   // 6 stages x (A block + B block) of one product and one 64-channel k-block
constexpr int kTileSmemBytes = kTileRingBytes + 8 * 2048 /*store staging*/ + 1024 /*align*/ + 2048 /*column vectors x2*/ + 128;

struct BwdTileParams {
  const float* lse;      // [2][N][L]
  const float* delta;    // [2][N][L]
  unsigned short* ds;    // [N][Lp][Lp] bf16
  unsigned short* pb;    // [N][Lp][Lp] bf16 (HAS_B)
  unsigned short* pa;    // [N][Lp][Lp] bf16 or null
  int N, L, Lp;
  int tiles_1d;          // Lp / 128
  int num_tiles;         // N * tiles_1d^2
  uint32_t idesc_fwd;    // M128 N128, forward operand format
  uint32_t idesc_bf16;   // M128 N128, bf16 x bf16
};

// clock64 accounting of one CTA (debug builds only): where the producer, the MMA issuer and epilogue warp 0 spend time
#ifdef COATTN_TRACE_BWD
#define BT_T0() long long bt_t0 = clock64()
#define BT_ACC(i) do { const long long bt_now = clock64(); bt_acc[i] += bt_now - bt_t0; bt_t0 = bt_now; } while (0)
#define BT_REPORT()                                                                                                   \
  do {                                                                                                                \
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kTileProducerWarp || warp == kTileMmaWarp))              \
      printf("bwd_tile<%d> warp %d: empty-wait %lld | d_empty-wait %lld | full-wait %lld | prologue %lld | d_full-wait %lld | " \
             "combine+store %lld | total %lld cycles\n", (int)HAS_B, warp, bt_acc[0], bt_acc[1], bt_acc[2], bt_acc[3],  \
             bt_acc[4], bt_acc[5], clock64() - bt_start);                                                              \
  } while (0)
#else
#define BT_T0() do {} while (0)
#define BT_ACC(i) do {} while (0)
#define BT_REPORT() do {} while (0)
#endif

template <bool HAS_B>
__global__ void __launch_bounds__(kTileThreads, 1)
bwd_tile_kernel(const __grid_constant__ CUtensorMap tm_qt, const __grid_constant__ CUtensorMap tm_bt,
                const __grid_constant__ CUtensorMap tm_dza, const __grid_constant__ CUtensorMap tm_btg,
                const __grid_constant__ CUtensorMap tm_atg, const __grid_constant__ CUtensorMap tm_dzb,
                BwdTileParams p) {
  extern __shared__ uint8_t smem_raw[];
  // operand blocks per stage: Q, B | dZ_a, B(bf16) | A(bf16), dZ_b   (64 channels x 128 positions = 16 KB each, stored as
  // two 64-position chunks of 64 channel rows x 128 B, 128-byte swizzle)
  // Stage = the operand pair (two 16 KB blocks) of ONE product for one k-block.  The loads and MMAs of a tile run product
  // by product (S: 16 MMAs, then dP_a, then dP_b): switching the accumulator every four MMAs stalls the tensor pipe
  // (measured on the attend kernel), three switches per tile do not.
  constexpr int kTileStages = 6;
  constexpr int kTileStageBytes = 2 * 16384;
  constexpr int kProducts = HAS_B ? 3 : 2;
  uint8_t* smem = align_1024(smem_raw);
  uint8_t* stage_out = smem + kTileRingBytes;                                      // 8 warps x 2 KB
  float* colv = reinterpret_cast<float*>(stage_out + 8 * 2048);                    // 2 x ([0,128) lse_b, [128,256) delta_b)
  uint64_t* bars = reinterpret_cast<uint64_t*>(colv + 512);
  uint64_t* full = bars;
  uint64_t* empty = bars + kTileStages;
  // accumulator sets in TMEM: S | dP_a | dP_b = 384 columns with the B branch (one set), 256 without (two sets, so
  // the epilogue of one tile overlaps the MMAs of the next)
  constexpr int kAccSets = HAS_B ? 1 : 2;
  constexpr uint32_t kAccStride = 256;
  uint64_t* d_full = empty + kTileStages;      // [2]
  uint64_t* d_empty = d_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef COATTN_TRACE_BWD
  long long bt_acc[6] = {0, 0, 0, 0, 0, 0};
  const long long bt_start = clock64();
#endif

  if (warp == kTileProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_qt); tma_prefetch_desc(&tm_bt); tma_prefetch_desc(&tm_dza);
    tma_prefetch_desc(&tm_btg); tma_prefetch_desc(&tm_atg); tma_prefetch_desc(&tm_dzb);
    for (int s = 0; s < kTileStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(d_full + a, 1); mbar_init(d_empty + a, 8); }
    fence_mbar_init();
  }
  if (warp == kTileMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr int kNumKb = kC / 64;
  const int per_sample = p.tiles_1d * p.tiles_1d;

  if (warp == kTileProducerWarp) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n = t / per_sample, r = t - n * per_sample;
        const int ipos = (r % p.tiles_1d) * 128, jpos = (r / p.tiles_1d) * 128;
        for (int pr = 0; pr < kProducts; ++pr) {
          // product 0: S = Q^T B (forward format) | 1: dP_a = dZ_a^T B (bf16) | 2: dP_b = A^T dZ_b (bf16)
          const CUtensorMap* ta = (pr == 0) ? &tm_qt : (pr == 1) ? &tm_dza : &tm_atg;    // i-side operand
          const CUtensorMap* tb = (pr == 0) ? &tm_bt : (pr == 1) ? &tm_btg : &tm_dzb;    // j-side operand
          for (int kb = 0; kb < kNumKb; ++kb, ++cnt) {
            const uint32_t s = cnt % kTileStages, ph = (cnt / kTileStages) & 1;
            BT_T0();
            mbar_wait(empty + s, ph ^ 1, 40);
            BT_ACC(0);
            mbar_arrive_expect_tx(full + s, kTileStageBytes);
            uint8_t* st = smem + s * kTileStageBytes;
            const int crow = n * kC + kb * 64;     // channel row of this k-block in a [N*C][Lp] plane
#pragma unroll
            for (int mc = 0; mc < 2; ++mc) {
              tma_load_2d(st + mc * 8192, ta, full + s, ipos + mc * 64, crow);
              tma_load_2d(st + 16384 + mc * 8192, tb, full + s, jpos + mc * 64, crow);
            }
          }
        }
      }
    }
  } else if (warp == kTileMmaWarp) {
    const uint32_t base = smem_u32(smem);
    uint32_t cnt = 0, it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const uint32_t acc = it % kAccSets, aph = (it / kAccSets) & 1;
      const uint32_t tacc = tmem + acc * kAccStride;
      BT_T0();
      warp_mbar_wait(d_empty + acc, aph ^ 1, lane, 43);   // the epilogue has read this set's previous accumulators
      BT_ACC(1);
      tc_fence_after();
      for (int pr = 0; pr < kProducts; ++pr) {
        const uint32_t idesc = (pr == 0) ? p.idesc_fwd : p.idesc_bf16;
        const uint32_t td = tacc + (uint32_t)pr * 128;
        for (int kb = 0; kb < kNumKb; ++kb, ++cnt) {
          const uint32_t s = cnt % kTileStages, ph = (cnt / kTileStages) & 1;
          BT_T0();
          warp_mbar_wait(full + s, ph, lane, 41);
          BT_ACC(2);
          tc_fence_after();
          const uint32_t sb = base + s * kTileStageBytes;
          // MN-major tiles: 64-position chunks 8192 B apart (LBO), 8-channel groups 1024 B apart (SBO); 16 channels = 2048 B
          const uint64_t da = make_sdesc_mn_sw128(sb, 8192, 1024), db = make_sdesc_mn_sw128(sb + 16384, 8192, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(td, da + 128 * k, db + 128 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(empty + s);
            if (pr == kProducts - 1 && kb == kNumKb - 1) umma_commit(d_full + acc);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int quad = warp & 3;            // TMEM lane quadrant
    const int half = warp >> 2;           // columns [64 half, 64 half + 64) of the tile
    const int et = threadIdx.x;           // 0..255
    uint8_t* stg = stage_out + warp * 2048;   // this warp's 32 rows x 64 B staging block
    // staging block addressing: 16-byte chunk c of row r sits at r*64 + ((c ^ ((r >> 1) & 3)) * 16) -- conflict-free both
    // for the row-per-lane writes and for the 8-rows-per-instruction reads that feed coalesced global stores
    const uint32_t wr_row = (uint32_t)lane * 64, wr_x = ((uint32_t)lane >> 1) & 3;
    const int rd_r = lane >> 2, rd_c = lane & 3;
    auto flush = [&](unsigned short* base, size_t tile_off, const uint32_t* pk) {
      // pk: this lane's row, 32 columns as 16 packed pairs
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(stg + wr_row + ((c ^ wr_x) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = rd_r + 8 * k;
        const uint4 v = *reinterpret_cast<const uint4*>(stg + r * 64 + ((rd_c ^ ((r >> 1) & 3)) << 4));
        *reinterpret_cast<uint4*>(base + tile_off + (size_t)r * p.Lp + rd_c * 8) = v;
      }
      __syncwarp();
    };
    // per-tile vectors are fetched one tile ahead so that their latency hides behind the previous tile's combine
    auto fetch = [&](int t, float& x, float& la, float& da) {
      const int n = t / per_sample, r = t - n * per_sample;
      const int i = (r % p.tiles_1d) * 128 + quad * 32 + lane;
      const int j = (r / p.tiles_1d) * 128 + (et & 127);
      x = (HAS_B && j < p.L) ? __ldg((et < 128 ? p.lse : p.delta) + (size_t)(p.N + n) * p.L + j) : 0.f;
      la = (i < p.L) ? __ldg(p.lse + (size_t)n * p.L + i) : 0.f;
      da = (i < p.L) ? __ldg(p.delta + (size_t)n * p.L + i) : 0.f;
    };
    float nx = 0.f, nla = 0.f, nda = 0.f;
    if ((int)blockIdx.x < p.num_tiles) fetch(blockIdx.x, nx, nla, nda);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int n = t / per_sample, r = t - n * per_sample;
      const int i0 = (r % p.tiles_1d) * 128, j0 = (r / p.tiles_1d) * 128;
      float* cv = colv + (it & 1) * 256;
      BT_T0();
      if (HAS_B) {
        // per-column normalisers of this j-tile: -lse_b log2(e) (so that P_b = exp2(fma(S, log2e, .))) and delta_b.
        // Double buffered by tile parity; the named barrier also keeps a fast warp from overwriting the buffer two
        // tiles ahead while a slow one still reads it.
        cv[et] = (et < 128) ? -nx * kLog2e : nx;
        named_bar_sync(1, 256);
      }
      const int i = i0 + quad * 32 + lane;
      const bool vi = i < p.L;
      const float nlse_a = -nla * kLog2e;
      const float del_a = nda;
      const uint32_t acc = it % kAccSets, aph = (it / kAccSets) & 1;
      const uint32_t taddr = tmem + acc * kAccStride + ((uint32_t)(quad * 32) << 16) + (uint32_t)half * 64;
      // element offset of (first row of this warp, first column of this warp's half) in the [N][Lp][Lp] outputs
      const size_t tile_off = ((size_t)n * p.Lp + i0 + quad * 32) * p.Lp + j0 + half * 64;
      const bool ragged = (i0 + 128 > p.L) || (j0 + 128 > p.L);   // tile-uniform: interior tiles need no masking
      if (t + (int)gridDim.x < p.num_tiles) fetch(t + gridDim.x, nx, nla, nda);
      BT_ACC(3);
      warp_mbar_wait(d_full + acc, aph, lane, 42);
      BT_ACC(4);
      tc_fence_after();
      // Chunk 0 (32 columns) is combined in place (its packed results fit in half of the registers it arrived in), then
      // chunk 1 is pulled out of TMEM and the accumulators are released BEFORE anything is written to memory, so the
      // next tile's MMAs overlap both flushes and the second combine.
      auto combine = [&](uint32_t (&sv)[32], uint32_t (&av)[32], uint32_t (&bv)[32], int jl0, auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
        for (int g = 0; g < 8; ++g) {          // 4 columns per step
          float4 nl = make_float4(0.f, 0.f, 0.f, 0.f), db = nl;
          if (HAS_B) {
            nl = *reinterpret_cast<const float4*>(cv + jl0 + 4 * g);
            db = *reinterpret_cast<const float4*>(cv + 128 + jl0 + 4 * g);
          }
          const float nlv[4] = {nl.x, nl.y, nl.z, nl.w}, dbv[4] = {db.x, db.y, db.z, db.w};
          float d[4], pa[4], pb[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 4 * g + e;
            const float sx = __uint_as_float(sv[k]);
            float a = fast_exp2(fmaf(sx, kLog2e, nlse_a));
            float b = HAS_B ? fast_exp2(fmaf(sx, kLog2e, nlv[e])) : 0.f;
            if (MASKED) {
              const bool ok = vi && (j0 + jl0 + k) < p.L;
              a = ok ? a : 0.f;
              b = ok ? b : 0.f;
            }
            float dd = a * (__uint_as_float(av[k]) - del_a);
            if (HAS_B) dd = fmaf(b, __uint_as_float(bv[k]) - dbv[e], dd);
            d[e] = dd; pa[e] = a; pb[e] = b;
          }
          // columns 4g..4g+3 are consumed: the first half of each array receives the packed results
          // (av <- dS, bv <- P_b, sv <- P_a)
          av[2 * g] = pack_bf16x2(d[0], d[1]);  av[2 * g + 1] = pack_bf16x2(d[2], d[3]);
          sv[2 * g] = pack_bf16x2(pa[0], pa[1]); sv[2 * g + 1] = pack_bf16x2(pa[2], pa[3]);
          if (HAS_B) { bv[2 * g] = pack_bf16x2(pb[0], pb[1]); bv[2 * g + 1] = pack_bf16x2(pb[2], pb[3]); }
        }
      };
      const int jl = half * 64;
      uint32_t s0[32], a0[32], b0[32], s1[32], a1[32], b1[32];
      tmem_ld32(taddr, s0);
      tmem_ld32(taddr + 128, a0);
      if (HAS_B) tmem_ld32(taddr + 256, b0);
      tmem_ld_wait();
      if (ragged) combine(s0, a0, b0, jl, std::true_type{}); else combine(s0, a0, b0, jl, std::false_type{});
      tmem_ld32(taddr + 32, s1);
      tmem_ld32(taddr + 128 + 32, a1);
      if (HAS_B) tmem_ld32(taddr + 256 + 32, b1);
      tmem_ld_wait();
      tc_fence_before();
      warp_mbar_arrive(d_empty + acc, lane);   // everything this warp needs from TMEM is in registers
      flush(p.ds, tile_off, a0);
      if (HAS_B) flush(p.pb, tile_off, b0);
      if (p.pa != nullptr) flush(p.pa, tile_off, s0);
      if (ragged) combine(s1, a1, b1, jl + 32, std::true_type{}); else combine(s1, a1, b1, jl + 32, std::false_type{});
      flush(p.ds, tile_off + 32, a1);
      if (HAS_B) flush(p.pb, tile_off + 32, b1);
      if (p.pa != nullptr) flush(p.pa, tile_off + 32, s1);
      BT_ACC(5);
    }
  }
  BT_REPORT();
  tc_fence_before();
  __syncthreads();
  if (warp == kTileMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// 16-bit [N][Lp][Lp] -> transposed per sample (32x32 tiles through shared memory); counterpart gradients only
__global__ void __launch_bounds__(256) transpose16_kernel(const unsigned short* __restrict__ src,
                                                          unsigned short* __restrict__ dst, int Lp) {
  __shared__ unsigned short t[32][34];
  const size_t base = (size_t)blockIdx.z * Lp * Lp;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = blockIdx.y * 32 + (threadIdx.x >> 5);
#pragma unroll
  for (int k = 0; k < 4; ++k) t[(threadIdx.x >> 5) + 8 * k][threadIdx.x & 31] = src[base + (size_t)(y0 + 8 * k) * Lp + x];
  __syncthreads();
  const int xo = blockIdx.y * 32 + (threadIdx.x & 31);
  const int yo0 = blockIdx.x * 32 + (threadIdx.x >> 5);
#pragma unroll
  for (int k = 0; k < 4; ++k) dst[base + (size_t)(yo0 + 8 * k) * Lp + xo] = t[threadIdx.x & 31][(threadIdx.x >> 5) + 8 * k];
}

// W [C_out][C_in] fp32 -> Wt [C_in][C_out] bf16 (B operand of dA += dQ W: dA[i][c] = sum_d dQt[i][d] Wt[c][d])
__global__ void transpose_w_kernel(const float* __restrict__ w, unsigned short* __restrict__ wt) {
  const int c = blockIdx.x;       // C_in
  const int d = threadIdx.x;      // C_out
  wt[c * kC + d] = cvt16<true>(w[d * kC + c]);
}

}  // namespace coattn
