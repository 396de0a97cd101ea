// sm_100a kernels of the co-attention hot path (reference: rgbd_segmentation_RAA.py:150-187 / :204-238).
//
//   prep_kernel     fp32 NCHW features -> bf16 operands in both layouts, zero padded to Lp   (:154-158)
//   project_kernel  Qt = At * W^T on tcgen05 (the nn.Linear of :159)                          (:159)
//   attend_kernel   fused affinity + softmax + attend, flash style, S never leaves the SM     (:160-170)
//   gate_kernel     1x1 gate conv + sigmoid + scale + concat with the original features       (:177-187)
//
// Notation: per sample A = V_a [C, L], B = V_b [C, L] (L = H'W' contiguous), Q = W A.
//   S[i,j]   = sum_c Q[c,i] B[c,j]
//   Z_a[:,i] = sum_j B[:,j] softmax_j(S[i,:])[j]      (pass 0: queries Q,  keys B, values B)
//   Z_b[:,j] = sum_i A[:,i] softmax_i(S[:,j])[i]      (pass 1: queries B,  keys Q, values A)
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace coattn {

constexpr int kC = 256;       // channels (all_channel of the reference ctor, :22)
constexpr int kLPad = 256;    // spatial padding granule of the 16-bit workspace (one CTA-pair query tile)

// ==============================================================================================
// prep: cast + transpose + pad
// ==============================================================================================
// src  : fp32 [C][L]              (one sample of V_a or V_b)
// x16  : bf16 [C][Lp]             same orientation (K-major "values" operand of the attend GEMM)
// xt   : bf16 [Lp][C]             transposed       (K-major "queries/keys" operand of the affinity GEMM)
constexpr int kPrepTileL = 64;
constexpr int kPrepThreads = 256;
constexpr int kPrepStride = kC + 2;  // bf16 elements; 129 words -> conflict-free column writes

struct PrepParams {
  const float* va;   // [N][C][L]
  const float* vb;   // [N][C][L]
  unsigned short* at;   // [N][Lp][C]   16-bit operands (f16 or bf16 bit patterns)
  unsigned short* bt;   // [N][Lp][C]
  unsigned short* a16;  // [N][C][Lp]
  unsigned short* b16;  // [N][C][Lp]
  int L, Lp;
  int only_b;            // 1: blockIdx.y indexes samples of V_b only (V_a is handled by project_fused_kernel)
};

template <bool BF16>
__global__ void __launch_bounds__(kPrepThreads) prep_kernel(PrepParams p) {
  __shared__ __align__(16) unsigned short tile[kPrepTileL * kPrepStride];
  const int n = p.only_b ? blockIdx.y : (blockIdx.y >> 1);
  const int which = p.only_b ? 1 : (blockIdx.y & 1);  // 0: A, 1: B
  const int l0 = blockIdx.x * kPrepTileL;
  const float* src = (which ? p.vb : p.va) + (size_t)n * kC * p.L;
  unsigned short* x16 = (which ? p.b16 : p.a16) + (size_t)n * kC * p.Lp;
  unsigned short* xt = (which ? p.bt : p.at) + (size_t)n * p.Lp * kC;

  const int tl = threadIdx.x & 63;   // position within the tile
  const int tc = threadIdx.x >> 6;   // 0..3
  const int l = l0 + tl;
  const bool valid = l < p.L;
#pragma unroll 8
  for (int k = 0; k < kC / 4; ++k) {
    const int c = tc + 4 * k;
    const float v = valid ? __ldg(src + (size_t)c * p.L + l) : 0.0f;
    const unsigned short mine = cvt16<BF16>(v);
    tile[tl * kPrepStride + c] = mine;
    // pack neighbouring positions: even lanes store 4 bytes
    const unsigned short next = __shfl_down_sync(0xffffffffu, mine, 1);
    if ((tl & 1) == 0) {
      *reinterpret_cast<uint32_t*>(x16 + (size_t)c * p.Lp + l) = (uint32_t)mine | ((uint32_t)next << 16);
    }
  }
  __syncthreads();
  const int cp = threadIdx.x & 127;  // channel pair
  const int r0 = threadIdx.x >> 7;   // 0..1
#pragma unroll 8
  for (int k = 0; k < kPrepTileL / 2; ++k) {
    const int r = r0 + 2 * k;
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&tile[r * kPrepStride + 2 * cp]);
    *reinterpret_cast<uint32_t*>(xt + (size_t)(l0 + r) * kC + 2 * cp) = w;
  }
}

// Vectorised variant for L % 4 == 0 (and 16-byte aligned inputs): float4 loads, 8-byte / 16-byte stores.
constexpr int kPrepStrideV = kC + 8;  // 16-bit elements; 528-byte rows keep 16-byte alignment for uint4 reads

template <bool BF16>
__global__ void __launch_bounds__(kPrepThreads) prep_kernel_vec4(PrepParams p) {
  __shared__ __align__(16) unsigned short tile[kPrepTileL * kPrepStrideV];
  const int n = p.only_b ? blockIdx.y : (blockIdx.y >> 1);
  const int which = p.only_b ? 1 : (blockIdx.y & 1);  // 0: A, 1: B
  const int l0 = blockIdx.x * kPrepTileL;
  const float* src = (which ? p.vb : p.va) + (size_t)n * kC * p.L;
  unsigned short* x16 = (which ? p.b16 : p.a16) + (size_t)n * kC * p.Lp;
  unsigned short* xt = (which ? p.bt : p.at) + (size_t)n * p.Lp * kC;

  const int col4 = threadIdx.x & 15;   // float4 column inside the 64-position tile
  const int crow = threadIdx.x >> 4;   // 0..15
  const int l = l0 + 4 * col4;
  const bool valid = l < p.L;          // L % 4 == 0: a float4 is entirely inside or outside
#pragma unroll 4
  for (int k = 0; k < kC / 16; ++k) {
    const int c = crow + 16 * k;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) v = __ldcs(reinterpret_cast<const float4*>(src + (size_t)c * p.L + l));
    const uint32_t lo = pack16x2<BF16>(v.x, v.y), hi = pack16x2<BF16>(v.z, v.w);
    *reinterpret_cast<uint2*>(x16 + (size_t)c * p.Lp + l) = make_uint2(lo, hi);
    unsigned short* t = tile + (4 * col4) * kPrepStrideV + c;
    t[0] = (unsigned short)(lo & 0xFFFFu);
    t[kPrepStrideV] = (unsigned short)(lo >> 16);
    t[2 * kPrepStrideV] = (unsigned short)(hi & 0xFFFFu);
    t[3 * kPrepStrideV] = (unsigned short)(hi >> 16);
  }
  __syncthreads();
  const int c8 = threadIdx.x & 31;   // group of 8 channels
  const int r0 = threadIdx.x >> 5;   // 0..7
#pragma unroll
  for (int k = 0; k < kPrepTileL / 8; ++k) {
    const int r = r0 + 8 * k;
    const uint4 w = *reinterpret_cast<const uint4*>(&tile[r * kPrepStrideV + 8 * c8]);
    *reinterpret_cast<uint4*>(xt + (size_t)(l0 + r) * kC + 8 * c8) = w;
  }
}

// fp32 [C][C] -> 16-bit [C][C]
template <bool BF16>
__global__ void cast_w_kernel(const float* __restrict__ w, unsigned short* __restrict__ w16, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w16[i] = cvt16<BF16>(w[i]);
}

// ==============================================================================================
// shared role layout of the two tcgen05 kernels: warps 0-3 own TMEM lane quadrants 0-3
// (softmax / epilogue), warp 4 = TMA producer, warp 5 = MMA issuer + TMEM allocator.
// ==============================================================================================
constexpr int kNumThreads = 192;
constexpr int kProducerWarp = 4;
constexpr int kMmaWarp = 5;

__device__ __forceinline__ uint8_t* align_1024(uint8_t* p) {
  const uint32_t a = smem_u32(p);
  return p + ((1024u - (a & 1023u)) & 1023u);
}

// One lane polls the barrier, the warp reconverges behind it (4 pollers per CTA instead of 128).
__device__ __forceinline__ void warp_mbar_wait(uint64_t* bar, uint32_t parity, int lane, int tag) {
  if (lane == 0) mbar_wait(bar, parity, tag);
  __syncwarp();
}
// Every lane has fenced its own tcgen05 traffic; one elected arrival per warp publishes it.
__device__ __forceinline__ void warp_mbar_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// ==============================================================================================
// project: Qt[n][i][co] = sum_ci At[n][i][ci] * W[co][ci]          (bf16 x bf16 -> fp32 -> bf16)
// one CTA per 128-row tile; A tile and the whole W live in shared memory (64 KB + 128 KB)
// ==============================================================================================
constexpr int kProjSmemBytes = 64 * 1024 + 128 * 1024 + 1024 /*align slack*/ + 64 /*barriers*/;

struct ProjectParams {
  unsigned short* qt;  // [N][Lp][C]
  int Lp;
};

template <bool BF16>
__global__ void __launch_bounds__(kNumThreads, 1)
project_kernel(const __grid_constant__ CUtensorMap tmap_at,  // [N*Lp][C], box {64, 128}
               const __grid_constant__ CUtensorMap tmap_w,   // [C][C],    box {64, 256}
               ProjectParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  uint8_t* sA = smem;               // 4 k-blocks x [128 rows x 128 B]
  uint8_t* sW = smem + 64 * 1024;   // 4 k-blocks x [256 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 192 * 1024);
  uint64_t* ab_full = bars + 0;
  uint64_t* d_full = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.y;
  const int row0 = n * p.Lp + blockIdx.x * 128;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_at);
    tma_prefetch_desc(&tmap_w);
    mbar_init(ab_full, 1);
    mbar_init(d_full, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      mbar_arrive_expect_tx(ab_full, 192 * 1024);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        tma_load_2d(sA + kb * 16384, &tmap_at, ab_full, kb * 64, row0);
        tma_load_2d(sW + kb * 32768, &tmap_w, ab_full, kb * 64, 0);
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_wait(ab_full, 0, 100);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc_16(128, 256, BF16);
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const uint64_t ad = make_sdesc_k_sw128(smem_u32(sA + (kk >> 2) * 16384 + (kk & 3) * 32));
        const uint64_t bd = make_sdesc_k_sw128(smem_u32(sW + (kk >> 2) * 32768 + (kk & 3) * 32));
        umma_ss(tmem, ad, bd, idesc, kk > 0);
      }
      umma_commit(d_full);
    }
  } else {
    // epilogue: thread = one output row
    mbar_wait(d_full, 0, 101);
    tc_fence_after();
    const int row = warp * 32 + lane;
    unsigned short* dst = p.qt + (size_t)(row0 + row) * kC;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int ch = 0; ch < 8; ++ch) {
      uint32_t v[32];
      tmem_ld32(taddr + ch * 32, v);
      tmem_ld_wait();
      uint4* d4 = reinterpret_cast<uint4*>(dst + ch * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = pack16x2<BF16>(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
        o.y = pack16x2<BF16>(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
        o.z = pack16x2<BF16>(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
        o.w = pack16x2<BF16>(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
        d4[q] = o;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ==============================================================================================
// project_fused: the A side of prep and the W projection in one kernel (L % 4 == 0 not required, but the
// launcher only uses it when the fp32 rows are 4-byte aligned, i.e. always).  Per 128-position tile:
//   warps 0-3 read the fp32 features [256 ch][128 pos] (thread = position, coalesced per channel row), convert to
//   16 bit, write A16 [C][Lp] (values operand of pass 1) and build the At tile directly in shared memory in the
//   K-major / 128-byte-swizzle layout the UMMA descriptor expects (16-byte chunk index XOR row % 8) -- At never
//   exists in HBM; W16 arrives by TMA; 16 x tcgen05.mma; the same warps drain TMEM into Qt [Lp][C].
// ==============================================================================================
struct ProjectFusedParams {
  const float* va;        // [N][C][L]
  unsigned short* a16;    // [N][C][Lp]
  unsigned short* qt;     // [N][Lp][C]
  int L, Lp;
};

constexpr int kPFStages = 2;
constexpr int kPFStageBytes = 16384 + 32768;   // At k-block [128 pos x 64 ch] + W k-block [256 out x 64 in]
constexpr int kProjFusedSmemBytes = kPFStages * kPFStageBytes + 1024 + 128;

// K (input channels) is streamed in four 64-channel blocks through a 2-stage ring, so that the conversion of block
// kb+1 overlaps the MMAs of block kb and two CTAs fit on an SM (96 KB shared memory, 256 TMEM columns each).
template <bool BF16>
__global__ void __launch_bounds__(kNumThreads, 2)
project_fused_kernel(const __grid_constant__ CUtensorMap tmap_w,   // [C][C], box {64, 256}
                     ProjectFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPFStages * kPFStageBytes);
  uint64_t* w_full = bars + 0;              // [2] TMA bytes of a W k-block
  uint64_t* a_full = bars + kPFStages;      // [2] 4 arrivals (one per converting warp)
  uint64_t* empty = bars + 2 * kPFStages;   // [2] MMAs of the stage completed
  uint64_t* d_full = bars + 3 * kPFStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.y;
  const int l0 = blockIdx.x * 128;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < kPFStages; ++s) { mbar_init(w_full + s, 1); mbar_init(a_full + s, 4); mbar_init(empty + s, 1); }
    mbar_init(d_full, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      for (int kb = 0; kb < 4; ++kb) {
        const int s = kb & 1;
        mbar_wait(empty + s, ((kb >> 1) & 1) ^ 1, 103);
        mbar_arrive_expect_tx(w_full + s, 32768);
        tma_load_2d(smem + s * kPFStageBytes + 16384, &tmap_w, w_full + s, kb * 64, 0);
      }
    }
  } else if (warp == kMmaWarp) {
    constexpr uint32_t idesc = make_idesc_16(128, 256, BF16);
    const uint32_t base = smem_u32(smem);
    for (int kb = 0; kb < 4; ++kb) {
      const int s = kb & 1;
      const uint32_t ph = (kb >> 1) & 1;
      warp_mbar_wait(w_full + s, ph, lane, 100);
      warp_mbar_wait(a_full + s, ph, lane, 102);
      tc_fence_after();
      const uint64_t ad0 = make_sdesc_k_sw128(base + s * kPFStageBytes);
      const uint64_t bd0 = make_sdesc_k_sw128(base + s * kPFStageBytes + 16384);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem, ad0 + 2 * k, bd0 + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(empty + s);
        if (kb == 3) umma_commit(d_full);
      }
      __syncwarp();
    }
  } else {
    // ---- convert: thread = two neighbouring positions (rows 2q, 2q+1 of the At tile) x 32 channels of the k-block.
    // Requires L even (float2 loads); the launcher falls back to the unfused kernels otherwise.
    const int q2 = (warp & 1) * 32 + lane;      // position pair 0..63
    const int chalf = warp >> 1;                // channels [32 * chalf, 32 * chalf + 32) of the k-block
    const int r0 = 2 * q2;
    const int lpos = l0 + r0;
    const bool valid2 = lpos < p.L;             // L even: both positions valid or both padding
    const float* src = p.va + (size_t)n * kC * p.L + lpos;
    unsigned short* a16 = p.a16 + (size_t)n * kC * p.Lp + lpos;
#pragma unroll 1
    for (int kb = 0; kb < 4; ++kb) {
      const int s = kb & 1;
      uint8_t* sA = smem + s * kPFStageBytes;
      float2 v[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const int c = kb * 64 + chalf * 32 + u;
        v[u] = valid2 ? __ldcs(reinterpret_cast<const float2*>(src + (size_t)c * p.L)) : make_float2(0.f, 0.f);
      }
      warp_mbar_wait(empty + s, ((kb >> 1) & 1) ^ 1, lane, 104);   // the MMAs that read this stage two blocks ago are done
#pragma unroll
      for (int g = 0; g < 4; ++g) {             // 8 channels = one 16-byte chunk of a row
        const int cl = chalf * 32 + g * 8;      // channel inside the k-block
        const int chunk = cl >> 3;
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          lo[t] = pack16x2<BF16>(v[g * 8 + 2 * t].x, v[g * 8 + 2 * t + 1].x);   // position r0
          hi[t] = pack16x2<BF16>(v[g * 8 + 2 * t].y, v[g * 8 + 2 * t + 1].y);   // position r0 + 1
        }
        *reinterpret_cast<uint4*>(sA + r0 * 128 + ((chunk ^ (r0 & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(sA + (r0 + 1) * 128 + ((chunk ^ ((r0 + 1) & 7)) << 4)) =
            make_uint4(hi[0], hi[1], hi[2], hi[3]);
#pragma unroll
        for (int u = 0; u < 8; ++u)     // A16[c][lpos .. lpos+1]: 4-byte stores, 128 contiguous bytes per warp
          *reinterpret_cast<uint32_t*>(a16 + (size_t)(kb * 64 + cl + u) * p.Lp) = pack16x2<BF16>(v[g * 8 + u].x, v[g * 8 + u].y);
      }
      fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor core (async proxy)
      warp_mbar_arrive(a_full + s, lane);
    }
    // ---- epilogue: thread = one output row of Qt
    warp_mbar_wait(d_full, 0, lane, 101);
    tc_fence_after();
    unsigned short* dst = p.qt + ((size_t)n * p.Lp + l0 + warp * 32 + lane) * kC;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int ch = 0; ch < 8; ++ch) {
      uint32_t v[32];
      tmem_ld32(taddr + ch * 32, v);
      tmem_ld_wait();
      uint4* d4 = reinterpret_cast<uint4*>(dst + ch * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = pack16x2<BF16>(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
        o.y = pack16x2<BF16>(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
        o.z = pack16x2<BF16>(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
        o.w = pack16x2<BF16>(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
        d4[q] = o;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ==============================================================================================
// MN-major path (default forward): the features are only CAST, never transposed.
//   cast_kernel        fp32 [N][C][L] -> 16-bit X[plane][N][C][Lp] (plane 0 = V_b, plane 1 = V_a), zero padded
//   project_mn_kernel  Q16[n][co][i] = sum_ci W[co][ci] A16[n][ci][i]  -> X plane 2, channel-major like its input:
//                      A operand = W16 (K-major), B operand = the A16 tile as an MN-major operand (positions contiguous)
// ==============================================================================================
struct CastParams {
  const float* va;
  const float* vb;
  unsigned short* x;   // [3][N][C][Lp]; planes 0 (V_b) and 1 (V_a) are written here
  int N, L, Lp;
  int Na;              // samples of V_a (= N, or the number of query frames when each is paired with several references)
  unsigned* status;    // status block of the workspace (include/coattn_b200.h, COATTN_STATUS_*) or null
  int first_plane;     // 0: both planes (grid.y = 2); 1 with grid.y = 1: V_a only
  const float* w;      // [C][C] fp32 similarity weights or null: the first C blocks of plane 0 also cast one row of W each
  unsigned short* w16; // [C][C] 16-bit (the cast_w launch of its own, 3 us + a launch gap per modality call, is gone)
};

// fp16 operand range guard: the block's largest |v| goes into the status block (atomicMax on the bits of a non-negative
// float orders like the float), and a value the fp16 conversion would clamp (> 65504, Inf, NaN) raises the overflow bit of
// its frame.  One shuffle reduction + at most three atomics per 256-thread block; nothing is done for bf16 operands
// (fp32 exponent range).
__device__ __forceinline__ void report_absmax(float m, bool nan, unsigned* status, int plane) {
  __shared__ float red_m[8];
  __shared__ int red_n[8];
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  nan = __any_sync(0xffffffffu, nan);
  if ((threadIdx.x & 31) == 0) { red_m[threadIdx.x >> 5] = m; red_n[threadIdx.x >> 5] = nan ? 1 : 0; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int bad = 0;
    for (int i = 0; i < 8; ++i) { m = fmaxf(m, red_m[i]); bad |= red_n[i]; }
    atomicMax(status + 1 + plane, __float_as_uint(m));
    if (bad || m > 65504.0f) atomicOr(status, plane ? 2u /*COATTN_STATUS_OVERFLOW_A*/ : 1u /*COATTN_STATUS_OVERFLOW_B*/);
  }
}

template <bool BF16, int VEC>
__global__ void __launch_bounds__(256) cast_kernel(CastParams p) {
  pdl_wait();
  const int row = blockIdx.x;                 // n * C + c
  const int plane = blockIdx.y + p.first_plane;   // 0: V_b, 1: V_a
  if (p.w != nullptr && blockIdx.y == 0 && row < kC) p.w16[row * kC + threadIdx.x] = cvt16<BF16>(p.w[row * kC + threadIdx.x]);
  if (plane == 1 && row >= p.Na * kC) return;
  const float* src = (plane ? p.va : p.vb) + (size_t)row * p.L;
  unsigned short* dst = p.x + ((size_t)plane * p.N * kC + row) * p.Lp;
  float amax = 0.f;
  bool nan = false;
  if constexpr (VEC == 4) {
    for (int i = threadIdx.x * 4; i < p.Lp; i += 256 * 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < p.L) v = __ldcs(reinterpret_cast<const float4*>(src + i));     // L % 4 == 0
      *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack16x2<BF16>(v.x, v.y), pack16x2<BF16>(v.z, v.w));
      if constexpr (!BF16) {
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));   // fmaxf drops NaN:
        nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);                          // tracked apart
      }
    }
  } else {
    // rows that are only 4-byte aligned (odd L: 61x81, 61x107): scalar loads, four pairs per thread in flight -- with one
    // pair per iteration a resident SM had 16 KB of loads outstanding and the kernel ran at 0.75 of the copy bandwidth
    for (int i0 = threadIdx.x * 2; i0 < p.Lp; i0 += 4 * 256 * 2) {
      float a[4], b[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * 512;
        a[k] = (i < p.L) ? __ldcs(src + i) : 0.f;
        b[k] = (i + 1 < p.L) ? __ldcs(src + i + 1) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * 512;
        if (i < p.Lp) *reinterpret_cast<uint32_t*>(dst + i) = pack16x2<BF16>(a[k], b[k]);
        if constexpr (!BF16) { amax = fmaxf(amax, fmaxf(fabsf(a[k]), fabsf(b[k]))); nan |= (a[k] != a[k]) | (b[k] != b[k]); }
      }
    }
  }
  if constexpr (!BF16) {
    if (p.status != nullptr) report_absmax(amax, nan, p.status, plane);
  }
}

// Producer side of the hot path (deeplab/deeplabv3_encoder.py:80-82): the ASPP tail
//   features = PReLU(BatchNorm_eval(bottleneck_conv_output))
// fused with the operand cast: ONE pass reads the conv output and writes the fp32 features (the concat's passthrough half,
// the auxiliary classifier and the reduce conv still want them) AND the zero-padded 16-bit plane of the workspace that
// project_mn / attend2 read -- the cast kernel and the two eager elementwise kernels (BN, PReLU) disappear.
//   y = x * scale[c] + shift[c]   (scale = gamma / sqrt(var + eps), shift = beta - mean * scale: eval-mode BN)
//   y = y >= 0 ? y : slope * y    (nn.PReLU() with its single parameter)
struct TailParams {
  const float* x;        // [N][C][L] bottleneck conv output (bias included)
  const float* scale;    // [C]
  const float* shift;    // [C]
  const float* slope;    // [1]
  float* y;              // [N][C][L] fp32 features, or null
  unsigned short* plane; // [N][C][Lp] 16-bit operand plane of this frame
  int L, Lp;
  unsigned* status;      // status block or null (fp16 range guard, as in cast_kernel)
  int status_plane;      // 0: V_b, 1: V_a
};

template <bool BF16, int VEC>
__global__ void __launch_bounds__(256) aspp_tail_kernel(TailParams p) {
  const int row = blockIdx.x;                 // n * C + c
  const int c = row % kC;
  const float sc = __ldg(p.scale + c), sh = __ldg(p.shift + c), a = __ldg(p.slope);
  const float* src = p.x + (size_t)row * p.L;
  float* dstf = p.y ? p.y + (size_t)row * p.L : nullptr;
  unsigned short* dst = p.plane + (size_t)row * p.Lp;
  float amax = 0.f;
  bool nan = false;
  auto act = [&](float v) { const float t = fmaf(v, sc, sh); return t >= 0.f ? t : a * t; };
  if constexpr (VEC == 4) {
    for (int i = threadIdx.x * 4; i < p.Lp; i += 256 * 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < p.L) {                          // L % 4 == 0
        const float4 x = __ldcs(reinterpret_cast<const float4*>(src + i));
        v = make_float4(act(x.x), act(x.y), act(x.z), act(x.w));
        if (dstf) *reinterpret_cast<float4*>(dstf + i) = v;
      }
      *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack16x2<BF16>(v.x, v.y), pack16x2<BF16>(v.z, v.w));
      if constexpr (!BF16) {
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
      }
    }
  } else {
    // four pairs per thread in flight, like cast_kernel's scalar path
    for (int i0 = threadIdx.x * 2; i0 < p.Lp; i0 += 4 * 256 * 2) {
      float x0[4], x1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * 512;
        x0[k] = (i < p.L) ? __ldcs(src + i) : 0.f;
        x1[k] = (i + 1 < p.L) ? __ldcs(src + i + 1) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * 512;
        const float v0 = (i < p.L) ? act(x0[k]) : 0.f;
        const float v1 = (i + 1 < p.L) ? act(x1[k]) : 0.f;
        if (dstf) { if (i < p.L) dstf[i] = v0; if (i + 1 < p.L) dstf[i + 1] = v1; }
        if (i < p.Lp) *reinterpret_cast<uint32_t*>(dst + i) = pack16x2<BF16>(v0, v1);
        if constexpr (!BF16) { amax = fmaxf(amax, fmaxf(fabsf(v0), fabsf(v1))); nan |= (v0 != v0) | (v1 != v1); }
      }
    }
  }
  if constexpr (!BF16) {
    if (p.status != nullptr) report_absmax(amax, nan, p.status, p.status_plane);
  }
}

// 16-bit features (coattn_forward16) whose rows cannot be read by TMA directly (L % 8 != 0 or a base pointer that is
// not 16-byte aligned): copied into the same zero-padded planes the cast writes.  CastParams::va / vb then point to
// 16-bit data.  Only the fallback of that entry point; aligned 16-bit features are consumed in place.
// WORDS = true: 4-byte accesses (needs 4-byte aligned base pointers); a row of odd L starts on an odd element in every
// second row, where a destination word is assembled from the halves of two source words.
template <bool WORDS>
__global__ void __launch_bounds__(256) pad16_kernel(CastParams p) {
  const int row = blockIdx.x;                 // n * C + c
  const int plane = blockIdx.y;               // 0: V_b, 1: V_a
  if (plane == 1 && row >= p.Na * kC) return;
  const unsigned short* base = reinterpret_cast<const unsigned short*>(plane ? p.va : p.vb);
  const size_t s0 = (size_t)row * p.L;        // first element of the row
  unsigned short* dst = p.x + ((size_t)plane * p.N * kC + row) * p.Lp;
  if constexpr (WORDS) {
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(base);
    uint32_t* dst32 = reinterpret_cast<uint32_t*>(dst);      // Lp is a multiple of 256: rows of the planes are 512-byte aligned
    const bool odd = (s0 & 1) != 0;
    for (int i = threadIdx.x; i < p.Lp / 2; i += 256) {
      const int e = 2 * i;
      uint32_t w = 0;
      if (e + 1 < p.L) {                      // both elements inside the row
        const size_t wi = (s0 + e) >> 1;
        w = odd ? (__ldcs(src32 + wi) >> 16) | (__ldcs(src32 + wi + 1) << 16) : __ldcs(src32 + wi);
      } else if (e < p.L) {                   // last element of an odd-length row: never touch the word after it
        w = __ldcs(base + s0 + e);
      }
      dst32[i] = w;
    }
  } else {
    for (int i = threadIdx.x; i < p.Lp; i += 256) dst[i] = (i < p.L) ? __ldcs(base + s0 + i) : (unsigned short)0;
  }
}

constexpr int kProjMnTile = 64;            // positions per tile
constexpr int kProjMnXStages = 2;
constexpr int kProjMnThreads = 320;          // warps 0-7 epilogue (TMEM lane quadrant x m-tile), 8 TMA producer, 9 MMA issuer
constexpr int kProjMnProducerWarp = 8;
constexpr int kProjMnMmaWarp = 9;
constexpr int kProjMnSmemBytes = 128 * 1024 + kProjMnXStages * 32 * 1024 + 8 * 4096 /*store staging*/ + 1024 + 128;

struct ProjectMnParams {
  unsigned short* q16;   // X plane 2: [N][C][Lp]
  int Lp;
  int tiles_per_sample;  // Lp / 64
  int num_tiles;         // N * Lp / 64
  int a_row0_base;       // first row of plane A16 in the X tensor map (= 1 * N * C)
  unsigned* status;      // status block of the workspace or null: |Q| beyond the fp16 range raises COATTN_STATUS_OVERFLOW_Q
};

// Persistent: W16 (128 KB) is loaded once per CTA and stays in shared memory; 64-position tiles of A16 stream through
// a 2-stage ring; two TMEM accumulator sets let the drain of tile t overlap the MMAs of tile t+1.
template <bool BF16>
__global__ void __launch_bounds__(kProjMnThreads, 1)
project_mn_kernel(const __grid_constant__ CUtensorMap tmap_w,   // W16 [C][C], box {64, 128}
                  const __grid_constant__ CUtensorMap tmap_x,   // X [3*N*C][Lp], box {64, 256}
                  ProjectMnParams p) {
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  uint8_t* sW = smem;                 // 2 m-tiles x 4 k-blocks x [128 rows x 128 B] = 128 KB
  uint8_t* sX = smem + 128 * 1024;    // stages x [256 channel rows x 128 B (64 positions)]
  uint8_t* sOut = sX + kProjMnXStages * 32768;      // 8 warps x [32 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + 8 * 4096);
  uint64_t* w_full = bars + 0;
  uint64_t* x_full = bars + 1;                      // [2]
  uint64_t* x_empty = x_full + kProjMnXStages;      // [2]
  uint64_t* d_full = x_empty + kProjMnXStages;      // [2]
  uint64_t* d_empty = d_full + 2;                   // [2] 8 arrivals (epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kProjMnProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    mbar_init(w_full, 1);
    for (int s = 0; s < kProjMnXStages; ++s) { mbar_init(x_full + s, 1); mbar_init(x_empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(d_full + b, 1); mbar_init(d_empty + b, 8); }
    fence_mbar_init();
  }
  if (warp == kProjMnMmaWarp) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kProjMnProducerWarp) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 128 * 1024);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sW + (mt * 4 + kb) * 16384, &tmap_w, w_full, kb * 64, mt * 128);
      uint32_t cnt = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++cnt) {
        const int n = t / p.tiles_per_sample;
        const int i0 = (t - n * p.tiles_per_sample) * kProjMnTile;
        const uint32_t s = cnt % kProjMnXStages, ph = (cnt / kProjMnXStages) & 1;
        mbar_wait(x_empty + s, ph ^ 1, 110);
        mbar_arrive_expect_tx(x_full + s, 32768);
        tma_load_2d(sX + s * 32768, &tmap_x, x_full + s, i0, p.a_row0_base + n * kC);
      }
    }
  } else if (warp == kProjMnMmaWarp) {
    constexpr uint32_t idesc = make_idesc_16_major(128, kProjMnTile, BF16, false, true);
    const uint32_t wbase = smem_u32(sW);
    const uint32_t xbase = smem_u32(sX);
    warp_mbar_wait(w_full, 0, lane, 100);
    uint32_t cnt = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++cnt) {
      const uint32_t s = cnt % kProjMnXStages, ph = (cnt / kProjMnXStages) & 1;
      const uint32_t b = cnt & 1, dph = (cnt >> 1) & 1;
      warp_mbar_wait(x_full + s, ph, lane, 111);
      warp_mbar_wait(d_empty + b, dph ^ 1, lane, 112);
      tc_fence_after();
      const uint64_t xd0 = make_sdesc_mn_sw128(xbase + s * 32768, 32768, 1024);
      if (elect_one()) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            const uint64_t ad = make_sdesc_k_sw128(wbase + (mt * 4 + (kk >> 2)) * 16384 + (kk & 3) * 32);
            umma_ss(tmem + b * 128 + mt * 64, ad, xd0 + (uint64_t)((kk * 2048) >> 4), idesc, kk > 0);
          }
        }
        umma_commit(x_empty + s);
        umma_commit(d_full + b);
      }
      __syncwarp();
    }
  } else {
    // epilogue: warp = (TMEM lane quadrant, m-tile); thread = one output channel, 64 positions = 128 bytes, staged
    // through a swizzled 4 KB block per warp so that every global store instruction writes 4 full 128-byte rows
    const int quad = warp & 3, mt = warp >> 2;
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + mt * 64;
    uint8_t* stg = sOut + warp * 4096;
    const int rd_r = lane >> 3, rd_c = lane & 7;
    uint32_t cnt = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++cnt) {
      const int n = t / p.tiles_per_sample;
      const int i0 = (t - n * p.tiles_per_sample) * kProjMnTile;
      const uint32_t b = cnt & 1, dph = (cnt >> 1) & 1;
      warp_mbar_wait(d_full + b, dph, lane, 101);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(taddr + b * 128, v0);
      tmem_ld32(taddr + b * 128 + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      warp_mbar_arrive(d_empty + b, lane);      // accumulators are in registers: the next tile's MMAs may overwrite them
      if constexpr (!BF16) {
        // fp16 range guard for Q = W V_a (the pack below saturates at +-65504): !(x <= limit) also catches NaN
        float qm = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) qm = fmaxf(qm, fmaxf(fabsf(__uint_as_float(v0[k])), fabsf(__uint_as_float(v1[k]))));
        if (p.status != nullptr && __any_sync(0xffffffffu, !(qm <= 65504.0f)) && lane == 0) atomicOr(p.status, 4u);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) =
            make_uint4(pack16x2<BF16>(__uint_as_float(v0[8 * q + 0]), __uint_as_float(v0[8 * q + 1])),
                       pack16x2<BF16>(__uint_as_float(v0[8 * q + 2]), __uint_as_float(v0[8 * q + 3])),
                       pack16x2<BF16>(__uint_as_float(v0[8 * q + 4]), __uint_as_float(v0[8 * q + 5])),
                       pack16x2<BF16>(__uint_as_float(v0[8 * q + 6]), __uint_as_float(v0[8 * q + 7])));
        *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + q) ^ (lane & 7)) << 4)) =
            make_uint4(pack16x2<BF16>(__uint_as_float(v1[8 * q + 0]), __uint_as_float(v1[8 * q + 1])),
                       pack16x2<BF16>(__uint_as_float(v1[8 * q + 2]), __uint_as_float(v1[8 * q + 3])),
                       pack16x2<BF16>(__uint_as_float(v1[8 * q + 4]), __uint_as_float(v1[8 * q + 5])),
                       pack16x2<BF16>(__uint_as_float(v1[8 * q + 6]), __uint_as_float(v1[8 * q + 7])));
      }
      __syncwarp();
      unsigned short* dst = p.q16 + ((size_t)n * kC + mt * 128 + quad * 32) * p.Lp + i0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = rd_r + 4 * k;
        const uint4 v = *reinterpret_cast<const uint4*>(stg + r * 128 + ((rd_c ^ (r & 7)) << 4));
        *reinterpret_cast<uint4*>(dst + (size_t)r * p.Lp + rd_c * 8) = v;
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kProjMnMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ==============================================================================================
// attend: persistent flash-style kernel over work items (sample n, pass p, 128-row query tile)
//
//   warps 0-3  softmax + drain: thread r owns query row r (TMEM lane r); they also stage the query tile:
//              each thread copies its 512-byte row global -> registers -> TMEM (A operand of the affinity MMA)
//   warp 4     TMA producer of key tiles   [64 positions x 256 channels]  (4-stage ring)
//   warp 5     MMA issuer (one thread) and TMEM allocator
//   warp 6     TMA producer of value tiles [256 channels x 64 positions]  (3-stage ring)
//
//   TMEM columns: [0,256) O accumulator (fp32) | [256,320) S/P buffer 0 | [320,384) S/P buffer 1 |
//                 [384,512) query tile, 128 x 256 16-bit values packed two per column
//   MMA order per item: S(0) | S(1) PV(0) | S(2) PV(1) | ... | PV(T-1); P(j) overwrites S(j) in place.
// ==============================================================================================
constexpr int kBM = 128;      // query rows per tile (TMEM lanes)
constexpr int kBN = 64;       // key/value positions per step
constexpr int kKStages = 4;
constexpr int kVStages = 3;
constexpr int kKBytes = kBN * kC * 2;      // 32 KB : 4 k-blocks x [ 64 rows x 128 B]
constexpr int kVBytes = kC * kBN * 2;      // 32 KB : [256 rows x 128 B]
constexpr int kAttendThreads = 224;
constexpr int kKProducerWarp = 4;
constexpr int kVProducerWarp = 6;
constexpr int kAttendSmemBytes = kKStages * kKBytes + kVStages * kVBytes + 1024 + 256;
constexpr uint32_t kTmemColsO = 0;         // O accumulator: 256 fp32 columns
constexpr uint32_t kTmemColsS = 256;       // two S/P buffers of kBN columns
constexpr uint32_t kTmemColsQ = 384;       // query tile: 128 columns of packed pairs
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units: O is only rescaled when the row max jumps by > 2^8

struct AttendParams {
  const unsigned short* t;  // [2][N][Lp][C] 16-bit: t[0] = Bt, t[1] = Qt (queries of pass p are t[1-p])
  float* z;     // [2][N][C][L]  raw attended features (pass 0: Z_a, pass 1: Z_b); may be null when cat_* are set
  float* lse;   // [2][N][L]     log-sum-exp of each softmax row (natural log)
  // fused gate epilogue (:177-184): when cat_a/cat_b are set the drain also writes Z * sigmoid(g.Z + b) into the
  // first C channels of the concat tensors and the gate values into mask
  float* cat_a;          // [N][2C][L] or null
  float* cat_b;          // [N][2C][L] or null
  float* mask;           // [2][N][L] or null
  const float* gate_w;   // [C]
  const float* gate_b;   // [1] or null
  int N, L, Lp;
  int q_tiles;   // ceil(L / 128)
  int kv_tiles;  // ceil(L / 64)
  int num_items; // 2 * N * q_tiles
  long long* trace;  // debug only (COATTN_TRACE builds): clock64 stamps of block 0
};

#ifdef COATTN_TRACE
#define TRACE_MMA(slot) do { if (blockIdx.x == 0 && lane == 0 && tcount < 96) p.trace[tcount * 8 + (slot)] = clock64(); } while (0)
#define TRACE_SM(slot) do { if (blockIdx.x == 0 && warp == 0 && lane == 0 && scount < 96) p.trace[scount * 8 + (slot)] = clock64(); } while (0)
#else
#define TRACE_MMA(slot) do {} while (0)
#define TRACE_SM(slot) do {} while (0)
#endif

template <bool BF16>
__global__ void __launch_bounds__(kAttendThreads, 1)
attend_kernel(const __grid_constant__ CUtensorMap tmap_k,  // T  [2*N*Lp][C],  box {64, 64}
              const __grid_constant__ CUtensorMap tmap_v,  // VV [2*N*C][Lp],  box {64, 256}
              AttendParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  uint8_t* sK = smem;
  uint8_t* sV = sK + kKStages * kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kVStages * kVBytes);
  uint64_t* q_full = bars + 0;                 // query tile staged in TMEM (one arrival per softmax warp)
  uint64_t* k_full = bars + 1;                 // [kKStages]
  uint64_t* k_empty = k_full + kKStages;       // [kKStages]
  uint64_t* v_full = k_empty + kKStages;       // [kVStages]
  uint64_t* v_empty = v_full + kVStages;       // [kVStages]
  uint64_t* s_full = v_empty + kVStages;       // [2]
  uint64_t* p_full = s_full + 2;               // [2]
  uint64_t* o_full = p_full + 2;               // one completion per PV step
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kKProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 4);
    for (int s = 0; s < kKStages; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
    for (int s = 0; s < kVStages; ++s) { mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(s_full + b, 1); mbar_init(p_full + b, 4); }
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int T = p.kv_tiles;

  if (warp == kKProducerWarp) {
    // ------------------------------------------------------------------ TMA producer: key tiles
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int np = item / p.q_tiles;
        const int krow0 = ((np & 1) * p.N + (np >> 1)) * p.Lp;
        for (int j = 0; j < T; ++j, ++cnt) {
          const uint32_t s = cnt % kKStages, ph = (cnt / kKStages) & 1;
          mbar_wait(k_empty + s, ph ^ 1, 2);
          mbar_arrive_expect_tx(k_full + s, kKBytes);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)
            tma_load_2d(sK + s * kKBytes + kb * (kBN * 128), &tmap_k, k_full + s, kb * 64, krow0 + j * kBN);
        }
      }
    }
  } else if (warp == kVProducerWarp) {
    // ------------------------------------------------------------------ TMA producer: value tiles
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int np = item / p.q_tiles;
        const int vrow0 = ((np & 1) * p.N + (np >> 1)) * kC;
        for (int j = 0; j < T; ++j, ++cnt) {
          const uint32_t s = cnt % kVStages, ph = (cnt / kVStages) & 1;
          mbar_wait(v_empty + s, ph ^ 1, 3);
          mbar_arrive_expect_tx(v_full + s, kVBytes);
          tma_load_2d(sV + s * kVBytes, &tmap_v, v_full + s, j * kBN, vrow0);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the (warp-uniform) control flow so that descriptors live in uniform registers;
    // only the tcgen05.mma / tcgen05.commit instructions themselves are issued by one elected lane.
    constexpr uint32_t idesc_s = make_idesc_16(kBM, kBN, BF16);
    constexpr uint32_t idesc_o = make_idesc_16(kBM, kC, BF16);
    uint32_t it = 0, kcnt = 0, vcnt = 0;
    uint32_t pphase0 = 0, pphase1 = 0;
    int tcount = 0; (void)tcount;
    const uint32_t tO = tmem + kTmemColsO;
    const uint32_t tQ = tmem + kTmemColsQ;
    const uint32_t sK_addr = smem_u32(sK);
    const uint32_t sV_addr = smem_u32(sV);
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      auto issue_s = [&](int j) {
        const uint32_t s = kcnt % kKStages, ph = (kcnt / kKStages) & 1;
        warp_mbar_wait(k_full + s, ph, lane, 10);
        tc_fence_after();
        const uint32_t tS = tmem + kTmemColsS + (uint32_t)(j & 1) * kBN;
        const uint64_t bd0 = make_sdesc_k_sw128(sK_addr + s * kKBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kC / 16; ++kk) {
            // k-block kk/4 is 8 KB further, the 16-element step inside a 128-byte row is 32 B (>>4 in the descriptor)
            const uint64_t bd = bd0 + (uint64_t)(((kk >> 2) * (kBN * 128) + (kk & 3) * 32) >> 4);
            umma_ts(tS, tQ + kk * 8, bd, idesc_s, kk > 0);
          }
          umma_commit(k_empty + s);
          umma_commit(s_full + (j & 1));
        }
        __syncwarp();
        ++kcnt;
      };
      // q_full(it) also implies that the softmax warps finished draining O of the previous item
      warp_mbar_wait(q_full, it & 1, lane, 11);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        TRACE_MMA(0);
        if (j + 1 < T) issue_s(j + 1);
        TRACE_MMA(1);
        const int b = j & 1;
        if (b == 0) { warp_mbar_wait(p_full + 0, pphase0, lane, 13); pphase0 ^= 1; }
        else        { warp_mbar_wait(p_full + 1, pphase1, lane, 13); pphase1 ^= 1; }
        TRACE_MMA(2);
        const uint32_t s = vcnt % kVStages, ph = (vcnt / kVStages) & 1;
        warp_mbar_wait(v_full + s, ph, lane, 14);
        tc_fence_after();
        TRACE_MMA(3);
        const uint32_t tP = tmem + kTmemColsS + (uint32_t)b * kBN;
        const uint64_t vd0 = make_sdesc_k_sw128(sV_addr + s * kVBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kBN / 16; ++kk) {
            umma_ts(tO, tP + kk * 8, vd0 + (uint64_t)((kk * 32) >> 4), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(v_empty + s);
          umma_commit(o_full);
        }
        __syncwarp();
        TRACE_MMA(4);
        ++tcount;
        ++vcnt;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + drain (1 thread = 1 row)
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t tO = tmem + lane_base + kTmemColsO;
    const uint32_t tQ = tmem + lane_base + kTmemColsQ;
    uint32_t sphase0 = 0, sphase1 = 0;
    uint32_t it = 0;
    int scount = 0; (void)scount;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int qt = item % p.q_tiles;
      const int np = item / p.q_tiles;
      const int pass = np & 1;
      const int n = np >> 1;
      const int row = qt * kBM + warp * 32 + lane;
      const uint32_t pv_base = it * (uint32_t)T;
      // ---- stage this thread's query row into TMEM.  All affinity MMAs of the previous item have completed
      // (s_full of its last tile was observed) and O has been drained, so Q and O may be overwritten.
      {
        const uint4* qsrc = reinterpret_cast<const uint4*>(
            p.t + ((size_t)((1 - pass) * p.N + n) * p.Lp + row) * kC);   // rows < Lp always exist (zero padded)
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t q[32];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const uint4 t4 = __ldg(qsrc + ch * 8 + v);
            q[4 * v + 0] = t4.x; q[4 * v + 1] = t4.y; q[4 * v + 2] = t4.z; q[4 * v + 3] = t4.w;
          }
          tmem_st32(tQ + ch * 32, q);
        }
        tmem_st_wait();
        tc_fence_before();
        warp_mbar_arrive(q_full, lane);
      }
      float m = -INFINITY, l = 0.0f;
      for (int j = 0; j < T; ++j) {
        const int b = j & 1;
        const uint32_t tS = tmem + lane_base + kTmemColsS + (uint32_t)b * kBN;
        if (b == 0) { warp_mbar_wait(s_full + 0, sphase0, lane, 20); sphase0 ^= 1; }
        else        { warp_mbar_wait(s_full + 1, sphase1, lane, 20); sphase1 ^= 1; }
        tc_fence_after();
        TRACE_SM(5);
        ++scount;
        uint32_t s0[32], s1[32];
        tmem_ld32(tS, s0);
        tmem_ld32(tS + 32, s1);
        tmem_ld_wait();
        if (j == T - 1) {
          const int nvalid = p.L - j * kBN;  // >= 1
          if (nvalid < kBN) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              if (k >= nvalid) s0[k] = 0xff800000u;        // -inf
              if (32 + k >= nvalid) s1[k] = 0xff800000u;
            }
          }
        }
        float tmax = __uint_as_float(s0[0]);
#pragma unroll
        for (int k = 1; k < 32; ++k) tmax = fmaxf(tmax, __uint_as_float(s0[k]));
#pragma unroll
        for (int k = 0; k < 32; ++k) tmax = fmaxf(tmax, __uint_as_float(s1[k]));
        if (j == 0) {
          m = tmax;
        } else {
          const bool need = (tmax - m) * kLog2e > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            const float m_new = fmaxf(m, tmax);
            const float scale = fast_exp2((m - m_new) * kLog2e);
            // PV(j-2) is complete (s_full(j) was observed), so the barrier is in phase j-1 or later
            warp_mbar_wait(o_full, (pv_base + (uint32_t)j - 1u) & 1u, lane, 21);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < kC / 32; ++ch) {
              uint32_t o[32];
              tmem_ld32(tO + ch * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * scale);
              tmem_st32(tO + ch * 32, o);
            }
            tmem_st_wait();
            l *= scale;
            m = m_new;
          }
        }
        const float neg_m = -m * kLog2e;
        uint32_t pk[32];
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;   // independent partial sums (ILP)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(s0[2 * k]), kLog2e, neg_m));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s0[2 * k + 1]), kLog2e, neg_m));
          pk[k] = pack16x2<BF16>(p0, p1);
          if constexpr (BF16) { l0 += bf16lo_to_f32(pk[k]); l1 += bf16hi_to_f32(pk[k]); }  // sum what the MMA sees
          else { l0 += p0; l1 += p1; }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(s1[2 * k]), kLog2e, neg_m));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s1[2 * k + 1]), kLog2e, neg_m));
          pk[16 + k] = pack16x2<BF16>(p0, p1);
          if constexpr (BF16) { l2 += bf16lo_to_f32(pk[16 + k]); l3 += bf16hi_to_f32(pk[16 + k]); }
          else { l2 += p0; l3 += p1; }
        }
        l += (l0 + l1) + (l2 + l3);
        tmem_st32(tS, pk);   // P (16-bit pairs) overlays the first 32 columns of the S buffer
        tmem_st_wait();
        tc_fence_before();
        warp_mbar_arrive(p_full + b, lane);
      }
      // drain: Z[c][row] = O[row][c] / l
      // Phases are waited one by one: after s_full(T-1) only PV(T-3) is known complete, so the barrier may
      // still be in phase T-2; a parity wait for phase T-1 alone would alias and pass early.
      if (T >= 2) warp_mbar_wait(o_full, (pv_base + (uint32_t)T - 2u) & 1u, lane, 23);
      warp_mbar_wait(o_full, (pv_base + (uint32_t)T - 1u) & 1u, lane, 22);
      tc_fence_after();
      const float inv = 1.0f / l;
      const bool valid = row < p.L;
      if (p.z != nullptr) {
        float* zcol = p.z + ((size_t)(pass * p.N + n) * kC) * p.L + row;
#pragma unroll 1
        for (int ch = 0; ch < kC / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(tO + ch * 32, o);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int k = 0; k < 32; ++k) zcol[(size_t)(ch * 32 + k) * p.L] = __uint_as_float(o[k]) * inv;
          }
        }
      }
      if (p.cat_a != nullptr) {
        // gate logit of this position: g . Z[:, row] + b  (two passes over the accumulator, it stays in TMEM)
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < kC / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(tO + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            d0 = fmaf(__ldg(p.gate_w + ch * 32 + k + 0), __uint_as_float(o[k + 0]), d0);
            d1 = fmaf(__ldg(p.gate_w + ch * 32 + k + 1), __uint_as_float(o[k + 1]), d1);
            d2 = fmaf(__ldg(p.gate_w + ch * 32 + k + 2), __uint_as_float(o[k + 2]), d2);
            d3 = fmaf(__ldg(p.gate_w + ch * 32 + k + 3), __uint_as_float(o[k + 3]), d3);
          }
        }
        const float logit = ((d0 + d1) + (d2 + d3)) * inv + (p.gate_b ? __ldg(p.gate_b) : 0.f);
        const float gate = 1.0f / (1.0f + __expf(-logit));
        const float sc = inv * gate;
        float* ccol = (pass ? p.cat_b : p.cat_a) + (size_t)n * 2 * kC * p.L + row;
#pragma unroll 1
        for (int ch = 0; ch < kC / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(tO + ch * 32, o);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int k = 0; k < 32; ++k) ccol[(size_t)(ch * 32 + k) * p.L] = __uint_as_float(o[k]) * sc;
          }
        }
        if (valid && p.mask != nullptr) p.mask[(size_t)(pass * p.N + n) * p.L + row] = gate;
      }
      if (valid) p.lse[(size_t)(pass * p.N + n) * p.L + row] = m + __logf(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ==============================================================================================
// passthrough: cat_x[n][C + c][:] = v_x[n][c][:]  (the second half of the concat, :186-187).  Pure HBM copy:
// 8 L C bytes per (sample, side).  Per sample the source is one contiguous run of C*L floats.
// ==============================================================================================
struct PassParams {
  const float* v_a;
  const float* v_b;
  float* cat_a;
  float* cat_b;
  int N;
  size_t plane;   // C * L floats
};

template <int VEC>
__global__ void __launch_bounds__(256) passthrough_kernel(PassParams p) {
  const int side = blockIdx.y / p.N;
  const int n = blockIdx.y - side * p.N;
  const float* src = (side ? p.v_b : p.v_a) + (size_t)n * p.plane;
  float* dst = (side ? p.cat_b : p.cat_a) + (size_t)n * 2 * p.plane + p.plane;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if constexpr (VEC == 4) {
    const size_t n4 = p.plane / 4;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      const float4 a = __ldcs(s4 + i), b = __ldcs(s4 + i + stride), c = __ldcs(s4 + i + 2 * stride),
                   d = __ldcs(s4 + i + 3 * stride);
      __stcs(d4 + i, a); __stcs(d4 + i + stride, b); __stcs(d4 + i + 2 * stride, c); __stcs(d4 + i + 3 * stride, d);
    }
    for (; i < n4; i += stride) __stcs(d4 + i, __ldcs(s4 + i));
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.plane; i += stride) dst[i] = src[i];
  }
}

// ==============================================================================================
// gate: m = sigmoid(g . Z[:,p] + b);  cat[0:C] = Z * m;  cat[C:2C] = V         (:177-187)
// HBM-bound: 16 L C bytes per (sample, side).  One block = kGatePos positions x all 256 channels;
// warp w owns channels [32w, 32w+32), lanes run along positions (coalesced, VEC floats each).
// ==============================================================================================
constexpr int kGateThreads = 512;   // 16 warps x 16 channels; gate_kernel<2, 2, 8>: 32 registers of Z per thread, two blocks per SM

struct GateParams {
  const float* z;      // [2][N][C][L] raw attended features (side 0: Z_a, side 1: Z_b)
  const float* v_a;    // [N][C][L] original features of frame A
  const float* v_b;    // [N][C][L] original features of frame B
  const float* gate_w; // [C]
  const float* gate_b; // [1] or nullptr
  float* cat_a;        // [N][2C][L]
  float* cat_b;        // [N][2C][L]
  int N, L;
};

// vector types of 1, 2 and 4 floats for the streaming loads / stores of the epilogue kernels
template <int VEC> struct GateVec;
template <> struct GateVec<1> { using type = float; };
template <> struct GateVec<2> { using type = float2; };
template <> struct GateVec<4> { using type = float4; };
template <int VEC> __device__ __forceinline__ void gate_ld(const float* src, float (&dst)[VEC]) {
  using V = typename GateVec<VEC>::type;
  const V t = __ldcs(reinterpret_cast<const V*>(src));
  if constexpr (VEC == 4) { dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w; }
  else if constexpr (VEC == 2) { dst[0] = t.x; dst[1] = t.y; }
  else dst[0] = t;
}
template <int VEC> __device__ __forceinline__ void gate_st(float* dst, const float (&src)[VEC]) {
  using V = typename GateVec<VEC>::type;
  V t;
  if constexpr (VEC == 4) { t.x = src[0]; t.y = src[1]; t.z = src[2]; t.w = src[3]; }
  else if constexpr (VEC == 2) { t.x = src[0]; t.y = src[1]; }
  else t = src[0];
  __stcs(reinterpret_cast<V*>(dst), t);
}

// VEC floats per lane (block = 32 * VEC positions x 256 channels); MINB resident blocks per SM: <4, 1> holds a [256 x 128]
// tile in registers with one block per SM, <2, 2> a [256 x 64] tile with two blocks per SM, whose load / reduce / store
// phases overlap each other; PB = passthrough rows in flight per thread.
template <int VEC, int MINB = 1, int PB = 16>
__global__ void __launch_bounds__(kGateThreads, MINB) gate_kernel(GateParams p) {
  constexpr int kWarps = kGateThreads / 32, kCh = kC / kWarps;     // 16 warps, 16 channels each
  __shared__ float part[kWarps][32 * VEC];
  __shared__ float gw[kC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int side = blockIdx.y / p.N;
  const int n = blockIdx.y - side * p.N;
  const int pos = (blockIdx.x * 32 + lane) * VEC;
  if (threadIdx.x < kC) gw[threadIdx.x] = p.gate_w[threadIdx.x];
  __syncthreads();
  const bool valid = pos < p.L;   // L % VEC == 0 is guaranteed by the launcher
  const float* z = p.z + (size_t)blockIdx.y * kC * p.L + pos;
  const float* v = (side ? p.v_b : p.v_a) + (size_t)n * kC * p.L + pos;
  float* cat = (side ? p.cat_b : p.cat_a) + (size_t)n * 2 * kC * p.L + pos;

  float zr[kCh][VEC];
  float dot[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) dot[e] = 0.f;
  if (valid) {
#pragma unroll
    for (int k = 0; k < kCh; ++k) gate_ld<VEC>(z + (size_t)(warp * kCh + k) * p.L, zr[k]);
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
      const float g = gw[warp * kCh + k];
#pragma unroll
      for (int e = 0; e < VEC; ++e) dot[e] = fmaf(g, zr[k][e], dot[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) part[warp][lane * VEC + e] = dot[e];
  // passthrough copy of the original features while the partial sums settle (all loads of a batch first, then its stores)
  if (valid) {
#pragma unroll
    for (int k0 = 0; k0 < kCh; k0 += PB) {
      float t[PB][VEC];
#pragma unroll
      for (int k = 0; k < PB; ++k) gate_ld<VEC>(v + (size_t)(warp * kCh + k0 + k) * p.L, t[k]);
#pragma unroll
      for (int k = 0; k < PB; ++k) gate_st<VEC>(cat + (size_t)(kC + warp * kCh + k0 + k) * p.L, t[k]);
    }
  }
  __syncthreads();
  if (valid) {
    const float bias = p.gate_b ? __ldg(p.gate_b) : 0.f;
    float mask[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float t = bias;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += part[w][lane * VEC + e];
      mask[e] = 1.0f / (1.0f + __expf(-t));
    }
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
      float t[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) t[e] = zr[k][e] * mask[e];
      gate_st<VEC>(cat + (size_t)(warp * kCh + k) * p.L, t);
    }
  }
}

// ==============================================================================================
// merge_gate_kernel: COATTN_FLAG_SPLIT_KEYS (few pairs: fewer work items than CTA pairs).  attend2 swept the key range
// of every item in `splits` parts; part s left Z_s (normalised over ITS keys) and lse_s.  Per position
//   lse = log sum_s exp(lse_s),   Z = sum_s exp(lse_s - lse) Z_s                      (the softmax over all keys)
// followed by the epilogue of gate_kernel (:175-187): gate logit, sigmoid, scale, concat; lse / mask / raw Z are kept
// when asked for.  HBM-bound: reads splits * C L * 4 bytes of parts per sample and side.
// ==============================================================================================
constexpr int kMaxKeySplits = 4;

struct MergeParams {
  const float* zp;     // [splits][passes][N][C][L]
  const float* lsep;   // [splits][passes][N][L]
  const float* v_a;    // [N / q_group][C][L] or null (no passthrough half: gated-only output)
  const float* v_b;    // [N][C][L] or null
  const float* gate_w; // [C]
  const float* gate_b; // [1] or null
  float* cat_a;        // [N][out_channels][L]
  float* cat_b;        // [N][out_channels][L] (unused with passes == 1)
  float* z;            // [passes][N][C][L] or null
  float* lse;          // [passes][N][L] or null
  float* mask;         // [passes][N][L] or null
  int N, L, splits, passes, out_channels, q_group;
};

template <int VEC>
__global__ void __launch_bounds__(kGateThreads) merge_gate_kernel(MergeParams p) {
  constexpr int kWarps = kGateThreads / 32, kCh = kC / kWarps;     // 16 warps, 16 channels each
  __shared__ float part[kWarps][32 * VEC];
  __shared__ float gw[kC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int side = blockIdx.y / p.N;
  const int n = blockIdx.y - side * p.N;
  const int pos = (blockIdx.x * 32 + lane) * VEC;
  if (threadIdx.x < kC) gw[threadIdx.x] = p.gate_w[threadIdx.x];
  __syncthreads();
  const bool valid = pos < p.L;   // L % VEC == 0 is guaranteed by the launcher
  const size_t part_stride = (size_t)p.passes * p.N;               // samples-and-sides per part
  const size_t sn = (size_t)side * p.N + n;
  float* cat = (side ? p.cat_b : p.cat_a) + (size_t)n * p.out_channels * p.L + pos;

  float wgt[kMaxKeySplits][VEC];
  float lse_out[VEC];
  float zr[kCh][VEC];
  float dot[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { dot[e] = 0.f; lse_out[e] = 0.f; }
  if (valid) {
    float ls[kMaxKeySplits][VEC];
#pragma unroll
    for (int s = 0; s < kMaxKeySplits; ++s)
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        ls[s][e] = (s < p.splits) ? __ldg(p.lsep + ((size_t)s * part_stride + sn) * p.L + pos + e) : -INFINITY;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float m = ls[0][e];
#pragma unroll
      for (int s = 1; s < kMaxKeySplits; ++s) m = fmaxf(m, ls[s][e]);
      float sum = 0.f;
#pragma unroll
      for (int s = 0; s < kMaxKeySplits; ++s) { wgt[s][e] = __expf(ls[s][e] - m); sum += wgt[s][e]; }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int s = 0; s < kMaxKeySplits; ++s) wgt[s][e] *= inv;
      lse_out[e] = m + __logf(sum);
    }
#pragma unroll
    for (int k = 0; k < kCh; ++k)
#pragma unroll
      for (int e = 0; e < VEC; ++e) zr[k][e] = 0.f;
    for (int s = 0; s < p.splits; ++s) {
      const float* z = p.zp + (((size_t)s * part_stride + sn) * kC + warp * kCh) * p.L + pos;
      float w[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {      // wgt[s] with a run-time s: select without indexing the register array
        w[e] = wgt[0][e];
#pragma unroll
        for (int t = 1; t < kMaxKeySplits; ++t) w[e] = (s == t) ? wgt[t][e] : w[e];
      }
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        if constexpr (VEC == 4) {
          const float4 t = __ldcs(reinterpret_cast<const float4*>(z + (size_t)k * p.L));
          zr[k][0] = fmaf(w[0], t.x, zr[k][0]); zr[k][1] = fmaf(w[1], t.y, zr[k][1]);
          zr[k][2] = fmaf(w[2], t.z, zr[k][2]); zr[k][3] = fmaf(w[3], t.w, zr[k][3]);
        } else {
          zr[k][0] = fmaf(w[0], __ldcs(z + (size_t)k * p.L), zr[k][0]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
      const float g = gw[warp * kCh + k];
#pragma unroll
      for (int e = 0; e < VEC; ++e) dot[e] = fmaf(g, zr[k][e], dot[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) part[warp][lane * VEC + e] = dot[e];
  // passthrough half of the concat while the partial sums settle
  const float* vsrc = side ? p.v_b : p.v_a;
  if (valid && vsrc != nullptr) {
    const float* v = vsrc + (size_t)(side ? n : n / p.q_group) * kC * p.L + pos;
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
      const int c = warp * kCh + k;
      if constexpr (VEC == 4) __stcs(reinterpret_cast<float4*>(cat + (size_t)(kC + c) * p.L), __ldcs(reinterpret_cast<const float4*>(v + (size_t)c * p.L)));
      else __stcs(cat + (size_t)(kC + c) * p.L, __ldcs(v + (size_t)c * p.L));
    }
  }
  __syncthreads();
  if (valid) {
    const float bias = p.gate_b ? __ldg(p.gate_b) : 0.f;
    float mask[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float t = bias;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += part[w][lane * VEC + e];
      mask[e] = 1.0f / (1.0f + __expf(-t));
    }
    float* zout = p.z ? p.z + (sn * kC) * p.L + pos : nullptr;
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
      const int c = warp * kCh + k;
      if constexpr (VEC == 4) {
        float4 t;
        t.x = zr[k][0] * mask[0]; t.y = zr[k][1] * mask[1]; t.z = zr[k][2] * mask[2]; t.w = zr[k][3] * mask[3];
        __stcs(reinterpret_cast<float4*>(cat + (size_t)c * p.L), t);
        if (zout) *reinterpret_cast<float4*>(zout + (size_t)c * p.L) = make_float4(zr[k][0], zr[k][1], zr[k][2], zr[k][3]);
      } else {
        __stcs(cat + (size_t)c * p.L, zr[k][0] * mask[0]);
        if (zout) zout[(size_t)c * p.L] = zr[k][0];
      }
    }
    if (warp == 0) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if (p.lse) p.lse[sn * p.L + pos + e] = lse_out[e];
        if (p.mask) p.mask[sn * p.L + pos + e] = mask[e];
      }
    }
  }
}

}  // namespace coattn
