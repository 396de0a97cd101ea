// bwd_flash: the L x L part of the co-attention backward without ever writing an L x L matrix to memory.
//
// What autograd does through rgbd_segmentation_RAA.py:160-170 (train.py:599), per sample, with
//   P_a[i,j] = exp(S[i,j] - lse_a[i])  (softmax over j, :165)      P_b[i,j] = exp(S[i,j] - lse_b[j])  (softmax over i, :164)
//   dP_a = dZ_a^T B      dP_b = A^T dZ_b      dS = P_a (dP_a - delta_a[i]) + P_b (dP_b - delta_b[j])
//   dQ[:,i]   = sum_j dS[i,j] B[:,j]                      (-> dW = dQ A^T, dA += W^T dQ by two small GEMMs afterwards)
//   dA[:,i]  += sum_j P_b[i,j] dZ_b[:,j]
//   dB[:,j]  += sum_i dS[i,j] Q[:,i] + sum_i P_a[i,j] dZ_a[:,i]       (only with no_grad_for_counterpart=False, :147-148)
// is a set of flash-attention-shaped sweeps: a CTA pair owns 256 "row" positions, keeps one fp32 accumulator O [256 x C]
// in TMEM and sweeps the "column" positions in tiles of 128; per tile it RECOMPUTES the affinity tile S (forward operand
// format, so that exp(S - lse) is the forward's softmax), optionally a second product T (dP_a or dP_b, bf16 operands),
// forms X = exp(S - n) or X = exp(S - n) (T - d) in registers (n, d: per-row or per-column vectors), stores X as 16-bit
// pairs in TMEM and issues O += X V with X as the TMEM operand -- the structure of attend2 (attend2_kernel.cuh), minus the
// online softmax (the normalisers are known) plus the second product.
//
// A work item = (kind, sample, 256-row tile) runs 1-3 PHASES that accumulate into the same O:
// (all gradient operands -- dZ_a, dZ_b and X -- are in the forward's operand format; with fp16 they carry the call-wide
//  power-of-two scale s of bwd_planes_kernel, delta is scaled on load and O is unscaled in the drain)
//   kind DQ   : phase A  S = Q_I^T B_J,  T = dZ_a,I^T B_J,  X = P_a (T - delta_a[I]),  V = B      (row vectors)
//               phase B  S = Q_I^T B_J,  T = A_I^T dZ_b,J,  X = P_b (T - delta_b[J]),  V = B      (column vectors)
//               -> dQ as bf16 in both layouts ([Lp][C] and [C][Lp]) for the two small GEMMs that follow
//   kind DAPB : phase    S = Q_I^T B_J,                     X = P_b,                   V = dZ_b   -> d_v_a += O
//   kind DB   : the same three with the roles of the frames swapped (rows = B positions), all into d_v_b += O
// The depth modality (B branch gradient dead, :240-247) runs phase A only.
//
// Shared memory (per CTA): two RESIDENT row-side operand tiles R1 (S) and R2 (T), 64 KB each; a 2-slot ring of 32 KB
// column-side operand halves (C1 and C2 of a tile alternate); one 32 KB V slot (the T-less items that run last in a
// cluster's list borrow the idle R2 for two more).  TMEM (per CTA): [0,256) O | [256,384) S | [384,512) T.
// A phase with T: X(j) is written over the S BUFFER -- a warp turns T(j) into X(j) in registers (t_free: the T buffer goes
// back), pulls S(j+1) into registers and stores X(j) over the first half of the S columns it has just read; a phase without
// T uses [384,448) and [448,512) as two alternating X buffers.
// Tensor-pipe order inside a phase with T:  S(0) T(0) S(1) | T(1) PV(0) S(2) | T(2) PV(1) S(3) | ...
//   T(j+1) only needs T(j) in registers and runs while X(j) is produced; S(j+2) follows PV(j) -- the reader of X(j) -- IN THE
//   PIPE (the MMAs of one issuing thread execute in order), which is what makes overwriting X(j) safe without a barrier.
//   History (clock64 traces under profiles/): ONE buffer shared by S and T, 6.0 k cycles per tile for 3.1 k of MMA; separate
//   buffers with X written over T (T(j+1) behind PV(j): the serial chain T -> X -> PV -> T), 4.5 k; this layout, 3.3 k.
#pragma once
#include "attend2_kernel.cuh"
#include "backward_kernels.cuh"

namespace coattn {

constexpr int kFRBytes = k2BM * kC * 2;          // 64 KB : row-side operand tile, 2 chunks of [256 ch x 64 positions]
constexpr int kFKBytes = (k2BN / 2) * kC * 2;    // 32 KB : this CTA's half of a column-side operand tile
constexpr int kFVBytes = (kC / 2) * k2BN * 2;    // 32 KB : this CTA's channels of a V tile
constexpr int kFKSlots = 2;
// G = X-producer warps per TMEM lane quadrant (they split the 128 columns of a tile and the 256 channels of the drain):
//   warps [0, 4G) X producers + drain | 4G TMA (R tiles, column slot 0) | 4G+1 MMA issuer | 4G+2 TMA (V) | 4G+3 TMA (column slot 1)
template <int G> struct FlashCfg {
  static constexpr int kXWarps = 4 * G;
  static constexpr int kThreads = 32 * (4 * G + 4);
  static constexpr int kCols = k2BN / G;      // tile columns per thread
  static constexpr int kLd = kCols / 32;      // 32-column TMEM loads per thread and product
  static constexpr int kCh = kC / G;          // channels per thread in the drain
};
constexpr int kFColvBytes = 2 * 256 * 4;         // [tile parity][normaliser x 128 | delta x 128]
constexpr int kFSmemBytes = 2 * kFRBytes + kFKSlots * kFKBytes + kFVBytes + kFColvBytes + 256;
static_assert(kFSmemBytes <= 232448, "bwd_flash shared memory exceeds the 227 KB per-CTA limit");
constexpr uint32_t kFTmemO = 0, kFTmemS = 256, kFTmemT = 384, kFTmemX = 384;
constexpr int kFMaxPhases = 6;
constexpr int kFMaps = 10;

// Debug builds (-DCOATTN_TRACE_FLASH): clock64 stamps of cluster 0's second item, per column tile:
//   MMA issuer  0 loop top | 1 T(j+1) issued | 2 v_full seen | 3 x_full seen (PV issue) | 4 S(j+2) issued
//   X warp 0    9 S(j+1) in registers | 11 T(j) seen | 12 X stored + arrived        (phase with T; other slots: older layouts)
#ifdef COATTN_TRACE_FLASH
__device__ long long g_flash_trace[64 * 16];
#define FTR(j, slot) do { if (blockIdx.x == 0 && lane == 0 && idx == 1 && (j) < 64) g_flash_trace[(j) * 16 + (slot)] = clock64(); } while (0)
// per item of cluster 0 (X warp 0): 0 item start | 1 + k end of phase k's sweep (last X stored) | 5 O complete | 6 drain done | 7 kind
__device__ long long g_flash_items[32 * 8];
#define FTI(slot, v) do { if (blockIdx.x == 0 && warp == 0 && lane == 0 && idx < 32) g_flash_items[idx * 8 + (slot)] = (v); } while (0)
#else
#define FTR(j, slot) do {} while (0)
#define FTI(slot, v) do {} while (0)
#endif

struct FlashMaps {
  CUtensorMap m[kFMaps];     // 16-bit planes [N * C][Lp]; box {64 positions, 256 channels} or {64 positions, 128 channels}
};

struct FlashPhase {
  int r1, c1;          // tensor maps (box 256 rows) of S = R1_I^T C1_J: row-side / column-side operand
  int r2, c2;          // ... of T = R2_I^T C2_J; r2 < 0: no second product (X = P)
  int v;               // tensor map (box 128 rows) of the PV operand
  int vec_col;         // 0: nvec / dvec are indexed by the item's row position, 1: by the swept column position
  const float* nvec;   // [N][L] log-sum-exp normaliser n of P = exp(S - n)
  const float* dvec;   // [N][L] delta of X = P (T - d); unused when r2 < 0
};

struct FlashKind {
  int phase0, phases;      // phases [phase0, phase0 + phases) of FlashParams::ph
  int out_mode;            // 0: O -> 16-bit, both layouts (dQ);  1: acc[n][c][row] += O (fp32, [N][C][L])
  unsigned short* out_t;   // mode 0: [N][Lp][C]
  unsigned short* out_c;   // mode 0: [N][C][Lp]
  float* acc;              // mode 1
};

struct FlashParams {
  FlashPhase ph[kFMaxPhases];
  FlashKind kind[2];
  int items[2];            // work items of kind 0 (the long ones) and kind 1: N * q_pairs each, or 0
  int ratio;               // cost of a kind-0 item in kind-1 items (rounded), for the static balance below
  int N, L, Lp;
  int q_pairs;             // ceil(L / 256)
  int kv_tiles;            // ceil(L / 128)
  uint32_t idesc_s, idesc_s_last;     // S: forward operand format, M 256 x N 128 (n_last in the ragged tile), MN-major operands
  uint32_t idesc_t, idesc_t_last;     // T: gradient operand format (= forward format; the gradient planes are scaled, see bwd_planes_kernel)
  uint32_t idesc_o;                   // PV: same format, M 256 x N 256, K-major
  const unsigned* absmax;             // bits of max |dZ| of the call: the gradient planes hold s dZ, s = grad_scale_from_absmax
                                      // (fp16 only; null: s = 1).  delta is multiplied by s on load, O by 1 / s in the drain.
};

// Static schedule.  Cluster k of K runs its kind-0 items (k, k + K, ...) first and then a contiguous range of kind-1 items.
// The clusters that got one kind-0 item fewer receive `ratio` kind-1 items more before the rest is dealt round; every warp
// of the pair evaluates the same closed form, so no work counter or broadcast is needed.
struct FlashSched {
  int long_cnt, short_start, short_cnt;
};
__device__ __forceinline__ FlashSched flash_schedule(int k, int K, int n0, int n1, int ratio) {
  FlashSched s;
  const int k_hi = (n0 % K == 0) ? K : (n0 % K);      // clusters holding ceil(n0 / K) long items: k < k_hi
  const int k_lo = K - k_hi;
  int extra_total = k_lo * ratio;
  if (extra_total > n1) extra_total = n1;
  const int rest = n1 - extra_total;
  auto cnt = [&](int c) {
    int x = rest / K + (c < rest % K ? 1 : 0);
    if (c >= k_hi && k_lo > 0) x += extra_total / k_lo + ((c - k_hi) < extra_total % k_lo ? 1 : 0);
    return x;
  };
  s.long_cnt = n0 / K + (k < n0 % K ? 1 : 0);
  s.short_cnt = cnt(k);
  int st = 0;
  for (int c = 0; c < k; ++c) st += cnt(c);
  s.short_start = st;
  return s;
}

// XBF: X (and the gradient planes) are bf16 (bf16 forward) instead of scaled fp16
template <int G, bool XBF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FlashCfg<G>::kThreads, 1)
bwd_flash_kernel(const __grid_constant__ FlashMaps maps, const __grid_constant__ FlashParams p) {
  pdl_wait();
  using Cfg = FlashCfg<G>;
  constexpr int kXWarps = Cfg::kXWarps, kCols = Cfg::kCols, kLd = Cfg::kLd, kCh = Cfg::kCh;
  constexpr int kFProducerWarp = kXWarps, kFMmaWarp = kXWarps + 1, kFVProducerWarp = kXWarps + 2, kFC2ProducerWarp = kXWarps + 3;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sR1 = smem;
  uint8_t* sR2 = sR1 + kFRBytes;
  uint8_t* sK = sR2 + kFRBytes;
  uint8_t* sV = sK + kFKSlots * kFKBytes;
  float* colv = reinterpret_cast<float*>(sV + kFVBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(colv) + kFColvBytes);
  uint64_t* r1_full = bars + 0;     // (L) row tiles of both CTAs landed
  uint64_t* r1_empty = bars + 1;    // every S MMA of the item completed
  uint64_t* r2_full = bars + 2;     // (L)
  uint64_t* r2_empty = bars + 3;    // every T MMA of the phase completed
  uint64_t* k_full = bars + 4;      // (L) [2]
  uint64_t* k_empty = bars + 6;     // [2]
  // V slots: slot 0 = sV; slots 1 and 2 = the two halves of R2, borrowed by the T-less items that run last (R2 holds nothing
  // then): with one slot PV(j+1) can only start a TMA round trip after PV(j) has completed (~2.4 k cycles per tile + hand-offs)
  uint64_t* v_full = bars + 19;     // (L) [3]
  uint64_t* v_empty = bars + 22;    // [3]
  uint64_t* s_full = bars + 10;     // S complete (both CTAs)
  uint64_t* s_free = bars + 11;     // (L) S sits in the registers of every softmax warp of the pair
  uint64_t* x_full = bars + 12;     // (L) [2] one per X buffer
  uint64_t* o_full = bars + 14;     // [2] the PV that read X buffer b has completed
  uint64_t* t_full = bars + 16;     // T complete (both CTAs)
  uint64_t* t_free = bars + 18;     // (L) T sits in the registers of every X-producer warp of the pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == kFProducerWarp && lane == 0) {
    for (int i = 0; i < kFMaps; ++i) tma_prefetch_desc(&maps.m[i]);
    mbar_init(r1_full, 1); mbar_init(r1_empty, 1); mbar_init(r2_full, 1); mbar_init(r2_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
    for (int v = 0; v < 3; ++v) { mbar_init(v_full + v, 1); mbar_init(v_empty + v, 1); }
    mbar_init(s_full, 1);
    mbar_init(t_full, 1);
    mbar_init(s_free, 2 * kXWarps);
    mbar_init(t_free, 2 * kXWarps);
    for (int b = 0; b < 2; ++b) { mbar_init(x_full + b, 2 * kXWarps); mbar_init(o_full + b, 1); }
    fence_mbar_init();
  }
  if (warp == kFMmaWarp) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int T = p.kv_tiles;
  const int n_last = ((p.L - (T - 1) * k2BN) + 15) & ~15;     // columns the MMAs touch in the ragged last tile
  const FlashSched sch = flash_schedule(cluster_id, num_clusters, p.items[0], p.items[1], p.ratio);
  const int my_items = sch.long_cnt + sch.short_cnt;
  // item idx of this cluster -> (kind, item id)
  auto decode = [&](int idx, int& kind, int& item) {
    if (idx < sch.long_cnt) { kind = 0; item = cluster_id + idx * num_clusters; }
    else { kind = 1; item = sch.short_start + (idx - sch.long_cnt); }
  };

  if (warp == kFProducerWarp) {
    // ------------------------------------------------------------------ TMA producer: resident row tiles + column tiles
    if (lane == 0) {
      const uint32_t r1_full_l = mapa_u32(smem_u32(r1_full), 0);
      const uint32_t r2_full_l = mapa_u32(smem_u32(r2_full), 0);
      uint32_t it = 0, r2cnt = 0;
      uint32_t kuse0 = 0;              // loads that went through column slot 0 so far (phase of k_empty[0] / k_full[0])
      for (int idx = 0; idx < my_items; ++idx, ++it) {
        int kd, item;
        decode(idx, kd, item);
        const FlashKind& K = p.kind[kd];
        const int qp = item % p.q_pairs, n = item / p.q_pairs;
        const int rpos0 = qp * (2 * k2BM) + (int)rank * k2BM;
        mbar_wait(r1_empty, (it & 1) ^ 1, 1);
        if (rank == 0) mbar_arrive_expect_tx(r1_full, 2 * kFRBytes);
#pragma unroll
        for (int mc = 0; mc < 2; ++mc)
          tma_load_2d_pair(sR1 + mc * 32768, &maps.m[p.ph[K.phase0].r1], r1_full_l, rpos0 + mc * 64, n * kC);
        for (int pi = K.phase0; pi < K.phase0 + K.phases; ++pi) {
          const FlashPhase& ph = p.ph[pi];
          const bool has_t = ph.r2 >= 0;
          if (has_t) {
            mbar_wait(r2_empty, (r2cnt & 1) ^ 1, 2);
            if (rank == 0) mbar_arrive_expect_tx(r2_full, 2 * kFRBytes);
#pragma unroll
            for (int mc = 0; mc < 2; ++mc)
              tma_load_2d_pair(sR2 + mc * 32768, &maps.m[ph.r2], r2_full_l, rpos0 + mc * 64, n * kC);
            ++r2cnt;
          }
          // Column-side tiles.  Slot 0 belongs to this warp, slot 1 to the second producer warp (each waits on EVERY release
          // of its own slot, so a parity wait can never be two phases behind).  With T: C1(j) -> slot 0, C2(j) -> slot 1 --
          // S runs about a tile ahead of T and one in-order producer would hold C1(j+2) back behind C2(j+1).  Without T:
          // C1(j) alternates, even tiles here, odd tiles in the other warp.
          // C1 == C2 (phase A with one operand format: S and dP_a both multiply B_J): ONE load per tile serves both products
          // and the two slots alternate like in a phase without T -- the column tiles are then double buffered.
          const bool alt = !has_t || ph.c2 == ph.c1;
          for (int j = 0; j < T; ++j) {
            if (alt && (j & 1)) continue;
            // this CTA's half of the column positions: 64 (n_last / 2 in the ragged last tile) x 256 channel rows
            const int kpos = j * k2BN + (int)rank * ((j == T - 1) ? (n_last / 2) : (k2BN / 2));
            mbar_wait(k_empty + 0, (kuse0 & 1) ^ 1, 3);
            if (rank == 0) mbar_arrive_expect_tx(k_full + 0, 2 * kFKBytes);
            tma_load_2d_pair(sK, &maps.m[ph.c1], mapa_u32(smem_u32(k_full + 0), 0), kpos, n * kC);
            ++kuse0;
          }
        }
      }
    }
  } else if (warp == kFC2ProducerWarp) {
    // ------------------------------------------------------------------ TMA producer: column slot 1 (C2, or odd C1 tiles)
    if (lane == 0) {
      uint32_t kuse1 = 0;
      for (int idx = 0; idx < my_items; ++idx) {
        int kd, item;
        decode(idx, kd, item);
        const FlashKind& K = p.kind[kd];
        const int n = item / p.q_pairs;
        for (int pi = K.phase0; pi < K.phase0 + K.phases; ++pi) {
          const FlashPhase& ph = p.ph[pi];
          const bool has_t = ph.r2 >= 0;
          const bool alt = !has_t || ph.c2 == ph.c1;
          for (int j = 0; j < T; ++j) {
            if (alt && !(j & 1)) continue;
            const int kpos = j * k2BN + (int)rank * ((j == T - 1) ? (n_last / 2) : (k2BN / 2));
            mbar_wait(k_empty + 1, (kuse1 & 1) ^ 1, 5);
            if (rank == 0) mbar_arrive_expect_tx(k_full + 1, 2 * kFKBytes);
            tma_load_2d_pair(sK + kFKBytes, &maps.m[alt ? ph.c1 : ph.c2], mapa_u32(smem_u32(k_full + 1), 0), kpos, n * kC);
            ++kuse1;
          }
        }
      }
    }
  } else if (warp == kFVProducerWarp) {
    // ------------------------------------------------------------------ TMA producer: V tiles
    if (lane == 0) {
      uint32_t vu0 = 0, vu1 = 0, vu2 = 0;      // loads that went through each V slot so far
      uint32_t ntv = 0, r2done = 0;      // V tiles of T-less phases so far; phases with T so far (completions of r2_empty)
      for (int idx = 0; idx < my_items; ++idx) {
        int kd, item;
        decode(idx, kd, item);
        const FlashKind& K = p.kind[kd];
        const int n = item / p.q_pairs;
        for (int pi = K.phase0; pi < K.phase0 + K.phases; ++pi) {
          const CUtensorMap* mv = &maps.m[p.ph[pi].v];
          const bool has_t = p.ph[pi].r2 >= 0;
          // Items of kind 1 that consist of ONE T-less phase borrow R2 for two more V slots.  They run after all kind-0 items
          // of the cluster (flash_schedule), so no phase with T -- no load of R2 -- can follow them: the R producer needs no
          // guard (it runs phases ahead of the PVs, and a parity wait cannot look back more than one completion).  Here:
          // every T MMA of the last phase that read R2 must have completed (this warp is at most one tile ahead of the PVs).
          const bool borrow = !has_t && K.phases == 1 && kd == 1;
          if (borrow && r2done > 0) mbar_wait(r2_empty, (r2done - 1) & 1, 7);
          for (int j = 0; j < T; ++j) {
            const uint32_t vs = borrow ? (ntv++ % 3u) : 0u;
            uint8_t* slot = vs == 0 ? sV : sR2 + (vs - 1) * kFVBytes;
            const uint32_t vu = vs == 0 ? vu0 : (vs == 1 ? vu1 : vu2);
            mbar_wait(v_empty + vs, (vu & 1) ^ 1, 4);
            if (vs == 0) ++vu0; else if (vs == 1) ++vu1; else ++vu2;
            if (rank == 0) mbar_arrive_expect_tx(v_full + vs, 2 * kFVBytes);
            const uint32_t v_full_l = mapa_u32(smem_u32(v_full + vs), 0);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
              tma_load_2d_pair(slot + kb * ((kC / 2) * 128), mv, v_full_l, j * k2BN + kb * 64, n * kC + (int)rank * (kC / 2));
          }
          if (has_t) ++r2done;
        }
      }
    }
  } else if (warp == kFMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; uniform control flow)
    if (rank == 0) {
      const int ksteps_last = n_last / 16;
      uint32_t it = 0, scnt = 0, r2cnt = 0, tcnt = 0;
      uint32_t vu0 = 0, vu1 = 0, vu2 = 0, ntv = 0;
      uint32_t kuse0 = 0, kuse1 = 0;
      uint32_t xuse0 = 0, xuse1 = 0;   // uses of X buffer 0 / 1 so far (phase of x_full[b] / o_full[b])
      uint32_t nt_tile = 0;            // tiles of phases without T so far (they alternate between the two X buffers)
      const uint32_t tO = tmem + kFTmemO, tS = tmem + kFTmemS, tT = tmem + kFTmemT;
      const uint64_t r1d0 = make_sdesc_mn_sw128(smem_u32(sR1), 32768, 1024);
      const uint64_t r2d0 = make_sdesc_mn_sw128(smem_u32(sR2), 32768, 1024);
      const uint32_t sK_addr = smem_u32(sK), sV_addr = smem_u32(sV), sR2_addr = smem_u32(sR2);
      for (int idx = 0; idx < my_items; ++idx, ++it) {
        int kd, item;
        decode(idx, kd, item);
        const FlashKind& K = p.kind[kd];
        bool pi0_trace = true;
        // one affinity-type product of column tile j: S into the S buffer (needs the previous S in registers), or T into
        // the T buffer (follows the previous PV in the pipe, which read X out of that buffer; T itself was in registers
        // before that X existed)
        bool shared_c = false;      // this phase's S and T read the same column tile
        auto issue_set = [&](bool is_t, int j, bool phase_has_t) {
          const bool alt = !phase_has_t || shared_c;
          const uint32_t s = alt ? (uint32_t)(j & 1) : (is_t ? 1u : 0u);
          // shared tile: S waits for it, T reuses it (its use count was taken by S) and releases it
          const bool reuse = shared_c && is_t;
          const uint32_t phs = ((s ? kuse1 : kuse0) - (reuse ? 1u : 0u)) & 1;
          if (!reuse) warp_mbar_wait(k_full + s, phs, lane, 10);
          if (pi0_trace && !is_t) FTR(j - 1, 13);
          if (!is_t && scnt > 0) warp_mbar_wait(s_free, (scnt - 1) & 1, lane, 12);
          if (is_t && tcnt > 0) warp_mbar_wait(t_free, (tcnt - 1) & 1, lane, 16);
          if (pi0_trace && !is_t) FTR(j - 1, 14);
          tc_fence_after();
          const uint64_t kd0 = make_sdesc_mn_sw128(sK_addr + s * kFKBytes, 32768, 1024);
          const uint64_t rd0 = is_t ? r2d0 : r1d0;
          const uint32_t idesc = is_t ? ((j == T - 1) ? p.idesc_t_last : p.idesc_t) : ((j == T - 1) ? p.idesc_s_last : p.idesc_s);
          const uint32_t td = is_t ? tT : tS;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kC / 16; ++kk)      // MN-major operands: 16 channel rows = 2048 B per K step
              umma2_ss(td, rd0 + (uint64_t)((kk * 2048) >> 4), kd0 + (uint64_t)((kk * 2048) >> 4), idesc, kk > 0);
            if (!(shared_c && !is_t)) umma2_commit_mc(k_empty + s, 3);
            umma2_commit_mc(is_t ? t_full : s_full, 3);
          }
          __syncwarp();
          if (!reuse) { if (s) ++kuse1; else ++kuse0; }
          if (!is_t) ++scnt; else ++tcnt;
        };
        warp_mbar_wait(r1_full, it & 1, lane, 11);
        tc_fence_after();
        bool first_pv = true;
        for (int pi = K.phase0; pi < K.phase0 + K.phases; ++pi) {
          const bool has_t = p.ph[pi].r2 >= 0;
          if (has_t) { warp_mbar_wait(r2_full, r2cnt & 1, lane, 15); tc_fence_after(); }
          pi0_trace = (pi == K.phase0);
          shared_c = has_t && p.ph[pi].c2 == p.ph[pi].c1;
          // Order of the products.  Without T:  S(0) | S(1) PV(0) | S(2) PV(1) | ...   (X alternates between two buffers)
          // With T:  S(0) T(0) S(1) | T(1) PV(0) S(2) | T(2) PV(1) S(3) | ...   X(j) lives in the S buffer (written over
          // S(j+1), which every warp pulls into registers first), so T(j+1) only waits for T(j) to sit in registers and runs
          // while X(j) is produced, and S(j+2) follows PV(j) -- the reader of X(j) -- in the pipe.  (X used to be written
          // over T: every tile then paid the serial chain T(j) -> X(j) -> PV(j) -> T(j+1), 4.5 k cycles for 3.07 k of MMA.)
          issue_set(false, 0, has_t);
          if (has_t) { issue_set(true, 0, has_t); if (T > 1) issue_set(false, 1, has_t); }
          for (int j = 0; j < T; ++j) {
            if (pi == K.phase0) FTR(j, 0);
            if (has_t) { if (j + 1 < T) issue_set(true, j + 1, has_t); }
            else if (j + 1 < T) issue_set(false, j + 1, has_t);
            if (pi == K.phase0) FTR(j, 1);
            // O += X(j) V(j)
            const uint32_t xb = has_t ? 0u : (nt_tile++ & 1u);
            const uint32_t vs = (!has_t && K.phases == 1 && kd == 1) ? (ntv++ % 3u) : 0u;      // see the V producer
            const uint32_t vu = vs == 0 ? vu0 : (vs == 1 ? vu1 : vu2);
            warp_mbar_wait(v_full + vs, vu & 1, lane, 14);
            if (vs == 0) ++vu0; else if (vs == 1) ++vu1; else ++vu2;
            if (pi == K.phase0) FTR(j, 2);
            warp_mbar_wait(x_full + xb, (xb ? xuse1 : xuse0) & 1, lane, 13);
            if (pi == K.phase0) FTR(j, 3);
            tc_fence_after();
            const uint32_t tX = tmem + kFTmemX + xb * (k2BN / 2);
            const uint64_t vd0 = make_sdesc_k_sw128(vs == 0 ? sV_addr : sR2_addr + (vs - 1) * kFVBytes);
            const int ksteps = (j == T - 1) ? ksteps_last : k2BN / 16;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < k2BN / 16; ++kk) {
                if (kk < ksteps) {
                  const uint64_t bd = vd0 + (uint64_t)(((kk >> 2) * ((kC / 2) * 128) + (kk & 3) * 32) >> 4);
                  // with T, group g's X sits at the start of ITS OWN S columns (16 tile columns per K step, 8 TMEM columns)
                  const uint32_t xa = has_t ? tmem + kFTmemS + (uint32_t)((kk / (kCols / 16)) * kCols + (kk % (kCols / 16)) * 8)
                                            : tX + kk * 8;
                  umma2_ts(tO, xa, bd, p.idesc_o, (!first_pv || kk > 0) ? 1u : 0u);
                }
              }
              umma2_commit_mc(v_empty + vs, 3);
              umma2_commit_mc(o_full + xb, 3);
            }
            __syncwarp();
            first_pv = false;
            if (xb) ++xuse1; else ++xuse0;
            if (has_t && j + 2 < T) issue_set(false, j + 2, has_t);
            if (pi == K.phase0) FTR(j, 4);
          }
          if (has_t) {
            if (elect_one()) umma2_commit_mc(r2_empty, 3);      // every T of the phase has completed: R2 may be replaced
            __syncwarp();
            ++r2cnt;
          }
        }
        if (elect_one()) umma2_commit_mc(r1_empty, 3);
        __syncwarp();
      }
    }
  } else if (warp < kXWarps) {
    // ------------------------------------------------------------------ X producers + drain
    const int g = warp >> 2;             // column group: columns [kCols g, kCols (g + 1)) of a tile, channels [kCh g, kCh (g + 1)) of O
    const int quad = warp & 3;
    const int rloc = quad * 32 + lane;
    const int et = threadIdx.x;          // 0 .. 128 G - 1
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t tSg = tmem + lane_base + kFTmemS + (uint32_t)(g * kCols);
    const uint32_t tTg = tmem + lane_base + kFTmemT + (uint32_t)(g * kCols);
    const uint32_t tOg = tmem + lane_base + kFTmemO + (uint32_t)(g * kCh);
    const uint32_t s_free_l = mapa_u32(smem_u32(s_free), 0);
    const uint32_t t_free_l = mapa_u32(smem_u32(t_free), 0);
    const uint32_t x_full_l0 = mapa_u32(smem_u32(x_full + 0), 0);
    const uint32_t x_full_l1 = mapa_u32(smem_u32(x_full + 1), 0);
    uint32_t scnt = 0, tcnt = 0, cvcnt = 0, nt_tile = 0;
    uint32_t xuse0 = 0, xuse1 = 0;
    uint32_t last_xb = 0;
    const float gs = (!XBF && p.absmax != nullptr) ? grad_scale_from_absmax(__ldg(p.absmax)) : 1.0f;
    const float inv_gs = 1.0f / gs;      // a power of two: exact
    for (int idx = 0; idx < my_items; ++idx) {
      int kd, item;
      decode(idx, kd, item);
      const FlashKind& K = p.kind[kd];
      const int qp = item % p.q_pairs, n = item / p.q_pairs;
      const int row = qp * (2 * k2BM) + (int)rank * k2BM + rloc;
      const bool vrow = row < p.L;
      FTI(1, 0ll); FTI(2, 0ll); FTI(3, 0ll); FTI(0, clock64()); FTI(7, (long long)kd);
      for (int pi = K.phase0; pi < K.phase0 + K.phases; ++pi) {
        const FlashPhase& ph = p.ph[pi];
        const bool has_t = ph.r2 >= 0;
        const bool vcol = ph.vec_col != 0;
        const float* nv = ph.nvec + (size_t)n * p.L;
        const float* dv = has_t ? ph.dvec + (size_t)n * p.L : nullptr;
        // row vectors live in registers for the whole phase
        const float nrow = (!vcol && vrow) ? -__ldg(nv + row) * kLog2e : 0.f;
        const float drow = (!vcol && vrow && has_t) ? __ldg(dv + row) * gs : 0.f;
        // column vectors: thread et stages element et & 127 of the normaliser (et < 128) or of delta; fetched one tile ahead
        auto fetch_col = [&](int j) -> float {
          const int pos = j * k2BN + (et & 127);
          if (!vcol || pos >= p.L) return 0.f;
          if (et < 128) return -__ldg(nv + pos) * kLog2e;
          return has_t ? __ldg(dv + pos) * gs : 0.f;
        };
        float cnext = fetch_col(0);
        // publish the column vectors of tile j (vcol phases) in the buffer of the current parity, prefetch those of tile j + 1.
        // The barrier also keeps a fast warp from overwriting the buffer of the tile before last.
        auto stage_cols = [&](int j) -> float* {
          float* cvj = colv + (cvcnt & 1) * 256;
          if (vcol) {
            if (et < 256) cvj[et] = cnext;
            named_bar_sync(1, 128 * G);
            ++cvcnt;
            if (j + 1 < T && et < 256) cnext = fetch_col(j + 1);
          }
          return cvj;
        };
        // the next S tile -> registers; the buffer goes back to the issuer at once
        auto load_s = [&](uint32_t (&sv)[kLd][32]) {
          warp_mbar_wait(s_full, scnt & 1, lane, 20);
          ++scnt;
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < kLd; ++c) tmem_ld32(tSg + c * 32, sv[c]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(s_free_l);
        };
        // P = exp(S - n) of tile j for this thread's columns (padding rows / ragged columns -> 0)
        auto exp_s = [&](const uint32_t (&sv)[kLd][32], const float* cvj, int j, float (&pr)[kCols]) {
#pragma unroll
          for (int c = 0; c < kLd; ++c)
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              float4 nc = make_float4(nrow, nrow, nrow, nrow);
              if (vcol) nc = *reinterpret_cast<const float4*>(cvj + g * kCols + c * 32 + k);
              pr[c * 32 + k + 0] = fast_exp2(fmaf(__uint_as_float(sv[c][k + 0]), kLog2e, nc.x));
              pr[c * 32 + k + 1] = fast_exp2(fmaf(__uint_as_float(sv[c][k + 1]), kLog2e, nc.y));
              pr[c * 32 + k + 2] = fast_exp2(fmaf(__uint_as_float(sv[c][k + 2]), kLog2e, nc.z));
              pr[c * 32 + k + 3] = fast_exp2(fmaf(__uint_as_float(sv[c][k + 3]), kLog2e, nc.w));
            }
          const int jc0 = j * k2BN + g * kCols;              // first column position of this thread
          if (!vrow) {
#pragma unroll
            for (int k = 0; k < kCols; ++k) pr[k] = 0.f;
          } else if ((j == T - 1) && (jc0 + kCols > p.L)) {
#pragma unroll
            for (int k = 0; k < kCols; ++k) if (jc0 + k >= p.L) pr[k] = 0.f;
          }
        };
        if (!has_t) {
          // ---- X = P: S(j) -> registers -> P -> one of the two X buffers in the T columns
          for (int j = 0; j < T; ++j) {
            float* cv = stage_cols(j);
            float pr[kCols];
            {
              uint32_t sv[kLd][32];
              load_s(sv);
              exp_s(sv, cv, j, pr);
            }
            uint32_t pk[kCols / 2];
#pragma unroll
            for (int k = 0; k < kCols / 2; ++k) pk[k] = pack16x2<XBF>(pr[2 * k], pr[2 * k + 1]);
            // X -> buffer xb once the PV that read its previous content has completed
            const uint32_t xb = nt_tile++ & 1u;
            const uint32_t xu = xb ? xuse1 : xuse0;
            if (xu > 0) { warp_mbar_wait(o_full + xb, (xu - 1) & 1, lane, 24); tc_fence_after(); }
            const uint32_t xaddr = tmem + lane_base + kFTmemX + xb * (k2BN / 2) + (uint32_t)(g * (kCols / 2));
            if constexpr (kCols == 64) tmem_st32(xaddr, pk);
            else tmem_st16(xaddr, pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(xb ? x_full_l1 : x_full_l0);
            if (xb) ++xuse1; else ++xuse0;
            last_xb = xb;
          }
        } else {
          // ---- X = P (T - d), software-pipelined: P(j) is in registers when T(j) arrives; T(j) -> X(j) in registers (the T
          // buffer goes back to the issuer), then S(j+1) -> registers, then X(j) is stored over the first half of the S columns
          // this warp has just read (no other warp's columns are touched, and s_full(j+1) implies that PV(j-1), which read
          // X(j-1) from there, has completed), then P(j+1) is exponentiated while PV(j) and T(j+1) run.
          float pr[kCols];
          float* cv = stage_cols(0);
          {
            uint32_t sv[kLd][32];
            load_s(sv);
            exp_s(sv, cv, 0, pr);
          }
          // T(j) -> X(j) = P(j) (T(j) - d) packed to 16 bits, in two 32-column chunks (P x 64 + T x 64 + the packed result
          // would not fit the register file); the T buffer goes back to the issuer as soon as the last chunk has left TMEM
          auto t_to_x = [&](const float* cvj, uint32_t (&pk)[kCols / 2]) {
            warp_mbar_wait(t_full, tcnt & 1, lane, 21);
            ++tcnt;
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < kLd; ++c) {
              uint32_t tv[32];
              tmem_ld32(tTg + c * 32, tv);
              tmem_ld_wait();
              if (c == kLd - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(t_free_l);
              }
#pragma unroll
              for (int k = 0; k < 32; k += 4) {
                float4 dc = make_float4(drow, drow, drow, drow);
                if (vcol) dc = *reinterpret_cast<const float4*>(cvj + 128 + g * kCols + c * 32 + k);
                const float x0 = pr[c * 32 + k + 0] * (__uint_as_float(tv[k + 0]) - dc.x);
                const float x1 = pr[c * 32 + k + 1] * (__uint_as_float(tv[k + 1]) - dc.y);
                const float x2 = pr[c * 32 + k + 2] * (__uint_as_float(tv[k + 2]) - dc.z);
                const float x3 = pr[c * 32 + k + 3] * (__uint_as_float(tv[k + 3]) - dc.w);
                pk[c * 16 + (k >> 1)] = pack16x2<XBF>(x0, x1);
                pk[c * 16 + (k >> 1) + 1] = pack16x2<XBF>(x2, x3);
              }
            }
          };
          auto store_x = [&](const uint32_t (&pk)[kCols / 2]) {
            // only the last tile of a phase (no S(j+1) before it) really waits here
            if (xuse0 > 0) { warp_mbar_wait(o_full + 0, (xuse0 - 1) & 1, lane, 24); tc_fence_after(); }
            if constexpr (kCols == 64) tmem_st32(tSg, pk);
            else tmem_st16(tSg, pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(x_full_l0);
            ++xuse0;
          };
#pragma unroll 1
          for (int j = 0; j + 1 < T; ++j) {
            uint32_t pk[kCols / 2];
            t_to_x(cv, pk);
            if (warp == 0 && pi == K.phase0) FTR(j, 11);
            cv = stage_cols(j + 1);
            uint32_t sv[kLd][32];
            load_s(sv);
            if (warp == 0 && pi == K.phase0) FTR(j, 9);
            store_x(pk);
            if (warp == 0 && pi == K.phase0) FTR(j, 12);
            exp_s(sv, cv, j + 1, pr);
          }
          {
            uint32_t pk[kCols / 2];
            t_to_x(cv, pk);
            store_x(pk);
          }
          last_xb = 0;
        }
        FTI(1 + (pi - K.phase0), clock64());
      }
      // ---- drain: PV completions arrive in order; the last PV of the item read buffer last_xb
      warp_mbar_wait(o_full + last_xb, ((last_xb ? xuse1 : xuse0) - 1) & 1, lane, 22);
      tc_fence_after();
      FTI(5, clock64());
      const int c0 = g * kCh;
      if (K.out_mode == 0) {
        // dQ: bf16, position-major [N][Lp][C] (contiguous 64-byte pieces per chunk) and channel-major [N][C][Lp]
        // (lanes = consecutive positions).  Padding rows are written as zeros: the GEMMs that follow contract over them.
        unsigned short* ot = K.out_t + ((size_t)n * p.Lp + row) * kC + c0;
        unsigned short* oc = K.out_c + ((size_t)n * kC + c0) * p.Lp + row;
#pragma unroll 1
        for (int ch = 0; ch < kCh / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(tOg + ch * 32, o);
          tmem_ld_wait();
          uint32_t w[16];
#pragma unroll
          for (int k = 0; k < 16; ++k)
            w[k] = vrow ? pack_bf16x2(__uint_as_float(o[2 * k]) * inv_gs, __uint_as_float(o[2 * k + 1]) * inv_gs) : 0u;
          uint4* d4 = reinterpret_cast<uint4*>(ot + ch * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) d4[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
#pragma unroll
          for (int k = 0; k < 32; ++k)
            oc[(size_t)(ch * 32 + k) * p.Lp] = (unsigned short)((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xFFFFu));
        }
      } else {
        float* acc = K.acc + ((size_t)n * kC + c0) * p.L + row;
#pragma unroll 1
        for (int ch = 0; ch < kCh / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(tOg + ch * 32, o);
          tmem_ld_wait();
          if (vrow) {
            float old[32];      // all loads before the first store (a run-time stride would serialise load -> add -> store)
#pragma unroll
            for (int k = 0; k < 32; ++k) old[k] = acc[(size_t)(ch * 32 + k) * p.L];
#pragma unroll
            for (int k = 0; k < 32; ++k) acc[(size_t)(ch * 32 + k) * p.L] = fmaf(__uint_as_float(o[k]), inv_gs, old[k]);
          }
        }
      }
      FTI(6, clock64());
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kFMmaWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 512);
  }
}

}  // namespace coattn
