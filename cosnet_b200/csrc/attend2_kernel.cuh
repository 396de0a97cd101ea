// attend2: the CTA-pair (cta_group::2) version of the fused affinity / softmax / attend kernel.
//
// One cluster of two CTAs owns a 256-row query tile (128 rows per CTA) of one (sample, pass) and sweeps all
// key/value tiles of 128 positions.  tcgen05.mma.cta_group::2 runs M = 256 MMAs across both SMs:
//   S(j)  = Q K(j)^T : M 256, N 128, K 256 -> 16 x (256x128x16); A = Q tile from shared memory (each CTA its own
//           128 rows), B = key tile, each CTA stages HALF of the 128 key rows (64 x 256, 32 KB)
//   O    += P(j) V(j): M 256, N 256, K 128 ->  8 x (256x256x16); A = P from TMEM, B = value tile, each CTA stages
//           HALF of the 256 channels (128 x 128, 32 KB)
// so every MMA runs at the full N >= 128 rate and each SM pulls half of each K/V tile from L2.
//
//   warps 0-7   softmax + drain, two warpgroups: warp w and w+4 own the same 32 query rows (TMEM lane quadrant
//               w % 4); warpgroup 0 handles key columns [0,64) of each tile and channels [0,128) of the drain,
//               warpgroup 1 the other halves.  Row max / row sum / gate dot are exchanged through shared memory.
//   warp 8      TMA producer: query tile + key tiles          warp 10   TMA producer: value tiles
//   warp 9      MMA issuer (leader CTA only) + TMEM allocator
//   warp 11     copies this CTA's 128 positions of the original fp32 features into channels [256,512) of the
//               concat (the passthrough half, :186-187) while the tensor pipe is busy -- ~3.5 B/clk per SM
//
//   TMEM columns (per CTA): [0,256) O accumulator | [256,384) S (one buffer) | [384,448) P buffer 0 | [448,512) P buffer 1
//   The softmax warps pull S(j) into registers and release the S buffer at once (s_free), so S(j+1) is computed while
//   they exponentiate; P(j) goes to its own buffer, so S never waits for a PV.
//   Barriers with a "(L)" are only used in the leader CTA and are signalled from both CTAs.
#pragma once
#include <type_traits>

#include "coattn_kernels.cuh"

namespace coattn {

constexpr int k2BM = 128;           // query rows per CTA (256 per pair)
constexpr int k2BN = 128;           // key/value positions per tile (pair wide)
constexpr int k2VStages = 2;
constexpr int k2QBytes = k2BM * kC * 2;          // 64 KB : 4 k-blocks x [128 rows x 128 B]
constexpr int k2KBytes = (k2BN / 2) * kC * 2;    // 32 KB : 4 k-blocks x [ 64 rows x 128 B]   (this CTA's key rows)
constexpr int k2VBytes = (kC / 2) * k2BN * 2;    // 32 KB : 2 k-blocks x [128 rows x 128 B]   (this CTA's channels)
constexpr uint32_t k2TmemO = 0;
constexpr uint32_t k2TmemS = 256;   // 128 columns: fp32 affinity tile
constexpr uint32_t k2TmemP = 384;   // 2 x 64 columns: 16-bit softmax numerators of two consecutive tiles

// G = softmax warps per TMEM lane quadrant ("column groups"): the G warps of a quadrant own the same 32 query rows and
// split the 128 key columns of a tile (and the 256 channels of the drain) G ways.
//   G = 2:  8 softmax warps, 64 columns / 128 channels per thread, 384 threads, 3 key stages
//   G = 4: 16 softmax warps, 32 columns /  64 channels per thread, 640 threads (<= 102 registers), 2 key stages
template <int G>
struct Attend2Cfg {
  static constexpr int kSoftmaxWarps = 4 * G;
  static constexpr int kKProducerWarp = 4 * G;
  static constexpr int kMmaWarp = 4 * G + 1;
  static constexpr int kVProducerWarp = 4 * G + 2;
  static constexpr int kCopyWarp = 4 * G + 3;   // passthrough half of the concat: cat[:, C + c, rows] = v[:, c, rows]
  static constexpr int kThreads = 32 * (4 * G + 4);
  static constexpr int kKStages = (G == 2) ? 3 : 2;
  static constexpr int kCols = k2BN / G;            // key columns per softmax thread
  static constexpr int kLoads = kCols / 32;         // 32-column TMEM loads per thread and tile
  static constexpr int kChans = kC / G;             // channels per thread in the drain / rescale
  static constexpr int kChunks = kChans / 32;
  static constexpr int kScratchBytes = 2 * G * 128 * 4;   // exchange buffer [parity][group][row]
  static constexpr int kSmemBytes = k2QBytes + kKStages * k2KBytes + k2VStages * k2VBytes + kScratchBytes + 256;
  static_assert(kSmemBytes <= 232448, "attend2 shared memory exceeds the 227 KB per-CTA limit");
};

// Debug builds (-DCOATTN_TRACE2): per-item clock64 stamps of CTA 0 (MMA issuer: slots 0-4, softmax warp 0: slots 8-12),
// dumped by the host after the third launch.  No device printf -- it perturbs the pipeline it is meant to observe.
#ifdef COATTN_TRACE2
__device__ long long g_attend2_trace[16 * 16];
#define TRG(slot) do { if (blockIdx.x == 0 && lane == 0 && it < 16) g_attend2_trace[it * 16 + (slot)] = clock64(); } while (0)
// per-tile stamps of item 3 (warm): MMA issuer slots 0-3, softmax warp 0 slots 4-7
__device__ long long g_attend2_tiles[32 * 8];
__device__ long long g_attend2_warps[2 * 32 * 16];   // [saw S | arrived P][tile][softmax warp] of item 3, CTA 0
#define TRT(j, slot) do { if (blockIdx.x == 0 && lane == 0 && it == 3 && (j) < 32) g_attend2_tiles[(j) * 8 + (slot)] = clock64(); } while (0)
#else
#define TRG(slot) do {} while (0)
#define TRT(j, slot) do {} while (0)
#endif

struct Attend2Params {
  float* z;     // [2][N][C][L] raw attended features or null
  float* lse;   // [2][N][L]
  void* cat_a;  // [N][2C][L] or null: fused gate epilogue (see AttendParams); fp32, or 16-bit elements with IO16
  void* cat_b;
  float* mask;  // [2][N][L] or null
  const float* gate_w;
  const float* gate_b;
  const void* v_a;    // [N][C][L] original features (fp32, or 16-bit with IO16), or null: when set (together with cat_*),
  const void* v_b;    //           the copy warp also writes the passthrough half of the concat (:186-187)
  int out_channels;   // channels per sample of cat_*: 2C (concat layout) or C (gated half only, no passthrough)
  int N, L, Lp;
  int q_pairs;   // ceil(L / 256)
  int kv_tiles;  // ceil(L / 128)
  int num_items; // passes * N * q_pairs
  int splits;    // 1, or key-range splits per item (COATTN_FLAG_SPLIT_KEYS, few items): unit = (item, part) sweeps key tiles
                 // [T part / splits, T (part + 1) / splits) and writes z / lse of ITS range to z[part], lse[part]
                 // ([splits][passes][N]...); merge_gate_kernel combines the parts.  No fused gate / passthrough then.
  int passes;    // 2, or 1 = frame-A outputs only (pass 0; test.py averages x1 only, test.py:301)
  int q_group;   // 1, or (passes == 1 only) pairs per query frame: pair n uses sample n / q_group of V_a and Q = W V_a
  // row of sample 0 in each tensor map (MN path).  Workspace planes X = [B16, A16, Q16]: xq_row0 = 2 N C, xb_row0 = 0,
  // v0_row0 = 0, v1_row0 = N C;  16-bit features consumed in place (IO16, one map per tensor): all 0
  int xq_row0;   // Q16 in tmap_q
  int xb_row0;   // V_b in tmap_k
  int v0_row0;   // values of pass 0 (V_b) in tmap_v
  int v1_row0;   // values of pass 1 (V_a) in tmap_v1
  unsigned* status;   // status block of the workspace or null (FOLD: |Q| beyond the fp16 range raises COATTN_STATUS_OVERFLOW_Q)
};

// exchange one float between the G threads that own the same query row (warps quad, quad + 4, ...): every thread gets
// the maximum / the sum over all G contributions, combined in group order so that all of them hold the same bits
template <int G>
__device__ __forceinline__ float group_exchange_max(float v, float* xbuf, uint32_t seq, int g, int row, int quad) {
  float* base = xbuf + (seq & 1u) * (G * 128);
  base[g * 128 + row] = v;
  named_bar_sync(1 + quad, 32 * G);
  float r = base[row];
#pragma unroll
  for (int i = 1; i < G; ++i) r = fmaxf(r, base[i * 128 + row]);
  return r;
}
template <int G>
__device__ __forceinline__ float group_exchange_sum(float v, float* xbuf, uint32_t seq, int g, int row, int quad) {
  float* base = xbuf + (seq & 1u) * (G * 128);
  base[g * 128 + row] = v;
  named_bar_sync(1 + quad, 32 * G);
  float r = base[row];
#pragma unroll
  for (int i = 1; i < G; ++i) r += base[i * 128 + row];
  return r;
}

// MN = false: queries / keys come from the position-major arrays T = [Bt, Qt] ([Lp][C], K-major operands).
// MN = true : queries / keys come from the channel-major arrays X = [B16, A16, Q16] ([C][Lp], the NCHW orientation)
//             as MN-major UMMA operands -- no transposed copies of the features exist at all.
// IO16 = true: the features (passthrough source) and the cat_* outputs are 16-bit (fp16, or bf16 with BF16); the
//             operands may then be read straight from the caller's tensors (one tensor map per tensor, see *_row0).
// SPLIT = true: the work units are (item, key-range part) pairs (Attend2Params::splits > 1); the default instantiation
//             has splits == 1 folded away at compile time, so its loops are the plain per-item sweeps.
// FOLD = true (MN only): the W projection (:158-159) happens INSIDE the kernel.  Every item first projects its own 256-row
//             query tile on the tensor cores -- pass 0: Q_I = (W A_I), pass 1: Q'_J = (W^T B_J), so that S^T = Q'^T A needs no
//             projected KEYS -- from the raw feature tile (TMA, MN-major A operand) and W (two 32 KB blocks through the key
//             ring) into the TMEM columns of S and P, which are idle between items; the softmax warps round it to 16 bits
//             and write it back over the raw tile as a K-major operand.  The projected plane Q16 and the project_mn launch
//             no longer exist in the forward; tmap_q then holds V_a and xq_row0 its first row.
template <bool BF16, bool MN, int G, bool IO16 = false, bool SPLIT = false, bool FOLD = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Attend2Cfg<G>::kThreads, 1)
attend2_kernel(const __grid_constant__ CUtensorMap tmap_q,  // !MN: T [2*N*Lp][C], box {64, 128};  MN: holds Q16 (FOLD: V_a), rows [..][C] x Lp, box {64, 256}
               const __grid_constant__ CUtensorMap tmap_k,  // !MN: T [2*N*Lp][C], box {64, 64};   MN: holds V_b, box {64, 256}
               const __grid_constant__ CUtensorMap tmap_v,  // holds V_b (values of pass 0), box {64, 128}
               const __grid_constant__ CUtensorMap tmap_v1, // holds V_a (values of pass 1), box {64, 128}
               const __grid_constant__ CUtensorMap tmap_w0, // FOLD: W16 [C][C], box {64, 64}  (K-major blocks, pass 0)
               const __grid_constant__ CUtensorMap tmap_w1, // FOLD: W16 [C][C], box {64, 256} (MN-major block = W^T, pass 1)
               Attend2Params p) {
  static_assert(!FOLD || MN, "the in-kernel projection reads channel-major feature tiles");
  using Cfg = Attend2Cfg<G>;
  const int kSplits = SPLIT ? p.splits : 1;
  constexpr int k2KStages = Cfg::kKStages;
  constexpr int k2SoftmaxWarps = Cfg::kSoftmaxWarps;
  constexpr int k2KProducerWarp = Cfg::kKProducerWarp, k2MmaWarp = Cfg::kMmaWarp, k2VProducerWarp = Cfg::kVProducerWarp,
                k2CopyWarp = Cfg::kCopyWarp;
  constexpr int k2ScratchBytes = Cfg::kScratchBytes;
  pdl_wait();      // launched with programmatic stream serialization (the cast kernel before it may still be draining)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + k2QBytes;
  uint8_t* sV = sK + k2KStages * k2KBytes;
  float* xbuf = reinterpret_cast<float*>(sV + k2VStages * k2VBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xbuf) + k2ScratchBytes);
  uint64_t* q_full = bars + 0;                  // (L) query tiles of both CTAs landed (tx bytes)
  uint64_t* q_empty = bars + 1;                 // every affinity MMA of the item completed
  uint64_t* k_full = bars + 2;                  // (L) [k2KStages] both halves of a key tile landed
  uint64_t* k_empty = k_full + k2KStages;       // [k2KStages]
  uint64_t* v_full = k_empty + k2KStages;       // (L) [k2VStages]
  uint64_t* v_empty = v_full + k2VStages;       // [k2VStages]
  uint64_t* s_full = v_empty + k2VStages;       // S(j) complete (both CTAs)
  uint64_t* s_free = s_full + 1;                // (L) S(j) sits in the registers of every softmax warp of the pair
  uint64_t* p_full = s_full + 2;                // (L) [2] one arrival per softmax warp of each CTA
  uint64_t* o_full = p_full + 2;                // [2] PV(j) complete, on barrier j & 1: a waiter can then never be two
                                                //     phases behind (the next completion on the same barrier needs P(j+2))
  uint64_t* proj_full = o_full + 2;             // FOLD: the projected query tile of the item is complete in TMEM (both CTAs)
  uint64_t* qk_ready = o_full + 3;              // FOLD (L): every softmax warp of the pair has written its rows of Q to shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();      // the swizzled operand tiles need 1024-byte alignment
  if (warp == k2KProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_v1);
    if constexpr (FOLD) { tma_prefetch_desc(&tmap_w0); tma_prefetch_desc(&tmap_w1); }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < k2KStages; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
    for (int s = 0; s < k2VStages; ++s) { mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 2 * k2SoftmaxWarps);
    for (int b = 0; b < 2; ++b) mbar_init(p_full + b, 2 * k2SoftmaxWarps);
    mbar_init(o_full + 0, 1);
    mbar_init(o_full + 1, 1);
    mbar_init(proj_full, 1);
    mbar_init(qk_ready, 2 * k2SoftmaxWarps);
    fence_mbar_init();
  }
  if (warp == k2MmaWarp) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int T = p.kv_tiles;
  // keys the MMAs touch in the last (ragged) tile: valid keys rounded up to the MMA's N granule of 16
  const int n_last = ((p.L - (T - 1) * k2BN) + 15) & ~15;

  if (warp == k2KProducerWarp) {
    // ------------------------------------------------------------------ TMA producer: query tile + key tiles
    if (lane == 0) {
      const uint32_t q_full_l = mapa_u32(smem_u32(q_full), 0);
      uint32_t it = 0, cnt = 0;
      for (int unit = cluster_id; unit < p.num_items * kSplits; unit += num_clusters, ++it) {
        const int item = unit / kSplits, part = unit % kSplits;
        const int j0 = (T * part) / kSplits, j1 = (T * (part + 1)) / kSplits;   // key tiles [j0, j1) of this unit
        const int qp = item % p.q_pairs;
        const int np = item / p.q_pairs;
        const int pass = (p.passes == 2) ? (np & 1) : 0;
        const int n = (p.passes == 2) ? (np >> 1) : np;
        mbar_wait(q_empty, (it & 1) ^ 1, 1);
        if (rank == 0) mbar_arrive_expect_tx(q_full, 2 * k2QBytes);
        if constexpr (MN) {
          // pass 0: queries Q16 (tmap_q), keys V_b (tmap_k);  pass 1: queries V_b, keys Q16
          // FOLD: tmap_q holds V_a: pass 0 queries = (W V_a) projected here, keys V_b; pass 1 queries = (W^T V_b), keys V_a
          const CUtensorMap* mq = pass ? &tmap_k : &tmap_q;
          const CUtensorMap* mk = pass ? &tmap_q : &tmap_k;
          const int nq = pass ? n : n / p.q_group;      // q_group > 1 only with passes == 1 (pass 0: queries from V_a)
          const int qch0 = pass ? p.xb_row0 + n * kC : p.xq_row0 + nq * kC;
          const int kch0 = pass ? p.xq_row0 + n * kC : p.xb_row0 + n * kC;
          const int qpos0 = qp * (2 * k2BM) + (int)rank * k2BM;
#pragma unroll
          for (int mc = 0; mc < 2; ++mc)     // two 64-position chunks x 256 channel rows
            tma_load_2d_pair(sQ + mc * 32768, mq, q_full_l, qpos0 + mc * 64, qch0);
          if constexpr (FOLD) {
            // W through the key ring, one 32 KB block per half of the 256 projected channels.  Block h, CTA r:
            //   pass 0  B operand [N = c_out][K = c_in] K-major: rows c_out in [128 h + 64 r, + 64), four 64-wide k-blocks
            //   pass 1  B operand W^T = [N = c_in][K = c_out] MN-major: columns c_in in [128 h + 64 r, + 64), all 256 rows
            for (int hh = 0; hh < 2; ++hh, ++cnt) {
              const uint32_t s = cnt % k2KStages, ph = (cnt / k2KStages) & 1;
              mbar_wait(k_empty + s, ph ^ 1, 4);
              if (rank == 0) mbar_arrive_expect_tx(k_full + s, 2 * k2KBytes);
              const uint32_t full_l = mapa_u32(smem_u32(k_full + s), 0);
              const int c0 = 128 * hh + 64 * (int)rank;
              if (pass == 0) {
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
                  tma_load_2d_pair(sK + s * k2KBytes + kb * ((k2BN / 2) * 128), &tmap_w0, full_l, kb * 64, c0);
              } else {
                tma_load_2d_pair(sK + s * k2KBytes, &tmap_w1, full_l, c0, 0);
              }
            }
          }
          for (int j = j0; j < j1; ++j, ++cnt) {
            const uint32_t s = cnt % k2KStages, ph = (cnt / k2KStages) & 1;
            mbar_wait(k_empty + s, ph ^ 1, 2);
            if (rank == 0) mbar_arrive_expect_tx(k_full + s, 2 * k2KBytes);
            const uint32_t full_l = mapa_u32(smem_u32(k_full + s), 0);
            // this CTA's half of the keys: 64 positions (n_last / 2 in the ragged last tile) x 256 channel rows
            const int kpos = j * k2BN + (int)rank * ((j == T - 1) ? (n_last / 2) : (k2BN / 2));
            tma_load_2d_pair(sK + s * k2KBytes, mk, full_l, kpos, kch0);
          }
        } else {
          const int qrow0 = ((1 - pass) * p.N + n) * p.Lp + qp * (2 * k2BM) + (int)rank * k2BM;
          const int krow_base = (pass * p.N + n) * p.Lp;
          const int krow0 = krow_base + (int)rank * (k2BN / 2);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) tma_load_2d_pair(sQ + kb * (k2BM * 128), &tmap_q, q_full_l, kb * 64, qrow0);
          for (int j = j0; j < j1; ++j, ++cnt) {
            const uint32_t s = cnt % k2KStages, ph = (cnt / k2KStages) & 1;
            mbar_wait(k_empty + s, ph ^ 1, 2);
            if (rank == 0) mbar_arrive_expect_tx(k_full + s, 2 * k2KBytes);
            const uint32_t full_l = mapa_u32(smem_u32(k_full + s), 0);
            // ragged last tile: the MMA only uses n_last (multiple of 16) keys, CTA r supplies keys [r, r+1) * n_last/2
            const int krow = (j == T - 1) ? krow_base + j * k2BN + (int)rank * (n_last / 2) : krow0 + j * k2BN;
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
              tma_load_2d_pair(sK + s * k2KBytes + kb * ((k2BN / 2) * 128), &tmap_k, full_l, kb * 64, krow);
          }
        }
      }
    }
  } else if (warp == k2VProducerWarp) {
    // ------------------------------------------------------------------ TMA producer: value tiles
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int unit = cluster_id; unit < p.num_items * kSplits; unit += num_clusters) {
        const int item = unit / kSplits, part = unit % kSplits;
        const int j0 = (T * part) / kSplits, j1 = (T * (part + 1)) / kSplits;   // key tiles [j0, j1) of this unit
        const int np = item / p.q_pairs;
        const int vpass = (p.passes == 2) ? (np & 1) : 0;
        const int vn = (p.passes == 2) ? (np >> 1) : np;
        const CUtensorMap* mv = vpass ? &tmap_v1 : &tmap_v;
        const int vrow0 = (vpass ? p.v1_row0 : p.v0_row0) + vn * kC + (int)rank * (kC / 2);
        for (int j = j0; j < j1; ++j, ++cnt) {
          const uint32_t s = cnt % k2VStages, ph = (cnt / k2VStages) & 1;
          mbar_wait(v_empty + s, ph ^ 1, 3);
          if (rank == 0) mbar_arrive_expect_tx(v_full + s, 2 * k2VBytes);
          const uint32_t full_l = mapa_u32(smem_u32(v_full + s), 0);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
            tma_load_2d_pair(sV + s * k2VBytes + kb * ((kC / 2) * 128), mv, full_l, j * k2BN + kb * 64, vrow0);
        }
      }
    }
  } else if (warp == k2CopyWarp) {
    // ------------------------------------------------------------------ passthrough copy (bit exact; fp32, or 16-bit with IO16)
    if (p.v_a != nullptr && p.cat_a != nullptr) {
      using Elem = typename std::conditional<IO16, unsigned short, float>::type;
      using Vec4 = typename std::conditional<IO16, uint2, float4>::type;     // four elements
      const bool vec = (p.L % 4 == 0) &&
                       (((reinterpret_cast<uintptr_t>(p.v_a) | reinterpret_cast<uintptr_t>(p.v_b) |
                          reinterpret_cast<uintptr_t>(p.cat_a) | reinterpret_cast<uintptr_t>(p.cat_b)) & (sizeof(Vec4) - 1)) == 0);
      for (int item = cluster_id; item < p.num_items; item += num_clusters) {
        const int qp = item % p.q_pairs;
        const int np = item / p.q_pairs;
        const int pass = (p.passes == 2) ? (np & 1) : 0;
        const int n = (p.passes == 2) ? (np >> 1) : np;
        const int row0 = qp * (2 * k2BM) + (int)rank * k2BM;
        const Elem* src = reinterpret_cast<const Elem*>(pass ? p.v_b : p.v_a) + (size_t)(pass ? n : n / p.q_group) * kC * p.L;
        Elem* dst = reinterpret_cast<Elem*>(pass ? p.cat_b : p.cat_a) + ((size_t)n * 2 * kC + kC) * p.L;
        if (vec) {
          const int r = row0 + 4 * lane;
          if (r < p.L) {
#pragma unroll 1
            for (int c = 0; c < kC; c += 8) {
              Vec4 t[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) t[u] = __ldcs(reinterpret_cast<const Vec4*>(src + (size_t)(c + u) * p.L + r));
#pragma unroll
              for (int u = 0; u < 8; ++u) __stcs(reinterpret_cast<Vec4*>(dst + (size_t)(c + u) * p.L + r), t[u]);
            }
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < kC; c += 4) {
            Elem t[16];      // four channels x four row groups in flight
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int r = row0 + (u & 3) * 32 + lane;
              t[u] = (r < p.L) ? __ldcs(src + (size_t)(c + (u >> 2)) * p.L + r) : Elem(0);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int r = row0 + (u & 3) * 32 + lane;
              if (r < p.L) __stcs(dst + (size_t)(c + (u >> 2)) * p.L + r, t[u]);
            }
          }
        }
      }
    }
  } else if (warp == k2MmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; uniform control flow)
    if (rank == 0) {
      // FOLD: the query tile is the projected one, written by the softmax warps as a K-major operand
      constexpr bool QMN = MN && !FOLD;
      constexpr uint32_t idesc_s = make_idesc_16_major(2 * k2BM, k2BN, BF16, QMN, MN);
      constexpr uint32_t idesc_o = make_idesc_16(2 * k2BM, kC, BF16);
      const uint32_t idesc_s_last = make_idesc_16_major(2 * k2BM, (uint32_t)n_last, BF16, QMN, MN);
      const int ksteps_last = n_last / 16;
      uint32_t it = 0, kcnt = 0, vcnt = 0;
      uint32_t rcnt = 0;      // blocks taken from the key ring (FOLD: S tiles + W blocks; otherwise == kcnt)
      uint32_t pphase0 = 0, pphase1 = 0;
      const uint32_t tO = tmem + k2TmemO;
      const uint32_t tS = tmem + k2TmemS;
      const uint64_t qd0 = QMN ? make_sdesc_mn_sw128(smem_u32(sQ), 32768, 1024) : make_sdesc_k_sw128(smem_u32(sQ));
      const uint32_t sK_addr = smem_u32(sK);
      const uint32_t sV_addr = smem_u32(sV);
      for (int unit = cluster_id; unit < p.num_items * kSplits; unit += num_clusters, ++it) {
        const int part = unit % kSplits;
        const int j0 = (T * part) / kSplits, j1 = (T * (part + 1)) / kSplits;   // key tiles [j0, j1) of this unit
        // S(j) of this item; kcnt counts every S tile of the kernel (key stage ring and s_free phases)
        auto issue_s = [&](int j) {
          const uint32_t s = rcnt % k2KStages, ph = (rcnt / k2KStages) & 1;
          warp_mbar_wait(k_full + s, ph, lane, 10);
          // the single S buffer: every softmax warp of the pair has pulled the previous tile into registers
          if (kcnt > 0) warp_mbar_wait(s_free, (kcnt - 1) & 1, lane, 12);
          tc_fence_after();
          const uint64_t kd0 = MN ? make_sdesc_mn_sw128(sK_addr + s * k2KBytes, 32768, 1024) : make_sdesc_k_sw128(sK_addr + s * k2KBytes);
          const uint32_t idesc = (j == T - 1) ? idesc_s_last : idesc_s;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kC / 16; ++kk) {
              // K-major: k-block kk/4 + 32 B per 16 channels inside the 128-byte row;  MN-major: 16 channel rows = 2048 B
              const uint64_t ad = qd0 + (uint64_t)((QMN ? kk * 2048 : ((kk >> 2) * (k2BM * 128) + (kk & 3) * 32)) >> 4);
              const uint64_t bd = kd0 + (uint64_t)((MN ? kk * 2048 : ((kk >> 2) * ((k2BN / 2) * 128) + (kk & 3) * 32)) >> 4);
              umma2_ss(tS, ad, bd, idesc, kk > 0);
            }
            umma2_commit_mc(k_empty + s, 3);
            umma2_commit_mc(s_full, 3);
            if (j == j1 - 1) umma2_commit_mc(q_empty, 3);    // last affinity tile of the unit: the query tile is free
          }
          __syncwarp();
          ++kcnt;
          ++rcnt;
        };
        TRG(0);
        warp_mbar_wait(q_full, it & 1, lane, 11);
        tc_fence_after();
        TRG(1);
        if constexpr (FOLD) {
          // Projection of this item's query tile into the TMEM columns of S and P (idle between items: the previous item's
          // PVs precede these MMAs in the pipe, its last S sits in registers once s_free says so):
          //   D[i, c] = sum_k raw[k, i] Wop[c, k]     M 256 (query rows of the pair) x N 128 per block x K 256
          const int np_ = (unit / kSplits) / p.q_pairs;
          const int pass_ = (p.passes == 2) ? (np_ & 1) : 0;
          const uint64_t rawd0 = make_sdesc_mn_sw128(smem_u32(sQ), 32768, 1024);
          const uint32_t idesc_p = pass_ ? make_idesc_16_major(2 * k2BM, 128, BF16, true, true)
                                         : make_idesc_16_major(2 * k2BM, 128, BF16, true, false);
          for (int hh = 0; hh < 2; ++hh, ++rcnt) {
            const uint32_t s = rcnt % k2KStages, ph = (rcnt / k2KStages) & 1;
            warp_mbar_wait(k_full + s, ph, lane, 15);
            if (hh == 0 && kcnt > 0) warp_mbar_wait(s_free, (kcnt - 1) & 1, lane, 16);
            tc_fence_after();
            const uint32_t sb = sK_addr + s * k2KBytes;
            const uint64_t wk0 = make_sdesc_k_sw128(sb), wm0 = make_sdesc_mn_sw128(sb, 32768, 1024);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < kC / 16; ++kk) {
                const uint64_t ad = rawd0 + (uint64_t)((kk * 2048) >> 4);
                const uint64_t bd = pass_ ? wm0 + (uint64_t)((kk * 2048) >> 4)
                                          : wk0 + (uint64_t)(((kk >> 2) * ((k2BN / 2) * 128) + (kk & 3) * 32) >> 4);
                umma2_ss(tmem + k2TmemS + (uint32_t)hh * 128, ad, bd, idesc_p, kk > 0);
              }
              umma2_commit_mc(k_empty + s, 3);
              if (hh == 1) umma2_commit_mc(proj_full, 3);
            }
            __syncwarp();
          }
          warp_mbar_wait(qk_ready, it & 1, lane, 17);      // Q (16-bit, K-major) replaced the raw tile in shared memory
          tc_fence_after();
        }
        // Tensor-pipe order per item: S(0) S(1) | S(2) PV(0) | S(3) PV(1) | ... | PV(T-1).  S(j+2) only needs the S buffer
        // back (s_free(j+1): the softmax warps hold S(j+1) in registers) and is computed while they work on tile j+1;
        // PV(j) follows when P(j) is complete.  Both conditions arrive when softmax(j) ends, and the affinity tile goes
        // first, so the next S is ready when the softmax warps come back for it.
        issue_s(j0);
        if (j1 - j0 > 1) issue_s(j0 + 1);
        TRG(2);
        for (int j = j0; j < j1; ++j) {
          const int b = (j - j0) & 1;
          TRT(j, 2);
          if (j + 2 < j1) issue_s(j + 2);
          TRT(j, 3);
          {
            const uint32_t s = vcnt % k2VStages, ph = (vcnt / k2VStages) & 1;
            warp_mbar_wait(v_full + s, ph, lane, 14);
          }
          // P(0) of an item is only produced after the previous item's O was drained, so no separate O barrier
          if (b == 0) { warp_mbar_wait(p_full + 0, pphase0, lane, 13); pphase0 ^= 1; }
          else        { warp_mbar_wait(p_full + 1, pphase1, lane, 13); pphase1 ^= 1; }
          tc_fence_after();
          if (j == 0) TRG(3);
          TRT(j, 0);
          const uint32_t s = vcnt % k2VStages;
          const uint32_t tP = tmem + k2TmemP + (uint32_t)b * (k2BN / 2);
          const uint64_t vd0 = make_sdesc_k_sw128(sV_addr + s * k2VBytes);
          const int ksteps = (j == T - 1) ? ksteps_last : k2BN / 16;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < k2BN / 16; ++kk) {
              if (kk < ksteps) {
                const uint64_t bd = vd0 + (uint64_t)(((kk >> 2) * ((kC / 2) * 128) + (kk & 3) * 32) >> 4);
                umma2_ts(tO, tP + kk * 8, bd, idesc_o, (j > j0 || kk > 0) ? 1u : 0u);
              }
            }
            umma2_commit_mc(v_empty + s, 3);
            umma2_commit_mc(o_full + b, 3);
          }
          __syncwarp();
          ++vcnt;
          if (j == T - 1) TRG(4);
          TRT(j, 1);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + drain
    constexpr int kCols = Cfg::kCols, kLoads = Cfg::kLoads, kChans = Cfg::kChans, kChunks = Cfg::kChunks;
    const int g = warp >> 2;             // column group: key columns [g kCols, (g+1) kCols) / channels [g kChans, (g+1) kChans)
    const int quad = warp & 3;           // TMEM lane quadrant
    const int rloc = quad * 32 + lane;   // query row inside this CTA's 128-row tile
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t tO = tmem + lane_base + k2TmemO + (uint32_t)(g * kChans);   // this group's channels
    const uint32_t p_full_l0 = mapa_u32(smem_u32(p_full + 0), 0);
    const uint32_t p_full_l1 = mapa_u32(smem_u32(p_full + 1), 0);
    const uint32_t s_free_l = mapa_u32(smem_u32(s_free), 0);
    uint32_t scnt = 0, it = 0, seq = 0;     // scnt: S tiles consumed so far (phase of s_full)
    // gate weights of this group's channels live in registers, one value per lane and 32-channel chunk, and are
    // broadcast by shuffles in the drain: 128 global loads per item used to queue behind the copy warp's traffic in the
    // LSU and made the gate dot the longest part of the drain
    float greg[kChunks];
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) greg[ch] = (p.cat_a != nullptr) ? __ldg(p.gate_w + g * kChans + ch * 32 + lane) : 0.f;
    // FOLD: round the projected query tile of item `item_it` to 16 bits and publish it as the affinity MMAs' A operand.
    // Called for the first item before its sweep, and for every following item from the MIDDLE of the previous item's drain
    // (after the gate-dot sweep over O, before the store sweep): the projection MMAs -- queued behind the last PV -- have
    // finished by then, and S(0), S(1) of the next item are computed while the stores drain, as they were before the
    // projection moved into the kernel.
    auto convert_q = [&](uint32_t item_it) {
        // the projected query tile: TMEM columns [256, 512) hold Q[row][c], c = column - 256 -> 16 bits -> shared memory as the
        // K-major operand of the affinity MMAs (four 64-channel k-blocks of [128 rows x 128 B], 16-byte chunks XOR row % 8),
        // over the raw tile, which every projection MMA has finished reading (proj_full)
        warp_mbar_wait(proj_full, item_it & 1, lane, 25);
        tc_fence_after();
        uint8_t* qrow = sQ + rloc * 128;
        float qmax = 0.f;
        // all of this thread's columns leave TMEM with ONE wait (four dependent load -> wait -> store rounds cost ~2.5 k
        // cycles per item on the path between the drain and the next item's first softmax; few other registers are live here)
        uint32_t o[kChunks][32];
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch) tmem_ld32(tmem + lane_base + k2TmemS + (uint32_t)(g * kChans + ch * 32), o[ch]);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch) {
          if constexpr (!BF16) {      // fp16 range guard (the pack below saturates at +-65504)
#pragma unroll
            for (int k = 0; k < 32; ++k) qmax = fmaxf(qmax, fabsf(__uint_as_float(o[ch][k])));
          }
          const int c = g * kChans + ch * 32;          // first channel of this chunk
          uint8_t* kb = qrow + (c >> 6) * (k2BM * 128);
          const int j0 = (c & 63) >> 3;                // first 16-byte chunk inside the 128-byte row
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(kb + (((j0 + q) ^ (rloc & 7)) << 4)) =
                make_uint4(pack16x2<BF16>(__uint_as_float(o[ch][8 * q + 0]), __uint_as_float(o[ch][8 * q + 1])),
                           pack16x2<BF16>(__uint_as_float(o[ch][8 * q + 2]), __uint_as_float(o[ch][8 * q + 3])),
                           pack16x2<BF16>(__uint_as_float(o[ch][8 * q + 4]), __uint_as_float(o[ch][8 * q + 5])),
                           pack16x2<BF16>(__uint_as_float(o[ch][8 * q + 6]), __uint_as_float(o[ch][8 * q + 7])));
        }
        if constexpr (!BF16) {
          if (p.status != nullptr && __any_sync(0xffffffffu, !(qmax <= 65504.0f)) && lane == 0) atomicOr(p.status, 4u);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(qk_ready), 0));
    };
    uint32_t pv_acc0 = 0, pv_acc1 = 0;      // PV tiles of even / odd local index completed by the earlier units
    for (int unit = cluster_id; unit < p.num_items * kSplits; unit += num_clusters, ++it) {
      const int item = unit / kSplits, part = unit - item * kSplits;
      const int j0 = (T * part) / kSplits, j1 = (T * (part + 1)) / kSplits;   // key tiles [j0, j1) of this unit
      const int Tu = j1 - j0;
      const int qp = item % p.q_pairs;
      const int np = item / p.q_pairs;
      const int pass = (p.passes == 2) ? (np & 1) : 0;
      const int n = (p.passes == 2) ? (np >> 1) : np;
      const int row = qp * (2 * k2BM) + (int)rank * k2BM + rloc;
      // phase index of PV(jj) (jj = tile index inside the unit) on o_full[jj & 1]: tiles of that parity in the earlier
      // units + (jj >> 1)
      const uint32_t pv_base0 = pv_acc0, pv_base1 = pv_acc1;
      pv_acc0 += (uint32_t)((Tu + 1) / 2);
      pv_acc1 += (uint32_t)(Tu / 2);
      auto wait_pv = [&](int jj, int tag) {      // PV(jj) complete; requires PV(jj - 2) to be known complete
        const uint32_t ph = ((jj & 1) ? pv_base1 : pv_base0) + (uint32_t)(jj >> 1);
        warp_mbar_wait(o_full + (jj & 1), ph & 1u, lane, tag);
      };
      // rows of this warp that lie entirely in the padding of the last query tile: no softmax math, P = 0
      // (zero MMA operands also draw less power, and this kernel runs against the power cap)
      const bool warp_is_padding = (qp * (2 * k2BM) + (int)rank * k2BM + quad * 32) >= p.L;
      if constexpr (FOLD) {
        if (it == 0) convert_q(0);
      }
      float m = -INFINITY, l = 0.0f;
      if (warp == 0) TRG(8);
      for (int j = j0; j < j1; ++j) {
        const int jr = j - j0;      // tile index inside the unit: buffer parities and PV phases count from the unit's start
        const int b = jr & 1;
        const uint32_t tSg = tmem + lane_base + k2TmemS + (uint32_t)(g * kCols);                       // this group's S columns
        const uint32_t tPg = tmem + lane_base + k2TmemP + (uint32_t)(b * (k2BN / 2) + g * (kCols / 2));  // ... and its P slot
        warp_mbar_wait(s_full, scnt & 1, lane, 20);
        ++scnt;
        tc_fence_after();
        if (warp_is_padding) {
          // nothing to read: hand the S buffer back, keep the exchange sequence in step, P = 0
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(s_free_l);
          if (jr == 0) (void)group_exchange_max<G>(0.f, xbuf, seq++, g, rloc, quad);
          else if (named_bar_red_or(1 + quad, 32 * G, false)) (void)group_exchange_max<G>(0.f, xbuf, seq++, g, rloc, quad);
          uint32_t zero[kCols / 2];
#pragma unroll
          for (int k = 0; k < kCols / 2; ++k) zero[k] = 0u;
          if (jr >= 2) { wait_pv(jr - 2, 24); tc_fence_after(); }
          if constexpr (kCols == 64) tmem_st32(tPg, zero); else tmem_st16(tPg, zero);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(b == 0 ? p_full_l0 : p_full_l1);
          continue;
        }
        if (warp == 0 && j == 0) TRG(9);
        if (warp == 0) TRT(j, 4);
#ifdef COATTN_TRACE2
        if (blockIdx.x == 0 && lane == 0 && it == 3 && j < 32) g_attend2_warps[j * 16 + warp] = clock64();
#endif
        uint32_t sv[kLoads][32];
#pragma unroll
        for (int c = 0; c < kLoads; ++c) tmem_ld32(tSg + c * 32, sv[c]);
        tmem_ld_wait();
        // S(j) is in registers: release the buffer so that S(j+1) is computed while this tile is exponentiated
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(s_free_l);
        if (j == T - 1) {
          const int nvalid = p.L - j * k2BN - g * kCols;   // may be <= 0 for the upper groups: everything masked
          if (nvalid < kCols) {
#pragma unroll
            for (int c = 0; c < kLoads; ++c)
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (c * 32 + k >= nvalid) sv[c][k] = 0xff800000u;        // -inf
          }
        }
        // four independent chains (a single deep chain of dependent max operations costs ~4 cycles per element)
        float h0 = __uint_as_float(sv[0][0]), h1 = __uint_as_float(sv[0][1]), h2 = __uint_as_float(sv[0][2]),
              h3 = __uint_as_float(sv[0][3]);
#pragma unroll
        for (int c = 0; c < kLoads; ++c)
#pragma unroll
          for (int k = (c == 0 ? 4 : 0); k < 32; k += 4) {
            h0 = fmaxf(h0, __uint_as_float(sv[c][k]));     h1 = fmaxf(h1, __uint_as_float(sv[c][k + 1]));
            h2 = fmaxf(h2, __uint_as_float(sv[c][k + 2])); h3 = fmaxf(h3, __uint_as_float(sv[c][k + 3]));
          }
        const float hmax = fmaxf(fmaxf(h0, h1), fmaxf(h2, h3));
        uint32_t pk[kCols / 2];
        // P = 2^(S log2e - m) for this thread's columns -> pk, returns their sum.  x = S log2e - m and the row sum run on
        // packed fp32 pairs (FFMA2 / FADD2): the pair ops halve two of the four per-element instructions, and every lane
        // rounds exactly like the scalar op did
        auto exp_tile = [&](float m_ref) -> float {
          const float neg_m = -m_ref * kLog2e;
          const uint64_t log2e2 = f32x2_pack(kLog2e, kLog2e), neg_m2 = f32x2_pack(neg_m, neg_m);
          uint64_t la = f32x2_pack(0.f, 0.f), lb = la;      // {l0, l1} (even pairs), {l2, l3} (odd pairs)
#pragma unroll
          for (int c = 0; c < kLoads; ++c)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              float x0, x1;
              f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(sv[c][2 * k]), __uint_as_float(sv[c][2 * k + 1])), log2e2, neg_m2), x0, x1);
              const float p0 = fast_exp2(x0);
              const float p1 = fast_exp2(x1);
              const uint32_t w = pack16x2<BF16>(p0, p1);
              pk[c * 16 + k] = w;
              const uint64_t pp = BF16 ? f32x2_pack(bf16lo_to_f32(w), bf16hi_to_f32(w)) : f32x2_pack(p0, p1);
              if (k & 1) lb = f32x2_add(lb, pp); else la = f32x2_add(la, pp);
            }
          float l0, l1, l2, l3;
          f32x2_unpack(la, l0, l1);
          f32x2_unpack(lb, l2, l3);
          return (l0 + l1) + (l2 + l3);
        };
        float lt;
        if (jr == 0) {
          // all groups of the row agree on the first tile's max: the reference point of the row
          m = group_exchange_max<G>(hmax, xbuf, seq++, g, rloc, quad);
          lt = exp_tile(m);
        } else {
          // lazy rescale: the reference max only moves when the row max jumps by more than 2^13 (fp16 P stays below
          // 8192 < 65504 and keeps its 11-bit precision at any magnitude; bf16 has the range for 2^24).  The slow
          // path costs ~3 k cycles and stalls the whole CTA pair, and it is taken for all 32 rows of a lane quadrant, so
          // with the usual 2^8 threshold spiky features (logit sigma ~ 6 in log2 units) hit it on most early tiles.
          // The tile is exponentiated against the CURRENT reference first; whether any row of the quadrant needs a new one is
          // then decided by ONE barrier-with-OR among the warps that own these rows (the per-tile exchange of the row maxima
          // through shared memory used to sit between the TMEM load and the first exponential of every tile), and only the
          // slow path exchanges the maxima and exponentiates again.  Decisions and results are the ones of the exchange.
          constexpr float kThreshold = BF16 ? 24.0f : 13.0f;
          lt = exp_tile(m);
          const bool need = (hmax - m) * kLog2e > kThreshold;
          if (named_bar_red_or(1 + quad, 32 * G, need)) {
            const float tmax = group_exchange_max<G>(hmax, xbuf, seq++, g, rloc, quad);
            const float m_new = fmaxf(m, tmax);
            const float scale = fast_exp2((m - m_new) * kLog2e);
            // S(j) follows PV(j-3) in the tensor pipe, so s_full(j) implies PV(j-3) is complete
            wait_pv(jr - 1, 21);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < kChunks; ++ch) {
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
            lt = exp_tile(m);
          }
        }
        l += lt;
        // packed P: the kCols keys of this group -> kCols / 2 columns of P buffer b, once PV(j-2) has read its previous
        // content (S(j) is issued ahead of PV(j-2), so s_full(j) does not imply it; the wait is almost always over)
        if (jr >= 2) { wait_pv(jr - 2, 24); tc_fence_after(); }
        if constexpr (kCols == 64) tmem_st32(tPg, pk); else tmem_st16(tPg, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(b == 0 ? p_full_l0 : p_full_l1);
        if (warp == 0) TRT(j, 5);
#ifdef COATTN_TRACE2
        if (blockIdx.x == 0 && lane == 0 && it == 3 && j < 32) g_attend2_warps[512 + j * 16 + warp] = clock64();
#endif
      }
      if (warp == 0) TRG(10);
      // ---- drain.  s_full(T-1) only implies PV(T-4): wait PV(T-2) first (completions are in order), then PV(T-1).
      if (Tu >= 2) wait_pv(Tu - 2, 23);
      wait_pv(Tu - 1, 22);
      tc_fence_after();
      if (warp == 0) TRG(11);
      l = group_exchange_sum<G>(l, xbuf, seq++, g, rloc, quad);
      if (warp == 0) TRG(5);
      const bool has_next = (unit + num_clusters) < p.num_items * kSplits;
      if (warp_is_padding) {     // nothing to store; keep the exchange sequence of the gate dot in step
        if (p.cat_a != nullptr) (void)group_exchange_sum<G>(0.f, xbuf, seq++, g, rloc, quad);
        if constexpr (FOLD) { if (has_next) convert_q(it + 1); }
        continue;
      }
      const float inv = 1.0f / l;
      const bool valid = row < p.L;
      const int c0 = g * kChans;
      // raw Z (kept for the backward pass): written in the same sweep over O as the gated output below; a sweep of its
      // own only when there is no fused gate (coattn_stage_attend)
      const size_t out_idx = (size_t)(part * p.passes + pass) * p.N + n;     // part = 0 unless the keys are split
      float* zcol = p.z ? p.z + (out_idx * kC + c0) * p.L + row : nullptr;
      if (zcol != nullptr && p.cat_a == nullptr) {
#pragma unroll 1
        for (int ch = 0; ch < kChunks; ++ch) {
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
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        {
          uint32_t oa[32], ob[32];
          tmem_ld32(tO, oa);
          tmem_ld_wait();
          if (warp == 0) TRG(6);
#pragma unroll
          for (int ch = 0; ch < kChunks; ++ch) {     // chunk ch + 1 streams out of TMEM while chunk ch is consumed
            uint32_t (&o)[32] = (ch & 1) ? ob : oa;
            uint32_t (&nx)[32] = (ch & 1) ? oa : ob;
            if (ch + 1 < kChunks) tmem_ld32(tO + (ch + 1) * 32, nx);
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              d0 = fmaf(__shfl_sync(0xffffffffu, greg[ch], k + 0), __uint_as_float(o[k + 0]), d0);
              d1 = fmaf(__shfl_sync(0xffffffffu, greg[ch], k + 1), __uint_as_float(o[k + 1]), d1);
              d2 = fmaf(__shfl_sync(0xffffffffu, greg[ch], k + 2), __uint_as_float(o[k + 2]), d2);
              d3 = fmaf(__shfl_sync(0xffffffffu, greg[ch], k + 3), __uint_as_float(o[k + 3]), d3);
            }
            if (ch + 1 < kChunks) tmem_ld_wait();
          }
        }
        float dot = (d0 + d1) + (d2 + d3);
        if (warp == 0) TRG(13);
        dot = group_exchange_sum<G>(dot, xbuf, seq++, g, rloc, quad);   // same summation order in every group
        if (warp == 0) TRG(14);
        if constexpr (FOLD) { if (has_next) convert_q(it + 1); }      // between the two sweeps over O
        const float logit = dot * inv + (p.gate_b ? __ldg(p.gate_b) : 0.f);
        const float gate = 1.0f / (1.0f + __expf(-logit));
        const float sc = inv * gate;
        using OutT = typename std::conditional<IO16, unsigned short, float>::type;
        OutT* ccol = reinterpret_cast<OutT*>(pass ? p.cat_b : p.cat_a) + ((size_t)n * p.out_channels + c0) * p.L + row;
        {
          uint32_t oa[32], ob[32];
          tmem_ld32(tO, oa);
          tmem_ld_wait();
          if (warp == 0) TRG(7);
#pragma unroll
          for (int ch = 0; ch < kChunks; ++ch) {
            uint32_t (&o)[32] = (ch & 1) ? ob : oa;
            uint32_t (&nx)[32] = (ch & 1) ? oa : ob;
            if (ch + 1 < kChunks) tmem_ld32(tO + (ch + 1) * 32, nx);
            if (valid) {
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                if constexpr (IO16) __stcs(ccol + (size_t)(ch * 32 + k) * p.L, cvt16<BF16>(__uint_as_float(o[k]) * sc));
                else __stcs(ccol + (size_t)(ch * 32 + k) * p.L, __uint_as_float(o[k]) * sc);
              }
              if (zcol != nullptr) {
#pragma unroll
                for (int k = 0; k < 32; ++k) zcol[(size_t)(ch * 32 + k) * p.L] = __uint_as_float(o[k]) * inv;
              }
            }
            if (ch + 1 < kChunks) tmem_ld_wait();
          }
        }
        if (valid && g == 0 && p.mask != nullptr) p.mask[(size_t)(pass * p.N + n) * p.L + row] = gate;
      }
      if constexpr (FOLD) { if (has_next && p.cat_a == nullptr) convert_q(it + 1); }      // no fused gate: after the only sweep
      if (valid && g == 0) p.lse[out_idx * p.L + row] = m + __logf(l);
      if (warp == 0) TRG(12);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA may leave (or free TMEM) while its partner can still signal it
  if (warp == k2MmaWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 512);
  }
}

}  // namespace coattn
