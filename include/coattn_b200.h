/*
 * coattn_b200 -- C ABI of the B200 (sm_100a) co-attention hot path.
 *
 * Drop-in boundary for the inline co-attention block of the reference model
 *   /root/reference/rgbd_segmentation_RAA.py:150-187  (RGB)   and   :204-238 (depth)
 * i.e. everything between the encoder outputs V_a, V_b and the inputs of reduce_channels_A/B.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the current CUDA device
 *     unless stated otherwise; the caller owns all buffers (workspace included);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it;
 *   - functions return 0 on success, a negative COATTN_E_* code for argument errors, or a positive
 *     cudaError_t value passed through from the runtime.  No exceptions cross this boundary;
 *   - no global mutable state: safe to call concurrently from several host threads on different
 *     devices (nn.DataParallel, reference train.py:493).
 *
 * Layouts (all row-major / NCHW contiguous, fp32):
 *   v_a, v_b   [N, 256, H, W]      encoder features of frame A / frame B          (:143-148)
 *   w          [256, 256]          *_similarity_weights.weight  (out, in)          (:27, :38)
 *   gate_w     [256]               gate.weight viewed [1,256,1,1]                  (:28, :39)
 *   gate_b     [1] or NULL         depth_gate.bias; NULL for the bias-free RGB gate
 *   cat_a/b    [N, 512, H, W]      concat([Z * sigmoid(gate(Z)), V], dim=1)        (:183-187)
 *   z          [2, N, 256, H*W]    raw attended features, z[0] = Z_a, z[1] = Z_b   (:169-170)
 *   lse        [2, N, H*W]         natural-log normalisers: lse[0][n][i] = log sum_j exp S[i,j]
 *                                  (softmax of :165), lse[1][n][j] = log sum_i exp S[i,j] (:164)
 *   mask       [2, N, H*W]         sigmoid gate values, mask[0] = input_mask_a, mask[1] = input_mask_b (:180-182)
 */
#ifndef COATTN_B200_H_
#define COATTN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COATTN_B200_ABI_VERSION 2

enum {
  COATTN_OK = 0,
  COATTN_E_NULL = -1,        /* a required pointer is NULL                                  */
  COATTN_E_SHAPE = -2,       /* n, h or w < 1, or c != 256                                  */
  COATTN_E_WORKSPACE = -3,   /* workspace too small or not 1024-byte aligned                */
  COATTN_E_ARCH = -4,        /* current device is not compute capability 10.x (no fallback) */
  COATTN_E_DRIVER = -5,      /* cuTensorMapEncodeTiled unavailable / failed                 */
  COATTN_E_ALIGN = -6,       /* a tensor pointer is not 16-byte aligned                     */
  COATTN_E_UNSUPPORTED = -7  /* valid request the library does not implement (yet)          */
};

/*
 * flags (bitwise or).  The 16-bit tensor-core operand format of the affinity/attend GEMMs:
 *   default (0)            IEEE fp16 operands, fp32 accumulation.  Same tcgen05 kind::f16 pipe and rate as
 *                          bf16 but a 2^-11 mantissa: meets rel-L2 <= 1e-3 on the module output at every
 *                          tested feature scale.  fp32 -> fp16 conversion saturates at +-65504.
 *   COATTN_FLAG_BF16       bf16 operands, fp32 accumulation (full fp32 exponent range; ~2e-3 rel-L2 at
 *                          train-like logit scales).
 */
#define COATTN_FLAG_BF16 1u
/*
 *   COATTN_FLAG_UNFUSED_GATE  run the gate / sigmoid / concat stage as its own HBM-bound kernel
 *                          (coattn_stage_gate) instead of fusing it into the drain of the attend kernel.
 *                          Same results to fp32 rounding; kept for per-kernel roofline measurements.
 */
#define COATTN_FLAG_UNFUSED_GATE 2u
/*
 *   COATTN_FLAG_SINGLE_CTA  run the attend stage with the single-CTA kernel (128-row query tiles, 64-position
 *                          key tiles) instead of the default CTA-pair kernel (tcgen05 cta_group::2, 256-row
 *                          query tiles, 128-position key tiles).  Same math; kept as a cross-check.
 */
#define COATTN_FLAG_SINGLE_CTA 4u
/*
 *   COATTN_FLAG_A_ONLY     compute only the frame-A outputs (cat_a, lse[0], mask[0], z[0]); cat_b is not written.
 *                          For test.py-style inference, which only averages x1 over the reference frames
 *                          (test.py:301-305): halves the attend work.  CTA-pair kernel only.
 */
#define COATTN_FLAG_A_ONLY 8u
/*
 *   COATTN_FLAG_UNFUSED_PREP  run prep (both frames) and project as separate kernels with the At operand
 *                          round-tripping through HBM, instead of the fused A-side prep + projection kernel.
 */
#define COATTN_FLAG_UNFUSED_PREP 16u
/*
 *   COATTN_FLAG_GATED_ONLY  cat_a / cat_b are [N, 256, H, W]: only Z * sigmoid(gate(Z)) is produced, the passthrough
 *                          half of the concat is not materialised.  For consumers that split the 3x3 reduce conv
 *                          (:188-189) into its two 256-channel halves, conv(cat, W) = conv(Zg, W[:, :256]) +
 *                          conv(V, W[:, 256:]), so the concat never exists.  CTA-pair kernel, fused gate only.
 */
#define COATTN_FLAG_GATED_ONLY 32u
/*
 *   COATTN_FLAG_KMAJOR     use the position-major (transposed, [L][C]) copies of the features as K-major tensor-core
 *                          operands for the affinity GEMM, as in the first version of the kernels.  Default: the
 *                          features stay channel-major ([C][L], their NCHW orientation) and are consumed as MN-major
 *                          operands, so the prep stage is a pure cast.  Same arithmetic, bit-identical results.
 */
#define COATTN_FLAG_KMAJOR 64u
/*
 *   COATTN_FLAG_SOFTMAX16  attend kernel with 16 softmax warps (four column groups per TMEM lane quadrant, 32 key columns
 *                          per thread, two key stages) instead of the default 8 (two groups, 64 columns per thread).
 *                          Measured ~3 % slower at 60x60 (the softmax chain is bound by the MUFU pipe, not by warp
 *                          count); kept as a cross-check of the column-group logic.
 */
#define COATTN_FLAG_SOFTMAX16 128u
/*
 *   COATTN_FLAG_SPLIT_KEYS  latency mode for a few pairs (one or two at 60x60: fewer work items than the 74 CTA pairs of
 *                          a B200): the key range of every (sample, pass, query tile) item is swept in up to 4 parts by
 *                          different CTA pairs, and a small HBM-bound kernel merges the parts (log-sum-exp weights) and
 *                          applies the gate / concat epilogue.  Same softmax, other summation order: results equal the
 *                          default path to the rounding of the 16-bit softmax numerators (1e-5 ... 2e-4 rel-L2), not bit for bit -- which is why it is opt-in (the default
 *                          path is batch invariant).  No effect (default path) when the batch already fills the GPU.
 *                          coattn_forward and coattn_forward_queries; not with the cross-check flags.
 */
#define COATTN_FLAG_SPLIT_KEYS 256u

/*
 *   COATTN_FLAG_PLANES_READY  the 16-bit operand planes of V_a and V_b in the workspace were already written by
 *                          coattn_stage_tail (the fused encoder tail, below): coattn_forward / coattn_stage_prep_project skip
 *                          their cast kernel.  v_a / v_b are still read for the passthrough half of the concat.
 *                          Default (channel-major) path only.
 */
#define COATTN_FLAG_PLANES_READY 512u

/*
 *   COATTN_FLAG_UNFOLDED   compute Q = W V_a with the stand-alone projection kernel (project_mn) into the workspace and let
 *                          the attend kernel read it as queries (pass 0) and KEYS (pass 1), as in the first version.
 *                          Default: the attend kernel projects the query tile of every work item itself (pass 0: W V_a,
 *                          pass 1: W^T V_b, so no projected keys are needed) -- north_star item 2, "the W projection is
 *                          pre-folded into V_a in the same kernel".  Same math; the frame-B outputs differ by the 16-bit
 *                          rounding of W^T V_b instead of W V_a (~1e-4 rel-L2).  Kept as a cross-check and for A/B timing.
 */
#define COATTN_FLAG_UNFOLDED 1024u

/*
 * Status block: the first COATTN_STATUS_WORDS 32-bit words of every workspace.  With fp16 operands (the default) the
 * fp32 -> fp16 conversions of the features and of Q = W V_a clamp at +-65504; instead of clipping silently the kernels
 * record it here (sticky bits; the library never clears them -- coattn_status_clear, or zero the words yourself):
 *   word 0   COATTN_STATUS_OVERFLOW_B / _A : a feature of V_b / V_a was outside the fp16 range (or Inf / NaN)
 *            COATTN_STATUS_OVERFLOW_Q      : an element of Q = W V_a was
 *   word 1,2 bits of max |v| over V_b / V_a seen by the cast kernels (fp32; compare as floats).  Features whose
 *            largest magnitude is below 2^-14 (fp16 subnormals) lose precision the same way: callers should scale them
 *            or select COATTN_FLAG_BF16, whose operands have the fp32 exponent range and never set anything here.
 */
#define COATTN_STATUS_WORDS 8
#define COATTN_STATUS_OVERFLOW_B 1u
#define COATTN_STATUS_OVERFLOW_A 2u
#define COATTN_STATUS_OVERFLOW_Q 4u
/* zero the status block (asynchronous on `stream`) */
int coattn_status_clear(void* workspace, void* stream);
/* copy the status block to `host_words` (COATTN_STATUS_WORDS entries) and wait for `stream` */
int coattn_status_read(const void* workspace, uint32_t* host_words, void* stream);

/* ABI version of the loaded library (== COATTN_B200_ABI_VERSION it was built with). */
int coattn_b200_abi_version(void);

/* Static, human readable description of a return code of this library. */
const char* coattn_b200_strerror(int code);

/* Bytes of scratch `coattn_forward` (and coattn_forward_queries / coattn_forward16 with n = nq * refs) needs for a batch of
 * n pairs of [c, h, w] features: six 16-bit operand planes of n * round_up(h w, 256) * c elements, W in 16 bits, and the
 * z / lse segments -- one copy per key-range part COATTN_FLAG_SPLIT_KEYS can use at this size (1 once the batch fills the
 * GPU).  Independent of the flags. */
int64_t coattn_workspace_bytes(int n, int c, int h, int w);

/*
 * Whole hot path for one modality (replaces :150-187 or :204-238):
 *   prep (16-bit cast/transposes) -> project (Q = W A) -> attend (both softmax axes, gate + sigmoid + scale
 *   fused into its drain) -> passthrough copy of the original features into the second half of the concat.
 * z, lse and mask are optional outputs kept for the backward pass; each may be NULL (lse and, with
 * COATTN_FLAG_UNFUSED_GATE, z are then routed to the workspace).
 */
int coattn_forward(const float* v_a, const float* v_b, const float* w, const float* gate_w,
                   const float* gate_b, float* cat_a, float* cat_b, float* z, float* lse, float* mask,
                   void* workspace, int64_t workspace_bytes, int n, int c, int h, int w_,
                   unsigned flags, void* stream);

/*
 * test.py-style inference (test.py:287-305): `nq` query (target) frames, each co-attended with `refs` reference frames;
 * pair p = (query p / refs, reference p), n = nq * refs pairs, frame-A outputs only (what test.py:301 keeps).
 *   v_a [nq, 256, H, W] query features, v_b [nq * refs, 256, H, W] reference features,
 *   cat_a [nq * refs, 512 (256 with COATTN_FLAG_GATED_ONLY), H, W]; workspace as for n = nq * refs pairs.
 * The reference re-encodes and re-projects the query for every pair; here its 16-bit cast and Q = W V_a are computed
 * once per query.  Same results as coattn_forward(COATTN_FLAG_A_ONLY) on the query features repeated `refs` times.
 */
int coattn_forward_queries(const float* v_a, const float* v_b, const float* w, const float* gate_w,
                           const float* gate_b, float* cat_a, void* workspace, int64_t workspace_bytes, int nq,
                           int refs, int c, int h, int w_, unsigned flags, void* stream);

/*
 * 16-bit feature interface ("V_a, V_b fp32 (or bf16) ... out cat_a, cat_b (same dtype as input)";
 * 8(f) N4: the producer -- deeplab/deeplabv3_encoder.py:80-82 under autocast -- hands over 16-bit features).
 *   v_a   [nq, 256, H, W]         16-bit: IEEE half, or bfloat16 with COATTN_FLAG_BF16 (= the operand format)
 *   v_b   [nq * refs, 256, H, W]  16-bit; pair p = (query frame p / refs, v_b[p]); refs == 1: ordinary pair batch
 *   cat_a, cat_b [nq * refs, 512 (256 with COATTN_FLAG_GATED_ONLY), H, W] 16-bit, same format; cat_b may be NULL with
 *         COATTN_FLAG_A_ONLY, which refs > 1 implies (test.py:301 keeps the frame-A output only)
 *   lse [passes, n, L], mask [passes, n, L] fp32, optional (NULL: not kept)
 *   w, gate_w, gate_b stay fp32 (the module's parameters).  Flags: COATTN_FLAG_BF16 | _A_ONLY | _GATED_ONLY, others
 *   -> COATTN_E_UNSUPPORTED.  Workspace as for coattn_forward with n = nq * refs.
 * With H*W % 8 == 0 and 16-byte aligned v_a / v_b the tensor cores' operands are read by TMA straight from the caller's
 * tensors (no cast, no copy; only Q = W V_a is written to the workspace); otherwise the features are first copied into
 * zero-padded planes.  The arithmetic is that of coattn_forward on the same 16-bit values; outputs are rounded to
 * 16 bits once, at the store, and the passthrough half of the concat is a bit copy of the inputs.
 */
int coattn_forward16(const void* v_a, const void* v_b, const float* w, const float* gate_w, const float* gate_b,
                     void* cat_a, void* cat_b, float* lse, float* mask, void* workspace, int64_t workspace_bytes,
                     int nq, int refs, int c, int h, int w_, unsigned flags, void* stream);

/*
 * Producer side: the tail of the ASPP encoder head, deeplab/deeplabv3_encoder.py:80-82,
 *     features = PReLU(BatchNorm(bottleneck(x)))         in eval mode (BN with running statistics),
 * fused with the 16-bit operand cast of the co-attention that consumes the features:
 *   x      [N, 256, H, W]  output of the 3x3 bottleneck conv (bias included)
 *   scale, shift [256]     eval-mode BN as an affine map: gamma / sqrt(var + eps), beta - mean * scale
 *   slope  [1]             the PReLU parameter
 *   y      [N, 256, H, W]  fp32 features (what the reference's encoder returns), or NULL if nobody needs them
 *   frame  0: these are frame A's features (V_a), 1: frame B's (V_b)
 * and writes the zero-padded 16-bit plane of that frame into `workspace` (laid out for n pairs, as coattn_forward
 * expects); a following coattn_forward(..., COATTN_FLAG_PLANES_READY) with the same workspace, n and operand format
 * starts at the projection.  One HBM pass instead of four (BN, PReLU, and the cast's read + write).
 */
int coattn_stage_tail(const float* x, const float* scale, const float* shift, const float* slope, float* y,
                      void* workspace, int64_t workspace_bytes, int frame, int n, int c, int h, int w_,
                      unsigned flags, void* stream);

/* ---- the four stages, exported individually for unit parity tests and per-kernel timing ---- */

/* stage 1 (:154-158): 16-bit operands.  Fills the workspace segments At, Bt, A16, B16 and W16. */
int coattn_stage_prep(const float* v_a, const float* v_b, const float* w, void* workspace,
                      int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags,
                      void* stream);
/* stages 1+2 of coattn_forward.  Default: cast both frames to 16 bit (channel-major, no transposes: B16, A16) and
 * project Q16 = W A16 with the A16 tile as an MN-major tensor-core operand.  With COATTN_FLAG_KMAJOR: V_b through the
 * transposing prep kernel, V_a through the fused convert + projection kernel (fills Bt, B16, A16, W16, Qt). */
int coattn_stage_prep_project(const float* v_a, const float* v_b, const float* w, void* workspace,
                              int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags,
                              void* stream);
/* stage 2 (:159): Qt = At W^T on the tensor cores. */
int coattn_stage_project(void* workspace, int64_t workspace_bytes, int n, int c, int h, int w_,
                         unsigned flags, void* stream);
/* stage 3 (:160-170): fused affinity / dual softmax / attend.  Writes z and lse. */
int coattn_stage_attend(float* z, float* lse, void* workspace, int64_t workspace_bytes, int n,
                        int c, int h, int w_, unsigned flags, void* stream);
/* stages 3+4 fused (:160-187): as stage 3, and the drain also writes Z * sigmoid(gate(Z)) into channels
 * [0, 256) of cat_a / cat_b and the gate values into mask.  If v_a / v_b are given (both or neither), the
 * passthrough half cat_x[:, 256:512] = v_x is written too (by a spare warp of the attend kernel; by the
 * passthrough kernel under COATTN_FLAG_SINGLE_CTA).  z, lse (-> workspace) and mask may be NULL. */
int coattn_stage_attend_gate(const float* v_a, const float* v_b, float* cat_a, float* cat_b, float* z,
                             float* lse, float* mask, const float* gate_w, const float* gate_b,
                             void* workspace, int64_t workspace_bytes, int n, int c, int h, int w_,
                             unsigned flags, void* stream);
/* stage 4b (:186-187): cat_x[:, 256:512] = v_x, the passthrough half of the concat. */
int coattn_stage_passthrough(const float* v_a, const float* v_b, float* cat_a, float* cat_b, int n,
                             int c, int h, int w_, void* stream);
/* stage 4 (:177-187), stand-alone: gate, sigmoid, scale, concat (both halves). */
int coattn_stage_gate(const float* z, const float* v_a, const float* v_b, const float* gate_w,
                      const float* gate_b, float* cat_a, float* cat_b, int n, int c, int h, int w_,
                      void* stream);

/*
 * Backward of `coattn_forward` for one modality: what autograd computes through :158-187 (train.py:599), with
 * the reference's semantics -- the B-side gate mask is a constant (:178-182) and V_b is a constant
 * (no_grad_for_counterpart, :144-148).  S is recomputed from the 16-bit operands and the saved `lse`; the
 * softmax matrices of the forward pass are not stored.
 *
 *   inputs   v_a, v_b, w, gate_w           as in coattn_forward
 *            z [2,N,256,L], lse [2,N,L], mask [2,N,L]      saved outputs of coattn_forward
 *            d_cat_a, d_cat_b [N,512,H,W]  gradients w.r.t. the two concat tensors; d_cat_b may be NULL
 *                                          (depth modality: the B branch is gradient dead, :240-247).  With
 *                                          COATTN_FLAG_GATED_ONLY they are [N,256,H,W], the gradients of the gated halves a
 *                                          gated-only forward returned (no passthrough term)
 *   outputs  d_v_a [N,256,H,W], d_w [256,256] (16-byte aligned, else COATTN_E_ALIGN), d_gate_w [256], d_gate_b [1]
 *            (may be NULL); all overwritten
 *            d_v_b [N,256,H,W] or NULL: gradient for the counterpart frame, only needed with
 *            no_grad_for_counterpart=False (:147-148); costs a second flash sweep with the roles of the frames swapped
 *            and a bf16 projection.  `counterpart` in the size query is kept for ABI stability and ignored.
 * Nothing of size L x L exists in the workspace: S, dP_a and dP_b are recomputed per 256 x 128 tile in TMEM
 * (csrc/bwd_flash_kernel.cuh); the scratch is 16-bit planes [N][256][round_up(L, 256)] and a few vectors, linear in L.
 *
 * A backward workspace BEGINS with the forward layout.  A caller that ran coattn_forward of the same call (same n, shape,
 * operand format, features) on this very buffer -- sized by coattn_backward_workspace_bytes -- and has not let anything
 * else write to it since may pass COATTN_FLAG_PLANES_READY: the 16-bit planes of V_a and V_b are then taken as they are
 * and the feature cast (the forward's operands regenerated from v_a / v_b) is skipped; results are bit-identical.  Without
 * the flag the backward assumes nothing about the workspace's content.
 */
int64_t coattn_backward_workspace_bytes(int n, int c, int h, int w, int counterpart);
int coattn_backward(const float* v_a, const float* v_b, const float* w, const float* gate_w, const float* z,
                    const float* lse, const float* mask, const float* d_cat_a, const float* d_cat_b,
                    float* d_v_a, float* d_v_b, float* d_w, float* d_gate_w, float* d_gate_b,
                    void* workspace, int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags,
                    void* stream);

/*
 * Debug/test view of the workspace: byte offset and byte size of a named segment
 * ("at", "bt", "qt", "a16", "b16", "w16", "z", "lse") for the given problem size, so tests can
 * compare intermediate operands with the oracle.  Returns 0 or COATTN_E_NULL for an unknown name.
 */
int coattn_workspace_segment(const char* name, int n, int c, int h, int w_, int64_t* offset,
                             int64_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* COATTN_B200_H_ */
