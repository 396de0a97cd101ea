"""test.py-style inference (reference test.py:278-305): every query (target) frame is co-attended with
`sample_range` reference frames of its sequence and the frame-A prediction x1 is averaged over them.

The reference runs the full model once per (query, reference) pair, so the query frame is re-encoded
`sample_range` times and the B-side head is computed although only `output[0]` is kept (test.py:293, :301).
Here the query is encoded ONCE, the references of all queries go through the encoders as one batch, the
co-attention runs as one batched launch computing the frame-A outputs only (half the attend work), with the query
side of the operator (16-bit cast, Q = W V_a) prepared once per query (`coattn_forward_queries`), and only the A-side
head is evaluated.  Queries are the unit that is sharded across GPUs
(`pair_batcher.shard_range`), so the mean over references never crosses a device.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .coattention import (coattention_forward16_raw, coattention_forward_raw, coattention_queries_raw,
                          modality_overlap_pays, run_modalities)


@torch.no_grad()
def coattention_frame_a(v_a, v_b, weight, gate_weight, gate_bias=None):
    """cat_a only ([N, 2C, H, W]) for N (query, reference) feature pairs."""
    cat_a, _, _, _ = coattention_forward_raw(v_a, v_b, weight, gate_weight, gate_bias, want_z=False, a_only=True)
    return cat_a


@torch.no_grad()
def coattention_queries_frame_a(v_a, v_b, weight, gate_weight, gate_bias=None, refs: int = 1):
    """cat_a only for Q query feature maps v_a [Q, C, H, W], each paired with `refs` consecutive reference maps of
    v_b [Q * refs, C, H, W]; the query side of the operator is prepared once per query (coattn_forward_queries)."""
    if v_a.dtype in (torch.float16, torch.bfloat16):      # half-precision model: 16-bit features in and out
        return coattention_forward16_raw(v_a, v_b, weight, gate_weight, gate_bias, refs=refs, a_only=True)[0]
    return coattention_queries_raw(v_a, v_b, weight, gate_weight, gate_bias, refs=refs)


@torch.no_grad()
def segment_with_references(model, target_rgb, target_depth, ref_rgbs, ref_depths):
    """Averaged frame-A mask of each query over its reference frames.

    target_rgb [Q,3,H,W], target_depth [Q,1,H,W], ref_rgbs [Q,R,3,H,W], ref_depths [Q,R,1,H,W]
    returns    [Q,1,H,W]  = mean_r model(target_q, ref_{q,r}, ...)[0]        (test.py:287-305)
    """
    q, r = ref_rgbs.shape[:2]
    size = target_rgb.shape[2:]
    v_a, _ = model.encoder(target_rgb)                        # once per query
    d_a = model.depth_encoder(target_depth)
    v_b, _ = model.encoder(ref_rgbs.flatten(0, 1))            # all references of all queries in one batch
    d_b = model.depth_encoder(ref_depths.flatten(0, 1))
    # the two modality calls are independent: when the batch leaves the last wave of the attend kernel partly empty they
    # run on two streams and fill each other's holes (same results, bit for bit)
    cat, dcat = run_modalities(
        lambda: coattention_queries_frame_a(v_a, v_b, model.rgb_similarity_weights.weight, model.gate.weight, None, refs=r),
        lambda: coattention_queries_frame_a(d_a, d_b, model.depth_similarity_weights.weight, model.depth_gate.weight,
                                            model.depth_gate.bias, refs=r),
        (d_a, d_b), overlap=modality_overlap_pays(q * r, v_a.shape[2], v_a.shape[3], passes=1, device=v_a.device))
    z = model.bn_A(model.reduce_channels_A(cat))                                             # :188, :190
    z = model.prelu(z + model.depth_weights(model.depth_bn(model.depth_reduce_channels(dcat))))  # :239-256
    x1 = torch.sigmoid(F.interpolate(model.segmentation_classifier_A(z), size, mode="bilinear"))  # :260-265
    return x1.view(q, r, *x1.shape[1:]).mean(dim=1)
