"""Drop-in `RGBDSegmentation_RAA` (ResNet + ASPP + Add, RGB + depth siamese model).

Same class name, constructor, `forward(rgbs_a, rgbs_b, depths_a, depths_b) -> (x1, x2, labels)`, `get_params`,
`load_state` and state_dict keys as the reference (rgbd_segmentation_RAA.py:18-268), so config.yaml, train.py and
test.py keep working; the two inline co-attention blocks (:150-187 and :204-238) are replaced by one call each
into the sm_100a kernels (`cosnet_b200.coattention`).  Everything else (encoders, 3x3 reduce convs, BN,
classifiers, upsampling) stays on cuDNN / ATen.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .backbone import DepthEncoder_ResNetASPP, Encoder, init_reference_style
from .coattention import (bn_eval_affine, coattention, coattention_planes_ready, encoder_tail, modality_overlap_pays,
                          run_modalities)

# legacy checkpoint prefixes -> current names (rgbd_segmentation_RAA.py:114-133); first match wins,
# "encoder.main_classifier" must be tested before the generic "encoder." rule
_LEGACY_PREFIXES = (
    ("encoder.layer5.", "encoder.aspp."),
    ("encoder.main_classifier", "encoder.main_classifier"),
    ("encoder.", "encoder.backbone."),
    ("linear_e.", "rgb_similarity_weights."),
    ("conv1.", "reduce_channels_A."),
    ("conv2.", "reduce_channels_B."),
    ("bn1.", "bn_A."),
    ("bn2.", "bn_B."),
    ("main_classifier1.", "segmentation_classifier_A."),
    ("main_classifier2.", "segmentation_classifier_B."),
)

_PARAM_SUBSETS = {
    # subset -> groups it contains (rgbd_segmentation_RAA.py:70-75)
    "none": (),
    "encoder": ("encoder",),
    "rgb_attention": ("rgb_attention",),
    "rgb": ("encoder", "rgb_attention"),
    "depth": ("depth",),
    "decoder": ("decoder",),
    "all": ("encoder", "rgb_attention", "depth", "decoder"),
}


class RGBDSegmentation_RAA(nn.Module):
    def __init__(self, block, num_blocks_of_layers_4_rgb, num_blocks_of_layers_4_depth, num_classes,
                 all_channel=256, all_dim=60 * 60, no_grad_for_counterpart=True):
        super().__init__()
        c = all_channel
        # RGB branch (:26-34)
        self.encoder = Encoder(3, block, num_blocks_of_layers_4_rgb, num_classes)
        self.rgb_similarity_weights = nn.Linear(c, c, bias=False)
        self.gate = nn.Conv2d(c, 1, kernel_size=1, bias=False)
        self.gate_s = nn.Sigmoid()
        self.reduce_channels_A = nn.Conv2d(2 * c, c, kernel_size=3, padding=1, bias=False)
        self.reduce_channels_B = nn.Conv2d(2 * c, c, kernel_size=3, padding=1, bias=False)
        self.bn_A = nn.BatchNorm2d(c)
        self.bn_B = nn.BatchNorm2d(c)
        self.prelu = nn.ReLU(inplace=True)
        # depth branch (:37-43)
        self.depth_encoder = DepthEncoder_ResNetASPP(256, block, num_blocks_of_layers_4_depth, num_classes)
        self.depth_similarity_weights = nn.Linear(c, c, bias=False)
        self.depth_gate = nn.Conv2d(c, 1, kernel_size=1, bias=True)
        self.depth_gate_s = nn.Sigmoid()
        self.depth_reduce_channels = nn.Conv2d(2 * c, c, kernel_size=3, padding=1, bias=False)
        self.depth_bn = nn.BatchNorm2d(c)
        self.depth_weights = nn.Conv2d(c, c, kernel_size=1, bias=True)
        # decoder (:47-49)
        self.segmentation_classifier_A = nn.Conv2d(c, num_classes, kernel_size=1, bias=True)
        self.segmentation_classifier_B = nn.Conv2d(c, num_classes, kernel_size=1, bias=True)
        self.softmax = nn.Sigmoid()

        self.no_grad_for_counterpart = no_grad_for_counterpart
        # `all_dim` is accepted for signature compatibility and ignored, like in the reference (:153)
        init_reference_style(self)   # :53-62 (nn.Linear keeps its default init)
        # the operator that replaces :150-187 / :204-238; tests may swap it for the CPU oracle
        self.coattention_impl = coattention
        # True: never materialise the concat -- conv(cat([Zg, V]), W) is evaluated as conv(Zg, W[:, :C]) + conv(V, W[:, C:])
        # (same parameters and state_dict; fp32 summation order differs by ~1e-6).  SURVEY.md 8f, row N3.
        self.split_reduce_conv = False
        # eval-mode forward under no_grad: run the RGB and the depth co-attention on two CUDA streams when that needs fewer
        # waves of the attend kernel (None = decide from the batch shape, True / False = force).  Bit-identical outputs.
        self.overlap_modalities = None
        # eval-mode forward under no_grad: the encoder tails (BN + PReLU) are fused with the
        # co-attention's 16-bit operand cast into one kernel each, the operators return the gated halves only, and the
        # reduce convs run in two halves with their BatchNorm (and the depth branch's 1x1 weighting) folded into the conv
        # weights -- no concat, no cast kernel, no stand-alone BN / PReLU launches around the hot path.  Same parameters
        # and state_dict; outputs equal the plain path to fp32 rounding (~1e-6).
        self.fuse_eval_path = True
        self._folded = {}

    # ------------------------------------------------------------------ optimiser groups (:65-100)
    def get_params(self, subset="none"):
        groups = {
            "encoder": [self.encoder],
            "rgb_attention": [self.rgb_similarity_weights, self.gate, self.reduce_channels_A, self.reduce_channels_B,
                              self.bn_A, self.bn_B],
            "depth": self.depth_encoder.get_params() + [self.depth_gate, self.depth_similarity_weights,
                                                        self.depth_reduce_channels, self.depth_bn, self.depth_weights],
            "decoder": [self.segmentation_classifier_A, self.segmentation_classifier_B],
        }
        out = []
        for g in _PARAM_SUBSETS.get(subset, ()):
            out.extend(groups[g])
        return out

    # ------------------------------------------------------------------ checkpoints (:103-136)
    @staticmethod
    def _current_key(key: str) -> str:
        if key.startswith("module."):      # saved from nn.DataParallel
            key = key[len("module."):]
        for old, new in _LEGACY_PREFIXES:
            if key.startswith(old):
                return new + key[len(old):]
        return key

    def load_state(self, state_dict):
        merged = self.state_dict().copy()
        for key, value in state_dict.items():
            merged[self._current_key(key)] = value
        self.load_state_dict(merged)

    # ------------------------------------------------------------------ forward (:139-268)
    def _encode_pair(self, enc, x_a, x_b, returns_tuple):
        """Frame A with grad, frame B under no_grad when `no_grad_for_counterpart` (:143-148, :198-203)."""
        out_a = enc(x_a)
        if self.no_grad_for_counterpart:
            with torch.no_grad():
                out_b = enc(x_b)
        else:
            out_b = enc(x_b)
        if returns_tuple:
            return out_a[0], out_b[0], out_b[1]    # the auxiliary map that survives is frame B's (:143, :146/148)
        return out_a, out_b, None

    @staticmethod
    def _split_conv(conv, gated, passthrough):
        """conv(cat([gated, passthrough], 1)) without the concat: the 3x3 weight [C, 2C, 3, 3] is applied in two halves."""
        c = gated.shape[1]
        w = conv.weight
        return (F.conv2d(gated, w[:, :c], None, conv.stride, conv.padding, conv.dilation) +
                F.conv2d(passthrough, w[:, c:], conv.bias, conv.stride, conv.padding, conv.dilation))

    def _folded_reduce(self, name, conv, bn, pointwise=None):
        """(weight [C, 2C, 3, 3], bias [C]) of conv -> BN_eval (-> 1x1 `pointwise` conv) as ONE convolution.  Cached until a
        parameter or a running statistic changes."""
        deps = [conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var] + (
            [pointwise.weight, pointwise.bias] if pointwise is not None else [])
        key = tuple((t.data_ptr(), t._version) for t in deps)
        hit = self._folded.get(name)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        with torch.no_grad():
            scale, shift = bn_eval_affine(bn)
            w = conv.weight * scale.view(-1, 1, 1, 1)
            b = shift if conv.bias is None else shift + conv.bias * scale
            if pointwise is not None:       # y = P (conv(x) * scale + shift) + p  is linear in conv's weights
                pw = pointwise.weight.view(pointwise.out_channels, pointwise.in_channels)
                w = torch.einsum("om,mikl->oikl", pw, w).contiguous()
                b = pw @ b + pointwise.bias
        self._folded[name] = (key, w, b)
        return w, b

    @staticmethod
    def _split_conv_folded(w, b, conv, gated, passthrough):
        c = gated.shape[1]
        return (F.conv2d(gated, w[:, :c], None, conv.stride, conv.padding, conv.dilation) +
                F.conv2d(passthrough, w[:, c:], b, conv.stride, conv.padding, conv.dilation))

    def _encode_pair_fused(self, enc, x_a, x_b, returns_tuple, tag):
        """Eval-mode encoders whose ASPP tail runs in the fused kernel: returns (V_a, V_b, labels) like `_encode_pair`, and
        leaves the 16-bit operand planes of both frames in the workspace tagged `tag`."""
        aspp = enc.aspp
        scale, shift = bn_eval_affine(aspp.bn)
        v_a = encoder_tail(aspp.pre_tail(enc.backbone(x_a)), scale, shift, aspp.prelu.weight, 0, tag)
        v_b = encoder_tail(aspp.pre_tail(enc.backbone(x_b)), scale, shift, aspp.prelu.weight, 1, tag)
        labels = None
        if returns_tuple:       # the auxiliary map that survives is frame B's (:143, :146/148)
            labels = enc.softmax(F.interpolate(enc.main_classifier(v_b), size=x_b.shape[2:], mode="bilinear"))
        return v_a, v_b, labels

    def _forward_eval_fused(self, rgbs_a, rgbs_b, depths_a, depths_b):
        input_size = rgbs_a.shape[2:]
        v_a, v_b, labels = self._encode_pair_fused(self.encoder, rgbs_a, rgbs_b, True, "rgb")
        d_a, d_b, _ = self._encode_pair_fused(self.depth_encoder, depths_a, depths_b, False, "depth")
        r_a, r_b = coattention_planes_ready(v_a, v_b, self.rgb_similarity_weights.weight, self.gate.weight, None, "rgb",
                                            gated_only=True)
        q_a, q_b = coattention_planes_ready(d_a, d_b, self.depth_similarity_weights.weight, self.depth_gate.weight,
                                            self.depth_gate.bias, "depth", gated_only=True)
        wa, ba = self._folded_reduce("A", self.reduce_channels_A, self.bn_A)
        wb, bb = self._folded_reduce("B", self.reduce_channels_B, self.bn_B)
        wd, bd = self._folded_reduce("D", self.depth_reduce_channels, self.depth_bn, self.depth_weights)
        z_a = self._split_conv_folded(wa, ba, self.reduce_channels_A, r_a, v_a)            # :188, :190
        z_b = self._split_conv_folded(wb, bb, self.reduce_channels_B, r_b, v_b)            # :189, :191
        dz_a = self._split_conv_folded(wd, bd, self.depth_reduce_channels, q_a, d_a)       # :239, :242, :245
        dz_b = self._split_conv_folded(wd, bd, self.depth_reduce_channels, q_b, d_b)       # :240-247
        z_a = self.prelu(z_a + dz_a)                            # :251, :256
        z_b = self.prelu(z_b + dz_b)                            # :252, :257
        x1 = self.softmax(F.interpolate(self.segmentation_classifier_A(z_a), input_size, mode="bilinear"))  # :260-265
        x2 = self.softmax(F.interpolate(self.segmentation_classifier_B(z_b), input_size, mode="bilinear"))
        return x1, x2, labels

    def _forward_eval(self, rgbs_a, rgbs_b, depths_a, depths_b):
        """Inference (eval mode, no autograd): the same operators as `forward`, with both encoders evaluated first so that
        the two co-attention calls -- independent of each other -- can share the GPU (`run_modalities`)."""
        if (self.fuse_eval_path and rgbs_a.dtype == torch.float32 and hasattr(self.encoder, "aspp")
                and hasattr(self.depth_encoder, "aspp") and self.encoder.aspp.prelu.weight.numel() == 1
                and self.depth_encoder.aspp.prelu.weight.numel() == 1):
            return self._forward_eval_fused(rgbs_a, rgbs_b, depths_a, depths_b)
        input_size = rgbs_a.shape[2:]
        v_a, v_b, labels = self._encode_pair(self.encoder, rgbs_a, rgbs_b, True)
        d_a, d_b, _ = self._encode_pair(self.depth_encoder, depths_a, depths_b, False)
        split = self.split_reduce_conv
        overlap = self.overlap_modalities
        if overlap is None:
            overlap = modality_overlap_pays(v_a.shape[0], v_a.shape[2], v_a.shape[3], 2, v_a.device)
        (r_a, r_b), (q_a, q_b) = run_modalities(
            lambda: coattention(v_a, v_b, self.rgb_similarity_weights.weight, self.gate.weight, None, gated_only=split),
            lambda: coattention(d_a, d_b, self.depth_similarity_weights.weight, self.depth_gate.weight, self.depth_gate.bias,
                                gated_only=split),
            (d_a, d_b), overlap)
        if split:
            z_a = self.bn_A(self._split_conv(self.reduce_channels_A, r_a, v_a))
            z_b = self.bn_B(self._split_conv(self.reduce_channels_B, r_b, v_b))
            dz_a = self.depth_weights(self.depth_bn(self._split_conv(self.depth_reduce_channels, q_a, d_a)))
            dz_b = self.depth_weights(self.depth_bn(self._split_conv(self.depth_reduce_channels, q_b, d_b)))
        else:
            z_a = self.bn_A(self.reduce_channels_A(r_a))            # :188, :190
            z_b = self.bn_B(self.reduce_channels_B(r_b))            # :189, :191
            dz_a = self.depth_weights(self.depth_bn(self.depth_reduce_channels(q_a)))       # :239, :242, :245
            dz_b = self.depth_weights(self.depth_bn(self.depth_reduce_channels(q_b)))       # :240-247
        z_a = self.prelu(z_a + dz_a)                            # :251, :256
        z_b = self.prelu(z_b + dz_b)                            # :252, :257
        x1 = self.softmax(F.interpolate(self.segmentation_classifier_A(z_a), input_size, mode="bilinear"))  # :260-265
        x2 = self.softmax(F.interpolate(self.segmentation_classifier_B(z_b), input_size, mode="bilinear"))
        return x1, x2, labels

    def forward(self, rgbs_a, rgbs_b, depths_a, depths_b):
        if (not self.training and not torch.is_grad_enabled() and rgbs_a.is_cuda and self.coattention_impl is coattention
                and self.overlap_modalities is not False):
            return self._forward_eval(rgbs_a, rgbs_b, depths_a, depths_b)
        input_size = rgbs_a.shape[2:]

        v_a, v_b, labels = self._encode_pair(self.encoder, rgbs_a, rgbs_b, True)
        if self.split_reduce_conv:
            g_a, g_b = self.coattention_impl(v_a, v_b, self.rgb_similarity_weights.weight, self.gate.weight, None,
                                             gated_only=True)
            z_a = self.bn_A(self._split_conv(self.reduce_channels_A, g_a, v_a))
            z_b = self.bn_B(self._split_conv(self.reduce_channels_B, g_b, v_b))
            del g_a, g_b
        else:
            cat_a, cat_b = self.coattention_impl(v_a, v_b, self.rgb_similarity_weights.weight, self.gate.weight, None)
            z_a = self.bn_A(self.reduce_channels_A(cat_a))          # :188, :190
            z_b = self.bn_B(self.reduce_channels_B(cat_b))          # :189, :191
            del cat_a, cat_b
        del v_a, v_b

        d_a, d_b, _ = self._encode_pair(self.depth_encoder, depths_a, depths_b, False)
        if self.split_reduce_conv:
            g_a, g_b = self.coattention_impl(d_a, d_b, self.depth_similarity_weights.weight, self.depth_gate.weight,
                                             self.depth_gate.bias, gated_only=True)
            dz_a = self.depth_weights(self.depth_bn(self._split_conv(self.depth_reduce_channels, g_a, d_a)))
            with torch.no_grad():
                dz_b = self.depth_weights(self.depth_bn(self._split_conv(self.depth_reduce_channels, g_b, d_b)))
            del g_a, g_b
        else:
            dcat_a, dcat_b = self.coattention_impl(d_a, d_b, self.depth_similarity_weights.weight, self.depth_gate.weight,
                                                   self.depth_gate.bias)
            dz_a = self.depth_weights(self.depth_bn(self.depth_reduce_channels(dcat_a)))       # :239, :242, :245
            with torch.no_grad():                                                               # :240-247
                dz_b = self.depth_weights(self.depth_bn(self.depth_reduce_channels(dcat_b)))
            del dcat_a, dcat_b
        del d_a, d_b

        z_a = self.prelu(z_a + dz_a)                            # :251, :256
        z_b = self.prelu(z_b + dz_b)                            # :252, :257
        x1 = self.softmax(F.interpolate(self.segmentation_classifier_A(z_a), input_size, mode="bilinear"))  # :260-265
        x2 = self.softmax(F.interpolate(self.segmentation_classifier_B(z_b), input_size, mode="bilinear"))
        return x1, x2, labels
